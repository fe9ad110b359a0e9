export NLML_HPE_LIB=build/dev/libnlml_tune.so
S="timeout 120 python scripts/time_gen.py"
$S 8 5 5 5 1404 40
$S 16 8 8 8 96 3
S="timeout 120 python scripts/sweep_gen.py"
for bcp in 3 4 5 6 7 8 10 12 16; do NLML_GEN_BCP=$bcp $S 8 5 5 5 1404 60; done
for bcp in 1 2; do NLML_GEN_BCP=$bcp $S 16 8 8 8 96 4; done
