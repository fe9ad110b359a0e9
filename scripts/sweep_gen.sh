export NLML_HPE_LIB=build/dev/libnlml_tune.so
S="timeout 120 python scripts/time_gen.py"
$S 8 5 5 5 1404 40
NLML_GEN_YTM=1 NLML_GEN_BCP=6 $S 8 5 5 5 1404 40
$S 5 3 3 3 1404 100
$S 16 8 8 8 96 3
S="timeout 120 python scripts/sweep_gen.py"
for bcp in 2 3 4 5 6 8; do NLML_GEN_BCP=$bcp $S 8 5 5 5 1404 60; done
for bcp in 3 6 9 12; do NLML_GEN_BCP=$bcp $S 5 3 3 3 1404 300; done
$S 8 8 8 8 96 20
NLML_GEN_BCP=2 $S 8 8 8 8 96 20
