# round-2 ncu captures (run under gpurun; one ncu family per call: all ncu runs in a call count as one)
set -x
# 1. MLP chain: launch list with DRAM bytes of one 151552-sample chunk
python scripts/prof_run.py mlp --n 151552 > gpurun_out/r2_mlp_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"linear_tc|split_planes" -c 27 --csv --log-file gpurun_out/r2_mlp_launches.csv python scripts/prof_run.py mlp --n 151552 > gpurun_out/r2_ncu_mlp.log 2>&1
# 2. (5,3,3,3) tensor-core fit kernel, full set (T=300 keeps the replays short)
python scripts/prof_run.py tucker --n 37888 --iters 300 --kernel tensor_core > gpurun_out/r2_tc_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tucker_fit_tc_kernel -s 1 -c 1 -f -o gpurun_out/r02_tucker_tc_full python scripts/prof_run.py tucker --n 37888 --iters 300 --kernel tensor_core > gpurun_out/r2_ncu_tc.log 2>&1
# 3. run-time-rank kernel on the (8,5,5,5) core
python scripts/sweep_gen.py 8 5 5 5 1404 30 > gpurun_out/r2_gen_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tucker_fit_gen_kernel -s 1 -c 1 -f -o gpurun_out/r02_tucker_gen_full python scripts/sweep_gen.py 8 5 5 5 1404 30 > gpurun_out/r2_ncu_gen.log 2>&1
# 4. projection GEMM + converged solve
python scripts/prof_run.py solve --n 303104 > gpurun_out/r2_solve_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"tucker_project_tc_kernel|tucker_fit_tps_kernel" -s 2 -c 2 -f -o gpurun_out/r02_tucker_solve_full python scripts/prof_run.py solve --n 303104 > gpurun_out/r2_ncu_solve.log 2>&1
ls -la gpurun_out/*.ncu-rep
