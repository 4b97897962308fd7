#!/usr/bin/env bash
# Second profiling pass of round 2 (dense path after the TMA-store epilogues and the fused backward): launch list of
# the eager C2 step, full captures of the two fused kernels.  Run on the GPU box through gpurun; outputs in gpurun_out/.
set -u
O=gpurun_out
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -c 120 --csv --log-file $O/r2b_launches_c2.csv python benchmarks/ncu_step.py c2 2 > $O/ncu_r2b_c2.log 2>&1
$NCU --metrics gpu__time_duration.sum -c 200 --csv --log-file $O/r2b_launches_c3.csv python benchmarks/ncu_step.py c3 1 > $O/ncu_r2b_c3.log 2>&1
$NCU --set full --import-source on -k "regex:k_dense_fwd_fused_ts|k_dense_bwd_fused" -c 2 --launch-skip 2 -o $O/r2b_fused_c2 -f python benchmarks/ncu_step.py c2 2 > $O/ncu_r2b_full.log 2>&1
ncu -i $O/r2b_fused_c2.ncu-rep --page raw --csv > $O/r2b_fused_c2.raw.csv 2>/dev/null
tail -n 2 $O/ncu_r2b_full.log; du -sh $O
