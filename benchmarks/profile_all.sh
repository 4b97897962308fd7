#!/usr/bin/env bash
# Round-2 profiling pass (run on the GPU box through gpurun): launch lists of every workload, full captures of the
# dominant kernels, a racecheck + memcheck pass over the new sparse kernels.  Outputs land in gpurun_out/.
set -u
O=gpurun_out
NCU="ncu --clock-control none"
for w in c2 c3 c1 c4 c5; do
  $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/r2_launches_${w}.csv python benchmarks/ncu_step.py $w 2 > $O/ncu_${w}.log 2>&1
done
# the driver's own command under ncu (graph replays are profiled node by node; capped)
$NCU --metrics gpu__time_duration.sum -c 300 --csv --log-file $O/r2_launches_bench_default.csv python bench.py --workload c2 --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_bench.log 2>&1
$NCU --set full --import-source on -k regex:k_dense_fwd_fused_ts -c 1 --launch-skip 1 -o $O/r2_fused_ts_c2 -f python benchmarks/ncu_step.py c2 2 > $O/ncu_full1.log 2>&1
$NCU --set full --import-source on -k "regex:k_tc_gemm|k_graph" -c 8 --launch-skip 8 -o $O/r2_c2_bwd -f python benchmarks/ncu_step.py c2 2 > $O/ncu_full2.log 2>&1
$NCU --set full --import-source on -k "regex:k_bucket_tiles|k_compact_emit|k_fine_row_spans|k_segment_reduce" -c 5 --launch-skip 5 -o $O/r2_c4_kernels -f python benchmarks/ncu_step.py c4 2 > $O/ncu_full3.log 2>&1
$NCU --set full --import-source on -k "regex:k_compact_onepass|k_dsum_down|k_dsum_reduce|k_segment_reduce" -c 8 --launch-skip 8 -o $O/r2_c5_kernels -f python benchmarks/ncu_step.py c5 2 > $O/ncu_full4.log 2>&1
$NCU --set full --import-source on -k "regex:k_tc_gemm|k_dense_fwd_fused" -c 12 -o $O/r2_c3_kernels -f python benchmarks/ncu_step.py c3 1 > $O/ncu_full5.log 2>&1
for r in r2_fused_ts_c2 r2_c2_bwd r2_c4_kernels r2_c5_kernels r2_c3_kernels; do
  ncu -i $O/$r.ncu-rep --page raw --csv > $O/$r.raw.csv 2>/dev/null
done
# gpurun brings back at most 64 MiB: keep the raw pages, drop the large reports (the dominant kernel's report stays)
rm -f $O/r2_c2_bwd.ncu-rep $O/r2_c4_kernels.ncu-rep $O/r2_c5_kernels.ncu-rep $O/r2_c3_kernels.ncu-rep
compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_round2.py -m gpu -q -x -k "bucketed and 2000-30 or determinism and cluster or c1_batch or mul_gradient" > $O/r2_racecheck.log 2>&1
compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_round2.py -m gpu -q -x -k "bucketed or determinism or c1_batch or hub or unbatched or link_loss" > $O/r2_memcheck.log 2>&1
tail -n 4 $O/r2_racecheck.log; tail -n 4 $O/r2_memcheck.log; du -sh $O
