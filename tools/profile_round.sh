# Profiling commands of round 1 (run on the GPU box from the repository root): launch lists of the bench step, of C5 (hit vs dense form)
# and per-launch metrics of the wide scoring kernel -> gpurun_out/*.csv; the summaries under profiles/ are made from those files.
set -x
mkdir -p gpurun_out
export GB2_PROFILE_RANGE=1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph-path --e2e-rows 1048576 > gpurun_out/bench_under_ncu.log 2>&1
unset GB2_PROFILE_RANGE
GB2_ONLY=c5 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c5.csv python tools/bench_configs.py > gpurun_out/c5_under_ncu.log 2>&1
GB2_ONLY=wide ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed -k regex:gb2_score_wide --clock-control none -c 9 --csv --log-file gpurun_out/wide_k2.csv python tools/bench_configs.py > gpurun_out/wide_under_ncu.log 2>&1
tail -3 gpurun_out/*.log
ls -la gpurun_out
