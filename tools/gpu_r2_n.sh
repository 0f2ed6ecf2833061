# round 2, GPU call N (1 GPU): final validation -- all GPU tests, smoke, bench (both arms), launch lists of the timed step, ncu full of the
# partition kernels and of K2 as they are at the end of the round
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/n_pytest.log 2>&1; tail -4 gpurun_out/n_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/n_smoke.log 2>&1; tail -1 gpurun_out/n_smoke.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/n_bench.json 2> gpurun_out/n_bench.err; tail -c 400 gpurun_out/n_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/n_bench_ref.json 2> gpurun_out/n_bench_ref.err; tail -c 300 gpurun_out/n_bench_ref.json
GB2_PROFILE_RANGE=kmers timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/n_launches_kmers.csv python bench.py --steps 2 --warmup 3 --no-kmer-e2e --no-cpu-baseline --no-graph-path > gpurun_out/n_launches_kmers.log 2>&1
GB2_PROFILE_RANGE=sequences timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/n_launches_seq.csv python bench.py --steps 2 --warmup 3 --no-kmer-e2e --no-cpu-baseline --no-graph-path > gpurun_out/n_launches_seq.log 2>&1
GB2_ONLY=c5 timeout 900 ncu --set full --import-source on --clock-control none -k regex:gb2_ds_scatter -s 4 -c 2 -o gpurun_out/n_ds_full -f python tools/bench_configs.py > gpurun_out/n_ds_full.log 2>&1
GB2_ONLY=c5 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/n_launches_c5.csv python tools/bench_configs.py > gpurun_out/n_c5_under_ncu.log 2>&1
GB2_ONLY=c5 GB2_JSON=gpurun_out/n_configs_c5.json timeout 600 python tools/bench_configs.py > gpurun_out/n_configs_c5.log 2>&1; grep partition gpurun_out/n_configs_c5.log | cut -c1-260
ls gpurun_out | grep "^n_"
