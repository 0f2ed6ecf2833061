# round 2, GPU call G (1 GPU): dense finalize by two partition passes (csrc/dense_sort.cu) + K5 scans: all GPU tests, C5 numbers, launch list, short bench
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/g_pytest.log 2>&1; tail -5 gpurun_out/g_pytest.log
GB2_ONLY=c5 GB2_JSON=gpurun_out/g_configs_c5.json timeout 600 python tools/bench_configs.py > gpurun_out/g_configs_c5.log 2>&1; tail -9 gpurun_out/g_configs_c5.log
GB2_ONLY=c5 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/g_launches_c5.csv python tools/bench_configs.py > gpurun_out/g_c5_under_ncu.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err; tail -c 1500 gpurun_out/g_bench.json
