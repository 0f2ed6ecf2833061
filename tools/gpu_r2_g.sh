# round 2, GPU call G (1 GPU): dense finalize by two partition passes (csrc/dense_sort.cu): parity tests, C5 numbers, launch list, ncu full of the scatter kernels
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_report.py tests/test_gpu_dropin.py tests/test_gpu_configs.py -x -q -m gpu > gpurun_out/g_pytest.log 2>&1; tail -5 gpurun_out/g_pytest.log
GB2_DS_THREADS=256 timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "dense" > gpurun_out/g_pytest_dense_256.log 2>&1; tail -2 gpurun_out/g_pytest_dense_256.log
GB2_ONLY=c5 GB2_JSON=gpurun_out/g_configs_c5.json timeout 600 python tools/bench_configs.py > gpurun_out/g_configs_c5.log 2>&1; tail -11 gpurun_out/g_configs_c5.log
GB2_ONLY=c5 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/g_launches_c5.csv python tools/bench_configs.py > gpurun_out/g_c5_under_ncu.log 2>&1
GB2_ONLY=c5 timeout 900 ncu --set full --import-source on --clock-control none -k regex:gb2_ds_scatter -s 4 -c 2 -o gpurun_out/g_ds_full -f python tools/bench_configs.py > gpurun_out/g_ds_full.log 2>&1
tail -3 gpurun_out/g_ds_full.log
