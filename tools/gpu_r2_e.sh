# round 2, GPU call E (1 GPU): K4 exact path / wide kernel split -- tests, C3 prep timing, wide numbers
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_kernels.py tests/test_gpu_random_motifs.py tests/test_gpu_configs.py tests/test_gpu_dropin.py -x -q 2>&1 | tail -15 > gpurun_out/e_pytest.log; tail -5 gpurun_out/e_pytest.log
GB2_MOTIF_TIMING=1 timeout 600 python tools/bench_c3.py --gpus 1 --out gpurun_out/e_c3_1gpu.json > gpurun_out/e_c3_1gpu.log 2>&1; grep -E "gb2_motif_create" gpurun_out/e_c3_1gpu.log | tail -6; tail -1 gpurun_out/e_c3_1gpu.log | cut -c1-900
GB2_ONLY=wide GB2_JSON=gpurun_out/e_configs_wide.json timeout 600 python tools/bench_configs.py > gpurun_out/e_configs_wide.log 2>&1; tail -4 gpurun_out/e_configs_wide.log
