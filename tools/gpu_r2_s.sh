# round 2, GPU call S (1 GPU): merged report table from Arrow buffers + report order on the device -- graph / drop-in tests, C4 phases on one GPU
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_graph.py tests/test_gpu_graph_scale.py tests/test_gpu_dropin.py tests/test_gpu_configs.py tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/s_pytest.log 2>&1; tail -4 gpurun_out/s_pytest.log
timeout 600 python tools/bench_genome.py --total-chroms 2 --out gpurun_out/s_c4_1gpu.json > gpurun_out/s_c4_1gpu.log 2>&1; tail -1 gpurun_out/s_c4_1gpu.log | cut -c1-1600
