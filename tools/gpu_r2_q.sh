# round 2, GPU call Q (1 GPU): sequence entry with three buffers in rotation -- parity and e2e rate
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sequences.py tests/test_gpu_c_abi.py -x -q -m gpu > gpurun_out/q_pytest.log 2>&1; tail -2 gpurun_out/q_pytest.log
timeout 300 python tools/bench_e2e.py --threads 16,12,0,16 --out gpurun_out/q_e2e.json 2>/dev/null | cut -c1-110
