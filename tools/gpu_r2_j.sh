# round 2, GPU call J (1 GPU): host packer threads of gb2_scan_host_sequences -- parity tests, e2e rate against the thread count
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sequences.py -x -q -m gpu > gpurun_out/j_pytest_seq.log 2>&1; tail -4 gpurun_out/j_pytest_seq.log
timeout 900 python tools/bench_e2e.py --threads 0,4,8,12,16 --out gpurun_out/j_e2e_threads2.json > gpurun_out/j_e2e_threads2.log 2> gpurun_out/j_e2e_threads2.err; cat gpurun_out/j_e2e_threads2.log | cut -c1-200
