# round 2, GPU call J (1 GPU): host packer threads of gb2_scan_host_sequences -- parity tests, e2e rate against the thread count
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sequences.py -x -q -m gpu > gpurun_out/j_pytest_seq.log 2>&1; tail -4 gpurun_out/j_pytest_seq.log
GB2_SCAN_TIMING=1 timeout 900 python tools/bench_e2e.py --threads 0,2,4,8,12,16,22 --out gpurun_out/j_e2e_threads.json > gpurun_out/j_e2e_threads.log 2> gpurun_out/j_e2e_threads.err; cat gpurun_out/j_e2e_threads.log | cut -c1-300; grep "packer threads" gpurun_out/j_e2e_threads.err | sort | uniq -c | head -20; grep "plan" gpurun_out/j_e2e_threads.err | tail -3 | cut -c1-400
