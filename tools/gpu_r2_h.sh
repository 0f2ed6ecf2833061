# round 2, GPU call H (2 GPUs): C4 merged table with the NCCL warm-up in gb2_comm_init + phase timers; 2-GPU tests; raw H2D rate at 1 and 2 ranks
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_dropin.py tests/test_gpu_graph.py -x -q -m gpu -k "two_gpus or world_2 or gpus_2 or comm" > gpurun_out/h_pytest_2gpu.log 2>&1; tail -4 gpurun_out/h_pytest_2gpu.log
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 tools/bench_genome.py --total-chroms 4 --out gpurun_out/h_c4_strong_2gpu.json > gpurun_out/h_c4_strong_2gpu.log 2>&1; tail -1 gpurun_out/h_c4_strong_2gpu.log | cut -c1-1800
for n in 1 2; do timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2962$n tools/h2d_probe.py --out gpurun_out/h_h2d_probe.jsonl > gpurun_out/h_h2d_$n.log 2>&1; tail -1 gpurun_out/h_h2d_$n.log | cut -c1-600; done
nvidia-smi topo -m > gpurun_out/h_topo.txt 2>&1; lscpu | grep -i -E "numa|socket|model name|^CPU\(s\)" >> gpurun_out/h_topo.txt; cat gpurun_out/h_topo.txt | head -30
