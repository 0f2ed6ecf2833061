# round 2, GPU call T (8 GPUs): C4 strong scaling at 8 GPUs with the merged table built from device-ordered columns and Arrow buffers
set -x
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29618 tools/bench_genome.py --total-chroms 16 --out gpurun_out/t_c4_strong_8gpu.json > gpurun_out/t_c4_strong_8gpu.log 2>&1; tail -1 gpurun_out/t_c4_strong_8gpu.log | cut -c1-1600
