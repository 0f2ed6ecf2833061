# round 2, GPU call F (8 GPUs): C3 sharded by motif at 8 GPUs (+ parity vs one GPU), bench --gpus 8 (parity key, e2e), C4 strong scaling
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 500 python tools/bench_c3.py --gpus 8 --out gpurun_out/f_c3_8gpu.json > gpurun_out/f_c3_8gpu.log 2>&1; tail -1 gpurun_out/f_c3_8gpu.log | cut -c1-1600
timeout 700 python bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/f_bench_8gpu.json 2> gpurun_out/f_bench_8gpu.err; tail -c 2500 gpurun_out/f_bench_8gpu.json; tail -3 gpurun_out/f_bench_8gpu.err | cut -c1-300
for n in 8 4 2; do
  timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n tools/bench_genome.py --total-chroms 16 --out gpurun_out/f_c4_strong_${n}gpu.json > gpurun_out/f_c4_strong_${n}gpu.log 2>&1; tail -1 gpurun_out/f_c4_strong_${n}gpu.log | cut -c1-1200
done
ls -la gpurun_out | tail -8
