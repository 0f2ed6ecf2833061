# round 2, GPU call K (8 GPUs): raw H2D rate of the box at 8 and 4 concurrent ranks, bench --gpus 8 (parity key, e2e with host packers),
# C4 strong scaling at 8 GPUs after the NCCL warm-up, C3 by motif at 8 GPUs
set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/k_topo.txt 2>&1; lscpu | grep -i -E "numa|socket|model name|^CPU\(s\)" >> gpurun_out/k_topo.txt; head -14 gpurun_out/k_topo.txt
for n in 8 4; do timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2962$n tools/h2d_probe.py --reps 4 --out gpurun_out/k_h2d_probe.jsonl > gpurun_out/k_h2d_$n.log 2>&1; tail -1 gpurun_out/k_h2d_$n.log | cut -c1-500; done
timeout 500 python bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/k_bench_8gpu.json 2> gpurun_out/k_bench_8gpu.err; tail -3 gpurun_out/k_bench_8gpu.err | cut -c1-300
python - <<'P'
import json
for ln in open('gpurun_out/k_bench_8gpu.json'):
    if ln.startswith('{'):
        d=json.loads(ln); print(d['value'], d['ms_per_step'], {k:v for k,v in d['e2e'].items() if k!='api'}, {k:(v['ms_per_step']) for k,v in d['e2e_variants'].items()}, d['parity'].get('ok'))
P
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29618 tools/bench_genome.py --total-chroms 16 --out gpurun_out/k_c4_strong_8gpu.json > gpurun_out/k_c4_strong_8gpu.log 2>&1; tail -1 gpurun_out/k_c4_strong_8gpu.log | cut -c1-1500
timeout 300 python tools/bench_c3.py --gpus 8 --out gpurun_out/k_c3_8gpu.json > gpurun_out/k_c3_8gpu.log 2>&1; tail -1 gpurun_out/k_c3_8gpu.log | cut -c1-900
