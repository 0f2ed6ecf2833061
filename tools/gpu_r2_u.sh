# round 2, GPU call U (8 GPUs): bench --gpus 8 with the end-of-round code (parity key, default packer policy: off above two ranks)
set -x
mkdir -p gpurun_out
timeout 400 python bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/u_bench_8gpu.json 2> gpurun_out/u_bench_8gpu.err; tail -1 gpurun_out/u_bench_8gpu.err | cut -c1-200
python - <<'P'
import json
for ln in open('gpurun_out/u_bench_8gpu.json'):
    if ln.startswith('{'):
        d=json.loads(ln); print(d['value'], d['ms_per_step'], {k:v for k,v in d['e2e'].items() if k!='api'}, {k:(v['ms_per_step']) for k,v in d['e2e_variants'].items()}, d['parity'].get('ok'))
P
