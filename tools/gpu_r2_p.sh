# round 2, GPU call P (4 GPUs): do the host packers pay with four ranks on the host?  (e2e with 6 packers per rank against the copy-engine-only variant of the same run)
set -x
mkdir -p gpurun_out
nproc
GB2_HOST_PACK_THREADS=6 timeout 500 python bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/p_bench_4gpu.json 2> gpurun_out/p_bench_4gpu.err; tail -1 gpurun_out/p_bench_4gpu.err | cut -c1-200
python - <<'P'
import json
for ln in open('gpurun_out/p_bench_4gpu.json'):
    if ln.startswith('{'):
        d=json.loads(ln); print(d['value'], d['ms_per_step'], {k:v for k,v in d['e2e'].items() if k!='api'}, {k:(v['ms_per_step']) for k,v in d['e2e_variants'].items()}, d['parity'].get('ok'))
P
