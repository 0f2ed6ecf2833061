# round 2, GPU call L (2 GPUs): do the host packers still pay when two ranks share the host?  e2e with 6 packer threads per rank against the
# copy-engine-only variant of the same run
set -x
mkdir -p gpurun_out
GB2_HOST_PACK_THREADS=6 timeout 500 python bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/l_bench_2gpu.json 2> gpurun_out/l_bench_2gpu.err; tail -2 gpurun_out/l_bench_2gpu.err | cut -c1-300
python - <<'P'
import json
for ln in open('gpurun_out/l_bench_2gpu.json'):
    if ln.startswith('{'):
        d=json.loads(ln); print(d['value'], d['ms_per_step'], {k:v for k,v in d['e2e'].items() if k!='api'}, {k:(v['ms_per_step']) for k,v in d['e2e_variants'].items()}, d['parity'].get('ok'))
P
