# round 2, GPU call H (2 GPUs): the tests a one-GPU box skips + bench --gpus 2 with the default packer policy, end-of-round code
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_dropin.py tests/test_gpu_graph.py -x -q -m gpu -k "two_gpus or world_2 or gpus_2 or comm" > gpurun_out/h_pytest_2gpu.log 2>&1; tail -3 gpurun_out/h_pytest_2gpu.log
timeout 500 python bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/h_bench_2gpu.json 2> gpurun_out/h_bench_2gpu.err; tail -1 gpurun_out/h_bench_2gpu.err | cut -c1-200
python - <<'P'
import json
for ln in open('gpurun_out/h_bench_2gpu.json'):
    if ln.startswith('{'):
        d=json.loads(ln); print(d['value'], d['ms_per_step'], {k:v for k,v in d['e2e'].items() if k!='api'}, {k:(v['ms_per_step']) for k,v in d['e2e_variants'].items()}, d['parity'].get('ok'))
P
