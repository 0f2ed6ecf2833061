# round 2, GPU call N (1 GPU): final validation -- all GPU tests, smoke, bench (both arms)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/n_pytest.log 2>&1; tail -4 gpurun_out/n_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/n_smoke.log 2>&1; tail -1 gpurun_out/n_smoke.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/n_bench.json 2> gpurun_out/n_bench.err; tail -c 300 gpurun_out/n_bench.json
python - <<'P'
import json
for ln in open('gpurun_out/n_bench.json'):
    if ln.startswith('{'):
        d=json.loads(ln); print(d['value'], d['ms_per_step'], d['roofline']['frac'], {k:v for k,v in d['e2e'].items() if k!='api'}, {k:round(v['ms_per_step'],2) for k,v in d['e2e_variants'].items()}, d['parity'].get('ok'), d['cpu_baseline']['value'])
P
