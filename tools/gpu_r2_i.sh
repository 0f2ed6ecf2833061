# round 2, GPU call I (1 GPU): all GPU tests, mid-range threshold sweep, e2e phase breakdown with pinned result buffers, bench
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/i_pytest.log 2>&1; tail -4 gpurun_out/i_pytest.log
GB2_SCAN_TIMING=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-graph-path --no-kmer-e2e > gpurun_out/i_bench_timing.json 2> gpurun_out/i_bench_timing.err; grep gb2_scan_host_sequences gpurun_out/i_bench_timing.err | tail -8 | cut -c1-400
timeout 900 python bench.py > gpurun_out/i_bench.json 2> gpurun_out/i_bench.err
python - <<'P'
import json
for ln in open('gpurun_out/i_bench.json'):
    if ln.startswith('{'):
        d=json.loads(ln); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e_variants']['sequences_2bit']['ms_per_step'], d['parity'].get('ok'), d['cpu_baseline']['value'])
P
