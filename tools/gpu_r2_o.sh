# round 2, GPU call O (1 GPU): the e2e step inside bench.py against tools/bench_e2e.py on the same box
set -x
mkdir -p gpurun_out
nproc
timeout 600 python tools/bench_e2e.py --threads 16,12,16 > gpurun_out/o_e2e.log 2>&1; cut -c1-120 gpurun_out/o_e2e.log
GB2_SCAN_TIMING=1 timeout 900 python bench.py --steps 20 --no-cpu-baseline --no-graph-path --no-kmer-e2e > gpurun_out/o_bench.json 2> gpurun_out/o_bench.err
grep "gb2_scan_host_sequences" gpurun_out/o_bench.err | head -12 | cut -c1-330
python - <<'P'
import json
for ln in open('gpurun_out/o_bench.json'):
    if ln.startswith('{'):
        d=json.loads(ln); print({k:v for k,v in d['e2e'].items() if k!='api'})
P
timeout 600 python tools/bench_e2e.py --threads 16 > gpurun_out/o_e2e2.log 2>&1; cut -c1-120 gpurun_out/o_e2e2.log
