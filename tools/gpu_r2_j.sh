# round 2, GPU call J (1 GPU): host packer threads of gb2_scan_host_sequences -- parity tests, e2e rate against the thread count, bench
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sequences.py tests/test_host_cpu.py -x -q > gpurun_out/j_pytest_seq.log 2>&1; tail -4 gpurun_out/j_pytest_seq.log
timeout 900 python tools/bench_e2e.py --threads 0,2,4,8,12,16 --out gpurun_out/j_e2e_threads.json > gpurun_out/j_e2e_threads.log 2> gpurun_out/j_e2e_threads.err; cat gpurun_out/j_e2e_threads.log | cut -c1-200
timeout 900 python bench.py > gpurun_out/j_bench.json 2> gpurun_out/j_bench.err; tail -2 gpurun_out/j_bench.err
python - <<'P'
import json
for ln in open('gpurun_out/j_bench.json'):
    if ln.startswith('{'):
        d=json.loads(ln); print(d['value'], d['ms_per_step'], {k:v for k,v in d['e2e'].items() if k!='api'}, {k:(v['ms_per_step']) for k,v in d['e2e_variants'].items()}, d['parity'].get('ok'), d['cpu_baseline']['value'])
P
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/j_bench_ref.json 2> gpurun_out/j_bench_ref.err; tail -c 600 gpurun_out/j_bench_ref.json
