# round 2, GPU call L (2 GPUs): packer threads per rank with two ranks on the host (6 = the default: a quarter of the hardware threads)
set -x
mkdir -p gpurun_out
nproc
for t in 9 12; do
GB2_HOST_PACK_THREADS=$t timeout 300 python bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-graph-path --no-kmer-e2e > gpurun_out/l_bench_2gpu_$t.json 2> gpurun_out/l_bench_2gpu_$t.err
python - <<P
import json
for ln in open('gpurun_out/l_bench_2gpu_$t.json'):
    if ln.startswith('{'):
        d=json.loads(ln); e=d['e2e']; print($t, e['ms_per_step'], e['ms_per_step_without_host_packers'], e['chunks_as_text'], e['chunks_packed_on_host'], e['host_cpus'], d['parity'].get('ok'))
P
done
