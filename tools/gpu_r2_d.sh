# round 2, GPU call D (1 GPU): tests touched since call C, C3 prep timing, ncu captures (wide kernel, K2, launch lists)
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_vcf.py tests/test_gpu_configs.py tests/test_gpu_sequences.py -x -q 2>&1 | tail -15 > gpurun_out/d_pytest.log; tail -5 gpurun_out/d_pytest.log
GB2_MOTIF_TIMING=1 timeout 600 python tools/bench_c3.py --gpus 1 --out gpurun_out/d_c3_1gpu.json > gpurun_out/d_c3_1gpu.log 2>&1; grep -E "gb2_motif_create|prep_this_rank" gpurun_out/d_c3_1gpu.log | cut -c1-400 | tail -8
GB2_ONLY=wide timeout 600 ncu --set full --clock-control none --import-source on -k regex:gb2_score_wide -s 3 -c 1 -o gpurun_out/d_wide48_full python tools/bench_configs.py > gpurun_out/d_wide_ncu.log 2>&1; tail -2 gpurun_out/d_wide_ncu.log
GB2_PROFILE_RANGE=kmers timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gb2_score_kernel -c 1 -o gpurun_out/d_k2_full python bench.py --steps 1 --warmup 3 --no-kmer-e2e --no-cpu-baseline --no-graph-path > gpurun_out/d_k2_ncu.log 2>&1; tail -2 gpurun_out/d_k2_ncu.log | cut -c1-300
GB2_PROFILE_RANGE=kmers timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/d_launches_kmers.csv python bench.py --steps 2 --warmup 3 --no-kmer-e2e --no-cpu-baseline --no-graph-path > gpurun_out/d_launches_kmers.log 2>&1
GB2_PROFILE_RANGE=sequences timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/d_launches_seq.csv python bench.py --steps 2 --warmup 3 --no-kmer-e2e --no-cpu-baseline --no-graph-path > gpurun_out/d_launches_seq.log 2>&1
ls -la gpurun_out | tail -12
