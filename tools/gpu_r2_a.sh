# round 2, GPU call A: new sequence-form tests, whole gpu suite, bench with both CTA sizes of the sequence kernel,
# fresh wide-kernel ncu numbers, ncu --set full of the sequence kernel
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv
python -m pytest tests/test_gpu_sequences.py -x -q 2>&1 | tail -15 > gpurun_out/a_pytest_seq.log; tail -5 gpurun_out/a_pytest_seq.log
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/a_pytest_all.log; tail -5 gpurun_out/a_pytest_all.log
python bench.py --steps 20 --warmup 3 > gpurun_out/a_bench_1024.json 2> gpurun_out/a_bench_1024.err; tail -c 3000 gpurun_out/a_bench_1024.json; tail -5 gpurun_out/a_bench_1024.err
GB2_SEQ_THREADS=512 python bench.py --steps 20 --warmup 3 --no-kmer-e2e --no-cpu-baseline --no-graph-path > gpurun_out/a_bench_512.json 2> gpurun_out/a_bench_512.err; tail -c 1500 gpurun_out/a_bench_512.json
GB2_ONLY=wide ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed -k regex:gb2_score_wide --clock-control none -c 9 --csv --log-file gpurun_out/a_wide_k2.csv python tools/bench_configs.py > gpurun_out/a_wide_under_ncu.log 2>&1
GB2_PROFILE_RANGE=sequences ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gb2_score_seq -c 1 -o gpurun_out/a_seq_full python bench.py --steps 1 --warmup 3 --no-kmer-e2e --no-cpu-baseline --no-graph-path > gpurun_out/a_seq_ncu.log 2>&1
tail -3 gpurun_out/a_seq_ncu.log
ls -la gpurun_out
