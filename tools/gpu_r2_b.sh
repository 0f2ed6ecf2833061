# round 2, GPU call B (1 GPU): whole gpu suite after the chunk-plan / batched-motif / many-motif changes, config numbers as JSON,
# C3 on one GPU, the reference arm, the bench line
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/b_pytest_all.log; tail -6 gpurun_out/b_pytest_all.log
for sec in wide narrow c5; do GB2_ONLY=$sec GB2_JSON=gpurun_out/b_configs_$sec.json python tools/bench_configs.py > gpurun_out/b_configs_$sec.log 2>&1; tail -12 gpurun_out/b_configs_$sec.log; done
python tools/bench_c3.py --gpus 1 --out gpurun_out/b_c3_1gpu.json > gpurun_out/b_c3_1gpu.log 2>&1; tail -3 gpurun_out/b_c3_1gpu.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/b_bench_ref.json 2> gpurun_out/b_bench_ref.err; tail -c 1500 gpurun_out/b_bench_ref.json; tail -3 gpurun_out/b_bench_ref.err
python bench.py --steps 20 --warmup 3 > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err; tail -c 2500 gpurun_out/b_bench.json; tail -5 gpurun_out/b_bench.err
ls -la gpurun_out
