# round 2, GPU call C (2 GPUs): hardware multi-GPU tests, bench --gpus 2 with its parity key, C3 sharded by motif, reference arm,
# wide / narrow kernel numbers after the R templating
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
python -m pytest tests/test_gpu_multi.py tests/test_gpu_sequences.py tests/test_gpu_vcf.py -x -q 2>&1 | tail -15 > gpurun_out/c_pytest_new.log; tail -5 gpurun_out/c_pytest_new.log
python -m pytest tests/test_gpu_dropin.py tests/test_gpu_graph.py tests/test_gpu_kernels.py tests/test_gpu_configs.py -x -q 2>&1 | tail -15 > gpurun_out/c_pytest_old.log; tail -5 gpurun_out/c_pytest_old.log
for sec in wide narrow; do GB2_ONLY=$sec GB2_JSON=gpurun_out/c_configs_$sec.json python tools/bench_configs.py > gpurun_out/c_configs_$sec.log 2>&1; tail -9 gpurun_out/c_configs_$sec.log; done
python tools/bench_c3.py --gpus 1 --out gpurun_out/c_c3_1gpu.json > gpurun_out/c_c3_1gpu.log 2>&1; tail -2 gpurun_out/c_c3_1gpu.log | cut -c1-1200
python tools/bench_c3.py --gpus 2 --out gpurun_out/c_c3_2gpu.json > gpurun_out/c_c3_2gpu.log 2>&1; tail -2 gpurun_out/c_c3_2gpu.log | cut -c1-1500
python bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/c_bench_2gpu.json 2> gpurun_out/c_bench_2gpu.err; tail -c 3000 gpurun_out/c_bench_2gpu.json; tail -5 gpurun_out/c_bench_2gpu.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/c_bench_ref.json 2> gpurun_out/c_bench_ref.err; tail -c 1200 gpurun_out/c_bench_ref.json; tail -3 gpurun_out/c_bench_ref.err
ls -la gpurun_out
