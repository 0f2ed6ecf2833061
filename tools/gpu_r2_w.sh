# round 2, GPU call W (4 GPUs): C4 weak form -- 8 chr22-sized chromosomes per GPU (32 in all, 1.6 Gb x 5,008 haplotypes), global q-values, end-of-round code
set -x
mkdir -p gpurun_out
timeout 160 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29614 tools/bench_genome.py --chroms-per-gpu 8 --out gpurun_out/w_c4_4gpu_8each.json > gpurun_out/w_c4_4gpu_8each.log 2>&1; tail -1 gpurun_out/w_c4_4gpu_8each.log | cut -c1-1700
