# round 2, GPU call Q (1 GPU): chunk size of gb2_scan_host_sequences against the end-to-end step (packers on, 16 threads)
set -x
mkdir -p gpurun_out
for cb in 67108864 33554432 16777216 8388608; do
  echo "chunk bases $cb"
  GB2_SEQ_CHUNK_BASES=$cb timeout 300 python tools/bench_e2e.py --threads 16,0 2>/dev/null | cut -c1-110
done
