# round 2, GPU call M (1 GPU): compute-sanitizer (memcheck, racecheck) over the new round-2 kernels; full test suite + bench sanity
set -x
mkdir -p gpurun_out
export GB2_SEQ_CHUNK_BASES=8192
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "dense_finalize_equals or k5 or goldens" > gpurun_out/m_memcheck_dense.log 2>&1; echo "memcheck dense rc=$?"; tail -3 gpurun_out/m_memcheck_dense.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "dense_finalize_equals" > gpurun_out/m_racecheck_dense.log 2>&1; echo "racecheck dense rc=$?"; tail -3 gpurun_out/m_racecheck_dense.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_sequences.py -x -q -m gpu -k "host_packers or equals_scan_host" > gpurun_out/m_memcheck_seq.log 2>&1; echo "memcheck seq rc=$?"; tail -3 gpurun_out/m_memcheck_seq.log
unset GB2_SEQ_CHUNK_BASES
grep -c "ERROR SUMMARY: 0 errors" gpurun_out/m_*.log
grep -h "ERROR SUMMARY" gpurun_out/m_*.log | sort | uniq -c
