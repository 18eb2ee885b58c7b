# round 2 (session 3), job l: compute-sanitizer memcheck on this session's new device paths — multi-device engine (slabs on one GPU:
# exchange count / split / append, peer copies), Wendland / Gaussian tested-walk kernels, CUDA-graph replay
set -x
LOG=gpurun_out/r3l_sanitizer.log
echo "== memcheck: tests/test_multi_gpu.py (dam break on 2 and 3 slabs, migration, adaptive dt, strided records, thin scenes, halo sliver)" >> $LOG
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_multi_gpu.py -m gpu -q -x 2>&1 | grep -E "=========|ERROR SUMMARY|passed|failed" | tail -6 >> $LOG
echo "== memcheck: kernel classes + step graphs (tests/test_gpu_parity.py -k 'kernel or graphs')" >> $LOG
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "kernel or graphs" 2>&1 | grep -E "=========|ERROR SUMMARY|passed|failed" | tail -6 >> $LOG
echo "== racecheck: tests/test_multi_gpu.py -k 'dam_break or sliver'" >> $LOG
timeout 1500 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_multi_gpu.py -m gpu -q -x -k "dam_break or sliver" 2>&1 | grep -E "=========|RACECHECK SUMMARY|passed|failed" | tail -6 >> $LOG
cat $LOG
