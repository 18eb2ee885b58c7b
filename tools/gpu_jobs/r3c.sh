# round 2 (session 3), job c: A/B of the force-record layout (interleaved 32-byte records) and the direct t factor
set -x
bash tools/tune_run.sh --scene dam_break_10M --also "" --warmup 60 --steps 30 2>&1 | grep -v "^+" > gpurun_out/r3c_tune_10M.txt
cat gpurun_out/r3c_tune_10M.txt
bash tools/tune_run.sh --scene dam_break_1M --also "" --warmup 60 --steps 60 2>&1 | grep -v "^+" > gpurun_out/r3c_tune_1M.txt
cat gpurun_out/r3c_tune_1M.txt
SPHB_LIB=$PWD/tune/libsphb_rec32t.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
