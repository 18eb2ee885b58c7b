# round 2 (session 3), job d: A/B of the mask-word prefetch in the force pass
set -x
bash tools/tune_run.sh --scene dam_break_10M --also "" --warmup 60 --steps 30 2>&1 | grep -v "^+" > gpurun_out/r3d_tune_10M.txt
cat gpurun_out/r3d_tune_10M.txt
bash tools/tune_run.sh --scene dam_break_1M --also "" --warmup 60 --steps 60 2>&1 | grep -v "^+" > gpurun_out/r3d_tune_1M.txt
cat gpurun_out/r3d_tune_1M.txt
