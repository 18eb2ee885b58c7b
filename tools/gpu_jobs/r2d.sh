bash tools/tune_run.sh --warmup 20 --steps 20 2>&1 | tee gpurun_out/r2d_tune.txt
