# round 2, job l: duo kernels — parity file + ncu full capture
set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -12
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_(force|density)_duo" -s 120 -c 2 -o gpurun_out/prof_r2l_duo python bench.py --no-cpu --steps 2 --warmup 60 > gpurun_out/prof_r2l.log 2>&1
tail -2 gpurun_out/prof_r2l.log | cut -c1-200
