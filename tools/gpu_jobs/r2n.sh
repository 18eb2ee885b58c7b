# round 2, job n: ncu full capture of the per-warp staged pair kernels
set -x
SPHB_PAIR_MODE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_(force|density)_stage" -s 120 -c 2 -o gpurun_out/prof_r2n_wstage python bench.py --no-cpu --steps 2 --warmup 60 > gpurun_out/prof_r2n.log 2>&1
tail -2 gpurun_out/prof_r2n.log | cut -c1-200
