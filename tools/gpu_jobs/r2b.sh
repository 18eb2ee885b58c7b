# round 2, job b: ncu full capture of the new pair kernels (after 60 warm-up steps: disordered state)
set -x
ncu --set full --clock-control none --import-source on -k regex:"k_(force|density)_mask16" -s 120 -c 2 -o gpurun_out/prof_r2b_pair python bench.py --no-cpu --steps 2 --warmup 60 > gpurun_out/prof_r2b_pair.log 2>&1
tail -3 gpurun_out/prof_r2b_pair.log | cut -c1-300
ls -la gpurun_out/
