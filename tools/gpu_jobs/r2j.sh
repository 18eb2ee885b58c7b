# round 2, job j: ncu full capture of the staged pair kernels (after 60 warm-up steps: disordered state)
set -x
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_(force|density)_stage" -s 120 -c 2 -o gpurun_out/prof_r2j_stage python bench.py --no-cpu --steps 2 --warmup 60 > gpurun_out/prof_r2j.log 2>&1
tail -2 gpurun_out/prof_r2j.log | cut -c1-200
