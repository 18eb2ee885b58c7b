ncu --set full --clock-control none --import-source on -k regex:"k_(force|density)_mask" -s 6 -c 2 -o gpurun_out/prof_mask3 python bench.py --no-cpu --steps 2 --warmup 2 > gpurun_out/prof_mask3.log 2>&1
tail -1 gpurun_out/prof_mask3.log | cut -c1-100
