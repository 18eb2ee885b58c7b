ncu --set full --clock-control none --import-source on -k regex:"k_(force|density)_mask" -s 6 -c 2 -o gpurun_out/prof_r1d_pair python bench.py --no-cpu --steps 2 --warmup 2 > gpurun_out/prof_r1d_pair.log 2>&1
ncu --set full --clock-control none -k regex:"k_(onesweep_pass|onesweep_hist|reorder|integrate|cell_keys|cell_ends|scan_apply)" -s 30 -c 10 -o gpurun_out/prof_r1d_stream python bench.py --no-cpu --steps 2 --warmup 2 > gpurun_out/prof_r1d_stream.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 400 --csv --log-file gpurun_out/launches_r1d.csv python bench.py --no-cpu --steps 3 --warmup 3 > gpurun_out/b_r1d.log 2>&1
tail -1 gpurun_out/b_r1d.log | cut -c1-120
