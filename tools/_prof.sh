ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 400 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --no-cpu --steps 3 --warmup 3 > gpurun_out/b_r1c.log 2>&1
tail -1 gpurun_out/b_r1c.log | cut -c1-300
