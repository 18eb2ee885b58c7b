# round 2, job t: the bench line as the driver runs it + per-launch ncu metrics of one step (10M and 1M) for profiles/traffic.json
set -x
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err
tail -c 3000 gpurun_out/r2t_bench.json; tail -3 gpurun_out/r2t_bench.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum
for sc in dam_break_10M dam_break_1M; do
  timeout 900 ncu --metrics $M --clock-control none -k regex:"^k_|::k_" -s 421 -c 7 --csv --log-file gpurun_out/r2t_step_$sc.csv python bench.py --scene $sc --also "" --no-cpu --steps 1 --warmup 3 > gpurun_out/r2t_ncu_$sc.log 2>&1
  tail -c 300 gpurun_out/r2t_ncu_$sc.log
done
