# round 2, final profiles of the default path on the headline workload (dam_break_10M) and on dam_break_1M:
#   1. ncu --set full of the two pair kernels (after the 60-step pre-roll)
#   2. the launch list of the bench command (gpu__time_duration.sum)
#   3. per-launch dram bytes / warp instructions of one step -> profiles/traffic.json (tools/ncu_step_metrics.py)
set -x
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_density_stage|k_force_mask16" -s 120 -c 2 -o gpurun_out/prof_r2_final_pair_10M python bench.py --no-cpu --also "" --steps 2 --warmup 3 > gpurun_out/prof_r2_final_pair_10M.log 2>&1
tail -c 200 gpurun_out/prof_r2_final_pair_10M.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_|::k_" -s 0 -c 600 --csv --log-file gpurun_out/r2_final_launches_10M.csv python bench.py --no-cpu --also "" --steps 2 --warmup 3 > gpurun_out/r2_final_launches_10M.log 2>&1
tail -c 200 gpurun_out/r2_final_launches_10M.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum
for sc in dam_break_10M dam_break_1M; do
  timeout 900 ncu --metrics $M --clock-control none -k regex:"^k_|::k_" -s 421 -c 7 --csv --log-file gpurun_out/r2_final_step_$sc.csv python bench.py --scene $sc --also "" --no-cpu --steps 1 --warmup 3 > gpurun_out/r2_final_ncu_$sc.log 2>&1
done
ls -la gpurun_out | tail -8
