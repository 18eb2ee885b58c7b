# round 2 (session 3), job 4e: launch list of the bench command on the final code; the other single-GPU workloads; small-scene drop-in benchmark
set -x
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_|::k_" -s 0 -c 700 --csv --log-file gpurun_out/r4_final_launches_10M.csv python bench.py --no-cpu --also "" --steps 2 --warmup 3 > gpurun_out/r4_final_launches_10M.log 2>&1
tail -c 200 gpurun_out/r4_final_launches_10M.log
one() { python bench.py --no-cpu --also "" "$@" 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['config']['workload'], d['value'], d['ms_per_step'], d['e2e']['value'], d['extra']['stage_ms_rank0'], d['extra']['max_neighbors'])"; }
one --scene fluid_drop_1M --warmup 5 --steps 40
one --scene dam_break_1M --warmup 5 --steps 100
one --scene dam_break_347k --warmup 5 --steps 100
( cd dropin/_ref && timeout 300 ./performance_test 2>&1 | grep -E "Benchmarking|Average FPS|frame time" | head -12 )
