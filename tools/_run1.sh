python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2
python bench.py > gpurun_out/bench_r1d.json 2> gpurun_out/bench_r1d.err; python -c "import json; d=json.load(open('gpurun_out/bench_r1d.json')); print('dam', d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], d['extra']['stage_ms_rank0'], d['cpu_baseline']['value'])"
python bench.py --no-cpu --steps 20 --scene dam_break_10M 2>&1 | tail -1 > gpurun_out/bench_r1d_10M.json; python -c "import json; d=json.load(open('gpurun_out/bench_r1d_10M.json')); print('10M', d['value'], d['ms_per_step'], d['e2e']['value'], d['extra']['stage_ms_rank0'])"
python bench.py --no-cpu --steps 20 --scene fluid_drop_1M 2>&1 | tail -1 > gpurun_out/bench_r1d_drop.json; python -c "import json; d=json.load(open('gpurun_out/bench_r1d_drop.json')); print('drop', d['value'], d['ms_per_step'], d['e2e']['value'], d['extra']['stage_ms_rank0'])"
python bench.py --no-cpu --steps 20 --scene dam_break_13k 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('13k', d['value'], d['ms_per_step'], d['extra']['stage_ms_rank0'])"
