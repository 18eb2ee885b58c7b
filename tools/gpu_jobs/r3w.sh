# round 2 (session 3), job w: lanes per particle at the large sizes (expected: no gain — the passes are issue / L1 bound there)
set -x
one() { python bench.py --no-cpu --also "" "$@" 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['extra']['stage_ms_rank0'])"; }
for L in 1 2 4; do
  echo "lanes $L dam_break_347k"; SPHB_LANES=$L one --scene dam_break_347k --warmup 60 --steps 100
  echo "lanes $L dam_break_1M"; SPHB_LANES=$L one --scene dam_break_1M --warmup 60 --steps 60
  echo "lanes $L dam_break_10M"; SPHB_LANES=$L one --scene dam_break_10M --warmup 60 --steps 20
done
