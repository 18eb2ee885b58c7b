#!/usr/bin/env bash
# run bench.py once with the in-tree libsphb.so and once per alternative build under tune/ (tuning sweeps)
# usage: tools/tune_run.sh [bench.py args...]   (default: --warmup 60 --steps 40)
args=("$@"); [ ${#args[@]} -eq 0 ] && args=(--warmup 60 --steps 40)
one() {
  python bench.py --no-cpu "${args[@]}" 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', d['value'], d['ms_per_step'], d['extra']['stage_ms_rank0'])"
}
one base
for lib in tune/libsphb_*.so; do
  [ -f "$lib" ] || continue
  name=$(basename $lib .so | sed 's/libsphb_//')
  SPHB_LIB=$PWD/$lib one $name
done
