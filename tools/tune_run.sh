#!/usr/bin/env bash
# run bench.py once per alternative build of libsphb.so under tune/ (tuning sweeps)
for lib in tune/libsphb_*.so; do
  name=$(basename $lib .so | sed 's/libsphb_//')
  SPHB_LIB=$PWD/$lib python bench.py --no-cpu --steps 10 "$@" 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$name', d['value'], d['ms_per_step'], d['extra']['stage_ms_rank0'])"
done
