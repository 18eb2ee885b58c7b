# round 2, job v: new parity tests (CFL-limited adaptive dt single + slabs, full-size oracle steps), then the whole GPU suite
set -x
nproc; free -g | head -2
timeout 1700 python -m pytest tests -m gpu -q -x -k "cfl_limited or full_size_step" 2>&1 | tail -8
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
