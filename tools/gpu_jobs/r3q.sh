# round 2 (session 3), job q: whole GPU suite (incl. the 10.7 M full-size oracle step), then compute-sanitizer on the new device paths
set -x
timeout 2400 python -m pytest tests -m gpu -q --durations=8 2>&1 | grep -v "Warning: Particle" | tail -16
bash tools/gpu_jobs/r3l.sh 2>&1 | tail -20
