# round 2, job w: compute-sanitizer memcheck + racecheck + synccheck on the 13k dam break (default fast path and strict), log kept under profiles/
set -x
for tool in memcheck racecheck synccheck; do
  echo "== $tool (fast path, 3 steps)" >> gpurun_out/r2w_sanitizer.log
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python bench.py --scene dam_break_13k --also "" --no-cpu --steps 1 --warmup 3 2>&1 | grep -E "=========|ERROR SUMMARY|RACECHECK SUMMARY" | tail -12 >> gpurun_out/r2w_sanitizer.log
done
echo "== memcheck (pair mode 1: both passes staged)" >> gpurun_out/r2w_sanitizer.log
SPHB_PAIR_MODE=1 timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python bench.py --scene dam_break_13k --also "" --no-cpu --steps 1 --warmup 3 2>&1 | grep -E "=========|ERROR SUMMARY" | tail -12 >> gpurun_out/r2w_sanitizer.log
echo "== racecheck (pair mode 1)" >> gpurun_out/r2w_sanitizer.log
SPHB_PAIR_MODE=1 timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python bench.py --scene dam_break_13k --also "" --no-cpu --steps 1 --warmup 3 2>&1 | grep -E "=========|RACECHECK SUMMARY" | tail -12 >> gpurun_out/r2w_sanitizer.log
echo "== memcheck (strict mode)" >> gpurun_out/r2w_sanitizer.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python bench.py --scene dam_break_13k --also "" --no-cpu --steps 1 --warmup 3 --math strict 2>&1 | grep -E "=========|ERROR SUMMARY" | tail -12 >> gpurun_out/r2w_sanitizer.log
cat gpurun_out/r2w_sanitizer.log
