# round 2, job a: parity of the new 16-bit mask kernels + first bench
set -x
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2a_tests.log
cat gpurun_out/r2a_tests.log
timeout 600 python bench.py --no-cpu --warmup 60 --steps 100 > gpurun_out/r2a_bench_1M.json 2> gpurun_out/r2a_bench_1M.err
tail -c 1500 gpurun_out/r2a_bench_1M.json
timeout 600 python bench.py --no-cpu --warmup 10 --steps 20 > gpurun_out/r2a_bench_1M_w10.json 2>> gpurun_out/r2a_bench_1M.err
tail -c 600 gpurun_out/r2a_bench_1M_w10.json
