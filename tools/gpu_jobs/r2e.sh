set -x
SPHB_LIB=$PWD/tune/libsphb_nopf.so ncu --set full --clock-control none --import-source on -k regex:"k_density_mask16" -s 40 -c 1 -o gpurun_out/prof_r2e_dens python bench.py --no-cpu --steps 2 --warmup 40 > gpurun_out/prof_r2e.log 2>&1
tail -2 gpurun_out/prof_r2e.log | cut -c1-200
