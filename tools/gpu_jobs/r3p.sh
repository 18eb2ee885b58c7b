# round 2 (session 3), job p: which halo copies carry a density different from the single context (build without the sliver, ghosts exported)?
set -x
SPHB_LIB=$PWD/tune/libsphb_nosliver_dbg.so timeout 1500 python tools/debug/slab_ghost_fields.py 47 2>&1 | tail -30
