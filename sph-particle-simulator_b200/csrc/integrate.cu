// Fused leapfrog integration + AABB boundary clamp + CFL reduction (the on-device timestep rule lives in neighbor.cu:
// k_cell_count sets the dt of a step).
//
// Replaces SPHEngine::integrate_leapfrog (reference src/sph_engine.cpp:290-310),
// ParticleSystem::apply_boundary_conditions (src/particle.cpp:122-153, serial and untimed in the
// reference), the max|v| loop of SPHEngine::compute_cfl_timestep (src/sph_engine.cpp:312-333) and
// "current_time_ += dt" (src/sph_engine.cpp:138) in ONE streaming pass: 48 B read + 32 B written per
// particle.  Every fp32 operation is issued with explicit round-to-nearest intrinsics in the
// reference's association order, so this stage is bit-exact in both math modes.
#include "sphb_internal.cuh"

namespace sphb {

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ void clamp_axis(float& p, float& v, float lo, float hi) {
    // particle.cpp:127-133: restitution 0.8 hard-coded, independent of params.damping
    if (p < lo) { p = lo; v = __fmul_rn(v, -0.8f); }
    else if (p > hi) { p = hi; v = __fmul_rn(v, -0.8f); }
}

__global__ void __launch_bounds__(kThreads) k_integrate(size_t n, float4* __restrict__ posm, float4* __restrict__ velid,
                                                        const float4* __restrict__ acc, IntegrateConsts ic,
                                                        DeviceScalars* sc, int* __restrict__ box) {
    const size_t s = (size_t)blockIdx.x * kThreads + threadIdx.x;
    const float dt = sc->dt;
    float v2 = 0.0f;
    if (s < n && !(__float_as_uint(velid[s].w) & 0x80000000u)) {   // bit 31 of the id word: halo copy (slab mode), not advanced
        float4 p = posm[s];
        float4 v = velid[s];
        const float4 a = acc[s];
        // v += 0.5f * a * dt   ((0.5f * a) * dt, sph_engine.cpp:299)
        const float hx = __fmul_rn(__fmul_rn(0.5f, a.x), dt);
        const float hy = __fmul_rn(__fmul_rn(0.5f, a.y), dt);
        const float hz = __fmul_rn(__fmul_rn(0.5f, a.z), dt);
        v.x = __fadd_rn(v.x, hx); v.y = __fadd_rn(v.y, hy); v.z = __fadd_rn(v.z, hz);
        // x += v * dt  (302)
        p.x = __fadd_rn(p.x, __fmul_rn(v.x, dt)); p.y = __fadd_rn(p.y, __fmul_rn(v.y, dt)); p.z = __fadd_rn(p.z, __fmul_rn(v.z, dt));
        // v += 0.5f * a * dt  (305), v *= damping (308)
        v.x = __fadd_rn(v.x, hx); v.y = __fadd_rn(v.y, hy); v.z = __fadd_rn(v.z, hz);
        v.x = __fmul_rn(v.x, ic.damping); v.y = __fmul_rn(v.y, ic.damping); v.z = __fmul_rn(v.z, ic.damping);
        clamp_axis(p.x, v.x, ic.xmin, ic.xmax);
        clamp_axis(p.y, v.y, ic.ymin, ic.ymax);
        clamp_axis(p.z, v.z, ic.zmin, ic.zmax);
        posm[s] = p;
        velid[s] = v;
        v2 = __fadd_rn(__fadd_rn(__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y)), __fmul_rn(v.z, v.z));
        if (__float_as_uint(v.w) == 0u) { sc->a0[0] = a.x; sc->a0[1] = a.y; sc->a0[2] = a.z; sc->a0_fresh = 1u; }
    }
    if (s == 0) sc->time = __fadd_rn(sc->time, dt);   // current_time_ += dt (sph_engine.cpp:138)
    unsigned bits = (v2 == v2) ? __float_as_uint(v2) : 0u;
    bits = __reduce_max_sync(0xffffffffu, bits);
    // one same-address atomic per warp serialises at L2 (334 k warps at 10 M particles): only warps that can
    // still raise the running maximum issue it
    if ((threadIdx.x & 31) == 0 && bits > *(volatile unsigned int*)&sc->max_v2_bits) atomicMax(&sc->max_v2_bits, bits);
    if (box) {
        // sparse scenes: bounding box of the new positions (ordered-int images of the floats), so the next
        // step can size its cell table from the particles instead of the much larger AABB
        int lo[3] = {0x7FFFFFFF, 0x7FFFFFFF, 0x7FFFFFFF}, hi[3] = {(int)0x80000000, (int)0x80000000, (int)0x80000000};
        if (s < n) {
            const float4 p = posm[s];
            const float c[3] = {p.x, p.y, p.z};
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                if (fabsf(c[a]) <= 3.4028235e38f) {
                    const int i = __float_as_int(c[a]);
                    lo[a] = hi[a] = i >= 0 ? i : i ^ 0x7FFFFFFF;
                }
            }
        }
        __shared__ int s_lo[3][kThreads / 32], s_hi[3][kThreads / 32];
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const int l = __reduce_min_sync(0xffffffffu, lo[a]), h = __reduce_max_sync(0xffffffffu, hi[a]);
            if (lane == 0) { s_lo[a][w] = l; s_hi[a][w] = h; }
        }
        __syncthreads();
        if (threadIdx.x < 3) {   // one atomic pair per axis per CTA
            const int a = threadIdx.x;
            int l = s_lo[a][0], h = s_hi[a][0];
#pragma unroll
            for (int k = 1; k < kThreads / 32; ++k) { l = min(l, s_lo[a][k]); h = max(h, s_hi[a][k]); }
            if (l != 0x7FFFFFFF) atomicMin(&box[a], l);
            if (h != (int)0x80000000) atomicMax(&box[3 + a], h);
        }
    }
}

}  // namespace

int launch_integrate(size_t n, float4* posm, float4* velid, const float4* acc, IntegrateConsts ic, DeviceScalars* sc,
                     int* d_box_or_null, cudaStream_t st) {
    k_integrate<<<(unsigned)((n + kThreads - 1) / kThreads), kThreads, 0, st>>>(n, posm, velid, acc, ic, sc, d_box_or_null);
    return 1;
}

}  // namespace sphb
