// Per-pair SPH arithmetic shared by the pair kernels.
//
// STRICT: every operation is an explicit round-to-nearest intrinsic (never contracted to FMA) in
// the association order of the reference's C++ expressions, so results are bit-identical to the
// reference compiled without -ffast-math:
//   density   equations::compute_density      reference src/sph_engine.cpp:370-383
//             CubicSplineKernel::W / W_3d      src/kernels.cpp:140-143, 58-66
//   pressure  equations::compute_pressure      src/sph_engine.cpp:385-388
//   force     compute_pressure_force           src/sph_engine.cpp:390-412 (gradW: kernels.cpp:145-148, 96-108)
//             compute_viscosity_force          src/sph_engine.cpp:414-432 (laplacianW: kernels.cpp:150-153, 130-138)
//             compute_acceleration             src/sph_engine.cpp:438-443
// FAST: the same formulas in branch-free B-spline form ((2-q)+^3 - 4(1-q)+^3 and derivatives), FMA
// contraction allowed, MUFU rsqrt instead of sqrt + divide, per-neighbour quotients m/(2 rho) folded
// once per particle by the density kernel.  The neighbour inclusion test is the exact one in both
// modes so neighbour sets (and counts) never differ.
#pragma once

#include "sphb_internal.cuh"

namespace sphb {

__device__ __forceinline__ int cell_coord(float p, float inv_cell) { return __float2int_rd(__fmul_rn(p, inv_cell)); }
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
__device__ __forceinline__ float pick_axis(const float4& p, int axis) { return axis == 0 ? p.x : (axis == 1 ? p.y : p.z); }

// slab mode: is the density of this particle needed here (owned or first halo layer)?
__device__ __forceinline__ bool wants_density(const PairArgs& a, const float4& p) {
    if (a.slab_axis < 0) return true;
    const int cell = cell_coord(pick_axis(p, a.slab_axis), a.grid.ref_inv_cell);
    return cell >= a.rho_lo && cell < a.rho_hi;
}
__device__ __forceinline__ bool is_ghost(const float4& velid) { return (__float_as_uint(velid.w) & 0x80000000u) != 0u; }

// dot(d, d) <= r2 exactly as SpatialHash::within_radius_squared (reference spatial_hash.h:70-73):
// (dx*dx + dy*dy) + dz*dz with every product and sum rounded separately.
__device__ __forceinline__ float dist2_exact(float dx, float dy, float dz) {
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// ---- strict ---------------------------------------------------------------------------------------
__device__ __forceinline__ float w_strict(const PairConsts& k, float d2) {
    const float q = __fdiv_rn(__fsqrt_rn(d2), k.h);
    if (q >= 0.0f && q <= 1.0f) {
        // sigma * (2.0f/3.0f - q*q + 0.5f*q*q*q)
        const float a = __fsub_rn(2.0f / 3.0f, __fmul_rn(q, q));
        const float b = __fmul_rn(__fmul_rn(__fmul_rn(0.5f, q), q), q);
        return __fmul_rn(k.sigma, __fadd_rn(a, b));
    } else if (q > 1.0f && q <= 2.0f) {
        const float t = __fsub_rn(2.0f, q);
        return __fmul_rn(k.sigma, __fmul_rn(__fmul_rn(__fmul_rn(1.0f / 6.0f, t), t), t));
    }
    return 0.0f;
}

struct ForceAccum {
    float px, py, pz;   // pressure force
    float vx, vy, vz;   // viscosity force
};

// One neighbour j != i.  r = p_i - p_j, d2 = dot(r, r) (exact), u = v_j - v_i.
__device__ __forceinline__ void force_pair_strict(const PairConsts& k, ForceAccum& f, float rx, float ry, float rz, float d2,
                                                  float ux, float uy, float uz, float P_i, float m_j, float rho_j) {
    const float r_len = __fsqrt_rn(d2);
    const float q = __fdiv_rn(r_len, k.h);
    const float P_j = __fmul_rn(k.gas_constant, __fsub_rn(rho_j, k.rest_density));
    const bool in1 = (q >= 0.0f && q <= 1.0f);
    const bool in2 = (q > 1.0f && q <= 2.0f);
    const float t = __fsub_rn(2.0f, q);
    if (!(r_len < 1e-6f)) {
        const float pressure_term = __fdiv_rn(__fadd_rn(P_i, P_j), __fmul_rn(2.0f, rho_j));
        float gx = 0.0f, gy = 0.0f, gz = 0.0f;
        if (in1 || in2) {
            // sigma * (-2q + 1.5 q q)  |  -sigma * (0.5 t t)
            const float s = in1 ? __fmul_rn(k.sigma, __fadd_rn(__fmul_rn(-2.0f, q), __fmul_rn(__fmul_rn(1.5f, q), q)))
                                : __fmul_rn(-k.sigma, __fmul_rn(__fmul_rn(0.5f, t), t));
            gx = __fmul_rn(s, __fdiv_rn(rx, r_len));
            gy = __fmul_rn(s, __fdiv_rn(ry, r_len));
            gz = __fmul_rn(s, __fdiv_rn(rz, r_len));
        }
        gx = __fdiv_rn(gx, k.h); gy = __fdiv_rn(gy, k.h); gz = __fdiv_rn(gz, k.h);
        const float sc = __fmul_rn(m_j, pressure_term);
        f.px = __fsub_rn(f.px, __fmul_rn(sc, gx));
        f.py = __fsub_rn(f.py, __fmul_rn(sc, gy));
        f.pz = __fsub_rn(f.pz, __fmul_rn(sc, gz));
    }
    float lap = 0.0f;
    if (in1) lap = __fdiv_rn(__fmul_rn(k.sigma, __fadd_rn(-2.0f, __fmul_rn(3.0f, q))), k.h_sq);
    else if (in2) lap = __fdiv_rn(__fmul_rn(k.sigma, t), k.h_sq);
    const float sv = __fmul_rn(__fdiv_rn(m_j, rho_j), k.viscosity);
    f.vx = __fadd_rn(f.vx, __fmul_rn(__fmul_rn(sv, ux), lap));
    f.vy = __fadd_rn(f.vy, __fmul_rn(__fmul_rn(sv, uy), lap));
    f.vz = __fadd_rn(f.vz, __fmul_rn(__fmul_rn(sv, uz), lap));
}

// (Fp + Fv + (0, g, 0) * m) / m
__device__ __forceinline__ float4 accel_strict(const PairConsts& k, const ForceAccum& f, float m) {
    float4 a;
    a.x = __fdiv_rn(__fadd_rn(__fadd_rn(f.px, f.vx), __fmul_rn(0.0f, m)), m);
    a.y = __fdiv_rn(__fadd_rn(__fadd_rn(f.py, f.vy), __fmul_rn(k.gravity, m)), m);
    a.z = __fdiv_rn(__fadd_rn(__fadd_rn(f.pz, f.vz), __fmul_rn(0.0f, m)), m);
    a.w = 0.0f;
    return a;
}

// ---- the other kernel classes behind create_kernel (reference src/kernels.cpp:166-236) -------------
// The reference engine calls W / gradW / laplacianW through a Kernel pointer (sph_engine.cpp:209, 232, 236), so Wendland
// C2 and the Gaussian drop into the same density / force sums.  STRICT: the classes' own operation order, every op a
// round-to-nearest intrinsic — bit-identical for Wendland C2; the Gaussian calls expf, whose CUDA and glibc
// implementations differ in the last bit, so it is held to a tolerance instead (tests/test_gpu_parity.py).
// KT = kKernelCubic forwards to the functions above.
template <int KT>
__device__ __forceinline__ float w_strict_k(const PairConsts& k, float d2) {
    if (KT == kKernelCubic) return w_strict(k, d2);
    if (KT == kKernelGaussian)   // norm * exp(-r_sq * sigma_sq_inv)   (kernels.cpp:207-210)
        return __fmul_rn(k.gnorm, expf(__fmul_rn(-d2, k.gssi)));
    // Wendland C2 (kernels.cpp:171-177): q >= 2 -> 0; tmp = 1 - 0.5 q; norm * tmp * tmp * tmp * tmp * (2 q + 1)
    const float q = __fdiv_rn(__fsqrt_rn(d2), k.h);
    if (q >= 2.0f) return 0.0f;
    const float tmp = __fsub_rn(1.0f, __fmul_rn(0.5f, q));
    const float t4 = __fmul_rn(__fmul_rn(__fmul_rn(__fmul_rn(k.wnorm, tmp), tmp), tmp), tmp);
    return __fmul_rn(t4, __fadd_rn(__fmul_rn(2.0f, q), 1.0f));
}

template <int KT>
__device__ __forceinline__ void force_pair_strict_k(const PairConsts& k, ForceAccum& f, float rx, float ry, float rz, float d2,
                                                    float ux, float uy, float uz, float P_i, float m_j, float rho_j) {
    if (KT == kKernelCubic) { force_pair_strict(k, f, rx, ry, rz, d2, ux, uy, uz, P_i, m_j, rho_j); return; }
    const float r_len = __fsqrt_rn(d2);
    const float P_j = __fmul_rn(k.gas_constant, __fsub_rn(rho_j, k.rest_density));
    float gx = 0.0f, gy = 0.0f, gz = 0.0f, lap = 0.0f;
    if (KT == kKernelGaussian) {
        const float e = expf(__fmul_rn(-d2, k.gssi));
        // gradW = -2 s norm exp * r (kernels.cpp:212-216); laplacianW = 2 s norm exp (2 s r_sq - 3) (218-222)
        const float s = __fmul_rn(__fmul_rn(__fmul_rn(-2.0f, k.gssi), k.gnorm), e);
        gx = __fmul_rn(s, rx); gy = __fmul_rn(s, ry); gz = __fmul_rn(s, rz);
        lap = __fmul_rn(__fmul_rn(__fmul_rn(__fmul_rn(2.0f, k.gssi), k.gnorm), e),
                        __fsub_rn(__fmul_rn(__fmul_rn(2.0f, k.gssi), d2), 3.0f));
    } else {
        const float q = __fdiv_rn(r_len, k.h);
        if (!(q >= 2.0f)) {
            const float tmp = __fsub_rn(1.0f, __fmul_rn(0.5f, q));
            if (!(r_len < 1e-6f)) {   // gradW (kernels.cpp:179-190): norm * (-5 tmp tmp tmp q) * (r / (r_len h))
                const float dW_dq = __fmul_rn(__fmul_rn(__fmul_rn(__fmul_rn(-5.0f, tmp), tmp), tmp), q);
                const float s = __fmul_rn(k.wnorm, dW_dq), d = __fmul_rn(r_len, k.h);
                gx = __fmul_rn(s, __fdiv_rn(rx, d)); gy = __fmul_rn(s, __fdiv_rn(ry, d)); gz = __fmul_rn(s, __fdiv_rn(rz, d));
            }
            // laplacianW (kernels.cpp:192-198): norm * (5 / h_sq) * tmp * tmp * (5 q - 3)
            lap = __fmul_rn(__fmul_rn(__fmul_rn(__fmul_rn(k.wnorm, __fdiv_rn(5.0f, k.h_sq)), tmp), tmp),
                            __fsub_rn(__fmul_rn(5.0f, q), 3.0f));
        }
    }
    if (!(r_len < 1e-6f)) {   // compute_pressure_force's own guard (sph_engine.cpp:403)
        const float pressure_term = __fdiv_rn(__fadd_rn(P_i, P_j), __fmul_rn(2.0f, rho_j));
        const float sc = __fmul_rn(m_j, pressure_term);
        f.px = __fsub_rn(f.px, __fmul_rn(sc, gx));
        f.py = __fsub_rn(f.py, __fmul_rn(sc, gy));
        f.pz = __fsub_rn(f.pz, __fmul_rn(sc, gz));
    }
    const float sv = __fmul_rn(__fdiv_rn(m_j, rho_j), k.viscosity);
    f.vx = __fadd_rn(f.vx, __fmul_rn(__fmul_rn(sv, ux), lap));
    f.vy = __fadd_rn(f.vy, __fmul_rn(__fmul_rn(sv, uy), lap));
    f.vz = __fadd_rn(f.vz, __fmul_rn(__fmul_rn(sv, uz), lap));
}

// ---- fast -----------------------------------------------------------------------------------------
__device__ __forceinline__ float fast_sqrt(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_rsqrt(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// W(q)/sigma * 6 = (2-q)+^3 - 4 (1-q)+^3
__device__ __forceinline__ float w_fast(const PairConsts& k, float d2) {
    const float q = fast_sqrt(d2) * k.inv_h;
    const float t2 = fmaxf(2.0f - q, 0.0f);
    const float t1 = fmaxf(1.0f - q, 0.0f);
    return (k.sigma * (1.0f / 6.0f)) * (t2 * t2 * t2 - 4.0f * (t1 * t1 * t1));
}

// A_j = m_j / (2 rho_j), B_j = A_j * P_j (folded by the density kernel).
__device__ __forceinline__ void force_pair_fast(const PairConsts& k, ForceAccum& f, float rx, float ry, float rz, float d2,
                                                float ux, float uy, float uz, float P_i, float A_j, float B_j) {
    const bool apart = d2 >= 1e-12f;                 // r_len >= 1e-6 (sph_engine.cpp:403)
    const float inv_len = apart ? fast_rsqrt(d2) : 0.0f;
    const float q = (d2 * inv_len) * k.inv_h;
    const float t2 = fmaxf(2.0f - q, 0.0f);
    const float t1 = fmaxf(1.0f - q, 0.0f);
    // dW/dq / sigma = 2 (1-q)+^2 - 0.5 (2-q)+^2 ;  d2W/dq2 / sigma = (2-q)+ - 4 (1-q)+
    const float gq = 2.0f * (t1 * t1) - 0.5f * (t2 * t2);
    const float lq = t2 - 4.0f * t1;
    const float cp = (A_j * P_i + B_j) * (k.sig_h * gq * inv_len);
    f.px -= cp * rx; f.py -= cp * ry; f.pz -= cp * rz;
    const float cv = (2.0f * k.viscosity * k.sig_h2) * (A_j * lq);
    f.vx += cv * ux; f.vy += cv * uy; f.vz += cv * uz;
}

// Wendland C2 / Gaussian in fast math (FMA contraction, MUFU sqrt / rsqrt / ex2): W without the normalisation constant
// (the caller multiplies the sum once) and the pair force with the same folded inputs as force_pair_fast.
template <int KT>
__device__ __forceinline__ float w_fast_k(const PairConsts& k, float d2) {   // W / wnorm (Wendland), W / gnorm (Gaussian)
    if (KT == kKernelGaussian) return __expf(-d2 * k.gssi);
    const float q = fast_sqrt(d2) * k.inv_h;
    const float tmp = fmaxf(1.0f - 0.5f * q, 0.0f);   // (1 - q/2)+: exactly 0 beyond the support
    const float t2 = tmp * tmp;
    return (t2 * t2) * (2.0f * q + 1.0f);
}
template <int KT>
__device__ __forceinline__ float w_norm_k(const PairConsts& k) { return KT == kKernelGaussian ? k.gnorm : k.wnorm; }

template <int KT>
__device__ __forceinline__ void force_pair_fast_k(const PairConsts& k, ForceAccum& f, float rx, float ry, float rz, float d2,
                                                  float ux, float uy, float uz, float P_i, float A_j, float B_j) {
    const bool apart = d2 >= 1e-12f;                 // r_len >= 1e-6 (sph_engine.cpp:403)
    float gf, lap;                                   // gradW = gf * r, laplacianW = lap
    if (KT == kKernelGaussian) {
        const float e = k.gnorm * __expf(-d2 * k.gssi);
        gf = apart ? -2.0f * k.gssi * e : 0.0f;
        lap = 2.0f * k.gssi * e * (2.0f * k.gssi * d2 - 3.0f);
    } else {
        const float inv_len = apart ? fast_rsqrt(d2) : 0.0f;
        const float q = (d2 * inv_len) * k.inv_h;    // coincident: q = 0, like the reference's length 0
        const float tmp = fmaxf(1.0f - 0.5f * q, 0.0f);
        const float t2 = tmp * tmp;
        gf = (k.wnorm * k.inv_h) * (-5.0f * t2 * tmp * q) * inv_len;
        lap = (5.0f * k.wnorm * k.inv_h * k.inv_h) * t2 * (5.0f * q - 3.0f);
    }
    const float cp = (A_j * P_i + B_j) * gf;
    f.px -= cp * rx; f.py -= cp * ry; f.pz -= cp * rz;
    const float cv = (2.0f * k.viscosity) * (A_j * lap);
    f.vx += cv * ux; f.vy += cv * uy; f.vz += cv * uz;
}

__device__ __forceinline__ float4 accel_fast(const PairConsts& k, const ForceAccum& f, float m) {
    const float inv_m = 1.0f / m;
    float4 a;
    a.x = (f.px + f.vx) * inv_m;
    a.y = (f.py + f.vy) * inv_m + k.gravity;
    a.z = (f.pz + f.vz) * inv_m;
    a.w = 0.0f;
    return a;
}

}  // namespace sphb
