// Issue-rate microbenchmark of the instructions the pair kernels are made of (sm_100a): each kernel runs a long
// unrolled chain of ONE instruction kind on independent registers, 8 warps per SMSP, and reports warp-instructions per
// clock per SMSP.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; run on one GPU.
#include <cstdio>
#include <cuda_runtime.h>

#define ITER 2048
#define NREG 8

template <int OP>
__global__ void __launch_bounds__(1024, 1) k(float* out, float a, float b, int iters) {
    float2 r[NREG];
#pragma unroll
    for (int i = 0; i < NREG; ++i) r[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i);
    const float2 A = make_float2(a, a * 1.0001f), B = make_float2(b, b * 0.9999f);
    unsigned m = threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NREG; ++i) {
            if (OP == 0) { r[i].x = fmaf(r[i].x, A.x, B.x); r[i].y = fmaf(r[i].y, A.y, B.y); }                 // 2 scalar FFMA (3 registers)
            if (OP == 1) r[i] = __ffma2_rn(r[i], A, B);                                                           // 1 FFMA2
            if (OP == 2) r[i] = __fadd2_rn(make_float2(A.x, A.x), r[i]);                                          // FADD2 with broadcast operand
            if (OP == 3) { asm volatile("fma.rn.sat.f32 %0, %0, %1, %2;" : "+f"(r[i].x) : "f"(A.x), "f"(B.x)); asm volatile("fma.rn.sat.f32 %0, %0, %1, %2;" : "+f"(r[i].y) : "f"(A.y), "f"(B.y)); }
            if (OP == 4) { r[i].x = fmaxf(r[i].x * 1.0f, A.x); r[i].y = fminf(r[i].y, B.y); asm volatile("" : "+f"(r[i].x), "+f"(r[i].y)); }  // FMNMX
            if (OP == 5) { asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(r[i].x)); asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(r[i].y)); }
            if (OP == 6) { unsigned x = __float_as_uint(r[i].x); asm volatile("shf.l.wrap.b32 %0, %1, %0, 1;" : "+r"(m) : "r"(x)); asm volatile("shf.l.wrap.b32 %0, %1, %0, 1;" : "+r"(x) : "r"(m)); r[i].x = __uint_as_float(x); }
            if (OP == 7) { r[i].x = r[i].x + A.x; r[i].y = r[i].y + A.y; asm volatile("" : "+f"(r[i].x), "+f"(r[i].y)); }  // 2 scalar FADD
            if (OP == 8) r[i] = __fmul2_rn(r[i], A);
            if (OP == 9) { r[i] = __ffma2_rn(r[i], A, B); r[(i + 1) % NREG].x = fmaxf(r[(i + 1) % NREG].x, A.x); asm volatile("" : "+f"(r[(i+1)%NREG].x)); }  // FFMA2 + FMNMX mix
            if (OP == 10) { r[i] = __ffma2_rn(r[i], A, B); r[(i + 1) % NREG].x = fmaf(r[(i + 1) % NREG].x, A.x, B.x); }  // FFMA2 + FFMA mix
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NREG; ++i) s += r[i].x + r[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + m;
}

template <int OP>
void run(const char* name, int instr_per_inner, float* d) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = 148;
    k<OP><<<blocks, 1024>>>(d, 1.0001f, 1e-3f, 16);
    cudaEventRecord(e0);
    k<OP><<<blocks, 1024>>>(d, 1.0001f, 1e-3f, ITER);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double warp_instr_per_smsp = (double)ITER * NREG * instr_per_inner * 8;   // 32 warps per SM = 8 per SMSP
    const double cycles = ms * 1e-3 * clk * 1e3;
    printf("%-34s %8.3f ms  %.3f warp-instr/clk/SMSP (at %d MHz nominal)\n", name, ms, warp_instr_per_smsp / cycles, clk / 1000);
}

int main() {
    float* d; cudaMalloc(&d, 148 * 1024 * sizeof(float));
    run<0>("FFMA x2 (3-reg scalar)", 2, d);
    run<1>("FFMA2", 1, d);
    run<2>("FADD2 (broadcast operand)", 1, d);
    run<3>("FFMA.SAT x2", 2, d);
    run<4>("FMNMX x2 (+FMUL)", 3, d);
    run<5>("MUFU.SQRT x2", 2, d);
    run<6>("SHF x2", 2, d);
    run<7>("FADD x2", 2, d);
    run<8>("FMUL2", 1, d);
    run<9>("FFMA2 + FMNMX", 2, d);
    run<10>("FFMA2 + FFMA", 2, d);
    cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
