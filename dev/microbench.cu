// Pipe-rate micro-benchmarks on B200 (inline PTX so nothing is hoisted): how fast can the SM
// decide "all_k p_k > T_k" with DSETP / ISETP / mixed / DPX packed-16 / FSETP chains?
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#define ITERS 4096

// 4 chains x 8 compares per asm block = 32 compares
#define CH8(OP, TY, P, a0,a1,a2,a3,a4,a5,a6,a7) \
    "setp." OP "." TY " " P ", " a0 ", %9;\n\t" \
    "setp." OP ".and." TY " " P ", " a1 ", %10, " P ";\n\t" \
    "setp." OP ".and." TY " " P ", " a2 ", %11, " P ";\n\t" \
    "setp." OP ".and." TY " " P ", " a3 ", %12, " P ";\n\t" \
    "setp." OP ".and." TY " " P ", " a4 ", %13, " P ";\n\t" \
    "setp." OP ".and." TY " " P ", " a5 ", %14, " P ";\n\t" \
    "setp." OP ".and." TY " " P ", " a6 ", %15, " P ";\n\t" \
    "setp." OP ".and." TY " " P ", " a7 ", %16, " P ";\n\t"

__device__ __forceinline__ unsigned blk_d(const double (&p)[8], const double (&T)[8]) {
    unsigned r;
    asm volatile("{\n\t.reg .pred q0,q1,q2,q3;\n\t"
        CH8("gt","f64","q0","%1","%2","%3","%4","%5","%6","%7","%8")
        CH8("gt","f64","q1","%2","%3","%4","%5","%6","%7","%8","%1")
        CH8("gt","f64","q2","%3","%4","%5","%6","%7","%8","%1","%2")
        CH8("gt","f64","q3","%4","%5","%6","%7","%8","%1","%2","%3")
        "or.pred q0, q0, q1;\n\tor.pred q2, q2, q3;\n\tor.pred q0, q0, q2;\n\tselp.u32 %0, 1, 0, q0;\n\t}"
        : "=r"(r) : "d"(p[0]),"d"(p[1]),"d"(p[2]),"d"(p[3]),"d"(p[4]),"d"(p[5]),"d"(p[6]),"d"(p[7]),
                    "d"(T[0]),"d"(T[1]),"d"(T[2]),"d"(T[3]),"d"(T[4]),"d"(T[5]),"d"(T[6]),"d"(T[7]));
    return r;
}
__device__ __forceinline__ unsigned blk_i(const unsigned (&p)[8], const unsigned (&T)[8]) {
    unsigned r;
    asm volatile("{\n\t.reg .pred q0,q1,q2,q3;\n\t"
        CH8("ge","u32","q0","%1","%2","%3","%4","%5","%6","%7","%8")
        CH8("ge","u32","q1","%2","%3","%4","%5","%6","%7","%8","%1")
        CH8("ge","u32","q2","%3","%4","%5","%6","%7","%8","%1","%2")
        CH8("ge","u32","q3","%4","%5","%6","%7","%8","%1","%2","%3")
        "or.pred q0, q0, q1;\n\tor.pred q2, q2, q3;\n\tor.pred q0, q0, q2;\n\tselp.u32 %0, 1, 0, q0;\n\t}"
        : "=r"(r) : "r"(p[0]),"r"(p[1]),"r"(p[2]),"r"(p[3]),"r"(p[4]),"r"(p[5]),"r"(p[6]),"r"(p[7]),
                    "r"(T[0]),"r"(T[1]),"r"(T[2]),"r"(T[3]),"r"(T[4]),"r"(T[5]),"r"(T[6]),"r"(T[7]));
    return r;
}
__device__ __forceinline__ unsigned blk_f(const float (&p)[8], const float (&T)[8]) {
    unsigned r;
    asm volatile("{\n\t.reg .pred q0,q1,q2,q3;\n\t"
        CH8("gt","f32","q0","%1","%2","%3","%4","%5","%6","%7","%8")
        CH8("gt","f32","q1","%2","%3","%4","%5","%6","%7","%8","%1")
        CH8("gt","f32","q2","%3","%4","%5","%6","%7","%8","%1","%2")
        CH8("gt","f32","q3","%4","%5","%6","%7","%8","%1","%2","%3")
        "or.pred q0, q0, q1;\n\tor.pred q2, q2, q3;\n\tor.pred q0, q0, q2;\n\tselp.u32 %0, 1, 0, q0;\n\t}"
        : "=r"(r) : "f"(p[0]),"f"(p[1]),"f"(p[2]),"f"(p[3]),"f"(p[4]),"f"(p[5]),"f"(p[6]),"f"(p[7]),
                    "f"(T[0]),"f"(T[1]),"f"(T[2]),"f"(T[3]),"f"(T[4]),"f"(T[5]),"f"(T[6]),"f"(T[7]));
    return r;
}
// mixed: 16 f64 compares (2 chains) interleaved with 16 u32 compares (2 chains)
__device__ __forceinline__ unsigned blk_mix(const double (&p)[8], const double (&T)[8], const unsigned (&ph)[8], const unsigned (&Th)[8]) {
    unsigned r;
    asm volatile("{\n\t.reg .pred q0,q1,q2,q3;\n\t"
        "setp.gt.f64 q0, %1, %5;\n\t setp.ge.u32 q2, %9, %13;\n\t setp.gt.f64 q1, %2, %5;\n\t setp.ge.u32 q3, %10, %13;\n\t"
        "setp.gt.and.f64 q0, %2, %6, q0;\n\t setp.ge.and.u32 q2, %10, %14, q2;\n\t setp.gt.and.f64 q1, %3, %6, q1;\n\t setp.ge.and.u32 q3, %11, %14, q3;\n\t"
        "setp.gt.and.f64 q0, %3, %7, q0;\n\t setp.ge.and.u32 q2, %11, %15, q2;\n\t setp.gt.and.f64 q1, %4, %7, q1;\n\t setp.ge.and.u32 q3, %12, %15, q3;\n\t"
        "setp.gt.and.f64 q0, %4, %8, q0;\n\t setp.ge.and.u32 q2, %12, %16, q2;\n\t setp.gt.and.f64 q1, %1, %8, q1;\n\t setp.ge.and.u32 q3, %9, %16, q3;\n\t"
        "setp.gt.and.f64 q0, %1, %6, q0;\n\t setp.ge.and.u32 q2, %9, %14, q2;\n\t setp.gt.and.f64 q1, %2, %7, q1;\n\t setp.ge.and.u32 q3, %10, %15, q3;\n\t"
        "setp.gt.and.f64 q0, %2, %7, q0;\n\t setp.ge.and.u32 q2, %10, %15, q2;\n\t setp.gt.and.f64 q1, %3, %8, q1;\n\t setp.ge.and.u32 q3, %11, %16, q3;\n\t"
        "setp.gt.and.f64 q0, %3, %8, q0;\n\t setp.ge.and.u32 q2, %11, %16, q2;\n\t setp.gt.and.f64 q1, %4, %5, q1;\n\t setp.ge.and.u32 q3, %12, %13, q3;\n\t"
        "setp.gt.and.f64 q0, %4, %5, q0;\n\t setp.ge.and.u32 q2, %12, %13, q2;\n\t setp.gt.and.f64 q1, %1, %6, q1;\n\t setp.ge.and.u32 q3, %9, %14, q3;\n\t"
        "or.pred q0, q0, q1;\n\tor.pred q2, q2, q3;\n\tor.pred q0, q0, q2;\n\tselp.u32 %0, 1, 0, q0;\n\t}"
        : "=r"(r) : "d"(p[0]),"d"(p[1]),"d"(p[2]),"d"(p[3]),"d"(T[0]),"d"(T[1]),"d"(T[2]),"d"(T[3]),
                    "r"(ph[0]),"r"(ph[1]),"r"(ph[2]),"r"(ph[3]),"r"(Th[0]),"r"(Th[1]),"r"(Th[2]),"r"(Th[3]));
    return r;
}
// DPX: 16 VIADDMNMX.S16x2 = 32 element-compares, 4 independent accumulators
__device__ __forceinline__ unsigned blk_dpx(const unsigned (&p)[8], const unsigned (&T)[8]) {
    unsigned a0 = 0x7fff7fffu, a1 = a0, a2 = a0, a3 = a0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        asm volatile("{ .reg .u32 t;\n\t"
                     "vadd2.s32.s32.s32 t, %4, %8, 0; \n\t}" : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3) : "r"(p[k]), "r"(p[k + 1]), "r"(p[k + 2]), "r"(p[k + 3]), "r"(T[k]));
    }
    return a0 ^ a1 ^ a2 ^ a3;
}

template <int KIND>
__global__ void __launch_bounds__(256) k_reg(const double* __restrict__ in, unsigned* out, int iters) {
    double T[8], p[8]; unsigned Th[8], ph[8]; float Tf[8], pf[8];
    for (int i = 0; i < 8; i++) { T[i] = in[64 + ((threadIdx.x + i) & 63)]; p[i] = in[(threadIdx.x * 3 + i) & 63];
        Th[i] = (unsigned)(__double_as_longlong(T[i]) >> 32); ph[i] = (unsigned)(__double_as_longlong(p[i]) >> 32); Tf[i] = (float)T[i]; pf[i] = (float)p[i]; }
    unsigned cnt = 0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {   // make every operand loop-variant (ptxas hoists invariant setp's): 8 integer adds per 32 compares
            if (KIND == 0 || KIND == 2 || KIND == 6) p[k] = __hiloint2double(__double2hiint(p[k]), __double2loint(p[k]) + it);
            if (KIND == 1 || KIND == 2 || KIND == 3 || KIND >= 6) ph[k] += it;
            if (KIND == 5) pf[k] = __uint_as_float(__float_as_uint(pf[k]) + it);
        }
        if (KIND == 0) cnt += blk_d(p, T);
        else if (KIND == 1) cnt += blk_i(ph, Th);
        else if (KIND == 2) cnt += blk_mix(p, T, ph, Th);
        else if (KIND == 3) { // DPX 16x2 via intrinsic, volatile-ish through data dependence
            unsigned a0 = cnt | 0x7fff0000u, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
#pragma unroll
            for (int k = 0; k < 4; k++) { a0 = __viaddmin_s16x2(ph[k], Th[k], a0); a1 = __viaddmin_s16x2(ph[k+1], Th[k], a1); a2 = __viaddmin_s16x2(ph[k+2], Th[k], a2); a3 = __viaddmin_s16x2(ph[k+3], Th[k], a3); }
            cnt += (a0 ^ a1 ^ a2 ^ a3) & 1; }
        else if (KIND == 5) cnt += blk_f(pf, Tf);
        else if (KIND == 6) {   // 16 DSETP (2 chains of 8, one asm block) + 16 VIADDMNMX.S16x2 (4 accumulators) interleaved by the scheduler
            unsigned a0 = cnt | 0x7fff0000u, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
            unsigned r;
            asm volatile("{\n\t.reg .pred q0,q1;\n\t"
                "setp.gt.f64 q0, %1, %9;\n\t setp.gt.f64 q1, %2, %9;\n\t"
                "setp.gt.and.f64 q0, %2, %10, q0;\n\t setp.gt.and.f64 q1, %3, %10, q1;\n\t"
                "setp.gt.and.f64 q0, %3, %11, q0;\n\t setp.gt.and.f64 q1, %4, %11, q1;\n\t"
                "setp.gt.and.f64 q0, %4, %12, q0;\n\t setp.gt.and.f64 q1, %5, %12, q1;\n\t"
                "setp.gt.and.f64 q0, %5, %13, q0;\n\t setp.gt.and.f64 q1, %6, %13, q1;\n\t"
                "setp.gt.and.f64 q0, %6, %14, q0;\n\t setp.gt.and.f64 q1, %7, %14, q1;\n\t"
                "setp.gt.and.f64 q0, %7, %15, q0;\n\t setp.gt.and.f64 q1, %8, %15, q1;\n\t"
                "setp.gt.and.f64 q0, %8, %16, q0;\n\t setp.gt.and.f64 q1, %1, %16, q1;\n\t"
                "or.pred q0, q0, q1;\n\tselp.u32 %0, 1, 0, q0;\n\t}"
                : "=r"(r) : "d"(p[0]),"d"(p[1]),"d"(p[2]),"d"(p[3]),"d"(p[4]),"d"(p[5]),"d"(p[6]),"d"(p[7]),
                            "d"(T[0]),"d"(T[1]),"d"(T[2]),"d"(T[3]),"d"(T[4]),"d"(T[5]),"d"(T[6]),"d"(T[7]));
#pragma unroll
            for (int k = 0; k < 4; k++) { a0 = __viaddmin_s16x2(ph[k], Th[k], a0); a1 = __viaddmin_s16x2(ph[k+1], Th[k], a1); a2 = __viaddmin_s16x2(ph[k+2], Th[k], a2); a3 = __viaddmin_s16x2(ph[k+3], Th[k], a3); }
            cnt += r + ((a0 ^ a1 ^ a2 ^ a3) & 1); }
        else if (KIND == 7) {   // 32 VIADDMNMX.S16x2 on 8 independent accumulators
            unsigned a[8];
#pragma unroll
            for (int j = 0; j < 8; j++) a[j] = (cnt + j) | 0x7fff0000u;
#pragma unroll
            for (int k = 0; k < 4; k++)
#pragma unroll
                for (int j = 0; j < 8; j++) a[j] = __viaddmin_s16x2(ph[(k + j) & 7], Th[k], a[j]);
            unsigned x = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) x ^= a[j];
            cnt += x & 1; }
        else if (KIND == 8) {   // 32 x (IADD + IMNMX as 32-bit viaddmin) on 8 accumulators
            int a[8];
#pragma unroll
            for (int j = 0; j < 8; j++) a[j] = (int)((cnt + j) | 0x7fff0000u);
#pragma unroll
            for (int k = 0; k < 4; k++)
#pragma unroll
                for (int j = 0; j < 8; j++) a[j] = __viaddmin_s32((int)ph[(k + j) & 7], (int)Th[k], a[j]);
            int x = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) x ^= a[j];
            cnt += x & 1; }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = cnt;
}
// with operands streamed from shared memory (broadcast LDS.128), like the real kernel
template <int KIND>
__global__ void __launch_bounds__(256) k_lds(const double* __restrict__ in, unsigned* out, int iters) {
    __shared__ __align__(16) double sp[1024];
    for (int i = threadIdx.x; i < 1024; i += 256) sp[i] = in[i & 63];
    __syncthreads();
    double T[8]; unsigned Th[8];
    for (int i = 0; i < 8; i++) { T[i] = in[64 + ((threadIdx.x + i) & 63)]; Th[i] = (unsigned)(__double_as_longlong(T[i]) >> 32); }
    unsigned cnt = 0;
    const double2* s2 = reinterpret_cast<const double2*>(sp);
    for (int it = 0; it < iters; it++) {
        double p[8]; unsigned ph[8];
        const double2* q = s2 + ((it * 4) & 511);
#pragma unroll
        for (int j = 0; j < 4; j++) { double2 v = q[j]; p[2*j] = v.x; p[2*j+1] = v.y; }
#pragma unroll
        for (int j = 0; j < 8; j++) ph[j] = (unsigned)(__double_as_longlong(p[j]) >> 32);
        if (KIND == 0) cnt += blk_d(p, T);
        else if (KIND == 1) cnt += blk_i(ph, Th);
        else cnt += blk_mix(p, T, ph, Th);
    }
    if (cnt == 0xffffffffu) out[0] = cnt;
}

// LDS broadcast cost: W = bytes per lane per load (4, 8, 16), uniform address
template <int W>
__global__ void __launch_bounds__(256) k_ldsonly(const double* __restrict__ in, unsigned* out, int iters) {
    __shared__ __align__(16) unsigned sp[4096];
    for (int i = threadIdx.x; i < 4096; i += 256) sp[i] = (unsigned)i;
    __syncthreads();
    unsigned acc = 0;
    for (int it = 0; it < iters; it++) {
        const unsigned base = (it * 64) & 4095 & ~63u;
#pragma unroll
        for (int j = 0; j < 16; j++) {
            if (W == 4) { unsigned v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(sp + base + j))); acc ^= v; }
            else if (W == 8) { unsigned v0, v1; asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v0), "=r"(v1) : "r"((unsigned)__cvta_generic_to_shared(sp + base + 2 * j))); acc ^= v0 ^ v1; }
            else { unsigned v0, v1, v2, v3; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"((unsigned)__cvta_generic_to_shared(sp + base + 4 * j))); acc ^= v0 ^ v1 ^ v2 ^ v3; }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
// per-lane distinct addresses (conflict-free), for comparison
template <int W>
__global__ void __launch_bounds__(256) k_ldsvec(const double* __restrict__ in, unsigned* out, int iters) {
    __shared__ __align__(16) unsigned sp[8192];
    for (int i = threadIdx.x; i < 8192; i += 256) sp[i] = (unsigned)i;
    __syncthreads();
    unsigned acc = 0;
    const int lane = threadIdx.x & 31;
    for (int it = 0; it < iters; it++) {
        const unsigned base = ((it * 128) & 4095);
#pragma unroll
        for (int j = 0; j < 16; j++) {
            if (W == 4) { unsigned v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(sp + base + lane + 32 * (j & 3)))); acc ^= v; }
            else { unsigned v0, v1, v2, v3; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"((unsigned)__cvta_generic_to_shared(sp + base + 4 * lane + 128 * (j & 3)))); acc ^= v0 ^ v1 ^ v2 ^ v3; }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <typename F> float timeit(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; r++) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
    return best;
}
int main() {
    double h[128]; for (int i = 0; i < 128; i++) h[i] = 0.5 + i * 0.001;
    double* d; unsigned* o; cudaMalloc(&d, sizeof h); cudaMalloc(&o, 4u << 20); cudaMemcpy(d, h, sizeof h, cudaMemcpyHostToDevice);
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    const int sms = pr.multiProcessorCount;
    for (int warps_per_sm : {8, 16, 32}) {
        const int grid = sms * warps_per_sm / 8;
        const double cmp = (double)grid * 256 * ITERS * 32;
        printf("--- %d warps/SM  (cmp/clk/SM assumes 1.965 GHz; FP64 pipe nominal = 32)\n", warps_per_sm);
        auto rep = [&](const char* n, float ms) { printf("%-34s %.3f ms  %.2f Tcmp/s  %.1f cmp/clk/SM\n", n, ms, cmp / ms / 1e9, cmp / (ms * 1e-3) / sms / 1.965e9); };
        rep("REG DSETP chain", timeit([&] { k_reg<0><<<grid, 256>>>(d, o, ITERS); }));
        rep("REG ISETP.u32 chain", timeit([&] { k_reg<1><<<grid, 256>>>(d, o, ITERS); }));
        rep("REG mixed DSETP/ISETP 1:1", timeit([&] { k_reg<2><<<grid, 256>>>(d, o, ITERS); }));
        rep("REG VIADDMNMX.S16x2 (2 el/instr)", timeit([&] { k_reg<3><<<grid, 256>>>(d, o, ITERS); }));
        rep("REG 16 DSETP + 16 DPX16x2 (48 el)", timeit([&] { k_reg<6><<<grid, 256>>>(d, o, ITERS); }) * (32.0f / 48.0f));
        rep("REG DPX16x2 8 accs (64 el)", timeit([&] { k_reg<7><<<grid, 256>>>(d, o, ITERS); }) * 0.5f);
        rep("REG VIADDMNMX.S32 8 accs (32 el)", timeit([&] { k_reg<8><<<grid, 256>>>(d, o, ITERS); }));
        rep("REG FSETP chain", timeit([&] { k_reg<5><<<grid, 256>>>(d, o, ITERS); }));
        {
            const double lds = (double)grid * 8 * ITERS * 16;   // warp-level LDS instructions per launch
            auto repl = [&](const char* n, float ms) { printf("%-34s %.3f ms  %.2f warp-LDS/clk/SM\n", n, ms, lds / (ms * 1e-3) / sms / 1.965e9); };
            repl("LDS.32 uniform", timeit([&] { k_ldsonly<4><<<grid, 256>>>(d, o, ITERS); }));
            repl("LDS.64 uniform", timeit([&] { k_ldsonly<8><<<grid, 256>>>(d, o, ITERS); }));
            repl("LDS.128 uniform", timeit([&] { k_ldsonly<16><<<grid, 256>>>(d, o, ITERS); }));
            repl("LDS.32 per-lane", timeit([&] { k_ldsvec<4><<<grid, 256>>>(d, o, ITERS); }));
            repl("LDS.128 per-lane", timeit([&] { k_ldsvec<16><<<grid, 256>>>(d, o, ITERS); }));
        }
        rep("LDS.128 + DSETP chain", timeit([&] { k_lds<0><<<grid, 256>>>(d, o, ITERS); }));
        rep("LDS.128 + ISETP chain", timeit([&] { k_lds<1><<<grid, 256>>>(d, o, ITERS); }));
        rep("LDS.128 + mixed", timeit([&] { k_lds<2><<<grid, 256>>>(d, o, ITERS); }));
    }
    return 0;
}
