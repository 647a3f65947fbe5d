// hvb200_kernels.cu — sm_100a kernels of the hv_approx hot path and their launchers.
//
// Path (reference file:line, relative to /root/reference):
//   direction generators   c/hvapprox.c:288-304,551-576,610-647 (DZ2019-HW), :239-250 (DZ2019-MC
//                          post-processing), :791-826 (Rphi-FWE+ mapping)
//   max-min + s^d          c/hvapprox.c:68-117 (get_expected_value), c/pow_int.h:16-74
//   sum                    c/hvapprox.c:251,681,857
//
// Layout in HBM (per plan = per GPU):
//   pts   [n_tiles][TP][d] FP64, transformed points p' = |ref - p| > 0, zero-padded to whole tiles
//   dirs  [batch][2][d]    FP64, per sample the reciprocal direction r and winv = RN(1/r)
//   chunk/CTA partial sums, one double per CTA; (hi, lo) result pair
//
// Kernels:
//   k_gen_hw      one thread per sample: x87-exact lattice point (integer emulation of the
//                 reference's long double arithmetic), d-1 Newton solves in FP64, polar->cartesian
//   k_mt_generate, k_zig_*   (hvb200_mcgen.cuh) the reference's MT19937+ziggurat normal stream, bit-exact
//   k_gen_mc      normals -> |z|, clamp, norm, r = norm/|z|
//   k_gen_rphi    u (host fixed-point recurrence) -> Fang-Wang map -> r = 1/w
//   k_prep_r      uploaded r -> winv
//   k_maxmin      persistent CTAs; a producer warp streams point tiles into shared memory with
//                 TMA bulk copies (cp.async.bulk + mbarrier, multi-stage); 224 consumer threads
//                 own R directions each and sweep every tile.
//                   EXACT : d DMUL + d min per sample*point, 1 max             (2d+1 FP64 ops)
//                   SCREEN: per element one DSETP against a per-direction threshold
//                           T_k < best/r_k (predicate chain); only candidates that may raise the
//                           running max are evaluated with the EXACT arithmetic.  max/min are
//                           exact and order-free, so both give bit-identical s.
//   k_maxmin_dpx  the same screen on conservatively quantised 16-bit keys with DPX VIADDMNMX.S16x2
//                 (two elements per instruction, no predicates); default for n' >= 256
//   k_quantize    point keys for k_maxmin_dpx
//   k_reduce      fixed-shape pairwise tree over CTA partial sums in double-double -> (hi, lo)
//
// No tensor cores: max-min is a semiring, not a contraction.  Compiled with -fmad=false so that
// a*b+c is never contracted where the reference (x86-64, no FMA) rounds twice; fma() is written
// explicitly where wanted.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <chrono>
#include <thread>
#include <mutex>
#include <algorithm>

#include <nvtx3/nvToolsExt.h>     // header-only: ranges are no-ops unless a profiler injects its library

#include "hvb200_tables.h"
#include "hvb200_internal.h"

typedef unsigned long long ull;

// NVTX range of the enclosing scope (plan creation, preparation, direction generators, max-min, reduction): the
// phases of one estimate show up by name on an Nsight Systems / Nsight Compute timeline.
struct NvtxScope {
    explicit NvtxScope(const char *name) { nvtxRangePushA(name); }
    ~NvtxScope() { nvtxRangePop(); }
};
#define HVB200_NVTX(name) NvtxScope nvtx_scope_(name)

#define CUDA_TRY(expr)                                                                       \
    do {                                                                                     \
        cudaError_t e_ = (expr);                                                             \
        if (e_ != cudaSuccess) {                                                             \
            hvb200_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_),         \
                             __FILE__, __LINE__);                                            \
            return -1;                                                                       \
        }                                                                                    \
    } while (0)

// ------------------------------------------------------------------------------------------------
// constants
// ------------------------------------------------------------------------------------------------
static constexpr int kConsumerWarps = 7;                         // 7 + producer = 8 warps: 2 CTAs/SM at 128 regs
static constexpr int kConsumerThreads = kConsumerWarps * 32;
static constexpr int kBlockThreads = kConsumerThreads + 32;   // + one producer warp
static constexpr int kStages = 3;
static constexpr int kMaxStageBytes = 24 * 1024;
static constexpr int kNewtonCap = 200;
static constexpr int kMaxFmTerms = 32;

__constant__ double c_fm_lead[HVB200_MAXD];
__constant__ double c_fm_h[HVB200_MAXD][kMaxFmTerms];
__constant__ int c_fm_nh[HVB200_MAXD];
__constant__ double c_Im[HVB200_MAXD];
__constant__ ull c_theta0_bits[32];

// x^e multiplication trees: same shapes as c/pow_int.h:22-56 (see oracle/hvapprox_oracle.c for
// the derivation of the table); step t computes v[t+1] = v[a]*v[b].
struct PowProg { unsigned char n; unsigned char op[7][2]; };
static constexpr PowProg kPowProg[32] = {
    {0, {}}, {0, {}},
    {1, {{0,0}}},
    {2, {{0,0},{1,0}}},
    {2, {{0,0},{1,1}}},
    {3, {{0,0},{1,1},{2,0}}},
    {3, {{0,0},{1,1},{2,1}}},
    {4, {{0,0},{1,1},{2,1},{3,0}}},
    {3, {{0,0},{1,1},{2,2}}},
    {4, {{0,0},{1,0},{2,2},{3,2}}},
    {4, {{0,0},{1,1},{2,2},{3,1}}},
    {5, {{0,0},{1,1},{2,2},{3,1},{4,0}}},
    {4, {{0,0},{1,1},{2,2},{3,2}}},
    {5, {{0,0},{1,1},{2,2},{3,2},{4,0}}},
    {5, {{0,0},{1,1},{2,2},{3,2},{4,1}}},
    {5, {{0,0},{1,1},{2,0},{3,3},{4,3}}},
    {4, {{0,0},{1,1},{2,2},{3,3}}},
    {5, {{0,0},{1,1},{2,2},{3,3},{4,0}}},
    {5, {{0,0},{1,0},{2,2},{3,2},{4,4}}},
    {6, {{0,0},{1,0},{2,2},{3,2},{4,4},{5,0}}},
    {5, {{0,0},{1,1},{2,2},{3,1},{4,4}}},
    {6, {{0,0},{1,1},{2,2},{3,1},{4,4},{5,0}}},
    {6, {{0,0},{1,1},{2,0},{3,3},{4,4},{5,1}}},
    {6, {{0,0},{1,0},{1,2},{3,3},{4,4},{5,2}}},
    {5, {{0,0},{1,1},{2,2},{3,2},{4,4}}},
    {6, {{0,0},{1,1},{2,2},{3,2},{4,4},{5,0}}},
    {6, {{0,0},{1,1},{2,2},{3,2},{4,0},{5,5}}},
    {6, {{0,0},{1,0},{2,2},{3,3},{4,4},{5,2}}},
    {6, {{0,0},{1,1},{2,2},{3,2},{4,1},{5,5}}},
    {7, {{0,0},{1,1},{2,2},{3,2},{4,1},{5,5},{6,0}}},
    {6, {{0,0},{1,1},{2,0},{3,3},{4,3},{5,5}}},
    {7, {{0,0},{1,1},{2,0},{3,3},{4,3},{5,5},{6,0}}},
};

template <int E>
__device__ __forceinline__ double pow_chain(double x)
{
    if constexpr (E == 0) {
        return 1.0;
    } else if constexpr (E <= 31) {
        constexpr PowProg P = kPowProg[E];
        double v[8];
        v[0] = x;
#pragma unroll
        for (int t = 0; t < P.n; t++) v[t + 1] = v[P.op[t][0]] * v[P.op[t][1]];
        return v[P.n];
    } else {   // generic square-and-multiply, c/pow_int.h:65-73
        double acc = (E & 1) ? x : 1.0;
#pragma unroll
        for (unsigned e = (unsigned)E >> 1; e; e >>= 1) {
            x *= x;
            if (e & 1u) acc *= x;
        }
        return acc;
    }
}

// runtime exponent (1..31) through the same tabulated multiplication programs
__device__ __forceinline__ double pow_chain_runtime(double x, int e)
{
    switch (e) {
#define HVB200_PC(E) case E: return pow_chain<E>(x);
        HVB200_PC(1) HVB200_PC(2) HVB200_PC(3) HVB200_PC(4) HVB200_PC(5) HVB200_PC(6) HVB200_PC(7) HVB200_PC(8)
        HVB200_PC(9) HVB200_PC(10) HVB200_PC(11) HVB200_PC(12) HVB200_PC(13) HVB200_PC(14) HVB200_PC(15) HVB200_PC(16)
        HVB200_PC(17) HVB200_PC(18) HVB200_PC(19) HVB200_PC(20) HVB200_PC(21) HVB200_PC(22) HVB200_PC(23) HVB200_PC(24)
        HVB200_PC(25) HVB200_PC(26) HVB200_PC(27) HVB200_PC(28) HVB200_PC(29) HVB200_PC(30) HVB200_PC(31)
#undef HVB200_PC
    }
    return 1.0;
}

// ------------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + 1-D TMA bulk copy (SASS: SYNCS.*, UBLKCP)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(ull *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(ull *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(ull *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(ull *bar, unsigned parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// Producer-side wait: the producer is 3 stages ahead, so it polls with a back-off instead of
// spinning (a spinning lane costs its SM sub-partition ~25 % of its issue slots).
__device__ __forceinline__ void mbar_wait_backoff(ull *bar, unsigned parity)
{
    unsigned done = 0;
    for (;;) {
        // try_wait with a suspend-time hint: the hardware parks the thread until the phase completes
        // or ~20 us pass, instead of returning after its short default time limit
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity), "r"(20000u) : "memory");
        if (done) break;
        __nanosleep(2000);
    }
}
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gsrc, unsigned bytes, ull *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void consumer_bar_sync()
{   // named barrier 1 over the consumer threads only (the producer warp never joins)
    asm volatile("bar.sync 1, %0;" ::"n"(kConsumerThreads) : "memory");
}

// ------------------------------------------------------------------------------------------------
// K1: max-min
// ------------------------------------------------------------------------------------------------
struct MaxminArgs {
    const double *pts;       // [n_tiles][tile_points][D]
    int n_tiles;
    int tile_points;
    int stage_bytes;         // bytes reserved per pipeline stage (>= tile bytes, multiple of 128)
    int staged;              // 1: per-chunk directions are staged in shared memory
    const double *dirs;      // [count][2][D]
    ull count;
    double *cta_sums;        // [gridDim.x], accumulated across launches
    double *values;          // optional [count]
    ull *slow_counter;       // optional diagnostic
};

enum { MODE_EXACT = 1, MODE_SCREEN = 2, MODE_DPX = 3, MODE_BP = 4 };

// Threshold for the screen: a T with  p <= T  =>  RN(p * r) <= best.
// winv = RN(1/r)  =>  best*winv*(1-2^-50) < (best/r)(1-2^-51); tiny values flush to 0 (= "always a
// candidate"), which is always safe.
__device__ __forceinline__ double screen_threshold(double best, double winv)
{
    const double t0 = best * winv;
    const double t = fma(t0, -0x1p-50, t0);
    return (t > 0x1p-900) ? t : 0.0;
}

template <int D, bool STAGED>
__device__ __forceinline__ void screen_slow_update(const double *__restrict__ p,       // smem, D doubles
                                                   const double *__restrict__ dir,     // r | winv (smem or global)
                                                   double &best, double (&T)[D])
{
    auto ld = [&](int i) { return STAGED ? dir[i] : __ldg(dir + i); };
    double m = p[0] * ld(0);
#pragma unroll
    for (int k = 1; k < D; k++) {
        const double t = p[k] * ld(k);
        m = (t < m) ? t : m;
    }
    if (m > best) {
        best = m;
#pragma unroll
        for (int k = 0; k < D; k++) T[k] = screen_threshold(m, ld(D + k));
    }
}

template <int D, int R, int G, int MODE, int MINB>
__global__ void __launch_bounds__(kBlockThreads, MINB) k_maxmin(const MaxminArgs a)
{
    static_assert((G * D) % 2 == 0, "a group of G points must be a whole number of 16-byte words");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr ull CH = (ull)kConsumerThreads * R;          // samples per chunk
    const int tile_doubles = a.tile_points * D;
    const unsigned tile_bytes = (unsigned)tile_doubles * 8u;
    const int stage_doubles = a.stage_bytes / 8;
    double *tiles = reinterpret_cast<double *>(smem_raw);
    double *dstage = tiles + (size_t)kStages * stage_doubles;                           // [CH][2D] if staged
    ull *full_bar = reinterpret_cast<ull *>(dstage + (a.staged ? CH * 2 * D : 0));
    ull *empty_bar = full_bar + kStages;
    ull *dfull_bar = empty_bar + kStages;
    ull *dempty_bar = dfull_bar + 1;
    double *red = reinterpret_cast<double *>(dempty_bar + 1);   // 8 doubles

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kStages; s++) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], kConsumerThreads / 32);
        }
        mbar_init(dfull_bar, 1);
        mbar_init(dempty_bar, kConsumerThreads / 32);
        mbar_fence_init();
    }
    __syncthreads();

    const ull n_chunks = (a.count + CH - 1) / CH;
    const bool resident = a.n_tiles <= kStages;           // whole point set fits: load once

    if (warp == kConsumerThreads / 32) {
        // ---------------- producer warp: TMA bulk copies of direction blocks and point tiles ----------------
        if (lane == 0) {
            ull q = 0, it = 0;
            for (ull chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x, it++) {
                if (a.staged) {
                    if (it > 0) mbar_wait_backoff(dempty_bar, (unsigned)((it - 1) & 1));
                    const ull first = chunk * CH;
                    const ull rows = (a.count - first < CH) ? a.count - first : CH;
                    const unsigned bytes = (unsigned)(rows * 2 * D * 8);
                    mbar_arrive_expect_tx(dfull_bar, bytes);
                    tma_load_1d(dstage, a.dirs + first * (2 * D), bytes, dfull_bar);
                }
                if (resident && it > 0) continue;
                for (int t = 0; t < a.n_tiles; t++, q++) {
                    const int stage = (int)(q % kStages);
                    if (q >= (ull)kStages) mbar_wait_backoff(&empty_bar[stage], (unsigned)(((q / kStages) - 1) & 1));
                    mbar_arrive_expect_tx(&full_bar[stage], tile_bytes);
                    tma_load_1d(tiles + (size_t)stage * stage_doubles, a.pts + (size_t)t * tile_doubles, tile_bytes,
                                &full_bar[stage]);
                }
            }
        }
        return;
    }

    // ---------------- consumers ----------------
    double acc = 0.0;
    ull slow = 0;
    ull q = 0, it = 0;
    for (ull chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x, it++) {
        ull sample[R];
        bool valid[R];
        double best[R];
        double T[R][D];      // SCREEN: thresholds;  EXACT: reciprocal direction r
        const double *dir[R];
        if (a.staged) mbar_wait(dfull_bar, (unsigned)(it & 1));
#pragma unroll
        for (int r = 0; r < R; r++) {
            sample[r] = chunk * CH + (ull)r * kConsumerThreads + tid;
            valid[r] = sample[r] < a.count;
            best[r] = 0.0;
            dir[r] = a.staged ? dstage + (size_t)(r * kConsumerThreads + tid) * (2 * D) : a.dirs + sample[r] * (2 * D);
            if (MODE == MODE_EXACT) {
#pragma unroll
                for (int k = 0; k < D; k++) T[r][k] = valid[r] ? (a.staged ? dir[r][k] : __ldg(dir[r] + k)) : 0.0;
            } else {
#pragma unroll
                for (int k = 0; k < D; k++) T[r][k] = valid[r] ? 0.0 : __longlong_as_double(0x7ff0000000000000LL);
            }
        }

        for (int t = 0; t < a.n_tiles; t++, q++) {
            const int stage = resident ? t : (int)(q % kStages);
            const unsigned parity = resident ? 0u : (unsigned)((q / kStages) & 1);
            mbar_wait(&full_bar[stage], parity);
            const double *tile = tiles + (size_t)stage * stage_doubles;
            const double2 *tile2 = reinterpret_cast<const double2 *>(tile);

#pragma unroll 1
            for (int g0 = 0; g0 < a.tile_points; g0 += G) {
                const double2 *gp = tile2 + (g0 * D) / 2;
                if (MODE == MODE_EXACT) {
                    // The textbook loop, every product formed and compared: per element one DMUL and one DSETP.  The
                    // running min of c/hvapprox.c:84-100 is a select (DSETP + 2 FSEL per element, which made the loop
                    // issue-bound at 53 % FP64-pipe utilisation); max_p min_k (p_k r_k) only needs to know whether
                    // EVERY product of a point exceeds the running max, so the comparisons chain through one
                    // predicate per (direction, point) — DSETP.GT.AND — and the min itself is formed for the few
                    // points that raise the max (~ln n per direction).  Same products, same comparisons, same bits.
                    bool c[R][G];
#pragma unroll
                    for (int r = 0; r < R; r++)
#pragma unroll
                        for (int g = 0; g < G; g++) c[r][g] = true;
#pragma unroll
                    for (int j = 0; j < (G * D) / 2; j++) {
                        // large D: keep the scheduler from hoisting every load and product of the group ahead of the
                        // comparisons (r[D] already fills the register file): shared-memory loads do not cross the fence
                        if (D > 16 && j > 0 && j % 8 == 0) asm volatile("" ::: "memory");
                        const double2 v = gp[j];
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            const int e = 2 * j + h, g = e / D, k = e % D;
                            const double pv = h ? v.y : v.x;
#pragma unroll
                            for (int r = 0; r < R; r++) c[r][g] = c[r][g] & (pv * T[r][k] > best[r]);
                        }
                    }
                    bool any = false;
#pragma unroll
                    for (int r = 0; r < R; r++)
#pragma unroll
                        for (int g = 0; g < G; g++) any = any | c[r][g];
                    if (any) {
#pragma unroll
                        for (int g = 0; g < G; g++)
#pragma unroll
                            for (int r = 0; r < R; r++)
                                if (c[r][g]) {
                                    const double *p = tile + (size_t)(g0 + g) * D;
                                    double m = p[0] * T[r][0];
#pragma unroll
                                    for (int k = 1; k < D; k++) {
                                        const double t = p[k] * T[r][k];
                                        m = (t < m) ? t : m;
                                    }
                                    best[r] = (m > best[r]) ? m : best[r];
                                    slow++;
                                }
                    }
                } else {
                    bool c[R][G];
#pragma unroll
                    for (int r = 0; r < R; r++)
#pragma unroll
                        for (int g = 0; g < G; g++) c[r][g] = true;
#pragma unroll
                    for (int j = 0; j < (G * D) / 2; j++) {
                        const double2 v = gp[j];
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            const int e = 2 * j + h, g = e / D, k = e % D;
                            const double pv = h ? v.y : v.x;
#pragma unroll
                            for (int r = 0; r < R; r++) c[r][g] = c[r][g] & (pv > T[r][k]);
                        }
                    }
                    bool any = false;
#pragma unroll
                    for (int r = 0; r < R; r++)
#pragma unroll
                        for (int g = 0; g < G; g++) any = any | c[r][g];
                    if (any) {
#pragma unroll
                        for (int g = 0; g < G; g++)
#pragma unroll
                            for (int r = 0; r < R; r++)
                                if (c[r][g]) {
                                    if (a.staged) screen_slow_update<D, true>(tile + (size_t)(g0 + g) * D, dir[r], best[r], T[r]);
                                    else screen_slow_update<D, false>(tile + (size_t)(g0 + g) * D, dir[r], best[r], T[r]);
                                    slow++;
                                }
                    }
                }
            }
            if (!resident) {
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_bar[stage]);
            }
        }
        if (a.staged) {
            __syncwarp();
            if (lane == 0) mbar_arrive(dempty_bar);
        }

#pragma unroll
        for (int r = 0; r < R; r++) {
            const double v = pow_chain<D>(best[r]);
            if (valid[r]) {
                acc += v;
                if (a.values) a.values[sample[r]] = v;
            }
        }
    }

    // fixed-shape block reduction over the 256 consumer threads
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, off);
    if (lane == 0) red[warp] = acc;
    if (a.slow_counter) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) slow += __shfl_down_sync(0xffffffffu, slow, off);
        if (lane == 0 && slow) atomicAdd(a.slow_counter, slow);
    }
    consumer_bar_sync();
    if (tid == 0) {
        double v[8];
#pragma unroll
        for (int w = 0; w < 8; w++) v[w] = (w < kConsumerWarps) ? red[w] : 0.0;
        a.cta_sums[blockIdx.x] += ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
    }
}

// ------------------------------------------------------------------------------------------------
// K1-DPX: the same exact screen on packed 16-bit keys (DPX VIADDMNMX.S16x2)
//
// The DSETP screen is limited by the SM's predicate-compare rate (~57 compares/clk/SM measured,
// whatever the operand type).  VIADDMNMX.S16x2 does min(acc, a+b) on two 16-bit lanes without
// touching predicates (45 instr = 90 element-compares/clk/SM measured).  Per objective k the
// points' coordinates and the thresholds are mapped through the SAME monotone map
//     f_k(x) = (x - lo_k) * scale_k ,  lo_k/hi_k = range of p'_k over the point set, scale_k = 32000/(hi_k-lo_k)
// to keys  KP = clamp(ceil f_k(p), 0, S)  (rounded up)  and  KT = floor f_k(T)  (rounded down;
// 0 if T < lo_k, S+1 if T >= hi_k), so that  p > T  =>  KP - KT >= 0.  A point is a candidate iff
// min_k (KP_k - KT_k) >= 0; candidates take the same exact FP64 slow path as the DSETP screen, so the
// per-sample result keeps the reference's bits.  One 32-bit word holds the keys of two points
// (point 2j in the low half, 2j+1 in the high half) for one objective; the threshold key is
// duplicated in both halves.  Padding points carry KP = -1 (never candidates).
// ------------------------------------------------------------------------------------------------
static constexpr int kKeyScale = 32000;
static constexpr int kDpxPivots = 64;
static int dpx_pivots()
{
    const char *e = getenv("HVB200_PIVOTS");
    const int v = e ? atoi(e) : kDpxPivots;
    return v < 0 ? 0 : (v > 1024 ? 1024 : v);
}

struct DpxArgs {
    const unsigned *keys;    // [n_tiles][tile_pairs][D] packed words
    int n_tiles;
    int tile_pairs;
    int stage_bytes;
    ull npoints;
    int n_pivots;            // first points of the set, evaluated exactly for every direction as a warm-up
    const double *pts;       // FP64 points, row-major [npoints_padded][D] (slow path)
    const double *dirs;      // [count][2][D]
    ull count;
    double *cta_sums;
    double *values;
    ull *slow_counter;
    double qscale[32], qlos[32];     // key map: scale_k and lo_k*scale_k*(1+2^-20)+2^-10 (see dpx_neg_key)
};

// Threshold key of objective k for a running max `best_m` = best*(1-2^-20):
//   x  = best_m * winv_k * scale_k - (lo_k*scale_k*(1+2^-20) + 2^-10)      <=  f_k(T_k) - margin,
//   KT = clamp(floor(x), 0, S+1)
// where T_k = best*winv_k*(1-2^-50) is the FP64 screen threshold.  The margin (>= 2^-21 relative,
// 2^-10 absolute in key units) dominates every rounding error of the two evaluations of f_k (each
// <= 2^-50 relative), so KT <= floor f_k(T_k) and  p > T_k  =>  KP = ceil f_k(p) >= KT  still holds; it
// only admits a few more false candidates.  Branch-free: 2 FP64 ops, one conversion, one clamp.
__device__ __forceinline__ unsigned dpx_neg_key(double best_m, double winv, double scale, double lo_s)
{
    const double x = fma(best_m * winv, scale, -lo_s);
    int kt = __double2int_rd(x);                  // NaN -> 0, +-inf saturate
    kt = max(0, min(kt, kKeyScale + 1));
    return ((unsigned)(-kt) & 0xffffu) * 0x10001u;
}

__global__ void __launch_bounds__(256) k_quantize(const double *__restrict__ pts, ull n, int d, int tile_pairs,
                                                  ull total_pairs, const double *__restrict__ qlo,
                                                  const double *__restrict__ qhi, const double *__restrict__ qscale,
                                                  unsigned *__restrict__ keys)
{
    const ull w = (ull)blockIdx.x * blockDim.x + threadIdx.x;      // word index = pair * d + k
    if (w >= total_pairs * (ull)d) return;
    const ull pair = w / d;
    const int k = (int)(w - pair * d);
    unsigned word = 0;
    for (int h = 0; h < 2; h++) {
        const ull idx = 2 * pair + h;
        unsigned key = 0xffffu;                                      // -1: padding, never a candidate
        if (idx < n) {
            const double p = pts[idx * d + k];
            double f = ceil((p - qlo[k]) * qscale[k]);
            f = f < 0.0 ? 0.0 : (f > (double)kKeyScale ? (double)kKeyScale : f);
            key = (unsigned)(int)f;
        }
        word |= key << (16 * h);
    }
    keys[w] = word;
    (void)tile_pairs; (void)qhi;
}

template <int D>
__device__ __forceinline__ void dpx_refresh_keys(double best, const double *__restrict__ winv, unsigned (&negT)[D],
                                                 const DpxArgs &a)
{
    const double best_m = best * (1.0 - 0x1p-20);
    if constexpr (D % 2 == 0) {
        const double2 *w2 = reinterpret_cast<const double2 *>(winv);       // rows of 2D doubles: 16-byte aligned
#pragma unroll
        for (int k = 0; k < D; k += 2) {
            const double2 w = __ldg(w2 + k / 2);
            negT[k] = dpx_neg_key(best_m, w.x, a.qscale[k], a.qlos[k]);
            negT[k + 1] = dpx_neg_key(best_m, w.y, a.qscale[k + 1], a.qlos[k + 1]);
        }
    } else {
#pragma unroll
        for (int k = 0; k < D; k++) negT[k] = dpx_neg_key(best_m, __ldg(winv + k), a.qscale[k], a.qlos[k]);
    }
}
template <int D>
__device__ __forceinline__ void dpx_slow_update(const double *__restrict__ p,      // global, D doubles
                                                const double *__restrict__ dir,    // global r | winv
                                                double &best, unsigned (&negT)[D], const DpxArgs &a)
{
    double m;
    if constexpr (D % 2 == 0) {
        const double2 *p2 = reinterpret_cast<const double2 *>(p), *r2 = reinterpret_cast<const double2 *>(dir);
        const double2 p0 = __ldg(p2), r0 = __ldg(r2);
        m = p0.x * r0.x;
        const double t1 = p0.y * r0.y;
        m = (t1 < m) ? t1 : m;
#pragma unroll
        for (int k = 2; k < D; k += 2) {
            const double2 pv = __ldg(p2 + k / 2), rv = __ldg(r2 + k / 2);
            const double ta = pv.x * rv.x, tb = pv.y * rv.y;
            m = (ta < m) ? ta : m;
            m = (tb < m) ? tb : m;
        }
    } else {
        m = __ldg(p) * __ldg(dir);
#pragma unroll
        for (int k = 1; k < D; k++) {
            const double t = __ldg(p + k) * __ldg(dir + k);
            m = (t < m) ? t : m;
        }
    }
    if (m > best) {
        best = m;
        dpx_refresh_keys<D>(m, dir + D, negT, a);
    }
}

template <int D, int R, int G, int MINB>
__global__ void __launch_bounds__(kBlockThreads, MINB) k_maxmin_dpx(const __grid_constant__ DpxArgs a)
{
    static_assert((G * D) % 4 == 0, "a group of G point pairs must be a whole number of 16-byte words");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr ull CH = (ull)kConsumerThreads * R;
    const int tile_words = a.tile_pairs * D;
    const unsigned tile_bytes = (unsigned)tile_words * 4u;
    const int stage_words = a.stage_bytes / 4;
    unsigned *tiles = reinterpret_cast<unsigned *>(smem_raw);
    ull *full_bar = reinterpret_cast<ull *>(smem_raw + (size_t)kStages * a.stage_bytes);
    ull *empty_bar = full_bar + kStages;
    double *red = reinterpret_cast<double *>(empty_bar + kStages + 2);
    double *pivots = red + 8;                                   // [n_pivots][D] FP64: the first points of the set

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kStages; s++) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], kConsumerThreads / 32);
        }
        mbar_fence_init();
    }
    const int n_piv = a.n_pivots;
    for (int i = tid; i < n_piv * D; i += kBlockThreads) pivots[i] = a.pts[i];
    __syncthreads();

    const ull n_chunks = (a.count + CH - 1) / CH;
    const ull my_chunks = (n_chunks > blockIdx.x) ? (n_chunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const bool resident = a.n_tiles <= kStages;
    const ull total_loads = resident ? (my_chunks ? (ull)a.n_tiles : 0) : my_chunks * (ull)a.n_tiles;

    if (warp == kConsumerThreads / 32) {
        if (lane == 0) {
            for (ull q = 0; q < total_loads; q++) {
                const int stage = (int)(q % kStages);
                if (q >= (ull)kStages) mbar_wait_backoff(&empty_bar[stage], (unsigned)(((q / kStages) - 1) & 1));
                mbar_arrive_expect_tx(&full_bar[stage], tile_bytes);
                tma_load_1d(tiles + (size_t)stage * stage_words, a.keys + (size_t)(q % (ull)a.n_tiles) * tile_words,
                            tile_bytes, &full_bar[stage]);
            }
        }
        return;
    }

    double acc = 0.0;
    ull slow = 0;
    ull q = 0;
    for (ull chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
        ull sample[R];
        bool valid[R];
        double best[R];
        unsigned negT[R][D];
#pragma unroll
        for (int r = 0; r < R; r++) {
            sample[r] = chunk * CH + (ull)r * kConsumerThreads + tid;
            valid[r] = sample[r] < a.count;
            best[r] = 0.0;
            if (valid[r]) {
                // warm-up: exact max-min over the pivot points (branch-free, all lanes busy) so that the
                // screen starts from a useful running max instead of paying a divergent slow-path entry
                // for each of the first ~ln(n_piv) records of every direction
                const double *dir = a.dirs + sample[r] * (2 * D);
                double rr[D];
#pragma unroll
                for (int k = 0; k < D; k++) rr[k] = __ldg(dir + k);
                double b = 0.0;
#pragma unroll 2
                for (int i = 0; i < n_piv; i++) {
                    const double *pp = pivots + i * D;
                    double m = pp[0] * rr[0];
#pragma unroll
                    for (int k = 1; k < D; k++) {
                        const double t = pp[k] * rr[k];
                        m = (t < m) ? t : m;
                    }
                    b = (m > b) ? m : b;
                }
                best[r] = b;
                dpx_refresh_keys<D>(b, dir + D, negT[r], a);
            } else {
                // tail samples beyond count get the "nothing passes" key
                const unsigned v = ((unsigned)(-(kKeyScale + 1)) & 0xffffu) * 0x10001u;
#pragma unroll
                for (int k = 0; k < D; k++) negT[r][k] = v;
            }
        }

        for (int t = 0; t < a.n_tiles; t++, q++) {
            const int stage = resident ? t : (int)(q % kStages);
            const unsigned parity = resident ? 0u : (unsigned)((q / kStages) & 1);
            mbar_wait(&full_bar[stage], parity);
            const uint4 *tile4 = reinterpret_cast<const uint4 *>(tiles + (size_t)stage * stage_words);
            const ull tile_first_point = (ull)t * a.tile_pairs * 2;

#pragma unroll 1
            for (int g0 = 0; g0 < a.tile_pairs; g0 += G) {
                const uint4 *gp = tile4 + (g0 * D) / 4;
                unsigned am[R][G];
#pragma unroll
                for (int r = 0; r < R; r++)
#pragma unroll
                    for (int g = 0; g < G; g++) am[r][g] = 0x7fff7fffu;
#pragma unroll
                for (int j = 0; j < (G * D) / 4; j++) {
                    const uint4 w = gp[j];
#pragma unroll
                    for (int h = 0; h < 4; h++) {
                        const int e = 4 * j + h, g = e / D, k = e % D;
                        const unsigned wh = h == 0 ? w.x : (h == 1 ? w.y : (h == 2 ? w.z : w.w));
#pragma unroll
                        for (int r = 0; r < R; r++) am[r][g] = __viaddmin_s16x2(wh, negT[r][k], am[r][g]);
                    }
                }
                // any halfword >= 0 ?  (sign bit clear)
                unsigned all_neg = 0x80008000u;
#pragma unroll
                for (int r = 0; r < R; r++)
#pragma unroll
                    for (int g = 0; g < G; g++) all_neg &= am[r][g];
                if (all_neg != 0x80008000u) {
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        // candidate bits of direction r: bit (2g+h) <=> point 2(g0+g)+h
                        unsigned cand = 0;
#pragma unroll
                        for (int g = 0; g < G; g++) {
                            const unsigned nm = ~am[r][g];
                            cand |= ((nm >> 15) & 1u) << (2 * g);
                            cand |= ((nm >> 31) & 1u) << (2 * g + 1);
                        }
                        while (cand) {
                            const int bit = __ffs((int)cand) - 1;
                            cand &= cand - 1;
                            const ull idx = tile_first_point + 2 * (ull)g0 + (ull)bit;
                            dpx_slow_update<D>(a.pts + idx * D, a.dirs + sample[r] * (2 * D), best[r], negT[r], a);
                            slow++;
                        }
                    }
                }
            }
            if (!resident) {
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_bar[stage]);
            }
        }

#pragma unroll
        for (int r = 0; r < R; r++) {
            const double v = pow_chain<D>(best[r]);
            if (valid[r]) {
                acc += v;
                if (a.values) a.values[sample[r]] = v;
            }
        }
    }

#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, off);
    if (lane == 0) red[warp] = acc;
    if (a.slow_counter) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) slow += __shfl_down_sync(0xffffffffu, slow, off);
        if (lane == 0 && slow) atomicAdd(a.slow_counter, slow);
    }
    consumer_bar_sync();
    if (tid == 0) {
        double v[8];
#pragma unroll
        for (int w = 0; w < 8; w++) v[w] = (w < kConsumerWarps) ? red[w] : 0.0;
        a.cta_sums[blockIdx.x] += ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
    }
}


// ------------------------------------------------------------------------------------------------
// Point ordering for the screens.  max over points is order-free (exact), so the processing order
// is a pure performance knob: a screen pays a divergent exact evaluation every time a point raises a
// direction's running max, and with the points in input order that happens ~ln n times per direction.
// k_probe_argmax evaluates a few thousand throw-away random directions exactly and counts how often
// each point is the arg-max; sorting the points by that count (most frequent winners first) makes the
// first 64 points (the warm-up set) contain the winner of most directions: measured 4.9 -> 0.9 later
// records per direction at n = 1e4, d = 10.  The probe directions need no parity with anything.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned hash_u32(unsigned x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
__global__ void __launch_bounds__(128) k_probe_argmax(const double *__restrict__ pts, unsigned n, int d, unsigned nprobe,
                                                      unsigned *__restrict__ wins)
{
    // one warp per probe direction; the lanes stride over the points
    const unsigned j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned lane = threadIdx.x & 31u;
    if (j >= nprobe) return;
    double r[HVB200_MAXD];
    for (int k = 0; k < d; k++) {
        // |N(0,1)| by Box-Muller on two hashed uniforms; only the direction matters, not its norm
        const unsigned h1 = hash_u32(j * 2654435761u + 2u * k + 1u), h2 = hash_u32(h1 ^ (0x9e3779b9u + k));
        const float u1 = (h1 >> 8) * (1.0f / 16777216.0f) + (0.5f / 16777216.0f), u2 = (h2 >> 8) * (1.0f / 16777216.0f);
        const float z = sqrtf(-2.0f * __logf(u1)) * __cosf(6.2831853f * u2);
        r[k] = 1.0 / ((double)fabsf(z) + 1e-6);
    }
    double best = -1.0;
    unsigned arg = 0xffffffffu;
    for (unsigned i = lane; i < n; i += 32) {
        const double *p = pts + (size_t)i * d;
        double m = __ldg(p) * r[0];
        for (int k = 1; k < d; k++) {
            const double t = __ldg(p + k) * r[k];
            m = (t < m) ? t : m;
        }
        if (m > best) { best = m; arg = i; }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const double ob = __shfl_down_sync(0xffffffffu, best, off);
        const unsigned oa = __shfl_down_sync(0xffffffffu, arg, off);
        if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
    }
    if (lane == 0 && arg != 0xffffffffu) atomicAdd(&wins[arg], 1u);
}
__global__ void __launch_bounds__(256) k_gather_rows(const double *__restrict__ src, const unsigned *__restrict__ order,
                                                     ull n, int d, double *__restrict__ dst)
{
    const ull w = (ull)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n * (ull)d) return;
    const ull i = w / d;
    const int k = (int)(w - i * d);
    dst[w] = src[(ull)order[i] * d + k];
}

// ------------------------------------------------------------------------------------------------
// K5: final reduction, double-double pairwise tree of fixed shape (1 CTA, 1024 threads)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dd_add(double &hi, double &lo, double bh, double bl)
{   // (hi,lo) += (bh,bl), Knuth two-sum + renormalisation
    const double s = hi + bh;
    const double bb = s - hi;
    const double err = (hi - (s - bb)) + (bh - bb);
    const double l = err + (lo + bl);
    hi = s + l;
    lo = l - (hi - s);
}
__global__ void __launch_bounds__(1024) k_reduce(const double *__restrict__ cta_sums, int n, double *out_hilo)
{
    __shared__ double sh[1024], sl[1024];
    const int tid = threadIdx.x;
    double hi = 0.0, lo = 0.0;
    for (int i = tid; i < n; i += 1024) dd_add(hi, lo, cta_sums[i], 0.0);
    sh[tid] = hi; sl[tid] = lo;
    __syncthreads();
    for (int off = 512; off > 0; off >>= 1) {
        if (tid < off) {
            double h = sh[tid], l = sl[tid];
            dd_add(h, l, sh[tid + off], sl[tid + off]);
            sh[tid] = h; sl[tid] = l;
        }
        __syncthreads();
    }
    if (tid == 0) { out_hilo[0] = sh[0]; out_hilo[1] = sl[0]; }
}
// Long runs of the pruned kernel leave one partial sum per (chunk, warp) — 1.6 M of them for 1e8 directions — which one
// CTA would walk for half a millisecond.  Stage 1: CTA g reduces the fixed slice [g S, (g + 1) S) to one (hi, lo) pair
// with the same tree; stage 2 adds the pairs.  The shape depends on n only, so the result stays deterministic.
static constexpr int kReduceSlice = 1 << 15;
__global__ void __launch_bounds__(1024) k_reduce_slices(const double *__restrict__ sums, ull n, double *__restrict__ pairs)
{
    __shared__ double sh[1024], sl[1024];
    const int tid = threadIdx.x;
    const ull i0 = (ull)blockIdx.x * kReduceSlice, i1 = min(n, i0 + (ull)kReduceSlice);
    double hi = 0.0, lo = 0.0;
    for (ull i = i0 + tid; i < i1; i += 1024) dd_add(hi, lo, sums[i], 0.0);
    sh[tid] = hi; sl[tid] = lo;
    __syncthreads();
    for (int off = 512; off > 0; off >>= 1) {
        if (tid < off) {
            double h = sh[tid], l = sl[tid];
            dd_add(h, l, sh[tid + off], sl[tid + off]);
            sh[tid] = h; sl[tid] = l;
        }
        __syncthreads();
    }
    if (tid == 0) { pairs[2 * blockIdx.x] = sh[0]; pairs[2 * blockIdx.x + 1] = sl[0]; }
}
__global__ void __launch_bounds__(1024) k_reduce_pairs(const double *__restrict__ pairs, int n, double *out_hilo)
{
    __shared__ double sh[1024], sl[1024];
    const int tid = threadIdx.x;
    double hi = 0.0, lo = 0.0;
    for (int i = tid; i < n; i += 1024) dd_add(hi, lo, pairs[2 * i], pairs[2 * i + 1]);
    sh[tid] = hi; sl[tid] = lo;
    __syncthreads();
    for (int off = 512; off > 0; off >>= 1) {
        if (tid < off) {
            double h = sh[tid], l = sl[tid];
            dd_add(h, l, sh[tid + off], sl[tid + off]);
            sh[tid] = h; sl[tid] = l;
        }
        __syncthreads();
    }
    if (tid == 0) { out_hilo[0] = sh[0]; out_hilo[1] = sl[0]; }
}

// ------------------------------------------------------------------------------------------------
// K2: DZ2019-HW directions
// ------------------------------------------------------------------------------------------------
struct HwDev {
    ull N;
    ull Nrecip;            // floor((2^64 - 1) / N): exact 64/32 division by a multiply-high and <= 2 corrections
    ull a[HVB200_MAXD];
    int d;
};
// floor(t / N) for t < N * 2^32 with the reciprocal above (the quotient estimate is at most 2 too small)
__device__ __forceinline__ ull div_by_n(ull t, ull N, ull Nrecip, ull &rem)
{
    ull q = __umul64hi(t, Nrecip);
    ull r = t - q * N;
    while (r >= N) { r -= N; q++; }
    rem = r;
    return q;
}

// RN64(x / N) as M * 2^e with M in [2^63, 2^64): bit-exact emulation of the x87 division
// (i+1) / (long double)nsamples at c/hvapprox.c:295.  Requires 0 < x < N <= 2^32.
struct X87Val { ull M; int e; };
__device__ __forceinline__ X87Val x87_div(ull x, ull N, ull Nrecip)
{
    const int lx = 63 - __clzll((long long)x), lN = 63 - __clzll((long long)N);
    int s = lN - lx;
    ull X = x << s;
    if (X < N) { X <<= 1; s++; }            // N <= X < 2N  ->  X/N in [1,2)
    const ull r0 = X - N;
    const ull t1 = r0 << 32;
    ull r1, r2;
    const ull q1 = div_by_n(t1, N, Nrecip, r1);
    const ull t2 = r1 << 31;
    const ull q2 = div_by_n(t2, N, Nrecip, r2);
    X87Val f;
    f.M = (1ull << 63) + (q1 << 31) + q2;
    f.e = -63 - s;
    const ull twice = r2 << 1;
    if (twice > N || (twice == N && (f.M & 1ull))) {
        f.M += 1ull;
        if (f.M == 0ull) { f.M = 1ull << 63; f.e += 1; }
    }
    return f;
}
// (double) fractl( RN64(f * a) ): c/hvapprox.c:297-298 followed by the cast of :615.
__device__ __forceinline__ double x87_mul_frac(X87Val f, ull a)
{
    const ull lo = f.M * a;
    const ull hi = __umul64hi(f.M, a);
    ull Mv;
    int ev;
    if (hi == 0ull) {
        Mv = lo; ev = f.e;                  // fits 64 bits: exact
    } else {
        int sh = 64 - __clzll((long long)hi);
        Mv = (hi << (64 - sh)) | (lo >> sh);
        const ull dropped = lo & ((1ull << sh) - 1ull), half = 1ull << (sh - 1);
        if (dropped > half || (dropped == half && (Mv & 1ull))) {
            Mv += 1ull;
            if (Mv == 0ull) { Mv = 1ull << 63; sh += 1; }
        }
        ev = f.e + sh;
    }
    if (ev >= 0) return 0.0;
    const int fbits = -ev;
    const ull fr = (fbits >= 64) ? Mv : (Mv & ((1ull << fbits) - 1ull));
    if (fr == 0ull) return 0.0;
    return __ull2double_rn(fr) * __longlong_as_double((long long)(1023 - fbits) << 52);
}

// F_m(b) = int_0^b sin^m via the reduction formula (same restatement as the oracle):
//   even m: lead*b - cos*sin*H(sin^2);  odd m: lead*(1-cos) - cos*sin^2*H(sin^2)
// sin and cos for the first quadrant, where every angle of the solve lives: reflect to [0, pi/4] about pi/4 and
// evaluate the two minimax kernels of that interval (the classic fdlibm coefficients, < 1 ulp) with explicit FMAs.
// No quadrant bookkeeping, no large-argument path: a third of the library call's instructions.
#ifndef HVB200_HW_LIBM_SINCOS
__device__ __forceinline__ void sincos_q1(double x, double *s, double *c)
{
    if (!(x >= 0.0 && x <= 1.5707963267948966192)) { sincos(x, s, c); return; }
    const bool refl = x > 0.78539816339744830962;
    const double y = refl ? (1.5707963267948966192 - x) + 6.123233995736766e-17 : x;
    const double z = y * y;
    double ps = 1.58969099521155010221e-10;
    ps = fma(ps, z, -2.50507602534068634195e-08);
    ps = fma(ps, z, 2.75573137070700676789e-06);
    ps = fma(ps, z, -1.98412698298579493134e-04);
    ps = fma(ps, z, 8.33333333332248946124e-03);
    ps = fma(ps, z, -1.66666666666666324348e-01);
    const double sy = fma(y * z, ps, y);
    double pc = -1.13596475577881948265e-11;
    pc = fma(pc, z, 2.08757232129817482790e-09);
    pc = fma(pc, z, -2.75573143513906633035e-07);
    pc = fma(pc, z, 2.48015872894767294178e-05);
    pc = fma(pc, z, -1.38888888888741095749e-03);
    pc = fma(pc, z, 4.16666666666666019037e-02);
    const double hz = 0.5 * z;
    const double w = 1.0 - hz;
    const double cy = w + (((1.0 - w) - hz) + z * z * pc);
    *s = refl ? cy : sy;
    *c = refl ? sy : cy;
}
#else
__device__ __forceinline__ void sincos_q1(double x, double *s, double *c) { sincos(x, s, c); }
#endif
__device__ __forceinline__ double fm_eval(int m, double b, double &sb, double &cb)
{
    sincos_q1(b, &sb, &cb);
    if (m == 0) return b;
    if (m == 1) return 1.0 - cb;
    const double s2 = sb * sb;
    const int nh = c_fm_nh[m];
    double H = c_fm_h[m][nh - 1];
    for (int t = nh - 2; t >= 0; t--) H = fma(H, s2, c_fm_h[m][t]);
    if (m & 1) return fma(c_fm_lead[m], 1.0 - cb, -(cb * s2) * H);
    return fma(c_fm_lead[m], b, -(cb * sb) * H);
}
__device__ __forceinline__ double ipow(double x, int e)
{
    double acc = (e & 1) ? x : 1.0;
    for (e >>= 1; e; e >>= 1) { x *= x; if (e & 1) acc *= x; }
    return acc;
}
// solve_inverse_int_of_power_sin (c/hvapprox.c:551-564) in FP64; returns sin/cos of the angle.
// Tabulated inverse theta = F_m^{-1}(u * I_m) on a uniform grid in u (HVB200_HWTAB intervals per exponent m, default 2048;
// 16384 and 131072 measured the same generator time on B200),
// built once per device by k_hw_table_init with the reference-style Newton.  k_gen_hw uses the linear
// interpolant only as the STARTING point of the same Newton iteration with the same stop rule
// |F(x) - t| <= 1e-15 (c/hvapprox.c:558), which cuts the mean iteration count from 4.2 to about 2 and
// makes it nearly uniform across a warp.  The first intervals (theta ~ u^(1/(m+1)) has unbounded slope
// at 0) and zero targets keep the reference's start at pi/2.
static constexpr int kHwTabDefault = 2048;
static int g_hw_ntab = 0;                            // intervals per exponent (HVB200_HWTAB, fixed at the first plan)

__device__ void hw_solve(double target, double u, int m, const double *__restrict__ tab, int ntab,
                         double &s_out, double &c_out, unsigned &capped)
{
    if (target == 0.0 && m < 32 && c_theta0_bits[m] != ~0ull) {
        sincos_q1(__longlong_as_double((long long)c_theta0_bits[m]), &s_out, &c_out);
        return;
    }
    if (m == 0) {
        // F_0(b) = b: the reference's single Newton step x = pi/2 - (pi/2 - t) returns t up to a
        // long double rounding (1e-19); in FP64 the same step would lose t below 1e-16 absolute.
        sincos_q1(target, &s_out, &c_out);
        return;
    }
    if (m == 1) {
        // F_1(b) = 1 - cos b = t has the closed form cos b = 1 - t, sin b = sqrt(t (2 - t)): exact up to rounding
        // (1 - t is exact for t >= 0.5, and below that its rounding error is under 1.2e-16 in F)
        c_out = 1.0 - target;
        s_out = sqrt(target * (2.0 - target));
        return;
    }
    const double HALF_PI = 1.57079632679489661923;
    double x = HALF_PI;
    double sb = 1.0, cb = 6.123233995736766e-17;     // sin, cos of (double)(pi/2)
    double f = c_Im[m] - target;
    if (tab != nullptr && fabs(f) > 1e-15) {
        const double g = u * ntab;
        const int gi = (int)g;
        if (gi >= 4 && gi < ntab) {
            const double *row = tab + (size_t)m * (ntab + 1);
            const double a = __ldg(row + gi), b = __ldg(row + gi + 1);
            x = a + (b - a) * (g - gi);
            f = fm_eval(m, x, sb, cb) - target;
            // One evaluation is enough when the start is close: with a = -f/F'(x) the Newton step, the root of
            //   f + F' e + F'' e^2/2 + F''' e^3/6 + F'''' e^4/24 + ... = 0
            // is the reverted series e = a - A a^2 + (2A^2 - B) a^3 - (5A^3 - 5AB + C) a^4 + ..., where (q = cot x)
            //   A = F''/2F' = m q/2,  B = F'''/6F' = (m(m-1) q^2 - m)/6,  C = F''''/24F' = (m(m-1)(m-2) q^3 - m(3m-2) q)/24.
            // Truncating after a^3 leaves a residual of F' |5A^3 - 5AB + C| a^4 (1 + O(m q a)) in F; when 1.5x that bound
            // is below the stop rule's tolerance the verifying evaluation of the loop below is skipped, and sin/cos
            // of x + e come from the rotation by e (Taylor series to e^6, |e| <= 1e-2: truncation < 1e-17).
#ifndef HVB200_HW_ALWAYS_VERIFY                       // development switch: always run the verifying loop below
            if (fabs(f) > 1e-15) {
                // one division: t = a / sin x, so that A a = (m/2) v, B a^2 = (m(m-1) v^2 - m a^2)/6,
                // C a^3 = (m(m-1)(m-2) v^3 - m(3m-2) v a^2)/24 with v = t cos x — no cot x needed
                const double gm2 = ipow(sb, m - 2);              // m >= 2 on this path
                const double g = gm2 * sb * sb;
                const double t = -f / (g * sb);
                const double a = t * sb, v = t * cb, md = (double)m;
                const double u = 0.5 * md * v, v2 = v * v, a2 = a * a;
                const double mm1 = md * (md - 1.0);
                const double Ba2 = fma(mm1, v2, -md * a2) * (1.0 / 6.0);
                const double Ca3 = v * fma(mm1 * (md - 2.0), v2, -md * fma(3.0, md, -2.0) * a2) * (1.0 / 24.0);
                const double e = a * (fma(u, fma(2.0, u, -1.0), 1.0) - Ba2);
                const double bound = 1.5 * g * fabs(a) * (fabs(Ca3) + 5.0 * fabs(u) * fma(u, u, fabs(Ba2)));
                if (fabs(a) + md * fabs(v) <= 1e-2 && bound <= 5e-16 && x + e <= HALF_PI) {
                    const double e2 = e * e;
                    const double sd = e * fma(e2 * fma(e2, -1.0 / 20.0, 1.0), -1.0 / 6.0, 1.0);
                    const double omc = 0.5 * e2 * fma(e2 * fma(e2, -1.0 / 30.0, 1.0), -1.0 / 12.0, 1.0);
                    s_out = sb + fma(cb, sd, -sb * omc);
                    c_out = cb - fma(sb, sd, cb * omc);
#ifdef HVB200_HW_CHECK                                // development switch: count accepted solves that miss the stop rule
                    double s2, c2;
                    if (fabs(fm_eval(m, x + e, s2, c2) - target) > 1e-15 || fabs(s2 - s_out) > 4e-16 || fabs(c2 - c_out) > 4e-16) capped++;
#endif
                    return;
                }
            }
#endif
        }
    }
    int it = 0;
    while (fabs(f) > 1e-15) {
        if (++it > kNewtonCap || !(fabs(x) < 1e300)) { capped++; break; }
        const double g = ipow(sb, m);
        x -= f / g;
        x = (x > HALF_PI) ? HALF_PI : x;             // F is convex and increasing on [0, pi/2]: a step from the left may overshoot
        f = fm_eval(m, x, sb, cb) - target;
    }
    s_out = sb; c_out = cb;
}
__global__ void __launch_bounds__(128) k_hw_table_init(double *__restrict__ tab, int mmax, int ntab)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (mmax + 1) * (ntab + 1)) return;
    const int m = idx / (ntab + 1), g = idx - m * (ntab + 1);
    double theta;
    if (m == 0) theta = ((double)g / ntab) * c_Im[0];
    else if (g == 0) theta = 0.0;
    else {
        // reference-style Newton from pi/2, run to a tighter residual than the generator needs
        const double target = ((double)g / ntab) * c_Im[m];
        double x = 1.57079632679489661923, sb = 1.0, cb = 6.123233995736766e-17;
        double f = c_Im[m] - target;
        for (int it = 0; it < kNewtonCap && fabs(f) > 2e-16; it++) {
            x -= f / ipow(sb, m);
            f = fm_eval(m, x, sb, cb) - target;
        }
        theta = x;
    }
    tab[idx] = theta;
}

// One DZ2019-HW direction (c/hvapprox.c:676-680 for sample index i): x87-exact lattice point, d-1 solves, polar ->
// cartesian in the reference's product order.  emit(k, r_k, winv_k) is called for k = d-1 ... 0 with the reciprocal
// weight r_k = |w_k| <= 1e-20 ? 1e20 : 1/w_k (:640-646) and winv_k = w_k (the clamp's 1e-20 when clamped): the
// screens only need a winv with winv <= (1/r)(1 + 2^-53) for their conservative thresholds, and r = RN(1/w) gives
// w = (1/r)/(1 + e), |e| <= 2^-53 — one FP64 division per component instead of two.
template <typename Emit>
__device__ __forceinline__ void hw_direction(const HwDev &P, ull i, const double *__restrict__ tab, int ntab, unsigned &capped, Emit emit)
{
    const int d = P.d;
    const ull idx1 = i + 1;
    const bool last = !(idx1 < P.N);                // last sample: all zeros (c/hvapprox.c:300-303)
    X87Val fac;
    if (!last) fac = x87_div(idx1, P.N, P.Nrecip);
    double prod = 1.0;
    for (int q = 0; q < d - 1; q++) {
        const int m = d - 2 - q;
        const double u = last ? 0.0 : x87_mul_frac(fac, P.a[q]);
        double s, c;
        hw_solve(u * c_Im[m], u, m, tab, ntab, s, c, capped);
        // w_{d-1} = c_0; w_{d-1-q} = (s_0..s_{q-1}) c_q; w_0 = s_0..s_{d-2}  (c/hvapprox.c:632-638)
        const double w = (q == 0) ? c : prod * c;
        prod = (q == 0) ? s : s * prod;
        const bool tiny = fabs(w) <= 1e-20;
        emit(d - 1 - q, tiny ? 1e20 : 1.0 / w, tiny ? 1e-20 : w);
    }
    const bool tiny = fabs(prod) <= 1e-20;
    emit(0, tiny ? 1e20 : 1.0 / prod, tiny ? 1e-20 : prod);
}

// The rows of a warp's 32 samples are contiguous in `dirs` (32 x 2d doubles), but a thread writing its own row stores
// with a stride of 2d doubles: 32 L1 tag cycles per store instruction, 20 of them per sample at d = 10 — the L1 tag stage,
// not the arithmetic, bounded the kernel (18 % fewer instructions changed nothing).  Rows are staged in shared memory
// (row stride 2d + 1 doubles: conflict-free) and the warp copies its block out with coalesced stores.
// Measured (2e7 directions): staged 3.38 vs 3.6-3.8 ms at d = 10, but slower for d <= 6 (the copy loop outweighs
// six stores) and for d >= 16 (the staging buffer costs occupancy), so STAGED is used for 7 <= d <= 12 only.
// Row j of the launch is position p0 + j of a shard, i.e. lattice index first + ((p >> shift) * stride << shift) +
// (p & (2^shift - 1)): a contiguous range (shift = 62) or the block-cyclic shard of one rank.
template <bool STAGED>
__global__ void __launch_bounds__(128) k_gen_hw(const HwDev P, ull first, int shift, ull stride, ull p0, ull count, double *__restrict__ dirs,
                                                ull *capped_counter, const double *__restrict__ tab, int ntab)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int d = P.d, row_len = 2 * d, row_stride = 2 * d + 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *wrows = reinterpret_cast<double *>(smem_raw) + (size_t)warp * 32 * row_stride;
    const ull j = (ull)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned capped = 0;
    if (j < count) {
        double *out = STAGED ? wrows + lane * row_stride : dirs + j * (ull)row_len;
        const ull p = p0 + j;
        const ull i = first + (((p >> shift) * stride) << shift) + (p & ((1ull << shift) - 1ull));
        hw_direction(P, i, tab, ntab, capped, [&](int k, double r, double winv) { out[k] = r; out[d + k] = winv; });
    }
    __syncwarp();
    const ull wfirst = (ull)blockIdx.x * blockDim.x + (ull)warp * 32;      // first sample of this warp
    if (STAGED && wfirst < count) {
        const int rows = (int)min((ull)32, count - wfirst);
        double *dst = dirs + wfirst * (ull)row_len;
        int row = 0, col = lane;
        while (col >= row_len) { col -= row_len; row++; }
        for (int idx = lane; idx < rows * row_len; idx += 32) {
            dst[idx] = wrows[row * row_stride + col];
            col += 32;
            while (col >= row_len) { col -= row_len; row++; }
        }
    }
    if (capped) atomicAdd(capped_counter, (ull)capped);
}
#include "hvb200_bp.cuh"
#include "hvb200_bpprep.cuh"
#include "hvb200_contrib.cuh"

static bool gen_hw_staged(int d) { return d >= 7 && d <= 12; }
static int gen_hw_smem_bytes(int d) { return gen_hw_staged(d) ? 128 * (2 * d + 1) * (int)sizeof(double) : 0; }

__device__ __forceinline__ double safe_winv(double r)
{
    const double wi = 1.0 / r;
    return (wi < __longlong_as_double(0x7ff0000000000000LL)) ? wi : 0.0;   // inf/NaN -> 0 (always candidate)
}

// K3: DZ2019-MC post-processing (c/hvapprox.c:241-250, euclidean_norm :203-212)
__global__ void __launch_bounds__(128) k_gen_mc(const double *__restrict__ normals, ull count, int d,
                                                double *__restrict__ dirs)
{
    const ull j = (ull)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    const double *z = normals + j * (ull)d;
    double *out = dirs + j * (ull)(2 * d);
    double a0 = fabs(z[0]), a1 = fabs(z[1]);
    a0 = (a0 < 1e-20) ? 1e-20 : a0;
    a1 = (a1 < 1e-20) ? 1e-20 : a1;
    double nrm = a0 * a0 + a1 * a1;
    for (int k = 2; k < d; k++) {
        double ak = fabs(z[k]);
        ak = (ak < 1e-20) ? 1e-20 : ak;
        nrm += ak * ak;
    }
    nrm = sqrt(nrm);
    for (int k = 0; k < d; k++) {
        double ak = fabs(z[k]);
        ak = (ak < 1e-20) ? 1e-20 : ak;
        const double r = nrm / ak;
        out[k] = r;
        out[d + k] = safe_winv(r);
    }
}

// K4: Rphi-FWE+ mapping (fang_wang_efficient_mapping_plus, c/hvapprox.c:791-826) + r = 1/w (:855-856)
__global__ void __launch_bounds__(128) k_gen_rphi(const double *__restrict__ uu, ull count, int d,
                                                  double *__restrict__ dirs)
{
    const ull j = (ull)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    const double HALF_PI = 1.57079632679489661923;
    // uu is objective-major: u_k of sample j at uu[k * count + j] (coalesced; the host writes one
    // contiguous run per objective)
    auto u = [&](int k) { return uu[(ull)k * count + j]; };
    double *out = dirs + j * (ull)(2 * d);
    double prod = 1.0;
    const int pairs = d / 2 - 1;
    auto emit = [&](int k, double w) {
        const double r = 1.0 / w;
        out[k] = r;
        out[d + k] = safe_winv(r);
    };
    for (int i = 0; i < pairs; i++) {
        const int l = d - 2 * i - 2;
        const double t = pow(u(l - 1), 1.0 / l);
        const double dl = sqrt(1.0 - t * t) * prod;
        prod *= t;
        double sa, ca;
        sincos(HALF_PI * u(l), &sa, &ca);
        emit(2 * i, dl * sa);
        emit(2 * i + 1, dl * ca);
    }
    double ang;
    if (d % 2 == 0) {
        ang = HALF_PI * u(0);
    } else {
        emit(d - 3, prod * u(0));
        prod *= sqrt(1 - u(0) * u(0));
        ang = HALF_PI * u(1);
    }
    double sa, ca;
    sincos(ang, &sa, &ca);
    emit(d - 2, prod * sa);
    emit(d - 1, prod * ca);
}

// uploaded reciprocal directions -> (r, winv)
__global__ void __launch_bounds__(128) k_prep_r(const double *__restrict__ rin, ull count, int d,
                                                double *__restrict__ dirs)
{
    const ull j = (ull)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    for (int k = 0; k < d; k++) {
        const double r = rin[j * (ull)d + k];
        dirs[j * (ull)(2 * d) + k] = r;
        dirs[j * (ull)(2 * d) + d + k] = safe_winv(r);
    }
}
__global__ void __launch_bounds__(128) k_extract_r(const double *__restrict__ dirs, ull count, int d,
                                                   double *__restrict__ rout)
{
    const ull j = (ull)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    for (int k = 0; k < d; k++) rout[j * (ull)d + k] = dirs[j * (ull)(2 * d) + k];
}

// ------------------------------------------------------------------------------------------------
// FP64 pipe micro-benchmarks (roofline denominator measured on the same GPU)
// ------------------------------------------------------------------------------------------------
template <int OP>
__global__ void __launch_bounds__(256) k_fp64_bench(double *out, int iters, double seed)
{
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = seed + 1e-3 * (threadIdx.x + i);
    const double b = 1.0 + 1e-9 * threadIdx.x, c = 1e-12;
    int cnt = 0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (OP == 0) x[i] = fma(x[i], b, c);
                else if (OP == 1) x[i] = x[i] * b;
                else cnt += (x[i] > b + (double)(it + u)) ? 1 : 0;   // DSETP (+ integer add)
            }
        }
    }
    double s = cnt;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i];
    if (s == 123.456) out[0] = s;
}

// ------------------------------------------------------------------------------------------------
// transform_and_filter on the device (c/hvapprox.c:30-66): p' = maximise_k ? p - ref : ref - p,
// rows with any p'_k <= 0 are dropped, the survivors keep their order.  One flag per point, an exclusive sum of
// the flags (cub), a scatter of the surviving rows.  The subtraction is the reference's single IEEE operation,
// so the compacted set is bit-identical to the host path.
// ------------------------------------------------------------------------------------------------
struct TfArgs {
    double ref[HVB200_MAXD];
    unsigned char maxi[HVB200_MAXD];
    int d;
};
__global__ void __launch_bounds__(256) k_tf_flags(const double *__restrict__ data, ull n, const TfArgs a, unsigned *__restrict__ flags)
{
    const ull i = (ull)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double *src = data + i * (ull)a.d;
    unsigned keep = 1u;
    for (int k = 0; k < a.d; k++) {
        const double v = a.maxi[k] ? src[k] - a.ref[k] : a.ref[k] - src[k];
        if (!(v > 0)) keep = 0u;
    }
    flags[i] = keep;
}
__global__ void __launch_bounds__(256) k_tf_scatter(const double *__restrict__ data, ull n, const TfArgs a,
                                                    const unsigned *__restrict__ flags, const unsigned *__restrict__ pos,
                                                    double *__restrict__ out, ull *__restrict__ kept)
{
    const ull w = (ull)blockIdx.x * blockDim.x + threadIdx.x;       // element index = point * d + k
    if (w >= n * (ull)a.d) return;
    const ull i = w / (ull)a.d;
    const int k = (int)(w - i * (ull)a.d);
    if (flags[i]) out[(ull)pos[i] * (ull)a.d + k] = a.maxi[k] ? data[w] - a.ref[k] : a.ref[k] - data[w];
    if (w == 0) *kept = (ull)pos[n - 1] + flags[n - 1];
}
// ---- nondominated pre-filter: a transformed point q with p_k >= q_k for all k for some other
// point p can never be the arg-max of max_p min_k p_k r_k (r > 0), so dropping it leaves every sample value — and
// with the per-chunk sums every bit of the estimate — unchanged.  One thread per q, the set streamed through shared
// memory; of identical points the first survives.  O(n^2 d) compares with an early exit on the first objective
// that fails.
constexpr int kNdThreads = 128;                 // points q per CTA
constexpr int kNdTileBytes = 32 * 1024;         // slice of candidate dominators p per CTA
__global__ void __launch_bounds__(256) k_fill_u32(unsigned *__restrict__ a, ull n, unsigned v)
{
    const ull i = (ull)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = v;
}
// grid (q blocks, p slices): CTA (x, y) tests its 128 points q against slice y of the set and clears flags[q] (preset
// to 1) when a dominator turns up — a single thread walking the whole set is latency-bound (4 ms at n = 1e4).
// Points q = qi * qstride for qi < qcount, flags indexed by qi (qstride = 1, qcount = n: the whole set; a strided
// sample first tells whether the full pass is worth its time).
// strict = 1: only strict dominance clears a flag (duplicates stay; the contribution approximation's ignore_dominated).
__global__ void __launch_bounds__(kNdThreads) k_nd_flags(const double *__restrict__ pts, ull n, int d, int tile, unsigned *flags,
                                                         ull qstride, ull qcount, int strict = 0)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *sp = reinterpret_cast<double *>(smem_raw);                  // [tile][d]
    const ull qi = (ull)blockIdx.x * blockDim.x + threadIdx.x;
    const ull q = qi * qstride;
    const ull t0 = (ull)blockIdx.y * (ull)tile;
    const int tn = (int)min((ull)tile, n - t0);
    for (int i = threadIdx.x; i < tn * d; i += kNdThreads) sp[i] = pts[t0 * (ull)d + i];
    __syncthreads();
    if (qi >= qcount || flags[qi] == 0u) return;                        // already dominated (benign race: only ever cleared)
    double qv[HVB200_MAXD];
    for (int k = 0; k < d; k++) qv[k] = pts[q * (ull)d + k];
    for (int j = 0; j < tn; j++) {
        const double *p = sp + j * d;
        bool greater = false;
        int k = 0;
        for (; k < d; k++) {
            const double pk = p[k];
            if (pk < qv[k]) break;
            greater |= pk > qv[k];
        }
        if (k == d && (greater || (!strict && t0 + (ull)j < q))) { flags[qi] = 0u; return; }
    }
}
__global__ void __launch_bounds__(256) k_gather_flagged(const double *__restrict__ src, ull n, int d, const unsigned *__restrict__ flags,
                                                        const unsigned *__restrict__ pos, double *__restrict__ dst, ull *__restrict__ kept)
{
    const ull w = (ull)blockIdx.x * blockDim.x + threadIdx.x;       // element index = point * d + k
    if (w >= n * (ull)d) return;
    const ull i = w / (ull)d;
    if (flags[i]) dst[(ull)pos[i] * (ull)d + (w - i * (ull)d)] = src[w];
    if (w == 0) *kept = (ull)pos[n - 1] + flags[n - 1];
}
// per-objective minimum and maximum of a set of positive coordinates (bit patterns of positive doubles order
// like unsigned integers): lo initialised to +inf bits, hi to 0
__global__ void __launch_bounds__(256) k_colrange(const double *__restrict__ pts, ull n, int d, ull *__restrict__ lo, ull *__restrict__ hi)
{
    __shared__ ull slo[HVB200_MAXD], shi[HVB200_MAXD];
    for (int k = threadIdx.x; k < d; k += blockDim.x) { slo[k] = 0x7ff0000000000000ull; shi[k] = 0ull; }
    __syncthreads();
    const ull total = n * (ull)d;
    for (ull w = (ull)blockIdx.x * blockDim.x + threadIdx.x; w < total; w += (ull)gridDim.x * blockDim.x) {
        const int k = (int)(w % (ull)d);
        const ull bits = (ull)__double_as_longlong(pts[w]);
        atomicMin(&slo[k], bits);
        atomicMax(&shi[k], bits);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < d; k += blockDim.x) { atomicMin(&lo[k], slo[k]); atomicMax(&hi[k], shi[k]); }
}

// ------------------------------------------------------------------------------------------------
// Batched point sets: many small sets against ONE direction set — the loop of the
// reference's command line tool (c/main-hvapprox.c:149-174) and of optimisers that score a population of
// fronts per generation.  CTA (x, set) keeps its set in shared memory and evaluates the direction chunks
// x, x + gridDim.x, ... with the textbook arithmetic (get_expected_value, c/hvapprox.c:68-117); the chunk
// -> CTA map is static and every CTA owns one slot of `partial`, so the sums are deterministic.
// ------------------------------------------------------------------------------------------------
static constexpr int kSetsThreads = 256, kSetsR = 2;
template <int D>
__global__ void __launch_bounds__(kSetsThreads) k_maxmin_sets(const double *__restrict__ pts_all, const unsigned *__restrict__ set_off,
                                                              const double *__restrict__ dirs, ull count, int d_runtime,
                                                              double *__restrict__ partial)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *sp = reinterpret_cast<double *>(smem_raw);
    __shared__ double red[kSetsThreads / 32];
    const int d = D ? D : d_runtime;
    const int set = blockIdx.y;
    const unsigned p0 = set_off[set], n = set_off[set + 1] - p0;
    for (unsigned i = threadIdx.x; i < n * (unsigned)d; i += kSetsThreads) sp[i] = pts_all[(size_t)p0 * d + i];
    __syncthreads();
    constexpr ull CH = (ull)kSetsThreads * kSetsR;
    const ull n_chunks = (count + CH - 1) / CH;
    double acc = 0.0;
    for (ull chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
#pragma unroll
        for (int r = 0; r < kSetsR; r++) {
            const ull smp = chunk * CH + (ull)r * kSetsThreads + threadIdx.x;
            if (smp >= count || n == 0) continue;
            const double *dir = dirs + smp * (ull)(2 * d);
            double best = 0.0;
            if constexpr (D > 0) {
                double rr[D];
#pragma unroll
                for (int k = 0; k < D; k++) rr[k] = __ldg(dir + k);
                for (unsigned i = 0; i < n; i++) {
                    const double *pp = sp + (size_t)i * D;
                    double m = pp[0] * rr[0];
#pragma unroll
                    for (int k = 1; k < D; k++) {
                        const double t = pp[k] * rr[k];
                        m = (t < m) ? t : m;
                    }
                    best = (m > best) ? m : best;
                }
                acc += pow_chain<D>(best);
            } else {
                for (unsigned i = 0; i < n; i++) {
                    const double *pp = sp + (size_t)i * d;
                    double m = pp[0] * __ldg(dir);
                    for (int k = 1; k < d; k++) {
                        const double t = pp[k] * __ldg(dir + k);
                        m = (t < m) ? t : m;
                    }
                    best = (m > best) ? m : best;
                }
                // pow_uint (c/pow_int.h:59-74): tabulated chains up to 31, square-and-multiply above
                double x = best, v = (d & 1) ? best : 1.0;
                if (d <= 31) v = pow_chain_runtime(best, d);
                else for (unsigned e = (unsigned)d >> 1; e; e >>= 1) { x *= x; if (e & 1u) v *= x; }
                acc += v;
            }
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0)
        partial[(size_t)set * gridDim.x + blockIdx.x] += ((red[0] + red[1]) + (red[2] + red[3])) + ((red[4] + red[5]) + (red[6] + red[7]));
}
// one CTA per set: fixed-shape double-double tree over the set's CTA partials
__global__ void __launch_bounds__(256) k_reduce_sets(const double *__restrict__ partial, int per_set, double *__restrict__ out_hilo)
{
    __shared__ double sh[256], sl[256];
    const int tid = threadIdx.x, set = blockIdx.x;
    double hi = 0.0, lo = 0.0;
    for (int i = tid; i < per_set; i += 256) dd_add(hi, lo, partial[(size_t)set * per_set + i], 0.0);
    sh[tid] = hi; sl[tid] = lo;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (tid < off) {
            double h = sh[tid], l = sl[tid];
            dd_add(h, l, sh[tid + off], sl[tid + off]);
            sh[tid] = h; sl[tid] = l;
        }
        __syncthreads();
    }
    if (tid == 0) { out_hilo[2 * set] = sh[0]; out_hilo[2 * set + 1] = sl[0]; }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct DevPlan {
    int device = 0;
    int d = 0;
    size_t n = 0;
    int n_tiles = 0, tile_points = 0;
    int R = 1, G = 2, R_exact = 1;
    int staged = 0, stage_bytes = 0;
    size_t smem_bytes = 0;
    double *pts = nullptr;   size_t pts_cap = 0;       // doubles
    unsigned *keys = nullptr; size_t keys_cap = 0;     // DPX screen: packed 16-bit keys (words)
    double *qdev = nullptr;              // 3 x HVB200_MAXD doubles: qlo, qhi, qscale of the point set (k_quantize, device preparation)
    bool resources_ready = false;        // stream / events / small buffers exist (workspace reuse)
    bool ordered = false;                // points already sorted by probe arg-max frequency
    // block-pruned kernel (hvb200_bp.cuh)
    std::vector<double> host_pts;        // host copy of the transformed points when the plan was built from one (saves a
                                         // 0.6 ms pageable read-back in bp_prepare)
    bool bp_ready = false, bp_begun = false, bp_probe_queued = false;
    void *share = nullptr;               // BpShare of a single-process multi-GPU call (not owned)
    BpPrepScratch prep;                  // scratch of the device-side preparation (hvb200_bpprep.cuh)
    bool bp_on_device = false;           // the preparation of this plan was queued on the device (own_stream, event prep.done)
    bool fused = false;                  // this run: k_maxmin_bp generates its DZ2019-HW directions itself
    double *gen_scratch = nullptr; size_t gen_scratch_cap = 0;   // [grid][2d][CH] per-CTA direction blocks (L2 resident)
    double *bp_pts = nullptr; size_t bp_pts_cap = 0;
    float *bp_piv32 = nullptr; unsigned *bp_pividx = nullptr;
    uint2 *bp_hits = nullptr; size_t bp_hits_cap = 0;    // per-CTA hit lists of phase C
    uint4 *bp_quads = nullptr; size_t bp_quads_cap = 0;  // per-CTA lists of surviving (direction, block) quads
    double *bp_chunk_sums = nullptr; size_t bp_chunk_cap = 0, bp_chunk_off = 0;   // one partial sum per (chunk, warp) of the run
    unsigned *bp_chunk_counter = nullptr;
    double *reduce_pairs = nullptr; size_t reduce_pairs_cap = 0;        // (hi, lo) of the slices of a long run's partial sums
    unsigned *bp_pkeys = nullptr, *bp_bpair = nullptr; size_t bp_keys_cap = 0, bp_pair_cap = 0;
    int bp_nb = 0, bp_nb_pad = 0, bp_seg_blocks = 0, bp_npiv = 0;
    size_t bp_smem = 0;
    double bp_qscale[HVB200_MAXD] = {0}, bp_qlos[HVB200_MAXD] = {0};
    double *pts_tmp = nullptr; size_t pts_tmp_cap = 0;
    unsigned *wins = nullptr; size_t wins_cap = 0;      // 2n words: win counts, then the order
    int k_tile_pairs = 0, k_n_tiles = 0, k_stage_bytes = 0;
    size_t k_smem_bytes = 0;
    int kR = 1, kG = 2;
    double qlo[HVB200_MAXD] = {0}, qhi[HVB200_MAXD] = {0}, qscale[HVB200_MAXD] = {0};
    double *dirs = nullptr;      size_t dirs_cap = 0;      // doubles (the workspace may serve plans of different d)
    const double *dirs_view = nullptr;   // when set, the consumers read this (cached direction sets) instead of `dirs`
    double *cta_sums = nullptr;  int cta_cap = 0;
    double *values = nullptr;    size_t values_cap = 0;
    double *out_hilo = nullptr;
    ull *counters = nullptr;     // [0] newton capped, [1] slow path
    double *staging_host[2] = {nullptr, nullptr}; size_t staging_cap[2] = {0, 0};   // pinned, double-buffered
    cudaEvent_t staging_free[2] = {nullptr, nullptr};   // recorded after the H2D copy out of buffer i
    int staging_next = 0, staging_cur = 0;
    double *staging_dev = nullptr;  size_t staging_dev_cap = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    int sm_count = 0;
    int grid = 0;                // CTAs of the max-min kernel
    int kernel_used = 0;
    std::vector<cudaEvent_t> ev_pool;
    struct Span { int a, b, kind; };
    std::vector<Span> spans;
    size_t ev_next = 0;
    cudaEvent_t ev_h2d0 = nullptr, ev_h2d1 = nullptr;
    struct McGen *mc = nullptr;      // device generator of the DZ2019-MC stream (hvb200_mcgen.cuh)
};

static int span_begin(DevPlan *dp, int kind);
static void span_end(DevPlan *dp, int b);
#include "hvb200_mcgen.cuh"


static bool g_const_ready[16] = {false};
static double *g_hw_tab[16] = {nullptr};          // per device: F_m^{-1} tables for k_gen_hw (m = 0..31)
static std::mutex g_const_mutex;

static int upload_constants(int device)
{
    std::lock_guard<std::mutex> lock(g_const_mutex);     // plans may be created from several host threads
    if (device < 16 && g_const_ready[device]) return 0;
    static double lead[HVB200_MAXD], h[HVB200_MAXD][kMaxFmTerms], Im[HVB200_MAXD];
    static int nh[HVB200_MAXD];
    const long double PI_L = 3.141592653589793238462643383279502884L;
    for (int m = 0; m < HVB200_MAXD; m++) {
        long double ld = 1.0L, hh[kMaxFmTerms];
        int cnt = 0;
        for (int j = m; j >= 2; j -= 2) {
            hh[cnt++] = ld / (long double)j;
            ld *= (long double)(j - 1) / (long double)j;
        }
        lead[m] = (double)ld;
        nh[m] = cnt;
        for (int t = 0; t < kMaxFmTerms; t++) h[m][t] = 0.0;
        for (int t = 0; t < cnt; t++) h[m][t] = (double)hh[cnt - 1 - t];
        // I_m = lead * (pi/2 for even m, 1 for odd m)  [c/hvapprox.c:480-548]
        Im[m] = (double)((m & 1) ? ld : ld * (PI_L / 2));
    }
    CUDA_TRY(cudaMemcpyToSymbol(c_fm_lead, lead, sizeof lead));
    CUDA_TRY(cudaMemcpyToSymbol(c_fm_h, h, sizeof h));
    CUDA_TRY(cudaMemcpyToSymbol(c_fm_nh, nh, sizeof nh));
    CUDA_TRY(cudaMemcpyToSymbol(c_Im, Im, sizeof Im));
    CUDA_TRY(cudaMemcpyToSymbol(c_theta0_bits, hvb200_theta0_bits, sizeof(ull) * 32));
    if (device < 16 && g_hw_tab[device] == nullptr && !getenv("HVB200_NO_HW_TABLE")) {
        const int mmax = 31;
        double *tab = nullptr;
        if (g_hw_ntab == 0) {
            const char *e = getenv("HVB200_HWTAB");
            const int v = e ? atoi(e) : kHwTabDefault;
            g_hw_ntab = v < 64 ? 64 : (v > (1 << 20) ? (1 << 20) : v);
        }
        CUDA_TRY(cudaMalloc(&tab, (size_t)(mmax + 1) * (g_hw_ntab + 1) * sizeof(double)));
        const int total = (mmax + 1) * (g_hw_ntab + 1);
        k_hw_table_init<<<(total + 127) / 128, 128>>>(tab, mmax, g_hw_ntab);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaDeviceSynchronize());
        g_hw_tab[device] = tab;
    }
    if (device < 16) g_const_ready[device] = true;
    return 0;
}

extern "C" int hvb200cu_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// ---- launch geometry ---------------------------------------------------------------------------
// Per dimension: R directions per thread, G points per inner-loop group, whether the per-chunk
// direction block (r | winv, 16*R*D bytes per thread) is staged in shared memory, and the number
// of CTAs per SM the register budget is capped for.  Shared memory per CTA:
//   kStages * stage_bytes  +  staged ? 256*R*2*D*8 : 0  +  barriers.
struct KernelCfg { int R, G, minb; bool staged; };
static constexpr KernelCfg cfg_for(int D)
{
    //             R  G  minb staged
    if (D <= 5)  return {2, 4, 2, true};
    if (D <= 12) return {3, 2, 2, false};
    if (D <= 22) return {1, 2, 2, true};
    // 23, 24: one CTA per SM as well (d = 23, n = 2000: EXACT 6.09 -> 3.38 ms per 4e5 directions, SCREEN 10.9 -> 3.26); below,
    // two CTAs with their 100-330 bytes of spills are faster (d = 17, 18, 20: 2-30 % slower with one)
    if (D <= 24) return {1, 2, 1, true};
    // 25..32: one CTA per SM (255 registers).  With two, r[D] and the thresholds spilled 650-900 bytes per thread and the
    // textbook loop ran at half the speed (d = 30, n = 2000: EXACT 9.16 -> 4.05 ms per 4e5 directions, DPX 4.07 -> 3.15)
    if (D <= 32) return {1, 2, 1, false};
    return {1, 2, 1, false};      // 33..64 (extension, EXACT only): r[D] alone is up to 128 registers -> one CTA per SM
}
static constexpr size_t kSmemPerSm = 227 * 1024;
static constexpr size_t kBarrierBytes = (2 * kStages + 2) * sizeof(ull) + 8 * sizeof(double);

typedef void (*maxmin_fn)(const MaxminArgs);
// The textbook loop keeps r (not thresholds) in registers and is bound by FP64 + ALU issue; two
// directions per thread measured best (d=10: {2,2} 541 Gsp/s vs {3,2} 484).  Same tile geometry as the
// screen configuration of the same D (only R and G differ).
static constexpr KernelCfg cfg_exact_for(int D)
{
    if (D <= 5)  return {2, 4, 2, true};
    if (D <= 12) return {2, 2, 2, false};
    return cfg_for(D);
}
template <int D>
static maxmin_fn pick_mode(int mode)
{
    constexpr KernelCfg c = cfg_for(D);
    constexpr KernelCfg e = cfg_exact_for(D);
    static_assert(c.staged == e.staged && c.minb == e.minb, "EXACT and SCREEN share the shared-memory geometry");
    if constexpr (D > 32) {
        return (maxmin_fn)k_maxmin<D, e.R, e.G, MODE_EXACT, e.minb>;       // the screens stop at 32 objectives
    } else {
        return mode == MODE_EXACT ? (maxmin_fn)k_maxmin<D, e.R, e.G, MODE_EXACT, e.minb>
                                  : (maxmin_fn)k_maxmin<D, c.R, c.G, MODE_SCREEN, c.minb>;
    }
}
template <int D>
static maxmin_fn pick_d(int d, int mode)
{
    if constexpr (D < 2) {
        return nullptr;
    } else {
        if (d == D) return pick_mode<D>(mode);
        return pick_d<D - 1>(d, mode);
    }
}
static maxmin_fn pick_kernel(int d, int mode) { return pick_d<HVB200_MAXD>(d, mode); }
// DPX screen geometry: R directions per thread, G point pairs per group ((G*D) % 4 == 0).
struct DpxCfg { int R, G, minb; };
static constexpr DpxCfg dpx_cfg_for(int D)
{
    //                          R  G  CTAs/SM      (measured on B200: d=10 {4,4,2} 2.10 Tsp/s; d=20 {3,4,2} 0.96, {4,2,2} 0.88)
    if (D <= 8)  return {4, 4, 3};
    if (D <= 12) return {4, 4, 2};
    if (D <= 20) return {3, 4, 2};
    if (D <= 24) return {2, (D % 2) ? 4 : 2, 2};      // (d = 23: 1.86 ms with two CTAs per SM, 1.98 with one)
    return {2, (D % 2) ? 4 : 2, 1};
}
typedef void (*dpx_fn)(const DpxArgs);
template <int D>
static dpx_fn pick_dpx(int d)
{
    if constexpr (D < 2) {
        return nullptr;
    } else {
        if (d == D) {
            constexpr DpxCfg c = dpx_cfg_for(D);
            return (dpx_fn)k_maxmin_dpx<D, c.R, c.G, c.minb>;
        }
        return pick_dpx<D - 1>(d);
    }
}
static KernelCfg cfg_runtime(int d) { return cfg_for(d); }

// Workspace cache: a DevPlan keeps its device buffers, stream and events when its plan is destroyed
// and is handed to the next plan created on the same device, so that repeated hv_approx calls (the
// common use: one call per generation of an optimiser) pay cudaMalloc/cudaFree only once.
// HVB200_NO_CACHE=1 disables it; at most kPoolMax workspaces are kept per device.
static std::mutex g_pool_mutex;
static std::vector<DevPlan *> g_pool[16];
static constexpr size_t kPoolMax = 4;
static bool pool_enabled()
{
    const char *e = getenv("HVB200_NO_CACHE");
    return !(e && e[0] == '1');
}
static DevPlan *pool_take(int device)
{
    if (device < 0 || device >= 16 || !pool_enabled()) return nullptr;
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    if (g_pool[device].empty()) return nullptr;
    DevPlan *dp = g_pool[device].back();
    g_pool[device].pop_back();
    return dp;
}
static void devplan_release(DevPlan *dp);
static bool pool_give(DevPlan *dp)
{
    if (dp->device < 0 || dp->device >= 16 || !pool_enabled()) return false;
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    if (g_pool[dp->device].size() >= kPoolMax) return false;
    g_pool[dp->device].push_back(dp);
    return true;
}

static int plan_upload_impl(struct hvb200_plan *plan, const double *pts_host, const double *pts_dev);
// Scratch of the plan-creation kernels (transform_and_filter, nondominated filter): one cached allocation per device
// (cudaMalloc/cudaFree cost more than the kernels: 1.3 of 2.1 ms at n = 5e4, d = 20); released again when it grew
// beyond 256 MB.  Users hold the device's tf_mutex (one per device: plans for different GPUs are built concurrently by
// hvb200_hv_approx_multi).
static std::mutex tf_mutex[16];
static void *tf_scratch[16] = {nullptr};
static size_t tf_cap[16] = {0};
struct TfScratch { double *raw, *out; unsigned *flags, *pos; ull *kept_dev; void *tmp; size_t tmp_bytes; };
static int tf_scratch_acquire(int device, size_t npoints, int d, TfScratch &sc)
{
    const int dev = device < 16 ? device : 15;
    const size_t elems = npoints * (size_t)d;
    size_t tmp_bytes = 0;
    if (cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, (unsigned *)nullptr, (unsigned *)nullptr, (int)npoints, (cudaStream_t)0) != cudaSuccess) {
        hvb200_set_error("cub scan sizing failed"); cudaGetLastError(); return -1;
    }
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t need = 2 * up(elems * sizeof(double)) + 2 * up(npoints * sizeof(unsigned)) + up(sizeof(ull)) + up(tmp_bytes ? tmp_bytes : 16);
    if (tf_cap[dev] < need) {
        cudaFree(tf_scratch[dev]);
        tf_scratch[dev] = nullptr; tf_cap[dev] = 0;
        if (cudaMalloc(&tf_scratch[dev], need) != cudaSuccess) {
            cudaGetLastError();
            hvb200_set_error("out of device memory (plan creation scratch, %zu points)", npoints);
            return -1;
        }
        tf_cap[dev] = need;
    }
    unsigned char *base = (unsigned char *)tf_scratch[dev];
    sc.raw = (double *)base;                          base += up(elems * sizeof(double));
    sc.out = (double *)base;                          base += up(elems * sizeof(double));
    sc.flags = (unsigned *)base;                      base += up(npoints * sizeof(unsigned));
    sc.pos = (unsigned *)base;                        base += up(npoints * sizeof(unsigned));
    sc.kept_dev = (ull *)base;                        base += up(sizeof(ull));
    sc.tmp = (void *)base;
    sc.tmp_bytes = tmp_bytes;
    return 0;
}
static void tf_scratch_trim(int device)
{
    const int dev = device < 16 ? device : 15;
    if (tf_cap[dev] > ((size_t)256 << 20)) { cudaFree(tf_scratch[dev]); tf_scratch[dev] = nullptr; tf_cap[dev] = 0; }
}

// HVB200_NDFILTER=0|1 forces the nondominated pre-filter off / on.  By default it is considered for 512 <= n' <= 2^18
// (below, the EXACT kernel reads every point anyway; above, the O(n'^2) pass takes longer than most runs) and runs
// only if a strided sample of 256 points, tested against the whole set first (0.05 ms), has at least 1/16 of its
// points dominated: a front gains nothing from the full pass (1.2 ms at n' = 1e4, 73 ms at 1e5, d = 10).
static int nd_filter_mode(size_t n)        // 0 off, 1 on, 2 decide by sample
{
    const char *env = getenv("HVB200_NDFILTER");
    if (env && env[0] == '0') return 0;
    if (env && env[0] == '1') return (n >= 2 && n <= ((size_t)1 << 22)) ? 1 : 0;      /* grid.y = n / tile <= 65535 */
    return (n >= 512 && n <= ((size_t)1 << 18)) ? 2 : 0;
}
constexpr int kNdSample = 256;
static int nd_tile_rows(size_t n, int d)
{
    int tile = kNdTileBytes / (8 * d);
    if (tile > 512) tile = 512;
    if ((size_t)tile > n) tile = (int)n;
    return tile;
}
// true when the full filter should run on src (device, n x d)
static bool nd_filter_decide(const TfScratch &sc, const double *src, size_t n, int d)
{
    const int mode = nd_filter_mode(n);
    if (mode != 2) return mode == 1;
    // (a small slice per CTA: the sample pass has 2 q-blocks only and needs the slices to fill the GPU)
    const int tile = std::max(32, std::min(nd_tile_rows(n, d), (int)(n / 296)));
    unsigned host[kNdSample];
    k_fill_u32<<<1, 256>>>(sc.pos, (ull)kNdSample, 1u);
    const dim3 grid((kNdSample + kNdThreads - 1) / kNdThreads, (unsigned)((n + tile - 1) / tile));
    k_nd_flags<<<grid, kNdThreads, (size_t)tile * d * sizeof(double)>>>(src, (ull)n, d, tile, sc.pos, (ull)(n / kNdSample), (ull)kNdSample);
    if (cudaMemcpy(host, sc.pos, sizeof host, cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); return false; }
    int dominated = 0;
    for (int i = 0; i < kNdSample; i++) dominated += host[i] ? 0 : 1;
    return dominated * 16 >= kNdSample;
}
// src -> dst (device, n x d rows): the nondominated rows in their original order
static int nd_filter_device(const TfScratch &sc, const double *src, double *dst, size_t n, int d, size_t *kept_out)
{
    const int tile = nd_tile_rows(n, d);
    ull kept = 0;
    k_fill_u32<<<(unsigned)((n + 255) / 256), 256>>>(sc.flags, (ull)n, 1u);
    const dim3 grid((unsigned)((n + kNdThreads - 1) / kNdThreads), (unsigned)((n + tile - 1) / tile));
    k_nd_flags<<<grid, kNdThreads, (size_t)tile * d * sizeof(double)>>>(src, (ull)n, d, tile, sc.flags, 1ull, (ull)n);
    size_t tmp_bytes = sc.tmp_bytes;
    bool ok = cub::DeviceScan::ExclusiveSum(sc.tmp, tmp_bytes, sc.flags, sc.pos, (int)n, (cudaStream_t)0) == cudaSuccess;
    if (ok) {
        k_gather_flagged<<<(unsigned)((n * (size_t)d + 255) / 256), 256>>>(src, (ull)n, d, sc.flags, sc.pos, dst, sc.kept_dev);
        ok = cudaMemcpy(&kept, sc.kept_dev, sizeof kept, cudaMemcpyDeviceToHost) == cudaSuccess;
    }
    if (!ok) {
        hvb200_set_error("nondominated filter failed: %s", cudaGetErrorString(cudaGetLastError()));
        return -1;
    }
    *kept_out = (size_t)kept;
    return 0;
}

// flags_host[i] = 1 unless some other point strictly dominates point i (transformed space: larger is better);
// duplicates keep their flags.  Used by the contribution approximation (ignore_dominated).
extern "C" int hvb200cu_nd_strict_flags(int device, const double *pts_host, size_t n, int d, unsigned *flags_host)
{
    CUDA_TRY(cudaSetDevice(device));
    std::lock_guard<std::mutex> lock(tf_mutex[device < 16 ? device : 15]);
    TfScratch sc;
    if (tf_scratch_acquire(device, n, d, sc)) return -1;
    CUDA_TRY(cudaMemcpy(sc.raw, pts_host, n * (size_t)d * sizeof(double), cudaMemcpyHostToDevice));
    const int tile = nd_tile_rows(n, d);
    k_fill_u32<<<(unsigned)((n + 255) / 256), 256>>>(sc.flags, (ull)n, 1u);
    const dim3 grid((unsigned)((n + kNdThreads - 1) / kNdThreads), (unsigned)((n + tile - 1) / tile));
    k_nd_flags<<<grid, kNdThreads, (size_t)tile * d * sizeof(double)>>>(sc.raw, (ull)n, d, tile, sc.flags, 1ull, (ull)n, 1);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpy(flags_host, sc.flags, n * sizeof(unsigned), cudaMemcpyDeviceToHost));
    tf_scratch_trim(device);
    return 0;
}

extern "C" int hvb200cu_plan_upload(struct hvb200_plan *plan, const double *pts_host)
{
    plan->n_dominated = 0;
    if (!plan->no_ndfilter && nd_filter_mode(plan->n)) {
        CUDA_TRY(cudaSetDevice(plan->device));
        std::lock_guard<std::mutex> lock(tf_mutex[plan->device < 16 ? plan->device : 15]);
        TfScratch sc;
        if (tf_scratch_acquire(plan->device, plan->n, plan->d, sc)) return -1;
        size_t kept = 0;
        if (cudaMemcpy(sc.raw, pts_host, plan->n * (size_t)plan->d * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) {
            hvb200_set_error("upload of the point set failed: %s", cudaGetErrorString(cudaGetLastError()));
            return -1;
        }
        kept = plan->n;
        if (nd_filter_decide(sc, sc.raw, plan->n, plan->d) && nd_filter_device(sc, sc.raw, sc.out, plan->n, plan->d, &kept)) return -1;
        int rc;
        if (kept < plan->n) {
            plan->n_dominated = plan->n - kept;
            plan->n = kept;
            rc = plan_upload_impl(plan, nullptr, sc.out);    // ends with a stream synchronisation: the scratch is free afterwards
        } else {
            // nothing filtered: the set is on the device already (scratch), a second upload from the host would cost more
            // than the device-side copy and range reduction
            rc = plan_upload_impl(plan, nullptr, sc.raw);
        }
        tf_scratch_trim(plan->device);
        return rc;
    }
    return plan_upload_impl(plan, pts_host, nullptr);
}

// transform_and_filter (c/hvapprox.c:30-66) on the device: flags -> exclusive scan -> scatter, then the
// nondominated filter; the surviving rows keep the reference's order.
extern "C" int hvb200cu_plan_upload_raw(struct hvb200_plan *plan, const double *data_host, size_t npoints,
                                        const double *ref, const unsigned char *maximise, size_t *kept_out)
{
    *kept_out = 0;
    plan->n_dominated = 0;
    CUDA_TRY(cudaSetDevice(plan->device));
    const int d = plan->d;
    if (npoints == 0) return 0;
    if (npoints >= 0x7fffffffull) { hvb200_set_error("on-device transform_and_filter supports < 2^31 points"); return -1; }
    std::lock_guard<std::mutex> lock(tf_mutex[plan->device < 16 ? plan->device : 15]);
    TfScratch sc;
    if (tf_scratch_acquire(plan->device, npoints, d, sc)) return -1;
    const size_t elems = npoints * (size_t)d;
    int rc = -1;
    ull kept = 0;
    TfArgs a;
    a.d = d;
    for (int k = 0; k < HVB200_MAXD; k++) { a.ref[k] = k < d ? ref[k] : 0.0; a.maxi[k] = k < d ? (maximise[k] ? 1 : 0) : 0; }
    do {
        if (cudaMemcpy(sc.raw, data_host, elems * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) break;
        k_tf_flags<<<(unsigned)((npoints + 255) / 256), 256>>>(sc.raw, (ull)npoints, a, sc.flags);
        size_t tmp_bytes = sc.tmp_bytes;
        if (cub::DeviceScan::ExclusiveSum(sc.tmp, tmp_bytes, sc.flags, sc.pos, (int)npoints, (cudaStream_t)0) != cudaSuccess) break;
        k_tf_scatter<<<(unsigned)((elems + 255) / 256), 256>>>(sc.raw, (ull)npoints, a, sc.flags, sc.pos, sc.out, sc.kept_dev);
        if (cudaMemcpy(&kept, sc.kept_dev, sizeof kept, cudaMemcpyDeviceToHost) != cudaSuccess) break;
        rc = 0;
    } while (0);
    if (rc) {
        const cudaError_t e = cudaGetLastError();
        hvb200_set_error("on-device transform_and_filter failed: %s", cudaGetErrorString(e));
    } else if (kept) {
        const double *final_pts = sc.out;
        if (nd_filter_decide(sc, sc.out, (size_t)kept, d)) {
            size_t kept2 = 0;
            rc = nd_filter_device(sc, sc.out, sc.raw, (size_t)kept, d, &kept2);
            if (rc == 0 && kept2 < kept) { plan->n_dominated = (size_t)kept - kept2; kept = kept2; final_pts = sc.raw; }
        }
        if (rc == 0) {
            plan->n = (size_t)kept;
            rc = plan_upload_impl(plan, nullptr, final_pts);      // ends with a stream synchronisation: the scratch is free afterwards
            *kept_out = (size_t)kept;
        }
    }
    tf_scratch_trim(plan->device);
    return rc;
}

static int plan_upload_impl(struct hvb200_plan *plan, const double *pts_host, const double *pts_dev)
{
    HVB200_NVTX("hvb200:plan_create");
    CUDA_TRY(cudaSetDevice(plan->device));
    if (upload_constants(plan->device)) return -1;
    DevPlan *dp = pool_take(plan->device);
    if (!dp) dp = new DevPlan();
    plan->dev = dp;
    dp->device = plan->device;
    dp->d = plan->d;
    dp->n = plan->n;
    dp->ordered = false;
    dp->bp_ready = dp->bp_begun = dp->bp_probe_queued = false;
    dp->share = plan->share;
    const int d = plan->d;
    const KernelCfg kc = cfg_runtime(d);
    dp->R = kc.R;
    dp->G = kc.G;
    dp->R_exact = d > 32 ? 1 : cfg_exact_for(d).R;
    dp->staged = kc.staged ? 1 : 0;
    // shared memory budget per CTA for `minb` CTAs per SM -> bytes per pipeline stage
    const size_t dir_bytes = kc.staged ? (size_t)kConsumerThreads * kc.R * 2 * d * 8 : 0;
    size_t per_cta = kSmemPerSm / kc.minb - 1024;                     // 1 KB/CTA is reserved by the driver
    size_t stage = (per_cta - dir_bytes - kBarrierBytes) / kStages;
    if (stage > (size_t)kMaxStageBytes) stage = kMaxStageBytes;
    stage &= ~(size_t)127;
    int tp = (int)(stage / (8 * d));
    tp &= ~3;                                      // multiple of 4 (covers G = 1, 2 and 4)
    if (tp < 4) { hvb200_set_error("internal: no room for a point tile (d=%d)", d); return -1; }
    if ((size_t)tp > ((plan->n + 3) & ~(size_t)3)) tp = (int)((plan->n + 3) & ~(size_t)3);
    dp->tile_points = tp;
    dp->stage_bytes = (int)((((size_t)tp * d * 8) + 127) & ~(size_t)127);
    dp->smem_bytes = (size_t)kStages * dp->stage_bytes + dir_bytes + kBarrierBytes;
    dp->n_tiles = (int)((plan->n + tp - 1) / tp);
    const size_t padded = (size_t)dp->n_tiles * tp * d;
    if (!dp->resources_ready) {
        CUDA_TRY(cudaMalloc(&dp->out_hilo, 2 * sizeof(double)));
        CUDA_TRY(cudaMalloc(&dp->counters, 8 * sizeof(ull)));
        CUDA_TRY(cudaMalloc(&dp->qdev, 3 * HVB200_MAXD * sizeof(double)));
        CUDA_TRY(cudaStreamCreateWithFlags(&dp->own_stream, cudaStreamNonBlocking));
        cudaDeviceProp prop;
        CUDA_TRY(cudaGetDeviceProperties(&prop, plan->device));
        dp->sm_count = prop.multiProcessorCount;
        CUDA_TRY(cudaEventCreate(&dp->ev_h2d0));
        CUDA_TRY(cudaEventCreate(&dp->ev_h2d1));
        dp->resources_ready = true;
    }
    cudaStream_t us = dp->own_stream;
    if (dp->pts_cap < padded) {
        cudaFree(dp->pts);
        dp->pts = nullptr; dp->pts_cap = 0;
        CUDA_TRY(cudaMalloc(&dp->pts, padded * sizeof(double)));
        dp->pts_cap = padded;
    }
    // zero only the padding rows behind the real points
    if (pts_dev) { CUDA_TRY(cudaMemcpyAsync(dp->pts, pts_dev, plan->n * (size_t)d * sizeof(double), cudaMemcpyDeviceToDevice, us)); dp->host_pts.clear(); }
    else {
        CUDA_TRY(cudaMemcpyAsync(dp->pts, pts_host, plan->n * (size_t)d * sizeof(double), cudaMemcpyHostToDevice, us));
        dp->host_pts.assign(pts_host, pts_host + plan->n * (size_t)d);
    }
    if (padded > plan->n * (size_t)d)
        CUDA_TRY(cudaMemsetAsync(dp->pts + plan->n * (size_t)d, 0, (padded - plan->n * (size_t)d) * sizeof(double), us));
    {
        // ---- per-objective range of the point set (key maps of the DPX screen and of the pruned kernel)
        for (int k = 0; k < HVB200_MAXD; k++) { dp->qlo[k] = INFINITY; dp->qhi[k] = -INFINITY; dp->qscale[k] = 0.0; }
        if (pts_dev) {
            // by a device reduction (the set never visits the host)
            ull init[2 * HVB200_MAXD], res[2 * HVB200_MAXD];
            for (int k = 0; k < HVB200_MAXD; k++) { init[k] = 0x7ff0000000000000ull; init[HVB200_MAXD + k] = 0ull; }
            ull *rng = reinterpret_cast<ull *>(dp->qdev);               // two thirds of it, overwritten below
            CUDA_TRY(cudaMemcpyAsync(rng, init, sizeof init, cudaMemcpyHostToDevice, us));
            k_colrange<<<dp->sm_count * 4, 256, 0, us>>>(dp->pts, (ull)plan->n, d, rng, rng + HVB200_MAXD);
            CUDA_TRY(cudaMemcpyAsync(res, rng, sizeof res, cudaMemcpyDeviceToHost, us));
            CUDA_TRY(cudaStreamSynchronize(us));
            for (int k = 0; k < d; k++) { memcpy(&dp->qlo[k], &res[k], 8); memcpy(&dp->qhi[k], &res[HVB200_MAXD + k], 8); }
        } else {
            for (size_t i = 0; i < plan->n; i++)
                for (int k = 0; k < d; k++) {
                    const double v = pts_host[i * (size_t)d + k];
                    if (v < dp->qlo[k]) dp->qlo[k] = v;
                    if (v > dp->qhi[k]) dp->qhi[k] = v;
                }
        }
        for (int k = 0; k < d; k++) {
            const double range = dp->qhi[k] - dp->qlo[k];
            dp->qscale[k] = (range > 0 && std::isfinite(range)) ? (double)kKeyScale / range : 0.0;
            if (!std::isfinite(dp->qscale[k])) dp->qscale[k] = 0.0;
        }
        double qhost[3 * HVB200_MAXD];
        memcpy(qhost, dp->qlo, HVB200_MAXD * sizeof(double));
        memcpy(qhost + HVB200_MAXD, dp->qhi, HVB200_MAXD * sizeof(double));
        memcpy(qhost + 2 * HVB200_MAXD, dp->qscale, HVB200_MAXD * sizeof(double));
        CUDA_TRY(cudaMemcpyAsync(dp->qdev, qhost, sizeof qhost, cudaMemcpyHostToDevice, us));
    }
    if (d <= 32) {
        // ---- DPX keys: monotone 16-bit quantisation on the device
        const DpxCfg dc = dpx_cfg_for(d);
        dp->kR = dc.R;
        dp->kG = dc.G;
        const size_t total_pairs_raw = (plan->n + 1) / 2;
        int tpairs = (16 * 1024 / (4 * d)) & ~3;
        if ((size_t)tpairs > ((total_pairs_raw + 3) & ~(size_t)3)) tpairs = (int)((total_pairs_raw + 3) & ~(size_t)3);
        dp->k_tile_pairs = tpairs;
        dp->k_n_tiles = (int)((total_pairs_raw + tpairs - 1) / tpairs);
        dp->k_stage_bytes = (int)((((size_t)tpairs * d * 4) + 127) & ~(size_t)127);
        dp->k_smem_bytes = (size_t)kStages * dp->k_stage_bytes + kBarrierBytes + (size_t)dpx_pivots() * d * sizeof(double);
        const size_t total_pairs = (size_t)dp->k_n_tiles * tpairs;
        const ull words = (ull)total_pairs * d;
        if (dp->keys_cap < words) {
            cudaFree(dp->keys);
            dp->keys = nullptr; dp->keys_cap = 0;
            CUDA_TRY(cudaMalloc(&dp->keys, words * sizeof(unsigned)));
            dp->keys_cap = words;
        }
        double *qdev = dp->qdev;
        k_quantize<<<(unsigned)((words + 255) / 256), 256, 0, us>>>(dp->pts, (ull)plan->n, d, tpairs, (ull)total_pairs, qdev, qdev + HVB200_MAXD, qdev + 2 * HVB200_MAXD, dp->keys);
        CUDA_TRY(cudaGetLastError());
    }
    // pts_host and qhost are caller/stack memory: the copies must have left them before returning; later
    // runs may use another stream, so everything issued here is complete when the plan is handed out
    CUDA_TRY(cudaStreamSynchronize(us));
    return 0;
}

extern "C" void hvb200cu_plan_free(struct hvb200_plan *plan)
{
    DevPlan *dp = (DevPlan *)plan->dev;
    if (!dp) return;
    plan->dev = nullptr;
    cudaSetDevice(dp->device);
    // all work of the plan is complete here (every run ends with a stream synchronisation)
    if (dp->resources_ready && pool_give(dp)) return;
    devplan_release(dp);
}
static void devplan_release(DevPlan *dp)
{
    cudaFree(dp->pts); cudaFree(dp->keys); cudaFree(dp->dirs); cudaFree(dp->cta_sums); cudaFree(dp->values);
    cudaFree(dp->out_hilo); cudaFree(dp->counters); cudaFree(dp->staging_dev); cudaFree(dp->qdev);
    cudaFree(dp->pts_tmp); cudaFree(dp->wins); cudaFree(dp->gen_scratch); cudaFree(dp->reduce_pairs);
    bp_prep_scratch_free(&dp->prep);
    cudaFree(dp->bp_hits); cudaFree(dp->bp_quads); cudaFree(dp->bp_chunk_sums); cudaFree(dp->bp_chunk_counter); cudaFree(dp->bp_pts); cudaFree(dp->bp_piv32); cudaFree(dp->bp_pividx); cudaFree(dp->bp_pkeys); cudaFree(dp->bp_bpair);
    for (int i = 0; i < 2; i++) {
        if (dp->staging_host[i]) cudaFreeHost(dp->staging_host[i]);
        if (dp->staging_free[i]) cudaEventDestroy(dp->staging_free[i]);
    }
    for (auto e : dp->ev_pool) cudaEventDestroy(e);
    if (dp->ev_h2d0) cudaEventDestroy(dp->ev_h2d0);
    if (dp->ev_h2d1) cudaEventDestroy(dp->ev_h2d1);
    if (dp->mc) { mcgen_free(dp->mc); delete dp->mc; }
    if (dp->own_stream) cudaStreamDestroy(dp->own_stream);
    delete dp;
}

static int span_begin(DevPlan *dp, int kind)
{
    while (dp->ev_pool.size() < dp->ev_next + 2) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return -1;
        dp->ev_pool.push_back(e);
    }
    const int a = (int)dp->ev_next, b = a + 1;
    dp->ev_next += 2;
    dp->spans.push_back({a, b, kind});
    cudaEventRecord(dp->ev_pool[a], dp->stream);
    return b;
}
static void span_end(DevPlan *dp, int b) { if (b >= 0) cudaEventRecord(dp->ev_pool[b], dp->stream); }


// Build the block-pruned kernel's inputs (hvb200_bp.cuh): pivots = the most frequent probe winners,
// blocks = 32 consecutive points of a kd-tree order, two key layouts.  Host work is O(n d log n), once per plan.
typedef void (*bp_fn)(const BpArgs);
static constexpr int bp_R_of(int D) { return HVB200_BP_R2 && D <= 16 ? 2 : 1; }
template <int D>
static bp_fn pick_bp(int d)
{
    if constexpr (D < 2) {
        return nullptr;
    } else {
        if (d == D) return (bp_fn)k_maxmin_bp<D, bp_R_of(D), false>;
        return pick_bp<D - 1>(d);
    }
}
// the variant that generates its DZ2019-HW directions itself (d <= 32: the reference's method stops there)
template <int D>
static bp_fn pick_bp_gen(int d)
{
    if constexpr (D < 2) {
        return nullptr;
    } else {
        if (d == D) return (bp_fn)k_maxmin_bp<D, bp_R_of(D), true>;
        return pick_bp_gen<D - 1>(d);
    }
}
// Host threads of one kd build: 2^g_kd_depth.  Several plans may be prepared at the same time — one process per GPU
// under torchrun (LOCAL_WORLD_SIZE), or the shard threads of a single-process multi-GPU call — so the build takes
// its share of the host's cores instead of 8 threads each (64 threads on a 32-core box at 8 GPUs).
static int bp_kd_depth()
{
    unsigned cores = std::thread::hardware_concurrency();
    if (cores == 0) cores = 8;
    int sharers = 1;
    if (const char *e = getenv("LOCAL_WORLD_SIZE")) sharers = std::max(1, atoi(e));
    const unsigned mine = std::max(1u, cores / (unsigned)sharers);
    return mine >= 8 ? 3 : (mine >= 4 ? 2 : (mine >= 2 ? 1 : 0));
}
static void bp_kd_order(const std::vector<double> &norm, int d, std::vector<unsigned> &idx, size_t lo, size_t hi, int max_depth, int depth = 0)
{
    const size_t cnt = hi - lo;
    if (cnt <= (size_t)kBpBlock) return;
    // widest objective of the node: one pass over its rows
    double mn[HVB200_MAXD], mx[HVB200_MAXD];
    for (int k = 0; k < d; k++) { mn[k] = INFINITY; mx[k] = -INFINITY; }
    for (size_t i = lo; i < hi; i++) {
        const double *row = &norm[(size_t)idx[i] * d];
        for (int k = 0; k < d; k++) { mn[k] = row[k] < mn[k] ? row[k] : mn[k]; mx[k] = row[k] > mx[k] ? row[k] : mx[k]; }
    }
    int best_k = 0;
    for (int k = 1; k < d; k++) if (mx[k] - mn[k] > mx[best_k] - mn[best_k]) best_k = k;
    size_t half = (cnt / 2 + kBpBlock - 1) / kBpBlock * kBpBlock;      // left part = whole blocks
    if (half >= cnt) half = cnt / 2;
    // median split on (key, index) pairs: contiguous keys for nth_element, index as the tie-break
    std::vector<std::pair<double, unsigned>> kv(cnt);
    for (size_t i = 0; i < cnt; i++) kv[i] = {norm[(size_t)idx[lo + i] * d + best_k], idx[lo + i]};
    std::nth_element(kv.begin(), kv.begin() + half, kv.end());
    for (size_t i = 0; i < cnt; i++) idx[lo + i] = kv[i].second;
    std::vector<std::pair<double, unsigned>>().swap(kv);
    if (depth < max_depth && cnt >= 2048) {
        // the two halves touch disjoint index ranges
        std::thread left([&] { bp_kd_order(norm, d, idx, lo, lo + half, max_depth, depth + 1); });
        bp_kd_order(norm, d, idx, lo + half, hi, max_depth, depth + 1);
        left.join();
    } else {
        bp_kd_order(norm, d, idx, lo, lo + half, max_depth, depth + 1);
        bp_kd_order(norm, d, idx, lo + half, hi, max_depth, depth + 1);
    }
}
static int bp_pivot_count()
{
    const char *e = getenv("HVB200_BP_PIVOTS");
    const int v = e ? atoi(e) : 320;
    return v < 1 ? 1 : (v > kBpMaxPivots ? kBpMaxPivots : v);
}
static size_t bp_npiv_for(size_t n)
{
    // pivots: up to 320, about an eighth of a small set (n=1000: 128 pivots 3.75 Tsp/s, 256 pivots 3.1)
    size_t npiv = std::min<size_t>(n, (size_t)bp_pivot_count());
    if (!getenv("HVB200_BP_PIVOTS")) npiv = std::min<size_t>(npiv, std::max<size_t>(16, n / 8));
    return npiv;
}

// What the pruned kernel's preparation builds on the host.  It depends on the transformed point set only, so the
// plans that hvb200_hv_approx_multi creates for the same set on several GPUs build it once and share it (BpShare).
struct BpHostPrep {
    std::vector<double> sorted;          // [n_pad][d] points in kd order, zero padded
    std::vector<float> piv32;            // [npiv][d]
    std::vector<unsigned> pividx;        // [npiv] row of each pivot in `sorted`
    std::vector<unsigned> pkeys;         // [nb_pad][W][32]
    std::vector<unsigned> bpair;         // [nb_pad/2][d]
    double qscale[HVB200_MAXD], qlos[HVB200_MAXD];
    int rc = 0;
};
struct BpShare { std::once_flag once; BpHostPrep prep; };
extern "C" void *hvb200cu_share_new(void) { return new (std::nothrow) BpShare(); }
extern "C" void hvb200cu_share_free(void *p) { delete (BpShare *)p; }

static int bp_alloc_outputs(DevPlan *dp, size_t pkeys_words, size_t bpair_words);
static int bp_device_build(DevPlan *dp);
// Stage 1, at the start of a run: geometry (it depends on n and d only) and the probe kernel, asynchronously.  The
// host part (stage 2) runs when the first max-min launch is due — by then the direction generator of the first
// batch has been queued, so the 1.6 ms of host work at cfg3 hide behind it instead of idling the GPU.
static int bp_prepare_begin(DevPlan *dp)
{
    const size_t n = dp->n;
    const int d = dp->d;
    cudaStream_t us = dp->own_stream;
    unsigned nprobe = 4096;
    while (nprobe > 256 && (double)nprobe * (double)n > 6e8) nprobe /= 2;
    if (dp->wins_cap < 2 * n) {
        cudaFree(dp->wins);
        dp->wins = nullptr; dp->wins_cap = 0;
        CUDA_TRY(cudaMalloc(&dp->wins, 2 * n * sizeof(unsigned)));
        dp->wins_cap = 2 * n;
    }
    const size_t npiv = bp_npiv_for(n);
    const int nb = (int)((n + kBpBlock - 1) / kBpBlock);
    const int nb_pad = (nb + 2 * kBpG - 1) / (2 * kBpG) * (2 * kBpG);
    dp->bp_nb = nb;
    dp->bp_nb_pad = nb_pad;
    dp->bp_npiv = (int)npiv;
    // segment = as many blocks as fit the shared-memory budget of one CTA (3 CTAs per SM up to d = 16, else 2)
    const size_t budget = (d <= 16 ? (HVB200_BP_MINB >= 4 ? 55 : 74) : 112) * 1024;
    int seg = (nb_pad + 31) / 32 * 32;
    while (seg > 32 && bp_smem_layout(d, bp_R_of(d), seg, (int)npiv).total > budget) seg -= 32;
    if (const char *e = getenv("HVB200_BP_SEG")) { const int v = atoi(e) / 32 * 32; if (v >= 32 && v < seg) seg = v; }
    dp->bp_seg_blocks = seg;
    dp->bp_smem = bp_smem_layout(d, bp_R_of(d), seg, (int)npiv).total;
    dp->bp_probe_queued = false;
    // HVB200_BP_PREP=host keeps round 1's host build (kd order by nth_element, keys on the host) as a cross-check
    const char *where = getenv("HVB200_BP_PREP");
    dp->bp_on_device = !(where && !strcmp(where, "host"));
    if (!dp->bp_on_device && !dp->share) {
        // (a shared host preparation probes on the device of whichever plan builds it, see bp_prepare_finish)
        CUDA_TRY(cudaMemsetAsync(dp->wins, 0, n * sizeof(unsigned), us));
        k_probe_argmax<<<(nprobe * 32 + 127) / 128, 128, 0, us>>>(dp->pts, (unsigned)n, d, nprobe, dp->wins);
        CUDA_TRY(cudaGetLastError());
        dp->bp_probe_queued = true;
    }
    if (dp->bp_on_device) {
        if (bp_alloc_outputs(dp, (size_t)nb_pad * ((d + 1) / 2) * kBpBlock, (size_t)(nb_pad / 2) * d)) return -1;
        for (int k = 0; k < HVB200_MAXD; k++) {
            dp->bp_qscale[k] = k < d ? dp->qscale[k] : 0.0;
            dp->bp_qlos[k] = k < d ? dp->qlo[k] * dp->qscale[k] * (1.0 + 0x1p-20) + 0x1p-10 : 0.0;
        }
        if (bp_device_build(dp)) return -1;
    }
    dp->bp_begun = true;
    return 0;
}

// FP32 probe pass (hvb200_bpprep.cuh): slices sized so that the grid holds a few waves of CTAs
static int probe32_launch(DevPlan *dp, BpPrepScratch *sc, unsigned nprobe)
{
    const size_t n = dp->n;
    const int d = dp->d;
    cudaStream_t us = dp->own_stream;
    const unsigned groups = (nprobe + kProbeThreads - 1) / kProbeThreads;
    unsigned nslices = std::max(1u, (unsigned)(dp->sm_count * 6) / groups);
    unsigned slice = (unsigned)((n + nslices - 1) / nslices);
    slice = std::max((slice + kProbeTile - 1) / kProbeTile * kProbeTile, (unsigned)kProbeTile);
    nslices = (unsigned)((n + slice - 1) / slice);
    if (sc->cap_probe < (size_t)nslices * nprobe) {
        cudaFree(sc->probe_partial); sc->probe_partial = nullptr; sc->cap_probe = 0;
        CUDA_TRY(cudaMalloc(&sc->probe_partial, (size_t)nslices * nprobe * sizeof(ull)));
        sc->cap_probe = (size_t)nslices * nprobe;
    }
    const dim3 grid(groups, nslices);
    auto launch = [&](auto dp_tag) {
        constexpr int DP = decltype(dp_tag)::value;
        k_probe32<DP><<<grid, kProbeThreads, 0, us>>>(dp->pts, (unsigned)n, d, nprobe, slice, sc->probe_partial);
    };
    // (objectives padded to the next size: the pad costs its share of the multiplies)
    if (d <= 4) launch(std::integral_constant<int, 4>());
    else if (d <= 6) launch(std::integral_constant<int, 6>());
    else if (d <= 8) launch(std::integral_constant<int, 8>());
    else if (d <= 10) launch(std::integral_constant<int, 10>());
    else if (d <= 12) launch(std::integral_constant<int, 12>());
    else if (d <= 16) launch(std::integral_constant<int, 16>());
    else if (d <= 20) launch(std::integral_constant<int, 20>());
    else if (d <= 24) launch(std::integral_constant<int, 24>());
    else if (d <= 32) launch(std::integral_constant<int, 32>());
    else if (d <= 48) launch(std::integral_constant<int, 48>());
    else launch(std::integral_constant<int, 64>());
    k_probe32_wins<<<(nprobe + 255) / 256, 256, 0, us>>>(sc->probe_partial, nprobe, nslices, (unsigned)n, dp->wins);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// device buffers the preparation fills: points in kd order, pivots, point keys, block bounds
static int bp_alloc_outputs(DevPlan *dp, size_t pkeys_words, size_t bpair_words)
{
    const int d = dp->d;
    const size_t n_pad = (size_t)dp->bp_nb_pad * kBpBlock;
    if (dp->bp_pts_cap < n_pad * (size_t)d) {
        cudaFree(dp->bp_pts);
        dp->bp_pts = nullptr; dp->bp_pts_cap = 0;
        CUDA_TRY(cudaMalloc(&dp->bp_pts, n_pad * (size_t)d * sizeof(double)));
        dp->bp_pts_cap = n_pad * (size_t)d;
    }
    if (!dp->bp_piv32) {
        CUDA_TRY(cudaMalloc(&dp->bp_piv32, (size_t)kBpMaxPivots * HVB200_MAXD * sizeof(float)));
        CUDA_TRY(cudaMalloc(&dp->bp_pividx, (size_t)kBpMaxPivots * sizeof(unsigned)));
    }
    if (dp->bp_keys_cap < pkeys_words) {
        cudaFree(dp->bp_pkeys);
        dp->bp_pkeys = nullptr; dp->bp_keys_cap = 0;
        CUDA_TRY(cudaMalloc(&dp->bp_pkeys, pkeys_words * sizeof(unsigned)));
        dp->bp_keys_cap = pkeys_words;
    }
    if (dp->bp_pair_cap < bpair_words) {
        cudaFree(dp->bp_bpair);
        dp->bp_bpair = nullptr; dp->bp_pair_cap = 0;
        CUDA_TRY(cudaMalloc(&dp->bp_bpair, bpair_words * sizeof(unsigned)));
        dp->bp_pair_cap = bpair_words;
    }
    return 0;
}

// The whole preparation on the device (hvb200_bpprep.cuh), queued on the plan's own stream; prep.done is recorded
// behind it.  Host work: the kd tree's shape (segment boundaries per level, from n alone) and ~8 launches per level.
static int bp_device_build(DevPlan *dp)
{
    HVB200_NVTX("hvb200:pruned_prepare_device");
    const size_t n = dp->n;
    const int d = dp->d;
    cudaStream_t us = dp->own_stream;
    BpPrepScratch *sc = &dp->prep;
    // ---- tree shape: a node of cnt > 32 points splits into [lo, lo + half) and [lo + half, hi), half = whole blocks
    std::vector<unsigned> all;                      // the levels' start arrays, one after the other (each ends with n)
    std::vector<size_t> level_off;
    std::vector<unsigned> level_nseg;
    size_t local_off = 0;
    unsigned local_nseg = 0;
    {
        std::vector<unsigned> cur = {0u, (unsigned)n}, next;
        for (;;) {
            bool any = false;
            next.clear();
            for (size_t s2 = 0; s2 + 1 < cur.size(); s2++) {
                const unsigned lo = cur[s2], cnt = cur[s2 + 1] - lo;
                next.push_back(lo);
                if (cnt > (unsigned)kBpBlock) {
                    unsigned half = (cnt / 2 + kBpBlock - 1) / kBpBlock * kBpBlock;
                    if (half >= cnt) half = cnt / 2;
                    next.push_back(lo + half);
                    any = true;
                }
            }
            next.push_back((unsigned)n);
            unsigned biggest = 0;
            for (size_t s2 = 0; s2 + 1 < cur.size(); s2++) biggest = std::max(biggest, cur[s2 + 1] - cur[s2]);
            if (!any || biggest <= (unsigned)kKdLocalMax) {
                // the rest of the tree is finished inside shared memory, one CTA per segment of this level
                local_off = all.size();
                local_nseg = any ? (unsigned)cur.size() - 1 : 0;
                all.insert(all.end(), cur.begin(), cur.end());
                break;
            }
            level_off.push_back(all.size());
            level_nseg.push_back((unsigned)cur.size() - 1);
            all.insert(all.end(), cur.begin(), cur.end());
            cur.swap(next);
        }
    }
    unsigned max_seg = 1;
    for (unsigned v : level_nseg) max_seg = std::max(max_seg, v);
    // ---- scratch
    if (sc->cap_n < n) {
        for (int i = 0; i < 2; i++) { cudaFree(sc->idx[i]); cudaFree(sc->keys[i]); cudaFree(sc->pkeys64[i]); sc->idx[i] = sc->keys[i] = nullptr; sc->pkeys64[i] = nullptr; }
        cudaFree(sc->pos); sc->pos = nullptr; sc->cap_n = 0;
        for (int i = 0; i < 2; i++) {
            CUDA_TRY(cudaMalloc(&sc->idx[i], n * sizeof(unsigned)));
            CUDA_TRY(cudaMalloc(&sc->keys[i], n * sizeof(unsigned)));
            CUDA_TRY(cudaMalloc(&sc->pkeys64[i], n * sizeof(ull)));
        }
        CUDA_TRY(cudaMalloc(&sc->pos, n * sizeof(unsigned)));
        sc->cap_n = n;
    }
    if (sc->cap_starts < all.size() + 2) {
        cudaFree(sc->starts); sc->starts = nullptr; sc->cap_starts = 0;
        CUDA_TRY(cudaMalloc(&sc->starts, (all.size() + 2 + 256) * sizeof(unsigned)));
        sc->cap_starts = all.size() + 2 + 256;
    }
    if (sc->cap_minmax < (size_t)max_seg * d * 2) {
        cudaFree(sc->minmax); sc->minmax = nullptr; sc->cap_minmax = 0;
        CUDA_TRY(cudaMalloc(&sc->minmax, (size_t)max_seg * d * 2 * sizeof(ull)));
        sc->cap_minmax = (size_t)max_seg * d * 2;
    }
    size_t b1 = 0, b2 = 0;
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, b1, (const unsigned *)nullptr, (unsigned *)nullptr, (const unsigned *)nullptr, (unsigned *)nullptr, (int)n, 0, 32, us));
    CUDA_TRY(cub::DeviceRadixSort::SortKeys(nullptr, b2, (const ull *)nullptr, (ull *)nullptr, (int)n, 0, 64, us));
    size_t b3 = 0;
    CUDA_TRY(cub::DeviceRadixSort::SortKeys(nullptr, b3, (const unsigned *)nullptr, (unsigned *)nullptr, (int)n, 0, 32, us));
    const size_t need_tmp = std::max(std::max(b1, b2), b3);
    if (sc->cub_bytes < need_tmp) {
        cudaFree(sc->cub_tmp); sc->cub_tmp = nullptr; sc->cub_bytes = 0;
        CUDA_TRY(cudaMalloc(&sc->cub_tmp, need_tmp));
        sc->cub_bytes = need_tmp;
    }
    if (!sc->done) CUDA_TRY(cudaEventCreateWithFlags(&sc->done, cudaEventDisableTiming));
    // the level table goes up first: a copy from pageable memory waits for the work already queued on its stream
    if (!all.empty()) CUDA_TRY(cudaMemcpyAsync(sc->starts, all.data(), all.size() * sizeof(unsigned), cudaMemcpyHostToDevice, us));
    const double *qlo = dp->qdev, *qhi = dp->qdev + HVB200_MAXD, *qscale = dp->qdev + 2 * HVB200_MAXD;
    const unsigned gn = (unsigned)((n + 255) / 256);
    unsigned nprobe = 4096;
    while (nprobe > 256 && (double)nprobe * (double)n > 6e8) nprobe /= 2;
    // ---- probe directions: how often is each point the arg-max (the pivots are the most frequent winners)
    CUDA_TRY(cudaMemsetAsync(dp->wins, 0, n * sizeof(unsigned), us));
    if (probe32_launch(dp, sc, nprobe)) return -1;
    dp->bp_probe_queued = true;
    // ---- kd order
    k_iota_u32<<<gn, 256, 0, us>>>(sc->idx[0], (unsigned)n);
    int cur = 0;
    for (size_t L = 0; L < level_nseg.size(); L++) {
        const unsigned nseg = level_nseg[L];
        const unsigned *starts = sc->starts + level_off[L];
        int nbits = 0;
        while ((1u << nbits) < nseg) nbits++;
        const int vb = std::min(std::max(kKdValueBitsMin, kKdKeyBits - nbits), 32 - nbits);
        const ull cells = (ull)nseg * d * 2;
        k_fill_u64<<<(unsigned)((cells + 255) / 256), 256, 0, us>>>(sc->minmax, cells, ~0ull, 0ull);
        k_kd_minmax<<<gn, 256, 0, us>>>(dp->pts, sc->idx[cur], (unsigned)n, d, starts, nseg, sc->minmax);
        k_kd_keys<<<gn, 256, 0, us>>>(dp->pts, sc->idx[cur], (unsigned)n, d, starts, nseg, sc->minmax, qhi, vb, sc->keys[0]);
        size_t tmp = sc->cub_bytes;
        CUDA_TRY(cub::DeviceRadixSort::SortPairs(sc->cub_tmp, tmp, (const unsigned *)sc->keys[0], sc->keys[1], (const unsigned *)sc->idx[cur], sc->idx[cur ^ 1],
                                                 (int)n, 0, vb + nbits, us));
        cur ^= 1;
    }
    if (local_nseg) k_kd_local<<<local_nseg, kKdLocalThreads, 0, us>>>(dp->pts, sc->idx[cur], d, sc->starts + local_off, qhi);
    // ---- points, keys, bounds in kd order
    const ull n_pad = (ull)dp->bp_nb_pad * kBpBlock;
    const int W = (d + 1) / 2;
    k_bp_gather<<<(unsigned)((n_pad * d + 255) / 256), 256, 0, us>>>(dp->pts, sc->idx[cur], (ull)n, n_pad, d, dp->bp_pts, sc->pos);
    k_bp_keys<<<(unsigned)(((ull)dp->bp_nb_pad * W * kBpBlock + 255) / 256), 256, 0, us>>>(dp->bp_pts, (ull)n, dp->bp_nb_pad, d, qlo, qscale, dp->bp_pkeys);
    k_bp_bounds<<<(unsigned)(((ull)(dp->bp_nb_pad / 2) * W * 32 + 255) / 256), 256, 0, us>>>(dp->bp_pkeys, dp->bp_nb_pad, d, dp->bp_bpair);
    // ---- pivots
    if (n <= (1u << 19) && nprobe <= 4096) {
        unsigned *k0 = reinterpret_cast<unsigned *>(sc->pkeys64[0]), *k1 = reinterpret_cast<unsigned *>(sc->pkeys64[1]);
        k_bp_pivot_keys<unsigned, 19><<<gn, 256, 0, us>>>(dp->wins, (unsigned)n, nprobe, k0);
        size_t tmp = sc->cub_bytes;
        CUDA_TRY(cub::DeviceRadixSort::SortKeys(sc->cub_tmp, tmp, (const unsigned *)k0, k1, (int)n, 0, 32, us));
        k_bp_pivots<unsigned, 19><<<(unsigned)((dp->bp_npiv * d + 255) / 256), 256, 0, us>>>(dp->pts, k1, sc->pos, dp->bp_npiv, d, dp->bp_piv32, dp->bp_pividx);
    } else {
        k_bp_pivot_keys<ull, 32><<<gn, 256, 0, us>>>(dp->wins, (unsigned)n, nprobe, sc->pkeys64[0]);
        size_t tmp = sc->cub_bytes;
        CUDA_TRY(cub::DeviceRadixSort::SortKeys(sc->cub_tmp, tmp, (const ull *)sc->pkeys64[0], sc->pkeys64[1], (int)n, 0, 46, us));
        k_bp_pivots<ull, 32><<<(unsigned)((dp->bp_npiv * d + 255) / 256), 256, 0, us>>>(dp->pts, sc->pkeys64[1], sc->pos, dp->bp_npiv, d, dp->bp_piv32, dp->bp_pividx);
    }
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(sc->done, us));
    return 0;
}

static int bp_host_build(DevPlan *dp, BpHostPrep &H)
{
    HVB200_NVTX("hvb200:pruned_prepare_host");
    const size_t n = dp->n;
    const int d = dp->d, W = (d + 1) / 2;
    cudaStream_t us = dp->own_stream;
    const bool dbg = getenv("HVB200_DEBUG_TIMING") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = now();
    unsigned nprobe = 4096;
    while (nprobe > 256 && (double)nprobe * (double)n > 6e8) nprobe /= 2;
    if (!dp->bp_probe_queued) {
        CUDA_TRY(cudaMemsetAsync(dp->wins, 0, n * sizeof(unsigned), us));
        k_probe_argmax<<<(nprobe * 32 + 127) / 128, 128, 0, us>>>(dp->pts, (unsigned)n, d, nprobe, dp->wins);
        CUDA_TRY(cudaGetLastError());
        dp->bp_probe_queued = true;
    }
    std::vector<unsigned> wins(n);
    std::vector<double> pts_rb;
    if (dp->host_pts.size() != n * (size_t)d || dp->ordered) {
        // built on the device, or reordered for the DPX screen since: read the points back
        pts_rb.resize(n * (size_t)d);
        CUDA_TRY(cudaMemcpyAsync(pts_rb.data(), dp->pts, n * (size_t)d * sizeof(double), cudaMemcpyDeviceToHost, us));
        CUDA_TRY(cudaStreamSynchronize(us));
    }
    const std::vector<double> &pts = pts_rb.empty() ? dp->host_pts : pts_rb;
    if (dbg) fprintf(stderr, "bp_prepare: d2h %.2f ms\n", now() - t0);
    // key map of the point set (the same monotone map as the DPX screen, any d <= 64)
    std::vector<double> qlo(d, INFINITY), qhi(d, -INFINITY), qscale(d, 0.0);
    for (size_t i = 0; i < n; i++)
        for (int k = 0; k < d; k++) {
            const double v = pts[i * d + k];
            if (v < qlo[k]) qlo[k] = v;
            if (v > qhi[k]) qhi[k] = v;
        }
    for (int k = 0; k < HVB200_MAXD; k++) { H.qscale[k] = 0.0; H.qlos[k] = 0.0; }
    for (int k = 0; k < d; k++) {
        const double range = qhi[k] - qlo[k];
        qscale[k] = (range > 0 && std::isfinite(range)) ? (double)kKeyScale / range : 0.0;
        if (!std::isfinite(qscale[k])) qscale[k] = 0.0;
        H.qscale[k] = qscale[k];
        H.qlos[k] = qlo[k] * qscale[k] * (1.0 + 0x1p-20) + 0x1p-10;
    }
    // kd-tree order on the normalised coordinates
    std::vector<double> norm(n * (size_t)d);
    for (size_t i = 0; i < n; i++)
        for (int k = 0; k < d; k++) norm[i * d + k] = qhi[k] > 0 ? pts[i * d + k] / qhi[k] : 0.0;
    std::vector<unsigned> idx(n);
    for (size_t i = 0; i < n; i++) idx[i] = (unsigned)i;
    if (dbg) fprintf(stderr, "bp_prepare: ranges+norm %.2f ms\n", now() - t0);
    bp_kd_order(norm, d, idx, 0, n, bp_kd_depth());
    if (dbg) fprintf(stderr, "bp_prepare: kd %.2f ms\n", now() - t0);
    std::vector<unsigned> pos(n);
    for (size_t i = 0; i < n; i++) pos[idx[i]] = (unsigned)i;
    const int nb = dp->bp_nb, nb_pad = dp->bp_nb_pad;
    const size_t n_pad = (size_t)nb_pad * kBpBlock;
    H.sorted.assign(n_pad * (size_t)d, 0.0);
    std::vector<double> &sorted = H.sorted;
    for (size_t i = 0; i < n; i++) memcpy(&sorted[i * d], &pts[(size_t)idx[i] * d], d * sizeof(double));
    // point keys KP = clamp(ceil f_k(p), 0, S); padding points -1, dummy objective of an odd d = S (always passes)
    std::vector<unsigned short> key(n_pad * (size_t)(2 * W));
    for (size_t i = 0; i < n_pad; i++)
        for (int k = 0; k < 2 * W; k++) {
            unsigned short v;
            if (i >= n) v = 0xffffu;
            else if (k >= d) v = (unsigned short)kKeyScale;
            else {
                double f = ceil((sorted[i * d + k] - qlo[k]) * qscale[k]);
                f = f < 0.0 ? 0.0 : (f > (double)kKeyScale ? (double)kKeyScale : f);
                v = (unsigned short)(int)f;
            }
            key[i * (2 * W) + k] = v;
        }
    H.pkeys.resize((size_t)nb_pad * W * kBpBlock);
    for (int b = 0; b < nb_pad; b++)
        for (int j = 0; j < W; j++)
            for (int l = 0; l < kBpBlock; l++) {
                const size_t i = (size_t)b * kBpBlock + l;
                H.pkeys[((size_t)b * W + j) * kBpBlock + l] = (unsigned)key[i * (2 * W) + 2 * j] | ((unsigned)key[i * (2 * W) + 2 * j + 1] << 16);
            }
    // block bounds: per-objective maximum key (signed: padding -1 never raises it), two blocks per word
    H.bpair.resize((size_t)(nb_pad / 2) * d);
    for (int pr = 0; pr < nb_pad / 2; pr++)
        for (int k = 0; k < d; k++) {
            unsigned word = 0;
            for (int h = 0; h < 2; h++) {
                const int b = 2 * pr + h;
                int m = -1;
                if (b < nb)
                    for (int l = 0; l < kBpBlock; l++) m = std::max(m, (int)(short)key[((size_t)b * kBpBlock + l) * (2 * W) + k]);
                word |= ((unsigned)m & 0xffffu) << (16 * h);
            }
            H.bpair[(size_t)pr * d + k] = word;
        }
    if (dbg) fprintf(stderr, "bp_prepare: sorted+keys %.2f ms\n", now() - t0);
    // (a copy into pageable memory blocks until the probe kernel is done: issued only now, after the kd build)
    CUDA_TRY(cudaMemcpyAsync(wins.data(), dp->wins, n * sizeof(unsigned), cudaMemcpyDeviceToHost, us));
    CUDA_TRY(cudaStreamSynchronize(us));               // probe wins are on the host now
    if (dbg) fprintf(stderr, "bp_prepare: probe wait %.2f ms\n", now() - t0);
    // pivots: the points that won most probes
    const size_t npiv = (size_t)dp->bp_npiv;
    std::vector<unsigned> byw(n);
    for (size_t i = 0; i < n; i++) byw[i] = (unsigned)i;
    std::partial_sort(byw.begin(), byw.begin() + npiv, byw.end(),
                      [&](unsigned x, unsigned y) { return wins[x] != wins[y] ? wins[x] > wins[y] : x < y; });
    H.piv32.resize(npiv * (size_t)d);
    H.pividx.resize(npiv);
    for (size_t i = 0; i < npiv; i++) {
        H.pividx[i] = pos[byw[i]];
        for (int k = 0; k < d; k++) H.piv32[i * d + k] = (float)pts[(size_t)byw[i] * d + k];
    }
    if (dbg) fprintf(stderr, "bp_prepare: pivots %.2f ms\n", now() - t0);
    return 0;
}

// Stage 2: host build (own or shared) + uploads.  Called by the first max-min launch of the plan.
static int bp_prepare_finish(DevPlan *dp)
{
    HVB200_NVTX("hvb200:pruned_prepare");
    cudaStream_t us = dp->own_stream;
    if (dp->bp_on_device) {
        // everything was queued by bp_prepare_begin: the run's stream waits for it on the device, the host does not
        if (dp->stream != us) CUDA_TRY(cudaStreamWaitEvent(dp->stream, dp->prep.done, 0));
        dp->bp_ready = true;
        return 0;
    }
    BpHostPrep own;
    BpHostPrep *H = &own;
    if (dp->share) {
        BpShare *sh = (BpShare *)dp->share;
        std::call_once(sh->once, [&] { sh->prep.rc = bp_host_build(dp, sh->prep); });
        H = &sh->prep;
        if (H->rc) return -1;
        CUDA_TRY(cudaSetDevice(dp->device));
    } else if (bp_host_build(dp, own)) return -1;
    if (bp_alloc_outputs(dp, H->pkeys.size(), H->bpair.size())) return -1;
    for (int k = 0; k < HVB200_MAXD; k++) { dp->bp_qscale[k] = H->qscale[k]; dp->bp_qlos[k] = H->qlos[k]; }
    CUDA_TRY(cudaMemcpyAsync(dp->bp_pts, H->sorted.data(), H->sorted.size() * sizeof(double), cudaMemcpyHostToDevice, us));
    CUDA_TRY(cudaMemcpyAsync(dp->bp_piv32, H->piv32.data(), H->piv32.size() * sizeof(float), cudaMemcpyHostToDevice, us));
    CUDA_TRY(cudaMemcpyAsync(dp->bp_pividx, H->pividx.data(), H->pividx.size() * sizeof(unsigned), cudaMemcpyHostToDevice, us));
    CUDA_TRY(cudaMemcpyAsync(dp->bp_pkeys, H->pkeys.data(), H->pkeys.size() * sizeof(unsigned), cudaMemcpyHostToDevice, us));
    CUDA_TRY(cudaMemcpyAsync(dp->bp_bpair, H->bpair.data(), H->bpair.size() * sizeof(unsigned), cudaMemcpyHostToDevice, us));
    CUDA_TRY(cudaStreamSynchronize(us));              // host vectors go out of scope; the run's stream may be another one
    dp->bp_ready = true;
    return 0;
}

// Sort the plan's points by how often they win a probe direction (see k_probe_argmax).  Runs once per
// plan, only for jobs large enough to amortise it (~0.5 ms at n = 1e4).
static int order_points(DevPlan *dp)
{
    const size_t n = dp->n;
    const int d = dp->d;
    unsigned nprobe = 4096;
    while (nprobe > 256 && (double)nprobe * (double)n > 6e8) nprobe /= 2;
    cudaStream_t us = dp->own_stream;
    if (dp->wins_cap < 2 * n) {
        cudaFree(dp->wins);
        dp->wins = nullptr; dp->wins_cap = 0;
        CUDA_TRY(cudaMalloc(&dp->wins, 2 * n * sizeof(unsigned)));
        dp->wins_cap = 2 * n;
    }
    if (dp->pts_tmp_cap < n * (size_t)d) {
        cudaFree(dp->pts_tmp);
        dp->pts_tmp = nullptr; dp->pts_tmp_cap = 0;
        CUDA_TRY(cudaMalloc(&dp->pts_tmp, n * (size_t)d * sizeof(double)));
        dp->pts_tmp_cap = n * (size_t)d;
    }
    CUDA_TRY(cudaMemsetAsync(dp->wins, 0, n * sizeof(unsigned), us));
    k_probe_argmax<<<(nprobe * 32 + 127) / 128, 128, 0, us>>>(dp->pts, (unsigned)n, d, nprobe, dp->wins);
    CUDA_TRY(cudaGetLastError());
    std::vector<unsigned> wins(n), order(n);
    CUDA_TRY(cudaMemcpyAsync(wins.data(), dp->wins, n * sizeof(unsigned), cudaMemcpyDeviceToHost, us));
    CUDA_TRY(cudaStreamSynchronize(us));
    // stable counting sort by descending win count (counts <= nprobe)
    std::vector<unsigned> start(nprobe + 2, 0);
    for (size_t i = 0; i < n; i++) start[nprobe - std::min(wins[i], nprobe) + 1]++;
    for (unsigned c = 0; c <= nprobe; c++) start[c + 1] += start[c];
    for (size_t i = 0; i < n; i++) order[start[nprobe - std::min(wins[i], nprobe)]++] = (unsigned)i;
    CUDA_TRY(cudaMemcpyAsync(dp->wins + n, order.data(), n * sizeof(unsigned), cudaMemcpyHostToDevice, us));
    k_gather_rows<<<(unsigned)((n * d + 255) / 256), 256, 0, us>>>(dp->pts, dp->wins + n, (ull)n, d, dp->pts_tmp);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(dp->pts, dp->pts_tmp, n * (size_t)d * sizeof(double), cudaMemcpyDeviceToDevice, us));
    if (dp->keys) {
        const size_t total_pairs = (size_t)dp->k_n_tiles * dp->k_tile_pairs;
        const ull words = (ull)total_pairs * d;
        k_quantize<<<(unsigned)((words + 255) / 256), 256, 0, us>>>(dp->pts, (ull)n, d, dp->k_tile_pairs, (ull)total_pairs,
                                                                   dp->qdev, dp->qdev + HVB200_MAXD, dp->qdev + 2 * HVB200_MAXD, dp->keys);
        CUDA_TRY(cudaGetLastError());
    }
    CUDA_TRY(cudaStreamSynchronize(us));           // `order` is host memory; later runs may use another stream
    dp->ordered = true;
    return 0;
}

// The dynamic shared-memory limit is an attribute of the KERNEL FUNCTION, shared by all host threads: it is always
// raised to the device maximum (never to the size one call needs — a concurrent call with a smaller point set
// would lower it under a launch in flight: "invalid argument").  The limit does not affect occupancy.
static int max_dyn_smem_limit()
{
    static int limit = 0;
    if (limit == 0) {
        int dev = 0, v = 0;
        // minus 1 KB: kernels with a few bytes of static shared memory must stay within the opt-in total
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) == cudaSuccess && v > 2048) limit = v - 1024;
        else limit = 226 * 1024;
    }
    return limit;
}

extern "C" int hvb200cu_run_begin(struct hvb200_plan *plan, void *stream, uint64_t max_batch, uint64_t *batch_out)
{
    DevPlan *dp = (DevPlan *)plan->dev;
    CUDA_TRY(cudaSetDevice(dp->device));
    dp->stream = stream ? (cudaStream_t)stream : dp->own_stream;
    dp->spans.clear();
    dp->ev_next = 0;
    memset(&plan->stats, 0, sizeof plan->stats);
    const int d = dp->d;
    // batch size: bounded by 2 GiB of direction buffer and by what the caller needs
    uint64_t batch = (2ull << 30) / (uint64_t)(16 * d);
    {
        const char *e = getenv("HVB200_BATCH");          // samples per launch (development knob)
        if (e && atoll(e) > 0) batch = (uint64_t)atoll(e);
    }
    batch -= batch % (uint64_t)(kConsumerThreads * 4);
    if (batch > max_batch) batch = max_batch;
    if (batch < 1) batch = 1;
    *batch_out = batch;

    // kernel selection + launch geometry
    // AUTO: the screens pay a divergent exact evaluation per candidate, which only amortises over a
    // few hundred points (measured on B200: n=100,d=3 EXACT 602 / DPX 243 / SCREEN 371 Gsp/s;
    // n=1000,d=5 EXACT 734 / SCREEN 930 / DPX 1277; n=1e4,d=10 EXACT 484 / SCREEN 1240 / DPX 1946).
    // The block-pruned kernel wins from n ~ 500 upwards (n=1000,d=5: 3.9 Tsp/s vs DPX 1.55; n=1e4,d=10: 9.8 vs
    // 2.3; n=5e4,d=20: 4.4 vs 1.0; d=50: 0.64 vs 0.0045 for the runtime-d kernel) but its plan-time
    // preparation (probe kernel, kd order and keys on the host) costs ~23 ns per coordinate of the point set
    // (2.3 ms at n=1e4,d=10; 20 ms at n=5e4,d=20), which the ~3x faster kernel earns back after about
    // 7e4*d directions whatever n is.
    int mode = MODE_DPX;
    if (plan->kernel == HVB200_KERNEL_EXACT) mode = MODE_EXACT;
    else if (plan->kernel == HVB200_KERNEL_SCREEN) mode = MODE_SCREEN;
    else if (plan->kernel == HVB200_KERNEL_BP) mode = MODE_BP;
    else if (plan->kernel == HVB200_KERNEL_AUTO) {
        if (dp->n < 256) mode = MODE_EXACT;
        else if (d > 32) mode = MODE_BP;
        // Round 2: the preparation runs on the device (0.2-0.4 ms of kernels), so whole calls — plan creation from host
        // buffers + one run — are faster with the pruned kernel from n*d ~ 16 K coordinates on whatever the run's length
        // (dev/auto_choice.py, pruned / DPX call time at 16 K ... 1 M directions: n=2000,d=10 0.90 ... 0.58; n=5000,d=10
        // 0.89 ... 0.45; n=2000,d=30 0.27 ... 0.33; n=20000,d=20 0.51 ... 0.34; n=1e5,d=3 0.67 ... 0.22), while small sets
        // (n=1000,d=5: 1.12 ... 0.67) still need ~65 K*d directions to earn it back.
        else if (dp->n >= 512 && (dp->bp_ready || plan->run_samples >= (uint64_t)65536 * (uint64_t)d ||
                                  (uint64_t)dp->n * (uint64_t)d >= 16384)) mode = MODE_BP;
    }
    if (mode == MODE_BP && dp->n / 32 + 64 >= ((size_t)1 << kBpPtBits)) mode = MODE_DPX;   // hit entries hold 23-bit granule (32-point) indices
    if (d > 32 && mode != MODE_BP) mode = MODE_EXACT;
    dp->kernel_used = mode;
    if ((mode == MODE_DPX || mode == MODE_SCREEN || mode == MODE_EXACT) && d <= 32 && !dp->ordered && dp->n >= 1024 && plan->run_samples >= (1u << 20) &&
        !getenv("HVB200_NO_ORDER")) {
        if (order_points(dp)) return -1;
    }
    if (mode == MODE_BP && !dp->bp_ready && bp_prepare_begin(dp)) return -1;
    const size_t smem = mode == MODE_DPX ? dp->k_smem_bytes : (mode == MODE_BP ? dp->bp_smem : dp->smem_bytes);
    int occ = 1;
    dp->fused = mode == MODE_BP && plan->want_fused_hw && d <= 32 && !getenv("HVB200_NO_FUSED");
    if (!dp->fused && dp->dirs_cap < batch * (size_t)(2 * d)) {       // (a fused run has no direction buffer)
        cudaFree(dp->dirs);
        dp->dirs = nullptr; dp->dirs_cap = 0;
        CUDA_TRY(cudaMalloc(&dp->dirs, batch * (size_t)(2 * d) * sizeof(double)));
        dp->dirs_cap = batch * (size_t)(2 * d);
    }
    if (mode == MODE_BP) {
        bp_fn fn = dp->fused ? pick_bp_gen<32>(d) : pick_bp<HVB200_MAXD>(d);
        CUDA_TRY(cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn_smem_limit()));
        CUDA_TRY(cudaFuncSetAttribute((const void *)fn, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void *)fn, kBpThreads, smem));
    } else if (mode == MODE_DPX) {
        dpx_fn fn = pick_dpx<32>(d);
        CUDA_TRY(cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn_smem_limit()));
        CUDA_TRY(cudaFuncSetAttribute((const void *)fn, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void *)fn, kBlockThreads, smem));
    } else {
        maxmin_fn fn = pick_kernel(d, mode);
        CUDA_TRY(cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn_smem_limit()));
        CUDA_TRY(cudaFuncSetAttribute((const void *)fn, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void *)fn, kBlockThreads, smem));
    }
    if (occ < 1) { hvb200_set_error("max-min kernel does not fit on an SM (d=%d)", d); return -1; }
    dp->grid = dp->sm_count * occ;
    if (mode == MODE_BP && dp->bp_hits_cap < (size_t)dp->grid * kBpHitCap) {
        cudaFree(dp->bp_hits);
        dp->bp_hits = nullptr; dp->bp_hits_cap = 0;
        CUDA_TRY(cudaMalloc(&dp->bp_hits, (size_t)dp->grid * kBpHitCap * sizeof(uint2)));
        dp->bp_hits_cap = (size_t)dp->grid * kBpHitCap;
    }
    if (mode == MODE_BP) {
        // worst case: every direction of a chunk survives in every block of a segment; 12 bytes per quad
        const size_t need = (size_t)dp->grid * (size_t)dp->bp_seg_blocks * (size_t)(kBpThreads * bp_R_of(d) / 4) * 12;
        if (dp->bp_quads_cap < need) {
            cudaFree(dp->bp_quads);
            dp->bp_quads = nullptr; dp->bp_quads_cap = 0;
            CUDA_TRY(cudaMalloc(&dp->bp_quads, need));
            dp->bp_quads_cap = need;
        }
    }
    if (dp->fused) {
        const size_t ch = (size_t)kBpThreads * bp_R_of(d);
        const size_t need = (size_t)dp->grid * ((size_t)d * ch + (size_t)d * ch / 2);      // r FP64 + winv FP32 per direction
        if (dp->gen_scratch_cap < need) {
            cudaFree(dp->gen_scratch);
            dp->gen_scratch = nullptr; dp->gen_scratch_cap = 0;
            CUDA_TRY(cudaMalloc(&dp->gen_scratch, need * sizeof(double)));
            dp->gen_scratch_cap = need;
        }
    }
    if (mode == MODE_BP) {
        const size_t ch = (size_t)kBpThreads * bp_R_of(d);
        const size_t need = (size_t)(kBpThreads / 32) * ((size_t)(plan->run_total / ch) + (size_t)(plan->run_total / batch) + 8);
        if (dp->bp_chunk_cap < need) {
            cudaFree(dp->bp_chunk_sums);
            dp->bp_chunk_sums = nullptr; dp->bp_chunk_cap = 0;
            CUDA_TRY(cudaMalloc(&dp->bp_chunk_sums, need * sizeof(double)));
            dp->bp_chunk_cap = need;
        }
        if (!dp->bp_chunk_counter) CUDA_TRY(cudaMalloc(&dp->bp_chunk_counter, sizeof(unsigned)));
        dp->bp_chunk_off = 0;
    }
    if (dp->cta_cap < dp->grid) {
        cudaFree(dp->cta_sums);
        CUDA_TRY(cudaMalloc(&dp->cta_sums, dp->grid * sizeof(double)));
        dp->cta_cap = dp->grid;
    }
    CUDA_TRY(cudaMemsetAsync(dp->cta_sums, 0, dp->cta_cap * sizeof(double), dp->stream));
    CUDA_TRY(cudaMemsetAsync(dp->counters, 0, 8 * sizeof(ull), dp->stream));
    return 0;
}

// Pinned host staging for host-produced streams (MC normals, Rphi u, uploaded directions).
// Two buffers alternate: the host fills buffer b+1 while the copy of buffer b is in flight; a buffer
// is handed out again only after the event recorded behind its last H2D copy has completed.
extern "C" double *hvb200cu_staging(struct hvb200_plan *plan, size_t ndoubles)
{
    DevPlan *dp = (DevPlan *)plan->dev;
    const int i = dp->staging_next;
    dp->staging_next ^= 1;
    dp->staging_cur = i;
    if (dp->staging_free[i] == nullptr) {
        if (cudaEventCreateWithFlags(&dp->staging_free[i], cudaEventDisableTiming) != cudaSuccess) {
            hvb200_set_error("cudaEventCreate failed");
            cudaGetLastError();
            return nullptr;
        }
    } else if (cudaEventSynchronize(dp->staging_free[i]) != cudaSuccess) {
        hvb200_set_error("cudaEventSynchronize(staging) failed: %s", cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    if (dp->staging_cap[i] < ndoubles) {
        if (dp->staging_host[i]) cudaFreeHost(dp->staging_host[i]);
        dp->staging_host[i] = nullptr; dp->staging_cap[i] = 0;
        if (cudaMallocHost(&dp->staging_host[i], ndoubles * sizeof(double)) != cudaSuccess) {
            hvb200_set_error("cudaMallocHost(%zu bytes) failed", ndoubles * sizeof(double));
            cudaGetLastError();
            return nullptr;
        }
        dp->staging_cap[i] = ndoubles;
    }
    return dp->staging_host[i];
}

static int stage_to_device(struct hvb200_plan *plan, DevPlan *dp, const double *host, size_t ndoubles)
{
    if (dp->staging_dev_cap < ndoubles) {
        cudaFree(dp->staging_dev);
        dp->staging_dev = nullptr; dp->staging_dev_cap = 0;
        CUDA_TRY(cudaMalloc(&dp->staging_dev, ndoubles * sizeof(double)));
        dp->staging_dev_cap = ndoubles;
    }
    const int b = span_begin(dp, 3);
    CUDA_TRY(cudaMemcpyAsync(dp->staging_dev, host, ndoubles * sizeof(double), cudaMemcpyHostToDevice, dp->stream));
    span_end(dp, b);
    // the pinned buffer may be refilled once this copy has drained; a caller-owned (pageable) source is
    // copied synchronously by the runtime, so the event is harmless there
    if (host == dp->staging_host[dp->staging_cur] && dp->staging_free[dp->staging_cur])
        CUDA_TRY(cudaEventRecord(dp->staging_free[dp->staging_cur], dp->stream));
    plan->stats.h2d_bytes += ndoubles * sizeof(double);
    return 0;
}

extern "C" int hvb200cu_gen_hw(struct hvb200_plan *plan, const hvb200_hw_params *hw, uint64_t i0, uint64_t count)
{
    return hvb200cu_gen_hw_mapped(plan, hw, i0, 62, 1, 0, count, 0);
}
// directions [i0, i0 + count) of the lattice into rows [offset, offset + count) of the direction buffer
extern "C" int hvb200cu_gen_hw_at(struct hvb200_plan *plan, const hvb200_hw_params *hw, uint64_t i0, uint64_t count, uint64_t offset)
{
    return hvb200cu_gen_hw_mapped(plan, hw, i0, 62, 1, 0, count, offset);
}
// positions [p0, p0 + count) of the shard (first, shift, stride) into rows [offset, offset + count) of the direction buffer
extern "C" int hvb200cu_gen_hw_mapped(struct hvb200_plan *plan, const hvb200_hw_params *hw, uint64_t first, int shift, uint64_t stride,
                                      uint64_t p0, uint64_t count, uint64_t offset)
{
    HVB200_NVTX("hvb200:gen_hw");
    DevPlan *dp = (DevPlan *)plan->dev;
    if ((offset + count) * (uint64_t)(2 * hw->d) > dp->dirs_cap) { hvb200_set_error("internal: direction buffer overflow"); return -1; }
    if (hw->nsamples > (1ull << 32)) { hvb200_set_error("DZ2019-HW on the device supports nsamples <= 2^32"); return -1; }
    HwDev P;
    P.N = hw->nsamples;
    P.Nrecip = hw->nsamples ? ~0ull / hw->nsamples : 0;
    P.d = hw->d;
    for (int k = 0; k < HVB200_MAXD; k++) P.a[k] = hw->a[k];
    const double *tab = dp->device < 16 ? g_hw_tab[dp->device] : nullptr;
    double *out = dp->dirs + offset * (uint64_t)(2 * hw->d);
    const unsigned grid = (unsigned)((count + 127) / 128);
    const int b = span_begin(dp, 1);
    if (gen_hw_staged(hw->d)) k_gen_hw<true><<<grid, 128, gen_hw_smem_bytes(hw->d), dp->stream>>>(P, first, shift, stride, p0, count, out, dp->counters, tab, g_hw_ntab);
    else k_gen_hw<false><<<grid, 128, 0, dp->stream>>>(P, first, shift, stride, p0, count, out, dp->counters, tab, g_hw_ntab);
    span_end(dp, b);
    CUDA_TRY(cudaGetLastError());
    plan->stats.dirgen_launches++;
    return 0;
}
extern "C" int hvb200cu_gen_mc(struct hvb200_plan *plan, const double *normals_host, uint64_t count)
{
    DevPlan *dp = (DevPlan *)plan->dev;
    if (stage_to_device(plan, dp, normals_host, count * (size_t)dp->d)) return -1;
    const int b = span_begin(dp, 1);
    k_gen_mc<<<(unsigned)((count + 127) / 128), 128, 0, dp->stream>>>(dp->staging_dev, count, dp->d, dp->dirs);
    span_end(dp, b);
    CUDA_TRY(cudaGetLastError());
    plan->stats.dirgen_launches++;
    return 0;
}
// DZ2019-MC with the stream generated on the device
extern "C" int hvb200cu_mc_begin(struct hvb200_plan *plan, uint32_t seed)
{
    DevPlan *dp = (DevPlan *)plan->dev;
    if (!dp->mc) dp->mc = new McGen();
    return mcgen_begin(dp, dp->mc, seed);
}
// the stream from pair `pair0` on (jump-ahead), for shards that own a pair range
extern "C" int hvb200cu_mc_begin_at(struct hvb200_plan *plan, uint32_t seed, uint64_t pair0)
{
    DevPlan *dp = (DevPlan *)plan->dev;
    CUDA_TRY(cudaSetDevice(dp->device));
    if (!dp->mc) dp->mc = new McGen();
    return mcgen_begin(dp, dp->mc, seed, pair0);
}
extern "C" int hvb200cu_mc_scan_segment(struct hvb200_plan *plan, uint32_t seed, uint64_t pair0, uint64_t npairs,
                                        int want_body, uint64_t out[20])
{
    HVB200_NVTX("hvb200:mc_scan_segment");
    DevPlan *dp = (DevPlan *)plan->dev;
    CUDA_TRY(cudaSetDevice(dp->device));
    if (!dp->mc) dp->mc = new McGen();
    dp->stream = dp->own_stream;             // outside a run: the plan's own stream, a fresh list of timing spans
    dp->spans.clear();
    dp->ev_next = 0;
    if (const int rc = mcgen_begin(dp, dp->mc, seed, pair0)) return rc;
    ull tmp[20];
    const int rc = mcgen_scan_segment(plan, dp, dp->mc, npairs, want_body, tmp);
    if (rc == 0) for (int i = 0; i < 20; i++) out[i] = tmp[i];
    return rc;
}
extern "C" int hvb200cu_mc_skip(struct hvb200_plan *plan, uint64_t nnormals)
{
    DevPlan *dp = (DevPlan *)plan->dev;
    while (nnormals) {
        const uint64_t step = nnormals < (1ull << 24) ? nnormals : (1ull << 24);
        const int rc = mcgen_normals(plan, dp, dp->mc, step, nullptr);
        if (rc) return rc;
        nnormals -= step;
    }
    return 0;
}
extern "C" int hvb200cu_gen_mc_device(struct hvb200_plan *plan, uint64_t count)
{
    HVB200_NVTX("hvb200:gen_mc");
    DevPlan *dp = (DevPlan *)plan->dev;
    const size_t ndoubles = count * (size_t)dp->d;
    if (dp->staging_dev_cap < ndoubles) {
        cudaFree(dp->staging_dev);
        dp->staging_dev = nullptr; dp->staging_dev_cap = 0;
        CUDA_TRY(cudaMalloc(&dp->staging_dev, ndoubles * sizeof(double)));
        dp->staging_dev_cap = ndoubles;
    }
    if (const int rc = mcgen_normals(plan, dp, dp->mc, ndoubles, dp->staging_dev)) return rc;
    const int b = span_begin(dp, 1);
    k_gen_mc<<<(unsigned)((count + 127) / 128), 128, 0, dp->stream>>>(dp->staging_dev, count, dp->d, dp->dirs);
    span_end(dp, b);
    CUDA_TRY(cudaGetLastError());
    plan->stats.dirgen_launches++;
    return 0;
}
extern "C" int hvb200cu_gen_rphi(struct hvb200_plan *plan, const double *u_host, uint64_t count)
{
    HVB200_NVTX("hvb200:gen_rphi");
    DevPlan *dp = (DevPlan *)plan->dev;
    if (stage_to_device(plan, dp, u_host, count * (size_t)(dp->d - 1))) return -1;
    const int b = span_begin(dp, 1);
    k_gen_rphi<<<(unsigned)((count + 127) / 128), 128, 0, dp->stream>>>(dp->staging_dev, count, dp->d, dp->dirs);
    span_end(dp, b);
    CUDA_TRY(cudaGetLastError());
    plan->stats.dirgen_launches++;
    return 0;
}
// ---- cached direction sets (Rphi-FWE+ directions depend on d and the sample index only) ----------
extern "C" int hvb200cu_dev_alloc(int device, size_t bytes, void **out)
{
    CUDA_TRY(cudaSetDevice(device));
    CUDA_TRY(cudaMalloc(out, bytes));
    return 0;
}
extern "C" void hvb200cu_dev_free(int device, void *p)
{
    if (cudaSetDevice(device) == cudaSuccess) cudaFree(p);
    cudaGetLastError();
}
// rows [0, count) of the plan's direction buffer -> dst (device), then wait: dst is published to other streams
extern "C" int hvb200cu_copy_dirs_out(struct hvb200_plan *plan, double *dst, uint64_t count, int wait)
{
    DevPlan *dp = (DevPlan *)plan->dev;
    CUDA_TRY(cudaMemcpyAsync(dst, dp->dirs, count * (size_t)(2 * dp->d) * sizeof(double), cudaMemcpyDeviceToDevice, dp->stream));
    if (wait) CUDA_TRY(cudaStreamSynchronize(dp->stream));
    return 0;
}
extern "C" int hvb200cu_copy_dev(struct hvb200_plan *plan, double *dst, const double *src, size_t doubles)
{
    DevPlan *dp = (DevPlan *)plan->dev;
    CUDA_TRY(cudaMemcpyAsync(dst, src, doubles * sizeof(double), cudaMemcpyDeviceToDevice, dp->stream));
    CUDA_TRY(cudaStreamSynchronize(dp->stream));
    return 0;
}
extern "C" void hvb200cu_plan_sync(struct hvb200_plan *plan)
{
    DevPlan *dp = (DevPlan *)plan->dev;
    if (!dp) return;
    cudaSetDevice(dp->device);
    if (dp->stream) cudaStreamSynchronize(dp->stream);
    cudaGetLastError();
}
extern "C" void hvb200cu_set_dirs_view(struct hvb200_plan *plan, const double *view)
{
    ((DevPlan *)plan->dev)->dirs_view = view;
}
extern "C" int hvb200cu_gen_uploaded(struct hvb200_plan *plan, const double *r_host, uint64_t count)
{
    DevPlan *dp = (DevPlan *)plan->dev;
    if (stage_to_device(plan, dp, r_host, count * (size_t)dp->d)) return -1;
    const int b = span_begin(dp, 1);
    k_prep_r<<<(unsigned)((count + 127) / 128), 128, 0, dp->stream>>>(dp->staging_dev, count, dp->d, dp->dirs);
    span_end(dp, b);
    CUDA_TRY(cudaGetLastError());
    plan->stats.dirgen_launches++;
    return 0;
}
extern "C" int hvb200cu_read_dirs(struct hvb200_plan *plan, uint64_t count, double *r_host)
{
    DevPlan *dp = (DevPlan *)plan->dev;
    double *tmp = nullptr;
    CUDA_TRY(cudaMalloc(&tmp, count * (size_t)dp->d * sizeof(double)));
    k_extract_r<<<(unsigned)((count + 127) / 128), 128, 0, dp->stream>>>(dp->dirs_view ? dp->dirs_view : dp->dirs, count, dp->d, tmp);
    cudaError_t e = cudaMemcpyAsync(r_host, tmp, count * (size_t)dp->d * sizeof(double), cudaMemcpyDeviceToHost, dp->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(dp->stream);
    cudaFree(tmp);
    if (e != cudaSuccess) { hvb200_set_error("read_dirs: %s", cudaGetErrorString(e)); return -1; }
    return 0;
}

// Grid of a k_maxmin_bp launch.  Long runs take every resident CTA and let the dynamic chunk hand-out even things out; a
// short run (default-sized calls: 262 144 directions are 512 chunks for 444 resident CTAs) would leave most of the GPU
// idle while a few CTAs work on a second chunk, so it gets ceil(chunks / rounds) CTAs: every CTA the same number of chunks.
static unsigned bp_grid_for(const DevPlan *dp, ull n_chunks)
{
    const ull full = (ull)dp->grid;
    if (n_chunks <= full) return (unsigned)(n_chunks ? n_chunks : 1);
    const ull rounds = (n_chunks + full - 1) / full;
    if (rounds > 3) return (unsigned)full;
    return (unsigned)((n_chunks + rounds - 1) / rounds);
}

// the arguments of k_maxmin_bp that do not depend on where the directions come from
static int bp_fill_args(DevPlan *dp, BpArgs &ka, uint64_t count, ull n_chunks)
{
    const int d = dp->d;
    memset(&ka, 0, sizeof ka);
    ka.pts = dp->bp_pts;
    ka.piv32 = dp->bp_piv32;
    ka.piv_idx = dp->bp_pividx;
    ka.n_piv = dp->bp_npiv;
    ka.pkeys = dp->bp_pkeys;
    ka.bpair = dp->bp_bpair;
    ka.nb = dp->bp_nb;
    ka.nb_pad = dp->bp_nb_pad;
    ka.seg_blocks = dp->bp_seg_blocks;
    ka.count = count;
    const size_t slots = (size_t)n_chunks * (kBpThreads / 32);           // one per (chunk, warp)
    if (dp->bp_chunk_off + slots > dp->bp_chunk_cap) { hvb200_set_error("internal: chunk buffer too small"); return -1; }
    ka.chunk_sums = dp->bp_chunk_sums + dp->bp_chunk_off;
    dp->bp_chunk_off += slots;
    ka.chunk_counter = dp->bp_chunk_counter;
    CUDA_TRY(cudaMemsetAsync(dp->bp_chunk_counter, 0, sizeof(unsigned), dp->stream));
    ka.slow_counter = dp->counters + 1;
    ka.capped_counter = dp->counters;
    ka.hits = dp->bp_hits;
    ka.quads = dp->bp_quads;
    {
        const char *e = getenv("HVB200_BP_EARLYOUT");
        ka.early_out = !(e && e[0] == '0');
    }
    for (int k = 0; k < HVB200_MAXD; k++) {
        ka.qscale[k] = k < d ? dp->bp_qscale[k] : 0.0;
        ka.qlos[k] = k < d ? dp->bp_qlos[k] : 0.0;
    }
    return 0;
}

extern "C" int hvb200cu_run_is_fused(struct hvb200_plan *plan) { return ((DevPlan *)plan->dev)->fused ? 1 : 0; }

// DZ2019-HW sum with the generator fused into the max-min kernel: `count` directions, position p of the launch being
// lattice index first + ((p >> shift) * stride << shift) + (p & (2^shift - 1)).  One launch whatever the count: there
// is no direction buffer to size it by.
extern "C" int hvb200cu_maxmin_hw_fused(struct hvb200_plan *plan, const hvb200_hw_params *hw, uint64_t first, int shift,
                                        uint64_t stride, uint64_t count)
{
    HVB200_NVTX("hvb200:maxmin_hw_fused");
    DevPlan *dp = (DevPlan *)plan->dev;
    const int d = dp->d;
    if (!dp->fused || dp->kernel_used != MODE_BP) { hvb200_set_error("internal: run was not set up for the fused kernel"); return -1; }
    if (count == 0) return 0;
    if (!dp->bp_ready && bp_prepare_finish(dp)) return -1;
    const ull chunk = (ull)kBpThreads * bp_R_of(d);
    const ull n_chunks = (count + chunk - 1) / chunk;
    if (n_chunks > 0xfffffff0ull) { hvb200_set_error("internal: too many chunks for one launch"); return -1; }
    const unsigned grid = bp_grid_for(dp, n_chunks);
    BpArgs ka;
    const int b = span_begin(dp, 0);
    if (bp_fill_args(dp, ka, count, n_chunks)) return -1;
    ka.hw.N = hw->nsamples;
    ka.hw.Nrecip = hw->nsamples ? ~0ull / hw->nsamples : 0;
    ka.hw.d = hw->d;
    for (int k = 0; k < HVB200_MAXD; k++) ka.hw.a[k] = hw->a[k];
    ka.gen_i0 = first;
    ka.gen_stride = stride;
    ka.gen_shift = shift;
    ka.gen_tab = dp->device < 16 ? g_hw_tab[dp->device] : nullptr;
    ka.gen_ntab = g_hw_ntab;
    ka.gen_scratch = dp->gen_scratch;
    bp_fn fn = pick_bp_gen<32>(d);
    fn<<<grid, kBpThreads, dp->bp_smem, dp->stream>>>(ka);
    span_end(dp, b);
    CUDA_TRY(cudaGetLastError());
    plan->stats.maxmin_launches++;
    return 0;
}

extern "C" int hvb200cu_maxmin(struct hvb200_plan *plan, uint64_t count, double *values_host)
{
    HVB200_NVTX("hvb200:maxmin");
    DevPlan *dp = (DevPlan *)plan->dev;
    const int d = dp->d;
    if (values_host && dp->values_cap < count) {
        cudaFree(dp->values);
        dp->values = nullptr; dp->values_cap = 0;
        CUDA_TRY(cudaMalloc(&dp->values, count * sizeof(double)));
        dp->values_cap = count;
    }
    MaxminArgs a;
    a.pts = dp->pts;
    a.n_tiles = dp->n_tiles;
    a.tile_points = dp->tile_points;
    a.stage_bytes = dp->stage_bytes;
    a.staged = dp->staged;
    a.dirs = dp->dirs_view ? dp->dirs_view : dp->dirs;
    a.count = count;
    a.cta_sums = dp->cta_sums;
    a.values = values_host ? dp->values : nullptr;
    a.slow_counter = dp->counters + 1;
    const bool dpx = dp->kernel_used == MODE_DPX, bp = dp->kernel_used == MODE_BP;
    if (bp && !dp->bp_ready && bp_prepare_finish(dp)) return -1;     // host part of the preparation, behind the queued generator
    const size_t smem = dpx ? dp->k_smem_bytes : (bp ? dp->bp_smem : dp->smem_bytes);
    const ull chunk = bp ? (ull)kBpThreads * bp_R_of(d)
                         : (ull)kConsumerThreads * (dpx ? dp->kR : (dp->kernel_used == MODE_EXACT ? dp->R_exact : dp->R));
    const ull n_chunks = (count + chunk - 1) / chunk;
    const unsigned grid = bp ? bp_grid_for(dp, n_chunks) : (unsigned)std::min<ull>((ull)dp->grid, n_chunks ? n_chunks : 1);
    const int b = span_begin(dp, 0);
    if (bp) {
        BpArgs ka;
        if (bp_fill_args(dp, ka, count, n_chunks)) return -1;
        ka.dirs = dp->dirs_view ? dp->dirs_view : dp->dirs;
        ka.values = a.values;
        bp_fn fn = pick_bp<HVB200_MAXD>(d);
        fn<<<grid, kBpThreads, smem, dp->stream>>>(ka);
    } else if (dpx) {
        DpxArgs ka;
        ka.keys = dp->keys;
        ka.n_tiles = dp->k_n_tiles;
        ka.tile_pairs = dp->k_tile_pairs;
        ka.stage_bytes = dp->k_stage_bytes;
        ka.npoints = dp->n;
        ka.n_pivots = (int)std::min<size_t>(dp->n, (size_t)dpx_pivots());
        ka.pts = dp->pts;
        ka.dirs = dp->dirs_view ? dp->dirs_view : dp->dirs;
        ka.count = count;
        ka.cta_sums = dp->cta_sums;
        ka.values = a.values;
        ka.slow_counter = dp->counters + 1;
        for (int k = 0; k < 32; k++) {
            ka.qscale[k] = dp->qscale[k];
            ka.qlos[k] = k < d ? dp->qlo[k] * dp->qscale[k] * (1.0 + 0x1p-20) + 0x1p-10 : 0.0;
        }
        dpx_fn fn = pick_dpx<32>(d);
        fn<<<grid, kBlockThreads, smem, dp->stream>>>(ka);
    } else {
        maxmin_fn fn = pick_kernel(d, dp->kernel_used);
        fn<<<grid, kBlockThreads, smem, dp->stream>>>(a);
    }
    span_end(dp, b);
    CUDA_TRY(cudaGetLastError());
    plan->stats.maxmin_launches++;
    if (values_host) {
        CUDA_TRY(cudaMemcpyAsync(values_host, dp->values, count * sizeof(double), cudaMemcpyDeviceToHost, dp->stream));
        CUDA_TRY(cudaStreamSynchronize(dp->stream));
    }
    return 0;
}

static int max_dyn_smem_limit();
// ---- epsilon indicator (the same max-min semiring, other operands) --------------------
// I_eps(A, B) = max_{b in B} min_{a in A} max_k v_k(a, b),  v_k = a_k - b_k (additive) or a_k / b_k (multiplicative)
// for a minimised objective and the operands swapped for a maximised one (c/epsilon.h:52-107).  Every v_k is one
// IEEE operation and max/min are exact, so the result carries the reference's bits whatever the evaluation order;
// the reference's early exits (:88-99) only prune.  One thread per point of B, A streamed through shared memory.
struct EpsArgs { signed char mm[HVB200_MAXD]; int d; };
static constexpr int kEpsThreads = 256, kEpsTile = 256;
// MODE: -1 all objectives minimised, +1 all maximised, 0 mixed (per-objective sign from e.mm) — the three
// specialisations of the reference (c/epsilon.h:166-176).  D > 0: compile-time dimension, the reference point of
// the thread lives in registers; D = 0: run-time dimension.
template <bool MULT, int MODE, int D>
__global__ void __launch_bounds__(kEpsThreads) k_epsilon(const double *__restrict__ A, ull na, const double *__restrict__ B, ull nb,
                                                         const EpsArgs e, double *__restrict__ block_max)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *sa = reinterpret_cast<double *>(smem_raw);                 // [kEpsTile][d]
    __shared__ double red[kEpsThreads / 32];
    const int d = D ? D : e.d;
    const ull b = (ull)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = b < nb;
    const double *pb = B + (valid ? b : 0) * (ull)d;
    double pbr[D ? D : 1];
    if constexpr (D > 0) {
#pragma unroll
        for (int k = 0; k < D; k++) pbr[k] = __ldg(pb + k);
    }
    double eps_min = __longlong_as_double(0x7ff0000000000000LL);
    for (ull t0 = 0; t0 < na; t0 += kEpsTile) {
        const int tn = (int)min((ull)kEpsTile, na - t0);
        __syncthreads();
        for (int i = threadIdx.x; i < tn * d; i += kEpsThreads) sa[i] = A[t0 * (ull)d + i];
        __syncthreads();
        for (int i = 0; i < tn; i++) {
            const double *pa = sa + (size_t)i * d;
            double emax = 0.0;
#pragma unroll
            for (int k = 0; k < (D ? D : d); k++) {
                const double x = pa[k], y = D ? pbr[D ? k : 0] : __ldg(pb + k);
                const int sgn = MODE ? MODE : (int)e.mm[k];
                double v;
                if (sgn < 0) v = MULT ? x / y : x - y;
                else if (sgn > 0) v = MULT ? y / x : y - x;
                else v = 0.0;
                emax = (k == 0 || v > emax) ? v : emax;
            }
            eps_min = (emax < eps_min) ? emax : eps_min;
        }
    }
    double v = valid ? eps_min : -__longlong_as_double(0x7ff0000000000000LL);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const double o = __shfl_down_sync(0xffffffffu, v, off);
        v = (o > v) ? o : v;
    }
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double m = red[0];
        for (int w = 1; w < kEpsThreads / 32; w++) m = (red[w] > m) ? red[w] : m;
        block_max[blockIdx.x] = m;
    }
}
typedef void (*eps_fn)(const double *, ull, const double *, ull, const EpsArgs, double *);
template <bool MULT, int MODE, int D>
static eps_fn pick_eps_d(int d)
{
    if constexpr (D < 2) {
        return (eps_fn)k_epsilon<MULT, MODE, 0>;
    } else {
        if (d == D) return (eps_fn)k_epsilon<MULT, MODE, D>;
        return pick_eps_d<MULT, MODE, D - 1>(d);
    }
}
static eps_fn pick_eps(bool mult, int mode, int d)
{
    if (mult) return mode < 0 ? pick_eps_d<true, -1, 32>(d) : (mode > 0 ? pick_eps_d<true, 1, 32>(d) : pick_eps_d<true, 0, 32>(d));
    return mode < 0 ? pick_eps_d<false, -1, 32>(d) : (mode > 0 ? pick_eps_d<false, 1, 32>(d) : pick_eps_d<false, 0, 32>(d));
}
extern "C" int hvb200cu_epsilon(int device, int mult, const double *A_host, size_t na, const double *B_host, size_t nb, int d,
                                const signed char *mm, double *out)
{
    CUDA_TRY(cudaSetDevice(device));
    const unsigned blocks = (unsigned)((nb + kEpsThreads - 1) / kEpsThreads);
    double *A = nullptr, *B = nullptr, *bm = nullptr;
    int rc = -1;
    std::vector<double> host(blocks);
    do {
        if (cudaMalloc(&A, na * (size_t)d * sizeof(double)) != cudaSuccess || cudaMalloc(&B, nb * (size_t)d * sizeof(double)) != cudaSuccess ||
            cudaMalloc(&bm, blocks * sizeof(double)) != cudaSuccess) break;
        if (cudaMemcpy(A, A_host, na * (size_t)d * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) break;
        if (cudaMemcpy(B, B_host, nb * (size_t)d * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) break;
        EpsArgs e;
        e.d = d;
        for (int k = 0; k < HVB200_MAXD; k++) e.mm[k] = k < d ? mm[k] : 0;
        const size_t smem = (size_t)kEpsTile * d * sizeof(double);
        int mode = mm[0];
        for (int k = 1; k < d; k++) if (mm[k] != mm[0]) mode = 0;
        eps_fn fn = pick_eps(mult != 0, mode, d);
        if (cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn_smem_limit()) != cudaSuccess) break;
        fn<<<blocks, kEpsThreads, smem>>>(A, (ull)na, B, (ull)nb, e, bm);
        if (cudaMemcpy(host.data(), bm, blocks * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) break;
        rc = 0;
    } while (0);
    if (rc) hvb200_set_error("epsilon indicator: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFree(A); cudaFree(B); cudaFree(bm);
    if (rc) return -1;
    double eps = mult ? 0.0 : -INFINITY;                  // c/epsilon.h:83
    for (unsigned i = 0; i < blocks; i++) eps = (host[i] > eps) ? host[i] : eps;
    *out = eps;
    return 0;
}

// ---- HypE weighted-hypervolume sampling, 2 objectives: the dominator counts of estimate_whv (c/whv_hype.c:152-195) ----
// One thread per sample, the normalised points streamed through shared memory.  A point dominates a sample unless
// sample[d] < p[d] for some d (the reference's test, :172-176, NaN included); the counts are integers, so any order works.
constexpr int kWhvTile = 2048;
__global__ void __launch_bounds__(256) k_whv_counts(const double2 *__restrict__ samples, ull nsamples,
                                                    const double2 *__restrict__ pts, unsigned npoints, unsigned *__restrict__ counts)
{
    __shared__ double2 sp[kWhvTile];
    const ull s = (ull)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = s < nsamples;
    const double2 smp = valid ? samples[s] : make_double2(0.0, 0.0);
    unsigned cnt = 0;
    for (unsigned t0 = 0; t0 < npoints; t0 += kWhvTile) {
        const unsigned tn = min((unsigned)kWhvTile, npoints - t0);
        __syncthreads();
        for (unsigned i = threadIdx.x; i < tn; i += blockDim.x) sp[i] = pts[t0 + i];
        __syncthreads();
#pragma unroll 4
        for (unsigned i = 0; i < tn; i++) {
            const double2 p = sp[i];
            cnt += (!(smp.x < p.x) && !(smp.y < p.y)) ? 1u : 0u;
        }
    }
    if (valid) counts[s] = cnt;
}
extern "C" int hvb200cu_whv_counts(int device, const double *pts_host, size_t npoints, const double *samples_host, size_t nsamples,
                                   unsigned *counts_host)
{
    CUDA_TRY(cudaSetDevice(device));
    double2 *P = nullptr, *S = nullptr;
    unsigned *C = nullptr;
    int rc = -1;
    do {
        if (cudaMalloc(&P, npoints * sizeof(double2)) != cudaSuccess || cudaMalloc(&S, nsamples * sizeof(double2)) != cudaSuccess ||
            cudaMalloc(&C, nsamples * sizeof(unsigned)) != cudaSuccess) break;
        if (cudaMemcpy(P, pts_host, npoints * sizeof(double2), cudaMemcpyHostToDevice) != cudaSuccess) break;
        if (cudaMemcpy(S, samples_host, nsamples * sizeof(double2), cudaMemcpyHostToDevice) != cudaSuccess) break;
        k_whv_counts<<<(unsigned)((nsamples + 255) / 256), 256>>>(S, (ull)nsamples, P, (unsigned)npoints, C);
        if (cudaMemcpy(counts_host, C, nsamples * sizeof(unsigned), cudaMemcpyDeviceToHost) != cudaSuccess) break;
        rc = 0;
    } while (0);
    if (rc) hvb200_set_error("whv_hype dominator counts: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFree(P); cudaFree(S); cudaFree(C);
    return rc;
}

// ---- generational distances: GD / IGD / IGD+ / averaged Hausdorff (c/igd.h:60-289) -------------------------------
// gd_common_helper_ (c/igd.h:61-127): for every point a of X the squared distance to the nearest point of Y,
//     min_y sum_k diff_k^2,  diff_k = a_k - y_k   (plain)   or   max(oriented difference, 0)   ("plus"),
// then a square root / power and a left-to-right sum over a.  The O(|X| |Y| d) part runs here, one thread per a with
// Y streamed through shared memory; products and sums are separate roundings in the reference's order (-fmad=false),
// min is exact, so each per-point value carries the reference's bits.  The tail (sqrt, pow_uint, ordered sum, powl)
// is replayed on the host in the reference's order, which makes the result bit-identical, not just close.
template <bool PLUS, int MODE, int D>
__global__ void __launch_bounds__(kEpsThreads) k_gd(const double *__restrict__ X, ull nx, const double *__restrict__ Y, ull ny,
                                                    const EpsArgs e, double *__restrict__ out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *sy = reinterpret_cast<double *>(smem_raw);                 // [kEpsTile][d]
    const int d = D ? D : e.d;
    const ull a = (ull)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = a < nx;
    const double *pa = X + (valid ? a : 0) * (ull)d;
    double par[D ? D : 1];
    if constexpr (D > 0) {
#pragma unroll
        for (int k = 0; k < D; k++) par[k] = __ldg(pa + k);
    }
    double min_dist = __longlong_as_double(0x7ff0000000000000LL);
    for (ull t0 = 0; t0 < ny; t0 += kEpsTile) {
        const int tn = (int)min((ull)kEpsTile, ny - t0);
        __syncthreads();
        for (int i = threadIdx.x; i < tn * d; i += kEpsThreads) sy[i] = Y[t0 * (ull)d + i];
        __syncthreads();
        for (int i = 0; i < tn; i++) {
            const double *pr = sy + (size_t)i * d;
            double dist = 0.0;
#pragma unroll
            for (int k = 0; k < (D ? D : d); k++) {
                const double a_d = D ? par[D ? k : 0] : __ldg(pa + k), r_d = pr[k];
                const int sgn = MODE ? MODE : (int)e.mm[k];
                double diff;
                if (PLUS) {
                    const double v = sgn < 0 ? r_d - a_d : (sgn > 0 ? a_d - r_d : 0.0);
                    diff = (v < 0.0) ? 0.0 : v;                        // MAX(v, 0.), c/maxminclamp.h:58
                } else {
                    diff = sgn != 0 ? a_d - r_d : 0.0;
                }
                dist += diff * diff;
            }
            min_dist = (min_dist > dist) ? dist : min_dist;
        }
    }
    if (valid) out[a] = min_dist;
}
typedef void (*gd_fn)(const double *, ull, const double *, ull, const EpsArgs, double *);
template <bool PLUS, int MODE, int D>
static gd_fn pick_gd_d(int d)
{
    if constexpr (D < 2) {
        return (gd_fn)k_gd<PLUS, MODE, 0>;
    } else {
        if (d == D) return (gd_fn)k_gd<PLUS, MODE, D>;
        return pick_gd_d<PLUS, MODE, D - 1>(d);
    }
}
static gd_fn pick_gd(bool plus, int mode, int d)
{
    if (plus) return mode < 0 ? pick_gd_d<true, -1, 16>(d) : (mode > 0 ? pick_gd_d<true, 1, 16>(d) : pick_gd_d<true, 0, 16>(d));
    // the plain difference is squared: the orientation does not matter, only ignored objectives (mm = 0) do
    return mode != 0 ? pick_gd_d<false, -1, 16>(d) : pick_gd_d<false, 0, 16>(d);
}
// pow_uint (c/pow_int.h:59-74) on the host, same multiplication trees as the device pow_chain
static double pow_uint_host(double x, unsigned e)
{
    if (e == 0) return 1.0;
    if (e <= 31) {
        const PowProg P = kPowProg[e];
        double v[8];
        v[0] = x;
        for (int t = 0; t < P.n; t++) v[t + 1] = v[P.op[t][0]] * v[P.op[t][1]];
        return v[P.n];
    }
    double acc = (e & 1) ? x : 1.0;
    for (e >>= 1; e; e >>= 1) { x *= x; if (e & 1u) acc *= x; }
    return acc;
}
// gd_common (c/igd.h:61-127, 159-174) for (X, Y) = (points_a, points_r)
extern "C" int hvb200cu_gd_common(int device, const double *X_host, size_t nx, const double *Y_host, size_t ny, int d,
                                  const signed char *mm, int plus, int psize, unsigned p, double *result)
{
    if (nx == 0) { *result = INFINITY; return 0; }
    CUDA_TRY(cudaSetDevice(device));
    const unsigned blocks = (unsigned)((nx + kEpsThreads - 1) / kEpsThreads);
    double *X = nullptr, *Y = nullptr, *out = nullptr;
    std::vector<double> host(nx);
    int rc = -1;
    do {
        if (cudaMalloc(&X, nx * (size_t)d * sizeof(double)) != cudaSuccess || cudaMalloc(&Y, (ny ? ny : 1) * (size_t)d * sizeof(double)) != cudaSuccess ||
            cudaMalloc(&out, nx * sizeof(double)) != cudaSuccess) break;
        if (cudaMemcpy(X, X_host, nx * (size_t)d * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) break;
        if (ny && cudaMemcpy(Y, Y_host, ny * (size_t)d * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) break;
        EpsArgs e;
        e.d = d;
        for (int k = 0; k < HVB200_MAXD; k++) e.mm[k] = k < d ? mm[k] : 0;
        int mode = mm[0];
        for (int k = 1; k < d; k++) if (mm[k] != mm[0]) mode = 0;
        gd_fn fn = pick_gd(plus != 0, mode, d);
        if (cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn_smem_limit()) != cudaSuccess) break;
        fn<<<blocks, kEpsThreads, (size_t)kEpsTile * d * sizeof(double)>>>(X, (ull)nx, Y, (ull)ny, e, out);
        if (cudaMemcpy(host.data(), out, nx * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) break;
        rc = 0;
    } while (0);
    if (rc) hvb200_set_error("generational distance: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFree(X); cudaFree(Y); cudaFree(out);
    if (rc) return -1;
    double gd = 0.0;
    for (size_t a = 0; a < nx; a++) {
        double md = host[a];
        if (md == 0.0) continue;                                        // zero_skip, c/igd.h:100-101,115
        if (p == 1) md = sqrt(md);
        else if (p % 2 == 0) md = pow_uint_host(md, p / 2);
        else md = pow_uint_host(sqrt(md), p);
        gd += md;
    }
    if (p == 1) *result = gd / (double)nx;
    else if (psize) *result = (double)powl(gd / (double)nx, 1.0 / p);
    else *result = (double)powl(gd, 1.0 / p) / (double)nx;
    return 0;
}

// ---- batched point sets ------------------------------------------------------------------------
struct SetsCtx {
    int kind = 0;                        // first member of every per-batch sink: 0 = batched sets, 1 = contributions
    double *pts = nullptr;
    unsigned *off = nullptr;
    double *partial = nullptr, *hilo = nullptr;
    int nsets = 0, grid_x = 0, d = 0;
    size_t smem = 0;
};
typedef void (*sets_fn)(const double *, const unsigned *, const double *, ull, int, double *);
template <int D>
static sets_fn pick_sets(int d)
{
    if constexpr (D < 2) {
        return (sets_fn)k_maxmin_sets<0>;
    } else {
        if (d == D) return (sets_fn)k_maxmin_sets<D>;
        return pick_sets<D - 1>(d);
    }
}
extern "C" size_t hvb200cu_sets_max_points(int d)
{
    return (size_t)(200 * 1024) / ((size_t)d * sizeof(double));     // one set must fit the shared memory of a CTA
}
extern "C" int hvb200cu_sets_begin(struct hvb200_plan *plan, const double *pts_host, const unsigned *off_host, int nsets, void **ctx_out)
{
    DevPlan *dp = (DevPlan *)plan->dev;
    CUDA_TRY(cudaSetDevice(dp->device));
    SetsCtx *c = new SetsCtx();
    *ctx_out = c;
    c->nsets = nsets;
    c->d = dp->d;
    const size_t total = off_host[nsets];
    unsigned nmax = 0;
    for (int i = 0; i < nsets; i++) nmax = std::max(nmax, off_host[i + 1] - off_host[i]);
    c->smem = (size_t)nmax * dp->d * sizeof(double) + 16;
    // about 4 CTAs per SM over all sets, at least one per set
    c->grid_x = std::max(1, (dp->sm_count * 4 + nsets - 1) / nsets);
    CUDA_TRY(cudaMalloc(&c->pts, std::max<size_t>(1, total) * dp->d * sizeof(double)));
    CUDA_TRY(cudaMalloc(&c->off, (size_t)(nsets + 1) * sizeof(unsigned)));
    CUDA_TRY(cudaMalloc(&c->partial, (size_t)nsets * c->grid_x * sizeof(double)));
    CUDA_TRY(cudaMalloc(&c->hilo, (size_t)nsets * 2 * sizeof(double)));
    CUDA_TRY(cudaMemcpy(c->pts, pts_host, total * dp->d * sizeof(double), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(c->off, off_host, (size_t)(nsets + 1) * sizeof(unsigned), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemset(c->partial, 0, (size_t)nsets * c->grid_x * sizeof(double)));
    sets_fn fn = pick_sets<16>(dp->d);
    CUDA_TRY(cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn_smem_limit()));
    return 0;
}
static int contrib_batch(struct hvb200_plan *plan, void *ctx, uint64_t count);
extern "C" int hvb200cu_sets_maxmin(struct hvb200_plan *plan, void *ctx, uint64_t count)
{
    if (*(const int *)ctx == 1) return contrib_batch(plan, ctx, count);
    DevPlan *dp = (DevPlan *)plan->dev;
    SetsCtx *c = (SetsCtx *)ctx;
    sets_fn fn = pick_sets<16>(dp->d);
    const int b = span_begin(dp, 0);
    // always grid_x CTAs per set (idle ones add 0): the partial slots keep one stride over all launches of a run
    fn<<<dim3((unsigned)c->grid_x, (unsigned)c->nsets), kSetsThreads, c->smem, dp->stream>>>(c->pts, c->off, dp->dirs_view ? dp->dirs_view : dp->dirs, (ull)count, dp->d, c->partial);
    span_end(dp, b);
    CUDA_TRY(cudaGetLastError());
    plan->stats.maxmin_launches++;
    return 0;
}
extern "C" int hvb200cu_sets_end(struct hvb200_plan *plan, void *ctx, double *hilo_host)
{
    DevPlan *dp = (DevPlan *)plan->dev;
    SetsCtx *c = (SetsCtx *)ctx;
    int rc = 0;
    if (hilo_host) {
        k_reduce_sets<<<c->nsets, 256, 0, dp->stream>>>(c->partial, c->grid_x, c->hilo);
        if (cudaMemcpyAsync(hilo_host, c->hilo, (size_t)c->nsets * 2 * sizeof(double), cudaMemcpyDeviceToHost, dp->stream) != cudaSuccess ||
            cudaStreamSynchronize(dp->stream) != cudaSuccess) {
            hvb200_set_error("batched sets: %s", cudaGetErrorString(cudaGetLastError()));
            rc = -1;
        }
    }
    cudaFree(c->pts); cudaFree(c->off); cudaFree(c->partial); cudaFree(c->hilo);
    delete c;
    return rc;
}

// ------------------------------------------------------------------------------------------------
// hypervolume-contribution approximation (hvb200_contrib.cuh): per batch of directions k_top2, a stable sort of
// (arg-max, delta) and one warp per point adding its run — deterministic whatever the schedule
// ------------------------------------------------------------------------------------------------
struct ContribCtx {
    int kind = 1;
    double *contrib = nullptr;           // [n] running sums of s1^d - s2^d
    double *delta[2] = {nullptr, nullptr};
    unsigned *arg[2] = {nullptr, nullptr};
    void *cub_tmp = nullptr;
    size_t cap = 0, cub_bytes = 0;
};
typedef void (*top2_fn)(const double *, int, int, int, const double *, ull, double *, unsigned *, double *);
template <int D>
static top2_fn pick_top2(int d)
{
    if constexpr (D < 2) {
        return nullptr;
    } else {
        if (d == D) return (top2_fn)k_top2<D>;
        return pick_top2<D - 1>(d);
    }
}
extern "C" int hvb200cu_contrib_begin(struct hvb200_plan *plan, void **ctx_out)
{
    DevPlan *dp = (DevPlan *)plan->dev;
    *ctx_out = nullptr;
    if (dp->d > 32) { hvb200_set_error("contribution approximation supports at most 32 objectives"); return -2; }
    ContribCtx *c = new (std::nothrow) ContribCtx();
    if (!c) { hvb200_set_error("out of host memory"); return -4; }
    *ctx_out = c;
    CUDA_TRY(cudaSetDevice(dp->device));
    CUDA_TRY(cudaMalloc(&c->contrib, dp->n * sizeof(double)));
    CUDA_TRY(cudaMemsetAsync(c->contrib, 0, dp->n * sizeof(double), dp->own_stream));
    CUDA_TRY(cudaStreamSynchronize(dp->own_stream));
    return 0;
}
static int contrib_batch(struct hvb200_plan *plan, void *ctx, uint64_t count)
{
    HVB200_NVTX("hvb200:contributions");
    DevPlan *dp = (DevPlan *)plan->dev;
    ContribCtx *c = (ContribCtx *)ctx;
    const int d = dp->d;
    if (count == 0) return 0;
    if (count > 0x7fffffffull) { hvb200_set_error("internal: contribution batch too large"); return -1; }
    if (c->cap < count) {
        for (int i = 0; i < 2; i++) { cudaFree(c->delta[i]); cudaFree(c->arg[i]); c->delta[i] = nullptr; c->arg[i] = nullptr; }
        c->cap = 0;
        for (int i = 0; i < 2; i++) {
            CUDA_TRY(cudaMalloc(&c->delta[i], count * sizeof(double)));
            CUDA_TRY(cudaMalloc(&c->arg[i], count * sizeof(unsigned)));
        }
        c->cap = count;
    }
    int key_bits = 1;
    while (key_bits < 32 && (1ull << key_bits) <= dp->n) key_bits++;
    size_t need = 0;
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, need, (const unsigned *)nullptr, (unsigned *)nullptr, (const double *)nullptr, (double *)nullptr,
                                             (int)count, 0, key_bits, dp->stream));
    if (c->cub_bytes < need) {
        cudaFree(c->cub_tmp); c->cub_tmp = nullptr; c->cub_bytes = 0;
        CUDA_TRY(cudaMalloc(&c->cub_tmp, need));
        c->cub_bytes = need;
    }
    top2_fn fn = pick_top2<32>(d);
    const size_t smem = 2 * (size_t)dp->stage_bytes + 2 * sizeof(ull) + 8 * sizeof(double);
    CUDA_TRY(cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn_smem_limit()));
    const ull chunk = (ull)kTop2Threads * (d <= 12 ? 2 : 1);
    const ull n_chunks = (count + chunk - 1) / chunk;
    const unsigned grid = (unsigned)std::min<ull>((ull)std::min(dp->grid, 2 * dp->sm_count), n_chunks);
    const int b = span_begin(dp, 0);
    fn<<<grid, kTop2Threads, smem, dp->stream>>>(dp->pts, dp->n_tiles, dp->tile_points, dp->stage_bytes, dp->dirs_view ? dp->dirs_view : dp->dirs,
                                                  (ull)count, c->delta[0], c->arg[0], dp->cta_sums);
    span_end(dp, b);
    CUDA_TRY(cudaGetLastError());
    size_t tmp = c->cub_bytes;
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(c->cub_tmp, tmp, (const unsigned *)c->arg[0], c->arg[1], (const double *)c->delta[0], c->delta[1],
                                             (int)count, 0, key_bits, dp->stream));
    k_contrib_accumulate<<<(unsigned)(((ull)dp->n * 32 + 255) / 256), 256, 0, dp->stream>>>(c->arg[1], c->delta[1], (ull)count, (unsigned)dp->n, c->contrib);
    CUDA_TRY(cudaGetLastError());
    plan->stats.maxmin_launches++;
    return 0;
}
extern "C" int hvb200cu_contrib_end(struct hvb200_plan *plan, void *ctx, double *contrib_host)
{
    DevPlan *dp = (DevPlan *)plan->dev;
    ContribCtx *c = (ContribCtx *)ctx;
    int rc = 0;
    if (c && contrib_host && c->contrib) {
        if (cudaMemcpyAsync(contrib_host, c->contrib, dp->n * sizeof(double), cudaMemcpyDeviceToHost, dp->stream) != cudaSuccess ||
            cudaStreamSynchronize(dp->stream) != cudaSuccess) {
            hvb200_set_error("reading the contributions back failed: %s", cudaGetErrorString(cudaGetLastError()));
            rc = -1;
        }
    }
    if (c) {
        cudaFree(c->contrib);
        for (int i = 0; i < 2; i++) { cudaFree(c->delta[i]); cudaFree(c->arg[i]); }
        cudaFree(c->cub_tmp);
        delete c;
    }
    return rc;
}

extern "C" int hvb200cu_run_end(struct hvb200_plan *plan, double out_hilo[2])
{
    HVB200_NVTX("hvb200:reduce");
    DevPlan *dp = (DevPlan *)plan->dev;
    {
        const int b = span_begin(dp, 2);
        if (dp->kernel_used == MODE_BP && dp->bp_chunk_off > (size_t)2 * kReduceSlice) {
            const int nsl = (int)((dp->bp_chunk_off + kReduceSlice - 1) / kReduceSlice);
            if (dp->reduce_pairs_cap < (size_t)nsl) {
                cudaFree(dp->reduce_pairs);
                dp->reduce_pairs = nullptr; dp->reduce_pairs_cap = 0;
                CUDA_TRY(cudaMalloc(&dp->reduce_pairs, (size_t)nsl * 2 * sizeof(double)));
                dp->reduce_pairs_cap = (size_t)nsl;
            }
            k_reduce_slices<<<nsl, 1024, 0, dp->stream>>>(dp->bp_chunk_sums, (ull)dp->bp_chunk_off, dp->reduce_pairs);
            k_reduce_pairs<<<1, 1024, 0, dp->stream>>>(dp->reduce_pairs, nsl, dp->out_hilo);
        } else if (dp->kernel_used == MODE_BP) k_reduce<<<1, 1024, 0, dp->stream>>>(dp->bp_chunk_sums, (int)dp->bp_chunk_off, dp->out_hilo);
        else k_reduce<<<1, 1024, 0, dp->stream>>>(dp->cta_sums, dp->grid, dp->out_hilo);
        span_end(dp, b);
        CUDA_TRY(cudaGetLastError());
        plan->stats.reduce_launches++;
    }
    double host[2];
    ull counters[8];
    CUDA_TRY(cudaMemcpyAsync(host, dp->out_hilo, sizeof host, cudaMemcpyDeviceToHost, dp->stream));
    CUDA_TRY(cudaMemcpyAsync(counters, dp->counters, sizeof counters, cudaMemcpyDeviceToHost, dp->stream));
    CUDA_TRY(cudaStreamSynchronize(dp->stream));
    if (out_hilo) { out_hilo[0] = host[0]; out_hilo[1] = host[1]; }
    plan->stats.newton_capped = counters[0];
    plan->stats.slow_path = counters[1];
    if (getenv("HVB200_DEBUG_COUNTERS"))
        fprintf(stderr, "hvb200 counters: newton capped=%llu exact evaluations of candidates=%llu\n", counters[0], counters[1]);
    plan->stats.npoints = dp->n;
    plan->stats.kernel = dp->kernel_used == MODE_EXACT ? HVB200_KERNEL_EXACT : (dp->kernel_used == MODE_DPX ? HVB200_KERNEL_DPX : (dp->kernel_used == MODE_BP ? HVB200_KERNEL_BP : HVB200_KERNEL_SCREEN));
    for (const auto &s : dp->spans) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, dp->ev_pool[s.a], dp->ev_pool[s.b]) != cudaSuccess) { cudaGetLastError(); continue; }
        if (s.kind == 0) plan->stats.maxmin_ms += ms;
        else if (s.kind == 1) plan->stats.dirgen_ms += ms;
        else if (s.kind == 2) plan->stats.reduce_ms += ms;
        else plan->stats.h2d_ms += ms;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Single-process multi-GPU exchange step: one NCCL all-gather of the per-GPU (hi, lo) pairs over
// NVLink.  libnccl is opened lazily with dlopen (no link-time dependency: a process that already
// loaded PyTorch's bundled NCCL gets that copy; without NCCL the caller combines the host copies).
// ------------------------------------------------------------------------------------------------
#include <dlfcn.h>
namespace {
typedef void *nccl_comm_t;
struct NcclApi {
    int (*CommInitAll)(nccl_comm_t *, int, const int *) = nullptr;
    int (*CommDestroy)(nccl_comm_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool ok = false, tried = false;
};
NcclApi g_nccl;
std::mutex g_nccl_mutex;
int g_nccl_ndev = 0;
nccl_comm_t g_nccl_comms[16];
double *g_nccl_gather[16] = {nullptr};

bool nccl_load()
{
    if (g_nccl.tried) return g_nccl.ok;
    g_nccl.tried = true;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return false;
    g_nccl.CommInitAll = (int (*)(nccl_comm_t *, int, const int *))dlsym(h, "ncclCommInitAll");
    g_nccl.CommDestroy = (int (*)(nccl_comm_t))dlsym(h, "ncclCommDestroy");
    g_nccl.AllGather = (int (*)(const void *, void *, size_t, int, nccl_comm_t, cudaStream_t))dlsym(h, "ncclAllGather");
    g_nccl.GroupStart = (int (*)())dlsym(h, "ncclGroupStart");
    g_nccl.GroupEnd = (int (*)())dlsym(h, "ncclGroupEnd");
    g_nccl.GetErrorString = (const char *(*)(int))dlsym(h, "ncclGetErrorString");
    g_nccl.ok = g_nccl.CommInitAll && g_nccl.AllGather && g_nccl.GroupStart && g_nccl.GroupEnd;
    return g_nccl.ok;
}
}  // namespace

// All-gather the (hi, lo) results of `n` plans (plan g on device g, results still in device memory
// after hvb200cu_run_end) and return the 2n doubles of device 0's copy.  Returns 1 if NCCL is not
// available (caller falls back to the host copies), -1 on error, 0 on success.
extern "C" int hvb200cu_nccl_allgather(struct hvb200_plan **plans, int n, double *pairs_out)
{
    HVB200_NVTX("hvb200:nccl_allgather");
    std::lock_guard<std::mutex> lock(g_nccl_mutex);
    if (n < 2 || n > 16 || !nccl_load()) return 1;
    if (g_nccl_ndev != n) {
        if (g_nccl_ndev) {
            for (int g = 0; g < g_nccl_ndev; g++) {
                if (g_nccl.CommDestroy) g_nccl.CommDestroy(g_nccl_comms[g]);
                cudaSetDevice(g); cudaFree(g_nccl_gather[g]); g_nccl_gather[g] = nullptr;
            }
            g_nccl_ndev = 0;
        }
        int devs[16];
        for (int g = 0; g < n; g++) devs[g] = g;
        const int rc = g_nccl.CommInitAll(g_nccl_comms, n, devs);
        if (rc != 0) { hvb200_set_error("ncclCommInitAll: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error"); return -1; }
        for (int g = 0; g < n; g++) {
            CUDA_TRY(cudaSetDevice(g));
            CUDA_TRY(cudaMalloc(&g_nccl_gather[g], 2 * 16 * sizeof(double)));
        }
        g_nccl_ndev = n;
    }
    g_nccl.GroupStart();
    for (int g = 0; g < n; g++) {
        DevPlan *dp = (DevPlan *)plans[g]->dev;
        const int rc = g_nccl.AllGather(dp->out_hilo, g_nccl_gather[g], 2, 8 /* ncclFloat64 */, g_nccl_comms[g], dp->own_stream);
        if (rc != 0) { g_nccl.GroupEnd(); hvb200_set_error("ncclAllGather: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error"); return -1; }
    }
    const int rc = g_nccl.GroupEnd();
    if (rc != 0) { hvb200_set_error("ncclGroupEnd: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error"); return -1; }
    for (int g = 0; g < n; g++) {
        DevPlan *dp = (DevPlan *)plans[g]->dev;
        CUDA_TRY(cudaSetDevice(g));
        CUDA_TRY(cudaStreamSynchronize(dp->own_stream));
    }
    CUDA_TRY(cudaSetDevice(0));
    CUDA_TRY(cudaMemcpy(pairs_out, g_nccl_gather[0], 2 * n * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" double hvb200cu_measure_fp64(int device, double *dmul_per_s, double *dsetp_per_s)
{
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return -1.0; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -1.0;
    double *out = nullptr;
    if (cudaMalloc(&out, 8) != cudaSuccess) return -1.0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4096, grid = prop.multiProcessorCount * 8, threads = 256;
    const double ops = (double)grid * threads * iters * 64.0;
    double res[3] = {0, 0, 0};
    for (int op = 0; op < 3; op++) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; rep++) {
            cudaEventRecord(e0);
            if (op == 0) k_fp64_bench<0><<<grid, threads>>>(out, iters, 1.0);
            else if (op == 1) k_fp64_bench<1><<<grid, threads>>>(out, iters, 1.0);
            else k_fp64_bench<2><<<grid, threads>>>(out, iters, 1.0);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < best) best = ms;
        }
        res[op] = ops / (best * 1e-3);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(out);
    if (cudaGetLastError() != cudaSuccess) return -1.0;
    if (dmul_per_s) *dmul_per_s = res[1];
    if (dsetp_per_s) *dsetp_per_s = res[2];
    return 2.0 * res[0];
}
