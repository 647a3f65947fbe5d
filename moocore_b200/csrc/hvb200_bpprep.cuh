// hvb200_bpprep.cuh — the block-pruned kernel's plan-time preparation on the device (included by hvb200_kernels.cu).
//
// What k_maxmin_bp needs besides the points (hvb200_bp.cuh): the points in kd-tree order (32 consecutive points = one
// block with a tight bounding box), their 16-bit keys in two layouts, the per-block key maxima, and the pivots.  Round 1
// built all of it on the host (1.6 - 2.3 ms at n = 1e4, d = 10; 122 ms at n = 1e5, d = 50) with the GPU idle; here it is
// a short sequence of kernels on the plan's own stream, and the first max-min launch waits for it with an event.
//
//   kd order     recursive median splits on the widest normalised objective, level by level.  The tree's SHAPE depends
//                on n only (a node of cnt > 32 points splits into whole blocks: left = ceil(cnt/2 / 32) * 32), so the host
//                uploads the segment boundaries of every level and the device only decides the CONTENT: per level
//                  k_kd_minmax   per (segment, objective) min / max of the raw coordinates (ordered bit patterns, atomics)
//                  k_kd_keys     key = segment << vb | fixed-point position of the point inside the segment's range of
//                                its widest objective (vb = 16 - segment bits, at least 12); leaves keep their order
//                  cub radix sort of (key, index): inside every segment the points end up ordered by that objective,
//                                so its left part is the lower "half" — a median split without a selection pass
//                Once every segment holds at most 1024 points, k_kd_local finishes the tree in ONE launch (a CTA per
//                segment, bitonic sorts in shared memory).
//                The order may differ from the host build's where coordinates agree to 2^-vb of the segment's range or two
//                objectives are equally wide; any order gives the same sample values (max over points is exact and
//                order-free) — the order only decides how tight the block boxes are.
//   k_bp_gather  points -> kd order (zero padding rows), inverse permutation
//   k_bp_keys    point keys [block][W][32] (two objectives per word)
//   k_bp_bounds  block maxima [block pair][d] (two blocks per word)
//   pivots       radix sort of ((nprobe - wins) << 32 | index): the most frequent probe winners first (ties: lower
//                index, as on the host), k_bp_pivots writes their FP32 copies and their rows in the kd order
#pragma once
#include <cub/device/device_radix_sort.cuh>

static constexpr int kKdKeyBits = 16;       // segment id + fixed-point position: two 8-bit radix passes per level
static constexpr int kKdValueBitsMin = 12;

__global__ void __launch_bounds__(256) k_iota_u32(unsigned *__restrict__ a, unsigned n)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = i;
}
__global__ void __launch_bounds__(256) k_fill_u64(ull *__restrict__ a, ull n, ull even, ull odd)
{
    const ull i = (ull)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = (i & 1ull) ? odd : even;
}
// segment of position i at this level: starts[s] <= i < starts[s + 1]
__device__ __forceinline__ unsigned kd_segment(const unsigned *__restrict__ starts, unsigned nseg, unsigned i)
{
    unsigned lo = 0, hi = nseg;                     // invariant: starts[lo] <= i < starts[hi]
    while (hi - lo > 1) {
        const unsigned mid = (lo + hi) >> 1;
        if (__ldg(starts + mid) <= i) lo = mid; else hi = mid;
    }
    return lo;
}
// minmax[(s * d + k) * 2 + {0, 1}] = min / max bit pattern of objective k over segment s (positive doubles order like
// their bit patterns).  One thread per point; a warp whose lanes share the segment reduces with shuffles first.
__global__ void __launch_bounds__(256) k_kd_minmax(const double *__restrict__ pts, const unsigned *__restrict__ idx, unsigned n, int d,
                                                   const unsigned *__restrict__ starts, unsigned nseg, ull *__restrict__ minmax)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < n;
    const unsigned seg = live ? kd_segment(starts, nseg, i) : 0xffffffffu;
    const bool split = live && (__ldg(starts + seg + 1) - __ldg(starts + seg) > (unsigned)kBpBlock);
    const unsigned seg0 = __shfl_sync(0xffffffffu, seg, 0);
    const bool uniform = __all_sync(0xffffffffu, seg == seg0 && split);
    if (!__any_sync(0xffffffffu, split)) return;
    const double *row = pts + (size_t)(live ? idx[i] : 0) * d;
    for (int k = 0; k < d; k++) {
        ull lo = split ? (ull)__double_as_longlong(__ldg(row + k)) : ~0ull, hi = split ? lo : 0ull;
        if (uniform) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const ull ol = __shfl_xor_sync(0xffffffffu, lo, off), oh = __shfl_xor_sync(0xffffffffu, hi, off);
                lo = ol < lo ? ol : lo;
                hi = oh > hi ? oh : hi;
            }
            if ((threadIdx.x & 31) == 0) {
                atomicMin(minmax + ((size_t)seg * d + k) * 2, lo);
                atomicMax(minmax + ((size_t)seg * d + k) * 2 + 1, hi);
            }
        } else if (split) {
            atomicMin(minmax + ((size_t)seg * d + k) * 2, lo);
            atomicMax(minmax + ((size_t)seg * d + k) * 2 + 1, hi);
        }
    }
}
__global__ void __launch_bounds__(256) k_kd_keys(const double *__restrict__ pts, const unsigned *__restrict__ idx, unsigned n, int d,
                                                 const unsigned *__restrict__ starts, unsigned nseg, const ull *__restrict__ minmax,
                                                 const double *__restrict__ qhi, int vb, unsigned *__restrict__ keys)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned seg = kd_segment(starts, nseg, i);
    unsigned fixed = 0;
    if (__ldg(starts + seg + 1) - __ldg(starts + seg) > (unsigned)kBpBlock) {
        // widest normalised objective of the segment (the first one among equals, like the host build)
        int bk = 0;
        double bw = -1.0, bmn = 0.0, bq = 1.0;
        for (int k = 0; k < d; k++) {
            const double q = __ldg(qhi + k);
            const double mn = q > 0 ? __longlong_as_double((long long)__ldg(minmax + ((size_t)seg * d + k) * 2)) / q : 0.0;
            const double mx = q > 0 ? __longlong_as_double((long long)__ldg(minmax + ((size_t)seg * d + k) * 2 + 1)) / q : 0.0;
            if (mx - mn > bw) { bw = mx - mn; bk = k; bmn = mn; bq = q; }
        }
        const double v = bq > 0 ? __ldg(pts + (size_t)idx[i] * d + bk) / bq : 0.0;
        const double f = bw > 0 ? (v - bmn) / bw : 0.0;
        const double top = (double)((1u << vb) - 1u);
        const double t = f * (double)(1u << vb);
        fixed = (unsigned)(t < 0.0 ? 0.0 : (t > top ? top : t));
    }
    keys[i] = (seg << vb) | fixed;
}
// ------------------------------------------------------------------------------------------------
// The bottom of the kd tree in one launch: a segment of at most kKdLocalMax points is finished by one CTA in shared
// memory — per local level every warp takes sub-segments (widest objective by shuffles, 32-bit fixed-point keys), then
// one CTA-wide bitonic sort of {sub-segment | key} orders every sub-segment at once; same split rule as the global
// levels (left part = whole 32-point blocks).  Replaces 5 global levels (~8 launches each) at n = 1e4.
// ------------------------------------------------------------------------------------------------
static constexpr int kKdLocalMax = 1024, kKdLocalThreads = 1024;
__global__ void __launch_bounds__(kKdLocalThreads) k_kd_local(const double *__restrict__ pts, unsigned *__restrict__ idx, int d,
                                                              const unsigned *__restrict__ starts, const double *__restrict__ qhi)
{
    constexpr int NW = kKdLocalThreads / 32, MAXSUB = kKdLocalMax / kBpBlock;      // 32 warps, at most 32 sub-segments
    __shared__ ull skey[kKdLocalMax];
    __shared__ unsigned sidx[kKdLocalMax];
    __shared__ unsigned sb[2][MAXSUB + 2];
    __shared__ unsigned nsub[2];
    __shared__ double swid[MAXSUB * HVB200_MAXD];       // normalised width of (sub-segment, objective)
    __shared__ double smin[MAXSUB * HVB200_MAXD];       // its normalised minimum
    const unsigned lo = starts[blockIdx.x], cnt = starts[blockIdx.x + 1] - lo;
    if (cnt <= (unsigned)kBpBlock) return;
    unsigned P = 64;
    while (P < cnt) P <<= 1;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (unsigned p = tid; p < cnt; p += kKdLocalThreads) sidx[p] = idx[lo + p];
    if (tid == 0) { sb[0][0] = 0; sb[0][1] = cnt; nsub[0] = 1; }
    int cur = 0;
    for (;;) {
        __syncthreads();
        const unsigned ns = nsub[cur];
        bool any = false;
        for (unsigned s2 = 0; s2 < ns; s2++) any |= sb[cur][s2 + 1] - sb[cur][s2] > (unsigned)kBpBlock;
        if (!any) break;
        // ---- (sub-segment, objective) ranges: one warp per pair
        for (unsigned t = warp; t < ns * (unsigned)d; t += NW) {
            const unsigned s2 = t / d;
            const int k = (int)(t - s2 * d);
            const unsigned b0 = sb[cur][s2], b1 = sb[cur][s2 + 1];
            if (b1 - b0 <= (unsigned)kBpBlock) continue;
            ull mn = ~0ull, mx = 0ull;
            for (unsigned p = b0 + lane; p < b1; p += 32) {
                const ull v = (ull)__double_as_longlong(__ldg(pts + (size_t)sidx[p] * d + k));
                mn = v < mn ? v : mn;
                mx = v > mx ? v : mx;
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const ull ol = __shfl_xor_sync(0xffffffffu, mn, off), oh = __shfl_xor_sync(0xffffffffu, mx, off);
                mn = ol < mn ? ol : mn;
                mx = oh > mx ? oh : mx;
            }
            if (lane == 0) {
                const double q = __ldg(qhi + k);
                const double a = q > 0 ? __longlong_as_double((long long)mn) / q : 0.0, b = q > 0 ? __longlong_as_double((long long)mx) / q : 0.0;
                swid[t] = b - a;
                smin[t] = a;
            }
        }
        __syncthreads();
        // ---- keys: one warp per sub-segment (the first widest objective, like the host build)
        for (unsigned s2 = warp; s2 < ns; s2 += NW) {
            const unsigned b0 = sb[cur][s2], b1 = sb[cur][s2 + 1];
            if (b1 - b0 <= (unsigned)kBpBlock) {
                for (unsigned p = b0 + lane; p < b1; p += 32) skey[p] = ((ull)s2 << 32) | (ull)(p - b0);      // a leaf keeps its order
                continue;
            }
            int bk = 0;
            double bw = -1.0;
            for (int k = 0; k < d; k++) { const double w = swid[s2 * d + k]; if (w > bw) { bw = w; bk = k; } }
            const double bmn = smin[s2 * d + bk], bq = __ldg(qhi + bk);
            for (unsigned p = b0 + lane; p < b1; p += 32) {
                const double v = bq > 0 ? __ldg(pts + (size_t)sidx[p] * d + bk) / bq : 0.0;
                const double t = (bw > 0 ? (v - bmn) / bw : 0.0) * 4294967296.0;
                const unsigned fixed = t <= 0.0 ? 0u : (t >= 4294967295.0 ? 0xffffffffu : (unsigned)t);
                skey[p] = ((ull)s2 << 32) | fixed;
            }
        }
        for (unsigned p = cnt + tid; p < P; p += kKdLocalThreads) { skey[p] = ~0ull; sidx[p] = 0xffffffffu; }
        __syncthreads();
        // ---- bitonic sort of (skey, sidx) over P slots: thread t owns the t-th compare-exchange of a step
        for (unsigned k2 = 2; k2 <= P; k2 <<= 1)
            for (unsigned j = k2 >> 1; j > 0; j >>= 1) {
                if ((unsigned)tid < P / 2) {
                    const unsigned i = 2 * j * ((unsigned)tid / j) + ((unsigned)tid % j), x = i + j;      // i has bit j clear
                    const ull a = skey[i], b = skey[x];
                    if (((i & k2) == 0) ? (a > b) : (a < b)) {
                        skey[i] = b; skey[x] = a;
                        const unsigned t = sidx[i]; sidx[i] = sidx[x]; sidx[x] = t;
                    }
                }
                __syncthreads();
            }
        // ---- next level's sub-segments
        if (tid == 0) {
            unsigned m = 0;
            for (unsigned s2 = 0; s2 < ns; s2++) {
                const unsigned b0 = sb[cur][s2], c = sb[cur][s2 + 1] - b0;
                sb[cur ^ 1][m++] = b0;
                if (c > (unsigned)kBpBlock) {
                    unsigned half = (c / 2 + kBpBlock - 1) / kBpBlock * kBpBlock;
                    if (half >= c) half = c / 2;
                    sb[cur ^ 1][m++] = b0 + half;
                }
            }
            sb[cur ^ 1][m] = cnt;
            nsub[cur ^ 1] = m;
        }
        cur ^= 1;
    }
    for (unsigned p = tid; p < cnt; p += kKdLocalThreads) idx[lo + p] = sidx[p];
}
__global__ void __launch_bounds__(256) k_bp_gather(const double *__restrict__ pts, const unsigned *__restrict__ idx, ull n, ull n_pad, int d,
                                                   double *__restrict__ sorted, unsigned *__restrict__ pos)
{
    const ull w = (ull)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_pad * (ull)d) return;
    const ull i = w / d;
    const int k = (int)(w - i * d);
    if (i < n) {
        const unsigned o = idx[i];
        sorted[w] = pts[(ull)o * d + k];
        if (k == 0) pos[o] = (unsigned)i;
    } else {
        sorted[w] = 0.0;
    }
}
// KP = clamp(ceil((p - lo) * scale), 0, S); padding points -1; the dummy objective of an odd d = S (always passes)
__device__ __forceinline__ unsigned bp_point_key(const double *__restrict__ sorted, ull i, ull n, int k, int d,
                                                 const double *__restrict__ qlo, const double *__restrict__ qscale)
{
    if (i >= n) return 0xffffu;
    if (k >= d) return (unsigned)kKeyScale;
    double f = ceil((sorted[i * d + k] - __ldg(qlo + k)) * __ldg(qscale + k));
    f = f < 0.0 ? 0.0 : (f > (double)kKeyScale ? (double)kKeyScale : f);
    return (unsigned)(int)f;
}
__global__ void __launch_bounds__(256) k_bp_keys(const double *__restrict__ sorted, ull n, int nb_pad, int d, const double *__restrict__ qlo,
                                                 const double *__restrict__ qscale, unsigned *__restrict__ pkeys)
{
    const int W = (d + 1) / 2;
    const ull w = (ull)blockIdx.x * blockDim.x + threadIdx.x;          // ((b * W) + j) * 32 + l
    if (w >= (ull)nb_pad * W * kBpBlock) return;
    const int l = (int)(w % kBpBlock);
    const ull bj = w / kBpBlock;
    const int j = (int)(bj % W);
    const ull b = bj / W;
    const ull i = b * kBpBlock + l;
    pkeys[w] = bp_point_key(sorted, i, n, 2 * j, d, qlo, qscale) | (bp_point_key(sorted, i, n, 2 * j + 1, d, qlo, qscale) << 16);
}
// block maxima from the point keys: one warp per (block pair, key word) — lane = point, one max-reduction per half-word
// and block (signed 16-bit: the padding key -1 never raises a maximum)
__global__ void __launch_bounds__(256) k_bp_bounds(const unsigned *__restrict__ pkeys, int nb_pad, int d, unsigned *__restrict__ bpair)
{
    const int W = (d + 1) / 2;
    const unsigned wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (wid >= (unsigned)(nb_pad / 2) * W) return;
    const unsigned pr = wid / W, j = wid - pr * W;
    int m[2][2];
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const unsigned w = __ldg(pkeys + ((size_t)(2 * pr + h) * W + j) * kBpBlock + lane);
        m[h][0] = __reduce_max_sync(0xffffffffu, (int)(short)(w & 0xffffu));
        m[h][1] = __reduce_max_sync(0xffffffffu, (int)(short)(w >> 16));
    }
    if (lane < 2 && 2 * (int)j + (int)lane < d)
        bpair[(size_t)pr * d + 2 * j + lane] = ((unsigned)m[0][lane] & 0xffffu) | (((unsigned)m[1][lane] & 0xffffu) << 16);
}
// pivot ranking keys: (nprobe - wins) in the high bits, the point index in the low IB bits (32-bit keys when the index
// fits 19 bits: 4 radix passes instead of 6)
template <typename K, int IB>
__global__ void __launch_bounds__(256) k_bp_pivot_keys(const unsigned *__restrict__ wins, unsigned n, unsigned nprobe, K *__restrict__ keys)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = ((K)(nprobe - min(wins[i], nprobe)) << IB) | (K)i;
}
template <typename K, int IB>
__global__ void __launch_bounds__(256) k_bp_pivots(const double *__restrict__ pts, const K *__restrict__ keys, const unsigned *__restrict__ pos,
                                                   int npiv, int d, float *__restrict__ piv32, unsigned *__restrict__ pividx)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= npiv * d) return;
    const int j = w / d, k = w - j * d;
    const unsigned o = (unsigned)(keys[j] & (((K)1 << IB) - 1));
    piv32[w] = (float)pts[(size_t)o * d + k];
    if (k == 0) pividx[j] = pos[o];
}

// ------------------------------------------------------------------------------------------------
// Probe directions in FP32, tiled.  The probes only rank points (which ones win most directions -> pivots), so
// neither FP64 nor exact ties matter.  One thread = one probe direction with its reciprocal weights in registers;
// the CTA streams a slice of the points through shared memory as floats (broadcast reads); grid = (probe groups,
// point slices).  Each (probe, slice) leaves one packed key {FP32 bits of the best min | ~index}: positive floats order
// like their bit patterns, and ~index makes the lower index win a tie.  k_probe32_wins reduces the slices.
// Round 1's kernel (one warp per probe, FP64, a dependent chain per point) took 0.45 ms at n = 1e4, d = 10 and
// 39 ms at n = 1e5, d = 50 — more than the rest of the preparation together.
// ------------------------------------------------------------------------------------------------
static constexpr int kProbeThreads = 256, kProbeTile = 64;
template <int DP>
__global__ void __launch_bounds__(kProbeThreads) k_probe32(const double *__restrict__ pts, unsigned n, int d, unsigned nprobe,
                                                           unsigned slice, ull *__restrict__ partial)
{
    __shared__ float tile[kProbeTile * DP];
    const unsigned j = blockIdx.x * kProbeThreads + threadIdx.x;          // probe direction
    float r[DP];
#pragma unroll
    for (int k = 0; k < DP; k++) {
        // |N(0,1)| by Box-Muller on two hashed uniforms; only the direction matters, not its norm
        const unsigned h1 = hash_u32(j * 2654435761u + 2u * k + 1u), h2 = hash_u32(h1 ^ (0x9e3779b9u + k));
        const float u1 = (h1 >> 8) * (1.0f / 16777216.0f) + (0.5f / 16777216.0f), u2 = (h2 >> 8) * (1.0f / 16777216.0f);
        const float z = sqrtf(-2.0f * __logf(u1)) * __cosf(6.2831853f * u2);
        r[k] = k < d ? 1.0f / (fabsf(z) + 1e-6f) : 0.0f;
    }
    const unsigned p0 = blockIdx.y * slice, p1 = min(n, p0 + slice);
    ull best = 0ull;
    for (unsigned t0 = p0; t0 < p1; t0 += kProbeTile) {
        const unsigned cnt = min((unsigned)kProbeTile, p1 - t0);
        __syncthreads();
        for (unsigned w = threadIdx.x; w < (unsigned)kProbeTile * DP; w += kProbeThreads) {
            const unsigned pi = w / DP, k = w - pi * DP;
            // objectives beyond d: +inf, the multiplicative identity of min (their weight is 0, see the clamp below)
            tile[w] = (pi < cnt && k < (unsigned)d) ? (float)__ldg(pts + (size_t)(t0 + pi) * d + k) : __int_as_float(0x7f800000);
        }
        __syncthreads();
        for (unsigned pi = 0; pi < cnt; pi++) {
            const float *pp = tile + pi * DP;
            float m = __int_as_float(0x7f800000);
#pragma unroll
            for (int k = 0; k + 1 < DP; k += 2) {
                // padded objectives give inf * 0 = NaN, which the 3-input min ignores
                m = bp_fmin3(m, pp[k] * r[k], pp[k + 1] * r[k + 1]);
            }
            if (DP & 1) m = fminf(m, pp[DP - 1] * r[DP - 1]);
            const ull key = ((ull)__float_as_uint(m) << 32) | (ull)(~(t0 + pi));
            best = key > best ? key : best;
        }
    }
    if (j < nprobe) partial[(size_t)blockIdx.y * nprobe + j] = best;
}
__global__ void __launch_bounds__(256) k_probe32_wins(const ull *__restrict__ partial, unsigned nprobe, unsigned nslices, unsigned n,
                                                      unsigned *__restrict__ wins)
{
    const unsigned j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nprobe) return;
    ull best = 0ull;
    for (unsigned s = 0; s < nslices; s++) {
        const ull v = partial[(size_t)s * nprobe + j];
        best = v > best ? v : best;
    }
    const unsigned arg = ~(unsigned)(best & 0xffffffffull);
    if (best != 0ull && arg < n) atomicAdd(&wins[arg], 1u);
}

// scratch of one device build, kept with the workspace
struct BpPrepScratch {
    unsigned *idx[2] = {nullptr, nullptr}, *keys[2] = {nullptr, nullptr}, *pos = nullptr, *starts = nullptr;
    ull *minmax = nullptr, *pkeys64[2] = {nullptr, nullptr}, *probe_partial = nullptr;
    size_t cap_probe = 0;
    double *q = nullptr;                 // [3][HVB200_MAXD]: qlo, qhi, qscale
    void *cub_tmp = nullptr;
    size_t cap_n = 0, cap_starts = 0, cap_minmax = 0, cub_bytes = 0;
    cudaEvent_t done = nullptr;
};
static void bp_prep_scratch_free(BpPrepScratch *sc)
{
    for (int i = 0; i < 2; i++) { cudaFree(sc->idx[i]); cudaFree(sc->keys[i]); cudaFree(sc->pkeys64[i]); }
    cudaFree(sc->pos); cudaFree(sc->starts); cudaFree(sc->minmax); cudaFree(sc->q); cudaFree(sc->cub_tmp); cudaFree(sc->probe_partial);
    if (sc->done) cudaEventDestroy(sc->done);
    *sc = BpPrepScratch();
}
