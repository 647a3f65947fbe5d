// hvb200_contrib.cuh — hypervolume-CONTRIBUTION approximation on the direction set of hv_approx (included by
// hvb200_kernels.cu).  SURVEY §8(f) rank 3, second half: the DZ2019 paper behind c/hvapprox.h:20-24 approximates
// contributions as well; the reference implements only the exact ones (c/hv_contrib.c), so this is an extension whose
// parity is UNPINNED — it is cross-checked against the exact hv_contributions at low d and against a brute-force
// restatement on shared directions.
//
// Estimator.  HV(P) ~ c_d/N sum_i s1_i^d with s1_i = max_p min_k p'_k r^(i)_k.  Removing point a changes the term of
// direction i only if a is the arg-max there, and then to s2_i^d (the second largest value).  Hence
//     HVC(a) = HV(P) - HV(P \ {a}) ~ c_d/N  sum_{i : argmax_i = a} (s1_i^d - s2_i^d).
// Duplicates of the arg-max give s2 = s1 (contribution 0, like the exact algorithm); dominated points must stay in the
// set (they can be the second best), so the plan is built without the nondominated pre-filter.
//
// k_top2<D>   the textbook loop with a predicate chain against the SECOND best value (a point matters only if every
//             product exceeds s2): per element one DMUL + one DSETP.GT.AND, the min is formed for the few points that
//             pass.  Point tiles arrive by TMA bulk copies into a 2-stage ring; R = 2 directions per thread up to d = 12.
//             Output per direction: delta = s1^d - s2^d (the reference's pow chains) and the arg-max; sum of s1^d per CTA.
// accumulate  stable radix sort of (arg-max, delta) per batch, then one warp per point adds its run of deltas with a
//             fixed lane-strided tree: the result does not depend on scheduling.
#pragma once

static constexpr int kTop2Threads = 256;
template <int D> struct Top2Cfg { static constexpr int R = D <= 12 ? 2 : 1, G = 2; };

template <int D>
__global__ void __launch_bounds__(kTop2Threads, 2) k_top2(const double *__restrict__ pts, int n_tiles, int tile_points, int stage_bytes,
                                                          const double *__restrict__ dirs, ull count, double *__restrict__ delta,
                                                          unsigned *__restrict__ arg, double *__restrict__ cta_sums)
{
    constexpr int R = Top2Cfg<D>::R, G = Top2Cfg<D>::G;
    static_assert((G * D) % 2 == 0, "a group of points is a whole number of 16-byte words");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *tiles = reinterpret_cast<double *>(smem_raw);                                   // [2][stage_bytes]
    ull *full_bar = reinterpret_cast<ull *>(smem_raw + 2 * (size_t)stage_bytes);           // [2]
    double *red = reinterpret_cast<double *>(full_bar + 2);                                 // [8]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int stage_doubles = stage_bytes / 8;
    const unsigned tile_bytes = (unsigned)tile_points * D * 8u;
    if (tid == 0) { mbar_init(&full_bar[0], 1); mbar_init(&full_bar[1], 1); mbar_fence_init(); }
    __syncthreads();
    constexpr ull CH = (ull)kTop2Threads * R;
    const ull n_chunks = (count + CH - 1) / CH;
    double acc = 0.0;
    ull q = 0;                                       // tiles consumed so far by this CTA (ring position)
    for (ull chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
        double r[R][D], s1[R], s2[R];
        unsigned am[R];
        ull sample[R];
#pragma unroll
        for (int rr = 0; rr < R; rr++) {
            sample[rr] = chunk * CH + (ull)rr * kTop2Threads + tid;
            const bool valid = sample[rr] < count;
#pragma unroll
            for (int k = 0; k < D; k++) r[rr][k] = valid ? __ldg(dirs + sample[rr] * (2 * D) + k) : 0.0;
            s1[rr] = s2[rr] = 0.0;
            am[rr] = 0xffffffffu;
        }
        if (tid == 0) {
            // first two tiles of this chunk (the previous chunk's readers are behind the barrier that ended it)
            for (int t = 0; t < min(2, n_tiles); t++) {
                const int st = (int)((q + t) & 1);
                mbar_arrive_expect_tx(&full_bar[st], tile_bytes);
                tma_load_1d(tiles + (size_t)st * stage_doubles, pts + (size_t)t * tile_points * D, tile_bytes, &full_bar[st]);
            }
        }
        for (int t = 0; t < n_tiles; t++, q++) {
            const int st = (int)(q & 1);
            mbar_wait(&full_bar[st], (unsigned)((q >> 1) & 1));
            const double *tile = tiles + (size_t)st * stage_doubles;
            const double2 *tile2 = reinterpret_cast<const double2 *>(tile);
#pragma unroll 1
            for (int g0 = 0; g0 < tile_points; g0 += G) {
                const double2 *gp = tile2 + (g0 * D) / 2;
                bool c[R][G];
#pragma unroll
                for (int rr = 0; rr < R; rr++)
#pragma unroll
                    for (int g = 0; g < G; g++) c[rr][g] = true;
#pragma unroll
                for (int j = 0; j < (G * D) / 2; j++) {
                    const double2 v = gp[j];
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const int e = 2 * j + h, g = e / D, k = e % D;
                        const double pv = h ? v.y : v.x;
#pragma unroll
                        for (int rr = 0; rr < R; rr++) c[rr][g] = c[rr][g] & (pv * r[rr][k] > s2[rr]);
                    }
                }
                bool any = false;
#pragma unroll
                for (int rr = 0; rr < R; rr++)
#pragma unroll
                    for (int g = 0; g < G; g++) any = any | c[rr][g];
                if (any) {
#pragma unroll
                    for (int g = 0; g < G; g++)
#pragma unroll
                        for (int rr = 0; rr < R; rr++)
                            if (c[rr][g]) {
                                const double *p = tile + (size_t)(g0 + g) * D;
                                double m = p[0] * r[rr][0];
#pragma unroll
                                for (int k = 1; k < D; k++) {
                                    const double tt = p[k] * r[rr][k];
                                    m = (tt < m) ? tt : m;
                                }
                                if (m > s1[rr]) { s2[rr] = s1[rr]; s1[rr] = m; am[rr] = (unsigned)(t * tile_points + g0 + g); }
                                else if (m > s2[rr]) s2[rr] = m;
                            }
                }
            }
            __syncthreads();                         // every thread is done with this stage
            if (tid == 0 && t + 2 < n_tiles) {
                mbar_arrive_expect_tx(&full_bar[st], tile_bytes);
                tma_load_1d(tiles + (size_t)st * stage_doubles, pts + (size_t)(t + 2) * tile_points * D, tile_bytes, &full_bar[st]);
            }
        }
#pragma unroll
        for (int rr = 0; rr < R; rr++) {
            if (sample[rr] < count) {
                const double v1 = pow_chain<D>(s1[rr]), v2 = pow_chain<D>(s2[rr]);
                acc += v1;
                delta[sample[rr]] = v1 - v2;
                arg[sample[rr]] = am[rr];
            }
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, off);
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    if (tid == 0) cta_sums[blockIdx.x] += ((red[0] + red[1]) + (red[2] + red[3])) + ((red[4] + red[5]) + (red[6] + red[7]));
}

// contrib[p] += sum of the deltas whose (sorted) arg-max is p: one warp per point, lanes stride over its run in
// direction order, fixed shuffle tree
__global__ void __launch_bounds__(256) k_contrib_accumulate(const unsigned *__restrict__ arg_sorted, const double *__restrict__ delta_sorted,
                                                            ull count, unsigned npoints, double *__restrict__ contrib)
{
    const unsigned p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (p >= npoints) return;
    // [lo, hi) = run of keys equal to p (lower bounds of p and p + 1)
    auto lower = [&](unsigned key) {
        ull a = 0, b = count;
        while (a < b) {
            const ull mid = (a + b) >> 1;
            if (__ldg(arg_sorted + mid) < key) a = mid + 1; else b = mid;
        }
        return a;
    };
    const ull lo = lower(p), hi = lower(p + 1);
    double s = 0.0;
    for (ull i = lo + lane; i < hi; i += 32) s += __ldg(delta_sorted + i);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
    if (lane == 0 && hi > lo) contrib[p] += s;
}
