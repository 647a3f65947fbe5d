// hvb200_bp.cuh — block-pruned max-min kernel (included by hvb200_kernels.cu).
//
// Generalises the reference's scalar early-out (c/hvapprox.c:104-108) into a two-level search that
// keeps the reference's bits: max_p min_k p_k r_k only needs the points that can still raise the
// running max, and whole blocks of points can be discarded by their per-objective maxima.
//
//   order     points sorted by their weakest normalised objective (blocks of 32 consecutive points
//             then share small coordinates in the same objective -> tight block bounds); the 64 points
//             that win most probe directions (k_probe_argmax) form a separate warm-up set
//   keys      the DPX 16-bit keys of k_maxmin_dpx (same monotone map, same conservative rounding):
//               pkeys [point][W]   two objectives per word            (lane = point)
//               bkeys [block][W]   per-objective maximum of the block (lane = block)
//               bpair [block pair][D]  two blocks per word            (lane = direction)
//   phase A   lane = direction: exact max-min over the warm-up set -> best, threshold keys
//   phase B   lane = direction: brute-force DPX screen of the BLOCK BOUNDS (n/32 virtual points) ->
//             per-direction bitmask of blocks that may still hold an improving point
//   phase C   warp = direction: the warp walks the direction's surviving blocks; 32 lanes re-check 32
//             block bounds at a time against the current thresholds, then test the 32 points of each
//             live block; candidates are evaluated exactly in FP64 by the warp (lane = objective,
//             shuffle min), and the thresholds are refreshed when the max rises
//
// Every discarded point/block fails a conservative key test against thresholds derived from a value
// that is <= the final max, so the final max — and s^d — is bit-identical to the brute-force kernels.
#pragma once

static constexpr int kBpBlock = 32;           // points per block = lanes per warp
static constexpr int kBpThreads = 256;        // 8 warps, all of them consumers
static constexpr int kBpR = 2;                // directions per thread in phases A and B
static constexpr int kBpG = 4;                // block pairs per group in phase B ((G*D) % 4 == 0)
static constexpr int kBpPivots = 64;

struct BpArgs {
    const double *pts;        // [n_pad][D] FP64, block order
    const double *pivots;     // [n_piv][D] FP64 warm-up set
    int n_piv;
    const unsigned *pkeys;    // [n_pad][W]
    const unsigned *bkeys;    // [nb][W]
    const unsigned *bpair;    // [nb_pad/2][D]
    int nb;                   // blocks
    int nb_pad;               // multiple of 2*kBpG
    int seg_blocks;           // blocks per segment (multiple of 32 and of 2*kBpG)
    int resident;             // 1: all block bounds stay in shared memory (loaded once per CTA, no CTA barriers later)
    const double *dirs;       // [count][2][D]
    ull count;
    double *cta_sums;
    double *values;
    ull *slow_counter;
    double qscale[32], qlos[32];
};

// key of one objective in "negated" form, single half-word (not duplicated)
__device__ __forceinline__ unsigned bp_neg_key16(double best_m, double winv, double scale, double lo_s)
{
    const double x = fma(best_m * winv, scale, -lo_s);
    int kt = __double2int_rd(x);
    kt = max(0, min(kt, kKeyScale + 1));
    return (unsigned)(-kt) & 0xffffu;
}

template <int D>
struct BpW { static constexpr int value = (D + 1) / 2; };

// mask bit -> block index within its 32-block word (phase B writes bit 4h+g of byte `grp` for block 8*grp + 2g + h)
__device__ __forceinline__ int bp_block_of_bit(int bit)
{
    return (bit & 24) | ((bit & 3) << 1) | ((bit >> 2) & 1);
}

// Thresholds of one direction in point packing (two objectives per word), computed by the whole warp:
// lane j < W builds word j, then every lane collects all W words.
template <int D>
__device__ __forceinline__ void bp_point_thresholds(double best, const double *__restrict__ winv, const BpArgs &a,
                                                    int lane, unsigned (&pT)[BpW<D>::value])
{
    constexpr int W = BpW<D>::value;
    const double best_m = best * (1.0 - 0x1p-20);
    unsigned mine = 0;
    if (lane < W) {
        const int k0 = 2 * lane, k1 = 2 * lane + 1;
        const unsigned lo = bp_neg_key16(best_m, __ldg(winv + k0), a.qscale[k0 < 32 ? k0 : 0], a.qlos[k0 < 32 ? k0 : 0]);
        unsigned hi = 0;                                       // dummy objective of an odd D: threshold key 0
        if (k1 < D) hi = bp_neg_key16(best_m, __ldg(winv + k1), a.qscale[k1 < 32 ? k1 : 0], a.qlos[k1 < 32 ? k1 : 0]);
        mine = lo | (hi << 16);
    }
#pragma unroll
    for (int j = 0; j < W; j++) pT[j] = __shfl_sync(0xffffffffu, mine, j);
}

template <int D>
__global__ void __launch_bounds__(kBpThreads, 3) k_maxmin_bp(const __grid_constant__ BpArgs a)
{
    constexpr int R = kBpR, G = kBpG, W = BpW<D>::value;
    static_assert((G * D) % 4 == 0, "group of block pairs must be whole 16-byte words");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr ull CH = (ull)kBpThreads * R;
    const int bnd_pairs = a.resident ? a.nb_pad / 2 : a.seg_blocks / 2;
    const int seg_words = a.seg_blocks / 32;                       // mask words per direction and segment
    unsigned *bnd = reinterpret_cast<unsigned *>(smem_raw);        // [bnd_pairs][D]
    unsigned *mask = bnd + (size_t)bnd_pairs * D;                  // [CH][seg_words]
    double *pivots = reinterpret_cast<double *>(mask + (size_t)CH * seg_words);
    ull *bar = reinterpret_cast<ull *>(pivots + (size_t)kBpPivots * D);
    double *red = reinterpret_cast<double *>(bar + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    for (int i = tid; i < a.n_piv * D; i += kBpThreads) pivots[i] = a.pivots[i];
    __syncthreads();
    unsigned phase = 0;                                            // parity of the bounds barrier
    if (a.resident) {
        // all block bounds fit: one TMA bulk copy per CTA, after which the warps never meet again
        if (tid == 0) {
            const unsigned bytes = (unsigned)(a.nb_pad / 2) * D * 4u;
            mbar_arrive_expect_tx(bar, bytes);
            tma_load_1d(bnd, a.bpair, bytes, bar);
        }
        mbar_wait(bar, phase);
        phase ^= 1u;
    }

    const ull n_chunks = (a.count + CH - 1) / CH;
    const int n_seg = (a.nb_pad + a.seg_blocks - 1) / a.seg_blocks;
    double acc = 0.0;
    ull slow = 0, c_surv = 0, c_rounds = 0, c_live = 0, c_dirs = 0;

    for (ull chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
        ull sample[R];
        bool valid[R];
        double best[R];
        unsigned negT[R][D];                                       // phase-B packing: key duplicated in both halves
#pragma unroll
        for (int r = 0; r < R; r++) {
            sample[r] = chunk * CH + (ull)r * kBpThreads + tid;
            valid[r] = sample[r] < a.count;
            best[r] = 0.0;
            if (valid[r]) {
                // ---- phase A: exact warm-up over the pivots
                const double *dir = a.dirs + sample[r] * (2 * D);
                double rr[D];
#pragma unroll
                for (int k = 0; k < D; k++) rr[k] = __ldg(dir + k);
                double b = 0.0;
#pragma unroll 2
                for (int i = 0; i < a.n_piv; i++) {
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
            }
        }

        for (int seg = 0; seg < n_seg; seg++) {
            const int seg0 = seg * a.seg_blocks;                                   // first block of the segment
            const int seg_nb = min(a.seg_blocks, a.nb_pad - seg0);                 // blocks in it (multiple of 2G)
            // ---- stage the segment's block bounds (pair layout) with one TMA bulk copy
            if (!a.resident) {
                __syncthreads();                                                   // previous users of bnd are done
                if (tid == 0) {
                    const unsigned bytes = (unsigned)(seg_nb / 2) * D * 4u;
                    mbar_arrive_expect_tx(bar, bytes);
                    tma_load_1d(bnd, a.bpair + (size_t)(seg0 / 2) * D, bytes, bar);
                }
            }
            // thresholds for phase B from the current best
#pragma unroll
            for (int r = 0; r < R; r++) {
                if (valid[r]) {
                    const double *winv = a.dirs + sample[r] * (2 * D) + D;
                    const double best_m = best[r] * (1.0 - 0x1p-20);
#pragma unroll
                    for (int k = 0; k < D; k++)
                        negT[r][k] = bp_neg_key16(best_m, __ldg(winv + k), a.qscale[k], a.qlos[k]) * 0x10001u;
                } else {
                    const unsigned v = ((unsigned)(-(kKeyScale + 1)) & 0xffffu) * 0x10001u;      // nothing passes
#pragma unroll
                    for (int k = 0; k < D; k++) negT[r][k] = v;
                }
            }
            if (!a.resident) {
                mbar_wait(bar, phase);
                phase ^= 1u;
            }
            __syncwarp();                 // the warp's mask rows of the previous segment/chunk are no longer read

            // ---- phase B: screen the block bounds, lane = direction
            unsigned *row[R];
#pragma unroll
            for (int r = 0; r < R; r++) row[r] = mask + (size_t)(r * kBpThreads + tid) * seg_words;
            const uint4 *bnd4 = reinterpret_cast<const uint4 *>(bnd + (a.resident ? (size_t)(seg0 / 2) * D : 0));
            unsigned wacc[R];
#pragma unroll
            for (int r = 0; r < R; r++) wacc[r] = 0;
            const int n_groups = seg_nb / (2 * G);
#pragma unroll 1
            for (int gi = 0; gi < n_groups; gi++) {
                const uint4 *gp = bnd4 + (size_t)gi * (G * D / 4);
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
                // 2G = 8 blocks per group -> 8 mask bits; 4 groups fill one 32-bit mask word.  Bit layout of
                // a group's byte: bit (4h + g) <-> block 2g + h of the group (see bp_block_of_bit)
                const int shift = (gi & 3) * (2 * G);
#pragma unroll
                for (int r = 0; r < R; r++) {
                    unsigned neg = 0;                                  // sign bits: low halves -> bits 0..3, high halves -> bits 16..19
#pragma unroll
                    for (int g = 0; g < G; g++) neg |= ((am[r][g] & 0x80008000u) >> 15) << g;
                    const unsigned byte = ((neg | (neg >> 12)) & 0xffu) ^ 0xffu;      // live = sign clear
                    wacc[r] |= byte << shift;
                    if ((gi & 3) == 3 || gi == n_groups - 1) { row[r][gi >> 2] = wacc[r]; wacc[r] = 0; }
                }
            }
            __syncwarp();                 // a warp's mask rows are written and read by that warp only

            // ---- phase C: warp = direction
            const int words_used = (seg_nb + 31) / 32;
#pragma unroll
            for (int r = 0; r < R; r++) {
#pragma unroll 1
                for (int owner = 0; owner < 32; owner++) {
                    const ull smp = chunk * CH + (ull)r * kBpThreads + (ull)(warp * 32 + owner);
                    if (smp >= a.count) continue;                                  // uniform in the warp
                    const unsigned *mrow = mask + (size_t)(r * kBpThreads + warp * 32 + owner) * seg_words;
                    // any surviving block at all?  (most directions are settled by the warm-up)
                    unsigned any = 0;
                    for (int wi = lane; wi < words_used; wi += 32) any |= mrow[wi];
                    if (__ballot_sync(0xffffffffu, any != 0) == 0) continue;
                    c_dirs++;
                    double bcur = __shfl_sync(0xffffffffu, best[r], owner);
                    const double *dir = a.dirs + smp * (2 * D);
                    const double rlane = (lane < D) ? __ldg(dir + lane) : 0.0;     // lane k keeps r_k for the exact evaluations
                    unsigned pT[W];
                    bp_point_thresholds<D>(bcur, dir + D, a, lane, pT);
#pragma unroll 1
                    for (int wi = 0; wi < words_used; wi++) {
                        const unsigned word = mrow[wi];
                        if (word == 0) continue;
                        c_surv += __popc(word); c_rounds++;
                        // re-check 32 block bounds in parallel against the current thresholds (lane = mask bit)
                        bool live = false;
                        const int wblk = seg0 + wi * 32;                       // first block of this mask word
                        const int blk = wblk + bp_block_of_bit(lane);
                        if (((word >> lane) & 1u) && blk < a.nb) {
                            const unsigned *bk = a.bkeys + (unsigned)blk * W;
                            unsigned m16 = 0x7fff7fffu;
#pragma unroll
                            for (int j = 0; j < W; j++) m16 = __viaddmin_s16x2(__ldg(bk + j), pT[j], m16);
                            live = (m16 & 0x80008000u) == 0;
                        }
                        unsigned lmask = __ballot_sync(0xffffffffu, live);
                        c_live += __popc(lmask);
                        // test the points of the live blocks, four blocks per round: the 4*W key loads are
                        // issued back to back so that their L1/L2 latencies overlap
                        while (lmask) {
                            int bb[4];
                            unsigned m16[4];
#pragma unroll
                            for (int u = 0; u < 4; u++) {
                                bb[u] = lmask ? wblk + bp_block_of_bit(__ffs((int)lmask) - 1) : -1;
                                lmask &= lmask - 1;                                     // (0 & anything) stays 0
                            }
                            unsigned kw[4][W];
#pragma unroll
                            for (int u = 0; u < 4; u++) {
                                const unsigned *pk = a.pkeys + ((unsigned)(bb[u] < 0 ? wblk : bb[u]) * kBpBlock + lane) * W;
#pragma unroll
                                for (int j = 0; j < W; j++) kw[u][j] = __ldg(pk + j);
                            }
                            unsigned cm[4];
#pragma unroll
                            for (int u = 0; u < 4; u++) {
                                m16[u] = 0x7fff7fffu;
#pragma unroll
                                for (int j = 0; j < W; j++) m16[u] = __viaddmin_s16x2(kw[u][j], pT[j], m16[u]);
                                cm[u] = __ballot_sync(0xffffffffu, bb[u] >= 0 && (m16[u] & 0x80008000u) == 0);
                            }
#pragma unroll
                            for (int u = 0; u < 4; u++) {
                                unsigned cmask = cm[u];
                                while (cmask) {
                                    const int q = __ffs((int)cmask) - 1;
                                    cmask &= cmask - 1;
                                    // exact FP64 evaluation by the warp: lane = objective
                                    const ull cp = (ull)bb[u] * kBpBlock + q;
                                    double prod = __longlong_as_double(0x7ff0000000000000LL);
                                    if (lane < D) prod = __ldg(a.pts + cp * D + lane) * rlane;
#pragma unroll
                                    for (int off = 16; off > 0; off >>= 1) {
                                        const double o = __shfl_xor_sync(0xffffffffu, prod, off);
                                        prod = (o < prod) ? o : prod;
                                    }
                                    slow++;
                                    if (prod > bcur) {                             // uniform: every lane holds the min
                                        bcur = prod;
                                        bp_point_thresholds<D>(bcur, dir + D, a, lane, pT);
                                    }
                                }
                            }
                        }
                    }
                    if (lane == owner) best[r] = bcur;
                }
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
        // `slow` was counted identically by all lanes of a warp: report lane 0's
        if (lane == 0 && slow) atomicAdd(a.slow_counter, slow);
        if (lane == 0) { atomicAdd(a.slow_counter + 1, c_surv); atomicAdd(a.slow_counter + 2, c_rounds); atomicAdd(a.slow_counter + 3, c_live); atomicAdd(a.slow_counter + 4, c_dirs); }
    }
    __syncthreads();
    if (tid == 0)
        a.cta_sums[blockIdx.x] += ((red[0] + red[1]) + (red[2] + red[3])) + ((red[4] + red[5]) + (red[6] + red[7]));
}

// ---- key construction ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_bp_pkeys(const double *__restrict__ pts, ull n, ull n_pad, int d, int W,
                                                  const double *__restrict__ qlo, const double *__restrict__ qscale,
                                                  unsigned *__restrict__ pkeys)
{
    const ull w = (ull)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_pad * (ull)W) return;
    const ull i = w / W;
    const int j = (int)(w - i * W);
    unsigned word = 0;
    for (int h = 0; h < 2; h++) {
        const int k = 2 * j + h;
        unsigned key;
        if (i >= n) key = 0xffffu;                                   // padding point: -1, never a candidate
        else if (k >= d) key = (unsigned)kKeyScale;                  // dummy objective of an odd d: always passes
        else {
            double f = ceil((pts[i * d + k] - qlo[k]) * qscale[k]);
            f = f < 0.0 ? 0.0 : (f > (double)kKeyScale ? (double)kKeyScale : f);
            key = (unsigned)(int)f;
        }
        word |= key << (16 * h);
    }
    pkeys[w] = word;
}
__global__ void __launch_bounds__(256) k_bp_bkeys(const unsigned *__restrict__ pkeys, int nb, int W, unsigned *__restrict__ bkeys)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= nb * W) return;
    const int b = w / W, j = w - b * W;
    unsigned m = 0x80008000u;                                        // (-32768, -32768)
    for (int q = 0; q < kBpBlock; q++) m = __vmaxs2(m, pkeys[((size_t)b * kBpBlock + q) * W + j]);
    bkeys[w] = m;
}
__global__ void __launch_bounds__(256) k_bp_pairs(const unsigned *__restrict__ bkeys, int nb, int nb_pad, int d, int W,
                                                  unsigned *__restrict__ bpair)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= (nb_pad / 2) * d) return;
    const int pr = w / d, k = w - pr * d;
    unsigned word = 0;
    for (int h = 0; h < 2; h++) {
        const int b = 2 * pr + h;
        unsigned key = 0xffffu;                                      // padding block: -1
        if (b < nb) key = (bkeys[(size_t)b * W + k / 2] >> (16 * (k & 1))) & 0xffffu;
        word |= key << (16 * h);
    }
    bpair[w] = word;
}
