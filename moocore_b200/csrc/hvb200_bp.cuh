// hvb200_bp.cuh — block-pruned max-min kernel (included by hvb200_kernels.cu).
//
// Generalises the reference's scalar early-out (c/hvapprox.c:104-108) into a two-level search that
// keeps the reference's bits: max_p min_k p_k r_k only needs the points that can still raise a lower
// bound of the max, and whole blocks of 32 points can be discarded by their per-objective maxima.
// Every phase runs with all lanes busy; the (direction, block) survivor matrix is transposed between
// the phases so that the lane axis is always the long one.
//
//   order     kd-tree order of the points (recursive median splits on the widest normalised
//             objective): 32 consecutive points form a block with a tight bounding box
//   keys      the DPX 16-bit keys of k_maxmin_dpx (same monotone map, same conservative rounding):
//               pkeys [block][W][32]     two objectives per word          (phase C, lane = point)
//               bpair [block pair][D]    two block maxima per word        (phase B, lane = direction)
//   phase A   lane = direction.  FP32 max-min over the pivots (the points that win most probe
//             directions, k_probe_argmax) picks one pivot per direction; that pivot is evaluated in
//             FP64 with the reference's arithmetic -> b0, a value some point really attains, hence a
//             lower bound of the max whatever the FP32 ranking did.  Threshold keys from b0.
//   phase B   lane = direction.  Brute-force DPX screen of the block bounds (n/32 virtual points):
//             bitmask of the blocks that may hold a point above b0; 32x32 bit transposes (warp
//             shuffles) turn it into one word per (block, 32 directions) in shared memory.
//   phase C   lane = point.  One thread per block first turns the block's direction words into "quads" — four
//             surviving directions of the block per entry, padded with a nothing-passes row — in ONE list per CTA
//             (global memory, L2 resident; its slots come from a shared counter).  The eight warps then take equal
//             shares of that list whatever blocks the quads belong to: a warp keeps the keys of the current block in
//             registers (reloaded when the block changes), tests its 32 points against the four directions'
//             thresholds (W VIADDMNMX + one vote each) and every candidate lane appends {direction, point} to
//             the warp's part of the CTA's hit list.
//   phase D   lane = candidate.  Exact FP64 evaluation (the reference's products and comparisons), folded
//             into the direction's max with a shared-memory atomicMax on the bit pattern (max is exact and
//             order-free, values are >= 0).
//   Large point sets are processed in segments of blocks; the thresholds are refreshed from the
//   running max between segments.
//
// Every discarded point/block fails a conservative key test against thresholds derived from a value
// that is <= the final max, so the final max — and s^d — is bit-identical to the brute-force kernels.
#pragma once

#ifndef HVB200_BP_R2
#define HVB200_BP_R2 1
#endif
#ifndef HVB200_BP_ENDBAR
#define HVB200_BP_ENDBAR 0       // 1: the warps of a CTA start every chunk together (development switch)
#endif
#ifndef HVB200_BP_MINB
#define HVB200_BP_MINB 3
#endif
static constexpr int kBpBlock = 32;           // points per block = lanes per warp
static constexpr int kBpThreads = 256;        // 8 warps
static constexpr int kBpG = 4;                // block pairs per group in phase B ((G*D) % 4 == 0)
static constexpr int kBpMaxPivots = 512;
static constexpr int kBpHitCap = 16384;       // (direction, block) pairs with candidates per chunk: per-CTA list in global memory (L2)
static constexpr int kBpPtBits = 23;          // hit entry = {direction << 23 | block, 32-bit mask of candidate points}

struct BpArgs {
    const double *pts;        // [n_pad][D] FP64, block order
    const float *piv32;       // [n_piv][D] FP32 copies of the pivots
    const unsigned *piv_idx;  // [n_piv] row of each pivot in pts
    int n_piv;
    const unsigned *pkeys;    // [nb][W][32]
    const unsigned *bpair;    // [nb_pad/2][D]
    int nb;                   // blocks
    int nb_pad;               // multiple of 2*kBpG
    int seg_blocks;           // blocks per segment (multiple of 32)
    const double *dirs;       // [count][2][D]
    ull count;
    double *chunk_sums;       // one partial sum per chunk of this launch (fixed tree inside the chunk): the result does
                              // not depend on which CTA ran which chunk
    unsigned *chunk_counter;  // dynamic chunk hand-out (zeroed before the launch)
    double *values;
    ull *slow_counter;
    uint2 *hits;              // [grid][kBpHitCap] hit lists: {direction of the chunk, point} per candidate
    uint4 *quads;             // [grid]{[seg_blocks * CH / 4] x 8 bytes of row offsets, then x 4 bytes of block indices}: the CTA's
                              // list of surviving (direction, block) pairs, four directions of one block per entry (L2 resident)
    double qscale[HVB200_MAXD], qlos[HVB200_MAXD];
    // GEN = true (DZ2019-HW sums): the CTA generates the directions of its chunk itself (hw_direction) into a private
    // scratch block [2D][CH] (objective-major: coalesced for the lane = direction phases), which stays in L2 — there is
    // no direction buffer in HBM and no separate generator launch.  Position p of the launch is lattice index
    // gen_i0 + ((p >> gen_shift) * gen_stride << gen_shift) + (p & (2^gen_shift - 1)): a contiguous range
    // (gen_shift = 62) or the block-cyclic shard of one rank (blocks of 2^gen_shift directions, every gen_stride-th).
    HwDev hw;
    ull gen_i0, gen_stride;
    int gen_shift;
    const double *gen_tab;
    int gen_ntab;
    double *gen_scratch;      // [grid]{[D][CH] FP64 r, [D][CH] FP32 winv rounded down} = 12 D CH bytes per CTA
    ull *capped_counter;
    int early_out;            // phase C may end a quad between threshold planes (d >= 24); HVB200_BP_EARLYOUT=0 switches it off
};

// shared-memory carve-up, used by the kernel and by the host to size the launch.
// Thresholds for phase C are stored as planes of 16-byte words [plane][CH + 1] (row CH = "nothing
// passes", the padding entry of the pair lists); the last plane holds the 1, 2 or 3 tail words of a row.
struct BpSmem {
    size_t thr, tail, bnd, mask, best, piv, list, bar, red, total;
    int mask_stride, Q, TW;
};
__host__ __device__ inline BpSmem bp_smem_layout(int D, int R, int seg_blocks, int n_piv)
{
    const int W = (D + 1) / 2, CH = kBpThreads * R, DW = CH / 32;
    BpSmem s;
    s.Q = W / 4;
    const int rem = W - 4 * s.Q;
    s.TW = rem == 3 ? 4 : rem;
    size_t o = 0;
    s.thr = o;  o += (size_t)(s.Q + (s.TW ? 1 : 0)) * (CH + 1) * 16;      // the tail words sit in plane Q (16-byte rows
    s.tail = o;                                                            // like the others: one address per direction)
    s.bnd = o;  o += (size_t)(seg_blocks / 2) * D * 4;
    s.mask_stride = seg_blocks + 1;
    s.mask = o; o += ((size_t)DW * s.mask_stride * 4 + 15) & ~(size_t)15;
    s.best = o; o += (size_t)CH * 8;
    s.piv = o;  o += ((size_t)n_piv * D * 4 + 15) & ~(size_t)15;
    s.list = o; o += (size_t)(kBpThreads / 32) * 32 * 16;         // per warp: 32 staged quads of the CTA's pair list
    s.bar = o;  o += 128;            // mbarrier, counters, per-warp hit counts, per-warp early-out state
    s.red = o;  o += 8 * 8;
    s.total = o;
    return s;
}

// key of one objective in "negated" form, single half-word (not duplicated)
__device__ __forceinline__ unsigned bp_neg_key16(double best_m, double winv, double scale, double lo_s)
{
    const double x = fma(best_m * winv, scale, -lo_s);
    int kt = __double2int_rd(x);
    kt = max(0, min(kt, kKeyScale + 1));
    return (unsigned)(-kt) & 0xffffu;
}

// mask bit -> block index within its 32-block word (phase B writes bit 4h+g of byte `grp` for block 8*grp + 2g + h)
__device__ __forceinline__ int bp_block_of_bit(int bit)
{
    return (bit & 24) | ((bit & 3) << 1) | ((bit >> 2) & 1);
}

// 32x32 bit-matrix transpose across a warp: on return bit i of lane j = bit j of lane i on entry.
__device__ __forceinline__ unsigned bp_transpose32(unsigned x, int lane)
{
#pragma unroll
    for (int j = 16; j >= 1; j >>= 1) {
        const unsigned m = j == 16 ? 0x0000ffffu : (j == 8 ? 0x00ff00ffu : (j == 4 ? 0x0f0f0f0fu : (j == 2 ? 0x33333333u : 0x55555555u)));
        const unsigned y = __shfl_xor_sync(0xffffffffu, x, j);
        x = (lane & j) ? ((x & ~m) | ((y >> j) & m)) : ((x & m) | ((y << j) & ~m));
    }
    return x;
}

// shared-memory loads by 32-bit shared-window address: phase C's inner loop addresses four threshold rows per
// iteration, and generic pointers cost it a window-base computation per load
__device__ __forceinline__ uint4 bp_lds128(unsigned addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
// (the staged quads are rewritten between batches: this one orders itself against the surrounding stores)
__device__ __forceinline__ uint4 bp_lds128_ordered(unsigned addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint2 bp_lds64(unsigned addr)
{
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ unsigned bp_lds32(unsigned addr)
{
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// 3-input FP32 min (FMNMX3, sm_100+): NaN operands are ignored like fminf
__device__ __forceinline__ float bp_fmin3(float a, float b, float c)
{
    float d;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
static_assert(kBpMaxPivots <= 512, "the packed arg-max key of phase A holds a 9-bit pivot index");

// exact value of one point for one direction, the reference's arithmetic (c/hvapprox.c:81-102)
// (rr points at r_0 of the direction; RS = stride between its components: 1 in the row-major direction buffer, CH in a
// CTA's generated scratch block, which is read with L2-coherent loads because the same kernel wrote it)
template <int D, int RS>
__device__ __forceinline__ double bp_exact(const double *__restrict__ pp, const double *rr)
{
    auto ldr = [&](int k) { return RS == 1 ? __ldg(rr + k) : __ldcg(rr + (size_t)k * RS); };
    double m = __ldg(pp) * ldr(0);
#pragma unroll
    for (int k = 1; k < D; k++) {
        const double t = __ldg(pp + k) * ldr(k);
        m = (t < m) ? t : m;
    }
    return m;
}

// cdirs: the chunk's directions — row-major rows of the direction buffer, or the CTA's objective-major scratch block
template <int D, int R, bool GEN>
__device__ __forceinline__ void bp_candidate(const BpArgs &a, const double *cdirs, unsigned dl, unsigned pt, long long *bestbits)
{
    constexpr int CH = kBpThreads * R;
    const double m = GEN ? bp_exact<D, CH>(a.pts + (size_t)pt * D, cdirs + dl)
                         : bp_exact<D, 1>(a.pts + (size_t)pt * D, cdirs + (size_t)dl * (2 * D));
    if (m > 0.0) atomicMax(bestbits + dl, __double_as_longlong(m));
}

template <int D, int R, bool GEN>
__device__ __noinline__ void bp_candidate_cold(const BpArgs &a, const double *cdirs, unsigned dl, unsigned pt, long long *bestbits)
{
    bp_candidate<D, R, GEN>(a, cdirs, dl, pt, bestbits);
}

template <int D, int R, bool GEN>
__global__ void __launch_bounds__(kBpThreads, (D <= 16 ? HVB200_BP_MINB : 2)) k_maxmin_bp(const __grid_constant__ BpArgs a)
{
    constexpr int G = kBpG, W = (D + 1) / 2, CH = kBpThreads * R, DW = CH / 32;
    constexpr int Q = W / 4, REM = W - 4 * Q, TW = REM == 3 ? 4 : REM, CHP = CH + 1;
    static_assert((G * D) % 4 == 0, "group of block pairs must be whole 16-byte words");
    static_assert(CH <= (1 << (32 - kBpPtBits)), "direction index must fit the candidate entry");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const BpSmem L = bp_smem_layout(D, R, a.seg_blocks, a.n_piv);
    uint4 *thr4 = reinterpret_cast<uint4 *>(smem_raw + L.thr);               // [Q][CHP]
    uint4 *thrt = thr4 + Q * CHP;                                            // [CHP] tail words of a row (TW of them used)
    unsigned *bnd = reinterpret_cast<unsigned *>(smem_raw + L.bnd);          // [seg_blocks/2][D]
    unsigned *maskT = reinterpret_cast<unsigned *>(smem_raw + L.mask);       // [DW][mask_stride]
    long long *bestbits = reinterpret_cast<long long *>(smem_raw + L.best);  // [CH]
    float *piv = reinterpret_cast<float *>(smem_raw + L.piv);                // [n_piv][D]
    ull *bar = reinterpret_cast<ull *>(smem_raw + L.bar);
    unsigned *ctr = reinterpret_cast<unsigned *>(bar + 1);                   // [1] next block of phase C, [2], [3] next chunk (alternating)
    unsigned *hcnt = ctr + 4;                                                // [8] hit entries of each warp's list
    unsigned *eost = hcnt + 8;                                               // [8] early-out state of each warp: quads seen | hits << 16 | on << 31
    double *red = reinterpret_cast<double *>(smem_raw + L.red);
    const int mstride = L.mask_stride;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint4 *sq = reinterpret_cast<uint4 *>(smem_raw + L.list) + warp * 32;                          // this warp's staged quads
    // the CTA's pair list: per quad four 16-bit threshold-row offsets, and the quad's block
    const size_t qcap = (size_t)a.seg_blocks * (CH / 4);
    unsigned short *qoff = reinterpret_cast<unsigned short *>(reinterpret_cast<unsigned char *>(a.quads) + (size_t)blockIdx.x * qcap * 12);
    unsigned *qblk = reinterpret_cast<unsigned *>(qoff + qcap * 4);          // 8 + 4 bytes per quad
    if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); ctr[1] = 0; }
    if (tid < kBpThreads / 32) eost[tid] = a.early_out ? 0x80000000u : 0u;
    // every warp appends to its own eighth of the CTA's hit list: the count is a warp-uniform register, no atomics
    constexpr unsigned WCAP = kBpHitCap / (kBpThreads / 32);
    uint2 *hits = a.hits + (size_t)blockIdx.x * kBpHitCap;
    uint2 *whits = hits + (size_t)warp * WCAP;
    for (int i = tid; i < a.n_piv * D; i += kBpThreads) piv[i] = a.piv32[i];
    {
        // row CH of the threshold planes: the key -(S+1) in every half-word -> no point passes
        const unsigned v = ((unsigned)(-(kKeyScale + 1)) & 0xffffu) * 0x10001u;
        if (tid < Q) thr4[tid * CHP + CH] = make_uint4(v, v, v, v);
        if (TW && tid == 0) thrt[CH] = make_uint4(v, v, v, v);
    }
    __syncthreads();
    unsigned phase = 0;                                            // parity of the bounds barrier
    const int n_seg = (a.nb_pad + a.seg_blocks - 1) / a.seg_blocks;
    const bool resident = n_seg == 1;
    if (resident) {
        // all block bounds fit: one TMA bulk copy per CTA
        if (tid == 0) {
            const unsigned bytes = (unsigned)(a.nb_pad / 2) * D * 4u;
            mbar_arrive_expect_tx(bar, bytes);
            tma_load_1d(bnd, a.bpair, bytes, bar);
        }
        mbar_wait(bar, phase);
        phase ^= 1u;
    }

    const ull n_chunks = (a.count + CH - 1) / CH;
    ull slow = 0;
    unsigned capped = 0;           // GEN: Newton solves stopped by the iteration cap

    // chunks are handed out dynamically (their cost varies with the directions); the first one is static
    ull chunk = blockIdx.x;
    unsigned cpar = 0;                // which of ctr[2], ctr[3] carries this chunk's successor
    while (chunk < n_chunks) {
        if (tid == 0) ctr[2 + cpar] = atomicAdd(a.chunk_counter, 1u);      // consumed behind the barriers of this chunk
        const ull chunk_base = chunk * CH;
        if (!resident && tid == 0) {
            // first segment of this chunk; the previous chunk's phase B is behind several CTA barriers
            const unsigned bytes = (unsigned)(min(a.seg_blocks, a.nb_pad) / 2) * D * 4u;
            mbar_arrive_expect_tx(bar, bytes);
            tma_load_1d(bnd, a.bpair, bytes, bar);
        }
        bool valid[R];
        // the CTA's scratch block: r in FP64 (the exact evaluations read it), winv = RN(1/r) only feeds the threshold keys
        // and is kept in FP32 rounded DOWN (a lower threshold lets more candidates through, never fewer): 120 instead
        // of 160 bytes per direction of L2 footprint
        constexpr size_t kScratchDoubles = (size_t)D * CH + (size_t)D * CH / 2;
        const double *cdirs = GEN ? a.gen_scratch + (size_t)blockIdx.x * kScratchDoubles : a.dirs + chunk_base * (2 * D);
        const float *cwinv = reinterpret_cast<const float *>(a.gen_scratch + (size_t)blockIdx.x * kScratchDoubles + (size_t)D * CH);
        // component k of direction dl of this chunk: reciprocal weight r (k < D) and winv (D <= k < 2D)
        auto dirv = [&](int dl, int k) {
            if constexpr (GEN) return k < D ? __ldcg(cdirs + (size_t)k * CH + dl) : (double)__ldcg(cwinv + (size_t)(k - D) * CH + dl);
            else return __ldg(cdirs + (size_t)dl * (2 * D) + k);
        };
        if constexpr (GEN) {
            // ---- generator: every thread produces its R directions of the chunk (the previous chunk's readers are
            //      behind the barrier that ended it; the other threads read these rows only in phase D, behind barriers)
#pragma unroll 1
            for (int r = 0; r < R; r++) {
                const ull p = chunk_base + (ull)r * kBpThreads + tid;
                if (p < a.count) {
                    const ull i = a.gen_i0 + (((p >> a.gen_shift) * a.gen_stride) << a.gen_shift) + (p & ((1ull << a.gen_shift) - 1ull));
                    double *o = a.gen_scratch + (size_t)blockIdx.x * kScratchDoubles + r * kBpThreads + tid;
                    float *ow = reinterpret_cast<float *>(a.gen_scratch + (size_t)blockIdx.x * kScratchDoubles + (size_t)D * CH) + r * kBpThreads + tid;
                    hw_direction(a.hw, i, a.gen_tab, a.gen_ntab, capped,
                                 [&](int k, double rk, double wk) { __stcg(o + (size_t)k * CH, rk); __stcg(ow + (size_t)k * CH, __double2float_rd(wk)); });
                }
            }
        }
        // ---- phase A: FP32 ranking of the pivots, exact value of the best one
        {
            float rf[R][D];
            unsigned bkey[R];          // best pivot so far: {FP32 bits of its min with the low 9 bits dropped | pivot index}
#pragma unroll
            for (int r = 0; r < R; r++) {
                const ull smp = chunk_base + (ull)r * kBpThreads + tid;
                valid[r] = smp < a.count;
                const int dl = valid[r] ? r * kBpThreads + tid : 0;      // (an invalid slot ranks direction 0 and is never summed)
#pragma unroll
                for (int k = 0; k < D; k++) rf[r][k] = (float)dirv(dl, k);
                bkey[r] = 0u;
            }
            // Two pivots per iteration.  The min over the objectives runs on 3-input FMNMX3 and the running arg-max is
            // one VIMNMX3.U32 on packed keys (positive floats order like their bit patterns; 9 index bits replace
            // the low mantissa bits, i.e. near-ties within 6e-5 are broken by the index — any pivot gives a valid b0).
            // 33 instructions per pivot pair and direction instead of 44: the ALU pipe is this kernel's busiest.
#pragma unroll 2
            for (int i = 0; i < a.n_piv; i += 2) {
                const int i1 = min(i + 1, a.n_piv - 1);
                const float *pa = piv + i * D, *pb = piv + i1 * D;
                float ka[D], kb[D];
#pragma unroll
                for (int k = 0; k < D; k++) { ka[k] = pa[k]; kb[k] = pb[k]; }
#pragma unroll
                for (int r = 0; r < R; r++) {
                    float ta[D], tb[D];
#pragma unroll
                    for (int k = 0; k < D; k++) { ta[k] = ka[k] * rf[r][k]; tb[k] = kb[k] * rf[r][k]; }
                    float ma = ta[0], mb = tb[0];
#pragma unroll
                    for (int k = 1; k + 1 < D; k += 2) { ma = bp_fmin3(ma, ta[k], ta[k + 1]); mb = bp_fmin3(mb, tb[k], tb[k + 1]); }
                    if constexpr (D % 2 == 0) { ma = fminf(ma, ta[D - 1]); mb = fminf(mb, tb[D - 1]); }
                    const unsigned keya = (__float_as_uint(ma) & 0xfffffe00u) | (unsigned)i;
                    const unsigned keyb = (__float_as_uint(mb) & 0xfffffe00u) | (unsigned)i1;
                    bkey[r] = max(bkey[r], max(keya, keyb));
                }
            }
            int bi[R];
#pragma unroll
            for (int r = 0; r < R; r++) bi[r] = (int)(bkey[r] & 0x1ffu);
#pragma unroll
            for (int r = 0; r < R; r++) {
                double b = 0.0;
                if (valid[r] && a.n_piv > 0) {
                    const int dl = r * kBpThreads + tid;
                    const double *pp = a.pts + (size_t)__ldg(a.piv_idx + bi[r]) * D;
                    const double m = GEN ? bp_exact<D, CH>(pp, cdirs + dl) : bp_exact<D, 1>(pp, cdirs + (size_t)dl * (2 * D));
                    b = (m > 0.0) ? m : 0.0;
                }
                bestbits[r * kBpThreads + tid] = __double_as_longlong(b);
            }
        }

        for (int seg = 0; seg < n_seg; seg++) {
            const int seg0 = seg * a.seg_blocks;                                   // first block of the segment
            const int seg_nb = min(a.seg_blocks, a.nb_pad - seg0);                 // blocks in it (multiple of 2G)
            // ---- thresholds from the current lower bound: phase-B packing in registers (key duplicated in
            //      both halves), phase-C packing (two objectives per word) in shared memory
            unsigned negT[R][D];
#pragma unroll
            for (int r = 0; r < R; r++) {
                const int dl = r * kBpThreads + tid;
                if (valid[r]) {
                    const double best_m = __longlong_as_double(bestbits[dl]) * (1.0 - 0x1p-20);
#pragma unroll
                    for (int k = 0; k < D; k++) negT[r][k] = bp_neg_key16(best_m, dirv(dl, D + k), a.qscale[k], a.qlos[k]);
                    unsigned tw[4 * Q + 4];
#pragma unroll
                    for (int j = 0; j < 4 * Q + 4; j++) {
                        tw[j] = 0;
                        if (2 * j < D) tw[j] = negT[r][2 * j];
                        if (2 * j + 1 < D) tw[j] |= negT[r][2 * j + 1] << 16;   // dummy objective of an odd D: threshold key 0
                    }
#pragma unroll
                    for (int q = 0; q < Q; q++) thr4[q * CHP + dl] = make_uint4(tw[4 * q], tw[4 * q + 1], tw[4 * q + 2], tw[4 * q + 3]);
                    if constexpr (TW == 1) *reinterpret_cast<unsigned *>(thrt + dl) = tw[4 * Q];
                    if constexpr (TW == 2) *reinterpret_cast<uint2 *>(thrt + dl) = make_uint2(tw[4 * Q], tw[4 * Q + 1]);
                    if constexpr (TW == 4) thrt[dl] = make_uint4(tw[4 * Q], tw[4 * Q + 1], tw[4 * Q + 2], tw[4 * Q + 3]);
#pragma unroll
                    for (int k = 0; k < D; k++) negT[r][k] *= 0x10001u;
                } else {
                    const unsigned v = ((unsigned)(-(kKeyScale + 1)) & 0xffffu) * 0x10001u;      // nothing passes
#pragma unroll
                    for (int k = 0; k < D; k++) negT[r][k] = v;
                }
            }
            if (!resident) {
                mbar_wait(bar, phase);
                phase ^= 1u;
            }

            // ---- phase B: screen the block bounds, lane = direction
            {
                const uint4 *bnd4 = reinterpret_cast<const uint4 *>(bnd + (resident ? (size_t)(seg0 / 2) * D : 0));
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
                    }
                    if ((gi & 3) == 3 || gi == n_groups - 1) {
                        // 32 blocks x 32 directions of this warp: transpose, lane = block afterwards
#pragma unroll
                        for (int r = 0; r < R; r++) {
                            const unsigned t = bp_transpose32(wacc[r], lane);
                            maskT[(r * (kBpThreads / 32) + warp) * mstride + (gi >> 2) * 32 + bp_block_of_bit(lane)] = t;
                            wacc[r] = 0;
                        }
                    }
                }
            }
            __syncthreads();                  // maskT, thresholds, bestbits complete; bnd no longer read
            if (!resident && tid == 0 && seg + 1 < n_seg) {
                const int nxt = min(a.seg_blocks, a.nb_pad - (seg0 + a.seg_blocks));
                const unsigned bytes = (unsigned)(nxt / 2) * D * 4u;
                mbar_arrive_expect_tx(bar, bytes);
                tma_load_1d(bnd, a.bpair + (size_t)((seg0 + a.seg_blocks) / 2) * D, bytes, bar);
            }

            // ---- phase C: the CTA's list of quads, then lane = point
            const int seg_real = min(seg_nb, a.nb - seg0);                         // real (non-padding) blocks
            // lane = (block, direction word): DW lanes share a block.  A segmented warp scan of the words' bit counts gives
            // every lane the position of its first direction in the block's run of quads; the slots of the run come from
            // the CTA's counter.  No lane expands more than one 32-bit word, so a crowded block costs what a sparse one does.
            {
                constexpr int BPS = 32 / DW;                                       // blocks per warp step
                const int sub = lane / DW, wl = lane % DW;
#pragma unroll 1
                for (int b0 = warp * BPS; b0 < seg_real; b0 += (kBpThreads / 32) * BPS) {
                    const int b = b0 + sub;
                    unsigned m = b < seg_real ? maskT[wl * mstride + b] : 0u;
                    const int cnt = __popc(m);
                    int incl = cnt;
#pragma unroll
                    for (int off = 1; off < DW; off <<= 1) {
                        const int t = __shfl_up_sync(0xffffffffu, incl, off, DW);
                        if (wl >= off) incl += t;
                    }
                    const int total = __shfl_sync(0xffffffffu, incl, DW - 1, DW);
                    if (__ballot_sync(0xffffffffu, total != 0) == 0u) continue;
                    const unsigned nq = (unsigned)(total + 3) >> 2;
                    unsigned qbase = 0;
                    if (wl == 0 && total) qbase = atomicAdd(ctr + 1, nq);
                    qbase = __shfl_sync(0xffffffffu, qbase, 0, DW);
                    unsigned short *run = qoff + (size_t)qbase * 4;                // 16-bit byte offsets of threshold rows
                    unsigned short *dst = run + (incl - cnt);
                    while (m) {
                        const unsigned off = (unsigned)(wl * 32 + __ffs((int)m) - 1) * 16u;
                        m &= m - 1;
                        __stcg(dst++, (unsigned short)off);             // the list lives in L2: written and read once per chunk
                    }
                    if (wl == DW - 1)
                        for (int p = total; p < (int)(nq * 4); p++) __stcg(run + p, (unsigned short)(CH * 16));      // row CH: nothing passes
                    for (unsigned i = wl; i < nq; i += DW) __stcg(qblk + qbase + i, (unsigned)b);
                }
            }
            __syncthreads();
            unsigned hcount = 0;                                                   // entries in this warp's hit list
            {
                const unsigned nquads = ctr[1];
                // early-out state of this warp (it outlives segments and chunks: parked in shared memory between phases)
                unsigned eo_cnt = eost[warp] & 0xffffu, eo_hit = (eost[warp] >> 16) & 0x7fffu;
                bool eo_on = (eost[warp] >> 31) != 0u;
                const unsigned qlo = (unsigned)((ull)nquads * (unsigned)warp / (kBpThreads / 32));
                const unsigned qhi = (unsigned)((ull)nquads * (unsigned)(warp + 1) / (kBpThreads / 32));
                unsigned tb_s = (unsigned)__cvta_generic_to_shared(thr4);
                unsigned sq_s = (unsigned)__cvta_generic_to_shared(sq);
                asm volatile("" : "+r"(tb_s), "+r"(sq_s));          // computed once: not rematerialised inside the loop
                const unsigned lt_mask = (1u << lane) - 1u;
                const uint2 *qo2 = reinterpret_cast<const uint2 *>(qoff);
                uint4 nxt = make_uint4(0u, 0u, 0u, 0u);
                if (qlo + lane < qhi) { const uint2 o = __ldcg(qo2 + qlo + lane); nxt = make_uint4(o.x, o.y, __ldcg(qblk + qlo + lane), 0u); }
#pragma unroll 1
                for (unsigned base = qlo; base < qhi; base += 32) {
                    const int nq = (int)min(32u, qhi - base);
                    sq[lane] = nxt;
                    __syncwarp();
                    if (base + 32 + lane < qhi) {                                   // next batch, under this one's work
                        const uint2 o = __ldcg(qo2 + base + 32 + lane);
                        nxt = make_uint4(o.x, o.y, __ldcg(qblk + base + 32 + lane), 0u);
                    }
                    int j = 0;
                    uint4 e = bp_lds128_ordered(sq_s);
#pragma unroll 1
                    while (j < nq) {
                        // a run of quads of one block: its keys stay in registers
                        const unsigned cur = e.z;
                        unsigned kw[W];
                        {
                            const unsigned *kp = a.pkeys + (size_t)(seg0 + (int)cur) * (W * 32) + lane;
#pragma unroll
                            for (int i = 0; i < W; i++) kw[i] = __ldg(kp + i * 32);
                        }
                        const unsigned pt = (unsigned)(seg0 + (int)cur) * kBpBlock + lane;
#pragma unroll 1
                        do {
                            j++;
                            const uint4 e_next = bp_lds128_ordered(sq_s + (unsigned)j * 16u);      // one quad ahead (entry 32 is never used)
                            const unsigned off[4] = {e.x & 0xffffu, e.x >> 16, e.y & 0xffffu, e.y >> 16};
                            unsigned ad[4];
#pragma unroll
                            for (int u = 0; u < 4; u++) ad[u] = tb_s + off[u];
                            unsigned m16[4] = {0x7fff7fffu, 0x7fff7fffu, 0x7fff7fffu, 0x7fff7fffu};
                            // Many objectives (d >= 24: three planes or more): a point that fails one objective stays out
                            // (the running minimum keeps its sign), and at d = 50 a block bound prunes little — 73 % of the
                            // blocks survive phase B at cfg5 — so nearly every quad dies long before its last plane.  From the
                            // second plane on a vote ends the quad when none of its 128 (point, direction) pairs is alive.
                            // Where the block bounds already prune well (fronts) the vote never fires and costs 2-4 %: every
                            // warp watches how often it does over windows of 64 quads, drops it below one quad in four and
                            // probes again every 16th window.
                            constexpr bool kEarlyOut = Q >= 3;
                            const bool try_out = kEarlyOut && eo_on;
                            bool live = true;
#pragma unroll
                            for (int q = 0; q < Q; q++) {
                                if (!kEarlyOut || live) {
                                    uint4 t[4];
#pragma unroll
                                    for (int u = 0; u < 4; u++) t[u] = bp_lds128(ad[u] + q * CHP * 16);
#pragma unroll
                                    for (int u = 0; u < 4; u++) {
                                        m16[u] = __viaddmin_s16x2(kw[4 * q + 0], t[u].x, m16[u]);
                                        m16[u] = __viaddmin_s16x2(kw[4 * q + 1], t[u].y, m16[u]);
                                        m16[u] = __viaddmin_s16x2(kw[4 * q + 2], t[u].z, m16[u]);
                                        m16[u] = __viaddmin_s16x2(kw[4 * q + 3], t[u].w, m16[u]);
                                    }
                                    if constexpr (kEarlyOut) {
                                        if (try_out && q >= 1 && (q + 1 < Q || TW > 0))
                                            live = __any_sync(0xffffffffu, ((m16[0] & 0x80008000u) == 0u) | ((m16[1] & 0x80008000u) == 0u) |
                                                                           ((m16[2] & 0x80008000u) == 0u) | ((m16[3] & 0x80008000u) == 0u));
                                    }
                                }
                            }
                            if constexpr (kEarlyOut) {
                                eo_hit += live ? 0u : 1u;
                                if ((++eo_cnt & 63u) == 0u) {
                                    eo_on = a.early_out && (eo_on ? eo_hit >= 16u : ((eo_cnt >> 6) & 15u) == 0u);
                                    eo_hit = 0u;
                                }
                                if (!live) { e = e_next; continue; }
                            }
                            if constexpr (TW == 1) {
                                unsigned t[4];
#pragma unroll
                                for (int u = 0; u < 4; u++) t[u] = bp_lds32(ad[u] + Q * CHP * 16);
#pragma unroll
                                for (int u = 0; u < 4; u++) m16[u] = __viaddmin_s16x2(kw[4 * Q], t[u], m16[u]);
                            }
                            if constexpr (TW == 2) {
                                uint2 t[4];
#pragma unroll
                                for (int u = 0; u < 4; u++) t[u] = bp_lds64(ad[u] + Q * CHP * 16);
#pragma unroll
                                for (int u = 0; u < 4; u++) {
                                    m16[u] = __viaddmin_s16x2(kw[4 * Q], t[u].x, m16[u]);
                                    m16[u] = __viaddmin_s16x2(kw[4 * Q + 1], t[u].y, m16[u]);
                                }
                            }
                            if constexpr (TW == 4) {
                                uint4 t[4];
#pragma unroll
                                for (int u = 0; u < 4; u++) t[u] = bp_lds128(ad[u] + Q * CHP * 16);
#pragma unroll
                                for (int u = 0; u < 4; u++) {
                                    m16[u] = __viaddmin_s16x2(kw[4 * Q], t[u].x, m16[u]);
                                    m16[u] = __viaddmin_s16x2(kw[4 * Q + 1], t[u].y, m16[u]);
                                    m16[u] = __viaddmin_s16x2(kw[4 * Q + 2], t[u].z, m16[u]);
                                }
                            }
                            // sign bits clear in both halves <=> candidate.  About a quarter of the pairs hold one: every
                            // candidate lane publishes its own {direction, point} entry
                            unsigned c[4];
#pragma unroll
                            for (int u = 0; u < 4; u++) c[u] = __ballot_sync(0xffffffffu, (m16[u] & 0x80008000u) == 0u);
                            if (c[0] | c[1] | c[2] | c[3]) {
#pragma unroll
                                for (int u = 0; u < 4; u++) {
                                    if (c[u]) {
                                        if ((c[u] >> lane) & 1u) {
                                            const unsigned slot = hcount + __popc(c[u] & lt_mask);
                                            if (slot < WCAP) whits[slot] = make_uint2(off[u] >> 4, pt);
                                            else {
                                                // the warp's hit list is full (never seen in practice): evaluate in place
                                                bp_candidate_cold<D, R, GEN>(a, cdirs, off[u] >> 4, pt, bestbits);
                                                slow++;
                                            }
                                        }
                                        hcount += __popc(c[u]);
                                    }
                                }
                            }
                            e = e_next;
                        } while (j < nq && e.z == cur);
                    }
                    __syncwarp();                 // the staged quads are overwritten by the next batch
                }
                if constexpr (Q >= 3) { if (lane == 0) eost[warp] = (eo_cnt & 0xffffu) | (eo_hit << 16) | (eo_on ? 0x80000000u : 0u); }
            }
            if (lane == 0) hcnt[warp] = min(hcount, WCAP);
            __syncthreads();

            // ---- phase D: lane = candidate, exact FP64 evaluation
            {
                // the eight warp lists as one flat index range: entry i lives in list w at i - first[w]
                unsigned first[kBpThreads / 32 + 1];
                first[0] = 0;
#pragma unroll
                for (int w = 0; w < kBpThreads / 32; w++) first[w + 1] = first[w] + hcnt[w];
                for (unsigned i = tid; i < first[kBpThreads / 32]; i += kBpThreads) {
                    unsigned w = 0, base = 0;
#pragma unroll
                    for (int t = 1; t < kBpThreads / 32; t++)
                        if (i >= first[t]) { w = t; base = first[t]; }
                    const uint2 e = hits[(size_t)w * WCAP + (i - base)];
                    bp_candidate<D, R, GEN>(a, cdirs, e.x, e.y, bestbits);
                    slow++;
                }
            }
            __syncthreads();
            if (tid == 0) ctr[1] = 0;
        }

        double acc = 0.0;
#pragma unroll
        for (int r = 0; r < R; r++) {
            const double v = pow_chain<D>(__longlong_as_double(bestbits[r * kBpThreads + tid]));
            if (valid[r]) {
                acc += v;
                if (a.values) a.values[chunk_base + (ull)r * kBpThreads + tid] = v;
            }
        }
        // fixed-shape reduction of the chunk: the lanes of a warp here, the 8 warp sums of every chunk in k_reduce (one
        // slot per (chunk, warp): the result does not depend on which CTA ran which chunk, and no CTA barrier is needed)
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, off);
        if (lane == 0) a.chunk_sums[chunk * (kBpThreads / 32) + warp] = acc;
        chunk = (ull)gridDim.x + ctr[2 + cpar];      // written by thread 0 before this chunk's barriers; its twin serves the next chunk
        cpar ^= 1u;
#if HVB200_BP_ENDBAR
        __syncthreads();
#endif
    }

#pragma unroll
    for (int off = 16; off > 0; off >>= 1) slow += __shfl_down_sync(0xffffffffu, slow, off);
    if (a.slow_counter && lane == 0 && slow) atomicAdd(a.slow_counter, slow);
    if (GEN && capped) atomicAdd(a.capped_counter, (ull)capped);
}
