// hvb200_mcgen.cuh — device generator for the DZ2019-MC direction stream (included by
// hvb200_kernels.cu; uses DevPlan, CUDA_TRY, span_begin/span_end from there).
//
// Reproduces, on the GPU, the exact stream of standard normals the reference draws on the host:
//   mt19937_seed / mt19937_gen / tempering        c/mt19937/mt19937.c:6-55, c/mt19937/mt19937.h:28-44
//   next64 = hi<<32|lo, next_double = (a*2^26+b)/2^53   c/mt19937/mt19937.h:46-57
//   256-layer ziggurat rng_standard_normal         c/rng.c:71-99 (tables c/ziggurat_constants.h)
//
// The stream is sequential by definition (a rejected draw shifts everything behind it), so it is
// decomposed into data-parallel passes over a batch of 64-bit "pairs" (every draw of the reference
// consumes exactly one pair):
//   k_mt_generate   one CTA advances the 624-word state block by block (3 dependent phases of up to
//                   227 independent words each) and writes tempered words
//   k_zig_classify  for EVERY pair j, in parallel: if an attempt of the ziggurat started at j, how
//                   many pairs L_j would it consume, would it emit a normal, and which value
//   scan #1         "which pairs really start an attempt": the skip counter s' = s>0 ? s-1 : L_j-1 is a
//                   composition of per-pair maps on {0..15}; maps are packed 4 bits per entry in a
//                   64-bit word and composed with an inclusive scan (associative, non-commutative)
//   scan #2         exclusive sum of "emits" -> position of each normal in the output
//   k_zig_scatter   writes the normals, and the pair index at which the next batch must start
// CUB's DeviceScan does the two scans.  log1p/exp in the two slow paths are CUDA's (<= 1 ulp from
// glibc): tail values can differ in the last bit, accept/reject decisions only at exact ties.
#pragma once
#include <cub/device/device_scan.cuh>

__constant__ ull c_zig_ki[256];
__constant__ double c_zig_wi[256];
__constant__ double c_zig_fi[256];

static constexpr int kMtN = 624, kMtM = 397;
static constexpr int kZigMaxL = 16;           // longest attempt (in pairs) the skip tables can express

struct McGen {
    unsigned *state = nullptr;                // 624 words
    unsigned *words = nullptr; size_t words_cap = 0;   // tempered output, 2 words per pair
    unsigned char *L = nullptr;               // pairs consumed by an attempt starting here (0 = needs lookahead)
    unsigned char *emit = nullptr;
    double *val = nullptr;
    ull *ftab = nullptr, *fscan = nullptr;
    unsigned *eflag = nullptr, *epos = nullptr;
    size_t pairs_cap = 0;
    void *cub_tmp = nullptr; size_t cub_tmp_bytes = 0;
    ull *scalars = nullptr;                   // [0] next start pair, [1] emitted total, [2] error flags
    ull *scalars_host = nullptr;              // pinned mirror
    size_t have_pairs = 0;                    // valid pairs at the head of `words`
    bool tables_ready = false;
    // parallel MT19937 generation by jump-ahead (hvb200_mtjump.c): start states of the segments, state sequences
    unsigned *seg_states = nullptr; size_t seg_cap = 0;      // [segments][624]
    unsigned *xbuf = nullptr; size_t xbuf_cap = 0;           // [jumps of one round][34 * 624]
};

// ---- MT19937 block generation ------------------------------------------------------------------
__device__ __forceinline__ unsigned mt_twist(unsigned cur, unsigned nxt, unsigned far_)
{
    const unsigned y = (cur & 0x80000000u) | (nxt & 0x7fffffffu);
    return far_ ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}
__device__ __forceinline__ unsigned mt_temper(unsigned y)
{
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}
// CTA s generates the blocks [s*seg_blocks, min((s+1)*seg_blocks, nblocks_total)) that follow ITS state
// states[s] (624 words each) and writes them, tempered or raw, to out + s*out_stride (+624 when the state itself
// is copied in front: the "state sequence" consumed by k_mt_jump).  The last CTA leaves its final state in
// final_state.  x[n] = x[n-227] ^ twist(x[n-624], x[n-623]): 227 consecutive words are independent, so a
// block is three dependent phases.  Old and new blocks live in two shared arrays (ping-pong), which leaves one
// barrier per phase; tempering and the global store of a block overlap the next block's phases.
__global__ void __launch_bounds__(256) k_mt_generate(const unsigned *states, unsigned *__restrict__ out,
                                                     int nblocks_total, int seg_blocks, size_t out_stride,
                                                     unsigned *final_state, int temper, int copy_state)   // final_state may alias states
{
    __shared__ unsigned xs[2][kMtN];
    const int t = threadIdx.x;
    const int first = blockIdx.x * seg_blocks;
    const int nblocks = min(seg_blocks, nblocks_total - first);
    const unsigned *state = states + (size_t)blockIdx.x * kMtN;
    unsigned *obase = out + (size_t)blockIdx.x * out_stride;
    for (int i = t; i < kMtN; i += 256) {
        const unsigned v = state[i];
        xs[0][i] = v;
        if (copy_state) obase[i] = v;
    }
    if (copy_state) obase += kMtN;
    __syncthreads();
    int cur = 0;
    for (int b = 0; b < nblocks; b++) {
        const unsigned *xo = xs[cur];
        unsigned *xn = xs[cur ^ 1];
        unsigned *o = obase + (size_t)b * kMtN;
        // phase 1: i in [0, 227)      old x[i], x[i+1], x[i+397]
        if (t < 227) {
            const unsigned v = mt_twist(xo[t], xo[t + 1], xo[t + kMtM]);
            xn[t] = v;
            o[t] = temper ? mt_temper(v) : v;
        }
        __syncthreads();
        // phase 2: i in [227, 454)    old x[i], x[i+1], new x[i-227]
        if (t < 227) {
            const int i = 227 + t;
            const unsigned v = mt_twist(xo[i], xo[i + 1], xn[i - 227]);
            xn[i] = v;
            o[i] = temper ? mt_temper(v) : v;
        }
        __syncthreads();
        // phase 3: i in [454, 624)    old x[i], x[i+1] (new x[0] for i = 623), new x[i-227]
        if (t < kMtN - 454) {
            const int i = 454 + t;
            const unsigned nxt = (i + 1 == kMtN) ? xn[0] : xo[i + 1];
            const unsigned v = mt_twist(xo[i], nxt, xn[i - 227]);
            xn[i] = v;
            o[i] = temper ? mt_temper(v) : v;
        }
        __syncthreads();
        cur ^= 1;
    }
    if (final_state != nullptr && blockIdx.x == gridDim.x - 1)
        for (int i = t; i < kMtN; i += 256) final_state[i] = xs[cur][i];
}

// Jump-ahead (hvb200_mtjump.c): the state J words ahead is the GF(2) convolution of the state sequence X with
// the polynomial t^J mod phi, given as the sorted list of its set coefficients:
//     out[p] = XOR over listed i of X[i + p],   p = 0..623.
// grid = (jumps, kMtSplits): CTA (j, sp) handles the coefficients in [sp*1248, (sp+1)*1248) — a 1872-word
// window of X in shared memory — and XORs its partial result into out (zeroed by the caller).
static constexpr int kMtSplits = 16, kMtSlice = 1248;       // 16 * 1248 = 19968 >= 19937
static constexpr int kMtSeqBlocks = 33;                      // 624 + 33*624 = 21216 >= 19937 + 623 words
static constexpr int kMtSeqWords = (kMtSeqBlocks + 1) * kMtN;
__global__ void __launch_bounds__(256) k_mt_jump(const unsigned *__restrict__ X, const unsigned short *__restrict__ pos,
                                                 const int *__restrict__ slice_off, unsigned *__restrict__ out)
{
    __shared__ unsigned xs[kMtSlice + kMtN];
    __shared__ unsigned short ps[kMtSlice];
    const int t = threadIdx.x, j = blockIdx.x, sp = blockIdx.y;
    const int base = sp * kMtSlice;
    const unsigned *x = X + (size_t)j * kMtSeqWords + base;
    for (int i = t; i < kMtSlice + kMtN; i += 256) xs[i] = x[i];
    const int p0 = slice_off[sp], n = slice_off[sp + 1] - p0;
    for (int i = t; i < n; i += 256) ps[i] = (unsigned short)(pos[p0 + i] - base);
    __syncthreads();
    unsigned a0 = 0, a1 = 0, a2 = 0;
    const bool third = t + 512 < kMtN;
    for (int i = 0; i < n; i++) {
        const int q = ps[i] + t;
        a0 ^= xs[q];
        a1 ^= xs[q + 256];
        if (third) a2 ^= xs[q + 512];
    }
    unsigned *o = out + (size_t)j * kMtN;
    atomicXor(o + t, a0);
    atomicXor(o + t + 256, a1);
    if (third) atomicXor(o + t + 512, a2);
}

// ---- ziggurat, one hypothetical attempt per pair -------------------------------------------------
__device__ __forceinline__ double pair_to_unit(unsigned w0, unsigned w1)
{
    const int a = (int)(w0 >> 5), b = (int)(w1 >> 6);
    return (a * 67108864.0 + b) / 9007199254740992.0;
}
// emit codes: 0 = no normal, 1 = normal, 2 = tail loop beyond the skip tables, 3 = the accept/reject comparison of a slow
// path is a near-tie (its two sides agree to tie_eps relative: CUDA's log1p/exp and glibc's may then decide differently,
// SURVEY H3) — 2 and 3 stop the device generator when the attempt is real, and the host generator takes over.
__global__ void __launch_bounds__(256) k_zig_classify(const unsigned *__restrict__ words, ull npairs,
                                                      unsigned char *__restrict__ Lout, unsigned char *__restrict__ emit,
                                                      double *__restrict__ val, ull *__restrict__ ftab, ull *scalars, double tie_eps)
{
    // the 256-layer tables are indexed by random bits: constant memory would serialise the 32 lanes of a warp
    __shared__ double s_wi[256], s_fi[256];
    __shared__ ull s_ki[256];
    s_wi[threadIdx.x] = c_zig_wi[threadIdx.x];
    s_fi[threadIdx.x] = c_zig_fi[threadIdx.x];
    s_ki[threadIdx.x] = c_zig_ki[threadIdx.x];
    __syncthreads();
    const ull j = (ull)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= npairs) return;
    const double R = __longlong_as_double((long long)HVB200_ZIG_R_BITS);
    const double INV_R = __longlong_as_double((long long)HVB200_ZIG_INV_R_BITS);
    const ull r = ((ull)words[2 * j] << 32) | words[2 * j + 1];
    const int layer = (int)(r & 0xff);
    const ull mag = (r >> 9) & 0x000fffffffffffffull;
    double x = (double)mag * s_wi[layer];
    if ((r >> 8) & 1) x = -x;
    int L = 1, em = 1;
    double v = x;
    if (!(mag < s_ki[layer])) {
        if (layer == 0) {               // tail: pairs j+1, j+2, then j+3, j+4, ...
            L = 0;
            for (int t = 0; 1 + 2 * (t + 1) <= kZigMaxL; t++) {
                const ull a = j + 1 + 2 * (ull)t;
                if (a + 1 >= npairs) { L = 0; em = 0; break; }          // needs pairs beyond this batch
                const double xx = -INV_R * log1p(-pair_to_unit(words[2 * a], words[2 * a + 1]));
                const double yy = -log1p(-pair_to_unit(words[2 * a + 2], words[2 * a + 3]));
                if (fabs((yy + yy) - xx * xx) <= tie_eps * (xx * xx)) { L = 1 + 2 * (t + 1); em = 3; break; }
                if (yy + yy > xx * xx) {
                    v = ((mag >> 8) & 1) ? -(R + xx) : R + xx;
                    L = 1 + 2 * (t + 1);
                    break;
                }
                if (1 + 2 * (t + 2) > kZigMaxL) { L = kZigMaxL; em = 2; break; }      // beyond the skip tables: reported if this attempt is real
            }
        } else {                        // wedge: pair j+1
            if (j + 1 >= npairs) { L = 0; em = 0; }
            else {
                const double u = pair_to_unit(words[2 * j + 2], words[2 * j + 3]);
                const double f_hi = s_fi[layer - 1], f_lo = s_fi[layer];
                L = 2;
                const double lhs = (f_hi - f_lo) * u + f_lo, rhs = exp(-0.5 * x * x);
                em = fabs(lhs - rhs) <= tie_eps * rhs ? 3 : (lhs < rhs ? 1 : 0);
            }
        }
    }
    if (L == 0) em = 0;
    Lout[j] = (unsigned char)L;
    emit[j] = (unsigned char)em;
    val[j] = v;
    // skip map on s in {0..15}: s > 0 -> s-1; s == 0 -> L-1.  An attempt that cannot be resolved in
    // this batch (L == 0) behaves like L = 1 here; it lies behind the last normal that is used.
    const ull first = (ull)((L ? L : 1) - 1);
    ftab[j] = 0xEDCBA98765432100ull | first;     // entry s (4 bits at 4s) = s-1 for s>=1, entry 0 = L-1
}
struct SkipCompose {
    // (g after f)[s] = g[f[s]]; operands arrive as (earlier, later)
    __device__ __forceinline__ ull operator()(const ull &f, const ull &g) const
    {
        ull out = 0;
#pragma unroll
        for (int s = 0; s < 16; s++) {
            const unsigned fs = (unsigned)(f >> (4 * s)) & 15u;
            out |= ((g >> (4 * fs)) & 15ull) << (4 * s);
        }
        return out;
    }
};
__global__ void __launch_bounds__(256) k_zig_flags(const ull *__restrict__ fscan, const unsigned char *__restrict__ emit,
                                                   ull npairs, unsigned *__restrict__ eflag, ull *scalars)
{
    const ull j = (ull)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= npairs) return;
    const bool start = (j == 0) || ((fscan[j - 1] & 15ull) == 0);      // skip counter before pair j, from s = 0
    // a tail loop longer than the skip tables express (1.5e-12 per pair) only matters when an attempt really starts here
    if (start && emit[j] >= 2) atomicOr(&scalars[2], emit[j] == 2 ? 1ull : 2ull);
    eflag[j] = (start && emit[j] == 1) ? 1u : 0u;
}
__global__ void __launch_bounds__(256) k_zig_scatter(const unsigned *__restrict__ eflag, const unsigned *__restrict__ epos,
                                                     const double *__restrict__ val, const unsigned char *__restrict__ L,
                                                     ull npairs, ull need, double *__restrict__ normals, ull *scalars)
{
    const ull j = (ull)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= npairs) return;
    if (eflag[j]) {
        const ull q = epos[j];
        if (q < need && normals) normals[q] = val[j];
        if (q == need - 1) scalars[0] = j + L[j];                      // next batch starts behind this attempt
    }
    if (j == npairs - 1) scalars[1] = (ull)epos[j] + eflag[j];         // normals available in this batch
}

// Range end of a pair-addressed segment (mcgen_count_range): the attempts partition the pairs, so exactly one real
// attempt start j satisfies j < limit <= j + L_j.  It reports where the attempts behind the limit begin and how many
// normals the attempts that start before the limit emit.
__global__ void __launch_bounds__(256) k_zig_range_end(const ull *__restrict__ fscan, const unsigned *__restrict__ eflag,
                                                       const unsigned *__restrict__ epos, const unsigned char *__restrict__ L,
                                                       ull npairs, ull limit, ull *scalars)
{
    const ull j = (ull)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= npairs || j >= limit) return;
    const bool start = (j == 0) || ((fscan[j - 1] & 15ull) == 0);
    const ull Lj = L[j] ? L[j] : 1;
    if (start && j + Lj >= limit) {
        scalars[0] = j + Lj;
        scalars[1] = (ull)epos[j] + eflag[j];
    }
}

// ---- host side -----------------------------------------------------------------------------------
static int mcgen_reserve(DevPlan *dp, McGen *g, size_t pairs)
{
    if (g->pairs_cap >= pairs) return 0;
    const size_t cap = pairs + pairs / 8 + 4096;
    // the head of `words` may hold pairs left over from the previous batch: carry them over
    unsigned *nw = nullptr;
    CUDA_TRY(cudaMalloc(&nw, (2 * cap + 2 * kMtN) * sizeof(unsigned)));
    if (g->words && g->have_pairs) {
        CUDA_TRY(cudaMemcpyAsync(nw, g->words, 2 * g->have_pairs * sizeof(unsigned), cudaMemcpyDeviceToDevice, dp->stream));
        CUDA_TRY(cudaStreamSynchronize(dp->stream));
    }
    cudaFree(g->words);
    g->words = nw;
    g->words_cap = 2 * cap + 2 * kMtN;
    cudaFree(g->L); cudaFree(g->emit); cudaFree(g->val); cudaFree(g->ftab); cudaFree(g->fscan);
    cudaFree(g->eflag); cudaFree(g->epos); cudaFree(g->cub_tmp);
    g->L = nullptr; g->emit = nullptr; g->val = nullptr; g->ftab = nullptr; g->fscan = nullptr;
    g->eflag = nullptr; g->epos = nullptr; g->cub_tmp = nullptr; g->pairs_cap = 0;
    CUDA_TRY(cudaMalloc(&g->L, cap));
    CUDA_TRY(cudaMalloc(&g->emit, cap));
    CUDA_TRY(cudaMalloc(&g->val, cap * sizeof(double)));
    CUDA_TRY(cudaMalloc(&g->ftab, cap * sizeof(ull)));
    CUDA_TRY(cudaMalloc(&g->fscan, cap * sizeof(ull)));
    CUDA_TRY(cudaMalloc(&g->eflag, cap * sizeof(unsigned)));
    CUDA_TRY(cudaMalloc(&g->epos, cap * sizeof(unsigned)));
    size_t t1 = 0, t2 = 0;
    CUDA_TRY(cub::DeviceScan::InclusiveScan(nullptr, t1, g->ftab, g->fscan, SkipCompose(), (int)cap, dp->stream));
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, t2, g->eflag, g->epos, (int)cap, dp->stream));
    g->cub_tmp_bytes = std::max(t1, t2);
    CUDA_TRY(cudaMalloc(&g->cub_tmp, g->cub_tmp_bytes));
    g->pairs_cap = cap;
    return 0;
}

static int mcgen_begin(DevPlan *dp, McGen *g, uint32_t seed, ull pair0 = 0)
{
    if (!g->tables_ready) {
        static double wi[256], fi[256];
        for (int i = 0; i < 256; i++) { memcpy(&wi[i], &hvb200_zig_wi_bits[i], 8); memcpy(&fi[i], &hvb200_zig_fi_bits[i], 8); }
        CUDA_TRY(cudaMemcpyToSymbol(c_zig_ki, hvb200_zig_ki, sizeof(ull) * 256));
        CUDA_TRY(cudaMemcpyToSymbol(c_zig_wi, wi, sizeof wi));
        CUDA_TRY(cudaMemcpyToSymbol(c_zig_fi, fi, sizeof fi));
        g->tables_ready = true;
    }
    if (!g->state) {
        CUDA_TRY(cudaMalloc(&g->state, kMtN * sizeof(unsigned)));
        CUDA_TRY(cudaMalloc(&g->scalars, 4 * sizeof(ull)));
        CUDA_TRY(cudaMallocHost(&g->scalars_host, 4 * sizeof(ull)));
    }
    // mt19937_seed (c/mt19937/mt19937.c:6-16): Knuth LCG fill, then the first draw regenerates the block
    unsigned key[kMtN];
    for (int i = 0; i < kMtN; i++) {
        key[i] = seed;
        seed = 1812433253u * (seed ^ (seed >> 30)) + (unsigned)i + 1u;
    }
    // a stream that starts at pair `pair0` (DZ2019-MC shards that own a pair range, hvb200_plan_mc_segment_scan): the
    // state window 2*pair0 words ahead, by the jump-ahead polynomial t^(2 pair0) mod phi applied on the host
    // (hvb200_mtjump.c: 64 squarings of a degree-19937 polynomial and one 20 560-word convolution, ~0.1 s)
    if (pair0) {
        static thread_local unsigned poly[kMtN], jumped[kMtN];
        if (hvb200_mtjump_poly(2 * pair0, poly) || hvb200_mtjump_apply_host(key, poly, jumped)) return -1;
        memcpy(key, jumped, sizeof key);
    }
    CUDA_TRY(cudaMemcpyAsync(g->state, key, sizeof key, cudaMemcpyHostToDevice, dp->stream));
    CUDA_TRY(cudaStreamSynchronize(dp->stream));          // `key` is on the stack
    g->have_pairs = 0;
    return 0;
}

// ---- parallel MT19937 generation -------------------------------------------------------------------
// A run of blocks is cut into segments of kMtSegBlocks blocks.  The start states of the segments come from
// log2(segments) rounds of jump-ahead: round k maps the states of segments [0, 2^k) to those of segments
// [2^k, 2^(k+1)) with the polynomial of J * 2^k words, J = one segment (k_mt_generate in "state sequence" mode,
// then k_mt_jump).  Then every segment is generated by its own CTA.  Bit-identical to the sequential
// generator by construction (tests: the DZ2019-MC goldens and the host/device cross-checks).
static constexpr int kMtSegBlocks = 256;                    // 159 744 words, ~65 us for one CTA
static constexpr int kMtLevels = 10;                        // up to 1024 segments per pass
struct MtJumpTables {
    unsigned short *pos = nullptr;                          // [kMtLevels][20000] sorted set coefficients
    int *slice_off = nullptr;                               // [kMtLevels][kMtSplits + 1]
    bool ready = false, failed = false;
};
static MtJumpTables g_mtjump[16];
static std::mutex g_mtjump_mutex;
static constexpr int kMtPosStride = 20000;

static int mtjump_tables(int device, MtJumpTables **out)
{
    *out = nullptr;
    if (device < 0 || device >= 16 || getenv("HVB200_MT_SERIAL")) return 0;
    std::lock_guard<std::mutex> lock(g_mtjump_mutex);
    MtJumpTables *T = &g_mtjump[device];
    if (T->failed) return 0;
    if (!T->ready) {
        static std::vector<unsigned short> h_pos;           // host copy, shared by all devices
        static std::vector<int> h_off;
        if (h_pos.empty()) {
            std::vector<unsigned> poly(kMtN), nxt(kMtN);
            if (hvb200_mtjump_poly((uint64_t)kMtSegBlocks * kMtN, poly.data())) { T->failed = true; return 0; }
            h_pos.assign((size_t)kMtLevels * kMtPosStride, 0);
            h_off.assign((size_t)kMtLevels * (kMtSplits + 1), 0);
            for (int k = 0; k < kMtLevels; k++) {
                int cnt = 0;
                int *off = &h_off[(size_t)k * (kMtSplits + 1)];
                for (int sp = 0; sp < kMtSplits; sp++) {
                    off[sp] = cnt;
                    for (int i = sp * kMtSlice; i < (sp + 1) * kMtSlice && i < 19937; i++)
                        if ((poly[i >> 5] >> (i & 31)) & 1u) h_pos[(size_t)k * kMtPosStride + cnt++] = (unsigned short)i;
                }
                off[kMtSplits] = cnt;
                if (k + 1 < kMtLevels) {
                    if (hvb200_mtjump_double(poly.data(), nxt.data())) { h_pos.clear(); T->failed = true; return 0; }
                    poly.swap(nxt);
                }
            }
        }
        CUDA_TRY(cudaMalloc(&T->pos, h_pos.size() * sizeof(unsigned short)));
        CUDA_TRY(cudaMalloc(&T->slice_off, h_off.size() * sizeof(int)));
        CUDA_TRY(cudaMemcpy(T->pos, h_pos.data(), h_pos.size() * sizeof(unsigned short), cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(T->slice_off, h_off.data(), h_off.size() * sizeof(int), cudaMemcpyHostToDevice));
        T->ready = true;
    }
    *out = T;
    return 0;
}

// `nblocks` further blocks of tempered words after g->state into `out`; g->state is advanced.
static int mcgen_generate(struct hvb200_plan *plan, DevPlan *dp, McGen *g, unsigned *out, int nblocks)
{
    MtJumpTables *T = nullptr;
    if (nblocks >= 2 * kMtSegBlocks && mtjump_tables(dp->device, &T)) return -1;
    if (T == nullptr) {
        k_mt_generate<<<1, 256, 0, dp->stream>>>(g->state, out, nblocks, nblocks, 0, g->state, 1, 0);
        CUDA_TRY(cudaGetLastError());
        plan->stats.dirgen_launches++;
        return 0;
    }
    const int max_seg = 1 << kMtLevels;
    while (nblocks > 0) {
        const int pass_blocks = std::min(nblocks, max_seg * kMtSegBlocks);
        const int nseg = (pass_blocks + kMtSegBlocks - 1) / kMtSegBlocks;
        if (g->seg_cap < (size_t)nseg) {
            cudaFree(g->seg_states);
            g->seg_states = nullptr; g->seg_cap = 0;
            CUDA_TRY(cudaMalloc(&g->seg_states, (size_t)nseg * kMtN * sizeof(unsigned)));
            g->seg_cap = (size_t)nseg;
        }
        int max_jumps = 0;
        for (int c = 1; c < nseg; c <<= 1) max_jumps = std::max(max_jumps, std::min(c, nseg - c));
        if (g->xbuf_cap < (size_t)max_jumps) {
            cudaFree(g->xbuf);
            g->xbuf = nullptr; g->xbuf_cap = 0;
            CUDA_TRY(cudaMalloc(&g->xbuf, (size_t)max_jumps * kMtSeqWords * sizeof(unsigned)));
            g->xbuf_cap = (size_t)max_jumps;
        }
        CUDA_TRY(cudaMemcpyAsync(g->seg_states, g->state, kMtN * sizeof(unsigned), cudaMemcpyDeviceToDevice, dp->stream));
        int level = 0;
        for (int c = 1; c < nseg; c <<= 1, level++) {
            const int cnt = std::min(c, nseg - c);
            // state sequences of segments [0, cnt), then their windows c segments ahead
            k_mt_generate<<<cnt, 256, 0, dp->stream>>>(g->seg_states, g->xbuf, cnt * kMtSeqBlocks, kMtSeqBlocks, (size_t)kMtSeqWords, nullptr, 0, 1);
            CUDA_TRY(cudaMemsetAsync(g->seg_states + (size_t)c * kMtN, 0, (size_t)cnt * kMtN * sizeof(unsigned), dp->stream));
            k_mt_jump<<<dim3((unsigned)cnt, kMtSplits), 256, 0, dp->stream>>>(g->xbuf, T->pos + (size_t)level * kMtPosStride,
                                                                              T->slice_off + (size_t)level * (kMtSplits + 1),
                                                                              g->seg_states + (size_t)c * kMtN);
            plan->stats.dirgen_launches += 2;
        }
        k_mt_generate<<<nseg, 256, 0, dp->stream>>>(g->seg_states, out, pass_blocks, kMtSegBlocks, (size_t)kMtSegBlocks * kMtN, g->state, 1, 0);
        CUDA_TRY(cudaGetLastError());
        plan->stats.dirgen_launches++;
        out += (size_t)pass_blocks * kMtN;
        nblocks -= pass_blocks;
    }
    return 0;
}

// Make `target` pairs resident at the head of g->words (generating what is missing), classify every pair and scan
// the skip maps from s = 0 at pair 0 (the head of the buffer is always an attempt start).
static int mcgen_fill_classify(struct hvb200_plan *plan, DevPlan *dp, McGen *g, size_t target)
{
    if (mcgen_reserve(dp, g, target)) return -1;
    if (g->have_pairs < target) {
        const size_t missing_words = 2 * (target - g->have_pairs);
        const int nblocks = (int)((missing_words + kMtN - 1) / kMtN);
        const int b = span_begin(dp, 1);
        const int rc = mcgen_generate(plan, dp, g, g->words + 2 * g->have_pairs, nblocks);
        span_end(dp, b);
        if (rc) return -1;
        g->have_pairs += (size_t)nblocks * kMtN / 2;
    }
    const ull P = g->have_pairs;
    const unsigned grid = (unsigned)((P + 255) / 256);
    const int b = span_begin(dp, 1);
    CUDA_TRY(cudaMemsetAsync(g->scalars, 0, 4 * sizeof(ull), dp->stream));
    // near-tie band of the slow paths' comparisons: 8 ulp (CUDA's and glibc's log1p / exp are each within 1 ulp of the
    // true value, the expressions around them add a few roundings); HVB200_MC_TIE_EPS widens it to exercise the fallback
    double tie_eps = 8.0 * 2.220446049250313e-16;
    if (const char *e = getenv("HVB200_MC_TIE_EPS")) { const double v = atof(e); if (v > 0.0) tie_eps = v; }
    k_zig_classify<<<grid, 256, 0, dp->stream>>>(g->words, P, g->L, g->emit, g->val, g->ftab, g->scalars, tie_eps);
    CUDA_TRY(cub::DeviceScan::InclusiveScan(g->cub_tmp, g->cub_tmp_bytes, g->ftab, g->fscan, SkipCompose(), (int)P, dp->stream));
    k_zig_flags<<<grid, 256, 0, dp->stream>>>(g->fscan, g->emit, P, g->eflag, g->scalars);
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(g->cub_tmp, g->cub_tmp_bytes, g->eflag, g->epos, (int)P, dp->stream));
    span_end(dp, b);
    CUDA_TRY(cudaGetLastError());
    plan->stats.dirgen_launches += 2;
    return 0;
}
// Drop the first `next` pairs of the buffer (they are consumed); the rest moves to the head.
static int mcgen_consume(DevPlan *dp, McGen *g, ull next)
{
    const ull P = g->have_pairs;
    const size_t left = (size_t)(P - next);
    if (left && next) {
        // ranges may overlap when more than half of the buffer is left: go through a bounce copy
        if (next >= left) {
            CUDA_TRY(cudaMemcpyAsync(g->words, g->words + 2 * next, 2 * left * sizeof(unsigned), cudaMemcpyDeviceToDevice, dp->stream));
        } else {
            unsigned *tmp = nullptr;
            CUDA_TRY(cudaMalloc(&tmp, 2 * left * sizeof(unsigned)));
            CUDA_TRY(cudaMemcpyAsync(tmp, g->words + 2 * next, 2 * left * sizeof(unsigned), cudaMemcpyDeviceToDevice, dp->stream));
            CUDA_TRY(cudaMemcpyAsync(g->words, tmp, 2 * left * sizeof(unsigned), cudaMemcpyDeviceToDevice, dp->stream));
            CUDA_TRY(cudaStreamSynchronize(dp->stream));
            cudaFree(tmp);
        }
    }
    g->have_pairs = left;
    return 0;
}
static int mcgen_read_scalars(DevPlan *dp, McGen *g)
{
    CUDA_TRY(cudaMemcpyAsync(g->scalars_host, g->scalars, 4 * sizeof(ull), cudaMemcpyDeviceToHost, dp->stream));
    CUDA_TRY(cudaStreamSynchronize(dp->stream));
    if (g->scalars_host[2] || getenv("HVB200_MC_FORCE_TAIL_OVERFLOW")) {
        if (g->scalars_host[2] & 2ull) hvb200_set_error("ziggurat accept/reject near-tie: the host generator decides");
        else hvb200_set_error("ziggurat tail loop longer than %d pairs (device limit)", kZigMaxL);
        return 2;                       // the caller continues this run with the host generator
    }
    return 0;
}

// Produce the next `need` normals of the stream into `normals_dev` (nullptr = discard).
static int mcgen_normals(struct hvb200_plan *plan, DevPlan *dp, McGen *g, ull need, double *normals_dev)
{
    if (need == 0) return 0;
    if (need > 0x7fff0000ull) { hvb200_set_error("internal: MC batch too large"); return -1; }
    size_t target = (size_t)(need + need / 32 + 2048);           // ~2.2 % of the pairs do not end up as a normal
    for (int attempt = 0; attempt < 8; attempt++) {
        if (mcgen_fill_classify(plan, dp, g, target)) return -1;
        const ull P = g->have_pairs;
        const unsigned grid = (unsigned)((P + 255) / 256);
        const int b = span_begin(dp, 1);
        k_zig_scatter<<<grid, 256, 0, dp->stream>>>(g->eflag, g->epos, g->val, g->L, P, need, normals_dev, g->scalars);
        span_end(dp, b);
        CUDA_TRY(cudaGetLastError());
        plan->stats.dirgen_launches++;
        if (const int rc = mcgen_read_scalars(dp, g)) return rc;
        // the last kZigMaxL pairs may hold attempts whose look-ahead leaves the batch: never end there
        if (g->scalars_host[1] >= need && g->scalars_host[0] + 2 * kZigMaxL <= P)
            return mcgen_consume(dp, g, g->scalars_host[0]);     // keep the pairs behind the last used attempt
        target += target / 16 + 4096;       // not enough accepted draws: extend the batch and redo
    }
    hvb200_set_error("internal: MC generator could not produce %llu normals", need);
    return -1;
}

// ---- pair-addressed segments (DZ2019-MC shards that do not parse the stream prefix) ---------------------------------
// A sample's position in the MT19937 stream is data dependent (c/rng.c:71-99: every rejected draw shifts what follows),
// so a shard that starts at sample i0 > 0 had to parse the whole prefix.  Instead every rank owns a PAIR range
// [pair0, pair0 + npairs) of the stream, reached by jump-ahead, and the ranks exchange what the parse of a range needs
// from its predecessors: the skip state s_in (how many pairs at the head of the range still belong to an attempt that
// started before it) and the number of normals emitted before it.  The parses for the 16 possible s_in coincide from the
// first pair at which all of them start an attempt (a few tens of pairs into the range: 98 % of the attempts consume one
// pair), so a range is described by
//     merge        the first such pair (relative to pair0),
//     head[s]      the normals emitted by the attempts that start in [s, merge) when s_in = s        (host walk over the
//                  device's own classification of the head, so that both passes take the same accept/reject decisions),
//     body         the normals emitted by the attempts that start in [merge, npairs)                  (one device pass),
//     skip_out     the pairs of the NEXT range that the last attempt of this one consumes.
// mcgen_scan_segment computes them; hvb200_mc_segments_resolve (host C) turns the gathered descriptions into each
// rank's first pair, the normals it discards to reach a sample boundary, and its sample range.
static constexpr int kMcHead = 4096;
static int mcgen_scan_segment(struct hvb200_plan *plan, DevPlan *dp, McGen *g, ull npairs, int want_body, ull out[20])
{
    // out: [0] merge, [1..16] head[s], [17] body, [18] skip_out, [19] reserved
    if (npairs != ~0ull && npairs < 4 * (ull)kMcHead) { hvb200_set_error("MC segment too short (%llu pairs)", npairs); return -1; }
    const size_t batch = (size_t)1 << 26;
    size_t target = 2 * (size_t)kMcHead;
    if (mcgen_fill_classify(plan, dp, g, target)) return -1;
    static thread_local unsigned char hL[kMcHead], hE[kMcHead];
    CUDA_TRY(cudaMemcpyAsync(hL, g->L, kMcHead, cudaMemcpyDeviceToHost, dp->stream));
    CUDA_TRY(cudaMemcpyAsync(hE, g->emit, kMcHead, cudaMemcpyDeviceToHost, dp->stream));
    CUDA_TRY(cudaStreamSynchronize(dp->stream));
    int pos[16];
    ull cnt[16];
    for (int s = 0; s < 16; s++) { pos[s] = s; cnt[s] = 0; }
    for (;;) {
        int lo = 0, hi = 0;
        for (int s = 1; s < 16; s++) { if (pos[s] < pos[lo]) lo = s; if (pos[s] > pos[hi]) hi = s; }
        if (pos[lo] == pos[hi]) break;
        const int p = pos[lo];
        if (p >= kMcHead - kZigMaxL || hL[p] == 0) { hvb200_set_error("MC segment: the 16 parses did not merge within %d pairs", kMcHead); return 3; }
        if (hE[p] >= 2) { hvb200_set_error("ziggurat tail loop beyond the device tables or accept/reject near-tie"); return 2; }
        // every parse standing at p takes the same step
        for (int s = 0; s < 16; s++) if (pos[s] == p) { pos[s] = p + hL[p]; cnt[s] += hE[p] == 1; }
    }
    const ull merge = (ull)pos[0];
    out[0] = merge;
    for (int s = 0; s < 16; s++) out[1 + s] = cnt[s];
    out[17] = out[18] = out[19] = 0;
    if (npairs == ~0ull || !want_body) return 0;
    // body: attempts starting in [merge, npairs), batch by batch; each batch begins at an attempt start
    if (mcgen_consume(dp, g, merge)) return -1;
    ull remaining = npairs - merge, body = 0;
    for (;;) {
        target = (size_t)std::min<ull>(batch, remaining + 64);
        if (mcgen_fill_classify(plan, dp, g, target)) return -1;
        const ull P = g->have_pairs;
        const ull limit = remaining + 2 * kZigMaxL <= P ? remaining : P - 2 * kZigMaxL;
        const int b = span_begin(dp, 1);
        k_zig_range_end<<<(unsigned)((P + 255) / 256), 256, 0, dp->stream>>>(g->fscan, g->eflag, g->epos, g->L, P, limit, g->scalars);
        span_end(dp, b);
        CUDA_TRY(cudaGetLastError());
        plan->stats.dirgen_launches++;
        if (const int rc = mcgen_read_scalars(dp, g)) return rc;
        const ull next = g->scalars_host[0];
        if (next < limit || next > limit + kZigMaxL) { hvb200_set_error("internal: MC segment scan lost its range end"); return -1; }
        body += g->scalars_host[1];
        if (next >= remaining) { out[17] = body; out[18] = next - remaining; return 0; }
        if (mcgen_consume(dp, g, next)) return -1;
        remaining -= next;
    }
}

static void mcgen_free(McGen *g)
{
    cudaFree(g->state); cudaFree(g->words); cudaFree(g->L); cudaFree(g->emit); cudaFree(g->val); cudaFree(g->ftab);
    cudaFree(g->fscan); cudaFree(g->eflag); cudaFree(g->epos); cudaFree(g->cub_tmp); cudaFree(g->scalars);
    cudaFree(g->seg_states); cudaFree(g->xbuf);
    if (g->scalars_host) cudaFreeHost(g->scalars_host);
    *g = McGen();
}
