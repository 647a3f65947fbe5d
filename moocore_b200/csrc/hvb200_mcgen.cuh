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
// Generates `nblocks` further blocks of 624 tempered words after the current state.
// x[n] = x[n-227] ^ twist(x[n-624], x[n-623]): 227 consecutive words are independent, so a block is
// three dependent phases.  Old and new blocks live in two shared arrays (ping-pong), which leaves
// one barrier per phase; tempering and the global store of a block overlap the next block's phases.
__global__ void __launch_bounds__(256) k_mt_generate(unsigned *__restrict__ state, unsigned *__restrict__ out, int nblocks)
{
    __shared__ unsigned xs[2][kMtN];
    const int t = threadIdx.x;
    for (int i = t; i < kMtN; i += 256) xs[0][i] = state[i];
    __syncthreads();
    int cur = 0;
    for (int b = 0; b < nblocks; b++) {
        const unsigned *xo = xs[cur];
        unsigned *xn = xs[cur ^ 1];
        unsigned *o = out + (size_t)b * kMtN;
        // phase 1: i in [0, 227)      old x[i], x[i+1], x[i+397]
        if (t < 227) {
            const unsigned v = mt_twist(xo[t], xo[t + 1], xo[t + kMtM]);
            xn[t] = v;
            o[t] = mt_temper(v);
        }
        __syncthreads();
        // phase 2: i in [227, 454)    old x[i], x[i+1], new x[i-227]
        if (t < 227) {
            const int i = 227 + t;
            const unsigned v = mt_twist(xo[i], xo[i + 1], xn[i - 227]);
            xn[i] = v;
            o[i] = mt_temper(v);
        }
        __syncthreads();
        // phase 3: i in [454, 624)    old x[i], x[i+1] (new x[0] for i = 623), new x[i-227]
        if (t < kMtN - 454) {
            const int i = 454 + t;
            const unsigned nxt = (i + 1 == kMtN) ? xn[0] : xo[i + 1];
            const unsigned v = mt_twist(xo[i], nxt, xn[i - 227]);
            xn[i] = v;
            o[i] = mt_temper(v);
        }
        __syncthreads();
        cur ^= 1;
    }
    for (int i = t; i < kMtN; i += 256) state[i] = xs[cur][i];
}

// ---- ziggurat, one hypothetical attempt per pair -------------------------------------------------
__device__ __forceinline__ double pair_to_unit(unsigned w0, unsigned w1)
{
    const int a = (int)(w0 >> 5), b = (int)(w1 >> 6);
    return (a * 67108864.0 + b) / 9007199254740992.0;
}
__global__ void __launch_bounds__(256) k_zig_classify(const unsigned *__restrict__ words, ull npairs,
                                                      unsigned char *__restrict__ Lout, unsigned char *__restrict__ emit,
                                                      double *__restrict__ val, ull *__restrict__ ftab, ull *scalars)
{
    const ull j = (ull)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= npairs) return;
    const double R = __longlong_as_double((long long)HVB200_ZIG_R_BITS);
    const double INV_R = __longlong_as_double((long long)HVB200_ZIG_INV_R_BITS);
    const ull r = ((ull)words[2 * j] << 32) | words[2 * j + 1];
    const int layer = (int)(r & 0xff);
    const ull mag = (r >> 9) & 0x000fffffffffffffull;
    double x = (double)mag * c_zig_wi[layer];
    if ((r >> 8) & 1) x = -x;
    int L = 1, em = 1;
    double v = x;
    if (!(mag < c_zig_ki[layer])) {
        if (layer == 0) {               // tail: pairs j+1, j+2, then j+3, j+4, ...
            L = 0;
            for (int t = 0; 1 + 2 * (t + 1) <= kZigMaxL; t++) {
                const ull a = j + 1 + 2 * (ull)t;
                if (a + 1 >= npairs) { L = 0; em = 0; break; }          // needs pairs beyond this batch
                const double xx = -INV_R * log1p(-pair_to_unit(words[2 * a], words[2 * a + 1]));
                const double yy = -log1p(-pair_to_unit(words[2 * a + 2], words[2 * a + 3]));
                if (yy + yy > xx * xx) {
                    v = ((mag >> 8) & 1) ? -(R + xx) : R + xx;
                    L = 1 + 2 * (t + 1);
                    break;
                }
                if (1 + 2 * (t + 2) > kZigMaxL) { atomicOr(&scalars[2], 1ull); L = kZigMaxL; em = 0; break; }
            }
        } else {                        // wedge: pair j+1
            if (j + 1 >= npairs) { L = 0; em = 0; }
            else {
                const double u = pair_to_unit(words[2 * j + 2], words[2 * j + 3]);
                const double f_hi = c_zig_fi[layer - 1], f_lo = c_zig_fi[layer];
                L = 2;
                em = ((f_hi - f_lo) * u + f_lo < exp(-0.5 * x * x)) ? 1 : 0;
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
                                                   ull npairs, unsigned *__restrict__ eflag)
{
    const ull j = (ull)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= npairs) return;
    const bool start = (j == 0) || ((fscan[j - 1] & 15ull) == 0);      // skip counter before pair j, from s = 0
    eflag[j] = (start && emit[j]) ? 1u : 0u;
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

static int mcgen_begin(DevPlan *dp, McGen *g, uint32_t seed)
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
    CUDA_TRY(cudaMemcpyAsync(g->state, key, sizeof key, cudaMemcpyHostToDevice, dp->stream));
    CUDA_TRY(cudaStreamSynchronize(dp->stream));          // `key` is on the stack
    g->have_pairs = 0;
    return 0;
}

// Produce the next `need` normals of the stream into `normals_dev` (nullptr = discard).
static int mcgen_normals(struct hvb200_plan *plan, DevPlan *dp, McGen *g, ull need, double *normals_dev)
{
    if (need == 0) return 0;
    if (need > 0x7fff0000ull) { hvb200_set_error("internal: MC batch too large"); return -1; }
    size_t target = (size_t)(need + need / 32 + 2048);           // ~1.4 % of draws are rejected or consume extra pairs
    for (int attempt = 0; attempt < 8; attempt++) {
        if (mcgen_reserve(dp, g, target)) return -1;
        if (g->have_pairs < target) {
            const size_t missing_words = 2 * (target - g->have_pairs);
            const int nblocks = (int)((missing_words + kMtN - 1) / kMtN);
            const int b = span_begin(dp, 1);
            k_mt_generate<<<1, 256, 0, dp->stream>>>(g->state, g->words + 2 * g->have_pairs, nblocks);
            span_end(dp, b);
            CUDA_TRY(cudaGetLastError());
            g->have_pairs += (size_t)nblocks * kMtN / 2;
            plan->stats.dirgen_launches++;
        }
        const ull P = g->have_pairs;
        const unsigned grid = (unsigned)((P + 255) / 256);
        const int b = span_begin(dp, 1);
        CUDA_TRY(cudaMemsetAsync(g->scalars, 0, 4 * sizeof(ull), dp->stream));
        k_zig_classify<<<grid, 256, 0, dp->stream>>>(g->words, P, g->L, g->emit, g->val, g->ftab, g->scalars);
        CUDA_TRY(cub::DeviceScan::InclusiveScan(g->cub_tmp, g->cub_tmp_bytes, g->ftab, g->fscan, SkipCompose(), (int)P, dp->stream));
        k_zig_flags<<<grid, 256, 0, dp->stream>>>(g->fscan, g->emit, P, g->eflag);
        CUDA_TRY(cub::DeviceScan::ExclusiveSum(g->cub_tmp, g->cub_tmp_bytes, g->eflag, g->epos, (int)P, dp->stream));
        k_zig_scatter<<<grid, 256, 0, dp->stream>>>(g->eflag, g->epos, g->val, g->L, P, need, normals_dev, g->scalars);
        span_end(dp, b);
        CUDA_TRY(cudaGetLastError());
        plan->stats.dirgen_launches += 3;
        CUDA_TRY(cudaMemcpyAsync(g->scalars_host, g->scalars, 4 * sizeof(ull), cudaMemcpyDeviceToHost, dp->stream));
        CUDA_TRY(cudaStreamSynchronize(dp->stream));
        if (g->scalars_host[2]) { hvb200_set_error("ziggurat tail loop longer than %d pairs (device limit)", kZigMaxL); return -1; }
        // the last kZigMaxL pairs may hold attempts whose look-ahead leaves the batch: never end there
        if (g->scalars_host[1] >= need && g->scalars_host[0] + 2 * kZigMaxL <= P) {
            // keep the pairs behind the last used attempt for the next batch
            const ull next = g->scalars_host[0];
            const size_t left = (size_t)(P - next);
            if (left) {
                // ranges may overlap when more than half of the buffer is left: go through a bounce copy
                if (2 * next >= 2 * left) {
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
        target += target / 16 + 4096;       // not enough accepted draws: extend the batch and redo
    }
    hvb200_set_error("internal: MC generator could not produce %llu normals", need);
    return -1;
}

static void mcgen_free(McGen *g)
{
    cudaFree(g->state); cudaFree(g->words); cudaFree(g->L); cudaFree(g->emit); cudaFree(g->val); cudaFree(g->ftab);
    cudaFree(g->fscan); cudaFree(g->eflag); cudaFree(g->epos); cudaFree(g->cub_tmp); cudaFree(g->scalars);
    if (g->scalars_host) cudaFreeHost(g->scalars_host);
    *g = McGen();
}
