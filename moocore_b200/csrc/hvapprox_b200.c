/* hvapprox_b200.c — host side (C99) of the B200 hv_approx path.
 *
 * Exports the three entry points of moocore's c/hvapprox.h:16-38 plus the extended API of
 * include/hvapprox_b200.h.  This file does what c/hvapprox.c does on the host outside its hot
 * loop — transform_and_filter (:30-66), construct_polar_a (:261-286), the final long double
 * scaling (:256-257, 691-692, 863-864) — and produces the two inherently sequential input
 * streams (MT19937+ziggurat normals for DZ2019-MC, the rounded Rphi recurrence) that are then
 * uploaded; every per-sample x per-point operation runs in the CUDA kernels of
 * hvb200_kernels.cu.  There is no CPU implementation of the max-min path here: if CUDA is not
 * usable the entry points report the error and return NaN.
 *
 * All file:line citations are relative to /root/reference.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <math.h>
#include <pthread.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "hvb200_tables.h"
#include "hvb200_internal.h"

#define TINY_WEIGHT 1e-20 /* ALMOST_ZERO_WEIGHT, c/hvapprox.c:9 */

/* ------------------------------------------------------------------------------------------ */
/* errors                                                                                     */
/* ------------------------------------------------------------------------------------------ */
static __thread char tls_error[512];

void hvb200_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(tls_error, sizeof tls_error, fmt, ap);
    va_end(ap);
}
const char *hvb200_last_error(void) { return tls_error; }
const char *hvb200_version(void) { return "moocore_b200 0.1 (hv_approx path of moocore 0.20, sm_100a)"; }
int hvb200_device_count(void) { return hvb200cu_device_count(); }

static double fail_nan(const char *where)
{
    fprintf(stderr, "moocore_b200: %s: %s\n", where, tls_error[0] ? tls_error : "unknown error");
    return NAN;
}

/* ------------------------------------------------------------------------------------------ */
/* transform_and_filter (c/hvapprox.c:30-66)                                                  */
/* ------------------------------------------------------------------------------------------ */
static size_t transform_and_filter(const double *data, size_t n, int d, const double *ref,
                                   const boolvec *maximise, double *out)
{
    size_t kept = 0;
    for (size_t i = 0; i < n; i++) {
        const double *src = data + i * (size_t)d;
        double *dst = out + kept * (size_t)d;
        int k;
        for (k = 0; k < d; k++) {
            const double v = maximise[k] ? src[k] - ref[k] : ref[k] - src[k];
            if (!(v > 0)) break;          /* v <= 0 (or NaN) : does not strictly dominate ref */
            dst[k] = v;
        }
        kept += (k == d);
    }
    return kept;
}

/* ------------------------------------------------------------------------------------------ */
/* plans                                                                                      */
/* ------------------------------------------------------------------------------------------ */
static int plan_create_shared(hvb200_plan **plan_out, const double *data, size_t npoints, dimension_t nobjs,
                              const double *ref, const boolvec *maximise, int device, void *share);
int hvb200_plan_create(hvb200_plan **plan_out, const double *data, size_t npoints, dimension_t nobjs,
                       const double *ref, const boolvec *maximise, int device)
{
    return plan_create_shared(plan_out, data, npoints, nobjs, ref, maximise, device, NULL);
}
static int plan_create_shared(hvb200_plan **plan_out, const double *data, size_t npoints, dimension_t nobjs,
                              const double *ref, const boolvec *maximise, int device, void *share)
{
    *plan_out = NULL;
    tls_error[0] = 0;
    const int d = (int)nobjs;
    if (d < 2 || d > HVB200_DIMENSION_MAX) {
        hvb200_set_error("nobjs = %d outside [2, %d]", d, HVB200_DIMENSION_MAX);
        return -2;
    }
    if (device < 0 || device >= hvb200cu_device_count()) {
        hvb200_set_error("CUDA device %d not available (%d visible); this library has no CPU fallback",
                         device, hvb200cu_device_count());
        return -3;
    }
    hvb200_plan *plan = calloc(1, sizeof *plan);
    if (!plan) { hvb200_set_error("out of host memory"); return -4; }
    plan->device = device;
    plan->d = d;
    plan->share = share;
    plan->kernel = HVB200_KERNEL_AUTO;
    const char *env = getenv("HVB200_KERNEL");
    if (env && !strcmp(env, "exact")) plan->kernel = HVB200_KERNEL_EXACT;
    if (env && !strcmp(env, "screen")) plan->kernel = HVB200_KERNEL_SCREEN;
    if (env && !strcmp(env, "dpx")) plan->kernel = HVB200_KERNEL_DPX;
    if (env && !strcmp(env, "pruned")) plan->kernel = HVB200_KERNEL_BP;
    /* transform_and_filter (c/hvapprox.c:30-66): on the device for large inputs (the reference flags its host
       loop as slow at 1M points, :35-42), on the host below 2^18 coordinates where a kernel launch, a scan and
       two synchronisations cost more than the loop.  HVB200_TRANSFORM=host|device forces one.  Same bits. */
    const char *tf = getenv("HVB200_TRANSFORM");
    int on_device = (uint64_t)npoints * (uint64_t)d >= (1u << 18);
    if (tf && !strcmp(tf, "host")) on_device = 0;
    if (tf && !strcmp(tf, "device")) on_device = 1;
    int rc;
    if (on_device) {
        unsigned char mx[HVB200_MAXD];
        for (int k = 0; k < d; k++) mx[k] = maximise[k] ? 1 : 0;
        size_t n = 0;
        rc = hvb200cu_plan_upload_raw(plan, data, npoints, ref, mx, &n);
        if (rc == 0 && n == 0) { if (plan->dev) hvb200cu_plan_free(plan); free(plan); return 0; }   /* caller returns 0.0, c/hvapprox.c:228-229 */
    } else {
        double *pts = malloc((npoints ? npoints : 1) * (size_t)d * sizeof *pts);
        if (!pts) { free(plan); hvb200_set_error("out of host memory"); return -4; }
        const size_t n = transform_and_filter(data, npoints, d, ref, maximise, pts);
        if (n == 0) { free(pts); free(plan); return 0; }     /* caller returns 0.0, c/hvapprox.c:228-229 */
        plan->n = n;
        rc = hvb200cu_plan_upload(plan, pts);
        free(pts);
    }
    if (rc) { if (plan->dev) hvb200cu_plan_free(plan); free(plan); return -1; }
    *plan_out = plan;
    return 0;
}
void hvb200_plan_destroy(hvb200_plan *plan)
{
    if (!plan) return;
    hvb200cu_plan_free(plan);
    free(plan);
}
size_t hvb200_plan_npoints(const hvb200_plan *plan) { return plan ? plan->n : 0; }
size_t hvb200_plan_ndominated(const hvb200_plan *plan) { return plan ? plan->n_dominated : 0; }
int hvb200_plan_set_kernel(hvb200_plan *plan, int kernel)
{
    if (!plan) { hvb200_set_error("NULL plan (no point strictly dominates ref)"); return -2; }
    if (kernel < HVB200_KERNEL_AUTO || kernel > HVB200_KERNEL_BP) { hvb200_set_error("bad kernel id %d", kernel); return -2; }
    plan->kernel = kernel;
    return 0;
}
int hvb200_plan_last_stats(const hvb200_plan *plan, hvb200_stats *out)
{
    if (!plan) { memset(out, 0, sizeof *out); hvb200_set_error("NULL plan (no point strictly dominates ref)"); return -2; }
    *out = plan->stats;
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* DZ2019-HW host part: construct_polar_a (c/hvapprox.c:261-286)                              */
/* ------------------------------------------------------------------------------------------ */
static int lattice_prime(int ldim)
{   /* smallest prime p with (p-1)/2 >= ldim; the reference tabulates it (:266-270) */
    for (int p = 3;; p += 2) {
        int prime = 1;
        for (int q = 3; q * q <= p; q += 2) if (p % q == 0) { prime = 0; break; }
        if (prime && (p - 1) / 2 >= ldim) return p;
    }
}
static void hw_params_init(hvb200_hw_params *hw, int d, uint64_t nsamples)
{
    static const long double PI_L = 3.141592653589793238462643383279502884L;
    memset(hw, 0, sizeof *hw);
    hw->d = d;
    hw->nsamples = nsamples;
    const int ldim = d - 1;
    const int p = lattice_prime(ldim);
    hw->a[0] = 1;
    for (int k = 1; k < ldim; k++) {
        long double t = fabsl(2 * cosl(2 * PI_L * k / p));
        t -= truncl(t);
        hw->a[k] = (uint64_t)llroundl(nsamples * t);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* DZ2019-MC host part: the NumPy-legacy MT19937 stream + 256-layer ziggurat                   */
/* (c/mt19937/mt19937.c:6-55, c/mt19937/mt19937.h:28-57, c/rng.c:71-99)                        */
/* ------------------------------------------------------------------------------------------ */
typedef struct { uint32_t x[624]; int pos; } mt_stream;

static void mt_seed(mt_stream *s, uint32_t seed)
{
    for (int i = 0; i < 624; i++) {
        s->x[i] = seed;
        seed = 1812433253u * (seed ^ (seed >> 30)) + (uint32_t)i + 1u;
    }
    s->pos = 624;
}
static void mt_twist(mt_stream *s)
{
    uint32_t *x = s->x;
    int i;
    for (i = 0; i < 624 - 397; i++) {
        const uint32_t y = (x[i] & 0x80000000u) | (x[i + 1] & 0x7fffffffu);
        x[i] = x[i + 397] ^ (y >> 1) ^ (-(y & 1u) & 0x9908b0dfu);
    }
    for (; i < 623; i++) {
        const uint32_t y = (x[i] & 0x80000000u) | (x[i + 1] & 0x7fffffffu);
        x[i] = x[i - 227] ^ (y >> 1) ^ (-(y & 1u) & 0x9908b0dfu);
    }
    const uint32_t y = (x[623] & 0x80000000u) | (x[0] & 0x7fffffffu);
    x[623] = x[396] ^ (y >> 1) ^ (-(y & 1u) & 0x9908b0dfu);
    s->pos = 0;
}
static inline uint32_t mt_u32(mt_stream *s)
{
    if (s->pos == 624) mt_twist(s);
    uint32_t y = s->x[s->pos++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}
static inline uint64_t mt_u64(mt_stream *s) { const uint64_t hi = mt_u32(s); return hi << 32 | mt_u32(s); }
static inline double mt_unit(mt_stream *s)
{
    const int32_t a = (int32_t)(mt_u32(s) >> 5), b = (int32_t)(mt_u32(s) >> 6);
    return (a * 67108864.0 + b) / 9007199254740992.0;
}
static inline double from_bits(uint64_t b) { double v; memcpy(&v, &b, sizeof v); return v; }

static double zig_normal(mt_stream *s)
{
    const double R = from_bits(HVB200_ZIG_R_BITS), INV_R = from_bits(HVB200_ZIG_INV_R_BITS);
    for (;;) {
        const uint64_t r = mt_u64(s);
        const int layer = (int)(r & 0xff);
        const uint64_t mag = (r >> 9) & 0x000fffffffffffffull;
        double x = (double)mag * from_bits(hvb200_zig_wi_bits[layer]);
        if ((r >> 8) & 1) x = -x;
        if (mag < hvb200_zig_ki[layer]) return x;          /* ~99.3 % */
        if (layer == 0) {                                    /* tail */
            for (;;) {
                const double xx = -INV_R * log1p(-mt_unit(s));
                const double yy = -log1p(-mt_unit(s));
                if (yy + yy > xx * xx) return ((mag >> 8) & 1) ? -(R + xx) : R + xx;
            }
        } else {                                             /* wedge */
            const double f_hi = from_bits(hvb200_zig_fi_bits[layer - 1]), f_lo = from_bits(hvb200_zig_fi_bits[layer]);
            if ((f_hi - f_lo) * mt_unit(s) + f_lo < exp(-0.5 * x * x)) return x;
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Rphi-FWE+ host part: Rphi_init / Rphi_next (c/hvapprox.c:695-747) in native long double     */
/* ------------------------------------------------------------------------------------------ */
static void rphi_alpha(long double *alpha, int ldim)
{
    alpha[0] = hvb200_rho[ldim];
    for (int k = 1; k < ldim; k++) alpha[k] = alpha[k - 1] * hvb200_rho[ldim];
}

/* Rphi_next (c/hvapprox.c:741-747): u <- (double) fractl(u + alpha), a sequential recurrence in x87
 * long double that is rounded to double at every step.  It is carried in 64-bit fixed point here:
 *   - alpha = rho^(k+1) >= rho^m = phi/(phi+1) > 0.5, so alpha = M * 2^-64 with M its 64-bit mantissa;
 *   - every state is a multiple of 2^-64 (0.5 is; an x87 sum of two such numbers rounds to a multiple of
 *     2^-64 or 2^-63; rounding to 53 bits only coarsens the grid), so U = u * 2^64 is an integer;
 *   - S = U + M exactly (65 bits).  S < 2^64: the x87 sum is exact and frac = S.  S >= 2^64: the sum
 *     has 65 significant bits and is rounded to nearest-even at bit 0, frac = sum - 1;
 *   - u' = RN53(frac): round the integer to 53 significant bits, ties to even.  A round-up to 2^64
 *     (u' == 1.0) is kept as the 65-bit value so that the next sum is still exact.
 * Checked bit-for-bit against the native x87 recurrence (tests/test_abi_cpu.py); ~3x faster than x87
 * per step and free of libm calls. */
static inline unsigned __int128 rphi_step_fixed(unsigned __int128 U, uint64_t M)
{
    const unsigned __int128 S = U + M;                       /* U <= 2^64, M < 2^64 */
    uint64_t F;
    if ((uint64_t)(S >> 64)) {
        unsigned __int128 q = S >> 1;                        /* sum = S * 2^-64 in [1, 2.x): 64-bit mantissa = S >> 1 */
        if ((uint64_t)S & 1u) q += (uint64_t)q & 1u;         /* ties to even */
        /* q * 2^-63 in [1, 2]: frac.  q == 2^64 (sum rounded to 2.0) gives frac 0; U = 2^64 with M >= 2^63
           gives sums in [1.5, 2) as well, so one subtraction of 2^63 is always enough below 2^64 */
        if ((uint64_t)(q >> 64)) return 0;
        F = ((uint64_t)q - ((uint64_t)1 << 63)) << 1;
    } else {
        F = (uint64_t)S;
    }
    if (F >> 53) {                                           /* more than 53 significant bits: RNE */
        const int r = 11 - __builtin_clzll(F);               /* bits to drop, 1..11 */
        const uint64_t unit = (uint64_t)1 << r, half = unit >> 1, low = F & (unit - 1);
        uint64_t hi = F - low;
        if (low > half || (low == half && (hi & unit))) {
            const uint64_t up = hi + unit;
            if (up < hi) return (unsigned __int128)1 << 64;  /* rounded up to 1.0 */
            hi = up;
        }
        F = hi;
    }
    return F;
}
static void rphi_mantissas(uint64_t *M, int ldim)
{
    long double alpha[HVB200_MAXD];
    rphi_alpha(alpha, ldim);
    for (int k = 0; k < ldim; k++) memcpy(&M[k], &alpha[k], 8);       /* x87: low 8 bytes = explicit 64-bit mantissa */
}
static inline double rphi_to_double(unsigned __int128 U)
{
    return (uint64_t)(U >> 64) ? 1.0 : (double)(uint64_t)U * 0x1p-64;   /* exact: U has <= 53 significant bits */
}
typedef struct { unsigned __int128 U; uint64_t M; uint64_t steps; double *out; size_t stride; } rphi_job;
static void *rphi_worker(void *arg)
{
    rphi_job *j = arg;
    unsigned __int128 U = j->U;
    const uint64_t M = j->M;
    if (j->out) {
        double *o = j->out;
        for (uint64_t i = 0; i < j->steps; i++, o += j->stride) { U = rphi_step_fixed(U, M); *o = rphi_to_double(U); }
    } else {
        for (uint64_t i = 0; i < j->steps; i++) U = rphi_step_fixed(U, M);
    }
    j->U = U;
    return NULL;
}
/* Advance all ldim sequences by `steps`; out (optional) receives u_k at step i in out[k*steps + i]
 * (one contiguous run per objective: no cache line is shared between the writer threads).  The
 * sequences of different objectives are independent: one host thread each when there is enough work. */
static void rphi_advance(unsigned __int128 *u, const uint64_t *M, int ldim, uint64_t steps, double *out)
{
    rphi_job jobs[HVB200_MAXD];
    pthread_t th[HVB200_MAXD];
    for (int k = 0; k < ldim; k++) {
        jobs[k].U = u[k]; jobs[k].M = M[k]; jobs[k].steps = steps;
        jobs[k].out = out ? out + (size_t)k * steps : NULL; jobs[k].stride = 1;
    }
    const int threaded = steps * (uint64_t)ldim >= 65536 && ldim > 1;
    int started[HVB200_MAXD] = {0};
    for (int k = 0; k < ldim; k++) {
        if (threaded && pthread_create(&th[k], NULL, rphi_worker, &jobs[k]) == 0) started[k] = 1;
        else rphi_worker(&jobs[k]);
    }
    for (int k = 0; k < ldim; k++) { if (started[k]) pthread_join(th[k], NULL); u[k] = jobs[k].U; }
}
/* Host-only debug/parity entry: the u-sequence of Rphi-FWE+ for samples [i0, i0+count). */
int hvb200_rphi_sequence(dimension_t nobjs, uint64_t i0, uint64_t count, double *u_out)
{
    const int ldim = (int)nobjs - 1;
    if (ldim < 1 || ldim > HVB200_MAXM) { hvb200_set_error("nobjs outside [2, %d]", HVB200_MAXM + 1); return -2; }
    uint64_t M[HVB200_MAXD];
    unsigned __int128 u[HVB200_MAXD];
    rphi_mantissas(M, ldim);
    for (int k = 0; k < ldim; k++) u[k] = (unsigned __int128)1 << 63;       /* seed 0.5, c/hvapprox.c:847-849 */
    rphi_advance(u, M, ldim, i0, NULL);
    double *soa = malloc((count ? count : 1) * (size_t)ldim * sizeof *soa);
    if (!soa) { hvb200_set_error("out of host memory"); return -4; }
    rphi_advance(u, M, ldim, count, soa);
    for (uint64_t i = 0; i < count; i++)
        for (int k = 0; k < ldim; k++) u_out[i * (size_t)ldim + k] = soa[(size_t)k * count + i];
    free(soa);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Rphi-FWE+ direction cache                                                                   */
/* ------------------------------------------------------------------------------------------ */
/* The Rphi-FWE+ directions depend on the number of objectives and the sample index only — not on the points,
   the reference point or nsamples (c/hvapprox.c:834-865: the recurrence starts at 0.5 for every call).  The
   host recurrence costs 5 ns per step and objective and made the reference's DEFAULT method 10-16x slower than
   DZ2019-HW at the default 262144 samples (2.2 ms at d=3, 9.5 ms at d=20 against 0.14 / 0.58 ms).  The prefix
   [0, valid) of the direction set is therefore kept on the device per (device, d), extended on demand from
   the saved recurrence state, and shared by all later calls.  Bounded (1 GiB per entry, 8 entries, grown by doubling); a run
   that would not fit streams its directions as before.  HVB200_NO_DIRCACHE=1 disables it. */
#define RPHI_CACHE_ENTRIES 8
#define RPHI_CACHE_MAX_BYTES ((uint64_t)1 << 30)
/* a cached direction buffer is shared by the cache entry and by the runs that read it: whoever drops the last
   reference frees it (all under g_rphi_mutex; a run drops its reference after hvb200cu_run_end, which is synchronous) */
typedef struct rphi_buf { double *dev; int device; int refs; } rphi_buf;
typedef struct {
    int used, device, d;
    rphi_buf *buf;                      /* dev: [cap][2][d] */
    uint64_t cap, valid;
    unsigned __int128 u[HVB200_MAXD];   /* recurrence state after `valid` steps */
    uint64_t stamp;
} rphi_cache_t;
static rphi_cache_t g_rphi_cache[RPHI_CACHE_ENTRIES];
static uint64_t g_rphi_stamp;
static pthread_mutex_t g_rphi_mutex = PTHREAD_MUTEX_INITIALIZER;
static void rphi_buf_unref_locked(rphi_buf *b)
{
    if (b && --b->refs == 0) { hvb200cu_dev_free(b->device, b->dev); free(b); }
}

/* ------------------------------------------------------------------------------------------ */
/* run                                                                                         */
/* ------------------------------------------------------------------------------------------ */
static int run_range_ex(hvb200_plan *plan, int method, uint64_t nsamples, uint32_t seed, uint64_t i0, uint64_t i1,
                        void *stream, double out_hilo[2], double *values_out, double *dirs_out, void *sets_ctx);
static int run_range(hvb200_plan *plan, int method, uint64_t nsamples, uint32_t seed, uint64_t i0, uint64_t i1,
                     void *stream, double out_hilo[2], double *values_out, double *dirs_out)
{
    return run_range_ex(plan, method, nsamples, seed, i0, i1, stream, out_hilo, values_out, dirs_out, NULL);
}
/* sets_ctx != NULL: the batch of directions is handed to the batched-sets kernel instead of the plan's own */
static int run_range_ex(hvb200_plan *plan, int method, uint64_t nsamples, uint32_t seed, uint64_t i0, uint64_t i1,
                        void *stream, double out_hilo[2], double *values_out, double *dirs_out, void *sets_ctx)
{
    if (!plan) { hvb200_set_error("NULL plan (no point strictly dominates ref)"); return -2; }
    const int d = plan->d;
    tls_error[0] = 0;
    if (i1 < i0 || i1 > nsamples) { hvb200_set_error("bad sample range [%llu, %llu) of %llu",
                                                     (unsigned long long)i0, (unsigned long long)i1, (unsigned long long)nsamples); return -2; }
    if (method == HVB200_METHOD_HW && d > MOOCORE_HVAPPROX_DIMENSION_MAX + 1) {
        hvb200_set_error("DZ2019-HW supports at most %d objectives", MOOCORE_HVAPPROX_DIMENSION_MAX + 1);
        return -2;
    }
    /* DZ2019-MC stream: generated on the device unless HVB200_MC=host asks for the host generator
       (kept as an independent cross-check of the device generator) */
    const char *mc_env = getenv("HVB200_MC");
    int mc_on_device = !(mc_env && !strcmp(mc_env, "host"));
    /* hvb200_plan_set_mc_start is one-shot: it applies to this run (if it is a device-generated DZ2019-MC run) and to no other */
    const int mc_seek = plan->mc_seek && method == HVB200_METHOD_MC && mc_on_device;
    plan->mc_seek = 0;
    uint64_t batch = 0;
    uint64_t want = i1 - i0 ? i1 - i0 : 1;
    if ((values_out || dirs_out) && want > (1u << 22)) want = 1u << 22;
    if (method != HVB200_METHOD_HW && want > (1u << 20)) want = 1u << 20;     /* MC stream: 1 Mi samples per batch */
    /* host-fed streams: smaller batches so that producing batch b+1 overlaps the GPU work on batch b */
    if ((method == HVB200_METHOD_RPHI || (method == HVB200_METHOD_MC && !mc_on_device)) && want > (1u << 18)) want = 1u << 18;
    plan->run_samples = (values_out || dirs_out) ? 0 : i1 - i0;
    plan->run_total = i1 - i0;
    /* a DZ2019-HW sum needs no direction buffer: the pruned max-min kernel can generate the lattice itself */
    plan->want_fused_hw = method == HVB200_METHOD_HW && !values_out && !dirs_out && !sets_ctx && plan->n > 0 &&
                          d <= MOOCORE_HVAPPROX_DIMENSION_MAX + 1;
    if (hvb200cu_run_begin(plan, stream, want, &batch)) return -1;
    if (hvb200cu_run_is_fused(plan)) {
        hvb200_hw_params hwf;
        hw_params_init(&hwf, d, nsamples);
        if (hvb200cu_maxmin_hw_fused(plan, &hwf, i0, 62, 1, i1 - i0)) return -1;
        if (hvb200cu_run_end(plan, out_hilo)) return -1;
        plan->stats.samples = i1 - i0;
        return 0;
    }

    const int rphi_cached = method == HVB200_METHOD_RPHI && i1 > 0 && !getenv("HVB200_NO_DIRCACHE") &&
                            i1 * (uint64_t)(16 * d) <= RPHI_CACHE_MAX_BYTES;
    hvb200_hw_params hw;
    mt_stream mt;
    uint64_t rphi_M[HVB200_MAXD];
    unsigned __int128 u[HVB200_MAXD];
    if (method == HVB200_METHOD_HW) {
        hw_params_init(&hw, d, nsamples);
    } else if (method == HVB200_METHOD_MC && mc_on_device) {
        /* the stream is sequential: skip the normals of samples [0, i0) */
        int rc;
        if (mc_seek) {
            /* a shard that owns a pair range of the stream (hvb200_plan_set_mc_start): jump there, then drop the
               normals in front of its first whole sample */
            rc = hvb200cu_mc_begin_at(plan, seed, plan->mc_pair);
            if (rc == 0) rc = hvb200cu_mc_skip(plan, plan->mc_discard);
        } else {
            rc = hvb200cu_mc_begin(plan, seed);
            if (rc == 0) rc = hvb200cu_mc_skip(plan, i0 * (uint64_t)d);
        }
        if (rc == 2) mc_on_device = 0;      /* a ziggurat tail loop beyond the device's skip tables: host generator */
        else if (rc) return -1;
    }
    if (method == HVB200_METHOD_MC && !mc_on_device) {
        mt_seed(&mt, seed);
        for (uint64_t j = 0; j < i0; j++) for (int k = 0; k < d; k++) (void)zig_normal(&mt);
    } else if (method == HVB200_METHOD_MC || method == HVB200_METHOD_HW) {
        ;
    } else if (method == HVB200_METHOD_RPHI) {
        rphi_mantissas(rphi_M, d - 1);
        for (int k = 0; k < d - 1; k++) u[k] = (unsigned __int128)1 << 63;       /* seed 0.5 */
        if (!rphi_cached) rphi_advance(u, rphi_M, d - 1, i0, NULL);
    } else {
        hvb200_set_error("unknown method %d", method);
        return -2;
    }

    /* Rphi-FWE+ with a cacheable prefix: make [0, i1) resident (extending the cached prefix from its saved
       recurrence state), then evaluate straight from the cache */
    if (rphi_cached) {
        const double *view = NULL;
        int rc = 0;
        pthread_mutex_lock(&g_rphi_mutex);
        rphi_cache_t *e = NULL, *victim = &g_rphi_cache[0];
        for (int i = 0; i < RPHI_CACHE_ENTRIES; i++) {
            rphi_cache_t *c = &g_rphi_cache[i];
            if (c->used && c->device == plan->device && c->d == d) { e = c; break; }
            if (!c->used) { if (victim->used) victim = c; }
            else if (victim->used && c->stamp < victim->stamp) victim = c;
        }
        if (!e) {
            /* the victim's buffer goes when its last reader is done */
            e = victim;
            if (e->used) rphi_buf_unref_locked(e->buf);
            memset(e, 0, sizeof *e);
            e->used = 1; e->device = plan->device; e->d = d;
            for (int k = 0; k < d - 1; k++) e->u[k] = (unsigned __int128)1 << 63;        /* seed 0.5 */
        }
        e->stamp = ++g_rphi_stamp;
        if (e->cap < i1) {
            uint64_t cap = (uint64_t)1 << 18;
            while (cap < i1) cap <<= 1;
            if (cap * (uint64_t)(16 * d) > RPHI_CACHE_MAX_BYTES) cap = RPHI_CACHE_MAX_BYTES / (uint64_t)(16 * d);
            void *nb = NULL;
            rphi_buf *nbuf = malloc(sizeof *nbuf);
            if (!nbuf) { hvb200_set_error("out of host memory"); rc = -1; }
            else if (hvb200cu_dev_alloc(plan->device, cap * (size_t)(16 * d), &nb)) { free(nbuf); rc = -1; }
            else {
                nbuf->dev = nb; nbuf->device = plan->device; nbuf->refs = 1;
                if (e->valid && hvb200cu_copy_dev(plan, nb, e->buf->dev, e->valid * (size_t)(2 * d))) {
                    hvb200cu_dev_free(plan->device, nb); free(nbuf); rc = -1;
                } else {
                    rphi_buf_unref_locked(e->buf);       /* freed now, or by the last run still reading it */
                    e->buf = nbuf;
                    e->cap = cap;
                }
            }
        }
        for (uint64_t b0 = e->valid; rc == 0 && b0 < i1; ) {
            const uint64_t cnt = (i1 - b0 < batch) ? i1 - b0 : batch;
            double *uu = hvb200cu_staging(plan, cnt * (size_t)(d - 1));
            if (!uu) { rc = -1; break; }
            /* advance a copy of the recurrence state: it is committed together with `valid` only when the GPU
               steps succeeded, so a failed batch leaves the entry consistent */
            unsigned __int128 unext[HVB200_MAXD];
            memcpy(unext, e->u, sizeof unext);
            rphi_advance(unext, rphi_M, d - 1, cnt, uu);
            if (hvb200cu_gen_rphi(plan, uu, cnt) || hvb200cu_copy_dirs_out(plan, e->buf->dev + b0 * (size_t)(2 * d), cnt, 1)) { rc = -1; break; }
            memcpy(e->u, unext, sizeof unext);
            b0 += cnt;
            e->valid = b0;                   /* the state e->u matches e->valid after every batch */
        }
        rphi_buf *held = NULL;
        if (rc == 0) { held = e->buf; held->refs++; view = held->dev; }
        pthread_mutex_unlock(&g_rphi_mutex);
        if (rc) return -1;
        for (uint64_t b0 = i0; b0 < i1; b0 += batch) {
            const uint64_t cnt = (i1 - b0 < batch) ? i1 - b0 : batch;
            hvb200cu_set_dirs_view(plan, view + b0 * (size_t)(2 * d));
            if (dirs_out && hvb200cu_read_dirs(plan, cnt, dirs_out + (b0 - i0) * (size_t)d)) rc = -1;
            if (rc == 0 && sets_ctx) {
                if (hvb200cu_sets_maxmin(plan, sets_ctx, cnt)) rc = -1;
            } else if (rc == 0 && (!dirs_out || values_out || out_hilo)) {
                if (plan->n && hvb200cu_maxmin(plan, cnt, values_out ? values_out + (b0 - i0) : NULL)) rc = -1;
            }
            if (rc) break;
        }
        hvb200cu_set_dirs_view(plan, NULL);
        if (rc == 0 && hvb200cu_run_end(plan, out_hilo)) rc = -1;       /* synchronous: no kernel reads `view` after it */
        else if (rc) hvb200cu_plan_sync(plan);
        pthread_mutex_lock(&g_rphi_mutex);
        rphi_buf_unref_locked(held);
        pthread_mutex_unlock(&g_rphi_mutex);
        if (rc) return -1;
        plan->stats.samples = i1 - i0;
        return 0;
    }

    for (uint64_t b0 = i0; b0 < i1; b0 += batch) {
        const uint64_t cnt = (i1 - b0 < batch) ? i1 - b0 : batch;
        if (method == HVB200_METHOD_HW) {
            if (hvb200cu_gen_hw(plan, &hw, b0, cnt)) return -1;
        } else if (method == HVB200_METHOD_MC && mc_on_device) {
            const int rc = hvb200cu_gen_mc_device(plan, cnt);
            if (rc == 2) {
                /* 1.5e-12 per pair: a tail loop the device's 16-entry skip tables cannot express.  The rest of the run
                   comes from the host generator (same stream), restarted at this batch */
                mc_on_device = 0;
                mt_seed(&mt, seed);
                for (uint64_t j = 0; j < b0; j++) for (int k = 0; k < d; k++) (void)zig_normal(&mt);
                b0 -= batch;
                continue;
            } else if (rc) return -1;
        } else if (method == HVB200_METHOD_MC) {
            double *z = hvb200cu_staging(plan, cnt * (size_t)d);
            if (!z) return -1;
            for (uint64_t j = 0; j < cnt * (uint64_t)d; j++) z[j] = zig_normal(&mt);
            if (hvb200cu_gen_mc(plan, z, cnt)) return -1;
        } else {
            double *uu = hvb200cu_staging(plan, cnt * (size_t)(d - 1));
            if (!uu) return -1;
            rphi_advance(u, rphi_M, d - 1, cnt, uu);
            if (hvb200cu_gen_rphi(plan, uu, cnt)) return -1;
        }
        if (dirs_out && hvb200cu_read_dirs(plan, cnt, dirs_out + (b0 - i0) * (size_t)d)) return -1;
        if (sets_ctx) {
            if (hvb200cu_sets_maxmin(plan, sets_ctx, cnt)) return -1;
        } else if (!dirs_out || values_out || out_hilo) {
            if (plan->n && hvb200cu_maxmin(plan, cnt, values_out ? values_out + (b0 - i0) : NULL)) return -1;
        }
    }
    if (hvb200cu_run_end(plan, out_hilo)) return -1;
    plan->stats.samples = i1 - i0;
    return 0;
}

int hvb200_plan_run(hvb200_plan *plan, int method, uint_fast32_t nsamples, uint32_t seed,
                    uint64_t i0, uint64_t i1, void *cuda_stream, double out_hilo[2])
{
    return run_range(plan, method, (uint64_t)nsamples, seed, i0, i1, cuda_stream, out_hilo, NULL, NULL);
}
/* ---- DZ2019-MC shards by pair ranges of the MT19937 stream (no rank parses another rank's part) ------------- */
/* pairs of tempered words per emitted normal of the 256-layer ziggurat (c/rng.c:71-99): 98.8 % of the attempts take one
   pair and emit; measured on 4e7 normals.  Only the balance of the last rank depends on it. */
#define HVB200_MC_PAIRS_PER_NORMAL 1.02201
int hvb200_mc_segment_layout(uint_fast32_t nsamples, dimension_t nobjs, int world, int rank, uint64_t *pair0, uint64_t *npairs)
{
    if (world < 1 || rank < 0 || rank >= world || !pair0 || !npairs) { hvb200_set_error("bad MC segment request (rank %d of %d)", rank, world); return -2; }
    const long double normals = (long double)nsamples * (long double)nobjs;
    uint64_t W = (uint64_t)(normals * (long double)HVB200_MC_PAIRS_PER_NORMAL / (long double)world);
    /* short streams are cheaper to parse from the start than to describe: the jump-ahead polynomial costs ~10 ms of host
       time per rank, a prefix parse 0.12 ms per million normals (cfg2, 5e6 normals on 2 GPUs: 1.3 ms from the start,
       10.6 ms by pair ranges).  HVB200_MC_SEGMENT_MIN (pairs per rank) moves the threshold (tests). */
    uint64_t wmin = (uint64_t)1 << 26;
    const char *e = getenv("HVB200_MC_SEGMENT_MIN");
    if (e && atoll(e) >= 4 * 4096) wmin = (uint64_t)atoll(e);
    if (world == 1 || W < wmin) { *pair0 = 0; *npairs = 0; return 1; }
    *pair0 = W * (uint64_t)rank;
    *npairs = rank == world - 1 ? UINT64_MAX : W;
    return 0;
}
int hvb200_plan_mc_segment_scan(hvb200_plan *plan, uint32_t seed, uint64_t pair0, uint64_t npairs, hvb200_mc_segment *out)
{
    if (!plan || !out) { hvb200_set_error("NULL plan or output"); return -2; }
    tls_error[0] = 0;
    uint64_t v[20];
    const int rc = hvb200cu_mc_scan_segment(plan, seed, pair0, npairs, npairs != UINT64_MAX, v);
    if (rc) return rc > 0 ? -10 - rc : -1;        /* -12: tail loop beyond the device tables, -13: parses did not merge */
    out->pair0 = pair0; out->npairs = npairs; out->merge = v[0];
    for (int s = 0; s < 16; s++) out->head[s] = v[1 + s];
    out->body = v[17]; out->skip_out = v[18];
    return 0;
}
int hvb200_mc_segments_resolve(const hvb200_mc_segment *segs, int world, dimension_t nobjs, uint_fast32_t nsamples,
                               hvb200_mc_start *starts)
{
    if (!segs || !starts || world < 1 || nobjs < 1) { hvb200_set_error("bad arguments"); return -2; }
    const uint64_t d = nobjs, N = (uint64_t)nsamples;
    uint64_t s_in = 0, Q = 0;        /* skip state at the head of segment g; normals emitted before it */
    if (segs[0].pair0 != 0) { hvb200_set_error("the first MC segment must start at pair 0"); return -2; }
    for (int g = 0; g < world; g++) {
        if (g + 1 < world && (segs[g].npairs == UINT64_MAX || segs[g + 1].pair0 != segs[g].pair0 + segs[g].npairs)) {
            hvb200_set_error("MC segments %d and %d are not adjacent", g, g + 1); return -2;
        }
        if (s_in > 15 || segs[g].merge < 15) { hvb200_set_error("corrupt MC segment %d", g); return -2; }
        /* the parse of segment g starts s_in pairs into it; the first normal it emits has index Q */
        const uint64_t first = Q;
        uint64_t i0 = (first + d - 1) / d;
        starts[g].pair = segs[g].pair0 + s_in;
        starts[g].discard = i0 * d - first;
        if (i0 > N) i0 = N;
        starts[g].i0 = i0;
        starts[g].i1 = N;
        if (g) starts[g - 1].i1 = i0;
        Q = first + segs[g].head[s_in] + segs[g].body;
        s_in = segs[g].skip_out;
    }
    for (int g = 0; g < world; g++) if (starts[g].i1 < starts[g].i0) starts[g].i1 = starts[g].i0;
    return 0;
}
int hvb200_plan_set_mc_start(hvb200_plan *plan, uint64_t pair, uint64_t discard)
{
    if (!plan) { hvb200_set_error("NULL plan (no point strictly dominates ref)"); return -2; }
    plan->mc_seek = 1; plan->mc_pair = pair; plan->mc_discard = discard;
    return 0;
}

/* Block-cyclic shard of a DZ2019-HW estimate: blocks b = first, first + stride, ... of `block` consecutive
   directions.  The cost of a direction varies smoothly along a Hua-Wang lattice (its first angle grows with the
   index), so contiguous shards are unbalanced (8 GPUs: 85 % efficiency) while interleaved blocks are not.
   Only the lattice is index-addressable: the MC and Rphi streams are sequential and stay on contiguous ranges. */
/* block size of the block-cyclic shards: 2^14 directions, halved (down to 2^10) until every rank gets >= 4 blocks.
   The per-direction cost drifts along the lattice; with b blocks per rank the ranks' loads differ by ~1/b of that
   drift (2^21-direction blocks, 6 per rank at cfg3 on 8 GPUs, left the last rank 4 % behind the first) and the
   kernels map shard positions to lattice indices arithmetically, so small blocks cost nothing. */
uint64_t hvb200_shard_block(uint_fast32_t nsamples, int world)
{
    uint64_t block = HVB200_SHARD_BLOCK;
    if (world < 1) world = 1;
    while (block > ((uint64_t)1 << 10) && block * (uint64_t)world * 4 > (uint64_t)nsamples) block >>= 1;
    return block;
}
int hvb200_plan_run_blocks(hvb200_plan *plan, int method, uint_fast32_t nsamples_, uint32_t seed,
                           uint64_t block, uint64_t first, uint64_t stride, void *cuda_stream, double out_hilo[2])
{
    (void)seed;
    if (!plan) { hvb200_set_error("NULL plan (no point strictly dominates ref)"); return -2; }
    const uint64_t nsamples = (uint64_t)nsamples_;
    const int d = plan->d;
    tls_error[0] = 0;
    if (method != HVB200_METHOD_HW) { hvb200_set_error("block-cyclic runs need an index-addressable direction set (DZ2019-HW)"); return -2; }
    if (d > MOOCORE_HVAPPROX_DIMENSION_MAX + 1) { hvb200_set_error("DZ2019-HW supports at most %d objectives", MOOCORE_HVAPPROX_DIMENSION_MAX + 1); return -2; }
    if (block == 0 || stride == 0 || first >= stride) { hvb200_set_error("bad block shard (block %llu, first %llu, stride %llu)",
                                                                         (unsigned long long)block, (unsigned long long)first, (unsigned long long)stride); return -2; }
    uint64_t total = 0;
    for (uint64_t lo = first * block; lo < nsamples; lo += stride * block)
        total += (nsamples - lo < block) ? nsamples - lo : block;
    uint64_t batch = 0;
    plan->run_samples = total;
    plan->run_total = total;
    plan->want_fused_hw = plan->n > 0 && (block & (block - 1)) == 0;      /* the fused kernel maps positions by shifts */
    if (hvb200cu_run_begin(plan, cuda_stream, total ? total : 1, &batch)) return -1;
    hvb200_hw_params hw;
    hw_params_init(&hw, d, nsamples);
    if (hvb200cu_run_is_fused(plan)) {
        /* one launch over the whole shard: position p of the shard is direction (first + (p / block) stride) block + p % block */
        if (hvb200cu_maxmin_hw_fused(plan, &hw, first * block, __builtin_ctzll(block), stride, total)) return -1;
        if (hvb200cu_run_end(plan, out_hilo)) return -1;
        plan->stats.samples = total;
        return 0;
    }
    if ((block & (block - 1)) == 0) {
        /* power-of-two blocks: the generator maps shard positions to lattice indices itself, one launch per buffer */
        for (uint64_t p0 = 0; p0 < total; p0 += batch) {
            const uint64_t cnt = (total - p0 < batch) ? total - p0 : batch;
            if (hvb200cu_gen_hw_mapped(plan, &hw, first * block, __builtin_ctzll(block), stride, p0, cnt, 0)) return -1;
            if (plan->n && hvb200cu_maxmin(plan, cnt, NULL)) return -1;
        }
        if (hvb200cu_run_end(plan, out_hilo)) return -1;
        plan->stats.samples = total;
        return 0;
    }
    /* other block sizes: the blocks of this shard are generated side by side into the direction buffer and evaluated
       by ONE max-min launch per full buffer: small launches lose ~15 % to start-up and tail */
    uint64_t filled = 0;
    for (uint64_t lo = first * block; lo < nsamples; lo += stride * block) {
        const uint64_t hi = (nsamples - lo < block) ? nsamples : lo + block;
        for (uint64_t b0 = lo; b0 < hi; ) {
            const uint64_t room = batch - filled;
            const uint64_t cnt = (hi - b0 < room) ? hi - b0 : room;
            if (hvb200cu_gen_hw_at(plan, &hw, b0, cnt, filled)) return -1;
            filled += cnt;
            b0 += cnt;
            if (filled == batch) {
                if (plan->n && hvb200cu_maxmin(plan, filled, NULL)) return -1;
                filled = 0;
            }
        }
    }
    if (filled && plan->n && hvb200cu_maxmin(plan, filled, NULL)) return -1;
    if (hvb200cu_run_end(plan, out_hilo)) return -1;
    plan->stats.samples = total;
    return 0;
}
int hvb200_plan_sample_values(hvb200_plan *plan, int method, uint_fast32_t nsamples, uint32_t seed,
                              uint64_t i0, uint64_t count, double *values_out)
{
    double hilo[2];
    return run_range(plan, method, (uint64_t)nsamples, seed, i0, i0 + count, NULL, hilo, values_out, NULL);
}
int hvb200_directions(int method, dimension_t nobjs, uint_fast32_t nsamples, uint32_t seed,
                      uint64_t i0, uint64_t count, double *r_out, int device)
{
    /* a one-point plan just to own the device buffers */
    const int d = (int)nobjs;
    double one[HVB200_MAXD], ref[HVB200_MAXD];
    boolvec mx[HVB200_MAXD];
    if (d < 2 || d > HVB200_DIMENSION_MAX) { hvb200_set_error("nobjs = %d outside [2, %d]", d, HVB200_DIMENSION_MAX); return -2; }
    for (int k = 0; k < d; k++) { one[k] = 0.0; ref[k] = 1.0; mx[k] = 0; }
    hvb200_plan *plan = NULL;
    int rc = hvb200_plan_create(&plan, one, 1, nobjs, ref, mx, device);
    if (rc || !plan) return rc ? rc : -1;
    rc = run_range(plan, method, (uint64_t)nsamples, seed, i0, i0 + count, NULL, NULL, NULL, r_out);
    hvb200_plan_destroy(plan);
    return rc;
}
int hvb200_plan_run_directions(hvb200_plan *plan, const double *r_host, uint64_t count,
                               double *values_out, double out_hilo[2])
{
    if (!plan) { hvb200_set_error("NULL plan (no point strictly dominates ref)"); return -2; }
    tls_error[0] = 0;
    uint64_t batch = 0;
    uint64_t want = count ? count : 1;
    if (want > (1u << 20)) want = 1u << 20;
    plan->run_samples = 0;
    plan->run_total = count;
    plan->want_fused_hw = 0;
    if (hvb200cu_run_begin(plan, NULL, want, &batch)) return -1;
    for (uint64_t b0 = 0; b0 < count; b0 += batch) {
        const uint64_t cnt = (count - b0 < batch) ? count - b0 : batch;
        if (hvb200cu_gen_uploaded(plan, r_host + b0 * (size_t)plan->d, cnt)) return -1;
        if (hvb200cu_maxmin(plan, cnt, values_out ? values_out + b0 : NULL)) return -1;
    }
    if (hvb200cu_run_end(plan, out_hilo)) return -1;
    plan->stats.samples = count;
    return 0;
}

/* (double)(c_d * (sum / (long double)N)) — c/hvapprox.c:256-257, 691-692, 863-864 */
double hvb200_finalize(dimension_t nobjs, uint_fast32_t nsamples, const double *hilo_pairs, int npairs)
{
    long double sum = 0.0L;
    for (int i = 0; i < npairs; i++) sum += (long double)hilo_pairs[2 * i] + (long double)hilo_pairs[2 * i + 1];
    const long double c_d = hvb200_cd[(int)nobjs];
    return (double)(c_d * (sum / (long double)nsamples));
}

/* ------------------------------------------------------------------------------------------ */
/* batched point sets                                                      */
/* ------------------------------------------------------------------------------------------ */
/* The estimates of `nsets` point sets that share ref, maximise and the direction set: set i holds the rows
   [cumsizes[i-1], cumsizes[i]) of `data` (cumsizes[-1] = 0), the layout and the loop of the reference's command
   line tool (c/main-hvapprox.c:149-174).  Directions are generated once; sets that fit a CTA's shared memory
   share one kernel launch per batch of directions, larger ones go through the single-set path.
   volumes_out[i] = 0.0 when no point of set i strictly dominates ref (c/hvapprox.c:228-229), NaN on failure. */
int hvb200_hv_approx_sets(int method, const double *data, const size_t *cumsizes, size_t nsets, dimension_t nobjs,
                          const double *ref, const boolvec *maximise, uint_fast32_t nsamples_, uint32_t seed,
                          int device, double *volumes_out)
{
    tls_error[0] = 0;
    const int d = (int)nobjs;
    const uint64_t nsamples = (uint64_t)nsamples_;
    for (size_t i = 0; i < nsets; i++) volumes_out[i] = NAN;
    if (d < 2 || d > HVB200_DIMENSION_MAX) { hvb200_set_error("nobjs = %d outside [2, %d]", d, HVB200_DIMENSION_MAX); return -2; }
    if (nsamples == 0) { hvb200_set_error("nsamples must be positive"); return -2; }
    if (nsets == 0) return 0;
    if (nsets > 65535) { hvb200_set_error("at most 65535 sets per call"); return -2; }
    const size_t total_rows = cumsizes[nsets - 1];
    double *pts = malloc((total_rows ? total_rows : 1) * (size_t)d * sizeof *pts);
    unsigned *off = malloc((nsets + 1) * sizeof *off);
    int *batched = malloc(nsets * sizeof *batched);
    if (!pts || !off || !batched) { free(pts); free(off); free(batched); hvb200_set_error("out of host memory"); return -4; }
    const size_t fit = hvb200cu_sets_max_points(d);
    size_t kept_total = 0, first_batched = nsets;
    int rc = 0;
    off[0] = 0;
    for (size_t i = 0; i < nsets; i++) {
        const size_t r0 = i ? cumsizes[i - 1] : 0, r1 = cumsizes[i];
        if (r1 < r0 || r1 > total_rows) { hvb200_set_error("cumsizes must be non-decreasing"); rc = -2; break; }
        size_t kept = transform_and_filter(data + r0 * (size_t)d, r1 - r0, d, ref, maximise, pts + kept_total * (size_t)d);
        batched[i] = kept > 0 && kept <= fit;
        if (kept == 0) volumes_out[i] = 0.0;
        if (!batched[i]) kept = 0;                       /* large sets: single-set path below */
        else if (first_batched == nsets) first_batched = i;
        kept_total += kept;
        off[i + 1] = (unsigned)kept_total;
    }
    if (rc == 0 && first_batched < nsets) {
        /* a plan on the first batched set owns the stream, the direction buffers and the generator state */
        hvb200_plan *plan = calloc(1, sizeof *plan);
        void *ctx = NULL;
        double *hilo = malloc(nsets * 2 * sizeof *hilo);
        if (!plan || !hilo) { hvb200_set_error("out of host memory"); rc = -4; }
        if (rc == 0) {
            plan->device = device;
            plan->d = d;
            plan->kernel = HVB200_KERNEL_EXACT;
            plan->n = off[first_batched + 1] - off[first_batched];
            if (device < 0 || device >= hvb200cu_device_count()) {
                hvb200_set_error("CUDA device %d not available (%d visible); this library has no CPU fallback", device, hvb200cu_device_count());
                rc = -3;
            } else if (hvb200cu_plan_upload(plan, pts + (size_t)off[first_batched] * d)) rc = -1;
        }
        if (rc == 0 && hvb200cu_sets_begin(plan, pts, off, (int)nsets, &ctx)) rc = -1;
        if (rc == 0 && run_range_ex(plan, method, nsamples, seed, 0, nsamples, NULL, NULL, NULL, NULL, ctx)) rc = -1;
        if (ctx && hvb200cu_sets_end(plan, ctx, rc == 0 ? hilo : NULL)) rc = -1;
        if (rc == 0)
            for (size_t i = 0; i < nsets; i++)
                if (batched[i]) volumes_out[i] = hvb200_finalize(nobjs, nsamples_, hilo + 2 * i, 1);
        if (plan && plan->dev) hvb200cu_plan_free(plan);
        free(plan);
        free(hilo);
    }
    /* sets too large for the batched kernel */
    for (size_t i = 0; rc == 0 && i < nsets; i++) {
        if (batched[i] || volumes_out[i] == 0.0) continue;
        const size_t r0 = i ? cumsizes[i - 1] : 0, r1 = cumsizes[i];
        hvb200_plan *plan = NULL;
        double hl[2];
        if (hvb200_plan_create(&plan, data + r0 * (size_t)d, r1 - r0, nobjs, ref, maximise, device) || !plan) { rc = -1; break; }
        if (run_range(plan, method, nsamples, seed, 0, nsamples, NULL, hl, NULL, NULL)) rc = -1;
        else volumes_out[i] = hvb200_finalize(nobjs, nsamples_, hl, 1);
        hvb200_plan_destroy(plan);
    }
    free(pts); free(off); free(batched);
    return rc;
}

/* ------------------------------------------------------------------------------------------ */
/* hypervolume-contribution approximation (second half)                     */
/* ------------------------------------------------------------------------------------------ */
/* contrib_out[i] ~ HV(P) - HV(P \ {p_i}) (P without its strictly dominated points when ignore_dominated) from the direction set of `method`: for every direction the arg-max point
   is credited with s1^d - s2^d (largest and second largest max-min value), see hvb200_contrib.cuh.  Points that do not
   strictly dominate ref get 0, like every point when none does.  The reference has no approximate implementation
   (c/hv_contrib.c is exact): parity is unpinned, the estimate is cross-checked against the exact contributions. */
int hvb200_hv_contributions_approx(int method, const double *data, size_t npoints, dimension_t nobjs, const double *ref,
                                   const boolvec *maximise, uint_fast32_t nsamples_, uint32_t seed, int device,
                                   int ignore_dominated, double *contrib_out, double *hv_out)
{
    tls_error[0] = 0;
    const int d = (int)nobjs;
    const uint64_t nsamples = (uint64_t)nsamples_;
    for (size_t i = 0; i < npoints; i++) contrib_out[i] = 0.0;
    if (hv_out) *hv_out = 0.0;
    if (d < 2 || d > MOOCORE_HVAPPROX_DIMENSION_MAX + 1) { hvb200_set_error("nobjs = %d outside [2, %d]", d, MOOCORE_HVAPPROX_DIMENSION_MAX + 1); return -2; }
    if (nsamples == 0) { hvb200_set_error("nsamples must be positive"); return -2; }
    if (device < 0 || device >= hvb200cu_device_count()) {
        hvb200_set_error("CUDA device %d not available (%d visible); this library has no CPU fallback", device, hvb200cu_device_count());
        return -3;
    }
    /* transform_and_filter (c/hvapprox.c:30-66) keeping the original row of every surviving point */
    double *pts = malloc((npoints ? npoints : 1) * (size_t)d * sizeof *pts);
    size_t *orig = malloc((npoints ? npoints : 1) * sizeof *orig);
    if (!pts || !orig) { free(pts); free(orig); hvb200_set_error("out of host memory"); return -4; }
    size_t kept = 0;
    for (size_t i = 0; i < npoints; i++) {
        int k;
        for (k = 0; k < d; k++) {
            const double v = maximise[k] ? data[i * (size_t)d + k] - ref[k] : ref[k] - data[i * (size_t)d + k];
            if (!(v > 0)) break;
            pts[kept * (size_t)d + k] = v;
        }
        if (k == d) orig[kept++] = i;
    }
    int rc = 0;
    if (ignore_dominated && kept > 1) {
        /* hv_contributions(..., ignore_dominated = true) of the reference (c/hv_contrib.c:348): strictly dominated points
           are taken out first (contribution 0) and do not shield the points that dominate them; duplicates stay */
        unsigned *flags = malloc(kept * sizeof *flags);
        if (!flags) { hvb200_set_error("out of host memory"); rc = -4; }
        else if (hvb200cu_nd_strict_flags(device, pts, kept, d, flags)) rc = -1;
        else {
            size_t m = 0;
            for (size_t j = 0; j < kept; j++)
                if (flags[j]) {
                    if (m != j) { memmove(pts + m * (size_t)d, pts + j * (size_t)d, (size_t)d * sizeof *pts); orig[m] = orig[j]; }
                    m++;
                }
            kept = m;
        }
        free(flags);
    }
    if (rc == 0 && kept) {
        hvb200_plan *plan = calloc(1, sizeof *plan);
        double *sums = malloc(kept * sizeof *sums);
        void *ctx = NULL;
        double hilo[2] = {0.0, 0.0};
        if (!plan || !sums) { hvb200_set_error("out of host memory"); rc = -4; }
        if (rc == 0) {
            plan->device = device;
            plan->d = d;
            plan->n = kept;
            plan->kernel = HVB200_KERNEL_EXACT;
            plan->no_ndfilter = 1;                    /* a dominated point can be the second best of a direction */
            if (hvb200cu_plan_upload(plan, pts)) rc = -1;
        }
        if (rc == 0 && hvb200cu_contrib_begin(plan, &ctx)) rc = -1;
        if (rc == 0 && run_range_ex(plan, method, nsamples, seed, 0, nsamples, NULL, hilo, NULL, NULL, ctx)) rc = -1;
        if (ctx && hvb200cu_contrib_end(plan, ctx, rc == 0 ? sums : NULL)) rc = -1;
        if (rc == 0) {
            const long double c_d = hvb200_cd[d];
            for (size_t j = 0; j < kept; j++) contrib_out[orig[j]] = (double)(c_d * ((long double)sums[j] / (long double)nsamples));
            if (hv_out) *hv_out = hvb200_finalize(nobjs, nsamples_, hilo, 1);
        }
        if (plan && plan->dev) hvb200cu_plan_free(plan);
        free(plan);
        free(sums);
    }
    free(pts);
    free(orig);
    return rc;
}

/* ------------------------------------------------------------------------------------------ */
/* epsilon indicator (c/epsilon.h:197-221; python/src/moocore/libmoocore.h:24-25)              */
/* ------------------------------------------------------------------------------------------ */
static double epsilon_common(int mult, const double *data, size_t n, dimension_t dim, const double *ref, size_t ref_size,
                             const boolvec *maximise)
{
    tls_error[0] = 0;
    const int d = (int)dim;
    if (d < 1 || d > HVB200_MAXD || n == 0 || ref_size == 0) {
        hvb200_set_error("epsilon: need 1..%d objectives and non-empty sets (d=%d, n=%zu, ref_size=%zu)", HVB200_MAXD, d, n, ref_size);
        return fail_nan(mult ? "epsilon_mult" : "epsilon_additive");
    }
    const char *env = getenv("HVB200_DEVICE");
    const int device = env ? atoi(env) : 0;
    if (device < 0 || device >= hvb200cu_device_count()) {
        hvb200_set_error("CUDA device %d not available (%d visible); this library has no CPU fallback", device, hvb200cu_device_count());
        return fail_nan(mult ? "epsilon_mult" : "epsilon_additive");
    }
    signed char mm[HVB200_MAXD];
    for (int k = 0; k < d; k++) mm[k] = maximise[k] ? 1 : -1;      /* minmax_from_boolvec, c/nondominated.h */
    double out = NAN;
    if (hvb200cu_epsilon(device, mult, data, n, ref, ref_size, d, mm, &out)) return fail_nan(mult ? "epsilon_mult" : "epsilon_additive");
    return out;
}
double epsilon_additive(const double *restrict data, size_t n, dimension_t dim, const double *restrict ref, size_t ref_size,
                        const boolvec *restrict maximise)
{
    return epsilon_common(0, data, n, dim, ref, ref_size, maximise);
}
double epsilon_mult(const double *restrict data, size_t n, dimension_t dim, const double *restrict ref, size_t ref_size,
                    const boolvec *restrict maximise)
{
    return epsilon_common(1, data, n, dim, ref, ref_size, maximise);
}

/* ------------------------------------------------------------------------------------------ */
/* IGD, IGD+, averaged Hausdorff distance (c/igd.h:196-287; libmoocore.h:20-22)                */
/* ------------------------------------------------------------------------------------------ */
static int gd_setup(const char *who, dimension_t dim, size_t n, size_t ref_size, const boolvec *maximise, signed char *mm, int *device)
{
    tls_error[0] = 0;
    const int d = (int)dim;
    if (d < 1 || d > HVB200_MAXD || ref_size == 0) {
        hvb200_set_error("%s: need 1..%d objectives and a non-empty reference set (d=%d, n=%zu, ref_size=%zu)", who, HVB200_MAXD, d, n, ref_size);
        return -1;
    }
    const char *env = getenv("HVB200_DEVICE");
    *device = env ? atoi(env) : 0;
    if (*device < 0 || *device >= hvb200cu_device_count()) {
        hvb200_set_error("CUDA device %d not available (%d visible); this library has no CPU fallback", *device, hvb200cu_device_count());
        return -1;
    }
    for (int k = 0; k < d; k++) mm[k] = maximise[k] ? 1 : -1;      /* minmax_from_boolvec */
    return 0;
}
double IGD(const double *restrict data, size_t npoints, dimension_t nobj, const double *restrict ref, size_t ref_size,
           const boolvec *restrict maximise)
{   /* IGD_minmax, c/igd.h:186-194: gd_common(ref as a, data as r, plus=false, psize=false, p=1) */
    signed char mm[HVB200_MAXD];
    int device;
    double v;
    if (gd_setup("IGD", nobj, npoints, ref_size, maximise, mm, &device)) return fail_nan("IGD");
    if (hvb200cu_gd_common(device, ref, ref_size, data, npoints, (int)nobj, mm, 0, 0, 1, &v)) return fail_nan("IGD");
    return v;
}
double IGD_plus(const double *restrict data, size_t npoints, dimension_t nobj, const double *restrict ref, size_t ref_size,
                const boolvec *restrict maximise)
{   /* IGD_plus_minmax, c/igd.h:232-241: plus=true, psize=true, p=1 */
    signed char mm[HVB200_MAXD];
    int device;
    double v;
    if (gd_setup("IGD_plus", nobj, npoints, ref_size, maximise, mm, &device)) return fail_nan("IGD_plus");
    if (hvb200cu_gd_common(device, ref, ref_size, data, npoints, (int)nobj, mm, 1, 1, 1, &v)) return fail_nan("IGD_plus");
    return v;
}
double avg_Hausdorff_dist(const double *restrict data, size_t npoints, dimension_t nobj, const double *restrict ref, size_t ref_size,
                          const boolvec *restrict maximise, unsigned int p)
{   /* avg_Hausdorff_dist_minmax, c/igd.h:257-273: MAX(GD_p, IGD_p), psize=true, p cast to uint_fast8_t */
    signed char mm[HVB200_MAXD];
    int device;
    double gd_p, igd_p;
    if (gd_setup("avg_Hausdorff_dist", nobj, npoints, ref_size, maximise, mm, &device)) return fail_nan("avg_Hausdorff_dist");
    const unsigned pp = (unsigned)(unsigned char)p;
    if (pp == 0) { hvb200_set_error("avg_Hausdorff_dist: p must be positive"); return fail_nan("avg_Hausdorff_dist"); }
    if (hvb200cu_gd_common(device, data, npoints, ref, ref_size, (int)nobj, mm, 0, 1, pp, &gd_p)) return fail_nan("avg_Hausdorff_dist");
    if (hvb200cu_gd_common(device, ref, ref_size, data, npoints, (int)nobj, mm, 0, 1, pp, &igd_p)) return fail_nan("avg_Hausdorff_dist");
    return (gd_p < igd_p) ? igd_p : gd_p;                          /* MAX(gd_p, igd_p) */
}

/* ------------------------------------------------------------------------------------------ */
/* whv_hype: HypE weighted-hypervolume sampling, 2 objectives (c/whv_hype.c, c/whv_hype.h)      */
/* ------------------------------------------------------------------------------------------ */
/* estimate_whv's sum (c/whv_hype.c:183-189): for every sample, 1.0/count added count times to ONE running double.
   Within a binade [2^(e-1), 2^e) every double is a multiple of u = 2^(e-53); if S + inc rounds to S + step and the
   rounding is not a tie, then S' + inc rounds to S' + step for every S' of that binade, so j additions are the exact
   sum S + j*step as long as it stays below 2^e.  Ties and binade crossings take literal additions. */
double hvb200_whv_ordered_sum(const unsigned *counts, size_t nsamples, int naive)
{
    double S = 0.0;
    for (size_t s = 0; s < nsamples; s++) {
        unsigned left = counts[s];
        if (!left) continue;
        const double inc = 1.0 / counts[s];
        while (!naive && left > 1 && S >= 1.0) {                 /* S >= 1 >= inc */
            int ex;
            (void)frexp(S, &ex);
            const double top = ldexp(1.0, ex), u = ldexp(1.0, ex - 53);
            const double t = S + inc;
            const double step = t - S;                           /* exact */
            unsigned j = 0;
            if (t < top && fabs(inc - step) != 0.5 * u) {        /* inc - step: the exact rounding error (Fast2Sum) */
                if (step == 0.0) { left = 0; break; }
                const double room = (top - S) / step;            /* additions that stay below top, +-1 */
                j = (room >= (double)left + 2.0) ? left : (room > 2.0 ? (unsigned)(room - 2.0) : 0u);
            }
            if (j == 0) { S = t; left--; continue; }
            S += (double)j * step;                               /* exact: a multiple of u below top */
            left -= j;
        }
        for (; left; left--) S += inc;
    }
    return S;
}

enum { WHV_UNIF, WHV_EXPO, WHV_GAUS };

static double whv_hype_common(int dist, const char *who, const double *points, int npoints, const double *ideal, const double *ref,
                              int nsamples, uint32_t seed, double mu_expo, const double *mu_gaus)
{
    tls_error[0] = 0;
    if (npoints < 0 || nsamples <= 0 || !ideal || !ref || (npoints > 0 && !points) || (dist == WHV_GAUS && !mu_gaus)) {
        hvb200_set_error("%s: need npoints >= 0, nsamples > 0 and non-null arguments (npoints=%d, nsamples=%d)", who, npoints, nsamples);
        return fail_nan(who);
    }
    const char *env = getenv("HVB200_DEVICE");
    const int device = env ? atoi(env) : 0;
    if (device < 0 || device >= hvb200cu_device_count()) {
        hvb200_set_error("CUDA device %d not available (%d visible); this library has no CPU fallback", device, hvb200cu_device_count());
        return fail_nan(who);
    }
    /* normalise(): (x - ideal) / (ref - ideal) with a zero range replaced by 1 (c/nondominated.h:878-901, called with
       the range [0, 1] and all-minimise by normalise01_inplace, c/whv_hype.c:205-212) */
    double diff[2];
    for (int k = 0; k < 2; k++) { diff[k] = ref[k] - ideal[k]; if (diff[k] == 0.0) diff[k] = 1.0; }
    const size_t np = (size_t)npoints, ns = (size_t)nsamples;
    double *samples = malloc(ns * 2 * sizeof(double));
    double *pts = malloc((np ? np : 1) * 2 * sizeof(double));
    unsigned *counts = calloc(ns, sizeof(unsigned));
    if (!samples || !pts || !counts) {
        free(samples); free(pts); free(counts);
        hvb200_set_error("%s: out of memory", who);
        return fail_nan(who);
    }
    mt_stream rng;
    mt_seed(&rng, seed);
    const double lower[2] = {0.0, 0.0}, range[2] = {1.0, 1.0};       /* hype_dist_new, c/whv_hype.c:94-107 */
    if (dist == WHV_UNIF) {                                           /* uniform_dist_sample, :78-92 */
        for (size_t i = 0; i < ns; i++)
            for (int k = 0; k < 2; k++) samples[2 * i + k] = lower[k] + mt_unit(&rng) * range[k];
    } else if (dist == WHV_EXPO) {                                    /* exp_dist_sample, :50-76 */
        const size_t half = (size_t)(int)(0.5 * nsamples);
        for (size_t i = 0; i < half; i++) {
            double x = mt_unit(&rng);
            samples[2 * i] = lower[0] - mu_expo * log(x);
            x = mt_unit(&rng);
            samples[2 * i + 1] = lower[1] + x * range[1];
        }
        for (size_t i = half; i < ns; i++) {
            double x = mt_unit(&rng);
            samples[2 * i] = lower[0] + x * range[0];
            x = mt_unit(&rng);
            samples[2 * i + 1] = lower[1] - mu_expo * log(x);
        }
    } else {                                                          /* gaussian_dist_sample, :30-48 */
        double mu[2];
        for (int k = 0; k < 2; k++) mu[k] = 0.0 + 1.0 * (mu_gaus[k] - ideal[k]) / diff[k];      /* :288 */
        const double sigma1 = 0.25, sigma2 = 0.25, rho = 1.0;
        const double sigma2rho = sigma2 * rho, nu = sigma2 * sqrt(1 - rho * rho);               /* c/rng.c:101-116 */
        for (size_t i = 0; i < ns; i++) {
            const double x1 = zig_normal(&rng);
            samples[2 * i] = mu[0] + x1 * sigma1;
            samples[2 * i + 1] = mu[1] + x1 * sigma2rho + nu * zig_normal(&rng);
        }
    }
    for (size_t i = 0; i < np; i++)
        for (int k = 0; k < 2; k++) pts[2 * i + k] = 0.0 + 1.0 * (points[2 * i + k] - ideal[k]) / diff[k];
    double whv = NAN;
    if (np == 0 || hvb200cu_whv_counts(device, pts, np, samples, ns, counts) == 0) {
        whv = hvb200_whv_ordered_sum(counts, ns, 0);
        double volume = 1.0;                                          /* calculate_volume_between_points, :197-203 */
        for (int k = 0; k < 2; k++) volume *= (ref[k] - ideal[k]);
        whv *= volume / nsamples;                                     /* Eq. 18, :236 */
    }
    free(samples); free(pts); free(counts);
    if (whv != whv && tls_error[0]) return fail_nan(who);
    return whv;
}
double whv_hype_unif(const double *points, int npoints, const double *ideal, const double *ref, int nsamples, uint32_t seed)
{
    return whv_hype_common(WHV_UNIF, "whv_hype_unif", points, npoints, ideal, ref, nsamples, seed, 0.0, NULL);
}
double whv_hype_expo(const double *points, int npoints, const double *ideal, const double *ref, int nsamples, uint32_t seed, double mu)
{
    return whv_hype_common(WHV_EXPO, "whv_hype_expo", points, npoints, ideal, ref, nsamples, seed, mu, NULL);
}
double whv_hype_gaus(const double *points, int npoints, const double *ideal, const double *ref, int nsamples, uint32_t seed,
                     const double *mu)
{
    return whv_hype_common(WHV_GAUS, "whv_hype_gaus", points, npoints, ideal, ref, nsamples, seed, 0.0, mu);
}

/* ------------------------------------------------------------------------------------------ */
/* single-process multi-GPU                                                                    */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    int method, device, rc;
    const double *data; size_t npoints; dimension_t nobjs; const double *ref; const boolvec *maximise;
    uint64_t nsamples, i0, i1; uint32_t seed;
    int stride;             /* > 1: block-cyclic shard `device` of `stride` (DZ2019-HW) */
    double hilo[2];
    int empty;
    hvb200_plan *plan;      /* kept alive until the exchange step */
    void *share;            /* host-side preparation shared by the shards (same point set on every GPU) */
    char err[512];
    /* DZ2019-MC by pair ranges: phase 1 = create + describe the own range, phase 2 = run from the resolved start */
    int phase;              /* 0: everything in one go, 1: plan + segment scan, 2: run */
    uint64_t pair0, npairs;
    hvb200_mc_segment seg;
    hvb200_mc_start start;
} shard_job;

/* how the last multi-GPU call of this thread combined its partial sums: 0 = one device (nothing to exchange),
   1 = NCCL all-gather over NVLink, 2 = host copies (NCCL unavailable or disabled) */
static __thread int tls_last_exchange;
int hvb200_multi_last_exchange(void) { return tls_last_exchange; }

static void *shard_main(void *arg)
{
    shard_job *j = arg;
    hvb200_plan *plan = NULL;
    if (j->phase == 2) {
        plan = j->plan;
        if (j->rc || j->empty || !plan) return NULL;
        j->rc = hvb200_plan_set_mc_start(plan, j->start.pair, j->start.discard);
        if (j->rc == 0) j->rc = hvb200_plan_run(plan, j->method, j->nsamples, j->seed, j->start.i0, j->start.i1, NULL, j->hilo);
        if (j->rc) snprintf(j->err, sizeof j->err, "%s", tls_error);
        return NULL;
    }
    j->hilo[0] = j->hilo[1] = 0.0;
    j->rc = plan_create_shared(&plan, j->data, j->npoints, j->nobjs, j->ref, j->maximise, j->device, j->share);
    if (j->rc == 0 && !plan) { j->empty = 1; return NULL; }
    if (j->phase == 1) {
        if (j->rc == 0) j->rc = hvb200_plan_mc_segment_scan(plan, j->seed, j->pair0, j->npairs, &j->seg);
        if (j->rc) snprintf(j->err, sizeof j->err, "%s", tls_error);
        j->plan = plan;
        return NULL;
    }
    if (j->rc == 0 && j->stride > 1 && j->method == HVB200_METHOD_HW)
        j->rc = hvb200_plan_run_blocks(plan, j->method, j->nsamples, j->seed, hvb200_shard_block(j->nsamples, j->stride), (uint64_t)j->device, (uint64_t)j->stride, NULL, j->hilo);
    else if (j->rc == 0) j->rc = hvb200_plan_run(plan, j->method, j->nsamples, j->seed, j->i0, j->i1, NULL, j->hilo);
    if (j->rc) snprintf(j->err, sizeof j->err, "%s", tls_error);
    j->plan = plan;
    return NULL;
}

double hvb200_hv_approx_multi(int method, const double *data, size_t npoints, dimension_t nobjs,
                              const double *ref, const boolvec *maximise,
                              uint_fast32_t nsamples, uint32_t seed, int ndevices)
{
    tls_error[0] = 0;
    const int avail = hvb200cu_device_count();
    if (avail < 1) { hvb200_set_error("no CUDA device visible; this library has no CPU fallback"); return fail_nan("hv_approx"); }
    if (ndevices < 1) ndevices = 1;
    if (ndevices > avail) ndevices = avail;
    if (ndevices > 16) ndevices = 16;
    if ((uint64_t)ndevices > (uint64_t)nsamples) ndevices = 1;
    shard_job jobs[16];
    pthread_t th[16];
    memset(jobs, 0, sizeof jobs);
    void *share = ndevices > 1 ? hvb200cu_share_new() : NULL;
    for (int g = 0; g < ndevices; g++) {
        shard_job *j = &jobs[g];
        j->method = method; j->device = g; j->share = share;
        j->data = data; j->npoints = npoints; j->nobjs = nobjs; j->ref = ref; j->maximise = maximise;
        j->nsamples = (uint64_t)nsamples; j->seed = seed;
        j->i0 = (uint64_t)nsamples * (uint64_t)g / (uint64_t)ndevices;
        j->i1 = (uint64_t)nsamples * (uint64_t)(g + 1) / (uint64_t)ndevices;
        j->stride = ndevices;
    }
    /* DZ2019-MC on several GPUs: every shard owns a pair range of the stream (nobody parses a prefix); two rounds of
       threads with the host-side resolve step in between.  HVB200_MC_SHARD=prefix keeps the contiguous sample ranges. */
    int mc_segments = 0;
    if (method == HVB200_METHOD_MC && ndevices > 1 && !(getenv("HVB200_MC") && !strcmp(getenv("HVB200_MC"), "host")) &&
        !(getenv("HVB200_MC_SHARD") && !strcmp(getenv("HVB200_MC_SHARD"), "prefix"))) {
        mc_segments = 1;
        for (int g = 0; g < ndevices; g++)
            if (hvb200_mc_segment_layout(nsamples, nobjs, ndevices, g, &jobs[g].pair0, &jobs[g].npairs) != 0) mc_segments = 0;
    }
    for (int round = 0; round < (mc_segments ? 2 : 1); round++) {
        for (int g = 0; g < ndevices; g++) jobs[g].phase = mc_segments ? round + 1 : 0;
        if (ndevices == 1) {
            shard_main(&jobs[0]);
        } else {
            int started[16] = {0};
            for (int g = 0; g < ndevices; g++) started[g] = pthread_create(&th[g], NULL, shard_main, &jobs[g]) == 0;
            for (int g = 0; g < ndevices; g++) {
                if (started[g]) pthread_join(th[g], NULL);
                else shard_main(&jobs[g]);               /* no thread: this shard runs inline, it is never dropped */
            }
        }
        if (mc_segments && round == 0) {
            int bad = 0, any_empty = 0;
            hvb200_mc_segment segs[16];
            hvb200_mc_start starts[16];
            for (int g = 0; g < ndevices; g++) { bad |= jobs[g].rc != 0; any_empty |= jobs[g].empty; segs[g] = jobs[g].seg; }
            if (any_empty) break;
            if (!bad && hvb200_mc_segments_resolve(segs, ndevices, nobjs, nsamples, starts) == 0) {
                for (int g = 0; g < ndevices; g++) jobs[g].start = starts[g];
            } else {
                /* a tail loop beyond the device tables, or parses that did not merge (both ~1e-12 per pair): fall back to
                   contiguous sample ranges, where the run itself switches to the host generator if it has to */
                for (int g = 0; g < ndevices; g++) {
                    if (!jobs[g].plan) continue;             /* no plan: that shard's failure stands and is reported */
                    jobs[g].rc = 0; jobs[g].err[0] = 0;
                    jobs[g].start.pair = 0; jobs[g].start.discard = jobs[g].i0 * (uint64_t)nobjs;
                    jobs[g].start.i0 = jobs[g].i0; jobs[g].start.i1 = jobs[g].i1;
                }
            }
        }
    }
    double pairs[32];
    int failed = -1, empty = 0;
    hvb200_plan *plans[16];
    for (int g = 0; g < ndevices; g++) {
        if (jobs[g].rc && failed < 0) failed = g;
        empty |= jobs[g].empty;
        plans[g] = jobs[g].plan;
        pairs[2 * g] = jobs[g].hilo[0];
        pairs[2 * g + 1] = jobs[g].hilo[1];
    }
    double result;
    if (failed >= 0) {
        hvb200_set_error("device %d: %s", failed, jobs[failed].err);
        result = fail_nan("hv_approx");
    } else if (empty) {
        result = 0.0;
    } else {
        /* exchange step: one NCCL all-gather of the 16-byte partials over NVLink; every GPU ends up
           with all pairs, device 0's copy is finalised (fixed order, long double).  Without NCCL the
           host copies of the same pairs are used. */
        tls_last_exchange = ndevices > 1 ? 2 : 0;
        if (ndevices > 1 && !getenv("HVB200_NO_NCCL")) {
            double gathered[32];
            const int rc = hvb200cu_nccl_allgather(plans, ndevices, gathered);
            if (rc == 0) { memcpy(pairs, gathered, 2 * (size_t)ndevices * sizeof(double)); tls_last_exchange = 1; }
            else if (rc < 0) fprintf(stderr, "moocore_b200: NCCL exchange failed (%s); using host copies\n", tls_error);
        }
        result = hvb200_finalize(nobjs, nsamples, pairs, ndevices);
    }
    for (int g = 0; g < ndevices; g++) hvb200_plan_destroy(jobs[g].plan);
    if (share) hvb200cu_share_free(share);
    return result;
}

static int env_ndevices(void)
{
    const char *e = getenv("HVB200_NGPUS");
    if (!e) return 1;
    const int v = atoi(e);
    return v < 1 ? 1 : v;
}

/* ------------------------------------------------------------------------------------------ */
/* the drop-in entry points (c/hvapprox.h:16-38)                                               */
/* ------------------------------------------------------------------------------------------ */
double hv_approx_normal(const double *restrict data, size_t npoints, dimension_t nobjs,
                        const double *restrict ref, const boolvec *restrict maximise,
                        uint_fast32_t nsamples, uint32_t random_seed)
{
    return hvb200_hv_approx_multi(HVB200_METHOD_MC, data, npoints, nobjs, ref, maximise, nsamples, random_seed, env_ndevices());
}
double hv_approx_hua_wang(const double *restrict data, size_t npoints, dimension_t nobjs,
                          const double *restrict ref, const boolvec *restrict maximise,
                          uint_fast32_t nsamples)
{
    return hvb200_hv_approx_multi(HVB200_METHOD_HW, data, npoints, nobjs, ref, maximise, nsamples, 0, env_ndevices());
}
double hv_approx_rphi_fang_wang_plus(const double *restrict data, size_t npoints, dimension_t nobjs,
                                     const double *restrict ref, const boolvec *restrict maximise,
                                     uint_fast32_t nsamples)
{
    return hvb200_hv_approx_multi(HVB200_METHOD_RPHI, data, npoints, nobjs, ref, maximise, nsamples, 0, env_ndevices());
}

double hvb200_measure_fp64_flops(int device, double *dmul_per_s, double *dsetp_per_s)
{
    return hvb200cu_measure_fp64(device, dmul_per_s, dsetp_per_s);
}
