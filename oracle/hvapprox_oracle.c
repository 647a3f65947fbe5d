/* oracle/hvapprox_oracle.c — CPU restatement of moocore's hv_approx path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under moocore_b200/ may include, link or call this
 * file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs use it, and only as the checker.
 *
 * Parity status: PINNED.  tests/test_oracle_cpu.py checks this file against
 *   (1) the reference's own known-answer values (R doctest r/tests/testthat/
 *       test-doctest-hv_approx.R:9-11, Python doctests python/src/moocore/_moocore.py:1090-1130),
 *   (2) golden vectors produced by the compiled reference (tests/golden/, generator
 *       tests/golden/make_golden.py), and
 *   (3) when oracle/_ref/libhvref.so is present, the reference itself on random windows.
 *
 * It is a restatement, not a copy: the pow chains are table-driven, the closed forms of
 * int_0^b sin^m are generated from the reduction formula instead of 32 hand-written cases,
 * and the constant tables come from oracle/gen_tables.py.  Where the reference's result is an
 * artefact of its own rounding noise (Newton solve with a zero target) the generated table
 * orc_theta0_bits holds what the reference produces.
 *
 * All file:line citations are relative to /root/reference.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "oracle_tables.h"

#define ORC_API __attribute__((visibility("default")))
enum { MAXDIM = 64 };

static const double TINY_WEIGHT = 1e-20; /* ALMOST_ZERO_WEIGHT, c/hvapprox.c:9 */

/* ------------------------------------------------------------------------------------------
 * x^e with the reference's multiplication trees (c/pow_int.h:16-74).
 *
 * Each exponent <= 31 is a straight-line program over registers v[0]=x, v[1], ...: step t
 * computes v[t+1] = v[a]*v[b]; the value of the last step is the result.  IEEE multiplication
 * commutes, so only the tree shape matters for the rounding, and the shapes below are the
 * ones the reference's switch produces (e.g. 7 -> ((x2*x2)*x2)*x, 23 -> x2, x3=x2*x, x5=x2*x3,
 * x10, x20, x20*x3).  Exponents > 31 use right-to-left square-and-multiply (pow_int.h:65-73).
 * ---------------------------------------------------------------------------------------- */
typedef struct { uint8_t n; uint8_t op[7][2]; } pow_prog;
static const pow_prog POW_PROG[32] = {
    /* 0*/ {0, {{0}}},
    /* 1*/ {0, {{0}}},
    /* 2*/ {1, {{0,0}}},
    /* 3*/ {2, {{0,0},{1,0}}},
    /* 4*/ {2, {{0,0},{1,1}}},
    /* 5*/ {3, {{0,0},{1,1},{2,0}}},
    /* 6*/ {3, {{0,0},{1,1},{2,1}}},
    /* 7*/ {4, {{0,0},{1,1},{2,1},{3,0}}},
    /* 8*/ {3, {{0,0},{1,1},{2,2}}},
    /* 9*/ {4, {{0,0},{1,0},{2,2},{3,2}}},
    /*10*/ {4, {{0,0},{1,1},{2,2},{3,1}}},
    /*11*/ {5, {{0,0},{1,1},{2,2},{3,1},{4,0}}},
    /*12*/ {4, {{0,0},{1,1},{2,2},{3,2}}},
    /*13*/ {5, {{0,0},{1,1},{2,2},{3,2},{4,0}}},
    /*14*/ {5, {{0,0},{1,1},{2,2},{3,2},{4,1}}},
    /*15*/ {5, {{0,0},{1,1},{2,0},{3,3},{4,3}}},
    /*16*/ {4, {{0,0},{1,1},{2,2},{3,3}}},
    /*17*/ {5, {{0,0},{1,1},{2,2},{3,3},{4,0}}},
    /*18*/ {5, {{0,0},{1,0},{2,2},{3,2},{4,4}}},
    /*19*/ {6, {{0,0},{1,0},{2,2},{3,2},{4,4},{5,0}}},
    /*20*/ {5, {{0,0},{1,1},{2,2},{3,1},{4,4}}},
    /*21*/ {6, {{0,0},{1,1},{2,2},{3,1},{4,4},{5,0}}},
    /*22*/ {6, {{0,0},{1,1},{2,0},{3,3},{4,4},{5,1}}},
    /*23*/ {6, {{0,0},{1,0},{1,2},{3,3},{4,4},{5,2}}},
    /*24*/ {5, {{0,0},{1,1},{2,2},{3,2},{4,4}}},
    /*25*/ {6, {{0,0},{1,1},{2,2},{3,2},{4,4},{5,0}}},
    /*26*/ {6, {{0,0},{1,1},{2,2},{3,2},{4,0},{5,5}}},
    /*27*/ {6, {{0,0},{1,0},{2,2},{3,3},{4,4},{5,2}}},
    /*28*/ {6, {{0,0},{1,1},{2,2},{3,2},{4,1},{5,5}}},
    /*29*/ {7, {{0,0},{1,1},{2,2},{3,2},{4,1},{5,5},{6,0}}},
    /*30*/ {6, {{0,0},{1,1},{2,0},{3,3},{4,3},{5,5}}},
    /*31*/ {7, {{0,0},{1,1},{2,0},{3,3},{4,3},{5,5},{6,0}}},
};

ORC_API double orc_pow_uint(double x, unsigned e)
{
    if (e == 0) return 1.0;
    if (e <= 31) {
        double v[8];
        v[0] = x;
        const pow_prog *p = &POW_PROG[e];
        for (int t = 0; t < p->n; t++) v[t + 1] = v[p->op[t][0]] * v[p->op[t][1]];
        return v[p->n];
    }
    double acc = (e & 1u) ? x : 1.0;
    for (e >>= 1; e; e >>= 1) {
        x *= x;
        if (e & 1u) acc *= x;
    }
    return acc;
}
static long double powl_chain(long double x, unsigned e)
{   /* long-double twin (c/pow_int.h:76-134): same trees. */
    if (e == 0) return 1.0L;
    if (e <= 31) {
        long double v[8];
        v[0] = x;
        const pow_prog *p = &POW_PROG[e];
        for (int t = 0; t < p->n; t++) v[t + 1] = v[p->op[t][0]] * v[p->op[t][1]];
        return v[p->n];
    }
    long double acc = (e & 1u) ? x : 1.0L;
    for (e >>= 1; e; e >>= 1) {
        x *= x;
        if (e & 1u) acc *= x;
    }
    return acc;
}

/* ------------------------------------------------------------------------------------------
 * transform_and_filter (c/hvapprox.c:30-66): p' = maximise ? p - ref : ref - p, rows with any
 * p'_k <= 0 dropped, survivors compacted in input order.  `maximise` is given as bytes here
 * (boolvec = uint8_t outside R, c/config.h:31-39).  Returns the number of survivors.
 * ---------------------------------------------------------------------------------------- */
ORC_API size_t orc_transform_and_filter(const double *data, size_t n, int d, const double *ref,
                                        const uint8_t *maximise, double *out)
{
    size_t kept = 0;
    for (size_t i = 0; i < n; i++) {
        double *row = out + kept * (size_t)d;
        int k = 0;
        for (; k < d; k++) {
            double v = maximise[k] ? data[i * (size_t)d + k] - ref[k] : ref[k] - data[i * (size_t)d + k];
            if (v <= 0) break;
            row[k] = v;
        }
        if (k == d) kept++;
    }
    return kept;
}

/* get_expected_value (c/hvapprox.c:68-117): (max_p min_k p_k r_k)^d.  The reference's 16-point
 * blocking and scalar early-out only reorder exact max/min operations, so a plain double loop
 * gives the same bits.  max starts at 0 (c/hvapprox.c:83). */
ORC_API double orc_maxmin(const double *pts, size_t n, int d, const double *r)
{
    double best = 0.0;
    for (size_t i = 0; i < n; i++) {
        const double *p = pts + i * (size_t)d;
        double m = p[0] * r[0];
        for (int k = 1; k < d; k++) {
            double t = p[k] * r[k];
            if (t < m) m = t;
        }
        if (m > best) best = m;
    }
    return best;
}
ORC_API double orc_expected_value(const double *pts, size_t n, int d, const double *r)
{
    return orc_pow_uint(orc_maxmin(pts, n, d, r), (unsigned)d);
}

/* Final scaling (c/hvapprox.c:256-257, 691-692, 863-864): (double)(c_d * (sum / (long double)N)). */
ORC_API double orc_finalize(double sum, int d, uint64_t nsamples)
{
    return (double)(orc_cd[d] * ((long double)sum / (long double)nsamples));
}

/* ==========================================================================================
 * DZ2019-HW (c/hvapprox.c:261-693)
 * ======================================================================================== */

/* I_m = int_0^{pi/2} sin^m (c/hvapprox.c:480-548).  The reference writes 2/3.L, 3*M_PIl/16, ...;
 * the same long doubles come out of  odd m: (m-1)!!/m!! as one correctly rounded division of
 * two exact integers;  even m=2k: RN64(C(2k,k) * pi_l) / 2^(2k+1)  (scaling by 2^j is exact). */
static const long double PI_L = 3.141592653589793238462643383279502884L;
static long double half_pi_integral(int m)
{
    if (m == 1) return 1.0L;
    if (m & 1) {
        uint64_t num = 1, den = 1;
        for (int t = 2; t < m; t += 2) num *= (uint64_t)t;      /* (m-1)!! */
        for (int t = 3; t <= m; t += 2) den *= (uint64_t)t;     /* m!!     */
        /* reduce so that both stay exactly representable and the quotient is the same real */
        uint64_t a = num, b = den;
        while (b) { uint64_t r_ = a % b; a = b; b = r_; }
        return (long double)(num / a) / (long double)(den / a);
    }
    int k = m / 2;
    uint64_t binom = 1;                                          /* C(2k,k) */
    for (int t = 1; t <= k; t++) binom = binom * (uint64_t)(k + t) / (uint64_t)t;
    return ldexpl((long double)binom * PI_L, -(2 * k + 1));
}
ORC_API void orc_half_pi_integral_bytes(int m, void *out16)
{
    long double v = half_pi_integral(m);
    memset(out16, 0, 16);
    memcpy(out16, &v, 10);
}

/* F_m(b) = int_0^b sin^m in double (c/hvapprox.c:367-477 evaluates per-m closed forms in double).
 * Restated through the reduction  F_m = ((m-1)/m) F_{m-2} - sin^{m-1} b cos b / m  unrolled to
 *   even m:  F = C b        - cos b sin b  * H(sin^2 b)
 *   odd  m:  F = C (1-cos b) - cos b sin^2 b * H(sin^2 b)
 * with H a polynomial with positive coefficients (Horner).  For even m this is the shape of the
 * reference's expressions; for odd m the reference expands in powers of cos b, which carries
 * ~1e-16 * sum|coef| of cancellation noise (the cause of its d >= 23 non-termination, SURVEY
 * fact 2).  Both agree to that noise, which is below the Newton tolerance for m <= 18. */
typedef struct { double lead; double h[32]; int nh; } fm_coef;
static fm_coef FM[MAXDIM];
static int fm_ready = 0;
static void fm_init(void)
{
    if (fm_ready) return;
    for (int m = 0; m < MAXDIM; m++) {
        fm_coef *c = &FM[m];
        memset(c, 0, sizeof *c);
        /* coefficients in long double, rounded once to double */
        long double lead = 1.0L;
        long double h[32];
        int nh = 0;
        /* walk the reduction from m down: term for exponent j (j = m, m-2, ...) is
           prod_{t>j, t≡m mod 2} ((t-1)/t) * (1/j) * sin^{j-1} cos */
        for (int j = m; j >= 2; j -= 2) {
            h[nh++] = lead / (long double)j;        /* multiplies sin^{j-1} cos */
            lead *= (long double)(j - 1) / (long double)j;
        }
        c->lead = (double)lead;
        /* store lowest power first: h[nh-1] belongs to the smallest j (2 or 3) */
        c->nh = nh;
        for (int t = 0; t < nh; t++) c->h[t] = (double)h[nh - 1 - t];
    }
    fm_ready = 1;
}
static double int_sin_pow(int m, double b)
{
    const fm_coef *c = &FM[m];
    if (m == 0) return b;
    const double cb = cos(b);
    if (m == 1) return 1.0 - cb;
    const double sb = sin(b);
    const double s2 = sb * sb;
    double H = c->h[c->nh - 1];
    for (int t = c->nh - 2; t >= 0; t--) H = H * s2 + c->h[t];
    if (m & 1) return c->lead * (1.0 - cb) - cb * s2 * H;
    return c->lead * b - cb * sb * H;
}
ORC_API double orc_int_sin_pow(int m, double b) { fm_init(); return int_sin_pow(m, b); }

/* solve_inverse_int_of_power_sin (c/hvapprox.c:551-564): Newton from pi/2 on F_m(x) = target,
 * long double iterate, derivative sin^m via sinl + powl chain, F in double, stop at |f| <= 1e-15.
 * Zero target: the stop angle is an artefact of the reference's rounding noise -> table.
 * The iteration cap (the reference has none) is reported through *capped. */
enum { NEWTON_CAP = 200 };
static long double solve_inverse(long double target, int m, int *capped)
{
    if (target == 0.0L && m < 32 && orc_theta0_bits[m] != UINT64_MAX) {
        double t0;
        memcpy(&t0, &orc_theta0_bits[m], 8);
        return (long double)t0;
    }
    long double x = PI_L / 2;
    long double f = half_pi_integral(m) - target;
    int it = 0;
    while (fabsl(f) > 1e-15) {
        if (++it > NEWTON_CAP) { if (capped) *capped += 1; break; }
        long double g = powl_chain(sinl(x), (unsigned)m);
        x -= f / g;
        f = (long double)int_sin_pow(m, (double)x) - target;
    }
    return x;
}
ORC_API double orc_solve_inverse_u(double u, int m)
{
    fm_init();
    return (double)solve_inverse((long double)u * half_pi_integral(m), m, NULL);
}

/* construct_polar_a (c/hvapprox.c:261-286): generator vector of the rank-1 lattice.
 * p = smallest prime with (p-1)/2 >= lattice dimension (table :266-270, regenerated here). */
static int lattice_prime(int ldim)
{
    if (ldim == 0) return 1;
    for (int p = 3;; p += 2) {
        int is_p = 1;
        for (int q = 3; q * q <= p; q += 2) if (p % q == 0) { is_p = 0; break; }
        if (is_p && (p - 1) / 2 >= ldim) return p;
    }
}
ORC_API void orc_hw_polar_a(int d, uint64_t nsamples, uint64_t *a)
{
    const int ldim = d - 1;
    const int p = lattice_prime(ldim);
    a[0] = 1;
    for (int k = 1; k < ldim; k++) {
        long double t = fabsl(2 * cosl(2 * PI_L * k / p));
        t -= truncl(t);
        a[k] = (uint64_t)llroundl(nsamples * t);
    }
}

/* compute_polar_sample (c/hvapprox.c:288-304) in native x87 long double. */
static void polar_sample(long double *u, int ldim, uint64_t i, uint64_t nsamples, const uint64_t *a)
{
    if (i + 1 < nsamples) {
        long double factor = (i + 1) / (long double)nsamples;
        for (int k = 0; k < ldim; k++) {
            long double v = factor * a[k];
            u[k] = v - truncl(v);
        }
    } else {
        for (int k = 0; k < ldim; k++) u[k] = 0.0L;
    }
}
ORC_API void orc_hw_polar_sample(int d, uint64_t i, uint64_t nsamples, double *u_out)
{
    uint64_t a[MAXDIM];
    long double u[MAXDIM];
    orc_hw_polar_a(d, nsamples, a);
    polar_sample(u, d - 1, i, nsamples, a);
    for (int k = 0; k < d - 1; k++) u_out[k] = (double)u[k];
}

/* One reciprocal direction (c/hvapprox.c:677-680 = compute_polar_sample, compute_theta :566-576,
 * compute_sin_cos_theta :610-619, compute_hua_wang_direction :626-647). */
static void hw_direction(double *r, int d, uint64_t i, uint64_t nsamples, const uint64_t *a, int *capped)
{
    long double u[MAXDIM];
    double s[MAXDIM], c[MAXDIM];
    const int ldim = d - 1;
    polar_sample(u, ldim, i, nsamples, a);
    for (int j = 0; j < ldim; j++) {
        const int m = d - 2 - j;
        long double th = solve_inverse(u[j] * half_pi_integral(m), m, capped);
        double t = (double)th;
        s[j] = sin(t);
        c[j] = cos(t);
    }
    /* w_0 = s_0 s_1 ... s_{d-2};  w_j = c_{d-1-j} * s_0..s_{d-2-j} (1<=j<=d-2);  w_{d-1} = c_0.
       The running product is built as s_0, s_1*(s_0), s_2*(s_1*s_0), ... (c/hvapprox.c:632-637). */
    double prod = s[0];
    r[d - 1] = c[0];
    for (int q = 1; q < ldim; q++) {       /* prod = s_0..s_{q-1}; pairs with c_q -> w_{d-1-q} */
        r[d - 1 - q] = prod * c[q];
        prod = s[q] * prod;
    }
    r[0] = prod;
    for (int j = 0; j < d; j++) r[j] = (fabs(r[j]) <= TINY_WEIGHT) ? 1.0 / TINY_WEIGHT : 1.0 / r[j];
}
ORC_API int orc_hw_directions(int d, uint64_t nsamples, uint64_t i0, uint64_t count, double *r_out)
{
    uint64_t a[MAXDIM];
    int capped = 0;
    fm_init();
    orc_hw_polar_a(d, nsamples, a);
    for (uint64_t j = 0; j < count; j++) hw_direction(r_out + j * (size_t)d, d, i0 + j, nsamples, a, &capped);
    return capped;
}
/* Window of the HW sum over already transformed points; left-to-right double sum like
 * c/hvapprox.c:676-682.  values (optional) receives the per-sample s^d. */
ORC_API double orc_hw_window(const double *pts, size_t n, int d, uint64_t nsamples,
                             uint64_t i0, uint64_t count, double *values)
{
    uint64_t a[MAXDIM];
    double r[MAXDIM];
    double sum = 0.0;
    fm_init();
    orc_hw_polar_a(d, nsamples, a);
    for (uint64_t j = 0; j < count; j++) {
        hw_direction(r, d, i0 + j, nsamples, a, NULL);
        double v = orc_expected_value(pts, n, d, r);
        if (values) values[j] = v;
        sum += v;
    }
    return sum;
}

/* ==========================================================================================
 * DZ2019-MC (c/hvapprox.c:221-258) with the NumPy-legacy MT19937 + ziggurat stream
 * ======================================================================================== */
typedef struct { uint32_t key[624]; int pos; } mt_state;

/* c/mt19937/mt19937.c:6-16 */
static void mt_seed(mt_state *st, uint32_t seed)
{
    for (int i = 0; i < 624; i++) {
        st->key[i] = seed;
        seed = 1812433253u * (seed ^ (seed >> 30)) + (uint32_t)i + 1u;
    }
    st->pos = 624;
}
/* c/mt19937/mt19937.c:18-55: x[i] <- x[i+397 mod 624] ^ twist(upper(x[i]) | lower(x[i+1])) */
static void mt_refill(mt_state *st)
{
    uint32_t *x = st->key;
    for (int i = 0; i < 624; i++) {
        uint32_t y = (x[i] & 0x80000000u) | (x[(i + 1) % 624] & 0x7fffffffu);
        x[i] = x[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    st->pos = 0;
}
/* c/mt19937/mt19937.h:28-44 */
static uint32_t mt_next(mt_state *st)
{
    if (st->pos == 624) mt_refill(st);
    uint32_t y = st->key[st->pos++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}
/* c/mt19937/mt19937.h:46-57: high word first; 53-bit double from 27+26 bits */
static uint64_t mt_next64(mt_state *st) { uint64_t hi = mt_next(st); return hi << 32 | mt_next(st); }
static double mt_next_double(mt_state *st)
{
    int32_t a = (int32_t)(mt_next(st) >> 5), b = (int32_t)(mt_next(st) >> 6);
    return (a * 67108864.0 + b) / 9007199254740992.0;
}
static double bits2d(uint64_t b) { double v; memcpy(&v, &b, 8); return v; }

/* rng_standard_normal (c/rng.c:71-99): 256-layer ziggurat. */
static double zig_normal(mt_state *st)
{
    const double R = bits2d(ORC_ZIG_R_BITS), INV_R = bits2d(ORC_ZIG_INV_R_BITS);
    for (;;) {
        uint64_t r = mt_next64(st);
        int idx = (int)(r & 0xff);
        int neg = (int)((r >> 8) & 1);
        uint64_t rabs = (r >> 9) & 0x000fffffffffffffull;
        double x = (double)rabs * bits2d(orc_zig_wi_bits[idx]);
        if (neg) x = -x;
        if (rabs < orc_zig_ki[idx]) return x;
        if (idx == 0) {
            for (;;) {
                double xx = -INV_R * log1p(-mt_next_double(st));
                double yy = -log1p(-mt_next_double(st));
                if (yy + yy > xx * xx) return ((rabs >> 8) & 1) ? -(R + xx) : R + xx;
            }
        } else {
            double f1 = bits2d(orc_zig_fi_bits[idx - 1]), f0 = bits2d(orc_zig_fi_bits[idx]);
            if ((f1 - f0) * mt_next_double(st) + f0 < exp(-0.5 * x * x)) return x;
        }
    }
}
ORC_API void orc_mt_raw(uint32_t seed, uint64_t count, uint32_t *out)
{
    mt_state st;
    mt_seed(&st, seed);
    for (uint64_t j = 0; j < count; j++) out[j] = mt_next(&st);
}
ORC_API void orc_mc_normals(uint32_t seed, uint64_t count, double *out)
{
    mt_state st;
    mt_seed(&st, seed);
    for (uint64_t j = 0; j < count; j++) out[j] = zig_normal(&st);
}
/* d normals -> reciprocal direction (c/hvapprox.c:239-250, euclidean_norm :203-212). */
static void mc_direction(double *r, int d, mt_state *st)
{
    for (int k = 0; k < d; k++) r[k] = zig_normal(st);
    for (int k = 0; k < d; k++) { double a = fabs(r[k]); r[k] = a < TINY_WEIGHT ? TINY_WEIGHT : a; }
    double nrm = r[0] * r[0] + r[1] * r[1];
    for (int k = 2; k < d; k++) nrm += r[k] * r[k];
    nrm = sqrt(nrm);
    for (int k = 0; k < d; k++) r[k] = nrm / r[k];
}
ORC_API double orc_mc_window(const double *pts, size_t n, int d, uint32_t seed,
                             uint64_t i0, uint64_t count, double *values, double *r_out)
{
    mt_state st;
    double r[MAXDIM];
    double sum = 0.0;
    mt_seed(&st, seed);
    for (uint64_t j = 0; j < i0 + count; j++) {
        mc_direction(r, d, &st);
        if (j < i0) continue;
        if (r_out) memcpy(r_out + (j - i0) * (size_t)d, r, (size_t)d * sizeof(double));
        if (pts) {
            double v = orc_expected_value(pts, n, d, r);
            if (values) values[j - i0] = v;
            sum += v;
        }
    }
    return sum;
}

/* ==========================================================================================
 * Rphi-FWE+ (c/hvapprox.c:695-865)
 * ======================================================================================== */
/* Rphi_init (:695-739): alpha_k = rho^(k+1) by repeated long double multiplication. */
static void rphi_alpha(long double *alpha, int ldim)
{
    alpha[0] = orc_rho[ldim];
    for (int k = 1; k < ldim; k++) alpha[k] = alpha[k - 1] * orc_rho[ldim];
}
/* fang_wang_efficient_mapping_plus (:791-826) */
static void fang_wang_plus(double *w, int d, const double *u)
{
    const double HALF_PI = 1.57079632679489661923;
    double prod = 1.0;
    const int pairs = d / 2 - 1;
    for (int i = 0; i < pairs; i++) {
        const int l = d - 2 * i - 2;
        double t = pow(u[l - 1], 1.0 / l);
        double dl = sqrt(1.0 - t * t) * prod;
        prod *= t;
        double ang = HALF_PI * u[l];
        w[2 * i] = dl * sin(ang);
        w[2 * i + 1] = dl * cos(ang);
    }
    double ang;
    if (d % 2 == 0) {
        ang = HALF_PI * u[0];
    } else {
        w[d - 3] = prod * u[0];
        prod *= sqrt(1 - u[0] * u[0]);
        ang = HALF_PI * u[1];
    }
    w[d - 2] = prod * sin(ang);
    w[d - 1] = prod * cos(ang);
}
ORC_API double orc_rphi_window(const double *pts, size_t n, int d, uint64_t i0, uint64_t count,
                               double *values, double *r_out, double *u_out)
{
    long double alpha[MAXDIM];
    double u[MAXDIM], w[MAXDIM];
    const int ldim = d - 1;
    double sum = 0.0;
    rphi_alpha(alpha, ldim);
    for (int k = 0; k < ldim; k++) u[k] = 0.5;
    for (uint64_t j = 0; j < i0 + count; j++) {
        for (int k = 0; k < ldim; k++) {          /* Rphi_next (:741-747) */
            long double v = u[k] + alpha[k];
            u[k] = (double)(v - truncl(v));
        }
        if (j < i0) continue;
        if (u_out) memcpy(u_out + (j - i0) * (size_t)ldim, u, (size_t)ldim * sizeof(double));
        fang_wang_plus(w, d, u);
        for (int k = 0; k < d; k++) w[k] = 1.0 / w[k];
        if (r_out) memcpy(r_out + (j - i0) * (size_t)d, w, (size_t)d * sizeof(double));
        if (pts) {
            double v = orc_expected_value(pts, n, d, w);
            if (values) values[j - i0] = v;
            sum += v;
        }
    }
    return sum;
}

/* ==========================================================================================
 * The three entry points (c/hvapprox.h:16-38) restated on top of the pieces above.
 * Types: dimension_t = uint_fast8_t (1 byte), boolvec = uint8_t, uint_fast32_t = 8 bytes on
 * x86-64 glibc — written as plain C types here to keep the oracle standalone.
 * ======================================================================================== */
static double *filtered(const double *data, size_t *n, int d, const double *ref, const uint8_t *maximise)
{
    double *pts = malloc((*n ? *n : 1) * (size_t)d * sizeof(double));
    *n = orc_transform_and_filter(data, *n, d, ref, maximise, pts);
    if (*n == 0) { free(pts); return NULL; }
    return pts;
}
ORC_API double orc_hv_approx_hua_wang(const double *data, size_t n, uint8_t d, const double *ref,
                                      const uint8_t *maximise, uint64_t nsamples)
{
    double *pts = filtered(data, &n, d, ref, maximise);
    if (!pts) return 0.0;
    double sum = orc_hw_window(pts, n, d, nsamples, 0, nsamples, NULL);
    free(pts);
    return orc_finalize(sum, d, nsamples);
}
ORC_API double orc_hv_approx_normal(const double *data, size_t n, uint8_t d, const double *ref,
                                    const uint8_t *maximise, uint64_t nsamples, uint32_t seed)
{
    double *pts = filtered(data, &n, d, ref, maximise);
    if (!pts) return 0.0;
    double sum = orc_mc_window(pts, n, d, seed, 0, nsamples, NULL, NULL);
    free(pts);
    return orc_finalize(sum, d, nsamples);
}
ORC_API double orc_hv_approx_rphi_fang_wang_plus(const double *data, size_t n, uint8_t d, const double *ref,
                                                 const uint8_t *maximise, uint64_t nsamples)
{
    double *pts = filtered(data, &n, d, ref, maximise);
    if (!pts) return 0.0;
    double sum = orc_rphi_window(pts, n, d, 0, nsamples, NULL, NULL, NULL);
    free(pts);
    return orc_finalize(sum, d, nsamples);
}


/* ------------------------------------------------------------------------------------------ */
/* epsilon indicator: restatement of c/epsilon.h:52-107 (epsilon_helper_) without its early exits, which only
 * prune: I = max_b min_a max_k v_k, v_k = a_k - b_k | a_k / b_k for a minimised objective, operands swapped
 * for a maximised one; start value 0 (multiplicative) or -inf (additive), c/epsilon.h:83.                    */
/* ------------------------------------------------------------------------------------------ */
__attribute__((visibility("default")))
double orc_epsilon(int mult, const double *data, size_t n, int dim, const double *ref, size_t ref_size, const unsigned char *maximise)
{
    double eps = mult ? 0.0 : -INFINITY;
    for (size_t b = 0; b < ref_size; b++) {
        const double *pb = ref + b * (size_t)dim;
        double eps_min = INFINITY;
        for (size_t a = 0; a < n; a++) {
            const double *pa = data + a * (size_t)dim;
            double eps_max = 0.0;
            for (int k = 0; k < dim; k++) {
                const double x = maximise[k] ? pb[k] : pa[k], y = maximise[k] ? pa[k] : pb[k];
                const double v = mult ? x / y : x - y;
                eps_max = (k == 0 || v > eps_max) ? v : eps_max;
            }
            eps_min = (eps_max < eps_min) ? eps_max : eps_min;
        }
        eps = (eps_min > eps) ? eps_min : eps;
    }
    return eps;
}


/* ------------------------------------------------------------------------------------------ */
/* generational distances: restatement of gd_common_helper_ (c/igd.h:61-127).  (a, r) = (points_a, points_r);
 * IGD = gd(ref, data, plus=0, psize=0, p=1), IGD+ = gd(ref, data, plus=1, psize=1, p=1),
 * avg Hausdorff = max(gd(data, ref, 0, 1, p), gd(ref, data, 0, 1, p))  (c/igd.h:186-194, 232-241, 257-273).   */
/* ------------------------------------------------------------------------------------------ */
__attribute__((visibility("default")))
double orc_gd_common(const double *pa_, size_t size_a, const double *pr_, size_t size_r, int dim, const unsigned char *maximise,
                     int plus, int psize, unsigned p)
{
    if (size_a == 0) return INFINITY;
    double gd = 0.0;
    for (size_t a = 0; a < size_a; a++) {
        const double *pa = pa_ + a * (size_t)dim;
        double min_dist = INFINITY;
        int zero = 0;
        for (size_t r = 0; r < size_r && !zero; r++) {
            const double *pr = pr_ + r * (size_t)dim;
            double dist = 0.0;
            for (int k = 0; k < dim; k++) {
                double diff;
                if (plus) {
                    const double v = maximise[k] ? pa[k] - pr[k] : pr[k] - pa[k];
                    diff = (v < 0.) ? 0. : v;
                } else {
                    diff = pa[k] - pr[k];
                }
                dist += diff * diff;
            }
            if (dist == 0) zero = 1;
            min_dist = (min_dist > dist) ? dist : min_dist;
        }
        if (zero) continue;
        if (p == 1) min_dist = sqrt(min_dist);
        else if (p % 2 == 0) min_dist = orc_pow_uint(min_dist, p / 2);
        else min_dist = orc_pow_uint(sqrt(min_dist), p);
        gd += min_dist;
    }
    if (p == 1) return gd / (double)size_a;
    if (psize) return (double)powl(gd / (double)size_a, 1.0 / p);
    return (double)powl(gd, 1.0 / p) / (double)size_a;
}

/* ------------------------------------------------------------------------------------------ */
/* whv_hype: restatement of c/whv_hype.c (2 objectives).  dist: 0 uniform (:78-92), 1 exponential (:50-76, mu = mu2[0]),
 * 2 bivariate normal around mu2 (:30-48, c/rng.c:101-116 with sigma 0.25, rho 1).  Samples of the unit square from the
 * MT19937 stream, points normalised to [ideal, ref] (c/nondominated.h:878-901), estimate_whv's ordered sum (:152-195)
 * written literally, then Eq. 18 (:236).                                                                      */
/* ------------------------------------------------------------------------------------------ */
__attribute__((visibility("default")))
double orc_whv_hype(int dist, const double *points, int npoints, const double *ideal, const double *ref, int nsamples,
                    uint32_t seed, const double *mu2)
{
    mt_state st;
    mt_seed(&st, seed);
    double diff[2];
    for (int k = 0; k < 2; k++) { diff[k] = ref[k] - ideal[k]; if (diff[k] == 0.0) diff[k] = 1; }
    double *smp = malloc(sizeof(double) * 2 * (size_t)nsamples);
    if (dist == 0) {
        for (int i = 0; i < nsamples; i++)
            for (int k = 0; k < 2; k++) smp[2 * i + k] = 0.0 + mt_next_double(&st) * 1.0;
    } else if (dist == 1) {
        const int half = (int)(0.5 * nsamples);
        const double mu = mu2[0];
        for (int i = 0; i < nsamples; i++) {
            const int expo_axis = i < half ? 0 : 1;
            for (int k = 0; k < 2; k++) {
                const double x = mt_next_double(&st);
                smp[2 * i + k] = (k == expo_axis) ? 0.0 - mu * log(x) : 0.0 + x * 1.0;
            }
        }
    } else {
        double mu[2];
        for (int k = 0; k < 2; k++) mu[k] = 0.0 + 1.0 * (mu2[k] - ideal[k]) / diff[k];
        const double sigma = 0.25, rho = 1.0, nu = sigma * sqrt(1 - rho * rho);
        for (int i = 0; i < nsamples; i++) {
            const double x1 = zig_normal(&st);
            smp[2 * i] = mu[0] + x1 * sigma;
            smp[2 * i + 1] = mu[1] + x1 * (sigma * rho) + nu * zig_normal(&st);
        }
    }
    double *p2 = malloc(sizeof(double) * 2 * (size_t)(npoints ? npoints : 1));
    for (int j = 0; j < 2 * npoints; j++) p2[j] = 0.0 + 1.0 * (points[j] - ideal[j & 1]) / diff[j & 1];
    double whv = 0.0;
    for (int s = 0; s < nsamples; s++) {
        unsigned dominated = 0;
        for (int j = 0; j < npoints; j++)
            if (!(smp[2 * s] < p2[2 * j]) && !(smp[2 * s + 1] < p2[2 * j + 1])) dominated++;
        for (unsigned j = 0; j < dominated; j++) whv += 1.0 / dominated;
    }
    free(smp); free(p2);
    double volume = 1.0;
    for (int k = 0; k < 2; k++) volume *= (ref[k] - ideal[k]);
    whv *= volume / nsamples;
    return whv;
}
