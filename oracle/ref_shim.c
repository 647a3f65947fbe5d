/* oracle/ref_shim.c — TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Compiles the UNMODIFIED reference hot path (c/hvapprox.c, c/rng.c,
 * c/mt19937/mt19937.c under /root/reference) into oracle/_ref/libhvref*.so and
 * exposes its file-static pieces so that parity tests can compare index windows,
 * single directions and single Newton solves against the real reference.
 *
 * The reference sources are #included from where they lie (-I$(REF)/c); nothing
 * is copied into this repository.  Build recipe: oracle/Makefile (target ref).
 *
 * Every export below is a thin trampoline around a reference function; the
 * reference file:line it reaches is quoted next to it.
 */
#define DEBUG 0
#include <stdint.h>
#include "config.h"     /* first on purpose: the d <= 64 build (oracle/Makefile, libhvref_d64.so) puts a patched copy
                           ahead of the reference's on the include path; its include guard then keeps the limit */
#include "ziggurat_constants.h" /* reference c/ziggurat_constants.h (static tables) */
#include "hvapprox.c"   /* reference c/hvapprox.c — brings the three public entry points + statics */
/* mt19937.h leaks the macros N and M (c/mt19937/mt19937.h:12-13). */
#undef N
#undef M

#include <string.h>

#define SHIM_API __attribute__((visibility("default")))

/* c/hvapprox.c:30-66 transform_and_filter.  Writes at most npoints*dim doubles to out,
 * returns the number of surviving points. */
SHIM_API size_t ref_transform_and_filter(const double *data, size_t npoints, int dim,
                                         const double *ref, const boolvec *maximise, double *out)
{
    size_t n = npoints;
    double *p = transform_and_filter(data, &n, (dimension_t)dim, ref, maximise);
    if (!p) return 0;
    memcpy(out, p, n * (size_t)dim * sizeof(double));
    free(p);
    return n;
}

/* c/hvapprox.c:68-117 get_expected_value: pow_uint(max_p min_k p_k*w_k, dim). */
SHIM_API double ref_expected_value(const double *points, size_t npoints, int dim, const double *w)
{
    return get_expected_value(points, npoints, (dimension_t)dim, w);
}

/* c/hvapprox.c:261-286 construct_polar_a(dim-1, nsamples) -> a[0..dim-2]. */
SHIM_API void ref_hw_polar_a(int dim, uint64_t nsamples, uint64_t *a_out)
{
    uint_fast32_t *a = construct_polar_a((dimension_t)(dim - 1), (uint_fast32_t)nsamples);
    for (int k = 0; k < dim - 1; k++) a_out[k] = (uint64_t)a[k];
    free(a);
}

/* c/hvapprox.c:288-304 compute_polar_sample: u_k as long double (16 bytes each, x87 layout)
 * and rounded to double. */
SHIM_API void ref_hw_polar_sample(int dim, uint64_t i, uint64_t nsamples,
                                  long double *u_ld_out, double *u_out)
{
    uint_fast32_t *a = construct_polar_a((dimension_t)(dim - 1), (uint_fast32_t)nsamples);
    long double u[MOOCORE_HVAPPROX_DIMENSION_MAX];
    compute_polar_sample(u, (dimension_t)(dim - 1), (uint_fast32_t)i, (uint_fast32_t)nsamples, a);
    for (int k = 0; k < dim - 1; k++) {
        if (u_ld_out) u_ld_out[k] = u[k];
        if (u_out) u_out[k] = (double)u[k];
    }
    free(a);
}

/* c/hvapprox.c:551-564 solve_inverse_int_of_power_sin(theta, m): target given as long double. */
SHIM_API long double ref_solve_inverse_ld(long double target, int m)
{
    return solve_inverse_int_of_power_sin(target, (dimension_t)m);
}
/* Same, target = u * I_m with u given as a double (exact for u == 0 and for tests). */
SHIM_API double ref_solve_inverse_u(double u, int m)
{
    long double t = (long double)u * int_power_of_sin_from_0_to_half_pi[m];
    return (double)solve_inverse_int_of_power_sin(t, (dimension_t)m);
}
/* c/hvapprox.c:367-477 int_of_power_of_sin_from_0_to_b. */
SHIM_API double ref_int_sin_pow(int m, double b)
{
    return int_of_power_of_sin_from_0_to_b((dimension_t)m, b);
}
/* c/hvapprox.c:480-548 table I_m as double / long double. */
SHIM_API long double ref_int_sin_pow_half_pi(int m)
{
    return int_power_of_sin_from_0_to_half_pi[m];
}
/* c/hvapprox.c:168-201 c_d table. */
SHIM_API long double ref_sphere_const(int d)
{
    return sphere_area_div_2_pow_d_times_d[d];
}

/* Reciprocal directions r[count][dim] of DZ2019-HW for indices i0..i0+count-1 of an
 * nsamples-point lattice: c/hvapprox.c:676-680 (the loop body minus get_expected_value). */
SHIM_API void ref_hw_directions(int dim, uint64_t nsamples, uint64_t i0, uint64_t count, double *r_out)
{
    const dimension_t d = (dimension_t)dim;
    const long double *int_all = compute_int_all((dimension_t)(d - 1));
    const uint_fast32_t *a = construct_polar_a((dimension_t)(d - 1), (uint_fast32_t)nsamples);
    long double theta[MOOCORE_HVAPPROX_DIMENSION_MAX];
    double st[MOOCORE_HVAPPROX_DIMENSION_MAX], ct[MOOCORE_HVAPPROX_DIMENSION_MAX];
    for (uint64_t j = 0; j < count; j++) {
        compute_polar_sample(theta, (dimension_t)(d - 1), (uint_fast32_t)(i0 + j), (uint_fast32_t)nsamples, a);
        compute_theta(theta, d, int_all);
        compute_sin_cos_theta(theta, (dimension_t)(d - 1), st, ct);
        compute_hua_wang_direction(r_out + j * (size_t)dim, d, st, ct);
    }
    free((void *)int_all);
    free((void *)a);
}

/* Window of the DZ2019-HW sum: values[j] = get_expected_value(direction i0+j) (may be NULL),
 * returns the reference's left-to-right double sum over the window (c/hvapprox.c:676-682).
 * `points` are already transformed (ref - p > 0). */
SHIM_API double ref_hw_window(const double *points, size_t npoints, int dim, uint64_t nsamples,
                              uint64_t i0, uint64_t count, double *values)
{
    const dimension_t d = (dimension_t)dim;
    const long double *int_all = compute_int_all((dimension_t)(d - 1));
    const uint_fast32_t *a = construct_polar_a((dimension_t)(d - 1), (uint_fast32_t)nsamples);
    long double theta[MOOCORE_HVAPPROX_DIMENSION_MAX];
    double st[MOOCORE_HVAPPROX_DIMENSION_MAX], ct[MOOCORE_HVAPPROX_DIMENSION_MAX], w[MOOCORE_HVAPPROX_DIMENSION_MAX + 1];
    double expected = 0.0;
    for (uint64_t j = 0; j < count; j++) {
        compute_polar_sample(theta, (dimension_t)(d - 1), (uint_fast32_t)(i0 + j), (uint_fast32_t)nsamples, a);
        compute_theta(theta, d, int_all);
        compute_sin_cos_theta(theta, (dimension_t)(d - 1), st, ct);
        compute_hua_wang_direction(w, d, st, ct);
        double v = get_expected_value(points, npoints, d, w);
        if (values) values[j] = v;
        expected += v;
    }
    free((void *)int_all);
    free((void *)a);
    return expected;
}

/* First `count` standard normals of the MC stream for `seed`: c/rng.h:7-13 + c/rng.c:71-99. */
SHIM_API void ref_mc_normals(uint32_t seed, uint64_t count, double *out)
{
    rng_state *rng = rng_new(seed);
    for (uint64_t j = 0; j < count; j++) out[j] = rng_standard_normal(rng);
    free(rng);
}
/* Raw tempered 32-bit outputs of the generator (c/mt19937/mt19937.h:28-44). */
SHIM_API void ref_mt_raw(uint32_t seed, uint64_t count, uint32_t *out)
{
    rng_state *rng = rng_new(seed);
    for (uint64_t j = 0; j < count; j++) out[j] = mt19937_next32(rng);
    free(rng);
}

/* DZ2019-MC directions (reciprocal form) and window sums: c/hvapprox.c:235-251.
 * Samples 0..i0-1 are generated and discarded so that the stream position is the reference's. */
SHIM_API double ref_mc_window(const double *points, size_t npoints, int dim, uint32_t seed,
                              uint64_t i0, uint64_t count, double *values, double *r_out)
{
    const dimension_t d = (dimension_t)dim;
    rng_state *rng = rng_new(seed);
    double w[MOOCORE_HVAPPROX_DIMENSION_MAX + 1];
    double expected = 0.0;
    for (uint64_t j = 0; j < i0 + count; j++) {
        for (dimension_t k = 0; k < d; k++) w[k] = rng_standard_normal(rng);
        if (j < i0) continue;
        for (dimension_t k = 0; k < d; k++) { w[k] = fabs(w[k]); w[k] = MAX(w[k], ALMOST_ZERO_WEIGHT); }
        double norm = euclidean_norm(w, d);
        for (dimension_t k = 0; k < d; k++) w[k] = norm / w[k];
        if (r_out) memcpy(r_out + (j - i0) * (size_t)dim, w, (size_t)dim * sizeof(double));
        if (points) {
            double v = get_expected_value(points, npoints, d, w);
            if (values) values[j - i0] = v;
            expected += v;
        }
    }
    free(rng);
    return expected;
}

/* Rphi-FWE+ window: c/hvapprox.c:844-858.  The recurrence is run from index 0. */
SHIM_API double ref_rphi_window(const double *points, size_t npoints, int dim,
                                uint64_t i0, uint64_t count, double *values, double *r_out, double *u_out)
{
    const dimension_t d = (dimension_t)dim;
    const long double *alpha = Rphi_init((dimension_t)(d - 1));
    double w[MOOCORE_HVAPPROX_DIMENSION_MAX + 1], u[MOOCORE_HVAPPROX_DIMENSION_MAX];
    for (dimension_t k = 0; k < d - 1; k++) u[k] = 0.5;
    double expected = 0.0;
    for (uint64_t j = 0; j < i0 + count; j++) {
        Rphi_next(u, (dimension_t)(d - 1), alpha);
        if (j < i0) continue;
        if (u_out) memcpy(u_out + (j - i0) * (size_t)(dim - 1), u, (size_t)(dim - 1) * sizeof(double));
        fang_wang_efficient_mapping_plus(w, d, u);
        for (dimension_t k = 0; k < d; k++) w[k] = 1. / w[k];
        if (r_out) memcpy(r_out + (j - i0) * (size_t)dim, w, (size_t)dim * sizeof(double));
        if (points) {
            double v = get_expected_value(points, npoints, d, w);
            if (values) values[j - i0] = v;
            expected += v;
        }
    }
    free((void *)alpha);
    return expected;
}
/* Rphi alpha vector as long double (c/hvapprox.c:695-739). */
SHIM_API void ref_rphi_alpha(int dim, long double *alpha_out)
{
    long double *alpha = Rphi_init((dimension_t)(dim - 1));
    for (int k = 0; k < dim - 1; k++) alpha_out[k] = alpha[k];
    free(alpha);
}

/* c/pow_int.h:59-74 */
SHIM_API double ref_pow_uint(double x, unsigned e) { return pow_uint(x, e); }

/* ABI facts used by tests. */
SHIM_API int ref_sizeof_dimension_t(void) { return (int)sizeof(dimension_t); }
SHIM_API int ref_sizeof_uint_fast32_t(void) { return (int)sizeof(uint_fast32_t); }
SHIM_API int ref_sizeof_boolvec(void) { return (int)sizeof(boolvec); }
SHIM_API int ref_dimension_max(void) { return MOOCORE_HVAPPROX_DIMENSION_MAX; }

/* Ziggurat tables (c/ziggurat_constants.h:66,154,284,1264-1266) — accessors so that
 * oracle/gen_ziggurat_tables.py can verify its own regenerated tables bit-for-bit. */
SHIM_API uint64_t ref_zig_ki(int i) { return ki_double[i]; }
SHIM_API double ref_zig_wi(int i) { return wi_double[i]; }
SHIM_API double ref_zig_fi(int i) { return fi_double[i]; }
SHIM_API double ref_zig_r(void) { return ziggurat_nor_r; }
SHIM_API double ref_zig_inv_r(void) { return ziggurat_nor_inv_r; }

/* Raw x87 bytes of long double table entries (ctypes' c_longdouble round-trips through
 * Python float and would lose the low 11 bits). */
SHIM_API void ref_sphere_const_bytes(int d, void *out16)
{
    long double v = sphere_area_div_2_pow_d_times_d[d];
    memset(out16, 0, 16);
    memcpy(out16, &v, 10);
}
SHIM_API void ref_rphi_alpha_bytes(int dim, void *out16xdm1)
{
    long double *alpha = Rphi_init((dimension_t)(dim - 1));
    memset(out16xdm1, 0, 16 * (size_t)(dim - 1));
    for (int k = 0; k < dim - 1; k++) memcpy((char *)out16xdm1 + 16 * k, &alpha[k], 10);
    free(alpha);
}
SHIM_API void ref_int_sin_pow_half_pi_bytes(int m, void *out16)
{
    long double v = int_power_of_sin_from_0_to_half_pi[m];
    memset(out16, 0, 16);
    memcpy(out16, &v, 10);
}
