/* include/hvapprox.h — the drop-in boundary.
 *
 * Declares exactly the three entry points of moocore's c/hvapprox.h:16-38, with the same
 * names, argument order, argument meaning and ABI types, so that a consumer that used to
 * compile c/hvapprox.c into itself (python/src/moocore/_ffi_build.py:39, r/src/Makevars:11,
 * c/Makefile:195) can link libmoocore_b200.so instead.
 *
 * ABI types (c/config.h:24-39; measured on x86-64 glibc):
 *   dimension_t   = uint_fast8_t   (1 byte)
 *   boolvec       = uint8_t, or int when compiled with -DR_PACKAGE (R logical vectors)
 *   uint_fast32_t = 8 bytes
 *
 * Behaviour kept from the reference:
 *   - data is row-major npoints x nobjs, caller-owned, read-only for the duration of the call;
 *   - returns 0.0 iff no point strictly dominates ref (c/hvapprox.c:228-229);
 *   - 2 <= nobjs <= MOOCORE_HVAPPROX_DIMENSION_MAX is a caller-enforced precondition.
 * Behaviour the reference does not define (it has no failure mode): if the CUDA path cannot
 * run (no device, out of memory, kernel fault) the functions print a diagnostic to stderr and
 * return NaN.  There is no CPU fallback.
 */
#ifndef HV_APPROX_H_
#define HV_APPROX_H_

#include <stddef.h>
#include <stdint.h>

#ifndef _MOOCORE_CONFIG_H_           /* stand-alone use: same typedefs as c/config.h */
typedef uint_fast8_t dimension_t;
# ifdef R_PACKAGE
typedef int boolvec;
# else
typedef uint8_t boolvec;
# endif
# define MOOCORE_HVAPPROX_DIMENSION_MAX 31
#endif

#ifndef MOOCORE_API
# if defined(__GNUC__) && __GNUC__ >= 4
#  define MOOCORE_API extern __attribute__((visibility("default")))
# else
#  define MOOCORE_API extern
# endif
#endif

#ifdef __cplusplus
# define HVB200_RESTRICT __restrict__
extern "C" {
#else
# define HVB200_RESTRICT restrict
#endif

/* DZ2019-MC: replaces c/hvapprox.c:221-258 (decl c/hvapprox.h:16-19). */
MOOCORE_API double hv_approx_normal(
    const double * HVB200_RESTRICT data, size_t npoints, dimension_t nobjs,
    const double * HVB200_RESTRICT ref, const boolvec * HVB200_RESTRICT maximise,
    uint_fast32_t nsamples, uint32_t random_seed);

/* DZ2019-HW: replaces c/hvapprox.c:656-693 (decl c/hvapprox.h:25-28). */
MOOCORE_API double hv_approx_hua_wang(
    const double * HVB200_RESTRICT data, size_t npoints, dimension_t nobjs,
    const double * HVB200_RESTRICT ref, const boolvec * HVB200_RESTRICT maximise,
    uint_fast32_t nsamples);

/* Rphi-FWE+: replaces c/hvapprox.c:834-865 (decl c/hvapprox.h:35-38). */
MOOCORE_API double hv_approx_rphi_fang_wang_plus(
    const double * HVB200_RESTRICT data, size_t npoints, dimension_t nobjs,
    const double * HVB200_RESTRICT ref, const boolvec * HVB200_RESTRICT maximise,
    uint_fast32_t nsamples);

#ifdef __cplusplus
}
#endif
#endif /* HV_APPROX_H_ */
