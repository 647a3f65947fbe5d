/* include/whv_hype_b200.h — the HypE weighted-hypervolume estimator of moocore (2 objectives), served by libmoocore_b200.so.
 *
 * Same names, argument order and ABI types as the reference (c/whv_hype.h:6-14, bound by CFFI through
 * python/src/moocore/libmoocore.h): `points` is npoints x 2 row-major FP64 (minimisation), `ideal` and `ref` bound the
 * sampled box.  The estimator draws nsamples points of the unit square from the reference's MT19937 stream
 * (uniform / half exponential in one objective + uniform in the other / bivariate normal around `mu`,
 * c/whv_hype.c:30-92), counts for each sample the points that dominate it (the O(nsamples * npoints) part, on the GPU),
 * adds 1/count once per dominator in the reference's order (:152-195, replayed on the host in closed form per binade)
 * and scales by the box volume / nsamples (:236): bit-identical to the reference.
 * On failure the functions print to stderr and return NaN — there is no CPU fallback.
 */
#ifndef WHV_HYPE_B200_H_
#define WHV_HYPE_B200_H_

#include <stdint.h>
#include "hvapprox.h"

#ifdef __cplusplus
extern "C" {
#endif

MOOCORE_API double whv_hype_unif(const double *points, int npoints, const double *ideal, const double *ref,
                                 int nsamples, uint32_t seed);
MOOCORE_API double whv_hype_expo(const double *points, int npoints, const double *ideal, const double *ref,
                                 int nsamples, uint32_t seed, double mu);
MOOCORE_API double whv_hype_gaus(const double *points, int npoints, const double *ideal, const double *ref,
                                 int nsamples, uint32_t seed, const double *mu);

/* Diagnostics (tests): the ordered sum of estimate_whv (c/whv_hype.c:183-189) over per-sample dominator counts —
 * for every sample, 1.0/count added count times to one running double.  naive != 0 runs that loop literally,
 * naive == 0 the O(nsamples) closed form (within a binade every addition of the same increment moves the sum by
 * the same multiple of its ulp); both return the same bits. */
MOOCORE_API double hvb200_whv_ordered_sum(const unsigned *counts, size_t nsamples, int naive);

#ifdef __cplusplus
}
#endif
#endif
