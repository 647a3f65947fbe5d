/* include/epsilon_b200.h — the epsilon and IGD-family indicator entry points of moocore, served by libmoocore_b200.so.
 *
 * Same names, argument order and ABI types as the reference (c/epsilon.h:197-221, bound by CFFI through
 * python/src/moocore/libmoocore.h:24-25): `data` is n x dim and `ref` ref_size x dim, row-major FP64;
 * maximise[k] selects the orientation of objective k.  Returns
 *
 *     I_eps(data, ref) = max_{b in ref} min_{a in data} max_k v_k(a, b)
 *
 * with v_k = a_k - b_k (additive) or a_k / b_k (multiplicative) for a minimised objective and the operands
 * swapped for a maximised one — bit-identical to the reference (one IEEE operation per v_k, exact max/min).
 * epsilon_mult requires strictly positive values (the reference's callers check that, _moocore.py:287-295).
 * The computation runs on the GPU (device HVB200_DEVICE, default 0); on failure the functions print to stderr
 * and return NaN — there is no CPU fallback.
 */
#ifndef EPSILON_B200_H_
#define EPSILON_B200_H_

#include "hvapprox.h"

#ifdef __cplusplus
extern "C" {
#endif

MOOCORE_API double epsilon_additive(const double *HVB200_RESTRICT data, size_t n, dimension_t dim,
                                    const double *HVB200_RESTRICT ref, size_t ref_size, const boolvec *HVB200_RESTRICT maximise);
MOOCORE_API double epsilon_mult(const double *HVB200_RESTRICT data, size_t n, dimension_t dim,
                                const double *HVB200_RESTRICT ref, size_t ref_size, const boolvec *HVB200_RESTRICT maximise);

/* Inverted generational distance, IGD+ and the averaged Hausdorff distance: c/igd.h:196-287, bound by CFFI through
 * python/src/moocore/libmoocore.h:20-22.  The O(n * ref_size * dim) nearest-point search runs on the GPU with the
 * reference's roundings, the square roots / powers and the ordered sum are replayed on the host: bit-identical. */
MOOCORE_API double IGD(const double *HVB200_RESTRICT data, size_t npoints, dimension_t nobj,
                       const double *HVB200_RESTRICT ref, size_t ref_size, const boolvec *HVB200_RESTRICT maximise);
MOOCORE_API double IGD_plus(const double *HVB200_RESTRICT data, size_t npoints, dimension_t nobj,
                            const double *HVB200_RESTRICT ref, size_t ref_size, const boolvec *HVB200_RESTRICT maximise);
MOOCORE_API double avg_Hausdorff_dist(const double *HVB200_RESTRICT data, size_t npoints, dimension_t nobj,
                                      const double *HVB200_RESTRICT ref, size_t ref_size, const boolvec *HVB200_RESTRICT maximise,
                                      unsigned int p);

#ifdef __cplusplus
}
#endif
#endif
