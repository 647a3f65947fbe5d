/* oracle/ref_eps_shim.c — TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Compiles the UNMODIFIED reference epsilon and IGD-family indicators (c/epsilon.h, c/igd.h under /root/reference,
 * header-only) into
 * oracle/_ref/libepsref.so.  The reference header is #included from where it lies (-I$(REF)/c); nothing is
 * copied into this repository.  Build recipe: oracle/Makefile (target ref).
 */
#define DEBUG 0
#include "epsilon.h"    /* reference c/epsilon.h */
#include "igd.h"        /* reference c/igd.h */

#define SHIM_API __attribute__((visibility("default")))

/* c/epsilon.h:197-208 */
SHIM_API double ref_epsilon_additive(const double *data, size_t n, int dim, const double *ref, size_t ref_size, const boolvec *maximise)
{
    return epsilon_additive(data, n, (dimension_t)dim, ref, ref_size, maximise);
}
/* c/epsilon.h:210-221 */
SHIM_API double ref_epsilon_mult(const double *data, size_t n, int dim, const double *ref, size_t ref_size, const boolvec *maximise)
{
    return epsilon_mult(data, n, (dimension_t)dim, ref, ref_size, maximise);
}

/* c/igd.h:196-207, 243-254, 276-287 */
SHIM_API double ref_IGD(const double *data, size_t n, int dim, const double *ref, size_t ref_size, const boolvec *maximise)
{
    return IGD(data, n, (dimension_t)dim, ref, ref_size, maximise);
}
SHIM_API double ref_IGD_plus(const double *data, size_t n, int dim, const double *ref, size_t ref_size, const boolvec *maximise)
{
    return IGD_plus(data, n, (dimension_t)dim, ref, ref_size, maximise);
}
SHIM_API double ref_avg_Hausdorff_dist(const double *data, size_t n, int dim, const double *ref, size_t ref_size, const boolvec *maximise, unsigned p)
{
    return avg_Hausdorff_dist(data, n, (dimension_t)dim, ref, ref_size, maximise, p);
}
