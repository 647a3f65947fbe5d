/* hvb200_internal.h — private interface between the host C layer (hvapprox_b200.c) and the
 * CUDA translation unit (hvb200_kernels.cu).  Plain C; no CUDA types leak to the C side. */
#ifndef HVB200_INTERNAL_H_
#define HVB200_INTERNAL_H_

#include <stddef.h>
#include <stdint.h>
#include "../../include/hvapprox_b200.h"
#include "../../include/epsilon_b200.h"
#include "../../include/whv_hype_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#ifndef HVB200_MAXD
#define HVB200_MAXD HVB200_DIMENSION_MAX
#endif

/* DZ2019-HW lattice parameters, built on the host (construct_polar_a, c/hvapprox.c:261-286). */
typedef struct hvb200_hw_params {
    uint64_t nsamples;                 /* N (<= 2^32 on the device path)          */
    uint64_t a[HVB200_MAXD];           /* generator vector a[0..d-2]              */
    int d;
} hvb200_hw_params;

struct hvb200_plan {
    int device;
    int d;                 /* number of objectives                                              */
    size_t n;              /* n' surviving points (after transform_and_filter and the nondominated filter) */
    size_t n_dominated;    /* points the nondominated pre-filter removed                        */
    int kernel;            /* requested HVB200_KERNEL_*                                         */
    uint64_t run_samples;  /* size of the run being started (lets the device side amortise one-off work) */
    uint64_t run_total;    /* samples of the run being started, whatever its outputs (sizes per-chunk buffers) */
    /* device side (owned by the CUDA TU) */
    void *dev;             /* struct dev_plan*                                                  */
    int want_fused_hw;     /* the run being started is a DZ2019-HW sum: the max-min kernel may generate its directions itself */
    int no_ndfilter;       /* keep dominated points (contribution approximation: they can be the second best) */
    int mc_seek;           /* the next DZ2019-MC run starts at pair mc_pair (hvb200_plan_set_mc_start) and discards mc_discard normals */
    uint64_t mc_pair, mc_discard;
    void *share;           /* preparation shared by the plans of one multi-GPU call (hvb200cu_share_new), or NULL */
    hvb200_stats stats;
};

/* error reporting (thread-local buffer lives in the C TU) */
void hvb200_set_error(const char *fmt, ...);

/* ---- implemented in hvb200_kernels.cu ---- */
int hvb200cu_device_count(void);
/* upload n x d transformed points (host, row-major) */
int hvb200cu_plan_upload(struct hvb200_plan *plan, const double *pts_host);
int hvb200cu_plan_upload_raw(struct hvb200_plan *plan, const double *data_host, size_t npoints,
                             const double *ref, const unsigned char *maximise, size_t *kept_out);
void hvb200cu_plan_free(struct hvb200_plan *plan);

/* A run = begin, then per batch {generate directions, maxmin}, then end. */
int hvb200cu_run_begin(struct hvb200_plan *plan, void *stream, uint64_t max_batch, uint64_t *batch_out);
/* direction producers: fill the plan's direction buffer with `count` samples */
int hvb200cu_gen_hw(struct hvb200_plan *plan, const hvb200_hw_params *hw, uint64_t i0, uint64_t count);
int hvb200cu_gen_hw_at(struct hvb200_plan *plan, const hvb200_hw_params *hw, uint64_t i0, uint64_t count, uint64_t offset);
/* positions [p0, p0 + count) of a shard: position p is lattice index first + ((p >> shift) * stride << shift) + (p & (2^shift - 1)) */
int hvb200cu_gen_hw_mapped(struct hvb200_plan *plan, const hvb200_hw_params *hw, uint64_t first, int shift, uint64_t stride,
                           uint64_t p0, uint64_t count, uint64_t offset);
int hvb200cu_gen_mc(struct hvb200_plan *plan, const double *normals_host, uint64_t count);     /* count*d normals */
/* DZ2019-MC stream generated on the device: begin(seed), skip n normals, then per batch */
int hvb200cu_mc_begin(struct hvb200_plan *plan, uint32_t seed);
int hvb200cu_mc_begin_at(struct hvb200_plan *plan, uint32_t seed, uint64_t pair0);      /* the stream from pair `pair0` on */
int hvb200cu_mc_scan_segment(struct hvb200_plan *plan, uint32_t seed, uint64_t pair0, uint64_t npairs, int want_body, uint64_t out[20]);
int hvb200cu_mc_skip(struct hvb200_plan *plan, uint64_t nnormals);
int hvb200cu_gen_mc_device(struct hvb200_plan *plan, uint64_t count);
int hvb200cu_gen_rphi(struct hvb200_plan *plan, const double *u_host, uint64_t count);         /* count*(d-1) u   */
int hvb200cu_gen_uploaded(struct hvb200_plan *plan, const double *r_host, uint64_t count);     /* count*d r       */
/* DZ2019-HW sums with the generator fused into the pruned max-min kernel (no direction buffer, one launch):
   run_is_fused tells whether run_begin chose that path for the run in progress */
int hvb200cu_run_is_fused(struct hvb200_plan *plan);
int hvb200cu_maxmin_hw_fused(struct hvb200_plan *plan, const hvb200_hw_params *hw, uint64_t first, int shift,
                             uint64_t stride, uint64_t count);
/* consume the direction buffer; values_host (optional) receives per-sample s^d */
int hvb200cu_maxmin(struct hvb200_plan *plan, uint64_t count, double *values_host);
/* copy the generated reciprocal directions of the current batch to the host (count x d) */
int hvb200cu_read_dirs(struct hvb200_plan *plan, uint64_t count, double *r_host);
int hvb200cu_run_end(struct hvb200_plan *plan, double out_hilo[2]);
/* epsilon indicator */
int hvb200cu_epsilon(int device, int mult, const double *A_host, size_t na, const double *B_host, size_t nb, int d,
                     const signed char *mm, double *out);
int hvb200cu_gd_common(int device, const double *X_host, size_t nx, const double *Y_host, size_t ny, int d,
                       const signed char *mm, int plus, int psize, unsigned p, double *result);
/* whv_hype: per-sample dominator counts (2 objectives; pts and samples are n x 2) */
int hvb200cu_whv_counts(int device, const double *pts_host, size_t npoints, const double *samples_host, size_t nsamples,
                        unsigned *counts_host);
/* cached direction sets */
int hvb200cu_dev_alloc(int device, size_t bytes, void **out);
void hvb200cu_dev_free(int device, void *p);
int hvb200cu_copy_dirs_out(struct hvb200_plan *plan, double *dst, uint64_t count, int wait);
int hvb200cu_copy_dev(struct hvb200_plan *plan, double *dst, const double *src, size_t doubles);
void hvb200cu_set_dirs_view(struct hvb200_plan *plan, const double *view);
/* wait for everything queued on the plan's stream (error paths that must not leave readers of a shared buffer behind) */
void hvb200cu_plan_sync(struct hvb200_plan *plan);
/* batched point sets against one direction set */
size_t hvb200cu_sets_max_points(int d);
int hvb200cu_sets_begin(struct hvb200_plan *plan, const double *pts_host, const unsigned *off_host, int nsets, void **ctx_out);
int hvb200cu_sets_maxmin(struct hvb200_plan *plan, void *ctx, uint64_t count);
int hvb200cu_sets_end(struct hvb200_plan *plan, void *ctx, double *hilo_host);

/* host-side preparation of the pruned kernel shared by the plans that one multi-GPU call creates for the same
   point set on several devices (built once, by whichever plan needs it first) */
void *hvb200cu_share_new(void);
void hvb200cu_share_free(void *share);

/* hypervolume-contribution approximation: a per-batch sink like the batched sets (passed as sets_ctx) */
int hvb200cu_nd_strict_flags(int device, const double *pts_host, size_t n, int d, unsigned *flags_host);
int hvb200cu_contrib_begin(struct hvb200_plan *plan, void **ctx_out);
int hvb200cu_contrib_end(struct hvb200_plan *plan, void *ctx, double *contrib_host);

/* pinned staging buffer for host-produced streams (MC normals, Rphi u); grows on demand */
double *hvb200cu_staging(struct hvb200_plan *plan, size_t ndoubles);

/* one NCCL all-gather of the plans' (hi, lo) device results; 1 = NCCL unavailable, 0 = ok, -1 = error */
int hvb200cu_nccl_allgather(struct hvb200_plan **plans, int n, double *pairs_out);

double hvb200cu_measure_fp64(int device, double *dmul_per_s, double *dsetp_per_s);

#ifdef __cplusplus
}
#endif
#endif
