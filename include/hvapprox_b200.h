/* include/hvapprox_b200.h — extended C-ABI of libmoocore_b200.so.
 *
 * Everything here is an ADDITION to the reference interface (the reference has none of it);
 * the three drop-in entry points live in include/hvapprox.h.  The additions exist for
 *   - multi-GPU sharding: a sample index range [i0, i1) of one estimate can be evaluated on
 *     one device and combined later (the reference's loop c/hvapprox.c:235,676,851 is the
 *     range [0, nsamples));
 *   - measurement: keep the point set resident in HBM across calls, time the kernels;
 *   - parity tests: per-sample values and the generated direction set.
 * Plain C types only (no torch, no CUDA types; a CUDA stream is passed as void*).
 *
 * Every function returning int returns 0 on success and a negative code on failure;
 * hvb200_last_error() then describes the failure (thread-local).
 */
#ifndef HVAPPROX_B200_H_
#define HVAPPROX_B200_H_

#include "hvapprox.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Direction-set generators; numbering follows c/main-hvapprox.c's -m option. */
enum {
    HVB200_METHOD_MC   = 1,   /* DZ2019-MC   hv_approx_normal                (c/hvapprox.c:221) */
    HVB200_METHOD_HW   = 2,   /* DZ2019-HW   hv_approx_hua_wang              (c/hvapprox.c:656) */
    HVB200_METHOD_RPHI = 3    /* Rphi-FWE+   hv_approx_rphi_fang_wang_plus   (c/hvapprox.c:834) */
};

/* Max-min kernel variants.  Both give bit-identical per-sample values. */
enum {
    HVB200_KERNEL_AUTO   = 0,
    HVB200_KERNEL_EXACT  = 1, /* d DMUL + d min + 1 max per sample*point (the textbook loop)   */
    HVB200_KERNEL_SCREEN = 2, /* exact threshold screen: 1 DSETP per element, products only for
                                 candidates that can raise the running max                      */
    HVB200_KERNEL_DPX    = 3, /* the same screen on conservatively quantised 16-bit keys with
                                 DPX VIADDMNMX.S16x2 (2 elements per instruction, no predicates) */
    HVB200_KERNEL_BP     = 4  /* block-pruned: blocks of 32 points are discarded by their
                                 per-objective maxima before any of their points is looked at     */
};

/* Largest dimension the library accepts.  The reference stops at 31
 * (MOOCORE_HVAPPROX_DIMENSION_MAX, c/config.h:28); 32..HVB200_DIMENSION_MAX is an extension
 * (c_d from the closed form, generic square-and-multiply for s^d) offered for DZ2019-MC and
 * Rphi-FWE+ only and only through this header, never through include/hvapprox.h. */
#define HVB200_DIMENSION_MAX 64

typedef struct hvb200_plan hvb200_plan;   /* opaque: transformed point set resident on one GPU */

typedef struct hvb200_stats {
    double maxmin_ms;      /* sum of device time of the max-min kernel launches of the last run */
    double dirgen_ms;      /* same for the direction generators                                  */
    double reduce_ms;      /* final reduction kernel                                             */
    double h2d_ms;         /* host->device copies issued by the last run (MC/Rphi streams)       */
    uint64_t maxmin_launches;
    uint64_t dirgen_launches;
    uint64_t reduce_launches;
    uint64_t samples;      /* i1 - i0 of the last run                                            */
    uint64_t npoints;      /* n' (points that strictly dominate ref)                             */
    uint64_t newton_capped;/* DZ2019-HW Newton solves stopped by the iteration cap               */
    uint64_t slow_path;    /* SCREEN kernel: candidate evaluations (diagnostic)                  */
    uint64_t h2d_bytes;    /* bytes copied host->device by the last run                          */
    int kernel;            /* HVB200_KERNEL_* actually used                                      */
} hvb200_stats;

MOOCORE_API const char *hvb200_last_error(void);
MOOCORE_API int hvb200_device_count(void);
MOOCORE_API const char *hvb200_version(void);

/* transform_and_filter (c/hvapprox.c:30-66) + upload.  *plan_out is NULL (and 0 is returned)
 * when no point strictly dominates ref — the caller then reports 0.0 like c/hvapprox.c:228. */
MOOCORE_API int hvb200_plan_create(hvb200_plan **plan_out, const double *data, size_t npoints,
                                   dimension_t nobjs, const double *ref, const boolvec *maximise,
                                   int device);
MOOCORE_API void hvb200_plan_destroy(hvb200_plan *plan);
MOOCORE_API size_t hvb200_plan_npoints(const hvb200_plan *plan);
/* Points removed by the nondominated pre-filter at plan creation (a point weakly dominated by another one can never
 * attain max_p min_k p'_k r_k, so the estimate keeps every bit; hvb200_plan_npoints counts the points that remain).
 * By default it runs for 512 <= n' <= 2^18 when a 256-point sample shows that at least 1/16 of the points are dominated;
 * HVB200_NDFILTER=0|1 forces it off / on. */
MOOCORE_API size_t hvb200_plan_ndominated(const hvb200_plan *plan);
MOOCORE_API int hvb200_plan_set_kernel(hvb200_plan *plan, int kernel);

/* Sum over samples i in [i0, i1) of (max_p min_k p'_k r^(i)_k)^d for the direction set of
 * `method` with `nsamples` directions in total, as a double-double (hi, lo).  `cuda_stream`
 * is a cudaStream_t (NULL = the plan's own stream).  Synchronous: returns when the result is
 * on the host. */
MOOCORE_API int hvb200_plan_run(hvb200_plan *plan, int method, uint_fast32_t nsamples, uint32_t seed,
                                uint64_t i0, uint64_t i1, void *cuda_stream, double out_hilo[2]);
/* The same sum over the direction blocks b = first, first + stride, first + 2 stride, ... where block b holds
 * the samples [b*block, min((b+1)*block, nsamples)): the block-cyclic shard of rank `first` among `stride`
 * ranks.  DZ2019-HW only (the lattice is index-addressable; its per-direction cost drifts along the index,
 * so interleaved blocks balance the ranks where contiguous ranges do not). */
#define HVB200_SHARD_BLOCK ((uint64_t)1 << 14)
/* the block size the library itself uses for `world` ranks: HVB200_SHARD_BLOCK, halved (down to 2^10) until every
 * rank gets at least 4 blocks */
MOOCORE_API uint64_t hvb200_shard_block(uint_fast32_t nsamples, int world);
MOOCORE_API int hvb200_plan_run_blocks(hvb200_plan *plan, int method, uint_fast32_t nsamples, uint32_t seed,
                                       uint64_t block, uint64_t first, uint64_t stride, void *cuda_stream,
                                       double out_hilo[2]);
MOOCORE_API int hvb200_plan_last_stats(const hvb200_plan *plan, hvb200_stats *out);

/* DZ2019-MC shards that do not parse each other's part of the stream.  A sample's position in the MT19937 stream is data
 * dependent (every rejected ziggurat draw shifts what follows, c/rng.c:71-99), so a contiguous sample range starting at
 * i0 > 0 costs a parse of the whole prefix.  Instead rank g owns the PAIR range [pair0, pair0 + npairs) of the stream
 * (one pair = two tempered words = one next64 of c/mt19937/mt19937.h:46-49; reached by MT19937 jump-ahead), describes
 * it with hvb200_plan_mc_segment_scan, the ranks all-gather the descriptions (160 bytes each), and
 * hvb200_mc_segments_resolve gives every rank the pair at which its parse starts, the normals it drops to reach a sample
 * boundary and the samples [i0, i1) it evaluates — a partition of [0, nsamples) whose per-sample values are those of the
 * sequential stream bit for bit.  Protocol per rank:
 *     hvb200_mc_segment_layout(nsamples, d, world, rank, &p0, &np)   (returns 1: too short to be worth it, use [i0,i1) ranges)
 *     hvb200_plan_mc_segment_scan(plan, seed, p0, np, &mine);  all-gather -> segs[world]
 *     hvb200_mc_segments_resolve(segs, world, d, nsamples, starts)
 *     hvb200_plan_set_mc_start(plan, starts[rank].pair, starts[rank].discard)
 *     hvb200_plan_run(plan, HVB200_METHOD_MC, nsamples, seed, starts[rank].i0, starts[rank].i1, ...)
 * The last rank's range is open-ended (npairs = UINT64_MAX): nobody depends on its description, so its scan only
 * looks at the head. */
typedef struct hvb200_mc_segment {
    uint64_t pair0, npairs;
    uint64_t merge;        /* first pair (relative to pair0) at which the parses of all 16 incoming skip states start an attempt */
    uint64_t head[16];     /* normals emitted by attempts starting in [pair0, pair0 + merge) for incoming skip state s          */
    uint64_t body;         /* normals emitted by attempts starting in [pair0 + merge, pair0 + npairs)                            */
    uint64_t skip_out;     /* pairs of the next segment that this segment's last attempt consumes                                */
} hvb200_mc_segment;
typedef struct hvb200_mc_start { uint64_t pair, discard, i0, i1; } hvb200_mc_start;
MOOCORE_API int hvb200_mc_segment_layout(uint_fast32_t nsamples, dimension_t nobjs, int world, int rank,
                                         uint64_t *pair0, uint64_t *npairs);
MOOCORE_API int hvb200_plan_mc_segment_scan(hvb200_plan *plan, uint32_t seed, uint64_t pair0, uint64_t npairs,
                                            hvb200_mc_segment *out);
MOOCORE_API int hvb200_mc_segments_resolve(const hvb200_mc_segment *segs, int world, dimension_t nobjs,
                                           uint_fast32_t nsamples, hvb200_mc_start *starts);
/* one-shot: applies to the NEXT run / sample_values call of the plan if that is a DZ2019-MC run with the device generator,
 * and is dropped by any other run */
MOOCORE_API int hvb200_plan_set_mc_start(hvb200_plan *plan, uint64_t pair, uint64_t discard);

/* (double)(c_d * (sum / (long double)nsamples)), c/hvapprox.c:256-257,691-692,863-864; the
 * `npairs` (hi, lo) partial sums are added in index order in long double. */
MOOCORE_API double hvb200_finalize(dimension_t nobjs, uint_fast32_t nsamples,
                                   const double *hilo_pairs, int npairs);

/* Parity/debug: per-sample values s_i^d for i in [i0, i0+count) copied to host memory. */
MOOCORE_API int hvb200_plan_sample_values(hvb200_plan *plan, int method, uint_fast32_t nsamples,
                                          uint32_t seed, uint64_t i0, uint64_t count, double *values_out);
/* Parity/debug: the reciprocal directions r^(i) (count x nobjs, row-major) as generated on
 * the device for method/nsamples/seed. */
MOOCORE_API int hvb200_directions(int method, dimension_t nobjs, uint_fast32_t nsamples, uint32_t seed,
                                  uint64_t i0, uint64_t count, double *r_out, int device);
/* Max-min over caller-supplied reciprocal directions (count x nobjs, host memory):
 * values_out (optional) gets s_i^d, out_hilo (optional) the sum. */
MOOCORE_API int hvb200_plan_run_directions(hvb200_plan *plan, const double *r_host, uint64_t count,
                                           double *values_out, double out_hilo[2]);

/* Host-only parity/debug: the Rphi-FWE+ u-sequence (c/hvapprox.c:741-747) of samples [i0, i0+count),
 * count x (nobjs-1) doubles row-major, as produced by the library (one host thread per objective). */
MOOCORE_API int hvb200_rphi_sequence(dimension_t nobjs, uint64_t i0, uint64_t count, double *u_out);

/* Host-only parity/debug of the MT19937 jump-ahead that parallelises the DZ2019-MC stream generator
 * (moocore_b200/csrc/hvb200_mtjump.c): poly = t^J mod phi(t) as 624 words (bit i of word w = coefficient
 * 32 w + i, phi = characteristic polynomial of c/mt19937/mt19937.c:18-55); _double squares it (2 J);
 * _apply_host returns the 624-word state window J words after state_in's. */
MOOCORE_API int hvb200_mtjump_poly(uint64_t J, uint32_t poly_out[624]);
MOOCORE_API int hvb200_mtjump_double(const uint32_t poly_in[624], uint32_t poly_out[624]);
MOOCORE_API int hvb200_mtjump_apply_host(const uint32_t state_in[624], const uint32_t poly[624], uint32_t state_out[624]);

/* Batched point sets: the estimates of `nsets` sets that share ref, maximise and ONE direction set.  Set i holds
 * the rows [cumsizes[i-1], cumsizes[i]) of `data` — the layout and the loop of c/main-hvapprox.c:149-174, which a
 * caller replaces by this one call.  volumes_out[i] = 0.0 when no point of set i strictly dominates ref, NaN on
 * failure (return value != 0).  Every volume equals the single-set entry point's up to the summation order
 * (<= 1e-14 relative). */
MOOCORE_API int hvb200_hv_approx_sets(int method, const double *data, const size_t *cumsizes, size_t nsets,
                                      dimension_t nobjs, const double *ref, const boolvec *maximise,
                                      uint_fast32_t nsamples, uint32_t seed, int device, double *volumes_out);

/* Hypervolume-CONTRIBUTION approximation on the direction set of hv_approx (the polar-coordinate estimator of the DZ2019
 * paper cited at c/hvapprox.h:20-24; the reference itself only has the exact c/hv_contrib.c, so this is an extension and
 * its parity is unpinned): contrib_out[i] ~ HV(P) - HV(P \ {p_i}) for each of the npoints rows (0 for rows that do not
 * strictly dominate ref), *hv_out (optional) the hypervolume estimate of the same direction set.  nobjs <= 32.
 * ignore_dominated != 0 follows hv_contributions(..., ignore_dominated = true) of c/hv_contrib.c:348: strictly dominated
 * points are removed first and get 0; with 0 they stay in the set and shield the points that dominate them. */
MOOCORE_API int hvb200_hv_contributions_approx(int method, const double *data, size_t npoints, dimension_t nobjs,
                                               const double *ref, const boolvec *maximise, uint_fast32_t nsamples,
                                               uint32_t seed, int device, int ignore_dominated, double *contrib_out, double *hv_out);

/* Single-process multi-GPU: evaluate one estimate sharded over `ndevices` GPUs
 * (devices 0..ndevices-1, contiguous index ranges, partial sums combined with one NCCL
 * all-gather when libnccl can be loaded, else by per-device copies). */
MOOCORE_API double hvb200_hv_approx_multi(int method, const double *data, size_t npoints, dimension_t nobjs,
                                          const double *ref, const boolvec *maximise,
                                          uint_fast32_t nsamples, uint32_t seed, int ndevices);

/* How the calling thread's last multi-GPU call (hvb200_hv_approx_multi, or a drop-in entry point under
 * HVB200_NGPUS=k) combined the per-GPU partial sums: 0 = single device, 1 = NCCL all-gather, 2 = host copies. */
MOOCORE_API int hvb200_multi_last_exchange(void);

/* FP64 pipe micro-benchmark used for the roofline denominator: returns measured DFMA
 * FLOP/s (2 flops per DFMA) of `device`, or a negative number on failure. */
MOOCORE_API double hvb200_measure_fp64_flops(int device, double *dmul_per_s, double *dsetp_per_s);

#ifdef __cplusplus
}
#endif
#endif
