/*
 * pbx.h -- C ABI of the B200-native Pibronic PIMC block estimator ("pbx").
 *
 * The reference (ngraymon/Pibronic) has no FFI: its hot path is the Python/numpy
 * functions of pibronic/pimc/pimc.py.  Each entry point below names the
 * reference interface it replaces (file:line relative to the reference tree):
 *
 *   pbx_plan_create      BoxData[PM].preprocess() + ModelVibronic[PM].precompute()
 *                        + ModelSampling.precompute()          pimc.py:647-692, 220-269, 409-424, 59-89
 *   pbx_sample_eval      the block loop of block_compute[_pm]   pimc.py:1358-1381, 1413-1449
 *                        (draw_sample 326-334, transform_sampled_coordinates 613-631,
 *                         build_o_matrix 1087-1129, build_scaling_factors 1076-1084,
 *                         build_denominator 1132-1136, diagonalize_coupling_matrix 1139-1174,
 *                         build_numerator 1177-1213)
 *   pbx_eval_coords      the same estimator on caller-supplied bead coordinates
 *                        (the reference's golden test feeds data.qTensor directly,
 *                         tests/pimc/test_pimc_explicit_example.py:183-205;
 *                         block_compute_rhoR_from_input_samples pimc.py:1273-1302)
 *   pbx_eval_stages      build_o_matrix / diagonalize_coupling_matrix / build_numerator
 *                        intermediates (omatrix, coupling_matrix, M_matrix)   pimc.py:1087-1187
 *   pbx_sample_coords    draw_sample + transform_sampled_coordinates only      pimc.py:326-334, 613-631
 *   pbx_chain_trace      build_numerator's bead chain on caller supplied M and O          pimc.py:1194-1209
 *   pbx_block_sums       per-block sums consumed by pibronic/stats/stats.py:38-55, 84-123
 *   pbx_stats_*          basic_jackknife_analysis                       stats/stats.py:271-299, jackknife.py:60-105
 *
 * Conventions
 *   - plain C: pointers + sizes, no C++/torch types.  All arrays are IEEE float64, C order.
 *   - the caller allocates every buffer.  "_dev" entry points take DEVICE pointers and are
 *     stream-ordered on `stream` (a cudaStream_t passed as void*, NULL = legacy default stream);
 *     they never synchronise.  "_host" entry points take HOST pointers, do the H2D/D2H copies
 *     themselves and return after the results are in the host buffers.
 *   - every function returns PBX_OK (0) or a negative pbx_status; pbx_last_error() gives the text
 *     for the calling thread.  Nothing throws across the ABI.
 *   - a plan is bound to one device; use one plan per (device, host thread).
 *   - there is NO CPU fallback: without a CUDA device pbx_plan_create fails with PBX_ERR_CUDA.
 */
#ifndef PBX_H
#define PBX_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PBX_ABI_VERSION 1
#define PBX_NSUMS 8   /* per-block sums, see pbx_block_sums_dev */

typedef enum pbx_status {
    PBX_OK = 0,
    PBX_ERR_ARG = -1,       /* bad argument (null pointer, size <= 0, P < 3, ...)            */
    PBX_ERR_MODEL = -2,     /* model not symmetric in the surfaces / omega <= 0 / N mismatch */
    PBX_ERR_CUDA = -3,      /* CUDA runtime error (text in pbx_last_error)                  */
    PBX_ERR_UNSUPPORTED = -4 /* size outside what the kernels handle                        */
} pbx_status;

/* flags for pbx_plan_create */
enum {
    PBX_FLAG_PM             = 1u << 0, /* also evaluate g(beta +/- delta_beta): block_compute_pm          */
    PBX_QUIRK_RHO_TRUNC     = 1u << 1, /* reference quirk pimc.py:1110-1111: rho(R) uses only the first   */
                                       /* min(A, A_rho) sampling surfaces (bit-parity runs with A_rho > A) */
    PBX_FLAG_M_TAU_PM       = 1u << 2, /* g+/- use exp(-tau+/- V), the consistent beta +/- delta_beta      */
                                       /* estimator; off = the reference's exp(-tau V) for all three       */
                                       /* (pimc.py:1183).  Needs PBX_FLAG_PM, a register-resident shape and */
                                       /* a library built with PBX_WITH_MTAU=1 (not the default build).     */
    PBX_FLAG_EIG_JACOBI     = 1u << 3, /* M = U exp(-tau lambda) U^T by a Jacobi eigensolve (reference's  */
                                       /* formulation); default is a scaling-and-squaring exp(-tau V)      */
    PBX_FLAG_FORCE_GENERIC  = 1u << 4, /* never use the register-resident small-A kernels                 */
    PBX_FLAG_NO_SCALING     = 1u << 5, /* skip the per-bead S scaling (golden test of the reference)       */
    PBX_FLAG_NO_WARPSPEC    = 1u << 6, /* fused sampler+estimator on the one-role kernel instead of the        */
                                       /* producer/consumer (warp-specialised) one; same results bit for bit  */
    PBX_FLAG_NO_FUSED_DMMA  = 1u << 7, /* 1 <= A <= 16 without a register-resident kernel: the blocked kernels   */
                                       /* through HBM scratch instead of the fused one-launch tensor-core kernel  */
    PBX_FLAG_PREFER_DMMA    = 1u << 8, /* use the fused tensor-core kernel even where a register-resident one exists */
    PBX_QUIRK_RHO_DOUBLE_SHIFT = 1u << 9 /* reference quirk pimc.py:1293-1298 (block_compute_rhoR_from_input_samples): rho(R) is   */
                                       /* evaluated at R - d_vib[a] - d_rho[a] instead of R - d_rho[a]; needs A_rho == A      */
};

/* which kernel family a plan runs on (pbx_plan_kernel_path) */
enum {
    PBX_PATH_GENERIC    = 0,  /* any A: warp per (sample, bead) + warp per sample, through HBM scratch          */
    PBX_PATH_REGISTER   = 1,  /* listed small shapes: one sample per thread, everything in registers            */
    PBX_PATH_BLOCKED    = 2,  /* A <= 16: four blocked kernels through HBM scratch (round-1 large-A path)       */
    PBX_PATH_FUSED_DMMA = 3   /* 1 <= A <= 16: one launch, one warp per sample, FP64 tensor cores, no scratch   */
};

/* coupled (vibronic) model: pibronic `coupled_model.json` arrays as loaded by ModelClass.load_model */
typedef struct pbx_model {
    int32_t A;                /* number of surfaces                         */
    int32_t N;                /* number of modes                            */
    const double *energy;     /* [A][A]                                     */
    const double *omega;      /* [N]                                        */
    const double *linear;     /* [N][A][A]   (may be NULL = zeros)          */
    const double *quadratic;  /* [N][N][A][A] (may be NULL = zeros)         */
} pbx_model;

/* sampling model rho: `sampling_model.json` arrays as loaded by ModelSampling.load_model */
typedef struct pbx_rho {
    int32_t A;                /* number of sampling surfaces A_rho          */
    int32_t N;                /* number of modes (must equal the model's)   */
    const double *energy;     /* [A_rho]                                    */
    const double *omega;      /* [N]                                        */
    const double *linear;     /* [N][A_rho]  (may be NULL = zeros)          */
} pbx_rho;

typedef struct pbx_plan pbx_plan;   /* opaque */

int pbx_abi_version(void);
/* optional parts compiled into this library: PBX_FEATURE_MTAU = the PBX_FLAG_M_TAU_PM kernels (build with PBX_WITH_MTAU=1) */
enum { PBX_FEATURE_MTAU = 1 };
int pbx_library_features(void);
const char *pbx_last_error(void);
/* number of CUDA devices visible (0 if none / driver missing) */
int pbx_device_count(void);

/* Builds all temperature dependent tables on the host (coth/csch/prefactors for tau, tau+, tau-;
 * shifts d, Delta; mixture weights; packed symmetric E_off/L_off/Q; sampler recurrence tables)
 * and uploads them to `device`.  beta in 1/eV.  device == -1 builds a host-only plan whose tables can
 * be read back with pbx_plan_table but which cannot launch anything (every other call fails). */
int pbx_plan_create(const pbx_model *vib, const pbx_rho *rho, int32_t beads, double beta,
                    double delta_beta, uint32_t flags, int32_t device, pbx_plan **out);
int pbx_plan_destroy(pbx_plan *plan);

/* Register-resident kernels exist for the shapes compiled into the library (csrc/shapes.def).  Further shapes can be
 * compiled at run time into their own shared library (pibronic_b200/jit.py runs nvcc on csrc/pbx_fast_inst.cu) and
 * registered here; plans created afterwards use them.  pbx_has_register_kernel: 1 if (A, N, A_rho) is known. */
int pbx_has_register_kernel(int32_t A, int32_t N, int32_t A_rho);
int pbx_register_shape_library(const char *path);

/* Introspection of the precomputed host tables (tests compare them with the reference's).
 * name is one of: "d_vib"[A][N] "d_rho"[Ar][N] "delta_vib"[A] "delta_rho"[Ar] "weights"[Ar]
 * "coth"[4][N] "csch"[4][N] (rows: vib tau, tau+, tau-, rho tau) "logpref"[3][A] "logpref_rho"[Ar]
 * "e_off"[AA] "l_off"[N][AA] "q_pack"[NN][AA] "samp"[P][N][3] "tau"[3].
 * Copies min(count, available) doubles, returns the number available (or a negative status). */
int64_t pbx_plan_table(const pbx_plan *plan, const char *name, double *out, int64_t count);
/* 1 if the plan runs on a register-resident small-A kernel, 0 if on the generic kernels */
int pbx_plan_is_fast(const pbx_plan *plan);
/* PBX_PATH_* of the kernels the fused / coordinate entry points of this plan launch */
int pbx_plan_kernel_path(const pbx_plan *plan);
/* number of kernel launches issued through this plan so far */
int64_t pbx_plan_launch_count(const pbx_plan *plan);
/* bytes of kernel parameters (the model tables of the register-resident kernels travel this way) sent host->device
 * with every estimator launch */
int64_t pbx_plan_launch_param_bytes(const pbx_plan *plan);

/* Fused sampler + estimator for global sample indices [first_sample, first_sample + n_samples).
 * Philox4x32-10 keyed by `seed`, counter = global sample index: results do not depend on how the
 * index range is split over calls, streams or GPUs.
 *   out4_dev : [4][n_samples] rows rho, g, g+, g- (rows 2,3 untouched unless PBX_FLAG_PM).
 *   out4_host: same rows with a row stride of ld_host >= n_samples doubles.
 *   sums_host: if not NULL, [ceil(n_samples/block_size)][PBX_NSUMS] per-block sums (see pbx_block_sums_dev). */
int pbx_sample_eval_dev(pbx_plan *plan, uint64_t seed, int64_t first_sample, int64_t n_samples,
                        double *out4_dev, void *stream);
int pbx_sample_eval_host(pbx_plan *plan, uint64_t seed, int64_t first_sample, int64_t n_samples,
                         double *out4_host, int64_t ld_host, int64_t block_size, double *sums_host);

/* Estimator on caller supplied bead coordinates R[n_samples][N][P] (the reference's qTensor[:,0]).
 * R_dev is read in place (no staging copy, no scratch for the register-resident and tensor-core kernels). */
int pbx_eval_coords_dev(pbx_plan *plan, const double *R_dev, int64_t n_samples, double *out4_dev,
                        void *stream);
int pbx_eval_coords_host(pbx_plan *plan, const double *R_host, int64_t n_samples, double *out4_host,
                         int64_t ld_host);

/* Sampler only: R_dev[n_samples][N][P], src_dev[n_samples] (mixture component; may be NULL). */
int pbx_sample_coords_dev(pbx_plan *plan, uint64_t seed, int64_t first_sample, int64_t n_samples,
                          double *R_dev, int32_t *src_dev, void *stream);

/* Intermediates of the estimator for given coordinates (any output may be NULL):
 *   o_rho [n][P][Ar], o_vib [3][n][P][A] (tau, tau+, tau-; all divided by S unless NO_SCALING),
 *   scale [n][P], v_mat [n][P][A][A], m_mat [n][P][A][A].  Device pointers. */
int pbx_eval_stages_dev(pbx_plan *plan, const double *R_dev, int64_t n_samples, double *o_rho,
                        double *o_vib, double *scale, double *v_mat, double *m_mat, void *stream);

/* Bead chain only (build_numerator, pimc.py:1194-1209): g[x] = tr prod_p ( m_mat[x][p] . diag(o_diag[x][p]) )
 * for caller supplied m_mat [n][P][A][A] and o_diag [n][P][A] (device pointers). */
int pbx_chain_trace_dev(pbx_plan *plan, const double *m_mat, const double *o_diag, int64_t n_samples,
                        double *g_out, void *stream);

/* Per-block sums of the estimator terms, blocks of `block_size` consecutive samples:
 *   sums_dev[n_blocks][PBX_NSUMS] = sum r, sum r+, sum r-, sum r^2, sum d1, sum d2, sum d1^2, sum d2^2
 *   with r = g/rho, r+- = g+-/rho, d1 = (g+ - g-)/(2 db rho), d2 = (g+ - 2g + g-)/(db^2 rho).
 * Deterministic (fixed summation order). */
int pbx_block_sums_dev(pbx_plan *plan, const double *out4_dev, int64_t n_samples, int64_t block_size,
                       double *sums_dev, void *stream);

/* On-device Z, E, Cv and their leave-one-out jackknife (what stats.basic_jackknife_analysis computes from the
 * four arrays, pibronic/stats/stats.py:38-55, 84-123, 271-299 and jackknife.py:60-105), WITHOUT the harmonic
 * contribution of the sampling model (the caller adds E_sampling / Cv_sampling).  Needs a PBX_FLAG_PM plan.
 *   stats_host[PBX_NSTATS] = Z, Z error, E, E error (0), Cv, Cv error (0), jk_E, jk_E error, jk_Cv, jk_Cv error
 * pbx_stats_dev : out4 is a DEVICE [4][n] array (row stride n); synchronises `stream`.
 * pbx_stats_host: out4 is a HOST array with row stride ld_host (copied to the device first).
 * pbx_stats_last: statistics of the most recent *_host call, from its device-resident copy (no transfer). */
#define PBX_NSTATS 10
int pbx_stats_dev(pbx_plan *plan, const double *out4_dev, int64_t n_samples, double *stats_host, void *stream);
int pbx_stats_host(pbx_plan *plan, const double *out4_host, int64_t ld_host, int64_t n_samples, double *stats_host);
int pbx_stats_last(pbx_plan *plan, double *stats_host);
/* Plan-free forms: the statistics depend on (beta, delta_beta) only.  out4 is [4][n] (row stride n / ld_host) on `device`
 * (pbx_stats_arrays_dev; synchronises `stream`) or on the host (pbx_stats_arrays_host: uploaded first). */
int pbx_stats_arrays_dev(const double *out4_dev, int64_t n_samples, double beta, double delta_beta, int32_t device,
                         double *stats_host, void *stream);
int pbx_stats_arrays_host(const double *out4_host, int64_t ld_host, int64_t n_samples, double beta, double delta_beta,
                          int32_t device, double *stats_host);

/* Self-test hook for the branch-free device math used by the kernels (pbx_device.cuh):
 * kind 0 ln(x) for x in (0,1], 1 sqrt(x), 2 exp(x) for x <= 0, 3 sin(2 pi x), 4 cos(2 pi x) for x in [0,1). */
int pbx_math_probe_dev(int32_t kind, const double *in_dev, double *out_dev, int64_t n, void *stream);

/* Measured FP64 FMA throughput of the device in TFLOP/s (dependent-chain DFMA micro-kernel);
 * the roofline denominator used by bench.py. */
int pbx_fp64_peak_tflops(int32_t device, double *tflops_out);
/* kind 0: the vector DFMA probe above; 1: the same through the FP64 tensor path (mma.sync m8n8k4, which shares the
 * FP64 units with the vector pipe on B200 but reaches a higher fraction of them); 2: the larger of the two. */
int pbx_fp64_peak_tflops_kind(int32_t device, int32_t kind, double *tflops_out);

#ifdef __cplusplus
}
#endif
#endif /* PBX_H */
