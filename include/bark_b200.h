/*
 * bark_b200.h  --  C ABI of the B200-native BARK hot path (libbark_b200.so).
 *
 * The reference (TobyBoyne/bark) has no FFI: its hot path is in-process
 * Python/numba.  Each entry point below names the reference function(s) it
 * replaces (paths relative to the reference repo root).  INTEGRATION.md shows
 * the ctypes stubs a maintainer adds on the reference side.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in `_host`;
 *   - sizes are int64_t, indices follow the reference's C-order layouts;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *   - no allocation and no host synchronisation inside: the caller owns all
 *     buffers (workspace sizes come from the *_bytes functions);
 *   - return value: 0 = ok, otherwise a BARK_E_* code; bark_last_error()
 *     returns a static string for the calling thread;
 *   - asynchronous failures (tree container overflow, leaf-column capacity,
 *     non-SPD matrix) are reported through per-chain status words the caller
 *     reads back (BARK_ST_* bits).
 */
#ifndef BARK_B200_H
#define BARK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BARK_ABI_VERSION 1

/* return codes */
#define BARK_OK 0
#define BARK_E_INVALID 1   /* bad argument                          */
#define BARK_E_CUDA 2      /* CUDA runtime error (see last error)   */
#define BARK_E_UNSUPPORTED 3

/* per-chain status bits (device side) */
#define BARK_ST_TREE_OVERFLOW 1u /* grow found < 2 inactive slots: reference raises OverflowError,
                                    src/bark/fitting/tree_proposals.py:57-58 */
#define BARK_ST_HYPER_MODE 2u    /* use_softplus_transform=False and sample_scale=False: reference raises
                                    NotImplementedError, src/bark/fitting/noise_scale_proposals.py:78-81 */
#define BARK_ST_COL_OVERFLOW 4u  /* leaf-column capacity p_cap exhausted (re-run with a larger p_cap)   */
#define BARK_ST_NOT_SPD 8u       /* Cholesky met a non-positive pivot                                  */
#define BARK_ST_TIMEOUT 16u      /* a device-side pipeline wait gave up (internal error; results invalid) */

/* NODE_RECORD_DTYPE (src/bark/forest.py:8-19): packed 26-byte records
 * is_leaf u8@0, feature_idx u32@1, threshold f32@5, left u32@9, right u32@13,
 * parent u32@17, depth u32@21, active u8@25. */
#define BARK_NODE_BYTES 26

/* Device-side struct-of-arrays view of `n_nodes` node records (lossless). */
typedef struct bark_nodes_soa {
    uint8_t* is_leaf;
    uint8_t* active;
    uint32_t* feature;
    float* threshold;
    uint32_t* left;
    uint32_t* right;
    uint32_t* parent;
    uint32_t* depth;
} bark_nodes_soa;

/* MCMC parameters: BARKTrainParamsNumba (src/bark/fitting/bark_sampler.py:48-92) */
typedef struct bark_params {
    double alpha;
    double beta;
    double proposal_weights[3]; /* grow, prune, change (normalised) */
    double gamma_prior_shape;
    double gamma_prior_rate;
    int32_t use_softplus_transform;
    int32_t sample_scale;
} bark_params;

int bark_abi_version(void);
const char* bark_last_error(void);

/* ---- a1: forest container codec (src/bark/forest.py:8-19) --------------------------------- */
/* AoS bytes (n_nodes*26) -> SoA and back; bit-for-bit lossless, stale fields included. */
int bark_nodes_unpack(const uint8_t* aos, int64_t n_nodes, bark_nodes_soa soa, void* stream);
int bark_nodes_pack(bark_nodes_soa soa, int64_t n_nodes, uint8_t* aos, void* stream);

/* ---- a2: traversal (pass_through_forest, src/bark/forest.py:28-67) ------------------------- */
/* nodes: n_forests*m*node_limit records (SoA); X (n_points, d) f64 row-major, shared by all forests;
 * feat_types (d) i32 (0 Cat, 1 Int, 2 Cont); leaves out: (n_forests, n_points, m) u32, C order. */
int bark_traverse(bark_nodes_soa nodes, int64_t n_forests, int64_t m, int64_t node_limit, const double* X,
                  int64_t n_points, int64_t d, const int32_t* feat_types, uint32_t* leaves, void* stream);

/* ---- a4: Gram (forest_gram_matrix & batched, src/bark/forest.py:78-98) --------------------- */
/* counts[b,i,j] = #{t: leaves_a[b,i,t] == leaves_b[b,j,t]}  (exact int32), computed as an int8 one-hot GEMM on
 * tcgen05 (UMMA kind::i8, s32 TMEM accumulators, operands streamed by the bulk-copy engine); the FP64 kernel matrix
 * can be produced in the same pass.
 * slots: upper bound on (leaf slot id + 1), i.e. ids must be < slots (<= 256); *status |= 1 otherwise.
 * counts (batch,na,nb) i32 and K (batch,na,nb) f64 are both optional (not both NULL).  K uses scale/noise/jitter/
 * add_diag exactly as bark_gram_to_kernel.  workspace: bark_gram_workspace_bytes(...) bytes. */
size_t bark_gram_workspace_bytes(int64_t batch, int64_t na, int64_t nb, int64_t m, int32_t slots);
int bark_gram_umma(const uint32_t* leaves_a, const uint32_t* leaves_b, int64_t batch, int64_t na, int64_t nb, int64_t m,
                   int32_t slots, int32_t* counts, double* K, const double* scale, const double* noise, double jitter,
                   int add_diag, uint32_t* status, void* workspace, void* stream);
/* K[b] = scale[b] * ((1.0/m) * counts[b]) + (jitter + noise[b]) * I   (same multiply order as
 * src/bark/fitting/bark_sampler.py:153-156; diag added only when add_diag != 0 and na == nb). */
int bark_gram_to_kernel(const int32_t* counts, int64_t batch, int64_t na, int64_t nb, int64_t m, const double* scale,
                        const double* noise, double jitter, int add_diag, double* K, void* stream);

/* ---- a5: batched FP64 Cholesky log-marginal-likelihood (mll, src/bark/fitting/quick_inverse.py:36-38;
 *          call sites src/bark/fitting/bark_sampler.py:160-162,269-272) ---------------------- */
/* K (batch, n, n) f64 row-major (lower triangle read), OVERWRITTEN (used as the factorisation workspace of a
 * square-root-free block Cholesky, LDL^T with 64x64 pivot blocks); y (n) f64 shared by the batch.
 * out_mll/out_logdet/out_quad (batch): 0.5*(-quad - logdet), log|K|, y^T K^-1 y.  status (batch) u32, OR-ed.
 * workspace: bark_mll_workspace_bytes(batch, n) bytes of device scratch. */
size_t bark_mll_workspace_bytes(int64_t batch, int64_t n);
int bark_mll_batched(double* K, int64_t batch, int64_t n, const double* y, double* out_mll, double* out_logdet,
                     double* out_quad, uint32_t* status, void* workspace, void* stream);

/* ---- a8-a13: device-resident MCMC (run_bark_sampler, src/bark/fitting/bark_sampler.py:95-284) */
typedef struct bark_mcmc_dims {
    int64_t chains;     /* chains resident on this device                      */
    int64_t n;          /* training points                                     */
    int64_t d;          /* features                                            */
    int64_t m;          /* trees                                               */
    int64_t node_limit; /* slots per tree (100 in the reference)               */
    int64_t p_cap;      /* leaf-column capacity per chain (multiple of 64)     */
} bark_mcmc_dims;

/* bytes of the opaque per-device MCMC workspace for these dims */
size_t bark_mcmc_workspace_bytes(const bark_mcmc_dims* dims);

/* Largest leaf-column capacity (multiple of 64) the sweep kernel's shared-memory working set allows for these
 * n / d / node_limit (p_cap and chains of `dims` are ignored); 0 if none.  The host side clamps its default and its
 * overflow retry (the analogue of the reference's fixed node container, src/bark/fitting/tree_proposals.py:45-58)
 * to this value. */
int64_t bark_mcmc_max_p_cap(const bark_mcmc_dims* dims);

/* Build chain state from (forest, noise, scale): leaf bitsets, leaf co-occurrence counts, B^-1, log-MLL.
 * X (n,d) row-major f64, y (n) f64, bounds (d,2) f64, feat_types (d) i32, noise/scale (chains) f64. */
int bark_mcmc_init(const bark_mcmc_dims* dims, void* workspace, bark_nodes_soa forest, const double* X,
                   const double* y, const double* bounds, const int32_t* feat_types, const double* noise,
                   const double* scale, void* stream);

/* Same with flags.  BARK_INIT_SKIP_NULL: root-only trees get no leaf column and do not count in m, i.e. the state
 * is the one of the kernel of batched_forest_gram_matrix_no_null (src/bark/forest.py:101-111) that the acquisition
 * model uses (src/bark/optimizer/opt_model.py:54-59).  Such a state is for bark_kinv_export only (no sweeps). */
#define BARK_INIT_SKIP_NULL 1
int bark_mcmc_init_ex(const bark_mcmc_dims* dims, void* workspace, bark_nodes_soa forest, const double* X,
                      const double* y, const double* bounds, const int32_t* feat_types, const double* noise,
                      const double* scale, int32_t flags, void* stream);

/* Run `n_sweeps` sweeps (m tree MH steps + 1 noise/scale MH step each, bark_sampler.py:216-284) on every
 * chain, device-resident.  RNG: Philox4x32-10 keyed by (seed, chain_offset + chain) and counted by
 * (sweep_offset + sweep, tree, slot) unless `tape` != NULL, in which case every random number is read from
 * tape[chain][sweep][m*5+3] (layout of oracle/bark_oracle.py: u_type,u_node,u_feat,u_rule,u_accept per tree,
 * then z_noise,z_scale,u_accept).  trace (optional): [chain][sweep][m+1][3] = log_q_prior, proposed mll,
 * accepted.  The forest SoA passed to bark_mcmc_init is updated in place (as the reference mutates its input,
 * bark_sampler.py:148,264). */
int bark_mcmc_sweeps(const bark_mcmc_dims* dims, void* workspace, bark_nodes_soa forest, const bark_params* params,
                     int64_t n_sweeps, uint64_t seed, int64_t chain_offset, int64_t sweep_offset, const double* tape,
                     double* trace, void* stream);

/* Same, with the refresh policy explicit: the running leaf-space state (B^-1, w, residual, log-det) of a chain is
 * rebuilt exactly when its noise/scale move is accepted -- where the reference rebuilds K^-1,
 * bark_sampler.py:276-282 -- and additionally every `refresh_every` sweeps (staggered over the chains; more often
 * for an ill-conditioned chain).  0: only on accept (the reference's behaviour); -1: the default policy
 * (8, or 0 when a tape is replayed). */
int bark_mcmc_sweeps_ex(const bark_mcmc_dims* dims, void* workspace, bark_nodes_soa forest, const bark_params* params,
                        int64_t n_sweeps, uint64_t seed, int64_t chain_offset, int64_t sweep_offset, const double* tape,
                        double* trace, int32_t refresh_every, void* stream);

/* Measurement variant (bench only; SYNCHRONISES the stream): same work without tape/trace, with CUDA events
 * recorded on `stream` around every kernel; returns the summed device time of the tree-sweep kernel and of the
 * hyper-step kernel in milliseconds (host pointers). */
int bark_mcmc_sweeps_timed(const bark_mcmc_dims* dims, void* workspace, bark_nodes_soa forest, const bark_params* params,
                           int64_t n_sweeps, uint64_t seed, int64_t chain_offset, int64_t sweep_offset,
                           float* ms_trees_host, float* ms_hyper_host, void* stream);
/* ... with the hyper step split: ms3_host[0..2] = tree sweep, noise/scale evaluation, exact refresh. */
int bark_mcmc_sweeps_timed3(const bark_mcmc_dims* dims, void* workspace, bark_nodes_soa forest, const bark_params* params,
                            int64_t n_sweeps, uint64_t seed, int64_t chain_offset, int64_t sweep_offset, float* ms3_host,
                            void* stream);

/* Read-out of per-chain scalars: each (chains) or NULL.  counters (chains, 16) u64:
 * [0 tree proposals issued, 1 valid, 2 accepted, 3 hyper issued, 4 hyper accepted, 5-7 grow/prune/change accepted,
 *  8-10 grow/prune/change valid, 11 sum over proposal blocks of extent^2 (one B^-1 V product pass each), 12 same over
 *  the blocks that accepted something (one update pass each), 13 leaf-bitset columns scanned (per block),
 *  14 sum over blocks of extent^2 x accepted proposals, 15 exact refreshes of the running state]. */
int bark_mcmc_read(const bark_mcmc_dims* dims, const void* workspace, double* noise, double* scale, double* mll,
                   uint32_t* status, uint64_t* counters, int32_t* p_used, void* stream);

/* Debug/verification export of one chain's leaf-space state (tests only): A (p_cap,p_cap) i32,
 * Binv (p_cap,p_cap) f64, colmap (m,node_limit) i32 (-1 = none), bits (p_cap, ceil(n/32)) u32. */
int bark_mcmc_export(const bark_mcmc_dims* dims, const void* workspace, int64_t chain, int32_t* A, double* Binv,
                     int32_t* colmap, uint32_t* bits, void* stream);

/* ---- inputs of the acquisition model (src/bark/optimizer/opt_model.py:54-59,83,101)
 * For every posterior sample of a workspace initialised by bark_mcmc_init[_ex] over the sample forests:
 *   kinv   (samples, n, n) f64 = K^-1,  K = scale * K0 + (1e-6 + noise) I   (np.linalg.inv at opt_model.py:59),
 *   kinv_y (samples, n)    f64 = K^-1 y   (the reference's lin_term is scale * K^-1 y, :101; quadr_term -scale^2 K^-1, :83)
 * from the leaf-space state by Woodbury: K^-1 = (I - Z B^-1 Z^T) / (noise + 1e-6) -- B is P x P with cond(B) << cond(K).
 * leaves (samples, n, m) u32: bark_traverse of the sample forests on the training inputs.  Either output may be NULL.
 * scratch: bark_kinv_scratch_bytes(dims) bytes. */
size_t bark_kinv_scratch_bytes(const bark_mcmc_dims* dims);
int bark_kinv_export(const bark_mcmc_dims* dims, const void* workspace, const uint32_t* leaves, double* kinv,
                     double* kinv_y, void* scratch, void* stream);

/* ---- a14-a15: posterior predictive (forest_predict, src/bark/tree_kernels/tree_gps.py:80-113;
 *               mixture_of_gaussians_as_normal :116-131; _predict src/bofire_mixed/surrogates/bark.py:71-94) */
/* Uses a workspace initialised by bark_mcmc_init over `dims->chains` = number of posterior samples
 * (forest = the sample forests, noise/scale = the sample hyper-parameters, X/y = training data).
 * candidates (n_c, d) f64 row-major.  scratch: bark_predict_scratch_bytes(dims, n_c) bytes.
 * mode 0: mu,var (samples, n_c) per sample (forest_predict, diag=True).
 * mode 1: mu,var (n_c): mixture moments of the per-sample Gaussians after the surrogate's un-standardisation:
 *         mu_j*y_std+y_mean, var_j*y_std^2 + add_noise*noise_j  (bark.py:83-91). */
size_t bark_predict_scratch_bytes(const bark_mcmc_dims* dims, int64_t n_c);
int bark_predict(const bark_mcmc_dims* dims, const void* workspace, bark_nodes_soa forest, const double* candidates,
                 int64_t n_c, int mode, double y_mean, double y_std, int add_noise, double* mu, double* var,
                 void* scratch, void* stream);

/* diag = False of forest_predict (src/bark/tree_kernels/tree_gps.py:107-112): per-sample mean mu (samples, n_c) and the
 * FULL matrix cov (samples, n_c, n_c) = scale - K_xX K^-1 K_Xx as the reference forms it (n_c <= 16384). */
size_t bark_predict_cov_scratch_bytes(const bark_mcmc_dims* dims, int64_t n_c);
int bark_predict_cov(const bark_mcmc_dims* dims, const void* workspace, bark_nodes_soa forest, const double* candidates,
                     int64_t n_c, double* mu, double* cov, void* scratch, void* stream);

/* Tensor-core predict (leaf-column extent p_max <= 768): per posterior sample, B^-1 is sliced once into 7 int8 digit
 * planes of a 54-bit fixed-point representation (bark_predict_prepare, into `prep`), then for every tile of 128
 * candidates z^T B^-1 z is an exact int8 one-hot GEMM on tcgen05 with masked int32 row sums (bark_predict_umma:
 * per-sample mu, var (samples, n_c)); bark_predict_mixture folds the samples as bark_predict mode 1 does.
 * slots: upper bound on (active node slot + 1) over the sample forests. */
size_t bark_predict_prep_bytes(const bark_mcmc_dims* dims, int32_t slots, int32_t p_max);
int bark_predict_prepare(const bark_mcmc_dims* dims, const void* workspace, bark_nodes_soa forest, int32_t slots,
                         int32_t p_max, void* prep, void* stream);
int bark_predict_umma(const bark_mcmc_dims* dims, const void* workspace, const void* prep, int32_t slots, int32_t p_max,
                      const double* candidates, int64_t n_c, double* mu, double* var, void* stream);
/* The same with the mixture over the samples folded inside the kernel (one persistent CTA per 128 candidates loops over
 * the samples): mu, var (n_c) as bark_predict mode 1 gives them, no (samples, n_c) intermediate. */
int bark_predict_umma_mixture(const bark_mcmc_dims* dims, const void* workspace, const void* prep, int32_t slots,
                              int32_t p_max, const double* candidates, int64_t n_c, double y_mean, double y_std,
                              int add_noise, double* mu, double* var, void* stream);
int bark_predict_mixture(const bark_mcmc_dims* dims, const void* workspace, const double* mu_s, const double* var_s,
                         int64_t n_c, double y_mean, double y_std, int add_noise, double* mu, double* var, void* stream);

/* ---- 8f-2: forests drawn from the BARK tree prior on the device (_sample_single_forest,
 * src/bark/fitting/bark_prior_sampler.py:15-62): every tree of `forest` (n_samples * m trees of node_limit slots, SoA) is
 * reset to a root leaf and grown by the depth prior alpha (1 + depth)^-beta with split rules drawn like a grow proposal's.
 * Philox streams keyed by (seed, sample, tree); *status |= BARK_ST_TREE_OVERFLOW if a tree runs out of slots. */
int bark_prior_sample(bark_nodes_soa forest, int64_t n_samples, int64_t m, int64_t node_limit, const double* bounds,
                      const int32_t* feat_types, int64_t d, double alpha, double beta, uint64_t seed, uint32_t* status,
                      void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BARK_B200_H */
