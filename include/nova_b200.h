/*
 * nova_b200.h -- C ABI of libnova_b200.so: the B200 (sm_100a) implementation of
 * NOVA_pointcloud's per-token diffusion-head sampling path and Chamfer scorer.
 *
 * The reference (zailaiyiwan123/NOVA_pointcloud) is pure Python and has no FFI; the
 * entry points below are what a binding for this path replaces, one by one:
 *
 *   nova_head_create/load/destroy  DiffusionMLP.__init__ + load_state_dict
 *                                  (diffnext/models/diffusion_mlp.py:81-87; key set SURVEY.md A.2)
 *   nova_head_forward              DiffusionMLP.forward(x, timestep, z, pred_ids)
 *                                  (diffnext/models/diffusion_mlp.py:89-99)
 *   nova_head_sample               Transformer3DModel.denoise: the S-step loop of
 *                                  head -> GuidanceScaler.scale -> FlowMatchEuler step
 *                                  (diffnext/models/transformers/transformer_3d.py:102-113,
 *                                   diffnext/models/guidance_scaler.py:46-87,
 *                                   diffnext/schedulers/scheduling_cfm.py:125-140)
 *   nova_euler_step                FlowMatchEulerDiscreteScheduler.step
 *                                  (diffnext/schedulers/scheduling_cfm.py:134-137)
 *   nova_chamfer_nn                the nearest-neighbour primitive under chamfer_distance
 *                                  (demo.py:38-55), distChamfer (train_newloss.py:316-349) and
 *                                  compute_chamfer_distance (test_optimize.py:354-383)
 *
 * Conventions
 *   - every function returns 0 on success, a negative NOVA_ERR_* otherwise;
 *     nova_last_error() returns a thread-local message for the last failure.
 *   - no exceptions, no CPU fallback, no hidden device allocation except the packed
 *     weights owned by a head handle; scratch space is a caller-provided workspace.
 *   - all data pointers are DEVICE pointers unless a parameter says "host";
 *     tensors are row-major and contiguous; `stream` is a cudaStream_t passed as void*.
 *   - no call synchronises the device; errors from asynchronous work surface on the
 *     caller's next synchronisation, as with any CUDA launch.
 *   - a head handle is immutable after nova_head_load: concurrent calls on different
 *     streams with different workspaces are safe.
 *   - token layout: one row per token, T = C*p*p values ordered (p_h, p_w, C) with the
 *     channel fastest, exactly PatchEmbed.patchify (diffnext/models/embeddings.py:152-154).
 */
#ifndef NOVA_B200_H_
#define NOVA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NOVA_B200_ABI_VERSION 2

#if defined(__GNUC__)
#define NOVA_API __attribute__((visibility("default")))
#else
#define NOVA_API
#endif

enum nova_status {
  NOVA_OK = 0,
  NOVA_ERR_INVALID = -1,     /* bad argument / unsupported shape */
  NOVA_ERR_CUDA = -2,        /* a CUDA runtime/driver call failed */
  NOVA_ERR_WORKSPACE = -3,   /* workspace too small */
  NOVA_ERR_NOT_LOADED = -4,  /* weights missing */
  NOVA_ERR_DEVICE = -5       /* not an sm_100 device */
};

enum nova_dtype {
  NOVA_F32 = 0,  /* SIMT fp32 arithmetic: the <=1e-5 parity mode */
  NOVA_BF16 = 1  /* bf16 operands on tcgen05 tensor cores, fp32 accumulate: the fast path */
};

typedef struct nova_head nova_head_t; /* opaque */

typedef struct nova_head_config {
  int32_t depth;      /* residual blocks (6 for mlp_d6w*, 3 for mlp_d3w1280) */
  int32_t width;      /* D, head width; multiple of 256, <= 2048 */
  int32_t cond_width; /* Dc, width of the condition z; multiple of 64 */
  int32_t token_dim;  /* T = C*p*p (3 for xyz points, 16 for the registry default); <= 64 */
  int32_t dtype;      /* nova_dtype: arithmetic type of the handle */
} nova_head_config;

/* guidance for nova_head_sample (GuidanceScaler, diffnext/models/guidance_scaler.py:74-87) */
typedef struct nova_guidance {
  float scale;  /* <= 1: off. > 1: z holds [cond; uncond] (2x rows), v = vu + (vc - vu)*scale */
  float trunc;  /* > 0: guidance is switched off once timestep < trunc (guidance_scaler.py:59-65) */
  float renorm; /* < 1: v *= clamp(|vc| / |v|, renorm, 1), norms per cloud (guidance_scaler.py:67-72) */
  /* three-pass forms (:78-85): z holds [cond; uncond; third] (3x rows).  At most one is used, image first:
   *   image_scale > 0:          v = renorm(vu + (vc - v3)*scale) + (v3 - vu)*image_scale      (third = image-only pass)
   *   spatiotemporal_scale > 0: v = renorm(vu + (vc - vu)*scale) + (vc - v3)*spatiotemporal_scale  (third = perturbed) */
  float image_scale;
  float spatiotemporal_scale;
} nova_guidance;

NOVA_API const char* nova_last_error(void);
NOVA_API int nova_abi_version(void);
/* 0 if the current device can run this library (compute capability 10.x), else NOVA_ERR_DEVICE. */
NOVA_API int nova_device_check(void);

NOVA_API int nova_head_create(const nova_head_config* cfg, nova_head_t** out);
NOVA_API int nova_head_destroy(nova_head_t* h);
NOVA_API int nova_head_get_config(const nova_head_t* h, nova_head_config* out);

/*
 * Load weights given in the reference's state_dict naming and shapes:
 * nn.Linear weights are (out,in); patch_embed.proj.weight is (D,C,p,p) and is permuted
 * to token order here; the 7 AdaLN projections are concatenated into one (20D,D) operand.
 * `src_dtype` (nova_dtype) is the element type of every array; `ptrs` are DEVICE pointers.
 * All 14 + 8*depth keys must be present; patch size p is taken as sqrt(T / C) from shapes.
 */
NOVA_API int nova_head_load(nova_head_t* h, int32_t n, const char* const* names, const void* const* ptrs,
                   const int64_t* numels, int32_t src_dtype, int32_t channels, void* stream);

/* Bytes of scratch needed for a call over `rows` head rows (M' = guidance passes x tokens). */
NOVA_API size_t nova_head_workspace_bytes(const nova_head_t* h, int64_t rows, int32_t num_steps);

/*
 * One velocity prediction.
 *   x_tok   [Bx, N, T] fp32   token-layout latent (Bx = B, or B/2 when rows pair up for guidance)
 *   t       [B] or [B, n] fp32 timesteps (t_per_token != 0 selects the per-token form)
 *   z       [B, N, Dc]        condition, element type = handle dtype
 *   pred_ids[B, n] int64 or NULL; NULL means every token (n == N)
 *   v_out   [B, n, T] fp32    velocity of the selected tokens (compact; the caller scatters)
 */
NOVA_API int nova_head_forward(const nova_head_t* h, const float* x_tok, const float* t, int32_t t_per_token,
                      const void* z, const int64_t* pred_ids, int64_t B, int64_t Bx, int64_t N, int64_t n,
                      float* v_out, void* workspace, size_t workspace_bytes, void* stream);

/*
 * DiffusionMLP.forward with a PRE-EMBEDDED input: a 3-D x passes through PatchEmbed.forward unchanged
 * (diffnext/models/embeddings.py:160-166), so the blocks start from the caller's rows.
 *   x_emb [B, N, D] handle dtype; t, t_per_token, z, v_out, workspace as nova_head_forward (same workspace size).
 * No pred_ids: the reference's scatter target patchify(x) does not exist for an embedded input.
 */
NOVA_API int nova_head_forward_embedded(const nova_head_t* h, const void* x_emb, const float* t, int32_t t_per_token,
                                        const void* z, int64_t B, int64_t N, float* v_out, void* workspace,
                                        size_t workspace_bytes, void* stream);

/*
 * The fused sampling loop (denoise): S Euler steps of the head with the condition
 * projection hoisted out of the loop and the latent kept in fp32.
 *   noise_tok [Bx, N, T] fp32   initial latent, token layout
 *   z         [B, N, Dc]        B = Bx (no guidance), 2*Bx ([cond; uncond]) or 3*Bx (three-pass guidance)
 *   pred_ids  [B, n] int64 or NULL (rows of the later passes must repeat the first Bx rows)
 *   timesteps [S] host fp32, sigmas [S+1] host fp64  (the scheduler's own values)
 *   x_out     [Bx, N, T] fp32   patchify(x_S).  Tokens outside pred_ids follow the
 *                               reference's x <- x + dt*x recurrence (two roundings per step).
 * The S-step loop touches only `workspace`; when the same (workspace address, shapes, schedule, guidance)
 * comes back the library replays a CUDA graph it captured on the second call (bit-identical results,
 * NOVA_B200_GRAPH=0 disables).  Keep the workspace address stable across calls to benefit.  If `stream` is
 * itself being captured the launches go into the caller's graph instead.
 */
NOVA_API int nova_head_sample(const nova_head_t* h, const float* noise_tok, const void* z, const int64_t* pred_ids,
                     int64_t B, int64_t Bx, int64_t N, int64_t n, const float* timesteps_host,
                     const double* sigmas_host, int32_t num_steps, const nova_guidance* guidance,
                     float* x_out, void* workspace, size_t workspace_bytes, void* stream);

/*
 * Set-by-set generation for a fixed condition, scheduled on the device: the accumulation loop of
 * Transformer3DModel.generate_frame (diffnext/models/transformers/transformer_3d.py:123-133) with the mask bookkeeping
 * of MaskEmbed.get_pred_mask (diffnext/models/embeddings.py:262-270) as device-side gathers / scatters.
 *   order          [Bx, N] int64  generation order of every cloud (a permutation of 0..N-1: argsort of uniforms)
 *   set_sizes_host [num_sets] host int32: set i predicts the tokens order[:, first_i : first_i + n_i], first_i = sum of
 *                  the sizes before it; empty sets are skipped (transformer_3d.py:120)
 *   noise_tok      [Bx, N, T] fp32: the initial latent of every token (a set reads only its own positions, so one
 *                  N(0,1) draw per token has the distribution of the reference's fresh tensor per set)
 *   guidance_scales_host [live sets] host fp32 or NULL: the decayed guidance scale of every non-empty set
 *                  (GuidanceScaler.decay_guidance_scale, transformer_3d.py:124); NULL = guidance->scale for all.
 *                  guidance->renorm must be >= 1 here (its norms need every set's full noise tensor).
 *   x_out          [Bx, N, T] fp32: tokens of every set written at their positions (x += sample * pred_mask, :133);
 *                  positions no set covers are left untouched.
 * One host call per pass: the launches of all sets are captured into ONE CUDA graph on the second call with the same
 * workspace, buffers, shapes, schedule, set sizes and guidance, and replayed afterwards.
 */
NOVA_API int nova_head_generate_sets(const nova_head_t* h, const float* noise_tok, const void* z, const int64_t* order,
                                     int64_t B, int64_t Bx, int64_t N, const int32_t* set_sizes_host, int32_t num_sets,
                                     const float* timesteps_host, const double* sigmas_host, int32_t num_steps,
                                     const nova_guidance* guidance, const float* guidance_scales_host, float* x_out,
                                     void* workspace, size_t workspace_bytes, void* stream);

/* prev = model_output * dt + sample, elementwise, two roundings in `dtype` (nova_dtype). */
NOVA_API int nova_euler_step(const void* model_output, const void* sample, double dt, void* prev, int64_t numel,
                    int32_t dtype, void* stream);

/*
 * Chamfer primitive: for every point of a its nearest neighbour in b and vice versa.
 *   a [B, N, 3] fp32, b [B, M, 3] fp32
 *   d1 [B, N] fp32 = min_j |a_i - b_j|  (Euclidean, NOT squared), idx1 [B, N] int32 or NULL
 *   d2 [B, M] fp32 = min_i |a_i - b_j|,                           idx2 [B, M] int32 or NULL
 * Exact difference form sum((x-y)^2) in fp32 (not the |x|^2+|y|^2-2xy matrix form).
 */
NOVA_API int nova_chamfer_nn(const float* a, const float* b, int64_t B, int64_t N, int64_t M, float* d1, float* d2,
                    int32_t* idx1, int32_t* idx2, void* stream);

/* cd [B] fp64 = mean_i d1[b, i] + mean_j d2[b, j]: the two means of chamfer_distance (demo.py:50-53) over the
 * distances nova_chamfer_nn returned, accumulated in double in a fixed order (deterministic), one launch. */
NOVA_API int nova_chamfer_pair_mean(const float* d1, const float* d2, int64_t B, int64_t N, int64_t M, double* cd,
                                    void* stream);

/*
 * Neighbourhood ops on the Chamfer tiling (callers either side of the sampling path).
 * All distances are Euclidean (not squared), exact difference form in fp32; ties keep the lowest index.
 *
 * nova_knn: q [B, Nq, 3], t [B, Nt, 3] fp32 -> dist [B, Nq, k] ascending, idx [B, Nq, k] int32 or NULL.
 *   Replaces the reference's topk(cdist(q, t), k, dim=-1, largest=False)
 *   (diffnext/models/transformers/transformer_pointcloud_nova.py:84-86,143-144).  1 <= k <= min(32, Nt).
 *
 * nova_local_density: points [B, N, 3] -> density [B, N] = mean of the k_neighbors smallest distances after
 *   the smallest one (self) is dropped.  Replaces compute_local_density(points, k_neighbors=8)
 *   (transformer_pointcloud_nova.py:81-89).  k_neighbors + 1 <= min(32, N), else an error, as topk raises there.
 *
 * nova_softmax_interp: targets [B, S, 3], points [B, N, 3] -> out [B, S, 3],
 *   out_i = sum_j softmax_j(-|t_i - p_j|) p_j.  Replaces the weighted average of feature_aware_interpolation
 *   (transformer_pointcloud_nova.py:142-150; the top-k indices computed there are not used by its result).
 */
NOVA_API int nova_knn(const float* q, const float* t, int64_t B, int64_t Nq, int64_t Nt, int32_t k, float* dist,
                      int32_t* idx, void* stream);
NOVA_API int nova_local_density(const float* points, int64_t B, int64_t N, int32_t k_neighbors, float* density,
                                void* stream);
NOVA_API int nova_softmax_interp(const float* targets, const float* points, int64_t B, int64_t S, int64_t N,
                                 float* out, void* stream);
/*
 * nova_farthest_point_sampling: points [B, N, 3], start_idx [B] int64 or NULL (= point 0) -> picked [B, num_samples]
 *   int64: picked[0] = start, picked[i] = the point farthest (squared distance, lowest index on ties) from the
 *   points picked so far.  This is the textbook algorithm farthest_point_sampling
 *   (transformer_pointcloud_nova.py:100-125) is named after; the reference's own loop takes `min` over a distance
 *   matrix that still holds its zero diagonal and therefore returns [start, 0, 0, ...] in exact arithmetic -- that
 *   form needs no kernel (the Python mirror offers it as mode="reference").  N <= 14000.
 */
NOVA_API int nova_farthest_point_sampling(const float* points, const int64_t* start_idx, int64_t B, int64_t N,
                                          int32_t num_samples, int64_t* picked, void* stream);

/*
 * nova_emd: earth mover's distance of equal-size clouds, a [B, N, 3], b [B, N, 3] fp32 -> emd [B] = mean distance of
 *   the minimum-cost perfect matching; assign [B, N] int32 or NULL (object of b matched to point i of a);
 *   status [B] int32 or NULL (bidding rounds used; negative: round budget exhausted, the matching is completed
 *   greedily and is not optimal).  Replaces dist = cdist(a, b); linear_sum_assignment(dist); mean(dist[rows, cols])
 *   (emd_approx train_newloss.py:352-377, earth_mover_distance demo.py:57-74, test_optimize.py:395-414).
 *   The reference solves it with the Hungarian method on the CPU; here it is Bertsekas' auction algorithm with
 *   epsilon scaling, one CTA per pair: the matching cost is within N * eps_final of the optimum (the mean within
 *   eps_final; the Python mirror uses 1e-5).  N <= 4096.  Deterministic.
 */
NOVA_API int nova_emd(const float* a, const float* b, int64_t B, int64_t N, float eps_final, int32_t max_rounds,
                      float* emd_out, int32_t* assign_out, int32_t* status_out, void* stream);

/*
 * Training-mode arithmetic either side of the head, forward only.
 *
 * nova_add_noise: x, noise [tokens, T] fp32, t_idx [tokens] int64 into the scheduler's training tables
 *   sigma_table / t_table [n_train] fp32 -> x_t = sigma * noise + (1 - sigma) * x (rounded like the reference:
 *   mul, mul, add) and, when t_out != NULL, t_out [tokens] = t_table[t_idx] (the per-token timestep the head takes).
 *   Replaces FlowMatchEulerDiscreteScheduler.add_noise (diffnext/schedulers/scheduling_cfm.py:106-117).
 *
 * nova_flow_loss: pred, noise, x [tokens, T] fp32, weight [tokens] fp32 or NULL (all ones) ->
 *   loss_tok [tokens] = mean_T((pred - (noise - x))^2) * weight / (sum(weight) + 1e-5), scratch2[0] = sum(loss_tok),
 *   scratch2[1] = sum(weight).  Deterministic (fixed-order single-block reductions).
 *   Replaces the loss of Transformer3DModel.get_losses (diffnext/models/transformers/transformer_3d.py:91-95).
 */
NOVA_API int nova_add_noise(const float* x, const float* noise, const float* sigma_table, const float* t_table,
                            const int64_t* t_idx, int64_t tokens, int32_t T, int32_t n_train, float* x_t,
                            float* t_out, void* stream);
NOVA_API int nova_flow_loss(const float* pred, const float* noise, const float* x, const float* weight,
                            int64_t tokens, int32_t T, float* loss_tok, float* scratch2, void* stream);

/*
 * Training step of the head: forward with saved activations, then the full backward pass --
 * what autograd does for the reference between Transformer3DModel.get_losses and loss.backward()
 * (diffnext/models/transformers/transformer_3d.py:79-100) through DiffusionMLP.forward with per-token timesteps
 * (diffnext/models/diffusion_mlp.py:56-99).  Rows are tokens: rows = batch x tokens, every row has its own timestep.
 *
 * nova_head_train_bytes:   bytes of the workspace both calls carve identically (saved activations + backward scratch).
 * nova_head_train_forward: x_tok [rows, T] fp32 noisy latent (token layout), t [rows] fp32, z [rows, Dc] (handle
 *                          dtype) -> v_out [rows, T] fp32; the activations the backward needs stay in `workspace`.
 * nova_head_backward:      dv [rows, T] fp32 = dLoss/dv, with the SAME x_tok, z and workspace as the forward call ->
 *                          grads[k] (fp32, reference state_dict shape of names[k], WRITTEN not accumulated; any subset
 *                          of the 14 + 8*depth keys, NULL entries skipped) and dz_out [rows, Dc] (handle dtype) or NULL.
 * Width must be a multiple of 256.  Matrix products run on the tcgen05 GEMM (bf16 handle) or the SIMT fp32 GEMM
 * (fp32 parity handle); weight gradients split their reduction over the rows into one batched launch.
 */
NOVA_API size_t nova_head_train_bytes(const nova_head_t* h, int64_t rows);
NOVA_API int nova_head_train_forward(const nova_head_t* h, const float* x_tok, const float* t, const void* z, int64_t rows,
                                     float* v_out, void* workspace, size_t workspace_bytes, void* stream);
NOVA_API int nova_head_backward(const nova_head_t* h, const float* dv, const float* x_tok, const void* z, int64_t rows,
                                int32_t n_grads, const char* const* names, float* const* grads, void* dz_out,
                                void* workspace, size_t workspace_bytes, void* stream);

/*
 * Multi-GPU: clouds are sharded data-parallel over one process per GPU, weights replicated, no collective inside the
 * denoise loop; the ONE collective of the path is an all-gather of the generated points after the last Euler step.
 * NCCL is resolved at first use (dlopen of libnccl.so.2; inside a PyTorch process the already-loaded NCCL is used).
 *
 * nova_comm_unique_id:  rank 0 fills a 128-byte id and ships it to the other ranks by any means (file, socket, MPI).
 * nova_comm_init_rank:  every rank, with its own device current (cudaSetDevice), joins the communicator.
 * nova_allgather:       recv [world_size * count_bytes] <- every rank's send [count_bytes], in rank order, on `stream`
 *                       (thin ncclAllGather; in-place when send == recv + rank * count_bytes).
 */
typedef struct nova_comm nova_comm_t; /* opaque (an ncclComm_t) */
NOVA_API int nova_comm_unique_id(char* out128);
NOVA_API int nova_comm_init_rank(const char* id128, int32_t world_size, int32_t rank, nova_comm_t** out);
NOVA_API int nova_comm_destroy(nova_comm_t* comm);
NOVA_API int nova_allgather(nova_comm_t* comm, const void* send, void* recv, int64_t count_bytes, void* stream);

/* Kernels launched by this library in the calling thread since the last reset (for bench.py). */
NOVA_API int64_t nova_launch_count(void);
NOVA_API void nova_launch_count_reset(void);

/*
 * In-situ kernel timing for bench.py: while enabled (per calling thread), CUDA events are recorded on
 * the launching stream around every launch of these kernel classes:
 *   0 = AdaLN statistics GEMMs (the modulation GEMM M x 2D x D of the fused step; M x 20D x D in the wide dataflow),
 *   1 = other GEMMs (fc1 / fc2 / condition), 2 = row kernels, 3 = prep, 4 = other,
 *   5 = the cluster chain kernel (small M: all stages of a step after the statistics GEMM in one launch),
 *   6 = the gate GEMM M x D x D with the block tail in its epilogue.
 * nova_profile_read sums the elapsed ms and launches per class (arrays of >= 7 entries), then clears.
 */
NOVA_API int nova_profile_enable(int32_t on);
NOVA_API int nova_profile_read(double* ms_by_class, int64_t* launches_by_class, int32_t n_classes);

/*
 * Test hook: C[M,N] = epilogue(A[M,K] W[N,K]^T + bias) with one named GEMM implementation.
 *   impl 0 = SIMT, 1 = tcgen05 cta_group::1, 2 = tcgen05 cta_group::2 (CTA pairs), 3 = tcgen05 default
 *   (1-3 bf16 only); epilogue 0 = bias, 1 = bias + SiLU.
 *   A, W, C are bf16 (dtype NOVA_BF16) or fp32 (NOVA_F32); bias fp32 [N] or NULL.
 */
NOVA_API int nova_debug_gemm(const void* A, const void* W, const float* bias, void* C, int64_t M, int64_t N, int64_t K,
                    int32_t dtype, int32_t impl, int32_t epilogue, void* stream);

/*
 * Test hook: the AdaLN statistics GEMM with the modulation fused into its epilogue (bf16, tcgen05).
 *   A [M,K], W [n_stats*D, K] and bias [n_stats*D] in the reference order (scale | shift | gate), x [M,D];
 *   h_out [M,D] = LN(x; 1e-6)(1 + scale) + shift;  gate_out [M,D] = gate (n_stats == 3) or NULL (n_stats == 2).
 * Synchronises the stream (it owns temporary buffers).
 */
NOVA_API int nova_debug_adaln_gemm(const void* A, const void* W, const float* bias, const void* x, void* h_out,
                                   void* gate_out, int64_t M, int64_t D, int64_t K, int32_t n_stats,
                                   int32_t cta_group, void* stream);

/*
 * Test hook: SM-clock stamps (cycles since kernel entry) of cluster 0 / CTA 0 of the last chain-kernel launch
 * (csrc/chain_tcgen05.cu), 8 slots per stage; recorded only when NOVA_B200_CHAIN_TIMELINE=1.  Synchronises the device.
 */
NOVA_API int nova_debug_chain_timeline(int64_t* out, int32_t n);

/*
 * The 4 host-mapped status words of the library: [0..2] are written by a tcgen05 kernel before it traps on a barrier
 * timeout (0xDEAD0000 | code, block, parity); [3] becomes 0xBAD1D5 when a gather / scatter kernel met a pred_id outside
 * [0, N).  Such an id is never used for indexing (the row reads token 0 / is not written), so nothing is accessed out
 * of bounds; the reference's gather raises there, a C caller checks this word after synchronising.
 */
NOVA_API int nova_debug_words(uint32_t* out4);
NOVA_API int nova_debug_words_clear(void);

#ifdef __cplusplus
}
#endif
#endif /* NOVA_B200_H_ */
