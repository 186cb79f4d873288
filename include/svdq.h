/*
 * svdq.h — C ABI of libsvdq.so, the B200 (sm_100a) implementation of the SVD-Hybrid merge hot
 * path of mgradyn/SVD-Quantization-Task-Merging.
 *
 * The reference has no FFI: its "operator API" is a set of module-level Python functions
 * (src/svd_hybrid/*.py, root quantization_utils.py).  Each entry point below names the
 * reference functions (file:line, relative to the reference repository root) whose arithmetic
 * it replaces; INTEGRATION.md shows the ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes only; every pointer is a DEVICE pointer unless the
 *     name ends in _host.  The library never allocates and never synchronises the device: the
 *     caller owns all buffers and passes the CUDA stream (cudaStream_t as void*).
 *   - return value: 0 = ok; < 0 = invalid argument (Python side raises ValueError);
 *     > 0 = cudaError_t (Python side raises RuntimeError).  svdq_last_error() returns a
 *     thread-local message for the last non-zero return.
 *   - dtype codes: 0 = float32, 1 = bfloat16, 2 = float16 (dtype of base / fine-tuned tensors).
 *   - mask strategy codes: 0 = union, 1 = intersection, 2 = majority (votes >= 0.5 * n_present).
 *   - NT = n_tasks is the stride of every per-task array; n_tasks <= 16 for svdq_tv_mask_gram (and for
 *     svdq_gram_reduce with a second Gram block), <= 32 everywhere else (see "wide path").
 *   - "tile" = tile_elems consecutive elements of one parameter (tile_elems % 1024 == 0);
 *     tiles are numbered parameter by parameter: tile_begin[p] .. tile_begin[p+1]-1.
 *   - tensor pointer tables: tensors[p*(NT+1) + 0] = base, [.. + 1 + t] = fine-tuned tensor of
 *     task t (NULL when task t lacks the parameter).  Tensor and mask pointers must be 16-byte
 *     aligned (source alignment of the TMA bulk copies of the staged kernels).
 */
#ifndef SVDQ_H_
#define SVDQ_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SVDQ_ABI_VERSION 7   /* 7: K14 operators (svdq_basis_project/expand, svdq_mask_offsets/select/scatter); 6: svdq_reconstruct_merge_basis; 5: svdq_mask_tile_counts, svdq_reload_merge; 4: svdq_host_kmeans, per-cluster svdq_param_average; 3: svdq_tv_mask_gram_bits, svdq_host_pack_mask */
#define SVDQ_MAX_STREAM_TASKS 16
#define SVDQ_MAX_TASKS 32
#define SVDQ_MAX_STAGES 8

int svdq_abi_version(void);
const char* svdq_last_error(void);

/* size in bytes of the scratch record svdq_rtvq_quantize / svdq_absmax_quantize need per CTA,
 * and the number of records (scratch = svdq_k4_scratch_bytes() bytes) */
int64_t svdq_k4_scratch_bytes(void);

/*
 * K1 — task-vector construction + tall-mask combination + masked Gram, one streaming pass.
 * Replaces: compute_task_vector (src/svd_hybrid/task_vector_loader.py:103-145),
 *           combine_masks / compute_{union,intersection,majority}_mask
 *           (src/svd_hybrid/mask_loader.py:412-485,488-648), apply_mask_to_tensor
 *           (mask_loader.py:651-679), stack_and_center and the T^T T half of compute_svd
 *           (src/svd_hybrid/basis.py:63-113,216-249).
 * masks: [P*NT] table of torch.bool storages (NULL entry = task has no mask for the parameter),
 *        or NULL when there are no masks at all.
 * packed/pmask_off: bit-packed combined mask, parameter p at word offset pmask_off[p] (a multiple of 4 words
 *       lets pass 2 stream the mask words through its TMA ring; any offset is accepted).
 * gram: [n_tiles][full ? 2 : 1][NT(NT+1)/2] fp32 partials (upper triangle, row-major);
 *       full == 1: the second block is the Gram over ALL elements (masked or not), the
 *       whole-model task Gram behind cluster weighting (src/svd_hybrid/clustering.py:227-232);
 *       full == 2: the second block is the Gram over the elements OUTSIDE the combined mask, the
 *       noise region of svd_include_noise (get_unmasked_portion, src/svd_hybrid/mask_loader.py:682-709;
 *       src/svd_hybrid/cli.py:336-338) -- svdq_gram_reduce then forms all = masked + complement.
 * count: [n_tiles] masked elements per tile.
 */
int svdq_tv_mask_gram(int dtype, int n_tasks, int mask_strategy, int full, int64_t n_tiles, int tile_elems,
                      const void* const* tensors, const uint8_t* const* masks, const int64_t* numel,
                      const int32_t* tile_param, const int32_t* tile_local, const int64_t* pmask_off,
                      uint32_t* packed, float* gram, uint32_t* count, void* stream);

/*
 * K1 with BIT-packed task masks: identical to svdq_tv_mask_gram except that every non-NULL entry of mask_bits
 * points to ceil(numel/8) bytes (padded to a multiple of 16, 16-byte aligned) holding element 8k+i in bit i of
 * byte k.  For inputs staged from host memory (load_task_masks, src/svd_hybrid/mask_loader.py:242-409, yields
 * host torch.bool tensors): the masks then cross PCIe at one bit per element.  svdq_host_pack_mask is the
 * matching host-side encoder (src/dst are HOST pointers; pure re-encoding of one mask, no combination -- the
 * vote of combine_masks stays on the device).
 */
int svdq_tv_mask_gram_bits(int dtype, int n_tasks, int mask_strategy, int full, int64_t n_tiles, int tile_elems,
                           const void* const* tensors, const uint8_t* const* mask_bits, const int64_t* numel,
                           const int32_t* tile_param, const int32_t* tile_local, const int64_t* pmask_off,
                           uint32_t* packed, float* gram, uint32_t* count, void* stream);
int svdq_host_pack_mask(const uint8_t* src_host, int64_t n, uint8_t* dst_host, int n_threads);
/* the same for a batch of masks (what one rank of a parameter-sharded merge owns): src_host[i] / n[i] / dst_host[i],
 * i < count; the batch's bytes are split evenly over n_threads host threads spawned once */
int svdq_host_pack_mask_batch(const uint8_t* const* src_host, const int64_t* n, uint8_t* const* dst_host, int64_t count,
                              int n_threads);

/*
 * Host-side k-means of cluster weighting (all pointers are HOST pointers; no CUDA call).
 * Replaces: compute_kmeans_clustering (src/svd_hybrid/clustering.py:123-156), i.e. the call
 *           sklearn.cluster.KMeans(n_clusters=k, random_state=seed, n_init=n_init).fit_predict(features)
 *           that cluster_tasks (clustering.py:198-245, from cli.py:530) makes with seed 42, n_init 10.
 * features_host: row-major [n x d] float32 (the engine passes an isometric n x n embedding of the whole-model
 * task Gram that K1 accumulates, rows = tasks in sorted-name order like clustering.py:87).  The procedure is
 * scikit-learn 1.9's, step by step (numpy RandomState stream, k-means++ with 2 + int(log k) local trials,
 * float32 Lloyd iterations, max_iter / tol stopping rules, first-best-wins selection over the n_init runs), so
 * labels_host [n] equals sklearn's labels_, not just its partition.  inertia_host (optional) = inertia_.
 */
int svdq_host_kmeans(const float* features_host, int n, int d, int k, uint32_t seed, int n_init, int max_iter,
                     double tol, int32_t* labels_host, double* inertia_host);

/*
 * Wide path, 17..32 task vectors (the Gram accumulators of that many tasks do not fit one thread):
 * svdq_mask_pack combines the N tall masks once (same reference lines as K1's mask part) into the packed
 * mask + per-tile counts; svdq_gram_staged then accumulates the Gram of all N tasks under that mask in one
 * pass over the inputs.  mask_mode selects the rows that enter the Gram: 0 = inside the combined mask, 1 = all
 * rows (whole-model Gram of cluster weighting), 2 = outside the mask (noise region).  svdq_gram_reduce
 * (full = 0), svdq_param_solve, svdq_project_exact, svdq_reconstruct_merge and svdq_diag_finalize accept
 * n_tasks <= 32 directly.
 */
int svdq_mask_pack(int n_tasks, int mask_strategy, int64_t n_tiles, int tile_elems, const uint8_t* const* masks,
                   const int64_t* numel, const int32_t* tile_param, const int32_t* tile_local,
                   const int64_t* pmask_off, uint32_t* packed, uint32_t* count, void* stream);
/* Single-pass Gram of ALL n_tasks <= 32 task vectors under the pre-combined mask: the inputs are read once,
 * staged through shared memory, and the 8x8 task blocks of the Gram are spread over the warps of a persistent CTA.
 * gram: [n_tiles][n(n+1)/2]; tile_elems must be a multiple of 1536; svdq_gram_reduce (full = 0) reduces it for
 * n_tasks <= 32.  Same reference lines as svdq_tv_mask_gram. */
int svdq_gram_staged(int dtype, int n_tasks, int mask_mode, int64_t n_tiles, int tile_elems,
                     const void* const* tensors, const int64_t* numel, const int32_t* tile_param,
                     const int32_t* tile_local, const int64_t* pmask_off, const uint8_t* has_mask,
                     const uint32_t* packed, float* gram, void* stream);

/*
 * K2a — fixed-order fp64 reduction of K1's tile partials per parameter.
 * gram_masked / gram_all: [P][NT*NT] full symmetric fp64 (gram_all may be NULL); dm: [P].
 * full == 2 (noise region): gram_noise [P][NT*NT] = Gram of the rows outside the mask, dm_noise [P] = their
 * count, non-zero only where the reference builds a noise basis (src/svd_hybrid/cli.py:330-338,
 * src/svd_hybrid/basis.py:455-466): the parameter has a mask, the masked count passes min_mask_size, and
 * at least one element is unmasked.  numel / has_mask / gram_noise / dm_noise may be NULL otherwise.
 * gram is CONSUMED: parameters with >= 512 tiles are reduced in two levels (64 tile ranges over 64 CTAs, then the
 * range sums in order) and the fp64 range sums overwrite the head of each range; count is left intact
 * (svdq_basis_offsets reads it).  The split depends only on the parameter's own tile count, so results do not
 * depend on how parameters are sharded.
 */
int svdq_gram_reduce(int n_tasks, int full, int64_t n_params, int min_mask_size, float* gram,
                     const uint32_t* count, const int64_t* tile_begin, const int64_t* numel, const uint8_t* has_mask,
                     double* gram_masked, double* gram_all, int64_t* dm, double* gram_noise, int64_t* dm_noise,
                     void* stream);

/*
 * K2b — per-parameter solve (one warp per parameter).
 * Replaces: the rest of compute_svd (src/svd_hybrid/basis.py:241), compute_energy_spectrum /
 *           select_rank (basis.py:116-213), construct_basis glue (basis.py:252-409),
 *           project_to_basis / compress_single_task (src/svd_hybrid/compress.py:6-56),
 *           asymmetric_quantization / multistage_residual_quantization
 *           (src/svd_hybrid/rtvq.py:4-103) on the low-energy block, dequantize_and_average
 *           (src/svd_hybrid/merge.py:61-141) and the gating of src/svd_hybrid/cli.py:319-343.
 * Inputs : gram_masked [P][NT*NT], dm [P], has_mask [P] (uint8), present [P] (bit t: task t has
 *          the parameter), weights [NT] fp64 by task position (NULL: see svdq_param_average),
 *          avg_order [NT] task positions in
 *          sorted-name order, sign_ref [P][NT*NT] optional Vh_ref[j][t] for test-only sign
 *          alignment (NULL in production).
 * Outputs (per parameter, stride NT; S = rtvq_stages):
 *   info [P][8] int32  : status (0 solved, 1 skipped: mask below min_mask_size, 2 empty),
 *                        n_active, r = min(Dm, n_active), k, r_eff, 0, 0, 0
 *   sv [P][NT] fp32 singular values; scal [P][4] = energy_retained, tail_add, mean_scale, 0
 *   coef [P][NT][NT] raw coefficients coef[t][j]; chigh [P][NT][NT] fp16 bits (j < k)
 *   codes [P][NT][S][NT] uint8 (i < r-k); qscale/qzp/qres [P][NT][S]
 *   chat [P][NT][NT] coefficients after fp16 / RTVQ round trip; cbar [P][NT] weighted average
 *   W [P][NT][NT] with u_d[j] = sum_t (tau_d[t] - mean_d) W[t][j]; gvec [P][NT] = W cbar
 *   V [P][NT][NT] fp64 right singular vectors V[t][j]
 */
int svdq_param_solve(int n_tasks, int64_t n_params, int center, float energy_threshold, int max_rank,
                     int min_mask_size, int rtvq_bits, int rtvq_stages,
                     const double* gram_masked, const int64_t* dm, const uint8_t* has_mask, const uint32_t* present,
                     const double* weights, const int32_t* avg_order, const double* sign_ref,
                     int32_t* info, float* sv, float* scal, float* coef, uint16_t* chigh, uint8_t* codes,
                     float* qscale, float* qzp, float* qres, float* chat, float* cbar, float* W, float* gvec,
                     double* V, void* stream);

/*
 * K2c — weighted average only: svdq_param_solve called with weights == NULL stops after the
 * coefficients / RTVQ / W; this entry point then forms cbar, gvec and scal[1] once the task weights
 * are known (cluster weighting derives them from the whole-model Gram on the host meanwhile).
 * Replaces dequantize_and_average (src/svd_hybrid/merge.py:89-141) and merge_with_clustering
 * (merge.py:586-626) + merge_cluster_results (src/svd_hybrid/clustering.py:374-425).
 * cluster_of / cluster_omega (both NULL = no clustering): cluster_of[t] in [0, NT) is the cluster index of task
 * position t, cluster_omega[c] the cross-cluster weight (softmax of the mean member weight, renormalised over the
 * clusters).  Per parameter the member weights are renormalised over the members of each cluster that HAVE the
 * parameter; a cluster none of whose members has it contributes zeros, mean included, exactly like the
 * reference (merge.py:289-290): scal[2] receives the resulting factor on the mean (1 otherwise), which
 * svdq_reconstruct_merge applies.
 */
int svdq_param_average(int n_tasks, int64_t n_params, const uint32_t* present, const double* weights,
                       const int32_t* avg_order, const int32_t* cluster_of, const double* cluster_omega,
                       const int32_t* info, const float* chat, const float* W, float* cbar,
                       float* gvec, float* scal, void* stream);

/*
 * K7 — exact projection on the STORED basis for selected (small) parameters, n_tasks <= 32.
 * Replaces project_to_basis / compress_single_task (src/svd_hybrid/compress.py:6-56) literally: the basis row is
 * rebuilt, rounded to fp16 when fp16_basis (the cast of src/svd_hybrid/cli.py:355-361 precedes the projection in
 * the reference) and contracted with (tau_t - mean) over the masked rows.  sel_tile_param / sel_tile_local list
 * the tiles of the selected parameters; proj receives one [NT*NT] partial c[t][j] per tile.
 * svdq_param_requantize then sums the partials of parameter p (tiles sel_tile_begin[p] .. sel_tile_begin[p+1]-1,
 * an empty range = keep the closed-form coefficients), overwrites coef and redoes chigh / codes / chat.
 * Call order: svdq_param_solve(weights = NULL) -> svdq_project_exact -> svdq_param_requantize -> svdq_param_average.
 * region: 0 = rows inside the combined mask, 1 = rows outside it (noise region; info / W of the noise solve).
 */
int svdq_project_exact(int dtype, int n_tasks, int fp16_basis, int center, int region, int64_t n_sel_tiles,
                       int tile_elems,
                       const void* const* tensors, const int64_t* numel, const int32_t* sel_tile_param,
                       const int32_t* sel_tile_local, const int64_t* pmask_off, const uint8_t* has_mask,
                       const uint32_t* packed, const int32_t* info, const float* W, float* proj, void* stream);
int svdq_param_requantize(int n_tasks, int64_t n_params, int rtvq_bits, int rtvq_stages, const int64_t* sel_tile_begin,
                          const float* proj, const uint32_t* present, const int32_t* info, float* coef,
                          uint16_t* chigh, uint8_t* codes, float* qscale, float* qzp, float* qres, float* chat,
                          void* stream);

/*
 * K3 — weighted reconstruction + merge (pass 2), optional fused diagnostics.
 * Replaces: reconstruct_from_coefficients (src/svd_hybrid/merge.py:144-194), the fp16 cast of
 *           the bases (src/svd_hybrid/cli.py:355-361), reconstruct_from_masked
 *           (src/svd_hybrid/mask_loader.py:712-763), merge_parameter / merge_all_parameters
 *           (merge.py:197-426), apply_merged_deltas (merge.py:429-552) and, with diag != 0,
 *           compute_reconstruction_error / compute_parameter_diagnostics
 *           (src/svd_hybrid/diagnostics.py:72-231).
 * out: [P] table of fp32 output tensors (merged = base + delta; parameters without a basis get
 *      a copy of base).  diag_partials: [n_tiles][4][NT] fp32 (sum e^2, sum |e|, sum rec^2, max |e| per task;
 *      may be NULL when diag == 0).
 * noise_*: outputs of a second svdq_param_solve over gram_noise / dm_noise (svd_include_noise): the unmasked
 *      positions then receive noise_shrink * (U_noise cbar_noise + mean) instead of 0 (merge.py:257-284,
 *      mask_loader.py:757-760).  All NULL = no noise region.  The diagnostics stay those of the masked region.
 */
int svdq_reconstruct_merge(int dtype, int n_tasks, int fp16_basis, int diag, int center, int64_t n_tiles,
                           int tile_elems, const void* const* tensors, const int64_t* numel,
                           const int32_t* tile_param, const int32_t* tile_local, const int64_t* pmask_off,
                           const uint8_t* has_mask, const uint32_t* packed, const int32_t* info, const float* W,
                           const float* cbar, const float* gvec, const float* scal, const float* chat,
                           float* const* out, float* diag_partials, const int32_t* noise_info, const float* noise_W,
                           const float* noise_cbar, const float* noise_gvec, const float* noise_scal,
                           float noise_shrink, void* stream);

/* diagnostics finalisation: out [P][NT][6] fp64 = absolute_error, relative_error,
 * max_absolute_error, mean_absolute_error, original_norm, reconstructed_norm
 * (src/svd_hybrid/diagnostics.py:110-117).  original_norm^2 is read off the diagonal of gram_masked
 * (the masked task vector's squared norm, already reduced in fp64 by svdq_gram_reduce). */
/*
 * K3c with the basis write-out fused in (the reference's default settings: svd_eval_reconstruction AND
 * svd_store_artifacts): svdq_reconstruct_merge(diag = 1) that ALSO stores the artifact bases of the masked region --
 * U_high [Dm x k], U_low [Dm x (r-k)] (fp16 when fp16_basis else fp32) and mean [Dm] (NULL table = no mean), rows
 * compacted to the elements inside the combined mask, exactly what svdq_write_basis would store (construct_basis
 * src/svd_hybrid/basis.py:363-364,398-407 after apply_mask_to_tensor src/svd_hybrid/mask_loader.py:675-679) -- so the
 * inputs are not read a third time.  tile_row_off from svdq_basis_offsets(region = 0).  Up to 8 tasks, no noise region.
 * chat and diag_partials both NULL = no diagnostics (svd_store_artifacts alone).
 */
int svdq_reconstruct_merge_basis(int dtype, int n_tasks, int fp16_basis, int center, int64_t n_tiles, int tile_elems,
                                 const void* const* tensors, const int64_t* numel, const int32_t* tile_param,
                                 const int32_t* tile_local, const int64_t* pmask_off, const uint8_t* has_mask,
                                 const uint32_t* packed, const int32_t* info, const float* W, const float* cbar,
                                 const float* gvec, const float* scal, const float* chat, float* const* out,
                                 float* diag_partials, const int64_t* tile_row_off, void* const* u_high,
                                 void* const* u_low, float* const* mean, void* stream);

int svdq_diag_finalize(int n_tasks, int64_t n_params, const float* diag_partials, const int64_t* tile_begin,
                       const int64_t* dm, const int32_t* info, const double* gram_masked, double* out, void* stream);

/*
 * K5 — materialise the bases in the reference artifact layout
 * (U_high [Dm x k], U_low [Dm x (r-k)], mean [Dm x 1], rows compacted through the mask:
 * src/svd_hybrid/basis.py:363-364,398-407; storage layout src/svd_hybrid/storage.py:76-106).
 * tile_row_off [n_tiles] is produced by svdq_basis_offsets from K1's counts.
 * region: 0 = the rows inside the combined mask; 1 = the rows outside it (the "noise" basis of
 * src/svd_hybrid/basis.py:455-466, with info / W of the noise solve; numel is needed for the offsets).
 */
/*
 * K11 -- merge from STORED artifacts (the reload path).
 * Replaces: merge_all_parameters / merge_parameter / reconstruct_from_coefficients (src/svd_hybrid/merge.py:144-426)
 *           as called by reconstruct_from_artifacts (src/svd_hybrid/reload.py:142-238), and the scatter of
 *           reconstruct_from_masked (src/svd_hybrid/mask_loader.py:712-763), for ALL parameters in one launch.
 * svdq_mask_tile_counts: masked elements per tile from the packed combined masks (count [n_tiles]); feed it to
 *           svdq_basis_offsets for tile_row_off.
 * svdq_reload_merge: kr [P][2] = (k, r) per parameter ((0, 0) = no basis -> zeros); u_high / u_low / mean [P] point to
 *           the stored bases (fp16 when basis_fp16 else fp32; mean fp32 or NULL), rows compacted to the region's
 *           elements; cbar [P][n_tasks] = averaged coefficients, c_high then c_low (dequantize_and_average,
 *           merge.py:61-141, on the host: a few numbers per parameter); out [P] = fp32 deltas.
 *           region 0: delta at the rows inside the mask, zeros elsewhere; region 1 (noise basis, merge.py:257-284):
 *           scale * value at the rows outside the mask, the rest of out untouched (call after region 0).
 */
int svdq_mask_tile_counts(int64_t n_tiles, int tile_elems, const int64_t* numel, const int32_t* tile_param,
                          const int32_t* tile_local, const int64_t* pmask_off, const uint8_t* has_mask,
                          const uint32_t* packed, uint32_t* count, void* stream);
int svdq_reload_merge(int basis_fp16, int n_tasks, int region, float scale, int64_t n_tiles, int tile_elems,
                      const int64_t* numel, const int32_t* tile_param, const int32_t* tile_local,
                      const int64_t* pmask_off, const uint8_t* has_mask, const uint32_t* packed, const int32_t* kr,
                      const void* const* u_high, const void* const* u_low, const float* const* mean, const float* cbar,
                      const int64_t* tile_row_off, float* const* out, void* stream);

int svdq_basis_offsets(int64_t n_params, int region, int tile_elems, const uint32_t* count, const int64_t* tile_begin,
                       const int64_t* numel, int64_t* tile_row_off, void* stream);
int svdq_write_basis(int dtype, int n_tasks, int fp16_basis, int center, int region, int64_t n_tiles, int tile_elems,
                     const void* const* tensors, const int64_t* numel, const int32_t* tile_param,
                     const int32_t* tile_local, const int64_t* pmask_off, const uint8_t* has_mask,
                     const uint32_t* packed, const int32_t* info, const float* W, const int64_t* tile_row_off,
                     void* const* u_high, void* const* u_low, float* const* mean, void* stream);

/*
 * K4 — quantisers on arbitrarily long fp32 tensors.
 * svdq_rtvq_quantize : asymmetric_quantization (src/svd_hybrid/rtvq.py:4-27 =
 *                      quantization_utils.py:76-99) for stages == 1, and
 *                      multistage_residual_quantization (rtvq.py:39-82) in general.
 *                      codes: [stages][codes_ld] uint8 (code_bytes 1) or int16 (code_bytes 2);
 *                      scale / zp / resnorm: [stages] fp32 device scalars.
 * svdq_rtvq_dequantize: asymmetric_dequantization / multistage_residual_dequantization
 *                      (rtvq.py:29-36,85-103; quantization_utils.py:137-172).
 * svdq_absmax_quantize: absmax_quantization (quantization_utils.py:60-73); q int8 / int16.
 * scratch: svdq_k4_scratch_bytes() bytes of device memory.
 * packed (optional, may be NULL): additionally receives the codes bit-packed, `bits` per code, [stages][packed_ld]
 *          32-bit words (bits in {1,2,4,8}); assembled with warp shuffles.  The reference stores one uint8 per
 *          code, so this is an extra, denser copy -- `codes` is always written in the reference layout.
 */
int svdq_rtvq_quantize(const float* x, int64_t n, int bits, int stages, void* codes, int64_t codes_ld,
                       int code_bytes, float* scale, float* zp, float* resnorm, void* scratch, uint32_t* packed,
                       int64_t packed_ld, void* stream);
int svdq_rtvq_dequantize(const void* codes, int64_t codes_ld, int code_bytes, int stages, int64_t n,
                         const float* scale, const float* zp, float* out, void* stream);
int svdq_absmax_quantize(const float* x, int64_t n, int bits, void* q, int code_bytes, float* scale, void* scratch,
                         void* stream);

/*
 * Mask utilities behind the fine-grained API.
 * svdq_combine_masks: compute_union/intersection/majority_mask (src/svd_hybrid/mask_loader.py:412-485)
 *                     over n_masks torch.bool tensors of n elements -> torch.bool out.
 * svdq_unpack_mask  : packed combined mask -> torch.bool.
 */
int svdq_combine_masks(const uint8_t* const* masks, int n_masks, int64_t n, int strategy, uint8_t* out, void* stream);
int svdq_unpack_mask(const uint32_t* packed, int64_t n, uint8_t* out, void* stream);

/*
 * K14 -- the operator-by-operator API of the artifact tools (fine-grained mirrors of the reference functions).
 * svdq_basis_project : c[cols] = U^T (delta - mean), U = rows x cols (fp16 or fp32, row-major, leading dimension ld),
 *                      cols <= 32 per call; project_to_basis (src/svd_hybrid/compress.py:6-21) with the centring of
 *                      compress_single_task (:24-56; mean may be NULL).  Fixed-order reduction (fp64 above the
 *                      per-thread partial sums).  scratch: svdq_project_scratch_bytes() bytes.
 * svdq_basis_expand  : out[rows] = scale * (U_high c_high + U_low c_low + mean); reconstruct_from_coefficients
 *                      (src/svd_hybrid/merge.py:144-194); n_low may be 0, mean may be NULL.
 * svdq_mask_offsets  : chunk_off[i] = number of kept elements before chunk i (chunks of svdq_select_chunk_elems()
 *                      elements), chunk_off[n_chunks] = total; kept(i) = (mask[i] != 0) != invert.
 * svdq_mask_select   : out = x.flatten()[mask] in ascending element order (apply_mask_to_tensor /
 *                      get_unmasked_portion, src/svd_hybrid/mask_loader.py:665-709); elem_bytes in {1, 2, 4, 8}.
 * svdq_mask_scatter  : the inverse, out[i] = values[rank of i among the kept elements]; other elements of out are left
 *                      as they are (reconstruct_from_masked, src/svd_hybrid/mask_loader.py:712-763).
 */
size_t svdq_project_scratch_bytes(void);
int svdq_select_chunk_elems(void);
int svdq_basis_project(int basis_fp16, const void* u, int64_t ld, int cols, int64_t rows, const float* delta,
                       const float* mean, float* c, void* scratch, void* stream);
int svdq_basis_expand(int basis_fp16, const void* u_high, int64_t ld_high, int k, const void* u_low, int64_t ld_low,
                      int n_low, int64_t rows, const float* c_high, const float* c_low, const float* mean, float scale,
                      float* out, void* stream);
int svdq_mask_offsets(const uint8_t* mask, int64_t n, int invert, int64_t* chunk_off, void* stream);
int svdq_mask_select(const void* x, int elem_bytes, const uint8_t* mask, int64_t n, int invert, const int64_t* chunk_off,
                     void* out, void* stream);
int svdq_mask_scatter(const void* values, int elem_bytes, const uint8_t* mask, int64_t n, int invert,
                      const int64_t* chunk_off, void* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SVDQ_H_ */
