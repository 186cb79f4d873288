// Argument blocks and launch entry points shared by the kernel translation units and the C ABI.
#pragma once
#include "svdq_common.cuh"
#include "k2_core.h"

namespace svdq {

struct K1Args {
    const void* const* tensors;     // [P][NT+1]: base, ft_0 .. ft_{NT-1}; a null ft = task lacks the parameter
    const uint8_t* const* masks;    // [P][NT] torch.bool storage, null entries = task has no mask; may be null
    const int64_t* numel;           // [P]
    const int32_t* tile_param;      // [n_tiles]
    const int32_t* tile_local;      // [n_tiles] tile index inside its parameter
    const int64_t* pmask_off;       // [P] word offset of the parameter's packed combined mask
    uint32_t* packed;               // bit i of word w = combined mask of element 32 w + i
    float* gram;                    // [n_tiles][FULL ? 2 : 1][G]: masked Gram [, Gram over all elements]
    uint32_t* count;                // [n_tiles] masked elements in the tile
    int tile_elems;                 // multiple of kStep
    int strategy;
    // pre-combined mask mode (n_tasks > 16 runs the Gram in several task-subset launches that must all see
    // the SAME combined mask): masks == nullptr, packed_in = packed combined masks written by k6_mask_pack,
    // has_mask_in[P] says which parameters have one.  packed is not written in this mode.
    const uint32_t* packed_in;
    const uint8_t* has_mask_in;
    // FULL only: 0 = second block is the Gram over ALL elements, 1 = over the elements OUTSIDE the combined mask
    // (the noise region of svd_include_noise; all = masked + complement is then formed by k2_gram_reduce)
    int second_complement;
    // pre-combined mask mode only: 0 = rows inside the combined mask, 1 = all rows, 2 = rows outside the mask
    int mask_mode;
    // 1 = the task masks are BIT-packed (element 8k+i = bit i of byte k; what svdq_host_pack_mask writes), so that
    // masks staged from the host cross PCIe at 1/8 of the bytes; 0 = torch.bool storage, one byte per element
    int mask_bits;
};

struct K2ReduceArgs {
    const float* gram;            // [n_tiles][FULL ? 2 : 1][G]
    const uint32_t* count;        // [n_tiles]
    const int64_t* tile_begin;    // [P+1] first tile of each parameter
    double* gram_masked;          // [P][NT*NT] full symmetric
    double* gram_all;             // [P][NT*NT] Gram over all elements, masked or not (FULL only; may be null)
    int64_t* dm;                  // [P]
    int nt;
    int full;                     // 0 masked only, 1 second block = all elements, 2 second block = complement
    // full == 2 (noise region, basis.py:455-466): complement Gram and its row count.  A parameter has a noise
    // region iff it has a mask, the masked region passes the svd_min_mask_size gate (cli.py:332) and at least
    // one element lies outside the mask.
    double* gram_noise;           // [P][NT*NT] or null
    int64_t* dm_noise;            // [P] or null
    const int64_t* numel;         // [P]   (full == 2)
    const uint8_t* has_mask;      // [P]   (full == 2)
    int min_mask_size;
};

struct K2SolveArgs {
    SolveConfig cfg;
    const double* gram_masked;    // [P][NT*NT]
    const int64_t* dm;            // [P]
    const uint8_t* has_mask;      // [P]
    const uint32_t* present;      // [P]
    const double* weights;        // [NT]
    const int32_t* avg_order;     // [NT]
    const int32_t* cluster_of;    // [NT] cluster index per task position, or null (svdq_param_average only)
    const double* omega;          // [NT] cross-cluster weights by cluster index (with cluster_of)
    const double* sign_ref;       // [P][NT*NT] or null
    // outputs, strides per parameter as in SolveOut
    int32_t* info; float* sv; float* scal; float* coef; uint16_t* chigh; uint8_t* codes;
    float* qscale; float* qzp; float* qres; float* chat; float* cbar; float* W; float* gvec; double* V;
};

struct K3Args {
    const void* const* tensors;   // [P][NT+1]
    const int64_t* numel;         // [P]
    const int32_t* tile_param;    // [n_tiles]
    const int32_t* tile_local;    // [n_tiles]
    const int64_t* pmask_off;     // [P]
    const uint8_t* has_mask;      // [P]
    const uint32_t* packed;       // packed combined masks written by K1
    const int32_t* info;          // [P][8] from K2
    const float* W;               // [P][NT*NT]  W[t][j]
    const float* cbar;            // [P][NT]
    const float* gvec;            // [P][NT]
    const float* scal;            // [P][4]
    const float* chat;            // [P][NT*NT]  chat[t][j] (DIAG)
    float* const* out;            // [P] merged fp32 tensors
    float* diag;                  // [n_tiles][4][NT] partials (DIAG): sum e^2, sum |e|, sum rec^2, max |e|
    int tile_elems;
    int center;
    // noise region (svd_include_noise): second solve over the rows outside the mask; all null = off
    const int32_t* info_n;        // [P][8]
    const float* W_n;             // [P][NT*NT]
    const float* cbar_n;          // [P][NT]
    const float* gvec_n;          // [P][NT]
    const float* scal_n;          // [P][4]
    float noise_shrink;
    // fused diagnostics, up to 8 tasks (k3c_merge_diag_compact): 0 = compact every tile; 2 = per tile, compaction only
    // where the combined mask keeps fewer than 55 % of the elements
    int diag_select;
    // optional (k3c_merge_diag_compact only): write the artifact bases in the same pass -- U_high [Dm x k], U_low
    // [Dm x (r-k)] (fp16 when the bases are stored in fp16, else fp32), mean [Dm], rows compacted to the masked
    // elements (the layout of svdq_write_basis); tile_row_off from svdq_basis_offsets.  All null = off.
    const int64_t* tile_row_off;
    void* const* u_high;
    void* const* u_low;
    float* const* mean_out;
};

struct K3DiagArgs {
    const float* diag;            // [n_tiles][4][NT]
    const double* gram_masked;    // [P][NT*NT]: its diagonal = sum orig^2 over the masked rows
    const int64_t* tile_begin;    // [P+1]
    const int64_t* dm;            // [P]
    const int32_t* info;          // [P][8]
    double* out;                  // [P][NT][6]: absolute_error, relative_error, max_absolute_error,
                                  //             mean_absolute_error, original_norm, reconstructed_norm
    int nt;
};

struct K6MaskArgs {              // mask combination + packing for the wide (17..32 tasks) path
    const uint8_t* const* masks;    // [P][n_tasks] (null entries allowed) or null
    const int64_t* numel;
    const int32_t* tile_param;
    const int32_t* tile_local;
    const int64_t* pmask_off;
    uint32_t* packed;
    uint32_t* count;                // [n_tiles]
    int tile_elems, strategy, n_tasks;
};

struct K7Args {                  // exact projection on the stored (fp16) basis for selected parameters
    const void* const* tensors;
    const int64_t* numel;
    const int32_t* tile_param;      // [n_sel_tiles] tiles of the selected parameters only
    const int32_t* tile_local;
    const int64_t* pmask_off;
    const uint8_t* has_mask;
    const uint32_t* packed;
    const int32_t* info;            // [P][8]
    const float* W;                 // [P][NT*NT]
    float* proj;                    // [n_sel_tiles][NT*NT] partial coefficients c[t][j]
    int tile_elems, center, fp16_basis;
    int invert;                     // 1 = project the rows OUTSIDE the combined mask (noise region)
};

struct K2RequantArgs {
    SolveConfig cfg;
    const int64_t* sel_tile_begin;  // [P+1] first selected tile of each parameter (empty range: not selected)
    const float* proj;              // [n_sel_tiles][NT*NT]
    const uint32_t* present;        // [P]
    const int32_t* info;            // [P][8]
    float* coef; uint16_t* chigh; uint8_t* codes; float* qscale; float* qzp; float* qres; float* chat;
};

constexpr int kK4MaxGrid = 148 * 8;

struct K4Stats {            // one record per CTA, reduced in CTA order by k4_finalize
    float lo, hi;
    int nan;
    int pad;
    double sumsq;
};

struct K5Args {
    const void* const* tensors;   // [P][NT+1]
    const int64_t* numel;
    const int32_t* tile_param;
    const int32_t* tile_local;
    const int64_t* pmask_off;
    const uint8_t* has_mask;
    const uint32_t* packed;
    const int32_t* info;          // [P][8]
    const float* W;               // [P][NT*NT]
    const int64_t* tile_row_off;  // [n_tiles]
    void* const* u_high;          // [P]  [Dm x k]      (fp16 when fp16_basis else fp32)
    void* const* u_low;           // [P]  [Dm x (r-k)]
    float* const* mean;           // [P]  [Dm] or null entries / null table
    int tile_elems;
    int center;
    int fp16_basis;
    int invert;                   // 1 = rows outside the combined mask (noise basis, basis.py:455-466)
    int n_tasks;                  // stride of the per-task tables (<= the kernel's compile-time task bound)
};

struct K11Args {                 // merge from stored artifacts (reload path)
    const int64_t* numel;           // [P]
    const int32_t* tile_param;      // [n_tiles]
    const int32_t* tile_local;
    const int64_t* pmask_off;       // [P] word offset of the parameter's packed combined mask
    const uint8_t* has_mask;        // [P]
    const uint32_t* packed;
    uint32_t* count;                // [n_tiles] (k11_tile_counts output)
    const int32_t* kr;              // [P][2]: k, r (0, 0: the parameter has no basis -> zeros)
    const void* const* u_high;      // [P] [Dm x k]     fp16 or fp32, rows compacted to the region's elements
    const void* const* u_low;       // [P] [Dm x (r-k)]
    const float* const* mean;       // [P] [Dm] or null entries / null table
    const float* cbar;              // [P][n_tasks] averaged coefficients (c_high then c_low)
    const int64_t* tile_row_off;    // [n_tiles]
    float* const* out;              // [P] fp32 deltas
    int tile_elems, n_tasks, region;
    float scale;                    // svd_noise_shrink for region 1, 1 otherwise
};

// launchers (one translation unit per kernel family; K1/K3/K5 additionally one per dtype)
template <int DT> cudaError_t k1_launch_dtype(int nt, const K1Args& a, int n_tiles, bool full, cudaStream_t st);
template <int DT> cudaError_t k3_launch_dtype(int nt, const K3Args& a, int n_tiles, bool fp16b, bool diag, cudaStream_t st);
template <int DT> cudaError_t k5_launch_dtype(int nt, const K5Args& a, int n_tiles, cudaStream_t st);
// staged persistent variants (TMA bulk copies into a shared-memory ring); cudaErrorNotSupported when nt > 8
template <int DT> cudaError_t k1s_launch_dtype(int nt, const K1Args& a, int n_tiles, bool full, int n_sm, cudaStream_t st);
template <int DT> cudaError_t k3s_launch_dtype(int nt, const K3Args& a, int n_tiles, bool fp16b, int n_sm, cudaStream_t st);
// tensor-core pass 1 (tcgen05) for 16-bit inputs, nt <= 8, single Gram block; cudaErrorNotSupported otherwise
template <int DT> cudaError_t k9_launch_dtype(int nt, const K1Args& a, int n_tiles, int n_sm, cudaStream_t st);
// tensor-core pass 2 (tcgen05) for bf16 inputs, nt <= 8, no diagnostics / noise; cudaErrorNotSupported otherwise
template <int DT> cudaError_t k10_launch_dtype(int nt, const K3Args& a, int n_tiles, bool fp16b, int n_sm, cudaStream_t st);
// tensor-core single-pass Gram (tcgen05, 3-piece bf16 split) for fp32 inputs under a pre-combined mask, 1..32 tasks;
// cudaErrorNotSupported otherwise.  chain = MMA steps per short accumulator chain (2, 4 or 8)
template <int DT> cudaError_t k12_launch_dtype(int n_tasks, const K1Args& a, int n_tiles, int n_sm, int chain, cudaStream_t st);
// tensor-core pass 2 of the wide path (tcgen05) for fp32 inputs, 2..21 tasks, no diagnostics / noise region;
// cudaErrorNotSupported otherwise
template <int DT> cudaError_t k13_launch_dtype(int n_tasks, const K3Args& a, int n_tiles, bool fp16b, int n_sm, cudaStream_t st);
// compacting pass 2 with fused diagnostics (only the elements inside the combined mask go through the arithmetic):
// up to 8 tasks, no noise region; cudaErrorNotSupported otherwise
template <int DT> cudaError_t k3c_launch_dtype(int nt, const K3Args& a, int n_tiles, bool fp16b, cudaStream_t st);
cudaError_t k11_counts_launch(const K11Args& a, int n_tiles, cudaStream_t st);
cudaError_t k11_merge_launch(const K11Args& a, int n_tiles, bool basis_fp16, cudaStream_t st);
cudaError_t k6_mask_pack_launch(const K6MaskArgs& a, int n_tiles, cudaStream_t st);
// single-pass staged Gram for 9..32 tasks under a pre-combined mask (K1Args in pre-combined mode)
template <int DT> cudaError_t k8_launch_dtype(int n_tasks, const K1Args& a, int n_tiles, int n_sm, cudaStream_t st);
template <int DT> cudaError_t k6_merge_launch_dtype(int n_tasks, const K3Args& a, int n_tiles, bool fp16b, bool diag,
                                                    cudaStream_t st);
cudaError_t k2_reduce_launch(const K2ReduceArgs& a, int n_params, cudaStream_t st);
cudaError_t k2_solve_launch(const K2SolveArgs& a, int n_params, cudaStream_t st);
cudaError_t k2_average_launch(const K2SolveArgs& a, int n_params, cudaStream_t st);
cudaError_t k2_requant_launch(const K2RequantArgs& a, int n_params, cudaStream_t st);
template <int DT> cudaError_t k7_launch_dtype(int nt, const K7Args& a, int n_tiles, cudaStream_t st);
cudaError_t k3_diag_launch(const K3DiagArgs& a, int n_params, cudaStream_t st);
cudaError_t k5_offsets_launch(const uint32_t* count, const int64_t* tile_begin, int64_t* tile_row_off, int n_params,
                              const int64_t* numel_if_inverted, int tile_elems, cudaStream_t st);
cudaError_t k4_rtvq_launch(const float* x, int64_t n, int bits, int stages, void* codes, int64_t codes_ld,
                           int code_bytes, float* scale, float* zp, float* resnorm, K4Stats* part, uint32_t* packed,
                           int64_t packed_ld, cudaStream_t st);
cudaError_t k4_dequant_launch(const void* codes, int64_t codes_ld, int code_bytes, int stages, int64_t n,
                              const float* scale, const float* zp, float* out, cudaStream_t st);
cudaError_t k4_absmax_launch(const float* x, int64_t n, int bits, void* q, int code_bytes, float* scale,
                             K4Stats* part, cudaStream_t st);
cudaError_t combine_masks_launch(const uint8_t* const* masks, int n_masks, int64_t n, int strategy, uint8_t* out,
                                 cudaStream_t st);
cudaError_t unpack_mask_launch(const uint32_t* packed, int64_t n, uint8_t* out, cudaStream_t st);
// K14: the operator-by-operator API (k14_operators.cu)
size_t k14_project_scratch_bytes();
int k14_select_chunk();
cudaError_t k14_project_launch(bool fp16, const void* U, int64_t ld, int cols, int64_t rows, const float* delta,
                               const float* mean, float* c, void* scratch, cudaStream_t st);
cudaError_t k14_expand_launch(bool fp16, const void* Uh, int64_t ldh, int k, const void* Ul, int64_t ldl, int nl,
                              int64_t rows, const float* ch, const float* cl, const float* mean, float scale,
                              float* out, cudaStream_t st);
cudaError_t k14_mask_offsets_launch(const uint8_t* mask, int64_t n, int invert, int64_t* chunk_off, cudaStream_t st);
cudaError_t k14_select_launch(bool scatter, int elem_bytes, const void* src, const uint8_t* mask, int64_t n, int invert,
                              const int64_t* chunk_off, void* dst, cudaStream_t st);

template <int DT> struct DTypeOf;
template <> struct DTypeOf<kF32> { using type = float; };
template <> struct DTypeOf<kBF16> { using type = __nv_bfloat16; };
template <> struct DTypeOf<kF16> { using type = __half; };

}  // namespace svdq
