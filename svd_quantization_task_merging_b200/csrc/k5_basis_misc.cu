// K5 — optional materialisation of the SVD bases in the reference's artifact layout, plus the
// small mask utilities behind the fine-grained API.
//
//   k5_tile_offsets   exclusive scan of K1's per-tile masked counts (row offset of every tile)
//   k5_write_basis    re-reads the inputs and writes U_high [Dm x k], U_low [Dm x (r-k)] (fp16 or
//                     fp32) and mean [Dm x 1] COMPACTED to the masked rows, the layout of
//                     construct_basis (src/svd_hybrid/basis.py:363-364,398-407) after
//                     apply_mask_to_tensor (src/svd_hybrid/mask_loader.py:675-679)
//   k_combine_masks   union / intersection / majority over torch.bool tensors
//                     (src/svd_hybrid/mask_loader.py:412-485) -> torch.bool
//   k_unpack_mask     packed combined mask -> torch.bool
#include "svdq_kernels.h"

#ifndef SVDQ_DTYPE
#define SVDQ_DTYPE 0
#endif

namespace svdq {

#if SVDQ_DTYPE == 0
__global__ void __launch_bounds__(32) k5_tile_offsets(const uint32_t* count, const int64_t* tile_begin,
                                                      int64_t* tile_row_off, const int64_t* numel_inv,
                                                      int tile_elems) {
    if (threadIdx.x != 0) return;
    const int p = blockIdx.x;
    int64_t acc = 0;
    for (int64_t t = tile_begin[p]; t < tile_begin[p + 1]; ++t) {
        tile_row_off[t] = acc;
        if (numel_inv) {       // rows OUTSIDE the mask: elements of the tile minus the masked count
            const int64_t lo = (t - tile_begin[p]) * (int64_t)tile_elems;
            const int64_t in_tile = min((int64_t)tile_elems, numel_inv[p] - lo);
            acc += in_tile - (int64_t)count[t];
        } else acc += count[t];
    }
}


#endif  // SVDQ_DTYPE == 0

template <typename T, int NT>
__global__ void __launch_bounds__(kBlock) k5_write_basis(const K5Args a) {
    constexpr int NTP = (NT + 3) & ~3;
    __shared__ __align__(16) float sWT[NT][NTP];
    __shared__ const void* s_ptr[NT + 1];
    __shared__ uint32_t s_warp[kBlock / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x;
    const int p = a.tile_param[tile];
    if (a.info[(int64_t)p * 8] != kSolved) return;
    const int n_active = a.info[(int64_t)p * 8 + 1], r = a.info[(int64_t)p * 8 + 2], k = a.info[(int64_t)p * 8 + 3];
    const int nlow = r - k;
    const int64_t numel = a.numel[p];
    const int64_t start = (int64_t)a.tile_local[tile] * a.tile_elems;
    const int64_t stop = min(start + (int64_t)a.tile_elems, numel);
    const bool has_mask = a.has_mask[p] != 0;
    if (tid <= NT) s_ptr[tid] = a.tensors[(int64_t)p * (NT + 1) + tid];
    for (int i = tid; i < NT * NTP; i += kBlock) {
        const int j = i / NTP, t = i % NTP;
        sWT[j][t] = (t < NT) ? a.W[(int64_t)p * NT * NT + t * NT + j] : 0.0f;
    }
    __syncthreads();
    const uint32_t* packed = has_mask ? a.packed + a.pmask_off[p] : nullptr;
    const float n_f = (float)(n_active > 0 ? n_active : 1);
    float* mean_out = a.mean ? a.mean[p] : nullptr;
    int64_t row_base = a.tile_row_off[tile];

    for (int64_t e0 = start; e0 < stop; e0 += kStep) {
        const int64_t e = e0 + (int64_t)tid * kVec;
        uint32_t bits = 0;
        if (e < stop) {
            const uint32_t valid = (e + kVec <= numel) ? 0xFu : ((1u << (int)(numel - e)) - 1u);
            if (has_mask) {
                const uint32_t w = __ldg(packed + (e >> 5));
                bits = ((a.invert ? ~w : w) >> (int)(e & 31)) & valid;
            } else bits = a.invert ? 0u : valid;
        }
        // block-wide exclusive scan of popc(bits) in element order
        const uint32_t mine = __popc(bits);
        uint32_t incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        uint32_t warp_off = 0, step_total = 0;
#pragma unroll
        for (int w = 0; w < kBlock / 32; ++w) {
            const uint32_t v = s_warp[w];
            if (w < warp) warp_off += v;
            step_total += v;
        }
        __syncthreads();
        int64_t row = row_base + warp_off + (incl - mine);
        row_base += step_total;
        if (bits == 0) continue;

        float b[kVec], x[NT][kVec], mean[kVec];
#pragma unroll
        for (int c = 0; c < kVec; ++c) {
            b[c] = (e + c < numel) ? Elem<T>::load1(s_ptr[0], e + c) : 0.0f;
            mean[c] = 0.0f;
        }
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            const void* fp = s_ptr[t + 1];
#pragma unroll
            for (int c = 0; c < kVec; ++c) {
                x[t][c] = (fp != nullptr && e + c < numel) ? Elem<T>::sub(Elem<T>::load1(fp, e + c), b[c]) : 0.0f;
                mean[c] += x[t][c];
            }
        }
#pragma unroll
        for (int c = 0; c < kVec; ++c) mean[c] = a.center ? __fdiv_rn(mean[c], n_f) : 0.0f;
#pragma unroll
        for (int c = 0; c < kVec; ++c) {
            if (!((bits >> c) & 1u)) continue;
            if (mean_out) mean_out[row] = mean[c];
            for (int j = 0; j < r; ++j) {
                float u = 0.0f;
#pragma unroll
                for (int t = 0; t < NT; ++t)
                    u = fmaf((s_ptr[t + 1] != nullptr) ? x[t][c] - mean[c] : 0.0f, sWT[j][t], u);
                void* dst = j < k ? a.u_high[p] : a.u_low[p];
                const int64_t off = j < k ? row * k + j : row * nlow + (j - k);
                if (a.fp16_basis) reinterpret_cast<__half*>(dst)[off] = __float2half_rn(u);
                else reinterpret_cast<float*>(dst)[off] = u;
            }
            ++row;
        }
    }
}

#if SVDQ_DTYPE == 0
// ---- mask utilities -------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_combine_masks(const uint8_t* const* masks, int n_masks, int64_t n,
                                                          int strategy, uint8_t* out) {
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock) {
        int votes = 0;
        for (int t = 0; t < n_masks; ++t) votes += __ldg(masks[t] + i) != 0;
        bool v;
        if (strategy == kUnion) v = votes > 0;
        else if (strategy == kIntersection) v = votes == n_masks;
        else v = 2 * votes >= n_masks;
        out[i] = v ? 1 : 0;
    }
}

__global__ void __launch_bounds__(kBlock) k_unpack_mask(const uint32_t* packed, int64_t n, uint8_t* out) {
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock)
        out[i] = (__ldg(packed + (i >> 5)) >> (int)(i & 31)) & 1u;
}

#endif  // SVDQ_DTYPE == 0

template <>
cudaError_t k5_launch_dtype<SVDQ_DTYPE>(int nt, const K5Args& a, int n_tiles, cudaStream_t st) {
    using T = DTypeOf<SVDQ_DTYPE>::type;
    if (n_tiles <= 0) return cudaSuccess;
    switch (nt) {
#define SVDQ_CASE(N) case N: k5_write_basis<T, N><<<n_tiles, kBlock, 0, st>>>(a); break;
        SVDQ_CASE(1) SVDQ_CASE(2) SVDQ_CASE(3) SVDQ_CASE(4) SVDQ_CASE(5) SVDQ_CASE(6) SVDQ_CASE(7) SVDQ_CASE(8)
        SVDQ_CASE(9) SVDQ_CASE(10) SVDQ_CASE(11) SVDQ_CASE(12) SVDQ_CASE(13) SVDQ_CASE(14) SVDQ_CASE(15) SVDQ_CASE(16)
#undef SVDQ_CASE
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

#if SVDQ_DTYPE == 0
cudaError_t k5_offsets_launch(const uint32_t* count, const int64_t* tile_begin, int64_t* tile_row_off, int n_params,
                              const int64_t* numel_if_inverted, int tile_elems, cudaStream_t st) {
    if (n_params <= 0) return cudaSuccess;
    k5_tile_offsets<<<n_params, 32, 0, st>>>(count, tile_begin, tile_row_off, numel_if_inverted, tile_elems);
    return cudaGetLastError();
}

cudaError_t combine_masks_launch(const uint8_t* const* masks, int n_masks, int64_t n, int strategy, uint8_t* out,
                                 cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    if (n_masks < 1 || n_masks > kMaxTasks || strategy < 0 || strategy > 2) return cudaErrorInvalidValue;
    int64_t g = (n + kBlock - 1) / kBlock;
    if (g > 148 * 16) g = 148 * 16;
    k_combine_masks<<<(int)g, kBlock, 0, st>>>(masks, n_masks, n, strategy, out);
    return cudaGetLastError();
}

cudaError_t unpack_mask_launch(const uint32_t* packed, int64_t n, uint8_t* out, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    int64_t g = (n + kBlock - 1) / kBlock;
    if (g > 148 * 16) g = 148 * 16;
    k_unpack_mask<<<(int)g, kBlock, 0, st>>>(packed, n, out);
    return cudaGetLastError();
}
#endif

}  // namespace svdq
