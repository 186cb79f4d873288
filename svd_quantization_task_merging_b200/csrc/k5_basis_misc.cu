// K5 — optional materialisation of the SVD bases in the reference's artifact layout, plus the
// small mask utilities behind the fine-grained API.
//
//   k5_tile_offsets   exclusive scan of K1's per-tile masked counts (row offset of every tile)
//   k5_write_basis    re-reads the inputs and writes U_high [Dm x k], U_low [Dm x (r-k)] (fp16 or
//                     fp32) and mean [Dm x 1] COMPACTED to the masked rows (staged through shared
//                     memory so the stores are 16-byte coalesced), the layout of
//                     construct_basis (src/svd_hybrid/basis.py:363-364,398-407) after
//                     apply_mask_to_tensor (src/svd_hybrid/mask_loader.py:675-679)
//   k_combine_masks   union / intersection / majority over torch.bool tensors
//                     (src/svd_hybrid/mask_loader.py:412-485) -> torch.bool
//   k_unpack_mask     packed combined mask -> torch.bool
#include "svdq_kernels.h"
#include "copy_out.cuh"

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

// One CTA per tile.  Per 1024-element step the CTA scans the mask bits (row index of every kept element),
// rebuilds the basis rows of its kept elements into a shared-memory image of the step's slice of U_high /
// U_low / mean -- the kept rows of a step are consecutive rows of the artifacts -- and then copies the three
// slices out with 16-byte coalesced stores (the smem image starts at the same offset modulo 16 bytes as the
// global slice, so head / tail handling is a few scalar elements).
// NT bounds the unrolled task loops (n = a.n_tasks <= NT tasks are real); VEC elements per thread and step: 4 up to
// 16 tasks, 2 above (the task values of a thread's elements stay in registers).
template <typename T, int NT, typename OUT, int VEC>
__global__ void __launch_bounds__(kBlock) k5_write_basis(const K5Args a) {
    constexpr int kStepV = kBlock * VEC;
    const int n = a.n_tasks;
    constexpr int NTP = (NT + 3) & ~3;
    constexpr int V = 16 / (int)sizeof(OUT);
    extern __shared__ __align__(16) unsigned char dyn[];
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
    if (tid <= NT) s_ptr[tid] = tid <= n ? a.tensors[(int64_t)p * (n + 1) + tid] : nullptr;
    for (int i = tid; i < NT * NTP; i += kBlock) {
        const int j = i / NTP, t = i % NTP;
        sWT[j][t] = (t < n && j < n) ? a.W[(int64_t)p * n * n + t * n + j] : 0.0f;
    }
    __syncthreads();
    uint32_t present_bits = 0;
#pragma unroll
    for (int t = 0; t < NT; ++t) present_bits |= (s_ptr[t + 1] != nullptr ? 1u : 0u) << t;
    __syncthreads();
    if (tid >= 1 && tid <= NT && s_ptr[tid] == nullptr) s_ptr[tid] = s_ptr[0];      // delta == 0, dropped below
    __syncthreads();
    const uint32_t* packed = has_mask ? a.packed + a.pmask_off[p] : nullptr;
    const float n_f = (float)(n_active > 0 ? n_active : 1);
    float* mean_out = a.mean ? a.mean[p] : nullptr;
    OUT* uh_out = reinterpret_cast<OUT*>(a.u_high[p]);
    OUT* ul_out = reinterpret_cast<OUT*>(a.u_low[p]);
    int64_t row_base = a.tile_row_off[tile];
    // smem image: [U_high slice | U_low slice | mean slice], each re-based per step (see k5_copy_out)
    OUT* s_u = reinterpret_cast<OUT*>(dyn);
    float* s_mean = reinterpret_cast<float*>(dyn + ((size_t)(kStepV * NT + 3 * V) * sizeof(OUT) + 15) / 16 * 16);

    for (int64_t e0 = start; e0 < stop; e0 += kStepV) {
        const int64_t e = e0 + (int64_t)tid * VEC;
        uint32_t bits = 0;
        const bool full = e + VEC <= numel;
        if (e < stop) {
            const uint32_t valid = full ? ((1u << VEC) - 1u) : ((1u << (int)(numel - e)) - 1u);
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
        int lr = (int)(warp_off + (incl - mine));                 // first row of this thread inside the step
        const int64_t gh = row_base * k, gl = row_base * nlow, gm = row_base;
        OUT* s_h = s_u + (gh % V);
        OUT* s_l = s_u + ((gh % V) + (int64_t)step_total * k + V - 1) / V * V + (gl % V);
        float* s_m = s_mean + (gm % 4);

        if (bits != 0) {
            float b[VEC], x[NT][VEC], mean[VEC];
            if constexpr (VEC == 4) {
                if (full) {
                    Elem<T>::load4(s_ptr[0], e, b);
#pragma unroll
                    for (int t = 0; t < NT; ++t) Elem<T>::load4(s_ptr[t + 1], e, x[t]);
                }
            }
            if (VEC != 4 || !full) {
#pragma unroll
                for (int c = 0; c < VEC; ++c) b[c] = (e + c < numel) ? Elem<T>::load1(s_ptr[0], e + c) : 0.0f;
#pragma unroll
                for (int t = 0; t < NT; ++t)
#pragma unroll
                    for (int c = 0; c < VEC; ++c)
                        x[t][c] = (e + c < numel) ? Elem<T>::load1(s_ptr[t + 1], e + c) : 0.0f;
            }
#pragma unroll
            for (int c = 0; c < VEC; ++c) mean[c] = 0.0f;
#pragma unroll
            for (int t = 0; t < NT; ++t)
#pragma unroll
                for (int c = 0; c < VEC; ++c) { x[t][c] = Elem<T>::sub(x[t][c], b[c]); mean[c] += x[t][c]; }
#pragma unroll
            for (int c = 0; c < VEC; ++c) mean[c] = a.center ? __fdiv_rn(mean[c], n_f) : 0.0f;
#pragma unroll
            for (int t = 0; t < NT; ++t)
#pragma unroll
                for (int c = 0; c < VEC; ++c) x[t][c] = ((present_bits >> t) & 1u) ? x[t][c] - mean[c] : 0.0f;
#pragma unroll
            for (int c = 0; c < VEC; ++c) {
                if (!((bits >> c) & 1u)) continue;
                if (mean_out) s_m[lr] = mean[c];
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    if (j >= r) break;
                    float u = 0.0f;
#pragma unroll
                    for (int t = 0; t < NT; ++t) u = fmaf(x[t][c], sWT[j][t], u);
                    if (j < k) s_h[lr * k + j] = OutCvt<OUT>::cvt(u);
                    else s_l[lr * nlow + (j - k)] = OutCvt<OUT>::cvt(u);
                }
                ++lr;
            }
        }
        __syncthreads();
        k5_copy_out<OUT>(uh_out, s_h, gh, (int)step_total * k, tid);
        k5_copy_out<OUT>(ul_out, s_l, gl, (int)step_total * nlow, tid);
        if (mean_out) k5_copy_out<float>(mean_out, s_m, gm, (int)step_total, tid);
        row_base += step_total;
        __syncthreads();
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

template <typename T, int NT, typename OUT, int VEC>
static cudaError_t k5_go(const K5Args& a, int n_tiles, cudaStream_t st) {
    constexpr int V = 16 / (int)sizeof(OUT);
    constexpr int kStepV = kBlock * VEC;
    const size_t dsm = ((size_t)(kStepV * NT + 3 * V) * sizeof(OUT) + 15) / 16 * 16 + (kStepV + 4) * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(k5_write_basis<T, NT, OUT, VEC>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm);
    if (e != cudaSuccess) return e;
    k5_write_basis<T, NT, OUT, VEC><<<n_tiles, kBlock, dsm, st>>>(a);
    return cudaGetLastError();
}

template <>
cudaError_t k5_launch_dtype<SVDQ_DTYPE>(int nt, const K5Args& a, int n_tiles, cudaStream_t st) {
    using T = DTypeOf<SVDQ_DTYPE>::type;
    if (n_tiles <= 0) return cudaSuccess;
    if (a.tile_elems % (kBlock * 4) != 0) return cudaErrorInvalidValue;
    switch (nt) {
#define SVDQ_CASE(N) \
    case N: return a.fp16_basis ? k5_go<T, N, __half, 4>(a, n_tiles, st) : k5_go<T, N, float, 4>(a, n_tiles, st);
        SVDQ_CASE(1) SVDQ_CASE(2) SVDQ_CASE(3) SVDQ_CASE(4) SVDQ_CASE(5) SVDQ_CASE(6) SVDQ_CASE(7) SVDQ_CASE(8)
        SVDQ_CASE(9) SVDQ_CASE(10) SVDQ_CASE(11) SVDQ_CASE(12) SVDQ_CASE(13) SVDQ_CASE(14) SVDQ_CASE(15) SVDQ_CASE(16)
#undef SVDQ_CASE
        default: break;
    }
    if (nt < 1 || nt > kMaxTasks) return cudaErrorInvalidValue;
    // 17..32 tasks: runtime task count under a compile-time bound, two elements per thread
    if (nt <= 24) return a.fp16_basis ? k5_go<T, 24, __half, 2>(a, n_tiles, st) : k5_go<T, 24, float, 2>(a, n_tiles, st);
    return a.fp16_basis ? k5_go<T, 32, __half, 2>(a, n_tiles, st) : k5_go<T, 32, float, 2>(a, n_tiles, st);
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
