// K11 — merge from STORED artifacts: the reload path (reference: merge_all_parameters / merge_parameter /
// reconstruct_from_coefficients src/svd_hybrid/merge.py:144-426 as driven by src/svd_hybrid/reload.py:142-238;
// scatter of the masked rows: reconstruct_from_masked src/svd_hybrid/mask_loader.py:712-763).
// The bases are read back in the artifact layout -- U_high [Dm x k], U_low [Dm x (r-k)] (fp16 or fp32), mean [Dm x 1],
// rows COMPACTED to the masked elements -- so this is pass 2 with U read instead of rebuilt: every parameter of the
// model in ONE launch (the reference and round 1 of this build: one matvec pair per parameter).
//   k11_tile_counts : masked elements per tile from the packed combined masks (row offsets via k5_tile_offsets)
//   k11_reload_merge: per 1024-element step a block-wide scan of the mask bits gives every kept element its row;
//                     delta[d] = U_high[row] . c_high + U_low[row] . c_low + mean[row]; region 0 writes the rows inside
//                     the mask and zeros elsewhere, region 1 (noise basis, merge.py:257-284) writes scale * value at
//                     the rows OUTSIDE the mask and leaves the rest as it is.
// Bound: HBM (2 r + 4 bytes read, 4 written per kept element).
#include "svdq_kernels.h"

namespace svdq {

__global__ void __launch_bounds__(kBlock) k11_tile_counts(const K11Args a) {
    __shared__ uint32_t s_w[kBlock / 32];
    const int tile = blockIdx.x, tid = threadIdx.x;
    const int p = a.tile_param[tile];
    const int64_t numel = a.numel[p];
    const int64_t start = (int64_t)a.tile_local[tile] * a.tile_elems;
    const int64_t stop = min(start + (int64_t)a.tile_elems, numel);
    uint32_t c = 0;
    if (a.has_mask[p]) {
        const uint32_t* packed = a.packed + a.pmask_off[p];
        for (int64_t w = (start >> 5) + tid; w * 32 < stop; w += kBlock) {
            uint32_t bits = __ldg(packed + w);
            const int64_t left = numel - w * 32;
            if (left < 32) bits &= (1u << (int)left) - 1u;
            c += __popc(bits);
        }
    } else if (tid == 0) {
        c = (uint32_t)(stop - start);
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if ((tid & 31) == 0) s_w[tid >> 5] = c;
    __syncthreads();
    if (tid == 0) {
        uint32_t t = 0;
        for (int w = 0; w < kBlock / 32; ++w) t += s_w[w];
        a.count[tile] = t;
    }
}

template <typename U>
__device__ __forceinline__ float k11_ld(const U* p, int64_t i);
template <> __device__ __forceinline__ float k11_ld<__half>(const __half* p, int64_t i) { return __half2float(p[i]); }
template <> __device__ __forceinline__ float k11_ld<float>(const float* p, int64_t i) { return __ldg(p + i); }

template <typename U>
__global__ void __launch_bounds__(kBlock) k11_reload_merge(const K11Args a) {
    __shared__ uint32_t s_warp[kBlock / 32];
    __shared__ float s_c[kMaxTasks];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x;
    const int p = a.tile_param[tile];
    const int r = min(max(a.kr[2 * p + 1], 0), a.n_tasks), k = min(max(a.kr[2 * p], 0), r);
    const int nlow = r - k;
    const int64_t numel = a.numel[p];
    const int64_t start = (int64_t)a.tile_local[tile] * a.tile_elems;
    const int64_t stop = min(start + (int64_t)a.tile_elems, numel);
    const bool has_mask = a.has_mask[p] != 0;
    const bool invert = a.region != 0;
    const U* uh = reinterpret_cast<const U*>(a.u_high[p]);
    const U* ul = reinterpret_cast<const U*>(a.u_low[p]);
    const float* mean = a.mean ? a.mean[p] : nullptr;
    float* out = a.out[p];
    const bool have = r > 0 && uh != nullptr;
    if (tid < kMaxTasks) s_c[tid] = (tid < r && tid < a.n_tasks) ? a.cbar[(int64_t)p * a.n_tasks + tid] : 0.0f;
    __syncthreads();
    const uint32_t* packed = has_mask ? a.packed + a.pmask_off[p] : nullptr;
    int64_t row_base = a.tile_row_off[tile];
    for (int64_t e0 = start; e0 < stop; e0 += kStep) {
        const int64_t e = e0 + (int64_t)tid * kVec;
        uint32_t bits = 0, valid = 0;
        if (e < stop) {
            valid = (e + kVec <= numel) ? 0xFu : ((1u << (int)(numel - e)) - 1u);
            if (has_mask) {
                const uint32_t w = __ldg(packed + (e >> 5));
                bits = ((invert ? ~w : w) >> (int)(e & 31)) & valid;
            } else bits = invert ? 0u : valid;
        }
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
        int64_t row = row_base + warp_off + (incl - mine);
#pragma unroll
        for (int c = 0; c < kVec; ++c) {
            if (!((valid >> c) & 1u)) continue;
            const bool on = (bits >> c) & 1u;
            float val = 0.0f;
            if (on && have) {
                float acc = 0.0f;
                for (int j = 0; j < k; ++j) acc = fmaf(k11_ld<U>(uh, row * k + j), s_c[j], acc);
                float acc2 = 0.0f;
                for (int j = 0; j < nlow; ++j) acc2 = fmaf(k11_ld<U>(ul, row * nlow + j), s_c[k + j], acc2);
                val = acc + acc2;                           // U_high c_high + U_low c_low (merge.py:184-186)
                if (mean) val += __ldg(mean + row);
                val *= a.scale;
            }
            if (on) ++row;
            if (!invert) out[e + c] = val;                  // region 0 owns the whole tensor: zeros outside the mask
            else if (on) out[e + c] = val;                  // region 1 fills the unmasked rows only
        }
        row_base += step_total;
        __syncthreads();
    }
}

cudaError_t k11_counts_launch(const K11Args& a, int n_tiles, cudaStream_t st) {
    if (n_tiles <= 0) return cudaSuccess;
    k11_tile_counts<<<n_tiles, kBlock, 0, st>>>(a);
    return cudaGetLastError();
}

cudaError_t k11_merge_launch(const K11Args& a, int n_tiles, bool basis_fp16, cudaStream_t st) {
    if (n_tiles <= 0) return cudaSuccess;
    if (a.n_tasks < 1 || a.n_tasks > kMaxTasks) return cudaErrorInvalidValue;
    if (basis_fp16) k11_reload_merge<__half><<<n_tiles, kBlock, 0, st>>>(a);
    else k11_reload_merge<float><<<n_tiles, kBlock, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace svdq
