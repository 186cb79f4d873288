// K7 — exact projection of the task vectors on the STORED basis, for small parameters.
//
// The reference casts the basis to fp16 (src/svd_hybrid/cli.py:355-361) and only then projects,
// c_t = fp16(U)^T (tau_t - mean) (src/svd_hybrid/compress.py:13-19, 35-37).  The closed form Sigma V^T used
// by K2 differs from that by the fp16 rounding noise of U averaged over the Dm rows (~2.4e-4 / sqrt(Dm)
// relative): negligible for large tensors, but enough to move an fp16 coefficient / RTVQ code by one step
// on the many small ones (biases, LayerNorm weights).  For parameters below a size threshold this kernel
// recomputes the coefficients the reference's way -- rebuild each basis row u_d = (tau_d - mean_d) W, round it
// to fp16, accumulate u_d (x) (tau_d - mean_d) over the masked rows -- and k2_param_requantize re-runs the
// fp16 / RTVQ step on them.  Costs one extra read of those parameters only (~0.2 % of the bytes of ViT-L-14).
// k7_project_exact<T, NT> serves N <= 8 (all N x N accumulators in registers); k7_project_exact_wide covers
// 9..32 task vectors by splitting the basis columns over blockIdx.y.
#include "svdq_kernels.h"

#ifndef SVDQ_DTYPE
#define SVDQ_DTYPE 0
#endif

namespace svdq {

template <typename T, int NT>
__global__ void __launch_bounds__(kBlock, 1) k7_project_exact(const K7Args a) {
    constexpr int NTP = (NT + 3) & ~3;
    constexpr int NACC = NT * NT;
    constexpr int kRows = 32;
    __shared__ __align__(16) float sWT[NT][NTP];
    __shared__ float red[kRows][kBlock + 1];
    __shared__ const void* s_ptr[NT + 1];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x;
    const int p = a.tile_param[tile];
    const int64_t numel = a.numel[p];
    const int64_t start = (int64_t)a.tile_local[tile] * a.tile_elems;
    const int64_t stop = min(start + (int64_t)a.tile_elems, numel);
    const bool solved = a.info[(int64_t)p * 8] == kSolved;
    const int n_active = a.info[(int64_t)p * 8 + 1];
    const int r = a.info[(int64_t)p * 8 + 4];                      // r_eff: null columns have a zero basis column
    const bool has_mask = a.has_mask[p] != 0;
    if (tid <= NT) s_ptr[tid] = a.tensors[(int64_t)p * (NT + 1) + tid];
    for (int i = tid; i < NT * NTP; i += kBlock) {
        const int j = i / NTP, t = i % NTP;
        sWT[j][t] = (t < NT) ? a.W[(int64_t)p * NT * NT + t * NT + j] : 0.0f;
    }
    __syncthreads();
    uint32_t present_bits = 0;
#pragma unroll
    for (int t = 0; t < NT; ++t) present_bits |= (s_ptr[t + 1] != nullptr ? 1u : 0u) << t;
    const uint32_t* packed = has_mask ? a.packed + a.pmask_off[p] : nullptr;
    const float n_f = (float)(n_active > 0 ? n_active : 1);

    float acc[NACC];                                               // acc[t * NT + j] = sum_d x_dt * u_dj
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = 0.0f;

    if (solved) {
        for (int64_t e0 = start; e0 < stop; e0 += kStep) {
            const int64_t e = e0 + (int64_t)tid * kVec;
            if (e >= stop) continue;
            const bool full = e + kVec <= numel;
            float b[kVec], x[NT][kVec], mean[kVec];
            if (full) Elem<T>::load4(s_ptr[0], e, b);
            else {
#pragma unroll
                for (int c = 0; c < kVec; ++c) b[c] = (e + c < numel) ? Elem<T>::load1(s_ptr[0], e + c) : 0.0f;
            }
#pragma unroll
            for (int c = 0; c < kVec; ++c) mean[c] = 0.0f;
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                const void* fp = s_ptr[t + 1];
                if (fp == nullptr) {
#pragma unroll
                    for (int c = 0; c < kVec; ++c) x[t][c] = 0.0f;
                    continue;
                }
                float f[kVec];
                if (full) Elem<T>::load4(fp, e, f);
                else {
#pragma unroll
                    for (int c = 0; c < kVec; ++c) f[c] = (e + c < numel) ? Elem<T>::load1(fp, e + c) : b[c];
                }
#pragma unroll
                for (int c = 0; c < kVec; ++c) { x[t][c] = Elem<T>::sub(f[c], b[c]); mean[c] += x[t][c]; }
            }
            uint32_t bits = full ? 0xFu : ((1u << (int)(numel - e)) - 1u);
            if (has_mask) {
                const uint32_t w = __ldg(packed + (e >> 5));
                bits &= ((a.invert ? ~w : w) >> (int)(e & 31)) & 0xFu;
            } else if (a.invert) bits = 0;
#pragma unroll
            for (int c = 0; c < kVec; ++c) mean[c] = a.center ? __fdiv_rn(mean[c], n_f) : 0.0f;
#pragma unroll
            for (int t = 0; t < NT; ++t)
#pragma unroll
                for (int c = 0; c < kVec; ++c) {
                    const bool on = ((present_bits >> t) & 1u) && ((bits >> c) & 1u);
                    x[t][c] = on ? x[t][c] - mean[c] : 0.0f;       // rows outside the mask contribute nothing
                }
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                if (j >= r) break;
                float u[kVec];
#pragma unroll
                for (int c = 0; c < kVec; ++c) u[c] = 0.0f;
#pragma unroll
                for (int t = 0; t < NT; ++t) {
                    const float w = sWT[j][t];
#pragma unroll
                    for (int c = 0; c < kVec; ++c) u[c] = fmaf(x[t][c], w, u[c]);
                }
                if (a.fp16_basis) {
#pragma unroll
                    for (int c = 0; c < kVec; ++c) u[c] = round_fp16(u[c]);
                }
#pragma unroll
                for (int t = 0; t < NT; ++t)
#pragma unroll
                    for (int c = 0; c < kVec; ++c) acc[t * NT + j] = fmaf(x[t][c], u[c], acc[t * NT + j]);
            }
        }
    }
    float* gout = a.proj + (int64_t)tile * NACC;
#pragma unroll
    for (int r0 = 0; r0 < NACC; r0 += kRows) {
#pragma unroll
        for (int q = 0; q < kRows; ++q)
            if (r0 + q < NACC) red[q][tid] = acc[r0 + q];
        __syncthreads();
#pragma unroll
        for (int rr = 0; rr < kRows / (kBlock / 32); ++rr) {
            const int q = warp * (kRows / (kBlock / 32)) + rr;
            if (r0 + q < NACC) {
                float s = 0.0f;
#pragma unroll
                for (int c = 0; c < kBlock / 32; ++c) s += red[q][lane + 32 * c];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                if (lane == 0) gout[r0 + q] = s;
            }
        }
        __syncthreads();
    }
}

// ---- 9..32 task vectors: runtime N (<= NTMAX), the N x r coefficient block is split over blockIdx.y in groups of
// JB basis columns so that the N x JB accumulators of a thread stay in registers; every column group re-reads
// the (small) selected parameters.  VEC elements per thread and step.
template <typename T, int VEC> struct K7Load;
template <typename T> struct K7Load<T, 4> {
    static __device__ __forceinline__ void load(const void* p, int64_t e, bool full, int64_t numel, float (&o)[4]) {
        if (full) Elem<T>::load4(p, e, o);
        else {
#pragma unroll
            for (int c = 0; c < 4; ++c) o[c] = (e + c < numel) ? Elem<T>::load1(p, e + c) : 0.0f;
        }
    }
};
template <typename T> struct K7Load<T, 2> {
    static __device__ __forceinline__ void load(const void* p, int64_t e, bool, int64_t numel, float (&o)[2]) {
#pragma unroll
        for (int c = 0; c < 2; ++c) o[c] = (e + c < numel) ? Elem<T>::load1(p, e + c) : 0.0f;
    }
};

template <typename T, int NTMAX, int JB, int VEC>
__global__ void __launch_bounds__(kBlock, 1) k7_project_exact_wide(const K7Args a, const int n_tasks) {
    constexpr int NACC = NTMAX * JB;
    constexpr int kRows = 32;
    constexpr int kStepW = kBlock * VEC;
    __shared__ float sW[JB][NTMAX];                                // sW[jj][t] = W[t][j0 + jj]
    __shared__ float red[kRows][kBlock + 1];
    __shared__ const void* s_ptr[NTMAX + 1];
    const int N = n_tasks;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x;
    const int j0 = blockIdx.y * JB;
    const int p = a.tile_param[tile];
    if (a.info[(int64_t)p * 8] != kSolved) return;
    const int r = a.info[(int64_t)p * 8 + 4];                      // r_eff
    if (j0 >= r) return;                                           // k2_param_requantize reads columns < r_eff only
    const int64_t numel = a.numel[p];
    const int64_t start = (int64_t)a.tile_local[tile] * a.tile_elems;
    const int64_t stop = min(start + (int64_t)a.tile_elems, numel);
    const int n_active = a.info[(int64_t)p * 8 + 1];
    const bool has_mask = a.has_mask[p] != 0;
    if (tid <= N) s_ptr[tid] = a.tensors[(int64_t)p * (N + 1) + tid];
    for (int i = tid; i < JB * NTMAX; i += kBlock) {
        const int jj = i / NTMAX, t = i % NTMAX;
        sW[jj][t] = (t < N && j0 + jj < r) ? a.W[(int64_t)p * N * N + t * N + j0 + jj] : 0.0f;
    }
    __syncthreads();
    const uint32_t* packed = has_mask ? a.packed + a.pmask_off[p] : nullptr;
    const float n_f = (float)(n_active > 0 ? n_active : 1);

    float acc[NACC];                                               // acc[t * JB + jj]
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = 0.0f;

    for (int64_t e0 = start; e0 < stop; e0 += kStepW) {
        const int64_t e = e0 + (int64_t)tid * VEC;
        if (e >= stop) continue;
        const bool full = e + VEC <= numel;
        float b[VEC], x[NTMAX][VEC], mean[VEC];
        K7Load<T, VEC>::load(s_ptr[0], e, full, numel, b);
#pragma unroll
        for (int c = 0; c < VEC; ++c) mean[c] = 0.0f;
#pragma unroll
        for (int t = 0; t < NTMAX; ++t) {
            const void* fp = t < N ? s_ptr[t + 1] : nullptr;
            if (fp == nullptr) {
#pragma unroll
                for (int c = 0; c < VEC; ++c) x[t][c] = 0.0f;
                continue;
            }
            float f[VEC];
            K7Load<T, VEC>::load(fp, e, full, numel, f);
#pragma unroll
            for (int c = 0; c < VEC; ++c) {
                x[t][c] = (e + c < numel) ? Elem<T>::sub(f[c], b[c]) : 0.0f;
                mean[c] += x[t][c];
            }
        }
        uint32_t bits = full ? ((1u << VEC) - 1u) : ((1u << (int)(numel - e)) - 1u);
        if (has_mask) {
            const uint32_t w = __ldg(packed + (e >> 5));
            bits &= ((a.invert ? ~w : w) >> (int)(e & 31)) & ((1u << VEC) - 1u);
        } else if (a.invert) bits = 0;
#pragma unroll
        for (int c = 0; c < VEC; ++c) mean[c] = a.center ? __fdiv_rn(mean[c], n_f) : 0.0f;
#pragma unroll
        for (int t = 0; t < NTMAX; ++t) {
            const bool present = t < N && s_ptr[t + 1] != nullptr;
#pragma unroll
            for (int c = 0; c < VEC; ++c) x[t][c] = (present && ((bits >> c) & 1u)) ? x[t][c] - mean[c] : 0.0f;
        }
#pragma unroll
        for (int jj = 0; jj < JB; ++jj) {
            float u[VEC];
#pragma unroll
            for (int c = 0; c < VEC; ++c) u[c] = 0.0f;
#pragma unroll
            for (int t = 0; t < NTMAX; ++t) {
                const float w = sW[jj][t];
#pragma unroll
                for (int c = 0; c < VEC; ++c) u[c] = fmaf(x[t][c], w, u[c]);
            }
            if (a.fp16_basis) {
#pragma unroll
                for (int c = 0; c < VEC; ++c) u[c] = round_fp16(u[c]);
            }
#pragma unroll
            for (int t = 0; t < NTMAX; ++t)
#pragma unroll
                for (int c = 0; c < VEC; ++c) acc[t * JB + jj] = fmaf(x[t][c], u[c], acc[t * JB + jj]);
        }
    }
    float* gout = a.proj + (int64_t)tile * N * N;
#pragma unroll
    for (int r0 = 0; r0 < NACC; r0 += kRows) {
#pragma unroll
        for (int q = 0; q < kRows; ++q)
            if (r0 + q < NACC) red[q][tid] = acc[r0 + q];
        __syncthreads();
#pragma unroll
        for (int rr = 0; rr < kRows / (kBlock / 32); ++rr) {
            const int q = warp * (kRows / (kBlock / 32)) + rr;
            if (r0 + q < NACC) {
                float s = 0.0f;
#pragma unroll
                for (int c = 0; c < kBlock / 32; ++c) s += red[q][lane + 32 * c];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                const int t = (r0 + q) / JB, j = j0 + (r0 + q) % JB;
                if (lane == 0 && t < N && j < N) gout[t * N + j] = s;
            }
        }
        __syncthreads();
    }
}

template <typename T, int NTMAX, int JB, int VEC>
static cudaError_t k7_wide_go(const K7Args& a, int n_tasks, int n_tiles, cudaStream_t st) {
    const dim3 grid(n_tiles, (n_tasks + JB - 1) / JB);
    k7_project_exact_wide<T, NTMAX, JB, VEC><<<grid, kBlock, 0, st>>>(a, n_tasks);
    return cudaGetLastError();
}

template <>
cudaError_t k7_launch_dtype<SVDQ_DTYPE>(int nt, const K7Args& a, int n_tiles, cudaStream_t st) {
    using T = DTypeOf<SVDQ_DTYPE>::type;
    if (n_tiles <= 0) return cudaSuccess;
    switch (nt) {
#define SVDQ_CASE(N) case N: k7_project_exact<T, N><<<n_tiles, kBlock, 0, st>>>(a); break;
        SVDQ_CASE(1) SVDQ_CASE(2) SVDQ_CASE(3) SVDQ_CASE(4) SVDQ_CASE(5) SVDQ_CASE(6) SVDQ_CASE(7) SVDQ_CASE(8)
#undef SVDQ_CASE
        default: break;
    }
    if (nt <= 8) return cudaGetLastError();
    if (nt <= 16) return k7_wide_go<T, 16, 4, 4>(a, nt, n_tiles, st);
    if (nt <= 24) return k7_wide_go<T, 24, 2, 2>(a, nt, n_tiles, st);
    if (nt <= kMaxTasks) return k7_wide_go<T, 32, 2, 2>(a, nt, n_tiles, st);
    return cudaErrorInvalidValue;
}

}  // namespace svdq
