// K6 — "wide" path for 17..32 task vectors (configs[3]: 20 tasks), where the N(N+1)/2 Gram
// accumulators no longer fit one thread's registers:
//   k6_mask_pack            combines the N tall masks once (runtime N) -> packed mask + per-tile counts;
//                           the Gram is then accumulated by k1_tv_mask_gram<., NT <= 16> launches over
//                           task-subset pairs in pre-combined-mask mode, all seeing this same mask
//   k6_reconstruct_merge    pass 2 with a runtime number of tasks: basis-row accumulators for all
//                           r <= 32 columns live in registers (two elements per thread), the tasks are
//                           streamed once; optional fused diagnostics accumulate in shared memory
// Same reference lines as K1 / K3.  Bound: HBM for the mask pass; pass 2 is shared-memory/FMA-bound at
// N = 20 (N*r = 400 FMAs per element) -- see DESIGN.md.
#include "svdq_kernels.h"
#include "k3_body.cuh"

#ifndef SVDQ_DTYPE
#define SVDQ_DTYPE 0
#endif

namespace svdq {

#if SVDQ_DTYPE == 0
__global__ void __launch_bounds__(kBlock) k6_mask_pack(const K6MaskArgs a) {
    __shared__ const uint8_t* s_mask[kMaxTasks];
    __shared__ uint32_t s_cnt[kBlock / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x;
    const int p = a.tile_param[tile];
    const int64_t numel = a.numel[p];
    const int64_t start = (int64_t)a.tile_local[tile] * a.tile_elems;
    const int64_t stop = min(start + (int64_t)a.tile_elems, numel);
    if (tid < a.n_tasks) s_mask[tid] = a.masks ? a.masks[(int64_t)p * a.n_tasks + tid] : nullptr;
    __syncthreads();
    int n_present = 0;
    for (int t = 0; t < a.n_tasks; ++t) n_present += s_mask[t] != nullptr;
    const bool has_mask = n_present > 0;
    const uint32_t thr_bytes = 0x01010101u * (uint32_t)(a.strategy == kUnion ? 1 : n_present);
    const bool majority = a.strategy == kMajority;
    uint32_t* packed = has_mask ? a.packed + a.pmask_off[p] : nullptr;
    uint32_t cnt = 0;
    for (int64_t e0 = start; e0 < stop; e0 += kStep) {
        const int64_t e = e0 + (int64_t)tid * kVec;
        uint32_t bits = 0;
        if (e < stop) {
            const bool full = e + kVec <= numel;
            const uint32_t valid = full ? 0xFu : ((1u << (int)(numel - e)) - 1u);
            if (has_mask) {
                uint32_t votes = 0;
                for (int t = 0; t < a.n_tasks; ++t) {
                    const uint8_t* mp = s_mask[t];
                    if (mp == nullptr) continue;
                    uint32_t w = 0;
                    if (full) w = ldg_stream_u32(mp + e);
                    else
                        for (int c = 0; c < kVec; ++c)
                            if (e + c < numel) w |= (uint32_t)__ldg(mp + e + c) << (8 * c);
                    votes += __vminu4(w, 0x01010101u);
                }
                if (majority) votes += votes;
                const uint32_t ge = __vcmpgeu4(votes, thr_bytes);
                bits = (((ge >> 7) & 1u) | ((ge >> 14) & 2u) | ((ge >> 21) & 4u) | ((ge >> 28) & 8u)) & valid;
            } else {
                bits = valid;
            }
        }
        cnt += __popc(bits);
        if (has_mask) {
            uint32_t w = bits << ((lane & 7) * 4);
            w |= __shfl_xor_sync(0xffffffffu, w, 1);
            w |= __shfl_xor_sync(0xffffffffu, w, 2);
            w |= __shfl_xor_sync(0xffffffffu, w, 4);
            if ((lane & 7) == 0 && e < stop) packed[e >> 5] = w;
        }
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if (lane == 0) s_cnt[warp] = cnt;
    __syncthreads();
    if (tid == 0) {
        uint32_t c = 0;
        for (int w = 0; w < kBlock / 32; ++w) c += s_cnt[w];
        a.count[tile] = c;
    }
}

cudaError_t k6_mask_pack_launch(const K6MaskArgs& a, int n_tiles, cudaStream_t st) {
    if (n_tiles <= 0) return cudaSuccess;
    if (a.n_tasks < 1 || a.n_tasks > kMaxTasks) return cudaErrorInvalidValue;
    k6_mask_pack<<<n_tiles, kBlock, 0, st>>>(a);
    return cudaGetLastError();
}
#endif  // SVDQ_DTYPE == 0

// two elements per thread
constexpr int kWVec = 2;
constexpr int kWStep = kBlock * kWVec;

template <typename T> struct Elem2;
template <> struct Elem2<float> {
    static __device__ __forceinline__ void load(const void* p, int64_t e, bool full, int64_t numel, float (&o)[2]) {
        const float* q = reinterpret_cast<const float*>(p);
        if (full) { const float2 v = __ldg(reinterpret_cast<const float2*>(q + e)); o[0] = v.x; o[1] = v.y; }
        else { o[0] = e < numel ? __ldg(q + e) : 0.0f; o[1] = e + 1 < numel ? __ldg(q + e + 1) : 0.0f; }
    }
};
template <> struct Elem2<__nv_bfloat16> {
    static __device__ __forceinline__ void load(const void* p, int64_t e, bool full, int64_t numel, float (&o)[2]) {
        const __nv_bfloat16* q = reinterpret_cast<const __nv_bfloat16*>(p);
        o[0] = e < numel ? __bfloat162float(q[e]) : 0.0f;
        o[1] = e + 1 < numel ? __bfloat162float(q[e + 1]) : 0.0f;
    }
};
template <> struct Elem2<__half> {
    static __device__ __forceinline__ void load(const void* p, int64_t e, bool full, int64_t numel, float (&o)[2]) {
        const __half* q = reinterpret_cast<const __half*>(p);
        o[0] = e < numel ? __half2float(q[e]) : 0.0f;
        o[1] = e + 1 < numel ? __half2float(q[e + 1]) : 0.0f;
    }
};

// RP: compile-time bound on the number of basis columns (24 or 32)
// NOISE (svd_include_noise): a second coefficient set for the rows outside the combined mask; every element
// picks the set of its own region (index 0 = masked rows, 1 = the rest), so the two elements of a thread may
// read different rows of sW -- the region stride is padded by 4 floats to keep the two 16-byte reads on
// different banks.
template <typename T, int RP, bool FP16B, bool DIAG, bool NOISE>
__global__ void __launch_bounds__(kBlock) k6_reconstruct_merge(const K3Args a, const int n_tasks) {
    extern __shared__ __align__(16) float dyn[];                  // DIAG: [5 * n_tasks][kBlock] accumulators
    constexpr int NREG = NOISE ? 2 : 1;
    constexpr int kRegStride = kMaxTasks * RP + 4;
    __shared__ __align__(16) float sWf[NREG * kRegStride];        // sW(g, t, j) = W_g[t][j]
    __shared__ float sChat[DIAG ? kMaxTasks : 1][RP];
    __shared__ float sCbar[NREG][RP], sSW[NREG][RP];
    __shared__ const void* s_ptr[kMaxTasks + 1];
#define SW(g, t, j) sWf[(g) * kRegStride + (t) * RP + (j)]
    const int N = n_tasks;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x;
    const int p = a.tile_param[tile];
    const int64_t numel = a.numel[p];
    const int64_t start = (int64_t)a.tile_local[tile] * a.tile_elems;
    const int64_t stop = min(start + (int64_t)a.tile_elems, numel);
    const int status = a.info[(int64_t)p * 8 + 0];
    const int n_active = a.info[(int64_t)p * 8 + 1];
    const int r = a.info[(int64_t)p * 8 + (DIAG ? 2 : 4)];
    const float tail_add = a.scal[(int64_t)p * 4 + 1];
    const bool has_mask = a.has_mask[p] != 0;
    const bool noise_on = NOISE && status == kSolved && a.info_n[(int64_t)p * 8 + 0] == kSolved;
    const int r_n = noise_on ? a.info_n[(int64_t)p * 8 + 4] : 0;
    const float tail_n = noise_on ? a.scal_n[(int64_t)p * 4 + 1] : 0.0f;
    if (tid <= N) s_ptr[tid] = a.tensors[(int64_t)p * (N + 1) + tid];
    for (int i = tid; i < kMaxTasks * RP; i += kBlock) {
        const int t = i / RP, j = i % RP;
        const bool ok = t < N && j < N;
        SW(0, t, j) = ok ? a.W[(int64_t)p * N * N + t * N + j] : 0.0f;
        if (NOISE) SW(NREG - 1, t, j) = (ok && noise_on) ? a.W_n[(int64_t)p * N * N + t * N + j] : 0.0f;
        if (DIAG) sChat[t][j] = ok ? a.chat[(int64_t)p * N * N + t * N + j] : 0.0f;
    }
    if (tid < RP) {
        sCbar[0][tid] = tid < N ? a.cbar[(int64_t)p * N + tid] : 0.0f;
        if (NOISE) sCbar[NREG - 1][tid] = (tid < N && noise_on) ? a.cbar_n[(int64_t)p * N + tid] : 0.0f;
    }
    if (DIAG) for (int i = tid; i < kDiagRows * N * kBlock; i += kBlock) dyn[i] = 0.0f;
    __syncthreads();
    if (tid < RP) {                                               // column sums of W over the active tasks
#pragma unroll
        for (int g = 0; g < NREG; ++g) {
            float s = 0.0f;
            for (int t = 0; t < N; ++t) if (s_ptr[t + 1] != nullptr) s += SW(g, t, tid);
            sSW[g][tid] = s;
        }
    }
    __syncthreads();
    const uint32_t* packed = has_mask ? a.packed + a.pmask_off[p] : nullptr;
    float* outp = a.out[p];
    const float n_f = (float)(n_active > 0 ? n_active : 1);

    for (int64_t e0 = start; e0 < stop; e0 += kWStep) {
        const int64_t e = e0 + (int64_t)tid * kWVec;
        if (e >= stop) continue;
        const bool full = e + kWVec <= numel;
        float b[kWVec];
        Elem2<T>::load(s_ptr[0], e, full, numel, b);
        float res[kWVec] = {b[0], b[1]};
        if (status == kSolved) {
            float u[RP][kWVec];
#pragma unroll
            for (int j = 0; j < RP; ++j) { u[j][0] = 0.0f; u[j][1] = 0.0f; }
            float mean[kWVec] = {0.0f, 0.0f};
            uint32_t bits = 0x3u;
            if (has_mask) bits = (__ldg(packed + (e >> 5)) >> (int)(e & 31)) & 0x3u;
            // region of each of the thread's two elements (0 = masked rows)
            const int g0 = (NOISE && !(bits & 1u)) ? NREG - 1 : 0, g1 = (NOISE && !(bits & 2u)) ? NREG - 1 : 0;
            for (int t = 0; t < N; ++t) {                        // stream the tasks once
                const void* fp = s_ptr[t + 1];
                if (fp == nullptr) continue;
                float f[kWVec];
                Elem2<T>::load(fp, e, full, numel, f);
                const float x0 = Elem<T>::sub(f[0], b[0]), x1 = Elem<T>::sub(f[1], b[1]);
                mean[0] += x0; mean[1] += x1;
#pragma unroll
                for (int j = 0; j < RP; j += 4) {
                    const float4 w = *reinterpret_cast<const float4*>(&SW(g0, t, j));
                    const float4 v = NOISE ? *reinterpret_cast<const float4*>(&SW(g1, t, j)) : w;
                    u[j + 0][0] = fmaf(x0, w.x, u[j + 0][0]); u[j + 0][1] = fmaf(x1, v.x, u[j + 0][1]);
                    u[j + 1][0] = fmaf(x0, w.y, u[j + 1][0]); u[j + 1][1] = fmaf(x1, v.y, u[j + 1][1]);
                    u[j + 2][0] = fmaf(x0, w.z, u[j + 2][0]); u[j + 2][1] = fmaf(x1, v.z, u[j + 2][1]);
                    u[j + 3][0] = fmaf(x0, w.w, u[j + 3][0]); u[j + 3][1] = fmaf(x1, v.w, u[j + 3][1]);
                }
            }
#pragma unroll
            for (int c = 0; c < kWVec; ++c) mean[c] = a.center ? __fdiv_rn(mean[c], n_f) : 0.0f;
            // u = (tau - mean) W = tau W - mean * colsum(W); fp16 round trip; contract with cbar
            float acc[kWVec] = {0.0f, 0.0f};
            const int gsel[kWVec] = {g0, g1};
            const int rsel[kWVec] = {g0 ? r_n : r, g1 ? r_n : r};
#pragma unroll
            for (int j = 0; j < RP; ++j) {
#pragma unroll
                for (int c = 0; c < kWVec; ++c) {
                    if (j < rsel[c]) {
                        float v = fmaf(-mean[c], sSW[gsel[c]][j], u[j][c]);
                        if (FP16B) v = round_fp16(v);
                        u[j][c] = v;
                        acc[c] = fmaf(v, sCbar[gsel[c]][j], acc[c]);
                    } else {
                        u[j][c] = 0.0f;
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < kWVec; ++c) {
                const bool m = (bits >> c) & 1u;
                float val = (acc[c] + mean[c]) + (gsel[c] ? tail_n : tail_add);
                if (NOISE && !m) val = noise_on ? __fmul_rn(val, a.noise_shrink) : 0.0f;
                res[c] = b[c] + ((m || NOISE) ? val : 0.0f);
            }
            if (DIAG) {
                for (int t = 0; t < N; ++t) {
                    const void* fp = s_ptr[t + 1];
                    if (fp == nullptr) continue;
                    float f[kWVec];
                    Elem2<T>::load(fp, e, full, numel, f);
                    const float x[kWVec] = {Elem<T>::sub(f[0], b[0]), Elem<T>::sub(f[1], b[1])};
                    float rec[kWVec] = {0.0f, 0.0f};
#pragma unroll
                    for (int j = 0; j < RP; ++j) {
                        const float ch = sChat[t][j];
                        rec[0] = fmaf(u[j][0], ch, rec[0]); rec[1] = fmaf(u[j][1], ch, rec[1]);
                    }
#pragma unroll
                    for (int c = 0; c < kWVec; ++c) {
                        if (!(((bits >> c) & 1u) && e + c < numel)) continue;
                        const float er = x[c] - rec[c];
                        float* d0 = dyn + (size_t)(0 * N + t) * kBlock + tid;
                        float* d1 = dyn + (size_t)(1 * N + t) * kBlock + tid;
                        float* d2 = dyn + (size_t)(2 * N + t) * kBlock + tid;
                        float* d3 = dyn + (size_t)(3 * N + t) * kBlock + tid;
                        *d0 = fmaf(er, er, *d0); *d1 += fabsf(er); *d2 = fmaf(rec[c], rec[c], *d2);
                        *d3 = fmaxf(*d3, fabsf(er));
                    }
                }
            }
        }
        if (full) *reinterpret_cast<float2*>(outp + e) = make_float2(res[0], res[1]);
        else if (e < numel) outp[e] = res[0];
    }

    if (DIAG) {
        __syncthreads();
        const int NR = kDiagRows * N;
        float* dout = a.diag + (int64_t)tile * NR;
        for (int row = warp; row < NR; row += kBlock / 32) {       // one warp reduces one row, fixed order
            const bool is_max = row >= 3 * N;
            float s = 0.0f;
            for (int c = 0; c < kBlock / 32; ++c) {
                const float v = dyn[(size_t)row * kBlock + lane + 32 * c];
                s = is_max ? fmaxf(s, v) : s + v;
            }
            for (int o = 16; o > 0; o >>= 1) {
                const float v = __shfl_xor_sync(0xffffffffu, s, o);
                s = is_max ? fmaxf(s, v) : s + v;
            }
            if (lane == 0) dout[row] = s;
        }
    }
}

template <typename T, int RP>
static cudaError_t launch_wide(const K3Args& a, int n_tasks, int n_tiles, bool fp16b, bool diag, cudaStream_t st) {
    const size_t dsm = diag ? (size_t)kDiagRows * n_tasks * kBlock * sizeof(float) : 0;
    cudaError_t e = cudaSuccess;
    const bool noise = a.info_n != nullptr;
#define SVDQ_GO2(F, D, Z)                                                                                        \
    do {                                                                                                         \
        if (dsm) e = cudaFuncSetAttribute(k6_reconstruct_merge<T, RP, F, D, Z>,                                  \
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm);                \
        if (e != cudaSuccess) return e;                                                                          \
        k6_reconstruct_merge<T, RP, F, D, Z><<<n_tiles, kBlock, dsm, st>>>(a, n_tasks);                          \
    } while (0)
#define SVDQ_GO(F, D)                                                                                            \
    do {                                                                                                         \
        if (noise) SVDQ_GO2(F, D, true);                                                                         \
        else SVDQ_GO2(F, D, false);                                                                              \
    } while (0)
    if (fp16b && diag) SVDQ_GO(true, true);
    else if (fp16b) SVDQ_GO(true, false);
    else if (diag) SVDQ_GO(false, true);
    else SVDQ_GO(false, false);
#undef SVDQ_GO
#undef SVDQ_GO2
    return cudaGetLastError();
}

template <>
cudaError_t k6_merge_launch_dtype<SVDQ_DTYPE>(int n_tasks, const K3Args& a, int n_tiles, bool fp16b, bool diag,
                                              cudaStream_t st) {
    using T = DTypeOf<SVDQ_DTYPE>::type;
    if (n_tiles <= 0) return cudaSuccess;
    if (n_tasks < 1 || n_tasks > kMaxTasks) return cudaErrorInvalidValue;
    if (n_tasks <= 24) return launch_wide<T, 24>(a, n_tasks, n_tiles, fp16b, diag, st);
    return launch_wide<T, 32>(a, n_tasks, n_tiles, fp16b, diag, st);
}

#undef SW
}  // namespace svdq
