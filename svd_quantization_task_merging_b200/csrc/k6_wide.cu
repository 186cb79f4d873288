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

// four elements per thread (two packed pairs)
constexpr int kWVec = 4;
constexpr int kWStep = kBlock * kWVec;

template <typename T> struct Elem2 {
    static __device__ __forceinline__ void load(const void* p, int64_t e, bool full, int64_t numel, float (&o)[kWVec]) {
        if (full) Elem<T>::load4(p, e, o);
        else {
#pragma unroll
            for (int c = 0; c < kWVec; ++c) o[c] = (e + c < numel) ? Elem<T>::load1(p, e + c) : 0.0f;
        }
    }
};

// RP: compile-time bound on the number of basis columns (24 or 32)
// NOISE (svd_include_noise): a second coefficient set for the rows outside the combined mask.  Every element
// belongs to exactly one region, so the task value is split into (masked part, unmasked part) -- one of them is
// zero -- and both parts are contracted with their own W into the SAME accumulators (uniform broadcast reads
// of W, no per-element select in the inner loop); the column sums, cbar and tails are picked per element.
template <typename T, int RP, bool FP16B, bool DIAG, bool NOISE>
__global__ void __launch_bounds__(kBlock, (RP <= 24 && !NOISE) ? 2 : 1) k6_reconstruct_merge(const K3Args a,
                                                                                             const int n_tasks) {
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

    constexpr int kH = kWVec / 2;
    for (int64_t e0 = start; e0 < stop; e0 += kWStep) {
        const int64_t e = e0 + (int64_t)tid * kWVec;
        if (e >= stop) continue;
        const bool full = e + kWVec <= numel;
        float b[kWVec];
        Elem2<T>::load(s_ptr[0], e, full, numel, b);
        float res[kWVec];
#pragma unroll
        for (int c = 0; c < kWVec; ++c) res[c] = b[c];
        if (status == kSolved) {
            float2 u[RP][kH];                                    // basis-row accumulators, packed element pairs
#pragma unroll
            for (int j = 0; j < RP; ++j)
#pragma unroll
                for (int h = 0; h < kH; ++h) u[j][h] = make_float2(0.0f, 0.0f);
            float mean[kWVec];
#pragma unroll
            for (int c = 0; c < kWVec; ++c) mean[c] = 0.0f;
            uint32_t bits = (1u << kWVec) - 1u;
            if (has_mask) bits = (__ldg(packed + (e >> 5)) >> (int)(e & 31)) & ((1u << kWVec) - 1u);
            // stream the tasks once, kUT at a time: their loads are issued back to back before any use (a task that
            // lacks the parameter, or t >= N, reads the base tensor -> zero delta; its W row is zero / unused)
            constexpr int kUT = 4;
            for (int t0 = 0; t0 < N; t0 += kUT) {
                float f[kUT][kWVec];
#pragma unroll
                for (int q = 0; q < kUT; ++q) {
                    const void* fp = (t0 + q < N) ? s_ptr[t0 + q + 1] : nullptr;
                    Elem2<T>::load(fp ? fp : s_ptr[0], e, full, numel, f[q]);
                }
#pragma unroll
                for (int q = 0; q < kUT; ++q) {
                    const int t = t0 + q;                        // sW rows t >= N are zero
                    float2 xm[kH], xn[NOISE ? kH : 1];
#pragma unroll
                    for (int h = 0; h < kH; ++h) {
                        const float x0 = Elem<T>::sub(f[q][2 * h], b[2 * h]);
                        const float x1 = Elem<T>::sub(f[q][2 * h + 1], b[2 * h + 1]);
                        mean[2 * h] += x0; mean[2 * h + 1] += x1;
                        if (NOISE) {
                            const bool m0 = (bits >> (2 * h)) & 1u, m1 = (bits >> (2 * h + 1)) & 1u;
                            xm[h] = make_float2(m0 ? x0 : 0.0f, m1 ? x1 : 0.0f);
                            xn[h] = make_float2(m0 ? 0.0f : x0, m1 ? 0.0f : x1);
                        } else xm[h] = make_float2(x0, x1);
                    }
#pragma unroll
                    for (int j = 0; j < RP; j += 4) {
                        const float4 w = *reinterpret_cast<const float4*>(&SW(0, t, j));
                        const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                        for (int z = 0; z < 4; ++z) {
                            const float2 w2 = make_float2(wv[z], wv[z]);
#pragma unroll
                            for (int h = 0; h < kH; ++h) u[j + z][h] = __ffma2_rn(xm[h], w2, u[j + z][h]);
                        }
                        if (NOISE) {
                            const float4 v = *reinterpret_cast<const float4*>(&SW(NREG - 1, t, j));
                            const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                            for (int z = 0; z < 4; ++z) {
                                const float2 v2 = make_float2(vv[z], vv[z]);
#pragma unroll
                                for (int h = 0; h < kH; ++h) u[j + z][h] = __ffma2_rn(xn[h], v2, u[j + z][h]);
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < kWVec; ++c) mean[c] = a.center ? __fdiv_rn(mean[c], n_f) : 0.0f;
            // u = (tau - mean) W = tau W - mean * colsum(W); fp16 round trip; contract with cbar
            float acc[kWVec];
            int gsel[kWVec], rsel[kWVec];
#pragma unroll
            for (int c = 0; c < kWVec; ++c) {
                acc[c] = 0.0f;
                gsel[c] = (NOISE && !((bits >> c) & 1u)) ? NREG - 1 : 0;
                rsel[c] = gsel[c] ? r_n : r;
            }
#pragma unroll
            for (int j = 0; j < RP; ++j) {
#pragma unroll
                for (int c = 0; c < kWVec; ++c) {
                    float& uc = (c & 1) ? u[j][c >> 1].y : u[j][c >> 1].x;
                    if (j < rsel[c]) {
                        float v = fmaf(-mean[c], sSW[gsel[c]][j], uc);
                        if (FP16B) v = round_fp16(v);
                        uc = v;
                        acc[c] = fmaf(v, sCbar[gsel[c]][j], acc[c]);
                    } else {
                        uc = 0.0f;
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
                    float2 rec2[kH];
#pragma unroll
                    for (int h = 0; h < kH; ++h) rec2[h] = make_float2(0.0f, 0.0f);
#pragma unroll
                    for (int j = 0; j < RP; ++j) {
                        const float ch = sChat[t][j];
                        const float2 ch2 = make_float2(ch, ch);
#pragma unroll
                        for (int h = 0; h < kH; ++h) rec2[h] = __ffma2_rn(u[j][h], ch2, rec2[h]);
                    }
                    float* d0 = dyn + (size_t)(0 * N + t) * kBlock + tid;
                    float* d1 = dyn + (size_t)(1 * N + t) * kBlock + tid;
                    float* d2 = dyn + (size_t)(2 * N + t) * kBlock + tid;
                    float* d3 = dyn + (size_t)(3 * N + t) * kBlock + tid;
                    float se = *d0, sa = *d1, sr = *d2, mx = *d3;
#pragma unroll
                    for (int c = 0; c < kWVec; ++c) {
                        if (!(((bits >> c) & 1u) && e + c < numel)) continue;
                        const float rc = (c & 1) ? rec2[c >> 1].y : rec2[c >> 1].x;
                        const float er = Elem<T>::sub(f[c], b[c]) - rc;
                        se = fmaf(er, er, se); sa += fabsf(er); sr = fmaf(rc, rc, sr); mx = fmaxf(mx, fabsf(er));
                    }
                    *d0 = se; *d1 = sa; *d2 = sr; *d3 = mx;
                }
            }
        }
        if (full) stg_stream_f4(outp + e, make_float4(res[0], res[1], res[2], res[3]));
        else {
#pragma unroll
            for (int c = 0; c < kWVec; ++c)
                if (e + c < numel) outp[e + c] = res[c];
        }
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
