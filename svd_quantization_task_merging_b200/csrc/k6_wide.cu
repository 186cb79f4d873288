// K6 — "wide" path for 17..32 task vectors (configs[3]: 20 tasks), where the N(N+1)/2 Gram
// accumulators no longer fit one thread's registers:
//   k6_mask_pack            combines the N tall masks once (runtime N) -> packed mask + per-tile counts;
//                           the Gram is then accumulated by k8_gram_staged under this mask
//   k6_reconstruct_merge    pass 2 with a runtime number of tasks (two elements per thread, the centred
//                           task vectors in registers as pairs of tasks); optional fused diagnostics
//                           accumulate in shared memory
// Same reference lines as K1 / K3.  Bound: HBM for the mask pass; pass 2 is issue-bound at N = 20
// (N*r = 400 FMAs per element) -- see DESIGN.md.
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
    // 16 elements (one 128-bit mask load per task) per thread and step, up to twenty tasks' loads in flight
    // (measured on ViT-L-14 x 20: 8 in flight 1.355 ms, 10: 1.185 ms, 20: 1.169 ms)
    constexpr int kMV = 16;
    for (int64_t e0 = start; e0 < stop; e0 += (int64_t)kBlock * kMV) {
        const int64_t e = e0 + (int64_t)tid * kMV;
        uint32_t bits = 0;
        if (e < stop) {
            const bool full = e + kMV <= numel;
            const uint32_t valid = full ? 0xFFFFu : ((1u << (int)(numel - e)) - 1u);
            if (has_mask) {
                uint32_t votes[4] = {0u, 0u, 0u, 0u};
                constexpr int kUT = 20;
                for (int t0 = 0; t0 < a.n_tasks; t0 += kUT) {
                    uint4 w[kUT];
#pragma unroll
                    for (int q = 0; q < kUT; ++q) {
                        const uint8_t* mp = (t0 + q < a.n_tasks) ? s_mask[t0 + q] : nullptr;
                        w[q] = make_uint4(0u, 0u, 0u, 0u);
                        if (mp == nullptr) continue;
                        if (full) {
                            asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                                         : "=r"(w[q].x), "=r"(w[q].y), "=r"(w[q].z), "=r"(w[q].w) : "l"(mp + e));
                        } else {
                            uint32_t ww[4] = {0u, 0u, 0u, 0u};
                            for (int c = 0; c < kMV; ++c)
                                if (e + c < numel) ww[c >> 2] |= (uint32_t)__ldg(mp + e + c) << (8 * (c & 3));
                            w[q] = make_uint4(ww[0], ww[1], ww[2], ww[3]);
                        }
                    }
#pragma unroll
                    for (int q = 0; q < kUT; ++q) {
                        votes[0] += __vminu4(w[q].x, 0x01010101u); votes[1] += __vminu4(w[q].y, 0x01010101u);
                        votes[2] += __vminu4(w[q].z, 0x01010101u); votes[3] += __vminu4(w[q].w, 0x01010101u);
                    }
                }
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    uint32_t v = votes[g];
                    if (majority) v += v;
                    const uint32_t ge = __vcmpgeu4(v, thr_bytes);
                    bits |= (((ge >> 7) & 1u) | ((ge >> 14) & 2u) | ((ge >> 21) & 4u) | ((ge >> 28) & 8u)) << (4 * g);
                }
                bits &= valid;
            } else {
                bits = valid;
            }
        }
        cnt += __popc(bits);
        if (has_mask) {                                  // two lanes share one 32-bit word of the packed mask
            uint32_t w = bits << ((lane & 1) * 16);
            w |= __shfl_xor_sync(0xffffffffu, w, 1);
            if ((lane & 1) == 0 && e < stop) packed[e >> 5] = w;
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

// ---- pass 2 for 17..32 task vectors ------------------------------------------------------------------------
// Two elements per thread.  All (N+1) loads of a step are issued back to back, the centred task vectors stay in
// registers as PAIRS OF TASKS (x_2q, x_2q+1), and every basis-row entry u_dj = sum_t x_dt W[t][j] is a chain of
// packed FMAs over task pairs: the second operand (W[2q][j], W[2q+1][j]) is one aligned 64-bit word of shared
// memory, so no register moves are needed to build packed operands (with two ELEMENTS per FMA every W entry had
// to be duplicated first: 21 % of the instructions of the previous version).  Even / odd partial sums are
// added at the end.  NTMAX (20, 24 or 32) bounds the unrolled loops; tasks t >= N read nothing and count as zero.
constexpr int kWVec = 2;
constexpr int kWStep = kBlock * kWVec;

template <typename T> using Elem2 = ElemPair<T>;

// NOISE (svd_include_noise): a second coefficient set for the rows outside the combined mask; an element picks the
// set of its own region (0 = masked rows, 1 = the rest) simply by the base address of its W / cbar reads.
template <typename T, int NTMAX, bool FP16B, bool DIAG, bool NOISE>
__global__ void __launch_bounds__(kBlock, (NTMAX <= 24 && !DIAG) ? 2 : 1) k6_reconstruct_merge(const K3Args a,
                                                                                               const int n_tasks) {
    extern __shared__ __align__(16) float dyn[];                  // DIAG: [4 * n_tasks][kBlock] accumulators
    constexpr int NREG = NOISE ? 2 : 1;
    constexpr int TP = NTMAX / 2;                                 // task pairs
    constexpr int kRegStride = NTMAX * TP + 2;                    // float2 units; +2: regions on different banks
    __shared__ __align__(16) float2 sW2[NREG * kRegStride];       // W2(g, j, q) = (W_g[2q][j], W_g[2q+1][j])
    __shared__ __align__(16) float2 sChat2[DIAG ? NTMAX : 1][TP]; // (chat[2q][j], chat[2q+1][j])
    __shared__ float sCbar[NREG][NTMAX];
    __shared__ const void* s_ptr[NTMAX + 1];
    __shared__ uint32_t s_present;
#define W2(g, j, q) sW2[(g) * kRegStride + (j) * TP + (q)]
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
    const float mean_scale = a.scal[(int64_t)p * 4 + 2];
    const bool has_mask = a.has_mask[p] != 0;
    const bool noise_on = NOISE && status == kSolved && a.info_n[(int64_t)p * 8 + 0] == kSolved;
    const int r_n = noise_on ? a.info_n[(int64_t)p * 8 + 4] : 0;
    const float tail_n = noise_on ? a.scal_n[(int64_t)p * 4 + 1] : 0.0f;
    const float mean_scale_n = noise_on ? a.scal_n[(int64_t)p * 4 + 2] : 1.0f;
    const int r_loop = NOISE ? max(r, r_n) : r;
    if (tid <= NTMAX) s_ptr[tid] = tid <= N ? a.tensors[(int64_t)p * (N + 1) + tid] : nullptr;
    for (int i = tid; i < NTMAX * TP; i += kBlock) {
        const int j = i / TP, q = i % TP;
        const int t0 = 2 * q, t1 = 2 * q + 1;
        auto at = [&](const float* M, int t) { return (t < N && j < N) ? M[(int64_t)p * N * N + t * N + j] : 0.0f; };
        W2(0, j, q) = make_float2(at(a.W, t0), at(a.W, t1));
        if (NOISE) W2(NREG - 1, j, q) = noise_on ? make_float2(at(a.W_n, t0), at(a.W_n, t1)) : make_float2(0.0f, 0.0f);
        if (DIAG) sChat2[j][q] = make_float2(at(a.chat, t0), at(a.chat, t1));
    }
    if (tid < NTMAX) {
        sCbar[0][tid] = tid < N ? a.cbar[(int64_t)p * N + tid] : 0.0f;
        if (NOISE) sCbar[NREG - 1][tid] = (tid < N && noise_on) ? a.cbar_n[(int64_t)p * N + tid] : 0.0f;
    }
    if (DIAG) for (int i = tid; i < kDiagRows * N * kBlock; i += kBlock) dyn[i] = 0.0f;
    __syncthreads();
    if (tid == 0) {
        uint32_t pb = 0;
        for (int t = 0; t < N; ++t) pb |= (s_ptr[t + 1] != nullptr ? 1u : 0u) << t;
        s_present = pb;
    }
    __syncthreads();
    const uint32_t present_bits = s_present;
    // a task that lacks the parameter (or t >= N) reads the base tensor: zero delta, dropped again after centring
    if (tid >= 1 && tid <= NTMAX && s_ptr[tid] == nullptr) s_ptr[tid] = s_ptr[0];
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
            // ---- all loads of the step back to back ----------------------------------------------------------
            float2 xp[TP][kWVec];                                // xp[q][c] = (x_2q, x_2q+1) of element c
#pragma unroll
            for (int q = 0; q < TP; ++q) {
                float f0[kWVec] = {b[0], b[1]}, f1[kWVec] = {b[0], b[1]};
                if (2 * q < N) Elem2<T>::load(s_ptr[2 * q + 1], e, full, numel, f0);
                if (2 * q + 1 < N) Elem2<T>::load(s_ptr[2 * q + 2], e, full, numel, f1);
#pragma unroll
                for (int c = 0; c < kWVec; ++c) xp[q][c] = make_float2(f0[c], f1[c]);
            }
            uint32_t bits = 0x3u;
            if (has_mask) bits = (__ldg(packed + (e >> 5)) >> (int)(e & 31)) & 0x3u;
            // ---- task vectors, mean over the active tasks (summed in task order), centring --------------------
            float mean[kWVec] = {0.0f, 0.0f};
#pragma unroll
            for (int q = 0; q < TP; ++q)
#pragma unroll
                for (int c = 0; c < kWVec; ++c) {
                    xp[q][c].x = Elem<T>::sub(xp[q][c].x, b[c]);
                    xp[q][c].y = Elem<T>::sub(xp[q][c].y, b[c]);
                    mean[c] += xp[q][c].x;
                    mean[c] += xp[q][c].y;
                }
#pragma unroll
            for (int c = 0; c < kWVec; ++c) mean[c] = a.center ? __fdiv_rn(mean[c], n_f) : 0.0f;
#pragma unroll
            for (int q = 0; q < TP; ++q)
#pragma unroll
                for (int c = 0; c < kWVec; ++c) {
                    xp[q][c].x = ((present_bits >> (2 * q)) & 1u) ? xp[q][c].x - mean[c] : 0.0f;
                    xp[q][c].y = ((present_bits >> (2 * q + 1)) & 1u) ? xp[q][c].y - mean[c] : 0.0f;
                }
            // ---- basis rows column by column, contraction with cbar (and chat for the diagnostics) -----------
            const float2* wbase[kWVec];
            const float* cbase[kWVec];
#pragma unroll
            for (int c = 0; c < kWVec; ++c) {
                const int g = (NOISE && !((bits >> c) & 1u)) ? NREG - 1 : 0;
                wbase[c] = sW2 + g * kRegStride;
                cbase[c] = sCbar[g];
            }
            constexpr int kJB = (NTMAX % 4 == 0 && !DIAG) ? 4 : 2;
            float acc[kWVec] = {0.0f, 0.0f};
            float2 rec2[DIAG ? TP : 1][kWVec];
            if (DIAG) {
#pragma unroll
                for (int q = 0; q < TP; ++q)
#pragma unroll
                    for (int c = 0; c < kWVec; ++c) rec2[q][c] = make_float2(0.0f, 0.0f);
            }
#pragma unroll
            for (int j0 = 0; j0 < NTMAX; j0 += kJB) {            // kJB columns at a time: 2 * kJB independent FMA chains
                if (j0 >= r_loop) break;
                float u[kJB][kWVec];
                float2 s2[kJB][kWVec];
#pragma unroll
                for (int z = 0; z < kJB; ++z)
#pragma unroll
                    for (int c = 0; c < kWVec; ++c) s2[z][c] = make_float2(0.0f, 0.0f);
#pragma unroll
                for (int q = 0; q < TP; q += 2) {
#pragma unroll
                    for (int z = 0; z < kJB; ++z)
#pragma unroll
                        for (int c = 0; c < kWVec; ++c) {
                            const float4 w = *reinterpret_cast<const float4*>(wbase[c] + (j0 + z) * TP + q);
                            s2[z][c] = __ffma2_rn(xp[q][c], make_float2(w.x, w.y), s2[z][c]);
                            s2[z][c] = __ffma2_rn(xp[q + 1][c], make_float2(w.z, w.w), s2[z][c]);
                        }
                }
#pragma unroll
                for (int z = 0; z < kJB; ++z) {
#pragma unroll
                    for (int c = 0; c < kWVec; ++c) u[z][c] = s2[z][c].x + s2[z][c].y;
                    if (FP16B) {
                        const float2 h = __half22float2(__float22half2_rn(make_float2(u[z][0], u[z][1])));
                        u[z][0] = h.x; u[z][1] = h.y;
                    }
                }
#pragma unroll
                for (int z = 0; z < kJB; ++z) {
                    // a column at or beyond r_loop has zero W and zero cbar / chat entries: contributes exactly 0
                    const int j = j0 + z;
#pragma unroll
                    for (int c = 0; c < kWVec; ++c) acc[c] = fmaf(u[z][c], cbase[c][j], acc[c]);
                    if (DIAG) {
#pragma unroll
                        for (int c = 0; c < kWVec; ++c) {
                            const float2 ud = make_float2(u[z][c], u[z][c]);
#pragma unroll
                            for (int q = 0; q < TP; ++q) rec2[q][c] = __ffma2_rn(ud, sChat2[j][q], rec2[q][c]);
                        }
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < kWVec; ++c) {
                const bool m = (bits >> c) & 1u;
                float val = fmaf(mean[c], (NOISE && !m) ? mean_scale_n : mean_scale, acc[c]) + ((NOISE && !m) ? tail_n : tail_add);
                if (NOISE && !m) val = noise_on ? __fmul_rn(val, a.noise_shrink) : 0.0f;
                res[c] = b[c] + ((m || NOISE) ? val : 0.0f);
            }
            if (DIAG) {
                // the diagnostics compare with the UNCENTRED masked task vector: re-read it (L2 hit)
#pragma unroll
                for (int q = 0; q < TP; ++q) {
#pragma unroll
                    for (int z = 0; z < 2; ++z) {
                        const int t = 2 * q + z;
                        if (t >= N || !((present_bits >> t) & 1u)) continue;
                        float f[kWVec];
                        Elem2<T>::load(s_ptr[t + 1], e, full, numel, f);
                        float* d0 = dyn + (size_t)(0 * N + t) * kBlock + tid;
                        float* d1 = dyn + (size_t)(1 * N + t) * kBlock + tid;
                        float* d2 = dyn + (size_t)(2 * N + t) * kBlock + tid;
                        float* d3 = dyn + (size_t)(3 * N + t) * kBlock + tid;
                        float se = *d0, sa = *d1, sr = *d2, mx = *d3;
#pragma unroll
                        for (int c = 0; c < kWVec; ++c) {
                            if (!(((bits >> c) & 1u) && e + c < numel)) continue;
                            const float rc = z ? rec2[q][c].y : rec2[q][c].x;
                            const float er = Elem<T>::sub(f[c], b[c]) - rc;
                            se = fmaf(er, er, se); sa += fabsf(er); sr = fmaf(rc, rc, sr); mx = fmaxf(mx, fabsf(er));
                        }
                        *d0 = se; *d1 = sa; *d2 = sr; *d3 = mx;
                    }
                }
            }
        }
        if (full) stg_stream_f2(outp + e, make_float2(res[0], res[1]));
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
#undef W2
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
    if (n_tasks <= 20) return launch_wide<T, 20>(a, n_tasks, n_tiles, fp16b, diag, st);   // the standard 20-task merge
    if (n_tasks <= 24) return launch_wide<T, 24>(a, n_tasks, n_tiles, fp16b, diag, st);
    return launch_wide<T, 32>(a, n_tasks, n_tiles, fp16b, diag, st);
}

}  // namespace svdq
