// K3 — pass 2 of the SVD-Hybrid merge: re-reads base + N fine-tuned tensors and the packed
// combined mask, rebuilds each basis row on the fly (u_d = (tau_d - mean_d) W, rounded to fp16
// when the bases are stored in fp16), contracts it with the averaged coefficients, scatters
// through the mask and writes merged = base + delta in the same pass.  Optionally accumulates
// the reference's per-task reconstruction diagnostics.
//
// Replaces (paths relative to /root/reference):
//   U = T V Sigma^-1 (implicit in torch.linalg.svd)      src/svd_hybrid/basis.py:241
//   fp16 cast of the bases                                src/svd_hybrid/cli.py:355-361
//   reconstruct_from_coefficients                         src/svd_hybrid/merge.py:180-192
//   reconstruct_from_masked (scatter)                     src/svd_hybrid/mask_loader.py:750-763
//   apply_merged_deltas                                   src/svd_hybrid/merge.py:486-488
//   compute_reconstruction_error (DIAG)                   src/svd_hybrid/diagnostics.py:101-117,205-216
// Bound: HBM (algorithmic bytes per element: (N+1)*sizeof(T) + 1/8 + 4).
#include "k3_body.cuh"

#ifndef SVDQ_DTYPE
#define SVDQ_DTYPE 0
#endif

namespace svdq {



// VEC = elements per thread and step: 4 up to 8 tasks; 2 with fused diagnostics (4*NT running reductions per thread)
// and above 8 tasks (NT x VEC task values per thread), where two elements in flight keep the register count low
// enough for more resident warps.
template <int NT, bool DIAG> struct K3Vec { static constexpr int value = (DIAG || NT > 8) ? 2 : 4; };

template <typename T, int NT, bool FP16B, bool DIAG, bool NOISE>
__global__ void __launch_bounds__(kBlock, (NT <= 8 || !DIAG) ? 2 : 1) k3_reconstruct_merge(const K3Args a) {
    constexpr int VEC = K3Vec<NT, DIAG>::value;
    constexpr int kStepV = kBlock * VEC;
    constexpr int NTP = (NT + 3) & ~3;
    __shared__ __align__(16) float sWT[NT][NTP];        // sWT[j][t] = W[t][j]
    __shared__ __align__(16) float sChatT[DIAG ? NT : 1][NTP];   // sChatT[j][t] = chat[t][j]
    __shared__ float sCbar[NT], sG[NT];
    __shared__ __align__(16) float sWTn[NOISE ? NT : 1][NTP];    // noise-region set (svd_include_noise)
    __shared__ float sCbarN[NOISE ? NT : 1], sGn[NOISE ? NT : 1];
    __shared__ const void* s_ptr[NT + 1];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x;
    const int p = a.tile_param[tile];
    const int64_t numel = a.numel[p];
    const int64_t start = (int64_t)a.tile_local[tile] * a.tile_elems;
    const int64_t stop = min(start + (int64_t)a.tile_elems, numel);
    const int status = a.info[(int64_t)p * 8 + 0];
    const int n_active = a.info[(int64_t)p * 8 + 1];
    // columns beyond r_eff are zero (null directions): their contribution is tail_add; the fused
    // diagnostics still walk all r columns so that a NaN coefficient shows up exactly as in the reference
    const int r = a.info[(int64_t)p * 8 + (DIAG ? 2 : 4)];
    const float tail_add = a.scal[(int64_t)p * 4 + 1];
    const float mean_scale = a.scal[(int64_t)p * 4 + 2];
    const bool has_mask = a.has_mask[p] != 0;

    if (tid <= NT) s_ptr[tid] = a.tensors[(int64_t)p * (NT + 1) + tid];
    for (int i = tid; i < NT * NTP; i += kBlock) {
        const int j = i / NTP, t = i % NTP;
        sWT[j][t] = (t < NT) ? a.W[(int64_t)p * NT * NT + t * NT + j] : 0.0f;
        if (DIAG) sChatT[j][t] = (t < NT) ? a.chat[(int64_t)p * NT * NT + t * NT + j] : 0.0f;
    }
    if (tid < NT) { sCbar[tid] = a.cbar[(int64_t)p * NT + tid]; sG[tid] = a.gvec[(int64_t)p * NT + tid]; }
    K3NoiseSet<NT> ns;
    if (NOISE) {
        ns.on = status == kSolved && a.info_n[(int64_t)p * 8 + 0] == kSolved;
        ns.r = a.info_n[(int64_t)p * 8 + 4];
        ns.tail_add = a.scal_n[(int64_t)p * 4 + 1];
        ns.mean_scale = a.scal_n[(int64_t)p * 4 + 2];
        ns.shrink = a.noise_shrink;
        ns.sWT = sWTn; ns.sCbar = sCbarN; ns.sG = sGn;
        for (int i = tid; i < NT * NTP; i += kBlock) {
            const int j = i / NTP, t = i % NTP;
            sWTn[j][t] = (t < NT) ? a.W_n[(int64_t)p * NT * NT + t * NT + j] : 0.0f;
        }
        if (tid < NT) { sCbarN[tid] = a.cbar_n[(int64_t)p * NT + tid]; sGn[tid] = a.gvec_n[(int64_t)p * NT + tid]; }
    }
    __syncthreads();

    const uint32_t* packed = has_mask ? a.packed + a.pmask_off[p] : nullptr;
    float* outp = a.out[p];
    const float inv_n_dummy = 0.0f; (void)inv_n_dummy;
    const float n_f = (float)(n_active > 0 ? n_active : 1);

    float dacc[DIAG ? kDiagRows * NT : 1];
    if (DIAG) {
#pragma unroll
        for (int i = 0; i < kDiagRows * NT; ++i) dacc[i] = 0.0f;
    }

    // tasks that lack the parameter: pointer replaced by the base tensor (delta == 0), see K1
    uint32_t present_bits = 0;
#pragma unroll
    for (int t = 0; t < NT; ++t) present_bits |= (s_ptr[t + 1] != nullptr ? 1u : 0u) << t;
    __syncthreads();
    if (tid >= 1 && tid <= NT && s_ptr[tid] == nullptr) s_ptr[tid] = s_ptr[0];
    __syncthreads();

    for (int64_t e0 = start; e0 < stop; e0 += kStepV) {
        const int64_t e = e0 + (int64_t)tid * VEC;
        if (e >= stop) continue;
        const bool full = e + VEC <= numel;
        float b[VEC];
        float res[VEC];
        if (status != kSolved) {
            // parameter without a basis (mask below svd_min_mask_size / no data): merged = base
            if constexpr (VEC == 4) {
                if (full) Elem<T>::load4(s_ptr[0], e, b);
                else {
#pragma unroll
                    for (int c = 0; c < VEC; ++c) b[c] = (e + c < numel) ? Elem<T>::load1(s_ptr[0], e + c) : 0.0f;
                }
            } else ElemPair<T>::load(s_ptr[0], e, full, numel, b);
#pragma unroll
            for (int c = 0; c < VEC; ++c) res[c] = b[c];
        } else {
            // ---- phase 1: all loads of the step back to back ----------------------------------------
            float x[NT][VEC];                    // fine-tuned values -> task vectors -> centred task vectors
            uint32_t pword = 0xffffffffu;
            if constexpr (VEC == 4) {
                if (full) {
                    Elem<T>::load4(s_ptr[0], e, b);
#pragma unroll
                    for (int t = 0; t < NT; ++t) Elem<T>::load4(s_ptr[t + 1], e, x[t]);
                } else {
#pragma unroll
                    for (int c = 0; c < VEC; ++c) b[c] = (e + c < numel) ? Elem<T>::load1(s_ptr[0], e + c) : 0.0f;
#pragma unroll
                    for (int t = 0; t < NT; ++t)
#pragma unroll
                        for (int c = 0; c < VEC; ++c)
                            x[t][c] = (e + c < numel) ? Elem<T>::load1(s_ptr[t + 1], e + c) : 0.0f;
                }
            } else {
                ElemPair<T>::load(s_ptr[0], e, full, numel, b);
#pragma unroll
                for (int t = 0; t < NT; ++t) ElemPair<T>::load(s_ptr[t + 1], e, full, numel, x[t]);
            }
            if (has_mask) pword = __ldg(packed + (e >> 5));
            k3_step<T, NT, FP16B, DIAG, NOISE, VEC>(b, x, pword, e, numel, r, present_bits, a.center, n_f, tail_add, mean_scale,
                                                    sWT, sChatT, sCbar, sG, res, dacc, ns);
        }
        if constexpr (VEC == 4) {
            if (full) stg_stream_f4(outp + e, make_float4(res[0], res[1], res[2], res[3]));
            else {
#pragma unroll
                for (int c = 0; c < VEC; ++c)
                    if (e + c < numel) outp[e + c] = res[c];
            }
        } else {
            if (full) stg_stream_f2(outp + e, make_float2(res[0], res[1]));
            else if (e < numel) outp[e] = res[0];
        }
    }

    if (DIAG) {
        // CTA reduction of the 5*NT diagnostic rows: sums for rows < 4*NT, max for the rest
        constexpr int NR = kDiagRows * NT;
        constexpr int kRows = 16;
        __shared__ float red[kRows][kBlock + 1];
        float* dout = a.diag + (int64_t)tile * NR;
#pragma unroll
        for (int r0 = 0; r0 < NR; r0 += kRows) {
#pragma unroll
            for (int rr = 0; rr < kRows; ++rr)
                if (r0 + rr < NR) red[rr][tid] = dacc[r0 + rr];
            __syncthreads();
#pragma unroll
            for (int q = 0; q < kRows / (kBlock / 32); ++q) {
                const int rr = warp * (kRows / (kBlock / 32)) + q;
                if (r0 + rr < NR) {
                    const bool is_max = (r0 + rr) >= 3 * NT;
                    float s = 0.0f;
#pragma unroll
                    for (int c = 0; c < kBlock / 32; ++c) {
                        const float v = red[rr][lane + 32 * c];
                        s = is_max ? fmaxf(s, v) : s + v;
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        const float v = __shfl_xor_sync(0xffffffffu, s, o);
                        s = is_max ? fmaxf(s, v) : s + v;
                    }
                    if (lane == 0) dout[r0 + rr] = s;
                }
            }
            __syncthreads();
        }
    }
}

#if SVDQ_DTYPE == 0
// ---- diagnostics finalisation: per (parameter, task) reduce the tile partials ------------------

__global__ void __launch_bounds__(32) k3_diag_finalize(const K3DiagArgs a) {
    const int p = blockIdx.x, t = threadIdx.x;
    if (t >= a.nt) return;
    const int NT = a.nt, NR = kDiagRows * NT;
    double se = 0.0, sa = 0.0, sr = 0.0;
    float mx = 0.0f;
    for (int64_t tl = a.tile_begin[p]; tl < a.tile_begin[p + 1]; ++tl) {
        const float* d = a.diag + tl * NR;
        se += (double)d[0 * NT + t]; sa += (double)d[1 * NT + t]; sr += (double)d[2 * NT + t];
        mx = fmaxf(mx, d[3 * NT + t]);
    }
    // ||orig_t||^2 over the masked rows is the diagonal of the (uncentred) masked Gram K1 already reduced in fp64
    const double so = a.gram_masked[((int64_t)p * NT + t) * NT + t];
    double* o = a.out + ((int64_t)p * NT + t) * 6;
    const bool solved = a.info[(int64_t)p * 8] == kSolved;
    const double dm = (double)a.dm[p];
    const double on = sqrt(so), en = sqrt(se);
    o[0] = solved ? en : 0.0;
    o[1] = solved ? (on > 1e-10 ? en / on : 0.0) : 0.0;
    o[2] = solved ? ((se != se) ? se : (double)mx) : 0.0;     // NaN anywhere -> NaN max (torch.max)
    o[3] = solved && dm > 0 ? sa / dm : 0.0;
    o[4] = solved ? on : 0.0;
    o[5] = solved ? sqrt(sr) : 0.0;
}

#endif  // SVDQ_DTYPE == 0

// ---- host-side dispatch ---------------------------------------------------------------------------
template <typename T, int NT>
static cudaError_t launch_nt(const K3Args& a, int n_tiles, bool fp16b, bool diag, cudaStream_t st) {
    if (n_tiles <= 0) return cudaSuccess;
    const bool noise = a.info_n != nullptr;
#define SVDQ_GO(F, D)                                                                          \
    do {                                                                                       \
        if (noise) k3_reconstruct_merge<T, NT, F, D, true><<<n_tiles, kBlock, 0, st>>>(a);     \
        else       k3_reconstruct_merge<T, NT, F, D, false><<<n_tiles, kBlock, 0, st>>>(a);    \
    } while (0)
    if (diag) {
        if (fp16b) SVDQ_GO(true, true);
        else       SVDQ_GO(false, true);
    } else {
        if (fp16b) SVDQ_GO(true, false);
        else       SVDQ_GO(false, false);
    }
#undef SVDQ_GO
    return cudaGetLastError();
}

template <>
cudaError_t k3_launch_dtype<SVDQ_DTYPE>(int nt, const K3Args& a, int n_tiles, bool fp16b, bool diag, cudaStream_t st) {
    using T = DTypeOf<SVDQ_DTYPE>::type;
    switch (nt) {
#define SVDQ_CASE(N) case N: return launch_nt<T, N>(a, n_tiles, fp16b, diag, st);
        SVDQ_CASE(1) SVDQ_CASE(2) SVDQ_CASE(3) SVDQ_CASE(4) SVDQ_CASE(5) SVDQ_CASE(6) SVDQ_CASE(7) SVDQ_CASE(8)
        SVDQ_CASE(9) SVDQ_CASE(10) SVDQ_CASE(11) SVDQ_CASE(12) SVDQ_CASE(13) SVDQ_CASE(14) SVDQ_CASE(15) SVDQ_CASE(16)
#undef SVDQ_CASE
        default: return cudaErrorInvalidValue;
    }
}

#if SVDQ_DTYPE == 0
cudaError_t k3_diag_launch(const K3DiagArgs& a, int n_params, cudaStream_t st) {
    if (n_params <= 0) return cudaSuccess;
    k3_diag_finalize<<<n_params, 32, 0, st>>>(a);
    return cudaGetLastError();
}
#endif

}  // namespace svdq
