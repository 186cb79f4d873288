// Per-step arithmetic of pass 2 (K3), shared by the direct-load kernel (k3_reconstruct_merge.cu) and
// the staged persistent kernel (k3s_reconstruct_merge_staged.cu): given one thread's four base values and
// N x 4 fine-tuned values it forms the task vectors, the mean across tasks, the basis rows
// u_d = (tau_d - mean_d) W (fp16 round trip when the bases are stored in fp16), the weighted
// reconstruction, and optionally the per-task diagnostic reductions.
#pragma once
#include "svdq_kernels.h"

namespace svdq {

constexpr int kDiagRows = 5;      // sum e^2, sum |e|, sum rec^2, sum orig^2, max |e|

template <int NT> struct K3Shared {
    static constexpr int NTP = (NT + 3) & ~3;
};

// x: in = fine-tuned values, scratch afterwards.  res: out = merged values.
template <typename T, int NT, bool FP16B, bool DIAG>
__device__ __forceinline__ void k3_step(const float (&b)[kVec], float (&x)[NT][kVec], const uint32_t pword,
                                        const int64_t e, const int64_t numel, const int r, const uint32_t present_bits,
                                        const int center, const float n_f, const float tail_add,
                                        const float (*sWT)[(NT + 3) & ~3], const float (*sChatT)[(NT + 3) & ~3],
                                        const float* sCbar, const float* sG, float (&res)[kVec],
                                        float (&dacc)[DIAG ? kDiagRows * NT : 1]) {
    struct { int center; } a{center};
    float mean[kVec];
#pragma unroll
    for (int c = 0; c < kVec; ++c) mean[c] = 0.0f;
#pragma unroll
    for (int t = 0; t < NT; ++t)
#pragma unroll
        for (int c = 0; c < kVec; ++c) {
            x[t][c] = Elem<T>::sub(x[t][c], b[c]);
            mean[c] += x[t][c];
        }
    const uint32_t bits = (pword >> (int)(e & 31)) & 0xFu;

    float orig[DIAG ? NT : 1][kVec];
    if (DIAG) {
#pragma unroll
        for (int t = 0; t < NT; ++t)
#pragma unroll
            for (int c = 0; c < kVec; ++c) orig[t][c] = x[t][c];
    }
    // mean over the active tasks = sum / n (basis.py:109); n a power of two: multiply by 1/n, exactly the same
    const int n_i = (int)n_f;
    const bool pow2 = (n_i & (n_i - 1)) == 0;
    const float inv_n = __fdiv_rn(1.0f, n_f);
#pragma unroll
    for (int c = 0; c < kVec; ++c) mean[c] = a.center ? (pow2 ? mean[c] * inv_n : __fdiv_rn(mean[c], n_f)) : 0.0f;
    if (present_bits == ((1u << NT) - 1u)) {
#pragma unroll
        for (int t = 0; t < NT; ++t)
#pragma unroll
            for (int c = 0; c < kVec; ++c) x[t][c] = x[t][c] - mean[c];
    } else {
#pragma unroll
        for (int t = 0; t < NT; ++t)
#pragma unroll
            for (int c = 0; c < kVec; ++c) x[t][c] = ((present_bits >> t) & 1u) ? x[t][c] - mean[c] : 0.0f;
    }

    float acc[kVec];
#pragma unroll
    for (int c = 0; c < kVec; ++c) acc[c] = 0.0f;
    float rec[DIAG ? NT : 1][kVec];
    if (DIAG) {
#pragma unroll
        for (int t = 0; t < NT; ++t)
#pragma unroll
            for (int c = 0; c < kVec; ++c) rec[t][c] = 0.0f;
    }
    if (FP16B || DIAG) {
#pragma unroll
        for (int j = 0; j < NT; ++j) {               // r <= NT columns; unrolled, uniform early exit
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
            const float cb = sCbar[j];
#pragma unroll
            for (int c = 0; c < kVec; ++c) {
                if (FP16B) u[c] = round_fp16(u[c]);
                acc[c] = fmaf(u[c], cb, acc[c]);
            }
            if (DIAG) {
#pragma unroll
                for (int t = 0; t < NT; ++t) {
                    const float ch = sChatT[j][t];
#pragma unroll
                    for (int c = 0; c < kVec; ++c) rec[t][c] = fmaf(u[c], ch, rec[t][c]);
                }
            }
        }
    } else {
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            const float g = sG[t];
#pragma unroll
            for (int c = 0; c < kVec; ++c) acc[c] = fmaf(x[t][c], g, acc[c]);
        }
    }
#pragma unroll
    for (int c = 0; c < kVec; ++c) {
        const bool m = (bits >> c) & 1u;
        const float val = (acc[c] + mean[c]) + tail_add;
        res[c] = b[c] + (m ? val : 0.0f);
    }
    if (DIAG) {
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            if (!((present_bits >> t) & 1u)) continue;
#pragma unroll
            for (int c = 0; c < kVec; ++c) {
                const bool m = ((bits >> c) & 1u) && (e + c < numel);
                if (m) {
                    const float er = orig[t][c] - rec[t][c];
                    dacc[0 * NT + t] = fmaf(er, er, dacc[0 * NT + t]);
                    dacc[1 * NT + t] += fabsf(er);
                    dacc[2 * NT + t] = fmaf(rec[t][c], rec[t][c], dacc[2 * NT + t]);
                    dacc[3 * NT + t] = fmaf(orig[t][c], orig[t][c], dacc[3 * NT + t]);
                    dacc[4 * NT + t] = fmaxf(dacc[4 * NT + t], fabsf(er));
                }
            }
        }
    }
}

}  // namespace svdq
