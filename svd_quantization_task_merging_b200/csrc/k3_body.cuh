// Per-step arithmetic of pass 2 (K3), shared by the direct-load kernel (k3_reconstruct_merge.cu) and
// the staged persistent kernel (k3s_reconstruct_merge_staged.cu): given one thread's four base values and
// N x 4 fine-tuned values it forms the task vectors, the mean across tasks, the basis rows
// u_d = (tau_d - mean_d) W (fp16 round trip when the bases are stored in fp16), the weighted
// reconstruction, and optionally the per-task diagnostic reductions.
#pragma once
#include "svdq_kernels.h"

namespace svdq {

// Block-uniform: does the combined mask keep fewer than 55 % of the tile's elements?  (the compacting pass 2 with
// fused diagnostics wins below that, the non-compacting one above: 43 % kept 5.8 vs 6.46 ms, 64 % 6.58 vs 6.45,
// 94 % 7.16 vs 6.45, no mask 7.11 vs 6.09 ms per ViT-L-14 x 8 merge.)  All threads of the CTA must call it.
__device__ __forceinline__ bool k3_tile_is_sparse(const uint32_t* packed, int64_t start, int64_t stop, int64_t numel) {
    __shared__ uint32_t s_kept[kBlock / 32];
    if (packed == nullptr) return false;
    uint32_t c = 0;
    for (int64_t w = (start >> 5) + threadIdx.x; w * 32 < stop; w += kBlock) {
        uint32_t bits = __ldg(packed + w);
        const int64_t left = numel - w * 32;
        if (left < 32) bits &= (1u << (int)left) - 1u;
        c += __popc(bits);
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0) s_kept[threadIdx.x >> 5] = c;
    __syncthreads();
    uint32_t kept = 0;
#pragma unroll
    for (int w = 0; w < kBlock / 32; ++w) kept += s_kept[w];
    __syncthreads();
    return (int64_t)kept * 20 < (stop - start) * 11;
}

constexpr int kDiagRows = 4;      // sum e^2, sum |e|, sum rec^2, max |e|  (sum orig^2 = diagonal of K1's masked Gram)

// Second coefficient set of a parameter: the basis of the elements OUTSIDE the combined mask (svd_include_noise,
// src/svd_hybrid/basis.py:455-466); its reconstruction is scaled by svd_noise_shrink (merge.py:257-284) and
// scattered to the unmasked positions (mask_loader.py:757-760).
template <int NT> struct K3NoiseSet {
    const float (*sWT)[(NT + 3) & ~3] = nullptr;
    const float* sCbar = nullptr;
    const float* sG = nullptr;
    int r = 0;
    float tail_add = 0.0f;
    float mean_scale = 1.0f;
    float shrink = 1.0f;
    bool on = false;
};

// Optional observer of pass 2's intermediate values (VEC = 2 only): the basis-row entries u_dj of the thread's two
// elements, after the fp16 round trip when the bases are stored in fp16, and the per-element mean across tasks.  The
// compacting diagnostics kernel uses it to write the artifact bases in the same pass (K3NoSink: nothing).
struct K3NoSink {
    static constexpr bool on = false;
    __device__ __forceinline__ void col(int, float2) const {}
    __device__ __forceinline__ void mean2(float, float) const {}
};

// x: in = fine-tuned values, scratch afterwards.  res: out = merged values.  VEC = elements per thread and call
// (4; 2 for the kernel with fused diagnostics, whose 4*NT running reductions otherwise cap it at 8 warps per SM).
template <typename T, int NT, bool FP16B, bool DIAG, bool NOISE = false, int VEC = kVec, class SINK = K3NoSink>
__device__ __forceinline__ void k3_step(const float (&b)[VEC], float (&x)[NT][VEC], const uint32_t pword,
                                        const int64_t e, const int64_t numel, const int r, const uint32_t present_bits,
                                        const int center, const float n_f, const float tail_add,
                                        const float mean_scale,
                                        const float (*sWT)[(NT + 3) & ~3], const float (*sChatT)[(NT + 3) & ~3],
                                        const float* sCbar, const float* sG, float (&res)[VEC],
                                        float (&dacc)[DIAG ? kDiagRows * NT : 1],
                                        const K3NoiseSet<NT>& ns = K3NoiseSet<NT>(), const SINK& sink = SINK()) {
    static_assert(!SINK::on || VEC == 2, "the sink sees two elements per call");
    struct { int center; } a{center};
    constexpr int kH = VEC / 2;
    float mean[VEC];
    {
        // sum over the tasks in task order, two elements per packed add (same rounding as the scalar adds)
        float2 m2[kH];
#pragma unroll
        for (int h = 0; h < kH; ++h) m2[h] = make_float2(0.0f, 0.0f);
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            Elem<T>::template subv<VEC>(x[t], b, x[t]);
#pragma unroll
            for (int h = 0; h < kH; ++h) m2[h] = __fadd2_rn(m2[h], make_float2(x[t][2 * h], x[t][2 * h + 1]));
        }
#pragma unroll
        for (int h = 0; h < kH; ++h) { mean[2 * h] = m2[h].x; mean[2 * h + 1] = m2[h].y; }
    }
    const uint32_t bits = (pword >> (int)(e & 31)) & ((1u << VEC) - 1u);

    float orig[DIAG ? NT : 1][VEC];
    if (DIAG) {
#pragma unroll
        for (int t = 0; t < NT; ++t)
#pragma unroll
            for (int c = 0; c < VEC; ++c) orig[t][c] = x[t][c];
    }
    // mean over the active tasks = sum / n (basis.py:109); n a power of two: multiply by 1/n, exactly the same
    const int n_i = (int)n_f;
    const bool pow2 = (n_i & (n_i - 1)) == 0;
    const float inv_n = __fdiv_rn(1.0f, n_f);
#pragma unroll
    for (int c = 0; c < VEC; ++c) mean[c] = a.center ? (pow2 ? mean[c] * inv_n : __fdiv_rn(mean[c], n_f)) : 0.0f;
    if constexpr (SINK::on) sink.mean2(mean[0], mean[1]);
    const bool all_present = present_bits == ((1u << NT) - 1u);
    if (all_present) {
        // x - mean as mean * -1 + x: one packed instruction per element pair, same single rounding
        const float2 neg1 = make_float2(-1.0f, -1.0f);
#pragma unroll
        for (int t = 0; t < NT; ++t)
#pragma unroll
            for (int h = 0; h < VEC / 2; ++h) {
                const float2 d = __ffma2_rn(make_float2(mean[2 * h], mean[2 * h + 1]), neg1, make_float2(x[t][2 * h], x[t][2 * h + 1]));
                x[t][2 * h] = d.x; x[t][2 * h + 1] = d.y;
            }
    } else {
#pragma unroll
        for (int t = 0; t < NT; ++t)
#pragma unroll
            for (int c = 0; c < VEC; ++c) x[t][c] = ((present_bits >> t) & 1u) ? x[t][c] - mean[c] : 0.0f;
    }

    // The two tall-skinny contractions run as packed 2-wide FMAs (fma.rn.f32x2): elements (0,1) and (2,3)
    // of the thread share an instruction; the fp16 round trip of the basis row uses the packed converts.
    float2 x2[NT][kH];
#pragma unroll
    for (int t = 0; t < NT; ++t)
#pragma unroll
        for (int h = 0; h < kH; ++h) x2[t][h] = make_float2(x[t][2 * h], x[t][2 * h + 1]);
    float2 acc2[kH];
#pragma unroll
    for (int h = 0; h < kH; ++h) acc2[h] = make_float2(0.0f, 0.0f);
    float2 rec2[DIAG ? NT : 1][kH];
    if (DIAG) {
#pragma unroll
        for (int t = 0; t < NT; ++t)
#pragma unroll
            for (int h = 0; h < kH; ++h) rec2[t][h] = make_float2(0.0f, 0.0f);
    }
    if (FP16B || DIAG || SINK::on) {
        // r <= NT columns, unrolled
        auto column = [&](const int j) {
            float2 u2[kH];
#pragma unroll
            for (int h = 0; h < kH; ++h) u2[h] = make_float2(0.0f, 0.0f);
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                const float w = sWT[j][t];
                const float2 w2 = make_float2(w, w);
#pragma unroll
                for (int h = 0; h < kH; ++h) u2[h] = __ffma2_rn(x2[t][h], w2, u2[h]);
            }
            const float cb = sCbar[j];
            const float2 cb2 = make_float2(cb, cb);
#pragma unroll
            for (int h = 0; h < kH; ++h) {
                if (FP16B) u2[h] = __half22float2(__float22half2_rn(u2[h]));
                acc2[h] = __ffma2_rn(u2[h], cb2, acc2[h]);
            }
            if constexpr (SINK::on) sink.col(j, u2[0]);
            if (DIAG) {
#pragma unroll
                for (int t = 0; t < NT; ++t) {
                    const float ch = sChatT[j][t];
                    const float2 ch2 = make_float2(ch, ch);
#pragma unroll
                    for (int h = 0; h < kH; ++h) rec2[t][h] = __ffma2_rn(u2[h], ch2, rec2[t][h]);
                }
            }
        };
        if (DIAG && NT > 1 && r >= NT - 1) {
            // the common case (full rank, or full rank minus the centring direction): straight-line code, one
            // uniform test instead of an exit test per column
#pragma unroll
            for (int j = 0; j < NT - 1; ++j) column(j);
            if (r == NT) column(NT - 1);
        } else {
#pragma unroll
            for (int j = 0; j < NT; ++j) {            // uniform early exit
                if (j >= r) break;
                column(j);
            }
        }
    } else {
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            const float g = sG[t];
            const float2 g2 = make_float2(g, g);
#pragma unroll
            for (int h = 0; h < kH; ++h) acc2[h] = __ffma2_rn(x2[t][h], g2, acc2[h]);
        }
    }
    float acc[VEC];
#pragma unroll
    for (int h = 0; h < kH; ++h) { acc[2 * h] = acc2[h].x; acc[2 * h + 1] = acc2[h].y; }
    // noise region: the same centred task vectors contracted with the second (unmasked-rows) coefficient set;
    // every element belongs to exactly one region, the select happens at the store
    float accn[NOISE ? VEC : 1];
    if (NOISE) {
        float2 an2[kH];
#pragma unroll
        for (int h = 0; h < kH; ++h) an2[h] = make_float2(0.0f, 0.0f);
        if (ns.on) {
            if (FP16B) {
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    if (j >= ns.r) break;
                    float2 u2[kH];
#pragma unroll
                    for (int h = 0; h < kH; ++h) u2[h] = make_float2(0.0f, 0.0f);
#pragma unroll
                    for (int t = 0; t < NT; ++t) {
                        const float w = ns.sWT[j][t];
                        const float2 w2 = make_float2(w, w);
#pragma unroll
                        for (int h = 0; h < kH; ++h) u2[h] = __ffma2_rn(x2[t][h], w2, u2[h]);
                    }
                    const float cb = ns.sCbar[j];
                    const float2 cb2 = make_float2(cb, cb);
#pragma unroll
                    for (int h = 0; h < kH; ++h) {
                        u2[h] = __half22float2(__float22half2_rn(u2[h]));
                        an2[h] = __ffma2_rn(u2[h], cb2, an2[h]);
                    }
                }
            } else {
#pragma unroll
                for (int t = 0; t < NT; ++t) {
                    const float g = ns.sG[t];
                    const float2 g2 = make_float2(g, g);
#pragma unroll
                    for (int h = 0; h < kH; ++h) an2[h] = __ffma2_rn(x2[t][h], g2, an2[h]);
                }
            }
        }
#pragma unroll
        for (int h = 0; h < kH; ++h) { accn[2 * h] = an2[h].x; accn[2 * h + 1] = an2[h].y; }
    }
    float rec[DIAG ? NT : 1][VEC];
    if (DIAG) {
#pragma unroll
        for (int t = 0; t < NT; ++t)
#pragma unroll
            for (int h = 0; h < kH; ++h) { rec[t][2 * h] = rec2[t][h].x; rec[t][2 * h + 1] = rec2[t][h].y; }
    }
#pragma unroll
    for (int c = 0; c < VEC; ++c) {
        const bool m = (bits >> c) & 1u;
        // mean_scale is 1 except under cluster weighting when a whole cluster lacks the parameter (that cluster
        // contributes zeros, mean included: merge.py:289-290); fma(mean, 1, acc) rounds exactly like acc + mean
        const float val = fmaf(mean[c], mean_scale, acc[c]) + tail_add;
        float other = 0.0f;
        if (NOISE) other = ns.on ? __fmul_rn(fmaf(mean[c], ns.mean_scale, accn[c]) + ns.tail_add, ns.shrink) : 0.0f;
        res[c] = b[c] + (m ? val : other);
    }
    if (DIAG) {
        // elements that count: inside the combined mask and inside the tensor (one mask word for all tasks)
        const int64_t left = numel - e;
        const uint32_t mb = bits & (left >= (int64_t)VEC ? ((1u << VEC) - 1u) : ((1u << (int)(left > 0 ? left : 0)) - 1u));
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            if (!all_present && !((present_bits >> t) & 1u)) continue;
#pragma unroll
            for (int c = 0; c < VEC; ++c) {
                if ((mb >> c) & 1u) {
                    const float er = orig[t][c] - rec[t][c];
                    dacc[0 * NT + t] = fmaf(er, er, dacc[0 * NT + t]);
                    dacc[1 * NT + t] += fabsf(er);
                    dacc[2 * NT + t] = fmaf(rec[t][c], rec[t][c], dacc[2 * NT + t]);
                    dacc[3 * NT + t] = fmaxf(dacc[3 * NT + t], fabsf(er));
                }
            }
        }
    }
}

}  // namespace svdq
