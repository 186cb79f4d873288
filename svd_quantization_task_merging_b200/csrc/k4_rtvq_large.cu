// K4 — multi-stage residual quantisation (RTVQ) and the root quantization_utils quantisers on
// arbitrarily long tensors.
//
// Replaces (paths relative to /root/reference):
//   asymmetric_quantization / _dequantization      src/svd_hybrid/rtvq.py:4-36, quantization_utils.py:76-99,137-172
//   multistage_residual_quantization / _deq.       src/svd_hybrid/rtvq.py:39-103
//   absmax_quantization                            quantization_utils.py:60-73
//
// The residual is never written to memory: pass q re-reads x and REPLAYS stages 0..q-2 with
// their (already final) scale / zero-point to rebuild the residual in registers bit-exactly,
// quantises stage q-1, and reduces min / max / sum-of-squares of the new residual for stage q
// in the same pass.  S stages cost S+1 reads of x and S code writes; all scalars stay on the
// device (no host synchronisation between stages).  Bound: HBM.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "svdq_kernels.h"

namespace svdq {

namespace cg = cooperative_groups;



struct K4PassArgs {
    const float* x;
    int64_t n;
    int bits;
    int pass;               // q: replay stages 0..q-2, quantise stage q-1 (if q >= 1), stats for stage q
    int stages;             // S: no stats when pass == S
    const float* scale;     // [S] device scalars, entries < pass valid
    const float* zp;        // [S]
    void* codes;            // [S][codes_ld] uint8 or int16
    int64_t codes_ld;
    uint32_t* packed;       // optional [S][packed_ld] words: codes bit-packed, `bits` per code, little-endian
    int64_t packed_ld;      //          (bits in {1, 2, 4, 8}); the reference layout stays one uint8 per code
    K4Stats* part;          // [gridDim.x]
};

__device__ __forceinline__ void stats_update(float v, float& lo, float& hi, int& nan, double& ss) {
    nan |= (v != v);
    lo = fminf(lo, v);      // fminf/fmaxf drop NaN; the nan flag restores torch.min/max propagation
    hi = fmaxf(hi, v);
    ss += (double)v * (double)v;
}

// The dequantised value of a code depends only on (stage, code): 2^bits values per stage.  They are
// tabulated once per CTA with the exactly-rounded IEEE expression (q - zp) / scale, so the per-element
// replay costs a shared-memory lookup instead of a division and stays bit-exact.
template <typename CodeT>
__global__ void __launch_bounds__(kBlock) k4_rtvq_pass(const K4PassArgs a) {
    __shared__ QuantScalars s_q[kCoreMaxStages];
    __shared__ float s_lut[sizeof(CodeT) == 1 ? kCoreMaxStages : 1][sizeof(CodeT) == 1 ? 256 : 1];
    const int tid = threadIdx.x;
    const int q = a.pass;
    if (tid < q && tid < kCoreMaxStages) { s_q[tid].scale = a.scale[tid]; s_q[tid].zp = a.zp[tid]; }
    __syncthreads();
    const bool want_stats = q < a.stages;
    if (sizeof(CodeT) == 1) {
        const int levels = 1 << a.bits;
        for (int i = tid; i < q * levels; i += kBlock) {
            const int s = i / levels, c = i % levels;
            s_lut[s][c] = asym_decode(c, s_q[s]);
        }
        __syncthreads();
    }
    float lo = __int_as_float(0x7f800000), hi = __int_as_float(0xff800000);
    int nan = 0;
    double ss = 0.0;
    CodeT* crow = q >= 1 ? reinterpret_cast<CodeT*>(a.codes) + (int64_t)(q - 1) * a.codes_ld : nullptr;
    const float qmax = (float)((1 << a.bits) - 1);

    const int64_t nvec = (a.n + kVec - 1) / kVec;
    constexpr int kUnroll = 2;                      // two 128-bit loads in flight per thread
    const int64_t stride = (int64_t)gridDim.x * kBlock;
    // uniform trip count for every thread of the grid (the packed-code path uses full-warp shuffles)
    const int64_t n_iter = (nvec + stride * kUnroll - 1) / (stride * kUnroll);
    for (int64_t it = 0; it < n_iter; ++it) {
        const int64_t v0 = (int64_t)blockIdx.x * kBlock + tid + it * stride * kUnroll;
        float r[kUnroll][kVec];
        bool fullv[kUnroll], act[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const int64_t e = (v0 + u * stride) * kVec;
            act[u] = v0 + u * stride < nvec;
            fullv[u] = act[u] && e + kVec <= a.n;
            if (fullv[u]) { const float4 t = ldg_stream_f4(a.x + e); r[u][0] = t.x; r[u][1] = t.y; r[u][2] = t.z; r[u][3] = t.w; }
            else {
#pragma unroll
                for (int c = 0; c < kVec; ++c) r[u][c] = (act[u] && e + c < a.n) ? __ldg(a.x + e + c) : 0.0f;
            }
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            if (!act[u]) continue;
            const int64_t e = (v0 + u * stride) * kVec;
            int code[kVec] = {0, 0, 0, 0};
            for (int s = 0; s < q; ++s) {
                const QuantScalars qs = s_q[s];
                const bool need_res = (s < q - 1) || want_stats;
#pragma unroll
                for (int c = 0; c < kVec; ++c) {
                    float t = rintf(f_add(f_mul(qs.scale, r[u][c]), qs.zp));
                    t = (t != t) ? 0.0f : fminf(fmaxf(t, 0.0f), qmax);       // NaN -> code 0 (torch .to(uint8))
                    code[c] = (int)t;
                    if (need_res) {
                        const float d = sizeof(CodeT) == 1 ? s_lut[s][code[c]] : asym_decode(code[c], qs);
                        r[u][c] = f_sub(r[u][c], d);
                    }
                }
            }
            if (q >= 1) {
                if (fullv[u]) {
                    if (sizeof(CodeT) == 1) {
                        const uint32_t w = (uint32_t)code[0] | ((uint32_t)code[1] << 8) | ((uint32_t)code[2] << 16) |
                                           ((uint32_t)code[3] << 24);
                        *reinterpret_cast<uint32_t*>(crow + e) = w;
                    } else {
                        uint2 w;
                        w.x = ((uint32_t)code[0] & 0xffffu) | ((uint32_t)code[1] << 16);
                        w.y = ((uint32_t)code[2] & 0xffffu) | ((uint32_t)code[3] << 16);
                        *reinterpret_cast<uint2*>(crow + e) = w;
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < kVec; ++c)
                        if (e + c < a.n) crow[e + c] = (CodeT)code[c];
                }
            }
            if (want_stats) {
#pragma unroll
                for (int c = 0; c < kVec; ++c)
                    if (e + c < a.n) stats_update(r[u][c], lo, hi, nan, ss);
            }
        }
        // ---- optional bit-packed copy of the codes, assembled with warp shuffles: a thread holds
        //      4 * bits bits, 32 / (4 * bits) neighbouring lanes share one 32-bit word ------------------
        if (a.packed != nullptr && q >= 1 && sizeof(CodeT) == 1) {
            const int lane = tid & 31;
            const int per_thread = kVec * a.bits;                 // 4, 8, 16 or 32 bits
            const int lanes_per_word = 32 / per_thread;           // 8, 4, 2 or 1
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const int64_t v = v0 + u * stride;                // vector index of this thread (lane-contiguous)
                uint32_t w = 0;
                if (act[u]) {
                    const int64_t e = v * kVec;
                    const uint8_t* cr = reinterpret_cast<const uint8_t*>(crow);
#pragma unroll
                    for (int c = 0; c < kVec; ++c)
                        if (e + c < a.n) w |= (uint32_t)cr[e + c] << (c * a.bits);
                    w <<= (lane % lanes_per_word) * per_thread;
                }
                for (int o = 1; o < lanes_per_word; o <<= 1) w |= __shfl_xor_sync(0xffffffffu, w, o);
                if (act[u] && (lane % lanes_per_word) == 0)
                    a.packed[(int64_t)(q - 1) * a.packed_ld + v / lanes_per_word] = w;
            }
        }
    }
    if (!want_stats) return;
    // CTA reduction in a fixed order
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        nan |= __shfl_xor_sync(0xffffffffu, nan, o);
        ss += __shfl_xor_sync(0xffffffffu, ss, o);
    }
    __shared__ K4Stats s_w[kBlock / 32];
    if ((tid & 31) == 0) { s_w[tid >> 5].lo = lo; s_w[tid >> 5].hi = hi; s_w[tid >> 5].nan = nan; s_w[tid >> 5].sumsq = ss; }
    __syncthreads();
    if (tid == 0) {
        K4Stats t = s_w[0];
        for (int w = 1; w < kBlock / 32; ++w) {
            t.lo = fminf(t.lo, s_w[w].lo); t.hi = fmaxf(t.hi, s_w[w].hi); t.nan |= s_w[w].nan; t.sumsq += s_w[w].sumsq;
        }
        t.pad = 0;
        a.part[blockIdx.x] = t;
    }
}

// one warp: reduce the per-CTA stats in a fixed order, emit scale / zero-point / residual norm of `stage`
__global__ void __launch_bounds__(32) k4_finalize(const K4Stats* part, int n_part, int bits, int stage,
                                                  float* scale, float* zp, float* resnorm) {
    const int lane = threadIdx.x;
    float lo = __int_as_float(0x7f800000), hi = __int_as_float(0xff800000);
    int nan = 0;
    double ss = 0.0;
    for (int i = lane; i < n_part; i += 32) {
        lo = fminf(lo, part[i].lo); hi = fmaxf(hi, part[i].hi); nan |= part[i].nan; ss += part[i].sumsq;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        nan |= __shfl_xor_sync(0xffffffffu, nan, o);
        ss += __shfl_xor_sync(0xffffffffu, ss, o);
    }
    if (lane != 0) return;
    if (nan) { lo = __int_as_float(0x7fc00000); hi = lo; }
    const QuantScalars q = asym_scalars(lo, hi, bits);
    scale[stage] = q.scale;
    zp[stage] = q.zp;
    resnorm[stage] = (float)sqrt(ss);
}

// sum of the stage dequantisations, left to right from zero (rtvq.py:91-101)
template <typename CodeT>
__global__ void __launch_bounds__(kBlock) k4_dequant(const void* codes, int64_t codes_ld, int stages, int64_t n,
                                                     const float* scale, const float* zp, float* out) {
    __shared__ QuantScalars s_q[kCoreMaxStages];
    if (threadIdx.x < stages) { s_q[threadIdx.x].scale = scale[threadIdx.x]; s_q[threadIdx.x].zp = zp[threadIdx.x]; }
    __syncthreads();
    const CodeT* c = reinterpret_cast<const CodeT*>(codes);
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock) {
        float acc = 0.0f;
        for (int s = 0; s < stages; ++s) acc = f_add(acc, asym_decode((int)c[(int64_t)s * codes_ld + i], s_q[s]));
        out[i] = acc;
    }
}

// ---- absmax (quantization_utils.py:60-73): s = (2^(b-1)-1) / max|x| ; q = round(s x), no clamp --
__global__ void __launch_bounds__(kBlock) k4_absmax_stats(const float* x, int64_t n, K4Stats* part) {
    float hi = 0.0f;
    int nan = 0;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock) {
        const float v = fabsf(__ldg(x + i));
        nan |= (v != v);
        hi = fmaxf(hi, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        nan |= __shfl_xor_sync(0xffffffffu, nan, o);
    }
    __shared__ float s_hi[kBlock / 32];
    __shared__ int s_nan[kBlock / 32];
    if ((threadIdx.x & 31) == 0) { s_hi[threadIdx.x >> 5] = hi; s_nan[threadIdx.x >> 5] = nan; }
    __syncthreads();
    if (threadIdx.x == 0) {
        K4Stats t; t.lo = 0.0f; t.hi = s_hi[0]; t.nan = s_nan[0]; t.pad = 0; t.sumsq = 0.0;
        for (int w = 1; w < kBlock / 32; ++w) { t.hi = fmaxf(t.hi, s_hi[w]); t.nan |= s_nan[w]; }
        part[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(32) k4_absmax_finalize(const K4Stats* part, int n_part, int bits, float* scale) {
    if (threadIdx.x != 0) return;
    float hi = part[0].hi;
    int nan = part[0].nan;
    for (int i = 1; i < n_part; ++i) { hi = fmaxf(hi, part[i].hi); nan |= part[i].nan; }
    if (nan) hi = __int_as_float(0x7fc00000);
    const float recip = f_div(1.0f, hi);
    scale[0] = f_mul(recip, (float)((1 << (bits - 1)) - 1));
}

template <typename CodeT>
__global__ void __launch_bounds__(kBlock) k4_absmax_quant(const float* x, int64_t n, const float* scale, CodeT* q) {
    const float s = scale[0];
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock) {
        const float t = rintf(f_mul(s, __ldg(x + i)));
        q[i] = (t != t) ? (CodeT)0 : (CodeT)(int)t;
    }
}

// ---- single-kernel path for tensors that fit the register file of the whole GPU (<= ~4.8 M elements) ---------
// One cooperative launch: every thread loads its R elements ONCE and keeps the residual in registers across all
// stages; per stage the CTAs publish min / max / sum-of-squares partials, meet at a grid barrier, every CTA
// reduces the (few hundred) partials in the same fixed order, then quantises, writes the codes and updates the
// residual.  x is read once, the codes are written once per stage, nothing else touches memory -- and the
// S + 1 pass launches + S finalize launches of the multi-pass path (launch-latency-bound below ~20 M elements:
// 74 us for one stage of 4.19 M elements) become one launch with S grid barriers.  Bit-identical codes / scale /
// zero point (min / max are order-free, the quantiser arithmetic is the same code); the residual norm is summed
// in fp64 in a different order (same value to ~1e-16 relative).
constexpr int kK4fThreads = 512;

template <int R>     // R = elements per thread, a multiple of 4
__global__ void __launch_bounds__(kK4fThreads, 2) k4_rtvq_fused(const K4PassArgs a, float* scale_out, float* zp_out,
                                                                float* resnorm_out) {
    cg::grid_group grid = cg::this_grid();
    __shared__ float s_lut[256];
    __shared__ K4Stats s_w[kK4fThreads / 32];
    __shared__ QuantScalars s_q;
    const int tid = threadIdx.x, lane = tid & 31;
    const int64_t nvec = (a.n + kVec - 1) / kVec;
    const int64_t vstride = (int64_t)gridDim.x * kK4fThreads;
    float r[R];
#pragma unroll
    for (int k = 0; k < R / kVec; ++k) {
        const int64_t v = (int64_t)k * vstride + (int64_t)blockIdx.x * kK4fThreads + tid;
        const int64_t e = v * kVec;
        if (v < nvec && e + kVec <= a.n) {
            const float4 t = ldg_stream_f4(a.x + e);
            r[k * 4 + 0] = t.x; r[k * 4 + 1] = t.y; r[k * 4 + 2] = t.z; r[k * 4 + 3] = t.w;
        } else {
#pragma unroll
            for (int c = 0; c < kVec; ++c) r[k * 4 + c] = (v < nvec && e + c < a.n) ? __ldg(a.x + e + c) : 0.0f;
        }
    }
    const float qmax = (float)((1 << a.bits) - 1);
    const int levels = 1 << a.bits;
    for (int s = 0; s < a.stages; ++s) {
        // ---- statistics of the residual before stage s ----------------------------------------------------------
        float lo = __int_as_float(0x7f800000), hi = __int_as_float(0xff800000);
        int nan = 0;
        double ss = 0.0;
#pragma unroll
        for (int k = 0; k < R / kVec; ++k) {
            const int64_t e = ((int64_t)k * vstride + (int64_t)blockIdx.x * kK4fThreads + tid) * kVec;
#pragma unroll
            for (int c = 0; c < kVec; ++c)
                if (e + c < a.n) stats_update(r[k * 4 + c], lo, hi, nan, ss);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
            nan |= __shfl_xor_sync(0xffffffffu, nan, o);
            ss += __shfl_xor_sync(0xffffffffu, ss, o);
        }
        if (lane == 0) { s_w[tid >> 5].lo = lo; s_w[tid >> 5].hi = hi; s_w[tid >> 5].nan = nan; s_w[tid >> 5].sumsq = ss; }
        __syncthreads();
        if (tid == 0) {
            K4Stats t = s_w[0];
            for (int w = 1; w < kK4fThreads / 32; ++w) {
                t.lo = fminf(t.lo, s_w[w].lo); t.hi = fmaxf(t.hi, s_w[w].hi); t.nan |= s_w[w].nan; t.sumsq += s_w[w].sumsq;
            }
            t.pad = 0;
            a.part[(int64_t)(s & 1) * gridDim.x + blockIdx.x] = t;      // double-buffered: one barrier per stage
            __threadfence();
        }
        grid.sync();
        // ---- every CTA reduces the partials in the same fixed order: one partial per thread (gridDim.x <= 512) ------
        {
            const K4Stats* part = a.part + (int64_t)(s & 1) * gridDim.x;
            float glo = __int_as_float(0x7f800000), ghi = __int_as_float(0xff800000);
            int gnan = 0;
            double gss = 0.0;
            if (tid < (int)gridDim.x) { const K4Stats t = part[tid]; glo = t.lo; ghi = t.hi; gnan = t.nan; gss = t.sumsq; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                glo = fminf(glo, __shfl_xor_sync(0xffffffffu, glo, o));
                ghi = fmaxf(ghi, __shfl_xor_sync(0xffffffffu, ghi, o));
                gnan |= __shfl_xor_sync(0xffffffffu, gnan, o);
                gss += __shfl_xor_sync(0xffffffffu, gss, o);
            }
            __syncthreads();                       // s_w is free (its readers finished before the grid barrier)
            if (lane == 0) { s_w[tid >> 5].lo = glo; s_w[tid >> 5].hi = ghi; s_w[tid >> 5].nan = gnan; s_w[tid >> 5].sumsq = gss; }
            __syncthreads();
            if (tid == 0) {
                K4Stats t = s_w[0];
                for (int w = 1; w < kK4fThreads / 32; ++w) {
                    t.lo = fminf(t.lo, s_w[w].lo); t.hi = fmaxf(t.hi, s_w[w].hi); t.nan |= s_w[w].nan; t.sumsq += s_w[w].sumsq;
                }
                if (t.nan) { t.lo = __int_as_float(0x7fc00000); t.hi = t.lo; }
                s_q = asym_scalars(t.lo, t.hi, a.bits);
                if (blockIdx.x == 0) { scale_out[s] = s_q.scale; zp_out[s] = s_q.zp; resnorm_out[s] = (float)sqrt(t.sumsq); }
            }
        }
        __syncthreads();
        const QuantScalars qs = s_q;
        for (int i = tid; i < levels; i += kK4fThreads) s_lut[i] = asym_decode(i, qs);
        __syncthreads();
        // ---- quantise stage s, write its codes, update the residual ----------------------------------------------
        uint8_t* crow = reinterpret_cast<uint8_t*>(a.codes) + (int64_t)s * a.codes_ld;
        const bool need_res = s + 1 < a.stages;
#pragma unroll
        for (int k = 0; k < R / kVec; ++k) {
            const int64_t v = (int64_t)k * vstride + (int64_t)blockIdx.x * kK4fThreads + tid;
            const int64_t e = v * kVec;
            int code[kVec];
#pragma unroll
            for (int c = 0; c < kVec; ++c) {
                float t = rintf(f_add(f_mul(qs.scale, r[k * 4 + c]), qs.zp));
                t = (t != t) ? 0.0f : fminf(fmaxf(t, 0.0f), qmax);           // NaN -> code 0 (torch .to(uint8))
                code[c] = (int)t;
                if (need_res) r[k * 4 + c] = f_sub(r[k * 4 + c], s_lut[code[c]]);
            }
            if (v < nvec) {
                if (e + kVec <= a.n) {
                    *reinterpret_cast<uint32_t*>(crow + e) = (uint32_t)code[0] | ((uint32_t)code[1] << 8) |
                                                             ((uint32_t)code[2] << 16) | ((uint32_t)code[3] << 24);
                } else {
#pragma unroll
                    for (int c = 0; c < kVec; ++c)
                        if (e + c < a.n) crow[e + c] = (uint8_t)code[c];
                }
            }
            if (a.packed != nullptr) {           // bit-packed copy: 32 / (4 bits) neighbouring lanes share a word
                const int per_thread = kVec * a.bits, lanes_per_word = 32 / per_thread;
                uint32_t w = 0;
                if (v < nvec) {
#pragma unroll
                    for (int c = 0; c < kVec; ++c)
                        if (e + c < a.n) w |= (uint32_t)code[c] << (c * a.bits);
                    w <<= (lane % lanes_per_word) * per_thread;
                }
                for (int o = 1; o < lanes_per_word; o <<= 1) w |= __shfl_xor_sync(0xffffffffu, w, o);
                if (v < nvec && (lane % lanes_per_word) == 0) a.packed[(int64_t)s * a.packed_ld + v / lanes_per_word] = w;
            }
        }
    }
}

// largest tensor the fused path takes (elements): all CTAs must be co-resident (cooperative launch)
static int64_t k4_fused_capacity(int* grid_max) {
    static int g = -1;
    if (g < 0) {
        int dev = 0, sms = 0, occ = 0;
        g = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess &&
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k4_rtvq_fused<32>, kK4fThreads, 0) == cudaSuccess && occ > 0) {
            int coop = 0;
            cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
            if (coop) g = sms * (occ < 2 ? occ : 2);
        }
        const char* v = getenv("SVDQ_K4_FUSED");
        if (v && atoi(v) == 0) g = 0;                 // A/B switch
    }
    *grid_max = g;
    return (int64_t)g * kK4fThreads * 32;
}

template <int R>
static cudaError_t k4_fused_launch(K4PassArgs a, float* scale, float* zp, float* resnorm, int grid_max, cudaStream_t st) {
    int64_t g = (a.n + (int64_t)kK4fThreads * R - 1) / ((int64_t)kK4fThreads * R);
    if (g < 1) g = 1;
    if (g > grid_max) return cudaErrorInvalidValue;
    void* args[] = {&a, &scale, &zp, &resnorm};
    return cudaLaunchCooperativeKernel((void*)k4_rtvq_fused<R>, dim3((unsigned)g), dim3(kK4fThreads), args, 0, st);
}

static int grid_for(int64_t n_threads_needed) {
    int64_t g = (n_threads_needed + kBlock - 1) / kBlock;
    if (g < 1) g = 1;
    if (g > kK4MaxGrid) g = kK4MaxGrid;
    return (int)g;
}

// Runs the whole S-stage quantisation of x[0..n) on `st`.  part must hold kK4MaxGrid records.
cudaError_t k4_rtvq_launch(const float* x, int64_t n, int bits, int stages, void* codes, int64_t codes_ld,
                           int code_bytes, float* scale, float* zp, float* resnorm, K4Stats* part, uint32_t* packed,
                           int64_t packed_ld, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    if (stages < 1 || stages > kCoreMaxStages || bits < 1 || bits > 16) return cudaErrorInvalidValue;
    if (code_bytes != 1 && code_bytes != 2) return cudaErrorInvalidValue;
    if (code_bytes == 1 && bits > 8) return cudaErrorInvalidValue;
    if (code_bytes == 2 && stages != 1) return cudaErrorInvalidValue;   // int16 codes: single-stage quantiser only
    const int grid = grid_for((n + kVec - 1) / kVec);
    K4PassArgs a;
    a.x = x; a.n = n; a.bits = bits; a.stages = stages; a.scale = scale; a.zp = zp; a.codes = codes;
    a.codes_ld = codes_ld; a.part = part;
    const bool packable = code_bytes == 1 && (bits == 1 || bits == 2 || bits == 4 || bits == 8);
    if (packed != nullptr && !packable) return cudaErrorInvalidValue;
    a.packed = packed; a.packed_ld = packed_ld;
    // small / medium tensors: one cooperative launch with the residual resident in registers
    int grid_max = 0;
    const int64_t cap = k4_fused_capacity(&grid_max);
    if (code_bytes == 1 && n <= cap && grid_max > 0 && 2 * grid_max <= kK4MaxGrid) {
        a.pass = 0;
        const int64_t per8 = (int64_t)grid_max * kK4fThreads * 8, per16 = per8 * 2;
        if (n <= per8) return k4_fused_launch<8>(a, scale, zp, resnorm, grid_max, st);
        if (n <= per16) return k4_fused_launch<16>(a, scale, zp, resnorm, grid_max, st);
        return k4_fused_launch<32>(a, scale, zp, resnorm, grid_max, st);
    }
    for (int q = 0; q <= stages; ++q) {
        a.pass = q;
        if (code_bytes == 1) k4_rtvq_pass<uint8_t><<<grid, kBlock, 0, st>>>(a);
        else                 k4_rtvq_pass<int16_t><<<grid, kBlock, 0, st>>>(a);
        if (q < stages) k4_finalize<<<1, 32, 0, st>>>(part, grid, bits, q, scale, zp, resnorm);
    }
    return cudaGetLastError();
}

cudaError_t k4_dequant_launch(const void* codes, int64_t codes_ld, int code_bytes, int stages, int64_t n,
                              const float* scale, const float* zp, float* out, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    if (stages < 1 || stages > kCoreMaxStages) return cudaErrorInvalidValue;
    const int grid = grid_for(n);
    if (code_bytes == 1) k4_dequant<uint8_t><<<grid, kBlock, 0, st>>>(codes, codes_ld, stages, n, scale, zp, out);
    else if (code_bytes == 2) k4_dequant<int16_t><<<grid, kBlock, 0, st>>>(codes, codes_ld, stages, n, scale, zp, out);
    else return cudaErrorInvalidValue;
    return cudaGetLastError();
}

cudaError_t k4_absmax_launch(const float* x, int64_t n, int bits, void* q, int code_bytes, float* scale,
                             K4Stats* part, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    if (bits < 2 || bits > 16 || (code_bytes != 1 && code_bytes != 2)) return cudaErrorInvalidValue;
    const int grid = grid_for(n);
    k4_absmax_stats<<<grid, kBlock, 0, st>>>(x, n, part);
    k4_absmax_finalize<<<1, 32, 0, st>>>(part, grid, bits, scale);
    if (code_bytes == 1) k4_absmax_quant<int8_t><<<grid, kBlock, 0, st>>>(x, n, scale, reinterpret_cast<int8_t*>(q));
    else                 k4_absmax_quant<int16_t><<<grid, kBlock, 0, st>>>(x, n, scale, reinterpret_cast<int16_t*>(q));
    return cudaGetLastError();
}

}  // namespace svdq
