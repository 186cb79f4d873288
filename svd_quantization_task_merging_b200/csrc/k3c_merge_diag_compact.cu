// K3c — pass 2 WITH the fused per-task reconstruction diagnostics (svd_eval_reconstruction, the reference's
// default), compacting variant for up to 8 task vectors.  Same inputs, outputs and reference lines as
// k3_reconstruct_merge<T, NT, FP16B, DIAG = true> (reconstruct_from_coefficients src/svd_hybrid/merge.py:180-192,
// reconstruct_from_masked src/svd_hybrid/mask_loader.py:750-763, apply_merged_deltas merge.py:486-488,
// compute_reconstruction_error src/svd_hybrid/diagnostics.py:101-117,205-216).
//
// Why.  With the diagnostics fused pass 2 executes ~300 instructions per element (N r FMAs for the basis row, N r for
// the per-task reconstructions, four running reductions per task) and is issue-bound at 57 % of the HBM roofline
// (profiles/r2_ncu_k3diag.csv) -- but only elements INSIDE the combined tall mask need any of it: outside, merged =
// base and nothing enters the diagnostics.  Intersection masks keep 43 % of the elements, majority masks 64 %.
//
// How.  Per step of 1024 elements a CTA (A) loads base + N fine-tuned values of its elements with coalesced 128-bit
// loads, writes merged = base for all of them, and drops the values of the KEPT elements into a compacted
// shared-memory buffer (block-wide prefix sum of the mask bits); (B) the threads then walk the compacted elements two at
// a time through the unchanged per-element arithmetic (k3_step: same instruction sequence, so merged weights are
// bit-identical to the non-compacting kernel) and store merged = base + delta at the elements' positions.  The
// diagnostics partials of a tile are summed in a different (still fixed, placement-independent) order.
// Bound: HBM ((N+1) s + 1/8 + 4 B per element) once the mask is sparse enough; issue otherwise.
#include <type_traits>

#include "k3_body.cuh"
#include "copy_out.cuh"

#ifndef SVDQ_DTYPE
#define SVDQ_DTYPE 0
#endif

namespace svdq {

constexpr int kCVec = 4;
constexpr int kCStep = kBlock * kCVec;            // 1024 elements per step

// 16- / 8-byte asynchronous copy global -> shared (four consecutive elements of a 4- / 2-byte type)
template <int BYTES> __device__ __forceinline__ void cp_async_vec(void* smem_dst, const void* gmem_src);
template <> __device__ __forceinline__ void cp_async_vec<16>(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
template <> __device__ __forceinline__ void cp_async_vec<8>(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// four staged elements of tensor row `t` for this thread, as fp32
template <typename T> struct StagedQuad;
template <> struct StagedQuad<float> {
    static __device__ __forceinline__ void load(const unsigned char* raw, int t, int tid, float (&o)[4]) {
        const float4 v = *reinterpret_cast<const float4*>(raw + ((size_t)t * kCStep + tid * 4) * 4);
        o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
    }
};
template <> struct StagedQuad<__nv_bfloat16> {
    static __device__ __forceinline__ void load(const unsigned char* raw, int t, int tid, float (&o)[4]) {
        const uint2 v = *reinterpret_cast<const uint2*>(raw + ((size_t)t * kCStep + tid * 4) * 2);
        o[0] = __uint_as_float(v.x << 16); o[1] = __uint_as_float(v.x & 0xffff0000u);
        o[2] = __uint_as_float(v.y << 16); o[3] = __uint_as_float(v.y & 0xffff0000u);
    }
};
template <> struct StagedQuad<__half> {
    static __device__ __forceinline__ void load(const unsigned char* raw, int t, int tid, float (&o)[4]) {
        const uint2 v = *reinterpret_cast<const uint2*>(raw + ((size_t)t * kCStep + tid * 4) * 2);
        const float2 fa = __half22float2(*reinterpret_cast<const __half2*>(&v.x));
        const float2 fb = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
        o[0] = fa.x; o[1] = fa.y; o[2] = fb.x; o[3] = fb.y;
    }
};

// drops the basis rows of the thread's two compacted elements into the shared-memory image of the step's slice of
// U_high / U_low / mean (rows lr, lr + 1 of the step); the CTA copies the slices out with 16-byte stores afterwards
template <typename OUT> struct K3BasisSink {
    static constexpr bool on = true;
    OUT* s_h; OUT* s_l; float* s_m;         // images, already re-based to the step (null s_h: off)
    int lr, k, nlow;
    bool two;
    __device__ __forceinline__ void col(int j, float2 u) const {
        if (s_h == nullptr) return;
        if (j < k) { s_h[lr * k + j] = OutCvt<OUT>::cvt(u.x); if (two) s_h[(lr + 1) * k + j] = OutCvt<OUT>::cvt(u.y); }
        else { s_l[lr * nlow + (j - k)] = OutCvt<OUT>::cvt(u.x); if (two) s_l[(lr + 1) * nlow + (j - k)] = OutCvt<OUT>::cvt(u.y); }
    }
    __device__ __forceinline__ void mean2(float m0, float m1) const {
        if (s_m == nullptr) return;
        s_m[lr] = m0;
        if (two) s_m[lr + 1] = m1;
    }
};

template <typename T, int NT, bool FP16B, bool DIAG, bool BASIS>
__global__ void __launch_bounds__(kBlock, 2) k3c_merge_diag_compact(const K3Args a) {
    constexpr int NTP = (NT + 3) & ~3;
    constexpr int NR = DIAG ? kDiagRows * NT : 1;
    constexpr int ES = (int)sizeof(T);
    __shared__ __align__(16) float sWT[NT][NTP];        // sWT[j][t] = W[t][j]
    __shared__ __align__(16) float sChatT[DIAG ? NT : 1][NTP];     // sChatT[j][t] = chat[t][j]
    __shared__ float sCbar[NT], sG[NT];
    __shared__ const void* s_ptr[NT + 1];
    __shared__ uint32_t s_wtot[kBlock / 32];
    // dyn: compacted values of the step's kept elements s_x[t][slot] (t = NT: base), s_idx[slot] = offset inside the
    // step, and the raw staging area of the NEXT step (cp.async lands there while this step's arithmetic runs);
    // the start of the area also serves the CTA reduction of the diagnostics at the end
    extern __shared__ __align__(16) float dyn[];
    float* s_x = dyn;                                             // [(NT + 1)][kCStep]
    uint16_t* s_idx = reinterpret_cast<uint16_t*>(s_x + (NT + 1) * kCStep);      // [kCStep]
    unsigned char* s_raw = reinterpret_cast<unsigned char*>(s_idx + kCStep);     // [(NT + 1)][kCStep] elements of T (row 0: base)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x;
    const int p = a.tile_param[tile];
    const int64_t numel = a.numel[p];
    const int64_t start = (int64_t)a.tile_local[tile] * a.tile_elems;
    const int64_t stop = min(start + (int64_t)a.tile_elems, numel);
    const int status = a.info[(int64_t)p * 8 + 0];
    const int n_active = a.info[(int64_t)p * 8 + 1];
    // with diagnostics or a basis write-out all r columns are walked (NaN coefficients show up as in the reference;
    // the artifacts hold all r columns), otherwise only the r_eff that are not numerically null
    const int r = a.info[(int64_t)p * 8 + ((DIAG || BASIS) ? 2 : 4)];
    const float tail_add = a.scal[(int64_t)p * 4 + 1];
    const float mean_scale = a.scal[(int64_t)p * 4 + 2];
    const bool has_mask = a.has_mask[p] != 0;
    // automatic mode (diag_select == 2): per tile -- compaction where the mask keeps fewer than 55 % of the elements, the
    // plain two-elements-per-thread walk (the loop of k3_reconstruct_merge) elsewhere and for parameters without a basis
    constexpr bool write_basis = BASIS;                   // every solved tile is compacted then (rows = compacted slots)
    const bool compact = a.diag_select != 2 || (write_basis && status == kSolved) ||
                         (status == kSolved && k3_tile_is_sparse(has_mask ? a.packed + a.pmask_off[p] : nullptr, start, stop, numel));

    if (tid <= NT) s_ptr[tid] = a.tensors[(int64_t)p * (NT + 1) + tid];
    for (int i = tid; i < NT * NTP; i += kBlock) {
        const int j = i / NTP, t = i % NTP;
        sWT[j][t] = (t < NT) ? a.W[(int64_t)p * NT * NT + t * NT + j] : 0.0f;
        if (DIAG) sChatT[j][t] = (t < NT) ? a.chat[(int64_t)p * NT * NT + t * NT + j] : 0.0f;
    }
    if (tid < NT) { sCbar[tid] = a.cbar[(int64_t)p * NT + tid]; sG[tid] = a.gvec[(int64_t)p * NT + tid]; }
    __syncthreads();
    const uint32_t* packed = has_mask ? a.packed + a.pmask_off[p] : nullptr;
    float* outp = a.out[p];
    const float n_f = (float)(n_active > 0 ? n_active : 1);
    float dacc[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) dacc[i] = 0.0f;
    // tasks that lack the parameter: pointer replaced by the base tensor (delta == 0), see K1
    uint32_t present_bits = 0;
#pragma unroll
    for (int t = 0; t < NT; ++t) present_bits |= (s_ptr[t + 1] != nullptr ? 1u : 0u) << t;
    __syncthreads();
    if (tid >= 1 && tid <= NT && s_ptr[tid] == nullptr) s_ptr[tid] = s_ptr[0];
    __syncthreads();
    const bool solved = status == kSolved;

    // asynchronous staging of one step: this thread's four elements of every tensor (whole vectors only; the ragged
    // end of a parameter is read directly)
    auto stage = [&](int64_t s0) {
        const int64_t e = s0 + (int64_t)tid * kCVec;
        if (e + kCVec <= numel && e < stop) {
#pragma unroll
            for (int t = 0; t <= NT; ++t) {
                if (t > 0 && !solved) break;
                cp_async_vec<4 * ES>(s_raw + ((size_t)t * kCStep + tid * 4) * ES,
                                     reinterpret_cast<const unsigned char*>(s_ptr[t]) + e * ES);
            }
        }
    };
    if (!compact) {
        constexpr int kStepV = kBlock * 2;
        for (int64_t e0 = start; e0 < stop; e0 += kStepV) {
            const int64_t e = e0 + (int64_t)tid * 2;
            if (e >= stop) continue;
            const bool full = e + 2 <= numel;
            float b[2], res[2];
            ElemPair<T>::load(s_ptr[0], e, full, numel, b);
            if (!solved) { res[0] = b[0]; res[1] = b[1]; }
            else {
                float x[NT][2];
#pragma unroll
                for (int t = 0; t < NT; ++t) ElemPair<T>::load(s_ptr[t + 1], e, full, numel, x[t]);
                const uint32_t pword = has_mask ? __ldg(packed + (e >> 5)) : 0xffffffffu;
                k3_step<T, NT, FP16B, DIAG, false, 2>(b, x, pword, e, numel, r, present_bits, a.center, n_f, tail_add, mean_scale,
                                                      sWT, sChatT, sCbar, sG, res, dacc);
            }
            if (full) stg_stream_f2(outp + e, make_float2(res[0], res[1]));
            else if (e < numel) outp[e] = res[0];
        }
    } else {
    using OUT = typename std::conditional<FP16B, __half, float>::type;
    constexpr int V = 16 / (int)sizeof(OUT);
    OUT* uh_out = write_basis ? reinterpret_cast<OUT*>(a.u_high[p]) : nullptr;
    OUT* ul_out = write_basis ? reinterpret_cast<OUT*>(a.u_low[p]) : nullptr;
    float* mean_out = (write_basis && a.mean_out) ? a.mean_out[p] : nullptr;
    // image of the step's artifact slices behind the staging area (allocated only when the bases are written)
    OUT* s_u = reinterpret_cast<OUT*>(s_raw + (size_t)(NT + 1) * kCStep * ES);
    float* s_mean = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(s_u) +
                                             ((size_t)(kCStep * NT + 3 * V) * sizeof(OUT) + 15) / 16 * 16);
    K3BasisSink<OUT> sink;
    sink.k = a.info[(int64_t)p * 8 + 3];
    sink.nlow = r - sink.k;
    int64_t row_base = write_basis ? a.tile_row_off[tile] : 0;
    stage(start);

    for (int64_t e0 = start; e0 < stop; e0 += kCStep) {
        // ---- A: mask bits and the block-wide prefix sum of the kept elements ------------------------------------------
        const int64_t e = e0 + (int64_t)tid * kCVec;
        const bool active = e < stop;
        const bool full = e + kCVec <= numel;
        uint32_t bits = 0;
        if (active && solved) {
            const uint32_t valid = full ? 0xFu : ((1u << (int)(numel - e)) - 1u);
            bits = (has_mask ? (__ldg(packed + (e >> 5)) >> (int)(e & 31)) & 0xFu : 0xFu) & valid;
        }
        const uint32_t mine = __popc(bits);
        uint32_t incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) s_wtot[warp] = incl;
        cp_async_commit_wait_all();                         // this step's raw values have landed (this thread's copies)
        __syncthreads();                                    // ... everybody's; and phase B of the previous step has read s_x
        uint32_t slot = incl - mine, kept = 0;
#pragma unroll
        for (int w = 0; w < kBlock / 32; ++w) {
            const uint32_t v = s_wtot[w];
            if (w < warp) slot += v;
            kept += v;
        }
        // ---- merged = base everywhere; the kept elements' values into the compacted buffer, tensor by tensor --------
        if (active) {
            float b[kCVec];
            if (full) StagedQuad<T>::load(s_raw, 0, tid, b);
            else {
#pragma unroll
                for (int c = 0; c < kCVec; ++c) b[c] = (e + c < numel) ? Elem<T>::load1(s_ptr[0], e + c) : 0.0f;
            }
            if (full) stg_stream_f4(outp + e, make_float4(b[0], b[1], b[2], b[3]));
            else {
#pragma unroll
                for (int c = 0; c < kCVec; ++c)
                    if (e + c < numel) outp[e + c] = b[c];
            }
            if (bits) {
                uint32_t sl = slot;
#pragma unroll
                for (int c = 0; c < kCVec; ++c) {
                    if ((bits >> c) & 1u) {
                        s_x[NT * kCStep + sl] = b[c];
                        s_idx[sl] = (uint16_t)(tid * kCVec + c);
                        ++sl;
                    }
                }
#pragma unroll
                for (int t = 0; t < NT; ++t) {
                    float x[kCVec];
                    if (full) StagedQuad<T>::load(s_raw, t + 1, tid, x);
                    else {
#pragma unroll
                        for (int c = 0; c < kCVec; ++c) x[c] = (e + c < numel) ? Elem<T>::load1(s_ptr[t + 1], e + c) : 0.0f;
                    }
                    sl = slot;
#pragma unroll
                    for (int c = 0; c < kCVec; ++c) {
                        if ((bits >> c) & 1u) { s_x[t * kCStep + sl] = x[c]; ++sl; }
                    }
                }
            }
        }
        __syncthreads();                                    // compacted buffer complete; s_raw free again
        if (e0 + kCStep < stop) stage(e0 + kCStep);         // the next step's loads fly while this step's arithmetic runs
        if (!solved) continue;                              // uniform: parameter without a basis, merged = base
        // ---- B: the kept elements, two at a time, through the unchanged per-element arithmetic ----------------------
        const int items = (int)(kept + 1) >> 1;
        const int64_t gh = row_base * sink.k, gl = row_base * sink.nlow, gm = row_base;
        sink.s_h = write_basis ? s_u + (gh % V) : nullptr;
        sink.s_l = s_u + ((gh % V) + (int64_t)kept * sink.k + V - 1) / V * V + (gl % V);
        sink.s_m = mean_out ? s_mean + (gm % 4) : nullptr;
        for (int it = tid; it < items; it += kBlock) {
            const bool two = 2 * it + 1 < (int)kept;
            float b2[2], x2[NT][2], res[2];
            const float2 bb = *reinterpret_cast<const float2*>(s_x + NT * kCStep + 2 * it);
            b2[0] = bb.x; b2[1] = two ? bb.y : 0.0f;
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                const float2 v = *reinterpret_cast<const float2*>(s_x + t * kCStep + 2 * it);
                x2[t][0] = v.x; x2[t][1] = two ? v.y : 0.0f;
            }
            const uint32_t i0 = s_idx[2 * it], i1 = two ? s_idx[2 * it + 1] : 0u;
            // both elements are inside the mask and inside the tensor: pword = 0b11 (0b01), "numel" = their count
            sink.lr = 2 * it;
            sink.two = two;
            if constexpr (BASIS)
                k3_step<T, NT, FP16B, DIAG, false, 2, K3BasisSink<OUT>>(b2, x2, two ? 3u : 1u, (int64_t)0, (int64_t)(two ? 2 : 1), r,
                                                                       present_bits, a.center, n_f, tail_add, mean_scale, sWT,
                                                                       sChatT, sCbar, sG, res, dacc, K3NoiseSet<NT>(), sink);
            else
                k3_step<T, NT, FP16B, DIAG, false, 2>(b2, x2, two ? 3u : 1u, (int64_t)0, (int64_t)(two ? 2 : 1), r, present_bits,
                                                      a.center, n_f, tail_add, mean_scale, sWT, sChatT, sCbar, sG, res, dacc);
            outp[e0 + i0] = res[0];
            if (two) outp[e0 + i1] = res[1];
        }
        if (write_basis) {
            __syncthreads();                                // the step's image is complete
            k5_copy_out<OUT>(uh_out, sink.s_h, gh, (int)kept * sink.k, tid);
            k5_copy_out<OUT>(ul_out, sink.s_l, gl, (int)kept * sink.nlow, tid);
            if (mean_out) k5_copy_out<float>(mean_out, sink.s_m, gm, (int)kept, tid);
            row_base += kept;
            // (the image is rewritten two barriers from here, in phase B of the next step)
        }
        // (the next step's first barrier separates these reads of s_x / s_idx from its writes)
    }
    cp_async_commit_wait_all();
    }

    if (!DIAG) return;
    // ---- CTA reduction of the 4*NT diagnostic rows: sums for rows < 3*NT, max for the rest --------------------------
    __syncthreads();
    constexpr int kRows = 16;
    float (*red)[kBlock + 1] = reinterpret_cast<float (*)[kBlock + 1]>(dyn);         // 16 x 257 floats <= the s_x area
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

template <typename T, int NT, bool DIAG, bool BASIS>
static cudaError_t launch_c(const K3Args& a, int n_tiles, bool fp16b, cudaStream_t st) {
    if (n_tiles <= 0) return cudaSuccess;
    constexpr size_t kCompact = (size_t)(NT + 1) * kCStep * 4 + kCStep * 2 + (size_t)(NT + 1) * kCStep * sizeof(T),
                     kRed = 16 * (kBlock + 1) * 4;
    // + the image of a step's artifact slices when the bases are written (U entries of NT columns, then the means)
    const size_t esz = fp16b ? 2 : 4, vv = 16 / esz;
    const size_t image = BASIS ? ((kCStep * NT + 3 * vv) * esz + 15) / 16 * 16 + (kCStep + 4) * 4 : 0;
    const size_t dsm = (kCompact > kRed ? kCompact : kRed) + image;      // the reduction scratch reuses the compaction buffer
    cudaError_t e;
    if (fp16b) {
        e = cudaFuncSetAttribute(k3c_merge_diag_compact<T, NT, true, DIAG, BASIS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm);
        if (e != cudaSuccess) return e;
        k3c_merge_diag_compact<T, NT, true, DIAG, BASIS><<<n_tiles, kBlock, dsm, st>>>(a);
    } else {
        e = cudaFuncSetAttribute(k3c_merge_diag_compact<T, NT, false, DIAG, BASIS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm);
        if (e != cudaSuccess) return e;
        k3c_merge_diag_compact<T, NT, false, DIAG, BASIS><<<n_tiles, kBlock, dsm, st>>>(a);
    }
    return cudaGetLastError();
}

// compacting pass 2 with fused diagnostics and / or the basis write-out: up to 8 tasks, no noise region;
// cudaErrorNotSupported otherwise
template <>
cudaError_t k3c_launch_dtype<SVDQ_DTYPE>(int nt, const K3Args& a, int n_tiles, bool fp16b, cudaStream_t st) {
    using T = DTypeOf<SVDQ_DTYPE>::type;
    if (a.info_n != nullptr || a.tile_elems % kCStep != 0) return cudaErrorNotSupported;
    const bool diag = a.diag != nullptr && a.chat != nullptr;
    const bool basis = a.u_high != nullptr;
    if (!diag && !basis) return cudaErrorNotSupported;                     // without diagnostics only the basis write-out uses it
    switch (nt) {
#define SVDQ_CASE(N) case N: return !basis ? launch_c<T, N, true, false>(a, n_tiles, fp16b, st) \
                                   : diag ? launch_c<T, N, true, true>(a, n_tiles, fp16b, st) : launch_c<T, N, false, true>(a, n_tiles, fp16b, st);
        SVDQ_CASE(1) SVDQ_CASE(2) SVDQ_CASE(3) SVDQ_CASE(4) SVDQ_CASE(5) SVDQ_CASE(6) SVDQ_CASE(7) SVDQ_CASE(8)
#undef SVDQ_CASE
        default: return cudaErrorNotSupported;
    }
}

}  // namespace svdq
