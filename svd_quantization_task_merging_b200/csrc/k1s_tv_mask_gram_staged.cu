// K1 (staged) — the same pass 1 as k1_tv_mask_gram.cu (task vectors + tall-mask combination + masked
// Gram, one partial per tile) as a PERSISTENT, warp-specialised kernel: one CTA per SM, a producer warp
// streams 1024-element chunks of the (N+1) tensors and N masks global -> shared with the TMA bulk-copy
// engine (cp.async.bulk, mbarrier complete_tx) into a STAGES-deep ring, 8 consumer warps do the
// arithmetic out of shared memory.  Up to STAGES x 44 KB (N = 8, fp32) of reads are in flight per SM
// without occupying registers.  Per-tile outputs are identical to the non-staged kernel (same
// per-thread element assignment, same reduction order).  Chunks that are not a whole 1024 elements
// (the tail of a parameter) are loaded directly by the consumers.
//
// Replaces the same reference lines as k1_tv_mask_gram.cu.  Bound: HBM.
#include "stage_pipe.cuh"
#include "svdq_kernels.h"

#ifndef SVDQ_DTYPE
#define SVDQ_DTYPE 0
#endif

namespace svdq {

constexpr int kK1Stages = 4;
constexpr int kRedRows = 32;

template <typename T> struct SmemElem;
template <> struct SmemElem<float> {
    static __device__ __forceinline__ void load4(const unsigned char* slot, int tid, float (&o)[4]) {
        const float4 v = *reinterpret_cast<const float4*>(slot + tid * 16);
        o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
    }
};
template <> struct SmemElem<__nv_bfloat16> {
    static __device__ __forceinline__ void load4(const unsigned char* slot, int tid, float (&o)[4]) {
        const uint2 v = *reinterpret_cast<const uint2*>(slot + tid * 8);
        o[0] = __uint_as_float(v.x << 16); o[1] = __uint_as_float(v.x & 0xffff0000u);
        o[2] = __uint_as_float(v.y << 16); o[3] = __uint_as_float(v.y & 0xffff0000u);
    }
};
template <> struct SmemElem<__half> {
    static __device__ __forceinline__ void load4(const unsigned char* slot, int tid, float (&o)[4]) {
        const uint2 v = *reinterpret_cast<const uint2*>(slot + tid * 8);
        const __half2 a = *reinterpret_cast<const __half2*>(&v.x), b = *reinterpret_cast<const __half2*>(&v.y);
        const float2 fa = __half22float2(a), fb = __half22float2(b);
        o[0] = fa.x; o[1] = fa.y; o[2] = fb.x; o[3] = fb.y;
    }
};

template <typename T, int NT>
__host__ __device__ constexpr int k1s_stage_bytes() { return (NT + 1) * kStep * (int)sizeof(T) + NT * kStep; }
template <typename T, int NT>
__host__ __device__ constexpr int k1s_smem_bytes() {
    return kK1Stages * k1s_stage_bytes<T, NT>() + kRedRows * (kBlock + 1) * 4 + 2 * kK1Stages * 8 + kK1Stages * 4 + 64;
}

template <typename T, int NT, bool FULL>
__global__ void __launch_bounds__(kBlock + 32, 1) k1s_tv_mask_gram(const K1Args a, const int n_tiles) {
    constexpr int G = tri_count(NT);
    constexpr int NACC = FULL ? 2 * G : G;
    constexpr int STAGES = kK1Stages;
    constexpr int kTensorBytes = kStep * (int)sizeof(T);
    constexpr int kStageBytes = k1s_stage_bytes<T, NT>();
    constexpr int kMaskOff = (NT + 1) * kTensorBytes;

    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* stage_base = smem;
    float(*red)[kBlock + 1] = reinterpret_cast<float(*)[kBlock + 1]>(smem + STAGES * kStageBytes);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * kStageBytes + kRedRows * (kBlock + 1) * 4);
    uint64_t* empty = full + STAGES;
    int* s_direct = reinterpret_cast<int*>(empty + STAGES);
    __shared__ const void* s_ptr[NT + 1];
    __shared__ const uint8_t* s_mask[NT];
    __shared__ uint32_t s_cnt[kBlock / 32];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kBlock / 32); }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == kBlock / 32) {
        // ================= producer warp: lane i owns stream i (tensor i for i <= NT, mask i-NT-1 after) ===
        // every lane issues its own TMA bulk copy, so the (N+1)+N copies of a chunk are issued in parallel
        PipeState ps;
        const bool is_tensor = lane <= NT;
        const bool is_mask = lane > NT && lane <= 2 * NT;
        const int kMaskChunk = a.mask_bits ? kStep / 8 : kStep;          // bytes of one task's mask per chunk
        const int my_bytes = is_tensor ? kTensorBytes : kMaskChunk;
        const int my_off = is_tensor ? lane * kTensorBytes : kMaskOff + (lane - NT - 1) * kStep;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int p = a.tile_param[tile];
            const int64_t numel = a.numel[p];
            const int64_t start = (int64_t)a.tile_local[tile] * a.tile_elems;
            const int64_t stop = min(start + (int64_t)a.tile_elems, numel);
            const unsigned char* my_ptr = nullptr;
            if (is_tensor) {
                const void* q = a.tensors[(int64_t)p * (NT + 1) + lane];
                my_ptr = reinterpret_cast<const unsigned char*>(q ? q : a.tensors[(int64_t)p * (NT + 1)]);
            } else if (is_mask && a.masks) {
                my_ptr = a.masks[(int64_t)p * NT + (lane - NT - 1)];
            }
            const int n_present = __popc(__ballot_sync(0xffffffffu, is_mask && my_ptr != nullptr));
            const int64_t my_esize = is_tensor ? (int64_t)sizeof(T) : 1;
            const bool bits_in = a.mask_bits != 0;
            for (int64_t e0 = start; e0 < stop; e0 += kStep) {
                if (lane == 0) mbar_wait(&empty[ps.stage], ps.phase ^ 1u);
                __syncwarp();
                unsigned char* sb = stage_base + ps.stage * kStageBytes;
                if (e0 + kStep <= numel) {
                    if (lane == 0) {
                        s_direct[ps.stage] = 0;
                        mbar_arrive_expect_tx(&full[ps.stage], (uint32_t)((NT + 1) * kTensorBytes + n_present * kMaskChunk));
                    }
                    __syncwarp();
                    if (my_ptr != nullptr)
                        bulk_g2s(sb + my_off, my_ptr + ((!is_tensor && bits_in) ? e0 / 8 : e0 * my_esize), my_bytes,
                                 &full[ps.stage]);
                } else if (lane == 0) {
                    s_direct[ps.stage] = 1;          // tail chunk: consumers load it themselves
                    mbar_arrive(&full[ps.stage]);
                }
                ps.advance<STAGES>();
            }
        }
        return;
    }

    // ======================= consumer warps ===========================================================
    PipeState ps;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int p = a.tile_param[tile];
        const int64_t numel = a.numel[p];
        const int64_t start = (int64_t)a.tile_local[tile] * a.tile_elems;
        const int64_t stop = min(start + (int64_t)a.tile_elems, numel);
        named_bar_sync(1, kBlock);                       // previous tile finished with s_ptr / s_mask / red
        if (tid <= NT) {
            const void* q = a.tensors[(int64_t)p * (NT + 1) + tid];
            s_ptr[tid] = q ? q : a.tensors[(int64_t)p * (NT + 1)];
        }
        if (tid < NT) s_mask[tid] = a.masks ? a.masks[(int64_t)p * NT + tid] : nullptr;
        named_bar_sync(1, kBlock);

        int n_present = 0;
#pragma unroll
        for (int t = 0; t < NT; ++t) n_present += s_mask[t] != nullptr;
        const bool has_mask = n_present > 0;
        const uint32_t thr_bytes = 0x01010101u * (uint32_t)(a.strategy == kUnion ? 1 : n_present);
        const bool majority = a.strategy == kMajority;
        uint32_t* packed = has_mask ? a.packed + a.pmask_off[p] : nullptr;

        const bool comp = a.second_complement != 0;
        const bool mask_bits = a.mask_bits != 0;
        float2 acc2[G];
#pragma unroll
        for (int i = 0; i < G; ++i) acc2[i] = make_float2(0.0f, 0.0f);
        uint32_t cnt = 0;

        for (int64_t e0 = start; e0 < stop; e0 += kStep) {
            const int64_t e = e0 + (int64_t)tid * kVec;
            const bool active = e < stop;
            const bool fullv = e + kVec <= numel;
            float b[kVec], f[NT][kVec];
            uint32_t mw[NT];
            mbar_wait(&full[ps.stage], ps.phase);
            const unsigned char* sb = stage_base + ps.stage * kStageBytes;
            if (!s_direct[ps.stage]) {
                SmemElem<T>::load4(sb, tid, b);
#pragma unroll
                for (int t = 0; t < NT; ++t) SmemElem<T>::load4(sb + (t + 1) * kTensorBytes, tid, f[t]);
#pragma unroll
                for (int t = 0; t < NT; ++t) {
                    if (s_mask[t] == nullptr) mw[t] = 0u;
                    else if (!mask_bits) mw[t] = *reinterpret_cast<const uint32_t*>(sb + kMaskOff + t * kStep + tid * 4);
                    else mw[t] = nibble_to_bytes((*reinterpret_cast<const uint32_t*>(sb + kMaskOff + t * kStep + (tid >> 3) * 4)
                                                  >> ((tid & 7) * 4)) & 0xFu);
                }
            } else if (active && fullv) {
                Elem<T>::load4(s_ptr[0], e, b);
#pragma unroll
                for (int t = 0; t < NT; ++t) Elem<T>::load4(s_ptr[t + 1], e, f[t]);
#pragma unroll
                for (int t = 0; t < NT; ++t) {
                    if (s_mask[t] == nullptr) mw[t] = 0u;
                    else if (!mask_bits) mw[t] = ldg_stream_u32(s_mask[t] + e);
                    else mw[t] = nibble_to_bytes((__ldg(reinterpret_cast<const uint32_t*>(s_mask[t]) + (e >> 5))
                                                  >> (int)(e & 31)) & 0xFu);
                }
            } else if (active) {
#pragma unroll
                for (int c = 0; c < kVec; ++c) b[c] = (e + c < numel) ? Elem<T>::load1(s_ptr[0], e + c) : 0.0f;
#pragma unroll
                for (int t = 0; t < NT; ++t) {
#pragma unroll
                    for (int c = 0; c < kVec; ++c) f[t][c] = (e + c < numel) ? Elem<T>::load1(s_ptr[t + 1], e + c) : 0.0f;
                    mw[t] = 0;
                    if (s_mask[t] != nullptr && mask_bits) {
                        mw[t] = nibble_to_bytes((__ldg(reinterpret_cast<const uint32_t*>(s_mask[t]) + (e >> 5))
                                                 >> (int)(e & 31)) & 0xFu);       // bits past numel are cut by `valid`
                    } else if (s_mask[t] != nullptr) {
#pragma unroll
                        for (int c = 0; c < kVec; ++c)
                            if (e + c < numel) mw[t] |= (uint32_t)__ldg(s_mask[t] + e + c) << (8 * c);
                    }
                }
            } else {
#pragma unroll
                for (int c = 0; c < kVec; ++c) b[c] = 0.0f;
#pragma unroll
                for (int t = 0; t < NT; ++t) {
                    mw[t] = 0;
#pragma unroll
                    for (int c = 0; c < kVec; ++c) f[t][c] = 0.0f;
                }
            }
            // the slot's data now lives in registers: hand it back to the producer before computing
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[ps.stage]);
            ps.advance<STAGES>();

            float d[NT][kVec];
#pragma unroll
            for (int t = 0; t < NT; ++t) Elem<T>::template subv<kVec>(f[t], b, d[t]);
            uint32_t bits = 0;
            if (active) {
                const uint32_t valid = fullv ? 0xFu : ((1u << (int)(numel - e)) - 1u);
                if (has_mask) {
                    uint32_t votes = 0;
#pragma unroll
                    for (int t = 0; t < NT; ++t) votes += __vminu4(mw[t], 0x01010101u);
                    if (majority) votes += votes;
                    const uint32_t ge = __vcmpgeu4(votes, thr_bytes);
                    bits = ((ge >> 7) & 1u) | ((ge >> 14) & 2u) | ((ge >> 21) & 4u) | ((ge >> 28) & 8u);
                    bits &= valid;
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
                if ((lane & 7) == 0 && active) packed[e >> 5] = w;
            }
            // packed 2-wide FMAs (fma.rn.f32x2): FULL pairs the masked Gram with the all-element Gram of the
            // same (i, j); otherwise two consecutive elements share one instruction
            if (FULL) {
#pragma unroll
                for (int c = 0; c < kVec; ++c) {
                    const bool m = (bits >> c) & 1u;
                    float2 v[NT];
#pragma unroll
                    for (int t = 0; t < NT; ++t) v[t] = make_float2(m ? d[t][c] : 0.0f, (comp && m) ? 0.0f : d[t][c]);
#pragma unroll
                    for (int i = 0; i < NT; ++i)
#pragma unroll
                        for (int j = i; j < NT; ++j)
                            acc2[tri_index(i, j, NT)] = __ffma2_rn(v[i], v[j], acc2[tri_index(i, j, NT)]);
                }
            } else {
#pragma unroll
                for (int c = 0; c < kVec; c += 2) {
                    const bool m0 = (bits >> c) & 1u, m1 = (bits >> (c + 1)) & 1u;
                    float2 v[NT];
#pragma unroll
                    for (int t = 0; t < NT; ++t) v[t] = make_float2(m0 ? d[t][c] : 0.0f, m1 ? d[t][c + 1] : 0.0f);
#pragma unroll
                    for (int i = 0; i < NT; ++i)
#pragma unroll
                        for (int j = i; j < NT; ++j)
                            acc2[tri_index(i, j, NT)] = __ffma2_rn(v[i], v[j], acc2[tri_index(i, j, NT)]);
                }
            }
        }

        // unpack the packed accumulators: rows [0, G) masked Gram, rows [G, 2G) all-element Gram (FULL)
        float acc[NACC];
#pragma unroll
        for (int i = 0; i < G; ++i) {
            if (FULL) { acc[i] = acc2[i].x; acc[G + i] = acc2[i].y; }
            else acc[i] = acc2[i].x + acc2[i].y;
        }
        // ---- CTA reduction (consumers only), same order as the non-staged kernel ----------------------
        float* gout = a.gram + (int64_t)tile * NACC;
#pragma unroll
        for (int r0 = 0; r0 < NACC; r0 += kRedRows) {
#pragma unroll
            for (int r = 0; r < kRedRows; ++r)
                if (r0 + r < NACC) red[r][tid] = acc[r0 + r];
            named_bar_sync(1, kBlock);
#pragma unroll
            for (int rr = 0; rr < kRedRows / (kBlock / 32); ++rr) {
                const int r = warp * (kRedRows / (kBlock / 32)) + rr;
                if (r0 + r < NACC) {
                    float s = 0.0f;
#pragma unroll
                    for (int c = 0; c < kBlock / 32; ++c) s += red[r][lane + 32 * c];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                    if (lane == 0) gout[r0 + r] = s;
                }
            }
            named_bar_sync(1, kBlock);
        }
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        if (lane == 0) s_cnt[warp] = cnt;
        named_bar_sync(1, kBlock);
        if (tid == 0) {
            uint32_t c = 0;
#pragma unroll
            for (int w = 0; w < kBlock / 32; ++w) c += s_cnt[w];
            a.count[tile] = c;
        }
    }
}

template <typename T, int NT>
static cudaError_t launch_staged(const K1Args& a, int n_tiles, bool full, int n_sm, cudaStream_t st) {
    if (n_tiles <= 0) return cudaSuccess;
    constexpr int smem = k1s_smem_bytes<T, NT>();
    const int grid = n_tiles < n_sm ? n_tiles : n_sm;
    cudaError_t e;
    if (full) {
        e = cudaFuncSetAttribute(k1s_tv_mask_gram<T, NT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        k1s_tv_mask_gram<T, NT, true><<<grid, kBlock + 32, smem, st>>>(a, n_tiles);
    } else {
        e = cudaFuncSetAttribute(k1s_tv_mask_gram<T, NT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        k1s_tv_mask_gram<T, NT, false><<<grid, kBlock + 32, smem, st>>>(a, n_tiles);
    }
    return cudaGetLastError();
}

// staged path exists for nt <= 8 (stage ring of 4 x 44 KB at N = 8, fp32); returns cudaErrorNotSupported otherwise
template <>
cudaError_t k1s_launch_dtype<SVDQ_DTYPE>(int nt, const K1Args& a, int n_tiles, bool full, int n_sm, cudaStream_t st) {
    using T = DTypeOf<SVDQ_DTYPE>::type;
    switch (nt) {
#define SVDQ_CASE(N) case N: return launch_staged<T, N>(a, n_tiles, full, n_sm, st);
        SVDQ_CASE(1) SVDQ_CASE(2) SVDQ_CASE(3) SVDQ_CASE(4) SVDQ_CASE(5) SVDQ_CASE(6) SVDQ_CASE(7) SVDQ_CASE(8)
#undef SVDQ_CASE
        default: return cudaErrorNotSupported;
    }
}

}  // namespace svdq
