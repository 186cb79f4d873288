// K1 — pass 1 of the SVD-Hybrid merge: one streaming read of base + N fine-tuned tensors
// (+ N tall masks) that forms the task vectors in registers, combines the tall masks, and
// accumulates the masked N x N Gram matrix of the task matrix, tile by tile.
//
// Replaces (paths relative to /root/reference):
//   compute_task_vector         src/svd_hybrid/task_vector_loader.py:142   (delta = ft - base)
//   combine_masks               src/svd_hybrid/mask_loader.py:412-485,604-609
//   apply_mask_to_tensor        src/svd_hybrid/mask_loader.py:675-679      (predicate, no compaction)
//   stack_and_center + the T^T T half of torch.linalg.svd   src/svd_hybrid/basis.py:103-111,241
//
// Layout: a "tile" is tile_elems consecutive elements of ONE parameter; one CTA per tile.
// Every tile writes its own partial Gram (G = N(N+1)/2 fp32) and masked count, so the result
// is deterministic and independent of how parameters are sharded over GPUs.  Bound: HBM
// (algorithmic bytes per element: (N+1)*sizeof(T) + N mask bytes + 1/8 packed-mask byte).
#include "svdq_kernels.h"

#ifndef SVDQ_DTYPE
#define SVDQ_DTYPE 0
#endif

namespace svdq {


template <typename T, int NT, bool FULL>
__global__ void __launch_bounds__(kBlock, (NT <= 8 && !FULL) ? 2 : 1) k1_tv_mask_gram(const K1Args a) {
    constexpr int G = tri_count(NT);
    constexpr int NACC = FULL ? 2 * G : G;
    constexpr int kRows = 32;                       // accumulator rows reduced per smem round
    __shared__ float red[kRows][kBlock + 1];
    __shared__ const void* s_ptr[NT + 1];
    __shared__ const uint8_t* s_mask[NT];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x;
    const int p = a.tile_param[tile];
    const int64_t numel = a.numel[p];
    const int64_t start = (int64_t)a.tile_local[tile] * a.tile_elems;
    const int64_t stop = min(start + (int64_t)a.tile_elems, numel);

    if (tid <= NT) {
        // a task that lacks the parameter reads the base tensor instead (delta == 0): the load
        // sequence stays branch-free; K2 drops the column through its `present` bit mask
        const void* q = a.tensors[(int64_t)p * (NT + 1) + tid];
        s_ptr[tid] = q ? q : a.tensors[(int64_t)p * (NT + 1)];
    }
    if (tid < NT) s_mask[tid] = a.masks ? a.masks[(int64_t)p * NT + tid] : nullptr;
    __syncthreads();

    int n_present = 0;
#pragma unroll
    for (int t = 0; t < NT; ++t) n_present += s_mask[t] != nullptr;
    const bool has_mask = n_present > 0;
    const bool all_masks = n_present == NT;
    const bool mask_bits = a.mask_bits != 0;
    // per-byte vote thresholds: union >= 1, intersection >= n_present, majority 2*votes >= n_present
    const uint32_t thr_bytes = 0x01010101u * (uint32_t)(a.strategy == kUnion ? 1 : n_present);
    const bool majority = a.strategy == kMajority;
    uint32_t* packed = has_mask ? a.packed + a.pmask_off[p] : nullptr;

    const bool comp = a.second_complement != 0;   // FULL: second block = Gram of the UNMASKED elements (noise region)
    // more than 10 tasks: the packed (float2) accumulators would not fit the register file -> plain fp32
    // accumulators and scalar FMAs (FULL above 10 tasks falls back to the packed accumulators and spills)
    constexpr bool kScalarAcc = NT > 10 && !FULL;
    float2 acc2[kScalarAcc ? 1 : G];
    float acc1[kScalarAcc ? G : 1];
#pragma unroll
    for (int i = 0; i < (kScalarAcc ? 1 : G); ++i) acc2[i] = make_float2(0.0f, 0.0f);
#pragma unroll
    for (int i = 0; i < (kScalarAcc ? G : 1); ++i) acc1[i] = 0.0f;
    uint32_t cnt = 0;

    for (int64_t e0 = start; e0 < stop; e0 += kStep) {          // uniform trip count per CTA
        const int64_t e = e0 + (int64_t)tid * kVec;
        const bool active = e < stop;
        const bool full = e + kVec <= numel;
        // ---- phase 1: issue every load of this step back to back (no use in between, so the
        //      (N+1) tensor loads and N mask loads of a thread are all in flight together) ----------
        float b[kVec], f[NT][kVec];
        uint32_t mw[NT];
        if (active && full) {
            Elem<T>::load4(s_ptr[0], e, b);
#pragma unroll
            for (int t = 0; t < NT; ++t) Elem<T>::load4(s_ptr[t + 1], e, f[t]);
            if (mask_bits) {                                     // bit-packed task masks (host-staged inputs)
#pragma unroll
                for (int t = 0; t < NT; ++t)
                    mw[t] = s_mask[t] ? nibble_to_bytes((__ldg(reinterpret_cast<const uint32_t*>(s_mask[t]) + (e >> 5))
                                                         >> (int)(e & 31)) & 0xFu) : 0u;
            } else if (all_masks) {
#pragma unroll
                for (int t = 0; t < NT; ++t) mw[t] = ldg_stream_u32(s_mask[t] + e);
            } else if (has_mask) {
#pragma unroll
                for (int t = 0; t < NT; ++t) mw[t] = s_mask[t] ? ldg_stream_u32(s_mask[t] + e) : 0u;
            }
        } else if (active) {                                     // last, partial vector of a parameter
#pragma unroll
            for (int c = 0; c < kVec; ++c) b[c] = (e + c < numel) ? Elem<T>::load1(s_ptr[0], e + c) : 0.0f;
#pragma unroll
            for (int t = 0; t < NT; ++t) {
#pragma unroll
                for (int c = 0; c < kVec; ++c) f[t][c] = (e + c < numel) ? Elem<T>::load1(s_ptr[t + 1], e + c) : 0.0f;
                mw[t] = 0;
                if (s_mask[t] != nullptr && mask_bits) {         // bits past numel are cut by `valid` below
                    mw[t] = nibble_to_bytes((__ldg(reinterpret_cast<const uint32_t*>(s_mask[t]) + (e >> 5))
                                             >> (int)(e & 31)) & 0xFu);
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
        // ---- phase 2: task vectors, combined mask ------------------------------------------------
        float d[NT][kVec];
#pragma unroll
        for (int t = 0; t < NT; ++t) Elem<T>::template subv<kVec>(f[t], b, d[t]);
        uint32_t bits = 0;                                       // 4 combined-mask bits of this thread
        if (active) {
            const uint32_t valid = full ? 0xFu : ((1u << (int)(numel - e)) - 1u);
            if (has_mask) {
                uint32_t votes = 0;                              // 4 byte lanes, one per element
#pragma unroll
                for (int t = 0; t < NT; ++t) votes += __vminu4(mw[t], 0x01010101u);
                if (majority) votes += votes;                    // 2 * votes (<= 64 per byte)
                const uint32_t ge = __vcmpgeu4(votes, thr_bytes);   // 0xff per byte where true
                bits = ((ge >> 7) & 1u) | ((ge >> 14) & 2u) | ((ge >> 21) & 4u) | ((ge >> 28) & 8u);
                bits &= valid;
            } else {
                bits = valid;
            }
        }
        cnt += __popc(bits);

        if (has_mask) {                 // eight lanes share one 32-bit word of the packed combined mask
            uint32_t w = bits << ((lane & 7) * 4);
            w |= __shfl_xor_sync(0xffffffffu, w, 1);
            w |= __shfl_xor_sync(0xffffffffu, w, 2);
            w |= __shfl_xor_sync(0xffffffffu, w, 4);
            if ((lane & 7) == 0 && e < stop) packed[e >> 5] = w;
        }

        // packed 2-wide FMAs (fma.rn.f32x2): FULL pairs the masked Gram with the all-element Gram of the
        // same (i, j) -- or, with second_complement, with the Gram of the unmasked elements; otherwise two consecutive elements share one instruction
        if constexpr (FULL) {
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
        } else if constexpr (kScalarAcc) {
#pragma unroll
            for (int c = 0; c < kVec; ++c) {
                const bool m = (bits >> c) & 1u;
                float v[NT];
#pragma unroll
                for (int t = 0; t < NT; ++t) v[t] = m ? d[t][c] : 0.0f;
#pragma unroll
                for (int i = 0; i < NT; ++i)
#pragma unroll
                    for (int j = i; j < NT; ++j)
                        acc1[tri_index(i, j, NT)] = fmaf(v[i], v[j], acc1[tri_index(i, j, NT)]);
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
        if constexpr (FULL) { acc[i] = acc2[i].x; acc[G + i] = acc2[i].y; }
        else if constexpr (kScalarAcc) acc[i] = acc1[i];
        else acc[i] = acc2[i].x + acc2[i].y;
    }
    // ---- CTA reduction in a fixed order: kRows accumulator rows per round through smem ----------
    float* gout = a.gram + (int64_t)tile * NACC;
#pragma unroll
    for (int r0 = 0; r0 < NACC; r0 += kRows) {
#pragma unroll
        for (int r = 0; r < kRows; ++r)
            if (r0 + r < NACC) red[r][tid] = acc[r0 + r];
        __syncthreads();
#pragma unroll
        for (int rr = 0; rr < kRows / (kBlock / 32); ++rr) {
            const int r = warp * (kRows / (kBlock / 32)) + rr;
            if (r0 + r < NACC) {
                float s = 0.0f;
#pragma unroll
                for (int c = 0; c < kBlock / 32; ++c) s += red[r][lane + 32 * c];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                if (lane == 0) gout[r0 + r] = s;
            }
        }
        __syncthreads();
    }
    // masked count
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    __shared__ uint32_t s_cnt[kBlock / 32];
    if (lane == 0) s_cnt[warp] = cnt;
    __syncthreads();
    if (tid == 0) {
        uint32_t c = 0;
#pragma unroll
        for (int w = 0; w < kBlock / 32; ++w) c += s_cnt[w];
        a.count[tile] = c;
    }
}

// ---- host-side dispatch --------------------------------------------------------------------------
template <typename T, int NT>
static cudaError_t launch_nt(const K1Args& a, int n_tiles, bool full, cudaStream_t st) {
    if (n_tiles <= 0) return cudaSuccess;
    if (full) k1_tv_mask_gram<T, NT, true><<<n_tiles, kBlock, 0, st>>>(a);
    else      k1_tv_mask_gram<T, NT, false><<<n_tiles, kBlock, 0, st>>>(a);
    return cudaGetLastError();
}

template <>
cudaError_t k1_launch_dtype<SVDQ_DTYPE>(int nt, const K1Args& a, int n_tiles, bool full, cudaStream_t st) {
    using T = DTypeOf<SVDQ_DTYPE>::type;
    switch (nt) {
#define SVDQ_CASE(N) case N: return launch_nt<T, N>(a, n_tiles, full, st);
        SVDQ_CASE(1) SVDQ_CASE(2) SVDQ_CASE(3) SVDQ_CASE(4) SVDQ_CASE(5) SVDQ_CASE(6) SVDQ_CASE(7) SVDQ_CASE(8)
        SVDQ_CASE(9) SVDQ_CASE(10) SVDQ_CASE(11) SVDQ_CASE(12) SVDQ_CASE(13) SVDQ_CASE(14) SVDQ_CASE(15) SVDQ_CASE(16)
#undef SVDQ_CASE
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace svdq
