// K8 — single-pass masked Gram for 9..32 task vectors (the standard 14- and 20-task merges).
//
// With more than 8 tasks the N(N+1)/2 accumulators of the Gram no longer fit one thread, and the first version of
// the wide path re-read the inputs once per pair of 8-task blocks (2-3x the traffic of pass 1).  Here the inputs
// are read from HBM exactly once: a persistent CTA per SM streams chunks of C elements of the (N+1) tensors into a
// shared-memory ring with the TMA bulk-copy engine, all 16 warps first turn the chunk into masked task vectors
// d_t = (ft_t - base) * m (fp32, [N][C] in shared memory; same rounding as ft - base on the tensors' dtype), and
// then the 8x8 task BLOCKS of the Gram are spread over the warps: a warp owns one block pair (bi, bj) for the whole
// kernel -- its 64 (36 on the diagonal) accumulators never leave registers -- and walks its share of the chunk's
// elements two at a time.  Accumulators are flushed once per tile in a fixed order, so the per-tile partial Grams
// (same [n_tiles][N(N+1)/2] layout as K1, consumed by k2_gram_reduce) are deterministic and independent of which
// CTA ran the tile.  The combined mask comes pre-packed from k6_mask_pack; mask_mode selects the rows (0 = inside
// the mask, 1 = all, 2 = outside) exactly as in k1_tv_mask_gram's pre-combined mode.
//
// Replaces the same reference lines as K1: compute_task_vector (src/svd_hybrid/task_vector_loader.py:142),
// apply_mask_to_tensor (src/svd_hybrid/mask_loader.py:675-679), stack_and_center + the T^T T half of
// torch.linalg.svd (src/svd_hybrid/basis.py:103-111,241).  Bound: HBM ((N+1)*sizeof(T) B per element) against
// ~N^2/2 FMAs per element from shared memory; at N = 20 both are within ~20 % of each other.
#include "svdq_kernels.h"
#include "stage_pipe.cuh"

#ifndef SVDQ_DTYPE
#define SVDQ_DTYPE 0
#endif

namespace svdq {

constexpr int kK8Threads = 512;
constexpr int kK8Warps = kK8Threads / 32;
constexpr int kK8Blk = 8;                                   // tasks per block

template <int NB> struct K8Cfg {
    static constexpr int kChunk = NB == 4 ? 512 : 768;      // elements per chunk (tile_elems must be a multiple)
    static constexpr int kPasses = kChunk / 64;             // a pass = 64 elements = 2 per lane
    static constexpr int kRows = NB * kK8Blk;               // rows of the task-vector buffer
};

// warp -> (block pair, slice, slices of the pair).  Diagonal pairs cost 36 FMAs per element, off-diagonal 64.
template <int NB>
__device__ __forceinline__ bool k8_pair_of_warp(int w, int& bi, int& bj, int& slice, int& n_slices) {
    if (NB == 2) {            // (0,0) x3, (1,1) x3, (0,1) x6
        if (w < 6) { bi = bj = w / 3; slice = w % 3; n_slices = 3; return true; }
        if (w < 12) { bi = 0; bj = 1; slice = w - 6; n_slices = 6; return true; }
        return false;
    }
    if (NB == 3) {            // three diagonal pairs x2, three off-diagonal pairs x3
        if (w < 6) { bi = bj = w / 2; slice = w % 2; n_slices = 2; return true; }
        if (w < 15) {
            const int o = (w - 6) / 3;
            bi = o == 2 ? 1 : 0; bj = o == 0 ? 1 : 2; slice = (w - 6) % 3; n_slices = 3; return true;
        }
        return false;
    }
    // NB == 4: four diagonal pairs x1, six off-diagonal pairs x2
    if (w < 4) { bi = bj = w; slice = 0; n_slices = 1; return true; }
    const int o = (w - 4) / 2;
    constexpr int kBi[6] = {0, 0, 0, 1, 1, 2}, kBj[6] = {1, 2, 3, 2, 3, 3};
    bi = kBi[o]; bj = kBj[o]; slice = (w - 4) % 2; n_slices = 2;
    return true;
}
// (block pair) -> first warp and number of warps, the inverse of the above
template <int NB>
__device__ __forceinline__ void k8_warps_of_pair(int bi, int bj, int& first, int& count) {
    if (NB == 2) {
        if (bi == bj) { first = bi * 3; count = 3; } else { first = 6; count = 6; }
    } else if (NB == 3) {
        if (bi == bj) { first = bi * 2; count = 2; }
        else { const int o = bi == 1 ? 2 : (bj == 1 ? 0 : 1); first = 6 + o * 3; count = 3; }
    } else {
        if (bi == bj) { first = bi; count = 1; }
        else {
            const int o = bi == 0 ? bj - 1 : (bi == 1 ? bj + 1 : 5);
            first = 4 + o * 2; count = 2;
        }
    }
}

// four consecutive staged elements (quad q of a tensor's chunk) as fp32
template <typename T> struct K8Quad;
template <> struct K8Quad<float> {
    static __device__ __forceinline__ void load(const unsigned char* slot, int q, float (&o)[4]) {
        const float4 v = *reinterpret_cast<const float4*>(slot + q * 16);
        o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
    }
};
template <> struct K8Quad<__nv_bfloat16> {
    static __device__ __forceinline__ void load(const unsigned char* slot, int q, float (&o)[4]) {
        const uint2 v = *reinterpret_cast<const uint2*>(slot + q * 8);
        o[0] = __uint_as_float(v.x << 16); o[1] = __uint_as_float(v.x & 0xffff0000u);
        o[2] = __uint_as_float(v.y << 16); o[3] = __uint_as_float(v.y & 0xffff0000u);
    }
};
template <> struct K8Quad<__half> {
    static __device__ __forceinline__ void load(const unsigned char* slot, int q, float (&o)[4]) {
        const uint2 v = *reinterpret_cast<const uint2*>(slot + q * 8);
        const float2 fa = __half22float2(*reinterpret_cast<const __half2*>(&v.x));
        const float2 fb = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
        o[0] = fa.x; o[1] = fa.y; o[2] = fb.x; o[3] = fb.y;
    }
};

template <typename T, int NB>
__global__ void __launch_bounds__(kK8Threads, 1) k8_gram_staged(const K1Args a, const int n_tasks, const int n_tiles,
                                                                const int n_stages) {
    using Cfg = K8Cfg<NB>;
    constexpr int C = Cfg::kChunk;
    constexpr int kRows = Cfg::kRows;
    constexpr int kTensorBytes = C * (int)sizeof(T);
    const int N = n_tasks;
    const int G = tri_count(N);
    const int stage_bytes = (N + 1) * kTensorBytes;
    extern __shared__ __align__(128) unsigned char smem[];
    float* dbuf = reinterpret_cast<float*>(smem);                                  // [kRows][C] masked task vectors
    float* scratch = dbuf + kRows * C;                                        // [kK8Warps][64] flush partials
    unsigned char* ring = reinterpret_cast<unsigned char*>(scratch + kK8Warps * 64);   // n_stages x stage_bytes
    __shared__ uint64_t full[4];
    __shared__ int s_direct[4];
    __shared__ const unsigned char* s_ptr[kMaxTasks + 1];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < n_stages; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    for (int i = tid; i < kRows * C; i += kK8Threads) dbuf[i] = 0.0f;              // rows >= N stay zero for good
    __syncthreads();

    int bi = 0, bj = 0, slice = 0, n_slices = 1;
    const bool has_pair = k8_pair_of_warp<NB>(warp, bi, bj, slice, n_slices);
    const bool diag = bi == bj;
    float acc[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) acc[i] = 0.0f;

    // ---- chunk cursors: the (tile, offset) sequence of this CTA.  Tile facts are re-read only when the tile
    //      changes; the load cursor (last warp only) runs n_stages chunks ahead of the compute cursor ------------
    struct Cursor {
        int tile, p;
        int64_t off, stop, numel;                       // off = element offset of the chunk inside its parameter
    };
    auto enter_tile = [&](Cursor& c) {
        if (c.tile >= n_tiles) return;
        c.p = a.tile_param[c.tile];
        c.numel = a.numel[c.p];
        c.off = (int64_t)a.tile_local[c.tile] * a.tile_elems;
        c.stop = min(c.off + (int64_t)a.tile_elems, c.numel);
    };
    auto advance = [&](Cursor& c) {                     // next chunk (tile >= n_tiles when exhausted)
        c.off += C;
        if (c.off >= c.stop) { c.tile += gridDim.x; enter_tile(c); }
    };
    // the LAST warp starts the loads of a chunk (lane t copies tensor t); in the 17..24-task configuration that
    // warp owns no block pair, so the issue never delays a compute warp
    auto issue = [&](const Cursor& c, int k) {
        const int stage = k % n_stages;
        if (c.off + C <= c.numel) {
            if (lane == 0) {
                s_direct[stage] = 0;
                mbar_arrive_expect_tx(&full[stage], (uint32_t)stage_bytes);
            }
            __syncwarp();
            unsigned char* sb = ring + (size_t)stage * stage_bytes;
            const void* const* tp = a.tensors + (int64_t)c.p * (N + 1);
            const unsigned char* base = reinterpret_cast<const unsigned char*>(tp[0]);
            for (int t = lane; t <= N; t += 32) {
                const unsigned char* src = tp[t] ? reinterpret_cast<const unsigned char*>(tp[t]) : base;
                bulk_g2s(sb + (size_t)t * kTensorBytes, src + c.off * (int64_t)sizeof(T), kTensorBytes, &full[stage]);
            }
        } else if (lane == 0) {
            s_direct[stage] = 1;                        // last, partial chunk of a parameter: read directly
            mbar_arrive(&full[stage]);
        }
    };
    const bool issuer = warp == kK8Warps - 1;

    Cursor cc{(int)blockIdx.x, 0, 0, 0, 0};
    enter_tile(cc);
    Cursor lc = cc;
    int k_load = 0;
    if (issuer) {
        for (; k_load < n_stages && lc.tile < n_tiles; ++k_load) { issue(lc, k_load); advance(lc); }
    }

    // Per chunk: phase A (all warps) fills the task-vector buffer, barrier, phase B (block pairs) reads it, barrier.
    // (A double-buffered variant with half-size chunks and one barrier per chunk was measured and was slower: the
    // per-chunk fixed costs outweigh the hidden skew.)
    int p_prev = -1;
    bool has_mask = false;
    const uint32_t* packed = nullptr;
    for (int k = 0; cc.tile < n_tiles; ++k) {
        const int stage = k % n_stages;
        const uint32_t parity = (uint32_t)(k / n_stages) & 1u;
        const int p = cc.p;
        const int64_t numel = cc.numel, e0 = cc.off;
        if (p != p_prev) {                              // per-parameter facts (uniform branch)
            if (tid <= N) {
                const void* q = a.tensors[(int64_t)p * (N + 1) + tid];
                s_ptr[tid] = reinterpret_cast<const unsigned char*>(q ? q : a.tensors[(int64_t)p * (N + 1)]);
            }
            has_mask = a.has_mask_in[p] != 0;
            packed = has_mask ? a.packed_in + a.pmask_off[p] : nullptr;
            p_prev = p;
        }
        // phase A work split: a thread owns four consecutive elements (one 128-bit shared-memory access per task)
        // and every TG-th task; its four mask bits come from one packed word, fetched before the wait
        constexpr int Q = C / 4;                         // element quads per chunk
        constexpr int TG = kK8Threads / Q;               // task groups (2 for C = 768, 4 for C = 512)
        const int q = tid % Q, tg = tid / Q;
        const int64_t eq = e0 + 4 * q;
        uint32_t keep4 = 0;                              // bit c: element eq + c enters this Gram
        if (tg < TG && eq < numel) {
            const int64_t left = numel - eq;
            const uint32_t valid = left >= 4 ? 0xFu : ((1u << (int)left) - 1u);
            uint32_t bits = 0xFu;
            if (a.mask_mode != 1) {
                const uint32_t w = has_mask ? (__ldg(packed + (eq >> 5)) >> (int)(eq & 31)) & 0xFu : 0xFu;
                bits = a.mask_mode == 2 ? (has_mask ? ~w & 0xFu : 0u) : w;
            }
            keep4 = bits & valid;
        }
        mbar_wait(&full[stage], parity);
        const bool direct = s_direct[stage] != 0;
        if (direct) __syncthreads();                     // s_ptr of a new parameter must be visible to direct reads
        // ---- phase A: raw chunk -> masked task vectors ----------------------------------------------------------
        const unsigned char* sb = ring + (size_t)stage * stage_bytes;
        if (tg < TG) {
            float b[4];
            if (!direct) K8Quad<T>::load(sb, q, b);
            else {
#pragma unroll
                for (int c = 0; c < 4; ++c) b[c] = (eq + c < numel) ? Elem<T>::load1(s_ptr[0], eq + c) : 0.0f;
            }
#pragma unroll 4
            for (int t = tg; t < N; t += TG) {
                float f[4];
                if (!direct) K8Quad<T>::load(sb + (size_t)(t + 1) * kTensorBytes, q, f);
                else {
#pragma unroll
                    for (int c = 0; c < 4; ++c) f[c] = (eq + c < numel) ? Elem<T>::load1(s_ptr[t + 1], eq + c) : b[c];
                }
                float4 d;
                d.x = (keep4 & 1u) ? Elem<T>::sub(f[0], b[0]) : 0.0f;
                d.y = (keep4 & 2u) ? Elem<T>::sub(f[1], b[1]) : 0.0f;
                d.z = (keep4 & 4u) ? Elem<T>::sub(f[2], b[2]) : 0.0f;
                d.w = (keep4 & 8u) ? Elem<T>::sub(f[3], b[3]) : 0.0f;
                *reinterpret_cast<float4*>(dbuf + t * C + 4 * q) = d;
            }
        }
        __syncthreads();                                 // task vectors complete; the raw stage is free again
        if (issuer && lc.tile < n_tiles) { issue(lc, k_load); advance(lc); ++k_load; }

        // ---- phase B: this warp's block pair over its share of the chunk ------------------------------------------
        if (has_pair) {
            const float* di = dbuf + bi * kK8Blk * C;
            const float* dj = dbuf + bj * kK8Blk * C;
            for (int pass = slice; pass < Cfg::kPasses; pass += n_slices) {
                const int el = pass * 64 + lane * 2;
                float2 xi[kK8Blk], xj[kK8Blk];
#pragma unroll
                for (int t = 0; t < kK8Blk; ++t) xi[t] = *reinterpret_cast<const float2*>(di + t * C + el);
                // the two products of an accumulator are issued a whole block apart (no back-to-back dependent FMAs)
                if (diag) {
#pragma unroll
                    for (int i = 0; i < kK8Blk; ++i)
#pragma unroll
                        for (int j = i; j < kK8Blk; ++j) acc[i * kK8Blk + j] = fmaf(xi[i].x, xi[j].x, acc[i * kK8Blk + j]);
#pragma unroll
                    for (int i = 0; i < kK8Blk; ++i)
#pragma unroll
                        for (int j = i; j < kK8Blk; ++j) acc[i * kK8Blk + j] = fmaf(xi[i].y, xi[j].y, acc[i * kK8Blk + j]);
                } else {
#pragma unroll
                    for (int t = 0; t < kK8Blk; ++t) xj[t] = *reinterpret_cast<const float2*>(dj + t * C + el);
#pragma unroll
                    for (int i = 0; i < kK8Blk; ++i)
#pragma unroll
                        for (int j = 0; j < kK8Blk; ++j) acc[i * kK8Blk + j] = fmaf(xi[i].x, xj[j].x, acc[i * kK8Blk + j]);
#pragma unroll
                    for (int i = 0; i < kK8Blk; ++i)
#pragma unroll
                        for (int j = 0; j < kK8Blk; ++j) acc[i * kK8Blk + j] = fmaf(xi[i].y, xj[j].y, acc[i * kK8Blk + j]);
                }
            }
        }

        // ---- next chunk; flush the accumulators when the tile ends (once per tile_elems / C chunks) ----------------
        const int tile_done = cc.tile;
        advance(cc);
        const bool tile_end = cc.tile != tile_done;
        if (tile_end) {
#pragma unroll
            for (int i = 0; i < 64; ++i) {
                float v = acc[i];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0) scratch[warp * 64 + i] = v;
                acc[i] = 0.0f;
            }
        }
        __syncthreads();                                 // task-vector buffer free again; flush partials complete
        if (tile_end) {
            float* gout = a.gram + (int64_t)tile_done * G;
            for (int idx = tid; idx < G; idx += kK8Threads) {
                int i = 0, rem = idx;                    // unpack the upper triangle (row-major, as tri_index)
                while (rem >= N - i) { rem -= N - i; ++i; }
                const int j = i + rem;
                int first, count;
                k8_warps_of_pair<NB>(i / kK8Blk, j / kK8Blk, first, count);
                const int slot = (i % kK8Blk) * kK8Blk + (j % kK8Blk);
                float sum = 0.0f;
                for (int w = 0; w < count; ++w) sum += scratch[(first + w) * 64 + slot];
                gout[idx] = sum;
            }
            // scratch is rewritten at the next flush, at least one chunk barrier away
        }
    }
}

template <typename T, int NB>
static cudaError_t k8_go(const K1Args& a, int n_tasks, int n_tiles, int n_sm, cudaStream_t st) {
    using Cfg = K8Cfg<NB>;
    if (a.tile_elems % Cfg::kChunk != 0) return cudaErrorInvalidValue;
    const size_t fixed = (size_t)Cfg::kRows * Cfg::kChunk * 4 + kK8Warps * 64 * 4;
    const size_t stage = (size_t)(n_tasks + 1) * Cfg::kChunk * sizeof(T);
    const size_t budget = 232448 - 1024;                 // 227 KB per CTA minus the static shared memory
    int stages = (int)((budget - fixed) / stage);
    if (stages > 4) stages = 4;
    if (stages < 2) return cudaErrorInvalidValue;
    const size_t dsm = fixed + stages * stage;
    cudaError_t e = cudaFuncSetAttribute(k8_gram_staged<T, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm);
    if (e != cudaSuccess) return e;
    const int grid = n_tiles < n_sm ? n_tiles : n_sm;
    k8_gram_staged<T, NB><<<grid, kK8Threads, dsm, st>>>(a, n_tasks, n_tiles, stages);
    return cudaGetLastError();
}

template <>
cudaError_t k8_launch_dtype<SVDQ_DTYPE>(int n_tasks, const K1Args& a, int n_tiles, int n_sm, cudaStream_t st) {
    using T = DTypeOf<SVDQ_DTYPE>::type;
    if (n_tiles <= 0) return cudaSuccess;
    if (n_tasks < 1 || n_tasks > kMaxTasks) return cudaErrorInvalidValue;
    if (n_tasks <= 16) return k8_go<T, 2>(a, n_tasks, n_tiles, n_sm, st);
    if (n_tasks <= 24) return k8_go<T, 3>(a, n_tasks, n_tiles, n_sm, st);
    return k8_go<T, 4>(a, n_tasks, n_tiles, n_sm, st);
}

}  // namespace svdq
