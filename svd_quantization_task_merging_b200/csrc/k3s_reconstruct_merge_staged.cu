// K3 (staged) — pass 2 (weighted reconstruction + merge) as a PERSISTENT, warp-specialised kernel:
// one CTA per SM, a producer warp streams 1024-element chunks of the (N+1) tensors global -> shared
// with the TMA bulk-copy engine (cp.async.bulk + mbarrier complete_tx) into a STAGES-deep ring,
// TWO groups of 8 consumer warps take alternate chunks (pass 2 carries ~100 instructions per element,
// so it needs 16 resident warps to keep the FMA pipe fed while the ring keeps HBM busy), run
// svdq::k3_step out of shared memory and write the merged values with streaming 128-bit stores.  Bit-identical to k3_reconstruct_merge.cu (same per-thread element assignment and
// arithmetic).  Tail chunks (< 1024 elements) are loaded directly by the consumers.
#include "k3_body.cuh"
#include "stage_pipe.cuh"

#ifndef SVDQ_DTYPE
#define SVDQ_DTYPE 0
#endif

namespace svdq {

constexpr int kK3Stages = 5;
constexpr int kK3Groups = 2;                       // consumer groups of kBlock threads
constexpr int kK3Consumers = kK3Groups * kBlock;

template <typename T> struct SmemElem3;
template <> struct SmemElem3<float> {
    static __device__ __forceinline__ void load4(const unsigned char* slot, int tid, float (&o)[4]) {
        const float4 v = *reinterpret_cast<const float4*>(slot + tid * 16);
        o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
    }
};
template <> struct SmemElem3<__nv_bfloat16> {
    static __device__ __forceinline__ void load4(const unsigned char* slot, int tid, float (&o)[4]) {
        const uint2 v = *reinterpret_cast<const uint2*>(slot + tid * 8);
        o[0] = __uint_as_float(v.x << 16); o[1] = __uint_as_float(v.x & 0xffff0000u);
        o[2] = __uint_as_float(v.y << 16); o[3] = __uint_as_float(v.y & 0xffff0000u);
    }
};
template <> struct SmemElem3<__half> {
    static __device__ __forceinline__ void load4(const unsigned char* slot, int tid, float (&o)[4]) {
        const uint2 v = *reinterpret_cast<const uint2*>(slot + tid * 8);
        const __half2 a = *reinterpret_cast<const __half2*>(&v.x), b = *reinterpret_cast<const __half2*>(&v.y);
        const float2 fa = __half22float2(a), fb = __half22float2(b);
        o[0] = fa.x; o[1] = fa.y; o[2] = fb.x; o[3] = fb.y;
    }
};

template <typename T, int NT>
__host__ __device__ constexpr int k3s_smem_bytes() {
    return kK3Stages * ((NT + 1) * kStep * (int)sizeof(T) + kStep / 8) + 2 * kK3Stages * 8 + kK3Stages * 4 + 64;
}

template <typename T, int NT, bool FP16B>
__global__ void __launch_bounds__(kK3Consumers + 32, 1) k3s_reconstruct_merge(const K3Args a, const int n_tiles) {
    constexpr int NTP = (NT + 3) & ~3;
    constexpr int STAGES = kK3Stages;
    constexpr int kTensorBytes = kStep * (int)sizeof(T);
    constexpr int kMaskBytes = kStep / 8;                          // 1024 mask bits of the chunk
    constexpr int kStageBytes = (NT + 1) * kTensorBytes + kMaskBytes;
    constexpr int kMaskOff = (NT + 1) * kTensorBytes;
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* stage_base = smem;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * kStageBytes);
    uint64_t* empty = full + STAGES;
    int* s_direct = reinterpret_cast<int*>(empty + STAGES);
    __shared__ __align__(16) float sWT[NT][NTP];
    __shared__ __align__(16) float sChatT[1][NTP];
    __shared__ float sCbar[NT], sG[NT];
    __shared__ const void* s_ptr[NT + 1];
    __shared__ uint32_t s_present;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int group = warp / (kBlock / 32);              // consumer group (kK3Groups = producer warp)
    const int tid = threadIdx.x % kBlock;                // thread index inside its consumer group
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kBlock / 32); }
        mbar_fence_init();
    }
    __syncthreads();

    if (group == kK3Groups) {
        // ================= producer warp: lane i issues the bulk copies of tensor i ===================
        PipeState ps;
        const bool is_tensor = lane <= NT;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int p = a.tile_param[tile];
            const int64_t numel = a.numel[p];
            const int64_t start = (int64_t)a.tile_local[tile] * a.tile_elems;
            const int64_t stop = min(start + (int64_t)a.tile_elems, numel);
            const bool solved = a.info[(int64_t)p * 8] == kSolved;
            // the mask words ride the ring only when their row starts 16-byte aligned (TMA source alignment);
            // otherwise the consumers read them directly
            const bool has_mask = a.has_mask[p] != 0 &&
                                  ((reinterpret_cast<uintptr_t>(a.packed + a.pmask_off[p]) & 15u) == 0);
            const unsigned char* my_ptr = nullptr;
            if (is_tensor) {
                const void* q = a.tensors[(int64_t)p * (NT + 1) + lane];
                my_ptr = reinterpret_cast<const unsigned char*>(q ? q : a.tensors[(int64_t)p * (NT + 1)]);
            } else if (lane == NT + 1 && has_mask) {
                my_ptr = reinterpret_cast<const unsigned char*>(a.packed + a.pmask_off[p]);   // 16-byte aligned rows
            }
            for (int64_t e0 = start; e0 < stop; e0 += kStep) {
                if (lane == 0) mbar_wait(&empty[ps.stage], ps.phase ^ 1u);
                __syncwarp();
                unsigned char* sb = stage_base + ps.stage * kStageBytes;
                if (e0 + kStep <= numel && solved) {
                    if (lane == 0) {
                        s_direct[ps.stage] = 0;
                        mbar_arrive_expect_tx(&full[ps.stage],
                                              (uint32_t)((NT + 1) * kTensorBytes + (has_mask ? kMaskBytes : 0)));
                    }
                    __syncwarp();
                    if (is_tensor)
                        bulk_g2s(sb + lane * kTensorBytes, my_ptr + e0 * (int64_t)sizeof(T), kTensorBytes, &full[ps.stage]);
                    else if (my_ptr != nullptr)
                        bulk_g2s(sb + kMaskOff, my_ptr + e0 / 8, kMaskBytes, &full[ps.stage]);
                } else if (lane == 0) {
                    s_direct[ps.stage] = 1;          // tail chunk, or a parameter without a basis (only base is read)
                    mbar_arrive(&full[ps.stage]);
                }
                ps.advance<STAGES>();
            }
        }
        return;
    }

    // ======================= consumers ==================================================================
    PipeState ps;
    int p_prev = -1;
    uint32_t chunk = 0;                                  // running chunk counter: group g owns chunks with chunk % 2 == g
    float dacc[1] = {0.0f};
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int p = a.tile_param[tile];
        const int64_t numel = a.numel[p];
        const int64_t start = (int64_t)a.tile_local[tile] * a.tile_elems;
        const int64_t stop = min(start + (int64_t)a.tile_elems, numel);
        if (p != p_prev) {                                   // per-parameter constants (uniform branch)
            named_bar_sync(1, kK3Consumers);
            if (group == 0 && tid <= NT) {
                const void* q = a.tensors[(int64_t)p * (NT + 1) + tid];
                s_ptr[tid] = q ? q : a.tensors[(int64_t)p * (NT + 1)];
            }
            if (threadIdx.x == 0) {
                uint32_t pb = 0;
                for (int t = 0; t < NT; ++t) pb |= (a.tensors[(int64_t)p * (NT + 1) + 1 + t] != nullptr ? 1u : 0u) << t;
                s_present = pb;
            }
            for (int i = threadIdx.x; i < NT * NTP; i += kK3Consumers) {
                const int j = i / NTP, t = i % NTP;
                sWT[j][t] = (t < NT) ? a.W[(int64_t)p * NT * NT + t * NT + j] : 0.0f;
            }
            if (group == 0 && tid < NT) { sCbar[tid] = a.cbar[(int64_t)p * NT + tid]; sG[tid] = a.gvec[(int64_t)p * NT + tid]; }
            named_bar_sync(1, kK3Consumers);
            p_prev = p;
        }
        const int status = a.info[(int64_t)p * 8 + 0];
        const int n_active = a.info[(int64_t)p * 8 + 1];
        // columns beyond r_eff are zero (null directions): their contribution is tail_add; the fused
    // diagnostics still walk all r columns so that a NaN coefficient shows up exactly as in the reference
    const int r = a.info[(int64_t)p * 8 + (false ? 2 : 4)];
        const float tail_add = a.scal[(int64_t)p * 4 + 1];
        const float mean_scale = a.scal[(int64_t)p * 4 + 2];
        const bool has_mask = a.has_mask[p] != 0;
        const uint32_t* packed = has_mask ? a.packed + a.pmask_off[p] : nullptr;
        const bool mask_in_ring = has_mask && ((reinterpret_cast<uintptr_t>(packed) & 15u) == 0);
        float* outp = a.out[p];
        const float n_f = (float)(n_active > 0 ? n_active : 1);
        const uint32_t present_bits = s_present;

        for (int64_t e0 = start; e0 < stop; e0 += kStep, ++chunk) {
            if ((int)(chunk % kK3Groups) != group) {         // the other group's chunk
                ps.advance<STAGES>();
                continue;
            }
            const int64_t e = e0 + (int64_t)tid * kVec;
            const bool active = e < stop;
            const bool fullv = e + kVec <= numel;
            float b[kVec], x[NT][kVec], res[kVec];
            uint32_t pword = 0xffffffffu;
            mbar_wait(&full[ps.stage], ps.phase);
            const unsigned char* sb = stage_base + ps.stage * kStageBytes;
            bool staged_mask = false;
            if (!s_direct[ps.stage]) {
                SmemElem3<T>::load4(sb, tid, b);
#pragma unroll
                for (int t = 0; t < NT; ++t) SmemElem3<T>::load4(sb + (t + 1) * kTensorBytes, tid, x[t]);
                if (mask_in_ring) pword = *reinterpret_cast<const uint32_t*>(sb + kMaskOff + (tid >> 3) * 4);
                staged_mask = mask_in_ring || !has_mask;
            } else if (active) {
                if (fullv) Elem<T>::load4(s_ptr[0], e, b);
                else {
#pragma unroll
                    for (int c = 0; c < kVec; ++c) b[c] = (e + c < numel) ? Elem<T>::load1(s_ptr[0], e + c) : 0.0f;
                }
                if (status == kSolved) {
                    if (fullv) {
#pragma unroll
                        for (int t = 0; t < NT; ++t) Elem<T>::load4(s_ptr[t + 1], e, x[t]);
                    } else {
#pragma unroll
                        for (int t = 0; t < NT; ++t)
#pragma unroll
                            for (int c = 0; c < kVec; ++c)
                                x[t][c] = (e + c < numel) ? Elem<T>::load1(s_ptr[t + 1], e + c) : 0.0f;
                    }
                }
            }
            if (!staged_mask && active && has_mask && status == kSolved) pword = __ldg(packed + (e >> 5));
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[ps.stage]);
            ps.advance<STAGES>();
            if (!active) continue;

            if (status != kSolved) {
#pragma unroll
                for (int c = 0; c < kVec; ++c) res[c] = b[c];
            } else {
                k3_step<T, NT, FP16B, false>(b, x, pword, e, numel, r, present_bits, a.center, n_f, tail_add, mean_scale, sWT, sChatT,
                                             sCbar, sG, res, dacc);
            }
            if (fullv) stg_stream_f4(outp + e, make_float4(res[0], res[1], res[2], res[3]));
            else {
#pragma unroll
                for (int c = 0; c < kVec; ++c)
                    if (e + c < numel) outp[e + c] = res[c];
            }
        }
    }
}

template <typename T, int NT>
static cudaError_t launch_staged(const K3Args& a, int n_tiles, bool fp16b, int n_sm, cudaStream_t st) {
    if (n_tiles <= 0) return cudaSuccess;
    constexpr int smem = k3s_smem_bytes<T, NT>();
    const int grid = n_tiles < n_sm ? n_tiles : n_sm;
    cudaError_t e;
    if (fp16b) {
        e = cudaFuncSetAttribute(k3s_reconstruct_merge<T, NT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        k3s_reconstruct_merge<T, NT, true><<<grid, kK3Consumers + 32, smem, st>>>(a, n_tiles);
    } else {
        e = cudaFuncSetAttribute(k3s_reconstruct_merge<T, NT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        k3s_reconstruct_merge<T, NT, false><<<grid, kK3Consumers + 32, smem, st>>>(a, n_tiles);
    }
    return cudaGetLastError();
}

// staged path: nt <= 8, no fused diagnostics (the DIAG variant stays on the direct-load kernel)
template <>
cudaError_t k3s_launch_dtype<SVDQ_DTYPE>(int nt, const K3Args& a, int n_tiles, bool fp16b, int n_sm, cudaStream_t st) {
    using T = DTypeOf<SVDQ_DTYPE>::type;
    switch (nt) {
#define SVDQ_CASE(N) case N: return launch_staged<T, N>(a, n_tiles, fp16b, n_sm, st);
        SVDQ_CASE(1) SVDQ_CASE(2) SVDQ_CASE(3) SVDQ_CASE(4) SVDQ_CASE(5) SVDQ_CASE(6) SVDQ_CASE(7) SVDQ_CASE(8)
#undef SVDQ_CASE
        default: return cudaErrorNotSupported;
    }
}

}  // namespace svdq
