// Copy of a shared-memory image of an artifact slice to global memory with 16-byte coalesced stores, shared by the
// basis writers (k5_basis_misc.cu, k3c_merge_diag_compact.cu).
#pragma once
#include "svdq_common.cuh"

namespace svdq {

template <typename OUT> struct OutCvt;
template <> struct OutCvt<__half> { static __device__ __forceinline__ __half cvt(float v) { return __float2half_rn(v); } };
template <> struct OutCvt<float> { static __device__ __forceinline__ float cvt(float v) { return v; } };

template <typename OUT>
__device__ __forceinline__ void k5_copy_out(OUT* __restrict__ dst, const OUT* __restrict__ src, int64_t g0, int n,
                                            int tid) {
    // dst + g0 .. + n  <-  src[0 .. n);  src is placed so that (src address) == (dst + g0 address) mod 16
    constexpr int V = 16 / (int)sizeof(OUT);
    const int head = min(n, (int)((V - (g0 % V)) % V));
    if (tid < head) dst[g0 + tid] = src[tid];
    const int nvec = (n - head) / V;
    const uint4* s4 = reinterpret_cast<const uint4*>(src + head);
    uint4* d4 = reinterpret_cast<uint4*>(dst + g0 + head);
    for (int i = tid; i < nvec; i += kBlock) d4[i] = s4[i];
    const int done = head + nvec * V;
    if (tid < n - done) dst[g0 + done + tid] = src[done + tid];
}

}  // namespace svdq
