// Shared device helpers for the SVD-Hybrid merge kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace svdq {

constexpr int kMaxTasks = 32;      // N <= 32 task vectors per merge
constexpr int kBlock = 256;        // threads per CTA of the streaming kernels
constexpr int kVec = 4;            // elements per thread per step (one 128-bit load for fp32)
constexpr int kStep = kBlock * kVec;   // elements one CTA covers per loop step (1024)
constexpr int kMaxStages = 8;      // RTVQ stages supported in-kernel

enum DType : int { kF32 = 0, kBF16 = 1, kF16 = 2 };
enum MaskStrategy : int { kUnion = 0, kIntersection = 1, kMajority = 2 };

__host__ __device__ constexpr int tri_count(int n) { return n * (n + 1) / 2; }
// packed index of (i, j), i <= j, row-major upper triangle
__host__ __device__ constexpr int tri_index(int i, int j, int n) { return i * n - i * (i - 1) / 2 + (j - i); }

// ---- streaming loads: read-once data bypasses L1 allocation -----------------------------
__device__ __forceinline__ float4 ldg_stream_f4(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ uint2 ldg_stream_u2(const void* p) {
    uint2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t ldg_stream_u32(const void* p) {
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void stg_stream_f4(float* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void stg_stream_f2(float* p, float2 v) {
    asm volatile("st.global.cs.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
}

// ---- element type adapters ------------------------------------------------------------------
// load4: four consecutive elements starting at element index e of tensor p (vector path needs
// 16 B (fp32) / 8 B (16-bit) alignment of p + e, guaranteed by the host for e % 4 == 0).
template <typename T> struct Elem;

template <> struct Elem<float> {
    static __device__ __forceinline__ void load4(const void* p, int64_t e, float (&out)[4]) {
        float4 v = ldg_stream_f4(reinterpret_cast<const float*>(p) + e);
        out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
    }
    static __device__ __forceinline__ float load1(const void* p, int64_t e) {
        return __ldg(reinterpret_cast<const float*>(p) + e);
    }
    // delta = finetuned - base in the tensors' own dtype (task_vector_loader.py:142)
    static __device__ __forceinline__ float sub(float f, float b) { return __fsub_rn(f, b); }
    template <int V> static __device__ __forceinline__ void subv(const float (&f)[V], const float (&b)[V], float (&d)[V]) {
#pragma unroll
        for (int c = 0; c < V; ++c) d[c] = __fsub_rn(f[c], b[c]);
    }
};

template <> struct Elem<__nv_bfloat16> {
    static __device__ __forceinline__ void load4(const void* p, int64_t e, float (&out)[4]) {
        uint2 v = ldg_stream_u2(reinterpret_cast<const __nv_bfloat16*>(p) + e);
        out[0] = __uint_as_float(v.x << 16); out[1] = __uint_as_float(v.x & 0xffff0000u);
        out[2] = __uint_as_float(v.y << 16); out[3] = __uint_as_float(v.y & 0xffff0000u);
    }
    static __device__ __forceinline__ float load1(const void* p, int64_t e) {
        return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[e]);
    }
    // bf16 - bf16 rounds the difference back to bf16
    static __device__ __forceinline__ float sub(float f, float b) {
        return __bfloat162float(__float2bfloat16_rn(__fsub_rn(f, b)));
    }
    // V (even) differences at once: the same fp32 subtract + round-to-nearest-even to bf16 as sub(), with the packed
    // convert (one F2FP on the ALU pipe per pair instead of two F2F on the quarter-rate XU pipe)
    template <int V> static __device__ __forceinline__ void subv(const float (&f)[V], const float (&b)[V], float (&d)[V]) {
        static_assert(V % 2 == 0, "pairs");
#pragma unroll
        for (int c = 0; c < V; c += 2) {
            const __nv_bfloat162 r = __floats2bfloat162_rn(__fsub_rn(f[c], b[c]), __fsub_rn(f[c + 1], b[c + 1]));
            const uint32_t w = *reinterpret_cast<const uint32_t*>(&r);
            d[c] = __uint_as_float(w << 16);
            d[c + 1] = __uint_as_float(w & 0xffff0000u);
        }
    }
};

template <> struct Elem<__half> {
    static __device__ __forceinline__ void load4(const void* p, int64_t e, float (&out)[4]) {
        uint2 v = ldg_stream_u2(reinterpret_cast<const __half*>(p) + e);
        __half2 a = *reinterpret_cast<__half2*>(&v.x), b = *reinterpret_cast<__half2*>(&v.y);
        float2 fa = __half22float2(a), fb = __half22float2(b);
        out[0] = fa.x; out[1] = fa.y; out[2] = fb.x; out[3] = fb.y;
    }
    static __device__ __forceinline__ float load1(const void* p, int64_t e) {
        return __half2float(reinterpret_cast<const __half*>(p)[e]);
    }
    static __device__ __forceinline__ float sub(float f, float b) {
        return __half2float(__float2half_rn(__fsub_rn(f, b)));
    }
    template <int V> static __device__ __forceinline__ void subv(const float (&f)[V], const float (&b)[V], float (&d)[V]) {
        static_assert(V % 2 == 0, "pairs");
#pragma unroll
        for (int c = 0; c < V; c += 2) {
            const float2 r = __half22float2(__floats2half2_rn(__fsub_rn(f[c], b[c]), __fsub_rn(f[c + 1], b[c + 1])));
            d[c] = r.x;
            d[c + 1] = r.y;
        }
    }
};

// round-trip through fp16 (the reference stores bases as .half() and upcasts again)
// two consecutive elements starting at an EVEN element index e (8- / 4-byte aligned vector access when `full`)
template <typename T> struct ElemPair;
template <> struct ElemPair<float> {
    static __device__ __forceinline__ void load(const void* p, int64_t e, bool full, int64_t numel, float (&o)[2]) {
        const float* q = reinterpret_cast<const float*>(p);
        if (full) {
            float2 v;
            asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(q + e));
            o[0] = v.x; o[1] = v.y;
        } else { o[0] = e < numel ? __ldg(q + e) : 0.0f; o[1] = e + 1 < numel ? __ldg(q + e + 1) : 0.0f; }
    }
};
template <> struct ElemPair<__nv_bfloat16> {
    static __device__ __forceinline__ void load(const void* p, int64_t e, bool full, int64_t numel, float (&o)[2]) {
        const __nv_bfloat16* q = reinterpret_cast<const __nv_bfloat16*>(p);
        if (full) {
            const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(q + e));
            o[0] = __uint_as_float(v << 16); o[1] = __uint_as_float(v & 0xffff0000u);
        } else { o[0] = e < numel ? __bfloat162float(q[e]) : 0.0f; o[1] = e + 1 < numel ? __bfloat162float(q[e + 1]) : 0.0f; }
    }
};
template <> struct ElemPair<__half> {
    static __device__ __forceinline__ void load(const void* p, int64_t e, bool full, int64_t numel, float (&o)[2]) {
        const __half* q = reinterpret_cast<const __half*>(p);
        if (full) {
            const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(q + e));
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&v));
            o[0] = f.x; o[1] = f.y;
        } else { o[0] = e < numel ? __half2float(q[e]) : 0.0f; o[1] = e + 1 < numel ? __half2float(q[e + 1]) : 0.0f; }
    }
};


// four mask bits (bit c = element c) -> four 0/1 bytes in the lanes of a 32-bit word (the layout torch.bool gives)
__device__ __forceinline__ uint32_t nibble_to_bytes(uint32_t nib) { return (nib * 0x00204081u) & 0x01010101u; }
__device__ __forceinline__ float round_fp16(float x) { return __half2float(__float2half_rn(x)); }

}  // namespace svdq
