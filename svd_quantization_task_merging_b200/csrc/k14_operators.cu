// K14 — the operator-by-operator API of the reference's artifact tools on the device (no library calls):
//
//   k14_project        c = U^T (delta - mean) for one stored basis block of up to 32 columns
//                      (project_to_basis, src/svd_hybrid/compress.py:6-21; compress_single_task's centring, :24-56)
//   k14_expand         out = scale * (U_high c_high + U_low c_low + mean)
//                      (reconstruct_from_coefficients, src/svd_hybrid/merge.py:144-194)
//   k14_mask_count / k14_chunk_scan / k14_select / k14_scatter
//                      tensor.flatten()[mask] and its inverse in ascending element order
//                      (apply_mask_to_tensor / get_unmasked_portion / reconstruct_from_masked,
//                      src/svd_hybrid/mask_loader.py:665-763)
//
// All of them are single streaming passes over U (rows x cols, row-major with a leading dimension) or over the
// tensor and its byte mask; the projections are reduced in a fixed order (per-thread fp32 partial sums, fp64 above),
// so results do not depend on the launch.
#include "svdq_kernels.h"

namespace svdq {

namespace {

constexpr int kOpBlock = 256;
constexpr int kOpCols = 32;           // columns per projection launch
constexpr int kOpGrid = 148 * 4;      // projection partial rows
constexpr int kSelChunk = 4096;       // elements per selection chunk (one CTA)
constexpr int kSelPer = kSelChunk / kOpBlock;

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__half v) { return __half2float(v); }

template <typename UT, int R>
__global__ void __launch_bounds__(kOpBlock) k14_project(const UT* __restrict__ U, int64_t ld, int cols, int64_t rows,
                                                        const float* __restrict__ delta,
                                                        const float* __restrict__ mean, double* __restrict__ partial) {
    float acc[R];
#pragma unroll
    for (int j = 0; j < R; ++j) acc[j] = 0.f;
    const int64_t stride = (int64_t)gridDim.x * kOpBlock;
    for (int64_t i = (int64_t)blockIdx.x * kOpBlock + threadIdx.x; i < rows; i += stride) {
        float x = __ldg(delta + i);
        if (mean) x -= __ldg(mean + i);
        const UT* row = U + i * ld;
#pragma unroll
        for (int j = 0; j < R; ++j)
            if (j < cols) acc[j] = fmaf(to_f32(row[j]), x, acc[j]);
    }
    __shared__ double s_red[kOpBlock / 32][R];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < R; ++j) {
        double v = (double)acc[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) s_red[warp][j] = v;
    }
    __syncthreads();
    if (threadIdx.x < R) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < kOpBlock / 32; ++w) v += s_red[w][threadIdx.x];
        partial[(int64_t)blockIdx.x * kOpCols + threadIdx.x] = v;
    }
}

__global__ void __launch_bounds__(32) k14_project_finish(const double* __restrict__ partial, int n_blocks, int cols,
                                                         float* __restrict__ c) {
    const int j = threadIdx.x;
    if (j >= cols) return;
    double v = 0.0;
    for (int b = 0; b < n_blocks; ++b) v += partial[(int64_t)b * kOpCols + j];
    c[j] = (float)v;
}

template <typename UT>
__global__ void __launch_bounds__(kOpBlock) k14_expand(const UT* __restrict__ Uh, int64_t ldh, int k,
                                                       const UT* __restrict__ Ul, int64_t ldl, int nl, int64_t rows,
                                                       const float* __restrict__ ch, const float* __restrict__ cl,
                                                       const float* __restrict__ mean, float scale,
                                                       float* __restrict__ out) {
    extern __shared__ float s_c[];
    for (int j = threadIdx.x; j < k + nl; j += kOpBlock) s_c[j] = j < k ? ch[j] : cl[j - k];
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * kOpBlock;
    for (int64_t i = (int64_t)blockIdx.x * kOpBlock + threadIdx.x; i < rows; i += stride) {
        // the reference forms the two matvecs separately, adds them, then the mean (merge.py:182-192)
        float hi = 0.f, lo = 0.f;
        const UT* rh = Uh + i * ldh;
        for (int j = 0; j < k; ++j) hi = fmaf(to_f32(rh[j]), s_c[j], hi);
        if (nl > 0) {
            const UT* rl = Ul + i * ldl;
            for (int j = 0; j < nl; ++j) lo = fmaf(to_f32(rl[j]), s_c[k + j], lo);
        }
        float v = hi + lo;
        if (mean) v += __ldg(mean + i);
        out[i] = v * scale;
    }
}

// ---- mask selection -------------------------------------------------------------------------------------
// kept(i) = (mask[i] != 0) != invert.  A thread owns kSelPer consecutive elements of its CTA's chunk.
__device__ __forceinline__ uint32_t sel_bits(const uint8_t* __restrict__ mask, int64_t lo, int64_t n, int invert) {
    uint32_t bits = 0;
    if (lo + kSelPer <= n && ((uintptr_t)(mask + lo) & 15u) == 0) {
        const uint4 w = __ldg(reinterpret_cast<const uint4*>(mask + lo));
        const uint32_t q[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int e = 0; e < 16; ++e) bits |= (((q[e >> 2] >> (8 * (e & 3))) & 0xffu) != 0u ? 1u : 0u) << e;
    } else {
#pragma unroll
        for (int e = 0; e < kSelPer; ++e)
            if (lo + e < n && mask[lo + e] != 0) bits |= 1u << e;
    }
    if (invert) {
        const int64_t left = n - lo;
        const uint32_t valid = left >= kSelPer ? 0xffffu : (left > 0 ? (1u << (int)left) - 1u : 0u);
        bits = ~bits & valid;
    }
    return bits;
}
static_assert(kSelPer == 16, "sel_bits reads one 16-byte vector per thread");

// block-wide exclusive prefix of one count per thread; returns the block total through s_tot
__device__ __forceinline__ uint32_t block_exclusive(uint32_t cnt, uint32_t* s_warp, uint32_t& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < kOpBlock / 32; ++w) {
        const uint32_t v = s_warp[w];
        if (w < warp) base += v;
        tot += v;
    }
    total = tot;
    return base + inc - cnt;
}

__global__ void __launch_bounds__(kOpBlock) k14_mask_count(const uint8_t* __restrict__ mask, int64_t n, int invert,
                                                           int64_t* __restrict__ chunk_off) {
    __shared__ uint32_t s_warp[kOpBlock / 32];
    const int64_t lo = (int64_t)blockIdx.x * kSelChunk + (int64_t)threadIdx.x * kSelPer;
    uint32_t total;
    (void)block_exclusive(__popc(sel_bits(mask, lo, n, invert)), s_warp, total);
    if (threadIdx.x == 0) chunk_off[blockIdx.x] = total;          // counts; k14_chunk_scan turns them into offsets
}

// in-place exclusive scan of the chunk counts by one CTA; chunk_off[n_chunks] = number of kept elements
__global__ void __launch_bounds__(1024) k14_chunk_scan(int64_t* __restrict__ chunk_off, int64_t n_chunks) {
    __shared__ int64_t s_warp[32];
    __shared__ int64_t s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int64_t b = 0; b < n_chunks; b += 1024) {
        const int64_t i = b + threadIdx.x;
        const int64_t cnt = i < n_chunks ? chunk_off[i] : 0;
        int64_t inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int64_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        int64_t base = s_carry, tot = 0;
        for (int w = 0; w < 32; ++w) {
            const int64_t v = s_warp[w];
            if (w < warp) base += v;
            tot += v;
        }
        if (i < n_chunks) chunk_off[i] = base + inc - cnt;
        __syncthreads();
        if (threadIdx.x == 0) s_carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) chunk_off[n_chunks] = s_carry;
}

template <typename E, bool SCATTER>
__global__ void __launch_bounds__(kOpBlock) k14_select(const E* __restrict__ src, const uint8_t* __restrict__ mask,
                                                       int64_t n, int invert, const int64_t* __restrict__ chunk_off,
                                                       E* __restrict__ dst) {
    __shared__ uint32_t s_warp[kOpBlock / 32];
    const int64_t lo = (int64_t)blockIdx.x * kSelChunk + (int64_t)threadIdx.x * kSelPer;
    uint32_t bits = sel_bits(mask, lo, n, invert);
    uint32_t total;
    int64_t o = chunk_off[blockIdx.x] + block_exclusive(__popc(bits), s_warp, total);
    while (bits) {
        const int e = __ffs(bits) - 1;
        bits &= bits - 1;
        if (SCATTER) dst[lo + e] = src[o++];          // compact values -> their element positions
        else dst[o++] = src[lo + e];                  // kept elements -> compact, ascending element order
    }
}

template <typename E>
cudaError_t select_go(bool scatter, const void* src, const uint8_t* mask, int64_t n, int invert,
                      const int64_t* chunk_off, void* dst, cudaStream_t st) {
    const int64_t chunks = (n + kSelChunk - 1) / kSelChunk;
    if (chunks > 0x7fffffff) return cudaErrorInvalidValue;
    if (scatter) k14_select<E, true><<<(int)chunks, kOpBlock, 0, st>>>((const E*)src, mask, n, invert, chunk_off, (E*)dst);
    else k14_select<E, false><<<(int)chunks, kOpBlock, 0, st>>>((const E*)src, mask, n, invert, chunk_off, (E*)dst);
    return cudaGetLastError();
}

template <typename UT>
cudaError_t project_go(const void* U, int64_t ld, int cols, int64_t rows, const float* delta, const float* mean,
                       float* c, double* scratch, cudaStream_t st) {
    int64_t g = (rows + kOpBlock - 1) / kOpBlock;
    if (g > kOpGrid) g = kOpGrid;
    if (g < 1) g = 1;
    if (cols <= 8) k14_project<UT, 8><<<(int)g, kOpBlock, 0, st>>>((const UT*)U, ld, cols, rows, delta, mean, scratch);
    else if (cols <= 16) k14_project<UT, 16><<<(int)g, kOpBlock, 0, st>>>((const UT*)U, ld, cols, rows, delta, mean, scratch);
    else k14_project<UT, 32><<<(int)g, kOpBlock, 0, st>>>((const UT*)U, ld, cols, rows, delta, mean, scratch);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    k14_project_finish<<<1, 32, 0, st>>>(scratch, (int)g, cols, c);
    return cudaGetLastError();
}

}  // namespace

size_t k14_project_scratch_bytes() { return (size_t)kOpGrid * kOpCols * sizeof(double); }
int k14_select_chunk() { return kSelChunk; }

cudaError_t k14_project_launch(bool fp16, const void* U, int64_t ld, int cols, int64_t rows, const float* delta,
                               const float* mean, float* c, void* scratch, cudaStream_t st) {
    if (cols < 1 || cols > kOpCols || rows < 0 || ld < cols) return cudaErrorInvalidValue;
    return fp16 ? project_go<__half>(U, ld, cols, rows, delta, mean, c, (double*)scratch, st)
                : project_go<float>(U, ld, cols, rows, delta, mean, c, (double*)scratch, st);
}

cudaError_t k14_expand_launch(bool fp16, const void* Uh, int64_t ldh, int k, const void* Ul, int64_t ldl, int nl,
                              int64_t rows, const float* ch, const float* cl, const float* mean, float scale,
                              float* out, cudaStream_t st) {
    if (k < 0 || nl < 0 || k + nl > 8192 || rows < 0 || ldh < k || (nl > 0 && ldl < nl)) return cudaErrorInvalidValue;
    if (rows == 0) return cudaSuccess;
    int64_t g = (rows + kOpBlock - 1) / kOpBlock;
    if (g > 148 * 8) g = 148 * 8;
    const size_t sm = (size_t)(k + nl + 1) * sizeof(float);
    if (fp16) k14_expand<__half><<<(int)g, kOpBlock, sm, st>>>((const __half*)Uh, ldh, k, (const __half*)Ul, ldl, nl, rows,
                                                               ch, cl, mean, scale, out);
    else k14_expand<float><<<(int)g, kOpBlock, sm, st>>>((const float*)Uh, ldh, k, (const float*)Ul, ldl, nl, rows, ch, cl,
                                                         mean, scale, out);
    return cudaGetLastError();
}

cudaError_t k14_mask_offsets_launch(const uint8_t* mask, int64_t n, int invert, int64_t* chunk_off, cudaStream_t st) {
    const int64_t chunks = (n + kSelChunk - 1) / kSelChunk;
    if (chunks > 0x7fffffff) return cudaErrorInvalidValue;
    if (chunks > 0) {
        k14_mask_count<<<(int)chunks, kOpBlock, 0, st>>>(mask, n, invert, chunk_off);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    k14_chunk_scan<<<1, 1024, 0, st>>>(chunk_off, chunks);
    return cudaGetLastError();
}

cudaError_t k14_select_launch(bool scatter, int elem_bytes, const void* src, const uint8_t* mask, int64_t n, int invert,
                              const int64_t* chunk_off, void* dst, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    switch (elem_bytes) {
        case 1: return select_go<uint8_t>(scatter, src, mask, n, invert, chunk_off, dst, st);
        case 2: return select_go<uint16_t>(scatter, src, mask, n, invert, chunk_off, dst, st);
        case 4: return select_go<uint32_t>(scatter, src, mask, n, invert, chunk_off, dst, st);
        case 8: return select_go<uint64_t>(scatter, src, mask, n, invert, chunk_off, dst, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace svdq
