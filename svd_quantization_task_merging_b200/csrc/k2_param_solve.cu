// K2 — the small batched per-parameter solve between the two streaming passes.
//   k2_gram_reduce : sums the per-tile partial Grams / counts of K1 in a fixed order (fp64)
//   k2_param_solve : one warp per parameter runs svdq::solve_param (k2_core.h): centring,
//                    Jacobi eigensolve, energy rank selection, closed-form coefficients,
//                    fp16 high block, multi-stage RTVQ low block, weighted average, W matrix.
// Replaces the rest of torch.linalg.svd (src/svd_hybrid/basis.py:241), select_rank
// (basis.py:199-211), project_to_basis (compress.py:18-19), the fp16 cast (compress.py:44-45),
// RTVQ (rtvq.py:55-79) and dequantize_and_average (merge.py:127-139).  Latency-bound, tiny.
#include "svdq_kernels.h"

namespace svdq {


// Parameters with many tiles (Llama-sized: lm_head has 32,064 tiles of 16,384 elements) are reduced in two levels so
// that more than one SM works on them: k2_gram_partial sums kRedSplit contiguous tile ranges per parameter and
// leaves each range's fp64 partial in place of the range's first two tiles; k2_gram_reduce adds the partials in
// range order.  The split depends only on the parameter's own tile count: the summation order -- and so every bit
// of the result -- is independent of how parameters are sharded over ranks.
constexpr int kRedSplit = 64;
constexpr int kRedSplitMin = 512;        // tiles; below that one CTA reduces the parameter directly
constexpr int kRedMaxBig = 1024;         // parameters scanned per blockIdx.y (bounds the shared list of big ones)

__device__ __forceinline__ void red_split(int64_t T, int64_t& L, int& ns) {
    L = (T + kRedSplit - 1) / kRedSplit;          // >= 8 tiles per range; the last range takes the remainder
    ns = (int)(T / L);
}

__global__ void __launch_bounds__(kBlock) k2_gram_partial(const K2ReduceArgs a, float* gram_rw, const int n_params) {
    const int s = blockIdx.x, tid = threadIdx.x;
    const int NT = a.nt, G = tri_count(NT), NACC = a.full ? 2 * G : G;
    __shared__ double s_part[kBlock / 32][tri_count(kMaxTasks)];
    __shared__ int s_big[kRedMaxBig];
    __shared__ int s_nbig;
    const int lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_nbig = 0;
    __syncthreads();
    const int p_lo = blockIdx.y * kRedMaxBig, p_hi = min(n_params, p_lo + kRedMaxBig);
    for (int p = p_lo + tid; p < p_hi; p += kBlock)
        if (a.tile_begin[p + 1] - a.tile_begin[p] >= kRedSplitMin) s_big[atomicAdd(&s_nbig, 1)] = p;
    __syncthreads();
    const int nbig = s_nbig;
    for (int b = 0; b < nbig; ++b) {               // order of the list is irrelevant: parameters are independent
        const int p = s_big[b];
        const int64_t t0 = a.tile_begin[p], t1 = a.tile_begin[p + 1];
        int64_t L;
        int ns;
        red_split(t1 - t0, L, ns);
        if (s >= ns) continue;                     // uniform per CTA
        const int64_t b0 = t0 + (int64_t)s * L, b1 = (s == ns - 1) ? t1 : b0 + L;
        for (int r = lane; r < NACC; r += 32) {
            double acc = 0.0;
            for (int64_t t = b0 + warp; t < b1; t += kBlock / 32) acc += (double)a.gram[t * NACC + r];
            s_part[warp][r] = acc;
        }
        __syncthreads();                           // every read of the range is done before its head is overwritten
        for (int r = tid; r < NACC; r += kBlock) {
            double m = 0.0;
            for (int w = 0; w < kBlock / 32; ++w) m += s_part[w][r];
            gram_rw[b0 * NACC + 2 * r] = __int_as_float(__double2loint(m));
            gram_rw[b0 * NACC + 2 * r + 1] = __int_as_float(__double2hiint(m));
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kBlock) k2_gram_reduce(const K2ReduceArgs a, const int n_big) {
    const int p = blockIdx.x, tid = threadIdx.x;
    const int NT = a.nt, G = tri_count(NT), NACC = a.full ? 2 * G : G;
    const int64_t t0 = a.tile_begin[p], t1 = a.tile_begin[p + 1];
    __shared__ double s_part[kBlock / 32][tri_count(kMaxTasks)];   // 528 >= 2 * tri_count(16): FULL needs nt <= 16
    __shared__ unsigned long long s_cnt[kBlock / 32];
    const int lane = tid & 31, warp = tid >> 5;
    const bool two_level = n_big >= 0 && t1 - t0 >= kRedSplitMin;

    if (two_level) {
        // partials of k2_gram_partial, added in range order; rows spread over the threads, warp 0's slot holds them
        int64_t L;
        int ns;
        red_split(t1 - t0, L, ns);
        for (int r = tid; r < NACC; r += kBlock) {
            double acc = 0.0;
            for (int s = 0; s < ns; ++s) {
                const float* q = a.gram + (t0 + (int64_t)s * L) * NACC + 2 * r;
                acc += __hiloint2double(__float_as_int(q[1]), __float_as_int(q[0]));
            }
            s_part[0][r] = acc;
        }
        for (int r = tid; r < NACC; r += kBlock)
            for (int w = 1; w < kBlock / 32; ++w) s_part[w][r] = 0.0;
    } else {
        // each warp owns tiles warp, warp+8, ...; inside a warp lane l owns accumulator rows l, l+32, ...
        for (int r = lane; r < NACC; r += 32) {
            double s = 0.0;
            for (int64_t t = t0 + warp; t < t1; t += kBlock / 32) s += (double)a.gram[t * NACC + r];
            s_part[warp][r] = s;
        }
    }
    unsigned long long c = 0;
    for (int64_t t = t0 + tid; t < t1; t += kBlock) c += a.count[t];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) s_cnt[warp] = c;
    __syncthreads();
    for (int r = tid; r < G; r += kBlock) {
        double m = 0.0, u = 0.0;
        for (int w = 0; w < kBlock / 32; ++w) {
            m += s_part[w][r];
            if (a.full) u += s_part[w][G + r];
        }
        // unpack the upper triangle
        int i = 0, rem = r;
        while (rem >= NT - i) { rem -= NT - i; ++i; }
        const int j = i + rem;
        a.gram_masked[(int64_t)p * NT * NT + i * NT + j] = m;
        a.gram_masked[(int64_t)p * NT * NT + j * NT + i] = m;
        if (a.full && a.gram_all) {
            const double all = a.full == 2 ? m + u : u;              // second block: all elements / complement
            a.gram_all[(int64_t)p * NT * NT + i * NT + j] = all;
            a.gram_all[(int64_t)p * NT * NT + j * NT + i] = all;
        }
        if (a.full == 2 && a.gram_noise) {
            a.gram_noise[(int64_t)p * NT * NT + i * NT + j] = u;
            a.gram_noise[(int64_t)p * NT * NT + j * NT + i] = u;
        }
    }
    if (tid == 0) {
        unsigned long long tot = 0;
        for (int w = 0; w < kBlock / 32; ++w) tot += s_cnt[w];
        a.dm[p] = (int64_t)tot;
        if (a.full == 2 && a.dm_noise) {
            const bool on = a.has_mask[p] != 0 && (int64_t)tot >= (int64_t)a.min_mask_size && tot > 0;
            a.dm_noise[p] = on ? a.numel[p] - (int64_t)tot : 0;
        }
    }
}


struct WarpLanes {
    int lane;
    static constexpr int nl = 32;
    __device__ __forceinline__ void sync() { __syncwarp(); }
};
// a whole CTA as the lane set: above 8 tasks a Jacobi round has more (pair, row) items than a warp has lanes, and the
// solve of the 17..32-task merges is pure latency (one CTA per parameter, 296 parameters); every sync() of the
// lane-SPMD code is reached uniformly, and no value depends on the number of lanes, so the results are the same bits
template <int THREADS> struct BlockLanes {
    int lane;
    static constexpr int nl = THREADS;
    __device__ __forceinline__ void sync() { __syncthreads(); }
};

template <class LN, int THREADS>
__global__ void __launch_bounds__(THREADS) k2_param_solve(const K2SolveArgs a) {
    __shared__ SolveScratch sc;
    const int p = blockIdx.x;
    const int NT = a.cfg.n_tasks, S = a.cfg.stages;
    const int64_t nn = (int64_t)NT * NT;
    SolveIn in;
    in.G = a.gram_masked + p * nn;
    in.dm = a.dm[p];
    in.has_mask = a.has_mask[p];
    in.present = a.present[p];
    in.weights = a.weights;
    in.avg_order = a.avg_order;
    in.sign_ref = a.sign_ref ? a.sign_ref + p * nn : nullptr;
    SolveOut out;
    out.info = a.info + (int64_t)p * 8;
    out.sv = a.sv + (int64_t)p * NT;
    out.scal = a.scal + (int64_t)p * 4;
    out.coef = a.coef + p * nn;
    out.chigh = a.chigh + p * nn;
    out.codes = a.codes + p * nn * S;
    out.qscale = a.qscale + (int64_t)p * NT * S;
    out.qzp = a.qzp + (int64_t)p * NT * S;
    out.qres = a.qres + (int64_t)p * NT * S;
    out.chat = a.chat + p * nn;
    out.cbar = a.cbar + (int64_t)p * NT;
    out.W = a.W + p * nn;
    out.gvec = a.gvec + (int64_t)p * NT;
    out.V = a.V + p * nn;
    LN ln{(int)threadIdx.x};
    solve_param(a.cfg, in, out, sc, ln);
}

// weighted average only (reads info / chat / W written by k2_param_solve)
__global__ void __launch_bounds__(32) k2_param_average(const K2SolveArgs a) {
    const int p = blockIdx.x;
    const int NT = a.cfg.n_tasks;
    const int64_t nn = (int64_t)NT * NT;
    SolveIn in;
    in.G = nullptr; in.dm = 0; in.has_mask = 0;
    in.present = a.present[p];
    in.weights = a.weights;
    in.avg_order = a.avg_order;
    in.cluster_of = a.cluster_of;
    in.omega = a.omega;
    in.sign_ref = nullptr;
    SolveOut out = {};
    out.info = a.info + (int64_t)p * 8;
    out.scal = a.scal + (int64_t)p * 4;
    out.chat = a.chat + p * nn;
    out.cbar = a.cbar + (int64_t)p * NT;
    out.W = a.W + p * nn;
    out.gvec = a.gvec + (int64_t)p * NT;
    WarpLanes ln{(int)threadIdx.x};
    average_param(a.cfg, in, out, ln);
}

// re-quantisation after the exact projection (K7): sum the tile partials in a fixed order (fp64), overwrite the
// coefficients of the selected parameters and redo their fp16 / RTVQ step
__global__ void __launch_bounds__(32) k2_param_requantize(const K2RequantArgs a) {
    const int p = blockIdx.x, lane = threadIdx.x;
    const int NT = a.cfg.n_tasks, S = a.cfg.stages;
    const int64_t nn = (int64_t)NT * NT;
    const int64_t t0 = a.sel_tile_begin[p], t1 = a.sel_tile_begin[p + 1];
    if (t1 <= t0 || a.info[(int64_t)p * 8] != kSolved) return;
    // only the columns with a basis direction are re-projected: a numerically-null direction keeps its
    // closed-form coefficient (round-off dust, like LAPACK's; an exact zero there would turn the
    // reference's 2-element RTVQ edge into NaN, rtvq.py:17)
    const int r = a.info[(int64_t)p * 8 + 4];
    float* coef = a.coef + p * nn;
    for (int i = lane; i < NT * NT; i += 32) {
        const int j = i % NT;
        if (j >= r) continue;
        double s = 0.0;
        for (int64_t tl = t0; tl < t1; ++tl) s += (double)a.proj[tl * nn + i];
        coef[i] = (float)s;
    }
    __syncwarp();
    SolveOut out = {};
    out.info = const_cast<int32_t*>(a.info) + (int64_t)p * 8;
    out.coef = coef;
    out.chigh = a.chigh + p * nn;
    out.codes = a.codes + p * nn * S;
    out.qscale = a.qscale + (int64_t)p * NT * S;
    out.qzp = a.qzp + (int64_t)p * NT * S;
    out.qres = a.qres + (int64_t)p * NT * S;
    out.chat = a.chat + p * nn;
    WarpLanes ln{lane};
    quantize_param(a.cfg, a.present[p], out, ln);
}

cudaError_t k2_requant_launch(const K2RequantArgs& a, int n_params, cudaStream_t st) {
    if (n_params <= 0) return cudaSuccess;
    if (a.cfg.n_tasks < 1 || a.cfg.n_tasks > kCoreMaxTasks) return cudaErrorInvalidValue;
    k2_param_requantize<<<n_params, 32, 0, st>>>(a);
    return cudaGetLastError();
}

cudaError_t k2_average_launch(const K2SolveArgs& a, int n_params, cudaStream_t st) {
    if (n_params <= 0) return cudaSuccess;
    if (a.cfg.n_tasks < 1 || a.cfg.n_tasks > kCoreMaxTasks) return cudaErrorInvalidValue;
    k2_param_average<<<n_params, 32, 0, st>>>(a);
    return cudaGetLastError();
}

cudaError_t k2_reduce_launch(const K2ReduceArgs& a, int n_params, cudaStream_t st) {
    if (n_params <= 0) return cudaSuccess;
    if (a.nt < 1 || a.nt > kMaxTasks || (a.full && a.nt > 16)) return cudaErrorInvalidValue;
    const dim3 grid(kRedSplit, (n_params + kRedMaxBig - 1) / kRedMaxBig);
    k2_gram_partial<<<grid, kBlock, 0, st>>>(a, const_cast<float*>(a.gram), n_params);
    k2_gram_reduce<<<n_params, kBlock, 0, st>>>(a, 0);
    return cudaGetLastError();
}

cudaError_t k2_solve_launch(const K2SolveArgs& a, int n_params, cudaStream_t st) {
    if (n_params <= 0) return cudaSuccess;
    if (a.cfg.n_tasks < 1 || a.cfg.n_tasks > kCoreMaxTasks) return cudaErrorInvalidValue;
    if (a.cfg.stages < 1 || a.cfg.stages > kCoreMaxStages || a.cfg.bits < 1 || a.cfg.bits > 8)
        return cudaErrorInvalidValue;
    if (a.cfg.n_tasks <= 8) k2_param_solve<WarpLanes, 32><<<n_params, 32, 0, st>>>(a);
    else if (a.cfg.n_tasks <= 16) k2_param_solve<BlockLanes<128>, 128><<<n_params, 128, 0, st>>>(a);     // 8 pairs x 16 rows
    else k2_param_solve<BlockLanes<256>, 256><<<n_params, 256, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace svdq
