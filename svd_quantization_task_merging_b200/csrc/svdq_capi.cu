// C ABI of libsvdq.so (see include/svdq.h).  Thin argument checking + kernel launches; no
// allocation, no device synchronisation, no hidden state besides the thread-local error text.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <exception>
#include <thread>
#include <vector>

#include "../../include/svdq.h"
#include "svdq_kernels.h"

namespace {

thread_local char g_err[512] = "";

int fail_arg(const char* fn, const char* what) {
    snprintf(g_err, sizeof(g_err), "%s: invalid argument: %s", fn, what);
    return -1;
}

int finish(const char* fn, cudaError_t e) {
    if (e == cudaSuccess) return 0;
    if (e == cudaErrorInvalidValue) {
        snprintf(g_err, sizeof(g_err), "%s: invalid argument (unsupported n_tasks / dtype / bits / stages)", fn);
        (void)cudaGetLastError();
        return -1;
    }
    snprintf(g_err, sizeof(g_err), "%s: CUDA error %d (%s)", fn, (int)e, cudaGetErrorString(e));
    return (int)e;
}

#define REQUIRE(cond, what) do { if (!(cond)) return fail_arg(__func__, what); } while (0)

// SVDQ_STAGED is a bit mask selecting the staged persistent (TMA ring) variant per pass: bit 0 = K1,
// bit 1 = K3 (A/B switch; 0 = direct-load kernels).  Default 3: measured on B200 (ViT-L-14 x 8) the ring
// wins for both passes (pass 1: 2.24 vs 3.19 ms, pass 2: 1.95 vs 2.16 ms).
int staged_mask() {
    static const int m = [] { const char* v = getenv("SVDQ_STAGED"); return v ? atoi(v) : 3; }();
    return m;
}
int sm_count() {
    static const int n = [] {
        int dev = 0, v = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) return 148;
        return v;
    }();
    return n;
}
bool aligned16(const void* p) { return ((uintptr_t)p & 15u) == 0; }

// SVDQ_TC (bit mask, default 3): bit 0 = tensor-core pass 1 for 16-bit inputs (tcgen05, k9_gram_tc.cu; up to 8 tasks,
// single Gram block), bit 1 = tensor-core pass 2 for bf16 inputs (k10_merge_tc.cu; up to 8 tasks, no diagnostics /
// noise region), bit 2 = tensor-core single-pass Gram of the wide path for fp32 inputs (k12_gram_wide_tc.cu; 3-piece
// bf16 split), bit 3 = tensor-core pass 2 of the wide path for fp32 inputs and fp16 bases (k13_merge_wide_tc.cu);
// 0 = the CUDA-core kernels (A/B switch).  SVDQ_TC_CHAIN = MMA steps per short accumulator chain of K12.
int tc_enabled() {                      // read per call: tests flip it inside one process
    const char* v = getenv("SVDQ_TC");
    return v ? atoi(v) : 15;
}
// SVDQ_COMPACT_DIAG: pass 2 with fused diagnostics (up to 8 tasks) compacts the elements inside the combined mask before
// the arithmetic (k3c_merge_diag_compact.cu).  1 (default) = per tile inside that kernel: compaction where the mask
// keeps fewer than 55 % of the elements, the plain walk elsewhere; 2 = compaction for every tile;
// 0 = non-compacting kernel for every tile (A/B switches)
int compact_diag() {
    const char* v = getenv("SVDQ_COMPACT_DIAG");
    return v ? atoi(v) : 1;
}
int tc_chain() {
    const char* v = getenv("SVDQ_TC_CHAIN");
    return v ? atoi(v) : 4;
}

cudaError_t k1_launch(int dtype, int nt, const svdq::K1Args& a, int n_tiles, bool full, cudaStream_t st) {
    if ((tc_enabled() & 1) && !full && nt <= 8 && dtype != svdq::kF32) {
        const cudaError_t e = dtype == svdq::kBF16 ? svdq::k9_launch_dtype<svdq::kBF16>(nt, a, n_tiles, sm_count(), st)
                                                   : svdq::k9_launch_dtype<svdq::kF16>(nt, a, n_tiles, sm_count(), st);
        if (e != cudaErrorNotSupported) return e;
    }
    if ((staged_mask() & 1) && nt <= 8) {
        cudaError_t e = cudaErrorNotSupported;
        switch (dtype) {
            case svdq::kF32:  e = svdq::k1s_launch_dtype<svdq::kF32>(nt, a, n_tiles, full, sm_count(), st); break;
            case svdq::kBF16: e = svdq::k1s_launch_dtype<svdq::kBF16>(nt, a, n_tiles, full, sm_count(), st); break;
            case svdq::kF16:  e = svdq::k1s_launch_dtype<svdq::kF16>(nt, a, n_tiles, full, sm_count(), st); break;
            default: break;
        }
        if (e != cudaErrorNotSupported) return e;
    }
    switch (dtype) {
        case svdq::kF32:  return svdq::k1_launch_dtype<svdq::kF32>(nt, a, n_tiles, full, st);
        case svdq::kBF16: return svdq::k1_launch_dtype<svdq::kBF16>(nt, a, n_tiles, full, st);
        case svdq::kF16:  return svdq::k1_launch_dtype<svdq::kF16>(nt, a, n_tiles, full, st);
        default:          return cudaErrorInvalidValue;
    }
}
cudaError_t k3_launch(int dtype, int nt, const svdq::K3Args& a, int n_tiles, bool fp16b, bool diag, cudaStream_t st) {
    if ((tc_enabled() & 8) && dtype == svdq::kF32 && nt > 16 && fp16b && !diag && a.info_n == nullptr) {
        const cudaError_t e = svdq::k13_launch_dtype<svdq::kF32>(nt, a, n_tiles, fp16b, sm_count(), st);
        if (e != cudaErrorNotSupported) return e;
    }
    if (nt > SVDQ_MAX_STREAM_TASKS) {           // 17..32 tasks: runtime-N kernel
        switch (dtype) {
            case svdq::kF32:  return svdq::k6_merge_launch_dtype<svdq::kF32>(nt, a, n_tiles, fp16b, diag, st);
            case svdq::kBF16: return svdq::k6_merge_launch_dtype<svdq::kBF16>(nt, a, n_tiles, fp16b, diag, st);
            case svdq::kF16:  return svdq::k6_merge_launch_dtype<svdq::kF16>(nt, a, n_tiles, fp16b, diag, st);
            default:          return cudaErrorInvalidValue;
        }
    }
    if ((tc_enabled() & 2) && dtype == svdq::kBF16 && nt <= 8 && !diag && a.info_n == nullptr) {
        const cudaError_t e = svdq::k10_launch_dtype<svdq::kBF16>(nt, a, n_tiles, fp16b, sm_count(), st);
        if (e != cudaErrorNotSupported) return e;
    }
    if (diag && nt <= 8 && a.info_n == nullptr && compact_diag() != 0) {
        const bool split = compact_diag() == 1 && a.packed != nullptr;      // without packed masks nothing is sparse
        if (split || compact_diag() == 2) {
            svdq::K3Args c = a;
            c.diag_select = split ? 2 : 0;
            cudaError_t e = cudaErrorNotSupported;
            switch (dtype) {
                case svdq::kF32:  e = svdq::k3c_launch_dtype<svdq::kF32>(nt, c, n_tiles, fp16b, st); break;
                case svdq::kBF16: e = svdq::k3c_launch_dtype<svdq::kBF16>(nt, c, n_tiles, fp16b, st); break;
                case svdq::kF16:  e = svdq::k3c_launch_dtype<svdq::kF16>(nt, c, n_tiles, fp16b, st); break;
                default: break;
            }
            if (e != cudaErrorNotSupported) return e;
        }
    }
    if ((staged_mask() & 2) && nt <= 8 && !diag && a.info_n == nullptr) {
        cudaError_t e = cudaErrorNotSupported;
        switch (dtype) {
            case svdq::kF32:  e = svdq::k3s_launch_dtype<svdq::kF32>(nt, a, n_tiles, fp16b, sm_count(), st); break;
            case svdq::kBF16: e = svdq::k3s_launch_dtype<svdq::kBF16>(nt, a, n_tiles, fp16b, sm_count(), st); break;
            case svdq::kF16:  e = svdq::k3s_launch_dtype<svdq::kF16>(nt, a, n_tiles, fp16b, sm_count(), st); break;
            default: break;
        }
        if (e != cudaErrorNotSupported) return e;
    }
    switch (dtype) {
        case svdq::kF32:  return svdq::k3_launch_dtype<svdq::kF32>(nt, a, n_tiles, fp16b, diag, st);
        case svdq::kBF16: return svdq::k3_launch_dtype<svdq::kBF16>(nt, a, n_tiles, fp16b, diag, st);
        case svdq::kF16:  return svdq::k3_launch_dtype<svdq::kF16>(nt, a, n_tiles, fp16b, diag, st);
        default:          return cudaErrorInvalidValue;
    }
}
cudaError_t k5_launch(int dtype, int nt, const svdq::K5Args& a, int n_tiles, cudaStream_t st) {
    switch (dtype) {
        case svdq::kF32:  return svdq::k5_launch_dtype<svdq::kF32>(nt, a, n_tiles, st);
        case svdq::kBF16: return svdq::k5_launch_dtype<svdq::kBF16>(nt, a, n_tiles, st);
        case svdq::kF16:  return svdq::k5_launch_dtype<svdq::kF16>(nt, a, n_tiles, st);
        default:          return cudaErrorInvalidValue;
    }
}

}  // namespace

extern "C" {

int svdq_abi_version(void) { return SVDQ_ABI_VERSION; }

const char* svdq_last_error(void) { return g_err; }

int64_t svdq_k4_scratch_bytes(void) { return (int64_t)sizeof(svdq::K4Stats) * svdq::kK4MaxGrid; }

static int tv_mask_gram_impl(const char* fn_name, int mask_bits, int dtype, int n_tasks, int mask_strategy, int full,
                             int64_t n_tiles, int tile_elems, const void* const* tensors, const uint8_t* const* masks,
                             const int64_t* numel, const int32_t* tile_param, const int32_t* tile_local,
                             const int64_t* pmask_off, uint32_t* packed, float* gram, uint32_t* count, void* stream) {
    REQUIRE(n_tasks >= 1 && n_tasks <= SVDQ_MAX_STREAM_TASKS, "n_tasks must be in [1, 16]");
    REQUIRE(dtype >= 0 && dtype <= 2, "dtype");
    REQUIRE(mask_strategy >= 0 && mask_strategy <= 2, "Unknown mask strategy");
    REQUIRE(tile_elems > 0 && tile_elems % svdq::kStep == 0, "tile_elems must be a positive multiple of 1024");
    REQUIRE(n_tiles >= 0 && n_tiles < (1ll << 31), "n_tiles");
    if (n_tiles == 0) return 0;
    REQUIRE(tensors && numel && tile_param && tile_local && gram && count, "null pointer");
    REQUIRE(!masks || (pmask_off && packed), "masks given without packed-mask storage");
    svdq::K1Args a;
    a.tensors = tensors; a.masks = masks; a.numel = numel; a.tile_param = tile_param; a.tile_local = tile_local;
    a.pmask_off = pmask_off; a.packed = packed; a.gram = gram; a.count = count; a.tile_elems = tile_elems;
    a.strategy = mask_strategy;
    a.packed_in = nullptr; a.has_mask_in = nullptr; a.second_complement = full == 2; a.mask_mode = 0;
    a.mask_bits = mask_bits;
    REQUIRE(full >= 0 && full <= 2, "full must be 0 (masked), 1 (+ all elements) or 2 (+ complement)");
    return finish(fn_name, k1_launch(dtype, n_tasks, a, (int)n_tiles, full != 0, (cudaStream_t)stream));
}

int svdq_tv_mask_gram(int dtype, int n_tasks, int mask_strategy, int full, int64_t n_tiles, int tile_elems,
                      const void* const* tensors, const uint8_t* const* masks, const int64_t* numel,
                      const int32_t* tile_param, const int32_t* tile_local, const int64_t* pmask_off,
                      uint32_t* packed, float* gram, uint32_t* count, void* stream) {
    return tv_mask_gram_impl(__func__, 0, dtype, n_tasks, mask_strategy, full, n_tiles, tile_elems, tensors, masks, numel,
                             tile_param, tile_local, pmask_off, packed, gram, count, stream);
}

int svdq_tv_mask_gram_bits(int dtype, int n_tasks, int mask_strategy, int full, int64_t n_tiles, int tile_elems,
                           const void* const* tensors, const uint8_t* const* mask_bits, const int64_t* numel,
                           const int32_t* tile_param, const int32_t* tile_local, const int64_t* pmask_off,
                           uint32_t* packed, float* gram, uint32_t* count, void* stream) {
    return tv_mask_gram_impl(__func__, 1, dtype, n_tasks, mask_strategy, full, n_tiles, tile_elems, tensors, mask_bits,
                             numel, tile_param, tile_local, pmask_off, packed, gram, count, stream);
}

// Host-side transfer encoding of task masks: torch.bool bytes -> bits (element 8k+i = bit i of byte k).  The work of
// a whole batch of masks is cut into equal byte ranges over n_threads host threads (spawned once per call); the
// inner loop is in host_pack.cpp (AVX2 when the CPU has it: 32 mask bytes -> one 32-bit word per instruction pair).
extern "C" void svdq_host_pack_range(const uint8_t* src, int64_t n, uint8_t* dst, int64_t lo, int64_t hi);

int svdq_host_pack_mask_batch(const uint8_t* const* src, const int64_t* n, uint8_t* const* dst, int64_t count, int n_threads) {
    REQUIRE(count >= 0 && (count == 0 || (src && n && dst)), "null pointer");
    REQUIRE(n_threads >= 1 && n_threads <= 256, "n_threads must be in [1, 256]");
    int64_t total = 0;
    for (int64_t i = 0; i < count; ++i) {
        REQUIRE(n[i] >= 0 && (n[i] == 0 || (src[i] && dst[i])), "null mask pointer");
        total += (n[i] + 7) / 8;
    }
    if (total == 0) return 0;
    // output bytes [lo, hi) of the concatenation of all masks' packed images, 64-byte granules
    auto work = [&](int64_t lo, int64_t hi) {
        int64_t off = 0;
        for (int64_t i = 0; i < count && off < hi; ++i) {
            const int64_t n_out = (n[i] + 7) / 8;
            const int64_t a = lo > off ? lo - off : 0, b = (hi - off) < n_out ? (hi - off) : n_out;
            if (a < b) svdq_host_pack_range(src[i], n[i], dst[i], a, b);
            off += n_out;
        }
    };
    if (n_threads == 1 || total < (1 << 16)) { work(0, total); return 0; }
    const int64_t per = ((total + n_threads - 1) / n_threads + 63) & ~(int64_t)63;
    // A thread that cannot be created (EAGAIN under a process / pid limit) must not take the process down: the
    // calling thread packs whatever has not been handed out.
    std::vector<std::thread> pool;
    int64_t handed = 0;
    try {
        pool.reserve(n_threads);
        for (int t = 0; t < n_threads; ++t) {
            const int64_t lo = (int64_t)t * per, hi = lo + per < total ? lo + per : total;
            if (lo >= hi) break;
            pool.emplace_back(work, lo, hi);
            handed = hi;
        }
    } catch (...) {
    }
    if (handed < total) work(handed, total);
    for (auto& th : pool) th.join();
    return 0;
}

int svdq_host_pack_mask(const uint8_t* src, int64_t n, uint8_t* dst, int n_threads) {
    REQUIRE(n >= 0 && (n == 0 || (src && dst)), "null pointer");
    return svdq_host_pack_mask_batch(&src, &n, &dst, 1, n_threads);
}

extern "C" int svdq_host_kmeans_impl(const float*, int, int, int, uint32_t, int, int, double, int32_t*, double*);

int svdq_host_kmeans(const float* features, int n, int d, int k, uint32_t seed, int n_init, int max_iter, double tol,
                     int32_t* labels, double* inertia) {
    REQUIRE(features && labels, "null pointer");
    REQUIRE(n >= 1 && d >= 1 && n <= 4096 && d <= 65536, "feature matrix shape");
    REQUIRE(k >= 1 && k <= n, "Invalid k for the number of samples");      // clustering.py:147-148
    REQUIRE(n_init >= 1 && max_iter >= 1 && tol >= 0.0, "n_init / max_iter / tol");
    try {
        return svdq_host_kmeans_impl(features, n, d, k, seed, n_init, max_iter, tol, labels, inertia);
    } catch (const std::exception& e) {      // allocation failure: report it, do not unwind through the C ABI
        snprintf(g_err, sizeof(g_err), "%s: %s", __func__, e.what());
        return -1;
    }
}

int svdq_mask_pack(int n_tasks, int mask_strategy, int64_t n_tiles, int tile_elems, const uint8_t* const* masks,
                   const int64_t* numel, const int32_t* tile_param, const int32_t* tile_local,
                   const int64_t* pmask_off, uint32_t* packed, uint32_t* count, void* stream) {
    REQUIRE(n_tasks >= 1 && n_tasks <= SVDQ_MAX_TASKS, "n_tasks must be in [1, 32]");
    REQUIRE(mask_strategy >= 0 && mask_strategy <= 2, "Unknown mask strategy");
    REQUIRE(tile_elems > 0 && tile_elems % svdq::kStep == 0, "tile_elems must be a positive multiple of 1024");
    REQUIRE(n_tiles >= 0 && n_tiles < (1ll << 31), "n_tiles");
    if (n_tiles == 0) return 0;
    REQUIRE(numel && tile_param && tile_local && count, "null pointer");
    REQUIRE(!masks || (pmask_off && packed), "masks given without packed-mask storage");
    svdq::K6MaskArgs a;
    a.masks = masks; a.numel = numel; a.tile_param = tile_param; a.tile_local = tile_local; a.pmask_off = pmask_off;
    a.packed = packed; a.count = count; a.tile_elems = tile_elems; a.strategy = mask_strategy; a.n_tasks = n_tasks;
    return finish(__func__, svdq::k6_mask_pack_launch(a, (int)n_tiles, (cudaStream_t)stream));
}

int svdq_gram_staged(int dtype, int n_tasks, int mask_mode, int64_t n_tiles, int tile_elems,
                     const void* const* tensors, const int64_t* numel, const int32_t* tile_param,
                     const int32_t* tile_local, const int64_t* pmask_off, const uint8_t* has_mask,
                     const uint32_t* packed, float* gram, void* stream) {
    REQUIRE(n_tasks >= 1 && n_tasks <= SVDQ_MAX_TASKS, "n_tasks must be in [1, 32]");
    REQUIRE(dtype >= 0 && dtype <= 2, "dtype");
    REQUIRE(mask_mode >= 0 && mask_mode <= 2, "mask_mode must be 0 (masked rows), 1 (all rows) or 2 (unmasked rows)");
    REQUIRE(tile_elems > 0 && tile_elems % 1536 == 0, "tile_elems must be a positive multiple of 1536 (chunks of 512 / 768)");
    REQUIRE(n_tiles >= 0 && n_tiles < (1ll << 31), "n_tiles");
    if (n_tiles == 0) return 0;
    REQUIRE(tensors && numel && tile_param && tile_local && pmask_off && has_mask && packed && gram, "null pointer");
    svdq::K1Args a;
    a.tensors = tensors; a.masks = nullptr; a.numel = numel; a.tile_param = tile_param; a.tile_local = tile_local;
    a.pmask_off = pmask_off; a.packed = nullptr; a.gram = gram; a.count = nullptr; a.tile_elems = tile_elems;
    a.strategy = 0; a.packed_in = packed; a.has_mask_in = has_mask; a.second_complement = 0; a.mask_mode = mask_mode;
    a.mask_bits = 0;
    cudaError_t e;
    if ((tc_enabled() & 4) && dtype == svdq::kF32) {
        e = svdq::k12_launch_dtype<svdq::kF32>(n_tasks, a, (int)n_tiles, sm_count(), tc_chain(), (cudaStream_t)stream);
        if (e != cudaErrorNotSupported) return finish(__func__, e);
    }
    switch (dtype) {
        case svdq::kF32:  e = svdq::k8_launch_dtype<svdq::kF32>(n_tasks, a, (int)n_tiles, sm_count(), (cudaStream_t)stream); break;
        case svdq::kBF16: e = svdq::k8_launch_dtype<svdq::kBF16>(n_tasks, a, (int)n_tiles, sm_count(), (cudaStream_t)stream); break;
        default:          e = svdq::k8_launch_dtype<svdq::kF16>(n_tasks, a, (int)n_tiles, sm_count(), (cudaStream_t)stream); break;
    }
    return finish(__func__, e);
}

int svdq_gram_reduce(int n_tasks, int full, int64_t n_params, int min_mask_size, float* gram,
                     const uint32_t* count, const int64_t* tile_begin, const int64_t* numel, const uint8_t* has_mask,
                     double* gram_masked, double* gram_all, int64_t* dm, double* gram_noise, int64_t* dm_noise,
                     void* stream) {
    REQUIRE(n_tasks >= 1 && n_tasks <= (full ? SVDQ_MAX_STREAM_TASKS : SVDQ_MAX_TASKS),
            "n_tasks must be in [1, 32] ([1, 16] with a second Gram block)");
    REQUIRE(n_params >= 0 && n_params < (1ll << 31), "n_params");
    REQUIRE(full >= 0 && full <= 2, "full must be 0 (masked), 1 (+ all elements) or 2 (+ complement)");
    if (n_params == 0) return 0;
    REQUIRE(gram && count && tile_begin && gram_masked && dm, "null pointer");
    REQUIRE(full == 2 || (!gram_noise && !dm_noise), "noise-region outputs need full == 2");
    REQUIRE(!dm_noise || (numel && has_mask), "dm_noise needs numel and has_mask");
    svdq::K2ReduceArgs a;
    a.gram = gram; a.count = count; a.tile_begin = tile_begin; a.gram_masked = gram_masked; a.gram_all = gram_all;
    a.dm = dm; a.nt = n_tasks; a.full = full;
    a.gram_noise = gram_noise; a.dm_noise = dm_noise; a.numel = numel; a.has_mask = has_mask;
    a.min_mask_size = min_mask_size;
    return finish(__func__, svdq::k2_reduce_launch(a, (int)n_params, (cudaStream_t)stream));
}

int svdq_param_solve(int n_tasks, int64_t n_params, int center, float energy_threshold, int max_rank,
                     int min_mask_size, int rtvq_bits, int rtvq_stages,
                     const double* gram_masked, const int64_t* dm, const uint8_t* has_mask, const uint32_t* present,
                     const double* weights, const int32_t* avg_order, const double* sign_ref,
                     int32_t* info, float* sv, float* scal, float* coef, uint16_t* chigh, uint8_t* codes,
                     float* qscale, float* qzp, float* qres, float* chat, float* cbar, float* W, float* gvec,
                     double* V, void* stream) {
    REQUIRE(n_tasks >= 1 && n_tasks <= SVDQ_MAX_TASKS, "n_tasks must be in [1, 32]");
    REQUIRE(rtvq_bits >= 1 && rtvq_bits <= 8, "Low bits must be in [1, 8]");
    REQUIRE(rtvq_stages >= 1 && rtvq_stages <= SVDQ_MAX_STAGES, "RTVQ stages must be in [1, 8]");
    REQUIRE(energy_threshold > 0.0f && energy_threshold <= 1.0f, "Energy threshold must be in (0, 1]");
    REQUIRE(n_params >= 0 && n_params < (1ll << 31), "n_params");
    if (n_params == 0) return 0;
    REQUIRE(gram_masked && dm && has_mask && present && avg_order, "null input pointer");
    REQUIRE(info && sv && scal && coef && chigh && codes && qscale && qzp && qres && chat && cbar && W && gvec && V,
            "null output pointer");
    svdq::K2SolveArgs a = {};
    a.cfg.n_tasks = n_tasks; a.cfg.center = center; a.cfg.energy_threshold = energy_threshold;
    a.cfg.max_rank = max_rank; a.cfg.min_mask_size = min_mask_size; a.cfg.bits = rtvq_bits; a.cfg.stages = rtvq_stages;
    a.gram_masked = gram_masked; a.dm = dm; a.has_mask = has_mask; a.present = present; a.weights = weights;
    a.avg_order = avg_order; a.sign_ref = sign_ref;
    a.info = info; a.sv = sv; a.scal = scal; a.coef = coef; a.chigh = chigh; a.codes = codes; a.qscale = qscale;
    a.qzp = qzp; a.qres = qres; a.chat = chat; a.cbar = cbar; a.W = W; a.gvec = gvec; a.V = V;
    return finish(__func__, svdq::k2_solve_launch(a, (int)n_params, (cudaStream_t)stream));
}

int svdq_param_average(int n_tasks, int64_t n_params, const uint32_t* present, const double* weights,
                       const int32_t* avg_order, const int32_t* cluster_of, const double* cluster_omega,
                       const int32_t* info, const float* chat, const float* W, float* cbar,
                       float* gvec, float* scal, void* stream) {
    REQUIRE(n_tasks >= 1 && n_tasks <= SVDQ_MAX_TASKS, "n_tasks must be in [1, 32]");
    REQUIRE(n_params >= 0 && n_params < (1ll << 31), "n_params");
    if (n_params == 0) return 0;
    REQUIRE(present && weights && avg_order && info && chat && W && cbar && gvec && scal, "null pointer");
    svdq::K2SolveArgs a = {};
    a.cfg.n_tasks = n_tasks;
    REQUIRE((cluster_of == nullptr) == (cluster_omega == nullptr), "cluster_of and cluster_omega go together");
    a.present = present; a.weights = weights; a.avg_order = avg_order;
    a.cluster_of = cluster_of; a.omega = cluster_omega;
    a.info = const_cast<int32_t*>(info); a.chat = const_cast<float*>(chat); a.W = const_cast<float*>(W);
    a.cbar = cbar; a.gvec = gvec; a.scal = scal;
    return finish(__func__, svdq::k2_average_launch(a, (int)n_params, (cudaStream_t)stream));
}

int svdq_project_exact(int dtype, int n_tasks, int fp16_basis, int center, int region, int64_t n_sel_tiles,
                       int tile_elems,
                       const void* const* tensors, const int64_t* numel, const int32_t* sel_tile_param,
                       const int32_t* sel_tile_local, const int64_t* pmask_off, const uint8_t* has_mask,
                       const uint32_t* packed, const int32_t* info, const float* W, float* proj, void* stream) {
    REQUIRE(n_tasks >= 1 && n_tasks <= SVDQ_MAX_TASKS, "n_tasks must be in [1, 32]");
    REQUIRE(dtype >= 0 && dtype <= 2, "dtype");
    REQUIRE(tile_elems > 0 && tile_elems % svdq::kStep == 0, "tile_elems must be a positive multiple of 1024");
    REQUIRE(n_sel_tiles >= 0 && n_sel_tiles < (1ll << 31), "n_sel_tiles");
    if (n_sel_tiles == 0) return 0;
    REQUIRE(tensors && numel && sel_tile_param && sel_tile_local && has_mask && info && W && proj, "null pointer");
    svdq::K7Args a;
    a.tensors = tensors; a.numel = numel; a.tile_param = sel_tile_param; a.tile_local = sel_tile_local;
    a.pmask_off = pmask_off; a.has_mask = has_mask; a.packed = packed; a.info = info; a.W = W; a.proj = proj;
    a.tile_elems = tile_elems; a.center = center; a.fp16_basis = fp16_basis; a.invert = region == 1;
    REQUIRE(region == 0 || region == 1, "region must be 0 (masked rows) or 1 (rows outside the mask)");
    cudaError_t e;
    switch (dtype) {
        case svdq::kF32:  e = svdq::k7_launch_dtype<svdq::kF32>(n_tasks, a, (int)n_sel_tiles, (cudaStream_t)stream); break;
        case svdq::kBF16: e = svdq::k7_launch_dtype<svdq::kBF16>(n_tasks, a, (int)n_sel_tiles, (cudaStream_t)stream); break;
        default:          e = svdq::k7_launch_dtype<svdq::kF16>(n_tasks, a, (int)n_sel_tiles, (cudaStream_t)stream); break;
    }
    return finish(__func__, e);
}

int svdq_param_requantize(int n_tasks, int64_t n_params, int rtvq_bits, int rtvq_stages, const int64_t* sel_tile_begin,
                          const float* proj, const uint32_t* present, const int32_t* info, float* coef,
                          uint16_t* chigh, uint8_t* codes, float* qscale, float* qzp, float* qres, float* chat,
                          void* stream) {
    REQUIRE(n_tasks >= 1 && n_tasks <= SVDQ_MAX_TASKS, "n_tasks must be in [1, 32]");
    REQUIRE(rtvq_bits >= 1 && rtvq_bits <= 8, "Low bits must be in [1, 8]");
    REQUIRE(rtvq_stages >= 1 && rtvq_stages <= SVDQ_MAX_STAGES, "RTVQ stages must be in [1, 8]");
    REQUIRE(n_params >= 0 && n_params < (1ll << 31), "n_params");
    if (n_params == 0) return 0;
    REQUIRE(sel_tile_begin && proj && present && info && coef && chigh && codes && qscale && qzp && qres && chat,
            "null pointer");
    svdq::K2RequantArgs a = {};
    a.cfg.n_tasks = n_tasks; a.cfg.bits = rtvq_bits; a.cfg.stages = rtvq_stages;
    a.sel_tile_begin = sel_tile_begin; a.proj = proj; a.present = present; a.info = info; a.coef = coef;
    a.chigh = chigh; a.codes = codes; a.qscale = qscale; a.qzp = qzp; a.qres = qres; a.chat = chat;
    return finish(__func__, svdq::k2_requant_launch(a, (int)n_params, (cudaStream_t)stream));
}

int svdq_reconstruct_merge(int dtype, int n_tasks, int fp16_basis, int diag, int center, int64_t n_tiles,
                           int tile_elems, const void* const* tensors, const int64_t* numel,
                           const int32_t* tile_param, const int32_t* tile_local, const int64_t* pmask_off,
                           const uint8_t* has_mask, const uint32_t* packed, const int32_t* info, const float* W,
                           const float* cbar, const float* gvec, const float* scal, const float* chat,
                           float* const* out, float* diag_partials, const int32_t* noise_info, const float* noise_W,
                           const float* noise_cbar, const float* noise_gvec, const float* noise_scal,
                           float noise_shrink, void* stream) {
    REQUIRE(n_tasks >= 1 && n_tasks <= SVDQ_MAX_TASKS, "n_tasks must be in [1, 32]");
    REQUIRE(dtype >= 0 && dtype <= 2, "dtype");
    REQUIRE(tile_elems > 0 && tile_elems % svdq::kStep == 0, "tile_elems must be a positive multiple of 1024");
    REQUIRE(n_tiles >= 0 && n_tiles < (1ll << 31), "n_tiles");
    if (n_tiles == 0) return 0;
    REQUIRE(tensors && numel && tile_param && tile_local && has_mask && info && W && cbar && gvec && scal && out,
            "null pointer");
    REQUIRE(!diag || (diag_partials && chat), "diag requested without diag_partials / chat");
    svdq::K3Args a;
    a.tensors = tensors; a.numel = numel; a.tile_param = tile_param; a.tile_local = tile_local;
    a.pmask_off = pmask_off; a.has_mask = has_mask; a.packed = packed; a.info = info; a.W = W; a.cbar = cbar;
    a.gvec = gvec; a.scal = scal; a.chat = chat; a.out = out; a.diag = diag_partials; a.tile_elems = tile_elems;
    a.center = center;
    REQUIRE(!noise_info || (noise_W && noise_cbar && noise_gvec && noise_scal && packed && pmask_off),
            "noise region requested without its W / cbar / gvec / scal tables or without packed masks");
    a.info_n = noise_info; a.W_n = noise_W; a.cbar_n = noise_cbar; a.gvec_n = noise_gvec; a.scal_n = noise_scal;
    a.noise_shrink = noise_shrink;
    a.diag_select = 0;
    a.tile_row_off = nullptr; a.u_high = nullptr; a.u_low = nullptr; a.mean_out = nullptr;
    return finish(__func__, k3_launch(dtype, n_tasks, a, (int)n_tiles, fp16_basis != 0, diag != 0,
                                            (cudaStream_t)stream));
}

int svdq_reconstruct_merge_basis(int dtype, int n_tasks, int fp16_basis, int center, int64_t n_tiles, int tile_elems,
                                 const void* const* tensors, const int64_t* numel, const int32_t* tile_param,
                                 const int32_t* tile_local, const int64_t* pmask_off, const uint8_t* has_mask,
                                 const uint32_t* packed, const int32_t* info, const float* W, const float* cbar,
                                 const float* gvec, const float* scal, const float* chat, float* const* out,
                                 float* diag_partials, const int64_t* tile_row_off, void* const* u_high,
                                 void* const* u_low, float* const* mean, void* stream) {
    REQUIRE(n_tasks >= 1 && n_tasks <= 8, "n_tasks must be in [1, 8] (use svdq_reconstruct_merge + svdq_write_basis above)");
    REQUIRE(dtype >= 0 && dtype <= 2, "dtype");
    REQUIRE(tile_elems > 0 && tile_elems % svdq::kStep == 0, "tile_elems must be a positive multiple of 1024");
    REQUIRE(n_tiles >= 0 && n_tiles < (1ll << 31), "n_tiles");
    if (n_tiles == 0) return 0;
    REQUIRE(tensors && numel && tile_param && tile_local && has_mask && info && W && cbar && gvec && scal && out &&
            tile_row_off && u_high && u_low, "null pointer");
    REQUIRE((chat != nullptr) == (diag_partials != nullptr), "chat and diag_partials go together (both NULL = no diagnostics)");
    svdq::K3Args a;
    a.tensors = tensors; a.numel = numel; a.tile_param = tile_param; a.tile_local = tile_local;
    a.pmask_off = pmask_off; a.has_mask = has_mask; a.packed = packed; a.info = info; a.W = W; a.cbar = cbar;
    a.gvec = gvec; a.scal = scal; a.chat = chat; a.out = out; a.diag = diag_partials; a.tile_elems = tile_elems;
    a.center = center;
    a.info_n = nullptr; a.W_n = nullptr; a.cbar_n = nullptr; a.gvec_n = nullptr; a.scal_n = nullptr; a.noise_shrink = 1.0f;
    a.diag_select = 0;
    a.tile_row_off = tile_row_off; a.u_high = u_high; a.u_low = u_low; a.mean_out = mean;
    cudaError_t e;
    switch (dtype) {
        case svdq::kF32:  e = svdq::k3c_launch_dtype<svdq::kF32>(n_tasks, a, (int)n_tiles, fp16_basis != 0, (cudaStream_t)stream); break;
        case svdq::kBF16: e = svdq::k3c_launch_dtype<svdq::kBF16>(n_tasks, a, (int)n_tiles, fp16_basis != 0, (cudaStream_t)stream); break;
        default:          e = svdq::k3c_launch_dtype<svdq::kF16>(n_tasks, a, (int)n_tiles, fp16_basis != 0, (cudaStream_t)stream); break;
    }
    return finish(__func__, e);
}

int svdq_diag_finalize(int n_tasks, int64_t n_params, const float* diag_partials, const int64_t* tile_begin,
                       const int64_t* dm, const int32_t* info, const double* gram_masked, double* out, void* stream) {
    REQUIRE(n_tasks >= 1 && n_tasks <= SVDQ_MAX_TASKS, "n_tasks must be in [1, 32]");
    REQUIRE(n_params >= 0 && n_params < (1ll << 31), "n_params");
    if (n_params == 0) return 0;
    REQUIRE(diag_partials && tile_begin && dm && info && gram_masked && out, "null pointer");
    svdq::K3DiagArgs a;
    a.gram_masked = gram_masked;
    a.diag = diag_partials; a.tile_begin = tile_begin; a.dm = dm; a.info = info; a.out = out; a.nt = n_tasks;
    return finish(__func__, svdq::k3_diag_launch(a, (int)n_params, (cudaStream_t)stream));
}

int svdq_basis_offsets(int64_t n_params, int region, int tile_elems, const uint32_t* count, const int64_t* tile_begin,
                       const int64_t* numel, int64_t* tile_row_off, void* stream) {
    REQUIRE(n_params >= 0 && n_params < (1ll << 31), "n_params");
    REQUIRE(region == 0 || region == 1, "region must be 0 (masked rows) or 1 (rows outside the mask)");
    if (n_params == 0) return 0;
    REQUIRE(count && tile_begin && tile_row_off, "null pointer");
    REQUIRE(region == 0 || (numel && tile_elems > 0), "region 1 needs numel and tile_elems");
    return finish(__func__, svdq::k5_offsets_launch(count, tile_begin, tile_row_off, (int)n_params,
                                                    region == 1 ? numel : nullptr, tile_elems,
                                                    (cudaStream_t)stream));
}

int svdq_write_basis(int dtype, int n_tasks, int fp16_basis, int center, int region, int64_t n_tiles, int tile_elems,
                     const void* const* tensors, const int64_t* numel, const int32_t* tile_param,
                     const int32_t* tile_local, const int64_t* pmask_off, const uint8_t* has_mask,
                     const uint32_t* packed, const int32_t* info, const float* W, const int64_t* tile_row_off,
                     void* const* u_high, void* const* u_low, float* const* mean, void* stream) {
    REQUIRE(n_tasks >= 1 && n_tasks <= SVDQ_MAX_TASKS, "n_tasks must be in [1, 32]");
    REQUIRE(dtype >= 0 && dtype <= 2, "dtype");
    REQUIRE(tile_elems > 0 && tile_elems % svdq::kStep == 0, "tile_elems must be a positive multiple of 1024");
    REQUIRE(n_tiles >= 0 && n_tiles < (1ll << 31), "n_tiles");
    if (n_tiles == 0) return 0;
    REQUIRE(tensors && numel && tile_param && tile_local && has_mask && info && W && tile_row_off && u_high && u_low,
            "null pointer");
    svdq::K5Args a;
    a.tensors = tensors; a.numel = numel; a.tile_param = tile_param; a.tile_local = tile_local;
    a.pmask_off = pmask_off; a.has_mask = has_mask; a.packed = packed; a.info = info; a.W = W;
    a.tile_row_off = tile_row_off; a.u_high = u_high; a.u_low = u_low; a.mean = mean; a.tile_elems = tile_elems;
    a.center = center; a.fp16_basis = fp16_basis; a.invert = region == 1; a.n_tasks = n_tasks;
    REQUIRE(region == 0 || region == 1, "region must be 0 (masked rows) or 1 (rows outside the mask)");
    return finish(__func__, k5_launch(dtype, n_tasks, a, (int)n_tiles, (cudaStream_t)stream));
}

int svdq_mask_tile_counts(int64_t n_tiles, int tile_elems, const int64_t* numel, const int32_t* tile_param,
                          const int32_t* tile_local, const int64_t* pmask_off, const uint8_t* has_mask,
                          const uint32_t* packed, uint32_t* count, void* stream) {
    REQUIRE(tile_elems > 0 && tile_elems % svdq::kStep == 0, "tile_elems must be a positive multiple of 1024");
    REQUIRE(n_tiles >= 0 && n_tiles < (1ll << 31), "n_tiles");
    if (n_tiles == 0) return 0;
    REQUIRE(numel && tile_param && tile_local && has_mask && count, "null pointer");
    svdq::K11Args a = {};
    a.numel = numel; a.tile_param = tile_param; a.tile_local = tile_local; a.pmask_off = pmask_off; a.has_mask = has_mask;
    a.packed = packed; a.count = count; a.tile_elems = tile_elems;
    return finish(__func__, svdq::k11_counts_launch(a, (int)n_tiles, (cudaStream_t)stream));
}

int svdq_reload_merge(int basis_fp16, int n_tasks, int region, float scale, int64_t n_tiles, int tile_elems,
                      const int64_t* numel, const int32_t* tile_param, const int32_t* tile_local,
                      const int64_t* pmask_off, const uint8_t* has_mask, const uint32_t* packed, const int32_t* kr,
                      const void* const* u_high, const void* const* u_low, const float* const* mean, const float* cbar,
                      const int64_t* tile_row_off, float* const* out, void* stream) {
    REQUIRE(n_tasks >= 1 && n_tasks <= SVDQ_MAX_TASKS, "n_tasks must be in [1, 32]");
    REQUIRE(region == 0 || region == 1, "region");
    REQUIRE(tile_elems > 0 && tile_elems % svdq::kStep == 0, "tile_elems must be a positive multiple of 1024");
    REQUIRE(n_tiles >= 0 && n_tiles < (1ll << 31), "n_tiles");
    if (n_tiles == 0) return 0;
    REQUIRE(numel && tile_param && tile_local && has_mask && kr && u_high && u_low && cbar && tile_row_off && out, "null pointer");
    svdq::K11Args a = {};
    a.numel = numel; a.tile_param = tile_param; a.tile_local = tile_local; a.pmask_off = pmask_off; a.has_mask = has_mask;
    a.packed = packed; a.kr = kr; a.u_high = u_high; a.u_low = u_low; a.mean = mean; a.cbar = cbar;
    a.tile_row_off = tile_row_off; a.out = out; a.tile_elems = tile_elems; a.n_tasks = n_tasks; a.region = region;
    a.scale = scale;
    return finish(__func__, svdq::k11_merge_launch(a, (int)n_tiles, basis_fp16 != 0, (cudaStream_t)stream));
}

int svdq_rtvq_quantize(const float* x, int64_t n, int bits, int stages, void* codes, int64_t codes_ld,
                       int code_bytes, float* scale, float* zp, float* resnorm, void* scratch, uint32_t* packed,
                       int64_t packed_ld, void* stream) {
    REQUIRE(n >= 0, "n");
    if (n == 0) return 0;
    REQUIRE(bits >= 1 && bits <= 16, "bits must be in [1, 16]");
    REQUIRE(stages >= 1 && stages <= SVDQ_MAX_STAGES, "RTVQ stages must be in [1, 8]");
    REQUIRE(code_bytes == 1 || code_bytes == 2, "code_bytes");
    REQUIRE(code_bytes == 2 || bits <= 8, "uint8 codes need bits <= 8");
    REQUIRE(codes_ld >= n && codes_ld % 4 == 0, "codes_ld must be >= n and a multiple of 4");
    REQUIRE(x && codes && scale && zp && resnorm && scratch, "null pointer");
    REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)codes & 7) == 0, "x must be 16-byte, codes 8-byte aligned");
    REQUIRE(!packed || packed_ld * 32 >= n * bits, "packed_ld too small");
    return finish(__func__, svdq::k4_rtvq_launch(x, n, bits, stages, codes, codes_ld, code_bytes, scale, zp, resnorm,
                                                 (svdq::K4Stats*)scratch, packed, packed_ld, (cudaStream_t)stream));
}

int svdq_rtvq_dequantize(const void* codes, int64_t codes_ld, int code_bytes, int stages, int64_t n,
                         const float* scale, const float* zp, float* out, void* stream) {
    REQUIRE(n >= 0, "n");
    if (n == 0) return 0;
    REQUIRE(stages >= 1 && stages <= SVDQ_MAX_STAGES, "stages must be in [1, 8]");
    REQUIRE(code_bytes == 1 || code_bytes == 2, "code_bytes");
    REQUIRE(codes && scale && zp && out, "null pointer");
    return finish(__func__, svdq::k4_dequant_launch(codes, codes_ld, code_bytes, stages, n, scale, zp, out,
                                                    (cudaStream_t)stream));
}

int svdq_absmax_quantize(const float* x, int64_t n, int bits, void* q, int code_bytes, float* scale, void* scratch,
                         void* stream) {
    REQUIRE(n >= 0, "n");
    if (n == 0) return 0;
    REQUIRE(bits >= 2 && bits <= 16, "bits must be in [2, 16]");
    REQUIRE(code_bytes == 1 || code_bytes == 2, "code_bytes");
    REQUIRE(x && q && scale && scratch, "null pointer");
    return finish(__func__, svdq::k4_absmax_launch(x, n, bits, q, code_bytes, scale, (svdq::K4Stats*)scratch,
                                                   (cudaStream_t)stream));
}

int svdq_combine_masks(const uint8_t* const* masks, int n_masks, int64_t n, int strategy, uint8_t* out, void* stream) {
    REQUIRE(n_masks >= 1, "Empty mask list");
    REQUIRE(n_masks <= SVDQ_MAX_TASKS, "at most 32 masks");
    REQUIRE(strategy >= 0 && strategy <= 2, "Unknown mask strategy");
    REQUIRE(n >= 0, "n");
    if (n == 0) return 0;
    REQUIRE(masks && out, "null pointer");
    return finish(__func__, svdq::combine_masks_launch(masks, n_masks, n, strategy, out, (cudaStream_t)stream));
}

int svdq_unpack_mask(const uint32_t* packed, int64_t n, uint8_t* out, void* stream) {
    REQUIRE(n >= 0, "n");
    if (n == 0) return 0;
    REQUIRE(packed && out, "null pointer");
    return finish(__func__, svdq::unpack_mask_launch(packed, n, out, (cudaStream_t)stream));
}

size_t svdq_project_scratch_bytes(void) { return svdq::k14_project_scratch_bytes(); }
int svdq_select_chunk_elems(void) { return svdq::k14_select_chunk(); }

int svdq_basis_project(int basis_fp16, const void* u, int64_t ld, int cols, int64_t rows, const float* delta,
                       const float* mean, float* c, void* scratch, void* stream) {
    REQUIRE(cols >= 1 && cols <= 32, "cols must be 1..32 (call once per block of 32 columns)");
    REQUIRE(rows >= 0 && ld >= cols, "rows / ld");
    REQUIRE(c && scratch, "null pointer");
    REQUIRE(rows == 0 || (u && delta), "null pointer");
    return finish(__func__, svdq::k14_project_launch(basis_fp16 != 0, u, ld, cols, rows, delta, mean, c, scratch,
                                                     (cudaStream_t)stream));
}

int svdq_basis_expand(int basis_fp16, const void* u_high, int64_t ld_high, int k, const void* u_low, int64_t ld_low,
                      int n_low, int64_t rows, const float* c_high, const float* c_low, const float* mean, float scale,
                      float* out, void* stream) {
    REQUIRE(k >= 0 && n_low >= 0 && k + n_low <= 8192, "column counts");
    REQUIRE(rows >= 0, "rows");
    if (rows == 0) return 0;
    REQUIRE(out, "null pointer");
    REQUIRE(k == 0 || (u_high && c_high && ld_high >= k), "high block");
    REQUIRE(n_low == 0 || (u_low && c_low && ld_low >= n_low), "low block");
    return finish(__func__, svdq::k14_expand_launch(basis_fp16 != 0, u_high, ld_high, k, u_low, ld_low, n_low, rows,
                                                    c_high, c_low, mean, scale, out, (cudaStream_t)stream));
}

int svdq_mask_offsets(const uint8_t* mask, int64_t n, int invert, int64_t* chunk_off, void* stream) {
    REQUIRE(n >= 0 && chunk_off, "n / chunk_off");
    REQUIRE(n == 0 || mask, "null pointer");
    return finish(__func__, svdq::k14_mask_offsets_launch(mask, n, invert != 0, chunk_off, (cudaStream_t)stream));
}

int svdq_mask_select(const void* x, int elem_bytes, const uint8_t* mask, int64_t n, int invert, const int64_t* chunk_off,
                     void* out, void* stream) {
    REQUIRE(n >= 0, "n");
    if (n == 0) return 0;
    REQUIRE(x && mask && chunk_off && out, "null pointer");
    return finish(__func__, svdq::k14_select_launch(false, elem_bytes, x, mask, n, invert != 0, chunk_off, out,
                                                    (cudaStream_t)stream));
}

int svdq_mask_scatter(const void* values, int elem_bytes, const uint8_t* mask, int64_t n, int invert,
                      const int64_t* chunk_off, void* out, void* stream) {
    REQUIRE(n >= 0, "n");
    if (n == 0) return 0;
    REQUIRE(values && mask && chunk_off && out, "null pointer");
    return finish(__func__, svdq::k14_select_launch(true, elem_bytes, values, mask, n, invert != 0, chunk_off, out,
                                                    (cudaStream_t)stream));
}

}  // extern "C"
