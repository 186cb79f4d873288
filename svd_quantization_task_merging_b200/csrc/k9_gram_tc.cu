// K9 — pass 1 (task vectors + tall-mask combination + masked Gram, one partial per tile) for 16-BIT inputs on the
// 5th-generation tensor cores (tcgen05.mma kind::f16, accumulators in TMEM).  Same inputs, outputs and reference
// lines as k1s_tv_mask_gram_staged.cu; used for bf16 / fp16 checkpoints with up to 8 task vectors and no second
// Gram block.  Why: with 2-byte elements pass 1 moves 18 B per element but the CUDA-core Gram still costs 36 packed
// FMAs + converts per element, so the fp32-FMA kernel is issue-bound at ~50 % of the HBM roofline
// (profiles/r1_ncu_full_llama_bf16_k1s_k3s.csv).  A task vector bf16(ft - base) is EXACTLY a bf16 number, so the
// Gram of the rounded task vectors is a plain bf16 x bf16 -> fp32 tensor-core product: no split, no loss.
//
// Mapping ("sliced block-diagonal Gram").  The Gram G = T^T T of the tall-skinny task matrix T [D x 8] has only
// 8 x 8 outputs, far below the 128 x N tile of one tcgen05.mma.  A 1024-element chunk is therefore cut into 16
// slices of 64 elements; rows of the MMA operand are (slice s, task i), so that
//     D[(s,i),(s',j)] = sum_k T[64 s + k][i] * T[64 s' + k][j],        M = N = 128, K = 16 per instruction,
// and the 16 diagonal 8 x 8 blocks (s == s') are the Gram partials of the slices; the off-diagonal blocks are
// unused (the tensor pipe has 0.25 cycles per element of work, the HBM budget is 0.77).  A and B are the SAME
// shared-memory tile: [group of 8 elements][task][8 elements] = canonical K-major, no swizzle (core matrix = 8 tasks
// x 16 B, LBO = 128 B to the next 8 elements, SBO = 1024 B to the next slice), so one descriptor serves both.
//
// Accuracy.  The tensor core adds into its fp32 accumulator with truncation (measured on B200: -3e-8 relative
// per accumulate, scratch/tcprobe), so an accumulator only ever chains the 4 instructions of ONE chunk (64-element
// dot products); the partials are then summed by CUDA cores in round-to-nearest fp32 in a fixed order (16 slices
// pairwise, 16 chunks per tile) and across tiles in fp64 by k2_gram_reduce as before.
//
// Roles (one persistent CTA per SM, 14 warps): warps 0-7 transform (raw ring -> masked bf16 task vectors in the MMA
// tile; tall-mask vote, packed mask, counts), warp 8 TMA producer (cp.async.bulk ring as in K1 staged), warp 9 issues
// the MMAs (one thread), warps 10-13 drain the accumulators (tcgen05.ld: a warp reads the TMEM lanes 32 (w % 4) ..).
// Bound: HBM (18 B per element at N = 8).
#include <type_traits>

#include "stage_pipe.cuh"
#include "svdq_kernels.h"
#include "tc_common.cuh"

#ifndef SVDQ_DTYPE
#define SVDQ_DTYPE 1
#endif

namespace svdq {

#if SVDQ_DTYPE != 0

constexpr int kTcMaxStages = 4;              // raw ring depth (stages of kTcStageElems elements): as many as fit next to
                                             // the fixed regions -- 4 without masks or with bit-packed masks, 3 with byte
                                             // masks (launch_tc); a 4th stage is worth 4 % on the mask-free Llama shard
constexpr int kTcStageElems = 2 * kStep;     // one raw stage = two 1024-element chunks: 4 KB per bulk copy (the producer
                                             // issues its copies one at a time, so 2 KB copies capped it near 4.7 TB/s)
constexpr int kTcTiles = 3;                  // MMA tile ring depth (accumulators: 2)
constexpr int kTcRowStride = kTcStageElems * 2 + 16;   // bytes between the tensors of a raw stage: +16 B skews the banks
                                             // so that the 8 task lanes of a quarter-warp read 8 different bank groups
// bytes between the task masks of a raw stage (runtime: 0 without masks, kTcStageElems / 8 for bit-packed masks,
// kTcStageElems for byte masks)
constexpr int kTcTileBytes = kStep * 8 * 2;  // MMA tile of one chunk: 1024 elements x 8 task rows x 2 B = 16 KB
constexpr int kTcTransform = 256;            // transform threads (warps 0-7)
constexpr int kTcThreads = 14 * 32;
constexpr int kTcTmemCols = 256;             // two 128-column accumulators

constexpr int kTcMaskOff = (9 * kTcRowStride + 15) / 16 * 16;      // the masks follow the 9 tensor rows of a stage
__host__ __device__ constexpr int tc_stage_bytes(int mask_stride) { return kTcMaskOff + 8 * mask_stride; }
// tiles + mask LUT + drain partials + barriers / flags; the raw ring follows
constexpr int kTcFixedBytes = kTcTiles * kTcTileBytes + 4096 + 4 * 64 * 4 * 2 + 512;
constexpr int kTcSmemLimit = 227 * 1024 - 1024;      // opt-in limit minus the static __shared__ variables

template <typename T> __device__ __forceinline__ uint32_t load_raw16(const void* p, int64_t e) {
    return (uint32_t) * (reinterpret_cast<const uint16_t*>(p) + e);
}

template <typename T, int NT>
__global__ void __launch_bounds__(kTcThreads, 1) k9_gram_tc(const K1Args a, const int n_tiles, const int n_stages,
                                                            const int mask_stride) {
    constexpr int G = tri_count(NT);
    constexpr int kMaskOff = kTcMaskOff;
    const uint32_t STAGES = (uint32_t)n_stages;
    const int kStageBytes = tc_stage_bytes(mask_stride);

    // (no swizzle: the MMA descriptors need 16-byte alignment only; keeping the pointer arithmetic on the __shared__
    // array itself lets the compiler emit LDS / STS instead of generic loads and stores)
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* tile_buf = smem;                                     // kTcTiles x 16 KB MMA tiles
    uint4* lut = reinterpret_cast<uint4*>(tile_buf + kTcTiles * kTcTileBytes);    // byte -> 8 x 16-bit lane masks
    unsigned char* stage_base = smem + kTcFixedBytes;                   // raw ring: n_stages x kStageBytes
    float* s_part = reinterpret_cast<float*>(lut + 256);                // [2][4][64] per-drain-warp Gram partials
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_part + 2 * 4 * 64);
    uint64_t* full = bars;                  // [kTcMaxStages] producer -> transform
    uint64_t* empty = full + kTcMaxStages;  // [kTcMaxStages] transform -> producer
    uint64_t* tfull = empty + kTcMaxStages; // [kTcTiles] transform -> MMA (tile written)
    uint64_t* tempty = tfull + kTcTiles;    // [kTcTiles] MMA -> transform (tile read)
    uint64_t* afull = tempty + kTcTiles;    // [2] MMA -> drain (accumulator complete)
    uint64_t* aempty = afull + 2;           // [2] drain -> MMA (accumulator read)
    int* s_direct = reinterpret_cast<int*>(aempty + 2);
    __shared__ const void* s_ptr[NT + 1];
    __shared__ const uint8_t* s_mask[NT];
    __shared__ uint32_t s_cnt[2 * (kTcTransform / 32)];
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < kTcMaxStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kTcTransform / 32); }
        for (int s = 0; s < kTcTiles; ++s) { mbar_init(&tfull[s], kTcTransform / 32); mbar_init(&tempty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&afull[s], 1); mbar_init(&aempty[s], 4); }
        mbar_fence_init();
    }
    if (tid < 256) {        // expansion table: bit c of the byte -> 16-bit lane c all ones
        uint32_t w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
            w[q] = (((uint32_t)tid >> (2 * q)) & 1u ? 0x0000FFFFu : 0u) | (((uint32_t)tid >> (2 * q + 1)) & 1u ? 0xFFFF0000u : 0u);
        lut[tid] = make_uint4(w[0], w[1], w[2], w[3]);
    }
    if (warp == 9) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(kTcTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;

    if (warp == 8) {
        // ================= TMA producer: lane i owns stream i (tensor i for i <= NT, mask i-NT-1 after) ==============
        PipeState ps;
        const bool is_tensor = lane <= NT;
        const bool is_mask = lane > NT && lane <= 2 * NT;
        const bool bits_in = a.mask_bits != 0;
        const int kMaskBytes = bits_in ? kTcStageElems / 8 : kTcStageElems;     // one task's mask bytes per stage
        const int my_bytes = is_tensor ? kTcStageElems * 2 : kMaskBytes;
        const int my_off = is_tensor ? lane * kTcRowStride : kMaskOff + (lane - NT - 1) * mask_stride;
        // Per-parameter values are re-read only when the CTA's next tile belongs to another parameter (consecutive
        // tiles of a CTA mostly come from the same large tensor): the dependent global loads of a tile boundary cost
        // every role a few hundred nanoseconds per tile otherwise.
        int prev_p = -1, n_present = 0;
        int64_t numel = 0;
        const unsigned char* my_ptr = nullptr;
        int p_ahead = a.tile_param[blockIdx.x], local_ahead = a.tile_local[blockIdx.x];    // grid <= n_tiles
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int p = p_ahead;                      // tile -> (parameter, position) is read one tile ahead
            const int64_t start = (int64_t)local_ahead * a.tile_elems;
            if (tile + (int)gridDim.x < n_tiles) {
                p_ahead = a.tile_param[tile + gridDim.x];
                local_ahead = a.tile_local[tile + gridDim.x];
            }
            if (p != prev_p) {
                prev_p = p;
                numel = a.numel[p];
                my_ptr = nullptr;
                if (is_tensor) {
                    const void* q = a.tensors[(int64_t)p * (NT + 1) + lane];
                    my_ptr = reinterpret_cast<const unsigned char*>(q ? q : a.tensors[(int64_t)p * (NT + 1)]);
                } else if (is_mask && a.masks) {
                    my_ptr = a.masks[(int64_t)p * NT + (lane - NT - 1)];
                }
                n_present = __popc(__ballot_sync(0xffffffffu, is_mask && my_ptr != nullptr));
            }
            const int64_t stop = min(start + (int64_t)a.tile_elems, numel);
            for (int64_t e0 = start; e0 < stop; e0 += kTcStageElems) {
                if (lane == 0) mbar_wait(&empty[ps.stage], ps.phase ^ 1u);
                __syncwarp();
                unsigned char* sb = stage_base + ps.stage * kStageBytes;
                if (e0 + kTcStageElems <= numel) {
                    if (lane == 0) {
                        s_direct[ps.stage] = 0;
                        mbar_arrive_expect_tx(&full[ps.stage], (uint32_t)((NT + 1) * kTcStageElems * 2 + n_present * kMaskBytes));
                    }
                    __syncwarp();
                    if (my_ptr != nullptr)
                        bulk_g2s(sb + my_off, my_ptr + (is_tensor ? e0 * 2 : (bits_in ? e0 / 8 : e0)), my_bytes, &full[ps.stage]);
                } else if (lane == 0) {
                    s_direct[ps.stage] = 1;          // tail of the parameter: the transform warps load it themselves
                    mbar_arrive(&full[ps.stage]);
                }
                ps.advance_n(STAGES);
            }
        }
    } else if (warp == 9) {
        // ================= MMA issuer: the whole warp walks the loops, one elected lane issues ==========================
        // (uniform control flow: the compiler needs no divergent-operand loops around the tcgen05 instructions)
        const bool leader = elect_one();
        PipeState tb, ab;                               // tile-buffer ring, accumulator ring
        const uint32_t idesc = tc_idesc_f16<T>(128, 128);
        const uint64_t desc0 = tc_smem_desc(smem_u32(tile_buf), 128, 1024);
        int prev_p = -1;
        int64_t numel = 0;
        int p_ahead = a.tile_param[blockIdx.x], local_ahead = a.tile_local[blockIdx.x];    // grid <= n_tiles
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int p = p_ahead;                      // tile -> (parameter, position) is read one tile ahead
            const int64_t start = (int64_t)local_ahead * a.tile_elems;
            if (tile + (int)gridDim.x < n_tiles) {
                p_ahead = a.tile_param[tile + gridDim.x];
                local_ahead = a.tile_local[tile + gridDim.x];
            }
            if (p != prev_p) { prev_p = p; numel = a.numel[p]; }
            const int64_t stop = min(start + (int64_t)a.tile_elems, numel);
            for (int64_t e0 = start; e0 < stop; e0 += kStep) {
                mbar_wait(&tfull[tb.stage], tb.phase);
                mbar_wait(&aempty[ab.stage], ab.phase ^ 1u);
                tc_fence_after();
                if (leader) {
                    const uint64_t d0 = desc0 + (uint64_t)(tb.stage * (kTcTileBytes >> 4));
#pragma unroll
                    for (int k = 0; k < 4; ++k) {       // 16 elements of every slice per instruction
                        const uint64_t d = d0 + (uint64_t)(k * (256 >> 4));
                        tc_mma_f16(tmem + ab.stage * 128, d, d, idesc, k ? 1u : 0u);
                    }
                    tc_commit(&tempty[tb.stage]);       // tile buffer may be overwritten once these MMAs have read it
                    tc_commit(&afull[ab.stage]);        // accumulator complete
                }
                __syncwarp();
                tb.advance<kTcTiles>();
                ab.advance<2>();
            }
        }
    } else if (warp >= 10) {
        // ================= drain warps: TMEM -> registers -> per-tile Gram partial =================================
        const int q = warp & 3;                         // TMEM lane quadrant this warp may read
        const int sl = lane >> 3;                       // slice within the quadrant; row i = lane & 7
        PipeState tb;
        uint32_t tile_par = 0;
        int prev_p = -1;
        int64_t numel = 0;
        int p_ahead = a.tile_param[blockIdx.x], local_ahead = a.tile_local[blockIdx.x];    // grid <= n_tiles
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int p = p_ahead;                      // tile -> (parameter, position) is read one tile ahead
            const int64_t start = (int64_t)local_ahead * a.tile_elems;
            if (tile + (int)gridDim.x < n_tiles) {
                p_ahead = a.tile_param[tile + gridDim.x];
                local_ahead = a.tile_local[tile + gridDim.x];
            }
            if (p != prev_p) { prev_p = p; numel = a.numel[p]; }
            const int64_t stop = min(start + (int64_t)a.tile_elems, numel);
            float acc[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = 0.0f;
            for (int64_t e0 = start; e0 < stop; e0 += kStep) {
                mbar_wait(&afull[tb.stage], tb.phase);
                tc_fence_after();
                uint32_t r[32];
                tc_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(tb.stage * 128 + q * 32), r);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&aempty[tb.stage]);
                tb.advance<2>();
                // this lane's row (slice 4q + sl, task i) sits in columns 8 sl .. 8 sl + 7 of the 32 just read
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint32_t lo = sl & 1 ? r[8 + j] : r[j], hi = sl & 1 ? r[24 + j] : r[16 + j];
                    float v = __uint_as_float(sl & 2 ? hi : lo);
                    v += __shfl_xor_sync(0xffffffffu, v, 8);
                    v += __shfl_xor_sync(0xffffffffu, v, 16);
                    acc[j] += v;
                }
            }
            float* part = s_part + tile_par * 256 + q * 64;
            if (lane < 8) {
#pragma unroll
                for (int j = 0; j < 8; ++j) part[lane * 8 + j] = acc[j];
            }
            named_bar_sync(2, 128);
            if (q == 0 && lane < 32) {
                float* gout = a.gram + (int64_t)tile * G;
                const float* pp = s_part + tile_par * 256;
                for (int idx = lane; idx < G; idx += 32) {
                    int i = 0, rem = idx;
                    while (rem >= NT - i) { rem -= NT - i; ++i; }
                    const int j = i + rem;
                    gout[idx] = ((pp[i * 8 + j] + pp[64 + i * 8 + j]) + (pp[128 + i * 8 + j] + pp[192 + i * 8 + j]));
                }
            }
            tile_par ^= 1u;       // the other half of s_part is free: its readers passed the barrier above one tile ago
        }
    } else {
        // ================= transform warps ==========================================================================
        PipeState ps, tb;
        const int t = lane & 7;                         // task row of this lane
        int pending = -1;                               // tile buffer written but not yet handed to the MMA warp
        const bool majority = a.strategy == kMajority;
        const bool mask_bits = a.mask_bits != 0;
        int prev_p = -1;
        int64_t numel = 0;
        bool has_mask = false;
        uint32_t thr_bytes = 0, tile_par = 0;
        uint32_t* packed = nullptr;
        int p_ahead = a.tile_param[blockIdx.x], local_ahead = a.tile_local[blockIdx.x];    // grid <= n_tiles
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int p = p_ahead;                      // tile -> (parameter, position) is read one tile ahead
            const int64_t start = (int64_t)local_ahead * a.tile_elems;
            if (tile + (int)gridDim.x < n_tiles) {
                p_ahead = a.tile_param[tile + gridDim.x];
                local_ahead = a.tile_local[tile + gridDim.x];
            }
            if (p != prev_p) {                          // same decision in every transform warp (same tile sequence)
                prev_p = p;
                numel = a.numel[p];
                named_bar_sync(1, kTcTransform);        // previous parameter finished with s_ptr / s_mask
                if (tid <= NT) {
                    const void* qp = a.tensors[(int64_t)p * (NT + 1) + tid];
                    s_ptr[tid] = qp ? qp : a.tensors[(int64_t)p * (NT + 1)];
                }
                if (tid < NT) s_mask[tid] = a.masks ? a.masks[(int64_t)p * NT + tid] : nullptr;
                named_bar_sync(1, kTcTransform);
                int n_present = 0;
#pragma unroll
                for (int u = 0; u < NT; ++u) n_present += s_mask[u] != nullptr;
                has_mask = n_present > 0;
                thr_bytes = 0x01010101u * (uint32_t)(a.strategy == kUnion ? 1 : n_present);
                packed = has_mask ? a.packed + a.pmask_off[p] : nullptr;
            }
            const int64_t stop = min(start + (int64_t)a.tile_elems, numel);
            uint32_t cnt = 0;

            for (int64_t s0 = start; s0 < stop; s0 += kTcStageElems) {
                mbar_wait(&full[ps.stage], ps.phase);
                const unsigned char* sbase = stage_base + ps.stage * kStageBytes;
                const bool direct = s_direct[ps.stage] != 0;
#pragma unroll 1
                for (int sub = 0; sub < kTcStageElems / kStep; ++sub) {
                    const int64_t e0 = s0 + (int64_t)sub * kStep;
                    if (e0 >= stop) break;
                    const bool last_sub = sub == kTcStageElems / kStep - 1 || e0 + kStep >= stop;
                    const unsigned char* sb = sbase + sub * (kStep * 2);          // this chunk inside the tensors' rows
                    unsigned char* tile_out = tile_buf + tb.stage * kTcTileBytes;
                    // Warp w owns elements 128 w .. 128 w + 127 of the chunk for every step, so the transform warps
                    // never wait for one another inside a tile.
                    if (!has_mask && !direct) {
                        // ---- fast path (whole chunk, no masks): all loads first; the async-proxy fence + hand-over of
                        // the PREVIOUS chunk's tile overlaps their latency
                        uint4 b[4], f[4];
#pragma unroll
                        for (int it = 0; it < 4; ++it) {
                            const int g = warp * 16 + it * 4 + (lane >> 3);
                            b[it] = *reinterpret_cast<const uint4*>(sb + g * 16);
                            f[it] = *reinterpret_cast<const uint4*>(sb + (t + 1) * kTcRowStride + g * 16);
                        }
                        if (pending >= 0) {
                            fence_async_smem();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&tfull[pending]);
                            pending = -1;
                        }
                        uint4 d[4];
#pragma unroll
                        for (int it = 0; it < 4; ++it) {
                            d[it] = make_uint4(0u, 0u, 0u, 0u);
                            if (t < NT) {
                                d[it].x = sub2<T>(f[it].x, b[it].x); d[it].y = sub2<T>(f[it].y, b[it].y);
                                d[it].z = sub2<T>(f[it].z, b[it].z); d[it].w = sub2<T>(f[it].w, b[it].w);
                            }
                        }
                        if (last_sub) {                  // the raw stage lives in registers now: hand it back
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&empty[ps.stage]);
                        }
                        mbar_wait(&tempty[tb.stage], tb.phase ^ 1u);   // the MMAs that read this tile buffer have finished
#pragma unroll
                        for (int it = 0; it < 4; ++it) {
                            const int g = warp * 16 + it * 4 + (lane >> 3);
                            *reinterpret_cast<uint4*>(tile_out + g * 128 + t * 16) = d[it];
                        }
                        cnt += kVec;                    // every element of the chunk counts
                        pending = (int)tb.stage;
                        tb.advance<kTcTiles>();
                        continue;
                    }
                    if (pending >= 0) {
                        fence_async_smem();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tfull[pending]);
                        pending = -1;
                    }
                    mbar_wait(&tempty[tb.stage], tb.phase ^ 1u);       // the MMAs that read this tile buffer have finished
                    // ---- combined mask: lane -> 4 consecutive elements; the 8 lanes of an octet end up with the same word
                    uint32_t w = 0xFFFFFFFFu;
                    {
                        const int64_t e = e0 + (int64_t)warp * 128 + lane * kVec;
                        const bool active = e < stop;
                        uint32_t bits = 0;
                        if (active) {
                            const uint32_t valid = (e + kVec <= numel) ? 0xFu : ((1u << (int)(numel - e)) - 1u);
                            if (has_mask) {
                                const int mt = warp * 32 + lane;       // this thread's word / nibble index in the chunk
                                const unsigned char* mb = sbase + kMaskOff + sub * (mask_bits ? kStep / 8 : kStep);
                                uint32_t votes = 0;
#pragma unroll
                                for (int u = 0; u < NT; ++u) {
                                    uint32_t mw = 0u;
                                    if (s_mask[u] != nullptr) {
                                        if (!direct) {
                                            mw = !mask_bits ? *reinterpret_cast<const uint32_t*>(mb + u * mask_stride + mt * 4)
                                                            : nibble_to_bytes((*reinterpret_cast<const uint32_t*>(mb + u * mask_stride + (mt >> 3) * 4)
                                                                               >> ((mt & 7) * 4)) & 0xFu);
                                        } else if (mask_bits) {
                                            mw = nibble_to_bytes((__ldg(reinterpret_cast<const uint32_t*>(s_mask[u]) + (e >> 5)) >> (int)(e & 31)) & 0xFu);
                                        } else {
#pragma unroll
                                            for (int c = 0; c < kVec; ++c)
                                                if (e + c < numel) mw |= (uint32_t)__ldg(s_mask[u] + e + c) << (8 * c);
                                        }
                                    }
                                    votes += __vminu4(mw, 0x01010101u);
                                }
                                if (majority) votes += votes;
                                const uint32_t ge = __vcmpgeu4(votes, thr_bytes);
                                bits = ((ge >> 7) & 1u) | ((ge >> 14) & 2u) | ((ge >> 21) & 4u) | ((ge >> 28) & 8u);
                                bits &= valid;
                            } else {
                                bits = valid;
                            }
                        }
                        cnt += __popc(bits);
                        w = bits << ((lane & 7) * 4);
                        w |= __shfl_xor_sync(0xffffffffu, w, 1);
                        w |= __shfl_xor_sync(0xffffffffu, w, 2);
                        w |= __shfl_xor_sync(0xffffffffu, w, 4);
                        if ((lane & 7) == 0 && has_mask && active) packed[e >> 5] = w;
                    }
                    // ---- masked task vectors into the MMA tile: lane -> (task t, group of 8 elements) -------------
#pragma unroll
                    for (int it = 0; it < 4; ++it) {
                        const int g = warp * 16 + it * 4 + (lane >> 3);
                        // elements 8 g .. 8 g + 7 sit in byte (lane >> 3) of the word held by octet `it`
                        const uint32_t mbyte = (__shfl_sync(0xffffffffu, w, it * 8) >> ((lane >> 3) * 8)) & 0xFFu;
                        uint4 d = make_uint4(0u, 0u, 0u, 0u);
                        if (t < NT && mbyte != 0u) {
                            uint4 b, f;
                            if (!direct) {
                                b = *reinterpret_cast<const uint4*>(sb + g * 16);
                                f = *reinterpret_cast<const uint4*>(sb + (t + 1) * kTcRowStride + g * 16);
                            } else {
                                const int64_t e = e0 + (int64_t)g * 8;
                                uint32_t bw[4] = {0u, 0u, 0u, 0u}, fw[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                                for (int c = 0; c < 8; ++c) {
                                    if (e + c < numel) {
                                        bw[c >> 1] |= load_raw16<T>(s_ptr[0], e + c) << (16 * (c & 1));
                                        fw[c >> 1] |= load_raw16<T>(s_ptr[t + 1], e + c) << (16 * (c & 1));
                                    }
                                }
                                b = make_uint4(bw[0], bw[1], bw[2], bw[3]);
                                f = make_uint4(fw[0], fw[1], fw[2], fw[3]);
                            }
                            d.x = sub2<T>(f.x, b.x); d.y = sub2<T>(f.y, b.y); d.z = sub2<T>(f.z, b.z); d.w = sub2<T>(f.w, b.w);
                            if (mbyte != 0xFFu) {
                                const uint4 m = lut[mbyte];
                                d.x &= m.x; d.y &= m.y; d.z &= m.z; d.w &= m.w;
                            }
                        }
                        *reinterpret_cast<uint4*>(tile_out + g * 128 + t * 16) = d;
                    }
                    if (last_sub) {
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&empty[ps.stage]);
                    }
                    pending = (int)tb.stage;
                    tb.advance<kTcTiles>();
                }
                ps.advance_n(STAGES);
            }
            cnt = __reduce_add_sync(0xffffffffu, cnt);
            uint32_t* sc = s_cnt + tile_par * (kTcTransform / 32);     // two sets: the other one may still be read
            if (lane == 0) sc[warp] = cnt;
            named_bar_sync(1, kTcTransform);
            if (tid == 0) {
                uint32_t c = 0;
#pragma unroll
                for (int w = 0; w < kTcTransform / 32; ++w) c += sc[w];
                a.count[tile] = c;
            }
            tile_par ^= 1u;
        }
        if (pending >= 0) {
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tfull[pending]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 9) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTcTmemCols));
}

template <typename T, int NT>
static cudaError_t launch_tc(const K1Args& a, int n_tiles, int n_sm, cudaStream_t st) {
    if (n_tiles <= 0) return cudaSuccess;
    const int mask_stride = a.masks == nullptr ? 0 : (a.mask_bits ? kTcStageElems / 8 : kTcStageElems);
    int n_stages = (kTcSmemLimit - kTcFixedBytes) / tc_stage_bytes(mask_stride);
    if (n_stages > kTcMaxStages) n_stages = kTcMaxStages;
    const int smem = kTcFixedBytes + n_stages * tc_stage_bytes(mask_stride);
    const int grid = n_tiles < n_sm ? n_tiles : n_sm;
    cudaError_t e = cudaFuncSetAttribute(k9_gram_tc<T, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemLimit);
    if (e != cudaSuccess) return e;
    k9_gram_tc<T, NT><<<grid, kTcThreads, smem, st>>>(a, n_tiles, n_stages, mask_stride);
    return cudaGetLastError();
}

#endif  // 16-bit dtypes

// tensor-core pass 1 exists for 16-bit inputs, nt <= 8, single Gram block; cudaErrorNotSupported otherwise
template <>
cudaError_t k9_launch_dtype<SVDQ_DTYPE>(int nt, const K1Args& a, int n_tiles, int n_sm, cudaStream_t st) {
#if SVDQ_DTYPE != 0
    using T = DTypeOf<SVDQ_DTYPE>::type;
    if (a.packed_in != nullptr) return cudaErrorNotSupported;
    switch (nt) {
#define SVDQ_CASE(N) case N: return launch_tc<T, N>(a, n_tiles, n_sm, st);
        SVDQ_CASE(1) SVDQ_CASE(2) SVDQ_CASE(3) SVDQ_CASE(4) SVDQ_CASE(5) SVDQ_CASE(6) SVDQ_CASE(7) SVDQ_CASE(8)
#undef SVDQ_CASE
        default: return cudaErrorNotSupported;
    }
#else
    return cudaErrorNotSupported;
#endif
}

}  // namespace svdq
