// K12 — single-pass masked Gram for 9..21 FP32 task vectors on the 5th-generation tensor cores (tcgen05.mma
// kind::f16 on a 3-piece bf16 split, accumulators in TMEM).  Same inputs, outputs and reference lines as
// k8_gram_staged.cu (compute_task_vector src/svd_hybrid/task_vector_loader.py:142, apply_mask_to_tensor
// src/svd_hybrid/mask_loader.py:675-679, stack_and_center + the T^T T half of torch.linalg.svd
// src/svd_hybrid/basis.py:103-111,241); the combined mask comes pre-packed from k6_mask_pack.
//
// Why.  At N = 20 the CUDA-core Gram costs 300 FMAs per element out of shared memory and runs at 38 % of the HBM
// roofline (profiles/r1_ncu_full_wide_k6_k8.csv: issue-bound).
//
// Split.  A masked task-vector element x (fp32) is written as x = h + m + l with h = bf16(x), m = bf16(x - h),
// l = bf16(x - h - m); both residuals are exact in fp32 and 3 x 8 significant bits hold the 24 of x, so the split is
// EXACT.  bf16 x bf16 products are exact in the tensor core's fp32 datapath.  Then, with hh = H^T H etc.,
//     G = hh + (hm + hm^T) + (hl + hl^T) + mm + (ml + ml^T)  [+ ll, 2^-32 relative, dropped].
// Rows of the MMA operand tile are (piece, slice, task): a 128-element tile buffer is cut into two slices of 64
// elements and row r = 42 piece + 21 slice + t holds piece `piece` of task t for the elements of that slice (126 rows,
// tasks >= N stay zero), K-major, canonical no-swizzle layout [group of 8 elements][128 rows][8 elements] (core matrix =
// 8 rows x 16 B; LBO = 2048 B, SBO = 128 B).  ONE tile serves as A and B of both instructions of a step (a step =
// 16 elements of EACH slice; these tiny-N MMAs are bound by the issue rate of the one thread that launches them and by
// their shared-memory operand reads, not by the tensor pipe, hence two slices per instruction and one issuing warp
// per instruction kind):
//     MMA 1: A = rows 0..63 (M = 64), B = rows 0..47 (h_s0, h_s1)        -> D1[h_s,i][h_s,j]              = hh per slice
//     MMA 2: A = rows 0..127, B = rows 42..137 (m_s0, m_s1, l_s0, l_s1)  -> D2[h_s,i | m_s,i][m_s,j | l_s,j] = hm hl / mm ml
// (blocks that pair different slices, and rows past 125, are never read).
//
// Accuracy.  The tensor core adds into its fp32 accumulator with truncation (measured on B200, scratch/tcprobe:
// -6e-8 relative per chained MMA on an all-positive sum) and any single fp32 accumulator collects round-off with
// the square root of its chain length.  Only hh carries full-size sums, so D1 chains just C1 instructions
// (16 C1 elements per slice) and is then drained and summed by CUDA-core warps in round-to-nearest fp32, two-level, in
// a fixed order; D2 holds sums of random-sign products that are 2^-8 .. 2^-16 of hh, so it chains across the tile and
// is drained once.  Across tiles k2_gram_reduce sums in fp64 as before.
//
// Roles (one persistent CTA per SM, 18 warps): warps 0-11 transform in three groups of four -- group k owns MMA tile
// buffer k, so three 128-element buffers are being filled at any time and nobody waits on a neighbour's loads --,
// warps 12-14 drain (TMEM lane quadrants 0-2), warp 15 TMA producer (cp.async.bulk ring: one lane per tensor, one
// for the stage's 64 B of packed mask), warps 16 / 17 issue MMA 1 / MMA 2 (one elected thread each).
// Bound: HBM ((N+1) x 4 B per element) / shared-memory bandwidth (MMA operand reads).
#include "stage_pipe.cuh"
#include "svdq_kernels.h"
#include "tc_common.cuh"

#ifndef SVDQ_DTYPE
#define SVDQ_DTYPE 0
#endif

namespace svdq {

#if SVDQ_DTYPE == 0

constexpr int kWGroups = 3;                    // transform groups (sub-chunk seq is filled by group seq % 3 ...)
constexpr int kWBufs = 4;                      // ... into MMA tile buffer seq % 4: a group never waits for its own MMAs
constexpr int kWGroupThreads = 128;
constexpr int kWTransformWarps = kWGroups * 4;
constexpr int kWThreads = 18 * 32;
constexpr int kWDrain0 = 12, kWProducer = 15, kWMma = 16;       // warps 16, 17: MMA 1, MMA 2
constexpr int kWTB = 128;                      // elements per MMA tile buffer: 2 slices x 4 steps of K = 16
constexpr int kWNP = 21;                       // task rows per (piece, slice) block
constexpr int kWGroupBytes = 2048;             // one group of 8 elements per slice: 128 rows x 16 B
constexpr int kWTileBytes = (kWTB / 16) * kWGroupBytes;    // 16 KB
constexpr int kWSteps = kWTB / 32;             // MMA steps per tile buffer
constexpr int kWN1 = 48, kWN2 = 96;            // N of MMA 1 / MMA 2 (multiples of 16 at M = 128)
constexpr int kWD1Bufs = 4;                    // D1 ring (48 TMEM columns each)
constexpr int kWD2Col = kWD1Bufs * kWN1;       // D2 ring: 2 x 96 columns behind the D1 ring
constexpr int kWTmemCols = 512;
constexpr int kWPad = 256;                     // readable bytes behind the last tile buffer (B of MMA 2 passes row 127)
constexpr int kWStage = 512;                   // elements per raw stage (one 2 KB bulk copy per tensor)
constexpr int kWSubs = kWStage / kWTB;
constexpr int kWRowStride = kWStage * 4 + 16;  // +16 B: the 8 task lanes of a quarter-warp read 8 different bank groups
constexpr int kWMaxTasks = 21;                 // 3 N <= 63 rows
constexpr int kWXLd = 43;                      // leading dimension of the drained D2 rows in shared memory (42 used)
constexpr int kWXRows = 4 * kWNP, kWHLd = 22, kWHRows = 2 * kWNP;
constexpr int kWFixed = kWBufs * kWTileBytes + kWPad + 4096 /*lut*/ + kWXRows * kWXLd * 4 + kWHRows * kWHLd * 4 + 512 /*barriers*/;
static_assert((kWXRows * kWXLd * 4 + kWHRows * kWHLd * 4) % 16 == 0, "barrier block alignment");

// Two fp32 values -> the three bf16 pieces of each, packed (low half = first value).  h = bf16(x) rounded to NEAREST, so
// that the residual x - h (exact, at most 16 significant bits) has a random sign and the cross terms h m stay sums of
// random-sign products (their long accumulator chains would otherwise collect the truncation bias of a same-sign
// sum); m = the top 16 bits of the residual (truncation), and what is left has at most 8 significant bits, i.e. IS a
// bf16 number.  Packed arithmetic (one F2FP / FFMA2 per pair), byte permutes to gather the high halves.
__device__ __forceinline__ void split2(float2 x, uint32_t& h, uint32_t& m, uint32_t& l) {
    const float2 neg1 = make_float2(-1.0f, -1.0f);
    const __nv_bfloat162 hb = __floats2bfloat162_rn(x.x, x.y);
    h = *reinterpret_cast<const uint32_t*>(&hb);
    const float2 hf = make_float2(__uint_as_float(h << 16), __uint_as_float(h & 0xffff0000u));
    const float2 r = __ffma2_rn(hf, neg1, x);                                   // x - h, exact
    const uint32_t ra = __float_as_uint(r.x), rb = __float_as_uint(r.y);
    m = __byte_perm(ra, rb, 0x7632);
    const float2 mf = make_float2(__uint_as_float(ra & 0xffff0000u), __uint_as_float(rb & 0xffff0000u));
    const float2 q = __ffma2_rn(mf, neg1, r);                                   // exact; <= 8 significant bits left
    l = __byte_perm(__float_as_uint(q.x), __float_as_uint(q.y), 0x7632);
}

__host__ __device__ constexpr int k12_stage_bytes(int n_tasks) { return (n_tasks + 1) * kWRowStride + 64; }

template <int C1>
__global__ void __launch_bounds__(kWThreads, 1) k12_gram_wide_tc(const K1Args a, const int n_tasks, const int n_tiles,
                                                                 const int n_stages) {
    constexpr int kChains = kWSteps / C1;                  // D1 chains per tile buffer
    const int N = n_tasks;
    const int G = tri_count(N);
    const int stage_bytes = k12_stage_bytes(N);
    const int mask_off = (N + 1) * kWRowStride;            // the stage's 64 B of packed mask

    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* tile_buf = smem;
    uint4* lut = reinterpret_cast<uint4*>(tile_buf + kWBufs * kWTileBytes + kWPad);
    float* s_x = reinterpret_cast<float*>(lut + 256);      // [84][kWXLd] drained D2 rows: [m_j | l_j] of the row's own slice
    float* s_hh = s_x + kWXRows * kWXLd;                   // [42][kWHLd] per-slice hh rows
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_hh + kWHRows * kWHLd);
    uint64_t* full = bars;                   // [4]        producer -> transform
    uint64_t* empty = full + 4;              // [4]        transform -> producer
    uint64_t* tfull = empty + 4;             // [kWBufs]   transform group -> MMA
    uint64_t* tempty = tfull + kWBufs;       // [kWBufs]   MMA -> transform groups
    uint64_t* d1full = tempty + kWBufs;      // [kWD1Bufs] MMA -> drains 0, 1
    uint64_t* d1empty = d1full + kWD1Bufs;   // [kWD1Bufs]
    uint64_t* d2full = d1empty + kWD1Bufs;   // [2]        MMA -> drains
    uint64_t* d2empty = d2full + 2;          // [2]
    int* s_flags = reinterpret_cast<int*>(d2empty + 2);    // [4] per stage: bit 0 = tensors not staged, bit 1 = mask not staged
    unsigned char* ring = reinterpret_cast<unsigned char*>(bars) + 512;
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < 4; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kWTransformWarps); }
        for (int s = 0; s < kWBufs; ++s) { mbar_init(&tfull[s], 4); mbar_init(&tempty[s], 2); }
        for (int s = 0; s < kWD1Bufs; ++s) { mbar_init(&d1full[s], 1); mbar_init(&d1empty[s], 3); }
        for (int s = 0; s < 2; ++s) { mbar_init(&d2full[s], 1); mbar_init(&d2empty[s], 3); }
        mbar_fence_init();
    }
    if (tid < 256) {        // expansion table: bit c of the byte -> 16-bit lane c all ones
        uint32_t w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
            w[q] = (((uint32_t)tid >> (2 * q)) & 1u ? 0x0000FFFFu : 0u) | (((uint32_t)tid >> (2 * q + 1)) & 1u ? 0xFFFF0000u : 0u);
        lut[tid] = make_uint4(w[0], w[1], w[2], w[3]);
    }
    // rows of tasks >= N are never written: zero the tile buffers (and the pad) once
    for (int i = tid; i < (kWBufs * kWTileBytes + kWPad) / 16; i += kWThreads)
        reinterpret_cast<uint4*>(tile_buf)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (warp == kWMma) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(kWTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    fence_async_smem();                     // the zeros above must be visible to the tensor core's reads
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;

    if (warp == kWProducer) {
        // ================= TMA producer: lane t copies tensor t, lane N + 1 the stage's mask bytes ====================
        uint32_t stage = 0, phase = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int p = a.tile_param[tile];
            const int64_t numel = a.numel[p];
            const int64_t start = (int64_t)a.tile_local[tile] * a.tile_elems;
            const int64_t stop = min(start + (int64_t)a.tile_elems, numel);
            const void* const* tp = a.tensors + (int64_t)p * (N + 1);
            const unsigned char* base = reinterpret_cast<const unsigned char*>(tp[0]);
            const unsigned char* mine = nullptr;
            if (lane <= N) mine = tp[lane] ? reinterpret_cast<const unsigned char*>(tp[lane]) : base;
            // the mask is staged when it is needed (modes 0 / 2), present, and its row starts 16-byte aligned
            const bool mask_staged = a.mask_mode != 1 && a.has_mask_in[p] != 0 && (a.pmask_off[p] & 3) == 0;
            const unsigned char* pk = reinterpret_cast<const unsigned char*>(a.packed_in + a.pmask_off[p]);
            for (int64_t e0 = start; e0 < stop; e0 += kWStage) {
                if (lane == 0) mbar_wait(&empty[stage], phase ^ 1u);
                __syncwarp();
                unsigned char* sb = ring + (size_t)stage * stage_bytes;
                if (e0 + kWStage <= numel) {
                    if (lane == 0) {
                        s_flags[stage] = mask_staged ? 0 : 2;
                        mbar_arrive_expect_tx(&full[stage], (uint32_t)((N + 1) * kWStage * 4 + (mask_staged ? kWStage / 8 : 0)));
                    }
                    __syncwarp();
                    if (mine) bulk_g2s(sb + (size_t)lane * kWRowStride, mine + e0 * 4, kWStage * 4, &full[stage]);
                    if (lane == N + 1 && mask_staged) bulk_g2s(sb + mask_off, pk + (e0 >> 3), kWStage / 8, &full[stage]);
                } else if (lane == 0) {
                    s_flags[stage] = 3;              // tail of the parameter: the transform warps load it themselves
                    mbar_arrive(&full[stage]);
                }
                if (++stage == (uint32_t)n_stages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp >= kWMma) {
        // ================= MMA issuers: warp 16 -> MMA 1 (D1 ring, short chains), warp 17 -> MMA 2 (D2, whole tile) ====
        // The whole warp walks the loops (uniform control flow, every lane polls the barriers); one elected lane issues.
        const bool second = warp == kWMma + 1;
        const bool leader = elect_one();
        PipeState tb, d1, d2;
        const uint32_t idesc = second ? tc_idesc(1u, 1u, 128, kWN2) : tc_idesc(1u, 1u, 64, kWN1);   // MMA 1 needs the h rows (0..41) only
        const uint64_t desc0 = tc_smem_desc(smem_u32(tile_buf), kWGroupBytes, 128);    // buffer 0, step 0, row 0
        const uint64_t b_row = second ? (uint64_t)(2 * kWNP) : 0ull;                  // start address field: 16-byte units
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int p = a.tile_param[tile];
            const int64_t numel = a.numel[p];
            const int64_t start = (int64_t)a.tile_local[tile] * a.tile_elems;
            const int64_t stop = min(start + (int64_t)a.tile_elems, numel);
            const int nsub = (int)((stop - start + kWTB - 1) / kWTB);
            if (second) {
                mbar_wait(&d2empty[d2.stage], d2.phase ^ 1u);
                tc_fence_after();
            }
            const uint32_t t2 = tmem + kWD2Col + d2.stage * kWN2;
            for (int sb = 0; sb < nsub; ++sb) {
                mbar_wait(&tfull[tb.stage], tb.phase);
                tc_fence_after();
                const uint64_t da = desc0 + (uint64_t)(tb.stage * (kWTileBytes >> 4));
                if (!second) {
#pragma unroll
                    for (int c = 0; c < kChains; ++c) {
                        mbar_wait(&d1empty[d1.stage], d1.phase ^ 1u);
                        tc_fence_after();
                        if (leader) {
                            const uint32_t t1 = tmem + d1.stage * kWN1;
#pragma unroll
                            for (int k = 0; k < C1; ++k) {
                                const uint64_t d = da + (uint64_t)((c * C1 + k) * (2 * kWGroupBytes >> 4));
                                tc_mma_f16(t1, d, d, idesc, k ? 1u : 0u);
                            }
                            tc_commit(&d1full[d1.stage]);
                        }
                        __syncwarp();
                        d1.advance<kWD1Bufs>();
                    }
                } else if (leader) {
#pragma unroll
                    for (int k = 0; k < kWSteps; ++k) {
                        const uint64_t d = da + (uint64_t)(k * (2 * kWGroupBytes >> 4));
                        tc_mma_f16(t2, d, d + b_row, idesc, (sb | k) ? 1u : 0u);
                    }
                }
                if (leader) tc_commit(&tempty[tb.stage]);       // this warp's MMAs have read the tile buffer
                __syncwarp();
                tb.advance<kWBufs>();
            }
            if (second) {
                if (leader) tc_commit(&d2full[d2.stage]);
                __syncwarp();
                d2.advance<2>();
            }
        }
    } else if (warp >= kWDrain0 && warp < kWDrain0 + 3) {
        // ================= drain warps: TMEM -> registers -> per-tile Gram ============================================
        // M = 128 accumulators: row r of D sits in TMEM lane r; warp q reads lanes 32 q .. 32 q + 31.
        // D1 rows: h_s,i = 21 s + i (42 rows: quadrants 0, 1).  D2 rows: h_s,i = 21 s + i, m_s,i = 42 + 21 s + i (84 rows);
        // D2 columns: m_s,j = 21 s + j, l_s,j = 42 + 21 s + j.
        const int q = warp - kWDrain0;                      // == warp % 4: the TMEM lane quadrant this warp may read
        const int row = 32 * q + lane;                      // D2 (M = 128): row r in TMEM lane r
        const bool slice1 = (row / kWNP) & 1;               // rows of slice 1 pair with the columns of slice 1
        const int row1 = 16 * q + (lane & 15);              // D1 (M = 64): row r in TMEM lane 32 (r / 16) + r % 16
        const bool row1_ok = lane < 16 && row1 < kWHRows;
        const bool slice1_d1 = row1 >= kWNP;
        const uint32_t lane_base = (uint32_t)(32 * q) << 16;
        PipeState d1, d2;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int p = a.tile_param[tile];
            const int64_t numel = a.numel[p];
            const int64_t start = (int64_t)a.tile_local[tile] * a.tile_elems;
            const int64_t stop = min(start + (int64_t)a.tile_elems, numel);
            const int nsub = (int)((stop - start + kWTB - 1) / kWTB);
            {                                               // hh rows 0..41 live in lanes 0..15 of quadrants 0, 1, 2
                float hi[kWNP], lo[kWNP];
#pragma unroll
                for (int j = 0; j < kWNP; ++j) { hi[j] = 0.0f; lo[j] = 0.0f; }
                const int n_chains = nsub * kChains;
                for (int c = 0; c < n_chains; ++c) {
                    mbar_wait(&d1full[d1.stage], d1.phase);
                    tc_fence_after();
                    uint32_t r[48];
                    {
                        uint32_t r0[16], r1[16], r2[16];
                        const uint32_t t1 = tmem + lane_base + d1.stage * kWN1;
                        tc_ld16_nowait(t1, r0); tc_ld16_nowait(t1 + 16, r1); tc_ld16_nowait(t1 + 32, r2);
                        tc_wait_ld();
#pragma unroll
                        for (int j = 0; j < 16; ++j) { r[j] = r0[j]; r[16 + j] = r1[j]; r[32 + j] = r2[j]; }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&d1empty[d1.stage]);
                    d1.advance<kWD1Bufs>();
#pragma unroll
                    for (int j = 0; j < kWNP; ++j) lo[j] += __uint_as_float(slice1_d1 ? r[kWNP + j] : r[j]);
                    if ((c & 15) == 15) {                   // two-level sum: 16 chains per low accumulator
#pragma unroll
                        for (int j = 0; j < kWNP; ++j) { hi[j] += lo[j]; lo[j] = 0.0f; }
                    }
                }
                if (row1_ok) {
#pragma unroll
                    for (int j = 0; j < kWNP; ++j) s_hh[row1 * kWHLd + j] = hi[j] + lo[j];
                }
            }
            // ---- tile end: D2 (whole-tile sums of the cross terms) -----------------------------------------------
            mbar_wait(&d2full[d2.stage], d2.phase);
            tc_fence_after();
            {
                const uint32_t t2 = tmem + lane_base + kWD2Col + d2.stage * kWN2;
#pragma unroll
                for (int s = 0; s < 6; ++s) {
                    uint32_t r16[16];
                    tc_ld16_nowait(t2 + s * 16, r16);
                    tc_wait_ld();
                    if (row < kWXRows) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int c = s * 16 + j;                    // static: column -> (block, slice, task)
                            if (c < 4 * kWNP) {
                                const int blk = c / (2 * kWNP), cs = (c / kWNP) & 1, tj = c % kWNP;
                                if (slice1 == (cs != 0)) s_x[row * kWXLd + blk * kWNP + tj] = __uint_as_float(r16[j]);
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&d2empty[d2.stage]);
            d2.advance<2>();
            named_bar_sync(2, 96);
            // G[i][j] = sum over slices of hh + mm + (Y + Y^T), Y = hm + hl + ml
            {
                float* gout = a.gram + (int64_t)tile * G;
                for (int idx = (warp - kWDrain0) * 32 + lane; idx < G; idx += 96) {
                    int i = 0, rem = idx;                    // unpack the upper triangle (row-major, as tri_index)
                    while (rem >= N - i) { rem -= N - i; ++i; }
                    const int j = i + rem;
                    float g2[2];
#pragma unroll
                    for (int sl = 0; sl < 2; ++sl) {
                        const float* xh_i = s_x + (kWNP * sl + i) * kWXLd;            // row h_s,i: [hm | hl]
                        const float* xh_j = s_x + (kWNP * sl + j) * kWXLd;
                        const float* xm_i = s_x + (2 * kWNP + kWNP * sl + i) * kWXLd;  // row m_s,i: [mm | ml]
                        const float* xm_j = s_x + (2 * kWNP + kWNP * sl + j) * kWXLd;
                        const float y_ij = (xh_i[j] + xh_i[kWNP + j]) + xm_i[kWNP + j];
                        const float y_ji = (xh_j[i] + xh_j[kWNP + i]) + xm_j[kWNP + i];
                        g2[sl] = ((y_ij + y_ji) + xm_i[j]) + s_hh[(kWNP * sl + i) * kWHLd + j];
                    }
                    gout[idx] = g2[0] + g2[1];
                }
            }
            named_bar_sync(2, 96);
        }
    } else if (warp < kWTransformWarps) {
        // ================= transform groups ==========================================================================
        const int grp = warp >> 2;
        const int tg = tid & (kWGroupThreads - 1);
        const int g = tg >> 3, tl = tg & 7;                 // group of 8 elements inside the buffer, task lane
        // element group g of the buffer = slice g / 8, K group g % 8; row = 42 piece + 21 slice + task
        unsigned char* tile_out0 = tile_buf + (g & 7) * kWGroupBytes + (g >> 3) * (kWNP * 16);
        uint32_t stage = 0, phase = 0;                      // raw ring position
        int seq = 0;                                        // running index of the 128-element sub-chunks of this CTA
        const int mode = a.mask_mode;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int p = a.tile_param[tile];
            const int64_t numel = a.numel[p];
            const int64_t start = (int64_t)a.tile_local[tile] * a.tile_elems;
            const int64_t stop = min(start + (int64_t)a.tile_elems, numel);
            const void* const* tp = a.tensors + (int64_t)p * (N + 1);
            const bool has_mask = a.has_mask_in[p] != 0;
            const uint8_t* pk = reinterpret_cast<const uint8_t*>(a.packed_in + a.pmask_off[p]);
            for (int64_t s0 = start; s0 < stop; s0 += kWStage) {
                mbar_wait(&full[stage], phase);
                const unsigned char* sbase = ring + (size_t)stage * stage_bytes;
                const int flags = s_flags[stage];
#pragma unroll 1
                for (int sub = 0; sub < kWSubs; ++sub, ++seq) {
                    const int64_t e0 = s0 + (int64_t)sub * kWTB;
                    if (e0 >= stop) break;                             // seq counts the sub-chunks that exist, like the MMA warp
                    if (seq % kWGroups != grp) continue;
                    const int64_t e = e0 + 8 * g;
                    const int64_t left = numel - e;
                    uint32_t mbyte = left >= 8 ? 0xFFu : (left <= 0 ? 0u : ((1u << (int)left) - 1u));
                    if (mode != 1 && mbyte != 0u) {
                        uint32_t w = 0xFFu;
                        if (has_mask) w = (flags & 2) ? (uint32_t)__ldg(pk + (e >> 3)) : (uint32_t)sbase[mask_off + sub * (kWTB / 8) + g];
                        mbyte &= mode == 2 ? (has_mask ? ~w & 0xFFu : 0u) : w;
                    }
                    // all loads of the sub-chunk first (base once, then this thread's tasks tl, tl + 8, tl + 16; a lane
                    // without a third task re-reads a valid row and does not store it)
                    const float2 neg1 = make_float2(-1.0f, -1.0f);
                    float2 b[4], f[3][4];
                    const bool staged = !(flags & 1);
                    if (staged) {
                        const unsigned char* src = sbase + (sub * kWTB + 8 * g) * 4;
                        const float4 b0 = *reinterpret_cast<const float4*>(src);
                        const float4 b1 = *reinterpret_cast<const float4*>(src + 16);
                        b[0] = make_float2(b0.x, b0.y); b[1] = make_float2(b0.z, b0.w);
                        b[2] = make_float2(b1.x, b1.y); b[3] = make_float2(b1.z, b1.w);
#pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            const int t = min(tl + 8 * k, N - 1);
                            const float4 f0 = *reinterpret_cast<const float4*>(src + (size_t)(t + 1) * kWRowStride);
                            const float4 f1 = *reinterpret_cast<const float4*>(src + (size_t)(t + 1) * kWRowStride + 16);
                            f[k][0] = make_float2(f0.x, f0.y); f[k][1] = make_float2(f0.z, f0.w);
                            f[k][2] = make_float2(f1.x, f1.y); f[k][3] = make_float2(f1.z, f1.w);
                        }
                    } else {
                        float bb[8];
#pragma unroll
                        for (int c = 0; c < 8; ++c) bb[c] = e + c < numel ? Elem<float>::load1(tp[0], e + c) : 0.0f;
#pragma unroll
                        for (int c = 0; c < 4; ++c) b[c] = make_float2(bb[2 * c], bb[2 * c + 1]);
#pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            const int t = min(tl + 8 * k, N - 1);
                            const void* ft = tp[t + 1] ? tp[t + 1] : tp[0];
                            float ff[8];
#pragma unroll
                            for (int c = 0; c < 8; ++c) ff[c] = e + c < numel ? Elem<float>::load1(ft, e + c) : 0.0f;
#pragma unroll
                            for (int c = 0; c < 4; ++c) f[k][c] = make_float2(ff[2 * c], ff[2 * c + 1]);
                        }
                    }
                    const bool partial = __any_sync(0xffffffffu, mbyte != 0xFFu);   // warp-uniform: no divergence below
                    uint4 km = make_uint4(~0u, ~0u, ~0u, ~0u);
                    if (partial) km = lut[mbyte];
                    const int buf = seq & (kWBufs - 1);
                    unsigned char* tile_out = tile_out0 + buf * kWTileBytes;
                    mbar_wait(&tempty[buf], (((uint32_t)seq >> 2) & 1u) ^ 1u);      // the MMAs that read this buffer last time have finished
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const int t = tl + 8 * k;
                        uint4 H, M, L;
                        // finetuned - base (task_vector_loader.py:142): b * -1 + f is the correctly rounded difference
                        split2(__ffma2_rn(b[0], neg1, f[k][0]), H.x, M.x, L.x);
                        split2(__ffma2_rn(b[1], neg1, f[k][1]), H.y, M.y, L.y);
                        split2(__ffma2_rn(b[2], neg1, f[k][2]), H.z, M.z, L.z);
                        split2(__ffma2_rn(b[3], neg1, f[k][3]), H.w, M.w, L.w);
                        if (partial) {
                            H.x &= km.x; H.y &= km.y; H.z &= km.z; H.w &= km.w;
                            M.x &= km.x; M.y &= km.y; M.z &= km.z; M.w &= km.w;
                            L.x &= km.x; L.y &= km.y; L.z &= km.z; L.w &= km.w;
                        }
                        if (t < N) {
                            *reinterpret_cast<uint4*>(tile_out + t * 16) = H;
                            *reinterpret_cast<uint4*>(tile_out + (2 * kWNP + t) * 16) = M;
                            *reinterpret_cast<uint4*>(tile_out + (4 * kWNP + t) * 16) = L;
                        }
                    }
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tfull[buf]);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[stage]);
                if (++stage == (uint32_t)n_stages) { stage = 0; phase ^= 1u; }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kWMma) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kWTmemCols));
}

template <int C1>
static cudaError_t k12_go(const K1Args& a, int n_tasks, int n_tiles, int n_sm, cudaStream_t st) {
    if (a.tile_elems % kWStage != 0) return cudaErrorNotSupported;
    const size_t stage = (size_t)k12_stage_bytes(n_tasks);
    const size_t budget = 232448 - 1024;                 // 227 KB per CTA minus the static shared memory
    int stages = (int)((budget - kWFixed) / stage);
    if (stages > 4) stages = 4;
    if (stages < 2) return cudaErrorNotSupported;
    const size_t dsm = kWFixed + stages * stage;
    cudaError_t e = cudaFuncSetAttribute(k12_gram_wide_tc<C1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm);
    if (e != cudaSuccess) return e;
    const int grid = n_tiles < n_sm ? n_tiles : n_sm;
    k12_gram_wide_tc<C1><<<grid, kWThreads, dsm, st>>>(a, n_tasks, n_tiles, stages);
    return cudaGetLastError();
}

#endif  // fp32

// tensor-core wide Gram exists for fp32 inputs in pre-combined mask mode, 2..21 tasks; cudaErrorNotSupported otherwise.
// chain = 16-element MMA steps per D1 accumulator chain (2, 4 or 8; fewer = less truncation bias, more drain work)
template <>
cudaError_t k12_launch_dtype<SVDQ_DTYPE>(int n_tasks, const K1Args& a, int n_tiles, int n_sm, int chain, cudaStream_t st) {
#if SVDQ_DTYPE == 0
    if (n_tiles <= 0) return cudaSuccess;
    if (n_tasks < 2 || n_tasks > kWMaxTasks || a.packed_in == nullptr) return cudaErrorNotSupported;
    switch (chain) {
        case 2: return k12_go<2>(a, n_tasks, n_tiles, n_sm, st);
        case 8: return k12_go<8>(a, n_tasks, n_tiles, n_sm, st);
        default: return k12_go<4>(a, n_tasks, n_tiles, n_sm, st);
    }
#else
    return cudaErrorNotSupported;
#endif
}

}  // namespace svdq
