// K13 — pass 2 of the wide path (basis rows, weighted reconstruction, mask scatter, merged = base + delta) for
// 9..21 FP32 task vectors on the 5th-generation tensor cores.  Same inputs, outputs and reference lines as
// k6_reconstruct_merge (reconstruct_from_coefficients src/svd_hybrid/merge.py:144-194, merge_parameter :197-312,
// apply_merged_deltas :429-552, the fp16 cast of the bases src/svd_hybrid/cli.py:355-361,
// reconstruct_from_masked src/svd_hybrid/mask_loader.py:750-763); no fused diagnostics, no noise region.
//
// Why.  With fp16 bases every basis entry u_dj = sum_t (tau_dt - mean_d) W_tj has to be formed and rounded before
// the contraction with cbar: N r = 400 FMAs per element at N = 20, and the CUDA-core kernel sits at 47 % of the HBM
// roofline (profiles/r1_ncu_full_wide_k6_k8.csv: issue-bound).
//
// Operands.  tau (fp32) is split EXACTLY into three bf16 pieces per element (as in k12_gram_wide_tc.cu) and written
// into the tile [group of 8 elements][64 rows][8 elements], row = 21 piece + task: read as an MN-major A operand
// this is A[element][(piece, task)] (M = 128 elements, K = 64).  W' = W - 1 (1^T W) / n (the centring folded in, formed
// in fp64) is split into three bf16 pieces too, and B[(piece, task)][(q, j)] = W'_q[task][j] for every tau piece, so
//     D[d][(q, j)] = sum_{piece, t} tau_piece[d][t] W'_q[t][j] = (tau_d W'_q)_j,     u_dj = D[d][(0,j)] + D[d][(1,j)] + D[d][(2,j)]
// (column (q, j) sits at 24 (j / 8) + 8 q + j % 8, so that a block of 8 basis columns is 24 adjacent accumulator columns),
// plus one column of ones that yields sum_t tau_dt for the mean.  N = 80 columns, four MMAs (K = 16 each) per 128
// elements, accumulators in TMEM; the products are exact, the fp32 accumulation chains 4 instructions.
// The epilogue warps read their element's row from TMEM, add the three W pieces, round to fp16 when the bases are
// stored in fp16, contract with cbar, add the mean, apply the packed tall mask and stream out merged = base + delta.
//
// Roles (one persistent CTA per SM, 22 warps): warps 0-11 transform in three groups of four (sub-chunk seq is filled
// by group seq % 3 into tile buffer seq % 4; at the start of a tile all twelve build the B operand of its parameter),
// warp 12 TMA producer, warp 13 MMA issuer, warps 14-21 epilogue (two sets of four TMEM lane quadrants, alternating
// sub-chunks).  Bound: HBM ((N+1) x 4 + 4 B per element) / instruction issue.
#include "stage_pipe.cuh"
#include "svdq_kernels.h"
#include "tc_common.cuh"

#ifndef SVDQ_DTYPE
#define SVDQ_DTYPE 0
#endif

namespace svdq {

#if SVDQ_DTYPE == 0

constexpr int kXGroups = 3;
constexpr int kXGroupThreads = 128;
constexpr int kXTransformWarps = kXGroups * 4;
constexpr int kXTransform = kXTransformWarps * 32;
constexpr int kXProducer = 12, kXMma = 13, kXEpi0 = 14, kXEpiWarps = 8;
constexpr int kXThreads = (kXEpi0 + kXEpiWarps) * 32;
constexpr int kXTB = 128;                      // elements per tile buffer = rows of one accumulator
constexpr int kXNP = 21;                       // task rows per piece (K = 3 x 21 + 1 zero row)
constexpr int kXJ = 24;                        // basis columns per W piece in the accumulator
constexpr int kXN = 80;                        // accumulator columns: 3 x 24 + the ones column (72) + padding
constexpr int kXBufs = 4;                      // tile buffers
constexpr int kXTileBytes = (kXTB / 8) * 1024; // [16 groups][64 rows][16 B]
constexpr int kXBaseRing = 8;                  // base values of the sub-chunks in flight (released by the epilogue)
constexpr int kXBaseBytes = kXTB * 4;
constexpr int kXBSlots = 3;                    // B operands (one per tile in flight)
constexpr int kXBBytes = (kXN / 8) * 1024;     // [10 column groups][8 K groups][8 columns][16 B]
constexpr int kXAccBufs = 4;                   // accumulators in TMEM (80 columns each)
constexpr int kXTmemCols = 512;
constexpr int kXStage = kXGroups * kXTB;        // elements per raw stage: one sub-chunk per transform group (1.5 KB per tensor)
constexpr int kXSubs = kXStage / kXTB;
constexpr int kXRowStride = kXStage * 4 + 16;
constexpr int kXMaxTasks = 21;
constexpr int kXTabBytes = 3 * kXNP * kXJ * 2; // W' pieces [q][t][j] bf16
constexpr int kXFixed = kXBufs * kXTileBytes + kXBaseRing * kXBaseBytes + kXBSlots * kXBBytes + kXTabBytes + 512;
static_assert(kXTabBytes % 16 == 0, "alignment of the barrier block");

__host__ __device__ constexpr int k13_stage_bytes(int n_tasks) { return (n_tasks + 1) * kXRowStride; }

__device__ __forceinline__ void k13_split2(float2 x, uint32_t& h, uint32_t& m, uint32_t& l) {
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

__device__ __forceinline__ void tc_ld8_nowait(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}

template <bool FP16B>
__global__ void __launch_bounds__(kXThreads, 1) k13_merge_wide_tc(const K3Args a, const int n_tasks, const int n_tiles,
                                                                  const int n_stages) {
    const int N = n_tasks;
    const int stage_bytes = k13_stage_bytes(N);

    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* tile_buf = smem;
    unsigned char* base_ring = tile_buf + kXBufs * kXTileBytes;
    unsigned char* b_ring = base_ring + kXBaseRing * kXBaseBytes;
    uint16_t* s_tab = reinterpret_cast<uint16_t*>(b_ring + kXBSlots * kXBBytes);          // [3][21][24]
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(s_tab) + kXTabBytes);
    uint64_t* full = bars;                   // [8]           producer -> transform
    uint64_t* empty = full + 8;              // [8]           transform -> producer
    uint64_t* tfull = empty + 8;             // [kXBufs]      transform group -> MMA
    uint64_t* tempty = tfull + kXBufs;       // [kXBufs]      MMA -> transform groups
    uint64_t* afull = tempty + kXBufs;       // [kXAccBufs]   MMA -> epilogue set
    uint64_t* aempty = afull + kXAccBufs;    // [kXAccBufs]   epilogue set -> MMA
    uint64_t* bempty = aempty + kXAccBufs;   // [kXBaseRing]  epilogue set (base values read) -> transform group
    uint64_t* wfull = bempty + kXBaseRing;   // [kXBSlots]    transform (B operand written) -> MMA
    uint64_t* wempty = wfull + kXBSlots;     // [kXBSlots]    MMA (tile finished) -> transform
    int* s_direct = reinterpret_cast<int*>(wempty + kXBSlots);                            // [8]
    unsigned char* ring = reinterpret_cast<unsigned char*>(bars) + 512;
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < 8; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kXTransformWarps); }
        for (int s = 0; s < kXBufs; ++s) { mbar_init(&tfull[s], 4); mbar_init(&tempty[s], 1); }
        for (int s = 0; s < kXAccBufs; ++s) { mbar_init(&afull[s], 1); mbar_init(&aempty[s], 4); }
        for (int s = 0; s < kXBaseRing; ++s) mbar_init(&bempty[s], 4);
        for (int s = 0; s < kXBSlots; ++s) { mbar_init(&wfull[s], kXTransformWarps); mbar_init(&wempty[s], 1); }
        mbar_fence_init();
    }
    // row 63 of every group and the rows of tasks >= N are never written: zero the tile buffers once
    for (int i = tid; i < kXBufs * kXTileBytes / 16; i += kXThreads)
        reinterpret_cast<uint4*>(tile_buf)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (warp == kXMma) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(kXTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;

    if (warp == kXProducer) {
        // ================= TMA producer: lane t copies tensor t ========================================================
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
            for (int64_t e0 = start; e0 < stop; e0 += kXStage) {
                if (lane == 0) mbar_wait(&empty[stage], phase ^ 1u);
                __syncwarp();
                unsigned char* sb = ring + (size_t)stage * stage_bytes;
                if (e0 + kXStage <= numel) {
                    if (lane == 0) {
                        s_direct[stage] = 0;
                        mbar_arrive_expect_tx(&full[stage], (uint32_t)((N + 1) * kXStage * 4));
                    }
                    __syncwarp();
                    if (mine) bulk_g2s(sb + (size_t)lane * kXRowStride, mine + e0 * 4, kXStage * 4, &full[stage]);
                } else if (lane == 0) {
                    s_direct[stage] = 1;             // tail of the parameter: the transform warps load it themselves
                    mbar_arrive(&full[stage]);
                }
                if (++stage == (uint32_t)n_stages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == kXMma) {
        // ================= MMA issuer: the whole warp walks the loops, one elected lane issues ==========================
        const bool leader = elect_one();
        PipeState tb, ab, wb;
        const uint32_t idesc = tc_idesc(1u, 1u, 128, kXN, /*a_mn=*/1u, /*b_mn=*/0u);
        // A, MN-major: groups of 8 elements (M) 1024 B apart, groups of 8 rows (K) 128 B apart
        const uint64_t desc_a0 = tc_smem_desc(smem_u32(tile_buf), /*lbo (K groups)=*/128, /*sbo (M groups)=*/1024);
        // B, K-major: groups of 8 K 128 B apart, groups of 8 columns 1024 B apart
        const uint64_t desc_b0 = tc_smem_desc(smem_u32(b_ring), /*lbo (K groups)=*/128, /*sbo (N groups)=*/1024);
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int p = a.tile_param[tile];
            const int64_t numel = a.numel[p];
            const int64_t start = (int64_t)a.tile_local[tile] * a.tile_elems;
            const int64_t stop = min(start + (int64_t)a.tile_elems, numel);
            const int nsub = (int)((stop - start + kXTB - 1) / kXTB);
            mbar_wait(&wfull[wb.stage], wb.phase);
            tc_fence_after();
            const uint64_t db = desc_b0 + (uint64_t)(wb.stage * (kXBBytes >> 4));
            for (int sb = 0; sb < nsub; ++sb) {
                mbar_wait(&tfull[tb.stage], tb.phase);
                mbar_wait(&aempty[ab.stage], ab.phase ^ 1u);
                tc_fence_after();
                if (leader) {
                    const uint64_t da = desc_a0 + (uint64_t)(tb.stage * (kXTileBytes >> 4));
                    const uint32_t td = tmem + ab.stage * kXN;
#pragma unroll
                    for (int k = 0; k < 4; ++k)         // K = 16 rows per instruction: two K groups = 256 B in A and in B
                        tc_mma_f16(td, da + (uint64_t)(k * 16), db + (uint64_t)(k * 16), idesc, k ? 1u : 0u);
                    tc_commit(&tempty[tb.stage]);
                    tc_commit(&afull[ab.stage]);
                }
                __syncwarp();
                tb.advance<kXBufs>();
                ab.advance<kXAccBufs>();
            }
            if (leader) tc_commit(&wempty[wb.stage]);   // the tile's MMAs have read its B operand
            __syncwarp();
            wb.advance<kXBSlots>();
        }
    } else if (warp >= kXEpi0) {
        // ================= epilogue warps ===========================================================================
        const int q = warp & 3;                         // TMEM lane quadrant (warp % 4)
        const int eset = (warp - kXEpi0) >> 2;          // sub-chunks seq = eset, eset + 2, ... belong to this set
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        int seq = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int p = a.tile_param[tile];
            const int64_t numel = a.numel[p];
            const int64_t start = (int64_t)a.tile_local[tile] * a.tile_elems;
            const int64_t stop = min(start + (int64_t)a.tile_elems, numel);
            const bool solved = a.info[(int64_t)p * 8 + 0] == kSolved;
            const int n_active = a.info[(int64_t)p * 8 + 1];
            const int r = solved ? a.info[(int64_t)p * 8 + 4] : 0;
            const float tail_add = a.scal[(int64_t)p * 4 + 1];
            const float mean_scale = a.scal[(int64_t)p * 4 + 2];
            const bool has_mask = a.has_mask[p] != 0;
            const uint32_t* packed = has_mask ? a.packed + a.pmask_off[p] : nullptr;
            float* outp = a.out[p];
            const float n_f = (float)(n_active > 0 ? n_active : 1);
            const bool pow2 = (n_active & (n_active - 1)) == 0;
            const float inv_n = __fdiv_rn(1.0f, n_f);
            const bool center = a.center != 0;
            float cb[kXJ];
#pragma unroll
            for (int j = 0; j < kXJ; ++j) cb[j] = (j < N && j < r) ? a.cbar[(int64_t)p * N + j] : 0.0f;
            // tile-local 32-bit indexing; this set owns the sub-chunks whose running number has its parity
            const int n_tile = (int)(stop - start);
            const int nsub = (n_tile + kXTB - 1) / kXTB;
            float* out_t = outp + start + q * 32 + lane;
            const uint32_t* pk_t = has_mask ? packed + (start >> 5) + q : nullptr;
            for (int sb = (eset ^ seq) & 1; sb < nsub; sb += 2) {
                const int my_seq = seq + sb;
                const int ab = my_seq & (kXAccBufs - 1), bslot = my_seq & (kXBaseRing - 1);
                const int o = sb * kXTB + q * 32;               // first element of this warp's row block inside the tile
                const uint32_t pw = (has_mask && solved && o < n_tile) ? __ldg(pk_t + sb * (kXTB / 32)) : 0xFFFFFFFFu;
                mbar_wait(&afull[ab], ((uint32_t)my_seq / kXAccBufs) & 1u);
                tc_fence_after();
                const float base_v = reinterpret_cast<const float*>(base_ring + bslot * kXBaseBytes)[q * 32 + lane];
                const uint32_t trow = tmem + lane_base + (uint32_t)(ab * kXN);
                float acc = 0.0f;
                uint32_t s_bits;
                tc_ld1_nowait(trow + 3 * kXJ, s_bits);
#pragma unroll
                for (int c = 0; c < kXJ / 8; ++c) {
                    if (8 * c < r) {                            // warp-uniform
                        // columns of chunk c: [W'_0 x 8 | W'_1 x 8 | W'_2 x 8] at 24 c
                        uint32_t d01[16], d2[8];
                        tc_ld16_nowait(trow + 24 * c, d01);
                        tc_ld8_nowait(trow + 24 * c + 16, d2);
                        tc_wait_ld();
                        float2 part = make_float2(0.0f, 0.0f);
#pragma unroll
                        for (int j = 0; j < 8; j += 2) {
                            float2 u = __fadd2_rn(make_float2(__uint_as_float(d01[8 + j]), __uint_as_float(d01[9 + j])),
                                                  make_float2(__uint_as_float(d2[j]), __uint_as_float(d2[j + 1])));
                            u = __fadd2_rn(make_float2(__uint_as_float(d01[j]), __uint_as_float(d01[j + 1])), u);
                            if (FP16B) u = __half22float2(__float22half2_rn(u));     // the stored basis is fp16 (cli.py:355-361)
                            part = __ffma2_rn(u, make_float2(cb[8 * c + j], cb[8 * c + j + 1]), part);
                        }
                        acc += part.x + part.y;
                    }
                }
                tc_wait_ld();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { mbar_arrive(&aempty[ab]); mbar_arrive(&bempty[bslot]); }
                float res = base_v;
                if (solved) {
                    const float ms = __uint_as_float(s_bits);
                    const float mean = center ? (pow2 ? ms * inv_n : __fdiv_rn(ms, n_f)) : 0.0f;
                    const float val = fmaf(mean, mean_scale, acc) + tail_add;
                    res = base_v + (((pw >> lane) & 1u) ? val : 0.0f);
                }
                if (o + lane < n_tile) __stcs(out_t + sb * kXTB, res);
            }
            seq += nsub;
        }
    } else {
        // ================= transform groups ==========================================================================
        const int grp = warp >> 2;
        const int tg = tid & (kXGroupThreads - 1);
        const int g = tg >> 3, tl = tg & 7;                 // group of 8 elements inside the buffer, task lane
        unsigned char* tile_out0 = tile_buf + g * 1024;
        uint32_t stage = 0, phase = 0;
        int seq = 0, tcount = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcount) {
            const int p = a.tile_param[tile];
            const int64_t numel = a.numel[p];
            const int64_t start = (int64_t)a.tile_local[tile] * a.tile_elems;
            const int64_t stop = min(start + (int64_t)a.tile_elems, numel);
            const void* const* tp = a.tensors + (int64_t)p * (N + 1);
            // ---- B operand of this parameter (all twelve warps) ----------------------------------------------------
            {
                const int slot = tcount % kXBSlots;
                const bool solved = a.info[(int64_t)p * 8 + 0] == kSolved;
                const int r = solved ? a.info[(int64_t)p * 8 + 4] : 0;
                const int n_active = a.info[(int64_t)p * 8 + 1];
                const float* W = a.W + (int64_t)p * N * N;
                named_bar_sync(1, kXTransform);             // the previous tile's table has been consumed
                // W' = W - 1 (1^T W) / n in fp64 (rows of the tasks that have the parameter), three bf16 pieces
                for (int i = tid; i < kXNP * kXJ; i += kXTransform) {
                    const int t = i / kXJ, j = i - t * kXJ;
                    uint32_t w1 = 0u, w2 = 0u, w3 = 0u;
                    if (t < N && j < N && j < r) {
                        double w = (double)W[t * N + j];
                        if (a.center && tp[t + 1] != nullptr) {
                            double s = 0.0;
                            for (int u = 0; u < N; ++u) s += (double)W[u * N + j];
                            w -= s / (double)(n_active > 0 ? n_active : 1);
                        }
                        w1 = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn((float)w));
                        const double r1 = w - (double)__uint_as_float(w1 << 16);
                        w2 = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn((float)r1));
                        const double r2 = r1 - (double)__uint_as_float(w2 << 16);
                        w3 = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn((float)r2));
                    }
                    s_tab[i] = (uint16_t)w1; s_tab[kXNP * kXJ + i] = (uint16_t)w2; s_tab[2 * kXNP * kXJ + i] = (uint16_t)w3;
                }
                named_bar_sync(1, kXTransform);
                mbar_wait(&wempty[slot], (((uint32_t)tcount / kXBSlots) & 1u) ^ 1u);     // the MMAs of the tile that used this slot are done
                // image [column group n / 8][K group k / 8][column n % 8][k % 8]: one 16-byte vector = 8 K of one column
                unsigned char* bimg = b_ring + slot * kXBBytes;
                for (int v = tid; v < (kXN / 8) * 8 * 8; v += kXTransform) {
                    const int ng = v >> 6, kg = (v >> 3) & 7, nn = v & 7;
                    const int n = ng * 8 + nn;
                    uint32_t w[8];
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) {
                        const int k = kg * 8 + kk;                    // row (piece, task) of A
                        const int t = k % kXNP;
                        uint32_t val = 0u;
                        if (k < 3 * kXNP && t < N) {
                            // column n = 24 c + 8 q + jj holds piece q of basis column j = 8 c + jj
                            if (n < 3 * kXJ) val = s_tab[((n % 24) >> 3) * kXNP * kXJ + t * kXJ + (n / 24) * 8 + (n & 7)];
                            else if (n == 3 * kXJ) val = 0x3F80u;     // 1.0: column 72 = sum over pieces and tasks = sum_t tau_t
                        }
                        w[kk] = val;
                    }
                    *reinterpret_cast<uint4*>(bimg + v * 16) =
                        make_uint4(w[0] | (w[1] << 16), w[2] | (w[3] << 16), w[4] | (w[5] << 16), w[6] | (w[7] << 16));
                }
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&wfull[slot]);
            }
            for (int64_t s0 = start; s0 < stop; s0 += kXStage) {
                mbar_wait(&full[stage], phase);
                const unsigned char* sbase = ring + (size_t)stage * stage_bytes;
                const bool direct = s_direct[stage] != 0;
                // sub-chunk `grp` of the stage belongs to this group; seq numbers the sub-chunks that exist (the last
                // stage of a tile may hold fewer than three), exactly as the MMA and epilogue warps count them
                const int n_here = (int)min((int64_t)kXSubs, (stop - s0 + kXTB - 1) / kXTB);
                if (grp < n_here) {
                    const int sub = grp;
                    const int my_seq = seq + sub;
                    const int64_t e0 = s0 + (int64_t)sub * kXTB;
                    const int64_t e = e0 + 8 * g;
                    const float2 neg1 = make_float2(-1.0f, -1.0f);
                    float2 b[4], f[3][4];
                    if (!direct) {
                        const unsigned char* src = sbase + (sub * kXTB + 8 * g) * 4;
                        const float4 b0 = *reinterpret_cast<const float4*>(src);
                        const float4 b1 = *reinterpret_cast<const float4*>(src + 16);
                        b[0] = make_float2(b0.x, b0.y); b[1] = make_float2(b0.z, b0.w);
                        b[2] = make_float2(b1.x, b1.y); b[3] = make_float2(b1.z, b1.w);
#pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            const int t = min(tl + 8 * k, N - 1);
                            const float4 f0 = *reinterpret_cast<const float4*>(src + (size_t)(t + 1) * kXRowStride);
                            const float4 f1 = *reinterpret_cast<const float4*>(src + (size_t)(t + 1) * kXRowStride + 16);
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
                    const int buf = my_seq & (kXBufs - 1), bslot = my_seq & (kXBaseRing - 1);
                    unsigned char* tile_out = tile_out0 + buf * kXTileBytes;
                    mbar_wait(&tempty[buf], (((uint32_t)my_seq / kXBufs) & 1u) ^ 1u);        // the MMAs that read this buffer last time have finished
                    mbar_wait(&bempty[bslot], (((uint32_t)my_seq / kXBaseRing) & 1u) ^ 1u);  // ... and the epilogue that read this base slot
                    if (tl == 0) {
                        float4* bo = reinterpret_cast<float4*>(base_ring + bslot * kXBaseBytes + g * 32);
                        bo[0] = make_float4(b[0].x, b[0].y, b[1].x, b[1].y);
                        bo[1] = make_float4(b[2].x, b[2].y, b[3].x, b[3].y);
                    }
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const int t = tl + 8 * k;
                        uint4 H, M, L;
                        // finetuned - base (task_vector_loader.py:142): b * -1 + f is the correctly rounded difference
                        k13_split2(__ffma2_rn(b[0], neg1, f[k][0]), H.x, M.x, L.x);
                        k13_split2(__ffma2_rn(b[1], neg1, f[k][1]), H.y, M.y, L.y);
                        k13_split2(__ffma2_rn(b[2], neg1, f[k][2]), H.z, M.z, L.z);
                        k13_split2(__ffma2_rn(b[3], neg1, f[k][3]), H.w, M.w, L.w);
                        if (t < N) {
                            *reinterpret_cast<uint4*>(tile_out + t * 16) = H;
                            *reinterpret_cast<uint4*>(tile_out + (kXNP + t) * 16) = M;
                            *reinterpret_cast<uint4*>(tile_out + (2 * kXNP + t) * 16) = L;
                        }
                    }
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tfull[buf]);
                }
                seq += n_here;
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[stage]);
                if (++stage == (uint32_t)n_stages) { stage = 0; phase ^= 1u; }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kXMma) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kXTmemCols));
}

template <bool FP16B>
static cudaError_t k13_go(const K3Args& a, int n_tasks, int n_tiles, int n_sm, cudaStream_t st) {
    if (a.tile_elems % kXStage != 0) return cudaErrorNotSupported;
    const size_t stage = (size_t)k13_stage_bytes(n_tasks);
    const size_t budget = 232448 - 1024;
    int stages = (int)((budget - kXFixed) / stage);
    if (stages > 8) stages = 8;
    if (stages < 3) return cudaErrorNotSupported;
    const size_t dsm = kXFixed + stages * stage;
    cudaError_t e = cudaFuncSetAttribute(k13_merge_wide_tc<FP16B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm);
    if (e != cudaSuccess) return e;
    const int grid = n_tiles < n_sm ? n_tiles : n_sm;
    k13_merge_wide_tc<FP16B><<<grid, kXThreads, dsm, st>>>(a, n_tasks, n_tiles, stages);
    return cudaGetLastError();
}

#endif  // fp32

// tensor-core pass 2 of the wide path: fp32 inputs, 2..21 tasks, no diagnostics / noise region; cudaErrorNotSupported otherwise
template <>
cudaError_t k13_launch_dtype<SVDQ_DTYPE>(int n_tasks, const K3Args& a, int n_tiles, bool fp16b, int n_sm, cudaStream_t st) {
#if SVDQ_DTYPE == 0
    if (n_tiles <= 0) return cudaSuccess;
    if (n_tasks < 2 || n_tasks > kXMaxTasks || a.info_n != nullptr) return cudaErrorNotSupported;
    return fp16b ? k13_go<true>(a, n_tasks, n_tiles, n_sm, st) : k13_go<false>(a, n_tasks, n_tiles, n_sm, st);
#else
    return cudaErrorNotSupported;
#endif
}

}  // namespace svdq
