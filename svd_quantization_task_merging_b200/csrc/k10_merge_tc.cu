// K10 — pass 2 (basis rows, weighted reconstruction, mask scatter, merged = base + delta) for bf16 inputs on the
// 5th-generation tensor cores.  Same inputs, outputs and reference lines as k3s_reconstruct_merge_staged.cu
// (reconstruct_from_coefficients src/svd_hybrid/merge.py:144-194, merge_parameter :197-312, apply_merged_deltas
// :429-552, the fp16 cast of the bases src/svd_hybrid/cli.py:355-361); used for bf16 checkpoints with up to 8 task
// vectors, no fused diagnostics, no noise region.  Why: pass 2 moves 22 B per element but the CUDA-core kernel
// executes ~155 thread instructions per element (bf16 unpack, centring, N r FMAs, fp16 round trip), issue-bound at
// ~50 % of the HBM roofline (profiles/r1_ncu_full_llama_bf16_k1s_k3s.csv).
//
// The basis row of element d is u_d = (tau_d - mean_d 1) W with tau_d the N task-vector values.  tau_d is EXACTLY a
// bf16 vector (bf16(ft - base)), W is split into three bf16 pieces W = W1 + W2 + W3 (24 bits), so
//     tau_d W = tau_d W1 + tau_d W2 + tau_d W3
// is a sum of exact bf16 x bf16 products accumulated in fp32 by the tensor core, and the centring enters as
// -mean_d (1^T W) with mean_d = (1^T tau_d) / n (1^T W is round-off of a column sum that is zero in exact arithmetic).
// One tcgen05.mma (M = 128 elements, N = 32, K = 16) per 128 elements:
//     A [128 x 16]  = [tau | tau]   (MN-major: the K1 tile layout [group of 8 elements][task][8 elements]; the second
//                                    K group ALIASES the first through a leading byte offset of 0)
//     B [16 x 32]   = [[W1, W3, 1, 0], [W2, 0, 0, 0]]   (K-major, 1 KB, rebuilt per tile, carried in every tile buffer)
//     D [128 x 32]  : columns 0-7 = tau (W1 + W2), 8-15 = tau W3, 16 = sum_t tau_t
// The epilogue warps read their element's row from TMEM (tcgen05.ld), finish u_j, round it to fp16 when the bases are
// stored in fp16, contract with cbar, add the mean, apply the packed tall mask and stream out merged = base + delta.
//
// Roles (one persistent CTA per SM, 18 warps): warps 0-7 transform (raw ring -> bf16 task vectors + base + B in the
// tile buffer), warp 8 TMA producer, warp 9 MMA issuer, warps 10-17 epilogue.  Bound: HBM (22 B per element).
#include "svdq_kernels.h"
#include "tc_common.cuh"

#ifndef SVDQ_DTYPE
#define SVDQ_DTYPE 1
#endif

namespace svdq {

#if SVDQ_DTYPE == 1

constexpr int kMtStages = 3;                   // raw ring depth (stages of kMtStageElems elements)
constexpr int kMtStageElems = 2 * kStep;
constexpr int kMtTiles = 4;                    // tile buffer ring (operands of the MMAs; released by the MMAs' completion)
constexpr int kMtBaseRing = 6;                 // ring of the chunks' base values (released by the epilogue warps)
constexpr int kMtRowStride = kMtStageElems * 2 + 16;
constexpr int kMtATile = kStep * 8 * 2;        // 16 KB: [128 groups][8 tasks][8 elements] bf16
constexpr int kMtBOff = kMtATile;              // 1 KB: B operand
constexpr int kMtTileBytes = kMtBOff + 1024;
constexpr int kMtBaseBytes = kStep * 2;        // 2 KB: base values of one chunk (bf16)
constexpr int kMtTransform = 256;
constexpr int kMtEpiWarps = 8;                 // 4 TMEM lane quadrants x 2 sets of row blocks (16 warps measured slower:
                                               // 72-register cap, four readers per TMEM quadrant)
constexpr int kMtThreads = (10 + kMtEpiWarps) * 32;
constexpr int kMtTmemCols = 512;               // two accumulator sets of 8 x 32 columns

__host__ __device__ constexpr int mt_stage_bytes() { return 9 * kMtRowStride; }
__host__ __device__ constexpr int mt_smem_bytes() {
    return kMtTiles * kMtTileBytes + kMtBaseRing * kMtBaseBytes + kMtStages * mt_stage_bytes() + 64 * 8 + 256;
}

__device__ __forceinline__ uint32_t bf16_bits_rn(float x) { return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(x)); }
__device__ __forceinline__ float bf16_bits_to_f32(uint32_t b) { return __uint_as_float(b << 16); }

// value of the B operand at (k, n) for this parameter: see the file comment.  W[t][j] with t >= nt or j >= r is 0.
// Centring is folded in: u = (tau - mean 1) W = tau (W - 1 (1^T W) / n) over the n tasks that have the parameter
// (1^T W is round-off of a sum that is zero in exact arithmetic; formed in fp64 and split together with W, so the
// pieces' column sums vanish to 2^-27 of |W| and a large common component of the task vectors cancels as exactly
// as with explicit centring).
template <int NT>
__device__ __forceinline__ uint32_t mt_b_value(const float* W, int r, int k, int n, uint32_t present, int n_active, int center) {
    const int t = k & 7, piece_row = k >> 3;
    if (t >= NT) return 0u;
    if (n < 16) {
        const int j = n & 7, blk = n >> 3;
        if (j >= r || j >= NT) return 0u;
        double w = (double)W[t * NT + j];
        if (center && ((present >> t) & 1u)) {
            double s = 0.0;
#pragma unroll
            for (int u = 0; u < NT; ++u) s += (double)W[u * NT + j];
            w -= s / (double)(n_active > 0 ? n_active : 1);
        }
        const uint32_t w1 = bf16_bits_rn((float)w);
        const double r1 = w - (double)bf16_bits_to_f32(w1);
        const uint32_t w2 = bf16_bits_rn((float)r1);
        const double r2 = r1 - (double)bf16_bits_to_f32(w2);
        const uint32_t w3 = bf16_bits_rn((float)r2);
        if (blk == 0) return piece_row == 0 ? w1 : w2;
        return piece_row == 0 ? w3 : 0u;
    }
    if (n == 16) return piece_row == 0 ? 0x3F80u : 0u;        // 1.0: column 16 = sum over the tasks
    return 0u;
}

template <int NT, bool FP16B>
__global__ void __launch_bounds__(kMtThreads, 1) k10_merge_tc(const K3Args a, const int n_tiles) {
    using T = __nv_bfloat16;
    constexpr int STAGES = kMtStages;
    constexpr int kStageBytes = mt_stage_bytes();

    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* tile_buf = smem;
    unsigned char* base_ring = tile_buf + kMtTiles * kMtTileBytes;
    unsigned char* stage_base = base_ring + kMtBaseRing * kMtBaseBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(stage_base + STAGES * kStageBytes);
    uint64_t* full = bars;                  // [STAGES] producer -> transform
    uint64_t* empty = full + STAGES;        // [STAGES] transform -> producer
    uint64_t* tfull = empty + STAGES;       // [kMtTiles] transform -> MMA
    uint64_t* tempty = tfull + kMtTiles;    // [kMtTiles] MMA (operands read) -> transform
    uint64_t* afull = tempty + kMtTiles;    // [2] MMA -> epilogue
    uint64_t* aempty = afull + 2;           // [2] epilogue -> MMA
    uint64_t* bempty = aempty + 2;          // [kMtBaseRing] epilogue (base values read) -> transform
    int* s_direct = reinterpret_cast<int*>(bempty + kMtBaseRing);
    __shared__ const void* s_ptr[NT + 1];
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kMtTransform / 32); }
        for (int s = 0; s < kMtTiles; ++s) { mbar_init(&tfull[s], kMtTransform / 32); mbar_init(&tempty[s], 1); }
        for (int s = 0; s < kMtBaseRing; ++s) mbar_init(&bempty[s], kMtEpiWarps);
        for (int s = 0; s < 2; ++s) { mbar_init(&afull[s], 1); mbar_init(&aempty[s], kMtEpiWarps); }
        mbar_fence_init();
    }
    if (warp == 9) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(kMtTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;

    if (warp == 8) {
        // ================= TMA producer: lane i streams tensor i (base, then the NT fine-tuned tensors) ============
        PipeState ps;
        const bool is_tensor = lane <= NT;
        // Per-parameter values are re-read only when the CTA's next tile belongs to another parameter: consecutive
        // tiles of a CTA mostly come from the same (large) tensor, and the dependent global loads of a tile boundary
        // otherwise cost every role a few hundred nanoseconds per tile.
        int prev_p = -1;
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
                if (is_tensor) {
                    const void* q = a.tensors[(int64_t)p * (NT + 1) + lane];
                    my_ptr = reinterpret_cast<const unsigned char*>(q ? q : a.tensors[(int64_t)p * (NT + 1)]);
                }
            }
            const int64_t stop = min(start + (int64_t)a.tile_elems, numel);
            for (int64_t e0 = start; e0 < stop; e0 += kMtStageElems) {
                if (lane == 0) mbar_wait(&empty[ps.stage], ps.phase ^ 1u);
                __syncwarp();
                unsigned char* sb = stage_base + ps.stage * kStageBytes;
                if (e0 + kMtStageElems <= numel) {
                    if (lane == 0) {
                        s_direct[ps.stage] = 0;
                        mbar_arrive_expect_tx(&full[ps.stage], (uint32_t)((NT + 1) * kMtStageElems * 2));
                    }
                    __syncwarp();
                    if (is_tensor) bulk_g2s(sb + lane * kMtRowStride, my_ptr + e0 * 2, kMtStageElems * 2, &full[ps.stage]);
                } else if (lane == 0) {
                    s_direct[ps.stage] = 1;          // tail of the parameter: the transform warps load it themselves
                    mbar_arrive(&full[ps.stage]);
                }
                ps.advance<STAGES>();
            }
        }
    } else if (warp == 9) {
        // ================= MMA issuer: the whole warp walks the loops, one elected lane issues ==========================
        const bool leader = elect_one();
        PipeState tb, ab;
        const uint32_t idesc = tc_idesc(1u, 1u, 128, 32, /*a_mn=*/1u, /*b_mn=*/0u);
        const uint64_t desc_a0 = tc_smem_desc(smem_u32(tile_buf), /*lbo (K groups)=*/0, /*sbo (M groups)=*/128);
        const uint64_t desc_b0 = tc_smem_desc(smem_u32(tile_buf) + kMtBOff, /*lbo (K chunks)=*/128, /*sbo (N groups)=*/256);
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
                    const uint64_t da = desc_a0 + (uint64_t)(tb.stage * (kMtTileBytes >> 4));
                    const uint64_t db = desc_b0 + (uint64_t)(tb.stage * (kMtTileBytes >> 4));
                    const uint32_t td = tmem + ab.stage * 256;
#pragma unroll
                    for (int m = 0; m < 8; ++m) tc_mma_f16(td + m * 32, da + (uint64_t)(m * (2048 >> 4)), db, idesc, 0u);
                    tc_commit(&tempty[tb.stage]);
                    tc_commit(&afull[ab.stage]);
                }
                __syncwarp();
                tb.advance<kMtTiles>();
                ab.advance<2>();
            }
        }
    } else if (warp >= 10) {
        // ================= epilogue warps ===========================================================================
        const int q = warp & 3;                         // TMEM lane quadrant
        constexpr int kSets = kMtEpiWarps / 4;          // row blocks m = mset, mset + kSets, ... belong to this warp
        const int mset = (warp - 10) >> 2;
        PipeState bb, ab;
        const bool center = a.center != 0;
        int prev_p = -1;
        int64_t numel = 0;
        bool solved = false, has_mask = false, pow2 = true;
        float tail_add = 0.0f, mean_scale = 0.0f, n_f = 1.0f, inv_n = 1.0f;
        const uint32_t* packed = nullptr;
        float* outp = nullptr;
        float cb[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) cb[j] = 0.0f;
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
                solved = a.info[(int64_t)p * 8 + 0] == kSolved;
                const int n_active = a.info[(int64_t)p * 8 + 1];
                const int r = a.info[(int64_t)p * 8 + 4];
                tail_add = a.scal[(int64_t)p * 4 + 1];
                mean_scale = a.scal[(int64_t)p * 4 + 2];
                has_mask = a.has_mask[p] != 0;
                packed = has_mask ? a.packed + a.pmask_off[p] : nullptr;
                outp = a.out[p];
                n_f = (float)(n_active > 0 ? n_active : 1);
                pow2 = (n_active & (n_active - 1)) == 0;
                inv_n = __fdiv_rn(1.0f, n_f);
#pragma unroll
                for (int j = 0; j < 8; ++j) cb[j] = (j < NT && j < r && solved) ? a.cbar[(int64_t)p * NT + j] : 0.0f;
            }
            const int64_t stop = min(start + (int64_t)a.tile_elems, numel);
            for (int64_t e0 = start; e0 < stop; e0 += kStep) {
                // this thread's elements of the chunk: loc = m * 128 + q * 32 + lane for m = mset, mset + kSets, ...
                const int left = (int)min(stop - e0, (int64_t)kStep) - (q * 32 + lane);       // element m valid iff m * 128 < left
                const int left_w = (int)min(stop - e0, (int64_t)kStep) - q * 32;              // ... for the warp's row block
                float* op = outp + e0 + q * 32 + lane;
                uint32_t pw[8 / kSets];
#pragma unroll
                for (int mm = 0; mm < 8 / kSets; ++mm)
                    pw[mm] = (has_mask && solved && (kSets * mm + mset) * 128 < left_w)
                                 ? __ldg(packed + (e0 >> 5) + (kSets * mm + mset) * 4 + q) : 0xFFFFFFFFu;
                mbar_wait(&afull[ab.stage], ab.phase);
                tc_fence_after();
                const uint16_t* bp = reinterpret_cast<const uint16_t*>(base_ring + bb.stage * kMtBaseBytes) + q * 32 + lane;
                const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab.stage * 256);
#pragma unroll
                for (int pr = 0; pr < 4 / kSets; ++pr) {        // two row blocks in flight
                    const int m0 = 2 * kSets * pr + mset, m1 = m0 + kSets;
                    uint32_t d0[16], d1[16], s0, s1;
                    tc_ld16_nowait(trow + m0 * 32, d0);
                    tc_ld1_nowait(trow + m0 * 32 + 16, s0);
                    tc_ld16_nowait(trow + m1 * 32, d1);
                    tc_ld1_nowait(trow + m1 * 32 + 16, s1);
                    const float base0 = bf16_bits_to_f32((uint32_t)bp[m0 * 128]);
                    const float base1 = bf16_bits_to_f32((uint32_t)bp[m1 * 128]);
                    tc_wait_ld();
                    if (pr == 4 / kSets - 1) {          // last read of this chunk's accumulators and base values
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) { mbar_arrive(&aempty[ab.stage]); mbar_arrive(&bempty[bb.stage]); }
                    }
                    float res0 = base0, res1 = base1;
                    if (solved) {
                        const float ms0 = __uint_as_float(s0), ms1 = __uint_as_float(s1);
                        const float mean0 = center ? (pow2 ? ms0 * inv_n : __fdiv_rn(ms0, n_f)) : 0.0f;
                        const float mean1 = center ? (pow2 ? ms1 * inv_n : __fdiv_rn(ms1, n_f)) : 0.0f;
                        // packed 2-wide math over column pairs (j, j + 1): add the W3 part, fp16 round trip of the
                        // basis entries, contraction with cbar in two partial sums (even / odd columns)
                        float2 p0 = make_float2(0.0f, 0.0f), p1 = make_float2(0.0f, 0.0f);
#pragma unroll
                        for (int j = 0; j < 8; j += 2) {
                            float2 u = __fadd2_rn(make_float2(__uint_as_float(d0[j]), __uint_as_float(d0[j + 1])),
                                                  make_float2(__uint_as_float(d0[8 + j]), __uint_as_float(d0[9 + j])));
                            float2 v = __fadd2_rn(make_float2(__uint_as_float(d1[j]), __uint_as_float(d1[j + 1])),
                                                  make_float2(__uint_as_float(d1[8 + j]), __uint_as_float(d1[9 + j])));
                            if (FP16B) {                // the stored basis is fp16 (cli.py:355-361)
                                u = __half22float2(__float22half2_rn(u));
                                v = __half22float2(__float22half2_rn(v));
                            }
                            const float2 c2 = make_float2(cb[j], cb[j + 1]);
                            p0 = __ffma2_rn(u, c2, p0);
                            p1 = __ffma2_rn(v, c2, p1);
                        }
                        const float a0 = p0.x + p0.y, a1 = p1.x + p1.y;
                        const float val0 = fmaf(mean0, mean_scale, a0) + tail_add;
                        const float val1 = fmaf(mean1, mean_scale, a1) + tail_add;
                        res0 = base0 + (((pw[2 * pr] >> lane) & 1u) ? val0 : 0.0f);
                        res1 = base1 + (((pw[2 * pr + 1] >> lane) & 1u) ? val1 : 0.0f);
                    }
                    if (m0 * 128 < left) __stcs(op + m0 * 128, res0);
                    if (m1 * 128 < left) __stcs(op + m1 * 128, res1);
                }
                bb.advance<kMtBaseRing>();
                ab.advance<2>();
            }
        }
    } else {
        // ================= transform warps ==========================================================================
        PipeState ps, tb, bb;
        const int t = lane & 7;
        int pending = -1;
        int prev_p = -1;
        int64_t numel = 0;
        uint4 bfrag[2] = {make_uint4(0u, 0u, 0u, 0u), make_uint4(0u, 0u, 0u, 0u)};
        int p_ahead = a.tile_param[blockIdx.x], local_ahead = a.tile_local[blockIdx.x];    // grid <= n_tiles
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int p = p_ahead;                      // tile -> (parameter, position) is read one tile ahead
            const int64_t start = (int64_t)local_ahead * a.tile_elems;
            if (tile + (int)gridDim.x < n_tiles) {
                p_ahead = a.tile_param[tile + gridDim.x];
                local_ahead = a.tile_local[tile + gridDim.x];
            }
            const bool new_param = p != prev_p;         // same decision in every transform warp (same tile sequence)
            if (new_param) {
                prev_p = p;
                numel = a.numel[p];
                named_bar_sync(1, kMtTransform);        // previous parameter finished with s_ptr
                if (tid <= NT) {
                    const void* qp = a.tensors[(int64_t)p * (NT + 1) + tid];
                    s_ptr[tid] = qp ? qp : a.tensors[(int64_t)p * (NT + 1)];
                }
                named_bar_sync(1, kMtTransform);
            }
            const int64_t stop = min(start + (int64_t)a.tile_elems, numel);
            // B operand of this parameter: warp 0 keeps its 1 KB image in registers (lane -> rows lane, lane + 32 of
            // the 64 sixteen-byte rows) and drops it into every tile buffer of the tile
            if (warp == 0 && new_param) {
                const bool solved = a.info[(int64_t)p * 8 + 0] == kSolved;
                const int r = solved ? a.info[(int64_t)p * 8 + 4] : 0;
                const int n_active = a.info[(int64_t)p * 8 + 1];
                uint32_t present = 0;
#pragma unroll
                for (int u = 0; u < NT; ++u) present |= (a.tensors[(int64_t)p * (NT + 1) + 1 + u] != nullptr ? 1u : 0u) << u;
                const float* W = a.W + (int64_t)p * NT * NT;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int row = lane + 32 * half;                  // = (n / 8) * 16 + (k / 8) * 8 + (n % 8)
                    const int n = (row >> 4) * 8 + (row & 7), kc = (row >> 3) & 1;
                    uint32_t v[8];
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) v[kk] = mt_b_value<NT>(W, r, kc * 8 + kk, n, present, n_active, a.center);
                    bfrag[half] = make_uint4(v[0] | (v[1] << 16), v[2] | (v[3] << 16), v[4] | (v[5] << 16), v[6] | (v[7] << 16));
                }
            }
            for (int64_t s0 = start; s0 < stop; s0 += kMtStageElems) {
                mbar_wait(&full[ps.stage], ps.phase);
                const unsigned char* sbase = stage_base + ps.stage * kStageBytes;
                const bool direct = s_direct[ps.stage] != 0;
#pragma unroll 1
                for (int sub = 0; sub < kMtStageElems / kStep; ++sub) {
                    const int64_t e0 = s0 + (int64_t)sub * kStep;
                    if (e0 >= stop) break;
                    const bool last_sub = sub == kMtStageElems / kStep - 1 || e0 + kStep >= stop;
                    const unsigned char* sb = sbase + sub * (kStep * 2);
                    unsigned char* tile_out = tile_buf + tb.stage * kMtTileBytes;
                    unsigned char* base_out = base_ring + bb.stage * kMtBaseBytes;
                    uint4 b[4], f[4];
                    if (!direct) {
#pragma unroll
                        for (int it = 0; it < 4; ++it) {
                            const int g = warp * 16 + it * 4 + (lane >> 3);
                            b[it] = *reinterpret_cast<const uint4*>(sb + g * 16);
                            f[it] = *reinterpret_cast<const uint4*>(sb + (t + 1) * kMtRowStride + g * 16);
                        }
                    } else {
#pragma unroll
                        for (int it = 0; it < 4; ++it) {
                            const int g = warp * 16 + it * 4 + (lane >> 3);
                            const int64_t e = e0 + (int64_t)g * 8;
                            uint32_t bw[4] = {0u, 0u, 0u, 0u}, fw[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                            for (int c = 0; c < 8; ++c) {
                                if (e + c < numel) {
                                    bw[c >> 1] |= (uint32_t)reinterpret_cast<const uint16_t*>(s_ptr[0])[e + c] << (16 * (c & 1));
                                    if (t < NT) fw[c >> 1] |= (uint32_t)reinterpret_cast<const uint16_t*>(s_ptr[t + 1])[e + c] << (16 * (c & 1));
                                }
                            }
                            b[it] = make_uint4(bw[0], bw[1], bw[2], bw[3]);
                            f[it] = t < NT ? make_uint4(fw[0], fw[1], fw[2], fw[3]) : b[it];
                        }
                    }
                    if (pending >= 0) {                 // hand the previous chunk's tile over while these loads land
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
                    if (last_sub) {                     // the raw stage lives in registers now: hand it back
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&empty[ps.stage]);
                    }
                    // only now is the destination needed: the MMAs that read this tile buffer / the epilogue warps
                    // that read this base slot have finished (the previous chunk's tile is already handed over)
                    mbar_wait(&tempty[tb.stage], tb.phase ^ 1u);
                    mbar_wait(&bempty[bb.stage], bb.phase ^ 1u);
#pragma unroll
                    for (int it = 0; it < 4; ++it) {
                        const int g = warp * 16 + it * 4 + (lane >> 3);
                        *reinterpret_cast<uint4*>(tile_out + g * 128 + t * 16) = d[it];
                        if (t == 0) *reinterpret_cast<uint4*>(base_out + g * 16) = b[it];
                    }
                    if (warp == 0) {
                        *reinterpret_cast<uint4*>(tile_out + kMtBOff + lane * 16) = bfrag[0];
                        *reinterpret_cast<uint4*>(tile_out + kMtBOff + (lane + 32) * 16) = bfrag[1];
                    }
                    pending = (int)tb.stage;
                    tb.advance<kMtTiles>();
                    bb.advance<kMtBaseRing>();
                }
                ps.advance<STAGES>();
            }
        }
        if (pending >= 0) {
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tfull[pending]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 9) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kMtTmemCols));
}

template <int NT>
static cudaError_t launch_mt(const K3Args& a, int n_tiles, bool fp16b, int n_sm, cudaStream_t st) {
    if (n_tiles <= 0) return cudaSuccess;
    constexpr int smem = mt_smem_bytes();
    const int grid = n_tiles < n_sm ? n_tiles : n_sm;
    cudaError_t e;
    if (fp16b) {
        e = cudaFuncSetAttribute(k10_merge_tc<NT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        k10_merge_tc<NT, true><<<grid, kMtThreads, smem, st>>>(a, n_tiles);
    } else {
        e = cudaFuncSetAttribute(k10_merge_tc<NT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        k10_merge_tc<NT, false><<<grid, kMtThreads, smem, st>>>(a, n_tiles);
    }
    return cudaGetLastError();
}

#endif  // bf16

// tensor-core pass 2 exists for bf16 inputs, nt <= 8, no diagnostics / noise region; cudaErrorNotSupported otherwise
template <>
cudaError_t k10_launch_dtype<SVDQ_DTYPE>(int nt, const K3Args& a, int n_tiles, bool fp16b, int n_sm, cudaStream_t st) {
#if SVDQ_DTYPE == 1
    if (a.info_n != nullptr) return cudaErrorNotSupported;
    switch (nt) {
#define SVDQ_CASE(N) case N: return launch_mt<N>(a, n_tiles, fp16b, n_sm, st);
        SVDQ_CASE(1) SVDQ_CASE(2) SVDQ_CASE(3) SVDQ_CASE(4) SVDQ_CASE(5) SVDQ_CASE(6) SVDQ_CASE(7) SVDQ_CASE(8)
#undef SVDQ_CASE
        default: return cudaErrorNotSupported;
    }
#else
    return cudaErrorNotSupported;
#endif
}

}  // namespace svdq
