// Per-parameter solve of the SVD-Hybrid path: centred Gram -> eigenpairs -> energy rank ->
// closed-form task coefficients -> fp16 high block + multi-stage RTVQ low block -> weighted
// average -> projection matrix for the reconstruction pass.
//
// The code is written "lane-SPMD": every function takes a Lanes object {lane, nl, sync()}.
// On the GPU one warp runs it (nl = 32, sync = __syncwarp) out of shared memory; compiled for
// the host (tests only, tests/hostcheck) the same source runs serially (nl = 1).  All
// cross-lane communication goes through the scratch arrays between sync() points, so both
// executions compute the same values.
//
// Reference behaviour restated here (paths relative to /root/reference):
//   centring            src/svd_hybrid/basis.py:103-111  (T - T.mean(dim=1)  <=>  G_c = H G H)
//   thin SVD            src/svd_hybrid/basis.py:241      (sigma_j = sqrt(lambda_j(G_c)), V = eigvecs)
//   rank selection      src/svd_hybrid/basis.py:147-156,199-211
//   projection          src/svd_hybrid/compress.py:13-19 (c_t = U^T (tau_t - mean) = Sigma V^T e_t)
//   fp16 high block     src/svd_hybrid/compress.py:44-45
//   RTVQ                src/svd_hybrid/rtvq.py:4-103
//   weighted average    src/svd_hybrid/merge.py:96-139
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define SVDQ_HD __host__ __device__ inline
#else
#define SVDQ_HD inline
#endif

namespace svdq {

constexpr int kCoreMaxTasks = 32;
constexpr int kCoreLd = kCoreMaxTasks + 1;      // padded leading dimension (bank spread)
constexpr int kCoreMaxStages = 8;

// ---- separately rounded fp32 arithmetic (torch-eager semantics: no FMA contraction) ---------
SVDQ_HD float f_mul(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);
#else
    return a * b;
#endif
}
SVDQ_HD float f_add(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fadd_rn(a, b);
#else
    return a + b;
#endif
}
SVDQ_HD float f_sub(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fsub_rn(a, b);
#else
    return a - b;
#endif
}
SVDQ_HD float f_div(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fdiv_rn(a, b);
#else
    return a / b;
#endif
}
SVDQ_HD uint32_t f_bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
SVDQ_HD float bits_f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
SVDQ_HD bool f_isnan(float f) { return (f_bits(f) & 0x7fffffffu) > 0x7f800000u; }

// ---- IEEE binary16 conversion, round-to-nearest-even (integer code: same bits on host/GPU) --
SVDQ_HD uint16_t f32_to_f16_bits(float f) {
    uint32_t x = f_bits(f);
    const uint32_t sign = (x >> 16) & 0x8000u;
    x &= 0x7fffffffu;
    if (x >= 0x7f800000u) return (uint16_t)(sign | 0x7c00u | (x > 0x7f800000u ? 0x0200u : 0u));
    if (x < 0x38800000u) {                       // below 2^-14: binary16 subnormal or zero
        if (x <= 0x33000000u) return (uint16_t)sign;   // <= 2^-25 rounds (ties-to-even) to zero
        const uint32_t e = x >> 23;
        const uint32_t m = (x & 0x7fffffu) | 0x800000u;
        const uint32_t shift = 126u - e;           // 14..24
        uint32_t q = m >> shift;
        const uint32_t rem = m & ((1u << shift) - 1u), half = 1u << (shift - 1u);
        if (rem > half || (rem == half && (q & 1u))) ++q;
        return (uint16_t)(sign | q);
    }
    const uint32_t e = (x >> 23) - 112u, m = x & 0x7fffffu;
    uint32_t q = (e << 10) | (m >> 13);
    const uint32_t rem = m & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (q & 1u))) ++q;   // carry into the exponent is right
    if (q >= 0x7c00u) q = 0x7c00u;
    return (uint16_t)(sign | q);
}
SVDQ_HD float f16_bits_to_f32(uint16_t h) {
    const uint32_t sign = ((uint32_t)h & 0x8000u) << 16, e = (h >> 10) & 0x1fu, m = h & 0x3ffu;
    if (e == 0) {
        const float v = (float)m * 5.9604644775390625e-8f;   // m * 2^-24, exact
        return sign ? -v : v;
    }
    if (e == 31) return bits_f(sign | 0x7f800000u | (m << 13));
    return bits_f(sign | ((e + 112u) << 23) | (m << 13));
}

// ---- asymmetric quantiser on a short vector held by ONE lane (n <= 32) -----------------------
// rtvq.py:4-27.  `int / tensor` in Python is reciprocal()*int: two roundings.
struct QuantScalars { float scale, zp; };

SVDQ_HD QuantScalars asym_scalars(float lo, float hi, int bits) {
    const float qmax = (float)((1 << bits) - 1);
    const float recip = f_div(1.0f, f_sub(hi, lo));
    QuantScalars q;
    q.scale = f_mul(recip, qmax);
    q.zp = -rintf(f_mul(q.scale, lo));
    return q;
}
SVDQ_HD int asym_code(float x, QuantScalars q, int bits) {
    const float qmax = (float)((1 << bits) - 1);
    float t = rintf(f_add(f_mul(q.scale, x), q.zp));
    if (f_isnan(t)) return 0;                    // torch: NaN -> clamp keeps NaN -> .to(uint8) = 0
    t = t < 0.0f ? 0.0f : (t > qmax ? qmax : t);
    return (int)t;
}
SVDQ_HD float asym_decode(int code, QuantScalars q) { return f_div(f_sub((float)code, q.zp), q.scale); }

// Multi-stage residual quantisation of x[0..n) (rtvq.py:39-103).  codes: [stages][ld_codes].
// deq_sum[i] = 0 + d_0 + d_1 + ... (left-to-right, as multistage_residual_dequantization).
SVDQ_HD void rtvq_short(const float* x, int n, int bits, int stages, uint8_t* codes, int ld_codes,
                        float* scale, float* zp, float* resnorm, float* deq_sum) {
    float res[kCoreMaxTasks];
    for (int i = 0; i < n; ++i) { res[i] = x[i]; deq_sum[i] = 0.0f; }
    for (int s = 0; s < stages; ++s) {
        double ss = 0.0;
        float lo = res[0], hi = res[0];
        bool nan = false;
        for (int i = 0; i < n; ++i) {
            const float v = res[i];
            ss += (double)v * (double)v;
            nan = nan || f_isnan(v);
            lo = v < lo ? v : lo;
            hi = v > hi ? v : hi;
        }
        if (nan) { lo = bits_f(0x7fc00000u); hi = lo; }     // torch.min/max propagate NaN
        resnorm[s] = (float)sqrt(ss);
        const QuantScalars q = asym_scalars(lo, hi, bits);
        scale[s] = q.scale;
        zp[s] = q.zp;
        for (int i = 0; i < n; ++i) {
            const int c = asym_code(res[i], q, bits);
            codes[s * ld_codes + i] = (uint8_t)c;
            const float d = asym_decode(c, q);
            res[i] = f_sub(res[i], d);
            deq_sum[i] = f_add(deq_sum[i], d);
        }
    }
}

// ---- energy rank selection (basis.py:147-156,199-211) -----------------------------------------
// cumsum accumulates in double and rounds each prefix to fp32 (ATen cpu_cum_base_kernel); the
// fp32 total is restated as the double sum rounded once (<= 1 ulp from ATen's lane order).
SVDQ_HD int select_rank_f32(const float* S, int r, float thr, int max_rank, int min_rank, float* energy_retained) {
    if (r <= 0) { *energy_retained = 0.0f; return 0; }
    double tot_d = 0.0;
    for (int i = 0; i < r; ++i) tot_d += (double)f_mul(S[i], S[i]);
    const float tot = (float)tot_d;
    const bool flat = tot < 1e-10f;
    float cum[kCoreMaxTasks];
    int below = 0;
    double acc = 0.0;
    for (int i = 0; i < r; ++i) {
        float c = 1.0f;
        if (!flat) { acc += (double)f_mul(S[i], S[i]); c = f_div((float)acc, tot); }
        cum[i] = c;
        if (c < thr) ++below;
    }
    int k = below + 1;
    if (k < min_rank) k = min_rank;
    if (max_rank > 0 && k > max_rank) k = max_rank;
    if (k > r) k = r;
    *energy_retained = cum[k - 1];
    return k;
}

// ---- scratch + I/O of one parameter solve -----------------------------------------------------
struct SolveScratch {
    double A[kCoreMaxTasks * kCoreLd];     // working matrix (Gram -> diagonal)
    double V[kCoreMaxTasks * kCoreLd];     // accumulated rotations, V[t][j]
    double Vs[kCoreMaxTasks * kCoreLd];    // sorted, sign-fixed eigenvectors Vs[t][j]
    double tmp[kCoreMaxTasks];
    double cs[2 * kCoreMaxTasks];          // (c, s) of the current Jacobi round
    int pq[2 * kCoreMaxTasks];             // (p, q) of the current Jacobi round (p = -1: no rotation)
    double sigma[kCoreMaxTasks];
    int idx[kCoreMaxTasks];                // active-task compaction: a -> task position
    int perm[kCoreMaxTasks];
    int n, r, k, r_eff;
};

struct SolveConfig {
    int n_tasks;          // NT: stride of every per-task array below
    int center;
    float energy_threshold;
    int max_rank;         // <= 0: no cap
    int min_mask_size;
    int bits, stages;
};

struct SolveIn {
    const double* G;          // [NT*NT] uncentred masked Gram, full symmetric, row-major
    int64_t dm;               // masked element count (rows of the task matrix)
    int has_mask;
    uint32_t present;         // bit t set: task t has this parameter
    const double* weights;    // [NT] merge weights by task position (un-normalised is fine); null: solve only,
                              //      average_param runs later
    const int32_t* avg_order; // [NT] task positions in sorted-name order (merge.py:89)
    const double* sign_ref;   // [NT*NT] optional Vh_ref[j][t] (test-only sign alignment) or null
    const int32_t* cluster_of = nullptr;  // [NT] cluster index (0 .. NT-1) of each task position, or null: no clustering
    const double* omega = nullptr;        // [NT] cross-cluster weight by cluster index (softmax of the cluster scores,
                                          //      renormalised: clustering.py:399-423); used only with cluster_of
};

struct SolveOut {             // all strides are NT (= cfg.n_tasks); S = cfg.stages
    int32_t* info;            // [8]: status, n_active, r, k, r_eff, 0, 0, 0
    float* sv;                // [NT] singular values (fp32), first r valid
    float* scal;              // [4]: energy_retained = cum[k-1]; tail_add = sum_{j>=r_eff} 0*cbar[j]
                              //      (0, or NaN when a zero-direction coefficient is NaN/inf, so the
                              //      reconstruction reproduces the reference's NaN propagation); 0; 0
    float* coef;              // [NT*NT] coef[t][j] raw fp32 coefficients
    uint16_t* chigh;          // [NT*NT] fp16 bits, j < k valid
    uint8_t* codes;           // [NT*S*NT] codes[t][s][i], i < r-k valid
    float* qscale;            // [NT*S]
    float* qzp;               // [NT*S]
    float* qres;              // [NT*S] residual norm before each stage
    float* chat;              // [NT*NT] chat[t][j]: coefficient actually used by merge/diagnostics
    float* cbar;              // [NT] weighted-average coefficients (first r valid)
    float* W;                 // [NT*NT] W[t][j]: u_d[j] = sum_t (tau_d[t] - mean_d) W[t][j]
    float* gvec;              // [NT] gvec[t] = sum_j W[t][j] cbar[j] (fp32-basis shortcut)
    double* V;                // [NT*NT] V[t][j] right singular vectors (sorted, sign-fixed)
};

enum SolveStatus : int { kSolved = 0, kSkippedSmallMask = 1, kSkippedEmpty = 2 };

// Parallel-order Jacobi eigensolver on the symmetric n x n matrix A (ld kCoreLd); V accumulates the
// rotations.  One sweep = n-1 (n even) rounds of a round-robin tournament; the <= n/2 rotations of a
// round touch disjoint index pairs, so their angles are computed side by side (one fp64 div/sqrt
// latency chain per ROUND instead of per rotation) and applied as A <- J^T A J in two phases.
// cs: scratch for the (c, s) of the current round, 2 * kCoreMaxTasks doubles.
template <class L>
SVDQ_HD void jacobi_eig(double* A, double* V, double* cs, int* pq, int n, L& ln) {
    for (int i = ln.lane; i < n; i += ln.nl)
        for (int j = 0; j < n; ++j) V[i * kCoreLd + j] = (i == j) ? 1.0 : 0.0;
    ln.sync();
    const int m = (n + 1) & ~1;                 // players in the tournament (one bye when n is odd)
    const int npairs = m / 2;
    for (int sweep = 0; sweep < 64 && n > 1; ++sweep) {
        double off = 0.0, dg = 0.0;
        for (int i = 0; i < n; ++i) {
            dg += A[i * kCoreLd + i] * A[i * kCoreLd + i];
            for (int j = i + 1; j < n; ++j) off += A[i * kCoreLd + j] * A[i * kCoreLd + j];
        }
        if (off <= 1e-34 * dg || off == 0.0) break;      // every lane sees the same values
        for (int round = 0; round < m - 1; ++round) {
            // ---- rotation angles of this round's disjoint pairs -------------------------------------
            for (int k = ln.lane; k < npairs; k += ln.nl) {
                int p = (k == 0) ? (m - 1) : (round + k) % (m - 1);
                int q = (k == 0) ? round : (round + m - 1 - k) % (m - 1);
                if (p > q) { const int t = p; p = q; q = t; }
                double c = 1.0, s = 0.0;
                bool rot = false;
                if (q < n) {
                    const double app = A[p * kCoreLd + p], aqq = A[q * kCoreLd + q], apq = A[p * kCoreLd + q];
                    if (!(fabs(apq) <= 1e-300 || fabs(apq) <= 1e-19 * sqrt(fabs(app * aqq)))) {
                        const double theta = (aqq - app) / (2.0 * apq);
                        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                        c = 1.0 / sqrt(t * t + 1.0);
                        s = t * c;
                        rot = true;
                    }
                }
                cs[2 * k] = c; cs[2 * k + 1] = s;
                pq[2 * k] = rot ? p : -1; pq[2 * k + 1] = q;
            }
            ln.sync();
            // ---- A <- A J and V <- V J : item = (pair k, row i) ----------------------------------------
            for (int it = ln.lane; it < npairs * n; it += ln.nl) {
                const int k = it / n, i = it % n;
                const int p = pq[2 * k], q = pq[2 * k + 1];
                if (p < 0) continue;
                const double c = cs[2 * k], s = cs[2 * k + 1];
                const double aip = A[i * kCoreLd + p], aiq = A[i * kCoreLd + q];
                A[i * kCoreLd + p] = c * aip - s * aiq;
                A[i * kCoreLd + q] = s * aip + c * aiq;
                const double vip = V[i * kCoreLd + p], viq = V[i * kCoreLd + q];
                V[i * kCoreLd + p] = c * vip - s * viq;
                V[i * kCoreLd + q] = s * vip + c * viq;
            }
            ln.sync();
            // ---- A <- J^T A : item = (pair k, column j) -------------------------------------------------
            for (int it = ln.lane; it < npairs * n; it += ln.nl) {
                const int k = it / n, j = it % n;
                const int p = pq[2 * k], q = pq[2 * k + 1];
                if (p < 0) continue;
                const double c = cs[2 * k], s = cs[2 * k + 1];
                const double apj = A[p * kCoreLd + j], aqj = A[q * kCoreLd + j];
                A[p * kCoreLd + j] = c * apj - s * aqj;
                A[q * kCoreLd + j] = s * apj + c * aqj;
            }
            ln.sync();
            // ---- the rotated pivots are zero by construction; keep A exactly symmetric ---------------
            for (int k = ln.lane; k < npairs; k += ln.nl) {
                const int p = pq[2 * k], q = pq[2 * k + 1];
                if (p < 0) continue;
                A[p * kCoreLd + q] = 0.0;
                A[q * kCoreLd + p] = 0.0;
            }
            ln.sync();
        }
        // symmetrise (the two phases round the mirrored entries differently)
        for (int i = ln.lane; i < n; i += ln.nl)
            for (int j = i + 1; j < n; ++j) {
                const double v = 0.5 * (A[i * kCoreLd + j] + A[j * kCoreLd + i]);
                A[i * kCoreLd + j] = v;
            }
        ln.sync();
        for (int i = ln.lane; i < n; i += ln.nl)
            for (int j = 0; j < i; ++j) A[i * kCoreLd + j] = A[j * kCoreLd + i];
        ln.sync();
    }
    ln.sync();
}

// fp16 high block + multi-stage RTVQ low block of the coefficients in out.coef (compress.py:44-51,
// rtvq.py:39-82): writes chigh, codes, qscale / qzp / qres and the round-tripped coefficients chat.
// Reads only info and coef, so it also serves as the re-quantisation step after an exact projection.
template <class L>
SVDQ_HD void quantize_param(const SolveConfig& cfg, const uint32_t present, const SolveOut& out, L& ln) {
    const int NT = cfg.n_tasks, S = cfg.stages;
    if (out.info[0] != kSolved) return;
    const int r = out.info[2], k = out.info[3];
    const int n_low = r - k;
    for (int t = ln.lane; t < NT; t += ln.nl) {
        if (!(present >> t & 1u)) continue;
        float c[kCoreMaxTasks];
        for (int j = 0; j < r; ++j) c[j] = out.coef[t * NT + j];
        for (int j = 0; j < k; ++j) {
            const uint16_t h = f32_to_f16_bits(c[j]);
            out.chigh[t * NT + j] = h;
            out.chat[t * NT + j] = f16_bits_to_f32(h);
        }
        if (n_low > 0) {
            float deq[kCoreMaxTasks];
            rtvq_short(c + k, n_low, cfg.bits, S, out.codes + (size_t)t * S * NT, NT,
                       out.qscale + t * S, out.qzp + t * S, out.qres + t * S, deq);
            for (int i = 0; i < n_low; ++i) out.chat[t * NT + k + i] = deq[i];
        }
    }
    ln.sync();
}

// Weighted average of the round-tripped coefficients in sorted-name order (merge.py:89-139), the
// NaN-propagation term of the zero directions, and the fp32-basis shortcut g = W cbar.  Reads only
// what solve_param wrote (info, chat, W), so it can run as its own tiny launch once the task weights
// are known (cluster weighting computes them on the host while solve_param runs).
template <class L>
SVDQ_HD void average_param(const SolveConfig& cfg, const SolveIn& in, const SolveOut& out, L& ln) {
    const int NT = cfg.n_tasks;
    if (out.info[0] != kSolved) return;
    const int r = out.info[2], r_eff = out.info[4];
    // Without clustering: weights renormalised over the tasks that have the parameter (merge.py:121-123).
    // With clustering (merge_with_clustering, merge.py:586-626): the member weights are renormalised over the
    // members of each cluster that have the parameter, every cluster's reconstruction (U c_c + mean) enters the
    // cross-cluster average with omega[c], and a cluster WITHOUT any such member contributes zeros -- mean
    // included (merge.py:289-290 via dequantize_and_average returning None).  All of it is linear, so it is one
    // average with w_eff[t] = omega[c(t)] w[t] / wsum_c and the mean scaled by sum of omega over the non-empty
    // clusters (= 1 unless a whole cluster lacks the parameter).
    double wsum = 0.0, mean_scale = 1.0;
    double wsum_c[kCoreMaxTasks];
    if (in.cluster_of == nullptr) {
        for (int o = 0; o < NT; ++o) {
            const int t = in.avg_order[o];
            if (in.present >> t & 1u) wsum += in.weights[t];
        }
    } else {
        for (int c = 0; c < NT; ++c) wsum_c[c] = 0.0;
        bool any[kCoreMaxTasks];
        for (int c = 0; c < NT; ++c) any[c] = false;
        for (int o = 0; o < NT; ++o) {
            const int t = in.avg_order[o];
            if (in.present >> t & 1u) { wsum_c[in.cluster_of[t]] += in.weights[t]; any[in.cluster_of[t]] = true; }
        }
        mean_scale = 0.0;
        for (int c = 0; c < NT; ++c) if (any[c]) mean_scale += in.omega[c];
    }
    for (int j = ln.lane; j < r; j += ln.nl) {
        float acc = 0.0f;
        for (int o = 0; o < NT; ++o) {
            const int t = in.avg_order[o];
            if (!(in.present >> t & 1u)) continue;
            const float w = in.cluster_of == nullptr
                                ? (float)(in.weights[t] / wsum)
                                : (float)(in.omega[in.cluster_of[t]] * (in.weights[t] / wsum_c[in.cluster_of[t]]));
            acc = f_add(acc, f_mul(out.chat[t * NT + j], w));
        }
        out.cbar[j] = acc;
    }
    if (ln.lane == 0) out.scal[2] = (float)mean_scale;
    ln.sync();
    if (ln.lane == 0) {
        float tail = 0.0f;
        for (int j = r_eff; j < r; ++j) tail = f_add(tail, f_mul(0.0f, out.cbar[j]));
        out.scal[1] = tail;
    }
    for (int t = ln.lane; t < NT; t += ln.nl) {
        double g = 0.0;
        if (in.present >> t & 1u)
            for (int j = 0; j < r_eff; ++j) g += (double)out.W[t * NT + j] * (double)out.cbar[j];
        out.gvec[t] = (float)g;
    }
    ln.sync();
}

// The whole per-parameter solve.  Every lane must call it with the same arguments.
template <class L>
SVDQ_HD void solve_param(const SolveConfig& cfg, const SolveIn& in, const SolveOut& out, SolveScratch& sc, L& ln) {
    const int NT = cfg.n_tasks, S = cfg.stages;

    // ---- clear outputs ----------------------------------------------------------------------
    for (int i = ln.lane; i < NT * NT; i += ln.nl) {
        out.coef[i] = 0.0f; out.chigh[i] = 0; out.chat[i] = 0.0f; out.W[i] = 0.0f; out.V[i] = 0.0;
    }
    for (int i = ln.lane; i < NT * S * NT; i += ln.nl) out.codes[i] = 0;
    for (int i = ln.lane; i < NT * S; i += ln.nl) { out.qscale[i] = 0.0f; out.qzp[i] = 0.0f; out.qres[i] = 0.0f; }
    for (int i = ln.lane; i < NT; i += ln.nl) { out.sv[i] = 0.0f; out.cbar[i] = 0.0f; out.gvec[i] = 0.0f; }

    // ---- active tasks, gating (cli.py:319-343) ----------------------------------------------
    int n = 0;
    for (int t = 0; t < NT; ++t)
        if (in.present >> t & 1u) { if (ln.lane == 0) sc.idx[n] = t; ++n; }
    int status = kSolved;
    if (n == 0 || in.dm <= 0) status = kSkippedEmpty;
    else if (in.has_mask && in.dm < (int64_t)cfg.min_mask_size) status = kSkippedSmallMask;
    if (status != kSolved) {
        if (ln.lane == 0) {
            out.info[0] = status; out.info[1] = n; out.info[2] = 0; out.info[3] = 0; out.info[4] = 0;
            out.info[5] = 0; out.info[6] = 0; out.info[7] = 0;
            out.scal[0] = 0.0f; out.scal[1] = 0.0f; out.scal[2] = 0.0f; out.scal[3] = 0.0f;
        }
        ln.sync();
        return;
    }
    ln.sync();

    // ---- compact Gram of the active tasks, centre: G_c = H G H --------------------------------
    for (int a = ln.lane; a < n; a += ln.nl)
        for (int b = 0; b < n; ++b) sc.A[a * kCoreLd + b] = in.G[sc.idx[a] * NT + sc.idx[b]];
    ln.sync();
    if (cfg.center) {
        for (int a = ln.lane; a < n; a += ln.nl) {
            double s = 0.0;
            for (int b = 0; b < n; ++b) s += sc.A[a * kCoreLd + b];
            sc.tmp[a] = s / n;
        }
        ln.sync();
        double mm = 0.0;
        for (int a = 0; a < n; ++a) mm += sc.tmp[a];
        mm /= n;
        for (int a = ln.lane; a < n; a += ln.nl)
            for (int b = 0; b < n; ++b) sc.A[a * kCoreLd + b] = sc.A[a * kCoreLd + b] - sc.tmp[a] - sc.tmp[b] + mm;
        ln.sync();
        // exact symmetry (the row/column means were summed in different orders)
        for (int a = ln.lane; a < n; a += ln.nl)
            for (int b = a + 1; b < n; ++b) {
                const double v = 0.5 * (sc.A[a * kCoreLd + b] + sc.A[b * kCoreLd + a]);
                sc.A[a * kCoreLd + b] = v;
            }
        ln.sync();
        for (int a = ln.lane; a < n; a += ln.nl)
            for (int b = 0; b < a; ++b) sc.A[a * kCoreLd + b] = sc.A[b * kCoreLd + a];
        ln.sync();
    }

    jacobi_eig(sc.A, sc.V, sc.cs, sc.pq, n, ln);

    // ---- sort descending, sign convention -------------------------------------------------------
    if (ln.lane == 0) {
        for (int i = 0; i < n; ++i) sc.perm[i] = i;
        for (int i = 0; i < n; ++i) {              // stable selection sort on the diagonal
            int best = i;
            for (int j = i + 1; j < n; ++j)
                if (sc.A[sc.perm[j] * kCoreLd + sc.perm[j]] > sc.A[sc.perm[best] * kCoreLd + sc.perm[best]]) best = j;
            const int pb = sc.perm[best];
            for (int j = best; j > i; --j) sc.perm[j] = sc.perm[j - 1];
            sc.perm[i] = pb;
        }
    }
    ln.sync();
    for (int j = ln.lane; j < n; j += ln.nl) {
        const int src = sc.perm[j];
        const double lam = sc.A[src * kCoreLd + src];
        // sqrt(|lambda|): a numerically-zero eigenvalue of the centred Gram comes out as +-1e-17 lambda_1;
        // LAPACK's singular value there is non-negative round-off dust, never an exact zero, and the
        // reference's 2-element RTVQ edge (rtvq.py:17) turns NaN only on an exact zero
        sc.sigma[j] = sqrt(fabs(lam));
        double sgn = 1.0;
        if (in.sign_ref) {
            double dot = 0.0;
            for (int a = 0; a < n; ++a) dot += sc.V[a * kCoreLd + src] * in.sign_ref[j * NT + sc.idx[a]];
            sgn = dot < 0.0 ? -1.0 : 1.0;
        } else {
            double big = 0.0;
            for (int a = 0; a < n; ++a) {
                const double v = sc.V[a * kCoreLd + src];
                if (fabs(v) > fabs(big)) big = v;
            }
            sgn = big < 0.0 ? -1.0 : 1.0;
        }
        for (int a = 0; a < n; ++a) sc.Vs[a * kCoreLd + j] = sgn * sc.V[a * kCoreLd + src];
    }
    ln.sync();

    // ---- thin-SVD shape, rank selection -----------------------------------------------------------
    const int r = in.dm < (int64_t)n ? (int)in.dm : n;       // torch.linalg.svd(full_matrices=False)
    if (ln.lane == 0) {
        float S32[kCoreMaxTasks];
        for (int j = 0; j < r; ++j) S32[j] = (float)sc.sigma[j];
        float er;
        const int k = select_rank_f32(S32, r, cfg.energy_threshold, cfg.max_rank, 1, &er);
        // columns whose singular value is numerically zero carry no direction: they get a zero
        // basis column (the reference's LAPACK vector there is round-off noise; its coefficient
        // dequantises to ~0 either way)
        int r_eff = 0;
        for (int j = 0; j < r; ++j)
            if (sc.sigma[j] > 1e-5 * sc.sigma[0] && sc.sigma[j] > 0.0) r_eff = j + 1;
        sc.n = n; sc.r = r; sc.k = k; sc.r_eff = r_eff;
        for (int j = 0; j < r; ++j) out.sv[j] = S32[j];
        out.scal[0] = er; out.scal[1] = 0.0f; out.scal[2] = 0.0f; out.scal[3] = 0.0f;
        out.info[0] = kSolved; out.info[1] = n; out.info[2] = r; out.info[3] = k; out.info[4] = r_eff;
        out.info[5] = 0; out.info[6] = 0; out.info[7] = 0;
    }
    ln.sync();
    const int k = sc.k, r_eff = sc.r_eff;
    const int n_low = r - k;

    // ---- per task: closed-form coefficients c[t][j] = sigma_j V[t][j] ---------------------------------
    for (int a = ln.lane; a < n; a += ln.nl) {
        const int t = sc.idx[a];
        for (int j = 0; j < r; ++j) out.coef[t * NT + j] = (float)(sc.sigma[j] * sc.Vs[a * kCoreLd + j]);
    }
    ln.sync();
    quantize_param(cfg, in.present, out, ln);

    // ---- projection matrix W = H V Sigma^-1 (independent of the task weights) ----------------------
    for (int a = ln.lane; a < n; a += ln.nl) {
        const int t = sc.idx[a];
        for (int j = 0; j < r_eff; ++j) {
            double v = sc.Vs[a * kCoreLd + j];
            if (cfg.center) {
                double m = 0.0;
                for (int b = 0; b < n; ++b) m += sc.Vs[b * kCoreLd + j];
                v -= m / n;
            }
            out.W[t * NT + j] = (float)(v / sc.sigma[j]);
        }
        for (int j = 0; j < n; ++j) out.V[t * NT + j] = sc.Vs[a * kCoreLd + j];
    }
    ln.sync();
    if (in.weights != nullptr) average_param(cfg, in, out, ln);
}


}  // namespace svdq
