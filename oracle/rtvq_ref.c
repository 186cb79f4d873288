/*
 * Plain-C restatement of the byte/integer pieces of the reference hot path.
 * TEST ORACLE ONLY (see oracle/__init__.py): only tests/, smoke() and the
 * cpu_baseline leg of bench.py may load this library.
 *
 * Build: oracle/build.py  (gcc -O2 -ffp-contract=off; no fast-math, so every
 * fp32 multiply, add and divide is a separately rounded IEEE operation, which
 * is what torch-eager on CPU does).
 *
 * Restated from (paths relative to /root/reference):
 *   ref_combine_masks   src/svd_hybrid/mask_loader.py:412-485
 *   ref_asym_quant      src/svd_hybrid/rtvq.py:4-27  (= quantization_utils.py:76-99)
 *   ref_asym_dequant    src/svd_hybrid/rtvq.py:29-36 (= quantization_utils.py:137-172)
 *   ref_rtvq            src/svd_hybrid/rtvq.py:39-103
 *   ref_select_rank     src/svd_hybrid/basis.py:116-213
 *   ref_absmax_quant    quantization_utils.py:60-73
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* strategy: 0 union, 1 intersection, 2 majority (votes >= 0.5 * n_masks) */
int ref_combine_masks(const uint8_t *const *masks, int n_masks, int64_t n, int strategy, uint8_t *out)
{
    if (n_masks <= 0) return -1;
    if (strategy < 0 || strategy > 2) return -2;
    for (int64_t i = 0; i < n; ++i) {
        int votes = 0;
        for (int t = 0; t < n_masks; ++t) votes += masks[t][i] != 0;
        uint8_t v;
        if (strategy == 0)      v = votes > 0;
        else if (strategy == 1) v = votes == n_masks;
        else                    v = (float)votes >= 0.5f * (float)n_masks;
        out[i] = v;
    }
    return 0;
}

/* torch.min / torch.max propagate NaN */
static void minmax_nanprop(const float *x, int64_t n, float *lo, float *hi)
{
    float a = x[0], b = x[0];
    int nan = isnan(x[0]);
    for (int64_t i = 1; i < n; ++i) {
        float v = x[i];
        if (isnan(v)) nan = 1;
        if (v < a) a = v;
        if (v > b) b = v;
    }
    if (nan) { a = NAN; b = NAN; }
    *lo = a; *hi = b;
}

static inline float clampf_nanprop(float v, float lo, float hi)
{
    if (isnan(v)) return v;
    return v < lo ? lo : (v > hi ? hi : v);
}

/* float -> integer code the way torch's .to(uint8/int16) lands on x86: NaN -> 0 */
static inline int32_t to_code(float v)
{
    if (isnan(v)) return 0;
    return (int32_t)v;
}

/* codes are written as int32 per element so one entry point serves uint8 and int16 */
int ref_asym_quant(const float *x, int64_t n, int bits, int32_t *q, float *scale_out, float *zp_out)
{
    if (n <= 0) return -1;
    float lo, hi;
    minmax_nanprop(x, n, &lo, &hi);
    const float qmax = (float)((1 << bits) - 1);
    /* Python `int / tensor` dispatches to Tensor.__rtruediv__ = reciprocal() * int:
       two roundings (1/(hi-lo), then * qmax), not one division. */
    const float recip = 1.0f / (hi - lo);
    const float scale = recip * qmax;
    const float zp = -1.0f * rintf(scale * lo);
    for (int64_t i = 0; i < n; ++i) {
        float t = scale * x[i];
        t = t + zp;
        q[i] = to_code(clampf_nanprop(rintf(t), 0.0f, qmax));
    }
    *scale_out = scale;
    *zp_out = zp;
    return 0;
}

void ref_asym_dequant(const int32_t *q, int64_t n, float scale, float zp, float *out)
{
    for (int64_t i = 0; i < n; ++i) {
        float t = (float)q[i] - zp;
        out[i] = t / scale;
    }
}

/*
 * Multi-stage residual quantisation.  codes: [stages][n] int32, scale/zp/resnorm: [stages],
 * deq_sum: [n] = left-to-right sum of the stage dequantisations starting from 0 (rtvq.py:91-101),
 * may be NULL.  Returns the number of payloads (0 for an empty tensor).
 */
int ref_rtvq(const float *x, int64_t n, int bits, int stages, int32_t *codes, float *scale, float *zp,
             float *resnorm, float *deq_sum)
{
    if (n == 0) return 0;
    float *res = (float *)malloc(sizeof(float) * (size_t)n);
    float *deq = (float *)malloc(sizeof(float) * (size_t)n);
    if (!res || !deq) { free(res); free(deq); return -1; }
    memcpy(res, x, sizeof(float) * (size_t)n);
    if (deq_sum) for (int64_t i = 0; i < n; ++i) deq_sum[i] = 0.0f;
    for (int s = 0; s < stages; ++s) {
        double ss = 0.0;
        for (int64_t i = 0; i < n; ++i) ss += (double)res[i] * (double)res[i];
        resnorm[s] = (float)sqrt(ss);
        ref_asym_quant(res, n, bits, codes + (size_t)s * (size_t)n, &scale[s], &zp[s]);
        ref_asym_dequant(codes + (size_t)s * (size_t)n, n, scale[s], zp[s], deq);
        for (int64_t i = 0; i < n; ++i) {
            res[i] = res[i] - deq[i];
            if (deq_sum) deq_sum[i] = deq_sum[i] + deq[i];
        }
    }
    free(res); free(deq);
    return stages;
}

/*
 * Energy rank selection.  S: r singular values (fp32).  cum_out (may be NULL) receives the
 * cumulative energy fractions.  max_rank <= 0 means "no cap".
 * torch details restated: S**2 in fp32; cumsum accumulates in double and rounds each prefix to
 * fp32 (ATen cpu_cum_base_kernel uses acc_type<float,false> = double); division in fp32; the
 * threshold is compared after rounding it to fp32.  The total is a fp32 reduction whose lane
 * order is an ATen implementation detail: it is restated as the double sum rounded to fp32,
 * which differs from torch by at most 1 ulp (ties within 1 ulp of the threshold are the only
 * inputs that can see it).
 */
int ref_select_rank(const float *S, int r, float thr, int max_rank, int min_rank, float *cum_out)
{
    if (r <= 0) return 0;
    double tot_d = 0.0;
    for (int i = 0; i < r; ++i) tot_d += (double)(S[i] * S[i]);
    const float tot = (float)tot_d;
    int below = 0;
    double acc = 0.0;
    for (int i = 0; i < r; ++i) {
        float c;
        if (tot < 1e-10f) c = 1.0f;
        else { acc += (double)(S[i] * S[i]); c = (float)acc / tot; }
        if (cum_out) cum_out[i] = c;
        if (c < thr) ++below;
    }
    int k = below + 1;
    if (k < min_rank) k = min_rank;
    if (max_rank > 0 && k > max_rank) k = max_rank;
    if (k > r) k = r;
    return k;
}

/* absmax quantiser of the root module: s = (2^(b-1)-1)/max|x|; q = round(s*x); no clamp */
int ref_absmax_quant(const float *x, int64_t n, int bits, int32_t *q, float *scale_out)
{
    if (n <= 0) return -1;
    float m = fabsf(x[0]);
    int nan = isnan(x[0]);
    for (int64_t i = 1; i < n; ++i) {
        float a = fabsf(x[i]);
        if (isnan(a)) nan = 1;
        if (a > m) m = a;
    }
    if (nan) m = NAN;
    const float recip = 1.0f / m;                       /* same __rtruediv__ path */
    const float s = recip * (float)((1 << (bits - 1)) - 1);
    for (int64_t i = 0; i < n; ++i) q[i] = to_code(rintf(s * x[i]));
    *scale_out = s;
    return 0;
}
