// Host-side k-means for cluster weighting: a lean restatement of the call the reference makes,
//   sklearn.cluster.KMeans(n_clusters=k, random_state=42, n_init=10).fit_predict(features)
// (reference: src/svd_hybrid/clustering.py:123-156, called from cluster_tasks :198-245 and cli.py:530).
//
// The reference's feature matrix is [N x P_total]; k-means only sees pairwise geometry, so the engine
// feeds an isometric N x N embedding of the whole-model task Gram (K1 by-product).  What is reproduced
// here is scikit-learn 1.9's procedure, step by step, so that the labels come out as sklearn's do:
//   * numpy RandomState(seed): MT19937 with init_genrand seeding; choice(p=...) = searchsorted(cdf, double, right);
//     uniform() = 53-bit doubles (two 32-bit draws each);
//   * fit(): X -= X.mean(0) in float32; tol = mean(var(X, 0)) * 1e-4;
//   * n_init runs from ONE random stream: k-means++ seeding with 2 + int(log k) local trials (distances in
//     float64, rounded to float32; float32 cumulative sums), then Lloyd iterations in float32
//     (argmin of ||c||^2 - 2 x.c, first minimum wins; empty clusters relocated to the farthest points;
//     strict-convergence / centre-shift stopping rules; final E-step when not strictly converged);
//   * the first run wins ties: a later run replaces the best only with a lower float32 inertia AND a different partition.
// No threads, no allocation beyond a few small vectors: ~20 us for 8 tasks (sklearn's call: 8-23 ms).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {

struct MT19937 {
    uint32_t mt[624];
    int idx;
    explicit MT19937(uint32_t seed) {
        mt[0] = seed;
        for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
        idx = 624;
    }
    uint32_t next32() {
        if (idx >= 624) {
            for (int i = 0; i < 624; ++i) {
                const uint32_t y = (mt[i] & 0x80000000u) | (mt[(i + 1) % 624] & 0x7fffffffu);
                mt[i] = mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            idx = 0;
        }
        uint32_t y = mt[idx++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    }
    double next_double() {                       // numpy legacy random_sample
        const uint32_t a = next32() >> 5, b = next32() >> 6;
        return (a * 67108864.0 + b) / 9007199254740992.0;
    }
};

struct Problem {
    int n, d, k;
    const float* X;                 // centred, row-major [n x d]
};

// squared distances of row c to every row, float64 arithmetic rounded to float32 and clamped at 0
// (sklearn.metrics.pairwise._euclidean_distances_upcast for float32 inputs)
void dist_row(const Problem& P, int c, float* out) {
    const float* xc = P.X + (size_t)c * P.d;
    double cc = 0.0;
    for (int f = 0; f < P.d; ++f) cc += (double)xc[f] * xc[f];
    for (int i = 0; i < P.n; ++i) {
        const float* xi = P.X + (size_t)i * P.d;
        double ii = 0.0, ci = 0.0;
        for (int f = 0; f < P.d; ++f) { ii += (double)xi[f] * xi[f]; ci += (double)xc[f] * xi[f]; }
        float v = (float)(-2.0 * ci + cc + ii);
        out[i] = v > 0.f ? v : 0.f;
    }
    out[c] = out[c];                // (X[candidates] is a copy, not `X is Y`: the diagonal is not forced to 0)
}

void kmeans_plusplus(const Problem& P, MT19937& rng, float* centers) {
    const int n = P.n, d = P.d, k = P.k;
    const int n_trials = 2 + (int)std::log((double)k);
    // first centre: RandomState.choice(n, p = 1/n as float32)
    const float pf = 1.0f / (float)n;           // sample_weight / sample_weight.sum() in float32
    std::vector<double> cdf(n);
    double acc = 0.0;
    for (int i = 0; i < n; ++i) { acc += (double)pf; cdf[i] = acc; }
    for (int i = 0; i < n; ++i) cdf[i] /= acc;
    const double u = rng.next_double();
    int first = 0;
    while (first < n && cdf[first] <= u) ++first;                 // searchsorted(side="right")
    if (first >= n) first = n - 1;
    std::memcpy(centers, P.X + (size_t)first * d, sizeof(float) * d);
    std::vector<float> closest(n), cum(n), cand_d((size_t)n_trials * n);
    dist_row(P, first, closest.data());
    float pot = 0.f;
    for (int i = 0; i < n; ++i) pot += closest[i];
    std::vector<double> rv(n_trials);
    std::vector<int> cand(n_trials);
    for (int c = 1; c < k; ++c) {
        for (int t = 0; t < n_trials; ++t) rv[t] = rng.next_double() * (double)pot;
        float s = 0.f;
        for (int i = 0; i < n; ++i) { s += closest[i]; cum[i] = s; }
        for (int t = 0; t < n_trials; ++t) {
            int j = 0;
            while (j < n && (double)cum[j] < rv[t]) ++j;          // searchsorted(side="left")
            cand[t] = j < n - 1 ? j : n - 1;
        }
        int best = 0;
        float best_pot = 0.f;
        for (int t = 0; t < n_trials; ++t) {
            float* dt = cand_d.data() + (size_t)t * n;
            dist_row(P, cand[t], dt);
            float p = 0.f;
            for (int i = 0; i < n; ++i) { dt[i] = closest[i] < dt[i] ? closest[i] : dt[i]; p += dt[i]; }
            if (t == 0 || p < best_pot) { best = t; best_pot = p; }
        }
        pot = best_pot;
        std::memcpy(closest.data(), cand_d.data() + (size_t)best * n, sizeof(float) * n);
        std::memcpy(centers + (size_t)c * d, P.X + (size_t)cand[best] * d, sizeof(float) * d);
    }
}

float sqdist(const float* a, const float* b, int d) {             // _euclidean_dense_dense (4-way unrolled sum)
    float r = 0.f;
    const int n4 = d / 4, rem = d % 4;
    for (int i = 0; i < n4; ++i, a += 4, b += 4)
        r += (a[0] - b[0]) * (a[0] - b[0]) + (a[1] - b[1]) * (a[1] - b[1]) + (a[2] - b[2]) * (a[2] - b[2]) +
             (a[3] - b[3]) * (a[3] - b[3]);
    for (int i = 0; i < rem; ++i) r += (a[i] - b[i]) * (a[i] - b[i]);
    return r;
}

// one Lloyd iteration (lloyd_iter_chunked_dense); update = false: E-step only
void lloyd_iter(const Problem& P, const float* c_old, float* c_new, float* w, int32_t* labels, float* shift, bool update) {
    const int n = P.n, d = P.d, k = P.k;
    std::vector<float> cn(k);
    for (int j = 0; j < k; ++j) {
        float s = 0.f;
        for (int f = 0; f < d; ++f) s += c_old[(size_t)j * d + f] * c_old[(size_t)j * d + f];
        cn[j] = s;
    }
    if (update) {
        std::memset(c_new, 0, sizeof(float) * (size_t)k * d);
        std::memset(w, 0, sizeof(float) * k);
    }
    for (int i = 0; i < n; ++i) {
        const float* xi = P.X + (size_t)i * d;
        int lab = 0;
        float best = 0.f;
        for (int j = 0; j < k; ++j) {
            float dot = 0.f;
            for (int f = 0; f < d; ++f) dot += xi[f] * c_old[(size_t)j * d + f];
            const float v = -2.0f * dot + cn[j];
            if (j == 0 || v < best) { best = v; lab = j; }
        }
        labels[i] = lab;
        if (update) {
            w[lab] += 1.0f;
            for (int f = 0; f < d; ++f) c_new[(size_t)lab * d + f] += xi[f];
        }
    }
    if (!update) return;
    // _relocate_empty_clusters_dense
    std::vector<int> empty;
    for (int j = 0; j < k; ++j) if (w[j] == 0.f) empty.push_back(j);
    if (!empty.empty()) {
        std::vector<float> dist(n);
        float mx = 0.f;
        for (int i = 0; i < n; ++i) {
            float s = 0.f;
            for (int f = 0; f < d; ++f) { const float t = P.X[(size_t)i * d + f] - c_old[(size_t)labels[i] * d + f]; s += t * t; }
            dist[i] = s;
            mx = s > mx ? s : mx;
        }
        if (mx != 0.f) {
            std::vector<char> used(n, 0);
            for (size_t e = 0; e < empty.size(); ++e) {           // farthest points first
                int far = -1;
                for (int i = 0; i < n; ++i) if (!used[i] && (far < 0 || dist[i] > dist[far])) far = i;
                if (far < 0) break;
                used[far] = 1;
                const int nc = empty[e], oc = labels[far];
                for (int f = 0; f < d; ++f) {
                    c_new[(size_t)oc * d + f] -= P.X[(size_t)far * d + f];
                    c_new[(size_t)nc * d + f] = P.X[(size_t)far * d + f];
                }
                w[nc] = 1.0f;
                w[oc] -= 1.0f;
            }
        }
    }
    // _average_centers
    int amax = 0;
    for (int j = 1; j < k; ++j) if (w[j] > w[amax]) amax = j;
    for (int j = 0; j < k; ++j) {
        if (w[j] > 0.f) {
            const float alpha = 1.0f / w[j];
            for (int f = 0; f < d; ++f) c_new[(size_t)j * d + f] *= alpha;
        } else {
            for (int f = 0; f < d; ++f) c_new[(size_t)j * d + f] = c_new[(size_t)amax * d + f];
        }
    }
    for (int j = 0; j < k; ++j) shift[j] = std::sqrt(sqdist(c_new + (size_t)j * d, c_old + (size_t)j * d, d));
}

bool same_clustering(const int32_t* a, const int32_t* b, int n, int k) {
    std::vector<int> map(k, -1);
    for (int i = 0; i < n; ++i) {
        if (map[a[i]] == -1) map[a[i]] = b[i];
        else if (map[a[i]] != b[i]) return false;
    }
    return true;
}

}  // namespace

extern "C" int svdq_host_kmeans_impl(const float* features, int n, int d, int k, uint32_t seed, int n_init, int max_iter,
                                     double tol_rel, int32_t* labels_out, double* inertia_out) {
    // centre (float32, like X -= X.mean(axis=0)) and tolerance (np.var on the un-centred data)
    std::vector<float> X((size_t)n * d);
    double var_mean = 0.0;
    for (int f = 0; f < d; ++f) {
        float s = 0.f;
        for (int i = 0; i < n; ++i) s += features[(size_t)i * d + f];
        const float mean = s / (float)n;
        float v = 0.f;
        for (int i = 0; i < n; ++i) {
            const float t = features[(size_t)i * d + f] - mean;
            v += t * t;
            X[(size_t)i * d + f] = t;
        }
        var_mean += (double)(v / (float)n);
    }
    const float tol = (float)(var_mean / d) * (float)tol_rel;
    Problem P{n, d, k, X.data()};
    MT19937 rng(seed);
    std::vector<float> ca((size_t)k * d), cb((size_t)k * d), w(k), shift(k);
    std::vector<int32_t> labels(n), labels_old(n), best(n);
    float best_inertia = 0.f;
    bool have = false;
    for (int run = 0; run < n_init; ++run) {
        kmeans_plusplus(P, rng, ca.data());
        float* centers = ca.data();
        float* centers_new = cb.data();
        std::fill(labels.begin(), labels.end(), -1);
        std::fill(labels_old.begin(), labels_old.end(), -1);
        bool strict = false;
        for (int it = 0; it < max_iter; ++it) {
            lloyd_iter(P, centers, centers_new, w.data(), labels.data(), shift.data(), true);
            float* t = centers; centers = centers_new; centers_new = t;
            if (labels == labels_old) { strict = true; break; }
            float tot = 0.f;
            for (int j = 0; j < k; ++j) tot += shift[j] * shift[j];
            if (tot <= tol) break;
            labels_old = labels;
        }
        if (!strict) lloyd_iter(P, centers, centers, w.data(), labels.data(), shift.data(), false);
        float inertia = 0.f;
        for (int i = 0; i < n; ++i) inertia += sqdist(P.X + (size_t)i * d, centers + (size_t)labels[i] * d, d);
        if (!have || (inertia < best_inertia && !same_clustering(labels.data(), best.data(), n, k))) {
            best = labels;
            best_inertia = inertia;
            have = true;
        }
    }
    std::memcpy(labels_out, best.data(), sizeof(int32_t) * n);
    if (inertia_out) *inertia_out = (double)best_inertia;
    return 0;
}
