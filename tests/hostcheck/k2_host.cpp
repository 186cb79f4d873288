// TEST-ONLY host build of the per-parameter solve core (csrc/k2_core.h) so that the exact source
// the GPU runs can be checked against the oracle on a machine without a GPU.  Never loaded by the
// product package; built by tests/hostcheck/build.py with g++ -ffp-contract=off.
#include "../../svd_quantization_task_merging_b200/csrc/k2_core.h"

using namespace svdq;

struct SerialLanes {
    int lane = 0;
    static constexpr int nl = 1;
    void sync() {}
};

extern "C" {

int k2_host_solve(int n_tasks, int center, float thr, int max_rank, int min_mask_size, int bits, int stages,
                  const double* G, int64_t dm, int has_mask, uint32_t present, const double* weights,
                  const int32_t* avg_order, const double* sign_ref,
                  int32_t* info, float* sv, float* scal, float* coef, uint16_t* chigh, uint8_t* codes, float* qscale,
                  float* qzp, float* qres, float* chat, float* cbar, float* W, float* gvec, double* V) {
    static SolveScratch sc;
    SolveConfig cfg{n_tasks, center, thr, max_rank, min_mask_size, bits, stages};
    SolveIn in{G, dm, has_mask, present, weights, avg_order, sign_ref};
    SolveOut out{info, sv, scal, coef, chigh, codes, qscale, qzp, qres, chat, cbar, W, gvec, V};
    SerialLanes ln;
    solve_param(cfg, in, out, sc, ln);
    return 0;
}

uint16_t k2_host_f32_to_f16(float f) { return f32_to_f16_bits(f); }
float k2_host_f16_to_f32(uint16_t h) { return f16_bits_to_f32(h); }

int k2_host_select_rank(const float* S, int r, float thr, int max_rank, int min_rank, float* er) {
    return select_rank_f32(S, r, thr, max_rank, min_rank, er);
}

void k2_host_rtvq_short(const float* x, int n, int bits, int stages, uint8_t* codes, int ld, float* scale, float* zp,
                        float* resnorm, float* deq) {
    rtvq_short(x, n, bits, stages, codes, ld, scale, zp, resnorm, deq);
}

}  // extern "C"
