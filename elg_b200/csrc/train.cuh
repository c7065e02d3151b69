// Training path (REINFORCE with the POMO shared baseline): shared declarations.
// reference: CVRP/train.py:83-148, TSP/train.py:80-146.
#pragma once
#include "common.cuh"

namespace elg {

// ---- generic strided / batched fp32 GEMM  C (+)= alpha * A * B  (train_bwd.cu) -------------------------------
// A(m, k) = A[m * sAm + k * sAk], B(k, n) = B[k * sBk + n * sBn] (B == NULL: all ones), C[m * ldc + n].
// batch index = b1 * nb2 + b2 with separate strides per level; splits > 1 splits K over CTAs and adds with atomics
// (C must then hold the value to accumulate onto).
struct GemmP {
  const float* A = nullptr;
  const float* B = nullptr;
  float* C = nullptr;
  int M = 0, N = 0, K = 0;
  long long sAm = 0, sAk = 0, sBk = 0, sBn = 0;
  int ldc = 0;
  int nb1 = 1, nb2 = 1;
  long long bA1 = 0, bA2 = 0, bB1 = 0, bB2 = 0, bC1 = 0, bC2 = 0;
  float alpha = 1.f;
  int accumulate = 0;
  int splits = 1;
};
int launch_gemm(const GemmP& p, cudaStream_t st);

// ---- one recorded decode step of one POMO row (written by replay_kernel) ---------------------------------------
struct StepRec {
  int cur;            // node the row stands on before the step
  int act;            // recorded action of the step
  float load;         // cvrp: remaining load; tsp: first node (as int bits)
  int active;         // 1 = the step carries gradient (policy step, row not finished, >= 2 selectable nodes)
  uint32_t mask[4];   // bit j set = node j masked
};

constexpr int TRAIN_MAX_NODES = 112;     // K'/V/E' of one instance are staged in shared memory by the backward kernel
// local-policy gradient accumulators (floats) filled by local_kernel<BWD>, consumed by local_fold_bwd_kernel
constexpr int LG_WO = 0;                   // [LE][LE]  d multi_head_combine.weight
constexpr int LG_BO = LG_WO + LE * LE;     // [LE]
constexpr int LG_WV = LG_BO + LE;          // [LE][LE]  d Wv
constexpr int LG_WE = LG_WV + LE * LE;     // [LE][4]   direct part of d init_emb.weight
constexpr int LG_BE = LG_WE + LE * 4;      // [LE]      direct part of d init_emb.bias
constexpr int LG_DU = LG_BE + LE;          // [LH][4]   d u_h[f]   (s_hp = u_h . f_p + t_hp)
constexpr int LG_DT = LG_DU + LH * 4;      // [LH][KT_MAX]  d t_hp
constexpr int LG_TOTAL = LG_DT + LH * KT_MAX;

struct DecodeBwdArgs {
  elg_tables t;
  const float* weights;
  elg_weight_layout_t L;
  const float* derived;
  const float* eplain;        // [B][N1][E] fp32 E'
  int problem, B, M, N1, NP, k_local, flags;
  float xi, clip;
  const StepRec* rec;         // [B][T][M]
  int T, t0, nT;              // steps [t0, t0 + nT) of this chunk
  const float* coef;          // [B][M]
  // materialised per row-step (row = (b * nT + (t - t0)) * M + m)
  float* add;                 // [rows][NP]      penalty + local score per node
  float* dx;                  // [rows][NP]      d loss / d pre-tanh score
  // accumulators
  float* dEp;                 // [B][N1][E]  (w.r.t. E' = enc Wo-fold / sqrt(E))
  float* dV;                  // [B][N1][E]
  float* dK;                  // [B][N1][E]  (w.r.t. K' = K log2(e)/sqrt(D))
  float* deb;                 // [B][N1]
  float* dqtab;               // [B][N1][E]
  float* dqfirst;             // [B][N1][E] (tsp)
  float* dwl;                 // [E]  d load column of Wq_last (cvrp)
  float* lg;                  // [LG_TOTAL]
};
int launch_replay(int problem, const float* demand, const int16_t* tours, int t_max, int B, int M, int N1, int T,
                  StepRec* rec, cudaStream_t st);
int launch_local(const DecodeBwdArgs& a, bool bwd, cudaStream_t st);
int launch_global_bwd(const DecodeBwdArgs& a, cudaStream_t st);
int launch_local_fold_bwd(const elg_model_desc* d, const elg_weight_layout_t& L, const float* weights, const float* derived,
                          const float* lg, float* grads, cudaStream_t st);

}  // namespace elg
