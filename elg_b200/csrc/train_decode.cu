// Backward of the decode step (training path): replay of the recorded tours, forward recompute and gradient of
// one decode step per POMO row, for every step of the rollout.
//
// reference (what is differentiated): CVRP_Decoder.forward CVRP/models.py:322-423, local_policy_att.forward
// CVRP/models.py:51-175 (TSP/models.py:244-303, 48-110), driven by J.backward() of CVRP/train.py:112-124.
//
// Row-steps are independent, so the work is split into kernels without cross-row communication that write small
// per-row-step vectors to HBM, and batched GEMMs (train_bwd.cu) that contract them over the rows of an instance:
//   replay_kernel        tours -> per (instance, step, row) state {cur, load, mask bits, action}
//   local_kernel<FWD>    penalty + local-policy score per node            -> ADD[row][node]
//   global_bwd_kernel    q, attention, score, softmax; d logits -> DX, O (for d E'), and in registers d V, d K';
//                        query-table gradient by atomics
//   local_kernel<BWD>    gradient of the local policy (register accumulators, flushed once per warp)
//   local_fold_bwd       chain rule through the constant-query folds -> gradients of the local policy's parameters
#include "train.cuh"
#include "umma.cuh"

namespace elg {

constexpr unsigned FULLM = 0xffffffffu;
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULLM, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULLM, v, o));
  return v;
}

// All-reduce of N (4 or 16) values at once: reduce-scatter over the high lane bits (N/2 + N/4 + ... shuffles), butterfly
// over the remaining bits, all-gather (N shuffles): 32 shuffles for 16 values instead of 80, 10 for 4 instead of 20.
template <int N, bool MAX>
__device__ __forceinline__ void warp_allreduce(float (&v)[N]) {
  const int lane = threadIdx.x & 31;
  constexpr int LOG = N == 16 ? 4 : 2;
  static_assert(N == 16 || N == 4, "N");
#pragma unroll
  for (int s = 0; s < LOG; ++s) {
    const int half = N >> (s + 1), bit = 16 >> s;
    const bool up = (lane & bit) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float send = up ? v[i] : v[i + half];
      const float keep = up ? v[i + half] : v[i];
      const float got = __shfl_xor_sync(FULLM, send, bit);
      v[i] = MAX ? fmaxf(keep, got) : keep + got;
    }
  }
  float r = v[0];
#pragma unroll
  for (int bit = 16 >> LOG; bit > 0; bit >>= 1) {
    const float got = __shfl_xor_sync(FULLM, r, bit);
    r = MAX ? fmaxf(r, got) : r + got;
  }
#pragma unroll
  for (int k = 0; k < N; ++k) v[k] = __shfl_sync(FULLM, r, k << (5 - LOG));
}

// ---- replay: the environment (CVRPEnv.step CVRP/CVRPEnv.py:190-249, TSPEnv.step TSP/TSPEnv.py:108-133) driven by the
// recorded actions; one warp per POMO row.  Same fp32 load recurrence and masks as phase C of the rollout kernels.
__global__ void __launch_bounds__(128) replay_kernel(int problem, const float* __restrict__ demand, const int16_t* __restrict__ tours,
                                                     int t_max, int B, int M, int N1, int T, StepRec* __restrict__ rec) {
  // one warp per POMO row; lane i owns node 32 w + i of every mask word (the too-large test is a ballot)
  const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (g >= B * M) return;
  const int b = g / M, m = g % M;
  const bool cvrp = problem == ELG_CVRP;
  const float* dem = cvrp ? demand + (size_t)b * N1 : nullptr;
  const int16_t* tour = tours + (size_t)g * t_max;
  const int W = (N1 + 31) / 32;
  float dj[4];
#pragma unroll
  for (int w = 0; w < 4; ++w) dj[w] = (cvrp && w * 32 + lane < N1) ? dem[w * 32 + lane] : -1.f;      // -1: never too large
  uint32_t vis[4] = {0, 0, 0, 0}, msk[4] = {0, 0, 0, 0};
  float load = 1.f;
  bool fin = false;
  int cur = 0, first = 0;
  const int tpol = cvrp ? 2 : 1;
  for (int t = 0; t < T; ++t) {
    const int act = t < t_max ? (int)tour[t] : 0;
    if (lane == 0) {
      StepRec r;
      r.cur = cur;
      r.act = act;
      r.load = cvrp ? load : __int_as_float(first);
      int open = 0;
      for (int w = 0; w < W; ++w) {
        const int nb = N1 - w * 32;
        const uint32_t fullw = nb >= 32 ? FULLM : ((1u << nb) - 1u);
        open += __popc(~msk[w] & fullw);
      }
      r.active = (t >= tpol && !fin && open >= 2) ? 1 : 0;
#pragma unroll
      for (int w = 0; w < 4; ++w) r.mask[w] = msk[w];
      rec[((size_t)b * T + t) * M + m] = r;
    }
    // environment step (every lane keeps the same scalar state)
    if (cvrp) {
      const bool at_depot = act == 0;
      load = at_depot ? 1.f : load - dem[act];
      vis[act >> 5] |= 1u << (act & 31);
      vis[0] = at_depot ? (vis[0] | 1u) : (vis[0] & ~1u);
      bool allv = true;
      const float lde = __fadd_rn(load, 1e-6f);
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        if (w < W) {
          const uint32_t big = __ballot_sync(FULLM, lde < dj[w]);
          const int nb = N1 - w * 32;
          const uint32_t fullw = nb >= 32 ? FULLM : ((1u << nb) - 1u);
          allv = allv && ((vis[w] & fullw) == fullw);
          msk[w] = vis[w] | big;
        }
      }
      fin = fin || allv;
      if (fin) msk[0] &= ~1u;
    } else {
      if (t == 0) first = act;
      vis[act >> 5] |= 1u << (act & 31);
      msk[act >> 5] = vis[act >> 5];
    }
    cur = act;
  }
}

int launch_replay(int problem, const float* demand, const int16_t* tours, int t_max, int B, int M, int N1, int T,
                  StepRec* rec, cudaStream_t st) {
  replay_kernel<<<(B * M + 3) / 4, 128, 0, st>>>(problem, demand, tours, t_max, B, M, N1, T, rec);
  ELG_LAUNCH_OK();
  return ELG_OK;
}

// =================================================================================================================
// Local policy, one warp per row-step.  Position p of the local sequence lives on lane p (and p + 32); channel c of
// the 32-wide local embedding lives on lane c; small per-warp shared-memory vectors move data between the two views.
//   ik_p = We f_p + be + PE(p);  s_hp = u_h . f_p + t_hp (constant query folded, log2 domain);  w_h = softmax_p
//   ikbar_h = sum_p w_hp ik_p = We fbar_h + be + sum_p w_hp PE(p);  o = Wv[head rows] ikbar_h;  mh = Wo o + bo
//   loc_p = mh . ik_p / sqrt(32) = (z . f_p + c0 + mh . PE(p)) / sqrt(32),  z = We^T mh, c0 = be . mh
// =================================================================================================================
constexpr int LW = 8;          // warps per CTA
constexpr int PS = 33;         // padded row stride of 32-wide tables

struct LocalSmem {
  float We[LE][4];
  float be[LE];
  float PE[KT_MAX][PS];
  float Wv[LE][PS];
  float Wo[LE][PS];
  float bo[LE];
  float u2[LH][4];
  float t2[LH][KT_MAX];
  // per warp
  float wl[LW][LH][KT_MAX];
  float ikb[LW][LH][PS];
  float dikb[LW][LH][PS];
  float vo[LW][LE];
  float vmh[LW][LE];
  float vdmh[LW][LE];
  float vdo[LW][LE];
  float vg[LW][KT_MAX];
  int ids[LW][KT_MAX];
};

template <bool CVRP, bool BWD>
__global__ void __launch_bounds__(LW * 32) local_kernel(DecodeBwdArgs A) {
  extern __shared__ __align__(16) unsigned char smraw[];
  LocalSmem& S = *reinterpret_cast<LocalSmem*>(smraw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int DEP = CVRP ? 1 : 0;
  constexpr int F = CVRP ? 3 : 2;
  const float* w = A.weights;
  const float* loc = A.derived + DER_LOC;
  for (int i = tid; i < LE * 4; i += LW * 32) { const int c = i >> 2, f = i & 3; S.We[c][f] = f < F ? w[A.L.loc_we + c * F + f] : 0.f; }
  for (int i = tid; i < LE; i += LW * 32) { S.be[i] = w[A.L.loc_be + i]; S.bo[i] = w[A.L.loc_bo + i]; }
  for (int i = tid; i < KT_MAX * LE; i += LW * 32) S.PE[i / LE][i % LE] = loc[LOC_PE + i];
  for (int i = tid; i < LE * LE; i += LW * 32) { S.Wv[i / LE][i % LE] = w[A.L.loc_wv + i]; S.Wo[i / LE][i % LE] = w[A.L.loc_wo + i]; }
  for (int i = tid; i < LH * 4; i += LW * 32) S.u2[i >> 2][i & 3] = loc[LOC_U + i];
  for (int i = tid; i < LH * KT_MAX; i += LW * 32) S.t2[i / KT_MAX][i % KT_MAX] = loc[LOC_T + i];
  __syncthreads();

  const int N1 = A.N1, NP = A.NP, M = A.M, NL = N1 - DEP, kloc = A.k_local;
  const float isl = 0.17677669529663687f;   // 1 / sqrt(LE)
  // gradient accumulators (BWD): lane = output row / channel
  float aWo[LE], aWv[LE];
  float aBo = 0.f, aWe[4] = {0.f, 0.f, 0.f, 0.f}, aDu[LH][3], aDt[LH][2];
  if (BWD) {
#pragma unroll
    for (int c = 0; c < LE; ++c) aWo[c] = aWv[c] = 0.f;
#pragma unroll
    for (int h = 0; h < LH; ++h) { aDu[h][0] = aDu[h][1] = aDu[h][2] = 0.f; aDt[h][0] = aDt[h][1] = 0.f; }
  }
  const long long total = (long long)A.B * A.nT * M;
  const int hl = lane >> 3;     // head of output row `lane`
  for (long long item = (long long)blockIdx.x * LW + warp; item < total; item += (long long)gridDim.x * LW) {
    const int m = (int)(item % M);
    const int tl = (int)((item / M) % A.nT);
    const int b = (int)(item / ((long long)M * A.nT));
    const StepRec rc = A.rec[((size_t)b * A.T + A.t0 + tl) * M + m];
    if (!rc.active) continue;
    const int cur = rc.cur;
    const uint8_t* nrow = reinterpret_cast<const uint8_t*>(A.t.nbr) + ((size_t)b * N1 + cur) * ELG_NBR_NODE_BYTES(N1);
    const float2* feat = reinterpret_cast<const float2*>(nrow + ELG_NBR_STRIDE);
    // ---- neighbour walk: first k unmasked entries of the distance-sorted list of `cur`
    int cnt = 0;
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = lane + 32 * i;
      int id = 0;
      bool valid = false;
      if (e < NL) {
        id = nrow[nbr_pos(e)];
        valid = !((rc.mask[id >> 5] >> (id & 31)) & 1u);
      }
      const uint32_t bal = __ballot_sync(FULLM, valid);
      const int rank = cnt + __popc(bal & ((1u << lane) - 1u));
      if (valid && rank < kloc) S.ids[warp][rank + DEP] = id;
      cnt += __popc(bal);
    }
    const int kk = min(cnt, kloc);
    const int np = kk + DEP;
    if (DEP && lane == 0) S.ids[warp][0] = 0;
    __syncwarp();
    float dmax = 0.f;
    if (kk > 0) dmax = __ldg(feat + S.ids[warp][kk - 1 + DEP]).x;
    const float r0d = CVRP ? (dmax != 0.f ? 1.f / (dmax + 1e-6f) : 1.f) : 1.f / (dmax + 1e-6f);
    const float r1d = dmax != 0.f ? 1.f / dmax : 1.f;
    const float rld = CVRP ? 1.f / rc.load : 0.f;
    const bool dep_masked = DEP && (rc.mask[0] & 1u);
    // ---- features of this lane's positions
    float f0[2], f1[2], f2[2], pen[2];
    int node[2];
    bool live[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int p = lane + 32 * s;
      f0[s] = f1[s] = f2[s] = pen[s] = 0.f;
      node[s] = 0;
      live[s] = p < np;
      if (live[s] && !(DEP && p == 0)) {
        const int nd = S.ids[warp][p];
        node[s] = nd;
        const float2 ft = __ldg(feat + nd);
        f0[s] = ft.x * r0d;
        f1[s] = ft.y;
        if (CVRP) {
          pen[s] = -(ft.x * r1d);
          f2[s] = A.t.demand[(size_t)b * N1 + nd] * rld;
        } else {
          pen[s] = -f0[s];
        }
      }
    }
    // ---- attention of the constant query (the four heads' reductions are batched)
    float wgt[LH][2], fbar[LH][3];
    {
      float sc[LH][2], mxs[LH];
#pragma unroll
      for (int h = 0; h < LH; ++h) {
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          const int p = lane + 32 * s;
          float v = -INFINITY;
          if (live[s]) {
            v = fmaf(S.u2[h][2], f2[s], fmaf(S.u2[h][1], f1[s], S.u2[h][0] * f0[s])) + S.t2[h][p];
            if (DEP && p == 0 && dep_masked) v = -INFINITY;
          }
          sc[h][s] = v;
        }
        mxs[h] = fmaxf(sc[h][0], sc[h][1]);
      }
      warp_allreduce<LH, true>(mxs);
      float red[16];
#pragma unroll
      for (int h = 0; h < LH; ++h) {
        const float mref = mxs[h] == -INFINITY ? 0.f : mxs[h];
        wgt[h][0] = exp2f(sc[h][0] - mref);
        wgt[h][1] = exp2f(sc[h][1] - mref);
        red[h] = wgt[h][0] + wgt[h][1];
        red[4 + 3 * h + 0] = wgt[h][0] * f0[0] + wgt[h][1] * f0[1];
        red[4 + 3 * h + 1] = wgt[h][0] * f1[0] + wgt[h][1] * f1[1];
        red[4 + 3 * h + 2] = wgt[h][0] * f2[0] + wgt[h][1] * f2[1];
      }
      warp_allreduce<16, false>(red);
#pragma unroll
      for (int h = 0; h < LH; ++h) {
        const float inv = red[h] > 0.f ? 1.f / red[h] : 0.f;
        wgt[h][0] *= inv;
        wgt[h][1] *= inv;
        S.wl[warp][h][lane] = wgt[h][0];
        S.wl[warp][h][lane + 32] = wgt[h][1];
        fbar[h][0] = red[4 + 3 * h + 0] * inv;
        fbar[h][1] = red[4 + 3 * h + 1] * inv;
        fbar[h][2] = red[4 + 3 * h + 2] * inv;
      }
    }
    __syncwarp();
    // ikbar_h[c], lane = c
    float ikb[LH];
    {
      float acc[LH] = {0.f, 0.f, 0.f, 0.f};
      for (int p = 0; p < np; ++p) {
        const float pe = S.PE[p][lane];
#pragma unroll
        for (int h = 0; h < LH; ++h) acc[h] = fmaf(S.wl[warp][h][p], pe, acc[h]);
      }
#pragma unroll
      for (int h = 0; h < LH; ++h) {
        ikb[h] = fmaf(S.We[lane][2], fbar[h][2], fmaf(S.We[lane][1], fbar[h][1], S.We[lane][0] * fbar[h][0])) + S.be[lane] + acc[h];
        S.ikb[warp][h][lane] = ikb[h];
      }
    }
    __syncwarp();
    // o[r], lane = r = h * 8 + d
    float ov = 0.f;
#pragma unroll
    for (int c = 0; c < LE; ++c) ov = fmaf(S.Wv[lane][c], S.ikb[warp][hl][c], ov);
    S.vo[warp][lane] = ov;
    __syncwarp();
    float mh = S.bo[lane];
#pragma unroll
    for (int c = 0; c < LE; ++c) mh = fmaf(S.Wo[lane][c], S.vo[warp][c], mh);
    S.vmh[warp][lane] = mh;
    __syncwarp();
    const size_t row = ((size_t)b * A.nT + tl) * M + m;
    if (!BWD) {      // the local scores themselves are only needed by the forward pass
      float zc[4] = {mh * S.We[lane][0], mh * S.We[lane][1], mh * S.We[lane][2], mh * S.be[lane]};
      warp_allreduce<4, false>(zc);
      const float z0 = zc[0], z1 = zc[1], z2 = zc[2], c0 = zc[3];
      float locv[2];
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const int p = min(lane + 32 * s, KT_MAX - 1);
        float pm = 0.f;
#pragma unroll
        for (int c = 0; c < LE; ++c) pm = fmaf(S.vmh[warp][c], S.PE[p][c], pm);
        locv[s] = (fmaf(z2, f2[s], fmaf(z1, f1[s], z0 * f0[s])) + c0 + pm) * isl;
      }
      float* add = A.add + row * NP;
      for (int j = lane; j < NP; j += 32) add[j] = (DEP && j == 0) ? 0.f : A.xi;
      __syncwarp();
#pragma unroll
      for (int s = 0; s < 2; ++s)
        if (live[s]) add[node[s]] = pen[s] + locv[s];
      continue;
    }
    // =========================== backward ===========================
    const float* dxr = A.dx + row * NP;
    float g[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      g[s] = live[s] ? dxr[node[s]] * isl : 0.f;
      S.vg[warp][lane + 32 * s] = g[s];
    }
    float dzc[4] = {g[0] * f0[0] + g[1] * f0[1], g[0] * f1[0] + g[1] * f1[1], g[0] * f2[0] + g[1] * f2[1], g[0] + g[1]};
    warp_allreduce<4, false>(dzc);
    const float dz0 = dzc[0], dz1 = dzc[1], dz2 = dzc[2], dc0 = dzc[3];
    __syncwarp();
    // d mh[c], lane = c
    float dmh = fmaf(S.We[lane][2], dz2, fmaf(S.We[lane][1], dz1, S.We[lane][0] * dz0)) + S.be[lane] * dc0;
    for (int p = 0; p < np; ++p) dmh = fmaf(S.vg[warp][p], S.PE[p][lane], dmh);
    aWe[0] = fmaf(mh, dz0, aWe[0]); aWe[1] = fmaf(mh, dz1, aWe[1]); aWe[2] = fmaf(mh, dz2, aWe[2]); aWe[3] = fmaf(mh, dc0, aWe[3]);
    aBo += dmh;
#pragma unroll
    for (int c = 0; c < LE; ++c) aWo[c] = fmaf(dmh, S.vo[warp][c], aWo[c]);
    S.vdmh[warp][lane] = dmh;
    __syncwarp();
    float dov = 0.f;      // d o[c], lane = c
#pragma unroll
    for (int k = 0; k < LE; ++k) dov = fmaf(S.Wo[k][lane], S.vdmh[warp][k], dov);
    S.vdo[warp][lane] = dov;
#pragma unroll
    for (int c = 0; c < LE; ++c) aWv[c] = fmaf(dov, S.ikb[warp][hl][c], aWv[c]);
    __syncwarp();
    float dfb[LH][3];
    {
      float red[16];
#pragma unroll
      for (int h = 0; h < LH; ++h) {
        float a = 0.f;
#pragma unroll
        for (int d = 0; d < LD; ++d) a = fmaf(S.Wv[h * LD + d][lane], S.vdo[warp][h * LD + d], a);
        S.dikb[warp][h][lane] = a;
        aWe[0] = fmaf(a, fbar[h][0], aWe[0]); aWe[1] = fmaf(a, fbar[h][1], aWe[1]); aWe[2] = fmaf(a, fbar[h][2], aWe[2]);
        aWe[3] += a;
        red[3 * h + 0] = S.We[lane][0] * a;
        red[3 * h + 1] = S.We[lane][1] * a;
        red[3 * h + 2] = S.We[lane][2] * a;
      }
      red[12] = red[13] = red[14] = red[15] = 0.f;
      warp_allreduce<16, false>(red);
#pragma unroll
      for (int h = 0; h < LH; ++h) { dfb[h][0] = red[3 * h]; dfb[h][1] = red[3 * h + 1]; dfb[h][2] = red[3 * h + 2]; }
    }
    __syncwarp();
    {
      float dw[LH][2], wb[LH];
#pragma unroll
      for (int h = 0; h < LH; ++h) {
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          const int p = min(lane + 32 * s, KT_MAX - 1);
          float a = fmaf(dfb[h][2], f2[s], fmaf(dfb[h][1], f1[s], dfb[h][0] * f0[s]));
#pragma unroll
          for (int c = 0; c < LE; ++c) a = fmaf(S.dikb[warp][h][c], S.PE[p][c], a);
          dw[h][s] = a;
        }
        wb[h] = wgt[h][0] * dw[h][0] + wgt[h][1] * dw[h][1];
      }
      warp_allreduce<LH, false>(wb);
#pragma unroll
      for (int h = 0; h < LH; ++h) {
        const float ds0 = wgt[h][0] * (dw[h][0] - wb[h]), ds1 = wgt[h][1] * (dw[h][1] - wb[h]);
        aDt[h][0] += ds0;
        aDt[h][1] += ds1;
        aDu[h][0] += ds0 * f0[0] + ds1 * f0[1];       // per-lane partial sums, reduced once at the end of the kernel
        aDu[h][1] += ds0 * f1[0] + ds1 * f1[1];
        aDu[h][2] += ds0 * f2[0] + ds1 * f2[1];
      }
    }
  }
  if (BWD) {
    float* lg = A.lg;
#pragma unroll
    for (int c = 0; c < LE; ++c) {
      atomicAdd(lg + LG_WO + lane * LE + c, aWo[c]);
      atomicAdd(lg + LG_WV + lane * LE + c, aWv[c]);
    }
    atomicAdd(lg + LG_BO + lane, aBo);
#pragma unroll
    for (int f = 0; f < 3; ++f) atomicAdd(lg + LG_WE + lane * 4 + f, aWe[f]);
    atomicAdd(lg + LG_BE + lane, aWe[3]);
#pragma unroll
    for (int h = 0; h < LH; ++h) {
      atomicAdd(lg + LG_DT + h * KT_MAX + lane, aDt[h][0]);
      atomicAdd(lg + LG_DT + h * KT_MAX + lane + 32, aDt[h][1]);
      const float u0 = warp_sum(aDu[h][0]), u1 = warp_sum(aDu[h][1]), u2 = warp_sum(aDu[h][2]);
      if (lane < 3) atomicAdd(lg + LG_DU + h * 4 + lane, lane == 0 ? u0 : (lane == 1 ? u1 : u2));
    }
  }
}

int launch_local(const DecodeBwdArgs& a, bool bwd, cudaStream_t st) {
  const size_t smem = sizeof(LocalSmem);
  static bool attr = false;
  if (!attr) {
    ELG_CUDA_OK(cudaFuncSetAttribute(local_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ELG_CUDA_OK(cudaFuncSetAttribute(local_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ELG_CUDA_OK(cudaFuncSetAttribute(local_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ELG_CUDA_OK(cudaFuncSetAttribute(local_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  const long long total = (long long)a.B * a.nT * a.M;
  long long want = (total + LW - 1) / LW;
  const int grid = (int)(want < 148 * 4 ? want : 148 * 4);
  const bool cvrp = a.problem == ELG_CVRP;
  if (cvrp && !bwd) local_kernel<true, false><<<grid, LW * 32, smem, st>>>(a);
  else if (cvrp && bwd) local_kernel<true, true><<<grid, LW * 32, smem, st>>>(a);
  else if (!bwd) local_kernel<false, false><<<grid, LW * 32, smem, st>>>(a);
  else local_kernel<false, true><<<grid, LW * 32, smem, st>>>(a);
  ELG_LAUNCH_OK();
  return ELG_OK;
}

// =================================================================================================================
// Global policy: one CTA per (instance, step), K', V, E' of the instance staged in shared memory (row stride 132
// floats: conflict-free both for "lane = node" row reads and "lane = 4 channels" column sweeps); one warp per row.
// =================================================================================================================
constexpr int GS = 132;        // table row stride (floats)
constexpr int WS = 113;        // stride of the per-head softmax-weight rows (odd: heads land in different banks)

__device__ __forceinline__ float dot16(const float* __restrict__ a, const float* __restrict__ b) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 x = *reinterpret_cast<const float4*>(a + 4 * i);
    const float4 y = *reinterpret_cast<const float4*>(b + 4 * i);
    s = fmaf(x.x, y.x, s); s = fmaf(x.y, y.y, s); s = fmaf(x.z, y.z, s); s = fmaf(x.w, y.w, s);
  }
  return s;
}

// NC consecutive 32-bit tensor-memory columns of this thread's lane <-> registers
template <int NC>
__device__ __forceinline__ void tmem_load(uint32_t taddr, float (&v)[NC]) {
  uint32_t r[NC];
#pragma unroll
  for (int c = 0; c < NC; c += 8) umma::ld8_nw(taddr + c, r + c);
  umma::wait_ld();
#pragma unroll
  for (int c = 0; c < NC; ++c) v[c] = umma::after_wait(r[c]);
}
template <int NC>
__device__ __forceinline__ void tmem_store(uint32_t taddr, const float (&v)[NC]) {
  uint32_t r[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) r[c] = __float_as_uint(v[c]);
#pragma unroll
  for (int c = 0; c < NC; c += 8) umma::st8(taddr + c, r + c);
  umma::wait_st();
}

constexpr int GTS = 2;         // rollout steps per CTA (tables staged once, accumulators flushed once)

// Per 8-row batch: phase 1a (warp = row) query, attention weights, attention output, scores, softmax, d logits, d o;
// phase 2a (warp = node slice, lane = 4 channels) d V += w^T d o over the batch's rows; phase 1b softmax backward
// (d s overwrites w) and d q; phase 2b d K' += d s^T q.  d V and d K' live in registers for the whole CTA.
template <bool CVRP, int GW>
__global__ void __launch_bounds__(GW * 32) global_bwd_kernel(DecodeBwdArgs A) {
  constexpr int GNJ = (TRAIN_MAX_NODES + GW - 1) / GW;     // nodes per warp in the accumulation phases
  extern __shared__ __align__(16) float gsm[];
  const int N1 = A.N1, NP = A.NP, M = A.M;
  float* sK = gsm;
  float* sV = sK + N1 * GS;
  float* sE = sV + N1 * GS;
  float* seb = sE + N1 * GS;            // [128]
  float* swl = seb + 128;               // [128]
  float* sdeb = swl + 128;              // [128]  d eb of this CTA (shared-memory atomics)
  float* pw = sdeb + 128;               // per warp: sq[128] so[128] sdo[128] sdx[128] sw[H][WS]
  __shared__ int sact[GW];
  __shared__ uint32_t sidx_all[GW * 32];     // per warp: compact list of selectable nodes (uint8 x 128)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nTB = (A.nT + GTS - 1) / GTS;
  const int b = blockIdx.x / nTB, tb = blockIdx.x % nTB;
  constexpr int PWF = 4 * 128 + H * WS;
  float* sq = pw + warp * PWF;
  float* so = sq + 128;
  float* sdo = so + 128;
  float* sdx = sdo + 128;
  float* sw = sdx + 128;
  {
    const float* gk = A.t.k + (size_t)b * N1 * E;
    const float* gv = A.t.v + (size_t)b * N1 * E;
    const float* ge = A.eplain + (size_t)b * N1 * E;
    for (int i = tid; i < N1 * (E / 4); i += GW * 32) {
      const int j = i >> 5, c = (i & 31) * 4;
      *reinterpret_cast<float4*>(sK + j * GS + c) = *reinterpret_cast<const float4*>(gk + j * E + c);
      *reinterpret_cast<float4*>(sV + j * GS + c) = *reinterpret_cast<const float4*>(gv + j * E + c);
      *reinterpret_cast<float4*>(sE + j * GS + c) = *reinterpret_cast<const float4*>(ge + j * E + c);
    }
    for (int i = tid; i < 128; i += GW * 32) {
      seb[i] = i < N1 ? A.t.eb[(size_t)b * N1 + i] : 0.f;
      swl[i] = A.derived[DER_WL + i];
      sdeb[i] = 0.f;
    }
  }
  __syncthreads();
  const int c4 = lane * 4, hl = lane >> 2;
  float4 dwl_acc = make_float4(0.f, 0.f, 0.f, 0.f);
  // d V, d K', d E' accumulators of this thread (GNJ nodes x 4 channels each) live in tensor memory: 2 KB per thread
  // that nothing else uses here, read-modify-written once per 12-row batch with tcgen05.ld / tcgen05.st
  constexpr int NC = GNJ * 4;
  static_assert(NC % 8 == 0 && 3 * NC * ((GW + 3) / 4) <= 512, "tensor-memory budget");
  __shared__ uint32_t s_tmem;
  if (warp == 0) umma::tmem_alloc(&s_tmem, 512);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tacc = s_tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(warp >> 2) * (3 * NC);
  {
    uint32_t zr[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
#pragma unroll
    for (int c = 0; c < 3 * NC; c += 8) umma::st8(tacc + c, zr);
    umma::wait_st();
  }
  const float ln2 = 0.6931471805599453f;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const int tl_end = min(A.nT, tb * GTS + GTS);
  // ---- the active rows of this CTA's steps, packed: batches of GW rows are taken from this list, so neither the tail
  // of a step (100 rows = 8 x 12 + 4) nor the sparse late steps leave warps idle at the batch barriers
  __shared__ uint16_t slist[GTS * 128];
  __shared__ int s_nact;
  if (tid == 0) s_nact = 0;
  __syncthreads();
  for (int e = tid; e < (tl_end - tb * GTS) * M; e += GW * 32) {
    const int tl = tb * GTS + e / M, m = e % M;
    if (A.rec[((size_t)b * A.T + A.t0 + tl) * M + m].active) slist[atomicAdd(&s_nact, 1)] = (uint16_t)(((tl - tb * GTS) << 8) | m);
  }
  __syncthreads();
  const int nact = s_nact;
  {
    for (int m0 = 0; m0 < nact; m0 += GW) {
      const bool valid = m0 + warp < nact;
      const int item = valid ? (int)slist[m0 + warp] : 0;
      const int tl = tb * GTS + (item >> 8), m = item & 255;
      const StepRec* recs = A.rec + ((size_t)b * A.T + A.t0 + tl) * M;
      const size_t row0 = ((size_t)b * A.nT + tl) * M;
      StepRec rc;
      rc.active = 0;
      if (valid) rc = recs[m];
      const bool act = valid && rc.active;
      if (lane == 0) sact[warp] = act ? 1 : 0;
      const size_t row = row0 + m;
      float* gdx = A.dx + row * NP;
      const int cur = rc.cur;
      const float load = CVRP ? rc.load : 0.f;
      const int first = CVRP ? 0 : __float_as_int(rc.load);
      if (act) {
        // ---- compact list of the row's selectable nodes: every loop below runs over these only (masked nodes have
        // weight, d logit and d score exactly 0); their count shrinks from N1 to 2 over the rollout
        uint8_t* sidx = reinterpret_cast<uint8_t*>(sidx_all + warp * 32);
        int cnt = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int j = lane + 32 * i;
          const bool open = j < N1 && !((rc.mask[i] >> lane) & 1u);
          const uint32_t bal = __ballot_sync(FULLM, open);
          if (open) sidx[cnt + __popc(bal & ((1u << lane) - 1u))] = (uint8_t)j;
          cnt += __popc(bal);
        }
        // weights / d scores of masked nodes must read as zero in the accumulation phases
        for (int i = lane; i < H * WS / 4; i += 32) reinterpret_cast<float4*>(sw)[i] = z4;
        *reinterpret_cast<float4*>(sdx + c4) = z4;
        if (c4 < NP) *reinterpret_cast<float4*>(gdx + c4) = z4;
        // ---- query (CVRP/models.py:336-340: Wq_last [enc[cur]; load]; TSP/models.py:258-260: q_first + Wq_last enc[cur])
        float4 q4 = *reinterpret_cast<const float4*>(A.t.qtab + ((size_t)b * N1 + cur) * E + c4);
        if (CVRP) {
          const float4 wl4 = *reinterpret_cast<const float4*>(swl + c4);
          q4.x = fmaf(load, wl4.x, q4.x); q4.y = fmaf(load, wl4.y, q4.y); q4.z = fmaf(load, wl4.z, q4.z); q4.w = fmaf(load, wl4.w, q4.w);
        } else {
          const float4 f4 = *reinterpret_cast<const float4*>(A.t.qfirst + ((size_t)b * N1 + first) * E + c4);
          q4.x += f4.x; q4.y += f4.y; q4.z += f4.z; q4.w += f4.w;
        }
        *reinterpret_cast<float4*>(sq + c4) = q4;
        __syncwarp();
        const int nslot = (cnt + 31) >> 5;          // node slots of 32 lanes over the compact list
        int jn[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) jn[i] = (lane + 32 * i) < cnt ? (int)sidx[lane + 32 * i] : -1;
        // ---- multi-head attention weights (log2 domain: K' carries log2(e)/sqrt(D))
#pragma unroll 1
        for (int h = 0; h < H; ++h) {
          float sc[4];
          float mx = -INFINITY;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float s = -INFINITY;
            if (i < nslot && jn[i] >= 0) s = dot16(sq + h * D, sK + jn[i] * GS + h * D);
            sc[i] = s;
            mx = fmaxf(mx, s);
          }
          mx = warp_max(mx);
          float sum = 0.f;
#pragma unroll
          for (int i = 0; i < 4; ++i) { sc[i] = exp2f(sc[i] - mx); sum += sc[i]; }
          sum = warp_sum(sum);
          const float inv = 1.f / sum;
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (jn[i] >= 0) sw[h * WS + jn[i]] = sc[i] * inv;
        }
        __syncwarp();
        // ---- attention output o[c], lane = 4 channels of head hl
        float4 o4 = z4;
        for (int k = 0; k < cnt; ++k) {
          const int j = sidx[k];
          const float wv = sw[hl * WS + j];
          const float4 v4 = *reinterpret_cast<const float4*>(sV + j * GS + c4);
          o4.x = fmaf(wv, v4.x, o4.x); o4.y = fmaf(wv, v4.y, o4.y); o4.z = fmaf(wv, v4.z, o4.z); o4.w = fmaf(wv, v4.w, o4.w);
        }
        *reinterpret_cast<float4*>(so + c4) = o4;
        __syncwarp();
        // ---- scores, clipping, softmax over the nodes, d logits
        float th[4], lg[4];
        float mx = -INFINITY;
        const float* addr = A.add + row * NP;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          lg[i] = -INFINITY; th[i] = 0.f;
          if (i < nslot && jn[i] >= 0) {
            const int j = jn[i];
            float s = seb[j];
            const float* er = sE + j * GS;
#pragma unroll 8
            for (int c = 0; c < E; c += 4) {
              const float4 a = *reinterpret_cast<const float4*>(so + c);
              const float4 e4 = *reinterpret_cast<const float4*>(er + c);
              s = fmaf(a.x, e4.x, s); s = fmaf(a.y, e4.y, s); s = fmaf(a.z, e4.z, s); s = fmaf(a.w, e4.w, s);
            }
            th[i] = tanhf(s + addr[j]);
            lg[i] = A.clip * th[i];
          }
          mx = fmaxf(mx, lg[i]);
        }
        mx = warp_max(mx);
        float pe[4], sum = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) { pe[i] = __expf(lg[i] - mx); sum += pe[i]; }
        sum = warp_sum(sum);
        const float inv = 1.f / sum;
        const float cf = A.coef[(size_t)b * M + m];
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (jn[i] >= 0) {
            const int j = jn[i];
            const float dl = cf * ((j == rc.act ? 1.f : 0.f) - pe[i] * inv);
            const float dxv = dl * A.clip * (1.f - th[i] * th[i]);
            sdx[j] = dxv;
            gdx[j] = dxv;
            if (dxv != 0.f) atomicAdd(sdeb + j, dxv);
          }
        }
        __syncwarp();
        // ---- d o = sum_j dx_j E'_j
        float4 d4 = z4;
        for (int k = 0; k < cnt; ++k) {
          const int j = sidx[k];
          const float x = sdx[j];
          const float4 e4 = *reinterpret_cast<const float4*>(sE + j * GS + c4);
          d4.x = fmaf(x, e4.x, d4.x); d4.y = fmaf(x, e4.y, d4.y); d4.z = fmaf(x, e4.z, d4.z); d4.w = fmaf(x, e4.w, d4.w);
        }
        *reinterpret_cast<float4*>(sdo + c4) = d4;
      }
      __syncthreads();
      // ---- phase 2a: d V[j][c] += w_r[h(c)][j] d o_r[c] and d E'[j][c] += dx_r[j] o_r[c] over the rows of the batch
      // (warp = nodes warp, warp + GW, ...; both accumulators fetched from tensor memory with one wait)
      {
        float accv[NC], acce[NC];
        {
          uint32_t r0[NC], r1[NC];
#pragma unroll
          for (int c = 0; c < NC; c += 8) umma::ld8_nw(tacc + c, r0 + c);
#pragma unroll
          for (int c = 0; c < NC; c += 8) umma::ld8_nw(tacc + 2 * NC + c, r1 + c);
          umma::wait_ld();
#pragma unroll
          for (int c = 0; c < NC; ++c) { accv[c] = umma::after_wait(r0[c]); acce[c] = umma::after_wait(r1[c]); }
        }
#pragma unroll 1
        for (int r = 0; r < GW; ++r) {
          if (!sact[r]) continue;
          const float* rw = pw + r * PWF + 4 * 128 + hl * WS;
          const float* rx = pw + r * PWF + 3 * 128;
          const float4 g4 = *reinterpret_cast<const float4*>(pw + r * PWF + 2 * 128 + c4);
          const float4 o4 = *reinterpret_cast<const float4*>(pw + r * PWF + 128 + c4);
#pragma unroll
          for (int i = 0; i < GNJ; ++i) {
            const int j = warp + GW * i;
            const float wv = j < N1 ? rw[j] : 0.f;
            const float xv = j < N1 ? rx[j] : 0.f;
            accv[4 * i + 0] = fmaf(wv, g4.x, accv[4 * i + 0]); accv[4 * i + 1] = fmaf(wv, g4.y, accv[4 * i + 1]);
            accv[4 * i + 2] = fmaf(wv, g4.z, accv[4 * i + 2]); accv[4 * i + 3] = fmaf(wv, g4.w, accv[4 * i + 3]);
            acce[4 * i + 0] = fmaf(xv, o4.x, acce[4 * i + 0]); acce[4 * i + 1] = fmaf(xv, o4.y, acce[4 * i + 1]);
            acce[4 * i + 2] = fmaf(xv, o4.z, acce[4 * i + 2]); acce[4 * i + 3] = fmaf(xv, o4.w, acce[4 * i + 3]);
          }
        }
        {
          uint32_t r0[NC], r1[NC];
#pragma unroll
          for (int c = 0; c < NC; ++c) { r0[c] = __float_as_uint(accv[c]); r1[c] = __float_as_uint(acce[c]); }
#pragma unroll
          for (int c = 0; c < NC; c += 8) umma::st8(tacc + c, r0 + c);
#pragma unroll
          for (int c = 0; c < NC; c += 8) umma::st8(tacc + 2 * NC + c, r1 + c);
          umma::wait_st();
        }
      }
      __syncthreads();
      if (act) {
        // ---- softmax backward per head: d s2_hj = ln2 * w_hj (d w_hj - sum_j w d w), in place over w (compact list)
        const uint8_t* sidx = reinterpret_cast<const uint8_t*>(sidx_all + warp * 32);
        int cnt = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int nb = N1 - 32 * i;
          const uint32_t fullw = nb >= 32 ? FULLM : (nb > 0 ? ((1u << nb) - 1u) : 0u);
          cnt += __popc(~rc.mask[i] & fullw);
        }
        int jn[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) jn[i] = (lane + 32 * i) < cnt ? (int)sidx[lane + 32 * i] : -1;
#pragma unroll 1
        for (int h = 0; h < H; ++h) {
          float dw[4], wv[4];
          float part = 0.f;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            dw[i] = 0.f; wv[i] = 0.f;
            if (jn[i] >= 0) {
              wv[i] = sw[h * WS + jn[i]];
              dw[i] = dot16(sdo + h * D, sV + jn[i] * GS + h * D);
            }
            part = fmaf(wv[i], dw[i], part);
          }
          const float wbar = warp_sum(part);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (jn[i] >= 0) sw[h * WS + jn[i]] = ln2 * wv[i] * (dw[i] - wbar);
        }
        __syncwarp();
        // ---- d q = sum_j d s2_hj K'_j ; scatter into the query-table gradient
        float4 g4 = z4;
        for (int k = 0; k < cnt; ++k) {
          const int j = sidx[k];
          const float x = sw[hl * WS + j];
          const float4 k4 = *reinterpret_cast<const float4*>(sK + j * GS + c4);
          g4.x = fmaf(x, k4.x, g4.x); g4.y = fmaf(x, k4.y, g4.y); g4.z = fmaf(x, k4.z, g4.z); g4.w = fmaf(x, k4.w, g4.w);
        }
        float* dq = A.dqtab + ((size_t)b * N1 + cur) * E + c4;
        atomicAdd(dq + 0, g4.x); atomicAdd(dq + 1, g4.y); atomicAdd(dq + 2, g4.z); atomicAdd(dq + 3, g4.w);
        if (CVRP) {
          dwl_acc.x = fmaf(load, g4.x, dwl_acc.x); dwl_acc.y = fmaf(load, g4.y, dwl_acc.y);
          dwl_acc.z = fmaf(load, g4.z, dwl_acc.z); dwl_acc.w = fmaf(load, g4.w, dwl_acc.w);
        } else {
          float* df = A.dqfirst + ((size_t)b * N1 + first) * E + c4;
          atomicAdd(df + 0, g4.x); atomicAdd(df + 1, g4.y); atomicAdd(df + 2, g4.z); atomicAdd(df + 3, g4.w);
        }
      }
      __syncthreads();
      // ---- phase 2b: d K'[j][c] += d s_r[h(c)][j] q_r[c]
      {
        float acc[NC];
        tmem_load<NC>(tacc + NC, acc);
#pragma unroll 1
        for (int r = 0; r < GW; ++r) {
          if (!sact[r]) continue;
          const float* rw = pw + r * PWF + 4 * 128 + hl * WS;
          const float4 g4 = *reinterpret_cast<const float4*>(pw + r * PWF + c4);
#pragma unroll
          for (int i = 0; i < GNJ; ++i) {
            const int j = warp + GW * i;
            const float wv = j < N1 ? rw[j] : 0.f;
            acc[4 * i + 0] = fmaf(wv, g4.x, acc[4 * i + 0]); acc[4 * i + 1] = fmaf(wv, g4.y, acc[4 * i + 1]);
            acc[4 * i + 2] = fmaf(wv, g4.z, acc[4 * i + 2]); acc[4 * i + 3] = fmaf(wv, g4.w, acc[4 * i + 3]);
          }
        }
        tmem_store<NC>(tacc + NC, acc);
      }
      __syncthreads();
    }
  }
  // ---- flush: tensor memory -> global accumulators
#pragma unroll 1
  for (int mi = 0; mi < 3; ++mi) {
    float acc[NC];
    tmem_load<NC>(tacc + mi * NC, acc);
    float* gM = (mi == 0 ? A.dV : (mi == 1 ? A.dK : A.dEp)) + (size_t)b * N1 * E;
#pragma unroll
    for (int i = 0; i < GNJ; ++i) {
      const int j = warp + GW * i;
      if (j < N1) {
        float* pm = gM + j * E + c4;
        atomicAdd(pm + 0, acc[4 * i + 0]); atomicAdd(pm + 1, acc[4 * i + 1]);
        atomicAdd(pm + 2, acc[4 * i + 2]); atomicAdd(pm + 3, acc[4 * i + 3]);
      }
    }
  }
  __syncthreads();
  for (int j = tid; j < N1; j += GW * 32) atomicAdd(A.deb + (size_t)b * N1 + j, sdeb[j]);
  if (CVRP) {
    atomicAdd(A.dwl + c4 + 0, dwl_acc.x); atomicAdd(A.dwl + c4 + 1, dwl_acc.y);
    atomicAdd(A.dwl + c4 + 2, dwl_acc.z); atomicAdd(A.dwl + c4 + 3, dwl_acc.w);
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(s_tmem, 512);
}

template <bool CVRP, int GW>
static int launch_global_bwd_t(const DecodeBwdArgs& a, cudaStream_t st) {
  size_t smem = ((size_t)3 * a.N1 * GS + 384 + (size_t)GW * (4 * 128 + H * WS)) * sizeof(float);
  // every CTA allocates all 512 tensor-memory columns: keep it to one CTA per SM (a second one would sit in
  // tcgen05.alloc until the first is done), also for small instances whose tables would leave room for two
  if (smem < 116 * 1024) smem = 116 * 1024;
  ELG_REQUIRE(a.M <= 128, ELG_EUNSUPPORTED, "training supports a POMO width of up to 128 rows (got %d)", a.M);
  ELG_REQUIRE(smem <= 226 * 1024, ELG_EUNSUPPORTED, "training supports up to %d nodes (needs %zu bytes of shared memory)", TRAIN_MAX_NODES, smem);
  ELG_CUDA_OK(cudaFuncSetAttribute(global_bwd_kernel<CVRP, GW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const unsigned grid = (unsigned)(a.B * ((a.nT + GTS - 1) / GTS));
  global_bwd_kernel<CVRP, GW><<<grid, GW * 32, smem, st>>>(a);
  ELG_LAUNCH_OK();
  return ELG_OK;
}

// 12 warps per CTA when the instance tables leave room for their scratch (N1 <= 101), else 8
int launch_global_bwd(const DecodeBwdArgs& a, cudaStream_t st) {
  const size_t smem12 = ((size_t)3 * a.N1 * GS + 384 + (size_t)12 * (4 * 128 + H * WS)) * sizeof(float);
  const bool wide = smem12 <= 226 * 1024;
  if (a.problem == ELG_CVRP) return wide ? launch_global_bwd_t<true, 12>(a, st) : launch_global_bwd_t<true, 8>(a, st);
  return wide ? launch_global_bwd_t<false, 12>(a, st) : launch_global_bwd_t<false, 8>(a, st);
}

// =================================================================================================================
// Chain rule through the constant-query folds of the local policy (one CTA):
//   q = Wq token;  a_h = Wk_h^T q_h / sqrt(8);  u_h = We^T a_h;  t_hp = a_h . (be + PE(p))
// =================================================================================================================
__global__ void local_fold_bwd_kernel(elg_model_desc d, elg_weight_layout_t L, const float* __restrict__ w,
                                      const float* __restrict__ der, const float* __restrict__ lg,
                                      float* __restrict__ grads) {
  __shared__ float q[LE], a[LH][LE], da[LH][LE], dq[LE], sdt[LH], We[LE][4];
  const int tid = threadIdx.x, nt = blockDim.x;
  const int F = d.problem == ELG_CVRP ? 3 : 2;
  const float isd = 0.35355339059327373f;   // 1 / sqrt(LD)
  const float* pe = der + DER_LOC + LOC_PE;
  for (int c = tid; c < LE; c += nt) {
    float s = 0.f;
    for (int i = 0; i < LE; ++i) s = fmaf(w[L.loc_wq + c * LE + i], w[L.loc_token + i], s);
    q[c] = s;
    for (int f = 0; f < 4; ++f) We[c][f] = f < F ? w[L.loc_we + c * F + f] : 0.f;
  }
  for (int h = tid; h < LH; h += nt) {
    float s = 0.f;
    for (int p = 0; p < KT_MAX; ++p) s += lg[LG_DT + h * KT_MAX + p];
    sdt[h] = s;
  }
  __syncthreads();
  for (int i = tid; i < LH * LE; i += nt) {
    const int h = i / LE, c = i % LE;
    float s = 0.f;
    for (int dd = 0; dd < LD; ++dd) s = fmaf(w[L.loc_wk + (h * LD + dd) * LE + c], q[h * LD + dd], s);
    a[h][c] = s * isd;
    float g = 0.f;
    for (int f = 0; f < 3; ++f) g = fmaf(We[c][f], lg[LG_DU + h * 4 + f], g);
    for (int p = 0; p < KT_MAX; ++p) g = fmaf(lg[LG_DT + h * KT_MAX + p], w[L.loc_be + c] + pe[p * LE + c], g);
    da[h][c] = g;
  }
  __syncthreads();
  for (int r = tid; r < LE; r += nt) {
    const int h = r / LD;
    float s = 0.f;
    for (int c = 0; c < LE; ++c) s = fmaf(w[L.loc_wk + r * LE + c], da[h][c], s);
    dq[r] = s * isd;
  }
  __syncthreads();
  for (int i = tid; i < LE * LE; i += nt) {
    const int r = i / LE, c = i % LE, h = r / LD;
    grads[L.loc_wk + i] += q[r] * da[h][c] * isd;
    grads[L.loc_wq + i] += dq[r] * w[L.loc_token + c];
    grads[L.loc_wv + i] += lg[LG_WV + i];
    grads[L.loc_wo + i] += lg[LG_WO + i];
  }
  for (int c = tid; c < LE; c += nt) {
    float s = 0.f;
    for (int r = 0; r < LE; ++r) s = fmaf(w[L.loc_wq + r * LE + c], dq[r], s);
    grads[L.loc_token + c] += s;
    grads[L.loc_bo + c] += lg[LG_BO + c];
    float gb = lg[LG_BE + c];
    for (int h = 0; h < LH; ++h) gb = fmaf(a[h][c], sdt[h], gb);
    grads[L.loc_be + c] += gb;
    for (int f = 0; f < F; ++f) {
      float g = lg[LG_WE + c * 4 + f];
      for (int h = 0; h < LH; ++h) g = fmaf(a[h][c], lg[LG_DU + h * 4 + f], g);
      grads[L.loc_we + c * F + f] += g;
    }
  }
}

int launch_local_fold_bwd(const elg_model_desc* d, const elg_weight_layout_t& L, const float* weights, const float* derived,
                          const float* lg, float* grads, cudaStream_t st) {
  local_fold_bwd_kernel<<<1, 256, 0, st>>>(*d, L, weights, derived, lg, grads);
  ELG_LAUNCH_OK();
  return ELG_OK;
}

}  // namespace elg
