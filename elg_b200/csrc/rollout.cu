// Persistent rollout kernel: the whole autoregressive construction loop of one aug-instance
// (decode step of the global POMO attention policy + local k-nearest attention policy + env step)
// runs inside one CTA, with the decoder keys/values/score matrix resident in shared memory.
//
// reference per-step path (file:line under the reference tree):
//   rollout                      CVRP/utils.py:7-29            TSP/utils.py:7-26
//   get_cur_feature              CVRP/CVRPEnv.py:291-318       TSP/TSPEnv.py:135-156
//   one_step_rollout             CVRP/CVRPModel.py:36-75       TSP/TSPModel.py:26-64
//   Decoder.forward              CVRP/models.py:322-423        TSP/models.py:244-303
//   local_policy_att.forward     CVRP/models.py:51-175         TSP/models.py:48-110
//   env.step                     CVRP/CVRPEnv.py:190-249       TSP/TSPEnv.py:108-133
//   reward                       CVRP/CVRPEnv.py:251-288       TSP/TSPEnv.py:158-184
//
// Work decomposition (DESIGN.md has the full story):
//   CTA   = one (aug-instance, tile of <= 64 POMO rows), 16 warps, persistent over work items.
//   smem  = K, V, E' of the instance (3 x N1 x 128 fp32, brought in by cp.async.bulk), the per-row
//           attention outputs, the folded local-policy tables and the bit-mask row state.
//   phase A  thread = (row, head): q . K over all unmasked nodes with a lazily rescaled online
//            softmax and the weighted V sum; K/V rows are warp-broadcast shared-memory reads.
//   phase B  warp = 4 rows: score tile (4 rows x 4 node-chunks per lane) against E'; then the local
//            policy with an octet of lanes per row (neighbour walk over the presorted list,
//            polar features, 4-head attention with the constant query folded into 3-vector dots);
//            logits = clip*tanh(score + penalty + local) + mask; argmax or Philox sampling.
//   phase C  env step on bit masks (load recurrence in fp32, visited/too-large masks, finished).
#include <string.h>
#include "common.cuh"

namespace elg {

constexpr int RW = 16;              // warps per CTA
constexpr int RT = RW * 32;         // threads per CTA
constexpr int MT_MAX = 64;          // rows per CTA
constexpr int N_RES_MAX = 112;      // nodes the resident kernel supports
constexpr int DS = 112;             // stride of the per-row dense penalty+local scratch
constexpr int TS = 36;              // padded row stride of the VPE / PE tables in smem
constexpr unsigned FULL = 0xffffffffu;

struct RolloutArgs {
  elg_tables t;
  const float* derived;
  int problem, B, M, N1, MT, tiles, k_local;
  float xi, clip;
  const int32_t* start_nodes;
  int mode;
  unsigned long long seed;
  int t_max;
  int16_t* tours;
  float* reward;
  int32_t* n_steps;
  float* logp;
  int32_t* work_counter;
  // single decode step from caller-provided state (model.one_step_rollout)
  int single_step;
  unsigned long long step_id;
  const int32_t* st_cur;
  const float* st_load;
  const int32_t* st_first;
  const uint32_t* st_mask;
  int32_t* out_selected;
  float* out_prob;
  float* out_logits;
};

// ---- shared-memory layout (offsets in floats) ---------------------------------------------------
struct SmemLayout {
  int k, v, e, o, eb, xy, dem, wl, u, tt, a, cv, vpe, pe, wct, bc, we, be;
  int cur, first, load, tlen, fin, logp, mask, vis, ctrl, bar;
  int total;     // floats
};
__host__ __device__ inline int r4(int x) { return (x + 3) & ~3; }
__host__ __device__ inline SmemLayout make_layout(int N1, int MT, int KT) {
  SmemLayout L;
  int o = 0;
  L.k = o; o += N1 * E;
  L.v = o; o += N1 * E;
  L.e = o; o += N1 * E;
  L.o = o; o += MT * E;
  L.eb = o; o += r4(N1);
  L.xy = o; o += r4(2 * N1);
  L.dem = o; o += r4(N1);
  L.wl = o; o += E;
  L.u = o; o += LH * 4;
  L.tt = o; o += LH * KT_MAX;
  L.a = o; o += LE * 4;
  L.cv = o; o += LE;
  L.vpe = o; o += KT * TS;
  L.pe = o; o += KT * TS;
  L.wct = o; o += LE * LE;
  L.bc = o; o += LE;
  L.we = o; o += LE * 4;
  L.be = o; o += LE;
  L.cur = o; o += MT;
  L.first = o; o += MT;
  L.load = o; o += MT;
  L.tlen = o; o += MT;
  L.fin = o; o += MT;
  L.logp = o; o += MT;
  L.mask = o; o += MT * 4;
  L.vis = o; o += MT * 4;
  L.ctrl = o; o += 4;
  L.bar = o; o += 4;
  L.total = o;
  return L;
}

// ---- small PTX helpers --------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// TMA bulk copy global -> shared (1D), completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// Philox4x32-10 (counter-based RNG for the sampling mode)
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  return ctr;
}

__device__ __forceinline__ float octet_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(FULL, v, 1));
  v = fmaxf(v, __shfl_xor_sync(FULL, v, 2));
  return fmaxf(v, __shfl_xor_sync(FULL, v, 4));
}
__device__ __forceinline__ float octet_sum(float v) {
  v += __shfl_xor_sync(FULL, v, 1);
  v += __shfl_xor_sync(FULL, v, 2);
  return v + __shfl_xor_sync(FULL, v, 4);
}

// =================================================================================================
template <int PROBLEM, int MAXE>
__global__ void __launch_bounds__(RT, 1) rollout_kernel(const RolloutArgs A) {
  constexpr bool CVRP = PROBLEM == ELG_CVRP;
  constexpr int DEP = CVRP ? 1 : 0;
  extern __shared__ __align__(128) float sm[];
  const int N1 = A.N1;
  const int KT = MAXE * 8;
  const SmemLayout L = make_layout(N1, A.MT, KT);
  float* sK = sm + L.k;
  float* sV = sm + L.v;
  float* sE = sm + L.e;
  float* sO = sm + L.o;
  float* sEb = sm + L.eb;
  float* sXY = sm + L.xy;
  float* sDem = sm + L.dem;
  float* sWL = sm + L.wl;
  const float* sU = sm + L.u;
  const float* sT = sm + L.tt;
  const float* sA = sm + L.a;
  const float* sCV = sm + L.cv;
  const float* sVPE = sm + L.vpe;
  const float* sPE = sm + L.pe;
  const float* sWCT = sm + L.wct;
  const float* sBC = sm + L.bc;
  const float* sWE = sm + L.we;
  const float* sBE = sm + L.be;
  int* sCur = reinterpret_cast<int*>(sm + L.cur);
  int* sFirst = reinterpret_cast<int*>(sm + L.first);
  float* sLoad = sm + L.load;
  float* sTlen = sm + L.tlen;
  int* sFin = reinterpret_cast<int*>(sm + L.fin);
  float* sLogp = sm + L.logp;
  uint32_t* sMask = reinterpret_cast<uint32_t*>(sm + L.mask);
  uint32_t* sVis = reinterpret_cast<uint32_t*>(sm + L.vis);
  int* sCtrl = reinterpret_cast<int*>(sm + L.ctrl);
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + L.bar);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- one-time: folded local-policy tables and the load column of Wq_last ---------------------
  {
    const float* loc = A.derived + DER_LOC;
    float* w = sm;
    for (int i = tid; i < E; i += RT) w[L.wl + i] = A.derived[DER_WL + i];
    for (int i = tid; i < LH * 4; i += RT) w[L.u + i] = loc[LOC_U + i];
    for (int i = tid; i < LH * KT_MAX; i += RT) w[L.tt + i] = loc[LOC_T + i];
    for (int i = tid; i < LE * 4; i += RT) { w[L.a + i] = loc[LOC_A + i]; w[L.we + i] = loc[LOC_WE + i]; }
    for (int i = tid; i < LE; i += RT) { w[L.cv + i] = loc[LOC_CV + i]; w[L.bc + i] = loc[LOC_BC + i]; w[L.be + i] = loc[LOC_BE + i]; }
    for (int i = tid; i < KT * LE; i += RT) {
      int p = i / LE, c = i % LE;
      w[L.vpe + p * TS + c] = loc[LOC_VPE + i];
      w[L.pe + p * TS + c] = loc[LOC_PE + i];
    }
    for (int i = tid; i < LE * LE; i += RT) w[L.wct + i] = loc[LOC_WCT + i];
    if (tid == 0) {
      mbar_init(bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
  }
  __syncthreads();

  const int total_work = A.B * A.tiles;
  uint32_t bar_phase = 0;
  const float inv_sqrt_le = 5.656854249492381f;   // sqrt(32), used as a divisor like the reference

  for (int iter = 0;; ++iter) {
    // ---- fetch work: dynamic (atomic counter) for rollouts, static for single decode steps ------
    int work;
    if (A.work_counter) {
      if (tid == 0) sCtrl[0] = atomicAdd(A.work_counter, 1);
      __syncthreads();
      work = sCtrl[0];
    } else {
      work = blockIdx.x + iter * gridDim.x;
    }
    if (work >= total_work) break;
    const int b = work / A.tiles, tile = work % A.tiles;
    const int row0 = tile * A.MT;
    const int nrows = min(A.MT, A.M - row0);

    // ---- stage the instance: K, V, E' by TMA bulk copies; small vectors by plain loads ---------
    if (tid == 0) {
      const uint32_t bytes = (uint32_t)N1 * E * sizeof(float);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(bar, 3 * bytes);
      bulk_g2s(sK, A.t.k + (size_t)b * N1 * E, bytes, bar);
      bulk_g2s(sV, A.t.v + (size_t)b * N1 * E, bytes, bar);
      bulk_g2s(sE, A.t.e + (size_t)b * N1 * E, bytes, bar);
    }
    for (int i = tid; i < N1; i += RT) {
      sEb[i] = A.t.eb[(size_t)b * N1 + i];
      sXY[2 * i] = A.t.xy[((size_t)b * N1 + i) * 2];
      sXY[2 * i + 1] = A.t.xy[((size_t)b * N1 + i) * 2 + 1];
      sDem[i] = CVRP ? A.t.demand[(size_t)b * N1 + i] : 0.f;
    }
    for (int r = tid; r < A.MT; r += RT) {
      const size_t g = (size_t)b * A.M + row0 + r;
      const bool ok = r < nrows;
      if (A.single_step && ok) {
        sCur[r] = A.st_cur[g];
        sFirst[r] = CVRP ? 0 : A.st_first[g];
        sLoad[r] = CVRP ? A.st_load[g] : 1.f;
#pragma unroll
        for (int w = 0; w < 4; ++w) { sMask[r * 4 + w] = A.st_mask[g * 4 + w]; sVis[r * 4 + w] = 0u; }
        sFin[r] = 0;
      } else {
        sCur[r] = 0; sFirst[r] = 0; sLoad[r] = 1.f;
#pragma unroll
        for (int w = 0; w < 4; ++w) { sMask[r * 4 + w] = 0u; sVis[r * 4 + w] = 0u; }
        sFin[r] = ok ? 0 : 1;
      }
      sTlen[r] = 0.f;
      sLogp[r] = 0.f;
    }
    mbar_wait(bar, bar_phase);
    bar_phase ^= 1;
    __syncthreads();

    // ---- the construction loop ----------------------------------------------------------------
    int t = 0;
    for (;; ++t) {
      const bool forced = !A.single_step && (t < 1 + DEP);
      if (!forced) {
        // ================= phase A: multi-head attention, thread = (row, head) ====================
        const int full_blocks = nrows >> 5, tail = nrows & 31;
        const int n_units = H * full_blocks + ((H * tail + 31) >> 5);
        for (int u = warp; u < n_units; u += RW) {
          int r, h;
          bool act;
          if (u < H * full_blocks) {
            h = u / full_blocks; r = (u % full_blocks) * 32 + lane; act = true;
          } else {
            int pi = (u - H * full_blocks) * 32 + lane;
            act = pi < H * tail;
            h = act ? pi / tail : 0;
            r = act ? full_blocks * 32 + pi % tail : 0;
          }
          act = act && !sFin[r];
          float q[D], o[D];
          uint32_t mw0 = FULL, mw1 = FULL, mw2 = FULL, mw3 = FULL;
          if (act) {
            const int cur = sCur[r];
            const float4* qp = reinterpret_cast<const float4*>(A.t.qtab + ((size_t)b * N1 + cur) * E + h * D);
            const float ld = sLoad[r];
#pragma unroll
            for (int d4 = 0; d4 < D / 4; ++d4) {
              float4 v4 = __ldg(qp + d4);
              if (CVRP) {
                float4 wl = *reinterpret_cast<const float4*>(sWL + h * D + d4 * 4);
                v4.x = fmaf(ld, wl.x, v4.x); v4.y = fmaf(ld, wl.y, v4.y);
                v4.z = fmaf(ld, wl.z, v4.z); v4.w = fmaf(ld, wl.w, v4.w);
              } else {
                float4 f4 = __ldg(reinterpret_cast<const float4*>(A.t.qfirst + ((size_t)b * N1 + sFirst[r]) * E + h * D) + d4);
                v4.x = f4.x + v4.x; v4.y = f4.y + v4.y; v4.z = f4.z + v4.z; v4.w = f4.w + v4.w;
              }
              q[d4 * 4] = v4.x; q[d4 * 4 + 1] = v4.y; q[d4 * 4 + 2] = v4.z; q[d4 * 4 + 3] = v4.w;
            }
            mw0 = sMask[r * 4]; mw1 = sMask[r * 4 + 1]; mw2 = sMask[r * 4 + 2]; mw3 = sMask[r * 4 + 3];
          } else {
#pragma unroll
            for (int d = 0; d < D; ++d) q[d] = 0.f;
          }
#pragma unroll
          for (int d = 0; d < D; ++d) o[d] = 0.f;
          float m = -INFINITY, l = 0.f;
          const float* kp = sK + h * D;
          const float* vp = sV + h * D;
          // K is pre-scaled by log2(e)/sqrt(D): scores live in the log2 domain, weights are 2^(s-m).
          // Only nodes that at least one lane of the warp still needs are visited (warp-uniform loop).
          auto key = [&](const int j, const bool use) {
            const float4* kr = reinterpret_cast<const float4*>(kp + j * E);
            float s = 0.f;
#pragma unroll
            for (int d4 = 0; d4 < D / 4; ++d4) {
              const float4 kk = kr[d4];
              s = fmaf(q[d4 * 4], kk.x, s); s = fmaf(q[d4 * 4 + 1], kk.y, s);
              s = fmaf(q[d4 * 4 + 2], kk.z, s); s = fmaf(q[d4 * 4 + 3], kk.w, s);
            }
            if (use) {
              if (s > m + 12.f) {          // lazy rescale of the running softmax reference
                const float c = exp2f(m - s);
                l *= c;
#pragma unroll
                for (int d = 0; d < D; ++d) o[d] *= c;
                m = s;
              }
              const float p = exp2f(s - m);
              l += p;
              const float4* vr = reinterpret_cast<const float4*>(vp + j * E);
#pragma unroll
              for (int d4 = 0; d4 < D / 4; ++d4) {
                const float4 vv = vr[d4];
                o[d4 * 4] = fmaf(p, vv.x, o[d4 * 4]); o[d4 * 4 + 1] = fmaf(p, vv.y, o[d4 * 4 + 1]);
                o[d4 * 4 + 2] = fmaf(p, vv.z, o[d4 * 4 + 2]); o[d4 * 4 + 3] = fmaf(p, vv.w, o[d4 * 4 + 3]);
              }
            }
          };
#define ELG_KEY_WORD(W, MW)                                                              \
          {                                                                                \
            const int nb = N1 - (W) * 32;                                                  \
            const uint32_t lim = nb >= 32 ? FULL : (nb <= 0 ? 0u : ((1u << nb) - 1u));     \
            uint32_t bits = __reduce_or_sync(FULL, ~(MW)) & lim;                           \
            while (bits) {                                                                 \
              const int jb = __ffs(bits) - 1;                                              \
              bits &= bits - 1;                                                            \
              key((W) * 32 + jb, !(((MW) >> jb) & 1u));                                    \
            }                                                                              \
          }
          ELG_KEY_WORD(0, mw0)
          ELG_KEY_WORD(1, mw1)
          ELG_KEY_WORD(2, mw2)
          ELG_KEY_WORD(3, mw3)
#undef ELG_KEY_WORD
          if (act) {
            const float inv = 1.f / l;
            float4* op = reinterpret_cast<float4*>(sO + r * E + h * D);
#pragma unroll
            for (int d4 = 0; d4 < D / 4; ++d4)
              op[d4] = make_float4(o[d4 * 4] * inv, o[d4 * 4 + 1] * inv, o[d4 * 4 + 2] * inv, o[d4 * 4 + 3] * inv);
          }
        }
        __syncthreads();
      }

      // ================= phase B + C: warp = 4 rows ==============================================
      const int r0 = warp * 4;
      int warp_live = 0;
      if (r0 < nrows) {
        const int rq = lane >> 3, s8 = lane & 7;       // octet (row within the warp), sub-lane
        const int myr = min(r0 + rq, nrows - 1);
        const bool row_ok = (r0 + rq) < nrows;
        int sel[4];
        float selp[4] = {1.f, 1.f, 1.f, 1.f};
        if (forced) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int gr = min(row0 + r0 + i, A.M - 1);
            sel[i] = (CVRP && t == 0) ? 0 : A.start_nodes[gr];
          }
        } else {
          const bool any_live = __any_sync(FULL, row_ok && !sFin[myr]);
          float acc[4][4];
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[i][c] = 0.f;
          if (any_live) {
            // ---- B2: score tile  acc[i][ch] = o[r0+i] . E'[lane + 32 ch] --------------------------
            int jj[4];
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) jj[ch] = min(lane + 32 * ch, N1 - 1);
            const int nch = (N1 + 31) >> 5;
            const float* ob = sO + r0 * E;
            const int rcl[4] = {0, min(1, nrows - 1 - r0), min(2, nrows - 1 - r0), min(3, nrows - 1 - r0)};
#pragma unroll 4
            for (int c4 = 0; c4 < E / 4; ++c4) {
              float4 ov[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) ov[i] = *reinterpret_cast<const float4*>(ob + rcl[i] * E + c4 * 4);
#pragma unroll
              for (int ch = 0; ch < 4; ++ch) {
                if (ch < nch) {
                  const float4 ev = *reinterpret_cast<const float4*>(sE + jj[ch] * E + ((c4 ^ (jj[ch] & 7)) << 2));
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    acc[i][ch] = fmaf(ov[i].x, ev.x, acc[i][ch]); acc[i][ch] = fmaf(ov[i].y, ev.y, acc[i][ch]);
                    acc[i][ch] = fmaf(ov[i].z, ev.z, acc[i][ch]); acc[i][ch] = fmaf(ov[i].w, ev.w, acc[i][ch]);
                  }
                }
              }
            }
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
              const float ebv = sEb[jj[ch]];
#pragma unroll
              for (int i = 0; i < 4; ++i) acc[i][ch] += ebv;
            }
            __syncwarp();
            // ---- B1: local policy, octet of lanes per row; scratch = this warp's (now dead) o rows ----
            uint8_t* ids = reinterpret_cast<uint8_t*>(sO + r0 * E);     // [4][64] bytes
            float* dense = sO + r0 * E + 64;                            // [4][DS]
            const bool rlive = row_ok && !sFin[myr];
            const int cur = sCur[myr];
            const float ldv = sLoad[myr];
            const float xc = sXY[2 * cur], yc = sXY[2 * cur + 1];
            // neighbour walk: first k valid entries of the presorted list of `cur`
            uint4 Lw = make_uint4(0, 0, 0, 0);
            if (rlive) Lw = __ldg(reinterpret_cast<const uint4*>(A.t.nbr + ((size_t)b * N1 + cur) * ELG_NBR_STRIDE) + s8);
            const int NL = N1 - DEP, iters = (NL + 7) >> 3, kloc = A.k_local;
            int cnt = 0;
#pragma unroll
            for (int it = 0; it < 16; ++it) {
              if (it >= iters) break;
              const uint32_t wsel = it < 4 ? Lw.x : (it < 8 ? Lw.y : (it < 12 ? Lw.z : Lw.w));
              const int id = (wsel >> ((it & 3) * 8)) & 0xff;
              const int e = it * 8 + s8;
              const bool valid = rlive && e < NL && cnt < kloc && !((sMask[myr * 4 + (id >> 5)] >> (id & 31)) & 1u);
              const uint32_t bal = __ballot_sync(FULL, valid);
              const uint32_t mine = (bal >> (lane & 24)) & 0xffu;
              const int rank = cnt + __popc(mine & ((1u << s8) - 1u));
              if (valid && rank < kloc) ids[rq * 64 + rank] = (uint8_t)id;
              cnt += __popc(mine);
              if (__all_sync(FULL, !rlive || cnt >= kloc)) break;
            }
            const int kk = min(cnt, kloc);
            const int np = rlive ? kk + DEP : 0;
            // default penalty xi everywhere (depot / neighbours are overwritten below)
#pragma unroll
            for (int i = 0; i < 4; ++i)
              for (int j = lane; j < N1; j += 32) dense[i * DS + j] = A.xi;
            __syncwarp();
            float dmax = 0.f;
            if (kk > 0) {
              const int nl = ids[rq * 64 + kk - 1];
              dmax = dist2(xc - sXY[2 * nl], yc - sXY[2 * nl + 1]);
            }
            float f0[MAXE], f1[MAXE], f2[MAXE], pen[MAXE];
            int node[MAXE];
#pragma unroll
            for (int e = 0; e < MAXE; ++e) {
              const int p = s8 + 8 * e;
              f0[e] = f1[e] = f2[e] = pen[e] = 0.f;
              node[e] = 0;
              if (p < np && !(DEP && p == 0)) {
                const int nd = ids[rq * 64 + p - DEP];
                node[e] = nd;
                const float dx = sXY[2 * nd] - xc, dy = sXY[2 * nd + 1] - yc;
                const float dd = dist2(xc - sXY[2 * nd], yc - sXY[2 * nd + 1]);
                if (CVRP) {
                  f0[e] = dmax != 0.f ? dd / (dmax + 1e-6f) : dd;
                  pen[e] = dmax != 0.f ? -(dd / dmax) : -dd;
                  f2[e] = sDem[nd] / ldv;
                } else {
                  f0[e] = dd / (dmax + 1e-6f);
                  pen[e] = -f0[e];
                }
                f1[e] = atan2f(dy, dx);
              }
            }
            // 4-head attention of the constant query over the local sequence
            float mown[4] = {0.f, 0.f, 0.f, 0.f};       // this lane's 4 rows of mh = Wo_l ol + bo_l
#pragma unroll
            for (int h = 0; h < LH; ++h) {
              const float u0 = sU[h * 4], u1 = sU[h * 4 + 1], u2 = sU[h * 4 + 2];
              float sc[MAXE];
              float mx = -INFINITY;
#pragma unroll
              for (int e = 0; e < MAXE; ++e) {
                const int p = s8 + 8 * e;
                float v = -INFINITY;
                if (p < np) {
                  v = fmaf(u2, f2[e], fmaf(u1, f1[e], u0 * f0[e])) + sT[h * KT_MAX + p];
                  if (DEP && p == 0 && (sMask[myr * 4] & 1u)) v = -INFINITY;
                }
                sc[e] = v;
                mx = fmaxf(mx, v);
              }
              mx = octet_max(mx);
              float g0 = 0.f, g1 = 0.f, g2 = 0.f, sum = 0.f;
              float vp8[LD];
#pragma unroll
              for (int d = 0; d < LD; ++d) vp8[d] = 0.f;
#pragma unroll
              for (int e = 0; e < MAXE; ++e) {
                const int p = s8 + 8 * e;
                const float w = (p < np && sc[e] != -INFINITY) ? expf(sc[e] - mx) : 0.f;
                sum += w;
                g0 = fmaf(w, f0[e], g0); g1 = fmaf(w, f1[e], g1); g2 = fmaf(w, f2[e], g2);
                if (p < np) {
                  const float4 va = *reinterpret_cast<const float4*>(sVPE + p * TS + h * LD);
                  const float4 vb = *reinterpret_cast<const float4*>(sVPE + p * TS + h * LD + 4);
                  vp8[0] = fmaf(w, va.x, vp8[0]); vp8[1] = fmaf(w, va.y, vp8[1]);
                  vp8[2] = fmaf(w, va.z, vp8[2]); vp8[3] = fmaf(w, va.w, vp8[3]);
                  vp8[4] = fmaf(w, vb.x, vp8[4]); vp8[5] = fmaf(w, vb.y, vp8[5]);
                  vp8[6] = fmaf(w, vb.z, vp8[6]); vp8[7] = fmaf(w, vb.w, vp8[7]);
                }
              }
              sum = octet_sum(sum);
              const float inv = sum > 0.f ? 1.f / sum : 0.f;
              g0 = octet_sum(g0) * inv; g1 = octet_sum(g1) * inv; g2 = octet_sum(g2) * inv;
#pragma unroll
              for (int d = 0; d < LD; ++d) {
                const float vps = octet_sum(vp8[d]) * inv;
                const int c = h * LD + d;
                // ol[c] = (Wv We)[c] . g + (Wv be)[c] + sum_p w_p (Wv PE(p))[c]
                const float ol = fmaf(sA[c * 4 + 2], g2, fmaf(sA[c * 4 + 1], g1, sA[c * 4] * g0)) + sCV[c] + vps;
                const float4 wc = *reinterpret_cast<const float4*>(sWCT + c * LE + s8 * 4);
                mown[0] = fmaf(wc.x, ol, mown[0]); mown[1] = fmaf(wc.y, ol, mown[1]);
                mown[2] = fmaf(wc.z, ol, mown[2]); mown[3] = fmaf(wc.w, ol, mown[3]);
              }
            }
            float z0 = 0.f, z1 = 0.f, z2 = 0.f, c0 = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int c = s8 * 4 + i;
              mown[i] += sBC[c];
              z0 = fmaf(sWE[c * 4], mown[i], z0); z1 = fmaf(sWE[c * 4 + 1], mown[i], z1);
              z2 = fmaf(sWE[c * 4 + 2], mown[i], z2); c0 = fmaf(sBE[c], mown[i], c0);
            }
            z0 = octet_sum(z0); z1 = octet_sum(z1); z2 = octet_sum(z2); c0 = octet_sum(c0);
            // loc_p = (We f_p + be + PE(p)) . mh / sqrt(32)
            float pem[MAXE];
#pragma unroll
            for (int e = 0; e < MAXE; ++e) pem[e] = 0.f;
#pragma unroll
            for (int src = 0; src < 8; ++src) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float mv = __shfl_sync(FULL, mown[i], (lane & 24) | src);
                const int c = src * 4 + i;
#pragma unroll
                for (int e = 0; e < MAXE; ++e) {
                  const int p = min(s8 + 8 * e, KT - 1);
                  pem[e] = fmaf(sPE[p * TS + c], mv, pem[e]);
                }
              }
            }
#pragma unroll
            for (int e = 0; e < MAXE; ++e) {
              const int p = s8 + 8 * e;
              if (p < np) {
                const float locv = (fmaf(f2[e], z2, fmaf(f1[e], z1, f0[e] * z0)) + c0 + pem[e]) / inv_sqrt_le;
                dense[rq * DS + node[e]] = pen[e] + locv;
              }
            }
            __syncwarp();
          }
          // ---- B3: logits, argmax / sampling --------------------------------------------------------
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = r0 + i;
            const bool live = r < nrows && !sFin[min(r, nrows - 1)];
            float best = -INFINITY;
            int bidx = 0x7fffffff;
            float lg[4];
            if (live) {
              const float* dn = sO + r0 * E + 64 + i * DS;
#pragma unroll
              for (int ch = 0; ch < 4; ++ch) {
                const int j = lane + 32 * ch;
                float v = -INFINITY;
                if (j < N1 && !((sMask[r * 4 + ch] >> lane) & 1u)) v = A.clip * tanhf(acc[i][ch] + dn[j]);
                lg[ch] = v;
                if (v > best) { best = v; bidx = j; }
              }
              if (A.out_logits) {
                float* lo = A.out_logits + ((size_t)b * A.M + row0 + r) * N1;
#pragma unroll
                for (int ch = 0; ch < 4; ++ch)
                  if (lane + 32 * ch < N1) lo[lane + 32 * ch] = lg[ch];
              }
            }
            // first-max argmax across the warp (ties -> lowest index, as torch.argmax)
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
              const float ov = __shfl_xor_sync(FULL, best, off);
              const int oi = __shfl_xor_sync(FULL, bidx, off);
              if (ov > best || (ov == best && oi < bidx)) { best = ov; bidx = oi; }
            }
            int choice = live ? bidx : 0;
            if (A.mode == ELG_SAMPLE && live) {
              float pr[4], psum = 0.f;
#pragma unroll
              for (int ch = 0; ch < 4; ++ch) { pr[ch] = lg[ch] == -INFINITY ? 0.f : expf(lg[ch] - best); psum += pr[ch]; }
              float tot = psum;
#pragma unroll
              for (int off = 16; off > 0; off >>= 1) tot += __shfl_xor_sync(FULL, tot, off);
              const unsigned long long grow = (unsigned long long)b * A.M + row0 + r;
              const unsigned long long stp = A.single_step ? A.step_id : (unsigned long long)t;
              const uint4 rnd = philox4x32(make_uint4((uint32_t)grow, (uint32_t)(grow >> 32), (uint32_t)stp, (uint32_t)(stp >> 32)),
                                           make_uint2((uint32_t)A.seed, (uint32_t)(A.seed >> 32)));
              const float target = ((rnd.x >> 8) + 0.5f) * (1.f / 16777216.f) * tot;
              float run = 0.f;
              int pick = -1;
              float pickp = 0.f;
#pragma unroll
              for (int ch = 0; ch < 4; ++ch) {
                float inc = pr[ch];
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                  const float nb = __shfl_up_sync(FULL, inc, off);
                  if (lane >= off) inc += nb;
                }
                const float cum = run + inc;
                const uint32_t hit = __ballot_sync(FULL, pr[ch] > 0.f && cum >= target);
                if (pick < 0 && hit) {
                  const int src = __ffs(hit) - 1;
                  pick = src + 32 * ch;
                  pickp = __shfl_sync(FULL, pr[ch], src);
                }
                run += __shfl_sync(FULL, inc, 31);
              }
              if (pick < 0) { pick = bidx; pickp = 1.f; }     // rounding fell off the end: take the mode
              choice = pick;
              selp[i] = pickp / tot;
            }
            sel[i] = choice;
          }
        }

        if (A.single_step) {
          if (lane < 4 && r0 + lane < nrows) {
            const size_t g = (size_t)b * A.M + row0 + r0 + lane;
            int sv = sel[0]; float pv = selp[0];
            if (lane == 1) { sv = sel[1]; pv = selp[1]; }
            if (lane == 2) { sv = sel[2]; pv = selp[2]; }
            if (lane == 3) { sv = sel[3]; pv = selp[3]; }
            A.out_selected[g] = sv;
            if (A.out_prob) A.out_prob[g] = pv;
          }
        } else {
          // ================= phase C: environment step (lanes = nodes for the too-large test) =======
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = r0 + i;
            if (r >= nrows) break;
            const int sl = sel[i];
            const int prev = sCur[r];
            const bool was_fin = sFin[r] != 0;
            uint32_t vw = lane < 4 ? sVis[r * 4 + lane] : 0u;
            if (lane == (sl >> 5)) vw |= 1u << (sl & 31);
            uint32_t mk;
            bool fin = was_fin;
            float ld = 1.f;
            if (CVRP) {
              const bool at_depot = sl == 0;
              ld = at_depot ? 1.f : sLoad[r] - sDem[sl];
              if (lane == 0) vw = at_depot ? (vw | 1u) : (vw & ~1u);
              uint32_t big[4];
#pragma unroll
              for (int ch = 0; ch < 4; ++ch) {
                const int j = lane + 32 * ch;
                big[ch] = __ballot_sync(FULL, j < N1 && (__fadd_rn(ld, 1e-6f) < sDem[min(j, N1 - 1)]));
              }
              const uint32_t bg = lane == 0 ? big[0] : (lane == 1 ? big[1] : (lane == 2 ? big[2] : big[3]));
              const int nb = N1 - lane * 32;
              const uint32_t fullw = nb >= 32 ? FULL : (nb <= 0 ? 0u : ((1u << nb) - 1u));
              const bool allv = __all_sync(FULL, lane >= 4 || (vw & fullw) == fullw);
              fin = was_fin || allv;
              mk = vw | bg;
              if (fin && lane == 0) mk &= ~1u;
            } else {
              mk = vw;
            }
            if (lane < 4) { sVis[r * 4 + lane] = vw; sMask[r * 4 + lane] = mk; }
            warp_live |= !fin;
            if (lane == 0) {
              if (t > 0) {
                float seg;
                if (A.t.unscaled) {
                  const float* ux = A.t.unscaled + (size_t)b * N1 * 2;
                  seg = rintf(seglen(ux[2 * prev] - ux[2 * sl], ux[2 * prev + 1] - ux[2 * sl + 1]));
                } else {
                  seg = seglen(sXY[2 * prev] - sXY[2 * sl], sXY[2 * prev + 1] - sXY[2 * sl + 1]);
                }
                sTlen[r] += seg;
              }
              if (!CVRP && t == 0) sFirst[r] = sl;
              sCur[r] = sl;
              sLoad[r] = ld;
              sFin[r] = fin ? 1 : 0;
              if (A.mode == ELG_SAMPLE && !was_fin) sLogp[r] += logf(selp[i]);
              if (t < A.t_max) A.tours[((size_t)b * A.M + row0 + r) * A.t_max + t] = (int16_t)sl;
            }
          }
        }
      }
      if (A.single_step) break;
      // ---- all rows finished?  (the barrier also publishes the new row state to phase A) ---------
      bool more = __syncthreads_or(warp_live) != 0;
      if (!CVRP) more = (t + 1) < N1;
      if (!more || t + 1 >= A.t_max) { ++t; break; }
    }

    // ---- epilogue: rewards ---------------------------------------------------------------------
    if (!A.single_step) {
      for (int r = tid; r < nrows; r += RT) {
        float len = sTlen[r];
        if (!CVRP) {                       // close the tour: last -> first
          const int a = sCur[r], f = sFirst[r];
          if (A.t.unscaled) {
            const float* ux = A.t.unscaled + (size_t)b * N1 * 2;
            len += rintf(seglen(ux[2 * a] - ux[2 * f], ux[2 * a + 1] - ux[2 * f + 1]));
          } else {
            len += seglen(sXY[2 * a] - sXY[2 * f], sXY[2 * a + 1] - sXY[2 * f + 1]);
          }
        }
        const size_t g = (size_t)b * A.M + row0 + r;
        A.reward[g] = -len;
        if (A.logp) A.logp[g] = sLogp[r];
      }
      if (tid == 0) A.n_steps[work] = t;
    }
    __syncthreads();
  }
}

// ---- host side ----------------------------------------------------------------------------------
static int pick_tiles(int M) { return (M + MT_MAX - 1) / MT_MAX; }

static int launch_rollout(const elg_model_desc* d, RolloutArgs& a, cudaStream_t st) {
  const int KT = d->local_k + (d->problem == ELG_CVRP ? 1 : 0);
  const int maxe = KT <= 32 ? 4 : (KT <= 48 ? 6 : 8);
  a.tiles = pick_tiles(a.M);
  a.MT = (((a.M + a.tiles - 1) / a.tiles) + 3) & ~3;
  const SmemLayout L = make_layout(a.N1, a.MT, maxe * 8);
  const size_t smem = (size_t)L.total * sizeof(float);
  ELG_REQUIRE(a.N1 <= N_RES_MAX, ELG_EUNSUPPORTED,
              "resident rollout kernel supports up to %d nodes (got %d); the streaming large-N path is not built yet", N_RES_MAX, a.N1);
  ELG_REQUIRE(smem <= 227 * 1024, ELG_EUNSUPPORTED, "instance does not fit in shared memory (%zu bytes for N1=%d, rows=%d)", smem, a.N1, a.MT);
  int dev = 0, sms = 148;
  ELG_CUDA_OK(cudaGetDevice(&dev));
  ELG_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int work = a.B * a.tiles;
  const int grid = work < sms ? work : sms;
#define ELG_RK(P, ME)                                                                                         \
  do {                                                                                                        \
    ELG_CUDA_OK(cudaFuncSetAttribute(rollout_kernel<P, ME>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    rollout_kernel<P, ME><<<grid, RT, smem, st>>>(a);                                                         \
  } while (0)
  if (d->problem == ELG_CVRP) {
    if (maxe == 4) ELG_RK(ELG_CVRP, 4); else if (maxe == 6) ELG_RK(ELG_CVRP, 6); else ELG_RK(ELG_CVRP, 8);
  } else {
    if (maxe == 4) ELG_RK(ELG_TSP, 4); else if (maxe == 6) ELG_RK(ELG_TSP, 6); else ELG_RK(ELG_TSP, 8);
  }
#undef ELG_RK
  ELG_LAUNCH_OK();
  return ELG_OK;
}

}  // namespace elg

using namespace elg;

extern "C" {

int elg_rollout_tiles(const elg_model_desc* d, int M, int N1) {
  (void)N1;
  if (check_desc(d) || M <= 0) return -1;
  return pick_tiles(M);
}

static int fill_common(const elg_model_desc* d, const float* derived, const elg_tables* t, int B, int M, int N1,
                       RolloutArgs& a) {
  int rc = check_desc(d);
  if (rc) return rc;
  ELG_REQUIRE(derived && t, ELG_EINVAL, "NULL pointer");
  ELG_REQUIRE(B > 0 && M > 0 && N1 > 1, ELG_EINVAL, "bad sizes B=%d M=%d N1=%d", B, M, N1);
  ELG_REQUIRE(t->xy && t->k && t->v && t->e && t->eb && t->qtab && t->nbr, ELG_EINVAL, "elg_tables has NULL members");
  ELG_REQUIRE(d->problem == ELG_TSP ? t->qfirst != nullptr : t->demand != nullptr, ELG_EINVAL,
              "tsp needs qfirst, cvrp needs demand");
  memset(&a, 0, sizeof(a));
  a.t = *t;
  a.derived = derived;
  a.problem = d->problem;
  a.B = B; a.M = M; a.N1 = N1;
  a.k_local = d->local_k;
  a.xi = d->xi; a.clip = d->clip;
  return ELG_OK;
}

int elg_rollout(const elg_model_desc* d, const float* derived, const elg_tables* t, int B, int M, int N1,
                const int32_t* start_nodes, int mode, uint64_t seed, int t_max, int16_t* tours, float* reward,
                int32_t* n_steps, float* logp, int32_t* work_counter, void* stream) {
  RolloutArgs a;
  int rc = fill_common(d, derived, t, B, M, N1, a);
  if (rc) return rc;
  ELG_REQUIRE(start_nodes && tours && reward && n_steps && work_counter, ELG_EINVAL, "NULL output/state pointer");
  ELG_REQUIRE(mode == ELG_GREEDY || mode == ELG_SAMPLE, ELG_EINVAL, "unknown mode %d", mode);
  const int need = d->problem == ELG_CVRP ? 2 * N1 + 2 : N1;
  ELG_REQUIRE(t_max >= need, ELG_EINVAL, "t_max=%d too small, need >= %d", t_max, need);
  a.start_nodes = start_nodes; a.mode = mode; a.seed = seed; a.t_max = t_max;
  a.tours = tours; a.reward = reward; a.n_steps = n_steps; a.logp = logp; a.work_counter = work_counter;
  return launch_rollout(d, a, (cudaStream_t)stream);
}

int elg_decode_step(const elg_model_desc* d, const float* derived, const elg_tables* t, int B, int M, int N1,
                    const int32_t* cur, const float* load, const int32_t* first, const uint32_t* mask_bits,
                    int mode, uint64_t seed, uint64_t step, int32_t* selected, float* prob, float* logits,
                    void* stream) {
  RolloutArgs a;
  int rc = fill_common(d, derived, t, B, M, N1, a);
  if (rc) return rc;
  ELG_REQUIRE(cur && mask_bits && selected, ELG_EINVAL, "NULL state pointer");
  ELG_REQUIRE(d->problem == ELG_TSP ? first != nullptr : load != nullptr, ELG_EINVAL, "tsp needs first, cvrp needs load");
  ELG_REQUIRE(mode == ELG_GREEDY || mode == ELG_SAMPLE, ELG_EINVAL, "unknown mode %d", mode);
  a.mode = mode; a.seed = seed; a.single_step = 1; a.step_id = step; a.t_max = 1;
  a.st_cur = cur; a.st_load = load; a.st_first = first; a.st_mask = mask_bits;
  a.out_selected = selected; a.out_prob = prob; a.out_logits = logits;
  a.work_counter = nullptr;   // static CTA -> work mapping
  return launch_rollout(d, a, (cudaStream_t)stream);
}

}  // extern "C"
