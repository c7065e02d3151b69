// Tensor-core rollout kernel (resident instances, greedy decoding): the same per-step path as rollout.cu
// (reference file:line list there), but the three dense contractions of the global POMO decoder
// (CVRP/models.py:330-352, TSP/models.py:252-272) all run on tcgen05 with every operand that changes per
// step living in TENSOR MEMORY:
//
//   S_h = Q_h K_h'^T        A = Q   (TMEM, written by tcgen05.st)      B = K' fp16 hi/lo (smem, resident)
//   O_h = P_h V_h           A = P_h (TMEM, written in place of S_h)    B = V^T fp16 hi/lo (smem, resident)
//   score = O E'^T          A = O   (TMEM)                             B = E' fp16 hi/lo (smem, resident)
//
// Split precision (x = hi + lo, fp16 each) with the two small cross terms issued BEFORE hi*hi into the same fp32
// accumulator keeps every contraction fp32-grade (tests/test_umma_selftest.py).
//
// CTA = one (aug-instance, tile of <= 112 POMO rows); TMEM lane = row.  16 warps: warp w owns TMEM lane quadrant
// q = w % 4 (rows 32q..32q+31) and sub-slot wsub = w / 4:
//   softmax  two independent groups of 8 warps (grp = wsub / 2): group g walks heads 4g..4g+3 one at a time through
//            its own S/P buffer, its own mbarrier and a 256-thread named barrier, so one group's tensor-core / TMEM
//            latency is covered by the other group's arithmetic.  thread = (row, key half kh = wsub & 1); the two
//            halves of a row exchange (max, sum) through shared memory + a 64-thread named barrier
//   B1       local policy, octet of lanes per row exactly as in rollout.cu (4-row tasks dealt to the groups, run
//            while the tensor core works on P.V / the next Q.K), results to shared memory ordered by node id
//   B3       thread = (row, quarter wsub of the node columns): clip*tanh(score + eb + {penalty+local | xi}) + mask,
//            first-max argmax combined across the 4 quarters through shared memory
//   C        env step: every thread of a row recomputes the scalar state, owns mask word wsub
//
// TMEM columns: [0,128) Q hi|lo, later O hi|lo   [128,256) O accumulators (8 heads x 16)
//               [256,256+2*N1p) S / P buffers of the two groups, later the score accumulator
#include "rollout_common.cuh"

namespace elg {

constexpr int TC_MT_MAX = 112;
constexpr uint32_t TC_COL_Q = 0, TC_COL_O = 128, TC_COL_S = 256;
constexpr float TC_P_SCALE_LOG2 = 10.f;      // softmax weights are 2^(s - m + 10): keeps small weights out of fp16 subnormals

struct TcLayout {
  int ops, eb, xy, dem, wl, u, tt, a, vpe, pw, pb, zw, zb;
  int cur, first, load, tlen, fin, logp, mask, vis, ids, add, nb, xch, ctrl, bar;
  int total;      // floats
};
__host__ __device__ inline int tc_r4(int x) { return (x + 3) & ~3; }
__host__ __device__ inline TcLayout make_tc_layout(int N1, int MT, int KT, int K1) {
  TcLayout L;
  const int N1p = (N1 + 15) & ~15;
  int o = 0;
  L.ops = o; o += 3 * N1p * 128;            // E' | K' | V^T, each fp16 hi + lo (N1p x 128 x 2 x 2 bytes)
  L.eb = o; o += N1p;
  L.xy = o; o += 2 * N1p;
  L.dem = o; o += N1p;
  L.wl = o; o += E;
  L.u = o; o += LH * 4;
  L.tt = o; o += LH * KT_MAX;
  L.a = o; o += LE * 4;                     // (Wv We)[c][0..2], (Wv be)[c]
  L.vpe = o; o += KT * TS;
  L.pw = o; o += KT * TS;
  L.pb = o; o += KT_MAX;
  L.zw = o; o += LE * 4;
  L.zb = o; o += 4;
  L.cur = o; o += MT;
  L.first = o; o += MT;
  L.load = o; o += MT;
  L.tlen = o; o += MT;
  L.fin = o; o += MT;
  L.logp = o; o += MT;
  L.mask = o; o += MT * 4;
  L.vis = o; o += MT * 4;
  L.ids = o; o += RW * 4 * (KT_MAX / 4);    // per octet: 64 uint8 neighbour ids
  L.add = o; o += tc_r4(MT * K1);           // penalty + local of each row's neighbours, ordered by node id
  L.nb = o; o += MT * 4;                    // neighbour bit mask per row
  L.xch = o; o += 2 * 4 * 128;              // softmax (max, sum) exchange; later the 4 argmax candidates per row
  L.ctrl = o; o += 4;
  L.bar = o; o += 12;                       // 4 mbarriers + TMEM base address
  L.total = o;
  return L;
}

__device__ __forceinline__ uint32_t pick4(const uint32_t (&w)[4], int i) {
  return i == 0 ? w[0] : (i == 1 ? w[1] : (i == 2 ? w[2] : (i == 3 ? w[3] : 0u)));
}
__device__ __forceinline__ void pair_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void group_sync(int id) { asm volatile("bar.sync %0, 256;" ::"r"(id) : "memory"); }

// =================================================================================================
template <int PROBLEM, int MAXE>
__global__ void __launch_bounds__(RT, 1) rollout_tc_kernel(const RolloutArgs A) {
  constexpr bool CVRP = PROBLEM == ELG_CVRP;
  constexpr int DEP = CVRP ? 1 : 0;
  extern __shared__ __align__(128) float sm[];
  const int N1 = A.N1;
  const int N1p = (N1 + 15) & ~15;
  const int W = (N1 + 31) >> 5;
  const int KT = MAXE * 8;
  const int K1 = A.k_local + DEP;
  const TcLayout L = make_tc_layout(N1, A.MT, KT, K1);
  const float* sEb = sm + L.eb;
  const float* sXY = sm + L.xy;
  const float* sDem = sm + L.dem;
  const float* sWL = sm + L.wl;
  const float* sU = sm + L.u;
  const float* sT = sm + L.tt;
  const float* sA = sm + L.a;
  const float* sVPE = sm + L.vpe;
  const float* sPW = sm + L.pw;
  const float* sPB = sm + L.pb;
  const float* sZW = sm + L.zw;
  const float* sZB = sm + L.zb;
  int* sCur = reinterpret_cast<int*>(sm + L.cur);
  int* sFirst = reinterpret_cast<int*>(sm + L.first);
  float* sLoad = sm + L.load;
  float* sTlen = sm + L.tlen;
  int* sFin = reinterpret_cast<int*>(sm + L.fin);
  float* sLogp = sm + L.logp;
  uint32_t* sMask = reinterpret_cast<uint32_t*>(sm + L.mask);
  uint32_t* sVis = reinterpret_cast<uint32_t*>(sm + L.vis);
  float* sAdd = sm + L.add;
  uint32_t* sNb = reinterpret_cast<uint32_t*>(sm + L.nb);
  float* sXm = sm + L.xch;              // [4][128]
  float* sXl = sm + L.xch + 512;        // [4][128]
  int* sCtrl = reinterpret_cast<int*>(sm + L.ctrl);
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + L.bar);      // TMA completion
  uint64_t* bar_sc = bar + 1;                                     // tcgen05.commit of the score MMA
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, wsub = warp >> 2, grp = wsub >> 1, kh = wsub & 1;     // softmax group (heads 4*grp..), key half
  uint64_t* bar_grp = bar + 2 + grp;                              // tcgen05.commit of this softmax group's MMAs
  const int row = q * 32 + lane;                   // TMEM lane = row of the tile
  const int rc = row < A.MT ? row : 0;             // clamped index into the per-row state arrays
  const int KH = N1p >> 1, CQ = N1p >> 2;          // keys per softmax thread, node columns per B3 thread

  {
    const float* loc = A.derived + DER_LOC;
    float* w = sm;
    for (int i = tid; i < E; i += RT) w[L.wl + i] = A.derived[DER_WL + i];
    for (int i = tid; i < LH * 4; i += RT) w[L.u + i] = loc[LOC_U + i];
    for (int i = tid; i < LH * KT_MAX; i += RT) w[L.tt + i] = loc[LOC_T + i];
    for (int i = tid; i < LE * 4; i += RT) {
      w[L.a + i] = (i & 3) == 3 ? loc[LOC_CV + (i >> 2)] : loc[LOC_A + i];
      w[L.zw + i] = loc[LOC_ZW + i];
    }
    for (int i = tid; i < KT_MAX; i += RT) w[L.pb + i] = loc[LOC_PB + i];
    for (int i = tid; i < 4; i += RT) w[L.zb + i] = loc[LOC_ZB + i];
    for (int i = tid; i < KT * LE; i += RT) {
      int p = i / LE, c = i % LE;
      w[L.vpe + p * TS + c] = loc[LOC_VPE + i];
      w[L.pw + p * TS + c] = loc[LOC_PW + i];
    }
    if (tid == 0) {
      mbar_init(bar, 1);
      mbar_init(bar + 1, 1);
      mbar_init(bar + 2, 1);
      mbar_init(bar + 3, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) umma::tmem_alloc(tmem_ptr, 512);
    umma::fence_before_sync();
  }
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tm = *tmem_ptr;
  const uint32_t tl = tm + ((uint32_t)(q * 32) << 16);          // this warp's lane quadrant

  // tensor-core operand addresses (shared memory) and instruction descriptors
  const uint32_t seg = (uint32_t)N1p * 512u;                    // bytes of one operand (hi + lo)
  const uint32_t half = (uint32_t)N1p * 256u;
  const uint32_t opE = umma::smem_addr(sm + L.ops), opK = opE + seg, opV = opK + seg;
  const uint32_t lboN = (uint32_t)N1p * 16u;                    // K' / E': N1p rows per 8-column chunk
  const uint32_t idescS = umma::make_idesc_f16(128, N1p), idescO = umma::make_idesc_f16(128, 16);
  const int nks = N1p >> 4;

  const int total_work = A.B * A.tiles;
  uint32_t bar_phase = 0, sc_phase = 0, grp_phase = 0;
  const bool leader = tid == grp * 256;                           // issues this group's MMAs
#ifdef ELG_PHASE_TIMING
  unsigned long long pclk[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif

  for (int iter = 0;; ++iter) {
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

    // ---- the instance's operands: one TMA bulk copy per operand, resident for the whole rollout ----------
    if (tid == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(bar, 3 * seg);
      const uint8_t* src = reinterpret_cast<const uint8_t*>(A.t.e) + (size_t)b * 3 * seg;
      uint8_t* dst = reinterpret_cast<uint8_t*>(sm + L.ops);
      bulk_g2s(dst, src, seg, bar);
      bulk_g2s(dst + seg, src + seg, seg, bar);
      bulk_g2s(dst + 2 * seg, src + 2 * seg, seg, bar);
    }
    for (int i = tid; i < N1p; i += RT) {
      const bool ok = i < N1;
      (sm + L.eb)[i] = ok ? A.t.eb[(size_t)b * N1 + i] : 0.f;
      (sm + L.xy)[2 * i] = ok ? A.t.xy[((size_t)b * N1 + i) * 2] : 0.f;
      (sm + L.xy)[2 * i + 1] = ok ? A.t.xy[((size_t)b * N1 + i) * 2 + 1] : 0.f;
      (sm + L.dem)[i] = (CVRP && ok) ? A.t.demand[(size_t)b * N1 + i] : 0.f;
    }
    for (int r = tid; r < A.MT; r += RT) {
      const size_t g = (size_t)b * A.M + row0 + r;
      const bool ok = r < nrows;
      if (A.single_step && ok) {
        sCur[r] = A.st_cur[g];
        sFirst[r] = CVRP ? 0 : A.st_first[g];
        sLoad[r] = CVRP ? A.st_load[g] : 1.f;
        sFin[r] = 0;
      } else {
        sCur[r] = 0; sFirst[r] = 0; sLoad[r] = 1.f;
        sFin[r] = ok ? 0 : 1;
      }
      sTlen[r] = 0.f;
      sLogp[r] = 0.f;
    }
    for (int i = tid; i < A.MT * 4; i += RT) {
      const int r = i >> 2, w = i & 3;
      sVis[i] = 0u;
      sMask[i] = (A.single_step && r < nrows && w < W) ? A.st_mask[((size_t)b * A.M + row0 + r) * W + w] : 0u;
    }
    mbar_wait(bar, bar_phase);
    bar_phase ^= 1;
    __syncthreads();

    const bool in_tile = row < nrows;
    int t = 0;
    for (;; ++t) {
      const bool forced = !A.single_step && (t < 1 + DEP);
      PHASE_T0();
      // row state at the start of the step (registers: phase C of this step overwrites the shared copies)
      const int cur0 = sCur[rc];
      const float ld0 = sLoad[rc];
      const bool act = in_tile && !sFin[rc];

      if (!forced) {
        // ---- valid-key bits of this row: unmasked and < N1 ----------------------------------------------
        uint32_t inv[4];
        {
          const uint4 m4 = *reinterpret_cast<const uint4*>(sMask + rc * 4);
          const uint32_t mw[4] = {m4.x, m4.y, m4.z, m4.w};
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            const int nb = N1 - w * 32;
            const uint32_t lim = nb >= 32 ? FULL : (nb > 0 ? ((1u << nb) - 1u) : 0u);
            inv[w] = act ? (~mw[w] & lim) : 0u;
          }
        }
        // ---- Q operand: q = Wq_last [enc[cur]; load] (cvrp) / q_first + Wq_last enc[cur] (tsp), fp16 hi/lo -> TMEM ----
        {
          uint32_t hw[16], lw[16];                     // heads 2*wsub, 2*wsub + 1: 32 consecutive k
          if (act) {
            const float4* qp = reinterpret_cast<const float4*>(A.t.qtab + ((size_t)b * N1 + cur0) * E + 2 * wsub * D);
            const float4* fp = CVRP ? nullptr : reinterpret_cast<const float4*>(A.t.qfirst + ((size_t)b * N1 + sFirst[rc]) * E + 2 * wsub * D);
#pragma unroll
            for (int d4 = 0; d4 < 2 * D / 4; ++d4) {
              float4 v4 = __ldg(qp + d4);
              if (CVRP) {
                const float4 wl = *reinterpret_cast<const float4*>(sWL + 2 * wsub * D + d4 * 4);
                v4.x = fmaf(ld0, wl.x, v4.x); v4.y = fmaf(ld0, wl.y, v4.y);
                v4.z = fmaf(ld0, wl.z, v4.z); v4.w = fmaf(ld0, wl.w, v4.w);
              } else {
                const float4 f4 = __ldg(fp + d4);
                v4.x = f4.x + v4.x; v4.y = f4.y + v4.y; v4.z = f4.z + v4.z; v4.w = f4.w + v4.w;
              }
              umma::split2_f16(v4.x, v4.y, hw[d4 * 2], lw[d4 * 2]);
              umma::split2_f16(v4.z, v4.w, hw[d4 * 2 + 1], lw[d4 * 2 + 1]);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) hw[i] = lw[i] = 0u;
          }
          umma::st16s<1>(tl + TC_COL_Q + 16 * wsub, hw);
          umma::st16s<1>(tl + TC_COL_Q + 64 + 16 * wsub, lw);
        }
        umma::wait_st();
        umma::fence_before_sync();
        __syncthreads();

        // MMA issue helpers (thread 0 only).  Per contraction: lo*hi, hi*lo, then hi*hi into one accumulator.
        auto issue_qk = [&](int head) {
          const uint32_t d = tm + TC_COL_S + grp * N1p;
          const uint32_t aHi = tm + TC_COL_Q + 8 * head, aLo = aHi + 64;
          const uint64_t bHi = umma::make_desc(opK + head * 2 * lboN, lboN, 128);
          const uint64_t bLo = umma::make_desc(opK + half + head * 2 * lboN, lboN, 128);
          umma::mma_f16_ts(d, aLo, bHi, idescS, false);
          umma::mma_f16_ts(d, aHi, bLo, idescS, true);
          umma::mma_f16_ts(d, aHi, bHi, idescS, true);
        };
        auto issue_pv = [&](int head) {
          const uint32_t d = tm + TC_COL_O + 16 * head;
          const uint32_t pHi = tm + TC_COL_S + grp * N1p, pLo = pHi + (N1p >> 1);
          const uint32_t vHi = opV + head * (N1p * 32), vLo = vHi + half;
          for (int ks = 0; ks < nks; ++ks)
            umma::mma_f16_ts(d, pLo + 8 * ks, umma::make_desc(vHi + ks * 512, 256, 128), idescO, ks > 0);
          for (int ks = 0; ks < nks; ++ks)
            umma::mma_f16_ts(d, pHi + 8 * ks, umma::make_desc(vLo + ks * 512, 256, 128), idescO, true);
          for (int ks = 0; ks < nks; ++ks)
            umma::mma_f16_ts(d, pHi + 8 * ks, umma::make_desc(vHi + ks * 512, 256, 128), idescO, true);
        };
        if (leader) {
          umma::fence_after_sync();
          issue_qk(4 * grp);
          umma::commit(bar_grp);
        }
        PHASE_MARK(0);

        // this thread's window of valid-key bits: keys [kh*KH, kh*KH + KH)
        uint32_t v0, v1;
        {
          const int s0 = kh * KH, wd = s0 >> 5, sh = s0 & 31;
          v0 = __funnelshift_r(pick4(inv, wd), pick4(inv, wd + 1), sh);
          v1 = __funnelshift_r(pick4(inv, wd + 1), pick4(inv, wd + 2), sh);
        }
        float lt0 = 0.f, lt1 = 0.f, lt2 = 0.f, lt3 = 0.f;       // softmax denominators of heads 4*grp + rho

#pragma unroll 1
        for (int rho = 0; rho < 4; ++rho) {
          mbar_wait(bar_grp, grp_phase);
          grp_phase ^= 1;
          umma::fence_after_sync();
          const uint32_t sb = tl + TC_COL_S + grp * N1p;
          uint32_t sr[56];
          // KH = 8..56 columns in at most three loads (32 + 16 + 8, register i <-> column i); columns past KH that the
          // fixed shapes drag along lie inside the allocation and are never used
          {
            const uint32_t s0 = sb + kh * KH;
            umma::ld32_nw(s0, sr);
            if (KH > 32) umma::ld16_nw(s0 + 32, sr + 32);
            if (KH > 48) umma::ld8_nw(s0 + 48, sr + 48);
          }
          umma::wait_ld();
          float mloc = -INFINITY;
#pragma unroll
          for (int c8 = 0; c8 < 7; ++c8)
            if (c8 * 8 < KH) {
#pragma unroll
              for (int i = c8 * 8; i < c8 * 8 + 8; ++i) {
                const bool valid = ((i < 32 ? (v0 >> i) : (v1 >> (i - 32))) & 1u) != 0u;
                const float s = valid ? umma::after_wait(sr[i]) : -INFINITY;
                sr[i] = __float_as_uint(s);
                mloc = fmaxf(mloc, s);
              }
            }
          const float moff = mloc == -INFINITY ? 0.f : mloc - TC_P_SCALE_LOG2;
          float lloc = 0.f;
#pragma unroll
          for (int c8 = 0; c8 < 7; ++c8)
            if (c8 * 8 < KH) {
#pragma unroll
              for (int i = c8 * 8; i < c8 * 8 + 8; ++i) {
                const float s = __uint_as_float(sr[i]);
                const float p = umma::ex2_raw(s - moff);
                lloc += p;
                sr[i] = __float_as_uint(p);
              }
            }
          // combine the two key halves of (row, head): common reference max, total denominator
          sXm[wsub * 128 + row] = mloc;
          sXl[wsub * 128 + row] = lloc;
          pair_sync(1 + grp * 4 + q);
          const float mo = sXm[(wsub ^ 1) * 128 + row], lo_ = sXl[(wsub ^ 1) * 128 + row];
          const float mm = fmaxf(mloc, mo);
          const float cown = mloc == -INFINITY ? 0.f : umma::ex2_raw(mloc - mm);
          const float coth = mo == -INFINITY ? 0.f : umma::ex2_raw(mo - mm);
          const float ltot = fmaf(lloc, cown, lo_ * coth);
          if (rho == 0) lt0 = ltot; else if (rho == 1) lt1 = ltot; else if (rho == 2) lt2 = ltot; else lt3 = ltot;
          // P (fp16 hi/lo, two keys per 32-bit column) in place of S: hi at [sb, sb + N1p/2), lo behind it
#pragma unroll
          for (int c8 = 0; c8 < 7; ++c8)
            if (c8 * 8 < KH) {
#pragma unroll
              for (int i = c8 * 4; i < c8 * 4 + 4; ++i) {       // keys (2i, 2i+1) -> hi word in sr[2i], lo word in sr[2i+1]
                uint32_t hwd, lwd;
                umma::split2_f16(__uint_as_float(sr[2 * i]) * cown, __uint_as_float(sr[2 * i + 1]) * cown, hwd, lwd);
                sr[2 * i] = hwd;
                sr[2 * i + 1] = lwd;
              }
            }
          umma::st_words<2>(sb + kh * (KH >> 1), sr, KH >> 1);
          umma::st_words<2>(sb + (N1p >> 1) + kh * (KH >> 1), sr + 1, KH >> 1);
          umma::wait_st();
          umma::fence_before_sync();
          group_sync(9 + grp);
          if (leader) {
            umma::fence_after_sync();
            issue_pv(4 * grp + rho);
            if (rho < 3) issue_qk(4 * grp + rho + 1);
            umma::commit(bar_grp);
          }
          PHASE_MARK(1);

          // ---- B1: local policy for rows [64 rho, 64 rho + 64), octet of lanes per row, while the tensor core runs ----
          if (rho < 2) {
            // 4-row tasks dealt alternately to the two groups so both carry the same local-policy load
            const int r0 = 4 * (((rho * 8 + kh * 4 + q) << 1) + grp);
            const bool own = r0 < nrows;
            const int rq = lane >> 3, s8 = lane & 7;
            const int myr = own ? min(r0 + rq, nrows - 1) : 0;
            const bool row_ok = own && (r0 + rq) < nrows;
            const bool rlive = row_ok && !sFin[myr];
            if (own && __any_sync(FULL, rlive)) {
              float addv[MAXE];
              int node[MAXE];
              uint8_t* ids = reinterpret_cast<uint8_t*>(sm + L.ids) + (warp * 4 + rq) * KT_MAX;
              const int cur = sCur[myr];
              const float ldv = sLoad[myr];
              const int NL = N1 - DEP, kloc = A.k_local;
              const uint32_t* mrow = sMask + myr * 4;
              const uint8_t* nrow = reinterpret_cast<const uint8_t*>(A.t.nbr) + ((size_t)b * N1 + cur) * ELG_NBR_NODE_BYTES(N1);
              const float2* feat = reinterpret_cast<const float2*>(nrow + ELG_NBR_STRIDE);      // (distance, angle) from cur
              int cnt = 0;
              {
                uint4 Lw = make_uint4(0, 0, 0, 0);
                if (rlive) Lw = __ldg(reinterpret_cast<const uint4*>(nrow) + s8);
                const int iters = (NL + 7) >> 3;
#pragma unroll
                for (int it = 0; it < 16; ++it) {
                  if (it >= iters) break;
                  const uint32_t wsel = it < 4 ? Lw.x : (it < 8 ? Lw.y : (it < 12 ? Lw.z : Lw.w));
                  const int id = (wsel >> ((it & 3) * 8)) & 0xff;
                  const int e = it * 8 + s8;
                  const bool valid = rlive && e < NL && cnt < kloc && !((mrow[id >> 5] >> (id & 31)) & 1u);
                  const uint32_t bal = __ballot_sync(FULL, valid);
                  const uint32_t mine = (bal >> (lane & 24)) & 0xffu;
                  const int rank = cnt + __popc(mine & ((1u << s8) - 1u));
                  if (valid && rank < kloc) ids[rank] = (uint8_t)id;
                  cnt += __popc(mine);
                  if (__all_sync(FULL, !rlive || cnt >= kloc)) break;
                }
              }
              const int kk = min(cnt, kloc);
              const int np = rlive ? kk + DEP : 0;
              __syncwarp();
              // features of this lane's entries: gathered from the per-instance pair table (same dist2 / atan2f values the
              // kernel used to recompute); the three divisions become one reciprocal each per row
              float2 ft[MAXE];
#pragma unroll
              for (int e = 0; e < MAXE; ++e) {
                const int p = s8 + 8 * e;
                node[e] = 0;
                ft[e] = make_float2(0.f, 0.f);
                if (p < np && !(DEP && p == 0)) {
                  node[e] = ids[p - DEP];
                  ft[e] = __ldg(feat + node[e]);
                }
              }
              float dmax = 0.f;
              if (kk > 0) dmax = __ldg(feat + ids[kk - 1]).x;
              const float r0d = dmax != 0.f ? 1.f / (dmax + 1e-6f) : 1.f;      // cvrp: cur_dist / (max + 1e-6); dmax == 0 -> dd itself
              const float r1d = dmax != 0.f ? 1.f / dmax : 1.f;                 // penalty -d / dmax (no eps, CVRP/models.py:380,403)
              const float rtsp = 1.f / (dmax + 1e-6f);
              const float rld = 1.f / ldv;
              float f0[MAXE], f1[MAXE], f2[MAXE];
#pragma unroll
              for (int e = 0; e < MAXE; ++e) {
                const int p = s8 + 8 * e;
                f0[e] = f1[e] = f2[e] = addv[e] = 0.f;
                if (p < np && !(DEP && p == 0)) {
                  const float dd = ft[e].x;
                  if (CVRP) {
                    f0[e] = dd * r0d;
                    addv[e] = -(dd * r1d);
                    f2[e] = sDem[node[e]] * rld;
                  } else {
                    f0[e] = dd * rtsp;
                    addv[e] = -f0[e];
                  }
                  f1[e] = ft[e].y;
                }
              }
              // 4-head attention of the constant query over the local sequence.  Per head the octet reduces
              // (sum, g0..2) to every lane and the 8 value columns reduce-scatter, so lane d ends up owning ol[h*8 + d].
              float olh[LH];
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
                    if (DEP && p == 0 && (mrow[0] & 1u)) v = -INFINITY;
                  }
                  sc[e] = v;
                  mx = fmaxf(mx, v);
                }
                mx = octet_max(mx);
                const float mref = mx == -INFINITY ? 0.f : mx;
                float g0 = 0.f, g1 = 0.f, g2 = 0.f, sum = 0.f;
                float vp8[LD];
#pragma unroll
                for (int d = 0; d < LD; ++d) vp8[d] = 0.f;
#pragma unroll
                for (int e = 0; e < MAXE; ++e) {
                  const int p = s8 + 8 * e;
                  const float w = umma::ex2_raw(sc[e] - mref);          // -inf (masked / beyond np) -> 0
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
                const float inv_s = sum > 0.f ? 1.f / sum : 0.f;
                g0 = octet_sum(g0) * inv_s; g1 = octet_sum(g1) * inv_s; g2 = octet_sum(g2) * inv_s;
                // reduce-scatter of vp8 over the octet: after three exchange steps lane s8 holds the total of vp8[s8]
                float r4[4], r2[2];
                {
                  const bool up = (s8 & 4) != 0;
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    const float send = up ? vp8[i] : vp8[i + 4];
                    const float keep = up ? vp8[i + 4] : vp8[i];
                    r4[i] = keep + __shfl_xor_sync(FULL, send, 4);
                  }
                }
                {
                  const bool up = (s8 & 2) != 0;
#pragma unroll
                  for (int i = 0; i < 2; ++i) {
                    const float send = up ? r4[i] : r4[i + 2];
                    const float keep = up ? r4[i + 2] : r4[i];
                    r2[i] = keep + __shfl_xor_sync(FULL, send, 2);
                  }
                }
                const bool up1 = (s8 & 1) != 0;
                const float vps = ((up1 ? r2[1] : r2[0]) + __shfl_xor_sync(FULL, up1 ? r2[0] : r2[1], 1)) * inv_s;
                // ol[c] = (Wv We)[c] . g + (Wv be)[c] + sum_p w_p (Wv PE(p))[c],  c = h*8 + s8
                const float4 a4 = *reinterpret_cast<const float4*>(sA + (h * LD + s8) * 4);
                olh[h] = fmaf(a4.z, g2, fmaf(a4.y, g1, a4.x * g0)) + a4.w + vps;
              }
              // z = ZW^T ol + ZB (3 feature weights + constant), partial over this lane's four ol, then octet sum
              float z0 = 0.f, z1 = 0.f, z2 = 0.f, c0 = 0.f;
#pragma unroll
              for (int h = 0; h < LH; ++h) {
                const float4 zw = *reinterpret_cast<const float4*>(sZW + (h * LD + s8) * 4);
                z0 = fmaf(zw.x, olh[h], z0); z1 = fmaf(zw.y, olh[h], z1);
                z2 = fmaf(zw.z, olh[h], z2); c0 = fmaf(zw.w, olh[h], c0);
              }
              z0 = octet_sum(z0) + sZB[0]; z1 = octet_sum(z1) + sZB[1]; z2 = octet_sum(z2) + sZB[2]; c0 = octet_sum(c0) + sZB[3];
              // positional part PW[p] . ol of this lane's entries: ol broadcast four columns at a time
              float pem[MAXE];
#pragma unroll
              for (int e = 0; e < MAXE; ++e) pem[e] = sPB[min(s8 + 8 * e, KT - 1)];
#pragma unroll
              for (int h = 0; h < LH; ++h) {
#pragma unroll
                for (int dq = 0; dq < 2; ++dq) {
                  const float o0 = __shfl_sync(FULL, olh[h], (lane & 24) | (dq * 4 + 0));
                  const float o1 = __shfl_sync(FULL, olh[h], (lane & 24) | (dq * 4 + 1));
                  const float o2 = __shfl_sync(FULL, olh[h], (lane & 24) | (dq * 4 + 2));
                  const float o3 = __shfl_sync(FULL, olh[h], (lane & 24) | (dq * 4 + 3));
#pragma unroll
                  for (int e = 0; e < MAXE; ++e) {
                    const int p = min(s8 + 8 * e, KT - 1);
                    const float4 pw = *reinterpret_cast<const float4*>(sPW + p * TS + h * LD + dq * 4);
                    pem[e] = fmaf(pw.w, o3, fmaf(pw.z, o2, fmaf(pw.y, o1, fmaf(pw.x, o0, pem[e]))));
                  }
                }
              }
#pragma unroll
              for (int e = 0; e < MAXE; ++e)      // penalty + local score (tables carry the 1/sqrt(32))
                addv[e] += fmaf(f2[e], z2, fmaf(f1[e], z1, f0[e] * z0)) + c0 + pem[e];

              // publish: neighbour bit mask of the row, and penalty + local ordered by node id
              uint32_t nbw[4] = {0u, 0u, 0u, 0u};
#pragma unroll
              for (int e = 0; e < MAXE; ++e) {
                const int p = s8 + 8 * e;
                if (p < np) {
                  const int nd = node[e];
                  const uint32_t bit = 1u << (nd & 31);
                  nbw[0] |= (nd >> 5) == 0 ? bit : 0u; nbw[1] |= (nd >> 5) == 1 ? bit : 0u;
                  nbw[2] |= (nd >> 5) == 2 ? bit : 0u; nbw[3] |= (nd >> 5) == 3 ? bit : 0u;
                }
              }
#pragma unroll
              for (int w = 0; w < 4; ++w) {
                nbw[w] |= __shfl_xor_sync(FULL, nbw[w], 1);
                nbw[w] |= __shfl_xor_sync(FULL, nbw[w], 2);
                nbw[w] |= __shfl_xor_sync(FULL, nbw[w], 4);
              }
              const int pc1 = __popc(nbw[0]), pc2 = pc1 + __popc(nbw[1]), pc3 = pc2 + __popc(nbw[2]);
#pragma unroll
              for (int e = 0; e < MAXE; ++e) {
                const int p = s8 + 8 * e;
                if (p < np) {
                  const int nd = node[e], wd = nd >> 5;
                  const uint32_t below = pick4(nbw, wd) & ((1u << (nd & 31)) - 1u);
                  const int rank = (wd == 0 ? 0 : (wd == 1 ? pc1 : (wd == 2 ? pc2 : pc3))) + __popc(below);
                  sAdd[myr * K1 + rank] = addv[e];
                }
              }
              if (row_ok && s8 < 4) sNb[myr * 4 + s8] = pick4(nbw, s8);
            }
            PHASE_MARK(2);
          }
        }

        // ---- O = P V accumulated: normalise, fp16 hi/lo -> O operand (TMEM, over the dead Q operand) ------------
        mbar_wait(bar_grp, grp_phase);
        grp_phase ^= 1;
        umma::fence_after_sync();
        {
          uint32_t orr[32], hw[16], lw[16];            // heads 4*grp + 2*kh, + 1: 32 consecutive accumulator columns
          umma::ld32_nw(tl + TC_COL_O + 16 * (4 * grp + 2 * kh), orr);
          umma::wait_ld();
#pragma unroll
          for (int i2 = 0; i2 < 2; ++i2) {
            const float lsel = kh ? (i2 ? lt3 : lt2) : (i2 ? lt1 : lt0);
            const float inv_l = act ? 1.f / lsel : 0.f;
#pragma unroll
            for (int d2 = 0; d2 < 8; ++d2)
              umma::split2_f16(act ? umma::after_wait(orr[i2 * 16 + 2 * d2]) * inv_l : 0.f,
                               act ? umma::after_wait(orr[i2 * 16 + 2 * d2 + 1]) * inv_l : 0.f, hw[i2 * 8 + d2], lw[i2 * 8 + d2]);
          }
          umma::st16s<1>(tl + TC_COL_Q + 8 * (4 * grp + 2 * kh), hw);
          umma::st16s<1>(tl + TC_COL_Q + 64 + 8 * (4 * grp + 2 * kh), lw);
          umma::wait_st();
          umma::fence_before_sync();
        }
        __syncthreads();
        if (tid == 0) {
          // scores = O E'^T (the reference's single-head matmul(mh_atten_out, single_head_key), Wo folded into E')
          umma::fence_after_sync();
          const uint32_t d = tm + TC_COL_S;
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            umma::mma_f16_ts(d, tm + TC_COL_Q + 64 + 8 * ks, umma::make_desc(opE + ks * 2 * lboN, lboN, 128), idescS, ks > 0);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            umma::mma_f16_ts(d, tm + TC_COL_Q + 8 * ks, umma::make_desc(opE + half + ks * 2 * lboN, lboN, 128), idescS, true);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            umma::mma_f16_ts(d, tm + TC_COL_Q + 8 * ks, umma::make_desc(opE + ks * 2 * lboN, lboN, 128), idescS, true);
          umma::commit(bar_sc);
        }
        PHASE_MARK(3);

        // ---- B3: logits of this thread's node columns [wsub*CQ, wsub*CQ + CQ) -----------------------------------
        mbar_wait(bar_sc, sc_phase);
        sc_phase ^= 1;
        umma::fence_after_sync();
        PHASE_MARK(7);
        {
          uint32_t xr[28];
          const int c0n = wsub * CQ;
          {
            const uint32_t s0 = tl + TC_COL_S + c0n;
            umma::ld16_nw(s0, xr);
            if (CQ > 16) umma::ld8_nw(s0 + 16, xr + 16);
            if (CQ > 24) umma::ld4_nw(s0 + 24, xr + 24);
          }
          umma::wait_ld();
          float best = -INFINITY;
          int bidx = 0x7fffffff;
          bool need_all = false;
          if (act) {
            const uint4 n4 = *reinterpret_cast<const uint4*>(sNb + rc * 4);
            const uint32_t nbw[4] = {n4.x, n4.y, n4.z, n4.w};
            const int wd = c0n >> 5, sh = c0n & 31;
            const uint32_t vwin = __funnelshift_r(pick4(inv, wd), pick4(inv, wd + 1), sh);      // unmasked & < N1
            const uint32_t nwin = __funnelshift_r(pick4(nbw, wd), pick4(nbw, wd + 1), sh);      // neighbour (or depot)
            int rbase = __popc(pick4(nbw, wd) & ((1u << sh) - 1u));
            rbase += wd > 0 ? __popc(nbw[0]) : 0;
            rbase += wd > 1 ? __popc(nbw[1]) : 0;
            rbase += wd > 2 ? __popc(nbw[2]) : 0;
            const float* arow = sAdd + rc * K1 + rbase;
            float* lo = A.out_logits ? A.out_logits + ((size_t)b * A.M + row0 + row) * N1 : nullptr;
            // pass 1: pre-activation x = score + eb + {penalty + local | xi}; -inf where masked; first maximum
            float xmax = -INFINITY;
            int xidx = 0x7fffffff;
#pragma unroll
            for (int c4 = 0; c4 < 7; ++c4)
              if (c4 * 4 < CQ) {
#pragma unroll
                for (int i = c4 * 4; i < c4 * 4 + 4; ++i) {
                  float x = -INFINITY;
                  if ((vwin >> i) & 1u) {
                    const bool isnb = (nwin >> i) & 1u;
                    const float add = isnb ? arow[__popc(nwin & ((1u << i) - 1u))] : A.xi;
                    x = (umma::after_wait(xr[i]) + sEb[c0n + i]) + add;
                    if (x > xmax) { xmax = x; xidx = c0n + i; }
                  }
                  xr[i] = __float_as_uint(x);
                }
              }
            // tanh is monotone, so the arg-max of clip*tanh(x) is the arg-max of x unless fp32 tanh maps a smaller x to
            // the same (or, by a 1-2 ulp wobble, larger) logit.  |d tanh| >= 4 e^(-2|x|) |dx| near the maximum, so only
            // pre-activations within w = max(1e-4, 5e-7 e^(2|xmax|)) of it can tie (w covers ~4 ulp of tanh); in the
            // saturated range that is a wide window, otherwise just the maximum -> a single tanh per row.
            const float thr = xmax - fmaxf(1e-4f, 5e-7f * __expf(2.f * fabsf(xmax)));
            int ncand = 0;
#pragma unroll
            for (int c4 = 0; c4 < 7; ++c4)
              if (c4 * 4 < CQ) {
#pragma unroll
                for (int i = c4 * 4; i < c4 * 4 + 4; ++i) ncand += (__uint_as_float(xr[i]) >= thr && __uint_as_float(xr[i]) != -INFINITY) ? 1 : 0;
              }
            need_all = ncand > 1 || lo != nullptr;
            if (ncand == 1) { best = A.clip * tanhf(xmax); bidx = xidx; }
          }
          const bool slow = __any_sync(FULL, need_all);
          if (slow) {        // rare (ties / saturation) or diagnostic (logits requested): every logit
            float* lo = (act && A.out_logits) ? A.out_logits + ((size_t)b * A.M + row0 + row) * N1 : nullptr;
            if (need_all) { best = -INFINITY; bidx = 0x7fffffff; }
#pragma unroll
            for (int c4 = 0; c4 < 7; ++c4)
              if (c4 * 4 < CQ) {
#pragma unroll
                for (int i = c4 * 4; i < c4 * 4 + 4; ++i) {
                  const int j = c0n + i;
                  const float x = __uint_as_float(xr[i]);
                  const float v = (need_all && x != -INFINITY) ? A.clip * tanhf(x) : -INFINITY;
                  if (need_all && v > best) { best = v; bidx = j; }
                  if (lo && j < N1) lo[j] = v;
                }
              }
          }
          umma::fence_before_sync();
          sXm[wsub * 128 + row] = best;
          reinterpret_cast<int*>(sXl)[wsub * 128 + row] = bidx;
        }
        PHASE_MARK(4);
      }
      __syncthreads();

      // ---- selection: first-max over the 4 column quarters (ties -> lowest index, as torch.argmax) ----------------
      int sl;
      if (forced) {
        sl = (CVRP && t == 0) ? 0 : A.start_nodes[min(row0 + row, A.M - 1)];
      } else {
        float bv = -INFINITY;
        int bi = 0x7fffffff;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float ov = sXm[c * 128 + row];
          const int oi = reinterpret_cast<const int*>(sXl)[c * 128 + row];
          if (ov > bv) { bv = ov; bi = oi; }
        }
        sl = act ? bi : 0;
      }
      bool live_after = false;
      if (A.single_step) {
        if (wsub == 0 && in_tile) {
          const size_t g = (size_t)b * A.M + row0 + row;
          A.out_selected[g] = sl;
          if (A.out_prob) A.out_prob[g] = 1.f;
        }
      } else if (in_tile) {
        // ================= phase C: environment step; this thread owns mask word `wsub` of its row =============
        const bool was_fin = !act;
        bool fin = was_fin;
        float ld = 1.f;
        const uint4 v4 = *reinterpret_cast<const uint4*>(sVis + rc * 4);
        uint32_t vw[4] = {v4.x, v4.y, v4.z, v4.w};
        const uint32_t sbit = 1u << (sl & 31);
        vw[0] |= (sl >> 5) == 0 ? sbit : 0u; vw[1] |= (sl >> 5) == 1 ? sbit : 0u;
        vw[2] |= (sl >> 5) == 2 ? sbit : 0u; vw[3] |= (sl >> 5) == 3 ? sbit : 0u;
        if (CVRP) {
          const bool at_depot = sl == 0;
          ld = at_depot ? 1.f : ld0 - sDem[sl];
          vw[0] = at_depot ? (vw[0] | 1u) : (vw[0] & ~1u);
          bool allv = true;
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            const int nb = N1 - w * 32;
            if (nb > 0) {
              const uint32_t fullw = nb >= 32 ? FULL : ((1u << nb) - 1u);
              allv = allv && ((vw[w] & fullw) == fullw);
            }
          }
          fin = was_fin || allv;
          if (wsub < W) {
            uint32_t big = 0u;
            const float lde = __fadd_rn(ld, 1e-6f);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int j = wsub * 32 + i;
              if (j < N1 && lde < sDem[j]) big |= 1u << i;
            }
            const uint32_t mine = pick4(vw, wsub);
            uint32_t mk = mine | big;
            if (wsub == 0 && fin) mk &= ~1u;               // finished rows may stay at the depot
            sVis[rc * 4 + wsub] = mine;
            sMask[rc * 4 + wsub] = mk;
          }
        } else {
          if (wsub < W) {
            const uint32_t mine = pick4(vw, wsub);
            sVis[rc * 4 + wsub] = mine;
            sMask[rc * 4 + wsub] = mine;
          }
        }
        live_after = !fin;
        if (wsub == 0) {
          if (t > 0) {
            float sg;
            if (A.t.unscaled) {
              const float* ux = A.t.unscaled + (size_t)b * N1 * 2;
              sg = rintf(seglen(ux[2 * cur0] - ux[2 * sl], ux[2 * cur0 + 1] - ux[2 * sl + 1]));
            } else {
              sg = seglen(sXY[2 * cur0] - sXY[2 * sl], sXY[2 * cur0 + 1] - sXY[2 * sl + 1]);
            }
            sTlen[rc] += sg;
          }
          if (!CVRP && t == 0) sFirst[rc] = sl;
          sCur[rc] = sl;
          sLoad[rc] = ld;
          sFin[rc] = fin ? 1 : 0;
          if (t < A.t_max) A.tours[((size_t)b * A.M + row0 + row) * A.t_max + t] = (int16_t)sl;
        }
      }
      PHASE_MARK(5);
      if (A.single_step) break;
      bool more = __syncthreads_or(live_after ? 1 : 0) != 0;
      PHASE_MARK(6);
      if (!CVRP) more = (t + 1) < N1;
      if (!more || t + 1 >= A.t_max) { ++t; break; }
    }

    // ---- epilogue: rewards ---------------------------------------------------------------------
    if (!A.single_step) {
      for (int r = tid; r < nrows; r += RT) {
        float len = sTlen[r];
        if (!CVRP) {
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
      if (tid == 0) A.n_steps[b * A.ns_stride + tile] = t;
    }
    __syncthreads();
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tm, 512);
#ifdef ELG_PHASE_TIMING
  if (tid == 0)
    for (int i = 0; i < 8; ++i) atomicAdd(&g_phase_clk[i], pclk[i]);
#endif
}

// ---- host side ----------------------------------------------------------------------------------
static int tc_maxe(const elg_model_desc* d) {
  const int KT = d->local_k + (d->problem == ELG_CVRP ? 1 : 0);
  return KT <= 32 ? 4 : (KT <= 48 ? 6 : 8);
}

// Largest row tile (multiple of 4, <= 112) whose layout fits the 227 KB of one SM; 0 = none
int rollout_tc_max_rows(const elg_model_desc* d, int N1) {
  if (N1 > N_RES_MAX) return 0;
  const int K1 = d->local_k + (d->problem == ELG_CVRP ? 1 : 0);
  for (int mt = TC_MT_MAX; mt >= 4; mt -= 4)
    if ((size_t)make_tc_layout(N1, mt, tc_maxe(d) * 8, K1).total * sizeof(float) <= 227 * 1024) return mt;
  return 0;
}

// Tiles per aug-instance the tensor-core kernel would use for (B, M, N1); 0 = the shape is not eligible
int rollout_tc_tiles(const elg_model_desc* d, int B, int M, int N1, int* mt_out) {
  const int cap = rollout_tc_max_rows(d, N1);
  if (cap == 0) return 0;
  int tiles = (M + cap - 1) / cap;
  int mt = (((M + tiles - 1) / tiles) + 3) & ~3;
  if (mt > cap) { mt = cap; tiles = (M + mt - 1) / mt; }
  if (mt_out) *mt_out = mt;
  (void)B;
  return tiles;
}

int launch_rollout_tc(const elg_model_desc* d, RolloutArgs& a, cudaStream_t st) {
  int mt = 0;
  const int tiles = rollout_tc_tiles(d, a.B, a.M, a.N1, &mt);
  ELG_REQUIRE(tiles > 0, ELG_EUNSUPPORTED, "tensor-core rollout does not fit (N1=%d)", a.N1);
  ELG_REQUIRE(a.mode == ELG_GREEDY, ELG_EUNSUPPORTED, "tensor-core rollout is greedy-only");
  a.tiles = tiles;
  a.MT = mt;
  const int maxe = tc_maxe(d);
  const int K1 = d->local_k + (d->problem == ELG_CVRP ? 1 : 0);
  const size_t smem = (size_t)make_tc_layout(a.N1, mt, maxe * 8, K1).total * sizeof(float);
  int dev = 0, sms = 148;
  ELG_CUDA_OK(cudaGetDevice(&dev));
  ELG_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int work = a.B * a.tiles;
  const int grid = work < sms ? work : sms;
#define ELG_TK(P, ME)                                                                                              \
  do {                                                                                                             \
    ELG_CUDA_OK(cudaFuncSetAttribute(rollout_tc_kernel<P, ME>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    rollout_tc_kernel<P, ME><<<grid, RT, smem, st>>>(a);                                                           \
  } while (0)
  if (d->problem == ELG_CVRP) {
    if (maxe == 4) ELG_TK(ELG_CVRP, 4); else if (maxe == 6) ELG_TK(ELG_CVRP, 6); else ELG_TK(ELG_CVRP, 8);
  } else {
    if (maxe == 4) ELG_TK(ELG_TSP, 4); else if (maxe == 6) ELG_TK(ELG_TSP, 6); else ELG_TK(ELG_TSP, 8);
  }
#undef ELG_TK
  ELG_LAUNCH_OK();
  return ELG_OK;
}

}  // namespace elg

#ifdef ELG_PHASE_TIMING
extern "C" int elg_debug_phase_clocks_tc(unsigned long long* out8, int reset) {
  ELG_CUDA_OK(cudaMemcpyFromSymbol(out8, elg::g_phase_clk, sizeof(unsigned long long) * 8));
  if (reset) { unsigned long long z[8] = {0}; ELG_CUDA_OK(cudaMemcpyToSymbol(elg::g_phase_clk, z, sizeof(z))); }
  return ELG_OK;
}
#endif
