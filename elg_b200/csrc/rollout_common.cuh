// Pieces shared by the rollout kernels (rollout.cu: fp32-pipe attention, resident + streaming;
// rollout_tc.cu: tensor-core attention, resident only).
#pragma once
#include <string.h>
#include "common.cuh"
#include "umma.cuh"

namespace elg {

constexpr int RW = 16;              // warps per CTA
constexpr int RT = RW * 32;         // threads per CTA
constexpr int MT_MAX = 64;          // rows per CTA
constexpr int N_RES_MAX = 112;      // nodes the resident variant supports
constexpr int N_STREAM_MAX = 8192;  // nodes the streaming variant supports (uint16 ids, sort kernel)
constexpr int DS = 128;             // stride of the per-row dense penalty+local scratch (one node chunk)
constexpr int TS = 36;              // padded row stride of the VPE / PE tables in smem
constexpr int SS = 116;             // resident: row stride of the staged scores (112 scores + 4 neighbour-mask words)
constexpr int A_HALF = 16384;       // resident: bytes of one fp16 A operand (64 rows x 128 k, K-major core matrices)
constexpr unsigned FULL = 0xffffffffu;

struct RolloutArgs {
  elg_tables t;
  const float* derived;
  int problem, B, M, N1, MT, tiles, k_local;
  int ns_stride;      // n_steps is indexed [aug-instance * ns_stride + tile] (ns_stride = elg_rollout_tiles())
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

// ---- streamed tensor-core rollout (rollout_stc.cu): tiles of elg_tables.et and the scratch elg_tables.ws ------------
constexpr int STC_TILE = ELG_TILE_NODES;       // nodes per tile
__host__ __device__ inline int stc_tiles(int N1) { return (N1 + STC_TILE - 1) / STC_TILE; }
struct StcWs {
  size_t mask, vis, nb, total;          // byte offsets: three bit masks [rows][Wp] uint32
  int Wp, NP;                           // mask words per row (4 per tile), padded node count
};
__host__ __device__ inline StcWs stc_ws_layout(long long rows, int N1) {
  StcWs w;
  w.Wp = stc_tiles(N1) * 4;
  w.NP = stc_tiles(N1) * STC_TILE;
  const size_t mb = (size_t)rows * w.Wp * 4;
  w.mask = 0; w.vis = mb; w.nb = 2 * mb;
  w.total = 3 * mb;
  return w;
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

// Optional per-phase cycle accounting (tools/phase_timing.py builds with -DELG_PHASE_TIMING; never in the shipped library)
#ifdef ELG_PHASE_TIMING
static __device__ unsigned long long g_phase_clk[16];   // one copy per translation unit
#define PHASE_T0() long long pt_ = clock64()
#define PHASE_MARK(i) do { if (threadIdx.x == 0) { const long long n_ = clock64(); pclk[i] += (unsigned long long)(n_ - pt_); pt_ = n_; } } while (0)
#else
#define PHASE_T0()
#define PHASE_MARK(i)
#endif


}  // namespace elg
