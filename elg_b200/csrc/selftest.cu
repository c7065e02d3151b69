// Diagnostic entry point: one split-precision tcgen05 GEMM  D[128][N] = A[rows][K] * B[N][K]^T
// (fp16 hi/lo operands, fp32 accumulation in TMEM), used by the test-suite to pin the descriptor /
// TMEM conventions in umma.cuh on real hardware before the rollout kernel relies on them.
#include "common.cuh"
#include "umma.cuh"

namespace elg {

__global__ void __launch_bounds__(128) umma_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                            float* __restrict__ D, int rowsA, int N, int K, int alias,
                                                            int terms) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int chunks = K / 8;
  const uint32_t lboA = (alias ? 64u : 128u) * 16u, lboB = (uint32_t)N * 16u;
  const uint32_t sizeA = chunks * lboA + (alias ? 1024u : 0u), sizeB = chunks * lboB;
  uint8_t* aHi = smem;
  uint8_t* aLo = aHi + sizeA;
  uint8_t* bHi = aLo + sizeA;
  uint8_t* bLo = bHi + sizeB;
  uint64_t* bar = reinterpret_cast<uint64_t*>(bLo + sizeB);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 1);

  for (uint32_t i = tid; i < (2 * sizeA + 2 * sizeB) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  __syncthreads();
  for (int i = tid; i < rowsA * K; i += 128) {
    int r = i / K, k = i % K;
    __half hi, lo;
    umma::split_f16(A[i], hi, lo);
    *reinterpret_cast<__half*>(aHi + umma::elem_off(r, k, lboA)) = hi;
    *reinterpret_cast<__half*>(aLo + umma::elem_off(r, k, lboA)) = lo;
  }
  for (int i = tid; i < N * K; i += 128) {
    int r = i / K, k = i % K;
    __half hi, lo;
    umma::split_f16(B[i], hi, lo);
    *reinterpret_cast<__half*>(bHi + umma::elem_off(r, k, lboB)) = hi;
    *reinterpret_cast<__half*>(bLo + umma::elem_off(r, k, lboB)) = lo;
  }
  umma::fence_async_smem();
  if (warp == 0) umma::tmem_alloc(tptr, 512);
  if (tid == 0) umma::mbar_init(bar, 1);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tbase = *tptr;

  const bool ts_mode = terms == 6;
  if (ts_mode) {
    // A operand into tensor memory: thread = lane = row; hi at columns [256, 256+K/2), lo at [320, 320+K/2)
    const int row = warp * 32 + lane;
    for (int c8 = 0; c8 < K / 16; ++c8) {
      uint32_t hw[8], lw[8];
      for (int i = 0; i < 8; ++i) {
        const int k = c8 * 16 + 2 * i;
        __half h0 = __float2half_rn(0.f), l0 = h0, h1 = h0, l1 = h0;
        if (row < rowsA) { umma::split_f16(A[(size_t)row * K + k], h0, l0); umma::split_f16(A[(size_t)row * K + k + 1], h1, l1); }
        hw[i] = umma::pack_h2(h0, h1);
        lw[i] = umma::pack_h2(l0, l1);
      }
      umma::st8(tbase + ((uint32_t)(warp * 32) << 16) + 256 + c8 * 8, hw);
      umma::st8(tbase + ((uint32_t)(warp * 32) << 16) + 320 + c8 * 8, lw);
    }
    umma::wait_st();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
  }
  if (tid == 0 && ts_mode) {
    const uint32_t idesc = umma::make_idesc_f16(128, N);
    for (int term = 0; term < 3; ++term) {
      const uint8_t* b = term == 1 ? bLo : bHi;
      for (int ks = 0; ks < K / 16; ++ks) {
        const uint64_t bd = umma::make_desc(umma::smem_addr(b) + ks * 2 * lboB, lboB, 128);
        umma::mma_f16_ts(tbase + (term ? 128u : 0u), tbase + (term == 2 ? 320u : 256u) + ks * 8, bd, idesc, !(ks == 0 && term <= 1));
      }
    }
    umma::commit(bar);
  }
  if (tid == 0 && !ts_mode) {
    const uint32_t idesc = umma::make_idesc_f16(128, N);
    // terms == 5: hi*hi into accumulator 0, the two cross terms into accumulator 1 (columns 128..), summed on read-out
    const bool split_acc = terms == 5;
    const bool cross_first = terms == 7;
    const int nterms = (split_acc || cross_first) ? 3 : terms;
    bool acc = false;
    for (int it = 0; it < nterms; ++it) {
      const int term = cross_first ? (it == 2 ? 0 : it + 1) : it;      // cross_first: hi*lo, lo*hi, then hi*hi
      if (split_acc && term == 1) acc = false;
      const uint8_t* a = term >= 2 ? aLo : aHi;
      const uint8_t* b = (term == 1 || term == 3) ? bLo : bHi;
      for (int ks = 0; ks < K / 16; ++ks) {
        const uint64_t ad = umma::make_desc(umma::smem_addr(a) + ks * 2 * lboA, lboA, 128);
        const uint64_t bd = umma::make_desc(umma::smem_addr(b) + ks * 2 * lboB, lboB, 128);
        umma::mma_f16_ss(tbase + ((split_acc && term >= 1) ? 128u : 0u), ad, bd, idesc, acc);
        acc = true;
      }
    }
    umma::commit(bar);
  }
  umma::mbar_wait(bar, 0);
  umma::fence_after_sync();
  for (int c = 0; c < N; c += 16) {
    float v[16];
    umma::ld16(tbase + ((uint32_t)(warp * 32) << 16) + c, v);
    if (terms == 5 || terms == 6) {
      float v2[16];
      umma::ld16(tbase + ((uint32_t)(warp * 32) << 16) + 128 + c, v2);
      for (int i = 0; i < 16; ++i) v[i] += v2[i];
    }
    for (int i = 0; i < 16; ++i) D[(size_t)(warp * 32 + lane) * N + c + i] = v[i];
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tbase, 512);
}

}  // namespace elg

using namespace elg;

extern "C" int elg_selftest_umma(const float* a, const float* b, float* d, int rows_a, int n, int k, int alias, int terms,
                                 void* stream) {
  ELG_REQUIRE(a && b && d, ELG_EINVAL, "NULL pointer");
  ELG_REQUIRE(n % 16 == 0 && n >= 16 && n <= 128 && k % 16 == 0 && k >= 16 && k <= 128, ELG_EINVAL, "need N,K multiples of 16 in [16,128]");
  ELG_REQUIRE(rows_a >= 1 && rows_a <= (alias ? 64 : 128) && terms >= 1 && terms <= 7, ELG_EINVAL, "bad rows/terms");
  const int chunks = k / 8;
  const size_t sizeA = (size_t)chunks * (alias ? 64 : 128) * 16 + (alias ? 1024 : 0), sizeB = (size_t)chunks * n * 16;
  const size_t smem = 2 * sizeA + 2 * sizeB + 64;
  ELG_CUDA_OK(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(a, b, d, rows_a, n, k, alias, terms);
  ELG_LAUNCH_OK();
  return ELG_OK;
}
