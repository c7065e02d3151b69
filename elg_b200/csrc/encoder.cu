// Encoder (model.pre_forward) and the decoder-side tables derived from the encoded nodes.
//
// reference: CVRP_Encoder / EncoderLayer / AddAndInstanceNormalization / FeedForward
//            (CVRP/models.py:199-269,506-562), TSP twins (TSP/models.py:134-194,387-423),
//            decoder.set_kv (CVRP/models.py:300-308, TSP/models.py:227-235).
//
// All encoder linears are batched over the B*N1 node rows of the whole batch so the 0.8 M
// weights are read once per batch (weight-stationary tiles), fp32 FMA accumulation (TF32/BF16
// inputs would break the 1e-4 logit tolerance).  Per layer:
//   qkv  = x [Wq;Wk;Wv]^T                     sgemm, N = 384
//   att  = softmax(q k^T / 4) v               per (aug-instance, head), K/V staged in smem
//   t    = att Wo^T + bo + x                  sgemm + bias + residual epilogue
//   x1   = instance_norm(t) * w + b           per aug-instance, stats over nodes
//   hid  = relu(x1 W1^T + b1)                 sgemm + bias + relu epilogue
//   t    = hid W2^T + b2 + x1                 sgemm + bias + residual epilogue
//   x    = instance_norm(t) * w + b
#include <cuda_fp16.h>
#include "common.cuh"
#include "umma.cuh"

namespace elg {

bool rollout_is_resident(const elg_model_desc* d, int N1);
int launch_neighbours(const elg_model_desc* d, const float* xy, const float* demand, int B, int N1, void* nbr, cudaStream_t stream);

// ---- embedding --------------------------------------------------------------------------------
__global__ void embed_kernel(int problem, const float* __restrict__ xy, const float* __restrict__ demand,
                             const float* __restrict__ w, elg_weight_layout_t L, long long rows, int N1,
                             float* __restrict__ x) {
  // thread = (node row, 4 channels): one 16-byte store; the row -> node index modulo is paid once per four outputs
  const long long total = rows * (E / 4);
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c0 = (int)(idx & (E / 4 - 1)) * 4;
    const long long r = idx / (E / 4);
    const int j = (int)(r % N1);
    const float2 p = *reinterpret_cast<const float2*>(xy + 2 * r);
    const float px = p.x, py = p.y;
    float v[4];
    if (problem == ELG_CVRP) {
      if (j == 0) {
        // F.linear: x0*w0 + x1*w1 then + bias
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c = c0 + i;
          v[i] = fmaf(py, w[L.emb_depot_w + c * 2 + 1], px * w[L.emb_depot_w + c * 2]) + w[L.emb_depot_b + c];
        }
      } else {
        const float dm = demand[r];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c = c0 + i;
          v[i] = fmaf(dm, w[L.emb_node_w + c * 3 + 2], fmaf(py, w[L.emb_node_w + c * 3 + 1], px * w[L.emb_node_w + c * 3])) +
                 w[L.emb_node_b + c];
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = c0 + i;
        v[i] = fmaf(py, w[L.emb_node_w + c * 2 + 1], px * w[L.emb_node_w + c * 2]) + w[L.emb_node_b + c];
      }
    }
    *reinterpret_cast<float4*>(x + r * E + c0) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

// ---- SGEMM  C[M][N] = A[M][K] * B[N][K]^T (+ epilogue) ----------------------------------------
// 128x128x8 tiles, 256 threads, 8x8 register tile per thread, register-prefetched double buffer.
enum { EPI_NONE = 0, EPI_BIAS = 1, EPI_BIAS_RELU = 2, EPI_BIAS_RES = 3, EPI_SWIZZLE = 4, EPI_UMMA_SPLIT = 5 };

template <int EPI>
__global__ void __launch_bounds__(256) sgemm_tn_kernel(const float* __restrict__ A, const float* __restrict__ Bm,
                                                       float* __restrict__ C, const float* __restrict__ bias,
                                                       const float* __restrict__ res, long long M, int N, int K,
                                                       int ldc, int N1) {
  constexpr int BM = 128, BN = 128, BK = 8;
  __shared__ __align__(16) float As[2][BK][BM];
  __shared__ __align__(16) float Bs[2][BK][BN];
  const int tid = threadIdx.x;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int lrow = tid >> 1, lk = (tid & 1) * 4;      // each thread loads one float4 of A and of B per k-tile
  const int ty = tid >> 4, tx = tid & 15;             // 16 x 16 threads, 8 x 8 outputs each
  const long long arow = m0 + lrow;
  const bool a_ok = arow < M;
  const float* ap = A + (a_ok ? arow : 0) * (long long)K + lk;
  const float* bp = Bm + (long long)(n0 + lrow) * K + lk;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float4 ra = a_ok ? *reinterpret_cast<const float4*>(ap) : make_float4(0, 0, 0, 0);
  float4 rb = *reinterpret_cast<const float4*>(bp);
  As[0][lk + 0][lrow] = ra.x; As[0][lk + 1][lrow] = ra.y; As[0][lk + 2][lrow] = ra.z; As[0][lk + 3][lrow] = ra.w;
  Bs[0][lk + 0][lrow] = rb.x; Bs[0][lk + 1][lrow] = rb.y; Bs[0][lk + 2][lrow] = rb.z; Bs[0][lk + 3][lrow] = rb.w;
  __syncthreads();
  const int nk = K / BK;
  for (int kt = 0; kt < nk; ++kt) {
    const int cur = kt & 1;
    if (kt + 1 < nk) {
      ra = a_ok ? *reinterpret_cast<const float4*>(ap + (kt + 1) * BK) : make_float4(0, 0, 0, 0);
      rb = *reinterpret_cast<const float4*>(bp + (kt + 1) * BK);
    }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[cur][k][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[cur][k][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[cur][k][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[cur][k][64 + tx * 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      const int nx = cur ^ 1;
      As[nx][lk + 0][lrow] = ra.x; As[nx][lk + 1][lrow] = ra.y; As[nx][lk + 2][lrow] = ra.z; As[nx][lk + 3][lrow] = ra.w;
      Bs[nx][lk + 0][lrow] = rb.x; Bs[nx][lk + 1][lrow] = rb.y; Bs[nx][lk + 2][lrow] = rb.z; Bs[nx][lk + 3][lrow] = rb.w;
    }
    __syncthreads();
  }
  // epilogue: rows m0 + {ty*4+i, 64+ty*4+i}, cols n0 + {tx*4+j, 64+tx*4+j}
#pragma unroll
  for (int ih = 0; ih < 2; ++ih)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long long m = m0 + ih * 64 + ty * 4 + i;
      if (m >= M) continue;
#pragma unroll
      for (int jh = 0; jh < 2; ++jh) {
        const int n = n0 + jh * 64 + tx * 4;
        float4 v = make_float4(acc[ih * 4 + i][jh * 4 + 0], acc[ih * 4 + i][jh * 4 + 1], acc[ih * 4 + i][jh * 4 + 2],
                               acc[ih * 4 + i][jh * 4 + 3]);
        if (EPI == EPI_BIAS || EPI == EPI_BIAS_RELU || EPI == EPI_BIAS_RES) {
          float4 bb = *reinterpret_cast<const float4*>(bias + n);
          v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
        }
        if (EPI == EPI_BIAS_RELU) {
          v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
        }
        if (EPI == EPI_BIAS_RES) {
          float4 rr = *reinterpret_cast<const float4*>(res + m * ldc + n);
          v.x += rr.x; v.y += rr.y; v.z += rr.z; v.w += rr.w;
        }
        if (EPI == EPI_UMMA_SPLIT) {
          // fp16 hi/lo halves in the K-major core-matrix layout the rollout kernel feeds to tcgen05 (umma.cuh):
          // per aug-instance 3 operands [E' | K' | V^T] of N1p*512 bytes; this epilogue writes E' = [hi | lo], each N1p rows x 128 k; element (j, k) at (k/8)*N1p*16 + (j/8)*128 + (j%8)*16 + (k%8)*2
          const int N1p = (N1 + 15) & ~15;
          const long long bi = m / N1;
          const int j = (int)(m % N1);
          uint8_t* base = reinterpret_cast<uint8_t*>(C) + bi * (3LL * N1p * 512) + (size_t)(n >> 3) * N1p * 16 +
                          (j >> 3) * 128 + (j & 7) * 16 + (n & 7) * 2;
          const float vv[4] = {v.x, v.y, v.z, v.w};
          __half hi[4], lo[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            hi[q] = __float2half_rn(vv[q]);
            lo[q] = __float2half_rn(vv[q] - __half2float(hi[q]));
          }
          *reinterpret_cast<uint2*>(base) = make_uint2((uint32_t)__half_as_ushort(hi[0]) | ((uint32_t)__half_as_ushort(hi[1]) << 16),
                                                       (uint32_t)__half_as_ushort(hi[2]) | ((uint32_t)__half_as_ushort(hi[3]) << 16));
          *reinterpret_cast<uint2*>(base + (size_t)N1p * 256) =
              make_uint2((uint32_t)__half_as_ushort(lo[0]) | ((uint32_t)__half_as_ushort(lo[1]) << 16),
                         (uint32_t)__half_as_ushort(lo[2]) | ((uint32_t)__half_as_ushort(lo[3]) << 16));
        } else {
          int nn = n;
          if (EPI == EPI_SWIZZLE) nn = eswz((int)(m % N1), n);
          *reinterpret_cast<float4*>(C + m * ldc + nn) = v;
        }
      }
    }
}


// ---- encoder self-attention on tcgen05 (N1 <= 112): one CTA per (aug-instance, head) ----------------------------
// S = Q K^T (SS form, M = 128 query rows, N = N1p keys, K = 16) -> TMEM; thread = query row: softmax in registers, P
// (fp16 hi/lo, two keys per column) back into TMEM over S; O = P V (TS form, N = 16) -> TMEM columns [112, 128).
// Split precision with the cross terms issued first (umma.cuh); q is pre-scaled by log2(e)/sqrt(D) (exp2 domain).
__global__ void __launch_bounds__(128, 3) enc_attention_tc_kernel(const float* __restrict__ qkv, int N1,
                                                                  float* __restrict__ att) {
  __shared__ __align__(128) uint8_t sQ[2 * 4096];          // A operand hi | lo: 128 rows x 16 k
  __shared__ __align__(128) uint8_t sK[2 * 112 * 32];      // B operand hi | lo: N1p keys x 16 k
  __shared__ __align__(128) uint8_t sV[2 * 112 * 32];      // B operand hi | lo: 16 d x N1p keys
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tptr;
  const int N1p = (N1 + 15) & ~15;
  const int h = blockIdx.x % H;
  const long long b = blockIdx.x / H;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int r = tid;                                        // query row = key row = TMEM lane
  const float* rowp = qkv + (b * N1 + r) * (3LL * E) + h * D;
  const uint32_t khalf = (uint32_t)N1p * 32u;
  {
    uint32_t qh[8], ql[8], kh[8], kl[8];
    float vv[D];
    if (r < N1) {
#pragma unroll
      for (int d4 = 0; d4 < D / 4; ++d4) {
        const float4 q4 = *reinterpret_cast<const float4*>(rowp + d4 * 4);
        const float4 k4 = *reinterpret_cast<const float4*>(rowp + E + d4 * 4);
        const float4 v4 = *reinterpret_cast<const float4*>(rowp + 2 * E + d4 * 4);
        const float sc = 0.36067376022224085f;              // log2(e) / sqrt(16)
        umma::split2_f16(q4.x * sc, q4.y * sc, qh[d4 * 2], ql[d4 * 2]);
        umma::split2_f16(q4.z * sc, q4.w * sc, qh[d4 * 2 + 1], ql[d4 * 2 + 1]);
        umma::split2_f16(k4.x, k4.y, kh[d4 * 2], kl[d4 * 2]);
        umma::split2_f16(k4.z, k4.w, kh[d4 * 2 + 1], kl[d4 * 2 + 1]);
        vv[d4 * 4] = v4.x; vv[d4 * 4 + 1] = v4.y; vv[d4 * 4 + 2] = v4.z; vv[d4 * 4 + 3] = v4.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) qh[i] = ql[i] = kh[i] = kl[i] = 0u;
#pragma unroll
      for (int d = 0; d < D; ++d) vv[d] = 0.f;
    }
    const uint32_t ro = (uint32_t)(r >> 3) * 128u + (uint32_t)(r & 7) * 16u;
    *reinterpret_cast<uint4*>(sQ + ro) = make_uint4(qh[0], qh[1], qh[2], qh[3]);
    *reinterpret_cast<uint4*>(sQ + 2048 + ro) = make_uint4(qh[4], qh[5], qh[6], qh[7]);
    *reinterpret_cast<uint4*>(sQ + 4096 + ro) = make_uint4(ql[0], ql[1], ql[2], ql[3]);
    *reinterpret_cast<uint4*>(sQ + 4096 + 2048 + ro) = make_uint4(ql[4], ql[5], ql[6], ql[7]);
    if (r < N1p) {
      const uint32_t lbo = (uint32_t)N1p * 16u;
      *reinterpret_cast<uint4*>(sK + ro) = make_uint4(kh[0], kh[1], kh[2], kh[3]);
      *reinterpret_cast<uint4*>(sK + lbo + ro) = make_uint4(kh[4], kh[5], kh[6], kh[7]);
      *reinterpret_cast<uint4*>(sK + khalf + ro) = make_uint4(kl[0], kl[1], kl[2], kl[3]);
      *reinterpret_cast<uint4*>(sK + khalf + lbo + ro) = make_uint4(kl[4], kl[5], kl[6], kl[7]);
#pragma unroll
      for (int d = 0; d < D; ++d) {
        __half hi, lo;
        umma::split_f16(vv[d], hi, lo);
        const uint32_t off = umma::elem_off(d, r, 256u);
        *reinterpret_cast<__half*>(sV + off) = hi;
        *reinterpret_cast<__half*>(sV + khalf + off) = lo;
      }
    }
  }
  umma::fence_async_smem();
  if (warp == 0) umma::tmem_alloc(&tptr, 128);
  if (tid == 0) umma::mbar_init(&bar, 1);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tm = tptr;
  const uint32_t tl = tm + ((uint32_t)(warp * 32) << 16);
  if (tid == 0) {
    const uint32_t idesc = umma::make_idesc_f16(128, N1p);
    const uint32_t lbo = (uint32_t)N1p * 16u;
    const uint64_t aHi = umma::make_desc(umma::smem_addr(sQ), 2048, 128), aLo = umma::make_desc(umma::smem_addr(sQ) + 4096, 2048, 128);
    const uint64_t bHi = umma::make_desc(umma::smem_addr(sK), lbo, 128), bLo = umma::make_desc(umma::smem_addr(sK) + khalf, lbo, 128);
    umma::mma_f16_ss(tm, aLo, bHi, idesc, false);
    umma::mma_f16_ss(tm, aHi, bLo, idesc, true);
    umma::mma_f16_ss(tm, aHi, bHi, idesc, true);
    umma::commit(&bar);
  }
  umma::mbar_wait(&bar, 0);
  umma::fence_after_sync();
  uint32_t sr[112];
  umma::ld32_nw(tl, sr);
  if (N1p > 32) umma::ld32_nw(tl + 32, sr + 32);
  if (N1p > 64) umma::ld32_nw(tl + 64, sr + 64);
  if (N1p > 96) umma::ld16_nw(tl + 96, sr + 96);
  umma::wait_ld();
  float mx = -INFINITY;
#pragma unroll
  for (int c = 0; c < 7; ++c)
    if (c * 16 < N1p) {
#pragma unroll
      for (int i = c * 16; i < c * 16 + 16; ++i) {
        const float sv = i < N1 ? umma::after_wait(sr[i]) : -INFINITY;
        sr[i] = __float_as_uint(sv);
        mx = fmaxf(mx, sv);
      }
    }
  const float moff = mx - 10.f;                             // weights 2^(s - max + 10): small ones stay out of fp16 subnormals
  float l = 0.f;
#pragma unroll
  for (int c = 0; c < 7; ++c)
    if (c * 16 < N1p) {
#pragma unroll
      for (int i = c * 8; i < c * 8 + 8; ++i) {
        const float p0 = umma::ex2_raw(__uint_as_float(sr[2 * i]) - moff), p1 = umma::ex2_raw(__uint_as_float(sr[2 * i + 1]) - moff);
        l += p0 + p1;
        uint32_t hw, lw;
        umma::split2_f16(p0, p1, hw, lw);
        sr[2 * i] = hw;
        sr[2 * i + 1] = lw;
      }
      umma::st8s<2>(tl + c * 8, sr + c * 16);                              // P hi: keys 16c .. 16c+15
      umma::st8s<2>(tl + (N1p >> 1) + c * 8, sr + c * 16 + 1);             // P lo
    }
  umma::wait_st();
  umma::fence_before_sync();
  __syncthreads();
  if (tid == 0) {
    umma::fence_after_sync();
    const uint32_t idesc = umma::make_idesc_f16(128, 16);
    const uint32_t vHi = umma::smem_addr(sV), vLo = vHi + khalf, pHi = tm, pLo = tm + (N1p >> 1), dO = tm + 112;
    const int nks = N1p >> 4;
    for (int ks = 0; ks < nks; ++ks) umma::mma_f16_ts(dO, pLo + 8 * ks, umma::make_desc(vHi + ks * 512, 256, 128), idesc, ks > 0);
    for (int ks = 0; ks < nks; ++ks) umma::mma_f16_ts(dO, pHi + 8 * ks, umma::make_desc(vLo + ks * 512, 256, 128), idesc, true);
    for (int ks = 0; ks < nks; ++ks) umma::mma_f16_ts(dO, pHi + 8 * ks, umma::make_desc(vHi + ks * 512, 256, 128), idesc, true);
    umma::commit(&bar);
  }
  umma::mbar_wait(&bar, 1);
  umma::fence_after_sync();
  {
    uint32_t orr[16];
    umma::ld16_nw(tl + 112, orr);
    umma::wait_ld();
    if (r < N1) {
      const float inv = 1.f / l;
      float* op = att + (b * N1 + r) * E + h * D;
#pragma unroll
      for (int d4 = 0; d4 < D / 4; ++d4)
        *reinterpret_cast<float4*>(op + d4 * 4) = make_float4(umma::after_wait(orr[d4 * 4]) * inv, umma::after_wait(orr[d4 * 4 + 1]) * inv,
                                                              umma::after_wait(orr[d4 * 4 + 2]) * inv, umma::after_wait(orr[d4 * 4 + 3]) * inv);
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tm, 128);
}

// ---- decoder keys / values as fp16 hi/lo tcgen05 B operands (rollout_tc.cu), segments 1 and 2 of elg_tables.e ----
//   K'  [N1p keys x 128 k]  element (j, c) at (c/8)*N1p*16 + (j/8)*128 + (j%8)*16 + (c%8)*2        (like E')
//   V^T per head h: [16 d x N1p keys]  element (d, j) at h*N1p*32 + (j/8)*256 + (d/8)*128 + (d%8)*16 + (j%8)*2
// each as [hi | lo] halves of N1p*256 bytes.  One CTA per aug-instance; padded key slots are written as zeros.
__global__ void __launch_bounds__(256) split_kv_kernel(const float* __restrict__ K, const float* __restrict__ V,
                                                       uint8_t* __restrict__ ops, int N1) {
  const int N1p = (N1 + 15) & ~15;
  const size_t b = blockIdx.x;
  const float* kp = K + b * N1 * E;
  const float* vp = V + b * N1 * E;
  uint8_t* eo = ops + b * 3 * ((size_t)N1p * 512);
  uint8_t* ko = eo + (size_t)N1p * 512;
  uint8_t* vo = ko + (size_t)N1p * 512;
  const uint32_t half = (uint32_t)N1p * 256u;
  // K': unit = (key j, 8-column chunk): 32 bytes in, one 16-byte core-matrix row of hi and of lo out; j % 8 fastest, so eight
  // lanes fill 128 contiguous bytes.  Key slots N1 .. N1p-1 are written as zeros here -- for K' and for E' (whose live rows
  // come from the GEMM epilogue) -- so the caller needs no memset of the operand buffer.
  for (int i = threadIdx.x; i < N1p * (E / 8); i += 256) {
    const int j = (i & 7) + ((i >> 3) / (E / 8)) * 8, c8 = (i >> 3) % (E / 8);
    const uint32_t off = (uint32_t)c8 * (uint32_t)N1p * 16u + (uint32_t)(j >> 3) * 128u + (uint32_t)(j & 7) * 16u;
    uint4 h = make_uint4(0u, 0u, 0u, 0u), l = h;
    if (j < N1) {
      const float4 a = *reinterpret_cast<const float4*>(kp + (size_t)j * E + c8 * 8);
      const float4 c = *reinterpret_cast<const float4*>(kp + (size_t)j * E + c8 * 8 + 4);
      umma::split2_f16(a.x, a.y, h.x, l.x);
      umma::split2_f16(a.z, a.w, h.y, l.y);
      umma::split2_f16(c.x, c.y, h.z, l.z);
      umma::split2_f16(c.z, c.w, h.w, l.w);
    } else {
      *reinterpret_cast<uint4*>(eo + off) = h;
      *reinterpret_cast<uint4*>(eo + half + off) = h;
    }
    *reinterpret_cast<uint4*>(ko + off) = h;
    *reinterpret_cast<uint4*>(ko + half + off) = l;
  }
  const int jp = N1p >> 1;
  for (int i = threadIdx.x; i < jp * E; i += 256) {               // V^T: key pairs of one (head, d) -> one 32-bit word
    const int c = i % E, j = (i / E) * 2;
    const float a = j < N1 ? vp[(size_t)j * E + c] : 0.f;
    const float bq = j + 1 < N1 ? vp[(size_t)(j + 1) * E + c] : 0.f;
    __half h0, l0, h1, l1;
    umma::split_f16(a, h0, l0);
    umma::split_f16(bq, h1, l1);
    const uint32_t off = (uint32_t)(c >> 4) * (uint32_t)N1p * 32u + umma::elem_off(c & 15, j, 256u);
    *reinterpret_cast<uint32_t*>(vo + off) = umma::pack_h2(h0, h1);
    *reinterpret_cast<uint32_t*>(vo + half + off) = umma::pack_h2(l0, l1);
  }
}

// ---- the same operands in tiles of ELG_TILE_NODES nodes for the streamed tensor-core rollout (rollout_stc.cu) ----------
// One CTA per (aug-instance, tile): [E' | K' | V^T], each 32 KB hi + 32 KB lo, in the layouts above with N1p = 128.
// E' is read back from the XOR-swizzled fp32 copy the fp32-pipe kernel uses (elg_tables.e); keys past N1 are zero.
__global__ void __launch_bounds__(256) split_tiles_kernel(const float* __restrict__ K, const float* __restrict__ V,
                                                          const float* __restrict__ Esw, uint8_t* __restrict__ et, int N1) {
  const int tiles = (N1 + ELG_TILE_NODES - 1) / ELG_TILE_NODES;
  const size_t b = blockIdx.x / tiles;
  const int tile = blockIdx.x % tiles, j0 = tile * ELG_TILE_NODES;
  uint8_t* out = et + (size_t)blockIdx.x * ELG_TILE_BYTES;
  for (int i = threadIdx.x; i < ELG_TILE_NODES * (E / 2); i += 256) {      // E', K': column pairs of one node -> one 32-bit word
    const int jl = i / (E / 2), c = (i % (E / 2)) * 2, j = j0 + jl;
    float2 ev = make_float2(0.f, 0.f), kv = make_float2(0.f, 0.f);
    if (j < N1) {
      const float* er = Esw + (b * N1 + j) * E;
      ev = make_float2(er[eswz(j, c)], er[eswz(j, c + 1)]);
      kv = *reinterpret_cast<const float2*>(K + (b * N1 + j) * E + c);
    }
    const uint32_t off = umma::elem_off(jl, c, ELG_TILE_NODES * 16u);
    __half h0, l0, h1, l1;
    umma::split_f16(ev.x, h0, l0); umma::split_f16(ev.y, h1, l1);
    *reinterpret_cast<uint32_t*>(out + off) = umma::pack_h2(h0, h1);
    *reinterpret_cast<uint32_t*>(out + 32768 + off) = umma::pack_h2(l0, l1);
    umma::split_f16(kv.x, h0, l0); umma::split_f16(kv.y, h1, l1);
    *reinterpret_cast<uint32_t*>(out + 65536 + off) = umma::pack_h2(h0, h1);
    *reinterpret_cast<uint32_t*>(out + 65536 + 32768 + off) = umma::pack_h2(l0, l1);
  }
  for (int i = threadIdx.x; i < (ELG_TILE_NODES / 2) * E; i += 256) {      // V^T: node pairs of one (head, d) -> one 32-bit word
    const int c = i % E, jl = (i / E) * 2, j = j0 + jl;
    const float a = j < N1 ? V[(b * N1 + j) * E + c] : 0.f;
    const float bq = j + 1 < N1 ? V[(b * N1 + j + 1) * E + c] : 0.f;
    __half h0, l0, h1, l1;
    umma::split_f16(a, h0, l0); umma::split_f16(bq, h1, l1);
    const uint32_t off = (uint32_t)(c >> 4) * (ELG_TILE_NODES * 32u) + umma::elem_off(c & 15, jl, 256u);
    *reinterpret_cast<uint32_t*>(out + 131072 + off) = umma::pack_h2(h0, h1);
    *reinterpret_cast<uint32_t*>(out + 131072 + 32768 + off) = umma::pack_h2(l0, l1);
  }
}

// ---- tcgen05 GEMM  C[M][N] = A[M][K] * W[N][K]^T (+ bias / relu / residual) ---------------------------------
// Split precision: x = hi + lo (fp16 pair, ~22 mantissa bits); D = A_hi W_hi + A_hi W_lo + A_lo W_hi accumulated in
// fp32 in TMEM (two accumulators: hi*hi | cross terms; fp32-matmul-grade accuracy, tools/umma_precision_experiment.py)
// -- plain TF32/BF16 inputs would break the 1e-4 logit tolerance.  One CTA = 128 rows; the fp32 activations are split on the fly into the K-major
// core-matrix operand layout (umma.cuh), the weights were pre-split at model load (split_weights_kernel) and arrive
// by TMA bulk copy from 32 KB (n-tile, 64-wide k-block) tiles.
// Persistent and warp-specialised (round 2): one CTA per SM walks m-tiles blockIdx.x, blockIdx.x + gridDim.x, ...
//   warps 0-7   epilogue: drain an accumulator set (TMEM lane = row; warp = (lane quadrant, 64-column half)) 16 columns at a
//               time through a swizzled 2 KB staging tile per warp, so every global store / residual load instruction covers
//               64 contiguous bytes of 8 rows (full sectors); residual rows are prefetched two chunks ahead
//   warps 8-15  builders: fp32 activation block (128 x 128) -> fp16 hi/lo operand tiles, double-buffered, so the block of
//               the next m-tile (or k super-block) is converted while the tensor core works on the current one;
//               lane = (row % 8, 8-wide k-chunk): 32-byte global sectors in, conflict-free 16-byte core-matrix rows out
//   warp 16     MMA issuer: split-precision MMAs of a 32-wide k-block as soon as its B stage and the A block have landed;
//               tcgen05.commit frees the stage / the A block; TWO accumulator sets (2 x 256 TMEM columns), so the
//               epilogue of one n-tile overlaps the MMAs of the next
//   warp 17     TMA producer: 16 KB weight stages (n-tile, 32-wide k-block: hi 8 KB + lo 8 KB) through a 5-stage ring
// K = 128 with any number of n-tiles, or one n-tile with K a multiple of 128 (the encoder's shapes).
constexpr int TG_A_BYTES = 65536, TG_B_BYTES = 16384, TG_STAGES = 5, TG_STAGE_TILE = 2048;
constexpr int TG_OFF_B = 2 * TG_A_BYTES, TG_OFF_STG = TG_OFF_B + TG_STAGES * TG_B_BYTES, TG_OFF_BAR = TG_OFF_STG + 8 * TG_STAGE_TILE;
constexpr int TG_SMEM = TG_OFF_BAR + 256;
constexpr int TG_THREADS = 576;      // 8 epilogue + 8 builder warps + MMA issuer + TMA producer
static_assert(TG_SMEM <= 232448, "tc_gemm_kernel shared memory");

__device__ __forceinline__ void tg_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(umma::smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tg_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(umma::smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void tg_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(umma::smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void tg_prefetch_l2(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tg_epi_sync() { asm volatile("bar.sync 2, 256;" ::: "memory"); }
__device__ __forceinline__ void tg_builders_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// cplane: element offset between the outputs of consecutive n-tiles (0 = n-tile nt writes columns [128 nt, 128 nt + 128)
// of C; otherwise n-tile nt writes columns [0, 128) of the plane C + nt * cplane -- several N = 128 products of one input)
template <int EPI>
__global__ void __launch_bounds__(TG_THREADS, 1) tc_gemm_kernel(const float* __restrict__ A, const uint8_t* __restrict__ Ws,
                                                                float* __restrict__ C, const float* __restrict__ bias,
                                                                const float* __restrict__ res, long long M, int N, int K, int ldc,
                                                                long long cplane) {
  extern __shared__ __align__(1024) uint8_t tsm[];
  uint8_t* aBuf = tsm;                                   // [2][hi 32 KB | lo 32 KB]: 128 rows x 128 k
  uint64_t* bars = reinterpret_cast<uint64_t*>(tsm + TG_OFF_BAR);
  uint64_t* b_full = bars;                      // [5] TMA landed
  uint64_t* b_empty = bars + TG_STAGES;         // [5] MMAs of the stage done
  uint64_t* a_full = bars + 2 * TG_STAGES;      // [2] builders wrote the A block
  uint64_t* a_empty = a_full + 2;               // [2] MMAs of the A block done
  uint64_t* acc_full = a_full + 4;              // [2] accumulator set complete
  uint64_t* acc_empty = a_full + 6;             // [2] accumulator set drained
  uint32_t* tptr = reinterpret_cast<uint32_t*>(a_full + 8);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) umma::tmem_alloc(tptr, 512);
  if (tid == 0) {
    for (int i = 0; i < 2 * TG_STAGES + 8; ++i) umma::mbar_init(bars + i, 1);
  }
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tbase = *tptr;
  const int nT = N >> 7, kS = K >> 7;
  const long long n_tiles = (M + 127) >> 7;

  if (warp == 17) {
    // ---------------- TMA producer (+ L2 prefetch of the activation / residual rows of the tiles ahead) ----------------
    const uint32_t b0 = umma::smem_addr(tsm + TG_OFF_B);
    const int ahead = 0;          // L2 bulk prefetch of the tiles ahead: measured slower (FFN2 707 -> 899 us), kept switched off
    auto prefetch_tile = [&](long long tile) {
      if (tile >= n_tiles) return;
      const long long row0 = tile * 128;
      const long long nrows = M - row0 < 128 ? M - row0 : 128;
      if (kS == 1) {            // rows of 512 bytes: the whole block is contiguous
        if (lane == 0) tg_prefetch_l2(A + row0 * K, (uint32_t)(nrows * 512));
      } else {
        for (long long r = lane; r < nrows; r += 32) tg_prefetch_l2(A + (row0 + r) * K, (uint32_t)(K * 4));
      }
      if (EPI == EPI_BIAS_RES && lane == 1) {
        if (ldc == 128) tg_prefetch_l2(res + row0 * 128, (uint32_t)(nrows * 512));
      }
    };
    uint32_t bi = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      if (ahead) prefetch_tile(tile + (long long)ahead * gridDim.x);
      __syncwarp();
      if (lane == 0) {
        for (int nt = 0; nt < nT; ++nt)
          for (int kb = 0; kb < (K >> 5); ++kb, ++bi) {
            const uint32_t st = bi % TG_STAGES;
            if (bi >= TG_STAGES) umma::mbar_wait(b_empty + st, ((bi / TG_STAGES) - 1) & 1);
            const uint8_t* src = Ws + ((size_t)nt * (K >> 6) + (kb >> 1)) * 32768 + (kb & 1) * 8192;
            tg_mbar_expect_tx(b_full + st, TG_B_BYTES);
            tg_bulk_g2s(b0 + st * TG_B_BYTES, src, 8192, b_full + st);                   // hi: 128 n x 32 k
            tg_bulk_g2s(b0 + st * TG_B_BYTES + 8192, src + 16384, 8192, b_full + st);    // lo
          }
      }
      __syncwarp();
    }
  } else if (warp == 16) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      const uint32_t idesc = umma::make_idesc_f16(128, 128);
      const uint32_t a0 = umma::smem_addr(aBuf), b0 = umma::smem_addr(tsm + TG_OFF_B);
      uint32_t bi = 0, ai = 0, ci = 0;
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ai += kS)
        for (int nt = 0; nt < nT; ++nt, ++ci) {
          const uint32_t acc = ci & 1;
          if (ci >= 2) umma::mbar_wait(acc_empty + acc, ((ci >> 1) - 1) & 1);
          umma::fence_after_sync();
          const uint32_t d0 = tbase + acc * 256;
          for (int ks = 0; ks < kS; ++ks) {
            const uint32_t aidx = ai + ks, slot = aidx & 1;
            if (nt == 0) umma::mbar_wait(a_full + slot, (aidx >> 1) & 1);
            const uint32_t aHi = a0 + slot * TG_A_BYTES, aLo = aHi + 32768;
            for (int kb = 0; kb < 4; ++kb, ++bi) {
              const uint32_t st = bi % TG_STAGES;
              umma::mbar_wait(b_full + st, (bi / TG_STAGES) & 1);
              umma::fence_after_sync();
              const uint32_t bHi = b0 + st * TG_B_BYTES, bLo = bHi + 8192;
#pragma unroll
              for (int term = 0; term < 3; ++term) {
                const uint32_t a = (term == 2 ? aLo : aHi) + kb * 4 * 2048;
                const uint32_t b2 = term == 1 ? bLo : bHi;
#pragma unroll
                for (int kk = 0; kk < 2; ++kk)
                  // the small cross terms get their own accumulator: added into the large hi*hi sums, the tensor core's
                  // internal alignment would truncate them (measured 4x error, tools/umma_precision_experiment.py)
                  umma::mma_f16_ss(d0 + (term ? 128u : 0u), umma::make_desc(a + kk * 4096, 2048, 128),
                                   umma::make_desc(b2 + kk * 4096, 2048, 128), idesc,
                                   !(ks == 0 && kb == 0 && (term == 0 || term == 1) && kk == 0));
              }
              umma::commit(b_empty + st);
            }
            if (nt == nT - 1) umma::commit(a_empty + slot);
          }
          umma::commit(acc_full + acc);
        }
    }
  } else if (warp >= 8) {
    // ---------------- builders: A operand blocks ----------------
    const int bw = warp - 8, r8 = lane & 7, cj = (bw & 3) * 4 + (lane >> 3), rh = bw >> 2;   // k in [8 cj, 8 cj + 8), rows [64 rh, 64 rh + 64)
    // rolling prefetch: the registers of a converted row group are refilled at once with the same group of the NEXT block
    // (next k super-block or next m-tile), so loads stay in flight through the store / barrier / a_empty wait
    float4 v[16];
    auto fetch = [&](long long tile, int ks, int it) {
      const long long row = tile * 128 + (rh * 8 + it) * 8 + r8;
      const float* src = A + row * K + ks * 128 + cj * 8;
      if (tile < n_tiles && row < M) {
        v[2 * it] = *reinterpret_cast<const float4*>(src);
        v[2 * it + 1] = *reinterpret_cast<const float4*>(src + 4);
      } else {
        v[2 * it] = v[2 * it + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
#pragma unroll
    for (int it = 0; it < 8; ++it) fetch(blockIdx.x, 0, it);
    uint32_t aidx = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      for (int ks = 0; ks < kS; ++ks, ++aidx) {
        const uint32_t slot = aidx & 1;
        const bool last_ks = ks == kS - 1;
        const long long ntile = last_ks ? tile + gridDim.x : tile;
        const int nks = last_ks ? 0 : ks + 1;
        if (aidx >= 2) umma::mbar_wait(a_empty + slot, ((aidx >> 1) - 1) & 1);
        uint8_t* aHi = aBuf + slot * TG_A_BYTES;
        uint8_t* aLo = aHi + 32768;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          uint4 h, l;
          umma::split2_f16(v[2 * it].x, v[2 * it].y, h.x, l.x);
          umma::split2_f16(v[2 * it].z, v[2 * it].w, h.y, l.y);
          umma::split2_f16(v[2 * it + 1].x, v[2 * it + 1].y, h.z, l.z);
          umma::split2_f16(v[2 * it + 1].z, v[2 * it + 1].w, h.w, l.w);
          fetch(ntile, nks, it);
          const uint32_t off = (uint32_t)cj * 2048u + (uint32_t)(rh * 8 + it) * 128u + (uint32_t)r8 * 16u;
          *reinterpret_cast<uint4*>(aHi + off) = h;
          *reinterpret_cast<uint4*>(aLo + off) = l;
        }
        umma::fence_async_smem();
        tg_builders_sync();
        if (tid == 256) tg_mbar_arrive(a_full + slot);
      }
    }
  } else {
    // ---------------- epilogue: TMEM lane = row; warp % 4 = lane quadrant (32 rows), warp / 4 = 64-column half ----------------
    const int q = warp & 3, ch = warp >> 2;
    float* stg = reinterpret_cast<float*>(tsm + TG_OFF_STG + warp * TG_STAGE_TILE);     // [32 rows][4 float4], float4 index ^ ((row >> 1) & 3)
    const int er = lane >> 2, ec = lane & 3;                                             // store side: rows 8 k + er, float4 ec
    // residual rows are requested TWO 16-column chunks ahead (across n-tile and m-tile boundaries), so their latency
    // hides behind the accumulator wait and the stores of the chunks in between
    float4 rr[2][4];
    auto fetch_res = [&](float4* dst, long long tile, int nt, int c) {
      const int col = (cplane ? 0 : nt * 128) + ch * 64 + c * 16 + 4 * ec;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const long long row = tile * 128 + 32 * q + 8 * k + er;
        dst[k] = (tile < n_tiles && row < M) ? *reinterpret_cast<const float4*>(res + (size_t)row * ldc + col) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    if (EPI == EPI_BIAS_RES) {
      fetch_res(rr[0], blockIdx.x, 0, 0);
      fetch_res(rr[1], blockIdx.x, 0, 1);
    }
    uint32_t ci = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const long long m0 = tile * 128 + 32 * q;
      for (int nt = 0; nt < nT; ++nt, ++ci) {
        const uint32_t acc = ci & 1;
        umma::mbar_wait(acc_full + acc, (ci >> 1) & 1);
        umma::fence_after_sync();
        float* Cp = cplane ? C + nt * cplane : C;
        const int nbase = (cplane ? 0 : nt * 128) + ch * 64;
        const uint32_t t0 = tbase + ((uint32_t)(32 * q) << 16) + acc * 256 + ch * 64;
        const bool last_nt = nt == nT - 1;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
          if (EPI == EPI_BIAS || EPI == EPI_BIAS_RELU || EPI == EPI_BIAS_RES)
            bb = *reinterpret_cast<const float4*>(bias + nt * 128 + ch * 64 + c * 16 + 4 * ec);
          uint32_t v[16], v2[16];
          umma::ld16_nw(t0 + c * 16, v);
          umma::ld16_nw(t0 + 128 + c * 16, v2);
          umma::wait_ld();
#pragma unroll
          for (int i4 = 0; i4 < 4; ++i4) {
            float4 o;
            o.x = umma::after_wait(v[4 * i4]) + umma::after_wait(v2[4 * i4]);
            o.y = umma::after_wait(v[4 * i4 + 1]) + umma::after_wait(v2[4 * i4 + 1]);
            o.z = umma::after_wait(v[4 * i4 + 2]) + umma::after_wait(v2[4 * i4 + 2]);
            o.w = umma::after_wait(v[4 * i4 + 3]) + umma::after_wait(v2[4 * i4 + 3]);
            *reinterpret_cast<float4*>(stg + lane * 16 + 4 * (i4 ^ ((lane >> 1) & 3))) = o;
          }
          __syncwarp();
          float4 o[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int r = 8 * k + er;
            o[k] = *reinterpret_cast<const float4*>(stg + r * 16 + 4 * (ec ^ ((r >> 1) & 3)));
            o[k].x += bb.x; o[k].y += bb.y; o[k].z += bb.z; o[k].w += bb.w;
            if (EPI == EPI_BIAS_RELU) { o[k].x = fmaxf(o[k].x, 0.f); o[k].y = fmaxf(o[k].y, 0.f); o[k].z = fmaxf(o[k].z, 0.f); o[k].w = fmaxf(o[k].w, 0.f); }
            if (EPI == EPI_BIAS_RES) { o[k].x += rr[c & 1][k].x; o[k].y += rr[c & 1][k].y; o[k].z += rr[c & 1][k].z; o[k].w += rr[c & 1][k].w; }
          }
          if (EPI == EPI_BIAS_RES) {       // chunk c + 2 in (tile, nt, c) order
            if (c < 2) fetch_res(rr[c & 1], tile, nt, c + 2);
            else fetch_res(rr[c & 1], last_nt ? tile + gridDim.x : tile, last_nt ? 0 : nt + 1, c - 2);
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const long long row = m0 + 8 * k + er;
            if (row < M) *reinterpret_cast<float4*>(Cp + (size_t)row * ldc + nbase + c * 16 + 4 * ec) = o[k];
          }
          __syncwarp();
        }
        umma::fence_before_sync();
        tg_epi_sync();
        if (tid == 0) tg_mbar_arrive(acc_empty + acc);
      }
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tbase, 512);
}

template <int EPI>
static int tc_gemm(const float* A, const float* derived, long long split_off, float* C, const float* bias, const float* res,
                   long long M, int N, int K, int ldc, cudaStream_t st, long long cplane = 0) {
  ELG_REQUIRE(N % 128 == 0 && K % 128 == 0, ELG_EUNSUPPORTED, "tc_gemm needs N%%128==0 and K%%128==0 (N=%d K=%d)", N, K);
  ELG_REQUIRE(N == 128 || K == 128, ELG_EUNSUPPORTED, "tc_gemm: K = 128 or a single n-tile (N=%d K=%d)", N, K);
  static bool attr = false;
  static int sms = 0;
  if (!attr) {
    ELG_CUDA_OK(cudaFuncSetAttribute(tc_gemm_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, TG_SMEM));
    int dev = 0;
    ELG_CUDA_OK(cudaGetDevice(&dev));
    ELG_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    attr = true;
  }
  const long long tiles = (M + 127) / 128;
  tc_gemm_kernel<EPI><<<(unsigned)(tiles < sms ? tiles : sms), TG_THREADS, TG_SMEM, st>>>(
      A, reinterpret_cast<const uint8_t*>(derived + split_off), C, bias, res, M, N, K, ldc, cplane);
  ELG_LAUNCH_OK();
  return ELG_OK;
}

// ---- fused feed-forward block on tcgen05:  out = relu(x1 W1^T + b1) W2^T + b2 + x1  (CVRP/models.py:249-269) ----------
// The hidden activations (rows x ff) never leave the SM: per 128-row tile and per 128-wide hidden chunk j
//   M1(j)  acc1[j & 1] = x1 W1_j^T            SS MMAs, A = x1 operand (shared memory), B = W1 n-tile j (TMA ring)
//   E(j)   epilogue warps: accumulator row -> + b1, ReLU -> fp16 hi/lo -> written IN PLACE over the accumulator columns
//          as the tensor-memory A operand of the second contraction (lane = row; the rollout kernels' P operand again)
//   M2(j)  acc2 += hid_j W2[:, chunk j]^T     TS MMAs, A in tensor memory, B = W2 k-blocks of chunk j (TMA ring)
// issued as M1(0) M1(1) M2(0) M1(2) M2(1) M1(3) M2(2) M2(3): the tensor pipe works on the next chunk while the epilogue
// warps convert the current one.  M1 keeps ONE accumulator and issues its cross terms first (lo stages before hi stages;
// same accuracy as two accumulators, tools/umma_precision_experiment.py); acc2 collects four chunks and keeps the
// cross terms in their own accumulator.  TMEM: [0,128) [128,256) acc1 / hid sets, [256,384) acc2, [384,512) acc2 cross.
// Weight stages are the 16 KB hi or lo halves of the pre-split 32 KB (n-tile, 64-wide k-block) tiles.
// Warps 0-3 epilogue, 4-7 builders (x1 block of the next tile), 8 MMA issuer, 9 TMA producer; persistent over m-tiles.
constexpr int TF_STAGES = 9, TF_THREADS = 320;
constexpr int TF_OFF_B = TG_A_BYTES, TF_OFF_STG = TF_OFF_B + TF_STAGES * 16384, TF_OFF_B1 = TF_OFF_STG + 4 * 4096;
constexpr int TF_MAX_FF = 512;
constexpr int TF_OFF_BAR = TF_OFF_B1 + TF_MAX_FF * 4, TF_SMEM = TF_OFF_BAR + 256;
static_assert(TF_SMEM <= 232448, "tc_ffn_kernel shared memory");
__device__ __forceinline__ void tf_group_sync(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }

__global__ void __launch_bounds__(TF_THREADS, 1) tc_ffn_kernel(const float* __restrict__ X, const uint8_t* __restrict__ W1s,
                                                               const float* __restrict__ b1, const uint8_t* __restrict__ W2s,
                                                               const float* __restrict__ b2, float* __restrict__ out,
                                                               float* __restrict__ hid_out, long long M, int FF) {
  extern __shared__ __align__(1024) uint8_t tsm[];
  uint8_t* aBuf = tsm;                                   // [2][hi 32 KB | lo 32 KB]: 128 rows x 128 k of x1
  float* sB1 = reinterpret_cast<float*>(tsm + TF_OFF_B1);
  uint64_t* bars = reinterpret_cast<uint64_t*>(tsm + TF_OFF_BAR);
  uint64_t* b_full = bars;                      // [TF_STAGES]
  uint64_t* b_empty = bars + TF_STAGES;         // [TF_STAGES]
  uint64_t* a_full = bars + 2 * TF_STAGES;      // [2] builders wrote the x1 block
  uint64_t* a_empty = a_full + 2;               // [2] first-contraction MMAs of the tile done
  uint64_t* acc1_full = a_full + 4;             // [2] M1 of the chunk complete
  uint64_t* hid_ready = a_full + 6;             // [2] hidden chunk written as the TMEM A operand
  uint64_t* acc2_full = a_full + 8;             // M2 of the tile's last chunk complete
  uint64_t* acc2_empty = a_full + 9;            // final epilogue has read acc2
  uint32_t* tptr = reinterpret_cast<uint32_t*>(a_full + 10);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) umma::tmem_alloc(tptr, 512);
  if (tid == 0) {
    for (int i = 0; i < 2 * TF_STAGES + 10; ++i) umma::mbar_init(bars + i, 1);
  }
  for (int i = tid; i < FF; i += TF_THREADS) sB1[i] = b1[i];
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tbase = *tptr;
  const int nC = FF >> 7;                       // hidden chunks (even)
  const long long n_tiles = (M + 127) >> 7;

  if (warp == 9) {
    // ---------------- TMA producer: stages in the order the MMA issuer consumes them ----------------
    if (lane == 0) {
      const uint32_t b0 = umma::smem_addr(tsm + TF_OFF_B);
      uint32_t bi = 0;
      auto push = [&](const uint8_t* src) {
        const uint32_t st = bi % TF_STAGES;
        if (bi >= TF_STAGES) umma::mbar_wait(b_empty + st, ((bi / TF_STAGES) - 1) & 1);
        tg_mbar_expect_tx(b_full + st, 16384);
        tg_bulk_g2s(b0 + st * 16384, src, 16384, b_full + st);
        ++bi;
      };
      auto push_chunk = [&](const uint8_t* t0) {       // two 32 KB tiles [hi | lo]: lo halves first, then hi halves
        push(t0 + 16384); push(t0 + 32768 + 16384); push(t0); push(t0 + 32768);
      };
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        push_chunk(W1s);
        for (int j = 0; j < nC; ++j) {
          if (j + 1 < nC) push_chunk(W1s + (size_t)(j + 1) * 65536);      // W1 n-tile j + 1: K = 128 -> two tiles
          push_chunk(W2s + (size_t)j * 65536);                            // W2 (one n-tile): k-blocks 2 j, 2 j + 1
        }
      }
    }
  } else if (warp == 8) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      const uint32_t idesc = umma::make_idesc_f16(128, 128);
      const uint32_t a0 = umma::smem_addr(aBuf), b0 = umma::smem_addr(tsm + TF_OFF_B);
      uint32_t bi = 0, g = 0, ti = 0;
      // one 128-wide contraction from four weight stages (lo kb0, lo kb1, hi kb0, hi kb1); SS: A from shared memory,
      // otherwise A = hidden chunk in tensor memory (hi words of k < 64 at +0, lo at +32, k >= 64 at +64 / +96)
      auto contract = [&](bool ss, uint32_t aHi, uint32_t aLo, uint32_t d_main, uint32_t d_cross, bool first) {
        bool acc_cross = !first, acc_main = !first || d_main == d_cross;
        // one pass over a 64-wide k-block: 4 MMAs, A part (hi / lo) x the stage's B half
        auto pass = [&](uint32_t bS, int kb, bool use_alo, uint32_t d, bool& accf) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const int ks = kb * 4 + kk;                 // k16-step inside the 128-wide contraction
            const uint64_t bd = umma::make_desc(bS + kk * 4096, 2048, 128);
            if (ss) {
              umma::mma_f16_ss(d, umma::make_desc((use_alo ? aLo : aHi) + ks * 4096, 2048, 128), bd, idesc, accf);
            } else {
              const uint32_t col = (ks < 4 ? 8u * ks : 64u + 8u * (ks - 4)) + (use_alo ? 32u : 0u);
              umma::mma_f16_ts(d, aHi + col, bd, idesc, accf);
            }
            accf = true;
          }
        };
        uint32_t stS[4];
#pragma unroll
        for (int sidx = 0; sidx < 4; ++sidx) stS[sidx] = (bi + sidx) % TF_STAGES;
        // lo stages: A_hi B_lo
#pragma unroll
        for (int sidx = 0; sidx < 2; ++sidx) {
          umma::mbar_wait(b_full + stS[sidx], ((bi + sidx) / TF_STAGES) & 1);
          umma::fence_after_sync();
          pass(b0 + stS[sidx] * 16384, sidx, false, d_cross, acc_cross);
          umma::commit(b_empty + stS[sidx]);
        }
        // hi stages: A_lo B_hi over the whole contraction, then A_hi B_hi (every cross term before the large products)
#pragma unroll
        for (int sidx = 2; sidx < 4; ++sidx) {
          umma::mbar_wait(b_full + stS[sidx], ((bi + sidx) / TF_STAGES) & 1);
          umma::fence_after_sync();
          pass(b0 + stS[sidx] * 16384, sidx & 1, true, d_cross, acc_cross);
        }
#pragma unroll
        for (int sidx = 2; sidx < 4; ++sidx) {
          pass(b0 + stS[sidx] * 16384, sidx & 1, false, d_main, acc_main);
          umma::commit(b_empty + stS[sidx]);
        }
        bi += 4;
      };
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++ti) {
        const uint32_t slot = 0;
        umma::mbar_wait(a_full + slot, ti & 1);
        umma::fence_after_sync();
        const uint32_t aHi = a0 + slot * TG_A_BYTES, aLo = aHi + 32768;
        auto m1 = [&](uint32_t gg) {
          const uint32_t set = tbase + (gg & 1) * 128;
          contract(true, aHi, aLo, set, set, true);
          umma::commit(acc1_full + (gg & 1));
        };
        m1(g);
        for (int j = 0; j < nC; ++j, ++g) {
          if (j + 1 < nC) {
            m1(g + 1);
            if (j + 2 == nC) umma::commit(a_empty + slot);      // every M1 of the tile issued: the x1 block is free once they complete
          }
          umma::mbar_wait(hid_ready + (g & 1), (g >> 1) & 1);
          if (j == 0 && ti >= 1) umma::mbar_wait(acc2_empty, (ti - 1) & 1);
          umma::fence_after_sync();
          contract(false, tbase + (g & 1) * 128, 0, tbase + 256, tbase + 384, j == 0);
        }
        umma::commit(acc2_full);
      }
    }
  } else if (warp >= 4) {
    // ---------------- builders: x1 block of the tile -> fp16 hi/lo operand (slot = tile parity) ----------------
    const int bw = warp - 4, r8 = lane & 7, cj = bw * 4 + (lane >> 3);        // k in [8 cj, 8 cj + 8)
    uint32_t ti = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++ti) {
      const long long m0 = tile * 128;
      const uint32_t slot = 0;      // one x1 operand block: rewritten while the tile's last M2 contractions run
      uint8_t* aHi = aBuf + slot * TG_A_BYTES;
      uint8_t* aLo = aHi + 32768;
#pragma unroll
      for (int hb = 0; hb < 2; ++hb) {
        float4 v[16];
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const long long row = m0 + (hb * 8 + it) * 8 + r8;
          const float* src = X + row * E + cj * 8;
          if (row < M) {
            v[2 * it] = *reinterpret_cast<const float4*>(src);
            v[2 * it + 1] = *reinterpret_cast<const float4*>(src + 4);
          } else {
            v[2 * it] = v[2 * it + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        if (hb == 0 && ti >= 1) umma::mbar_wait(a_empty + slot, (ti - 1) & 1);
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          uint4 h, l;
          umma::split2_f16(v[2 * it].x, v[2 * it].y, h.x, l.x);
          umma::split2_f16(v[2 * it].z, v[2 * it].w, h.y, l.y);
          umma::split2_f16(v[2 * it + 1].x, v[2 * it + 1].y, h.z, l.z);
          umma::split2_f16(v[2 * it + 1].z, v[2 * it + 1].w, h.w, l.w);
          const uint32_t off = (uint32_t)cj * 2048u + (uint32_t)(hb * 8 + it) * 128u + (uint32_t)r8 * 16u;
          *reinterpret_cast<uint4*>(aHi + off) = h;
          *reinterpret_cast<uint4*>(aLo + off) = l;
        }
      }
      umma::fence_async_smem();
      tf_group_sync(1);
      if (tid == 128) tg_mbar_arrive(a_full + slot);
    }
  } else {
    // ---------------- epilogue warps: TMEM lane = row, warp = lane quadrant ----------------
    float* stg = reinterpret_cast<float*>(tsm + TF_OFF_STG + warp * 4096);     // [32 rows][8 float4], float4 index ^ (row & 7)
    const int er = lane >> 3, ec = lane & 7;
    const uint32_t tl = tbase + ((uint32_t)(32 * warp) << 16);
    float4 rr[8];                             // residual (x1) rows of the next 32-column chunk of the final epilogue
    auto fetch_res = [&](long long tile, int c) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const long long row = tile * 128 + 32 * warp + 4 * k + er;
        rr[k] = (tile < n_tiles && row < M) ? *reinterpret_cast<const float4*>(X + (size_t)row * E + c * 32 + 4 * ec)
                                            : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    uint32_t g = 0, ti = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++ti) {
      const long long myrow = tile * 128 + 32 * warp + lane;
      for (int j = 0; j < nC; ++j, ++g) {
        // ---- E(j): accumulator row -> + b1, ReLU -> fp16 hi/lo A operand, in place ----
        const uint32_t set = tl + (g & 1) * 128;
        umma::mbar_wait(acc1_full + (g & 1), (g >> 1) & 1);
        umma::fence_after_sync();
#pragma unroll 1
        for (int hh = 0; hh < 2; ++hh) {                // hidden columns [64 hh, 64 hh + 64) of the chunk
          uint32_t v[64];
          umma::ld32_nw(set + 64 * hh, v);
          umma::ld32_nw(set + 64 * hh + 32, v + 32);
          umma::wait_ld();
          uint32_t hw[32], lw[32];
          const float* bp = sB1 + j * 128 + 64 * hh;
          float* hp = hid_out && myrow < M ? hid_out + (size_t)myrow * FF + j * 128 + 64 * hh : nullptr;
#pragma unroll
          for (int c4 = 0; c4 < 16; ++c4) {
            const float4 bb = *reinterpret_cast<const float4*>(bp + 4 * c4);
            float4 o;
            o.x = fmaxf(umma::after_wait(v[4 * c4]) + bb.x, 0.f);
            o.y = fmaxf(umma::after_wait(v[4 * c4 + 1]) + bb.y, 0.f);
            o.z = fmaxf(umma::after_wait(v[4 * c4 + 2]) + bb.z, 0.f);
            o.w = fmaxf(umma::after_wait(v[4 * c4 + 3]) + bb.w, 0.f);
            if (hp) *reinterpret_cast<float4*>(hp + 4 * c4) = o;          // training keeps the hidden activations
            umma::split2_f16(o.x, o.y, hw[2 * c4], lw[2 * c4]);
            umma::split2_f16(o.z, o.w, hw[2 * c4 + 1], lw[2 * c4 + 1]);
          }
          umma::st16s<1>(set + 64 * hh, hw);
          umma::st16s<1>(set + 64 * hh + 16, hw + 16);
          umma::st16s<1>(set + 64 * hh + 32, lw);
          umma::st16s<1>(set + 64 * hh + 48, lw + 16);
        }
        umma::wait_st();
        umma::fence_before_sync();
        tf_group_sync(2);
        if (tid == 0) tg_mbar_arrive(hid_ready + (g & 1));
        if (j == nC - 2) fetch_res(tile, 0);            // residual rows of the final epilogue's first chunk
      }
      // ---- F: out = acc2 + acc2 cross + b2 + x1 ----
      umma::mbar_wait(acc2_full, ti & 1);
      umma::fence_after_sync();
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        const float4 bb = *reinterpret_cast<const float4*>(b2 + c * 32 + 4 * ec);
        uint32_t v[32], v2[32];
        umma::ld32_nw(tl + 256 + c * 32, v);
        umma::ld32_nw(tl + 384 + c * 32, v2);
        umma::wait_ld();
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) {
          float4 o;
          o.x = umma::after_wait(v[4 * i4]) + umma::after_wait(v2[4 * i4]);
          o.y = umma::after_wait(v[4 * i4 + 1]) + umma::after_wait(v2[4 * i4 + 1]);
          o.z = umma::after_wait(v[4 * i4 + 2]) + umma::after_wait(v2[4 * i4 + 2]);
          o.w = umma::after_wait(v[4 * i4 + 3]) + umma::after_wait(v2[4 * i4 + 3]);
          *reinterpret_cast<float4*>(stg + lane * 32 + 4 * (i4 ^ (lane & 7))) = o;
        }
        if (c == 3) {                                   // acc2 is in registers / staging: the next tile's M2(0) may overwrite it
          umma::fence_before_sync();
          tf_group_sync(2);
          if (tid == 0) tg_mbar_arrive(acc2_empty);
        }
        __syncwarp();
        float4 o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int r = 4 * k + er;
          o[k] = *reinterpret_cast<const float4*>(stg + r * 32 + 4 * (ec ^ (r & 7)));
          o[k].x += bb.x + rr[k].x; o[k].y += bb.y + rr[k].y; o[k].z += bb.z + rr[k].z; o[k].w += bb.w + rr[k].w;
        }
        if (c < 3) fetch_res(tile, c + 1);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const long long row = tile * 128 + 32 * warp + 4 * k + er;
          if (row < M) *reinterpret_cast<float4*>(out + (size_t)row * E + c * 32 + 4 * ec) = o[k];
        }
        __syncwarp();
      }
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tbase, 512);
}

static int tc_ffn(const float* x1, const float* derived, long long off_w1, long long off_w2, const float* b1, const float* b2,
                  float* out, float* hid_out, long long rows, int ff, cudaStream_t st) {
  ELG_REQUIRE(ff % 256 == 0 && ff <= TF_MAX_FF, ELG_EUNSUPPORTED, "tc_ffn: ff must be a multiple of 256, at most %d (got %d)", TF_MAX_FF, ff);
  static bool attr = false;
  static int sms = 0;
  if (!attr) {
    ELG_CUDA_OK(cudaFuncSetAttribute(tc_ffn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TF_SMEM));
    int dev = 0;
    ELG_CUDA_OK(cudaGetDevice(&dev));
    ELG_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    attr = true;
  }
  const long long tiles = (rows + 127) / 128;
  tc_ffn_kernel<<<(unsigned)(tiles < sms ? tiles : sms), TF_THREADS, TF_SMEM, st>>>(
      x1, reinterpret_cast<const uint8_t*>(derived + off_w1), b1, reinterpret_cast<const uint8_t*>(derived + off_w2), b2, out,
      hid_out, rows, ff);
  ELG_LAUNCH_OK();
  return ELG_OK;
}

// ---- encoder self-attention: one CTA per (aug-instance, head) --------------------------------------
// qkv rows are [q(128) | k(128) | v(128)]; scores q.k/4, softmax over keys, weighted values.
constexpr int ATT_TK = 512;     // keys staged per tile (64 KB of K+V)
__global__ void __launch_bounds__(128) enc_attention_kernel(const float* __restrict__ qkv, int N1,
                                                            float* __restrict__ att) {
  extern __shared__ __align__(16) float sm[];
  float* sk = sm;                      // [ATT_TK][16]
  float* sv = sm + ATT_TK * D;         // [ATT_TK][16]
  const int h = blockIdx.x % H;
  const long long b = blockIdx.x / H;
  const float* base = qkv + b * N1 * (3LL * E);
  for (int i0 = 0; i0 < N1; i0 += blockDim.x) {
    const int i = i0 + threadIdx.x;
    const bool ok = i < N1;
    float q[D], o[D];
    float m = -INFINITY, l = 0.f;
#pragma unroll
    for (int d4 = 0; d4 < D / 4; ++d4) {
      float4 t = ok ? *reinterpret_cast<const float4*>(base + (long long)i * 3 * E + h * D + d4 * 4) : make_float4(0, 0, 0, 0);
      q[d4 * 4] = t.x * 0.25f; q[d4 * 4 + 1] = t.y * 0.25f; q[d4 * 4 + 2] = t.z * 0.25f; q[d4 * 4 + 3] = t.w * 0.25f;
      o[d4 * 4] = o[d4 * 4 + 1] = o[d4 * 4 + 2] = o[d4 * 4 + 3] = 0.f;
    }
    for (int j0 = 0; j0 < N1; j0 += ATT_TK) {
      const int nj = min(ATT_TK, N1 - j0);
      __syncthreads();
      for (int t = threadIdx.x; t < nj * (D / 4); t += blockDim.x) {
        int j = t / (D / 4), d4 = t % (D / 4);
        const float* rowp = base + (long long)(j0 + j) * 3 * E + h * D + d4 * 4;
        *reinterpret_cast<float4*>(sk + j * D + d4 * 4) = *reinterpret_cast<const float4*>(rowp + E);
        *reinterpret_cast<float4*>(sv + j * D + d4 * 4) = *reinterpret_cast<const float4*>(rowp + 2 * E);
      }
      __syncthreads();
      for (int j = 0; j < nj; ++j) {
        float s = 0.f;
#pragma unroll
        for (int d4 = 0; d4 < D / 4; ++d4) {
          float4 kk = *reinterpret_cast<const float4*>(sk + j * D + d4 * 4);
          s = fmaf(q[d4 * 4], kk.x, s); s = fmaf(q[d4 * 4 + 1], kk.y, s);
          s = fmaf(q[d4 * 4 + 2], kk.z, s); s = fmaf(q[d4 * 4 + 3], kk.w, s);
        }
        if (s > m + 8.f) {           // lazy rescale: only when the running reference falls far behind
          float c = expf(m - s);
          l *= c;
#pragma unroll
          for (int d = 0; d < D; ++d) o[d] *= c;
          m = s;
        }
        float p = expf(s - m);
        l += p;
#pragma unroll
        for (int d4 = 0; d4 < D / 4; ++d4) {
          float4 vv = *reinterpret_cast<const float4*>(sv + j * D + d4 * 4);
          o[d4 * 4] = fmaf(p, vv.x, o[d4 * 4]); o[d4 * 4 + 1] = fmaf(p, vv.y, o[d4 * 4 + 1]);
          o[d4 * 4 + 2] = fmaf(p, vv.z, o[d4 * 4 + 2]); o[d4 * 4 + 3] = fmaf(p, vv.w, o[d4 * 4 + 3]);
        }
      }
    }
    if (ok) {
      const float inv = 1.f / l;
      float* op = att + (b * N1 + i) * E + h * D;
#pragma unroll
      for (int d4 = 0; d4 < D / 4; ++d4)
        *reinterpret_cast<float4*>(op + d4 * 4) =
            make_float4(o[d4 * 4] * inv, o[d4 * 4 + 1] * inv, o[d4 * 4 + 2] * inv, o[d4 * 4 + 3] * inv);
    }
  }
}

// ---- instance norm over the nodes of one aug-instance (input already holds x + sublayer(x)) -------
// nn.InstanceNorm1d(E, affine=True, track_running_stats=False): biased variance, eps 1e-5.
__global__ void __launch_bounds__(E) instance_norm_kernel(const float* __restrict__ t, const float* __restrict__ w,
                                                          const float* __restrict__ bsh, int N1,
                                                          float* __restrict__ out) {
  const long long b = blockIdx.x;
  const int c = threadIdx.x;
  const float* p = t + b * N1 * E + c;
  float s = 0.f;
  for (int n = 0; n < N1; ++n) s += p[(long long)n * E];
  const float mean = s / (float)N1;
  float v = 0.f;
  for (int n = 0; n < N1; ++n) {
    float d = p[(long long)n * E] - mean;
    v = fmaf(d, d, v);
  }
  const float rstd = 1.f / sqrtf(v / (float)N1 + 1e-5f);
  const float g = w[c], be = bsh[c];
  float* q = out + b * N1 * E + c;
  for (int n = 0; n < N1; ++n) q[(long long)n * E] = (p[(long long)n * E] - mean) * rstd * g + be;
}

// Resident instances (N1 <= 112): the thread's channel column of the instance is read ONCE into registers (all loads in
// flight together), statistics and output come from the registers -- same summation order as the kernel above, so the
// results are bit-identical; HBM traffic 1 read + 1 write instead of up to 3 reads + 1 write.
constexpr int IN_NMAX = 112;
__global__ void __launch_bounds__(E) instance_norm_reg_kernel(const float* __restrict__ t, const float* __restrict__ w,
                                                              const float* __restrict__ bsh, int N1,
                                                              float* __restrict__ out) {
  const long long b = blockIdx.x;
  const int c = threadIdx.x;
  const float* p = t + b * N1 * E + c;
  float x[IN_NMAX];
#pragma unroll
  for (int n = 0; n < IN_NMAX; ++n) x[n] = n < N1 ? p[(long long)n * E] : 0.f;
  float s = 0.f;
#pragma unroll
  for (int n = 0; n < IN_NMAX; ++n)
    if (n < N1) s += x[n];
  const float mean = s / (float)N1;
  float v = 0.f;
#pragma unroll
  for (int n = 0; n < IN_NMAX; ++n)
    if (n < N1) {
      const float d = x[n] - mean;
      v = fmaf(d, d, v);
    }
  const float rstd = 1.f / sqrtf(v / (float)N1 + 1e-5f);
  const float g = w[c], be = bsh[c];
  float* q = out + b * N1 * E + c;
#pragma unroll
  for (int n = 0; n < IN_NMAX; ++n)
    if (n < N1) q[(long long)n * E] = (x[n] - mean) * rstd * g + be;
}

static void launch_instance_norm(const float* t, const float* w, const float* bsh, int B, int N1, float* out, cudaStream_t st) {
  if (N1 <= IN_NMAX) instance_norm_reg_kernel<<<(unsigned)B, E, 0, st>>>(t, w, bsh, N1, out);
  else instance_norm_kernel<<<(unsigned)B, E, 0, st>>>(t, w, bsh, N1, out);
}

// ---- score bias eb[r] = enc[r] . (bo / sqrt(E)) : one warp per node row -----------------------------
__global__ void row_dot_kernel(const float* __restrict__ x, const float* __restrict__ v, long long rows,
                               float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long r = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  if (r >= rows) return;
  float4 a = *reinterpret_cast<const float4*>(x + r * E + lane * 4);
  float4 bq = *reinterpret_cast<const float4*>(v + lane * 4);
  float s = a.x * bq.x;
  s = fmaf(a.y, bq.y, s); s = fmaf(a.z, bq.z, s); s = fmaf(a.w, bq.w, s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[r] = s;
}

template <int EPI>
static int gemm(const float* A, const float* Bm, float* C, const float* bias, const float* res, long long M, int N,
                int K, int ldc, int N1, cudaStream_t st) {
  ELG_REQUIRE(N % 128 == 0 && K % 8 == 0, ELG_EUNSUPPORTED, "sgemm needs N%%128==0 and K%%8==0 (N=%d K=%d)", N, K);
  dim3 grid((unsigned)((M + 127) / 128), N / 128);
  sgemm_tn_kernel<EPI><<<grid, 256, 0, st>>>(A, Bm, C, bias, res, M, N, K, ldc, N1);
  ELG_LAUNCH_OK();
  return ELG_OK;
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
#define ELG_TRY(expr) do { int _rc = (expr); if (_rc) return _rc; } while (0)

}  // namespace elg

using namespace elg;

extern "C" {

size_t elg_encode_workspace_bytes(const elg_model_desc* d, int B, int N1) {
  if (check_desc(d) || B <= 0 || N1 <= 0) return 0;
  size_t rows = (size_t)B * N1;
  // x, x1, t (E each), att (E), qkv (3E), hid (ff)
  return align_up(rows * (size_t)(E * 4 + 3 * E + d->ff) * sizeof(float), 256) + 256;
}


// Shared body of elg_encode / elg_encode_train.  With `saved` != NULL every layer keeps its activations
// (layout: train_saved_* in common.cuh) for elg_reinforce_backward instead of reusing one workspace.
static int encode_impl(const elg_model_desc* d, const float* weights, const float* derived, const elg_tables* t, int B,
                       int N1, void* workspace, size_t workspace_bytes, float* saved, void* stream) {
  elg_weight_layout_t L;
  ELG_TRY(elg_weight_layout(d, &L));
  ELG_REQUIRE(weights && derived && t && (workspace || saved), ELG_EINVAL, "NULL pointer");
  ELG_REQUIRE(B > 0 && N1 > 1, ELG_EINVAL, "bad batch/node count");
  ELG_REQUIRE(t->xy && t->enc && t->k && t->v && t->e && t->eb && t->qtab, ELG_EINVAL, "elg_tables has NULL members");
  ELG_REQUIRE(d->problem == ELG_TSP ? t->qfirst != nullptr : t->demand != nullptr, ELG_EINVAL,
              "tsp needs qfirst, cvrp needs demand");
  ELG_REQUIRE(saved || workspace_bytes >= elg_encode_workspace_bytes(d, B, N1), ELG_ENOMEM, "workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const long long rows = (long long)B * N1;
  float *x, *x1, *tt, *att, *qkv, *hid, *tt2;
  if (saved) {
    x = saved + train_saved_off(0, rows, d->ff, TS_XIN);
    x1 = tt = att = qkv = hid = tt2 = nullptr;
  } else {
    x = reinterpret_cast<float*>(align_up((size_t)workspace, 256));
    x1 = x + rows * E;
    tt = x1 + rows * E;
    att = tt + rows * E;
    qkv = att + rows * E;
    hid = qkv + rows * 3 * E;
    tt2 = tt;
  }
  const float* w = weights;

  {
    long long total = rows * (E / 4);
    int grid = (int)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
    embed_kernel<<<grid, 256, 0, st>>>(d->problem, t->xy, t->demand, w, L, rows, N1, x);
    ELG_LAUNCH_OK();
  }
  static bool attr_set = false;
  const int att_smem = 2 * ATT_TK * D * (int)sizeof(float);
  if (!attr_set) {
    ELG_CUDA_OK(cudaFuncSetAttribute(enc_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, att_smem));
    attr_set = true;
  }
  for (int l = 0; l < d->layers; ++l) {
    const auto& y = L.layer[l];
    float* xout = (l == d->layers - 1) ? t->enc : x;
    if (saved) {
      x = saved + train_saved_off(l, rows, d->ff, TS_XIN);
      qkv = saved + train_saved_off(l, rows, d->ff, TS_QKV);
      att = saved + train_saved_off(l, rows, d->ff, TS_ATT);
      tt = saved + train_saved_off(l, rows, d->ff, TS_T1);
      x1 = saved + train_saved_off(l, rows, d->ff, TS_X1);
      hid = saved + train_saved_off(l, rows, d->ff, TS_HID);
      tt2 = saved + train_saved_off(l, rows, d->ff, TS_T2);
      if (l < d->layers - 1) xout = saved + train_saved_off(l + 1, rows, d->ff, TS_XIN);
    }
    const long long so = split_off_layer(l, d->ff);
    ELG_TRY(tc_gemm<EPI_NONE>(x, derived, so, qkv, nullptr, nullptr, rows, 3 * E, E, 3 * E, st));
    if (N1 <= 112) enc_attention_tc_kernel<<<(unsigned)(B * H), 128, 0, st>>>(qkv, N1, att);
    else enc_attention_kernel<<<(unsigned)(B * H), 128, att_smem, st>>>(qkv, N1, att);
    ELG_LAUNCH_OK();
    ELG_TRY(tc_gemm<EPI_BIAS_RES>(att, derived, so + 3LL * E * E, tt, w + y.bo, x, rows, E, E, E, st));
    launch_instance_norm(tt, w + y.n1w, w + y.n1b, B, N1, x1, st);
    ELG_LAUNCH_OK();
    if (d->ff % 256 == 0 && d->ff <= TF_MAX_FF) {      // fused feed-forward block: the hidden activations stay on chip
      ELG_TRY(tc_ffn(x1, derived, so + 4LL * E * E, so + 4LL * E * E + (long long)d->ff * E, w + y.b1, w + y.b2, tt2,
                     saved ? hid : nullptr, rows, d->ff, st));
    } else {
      ELG_TRY(tc_gemm<EPI_BIAS_RELU>(x1, derived, so + 4LL * E * E, hid, w + y.b1, nullptr, rows, d->ff, E, d->ff, st));
      ELG_TRY(tc_gemm<EPI_BIAS_RES>(hid, derived, so + 4LL * E * E + (long long)d->ff * E, tt2, w + y.b2, x1, rows, E, d->ff, E, st));
    }
    launch_instance_norm(tt2, w + y.n2w, w + y.n2b, B, N1, xout, st);
    ELG_LAUNCH_OK();
  }
  // decoder-side tables from the encoded nodes
  const float* enc = t->enc;
  const long long sd = split_off_dec(d->layers, d->ff);
  // decoder tables K', V, qtab (and qfirst, tsp) are products of the same input with four consecutive pre-split weight
  // matrices: one launch with n-tile i written to plane i when the caller's buffers are equally spaced (the engine's are)
  const int n_dec = d->problem == ELG_TSP ? 4 : 3;
  const long long plane = t->v - t->k;
  const bool fused_dec = plane >= (long long)rows * E && t->qtab - t->v == plane && (n_dec == 3 || t->qfirst - t->qtab == plane);
  if (fused_dec) {
    ELG_TRY(tc_gemm<EPI_NONE>(enc, derived, sd, t->k, nullptr, nullptr, rows, n_dec * E, E, E, st, plane));
  } else {
    ELG_TRY(tc_gemm<EPI_NONE>(enc, derived, sd, t->k, nullptr, nullptr, rows, E, E, E, st));
    ELG_TRY(tc_gemm<EPI_NONE>(enc, derived, sd + 1LL * E * E, t->v, nullptr, nullptr, rows, E, E, E, st));
  }
  if (rollout_is_resident(d, N1)) {
    ELG_TRY(gemm<EPI_UMMA_SPLIT>(enc, derived + DER_WET, reinterpret_cast<float*>(t->e), nullptr, nullptr, rows, E, E, E, N1, st));
    split_kv_kernel<<<(unsigned)B, 256, 0, st>>>(t->k, t->v, reinterpret_cast<uint8_t*>(t->e), N1);
    ELG_LAUNCH_OK();
  } else {
    ELG_TRY(gemm<EPI_SWIZZLE>(enc, derived + DER_WET, reinterpret_cast<float*>(t->e), nullptr, nullptr, rows, E, E, E, N1, st));
    if (t->et) {
      split_tiles_kernel<<<(unsigned)(B * ((N1 + ELG_TILE_NODES - 1) / ELG_TILE_NODES)), 256, 0, st>>>(
          t->k, t->v, reinterpret_cast<const float*>(t->e), reinterpret_cast<uint8_t*>(t->et), N1);
      ELG_LAUNCH_OK();
    }
  }
  if (!fused_dec) {
    ELG_TRY(tc_gemm<EPI_NONE>(enc, derived, sd + 2LL * E * E, t->qtab, nullptr, nullptr, rows, E, E, E, st));
    if (d->problem == ELG_TSP)
      ELG_TRY(tc_gemm<EPI_NONE>(enc, derived, sd + 3LL * E * E, t->qfirst, nullptr, nullptr, rows, E, E, E, st));
  }
  row_dot_kernel<<<(unsigned)((rows * 32 + 255) / 256), 256, 0, st>>>(enc, derived + DER_BE, rows, t->eb);
  ELG_LAUNCH_OK();
  if (t->nbr) ELG_TRY(launch_neighbours(d, t->xy, t->demand, B, N1, t->nbr, st));
  if (saved)   // plain fp32 E' = enc (Wo/sqrt(E)) for the backward's forward recompute
    ELG_TRY(gemm<EPI_NONE>(enc, derived + DER_WET, saved + train_saved_eplain(d->layers, rows, d->ff), nullptr, nullptr, rows, E, E, E, N1, st));
  return ELG_OK;
}

int elg_encode(const elg_model_desc* d, const float* weights, const float* derived, const elg_tables* t, int B,
               int N1, void* workspace, size_t workspace_bytes, void* stream) {
  ELG_REQUIRE(workspace, ELG_EINVAL, "NULL workspace");
  return encode_impl(d, weights, derived, t, B, N1, workspace, workspace_bytes, nullptr, stream);
}

size_t elg_train_saved_bytes(const elg_model_desc* d, int B, int N1) {
  if (check_desc(d) || B <= 0 || N1 <= 0) return 0;
  return (size_t)train_saved_total(d->layers, (long long)B * N1, d->ff) * sizeof(float);
}

int elg_encode_train(const elg_model_desc* d, const float* weights, const float* derived, const elg_tables* t, int B,
                     int N1, void* saved, size_t saved_bytes, void* stream) {
  ELG_REQUIRE(saved && ((size_t)saved & 15) == 0, ELG_EINVAL, "saved-activation buffer must be 16-byte aligned");
  ELG_REQUIRE(saved_bytes >= elg_train_saved_bytes(d, B, N1), ELG_ENOMEM, "saved-activation buffer too small");
  return encode_impl(d, weights, derived, t, B, N1, nullptr, 0, reinterpret_cast<float*>(saved), stream);
}

}  // extern "C"
