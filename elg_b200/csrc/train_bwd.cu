// Training path: REINFORCE objective, backward through decoder tables and encoder, Adam.
//
// reference: the training-step body CVRP/train.py:104-125 (TSP/train.py:100-122): rollout in sample mode, POMO shared
// baseline, max-advantage scaling, J.backward(), torch.optim.Adam(lr, weight_decay=1e-6).step().
// What autograd differentiates there is re-derived here by hand: the decode step (train_decode.cu), the decoder-side
// tables (K, V, single-head key with the combine layer folded in, query tables) and the encoder
// (CVRP/models.py:199-269,506-562).
#include "train.cuh"

namespace elg {

// ---------------------------------------------------------------------------------------------------------------
// Generic strided / batched SGEMM (fp32 FMA), BK = 16, 256 threads, (BM/TM) x (BN/TN) thread grid, register-prefetched
// k-tiles (the next tile's global loads overlap the current tile's arithmetic).
// ---------------------------------------------------------------------------------------------------------------
template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(256) gemm_kernel(GemmP p) {
  constexpr int BK = 16;
  static_assert((BM / TM) * (BN / TN) == 256, "thread grid");
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int batch = blockIdx.z / p.splits, split = blockIdx.z % p.splits;
  const int b1 = batch / p.nb2, b2 = batch % p.nb2;
  const float* A = p.A + b1 * p.bA1 + b2 * p.bA2;
  const float* B = p.B ? p.B + b1 * p.bB1 + b2 * p.bB2 : nullptr;
  float* C = p.C + b1 * p.bC1 + b2 * p.bC2;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  int kper = (p.K + p.splits - 1) / p.splits;
  kper = (kper + BK - 1) / BK * BK;
  const int kbeg = split * kper;
  const int kend = min(p.K, kbeg + kper);
  const int ty = tid / (BN / TN), tx = tid % (BN / TN);
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  const bool a_kfast = p.sAk == 1;
  const bool b_nfast = p.sBn == 1;
  constexpr int NA = BM * BK / 256, NB = (BN * BK + 255) / 256;
  float ra[NA], rb[NB];
  // global -> registers for the k-tile starting at k0 (zero outside the matrix / this split's k-range)
  auto fetch = [&](int k0) {
#pragma unroll
    for (int u = 0; u < NA; ++u) {
      const int e = tid + u * 256;
      int m, k;
      if (a_kfast) { k = e % BK; m = e / BK; } else { m = e % BM; k = e / BM; }
      const int gm = m0 + m, gk = k0 + k;
      ra[u] = (gm < p.M && gk < kend) ? A[gm * p.sAm + gk * p.sAk] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < NB; ++u) {
      const int e = tid + u * 256;
      int n, k;
      if (b_nfast) { n = e % BN; k = e / BN; } else { k = e % BK; n = e / BK; }
      const int gn = n0 + n, gk = k0 + k;
      rb[u] = (e < BN * BK && gn < p.N && gk < kend) ? (B ? B[gk * p.sBk + gn * p.sBn] : 1.f) : 0.f;
    }
  };
  auto stage = [&]() {
#pragma unroll
    for (int u = 0; u < NA; ++u) {
      const int e = tid + u * 256;
      if (a_kfast) As[e % BK][e / BK] = ra[u]; else As[e / BM][e % BM] = ra[u];
    }
#pragma unroll
    for (int u = 0; u < NB; ++u) {
      const int e = tid + u * 256;
      if (e < BN * BK) { if (b_nfast) Bs[e / BN][e % BN] = rb[u]; else Bs[e % BK][e / BK] = rb[u]; }
    }
  };
  if (kbeg < kend) fetch(kbeg);
  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    stage();
    __syncthreads();
    if (k0 + BK < kend) fetch(k0 + BK);      // next tile's loads are in flight during this tile's arithmetic
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[k][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[k][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int gm = m0 + ty * TM + i;
    if (gm >= p.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int gn = n0 + tx * TN + j;
      if (gn >= p.N) continue;
      float* c = C + (long long)gm * p.ldc + gn;
      const float v = p.alpha * acc[i][j];
      if (p.splits > 1) atomicAdd(c, v);
      else if (p.accumulate) *c += v;
      else *c = v;
    }
  }
}

int launch_gemm(const GemmP& p, cudaStream_t st) {
  ELG_REQUIRE(p.A && p.C && p.M > 0 && p.N > 0 && p.K > 0 && p.splits >= 1 && p.nb1 >= 1 && p.nb2 >= 1, ELG_EINVAL, "bad gemm arguments");
  ELG_REQUIRE(p.splits == 1 || p.accumulate, ELG_EINVAL, "split-K gemm accumulates into C");
  const unsigned z = (unsigned)(p.nb1 * p.nb2 * p.splits);
  ELG_REQUIRE(z <= 65535u, ELG_EINVAL, "gemm batch x splits too large");
  if (p.N <= 16) {
    dim3 grid((p.M + 127) / 128, (p.N + 15) / 16, z);
    gemm_kernel<128, 16, 4, 2><<<grid, 256, 0, st>>>(p);
  } else {
    dim3 grid((p.M + 63) / 64, (p.N + 63) / 64, z);
    gemm_kernel<64, 64, 4, 4><<<grid, 256, 0, st>>>(p);
  }
  ELG_LAUNCH_OK();
  return ELG_OK;
}

// C[rows][N] (+)= alpha * A[rows][K] * W, W row-major [K][N]  ("dX = dY W")
static int gemm_nn(const float* A, int lda, const float* W, int ldw, float* C, int ldc, long long rows, int N, int K,
                   float alpha, int accumulate, cudaStream_t st) {
  GemmP p;
  p.A = A; p.B = W; p.C = C; p.M = (int)rows; p.N = N; p.K = K;
  p.sAm = lda; p.sAk = 1; p.sBk = ldw; p.sBn = 1; p.ldc = ldc; p.alpha = alpha; p.accumulate = accumulate;
  return launch_gemm(p, st);
}
// C[rows][N] (+)= alpha * A[rows][K] * W^T, W row-major [N][K]
static int gemm_nt(const float* A, int lda, const float* W, int ldw, float* C, int ldc, long long rows, int N, int K,
                   float alpha, int accumulate, cudaStream_t st) {
  GemmP p;
  p.A = A; p.B = W; p.C = C; p.M = (int)rows; p.N = N; p.K = K;
  p.sAm = lda; p.sAk = 1; p.sBk = 1; p.sBn = ldw; p.ldc = ldc; p.alpha = alpha; p.accumulate = accumulate;
  return launch_gemm(p, st);
}
// C[M][N] += alpha * A^T B with A [rows][M] (lda), B [rows][N] (ldb): weight gradients, split over the rows
static int gemm_tn_acc(const float* A, int lda, const float* Bm, int ldb, float* C, int ldc, int M, int N, long long rows,
                       float alpha, cudaStream_t st) {
  GemmP p;
  p.A = A; p.B = Bm; p.C = C; p.M = M; p.N = N; p.K = (int)rows;
  p.sAm = 1; p.sAk = lda; p.sBk = ldb; p.sBn = 1; p.ldc = ldc; p.alpha = alpha; p.accumulate = 1;
  const int tiles = ((M + 63) / 64) * ((N + 63) / 64);
  int splits = (int)((rows + 95) / 96);
  const int cap = (148 * 4 + tiles - 1) / tiles;
  if (splits > cap) splits = cap;
  if (splits < 1) splits = 1;
  p.splits = splits;
  return launch_gemm(p, st);
}

// ---------------------------------------------------------------------------------------------------------------
// small kernels
// ---------------------------------------------------------------------------------------------------------------
// out[c] += sum_r x[r][c]
__global__ void colsum_kernel(const float* __restrict__ x, long long rows, int N, int ld, int rows_per_cta,
                              float* __restrict__ out) {
  const long long r0 = (long long)blockIdx.x * rows_per_cta;
  const long long r1 = min(rows, r0 + rows_per_cta);
  for (int c = threadIdx.x; c < N; c += blockDim.x) {
    float s = 0.f;
    for (long long r = r0; r < r1; ++r) s += x[r * ld + c];
    atomicAdd(out + c, s);
  }
}
static int colsum(const float* x, long long rows, int N, int ld, float* out, cudaStream_t st) {
  const int per = 64;
  colsum_kernel<<<(unsigned)((rows + per - 1) / per), 128, 0, st>>>(x, rows, N, ld, per, out);
  ELG_LAUNCH_OK();
  return ELG_OK;
}

// g *= (h > 0)
__global__ void relu_bwd_kernel(float* __restrict__ g, const float* __restrict__ h, long long n4) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 gv = reinterpret_cast<float4*>(g)[i];
    const float4 hv = reinterpret_cast<const float4*>(h)[i];
    gv.x = hv.x > 0.f ? gv.x : 0.f; gv.y = hv.y > 0.f ? gv.y : 0.f;
    gv.z = hv.z > 0.f ? gv.z : 0.f; gv.w = hv.w > 0.f ? gv.w : 0.f;
    reinterpret_cast<float4*>(g)[i] = gv;
  }
}

// dst[i * stride] += src[i]
__global__ void add_strided_kernel(float* __restrict__ dst, int stride, const float* __restrict__ src, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[(long long)i * stride] += src[i];
}

// InstanceNorm1d backward over the nodes of one aug-instance (thread = channel), in place on g.
//   xhat = (t - mean) rstd;  d gamma += sum g xhat;  d beta += sum g;
//   d t = rstd (g gamma - mean_n(g gamma) - xhat mean_n(g gamma xhat))
__global__ void __launch_bounds__(E) instance_norm_bwd_kernel(const float* __restrict__ t, float* __restrict__ g,
                                                              const float* __restrict__ gamma, int N1,
                                                              float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const long long b = blockIdx.x;
  const int c = threadIdx.x;
  const float* p = t + b * N1 * E + c;
  float* q = g + b * N1 * E + c;
  float s = 0.f;
  for (int n = 0; n < N1; ++n) s += p[(long long)n * E];
  const float mean = s / (float)N1;
  float v = 0.f;
  for (int n = 0; n < N1; ++n) {
    const float d = p[(long long)n * E] - mean;
    v = fmaf(d, d, v);
  }
  const float rstd = 1.f / sqrtf(v / (float)N1 + 1e-5f);
  const float gm = gamma[c];
  float s1 = 0.f, s2 = 0.f;
  for (int n = 0; n < N1; ++n) {
    const float xh = (p[(long long)n * E] - mean) * rstd;
    const float gy = q[(long long)n * E];
    s1 += gy;
    s2 = fmaf(gy, xh, s2);
  }
  atomicAdd(dgamma + c, s2);
  atomicAdd(dbeta + c, s1);
  const float m1 = s1 * gm / (float)N1, m2 = s2 * gm / (float)N1;
  for (int n = 0; n < N1; ++n) {
    const float xh = (p[(long long)n * E] - mean) * rstd;
    q[(long long)n * E] = rstd * (q[(long long)n * E] * gm - m1 - xh * m2);
  }
}

// Encoder self-attention backward, one CTA per (aug-instance, head); q, k, v, dO of the head in shared memory.
// Phase 1 (thread = query i): row max / denominator, D_i = dO_i . O_i, then dq_i.  Phase 2 (thread = key j):
// dk_j, dv_j, recomputing p_ij from the row statistics.  No atomics, deterministic.
__global__ void __launch_bounds__(128) enc_attention_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ att,
                                                                const float* __restrict__ gatt, int N1,
                                                                float* __restrict__ gqkv) {
  extern __shared__ __align__(16) float asm_[];
  float* sq = asm_;
  float* sk = sq + N1 * D;
  float* sv = sk + N1 * D;
  float* sg = sv + N1 * D;
  float* smx = sg + N1 * D;
  float* sl = smx + N1;
  float* sD = sl + N1;
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const float* base = qkv + (size_t)b * N1 * 3 * E + h * D;
  for (int i = threadIdx.x; i < N1 * D; i += blockDim.x) {
    const int n = i / D, d = i % D;
    sq[i] = base[(size_t)n * 3 * E + d];
    sk[i] = base[(size_t)n * 3 * E + E + d];
    sv[i] = base[(size_t)n * 3 * E + 2 * E + d];
    sg[i] = gatt[((size_t)b * N1 + n) * E + h * D + d];
  }
  __syncthreads();
  const float scale = 0.25f;   // 1 / sqrt(D)
  for (int i = threadIdx.x; i < N1; i += blockDim.x) {
    float q[D], go[D];
#pragma unroll
    for (int d = 0; d < D; ++d) { q[d] = sq[i * D + d]; go[d] = sg[i * D + d]; }
    float mx = -INFINITY;
    for (int j = 0; j < N1; ++j) {
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < D; ++d) s = fmaf(q[d], sk[j * D + d], s);
      mx = fmaxf(mx, s * scale);
    }
    float l = 0.f;
    for (int j = 0; j < N1; ++j) {
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < D; ++d) s = fmaf(q[d], sk[j * D + d], s);
      l += __expf(s * scale - mx);
    }
    float Di = 0.f;
    const float* ao = att + ((size_t)b * N1 + i) * E + h * D;
#pragma unroll
    for (int d = 0; d < D; ++d) Di = fmaf(go[d], ao[d], Di);
    smx[i] = mx; sl[i] = 1.f / l; sD[i] = Di;
    float dq[D];
#pragma unroll
    for (int d = 0; d < D; ++d) dq[d] = 0.f;
    const float il = 1.f / l;
    for (int j = 0; j < N1; ++j) {
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int d = 0; d < D; ++d) { s = fmaf(q[d], sk[j * D + d], s); dp = fmaf(go[d], sv[j * D + d], dp); }
      const float pij = __expf(s * scale - mx) * il;
      const float dsv = pij * (dp - Di) * scale;
#pragma unroll
      for (int d = 0; d < D; ++d) dq[d] = fmaf(dsv, sk[j * D + d], dq[d]);
    }
    float* out = gqkv + ((size_t)b * N1 + i) * 3 * E + h * D;
#pragma unroll
    for (int d = 0; d < D; ++d) out[d] = dq[d];
  }
  __syncthreads();
  for (int j = threadIdx.x; j < N1; j += blockDim.x) {
    float k[D], v[D], dk[D], dv[D];
#pragma unroll
    for (int d = 0; d < D; ++d) { k[d] = sk[j * D + d]; v[d] = sv[j * D + d]; dk[d] = 0.f; dv[d] = 0.f; }
    for (int i = 0; i < N1; ++i) {
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int d = 0; d < D; ++d) { s = fmaf(sq[i * D + d], k[d], s); dp = fmaf(sg[i * D + d], v[d], dp); }
      const float pij = __expf(s * scale - smx[i]) * sl[i];
      const float dsv = pij * (dp - sD[i]) * scale;
#pragma unroll
      for (int d = 0; d < D; ++d) { dk[d] = fmaf(dsv, sq[i * D + d], dk[d]); dv[d] = fmaf(pij, sg[i * D + d], dv[d]); }
    }
    float* out = gqkv + ((size_t)b * N1 + j) * 3 * E + h * D;
#pragma unroll
    for (int d = 0; d < D; ++d) { out[E + d] = dk[d]; out[2 * E + d] = dv[d]; }
  }
}

// embedding backward: thread = channel; each CTA reduces a block of node rows
__global__ void __launch_bounds__(E) embed_bwd_kernel(int problem, const float* __restrict__ xy, const float* __restrict__ demand,
                                                      const float* __restrict__ g, long long rows, int N1, int rows_per_cta,
                                                      elg_weight_layout_t L, float* __restrict__ grads) {
  const int c = threadIdx.x;
  const long long r0 = (long long)blockIdx.x * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
  float dn[3] = {0.f, 0.f, 0.f}, dnb = 0.f, dd[2] = {0.f, 0.f}, ddb = 0.f;
  for (long long r = r0; r < r1; ++r) {
    const float gv = g[r * E + c];
    const float px = xy[2 * r], py = xy[2 * r + 1];
    if (problem == ELG_CVRP && (r % N1) == 0) {
      dd[0] = fmaf(gv, px, dd[0]); dd[1] = fmaf(gv, py, dd[1]); ddb += gv;
    } else {
      dn[0] = fmaf(gv, px, dn[0]); dn[1] = fmaf(gv, py, dn[1]); dnb += gv;
      if (problem == ELG_CVRP) dn[2] = fmaf(gv, demand[r], dn[2]);
    }
  }
  if (problem == ELG_CVRP) {
    atomicAdd(grads + L.emb_depot_w + c * 2, dd[0]); atomicAdd(grads + L.emb_depot_w + c * 2 + 1, dd[1]);
    atomicAdd(grads + L.emb_depot_b + c, ddb);
    atomicAdd(grads + L.emb_node_w + c * 3, dn[0]); atomicAdd(grads + L.emb_node_w + c * 3 + 1, dn[1]);
    atomicAdd(grads + L.emb_node_w + c * 3 + 2, dn[2]);
  } else {
    atomicAdd(grads + L.emb_node_w + c * 2, dn[0]); atomicAdd(grads + L.emb_node_w + c * 2 + 1, dn[1]);
  }
  atomicAdd(grads + L.emb_node_b + c, dnb);
}

// POMO shared baseline + max-advantage scaling (CVRP/train.py:113-121; TSP/train.py:109-119): one CTA per instance.
__global__ void advantage_stats_kernel(const float* __restrict__ reward, int M, float* __restrict__ mean_out,
                                       float* __restrict__ max_out, int* __restrict__ any_zero) {
  __shared__ float red[32];
  const int b = blockIdx.x;
  const float* r = reward + (size_t)b * M;
  float s = 0.f;
  for (int m = threadIdx.x; m < M; m += blockDim.x) s += r[m];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  float tot = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) tot += red[i];
  const float mean = tot / (float)M;
  __syncthreads();
  float mx = -INFINITY;
  for (int m = threadIdx.x; m < M; m += blockDim.x) mx = fmaxf(mx, r[m] - mean);
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m2 = -INFINITY;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) m2 = fmaxf(m2, red[i]);
    mean_out[b] = mean;
    max_out[b] = m2;
    if (m2 == 0.f) atomicOr(any_zero, 1);
  }
}
__global__ void advantage_coef_kernel(int problem, const float* __restrict__ reward, const float* __restrict__ logp,
                                      const float* __restrict__ mean, const float* __restrict__ mx,
                                      const int* __restrict__ any_zero, int B, int M, int scale_norm,
                                      float* __restrict__ coef, float* __restrict__ loss) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  float contrib = 0.f;
  if (g < B * M) {
    const int b = g / M;
    float c = -(reward[g] - mean[b]);
    if (scale_norm && (problem == ELG_CVRP || !*any_zero)) c = c / mx[b];
    c = c / (float)(B * M);
    coef[g] = c;
    if (logp) contrib = c * logp[g];
  }
  for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
  if (loss && (threadIdx.x & 31) == 0) atomicAdd(loss, contrib);
}

// torch.optim.Adam (L2 weight decay folded into the gradient, bias corrections from the host in double precision)
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            long long n, float step_size, float bc2_sqrt, float beta1, float beta2, float eps,
                            float weight_decay, float grad_scale) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float pv = p[i];
    const float gv = fmaf(weight_decay, pv, g[i] * grad_scale);
    const float mv = beta1 * m[i] + (1.f - beta1) * gv;
    const float vv = beta2 * v[i] + (1.f - beta2) * gv * gv;
    m[i] = mv;
    v[i] = vv;
    p[i] = pv - step_size * (mv / (sqrtf(vv) / bc2_sqrt + eps));
  }
}

static inline size_t al256(size_t v) { return (v + 255) & ~(size_t)255; }
#define ELG_TRY(expr) do { int _rc = (expr); if (_rc) return _rc; } while (0)

// workspace layout (floats, every part 64-float aligned); the first parts are documented in the header so that tests
// can read the table gradients back
struct TrainWs {
  long long dEp, dK, dV, dqtab, dqfirst, deb, dwl, lg, stats, gx, ghid, gatt, gqkv, rec, coef, chunk, total;
};
static TrainWs train_ws(const elg_model_desc* d, int B, int M, int N1, int t_max) {
  TrainWs w;
  const long long rows = (long long)B * N1;
  long long o = 0;
  auto take = [&](long long n) { long long r = o; o += (n + 63) & ~63LL; return r; };
  w.dEp = take(rows * E); w.dK = take(rows * E); w.dV = take(rows * E); w.dqtab = take(rows * E);
  w.dqfirst = take(d->problem == ELG_TSP ? rows * E : 0);
  w.deb = take(rows); w.dwl = take(E); w.lg = take(LG_TOTAL);
  w.stats = take(2LL * B + 64);
  w.gx = take(rows * E); w.ghid = take(rows * d->ff); w.gatt = take(rows * E); w.gqkv = take(rows * 3 * E);
  w.rec = take((long long)B * t_max * M * (long long)(sizeof(StepRec) / sizeof(float)));
  w.coef = take((long long)B * M);
  w.chunk = o;
  w.total = o;
  return w;
}
static inline long long chunk_row_floats(int NP) { return 2LL * NP; }

}  // namespace elg

using namespace elg;

extern "C" {

size_t elg_train_workspace_bytes(const elg_model_desc* d, int B, int M, int N1, int t_max, int chunk_steps) {
  if (check_desc(d) || B <= 0 || M <= 0 || N1 <= 1 || t_max <= 0 || chunk_steps <= 0) return 0;
  const TrainWs w = train_ws(d, B, M, N1, t_max);
  const int NP = (N1 + 3) & ~3;
  return (size_t)(w.chunk + (long long)B * chunk_steps * M * chunk_row_floats(NP) + 64) * sizeof(float);
}

int elg_train_workspace_layout(const elg_model_desc* d, int B, int M, int N1, int t_max, int64_t* out8) {
  ELG_TRY(check_desc(d));
  ELG_REQUIRE(out8, ELG_EINVAL, "NULL out");
  const TrainWs w = train_ws(d, B, M, N1, t_max);
  out8[0] = w.dEp; out8[1] = w.dK; out8[2] = w.dV; out8[3] = w.dqtab; out8[4] = w.dqfirst; out8[5] = w.deb;
  out8[6] = w.dwl; out8[7] = w.lg;
  return ELG_OK;
}

int elg_reinforce_backward(const elg_model_desc* d, const float* weights, const float* derived, const elg_tables* t,
                           const void* saved, int B, int M, int N1, const int16_t* tours, int t_max, int T,
                           const float* reward, const float* logp, int scale_norm, float* grads, float* loss,
                           void* workspace, size_t workspace_bytes, void* stream) {
  elg_weight_layout_t L;
  ELG_TRY(elg_weight_layout(d, &L));
  ELG_REQUIRE(weights && derived && t && saved && tours && reward && grads && workspace, ELG_EINVAL, "NULL pointer");
  ELG_REQUIRE(B > 0 && M > 0 && N1 > 1 && T > 0 && T <= t_max, ELG_EINVAL, "bad sizes");
  ELG_REQUIRE(N1 <= TRAIN_MAX_NODES && elg_rollout_resident(d, N1) == 1, ELG_EUNSUPPORTED,
              "the training path supports resident instances only (elg_rollout_resident: up to %d nodes, 108 for cvrp with k = 40)", TRAIN_MAX_NODES);
  ELG_REQUIRE(((size_t)workspace & 255) == 0, ELG_EINVAL, "workspace must be 256-byte aligned");
  ELG_REQUIRE(workspace_bytes >= elg_train_workspace_bytes(d, B, M, N1, t_max, 1), ELG_ENOMEM, "workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const bool cvrp = d->problem == ELG_CVRP;
  const long long rows = (long long)B * N1;
  const int NP = (N1 + 3) & ~3;
  const TrainWs w = train_ws(d, B, M, N1, t_max);
  float* ws = reinterpret_cast<float*>(workspace);
  const float* sv = reinterpret_cast<const float*>(saved);
  const long long avail = (long long)(workspace_bytes / sizeof(float)) - w.chunk - 64;
  int nT = (int)(avail / ((long long)B * M * chunk_row_floats(NP)));
  if (nT > 256) nT = 256;
  if (nT > T) nT = T;
  ELG_REQUIRE(nT >= 1, ELG_ENOMEM, "workspace too small for one step");

  // ---- zero the accumulators and the gradient
  ELG_CUDA_OK(cudaMemsetAsync(ws, 0, (size_t)w.gx * sizeof(float), st));            // table gradients, dwl, lg, stats
  ELG_CUDA_OK(cudaMemsetAsync(grads, 0, (size_t)L.total * sizeof(float), st));
  if (loss) ELG_CUDA_OK(cudaMemsetAsync(loss, 0, sizeof(float), st));
  float* mean = ws + w.stats;
  float* mx = mean + B;
  int* any_zero = reinterpret_cast<int*>(mx + B);
  float* coef = ws + w.coef;
  advantage_stats_kernel<<<B, 128, 0, st>>>(reward, M, mean, mx, any_zero);
  ELG_LAUNCH_OK();
  advantage_coef_kernel<<<(B * M + 127) / 128, 128, 0, st>>>(d->problem, reward, logp, mean, mx, any_zero, B, M, scale_norm, coef, loss);
  ELG_LAUNCH_OK();
  StepRec* rec = reinterpret_cast<StepRec*>(ws + w.rec);
  ELG_TRY(launch_replay(d->problem, t->demand, tours, t_max, B, M, N1, T, rec, st));

  // ---- decode backward, a chunk of steps at a time
  DecodeBwdArgs a;
  a.t = *t; a.weights = weights; a.L = L; a.derived = derived;
  a.eplain = sv + train_saved_eplain(d->layers, rows, d->ff);
  a.problem = d->problem; a.B = B; a.M = M; a.N1 = N1; a.NP = NP; a.k_local = d->local_k; a.flags = d->flags;
  a.xi = d->xi; a.clip = d->clip; a.rec = rec; a.T = T; a.coef = coef;
  a.dqtab = ws + w.dqtab; a.dqfirst = cvrp ? nullptr : ws + w.dqfirst; a.dwl = ws + w.dwl; a.lg = ws + w.lg;
  a.dV = ws + w.dV; a.dK = ws + w.dK; a.deb = ws + w.deb; a.dEp = ws + w.dEp;
  const long long crow = (long long)B * nT * M;
  float* cb = ws + w.chunk;
  a.add = cb; cb += crow * NP;
  a.dx = cb;
  for (int t0 = cvrp ? 2 : 1; t0 < T; t0 += nT) {
    a.t0 = t0;
    a.nT = (T - t0) < nT ? (T - t0) : nT;
    ELG_TRY(launch_local(a, false, st));
    ELG_TRY(launch_global_bwd(a, st));
    if (d->flags & ELG_FLAG_ENSEMBLE) ELG_TRY(launch_local(a, true, st));      // no local policy: its parameters get no gradient
  }
  if (d->flags & ELG_FLAG_ENSEMBLE) ELG_TRY(launch_local_fold_bwd(d, L, weights, derived, ws + w.lg, grads, st));

  // ---- decoder-side tables -> encoded nodes and decoder weights
  const float* enc = t->enc;
  float* gx = ws + w.gx;
  const float ise = 0.08838834764831845f;       // 1 / sqrt(E)
  const float ck = 0.36067376022224085f;        // log2(e) / sqrt(D)
  const int ldq = cvrp ? E + 1 : E;
  // E'[j][i] = sum_k enc[j][k] Wo[k][i] / sqrt(E);  eb[j] = enc[j] . bo / sqrt(E)
  ELG_TRY(gemm_nt(ws + w.dEp, E, weights + L.dec_wo, E, gx, E, rows, E, E, ise, 0, st));
  ELG_TRY(gemm_tn_acc(enc, E, ws + w.dEp, E, grads + L.dec_wo, E, E, E, rows, ise, st));
  ELG_TRY(gemm_nn(ws + w.deb, 1, weights + L.dec_bo, E, gx, E, rows, E, 1, ise, 1, st));
  ELG_TRY(gemm_tn_acc(enc, E, ws + w.deb, 1, grads + L.dec_bo, 1, E, 1, rows, ise, st));
  // K' = enc (ck Wk)^T, V = enc Wv^T
  ELG_TRY(gemm_nn(ws + w.dK, E, weights + L.dec_wk, E, gx, E, rows, E, E, ck, 1, st));
  ELG_TRY(gemm_tn_acc(ws + w.dK, E, enc, E, grads + L.dec_wk, E, E, E, rows, ck, st));
  ELG_TRY(gemm_nn(ws + w.dV, E, weights + L.dec_wv, E, gx, E, rows, E, E, 1.f, 1, st));
  ELG_TRY(gemm_tn_acc(ws + w.dV, E, enc, E, grads + L.dec_wv, E, E, E, rows, 1.f, st));
  // qtab = enc Wq_node^T (+ load * w_load per row, accumulated in dwl)
  ELG_TRY(gemm_nn(ws + w.dqtab, E, weights + L.dec_wq_last, ldq, gx, E, rows, E, E, 1.f, 1, st));
  ELG_TRY(gemm_tn_acc(ws + w.dqtab, E, enc, E, grads + L.dec_wq_last, ldq, E, E, rows, 1.f, st));
  if (cvrp) {
    add_strided_kernel<<<1, E, 0, st>>>(grads + L.dec_wq_last + E, ldq, ws + w.dwl, E);
    ELG_LAUNCH_OK();
  } else {
    ELG_TRY(gemm_nn(ws + w.dqfirst, E, weights + L.dec_wq_first, E, gx, E, rows, E, E, 1.f, 1, st));
    ELG_TRY(gemm_tn_acc(ws + w.dqfirst, E, enc, E, grads + L.dec_wq_first, E, E, E, rows, 1.f, st));
  }

  // ---- encoder backward
  float* ghid = ws + w.ghid;
  float* gatt = ws + w.gatt;
  float* gqkv = ws + w.gqkv;
  const int ff = d->ff;
  const size_t att_smem = (size_t)(4 * N1 * D + 3 * N1) * sizeof(float);
  ELG_CUDA_OK(cudaFuncSetAttribute(enc_attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)att_smem));
  for (int l = d->layers - 1; l >= 0; --l) {
    const auto& y = L.layer[l];
    const float* xin = sv + train_saved_off(l, rows, ff, TS_XIN);
    const float* qkv = sv + train_saved_off(l, rows, ff, TS_QKV);
    const float* att = sv + train_saved_off(l, rows, ff, TS_ATT);
    const float* t1 = sv + train_saved_off(l, rows, ff, TS_T1);
    const float* x1 = sv + train_saved_off(l, rows, ff, TS_X1);
    const float* hid = sv + train_saved_off(l, rows, ff, TS_HID);
    const float* t2 = sv + train_saved_off(l, rows, ff, TS_T2);
    instance_norm_bwd_kernel<<<(unsigned)B, E, 0, st>>>(t2, gx, weights + y.n2w, N1, grads + y.n2w, grads + y.n2b);
    ELG_LAUNCH_OK();
    // t2 = x1 + hid W2^T + b2
    ELG_TRY(colsum(gx, rows, E, E, grads + y.b2, st));
    ELG_TRY(gemm_tn_acc(gx, E, hid, ff, grads + y.w2, ff, E, ff, rows, 1.f, st));
    ELG_TRY(gemm_nn(gx, E, weights + y.w2, ff, ghid, ff, rows, ff, E, 1.f, 0, st));
    relu_bwd_kernel<<<148 * 8, 256, 0, st>>>(ghid, hid, rows * ff / 4);
    ELG_LAUNCH_OK();
    // hid = relu(x1 W1^T + b1)
    ELG_TRY(colsum(ghid, rows, ff, ff, grads + y.b1, st));
    ELG_TRY(gemm_tn_acc(ghid, ff, x1, E, grads + y.w1, E, ff, E, rows, 1.f, st));
    ELG_TRY(gemm_nn(ghid, ff, weights + y.w1, E, gx, E, rows, E, ff, 1.f, 1, st));
    instance_norm_bwd_kernel<<<(unsigned)B, E, 0, st>>>(t1, gx, weights + y.n1w, N1, grads + y.n1w, grads + y.n1b);
    ELG_LAUNCH_OK();
    // t1 = xin + att Wo^T + bo
    ELG_TRY(colsum(gx, rows, E, E, grads + y.bo, st));
    ELG_TRY(gemm_tn_acc(gx, E, att, E, grads + y.wo, E, E, E, rows, 1.f, st));
    ELG_TRY(gemm_nn(gx, E, weights + y.wo, E, gatt, E, rows, E, E, 1.f, 0, st));
    enc_attention_bwd_kernel<<<(unsigned)(B * H), 128, att_smem, st>>>(qkv, att, gatt, N1, gqkv);
    ELG_LAUNCH_OK();
    // qkv = xin [Wq; Wk; Wv]^T
    ELG_TRY(gemm_tn_acc(gqkv, 3 * E, xin, E, grads + y.wq, E, 3 * E, E, rows, 1.f, st));
    ELG_TRY(gemm_nn(gqkv, 3 * E, weights + y.wq, E, gx, E, rows, E, 3 * E, 1.f, 1, st));
  }
  {
    const int per = 32;
    embed_bwd_kernel<<<(unsigned)((rows + per - 1) / per), E, 0, st>>>(d->problem, t->xy, t->demand, gx, rows, N1, per, L, grads);
    ELG_LAUNCH_OK();
  }
  return ELG_OK;
}

int elg_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, int64_t step, float lr,
                  float beta1, float beta2, float eps, float weight_decay, float grad_scale, void* stream) {
  ELG_REQUIRE(params && grads && exp_avg && exp_avg_sq && n > 0 && step >= 1, ELG_EINVAL, "bad adam arguments");
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  adam_kernel<<<148 * 4, 256, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, n, (float)(lr / bc1),
                                                         (float)sqrt(bc2), beta1, beta2, eps, weight_decay, grad_scale);
  ELG_LAUNCH_OK();
  return ELG_OK;
}

}  // extern "C"
