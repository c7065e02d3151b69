// Training-data generators on the device (Philox4x32-10 streams): uniform, cluster and mixed instance distributions.
//
// reference: generate_vrp_data CVRP/generate_data.py:9-92, generate_tsp_data TSP/generate_data.py:9-57.
// The reference draws from torch / numpy generators on the host, so instances cannot be bit-identical; what is kept is
// the distribution: uniform points in [0,1)^2; `cluster`: n_cluster centres lower + (upper - lower) U, the points
// split into n_cluster consecutive groups of floor(P / n_cluster) (the last one takes the remainder), each N(centre, std)
// clamped to [0,1], and for cvrp the depot is one uniformly chosen point of the P = N + 1 (the rest keep their order);
// `mixed`: N uniform points of which a uniformly chosen half (without replacement) is redrawn around n_cluster_mix
// centres the same way, cvrp depot uniform; demands uniform integers 1..9 divided by the capacity of the size.
#include "rollout_common.cuh"

namespace elg {

__device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }     // [0, 1)

// one CTA per instance
__global__ void __launch_bounds__(128) generate_kernel(int problem, int kind, int N, int n_cluster, float lower, float upper,
                                                       float stdv, float capacity, unsigned long long seed,
                                                       float* __restrict__ depot, float* __restrict__ node,
                                                       float* __restrict__ demand) {
  extern __shared__ float gs[];
  const int b = blockIdx.x, tid = threadIdx.x;
  const bool cvrp = problem == ELG_CVRP;
  const int P = (cvrp && kind == 1) ? N + 1 : N;       // points drawn together
  float* pts = gs;                                       // [P][2]
  int* perm = reinterpret_cast<int*>(gs + 2 * (N + 1));  // [N]
  __shared__ float cx[16], cy[16];
  __shared__ int depot_idx;
  const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  auto rnd = [&](uint32_t stream, uint32_t i) { return philox4x32(make_uint4(i, (uint32_t)b, stream, 0x9e3779b9u), key); };
  if (tid < n_cluster) {
    const uint4 r = rnd(1, tid);
    cx[tid] = lower + (upper - lower) * u01(r.x);
    cy[tid] = lower + (upper - lower) * u01(r.y);
  }
  if (tid == 0) depot_idx = (int)(u01(rnd(2, 0).x) * (float)(N + 1));
  for (int i = tid; i < N; i += blockDim.x) perm[i] = i;
  __syncthreads();
  if (kind == 2 && tid == 0) {
    // uniformly chosen half without replacement: partial Fisher-Yates
    const int half = N / 2;
    for (int i = 0; i < half; ++i) {
      const int j = i + (int)(u01(rnd(3, i).x) * (float)(N - i));
      const int t = perm[i]; perm[i] = perm[min(j, N - 1)]; perm[min(j, N - 1)] = t;
    }
  }
  for (int i = tid; i < P; i += blockDim.x) {
    const uint4 r = rnd(4, i);
    float x = u01(r.x), y = u01(r.y);
    if (kind == 1) {
      const int g = P / n_cluster;
      const int c = min(i / max(g, 1), n_cluster - 1);
      const float rad = sqrtf(-2.f * logf(1.f - u01(r.z))), ang = 6.283185307179586f * u01(r.w);
      x = fminf(fmaxf(cx[c] + stdv * rad * cosf(ang), 0.f), 1.f);
      y = fminf(fmaxf(cy[c] + stdv * rad * sinf(ang), 0.f), 1.f);
    }
    pts[2 * i] = x; pts[2 * i + 1] = y;
  }
  __syncthreads();
  if (kind == 2) {
    const int half = N / 2, g = N / n_cluster / 2;
    for (int i = tid; i < half; i += blockDim.x) {
      const int c = min(i / max(g, 1), n_cluster - 1);
      const uint4 r = rnd(5, i);
      const float rad = sqrtf(-2.f * logf(1.f - u01(r.z))), ang = 6.283185307179586f * u01(r.w);
      const int idx = perm[i];
      pts[2 * idx] = fminf(fmaxf(cx[c] + stdv * rad * cosf(ang), 0.f), 1.f);
      pts[2 * idx + 1] = fminf(fmaxf(cy[c] + stdv * rad * sinf(ang), 0.f), 1.f);
    }
    __syncthreads();
  }
  float* out = node + (size_t)b * N * 2;
  if (cvrp && kind == 1) {
    for (int i = tid; i < P; i += blockDim.x) {
      if (i == depot_idx) { depot[2 * b] = pts[2 * i]; depot[2 * b + 1] = pts[2 * i + 1]; }
      else { const int o = i < depot_idx ? i : i - 1; out[2 * o] = pts[2 * i]; out[2 * o + 1] = pts[2 * i + 1]; }
    }
  } else {
    for (int i = tid; i < N; i += blockDim.x) { out[2 * i] = pts[2 * i]; out[2 * i + 1] = pts[2 * i + 1]; }
    if (cvrp && tid == 0) { const uint4 r = rnd(6, 0); depot[2 * b] = u01(r.x); depot[2 * b + 1] = u01(r.y); }
  }
  if (cvrp)
    for (int i = tid; i < N; i += blockDim.x)
      demand[(size_t)b * N + i] = (float)(1 + min((int)(u01(rnd(7, i).x) * 9.f), 8)) / capacity;
}

}  // namespace elg

using namespace elg;

extern "C" int elg_generate_problems(int problem, int kind, int n, int n_nodes, int n_cluster, float lower, float upper,
                                     float std, float capacity, uint64_t seed, float* depot_xy, float* node_xy,
                                     float* node_demand, void* stream) {
  ELG_REQUIRE(problem == ELG_TSP || problem == ELG_CVRP, ELG_EINVAL, "unknown problem %d", problem);
  ELG_REQUIRE(kind >= 0 && kind <= 2, ELG_EINVAL, "data kind must be 0 (uniform), 1 (cluster) or 2 (mixed)");
  ELG_REQUIRE(n > 0 && n_nodes > 1 && node_xy, ELG_EINVAL, "bad sizes / NULL output");
  ELG_REQUIRE(problem == ELG_TSP || (depot_xy && node_demand && capacity > 0.f), ELG_EINVAL, "cvrp needs depot, demand and a capacity");
  ELG_REQUIRE(kind == 0 || (n_cluster >= 1 && n_cluster <= 16), ELG_EINVAL, "n_cluster must be in [1, 16]");
  const size_t smem = (size_t)(2 * (n_nodes + 1) + n_nodes) * sizeof(float);
  ELG_REQUIRE(smem <= 200 * 1024, ELG_EUNSUPPORTED, "problem_size too large for the generator");
  ELG_CUDA_OK(cudaFuncSetAttribute(generate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  generate_kernel<<<(unsigned)n, 128, smem, (cudaStream_t)stream>>>(problem, kind, n_nodes, kind == 0 ? 1 : n_cluster, lower, upper, std,
                                                                  capacity, (unsigned long long)seed, depot_xy, node_xy, node_demand);
  ELG_LAUNCH_OK();
  return ELG_OK;
}
