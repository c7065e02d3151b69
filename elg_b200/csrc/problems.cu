// Problem loading (x8 augmentation), neighbour lists, environment step on bit masks,
// per-step features (API compatibility) and tour length.
#include "common.cuh"

namespace elg {

bool rollout_is_resident(const elg_model_desc* d, int N1);

// ---- x8 augmentation + depot/node concatenation ---------------------------------------------
// reference: augment_xy_data_by_8_fold (CVRP/utils.py:69-87), load_random_problems (CVRP/CVRPEnv.py:125-150)
__global__ void load_problems_kernel(int problem, const float* __restrict__ depot, const float* __restrict__ nodes,
                                     const float* __restrict__ demand, int n, int n_nodes, int aug,
                                     float* __restrict__ xy_out, float* __restrict__ dem_out) {
  const int N1 = n_nodes + (problem == ELG_CVRP ? 1 : 0);
  const long long total = (long long)aug * n * N1;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int j = (int)(idx % N1);
    long long bi = idx / N1;
    int i = (int)(bi % n), a = (int)(bi / n);
    float x, y, dm = 0.f;
    if (problem == ELG_CVRP && j == 0) {
      x = depot[2 * i]; y = depot[2 * i + 1];
    } else {
      int jj = j - (problem == ELG_CVRP ? 1 : 0);
      x = nodes[((long long)i * n_nodes + jj) * 2];
      y = nodes[((long long)i * n_nodes + jj) * 2 + 1];
      if (demand) dm = demand[(long long)i * n_nodes + jj];
    }
    float ox, oy;
    switch (a) {
      case 0: ox = x; oy = y; break;
      case 1: ox = 1.f - x; oy = y; break;
      case 2: ox = x; oy = 1.f - y; break;
      case 3: ox = 1.f - x; oy = 1.f - y; break;
      case 4: ox = y; oy = x; break;
      case 5: ox = 1.f - y; oy = x; break;
      case 6: ox = y; oy = 1.f - x; break;
      default: ox = 1.f - y; oy = 1.f - x; break;
    }
    xy_out[idx * 2] = ox;
    xy_out[idx * 2 + 1] = oy;
    if (dem_out) dem_out[idx] = dm;
  }
}

__global__ void pairwise_dist_kernel(const float* __restrict__ xy, int B, int N1, float* __restrict__ out) {
  const long long total = (long long)B * N1 * N1;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int j = (int)(idx % N1);
    long long bi = idx / N1;
    int i = (int)(bi % N1);
    long long b = bi / N1;
    const float* p = xy + b * N1 * 2;
    out[idx] = dist2(p[2 * i] - p[2 * j], p[2 * i + 1] - p[2 * j + 1]);
  }
}

// ---- neighbour lists ------------------------------------------------------------------------
// For every node i of every aug-instance: all candidate nodes (customers 1..N for cvrp, every node
// for tsp) ordered by (distance from i, index).  The decode step takes the first k *valid* entries,
// which is what torch.topk(k, largest=False) over the masked distance row yields in the reference
// (CVRP/models.py:74,375; TSP/models.py:62,286) -- computed once instead of twice per step.
__global__ void __launch_bounds__(128) neighbour_kernel(int problem, const float* __restrict__ xy, const float* __restrict__ demand, int N1,
                                                        uint8_t* __restrict__ nbr) {
  __shared__ __align__(16) uint32_t skey[128];      // rank keys: distance bits (non-negative floats order like integers), +inf for non-candidates
  __shared__ int scnt[128];                         // candidates per "strictly smaller" count: > 1 marks a group of equal distances
  const int i = blockIdx.x % N1, b = blockIdx.x / N1;
  const float* p = xy + (size_t)b * N1 * 2;
  const int j0 = problem == ELG_CVRP ? 1 : 0;
  const float xi = p[2 * i], yi = p[2 * i + 1];
  uint8_t* row = nbr + ((size_t)b * N1 + i) * ELG_NBR_NODE_BYTES(N1);
  for (int e = threadIdx.x; e < ELG_NBR_STRIDE; e += blockDim.x) row[e] = 0;
  // pair features of get_cur_feature (CVRP/CVRPEnv.py:291-318, TSP/TSPEnv.py:135-156) for cur = i: distance and
  // polar angle to every node, so the decode step gathers them instead of recomputing sqrt / atan2 per neighbour
  float2* feat = reinterpret_cast<float2*>(row + ELG_NBR_STRIDE);
  // the same values once more in list (rank) order, with the node's demand and id: one 16-byte record per list entry
  // (rollout_tc.cu reads the records of the chosen ranks instead of chasing list byte -> pair table -> demand)
  float4* rec = reinterpret_cast<float4*>(row + ELG_NBR_STRIDE + ELG_NBR_PAIR_BYTES(N1));
  // resident instances have N1 <= 112 <= blockDim.x: thread j owns node j (distance, angle computed once, used twice)
  const int j = threadIdx.x;
  const bool cand = j >= j0 && j < N1;
  float dj = 0.f, thj = 0.f;
  if (j < N1) {
    dj = dist2(xi - p[2 * j], yi - p[2 * j + 1]);
    thj = atan2f(p[2 * j + 1] - yi, p[2 * j] - xi);
    feat[j] = make_float2(dj, thj);
  }
  const uint32_t bj = __float_as_uint(dj);
  skey[j] = cand ? bj : 0x7f800000u;
  scnt[j] = 0;
  __syncthreads();
  // rank = candidates ordered by (distance, index) before me.  The strictly smaller ones are counted on the bit patterns,
  // four keys per shared-memory load; equal distances (rare for random instances, common on integer library coordinates)
  // show up as a "strictly smaller" count shared by several candidates and are resolved by index.
  int lt = 0;
  if (cand) {
    const uint4* k4 = reinterpret_cast<const uint4*>(skey);
    const int n4 = (N1 + 3) >> 2;
#pragma unroll 4
    for (int c = 0; c < n4; ++c) {
      const uint4 kk = k4[c];
      lt += (kk.x < bj ? 1 : 0) + (kk.y < bj ? 1 : 0) + (kk.z < bj ? 1 : 0) + (kk.w < bj ? 1 : 0);
    }
    atomicAdd(&scnt[lt], 1);
  }
  __syncthreads();
  if (cand) {
    int r = lt;
    if (scnt[lt] > 1)
      for (int k = j0; k < j; ++k) r += skey[k] == bj ? 1 : 0;
    row[nbr_pos(r)] = (uint8_t)j;
    rec[r] = make_float4(dj, thj, demand ? demand[(size_t)b * N1 + j] : 0.f, __int_as_float(j));
  }
  for (int e = N1 - j0 + threadIdx.x; e < N1; e += blockDim.x) rec[e] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// ---- neighbour lists for the streaming variant (N1 > 112): uint16 ids, linear order ------------------
// One CTA per (aug-instance, node): bitonic sort of (distance bits << 32 | index) keys in shared memory.
__global__ void __launch_bounds__(512) neighbour_sort_kernel(int problem, const float* __restrict__ xy, int N1,
                                                             uint16_t* __restrict__ out, int stride, int P) {
  extern __shared__ unsigned long long skeys[];
  const int i = blockIdx.x % N1, b = blockIdx.x / N1;
  const float* p = xy + (size_t)b * N1 * 2;
  const int j0 = problem == ELG_CVRP ? 1 : 0;
  const int NL = N1 - j0;
  const float xi = p[2 * i], yi = p[2 * i + 1];
  for (int e = threadIdx.x; e < P; e += blockDim.x) {
    unsigned long long key = ~0ull;
    if (e < NL) {
      const int j = e + j0;
      const float d = dist2(xi - p[2 * j], yi - p[2 * j + 1]);
      key = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)j;
    }
    skeys[e] = key;
  }
  __syncthreads();
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int e = threadIdx.x; e < P; e += blockDim.x) {
        const int partner = e ^ j;
        if (partner > e) {
          const unsigned long long a = skeys[e], c = skeys[partner];
          const bool up = (e & k) == 0;
          if ((a > c) == up) { skeys[e] = c; skeys[partner] = a; }
        }
      }
      __syncthreads();
    }
  }
  uint16_t* row = out + ((size_t)b * N1 + i) * stride;
  for (int e = threadIdx.x; e < stride; e += blockDim.x) row[e] = e < NL ? (uint16_t)(skeys[e] & 0xffffu) : (uint16_t)0;
}

// ---- environment step on bit masks (one warp per row, W = ceil(N1/32) words per row) ----------------
// reference: CVRPEnv.step (CVRP/CVRPEnv.py:190-249), TSPEnv.step (TSP/TSPEnv.py:108-133)
__global__ void env_step_kernel(int problem, const float* __restrict__ demand, int B, int M, int N1,
                                const int32_t* __restrict__ selected, float* __restrict__ load,
                                uint32_t* __restrict__ visited, uint32_t* __restrict__ mask,
                                uint8_t* __restrict__ finished, float* __restrict__ ninf_mask,
                                int32_t* __restrict__ n_unfinished) {
  const int lane = threadIdx.x & 31;
  const long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  if (row >= (long long)B * M) return;
  const int W = (N1 + 31) >> 5;
  const int b = (int)(row / M);
  const int sel = selected[row];
  uint32_t* vis = visited + row * W;
  uint32_t* msk = mask + row * W;
  bool fin = false;
  if (problem == ELG_CVRP) {
    const float* dem = demand + (size_t)b * N1;
    const bool at_depot = sel == 0;
    float ld = load[row];
    ld = at_depot ? 1.f : ld - dem[sel];
    bool all_vis = true;
    for (int w = 0; w < W; ++w) {
      uint32_t vw = vis[w];
      if (w == (sel >> 5)) vw |= 1u << (sel & 31);
      if (w == 0) vw = at_depot ? (vw | 1u) : (vw & ~1u);
      const int j = w * 32 + lane;
      const uint32_t big = __ballot_sync(0xffffffffu, j < N1 && (__fadd_rn(ld, 1e-6f) < dem[min(j, N1 - 1)]));
      const int nb = N1 - w * 32;
      const uint32_t full = nb >= 32 ? 0xffffffffu : ((1u << nb) - 1u);
      all_vis = all_vis && ((vw & full) == full);
      __syncwarp();
      if (lane == 0) { vis[w] = vw; msk[w] = vw | big; }
    }
    fin = finished[row] || all_vis;
    __syncwarp();
    if (lane == 0) {
      if (fin) msk[0] &= ~1u;
      load[row] = ld;
      finished[row] = fin ? 1 : 0;
    }
  } else {
    if (lane == 0) {
      const uint32_t vw = vis[sel >> 5] | (1u << (sel & 31));
      vis[sel >> 5] = vw;
      msk[sel >> 5] = vw;
    }
  }
  __syncwarp();
  if (ninf_mask) {
    for (int j = lane; j < N1; j += 32)
      ninf_mask[row * N1 + j] = ((msk[j >> 5] >> (j & 31)) & 1u) ? -INFINITY : 0.f;
  }
  if (n_unfinished && lane == 0 && problem == ELG_CVRP && !fin) atomicAdd(n_unfinished, 1);
}

// ---- per-step features (API compatibility only; the decode kernel recomputes what it needs) ----
// reference: CVRPEnv.get_cur_feature (CVRP/CVRPEnv.py:291-318), TSPEnv.get_local_feature (TSP/TSPEnv.py:135-156)
__global__ void cur_feature_kernel(const float* __restrict__ xy, const float* __restrict__ demand,
                                   const float* __restrict__ load, const int32_t* __restrict__ cur, int B, int M,
                                   int N1, float* __restrict__ cur_dist, float* __restrict__ cur_theta,
                                   float* __restrict__ rel_xy, float* __restrict__ norm_demand) {
  const long long total = (long long)B * M * N1;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int j = (int)(idx % N1);
    long long row = idx / N1;
    long long b = row / M;
    const float* p = xy + b * N1 * 2;
    int c = cur[row];
    float rx = p[2 * j] - p[2 * c], ry = p[2 * j + 1] - p[2 * c + 1];
    cur_dist[idx] = dist2(p[2 * c] - p[2 * j], p[2 * c + 1] - p[2 * j + 1]);
    cur_theta[idx] = atan2f(ry, rx);
    rel_xy[idx * 2] = rx;
    rel_xy[idx * 2 + 1] = ry;
    if (norm_demand) norm_demand[idx] = demand[b * N1 + j] / load[row];
  }
}

// ---- closed-tour length -----------------------------------------------------------------------
// reference: _get_reward / compute_unscaled_reward (CVRP/CVRPEnv.py:251-288),
//            _get_travel_distance / compute_unscaled_distance (TSP/TSPEnv.py:158-184)
__global__ void tour_length_kernel(const float* __restrict__ xy, int Bxy, const int64_t* __restrict__ tours, int B,
                                   int M, int T, int N1, int round_edges, float* __restrict__ out) {
  const long long row = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (row >= (long long)B * M) return;
  const long long b = row / M;
  const float* p = xy + (Bxy == 1 ? 0 : b) * (long long)N1 * 2;
  const int64_t* t = tours + row * T;
  int first = (int)t[0], prev = first;
  float sum = 0.f;
  for (int s = 1; s <= T; ++s) {
    int nxt = s < T ? (int)t[s] : first;
    float d = seglen(p[2 * prev] - p[2 * nxt], p[2 * prev + 1] - p[2 * nxt + 1]);
    if (round_edges) d = rintf(d);
    sum += d;
    prev = nxt;
  }
  out[row] = sum;
}

static inline int grid_for(long long total, int block) {
  long long g = (total + block - 1) / block;
  return (int)(g > 148LL * 32 ? 148 * 32 : (g < 1 ? 1 : g));
}

}  // namespace elg

using namespace elg;

extern "C" {

int elg_load_problems(int problem, const float* depot_xy, const float* node_xy, const float* node_demand, int n,
                      int n_nodes, int aug, float* xy_out, float* demand_out, void* stream) {
  ELG_REQUIRE(problem == ELG_TSP || problem == ELG_CVRP, ELG_EINVAL, "unknown problem %d", problem);
  ELG_REQUIRE(aug == 1 || aug == 8, ELG_EUNSUPPORTED, "aug_factor must be 1 or 8 (got %d)", aug);
  ELG_REQUIRE(n > 0 && n_nodes > 0 && node_xy && xy_out, ELG_EINVAL, "bad sizes/pointers");
  ELG_REQUIRE(problem == ELG_TSP || (depot_xy && node_demand && demand_out), ELG_EINVAL, "cvrp needs depot/demand");
  const int N1 = n_nodes + (problem == ELG_CVRP ? 1 : 0);
  long long total = (long long)aug * n * N1;
  load_problems_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      problem, depot_xy, node_xy, problem == ELG_CVRP ? node_demand : nullptr, n, n_nodes, aug, xy_out,
      problem == ELG_CVRP ? demand_out : nullptr);
  ELG_LAUNCH_OK();
  return ELG_OK;
}

int elg_rollout_resident(const elg_model_desc* d, int N1) {
  if (check_desc(d) || N1 <= 1) return -1;
  return rollout_is_resident(d, N1) ? 1 : 0;
}

size_t elg_e_bytes(const elg_model_desc* d, int B, int N1) {
  if (check_desc(d) || B <= 0 || N1 <= 1) return 0;
  if (rollout_is_resident(d, N1)) return (size_t)B * 3 * ((N1 + 15) & ~15) * 512;     // E' | K' | V^T, each fp16 hi + lo, rows padded to 16
  return (size_t)B * N1 * 128 * sizeof(float);
}

size_t elg_et_bytes(const elg_model_desc* d, int B, int N1) {
  if (check_desc(d) || B <= 0 || N1 <= 1 || rollout_is_resident(d, N1)) return 0;
  return (size_t)B * ((N1 + ELG_TILE_NODES - 1) / ELG_TILE_NODES) * ELG_TILE_BYTES;
}

size_t elg_nbr_bytes(const elg_model_desc* d, int B, int N1) {
  if (check_desc(d) || B <= 0 || N1 <= 1) return 0;
  if (rollout_is_resident(d, N1)) return (size_t)B * N1 * ELG_NBR_NODE_BYTES(N1);
  const int NL = N1 - (d->problem == ELG_CVRP ? 1 : 0);
  return (size_t)B * N1 * ELG_NBR16_STRIDE(NL) * sizeof(uint16_t);
}

int elg_pairwise_dist(const float* xy, int B, int N1, float* dist_out, void* stream) {
  ELG_REQUIRE(xy && dist_out && B > 0 && N1 > 0, ELG_EINVAL, "bad sizes/pointers");
  pairwise_dist_kernel<<<grid_for((long long)B * N1 * N1, 256), 256, 0, (cudaStream_t)stream>>>(xy, B, N1, dist_out);
  ELG_LAUNCH_OK();
  return ELG_OK;
}

int elg_env_step(int problem, const float* demand, int B, int M, int N1, const int32_t* selected, float* load,
                 uint32_t* visited_bits, uint32_t* mask_bits, uint8_t* finished, float* ninf_mask,
                 int32_t* n_unfinished, void* stream) {
  ELG_REQUIRE(problem == ELG_TSP || problem == ELG_CVRP, ELG_EINVAL, "unknown problem %d", problem);
  ELG_REQUIRE(N1 > 0, ELG_EINVAL, "bad node count %d", N1);
  ELG_REQUIRE(selected && visited_bits && mask_bits, ELG_EINVAL, "NULL state pointer");
  ELG_REQUIRE(problem == ELG_TSP || (demand && load && finished), ELG_EINVAL, "cvrp needs demand/load/finished");
  long long threads = (long long)B * M * 32;
  env_step_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      problem, demand, B, M, N1, selected, load, visited_bits, mask_bits, finished, ninf_mask, n_unfinished);
  ELG_LAUNCH_OK();
  return ELG_OK;
}

int elg_cur_feature(const float* xy, const float* demand, const float* load, const int32_t* cur, int B, int M,
                    int N1, float* cur_dist, float* cur_theta, float* rel_xy, float* norm_demand, void* stream) {
  ELG_REQUIRE(xy && cur && cur_dist && cur_theta && rel_xy, ELG_EINVAL, "NULL pointer");
  ELG_REQUIRE(!norm_demand || (demand && load), ELG_EINVAL, "norm_demand needs demand and load");
  cur_feature_kernel<<<grid_for((long long)B * M * N1, 256), 256, 0, (cudaStream_t)stream>>>(
      xy, demand, load, cur, B, M, N1, cur_dist, cur_theta, rel_xy, norm_demand);
  ELG_LAUNCH_OK();
  return ELG_OK;
}

int elg_tour_length(const float* xy, int Bxy, const int64_t* tours, int B, int M, int T, int N1, int round_edges,
                    float* out, void* stream) {
  ELG_REQUIRE(xy && tours && out && T > 0, ELG_EINVAL, "bad sizes/pointers");
  ELG_REQUIRE(Bxy == 1 || Bxy == B, ELG_EINVAL, "xy batch must be 1 or B");
  long long rows = (long long)B * M;
  tour_length_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, (cudaStream_t)stream>>>(xy, Bxy, tours, B, M, T, N1,
                                                                                        round_edges, out);
  ELG_LAUNCH_OK();
  return ELG_OK;
}

}  // extern "C"

// neighbour lists are built by elg_encode (encoder.cu) through this launcher:
// resident variant (N1 <= 112): uint8 ids, 8-way interleaved, ELG_NBR_STRIDE bytes per node;
// streaming variant: uint16 ids in rank order, ELG_NBR16_STRIDE(NL) entries per node.
namespace elg {
int launch_neighbours(const elg_model_desc* d, const float* xy, const float* demand, int B, int N1, void* nbr, cudaStream_t stream) {
  const int problem = d->problem;
  if (rollout_is_resident(d, N1)) {
    neighbour_kernel<<<(unsigned)B * N1, 128, 0, stream>>>(problem, xy, problem == ELG_CVRP ? demand : nullptr, N1, reinterpret_cast<uint8_t*>(nbr));
    ELG_LAUNCH_OK();
    return ELG_OK;
  }
  const int NL = N1 - (problem == ELG_CVRP ? 1 : 0);
  ELG_REQUIRE(N1 <= 8192, ELG_EUNSUPPORTED, "neighbour lists support up to 8192 nodes (got %d)", N1);
  int P = 64;
  while (P < NL) P <<= 1;
  const int stride = ELG_NBR16_STRIDE(NL);
  const size_t smem = (size_t)P * sizeof(unsigned long long);
  ELG_CUDA_OK(cudaFuncSetAttribute(neighbour_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  neighbour_sort_kernel<<<(unsigned)B * N1, 512, smem, stream>>>(problem, xy, N1, reinterpret_cast<uint16_t*>(nbr), stride, P);
  ELG_LAUNCH_OK();
  return ELG_OK;
}
}  // namespace elg
