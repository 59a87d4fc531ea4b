// Exact k-nearest-neighbour search on a uniform grid hash, and the label vote on top of it:
//   * match_labels_to_vtx (ovo/utils/eval_utils.py:13-44): scipy KDTree.query(mesh_vtx, k=5) + torch.mode over the 5 labels;
//   * same_instance (ovo/utils/instance_utils.py:5-24): Open3D compute_point_cloud_distance = nearest-neighbour distance.
// Both are brute CPU tree searches in the reference.  Here: counting sort of the points by cell (HBM-bound, N*(12+16+8) B),
// then one thread per query walks the shells of cells around its own cell until the k-th distance is provably final.
// Distances are evaluated in double like the KD-tree does (float32 coordinates convert exactly), ties break on the point index.
#include <algorithm>
#include <cmath>
#include <vector>

#include "common.cuh"

namespace ovo {

constexpr int kKnnMaxK = 8;
constexpr int kKnnMaxRing = 10;     // shells walked on the grid before a query falls back to the exhaustive scan

struct KnnGrid {
  float ox, oy, oz, cell, inv_cell;
  int nx, ny, nz;
};

__device__ __forceinline__ unsigned f2ord(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
inline float ord2f_host(unsigned u) {
  u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

// mm[0..2] = min xyz, mm[3..5] = max xyz as order-preserving unsigned keys
__global__ void knn_bounds_kernel(const float* __restrict__ xyz, long long N, unsigned* __restrict__ mm) {
  unsigned lo[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, hi[3] = {0u, 0u, 0u};
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < N; i += static_cast<long long>(gridDim.x) * blockDim.x)
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const unsigned k = f2ord(xyz[3 * i + a]);
      lo[a] = min(lo[a], k); hi[a] = max(hi[a], k);
    }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    for (int o = 16; o > 0; o >>= 1) {
      lo[a] = min(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
      hi[a] = max(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
    }
    if ((threadIdx.x & 31) == 0) { atomicMin(mm + a, lo[a]); atomicMax(mm + 3 + a, hi[a]); }
  }
}

__device__ __forceinline__ int axis_cell(float v, float o, float inv, int n) {
  const int c = static_cast<int>(floorf((v - o) * inv));
  return min(max(c, 0), n - 1);
}
__device__ __forceinline__ int cell_index(const KnnGrid& g, float x, float y, float z) {
  return (axis_cell(z, g.oz, g.inv_cell, g.nz) * g.ny + axis_cell(y, g.oy, g.inv_cell, g.ny)) * g.nx + axis_cell(x, g.ox, g.inv_cell, g.nx);
}

__global__ void knn_count_kernel(const float* __restrict__ xyz, long long N, KnnGrid g, int* __restrict__ cell_of_pt, int* __restrict__ counts) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= N) return;
  const int c = cell_index(g, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
  cell_of_pt[i] = c;
  atomicAdd(counts + c, 1);
}

// sum of count^2 over the cells = (points) x (mean occupancy of the cell a point lives in): the cost a query near the points pays
__global__ void knn_occupancy_kernel(const int* __restrict__ counts, int n_cells, unsigned long long* __restrict__ sum2) {
  unsigned long long acc = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_cells; i += gridDim.x * blockDim.x) {
    const unsigned long long c = static_cast<unsigned long long>(counts[i]);
    acc += c * c;
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0 && acc) atomicAdd(sum2, acc);
}

// exclusive scan, three passes: per-block (1024 items) scan + block totals, scan of the totals by one block, offset add
__global__ void __launch_bounds__(256) scan_blocks_kernel(int* __restrict__ data, int n, int* __restrict__ totals) {
  __shared__ int s_warp[8];
  const int base = blockIdx.x * 1024 + threadIdx.x * 4;
  int v[4], sum = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) { v[i] = base + i < n ? data[base + i] : 0; sum += v[i]; }
  int inc = sum;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = lane < 8 ? s_warp[lane] : 0;
    for (int o = 1; o < 8; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
    if (lane < 8) s_warp[lane] = w;
  }
  __syncthreads();
  int run = inc - sum + (warp > 0 ? s_warp[warp - 1] : 0);
#pragma unroll
  for (int i = 0; i < 4; ++i) { if (base + i < n) data[base + i] = run; run += v[i]; }
  if (threadIdx.x == 255) totals[blockIdx.x] = s_warp[7];
}
__global__ void __launch_bounds__(1024) scan_totals_kernel(int* __restrict__ totals, int nb) {
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int b0 = 0; b0 < nb; b0 += 1024) {
    const int i = b0 + threadIdx.x;
    const int v = i < nb ? totals[i] : 0;
    int inc = v;
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      int w = s_warp[lane];
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
      s_warp[lane] = w;
    }
    __syncthreads();
    const int carry = s_carry;
    if (i < nb) totals[i] = carry + inc - v + (warp > 0 ? s_warp[warp - 1] : 0);
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = carry + s_warp[31];
    __syncthreads();
  }
}
__global__ void scan_add_kernel(int* __restrict__ data, int n, const int* __restrict__ totals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) data[i] += totals[i >> 10];
}

// sorted[pos] = (x, y, z, original index); the order inside a cell is arbitrary (the search breaks ties on the index)
__global__ void knn_scatter_kernel(const float* __restrict__ xyz, long long N, const int* __restrict__ cell_of_pt,
                                   const int* __restrict__ starts, int* __restrict__ cursor, float4* __restrict__ sorted) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= N) return;
  const int c = cell_of_pt[i];
  const int pos = starts[c] + atomicAdd(cursor + c, 1);
  sorted[pos] = make_float4(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], __int_as_float(static_cast<int>(i)));
}

// K best (distance^2, index) pairs in registers, ascending; one fully unrolled bubble pass per insertion
template <int K>
struct TopK {
  double d[K];
  int id[K];
  __device__ void init() {
#pragma unroll
    for (int i = 0; i < K; ++i) { d[i] = INFINITY; id[i] = 0x7fffffff; }
  }
  __device__ __forceinline__ void push(double dd, int idx) {
    if (dd > d[K - 1] || (dd == d[K - 1] && idx >= id[K - 1])) return;
    d[K - 1] = dd; id[K - 1] = idx;
#pragma unroll
    for (int j = K - 1; j > 0; --j)
      if (d[j] < d[j - 1] || (d[j] == d[j - 1] && id[j] < id[j - 1])) {
        const double td = d[j]; d[j] = d[j - 1]; d[j - 1] = td;
        const int ti = id[j]; id[j] = id[j - 1]; id[j - 1] = ti;
      }
  }
  __device__ __forceinline__ double kth(int k) const {
    double v = d[0];
#pragma unroll
    for (int i = 1; i < K; ++i) if (i == k - 1) v = d[i];
    return v;
  }
};

__device__ __forceinline__ double dist2(const float4& p, float qx, float qy, float qz) {
  const double dx = static_cast<double>(p.x) - qx, dy = static_cast<double>(p.y) - qy, dz = static_cast<double>(p.z) - qz;
  return dx * dx + dy * dy + dz * dz;
}

// Points [p0, p1) of the sorted array against one query, by one WARP: each lane loads one point (coalesced float4) and
// evaluates its float distance; the few candidates that are not rejected by the current k-th distance are broadcast one by one
// and pushed by every lane into its own (identical) copy of the top-k list, in double.  The float distance only rejects, and with
// a margin above its rounding error, so the result is the double-precision one.
template <int K>
__device__ __forceinline__ void scan_run(const float4* __restrict__ sorted, int p0, int p1, float qx, float qy, float qz, int k,
                                         TopK<K>& top, float& kth_f) {
  const int lane = threadIdx.x & 31;
  for (int base = p0; base < p1; base += 32) {
    const int p = base + lane;
    float4 pt = make_float4(0.f, 0.f, 0.f, 0.f);
    bool pass = false;
    if (p < p1) {
      pt = __ldg(sorted + p);
      const float fx = pt.x - qx, fy = pt.y - qy, fz = pt.z - qz;
      pass = fx * fx + fy * fy + fz * fz <= kth_f;
    }
    unsigned m = __ballot_sync(0xffffffffu, pass);
    while (m) {
      const int src = __ffs(m) - 1;
      m &= m - 1;
      float4 c;
      c.x = __shfl_sync(0xffffffffu, pt.x, src); c.y = __shfl_sync(0xffffffffu, pt.y, src);
      c.z = __shfl_sync(0xffffffffu, pt.z, src); c.w = __shfl_sync(0xffffffffu, pt.w, src);
      top.push(dist2(c, qx, qy, qz), __float_as_int(c.w));
    }
    kth_f = __double2float_ru(top.kth(k) * 1.000004);
  }
}

// One warp per query.  Shell r = the cells at Chebyshev distance r from the query's cell.  After shell r every point inside the
// cube of (2r+1)^3 cells has been seen; the k-th distance is final once it does not exceed the distance from the query to the
// nearest face of that cube behind which unseen cells remain.
template <int K>
__global__ void __launch_bounds__(128)
    knn_query_kernel(const float4* __restrict__ sorted, const int* __restrict__ starts, KnnGrid g, const float* __restrict__ queries,
                     long long Q, int k, int32_t* __restrict__ idx_out, double* __restrict__ dist_out, int* __restrict__ overflow,
                     int* __restrict__ n_overflow) {
  const long long q = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;   // uniform over the warp
  if (q >= Q) return;
  const float qx = __ldg(queries + 3 * q), qy = __ldg(queries + 3 * q + 1), qz = __ldg(queries + 3 * q + 2);
  const int cx = axis_cell(qx, g.ox, g.inv_cell, g.nx), cy = axis_cell(qy, g.oy, g.inv_cell, g.ny), cz = axis_cell(qz, g.oz, g.inv_cell, g.nz);
  TopK<K> top;
  top.init();
  float kth_f = INFINITY;
  bool done = false;
  for (int r = 0; r <= kKnnMaxRing && !done; ++r) {
    // the shell's runs of cells: a row on a face of the cube contributes one contiguous run of cells, any other row its two end
    // cells.  Each lane fetches the bounds of one run, then the warp scans the non-empty ones (most are empty in sparse regions).
    const int w = 2 * r + 1, nslots = 2 * w * w, lane = threadIdx.x & 31;
    for (int s0 = 0; s0 < nslots; s0 += 32) {
      const int slot = s0 + lane;
      int p0 = 0, p1 = 0;
      if (slot < nslots) {
        const int idx = slot >> 1, side = slot & 1;
        const int dz = idx / w - r, dy = idx % w - r;
        const int z = cz + dz, y = cy + dy;
        if (z >= 0 && z < g.nz && y >= 0 && y < g.ny) {
          const size_t row = (static_cast<size_t>(z) * g.ny + y) * g.nx;
          if (abs(dz) == r || abs(dy) == r) {
            if (side == 0) { p0 = __ldg(starts + row + max(cx - r, 0)); p1 = __ldg(starts + row + min(cx + r, g.nx - 1) + 1); }
          } else {
            const int x = side == 0 ? cx - r : cx + r;
            if (x >= 0 && x < g.nx) { p0 = __ldg(starts + row + x); p1 = __ldg(starts + row + x + 1); }
          }
        }
      }
      unsigned m = __ballot_sync(0xffffffffu, p1 > p0);
      while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        scan_run(sorted, __shfl_sync(0xffffffffu, p0, src), __shfl_sync(0xffffffffu, p1, src), qx, qy, qz, k, top, kth_f);
      }
    }
    // margin to the unseen part of the grid
    double margin = INFINITY;
    const int c[3] = {cx, cy, cz}, n[3] = {g.nx, g.ny, g.nz};
    const float qv[3] = {qx, qy, qz}, o[3] = {g.ox, g.oy, g.oz};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      if (c[a] - r > 0) margin = fmin(margin, static_cast<double>(qv[a]) - (static_cast<double>(o[a]) + static_cast<double>(c[a] - r) * g.cell));
      if (c[a] + r < n[a] - 1) margin = fmin(margin, (static_cast<double>(o[a]) + static_cast<double>(c[a] + r + 1) * g.cell) - static_cast<double>(qv[a]));
    }
    // a point is binned with float arithmetic: it may sit up to ~1e-3 cells on the other side of a cell plane
    margin -= 1e-3 * static_cast<double>(g.cell);
    if (margin == INFINITY) done = true;                                   // the whole grid has been walked
    else if (margin > 0.0 && top.kth(k) <= margin * margin) done = true;
  }
  if ((threadIdx.x & 31) != 0) return;
  if (!done) {
    overflow[atomicAdd(n_overflow, 1)] = static_cast<int>(q);
    return;
  }
#pragma unroll
  for (int i = 0; i < K; ++i)
    if (i < k) {
      idx_out[q * k + i] = top.id[i];
      if (dist_out) dist_out[q * k + i] = sqrt(top.d[i]);
    }
}

// exhaustive scan for the queries the grid walk gave up on (far from every point): one block per query
__global__ void __launch_bounds__(256)
    knn_brute_kernel(const float4* __restrict__ sorted, long long N, const float* __restrict__ queries, const int* __restrict__ overflow,
                     int k, int32_t* __restrict__ idx_out, double* __restrict__ dist_out) {
  __shared__ double s_d[256 * kKnnMaxK];
  __shared__ int s_i[256 * kKnnMaxK];
  const long long q = overflow[blockIdx.x];
  const float qx = queries[3 * q], qy = queries[3 * q + 1], qz = queries[3 * q + 2];
  TopK<kKnnMaxK> top;
  top.init();
  float kth_f = INFINITY;
  for (long long p = threadIdx.x; p < N; p += blockDim.x) {
    const float4 pt = sorted[p];
    const float fx = pt.x - qx, fy = pt.y - qy, fz = pt.z - qz;
    if (fx * fx + fy * fy + fz * fz > kth_f) continue;
    top.push(dist2(pt, qx, qy, qz), __float_as_int(pt.w));
    kth_f = __double2float_ru(top.kth(k) * 1.000004);
  }
  constexpr int K = kKnnMaxK;
#pragma unroll
  for (int i = 0; i < K; ++i) { s_d[threadIdx.x * K + i] = top.d[i]; s_i[threadIdx.x * K + i] = top.id[i]; }
  __syncthreads();
  // tree merge of the per-thread lists
  for (int stride = 128; stride > 0; stride >>= 1) {
    if (threadIdx.x < stride) {
#pragma unroll
      for (int i = 0; i < K; ++i) top.push(s_d[(threadIdx.x + stride) * K + i], s_i[(threadIdx.x + stride) * K + i]);
#pragma unroll
      for (int i = 0; i < K; ++i) { s_d[threadIdx.x * K + i] = top.d[i]; s_i[threadIdx.x * K + i] = top.id[i]; }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < K; ++i)
      if (i < k) {
        idx_out[q * k + i] = top.id[i];
        if (dist_out) dist_out[q * k + i] = sqrt(top.d[i]);
      }
  }
}

// torch.mode over the k labels of each query (eval_utils.py:31-32): the most frequent value, the smallest one on ties
__global__ void knn_mode_kernel(const int32_t* __restrict__ labels, const int32_t* __restrict__ idx, long long Q, int k,
                                int32_t* __restrict__ out) {
  const long long q = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (q >= Q) return;
  int32_t v[kKnnMaxK];
  for (int i = 0; i < k; ++i) v[i] = labels[idx[q * k + i]];
  int best = v[0], best_n = 0;
  for (int i = 0; i < k; ++i) {
    int n = 0;
    for (int j = 0; j < k; ++j) n += v[j] == v[i];
    if (n > best_n || (n == best_n && v[i] < best)) { best = v[i]; best_n = n; }
  }
  out[q] = best;
}

}  // namespace ovo

static thread_local float g_knn_cell = 0.f;
static thread_local int g_knn_cells = 0, g_knn_overflow = 0;

extern "C" {

void ovo_knn_stats(float* cell_size, int* n_cells, int* n_fallback) {
  if (cell_size) *cell_size = g_knn_cell;
  if (n_cells) *n_cells = g_knn_cells;
  if (n_fallback) *n_fallback = g_knn_overflow;
}

int ovo_knn(const float* points_dev, int64_t N, const float* queries_dev, int64_t Q, int k, float cell_size,
            int32_t* idx_out_dev, double* dist_out_dev, void* stream_) {
  using namespace ovo;
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  OVO_REQUIRE(points_dev && queries_dev && idx_out_dev, "ovo_knn: null argument");
  OVO_REQUIRE(k >= 1 && k <= kKnnMaxK, "ovo_knn: k=%d outside [1, %d]", k, kKnnMaxK);
  OVO_REQUIRE(N >= k && N < (1LL << 31) && Q >= 0 && Q < (1LL << 31), "ovo_knn: need k <= N < 2^31 points (N=%lld, Q=%lld)", (long long)N, (long long)Q);
  if (Q == 0) return OVO_OK;
  keep_default_mempool_cached();
  AsyncTemps tmp(s);
  unsigned* mm = nullptr;
  OVO_CUDA(tmp.alloc(&mm, 6 * sizeof(unsigned)));
  const unsigned init[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
  unsigned got[6];
  OVO_CUDA(cudaMemcpyAsync(mm, init, sizeof(init), cudaMemcpyHostToDevice, s));
  knn_bounds_kernel<<<std::min<long long>(ceil_div(N, 256), 1184), 256, 0, s>>>(points_dev, N, mm);
  OVO_CHECK_LAUNCH();
  OVO_CUDA(cudaMemcpyAsync(got, mm, sizeof(got), cudaMemcpyDeviceToHost, s));
  OVO_CUDA(cudaStreamSynchronize(s));
  float lo[3], hi[3];
  for (int a = 0; a < 3; ++a) { lo[a] = ord2f_host(got[a]); hi[a] = ord2f_host(got[3 + a]); }
  OVO_REQUIRE(std::isfinite(lo[0]) && std::isfinite(lo[1]) && std::isfinite(lo[2]) && std::isfinite(hi[0]) && std::isfinite(hi[1]) && std::isfinite(hi[2]),
              "ovo_knn: non-finite point coordinates");
  const double ex = std::max(hi[0] - lo[0], 1e-6f), ey = std::max(hi[1] - lo[1], 1e-6f), ez = std::max(hi[2] - lo[2], 1e-6f);
  double cell = cell_size;
  if (!(cell > 0)) {
    // about 8 points per occupied cell, for points spread in the volume or (the usual case: depth maps) on surfaces
    const double vol = std::cbrt(8.0 * ex * ey * ez / static_cast<double>(N));
    const double surf = std::sqrt(8.0 * 2.0 * (ex * ey + ey * ez + ex * ez) / static_cast<double>(N));
    cell = std::max(vol, std::min(surf, std::max({ex, ey, ez})));
  }
  constexpr double kMaxCells = static_cast<double>(1 << 25);
  auto fit = [&](double c, KnnGrid* g) {
    const double nx = std::floor(ex / c) + 1, ny = std::floor(ey / c) + 1, nz = std::floor(ez / c) + 1;
    if (!(nx * ny * nz <= kMaxCells && nx <= 4096 && ny <= 4096 && nz <= 4096)) return false;
    g->nx = static_cast<int>(nx); g->ny = static_cast<int>(ny); g->nz = static_cast<int>(nz);
    g->ox = lo[0]; g->oy = lo[1]; g->oz = lo[2];
    g->cell = static_cast<float>(c); g->inv_cell = 1.0f / g->cell;
    return true;
  };
  KnnGrid g;
  while (!fit(cell, &g)) cell *= 1.26;
  int *cell_of_pt = nullptr, *starts = nullptr, *cursor = nullptr, *totals = nullptr, *overflow = nullptr, *n_overflow = nullptr;
  unsigned long long* sum2 = nullptr;
  float4* sorted = nullptr;
  OVO_CUDA(tmp.alloc(&cell_of_pt, N * sizeof(int)));
  OVO_CUDA(tmp.alloc(&sum2, sizeof(unsigned long long)));
  int n_cells = 0;
  // Count the points per cell; when the density is uneven (a depth-map surface inside a sparse volume) the cells that hold the
  // points are crowded: halve the cell while a point shares its cell with more than ~16 others on average (auto cell size only).
  for (int pass = 0;; ++pass) {
    n_cells = g.nx * g.ny * g.nz;
    OVO_CUDA(tmp.alloc(&starts, (static_cast<size_t>(n_cells) + 1) * sizeof(int)));
    OVO_CUDA(cudaMemsetAsync(starts, 0, (static_cast<size_t>(n_cells) + 1) * sizeof(int), s));
    knn_count_kernel<<<ceil_div(N, 256), 256, 0, s>>>(points_dev, N, g, cell_of_pt, starts);
    OVO_CHECK_LAUNCH();
    KnnGrid finer;
    if (cell_size > 0 || pass >= 5 || !fit(cell * 0.5, &finer)) break;
    unsigned long long h = 0;
    OVO_CUDA(cudaMemsetAsync(sum2, 0, sizeof(unsigned long long), s));
    knn_occupancy_kernel<<<std::min(ceil_div(n_cells, 256), 1184), 256, 0, s>>>(starts, n_cells, sum2);
    OVO_CHECK_LAUNCH();
    OVO_CUDA(cudaMemcpyAsync(&h, sum2, sizeof(h), cudaMemcpyDeviceToHost, s));
    OVO_CUDA(cudaStreamSynchronize(s));
    if (static_cast<double>(h) / static_cast<double>(N) <= 16.0) break;
    tmp.release(starts);
    cell *= 0.5;
    g = finer;
  }
  const int nb = ceil_div(n_cells + 1, 1024);
  OVO_CUDA(tmp.alloc(&cursor, static_cast<size_t>(n_cells) * sizeof(int)));
  OVO_CUDA(tmp.alloc(&totals, static_cast<size_t>(nb) * sizeof(int)));
  OVO_CUDA(tmp.alloc(&sorted, N * sizeof(float4)));
  OVO_CUDA(tmp.alloc(&overflow, Q * sizeof(int)));
  OVO_CUDA(tmp.alloc(&n_overflow, sizeof(int)));
  OVO_CUDA(cudaMemsetAsync(cursor, 0, static_cast<size_t>(n_cells) * sizeof(int), s));
  OVO_CUDA(cudaMemsetAsync(n_overflow, 0, sizeof(int), s));
  {
    ProfScope prof(s, PROF_OTHER, 0.0, static_cast<double>(N) * 36);
    scan_blocks_kernel<<<nb, 256, 0, s>>>(starts, n_cells + 1, totals);
    OVO_CHECK_LAUNCH();
    scan_totals_kernel<<<1, 1024, 0, s>>>(totals, nb);
    OVO_CHECK_LAUNCH();
    scan_add_kernel<<<ceil_div(n_cells + 1, 256), 256, 0, s>>>(starts, n_cells + 1, totals);
    OVO_CHECK_LAUNCH();
    knn_scatter_kernel<<<ceil_div(N, 256), 256, 0, s>>>(points_dev, N, cell_of_pt, starts, cursor, sorted);
    OVO_CHECK_LAUNCH();
  }
  {
    ProfScope prof(s, PROF_OTHER, 0.0, static_cast<double>(Q) * (12 + 12.0 * k));
    if (k == 1) knn_query_kernel<1><<<ceil_div(Q * 32, 128), 128, 0, s>>>(sorted, starts, g, queries_dev, Q, k, idx_out_dev, dist_out_dev, overflow, n_overflow);
    else if (k <= 5) knn_query_kernel<5><<<ceil_div(Q * 32, 128), 128, 0, s>>>(sorted, starts, g, queries_dev, Q, k, idx_out_dev, dist_out_dev, overflow, n_overflow);
    else knn_query_kernel<kKnnMaxK><<<ceil_div(Q * 32, 128), 128, 0, s>>>(sorted, starts, g, queries_dev, Q, k, idx_out_dev, dist_out_dev, overflow, n_overflow);
    OVO_CHECK_LAUNCH();
  }
  int n_over = 0;
  OVO_CUDA(cudaMemcpyAsync(&n_over, n_overflow, sizeof(int), cudaMemcpyDeviceToHost, s));
  OVO_CUDA(cudaStreamSynchronize(s));
  g_knn_cell = g.cell; g_knn_cells = n_cells; g_knn_overflow = n_over;
  if (n_over > 0) {
    knn_brute_kernel<<<n_over, 256, 0, s>>>(sorted, N, queries_dev, overflow, k, idx_out_dev, dist_out_dev);
    OVO_CHECK_LAUNCH();
  }
  return OVO_OK;   // `tmp` releases the temporaries in stream order
}

int ovo_knn_mode(const int32_t* labels_dev, const int32_t* idx_dev, int64_t Q, int k, int32_t* out_dev, void* stream) {
  OVO_REQUIRE(labels_dev && idx_dev && out_dev && Q > 0 && k >= 1 && k <= ovo::kKnnMaxK, "ovo_knn_mode: bad arguments");
  ovo::knn_mode_kernel<<<ovo::ceil_div(Q, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(labels_dev, idx_dev, Q, k, out_dev);
  OVO_CHECK_LAUNCH();
  return OVO_OK;
}

}  // extern "C"
