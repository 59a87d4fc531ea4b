// 3D association, instance vote, dense/instance fusion and query kernels (HBM-bound streaming passes).
//
// Replaces, on device and without per-mask host syncs:
//   geometry_utils.compute_camera_frustum_corners / compute_frustum_point_ids  (geometry_utils.py:99-129,252-277)
//   geometry_utils.match_3d_points_to_2d_pixels / project_3d_points           (geometry_utils.py:26-89)
//   geometry_utils.depth_filter                                               (geometry_utils.py:92-96)
//   OVO._match_and_track_instances / _track_objects                          (ovo.py:204-229,240-282)
//   Instance3D avg_pooling fusion                                            (instance3d.py:19-21)
//   clip_utils.clip_cosine_similarity / OVO.classify_instances               (clip_utils.py:16-19, ovo.py:486-491)
//
// Floating point that feeds integer decisions uses __fmul_rn/__fadd_rn/__fdiv_rn in a fixed left-to-right
// order (no FMA contraction) so the results are bit-identical to oracle/fusion.py.
#include <vector>

#include "common.cuh"
#include "gemm.cuh"
#include "p2p.cuh"

namespace ovo {

struct FrameGeom {
  float lo[3], hi[3];
  float planes[6][4];
  int dmin_bits, dmax_bits;  // depth>0 min / max as int bit patterns (positive floats order like ints)
};

struct FrameDev {
  float c2w[16], w2c[16], K[9];
  float match_th;
  int h, w, H, W;
  int has_ratio;
  float ratio_h, ratio_w;
  int crop_edge;
};

// ------------------------------------------------------------------------------------------ depth min/max
__global__ void depth_minmax_kernel(const float* __restrict__ depth, int n, FrameGeom* g) {
  int lmin = 0x7f800000, lmax = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float d = depth[i];
    if (d > 0.f) {
      const int b = __float_as_int(d);
      lmin = min(lmin, b);
      lmax = max(lmax, b);
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    lmin = min(lmin, __shfl_xor_sync(0xffffffffu, lmin, o));
    lmax = max(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(&g->dmin_bits, lmin);
    atomicMax(&g->dmax_bits, lmax);
  }
}

// positive floats order like their bit patterns: the range is reduced as two ints and IS the two floats
__global__ void depth_range_init_kernel(int* r) { r[0] = 0x7f800000; r[1] = 0; }
__global__ void depth_range_kernel(const float* __restrict__ depth, int n, int* __restrict__ r) {
  int lmin = 0x7f800000, lmax = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float d = depth[i];
    if (d > 0.f) { lmin = min(lmin, __float_as_int(d)); lmax = max(lmax, __float_as_int(d)); }
  }
  for (int o = 16; o > 0; o >>= 1) {
    lmin = min(lmin, __shfl_xor_sync(0xffffffffu, lmin, o));
    lmax = max(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
  }
  if ((threadIdx.x & 31) == 0) { atomicMin(&r[0], lmin); atomicMax(&r[1], lmax); }
}

// F depth maps [F][h][w] in ONE launch: the high-pass filter of every map (out, may be null) and the [min, max] of its values > 0
// (ranges [F][2] as int bit patterns, initialised by depth_range_init_batch_kernel)
__global__ void depth_range_init_batch_kernel(int* r, int F) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < F) { r[2 * i] = 0x7f800000; r[2 * i + 1] = 0; }
}

__device__ __forceinline__ float dot4_rn(const float* m, float x, float y, float z) {

  float acc = __fmul_rn(m[0], x);
  acc = __fadd_rn(acc, __fmul_rn(m[1], y));
  acc = __fadd_rn(acc, __fmul_rn(m[2], z));
  return __fadd_rn(acc, m[3]);
}

// corners (geometry_utils.py:110-129), planes (:163-207), AABB (:210-221); one thread, fixed op order.
__device__ __forceinline__ void frustum_setup_body(FrameGeom* g, const FrameDev* f) {
  const float dmin = __int_as_float(g->dmin_bits), dmax = __int_as_float(g->dmax_bits);
  const float wf = static_cast<float>(f->w), hf = static_cast<float>(f->h);
  const float px[8] = {0.f, wf, 0.f, wf, 0.f, wf, 0.f, wf};
  const float py[8] = {0.f, 0.f, hf, hf, 0.f, 0.f, hf, hf};
  float c[8][3];
  for (int i = 0; i < 8; ++i) {
    const float pz = i < 4 ? dmin : dmax;
    const float x = __fdiv_rn(__fmul_rn(__fsub_rn(px[i], f->K[2]), pz), f->K[0]);
    const float y = __fdiv_rn(__fmul_rn(__fsub_rn(py[i], f->K[5]), pz), f->K[4]);
    for (int r = 0; r < 3; ++r) c[i][r] = dot4_rn(&f->c2w[4 * r], x, y, pz);
  }
  for (int r = 0; r < 3; ++r) {
    float lo = c[0][r], hi = c[0][r];
    for (int i = 1; i < 8; ++i) {
      lo = fminf(lo, c[i][r]);
      hi = fmaxf(hi, c[i][r]);
    }
    g->lo[r] = lo;
    g->hi[r] = hi;
  }
  const int pairs[6][4] = {{2, 0, 1, 0}, {6, 4, 5, 4}, {4, 0, 2, 0}, {7, 3, 1, 3}, {5, 1, 3, 1}, {6, 2, 0, 2}};
  for (int i = 0; i < 6; ++i) {
    float a[3], b[3];
    for (int r = 0; r < 3; ++r) {
      a[r] = __fsub_rn(c[pairs[i][0]][r], c[pairs[i][1]][r]);
      b[r] = __fsub_rn(c[pairs[i][2]][r], c[pairs[i][3]][r]);
    }
    const float n0 = __fsub_rn(__fmul_rn(a[1], b[2]), __fmul_rn(a[2], b[1]));
    const float n1 = __fsub_rn(__fmul_rn(a[2], b[0]), __fmul_rn(a[0], b[2]));
    const float n2 = __fsub_rn(__fmul_rn(a[0], b[1]), __fmul_rn(a[1], b[0]));
    const float d =
        -__fadd_rn(__fadd_rn(__fmul_rn(n0, c[i][0]), __fmul_rn(n1, c[i][1])), __fmul_rn(n2, c[i][2]));
    g->planes[i][0] = n0; g->planes[i][1] = n1; g->planes[i][2] = n2; g->planes[i][3] = d;
  }
}
__global__ void frustum_setup_kernel(FrameGeom* g, const FrameDev* f) { frustum_setup_body(g, f); }

// ------------------------------------------------------------------------------------------ depth filter
// torchvision _get_gaussian_kernel1d(7, 2.5, float32) bit patterns (see oracle/fusion.py)
__constant__ uint32_t kGauss7Bits[7] = {1035802123u, 1041042090u, 1043549328u, 1044527997u,
                                        1043549328u, 1041042090u, 1035802123u};

__device__ __forceinline__ void depth_filter_body(const float* __restrict__ depth, int h, int w, float* __restrict__ out, int x, int y) {
  if (x >= w || y >= h) return;
  float acc = 0.f;
#pragma unroll
  for (int dy = 0; dy < 7; ++dy) {
    int yy = y + dy - 3;
    yy = yy < 0 ? -yy : (yy >= h ? 2 * h - 2 - yy : yy);  // reflect (no edge repeat)
    const float ky = __uint_as_float(kGauss7Bits[dy]);
#pragma unroll
    for (int dx = 0; dx < 7; ++dx) {
      int xx = x + dx - 3;
      xx = xx < 0 ? -xx : (xx >= w ? 2 * w - 2 - xx : xx);
      const float k2 = __fmul_rn(ky, __uint_as_float(kGauss7Bits[dx]));
      acc = __fadd_rn(acc, __fmul_rn(k2, __ldg(depth + static_cast<size_t>(yy) * w + xx)));
    }
  }
  const float d = depth[static_cast<size_t>(y) * w + x];
  out[static_cast<size_t>(y) * w + x] = fabsf(__fsub_rn(d, acc)) > 0.05f ? -1.f : d;
}
__global__ void depth_filter_kernel(const float* __restrict__ depth, int h, int w, float* __restrict__ out) {
  depth_filter_body(depth, h, w, out, blockIdx.x * blockDim.x + threadIdx.x, blockIdx.y * blockDim.y + threadIdx.y);
}
__global__ void depth_filter_range_batch_kernel(const float* __restrict__ depth, int h, int w, float* __restrict__ out, int* __restrict__ ranges) {
  const size_t off = static_cast<size_t>(blockIdx.z) * h * w;
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (out != nullptr) depth_filter_body(depth + off, h, w, out + off, x, y);
  if (ranges != nullptr) {
    int lmin = 0x7f800000, lmax = 0;
    if (x < w && y < h) {
      const float d = depth[off + static_cast<size_t>(y) * w + x];
      if (d > 0.f) { lmin = __float_as_int(d); lmax = lmin; }
    }
    for (int o = 16; o > 0; o >>= 1) {
      lmin = min(lmin, __shfl_xor_sync(0xffffffffu, lmin, o));
      lmax = max(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
    }
    if (((threadIdx.y * blockDim.x + threadIdx.x) & 31) == 0 && lmax != 0) {
      atomicMin(&ranges[2 * blockIdx.z], lmin);
      atomicMax(&ranges[2 * blockIdx.z + 1], lmax);
    }
  }
}

// ------------------------------------------------------------------------------------------ mask areas
__global__ void seg_area_kernel(const int32_t* __restrict__ seg, int n, int n_masks, int32_t* __restrict__ area) {
  extern __shared__ int32_t hist[];
  for (int i = threadIdx.x; i < n_masks; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int s = seg[i];
    if (s >= 0 && s < n_masks) atomicAdd(&hist[s], 1);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_masks; i += blockDim.x)
    if (hist[i]) atomicAdd(&area[i], hist[i]);
}

// ------------------------------------------------------------------------------------------ pass 1
// One streaming pass over the map: cull + project + depth test + seg lookup + vote histogram + match list.
// votes: [n_masks][n_ins + 1] (column 0 = unassigned points, column 1+id = points already carrying id).
constexpr int kP1Threads = 256;

__global__ void __launch_bounds__(kP1Threads)
    associate_pass1_kernel(const float* __restrict__ xyz, const int32_t* __restrict__ ins_ids, long long N,
                           const float* __restrict__ depth, const int32_t* __restrict__ seg_map,
                           const FrameGeom* __restrict__ geom, const FrameDev* __restrict__ fr, int n_masks,
                           int n_ins, int32_t* __restrict__ votes, int2* __restrict__ match_list,
                           int32_t* __restrict__ counters /* [0]=list len, [1]=n_matched */, int smem_votes) {
  // when the vote table fits, votes are first accumulated per block in shared memory (few hot global addresses
  // otherwise: every matched point of a mask hits the same counter)
  extern __shared__ int32_t s_votes[];
  const int n_votes = n_masks * (n_ins + 1);
  if (smem_votes)
    for (int i = threadIdx.x; i < n_votes; i += blockDim.x) s_votes[i] = 0;
  __shared__ float s_xyz[kP1Threads * 3];
  __shared__ FrameGeom s_g;
  __shared__ FrameDev s_f;
  __shared__ int s_warp_cnt[kP1Threads / 32], s_base, s_matched;
  int n_matched_local = 0;
  if (threadIdx.x == 0) s_matched = 0;
  if (threadIdx.x < sizeof(FrameGeom) / 4) reinterpret_cast<int*>(&s_g)[threadIdx.x] = reinterpret_cast<const int*>(geom)[threadIdx.x];
  for (int i = threadIdx.x; i < static_cast<int>(sizeof(FrameDev) / 4); i += blockDim.x)
    reinterpret_cast<int*>(&s_f)[i] = reinterpret_cast<const int*>(fr)[i];
  const int lane = threadIdx.x & 31;

  for (long long base = static_cast<long long>(blockIdx.x) * kP1Threads; base < N;
       base += static_cast<long long>(gridDim.x) * kP1Threads) {
    __syncthreads();
    // coalesced, vectorised stage of 256 points (3072 B) through shared memory
    const long long fbase = base * 3;
    const long long fend = min(N * 3, fbase + kP1Threads * 3);
    if (fend - fbase == kP1Threads * 3) {  // fbase*4 is a multiple of 3072 B -> 16 B aligned
      const float4* src = reinterpret_cast<const float4*>(xyz + fbase);
      if (threadIdx.x < kP1Threads * 3 / 4) reinterpret_cast<float4*>(s_xyz)[threadIdx.x] = __ldg(src + threadIdx.x);
    } else {
      for (int i = threadIdx.x; i < fend - fbase; i += kP1Threads) s_xyz[i] = xyz[fbase + i];
    }
    __syncthreads();
    const long long idx = base + threadIdx.x;
    bool matched = false;
    int seg = -1;
    if (idx < N) {
      const float x = s_xyz[threadIdx.x * 3], y = s_xyz[threadIdx.x * 3 + 1], z = s_xyz[threadIdx.x * 3 + 2];
      bool in = x >= s_g.lo[0] && x <= s_g.hi[0] && y >= s_g.lo[1] && y <= s_g.hi[1] && z >= s_g.lo[2] &&
                z <= s_g.hi[2];
      if (in) {
#pragma unroll
        for (int p = 0; p < 6; ++p) in = in && (dot4_rn(s_g.planes[p], x, y, z) <= 0.f);
      }
      if (in) {
        const float lx = dot4_rn(&s_f.w2c[0], x, y, z), ly = dot4_rn(&s_f.w2c[4], x, y, z);
        const float lz = dot4_rn(&s_f.w2c[8], x, y, z), lw = dot4_rn(&s_f.w2c[12], x, y, z);
        const float X = __fdiv_rn(lx, lw), Y = __fdiv_rn(ly, lw), Z = __fdiv_rn(lz, lw);
        float ph[3];
#pragma unroll
        for (int r = 0; r < 3; ++r)
          ph[r] = __fadd_rn(__fadd_rn(__fmul_rn(s_f.K[3 * r], X), __fmul_rn(s_f.K[3 * r + 1], Y)),
                            __fmul_rn(s_f.K[3 * r + 2], Z));
        const float uf = rintf(__fdiv_rn(ph[0], ph[2])), vf = rintf(__fdiv_rn(ph[1], ph[2]));
        if (fabsf(uf) < 2e9f && fabsf(vf) < 2e9f) {  // also rejects NaN/Inf
          int u = static_cast<int>(uf), v = static_cast<int>(vf);
          if (u >= 0 && v >= 0 && u < s_f.w && v < s_f.h) {
            const float d = __ldg(depth + static_cast<size_t>(v) * s_f.w + u);
            if (fabsf(__fsub_rn(lz, d)) < s_f.match_th && d != 0.f) {
              matched = true;
              if (s_f.has_ratio) {  // ovo.py:218-221
                u += s_f.crop_edge;
                v += s_f.crop_edge;
                v = static_cast<int>(__fmul_rn(static_cast<float>(v), s_f.ratio_h));
                u = static_cast<int>(__fmul_rn(static_cast<float>(u), s_f.ratio_w));
              }
              if (u >= 0 && v >= 0 && u < s_f.W && v < s_f.H) seg = __ldg(seg_map + static_cast<size_t>(v) * s_f.W + u);
              if (seg >= n_masks) seg = -1;
            }
          }
        }
      }
    }
    // bookkeeping: warp-aggregated inside the block, ONE global atomic per block and 256 points for the list cursor (every
    // warp hitting the same two global counters serialised ~60k same-address atomics per pass); the matched count is kept
    // per thread and added once per block at the end
    const unsigned m_all = __ballot_sync(0xffffffffu, matched);
    if (lane == 0) n_matched_local += __popc(m_all);
    const bool listed = matched && seg >= 0;
    const unsigned m_list = __ballot_sync(0xffffffffu, listed);
    const int warp = threadIdx.x >> 5;
    if (lane == 0) s_warp_cnt[warp] = __popc(m_list);
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
#pragma unroll
      for (int w = 0; w < kP1Threads / 32; ++w) { const int c = s_warp_cnt[w]; s_warp_cnt[w] = tot; tot += c; }
      s_base = tot ? atomicAdd(&counters[0], tot) : 0;
    }
    __syncthreads();
    if (listed) {
      const int pos = s_base + s_warp_cnt[warp] + __popc(m_list & ((1u << lane) - 1));
      match_list[pos] = make_int2(static_cast<int>(idx), seg);
      int id = ins_ids[idx];
      if (id >= n_ins) id = -1;  // ids the host does not know about count as unassigned
      const int key = seg * (n_ins + 1) + (id + 1);
      const unsigned peers = __match_any_sync(m_list, key);
      if (lane == __ffs(peers) - 1) atomicAdd(smem_votes ? &s_votes[key] : &votes[key], __popc(peers));
    }
  }
  if (n_matched_local) atomicAdd(&s_matched, n_matched_local);
  __syncthreads();
  if (threadIdx.x == 0 && s_matched) atomicAdd(&counters[1], s_matched);
  if (smem_votes) {
    __syncthreads();
    for (int i = threadIdx.x; i < n_votes; i += blockDim.x)
      if (s_votes[i]) atomicAdd(&votes[i], s_votes[i]);
  }
}

// ------------------------------------------------------------------------------------------ vote reduce
// One block per mask: n_unassigned, n_assigned, mode of assigned ids (ties -> smallest id = torch.mode on CPU).
__device__ __forceinline__ void vote_reduce_body(const int32_t* __restrict__ votes, int n_ins, const int32_t* __restrict__ area,
                                                 ovo_vote_row* __restrict__ rows) {
  const int m = blockIdx.x;
  const int32_t* row = votes + static_cast<size_t>(m) * (n_ins + 1);
  int best_cnt = 0, best_id = 0x7fffffff, total = 0;
  for (int i = threadIdx.x; i < n_ins; i += blockDim.x) {
    const int c = row[1 + i];
    total += c;
    if (c > best_cnt || (c == best_cnt && c > 0 && i < best_id)) {
      best_cnt = c;
      best_id = i;
    }
  }
  __shared__ int s_cnt[32], s_id[32], s_tot[32];
  for (int o = 16; o > 0; o >>= 1) {
    const int oc = __shfl_xor_sync(0xffffffffu, best_cnt, o), oi = __shfl_xor_sync(0xffffffffu, best_id, o);
    total += __shfl_xor_sync(0xffffffffu, total, o);
    if (oc > best_cnt || (oc == best_cnt && oi < best_id)) {
      best_cnt = oc;
      best_id = oi;
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    s_cnt[warp] = best_cnt; s_id[warp] = best_id; s_tot[warp] = total;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int nw = blockDim.x >> 5;
    for (int i = 1; i < nw; ++i) {
      total += s_tot[i];
      if (s_cnt[i] > best_cnt || (s_cnt[i] == best_cnt && s_id[i] < best_id)) {
        best_cnt = s_cnt[i];
        best_id = s_id[i];
      }
    }
    ovo_vote_row r;
    r.n_unassigned = row[0];
    r.n_assigned = total;
    r.n_matched = total + row[0];
    r.mode_id = best_cnt > 0 ? best_id : -1;
    r.ins_id = -1;
    r.is_new = 0;
    r.area = area[m];
    r.reserved = 0;
    rows[m] = r;
  }
}
__global__ void vote_reduce_kernel(const int32_t* __restrict__ votes, int n_ins, const int32_t* __restrict__ area,
                                   ovo_vote_row* __restrict__ rows) {
  vote_reduce_body(votes, n_ins, area, rows);
}

// New ids are allocated in mask order (ovo.py:255,271-273): every mask decides in parallel, then an ordered
// prefix sum over the "wants a new instance" flags hands out next_ins_id, next_ins_id+1, ...  One block.
__device__ __forceinline__ void vote_decide_body(ovo_vote_row* rows, int n_masks, int track_th, int32_t* mask_ins, int32_t* next_ins_id) {
  __shared__ int s_warp[8];
  __shared__ int s_base;
  if (threadIdx.x == 0) s_base = *next_ins_id;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int m0 = 0; m0 < n_masks; m0 += blockDim.x) {
    const int m = m0 + threadIdx.x;
    ovo_vote_row r;
    int want_new = 0, ins = -1;
    if (m < n_masks) {
      r = rows[m];
      if (r.n_matched > track_th) {
        if (r.n_assigned > track_th) ins = r.mode_id;
        else if (r.n_unassigned > track_th) want_new = 1;
      }
    }
    // ordered exclusive prefix of want_new over the block
    const unsigned bal = __ballot_sync(0xffffffffu, want_new);
    const int in_warp = __popc(bal & ((1u << lane) - 1));
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    int before = 0, total = 0;
    for (int w = 0; w < 8; ++w) {
      if (w < warp) before += s_warp[w];
      total += s_warp[w];
    }
    if (m < n_masks) {
      if (want_new) { ins = s_base + before + in_warp; r.is_new = 1; }
      r.ins_id = ins;
      rows[m] = r;
      mask_ins[m] = ins;
    }
    __syncthreads();
    if (threadIdx.x == 0) s_base += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) *next_ins_id = s_base;
}
__global__ void __launch_bounds__(256)
    vote_decide_kernel(ovo_vote_row* rows, int n_masks, int track_th, int32_t* mask_ins, int32_t* next_ins_id) {
  vote_decide_body(rows, n_masks, track_th, mask_ins, next_ins_id);
}

// Per-mask reduce + decisions by ONE block of 256 threads (a warp per mask): the tail of the batched vote.  votes
// [n_masks][n_ins+1] was written by other blocks / other GPUs: read through L2.
__device__ __forceinline__ void vote_finish_block(const int32_t* votes, int n_ins, int n_masks, const int32_t* __restrict__ area,
                                                  ovo_vote_row* rows, int track_th, int32_t* mask_ins, int32_t* next_ins_id) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int m = warp; m < n_masks; m += nw) {
    const int32_t* row = votes + static_cast<size_t>(m) * (n_ins + 1);
    int best_cnt = 0, best_id = 0x7fffffff, total = 0;
    for (int i = lane; i < n_ins; i += 32) {
      const int c = __ldcg(row + 1 + i);
      total += c;
      if (c > best_cnt || (c == best_cnt && c > 0 && i < best_id)) { best_cnt = c; best_id = i; }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const int oc = __shfl_xor_sync(0xffffffffu, best_cnt, o), oi = __shfl_xor_sync(0xffffffffu, best_id, o);
      total += __shfl_xor_sync(0xffffffffu, total, o);
      if (oc > best_cnt || (oc == best_cnt && oi < best_id)) { best_cnt = oc; best_id = oi; }
    }
    if (lane == 0) {
      ovo_vote_row r;
      r.n_unassigned = __ldcg(row);
      r.n_assigned = total;
      r.n_matched = total + r.n_unassigned;
      r.mode_id = best_cnt > 0 ? best_id : -1;
      r.ins_id = -1; r.is_new = 0; r.area = area[m]; r.reserved = 0;
      rows[m] = r;
    }
  }
  __syncthreads();
  vote_decide_body(rows, n_masks, track_th, mask_ins, next_ins_id);
}

// ------------------------------------------------------------------------------------------ pass 2
__global__ void associate_pass2_kernel(const int2* __restrict__ match_list, const int32_t* __restrict__ counters,
                                       const int32_t* __restrict__ mask_ins, int32_t* __restrict__ ins_ids) {
  const int n = counters[0];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int2 e = match_list[i];
    const int id = mask_ins[e.y];
    if (id >= 0 && ins_ids[e.x] == -1) ins_ids[e.x] = id;  // assigned points never change (ovo.py:274,280)
  }
}

// ------------------------------------------------------------------------------------------ batched association
// All keyframes of a batch against the map in ONE pass over xyz (ovo_map_associate_batch): the geometry of keyframe f does not
// depend on the instance ids, only its vote does.  Pass A therefore writes, for every keyframe f and point p, the mask the
// point was matched into (seg_of_pt[f][p], int16, -1 = none); the votes are then taken keyframe by keyframe over these dense
// rows (coalesced 2 + 4 B per point) with the id decisions left on the device, and the host reads all rows back once.
struct BatchFrame {
  FrameDev fr;
  FrameGeom geom;
  const float* depth_raw;     // frustum from the raw depth (ovo.py:209)
  const float* range;         // optional [2]: min / max of the raw depth > 0 computed elsewhere (then depth_raw is not scanned)
  const float* depth_used;    // matching against the filtered depth (ovo.py:213-216)
  float* depth_filtered;      // workspace (nullptr: no filter)
  const int32_t* seg_map;
  int n_masks, track_th;
  int votes_off;              // offset of this keyframe's table [n_matched, 0, 0, 0 | votes] in the batch's table buffer (ints)
  int nm_off;                 // offset of its n_matched counter (the table's first int)
};

__global__ void batch_depth_minmax_kernel(BatchFrame* __restrict__ fr) {
  BatchFrame& b = fr[blockIdx.y];
  if (b.range != nullptr) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      b.geom.dmin_bits = __float_as_int(b.range[0]);
      b.geom.dmax_bits = __float_as_int(b.range[1]);
    }
    return;
  }
  const int n = b.fr.h * b.fr.w;
  int lmin = 0x7f800000, lmax = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float d = b.depth_raw[i];
    if (d > 0.f) {
      const int v = __float_as_int(d);
      lmin = min(lmin, v);
      lmax = max(lmax, v);
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    lmin = min(lmin, __shfl_xor_sync(0xffffffffu, lmin, o));
    lmax = max(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(&b.geom.dmin_bits, lmin);
    atomicMax(&b.geom.dmax_bits, lmax);
  }
}

__global__ void batch_frustum_setup_kernel(BatchFrame* fr) { frustum_setup_body(&fr[blockIdx.x].geom, &fr[blockIdx.x].fr); }

__global__ void batch_depth_filter_kernel(const BatchFrame* __restrict__ fr) {
  const BatchFrame& b = fr[blockIdx.z];
  if (b.depth_filtered == nullptr) return;
  depth_filter_body(b.depth_raw, b.fr.h, b.fr.w, b.depth_filtered, blockIdx.x * blockDim.x + threadIdx.x, blockIdx.y * blockDim.y + threadIdx.y);
}

__global__ void batch_seg_area_kernel(const BatchFrame* __restrict__ fr, int32_t* __restrict__ area, int area_stride) {
  extern __shared__ int32_t hist[];
  const BatchFrame& b = fr[blockIdx.y];
  const int n_masks = b.n_masks, n = b.fr.H * b.fr.W;
  for (int i = threadIdx.x; i < n_masks; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int s = b.seg_map[i];
    if (s >= 0 && s < n_masks) atomicAdd(&hist[s], 1);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_masks; i += blockDim.x)
    if (hist[i]) atomicAdd(&area[blockIdx.y * area_stride + i], hist[i]);
}

// cull + project + depth test + seg lookup of one point against one keyframe (same operations, same order as pass 1 /
// oracle/fusion.py associate); returns matched, *seg = mask index or -1
__device__ __forceinline__ bool match_point_seg(float x, float y, float z, const FrameGeom& g, const FrameDev& f,
                                                const float* __restrict__ depth, const int32_t* __restrict__ seg_map,
                                                int n_masks, int* seg) {
  *seg = -1;
  bool in = x >= g.lo[0] && x <= g.hi[0] && y >= g.lo[1] && y <= g.hi[1] && z >= g.lo[2] && z <= g.hi[2];
  if (!in) return false;
#pragma unroll
  for (int p = 0; p < 6; ++p) in = in && (dot4_rn(g.planes[p], x, y, z) <= 0.f);
  if (!in) return false;
  const float lx = dot4_rn(&f.w2c[0], x, y, z), ly = dot4_rn(&f.w2c[4], x, y, z);
  const float lz = dot4_rn(&f.w2c[8], x, y, z), lw = dot4_rn(&f.w2c[12], x, y, z);
  const float X = __fdiv_rn(lx, lw), Y = __fdiv_rn(ly, lw), Z = __fdiv_rn(lz, lw);
  float ph[3];
#pragma unroll
  for (int r = 0; r < 3; ++r)
    ph[r] = __fadd_rn(__fadd_rn(__fmul_rn(f.K[3 * r], X), __fmul_rn(f.K[3 * r + 1], Y)), __fmul_rn(f.K[3 * r + 2], Z));
  const float uf = rintf(__fdiv_rn(ph[0], ph[2])), vf = rintf(__fdiv_rn(ph[1], ph[2]));
  if (!(fabsf(uf) < 2e9f && fabsf(vf) < 2e9f)) return false;   // also rejects NaN/Inf
  int u = static_cast<int>(uf), v = static_cast<int>(vf);
  if (u < 0 || v < 0 || u >= f.w || v >= f.h) return false;
  const float d = __ldg(depth + static_cast<size_t>(v) * f.w + u);
  if (!(fabsf(__fsub_rn(lz, d)) < f.match_th && d != 0.f)) return false;
  if (f.has_ratio) {  // ovo.py:218-221
    u += f.crop_edge;
    v += f.crop_edge;
    v = static_cast<int>(__fmul_rn(static_cast<float>(v), f.ratio_h));
    u = static_cast<int>(__fmul_rn(static_cast<float>(u), f.ratio_w));
  }
  if (u >= 0 && v >= 0 && u < f.W && v < f.H) {
    const int s = __ldg(seg_map + static_cast<size_t>(v) * f.W + u);
    *seg = s < n_masks ? s : -1;
  }
  return true;
}

constexpr int kBatchFramesPerPass = 16;   // keyframes whose constants sit in shared memory at once

__global__ void __launch_bounds__(kP1Threads)
    associate_batch_pass_kernel(const float* __restrict__ xyz, long long N, const BatchFrame* __restrict__ frames, int f0, int nf,
                                int16_t* __restrict__ seg_of_pt /* [F][stride] */, long long stride,
                                int32_t* __restrict__ tables /* n_matched of keyframe f at frames[f].nm_off */) {
  __shared__ float s_xyz[kP1Threads * 3];
  __shared__ BatchFrame s_fr[kBatchFramesPerPass];
  __shared__ int s_matched[kBatchFramesPerPass];
  for (int i = threadIdx.x; i < nf * static_cast<int>(sizeof(BatchFrame) / 4); i += blockDim.x)
    reinterpret_cast<int*>(s_fr)[i] = reinterpret_cast<const int*>(frames + f0)[i];
  if (threadIdx.x < kBatchFramesPerPass) s_matched[threadIdx.x] = 0;
  const int lane = threadIdx.x & 31;
  for (long long base = static_cast<long long>(blockIdx.x) * kP1Threads; base < N;
       base += static_cast<long long>(gridDim.x) * kP1Threads) {
    __syncthreads();
    const long long fbase = base * 3;
    const long long fend = min(N * 3, fbase + kP1Threads * 3);
    if (fend - fbase == kP1Threads * 3) {
      const float4* src = reinterpret_cast<const float4*>(xyz + fbase);
      if (threadIdx.x < kP1Threads * 3 / 4) reinterpret_cast<float4*>(s_xyz)[threadIdx.x] = __ldg(src + threadIdx.x);
    } else {
      for (int i = threadIdx.x; i < fend - fbase; i += kP1Threads) s_xyz[i] = xyz[fbase + i];
    }
    __syncthreads();
    const long long idx = base + threadIdx.x;
    const bool live = idx < N;
    const float x = s_xyz[threadIdx.x * 3], y = s_xyz[threadIdx.x * 3 + 1], z = s_xyz[threadIdx.x * 3 + 2];
    for (int f = 0; f < nf; ++f) {
      int seg = -1;
      const bool matched = live && match_point_seg(x, y, z, s_fr[f].geom, s_fr[f].fr, s_fr[f].depth_used, s_fr[f].seg_map,
                                                   s_fr[f].n_masks, &seg);
      if (live) seg_of_pt[static_cast<size_t>(f0 + f) * stride + idx] = static_cast<int16_t>(seg);
      const unsigned bal = __ballot_sync(0xffffffffu, matched);
      if (lane == 0 && bal) atomicAdd(&s_matched[f], __popc(bal));
    }
  }
  __syncthreads();
  if (threadIdx.x < nf && s_matched[threadIdx.x]) atomicAdd(&tables[s_fr[threadIdx.x].nm_off], s_matched[threadIdx.x]);
}

// Keyframe f of the batch: first give the points keyframe f-1 matched their new ids (its decisions are final), then vote.
// One coalesced scan of the two dense rows and the ids, 8 points per thread (16-byte loads of the int16 rows: the rows are
// padded to a multiple of 8 points); votes go to shared memory when the table fits.
// Fused tail (tail.enabled): the LAST block to finish its share of the scan (ticket counter) goes on alone — on a sharded map it
// pushes the table into every peer's inbox over NVLink, waits for theirs and sums (p2p.cuh), then reduces the votes per mask and
// takes the id decisions: one launch per keyframe instead of scan + exchange + reduce + decide.
struct VoteTail {
  int enabled;
  int32_t* done;                 // ticket counter (0 between launches)
  int32_t* table;                // [n_matched, 0, 0, 0 | votes]
  const int32_t* area; ovo_vote_row* rows; int32_t* mask_ins; int32_t* next_ins_id; int32_t* n_matched_out; int track_th;
  int world, rank, slots, slot, parity, epoch; long long table_cap;
  XchgPeers peers;
};

__global__ void __launch_bounds__(256)
    batch_vote_scan_kernel(const int16_t* __restrict__ seg_prev, const int32_t* __restrict__ mask_ins_prev,
                           const int16_t* __restrict__ seg_cur, int32_t* __restrict__ ins_ids, long long N,
                           const int32_t* __restrict__ n_ins_ptr, int n_masks, int32_t* __restrict__ votes, int smem_ints,
                           const __grid_constant__ VoteTail tail) {
  // votes of a block are gathered in a STATIC 10 KB table when the keyframe's compact table fits (the instance count is only known
  // on the device, so the size cannot be a launch parameter); larger tables take RED.ADD straight to L2 (500k votes of a 2M-point
  // keyframe over ~100 hot addresses cost ~25 us that way).  10 KB + the 1 KB the driver reserves per block is what fits on an SM
  // BESIDE a resident encoder GEMM CTA (214.6 KB of the 228 KB, 53.7k of the 64k registers; this kernel: 256 threads x 40): the
  // per-keyframe chain of a batch then advances while the encoder runs instead of waiting for the gaps between its launches.
  constexpr int kSmemVotes = 2560;
  __shared__ int32_t s_votes[kSmemVotes];
  (void)smem_ints;
  const int n_ins = *n_ins_ptr;
  const int n_votes = n_masks * (n_ins + 1);
  const bool use_smem = seg_cur != nullptr && n_votes <= kSmemVotes;
  if (use_smem)
    for (int i = threadIdx.x; i < n_votes; i += blockDim.x) s_votes[i] = 0;
  __syncthreads();
  const long long n8 = (N + 7) >> 3;
  for (long long q = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; q < n8;
       q += static_cast<long long>(gridDim.x) * blockDim.x) {
    uint4 vp = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu), vc = vp;
    if (seg_prev) vp = *reinterpret_cast<const uint4*>(seg_prev + 8 * q);
    if (seg_cur) vc = *reinterpret_cast<const uint4*>(seg_cur + 8 * q);
    // any point of the eight matched by either keyframe?  (int16 -1 = 0xffff; mask indices are < 0x8000)
    const uint32_t neg = (vp.x & vp.y & vp.z & vp.w & vc.x & vc.y & vc.z & vc.w & 0x80008000u);
    if (neg == 0x80008000u) continue;
    const int16_t* sp8 = reinterpret_cast<const int16_t*>(&vp);
    const int16_t* sc8 = reinterpret_cast<const int16_t*>(&vc);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const long long p = 8 * q + j;
      const int sp = sp8[j], sc = sc8[j];
      if ((sp < 0 && sc < 0) || p >= N) continue;
      int id = ins_ids[p];
      if (sp >= 0 && id == -1) {   // assigned points never change (ovo.py:274,280)
        const int nid = mask_ins_prev[sp];
        if (nid >= 0) { id = nid; ins_ids[p] = nid; }
      }
      if (sc >= 0) {
        if (id >= n_ins) id = -1;  // ids the decisions do not know about count as unassigned
        const int key = sc * (n_ins + 1) + (id + 1);
        atomicAdd(use_smem ? &s_votes[key] : &votes[key], 1);
      }
    }
  }
  if (use_smem) {
    __syncthreads();
    for (int i = threadIdx.x; i < n_votes; i += blockDim.x)
      if (s_votes[i]) atomicAdd(&votes[i], s_votes[i]);
  }
  if (!tail.enabled) return;
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(tail.done, 1) == static_cast<int>(gridDim.x) - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (threadIdx.x == 0) *tail.done = 0;
  if (tail.world > 1)
    xchg_block(tail.peers, tail.rank, tail.world, tail.slots, tail.slot, tail.parity, tail.epoch, tail.table_cap, tail.table,
               4 + max(n_masks, 1) * (n_ins + 1));
  if (threadIdx.x == 0) *tail.n_matched_out = __ldcg(tail.table);
  vote_finish_block(tail.table + 4, n_ins, n_masks, tail.area, tail.rows, tail.track_th, tail.mask_ins, tail.next_ins_id);
}

// ------------------------------------------------------------------------------------------ persistent batch votes
// ALL keyframes of a batch in ONE launch (sharded maps: 64 dependent per-keyframe launches, each waiting for the slowest of 8
// GPUs under a busy encoder, made the vote chain the critical path).  One 256-thread block per SM at most, resident for the
// whole batch; per keyframe f:
//   every block : ids of keyframe f-1 -> its points, votes of keyframe f of its slice of the map (static shared-memory table,
//                 flushed to the keyframe's table in L2), then a ticket on `sync[0]`;
//   block 0     : waits until every block's ticket for f is in, pushes the table into the peers' inboxes / waits for theirs /
//                 sums (p2p.cuh), reduces the votes per mask, takes the id decisions, releases `sync[1] = f + 1`;
//   every block : waits for that release (ld.acquire.gpu) before touching keyframe f+1.
// The blocks of a batch need not all be resident at once for progress: a waiting block only polls, and the ones not yet
// scheduled are ahead of every later kernel in the block scheduler's queue.
struct VotePersist {
  const BatchFrame* frames;      // n_masks, track_th, votes_off per keyframe
  int F;
  const int16_t* seg; long long stride;      // dense rows [F][stride]
  int32_t* ins_ids; long long N;
  int32_t* tables;
  const int32_t* area; ovo_vote_row* rows; int32_t* mask_ins; int rows_stride;     // [F][rows_stride]
  int32_t* next_ins_id; int32_t* n_matched_out;                                   // device counter, [F]
  int32_t* sync;                 // [0] block tickets, [1] released keyframes (both zero at launch)
  int world, rank, slots; long long table_cap;
  XchgPeers peers;
  int epoch[64];                 // the exchange's epoch of every keyframe slot
};

__device__ __forceinline__ int ld_acquire_gpu(const int32_t* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// kSmem: gather a block's votes in a static 10 KB table first (single GPU: 500k votes per keyframe); without it (sharded maps: a
// few 10^4 votes per keyframe and rank, RED.ADD to L2) the resident block leaves the SM's shared memory to the encoder — two
// attention CTAs need 226.5 of the 228 KB.
// (<= 40 registers, forced: 256 threads x 40 is what fits beside a resident encoder GEMM CTA; at 60 the block kept the next GEMM's
// CTA off its SM for the whole batch and the step got SLOWER)
template <bool kSmem>
__global__ void __launch_bounds__(256, 6) batch_vote_persistent_kernel(const __grid_constant__ VotePersist P) {
  constexpr int kSmemVotes = kSmem ? 2560 : 1;
  __shared__ int32_t s_votes[kSmemVotes];
  const long long n8 = (P.N + 7) >> 3;
  for (int f = 0; f <= P.F; ++f) {          // f == F: only the ids of the last keyframe
    const int n_ins = ld_acquire_gpu(P.next_ins_id);
    const bool vote = f < P.F;
    const int n_masks = vote ? P.frames[f].n_masks : 0;
    int32_t* table = vote ? P.tables + P.frames[f].votes_off : nullptr;
    int32_t* votes = vote ? table + 4 : nullptr;
    const int n_votes = n_masks * (n_ins + 1);
    const bool use_smem = kSmem && vote && n_votes <= kSmemVotes;
    if (use_smem)
      for (int i = threadIdx.x; i < n_votes; i += blockDim.x) s_votes[i] = 0;
    __syncthreads();
    const int16_t* seg_prev = f > 0 ? P.seg + static_cast<size_t>(f - 1) * P.stride : nullptr;
    const int16_t* seg_cur = vote ? P.seg + static_cast<size_t>(f) * P.stride : nullptr;
    const int32_t* mask_ins_prev = f > 0 ? P.mask_ins + static_cast<size_t>(f - 1) * P.rows_stride : nullptr;
    for (long long q = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; q < n8;
         q += static_cast<long long>(gridDim.x) * blockDim.x) {
      uint4 vp = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu), vc = vp;
      if (seg_prev) vp = *reinterpret_cast<const uint4*>(seg_prev + 8 * q);
      if (seg_cur) vc = *reinterpret_cast<const uint4*>(seg_cur + 8 * q);
      const uint32_t neg = (vp.x & vp.y & vp.z & vp.w & vc.x & vc.y & vc.z & vc.w & 0x80008000u);
      if (neg == 0x80008000u) continue;
      const int16_t* sp8 = reinterpret_cast<const int16_t*>(&vp);
      const int16_t* sc8 = reinterpret_cast<const int16_t*>(&vc);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const long long p = 8 * q + j;
        const int sp = sp8[j], sc = sc8[j];
        if ((sp < 0 && sc < 0) || p >= P.N) continue;
        int id = P.ins_ids[p];
        if (sp >= 0 && id == -1) {
          const int nid = __ldcg(mask_ins_prev + sp);     // written by block 0 during this launch: L2, not a stale L1 line
          if (nid >= 0) { id = nid; P.ins_ids[p] = nid; }
        }
        if (sc >= 0) {
          if (id >= n_ins) id = -1;
          const int key = sc * (n_ins + 1) + (id + 1);
          atomicAdd(use_smem ? &s_votes[key] : &votes[key], 1);
        }
      }
    }
    if (!vote) break;
    if (use_smem) {
      __syncthreads();
      for (int i = threadIdx.x; i < n_votes; i += blockDim.x)
        if (s_votes[i]) atomicAdd(&votes[i], s_votes[i]);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(&P.sync[0], 1);
    if (blockIdx.x == 0) {
      if (threadIdx.x == 0)
        while (ld_acquire_gpu(&P.sync[0]) < static_cast<int>(gridDim.x) * (f + 1)) __nanosleep(100);
      __syncthreads();
      if (P.world > 1)
        xchg_block(P.peers, P.rank, P.world, P.slots, f, P.epoch[f] & 1, P.epoch[f], P.table_cap, table, 4 + max(n_masks, 1) * (n_ins + 1));
      if (threadIdx.x == 0) P.n_matched_out[f] = __ldcg(table);
      vote_finish_block(votes, n_ins, n_masks, P.area + static_cast<size_t>(f) * P.rows_stride, P.rows + static_cast<size_t>(f) * P.rows_stride,
                        P.frames[f].track_th, P.mask_ins + static_cast<size_t>(f) * P.rows_stride, P.next_ins_id);
      __threadfence();
      __syncthreads();
      if (threadIdx.x == 0) atomicExch(&P.sync[1], f + 1);
    } else {
      if (threadIdx.x == 0)
        while (ld_acquire_gpu(&P.sync[1]) < f + 1) __nanosleep(500);   // (the exchange takes microseconds: poll gently, the SM is shared)
      __syncthreads();
    }
  }
}

// the same tail alone (staged protocol: the table was summed by a host-launched collective; or an empty shard)
__global__ void __launch_bounds__(256) batch_vote_finish_kernel(const __grid_constant__ VoteTail tail, int n_masks) {
  const int n_ins = *tail.next_ins_id;
  if (tail.world > 1)
    xchg_block(tail.peers, tail.rank, tail.world, tail.slots, tail.slot, tail.parity, tail.epoch, tail.table_cap, tail.table,
               4 + max(n_masks, 1) * (n_ins + 1));
  if (threadIdx.x == 0) *tail.n_matched_out = __ldcg(tail.table);
  vote_finish_block(tail.table + 4, n_ins, n_masks, tail.area, tail.rows, tail.track_th, tail.mask_ins, tail.next_ins_id);
}

// ------------------------------------------------------------------------------------------ dense fusion
// The dense per-point bank keeps, per map point, the running mean of the descriptors of the masks it fell into, as TWO
// bf16 planes: `hi` = the mean rounded to bf16 (the operand of the cosine query, read once from HBM by ovo_query_dense) and
// `lo` = bf16(mean - hi), the part of the mean the first plane cannot hold.  hi + lo carries 16-17 significant bits, so the
// increment (e - f)/c of a long stream (c in the hundreds: below half a bf16 ulp of f) is not lost.
// Update of one point by the k descriptors e_1..e_k a pass brings (keyframe order), every operation rounded to f32, no FMA
// contraction (oracle/fusion.py dense_update restates it):
//     f = hi + lo;  s = e_1 + ... + e_k  (e_i = the descriptor rounded to bf16);  c' = c + k
//     f' = f + (s - k*f) * (1/c');  hi' = bf16(f');  lo' = bf16(f' - hi')
// One warp per touched point: lane l owns elements 8*(l + 32*i) .. +7 of the row (16 B per access).
template <int kVecPerLane>
struct DenseRow {
  uint4 rh[kVecPerLane], rl[kVecPerLane];   // the two planes as loaded (unpacked only in dense_row_finish: 8 registers, not 32)
  float s[kVecPerLane * 8];                  // sum of this pass's descriptors
};

__device__ __forceinline__ void unpack_bf16x8(const uint4& u, float* o) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    o[2 * j] = __uint_as_float(w[j] << 16);
    o[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u);
  }
}

// kFull: the part is exactly kVecPerLane * 256 columns wide (no per-vector guards: straight-line code)
template <int kVecPerLane, bool kFull = false>
__device__ __forceinline__ void dense_row_load(DenseRow<kVecPerLane>& r, const __nv_bfloat16* hi, const __nv_bfloat16* lo,
                                               size_t p, int D, int Dp, int lane) {
  const uint4* ph = reinterpret_cast<const uint4*>(hi + p * D);
  const uint4* pl = reinterpret_cast<const uint4*>(lo + p * D);
  const int nvec = Dp >> 3;
#pragma unroll
  for (int i = 0; i < kVecPerLane; ++i) {
    const int v = lane + 32 * i;
    if (kFull || v < nvec) { r.rh[i] = ph[v]; r.rl[i] = pl[v]; }
  }
}

// s = e (first descriptor of the pass) / s += e, descriptor rows in bf16
template <int kVecPerLane, bool kFull = false>
__device__ __forceinline__ void dense_row_set(DenseRow<kVecPerLane>& r, const uint4 (&e)[kVecPerLane], int Dp, int lane) {
  const int nvec = Dp >> 3;
#pragma unroll
  for (int i = 0; i < kVecPerLane; ++i)
    if (kFull || lane + 32 * i < nvec) unpack_bf16x8(e[i], &r.s[8 * i]);
}
template <int kVecPerLane, bool kFull = false>
__device__ __forceinline__ void dense_row_add(DenseRow<kVecPerLane>& r, const uint4 (&e)[kVecPerLane], int Dp, int lane) {
  const int nvec = Dp >> 3;
#pragma unroll
  for (int i = 0; i < kVecPerLane; ++i) {
    if (kFull || lane + 32 * i < nvec) {
      float a[8];
      unpack_bf16x8(e[i], a);
#pragma unroll
      for (int j = 0; j < 8; ++j) r.s[8 * i + j] = __fadd_rn(r.s[8 * i + j], a[j]);
    }
  }
}

template <int kVecPerLane, bool kFull = false>
__device__ __forceinline__ void dense_row_finish(DenseRow<kVecPerLane>& r, int k, int c_new, __nv_bfloat16* hi,
                                                 __nv_bfloat16* lo, size_t p, int D, int Dp, int lane) {
  const float kf = static_cast<float>(k), inv = __fdiv_rn(1.0f, static_cast<float>(c_new));
  uint4* ph = reinterpret_cast<uint4*>(hi + p * D);
  uint4* pl = reinterpret_cast<uint4*>(lo + p * D);
  const int nvec = Dp >> 3;
#pragma unroll
  for (int i = 0; i < kVecPerLane; ++i) {
    const int v = lane + 32 * i;
    if (kFull || v < nvec) {
      float a[8], b[8];
      unpack_bf16x8(r.rh[i], a);
      unpack_bf16x8(r.rl[i], b);
      uint32_t wh[4], wl[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float f1[2];
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const float f = __fadd_rn(a[2 * j + t], b[2 * j + t]);
          f1[t] = __fadd_rn(f, __fmul_rn(__fsub_rn(r.s[8 * i + 2 * j + t], __fmul_rn(kf, f)), inv));
        }
        wh[j] = pack_bf16(f1[0], f1[1]);
        const float h0 = __uint_as_float(wh[j] << 16), h1 = __uint_as_float(wh[j] & 0xffff0000u);
        wl[j] = pack_bf16(__fsub_rn(f1[0], h0), __fsub_rn(f1[1], h1));
      }
      ph[v] = make_uint4(wh[0], wh[1], wh[2], wh[3]);
      pl[v] = make_uint4(wl[0], wl[1], wl[2], wl[3]);
    }
  }
}

template <int kVecPerLane, bool kFull = false>
__device__ __forceinline__ void load_desc_row(uint4 (&e)[kVecPerLane], const __nv_bfloat16* feats, int row, int D, int Dp, int lane) {
  const uint4* fr = reinterpret_cast<const uint4*>(feats + static_cast<size_t>(row) * D);
  const int nvec = Dp >> 3;
#pragma unroll
  for (int i = 0; i < kVecPerLane; ++i) {
    const int v = lane + 32 * i;
    if (kFull || v < nvec) e[i] = __ldg(fr + v);
  }
}

// one keyframe, from its match list: one warp per matched point (a point appears once in a keyframe's list)
template <int kVecPerLane>
__global__ void __launch_bounds__(256)
    fuse_dense_kernel(const int2* __restrict__ match_list, int n, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                      int32_t* __restrict__ counts, int D, const __nv_bfloat16* __restrict__ feats,
                      const int32_t* __restrict__ mask_row) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  for (int e = blockIdx.x * warps_per_block + (threadIdx.x >> 5); e < n; e += gridDim.x * warps_per_block) {
    const int2 m = match_list[e];
    const int r = mask_row[m.y];
    if (r < 0) continue;
    const int c = counts[m.x] + 1;
    __syncwarp();
    if (lane == 0) counts[m.x] = c;
    DenseRow<kVecPerLane> row;
    uint4 ev[kVecPerLane];
    load_desc_row<kVecPerLane>(ev, feats, r, D, D, lane);
    dense_row_load<kVecPerLane>(row, hi, lo, static_cast<size_t>(m.x), D, D, lane);
    dense_row_set<kVecPerLane>(row, ev, D, lane);
    dense_row_finish<kVecPerLane>(row, 1, c, hi, lo, static_cast<size_t>(m.x), D, D, lane);
  }
}

// ------------------------------------------------------------------------------------------ batched dense fusion
// Several keyframes fused in ONE pass over the bank: a point matched in k keyframes of the batch has its two 2 KB rows
// read once and written once instead of k times, and its k descriptors are summed in registers (one add per element and
// keyframe).  Input: seg_of_pt [F][N] int16 = the mask each keyframe matched the point into (-1 = none), written by the
// batched association or scattered from the keyframes' match lists.
__global__ void scatter_matches_kernel(const int2* __restrict__ list, int n, int16_t* __restrict__ seg_of_pt) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int2 e = list[i];
    seg_of_pt[e.x] = static_cast<int16_t>(e.y);
  }
}

// A warp scans 32 consecutive points (lane = point) for "matched by some keyframe", then takes the touched points one at a
// time: lane f looks up the descriptor row of keyframe f for the point (seg_of_pt -> mask_row: F independent two-step
// gathers in one round), the rows are broadcast with shuffles and two descriptor rows are in flight per step.
// A warp owns kVecPerLane * 256 COLUMNS of the rows (blockIdx.y = which part): with 512 columns per warp the kernel needs
// ~80 registers, so 24 warps per SM are resident instead of 16 (the pass is latency-bound: ncu long_scoreboard).  The
// counts are therefore not written here (another part's warp may still need the old value): fuse_counts_kernel follows.
// 128-thread blocks of <= 80 registers: one of them fits on an SM beside a resident encoder GEMM CTA (see batch_scan), seven when
// the SM is free.
template <int kVecPerLane, bool kFull>
__global__ void __launch_bounds__(128, kVecPerLane <= 2 ? 6 : 1)
    fuse_dense_batch_kernel(const int16_t* __restrict__ seg_of_pt /* [F][stride] */, long long stride, int F, long long N,
                            __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, const int32_t* __restrict__ counts,
                            int D, const __nv_bfloat16* __restrict__ feats, const int32_t* __restrict__ mask_row /* [F][n_masks] */,
                            int n_masks) {
  const int lane = threadIdx.x & 31;
  const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const int groups = (F + 31) >> 5;   // keyframes are looked up 32 at a time (lane = keyframe)
  const int col0 = blockIdx.y * kVecPerLane * 256;   // first column of this warp's part
  const int Dp = min(D - col0, kVecPerLane * 256);   // columns of this part
  for (long long p0 = warp * 32; p0 < N; p0 += n_warps * 32) {
    const long long pl = p0 + lane;
    bool any = false;
    if (pl < N)
      for (int f = 0; f < F; ++f) any = any || (seg_of_pt[static_cast<size_t>(f) * stride + pl] >= 0);
    unsigned todo_pts = __ballot_sync(0xffffffffu, any);
    while (todo_pts) {
      const int src = __ffs(todo_pts) - 1;
      todo_pts &= todo_pts - 1;
      const size_t p = static_cast<size_t>(p0 + src);
      int rlane[2] = {-1, -1};   // descriptor row of keyframe 32*g + lane for this point (F <= 64)
      int k = 0;
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const int f = 32 * g + lane;
        if (g < groups && f < F) {
          const int m = seg_of_pt[static_cast<size_t>(f) * stride + p];
          if (m >= 0 && m < n_masks) rlane[g] = mask_row[f * n_masks + m];
        }
        k += __popc(__ballot_sync(0xffffffffu, rlane[g] >= 0));
      }
      if (k == 0) continue;   // matched only into masks that produced no descriptor
      DenseRow<kVecPerLane> row;
      dense_row_load<kVecPerLane, kFull>(row, hi + col0, lo + col0, p, D, Dp, lane);
      const int c = counts[p];
      bool have = false;
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        unsigned todo = __ballot_sync(0xffffffffu, rlane[g] >= 0);
        while (todo) {
          const int f0 = __ffs(todo) - 1;
          todo &= todo - 1;
          const int f1 = todo ? __ffs(todo) - 1 : -1;
          if (f1 >= 0) todo &= todo - 1;
          const int r0 = __shfl_sync(0xffffffffu, rlane[g], f0);
          const int r1 = __shfl_sync(0xffffffffu, rlane[g], f1 >= 0 ? f1 : f0);
          uint4 e0[kVecPerLane], e1[kVecPerLane];
          load_desc_row<kVecPerLane, kFull>(e0, feats + col0, r0, D, Dp, lane);
          if (f1 >= 0) load_desc_row<kVecPerLane, kFull>(e1, feats + col0, r1, D, Dp, lane);
          if (have) dense_row_add<kVecPerLane, kFull>(row, e0, Dp, lane);
          else dense_row_set<kVecPerLane, kFull>(row, e0, Dp, lane);
          have = true;
          if (f1 >= 0) dense_row_add<kVecPerLane, kFull>(row, e1, Dp, lane);
        }
      }
      dense_row_finish<kVecPerLane, kFull>(row, k, c + k, hi + col0, lo + col0, p, D, Dp, lane);
    }
  }
}

// counts[p] += number of keyframes of the pass that brought a descriptor for p (the k of fuse_dense_batch_kernel)
__global__ void fuse_counts_kernel(const int16_t* __restrict__ seg_of_pt, long long stride, int F, long long N,
                                   int32_t* __restrict__ counts, const int32_t* __restrict__ mask_row, int n_masks) {
  for (long long p = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; p < N;
       p += static_cast<long long>(gridDim.x) * blockDim.x) {
    int k = 0;
    for (int f = 0; f < F; ++f) {
      const int m = seg_of_pt[static_cast<size_t>(f) * stride + p];
      if (m >= 0 && m < n_masks && mask_row[f * n_masks + m] >= 0) ++k;
    }
    if (k) counts[p] += k;
  }
}

__global__ void bank_update_mean_kernel(float* __restrict__ bank, int32_t* __restrict__ counts, int D,
                                        const float* __restrict__ feats, const int32_t* __restrict__ rows, int n) {
  const int i = blockIdx.x;
  if (i >= n) return;
  const int r = rows[i];
  if (r < 0) return;
  const int c = counts[r] + 1;
  __syncthreads();
  if (threadIdx.x == 0) counts[r] = c;
  const float cf = static_cast<float>(c);
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const float b = bank[static_cast<size_t>(r) * D + d];
    bank[static_cast<size_t>(r) * D + d] = __fadd_rn(b, __fdiv_rn(__fsub_rn(feats[static_cast<size_t>(i) * D + d], b), cf));
  }
}

// avg_pooling with k more views (instance3d.py:19-21): mean_{n+k} = (n * mean_n + sum of the k new descriptors) / (n + k).
// quads [m][4] = (bank row, views already behind the row n, first and one-past-last position in idx); idx = descriptor-store rows.
__global__ void bank_add_views_kernel(float* __restrict__ bank, int D, const float* __restrict__ store,
                                      const int32_t* __restrict__ quads, const int32_t* __restrict__ idx) {
  const int br = quads[4 * blockIdx.x], n = quads[4 * blockIdx.x + 1], i0 = quads[4 * blockIdx.x + 2], i1 = quads[4 * blockIdx.x + 3];
  const float nf = static_cast<float>(n), inv = __fdiv_rn(1.0f, static_cast<float>(n + (i1 - i0)));
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float acc = __fmul_rn(bank[static_cast<size_t>(br) * D + d], nf);
    for (int i = i0; i < i1; ++i) acc = __fadd_rn(acc, store[static_cast<size_t>(idx[i]) * D + d]);
    bank[static_cast<size_t>(br) * D + d] = __fmul_rn(acc, inv);
  }
}

// ------------------------------------------------------------------------------------------ query helpers
__global__ void f32_to_bf16_pad_kernel(const float* __restrict__ src, int rows, int cols, __nv_bfloat16* __restrict__ dst,
                                       int rows_pad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows_pad * cols) return;
  const int r = i / cols;
  dst[i] = __float2bfloat16_rn(r < rows ? src[i] : 0.f);
}

// one warp per (instance, query) dot product in f32, lanes stride the feature dimension
__global__ void query_instances_kernel(const float* __restrict__ bank, const int32_t* __restrict__ rows, int I, int D,
                                       const float* __restrict__ text, int Q, float* __restrict__ out) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= I * Q) return;
  const int i = w / Q, q = w - i * Q;
  const size_t br = rows ? static_cast<size_t>(rows[i]) : static_cast<size_t>(i);
  float acc = 0.f;
  for (int d = lane; d < D; d += 32) acc = fmaf(bank[br * D + d], text[static_cast<size_t>(q) * D + d], acc);
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) out[w] = acc;
}

__global__ void classify_kernel(const float* __restrict__ sim, long long n, int Q, float th, int32_t* __restrict__ cls,
                                float* __restrict__ conf) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* row = sim + i * Q;
  float best = row[0];
  int bi = 0;
  for (int q = 1; q < Q; ++q) {
    const float v = row[q];
    if (v > best) {  // first maximum wins, like torch.argmax
      best = v;
      bi = q;
    }
  }
  const bool keep = best > th;
  cls[i] = keep ? bi : -1;
  conf[i] = keep ? best : 0.f;
}


// ------------------------------------------------------------------------------------------ mask merge
// OVO._fuse_masks_with_same_ins_id (ovo.py:284-324): masks voted to the same instance are OR-ed into one row;
// out[r] = OR of the masks m with group[m] == r; areas[r] = popcount(out[r]).  One thread per 4 pixels.
__global__ void merge_masks_kernel(const uint8_t* __restrict__ masks, int M, int npix4, const int32_t* __restrict__ group,
                                   int R, uint8_t* __restrict__ out, int32_t* __restrict__ areas) {
  const int r = blockIdx.y;
  int local = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npix4; i += gridDim.x * blockDim.x) {
    uint32_t acc = 0;
    for (int m = 0; m < M; ++m)
      if (group[m] == r) acc |= reinterpret_cast<const uint32_t*>(masks + static_cast<size_t>(m) * npix4 * 4)[i];
    // normalise every non-zero byte to 1
    acc = ((acc | (acc >> 1) | (acc >> 2) | (acc >> 3) | (acc >> 4) | (acc >> 5) | (acc >> 6) | (acc >> 7)) & 0x01010101u);
    reinterpret_cast<uint32_t*>(out + static_cast<size_t>(r) * npix4 * 4)[i] = acc;
    local += __popc(acc);
  }
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(&areas[r], local);
}

// ------------------------------------------------------------------------------------------ view fusion
// Instance3D.update_clip (instance3d.py:157-189) for a batch of instances: instance j fuses the descriptor rows
// store[idx[off[j] .. off[j+1])] with mode 0 = avg_pooling (:19-21), 1 = l1_medoid (:9-12), 2 = cossim_medoid
// (:14-17) into bank[out_rows[j]]; chosen[j] = medoid index (0 for avg).  One block per instance.
__global__ void __launch_bounds__(256)
    fuse_views_kernel(const float* __restrict__ store, int D, const int32_t* __restrict__ idx, const int32_t* __restrict__ off,
                      int mode, float* __restrict__ bank, const int32_t* __restrict__ out_rows, int32_t* __restrict__ chosen) {
  const int j = blockIdx.x;
  const int beg = off[j], n = off[j + 1] - beg;
  if (n <= 0) return;
  float* dst = bank + static_cast<size_t>(out_rows[j]) * D;
  __shared__ float s_red[8];
  __shared__ float s_best;
  __shared__ int s_besti;
  if (mode == 0 || n == 1) {
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
      float acc = 0.f;
      for (int i = 0; i < n; ++i) acc += store[static_cast<size_t>(idx[beg + i]) * D + d];
      dst[d] = n == 1 ? acc : acc / static_cast<float>(n);
    }
    if (threadIdx.x == 0 && chosen) chosen[j] = 0;
    return;
  }
  if (threadIdx.x == 0) { s_best = mode == 1 ? INFINITY : -INFINITY; s_besti = 0; }
  __syncthreads();
  for (int a = 0; a < n; ++a) {
    const float* ra = store + static_cast<size_t>(idx[beg + a]) * D;
    float score = 0.f;  // sum over b of L1(a,b)  or  sum over b of cos(a,b)
    for (int b = 0; b < n; ++b) {
      const float* rb = store + static_cast<size_t>(idx[beg + b]) * D;
      float p0 = 0.f, p1 = 0.f, p2 = 0.f;
      for (int d = threadIdx.x; d < D; d += blockDim.x) {
        const float x = ra[d], y = rb[d];
        if (mode == 1) p0 += fabsf(x - y);
        else { p0 += x * y; p1 += x * x; p2 += y * y; }
      }
      float v[3] = {p0, p1, p2};
      for (int t = 0; t < 3; ++t) {
        float r = v[t];
        for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = r;
        __syncthreads();
        r = 0.f;
        for (int w = 0; w < (blockDim.x >> 5); ++w) r += s_red[w];
        v[t] = r;
      }
      score += mode == 1 ? v[0] : v[0] / fmaxf(sqrtf(v[1]) * sqrtf(v[2]), 1e-8f);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      if ((mode == 1 && score < s_best) || (mode == 2 && score > s_best)) { s_best = score; s_besti = a; }
    }
    __syncthreads();
  }
  const float* src = store + static_cast<size_t>(idx[beg + s_besti]) * D;
  for (int d = threadIdx.x; d < D; d += blockDim.x) dst[d] = src[d];
  if (threadIdx.x == 0 && chosen) chosen[j] = s_besti;
}


// ------------------------------------------------------------------------------------------ S2: mask NMS + seg-map
// masks_update / mask_nms / mask2segmap (ovo/utils/segment_utils.py:173-259,12-27) on the device: masks are
// bit-packed, pairwise intersections are popcounts, the keep rules and the paint order follow the reference
// (including tril(diagonal=1) keeping the first super-diagonal, segment_utils.py:236).
__global__ void bitpack_masks_kernel(const uint8_t* __restrict__ masks, int M, int npix, int nwords, uint32_t* __restrict__ bits) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = blockIdx.y;
  if (w >= nwords) return;
  const uint8_t* src = masks + static_cast<size_t>(m) * npix + static_cast<size_t>(w) * 32;
  uint32_t v = 0;
  const int n = min(32, npix - w * 32);
  for (int i = 0; i < n; ++i) v |= (src[i] != 0 ? 1u : 0u) << i;
  bits[static_cast<size_t>(m) * nwords + w] = v;
}

// rank[i] = position of mask i when sorted by descending key (ties: lower index first = stable sort)
__global__ void rank_desc_kernel(const float* __restrict__ key, int M, int32_t* __restrict__ rank, int32_t* __restrict__ order) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  const float k = key[i];
  int r = 0;
  for (int j = 0; j < M; ++j) {
    const float kj = key[j];
    r += (kj > k) || (kj == k && j < i);
  }
  rank[i] = r;
  order[r] = i;
}

// inter[a][b] = |mask(order[a]) & mask(order[b])| for a <= b (sorted order); one block per pair
__global__ void mask_intersections_kernel(const uint32_t* __restrict__ bits, int nwords, const int32_t* __restrict__ order,
                                          int M, int32_t* __restrict__ inter) {
  const int a = blockIdx.y, b = blockIdx.x;
  if (b < a) return;
  const uint32_t* ba = bits + static_cast<size_t>(order[a]) * nwords;
  const uint32_t* bb = bits + static_cast<size_t>(order[b]) * nwords;
  int acc = 0;
  for (int w = threadIdx.x; w < nwords; w += blockDim.x) acc += __popc(ba[w] & bb[w]);
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ int s[8];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int i = 0; i < (blockDim.x >> 5); ++i) t += s[i];
    inter[a * M + b] = t;
    inter[b * M + a] = t;
  }
}

// keep rules of mask_nms (segment_utils.py:225-257), one thread per column j of the sorted order
__global__ void mask_nms_decide_kernel(const int32_t* __restrict__ inter, const float* __restrict__ scores,
                                       const int32_t* __restrict__ order, int M, float iou_thr, float score_thr, float inner_lim,
                                       uint8_t* __restrict__ keep_sorted) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  const float area_j = static_cast<float>(inter[j * M + j]);
  float iou_max = 0.f, inner_u = 0.f, inner_l = 0.f;
  // inner[r][c] is written for pairs (i<=j'): [i][j'] when ri<0.5 && rj>=0.85, [j'][i] when ri>=0.85 && rj<0.5
  for (int i = 0; i < M; ++i) {
    const float it = static_cast<float>(inter[i * M + j]);
    const float area_i = static_cast<float>(inter[i * M + i]);
    if (i < j) {  // pair (i, j), i < j: column j of the upper triangle
      const float uni = __fsub_rn(__fadd_rn(area_i, area_j), it);
      iou_max = fmaxf(iou_max, __fdiv_rn(it, uni));
      const float ri = __fdiv_rn(it, area_i), rj = __fdiv_rn(it, area_j);
      if (ri < 0.5f && rj >= 0.85f) inner_u = fmaxf(inner_u, __fsub_rn(1.f, __fmul_rn(rj, ri)));   // inner[i][j], upper
      // tril(diagonal=1) also keeps inner[j-1][j]
      if (i == j - 1 && ri < 0.5f && rj >= 0.85f) inner_l = fmaxf(inner_l, __fsub_rn(1.f, __fmul_rn(rj, ri)));
    } else if (i > j) {  // pair (j, i) with j < i: entry [i][j] (lower triangle) set when r_j >= 0.85 && r_i < 0.5
      const float rjj = __fdiv_rn(it, area_j), rii = __fdiv_rn(it, area_i);
      if (rjj >= 0.85f && rii < 0.5f) inner_l = fmaxf(inner_l, __fsub_rn(1.f, __fmul_rn(rii, rjj)));
    }
  }
  const bool keep = iou_max <= iou_thr && scores[order[j]] > score_thr && inner_u <= inner_lim && inner_l <= inner_lim;
  keep_sorted[j] = keep ? 1 : 0;
}

__global__ void gather_u8_kernel(const uint8_t* __restrict__ src, const int32_t* __restrict__ idx, int n, uint8_t* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[idx[i]];
}

// mask2segmap: pixel gets the position (in descending-stability order) of the first mask that contains it
__global__ void paint_segmap_kernel(const uint8_t* __restrict__ masks, const int32_t* __restrict__ order, int M, int npix,
                                    int32_t* __restrict__ seg, uint8_t* __restrict__ maps_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix) return;
  int s = -1;
  for (int k = 0; k < M; ++k) {
    const uint8_t v = masks[static_cast<size_t>(order[k]) * npix + i] != 0;
    maps_out[static_cast<size_t>(k) * npix + i] = v;
    if (v && s < 0) s = k;
  }
  seg[i] = s;
}


// ------------------------------------------------------------------------------------------ map producer
// VanillaMapper.map (ovo/slam/vanilla_mapper.py:46-85): depth pixels that are not explained by a map point yet are
// un-projected and appended.  Same cull + project + depth test as the association (match_th 0.03, raw depth).
__device__ __forceinline__ bool match_point_rn(float x, float y, float z, const FrameGeom& g, const FrameDev& f,
                                               const float* __restrict__ depth, int& u, int& v) {
  bool in = x >= g.lo[0] && x <= g.hi[0] && y >= g.lo[1] && y <= g.hi[1] && z >= g.lo[2] && z <= g.hi[2];
  if (!in) return false;
#pragma unroll
  for (int p = 0; p < 6; ++p) in = in && (dot4_rn(g.planes[p], x, y, z) <= 0.f);
  if (!in) return false;
  const float lx = dot4_rn(&f.w2c[0], x, y, z), ly = dot4_rn(&f.w2c[4], x, y, z);
  const float lz = dot4_rn(&f.w2c[8], x, y, z), lw = dot4_rn(&f.w2c[12], x, y, z);
  const float X = __fdiv_rn(lx, lw), Y = __fdiv_rn(ly, lw), Z = __fdiv_rn(lz, lw);
  float ph[3];
#pragma unroll
  for (int r = 0; r < 3; ++r)
    ph[r] = __fadd_rn(__fadd_rn(__fmul_rn(f.K[3 * r], X), __fmul_rn(f.K[3 * r + 1], Y)), __fmul_rn(f.K[3 * r + 2], Z));
  const float uf = rintf(__fdiv_rn(ph[0], ph[2])), vf = rintf(__fdiv_rn(ph[1], ph[2]));
  if (!(fabsf(uf) < 2e9f && fabsf(vf) < 2e9f)) return false;
  u = static_cast<int>(uf);
  v = static_cast<int>(vf);
  if (u < 0 || v < 0 || u >= f.w || v >= f.h) return false;
  const float d = __ldg(depth + static_cast<size_t>(v) * f.w + u);
  return fabsf(__fsub_rn(lz, d)) < f.match_th && d != 0.f;
}

__global__ void mark_mapped_pixels_kernel(const float* __restrict__ xyz, long long N, const float* __restrict__ depth,
                                          const FrameGeom* __restrict__ geom, const FrameDev* __restrict__ fr,
                                          uint8_t* __restrict__ mapped) {
  __shared__ FrameGeom s_g;
  __shared__ FrameDev s_f;
  if (threadIdx.x < sizeof(FrameGeom) / 4) reinterpret_cast<int*>(&s_g)[threadIdx.x] = reinterpret_cast<const int*>(geom)[threadIdx.x];
  for (int i = threadIdx.x; i < static_cast<int>(sizeof(FrameDev) / 4); i += blockDim.x)
    reinterpret_cast<int*>(&s_f)[i] = reinterpret_cast<const int*>(fr)[i];
  __syncthreads();
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < N;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    int u, v;
    if (match_point_rn(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], s_g, s_f, depth, u, v)) mapped[static_cast<size_t>(v) * s_f.w + u] = 1;
  }
}

// flag[(y/ds)*(wd)+(x/ds)] = 1 iff pixel (y,x) (y,x multiples of ds) yields a new point: depth > 0 and, when the map
// is not empty, the whole k x k neighbourhood is valid and unmapped (~maxpool(~mask), padding counts as valid)
__global__ void new_point_flags_kernel(const float* __restrict__ depth, const uint8_t* __restrict__ mapped, int h, int w,
                                       int ds, int kpool, int erode, int hd, int wd, int32_t* __restrict__ flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= hd * wd) return;
  const int y = (i / wd) * ds, x = (i % wd) * ds;
  bool ok = depth[static_cast<size_t>(y) * w + x] > 0.f;
  if (erode) {
    ok = ok && !mapped[static_cast<size_t>(y) * w + x];
    const int r = kpool / 2;
    for (int dy = -r; dy <= r && ok; ++dy)
      for (int dx = -r; dx <= r; ++dx) {
        const int yy = y + dy, xx = x + dx;
        if (yy < 0 || xx < 0 || yy >= h || xx >= w) continue;
        if (!(depth[static_cast<size_t>(yy) * w + xx] > 0.f) || mapped[static_cast<size_t>(yy) * w + xx]) { ok = false; break; }
      }
  }
  flags[i] = ok ? 1 : 0;
}

// single-block exclusive scan (n <= a few 100k): flags -> positions, total in *count
__global__ void __launch_bounds__(1024) scan_flags_kernel(const int32_t* __restrict__ flags, int n, int32_t* __restrict__ pos,
                                                          int32_t* __restrict__ count) {
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < n; base += 1024) {
    const int i = base + threadIdx.x;
    const int f = i < n ? flags[i] : 0;
    int incl = f;
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int w = s_warp[lane];
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += t;
      }
      s_warp[lane] = w;
    }
    __syncthreads();
    const int before = s_carry + (warp ? s_warp[warp - 1] : 0) + incl - f;
    if (i < n) pos[i] = before;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = before + f;
    __syncthreads();
  }
  if (threadIdx.x == 0) *count = s_carry;
}

__global__ void emit_points_kernel(const int32_t* __restrict__ flags, const int32_t* __restrict__ pos, int hd, int wd, int ds,
                                   const float* __restrict__ depth, const uint8_t* __restrict__ rgb, int w,
                                   const FrameDev* __restrict__ fr, long long capacity_left, float* __restrict__ xyz_out,
                                   int32_t* __restrict__ ids_out, int32_t* __restrict__ ins_out, uint8_t* __restrict__ col_out,
                                   int base_id) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= hd * wd || !flags[i]) return;
  const int k = pos[i];
  if (k >= capacity_left) return;
  const int y = (i / wd) * ds, x = (i % wd) * ds;
  const float d = depth[static_cast<size_t>(y) * w + x];
  // vanilla_mapper.py:73-79 in a fixed f32 order
  const float X = __fdiv_rn(__fmul_rn(__fsub_rn(static_cast<float>(x), fr->K[2]), d), fr->K[0]);
  const float Y = __fdiv_rn(__fmul_rn(__fsub_rn(static_cast<float>(y), fr->K[5]), d), fr->K[4]);
#pragma unroll
  for (int r = 0; r < 3; ++r) xyz_out[3 * static_cast<size_t>(k) + r] = dot4_rn(&fr->c2w[4 * r], X, Y, d);
  ids_out[k] = base_id + k;
  ins_out[k] = -1;
  if (col_out != nullptr && rgb != nullptr) {
    const uint8_t* c = rgb + (static_cast<size_t>(y) * w + x) * 3;
    col_out[3 * static_cast<size_t>(k)] = c[0]; col_out[3 * static_cast<size_t>(k) + 1] = c[1]; col_out[3 * static_cast<size_t>(k) + 2] = c[2];
  }
}

}  // namespace ovo

// =============================================================================================== handle
// Control block of one association: uploaded with ONE pinned H2D copy (frame constants, geometry init, counters)
// and read back with ONE D2H copy (counters + per-mask rows, which follow the header in the same allocation).
struct CtlHeader {
  ovo::FrameDev frame;
  ovo::FrameGeom geom;
  int32_t counters[4];  // [0] match-list length, [1] n_matched, [2] next_ins_id
};

struct ovo_map {
  uint8_t* ctl = nullptr; uint8_t* h_ctl = nullptr;
  static constexpr int kSlots = 64;
  ovo::FrameGeom* geom = nullptr;
  ovo::FrameDev* frame = nullptr;
  int32_t* counters = nullptr;  // [0] list len, [1] n_matched, [2] next_ins_id
  int32_t* votes = nullptr; size_t votes_cap = 0;
  int32_t* area = nullptr; int32_t* mask_ins = nullptr; ovo_vote_row* rows = nullptr; int masks_cap = 0;
  int2* scratch_list = nullptr; size_t scratch_cap = 0;
  float* depth_f = nullptr; size_t depth_cap = 0;
  int16_t* seg_dense = nullptr; size_t seg_dense_cap = 0;
  uint8_t* mapped = nullptr; size_t mapped_cap = 0; int32_t* pix_flags = nullptr; size_t pix_cap = 0;   // map producer scratch   // [F][N] scratch of ovo_map_fuse_dense_batch
  int2* slot_list[kSlots] = {}; size_t slot_cap[kSlots] = {}; int slot_n[kSlots] = {};
  __nv_bfloat16* text_bf16 = nullptr; size_t text_cap = 0;
  // pinned host staging
  ovo_vote_row* h_rows = nullptr; int32_t* h_counters = nullptr; ovo::FrameDev* h_frame = nullptr;
  // batched association (ovo_map_associate_batch / ovo_map_batch_*): control block [n_matched[Fcap] | next_ins_id,pad | frames[Fcap] |
  // rows[Fcap][bt_stride]] on the device and mirrored in pinned host memory, per-keyframe areas / mask->instance tables, filtered
  // depth maps, the vote tables and the dense rows seg_of_pt [F][N] (shared with ovo_map_fuse_dense_batch)
  uint8_t* bctl = nullptr; uint8_t* h_bctl = nullptr; int bt_fcap = 0, bt_stride = 0;
  int32_t* bt_area = nullptr; int32_t* bt_mask_ins = nullptr;
  float* bt_depth = nullptr; size_t bt_depth_cap = 0;
  int32_t* bt_tables = nullptr; size_t bt_tables_cap = 0;
  bool bt_last_applied = false;   // the persistent vote kernel already gave the last keyframe's points their ids
  bool bt_valid = false; int bt_F = 0; int64_t bt_N = 0; int bt_next = 0; int32_t* bt_user_tables = nullptr;
  std::vector<int> bt_n_masks, bt_track_th, bt_votes_off, bt_table_len, bt_slots;
  int64_t dense_N = 0, dense_stride = 0; int dense_F = 0; std::vector<int> dense_slots;   // what seg_dense currently holds (from a batched association)
  __nv_bfloat16* feats_bf16 = nullptr; size_t feats_cap = 0;
  // association split in two calls (ovo_map_vote / ovo_map_apply): state of the pending keyframe
  bool pend_launched = false; cudaEvent_t pend_event = nullptr; cudaStream_t pend_stream = nullptr;
  bool pend_valid = false; int64_t pend_N = 0; int pend_n_ins = 0, pend_n_masks = 0, pend_slot = 0, pend_track_th = 0;
};

namespace {
template <typename T>
int grow(T** p, size_t* cap, size_t need) {
  if (need <= *cap) return OVO_OK;
  if (*p) cudaFree(*p);
  *p = nullptr;
  const size_t n = need + need / 4 + 64;
  if (cudaMalloc(reinterpret_cast<void**>(p), n * sizeof(T)) != cudaSuccess) {
    *cap = 0;
    cudaGetLastError();
    return ovo::set_error(OVO_E_NOMEM, "workspace allocation of %zu bytes failed", n * sizeof(T));
  }
  *cap = n;
  return OVO_OK;
}
}  // namespace

static int ctl_alloc(ovo_map* m, int masks_cap) {
  cudaFree(m->ctl); cudaFreeHost(m->h_ctl); cudaFree(m->area);   // (mask_ins points INTO the area allocation: not freed on its own)
  m->ctl = nullptr; m->h_ctl = nullptr; m->area = nullptr; m->mask_ins = nullptr;
  const size_t bytes = sizeof(CtlHeader) + static_cast<size_t>(masks_cap) * sizeof(ovo_vote_row);
  if (cudaMalloc(&m->ctl, bytes) != cudaSuccess || cudaMallocHost(&m->h_ctl, bytes) != cudaSuccess ||
      cudaMalloc(&m->area, 2 * static_cast<size_t>(masks_cap) * sizeof(int32_t)) != cudaSuccess)
    return ovo::set_error(OVO_E_CUDA, "ovo_map: control block allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
  m->mask_ins = m->area + masks_cap;
  m->masks_cap = masks_cap;
  CtlHeader* h = reinterpret_cast<CtlHeader*>(m->ctl);
  m->frame = &h->frame; m->geom = &h->geom; m->counters = h->counters;
  m->rows = reinterpret_cast<ovo_vote_row*>(m->ctl + sizeof(CtlHeader));
  CtlHeader* hh = reinterpret_cast<CtlHeader*>(m->h_ctl);
  m->h_frame = &hh->frame; m->h_counters = hh->counters;
  m->h_rows = reinterpret_cast<ovo_vote_row*>(m->h_ctl + sizeof(CtlHeader));
  return OVO_OK;
}

extern "C" {

int ovo_map_create(ovo_map_t** out) {
  OVO_REQUIRE(out != nullptr, "ovo_map_create: null out");
  ovo::keep_default_mempool_cached();
  ovo_map* m = new ovo_map();
  if (ctl_alloc(m, 256) != OVO_OK) {
    ovo_map_destroy(m);
    return OVO_E_CUDA;
  }
  *out = m;
  return OVO_OK;
}

int ovo_map_reserve(ovo_map_t* m, int64_t max_points, int max_instances, int max_masks, int64_t max_matches) {
  OVO_REQUIRE(m != nullptr, "ovo_map_reserve: null handle");
  OVO_REQUIRE(!m->pend_valid, "ovo_map_reserve: an association is pending (call before ovo_map_vote / after ovo_map_apply)");
  if (max_masks > 0 && max_masks + 1 > m->masks_cap) OVO_TRY(ctl_alloc(m, max_masks + 64));
  if (max_masks > 0 && max_instances > 0) {
    const size_t need = static_cast<size_t>(max_masks) * (static_cast<size_t>(max_instances) + 1);
    OVO_REQUIRE(need < (1ull << 28), "ovo_map_reserve: vote table too large (%d masks x %d instances)", max_masks, max_instances);
    OVO_TRY(grow(&m->votes, &m->votes_cap, need));
  }
  if (max_points > 0) OVO_TRY(grow(&m->scratch_list, &m->scratch_cap, static_cast<size_t>(max_points)));
  if (max_matches > 0)
    for (int i = 0; i < ovo_map::kSlots; ++i)
      if (static_cast<size_t>(max_matches) > m->slot_cap[i]) {
        OVO_TRY(grow(&m->slot_list[i], &m->slot_cap[i], static_cast<size_t>(max_matches)));
        m->slot_n[i] = 0;   // a re-allocated slot no longer holds its keyframe's match list
      }
  return OVO_OK;
}

void ovo_map_destroy(ovo_map_t* m) {
  if (!m) return;
  cudaFree(m->ctl); cudaFreeHost(m->h_ctl); cudaFree(m->votes); cudaFree(m->area);
  cudaFree(m->scratch_list); cudaFree(m->depth_f); cudaFree(m->seg_dense); cudaFree(m->mapped); cudaFree(m->pix_flags); cudaFree(m->text_bf16);
  cudaFree(m->bctl); cudaFreeHost(m->h_bctl); cudaFree(m->bt_area); cudaFree(m->bt_depth); cudaFree(m->bt_tables); cudaFree(m->feats_bf16);
  for (int i = 0; i < ovo_map::kSlots; ++i) cudaFree(m->slot_list[i]);
  if (m->pend_event) cudaEventDestroy(m->pend_event);
  delete m;
}

int ovo_depth_range(const float* depth_dev, int64_t n, float* range_out_dev, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  OVO_REQUIRE(depth_dev && range_out_dev && n > 0 && n < (1LL << 31), "ovo_depth_range: bad arguments");
  ovo::depth_range_init_kernel<<<1, 1, 0, stream>>>(reinterpret_cast<int*>(range_out_dev));
  OVO_CHECK_LAUNCH();
  ovo::depth_range_kernel<<<32, 256, 0, stream>>>(depth_dev, static_cast<int>(n), reinterpret_cast<int*>(range_out_dev));
  OVO_CHECK_LAUNCH();
  return OVO_OK;
}

int ovo_depth_filter_batch(const float* depth_dev, int n_frames, int h, int w, float* out_dev, float* ranges_out_dev, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  OVO_REQUIRE(depth_dev && (out_dev || ranges_out_dev) && n_frames > 0 && h >= 4 && w >= 4, "ovo_depth_filter_batch: bad arguments");
  if (ranges_out_dev) {
    ovo::depth_range_init_batch_kernel<<<ovo::ceil_div(n_frames, 128), 128, 0, stream>>>(reinterpret_cast<int*>(ranges_out_dev), n_frames);
    OVO_CHECK_LAUNCH();
  }
  ovo::depth_filter_range_batch_kernel<<<dim3(ovo::ceil_div(w, 32), ovo::ceil_div(h, 4), n_frames), dim3(32, 4), 0, stream>>>(
      depth_dev, h, w, out_dev, reinterpret_cast<int*>(ranges_out_dev));
  OVO_CHECK_LAUNCH();
  return OVO_OK;
}

int ovo_depth_filter(const float* depth_dev, int h, int w, float* out_dev, void* stream) {
  OVO_REQUIRE(depth_dev && out_dev && h >= 4 && w >= 4, "ovo_depth_filter: bad arguments");
  dim3 b(32, 8), g(ovo::ceil_div(w, 32), ovo::ceil_div(h, 8));
  ovo::depth_filter_kernel<<<g, b, 0, static_cast<cudaStream_t>(stream)>>>(depth_dev, h, w, out_dev);
  OVO_CHECK_LAUNCH();
  return OVO_OK;
}

// Phase 1 of the association: frustum, depth filter, mask areas, pass 1 (votes + match list).  No host sync.
static int associate_vote(ovo_map_t* m, const float* xyz_dev, const int32_t* ins_ids_dev, int64_t N, const ovo_frame* f,
                          int n_ins_in, int kf_slot, cudaStream_t stream) {
  OVO_REQUIRE(m && f, "ovo_map_associate: null argument");
  const int next_ins_id_v = n_ins_in;
  const int* next_ins_id = &next_ins_id_v;
  OVO_REQUIRE(N >= 0 && N < (1LL << 31), "ovo_map_associate: N out of range");
  OVO_REQUIRE(kf_slot >= 0 && kf_slot < ovo_map::kSlots, "ovo_map_associate: kf_slot out of range");
  OVO_REQUIRE(f->n_masks >= 0 && f->n_masks <= 8192, "ovo_map_associate: n_masks out of range");
  OVO_REQUIRE(f->depth_dev && f->seg_map_dev && f->h > 0 && f->w > 0 && f->H > 0 && f->W > 0, "ovo_map_associate: bad frame");
  OVO_REQUIRE(f->depth_range_dev == nullptr, "ovo_map_associate: pre-filtered depth (depth_range_dev) is a feature of the batched calls");
  OVO_REQUIRE(N == 0 || (xyz_dev && ins_ids_dev), "ovo_map_associate: null map");
  const int n_masks = f->n_masks, n_ins = *next_ins_id;
  OVO_REQUIRE(n_ins >= 0, "ovo_map_associate: negative next_ins_id");
  const size_t votes_need = static_cast<size_t>(n_masks > 0 ? n_masks : 1) * (n_ins + 1);
  OVO_REQUIRE(votes_need < (1ull << 28), "ovo_map_associate: vote table too large (%d masks x %d instances)", n_masks, n_ins);

  // workspaces
  OVO_TRY(grow(&m->votes, &m->votes_cap, votes_need));
  if (n_masks + 1 > m->masks_cap) OVO_TRY(ctl_alloc(m, n_masks + 64));
  OVO_TRY(grow(&m->scratch_list, &m->scratch_cap, static_cast<size_t>(N > 0 ? N : 1)));
  const int npix = f->h * f->w;
  OVO_TRY(grow(&m->depth_f, &m->depth_cap, static_cast<size_t>(npix)));

  // frame constants
  ovo::FrameDev* hf = m->h_frame;
  memcpy(hf->c2w, f->c2w, sizeof(hf->c2w)); memcpy(hf->w2c, f->w2c, sizeof(hf->w2c)); memcpy(hf->K, f->K, sizeof(hf->K));
  hf->match_th = f->match_th; hf->h = f->h; hf->w = f->w; hf->H = f->H; hf->W = f->W;
  hf->has_ratio = f->has_ratio; hf->ratio_h = f->ratio_h; hf->ratio_w = f->ratio_w; hf->crop_edge = f->crop_edge;
  CtlHeader* hh = reinterpret_cast<CtlHeader*>(m->h_ctl);
  memset(&hh->geom, 0, sizeof(hh->geom));
  hh->geom.dmin_bits = 0x7f800000; hh->geom.dmax_bits = 0;
  m->h_counters[0] = 0; m->h_counters[1] = 0; m->h_counters[2] = n_ins; m->h_counters[3] = 0;
  OVO_CUDA(cudaMemcpyAsync(m->ctl, m->h_ctl, sizeof(CtlHeader), cudaMemcpyHostToDevice, stream));   // one pinned upload
  OVO_CUDA(cudaMemsetAsync(m->votes, 0, votes_need * sizeof(int32_t), stream));
  OVO_CUDA(cudaMemsetAsync(m->area, 0, static_cast<size_t>(n_masks + 1) * sizeof(int32_t), stream));

  const int sms = ovo::num_sms();
  // frustum from the RAW depth (ovo.py:209), match against the filtered depth (ovo.py:213-216)
  ovo::depth_minmax_kernel<<<sms, 256, 0, stream>>>(f->depth_dev, npix, m->geom);
  OVO_CHECK_LAUNCH();
  ovo::frustum_setup_kernel<<<1, 1, 0, stream>>>(m->geom, m->frame);
  OVO_CHECK_LAUNCH();
  const float* depth_used = f->depth_dev;
  if (f->depth_filter) {
    OVO_TRY(ovo_depth_filter(f->depth_dev, f->h, f->w, m->depth_f, stream));
    depth_used = m->depth_f;
  }
  if (n_masks > 0) {
    ovo::seg_area_kernel<<<sms, 256, n_masks * sizeof(int32_t), stream>>>(f->seg_map_dev, f->H * f->W, n_masks, m->area);
    OVO_CHECK_LAUNCH();
  }
  if (N > 0) {
    const int blocks = static_cast<int>(std::min<long long>((N + ovo::kP1Threads - 1) / ovo::kP1Threads, sms * 8LL));
    const size_t vbytes = votes_need * sizeof(int32_t);
    const int smem_votes = vbytes <= 32 * 1024 ? 1 : 0;
    ovo::associate_pass1_kernel<<<blocks, ovo::kP1Threads, smem_votes ? vbytes : 0, stream>>>(
        xyz_dev, ins_ids_dev, N, depth_used, f->seg_map_dev, m->geom, m->frame, n_masks, n_ins, m->votes, m->scratch_list,
        m->counters, smem_votes);
    OVO_CHECK_LAUNCH();
  }
  m->pend_N = N; m->pend_n_ins = n_ins; m->pend_n_masks = n_masks; m->pend_slot = kf_slot; m->pend_track_th = f->track_th;
  m->pend_valid = true;
  return OVO_OK;
}

// Phase 2: per-mask reduce of the (possibly all-reduced) vote table, id decisions, pass 2 and the read-back ENQUEUED (no host
// sync): associate_finish waits for it.
static int associate_apply_launch(ovo_map_t* m, int32_t* ins_ids_dev, cudaStream_t stream) {
  if (!m->pend_valid) return ovo::set_error(OVO_E_STATE, "ovo_map_apply called without a pending ovo_map_vote");
  const int64_t N = m->pend_N;
  const int n_masks = m->pend_n_masks, n_ins = m->pend_n_ins, track_th = m->pend_track_th;
  const int sms = ovo::num_sms();
  if (n_masks > 0) {
    ovo::vote_reduce_kernel<<<n_masks, 128, 0, stream>>>(m->votes, n_ins, m->area, m->rows);
    OVO_CHECK_LAUNCH();
    ovo::vote_decide_kernel<<<1, 256, 0, stream>>>(m->rows, n_masks, track_th, m->mask_ins, m->counters + 2);
    OVO_CHECK_LAUNCH();
    if (N > 0) {
      ovo::associate_pass2_kernel<<<sms * 4, 256, 0, stream>>>(m->scratch_list, m->counters, m->mask_ins, ins_ids_dev);
      OVO_CHECK_LAUNCH();
    }
  }
  OVO_CUDA(cudaMemcpyAsync(m->h_counters, m->counters, 4 * sizeof(int32_t) + static_cast<size_t>(n_masks) * sizeof(ovo_vote_row),
                           cudaMemcpyDeviceToHost, stream));   // counters + rows are contiguous: one read-back
  if (!m->pend_event) OVO_CUDA(cudaEventCreateWithFlags(&m->pend_event, cudaEventDisableTiming));
  OVO_CUDA(cudaEventRecord(m->pend_event, stream));
  m->pend_stream = stream;
  m->pend_launched = true;
  return OVO_OK;
}

// The one host synchronisation of an association: waits for the read-back, hands the rows to the caller, keeps the match list.
static int associate_finish(ovo_map_t* m, int* next_ins_id, ovo_vote_row* votes_host, int* n_matched_host) {
  OVO_REQUIRE(m && next_ins_id && votes_host && n_matched_host, "ovo_map_associate: null argument");
  if (!m->pend_valid || !m->pend_launched) return ovo::set_error(OVO_E_STATE, "no association is in flight on this handle");
  m->pend_valid = false; m->pend_launched = false;
  cudaStream_t stream = m->pend_stream;
  const int n_masks = m->pend_n_masks, kf_slot = m->pend_slot;
  OVO_CUDA(cudaEventSynchronize(m->pend_event));
  if (n_masks > 0) memcpy(votes_host, m->h_rows, n_masks * sizeof(ovo_vote_row));
  *n_matched_host = m->h_counters[1];
  *next_ins_id = m->h_counters[2];
  // keep this keyframe's matched list for the (delayed) dense fusion
  const int n_list = m->h_counters[0];
  OVO_TRY(grow(&m->slot_list[kf_slot], &m->slot_cap[kf_slot], static_cast<size_t>(n_list > 0 ? n_list : 1)));
  if (n_list > 0)
    OVO_CUDA(cudaMemcpyAsync(m->slot_list[kf_slot], m->scratch_list, n_list * sizeof(int2), cudaMemcpyDeviceToDevice, stream));
  m->slot_n[kf_slot] = n_list;
  for (size_t i = 0; i < m->dense_slots.size(); ++i)
    if (m->dense_slots[i] == kf_slot) m->dense_slots[i] = -1;   // the slot's dense row (of an earlier batch) is stale now
  return OVO_OK;
}

static int associate_apply(ovo_map_t* m, int32_t* ins_ids_dev, int* next_ins_id, ovo_vote_row* votes_host,
                           int* n_matched_host, cudaStream_t stream) {
  OVO_REQUIRE(m && next_ins_id && votes_host && n_matched_host, "ovo_map_associate: null argument");
  OVO_TRY(associate_apply_launch(m, ins_ids_dev, stream));
  return associate_finish(m, next_ins_id, votes_host, n_matched_host);
}

// ovo_map_associate in two halves: everything enqueued, nothing waited for (the caller goes on with its own work while the
// GPU runs the association) ...
int ovo_map_associate_launch(ovo_map_t* m, const float* xyz_dev, int32_t* ins_ids_dev, int64_t N, const ovo_frame* f, int next_ins_id,
                             int kf_slot, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  OVO_REQUIRE(m && !m->pend_valid, "ovo_map_associate_launch: an association is already in flight on this handle (ovo_map_associate_wait first)");
  ovo::ProfScope prof(stream, ovo::PROF_ASSOC, 0.0, static_cast<double>(N) * 20 + (f ? static_cast<double>(f->h) * f->w * 8 : 0.0));
  OVO_TRY(associate_vote(m, xyz_dev, ins_ids_dev, N, f, next_ins_id, kf_slot, stream));
  const int r = associate_apply_launch(m, ins_ids_dev, stream);
  if (r != OVO_OK) m->pend_valid = false;
  return r;
}
// ... and the one host synchronisation, whenever the caller needs the rows (votes_host [n_masks of the launch]).
int ovo_map_associate_wait(ovo_map_t* m, int* next_ins_id, ovo_vote_row* votes_host, int* n_matched_host) {
  return associate_finish(m, next_ins_id, votes_host, n_matched_host);
}


int ovo_map_associate(ovo_map_t* m, const float* xyz_dev, int32_t* ins_ids_dev, int64_t N, const ovo_frame* f,
                      int* next_ins_id, ovo_vote_row* votes_host, int* n_matched_host, int kf_slot, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  OVO_REQUIRE(next_ins_id && votes_host && n_matched_host, "ovo_map_associate: null argument");
  ovo::ProfScope prof(stream, ovo::PROF_ASSOC, 0.0, static_cast<double>(N) * 20 + (f ? static_cast<double>(f->h) * f->w * 8 : 0.0));
  OVO_TRY(associate_vote(m, xyz_dev, ins_ids_dev, N, f, *next_ins_id, kf_slot, stream));
  return associate_apply(m, ins_ids_dev, next_ins_id, votes_host, n_matched_host, stream);
}

int ovo_map_vote(ovo_map_t* m, const float* xyz_dev, const int32_t* ins_ids_dev, int64_t N, const ovo_frame* f, int n_ins,
                 int32_t* table_out_dev, int kf_slot, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  OVO_REQUIRE(table_out_dev != nullptr, "ovo_map_vote: null table");
  OVO_TRY(associate_vote(m, xyz_dev, ins_ids_dev, N, f, n_ins, kf_slot, stream));
  const size_t n = static_cast<size_t>(m->pend_n_masks > 0 ? m->pend_n_masks : 1) * (n_ins + 1);
  OVO_CUDA(cudaMemcpyAsync(table_out_dev, m->votes, n * sizeof(int32_t), cudaMemcpyDeviceToDevice, stream));
  OVO_CUDA(cudaMemcpyAsync(table_out_dev + n, m->counters + 1, sizeof(int32_t), cudaMemcpyDeviceToDevice, stream));
  return static_cast<int>(n + 1);
}

int ovo_map_apply(ovo_map_t* m, const int32_t* table_in_dev, int32_t* ins_ids_dev, int* next_ins_id,
                  ovo_vote_row* votes_host, int* n_matched_host, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  OVO_REQUIRE(m && table_in_dev, "ovo_map_apply: null argument");
  if (!m->pend_valid) return ovo::set_error(OVO_E_STATE, "ovo_map_apply called without a pending ovo_map_vote");
  const size_t n = static_cast<size_t>(m->pend_n_masks > 0 ? m->pend_n_masks : 1) * (m->pend_n_ins + 1);
  OVO_CUDA(cudaMemcpyAsync(m->votes, table_in_dev, n * sizeof(int32_t), cudaMemcpyDeviceToDevice, stream));
  OVO_CUDA(cudaMemcpyAsync(m->counters + 1, table_in_dev + n, sizeof(int32_t), cudaMemcpyDeviceToDevice, stream));
  return associate_apply(m, ins_ids_dev, next_ins_id, votes_host, n_matched_host, stream);
}

// ----------------------------------------------------------------------------------------------- batched association
namespace {
struct BatchCtl {   // byte offsets inside the batch control block
  size_t n_matched, next, frames, rows, end;
  BatchCtl(int fcap, int stride) {
    n_matched = 0;
    next = (static_cast<size_t>(fcap) * 4 + 15) & ~size_t(15);
    frames = next + 16;                                                      // 16-byte aligned: BatchFrame holds pointers
    rows = (frames + static_cast<size_t>(fcap) * sizeof(ovo::BatchFrame) + 15) & ~size_t(15);
    end = rows + static_cast<size_t>(fcap) * stride * sizeof(ovo_vote_row);
  }
};
}  // namespace

static int batch_ctl_alloc(ovo_map* m, int F, int max_masks) {
  if (F <= m->bt_fcap && max_masks <= m->bt_stride) return OVO_OK;
  const int fcap = std::max(F, m->bt_fcap), stride = std::max((max_masks + 63) / 64 * 64, m->bt_stride);
  cudaFree(m->bctl); cudaFreeHost(m->h_bctl); cudaFree(m->bt_area);
  m->bctl = nullptr; m->h_bctl = nullptr; m->bt_area = nullptr; m->bt_fcap = 0; m->bt_stride = 0;
  const BatchCtl L(fcap, stride);
  if (cudaMalloc(&m->bctl, L.end) != cudaSuccess || cudaMallocHost(&m->h_bctl, L.end) != cudaSuccess ||
      cudaMalloc(&m->bt_area, 2 * static_cast<size_t>(fcap) * stride * sizeof(int32_t)) != cudaSuccess)
    return ovo::set_error(OVO_E_CUDA, "ovo_map: batch control block allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
  m->bt_mask_ins = m->bt_area + static_cast<size_t>(fcap) * stride;
  m->bt_fcap = fcap; m->bt_stride = stride;
  return OVO_OK;
}

int ovo_map_batch_begin(ovo_map_t* m, const float* xyz_dev, const int32_t* ins_ids_dev, int64_t N, const ovo_frame* frames, int F,
                        const int* kf_slots, int next_ins_id, int32_t* tables_dev, int64_t tables_cap, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  OVO_REQUIRE(m && frames && F > 0 && F <= 64, "ovo_map_batch_begin: 1..64 keyframes per batch (got %d)", F);
  OVO_REQUIRE(N >= 0 && N < (1LL << 31) && next_ins_id >= 0, "ovo_map_batch_begin: N / next_ins_id out of range");
  OVO_REQUIRE(N == 0 || (xyz_dev && ins_ids_dev), "ovo_map_batch_begin: null map");
  OVO_REQUIRE(!m->pend_valid && !m->bt_valid, "ovo_map_batch_begin: another association is pending on this handle");
  int max_masks = 1;
  size_t npix_max = 0;
  for (int f = 0; f < F; ++f) {
    const ovo_frame& fr = frames[f];
    OVO_REQUIRE(fr.n_masks >= 0 && fr.n_masks <= 8192, "ovo_map_batch_begin: n_masks out of range (keyframe %d)", f);
    OVO_REQUIRE(fr.depth_dev && fr.seg_map_dev && fr.h > 0 && fr.w > 0 && fr.H > 0 && fr.W > 0, "ovo_map_batch_begin: bad frame %d", f);
    OVO_REQUIRE(!kf_slots || (kf_slots[f] >= 0 && kf_slots[f] < ovo_map::kSlots), "ovo_map_batch_begin: kf_slot out of range");
    max_masks = std::max(max_masks, fr.n_masks);
    npix_max = std::max(npix_max, static_cast<size_t>(fr.h) * fr.w);
  }
  OVO_TRY(batch_ctl_alloc(m, F, max_masks));
  OVO_TRY(grow(&m->bt_depth, &m->bt_depth_cap, static_cast<size_t>(F) * npix_max));
  const int64_t stride = ((N > 0 ? N : 1) + 7) & ~int64_t(7);   // rows padded to 8 points: 16-byte loads in the vote scan
  OVO_TRY(grow(&m->seg_dense, &m->seg_dense_cap, static_cast<size_t>(F) * static_cast<size_t>(stride)));
  // vote tables: keyframe f may see up to next_ins_id + (masks of the keyframes before it) instances; each table starts with a
  // 4-int header whose first int counts the keyframe's matched points, so that a sharded map sums both with one exchange (and
  // only the compact front of the table, 4 + n_masks x (n_ins+1) ints, has to travel)
  m->bt_n_masks.assign(F, 0); m->bt_track_th.assign(F, 0); m->bt_votes_off.assign(F, 0); m->bt_table_len.assign(F, 0);
  m->bt_slots.assign(F, -1);
  size_t off = 0;
  int bound = next_ins_id;
  for (int f = 0; f < F; ++f) {
    const int nm = frames[f].n_masks;
    const size_t len = 4 + static_cast<size_t>(nm > 0 ? nm : 1) * (bound + 1);
    OVO_REQUIRE(len < (1ull << 28), "ovo_map_batch_begin: vote table too large (%d masks x %d instances)", nm, bound);
    m->bt_n_masks[f] = nm; m->bt_track_th[f] = frames[f].track_th; m->bt_votes_off[f] = static_cast<int>(off);
    m->bt_table_len[f] = static_cast<int>(len);
    if (kf_slots) m->bt_slots[f] = kf_slots[f];
    off += (len + 3) & ~size_t(3);
    OVO_REQUIRE(off < (1ull << 30), "ovo_map_batch_begin: vote tables of the batch too large");
    bound += nm;
  }
  int32_t* tables = tables_dev;
  if (tables_dev) {
    OVO_REQUIRE(static_cast<size_t>(tables_cap) >= off, "ovo_map_batch_begin: tables buffer too small (%lld < %zu ints)", (long long)tables_cap, off);
  } else {
    OVO_TRY(grow(&m->bt_tables, &m->bt_tables_cap, off));
    tables = m->bt_tables;
  }
  m->bt_user_tables = tables;

  const BatchCtl L(m->bt_fcap, m->bt_stride);
  int32_t* h_nm = reinterpret_cast<int32_t*>(m->h_bctl + L.n_matched);
  int32_t* h_next = reinterpret_cast<int32_t*>(m->h_bctl + L.next);
  ovo::BatchFrame* hf = reinterpret_cast<ovo::BatchFrame*>(m->h_bctl + L.frames);
  for (int f = 0; f < F; ++f) {
    const ovo_frame& fr = frames[f];
    ovo::BatchFrame& b = hf[f];
    memset(&b, 0, sizeof(b));
    memcpy(b.fr.c2w, fr.c2w, sizeof(b.fr.c2w)); memcpy(b.fr.w2c, fr.w2c, sizeof(b.fr.w2c)); memcpy(b.fr.K, fr.K, sizeof(b.fr.K));
    b.fr.match_th = fr.match_th; b.fr.h = fr.h; b.fr.w = fr.w; b.fr.H = fr.H; b.fr.W = fr.W;
    b.fr.has_ratio = fr.has_ratio; b.fr.ratio_h = fr.ratio_h; b.fr.ratio_w = fr.ratio_w; b.fr.crop_edge = fr.crop_edge;
    b.geom.dmin_bits = 0x7f800000; b.geom.dmax_bits = 0;
    b.depth_raw = fr.depth_dev;
    b.range = fr.depth_range_dev;
    const bool filter_here = fr.depth_filter && fr.depth_range_dev == nullptr;
    b.depth_filtered = filter_here ? m->bt_depth + static_cast<size_t>(f) * npix_max : nullptr;
    b.depth_used = filter_here ? b.depth_filtered : fr.depth_dev;
    b.seg_map = fr.seg_map_dev; b.n_masks = fr.n_masks; b.track_th = fr.track_th; b.votes_off = m->bt_votes_off[f];
    b.nm_off = m->bt_votes_off[f];
    h_nm[f] = 0;
  }
  h_next[0] = next_ins_id; h_next[1] = h_next[2] = h_next[3] = 0;
  OVO_CUDA(cudaMemcpyAsync(m->bctl, m->h_bctl, L.frames + static_cast<size_t>(F) * sizeof(ovo::BatchFrame), cudaMemcpyHostToDevice, stream));
  OVO_CUDA(cudaMemsetAsync(tables, 0, off * sizeof(int32_t), stream));
  OVO_CUDA(cudaMemsetAsync(m->bt_area, 0, static_cast<size_t>(m->bt_fcap) * m->bt_stride * sizeof(int32_t), stream));
  ovo::BatchFrame* dfr = reinterpret_cast<ovo::BatchFrame*>(m->bctl + L.frames);
  const int sms = ovo::num_sms();
  int hmax = 0, wmax = 0;
  for (int f = 0; f < F; ++f) { hmax = std::max(hmax, frames[f].h); wmax = std::max(wmax, frames[f].w); }
  ovo::batch_depth_minmax_kernel<<<dim3(32, F), 256, 0, stream>>>(dfr);
  OVO_CHECK_LAUNCH();
  ovo::batch_frustum_setup_kernel<<<F, 1, 0, stream>>>(dfr);
  OVO_CHECK_LAUNCH();
  ovo::batch_depth_filter_kernel<<<dim3(ovo::ceil_div(wmax, 32), ovo::ceil_div(hmax, 4), F), dim3(32, 4), 0, stream>>>(dfr);   // 128-thread blocks: co-resident with the encoder's GEMM CTAs
  OVO_CHECK_LAUNCH();
  ovo::batch_seg_area_kernel<<<dim3(32, F), 256, max_masks * sizeof(int32_t), stream>>>(dfr, m->bt_area, m->bt_stride);
  OVO_CHECK_LAUNCH();
  if (N > 0) {
    const int blocks = static_cast<int>(std::min<long long>((N + ovo::kP1Threads - 1) / ovo::kP1Threads, sms * 64LL));   // short-lived blocks
    for (int f0 = 0; f0 < F; f0 += ovo::kBatchFramesPerPass) {
      ovo::associate_batch_pass_kernel<<<blocks, ovo::kP1Threads, 0, stream>>>(xyz_dev, N, dfr, f0, std::min(ovo::kBatchFramesPerPass, F - f0),
                                                                                m->seg_dense, stride, tables);
      OVO_CHECK_LAUNCH();
    }
  }
  m->bt_valid = true; m->bt_F = F; m->bt_N = N; m->bt_next = next_ins_id; m->bt_last_applied = false;
  m->dense_slots.assign(m->bt_slots.begin(), m->bt_slots.end()); m->dense_N = N; m->dense_F = F; m->dense_stride = stride;
  for (int f = 0; f < F; ++f)
    if (m->bt_slots[f] >= 0) m->slot_n[m->bt_slots[f]] = 0;   // these slots hold dense rows now, not lists
  return OVO_OK;
}

static ovo::VoteTail make_tail(ovo_map* m, int f, ovo_xchg* x) {
  const BatchCtl L(m->bt_fcap, m->bt_stride);
  ovo::VoteTail t;
  memset(&t, 0, sizeof(t));
  t.enabled = 1;
  t.done = reinterpret_cast<int32_t*>(m->bctl + L.next) + 2;     // zeroed by the upload of ovo_map_batch_begin, reset by every tail
  t.table = m->bt_user_tables + m->bt_votes_off[f];
  t.area = m->bt_area + static_cast<size_t>(f) * m->bt_stride;
  t.rows = reinterpret_cast<ovo_vote_row*>(m->bctl + L.rows) + static_cast<size_t>(f) * m->bt_stride;
  t.mask_ins = m->bt_mask_ins + static_cast<size_t>(f) * m->bt_stride;
  t.next_ins_id = reinterpret_cast<int32_t*>(m->bctl + L.next);
  t.n_matched_out = reinterpret_cast<int32_t*>(m->bctl + L.n_matched) + f;
  t.track_th = m->bt_track_th[f];
  t.world = 1;
  if (x != nullptr && x->world > 1) {
    const int epoch = ++(*x->epochs)[f];
    t.world = x->world; t.rank = x->rank; t.slots = x->slots; t.slot = f; t.parity = epoch & 1; t.epoch = epoch;
    t.table_cap = x->table_cap; t.peers = x->peers;
  }
  return t;
}

// scan of keyframe f (ids of keyframe f-1 first, then its votes); tail != nullptr: the fused exchange + decisions follow in the
// same launch
static int batch_scan(ovo_map_t* m, int f, int32_t* ins_ids_dev, const ovo::VoteTail* tail, cudaStream_t stream) {
  const BatchCtl L(m->bt_fcap, m->bt_stride);
  const int64_t N = m->bt_N;
  int32_t* table = m->bt_user_tables + m->bt_votes_off[f];
  const int len = m->bt_table_len[f];
  const int32_t* d_next = reinterpret_cast<const int32_t*>(m->bctl + L.next);
  ovo::VoteTail none;
  memset(&none, 0, sizeof(none));
  if (N > 0) {
    const size_t st = static_cast<size_t>(m->dense_stride);
    const int16_t* prev = f > 0 ? m->seg_dense + static_cast<size_t>(f - 1) * st : nullptr;
    const int32_t* mi_prev = f > 0 ? m->bt_mask_ins + static_cast<size_t>(f - 1) * m->bt_stride : nullptr;
    const int smem_ints = 0;
    (void)len;
    const int blocks = static_cast<int>(std::min<long long>(((N + 7) / 8 + 255) / 256, ovo::num_sms() * 8LL));
    ovo::batch_vote_scan_kernel<<<blocks, 256, 0, stream>>>(prev, mi_prev, m->seg_dense + static_cast<size_t>(f) * st,
                                                           ins_ids_dev, N, d_next, m->bt_n_masks[f], table + 4, smem_ints,
                                                           tail ? *tail : none);
    OVO_CHECK_LAUNCH();
  } else if (tail) {   // an empty shard still takes part in the exchange and takes the decisions
    ovo::batch_vote_finish_kernel<<<1, 256, 0, stream>>>(*tail, m->bt_n_masks[f]);
    OVO_CHECK_LAUNCH();
  }
  return OVO_OK;
}

// votes of keyframe f on this handle's points (after applying the decisions of keyframe f-1); *table_dev / *table_len = the
// keyframe's table [n_matched, 0, 0, 0 | n_masks x (n_ins+1) votes] inside the batch's table buffer, to be summed over shards
int ovo_map_batch_vote(ovo_map_t* m, int f, int32_t* ins_ids_dev, int32_t** table_dev, int* table_len, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  OVO_REQUIRE(m && m->bt_valid && f >= 0 && f < m->bt_F, "ovo_map_batch_vote: no batch pending or keyframe %d out of range", f);
  OVO_TRY(batch_scan(m, f, ins_ids_dev, nullptr, stream));
  if (table_dev) *table_dev = m->bt_user_tables + m->bt_votes_off[f];
  if (table_len) *table_len = m->bt_table_len[f];
  return OVO_OK;
}

// decisions of keyframe f from its (summed) table: per-mask reduce, ordered id allocation; next_ins_id stays on the device
int ovo_map_batch_decide(ovo_map_t* m, int f, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  OVO_REQUIRE(m && m->bt_valid && f >= 0 && f < m->bt_F, "ovo_map_batch_decide: no batch pending or keyframe %d out of range", f);
  const ovo::VoteTail t = make_tail(m, f, nullptr);
  ovo::batch_vote_finish_kernel<<<1, 256, 0, stream>>>(t, m->bt_n_masks[f]);
  OVO_CHECK_LAUNCH();
  return OVO_OK;
}

// ids of the last keyframe's matches, then ONE read-back of every keyframe's rows and ONE host synchronisation
int ovo_map_batch_end(ovo_map_t* m, int32_t* ins_ids_dev, int* next_ins_id, ovo_vote_row* votes_host, int votes_stride,
                      int* n_matched_host, int32_t* mask_ins_out_dev, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  OVO_REQUIRE(m && m->bt_valid, "ovo_map_batch_end: no batch pending");
  OVO_REQUIRE(next_ins_id && votes_host && n_matched_host && votes_stride > 0, "ovo_map_batch_end: null argument");
  m->bt_valid = false;
  const BatchCtl L(m->bt_fcap, m->bt_stride);
  const int F = m->bt_F;
  const int64_t N = m->bt_N;
  for (int f = 0; f < F; ++f) OVO_REQUIRE(m->bt_n_masks[f] <= votes_stride, "ovo_map_batch_end: votes_stride %d < n_masks %d", votes_stride, m->bt_n_masks[f]);
  if (N > 0 && m->bt_n_masks[F - 1] > 0 && !m->bt_last_applied) {
    const int blocks = static_cast<int>(std::min<long long>(((N + 7) / 8 + 255) / 256, ovo::num_sms() * 8LL));
    ovo::batch_vote_scan_kernel<<<blocks, 256, 0, stream>>>(m->seg_dense + static_cast<size_t>(F - 1) * static_cast<size_t>(m->dense_stride),
                                                           m->bt_mask_ins + static_cast<size_t>(F - 1) * m->bt_stride, nullptr, ins_ids_dev, N,
                                                           reinterpret_cast<const int32_t*>(m->bctl + L.next), 0, nullptr, 0, ovo::VoteTail{});
    OVO_CHECK_LAUNCH();
  }
  if (mask_ins_out_dev)
    for (int f = 0; f < F; ++f) {
      if (m->bt_n_masks[f] < votes_stride)
        OVO_CUDA(cudaMemsetAsync(mask_ins_out_dev + static_cast<size_t>(f) * votes_stride, 0xff, static_cast<size_t>(votes_stride) * 4, stream));
      if (m->bt_n_masks[f] > 0)
        OVO_CUDA(cudaMemcpyAsync(mask_ins_out_dev + static_cast<size_t>(f) * votes_stride, m->bt_mask_ins + static_cast<size_t>(f) * m->bt_stride,
                                 static_cast<size_t>(m->bt_n_masks[f]) * 4, cudaMemcpyDeviceToDevice, stream));
    }
  const size_t bytes = L.rows + static_cast<size_t>(F) * m->bt_stride * sizeof(ovo_vote_row);
  OVO_CUDA(cudaMemcpyAsync(m->h_bctl, m->bctl, bytes, cudaMemcpyDeviceToHost, stream));
  OVO_CUDA(cudaStreamSynchronize(stream));   // the one host sync of the batch
  const int32_t* h_nm = reinterpret_cast<const int32_t*>(m->h_bctl + L.n_matched);
  const ovo_vote_row* h_rows = reinterpret_cast<const ovo_vote_row*>(m->h_bctl + L.rows);
  for (int f = 0; f < F; ++f) {
    n_matched_host[f] = h_nm[f];
    if (m->bt_n_masks[f] > 0)
      memcpy(votes_host + static_cast<size_t>(f) * votes_stride, h_rows + static_cast<size_t>(f) * m->bt_stride, m->bt_n_masks[f] * sizeof(ovo_vote_row));
  }
  *next_ins_id = *reinterpret_cast<const int32_t*>(m->h_bctl + L.next);
  return OVO_OK;
}

int ovo_map_batch_info(ovo_map_t* m, int f, int* n_masks, const int32_t** n_ins_dev) {
  OVO_REQUIRE(m && m->bt_valid && f >= 0 && f < m->bt_F, "ovo_map_batch_info: no batch pending or keyframe %d out of range", f);
  const BatchCtl L(m->bt_fcap, m->bt_stride);
  if (n_masks) *n_masks = m->bt_n_masks[f];
  if (n_ins_dev) *n_ins_dev = reinterpret_cast<const int32_t*>(m->bctl + L.next);
  return OVO_OK;
}

// all keyframes of the batch, ONE launch each: scan + (device-side exchange over the shards) + decisions
static int batch_run_fused(ovo_map_t* m, ovo_xchg* xchg, int32_t* ins_ids_dev, cudaStream_t stream) {
  for (int f = 0; f < m->bt_F; ++f) {
    const ovo::VoteTail t = make_tail(m, f, xchg);
    const int r = batch_scan(m, f, ins_ids_dev, &t, stream);
    if (r != OVO_OK) { m->bt_valid = false; return r; }
  }
  return OVO_OK;
}

// all keyframes of the batch in ONE persistent launch (batch_vote_persistent_kernel); the ids of the last keyframe included
static int batch_run_persistent(ovo_map_t* m, ovo_xchg* xchg, int32_t* ins_ids_dev, cudaStream_t stream) {
  const BatchCtl L(m->bt_fcap, m->bt_stride);
  ovo::VotePersist P;
  memset(&P, 0, sizeof(P));
  P.frames = reinterpret_cast<const ovo::BatchFrame*>(m->bctl + L.frames);
  P.F = m->bt_F;
  P.seg = m->seg_dense; P.stride = m->dense_stride;
  P.ins_ids = ins_ids_dev; P.N = m->bt_N;
  P.tables = m->bt_user_tables;
  P.area = m->bt_area; P.rows = reinterpret_cast<ovo_vote_row*>(m->bctl + L.rows); P.mask_ins = m->bt_mask_ins; P.rows_stride = m->bt_stride;
  P.next_ins_id = reinterpret_cast<int32_t*>(m->bctl + L.next);
  P.n_matched_out = reinterpret_cast<int32_t*>(m->bctl + L.n_matched);
  P.sync = reinterpret_cast<int32_t*>(m->bctl + L.next) + 2;          // [2], [3] of the 4-int state: zeroed by the upload of batch_begin
  P.world = 1;
  if (xchg != nullptr && xchg->world > 1) {
    P.world = xchg->world; P.rank = xchg->rank; P.slots = xchg->slots; P.table_cap = xchg->table_cap; P.peers = xchg->peers;
    for (int f = 0; f < m->bt_F; ++f) P.epoch[f] = ++(*xchg->epochs)[f];
  }
  const long long n8 = (m->bt_N + 7) / 8;
  const int blocks = static_cast<int>(std::max<long long>(1, std::min<long long>((n8 + 255) / 256, ovo::num_sms())));
  if (P.world > 1) ovo::batch_vote_persistent_kernel<false><<<blocks, 256, 0, stream>>>(P);
  else ovo::batch_vote_persistent_kernel<true><<<blocks, 256, 0, stream>>>(P);
  OVO_CHECK_LAUNCH();
  m->bt_last_applied = true;
  return OVO_OK;
}

int ovo_map_associate_batch_sharded(ovo_map_t* m, ovo_xchg_t* xchg, const float* xyz_dev, int32_t* ins_ids_dev, int64_t N,
                                    const ovo_frame* frames, int F, const int* kf_slots, int* next_ins_id, ovo_vote_row* votes_host,
                                    int votes_stride, int* n_matched_host, int32_t* mask_ins_out_dev, void* stream_) {
  OVO_REQUIRE(xchg && next_ins_id && votes_host && n_matched_host, "ovo_map_associate_batch_sharded: null argument");
  OVO_REQUIRE(F <= xchg->slots, "ovo_map_associate_batch_sharded: %d keyframes but the exchange has %d slots", F, xchg->slots);
  for (int r = 0; r < xchg->world; ++r) OVO_REQUIRE(xchg->peers.inbox[r] != nullptr, "ovo_map_associate_batch_sharded: peer %d not opened", r);
  OVO_TRY(ovo_map_batch_begin(m, xyz_dev, ins_ids_dev, N, frames, F, kf_slots, *next_ins_id, nullptr, 0, stream_));
  for (int f = 0; f < F; ++f)
    if (m->bt_table_len[f] > xchg->table_cap) {
      m->bt_valid = false;
      return ovo::set_error(OVO_E_INVALID, "ovo_map_associate_batch_sharded: table of keyframe %d (%d ints) exceeds the exchange's %lld", f,
                            m->bt_table_len[f], xchg->table_cap);
    }
  // one launch per keyframe with the exchange fused into its tail (default; measured faster beside the encoder: N=8 5873 vs
  // 5616 keyframes/s), or OVO_B200_VOTE=persistent: one persistent launch for the whole batch
  const char* mode = getenv("OVO_B200_VOTE");
  if (mode && mode[0] == 'p') OVO_TRY(batch_run_persistent(m, xchg, ins_ids_dev, static_cast<cudaStream_t>(stream_)));
  else OVO_TRY(batch_run_fused(m, xchg, ins_ids_dev, static_cast<cudaStream_t>(stream_)));
  return ovo_map_batch_end(m, ins_ids_dev, next_ins_id, votes_host, votes_stride, n_matched_host, mask_ins_out_dev, stream_);
}

int ovo_map_associate_batch(ovo_map_t* m, const float* xyz_dev, int32_t* ins_ids_dev, int64_t N, const ovo_frame* frames, int F,
                            const int* kf_slots, int* next_ins_id, ovo_vote_row* votes_host, int votes_stride, int* n_matched_host,
                            int32_t* mask_ins_out_dev, void* stream_) {
  OVO_REQUIRE(next_ins_id && votes_host && n_matched_host, "ovo_map_associate_batch: null argument");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  double px = 0;
  for (int f = 0; frames && f < F; ++f) px += static_cast<double>(frames[f].h) * frames[f].w * 8;
  ovo::ProfScope prof(stream, ovo::PROF_ASSOC, 0.0, static_cast<double>(N) * (12.0 + F * 14.0) + px);
  OVO_TRY(ovo_map_batch_begin(m, xyz_dev, ins_ids_dev, N, frames, F, kf_slots, *next_ins_id, nullptr, 0, stream_));
  {
    const char* mode = getenv("OVO_B200_VOTE");
    if (N > 0 && mode && mode[0] == 'p') OVO_TRY(batch_run_persistent(m, nullptr, ins_ids_dev, stream));
    else OVO_TRY(batch_run_fused(m, nullptr, ins_ids_dev, stream));
  }
  return ovo_map_batch_end(m, ins_ids_dev, next_ins_id, votes_host, votes_stride, n_matched_host, mask_ins_out_dev, stream_);
}

int ovo_map_get_matches(ovo_map_t* m, int kf_slot, int32_t* pairs_dev, int max_pairs, void* stream) {
  OVO_REQUIRE(m && kf_slot >= 0 && kf_slot < ovo_map::kSlots, "ovo_map_get_matches: bad slot");
  const int n = m->slot_n[kf_slot];
  OVO_REQUIRE(max_pairs >= n, "ovo_map_get_matches: buffer too small (%d < %d)", max_pairs, n);
  if (n > 0)
    OVO_CUDA(cudaMemcpyAsync(pairs_dev, m->slot_list[kf_slot], n * sizeof(int2), cudaMemcpyDeviceToDevice,
                             static_cast<cudaStream_t>(stream)));
  return n;
}

// descriptors f32 [R,D] -> the handle's bf16 staging copy (the update rule adds bf16-rounded descriptors)
static int stage_feats(ovo_map_t* m, const float* feats_dev, int R, int D, cudaStream_t stream) {
  const size_t n = static_cast<size_t>(R) * D;
  OVO_TRY(grow(&m->feats_bf16, &m->feats_cap, n));
  ovo::f32_to_bf16_pad_kernel<<<ovo::ceil_div(static_cast<long long>(n), 256), 256, 0, stream>>>(feats_dev, R, D, m->feats_bf16, R);
  OVO_CHECK_LAUNCH();
  return OVO_OK;
}

int ovo_map_fuse_dense(ovo_map_t* m, int kf_slot, void* bank_dev, void* bank_lo_dev, int32_t* counts_dev, int64_t N, int D,
                       const float* feats_dev, int n_rows, const int32_t* mask_row_dev, int n_masks, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  OVO_REQUIRE(m && kf_slot >= 0 && kf_slot < ovo_map::kSlots, "ovo_map_fuse_dense: bad slot");
  OVO_REQUIRE(bank_dev && bank_lo_dev && counts_dev && feats_dev && mask_row_dev, "ovo_map_fuse_dense: null argument");
  OVO_REQUIRE(D > 0 && D % 8 == 0 && n_rows > 0, "ovo_map_fuse_dense: D must be a multiple of 8, n_rows > 0");
  (void)N; (void)n_masks;
  bool dense_only = false;
  for (size_t i = 0; i < m->dense_slots.size(); ++i) dense_only = dense_only || (m->dense_slots[i] == kf_slot);
  if (dense_only) {   // the slot was filled by a batched association: it has a dense row, not a list
    const int slot1[1] = {kf_slot};
    return ovo_map_fuse_dense_batch(m, slot1, 1, bank_dev, bank_lo_dev, counts_dev, N, D, feats_dev, n_rows, mask_row_dev, n_masks, stream_);
  }
  const int n = m->slot_n[kf_slot];
  if (n == 0) return OVO_OK;
  OVO_REQUIRE(D <= 8 * 32 * 8, "ovo_map_fuse_dense: D > 2048 unsupported");
  const int blocks = std::min(ovo::ceil_div(n, 8), ovo::num_sms() * 8);
  ovo::ProfScope prof(stream, ovo::PROF_FUSE, 0.0, static_cast<double>(n) * (8.0 * D + 8));
  OVO_TRY(stage_feats(m, feats_dev, n_rows, D, stream));
  auto* hi = static_cast<__nv_bfloat16*>(bank_dev);
  auto* lo = static_cast<__nv_bfloat16*>(bank_lo_dev);
  if (D <= 1024)
    ovo::fuse_dense_kernel<4><<<blocks, 256, 0, stream>>>(m->slot_list[kf_slot], n, hi, lo, counts_dev, D, m->feats_bf16, mask_row_dev);
  else
    ovo::fuse_dense_kernel<8><<<blocks, 256, 0, stream>>>(m->slot_list[kf_slot], n, hi, lo, counts_dev, D, m->feats_bf16, mask_row_dev);
  OVO_CHECK_LAUNCH();
  return OVO_OK;
}

int ovo_map_fuse_dense_batch(ovo_map_t* m, const int* kf_slots_host, int n_slots, void* bank_dev, void* bank_lo_dev,
                             int32_t* counts_dev, int64_t N, int D, const float* feats_dev, int n_rows,
                             const int32_t* mask_row_dev, int n_masks, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  OVO_REQUIRE(m && kf_slots_host && bank_dev && bank_lo_dev && counts_dev && feats_dev && mask_row_dev, "ovo_map_fuse_dense_batch: null argument");
  OVO_REQUIRE(n_slots > 0 && n_slots <= 64 && N > 0 && N < (1LL << 31) && D > 0 && D % 8 == 0 && D <= 2048 && n_masks > 0 && n_masks < 32768 && n_rows > 0,
              "ovo_map_fuse_dense_batch: bad shape (slots %d, N %lld, D %d, masks %d)", n_slots, (long long)N, D, n_masks);
  // the dense rows a batched association left in the handle are used as they are when they cover exactly these slots in order
  bool direct = m->dense_N == N && m->dense_F >= n_slots && static_cast<int>(m->dense_slots.size()) >= n_slots;
  int first = -1;
  if (direct) {
    for (size_t i = 0; i < m->dense_slots.size() && first < 0; ++i)
      if (m->dense_slots[i] == kf_slots_host[0]) first = static_cast<int>(i);
    direct = first >= 0 && first + n_slots <= static_cast<int>(m->dense_slots.size());
    for (int i = 0; direct && i < n_slots; ++i) direct = m->dense_slots[first + i] == kf_slots_host[i];
  }
  const int16_t* rows = nullptr;
  int64_t stride = 0;
  double touched = 0;
  if (direct) {
    stride = m->dense_stride;
    rows = m->seg_dense + static_cast<size_t>(first) * stride;
    touched = static_cast<double>(N) * 0.25;   // profiling estimate only
  } else {
    long long total = 0;
    for (int i = 0; i < n_slots; ++i) {
      OVO_REQUIRE(kf_slots_host[i] >= 0 && kf_slots_host[i] < ovo_map::kSlots, "ovo_map_fuse_dense_batch: bad slot");
      for (size_t j = 0; j < m->dense_slots.size(); ++j)
        OVO_REQUIRE(m->dense_slots[j] != kf_slots_host[i], "ovo_map_fuse_dense_batch: slot %d was filled by a batched association; pass that batch's slots in its order", kf_slots_host[i]);
      total += m->slot_n[kf_slots_host[i]];
    }
    if (total == 0) return OVO_OK;
    stride = (N + 7) & ~int64_t(7);
    const size_t need = static_cast<size_t>(n_slots) * stride;
    OVO_TRY(grow(&m->seg_dense, &m->seg_dense_cap, need));
    m->dense_slots.clear(); m->dense_N = 0; m->dense_F = 0;
    OVO_CUDA(cudaMemsetAsync(m->seg_dense, 0xff, need * sizeof(int16_t), stream));   // -1 everywhere
    const int sms = ovo::num_sms();
    for (int i = 0; i < n_slots; ++i) {
      const int sl = kf_slots_host[i], n = m->slot_n[sl];
      if (n == 0) continue;
      ovo::scatter_matches_kernel<<<std::min(ovo::ceil_div(n, 256), sms * 4), 256, 0, stream>>>(m->slot_list[sl], n, m->seg_dense + static_cast<size_t>(i) * stride);
      OVO_CHECK_LAUNCH();
    }
    rows = m->seg_dense;
    touched = static_cast<double>(total);
  }
  ovo::ProfScope prof(stream, ovo::PROF_FUSE, 0.0, touched * (8.0 * D + 8));
  OVO_TRY(stage_feats(m, feats_dev, n_rows, D, stream));
  auto* hi = static_cast<__nv_bfloat16*>(bank_dev);
  auto* lo = static_cast<__nv_bfloat16*>(bank_lo_dev);
  const int parts = ovo::ceil_div(D, 512);   // a warp owns 512 columns of a row
  // short-lived blocks (a warp = ONE chunk of 32 points): a block that squats on an SM for the whole pass keeps the encoder's next
  // GEMM CTA (215 KB of shared memory, 54k registers) off that SM; blocks of a few microseconds give the SM back at once
  const int blocks = static_cast<int>((N + 127) / 128);
  if (D % 512 == 0)
    ovo::fuse_dense_batch_kernel<2, true><<<dim3(blocks, parts), 128, 0, stream>>>(rows, stride, n_slots, N, hi, lo, counts_dev, D, m->feats_bf16, mask_row_dev, n_masks);
  else
    ovo::fuse_dense_batch_kernel<2, false><<<dim3(blocks, parts), 128, 0, stream>>>(rows, stride, n_slots, N, hi, lo, counts_dev, D, m->feats_bf16, mask_row_dev, n_masks);
  OVO_CHECK_LAUNCH();
  ovo::fuse_counts_kernel<<<static_cast<int>((N + 255) / 256), 256, 0, stream>>>(rows, stride, n_slots, N, counts_dev, mask_row_dev, n_masks);
  OVO_CHECK_LAUNCH();
  return OVO_OK;
}

int ovo_bank_add_views(float* bank_dev, int D, const float* store_dev, const int32_t* quads_dev, const int32_t* idx_dev, int n, void* stream) {
  OVO_REQUIRE(bank_dev && store_dev && quads_dev && idx_dev && D > 0 && n >= 0, "ovo_bank_add_views: bad arguments");
  if (n == 0) return OVO_OK;
  ovo::bank_add_views_kernel<<<n, 256, 0, static_cast<cudaStream_t>(stream)>>>(bank_dev, D, store_dev, quads_dev, idx_dev);
  OVO_CHECK_LAUNCH();
  return OVO_OK;
}

int ovo_bank_update_mean(float* bank_dev, int32_t* counts_dev, int D, const float* feats_dev, const int32_t* rows_dev,
                         int n, void* stream) {
  OVO_REQUIRE(bank_dev && counts_dev && feats_dev && rows_dev && D > 0, "ovo_bank_update_mean: bad arguments");
  if (n <= 0) return OVO_OK;
  ovo::bank_update_mean_kernel<<<n, 256, 0, static_cast<cudaStream_t>(stream)>>>(bank_dev, counts_dev, D, feats_dev, rows_dev, n);
  OVO_CHECK_LAUNCH();
  return OVO_OK;
}

int ovo_query_dense(ovo_map_t* m, const void* bank_dev, int64_t N, int D, const float* text_dev, int Q, float* out_dev,
                    void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  OVO_REQUIRE(m && bank_dev && text_dev && out_dev, "ovo_query_dense: null argument");
  OVO_REQUIRE(N > 0 && N < (1LL << 31) && D > 0 && D % 8 == 0 && Q > 0, "ovo_query_dense: bad shape N=%lld D=%d Q=%d", (long long)N, D, Q);
  for (int q0 = 0; q0 < Q; q0 += 256) {
    const int qn = std::min(256, Q - q0);
    const int qpad = qn <= 32 ? 32 : (qn <= 64 ? 64 : (qn <= 128 ? 128 : 256));
    OVO_TRY(grow(&m->text_bf16, &m->text_cap, static_cast<size_t>(256) * D));
    ovo::f32_to_bf16_pad_kernel<<<ovo::ceil_div(static_cast<long long>(qpad) * D, 256), 256, 0, stream>>>(
        text_dev + static_cast<size_t>(q0) * D, qn, D, m->text_bf16, qpad);
    OVO_CHECK_LAUNCH();
    ovo::EpiParams ep;
    ep.out = out_dev + q0;
    ep.ldo = Q;
    ep.prof_cls = ovo::PROF_QUERY;
    OVO_TRY(ovo::launch_gemm(ovo::EPI_F32, static_cast<const __nv_bfloat16*>(bank_dev), D, m->text_bf16, D,
                             static_cast<int>(N), qn, D, ep, stream, qpad));
  }
  return OVO_OK;
}

int ovo_query_instances(const float* bank_dev, const int32_t* rows_dev, int I, int D, const float* text_dev, int Q,
                        float* out_dev, void* stream) {
  OVO_REQUIRE(bank_dev && text_dev && out_dev && I > 0 && D > 0 && Q > 0, "ovo_query_instances: bad arguments");
  const long long threads = static_cast<long long>(I) * Q * 32;
  ovo::query_instances_kernel<<<ovo::ceil_div(threads, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(bank_dev, rows_dev, I, D, text_dev, Q, out_dev);
  OVO_CHECK_LAUNCH();
  return OVO_OK;
}

int ovo_merge_masks(const uint8_t* masks_dev, int M, int H, int W, const int32_t* group_dev, int R, uint8_t* out_dev,
                    int32_t* areas_dev, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  OVO_REQUIRE(masks_dev && group_dev && out_dev && areas_dev && M > 0 && R > 0, "ovo_merge_masks: bad arguments");
  OVO_REQUIRE((static_cast<long long>(H) * W) % 4 == 0 && (reinterpret_cast<uintptr_t>(masks_dev) & 3) == 0 &&
                  (reinterpret_cast<uintptr_t>(out_dev) & 3) == 0, "ovo_merge_masks: H*W must be a multiple of 4 and buffers 4-byte aligned");
  const int npix4 = H * W / 4;
  OVO_CUDA(cudaMemsetAsync(areas_dev, 0, R * sizeof(int32_t), stream));
  ovo::merge_masks_kernel<<<dim3(std::min(ovo::ceil_div(npix4, 256), 64), R), 256, 0, stream>>>(masks_dev, M, npix4, group_dev, R, out_dev, areas_dev);
  OVO_CHECK_LAUNCH();
  return OVO_OK;
}

int ovo_fuse_views(const float* store_dev, int D, const int32_t* idx_dev, const int32_t* off_dev, int n_instances, int mode,
                   float* bank_dev, const int32_t* out_rows_dev, int32_t* chosen_dev, void* stream) {
  OVO_REQUIRE(store_dev && idx_dev && off_dev && bank_dev && out_rows_dev && D > 0 && mode >= 0 && mode <= 2, "ovo_fuse_views: bad arguments");
  if (n_instances <= 0) return OVO_OK;
  ovo::fuse_views_kernel<<<n_instances, 256, 0, static_cast<cudaStream_t>(stream)>>>(store_dev, D, idx_dev, off_dev, mode, bank_dev, out_rows_dev, chosen_dev);
  OVO_CHECK_LAUNCH();
  return OVO_OK;
}

int ovo_map_integrate(ovo_map_t* m, float* xyz_dev, int32_t* ids_dev, int32_t* ins_ids_dev, uint8_t* colors_dev, int64_t N,
                      int64_t capacity, const float* depth_dev, const uint8_t* rgb_dev, int h, int w, const float* c2w,
                      const float* w2c, const float* K, float match_th, int downscale, int k_pool, int next_point_id,
                      int* n_new_host, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  OVO_REQUIRE(m && xyz_dev && ids_dev && ins_ids_dev && depth_dev && c2w && w2c && K && n_new_host, "ovo_map_integrate: null argument");
  OVO_REQUIRE(N >= 0 && capacity >= N && h > 0 && w > 0 && downscale >= 1 && k_pool >= 1 && (k_pool & 1), "ovo_map_integrate: bad arguments");
  const int npix = h * w, hd = (h + downscale - 1) / downscale, wd = (w + downscale - 1) / downscale;
  OVO_TRY(grow(&m->mapped, &m->mapped_cap, static_cast<size_t>(npix)));
  OVO_TRY(grow(&m->pix_flags, &m->pix_cap, 2 * static_cast<size_t>(hd) * wd));
  ovo::FrameDev* hf = m->h_frame;
  memset(hf, 0, sizeof(*hf));
  memcpy(hf->c2w, c2w, sizeof(hf->c2w)); memcpy(hf->w2c, w2c, sizeof(hf->w2c)); memcpy(hf->K, K, sizeof(hf->K));
  hf->match_th = match_th; hf->h = h; hf->w = w; hf->H = h; hf->W = w;
  CtlHeader* hh = reinterpret_cast<CtlHeader*>(m->h_ctl);
  memset(&hh->geom, 0, sizeof(hh->geom));
  hh->geom.dmin_bits = 0x7f800000; hh->geom.dmax_bits = 0;
  m->h_counters[0] = m->h_counters[1] = m->h_counters[2] = m->h_counters[3] = 0;
  OVO_CUDA(cudaMemcpyAsync(m->ctl, m->h_ctl, sizeof(CtlHeader), cudaMemcpyHostToDevice, stream));
  const int sms = ovo::num_sms();
  const int erode = N > 0 ? 1 : 0;   // vanilla_mapper.py:58: suppression + pooling only once the map is not empty
  if (N > 0) {
    OVO_CUDA(cudaMemsetAsync(m->mapped, 0, npix, stream));
    ovo::depth_minmax_kernel<<<sms, 256, 0, stream>>>(depth_dev, npix, m->geom);
    OVO_CHECK_LAUNCH();
    ovo::frustum_setup_kernel<<<1, 1, 0, stream>>>(m->geom, m->frame);
    OVO_CHECK_LAUNCH();
    ovo::mark_mapped_pixels_kernel<<<static_cast<int>(std::min<long long>((N + 255) / 256, sms * 8LL)), 256, 0, stream>>>(
        xyz_dev, N, depth_dev, m->geom, m->frame, m->mapped);
    OVO_CHECK_LAUNCH();
  }
  int32_t* flags = m->pix_flags;
  int32_t* pos = m->pix_flags + static_cast<size_t>(hd) * wd;
  ovo::new_point_flags_kernel<<<ovo::ceil_div(hd * wd, 256), 256, 0, stream>>>(depth_dev, m->mapped, h, w, downscale, k_pool > 1 ? k_pool : 1,
                                                                             erode && true, hd, wd, flags);
  OVO_CHECK_LAUNCH();
  ovo::scan_flags_kernel<<<1, 1024, 0, stream>>>(flags, hd * wd, pos, m->counters);
  OVO_CHECK_LAUNCH();
  ovo::emit_points_kernel<<<ovo::ceil_div(hd * wd, 256), 256, 0, stream>>>(flags, pos, hd, wd, downscale, depth_dev, rgb_dev, w, m->frame,
                                                                         capacity - N, xyz_dev + 3 * N, ids_dev + N, ins_ids_dev + N,
                                                                         colors_dev ? colors_dev + 3 * N : nullptr, next_point_id);
  OVO_CHECK_LAUNCH();
  OVO_CUDA(cudaMemcpyAsync(m->h_counters, m->counters, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
  OVO_CUDA(cudaStreamSynchronize(stream));
  *n_new_host = m->h_counters[0];
  if (N + *n_new_host > capacity) return ovo::set_error(OVO_E_NOMEM, "ovo_map_integrate: %d new points exceed the capacity (%lld + %d > %lld)", *n_new_host, (long long)N, *n_new_host, (long long)capacity);
  return OVO_OK;
}

int ovo_mask_nms(const uint8_t* masks_dev, const float* scores_dev, int M, int H, int W, float iou_thr, float score_thr,
                 float inner_thr, uint8_t* keep_dev, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  OVO_REQUIRE(masks_dev && scores_dev && keep_dev && M > 0 && M <= 4096 && H > 0 && W > 0, "ovo_mask_nms: bad arguments");
  const int npix = H * W, nwords = ovo::ceil_div(npix, 32);
  uint32_t* bits = nullptr; int32_t *rank = nullptr, *order = nullptr, *inter = nullptr; uint8_t* keep_sorted = nullptr;
  // small, rare call (once per keyframe at most): stream-ordered temporaries
  OVO_CUDA(cudaMallocAsync(&bits, static_cast<size_t>(M) * nwords * 4, stream));
  OVO_CUDA(cudaMallocAsync(&rank, M * 4, stream));
  OVO_CUDA(cudaMallocAsync(&order, M * 4, stream));
  OVO_CUDA(cudaMallocAsync(&inter, static_cast<size_t>(M) * M * 4, stream));
  OVO_CUDA(cudaMallocAsync(&keep_sorted, M, stream));
  ovo::bitpack_masks_kernel<<<dim3(ovo::ceil_div(nwords, 256), M), 256, 0, stream>>>(masks_dev, M, npix, nwords, bits);
  OVO_CHECK_LAUNCH();
  ovo::rank_desc_kernel<<<ovo::ceil_div(M, 128), 128, 0, stream>>>(scores_dev, M, rank, order);
  OVO_CHECK_LAUNCH();
  ovo::mask_intersections_kernel<<<dim3(M, M), 256, 0, stream>>>(bits, nwords, order, M, inter);
  OVO_CHECK_LAUNCH();
  ovo::mask_nms_decide_kernel<<<ovo::ceil_div(M, 128), 128, 0, stream>>>(inter, scores_dev, order, M, iou_thr, score_thr,
                                                                              static_cast<float>(1.0 - static_cast<double>(inner_thr)), keep_sorted);
  OVO_CHECK_LAUNCH();
  // keep flags back in ORIGINAL mask order (masks_update/filter keep the original order, segment_utils.py:188-193)
  ovo::gather_u8_kernel<<<ovo::ceil_div(M, 128), 128, 0, stream>>>(keep_sorted, rank, M, keep_dev);
  OVO_CHECK_LAUNCH();
  cudaFreeAsync(bits, stream); cudaFreeAsync(rank, stream); cudaFreeAsync(order, stream); cudaFreeAsync(inter, stream);
  cudaFreeAsync(keep_sorted, stream);
  return OVO_OK;
}

int ovo_mask2segmap(const uint8_t* masks_dev, const float* stability_dev, int M, int H, int W, int32_t* seg_map_dev,
                    uint8_t* maps_out_dev, int32_t* order_dev, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  OVO_REQUIRE(masks_dev && stability_dev && seg_map_dev && maps_out_dev && order_dev && M > 0 && H > 0 && W > 0, "ovo_mask2segmap: bad arguments");
  int32_t* rank = nullptr;
  OVO_CUDA(cudaMallocAsync(&rank, M * 4, stream));
  ovo::rank_desc_kernel<<<ovo::ceil_div(M, 128), 128, 0, stream>>>(stability_dev, M, rank, order_dev);
  OVO_CHECK_LAUNCH();
  ovo::paint_segmap_kernel<<<ovo::ceil_div(H * W, 256), 256, 0, stream>>>(masks_dev, order_dev, M, H * W, seg_map_dev, maps_out_dev);
  OVO_CHECK_LAUNCH();
  cudaFreeAsync(rank, stream);
  return OVO_OK;
}

int ovo_classify(const float* sim_dev, int64_t n, int Q, float th, int32_t* cls_dev, float* conf_dev, void* stream) {
  OVO_REQUIRE(sim_dev && cls_dev && conf_dev && n > 0 && Q > 0, "ovo_classify: bad arguments");
  ovo::classify_kernel<<<ovo::ceil_div(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(sim_dev, n, Q, th, cls_dev, conf_dev);
  OVO_CHECK_LAUNCH();
  return OVO_OK;
}

}  // extern "C"
