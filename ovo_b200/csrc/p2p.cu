// Device-side exchange of the per-keyframe vote tables of a map sharded over the GPUs of one box: the ranks'
// partial tables [n_matched | votes n_masks x (n_ins+1)] are SUMMED on every rank without a host-launched collective.
//
// Every rank owns an inbox [parity 2][slot][source rank][table_cap ints] + flags [slot][source rank] in plain cudaMalloc
// memory whose IPC handle the peers have opened (NVLink peer mapping through NVSwitch).  One kernel per keyframe and rank:
//   block r < world : copies this rank's table into peer r's inbox (16-byte stores over NVLink), fences system-wide and
//                     raises flag[slot][my rank] there to the exchange's epoch (st.release.sys);
//   every block     : waits until all `world` flags of its OWN inbox have reached the epoch (ld.acquire.sys on local
//                     memory), then sums the world tables of its element range into the caller's table (integer sums: the
//                     result is identical on every rank, whatever the arrival order).
// Only the COMPACT table travels: the instance count n_ins lives on the device (ovo_map_batch_*), the kernel reads it there.
// Reuse of an inbox buffer is safe with two parities: a peer can only be one exchange of the same slot ahead (it needed this
// rank's push of that exchange to finish its own), and that exchange uses the other parity.
#include <algorithm>
#include <cstring>

#include "p2p.cuh"

namespace ovo {

__device__ __forceinline__ int compact_len(int n_bound, const int32_t* n_ins_dev, int n_masks) {
  // header (4 ints) + n_masks x (n_ins + 1) votes
  return n_ins_dev != nullptr ? min(n_bound, 4 + max(n_masks, 1) * (*n_ins_dev + 1)) : n_bound;
}

// small tables: one block does everything (push, flags, wait, sum)
__global__ void __launch_bounds__(512)
    vote_exchange_kernel(XchgPeers peers, int rank, int world, int slots, int slot, int parity, int epoch, long long table_cap,
                         int32_t* table, int n_bound, const int32_t* __restrict__ n_ins_dev, int n_masks) {
  xchg_block(peers, rank, world, slots, slot, parity, epoch, table_cap, table, compact_len(n_bound, n_ins_dev, n_masks));
}

// large tables, two launches (the sum overwrites `table`, so every push must have read it first: stream order):
// block r copies the table into peer r's inbox and raises the flag there ...
__global__ void __launch_bounds__(256)
    vote_push_kernel(XchgPeers peers, int rank, int world, int slots, int slot, int parity, int epoch, long long table_cap,
                     const int32_t* __restrict__ table, int n_bound, const int32_t* __restrict__ n_ins_dev, int n_masks) {
  const int n4 = (compact_len(n_bound, n_ins_dev, n_masks) + 3) >> 2;
  const size_t box = (static_cast<size_t>(parity) * slots + slot) * world;
  const int dst = blockIdx.x;
  int4* out = reinterpret_cast<int4*>(peers.inbox[dst] + (box + rank) * table_cap);
  const int4* in = reinterpret_cast<const int4*>(table);
  for (int i = threadIdx.x; i < n4; i += blockDim.x) out[i] = in[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) st_release_sys(peers.flags[dst] + static_cast<size_t>(slot) * world + rank, epoch);
}
// ... then every block waits for the world flags of the own inbox and sums its element range
__global__ void __launch_bounds__(256)
    vote_sum_kernel(XchgPeers peers, int rank, int world, int slots, int slot, int parity, int epoch, long long table_cap,
                    int32_t* __restrict__ table, int n_bound, const int32_t* __restrict__ n_ins_dev, int n_masks) {
  const int n4 = (compact_len(n_bound, n_ins_dev, n_masks) + 3) >> 2;
  const size_t box = (static_cast<size_t>(parity) * slots + slot) * world;
  if (threadIdx.x < world) {
    const int32_t* f = peers.flags[rank] + static_cast<size_t>(slot) * world + threadIdx.x;
    while (ld_acquire_sys(f) < epoch) {
    }
  }
  __syncthreads();
  const int4* mine = reinterpret_cast<const int4*>(peers.inbox[rank] + box * table_cap);
  const size_t stride4 = static_cast<size_t>(table_cap) >> 2;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
    int4 acc = make_int4(0, 0, 0, 0);
    for (int s = 0; s < world; ++s) {
      const int4 v = __ldcg(mine + s * stride4 + i);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    reinterpret_cast<int4*>(table)[i] = acc;
  }
}

// Packing of freshly mapped points for the fixed-size all-to-all of a sharded map (ovo_b200/sharding.py route_new_points_fixed):
// record slot dst * cap + (position of the point among this rank's points of shard dst, creation order) <- (x, y, z, id); unused
// slots keep a far-away sentinel (id -1).  ONE block: every thread takes a contiguous run of points, counts them per shard, the
// per-shard counts are scanned over the threads (stable order, deterministic), then the records are scattered.
__device__ __forceinline__ int shard_of(float x, float y, float z, float cell, int world) {
  const long long vx = static_cast<long long>(floorf(__fdiv_rn(x, cell))), vy = static_cast<long long>(floorf(__fdiv_rn(y, cell))),
                  vz = static_cast<long long>(floorf(__fdiv_rn(z, cell)));
  const long long h = (vx * 73856093LL) ^ (vy * 19349663LL) ^ (vz * 83492791LL);
  return static_cast<int>(((h % world) + world) % world);
}

__global__ void __launch_bounds__(1024)
    route_pack_kernel(const float* __restrict__ xyz, const int32_t* __restrict__ ids, int n, int world, float cell, int cap, float far,
                      float4* __restrict__ rec, int32_t* __restrict__ overflow) {
  // tiles of 1024 consecutive points (coalesced reads); inside a tile a point's position in its shard's run = the shard's running
  // base + the points of that shard in lower warps of the tile + those in lower lanes of its warp (stable: creation order)
  __shared__ int s_wcnt[16][32];     // per shard, per warp: points of the tile
  __shared__ int s_base[16];         // per shard: points of the earlier tiles
  const int t = threadIdx.x, T = blockDim.x, warp = t >> 5, lane = t & 31;
  const float4 sentinel = make_float4(far, far, far, __int_as_float(-1));
  for (int i = t; i < world * cap; i += T) rec[i] = sentinel;
  if (t < 16) s_base[t] = 0;
  __syncthreads();
  for (int base = 0; base < n; base += T) {
    const int i = base + t;
    float x = 0.f, y = 0.f, z = 0.f;
    int d = -1;
    if (i < n) {
      x = xyz[3 * i]; y = xyz[3 * i + 1]; z = xyz[3 * i + 2];
      d = shard_of(x, y, z, cell, world);
    }
    int rank = 0;
    for (int k = 0; k < world; ++k) {
      const unsigned bal = __ballot_sync(0xffffffffu, d == k);
      if (d == k) rank = __popc(bal & ((1u << lane) - 1));
      if (lane == 0) s_wcnt[k][warp] = __popc(bal);
    }
    __syncthreads();
    int before = 0;
    if (d >= 0)
      for (int w = 0; w < warp; ++w) before += s_wcnt[d][w];
    const int p = d >= 0 ? s_base[d] + before + rank : 0;
    if (d >= 0 && p < cap) rec[static_cast<size_t>(d) * cap + p] = make_float4(x, y, z, __int_as_float(ids[i]));
    __syncthreads();
    if (t < world) {
      int tot = 0;
      for (int w = 0; w < 32; ++w) tot += s_wcnt[t][w];
      s_base[t] += tot;
    }
    __syncthreads();
  }
  if (t < world && s_base[t] > cap) atomicAdd(overflow, s_base[t] - cap);
}

}  // namespace ovo

extern "C" {

int ovo_xchg_create(int rank, int world, int slots, int64_t table_ints, ovo_xchg_t** out) {
  OVO_REQUIRE(out && world >= 1 && world <= 16 && rank >= 0 && rank < world && slots > 0 && slots <= 64 && table_ints > 0,
              "ovo_xchg_create: bad arguments (world <= 16, slots <= 64)");
  ovo_xchg* x = new ovo_xchg();
  x->rank = rank; x->world = world; x->slots = slots;
  x->table_cap = (table_ints + 3) & ~int64_t(3);
  x->inbox_ints = static_cast<size_t>(2) * slots * world * x->table_cap;
  const size_t total = (x->inbox_ints + static_cast<size_t>(slots) * world + 64) * sizeof(int32_t);
  if (cudaMalloc(reinterpret_cast<void**>(&x->base), total) != cudaSuccess) {
    cudaGetLastError();
    delete x;
    return ovo::set_error(OVO_E_NOMEM, "ovo_xchg_create: %zu bytes for the inbox", total);
  }
  cudaMemset(x->base, 0, total);
  cudaDeviceSynchronize();
  x->peers.inbox[rank] = x->base;
  x->peers.flags[rank] = x->base + x->inbox_ints;
  x->epochs = new std::vector<int>(slots, 0);
  *out = x;
  return OVO_OK;
}

int ovo_xchg_ipc_handle(ovo_xchg_t* x, void* handle_out_64) {
  OVO_REQUIRE(x && handle_out_64, "ovo_xchg_ipc_handle: null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  cudaIpcMemHandle_t h;
  OVO_CUDA(cudaIpcGetMemHandle(&h, x->base));
  memcpy(handle_out_64, &h, sizeof(h));
  return OVO_OK;
}

int ovo_xchg_open_peers(ovo_xchg_t* x, const void* handles_world_x_64) {
  OVO_REQUIRE(x && handles_world_x_64, "ovo_xchg_open_peers: null argument");
  for (int r = 0; r < x->world; ++r) {
    if (r == x->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, static_cast<const uint8_t*>(handles_world_x_64) + 64 * r, sizeof(h));
    void* p = nullptr;
    OVO_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    x->opened[r] = p;
    x->peers.inbox[r] = static_cast<int32_t*>(p);
    x->peers.flags[r] = static_cast<int32_t*>(p) + x->inbox_ints;
  }
  return OVO_OK;
}

int ovo_xchg_exchange(ovo_xchg_t* x, int32_t* table_dev, int n_ints, const int32_t* n_ins_dev, int n_masks, int slot, void* stream) {
  OVO_REQUIRE(x && table_dev && n_ints > 0 && n_ints <= x->table_cap && slot >= 0 && slot < x->slots,
              "ovo_xchg_exchange: table of %d ints / slot %d outside the exchange's limits (%lld ints, %d slots)", n_ints, slot,
              x ? x->table_cap : 0LL, x ? x->slots : 0);
  OVO_REQUIRE((reinterpret_cast<uintptr_t>(table_dev) & 15) == 0, "ovo_xchg_exchange: table must be 16-byte aligned");
  for (int r = 0; r < x->world; ++r) OVO_REQUIRE(x->peers.inbox[r] != nullptr, "ovo_xchg_exchange: peer %d not opened", r);
  const int epoch = ++(*x->epochs)[slot];
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n_ints <= 32 * 1024) {
    ovo::vote_exchange_kernel<<<1, 512, 0, st>>>(x->peers, x->rank, x->world, x->slots, slot, epoch & 1, epoch, x->table_cap, table_dev,
                                                 n_ints, n_ins_dev, n_masks);
    OVO_CHECK_LAUNCH();
  } else {
    ovo::vote_push_kernel<<<x->world, 256, 0, st>>>(x->peers, x->rank, x->world, x->slots, slot, epoch & 1, epoch, x->table_cap, table_dev,
                                                   n_ints, n_ins_dev, n_masks);
    OVO_CHECK_LAUNCH();
    ovo::vote_sum_kernel<<<std::min(64, ovo::ceil_div(n_ints, 4 * 256)), 256, 0, st>>>(x->peers, x->rank, x->world, x->slots, slot, epoch & 1,
                                                                                    epoch, x->table_cap, table_dev, n_ints, n_ins_dev, n_masks);
    OVO_CHECK_LAUNCH();
  }
  return OVO_OK;
}

int ovo_route_pack(const float* xyz_dev, const int32_t* ids_dev, int n, int world, float cell, int cap_per_dst, float far_value,
                   float* records_out_dev, int32_t* overflow_dev, void* stream) {
  OVO_REQUIRE(xyz_dev && ids_dev && records_out_dev && overflow_dev && n >= 0 && world >= 1 && world <= 16 && cap_per_dst > 0 && cell > 0.f,
              "ovo_route_pack: bad arguments (world <= 16)");
  OVO_REQUIRE((reinterpret_cast<uintptr_t>(records_out_dev) & 15) == 0, "ovo_route_pack: records must be 16-byte aligned");
  ovo::route_pack_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(
      xyz_dev, ids_dev, n, world, cell, cap_per_dst, far_value, reinterpret_cast<float4*>(records_out_dev), overflow_dev);
  OVO_CHECK_LAUNCH();
  return OVO_OK;
}

void ovo_xchg_destroy(ovo_xchg_t* x) {
  if (!x) return;
  cudaDeviceSynchronize();
  for (int r = 0; r < x->world; ++r)
    if (x->opened[r]) cudaIpcCloseMemHandle(x->opened[r]);
  cudaFree(x->base);
  delete x->epochs;
  delete x;
}

}  // extern "C"
