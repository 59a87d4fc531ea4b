// Device-side exchange of vote tables over peer memory (NVLink / NVSwitch): shared by p2p.cu (stand-alone exchange kernel) and
// map.cu (the exchange fused into the tail of the vote kernel).  See p2p.cu for the protocol.
#pragma once
#include <vector>

#include "common.cuh"

namespace ovo {

struct XchgPeers {
  int32_t* inbox[16];
  int32_t* flags[16];
};

__device__ __forceinline__ void st_release_sys(int32_t* p, int32_t v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int32_t ld_acquire_sys(const int32_t* p) {
  int32_t v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// ONE block: push `table` (n ints, 16-byte aligned, padded to 4) into every rank's inbox, raise the flags, wait for the world
// tables of this exchange in the own inbox, and leave their sum in `table`.  Every thread of the block must call it.
__device__ __forceinline__ void xchg_block(const XchgPeers& peers, int rank, int world, int slots, int slot, int parity, int epoch,
                                           long long table_cap, int32_t* table, int n) {
  const int n4 = (n + 3) >> 2;
  const size_t box = (static_cast<size_t>(parity) * slots + slot) * world;   // [parity][slot][src]
  const int4* in = reinterpret_cast<const int4*>(table);
  for (int dst = 0; dst < world; ++dst) {
    int4* out = reinterpret_cast<int4*>(peers.inbox[dst] + (box + rank) * table_cap);
    for (int i = threadIdx.x; i < n4; i += blockDim.x) out[i] = __ldcg(in + i);
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < world) st_release_sys(peers.flags[threadIdx.x] + static_cast<size_t>(slot) * world + rank, epoch);
  if (threadIdx.x < world) {
    const int32_t* f = peers.flags[rank] + static_cast<size_t>(slot) * world + threadIdx.x;
    while (ld_acquire_sys(f) < epoch) {
    }
  }
  __syncthreads();
  const int4* mine = reinterpret_cast<const int4*>(peers.inbox[rank] + box * table_cap);
  const size_t stride4 = static_cast<size_t>(table_cap) >> 2;
  for (int i = threadIdx.x; i < n4; i += blockDim.x) {
    int4 acc = make_int4(0, 0, 0, 0);
    for (int s = 0; s < world; ++s) {
      const int4 v = __ldcg(mine + s * stride4 + i);     // L2 (the peers wrote it there), never a stale L1 line
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    reinterpret_cast<int4*>(table)[i] = acc;
  }
  __syncthreads();
}

}  // namespace ovo

struct ovo_xchg {
  int rank = 0, world = 1, slots = 0;
  long long table_cap = 0;      // ints per table (multiple of 4)
  int32_t* base = nullptr;      // local allocation: inbox [2][slots][world][table_cap] then flags [slots][world]
  size_t inbox_ints = 0;
  ovo::XchgPeers peers{};
  void* opened[16] = {};
  std::vector<int>* epochs = nullptr;
};
