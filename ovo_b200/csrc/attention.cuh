// Flash-style softmax attention forward for head_dim 64 on tcgen05 (reference: F.scaled_dot_product_attention
// at pe.py:145-147 for the vision tower, nn.MultiheadAttention with the causal mask pe.py:621-627 for text).
//
// One CTA = one (image, head, 128-query tile).  All of Q (128x64), K (seq_pad x 64) and V^T (64 x seq_pad)
// for the head are TMA-loaded once into 128B-swizzled shared memory (<= 176 KB at seq_pad 640).
// Per 128-key block: S = Q.K^T (tcgen05, accumulator in TMEM, double buffered so block j+1's QK^T overlaps
// block j's softmax) -> tcgen05.ld -> online softmax in registers (one thread per query row, exp2f) ->
// P (bf16) written to swizzled smem -> O_j = P.V_j (tcgen05) -> tcgen05.ld -> rescale-and-accumulate in
// registers.  q/k/v^T are produced in exactly this layout by the QKV GEMM epilogue (gemm.cuh EPI_QKV).
#pragma once
#include "ptx.cuh"

namespace ovo {

constexpr int kAttnThreads = 128;
constexpr int kAttnMaxBlocks = 5;  // seq_pad <= 640

struct AttnSmem {
  static constexpr int kQ = 128 * 64 * 2;             // 16 KB
  static constexpr int kKBlock = 128 * 64 * 2;        // 16 KB per 128 keys
  static constexpr int kVBlock = 64 * 64 * 2;         // 8 KB per 64 keys (V^T tile: 64 d-rows x 64 keys)
  static constexpr int kP = 2 * 128 * 64 * 2;         // 32 KB: P as two K-major 128x64 tiles
  static constexpr int bytes(int nblk) { return kQ + nblk * kKBlock + 2 * nblk * kVBlock + kP + 1024 + 256; }
};

__global__ void __launch_bounds__(kAttnThreads, 1)
    attention_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                         const __grid_constant__ CUtensorMap tmVt, __nv_bfloat16* __restrict__ out, int seq,
                         int seq_pad, int heads, int ld_out, float scale_log2e, int causal) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int nblk = seq_pad / 128;
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + AttnSmem::kQ;
  uint8_t* sV = sK + nblk * AttnSmem::kKBlock;
  uint8_t* sP = sV + 2 * nblk * AttnSmem::kVBlock;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + AttnSmem::kP);
  uint64_t* bar_k = bars;                      // [nblk]  (bar_k[0] also covers Q)
  uint64_t* bar_v = bars + kAttnMaxBlocks;     // [nblk]
  uint64_t* bar_s = bars + 2 * kAttnMaxBlocks; // [2] S buffer ready
  uint64_t* bar_o = bar_s + 2;                 // [1] P.V ready
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_o + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int qt = blockIdx.x;   // query tile
  const int bh = blockIdx.y;   // image * heads + head
  const int b = bh / heads, head = bh - b * heads;
  const int q0 = qt * 128;
  // causal: keys beyond the last query of this tile are never needed
  const int nblk_used = causal ? min(nblk, qt + 1) : nblk;

  if (tid == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmVt);
    for (int i = 0; i < nblk; ++i) { mbar_init(&bar_k[i], 1); mbar_init(&bar_v[i], 1); }
    mbar_init(&bar_s[0], 1); mbar_init(&bar_s[1], 1); mbar_init(&bar_o[0], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(tmem_slot);  // S0 [0,128) S1 [128,256) O [256,320)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S[2] = {tmem_base, tmem_base + 128};
  const uint32_t tmem_O = tmem_base + 256;

  constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128);
  constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64);
  const uint64_t qdesc = umma_desc_sw128(smem_u32(sQ));

  if (tid == 0) {
    for (int j = 0; j < nblk_used; ++j) {
      mbar_arrive_expect_tx(&bar_k[j], AttnSmem::kKBlock + (j == 0 ? AttnSmem::kQ : 0));
      if (j == 0) tma_load_2d(sQ, &tmQ, &bar_k[0], 0, bh * seq_pad + q0);
      tma_load_2d(sK + j * AttnSmem::kKBlock, &tmK, &bar_k[j], 0, bh * seq_pad + j * 128);
    }
    for (int j = 0; j < nblk_used; ++j) {
      mbar_arrive_expect_tx(&bar_v[j], 2 * AttnSmem::kVBlock);
      tma_load_2d(sV + (2 * j) * AttnSmem::kVBlock, &tmVt, &bar_v[j], j * 128, bh * 64);
      tma_load_2d(sV + (2 * j + 1) * AttnSmem::kVBlock, &tmVt, &bar_v[j], j * 128 + 64, bh * 64);
    }
    // S_0 = Q . K_0^T
    mbar_wait(&bar_k[0], 0);
    tc_fence_after();
    const uint64_t kdesc = umma_desc_sw128(smem_u32(sK));
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_bf16(tmem_S[0], qdesc + 2 * k, kdesc + 2 * k, idesc_s, k != 0);
    umma_commit(&bar_s[0]);
  }

  // per-thread state: one query row each
  const int qrow = q0 + tid;
  float m_run = -INFINITY, l_run = 0.f;
  float o_acc[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) o_acc[i] = 0.f;
  const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
  const int r8 = tid & 7;
  uint8_t* p_row = sP + (tid >> 3) * 1024 + r8 * 128;

  for (int j = 0; j < nblk_used; ++j) {
    const int buf = j & 1;
    // issue next block's QK^T as early as possible
    if (tid == 0 && j + 1 < nblk_used) {
      mbar_wait(&bar_k[j + 1], 0);
      tc_fence_after();
      const uint64_t kdesc = umma_desc_sw128(smem_u32(sK + (j + 1) * AttnSmem::kKBlock));
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16(tmem_S[buf ^ 1], qdesc + 2 * k, kdesc + 2 * k, idesc_s, k != 0);
      umma_commit(&bar_s[buf ^ 1]);
    }
    __syncwarp();
    mbar_wait(&bar_s[buf], (j >> 1) & 1);
    tc_fence_after();

    const int kv0 = j * 128;
    int kv_hi = seq - kv0;                                   // keys >= seq are padding
    if (causal) kv_hi = min(kv_hi, qrow - kv0 + 1);          // keys > query are masked
    // pass 1: row max
    float m_blk = -INFINITY;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_S[buf] + lane_off + c * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (c * 32 + i < kv_hi) m_blk = fmaxf(m_blk, __uint_as_float(v[i]));
    }
    const float m_new = fmaxf(m_run, m_blk);
    const float m_scaled = (m_new == -INFINITY) ? 0.f : m_new * scale_log2e;
    const float alpha = (m_run == -INFINITY) ? 0.f : exp2f(m_run * scale_log2e - m_scaled);
    // pass 2: p = exp2(s*scale - m), row sum, P -> swizzled smem (bf16)
    float l_blk = 0.f;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_S[buf] + lane_off + c * 32, v);
      tmem_ld_wait();
      float p[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float e = exp2f(__uint_as_float(v[i]) * scale_log2e - m_scaled);
        p[i] = (c * 32 + i < kv_hi) ? e : 0.f;
      }
      uint8_t* tile = p_row + (c >> 1) * (128 * 128);  // keys 0-63 -> tile 0, 64-127 -> tile 1
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        // round to bf16 first so that the row sum matches what the tensor core multiplies
        uint4 u;
        u.x = pack_bf16(p[8 * g], p[8 * g + 1]); u.y = pack_bf16(p[8 * g + 2], p[8 * g + 3]);
        u.z = pack_bf16(p[8 * g + 4], p[8 * g + 5]); u.w = pack_bf16(p[8 * g + 6], p[8 * g + 7]);
        const int chunk = (c & 1) * 4 + g;  // 16-byte chunk index inside the 128-byte row
        *reinterpret_cast<uint4*>(tile + ((chunk ^ r8) << 4)) = u;
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) l_blk += p[i];
    }
    l_run = l_run * alpha + l_blk;
    m_run = m_new;
#pragma unroll
    for (int i = 0; i < 64; ++i) o_acc[i] *= alpha;

    // P visible to the async proxy, then O_j = P . V_j
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      mbar_wait(&bar_v[j], 0);
      tc_fence_after();
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const uint64_t pdesc = umma_desc_sw128(smem_u32(sP + t * (128 * 128)));
        const uint64_t vdesc = umma_desc_sw128(smem_u32(sV + (2 * j + t) * AttnSmem::kVBlock));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_O, pdesc + 2 * k, vdesc + 2 * k, idesc_o, (t | k) != 0);
      }
      umma_commit(&bar_o[0]);
    }
    __syncwarp();
    mbar_wait(&bar_o[0], j & 1);
    tc_fence_after();
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_O + lane_off + c * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) o_acc[c * 32 + i] += __uint_as_float(v[i]);
    }
    tc_fence_before();  // order these TMEM reads before the next iteration's MMAs (after the next barrier)
  }

  if (qrow < seq) {
    const float inv = 1.f / l_run;
    __nv_bfloat16* dst = out + (static_cast<size_t>(b) * seq + qrow) * ld_out + head * 64;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      uint4 u;
      u.x = pack_bf16(o_acc[8 * g] * inv, o_acc[8 * g + 1] * inv);
      u.y = pack_bf16(o_acc[8 * g + 2] * inv, o_acc[8 * g + 3] * inv);
      u.z = pack_bf16(o_acc[8 * g + 4] * inv, o_acc[8 * g + 5] * inv);
      u.w = pack_bf16(o_acc[8 * g + 6] * inv, o_acc[8 * g + 7] * inv);
      *reinterpret_cast<uint4*>(dst + 8 * g) = u;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace ovo
