// Flash-style softmax attention forward for head_dim 64 on tcgen05 (reference: F.scaled_dot_product_attention
// at pe.py:145-147 for the vision tower, nn.MultiheadAttention with the causal mask pe.py:621-627 for text).
//
// One CTA = one (image, head, 128-query tile), 256 threads: two threads per query row, each owning half of the keys
// of a block and half of the output columns (the softmax is issue/latency bound, so warps per SM matter).  TWO CTAs
// are resident per SM (112 KB smem, 256 TMEM columns each): one CTA's softmax overlaps the other's tensor-core work.
// Q (128x64) is TMA-loaded once; K (128x64) and V (128x64, used as an MN-major B operand: no transposed copy of V is
// ever made) blocks stream through 2-slot rings of 128B-swizzled shared memory.  Per 128-key block: S = Q.K^T (tcgen05, accumulator in TMEM) -> tcgen05.ld -> online softmax in
// registers (one thread per query row, exp2f) -> P (bf16) written to swizzled smem -> O_j = P.V_j (tcgen05,
// issued together with the next block's Q.K^T) -> tcgen05.ld -> rescale-and-accumulate in registers.
// q/k/v are produced in exactly this layout by the QKV GEMM epilogue (gemm.cuh EPI_QKV).
#pragma once
#include "ptx.cuh"

namespace ovo {

constexpr int kAttnThreads = 256;
constexpr int kAttnMaxBlocks = 5;  // seq_pad <= 640

struct AttnSmem {
  static constexpr int kQ = 128 * 64 * 2;       // 16 KB
  static constexpr int kKBlock = 128 * 64 * 2;  // 16 KB per 128 keys
  static constexpr int kVBlock = 64 * 64 * 2;   // 8 KB per 64 keys (V rows of 128 B); one 128-key block = 2 of them, one TMA
  static constexpr int kP = 2 * 128 * 64 * 2;   // 32 KB: P as two K-major 128x64 tiles
  static constexpr int kBytes = kQ + 2 * kKBlock + 4 * kVBlock + kP + 256 + 512;  // 112 KB + barriers + row exchange
};

__global__ void __launch_bounds__(kAttnThreads, 2)
    attention_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                         const __grid_constant__ CUtensorMap tmV, __nv_bfloat16* __restrict__ out, int seq,
                         int seq_pad, int heads, int ld_out, float scale_log2e, int causal, int dbg) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();  // the 128B swizzle needs 1024-byte aligned tiles
  const int nblk = seq_pad / 128;
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + AttnSmem::kQ;                 // 2 slots
  uint8_t* sV = sK + 2 * AttnSmem::kKBlock;        // 2 slots x 2 tiles
  uint8_t* sP = sV + 4 * AttnSmem::kVBlock;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + AttnSmem::kP);
  uint64_t* bar_k = bars;       // [2]  (bar_k[0] also covers Q on its first use)
  uint64_t* bar_v = bars + 2;   // [2]
  uint64_t* bar_s = bars + 4;   // S ready
  uint64_t* bar_o = bars + 5;   // P.V ready
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
  float* s_x = reinterpret_cast<float*>(sP + AttnSmem::kP + 256);  // [128] per-row exchange between the two halves

  const int tid = threadIdx.x, warp = tid >> 5;
  const int qt = blockIdx.x;  // query tile
  const int bh = blockIdx.y;  // image * heads + head
  const int b = bh / heads, head = bh - b * heads;
  const int q0 = qt * 128;
  // causal: keys beyond the last query of this tile are never needed
  const int nb = causal ? min(nblk, qt + 1) : nblk;

  if (tid == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    for (int i = 0; i < 6; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<256>(tmem_slot);  // S [0,128)  O [128,192)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_S = *tmem_slot;
  const uint32_t tmem_O = tmem_S + 128;

  constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128);
  constexpr uint32_t idesc_o = umma_idesc_bf16_bmn(128, 64);
  const uint64_t qdesc = umma_desc_sw128(smem_u32(sQ));

  auto load_k = [&](int j) {  // K block j -> slot j&1
    const int s = j & 1;
    mbar_arrive_expect_tx(&bar_k[s], AttnSmem::kKBlock + (j == 0 ? AttnSmem::kQ : 0));
    if (j == 0) tma_load_2d(sQ, &tmQ, &bar_k[0], 0, bh * seq_pad + q0);
    tma_load_2d(sK + s * AttnSmem::kKBlock, &tmK, &bar_k[s], 0, bh * seq_pad + j * 128);
  };
  auto load_v = [&](int j) {  // V block j (128 keys x 64) -> slot j&1
    const int s = j & 1;
    mbar_arrive_expect_tx(&bar_v[s], 2 * AttnSmem::kVBlock);
    tma_load_2d(sV + (2 * s) * AttnSmem::kVBlock, &tmV, &bar_v[s], 0, bh * seq_pad + j * 128);
  };
  auto issue_qk = [&](int j) {  // S = Q . K_j^T
    mbar_wait(&bar_k[j & 1], (j >> 1) & 1);
    tc_fence_after();
    const uint64_t kdesc = umma_desc_sw128(smem_u32(sK + (j & 1) * AttnSmem::kKBlock));
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_bf16(tmem_S, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k != 0);
    umma_commit(bar_s);
  };

  if (tid == 0) {
    load_k(0);
    if (nb > 1) load_k(1);
    load_v(0);
    if (nb > 1) load_v(1);
    issue_qk(0);
  }

  // per-thread state: thread (row, half) owns query row `row`, keys [64*half, 64*half+64) of every block and output
  // columns [32*half, 32*half+32)
  const int row = tid & 127, half = tid >> 7;
  const int qrow = q0 + row;
  float m_run = -INFINITY, l_run = 0.f;  // l_run: partial row sum over this thread's keys (same m for both halves)
  float o_acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) o_acc[i] = 0.f;
  const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
  const int r8 = row & 7;
  uint8_t* p_row = sP + half * (128 * 128) + (row >> 3) * 1024 + r8 * 128;  // P tile `half` = this thread's 64 keys

  for (int j = 0; j < nb; ++j) {
    mbar_wait(bar_s, j & 1);
    tc_fence_after();
    // K slot j&1 has been consumed by Q.K_j^T: refill it with block j+2
    if (tid == 0 && j + 2 < nb) load_k(j + 2);

    const int kv0 = j * 128 + half * 64;             // first key of this thread's half
    int kv_hi = seq - kv0;                           // keys >= seq are padding
    if (causal) kv_hi = min(kv_hi, qrow - kv0 + 1);  // keys > query are masked
    // blocks that are entirely valid (all but the last one, and no causal diagonal) skip the per-element masking
    const bool full = !causal && (j * 128 + 128 <= seq);   // CTA uniform
    // pass 1: row max over this thread's 64 keys (independent partial maxima: no long dependent chain)
    float mp[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll 1
    for (int c = 0; c < ((dbg & 1) ? 0 : 2); ++c) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_S + lane_off + half * 64 + c * 32, v);
      tmem_ld_wait();
      if (full) {
#pragma unroll
        for (int i = 0; i < 32; ++i) mp[i & 3] = fmaxf(mp[i & 3], __uint_as_float(v[i]));
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (c * 32 + i < kv_hi) mp[i & 3] = fmaxf(mp[i & 3], __uint_as_float(v[i]));
      }
    }
    float m_blk = fmaxf(fmaxf(mp[0], mp[1]), fmaxf(mp[2], mp[3]));
    // both halves of a row must use the same maximum: half 1 publishes, half 0 combines and publishes back
    if (!(dbg & 8)) {
      if (half == 1) s_x[row] = m_blk;
      __syncthreads();
      if (half == 0) { m_blk = fmaxf(m_blk, s_x[row]); s_x[row] = m_blk; }
      __syncthreads();
      if (half == 1) m_blk = s_x[row];
    }
    const float m_new = fmaxf(m_run, m_blk);
    const float m_scaled = (m_new == -INFINITY) ? 0.f : m_new * scale_log2e;
    const float alpha = (m_run == -INFINITY) ? 0.f : fast_ex2(m_run * scale_log2e - m_scaled);
    // pass 2: p = exp2(s*scale - m), partial row sum, P -> swizzled smem (bf16)
    float lp[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
    for (int c = 0; c < ((dbg & 1) ? 0 : 2); ++c) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_S + lane_off + half * 64 + c * 32, v);
      tmem_ld_wait();
      float p[32];
      if (full) {
#pragma unroll
        for (int i = 0; i < 32; ++i) p[i] = fast_ex2(fmaf(__uint_as_float(v[i]), scale_log2e, -m_scaled));
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float e = fast_ex2(fmaf(__uint_as_float(v[i]), scale_log2e, -m_scaled));
          p[i] = (c * 32 + i < kv_hi) ? e : 0.f;
        }
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 u;
        u.x = pack_bf16(p[8 * g], p[8 * g + 1]); u.y = pack_bf16(p[8 * g + 2], p[8 * g + 3]);
        u.z = pack_bf16(p[8 * g + 4], p[8 * g + 5]); u.w = pack_bf16(p[8 * g + 6], p[8 * g + 7]);
        const int chunk = c * 4 + g;  // 16-byte chunk index inside the 128-byte row
        *reinterpret_cast<uint4*>(p_row + ((chunk ^ r8) << 4)) = u;
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) lp[i & 3] += p[i];
    }
    l_run = l_run * alpha + ((lp[0] + lp[1]) + (lp[2] + lp[3]));
    m_run = m_new;
#pragma unroll
    for (int i = 0; i < 32; ++i) o_acc[i] *= alpha;

    // P visible to the async proxy and every thread done reading S; then O_j = P.V_j and S_{j+1} = Q.K_{j+1}^T
    if (!(dbg & 4)) fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      mbar_wait(&bar_v[j & 1], (j >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const uint64_t pdesc = umma_desc_sw128(smem_u32(sP + t * (128 * 128)));
        // keys 64t + 16k .. +15 of the block: 16 rows of 128 B = 2048 B per K=16 step
        const uint64_t vdesc = umma_desc_sw128(smem_u32(sV + (2 * (j & 1) + t) * AttnSmem::kVBlock));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_O, pdesc + 2 * k, vdesc + 128 * k, idesc_o, (t | k) != 0);
      }
      umma_commit(bar_o);
      if (j + 1 < nb) issue_qk(j + 1);
    }
    __syncwarp();
    mbar_wait(bar_o, j & 1);
    tc_fence_after();
    // V slot j&1 has been consumed by P.V_j: refill it with block j+2
    if (tid == 0 && j + 2 < nb) load_v(j + 2);
    if (!(dbg & 2)) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_O + lane_off + half * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) o_acc[i] += __uint_as_float(v[i]);
    }
    tc_fence_before();  // these TMEM reads are ordered before the next P.V (issued after the next __syncthreads)
  }

  // total row sum = sum of the two halves' partial sums
  if (half == 1) s_x[row] = l_run;
  __syncthreads();
  if (half == 0) { l_run += s_x[row]; s_x[row] = l_run; }
  __syncthreads();
  if (half == 1) l_run = s_x[row];
  if (qrow < seq) {
    const float inv = 1.f / l_run;
    __nv_bfloat16* dst = out + (static_cast<size_t>(b) * seq + qrow) * ld_out + head * 64 + half * 32;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
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
    tmem_dealloc<256>(tmem_S);
  }
}

}  // namespace ovo
