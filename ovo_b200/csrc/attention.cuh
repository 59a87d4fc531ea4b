// Flash-style softmax attention forward for head_dim 64 on tcgen05 (reference: F.scaled_dot_product_attention
// at pe.py:145-147 for the vision tower, nn.MultiheadAttention with the causal mask pe.py:621-627 for text).
//
// One CTA = one (image, head, 128-query tile), 256 threads: two threads per query row, each owning half of the keys
// of a block and half of the output columns.  TWO CTAs are resident per SM (114 KB smem, 256 TMEM columns each): one
// CTA's softmax overlaps the other's tensor-core work.  Q (128x64) is TMA-loaded once; K (128x64) and V (128x64, used
// as an MN-major B operand: no transposed copy of V is ever made) blocks stream through 2-slot rings of 128B-swizzled
// shared memory.
//
// Per 128-key block: S = Q.K^T (tcgen05, accumulator in TMEM) -> ONE pass over S: tcgen05.ld, p = exp2(s*scale - m),
// P (bf16) -> swizzled smem -> O += P.V accumulated IN TMEM across all blocks (issued together with the next block's
// Q.K^T).  The reference maximum m of a row is fixed after the first block (the only block that needs a separate max
// pass) and only raised — with a rescale of the TMEM accumulator by tcgen05.ld/st — if a later block's maximum exceeds
// it by more than 2^kRescaleLog2: f32/bf16 carry an 8-bit exponent, so p up to 2^32 is exact enough and the final
// O / l division cancels the common factor.  Per block this removes a second pass over S, the per-block read-back of O
// and the cross-half maximum exchange (two block-wide barriers) of the textbook online softmax.
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
  // 112 KB + barriers.  Two CTAs per SM need 2 * (kBytes + 1 KB) <= 228 KB, so there is no room for a dedicated row-exchange
  // buffer: the rare cross-half exchanges (first block's maximum, a rescale, the final row sum) borrow the P tile while no
  // MMA reads it.
  static constexpr int kBytes = kQ + 2 * kKBlock + 4 * kVBlock + kP + 256;
};

constexpr float kRescaleLog2 = 32.f;   // raise a row's reference maximum only when a block exceeds it by 2^32

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
  uint64_t* bar_o = bars + 5;   // P.V done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
  float* s_x = reinterpret_cast<float*>(sP);   // [half][128] scratch inside the P tile (only while no P.V is in flight)

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

  // thread (row, half) owns query row `row`, keys [64*half, 64*half+64) of every block and output columns
  // [32*half, 32*half+32) of the accumulator in TMEM
  const int row = tid & 127, half = tid >> 7;
  const int qrow = q0 + row;
  float m_used = -INFINITY;      // the row's reference maximum (raw score units), identical in both halves
  float l_run = 0.f;             // partial row sum over this thread's keys
  float m_loc = -INFINITY;       // maximum this thread has seen over its keys so far
  int rescale = 0;               // CTA-uniform: some row's maximum outgrew its reference by 2^kRescaleLog2 in the last block
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

    if (j == 0) {
      // first block: the reference maximum of the row = its maximum over block 0 (both halves)
      float mp[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
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
      const float mine = fmaxf(fmaxf(mp[0], mp[1]), fmaxf(mp[2], mp[3]));
      s_x[half * 128 + row] = mine;
      __syncthreads();
      m_used = fmaxf(mine, s_x[(1 - half) * 128 + row]);
      m_loc = mine;
      __syncthreads();                               // the scratch lives in the P tile that is written next
    } else {
      // P.V_{j-1} has completed (it was issued before Q.K_j^T): the P tile, V slot (j-1)&1 and the accumulator are free
      mbar_wait(bar_o, (j - 1) & 1);
      tc_fence_after();
      if (tid == 0 && j + 1 < nb) load_v(j + 1);
      if (rescale) {   // rare, CTA-uniform (dbg 16: whenever a maximum grows, for the tests)
        s_x[half * 128 + row] = m_loc;
        __syncthreads();
        const float m_new = fmaxf(m_used, fmaxf(m_loc, s_x[(1 - half) * 128 + row]));   // the same in both halves of the row
        const float alpha = (m_new > m_used) ? fast_ex2((m_used - m_new) * scale_log2e) : 1.f;
        uint32_t v[32];
        tmem_ld_32x32(tmem_O + lane_off + half * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
        tmem_st_32x32(tmem_O + lane_off + half * 32, v);
        tmem_st_wait();
        l_run *= alpha;
        m_used = m_new;
        __syncthreads();                             // scratch reads done before P is written again
      }
    }
    const float m_scaled = (m_used == -INFINITY) ? 0.f : m_used * scale_log2e;

    // single pass: p = exp2(s*scale - m), partial row sum, block maximum, P -> swizzled smem (bf16)
    float lp[4] = {0.f, 0.f, 0.f, 0.f};
    float mp[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_S + lane_off + half * 64 + c * 32, v);
      tmem_ld_wait();
      float p[32];
      if (full) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float sv = __uint_as_float(v[i]);
          p[i] = fast_ex2(fmaf(sv, scale_log2e, -m_scaled));
          mp[i & 3] = fmaxf(mp[i & 3], sv);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float sv = __uint_as_float(v[i]);
          const bool ok = c * 32 + i < kv_hi;
          p[i] = ok ? fast_ex2(fmaf(sv, scale_log2e, -m_scaled)) : 0.f;
          if (ok) mp[i & 3] = fmaxf(mp[i & 3], sv);
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
    l_run += (lp[0] + lp[1]) + (lp[2] + lp[3]);
    m_loc = fmaxf(m_loc, fmaxf(fmaxf(mp[0], mp[1]), fmaxf(mp[2], mp[3])));
    const int need = m_loc > m_used + ((dbg & 16) ? 0.f : kRescaleLog2 / scale_log2e);

    // P visible to the async proxy and every thread done reading S; then O += P.V_j and S_{j+1} = Q.K_{j+1}^T
    fence_proxy_async_smem();
    tc_fence_before();
    rescale = __syncthreads_or(need);
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
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_O, pdesc + 2 * k, vdesc + 128 * k, idesc_o, (j | t | k) != 0);
      }
      umma_commit(bar_o);
      if (j + 1 < nb) issue_qk(j + 1);
    }
  }

  // total row sum = sum of the two halves' partial sums; the accumulator is complete once the last P.V has landed
  mbar_wait(bar_o, (nb - 1) & 1);                    // ... and the P tile is free to serve as scratch
  tc_fence_after();
  s_x[half * 128 + row] = l_run;
  __syncthreads();
  l_run += s_x[(1 - half) * 128 + row];
  uint32_t v[32];
  tmem_ld_32x32(tmem_O + lane_off + half * 32, v);   // warp-collective: every lane, also the padding rows
  tmem_ld_wait();
  if (qrow < seq) {
    const float inv = 1.f / l_run;
    __nv_bfloat16* dst = out + (static_cast<size_t>(b) * seq + qrow) * ld_out + head * 64 + half * 32;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      uint4 u;
      u.x = pack_bf16(__uint_as_float(v[8 * g]) * inv, __uint_as_float(v[8 * g + 1]) * inv);
      u.y = pack_bf16(__uint_as_float(v[8 * g + 2]) * inv, __uint_as_float(v[8 * g + 3]) * inv);
      u.z = pack_bf16(__uint_as_float(v[8 * g + 4]) * inv, __uint_as_float(v[8 * g + 5]) * inv);
      u.w = pack_bf16(__uint_as_float(v[8 * g + 6]) * inv, __uint_as_float(v[8 * g + 7]) * inv);
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
