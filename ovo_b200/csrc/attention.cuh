// Flash-style softmax attention forward for head_dim 64 on tcgen05 (reference: F.scaled_dot_product_attention
// at pe.py:145-147 for the vision tower, nn.MultiheadAttention with the causal mask pe.py:621-627 for text).
//
// One CTA = one (image, head, 128-query tile), 256 softmax threads: two threads per query row, each owning half of the keys
// of a block and half of the output columns (+ one MMA warp and one TMA producer warp, below).  TWO CTAs are resident per SM (84 KB smem, 256 TMEM columns each).
// Q (128x64) is TMA-loaded once; K and V stream in 64-key blocks through 4-slot rings of 128B-swizzled shared memory
// (V is consumed as an MN-major B operand: no transposed copy of V is ever made).  P never touches shared memory: the softmax
// threads store it as bf16 pairs into tensor memory (tcgen05.st, 16 columns per thread) and P.V takes its A operand from there
// (tcgen05.mma with a tensor-memory A operand).
//
// Software pipeline over 64-key blocks with TWO score buffers in TMEM (S0, S1: 64 columns each) and two P buffers (32 columns each):
//     tensor core : ... P.V_{j-1} | Q.K_{j+1}^T |          P.V_j | Q.K_{j+2}^T | ...
//     softmax     :        block j (reads S_j, writes P_j)  |  block j+1 (S_{j+1} was issued one block earlier) ...
// so the MMA -> tcgen05.ld -> exp2 -> P -> MMA dependency chain of a single-buffered loop is broken: while the threads
// do the softmax of block j+1 the tensor core finishes P.V_j and computes S_{j+2}.
// The output accumulator O (128x64 f32) stays in TMEM across all blocks (P.V accumulates in place).  A row's reference
// maximum m is fixed after the first block (the only one that needs a separate max pass) and is only raised — with a
// tcgen05.ld/st rescale of the accumulator — if a later block exceeds it by more than 2^kRescaleLog2: f32/bf16 carry an
// 8-bit exponent, so p up to 2^32 loses nothing and the final O / l division cancels the common factor.
// Warp specialisation: the 256 softmax threads issue neither MMAs nor loads — a ninth warp waits on the "P tile ready"
// mbarrier and issues P.V and the Q.K^T two blocks ahead; a tenth warp is the TMA producer: it refills a K / V ring slot as soon
// as the MMA that read it has completed (full/empty mbarrier pairs per slot) and starts the next work item's loads while the
// softmax threads are still in the current item's epilogue.  (A clock64 trace of one CTA, tools/attn_trace.py, showed that a
// TMA issue costs the issuing thread 150-200 cycles; with softmax thread 0 doing them, 350 of the ~1850 cycles of every block
// were spent there while the other 255 threads waited at the block's barrier.)  The softmax threads synchronise among
// themselves on a named barrier (bar.red.or also carries the rare rescale request).
// The kernel is PERSISTENT: 2 CTAs per SM each loop over (image, head, query tile) work items, so barrier set-up, the TMEM
// allocation and the tensor-map fetch are paid once per CTA (measured: a third of the non-persistent kernel's time was
// per-CTA fixed cost) and the next item's Q/K/V loads are in flight while the current item finishes.
// q/k/v are produced in exactly this layout by the QKV GEMM epilogue (gemm.cuh EPI_QKV).
#pragma once
#include "ptx.cuh"

namespace ovo {

constexpr int kAttnThreads = 320;   // warps 0..7: softmax (two threads per query row); warp 8: tcgen05.mma issue; warp 9: TMA producer
constexpr int kAttnMaxBlocks = 5;  // seq_pad <= 640 (host-side limit of the q/k/v buffers, in 128-query tiles)

constexpr int kAttnKB = 64;        // keys per block

struct AttnSmem {
  static constexpr int kQ = 128 * 64 * 2;       // 16 KB
  static constexpr int kKBlock = 64 * 64 * 2;   // 8 KB per 64 keys, 4 slots
  static constexpr int kVBlock = 64 * 64 * 2;   // 8 KB per 64 keys (V rows of 128 B), 4 slots
  static constexpr int kP = 2048;               // 2 KB of exchange scratch (2 x 256 floats), 2 buffers.  P itself lives in tensor memory
  // 84 KB + barriers; two CTAs per SM (tensor memory: 2 x 256 columns).  The scratch carries the rare cross-half exchanges (first
  // block's maximum and tail-key dot product, a rescale, the final row sum).
  static constexpr int kBytes = kQ + 4 * kKBlock + 4 * kVBlock + 2 * kP + 256;   // 24 mbarriers + the TMEM slot in the last 256 B
};

constexpr float kRescaleLog2 = 32.f;   // raise a row's reference maximum only when a block exceeds it by 2^32

__global__ void __launch_bounds__(kAttnThreads, 2)
    attention_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                         const __grid_constant__ CUtensorMap tmV, __nv_bfloat16* __restrict__ out, int seq,
                         int seq_pad, int heads, int ld_out, float scale_log2e, int causal, int dbg, int qtiles, int n_items,
                         const __nv_bfloat16* __restrict__ k_glob, const __nv_bfloat16* __restrict__ v_glob, int tail1,
                         long long* __restrict__ trace /* measurement tap (ovo_attn_trace): clock64 stamps of CTA 0, else null */) {
  griddep_launch();
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();  // the 128B swizzle needs 1024-byte aligned tiles
  const int nblk = seq_pad / kAttnKB;
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + AttnSmem::kQ;                 // 4 slots
  uint8_t* sV = sK + 4 * AttnSmem::kKBlock;        // 4 slots
  uint8_t* sP = sV + 4 * AttnSmem::kVBlock;        // 2 buffers
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * AttnSmem::kP);
  uint64_t* bar_k = bars;        // [4]  (bar_k[0] also covers Q on its first use)
  uint64_t* bar_v = bars + 4;    // [4]
  uint64_t* bar_s = bars + 8;    // [2]  S buffer ready
  uint64_t* bar_o = bars + 10;   // [2]  P.V of a block with this parity done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // provably warp-uniform: the MMA warp's control flow and descriptors stay
                                                            // on the uniform datapath (see the MMA warp below)
  // trace[(item * 16 + block) * 16 + phase]: softmax thread 0 writes phases 0..5, the MMA thread 6..12; block 15 = item level
  const bool tr0 = trace != nullptr && blockIdx.x == 0 && tid == 0, tr8 = trace != nullptr && blockIdx.x == 0 && tid == 256;
#define ATTN_STAMP(on, blk, ph) do { if (on) trace[((n_done & 7) * 16 + (blk)) * 16 + (ph)] = clock64(); } while (0)
  uint64_t* bar_q = bars + 13;   // Q tile of the current item landed
  uint64_t* bar_p = bars + 14;   // [2] P tile written and S buffer consumed by every softmax thread
  uint64_t* bar_kfree = bars + 16;   // [4] K slot consumed by its Q.K^T (tcgen05.commit): the producer may refill it
  uint64_t* bar_vfree = bars + 20;   // [4] V slot consumed by its P.V
  // rescale requests, per warp pair and block (4 slots): the global index of the block in which some row of the pair outgrew
  // its reference maximum; read two blocks later (block indices never repeat, so the slots are never cleared)
  volatile int* s_flag = reinterpret_cast<volatile int*>(bars + 24) + (warp & 3) * 4;
  if (tid == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    for (int i = 0; i < 12; ++i) mbar_init(&bars[i], 1);
    mbar_init(bar_q, 1);
    mbar_init(&bar_p[0], 8);      // one arrival per softmax warp
    mbar_init(&bar_p[1], 8);
    for (int i = 0; i < 16; ++i) reinterpret_cast<volatile int*>(bars + 24)[i] = -1;
    for (int i = 0; i < 4; ++i) { mbar_init(&bar_kfree[i], 1); mbar_init(&bar_vfree[i], 1); }
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc<256>(tmem_slot);  // S0 [0,64)  S1 [64,128)  O [128,192)  P0 [192,224)  P1 [224,256)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_S = *tmem_slot;
  const uint32_t tmem_O = tmem_S + 128;
  const uint32_t tmem_P = tmem_S + 192;   // P0 [192,224)  P1 [224,256): bf16 pairs, the A operand of P.V (never in shared memory)

  constexpr uint32_t idesc_s = umma_idesc_bf16(128, 64);
  constexpr uint32_t idesc_o = umma_idesc_bf16_bmn(128, 64);
  const uint64_t qdesc = umma_desc_sw128(smem_u32(sQ));

  // Ring slots and mbarrier parities are driven by `g`, the number of key blocks this CTA has processed so far plus the block
  // index inside the current item, so the rings keep rolling across work items.
  int n_done = 0;                       // work items this CTA has finished (parity of bar_q)
  int bh = 0, q0 = 0, nb = 0, g0 = 0;   // current item: (image*heads + head), first query, key blocks, global index of its block 0
  auto load_k = [&](int j) {  // K block j of the item -> slot (g0+j)&3
    const int s = (g0 + j) & 3;
    mbar_arrive_expect_tx(&bar_k[s], AttnSmem::kKBlock);
    tma_load_2d(sK + s * AttnSmem::kKBlock, &tmK, &bar_k[s], 0, bh * seq_pad + j * kAttnKB);
  };
  auto load_v = [&](int j) {  // V block j of the item -> slot (g0+j)&3
    const int s = (g0 + j) & 3;
    mbar_arrive_expect_tx(&bar_v[s], AttnSmem::kVBlock);
    tma_load_2d(sV + s * AttnSmem::kVBlock, &tmV, &bar_v[s], 0, bh * seq_pad + j * kAttnKB);
  };
  auto issue_qk = [&](int j) {  // S_{g&1} = Q . K_j^T   (called by every lane of the MMA warp; one elected lane issues)
    const int g = g0 + j;
    mbar_wait(&bar_k[g & 3], (g >> 2) & 1);
    tc_fence_after();
    ATTN_STAMP(tr8, j - 2 >= 0 ? j - 2 : 14, 11);
    const uint64_t kdesc = umma_desc_sw128(smem_u32(sK + (g & 3) * AttnSmem::kKBlock));
    if (elect_one()) {
      if (!(dbg & 2)) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_S + (g & 1) * 64, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k != 0);
      }
      umma_commit(&bar_s[g & 1]);
      umma_commit(&bar_kfree[g & 3]);
    }
    __syncwarp();
    ATTN_STAMP(tr8, j - 2 >= 0 ? j - 2 : 14, 12);
  };

  const int row = tid & 127, half = tid >> 7;
  const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
  const int r8 = row & 7;                           // position inside an 8-row swizzle group of the Q tile (tail-key dot product)

  for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n_done) {
  const int qt = item % qtiles;
  bh = item / qtiles;
  const int b = bh / heads, head = bh - b * heads;
  q0 = qt * 128;
  // causal: keys beyond the last query of this tile are never needed; blocks of pure padding are skipped
  // (dbg bits are measurement aids, results are wrong with them: 1 = no softmax math / P stores, 2 = no MMA, 4 = one key block)
  // tail1 (seq % 64 == 1, e.g. 24*24 patches + the class token = 577): the single key of the last block does not get a
  // 64-key block of its own — the softmax threads take it as a rank-1 term (q.k_last on the FMA pipe, p_last * v_last added
  // to the accumulator in the epilogue), which removes one of ten block iterations of every work item.
  nb = (dbg & 4) ? 1 : min(causal ? min(nblk, 2 * qt + 2) : nblk, (seq - tail1 + kAttnKB - 1) / kAttnKB);

  if (warp == 9) {
    // ---------------------------------------------------------------- TMA producer warp (one lane)
    // Full/empty rings: a slot's "free" barrier completes once per use (committed by the MMA thread behind the MMAs that read
    // it) and cannot complete again before this thread has refilled the slot, so a parity wait can never fall a phase behind.
    if (tid == 288) {
      if (n_done > 0) {
        // the previous item's last P.V (hence every MMA of that item: one thread's MMAs complete in order) has read Q and the rings
        const int gl = g0 - 1;
        mbar_wait(&bar_vfree[gl & 3], (gl >> 2) & 1);
      }
      mbar_arrive_expect_tx(bar_q, AttnSmem::kQ);
      tma_load_2d(sQ, &tmQ, bar_q, 0, bh * seq_pad + q0);
      for (int j = 0; j < 4 && j < nb; ++j) load_k(j);
      for (int j = 0; j < 4 && j < nb; ++j) load_v(j);
      for (int j = 0; j + 4 < nb; ++j) {
        const int g = g0 + j;
        mbar_wait(&bar_kfree[g & 3], (g >> 2) & 1);
        load_k(j + 4);
        mbar_wait(&bar_vfree[g & 3], (g >> 2) & 1);
        load_v(j + 4);
      }
    }
    __syncwarp();
    g0 += nb;
    continue;
  }
  if (warp == 8) {
    // ---------------------------------------------------------------- MMA warp
    // Every lane runs the (warp-uniform) control flow and the barrier waits; ONE elected lane issues the tcgen05 instructions.
    // With the whole section under `if (tid == 256)` the compiler could not prove the descriptors uniform and moved each one
    // into the uniform registers through an ELECT / R2UR.BROADCAST loop: ~100 cycles per tcgen05.mma, 1300-1500 cycles of
    // issue per key block (clock64 trace, tools/attn_trace.py) — as long as the softmax of a block.
    // S buffers: consumed (bar_p of the previous item's last two blocks was waited for); the accumulator is overwritten only
    // by P.V_0, which waits for the softmax threads' first bar_p arrival of this item, i.e. for the end of their previous epilogue
    mbar_wait(bar_q, n_done & 1);
    issue_qk(0);
    if (nb > 1) issue_qk(1);
    for (int j = 0; j < nb; ++j) {
      const int g = g0 + j;
      mbar_wait(&bar_p[g & 1], (g >> 1) & 1);      // P_j written (tensor memory), S buffer g&1 consumed
      ATTN_STAMP(tr8, j, 6);
      tc_fence_after();
      mbar_wait(&bar_v[g & 3], (g >> 2) & 1);
      tc_fence_after();
      ATTN_STAMP(tr8, j, 8);
      // keys 16k .. 16k+15 of the block: 16 rows of 128 B = 2048 B per K=16 step
      const uint64_t vdesc = umma_desc_sw128(smem_u32(sV + (g & 3) * AttnSmem::kVBlock));
      if (elect_one()) {
        if (!(dbg & 2)) {              // P from tensor memory: 8 columns (16 keys) per K step
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ts(tmem_O, tmem_P + (g & 1) * 32 + 8 * k, vdesc + 128 * k, idesc_o, (j | k) != 0);
        }
        umma_commit(&bar_o[g & 1]);
        umma_commit(&bar_vfree[g & 3]);
      }
      __syncwarp();
      ATTN_STAMP(tr8, j, 10);
      if (j + 2 < nb) issue_qk(j + 2);
      ATTN_STAMP(tr8, j, 7);
    }
    g0 += nb;
    continue;
  }

  // thread (row, half) owns query row `row`, keys [32*half, 32*half+32) of every block and output columns
  // [32*half, 32*half+32) of the accumulator in TMEM
  const int qrow = q0 + row;
  float m_used = -INFINITY;      // the row's reference maximum (raw score units), identical in both halves
  float l_run = 0.f;             // partial row sum over this thread's keys
  float m_loc = -INFINITY;       // maximum this thread has seen over its keys so far
  int rescale = 0;               // uniform over the warp pair: some row of the pair outgrew its reference by 2^kRescaleLog2 two blocks ago
  // The two threads of a row sit in warps w and w + 4 (same TMEM lane quarter).  They synchronise only with each other (a 64-thread
  // named barrier for the first block's maximum, the final row sum and the rare rescale); no barrier spans the CTA: every warp
  // arrives on the "P ready" mbarrier by itself, so the eight softmax warps are free to drift apart.
  const int pair_bar = 1 + (warp & 3);
  float s_tail = -INFINITY;      // raw score of the tail key (tail1), identical in both halves
  // this half's 32 dims of the tail key / 32 output columns of its V row (bf16, 64 B each): pulled into L1 now, read when used
  const uint4* kt_ptr = reinterpret_cast<const uint4*>(k_glob + (static_cast<size_t>(bh) * seq_pad + (seq - 1)) * 64 + half * 32);
  const uint4* vt_ptr = reinterpret_cast<const uint4*>(v_glob + (static_cast<size_t>(bh) * seq_pad + (seq - 1)) * 64 + half * 32);
  if (tail1 && (tid & 31) == 0) {
    asm volatile("prefetch.global.L1 [%0];" ::"l"(kt_ptr));
    asm volatile("prefetch.global.L1 [%0];" ::"l"(vt_ptr));
  }

  ATTN_STAMP(tr0, 15, 0);
  for (int j = 0; j < nb; ++j) {
    const int g = g0 + j;                              // global block index of this CTA: ring slots and parities
    ATTN_STAMP(tr0, j, 0);
    mbar_wait(&bar_s[g & 1], (g >> 1) & 1);
    ATTN_STAMP(tr0, j, 1);
    tc_fence_after();
    const uint32_t s_addr = tmem_S + lane_off + (g & 1) * 64 + half * 32;
    float* s_x = reinterpret_cast<float*>(sP + (g & 1) * AttnSmem::kP);   // exchange scratch

    const int kv0 = j * kAttnKB + half * 32;         // first key of this thread's half
    int kv_hi = seq - kv0;                           // keys >= seq are padding
    if (causal) kv_hi = min(kv_hi, qrow - kv0 + 1);  // keys > query are masked
    // blocks that are entirely valid (no padding, no causal diagonal) skip the per-element masking
    const bool full = !causal && (j * kAttnKB + kAttnKB <= seq);   // CTA uniform

    uint32_t v[32];
    tmem_ld_32x32(s_addr, v);
    tmem_ld_wait();
    ATTN_STAMP(tr0, j, 2);

    if (j == 0) {
      // first block: the reference maximum of the row = its maximum over block 0 (both halves)
      float mp[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      if (full) {
#pragma unroll
        for (int i = 0; i < 32; ++i) mp[i & 3] = fmaxf(mp[i & 3], __uint_as_float(v[i]));
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (i < kv_hi) mp[i & 3] = fmaxf(mp[i & 3], __uint_as_float(v[i]));
      }
      const float mine = fmaxf(fmaxf(mp[0], mp[1]), fmaxf(mp[2], mp[3]));
      s_x[half * 128 + row] = mine;
      float dot = 0.f;
      if (tail1) {
        // Q landed before S_0 was computed; the wait (already complete) orders this thread's generic-proxy reads after the TMA
        mbar_wait(bar_q, n_done & 1);
        float d4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint4 qv = *reinterpret_cast<const uint4*>(sQ + row * 128 + (((half * 4 + c) ^ r8) << 4));
          const uint4 kv = __ldg(kt_ptr + c);
          const uint32_t qa[4] = {qv.x, qv.y, qv.z, qv.w}, ka[4] = {kv.x, kv.y, kv.z, kv.w};
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            d4[w] = fmaf(__uint_as_float(qa[w] << 16), __uint_as_float(ka[w] << 16), d4[w]);
            d4[w] = fmaf(__uint_as_float(qa[w] & 0xffff0000u), __uint_as_float(ka[w] & 0xffff0000u), d4[w]);
          }
        }
        dot = (d4[0] + d4[1]) + (d4[2] + d4[3]);
        s_x[256 + half * 128 + row] = dot;
      }
      named_bar_sync(pair_bar, 64);
      m_used = fmaxf(mine, s_x[(1 - half) * 128 + row]);
      if (tail1) {
        s_tail = dot + s_x[256 + (1 - half) * 128 + row];   // a + b == b + a: the same value in both halves
        m_used = fmaxf(m_used, s_tail);
      }
      m_loc = mine;   // (a thread's scratch slot is next written in a rescale at block >= 2 or in the epilogue: S of that block exists
                      // only after every warp arrived for block 0, i.e. after the partner's read above)
    } else {
      // P buffer g&1 is free again without a wait of its own: S_j is ready (waited for above), so Q.K_j^T has completed, and
      // with it every MMA its thread issued earlier (tcgen05.mma execute in issue order) — P.V_{j-2}, the buffer's last reader,
      // was issued before Q.K_j^T.  The same chain (the pair's writes of block j-2 -> bar_p -> MMA warp -> commit -> bar_s) makes the
      // pair's rescale request of block j-2 visible here.
      if (j >= 2) rescale = (s_flag[(g - 2) & 3] == g - 2);
      if (rescale) {   // rare, uniform over the warp pair (dbg 16: whenever a maximum grows, for the tests)
        rescale = 0;
        mbar_wait(&bar_o[(g - 1) & 1], ((g - 1) >> 1) & 1);    // every P.V so far has landed: the accumulator is stable
        tc_fence_after();                                     // (P.V_j cannot start before this warp has arrived for block j)
        s_x[half * 128 + row] = m_loc;
        named_bar_sync(pair_bar, 64);
        const float m_new = fmaxf(m_used, fmaxf(m_loc, s_x[(1 - half) * 128 + row]));   // the same in both halves of the row
        const float alpha = (m_new > m_used) ? fast_ex2((m_used - m_new) * scale_log2e) : 1.f;
        uint32_t o[32];
        tmem_ld_32x32(tmem_O + lane_off + half * 32, o);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
        tmem_st_32x32(tmem_O + lane_off + half * 32, o);
        tmem_st_wait();
        l_run *= alpha;
        m_used = m_new;
        named_bar_sync(pair_bar, 64);                       // scratch reads done before the slot can be written again
      }
    }
    const float m_scaled = (m_used == -INFINITY) ? 0.f : m_used * scale_log2e;
    ATTN_STAMP(tr0, j, 3);

    // p = exp2(s*scale - m), partial row sum, block maximum, P -> swizzled smem (bf16)
    float lp[4] = {0.f, 0.f, 0.f, 0.f};
    float mp[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    float p[32];
    if (dbg & 1) {
#pragma unroll
      for (int i = 0; i < 32; ++i) p[i] = 0.f;
    } else if (full) {
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
        const bool ok = i < kv_hi;
        p[i] = ok ? fast_ex2(fmaf(sv, scale_log2e, -m_scaled)) : 0.f;
        if (ok) mp[i & 3] = fmaxf(mp[i & 3], sv);
      }
    }
    if (dbg & 1) {
    } else {                // P -> tensor memory: this thread's 32 keys = 16 columns of bf16 pairs in its own lane
      uint32_t pw[16];
#pragma unroll
      for (int w = 0; w < 16; ++w) pw[w] = pack_bf16(p[2 * w], p[2 * w + 1]);
      tmem_st_32x16(tmem_P + lane_off + (g & 1) * 32 + half * 16, pw);
      tmem_st_wait();
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) lp[i & 3] += p[i];
    l_run += (lp[0] + lp[1]) + (lp[2] + lp[3]);
    m_loc = fmaxf(m_loc, fmaxf(fmaxf(mp[0], mp[1]), fmaxf(mp[2], mp[3])));
    const int need = m_loc > m_used + ((dbg & 16) ? 0.f : kRescaleLog2 / scale_log2e);

    // P_j stored to tensor memory (tcgen05.wait::st above) and every thread done reading S_j; then O += P_j.V_j and S buffer j&1 <- Q.K_{j+2}^T
    ATTN_STAMP(tr0, j, 4);
    if (need) s_flag[g & 3] = g;      // picked up by both warps of the pair at block j + 2
    tc_fence_before();
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(&bar_p[g & 1]);   // one of 8 arrivals: this warp's P rows are stored, its S rows consumed
    ATTN_STAMP(tr0, j, 5);
  }
  ATTN_STAMP(tr0, 15, 1);

  // the accumulator is complete once the last P.V has landed (tcgen05.mma of one thread complete in order);
  // total row sum = sum of the two halves' partial sums
  mbar_wait(&bar_o[(g0 + nb - 1) & 1], ((g0 + nb - 1) >> 1) & 1);
  tc_fence_after();
  ATTN_STAMP(tr0, 15, 2);
  float* s_x = reinterpret_cast<float*>(sP);
  s_x[half * 128 + row] = l_run;
  named_bar_sync(pair_bar, 64);
  l_run += s_x[(1 - half) * 128 + row];
  uint32_t v[32];
  tmem_ld_32x32(tmem_O + lane_off + half * 32, v);   // warp-collective: every lane, also the padding rows
  tmem_ld_wait();
  if (tail1) {   // the tail key against the row's FINAL reference maximum: l += p, O += p * v_last
    const float pt = fast_ex2(fmaf(s_tail, scale_log2e, -m_used * scale_log2e));
    l_run += pt;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const uint4 vv = __ldg(vt_ptr + c);
      const uint32_t va[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        v[8 * c + 2 * w] = __float_as_uint(fmaf(pt, __uint_as_float(va[w] << 16), __uint_as_float(v[8 * c + 2 * w])));
        v[8 * c + 2 * w + 1] = __float_as_uint(fmaf(pt, __uint_as_float(va[w] & 0xffff0000u), __uint_as_float(v[8 * c + 2 * w + 1])));
      }
    }
  }
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
  // the partner has read this thread's scratch slot before it is written again (next item's block 0); this warp's accumulator
  // reads precede its first bar_p arrival of the next item, which P.V_0 (the first MMA to overwrite the accumulator) waits for
  tc_fence_before();
  named_bar_sync(pair_bar, 64);
  tc_fence_after();
  ATTN_STAMP(tr0, 15, 3);
  g0 += nb;
  }  // work items
#undef ATTN_STAMP

  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_S);
  }
}

}  // namespace ovo
