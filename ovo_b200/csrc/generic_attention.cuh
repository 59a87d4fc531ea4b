// Softmax attention for head dims other than 64 (reference: F.scaled_dot_product_attention at pe.py:145-147 with the 2D
// RoPE of rope.py:303-347) — the "ViT-H/14-shaped" encoder of BASELINE config 4 has width 1280 / 16 heads = head_dim 80.
//
// The tcgen05 kernel of attention.cuh is built around 128-byte (64 x bf16) swizzled rows; this one is the portable
// flash-style kernel on mma.sync.m16n8k16 (same structure as the Hiera window attention, sam_attention.cuh): one CTA =
// (64 queries, image, head), 4 warps x 16 query rows, K/V staged 64 keys at a time.  It reads q/k/v straight from the
// f32 output of the QKV GEMM ([tokens, 3*width], in_proj order q|k|v, heads contiguous), applies the rotary embedding
// while staging (one bf16 rounding, after the rotation — as the fused QKV epilogue of the head_dim-64 path does) and
// handles sequence lengths that are not a multiple of the key block (577) and the causal mask per element.
#pragma once
#include "mma_sync.cuh"

namespace ovo {

struct GenAttnParams {
  const float* qkv;            // [n_seq*seq, 3*width] f32
  __nv_bfloat16* out;          // [n_seq*seq, width]
  int seq, heads, width;
  int causal;
  const float2* rope;          // [grid+1][HD/4] (cos, sin) of r*theta_i, or nullptr
  int rope_grid;               // token t > 0 sits at (y, x) = divmod(t-1, grid); t = 0 is the class token (angle 0)
  float scale_log2e;
};

template <int HD>
__global__ void __launch_bounds__(128) generic_attention_kernel(GenAttnParams p) {
  constexpr int HDP = (HD + 15) / 16 * 16;     // contraction length of Q.K^T (zero padded)
  constexpr int ROW = HDP + 8;                 // smem row stride in bf16: rows 16 B apart modulo 128 B
  constexpr int KS = HDP / 16, NTO = HD / 8, QUART = HD / 4;
  static_assert(HD % 8 == 0 && HD <= 128, "head_dim must be a multiple of 8 up to 128");
  __shared__ __align__(16) __nv_bfloat16 sQ[64 * ROW];
  __shared__ __align__(16) __nv_bfloat16 sK[64 * ROW];
  __shared__ __align__(16) __nv_bfloat16 sV[64 * ROW];
  griddep_launch();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int img = blockIdx.y, head = blockIdx.z;
  const int q0 = blockIdx.x * 64;
  const size_t ld = 3 * static_cast<size_t>(p.width);
  const float* base = p.qkv + static_cast<size_t>(img) * p.seq * ld + head * HD;

  // stage rows [t0, t0+64) of q (which = 0), k (1) or v (2): f32 pairs -> (rotate) -> bf16 pairs
  auto stage = [&](__nv_bfloat16* dst, int t0, int which) {
    const float* src = base + which * p.width;
    for (int it = tid; it < 64 * (HDP / 2); it += 128) {
      const int r = it / (HDP / 2), i = it - r * (HDP / 2);   // row, pair
      const int t = t0 + r;
      float a = 0.f, b = 0.f;
      if (t < p.seq && 2 * i < HD) {
        const float2 v = *reinterpret_cast<const float2*>(src + static_cast<size_t>(t) * ld + 2 * i);
        a = v.x; b = v.y;
        if (which < 2 && p.rope != nullptr && t > 0) {
          const int rr = (i < QUART ? (t - 1) % p.rope_grid : (t - 1) / p.rope_grid) + 1;
          const float2 cs = p.rope[rr * QUART + (i < QUART ? i : i - QUART)];
          const float ra = a * cs.x - b * cs.y, rb = b * cs.x + a * cs.y;
          a = ra; b = rb;
        }
      }
      *reinterpret_cast<uint32_t*>(dst + r * ROW + 2 * i) = pack_bf16(a, b);
    }
  };
  stage(sQ, q0, 0);
  __syncthreads();
  const bool active = q0 + warp * 16 < p.seq;
  uint32_t qf[KS][4];
  if (active) {
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) ldsm_x4(qf[ks], sQ + (warp * 16 + (lane & 15)) * ROW + ks * 16 + ((lane >> 4) << 3));
  }
  float o[NTO][4];
#pragma unroll
  for (int i = 0; i < NTO; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  const int g = lane >> 2, t4 = lane & 3;
  const int qa = q0 + warp * 16 + g, qb = qa + 8;           // the two query rows of this thread
  const int k_end = p.causal ? min(p.seq, q0 + 64) : p.seq; // keys past the tile's last query are never needed

  for (int kb = 0; kb < k_end; kb += 64) {
    __syncthreads();
    stage(sK, kb, 1);
    stage(sV, kb, 2);
    __syncthreads();
    if (!active) continue;
    float s[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        uint32_t bf[2];
        ldsm_x2(bf, sK + (nt * 8 + (lane & 7)) * ROW + ks * 16 + (((lane >> 3) & 1) << 3));
        mma_bf16_16816(s[nt], qf[ks], bf);
      }
      // padding keys and the causal mask, per element
      const int k0 = kb + nt * 8 + 2 * t4;
      if (k0 >= p.seq || (p.causal && k0 > qa)) s[nt][0] = -INFINITY;
      if (k0 + 1 >= p.seq || (p.causal && k0 + 1 > qa)) s[nt][1] = -INFINITY;
      if (k0 >= p.seq || (p.causal && k0 > qb)) s[nt][2] = -INFINITY;
      if (k0 + 1 >= p.seq || (p.causal && k0 + 1 > qb)) s[nt][3] = -INFINITY;
    }
    float bm0 = -INFINITY, bm1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      bm0 = fmaxf(bm0, fmaxf(s[nt][0], s[nt][1]));
      bm1 = fmaxf(bm1, fmaxf(s[nt][2], s[nt][3]));
    }
    bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1)); bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
    bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1)); bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
    const float nm0 = fmaxf(m0, bm0), nm1 = fmaxf(m1, bm1);
    const float c0 = (m0 == -INFINITY) ? 0.f : fast_ex2((m0 - nm0) * p.scale_log2e);
    const float c1 = (m1 == -INFINITY) ? 0.f : fast_ex2((m1 - nm1) * p.scale_log2e);
    m0 = nm0; m1 = nm1;
    const float ms0 = (m0 == -INFINITY) ? 0.f : m0 * p.scale_log2e, ms1 = (m1 == -INFINITY) ? 0.f : m1 * p.scale_log2e;
    float rs0 = 0.f, rs1 = 0.f;
    uint32_t pf[4][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float e0 = fast_ex2(fmaf(s[nt][0], p.scale_log2e, -ms0)), e1 = fast_ex2(fmaf(s[nt][1], p.scale_log2e, -ms0));
      const float e2 = fast_ex2(fmaf(s[nt][2], p.scale_log2e, -ms1)), e3 = fast_ex2(fmaf(s[nt][3], p.scale_log2e, -ms1));
      rs0 += e0 + e1; rs1 += e2 + e3;
      pf[nt >> 1][(nt & 1) * 2] = pack_bf16(e0, e1);
      pf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(e2, e3);
    }
    l0 = l0 * c0 + rs0; l1 = l1 * c1 + rs1;
#pragma unroll
    for (int i = 0; i < NTO; ++i) { o[i][0] *= c0; o[i][1] *= c0; o[i][2] *= c1; o[i][3] *= c1; }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)
#pragma unroll
      for (int nt = 0; nt < NTO; ++nt) {
        uint32_t bf[2];
        ldsm_x2_trans(bf, sV + (kk * 16 + (lane & 15)) * ROW + nt * 8);
        mma_bf16_16816(o[nt], pf[kk], bf);
      }
  }
  if (!active) return;
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int q = h ? qb : qa;
    if (q >= p.seq) continue;
    const float inv = 1.f / (h ? l1 : l0);
    __nv_bfloat16* dst = p.out + (static_cast<size_t>(img) * p.seq + q) * p.width + head * HD + 2 * t4;
#pragma unroll
    for (int nt = 0; nt < NTO; ++nt)
      *reinterpret_cast<uint32_t*>(dst + nt * 8) = pack_bf16(o[nt][2 * h] * inv, o[nt][2 * h + 1] * inv);
  }
}

}  // namespace ovo
