// Windowed / global multi-head attention of the Hiera trunk (reference: thirdParty/segment-anything-2/sam2/modeling/
// backbones/hieradet.py:56-81 MultiScaleAttention.forward, window partition backbones/utils.py:16-63) for head_dim 72.
//
// The trunk keeps tokens in row-major spatial order; a window is addressed by index arithmetic, so window_partition /
// window_unpartition never materialise.  Q-pooling blocks (MaxPool2d(2,2) on q, hieradet.py:66-70) take the element-wise
// max of the four source tokens while the Q tile is staged.
//
// Flash-style: one CTA = (64 queries, window, head), 4 warps x 16 query rows; K/V staged 64 keys at a time in shared
// memory, S = Q K^T and O += P V on mma.sync.m16n8k16 bf16 (head_dim 72 is padded to 80 for the QK^T contraction and
// uses 9 n-tiles for PV), online softmax in registers.  Windows here have 16..4096 keys: at these shapes (16-64-256
// token windows, head_dim 72) a tcgen05 tile would be mostly padding; the trunk's FLOPs are in its GEMMs (gemm.cuh).
#pragma once
#include "mma_sync.cuh"

namespace ovo {

constexpr int kSamHd = 72;
constexpr int kSamHdPad = 80;
constexpr int kSamRow = 88;   // smem row stride in bf16 (176 B: 8 consecutive rows hit 8 distinct 16-B bank groups)
constexpr int kSamKB = 64;    // keys per block
constexpr int kSamQB = 64;    // queries per CTA

struct WinAttnParams {
  const __nv_bfloat16* qkv;   // [grid*grid, 3*dim_out]
  __nv_bfloat16* out;         // [grid_out*grid_out, dim_out]
  int grid;                   // input token grid side
  int ws;                     // window side on the input grid (== grid for global attention)
  int heads, dim_out;
  int q_pool;                 // 1: queries are 2x2 max-pooled
  float scale_log2e;          // head_dim^-0.5 * log2(e)
  int wins2;                  // windows per image: blockIdx.y = image * wins2 + window (several images per launch)
};

__device__ __forceinline__ uint4 bf16x8_max(uint4 a, uint4 b) {
  uint4 r;
  const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
  const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b);
  __nv_bfloat162* pr = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) pr[i] = __hmax2(pa[i], pb[i]);
  return r;
}

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// QB queries per CTA (QB/16 warps); K/V blocks of 64 keys double-buffered with cp.async so the next block's global loads
// overlap this block's MMAs.  Dynamic shared memory: Q [QB] + 2 stages x (K [64] + V [64]) rows of kSamRow bf16.
template <int QB>
constexpr int hiera_attn_smem_bytes() { return (QB + 4 * kSamKB) * kSamRow * 2; }

template <int QB>
__global__ void __launch_bounds__(QB * 2) hiera_attention_kernel(WinAttnParams p) {
  griddep_launch();
  extern __shared__ __align__(16) uint8_t hiera_smem[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(hiera_smem);
  __nv_bfloat16* sKV = sQ + QB * kSamRow;                 // stage s: K at sKV + s*2*64*row, V right after K
  constexpr int kThreads = QB * 2;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int head = blockIdx.z;
  const int wins = p.grid / p.ws;
  const int img = blockIdx.y / p.wins2, win = blockIdx.y - img * p.wins2;
  const int wy = win / wins, wx = win - wy * wins;
  const int nk = p.ws * p.ws;
  const int wsq = p.q_pool ? p.ws >> 1 : p.ws;        // window side on the query / output grid
  const int nq = wsq * wsq;
  const int grid_out = p.q_pool ? p.grid >> 1 : p.grid;
  const int q0 = blockIdx.x * QB;
  const int ld = 3 * p.dim_out;
  const __nv_bfloat16* qbase = p.qkv + static_cast<size_t>(img) * p.grid * p.grid * ld + head * kSamHd;
  const __nv_bfloat16* kbase = qbase + p.dim_out;
  const __nv_bfloat16* vbase = kbase + p.dim_out;
  const int nblocks = (nk + kSamKB - 1) / kSamKB;

  auto prefetch = [&](int blk, int stage) {
    __nv_bfloat16* sK = sKV + stage * 2 * kSamKB * kSamRow;
    __nv_bfloat16* sV = sK + kSamKB * kSamRow;
    for (int it = tid; it < kSamKB * 9; it += kThreads) {
      const int r = it / 9, c = it - r * 9;
      const int ki = blk * kSamKB + r;
      if (ki < nk) {
        const int ky = ki / p.ws, kx = ki - ky * p.ws;
        const size_t t = static_cast<size_t>(wy * p.ws + ky) * p.grid + wx * p.ws + kx;
        cp_async16(sK + r * kSamRow + c * 8, kbase + t * ld + c * 8);
        cp_async16(sV + r * kSamRow + c * 8, vbase + t * ld + c * 8);
      } else {   // rows past the window: zeros (P is 0 there, but 0 * garbage could be NaN)
        *reinterpret_cast<uint4*>(sK + r * kSamRow + c * 8) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(sV + r * kSamRow + c * 8) = make_uint4(0, 0, 0, 0);
      }
    }
  };
  prefetch(0, 0);
  cp_async_commit();
  // zero pad columns 72..79 of every K/V row (never touched by cp.async) while block 0 is in flight
  for (int r = tid; r < 4 * kSamKB; r += kThreads) *reinterpret_cast<uint4*>(sKV + r * kSamRow + 72) = make_uint4(0, 0, 0, 0);

  // ---- stage Q (QB rows x 10 chunks of 8 bf16; chunk 9 is the zero pad 72..79)
  for (int it = tid; it < QB * 10; it += kThreads) {
    const int r = it / 10, c = it - r * 10;
    uint4 v = make_uint4(0, 0, 0, 0);
    const int qi = q0 + r;
    if (c < 9 && qi < nq) {
      const int qy = qi / wsq, qx = qi - qy * wsq;
      if (p.q_pool) {
        const size_t t00 = static_cast<size_t>(wy * p.ws + 2 * qy) * p.grid + wx * p.ws + 2 * qx;
        const uint4 a = *reinterpret_cast<const uint4*>(qbase + t00 * ld + c * 8);
        const uint4 b = *reinterpret_cast<const uint4*>(qbase + (t00 + 1) * ld + c * 8);
        const uint4 cc = *reinterpret_cast<const uint4*>(qbase + (t00 + p.grid) * ld + c * 8);
        const uint4 d = *reinterpret_cast<const uint4*>(qbase + (t00 + p.grid + 1) * ld + c * 8);
        v = bf16x8_max(bf16x8_max(a, b), bf16x8_max(cc, d));
      } else {
        const size_t t = static_cast<size_t>(wy * p.ws + qy) * p.grid + wx * p.ws + qx;
        v = *reinterpret_cast<const uint4*>(qbase + t * ld + c * 8);
      }
    }
    *reinterpret_cast<uint4*>(sQ + r * kSamRow + c * 8) = v;
  }
  __syncthreads();

  const bool active = q0 + warp * 16 < nq;   // warp uniform; rows >= nq of the last tile are zero padding
  uint32_t qf[5][4];
  if (active) {
#pragma unroll
    for (int ks = 0; ks < 5; ++ks)
      ldsm_x4(qf[ks], sQ + (warp * 16 + (lane & 15)) * kSamRow + ks * 16 + ((lane >> 4) << 3));
  }
  float o[9][4];
#pragma unroll
  for (int i = 0; i < 9; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

  for (int blk = 0; blk < nblocks; ++blk) {
    const int kb = blk * kSamKB;
    if (blk + 1 < nblocks) {
      prefetch(blk + 1, (blk + 1) & 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();   // block `blk` has landed for every thread
    const __nv_bfloat16* sK = sKV + (blk & 1) * 2 * kSamKB * kSamRow;
    const __nv_bfloat16* sV = sK + kSamKB * kSamRow;
    if (active) {
    const int nvalid = min(kSamKB, nk - kb);     // multiple of 16
    const int ntiles = nvalid >> 3;

    float s[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
      if (nt < ntiles) {
#pragma unroll
        for (int ks = 0; ks < 5; ++ks) {
          uint32_t bf[2];
          ldsm_x2(bf, sK + (nt * 8 + (lane & 7)) * kSamRow + ks * 16 + (((lane >> 3) & 1) << 3));
          mma_bf16_16816(s[nt], qf[ks], bf);
        }
      }
    }
    // online softmax (rows g = lane/4 and g+8)
    float bm0 = -INFINITY, bm1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
      if (nt < ntiles) {
        bm0 = fmaxf(bm0, fmaxf(s[nt][0], s[nt][1]));
        bm1 = fmaxf(bm1, fmaxf(s[nt][2], s[nt][3]));
      }
    bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1)); bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
    bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1)); bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
    const float nm0 = fmaxf(m0, bm0), nm1 = fmaxf(m1, bm1);
    const float c0 = fast_ex2((m0 - nm0) * p.scale_log2e), c1 = fast_ex2((m1 - nm1) * p.scale_log2e);
    m0 = nm0; m1 = nm1;
    const float ms0 = m0 * p.scale_log2e, ms1 = m1 * p.scale_log2e;
    float rs0 = 0.f, rs1 = 0.f;
    uint32_t pf[4][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      float e0 = 0.f, e1 = 0.f, e2 = 0.f, e3 = 0.f;
      if (nt < ntiles) {
        e0 = fast_ex2(fmaf(s[nt][0], p.scale_log2e, -ms0)); e1 = fast_ex2(fmaf(s[nt][1], p.scale_log2e, -ms0));
        e2 = fast_ex2(fmaf(s[nt][2], p.scale_log2e, -ms1)); e3 = fast_ex2(fmaf(s[nt][3], p.scale_log2e, -ms1));
      }
      rs0 += e0 + e1; rs1 += e2 + e3;
      pf[nt >> 1][(nt & 1) * 2] = pack_bf16(e0, e1);
      pf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(e2, e3);
    }
    l0 = l0 * c0 + rs0; l1 = l1 * c1 + rs1;
#pragma unroll
    for (int i = 0; i < 9; ++i) { o[i][0] *= c0; o[i][1] *= c0; o[i][2] *= c1; o[i][3] *= c1; }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      if (kk * 16 < nvalid) {
#pragma unroll
        for (int nt = 0; nt < 9; ++nt) {
          uint32_t bf[2];
          ldsm_x2_trans(bf, sV + (kk * 16 + (lane & 15)) * kSamRow + nt * 8);
          mma_bf16_16816(o[nt], pf[kk], bf);
        }
      }
    }
    }
    __syncthreads();   // everyone is done with this stage before the prefetch two blocks ahead overwrites it
  }
  if (!active) return;
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.f / l0, i1 = 1.f / l1;
  const int g = lane >> 2, t4 = lane & 3;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int qi = q0 + warp * 16 + g + 8 * h;
    if (qi >= nq) continue;
    const int qy = qi / wsq, qx = qi - qy * wsq;
    const size_t row = static_cast<size_t>(wy * wsq + qy) * grid_out + wx * wsq + qx;
    __nv_bfloat16* dst = p.out + (static_cast<size_t>(img) * grid_out * grid_out + row) * p.dim_out + head * kSamHd + 2 * t4;
    const float inv = h ? i1 : i0;
#pragma unroll
    for (int nt = 0; nt < 9; ++nt)
      *reinterpret_cast<uint32_t*>(dst + nt * 8) = pack_bf16(o[nt][2 * h] * inv, o[nt][2 * h + 1] * inv);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// SAM-2 mask decoder, token -> image cross attention (sam/transformer.py:192-197,124-130): 8 query tokens per prompt
// against n_keys image tokens, 8 heads of 16.  One CTA per prompt, one warp per head; K/V tiles of 64 keys x 8 heads
// (256-byte rows: fully coalesced) double-buffered with cp.async; S^T = Q K^T and O = P V on mma.sync.m16n8k16 (the 8
// queries fill half of the 16-row tile).  q [P*8,128] bf16; k,v [kv_batch * n_keys, 128] bf16 (kv_stride = 0: the same
// keys for every prompt, layer 0); out [P*8,128] bf16.
constexpr int kT2iRow = 136;   // 128 + 8 bf16: 272-byte rows, 8 consecutive rows hit distinct 16-B bank groups
constexpr int kT2iSmemBytes = 2 * 2 * 64 * kT2iRow * 2;

__global__ void __launch_bounds__(256) sam_t2i_attn_mma_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k,
                                                               const __nv_bfloat16* __restrict__ v, size_t kv_stride, int n_keys,
                                                               __nv_bfloat16* __restrict__ out) {
  griddep_launch();
  extern __shared__ __align__(16) uint8_t t2i_smem[];
  __nv_bfloat16* sKV = reinterpret_cast<__nv_bfloat16*>(t2i_smem);
  const int prompt = blockIdx.x, tid = threadIdx.x, head = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const __nv_bfloat16* kb = k + static_cast<size_t>(prompt) * kv_stride;
  const __nv_bfloat16* vb = v + static_cast<size_t>(prompt) * kv_stride;
  const int nblocks = n_keys / 64;

  auto prefetch = [&](int blk, int stage) {
    __nv_bfloat16* sK = sKV + stage * 2 * 64 * kT2iRow;
    __nv_bfloat16* sV = sK + 64 * kT2iRow;
    for (int it = tid; it < 64 * 16; it += 256) {
      const int r = it >> 4, c = it & 15;
      const size_t src = (static_cast<size_t>(blk) * 64 + r) * 128 + c * 8;
      cp_async16(sK + r * kT2iRow + c * 8, kb + src);
      cp_async16(sV + r * kT2iRow + c * 8, vb + src);
    }
  };
  prefetch(0, 0);
  cp_async_commit();
  // A fragment of Q (rows 0..7 = the 8 tokens, rows 8..15 zero), head_dim 16 = one k-step
  uint32_t qf[4];
  {
    const __nv_bfloat16* qp = q + (static_cast<size_t>(prompt) * 8 + g) * 128 + head * 16 + 2 * t4;
    qf[0] = *reinterpret_cast<const uint32_t*>(qp);
    qf[2] = *reinterpret_cast<const uint32_t*>(qp + 8);
    qf[1] = 0; qf[3] = 0;
  }
  const float sc = 0.25f * 1.4426950408889634f;   // 16^-0.5 * log2(e)
  float o[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
  float m0 = -INFINITY, l0 = 0.f;
  for (int blk = 0; blk < nblocks; ++blk) {
    if (blk + 1 < nblocks) {
      prefetch(blk + 1, (blk + 1) & 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const __nv_bfloat16* sK = sKV + (blk & 1) * 2 * 64 * kT2iRow;
    const __nv_bfloat16* sV = sK + 64 * kT2iRow;
    float s[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
      uint32_t bf[2];
      ldsm_x2(bf, sK + (nt * 8 + (lane & 7)) * kT2iRow + head * 16 + (((lane >> 3) & 1) << 3));
      mma_bf16_16816(s[nt], qf, bf);
    }
    float bm = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) bm = fmaxf(bm, fmaxf(s[nt][0], s[nt][1]));
    bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, 1)); bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, 2));
    const float nm = fmaxf(m0, bm);
    const float c0 = fast_ex2((m0 - nm) * sc);
    m0 = nm;
    const float ms = m0 * sc;
    float rs = 0.f;
    uint32_t pf[4][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float e0 = fast_ex2(fmaf(s[nt][0], sc, -ms)), e1 = fast_ex2(fmaf(s[nt][1], sc, -ms));
      rs += e0 + e1;
      pf[nt >> 1][(nt & 1) * 2] = pack_bf16(e0, e1);
      pf[nt >> 1][(nt & 1) * 2 + 1] = 0;            // rows 8..15 are padding
    }
    l0 = l0 * c0 + rs;
#pragma unroll
    for (int i = 0; i < 2; ++i) { o[i][0] *= c0; o[i][1] *= c0; }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        uint32_t bf[2];
        ldsm_x2_trans(bf, sV + (kk * 16 + (lane & 15)) * kT2iRow + head * 16 + nt * 8);
        mma_bf16_16816(o[nt], pf[kk], bf);
      }
    __syncthreads();
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  const float inv = 1.f / l0;
  __nv_bfloat16* dst = out + (static_cast<size_t>(prompt) * 8 + g) * 128 + head * 16 + 2 * t4;
  *reinterpret_cast<uint32_t*>(dst) = pack_bf16(o[0][0] * inv, o[0][1] * inv);
  *reinterpret_cast<uint32_t*>(dst + 8) = pack_bf16(o[1][0] * inv, o[1][1] * inv);
}

}  // namespace ovo
