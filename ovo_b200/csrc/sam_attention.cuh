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
#include "ptx.cuh"

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
};

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x2(uint32_t (&r)[2], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];\n" : "=r"(r[0]), "=r"(r[1]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x2_trans(uint32_t (&r)[2], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];\n" : "=r"(r[0]), "=r"(r[1]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__device__ __forceinline__ uint4 bf16x8_max(uint4 a, uint4 b) {
  uint4 r;
  const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
  const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b);
  __nv_bfloat162* pr = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) pr[i] = __hmax2(pa[i], pb[i]);
  return r;
}

__global__ void __launch_bounds__(128) hiera_attention_kernel(WinAttnParams p) {
  __shared__ __align__(16) __nv_bfloat16 sQ[kSamQB * kSamRow];
  __shared__ __align__(16) __nv_bfloat16 sK[kSamKB * kSamRow];
  __shared__ __align__(16) __nv_bfloat16 sV[kSamKB * kSamRow];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int head = blockIdx.z;
  const int wins = p.grid / p.ws;
  const int wy = blockIdx.y / wins, wx = blockIdx.y - wy * wins;
  const int nk = p.ws * p.ws;
  const int wsq = p.q_pool ? p.ws >> 1 : p.ws;        // window side on the query / output grid
  const int nq = wsq * wsq;
  const int grid_out = p.q_pool ? p.grid >> 1 : p.grid;
  const int q0 = blockIdx.x * kSamQB;
  const int ld = 3 * p.dim_out;
  const __nv_bfloat16* qbase = p.qkv + head * kSamHd;
  const __nv_bfloat16* kbase = qbase + p.dim_out;
  const __nv_bfloat16* vbase = kbase + p.dim_out;

  // ---- stage Q (64 rows x 10 chunks of 8 bf16; chunk 9 is the zero pad 72..79)
  for (int it = tid; it < kSamQB * 10; it += 128) {
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

  for (int kb = 0; kb < nk; kb += kSamKB) {
    __syncthreads();   // previous block fully consumed
    for (int it = tid; it < kSamKB * 10; it += 128) {
      const int r = it / 10, c = it - r * 10;
      uint4 kv = make_uint4(0, 0, 0, 0), vv = make_uint4(0, 0, 0, 0);
      const int ki = kb + r;
      if (c < 9 && ki < nk) {
        const int ky = ki / p.ws, kx = ki - ky * p.ws;
        const size_t t = static_cast<size_t>(wy * p.ws + ky) * p.grid + wx * p.ws + kx;
        kv = *reinterpret_cast<const uint4*>(kbase + t * ld + c * 8);
        vv = *reinterpret_cast<const uint4*>(vbase + t * ld + c * 8);
      }
      *reinterpret_cast<uint4*>(sK + r * kSamRow + c * 8) = kv;
      *reinterpret_cast<uint4*>(sV + r * kSamRow + c * 8) = vv;
    }
    __syncthreads();
    if (!active) continue;
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
    const float c0 = exp2f((m0 - nm0) * p.scale_log2e), c1 = exp2f((m1 - nm1) * p.scale_log2e);
    m0 = nm0; m1 = nm1;
    float rs0 = 0.f, rs1 = 0.f;
    uint32_t pf[4][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      float e0 = 0.f, e1 = 0.f, e2 = 0.f, e3 = 0.f;
      if (nt < ntiles) {
        e0 = exp2f((s[nt][0] - m0) * p.scale_log2e); e1 = exp2f((s[nt][1] - m0) * p.scale_log2e);
        e2 = exp2f((s[nt][2] - m1) * p.scale_log2e); e3 = exp2f((s[nt][3] - m1) * p.scale_log2e);
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
    __nv_bfloat16* dst = p.out + row * p.dim_out + head * kSamHd + 2 * t4;
    const float inv = h ? i1 : i0;
#pragma unroll
    for (int nt = 0; nt < 9; ++nt)
      *reinterpret_cast<uint32_t*>(dst + nt * 8) = pack_bf16(o[nt][2 * h] * inv, o[nt][2 * h + 1] * inv);
  }
}

}  // namespace ovo
