// SAM-2 mask proposal, image path behind the C ABI of include/ovo_b200.h.
// Reference: thirdParty/segment-anything-2/sam2/ — utils/transforms.py, modeling/backbones/{hieradet,image_encoder,utils}.py,
// modeling/sam2_base.py:467-479, sam2_image_predictor.py, modeling/sam/{prompt_encoder,transformer,mask_decoder}.py,
// automatic_mask_generator.py, utils/amg.py; and ovo/entities/mask_generator.py:102-120.
//
// Every linear layer / 1x1 conv / transposed conv runs on the tcgen05 GEMM of gemm.cuh; attention inside the Hiera
// windows runs on sam_attention.cuh; the decoder's tiny attentions (8 tokens, head_dim 16/32) and the mask
// post-processing are plain CUDA-core kernels (HBM / latency bound).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <vector>

#include "common.cuh"
#include "gemm.cuh"
#include "sam_attention.cuh"

namespace ovo {
namespace {

constexpr int kC = 256;        // d_model of neck / prompt encoder / mask decoder (sam2_base.py:207-243)
constexpr int kTok = 8;        // obj_score + iou + 4 mask tokens + point + padding point (mask_decoder.py:186-204)
constexpr int kInt = 128;      // internal dim of the cross attentions (attention_downsample_rate 2)

// ------------------------------------------------------------------------------------------- small kernels
// LayerNorm over rows of `width` (any multiple of 4 up to 2048), optional row-modulo additive table written as a second
// bf16 output: out16 = bf16(y), out16b = bf16(y + add[row % add_mod]).
__global__ void __launch_bounds__(256)
    sam_ln_kernel(const float* __restrict__ x, int rows, int width, const float* __restrict__ g, const float* __restrict__ b,
                  float eps, float* __restrict__ out32, __nv_bfloat16* __restrict__ out16, __nv_bfloat16* __restrict__ out16b,
                  const float* __restrict__ add, int add_mod) {
  griddep_launch();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * width);
  const int nvec = width >> 2;
  float4 v[16];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int idx = lane + 32 * i;
    if (idx < nvec) { v[i] = xr[idx]; sum += v[i].x + v[i].y + v[i].z + v[i].w; }
  }
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / width;
  float var = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int idx = lane + 32 * i;
    if (idx < nvec) {
      const float a = v[i].x - mean, bb = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      var += a * a + bb * bb + c * c + d * d;
    }
  }
  for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
  const float rstd = rsqrtf(var / width + eps);
  const float4* ar = add ? reinterpret_cast<const float4*>(add + static_cast<size_t>(row % add_mod) * width) : nullptr;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int idx = lane + 32 * i;
    if (idx < nvec) {
      const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + idx);
      const float4 bb = __ldg(reinterpret_cast<const float4*>(b) + idx);
      float4 o;
      o.x = (v[i].x - mean) * rstd * gg.x + bb.x; o.y = (v[i].y - mean) * rstd * gg.y + bb.y;
      o.z = (v[i].z - mean) * rstd * gg.z + bb.z; o.w = (v[i].w - mean) * rstd * gg.w + bb.w;
      if (out32) reinterpret_cast<float4*>(out32 + static_cast<size_t>(row) * width)[idx] = o;
      if (out16) reinterpret_cast<uint2*>(out16 + static_cast<size_t>(row) * width)[idx] = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
      if (out16b) {
        const float4 a = __ldg(ar + idx);
        reinterpret_cast<uint2*>(out16b + static_cast<size_t>(row) * width)[idx] = make_uint2(pack_bf16(o.x + a.x, o.y + a.y), pack_bf16(o.z + a.z, o.w + a.w));
      }
    }
  }
}

// Trunk LayerNorm (hieradet.py norm1/norm2, eps 1e-6) -> bf16 GEMM operand.  Rows are short (144..1152 floats), so one warp
// takes R rows per iteration and issues all of their loads before the first reduction (NV float4 per lane per row).
template <int NV, int R>
__global__ void __launch_bounds__(256)
    sam_ln_rows_kernel(const float* __restrict__ x, int rows, int width, const float* __restrict__ g, const float* __restrict__ b,
                       float eps, __nv_bfloat16* __restrict__ out16) {
  griddep_launch();
  const int lane = threadIdx.x & 31;
  const int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int nwarps = gridDim.x * (blockDim.x >> 5);
  const int nvec = width >> 2;
  const float inv_w = 1.f / static_cast<float>(width);
  for (int r0 = warp * R; r0 < rows; r0 += nwarps * R) {
    float4 v[R][NV];
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(min(r0 + j, rows - 1)) * width);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int idx = lane + 32 * i;
        v[j][i] = idx < nvec ? xr[idx] : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int j = 0; j < R; ++j) {
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) sum += (v[j][i].x + v[j][i].y) + (v[j][i].z + v[j][i].w);
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      const float mean = sum * inv_w;
      float var = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        if (lane + 32 * i < nvec) {
          const float a = v[j][i].x - mean, bb = v[j][i].y - mean, c = v[j][i].z - mean, d = v[j][i].w - mean;
          var += a * a + bb * bb + c * c + d * d;
        }
      }
      for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
      const float rstd = rsqrtf(var * inv_w + eps);
      const int r = r0 + j;
      if (r >= rows) continue;
      uint2* orow = reinterpret_cast<uint2*>(out16 + static_cast<size_t>(r) * width);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int idx = lane + 32 * i;
        if (idx < nvec) {
          const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + idx), bb = __ldg(reinterpret_cast<const float4*>(b) + idx);
          orow[idx] = make_uint2(pack_bf16((v[j][i].x - mean) * rstd * gg.x + bb.x, (v[j][i].y - mean) * rstd * gg.y + bb.y),
                                 pack_bf16((v[j][i].z - mean) * rstd * gg.z + bb.z, (v[j][i].w - mean) * rstd * gg.w + bb.w));
        }
      }
    }
  }
}

// The decoder's image-side LayerNorm (norm4, sam/transformer.py:209-210) over [prompts*4096, 256] is pure HBM streaming
// (3 GB per call at 256 prompts): one warp normalises 4 rows per iteration and issues all of their loads (8 float4 per
// lane + the positional table) before the first reduction, so enough bytes are in flight to cover the DRAM latency.
__global__ void __launch_bounds__(256)
    sam_ln256_kernel(const float* __restrict__ x, int rows, const float* __restrict__ g, const float* __restrict__ b, float eps,
                     float* __restrict__ out32, __nv_bfloat16* __restrict__ out16, __nv_bfloat16* __restrict__ out16b,
                     const float* __restrict__ add, int add_mod) {
  griddep_launch();
  const int lane = threadIdx.x & 31;
  const int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int nwarps = gridDim.x * (blockDim.x >> 5);
  const float4 g0 = __ldg(reinterpret_cast<const float4*>(g) + lane), g1 = __ldg(reinterpret_cast<const float4*>(g) + 32 + lane);
  const float4 b0 = __ldg(reinterpret_cast<const float4*>(b) + lane), b1 = __ldg(reinterpret_cast<const float4*>(b) + 32 + lane);
  for (int r0 = warp * 4; r0 < rows; r0 += nwarps * 4) {
    float4 v[4][2], a[4][2];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = min(r0 + j, rows - 1);
      const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(r) * 256);
      v[j][0] = xr[lane]; v[j][1] = xr[32 + lane];
      if (out16b) {
        const float4* ar = reinterpret_cast<const float4*>(add + static_cast<size_t>(r % add_mod) * 256);
        a[j][0] = __ldg(ar + lane); a[j][1] = __ldg(ar + 32 + lane);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float sum = v[j][0].x + v[j][0].y + v[j][0].z + v[j][0].w + v[j][1].x + v[j][1].y + v[j][1].z + v[j][1].w;
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      const float mean = sum * (1.f / 256.f);
      float4 d0 = make_float4(v[j][0].x - mean, v[j][0].y - mean, v[j][0].z - mean, v[j][0].w - mean);
      float4 d1 = make_float4(v[j][1].x - mean, v[j][1].y - mean, v[j][1].z - mean, v[j][1].w - mean);
      float var = d0.x * d0.x + d0.y * d0.y + d0.z * d0.z + d0.w * d0.w + d1.x * d1.x + d1.y * d1.y + d1.z * d1.z + d1.w * d1.w;
      for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
      const float rstd = rsqrtf(var * (1.f / 256.f) + eps);
      const int r = r0 + j;
      if (r >= rows) continue;
      const float4 y0 = make_float4(d0.x * rstd * g0.x + b0.x, d0.y * rstd * g0.y + b0.y, d0.z * rstd * g0.z + b0.z, d0.w * rstd * g0.w + b0.w);
      const float4 y1 = make_float4(d1.x * rstd * g1.x + b1.x, d1.y * rstd * g1.y + b1.y, d1.z * rstd * g1.z + b1.z, d1.w * rstd * g1.w + b1.w);
      const size_t base = static_cast<size_t>(r) * 64;   // in float4 / uint2 units
      if (out32) { reinterpret_cast<float4*>(out32)[base + lane] = y0; reinterpret_cast<float4*>(out32)[base + 32 + lane] = y1; }
      if (out16) {
        reinterpret_cast<uint2*>(out16)[base + lane] = make_uint2(pack_bf16(y0.x, y0.y), pack_bf16(y0.z, y0.w));
        reinterpret_cast<uint2*>(out16)[base + 32 + lane] = make_uint2(pack_bf16(y1.x, y1.y), pack_bf16(y1.z, y1.w));
      }
      if (out16b) {
        reinterpret_cast<uint2*>(out16b)[base + lane] = make_uint2(pack_bf16(y0.x + a[j][0].x, y0.y + a[j][0].y), pack_bf16(y0.z + a[j][0].z, y0.w + a[j][0].w));
        reinterpret_cast<uint2*>(out16b)[base + 32 + lane] = make_uint2(pack_bf16(y1.x + a[j][1].x, y1.y + a[j][1].y), pack_bf16(y1.z + a[j][1].z, y1.w + a[j][1].w));
      }
    }
  }
}

// out16[i] = bf16(a[i] + b[(i / width % b_mod) * width + i % width])   (b optional): f32 -> bf16 GEMM operands
__global__ void add_cast_kernel(const float* __restrict__ a, const float* __restrict__ b, int b_mod, int width, size_t n4,
                                __nv_bfloat16* __restrict__ out16, float* __restrict__ out32) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 v = reinterpret_cast<const float4*>(a)[i];
  if (b) {
    const size_t e = i * 4;
    const size_t row = e / width, col = e - row * width;
    const float4 w = *reinterpret_cast<const float4*>(b + (row % b_mod) * width + col);
    v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
  }
  if (out16) reinterpret_cast<uint2*>(out16)[i] = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
  if (out32) reinterpret_cast<float4*>(out32)[i] = v;
}

// ---- image transform (utils/transforms.py:15-40): ToTensor, Resize (bilinear, antialias), Normalize
__global__ void sam_resize_h_kernel(const uint8_t* __restrict__ rgb, int H, int W, const int* __restrict__ xmin,
                                    const int* __restrict__ xsize, const float* __restrict__ xw, int kx, int S,
                                    float* __restrict__ tmp) {
  const int ox = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (ox >= S) return;
  const int x0 = xmin[ox], n = xsize[ox];
  const float* w = xw + static_cast<size_t>(ox) * kx;
  const uint8_t* src = rgb + (static_cast<size_t>(y) * W + x0) * 3;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  for (int k = 0; k < n; ++k) {
    const float wk = w[k];
    a0 += wk * (static_cast<float>(src[3 * k]) / 255.f);
    a1 += wk * (static_cast<float>(src[3 * k + 1]) / 255.f);
    a2 += wk * (static_cast<float>(src[3 * k + 2]) / 255.f);
  }
  float* dst = tmp + static_cast<size_t>(y) * S + ox;
  dst[0] = a0; dst[static_cast<size_t>(H) * S] = a1; dst[static_cast<size_t>(2) * H * S] = a2;
}
__global__ void sam_resize_v_kernel(const float* __restrict__ tmp, int H, const int* __restrict__ ymin,
                                    const int* __restrict__ ysize, const float* __restrict__ yw, int ky, int S,
                                    float* __restrict__ px) {
  const int ox = blockIdx.x * blockDim.x + threadIdx.x, oy = blockIdx.y, c = blockIdx.z;
  if (ox >= S) return;
  const int y0 = ymin[oy], n = ysize[oy];
  const float* w = yw + static_cast<size_t>(oy) * ky;
  const float* src = tmp + (static_cast<size_t>(c) * H + y0) * S + ox;
  float a = 0.f;
  for (int k = 0; k < n; ++k) a += w[k] * src[static_cast<size_t>(k) * S];
  const float mean = c == 0 ? 0.485f : (c == 1 ? 0.456f : 0.406f);
  const float sd = c == 0 ? 0.229f : (c == 1 ? 0.224f : 0.225f);
  px[(static_cast<size_t>(c) * S + oy) * S + ox] = (a - mean) / sd;
}
// PatchEmbed conv 7x7 stride 4 pad 3 (backbones/utils.py:65-95) as im2col rows [g*g, kpad] bf16, column = c*49 + ky*7 + kx
__global__ void sam_im2col_kernel(const float* __restrict__ px_all, int S, int g, int kpad, __nv_bfloat16* __restrict__ out_all) {
  const int img = blockIdx.x / (g * g), t = blockIdx.x - img * g * g;
  const float* px = px_all + static_cast<size_t>(img) * 3 * S * S;
  __nv_bfloat16* out = out_all + static_cast<size_t>(img) * g * g * kpad;
  const int ty = t / g, tx = t - ty * g;
  for (int j = threadIdx.x; j < kpad; j += blockDim.x) {
    float v = 0.f;
    if (j < 147) {
      const int c = j / 49, r = j - c * 49, ky = r / 7, kx = r - ky * 7;
      const int y = 4 * ty - 3 + ky, x = 4 * tx - 3 + kx;
      if (y >= 0 && y < S && x >= 0 && x < S) v = px[(static_cast<size_t>(c) * S + y) * S + x];
    }
    out[static_cast<size_t>(t) * kpad + j] = __float2bfloat16_rn(v);
  }
}
// MaxPool2d(2,2) on a [g,g,C] f32 token grid -> [g/2,g/2,C]  (hieradet.py:25-37, the shortcut of a transition block)
__global__ void sam_maxpool_kernel(const float* __restrict__ in, int n_img, int g, int C, float* __restrict__ out) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int c4 = C >> 2, go = g >> 1;
  if (i >= static_cast<size_t>(n_img) * go * go * c4) return;
  const int c = i % c4;
  const size_t t_all = i / c4;
  const size_t img = t_all / (static_cast<size_t>(go) * go), t = t_all - img * go * go;
  const int y = t / go, x = t - static_cast<size_t>(y) * go;
  const float4* p = reinterpret_cast<const float4*>(in) + img * g * g * c4;
  const size_t r0 = (static_cast<size_t>(2 * y) * g + 2 * x) * c4 + c;
  const float4 a = p[r0], b = p[r0 + c4], d = p[r0 + static_cast<size_t>(g) * c4], e = p[r0 + static_cast<size_t>(g) * c4 + c4];
  reinterpret_cast<float4*>(out)[i] = make_float4(fmaxf(fmaxf(a.x, b.x), fmaxf(d.x, e.x)), fmaxf(fmaxf(a.y, b.y), fmaxf(d.y, e.y)),
                                                  fmaxf(fmaxf(a.z, b.z), fmaxf(d.z, e.z)), fmaxf(fmaxf(a.w, b.w), fmaxf(d.w, e.w)));
}
// FPN top-down (image_encoder.py:117-128, nearest x2) + decoder-side constants:
//   embed = lat2 + up(lat3)   (no_mem_embed is already in lat2's bias);  src = embed + no_mask_embed
//   src_bf = bf16(src), srcpe_bf = bf16(src + dense_pe)
__global__ void sam_embed_kernel(const float* __restrict__ lat2, const float* __restrict__ lat3, int g, const float* __restrict__ no_mask,
                                 const float* __restrict__ dense_pe, float* __restrict__ embed, float* __restrict__ src,
                                 __nv_bfloat16* __restrict__ src_bf, __nv_bfloat16* __restrict__ srcpe_bf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;   // over n_img * g*g * 256 (blockIdx.y = image)
  if (i >= g * g * kC) return;
  const size_t o = static_cast<size_t>(blockIdx.y) * g * g * kC + i;
  const int c = i % kC, t = i / kC, y = t / g, x = t - y * g;
  const float e = lat2[o] + lat3[(static_cast<size_t>(blockIdx.y) * (g >> 1) * (g >> 1) + (y >> 1) * (g >> 1) + (x >> 1)) * kC + c];
  embed[o] = e;
  const float s = e + no_mask[c];
  src[o] = s;
  src_bf[o] = __float2bfloat16_rn(s);
  srcpe_bf[o] = __float2bfloat16_rn(s + dense_pe[i]);
}
// feature map [2g,2g,C] -> [g*g, (ky*2+kx)*C + c]: the layout in which a k2 s2 transposed conv (a GEMM with N = 4*C_out)
// meets its high-resolution skip feature (mask_decoder.py:214-217)
__global__ void sam_subpixel_kernel(const float* __restrict__ in, int g, int C, float* __restrict__ out) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(g) * g * 4 * C) return;
  const int c = i % C, sub = (i / C) & 3;
  const size_t t = i / (4 * C);
  const int y = t / g, x = t - static_cast<size_t>(y) * g;
  out[i] = in[((static_cast<size_t>(2 * y + (sub >> 1)) * 2 * g) + 2 * x + (sub & 1)) * C + c];
}

// feat_s0 [4g,4g,32] in the order the SECOND transposed conv's GEMM rows come in: up1 pixels are stored as
// (y, x, sub1) over the g x g grid, so row = (y*g + x)*4 + sub1 is up1 pixel (2y+ky1, 2x+kx1) and column
// sub2*32 + c is final pixel (2Y+ky2, 2X+kx2).
__global__ void sam_subpixel2_kernel(const float* __restrict__ in, int g, float* __restrict__ out) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(g) * g * 16 * 32) return;
  const int c = i & 31, sub2 = (i >> 5) & 3, sub1 = (i >> 7) & 3;
  const size_t t = i >> 9;
  const int y = t / g, x = t - static_cast<size_t>(y) * g;
  const int Y = 2 * (2 * y + (sub1 >> 1)) + (sub2 >> 1), X = 2 * (2 * x + (sub1 & 1)) + (sub2 & 1);
  out[i] = in[(static_cast<size_t>(Y) * 4 * g + X) * 32 + c];
}

// ---- prompt encoder (prompt_encoder.py:81-104, position_encoding.py:129-158): tokens [P,8,256] =
// [out_tokens(6); pe(point + 0.5) + point_embeddings[1]; not_a_point_embed]
__global__ void sam_tokens_kernel(const float* __restrict__ pts, int P, float image_size, const float* __restrict__ gauss,
                                  const float* __restrict__ point_embed, const float* __restrict__ not_a_point,
                                  const float* __restrict__ out_tokens, float* __restrict__ tokens) {
  const int p = blockIdx.x, c = threadIdx.x;   // 256 threads
  float* t = tokens + static_cast<size_t>(p) * kTok * kC;
#pragma unroll
  for (int j = 0; j < 6; ++j) t[j * kC + c] = out_tokens[j * kC + c];
  const float x = (pts[2 * p] + 0.5f) / image_size, y = (pts[2 * p + 1] + 0.5f) / image_size;
  const int f = c & 127;
  float v = (2.f * x - 1.f) * gauss[f] + (2.f * y - 1.f) * gauss[128 + f];
  v = 2.f * 3.14159265358979323846f * v;
  t[6 * kC + c] = (c < 128 ? sinf(v) : cosf(v)) + point_embed[c];
  t[7 * kC + c] = not_a_point[c];
}

// ---- decoder attentions (sam/transformer.py:255-286); all operands bf16, f32 math
// token self-attention: 8 tokens, 8 heads of 32.  q,k,v [P*8, 256] -> out [P*8, 256].  One CTA (256 threads) per prompt.
__global__ void __launch_bounds__(256) sam_self_attn_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k,
                                                            const __nv_bfloat16* __restrict__ v, __nv_bfloat16* __restrict__ out) {
  __shared__ float sq[kTok][kC], sk[kTok][kC], sv[kTok][kC];
  __shared__ float sp[8][kTok][kTok];
  const size_t base = static_cast<size_t>(blockIdx.x) * kTok * kC;
  for (int i = threadIdx.x; i < kTok * kC; i += 256) {
    sq[0][i] = __bfloat162float(q[base + i]); sk[0][i] = __bfloat162float(k[base + i]); sv[0][i] = __bfloat162float(v[base + i]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 8 * kTok * kTok; i += 256) {
    const int h = i >> 6, a = (i >> 3) & 7, b = i & 7;
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < 32; ++d) s += sq[a][h * 32 + d] * sk[b][h * 32 + d];
    sp[h][a][b] = s * 0.17677669529663687f;   // 32^-0.5
  }
  __syncthreads();
  if (threadIdx.x < 64) {
    const int h = threadIdx.x >> 3, a = threadIdx.x & 7;
    float m = -INFINITY;
    for (int b = 0; b < kTok; ++b) m = fmaxf(m, sp[h][a][b]);
    float l = 0.f;
    for (int b = 0; b < kTok; ++b) { const float e = expf(sp[h][a][b] - m); sp[h][a][b] = e; l += e; }
    for (int b = 0; b < kTok; ++b) sp[h][a][b] /= l;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kTok * kC; i += 256) {
    const int a = i >> 8, c = i & 255, h = c >> 5;
    float acc = 0.f;
#pragma unroll
    for (int b = 0; b < kTok; ++b) acc += sp[h][a][b] * sv[b][c];
    out[base + i] = __float2bfloat16_rn(acc);
  }
}

// image -> token cross attention: every image token attends to the 8 prompt tokens, 8 heads of 16.
// q [q_batch * n_img, 128] (q_stride = 0: shared queries, layer 0); k,v [P*8,128]; out [P*n_img,128].
// One thread per (prompt, image token, head); a warp = one head x 32 consecutive tokens, so every k/v read from shared
// memory is a warp-wide broadcast (one wavefront) and a block (8 warps = 8 heads) covers 32 whole 256-byte token rows.
__global__ void __launch_bounds__(256) sam_i2t_attn_kernel(const __nv_bfloat16* __restrict__ q, size_t q_stride, const __nv_bfloat16* __restrict__ k,
                                                           const __nv_bfloat16* __restrict__ v, int n_img, __nv_bfloat16* __restrict__ out) {
  __shared__ __align__(16) float sk[8][kTok][16], sv[8][kTok][16];   // [head][token][dim]
  const int p = blockIdx.y;
  for (int i = threadIdx.x; i < kTok * kInt; i += 256) {
    const int a = i >> 7, h = (i >> 4) & 7, d = i & 15;
    sk[h][a][d] = __bfloat162float(k[static_cast<size_t>(p) * kTok * kInt + i]) * 0.25f;   // 16^-0.5 folded into k
    sv[h][a][d] = __bfloat162float(v[static_cast<size_t>(p) * kTok * kInt + i]);
  }
  __syncthreads();
  const int h = threadIdx.x >> 5;
  const int t = blockIdx.x * 32 + (threadIdx.x & 31);
  if (t >= n_img) return;
  const uint4* qp = reinterpret_cast<const uint4*>(q + static_cast<size_t>(p) * q_stride + static_cast<size_t>(t) * kInt + h * 16);
  uint4 raw[2] = {qp[0], qp[1]};
  const __nv_bfloat16* qb = reinterpret_cast<const __nv_bfloat16*>(raw);
  float qf[16];
#pragma unroll
  for (int d = 0; d < 16; ++d) qf[d] = __bfloat162float(qb[d]);
  float s[kTok], m = -INFINITY;
#pragma unroll
  for (int a = 0; a < kTok; ++a) {
    float x = 0.f;
#pragma unroll
    for (int d4 = 0; d4 < 4; ++d4) {
      const float4 kk = *reinterpret_cast<const float4*>(&sk[h][a][4 * d4]);
      x += qf[4 * d4] * kk.x + qf[4 * d4 + 1] * kk.y + qf[4 * d4 + 2] * kk.z + qf[4 * d4 + 3] * kk.w;
    }
    s[a] = x; m = fmaxf(m, x);
  }
  float l = 0.f;
#pragma unroll
  for (int a = 0; a < kTok; ++a) { s[a] = __expf(s[a] - m); l += s[a]; }
  const float inv = 1.f / l;
  float o[16];
#pragma unroll
  for (int d = 0; d < 16; ++d) o[d] = 0.f;
#pragma unroll
  for (int a = 0; a < kTok; ++a)
#pragma unroll
    for (int d4 = 0; d4 < 4; ++d4) {
      const float4 vv = *reinterpret_cast<const float4*>(&sv[h][a][4 * d4]);
      o[4 * d4] += s[a] * vv.x; o[4 * d4 + 1] += s[a] * vv.y; o[4 * d4 + 2] += s[a] * vv.z; o[4 * d4 + 3] += s[a] * vv.w;
    }
  uint4 w[2];
  uint32_t* wp = reinterpret_cast<uint32_t*>(w);
#pragma unroll
  for (int d = 0; d < 8; ++d) wp[d] = pack_bf16(o[2 * d] * inv, o[2 * d + 1] * inv);
  uint4* op = reinterpret_cast<uint4*>(out + (static_cast<size_t>(p) * n_img + t) * kInt + h * 16);
  op[0] = w[0]; op[1] = w[1];
}

// gathers rows (token index `tok` of every prompt) of hs [P,8,256] f32 as bf16 [P,256]
__global__ void sam_pick_token_kernel(const float* __restrict__ hs, int P, int tok, __nv_bfloat16* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * kC) return;
  out[i] = __float2bfloat16_rn(hs[(static_cast<size_t>(i / kC) * kTok + tok) * kC + (i % kC)]);
}
// iou = sigmoid(head)[:, 1:4]   (mask_decoder.py:229, sam2_utils.py:134-135, multimask slice :141-143)
__global__ void sam_iou_kernel(const float* __restrict__ head /*[P,4]*/, int P, float* __restrict__ iou /*[P,3]*/) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * 3) return;
  const float x = head[(i / 3) * 4 + 1 + i % 3];
  iou[i] = 1.f / (1.f + expf(-x));
}

// ------------------------------------------------------------------------------------------- AMG post-processing
// F.interpolate(bilinear, align_corners=False) sample of a [h,w] logit map at output pixel (y,x) of [H,W]
// (ATen upsample_bilinear2d: src = scale*(dst+0.5)-0.5 clamped at 0, lambda in f32)
struct Bilin { int y0, y1, x0, x1; float ly, lx; };
__device__ __forceinline__ void bilin_axis(int d, float scale, int n, int& i0, int& i1, float& l) {
  float s = __fadd_rn(__fmul_rn(scale, __fadd_rn(static_cast<float>(d), 0.5f)), -0.5f);
  if (s < 0.f) s = 0.f;
  i0 = static_cast<int>(s);
  if (i0 > n - 1) i0 = n - 1;
  i1 = i0 + (i0 < n - 1 ? 1 : 0);
  l = __fadd_rn(s, -static_cast<float>(i0));
}
__device__ __forceinline__ float bilin_sample(const float* __restrict__ m, int w, int y0, int y1, int x0, int x1, float ly, float lx) {
  const float hy = __fadd_rn(1.f, -ly), hx = __fadd_rn(1.f, -lx);
  // ATen: w0y * (w0x * a + w1x * b) + w1y * (w0x * c + w1x * d)
  const float top = __fadd_rn(__fmul_rn(hx, m[y0 * w + x0]), __fmul_rn(lx, m[y0 * w + x1]));
  const float bot = __fadd_rn(__fmul_rn(hx, m[y1 * w + x0]), __fmul_rn(lx, m[y1 * w + x1]));
  return __fadd_rn(__fmul_rn(hy, top), __fmul_rn(ly, bot));
}

struct MaskStat { int hi, lo, area, x0, y0, x1, y1, pad; };   // counts of logit > +off, > -off, > 0; bbox of > 0

// candidates = masks with iou > pred_iou_thresh (cand[j] = flattened index).  grid (row bands, n_cand); the per-column
// interpolation table (x0, x1, lambda) is staged once per block in shared memory, the per-row one lives in registers.
constexpr int kAmgMaxW = 2048;
__global__ void __launch_bounds__(256) amg_stats_kernel(const float* __restrict__ low, const int* __restrict__ cand, int h, int w, int H, int W,
                                                        float off, MaskStat* __restrict__ stats) {
  __shared__ short s_x0[kAmgMaxW], s_x1[kAmgMaxW];
  __shared__ float s_lx[kAmgMaxW];
  const int j = blockIdx.y;
  const float* m = low + static_cast<size_t>(cand[j]) * h * w;
  const float sy = static_cast<float>(h) / H, sx = static_cast<float>(w) / W;
  for (int x = threadIdx.x; x < W; x += 256) {
    int xa, xb; float lx;
    bilin_axis(x, sx, w, xa, xb, lx);
    s_x0[x] = static_cast<short>(xa); s_x1[x] = static_cast<short>(xb); s_lx[x] = lx;
  }
  __syncthreads();
  const int rows_per = (H + gridDim.x - 1) / gridDim.x;
  const int r_lo = blockIdx.x * rows_per, r_hi = min(H, r_lo + rows_per);
  int hi = 0, lo = 0, area = 0, x0 = W, y0 = H, x1 = -1, y1 = -1;
  for (int y = r_lo; y < r_hi; ++y) {
    int ya, yb; float ly;
    bilin_axis(y, sy, h, ya, yb, ly);
    const float hy = __fadd_rn(1.f, -ly);
    const float* ra = m + ya * w;
    const float* rb = m + yb * w;
    for (int x = threadIdx.x; x < W; x += 256) {
      const int xa = s_x0[x], xb = s_x1[x];
      const float lx = s_lx[x], hx = __fadd_rn(1.f, -lx);
      const float top = __fadd_rn(__fmul_rn(hx, ra[xa]), __fmul_rn(lx, ra[xb]));
      const float bot = __fadd_rn(__fmul_rn(hx, rb[xa]), __fmul_rn(lx, rb[xb]));
      const float v = __fadd_rn(__fmul_rn(hy, top), __fmul_rn(ly, bot));
      hi += v > off; lo += v > -off;
      if (v > 0.f) { ++area; x0 = min(x0, x); x1 = max(x1, x); y0 = min(y0, y); y1 = max(y1, y); }
    }
  }
  for (int s = 16; s > 0; s >>= 1) {
    hi += __shfl_xor_sync(0xffffffffu, hi, s); lo += __shfl_xor_sync(0xffffffffu, lo, s); area += __shfl_xor_sync(0xffffffffu, area, s);
    x0 = min(x0, __shfl_xor_sync(0xffffffffu, x0, s)); y0 = min(y0, __shfl_xor_sync(0xffffffffu, y0, s));
    x1 = max(x1, __shfl_xor_sync(0xffffffffu, x1, s)); y1 = max(y1, __shfl_xor_sync(0xffffffffu, y1, s));
  }
  if ((threadIdx.x & 31) == 0) {
    MaskStat* st = stats + j;
    atomicAdd(&st->hi, hi); atomicAdd(&st->lo, lo); atomicAdd(&st->area, area);
    atomicMin(&st->x0, x0); atomicMin(&st->y0, y0); atomicMax(&st->x1, x1); atomicMax(&st->y1, y1);
  }
}
__global__ void amg_stats_init_kernel(MaskStat* st, int n, int H, int W) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) st[i] = MaskStat{0, 0, 0, W, H, -1, -1, 0};
}
// binarise the kept masks in their final order: out[k] = upsample(low[sel[k]]) > 0
__global__ void __launch_bounds__(256) amg_write_masks_kernel(const float* __restrict__ low, const int* __restrict__ sel, int h, int w, int H, int W,
                                                              uint8_t* __restrict__ out) {
  const int k = blockIdx.y;
  const float* m = low + static_cast<size_t>(sel[k]) * h * w;
  const float sy = static_cast<float>(h) / H, sx = static_cast<float>(w) / W;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < H * W; i += gridDim.x * 256) {
    const int y = i / W, x = i - y * W;
    int ya, yb, xa, xb; float ly, lx;
    bilin_axis(y, sy, h, ya, yb, ly);
    bilin_axis(x, sx, w, xa, xb, lx);
    out[static_cast<size_t>(k) * H * W + i] = bilin_sample(m, w, ya, yb, xa, xb, ly, lx) > 0.f ? 1 : 0;
  }
}

// counters: [0] n_cand, [1] K (final)
// predicted-IoU filter (automatic_mask_generator.py:331-334): ordered compaction of the flattened [P*3] list
__global__ void amg_candidates_kernel(const float* __restrict__ iou, int n, float thr, int* __restrict__ cand, int* __restrict__ counters) {
  __shared__ int s_off;
  if (threadIdx.x == 0) s_off = 0;
  __syncthreads();
  for (int base = 0; base < n; base += blockDim.x) {   // blockDim.x == 1024, one block
    const int i = base + threadIdx.x;
    const bool keep = i < n && iou[i] > thr;
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    __shared__ int s_warp[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    int before = 0;
    for (int w = 0; w < warp; ++w) before += s_warp[w];
    int total = 0;
    for (int w = 0; w < 32; ++w) total += s_warp[w];
    if (keep) cand[s_off + before + __popc(bal & ((1u << lane) - 1))] = i;
    __syncthreads();
    if (threadIdx.x == 0) s_off += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) counters[0] = s_off;
}

// stability filter (:337-342, utils/amg.py:158-178), boxes (:305-348), box NMS (torchvision nms: stable descending sort by
// predicted IoU, greedy, IoU = inter / (a_i + a_j - inter) > thr suppresses).  One block of 1024 threads, n_cand <= 1024.
__global__ void __launch_bounds__(1024) amg_decide_kernel(const MaskStat* __restrict__ stats, const int* __restrict__ cand, const float* __restrict__ iou,
                                                          float stab_thr, float nms_thr, int* __restrict__ counters, int* __restrict__ sel,
                                                          float* __restrict__ iou_out, float* __restrict__ stab_out, int32_t* __restrict__ boxes_out,
                                                          int32_t* __restrict__ src_out, int max_out) {
  __shared__ float s_score[1024], s_stab[1024];
  __shared__ int s_box[1024][4];
  __shared__ int s_idx[1024];      // candidate slot in sorted position
  __shared__ unsigned char s_dead[1024];
  __shared__ int s_n;
  const int n_cand = counters[0];
  const int t = threadIdx.x;
  // 1. stability filter, ordered compaction into shared arrays (slot order = flattened index order)
  if (t == 0) {
    int n = 0;
    for (int j = 0; j < n_cand && j < 1024; ++j) {
      const MaskStat st = stats[j];
      const float stab = __fdiv_rn(static_cast<float>(st.hi), static_cast<float>(st.lo));   // 0/0 -> NaN -> rejected
      if (stab >= stab_thr) {
        s_score[n] = iou[cand[j]]; s_stab[n] = stab; s_idx[n] = cand[j];
        const bool empty = st.area == 0;
        s_box[n][0] = empty ? 0 : st.x0; s_box[n][1] = empty ? 0 : st.y0; s_box[n][2] = empty ? 0 : st.x1; s_box[n][3] = empty ? 0 : st.y1;
        ++n;
      }
    }
    s_n = n;
  }
  __syncthreads();
  const int n = s_n;
  // 2. stable descending rank
  int rank = 0;
  if (t < n) {
    const float my = s_score[t];
    for (int j = 0; j < n; ++j) rank += (s_score[j] > my) || (s_score[j] == my && j < t);
  }
  __shared__ int s_order[1024];
  if (t < n) s_order[rank] = t;
  if (t < 1024) s_dead[t] = 0;
  __syncthreads();
  // 3. greedy suppression in sorted order
  for (int a = 0; a < n; ++a) {
    if (!s_dead[a]) {
      const int ia = s_order[a];
      const float ax0 = s_box[ia][0], ay0 = s_box[ia][1], ax1 = s_box[ia][2], ay1 = s_box[ia][3];
      const float aa = __fmul_rn(ax1 - ax0, ay1 - ay0);
      const int b = a + 1 + t;
      if (b < n && !s_dead[b]) {
        const int ib = s_order[b];
        const float bx0 = s_box[ib][0], by0 = s_box[ib][1], bx1 = s_box[ib][2], by1 = s_box[ib][3];
        const float ab = __fmul_rn(bx1 - bx0, by1 - by0);
        const float w = fmaxf(0.f, fminf(ax1, bx1) - fmaxf(ax0, bx0)), h = fmaxf(0.f, fminf(ay1, by1) - fmaxf(ay0, by0));
        const float inter = __fmul_rn(w, h);
        if (__fdiv_rn(inter, __fadd_rn(__fadd_rn(aa, ab), -inter)) > nms_thr) s_dead[b] = 1;
      }
    }
    __syncthreads();
  }
  // 4. survivors in sorted order
  if (t == 0) {
    int k = 0;
    for (int a = 0; a < n; ++a) {
      if (s_dead[a]) continue;
      if (k < max_out) {
        const int i = s_order[a];
        sel[k] = s_idx[i]; iou_out[k] = s_score[i]; stab_out[k] = s_stab[i]; src_out[k] = s_idx[i];
        for (int c = 0; c < 4; ++c) boxes_out[4 * k + c] = s_box[i][c];
      }
      ++k;
    }
    counters[1] = k;
  }
}
__global__ void mul_kernel(const float* a, const float* b, int n, float* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] * b[i];
}
// compaction of the masks kept by OVO's NMS: dst[k] = src[idx[k]]
__global__ void gather_masks_kernel(const uint8_t* __restrict__ src, const int* __restrict__ idx, size_t npix16, uint8_t* __restrict__ dst) {
  const uint4* s4 = reinterpret_cast<const uint4*>(src) + static_cast<size_t>(idx[blockIdx.y]) * npix16;
  uint4* d4 = reinterpret_cast<uint4*>(dst) + static_cast<size_t>(blockIdx.y) * npix16;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < npix16; i += static_cast<size_t>(gridDim.x) * blockDim.x) d4[i] = s4[i];
}
// AMG point grid (utils/amg.py:181-189) in model-frame pixels the way the reference's f32 arithmetic produces them
// (automatic_mask_generator.py:264-265,308-310; utils/transforms.py:59-65): f32(grid*W) / W * S
__global__ void amg_points_kernel(int n, int H, int W, float S, float* __restrict__ pts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * n) return;
  const int iy = i / n, ix = i - iy * n;
  // np.linspace(offset, 1 - offset, n) in f64: start + k * step, step = (stop - start) / (n - 1)
  const double off = 1.0 / (2.0 * n), step = n > 1 ? ((1.0 - off) - off) / (n - 1) : 0.0;
  const double gx = (n > 1 && ix == n - 1) ? 1.0 - off : off + ix * step, gy = (n > 1 && iy == n - 1) ? 1.0 - off : off + iy * step;
  const float px = static_cast<float>(gx * W), py = static_cast<float>(gy * H);
  pts[2 * i] = __fmul_rn(__fdiv_rn(px, static_cast<float>(W)), S);
  pts[2 * i + 1] = __fmul_rn(__fdiv_rn(py, static_cast<float>(H)), S);
}

template <typename T>
int dalloc(T** p, size_t n) {
  if (cudaMalloc(reinterpret_cast<void**>(p), std::max<size_t>(n, 1) * sizeof(T)) != cudaSuccess) {
    cudaGetLastError();
    return set_error(OVO_E_NOMEM, "sam workspace allocation of %zu bytes failed", n * sizeof(T));
  }
  return OVO_OK;
}

// ATen _upsample_bilinear2d_aa weights for one axis (aten/src/ATen/native/cpu/UpSampleKernel.cpp)
void aa_axis(int n_in, int n_out, std::vector<int>& mins, std::vector<int>& sizes, std::vector<float>& ws, int* kmax) {
  const float scale = static_cast<float>(n_in) / n_out;
  const float support = scale >= 1.f ? scale : 1.f;
  const float inv = scale >= 1.f ? 1.f / scale : 1.f;
  const int k = static_cast<int>(std::ceil(support)) * 2 + 1;
  *kmax = k;
  mins.assign(n_out, 0); sizes.assign(n_out, 0); ws.assign(static_cast<size_t>(n_out) * k, 0.f);
  for (int i = 0; i < n_out; ++i) {
    const float center = scale * (i + 0.5f);
    const int xmin = std::max(0, static_cast<int>(center - support + 0.5f));
    const int xmax = std::min(n_in, static_cast<int>(center + support + 0.5f));
    float total = 0.f;
    for (int j = 0; j < xmax - xmin; ++j) {
      const float w = std::max(0.f, 1.f - std::fabs((j + xmin - center + 0.5f) * inv));
      ws[static_cast<size_t>(i) * k + j] = w; total += w;
    }
    for (int j = 0; j < xmax - xmin; ++j) ws[static_cast<size_t>(i) * k + j] /= total;
    mins[i] = xmin; sizes[i] = xmax - xmin;
  }
}

}  // namespace
}  // namespace ovo

using namespace ovo;

struct ovo_sam {
  ovo_sam_cfg cfg{};
  ovo_sam_weights w{};
  std::vector<ovo_hiera_block> blocks;
  std::vector<ovo_sam_dec_layer> layers;
  int S = 0, g = 0;            // image size, embedding grid (S/16)
  int max_batch = 1;           // images the trunk can take in one pass (activation buffers are sized for it)
  int n_set = 0, cur = 0;      // images of the last set_image(s) call; the one the decoder works on
  int max_h = 0, max_w = 0, max_p = 0;
  // transform
  int tab_h = -1, tab_w = -1, kx = 0, ky = 0;
  int *xmin = nullptr, *xsize = nullptr, *ymin = nullptr, *ysize = nullptr;
  float *xw = nullptr, *yw = nullptr, *tmp = nullptr, *pixels = nullptr;
  __nv_bfloat16* patches = nullptr;
  // trunk
  float *xa = nullptr, *xb = nullptr, *tshort = nullptr;
  __nv_bfloat16 *xn = nullptr, *qkv = nullptr, *att = nullptr, *hid = nullptr;
  __nv_bfloat16* stage_bf[4] = {nullptr, nullptr, nullptr, nullptr};
  int stage_grid[4] = {0, 0, 0, 0}, stage_dim[4] = {0, 0, 0, 0};
  // neck / image features
  float *lat3 = nullptr, *lat2 = nullptr, *embed = nullptr, *feat_s0 = nullptr, *feat_s1 = nullptr, *s0_sub = nullptr, *s1_sub = nullptr;
  float* src = nullptr;
  __nv_bfloat16 *src_bf = nullptr, *srcpe_bf = nullptr, *k0 = nullptr, *v0 = nullptr, *qi0 = nullptr;
  // decoder (sized for max_p prompts)
  float *tokens = nullptr, *queries = nullptr, *tq_tmp = nullptr, *keys = nullptr, *keys_pre = nullptr, *dc1 = nullptr, *hyper = nullptr, *iou_head = nullptr;
  __nv_bfloat16 *tq_bf = nullptr, *tqpe_bf = nullptr, *t_q = nullptr, *t_k = nullptr, *t_v = nullptr, *t_o = nullptr, *t_mlp = nullptr;
  __nv_bfloat16 *keys_bf = nullptr, *keyspe_bf = nullptr, *big_q = nullptr, *big_k = nullptr, *big_v = nullptr, *up1 = nullptr, *up2 = nullptr;
  __nv_bfloat16 *tok_a = nullptr, *tok_b = nullptr;
  // post-processing
  MaskStat* stats = nullptr;
  int *cand = nullptr, *sel = nullptr;
  float* low_all = nullptr; float* iou_all = nullptr; float* points = nullptr;
  const float* low_override = nullptr; const float* iou_override = nullptr; int override_p = 0;   // measurement aid, see ovo_sam_override_logits
  uint8_t* masks_tmp = nullptr; float* score_tmp = nullptr; uint8_t* keep_tmp = nullptr; float* stab_tmp = nullptr; float* iou_tmp = nullptr;
  int32_t* box_tmp = nullptr; int32_t* src_tmp = nullptr; int32_t* order_tmp = nullptr;
  int* counters = nullptr;
  size_t masks_tmp_bytes = 0; uint8_t* masks_tmp2 = nullptr;
  // CUDA graphs of the static-shape launch sequences (trunk + neck: ~350 launches of 5-20 us each; decoder on the handle's
  // own buffers: ~75 launches): the first call per key runs eagerly, the second captures, later calls replay
  struct GraphEntry { cudaGraphExec_t exec = nullptr; long long launches = 0; int warm = 0; };
  std::map<long long, GraphEntry> graphs;
  cudaStream_t cap_stream = nullptr;
  bool use_graphs = true;
  std::vector<void*> owned;
};

namespace {

template <typename T>
int salloc(ovo_sam* s, T** p, size_t n) {
  OVO_TRY(dalloc(p, n));
  s->owned.push_back(*p);
  return OVO_OK;
}

int gemm(int epi, const __nv_bfloat16* A, int lda, const void* W, int ldw, int M, int N, int K, const float* bias, void* out, int ldo,
         const float* resid, int ldr, int resid_mod, cudaStream_t st) {
  EpiParams ep;
  ep.out = out; ep.ldo = ldo; ep.bias = bias; ep.resid = resid; ep.ldr = ldr; ep.resid_mod = resid_mod;
  return launch_gemm(epi, A, lda, static_cast<const __nv_bfloat16*>(W), ldw, M, N, K, ep, st);
}

int ln(const float* x, int rows, int width, const float* g, const float* b, float eps, float* o32, __nv_bfloat16* o16,
       __nv_bfloat16* o16b, const float* add, int add_mod, cudaStream_t st) {
  OVO_REQUIRE(width % 4 == 0 && width <= 2048, "sam layernorm: unsupported width %d", width);
  ProfScope prof(st, PROF_LN, 0.0, static_cast<double>(rows) * width * 8.0);
  if (width == 256 && rows >= 65536) {
    sam_ln256_kernel<<<num_sms() * 8, 256, 0, st>>>(x, rows, g, b, eps, o32, o16, o16b, add, add_mod);
    OVO_CHECK_LAUNCH();
    return OVO_OK;
  }
  if (o16 != nullptr && o32 == nullptr && o16b == nullptr && rows >= 1024) {   // the trunk's LayerNorms
    const int nv = ceil_div(width, 128);
    const int blocks = std::min(num_sms() * 8, ceil_div(rows, 8));
    bool done = true;
    if (nv <= 2) sam_ln_rows_kernel<2, 4><<<std::min(blocks, ceil_div(rows, 32)), 256, 0, st>>>(x, rows, width, g, b, eps, o16);
    else if (nv <= 3) sam_ln_rows_kernel<3, 2><<<std::min(blocks, ceil_div(rows, 16)), 256, 0, st>>>(x, rows, width, g, b, eps, o16);
    else if (nv <= 5) sam_ln_rows_kernel<5, 2><<<std::min(blocks, ceil_div(rows, 16)), 256, 0, st>>>(x, rows, width, g, b, eps, o16);
    else if (nv <= 9) sam_ln_rows_kernel<9, 1><<<blocks, 256, 0, st>>>(x, rows, width, g, b, eps, o16);
    else done = false;
    if (done) {
      OVO_CHECK_LAUNCH();
      return OVO_OK;
    }
  }
  sam_ln_kernel<<<ceil_div(rows, 8), 256, 0, st>>>(x, rows, width, g, b, eps, o32, o16, o16b, add, add_mod);
  OVO_CHECK_LAUNCH();
  return OVO_OK;
}

int add_cast(const float* a, const float* b, int b_mod, int width, size_t n, __nv_bfloat16* o16, float* o32, cudaStream_t st) {
  ProfScope prof(st, PROF_OTHER, 0.0, static_cast<double>(n) * 6.0);
  add_cast_kernel<<<ceil_div(static_cast<long long>(n / 4), 256), 256, 0, st>>>(a, b, b_mod, width, n / 4, o16, o32);
  OVO_CHECK_LAUNCH();
  return OVO_OK;
}

int g_hiera_qb128_min = 128;   // queries per window from which the 128-query CTA is used (OVO_B200_HIERA_QB128_MIN: tuning)

template <typename F>
int graphed(ovo_sam* s, long long key, cudaStream_t st, F&& fn) {
  ovo_sam::GraphEntry& ge = s->graphs[key];
  if (!s->use_graphs || ge.warm == 0 || profiling()) {
    OVO_TRY(fn(st));
    if (!profiling()) ge.warm = 1;
    return OVO_OK;
  }
  if (ge.exec == nullptr) {
    const long long before = ovo_launch_count(0);
    cudaGraph_t graph = nullptr;
    if (!s->cap_stream) OVO_CUDA(cudaStreamCreateWithFlags(&s->cap_stream, cudaStreamNonBlocking));
    OVO_CUDA(cudaStreamBeginCapture(s->cap_stream, cudaStreamCaptureModeThreadLocal));
    const int r = fn(s->cap_stream);
    const cudaError_t ce = cudaStreamEndCapture(s->cap_stream, &graph);
    if (r != OVO_OK) { if (graph) cudaGraphDestroy(graph); return r; }
    if (ce != cudaSuccess) return set_error(OVO_E_CUDA, "sam graph capture failed: %s", cudaGetErrorString(ce));
    const cudaError_t ie = cudaGraphInstantiate(&ge.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) { ge.exec = nullptr; return set_error(OVO_E_CUDA, "sam graph instantiate failed: %s", cudaGetErrorString(ie)); }
    ge.launches = ovo_launch_count(0) - before;
    count_launch(-static_cast<int>(ge.launches));   // captured, not executed yet
  }
  OVO_CUDA(cudaGraphLaunch(ge.exec, st));
  count_launch(static_cast<int>(ge.launches));
  return OVO_OK;
}

// Hiera trunk from the im2col'd patches (hieradet.py:274-291) + neck + decoder-side per-image constants
int run_trunk(ovo_sam* s, int B, int n_blocks, float* block_out, cudaStream_t st) {
  const ovo_sam_cfg& c = s->cfg;
  int grid = s->S / 4;
  const int T0 = grid * grid;
  float* x = s->xa;
  float* xo = s->xb;
  // patch embed + bias + pos embed
  OVO_TRY(gemm(EPI_F32_RESID, s->patches, s->w.patch_kpad, s->w.patch_w, s->w.patch_kpad, B * T0, c.embed_dim, s->w.patch_kpad, s->w.patch_b,
               x, c.embed_dim, s->w.pos, c.embed_dim, T0, st));
  const int nb = n_blocks < 0 ? c.n_blocks : std::min(n_blocks, c.n_blocks);
  int stage = 0, last_dim = c.embed_dim;
  for (int i = 0; i < nb; ++i) {
    const ovo_hiera_block& b = s->blocks[i];
    OVO_REQUIRE(b.grid_in == grid, "sam block %d: grid mismatch", i);
    const int T = B * grid * grid;               // token rows of all images
    const int ws = b.window > 0 ? b.window : grid;
    OVO_REQUIRE(grid % ws == 0 && (ws * ws) % 16 == 0 && (!b.q_pool || ws % 2 == 0), "sam block %d: unsupported window %d on grid %d", i, ws, grid);
    OVO_REQUIRE(b.dim_out == b.heads * kSamHd, "sam block %d: head_dim must be 72", i);
    OVO_TRY(ln(x, T, b.dim, b.norm1_w, b.norm1_b, c.trunk_ln_eps, nullptr, s->xn, nullptr, nullptr, 1, st));
    const float* resid = x;
    float* dst = x;
    int grid_out = grid;
    if (b.short_w != nullptr) {   // transition: shortcut = maxpool(proj(norm1(x)))  (hieradet.py:139-141)
      OVO_TRY(gemm(EPI_F32, s->xn, b.dim, b.short_w, b.dim, T, b.dim_out, b.dim, b.short_b, s->tshort, b.dim_out, nullptr, 0, 0, st));
      if (b.q_pool) {
        grid_out = grid / 2;
        sam_maxpool_kernel<<<ceil_div(static_cast<long long>(B) * grid_out * grid_out * (b.dim_out / 4), 256), 256, 0, st>>>(s->tshort, B, grid, b.dim_out, xo);
        OVO_CHECK_LAUNCH();
      } else {
        OVO_CUDA(cudaMemcpyAsync(xo, s->tshort, sizeof(float) * T * b.dim_out, cudaMemcpyDeviceToDevice, st));
      }
      resid = xo; dst = xo;
    }
    OVO_TRY(gemm(EPI_BF16, s->xn, b.dim, b.qkv_w, b.dim, T, 3 * b.dim_out, b.dim, b.qkv_b, s->qkv, 3 * b.dim_out, nullptr, 0, 0, st));
    {
      WinAttnParams p;
      p.qkv = s->qkv; p.out = s->att; p.grid = grid; p.ws = ws; p.heads = b.heads; p.dim_out = b.dim_out; p.q_pool = b.q_pool;
      p.scale_log2e = 1.4426950408889634f / sqrtf(static_cast<float>(kSamHd));
      const int nq = b.q_pool ? ws * ws / 4 : ws * ws;
      const int wins = (grid / ws) * (grid / ws);
      p.wins2 = wins;
      ProfScope prof(st, PROF_ATTN, 4.0 * B * wins * b.heads * static_cast<double>(nq) * ws * ws * kSamHd, 0.0);
      if (nq >= g_hiera_qb128_min) {
        hiera_attention_kernel<128><<<dim3(ceil_div(nq, 128), B * wins, b.heads), 256, hiera_attn_smem_bytes<128>(), st>>>(p);
      } else {
        hiera_attention_kernel<64><<<dim3(ceil_div(nq, 64), B * wins, b.heads), 128, hiera_attn_smem_bytes<64>(), st>>>(p);
      }
      OVO_CHECK_LAUNCH();
    }
    const int To = B * grid_out * grid_out;
    OVO_TRY(gemm(EPI_F32_RESID, s->att, b.dim_out, b.proj_w, b.dim_out, To, b.dim_out, b.dim_out, b.proj_b, dst, b.dim_out, resid, b.dim_out, 0, st));
    if (dst != x) std::swap(x, xo);
    grid = grid_out;
    OVO_TRY(ln(x, To, b.dim_out, b.norm2_w, b.norm2_b, c.trunk_ln_eps, nullptr, s->xn, nullptr, nullptr, 1, st));
    OVO_TRY(gemm(EPI_BF16_GELU, s->xn, b.dim_out, b.fc1_w, b.dim_out, To, 4 * b.dim_out, b.dim_out, b.fc1_b, s->hid, 4 * b.dim_out, nullptr, 0, 0, st));
    OVO_TRY(gemm(EPI_F32_RESID, s->hid, 4 * b.dim_out, b.fc2_w, 4 * b.dim_out, To, b.dim_out, 4 * b.dim_out, b.fc2_b, x, b.dim_out, x, b.dim_out, 0, st));
    last_dim = b.dim_out;
    if (stage < 4 && i == c.stage_end[stage]) {
      OVO_REQUIRE(s->stage_grid[stage] == grid && s->stage_dim[stage] == b.dim_out, "sam: stage %d geometry mismatch", stage);
      OVO_TRY(add_cast(x, nullptr, 1, b.dim_out, static_cast<size_t>(To) * b.dim_out, s->stage_bf[stage], nullptr, st));
      ++stage;
    }
  }
  if (block_out) OVO_CUDA(cudaMemcpyAsync(block_out, x, sizeof(float) * grid * grid * last_dim, cudaMemcpyDeviceToDevice, st));   // image 0
  if (nb < c.n_blocks) return OVO_OK;
  // ---- neck (image_encoder.py:102-134) with conv_s0/conv_s1 folded (sam2_base.py:467-479)
  const int g = s->g;
  OVO_TRY(gemm(EPI_F32, s->stage_bf[3], s->stage_dim[3], s->w.neck3_w, s->stage_dim[3], B * (g / 2) * (g / 2), kC, s->stage_dim[3], s->w.neck3_b, s->lat3, kC, nullptr, 0, 0, st));
  OVO_TRY(gemm(EPI_F32, s->stage_bf[2], s->stage_dim[2], s->w.neck2_w, s->stage_dim[2], B * g * g, kC, s->stage_dim[2], s->w.neck2_b, s->lat2, kC, nullptr, 0, 0, st));
  OVO_TRY(gemm(EPI_F32, s->stage_bf[1], s->stage_dim[1], s->w.s1_w, s->stage_dim[1], B * 4 * g * g, 64, s->stage_dim[1], s->w.s1_b, s->feat_s1, 64, nullptr, 0, 0, st));
  OVO_TRY(gemm(EPI_F32, s->stage_bf[0], s->stage_dim[0], s->w.s0_w, s->stage_dim[0], B * 16 * g * g, 32, s->stage_dim[0], s->w.s0_b, s->feat_s0, 32, nullptr, 0, 0, st));
  sam_embed_kernel<<<dim3(ceil_div(g * g * kC, 256), B), 256, 0, st>>>(s->lat2, s->lat3, g, s->w.no_mask_embed, s->w.dense_pe, s->embed, s->src, s->src_bf, s->srcpe_bf);
  OVO_CHECK_LAUNCH();
  const size_t HWs = static_cast<size_t>(g) * g;
  for (int b = 0; b < B; ++b) {
    sam_subpixel_kernel<<<ceil_div(static_cast<long long>(g) * g * 4 * 64, 256), 256, 0, st>>>(s->feat_s1 + b * 4 * HWs * 64, g, 64, s->s1_sub + b * 4 * HWs * 64);
    OVO_CHECK_LAUNCH();
    sam_subpixel2_kernel<<<ceil_div(static_cast<long long>(g) * g * 16 * 32, 256), 256, 0, st>>>(s->feat_s0 + b * 16 * HWs * 32, g, s->s0_sub + b * 16 * HWs * 32);
    OVO_CHECK_LAUNCH();
  }
  // layer-0 projections of the (prompt independent) image side: k, v of token->image and q of image->token
  const ovo_sam_dec_layer& L0 = s->layers[0];
  const int HW = g * g;
  OVO_TRY(gemm(EPI_BF16, s->srcpe_bf, kC, L0.t2i.k_w, kC, B * HW, kInt, kC, L0.t2i.k_b, s->k0, kInt, nullptr, 0, 0, st));
  OVO_TRY(gemm(EPI_BF16, s->src_bf, kC, L0.t2i.v_w, kC, B * HW, kInt, kC, L0.t2i.v_b, s->v0, kInt, nullptr, 0, 0, st));
  OVO_TRY(gemm(EPI_BF16, s->srcpe_bf, kC, L0.i2t.q_w, kC, B * HW, kInt, kC, L0.i2t.q_b, s->qi0, kInt, nullptr, 0, 0, st));
  return OVO_OK;
}

int copy_taps(ovo_sam* s, float* embed, float* s0, float* s1, cudaStream_t st) {
  const int g = s->g;
  const size_t HW = static_cast<size_t>(g) * g, b = s->cur;
  if (embed) OVO_CUDA(cudaMemcpyAsync(embed, s->embed + b * HW * kC, sizeof(float) * HW * kC, cudaMemcpyDeviceToDevice, st));
  if (s0) OVO_CUDA(cudaMemcpyAsync(s0, s->feat_s0 + b * 16 * HW * 32, sizeof(float) * 16 * HW * 32, cudaMemcpyDeviceToDevice, st));
  if (s1) OVO_CUDA(cudaMemcpyAsync(s1, s->feat_s1 + b * 4 * HW * 64, sizeof(float) * 4 * HW * 64, cudaMemcpyDeviceToDevice, st));
  return OVO_OK;
}

int patches_from_pixels(ovo_sam* s, int B, cudaStream_t st) {
  ProfScope prof(st, PROF_PRE, 0.0, 0.0);
  sam_im2col_kernel<<<B * (s->S / 4) * (s->S / 4), 160, 0, st>>>(s->pixels, s->S, s->S / 4, s->w.patch_kpad, s->patches);
  OVO_CHECK_LAUNCH();
  return OVO_OK;
}

int trunk_from_pixels(ovo_sam* s, int B, int n_blocks, float* block_out, cudaStream_t st) {
  s->n_set = B; s->cur = 0;
  if ((n_blocks < 0 || n_blocks >= s->cfg.n_blocks) && block_out == nullptr)
    return graphed(s, 1 | (static_cast<long long>(B) << 8), st, [&](cudaStream_t cs) {
      OVO_TRY(patches_from_pixels(s, B, cs));
      return run_trunk(s, B, -1, nullptr, cs);
    });
  OVO_TRY(patches_from_pixels(s, B, st));
  return run_trunk(s, B, n_blocks, block_out, st);
}

// Attention.forward on the token side: out_proj(attn(...)) handled by the caller; this projects q/k/v of tokens.
int tok_lin(ovo_sam* s, const __nv_bfloat16* a, const void* w, const float* b, int rows, int n, int k, __nv_bfloat16* out, cudaStream_t st) {
  return gemm(EPI_BF16, a, k, w, k, rows, n, k, b, out, n, nullptr, 0, 0, st);
}

// token -> image attention block: queries = LN(queries + out_proj(attn(q = queries+pe, k, v)))
int t2i_block(ovo_sam* s, const ovo_sam_attn& A, const __nv_bfloat16* kbuf, const __nv_bfloat16* vbuf, size_t kv_stride, int P,
              const float* nw, const float* nb, cudaStream_t st) {
  const int R = P * kTok, HW = s->g * s->g;
  OVO_TRY(add_cast(s->queries, s->tokens, R, kC, static_cast<size_t>(R) * kC, s->tqpe_bf, nullptr, st));
  OVO_TRY(tok_lin(s, s->tqpe_bf, A.q_w, A.q_b, R, kInt, kC, s->t_q, st));
  {
    ProfScope prof(st, PROF_ATTN, 4.0 * P * 8 * kTok * static_cast<double>(HW) * 16, 0.0);
    OVO_REQUIRE(HW % 64 == 0, "sam decoder: image token count %d must be a multiple of 64", HW);
    sam_t2i_attn_mma_kernel<<<P, 256, kT2iSmemBytes, st>>>(s->t_q, kbuf, vbuf, kv_stride, HW, s->t_o);
    OVO_CHECK_LAUNCH();
  }
  OVO_TRY(gemm(EPI_F32_RESID, s->t_o, kInt, A.o_w, kInt, R, kC, kInt, A.o_b, s->tq_tmp, kC, s->queries, kC, 0, st));
  return ln(s->tq_tmp, R, kC, nw, nb, 1e-5f, s->queries, s->tq_bf, nullptr, nullptr, 1, st);
}

}  // namespace

extern "C" {

int ovo_sam_create(const ovo_sam_cfg* cfg, const ovo_sam_weights* w, int max_h, int max_w, int max_prompts, ovo_sam_t** out) {
  OVO_REQUIRE(cfg && w && out, "ovo_sam_create: null argument");
  OVO_REQUIRE(cfg->image_size % 64 == 0 && cfg->n_blocks > 0 && cfg->decoder_depth >= 1, "ovo_sam_create: bad config");
  OVO_REQUIRE(w->patch_kpad % 8 == 0 && w->patch_kpad >= 147, "ovo_sam_create: patch_kpad must be a multiple of 8 >= 147");
  keep_default_mempool_cached();
  ovo_sam* s = new ovo_sam();
  s->cfg = *cfg; s->w = *w;
  s->blocks.assign(w->blocks, w->blocks + cfg->n_blocks);
  s->layers.assign(w->layers, w->layers + cfg->decoder_depth);
  s->w.blocks = s->blocks.data(); s->w.layers = s->layers.data();
  s->S = cfg->image_size; s->g = cfg->image_size / 16;
  s->max_h = max_h; s->max_w = max_w; s->max_p = max_prompts;
  s->max_batch = cfg->max_batch > 1 ? cfg->max_batch : 1;
  {
    const char* env = getenv("OVO_B200_GRAPHS");
    s->use_graphs = !(env && env[0] == '0');
    const char* q = getenv("OVO_B200_HIERA_QB128_MIN");
    if (q) g_hiera_qb128_min = atoi(q);
  }
  const int S = s->S, g = s->g, g0 = S / 4;
  // geometry walk: buffer sizes and the stage outputs
  size_t max_x = 0, max_qkv = 0, max_hid = 0, max_short = 0, max_xn = 0, max_att = 0;
  {
    int grid = g0, stage = 0;
    for (int i = 0; i < cfg->n_blocks; ++i) {
      const ovo_hiera_block& b = s->blocks[i];
      const size_t T = static_cast<size_t>(grid) * grid;
      const int go = b.q_pool ? grid / 2 : grid;
      const size_t To = static_cast<size_t>(go) * go;
      max_x = std::max({max_x, T * b.dim, To * b.dim_out});
      max_xn = std::max({max_xn, T * b.dim, To * b.dim_out});
      max_qkv = std::max(max_qkv, T * 3 * b.dim_out);
      max_att = std::max(max_att, To * b.dim_out);
      max_hid = std::max(max_hid, To * 4 * b.dim_out);
      if (b.short_w) max_short = std::max(max_short, T * b.dim_out);
      grid = go;
      if (stage < 4 && i == cfg->stage_end[stage]) { s->stage_grid[stage] = grid; s->stage_dim[stage] = b.dim_out; ++stage; }
    }
    if (stage != 4 || s->stage_grid[2] != g || s->stage_grid[3] != g / 2 || s->stage_grid[1] != 2 * g || s->stage_grid[0] != 4 * g) {
      delete s;
      return set_error(OVO_E_INVALID, "ovo_sam_create: the trunk must have four stages at strides 4/8/16/32");
    }
  }
  int r = OVO_OK;
  auto A = [&](auto** p, size_t n) { if (r == OVO_OK) r = salloc(s, p, n); };
  const size_t MB = static_cast<size_t>(s->max_batch);
  A(&s->tmp, static_cast<size_t>(3) * max_h * S); A(&s->pixels, MB * 3 * S * S);
  A(&s->patches, MB * g0 * g0 * w->patch_kpad);
  A(&s->xa, MB * max_x); A(&s->xb, MB * max_x); A(&s->tshort, MB * max_short); A(&s->xn, MB * max_xn); A(&s->qkv, MB * max_qkv); A(&s->att, MB * max_att); A(&s->hid, MB * max_hid);
  for (int k = 0; k < 4; ++k) A(&s->stage_bf[k], MB * s->stage_grid[k] * s->stage_grid[k] * s->stage_dim[k]);
  const size_t HW = static_cast<size_t>(g) * g;
  A(&s->lat3, MB * HW / 4 * kC); A(&s->lat2, MB * HW * kC); A(&s->embed, MB * HW * kC); A(&s->feat_s0, MB * 16 * HW * 32); A(&s->feat_s1, MB * 4 * HW * 64);
  A(&s->s0_sub, MB * 16 * HW * 32); A(&s->s1_sub, MB * 4 * HW * 64); A(&s->src, MB * HW * kC); A(&s->src_bf, MB * HW * kC); A(&s->srcpe_bf, MB * HW * kC);
  A(&s->k0, MB * HW * kInt); A(&s->v0, MB * HW * kInt); A(&s->qi0, MB * HW * kInt);
  const size_t P = max_prompts, R = P * kTok;
  A(&s->tokens, R * kC); A(&s->queries, R * kC); A(&s->tq_tmp, R * kC); A(&s->tq_bf, R * kC); A(&s->tqpe_bf, R * kC);
  A(&s->t_q, R * kC); A(&s->t_k, R * kC); A(&s->t_v, R * kC); A(&s->t_o, R * kC); A(&s->t_mlp, R * 2048);
  A(&s->keys, P * HW * kC); A(&s->keys_pre, P * HW * kC); A(&s->keys_bf, P * HW * kC); A(&s->keyspe_bf, P * HW * kC);
  A(&s->big_q, P * HW * kInt); A(&s->big_k, P * HW * kInt); A(&s->big_v, P * HW * kInt);
  A(&s->up1, P * HW * 4 * 64);
  A(&s->hyper, P * 4 * 32); A(&s->iou_head, P * 4); A(&s->tok_a, P * kC); A(&s->tok_b, P * kC);
  A(&s->stats, P * 3); A(&s->cand, P * 3); A(&s->sel, P * 3);
  A(&s->low_all, P * 3 * 16 * HW); A(&s->iou_all, P * 3); A(&s->points, P * 2);
  A(&s->score_tmp, P * 3); A(&s->keep_tmp, P * 3); A(&s->stab_tmp, P * 3); A(&s->iou_tmp, P * 3); A(&s->box_tmp, P * 3 * 4);
  A(&s->src_tmp, P * 3); A(&s->order_tmp, P * 3); A(&s->counters, 4);
  if (r != OVO_OK) { ovo_sam_destroy(s); return r; }
  if (cudaFuncSetAttribute(hiera_attention_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, hiera_attn_smem_bytes<128>()) != cudaSuccess ||
      cudaFuncSetAttribute(hiera_attention_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, hiera_attn_smem_bytes<64>()) != cudaSuccess ||
      cudaFuncSetAttribute(sam_t2i_attn_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kT2iSmemBytes) != cudaSuccess) {
    ovo_sam_destroy(s);
    return set_error(OVO_E_CUDA, "ovo_sam_create: cudaFuncSetAttribute(shared memory) failed: %s", cudaGetErrorString(cudaGetLastError()));
  }
  *out = s;
  return OVO_OK;
}

void ovo_sam_destroy(ovo_sam_t* s) {
  if (!s) return;
  for (auto& ge : s->graphs) if (ge.second.exec) cudaGraphExecDestroy(ge.second.exec);
  if (s->cap_stream) cudaStreamDestroy(s->cap_stream);
  for (void* p : s->owned) cudaFree(p);
  for (void* p : {static_cast<void*>(s->xmin), static_cast<void*>(s->xsize), static_cast<void*>(s->ymin), static_cast<void*>(s->ysize),
                  static_cast<void*>(s->xw), static_cast<void*>(s->yw), static_cast<void*>(s->masks_tmp), static_cast<void*>(s->masks_tmp2)})
    if (p) cudaFree(p);
  delete s;
}

int ovo_sam_set_pixels(ovo_sam_t* s, const float* pixels_dev, float* embed_out, float* s0_out, float* s1_out, int n_blocks,
                       float* block_out, void* stream_) {
  OVO_REQUIRE(s && pixels_dev, "ovo_sam_set_pixels: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  OVO_CUDA(cudaMemcpyAsync(s->pixels, pixels_dev, sizeof(float) * 3 * s->S * s->S, cudaMemcpyDeviceToDevice, st));
  OVO_TRY(trunk_from_pixels(s, 1, n_blocks, block_out, st));
  if (n_blocks < 0 || n_blocks >= s->cfg.n_blocks) OVO_TRY(copy_taps(s, embed_out, s0_out, s1_out, st));
  return OVO_OK;
}

// SAM2Transforms.__call__ for one frame -> s->pixels[slot]
static int resize_to_pixels(ovo_sam_t* s, const uint8_t* rgb_dev, int H, int W, int slot, cudaStream_t st) {
  const int S = s->S;
  if (s->tab_h != H || s->tab_w != W) {   // resize tables for this frame size (host, once per size)
    OVO_CUDA(cudaStreamSynchronize(st));
    for (void* p : {static_cast<void*>(s->xmin), static_cast<void*>(s->xsize), static_cast<void*>(s->ymin), static_cast<void*>(s->ysize),
                    static_cast<void*>(s->xw), static_cast<void*>(s->yw)})
      if (p) cudaFree(p);
    std::vector<int> mn, sz; std::vector<float> ws;
    aa_axis(W, S, mn, sz, ws, &s->kx);
    OVO_TRY(dalloc(&s->xmin, S)); OVO_TRY(dalloc(&s->xsize, S)); OVO_TRY(dalloc(&s->xw, ws.size()));
    OVO_CUDA(cudaMemcpy(s->xmin, mn.data(), sizeof(int) * S, cudaMemcpyHostToDevice));
    OVO_CUDA(cudaMemcpy(s->xsize, sz.data(), sizeof(int) * S, cudaMemcpyHostToDevice));
    OVO_CUDA(cudaMemcpy(s->xw, ws.data(), sizeof(float) * ws.size(), cudaMemcpyHostToDevice));
    aa_axis(H, S, mn, sz, ws, &s->ky);
    OVO_TRY(dalloc(&s->ymin, S)); OVO_TRY(dalloc(&s->ysize, S)); OVO_TRY(dalloc(&s->yw, ws.size()));
    OVO_CUDA(cudaMemcpy(s->ymin, mn.data(), sizeof(int) * S, cudaMemcpyHostToDevice));
    OVO_CUDA(cudaMemcpy(s->ysize, sz.data(), sizeof(int) * S, cudaMemcpyHostToDevice));
    OVO_CUDA(cudaMemcpy(s->yw, ws.data(), sizeof(float) * ws.size(), cudaMemcpyHostToDevice));
    s->tab_h = H; s->tab_w = W;
  }
  ProfScope prof(st, PROF_PRE, 0.0, static_cast<double>(H) * W * 3 + 3.0 * S * S * 4);
  sam_resize_h_kernel<<<dim3(ceil_div(S, 128), H), 128, 0, st>>>(rgb_dev, H, W, s->xmin, s->xsize, s->xw, s->kx, S, s->tmp);
  OVO_CHECK_LAUNCH();
  sam_resize_v_kernel<<<dim3(ceil_div(S, 128), S, 3), 128, 0, st>>>(s->tmp, H, s->ymin, s->ysize, s->yw, s->ky, S, s->pixels + static_cast<size_t>(slot) * 3 * S * S);
  OVO_CHECK_LAUNCH();
  return OVO_OK;
}

int ovo_sam_set_image(ovo_sam_t* s, const uint8_t* rgb_dev, int H, int W, float* pixels_out, float* embed_out, float* s0_out,
                      float* s1_out, int n_blocks, float* block_out, void* stream_) {
  OVO_REQUIRE(s && rgb_dev, "ovo_sam_set_image: null argument");
  OVO_REQUIRE(H > 0 && W > 0 && H <= s->max_h && W <= s->max_w, "ovo_sam_set_image: frame %dx%d exceeds the %dx%d the handle was created for", H, W, s->max_h, s->max_w);
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  const int S = s->S;
  OVO_TRY(resize_to_pixels(s, rgb_dev, H, W, 0, st));
  if (pixels_out) OVO_CUDA(cudaMemcpyAsync(pixels_out, s->pixels, sizeof(float) * 3 * S * S, cudaMemcpyDeviceToDevice, st));
  OVO_TRY(trunk_from_pixels(s, 1, n_blocks, block_out, st));
  if (n_blocks < 0 || n_blocks >= s->cfg.n_blocks) OVO_TRY(copy_taps(s, embed_out, s0_out, s1_out, st));
  return OVO_OK;
}

int ovo_sam_set_images(ovo_sam_t* s, const uint8_t* rgb_dev, int n, int H, int W, void* stream_) {
  OVO_REQUIRE(s && rgb_dev, "ovo_sam_set_images: null argument");
  OVO_REQUIRE(n > 0 && n <= s->max_batch, "ovo_sam_set_images: %d frames, handle created for batches of %d", n, s->max_batch);
  OVO_REQUIRE(H > 0 && W > 0 && H <= s->max_h && W <= s->max_w, "ovo_sam_set_images: frame %dx%d exceeds the %dx%d the handle was created for", H, W, s->max_h, s->max_w);
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  for (int b = 0; b < n; ++b) OVO_TRY(resize_to_pixels(s, rgb_dev + static_cast<size_t>(b) * H * W * 3, H, W, b, st));
  return trunk_from_pixels(s, n, -1, nullptr, st);
}

int ovo_sam_select_image(ovo_sam_t* s, int index) {
  OVO_REQUIRE(s && index >= 0 && index < s->n_set, "ovo_sam_select_image: index %d outside the %d images set", index, s ? s->n_set : 0);
  s->cur = index;
  return OVO_OK;
}

static int sam_predict_impl(ovo_sam_t* s, const float* points_dev, int P, float* low_out, float* iou_out, cudaStream_t st);

int ovo_sam_predict(ovo_sam_t* s, const float* points_dev, int P, float* low_out, float* iou_out, void* stream_) {
  OVO_REQUIRE(s && points_dev, "ovo_sam_predict: null argument");
  OVO_REQUIRE(P > 0 && P <= s->max_p, "ovo_sam_predict: %d prompts, handle sized for %d", P, s->max_p);
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  if (points_dev == s->points && low_out == s->low_all && iou_out == s->iou_all)   // the handle's own buffers: static launch sequence
    return graphed(s, (2ll << 32) | (static_cast<long long>(s->cur) << 16) | P, st, [&](cudaStream_t cs) { return sam_predict_impl(s, points_dev, P, low_out, iou_out, cs); });
  return sam_predict_impl(s, points_dev, P, low_out, iou_out, st);
}

static int sam_predict_impl(ovo_sam_t* s, const float* points_dev, int P, float* low_out, float* iou_out, cudaStream_t st) {
  const int g = s->g, HW = g * g, R = P * kTok;
  const size_t PHW = static_cast<size_t>(P) * HW;
  OVO_REQUIRE(s->n_set > 0 && s->cur < s->n_set, "ovo_sam_predict: no image set");
  // per-image features of the selected image (ovo_sam_select_image)
  const size_t slot = static_cast<size_t>(s->cur);
  const float* f_src = s->src + slot * HW * kC;
  const __nv_bfloat16* f_k0 = s->k0 + slot * HW * kInt;
  const __nv_bfloat16* f_v0 = s->v0 + slot * HW * kInt;
  const __nv_bfloat16* f_qi0 = s->qi0 + slot * HW * kInt;
  const float* f_s1_sub = s->s1_sub + slot * 4 * HW * 64;
  const float* f_s0_sub = s->s0_sub + slot * 16 * HW * 32;
  OVO_REQUIRE(PHW * 4 < (1ull << 31), "ovo_sam_predict: too many prompts for 32-bit GEMM row indices");
  sam_tokens_kernel<<<P, 256, 0, st>>>(points_dev, P, static_cast<float>(s->S), s->w.gauss, s->w.point_embed, s->w.not_a_point, s->w.out_tokens, s->tokens);
  OVO_CHECK_LAUNCH();
  const int depth = s->cfg.decoder_depth;
  for (int l = 0; l < depth; ++l) {
    const ovo_sam_dec_layer& L = s->layers[l];
    // (1) token self attention (sam/transformer.py:183-190)
    if (l == 0) {
      OVO_TRY(add_cast(s->tokens, nullptr, 1, kC, static_cast<size_t>(R) * kC, s->tq_bf, nullptr, st));
      OVO_TRY(tok_lin(s, s->tq_bf, L.self_attn.q_w, L.self_attn.q_b, R, kC, kC, s->t_q, st));
      OVO_TRY(tok_lin(s, s->tq_bf, L.self_attn.k_w, L.self_attn.k_b, R, kC, kC, s->t_k, st));
    } else {
      OVO_TRY(add_cast(s->queries, s->tokens, R, kC, static_cast<size_t>(R) * kC, s->tqpe_bf, nullptr, st));
      OVO_TRY(tok_lin(s, s->tqpe_bf, L.self_attn.q_w, L.self_attn.q_b, R, kC, kC, s->t_q, st));
      OVO_TRY(tok_lin(s, s->tqpe_bf, L.self_attn.k_w, L.self_attn.k_b, R, kC, kC, s->t_k, st));
    }
    OVO_TRY(tok_lin(s, s->tq_bf, L.self_attn.v_w, L.self_attn.v_b, R, kC, kC, s->t_v, st));
    sam_self_attn_kernel<<<P, 256, 0, st>>>(s->t_q, s->t_k, s->t_v, s->t_o);
    OVO_CHECK_LAUNCH();
    OVO_TRY(gemm(l == 0 ? EPI_F32 : EPI_F32_RESID, s->t_o, kC, L.self_attn.o_w, kC, R, kC, kC, L.self_attn.o_b, s->tq_tmp, kC,
                 l == 0 ? nullptr : s->queries, kC, 0, st));
    OVO_TRY(ln(s->tq_tmp, R, kC, L.norm_w[0], L.norm_b[0], 1e-5f, s->queries, s->tq_bf, nullptr, nullptr, 1, st));
    // (2) token -> image cross attention (:192-197)
    if (l == 0) {
      OVO_TRY(t2i_block(s, L.t2i, f_k0, f_v0, 0, P, L.norm_w[1], L.norm_b[1], st));
    } else {
      OVO_TRY(gemm(EPI_BF16, s->keyspe_bf, kC, L.t2i.k_w, kC, static_cast<int>(PHW), kInt, kC, L.t2i.k_b, s->big_k, kInt, nullptr, 0, 0, st));
      OVO_TRY(gemm(EPI_BF16, s->keys_bf, kC, L.t2i.v_w, kC, static_cast<int>(PHW), kInt, kC, L.t2i.v_b, s->big_v, kInt, nullptr, 0, 0, st));
      OVO_TRY(t2i_block(s, L.t2i, s->big_k, s->big_v, static_cast<size_t>(HW) * kInt, P, L.norm_w[1], L.norm_b[1], st));
    }
    // (3) MLP on the tokens (:199-202)
    OVO_TRY(gemm(EPI_BF16_RELU, s->tq_bf, kC, L.mlp0_w, kC, R, 2048, kC, L.mlp0_b, s->t_mlp, 2048, nullptr, 0, 0, st));
    OVO_TRY(gemm(EPI_F32_RESID, s->t_mlp, 2048, L.mlp1_w, 2048, R, kC, 2048, L.mlp1_b, s->tq_tmp, kC, s->queries, kC, 0, st));
    OVO_TRY(ln(s->tq_tmp, R, kC, L.norm_w[2], L.norm_b[2], 1e-5f, s->queries, s->tq_bf, nullptr, nullptr, 1, st));
    // (4) image -> token cross attention (:204-210)
    OVO_TRY(add_cast(s->queries, s->tokens, R, kC, static_cast<size_t>(R) * kC, s->tqpe_bf, nullptr, st));
    OVO_TRY(tok_lin(s, s->tqpe_bf, L.i2t.k_w, L.i2t.k_b, R, kInt, kC, s->t_k, st));
    OVO_TRY(tok_lin(s, s->tq_bf, L.i2t.v_w, L.i2t.v_b, R, kInt, kC, s->t_v, st));
    const __nv_bfloat16* qsrc = f_qi0;
    size_t qstride = 0;
    if (l > 0) {
      OVO_TRY(gemm(EPI_BF16, s->keyspe_bf, kC, L.i2t.q_w, kC, static_cast<int>(PHW), kInt, kC, L.i2t.q_b, s->big_q, kInt, nullptr, 0, 0, st));
      qsrc = s->big_q; qstride = static_cast<size_t>(HW) * kInt;
    }
    {
      ProfScope prof(st, PROF_ATTN, 4.0 * static_cast<double>(PHW) * 8 * kTok * 16, 0.0);
      sam_i2t_attn_kernel<<<dim3(ceil_div(HW, 32), P), 256, 0, st>>>(qsrc, qstride, s->t_k, s->t_v, HW, s->big_k /* out */);
      OVO_CHECK_LAUNCH();
    }
    OVO_TRY(gemm(EPI_F32_RESID, s->big_k, kInt, L.i2t.o_w, kInt, static_cast<int>(PHW), kC, kInt, L.i2t.o_b, s->keys_pre, kC,
                 l == 0 ? f_src : s->keys, kC, l == 0 ? HW : 0, st));
    // (the f32 copy is only the NEXT layer's residual: the last layer skips it — 1 GB less traffic at 256 prompts.  Fusing this
    //  LayerNorm into the out-projection's epilogue was measured: the two-pass epilogue is slower than this 74 %-of-DRAM-peak pass.)
    OVO_TRY(ln(s->keys_pre, static_cast<int>(PHW), kC, L.norm_w[3], L.norm_b[3], 1e-5f, l + 1 < depth ? s->keys : nullptr, s->keys_bf, s->keyspe_bf, s->w.dense_pe, HW, st));
  }
  // final token -> image attention + norm (sam/transformer.py:124-132)
  OVO_TRY(gemm(EPI_BF16, s->keyspe_bf, kC, s->w.final_attn.k_w, kC, static_cast<int>(PHW), kInt, kC, s->w.final_attn.k_b, s->big_k, kInt, nullptr, 0, 0, st));
  OVO_TRY(gemm(EPI_BF16, s->keys_bf, kC, s->w.final_attn.v_w, kC, static_cast<int>(PHW), kInt, kC, s->w.final_attn.v_b, s->big_v, kInt, nullptr, 0, 0, st));
  OVO_TRY(t2i_block(s, s->w.final_attn, s->big_k, s->big_v, static_cast<size_t>(HW) * kInt, P, s->w.norm_final_w, s->w.norm_final_b, st));
  // hs = s->queries.  Hypernetwork MLPs on the 4 mask tokens (mask_decoder.py:219-224) and the IoU head (:229) first: the
  // mask product is fused into the last up-scaling GEMM below.
  for (int i = 0; i < 4; ++i) {
    sam_pick_token_kernel<<<ceil_div(P * kC, 256), 256, 0, st>>>(s->queries, P, 2 + i, s->tok_a);
    OVO_CHECK_LAUNCH();
    OVO_TRY(gemm(EPI_BF16_RELU, s->tok_a, kC, s->w.hyper_w[i][0], kC, P, kC, kC, s->w.hyper_b[i][0], s->tok_b, kC, nullptr, 0, 0, st));
    OVO_TRY(gemm(EPI_BF16_RELU, s->tok_b, kC, s->w.hyper_w[i][1], kC, P, kC, kC, s->w.hyper_b[i][1], s->tok_a, kC, nullptr, 0, 0, st));
    OVO_TRY(gemm(EPI_F32, s->tok_a, kC, s->w.hyper_w[i][2], kC, P, 32, kC, s->w.hyper_b[i][2], s->hyper + i * 32, 128, nullptr, 0, 0, st));
  }
  sam_pick_token_kernel<<<ceil_div(P * kC, 256), 256, 0, st>>>(s->queries, P, 1, s->tok_a);
  OVO_CHECK_LAUNCH();
  OVO_TRY(gemm(EPI_BF16_RELU, s->tok_a, kC, s->w.iou_w[0], kC, P, 256, kC, s->w.iou_b[0], s->tok_b, 256, nullptr, 0, 0, st));
  OVO_TRY(gemm(EPI_BF16_RELU, s->tok_b, 256, s->w.iou_w[1], 256, P, 256, 256, s->w.iou_b[1], s->tok_a, 256, nullptr, 0, 0, st));
  OVO_TRY(gemm(EPI_F32, s->tok_a, 256, s->w.iou_w[2], 256, P, 4, 256, s->w.iou_b[2], s->iou_head, 4, nullptr, 0, 0, st));
  float* low = low_out ? low_out : s->low_all;
  float* iou = iou_out ? iou_out : s->iou_all;
  // Up-scaling (mask_decoder.py:210-217): the two k2 s2 transposed convs are GEMMs with N = 4*C_out whose epilogues do the
  // rest — (1) + feat_s1, LayerNorm2d over each sub-pixel's 64 channels, GELU -> up1 bf16 [P*g*g, (sub1, 64)];
  // (2) + feat_s0, GELU, and the product with the hyper-network vectors (:225-226) -> mask logits.  Neither the
  // [P,64,2g,2g] nor the [P,32,4g,4g] up-scaled embedding is ever written to memory.
  {
    EpiParams ep;
    ep.out = s->up1; ep.ldo = 256; ep.bias = s->w.up0_b; ep.resid = f_s1_sub; ep.ldr = 256; ep.resid_mod = HW;
    ep.ln_w = s->w.up_ln_w; ep.ln_b = s->w.up_ln_b;
    OVO_TRY(launch_gemm(EPI_UP_LN, s->keys_bf, kC, static_cast<const __nv_bfloat16*>(s->w.up0_w), kC, static_cast<int>(PHW), 256, kC, ep, st));
  }
  {
    EpiParams ep;
    ep.bias = s->w.up1_b; ep.resid = f_s0_sub; ep.ldr = 128; ep.resid_mod = 4 * HW;
    ep.dot_w = s->hyper; ep.dot_out = low; ep.dot_g = g;
    OVO_TRY(launch_gemm(EPI_GELU_DOT, s->up1, 64, static_cast<const __nv_bfloat16*>(s->w.up1_w), 64, static_cast<int>(PHW * 4), 128, 64, ep, st));
  }
  {
    ProfScope prof(st, PROF_OTHER, 0.0, 0.0);
    sam_iou_kernel<<<ceil_div(P * 3, 256), 256, 0, st>>>(s->iou_head, P, iou);
    OVO_CHECK_LAUNCH();
  }
  return OVO_OK;
}


int ovo_sam_postprocess(ovo_sam_t* s, const float* low_dev, const float* iou_dev, int P, int h, int w, int H, int W,
                        const ovo_amg_params* prm, uint8_t* masks_out, float* iou_out, float* stab_out, int32_t* boxes_out,
                        int32_t* src_out, int max_out, int* n_out, void* stream_) {
  OVO_REQUIRE(s && low_dev && iou_dev && prm && masks_out && iou_out && stab_out && boxes_out && src_out && n_out, "ovo_sam_postprocess: null argument");
  OVO_REQUIRE(P > 0 && P <= s->max_p && P * 3 <= 1024, "ovo_sam_postprocess: %d prompts unsupported (at most %d and 341)", P, s->max_p);
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  const int n = P * 3;
  ProfScope prof(st, PROF_OTHER, 0.0, 0.0);
  amg_candidates_kernel<<<1, 1024, 0, st>>>(iou_dev, n, prm->pred_iou_thresh, s->cand, s->counters);
  OVO_CHECK_LAUNCH();
  int counts[2] = {0, 0};
  OVO_CUDA(cudaMemcpyAsync(counts, s->counters, sizeof(int), cudaMemcpyDeviceToHost, st));
  OVO_CUDA(cudaStreamSynchronize(st));
  const int n_cand = counts[0];
  if (n_cand == 0) { *n_out = 0; return OVO_OK; }
  amg_stats_init_kernel<<<ceil_div(n_cand, 256), 256, 0, st>>>(s->stats, n_cand, H, W);
  OVO_CHECK_LAUNCH();
  OVO_REQUIRE(W <= kAmgMaxW && h < 32768 && w < 32768, "ovo_sam_postprocess: frame width %d unsupported (max %d)", W, kAmgMaxW);
  const int tiles = std::max(1, std::min(ceil_div(H * W, 256 * 8), 64));
  amg_stats_kernel<<<dim3(std::min(H, 16), n_cand), 256, 0, st>>>(low_dev, s->cand, h, w, H, W, prm->stability_offset, s->stats);
  OVO_CHECK_LAUNCH();
  amg_decide_kernel<<<1, 1024, 0, st>>>(s->stats, s->cand, iou_dev, prm->stability_thresh, prm->box_nms_thresh, s->counters, s->sel,
                                       iou_out, stab_out, boxes_out, src_out, max_out);
  OVO_CHECK_LAUNCH();
  OVO_CUDA(cudaMemcpyAsync(counts, s->counters, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
  OVO_CUDA(cudaStreamSynchronize(st));
  const int K = counts[1];
  OVO_REQUIRE(K <= max_out, "ovo_sam_postprocess: %d masks survive, room for %d", K, max_out);
  *n_out = K;
  if (K == 0) return OVO_OK;
  amg_write_masks_kernel<<<dim3(tiles, K), 256, 0, st>>>(low_dev, s->sel, h, w, H, W, masks_out);
  OVO_CHECK_LAUNCH();
  return OVO_OK;
}

// grid prompts + decoder + AMG filters + OVO's second stage for the image selected in the handle
static int generate_selected(ovo_sam_t* s, int H, int W, const ovo_amg_params* prm, int32_t* seg_map_dev, uint8_t* masks_out_dev,
                             int max_masks, int* n_masks, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  const int n = prm->points_per_side, P = n * n;
  OVO_REQUIRE(n > 0 && P <= s->max_p && P * 3 <= 1024, "ovo_sam_generate: points_per_side %d unsupported (handle sized for %d prompts)", n, s->max_p);
  OVO_REQUIRE((static_cast<size_t>(H) * W) % 16 == 0, "ovo_sam_generate: H*W must be a multiple of 16");
  amg_points_kernel<<<ceil_div(P, 128), 128, 0, st>>>(n, H, W, static_cast<float>(s->S), s->points);
  OVO_CHECK_LAUNCH();
  // every filter of the AMG is per mask, so all prompts go through the decoder in one batch (the reference's batches of 64
  // only bound its memory, automatic_mask_generator.py:270-276)
  OVO_TRY(ovo_sam_predict(s, s->points, P, s->low_all, s->iou_all, stream_));
  if (s->low_override != nullptr) {   // the network ran; its (random-weight) logits are replaced before the AMG post-processing
    OVO_REQUIRE(s->override_p == P, "ovo_sam_generate: logit override holds %d prompts, the grid has %d", s->override_p, P);
    const size_t per = static_cast<size_t>(16) * s->g * s->g;
    OVO_CUDA(cudaMemcpyAsync(s->low_all, s->low_override, static_cast<size_t>(P) * 3 * per * sizeof(float), cudaMemcpyDeviceToDevice, st));
    OVO_CUDA(cudaMemcpyAsync(s->iou_all, s->iou_override, static_cast<size_t>(P) * 3 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  const size_t need = static_cast<size_t>(P) * 3 * H * W;
  if (s->masks_tmp_bytes < need) {
    OVO_CUDA(cudaStreamSynchronize(st));
    if (s->masks_tmp) cudaFree(s->masks_tmp);
    if (s->masks_tmp2) cudaFree(s->masks_tmp2);
    s->masks_tmp = nullptr; s->masks_tmp2 = nullptr; s->masks_tmp_bytes = 0;
    OVO_TRY(dalloc(&s->masks_tmp, need)); OVO_TRY(dalloc(&s->masks_tmp2, need));
    s->masks_tmp_bytes = need;
  }
  int K = 0;
  OVO_TRY(ovo_sam_postprocess(s, s->low_all, s->iou_all, P, 4 * s->g, 4 * s->g, H, W, prm, s->masks_tmp, s->iou_tmp, s->stab_tmp, s->box_tmp,
                              s->src_tmp, P * 3, &K, stream_));
  *n_masks = 0;
  if (K == 0) return OVO_OK;
  // OVO's second stage (mask_generator.py:113-119): masks_update (score = stability * predicted_iou) then mask2segmap
  mul_kernel<<<ceil_div(K, 256), 256, 0, st>>>(s->stab_tmp, s->iou_tmp, K, s->score_tmp);
  OVO_CHECK_LAUNCH();
  OVO_TRY(ovo_mask_nms(s->masks_tmp, s->score_tmp, K, H, W, prm->nms_iou_th, prm->nms_score_th, prm->nms_inner_th, s->keep_tmp, stream_));
  std::vector<uint8_t> keep(K);
  OVO_CUDA(cudaMemcpyAsync(keep.data(), s->keep_tmp, K, cudaMemcpyDeviceToHost, st));
  OVO_CUDA(cudaStreamSynchronize(st));
  std::vector<int> idx;
  for (int i = 0; i < K; ++i) if (keep[i]) idx.push_back(i);
  const int M = static_cast<int>(idx.size());
  OVO_REQUIRE(M <= max_masks, "ovo_sam_generate: %d masks, room for %d", M, max_masks);
  if (M == 0) return OVO_OK;
  OVO_CUDA(cudaMemcpyAsync(s->order_tmp, idx.data(), sizeof(int) * M, cudaMemcpyHostToDevice, st));
  const size_t npix16 = static_cast<size_t>(H) * W / 16;
  gather_masks_kernel<<<dim3(std::min<int>(64, ceil_div(static_cast<long long>(npix16), 256)), M), 256, 0, st>>>(s->masks_tmp, s->order_tmp, npix16, s->masks_tmp2);
  OVO_CHECK_LAUNCH();
  // stability of the kept masks, same order
  std::vector<float> stab(K);
  OVO_CUDA(cudaMemcpyAsync(stab.data(), s->stab_tmp, sizeof(float) * K, cudaMemcpyDeviceToHost, st));
  OVO_CUDA(cudaStreamSynchronize(st));
  std::vector<float> stab_kept(M);
  for (int i = 0; i < M; ++i) stab_kept[i] = stab[idx[i]];
  OVO_CUDA(cudaMemcpyAsync(s->score_tmp, stab_kept.data(), sizeof(float) * M, cudaMemcpyHostToDevice, st));
  OVO_TRY(ovo_mask2segmap(s->masks_tmp2, s->score_tmp, M, H, W, seg_map_dev, masks_out_dev, s->src_tmp, stream_));
  OVO_CUDA(cudaStreamSynchronize(st));   // idx / stab_kept are host temporaries of this call
  *n_masks = M;
  return OVO_OK;
}

int ovo_sam_override_logits(ovo_sam_t* s, const float* low_dev, const float* iou_dev, int n_prompts) {
  OVO_REQUIRE(s && ((low_dev && iou_dev && n_prompts > 0) || (!low_dev && !iou_dev)), "ovo_sam_override_logits: bad arguments");
  s->low_override = low_dev; s->iou_override = iou_dev; s->override_p = low_dev ? n_prompts : 0;
  return OVO_OK;
}

int ovo_sam_generate(ovo_sam_t* s, const uint8_t* rgb_dev, int H, int W, const ovo_amg_params* prm, int32_t* seg_map_dev,
                     uint8_t* masks_out_dev, int max_masks, int* n_masks, void* stream_) {
  OVO_REQUIRE(s && rgb_dev && prm && seg_map_dev && masks_out_dev && n_masks, "ovo_sam_generate: null argument");
  OVO_TRY(ovo_sam_set_image(s, rgb_dev, H, W, nullptr, nullptr, nullptr, nullptr, -1, nullptr, stream_));
  return generate_selected(s, H, W, prm, seg_map_dev, masks_out_dev, max_masks, n_masks, stream_);
}

int ovo_sam_generate_batch(ovo_sam_t* s, const uint8_t* rgb_dev, int n_frames, int H, int W, const ovo_amg_params* prm,
                           int32_t* seg_maps_dev, uint8_t* masks_out_dev, int max_masks, int* n_masks_host, void* stream_) {
  OVO_REQUIRE(s && rgb_dev && prm && seg_maps_dev && masks_out_dev && n_masks_host, "ovo_sam_generate_batch: null argument");
  OVO_TRY(ovo_sam_set_images(s, rgb_dev, n_frames, H, W, stream_));     // ONE trunk pass over all frames
  const size_t npix = static_cast<size_t>(H) * W;
  for (int b = 0; b < n_frames; ++b) {
    OVO_TRY(ovo_sam_select_image(s, b));
    OVO_TRY(generate_selected(s, H, W, prm, seg_maps_dev + b * npix, masks_out_dev + static_cast<size_t>(b) * max_masks * npix, max_masks,
                              n_masks_host + b, stream_));
  }
  return OVO_OK;
}

}  // extern "C"
