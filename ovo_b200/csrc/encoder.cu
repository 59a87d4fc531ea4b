// PE ViT region encoder + text tower (reference: thirdParty/perception_models/core/vision_encoder/pe.py,
// rope.py, transforms.py; ovo/entities/textregion.py) behind the C ABI of include/ovo_b200.h.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <vector>

#include "attention.cuh"
#include "generic_attention.cuh"
#include "common.cuh"
#include "gemm.cuh"

namespace ovo {

// =========================================================================================== small kernels
// LayerNorm over rows of `width` f32 (pe.py:183-184,347-348; eps 1e-5): one warp per row, two-pass in registers.
// Writes bf16 (GEMM A operand) and/or f32.
template <int kMaxPerLane>
__global__ void __launch_bounds__(256)
    layernorm_kernel(const float* __restrict__ x, int rows, int width, const float* __restrict__ g,
                     const float* __restrict__ b, float eps, __nv_bfloat16* __restrict__ out_bf16,
                     float* __restrict__ out_f32, const int* __restrict__ gather /* optional row indices */) {
  griddep_launch();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int src = gather ? gather[row] : row;
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(src) * width);
  const int nvec = width >> 2;
  float4 v[kMaxPerLane];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    const int idx = lane + 32 * i;
    if (idx < nvec) {
      v[i] = xr[idx];
      sum += v[i].x + v[i].y + v[i].z + v[i].w;
    }
  }
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / width;
  float var = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    const int idx = lane + 32 * i;
    if (idx < nvec) {
      const float a = v[i].x - mean, bb = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      var += a * a + bb * bb + c * c + d * d;
    }
  }
  for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
  const float rstd = rsqrtf(var / width + eps);
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    const int idx = lane + 32 * i;
    if (idx < nvec) {
      const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + idx);
      const float4 bb = __ldg(reinterpret_cast<const float4*>(b) + idx);
      float4 o;
      o.x = (v[i].x - mean) * rstd * gg.x + bb.x;
      o.y = (v[i].y - mean) * rstd * gg.y + bb.y;
      o.z = (v[i].z - mean) * rstd * gg.z + bb.z;
      o.w = (v[i].w - mean) * rstd * gg.w + bb.w;
      if (out_f32) reinterpret_cast<float4*>(out_f32 + static_cast<size_t>(row) * width)[idx] = o;
      if (out_bf16) {
        uint2 u;
        u.x = pack_bf16(o.x, o.y);
        u.y = pack_bf16(o.z, o.w);
        reinterpret_cast<uint2*>(out_bf16 + static_cast<size_t>(row) * width)[idx] = u;
      }
    }
  }
}

// x[b*(P+1)] = class_embedding + positional_embedding[0]   (pe.py:512-519)
__global__ void cls_rows_kernel(float* __restrict__ x, int n_img, int tokens, int width, const float* __restrict__ cls_pos0) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_img * width) return;
  const int b = i / width, d = i - b * width;
  x[static_cast<size_t>(b) * tokens * width + d] = cls_pos0[d];
}

// ------------------------------------------------------------------------------------------- E1 preprocess
struct ImgJob {  // one 336x336 output image cut from a frame
  int frame, y1, x1, h, w;   // source rectangle
  int tab_x, tab_y;          // offsets (in entries) of the per-axis tables
  int kx, ky;                // taps per output (table row length)
};

// horizontal anti-aliased pass: tmp[img][c][y][ox] = sum_k wx[ox][k] * src[y1+y][x1+xmin[ox]+k][c] / 255
__global__ void aa_resize_h_kernel(const uint8_t* __restrict__ rgb, int H, int W, const ImgJob* __restrict__ jobs,
                                   const int* __restrict__ tab_min, const int* __restrict__ tab_size,
                                   const float* __restrict__ tab_w, int S, float* __restrict__ tmp, int tmp_h, int slot0) {
  const ImgJob j = jobs[blockIdx.z];
  const int ox = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (ox >= S || y >= j.h) return;
  const int xmin = tab_min[j.tab_x + ox], n = tab_size[j.tab_x + ox];
  const float* w = tab_w + static_cast<size_t>(j.tab_x + ox) * j.kx;
  const uint8_t* src = rgb + ((static_cast<size_t>(j.frame) * H + j.y1 + y) * W + j.x1 + xmin) * 3;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  for (int k = 0; k < n; ++k) {
    const float wk = w[k];
    a0 += wk * (static_cast<float>(src[3 * k]) / 255.f);
    a1 += wk * (static_cast<float>(src[3 * k + 1]) / 255.f);
    a2 += wk * (static_cast<float>(src[3 * k + 2]) / 255.f);
  }
  float* dst = tmp + (static_cast<size_t>(blockIdx.z + slot0) * 3 * tmp_h + y) * S + ox;
  dst[0] = a0;
  dst[static_cast<size_t>(tmp_h) * S] = a1;
  dst[static_cast<size_t>(2) * tmp_h * S] = a2;
}

// vertical pass + Normalize(0.5, 0.5) + im2col to patch-major bf16 rows [img*P + patch][c*p*p + ky*p + kx]
__global__ void aa_resize_v_patch_kernel(const float* __restrict__ tmp, int tmp_h, const ImgJob* __restrict__ jobs,
                                         const int* __restrict__ tab_min, const int* __restrict__ tab_size,
                                         const float* __restrict__ tab_w, int S, int patch, int kpad,
                                         __nv_bfloat16* __restrict__ patches, int slot0) {
  const ImgJob j = jobs[blockIdx.z];
  const int slot = blockIdx.z + slot0;
  const int ox = blockIdx.x * blockDim.x + threadIdx.x;
  const int oy = blockIdx.y;
  if (ox >= S) return;
  const int ymin = tab_min[j.tab_y + oy], n = tab_size[j.tab_y + oy];
  const float* w = tab_w + static_cast<size_t>(j.tab_y + oy) * j.ky;
  const int grid = S / patch;
  const int prow = slot * grid * grid + (oy / patch) * grid + ox / patch;
  const int pcol = (oy % patch) * patch + ox % patch;
  for (int c = 0; c < 3; ++c) {
    const float* src = tmp + ((static_cast<size_t>(slot) * 3 + c) * tmp_h + ymin) * S + ox;
    float a = 0.f;
    for (int k = 0; k < n; ++k) a += w[k] * src[static_cast<size_t>(k) * S];
    a = (a - 0.5f) / 0.5f;
    patches[static_cast<size_t>(prow) * kpad + c * patch * patch + pcol] = __float2bfloat16_rn(a);
  }
}

// test tap: normalised f32 pixels [n,3,S,S] -> patch-major bf16
__global__ void pixels_to_patches_kernel(const float* __restrict__ px, int n_img, int S, int patch, int kpad,
                                         __nv_bfloat16* __restrict__ patches) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t total = static_cast<size_t>(n_img) * 3 * S * S;
  if (i >= total) return;
  const int ox = i % S, oy = (i / S) % S, c = (i / (static_cast<size_t>(S) * S)) % 3, img = i / (static_cast<size_t>(3) * S * S);
  const int grid = S / patch;
  const size_t prow = static_cast<size_t>(img) * grid * grid + (oy / patch) * grid + ox / patch;
  patches[prow * kpad + c * patch * patch + (oy % patch) * patch + ox % patch] = __float2bfloat16_rn(px[i]);
}

// ------------------------------------------------------------------------------------------- E3 canvas
// resize_features (textregion.py:9-28): bilinear up-sample (align_corners=False) of the global image's
// tokens to [ph,pw], then 0.5*up + tokens of the crop that owns the cell.  tokens: [n_img, 1+g*g, width].
__global__ void token_canvas_kernel(const float* __restrict__ tokens_all, int g, int width, int nh, int nw,
                                    float* __restrict__ canvas_all, int imgs_per_frame) {
  const int ph = nh * g, pw = nw * g;
  const int p = blockIdx.x;  // canvas cell; blockIdx.y = frame
  const float* tokens = tokens_all + static_cast<size_t>(blockIdx.y) * imgs_per_frame * (g * g + 1) * width;
  float* canvas = canvas_all + static_cast<size_t>(blockIdx.y) * ph * pw * width;
  const int py = p / pw, px = p - py * pw;
  const float sy = fmaxf((static_cast<float>(g) / ph) * (py + 0.5f) - 0.5f, 0.f);
  const float sx = fmaxf((static_cast<float>(g) / pw) * (px + 0.5f) - 0.5f, 0.f);
  const int y0 = min(static_cast<int>(sy), g - 1), x0 = min(static_cast<int>(sx), g - 1);
  const int y1 = min(y0 + 1, g - 1), x1 = min(x0 + 1, g - 1);
  const float ly = sy - y0, lx = sx - x0;
  const size_t tstride = static_cast<size_t>(g) * g + 1;
  const float* G = tokens + width;  // global image, skip cls row
  const int crop = 1 + (py / g) * nw + (px / g);
  const float* C = tokens + (crop * tstride + 1 + (py % g) * g + (px % g)) * width;
  for (int d = threadIdx.x; d < width; d += blockDim.x) {
    const float top = G[(static_cast<size_t>(y0) * g + x0) * width + d] * (1.f - lx) + G[(static_cast<size_t>(y0) * g + x1) * width + d] * lx;
    const float bot = G[(static_cast<size_t>(y1) * g + x0) * width + d] * (1.f - lx) + G[(static_cast<size_t>(y1) * g + x1) * width + d] * lx;
    const float up = top * (1.f - ly) + bot * ly;
    canvas[static_cast<size_t>(p) * width + d] = 0.5f * up + C[d];
  }
}

// ------------------------------------------------------------------------------------------- E4 feature masks
// get_features_mask (textregion.py:145-161) as used at :187 (`<= 0` is padded): a token belongs to the mask
// iff any bilinear tap with non-zero weight is set.
__global__ void feature_mask_kernel(const uint8_t* __restrict__ masks, int M, int H, int W, int ph, int pw,
                                    uint8_t* __restrict__ fmask, int* __restrict__ cnt) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = blockIdx.y;
  if (p >= ph * pw) return;
  const int py = p / pw, px = p - py * pw;
  const float sy = fmaxf((static_cast<float>(H) / ph) * (py + 0.5f) - 0.5f, 0.f);
  const float sx = fmaxf((static_cast<float>(W) / pw) * (px + 0.5f) - 0.5f, 0.f);
  const int y0 = min(static_cast<int>(sy), H - 1), x0 = min(static_cast<int>(sx), W - 1);
  const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
  const float ly = sy - y0, lx = sx - x0;
  const uint8_t* mk = masks + static_cast<size_t>(m) * H * W;
  bool on = false;
  if ((1.f - ly) > 0.f && (1.f - lx) > 0.f) on |= mk[static_cast<size_t>(y0) * W + x0] != 0;
  if ((1.f - ly) > 0.f && lx > 0.f) on |= mk[static_cast<size_t>(y0) * W + x1] != 0;
  if (ly > 0.f && (1.f - lx) > 0.f) on |= mk[static_cast<size_t>(y1) * W + x0] != 0;
  if (ly > 0.f && lx > 0.f) on |= mk[static_cast<size_t>(y1) * W + x1] != 0;
  fmask[static_cast<size_t>(m) * ph * pw + p] = on ? 1 : 0;
  if (on) atomicAdd(&cnt[m], 1);
}

// ------------------------------------------------------------------------------------------- E5 masked mean
// mean[m] = sum_{p in mask m} canvas[frame(m)][p] / cnt[m]  (uniform softmax of textregion.py:183-189).
// Work unit = (256 feature columns) x (group of <= 8 masks of one frame) x (64-token slice): the 8 masks share
// every canvas read, the loads of a thread are independent (unrolled); per-slice partial sums are written to
// acc[slice][mask][column] and added in slice order by mean_finalize_kernel (deterministic, no atomics).
// Empty masks (cnt 0) get a zero row, fixed up in l2_normalize_kernel.
constexpr int kMeanMasks = 8;
constexpr int kMeanSlice = 64;      // tokens staged per shared-memory refill
constexpr int kMeanMaxSplits = 16;  // token slices per mask (bounds the partial-sum buffer)
struct MaskGroup { int frame, m0, n, pad; };

__global__ void __launch_bounds__(256)
    masked_sum_kernel(const float* __restrict__ canvas_all, int P, int width, const uint8_t* __restrict__ fmask,
                      const MaskGroup* __restrict__ groups, float* __restrict__ partial, int M, int tokens_per_split) {
  const MaskGroup gr = groups[blockIdx.y];
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  const int pbeg = blockIdx.z * tokens_per_split, pend = min(P, pbeg + tokens_per_split);
  __shared__ uint8_t s_f[kMeanMasks][kMeanSlice];
  float acc[kMeanMasks];
#pragma unroll
  for (int i = 0; i < kMeanMasks; ++i) acc[i] = 0.f;
  for (int p0 = pbeg; p0 < pend; p0 += kMeanSlice) {
    const int pe = min(kMeanSlice, pend - p0);
    __syncthreads();
    for (int i = threadIdx.x; i < kMeanMasks * kMeanSlice; i += blockDim.x) {
      const int mi = i / kMeanSlice, pi = i % kMeanSlice;
      s_f[mi][pi] = (mi < gr.n && pi < pe) ? fmask[static_cast<size_t>(gr.m0 + mi) * P + p0 + pi] : 0;
    }
    __syncthreads();
    if (d < width) {
      const float* canvas = canvas_all + (static_cast<size_t>(gr.frame) * P + p0) * width + d;
#pragma unroll 8
      for (int pi = 0; pi < pe; ++pi) {
        const float c = canvas[static_cast<size_t>(pi) * width];
#pragma unroll
        for (int i = 0; i < kMeanMasks; ++i)
          if (s_f[i][pi]) acc[i] += c;
      }
    }
  }
  if (d < width) {
#pragma unroll
    for (int i = 0; i < kMeanMasks; ++i)
      if (i < gr.n) partial[(static_cast<size_t>(blockIdx.z) * M + gr.m0 + i) * width + d] = acc[i];
  }
}

__global__ void mean_finalize_kernel(const float* __restrict__ partial, int splits, const int* __restrict__ cnt, int M,
                                     int width, __nv_bfloat16* __restrict__ mean) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t n = static_cast<size_t>(M) * width;
  if (i >= n) return;
  float acc = 0.f;
  for (int z = 0; z < splits; ++z) acc += partial[z * n + i];
  const int c = cnt[i / width];
  mean[i] = __float2bfloat16_rn(c > 0 ? acc / static_cast<float>(c) : 0.f);
}

// F.normalize(dim=-1) (textregion.py:194): one warp per row.  Rows whose mask covered no token (all keys
// padded at textregion.py:187 -> the MHA attends to nothing) become normalize(out_proj.bias @ proj).
__global__ void l2_normalize_kernel(float* __restrict__ x, int rows, int dim, const int* __restrict__ cnt,
                                    const float* __restrict__ empty_vec) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float* r = x + static_cast<size_t>(row) * dim;
  if (cnt != nullptr && cnt[row] == 0)
    for (int d = lane; d < dim; d += 32) r[d] = empty_vec[d];
  __syncwarp();
  float ss = 0.f;
  for (int d = lane; d < dim; d += 32) ss += r[d] * r[d];
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
  for (int d = lane; d < dim; d += 32) r[d] *= inv;
}

// get_embed_txt_similarity's embedding rule (clip_generator.py:170-173,193-196): per query
// normalize(mean_over_templates(normalize(e))).  One warp per query; emb [Q*T, D] -> out [Q, D].
__global__ void text_bank_kernel(const float* __restrict__ emb, int Q, int T, int D, float* __restrict__ out) {
  const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (q >= Q) return;
  float* o = out + static_cast<size_t>(q) * D;
  for (int d = lane; d < D; d += 32) o[d] = 0.f;
  for (int t = 0; t < T; ++t) {
    const float* e = emb + (static_cast<size_t>(q) * T + t) * D;
    float ss = 0.f;
    for (int d = lane; d < D; d += 32) ss += e[d] * e[d];
    for (int k = 16; k > 0; k >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, k);
    const float nrm = sqrtf(ss);
    for (int d = lane; d < D; d += 32) o[d] += e[d] / nrm;
  }
  float ss = 0.f;
  for (int d = lane; d < D; d += 32) { o[d] /= static_cast<float>(T); ss += o[d] * o[d]; }
  for (int k = 16; k > 0; k >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, k);
  const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
  for (int d = lane; d < D; d += 32) o[d] *= inv;
}

// ------------------------------------------------------------------------------------------- text tower
// x[t*ctx+s] = token_embedding[tok] + positional_embedding[s] (pe.py:672-680); eot[t] = t*ctx + argmax_s tok
__global__ void text_embed_kernel(const int32_t* __restrict__ tokens, int T, int ctx, int width,
                                  const float* __restrict__ emb, const float* __restrict__ pos, float* __restrict__ x,
                                  int* __restrict__ eot_rows) {
  const int row = blockIdx.x;  // t*ctx + s
  const int t = row / ctx, s = row - t * ctx;
  const int tok = tokens[row];
  for (int d = threadIdx.x; d < width; d += blockDim.x)
    x[static_cast<size_t>(row) * width + d] = emb[static_cast<size_t>(tok) * width + d] + pos[static_cast<size_t>(s) * width + d];
  if (s == 0 && threadIdx.x == 0) {
    int best = tokens[row], bi = 0;
    for (int i = 1; i < ctx; ++i)
      if (tokens[row + i] > best) { best = tokens[row + i]; bi = i; }
    eot_rows[t] = t * ctx + bi;
  }
}

// =========================================================================================== crop-based descriptors
// CLIPGenerator.extract_clip, crop branch (ovo/entities/clip_generator.py:136-158): per mask a masked crop and a margin
// crop (ovo/utils/segment_utils.py:29-182), each through the full encode_image (pe.py:535-543), then fuse_clips
// (ovo/utils/clip_utils.py:21-48).

// batched_mask_to_box + batched_box_xyxy_to_xywh (segment_utils.py:43-104): one block per mask; edges are the min / max set
// row / column INDEX, so w = right - left, h = bottom - top (the reference's convention, kept); empty mask -> 0,0,0,0.
__global__ void __launch_bounds__(1024) mask_boxes_kernel(const uint8_t* __restrict__ masks, int H, int W, int32_t* __restrict__ xywh) {
  const uint8_t* m = masks + static_cast<size_t>(blockIdx.x) * H * W;
  int x0 = W, x1 = -1, y0 = H, y1 = -1;
  const int n = H * W;
  if ((W & 3) == 0 && (reinterpret_cast<uintptr_t>(m) & 3) == 0) {
    const uint32_t* m4 = reinterpret_cast<const uint32_t*>(m);
#pragma unroll 8
    for (int i = threadIdx.x; i < n / 4; i += blockDim.x) {
      const uint32_t v = __ldg(m4 + i);
      if (v == 0) continue;
      const int y = (4 * i) / W, x = 4 * i - y * W;
      const int lo = x + ((v & 0xffu) ? 0 : (v & 0xff00u) ? 1 : (v & 0xff0000u) ? 2 : 3);
      const int hi = x + ((v & 0xff000000u) ? 3 : (v & 0xff0000u) ? 2 : (v & 0xff00u) ? 1 : 0);
      x0 = min(x0, lo); x1 = max(x1, hi); y0 = min(y0, y); y1 = max(y1, y);
    }
  } else {
    for (int i = threadIdx.x; i < n; i += blockDim.x)
      if (m[i]) {
        const int y = i / W, x = i - y * W;
        x0 = min(x0, x); x1 = max(x1, x); y0 = min(y0, y); y1 = max(y1, y);
      }
  }
  __shared__ int s[4][32];
  for (int o = 16; o > 0; o >>= 1) {
    x0 = min(x0, __shfl_xor_sync(0xffffffffu, x0, o)); x1 = max(x1, __shfl_xor_sync(0xffffffffu, x1, o));
    y0 = min(y0, __shfl_xor_sync(0xffffffffu, y0, o)); y1 = max(y1, __shfl_xor_sync(0xffffffffu, y1, o));
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { s[0][warp] = x0; s[1][warp] = x1; s[2][warp] = y0; s[3][warp] = y1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < (blockDim.x >> 5); ++i) {
      x0 = min(x0, s[0][i]); x1 = max(x1, s[1][i]); y0 = min(y0, s[2][i]); y1 = max(y1, s[3][i]);
    }
    int32_t* o = xywh + 4 * blockIdx.x;
    if (x1 < x0) { o[0] = o[1] = o[2] = o[3] = 0; }
    else { o[0] = x0; o[1] = y0; o[2] = x1 - x0; o[3] = y1 - y0; }
  }
}

// One crop = a virtual source image of vh x vw pixels: inside [oy, oy+h) x [ox, ox+w) it shows the frame rectangle at (x, y)
// (times the mask when mask >= 0: get_seg_img, segment_utils.py:138-142), zero elsewhere (pad_img, :149-157).
struct CropJob { int mask, x, y, w, h, ox, oy, vw, vh; };

// torchvision F.resize weights of one crop (ATen _upsample_bilinear2d_aa), both axes: thread per output index
__global__ void crop_tables_kernel(const CropJob* __restrict__ jobs, int L, int kmax, int* __restrict__ tab_min,
                                   int* __restrict__ tab_size, float* __restrict__ tab_w) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= L) return;
  const CropJob j = jobs[blockIdx.z];
  const int axis = blockIdx.y;                              // 0 = x, 1 = y
  const int n_in = axis == 0 ? j.vw : j.vh;
  const double scale = static_cast<double>(n_in) / L;
  const double support = scale > 1.0 ? scale : 1.0, inv = 1.0 / support;
  const double center = scale * (o + 0.5);
  const int lo = max(0, static_cast<int>(center - support + 0.5));
  const int hi = min(n_in, static_cast<int>(center + support + 0.5));
  const size_t e = (static_cast<size_t>(blockIdx.z) * 2 + axis) * L + o;
  double total = 0.0;
  for (int t = lo; t < hi; ++t) total += fmax(0.0, 1.0 - fabs((t - center + 0.5) * inv));
  tab_min[e] = lo; tab_size[e] = min(hi - lo, kmax);
  float* w = tab_w + e * kmax;
  for (int t = lo; t < hi && t - lo < kmax; ++t) w[t - lo] = static_cast<float>(fmax(0.0, 1.0 - fabs((t - center + 0.5) * inv)) / total);
}

// F.resize(crop uint8, (L, L)) (segment_utils.py:131-135): f32 anti-aliased bilinear (horizontal taps first, like ATen's
// separable passes), torch.round, uint8.  Output HWC so that the second resize + im2col (E1's kernels) reads it like a frame.
__global__ void __launch_bounds__(128)
    crop_resize_kernel(const uint8_t* __restrict__ rgb, int H, int W, const uint8_t* __restrict__ masks, const CropJob* __restrict__ jobs,
                       int L, int kmax, const int* __restrict__ tab_min, const int* __restrict__ tab_size,
                       const float* __restrict__ tab_w, uint8_t* __restrict__ out) {
  const int ox = blockIdx.x * blockDim.x + threadIdx.x;
  const int oy = blockIdx.y;
  if (ox >= L) return;
  const CropJob j = jobs[blockIdx.z];
  const size_t ex = (static_cast<size_t>(blockIdx.z) * 2) * L + ox, ey = (static_cast<size_t>(blockIdx.z) * 2 + 1) * L + oy;
  const int xmin = tab_min[ex], nx = tab_size[ex], ymin = tab_min[ey], ny = tab_size[ey];
  const float* wx = tab_w + ex * kmax;
  const float* wy = tab_w + ey * kmax;
  const uint8_t* mk = j.mask >= 0 ? masks + static_cast<size_t>(j.mask) * H * W : nullptr;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  for (int ky = 0; ky < ny; ++ky) {
    const int vy = ymin + ky - j.oy;                          // row inside the crop rectangle
    float h0 = 0.f, h1 = 0.f, h2 = 0.f;
    if (vy >= 0 && vy < j.h) {
      const size_t rowbase = static_cast<size_t>(j.y + vy) * W + j.x;
      for (int kx = 0; kx < nx; ++kx) {
        const int vx = xmin + kx - j.ox;
        if (vx < 0 || vx >= j.w) continue;
        if (mk && !mk[rowbase + vx]) continue;
        const uint8_t* px = rgb + (rowbase + vx) * 3;
        const float wk = wx[kx];
        h0 += wk * static_cast<float>(px[0]);
        h1 += wk * static_cast<float>(px[1]);
        h2 += wk * static_cast<float>(px[2]);
      }
    }
    const float wk = wy[ky];
    a0 += wk * h0; a1 += wk * h1; a2 += wk * h2;
  }
  uint8_t* o = out + ((static_cast<size_t>(blockIdx.z) * L + oy) * L + ox) * 3;
  o[0] = static_cast<uint8_t>(min(max(__float2int_rn(a0), 0), 255));
  o[1] = static_cast<uint8_t>(min(max(__float2int_rn(a1), 0), 255));
  o[2] = static_cast<uint8_t>(min(max(__float2int_rn(a2), 0), 255));
}

__global__ void f32_to_bf16_kernel(const float* __restrict__ x, size_t n4, __nv_bfloat16* __restrict__ o) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 v = reinterpret_cast<const float4*>(x)[i];
  uint2 u;
  u.x = pack_bf16(v.x, v.y); u.y = pack_bf16(v.z, v.w);
  reinterpret_cast<uint2*>(o)[i] = u;
}

// AttentionPooling's attention (pe.py:81-84): ONE learned query per head against all tokens of an image.
// kv f32 [n*seq, 2*width] (k | v), q f32 [width] already scaled by head_dim^-0.5.  grid (n_img, heads), 256 threads.
__global__ void __launch_bounds__(256)
    pool_attention_kernel(const float* __restrict__ kv, const float* __restrict__ q, int seq, int width, int hd,
                          __nv_bfloat16* __restrict__ out) {
  extern __shared__ float sm[];            // [seq] scores, then [256] partial sums
  float* sc = sm;
  float* red = sm + seq;
  __shared__ float s_stat[2];
  const int img = blockIdx.x, head = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  const float* base = kv + static_cast<size_t>(img) * seq * 2 * width + head * hd;
  const float* qh = q + head * hd;
  // four keys per warp iteration: their loads are independent, so they overlap instead of paying one memory latency per key
  for (int j0 = warp * 4; j0 < seq; j0 += nwarp * 4) {
    float d[4] = {0.f, 0.f, 0.f, 0.f};
    for (int i = lane; i < hd; i += 32) {
      const float qv = qh[i];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (j0 + u < seq) d[u] += qv * __ldg(base + static_cast<size_t>(j0 + u) * 2 * width + i);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      for (int o = 16; o > 0; o >>= 1) d[u] += __shfl_xor_sync(0xffffffffu, d[u], o);
      if (lane == 0 && j0 + u < seq) sc[j0 + u] = d[u];
    }
  }
  __syncthreads();
  if (warp == 0) {
    float mx = -INFINITY;
    for (int j = lane; j < seq; j += 32) mx = fmaxf(mx, sc[j]);
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int j = lane; j < seq; j += 32) { const float p = expf(sc[j] - mx); sc[j] = p; sum += p; }
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) { s_stat[0] = mx; s_stat[1] = sum; }
  }
  __syncthreads();
  const float inv = 1.f / s_stat[1];
  const int slices = static_cast<int>(blockDim.x) / hd;   // hd <= blockDim.x (checked by the host)
  const int d = static_cast<int>(threadIdx.x) % hd, slice = static_cast<int>(threadIdx.x) / hd;
  float acc = 0.f;
  if (slice < slices) {
    const float* v = base + width + d;
    float a4[4] = {0.f, 0.f, 0.f, 0.f};
    int j = slice;
    for (; j + 7 * slices < seq; j += 8 * slices) {          // eight independent loads in flight
      float t[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) t[u] = __ldg(v + static_cast<size_t>(j + u * slices) * 2 * width);
#pragma unroll
      for (int u = 0; u < 8; ++u) a4[u & 3] += sc[j + u * slices] * t[u];
    }
    for (; j < seq; j += slices) a4[0] += sc[j] * __ldg(v + static_cast<size_t>(j) * 2 * width);
    acc = (a4[0] + a4[1]) + (a4[2] + a4[3]);
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  if (slice == 0) {
    for (int sidx = 1; sidx < slices; ++sidx) acc += red[sidx * hd + d];
    out[static_cast<size_t>(img) * width + head * hd + d] = __float2bfloat16_rn(acc * inv);
  }
}

// fuse_clips (clip_utils.py:21-48) for the crops of ONE frame: g [D] (the frame's global descriptor, repeated per mask at
// clip_generator.py:153), seg / bbox [M, D], all unit norm.  One block; warps stride over the masks; the soft-max of the
// `hovsg` / `concept_fusion` types runs ACROSS the masks of the frame (dim=0), so the cosines are staged in shared memory.
// embed_type: 1 fixed_weights, 2 hovsg, 3 adaptive_weights, 4 concept_fusion.
__device__ __forceinline__ float warp_sum(float v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__global__ void __launch_bounds__(1024)
    fuse_clips_kernel(const float* __restrict__ g, const float* __restrict__ seg, const float* __restrict__ bbox, int M, int D,
                      int embed_type, float w_masked, float w_global, float* __restrict__ out) {
  extern __shared__ float s_cos[];          // [M]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  const float eps = 1e-6f;                  // torch.nn.CosineSimilarity(eps=1e-6)
  float gg = 0.f;
  for (int d = lane; d < D; d += 32) gg += g[d] * g[d];
  const float gn = fmaxf(sqrtf(warp_sum(gg)), eps);
  for (int m = warp; m < M; m += nwarp) {
    const float* s = seg + static_cast<size_t>(m) * D;
    const float* b = bbox + static_cast<size_t>(m) * D;
    float* o = out + static_cast<size_t>(m) * D;
    float wl = w_masked;
    if (embed_type == 3) {
      float sb = 0.f, ss = 0.f, bb = 0.f;
      for (int d = lane; d < D; d += 32) { sb += s[d] * b[d]; ss += s[d] * s[d]; bb += b[d] * b[d]; }
      sb = warp_sum(sb); ss = warp_sum(ss); bb = warp_sum(bb);
      wl = sb / (fmaxf(sqrtf(ss), eps) * fmaxf(sqrtf(bb), eps)) * w_masked;
    }
    float ll = 0.f, gl = 0.f;
    for (int d = lane; d < D; d += 32) {
      const float l = embed_type == 4 ? b[d] : s[d] * wl + b[d] * (1.f - wl);
      o[d] = l;
      ll += l * l; gl += g[d] * l;
    }
    ll = warp_sum(ll); gl = warp_sum(gl);
    const float ln = sqrtf(ll);
    if (embed_type != 4) {                  // clip_l = normalize(...): F.normalize eps 1e-12
      const float inv = 1.f / fmaxf(ln, 1e-12f);
      for (int d = lane; d < D; d += 32) o[d] *= inv;
      gl *= inv; ll = ll * inv * inv;
    }
    if (lane == 0) s_cos[m] = gl / (gn * fmaxf(sqrtf(ll), eps));
  }
  __syncthreads();
  float mx = -INFINITY, den = 1.f;
  if (embed_type == 2 || embed_type == 4) {
    for (int m = lane; m < M; m += 32) mx = fmaxf(mx, s_cos[m]);
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int m = lane; m < M; m += 32) sum += expf(s_cos[m] - mx);
    den = warp_sum(sum);
  }
  for (int m = warp; m < M; m += nwarp) {
    float* o = out + static_cast<size_t>(m) * D;
    float wg = w_global;
    if (embed_type == 2 || embed_type == 4) wg = expf(s_cos[m] - mx) / den;
    else if (embed_type == 3) wg = s_cos[m] * w_global;
    float ss = 0.f;
    for (int d = lane; d < D; d += 32) {
      const float v = g[d] * wg + o[d] * (1.f - wg);
      o[d] = v; ss += v * v;
    }
    const float inv = 1.f / fmaxf(sqrtf(warp_sum(ss)), 1e-12f);
    for (int d = lane; d < D; d += 32) o[d] *= inv;
  }
}

// return_all (clip_generator.py:151-152): out [M,3,D] = (global, masked crop, margin crop)
__global__ void stack_clips_kernel(const float* __restrict__ g, const float* __restrict__ seg, const float* __restrict__ bbox,
                                   int M, int D, float* __restrict__ out) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(M) * 3 * D) return;
  const int d = i % D, k = (i / D) % 3;
  const size_t m = i / (static_cast<size_t>(3) * D);
  out[i] = k == 0 ? g[d] : k == 1 ? seg[m * D + d] : bbox[m * D + d];
}

// siglip_cosine_similarity (clip_utils.py:10-14): sim <- sigmoid(sim * exp(logit_scale) + logit_bias), in place
__global__ void sigmoid_affine_kernel(float* __restrict__ sim, size_t n, float scale, float bias) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) sim[i] = 1.f / (1.f + expf(-(sim[i] * scale + bias)));
}

// ------------------------------------------------------------------------------------------- embed_type `learned`
// WeightsPredictorMerger (ovo/entities/clips_merging.py:26-56): a small post-norm transformer over the THREE descriptors of a mask
// (global, masked crop, margin crop), an MLP on the flattened tokens, soft-max weights over the three, weighted sum, L2 norm.
// Self-attention over 3 tokens: one thread per (mask, head).  qkv f32 [3B, 3d] (q | k | v) -> out bf16 [3B, d].
__global__ void merger_attention_kernel(const float* __restrict__ qkv, int B, int d, int nhead, __nv_bfloat16* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * nhead) return;
  const int b = i / nhead, h = i - b * nhead, hd = d / nhead;
  const float* base = qkv + static_cast<size_t>(b) * 3 * 3 * d + h * hd;
  float sc[3][3];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int c = 0; c < 3; ++c) sc[a][c] = 0.f;
  for (int e = 0; e < hd; ++e) {
    float q[3], k[3];
#pragma unroll
    for (int t = 0; t < 3; ++t) { q[t] = base[static_cast<size_t>(t) * 3 * d + e]; k[t] = base[static_cast<size_t>(t) * 3 * d + d + e]; }
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int c = 0; c < 3; ++c) sc[a][c] += q[a] * k[c];
  }
  const float scale = rsqrtf(static_cast<float>(hd));
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float m = fmaxf(fmaxf(sc[a][0], sc[a][1]), sc[a][2]) * scale;
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) { sc[a][c] = expf(sc[a][c] * scale - m); sum += sc[a][c]; }
#pragma unroll
    for (int c = 0; c < 3; ++c) sc[a][c] /= sum;
  }
  for (int e = 0; e < hd; ++e) {
    float v[3];
#pragma unroll
    for (int t = 0; t < 3; ++t) v[t] = base[static_cast<size_t>(t) * 3 * d + 2 * d + e];
#pragma unroll
    for (int a = 0; a < 3; ++a)
      out[(static_cast<size_t>(b) * 3 + a) * d + h * hd + e] = __float2bfloat16_rn(sc[a][0] * v[0] + sc[a][1] * v[1] + sc[a][2] * v[2]);
  }
}

// nn.LeakyReLU (slope 0.01, clips_merging.py:6-11) on an f32 GEMM output -> bf16 operand of the next linear
__global__ void leaky_relu_bf16_kernel(const float* __restrict__ x, size_t n, __nv_bfloat16* __restrict__ o) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) { const float v = x[i]; o[i] = __float2bfloat16_rn(v > 0.f ? v : 0.01f * v); }
}

// clips_merging.py:48-55: soft-max over the three clips (per channel when the MLP emits 3*D weights, per clip when it emits 3),
// weighted sum, F.normalize.  One warp per mask.
__global__ void merger_combine_kernel(const float* __restrict__ clips, const float* __restrict__ wts, int B, int D, int o_dim,
                                      float* __restrict__ out) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= B) return;
  const float* c = clips + static_cast<size_t>(b) * 3 * D;
  const float* w = wts + static_cast<size_t>(b) * o_dim;
  float* o = out + static_cast<size_t>(b) * D;
  float ss = 0.f;
  for (int d = lane; d < D; d += 32) {
    float l[3];
#pragma unroll
    for (int t = 0; t < 3; ++t) l[t] = o_dim == 3 ? w[t] : w[t * D + d];
    const float m = fmaxf(fmaxf(l[0], l[1]), l[2]);
    const float e0 = expf(l[0] - m), e1 = expf(l[1] - m), e2 = expf(l[2] - m);
    const float v = (c[d] * e0 + c[D + d] * e1 + c[2 * D + d] * e2) / (e0 + e1 + e2);
    o[d] = v; ss += v * v;
  }
  ss = warp_sum(ss);
  const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
  for (int d = lane; d < D; d += 32) o[d] *= inv;
}

}  // namespace ovo

// =============================================================================================== handle
using namespace ovo;

struct AaTable {
  int offset;  // entry offset into the device tables
  int k;       // taps per row
};

struct ovo_encoder {
  ovo_vit_cfg cfg{};
  ovo_vit_weights w{};
  std::vector<ovo_block_weights> blocks, text_blocks;
  int max_images = 0, max_h = 0, max_w = 0, max_masks = 0;
  int grid = 0, patches = 0, seq = 0, seq_pad = 0;
  size_t rows_cap = 0;  // activation rows
  // activations
  __nv_bfloat16 *patch_buf = nullptr, *xn = nullptr, *q = nullptr, *k = nullptr, *vt = nullptr, *attn = nullptr,
                *hmid = nullptr, *mean = nullptr, *pooled_in = nullptr;
  float2* rope_tab = nullptr;
  int hd = 64;                    // vision head_dim; != 64 takes the generic path (generic_attention.cuh)
  float* qkv_f32 = nullptr;       // [rows, 3*width] f32: QKV GEMM output of the generic path
  float2* rope_gen = nullptr;     // [grid+1][hd/4] rotary table of the generic path
  float *x = nullptr, *xfinal = nullptr, *canvas = nullptr,
        *resize_tmp = nullptr;
  uint8_t* fmask = nullptr;
  float* mean_acc = nullptr;
  MaskGroup* groups_dev = nullptr;
  int *fcnt = nullptr, *eot_rows = nullptr;
  // AA resize tables
  int *tab_min = nullptr, *tab_size = nullptr;
  float* tab_w = nullptr;
  int tab_entries = 0, tab_cap = 0, tab_wcap = 0, tab_wused = 0;
  std::map<std::pair<int, int>, AaTable> tables;
  ImgJob* jobs_dev = nullptr;
  int jobs_cap = 0;
  std::vector<ImgJob> jobs_host;   // last uploaded job list (re-uploaded only when the frame geometry changes)
  int loaded_images = 0;
  // CUDA graphs of the transformer stack, keyed by (n_img, n_layers, ln_post)
  struct GraphEntry { cudaGraphExec_t exec = nullptr; long long launches = 0; int warm = 0; };
  std::map<long long, GraphEntry> graphs;
  bool use_graphs = true;
  cudaStream_t cap_stream = nullptr;  // capture happens here (the caller's stream may be the legacy stream)
  // crop-based descriptors (ovo_encoder_set_pool_head / ovo_encode_crops)
  bool has_head = false;
  ovo_pool_head_weights head{};
  float *ph_x1 = nullptr, *ph_x2 = nullptr, *ph_emb = nullptr;
  __nv_bfloat16 *ph_a = nullptr, *ph_h = nullptr;
  int32_t* crop_boxes = nullptr;
  CropJob* crop_jobs = nullptr;
  std::vector<CropJob> crop_jobs_host;
  uint8_t* crop_u8 = nullptr;
  size_t crop_u8_cap = 0;
  int *ctab_min = nullptr, *ctab_size = nullptr;
  float* ctab_w = nullptr;
  size_t ctab_cap = 0, ctab_wcap = 0;
};

long long* g_attn_trace = nullptr;   // ovo_attn_trace: device buffer [8 items][16 blocks][8 phases] of clock64 stamps, or null
int g_attn_debug = 0;   // tuning experiments (attention.cuh dbg bits), set through ovo_set_gemm_cluster bits 24..31

namespace {

template <typename T>
int dmalloc(T** p, size_t n) {
  if (cudaMalloc(reinterpret_cast<void**>(p), n * sizeof(T)) != cudaSuccess) {
    cudaGetLastError();
    return set_error(OVO_E_NOMEM, "encoder workspace allocation of %zu bytes failed", n * sizeof(T));
  }
  if (cudaMemset(*p, 0, n * sizeof(T)) != cudaSuccess) return set_error(OVO_E_CUDA, "memset failed");
  return OVO_OK;
}

int launch_ln(const float* x, int rows, int width, const float* g, const float* b, float eps, __nv_bfloat16* o16,
              float* o32, const int* gather, cudaStream_t s) {
  OVO_REQUIRE(width % 4 == 0 && width <= 32 * 4 * 16, "layernorm: unsupported width %d", width);
  const int blocks = ceil_div(rows, 8);
  ProfScope prof(s, PROF_LN, 0.0, static_cast<double>(rows) * width * (4.0 + (o16 ? 2.0 : 0.0) + (o32 ? 4.0 : 0.0)));
  if (width <= 1024)
    layernorm_kernel<8><<<blocks, 256, 0, s>>>(x, rows, width, g, b, eps, o16, o32, gather);
  else
    layernorm_kernel<16><<<blocks, 256, 0, s>>>(x, rows, width, g, b, eps, o16, o32, gather);
  OVO_CHECK_LAUNCH();
  return OVO_OK;
}

int launch_attention(ovo_encoder* e, int n_seq, int seq, int seq_pad, int heads, int width, bool causal, cudaStream_t s) {
  OVO_REQUIRE(seq_pad % 128 == 0 && seq_pad / 128 <= kAttnMaxBlocks, "attention: seq_pad %d unsupported", seq_pad);
  CUtensorMap tq, tk, tv;
  const uint64_t bh = static_cast<uint64_t>(n_seq) * heads;
  OVO_TRY(make_tmap_bf16_2d(&tq, e->q, bh * seq_pad, 64, 64, 128, 64));
  OVO_TRY(make_tmap_bf16_2d(&tk, e->k, bh * seq_pad, 64, 64, kAttnKB, 64));
  OVO_TRY(make_tmap_bf16_2d(&tv, e->vt, bh * seq_pad, 64, 64, kAttnKB, 64));   // V [b,h,seq_pad,64]: MN-major B operand of P.V
  const int smem = AttnSmem::kBytes;
  static bool attr_set = false;
  if (!attr_set) {
    OVO_CUDA(cudaFuncSetAttribute(attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  const int qtiles = ceil_div(seq, 128);
  const float scale_log2e = 0.125f * 1.4426950408889634f;  // head_dim 64
  ProfScope prof(s, PROF_ATTN, 4.0 * static_cast<double>(bh) * seq * (causal ? 0.5 * seq : seq) * 64, 4.0 * static_cast<double>(bh) * seq * 64 * 2);
  const int n_items = qtiles * static_cast<int>(bh);
  // one key left over after the 64-key blocks (577 = 9 * 64 + 1): folded into the epilogue instead of a tenth block
  static const bool tail_off = getenv("OVO_B200_ATTN_TAIL") && getenv("OVO_B200_ATTN_TAIL")[0] == '0';   // A/B measurement aid
  const int tail1 = (!causal && seq > 64 && seq % kAttnKB == 1 && !tail_off) ? 1 : 0;
  attention_fwd_kernel<<<std::min(n_items, 2 * num_sms()), kAttnThreads, smem, s>>>(
      tq, tk, tv, e->attn, seq, seq_pad, heads, width, scale_log2e, causal ? 1 : 0, g_attn_debug, qtiles, n_items, e->k, e->vt, tail1, g_attn_trace);
  OVO_CHECK_LAUNCH();
  return OVO_OK;
}

// pe.py:216-225 for `n_layers` blocks over rows = n_seq * seq
int run_blocks(ovo_encoder* e, const std::vector<ovo_block_weights>& blocks, int n_layers, int n_seq, int seq,
               int seq_pad, int heads, int width, int mlp, bool rope, bool causal, cudaStream_t s) {
  const int rows = n_seq * seq;
  for (int l = 0; l < n_layers; ++l) {
    const ovo_block_weights& b = blocks[l];
    OVO_TRY(launch_ln(e->x, rows, width, b.ln1_w, b.ln1_b, e->cfg.ln_eps, e->xn, nullptr, nullptr, s));
    if (rope && e->hd != 64) {
      // generic head_dim: f32 q|k|v straight from the GEMM, rotary embedding + bf16 rounding while the attention kernel stages
      EpiParams qe;
      qe.out = e->qkv_f32; qe.ldo = 3 * width; qe.bias = b.qkv_b;
      OVO_TRY(launch_gemm(EPI_F32, e->xn, width, static_cast<const __nv_bfloat16*>(b.qkv_w), width, rows, 3 * width, width, qe, s));
      GenAttnParams gp;
      gp.qkv = e->qkv_f32; gp.out = e->attn; gp.seq = seq; gp.heads = heads; gp.width = width; gp.causal = causal ? 1 : 0;
      gp.rope = e->rope_gen; gp.rope_grid = e->grid;
      gp.scale_log2e = 1.4426950408889634f / sqrtf(static_cast<float>(e->hd));
      ProfScope prof(s, PROF_ATTN, 4.0 * n_seq * heads * static_cast<double>(seq) * seq * e->hd, 0.0);
      if (e->hd == 80) generic_attention_kernel<80><<<dim3(ceil_div(seq, 64), n_seq, heads), 128, 0, s>>>(gp);
      else if (e->hd == 96) generic_attention_kernel<96><<<dim3(ceil_div(seq, 64), n_seq, heads), 128, 0, s>>>(gp);
      else if (e->hd == 32) generic_attention_kernel<32><<<dim3(ceil_div(seq, 64), n_seq, heads), 128, 0, s>>>(gp);
      else return set_error(OVO_E_INVALID, "head_dim %d has no attention kernel (64 fast path; 32, 80, 96 generic)", e->hd);
      OVO_CHECK_LAUNCH();
    } else {
    EpiParams qkv;
    qkv.bias = b.qkv_b; qkv.q = e->q; qkv.k = e->k; qkv.vt = e->vt;
    qkv.rope_tab = rope ? e->rope_tab : nullptr; qkv.rope_grid = e->grid;
    qkv.seq = seq; qkv.seq_pad = seq_pad; qkv.heads = heads; qkv.width = width;
    OVO_TRY(launch_gemm(EPI_QKV, e->xn, width, static_cast<const __nv_bfloat16*>(b.qkv_w), width, rows, 3 * width, width, qkv, s));
    OVO_TRY(launch_attention(e, n_seq, seq, seq_pad, heads, width, causal, s));
    }
    EpiParams op;
    op.out = e->x; op.ldo = width; op.bias = b.out_b; op.resid = e->x; op.ldr = width;
    OVO_TRY(launch_gemm(EPI_F32_RESID, e->attn, width, static_cast<const __nv_bfloat16*>(b.out_w), width, rows, width, width, op, s));
    OVO_TRY(launch_ln(e->x, rows, width, b.ln2_w, b.ln2_b, e->cfg.ln_eps, e->xn, nullptr, nullptr, s));
    EpiParams fc;
    fc.out = e->hmid; fc.ldo = mlp; fc.bias = b.fc_b;
    OVO_TRY(launch_gemm(EPI_BF16_GELU, e->xn, width, static_cast<const __nv_bfloat16*>(b.fc_w), width, rows, mlp, width, fc, s));
    EpiParams pj;
    pj.out = e->x; pj.ldo = width; pj.bias = b.proj_b; pj.resid = e->x; pj.ldr = width;
    OVO_TRY(launch_gemm(EPI_F32_RESID, e->hmid, mlp, static_cast<const __nv_bfloat16*>(b.proj_w), mlp, rows, width, mlp, pj, s));
  }
  return OVO_OK;
}

// ATen _upsample_bilinear2d_aa weights for one axis: returns taps per output
int aa_axis(int n_in, int n_out, std::vector<int>& mins, std::vector<int>& sizes, std::vector<float>& ws) {
  const double scale = static_cast<double>(n_in) / n_out;
  const double support = std::max(scale, 1.0), inv = 1.0 / std::max(scale, 1.0);
  const int kmax = static_cast<int>(std::ceil(support)) * 2 + 1;
  mins.resize(n_out); sizes.resize(n_out); ws.assign(static_cast<size_t>(n_out) * kmax, 0.f);
  for (int i = 0; i < n_out; ++i) {
    const double center = scale * (i + 0.5);
    const int lo = std::max(0, static_cast<int>(center - support + 0.5));
    const int hi = std::min(n_in, static_cast<int>(center + support + 0.5));
    double total = 0;
    std::vector<double> t(hi - lo);
    for (int j = lo; j < hi; ++j) {
      t[j - lo] = std::max(0.0, 1.0 - std::fabs((j - center + 0.5) * inv));
      total += t[j - lo];
    }
    mins[i] = lo; sizes[i] = hi - lo;
    for (int j = 0; j < hi - lo; ++j) ws[static_cast<size_t>(i) * kmax + j] = static_cast<float>(t[j] / total);
  }
  return kmax;
}

int get_table(ovo_encoder* e, int n_in, AaTable* out, cudaStream_t s) {
  const int S = e->cfg.image_size;
  auto it = e->tables.find({n_in, S});
  if (it != e->tables.end()) { *out = it->second; return OVO_OK; }
  std::vector<int> mins, sizes; std::vector<float> ws;
  const int k = aa_axis(n_in, S, mins, sizes, ws);
  OVO_REQUIRE(e->tab_entries + S <= e->tab_cap && e->tab_wused + S * k <= e->tab_wcap, "resize table space exhausted");
  // the weight table of this axis starts at tab_wused; kernels index it as (offset_entries + o) * k, so keep
  // the entry offset consistent by giving each table its own weight base = entry_offset * k.
  const int entry_off = ceil_div(e->tab_wused, k) > e->tab_entries ? ceil_div(e->tab_wused, k) : e->tab_entries;
  OVO_REQUIRE(entry_off + S <= e->tab_cap && static_cast<size_t>(entry_off + S) * k <= static_cast<size_t>(e->tab_wcap), "resize table space exhausted");
  OVO_CUDA(cudaMemcpyAsync(e->tab_min + entry_off, mins.data(), S * sizeof(int), cudaMemcpyHostToDevice, s));
  OVO_CUDA(cudaMemcpyAsync(e->tab_size + entry_off, sizes.data(), S * sizeof(int), cudaMemcpyHostToDevice, s));
  OVO_CUDA(cudaMemcpyAsync(e->tab_w + static_cast<size_t>(entry_off) * k, ws.data(), static_cast<size_t>(S) * k * sizeof(float), cudaMemcpyHostToDevice, s));
  OVO_CUDA(cudaStreamSynchronize(s));  // host vectors go out of scope; tables are built once per size
  e->tab_entries = entry_off + S;
  e->tab_wused = (entry_off + S) * k;
  AaTable t{entry_off, k};
  e->tables[{n_in, S}] = t;
  *out = t;
  return OVO_OK;
}

}  // namespace

extern "C" {

void ovo_attn_trace(long long* buf_dev) { g_attn_trace = buf_dev; }


int ovo_encoder_create(const ovo_vit_cfg* cfg, const ovo_vit_weights* w, int max_images, int max_h, int max_w,
                       int max_masks, ovo_encoder_t** out) {
  OVO_REQUIRE(cfg && w && out, "ovo_encoder_create: null argument");
  OVO_REQUIRE(cfg->heads > 0 && cfg->width % cfg->heads == 0, "width %d is not a multiple of heads %d", cfg->width, cfg->heads);
  {
    const int hd = cfg->width / cfg->heads;
    OVO_REQUIRE(cfg->width % 8 == 0 && (hd == 64 || hd == 32 || hd == 80 || hd == 96),
                "vision head_dim %d unsupported (64 = tcgen05 path; 32, 80, 96 = generic path)", hd);
  }
  OVO_REQUIRE(cfg->image_size % cfg->patch_size == 0 && cfg->image_size / cfg->patch_size < kRopeRowsMax, "image_size must be a multiple of patch_size (grid < 40)");
  OVO_REQUIRE(max_images > 0 && max_masks > 0 && max_h > 0 && max_w > 0, "ovo_encoder_create: bad limits");
  OVO_REQUIRE(w->patch_kpad % 64 == 0 && w->patch_kpad >= 3 * cfg->patch_size * cfg->patch_size, "patch_kpad must be a multiple of 64");
  if (cfg->text_layers > 0)
    OVO_REQUIRE(cfg->text_heads > 0 && cfg->text_width / cfg->text_heads == 64 && cfg->text_ctx <= 128 && w->text_blocks,
                "text tower: head_dim must be 64 and ctx <= 128");
  OVO_REQUIRE(ovo::ceil_div(cfg->image_size / cfg->patch_size * (cfg->image_size / cfg->patch_size) + 1, 128) <= kAttnMaxBlocks,
              "sequence too long for the attention kernel");
  ovo_encoder* e = new ovo_encoder();
  e->cfg = *cfg; e->w = *w;
  e->blocks.assign(w->blocks, w->blocks + cfg->layers);
  if (cfg->text_layers > 0) e->text_blocks.assign(w->text_blocks, w->text_blocks + cfg->text_layers);
  e->max_images = max_images; e->max_h = max_h; e->max_w = max_w; e->max_masks = max_masks;
  e->hd = cfg->width / cfg->heads;
  e->grid = cfg->image_size / cfg->patch_size; e->patches = e->grid * e->grid; e->seq = e->patches + 1;
  e->seq_pad = ceil_div(e->seq, 128) * 128;
  const char* env = getenv("OVO_B200_GRAPHS");
  e->use_graphs = !(env && env[0] == '0');

  const int W = std::max(cfg->width, cfg->text_width), F = std::max(cfg->mlp_width, cfg->text_mlp_width);
  const int heads = std::max(cfg->heads, cfg->text_heads);
  constexpr int kMaxText = 256;                                     // strings per ovo_encode_text call
  const size_t rows = std::max(static_cast<size_t>(max_images) * e->seq, static_cast<size_t>(kMaxText) * std::max(cfg->text_ctx, 1));
  e->rows_cap = rows;
  const size_t rows_pad = rows + 128;
  const size_t qk_elems = std::max(static_cast<size_t>(max_images) * e->seq_pad, static_cast<size_t>(kMaxText) * 128) * heads * 64;
  const int nh_max = std::max(max_h / cfg->image_size, 1), nw_max = std::max(max_w / cfg->image_size, 1);
  const size_t pmax = static_cast<size_t>(nh_max) * nw_max * e->patches;
  int r = OVO_OK;
  r |= dmalloc(&e->patch_buf, static_cast<size_t>(max_images) * e->patches * w->patch_kpad + 64);
  r |= dmalloc(&e->x, rows_pad * W);
  r |= dmalloc(&e->xfinal, rows_pad * W);
  r |= dmalloc(&e->xn, rows_pad * W);
  r |= dmalloc(&e->attn, rows_pad * W);
  r |= dmalloc(&e->hmid, rows_pad * F);
  r |= dmalloc(&e->q, qk_elems);
  r |= dmalloc(&e->k, qk_elems);
  r |= dmalloc(&e->vt, qk_elems);
  r |= dmalloc(&e->rope_tab, static_cast<size_t>(e->grid + 1) * 16);
  if (e->hd != 64) {
    r |= dmalloc(&e->qkv_f32, static_cast<size_t>(max_images) * e->seq * 3 * cfg->width + 64);
    r |= dmalloc(&e->rope_gen, static_cast<size_t>(e->grid + 1) * (e->hd / 4));
  }
  r |= dmalloc(&e->canvas, std::max(pmax, static_cast<size_t>(max_images) * e->patches) * cfg->width);
  r |= dmalloc(&e->fmask, static_cast<size_t>(max_masks) * pmax);
  r |= dmalloc(&e->fcnt, static_cast<size_t>(max_masks));
  r |= dmalloc(&e->mean, static_cast<size_t>(max_masks + 128) * cfg->width);
  r |= dmalloc(&e->mean_acc, static_cast<size_t>(kMeanMaxSplits) * max_masks * cfg->width);
  r |= dmalloc(&e->groups_dev, static_cast<size_t>(max_masks) + static_cast<size_t>(max_images));
  r |= dmalloc(&e->resize_tmp, static_cast<size_t>(max_images) * 3 * max_h * cfg->image_size);
  r |= dmalloc(&e->eot_rows, rows_pad);
  r |= dmalloc(&e->pooled_in, rows_pad * static_cast<size_t>(W) / 8 + static_cast<size_t>(W) * 128);
  e->tab_cap = 16 * cfg->image_size; e->tab_wcap = e->tab_cap * 64;
  r |= dmalloc(&e->tab_min, static_cast<size_t>(e->tab_cap));
  r |= dmalloc(&e->tab_size, static_cast<size_t>(e->tab_cap));
  r |= dmalloc(&e->tab_w, static_cast<size_t>(e->tab_wcap));
  e->jobs_cap = max_images + 1;   // + the frame's global image of the crop-based path
  r |= dmalloc(&e->jobs_dev, static_cast<size_t>(max_images) + 1);
  if (r != OVO_OK) { ovo_encoder_destroy(e); return OVO_E_NOMEM; }

  // 2D RoPE table (rope.py:315-340): both axes share theta_i = 10000^(-2i/32); row r holds
  // (cos, sin)(r * theta_i) for r = coordinate + 1 (r = 0 is the cls token's zero angle).
  {
    std::vector<float2> tab(static_cast<size_t>(e->grid + 1) * 16);
    for (int r = 0; r <= e->grid; ++r)
      for (int i = 0; i < 16; ++i) {
        const float theta = 1.0f / powf(10000.0f, static_cast<float>(2 * i) / 32.0f);
        const float ang = static_cast<float>(r) * theta;
        tab[static_cast<size_t>(r) * 16 + i] = make_float2(cosf(ang), sinf(ang));
      }
    if (cudaMemcpy(e->rope_tab, tab.data(), tab.size() * sizeof(float2), cudaMemcpyHostToDevice) != cudaSuccess) {
      ovo_encoder_destroy(e);
      return set_error(OVO_E_CUDA, "rope table upload failed");
    }
  }
  if (e->hd != 64) {   // rope.py:315-340 for a general head_dim: hd/4 frequencies per axis, theta_i = 10000^(-2i / (hd/2))
    const int qn = e->hd / 4;
    std::vector<float2> tab(static_cast<size_t>(e->grid + 1) * qn);
    for (int r = 0; r <= e->grid; ++r)
      for (int i = 0; i < qn; ++i) {
        const float theta = 1.0f / powf(10000.0f, static_cast<float>(2 * i) / static_cast<float>(e->hd / 2));
        const float ang = static_cast<float>(r) * theta;
        tab[static_cast<size_t>(r) * qn + i] = make_float2(cosf(ang), sinf(ang));
      }
    if (cudaMemcpy(e->rope_gen, tab.data(), tab.size() * sizeof(float2), cudaMemcpyHostToDevice) != cudaSuccess) {
      ovo_encoder_destroy(e);
      return set_error(OVO_E_CUDA, "rope table upload failed");
    }
  }
  *out = e;
  return OVO_OK;
}

void ovo_encoder_destroy(ovo_encoder_t* e) {
  if (!e) return;
  for (auto& g : e->graphs) if (g.second.exec) cudaGraphExecDestroy(g.second.exec);
  if (e->cap_stream) cudaStreamDestroy(e->cap_stream);
  cudaFree(e->patch_buf); cudaFree(e->x); cudaFree(e->xfinal); cudaFree(e->xn); cudaFree(e->attn); cudaFree(e->hmid);
  cudaFree(e->q); cudaFree(e->k); cudaFree(e->vt); cudaFree(e->rope_tab); cudaFree(e->canvas); cudaFree(e->qkv_f32); cudaFree(e->rope_gen);
  cudaFree(e->fmask); cudaFree(e->mean_acc); cudaFree(e->groups_dev); cudaFree(e->fcnt); cudaFree(e->mean); cudaFree(e->resize_tmp); cudaFree(e->eot_rows);
  cudaFree(e->pooled_in); cudaFree(e->tab_min); cudaFree(e->tab_size); cudaFree(e->tab_w); cudaFree(e->jobs_dev);
  cudaFree(e->ph_x1); cudaFree(e->ph_x2); cudaFree(e->ph_emb); cudaFree(e->ph_a); cudaFree(e->ph_h); cudaFree(e->crop_boxes);
  cudaFree(e->crop_jobs); cudaFree(e->crop_u8); cudaFree(e->ctab_min); cudaFree(e->ctab_size); cudaFree(e->ctab_w);
  delete e;
}

int ovo_encoder_preprocess(ovo_encoder_t* e, const uint8_t* rgb_dev, int n_frames, int H, int W, int* n_img_per_frame,
                           void* stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  OVO_REQUIRE(e && rgb_dev && n_frames > 0, "ovo_encoder_preprocess: bad arguments");
  OVO_REQUIRE(H > 0 && W > 0 && H <= e->max_h && W <= e->max_w, "frame %dx%d exceeds encoder limits %dx%d", H, W, e->max_h, e->max_w);
  const int S = e->cfg.image_size;
  // crop boxes of the multi_resolution strategy (textregion.py:114-128)
  const int nh = std::max(H / S, 1), nw = std::max(W / S, 1);
  const int ch = ceil_div(H, nh), cw = ceil_div(W, nw);
  const int per = 1 + nh * nw;
  OVO_REQUIRE(n_frames * per <= e->max_images, "%d frames x %d images exceed max_images %d", n_frames, per, e->max_images);
  AaTable tgx, tgy, tcx, tcy;
  OVO_TRY(get_table(e, W, &tgx, s)); OVO_TRY(get_table(e, H, &tgy, s));
  OVO_TRY(get_table(e, cw, &tcx, s)); OVO_TRY(get_table(e, ch, &tcy, s));
  std::vector<ImgJob> jobs;
  for (int f = 0; f < n_frames; ++f) {
    jobs.push_back({f, 0, 0, H, W, tgx.offset, tgy.offset, tgx.k, tgy.k});
    for (int hi = 0; hi < nh; ++hi)
      for (int wi = 0; wi < nw; ++wi) {
        int y1 = hi * ch, x1 = wi * cw;
        const int y2 = std::min(y1 + ch, H), x2 = std::min(x1 + cw, W);
        y1 = std::max(y2 - ch, 0); x1 = std::max(x2 - cw, 0);
        jobs.push_back({f, y1, x1, y2 - y1, x2 - x1, tcx.offset, tcy.offset, tcx.k, tcy.k});
      }
  }
  const int n_img = static_cast<int>(jobs.size());
  if (jobs.size() != e->jobs_host.size() || memcmp(jobs.data(), e->jobs_host.data(), jobs.size() * sizeof(ImgJob)) != 0) {
    e->jobs_host = jobs;  // persistent host copy: the async upload may outlive this call
    OVO_CUDA(cudaMemcpyAsync(e->jobs_dev, e->jobs_host.data(), jobs.size() * sizeof(ImgJob), cudaMemcpyHostToDevice, s));
    OVO_CUDA(cudaStreamSynchronize(s));
  }
  ProfScope prof(s, PROF_PRE, 0.0, static_cast<double>(n_frames) * H * W * 3 + static_cast<double>(n_img) * e->patches * e->w.patch_kpad * 2);
  aa_resize_h_kernel<<<dim3(ceil_div(S, 128), H, n_img), 128, 0, s>>>(rgb_dev, H, W, e->jobs_dev, e->tab_min, e->tab_size, e->tab_w, S, e->resize_tmp, e->max_h, 0);
  OVO_CHECK_LAUNCH();
  aa_resize_v_patch_kernel<<<dim3(ceil_div(S, 128), S, n_img), 128, 0, s>>>(e->resize_tmp, e->max_h, e->jobs_dev, e->tab_min, e->tab_size, e->tab_w, S, e->cfg.patch_size, e->w.patch_kpad, e->patch_buf, 0);
  OVO_CHECK_LAUNCH();
  e->loaded_images = n_img;
  if (n_img_per_frame) *n_img_per_frame = per;
  return OVO_OK;
}

int ovo_encoder_load_pixels(ovo_encoder_t* e, const float* pixels_dev, int n_img, void* stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  OVO_REQUIRE(e && pixels_dev && n_img > 0 && n_img <= e->max_images, "ovo_encoder_load_pixels: bad arguments");
  const int S = e->cfg.image_size;
  const size_t total = static_cast<size_t>(n_img) * 3 * S * S;
  pixels_to_patches_kernel<<<ceil_div(total, 256), 256, 0, s>>>(pixels_dev, n_img, S, e->cfg.patch_size, e->w.patch_kpad, e->patch_buf);
  OVO_CHECK_LAUNCH();
  e->loaded_images = n_img;
  return OVO_OK;
}

static int forward_eager(ovo_encoder* e, int n_img, int n_layers, int apply_ln_post, cudaStream_t s) {
  const ovo_vit_cfg& c = e->cfg;
  const int rows = n_img * e->seq;
  // patch embed (pe.py:509-519) + cls row, then ln_pre (pe.py:524)
  EpiParams pe;
  pe.out = e->x; pe.ldo = c.width; pe.pos = e->w.pos; pe.patches = e->patches;
  OVO_TRY(launch_gemm(EPI_PATCH, e->patch_buf, e->w.patch_kpad, static_cast<const __nv_bfloat16*>(e->w.patch_w), e->w.patch_kpad,
                      n_img * e->patches, c.width, e->w.patch_kpad, pe, s));
  cls_rows_kernel<<<ceil_div(n_img * c.width, 256), 256, 0, s>>>(e->x, n_img, e->seq, c.width, e->w.cls_pos0);
  OVO_CHECK_LAUNCH();
  OVO_TRY(launch_ln(e->x, rows, c.width, e->w.ln_pre_w, e->w.ln_pre_b, c.ln_eps, nullptr, e->x, nullptr, s));
  OVO_TRY(run_blocks(e, e->blocks, n_layers, n_img, e->seq, e->seq_pad, c.heads, c.width, c.mlp_width, true, false, s));
  if (apply_ln_post) {
    OVO_TRY(launch_ln(e->x, rows, c.width, e->w.ln_post_w, e->w.ln_post_b, c.ln_eps, nullptr, e->xfinal, nullptr, s));
  } else {
    OVO_CUDA(cudaMemcpyAsync(e->xfinal, e->x, static_cast<size_t>(rows) * c.width * sizeof(float), cudaMemcpyDeviceToDevice, s));
  }
  return OVO_OK;
}

int ovo_encoder_forward(ovo_encoder_t* e, int n_img, int n_layers, int apply_ln_post, float* tokens_out_dev, void* stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  OVO_REQUIRE(e && n_img > 0 && n_img <= e->max_images, "ovo_encoder_forward: bad n_img %d", n_img);
  if (n_layers < 0 || n_layers > e->cfg.layers) n_layers = e->cfg.layers;
  const long long key = (static_cast<long long>(n_img) << 16) | (n_layers << 1) | (apply_ln_post ? 1 : 0);
  ovo_encoder::GraphEntry& ge = e->graphs[key];
  if (!e->use_graphs || ge.warm == 0 || profiling()) {
    // first call per shape runs eagerly (sets kernel attributes, validates), later calls replay a graph
    OVO_TRY(forward_eager(e, n_img, n_layers, apply_ln_post, s));
    if (!profiling()) ge.warm = 1;
  } else {
    if (ge.exec == nullptr) {
      const long long before = ovo_launch_count(0);
      cudaGraph_t graph = nullptr;
      if (!e->cap_stream) {
        // captured kernel nodes inherit the capture stream's priority: the encoder's (resource-hungry, one CTA per SM) kernels get the
        // highest one, so that map kernels running beside them on another stream only take what the encoder leaves (OVO_B200_ENC_PRIO=0: off)
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        const char* pe = getenv("OVO_B200_ENC_PRIO");
        OVO_CUDA(cudaStreamCreateWithPriority(&e->cap_stream, cudaStreamNonBlocking, (pe && pe[0] == '0') ? lo : hi));
      }
      OVO_CUDA(cudaStreamBeginCapture(e->cap_stream, cudaStreamCaptureModeThreadLocal));
      const int r = forward_eager(e, n_img, n_layers, apply_ln_post, e->cap_stream);
      const cudaError_t ce = cudaStreamEndCapture(e->cap_stream, &graph);
      if (r != OVO_OK) { if (graph) cudaGraphDestroy(graph); return r; }
      if (ce != cudaSuccess) return set_error(OVO_E_CUDA, "graph capture failed: %s", cudaGetErrorString(ce));
      const cudaError_t ie = cudaGraphInstantiate(&ge.exec, graph, 0);
      cudaGraphDestroy(graph);
      if (ie != cudaSuccess) { ge.exec = nullptr; return set_error(OVO_E_CUDA, "graph instantiate failed: %s", cudaGetErrorString(ie)); }
      ge.launches = ovo_launch_count(0) - before;
      count_launch(-static_cast<int>(ge.launches));  // captured, not executed yet
    }
    OVO_CUDA(cudaGraphLaunch(ge.exec, s));
    count_launch(static_cast<int>(ge.launches));
  }
  if (tokens_out_dev)
    OVO_CUDA(cudaMemcpyAsync(tokens_out_dev, e->xfinal, static_cast<size_t>(n_img) * e->seq * e->cfg.width * sizeof(float), cudaMemcpyDeviceToDevice, s));
  return OVO_OK;
}

// E3-E5 for n_frames frames whose images start at img0 (imgs_per_frame each); masks concatenated over frames.
static int pool_regions_batch(ovo_encoder* e, int img0, int n_frames, int H, int W, const uint8_t* masks_dev,
                              const int* masks_per_frame, float* out_dev, cudaStream_t s) {
  const ovo_vit_cfg& c = e->cfg;
  const int S = c.image_size, g = e->grid;
  const int nh = std::max(H / S, 1), nw = std::max(W / S, 1), per = 1 + nh * nw;
  OVO_REQUIRE(H <= e->max_h && W <= e->max_w && img0 >= 0 && img0 + n_frames * per <= e->max_images, "pool_regions: frames out of range");
  const int ph = nh * g, pw = nw * g, P = ph * pw;
  std::vector<MaskGroup> groups;
  int M = 0;
  for (int f = 0; f < n_frames; ++f) {
    OVO_REQUIRE(masks_per_frame[f] >= 0, "pool_regions: negative mask count");
    for (int m = 0; m < masks_per_frame[f]; m += kMeanMasks)
      groups.push_back({f, M + m, std::min(kMeanMasks, masks_per_frame[f] - m), 0});
    M += masks_per_frame[f];
  }
  if (M == 0) return OVO_OK;
  OVO_REQUIRE(M <= e->max_masks, "pool_regions: %d masks exceed max_masks %d", M, e->max_masks);
  const float* tokens = e->xfinal + static_cast<size_t>(img0) * e->seq * c.width;
  ProfScope prof(s, PROF_POOL, 0.0, static_cast<double>(n_frames) * P * c.width * 8 + static_cast<double>(M) * H * W);
  OVO_CUDA(cudaMemcpyAsync(e->groups_dev, groups.data(), groups.size() * sizeof(MaskGroup), cudaMemcpyHostToDevice, s));
  token_canvas_kernel<<<dim3(P, n_frames), 256, 0, s>>>(tokens, g, c.width, nh, nw, e->canvas, per);
  OVO_CHECK_LAUNCH();
  OVO_CUDA(cudaMemsetAsync(e->fcnt, 0, M * sizeof(int), s));
  feature_mask_kernel<<<dim3(ceil_div(P, 128), M), 128, 0, s>>>(masks_dev, M, H, W, ph, pw, e->fmask, e->fcnt);
  OVO_CHECK_LAUNCH();
  const int tokens_per_split = std::max(kMeanSlice, ceil_div(ceil_div(P, kMeanMaxSplits), kMeanSlice) * kMeanSlice);
  const int splits = ceil_div(P, tokens_per_split);
  masked_sum_kernel<<<dim3(ceil_div(c.width, 256), static_cast<unsigned>(groups.size()), splits), 256, 0, s>>>(
      e->canvas, P, c.width, e->fmask, e->groups_dev, e->mean_acc, M, tokens_per_split);
  OVO_CHECK_LAUNCH();
  mean_finalize_kernel<<<ceil_div(static_cast<long long>(M) * c.width, 256), 256, 0, s>>>(e->mean_acc, splits, e->fcnt, M, c.width, e->mean);
  OVO_CHECK_LAUNCH();
  EpiParams ep;
  ep.out = out_dev; ep.ldo = c.output_dim; ep.bias = e->w.pool_b;
  OVO_TRY(launch_gemm(EPI_F32, e->mean, c.width, static_cast<const __nv_bfloat16*>(e->w.pool_w), c.width, M, c.output_dim, c.width, ep, s));
  l2_normalize_kernel<<<ceil_div(M, 8), 256, 0, s>>>(out_dev, M, c.output_dim, e->fcnt, e->w.pool_b_empty);
  OVO_CHECK_LAUNCH();
  return OVO_OK;
}

int ovo_encoder_pool_regions(ovo_encoder_t* e, int img0, int H, int W, const uint8_t* masks_dev, int M, float* out_dev,
                             void* stream_) {
  OVO_REQUIRE(e && masks_dev && out_dev, "ovo_encoder_pool_regions: null argument");
  OVO_REQUIRE(M > 0 && M <= e->max_masks, "ovo_encoder_pool_regions: M=%d outside (0, %d]", M, e->max_masks);
  return pool_regions_batch(e, img0, 1, H, W, masks_dev, &M, out_dev, static_cast<cudaStream_t>(stream_));
}

int ovo_encode_regions(ovo_encoder_t* e, const uint8_t* rgb_dev, int n_frames, int H, int W, const uint8_t* masks_dev,
                       const int* masks_per_frame_host, float* out_dev, void* stream) {
  OVO_REQUIRE(e && masks_per_frame_host && masks_dev && out_dev, "ovo_encode_regions: null argument");
  int per = 0;
  OVO_TRY(ovo_encoder_preprocess(e, rgb_dev, n_frames, H, W, &per, stream));
  OVO_TRY(ovo_encoder_forward(e, n_frames * per, -1, 1, nullptr, stream));
  return pool_regions_batch(e, 0, n_frames, H, W, masks_dev, masks_per_frame_host, out_dev, static_cast<cudaStream_t>(stream));
}

int ovo_encode_text(ovo_encoder_t* e, const int32_t* tokens_dev, int T, float* out_dev, void* stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  OVO_REQUIRE(e && tokens_dev && out_dev && T > 0, "ovo_encode_text: bad arguments");
  const ovo_vit_cfg& c = e->cfg;
  OVO_REQUIRE(c.text_layers > 0, "encoder was created without a text tower");
  const int ctx = c.text_ctx, W = c.text_width;
  OVO_REQUIRE(T <= 256, "ovo_encode_text: at most 256 strings per call (got %d)", T);
  text_embed_kernel<<<T * ctx, 256, 0, s>>>(tokens_dev, T, ctx, W, e->w.tok_emb, e->w.text_pos, e->x, e->eot_rows);
  OVO_CHECK_LAUNCH();
  OVO_TRY(run_blocks(e, e->text_blocks, c.text_layers, T, ctx, 128, c.text_heads, W, c.text_mlp_width, false, true, s));
  // ln_final on the EOT rows only (LayerNorm is row-wise), then @ text_projection (pe.py:684-693)
  OVO_TRY(launch_ln(e->x, T, W, e->w.ln_final_w, e->w.ln_final_b, c.ln_eps, e->pooled_in, nullptr, e->eot_rows, s));
  EpiParams ep;
  ep.out = out_dev; ep.ldo = c.text_output_dim;
  OVO_TRY(launch_gemm(EPI_F32, e->pooled_in, W, static_cast<const __nv_bfloat16*>(e->w.text_proj_w), W, T, c.text_output_dim, W, ep, s));
  return OVO_OK;
}

int ovo_text_bank(ovo_encoder_t* e, const int32_t* tokens_dev, int Q, int T, float* out_dev, void* stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  OVO_REQUIRE(e && tokens_dev && out_dev && Q > 0 && T > 0, "ovo_text_bank: bad arguments");
  OVO_REQUIRE(T <= 256, "ovo_text_bank: at most 256 templates per query");
  const int D = e->cfg.text_output_dim;
  float* emb = e->xfinal;  // scratch: [chunk*T, D] f32 (xfinal is not live during a text pass)
  OVO_REQUIRE(static_cast<size_t>(256) * D <= (e->rows_cap + 128) * static_cast<size_t>(std::max(e->cfg.width, e->cfg.text_width)), "ovo_text_bank: scratch too small");
  const int qchunk = std::max(1, 256 / T);
  for (int q0 = 0; q0 < Q; q0 += qchunk) {
    const int qn = std::min(qchunk, Q - q0);
    OVO_TRY(ovo_encode_text(e, tokens_dev + static_cast<size_t>(q0) * T * e->cfg.text_ctx, qn * T, emb, stream_));
    text_bank_kernel<<<ceil_div(qn, 8), 256, 0, s>>>(emb, qn, T, D, out_dev + static_cast<size_t>(q0) * D);
    OVO_CHECK_LAUNCH();
  }
  return OVO_OK;
}

// ------------------------------------------------------------------------------------------- crop-based descriptors
int ovo_encoder_set_pool_head(ovo_encoder_t* e, const ovo_pool_head_weights* w) {
  OVO_REQUIRE(e && w, "ovo_encoder_set_pool_head: null argument");
  const ovo_vit_cfg& c = e->cfg;
  OVO_REQUIRE(w->heads > 0 && c.width % w->heads == 0 && c.width / w->heads <= 256, "pool head: heads %d unsupported for width %d", w->heads, c.width);
  OVO_REQUIRE(w->mlp_width > 0 && w->mlp_width % 8 == 0, "pool head: bad mlp_width %d", w->mlp_width);
  OVO_REQUIRE(w->q && w->kv_w && w->kv_b && w->out_w && w->out_b && w->ln_w && w->ln_b && w->fc_w && w->fc_b && w->proj_w && w->proj_b && w->vis_proj_w,
              "pool head: null weight pointer");
  // keys | values of every token live (f32) in the MLP's hidden buffer: [rows, 2*width] f32 <= [rows, mlp] bf16
  OVO_REQUIRE(std::max(c.mlp_width, c.text_mlp_width) >= 4 * c.width, "pool head: hidden buffer too small for the key/value projections");
  if (!e->ph_x1) {
    const size_t rows = static_cast<size_t>(e->max_images) + 128;
    int r = OVO_OK;
    r |= dmalloc(&e->ph_x1, rows * c.width);
    r |= dmalloc(&e->ph_x2, rows * c.width);
    r |= dmalloc(&e->ph_a, rows * c.width);
    r |= dmalloc(&e->ph_h, rows * w->mlp_width);
    r |= dmalloc(&e->ph_emb, (static_cast<size_t>(2) * e->max_masks + 1 + 128) * c.output_dim);
    r |= dmalloc(&e->crop_boxes, static_cast<size_t>(e->max_masks) * 4);
    r |= dmalloc(&e->crop_jobs, static_cast<size_t>(2) * e->max_masks);
    if (r != OVO_OK) return OVO_E_NOMEM;
  }
  e->head = *w;
  e->has_head = true;
  return OVO_OK;
}

// pe.py:493 (attn_pool) + :540-541 (@ proj) on the n images of the last forward (tokens after ln_post in xfinal)
static int pool_head(ovo_encoder* e, int n, float* out, cudaStream_t s) {
  const ovo_vit_cfg& c = e->cfg;
  const ovo_pool_head_weights& h = e->head;
  const int W = c.width, rows = n * e->seq, hd = W / h.heads, mlp = h.mlp_width;
  typedef const __nv_bfloat16* bfp;
  {
    const size_t n4 = static_cast<size_t>(rows) * W / 4;
    f32_to_bf16_kernel<<<ceil_div(n4, 256), 256, 0, s>>>(e->xfinal, n4, e->xn);
    OVO_CHECK_LAUNCH();
  }
  float* kv = reinterpret_cast<float*>(e->hmid);
  EpiParams kp;
  kp.out = kv; kp.ldo = 2 * W; kp.bias = h.kv_b;
  OVO_TRY(launch_gemm(EPI_F32, e->xn, W, static_cast<bfp>(h.kv_w), W, rows, 2 * W, W, kp, s));
  {
    ProfScope prof(s, PROF_POOL, 4.0 * n * e->seq * W, static_cast<double>(rows) * 2 * W * 4);
    pool_attention_kernel<<<dim3(n, h.heads), 256, (e->seq + 256) * sizeof(float), s>>>(kv, h.q, e->seq, W, hd, e->ph_a);
    OVO_CHECK_LAUNCH();
  }
  EpiParams op;
  op.out = e->ph_x1; op.ldo = W; op.bias = h.out_b;
  OVO_TRY(launch_gemm(EPI_F32, e->ph_a, W, static_cast<bfp>(h.out_w), W, n, W, W, op, s));
  OVO_TRY(launch_ln(e->ph_x1, n, W, h.ln_w, h.ln_b, c.ln_eps, e->ph_a, nullptr, nullptr, s));
  EpiParams fc;
  fc.out = e->ph_h; fc.ldo = mlp; fc.bias = h.fc_b;
  OVO_TRY(launch_gemm(EPI_BF16_GELU, e->ph_a, W, static_cast<bfp>(h.fc_w), W, n, mlp, W, fc, s));
  EpiParams pj;
  pj.out = e->ph_x2; pj.ldo = W; pj.bias = h.proj_b; pj.resid = e->ph_x1; pj.ldr = W;
  OVO_TRY(launch_gemm(EPI_F32_RESID, e->ph_h, mlp, static_cast<bfp>(h.proj_w), mlp, n, W, mlp, pj, s));
  {
    const size_t n4 = static_cast<size_t>(n) * W / 4;
    f32_to_bf16_kernel<<<ceil_div(n4, 256), 256, 0, s>>>(e->ph_x2, n4, e->ph_a);
    OVO_CHECK_LAUNCH();
  }
  EpiParams vp;
  vp.out = out; vp.ldo = c.output_dim;
  OVO_TRY(launch_gemm(EPI_F32, e->ph_a, W, static_cast<bfp>(h.vis_proj_w), W, n, c.output_dim, W, vp, s));
  return OVO_OK;
}

int ovo_encode_images(ovo_encoder_t* e, const float* pixels_dev, int n, float* out_dev, void* stream) {
  OVO_REQUIRE(e && pixels_dev && out_dev && n > 0, "ovo_encode_images: bad arguments");
  OVO_REQUIRE(e->has_head, "ovo_encode_images: call ovo_encoder_set_pool_head first");
  OVO_TRY(ovo_encoder_load_pixels(e, pixels_dev, n, stream));
  OVO_TRY(ovo_encoder_forward(e, n, -1, 1, nullptr, stream));
  return pool_head(e, n, out_dev, static_cast<cudaStream_t>(stream));
}

int ovo_mask_boxes(const uint8_t* masks_dev, int M, int H, int W, int32_t* xywh_dev, void* stream) {
  OVO_REQUIRE(masks_dev && xywh_dev && M > 0 && H > 0 && W > 0, "ovo_mask_boxes: bad arguments");
  mask_boxes_kernel<<<M, 1024, 0, static_cast<cudaStream_t>(stream)>>>(masks_dev, H, W, xywh_dev);
  OVO_CHECK_LAUNCH();
  return OVO_OK;
}

int ovo_fuse_clips(const float* g_dev, const float* seg_dev, const float* bbox_dev, int M, int D, int embed_type,
                   float w_masked, float w_global, float* out_dev, void* stream) {
  OVO_REQUIRE(g_dev && seg_dev && bbox_dev && out_dev && M > 0 && D > 0, "ovo_fuse_clips: bad arguments");
  OVO_REQUIRE(embed_type >= OVO_EMBED_FIXED_WEIGHTS && embed_type <= OVO_EMBED_CONCEPT_FUSION, "ovo_fuse_clips: embed_type %d has no fusion rule", embed_type);
  OVO_REQUIRE(M <= 10000, "ovo_fuse_clips: at most 10000 masks per frame");
  fuse_clips_kernel<<<1, 1024, M * sizeof(float), static_cast<cudaStream_t>(stream)>>>(g_dev, seg_dev, bbox_dev, M, D, embed_type, w_masked, w_global, out_dev);
  OVO_CHECK_LAUNCH();
  return OVO_OK;
}

int ovo_siglip_similarity(float* sim_dev, int64_t n, float logit_scale, float logit_bias, void* stream) {
  OVO_REQUIRE(sim_dev && n > 0, "ovo_siglip_similarity: bad arguments");
  sigmoid_affine_kernel<<<ceil_div(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(sim_dev, static_cast<size_t>(n), expf(logit_scale), logit_bias);
  OVO_CHECK_LAUNCH();
  return OVO_OK;
}

}  // extern "C"

template <typename T>
static int grow(T** p, size_t* cap, size_t need) {
  if (*cap >= need) return OVO_OK;
  cudaDeviceSynchronize();   // the old buffer may still be in use by queued kernels
  cudaFree(*p);
  *p = nullptr; *cap = 0;
  if (cudaMalloc(reinterpret_cast<void**>(p), need * sizeof(T)) != cudaSuccess) {
    cudaGetLastError();
    return set_error(OVO_E_NOMEM, "crop workspace allocation of %zu bytes failed", need * sizeof(T));
  }
  *cap = need;
  return OVO_OK;
}

extern "C" {

int ovo_encode_crops(ovo_encoder_t* e, const uint8_t* rgb_dev, int H, int W, const uint8_t* masks_dev, int M,
                     const ovo_crop_params* prm, float* out_dev, uint8_t* crops_out_dev, void* stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  OVO_REQUIRE(e && rgb_dev && masks_dev && prm && out_dev, "ovo_encode_crops: null argument");
  OVO_REQUIRE(e->has_head, "ovo_encode_crops: call ovo_encoder_set_pool_head first");
  OVO_REQUIRE(M > 0 && M <= e->max_masks, "ovo_encode_crops: M=%d outside (0, %d]", M, e->max_masks);
  OVO_REQUIRE(H > 0 && W > 0 && H <= e->max_h && W <= e->max_w, "frame %dx%d exceeds encoder limits %dx%d", H, W, e->max_h, e->max_w);
  OVO_REQUIRE(prm->embed_type >= OVO_EMBED_VANILLA && prm->embed_type <= OVO_EMBED_CONCEPT_FUSION, "ovo_encode_crops: unknown embed_type %d", prm->embed_type);
  const int L = prm->mask_res, S = e->cfg.image_size, D = e->cfg.output_dim;
  OVO_REQUIRE(L > 0 && L <= e->max_h && L <= 2048, "ovo_encode_crops: mask_res %d outside (0, min(max_h, 2048)]", L);
  const bool vanilla = prm->embed_type == OVO_EMBED_VANILLA;

  // 1. boxes -> host (they decide the geometry of every crop; the one synchronisation of this path)
  mask_boxes_kernel<<<M, 1024, 0, s>>>(masks_dev, H, W, e->crop_boxes);
  OVO_CHECK_LAUNCH();
  std::vector<int32_t> bx(static_cast<size_t>(M) * 4);
  OVO_CUDA(cudaMemcpyAsync(bx.data(), e->crop_boxes, bx.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  OVO_CUDA(cudaStreamSynchronize(s));

  // 2. crop jobs: masked crops first, then margin crops (clip_generator.py:148)
  std::vector<CropJob>& jobs = e->crop_jobs_host;
  jobs.clear();
  int maxdim = 1;
  for (int m = 0; m < M; ++m) {
    const int x = bx[4 * m], y = bx[4 * m + 1], w = bx[4 * m + 2], h = bx[4 * m + 3];
    if (vanilla) {   // pad_img (segment_utils.py:149-157): centred in a zero square of side max(w, h)
      const int side = std::max(w, h);
      OVO_REQUIRE(side > 0, "mask %d: empty crop (box %dx%d); the reference's F.resize raises here", m, w, h);
      jobs.push_back({m, x, y, w, h, h > w ? (h - w) / 2 : 0, h > w ? 0 : (w - h) / 2, side, side});
    } else {
      OVO_REQUIRE(w > 0 && h > 0, "mask %d: empty crop (box %dx%d); the reference's F.resize raises here", m, w, h);
      jobs.push_back({m, x, y, w, h, 0, 0, w, h});
    }
    maxdim = std::max(maxdim, std::max(jobs.back().vw, jobs.back().vh));
  }
  if (!vanilla)
    for (int m = 0; m < M; ++m) {   // increase_bbox_by_margin (segment_utils.py:159-182); slicing clamps right / bottom
      int x = bx[4 * m] - prm->bbox_margin, y = bx[4 * m + 1] - prm->bbox_margin;
      int w = bx[4 * m + 2] + 2 * prm->bbox_margin, h = bx[4 * m + 3] + 2 * prm->bbox_margin;
      if (x < 0) { w += x; x = 0; }
      if (y < 0) { h += y; y = 0; }
      w = std::min(x + w, W) - x; h = std::min(y + h, H) - y;
      OVO_REQUIRE(w > 0 && h > 0, "mask %d: empty margin crop", m);
      jobs.push_back({-1, x, y, w, h, 0, 0, w, h});
      maxdim = std::max(maxdim, std::max(w, h));
    }
  const int n_crops = static_cast<int>(jobs.size()), has_g = vanilla ? 0 : 1, n_total = n_crops + has_g;
  const int kmax = static_cast<int>(std::ceil(std::max(static_cast<double>(maxdim) / L, 1.0))) * 2 + 1;
  const int cmax = e->max_images;
  OVO_CUDA(cudaMemcpyAsync(e->crop_jobs, jobs.data(), jobs.size() * sizeof(CropJob), cudaMemcpyHostToDevice, s));
  OVO_TRY(grow(&e->crop_u8, &e->crop_u8_cap, static_cast<size_t>(cmax) * L * L * 3));
  {
    size_t cap = e->ctab_cap;
    OVO_TRY(grow(&e->ctab_min, &cap, static_cast<size_t>(cmax) * 2 * L));
    cap = e->ctab_cap;
    OVO_TRY(grow(&e->ctab_size, &cap, static_cast<size_t>(cmax) * 2 * L));
    e->ctab_cap = cap;
    OVO_TRY(grow(&e->ctab_w, &e->ctab_wcap, static_cast<size_t>(cmax) * 2 * L * kmax));
  }
  // second stage (the encoder's own Resize((S,S)) + Normalize, clip_generator.py:119): E1's kernels read the uint8
  // crops like frames of L x L pixels; the frame's global image is job 0
  AaTable tl, tgx, tgy;
  OVO_TRY(get_table(e, L, &tl, s));
  OVO_TRY(get_table(e, W, &tgx, s));
  OVO_TRY(get_table(e, H, &tgy, s));
  {
    std::vector<ImgJob>& ij = e->jobs_host;
    ij.clear();
    ij.push_back({0, 0, 0, H, W, tgx.offset, tgy.offset, tgx.k, tgy.k});
    for (int k = 0; k < cmax; ++k) ij.push_back({k, 0, 0, L, L, tl.offset, tl.offset, tl.k, tl.k});
    OVO_CUDA(cudaMemcpyAsync(e->jobs_dev, ij.data(), ij.size() * sizeof(ImgJob), cudaMemcpyHostToDevice, s));
    OVO_CUDA(cudaStreamSynchronize(s));
    ij.clear();   // not the job list ovo_encoder_preprocess caches: force its next upload
  }
  const int patch = e->cfg.patch_size, kpad = e->w.patch_kpad;
  for (int i0 = 0; i0 < n_total;) {
    const int n = std::min(cmax, n_total - i0);
    const int g_here = (has_g && i0 == 0) ? 1 : 0;
    const int nc = n - g_here, c0 = i0 + g_here - has_g;
    if (nc > 0) {
      ProfScope prof(s, PROF_PRE, 0.0, static_cast<double>(nc) * L * L * 3 * 2);
      crop_tables_kernel<<<dim3(ceil_div(L, 128), 2, nc), 128, 0, s>>>(e->crop_jobs + c0, L, kmax, e->ctab_min, e->ctab_size, e->ctab_w);
      OVO_CHECK_LAUNCH();
      crop_resize_kernel<<<dim3(ceil_div(L, 128), L, nc), 128, 0, s>>>(rgb_dev, H, W, masks_dev, e->crop_jobs + c0, L, kmax, e->ctab_min,
                                                                      e->ctab_size, e->ctab_w, e->crop_u8);
      OVO_CHECK_LAUNCH();
      if (crops_out_dev)
        OVO_CUDA(cudaMemcpyAsync(crops_out_dev + static_cast<size_t>(c0) * L * L * 3, e->crop_u8, static_cast<size_t>(nc) * L * L * 3,
                                 cudaMemcpyDeviceToDevice, s));
      aa_resize_h_kernel<<<dim3(ceil_div(S, 128), L, nc), 128, 0, s>>>(e->crop_u8, L, L, e->jobs_dev + 1, e->tab_min, e->tab_size, e->tab_w, S,
                                                                      e->resize_tmp, e->max_h, g_here);
      OVO_CHECK_LAUNCH();
      aa_resize_v_patch_kernel<<<dim3(ceil_div(S, 128), S, nc), 128, 0, s>>>(e->resize_tmp, e->max_h, e->jobs_dev + 1, e->tab_min, e->tab_size,
                                                                            e->tab_w, S, patch, kpad, e->patch_buf, g_here);
      OVO_CHECK_LAUNCH();
    }
    if (g_here) {
      aa_resize_h_kernel<<<dim3(ceil_div(S, 128), H, 1), 128, 0, s>>>(rgb_dev, H, W, e->jobs_dev, e->tab_min, e->tab_size, e->tab_w, S,
                                                                     e->resize_tmp, e->max_h, 0);
      OVO_CHECK_LAUNCH();
      aa_resize_v_patch_kernel<<<dim3(ceil_div(S, 128), S, 1), 128, 0, s>>>(e->resize_tmp, e->max_h, e->jobs_dev, e->tab_min, e->tab_size,
                                                                           e->tab_w, S, patch, kpad, e->patch_buf, 0);
      OVO_CHECK_LAUNCH();
    }
    e->loaded_images = n;
    OVO_TRY(ovo_encoder_forward(e, n, -1, 1, nullptr, stream_));
    OVO_TRY(pool_head(e, n, e->ph_emb + static_cast<size_t>(i0) * D, s));
    i0 += n;
  }
  l2_normalize_kernel<<<ceil_div(n_total, 8), 256, 0, s>>>(e->ph_emb, n_total, D, nullptr, nullptr);
  OVO_CHECK_LAUNCH();
  if (vanilla) {
    OVO_CUDA(cudaMemcpyAsync(out_dev, e->ph_emb, static_cast<size_t>(M) * D * sizeof(float), cudaMemcpyDeviceToDevice, s));
  } else if (prm->return_all) {
    stack_clips_kernel<<<ceil_div(static_cast<long long>(M) * 3 * D, 256), 256, 0, s>>>(e->ph_emb, e->ph_emb + D, e->ph_emb + static_cast<size_t>(1 + M) * D, M, D, out_dev);
    OVO_CHECK_LAUNCH();
  } else if (prm->embed_type == OVO_EMBED_VANILLA) {
  } else {
    fuse_clips_kernel<<<1, 1024, M * sizeof(float), s>>>(e->ph_emb, e->ph_emb + D, e->ph_emb + static_cast<size_t>(1 + M) * D, M, D, prm->embed_type,
                                                        prm->w_masked, prm->w_global, out_dev);
    OVO_CHECK_LAUNCH();
  }
  return OVO_OK;
}

int ovo_merge_clips_learned(const ovo_merger_weights* w, const float* clips_dev, int B, float* out_dev, void* stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  OVO_REQUIRE(w && clips_dev && out_dev && B > 0, "ovo_merge_clips_learned: bad arguments");
  const int d = w->d_model, ff = w->dim_feedforward, R = 3 * B;
  OVO_REQUIRE(d > 0 && w->nhead > 0 && d % w->nhead == 0 && d % 8 == 0 && d <= 2048 && ff > 0 && ff % 8 == 0, "merger: unsupported d_model %d / nhead %d / ff %d", d, w->nhead, ff);
  OVO_REQUIRE(w->n_layers >= 0 && (w->n_layers == 0 || w->layers) && w->n_linear >= 1 && w->mlp_w && w->mlp_b && w->mlp_out, "merger: missing weights");
  const int o_dim = w->mlp_out[w->n_linear - 1];
  OVO_REQUIRE(o_dim == 3 || o_dim == 3 * d, "merger: the MLP must emit 3 or 3*d_model weights (got %d)", o_dim);
  int maxdim = 3 * d;
  for (int j = 0; j < w->n_linear; ++j) {
    OVO_REQUIRE(w->mlp_out[j] > 0 && (j + 1 == w->n_linear || w->mlp_out[j] % 8 == 0), "merger: MLP width %d unsupported", w->mlp_out[j]);
    maxdim = std::max(maxdim, w->mlp_out[j]);
  }
  typedef const __nv_bfloat16* bfp;
  keep_default_mempool_cached();
  float *x = nullptr, *x2 = nullptr, *qkv = nullptr, *y = nullptr;
  __nv_bfloat16 *xb = nullptr, *att = nullptr, *h = nullptr, *a0 = nullptr, *a1 = nullptr;
  const size_t Rp = static_cast<size_t>(R) + 128, Bp = static_cast<size_t>(B) + 128;
  AsyncTemps tmp(s);
  OVO_CUDA(tmp.alloc(&x, Rp * d * 4));
  OVO_CUDA(tmp.alloc(&x2, Rp * d * 4));
  OVO_CUDA(tmp.alloc(&qkv, Rp * 3 * d * 4));
  OVO_CUDA(tmp.alloc(&xb, Rp * d * 2));
  OVO_CUDA(tmp.alloc(&att, Rp * d * 2));
  OVO_CUDA(tmp.alloc(&h, Rp * ff * 2));
  OVO_CUDA(tmp.alloc(&y, Bp * maxdim * 4));
  OVO_CUDA(tmp.alloc(&a0, Bp * maxdim * 2));
  OVO_CUDA(tmp.alloc(&a1, Bp * maxdim * 2));
  OVO_CUDA(cudaMemcpyAsync(x, clips_dev, static_cast<size_t>(R) * d * 4, cudaMemcpyDeviceToDevice, s));
  {
    const size_t n4 = static_cast<size_t>(R) * d / 4;
    f32_to_bf16_kernel<<<ceil_div(n4, 256), 256, 0, s>>>(x, n4, xb);
    OVO_CHECK_LAUNCH();
  }
  for (int l = 0; l < w->n_layers; ++l) {   // nn.TransformerEncoderLayer, norm_first=False, relu (clips_merging.py:29-36)
    const ovo_merger_layer& L = w->layers[l];
    EpiParams qe;
    qe.out = qkv; qe.ldo = 3 * d; qe.bias = L.in_b;
    OVO_TRY(launch_gemm(EPI_F32, xb, d, static_cast<bfp>(L.in_w), d, R, 3 * d, d, qe, s));
    merger_attention_kernel<<<ceil_div(B * w->nhead, 128), 128, 0, s>>>(qkv, B, d, w->nhead, att);
    OVO_CHECK_LAUNCH();
    EpiParams oe;
    oe.out = x2; oe.ldo = d; oe.bias = L.out_b; oe.resid = x; oe.ldr = d;
    OVO_TRY(launch_gemm(EPI_F32_RESID, att, d, static_cast<bfp>(L.out_w), d, R, d, d, oe, s));
    OVO_TRY(launch_ln(x2, R, d, L.ln1_w, L.ln1_b, w->ln_eps, xb, x, nullptr, s));
    EpiParams f1;
    f1.out = h; f1.ldo = ff; f1.bias = L.ff1_b;
    OVO_TRY(launch_gemm(EPI_BF16_RELU, xb, d, static_cast<bfp>(L.ff1_w), d, R, ff, d, f1, s));
    EpiParams f2;
    f2.out = x2; f2.ldo = d; f2.bias = L.ff2_b; f2.resid = x; f2.ldr = d;
    OVO_TRY(launch_gemm(EPI_F32_RESID, h, ff, static_cast<bfp>(L.ff2_w), ff, R, d, ff, f2, s));
    OVO_TRY(launch_ln(x2, R, d, L.ln2_w, L.ln2_b, w->ln_eps, xb, x, nullptr, s));
  }
  // MLP on the flattened tokens: xb [3B, d] row-major IS [B, 3d]
  const __nv_bfloat16* cur = xb;
  int in_dim = 3 * d;
  for (int j = 0; j < w->n_linear; ++j) {
    const int out_dim = w->mlp_out[j];
    EpiParams me;
    me.out = y; me.ldo = out_dim; me.bias = w->mlp_b[j];
    OVO_TRY(launch_gemm(EPI_F32, cur, in_dim, static_cast<bfp>(w->mlp_w[j]), in_dim, B, out_dim, in_dim, me, s));
    if (j + 1 < w->n_linear) {
      __nv_bfloat16* nxt = (j & 1) ? a1 : a0;
      const size_t n = static_cast<size_t>(B) * out_dim;
      leaky_relu_bf16_kernel<<<ceil_div(n, 256), 256, 0, s>>>(y, n, nxt);
      OVO_CHECK_LAUNCH();
      cur = nxt;
    }
    in_dim = out_dim;
  }
  merger_combine_kernel<<<ceil_div(B, 8), 256, 0, s>>>(clips_dev, y, B, d, o_dim, out_dev);
  OVO_CHECK_LAUNCH();
  return OVO_OK;   // `tmp` releases the temporaries in stream order
}

}  // extern "C"
