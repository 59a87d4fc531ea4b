// Persistent, warp-specialised bf16 GEMM for sm_100a:  C[M,N] = A[M,K] . B[N,K]^T  (+ fused epilogue)
//   warp 0      : TMA producer (cp.async.bulk.tensor, 128B-swizzled 128x64 / BNx64 tiles, mbarrier ring)
//   warp 1      : tcgen05.mma issuer (one elected thread), accumulators double-buffered in TMEM
//   warps 2..9  : epilogue (tcgen05.ld 32x32b -> registers -> fused epilogue -> global); two warps per TMEM
//                 lane quadrant, each taking every other 32-column chunk
// Both operands are K-major (row-major with K contiguous), which is how activations [tokens, width]
// and nn.Linear weights [out, in] already sit in memory, so no transposes are ever materialised.
//
// Used for every linear layer of the ViT / text tower (reference pe.py:125,150,190-198,509), the region
// projection (textregion.py:183-195) and the dense text-vs-map cosine query (clip_utils.py:16-19).
#pragma once
#include "ptx.cuh"

namespace ovo {

enum EpiKind : int {
  EPI_F32 = 0,        // out_f32 = acc (+bias)
  EPI_BF16 = 1,       // out_bf16 = acc (+bias)
  EPI_BF16_GELU = 2,  // out_bf16 = gelu(acc + bias)            (pe.py:190-198)
  EPI_F32_RESID = 3,  // out_f32 = acc + bias + resid            (pe.py:221-224)
  EPI_QKV = 4,        // split q/k/v head-major [b,h,seq_pad,64], 2D-RoPE on q,k (pe.py:125-143, rope.py:40-62)
  EPI_PATCH = 5,      // out_f32[token row] = acc + pos_emb      (pe.py:509-519)
  EPI_BF16_RELU = 6,  // out_bf16 = relu(acc + bias)            (SAM-2 two-way transformer MLP, sam/transformer.py:161-163)
  EPI_GELU_DOT = 7,   // SAM-2 mask decoder, second transposed conv fused with the hyper-network product (mask_decoder.py:217,
                      // 225-226): v = gelu(acc + bias + resid[row % resid_mod]) is one final pixel's 32 channels per 32-column
                      // chunk; masks[p, m, Y, X] = v . dot_w[p, 1+m, :] for the 3 multimask tokens.  The up-scaled embedding is
                      // never written.
  EPI_UP_LN = 8,      // SAM-2 mask decoder, first transposed conv + LayerNorm2d + GELU (mask_decoder.py:214-216): per 64-column
                      // group (one sub-pixel): out_bf16 = gelu(ln_w * normalize(acc + bias + resid[row % resid_mod]) + ln_b)
};

struct EpiParams {
  void* out = nullptr;
  int ldo = 0;
  const float* bias = nullptr;
  const float* resid = nullptr;
  int ldr = 0;
  int resid_mod = 0;           // > 0: the residual row is (row % resid_mod) — one [resid_mod, N] table shared by every batch
                               // item (SAM-2 mask decoder: image embedding / high-res features broadcast over the prompts).
                               // EPI_F32_RESID adds it after the bias; EPI_BF16_GELU adds it BEFORE the GELU when resid != nullptr
  // EPI_QKV
  __nv_bfloat16* q = nullptr;
  __nv_bfloat16* k = nullptr;
  __nv_bfloat16* vt = nullptr;
  const float2* rope_tab = nullptr;  // [grid+1][16] (cos, sin) of r*theta_i, or nullptr (text tower: no RoPE)
  int rope_grid = 0;                 // patch grid side (24); token t>0 sits at (y,x) = divmod(t-1, grid)
  int seq = 0, seq_pad = 0, heads = 0, width = 0;
  // EPI_PATCH
  const float* pos = nullptr;  // [1 + patches, width]
  int patches = 0;             // patches per image (576)
  // EPI_GELU_DOT
  const float* dot_w = nullptr;   // [P, 4, 32] hyper-network outputs
  float* dot_out = nullptr;       // [P, 3, 4g, 4g] mask logits
  int dot_g = 0;                  // image-embedding grid side g (rows are (p, y, x, sub1) over g x g, columns (sub2, 32 channels))
  // EPI_UP_LN
  const float* ln_w = nullptr;    // [64]
  const float* ln_b = nullptr;
  int prof_cls = 0;            // profiling class of this launch (common.cuh ProfClass; host side only)
  int debug = 0;               // tuning experiments: 1 = no epilogue, 2 = no MMA, 4 = no TMA loads, 8 = no GELU math,
                               // 16 = no global stores, 32 = no bias loads, 64 = no residual prefetch
};

constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kGemmThreads = 320;   // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (two per TMEM lane quadrant)
constexpr int kRopeRowsMax = 40;     // rows of the per-axis RoPE table staged in shared memory (grid+1 <= 40)

template <int BN>
struct GemmCfg {
  static constexpr int kStages = BN >= 256 ? 4 : (BN >= 128 ? 6 : 8);
  static constexpr int kBytesA = kBM * kBK * 2;
  static constexpr int kBytesB = BN * kBK * 2;
  static constexpr int kStageBytes = kBytesA + kBytesB;
  static constexpr int kTmemCols = 2 * BN < 32 ? 32 : 2 * BN;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/ + kRopeRowsMax * 17 * 8 /*rope table*/ +
                                    8 * 2048 /*epilogue staging, 32 rows x 64 B per warp*/;
};

// Exact (erf) GELU of nn.GELU (pe.py:301).  erf by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, far below the bf16
// rounding of the output): ~14 instructions instead of erff's ~25 — the fc epilogue is issue-bound otherwise.
__device__ __forceinline__ float gelu_erf(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = fast_rcp(fmaf(0.3275911f, z, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float e = 1.0f - p * t * fast_ex2(-1.4426950408889634f * z * z);   // erf(|x|/sqrt2)
  return 0.5f * x * (1.0f + copysignf(e, x));
}

// One thread owns `row` and 32 consecutive accumulator columns starting at `col`.
template <int EPI>
__device__ __forceinline__ void epilogue_store(const EpiParams& ep, int row, int col, const uint32_t (&v)[32], int M,
                                               int N, const float2* s_rope) {
  if (row >= M || col >= N) return;
  const int ncol = min(32, N - col);
  float acc[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) acc[j] = __uint_as_float(v[j]);
  if (ep.bias != nullptr) {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < ncol) acc[j] += __ldg(ep.bias + col + j);
  }

  if constexpr (EPI == EPI_F32 || EPI == EPI_F32_RESID) {
    float* out = static_cast<float*>(ep.out) + static_cast<size_t>(row) * ep.ldo + col;
    if constexpr (EPI == EPI_F32_RESID) {
      const float* r = ep.resid + static_cast<size_t>(ep.resid_mod > 0 ? row % ep.resid_mod : row) * ep.ldr + col;
      if (ncol == 32 && (ep.ldr & 3) == 0) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 t = *reinterpret_cast<const float4*>(r + j);
          acc[j] += t.x; acc[j + 1] += t.y; acc[j + 2] += t.z; acc[j + 3] += t.w;
        }
      } else {
        for (int j = 0; j < ncol; ++j) acc[j] += r[j];
      }
    }
    if (ncol == 32 && (ep.ldo & 3) == 0) {
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(out + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
    } else {
      for (int j = 0; j < ncol; ++j) out[j] = acc[j];
    }
  } else if constexpr (EPI == EPI_BF16 || EPI == EPI_BF16_GELU || EPI == EPI_BF16_RELU) {
    if constexpr (EPI == EPI_BF16_GELU) {
      if (ep.resid != nullptr) {
        const float* r = ep.resid + static_cast<size_t>(ep.resid_mod > 0 ? row % ep.resid_mod : row) * ep.ldr + col;
        for (int j = 0; j < ncol; ++j) acc[j] += r[j];
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[j] = gelu_erf(acc[j]);
    }
    if constexpr (EPI == EPI_BF16_RELU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[j] = fmaxf(acc[j], 0.f);
    }
    __nv_bfloat16* out = static_cast<__nv_bfloat16*>(ep.out) + static_cast<size_t>(row) * ep.ldo + col;
    if (ncol == 32 && (ep.ldo & 7) == 0) {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint4 t;
        t.x = pack_bf16(acc[j], acc[j + 1]); t.y = pack_bf16(acc[j + 2], acc[j + 3]);
        t.z = pack_bf16(acc[j + 4], acc[j + 5]); t.w = pack_bf16(acc[j + 6], acc[j + 7]);
        *reinterpret_cast<uint4*>(out + j) = t;
      }
    } else {
      for (int j = 0; j < ncol; ++j) out[j] = __float2bfloat16_rn(acc[j]);
    }
  } else if constexpr (EPI == EPI_QKV) {
    // column layout of in_proj: [q | k | v], each `width` wide, heads of 64 (pe.py:128-140)
    const int which = col / ep.width;
    const int within = col - which * ep.width;
    const int head = within >> 6;
    const int d0 = within & 63;  // 0 or 32
    const int b = row / ep.seq;
    const int t = row - b * ep.seq;
    const size_t bh = static_cast<size_t>(b) * ep.heads + head;
    {
      if (ep.rope_tab != nullptr && which < 2) {
        // 2D RoPE with a cls token (rope.py:315-347): interleaved pairs (2i, 2i+1); pairs 0..15 of a head
        // (d0 == 0) rotate by (x+1)*theta_i, pairs 16..31 (d0 == 32) by (y+1)*theta_i, the cls token by 0.
        int r = 0;
        if (t > 0) r = (d0 == 0 ? (t - 1) % ep.rope_grid : (t - 1) / ep.rope_grid) + 1;
        const float2* tab = s_rope + r * 17;
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const float2 cs = tab[j >> 1];
          const float a = acc[j], bb = acc[j + 1];
          acc[j] = a * cs.x - bb * cs.y;
          acc[j + 1] = bb * cs.x + a * cs.y;
        }
      }
      __nv_bfloat16* dst = (which == 0 ? ep.q : (which == 1 ? ep.k : ep.vt)) + (bh * ep.seq_pad + t) * 64 + d0;
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint4 u;
        u.x = pack_bf16(acc[j], acc[j + 1]); u.y = pack_bf16(acc[j + 2], acc[j + 3]);
        u.z = pack_bf16(acc[j + 4], acc[j + 5]); u.w = pack_bf16(acc[j + 6], acc[j + 7]);
        *reinterpret_cast<uint4*>(dst + j) = u;
      }
    }
  } else if constexpr (EPI == EPI_PATCH) {
    const int b = row / ep.patches;
    const int p = row - b * ep.patches;
    const size_t orow = static_cast<size_t>(b) * (ep.patches + 1) + 1 + p;
    float* out = static_cast<float*>(ep.out) + orow * ep.ldo + col;
    const float* pos = ep.pos + static_cast<size_t>(1 + p) * ep.ldo + col;
    for (int j = 0; j < ncol; ++j) out[j] = acc[j] + __ldg(pos + j);
  }
}

// ---- thread-local fused epilogues of the SAM-2 mask decoder -------------------------------------------------------
// One thread owns `row` and the 32 accumulator columns of chunk `col`: exactly the 32 channels of one output pixel.
__device__ __forceinline__ void epilogue_gelu_dot(const EpiParams& ep, int row, int col, const uint32_t (&v)[32], int M) {
  if (row >= M) return;
  float x[32];
  const float4* b4 = reinterpret_cast<const float4*>(ep.bias + col);
  const float4* r4 = reinterpret_cast<const float4*>(ep.resid + static_cast<size_t>(row % ep.resid_mod) * ep.ldr + col);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 b = __ldg(b4 + j), r = __ldg(r4 + j);
    x[4 * j] = gelu_erf(__uint_as_float(v[4 * j]) + b.x + r.x);
    x[4 * j + 1] = gelu_erf(__uint_as_float(v[4 * j + 1]) + b.y + r.y);
    x[4 * j + 2] = gelu_erf(__uint_as_float(v[4 * j + 2]) + b.z + r.z);
    x[4 * j + 3] = gelu_erf(__uint_as_float(v[4 * j + 3]) + b.w + r.w);
  }
  const int g = ep.dot_g, per = 4 * g * g;
  const int p = row / per, rem = row - p * per;
  const int t = rem >> 2, sub1 = rem & 3, sub2 = col >> 5;
  const int y = t / g, xq = t - y * g;
  const int S = 4 * g;
  const int Y = 4 * y + 2 * (sub1 >> 1) + (sub2 >> 1), X = 4 * xq + 2 * (sub1 & 1) + (sub2 & 1);
  const float4* w4 = reinterpret_cast<const float4*>(ep.dot_w + static_cast<size_t>(p) * 128 + 32);   // tokens 1..3 (multimask)
  float* o = ep.dot_out + (static_cast<size_t>(p) * 3 * S + Y) * S + X;
#pragma unroll
  for (int m = 0; m < 3; ++m) {
    float a = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 w = __ldg(w4 + m * 8 + j);   // the same address for the whole warp: a broadcast
      a += x[4 * j] * w.x + x[4 * j + 1] * w.y + x[4 * j + 2] * w.z + x[4 * j + 3] * w.w;
    }
    o[static_cast<size_t>(m) * S * S] = a;
  }
}

// ---- coalesced epilogue -----------------------------------------------------------------------------------------
// After tcgen05.ld a thread owns one output ROW (32 columns): storing that directly makes every warp instruction
// touch 32 different cache lines (measured: the epilogue, not the MMA, bounded the GEMM).  Instead each epilogue
// warp exchanges 32 rows x 64 B through a private, XOR-swizzled (bank-conflict-free) shared-memory tile so that
// afterwards lane l holds, for i = 0..3, the 16-byte piece (l & 3) of row 8*i + (l >> 2): a warp store instruction
// then writes 8 rows x 64 contiguous bytes (full 32-byte sectors).
// (explicit st.shared / ld.shared: through the generic `uint8_t*` the compiler emitted generic ST.E / LD.E, which take the
// slower generic-to-shared path of the LSU)
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void stage_exchange(uint8_t* st, int lane, const uint32_t (&in)[16], uint4 (&out)[4]) {
  __syncwarp();  // the previous exchange has been read by everyone
  const uint32_t base = smem_u32(st);
  const int sw = (lane >> 1) & 3;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    sts128(base + lane * 64 + ((j ^ sw) << 4), in[4 * j], in[4 * j + 1], in[4 * j + 2], in[4 * j + 3]);
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = 8 * i + (lane >> 2);
    out[i] = lds128(base + r * 64 + (((lane & 3) ^ ((r >> 1) & 3)) << 4));
  }
}

// One warp handles rows row0..row0+31 (lane = row) x columns col..col+31.  Every lane of the warp must call this.
// Residual prefetch (EPI_F32_RESID): the 8 float4 a lane adds AFTER the shared-memory exchange (rows row0 + 8i + lane/4,
// columns col + 16h + 4*(lane&3)) are requested one chunk ahead — the first chunk of a tile while the epilogue warps still
// wait for the accumulator — so the DRAM/L2 latency of the f32 residual stream hides behind the MMA main loop.
struct ResidPrefetch {
  float4 v[8];
  bool valid;
};
__device__ __forceinline__ bool resid_fast_ok(const EpiParams& ep) {
  return ep.resid != nullptr && (ep.ldo & 3) == 0 && (reinterpret_cast<uintptr_t>(ep.out) & 15) == 0 && (ep.ldr & 3) == 0 &&
         (reinterpret_cast<uintptr_t>(ep.resid) & 15) == 0;
}
__device__ __forceinline__ void resid_prefetch(const EpiParams& ep, int row0, int lane, int col, int M, int N, ResidPrefetch& pf) {
  pf.valid = (col + 32 <= N);
  if (!pf.valid) return;
  const int sub = lane & 3, rsub = lane >> 2;
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = row0 + 8 * i + rsub;
      if (r < M)
        pf.v[h * 4 + i] = *reinterpret_cast<const float4*>(ep.resid + static_cast<size_t>(ep.resid_mod > 0 ? r % ep.resid_mod : r) * ep.ldr + col + 16 * h + 4 * sub);
    }
}

template <int EPI>
__device__ __forceinline__ void epilogue_chunk(const EpiParams& ep, int row0, int lane, int col, const uint32_t (&v)[32],
                                               int M, int N, const float2* s_rope, uint8_t* st, const ResidPrefetch* pf = nullptr) {
  if (col >= N) return;  // warp uniform
  if constexpr (EPI == EPI_GELU_DOT) {
    epilogue_gelu_dot(ep, row0 + lane, col, v, M);
    return;
  }
  bool fast = (col + 32 <= N);
  if constexpr (EPI == EPI_F32 || EPI == EPI_F32_RESID)
    fast = fast && (ep.ldo & 3) == 0 && (reinterpret_cast<uintptr_t>(ep.out) & 15) == 0 &&
           (EPI != EPI_F32_RESID || ((ep.ldr & 3) == 0 && (reinterpret_cast<uintptr_t>(ep.resid) & 15) == 0));
  if constexpr (EPI == EPI_BF16 || EPI == EPI_BF16_GELU || EPI == EPI_BF16_RELU) fast = fast && (ep.ldo & 7) == 0 && (reinterpret_cast<uintptr_t>(ep.out) & 15) == 0;
  if constexpr (EPI == EPI_BF16_GELU) fast = fast && (ep.resid == nullptr || ((ep.ldr & 3) == 0 && (reinterpret_cast<uintptr_t>(ep.resid) & 15) == 0));
  if constexpr (EPI == EPI_PATCH) fast = fast && (ep.ldo & 3) == 0 && (reinterpret_cast<uintptr_t>(ep.out) & 15) == 0 && (reinterpret_cast<uintptr_t>(ep.pos) & 15) == 0;
  if (!fast) {  // ragged N / unaligned output (e.g. the [N_points, 20] query result): per-thread row stores
    epilogue_store<EPI>(ep, row0 + lane, col, v, M, N, s_rope);
    return;
  }
  float acc[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) acc[j] = __uint_as_float(v[j]);
  if (ep.bias != nullptr && !(ep.debug & 32)) {
    const float4* b4 = reinterpret_cast<const float4*>(ep.bias + col);  // col % 32 == 0 -> 128-byte aligned
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 t = __ldg(b4 + j);
      acc[4 * j] += t.x; acc[4 * j + 1] += t.y; acc[4 * j + 2] += t.z; acc[4 * j + 3] += t.w;
    }
  }
  const int sub = lane & 3, rsub = lane >> 2;

  if constexpr (EPI == EPI_PATCH) {
    // patch-embed rows -> token rows (skip the cls slot of every image) + absolute position embedding (pe.py:509-519)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      uint32_t in[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) in[j] = __float_as_uint(acc[16 * h + j]);
      uint4 o[4];
      stage_exchange(st, lane, in, o);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = row0 + 8 * i + rsub;
        if (r >= M) continue;
        const int c = col + 16 * h + 4 * sub;
        const int b = r / ep.patches, p = r - b * ep.patches;
        const float4 pe = __ldg(reinterpret_cast<const float4*>(ep.pos + static_cast<size_t>(1 + p) * ep.ldo + c));
        const float4 val = make_float4(__uint_as_float(o[i].x) + pe.x, __uint_as_float(o[i].y) + pe.y,
                                       __uint_as_float(o[i].z) + pe.z, __uint_as_float(o[i].w) + pe.w);
        *reinterpret_cast<float4*>(static_cast<float*>(ep.out) + (static_cast<size_t>(b) * (ep.patches + 1) + 1 + p) * ep.ldo + c) = val;
      }
    }
  } else if constexpr (EPI == EPI_F32 || EPI == EPI_F32_RESID) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {  // two halves of 16 f32 columns = 64 B per row
      uint32_t in[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) in[j] = __float_as_uint(acc[16 * h + j]);
      uint4 o[4];
      stage_exchange(st, lane, in, o);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = row0 + 8 * i + rsub;
        if (r >= M) continue;
        const int c = col + 16 * h + 4 * sub;
        float4 val = make_float4(__uint_as_float(o[i].x), __uint_as_float(o[i].y), __uint_as_float(o[i].z), __uint_as_float(o[i].w));
        if constexpr (EPI == EPI_F32_RESID) {
          const float4 rr = (pf != nullptr && pf->valid) ? pf->v[h * 4 + i]
                                                         : *reinterpret_cast<const float4*>(ep.resid + static_cast<size_t>(ep.resid_mod > 0 ? r % ep.resid_mod : r) * ep.ldr + c);
          val.x += rr.x; val.y += rr.y; val.z += rr.z; val.w += rr.w;
        }
        *reinterpret_cast<float4*>(static_cast<float*>(ep.out) + static_cast<size_t>(r) * ep.ldo + c) = val;
      }
    }
  } else if constexpr (EPI == EPI_BF16 || EPI == EPI_BF16_GELU || EPI == EPI_BF16_RELU) {
    if constexpr (EPI == EPI_BF16_GELU) {
      if (ep.resid != nullptr && row0 + lane < M) {   // lane = row: 128 contiguous bytes of its residual row
        const int rr = row0 + lane;
        const float4* r4 = reinterpret_cast<const float4*>(ep.resid + static_cast<size_t>(ep.resid_mod > 0 ? rr % ep.resid_mod : rr) * ep.ldr + col);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 t = __ldg(r4 + j);
          acc[4 * j] += t.x; acc[4 * j + 1] += t.y; acc[4 * j + 2] += t.z; acc[4 * j + 3] += t.w;
        }
      }
      if (!(ep.debug & 8)) {
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = gelu_erf(acc[j]);
      }
    }
    if constexpr (EPI == EPI_BF16_RELU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[j] = fmaxf(acc[j], 0.f);
    }
    uint32_t in[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) in[j] = pack_bf16(acc[2 * j], acc[2 * j + 1]);
    uint4 o[4];
    stage_exchange(st, lane, in, o);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = row0 + 8 * i + rsub;
      if (r < M && !(ep.debug & 16)) *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(ep.out) + static_cast<size_t>(r) * ep.ldo + col + 8 * sub) = o[i];
    }
  } else if constexpr (EPI == EPI_QKV) {
    const int which = col / ep.width;  // warp uniform: 0 q, 1 k, 2 v
    const int within = col - which * ep.width;
    const int head = within >> 6, d0 = within & 63;
    const int row = row0 + lane;
    if (ep.rope_tab != nullptr && which < 2) {
      // 2D RoPE with a cls token (rope.py:315-347): interleaved pairs (2i, 2i+1); pairs 0..15 of a head
      // (d0 == 0) rotate by (x+1)*theta_i, pairs 16..31 (d0 == 32) by (y+1)*theta_i, the cls token by 0.
      const int t = row % ep.seq;
      int r = 0;
      if (t > 0) r = (d0 == 0 ? (t - 1) % ep.rope_grid : (t - 1) / ep.rope_grid) + 1;
      const float2* tab = s_rope + r * 17;
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        const float2 cs = tab[j >> 1];
        const float a = acc[j], bb = acc[j + 1];
        acc[j] = a * cs.x - bb * cs.y;
        acc[j + 1] = bb * cs.x + a * cs.y;
      }
    }
    uint32_t in[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) in[j] = pack_bf16(acc[2 * j], acc[2 * j + 1]);
    uint4 o[4];
    stage_exchange(st, lane, in, o);
    __nv_bfloat16* base = which == 0 ? ep.q : (which == 1 ? ep.k : ep.vt);   // V is [b, head, seq_pad, 64] like K (MN-major B operand of P.V)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = row0 + 8 * i + rsub;
      if (r >= M) continue;
      const int b = r / ep.seq, t = r - b * ep.seq;
      *reinterpret_cast<uint4*>(base + ((static_cast<size_t>(b) * ep.heads + head) * ep.seq_pad + t) * 64 + d0 + 8 * sub) = o[i];
    }
  }
}

// EPI_UP_LN: one warp handles rows row0..row0+31 x the 64 columns of chunks `col` and `col`+32 (one sub-pixel's channels).
__device__ __forceinline__ void epilogue_up_ln(const EpiParams& ep, int row0, int lane, int col, const uint32_t (&v0)[32],
                                               const uint32_t (&v1)[32], int M, uint8_t* st) {
  const int row = row0 + lane;
  float x[64];
  {
    const int rr = row < M ? row : M - 1;
    const float4* b4 = reinterpret_cast<const float4*>(ep.bias + col);
    const float4* r4 = reinterpret_cast<const float4*>(ep.resid + static_cast<size_t>(ep.resid_mod > 0 ? rr % ep.resid_mod : rr) * ep.ldr + col);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float4 b = __ldg(b4 + j), r = __ldg(r4 + j);
      const uint32_t* src = j < 8 ? &v0[4 * j] : &v1[4 * (j - 8)];
      x[4 * j] = __uint_as_float(src[0]) + b.x + r.x; x[4 * j + 1] = __uint_as_float(src[1]) + b.y + r.y;
      x[4 * j + 2] = __uint_as_float(src[2]) + b.z + r.z; x[4 * j + 3] = __uint_as_float(src[3]) + b.w + r.w;
    }
  }
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < 64; ++j) sum += x[j];
  const float u = sum * (1.f / 64.f);
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < 64; ++j) { x[j] -= u; q += x[j] * x[j]; }
  const float r = 1.f / sqrtf(q * (1.f / 64.f) + 1e-6f);   // LayerNorm2d (sam2_utils.py:141-153)
  const int sub = lane & 3, rsub = lane >> 2;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    uint32_t in[16];
    const float4* w4 = reinterpret_cast<const float4*>(ep.ln_w + ((col + 32 * h) & 63));
    const float4* c4 = reinterpret_cast<const float4*>(ep.ln_b + ((col + 32 * h) & 63));
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 w = __ldg(w4 + j), c = __ldg(c4 + j);
      const float* xs = &x[32 * h + 4 * j];
      in[2 * j] = pack_bf16(gelu_erf(w.x * (xs[0] * r) + c.x), gelu_erf(w.y * (xs[1] * r) + c.y));
      in[2 * j + 1] = pack_bf16(gelu_erf(w.z * (xs[2] * r) + c.z), gelu_erf(w.w * (xs[3] * r) + c.w));
    }
    uint4 o[4];
    stage_exchange(st, lane, in, o);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int rw = row0 + 8 * i + rsub;
      if (rw < M) *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(ep.out) + static_cast<size_t>(rw) * ep.ldo + col + 32 * h + 8 * sub) = o[i];
    }
  }
}

// CS = thread-block cluster size along M.  The CS CTAs of a cluster work on CS consecutive M tiles of the SAME
// N tile; each loads 1/CS of the B tile and TMA-multicasts it to all of them, so L2 operand traffic per CTA and
// k-block drops from 16 KB + BN*128 B to 16 KB + BN*128/CS B (the kernel is otherwise L2-bandwidth bound).
// A smem stage may only be refilled when EVERY CTA of the cluster has consumed it: the MMA warp's
// tcgen05.commit arrives (multicast) on the `empty` barrier of all CTAs, which therefore counts CS arrivals.
template <int BN, int EPI, int CS>
__global__ void __launch_bounds__(kGemmThreads, 1)
    gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N,
                        int K, EpiParams ep) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + Cfg::kStages * Cfg::kBytesA;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + Cfg::kStages;
  uint64_t* tmem_full = bars + 2 * Cfg::kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float2* s_rope = reinterpret_cast<float2*>(smem + Cfg::kStages * Cfg::kStageBytes + 256);
  uint8_t* s_stage = reinterpret_cast<uint8_t*>(s_rope) + kRopeRowsMax * 17 * 8;

  griddep_launch();   // the next kernel of the stream may be scheduled as SMs free up (it waits for our completion itself)
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_m = (M + kBM - 1) / kBM;
  const int tiles_n = (N + BN - 1) / BN;
  const int groups_m = (tiles_m + CS - 1) / CS;      // cluster tiles along M
  const int num_ctiles = groups_m * tiles_n;
  const int num_kb = (K + kBK - 1) / kBK;
  const int crank = CS > 1 ? static_cast<int>(cluster_ctarank()) : 0;
  const int cluster_id = blockIdx.x / CS;
  const int num_clusters = gridDim.x / CS;
  constexpr uint16_t kMask = static_cast<uint16_t>((1u << CS) - 1);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < Cfg::kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], CS);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  if constexpr (EPI == EPI_QKV) {
    if (ep.rope_tab != nullptr)
      for (int i = threadIdx.x; i < (ep.rope_grid + 1) * 16; i += blockDim.x) s_rope[(i >> 4) * 17 + (i & 15)] = ep.rope_tab[i];
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CS > 1) cluster_sync_all();   // peers' barriers are initialised before anyone signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // PDL: everything above (barrier init, TMEM allocation, tensor-map prefetch, the constant RoPE table) overlapped the
  // previous kernel's tail; operands, residuals and outputs are only touched from here on
  griddep_wait();

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0, phase = 0;
      for (int ct = cluster_id; ct < num_ctiles; ct += num_clusters) {
        const int tm = (ct / tiles_n) * CS + crank, tn = ct % tiles_n;   // N fastest: concurrent CTAs spread over the weight tiles
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          if (ep.debug & 4) { mbar_arrive(&full[stage]); if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; } continue; }
          mbar_arrive_expect_tx(&full[stage], Cfg::kStageBytes);
          tma_load_2d(sA + stage * Cfg::kBytesA, &tmA, &full[stage], kb * kBK, tm * kBM);
          if constexpr (CS == 1) {
            tma_load_2d(sB + stage * Cfg::kBytesB, &tmB, &full[stage], kb * kBK, tn * BN);
          } else {
            constexpr int kSliceRows = BN / CS;
            tma_load_2d_multicast(sB + stage * Cfg::kBytesB + crank * kSliceRows * 128, &tmB, &full[stage], kb * kBK,
                                  tn * BN + crank * kSliceRows, kMask);
          }
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kBM, BN);
      int stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int ct = cluster_id; ct < num_ctiles; ct += num_clusters) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint64_t adesc = umma_desc_sw128(smem_u32(sA + stage * Cfg::kBytesA));
          const uint64_t bdesc = umma_desc_sw128(smem_u32(sB + stage * Cfg::kBytesB));
          if (!(ep.debug & 2)) {
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k)  // +32 B per K=16 step inside the 128 B swizzle row
              umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          }
          if constexpr (CS == 1) umma_commit(&empty[stage]);
          else umma_commit_multicast(&empty[stage], kMask);
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    const int quad = warp & 3;         // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;  // which of the two warps of that quadrant: takes chunks half, half+2, ...
    int acc = 0, acc_phase = 0;
    for (int ct = cluster_id; ct < num_ctiles; ct += num_clusters) {
      const int tm = (ct / tiles_n) * CS + crank, tn = ct % tiles_n;
      if constexpr (EPI == EPI_F32_RESID) {
        const bool pre_ok = resid_fast_ok(ep) && !(ep.debug & 64);
        ResidPrefetch nxt;
        nxt.valid = false;
        if (pre_ok) resid_prefetch(ep, tm * kBM + quad * 32, lane, tn * BN + half * 32, M, N, nxt);   // before the accumulator is ready
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
#pragma unroll 1
        for (int c = half; c < BN / 32; c += 2) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN + c * 32, v);
          ResidPrefetch cur = nxt;
          nxt.valid = false;
          if (pre_ok && c + 2 < BN / 32) resid_prefetch(ep, tm * kBM + quad * 32, lane, tn * BN + (c + 2) * 32, M, N, nxt);
          tmem_ld_wait();
          if (!(ep.debug & 1)) epilogue_chunk<EPI>(ep, tm * kBM + quad * 32, lane, tn * BN + c * 32, v, M, N, s_rope, s_stage + (warp - 2) * 2048, &cur);
        }
      } else if constexpr (EPI == EPI_UP_LN) {
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
#pragma unroll 1
        for (int c = 2 * half; c + 1 < BN / 32; c += 4) {   // adjacent chunk pairs = the 64 channels of one sub-pixel
          uint32_t v0[32], v1[32];
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN + c * 32, v0);
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN + (c + 1) * 32, v1);
          tmem_ld_wait();
          if (tn * BN + c * 32 + 64 <= N) epilogue_up_ln(ep, tm * kBM + quad * 32, lane, tn * BN + c * 32, v0, v1, M, s_stage + (warp - 2) * 2048);
        }
      } else {
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int c = half; c < BN / 32; c += 2) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN + c * 32, v);
        tmem_ld_wait();
        if (!(ep.debug & 1)) epilogue_chunk<EPI>(ep, tm * kBM + quad * 32, lane, tn * BN + c * 32, v, M, N, s_rope, s_stage + (warp - 2) * 2048);
      }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (CS > 1) cluster_sync_all();   // no CTA leaves while a peer may still multicast into it / signal it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------- host
int make_tmap_bf16_2d(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                      uint32_t box_rows, uint32_t box_cols);
int num_sms();

// Launches C = A.B^T with the given epilogue.  lda/ldb in elements (multiples of 8), pointers 16 B aligned.
int launch_gemm(int epi, const __nv_bfloat16* A, int lda, const __nv_bfloat16* B, int ldb, int M, int N, int K,
                const EpiParams& ep, cudaStream_t stream, int force_bn = 0);

}  // namespace ovo
