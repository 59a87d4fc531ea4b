// Error state, TMA descriptor creation, GEMM dispatch, and the GEMM test tap of the C ABI.
#include "common.cuh"
#include "gemm.cuh"

#include <cudaTypedefs.h>
#include <cstdlib>
#include <mutex>
#include <vector>

extern int g_attn_debug;

namespace ovo {

static thread_local char g_err[512] = "";
static thread_local long long g_launches = 0;

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
void count_launch(int n) { g_launches += n; }

struct ProfRec { int cls; cudaEvent_t a, b; double flops, bytes; };
static bool g_prof = false;
static std::vector<ProfRec> g_recs;
bool profiling() { return g_prof; }
ProfScope::ProfScope(cudaStream_t stream, int cls, double flops, double bytes) : s(stream), idx(-1) {
  if (!g_prof) return;
  ProfRec r{cls, nullptr, nullptr, flops, bytes};
  cudaEventCreate(&r.a); cudaEventCreate(&r.b);
  cudaEventRecord(r.a, s);
  g_recs.push_back(r);
  idx = static_cast<int>(g_recs.size()) - 1;
}
ProfScope::~ProfScope() { if (idx >= 0) cudaEventRecord(g_recs[idx].b, s); }

int num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

// cuTensorMapEncodeTiled is fetched through the runtime so the library has no link-time libcuda dependency
// (it must load on a machine without a driver for the symbol-export test).
static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  });
  return fn;
}

int make_tmap_bf16_2d(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                      uint32_t box_rows, uint32_t box_cols) {
  auto enc = get_encode();
  if (!enc) return set_error(OVO_E_CUDA, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (ld_elems & 7) != 0)
    return set_error(OVO_E_INVALID, "TMA operand must be 16-byte aligned with a row stride multiple of 8 elements");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld_elems * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(OVO_E_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return OVO_OK;
}

static bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("OVO_B200_PDL");
    v = (e && e[0] == '1') ? 1 : 0;   // opt-in: measured 0-1 % on graph-replayed ViT / SAM-2 trunks (launch gaps are not the limit)
  }
  return v == 1;
}

template <int BN, int EPI, int CS>
static int launch_one(const CUtensorMap& ta, const CUtensorMap& tb, int M, int N, int K, const EpiParams& ep,
                      cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  auto kern = gemm_bf16_tn_kernel<BN, EPI, CS>;
  static bool attr_set = false;
  if (!attr_set) {
    OVO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  const int ctiles = ceil_div(ceil_div(M, kBM), CS) * ceil_div(N, BN);   // cluster tiles
  const int max_clusters = num_sms() / CS;
  const int grid = (ctiles < max_clusters ? ctiles : max_clusters) * CS;
  ProfScope prof(stream, ep.prof_cls, 2.0 * M * N * K, 2.0 * (static_cast<double>(M) * K + static_cast<double>(N) * K) + 4.0 * M * N);
  {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kGemmThreads);
    cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if constexpr (CS > 1) {
      attr[na].id = cudaLaunchAttributeClusterDimension;
      attr[na].val.clusterDim.x = CS;
      attr[na].val.clusterDim.y = 1;
      attr[na].val.clusterDim.z = 1;
      ++na;
    }
    if (pdl_enabled()) {   // programmatic dependent launch: our prologue overlaps the previous kernel's tail (ptx.cuh)
      attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[na].val.programmaticStreamSerializationAllowed = 1;
      ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    OVO_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, M, N, K, ep));
  }
  OVO_CHECK_LAUNCH();
  return OVO_OK;
}

void keep_default_mempool_cached() {
  int dev = 0;
  cudaMemPool_t pool;
  if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    unsigned long long thr = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  }
  cudaGetLastError();
}

static int g_gemm_debug = 0;
static int g_query_cluster = 2;   // cluster size of the wide (Q > 64) dense query: the text tile is multicast (set by tuning)
static int g_gemm_cluster = -1;   // -1 = auto; OVO_B200_GEMM_CLUSTER=1|2|4 forces a cluster size (tuning aid)

template <int EPI>
static int launch_bn(int bn, const __nv_bfloat16* A, int lda, const __nv_bfloat16* B, int ldb, int M, int N, int K,
                     const EpiParams& ep, cudaStream_t stream) {
  if (g_gemm_cluster == -1) {
    const char* e = getenv("OVO_B200_GEMM_CLUSTER");
    g_gemm_cluster = e ? atoi(e) : 0;
  }
  // clusters pay off when there are enough M tiles to share a B tile; the narrow tiles (query, pooling) stay 1-CTA
  int cs = 1;
  // measured on B200: the GEMM was epilogue-bound, not L2-bound, and multicast brought nothing -> default 1
  if (bn >= 128 && ceil_div(M, kBM) >= 2) cs = g_gemm_cluster > 0 ? g_gemm_cluster : (ep.prof_cls == PROF_QUERY ? g_query_cluster : 1);
  if (bn < 128 || ceil_div(M, kBM) < cs) cs = 1;
  CUtensorMap ta, tb;
  OVO_TRY(make_tmap_bf16_2d(&ta, A, M, K, lda, kBM, kBK));
  OVO_TRY(make_tmap_bf16_2d(&tb, B, N, K, ldb, bn / cs, kBK));
  switch (bn * 10 + cs) {
    case 321: return launch_one<32, EPI, 1>(ta, tb, M, N, K, ep, stream);
    case 641: return launch_one<64, EPI, 1>(ta, tb, M, N, K, ep, stream);
    case 1281: return launch_one<128, EPI, 1>(ta, tb, M, N, K, ep, stream);
    case 1282: return launch_one<128, EPI, 2>(ta, tb, M, N, K, ep, stream);
    case 1284: return launch_one<128, EPI, 4>(ta, tb, M, N, K, ep, stream);
    case 2561: return launch_one<256, EPI, 1>(ta, tb, M, N, K, ep, stream);
    case 2562: return launch_one<256, EPI, 2>(ta, tb, M, N, K, ep, stream);
    case 2564: return launch_one<256, EPI, 4>(ta, tb, M, N, K, ep, stream);
  }
  return set_error(OVO_E_INVALID, "unsupported BN %d / cluster %d", bn, cs);
}

// Pick the N tile: the widest tile that does not waste more SM-rounds than a narrower one.
static int pick_bn(int M, int N) {
  if (N <= 32) return 32;
  if (N <= 64) return 64;
  if (N <= 128) return 128;
  const int sms = num_sms();
  const long long t256 = static_cast<long long>(ceil_div(M, kBM)) * ceil_div(N, 256);
  const long long t128 = static_cast<long long>(ceil_div(M, kBM)) * ceil_div(N, 128);
  const long long cost256 = 2LL * ceil_div(t256, sms);
  const long long cost128 = 1LL * ceil_div(t128, sms);
  return cost256 <= cost128 ? 256 : 128;
}

int launch_gemm(int epi, const __nv_bfloat16* A, int lda, const __nv_bfloat16* B, int ldb, int M, int N, int K,
                const EpiParams& ep_in, cudaStream_t stream, int force_bn) {
  EpiParams ep = ep_in;
  ep.debug = g_gemm_debug;
  if (M <= 0 || N <= 0 || K <= 0) return set_error(OVO_E_INVALID, "gemm: empty problem %dx%dx%d", M, N, K);
  const int bn = force_bn > 0 ? force_bn : pick_bn(M, N);
  switch (epi) {
    case EPI_F32: return launch_bn<EPI_F32>(bn, A, lda, B, ldb, M, N, K, ep, stream);
    case EPI_BF16: return launch_bn<EPI_BF16>(bn, A, lda, B, ldb, M, N, K, ep, stream);
    case EPI_BF16_GELU: return launch_bn<EPI_BF16_GELU>(bn, A, lda, B, ldb, M, N, K, ep, stream);
    case EPI_F32_RESID: return launch_bn<EPI_F32_RESID>(bn, A, lda, B, ldb, M, N, K, ep, stream);
    case EPI_QKV: return launch_bn<EPI_QKV>(bn, A, lda, B, ldb, M, N, K, ep, stream);
    case EPI_PATCH: return launch_bn<EPI_PATCH>(bn, A, lda, B, ldb, M, N, K, ep, stream);
    case EPI_BF16_RELU: return launch_bn<EPI_BF16_RELU>(bn, A, lda, B, ldb, M, N, K, ep, stream);
    case EPI_GELU_DOT: return launch_bn<EPI_GELU_DOT>(bn, A, lda, B, ldb, M, N, K, ep, stream);
    case EPI_UP_LN: return launch_bn<EPI_UP_LN>(bn >= 128 ? bn : 128, A, lda, B, ldb, M, N, K, ep, stream);
  }
  return set_error(OVO_E_INVALID, "unknown epilogue %d", epi);
}

}  // namespace ovo

extern "C" {

const char* ovo_last_error(void) { return ovo::g_err; }
int ovo_version(void) { return 100; }
long long ovo_launch_count(int reset) {
  long long v = ovo::g_launches;
  if (reset) ovo::g_launches = 0;
  return v;
}

void ovo_profile_begin(void) {
  for (auto& r : ovo::g_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  ovo::g_recs.clear();
  ovo::g_prof = true;
}

int ovo_profile_report(int n_classes, float* ms, double* flops, double* bytes, int* counts) {
  ovo::g_prof = false;
  if (n_classes < ovo::PROF_NCLASS) return ovo::set_error(OVO_E_INVALID, "ovo_profile_report: need room for %d classes", (int)ovo::PROF_NCLASS);
  for (int i = 0; i < n_classes; ++i) { ms[i] = 0.f; flops[i] = 0.0; bytes[i] = 0.0; counts[i] = 0; }
  for (auto& r : ovo::g_recs) {
    float t = 0.f;
    if (cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) {
      ms[r.cls] += t; flops[r.cls] += r.flops; bytes[r.cls] += r.bytes; counts[r.cls] += 1;
    }
    cudaEventDestroy(r.a); cudaEventDestroy(r.b);
  }
  ovo::g_recs.clear();
  cudaGetLastError();
  return ovo::PROF_NCLASS;
}

void ovo_set_gemm_cluster(int cluster_size) {
  ovo::g_gemm_cluster = cluster_size & 0xff;
  ovo::g_gemm_debug = (cluster_size >> 8) & 0xff;   // bits 8..15: tuning experiments (EpiParams::debug)
  if ((cluster_size >> 16) & 0xff) ovo::g_query_cluster = (cluster_size >> 16) & 0xff;   // bits 16..23: query cluster
  g_attn_debug = (cluster_size >> 24) & 0x7f;                                            // bits 24..30: attention experiments
}

int ovo_gemm_bf16(const void* A_dev, int lda, const void* B_dev, int ldb, int M, int N, int K, const float* bias_dev,
                  float* C_dev, int ldc, int force_bn, void* stream) {
  ovo::EpiParams ep;
  ep.out = C_dev;
  ep.ldo = ldc;
  ep.bias = bias_dev;
  return ovo::launch_gemm(ovo::EPI_F32, static_cast<const __nv_bfloat16*>(A_dev), lda,
                          static_cast<const __nv_bfloat16*>(B_dev), ldb, M, N, K, ep,
                          static_cast<cudaStream_t>(stream), force_bn);
}


// Tuning / measurement tap: times `iters` launches of C = A.B^T with the given fused epilogue on synthetic operands
// (0 f32, 1 bf16, 2 bf16+GELU, 3 f32+residual, 6 bf16+ReLU) with CUDA events on `stream`; *ms_out = average per launch.
int ovo_gemm_bench(int epi, int M, int N, int K, int force_bn, int iters, float* ms_out, void* stream_) {
  using namespace ovo;
  OVO_REQUIRE(M > 0 && N > 0 && K > 0 && K % 8 == 0 && iters > 0 && ms_out, "ovo_gemm_bench: bad arguments");
  OVO_REQUIRE(epi == EPI_F32 || epi == EPI_BF16 || epi == EPI_BF16_GELU || epi == EPI_F32_RESID || epi == EPI_BF16_RELU || epi == EPI_QKV, "ovo_gemm_bench: epilogue %d unsupported", epi);
  OVO_REQUIRE(epi != EPI_QKV || (M % 577 == 0 && N % 192 == 0), "ovo_gemm_bench: the QKV epilogue is timed at ViT shapes (M = images*577, N = 3*width, heads of 64)");
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  __nv_bfloat16 *A = nullptr, *B = nullptr; float *bias = nullptr, *resid = nullptr; void* out = nullptr;
  const int ldo = (N + 7) / 8 * 8;
  OVO_CUDA(cudaMalloc(&A, sizeof(__nv_bfloat16) * M * (size_t)K)); OVO_CUDA(cudaMalloc(&B, sizeof(__nv_bfloat16) * N * (size_t)K));
  OVO_CUDA(cudaMalloc(&bias, sizeof(float) * ldo)); OVO_CUDA(cudaMalloc(&resid, sizeof(float) * M * (size_t)ldo)); OVO_CUDA(cudaMalloc(&out, sizeof(float) * M * (size_t)ldo));
  OVO_CUDA(cudaMemsetAsync(A, 0x3c, sizeof(__nv_bfloat16) * M * (size_t)K, st)); OVO_CUDA(cudaMemsetAsync(B, 0x3c, sizeof(__nv_bfloat16) * N * (size_t)K, st));
  OVO_CUDA(cudaMemsetAsync(bias, 0, sizeof(float) * ldo, st)); OVO_CUDA(cudaMemsetAsync(resid, 0, sizeof(float) * M * (size_t)ldo, st));
  EpiParams ep;
  ep.out = out; ep.ldo = ldo; ep.bias = bias; ep.prof_cls = PROF_GEMM;
  if (epi == EPI_F32_RESID) { ep.resid = resid; ep.ldr = ldo; }
  __nv_bfloat16* qkv = nullptr;
  if (epi == EPI_QKV) {   // head-major q/k/v [images, heads, 640, 64], no RoPE table (the rotation math is not timed)
    const size_t per = static_cast<size_t>(M / 577) * (N / 192) * 640 * 64;
    OVO_CUDA(cudaMalloc(&qkv, sizeof(__nv_bfloat16) * 3 * per));
    ep.q = qkv; ep.k = qkv + per; ep.vt = qkv + 2 * per;
    ep.seq = 577; ep.seq_pad = 640; ep.heads = N / 192; ep.width = N / 3;
  }
  int rc = OVO_OK;
  for (int i = 0; i < 3 && rc == OVO_OK; ++i) rc = launch_gemm(epi, A, K, B, K, M, N, K, ep, st, force_bn);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, st);
  for (int i = 0; i < iters && rc == OVO_OK; ++i) rc = launch_gemm(epi, A, K, B, K, M, N, K, ep, st, force_bn);
  cudaEventRecord(e1, st);
  cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  *ms_out = ms / iters;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(A); cudaFree(B); cudaFree(bias); cudaFree(resid); cudaFree(out); cudaFree(qkv);
  return rc;
}

}  // extern "C"
