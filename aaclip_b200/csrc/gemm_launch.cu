// Host launcher + C-ABI entry for the tcgen05 GEMM (gemm_sm100.cuh).
#include <stdarg.h>
#include "common.cuh"
#include "gemm_sm100.cuh"
#include "internal.h"
#include "../../include/aaclip_b200.h"

namespace {

template <int CG, int ACT, int OUT>
int launch_one(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const gemm::Args& args, int num_sms,
               cudaStream_t stream) {
  using C = gemm::Cfg<CG>;
  auto kern = gemm::gemm_kernel<CG, ACT, OUT>;
  static bool configured = false;  // per instantiation
  if (!configured) {
    AACLIP_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    configured = true;
  }
  const int tiles = ((args.M + C::BM * CG - 1) / (C::BM * CG)) * ((args.N + C::BN - 1) / C::BN);
  int clusters = num_sms / CG;
  if (clusters > tiles) clusters = tiles;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(clusters * CG);
  cfg.blockDim = dim3(C::THREADS);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = host::pdl_enabled() ? 2 : 1;
  AACLIP_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmC, args));
  return host::OK;
}

template <int CG>
int dispatch(int act, int out_mode, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC,
             const gemm::Args& a, int sms,
             cudaStream_t s) {
  using namespace gemm;
#define CASE(A_, O_) \
  if (act == A_ && out_mode == O_) return launch_one<CG, A_, O_>(tmA, tmB, tmC, a, sms, s);
  CASE(ACT_NONE, OUT_BF16)
  CASE(ACT_GELU_ERF, OUT_BF16)
  CASE(ACT_QUICK_GELU, OUT_BF16)
  CASE(ACT_NONE, OUT_F32_RESID)
  CASE(ACT_NONE, OUT_F32)
  CASE(ACT_LEAKY, OUT_F32)
  CASE(ACT_NONE, OUT_F32_PATCH)
  CASE(ACT_NONE, OUT_DOTS)
  CASE(ACT_LEAKY, OUT_DOTS)
#undef CASE
  return host::fail(host::ERR_INVALID, "gemm: unsupported epilogue (act=%d, out=%d)", act, out_mode);
}

}  // namespace

int k::launch_gemm(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const float* bias, void* out,
                   int ldo, int act, int out_mode, const float* pos, int P, int cta_group, cudaStream_t stream,
                   const float* anchors, void* partials, int dots_cols) {
  if (M <= 0 || N <= 0 || K <= 0) return host::fail(host::ERR_INVALID, "gemm: empty problem %dx%dx%d", M, N, K);
  if (N % 32 != 0) return host::fail(host::ERR_INVALID, "gemm: N=%d must be a multiple of 32", N);
  if (lda % 8 != 0 || ldw % 8 != 0 || lda < K || ldw < K)
    return host::fail(host::ERR_INVALID, "gemm: operand pitches (%d, %d) must be multiples of 8 and >= K=%d", lda, ldw, K);
  if (out_mode == gemm::OUT_F32_PATCH && (pos == nullptr || P <= 0 || M % P != 0))
    return host::fail(host::ERR_INVALID, "gemm: patch epilogue needs pos and M %% P == 0");
  if (cta_group != 1 && cta_group != 2) return host::fail(host::ERR_INVALID, "gemm: cta_group must be 1 or 2");
  if (out_mode == gemm::OUT_DOTS && (!anchors || !partials || dots_cols <= 0 || dots_cols % 128 != 0 || dots_cols > N ||
                                     out == nullptr))
    return host::fail(host::ERR_INVALID, "gemm: dots epilogue needs anchors, partials and dots_cols %% 128 == 0 (got %d)",
                      dots_cols);
  int dev = 0;
  AACLIP_CUDA_CHECK(cudaGetDevice(&dev));
  const int sms = host::sm_count(dev);
  CUtensorMap tmA, tmB;
  int rc = host::make_tmap_2d(&tmA, A, M, K, lda, 128);
  if (rc) return rc;
  rc = host::make_tmap_2d(&tmB, W, N, K, ldw, cta_group == 1 ? 256 : 128);
  if (rc) return rc;
  // output tiles leave through TMA: box = 32 rows x 128 B (64 bf16 / 32 fp32), 128B swizzle
  CUtensorMap tmC;
  memset(&tmC, 0, sizeof tmC);
  if (out_mode != gemm::OUT_F32_PATCH && !(out_mode == gemm::OUT_DOTS && out == nullptr)) {
    const bool obf = (out_mode == gemm::OUT_BF16);
    if (ldo % (obf ? 8 : 4) != 0 || ldo < N) return host::fail(host::ERR_INVALID, "gemm: output pitch %d", ldo);
    rc = host::make_tmap_out(&tmC, out, M, N, ldo, obf);
    if (rc) return rc;
  }
  gemm::Args a;
  a.M = M; a.N = N; a.K = K; a.bias = bias; a.out = out; a.ldo = ldo; a.pos = pos; a.P = P;
  a.anchors = anchors; a.partials = static_cast<float4*>(partials); a.dots_cols = dots_cols;
  return cta_group == 1 ? dispatch<1>(act, out_mode, tmA, tmB, tmC, a, sms, stream)
                        : dispatch<2>(act, out_mode, tmA, tmB, tmC, a, sms, stream);
}

extern "C" int aaclip_gemm_bf16(const void* A, int lda, const void* W, int ldw, int M, int N, int K,
                                const float* bias, void* out, int ldo, int act, int out_mode, const float* pos, int P,
                                int cta_group, void* stream) {
  if (out_mode == gemm::OUT_DOTS) return host::fail(host::ERR_INVALID, "gemm: the dots epilogue is internal to the engine");
  return k::launch_gemm(A, lda, W, ldw, M, N, K, bias, out, ldo, act, out_mode, pos, P, cta_group,
                        static_cast<cudaStream_t>(stream));
}
