// Host launcher + C-ABI entry for the tcgen05 GEMM (gemm_sm100.cuh).
#include <stdarg.h>
#include "common.cuh"
#include "gemm_sm100.cuh"
#include "internal.h"
#include "../../include/aaclip_b200.h"

namespace {

template <int CG, int BN, int ACT, int OUT, int LNF = 0, int RV = 0>
int launch_one(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmD,
               const gemm::Args& args, int num_sms, cudaStream_t stream) {
  using C = gemm::Cfg<CG, OUT == gemm::OUT_F32_RESID_LN, RV, BN>;
  auto kern = gemm::gemm_kernel<CG, ACT, OUT, LNF, RV, BN>;
  // the opt-in to > 48 KB of dynamic shared memory belongs to the CURRENT device's context: remember it per
  // (instantiation, device), so that a second context on another GPU of the same process is configured too
  static bool configured[64] = {false};
  int dev = 0;
  AACLIP_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    AACLIP_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  const int tiles = ((args.M + C::BM * CG - 1) / (C::BM * CG)) * ((args.N + C::BN - 1) / C::BN);   // C::BN: 256 or 128
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
  AACLIP_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmC, tmD, args));
  return host::OK;
}

template <int CG, int BN>
int dispatch(int act, int out_mode, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC,
             const CUtensorMap& tmD, const gemm::Args& a, int sms, cudaStream_t s) {
  using namespace gemm;
  if (a.ln_part != nullptr) {   // LayerNorm folded into this (consumer) GEMM: bf16 outputs only
#define CASE_LN(A_) \
  if (act == A_ && out_mode == OUT_BF16) return launch_one<CG, BN, A_, OUT_BF16, 1>(tmA, tmB, tmC, tmD, a, sms, s);
    CASE_LN(ACT_NONE)
    CASE_LN(ACT_GELU_ERF)
    CASE_LN(ACT_QUICK_GELU)
#undef CASE_LN
    return host::fail(host::ERR_INVALID, "gemm: the folded-LayerNorm epilogue exists for bf16 outputs only (act=%d, out=%d)",
                      act, out_mode);
  }
#define CASE(A_, O_) \
  if (act == A_ && out_mode == O_) return launch_one<CG, BN, A_, O_>(tmA, tmB, tmC, tmD, a, sms, s);
  CASE(ACT_NONE, OUT_BF16)
  CASE(ACT_GELU_ERF, OUT_BF16)
  CASE(ACT_QUICK_GELU, OUT_BF16)
  CASE(ACT_NONE, OUT_F32_RESID)
  if (act == ACT_NONE && out_mode == OUT_F32_RESID_LN) {
    // AACLIP_RLN_VARIANT=1 selects the round-1 epilogue (one x_old chunk in flight per warp) for A/B runs; the default
    // (2) keeps a 3-deep x_old ring per warp.  Measured on B200 (tools/rln_probe.py, profiles/r2_rln_variants.txt):
    // out_proj 98.8 vs 96.8 us alone, 2.67 vs 2.61 ms per step in the bench - the epilogue is not what bounds this GEMM.
    static const int rv = getenv("AACLIP_RLN_VARIANT") ? atoi(getenv("AACLIP_RLN_VARIANT")) : 2;
    if constexpr (BN == 256) {
      if (rv == 1) return launch_one<CG, BN, ACT_NONE, OUT_F32_RESID_LN, 0, 1>(tmA, tmB, tmC, tmD, a, sms, s);
    }
    return launch_one<CG, BN, ACT_NONE, OUT_F32_RESID_LN, 0, 2>(tmA, tmB, tmC, tmD, a, sms, s);
  }
  CASE(ACT_NONE, OUT_F32)
  CASE(ACT_LEAKY, OUT_F32)
  CASE(ACT_LEAKY, OUT_BF16)
  CASE(ACT_NONE, OUT_F32_PATCH)
  CASE(ACT_NONE, OUT_DOTS)
  CASE(ACT_LEAKY, OUT_DOTS)
#undef CASE
  return host::fail(host::ERR_INVALID, "gemm: unsupported epilogue (act=%d, out=%d)", act, out_mode);
}

}  // namespace

int k::launch_gemm(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const float* bias, void* out,
                   int ldo, int act, int out_mode, const float* pos, int P, int cta_group, cudaStream_t stream,
                   const float* anchors, void* partials, int dots_cols, const k::LnFold* ln) {
  if (M <= 0 || N <= 0 || K <= 0) return host::fail(host::ERR_INVALID, "gemm: empty problem %dx%dx%d", M, N, K);
  if (N % 32 != 0) return host::fail(host::ERR_INVALID, "gemm: N=%d must be a multiple of 32", N);
  if (lda % 8 != 0 || ldw % 8 != 0 || lda < K || ldw < K)
    return host::fail(host::ERR_INVALID, "gemm: operand pitches (%d, %d) must be multiples of 8 and >= K=%d", lda, ldw, K);
  if (out_mode == gemm::OUT_F32_PATCH && (pos == nullptr || P <= 0 || M % P != 0))
    return host::fail(host::ERR_INVALID, "gemm: patch epilogue needs pos and M %% P == 0");
  if (cta_group < 0 || cta_group > 3)
    return host::fail(host::ERR_INVALID, "gemm: cta_group must be 0 (auto), 1, 2 or 3 (cta_group 1 with 128-column tiles)");
  if (out_mode == gemm::OUT_DOTS && (!anchors || !partials || dots_cols <= 0 || dots_cols % 128 != 0 || dots_cols > N ||
                                     out == nullptr))
    return host::fail(host::ERR_INVALID, "gemm: dots epilogue needs anchors, partials and dots_cols %% 128 == 0 (got %d)",
                      dots_cols);
  const bool rln = (out_mode == gemm::OUT_F32_RESID_LN);
  if (rln && (!ln || !ln->xb || !ln->part_out || N % 256 != 0 || ln->ldxb % 8 != 0 || ln->ldxb < N))
    return host::fail(host::ERR_INVALID, "gemm: the residual+statistics epilogue needs xb, part_out and N %% 256 == 0");
  if (cta_group == 3 && N % 128 != 0) return host::fail(host::ERR_INVALID, "gemm: 128-column tiles need N %% 128 == 0 (N=%d)", N);
  if (!rln && ln && ln->part_in && (!ln->colsum || ln->slices <= 0 || out_mode != gemm::OUT_BF16))
    return host::fail(host::ERR_INVALID, "gemm: the folded-LayerNorm epilogue needs colsum, slices and a bf16 output");
  int dev = 0;
  AACLIP_CUDA_CHECK(cudaGetDevice(&dev));
  const int sms = host::sm_count(dev) > 0 ? host::sm_count(dev) : 148;
  if (cta_group == 0) {
    // Tile shape by a wave model.  A CTA pair's 256 x 256 tile and a single CTA's 128 x 256 tile cost a SM the same
    // time (the pair form moves half the W bytes per SM: ~10 % faster per tile), a 128 x 128 tile roughly 0.6 of it;
    // the launch takes ceil(tiles / resident tiles) rounds.  Large batches always come out as the pair form; small
    // ones (tiles < SMs: every tile's K loop is a serial chain) as the shape that fills the machine in fewer rounds.
    auto rounds = [&](int bm, int bn, int slots) {
      const long long tiles = (long long)((M + bm - 1) / bm) * ((N + bn - 1) / bn);
      return double((tiles + slots - 1) / slots);
    };
    const double c2 = rounds(256, 256, sms / 2) * 1.0, c1 = rounds(128, 256, sms) * 1.1;
    const double c3 = (N % 128 == 0 && out_mode != gemm::OUT_F32_PATCH) ? rounds(128, 128, sms) * 0.6 : 1e30;
    cta_group = (c2 <= c1 && c2 <= c3) ? 2 : (c1 <= c3 ? 1 : 3);
  }
  CUtensorMap tmA, tmB;
  int rc = host::make_tmap_2d(&tmA, A, M, K, lda, 128);
  if (rc) return rc;
  rc = host::make_tmap_2d(&tmB, W, N, K, ldw, cta_group == 1 ? 256 : 128);   // rows of W one CTA stages per k-block
  if (rc) return rc;
  // output tiles leave through TMA: box = 32 rows x 128 B (64 bf16 / 32 fp32), 128B swizzle
  CUtensorMap tmC, tmD;
  memset(&tmC, 0, sizeof tmC);
  memset(&tmD, 0, sizeof tmD);
  if (out_mode != gemm::OUT_F32_PATCH && !(out_mode == gemm::OUT_DOTS && out == nullptr)) {
    const bool obf = (out_mode == gemm::OUT_BF16);
    if (ldo % (obf ? 8 : 4) != 0 || ldo < N) return host::fail(host::ERR_INVALID, "gemm: output pitch %d", ldo);
    rc = host::make_tmap_out(&tmC, out, M, N, ldo, obf);
    if (rc) return rc;
  }
  gemm::Args a;
  a.M = M; a.N = N; a.K = K; a.bias = bias; a.out = out; a.ldo = ldo; a.pos = pos; a.P = P;
  a.anchors = anchors; a.partials = static_cast<float4*>(partials); a.dots_cols = dots_cols;
  a.ln_part = nullptr; a.ln_slices = 0; a.ln_width = K; a.ln_eps = 0.f; a.ln_colsum = nullptr; a.part_out = nullptr;
  a.xb = nullptr; a.ldxb = 0;
  if (rln) {
    if ((reinterpret_cast<uintptr_t>(ln->xb) & 15u) != 0) return host::fail(host::ERR_INVALID, "gemm: xb must be 16-byte aligned");
    a.part_out = static_cast<float2*>(ln->part_out);
    a.xb = static_cast<__nv_bfloat16*>(ln->xb); a.ldxb = ln->ldxb;
    rc = host::make_tmap_out(&tmD, ln->xb, M, N, ln->ldxb, true);   // variants that stage the bf16 copy for a TMA store
    if (rc) return rc;
  } else if (ln && ln->part_in) {
    a.ln_part = static_cast<const float2*>(ln->part_in);
    a.ln_slices = ln->slices; a.ln_eps = ln->eps; a.ln_colsum = ln->colsum;
  }
  return cta_group == 1   ? dispatch<1, 256>(act, out_mode, tmA, tmB, tmC, tmD, a, sms, stream)
         : cta_group == 3 ? dispatch<1, 128>(act, out_mode, tmA, tmB, tmC, tmD, a, sms, stream)
                          : dispatch<2, 256>(act, out_mode, tmA, tmB, tmC, tmD, a, sms, stream);
}

extern "C" int aaclip_gemm_bf16(const void* A, int lda, const void* W, int ldw, int M, int N, int K,
                                const float* bias, void* out, int ldo, int act, int out_mode, const float* pos, int P,
                                int cta_group, void* stream) {
  host::PointerDeviceGuard dev_guard(A);
  if (out_mode == gemm::OUT_DOTS) return host::fail(host::ERR_INVALID, "gemm: the dots epilogue is internal to the engine");
  return k::launch_gemm(A, lda, W, ldw, M, N, K, bias, out, ldo, act, out_mode, pos, P, cta_group,
                        static_cast<cudaStream_t>(stream));
}

// Building blocks of the folded-LayerNorm schedule, exported so the parity tests can pin them on their own:
//   aaclip_gemm_resid_ln  x <- x + A W^T + bias (fp32, in place); xb <- bf16(x); part[r][N/128] <- (sum, sum sq) per slice
//   aaclip_gemm_lnfold    out(bf16) <- act(rstd_r (A Wf^T - mean_r colsum) + bias) with the row statistics from part
extern "C" int aaclip_gemm_resid_ln(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const float* bias,
                                    float* x, int ldx, void* xb, int ldxb, void* part_out, int cta_group, void* stream) {
  host::PointerDeviceGuard dev_guard(A);
  k::LnFold ln;
  ln.xb = xb; ln.ldxb = ldxb; ln.part_out = part_out;
  return k::launch_gemm(A, lda, W, ldw, M, N, K, bias, x, ldx, gemm::ACT_NONE, gemm::OUT_F32_RESID_LN, nullptr, 0, cta_group,
                        static_cast<cudaStream_t>(stream), nullptr, nullptr, 0, &ln);
}
extern "C" int aaclip_gemm_lnfold(const void* A, int lda, const void* Wf, int ldw, int M, int N, int K, const float* bias,
                                  const float* colsum, const void* part, int slices, float eps, void* out, int ldo, int act,
                                  int cta_group, void* stream) {
  host::PointerDeviceGuard dev_guard(A);
  k::LnFold ln;
  ln.part_in = part; ln.slices = slices; ln.eps = eps; ln.colsum = colsum;
  return k::launch_gemm(A, lda, Wf, ldw, M, N, K, bias, out, ldo, act, gemm::OUT_BF16, nullptr, 0, cta_group,
                        static_cast<cudaStream_t>(stream), nullptr, nullptr, 0, &ln);
}
