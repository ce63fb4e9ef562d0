// Device-side image transform: the reference's loader-side Compose (dataset/__init__.py:127-136)
//
//     transforms.Resize((S, S), Image.BICUBIC)   PIL resample on uint8 (antialiased when shrinking)
//     transforms.ToTensor()                      HWC uint8 -> CHW float32 / 255
//     transforms.Normalize(mean, std)            (x - mean) / std
//
// BIT-EXACT with PIL + torchvision: the resample is Pillow's integer arithmetic (Resample.c: Keys cubic a = -0.5,
// support 2 * max(scale, 1), per-output-pixel window, coefficients normalised in double and rounded to 22 fractional
// bits, pixel = clip8((2^21 + sum p * k) >> 22), horizontal pass first, uint8 intermediate); the float part uses
// IEEE division / subtraction (no reciprocal, no FMA contraction).  Coefficient tables are built on the host exactly
// as precompute_coeffs / normalize_coeffs_8bpc do, cached per (in, out) size pair and kept in device memory.
//
// Both kernels are HBM/L2-bound byte work (no tensor cores): per image 3*H0*W0 B read, 3*H0*S B intermediate written
// and re-read (L2), 12*S*S B written.
//   resample_h_kernel   4 input rows per CTA, rows staged in shared memory as RGBX words (one 32-bit LDS per tap
//                       yields all three channels; thread stride ~scale words -> conflict-free for odd strides)
//   resample_v_kernel   one output row per CTA: thread = byte column (x*3+c) so every tap is one coalesced byte row
//                       read; results are normalised into a planar smem tile and written as full float rows
#include <limits.h>
#include <math.h>
#include <stdarg.h>
#include <string.h>
#include <algorithm>
#include <map>
#include <mutex>
#include <vector>
#include "common.cuh"
#include "internal.h"
#include "../../include/aaclip_b200.h"

namespace {

constexpr int PRECISION_BITS = 32 - 8 - 2;
constexpr int H_ROWS = 8;       // input rows per CTA in the horizontal pass
constexpr int H_THREADS = 256;
constexpr int V_THREADS = 256;

struct Table {          // device copies
  int ksize = 0;
  int* bounds = nullptr;  // [out][2] = (first input index, tap count)
  int* kk = nullptr;      // [out][ksize] fixed-point coefficients
};

double bicubic_filter(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}

// Pillow Resample.c:precompute_coeffs + normalize_coeffs_8bpc for the whole-image box; a pass whose sizes agree is
// skipped by Pillow, which equals a single tap of weight 1.0 (clip8((2^21 + p * 2^22) >> 22) == p).
void build_table(int in_size, int out_size, int* ksize_out, std::vector<int>* bounds, std::vector<int>* kk) {
  if (in_size == out_size) {
    *ksize_out = 1;
    bounds->resize(2 * (size_t)out_size);
    kk->assign((size_t)out_size, 1 << PRECISION_BITS);
    for (int i = 0; i < out_size; ++i) { (*bounds)[2 * i] = i; (*bounds)[2 * i + 1] = 1; }
    return;
  }
  double scale = (double)in_size / out_size, filterscale = scale;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = 2.0 * filterscale;
  const int ksize = (int)ceil(support) * 2 + 1;
  *ksize_out = ksize;
  bounds->assign(2 * (size_t)out_size, 0);
  kk->assign((size_t)out_size * ksize, 0);
  std::vector<double> k(ksize);
  const double ss = 1.0 / filterscale;
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale;
    double ww = 0.0;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    for (int x = 0; x < xmax; ++x) {
      const double w = bicubic_filter((x + xmin - center + 0.5) * ss);
      k[x] = w;
      ww += w;
    }
    for (int x = 0; x < xmax; ++x) {
      if (ww != 0.0) k[x] /= ww;
      (*kk)[(size_t)xx * ksize + x] =
          k[x] < 0 ? (int)(-0.5 + k[x] * (1 << PRECISION_BITS)) : (int)(0.5 + k[x] * (1 << PRECISION_BITS));
    }
    (*bounds)[2 * xx] = xmin;
    (*bounds)[2 * xx + 1] = xmax;
  }
}

std::mutex g_mu;
std::map<long long, Table> g_tables;   // key: device << 48 | in << 24 | out

int get_table(int device, int in_size, int out_size, Table* out) {
  std::lock_guard<std::mutex> lock(g_mu);
  const long long key = ((long long)device << 48) | ((long long)in_size << 24) | (long long)out_size;
  auto it = g_tables.find(key);
  if (it != g_tables.end()) { *out = it->second; return host::OK; }
  std::vector<int> b, k;
  Table t;
  build_table(in_size, out_size, &t.ksize, &b, &k);
  AACLIP_CUDA_CHECK(cudaMalloc(&t.bounds, b.size() * sizeof(int)));
  AACLIP_CUDA_CHECK(cudaMalloc(&t.kk, k.size() * sizeof(int)));
  // synchronous copies (first use of a size pair only): the tables are read-only afterwards, on any stream
  AACLIP_CUDA_CHECK(cudaMemcpy(t.bounds, b.data(), b.size() * sizeof(int), cudaMemcpyHostToDevice));
  AACLIP_CUDA_CHECK(cudaMemcpy(t.kk, k.data(), k.size() * sizeof(int), cudaMemcpyHostToDevice));
  g_tables[key] = t;
  *out = t;
  return host::OK;
}

// ToTensor + Normalize as a 3 x 256 table: out = (float(v) / 255 - mean[c]) / std[c], evaluated on the host in IEEE
// single precision in torch's order (x86-64 SSE float arithmetic == the GPU's __fdiv_rn / __fsub_rn), cached per
// (device, mean, std) in device memory.
std::map<std::vector<uint32_t>, float*> g_luts;
int get_lut(int device, const float* mean, const float* stdv, const float** out) {
  std::lock_guard<std::mutex> lock(g_mu);
  std::vector<uint32_t> key(7);
  key[0] = (uint32_t)device;
  memcpy(&key[1], mean, 12);
  memcpy(&key[4], stdv, 12);
  auto it = g_luts.find(key);
  if (it != g_luts.end()) { *out = it->second; return host::OK; }
  std::vector<float> lut(3 * 256);
  for (int c = 0; c < 3; ++c)
    for (int v = 0; v < 256; ++v) {
      volatile float t = (float)v / 255.0f;   // volatile: one rounding per operation, no contraction
      volatile float u = t - mean[c];
      lut[c * 256 + v] = u / stdv[c];
    }
  float* d = nullptr;
  AACLIP_CUDA_CHECK(cudaMalloc(&d, lut.size() * sizeof(float)));
  AACLIP_CUDA_CHECK(cudaMemcpy(d, lut.data(), lut.size() * sizeof(float), cudaMemcpyHostToDevice));
  g_luts[key] = d;
  *out = d;
  return host::OK;
}

__device__ __forceinline__ uint32_t clip8(int v) {
  v >>= PRECISION_BITS;   // arithmetic shift, as Pillow's lookup index
  return (uint32_t)min(max(v, 0), 255);
}

// in u8 [rows, W0, 3] -> out u8 [rows, S, 3].  One CTA = H_ROWS input rows; a thread owns output column x for all of
// them: per tap ONE coefficient load feeds H_ROWS x 3 multiply-adds (taps outer, rows inner: 24 independent
// accumulators), and the bytes of a row reach shared memory through coalesced 32-bit loads.
__global__ void __launch_bounds__(H_THREADS)
resample_h_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, long long rows, int W0, int S,
                  const int* __restrict__ bounds, const int* __restrict__ kk, int ksize) {
  extern __shared__ uint32_t px[];   // [H_ROWS][W0] RGBX
  const long long r0 = (long long)blockIdx.x * H_ROWS;
  const int nr = (int)min((long long)H_ROWS, rows - r0);
  const uint8_t* src = in + r0 * W0 * 3;
  const int npx = nr * W0;
  if ((reinterpret_cast<uintptr_t>(src) & 3u) == 0) {
    // 4 pixels = 3 words: coalesced word loads, repacked to RGBX
    const uint32_t* w = reinterpret_cast<const uint32_t*>(src);
    const int nq = npx >> 2;
    for (int q = threadIdx.x; q < nq; q += (int)blockDim.x) {
      const uint32_t a = __ldg(w + 3 * q), b = __ldg(w + 3 * q + 1), c = __ldg(w + 3 * q + 2);
      px[4 * q + 0] = a & 0xFFFFFFu;
      px[4 * q + 1] = (a >> 24) | ((b & 0xFFFFu) << 8);
      px[4 * q + 2] = (b >> 16) | ((c & 0xFFu) << 16);
      px[4 * q + 3] = c >> 8;
    }
    for (int i = (nq << 2) + threadIdx.x; i < npx; i += (int)blockDim.x) {
      const uint8_t* p = src + 3LL * i;
      px[i] = (uint32_t)__ldg(p) | ((uint32_t)__ldg(p + 1) << 8) | ((uint32_t)__ldg(p + 2) << 16);
    }
  } else {
    for (int i = threadIdx.x; i < npx; i += (int)blockDim.x) {
      const uint8_t* p = src + 3LL * i;
      px[i] = (uint32_t)__ldg(p) | ((uint32_t)__ldg(p + 1) << 8) | ((uint32_t)__ldg(p + 2) << 16);
    }
  }
  __syncthreads();
  uint8_t* dst = out + r0 * S * 3;
  for (int x = threadIdx.x; x < S; x += (int)blockDim.x) {
    const int xmin = __ldg(bounds + 2 * x), cnt = __ldg(bounds + 2 * x + 1);
    const int* k = kk + (long long)x * ksize;
    const uint32_t* col = px + xmin;
    int acc[H_ROWS][3];
#pragma unroll
    for (int r = 0; r < H_ROWS; ++r) acc[r][0] = acc[r][1] = acc[r][2] = 1 << (PRECISION_BITS - 1);
    for (int t = 0; t < cnt; ++t) {
      const int c = __ldg(k + t);
#pragma unroll
      for (int r = 0; r < H_ROWS; ++r) {
        // rows past nr read stale smem of this CTA (in bounds); their results are never stored
        const uint32_t p = col[r * W0 + t];
        acc[r][0] += (int)(p & 255u) * c;
        acc[r][1] += (int)((p >> 8) & 255u) * c;
        acc[r][2] += (int)(p >> 16) * c;
      }
    }
#pragma unroll
    for (int r = 0; r < H_ROWS; ++r) {
      if (r < nr) {
        uint8_t* d = dst + 3LL * (r * S + x);
        d[0] = (uint8_t)clip8(acc[r][0]);
        d[1] = (uint8_t)clip8(acc[r][1]);
        d[2] = (uint8_t)clip8(acc[r][2]);
      }
    }
  }
}

// in u8 [B, H0, S, 3] -> out fp32 [B, 3, S, S] (or u8 [B, S, S, 3] when out_u8 != null: the bare PIL resize).
// One CTA = one output row: a thread owns 4 adjacent byte columns (x*3+c), so every tap is one coalesced 32-bit row
// read (byte loads when the row pitch 3*S is not a multiple of 4); the row's coefficients sit in shared memory.
__global__ void __launch_bounds__(V_THREADS)
resample_v_kernel(const uint8_t* __restrict__ in, float* __restrict__ out, uint8_t* __restrict__ out_u8, int H0, int S,
                  const int* __restrict__ bounds, const int* __restrict__ kk, int ksize, const float* __restrict__ lut) {
  extern __shared__ float plane[];   // [3][S] floats, then ksize coefficients
  int* coef = reinterpret_cast<int*>(plane + 3 * S);
  const int y = blockIdx.x, b = blockIdx.y;
  const int ymin = __ldg(bounds + 2 * y), cnt = __ldg(bounds + 2 * y + 1);
  for (int t = threadIdx.x; t < cnt; t += V_THREADS) coef[t] = __ldg(kk + (long long)y * ksize + t);
  __syncthreads();
  const int W3 = S * 3;
  const uint8_t* src = in + ((long long)b * H0 + ymin) * W3;
  auto emit = [&](int j, int s) {
    const uint32_t v = clip8(s);
    if (out_u8) {
      out_u8[((long long)b * S + y) * W3 + j] = (uint8_t)v;
    } else {
      const int x = j / 3, c = j - 3 * x;
      plane[c * S + x] = __ldg(lut + c * 256 + v);   // ToTensor + Normalize (see get_lut)
    }
  };
  if ((W3 & 3) == 0 && (reinterpret_cast<uintptr_t>(in) & 3u) == 0) {
    const int W4 = W3 >> 2;
    const uint32_t* src4 = reinterpret_cast<const uint32_t*>(src);
    for (int q = threadIdx.x; q < W4; q += V_THREADS) {
      int s0 = 1 << (PRECISION_BITS - 1), s1 = s0, s2 = s0, s3 = s0;
#pragma unroll 2
      for (int t = 0; t < cnt; ++t) {
        const uint32_t p = __ldg(src4 + (long long)t * W4 + q);
        const int c = coef[t];
        s0 += (int)(p & 255u) * c;
        s1 += (int)((p >> 8) & 255u) * c;
        s2 += (int)((p >> 16) & 255u) * c;
        s3 += (int)(p >> 24) * c;
      }
      emit(4 * q, s0); emit(4 * q + 1, s1); emit(4 * q + 2, s2); emit(4 * q + 3, s3);
    }
  } else {
    for (int j = threadIdx.x; j < W3; j += V_THREADS) {
      int s = 1 << (PRECISION_BITS - 1);
      for (int t = 0; t < cnt; ++t) s += (int)__ldg(src + (long long)t * W3 + j) * coef[t];
      emit(j, s);
    }
  }
  if (out_u8) return;
  __syncthreads();
  for (int j = threadIdx.x; j < W3; j += V_THREADS) {
    const int c = j / S, x = j - c * S;
    out[(((long long)b * 3 + c) * S + y) * S + x] = plane[j];
  }
}

int run(const uint8_t* images, int B, int H0, int W0, int S, const float* mean, const float* stdv, uint8_t* scratch,
        float* out, uint8_t* out_u8, cudaStream_t st) {
  if (B <= 0) return host::OK;
  if (!images || (!out && !out_u8)) return host::fail(host::ERR_INVALID, "preprocess: null argument");
  if (H0 < 1 || W0 < 1 || S < 1 || H0 >= (1 << 24) || W0 >= (1 << 24) || S >= (1 << 16))
    return host::fail(host::ERR_INVALID, "preprocess: sizes H0=%d W0=%d S=%d", H0, W0, S);
  const size_t smem_h = (size_t)H_ROWS * W0 * sizeof(uint32_t);
  if (smem_h > 200 * 1024) return host::fail(host::ERR_INVALID, "preprocess: W0=%d too wide (max 6400)", W0);
  if (W0 != S && !scratch) return host::fail(host::ERR_INVALID, "preprocess: scratch of B*H0*S*3 bytes required");
  int dev = 0;
  AACLIP_CUDA_CHECK(cudaGetDevice(&dev));
  Table th, tv;
  int rc = get_table(dev, W0, S, &th);
  if (rc) return rc;
  rc = get_table(dev, H0, S, &tv);
  if (rc) return rc;
  const uint8_t* vin = images;
  if (W0 != S) {   // Pillow skips a pass whose size does not change
    if (smem_h > 48 * 1024)
      AACLIP_CUDA_CHECK(cudaFuncSetAttribute(resample_h_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_h));
    const long long rows = (long long)B * H0;
    const long long grid = (rows + H_ROWS - 1) / H_ROWS;
    if (grid > INT_MAX) return host::fail(host::ERR_INVALID, "preprocess: %lld rows", rows);
    // block size: the S output columns split evenly over ceil(S / 256) passes (336 -> 2 passes of 168 -> 192 threads)
    const int passes = (S + H_THREADS - 1) / H_THREADS;
    const int threads = std::min(H_THREADS, (((S + passes - 1) / passes) + 31) / 32 * 32);
    resample_h_kernel<<<(unsigned)grid, threads, smem_h, st>>>(images, scratch, rows, W0, S, th.bounds, th.kk, th.ksize);
    AACLIP_CUDA_CHECK(cudaGetLastError());
    vin = scratch;
  }
  static const float CLIP_MEAN[3] = {0.48145466f, 0.4578275f, 0.40821073f};   // dataset/__init__.py:130-133
  static const float CLIP_STD[3] = {0.26862954f, 0.26130258f, 0.27577711f};
  const float* m = mean ? mean : CLIP_MEAN;
  const float* d = stdv ? stdv : CLIP_STD;
  const float* lut = nullptr;
  if (!out_u8) { rc = get_lut(dev, m, d, &lut); if (rc) return rc; }
  if (B > 65535) return host::fail(host::ERR_INVALID, "preprocess: B=%d > 65535", B);
  resample_v_kernel<<<dim3(S, B), V_THREADS, 3 * (size_t)S * sizeof(float) + (size_t)tv.ksize * sizeof(int), st>>>(vin, out, out_u8, H0, S, tv.bounds, tv.kk,
                                                                                  tv.ksize, lut);
  AACLIP_CUDA_CHECK(cudaGetLastError());
  return host::OK;
}

}  // namespace

int k::launch_preprocess_u8(const uint8_t* images, int B, int H0, int W0, int S, const float* mean, const float* stdv,
                            uint8_t* scratch, float* out, cudaStream_t stream) {
  return run(images, B, H0, W0, S, mean, stdv, scratch, out, nullptr, stream);
}

extern "C" long long aaclip_preprocess_scratch_bytes(int B, int H0, int W0, int S) {
  if (B <= 0 || H0 <= 0 || S <= 0 || W0 == S) return 0;
  return 3LL * B * H0 * S;
}

extern "C" int aaclip_preprocess_u8(const uint8_t* images, int B, int H0, int W0, int S, const float* host_mean,
                                    const float* host_std, uint8_t* scratch, float* out, void* stream) {
  host::PointerDeviceGuard dev_guard(images);
  return run(images, B, H0, W0, S, host_mean, host_std, scratch, out, nullptr, static_cast<cudaStream_t>(stream));
}

extern "C" int aaclip_resize_bicubic_u8(const uint8_t* images, int B, int H0, int W0, int S, uint8_t* scratch,
                                        uint8_t* out_u8, void* stream) {
  host::PointerDeviceGuard dev_guard(images);
  return run(images, B, H0, W0, S, nullptr, nullptr, scratch, nullptr, out_u8, static_cast<cudaStream_t>(stream));
}
