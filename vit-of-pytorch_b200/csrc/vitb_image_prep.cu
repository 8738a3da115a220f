// vitb_image_prep.cu — the input transform in front of the encoder, on the device.
//
// Replaces the per-sample CPU work of the reference's loaders (src/data_loaders.py:36-48,69-82,102-114):
//   Resize (Pillow bilinear, 8 bits per channel) -> RandomHorizontalFlip -> ToTensor -> Normalize
// and, optionally, the patch extraction of the Conv2d patch embedding (src/model.py:179,197), so a CIFAR
// batch crosses PCIe as 3 KB per image instead of 602 KB and the fp32 NCHW image need never exist.
//
// Integer work, HBM-bound on its outputs: per image 3*S*S fp32 (or bf16 patch-operand) bytes written against
// H*W*3 bytes read.  One CTA per (image, band of output rows): the band's source rows go through the horizontal
// pass into shared memory once (planar bytes), the vertical pass reads them back conflict-free (a warp reads
// 32 consecutive bytes per tap), results leave as 16-byte (image) / 8-byte (patch operand) stores.
#include <math.h>

#include "../../include/vitb200.h"
#include "vitb_common.cuh"
#include "vitb_image_prep_core.h"

namespace {

constexpr int kPrepThreads = 256;

__global__ void __launch_bounds__(kPrepThreads) image_prep_kernel(const vitb_prep::Args a) {
  extern __shared__ __align__(16) uint8_t prep_smem[];
  vitb_prep::phase1(a, blockIdx.x, threadIdx.x, kPrepThreads, prep_smem);
  __syncthreads();
  vitb_prep::phase2(a, blockIdx.x, threadIdx.x, kPrepThreads, prep_smem);
}

inline double triangle(double x) {
  if (x < 0.0) x = -x;
  return x < 1.0 ? 1.0 - x : 0.0;
}

inline double axis_scale(int in_size, int out_size) { return static_cast<double>(in_size) / static_cast<double>(out_size); }

inline int axis_ksize(int in_size, int out_size) {
  double fs = axis_scale(in_size, out_size);
  if (fs < 1.0) fs = 1.0;
  return static_cast<int>(ceil(fs)) * 2 + 1;
}

}  // namespace

extern "C" {

// Host-only: Pillow's precompute_coeffs + normalize_coeffs_8bpc for the bilinear filter over a whole axis.
int vitb_resize_tables_host(int in_size, int out_size, int32_t* bounds_host, int32_t* coeffs_host,
                            int coeffs_capacity, int* ksize_host) {
  VITB_REQUIRE(in_size > 0 && out_size > 0, VITB_ERR_BAD_ARG, "resize_tables: sizes must be positive");
  const int ksize = axis_ksize(in_size, out_size);
  if (ksize_host) *ksize_host = ksize;
  if (!bounds_host && !coeffs_host) return VITB_OK;   // size query
  VITB_REQUIRE(bounds_host && coeffs_host, VITB_ERR_BAD_ARG, "resize_tables: both tables or neither");
  VITB_REQUIRE(static_cast<long long>(coeffs_capacity) >= static_cast<long long>(out_size) * ksize, VITB_ERR_WORKSPACE,
               "resize_tables: coeffs_capacity %d < %d x %d", coeffs_capacity, out_size, ksize);
  const double scale = axis_scale(in_size, out_size);
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = 1.0 * filterscale;
  const double ss = 1.0 / filterscale;
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale;
    int xmin = static_cast<int>(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = static_cast<int>(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    const int n = xmax - xmin;
    double ww = 0.0;
    for (int x = 0; x < n; ++x) ww += triangle((x + xmin - center + 0.5) * ss);
    int32_t* k = coeffs_host + static_cast<size_t>(xx) * ksize;
    for (int x = 0; x < ksize; ++x) {
      double w = 0.0;
      if (x < n) {
        w = triangle((x + xmin - center + 0.5) * ss);   // the same double the sum above saw
        if (ww != 0.0) w /= ww;
      }
      const double f = w * static_cast<double>(1 << vitb_prep::kPrecisionBits);
      k[x] = w < 0.0 ? static_cast<int32_t>(-0.5 + f) : static_cast<int32_t>(0.5 + f);
    }
    bounds_host[2 * xx] = xmin;
    bounds_host[2 * xx + 1] = n;
  }
  return VITB_OK;
}

int vitb_image_prep(const uint8_t* src, int B, int H, int W, int C, int out_h, int out_w, const int32_t* xbounds,
                    const int32_t* xcoeffs, int xksize, const int32_t* ybounds, const int32_t* ycoeffs, int yksize,
                    const uint8_t* flip, const float* lut, float* out_img, int P, int ldk, void* cols_hi,
                    void* cols_lo, uint8_t* out_u8, void* stream_) {
  int st = vitb_check_device();
  if (st != VITB_OK) return st;
  if (B == 0) return VITB_OK;
  VITB_REQUIRE(src && lut && B > 0 && H > 0 && W > 0 && C > 0 && C <= 4 && out_h > 0 && out_w > 0, VITB_ERR_BAD_ARG,
               "image_prep: bad args");
  VITB_REQUIRE(out_img || cols_hi || out_u8, VITB_ERR_BAD_ARG, "image_prep: no output requested");
  VITB_REQUIRE((out_w == W) == (xbounds == nullptr) && (out_h == H) == (ybounds == nullptr), VITB_ERR_BAD_ARG,
               "image_prep: a pass needs its tables exactly when that axis changes size (Pillow skips the pass otherwise)");
  VITB_REQUIRE(!xbounds || (xcoeffs && xksize == axis_ksize(W, out_w)), VITB_ERR_BAD_ARG,
               "image_prep: column tables do not belong to %d -> %d", W, out_w);
  VITB_REQUIRE(!ybounds || (ycoeffs && yksize == axis_ksize(H, out_h)), VITB_ERR_BAD_ARG,
               "image_prep: row tables do not belong to %d -> %d", H, out_h);
  VITB_REQUIRE(!cols_lo || cols_hi, VITB_ERR_BAD_ARG, "image_prep: cols_lo without cols_hi");
  vitb_prep::Args a = {};
  a.src = src; a.B = B; a.H = H; a.W = W; a.C = C; a.out_h = out_h; a.out_w = out_w;
  a.xb = xbounds; a.xc = xcoeffs; a.xk = xksize; a.yb = ybounds; a.yc = ycoeffs; a.yk = yksize;
  a.flip = flip; a.lut = lut; a.out_img = out_img; a.out_u8 = out_u8;
  a.vec4_img = (out_w % 4 == 0) && (reinterpret_cast<uintptr_t>(out_img) % 16 == 0);
  a.cols_hi = reinterpret_cast<__nv_bfloat16*>(cols_hi);
  a.cols_lo = reinterpret_cast<__nv_bfloat16*>(cols_lo);
  if (cols_hi) {
    VITB_REQUIRE(P > 0 && out_h >= P && out_w >= P, VITB_ERR_BAD_ARG, "image_prep: patch size %d does not fit %dx%d", P,
                 out_h, out_w);
    VITB_REQUIRE(ldk % 8 == 0 && ldk >= C * P * P, VITB_ERR_UNSUPPORTED_SHAPE,
                 "image_prep: ldk=%d must be a multiple of 8 and >= %d", ldk, C * P * P);
    VITB_REQUIRE(reinterpret_cast<uintptr_t>(cols_hi) % 16 == 0 && reinterpret_cast<uintptr_t>(cols_lo) % 16 == 0,
                 VITB_ERR_BAD_ARG, "image_prep: patch operands must be 16-byte aligned");
    a.P = P; a.ldk = ldk; a.gh = out_h / P; a.gw = out_w / P;
  }
  int band = 0, rows_cap = 0;
  const size_t smem = vitb_prep::choose_band(H, out_h, ybounds != nullptr, C, out_w, &band, &rows_cap);
  VITB_REQUIRE(smem <= 200 * 1024, VITB_ERR_UNSUPPORTED_SHAPE,
               "image_prep: one output row needs %zu bytes of shared memory (%d source rows x %d columns)", smem, rows_cap,
               out_w);
  a.band_rows = band;
  a.rows_cap = rows_cap;
  const long long bands = (out_h + band - 1) / band;
  VITB_REQUIRE(bands * B < (1ll << 31), VITB_ERR_UNSUPPORTED_SHAPE, "image_prep: grid too large");
  if (smem > 48 * 1024) {
    VITB_CUDA_CHECK(cudaFuncSetAttribute(image_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem)));
  }
  image_prep_kernel<<<static_cast<unsigned>(bands * B), kPrepThreads, smem, reinterpret_cast<cudaStream_t>(stream_)>>>(a);
  VITB_LAUNCH_CHECK("image_prep_kernel");
  return VITB_OK;
}

}  // extern "C"
