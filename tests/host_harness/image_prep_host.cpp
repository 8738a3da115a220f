// tests/host_harness/image_prep_host.cpp — TEST SCAFFOLDING (built by tests/test_image_prep_host.py with g++).
//
// Walks the blocks and threads of image_prep_kernel (vit-of-pytorch_b200/csrc/vitb_image_prep.cu) on the host over
// the SAME phase bodies the kernel compiles (vitb_image_prep_core.h): phase 1 for every thread of a block, then —
// where the kernel has its __syncthreads() — phase 2 for every thread.  It exists so the kernel's indexing and
// integer arithmetic can be compared with the oracle in the build container, which has no GPU; it is not part
// of libvitb200.so and no product code links it.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../vit-of-pytorch_b200/csrc/vitb_image_prep_core.h"

extern "C" int image_prep_host(const uint8_t* src, int B, int H, int W, int C, int out_h, int out_w,
                               const int32_t* xb, const int32_t* xc, int xk, const int32_t* yb, const int32_t* yc,
                               int yk, const uint8_t* flip, const float* lut, float* out_img, int P, int ldk,
                               uint16_t* cols_hi, uint16_t* cols_lo, uint8_t* out_u8, int nthreads,
                               int* band_rows_out, int* rows_cap_out) {
  vitb_prep::Args a;
  memset(&a, 0, sizeof(a));
  a.src = src; a.B = B; a.H = H; a.W = W; a.C = C; a.out_h = out_h; a.out_w = out_w;
  a.xb = xb; a.xc = xc; a.xk = xk; a.yb = yb; a.yc = yc; a.yk = yk;
  a.flip = flip; a.lut = lut; a.out_img = out_img; a.out_u8 = out_u8;
  a.vec4_img = (out_w % 4 == 0) && (reinterpret_cast<uintptr_t>(out_img) % 16 == 0);
  a.cols_hi = reinterpret_cast<__nv_bfloat16*>(cols_hi);
  a.cols_lo = reinterpret_cast<__nv_bfloat16*>(cols_lo);
  if (cols_hi) { a.P = P; a.ldk = ldk; a.gh = out_h / P; a.gw = out_w / P; }
  int band = 0, cap = 0;
  const size_t bytes = vitb_prep::choose_band(H, out_h, yb != nullptr, C, out_w, &band, &cap);
  a.band_rows = band;
  a.rows_cap = cap;
  if (band_rows_out) *band_rows_out = band;
  if (rows_cap_out) *rows_cap_out = cap;
  const int bands = (out_h + band - 1) / band;
  // guard bytes after the scratch catch writes past the size the launch would have requested
  std::vector<uint8_t> scratch(bytes + 64);
  for (int blk = 0; blk < bands * B; ++blk) {
    memset(scratch.data(), 0xCD, scratch.size());
    for (int t = 0; t < nthreads; ++t) vitb_prep::phase1(a, blk, t, nthreads, scratch.data());
    for (size_t g = bytes; g < scratch.size(); ++g)
      if (scratch[g] != 0xCD) return -100;
    for (int t = 0; t < nthreads; ++t) vitb_prep::phase2(a, blk, t, nthreads, scratch.data());
  }
  return 0;
}
