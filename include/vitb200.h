/* vitb200.h — C ABI of libvitb200.so, the sm_100a (B200) kernel library behind the
 * reference-shaped nn.Module API of the ViT encoder hot path.
 *
 * The reference (sea-with-sakura/ViT-of-Pytorch) has no FFI of its own: every op on the path is an
 * ATen call made from src/model.py / res-vit/model.py.  Each entry point below therefore cites the
 * reference LINES whose device work it replaces; INTEGRATION.md shows the ctypes binding a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is DEVICE memory unless its name ends in _host.
 *   - the library never allocates, frees or retains device memory and never synchronises the
 *     device; scratch space is passed in by the caller (workspace / workspace_bytes).
 *   - every launching call takes the CUDA stream as a void* (cudaStream_t).
 *   - return value: 0 = OK, negative = error (VITB_ERR_*); text via vitb_last_error().
 *   - a device that is not compute capability 10.x is a hard error (VITB_ERR_UNSUPPORTED_ARCH):
 *     there is no fallback path of any kind.
 *   - dtype codes: 0 = float32, 1 = bfloat16.
 */
#ifndef VITB200_H_
#define VITB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VITB200_VERSION 100

#define VITB_F32 0
#define VITB_BF16 1

/* ---- library ------------------------------------------------------------------------------- */
int vitb_version(void);
int vitb_last_error(char* buf, size_t n);
int vitb_device_check(void);

/* ---- dense contraction (tcgen05 / TMEM / TMA) ------------------------------------------------
 * D[M,N] = epilogue( sum_seg  A_seg[M,K_seg] * B_seg[N,K_seg]^T )
 *
 * Replaces every dense contraction of the path:
 *   LinearGeneral tensordot        src/model.py:61-63 (q/k/v :86-88, out :99)
 *   nn.Linear fc1/fc2/classifier   src/model.py:43-48,210 ; res-vit/model.py:265-271,294,315-317,679
 *   Conv2d patch embedding (GEMM)  src/model.py:197      ; res-vit/model.py:602
 *   LoRA / router / approximators  res-vit/model.py:110-117,178,188,329-333
 *   and all of their autograd backward GEMMs (dgrad, wgrad).
 *
 * Operands are bf16.  Up to three (A,B,K) segments accumulate into ONE TMEM accumulator before the
 * epilogue runs: segment 2 carries the LoRA rank-r update (x*A^T)[M,r] x B[N,r]; three segments
 * carry the bf16x3 split used by the fp32 parity mode.
 *   a_mn_major = 0: A_seg is [M,K] row-major (K contiguous), leading dimension lda (elements)
 *   a_mn_major = 1: A_seg is stored [K,M] (M contiguous)  — used by wgrad (A = dY^T)
 *   b_mn_major = 0: B_seg is [N,K] row-major (nn.Linear weight layout)
 *   b_mn_major = 1: B_seg is stored [K,N] (N contiguous)  — LinearGeneral layout / dgrad / wgrad
 * Epilogue, in this order:  v = acc ; v += bias[n] ; v += row_bias[m / row_bias_group, n] ;
 *   (GELU: optionally store v to d2 (pre-activation), v = gelu_erf(v)) ;
 *   (GELU_BWD: v *= gelu_erf'(aux[m,n])) ; v += residual[rm, n] ; D[om, n] (+)= v
 * where for row_remap_group = g > 0 (patch-embedding scatter):  om = m + m/g + 1, rm = m%g + 1,
 * otherwise om = rm = m.
 */
#define VITB_EPI_NONE 0
#define VITB_EPI_GELU 1
#define VITB_EPI_GELU_BWD 2

typedef struct vitb_gemm_params {
  int32_t struct_bytes; /* sizeof(vitb_gemm_params) — ABI guard */
  int32_t M, N;
  int32_t num_segments; /* 1..3 */
  const void* A[3];
  const void* B[3];
  int64_t lda[3];
  int64_t ldb[3];
  int32_t K[3];
  int32_t a_mn_major;
  int32_t b_mn_major;
  int32_t split_k; /* >= 1; > 1 requires accumulate = 1 and d_dtype = f32 */
  int32_t epilogue; /* VITB_EPI_* */
  void* D;
  int64_t ldd;
  int32_t d_dtype;
  int32_t accumulate; /* 1: D += v via fp32 atomics (wgrad / split-K) */
  void* D2; /* optional bf16 pre-activation output for VITB_EPI_GELU */
  int64_t ldd2;
  const float* bias; /* [N] or NULL */
  const float* row_bias; /* [ceil(M/row_bias_group), N] or NULL */
  int32_t row_bias_group;
  int32_t row_remap_group;
  const void* residual; /* [*, N] or NULL */
  int64_t ldr;
  int32_t r_dtype;
  int32_t _pad0;
  const void* aux; /* bf16 [M,N] for VITB_EPI_GELU_BWD */
  int64_t ldaux;
} vitb_gemm_params;

int vitb_gemm(const vitb_gemm_params* p, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VITB200_H_ */
