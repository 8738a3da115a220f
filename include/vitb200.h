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
/* sizeof of an ABI struct as compiled into the library: 0 = vitb_gemm_params, 1 = vitb_attn_params. */
int vitb_struct_size(int which);

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
 *   (GELU_DG: store gelu_erf'(v) to d2, v = gelu_erf(v)) ; (GELU_BWD: v *= gelu_erf'(aux[m,n])) ;
 *   (MUL_AUX: v *= aux[m,n]) ; v += residual[rm, n] ; D[om, n] (+)= v ; colsum[n] += v
 * where for row_remap_group = g > 0 (patch-embedding scatter):  om = m + m/g + 1, rm = m%g + 1,
 * otherwise om = rm = m.
 */
#define VITB_EPI_NONE 0
#define VITB_EPI_GELU 1
#define VITB_EPI_GELU_BWD 2
#define VITB_EPI_GELU_DG 3  /* like GELU, but d2 receives gelu_erf'(v): the backward then only multiplies */
#define VITB_EPI_MUL_AUX 4  /* v *= aux[m,n]  (aux = the d2 of a VITB_EPI_GELU_DG forward) */

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
  void* D2; /* pre-activation (VITB_EPI_GELU, optional) or derivative (VITB_EPI_GELU_DG) output, same dtype as D */
  int64_t ldd2;
  const float* bias; /* [N] or NULL */
  const float* row_bias; /* [ceil(M/row_bias_group), N] or NULL */
  int32_t row_bias_group;
  int32_t row_remap_group;
  const void* residual; /* [*, N] or NULL */
  int64_t ldr;
  int32_t r_dtype;
  int32_t _pad0;
  const void* aux; /* [M,N] pre-activation (VITB_EPI_GELU_BWD) or derivative (VITB_EPI_MUL_AUX), same dtype as D */
  int64_t ldaux;
  float* colsum; /* optional [N]: += column sums of the values written to D (bias gradient of the producer) */
  /* Column groups (the merged q|k|v projection and its weight gradient): N = n_groups * Ng.  Column n belongs to group
   * g = n / Ng; its B rows / columns live at B_seg + g * b_group_stride (elements; every segment uses the same stride),
   * i.e. the n_groups weight matrices only have to sit at a uniform distance, not in one [K, N] array.  With
   * d_group_stride != 0 the output is grouped the same way: D(m, n) at D + g * d_group_stride + m * ldd + (n - g * Ng)
   * (fp32 accumulate outputs only: the n_groups weight gradients).  n_groups <= 1: plain GEMM.  Ng must be a multiple
   * of the N tile (256, or 128 when that pads N less). */
  int32_t n_groups;
  int32_t _pad1;
  int64_t b_group_stride;
  int64_t d_group_stride;
  /* optional DEVICE scalar: the number of rows of A / D that actually hold data (<= M).  Tiles past it are skipped; M
   * stays the capacity the buffers and tensor maps are built for (rows between *m_dev and the end of its last 128-row
   * tile are computed from whatever the buffers hold and written — they belong to the caller's scratch capacity).
   * Token-major GEMMs only (a_mn_major = 0, split_k <= 1, no colsum). */
  const int32_t* m_dev;
} vitb_gemm_params;

int vitb_gemm(const vitb_gemm_params* p, void* stream);

/* ---- LayerNorm ----------------------------------------------------------------------------------
 * nn.LayerNorm(D, eps) — src/model.py:108,114,146 (calls :119,:127,:155); res-vit/model.py:119-130.
 * x is [rows, D] with row stride x_row_stride (elements); D must be a multiple of 128.
 * Outputs (each optional, contiguous [rows, D]): y_f32, y_bf16 (GEMM operand), y_bf16_lo (the
 * x - bf16(x) half for the bf16x3 fp32-parity GEMMs); mean/rstd [rows] are saved for the backward.
 */
int vitb_layernorm_fwd(const void* x, int x_dtype, int64_t x_row_stride, int rows, int D,
                       const float* gamma, const float* beta, float eps, float* y_f32, void* y_bf16,
                       void* y_bf16_lo, float* mean, float* rstd, void* stream);

/* Backward of the above.  dx = LN'(dy) + dres (dres = gradient arriving on the residual branch,
 * optional).  dgamma/dbeta/dcolsum [D] are ACCUMULATED (+=); dcolsum = column sums of dx, i.e. the
 * bias gradient of the Linear whose output (plus residual) fed this LayerNorm. */
int vitb_layernorm_bwd(const void* dy, int dy_dtype, const float* x, int64_t x_row_stride,
                       const float* mean, const float* rstd, const float* gamma, int rows, int D,
                       const float* dres, int64_t dres_row_stride, float* dx_f32,
                       int64_t dx_row_stride, void* dx_bf16, void* dx_bf16_lo, float* dgamma,
                       float* dbeta, float* dcolsum, void* stream);
/* Same, when only every dres_every-th row has a residual-branch gradient (dres row r / dres_every belongs to row r,
 * r % dres_every == 0): the class-token-only last encoder block, where of an image's N rows only row 0 is read
 * downstream (src/model.py:155,210). */
int vitb_layernorm_bwd_sparse_res(const void* dy, int dy_dtype, const float* x, int64_t x_row_stride,
                                  const float* mean, const float* rstd, const float* gamma, int rows, int D,
                                  const float* dres, int64_t dres_row_stride, int dres_every, float* dx_f32,
                                  int64_t dx_row_stride, void* dx_bf16, void* dx_bf16_lo, float* dgamma,
                                  float* dbeta, float* dcolsum, void* stream);

/* ---- attention ----------------------------------------------------------------------------------
 * softmax(q k^T / sqrt(dh)) v with no mask and no dropout — SelfAttention.forward
 * src/model.py:90-97, Attention.forward res-vit/model.py:273-293 — and its backward.
 * Element (b, n, h, d) of a tensor lives at base + b*batch_stride + n*row_stride + h*head_dim + d
 * (strides in elements), so q/k/v can alias one packed [T, 3D] projection output.
 *   *_tc  : tcgen05/TMEM kernels, bf16, Nq == Nk; gradients dq/dk/dv are bf16.  vitb_attn_fwd_tc takes head_dim
 *           64..128 (multiples of 16) and any token count; vitb_attn_bwd_tc head_dim 64 and <= 256 tokens;
 *           vitb_attn_bwd_tc_long (below) any token count and head_dim 64..128.
 *   *_simt: CUDA-core fp32 math, dtype f32 or bf16, any head_dim, Nk <= 320, Nq != Nk allowed: the fp32 parity mode
 *           and asymmetric query / key counts; gradients dq/dk/dv are FP32 and dq must be zeroed by the caller
 *           (atomic accumulation).
 * lse is [B, H, Nq] fp32 (log-sum-exp of the scaled scores), written by fwd and read by bwd.
 */
typedef struct vitb_attn_params {
  int32_t struct_bytes;
  int32_t dtype;
  int32_t B, H, Nq, Nk, head_dim;
  int32_t _pad0;
  const void* q;
  const void* k;
  const void* v;
  void* o;
  float* lse;
  int64_t q_batch_stride, q_row_stride;
  int64_t k_batch_stride, k_row_stride;
  int64_t v_batch_stride, v_row_stride;
  int64_t o_batch_stride, o_row_stride;
  const void* dout;
  int64_t do_batch_stride, do_row_stride;
  void* dq;
  void* dk;
  void* dv;
  int64_t dq_batch_stride, dq_row_stride;
  int64_t dk_batch_stride, dk_row_stride;
  int64_t dv_batch_stride, dv_row_stride;
} vitb_attn_params;

int vitb_attn_supported_tc(int head_dim, int Nq, int Nk);      /* forward AND backward on tcgen05 (head_dim 64, <= 256 tokens) */
int vitb_attn_fwd_supported_tc(int head_dim, int Nq, int Nk);  /* forward only: also 64 < head_dim <= 128, <= 320 tokens (ViT-H/14) */
int vitb_attn_fwd_tc(const vitb_attn_params* p, void* stream);
int vitb_attn_bwd_tc(const vitb_attn_params* p, void* stream);
/* Persistent warp-specialised generation of the tcgen05 kernels (vitb_attention_ws.cu): one resident CTA per SM walks a
 * contiguous range of (image, head[, query tile]) items; a TMA producer warp, a single-thread tcgen05 issuer and eight
 * CUDA-core warps overlap the loads, MMAs and softmax arithmetic of neighbouring items.  Same contract and shapes as
 * vitb_attn_fwd_tc / vitb_attn_bwd_tc (bf16, head_dim 64, Nq == Nk <= 256; the forward up to 240 tokens). */
int vitb_attn_ws_supported(int which /* 0 forward, 1 backward */, int head_dim, int Nq, int Nk);
int vitb_attn_fwd_ws(const vitb_attn_params* p, void* stream);
int vitb_attn_bwd_ws(const vitb_attn_params* p, void* stream);
int vitb_attn_fwd_simt(const vitb_attn_params* p, void* stream);
int vitb_attn_bwd_simt(const vitb_attn_params* p, void* stream);
/* Single-query attention (Nq == 1: the class-token row of the LAST encoder block, whose other rows nobody reads;
 * src/model.py:155,210), bf16, head_dim 64, <= 256 keys, 16-byte aligned rows.  The forward is vitb_attn_fwd_simt
 * (it picks the bandwidth-bound kernel for these shapes); this is the backward with BF16 gradients: dq [B,1,H*64],
 * dk / dv rows written whole through their strides (they may alias a packed [T, 2D] buffer), nothing to zero. */
/* tcgen05 backward for ANY number of tokens (bf16, head_dim 64, Nq == Nk): 384 px inputs give 577 tokens
 * (src/config.py:12).  Key blocks of 256 are spread over CTAs, each streams every query tile; dq_acc is an fp32
 * [B, N, H*64] scratch buffer ZEROED BY THE CALLER that collects dQ (red.global.add) before it is written to p->dq as bf16.
 * Gradients are bf16 as in vitb_attn_bwd_tc. */
int vitb_attn_bwd_long_supported(int head_dim, int Nq, int Nk);
int vitb_attn_bwd_tc_long(const vitb_attn_params* p, float* dq_acc, void* stream);
int vitb_attn_q1_supported(int head_dim, int Nk);
int vitb_attn_q1_bwd(const vitb_attn_params* p, void* stream);

/* ---- operand preparation / embedding stage ---------------------------------------------------- */
/* hi = bf16(x); lo = bf16(x - hi) (optional).  Weight shadows and fp32-parity operands. */
int vitb_cast_split(const float* x, int64_t n, void* hi, void* lo, void* stream);
/* Patch extraction for the Conv2d(3,D,P,P) patch embedding (src/model.py:179,197;
 * res-vit/model.py:543,602): img [B,C,H,W] fp32 -> rows (b,py,px) x k (c,ph,pw), K padded to ldk. */
int vitb_im2col(const float* img, int B, int C, int H, int W, int P, int ldk, void* hi, void* lo,
                void* stream);
/* x[b,0,:] = cls + pos[0]  (cls_token.repeat + cat + pos add, src/model.py:203-204,17). */
int vitb_cls_rows(float* x, int B, int N, int D, const float* cls, const float* pos, void* stream);
/* Backward of the embedding stage from dx [B,N,D]: dpos [N,D] += sum_b dx; dcls [D] += sum_b dx[:,0];
 * dbias [D] += sum over patch rows; dpatch [B*(N-1), D] = bf16 patch rows (wgrad operand). */
int vitb_embed_bwd(const float* dx, int B, int N, int D, float* dpos, float* dcls, float* dbias,
                   void* dpatch_hi, void* dpatch_lo, void* stream);
/* out = dy * gelu_erf'(z), elementwise over n values of dtype f32 or bf16 (backward of nn.GELU(),
 * src/model.py:33,44; res-vit/model.py:154,158,160,312). */
int vitb_gelu_bwd(const void* dy, const void* z, void* out, int64_t n, int dtype, void* stream);
/* nn.Dropout of PositionEmbs / MlpBlock / EncoderBlock (src/model.py:11-20,34-50,110-123), training mode:
 * y = [residual +] keep * x / (1 - p), keep ~ Bernoulli(1 - p) drawn with Philox4x32-10 from (seed, draw counter).
 * x: n values of dtype f32 or bf16; y has the dtype of x, or fp32 when a (fp32) residual is given; mask: n bytes
 * (0 / 1), kept for the backward.  state: two uint64 in DEVICE memory, zeroed once by the caller: state[0] is the draw
 * counter, advanced by every call ON THE DEVICE, so a CUDA graph that contains the call draws a new mask per replay.
 * The mask stream is this library's own: it is not torch's CUDA generator stream (which the reference's CPU runs do not
 * share either); parity for dropout is statistical (tests/test_kernels_gpu.py). */
int vitb_dropout_fwd(const void* x, int x_dtype, const float* residual, void* y, uint8_t* mask, int64_t n, float p,
                     uint64_t seed, uint64_t* state, void* stream);
/* dx = dy * mask / (1 - p); dy / dx of dtype f32 or bf16 independently. */
int vitb_dropout_bwd(const void* dy, int dy_dtype, const uint8_t* mask, void* dx, int dx_dtype, int64_t n, float p,
                     void* stream);
/* out[c] += sum_r x[r,c]  (bias gradients). */
int vitb_colsum(const void* x, int x_dtype, int rows, int cols, int64_t ld, float* out, void* stream);
/* Same over a packed [rows, 3*seg_cols] buffer (dq|dk|dv): segment i accumulates into out_i[seg_cols]. */
int vitb_colsum3(const void* x, int x_dtype, int rows, int seg_cols, int64_t ld, float* out0, float* out1,
                 float* out2, void* stream);

/* ---- gradient exchange over NVLink peer memory ---------------------------------------------------
 * all-reduce (sum x scale: scale = 1 / world gives the average) of the flat fp32 gradient buffer of a data-parallel step —
 * nn.DataParallel's gradient reduction, src/train.py:128-129 — as ONE kernel per rank.  bufs[world]: every rank's mapping
 * of every rank's buffer (symmetric memory; this rank's own buffer is bufs[rank]); pads[world]: the same for a signal pad of
 * vitb_p2p_pad_words() uint32 words per rank, ZEROED ONCE by the caller before the first call; multicast: the NVSwitch
 * multicast address of the buffer (the switch then does the sum: NVLS) or null (peer loads / stores).  n fp32 elements, a
 * multiple of 4.  Every rank must call it, with the same n, once per step, on a stream with nothing else in flight (the
 * blocks of all ranks meet at two in-kernel barriers).  Replayable from a CUDA graph. */
int vitb_p2p_pad_words(void);
int vitb_p2p_allreduce(float* const* bufs, uint32_t* const* pads, float* multicast, int rank, int world, int64_t n,
                       float scale, void* stream);

/* ---- Res-ViT routing ---------------------------------------------------------------------------
 * Decision tail of RouterModule.forward (res-vit/model.py:189-211) + _router2indices (:169-173).
 * logits [T, bs, 2] fp32 (T = B*N tokens, token n = t %% N is "reserved" when n < reserve_initials):
 *   soft   = softmax(logits, -1)
 *   entropy_sum += -sum over non-reserved tokens of p*log(p + 1e-8)   (caller divides by B*(N-r0)*bs)
 *   hard   = one_hot(argmax softmax((logits + noise)/tau))  in training (noise = Gumbel sample supplied by
 *            the caller, exactly -log(Exponential(1)) as torch's F.gumbel_softmax draws it; ysoft saved)
 *          = one_hot(argmax soft)                            in eval        (first index wins ties)
 *   reserved tokens are forced to (0, 1);  indices[t] = sum_i hard[t,i,1] * 2^(bs-1-i)  (fp32 integers)
 */
int vitb_router_decide_fwd(const float* logits, const float* noise, int T, int N, int block_size,
                           int reserve_initials, int training, float tau, float* soft, float* hard,
                           float* ysoft, float* indices, float* entropy_sum, void* stream);
/* d_logits = softmax'(soft; d_soft + d_entropy*entropy_scale*dH/dp) + (training) softmax'(ysoft; d_hard)/tau;
 * reserved tokens receive no entropy / straight-through gradient.  d_soft, d_hard, d_entropy optional. */
int vitb_router_decide_bwd(const float* soft, const float* ysoft, const float* d_soft, const float* d_hard,
                           const float* d_entropy, float entropy_scale, int T, int N, int block_size,
                           int reserve_initials, int training, float tau, float* d_logits, void* stream);
/* Global router feature: out[b,c] = mean_{n >= reserve_initials} x[b,n,c] (res-vit/model.py:180-184); x is
 * [B,N,C] f32/bf16.  bwd writes dx[b,n,c] = dg[b,c]/(N-r0) for n >= r0, 0 otherwise. */
int vitb_token_mean_fwd(const void* x, int dtype, int B, int N, int C, int reserve_initials, float* out, void* stream);
int vitb_token_mean_bwd(const float* dg, int dtype, int B, int N, int C, int reserve_initials, void* dx, void* stream);
/* out[t,:] = ((member_mask >> (int)index[t]) & 1) ? a[t,:] : b[t,:]; a or b may be NULL (= zeros).
 * torch.isin + blend (res-vit/model.py:469-472,487,524) and approximator row selection (:349-368). */
int vitb_select_rows(const void* a, const void* b, const float* index, uint32_t member_mask, int rows, int cols,
                     int dtype, void* out, void* stream);
/* Same; additionally *any_member = 1 (int32, device, never cleared here) when at least one row was a member: whether an
 * approximator's key occurred in the batch, i.e. whether the reference's module ran at all (res-vit/model.py:363-367). */
int vitb_select_rows_flag(const void* a, const void* b, const float* index, uint32_t member_mask, int rows, int cols,
                          int dtype, void* out, int32_t* any_member, void* stream);

/* Device-side row compaction for Res-ViT's inference-time token skipping (res-vit/model.py:503-524): nothing
 * about the selection travels to the host.  index [T] fp32 packed router indices (as vitb_select_rows):
 *   vitb_compact_rows  rows[*count ...] receives every t with ((member_mask >> (int)index[t]) & 1), *count grows by their
 *                      number (zero it first); the order of the list is unspecified (every consumer is row-wise).
 *   vitb_gather_rows   dst[i, :] = src[rows[i], :] for i < *count     (max_rows bounds the launch; cols of dtype f32/bf16,
 *   vitb_scatter_rows  dst[rows[i], :] = src[i, :] for i < *count      rows 16-byte aligned and a multiple of 16 bytes)
 * The GEMMs between a gather and a scatter take their row count from the same device scalar (vitb_gemm_params.m_dev). */
int vitb_compact_rows(const float* index, uint32_t member_mask, int T, int32_t* rows, int32_t* count, void* stream);
int vitb_gather_rows(const void* src, int64_t src_ld, int dtype, const int32_t* rows, const int32_t* count, int max_rows,
                     int cols, void* dst, int64_t dst_ld, void* stream);
int vitb_scatter_rows(const void* src, int64_t src_ld, int dtype, const int32_t* rows, const int32_t* count, int max_rows,
                      int cols, void* dst, int64_t dst_ld, void* stream);

/* Res-ViT scalar losses; the reduction also writes the gradient its backward needs (no second pass).
 * A grid of blocks with one atomic per block: vitb_distill_loss is one launch, vitb_active_loss three small ones (sum,
 * gradient, scalars) behind a 4-byte memset of *ratio_out, which is REQUIRED (it carries the sum between the launches).
 * vitb_distill_loss — DistillLoss, res-vit/model.py:40-59: *loss_acc += mean((s - t)^2) over [rows, cols] (rows may be
 *   strided: the class-token rows of student / teacher); d_student [rows, cols] fp32 (optional) = 2 (s - t) / (rows cols).
 * vitb_active_loss — ActiveLoss, res-vit/model.py:61-85: ratio = mean of probs[b, n >= reserve_initials, j] (probs
 *   [B, N, L] fp32); *loss = (ratio + *shift_dev - target)^2; d_probs (optional) = 2 (ratio + shift - target) / count on the
 *   counted entries, 0 on the reserved tokens.  shift_dev (optional device scalar): global-batch mean minus this shard's
 *   mean under data parallelism. */
int vitb_distill_loss(const void* student, int64_t s_row_stride, const void* teacher, int64_t t_row_stride, int dtype, int rows,
                      int cols, float* loss_acc, float* d_student, void* stream);
int vitb_active_loss(const float* probs, int B, int N, int L, int reserve_initials, float target, const float* shift_dev,
                     float* ratio_out, float* loss, float* d_probs, void* stream);

/* ---- input transform in front of the encoder --------------------------------------
 * The per-sample CPU work of the reference's loaders — transforms.Compose([Resize(S), RandomHorizontalFlip(),
 * ToTensor(), Normalize(mean, std)]) at src/data_loaders.py:36-48 (CIFAR-10), :69-82 (CIFAR-100), :102-114
 * (ImageNet, Resize((S,S))) — on the device, bit-exact with torchvision + Pillow: the resize is Pillow's 8-bit
 * bilinear resample (22-bit fixed-point weights, horizontal pass rounded to a byte before the vertical pass).
 *
 * vitb_resize_tables_host is HOST code (no device work): the windows and weights of one axis in_size -> out_size,
 * Pillow's precompute_coeffs + normalize_coeffs_8bpc.  bounds_host [out_size, 2] = (first source index, taps),
 * coeffs_host [out_size, ksize]; *ksize_host = taps per window.  With both table pointers NULL it only reports ksize.
 * The caller uploads the tables once per (in_size, out_size) and keeps them on the device.
 *
 * vitb_image_prep: src [B,H,W,C] uint8 (C <= 4) -> any of
 *   out_img  [B,C,out_h,out_w] fp32  — what the reference's loader yields and VisionTransformer.forward consumes,
 *   cols_hi  [B*gh*gw, ldk] bf16     — the patch-embedding GEMM operand (same layout as vitb_im2col; cols_lo = the
 *                                      bf16x3 low half); columns >= C*P*P are NOT written (zero them once),
 *   out_u8   [B,out_h,out_w,C] uint8 — the resized bytes after the flip.
 * x/y tables must be NULL exactly when that axis keeps its size (Pillow skips the pass).  flip [B] (optional):
 * non-zero mirrors the width axis — the caller draws it (torch.rand(1) < 0.5 per image, as RandomHorizontalFlip does).
 * lut [C,256] fp32: the normalised value of each byte, ((v/255) - mean_c)/std_c evaluated by the caller in the
 * reference's arithmetic. */
int vitb_resize_tables_host(int in_size, int out_size, int32_t* bounds_host, int32_t* coeffs_host,
                            int coeffs_capacity, int* ksize_host);
int vitb_image_prep(const uint8_t* src, int B, int H, int W, int C, int out_h, int out_w, const int32_t* xbounds,
                    const int32_t* xcoeffs, int xksize, const int32_t* ybounds, const int32_t* ycoeffs, int yksize,
                    const uint8_t* flip, const float* lut, float* out_img, int P, int ldk, void* cols_hi,
                    void* cols_lo, uint8_t* out_u8, void* stream);

/* ---- loss and optimizer ------------------------------------------------------------------------- */
/* nn.CrossEntropyLoss (mean) — src/train.py:151,22; res-vit/model.py:550,681.
 * loss[0] = mean_b(lse_b - logits[b,label_b]); dlogits = (softmax - onehot)/B (optional). */
int vitb_cross_entropy(const float* logits, const int64_t* labels, int B, int C, float* loss,
                       float* dlogits, void* stream);
/* torch.optim.SGD(momentum) over a flat buffer (src/train.py:154-158), refreshing the bf16 shadow.
 * hyper_dev (optional DEVICE array of 4 floats {lr, momentum, dampening, weight_decay}) overrides the by-value arguments,
 * so an LR scheduler (OneCycleLR cycles lr AND momentum) keeps driving a captured CUDA graph. */
int vitb_sgd_momentum(float* p, const float* g, float* m, int64_t n, float lr, const float* hyper_dev, float momentum,
                      float dampening, float weight_decay, int nesterov, int first_step,
                      void* shadow_hi, void* shadow_lo, void* stream);
/* torch.optim.AdamW over a flat buffer (res-vit/train.py:272-277); grad_scale_dev (optional device
 * scalar) carries the clip_grad_norm_ coefficient (res-vit/train.py:65).  hyper_dev (optional DEVICE array of 4 floats
 * {lr, beta1, beta2, weight_decay}) / step_dev (optional device scalar) override the by-value arguments so a captured
 * CUDA graph follows the scheduler and the bias correction. */
int vitb_adamw(float* p, const float* g, float* m, float* v, int64_t n, float lr, const float* hyper_dev, float beta1,
               float beta2, float eps, float weight_decay, int step, const int* step_dev,
               const float* grad_scale_dev, void* shadow_hi, void* shadow_lo, void* stream);
/* The same update when single parameters of the flat buffer may have to be left alone: torch.optim.AdamW skips a parameter
 * whose .grad is None — no weight decay, no moment decay, and its own step count (the bias corrections) does not advance.
 * In Res-ViT that is a BlockPathApproximators member whose key did not occur in the batch (res-vit/model.py:349-368,
 * res-vit/train.py:272-277).  seg_end[nseg]: ascending end offsets of the parameters in the flat buffer; seg_flag[nseg]:
 * index into flags[] (int32, device; non-zero = received a gradient this step) or -1 = always updated; seg_step[nseg]: the
 * parameters' own step counts (fp32, device), advanced here for the live ones before the update.  hyper_dev = {lr, beta1,
 * beta2, weight_decay} on the device. */
int vitb_adamw_segments(float* p, const float* g, float* m, float* v, int64_t n, const float* hyper_dev, float eps,
                        const float* grad_scale_dev, void* shadow_hi, void* shadow_lo, const int64_t* seg_end,
                        const int32_t* seg_flag, const int32_t* flags, float* seg_step, int nseg, void* stream);
int vitb_sumsq(const float* x, int64_t n, float* out, void* stream);
int vitb_clip_coef(const float* sumsq, float max_norm, float* coef, float* norm_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VITB200_H_ */
