/*
 * pio_b200.h — C ABI of the B200-native Perceiver IO attention stack (libpio_b200.so).
 *
 * The reference (JOBR0/PerceiverIO_Pytorch) has no FFI layer: its operator interface for this path is the
 * nn.Module surface of perceiver_io/transformer_primitives.py (Attention :18-180, MLP :183-216,
 * SelfAttention :219-297, CrossAttention :300-406) driven by PerceiverEncoder / PerceiverDecoder
 * (perceiver_io/perceiver.py:98-107, :166-180).  Each entry point below names the reference lines whose
 * arithmetic it replaces.  The Python host mirror (perceiverio_pytorch_b200/primitives.py, perceiver.py) keeps
 * the reference's class names, constructor kwargs and state_dict layout and lowers every forward to these calls.
 *
 * Conventions
 *  - All pointers are DEVICE pointers owned by the caller (PyTorch); the library never allocates, frees or
 *    retains them.  `stream` is a cudaStream_t passed as void*; all work is enqueued on it, no host sync, so
 *    every call is CUDA-graph capturable.
 *  - bf16 matrices are row-major with a leading dimension (in elements) that is a multiple of 8 (16 bytes, the
 *    TMA global-stride rule) and a 16-byte aligned base.  Logical widths may be odd (261, 322, 1026): TMA
 *    zero-fills out-of-bounds columns, so no padding content is ever read.
 *  - 16-bit format: every args struct with 16-bit operands or outputs carries an `fp16` field.  0 (default): bfloat16.
 *    1: IEEE half — same tensor-core rate (tcgen05.mma kind::f16 takes both), 11 instead of 8 mantissa bits, i.e. 8x less
 *    operand rounding, at the price of fp16's range (conversions saturate at 65504).  All 16-bit buffers of one call use
 *    the same format.  The optical-flow recipe needs it: with bf16 operands the reference ALGORITHM itself misses the
 *    1e-2 bound there, and the reference's own mixed-precision mode is fp16 autocast
 *    (flow_perceiver.py:14,129).
 *  - Return value: 0 on success, negative pio_status otherwise; pio_last_error() gives a thread-local message.
 *    Nothing throws, aborts or falls back to a CPU path.
 */
#ifndef PIO_B200_H_
#define PIO_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PIO_ABI_VERSION 15

typedef enum pio_status {
  PIO_OK = 0,
  PIO_ERR_INVALID_ARGUMENT = -1, /* bad shape / alignment / null pointer */
  PIO_ERR_UNSUPPORTED = -2,      /* shape outside what the kernels implement */
  PIO_ERR_CUDA = -3,             /* CUDA runtime / driver / launch error */
  PIO_ERR_ARCH = -4              /* device is not sm_100 */
} pio_status;

int pio_abi_version(void);
const char* pio_last_error(void);
/* 0 if the current device can run the kernels (compute capability 10.x), PIO_ERR_ARCH otherwise. */
int pio_check_device(void);

/* ---------------------------------------------------------------------------------------------------------
 * LayerNorm + cast: y_bf16[r, 0:C] = (x[r] - mean) * rsqrt(var + eps) * gamma + beta, columns C..ldy-1 := 0.
 * Replaces nn.LayerNorm at transformer_primitives.py:281 (layer_norm1), :292 / :401 (layer_norm2),
 * :379-380 (layer_norm_kv, layer_norm_q).  gamma/beta may be NULL (plain normalisation); if `normalize` is 0
 * the kernel only casts (used for fp32 activations that feed a GEMM without a LayerNorm).
 * HBM-bound: 4*C bytes read + 2*ldy bytes written per row.
 * --------------------------------------------------------------------------------------------------------- */
typedef struct pio_layernorm_args {
  const float* x;   /* [rows, C], row stride ldx (elements) */
  int64_t ldx;
  void* y;          /* bf16 [rows, ldy] */
  int64_t ldy;
  const float* gamma;
  const float* beta;
  int64_t rows;
  int32_t C;
  int32_t normalize;
  float eps;
  int32_t split;    /* validation mode (bf16 x 2 split operands): y has ldy = 3 * pad8(C) columns laid out as
                       [hi | lo | hi] (split = 1, the A side of a product) or [hi | hi | lo] (split = 2, the B side) with
                       hi = bf16(v), lo = bf16(v - hi); one K = 3 * pad8(C) GEMM of an A-side by a B-side operand
                       evaluates hi*hi + lo*hi + hi*lo in fp32 */
  int32_t fp16;     /* 16-bit format of y: 0 = bf16, 1 = fp16 (not with split) */
} pio_layernorm_args;
int pio_layernorm_bf16(const pio_layernorm_args* a, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Batched GEMM with fused epilogue on tcgen05 tensor cores (TMA-staged operands, TMEM accumulators).  Large
 * problems run on CTA pairs (tcgen05.mma.cta_group::2, 256 x 256 tiles, residual prefetched and results stored by
 * TMA); small or oddly laid out ones on a single-CTA kernel with 128 x {64,128,256} tiles:
 *   acc[z] = A[z] (M x K, bf16, K contiguous)  x  B[z]
 *      B[z] is N x K with K contiguous (b_mn_major = 0: an nn.Linear weight [out, in] or K^T of attention), or
 *      B[z] is K x N with N contiguous (b_mn_major = 1: the V operand of P.V)
 *   v = alpha * acc + bias (per column: bias_mode 1, per row: bias_mode 2)
 *   v = gelu_erf(v) if act == 1
 *   v += residual[z][m, n]  (fp32)
 *   out_f32[z][m, n] = v and/or out_bf16[z][m, n] = bf16(v)
 * Replaces nn.Linear at transformer_primitives.py:93-95 (proj_q/k/v), :110 (final), :212-216 (fc1 + GELU, fc2),
 * the residual adds at :287,:292,:397,:401, perceiver.py:179 (final_layer), and — for head sizes the streaming
 * kernel does not cover — the two matmuls of Attention.attend (:138, :163).
 * --------------------------------------------------------------------------------------------------------- */
typedef struct pio_gemm_args {
  const void* A; int64_t lda; int64_t strideA;
  const void* B; int64_t ldb; int64_t strideB;
  int32_t b_mn_major;
  int32_t M, N, K, batch;
  const float* bias; int32_t bias_mode;
  int32_t act;
  float alpha;
  const float* residual; int64_t ldr; int64_t strideR;
  float* out_f32; int64_t ldo32; int64_t strideO32;
  void* out_bf16; int64_t ldo16; int64_t strideO16;
  /* debug / tuning: 0 = default */
  int32_t tile_n;       /* 64, 128, 256 or 0 = auto */
  int32_t max_ctas;     /* 0 = number of SMs */
  int32_t cluster_m;    /* CTAs per cluster along M sharing one multicast B tile: 1, 2, 4 or 0 = auto */
  int32_t kernel;       /* 0 = auto, 1 = single-CTA kernel (128 x tile_n tiles), 2 = CTA-pair kernel (cta_group::2,
                           256 x 256 tiles, TMA epilogue; needs K-major B, exactly one output and 16-byte aligned
                           output / residual rows) */
  /* LayerNorm fused into the projections around it (both kernels, batch == 1; SelfAttention :281 / :292):
   *  - producer side (the GEMM that writes the fp32 residual stream x): additionally writes out_bf16 = bf16(x), the
   *    UN-normalised rows, and stores the partial statistics (sum_n x[m,n], sum_n x[m,n]^2) of every 128-column
   *    half-tile of the row into its own slot: row_stats_out[m][2 * (n / T) + (n % T) / (T / 2)], T = the kernel's tile
   *    width (256 in the CTA-pair kernel, 64 .. 256 in the single-CTA kernel); slots the chosen tile width does not use
   *    are zeroed (plain stores: no atomics, nothing to zero beforehand, bit-reproducible).  The caller sizes the buffer
   *    with pio_gemm_stats_parts(M, N) (or, with a forced tile width, 2 * ceil(N / T)); the value for the CTA-pair
   *    kernel is 4 * ceil(N / 256): the producer of the (hi, lo) stream may split a tile's columns over four warps;
   *  - consumer side (the GEMM that multiplies LN(x) by W): A is that bf16(x), B is W' = W * diag(gamma), and the
   *    epilogue applies the normalisation per output row:
   *        v = rstd_m * (alpha * acc - mean_m * ln_colsum[n]) + bias[n],   ln_colsum[n] = sum_k W'[n, k],
   *        (sum, sumsq) = the row's row_stats_parts partials added in index order,
   *        mean_m = sum / ln_channels, rstd_m = rsqrt(sumsq / ln_channels - mean_m^2 + ln_eps)
   *    (bias must already contain W * beta). */
  float* row_stats_out;       /* [M][row_stats_parts][2] or NULL */
  const float* row_stats_in;  /* [M][row_stats_parts][2] or NULL */
  const float* ln_colsum;     /* [N], required with row_stats_in */
  int32_t ln_channels;
  float ln_eps;
  /* Walk the output tiles from the last M rows to the first (CTA-pair kernel; ignored by the single-CTA kernel).  A
   * consumer that starts where its producer finished finds that producer's last-written rows still in L2. */
  int32_t reverse_tiles;
  int32_t row_stats_parts;    /* partial statistics per row in row_stats_out / row_stats_in (see above) */
  int32_t fp16;               /* 16-bit format of A, B and out_bf16: 0 = bf16, 1 = fp16 */
  /* Residual stream as a PAIR of 16-bit arrays instead of fp32 (CTA-pair kernel only, batch == 1; the producer GEMMs of
   * the latent tower are bound by HBM bytes): value = hi + lo with hi = round16(v),
   * lo = round16(v - hi) — 16 (bf16) or 22 (fp16) significant bits, and hi IS the raw 16-bit copy the fused-LayerNorm
   * consumer multiplies, so the stream costs 4 bytes per element to write instead of 6.
   *  - out_lo16 != NULL (with out_bf16 != NULL and out_f32 == NULL): out_bf16 = hi, out_lo16 = lo, both with pitch ldo16;
   *  - residual_hi16 / residual_lo16 != NULL (with residual == NULL): v += hi + lo, both with pitch ldr16. */
  void* out_lo16;
  const void* residual_hi16;
  const void* residual_lo16;
  int64_t ldr16;
} pio_gemm_args;
int pio_gemm_bf16(const pio_gemm_args* a, void* stream);
/* Number of statistics slots per row that a fused-LayerNorm producer GEMM of shape M x N writes with the automatic
 * kernel / tile choice on the current device (what to pass as row_stats_parts; larger values are allowed, the unused
 * slots are zeroed — but the consumer reads every slot, so do not oversize). */
int pio_gemm_stats_parts(int32_t M, int32_t N);
/* 1 if pio_gemm_bf16 runs an M x N problem (batch 1, tuning fields 0) on the CTA-pair kernel on the current device — the
 * kernel that implements out_lo16 / residual_hi16 / residual_lo16 —, else 0. */
int pio_gemm_pair_kernel(int32_t M, int32_t N);

/* ---------------------------------------------------------------------------------------------------------
 * Row softmax for the materialised attention path: P[b, i, :] = softmax(scale * S[b, i, :]) with masked keys
 * excluded; rows with row_keep[b, i] == 0 are written as zeros.
 * Replaces transformer_primitives.py:146-158 and the wipe at :168-175.
 * --------------------------------------------------------------------------------------------------------- */
typedef struct pio_softmax_args {
  const float* S; int64_t lds; int64_t strideS;   /* [batch, rows, cols] */
  void* P; int64_t ldp; int64_t strideP;          /* bf16 [batch, rows, ldp]; columns cols..ldp-1 := 0 */
  const uint8_t* key_mask; int64_t stride_km;     /* [batch, cols] 1 = attend, or NULL */
  const uint8_t* row_keep; int64_t stride_rk;     /* [batch, rows] 1 = keep, or NULL */
  int32_t batch, rows, cols;
  float scale;
  int32_t split;    /* validation mode: P has ldp = 3 * pad8(cols) columns laid out as [hi | lo | hi] */
  /* General attention arguments of Attention.attend (all optional, NULL = absent); no recipe of the reference passes
   * them, they complete the operator interface:
   *   logits = (S + bias) * scale                    (transformer_primitives.py:143-147: the bias is added BEFORE the scale)
   *   logits[dense_mask == 0] = -1e30                (:149-156; a row with no valid entry becomes uniform)
   *   P_f32 = softmax(logits)                        (:158, what return_matrix hands back at :177-178)
   *   P (bf16, the operand of P.V) = P_f32, but rows with no valid entry (or row_keep == 0) are zeros (:168-175) */
  const uint8_t* dense_mask; int64_t dm_stride_b; int64_t dm_stride_r;                  /* [batch, rows, cols] */
  const float* bias; int64_t bias_stride_b; int64_t bias_stride_r; int64_t bias_stride_c; /* broadcast strides (elements) */
  float* P_f32; int64_t ldpf; int64_t stridePf;                                          /* [batch, rows, cols] */
  int32_t fp16;     /* 16-bit format of P: 0 = bf16, 1 = fp16 (not with split) */
} pio_softmax_args;
int pio_softmax_bf16(const pio_softmax_args* a, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Streaming (flash-style) attention on tcgen05: per (batch, head, 128-query tile[, key split]) the kernel
 * streams key/value tiles with TMA, keeps S and O in TMEM, and applies an online softmax.
 *   Q [B, Nq, H*dqk], K [B, Nk, H*dqk], V [B, Nk, H*dv]  (bf16, heads contiguous inside a row)
 * Output, num_splits == 1:  O_bf16 [B, Nq, H*dv] normalised (rows with row_keep == 0 zeroed).
 * Output, num_splits  > 1 or partial != 0: un-normalised fp32 partial O [splits, B, H, Nq, dv] plus running
 *   max m and sum l [splits, B, H, Nq] (base-e logits), to be merged by pio_attention_combine — the same
 *   contract the multi-GPU key-axis shard uses.
 * Replaces Attention.attend, transformer_primitives.py:117-180, without materialising [B,H,Nq,Nk].
 * Q may have batch stride 0 (the encoder's latent queries are a stride-0 broadcast, position_encoding.py:120).
 * --------------------------------------------------------------------------------------------------------- */
typedef struct pio_attention_args {
  const void* Q; int64_t ldq; int64_t strideQ;
  const void* K; int64_t ldk; int64_t strideK;
  const void* V; int64_t ldv; int64_t strideV;
  int32_t B, H, Nq, Nk, dqk, dv;
  float scale;                                   /* 1/sqrt(dqk) */
  const uint8_t* key_mask; int64_t stride_km;    /* [B, Nk] or NULL */
  const uint8_t* row_keep; int64_t stride_rk;    /* [B, Nq] or NULL */
  void* O; int64_t ldo; int64_t strideO;         /* bf16 [B, Nq, ldo] */
  int32_t num_splits;                            /* key-axis splits inside this GPU (>= 1) */
  int32_t partial;                               /* 1: always emit (O, m, l) partials */
  float* O_part; float* m_part; float* l_part;
  int32_t fp16;                                  /* 16-bit format of Q, K, V, P and O: 0 = bf16, 1 = fp16 */
} pio_attention_args;
int pio_attention_fwd(const pio_attention_args* a, void* stream);
/* 0 if pio_attention_fwd supports these head sizes, PIO_ERR_UNSUPPORTED otherwise (host picks the GEMM path). */
int pio_attention_supported(int32_t dqk, int32_t dv);
/* Key-tile width (64 or 128) the streaming kernel uses for these head sizes, or PIO_ERR_UNSUPPORTED; a key split
 * must give every split at least one tile. */
int pio_attention_key_tile(int32_t dqk, int32_t dv, int32_t same_kv);

/* Merge `parts` partial results (from key splits and/or gathered from other ranks):
 *   O[b, i, h*dv + :] = sum_p O_p * exp(m_p - M) / sum_p l_p * exp(m_p - M),  M = max_p m_p.
 * Part p lives at O_part + p*part_stride_O, m_part/l_part + p*part_stride_ml (0 = densely packed).  Outputs: the
 * normalised bf16 O (if O != NULL) and/or the merged, still un-normalised partial (O_out_part, m_out, l_out) that a
 * rank contributes to the cross-GPU exchange of the key-sharded encoder. */
typedef struct pio_combine_args {
  const float* O_part; const float* m_part; const float* l_part;   /* [parts][B, H, Nq, dv] / [parts][B, H, Nq] */
  int64_t part_stride_O, part_stride_ml;
  int32_t parts, B, H, Nq, dv;
  const uint8_t* row_keep; int64_t stride_rk;
  void* O; int64_t ldo; int64_t strideO;                          /* bf16 [B, Nq, ldo] or NULL */
  float* O_out_part; float* m_out; float* l_out;                  /* fp32 [B, H, Nq, dv] / [B, H, Nq] or NULL */
  /* Fused exchange over NVLink peer memory (the key-sharded encoder on one NVSwitch box): if part_ptrs != NULL it is a
   * DEVICE array of `parts` base pointers, one per rank, each addressing that rank's packed partial
   * [O (rows*dv) | m (rows) | l (rows)] in peer-mapped (symmetric) memory; the kernel then loads every rank's partial
   * straight through NVLink instead of reading a gathered copy, and O_part / m_part / l_part / part_stride_* are
   * ignored.  The caller orders the ranks' writes before this launch (a symmetric-memory barrier on the stream). */
  const float* const* part_ptrs;
  int32_t fp16;                                                   /* 16-bit format of O: 0 = bf16, 1 = fp16 */
  /* optional output: row_alive[b, i] = 1 if query row i of sample b was kept and saw at least one valid key on some
   * part (else the row is wiped, transformer_primitives.py:168-175).  With it the key-sharded encoder needs no separate
   * "does any rank hold a valid key" reduction: the partial sums carry that information. */
  uint8_t* row_alive; int64_t stride_ra;
} pio_combine_args;
int pio_attention_combine(const pio_combine_args* a, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Query-tiled decoder attention (single head): very many output queries attend over a short latent array — the
 * decoder's cross-attend (perceiver.py:166-180 -> CrossAttention.forward, transformer_primitives.py:371-399 ->
 * Attention.attend :117-180); optical flow: 182,528 queries x 2048 latents.
 *   out[b, i, :] = softmax_j(scale * Q[b, i, :] . K[b, j, :]) . V[b, :, :] + bias (+ residual[b, i, :])      (fp32)
 * K and V are distinct Nk x {dqk, dv} matrices (dqk, dv <= 384): the host folds the query projection into K and the
 * output projection into V (K' = k Wq, V' = v Wf^T), so `out` is already the attention block's output
 * incl. `final`, and no [Nq, Nk] matrix is ever written.  Runs on CTA pairs (cta_group::2): each CTA owns 128 queries
 * and stages half of every latent tile; K / V stream from L2 (they are a few MB), Q and out touch HBM once.
 * Rows with row_keep == 0 (or without any valid key) come out as bias (+ residual), as the reference's wiped rows do
 * (:168-175 then :110).  A per-key logit bias can be carried as an extra contraction column (Q column = 1).
 * --------------------------------------------------------------------------------------------------------- */
typedef struct pio_decoder_attention_args {
  const void* Q; int64_t ldq; int64_t strideQ;   /* 16-bit [B, Nq, ldq]; batch stride 0 = shared by every b */
  const void* K; int64_t ldk; int64_t strideK;   /* 16-bit [B, Nk, ldk] */
  const void* V; int64_t ldv; int64_t strideV;   /* 16-bit [B, Nk, ldv] */
  int32_t B, Nq, Nk, dqk, dv;
  float scale;
  const uint8_t* key_mask; int64_t stride_km;    /* [B, Nk] 1 = attend, or NULL */
  const uint8_t* row_keep; int64_t stride_rk;    /* [B, Nq] 1 = keep, or NULL */
  const float* bias;                             /* [dv] or NULL */
  const float* residual; int64_t ldr; int64_t strideR;   /* fp32 [B, Nq, ldr] or NULL */
  float* out; int64_t ldo; int64_t strideO;      /* fp32 [B, Nq, ldo] */
  int32_t fp16;                                  /* 16-bit format of Q, K, V, P and out_ln: 0 = bf16, 1 = fp16 */
  /* optional fused LayerNorm of the output rows (CrossAttention.layer_norm2, transformer_primitives.py:401): if out_ln
   * != NULL the epilogue also writes out_ln[b, i, :] = LayerNorm(out[b, i, :dv]) * ln_gamma + ln_beta as 16-bit rows
   * (columns dv..ld_ln-1 := 0) — the operand of the block's MLP, so no LayerNorm kernel re-reads `out`. */
  void* out_ln; int64_t ld_ln; int64_t stride_ln;
  const float* ln_gamma; const float* ln_beta; float ln_eps;
} pio_decoder_attention_args;
int pio_decoder_attention_fwd(const pio_decoder_attention_args* a, void* stream);
/* 0 if pio_decoder_attention_fwd covers these head sizes, PIO_ERR_UNSUPPORTED otherwise. */
int pio_decoder_attention_supported(int32_t dqk, int32_t dv);

/* ---------------------------------------------------------------------------------------------------------
 * fp32 SIMT linear for very narrow outputs (N <= 16): y[m, :] = x[m, :] . W^T + bias, everything fp32.
 * Replaces the decoder's `final_layer` (perceiver.py:179) when it projects to a handful of channels — the optical
 * flow head is 322 -> 2 and carries most of the model's bf16 error budget, while being 0.01 % of
 * the FLOPs.  HBM-bound: 4*K bytes read per row.
 * --------------------------------------------------------------------------------------------------------- */
typedef struct pio_linear_f32_args {
  const float* x; int64_t ldx;      /* [M, K] */
  const float* w; int64_t ldw;      /* [N, K] (nn.Linear layout) */
  const float* bias;                /* [N] or NULL */
  float* y; int64_t ldy;            /* [M, N] */
  int64_t M;
  int32_t N, K;
  /* optional second operand, y += x2 . w2^T: x2 is a 16-bit [M, K2] matrix (bf16, or fp16 if x2_fp16), w2 fp32 [N, K2].
   * The decoder tail uses it: final_layer(x + fc2(h)) = x . Wfinal^T + h . (Wfinal W2)^T + const, so for a narrow head
   * fc2, its fp32 output array and the separate head all collapse into this one pass (perceiver.py:177-179). */
  const void* x2; int64_t ldx2;
  const float* w2; int64_t ldw2;
  int32_t K2;
  int32_t x2_fp16;
} pio_linear_f32_args;
int pio_linear_f32(const pio_linear_f32_args* a, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Input-side glue fused into the encoder's LayerNorm.  The reference's preprocessors build
 * the encoder input as cat([features [B, N, Cf], broadcast(pos [N, Cp])], -1) (io_processors/preprocessors.py:180-199;
 * the Fourier table is batch-invariant: position_encoding.py:173-183 uses pos[0] only) and CrossAttention then
 * normalises it (transformer_primitives.py:379).  This entry point reads the per-sample features (any strides, e.g.
 * the NCHW image itself) and the [N, Cp] table and writes LayerNorm(cat(...)) as bf16 rows directly: the [B, N, C]
 * fp32 array (3.35 GB for the ImageNet-pixels recipe at batch 64) is never materialised.
 *   y[b*N + n, c] = bf16(((x_c - mean) * rstd) * gamma[c] + beta[c]),  x = [feat[b, n, :Cf] | pos[n, :Cp]],
 *   columns Cf+Cp .. ldy-1 are zeros.  The table part of the statistics is accumulated once per position
 *   (sum and sum of squares), the feature part per row; mean / variance are combined in fp32.
 * Needs N % 4 == 0, Cf + Cp <= 1208, ldy = pad8(Cf + Cp), 16-byte aligned pos / y.
 * --------------------------------------------------------------------------------------------------------- */
typedef struct pio_layernorm_concat_args {
  const float* feat; int64_t feat_stride_b; int64_t feat_stride_n; int64_t feat_stride_c;   /* elements */
  const float* pos;                      /* [N, Cp] contiguous */
  void* y; int64_t ldy;                  /* bf16 [B * N, ldy] contiguous rows */
  const float* gamma; const float* beta; /* [Cf + Cp] or NULL */
  int32_t B, N, Cf, Cp;
  float eps;
  int32_t fp16;                          /* 16-bit format of y: 0 = bf16, 1 = fp16 */
} pio_layernorm_concat_args;
int pio_layernorm_concat_bf16(const pio_layernorm_concat_args* a, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Order-independent 128-bit content hash of a device buffer of 32-bit words: out2[0..1] += hash (the caller zeroes
 * out2; hashes of several buffers accumulate when given different seeds).  The encode-once latent cache of
 * PerceiverEncoder (the reference's chunked decoders re-run the encoder on identical inputs,
 * multimodal_perceiver.py:146-161) keys on it instead of keeping and comparing a private copy of the input array.
 * HBM-bound: 4 bytes read per word.  data must be 16-byte aligned.
 * --------------------------------------------------------------------------------------------------------- */
int pio_hash_words(const void* data, int64_t n_words, uint64_t seed, uint64_t* out2, void* stream);

/* Per-launch device timing (bench.py's roofline): while enabled, every entry point brackets its kernel launch with
 * CUDA events on the launching stream.  pio_profile_read drains the records into out[family*4 + {ms, flops, bytes,
 * launches}] for the families {0 layernorm, 1 gemm, 2 softmax, 3 attention, 4 combine, 5 linear_f32}.  Do not enable while a
 * stream is being captured into a CUDA graph. */
void pio_profile_enable(int on);
int pio_profile_read(double* out, int n_families);

/* Number of kernels launched by this library in the calling process (bench.py's gpu_launches). */
int64_t pio_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* PIO_B200_H_ */
