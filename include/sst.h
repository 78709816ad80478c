/* libsst.so -- C ABI of the B200-native Silent Speech Transformer hot path.
 *
 * The reference (ChristianSquadro/EMG-based-Speech-Recognition-with-heterogenous-data) is pure PyTorch and has
 * no FFI of its own; the seam it offers is the nn.Module API of speech_recognition/architecture.py:51-188.  Each
 * entry point below replaces the torch ops issued by the reference lines cited next to it (paths relative to
 * /root/reference/speech_recognition/).  Conventions:
 *   - plain pointers + sizes, no torch types; every pointer is a DEVICE pointer unless stated otherwise;
 *   - the caller owns all memory (outputs, saved statistics, workspaces); the library never allocates or
 *     retains device memory;
 *   - every call only ENQUEUES work on `stream` (a cudaStream_t passed as void*), no device synchronisation;
 *   - return 0 on success, a negative SST_E_* code otherwise; message via sst_last_error() (thread local);
 *   - dtype arguments are SST_F32 (parity mode: CUDA-core fp32 kernels) or SST_BF16 (speed mode: tcgen05 /
 *     tensor-core kernels, fp32 accumulation and fp32 saved statistics).
 */
#ifndef SST_H_
#define SST_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SST_F32 0
#define SST_BF16 1

#define SST_OK 0
#define SST_E_ARG (-1)      /* bad shape / alignment / flag combination */
#define SST_E_ARCH (-2)     /* device is not sm_100 */
#define SST_E_LAUNCH (-3)   /* CUDA launch / runtime failure */
#define SST_E_UNSUPPORTED (-4)
#define SST_E_COMM (-5)     /* NCCL could not be loaded / a collective failed */

const char* sst_version(void);
const char* sst_last_error(void);
/* number of CUDA kernels this library has launched in this process (bench.py's gpu_launches) */
long long sst_launch_count(void);
/* 0 when the current device is a Blackwell sm_100 part, SST_E_ARCH otherwise. */
int sst_device_check(void);

/* ------------------------------------------------------------------------------------------------------------
 * GEMM family.  Replaces: nn.Linear (architecture.py:59,70,71; transformer.py:36,38,95,97), the per-head einsum
 * projections (transformer.py:172-174,209), nn.Conv1d k=3/k=1 as implicit GEMM over channels-last activations
 * (architecture.py:26,28,32) and every autograd-generated dgrad / wgrad of those.
 *
 *   layout SST_GEMM_TN :  C[m,n] = sum_k A(m,k) * B[n*ldb + k]
 *        A(m,k) = Amat[(m + a_row_shift[s]) * lda + a_col0[s] + (k - s*Kseg)],  s = k / Kseg,  Kseg = K / n_seg
 *        (rows outside [0, a_rows) read as zero).  n_seg = 3 expresses the three taps of a k=3 convolution over a
 *        time-padded channels-last activation; n_seg = 1 is a plain GEMM.
 *   layout SST_GEMM_TN_BMN (input gradients):  as TN, but B is read as a (K, N) row-major matrix, B(k,n) = Bmat[k*ldb + n]:
 *        dx = dy . W multiplies by the weight in ITS OWN (N_out, K_in) layout -- no transposed copy of W is ever built.
 *   layout SST_GEMM_NT_MN (weight gradients):  C[m,n] (+)= sum_k A[k*lda + m] * B(k,n)
 *        B(k,n) = Bmat[(k + b_row_shift[s]) * ldb + b_col0[s] + (n - s*Nseg)],  s = n / Nseg,  Nseg = N / n_seg
 *        (rows outside [0, b_rows) read as zero).
 *   epilogue, in this order:  v = alpha*acc;  BIAS: v += bias[n];  RELU: v = max(v,0);
 *        DROPOUT: v = keep(seed, m*N+n) ? v/(1-p) : 0;   MULMASK: v *= (aux[m*ldaux+n] > 0 ? mask_scale : 0);
 *        ACCUM: v += C_old;   then stored as out_dtype.
 *   row remap (conv outputs): with remap_P > 0 the virtual row m maps to chunk = m / P, t = m % P - remap_j0 and is
 *        stored at row chunk*remap_T + t iff 0 <= t < remap_T (other rows are dropped).
 *   dtype SST_BF16 runs on tcgen05 tensor cores (TMA-fed, TMEM accumulators); SST_F32 on CUDA cores.
 * ---------------------------------------------------------------------------------------------------------- */
#define SST_GEMM_TN 0
#define SST_GEMM_NT_MN 1
#define SST_GEMM_TN_BMN 2

#define SST_EPI_BIAS 1
#define SST_EPI_RELU 2
#define SST_EPI_DROPOUT 4
#define SST_EPI_MULMASK 8
#define SST_EPI_ACCUM 16

typedef struct SstGemmDesc {
  int32_t dtype;       /* element type of A and B */
  int32_t out_dtype;   /* element type of C (and of C_old for ACCUM) */
  int32_t aux_dtype;   /* element type of aux (MULMASK) */
  int32_t layout;
  int64_t M, N, K;
  int64_t lda, ldb, ldc, ldaux;
  int64_t a_rows, a_cols; /* TN: extent of Amat (rows, columns) ;  NT_MN: rows = K extent, cols >= M */
  int64_t b_rows, b_cols; /* NT_MN: extent of Bmat ; TN: rows = N, cols = K */
  int32_t n_seg;
  int32_t a_row_shift[3];
  int32_t a_col0[3];
  int32_t b_row_shift[3];
  int32_t b_col0[3];
  int32_t epilogue;    /* SST_EPI_* bits */
  float alpha;
  float mask_scale;
  float drop_p;
  uint64_t seed;
  int32_t remap_P, remap_T, remap_j0;
  int32_t force_simt;  /* debugging / cross-check: run the CUDA-core kernel even for bf16 */
  /* Segmented output columns (fp32 outputs; 0 = plain row-major C).  Column n of the result lands at
   *   C + out_grp_off[n / out_grp_cols] + ((n % out_grp_cols) / out_seg_cols) * out_seg_stride + m * ldc + n % out_seg_cols
   * (element offsets).  With out_seg_cols = dh, out_seg_stride = D*dh, ldc = dh, out_grp_cols = H*dh this writes the
   * weight gradient of the fused per-head projections straight into the reference's (H, D, dh) parameter layout of
   * w_q / w_k / w_v (transformer.py:146-149), one group per tensor -- no packed temporary, no permute.
   * out_seg_cols and out_grp_cols must be multiples of 32, N a multiple of 32. */
  int64_t out_seg_cols, out_seg_stride, out_grp_cols, out_grp_off[3];
  /* Column accumulators fused into the epilogue (tensor-core path, bf16 C stored through the staged path: layout TN / TN_BMN,
   * N % 32 == 0, ldc % 8 == 0, no ACCUM, no segmented output).  Over the rows that are actually stored, of the values AS
   * STORED (rounded to bf16) -- what a separate pass over C would read:
   *   col_acc_mode 1:  ((float*)col_acc)[n] += sum_m C[m,n]            the bias gradient of the layer whose output gradient C is
   *                    (replaces sst_colsum_accum over C: linear1.bias, transformer.py:61)
   *   col_acc_mode 2:  ((double*)col_acc)[g*2G + n%G] += sum_m C[m,n],  [g*2G + G + n%G] += sum_m C[m,n]^2,  g = n / G,
   *                    G = col_acc_grp (0: N): BatchNorm batch statistics of a conv output in sst_colstats' layout per group of G
   *                    channels (architecture.py:27,29,33); the call zeroes the 2*N doubles first. */
  void* col_acc;
  int32_t col_acc_mode, col_acc_grp;
} SstGemmDesc;

int sst_gemm(const SstGemmDesc* d, const void* A, const void* B, void* C, const void* bias /*fp32[N]*/,
             const void* aux, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Normalisation (HBM-bound, 128-bit vectorised; all statistics fp32/fp64).
 *  sst_layernorm_fwd/bwd : y = LN(x + dropout(r)) * gamma + beta, eps 1e-5, post-LN residual form of
 *      transformer.py:59-60,62-63,124-133.  `s_out` receives the pre-norm sum (may alias r); backward takes it
 *      back as `s`, returns ds (gradient of both x and the undropped r) and, when drop_p > 0, dr = ds*keep/(1-p).
 *      dgamma/dbeta are ACCUMULATED (+=) in fp32.
 *  sst_colstats / sst_bn_finalize / sst_bn_apply / sst_bn_bwd : training-mode nn.BatchNorm1d of ResBlock
 *      (architecture.py:27,29,33,40-48) over channels-last conv outputs (rows = chunk*T + t, pitch ld):
 *      per-channel sum / sum of squares in double, mean / invstd (+ running-stat update, momentum 0.1, unbiased
 *      variance), then out = relu?(bnA(xa) [+ bnB(xb)]) written into the time-padded layout
 *      (chunk, lead + t, C) with zero halo rows that the next implicit-GEMM convolution reads.
 * ---------------------------------------------------------------------------------------------------------- */
int sst_layernorm_fwd(int dtype, int64_t rows, int D, const void* x, const void* r, float drop_p, uint64_t seed,
                      const float* gamma, const float* beta, void* y, void* s_out, float* mean, float* rstd, float eps,
                      void* stream);
int sst_layernorm_bwd(int dtype, int64_t rows, int D, const void* dy, const void* s, const float* mean, const float* rstd,
                      const float* gamma, void* ds, void* dr, float drop_p, uint64_t seed, float* dgamma, float* dbeta,
                      void* stream);
/*  sst_gelu_dropout_fwd/bwd : the exact (erf) GELU variant of the feed-forward activation (north_star "GELU FFN layers"; the code
 *      as shipped uses ReLU, transformer.py:45,61 -- ReLU stays the default and rides in the GEMM epilogue).
 *      y = keep ? gelu(x)/(1-p) : 0 ; dx = dy * gelu'(x) * keep/(1-p), keep = philox_keep16(seed, row*cols + col); x is the
 *      saved PRE-activation.  dx may alias dy.  cols and pitches multiples of 8. */
int sst_gelu_dropout_fwd(int dtype, int64_t rows, int cols, const void* x, int64_t ldx, float drop_p, uint64_t seed, void* y, int64_t ldy,
                         void* stream);
int sst_gelu_dropout_bwd(int dtype, int64_t rows, int cols, const void* dy, int64_t lddy, const void* x, int64_t ldx, float drop_p,
                         uint64_t seed, void* dx, int64_t lddx, void* stream);
int sst_colstats(int dtype, const void* x, int64_t rows, int C, int64_t ld, double* stats /*[2*C]*/, void* stream);
/* out[c] += sum_rows x[r, c]  (bias gradients; fp32 accumulate) */
int sst_colsum_accum(int dtype, const void* x, int64_t rows, int C, int64_t ld, float* out, void* stream);
int sst_bn_finalize(const double* stats, int64_t count, int C, float eps, float momentum, float* mean, float* invstd,
                    float* running_mean, float* running_var, int training, void* stream);
int sst_bn_apply(int dtype, int64_t n_chunks, int T, int C, const void* xa, int64_t lda, const float* mean_a,
                 const float* invstd_a, const float* gamma_a, const float* beta_a, const void* xb, int64_t ldb,
                 const float* mean_b, const float* invstd_b, const float* gamma_b, const float* beta_b, int relu, void* out,
                 int lead, int trail, void* stream);
/* y may be NULL: the ReLU mask is then recomputed from xa / xb and the BN constants (bit-identical sign, one tensor less
 * to read in each of the two passes); beta_* are only used for that. */
int sst_bn_bwd(int dtype, int64_t n_chunks, int T, int C, const void* dout, int64_t ld_dout, const void* y, int y_lead,
               int y_trail, int relu, const void* xa, int64_t lda, const float* mean_a, const float* invstd_a,
               const float* gamma_a, const float* beta_a, void* dxa, int64_t ld_dxa, int lead_a, int trail_a,
               float* dgamma_a, float* dbeta_a, const void* xb, int64_t ldb, const float* mean_b, const float* invstd_b,
               const float* gamma_b, const float* beta_b, void* dxb, int64_t ld_dxb, int lead_b, int trail_b,
               float* dgamma_b, float* dbeta_b, double* red /*[3*C]*/, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Attention.  Replaces MultiHeadAttention.forward between the projections (transformer.py:177-208) together
 * with LearnedRelativePositionalEmbedding (transformer.py:260-403): logits = mask(q.k * scale) + relpos(q),
 * softmax, dropout on the probabilities, probs.v -- without materialising (B,H,L,L).
 *   q/k/v/o are token matrices: row = b*L + t (pitch ld*; or q_off[b] + t / k_off[b] + t, see the descriptor), head h
 *   occupies columns [h*dh, (h+1)*dh).
 *   masks are SET to -1e8 exactly as masked_fill does: causal (j > i), keys j >= k_lens[b] (or k_pad), and, with
 *   mask_q_rows, whole rows i >= q_lens[b] (or q_pad); fully masked rows therefore give the reference's uniform softmax.
 *   rel_dist R > 0 adds bias[i][j] = q_i . E[h][j-i+R-1] for |j-i| < R and -1e8 otherwise; E is
 *   (H, 2R-1, dh) in the compute dtype and receives no gradient (Q2).
 *   lse is float[2*B*H*Lq]: row maximum at [(b*H+h)*Lq + i] and log(sum exp(logit - max)) at [B*H*Lq + ...] (kept apart
 *   because a fully masked row has max = -1e8, where max + log(sum) would round the sum away); saved for backward.
 *   dtype SST_F32: CUDA-core kernels; SST_BF16: tensor-core flash kernels.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct SstAttnDesc {
  int32_t dtype;
  int32_t B, H, Lq, Lk, dh;
  int64_t ldq, ldk, ldv, ldo;
  int32_t causal;
  int32_t mask_q_rows;
  int32_t rel_dist;
  float scale;
  float drop_p;
  uint64_t seed;
  int32_t force_simt;
  /* optional per-position padding masks (device uint8, nonzero = padded), OR-ed with the length masks: q_pad[b*Lq + i]
   * masks the whole query row, k_pad[b*Lk + j] the key -- for padding that is not a suffix (a PAD token generated in the
   * middle of a greedy prefix, greedy_search.py:21 + architecture.py:174) */
  const uint8_t* q_pad;
  const uint8_t* k_pad;
  /* optional PACKED (variable-length) layouts, SURVEY.md 8(f) N2/N4 -- device int64[B], NULL = the padded layout b*Lq / b*Lk:
   *   q_off[b] = row of batch entry b's first query in q / o / dO / dq; entry b owns rows [q_off[b], q_off[b] + q_lens[b])
   *              (q_lens required, every length >= 1).  Rows i >= q_lens[b] do not exist: nothing is read for or written to
   *              them (they are the next entry's rows), Lq is only the maximum length (tile grid, lse / delta index space).
   *   k_off[b] = the same for k / v / dk / dv with k_lens (required).  Keys j >= k_lens[b] get probability exactly 0, as the
   *              masked padding keys of the padded layout do for every row that has one real key.
   *   Entries may share key rows in FORWARD calls (k_off[b] = 0 for all b: the hypotheses of a beam search attend to one
   *   encoder memory, BeamSearch.py:111 without the memory.repeat); backward needs disjoint ranges. */
  const int64_t* q_off;
  const int64_t* k_off;
  int64_t q_rows_total, k_rows_total;   /* rows the packed q-side / k-side matrices hold (bounds of the TMA descriptors) */
} SstAttnDesc;

int sst_attn_fwd(const SstAttnDesc* d, const void* q, const void* k, const void* v, const void* E, const int32_t* q_lens,
                 const int32_t* k_lens, void* o, float* lse, void* stream);
/* dq/dk/dv use the pitches of q/k/v; delta is a caller-provided float[B*H*Lq] scratch; ws a caller-provided scratch of
 * sst_attn_bwd_workspace_bytes(d) bytes (16-byte aligned; 0 bytes / NULL allowed for the CUDA-core path): the tensor-core
 * path hands the recomputed probabilities and logit gradients of every (query tile, key tile) pair from its dQ kernel to
 * its dK/dV kernel through it instead of recomputing them. */
size_t sst_attn_bwd_workspace_bytes(const SstAttnDesc* d);
int sst_attn_bwd(const SstAttnDesc* d, const void* q, const void* k, const void* v, const void* E, const int32_t* q_lens,
                 const int32_t* k_lens, const void* o, const float* lse, const void* dO, void* dq, void* dk, void* dv,
                 float* delta, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Losses (recognition_model.py:93-107).
 *  sst_ctc_loss : log_softmax over C classes + CTC (blank = `blank`, reduction 'mean' = mean_b nll_b / max(len_b,1),
 *      zero_infinity False).  Per utterance one warp runs the alpha and one the beta recursion in log space (lattice
 *      states spread over the lanes, the utterance's log-probs staged in shared memory); the gradient is written
 *      directly w.r.t. the raw logits (softmax - occupancy), scaled by gcoef, zero for t >= in_lens[b].
 *      Workspaces: lp_ws float[B*L*C], alpha_ws float[2*B*L*(2*Smax+1)] (alpha then beta lattice), nll float[B].
 *      targets: int64 (B, Smax), Smax <= 255 (the lattice of 2*Smax+1 states is spread over one warp's registers; the
 *      reference accepts any length, its corpus has <= ~120 phones per utterance): longer targets return SST_E_ARG.
 *  sst_ce_sumexp_loss : LabelSmoothingLoss.py:13-15 -- (1-eps)*CE(ignore_index, mean over kept) + eps/S*sum(exp(logits)),
 *      the sum running over every row including ignored ones; S = target length; gradient scaled by gcoef.
 *      row_ws float[2*rows].
 * ---------------------------------------------------------------------------------------------------------- */
int sst_ctc_loss(int logits_dtype, int grad_dtype, int B, int L, int C, int blank, const void* logits, int64_t ld,
                 const int64_t* targets, int Smax, const int32_t* in_lens, const int32_t* tgt_lens, float gcoef, float* lp_ws,
                 float* alpha_ws, float* nll, void* grad, int64_t ldg, float* loss_out, void* stream);
/* CTC best-path decode (BASELINE.json config 5): arg-max per frame, merge repeats, drop `blank`.  logits (B*L, ld);
 * out_ids int32 (B, L) padded with -1, out_lens int32[B].  The reference has no CTC decode; the oracle is
 * torch.argmax + collapse on the reference's w_aux logits. */
int sst_ctc_greedy(int logits_dtype, int B, int L, int C, int blank, const void* logits, int64_t ld, const int32_t* in_lens,
                   int32_t* out_ids, int32_t* out_lens, void* stream);
/* One step of the reference's greedy attention-decoder search on the device (greedy_search.py:22-37: softmax/argmax of the last
 * position, append, stop test): tokens[b*tok_stride_b + pos*tok_stride_pos] = arg-max_c logits[b*row_stride + c] (lowest index wins
 * ties, as torch.argmax), done[b] |= (token == eos), n_done[0] = number of samples with done set -- the only thing the host has
 * to look at. */
int sst_greedy_pick(int logits_dtype, int B, int C, const void* logits, int64_t row_stride, int64_t* tokens, int64_t tok_stride_b,
                    int64_t tok_stride_pos, int pos, int eos, uint8_t* done, int32_t* n_done, void* stream);
int sst_ce_sumexp_loss(int logits_dtype, int grad_dtype, int64_t rows, int S, int C, const void* logits, int64_t ld,
                       const int64_t* target, int ignore, float eps, int64_t n_valid, float gcoef, float* row_ws, void* grad,
                       int64_t ldg, float* loss_out, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Data movement and optimiser.
 *  sst_shift_left       : x[:, :-r] = x[:, r:]; x[:, -r:] = 0 per chunk, in place        (architecture.py:104-108)
 *  sst_im2col_first     : (n, Tin, 8) fp32 EMG -> (n*Tin/2, 32) patches [3 taps x 8 | centre tap x 8] so that the first
 *                         ResBlock's conv1 (k3,s2) and residual_path (k1,s2) are one GEMM (architecture.py:26,32,55)
 *  sst_gather_rows_pad  : decollate_tensor + pad_sequence(padding_value=fill)             (data_utils.py:176-185,
 *                         architecture.py:116-117);  sst_scatter_rows is its adjoint (din pre-zeroed)
 *  sst_embed_posenc_fwd : embedding_tgt(y) + pe[b]/D, dropout                             (architecture.py:126-127, Q10)
 *  sst_embed_bwd        : dW[y] += dout, skipping padding_idx rows (fp32 atomics)
 *  sst_permute3_cast    : out[i*o0+j*o1+k*o2] (+)= in[i*s0+j*s1+k*s2] with dtype conversion (weight packing / grad unpacking)
 *  sst_permute3_plan / sst_permute3_cast_batch : the same for a whole TABLE of permutes in one launch (the weight repack
 *                         after an optimizer step is ~230 of them).  The host fills in, out, dims, strides, dtypes and accumulate of
 *                         every item, sst_permute3_plan (host-only, no GPU work) assigns mode/first_block/nblocks and the
 *                         grid size, the table is copied to the device once and replayed while the pointers stay valid
 *  sst_adamw            : torch.optim.AdamW step over a flat fp32 buffer, `step` 1-based  (recognition_model.py:293);
 *                         p_bf16 (nullable): same-shape bf16 buffer that receives the updated parameters (the GEMM operands)
 * ---------------------------------------------------------------------------------------------------------- */
int sst_shift_left(float* x, int64_t n_chunks, int T, int Cc, int r, void* stream);
int sst_im2col_first(int out_dtype, const float* x, void* col, int64_t n_chunks, int Tin, void* stream);
int sst_gather_rows_pad(int dtype, const void* in, void* out, const int64_t* offs, const int32_t* lens, int B, int Lmax, int D,
                        float fill, void* stream);
int sst_scatter_rows(int dtype, const void* dout, void* din, const int64_t* offs, const int32_t* lens, int B, int Lmax, int D,
                     void* stream);
int sst_embed_posenc_fwd(int out_dtype, const int64_t* y, const float* W, const float* pe, void* out, int B, int S, int D,
                         float drop_p, uint64_t seed, void* stream);
int sst_embed_bwd(int dtype, const int64_t* y, const void* dout, float* dW, int B, int S, int D, int pad_idx, float drop_p,
                  uint64_t seed, void* stream);
int sst_permute3_cast(int in_dtype, int out_dtype, const void* in, void* out, int64_t d0, int64_t d1, int64_t d2, int64_t s0,
                      int64_t s1, int64_t s2, int64_t o0, int64_t o1, int64_t o2, int accumulate, void* stream);
typedef struct SstPermuteItem {
  const void* in; void* out;
  int64_t d0, d1, d2, s0, s1, s2, o0, o1, o2;
  int32_t in_dtype, out_dtype, accumulate;
  int32_t mode, first_block, nblocks;          /* filled by sst_permute3_plan */
} SstPermuteItem;
int sst_permute3_plan(SstPermuteItem* items_host, int n_items, int* total_blocks);
int sst_permute3_cast_batch(const SstPermuteItem* items_dev, int n_items, int total_blocks, void* stream);
/* ------------------------------------------------------------------------------------------------------------
 * Gradient exchange of the data-parallel path: replaces nn.DataParallel's gather (recognition_model.py:284).
 * One process per GPU; rank 0 draws a 128-byte id (sst_comm_unique_id) and hands it to every rank over any host channel;
 * sst_comm_init (collective, uses the calling thread's current device) returns an opaque communicator;
 * sst_comm_allreduce_bucket reduces `count` elements of a gradient bucket IN PLACE across the ranks (sum, or mean when
 * `average`) -- it only enqueues on `stream`, so the caller orders it after the backward stage that completes the bucket and
 * before the optimizer (sst_b200/train.py GradSync).  NCCL is dlopen'ed at the first call: `lib_path` (or the environment
 * variable SST_NCCL_LIB, or the loader's libnccl.so.2) names the library, normally the one the host framework already uses.
 * sst_comm_nccl_version: e.g. 22809, or a negative SST_E_* when NCCL cannot be loaded.
 * ---------------------------------------------------------------------------------------------------------- */
int sst_comm_unique_id(void* id_out /* 128 bytes, host */, const char* lib_path);
int sst_comm_init(const void* id /* 128 bytes, host */, int rank, int world, const char* lib_path, void** comm_out);
int sst_comm_allreduce_bucket(void* comm, void* buf, int64_t count, int dtype, int average, void* stream);
int sst_comm_destroy(void* comm);
int sst_comm_nccl_version(const char* lib_path);

/*  sst_set_dropout_salt : register (NULL: unregister) a device uint64 every dropout-drawing kernel launched afterwards adds to its
 *                         seed.  Seeds, the learning rate and Adam's bias corrections cross this ABI by value, which a captured
 *                         CUDA graph replays unchanged: with a salt registered, sst_write_scalars in front
 *                         of each replay (salt) and sst_adamw_dev (lr, 1 - beta1^t, sqrt(1 - beta2^t) from a device float[3])
 *                         make a replayed training step draw fresh masks and take the scheduled step.  Process-wide.
 *  sst_write_scalars    : nbytes (4..64, multiple of 4) from HOST memory to device memory as a kernel parameter -- captured by
 *                         value when the call returns, ordered on `stream` like any launch. */
int sst_set_dropout_salt(const uint64_t* salt_dev);
int sst_write_scalars(void* dst_dev, const void* src_host, int nbytes, void* stream);
int sst_adamw_dev(float* p, const float* g, float* m, float* v, int64_t n, const float* hyper_dev, float beta1, float beta2,
                  float eps, float wd, void* p_bf16, void* stream);
int sst_adamw(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps, float wd,
              int64_t step, void* p_bf16, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SST_H_ */
