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

const char* sst_version(void);
const char* sst_last_error(void);
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
} SstGemmDesc;

int sst_gemm(const SstGemmDesc* d, const void* A, const void* B, void* C, const void* bias /*fp32[N]*/,
             const void* aux, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SST_H_ */
