/* fcd_b200.h — C ABI of libfcd_b200.so, the B200 (sm_100a) kernels behind the FCD-GAN hot path.
 *
 * The reference (Cwuwhu/FCD-GAN-pytorch) has NO native code and NO FFI: its
 * "operator interface" for this path is the set of torch library calls made from Module.py, Loss.py
 * and ssim.py.  Each entry point below therefore cites the reference call site(s) whose arithmetic it
 * replaces.  INTEGRATION.md shows the ctypes binding a maintainer would add on the reference side.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; sizes are element counts;
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous on that stream and
 *     re-entrant per (device, stream); no entry point allocates device memory;
 *   - return value: 0 (FCD_OK) or an FCD_ERR_* code; fcd_last_error() returns the text for the
 *     calling host thread.  Nothing aborts or throws.
 *   - tensors inside the path are NHWC with an explicit pixel pitch `*_ld` (elements between two
 *     consecutive pixels) so that channel slices of a concatenation buffer are addressed in place;
 *   - a "split" tensor is two bf16 planes (hi, lo): value = hi + lo.  Convolution inputs are split,
 *     convolution outputs are fp32.  lo == NULL selects single-plane bf16 ("fast")
 *     arithmetic.
 */
#ifndef FCD_B200_H
#define FCD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FCD_OK 0
#define FCD_ERR_ARG 1     /* bad shape / unsupported configuration            */
#define FCD_ERR_CUDA 2    /* a CUDA runtime / driver call failed              */
#define FCD_ERR_UNSUPPORTED 3

#define FCD_ACT_NONE 0
#define FCD_ACT_RELU 1    /* Module.py:28,31                                   */
#define FCD_ACT_PRELU 2   /* Module.py:147,179 (one shared slope)              */
#define FCD_ACT_LEAKY 3   /* Module.py:197,201,205,209,215 (0.2)               */

#define FCD_ENGINE_AUTO 0
#define FCD_ENGINE_SIMT 1   /* fp32 CUDA-core implicit GEMM (any shape)         */
#define FCD_ENGINE_TC 2     /* tcgen05 + TMA implicit GEMM                      */

#define FCD_LOSS_L1 0     /* Loss.py:69  (CNetLoss)                            */
#define FCD_LOSS_MSE 1    /* Loss.py:103 (CGeneratorLoss)                      */

const char* fcd_last_error(void);
int fcd_version(void);
/* 1 if the tcgen05 engine can serve this convolution, else 0 (then AUTO uses SIMT). */
int fcd_conv2d_tc_supported(int Cin_p, int Cout_p, int KH, int KW, int stride);

/* ---- weight packing ------------------------------------------------------------------------
 * Packs an OIHW fp32 torch weight (Module.py:26,29,146,155,158,177,180,196-207) into the split-bf16
 * layout the conv kernels read: [tap][rows][cols] with cols contiguous, zero padded.
 *   mode 0 (forward): tap = r*KW+s,                 rows = Cout_p, cols = Cin_p
 *   mode 1 (dgrad)  : tap = (KH-1-r)*KW+(KW-1-s),   rows = Cin_p,  cols = Cout_p
 */
int fcd_pack_conv_weight(const float* w_oihw, int Cout, int Cin, int KH, int KW, int Cout_p, int Cin_p, int mode,
                         void* w_hi, void* w_lo, void* stream);

/* ---- convolution (replaces nn.Conv2d forward, Module.py:26-223) ------------------------------
 * z[n,oh,ow,co] = bias[co] + sum_{r,s,ci} x[n, oh*stride+r-pad, ow*stride+s-pad, ci] * w[r*KW+s][co][ci]
 * x: split NHWC (N,H,W,Cin_p) pitch x_ld; w: packed mode-0 weights; z: fp32 NHWC (N,OH,OW,Cout_p).
 * With stride 1, dgrad of a convolution is this same call on dz with the mode-1 weights and
 * pad' = K-1-pad.  If stat_sum/stat_sqsum are non-NULL the per-channel sum and sum of squares of z
 * (double[Cout_p], the BatchNorm2d batch statistics of Module.py:27,30,156,178,181,200,204,208) are
 * ACCUMULATED into them.
 */
int fcd_conv2d_fwd(const void* x_hi, const void* x_lo, int x_ld, const void* w_hi, const void* w_lo,
                   const float* bias, float* z, int z_ld, int N, int H, int W, int Cin_p, int Cout_p, int KH,
                   int KW, int stride, int pad, double* stat_sum, double* stat_sqsum, int engine, void* stream);

/* dgrad for stride-2 convolutions (Module.py:196-207 backward w.r.t. the input):
 * dx[n,h,w,ci] = sum_{r,s,co : (h+pad-r)%2==0,...} dz[n,(h+pad-r)/2,(w+pad-s)/2,co] * w[r*KW+s][co][ci]
 * w are the FORWARD (mode 0) packed weights. */
int fcd_conv2d_dgrad_strided(const void* dz_hi, const void* dz_lo, int dz_ld, const void* w_hi, const void* w_lo,
                             float* dx, int dx_ld, int N, int H, int W, int Cin_p, int Cout_p, int KH, int KW,
                             int stride, int pad, void* stream);

/* wgrad (nn.Conv2d backward w.r.t. weight and bias):
 * dw[co][ci][r][s] (+)= sum_{n,oh,ow} dz[n,oh,ow,co] * x[n,oh*stride+r-pad,ow*stride+s-pad,ci]   (OIHW fp32,
 * logical Cout x Cin);  db[co] (+)= sum dz.  `accumulate` = 0 overwrites, 1 adds.  `workspace` is
 * caller-owned scratch of at least fcd_conv2d_wgrad_workspace(...) bytes. */
size_t fcd_conv2d_wgrad_workspace(int N, int H, int W, int Cin_p, int Cout_p, int KH, int KW, int stride, int pad,
                                  int engine);
int fcd_conv2d_wgrad(const void* x_hi, const void* x_lo, int x_ld, const void* dz_hi, const void* dz_lo,
                     int dz_ld, float* dw_oihw, float* db, int N, int H, int W, int Cin, int Cin_p, int Cout,
                     int Cout_p, int KH, int KW, int stride, int pad, int accumulate, void* workspace,
                     size_t workspace_bytes, int engine, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FCD_B200_H */
