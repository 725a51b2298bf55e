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

#define FCD_RASTER_F32 0  /* element types of a source raster (fcd_tiles_gather)  */
#define FCD_RASTER_U16 1
#define FCD_RASTER_I16 2
#define FCD_RASTER_U8 3

#define FCD_LOSS_L1 0     /* Loss.py:69  (CNetLoss)                            */
#define FCD_LOSS_MSE 1    /* Loss.py:103 (CGeneratorLoss)                      */

const char* fcd_last_error(void);
int fcd_version(void);
/* Library-wide switches for A/B measurements: "wgrad_halo" (default 1) = halo-reuse weight-gradient kernel; "conv_halo" (1) =
 * halo-reuse forward / dgrad kernel; "conv_tma_out" (1) = TMA-store epilogue; "conv_halo_slots" (0 = automatic) = cap of the halo
 * kernel's input-plane ring (shared memory it does not take deepens the epilogue's output staging ring). */
int fcd_set_option(const char* name, int value);
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
 * ACCUMULATED into them.  `addend` (optional, fp32 NHWC, pitch addend_ld) is added to z in the epilogue:
 * it carries the identity branch of a residual block (Module.py:190) through dgrad and sums the
 * gradients of an activation that has two consumers.
 */
int fcd_conv2d_fwd(const void* x_hi, const void* x_lo, int x_ld, const void* w_hi, const void* w_lo,
                   const float* bias, const float* addend, int addend_ld, float* z, int z_ld, int N, int H, int W, int Cin_p, int Cout_p, int KH,
                   int KW, int stride, int pad, double* stat_sum, double* stat_sqsum, int engine, void* stream);

/* dgrad for stride-2 convolutions (Module.py:196-207 backward w.r.t. the input):
 * dx[n,h,w,ci] = sum_{r,s : (h+pad-r)%stride==0,...} dz[n,(h+pad-r)/stride,(w+pad-s)/stride,co] * w[r][s][co][ci]
 * w = FORWARD (mode 0) packed weights (SIMT engine), wT = mode-1 packed weights (tcgen05 engine: the input pixels are
 * split into stride x stride parity classes, each a stride-1 implicit GEMM over its subset of taps).  Either may be
 * NULL; engine AUTO prefers tcgen05 when wT is given and the channel counts are multiples of 64. */
int fcd_conv2d_dgrad_strided(const void* dz_hi, const void* dz_lo, int dz_ld, const void* w_hi, const void* w_lo,
                             const void* wT_hi, const void* wT_lo, const float* addend, int addend_ld, float* dx, int dx_ld,
                             int N, int H, int W, int Cin_p, int Cout_p, int KH, int KW, int stride, int pad, int engine,
                             void* stream);

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


/* ---- generic tap-list convolution (tcgen05 engine only; channel counts multiples of 64) ----------------------
 * z[n,oh,ow,:] = bias + addend + sum_{i<n_r, j<n_s} x[n, oh*csh + dh0 + i*dh_step, ow*csw + dw0 + j*dw_step, :] . w[i*n_s+j]
 * with x (N, XH, XW, Cin_p) split, out-of-bounds pixels = 0, w packed [n_r*n_s][Cout_p][Cin_p] (fcd_pack_conv_weight
 * mode 0 of an OIHW tensor with KH = n_r, KW = n_s).  csh / csw > 1 read every csh-th / csw-th pixel (TMA element
 * strides).  This is how the 13-band layers of the Generator (Module.py:146,158) and the Segmentor's first conv
 * (Module.py:26) run: four adjacent pixels are packed into the 64-wide channel axis (fcd_stage_nchw_to_split_pack4),
 * which cuts a 9x9 filter row from 9 taps to 3. */
int fcd_conv2d_taps_fwd(const void* x_hi, const void* x_lo, int x_ld, int XH, int XW, const void* w_hi, const void* w_lo,
                        const float* bias, const float* addend, int addend_ld, float* z, int z_ld, int N, int OH, int OW,
                        int Cin_p, int Cout_p, int n_r, int n_s, int dh0, int dh_step, int dw0, int dw_step, int csh, int csw,
                        void* stream);
/* dw[co][ci][r][s] (+)= sum_{n, h<GH, w<GW} x[n, h + r + dh0, w + s*dw_step + dw0, ci] * dz[n,h,w,co]  (fp32 OIHW with
 * KH = n_r, KW = n_s, logical Cout x Cin);  db[co] (+)= sum dz. */
size_t fcd_conv2d_taps_wgrad_workspace(int Cin_p, int Cout_p, int n_r, int n_s);
int fcd_conv2d_taps_wgrad(const void* x_hi, const void* x_lo, int x_ld, int XH, int XW, const void* dz_hi, const void* dz_lo,
                          int dz_ld, int GH, int GW, float* dw, float* db, int N, int Cin, int Cin_p, int Cout, int Cout_p,
                          int n_r, int n_s, int dh0, int dw0, int dw_step, int accumulate, void* workspace,
                          size_t workspace_bytes, void* stream);
/* NCHW fp32 (N, C <= 16, H, W) -> split NHWC (N, H, W+M, 64): dst[n,h,w'',j*16+c] = src[n,c,h,w''-M+j], j = 0..3. */
int fcd_stage_nchw_to_split_pack4(const float* src, int N, int C, int H, int W, int M, void* dst_hi, void* dst_lo, void* stream);
/* NCHW fp32 (N, C, H, W) -> split NHWC (N, H, W+M, Kp): dst[n,h,w'',j*C+c] = src[n,c,h,w''-M+j], j < P, zero for k >= P*C.
 * P = the filter width makes a whole filter row ONE tap (9 pixels x 13 bands = 117 -> Kp = 128 for Module.py:146,158). */
int fcd_stage_nchw_to_split_rowpack(const float* src, int N, int C, int H, int W, int M, int P, int Kp, void* dst_hi,
                                    void* dst_lo, void* stream);

/* ==== HBM-bound glue between the convolutions (elementwise.cu) =====================================
 * "split" outputs are conv operands (bf16 hi/lo planes); fp32 NHWC tensors are conv results / gradients. */

/* `.to(device)` boundary (Demo_USSS.py:147-148): NCHW fp32 (N,C,H,W) -> split NHWC with Cp >= C zero-padded
 * channels.  `mask` (optional, (N,1,H,W)) applies x*(1-mask): the soft masking of Demo_RSSS.py:290-291. */
int fcd_stage_nchw_to_split(const float* src, const float* mask, int N, int C, int H, int W, void* dst_hi, void* dst_lo,
                            int dst_ld, int Cp, void* stream);
/* NCHW fp32 -> split NHWC (N, OH, OW, Kp) im2col rows of a 3x3 / stride 2 / pad 1 convolution (Module.py:196, the first
 * discriminator layer): dst[n,oh,ow,(r*3+s)*C + c] = src[n,c,2oh-1+r,2ow-1+s], zero outside the image and for k >= 9*C. */
int fcd_stage_im2col3x3s2(const float* src, int N, int C, int H, int W, void* dst_hi, void* dst_lo, int Kp, void* stream);
/* fp32 NHWC (pitch src_ld) -> NCHW fp32; accumulate != 0 adds into dst. */
int fcd_unstage_f32_to_nchw(const float* src, int src_ld, int N, int C, int H, int W, float* dst, int accumulate,
                            void* stream);
/* split NHWC -> NCHW fp32 and NCHW fp32 -> fp32 NHWC (zero padded to Cp): the boundaries of the stand-alone
 * building blocks DoubleConv / Down / Up / ResidualBlock (Module.py:18-79,174-190). */
int fcd_unstage_split_to_nchw(const void* src_hi, const void* src_lo, int src_ld, int N, int C, int H, int W, float* dst,
                               void* stream);
int fcd_stage_nchw_to_f32(const float* src, int N, int C, int H, int W, float* dst, int dst_ld, int Cp, void* stream);
int fcd_f32_to_split(const float* src, int src_ld, void* dst_hi, void* dst_lo, int dst_ld, long long npix, int Cp,
                     void* stream);
/* dst[pix][c] += src[pix][c] on fp32 NHWC tensors (gradient of a tensor with two consumers, Module.py:168,190). */
int fcd_add_f32(float* dst, int dst_ld, const float* src, int src_ld, long long npix, int Cp, void* stream);

/* nn.BatchNorm2d (Module.py:27,30,156,178,181,200,204,208).  fcd_bn_stats ACCUMULATES per-channel sum / sum of
 * squares of z into double[Cp] buffers (the tcgen05 conv epilogue can produce the same numbers);
 * fcd_bn_finalize turns them (training != 0: batch statistics, biased variance, running stats updated with
 * the unbiased one and `momentum`) or the running statistics (training == 0) into y = z*scale + shift. */
int fcd_bn_stats(const float* z, int z_ld, long long npix, int Cp, double* sum, double* sqsum, void* stream);
int fcd_bn_finalize(const double* sum, const double* sqsum, double count, const float* gamma, const float* beta,
                    float* running_mean, float* running_var, int C, int Cp, float momentum, float eps, int training,
                    float* scale, float* shift, float* mean, float* invstd, void* stream);
/* out = act(z*scale + shift) [+ residual]  (scale == NULL: no BN).  ReLU Module.py:28,31; PReLU 147,179
 * (slope_ptr -> the learnable slope); LeakyReLU(0.2) 197-215 (slope_const); residual add 168,190. */
int fcd_bn_act_fwd(const float* z, int z_ld, const float* scale, const float* shift, int act, const float* slope_ptr,
                   float slope_const, const void* res_hi, const void* res_lo, int res_ld, void* out_hi, void* out_lo,
                   int out_ld, long long npix, int Cp, void* stream);
/* backward of the above in three steps: per-channel reductions (s1 = sum dy, s2 = sum dy*xhat, dslope),
 * finalize (c1, c2, dgamma, dbeta, dslope), apply (dz written split = operand of dgrad / wgrad). */
int fcd_bn_act_bwd_reduce(const float* da, int da_ld, const float* z, int z_ld, const float* scale, const float* shift,
                          const float* mean, const float* invstd, int act, const float* slope_ptr, float slope_const,
                          long long npix, int Cp, double* s1, double* s2, double* dslope, void* stream);
/* global_s1 / global_s2 / global_count (optional, SyncBN — no reference counterpart): the sums over ALL ranks'
 * batches; c1, c2 (the means that enter dz) then come from them, the parameter gradients still from the local sums.
 * dbias (optional): gradient of the producing convolution's bias in closed form from the same sums,
 * scale * (s1 - count*c1 - c2 * sum xhat) with sum xhat = (zsum - count*mean) * invstd (scale null: s1); (+)= when
 * dbias_accumulate. */
int fcd_bn_bwd_finalize(const double* s1, const double* s2, double count, int training, int C, int Cp, float* c1,
                        float* c2, float* dgamma, float* dbeta, int accumulate, const double* ds, float* dslope,
                        const float* scale, const double* zsum, const float* mean, const float* invstd, float* dbias,
                        int dbias_accumulate, const double* global_s1, const double* global_s2, double global_count,
                        void* stream);
int fcd_bn_act_bwd_apply(const float* da, int da_ld, const float* z, int z_ld, const float* scale, const float* shift,
                         const float* mean, const float* invstd, const float* c1, const float* c2, int act,
                         const float* slope_ptr, float slope_const, void* dz_hi, void* dz_lo, int dz_ld, long long npix,
                         int Cp, void* stream);

/* nn.MaxPool2d(2) (Module.py:44), floor semantics; backward routes to the first maximum like torch. */
int fcd_maxpool2_fwd(const void* in_hi, const void* in_lo, int in_ld, int N, int H, int W, int Cp, void* out_hi,
                     void* out_lo, int out_ld, void* stream);
int fcd_maxpool2_bwd(const float* d_out, int dout_ld, const void* in_hi, const void* in_lo, int in_ld, int N, int H,
                     int W, int Cp, float* d_in, int din_ld, int accumulate, void* stream);
/* nn.Upsample(x2, bilinear, align_corners=True) + F.pad to the skip size + torch.cat (Module.py:60,70-78):
 * writes the (N,H,W) channel slice of the concat buffer, zero outside [pad_top, pad_top+2h) x [pad_left, ..). */
int fcd_upsample2x_bilinear_fwd(const void* in_hi, const void* in_lo, int in_ld, int N, int h, int w, int Cp,
                                void* out_hi, void* out_lo, int out_ld, int H, int W, int pad_top, int pad_left,
                                void* stream);
int fcd_upsample2x_bilinear_bwd(const float* d_out, int dout_ld, int N, int h, int w, int Cp, int H, int W, int pad_top,
                                int pad_left, float* d_in, int din_ld, void* stream);
/* nn.ConvTranspose2d(k=2, s=2) (Module.py:63) = four 1x1 convolutions (fcd_conv2d_fwd) + these pixel shuffles. */
int fcd_convT2x2_shuffle_fwd(const float* src, long long plane_stride, int src_ld, int N, int h, int w, int Cp,
                             void* out_hi, void* out_lo, int out_ld, int H, int W, int pad_top, int pad_left,
                             void* stream);
int fcd_convT2x2_shuffle_bwd(const float* d_out, int dout_ld, int N, int h, int w, int Cp, int H, int W, int pad_top,
                             int pad_left, void* g_hi, void* g_lo, long long plane_stride, int g_ld, void* stream);

/* ==== heads and inline masking arithmetic (heads.cu) ================================================ */
/* OutConv: Conv2d 1x1 (Cin -> n_out <= 4) + Sigmoid -> NCHW fp32 change-density map (Module.py:82-90). */
int fcd_outconv_sigmoid_fwd(const void* x_hi, const void* x_lo, int x_ld, int Cin, const float* w, const float* b,
                            int n_out, int N, int H, int W, float* out_nchw, void* stream);
int fcd_outconv_sigmoid_bwd(const float* dout_nchw, const float* out_nchw, const void* x_hi, const void* x_lo, int x_ld,
                            int Cin, const float* w, int n_out, int N, int H, int W, float* dx, int dx_ld, float* dw,
                            float* db, int accumulate, double* scratch, void* stream);
/* Discriminator head (Module.py:212-223): pooled = AdaptiveAvgPool2d(1)(fx - fy); dense layers with
 * act 0 none / 3 LeakyReLU(0.2) / 4 sigmoid. */
int fcd_gap_diff_fwd(const void* x_hi, const void* x_lo, const void* y_hi, const void* y_lo, int ld, int N, int HW, int C,
                     float* pooled, void* stream);
int fcd_gap_diff_bwd(const float* dpooled, int N, int HW, int C, float* dfx, float* dfy, int ld, void* stream);
int fcd_fc_fwd(const float* in, const float* w, const float* b, int N, int I, int O, int act, float* pre, float* out,
               void* stream);
int fcd_fc_bwd(const float* dout, const float* pre, const float* out, const float* in, const float* w, int N, int I, int O,
               int act, float* du, float* din, float* dw, float* db, int accumulate, void* stream);
/* out = (a*(1-region) + b*region) * (1 - mask)  on NCHW fp32 (Demo_RSSS.py:290-300, Demo_WSSS.py:261-279);
 * region/b optional.  Backward w.r.t. mask only (a, b are data). */
int fcd_mask_fwd(const float* a, const float* b, const float* region, const float* mask, int N, int C, int H, int W,
                 float* out, void* stream);
int fcd_mask_bwd(const float* dout, const float* a, const float* b, const float* region, int N, int C, int H, int W,
                 float* dmask, int accumulate, void* stream);

/* ==== loss stack (losses.cu): NCHW fp32 boundary tensors ============================================ */
/* Masked reconstruction loss of CNetLoss (kind FCD_LOSS_L1, Loss.py:76-87) / CGeneratorLoss (FCD_LOSS_MSE,
 * Loss.py:109-119, samples with sum(1-cmap) == 0 skipped):  out2[0] = generator_loss, out2[1] = mean|cmap|.
 * sums: double[3*B] scratch kept for the backward (numerator_i, sum_p (1-cmap_i), sum_p |cmap_i|).
 * tm, gm (optional): the masked images t*(1-cmap), g*(1-cmap) consumed by MS-SSIM (Loss.py:78-79,93). */
int fcd_masked_recon_fwd(const float* t, const float* g, const float* cmap, int B, int C, int H, int W, int kind,
                         double* sums, float* out2, float* tm, float* gm, void* stream);
/* g_gen / g_l1: device scalars d(total)/d(generator_loss), d(total)/d(l1_loss) (NULL = 0); dtm, dgm: optional
 * gradients w.r.t. the masked images; dt, dg, dcmap: outputs (each optional). */
int fcd_masked_recon_bwd(const float* t, const float* g, const float* cmap, int B, int C, int H, int W, int kind,
                         const double* sums, const float* g_gen, const float* g_l1, const float* dtm, const float* dgm,
                         float* dt, float* dg, float* dcmap, void* stream);
/* region_loss (Loss.py:127-141): n = elements per sample, HW = pixels per sample; sums double[2*B]. */
int fcd_region_loss_fwd(const float* cmap, const float* region, int B, long long n, long long HW, int kind, double* sums,
                        float* out, void* stream);
int fcd_region_loss_bwd(const float* cmap, const float* region, int B, long long n, long long HW, int kind,
                        const double* sums, const float* gout, float* dcmap, void* stream);
/* mean(x) / mean|x| / mean(x^2) (mode 0/1/2): WGAN terms Demo_RSSS.py:304,324, nc_loss Demo_WSSS.py:299, l1 315. */
int fcd_mean_fwd(const float* x, long long n, int mode, double* acc, float* out, void* stream);
int fcd_mean_bwd(const float* x, long long n, int mode, const float* gout, float* dx, void* stream);
/* One SSIM level (ssim.py:55-92) on planes = B*C images of H x W: separable valid Gaussian blur (win: device
 * float[win_size], win_size odd <= 11) of X, Y, X^2, Y^2, XY; sums double[2*planes] receives the spatial SUMS of
 * the ssim map and the cs map; dmaps (optional) float[5*planes*OH*OW] receives d(which)/d(mu1,mu2,e11,e22,e12)
 * (which: 0 = cs, 1 = ssim) for the backward kernel. */
int fcd_ssim_level_fwd(const float* X, const float* Y, int planes, int H, int W, const float* win, int win_size, float C1,
                       float C2, double* sums, float* dmaps, int which, void* stream);
/* dX, dY (+)= coef[plane] * adjoint-blur(dmaps) chain rule (autograd of ssim.py:79-91). */
int fcd_ssim_level_bwd(const float* dmaps, const float* X, const float* Y, int planes, int H, int W, const float* win,
                       int win_size, const float* coef, float* dX, float* dY, int accumulate, void* stream);
/* F.avg_pool2d(kernel_size=2, padding=size%2) between MS-SSIM levels (ssim.py:215-216). */
int fcd_avgpool2_fwd(const float* in, int planes, int H, int W, int pad_h, int pad_w, float* out, void* stream);
int fcd_avgpool2_bwd(const float* dout, int planes, int H, int W, int pad_h, int pad_w, float* din, int accumulate,
                     void* stream);
/* relu / weighted geometric mean over levels / mean (ssim.py:143-150, 218-225).  sums: double[levels][2][planes];
 * counts: device double[levels] (blurred pixels per level); prod float[planes]; out: 1 (size_average) or B values;
 * coef float[levels][planes] = d out / d (spatial sum of the level's map). */
int fcd_msssim_combine_fwd(const double* sums, const double* counts, const float* weights, int levels, int planes, int C,
                           int size_average, int use_relu, float* prod, float* out, void* stream);
int fcd_msssim_combine_bwd(const double* sums, const double* counts, const float* weights, int levels, int planes, int C,
                           int size_average, int use_relu, const float* prod, const float* gout, float* coef, void* stream);

/* Feature-space MSE of PerceptionLoss (Loss.py:36,48,59).  The VGG16 feature stack (Loss.py:25, torchvision vgg16.features[0..29]:
 * thirteen 3x3 convolutions + ReLU + four 2x2 max-pools — fcd_conv2d_fwd / fcd_bn_act_fwd with scale == NULL / fcd_maxpool2_fwd)
 * runs on ONE batch whose first half holds the masked target images and whose second half the masked generated images; f is a
 * split NHWC feature tensor of 2*npix pixels, `half_elems` = element offset of the second half.  fwd: *acc += sum (a - b)^2
 * (double, caller zero-initialises).  bwd: grad[first half] (+)= g*(a - b), grad[second half] (+)= -g*(a - b) with
 * g = *gout * scale (scale = 2 * weight / element count), fp32 NHWC with pitch grad_ld. */
int fcd_mse_halves_fwd(const void* f_hi, const void* f_lo, int f_ld, long long half_elems, long long npix, int Cp, double* acc,
                       void* stream);
int fcd_mse_halves_bwd(const void* f_hi, const void* f_lo, int f_ld, long long half_elems, long long npix, int Cp,
                       const float* gout, float scale, float* grad, int grad_ld, long long grad_half_elems, int accumulate,
                       void* stream);

/* ---- rasters either side of the hot path ------------------------------------------------------
 * All pointers are device pointers.  geom is int[n_tiles][6], one row per tile of the grid, computed once on the host
 * (fcdgan_b200.raster.TileGrid = the geometry of GDALDataset.slice_assign, data_utils.py:151-176); items is int[B], the tile
 * index of every batch slot (null = tiles 0..B-1).
 *
 * fcd_tiles_gather: GDALDataset.__getitem__ (data_utils.py:94-123) + NORMALIZE.forward (CommonFunc.py:208-224).
 *   raster [C][H][W] of `dtype`; geom[t] = {read_x, read_y, read_w, read_h, write_x, write_y}; out [B][C][patch_h][patch_w]
 *   fp32 = 0 outside the write window, float((double(v) - mean[c]) / std[c]) inside (mean/std null: no normalisation). */
int fcd_tiles_gather(const void* raster, int dtype, int C, int H, int W, const int* geom, const int* items, int B,
                     int patch_w, int patch_h, const double* mean, const double* stdv, float* out, void* stream);
/* Dataset_mean / Dataset_std reductions (CommonFunc.py:436-499) over un-normalised tiles: valid pixels = fp32 band sum of x
 * non-zero.  centre null: sums[b][0|1][c] += sum of x|y, counts[b] += #valid; centre = [meanX | meanY]: sums of squared
 * deviations (counts may be null).  sums / counts must be zero-initialised. */
int fcd_tiles_moments(const float* x_tiles, const float* y_tiles, int B, int C, int npix_per_tile, const double* centre,
                      double* sums, long long* counts, void* stream);
/* GDALDataset.GDALwriteDefault (data_utils.py:178-213): geom[t] = {pad_x, pad_y, slice_x, slice_y, slice_w, slice_h};
 * tiles [B][1][patch_h][patch_w] -> raster [H][W]. */
int fcd_tiles_scatter(const float* tiles, const int* geom, const int* items, int B, int patch_w, int patch_h, float* raster,
                      int H, int W, void* stream);
/* Demo_USSS.py:349-362 + Evaluator._generate_matrix_bymap (metrics.py:74-80), same geom as fcd_tiles_scatter:
 * counts[i*2+j] += #{centre-crop pixels : int16(ref) == gt_map[i] && (cmap > thresh) == pre_map[j]}; counts is int64[4]. */
int fcd_confusion_accumulate(const float* cmap, const float* ref, const int* geom, const int* items, int B, int patch_w,
                             int patch_h, float thresh, int gt0, int gt1, int pre0, int pre1, long long* counts, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FCD_B200_H */
