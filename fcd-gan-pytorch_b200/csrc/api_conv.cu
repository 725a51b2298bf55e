// api_conv.cu — C-ABI entry points for the convolution family + library-wide error plumbing.
#include <stdarg.h>
#include <string.h>

#include "fcd_common.cuh"

namespace fcd {

static thread_local char g_err[512] = "";

void set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    int n = snprintf(g_err, sizeof(g_err), "[fcd_b200 err %d] ", code);
    vsnprintf(g_err + n, sizeof(g_err) - n, fmt, ap);
    va_end(ap);
}

int current_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    return (dev >= 0 && dev < FCD_MAX_DEVICES) ? dev : 0;
}

// per device: a process may drive several GPUs (the reference pins one script per GPU, SURVEY.md §2.1, but nothing here
// relies on it)
int sm_count() {
    static int n[FCD_MAX_DEVICES] = {};
    const int dev = current_device();
    if (n[dev] == 0) {
        if (cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n[dev] <= 0) n[dev] = 148;
    }
    return n[dev];
}

// implemented in conv_tc.cu / conv_simt.cu / wgrad_tc.cu
bool conv_tc_supported(int Cin_p, int Cout_p, int KH, int KW, int stride);
int conv2d_fwd_tc(const void*, const void*, int, const void*, const void*, const float*, const float*, int, float*, int,
                  int, int, int, int, int, int, int, int, int, double*, double*, cudaStream_t);
int conv2d_dgrad_strided_tc(const void*, const void*, int, const void*, const void*, const float*, int, float*, int, int, int,
                            int, int, int, int, int, int, int, cudaStream_t);
int conv2d_fwd_simt(const void*, const void*, int, const void*, const void*, const float*, const float*, int, float*,
                    int, int, int, int, int, int, int, int, int, int, double*, double*, cudaStream_t);
int conv2d_dgrad_strided_simt(const void*, const void*, int, const void*, const void*, const float*, int, float*, int,
                              int, int, int, int, int, int, int, int, int, cudaStream_t);
int conv2d_wgrad_simt(const void*, const void*, int, const void*, const void*, int, float*, float*, int, int, int, int,
                      int, int, int, int, int, int, int, int, cudaStream_t);
int pack_conv_weight(const float*, int, int, int, int, int, int, int, void*, void*, cudaStream_t);
bool wgrad_tc_supported(int Cin_p, int Cout_p, int KH, int KW, int stride);
size_t wgrad_tc_workspace(int N, int H, int W, int Cin_p, int Cout_p, int KH, int KW, int pad);
int conv2d_wgrad_tc(const void*, const void*, int, const void*, const void*, int, float*, int, int, int, int, int, int,
                    int, int, int, int, int, int, void*, size_t, cudaStream_t);
int conv2d_taps_tc(const void*, const void*, int, int, int, const void*, const void*, const float*, const float*, int, float*,
                   int, int, int, int, int, int, int, int, int, int, int, int, int, int, cudaStream_t);
int conv2d_taps_wgrad_tc(const void*, const void*, int, int, int, const void*, const void*, int, int, int, float*, int, int, int,
                         int, int, int, int, int, int, int, int, void*, size_t, cudaStream_t);
extern bool g_wgrad_halo_enabled;
extern bool g_conv_halo_enabled;
extern bool g_conv_tma_out;
extern int g_conv_halo_slots;
int channel_sum_split(const void* hi, const void* lo, int ld, long long npix, int C, float* out, int accumulate,
                      cudaStream_t stream);

}  // namespace fcd

using namespace fcd;

extern "C" {

const char* fcd_last_error(void) { return g_err; }
int fcd_version(void) { return 100; }

int fcd_set_option(const char* name, int value) {
    FCD_CHECK_ARG(name, "fcd_set_option: null name");
    if (strcmp(name, "wgrad_halo") == 0) {
        g_wgrad_halo_enabled = value != 0;
        return FCD_OK;
    }
    if (strcmp(name, "conv_tma_out") == 0) {
        g_conv_tma_out = value != 0;
        return FCD_OK;
    }
    if (strcmp(name, "conv_halo_slots") == 0) {
        g_conv_halo_slots = value;
        return FCD_OK;
    }
    if (strcmp(name, "conv_halo") == 0) {
        g_conv_halo_enabled = value != 0;
        return FCD_OK;
    }
    set_error(FCD_ERR_ARG, "fcd_set_option: unknown option '%s'", name);
    return FCD_ERR_ARG;
}

int fcd_conv2d_tc_supported(int Cin_p, int Cout_p, int KH, int KW, int stride) {
    return conv_tc_supported(Cin_p, Cout_p, KH, KW, stride) ? 1 : 0;
}

int fcd_pack_conv_weight(const float* w_oihw, int Cout, int Cin, int KH, int KW, int Cout_p, int Cin_p, int mode,
                         void* w_hi, void* w_lo, void* stream) {
    FCD_CHECK_ARG(w_oihw && w_hi, "fcd_pack_conv_weight: null pointer");
    FCD_CHECK_ARG(Cout_p >= Cout && Cin_p >= Cin && (mode == 0 || mode == 1), "fcd_pack_conv_weight: bad dims/mode");
    return pack_conv_weight(w_oihw, Cout, Cin, KH, KW, Cout_p, Cin_p, mode, w_hi, w_lo, as_stream(stream));
}

int fcd_conv2d_fwd(const void* x_hi, const void* x_lo, int x_ld, const void* w_hi, const void* w_lo,
                   const float* bias, const float* addend, int addend_ld, float* z, int z_ld, int N, int H, int W, int Cin_p, int Cout_p, int KH, int KW,
                   int stride, int pad, double* stat_sum, double* stat_sqsum, int engine, void* stream) {
    FCD_CHECK_ARG(x_hi && w_hi && z, "fcd_conv2d_fwd: null pointer");
    FCD_CHECK_ARG(N > 0 && H > 0 && W > 0 && Cin_p > 0 && Cout_p > 0 && KH > 0 && KW > 0 && stride > 0 && pad >= 0,
                  "fcd_conv2d_fwd: bad dims N=%d H=%d W=%d Cin_p=%d Cout_p=%d", N, H, W, Cin_p, Cout_p);
    FCD_CHECK_ARG((stat_sum == nullptr) == (stat_sqsum == nullptr), "fcd_conv2d_fwd: give both stat buffers or none");
    const bool tc_ok = conv_tc_supported(Cin_p, Cout_p, KH, KW, stride);
    if (engine == FCD_ENGINE_TC && !tc_ok) {
        set_error(FCD_ERR_UNSUPPORTED, "fcd_conv2d_fwd: tcgen05 engine does not take Cin_p=%d Cout_p=%d k=%dx%d s=%d",
                  Cin_p, Cout_p, KH, KW, stride);
        return FCD_ERR_UNSUPPORTED;
    }
    if (engine == FCD_ENGINE_TC || (engine == FCD_ENGINE_AUTO && tc_ok))
        return conv2d_fwd_tc(x_hi, x_lo, x_ld, w_hi, w_lo, bias, addend, addend_ld, z, z_ld, N, H, W, Cin_p, Cout_p, KH,
                             KW, stride, pad, stat_sum, stat_sqsum, as_stream(stream));
    return conv2d_fwd_simt(x_hi, x_lo, x_ld, w_hi, w_lo, bias, addend, addend_ld, z, z_ld, N, H, W, Cin_p, Cout_p, KH,
                           KW, stride, pad, stat_sum, stat_sqsum, as_stream(stream));
}

int fcd_conv2d_dgrad_strided(const void* dz_hi, const void* dz_lo, int dz_ld, const void* w_hi, const void* w_lo,
                             const void* wT_hi, const void* wT_lo, const float* addend, int addend_ld, float* dx, int dx_ld,
                             int N, int H, int W, int Cin_p, int Cout_p, int KH, int KW, int stride, int pad, int engine,
                             void* stream) {
    FCD_CHECK_ARG(dz_hi && dx && (w_hi || wT_hi), "fcd_conv2d_dgrad_strided: null pointer");
    const bool tc_ok = wT_hi != nullptr && conv_tc_supported(Cout_p, Cin_p, KH, KW, stride) && stride <= KH && stride <= KW;
    if (engine == FCD_ENGINE_TC && !tc_ok) {
        set_error(FCD_ERR_UNSUPPORTED, "fcd_conv2d_dgrad_strided: tcgen05 engine needs mode-1 weights and 64-multiple channels");
        return FCD_ERR_UNSUPPORTED;
    }
    if (engine == FCD_ENGINE_TC || (engine == FCD_ENGINE_AUTO && tc_ok))
        return conv2d_dgrad_strided_tc(dz_hi, dz_lo, dz_ld, wT_hi, wT_lo, addend, addend_ld, dx, dx_ld, N, H, W, Cin_p, Cout_p,
                                       KH, KW, stride, pad, as_stream(stream));
    FCD_CHECK_ARG(w_hi, "fcd_conv2d_dgrad_strided: the SIMT engine needs the mode-0 packed weights");
    return conv2d_dgrad_strided_simt(dz_hi, dz_lo, dz_ld, w_hi, w_lo, addend, addend_ld, dx, dx_ld, N, H, W, Cin_p, Cout_p, KH, KW, stride,
                                     pad, as_stream(stream));
}

size_t fcd_conv2d_wgrad_workspace(int N, int H, int W, int Cin_p, int Cout_p, int KH, int KW, int stride, int pad,
                                  int engine) {
    if (engine != FCD_ENGINE_SIMT && wgrad_tc_supported(Cin_p, Cout_p, KH, KW, stride))
        return wgrad_tc_workspace(N, H, W, Cin_p, Cout_p, KH, KW, pad);
    return 0;
}

int fcd_conv2d_wgrad(const void* x_hi, const void* x_lo, int x_ld, const void* dz_hi, const void* dz_lo, int dz_ld,
                     float* dw_oihw, float* db, int N, int H, int W, int Cin, int Cin_p, int Cout, int Cout_p, int KH,
                     int KW, int stride, int pad, int accumulate, void* workspace, size_t workspace_bytes, int engine,
                     void* stream) {
    FCD_CHECK_ARG(x_hi && dz_hi && dw_oihw, "fcd_conv2d_wgrad: null pointer");
    const bool tc_ok = wgrad_tc_supported(Cin_p, Cout_p, KH, KW, stride);
    if (engine == FCD_ENGINE_TC && !tc_ok) {
        set_error(FCD_ERR_UNSUPPORTED, "fcd_conv2d_wgrad: tcgen05 engine does not take this shape");
        return FCD_ERR_UNSUPPORTED;
    }
    if (engine == FCD_ENGINE_TC || (engine == FCD_ENGINE_AUTO && tc_ok)) {
        int rc = conv2d_wgrad_tc(x_hi, x_lo, x_ld, dz_hi, dz_lo, dz_ld, dw_oihw, N, H, W, Cin, Cin_p, Cout, Cout_p, KH,
                                 KW, stride, pad, accumulate, workspace, workspace_bytes, as_stream(stream));
        if (rc) return rc;
        if (db) {
            const int OH = (H + 2 * pad - KH) / stride + 1, OW = (W + 2 * pad - KW) / stride + 1;
            return channel_sum_split(dz_hi, dz_lo, dz_ld, 1LL * N * OH * OW, Cout, db, accumulate, as_stream(stream));
        }
        return FCD_OK;
    }
    return conv2d_wgrad_simt(x_hi, x_lo, x_ld, dz_hi, dz_lo, dz_ld, dw_oihw, db, N, H, W, Cin, Cin_p, Cout, Cout_p, KH,
                             KW, stride, pad, accumulate, as_stream(stream));
}

int fcd_conv2d_taps_fwd(const void* x_hi, const void* x_lo, int x_ld, int XH, int XW, const void* w_hi, const void* w_lo,
                        const float* bias, const float* addend, int addend_ld, float* z, int z_ld, int N, int OH, int OW,
                        int Cin_p, int Cout_p, int n_r, int n_s, int dh0, int dh_step, int dw0, int dw_step, int csh, int csw,
                        void* stream) {
    FCD_CHECK_ARG(x_hi && w_hi && z, "fcd_conv2d_taps_fwd: null pointer");
    return conv2d_taps_tc(x_hi, x_lo, x_ld, XH, XW, w_hi, w_lo, bias, addend, addend_ld, z, z_ld, N, OH, OW, Cin_p, Cout_p, n_r,
                          n_s, dh0, dh_step, dw0, dw_step, csh, csw, as_stream(stream));
}

size_t fcd_conv2d_taps_wgrad_workspace(int Cin_p, int Cout_p, int n_r, int n_s) {
    return sizeof(float) * static_cast<size_t>(n_r) * n_s * Cin_p * Cout_p;
}

int fcd_conv2d_taps_wgrad(const void* x_hi, const void* x_lo, int x_ld, int XH, int XW, const void* dz_hi, const void* dz_lo,
                          int dz_ld, int GH, int GW, float* dw, float* db, int N, int Cin, int Cin_p, int Cout, int Cout_p,
                          int n_r, int n_s, int dh0, int dw0, int dw_step, int accumulate, void* workspace,
                          size_t workspace_bytes, void* stream) {
    FCD_CHECK_ARG(x_hi && dz_hi && dw, "fcd_conv2d_taps_wgrad: null pointer");
    int rc = conv2d_taps_wgrad_tc(x_hi, x_lo, x_ld, XH, XW, dz_hi, dz_lo, dz_ld, GH, GW, dw, N, Cin, Cin_p, Cout, Cout_p, n_r, n_s,
                                  dh0, dw0, dw_step, accumulate, workspace, workspace_bytes, as_stream(stream));
    if (rc) return rc;
    if (db) return channel_sum_split(dz_hi, dz_lo, dz_ld, 1LL * N * GH * GW, Cout, db, accumulate, as_stream(stream));
    return FCD_OK;
}

}  // extern "C"
