// elementwise.cu — the HBM-bound glue between convolutions: staging NCHW<->NHWC, BatchNorm
// finalize / apply+activation (forward and backward), 2x2 max-pool, bilinear x2 up-sampling, the
// pixel-shuffle halves of ConvTranspose2d(k2,s2).  All kernels stream 8 channels (16 B of bf16 per
// plane / 32 B of fp32) per thread with the channel index fastest, so warps touch whole 128-byte lines.
//
// Reference call sites: nn.BatchNorm2d Module.py:27,30,156,178,181,200,204,208; ReLU/PReLU/LeakyReLU
// Module.py:28,31,147,179,197-209; MaxPool2d Module.py:44; Upsample(bilinear, align_corners=True)
// Module.py:60; F.pad + torch.cat Module.py:73-78; ConvTranspose2d Module.py:63.
#include "fcd_common.cuh"

namespace fcd {
namespace {

constexpr int NT = 256;
constexpr long long IDX32_MAX = (1LL << 32) - 2 * NT;   // element counts below this index with 32-bit arithmetic

struct F8 {
    float v[8];
};
__device__ __forceinline__ F8 ld_f32x8(const float* p) {
    F8 r;
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
__device__ __forceinline__ void st_f32x8(float* p, const F8& r) {
    *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(r.v[4], r.v[5], r.v[6], r.v[7]);
}
__device__ __forceinline__ F8 ld_bf16x8(const __nv_bfloat16* p) {
    F8 r;
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(h[i]);
        r.v[2 * i] = f.x;
        r.v[2 * i + 1] = f.y;
    }
    return r;
}
__device__ __forceinline__ F8 ld_split8(const __nv_bfloat16* hi, const __nv_bfloat16* lo, size_t off) {
    F8 r = ld_bf16x8(hi + off);
    if (lo) {
        const F8 l = ld_bf16x8(lo + off);
#pragma unroll
        for (int i = 0; i < 8; ++i) r.v[i] += l.v[i];
    }
    return r;
}
__device__ __forceinline__ void st_split8(__nv_bfloat16* hi, __nv_bfloat16* lo, size_t off, const F8& r) {
    uint4 uh, ul;
    __nv_bfloat162* ph = reinterpret_cast<__nv_bfloat162*>(&uh);
    __nv_bfloat162* pl = reinterpret_cast<__nv_bfloat162*>(&ul);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __nv_bfloat16 h0 = __float2bfloat16_rn(r.v[2 * i]), h1 = __float2bfloat16_rn(r.v[2 * i + 1]);
        ph[i] = __halves2bfloat162(h0, h1);
        pl[i] = __halves2bfloat162(__float2bfloat16_rn(r.v[2 * i] - __bfloat162float(h0)),
                                   __float2bfloat16_rn(r.v[2 * i + 1] - __bfloat162float(h1)));
    }
    *reinterpret_cast<uint4*>(hi + off) = uh;
    if (lo) *reinterpret_cast<uint4*>(lo + off) = ul;
}

// ---------------------------------------------------------------------------------------------
// staging
// ---------------------------------------------------------------------------------------------
// NCHW fp32 (N,C,H,W) -> split NHWC with Cp >= C channels (zero padded).  Optional per-pixel factor
// (1 - mask[n,0,h,w]) fuses the soft masking x*(1-cmap) of Demo_RSSS.py:290-291.
// Block = 32 consecutive pixels of one image x all channels, through shared memory: the NCHW reads are coalesced along the
// pixels (lane = pixel), the NHWC writes along the channels (8 lanes x 16 bytes = one pixel's 128-byte row) — a thread
// writing its whole pixel row by itself scatters 16-byte pieces over 32 different lines per store instruction.
constexpr int ST_PIX = 32;
constexpr int ST_CH = 256;      // channels per block (blockIdx.z walks wider tensors in chunks)
__global__ void stage_kernel(const float* __restrict__ src, const float* __restrict__ mask, int C, int Cp, long long HW,
                             long long npix, __nv_bfloat16* hi, __nv_bfloat16* lo, int ld) {
    extern __shared__ float st_tile[];                 // [ST_PIX][cw + 1]
    const int cbase = blockIdx.z * ST_CH;
    const int cw = (Cp - cbase) < ST_CH ? (Cp - cbase) : ST_CH;
    const int pitch = cw + 1;
    const long long n = blockIdx.y;
    const long long p0 = blockIdx.x * 1LL * ST_PIX;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const long long p = p0 + lane;
    const bool in = p < HW;
    const float f = (mask && in) ? 1.f - mask[n * HW + p] : 1.f;
    const float* s = src + n * C * HW + p;
    for (int c = warp; c < cw; c += nwarp)
        st_tile[lane * pitch + c] = (in && cbase + c < C) ? s[(cbase + c) * HW] * f : 0.f;
    __syncthreads();
    const int cg = cw / 8;
    for (int slot = threadIdx.x; slot < ST_PIX * cg; slot += blockDim.x) {
        const int i = slot / cg, c0 = (slot - i * cg) * 8;
        if (p0 + i >= HW) continue;
        F8 r;
#pragma unroll
        for (int k = 0; k < 8; ++k) r.v[k] = st_tile[i * pitch + c0 + k];
        st_split8(hi, lo, static_cast<size_t>(n * HW + p0 + i) * ld + cbase + c0, r);
    }
}

// NCHW fp32 (N,C,H,W) -> split NHWC (N,OH,OW,Kp) holding, per OUTPUT pixel of a 3x3 / stride 2 / pad 1 convolution, its whole
// receptive field: dst[n,oh,ow,(r*3+s)*C + c] = src[n,c,2*oh-1+r,2*ow-1+s] (0 outside the image, 0 for k >= 9*C).  The first
// discriminator layer (Module.py:196, 13 -> 64 channels) then is ONE K = 128 GEMM over a quarter of the pixels instead of 9 taps
// of a 64-channel zero-padded tensor (engine.py: conv_im2col_s2).  One thread = one (pixel, 8 consecutive k).
template <int PIX>
__global__ void stage_im2col_s2_kernel(const float* __restrict__ src, int C, int H, int W, int OH, int OW, int Kp,
                                       __nv_bfloat16* hi, __nv_bfloat16* lo) {
    // block = PIX consecutive output pixels of one output row (128 where the row is long enough: the kernel is bound by the
    // load -> barrier -> store round trip of a block, so more pixels per block = more bytes in flight): the 3 input rows x (2*32 + 1) pixels x C channels they read are
    // staged in shared memory with coalesced loads, then every (pixel, 8 consecutive k) slot is written as one 16-byte piece
    extern __shared__ float im_tile[];                 // [3][C][IM_W]
    constexpr int IM_W = 2 * PIX + 1;
    const int n = blockIdx.z, oh = blockIdx.y, ow0 = blockIdx.x * PIX;
    const long long HW = static_cast<long long>(H) * W;
    const float* sn = src + static_cast<long long>(n) * C * HW;
    const int iw0 = 2 * ow0 - 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    for (int rc = warp; rc < 3 * C; rc += nwarp) {          // one (input row, channel) line per warp pass, lanes along the row
        const int r = rc / C, c = rc - r * C;
        const int ih = 2 * oh - 1 + r;
        const bool row_in = ih >= 0 && ih < H;
        const float* line = sn + c * HW + static_cast<long long>(row_in ? ih : 0) * W;
        for (int x = lane; x < IM_W; x += 32) {
            const int iw = iw0 + x;
            im_tile[rc * IM_W + x] = (row_in && iw >= 0 && iw < W) ? __ldg(line + iw) : 0.f;
        }
    }
    // k -> offset of (tap row, channel, tap column) inside the staged tile (or -1 for the zero padding k >= 9*C)
    auto k_off = [&](int k) {
        const int tap = k / C, c = k - tap * C;
        const int tr = tap / 3, ts = tap - tr * 3;
        return tap < 9 ? (tr * C + c) * IM_W + ts : -1;
    };
    const int kg = Kp / 8;
    if (blockDim.x % kg == 0) {
        // every slot a thread visits has the same k group (the slot stride is a multiple of kg): its 8 offsets live in registers
        // (a shared-memory table read with stride 8 is a 4-way bank conflict per element — measured 3x the HBM time)
        const int k0 = (threadIdx.x % kg) * 8;
        int off[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) off[j] = k_off(k0 + j);
        __syncthreads();
        for (int i = threadIdx.x / kg; i < PIX; i += blockDim.x / kg) {
            if (ow0 + i >= OW) break;
            F8 r;
#pragma unroll
            for (int j = 0; j < 8; ++j) r.v[j] = off[j] >= 0 ? im_tile[off[j] + 2 * i] : 0.f;
            const size_t pix = (static_cast<size_t>(n) * OH + oh) * OW + ow0 + i;
            st_split8(hi, lo, pix * Kp + k0, r);
        }
        return;
    }
    int* lut = reinterpret_cast<int*>(im_tile + 3 * C * IM_W);
    for (int k = threadIdx.x; k < Kp; k += blockDim.x) lut[k] = k_off(k);
    __syncthreads();
    for (int slot = threadIdx.x; slot < PIX * kg; slot += blockDim.x) {
        const int i = slot / kg, k0 = (slot - i * kg) * 8;
        if (ow0 + i >= OW) continue;
        F8 r;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int off = lut[k0 + j];
            r.v[j] = off >= 0 ? im_tile[off + 2 * i] : 0.f;
        }
        const size_t pix = (static_cast<size_t>(n) * OH + oh) * OW + ow0 + i;
        st_split8(hi, lo, pix * Kp + k0, r);
    }
}

// NCHW fp32 (N,C<=16,H,W) -> split NHWC (N,H,W+M,64) with FOUR horizontally adjacent pixels packed into the channel
// axis: dst[n,h,w'',j*16+c] = src[n,c,h,w''-M+j] (0 outside).  A 9x9 convolution over 13 bands then needs
// ceil(9/4) = 3 taps of 64 channels per filter row instead of 9 (engine.py: conv_small_in / conv_small_out).
__global__ void stage_pack4_kernel(const float* __restrict__ src, int C, int H, int W, int M, __nv_bfloat16* hi,
                                   __nv_bfloat16* lo) {
    // block = 32 consecutive packed pixels of one row; the 35 source pixels x C channels they read go through shared memory
    __shared__ float pk_tile[16][ST_PIX + 4];
    const int n = blockIdx.z, h = blockIdx.y, wq0 = blockIdx.x * ST_PIX;
    const int Wp = W + M;
    const long long HW = static_cast<long long>(H) * W;
    const float* s = src + static_cast<long long>(n) * C * HW + static_cast<long long>(h) * W;
    for (int e = threadIdx.x; e < 16 * (ST_PIX + 3); e += blockDim.x) {
        const int x = e % (ST_PIX + 3), c = e / (ST_PIX + 3);
        const int w = wq0 - M + x;
        pk_tile[c][x] = (c < C && w >= 0 && w < W) ? s[c * HW + w] : 0.f;
    }
    __syncthreads();
    const int i = threadIdx.x >> 3, g = threadIdx.x & 7;      // 256 threads = 32 pixels x 8 channel groups
    if (wq0 + i >= Wp) return;
    const int j = g >> 1, c0 = (g & 1) * 8;
    F8 r;
#pragma unroll
    for (int k = 0; k < 8; ++k) r.v[k] = pk_tile[c0 + k][i + j];
    const size_t pix = (static_cast<size_t>(n) * H + h) * Wp + wq0 + i;
    st_split8(hi, lo, pix * 64 + g * 8, r);
}

// NCHW fp32 (N,C,H,W) -> split NHWC (N,H,W+M,Kp) with P horizontally adjacent pixels packed TIGHTLY into the channel axis:
// dst[n,h,w'',j*C+c] = src[n,c,h,w''-M+j] for j < P (0 outside the image, 0 for k >= P*C).  With P = the filter width a whole
// filter row of a KxK convolution over few bands is ONE tap: 9 x 13 = 117 -> K = 128 for the Generator's 9x9 layers
// (Module.py:146,158) instead of 3 taps of 64 in the 4-pixel form (91 % instead of 61 % of the MMA work is useful).
template <int PIX>
__global__ void stage_rowpack_kernel(const float* __restrict__ src, int C, int H, int W, int M, int P, int Kp,
                                     __nv_bfloat16* hi, __nv_bfloat16* lo) {
    // block = PIX consecutive packed pixels of one row (see stage_im2col_s2_kernel); the (PIX + P - 1) source pixels x C channels they read go through
    // shared memory (coalesced loads), then every (pixel, 8 consecutive k) slot is written as one 16-byte piece per plane
    extern __shared__ float rp_tile[];                   // [C][TW] + k -> offset table
    const int TW = (PIX + P - 1) | 1;                 // odd pitch: the channel lines start in different banks
    const int n = blockIdx.z, h = blockIdx.y, wq0 = blockIdx.x * PIX;
    const int Wp = W + M;
    const long long HW = static_cast<long long>(H) * W;
    const float* s = src + static_cast<long long>(n) * C * HW + static_cast<long long>(h) * W;
    for (int e = threadIdx.x; e < C * TW; e += blockDim.x) {
        const int c = e / TW, x = e - c * TW;
        const int w = wq0 - M + x;
        rp_tile[e] = (w >= 0 && w < W) ? __ldg(s + c * HW + w) : 0.f;
    }
    auto k_off = [&](int k) {
        const int j = k / C, c = k - j * C;
        return j < P ? c * TW + j : -1;
    };
    const int kg = Kp / 8;
    if (blockDim.x % kg == 0) {       // a thread keeps one k group: its 8 offsets live in registers (see stage_im2col_s2_kernel)
        const int k0 = (threadIdx.x % kg) * 8;
        int off[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) off[j] = k_off(k0 + j);
        __syncthreads();
        for (int i = threadIdx.x / kg; i < PIX; i += blockDim.x / kg) {
            if (wq0 + i >= Wp) break;
            F8 r;
#pragma unroll
            for (int j = 0; j < 8; ++j) r.v[j] = off[j] >= 0 ? rp_tile[off[j] + i] : 0.f;
            const size_t pix = (static_cast<size_t>(n) * H + h) * Wp + wq0 + i;
            st_split8(hi, lo, pix * Kp + k0, r);
        }
        return;
    }
    int* lut = reinterpret_cast<int*>(rp_tile + C * TW);
    for (int k = threadIdx.x; k < Kp; k += blockDim.x) lut[k] = k_off(k);
    __syncthreads();
    for (int slot = threadIdx.x; slot < PIX * kg; slot += blockDim.x) {
        const int i = slot / kg, k0 = (slot - i * kg) * 8;
        if (wq0 + i >= Wp) continue;
        F8 r;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int off = lut[k0 + j];
            r.v[j] = off >= 0 ? rp_tile[off + i] : 0.f;
        }
        const size_t pix = (static_cast<size_t>(n) * H + h) * Wp + wq0 + i;
        st_split8(hi, lo, pix * Kp + k0, r);
    }
}

// fp32 NHWC (pitch ld) -> NCHW fp32 (N,C,H,W); `accumulate` adds into dst.
__global__ void unstage_kernel(const float* __restrict__ src, int ld, int C, long long HW, long long npix,
                               float* __restrict__ dst, int accumulate) {
    const long long pix = blockIdx.x * 1LL * blockDim.x + threadIdx.x;
    if (pix >= npix) return;
    const long long n = pix / HW, p = pix - n * HW;
    const float* s = src + static_cast<size_t>(pix) * ld;
    float* d = dst + n * C * HW + p;
    for (int c = 0; c < C; ++c) {
        const float v = s[c];
        d[c * HW] = accumulate ? d[c * HW] + v : v;
    }
}

// split NHWC (pitch ld) -> NCHW fp32 (block boundary of the stand-alone DoubleConv/Down/Up/ResidualBlock modules)
__global__ void unstage_split_kernel(const __nv_bfloat16* hi, const __nv_bfloat16* lo, int ld, int C, long long HW,
                                     long long npix, float* __restrict__ dst) {
    const long long pix = blockIdx.x * 1LL * blockDim.x + threadIdx.x;
    if (pix >= npix) return;
    const long long n = pix / HW, p = pix - n * HW;
    float* d = dst + n * C * HW + p;
    for (int c0 = 0; c0 < C; c0 += 8) {
        const F8 v = ld_split8(hi, lo, static_cast<size_t>(pix) * ld + c0);
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (c0 + j < C) d[(c0 + j) * HW] = v.v[j];
    }
}

// NCHW fp32 -> fp32 NHWC (pitch ld, Cp channels, zero padded): an incoming boundary gradient
__global__ void stage_f32_kernel(const float* __restrict__ src, int C, int Cp, long long HW, long long npix,
                                 float* __restrict__ dst, int ld) {
    const long long pix = blockIdx.x * 1LL * blockDim.x + threadIdx.x;
    if (pix >= npix) return;
    const long long n = pix / HW, p = pix - n * HW;
    const float* s = src + n * C * HW + p;
    for (int c0 = 0; c0 < Cp; c0 += 8) {
        F8 r;
#pragma unroll
        for (int j = 0; j < 8; ++j) r.v[j] = (c0 + j < C) ? s[(c0 + j) * HW] : 0.f;
        st_f32x8(dst + static_cast<size_t>(pix) * ld + c0, r);
    }
}

// ---------------------------------------------------------------------------------------------
// BatchNorm statistics -> per-channel affine
// ---------------------------------------------------------------------------------------------
__global__ void bn_stats_kernel(const float* __restrict__ z, int ld, long long npix, int Cp, double* sum, double* sq) {
    // blockDim = 256; channel group (8 ch) = tid % cg; pixel lane = tid / cg
    const int cg = Cp / 8;
    const int g = threadIdx.x % cg, lane = threadIdx.x / cg, lanes = blockDim.x / cg;
    float a[8] = {0, 0, 0, 0, 0, 0, 0, 0}, b[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (long long p = blockIdx.x * 1LL * lanes + lane; p < npix; p += 1LL * gridDim.x * lanes) {
        const F8 v = ld_f32x8(z + static_cast<size_t>(p) * ld + g * 8);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            a[j] += v.v[j];
            b[j] = fmaf(v.v[j], v.v[j], b[j]);
        }
    }
    __shared__ float red[NT * 16];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        red[threadIdx.x * 16 + j] = a[j];
        red[threadIdx.x * 16 + 8 + j] = b[j];
    }
    __syncthreads();
    // thread t < cg*16 reduces column (group t/16, slot t%16) over the pixel lanes
    for (int t = threadIdx.x; t < cg * 16; t += blockDim.x) {
        const int gg = t / 16, slot = t % 16;
        double acc = 0.0;
        for (int l = 0; l < lanes; ++l) acc += red[(l * cg + gg) * 16 + slot];
        if (slot < 8)
            atomicAdd(sum + gg * 8 + slot, acc);
        else
            atomicAdd(sq + gg * 8 + slot - 8, acc);
    }
}

// mode: 1 = training (batch statistics, running stats updated), 0 = eval (running statistics)
__global__ void bn_finalize_kernel(const double* sum, const double* sq, double count, const float* gamma,
                                   const float* beta, float* running_mean, float* running_var, int C, int Cp,
                                   float momentum, float eps, int mode, float* scale, float* shift, float* mean_out,
                                   float* invstd_out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= Cp) return;
    if (c >= C) {
        scale[c] = 0.f; shift[c] = 0.f; mean_out[c] = 0.f; invstd_out[c] = 0.f;
        return;
    }
    double mean, var;
    if (mode) {
        mean = sum[c] / count;
        var = sq[c] / count - mean * mean;
        if (var < 0.0) var = 0.0;
        const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
        running_mean[c] = static_cast<float>((1.0 - momentum) * running_mean[c] + momentum * mean);
        running_var[c] = static_cast<float>((1.0 - momentum) * running_var[c] + momentum * unbiased);
    } else {
        mean = running_mean[c];
        var = running_var[c];
    }
    const double invstd = 1.0 / sqrt(var + static_cast<double>(eps));
    const double sc = gamma[c] * invstd;
    scale[c] = static_cast<float>(sc);
    shift[c] = static_cast<float>(beta[c] - mean * sc);
    mean_out[c] = static_cast<float>(mean);
    invstd_out[c] = static_cast<float>(invstd);
}

// out = act(z*scale + shift) (+ residual), written as a split tensor and/or fp32
__global__ void bn_act_fwd_kernel(const float* __restrict__ z, int z_ld, const float* __restrict__ scale,
                                  const float* __restrict__ shift, int act, const float* slope_ptr, float slope_const,
                                  const __nv_bfloat16* res_hi, const __nv_bfloat16* res_lo, int res_ld,
                                  __nv_bfloat16* out_hi, __nv_bfloat16* out_lo, int out_ld, long long npix, int Cp) {
    const int cg = Cp / 8;
    const long long idx = blockIdx.x * 1LL * blockDim.x + threadIdx.x;
    if (idx >= npix * cg) return;
    const long long pix = idx / cg;
    const int c0 = static_cast<int>(idx - pix * cg) * 8;
    const float slope = slope_ptr ? *slope_ptr : slope_const;
    F8 v = ld_f32x8(z + static_cast<size_t>(pix) * z_ld + c0);
    if (scale) {
        const F8 sc = ld_f32x8(scale + c0), sh = ld_f32x8(shift + c0);
#pragma unroll
        for (int j = 0; j < 8; ++j) v.v[j] = fmaf(v.v[j], sc.v[j], sh.v[j]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) v.v[j] = act_fwd(act, v.v[j], slope);
    if (res_hi) {
        const F8 r = ld_split8(res_hi, res_lo, static_cast<size_t>(pix) * res_ld + c0);
#pragma unroll
        for (int j = 0; j < 8; ++j) v.v[j] += r.v[j];
    }
    st_split8(out_hi, out_lo, static_cast<size_t>(pix) * out_ld + c0, v);
}

// s1[c] += sum dy, s2[c] += sum dy * xhat, dslope += sum da * u * [u <= 0]   (dy = da * act'(u))
__global__ void bn_act_bwd_reduce_kernel(const float* __restrict__ da, int da_ld, const float* __restrict__ z, int z_ld,
                                         const float* __restrict__ scale, const float* __restrict__ shift,
                                         const float* __restrict__ mean, const float* __restrict__ invstd, int act,
                                         const float* slope_ptr, float slope_const, long long npix, int Cp, double* s1,
                                         double* s2, double* dslope) {
    const int cg = Cp / 8;
    const int g = threadIdx.x % cg, lane = threadIdx.x / cg, lanes = blockDim.x / cg;
    const float slope = slope_ptr ? *slope_ptr : slope_const;
    float a[8] = {0, 0, 0, 0, 0, 0, 0, 0}, b[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    float ds = 0.f;
    F8 sc, sh, mu, is;
#pragma unroll
    for (int j = 0; j < 8; ++j) { sc.v[j] = 1.f; sh.v[j] = 0.f; mu.v[j] = 0.f; is.v[j] = 1.f; }
    if (scale) {
        sc = ld_f32x8(scale + g * 8); sh = ld_f32x8(shift + g * 8);
        mu = ld_f32x8(mean + g * 8); is = ld_f32x8(invstd + g * 8);
    }
    for (long long p = blockIdx.x * 1LL * lanes + lane; p < npix; p += 1LL * gridDim.x * lanes) {
        const F8 zz = ld_f32x8(z + static_cast<size_t>(p) * z_ld + g * 8);
        const F8 dd = ld_f32x8(da + static_cast<size_t>(p) * da_ld + g * 8);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float u = fmaf(zz.v[j], sc.v[j], sh.v[j]);
            const float dy = dd.v[j] * act_grad(act, u, slope);
            a[j] += dy;
            b[j] = fmaf(dy, (zz.v[j] - mu.v[j]) * is.v[j], b[j]);
            if (act == FCD_ACT_PRELU && !(u > 0.f)) ds = fmaf(dd.v[j], u, ds);
        }
    }
    __shared__ float red[NT * 16];
    __shared__ float red_ds[NT / 32];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        red[threadIdx.x * 16 + j] = a[j];
        red[threadIdx.x * 16 + 8 + j] = b[j];
    }
    if (dslope) {
        ds = warp_sum(ds);
        if ((threadIdx.x & 31) == 0) red_ds[threadIdx.x >> 5] = ds;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < cg * 16; t += blockDim.x) {
        const int gg = t / 16, slot = t % 16;
        double acc = 0.0;
        for (int l = 0; l < lanes; ++l) acc += red[(l * cg + gg) * 16 + slot];
        if (slot < 8)
            atomicAdd(s1 + gg * 8 + slot, acc);
        else
            atomicAdd(s2 + gg * 8 + slot - 8, acc);
    }
    if (dslope && threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < NT / 32; ++w) t += red_ds[w];
        atomicAdd(dslope, t);
    }
}

// per-channel backward coefficients + parameter gradients:
//   training: c1 = s1/count, c2 = s2/count;  eval: c1 = c2 = 0.   dgamma (+)= s2, dbeta (+)= s1, dslope (+)= ds.
// dbias (optional) = gradient of the bias of the convolution that produced z = sum over pixels of dz, in closed form from
// the same sums (dz = scale * (dy - c1 - xhat * c2)  =>  sum dz = scale * (s1 - count*c1 - c2 * sum xhat)); this replaces a
// full read of dz by the weight-gradient call.  In training mode the result is the analytic zero up to rounding, exactly
// like the reference's autograd value.
__global__ void bn_bwd_finalize_kernel(const double* s1, const double* s2, double count, int train, int C, int Cp,
                                       float* c1, float* c2, float* dgamma, float* dbeta, int accumulate,
                                       const double* ds, float* dslope, const float* scale, const double* zsum,
                                       const float* mean, const float* invstd, float* dbias, int dbias_accumulate,
                                       const double* gs1, const double* gs2, double gcount) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c == 0 && dslope) *dslope = (accumulate ? *dslope : 0.f) + static_cast<float>(*ds);
    if (c >= Cp) return;
    // the means that enter dz are those of the WHOLE batch the statistics were taken over: the local sums, or (SyncBN) the sums
    // over all ranks; parameter gradients always come from the LOCAL sums (the gradient exchange averages them afterwards)
    const double m1 = gs1 ? gs1[c] / gcount : s1[c] / count;
    const double m2 = gs2 ? gs2[c] / gcount : s2[c] / count;
    const float k1 = (train && c < C) ? static_cast<float>(m1) : 0.f;
    const float k2 = (train && c < C) ? static_cast<float>(m2) : 0.f;
    c1[c] = k1;
    c2[c] = k2;
    if (c < C) {
        if (dgamma) dgamma[c] = (accumulate ? dgamma[c] : 0.f) + static_cast<float>(s2[c]);
        if (dbeta) dbeta[c] = (accumulate ? dbeta[c] : 0.f) + static_cast<float>(s1[c]);
        if (dbias) {
            double v = s1[c];
            if (scale) {
                const double sum_xhat = (zsum && mean && invstd)
                                            ? (zsum[c] - count * static_cast<double>(mean[c])) * static_cast<double>(invstd[c])
                                            : 0.0;
                v = static_cast<double>(scale[c]) * (s1[c] - count * static_cast<double>(k1) - static_cast<double>(k2) * sum_xhat);
            }
            dbias[c] = (dbias_accumulate ? dbias[c] : 0.f) + static_cast<float>(v);
        }
    }
}

// A thread owns ONE 8-channel group and walks APPLY_PIX pixels with it, so the six per-channel vectors are loaded once per
// thread instead of once per pixel (they were 12 of the 16 load instructions of the one-pixel form).
constexpr int APPLY_PIX = 4;
__global__ void bn_act_bwd_apply_kernel(const float* __restrict__ da, int da_ld, const float* __restrict__ z, int z_ld,
                                        const float* __restrict__ scale, const float* __restrict__ shift,
                                        const float* __restrict__ mean, const float* __restrict__ invstd,
                                        const float* __restrict__ c1, const float* __restrict__ c2, int act,
                                        const float* slope_ptr, float slope_const, __nv_bfloat16* dz_hi,
                                        __nv_bfloat16* dz_lo, int dz_ld, long long npix, int Cp, long long pstride) {
    const int cg = Cp / 8;
    const long long slot = blockIdx.x * 1LL * blockDim.x + threadIdx.x;     // (pixel slot, channel group)
    const long long pslot = slot / cg;
    const int c0 = static_cast<int>(slot - pslot * cg) * 8;
    if (pslot >= pstride) return;                                           // pstride = number of pixel slots
    const float slope = slope_ptr ? *slope_ptr : slope_const;
    F8 sc, sh, mu, is, k1, k2;
    if (scale) {
        sc = ld_f32x8(scale + c0); sh = ld_f32x8(shift + c0); mu = ld_f32x8(mean + c0);
        is = ld_f32x8(invstd + c0); k1 = ld_f32x8(c1 + c0); k2 = ld_f32x8(c2 + c0);
    }
#pragma unroll
    for (int it = 0; it < APPLY_PIX; ++it) {
        const long long pix = pslot + it * pstride;
        if (pix >= npix) break;
        const F8 dd = ld_f32x8(da + static_cast<size_t>(pix) * da_ld + c0);
        F8 out;
        if (scale) {
            const F8 zz = ld_f32x8(z + static_cast<size_t>(pix) * z_ld + c0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float u = fmaf(zz.v[j], sc.v[j], sh.v[j]);
                const float dy = dd.v[j] * act_grad(act, u, slope);
                out.v[j] = sc.v[j] * (dy - k1.v[j] - (zz.v[j] - mu.v[j]) * is.v[j] * k2.v[j]);
            }
        } else if (act != FCD_ACT_NONE) {
            const F8 zz = ld_f32x8(z + static_cast<size_t>(pix) * z_ld + c0);
#pragma unroll
            for (int j = 0; j < 8; ++j) out.v[j] = dd.v[j] * act_grad(act, zz.v[j], slope);
        } else {
            out = dd;
        }
        st_split8(dz_hi, dz_lo, static_cast<size_t>(pix) * dz_ld + c0, out);
    }
}

// ---------------------------------------------------------------------------------------------
// MaxPool2d(2)  (floor: an odd last row / column is dropped)
// ---------------------------------------------------------------------------------------------
template <typename I>
__global__ void maxpool_fwd_kernel(const __nv_bfloat16* in_hi, const __nv_bfloat16* in_lo, int in_ld, int N, int H, int W,
                                   int Cp, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo, int out_ld) {
    const int OH = H / 2, OW = W / 2, cg = Cp / 8;
    const I idx = static_cast<I>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (static_cast<long long>(idx) >= 1LL * N * OH * OW * cg) return;
    I t = idx;
    const int c0 = static_cast<int>(t % cg) * 8; t /= cg;
    const int ow = static_cast<int>(t % OW); t /= OW;
    const int oh = static_cast<int>(t % OH);
    const int n = static_cast<int>(t / OH);
    const size_t base = ((static_cast<size_t>(n) * H + 2 * oh) * W + 2 * ow) * in_ld + c0;
    F8 m = ld_split8(in_hi, in_lo, base);
    const size_t offs[3] = {static_cast<size_t>(in_ld), static_cast<size_t>(W) * in_ld, static_cast<size_t>(W + 1) * in_ld};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const F8 v = ld_split8(in_hi, in_lo, base + offs[k]);
#pragma unroll
        for (int j = 0; j < 8; ++j) m.v[j] = v.v[j] > m.v[j] ? v.v[j] : m.v[j];
    }
    st_split8(out_hi, out_lo, ((static_cast<size_t>(n) * OH + oh) * OW + ow) * out_ld + c0, m);
}

// d_in[h,w] (+)= d_out[h/2,w/2] iff (h,w) is the first maximum of its window (torch keeps the first index).
// One thread = one 2x2 WINDOW x 8 channels: the four inputs and the output gradient are read once, the four input gradients
// written once (the per-input-pixel form read every window four times: 2.8 TB/s of algorithmic bytes).  Pixels of an odd last
// row / column belong to no window: gradient 0, written by the threads of the last window column / row.
template <typename I>
__global__ void maxpool_bwd_kernel(const float* __restrict__ d_out, int dout_ld, const __nv_bfloat16* in_hi,
                                   const __nv_bfloat16* in_lo, int in_ld, int N, int H, int W, int Cp, float* d_in,
                                   int din_ld, int accumulate) {
    const int OH = H / 2, OW = W / 2, cg = Cp / 8;
    const I idx = static_cast<I>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (static_cast<long long>(idx) >= 1LL * N * OH * OW * cg) return;
    I t = idx;
    const int c0 = static_cast<int>(t % cg) * 8; t /= cg;
    const int ow = static_cast<int>(t % OW); t /= OW;
    const int oh = static_cast<int>(t % OH);
    const int n = static_cast<int>(t / OH);
    const size_t pix0 = (static_cast<size_t>(n) * H + 2 * oh) * W + 2 * ow;
    const size_t poff[4] = {0, 1, static_cast<size_t>(W), static_cast<size_t>(W) + 1};
    F8 m = ld_split8(in_hi, in_lo, pix0 * in_ld + c0);
    int arg[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int k = 1; k < 4; ++k) {
        const F8 v = ld_split8(in_hi, in_lo, (pix0 + poff[k]) * in_ld + c0);
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (v.v[j] > m.v[j]) {
                m.v[j] = v.v[j];
                arg[j] = k;
            }
    }
    const F8 go = ld_f32x8(d_out + ((static_cast<size_t>(n) * OH + oh) * OW + ow) * dout_ld + c0);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float* dst = d_in + (pix0 + poff[k]) * din_ld + c0;
        F8 g;
#pragma unroll
        for (int j = 0; j < 8; ++j) g.v[j] = (arg[j] == k) ? go.v[j] : 0.f;
        if (accumulate) {
            const F8 old = ld_f32x8(dst);
#pragma unroll
            for (int j = 0; j < 8; ++j) g.v[j] += old.v[j];
        }
        st_f32x8(dst, g);
    }
    if (!accumulate) {       // the dropped odd row / column: zero gradient (accumulate: nothing to add)
        F8 z;
#pragma unroll
        for (int j = 0; j < 8; ++j) z.v[j] = 0.f;
        const bool last_w = (W & 1) && ow == OW - 1, last_h = (H & 1) && oh == OH - 1;
        const size_t row0 = (static_cast<size_t>(n) * H + 2 * oh) * W;
        if (last_w) {
            st_f32x8(d_in + (row0 + W - 1) * din_ld + c0, z);
            st_f32x8(d_in + (row0 + W + W - 1) * din_ld + c0, z);
        }
        if (last_h) {
            const size_t rowl = (static_cast<size_t>(n) * H + H - 1) * W;
            st_f32x8(d_in + (rowl + 2 * ow) * din_ld + c0, z);
            st_f32x8(d_in + (rowl + 2 * ow + 1) * din_ld + c0, z);
            if (last_w) st_f32x8(d_in + (rowl + W - 1) * din_ld + c0, z);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// bilinear x2 up-sampling (align_corners=True) written into a (possibly larger, zero padded) destination
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void src_coord(int o, float ratio, int in, int& i0, int& i1, float& l1) {
    const float f = ratio * o;  // torch: area_pixel_compute_source_index(align_corners=True)
    i0 = static_cast<int>(f);
    if (i0 > in - 1) i0 = in - 1;
    i1 = i0 + (i0 < in - 1 ? 1 : 0);
    l1 = f - i0;
}

template <typename I>
__global__ void upsample_fwd_kernel(const __nv_bfloat16* in_hi, const __nv_bfloat16* in_lo, int in_ld, int N, int h, int w,
                                    int Cp, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo, int out_ld, int H, int W,
                                    int pad_top, int pad_left) {
    const int cg = Cp / 8, uh = 2 * h, uw = 2 * w;
    const I idx = static_cast<I>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (static_cast<long long>(idx) >= 1LL * N * H * W * cg) return;
    I t = idx;
    const int c0 = static_cast<int>(t % cg) * 8; t /= cg;
    const int X = static_cast<int>(t % W); t /= W;
    const int Y = static_cast<int>(t % H);
    const int n = static_cast<int>(t / H);
    F8 o;
#pragma unroll
    for (int j = 0; j < 8; ++j) o.v[j] = 0.f;
    const int oy = Y - pad_top, ox = X - pad_left;
    if (oy >= 0 && oy < uh && ox >= 0 && ox < uw) {
        const float rh = uh > 1 ? static_cast<float>(h - 1) / (uh - 1) : 0.f;
        const float rw = uw > 1 ? static_cast<float>(w - 1) / (uw - 1) : 0.f;
        int y0, y1, x0, x1;
        float ly, lx;
        src_coord(oy, rh, h, y0, y1, ly);
        src_coord(ox, rw, w, x0, x1, lx);
        const size_t b = static_cast<size_t>(n) * h * w;
        const F8 v00 = ld_split8(in_hi, in_lo, (b + static_cast<size_t>(y0) * w + x0) * in_ld + c0);
        const F8 v01 = ld_split8(in_hi, in_lo, (b + static_cast<size_t>(y0) * w + x1) * in_ld + c0);
        const F8 v10 = ld_split8(in_hi, in_lo, (b + static_cast<size_t>(y1) * w + x0) * in_ld + c0);
        const F8 v11 = ld_split8(in_hi, in_lo, (b + static_cast<size_t>(y1) * w + x1) * in_ld + c0);
        const float hy = 1.f - ly, hx = 1.f - lx;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            o.v[j] = hy * (hx * v00.v[j] + lx * v01.v[j]) + ly * (hx * v10.v[j] + lx * v11.v[j]);
    }
    st_split8(out_hi, out_lo, ((static_cast<size_t>(n) * H + Y) * W + X) * out_ld + c0, o);
}

// gather form of the backward: every source pixel sums the (<= 6x6) destination pixels that read it
template <typename I>
__global__ void upsample_bwd_kernel(const float* __restrict__ d_out, int dout_ld, int N, int h, int w, int Cp, int H,
                                    int W, int pad_top, int pad_left, float* d_in, int din_ld) {
    const int cg = Cp / 8, uh = 2 * h, uw = 2 * w;
    const I idx = static_cast<I>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (static_cast<long long>(idx) >= 1LL * N * h * w * cg) return;
    I t = idx;
    const int c0 = static_cast<int>(t % cg) * 8; t /= cg;
    const int x = static_cast<int>(t % w); t /= w;
    const int y = static_cast<int>(t % h);
    const int n = static_cast<int>(t / h);
    const float rh = uh > 1 ? static_cast<float>(h - 1) / (uh - 1) : 0.f;
    const float rw = uw > 1 ? static_cast<float>(w - 1) / (uw - 1) : 0.f;
    // candidate destination range: src(o) in (y-1, y+1)
    int oy_lo = rh > 0.f ? static_cast<int>(floorf((y - 1) / rh)) - 1 : 0;
    int oy_hi = rh > 0.f ? static_cast<int>(ceilf((y + 1) / rh)) + 1 : uh - 1;
    int ox_lo = rw > 0.f ? static_cast<int>(floorf((x - 1) / rw)) - 1 : 0;
    int ox_hi = rw > 0.f ? static_cast<int>(ceilf((x + 1) / rw)) + 1 : uw - 1;
    oy_lo = max(oy_lo, 0); ox_lo = max(ox_lo, 0);
    oy_hi = min(oy_hi, uh - 1); ox_hi = min(ox_hi, uw - 1);
    F8 g;
#pragma unroll
    for (int j = 0; j < 8; ++j) g.v[j] = 0.f;
    for (int oy = oy_lo; oy <= oy_hi; ++oy) {
        int y0, y1;
        float ly;
        src_coord(oy, rh, h, y0, y1, ly);
        float wy = 0.f;
        if (y0 == y) wy += 1.f - ly;
        if (y1 == y) wy += (y1 != y0) ? ly : ly;  // when y1 == y0 (last row) both weights land on the same pixel
        if (y0 == y && y1 == y) wy = 1.f;
        if (wy == 0.f) continue;
        const int Y = oy + pad_top;
        if (Y < 0 || Y >= H) continue;
        for (int ox = ox_lo; ox <= ox_hi; ++ox) {
            int x0, x1;
            float lx;
            src_coord(ox, rw, w, x0, x1, lx);
            float wx = 0.f;
            if (x0 == x) wx += 1.f - lx;
            if (x1 == x) wx += lx;
            if (x0 == x && x1 == x) wx = 1.f;
            if (wx == 0.f) continue;
            const int X = ox + pad_left;
            if (X < 0 || X >= W) continue;
            const F8 v = ld_f32x8(d_out + ((static_cast<size_t>(n) * H + Y) * W + X) * dout_ld + c0);
            const float ww = wy * wx;
#pragma unroll
            for (int j = 0; j < 8; ++j) g.v[j] = fmaf(ww, v.v[j], g.v[j]);
        }
    }
    st_f32x8(d_in + ((static_cast<size_t>(n) * h + y) * w + x) * din_ld + c0, g);
}

// ---------------------------------------------------------------------------------------------
// ConvTranspose2d(k2,s2) = four 1x1 convolutions + pixel shuffle.  These two kernels are the shuffles.
// ---------------------------------------------------------------------------------------------
// src fp32 [4][N,h,w,Cp] (one plane per (i,j) sub-pixel)  ->  split dst slice (N,H,W) at (2y+i+pad_top, 2x+j+pad_left);
// pixels of dst outside the 2h x 2w region are zero filled.
__global__ void shuffle_up_kernel(const float* __restrict__ src, long long plane_stride, int src_ld, int N, int h, int w,
                                  int Cp, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo, int out_ld, int H, int W,
                                  int pad_top, int pad_left) {
    const int cg = Cp / 8;
    const long long idx = blockIdx.x * 1LL * blockDim.x + threadIdx.x;
    if (idx >= 1LL * N * H * W * cg) return;
    long long t = idx;
    const int c0 = static_cast<int>(t % cg) * 8; t /= cg;
    const int X = static_cast<int>(t % W); t /= W;
    const int Y = static_cast<int>(t % H);
    const int n = static_cast<int>(t / H);
    F8 o;
#pragma unroll
    for (int j = 0; j < 8; ++j) o.v[j] = 0.f;
    const int oy = Y - pad_top, ox = X - pad_left;
    if (oy >= 0 && oy < 2 * h && ox >= 0 && ox < 2 * w) {
        const int sub = (oy & 1) * 2 + (ox & 1);
        o = ld_f32x8(src + sub * plane_stride + ((static_cast<size_t>(n) * h + (oy >> 1)) * w + (ox >> 1)) * src_ld + c0);
    }
    st_split8(out_hi, out_lo, ((static_cast<size_t>(n) * H + Y) * W + X) * out_ld + c0, o);
}

// d_out fp32 (N,H,W) slice -> split [4][N,h,w,Cp] planes (the dz operand of the four 1x1 dgrad / wgrad calls)
__global__ void shuffle_down_kernel(const float* __restrict__ d_out, int dout_ld, int N, int h, int w, int Cp, int H,
                                    int W, int pad_top, int pad_left, __nv_bfloat16* g_hi, __nv_bfloat16* g_lo,
                                    long long plane_stride, int g_ld) {
    const int cg = Cp / 8;
    const long long idx = blockIdx.x * 1LL * blockDim.x + threadIdx.x;
    if (idx >= 4LL * N * h * w * cg) return;
    long long t = idx;
    const int c0 = static_cast<int>(t % cg) * 8; t /= cg;
    const int x = static_cast<int>(t % w); t /= w;
    const int y = static_cast<int>(t % h); t /= h;
    const int n = static_cast<int>(t % N);
    const int sub = static_cast<int>(t / N);
    const int Y = 2 * y + (sub >> 1) + pad_top, X = 2 * x + (sub & 1) + pad_left;
    F8 o;
#pragma unroll
    for (int j = 0; j < 8; ++j) o.v[j] = 0.f;
    if (Y >= 0 && Y < H && X >= 0 && X < W) o = ld_f32x8(d_out + ((static_cast<size_t>(n) * H + Y) * W + X) * dout_ld + c0);
    const size_t off = sub * plane_stride + ((static_cast<size_t>(n) * h + y) * w + x) * g_ld + c0;
    st_split8(g_hi, g_lo, off, o);
}

// fp32 NHWC -> split NHWC copy (channel slice aware); used to turn a gradient into a conv operand
__global__ void f32_to_split_kernel(const float* __restrict__ src, int src_ld, __nv_bfloat16* hi, __nv_bfloat16* lo,
                                    int dst_ld, long long npix, int Cp) {
    const int cg = Cp / 8;
    const long long idx = blockIdx.x * 1LL * blockDim.x + threadIdx.x;
    if (idx >= npix * cg) return;
    const long long pix = idx / cg;
    const int c0 = static_cast<int>(idx - pix * cg) * 8;
    st_split8(hi, lo, static_cast<size_t>(pix) * dst_ld + c0, ld_f32x8(src + static_cast<size_t>(pix) * src_ld + c0));
}

// dst += src on fp32 NHWC (pitch aware)
__global__ void add_f32_kernel(float* __restrict__ dst, int dst_ld, const float* __restrict__ src, int src_ld,
                               long long npix, int Cp) {
    const int cg = Cp / 8;
    const long long idx = blockIdx.x * 1LL * blockDim.x + threadIdx.x;
    if (idx >= npix * cg) return;
    const long long pix = idx / cg;
    const int c0 = static_cast<int>(idx - pix * cg) * 8;
    F8 a = ld_f32x8(dst + static_cast<size_t>(pix) * dst_ld + c0);
    const F8 b = ld_f32x8(src + static_cast<size_t>(pix) * src_ld + c0);
#pragma unroll
    for (int j = 0; j < 8; ++j) a.v[j] += b.v[j];
    st_f32x8(dst + static_cast<size_t>(pix) * dst_ld + c0, a);
}


// ---------------------------------------------------------------------------------------------
// feature-space MSE of the perception loss (Loss.py:36,48,59): the VGG16 stack runs on ONE batch holding the masked
// target images in its first half and the masked generated images in its second half, so the loss is the mean squared
// difference of the two halves of a split NHWC feature tensor.
// ---------------------------------------------------------------------------------------------
__global__ void mse_halves_fwd_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo, int ld,
                                      long long half_elems, long long npix, int Cp, double* __restrict__ acc) {
    const int cg = Cp / 8;
    const long long total = npix * cg, stride = 1LL * gridDim.x * blockDim.x;
    float part = 0.f;
    for (long long idx = blockIdx.x * 1LL * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const long long pix = idx / cg;
        const size_t off = static_cast<size_t>(pix) * ld + static_cast<int>(idx - pix * cg) * 8;
        const F8 a = ld_split8(hi, lo, off), b = ld_split8(hi, lo, off + half_elems);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float d = a.v[j] - b.v[j];
            part = fmaf(d, d, part);
        }
    }
    __shared__ double red[NT / 32];
    double w = warp_sum(static_cast<double>(part));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = w;
    __syncthreads();
    if (threadIdx.x < 32) {
        w = threadIdx.x < NT / 32 ? red[threadIdx.x] : 0.0;
        w = warp_sum(w);
        if (threadIdx.x == 0) atomicAdd(acc, w);
    }
}

// grad[first half] (+)= g * (a - b), grad[second half] (+)= -g * (a - b), g = *gout * scale
__global__ void mse_halves_bwd_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo, int ld,
                                      long long half_elems, long long npix, int Cp, const float* __restrict__ gout, float scale,
                                      float* __restrict__ grad, int grad_ld, long long grad_half_elems, int accumulate) {
    const int cg = Cp / 8;
    const long long idx = blockIdx.x * 1LL * blockDim.x + threadIdx.x;
    if (idx >= npix * cg) return;
    const long long pix = idx / cg;
    const int c0 = static_cast<int>(idx - pix * cg) * 8;
    const size_t off = static_cast<size_t>(pix) * ld + c0;
    const F8 a = ld_split8(hi, lo, off), b = ld_split8(hi, lo, off + half_elems);
    const float g = __ldg(gout) * scale;
    float* ga = grad + static_cast<size_t>(pix) * grad_ld + c0;
    float* gb = ga + grad_half_elems;
    F8 ra, rb;
    if (accumulate) {
        ra = ld_f32x8(ga);
        rb = ld_f32x8(gb);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float d = g * (a.v[j] - b.v[j]);
        ra.v[j] = accumulate ? ra.v[j] + d : d;
        rb.v[j] = accumulate ? rb.v[j] - d : -d;
    }
    st_f32x8(ga, ra);
    st_f32x8(gb, rb);
}

inline unsigned blocks_for(long long n) { return static_cast<unsigned>((n + NT - 1) / NT); }

inline int reduce_grid(long long npix, int lanes) {
    long long want = (npix + lanes * 8 - 1) / (lanes * 8);
    const long long cap = 8LL * sm_count();
    if (want > cap) want = cap;
    if (want < 1) want = 1;
    return static_cast<int>(want);
}

}  // namespace
}  // namespace fcd

using namespace fcd;

#define BF(p) reinterpret_cast<__nv_bfloat16*>(p)
#define CBF(p) reinterpret_cast<const __nv_bfloat16*>(p)

extern "C" {

int fcd_stage_nchw_to_split(const float* src, const float* mask, int N, int C, int H, int W, void* dst_hi, void* dst_lo,
                            int dst_ld, int Cp, void* stream) {
    FCD_CHECK_ARG(src && dst_hi && Cp % 8 == 0 && Cp >= C && dst_ld % 8 == 0, "fcd_stage_nchw_to_split: bad arguments");
    const long long npix = 1LL * N * H * W, HW = 1LL * H * W;
    FCD_CHECK_ARG(N <= 65535, "fcd_stage_nchw_to_split: at most 65535 images per call");
    const size_t smem = sizeof(float) * ST_PIX * ((Cp < ST_CH ? Cp : ST_CH) + 1);
    stage_kernel<<<dim3(static_cast<unsigned>((HW + ST_PIX - 1) / ST_PIX), N, (Cp + ST_CH - 1) / ST_CH), NT, smem,
                   as_stream(stream)>>>(
        src, mask, C, Cp, HW, npix, BF(dst_hi), BF(dst_lo), dst_ld);
    FCD_LAUNCH_OK();
    return FCD_OK;
}

int fcd_stage_im2col3x3s2(const float* src, int N, int C, int H, int W, void* dst_hi, void* dst_lo, int Kp, void* stream) {
    FCD_CHECK_ARG(src && dst_hi && N > 0 && C > 0 && H > 0 && W > 0, "fcd_stage_im2col3x3s2: bad arguments");
    FCD_CHECK_ARG(Kp % 8 == 0 && Kp >= 9 * C, "fcd_stage_im2col3x3s2: Kp must be a multiple of 8 and >= 9*C");
    const int OH = (H + 2 - 3) / 2 + 1, OW = (W + 2 - 3) / 2 + 1;
    FCD_CHECK_ARG(N <= 65535 && OH <= 65535, "fcd_stage_im2col3x3s2: dims exceed the launch grid");
    auto smem_for = [&](int pix) { return sizeof(float) * 3 * C * (2 * pix + 1) + sizeof(int) * Kp; };   // staged tile + k -> offset table
    FCD_CHECK_ARG(smem_for(ST_PIX) <= 48 * 1024, "fcd_stage_im2col3x3s2: too many channels (%d)", C);
    if (OW >= 128 && smem_for(128) <= 48 * 1024)
        stage_im2col_s2_kernel<128><<<dim3((OW + 127) / 128, OH, N), NT, smem_for(128), as_stream(stream)>>>(
            src, C, H, W, OH, OW, Kp, BF(dst_hi), BF(dst_lo));
    else
        stage_im2col_s2_kernel<ST_PIX><<<dim3((OW + ST_PIX - 1) / ST_PIX, OH, N), NT, smem_for(ST_PIX), as_stream(stream)>>>(
            src, C, H, W, OH, OW, Kp, BF(dst_hi), BF(dst_lo));
    FCD_LAUNCH_OK();
    return FCD_OK;
}

int fcd_stage_nchw_to_split_pack4(const float* src, int N, int C, int H, int W, int M, void* dst_hi, void* dst_lo, void* stream) {
    FCD_CHECK_ARG(src && dst_hi && C >= 1 && C <= 16 && M >= 0, "fcd_stage_nchw_to_split_pack4: needs 1 <= C <= 16");
    FCD_CHECK_ARG(N <= 65535 && H <= 65535 && M <= 4, "fcd_stage_nchw_to_split_pack4: dims exceed the launch grid / margin > 4");
    stage_pack4_kernel<<<dim3((W + M + ST_PIX - 1) / ST_PIX, H, N), NT, 0, as_stream(stream)>>>(src, C, H, W, M, BF(dst_hi),
                                                                                              BF(dst_lo));
    FCD_LAUNCH_OK();
    return FCD_OK;
}

int fcd_stage_nchw_to_split_rowpack(const float* src, int N, int C, int H, int W, int M, int P, int Kp, void* dst_hi,
                                    void* dst_lo, void* stream) {
    FCD_CHECK_ARG(src && dst_hi && C >= 1 && P >= 1 && M >= 0, "fcd_stage_nchw_to_split_rowpack: bad arguments");
    FCD_CHECK_ARG(Kp % 8 == 0 && Kp >= P * C, "fcd_stage_nchw_to_split_rowpack: Kp must be a multiple of 8 and >= P*C");
    FCD_CHECK_ARG(N <= 65535 && H <= 65535, "fcd_stage_nchw_to_split_rowpack: dims exceed the launch grid");
    auto smem_for = [&](int pix) { return sizeof(float) * C * ((pix + P - 1) | 1) + sizeof(int) * Kp; };
    FCD_CHECK_ARG(smem_for(ST_PIX) <= 48 * 1024, "fcd_stage_nchw_to_split_rowpack: too many channels (%d) / pixels per pack (%d)", C, P);
    if (W + M >= 128 && smem_for(128) <= 48 * 1024)
        stage_rowpack_kernel<128><<<dim3((W + M + 127) / 128, H, N), NT, smem_for(128), as_stream(stream)>>>(
            src, C, H, W, M, P, Kp, BF(dst_hi), BF(dst_lo));
    else
        stage_rowpack_kernel<ST_PIX><<<dim3((W + M + ST_PIX - 1) / ST_PIX, H, N), NT, smem_for(ST_PIX), as_stream(stream)>>>(
            src, C, H, W, M, P, Kp, BF(dst_hi), BF(dst_lo));
    FCD_LAUNCH_OK();
    return FCD_OK;
}

int fcd_unstage_f32_to_nchw(const float* src, int src_ld, int N, int C, int H, int W, float* dst, int accumulate,
                            void* stream) {
    FCD_CHECK_ARG(src && dst, "fcd_unstage_f32_to_nchw: null pointer");
    const long long npix = 1LL * N * H * W;
    unstage_kernel<<<blocks_for(npix), NT, 0, as_stream(stream)>>>(src, src_ld, C, 1LL * H * W, npix, dst, accumulate);
    FCD_LAUNCH_OK();
    return FCD_OK;
}

int fcd_unstage_split_to_nchw(const void* src_hi, const void* src_lo, int src_ld, int N, int C, int H, int W, float* dst,
                               void* stream) {
    FCD_CHECK_ARG(src_hi && dst && src_ld % 8 == 0, "fcd_unstage_split_to_nchw: bad arguments");
    const long long npix = 1LL * N * H * W;
    unstage_split_kernel<<<blocks_for(npix), NT, 0, as_stream(stream)>>>(CBF(src_hi), CBF(src_lo), src_ld, C, 1LL * H * W,
                                                                          npix, dst);
    FCD_LAUNCH_OK();
    return FCD_OK;
}

int fcd_stage_nchw_to_f32(const float* src, int N, int C, int H, int W, float* dst, int dst_ld, int Cp, void* stream) {
    FCD_CHECK_ARG(src && dst && Cp % 8 == 0 && Cp >= C && dst_ld % 4 == 0, "fcd_stage_nchw_to_f32: bad arguments");
    const long long npix = 1LL * N * H * W;
    stage_f32_kernel<<<blocks_for(npix), NT, 0, as_stream(stream)>>>(src, C, Cp, 1LL * H * W, npix, dst, dst_ld);
    FCD_LAUNCH_OK();
    return FCD_OK;
}

int fcd_f32_to_split(const float* src, int src_ld, void* dst_hi, void* dst_lo, int dst_ld, long long npix, int Cp,
                     void* stream) {
    FCD_CHECK_ARG(src && dst_hi && Cp % 8 == 0 && src_ld % 4 == 0 && dst_ld % 8 == 0, "fcd_f32_to_split: bad arguments");
    f32_to_split_kernel<<<blocks_for(npix * (Cp / 8)), NT, 0, as_stream(stream)>>>(src, src_ld, BF(dst_hi), BF(dst_lo),
                                                                                    dst_ld, npix, Cp);
    FCD_LAUNCH_OK();
    return FCD_OK;
}

int fcd_add_f32(float* dst, int dst_ld, const float* src, int src_ld, long long npix, int Cp, void* stream) {
    FCD_CHECK_ARG(dst && src && Cp % 8 == 0 && dst_ld % 4 == 0 && src_ld % 4 == 0, "fcd_add_f32: bad arguments");
    add_f32_kernel<<<blocks_for(npix * (Cp / 8)), NT, 0, as_stream(stream)>>>(dst, dst_ld, src, src_ld, npix, Cp);
    FCD_LAUNCH_OK();
    return FCD_OK;
}

int fcd_bn_stats(const float* z, int z_ld, long long npix, int Cp, double* sum, double* sqsum, void* stream) {
    FCD_CHECK_ARG(z && sum && sqsum, "fcd_bn_stats: null pointer");
    const int cg = Cp / 8;
    FCD_CHECK_ARG(Cp % 8 == 0 && cg <= NT && NT % cg == 0, "fcd_bn_stats: Cp/8 must divide 256 (Cp=%d)", Cp);
    bn_stats_kernel<<<reduce_grid(npix, NT / cg), NT, 0, as_stream(stream)>>>(z, z_ld, npix, Cp, sum, sqsum);
    FCD_LAUNCH_OK();
    return FCD_OK;
}

int fcd_bn_finalize(const double* sum, const double* sqsum, double count, const float* gamma, const float* beta,
                    float* running_mean, float* running_var, int C, int Cp, float momentum, float eps, int training,
                    float* scale, float* shift, float* mean, float* invstd, void* stream) {
    FCD_CHECK_ARG(gamma && beta && running_mean && running_var && scale && shift && mean && invstd,
                  "fcd_bn_finalize: null pointer");
    FCD_CHECK_ARG(!training || (sum && sqsum && count > 0), "fcd_bn_finalize: training mode needs statistics");
    bn_finalize_kernel<<<(Cp + 127) / 128, 128, 0, as_stream(stream)>>>(sum, sqsum, count, gamma, beta, running_mean,
                                                                        running_var, C, Cp, momentum, eps, training,
                                                                        scale, shift, mean, invstd);
    FCD_LAUNCH_OK();
    return FCD_OK;
}

int fcd_bn_act_fwd(const float* z, int z_ld, const float* scale, const float* shift, int act, const float* slope_ptr,
                   float slope_const, const void* res_hi, const void* res_lo, int res_ld, void* out_hi, void* out_lo,
                   int out_ld, long long npix, int Cp, void* stream) {
    FCD_CHECK_ARG(z && out_hi && Cp % 8 == 0 && z_ld % 4 == 0 && out_ld % 8 == 0, "fcd_bn_act_fwd: bad arguments");
    FCD_CHECK_ARG((scale == nullptr) == (shift == nullptr), "fcd_bn_act_fwd: scale and shift go together");
    bn_act_fwd_kernel<<<blocks_for(npix * (Cp / 8)), NT, 0, as_stream(stream)>>>(
        z, z_ld, scale, shift, act, slope_ptr, slope_const, CBF(res_hi), CBF(res_lo), res_ld, BF(out_hi), BF(out_lo),
        out_ld, npix, Cp);
    FCD_LAUNCH_OK();
    return FCD_OK;
}

int fcd_bn_act_bwd_reduce(const float* da, int da_ld, const float* z, int z_ld, const float* scale, const float* shift,
                          const float* mean, const float* invstd, int act, const float* slope_ptr, float slope_const,
                          long long npix, int Cp, double* s1, double* s2, double* dslope, void* stream) {
    FCD_CHECK_ARG(da && z && s1 && s2, "fcd_bn_act_bwd_reduce: null pointer");
    const int cg = Cp / 8;
    FCD_CHECK_ARG(Cp % 8 == 0 && cg <= NT && NT % cg == 0, "fcd_bn_act_bwd_reduce: Cp/8 must divide 256 (Cp=%d)", Cp);
    bn_act_bwd_reduce_kernel<<<reduce_grid(npix, NT / cg), NT, 0, as_stream(stream)>>>(
        da, da_ld, z, z_ld, scale, shift, mean, invstd, act, slope_ptr, slope_const, npix, Cp, s1, s2, dslope);
    FCD_LAUNCH_OK();
    return FCD_OK;
}

int fcd_bn_bwd_finalize(const double* s1, const double* s2, double count, int training, int C, int Cp, float* c1,
                        float* c2, float* dgamma, float* dbeta, int accumulate, const double* ds, float* dslope,
                        const float* scale, const double* zsum, const float* mean, const float* invstd, float* dbias,
                        int dbias_accumulate, const double* global_s1, const double* global_s2, double global_count,
                        void* stream) {
    FCD_CHECK_ARG(s1 && s2 && c1 && c2 && count > 0, "fcd_bn_bwd_finalize: bad arguments");
    FCD_CHECK_ARG((global_s1 == nullptr) == (global_s2 == nullptr) && (!global_s1 || global_count > 0),
                  "fcd_bn_bwd_finalize: global sums go together with a positive global count");
    bn_bwd_finalize_kernel<<<(Cp + 127) / 128, 128, 0, as_stream(stream)>>>(s1, s2, count, training, C, Cp, c1, c2,
                                                                            dgamma, dbeta, accumulate, ds, dslope, scale,
                                                                            zsum, mean, invstd, dbias, dbias_accumulate,
                                                                            global_s1, global_s2, global_count);
    FCD_LAUNCH_OK();
    return FCD_OK;
}

int fcd_bn_act_bwd_apply(const float* da, int da_ld, const float* z, int z_ld, const float* scale, const float* shift,
                         const float* mean, const float* invstd, const float* c1, const float* c2, int act,
                         const float* slope_ptr, float slope_const, void* dz_hi, void* dz_lo, int dz_ld, long long npix,
                         int Cp, void* stream) {
    FCD_CHECK_ARG(da && dz_hi && Cp % 8 == 0, "fcd_bn_act_bwd_apply: bad arguments");
    FCD_CHECK_ARG(!scale || (z && shift && mean && invstd && c1 && c2), "fcd_bn_act_bwd_apply: BN path needs all vectors");
    FCD_CHECK_ARG(act == FCD_ACT_NONE || z, "fcd_bn_act_bwd_apply: activation backward needs z");
    const long long pslots = (npix + APPLY_PIX - 1) / APPLY_PIX;           // pixel slots, each walks APPLY_PIX pixels
    bn_act_bwd_apply_kernel<<<blocks_for(pslots * (Cp / 8)), NT, 0, as_stream(stream)>>>(
        da, da_ld, z, z_ld, scale, shift, mean, invstd, c1, c2, act, slope_ptr, slope_const, BF(dz_hi), BF(dz_lo), dz_ld,
        npix, Cp, pslots);
    FCD_LAUNCH_OK();
    return FCD_OK;
}

int fcd_maxpool2_fwd(const void* in_hi, const void* in_lo, int in_ld, int N, int H, int W, int Cp, void* out_hi,
                     void* out_lo, int out_ld, void* stream) {
    FCD_CHECK_ARG(in_hi && out_hi && H >= 2 && W >= 2 && Cp % 8 == 0, "fcd_maxpool2_fwd: bad arguments");
    const long long total = 1LL * N * (H / 2) * (W / 2) * (Cp / 8);
    if (total < IDX32_MAX)     // 32-bit index arithmetic: a 64-bit division chain per thread costs more than its 64 bytes of traffic
        maxpool_fwd_kernel<unsigned><<<blocks_for(total), NT, 0, as_stream(stream)>>>(
            CBF(in_hi), CBF(in_lo), in_ld, N, H, W, Cp, BF(out_hi), BF(out_lo), out_ld);
    else
        maxpool_fwd_kernel<long long><<<blocks_for(total), NT, 0, as_stream(stream)>>>(
            CBF(in_hi), CBF(in_lo), in_ld, N, H, W, Cp, BF(out_hi), BF(out_lo), out_ld);
    FCD_LAUNCH_OK();
    return FCD_OK;
}

int fcd_maxpool2_bwd(const float* d_out, int dout_ld, const void* in_hi, const void* in_lo, int in_ld, int N, int H,
                     int W, int Cp, float* d_in, int din_ld, int accumulate, void* stream) {
    FCD_CHECK_ARG(d_out && in_hi && d_in && Cp % 8 == 0, "fcd_maxpool2_bwd: bad arguments");
    FCD_CHECK_ARG(H >= 2 && W >= 2, "fcd_maxpool2_bwd: needs H, W >= 2");
    const long long total = 1LL * N * (H / 2) * (W / 2) * (Cp / 8);
    if (total < IDX32_MAX)
        maxpool_bwd_kernel<unsigned><<<blocks_for(total), NT, 0, as_stream(stream)>>>(
            d_out, dout_ld, CBF(in_hi), CBF(in_lo), in_ld, N, H, W, Cp, d_in, din_ld, accumulate);
    else
        maxpool_bwd_kernel<long long><<<blocks_for(total), NT, 0, as_stream(stream)>>>(
            d_out, dout_ld, CBF(in_hi), CBF(in_lo), in_ld, N, H, W, Cp, d_in, din_ld, accumulate);
    FCD_LAUNCH_OK();
    return FCD_OK;
}

int fcd_upsample2x_bilinear_fwd(const void* in_hi, const void* in_lo, int in_ld, int N, int h, int w, int Cp,
                                void* out_hi, void* out_lo, int out_ld, int H, int W, int pad_top, int pad_left,
                                void* stream) {
    FCD_CHECK_ARG(in_hi && out_hi && Cp % 8 == 0 && H >= 2 * h + pad_top && W >= 2 * w + pad_left && pad_top >= 0 &&
                      pad_left >= 0,
                  "fcd_upsample2x_bilinear_fwd: bad arguments");
    const long long total = 1LL * N * H * W * (Cp / 8);
    if (total < IDX32_MAX)
        upsample_fwd_kernel<unsigned><<<blocks_for(total), NT, 0, as_stream(stream)>>>(
            CBF(in_hi), CBF(in_lo), in_ld, N, h, w, Cp, BF(out_hi), BF(out_lo), out_ld, H, W, pad_top, pad_left);
    else
        upsample_fwd_kernel<long long><<<blocks_for(total), NT, 0, as_stream(stream)>>>(
            CBF(in_hi), CBF(in_lo), in_ld, N, h, w, Cp, BF(out_hi), BF(out_lo), out_ld, H, W, pad_top, pad_left);
    FCD_LAUNCH_OK();
    return FCD_OK;
}

int fcd_upsample2x_bilinear_bwd(const float* d_out, int dout_ld, int N, int h, int w, int Cp, int H, int W, int pad_top,
                                int pad_left, float* d_in, int din_ld, void* stream) {
    FCD_CHECK_ARG(d_out && d_in && Cp % 8 == 0, "fcd_upsample2x_bilinear_bwd: bad arguments");
    const long long total = 1LL * N * h * w * (Cp / 8);
    if (total < IDX32_MAX)
        upsample_bwd_kernel<unsigned><<<blocks_for(total), NT, 0, as_stream(stream)>>>(
            d_out, dout_ld, N, h, w, Cp, H, W, pad_top, pad_left, d_in, din_ld);
    else
        upsample_bwd_kernel<long long><<<blocks_for(total), NT, 0, as_stream(stream)>>>(
            d_out, dout_ld, N, h, w, Cp, H, W, pad_top, pad_left, d_in, din_ld);
    FCD_LAUNCH_OK();
    return FCD_OK;
}

int fcd_convT2x2_shuffle_fwd(const float* src, long long plane_stride, int src_ld, int N, int h, int w, int Cp,
                             void* out_hi, void* out_lo, int out_ld, int H, int W, int pad_top, int pad_left,
                             void* stream) {
    FCD_CHECK_ARG(src && out_hi && Cp % 8 == 0, "fcd_convT2x2_shuffle_fwd: bad arguments");
    shuffle_up_kernel<<<blocks_for(1LL * N * H * W * (Cp / 8)), NT, 0, as_stream(stream)>>>(
        src, plane_stride, src_ld, N, h, w, Cp, BF(out_hi), BF(out_lo), out_ld, H, W, pad_top, pad_left);
    FCD_LAUNCH_OK();
    return FCD_OK;
}

int fcd_convT2x2_shuffle_bwd(const float* d_out, int dout_ld, int N, int h, int w, int Cp, int H, int W, int pad_top,
                             int pad_left, void* g_hi, void* g_lo, long long plane_stride, int g_ld, void* stream) {
    FCD_CHECK_ARG(d_out && g_hi && Cp % 8 == 0, "fcd_convT2x2_shuffle_bwd: bad arguments");
    shuffle_down_kernel<<<blocks_for(4LL * N * h * w * (Cp / 8)), NT, 0, as_stream(stream)>>>(
        d_out, dout_ld, N, h, w, Cp, H, W, pad_top, pad_left, BF(g_hi), BF(g_lo), plane_stride, g_ld);
    FCD_LAUNCH_OK();
    return FCD_OK;
}

int fcd_mse_halves_fwd(const void* f_hi, const void* f_lo, int f_ld, long long half_elems, long long npix, int Cp, double* acc,
                       void* stream) {
    FCD_CHECK_ARG(f_hi && acc && Cp % 8 == 0 && f_ld % 8 == 0 && half_elems % 8 == 0 && npix > 0, "fcd_mse_halves_fwd: bad arguments");
    long long blocks = (npix * (Cp / 8) + NT - 1) / NT;
    const long long cap = 8LL * sm_count();
    if (blocks > cap) blocks = cap;
    mse_halves_fwd_kernel<<<static_cast<unsigned>(blocks), NT, 0, as_stream(stream)>>>(CBF(f_hi), CBF(f_lo), f_ld, half_elems, npix,
                                                                                        Cp, acc);
    FCD_LAUNCH_OK();
    return FCD_OK;
}

int fcd_mse_halves_bwd(const void* f_hi, const void* f_lo, int f_ld, long long half_elems, long long npix, int Cp,
                       const float* gout, float scale, float* grad, int grad_ld, long long grad_half_elems, int accumulate,
                       void* stream) {
    FCD_CHECK_ARG(f_hi && gout && grad && Cp % 8 == 0 && f_ld % 8 == 0 && grad_ld % 4 == 0 && half_elems % 8 == 0 &&
                      grad_half_elems % 4 == 0, "fcd_mse_halves_bwd: bad arguments");
    mse_halves_bwd_kernel<<<blocks_for(npix * (Cp / 8)), NT, 0, as_stream(stream)>>>(
        CBF(f_hi), CBF(f_lo), f_ld, half_elems, npix, Cp, gout, scale, grad, grad_ld, grad_half_elems, accumulate);
    FCD_LAUNCH_OK();
    return FCD_OK;
}

}  // extern "C"
