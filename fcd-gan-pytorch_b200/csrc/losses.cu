// losses.cu — the pixel-wise loss stack of Loss.py / ssim.py as fused, HBM-bound fp32 kernels.
//
//   * masked reconstruction loss  (CNetLoss Loss.py:73-95 [L1], CGeneratorLoss Loss.py:108-124 [MSE]):
//       m = 1 - cmap;  gen = (1/B) sum_i  sum_{c,p} crit(m (t - g)) / (C * sum_p m_i);  l1 = mean |cmap|;
//       one pass reads t, g, cmap once, produces the per-sample sums AND the masked images t*m, g*m that
//       MS-SSIM consumes; one backward pass produces d g, d cmap (numerator + denominator + l1 + SSIM paths).
//   * region loss                 (Loss.py:127-141), mean|x| / mean x^2 / mean x (Demo_WSSS.py:299,315; WGAN means)
//   * SSIM / MS-SSIM              (ssim.py:26-92,153-225): per level one tiled kernel does the separable
//       "valid" Gaussian blur of the five moments in shared memory, the cs / ssim maps and the per-(b,c)
//       spatial sums (warp shuffle + one double atomic per block); the backward kernel applies the adjoint
//       blur to the five partial-derivative maps.  2x2 average pooling between levels (ssim.py:215-216,
//       padding = size % 2, count_include_pad) has its own pair of kernels.
//
// All tensors here are NCHW fp32 boundary tensors (the loss classes sit outside the NHWC network engine).
#include "fcd_common.cuh"

namespace fcd {
namespace {

constexpr int NT = 256;
constexpr int MAX_WS = 11;  // ssim.py default win_size; larger windows are rejected (FCD_ERR_UNSUPPORTED)

__device__ __forceinline__ float block_sum(float v, float* red /* [NT/32] */) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
    if (threadIdx.x < 32) {
        t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
        t = warp_sum(t);
    }
    return t;  // valid in warp 0
}

// ---------------------------------------------------------------------------------------------------
// masked reconstruction loss
// ---------------------------------------------------------------------------------------------------
// grid (chunks, B).  sums[0][i] = numerator_i, sums[1][i] = sum_p m_i, sums[2][i] = sum_p |cmap_i|
__global__ void masked_recon_fwd_kernel(const float* __restrict__ t, const float* __restrict__ g,
                                        const float* __restrict__ cmap, int B, int C, long long HW, int kind,
                                        double* sums, float* __restrict__ tm, float* __restrict__ gm) {
    const int i = blockIdx.y;
    float num = 0.f, den = 0.f, l1 = 0.f;
    for (long long p = blockIdx.x * 1LL * blockDim.x + threadIdx.x; p < HW; p += 1LL * gridDim.x * blockDim.x) {
        const float cm = cmap[i * HW + p];
        const float m = 1.f - cm;
        den += m;
        l1 += fabsf(cm);
        for (int c = 0; c < C; ++c) {
            const long long idx = (static_cast<long long>(i) * C + c) * HW + p;
            const float a = t[idx] * m, b = g[idx] * m;
            if (tm) {
                tm[idx] = a;
                gm[idx] = b;
            }
            const float d = a - b;
            num += kind == FCD_LOSS_L1 ? fabsf(d) : d * d;
        }
    }
    __shared__ float red[NT / 32];
    num = block_sum(num, red);
    den = block_sum(den, red);
    l1 = block_sum(l1, red);
    if (threadIdx.x == 0) {
        atomicAdd(sums + i, static_cast<double>(num));
        atomicAdd(sums + B + i, static_cast<double>(den));
        atomicAdd(sums + 2 * B + i, static_cast<double>(l1));
    }
}

// out[0] = generator loss, out[1] = mean |cmap|
__global__ void masked_recon_finalize_kernel(const double* sums, int B, int C, long long HW, int kind, float* out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double gen = 0.0, l1 = 0.0;
    for (int i = 0; i < B; ++i) {
        const double num = sums[i], den = sums[B + i];
        l1 += sums[2 * B + i];
        if (kind == FCD_LOSS_MSE && den == 0.0) continue;  // Loss.py:116 skips all-changed samples
        // crit(mean over C*HW) * HW / den  (Loss.py:83,118); fp32 like the reference's scalar arithmetic
        const float mean = static_cast<float>(num / (static_cast<double>(C) * HW));
        gen += static_cast<double>(mean * static_cast<float>(HW) / static_cast<float>(den));
    }
    out[0] = static_cast<float>(gen / B);
    out[1] = static_cast<float>(l1 / (static_cast<double>(B) * HW));
}

// One pass: dg, dt (optional), dcmap from  g_gen (d loss / d generator_loss), g_l1 (d / d l1_loss) and the optional
// gradients w.r.t. the masked images (from MS-SSIM).
__global__ void masked_recon_bwd_kernel(const float* __restrict__ t, const float* __restrict__ g,
                                        const float* __restrict__ cmap, int B, int C, long long HW, int kind,
                                        const double* sums, const float* g_gen_p, const float* g_l1_p,
                                        const float* __restrict__ dtm, const float* __restrict__ dgm, float* __restrict__ dt,
                                        float* __restrict__ dg, float* __restrict__ dcmap) {
    const int i = blockIdx.y;
    const float g_gen = g_gen_p ? *g_gen_p : 0.f, g_l1 = g_l1_p ? *g_l1_p : 0.f;
    const double num = sums[i], den = sums[B + i];
    const bool skip = (kind == FCD_LOSS_MSE && den == 0.0);
    // loss_i = num / (C * den) / B
    const float s = skip ? 0.f : static_cast<float>(g_gen / (static_cast<double>(B) * C * den));
    const float dden = skip ? 0.f : static_cast<float>(-static_cast<double>(g_gen) * num / (static_cast<double>(B) * C * den * den));
    const float kl1 = g_l1 / (static_cast<float>(B) * static_cast<float>(HW));
    for (long long p = blockIdx.x * 1LL * blockDim.x + threadIdx.x; p < HW; p += 1LL * gridDim.x * blockDim.x) {
        const float cm = cmap[i * HW + p];
        const float m = 1.f - cm;
        float dm = dden;
        for (int c = 0; c < C; ++c) {
            const long long idx = (static_cast<long long>(i) * C + c) * HW + p;
            const float tv = t[idx], gv = g[idx];
            const float d = tv - gv, md = m * d;
            float dd;  // d num / d (m*d)
            if (kind == FCD_LOSS_L1)
                dd = md > 0.f ? 1.f : (md < 0.f ? -1.f : 0.f);
            else
                dd = 2.f * md;
            float gt = s * dd * m, gg = -s * dd * m;
            dm += s * dd * d;
            if (dtm) {
                const float a = dtm[idx], b = dgm[idx];
                gt += a * m;
                gg += b * m;
                dm += a * tv + b * gv;
            }
            if (dg) dg[idx] = gg;
            if (dt) dt[idx] = gt;
        }
        if (dcmap) dcmap[i * HW + p] = -dm + kl1 * (cm > 0.f ? 1.f : (cm < 0.f ? -1.f : 0.f));
    }
}

// ---------------------------------------------------------------------------------------------------
// region loss + simple means
// ---------------------------------------------------------------------------------------------------
// sums[0][i] = sum crit(cmap*R), sums[1][i] = sum R       (n elements per sample)
__global__ void region_fwd_kernel(const float* __restrict__ cmap, const float* __restrict__ region, int B, long long n,
                                  int kind, double* sums) {
    const int i = blockIdx.y;
    float s = 0.f, r = 0.f;
    for (long long p = blockIdx.x * 1LL * blockDim.x + threadIdx.x; p < n; p += 1LL * gridDim.x * blockDim.x) {
        const float R = region[i * n + p], v = cmap[i * n + p] * R;
        r += R;
        s += kind == FCD_LOSS_L1 ? fabsf(v) : v * v;
    }
    __shared__ float red[NT / 32];
    s = block_sum(s, red);
    r = block_sum(r, red);
    if (threadIdx.x == 0) {
        atomicAdd(sums + i, static_cast<double>(s));
        atomicAdd(sums + B + i, static_cast<double>(r));
    }
}
__global__ void region_finalize_kernel(const double* sums, int B, long long n, long long HW, float* out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double acc = 0.0;
    for (int i = 0; i < B; ++i) {
        if (sums[B + i] == 0.0) continue;  // Loss.py:136
        const float mean = static_cast<float>(sums[i] / static_cast<double>(n));
        acc += static_cast<double>(mean * static_cast<float>(HW) / static_cast<float>(sums[B + i]));
    }
    out[0] = static_cast<float>(acc / B);
}
__global__ void region_bwd_kernel(const float* __restrict__ cmap, const float* __restrict__ region, int B, long long n,
                                  long long HW, int kind, const double* sums, const float* gout, float* __restrict__ dcmap) {
    const int i = blockIdx.y;
    const double R = sums[B + i];
    const float k = R == 0.0 ? 0.f : static_cast<float>(static_cast<double>(*gout) * HW / (static_cast<double>(B) * n * R));
    for (long long p = blockIdx.x * 1LL * blockDim.x + threadIdx.x; p < n; p += 1LL * gridDim.x * blockDim.x) {
        const float r = region[i * n + p], v = cmap[i * n + p] * r;
        const float d = kind == FCD_LOSS_L1 ? (v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f)) : 2.f * v;
        dcmap[i * n + p] = k * d * r;
    }
}

// mode 0: mean x, 1: mean |x|, 2: mean x^2
__global__ void mean_fwd_kernel(const float* __restrict__ x, long long n, int mode, double* acc) {
    float s = 0.f;
    for (long long p = blockIdx.x * 1LL * blockDim.x + threadIdx.x; p < n; p += 1LL * gridDim.x * blockDim.x) {
        const float v = x[p];
        s += mode == 0 ? v : (mode == 1 ? fabsf(v) : v * v);
    }
    __shared__ float red[NT / 32];
    s = block_sum(s, red);
    if (threadIdx.x == 0) atomicAdd(acc, static_cast<double>(s));
}
__global__ void mean_finalize_kernel(const double* acc, long long n, float* out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = static_cast<float>(acc[0] / static_cast<double>(n));
}
__global__ void mean_bwd_kernel(const float* __restrict__ x, long long n, int mode, const float* gout, float* __restrict__ dx) {
    const float k = *gout / static_cast<float>(n);
    for (long long p = blockIdx.x * 1LL * blockDim.x + threadIdx.x; p < n; p += 1LL * gridDim.x * blockDim.x) {
        const float v = x[p];
        dx[p] = mode == 0 ? k : (mode == 1 ? k * (v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f)) : 2.f * k * v);
    }
}

// ---------------------------------------------------------------------------------------------------
// SSIM level
// ---------------------------------------------------------------------------------------------------
constexpr int TW = 32, TH = 16;
constexpr int IN_W = TW + MAX_WS - 1, IN_H = TH + MAX_WS - 1;  // 42 x 26

struct SsimDims {
    int H, W, OH, OW;   // input and blurred ("valid") sizes
    int wh, ww;         // effective window sizes per dimension (1 = dimension skipped, ssim.py:45-50)
    float C1, C2;
};

// partial derivatives of f (cs or ssim) w.r.t. (mu1, mu2, e11, e22, e12), e = blurred second moments
__device__ __forceinline__ void ssim_point(float mu1, float mu2, float e11, float e22, float e12, float C1, float C2,
                                           float& ssim, float& cs, int which, float (&d)[5]) {
    const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
    const float s1 = e11 - mu1_sq, s2 = e22 - mu2_sq, s12 = e12 - mu12;
    const float A2 = 2.f * s12 + C2, B2 = s1 + s2 + C2;
    const float A1 = 2.f * mu12 + C1, B1 = mu1_sq + mu2_sq + C1;
    cs = A2 / B2;
    const float L = A1 / B1;
    ssim = L * cs;
    const float iB2 = 1.f / B2;
    // cs partials
    float dcs_mu1 = (-2.f * mu2 + 2.f * mu1 * cs) * iB2;
    float dcs_mu2 = (-2.f * mu1 + 2.f * mu2 * cs) * iB2;
    float dcs_e11 = -cs * iB2, dcs_e22 = -cs * iB2, dcs_e12 = 2.f * iB2;
    if (which == 0) {
        d[0] = dcs_mu1; d[1] = dcs_mu2; d[2] = dcs_e11; d[3] = dcs_e22; d[4] = dcs_e12;
    } else {
        const float iB1 = 1.f / B1;
        const float dL_mu1 = (2.f * mu2 - 2.f * mu1 * L) * iB1, dL_mu2 = (2.f * mu1 - 2.f * mu2 * L) * iB1;
        d[0] = cs * dL_mu1 + L * dcs_mu1;
        d[1] = cs * dL_mu2 + L * dcs_mu2;
        d[2] = L * dcs_e11; d[3] = L * dcs_e22; d[4] = L * dcs_e12;
    }
}

// grid (tiles_x, tiles_y, planes).  sums: [0][plane] += sum ssim_map, [1][plane] += sum cs_map.
// dmaps (optional): [5][planes][OH][OW] partial derivatives of `which` (0 cs, 1 ssim).
__global__ void __launch_bounds__(NT)
ssim_fwd_kernel(const float* __restrict__ X, const float* __restrict__ Y, SsimDims d, const float* __restrict__ win_h,
                const float* __restrict__ win_w, int planes, double* sums, float* __restrict__ dmaps, int which) {
    __shared__ float sx[IN_H][IN_W + 1], sy[IN_H][IN_W + 1];
    __shared__ float hz[5][IN_H][TW + 1];
    __shared__ float gh[MAX_WS], gw[MAX_WS];
    __shared__ float red[NT / 32];
    const int plane = blockIdx.z;
    const int r0 = blockIdx.y * TH, c0 = blockIdx.x * TW;
    const float* xp = X + static_cast<size_t>(plane) * d.H * d.W;
    const float* yp = Y + static_cast<size_t>(plane) * d.H * d.W;
    if (threadIdx.x < d.wh) gh[threadIdx.x] = d.wh == 1 ? 1.f : win_h[threadIdx.x];
    if (threadIdx.x >= 32 && threadIdx.x < 32 + d.ww) gw[threadIdx.x - 32] = d.ww == 1 ? 1.f : win_w[threadIdx.x - 32];
    const int in_h = TH + d.wh - 1, in_w = TW + d.ww - 1;
    for (int idx = threadIdx.x; idx < in_h * in_w; idx += NT) {
        const int r = idx / in_w, c = idx - r * in_w;
        const int gy = r0 + r, gx = c0 + c;
        const bool ok = gy < d.H && gx < d.W;
        sx[r][c] = ok ? xp[static_cast<size_t>(gy) * d.W + gx] : 0.f;
        sy[r][c] = ok ? yp[static_cast<size_t>(gy) * d.W + gx] : 0.f;
    }
    __syncthreads();
    // horizontal pass of the five moments
    for (int idx = threadIdx.x; idx < in_h * TW; idx += NT) {
        const int r = idx / TW, c = idx - r * TW;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f;
        for (int k = 0; k < d.ww; ++k) {
            const float w = gw[k], x = sx[r][c + k], y = sy[r][c + k];
            a0 = fmaf(w, x, a0);
            a1 = fmaf(w, y, a1);
            a2 = fmaf(w, x * x, a2);
            a3 = fmaf(w, y * y, a3);
            a4 = fmaf(w, x * y, a4);
        }
        hz[0][r][c] = a0; hz[1][r][c] = a1; hz[2][r][c] = a2; hz[3][r][c] = a3; hz[4][r][c] = a4;
    }
    __syncthreads();
    float acc_ssim = 0.f, acc_cs = 0.f;
    for (int idx = threadIdx.x; idx < TH * TW; idx += NT) {
        const int r = idx / TW, c = idx - r * TW;
        const int oy = r0 + r, ox = c0 + c;
        if (oy >= d.OH || ox >= d.OW) continue;
        float m[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
        for (int k = 0; k < d.wh; ++k) {
            const float w = gh[k];
#pragma unroll
            for (int j = 0; j < 5; ++j) m[j] = fmaf(w, hz[j][r + k][c], m[j]);
        }
        float ssim, cs, dd[5];
        ssim_point(m[0], m[1], m[2], m[3], m[4], d.C1, d.C2, ssim, cs, which, dd);
        acc_ssim += ssim;
        acc_cs += cs;
        if (dmaps) {
            const size_t stride = static_cast<size_t>(planes) * d.OH * d.OW;
            const size_t o = (static_cast<size_t>(plane) * d.OH + oy) * d.OW + ox;
#pragma unroll
            for (int j = 0; j < 5; ++j) dmaps[j * stride + o] = dd[j];
        }
    }
    if (sums) {
        acc_ssim = block_sum(acc_ssim, red);
        acc_cs = block_sum(acc_cs, red);
        if (threadIdx.x == 0) {
            atomicAdd(sums + plane, static_cast<double>(acc_ssim));
            atomicAdd(sums + planes + plane, static_cast<double>(acc_cs));
        }
    }
}

// Adjoint blur of the five derivative maps, combined into dX, dY:
//   dX(p) = coef * (T[mu1] + 2 X(p) T[e11] + Y(p) T[e12]),  dY(p) = coef * (T[mu2] + 2 Y(p) T[e22] + X(p) T[e12]),
//   T[m](p) = sum_{ky,kx} gh[ky] gw[kx] m(p - k)   (zero outside the blurred map).
__global__ void __launch_bounds__(NT)
ssim_bwd_kernel(const float* __restrict__ dmaps, const float* __restrict__ X, const float* __restrict__ Y, SsimDims d,
                const float* __restrict__ win_h, const float* __restrict__ win_w, int planes, const float* __restrict__ coef,
                float* __restrict__ dX, float* __restrict__ dY, int accumulate) {
    __shared__ float sm[5][IN_H][IN_W + 1];
    __shared__ float hz[5][IN_H][TW + 1];
    __shared__ float gh[MAX_WS], gw[MAX_WS];
    const int plane = blockIdx.z;
    const int r0 = blockIdx.y * TH, c0 = blockIdx.x * TW;
    if (threadIdx.x < d.wh) gh[threadIdx.x] = d.wh == 1 ? 1.f : win_h[threadIdx.x];
    if (threadIdx.x >= 32 && threadIdx.x < 32 + d.ww) gw[threadIdx.x - 32] = d.ww == 1 ? 1.f : win_w[threadIdx.x - 32];
    const int in_h = TH + d.wh - 1, in_w = TW + d.ww - 1;
    const size_t stride = static_cast<size_t>(planes) * d.OH * d.OW;
    const float* base = dmaps + static_cast<size_t>(plane) * d.OH * d.OW;
    // tile row r <-> blurred row qy = r0 - (wh-1) + r ; col c <-> qx = c0 - (ww-1) + c
    for (int idx = threadIdx.x; idx < in_h * in_w; idx += NT) {
        const int r = idx / in_w, c = idx - r * in_w;
        const int qy = r0 - (d.wh - 1) + r, qx = c0 - (d.ww - 1) + c;
        const bool ok = qy >= 0 && qy < d.OH && qx >= 0 && qx < d.OW;
        const size_t o = ok ? static_cast<size_t>(qy) * d.OW + qx : 0;
#pragma unroll
        for (int j = 0; j < 5; ++j) sm[j][r][c] = ok ? base[j * stride + o] : 0.f;
    }
    __syncthreads();
    // horizontal: Th(r, x) = sum_k gw[k] m(r, x - k)  -> tile col (x - c0) + (ww-1) - k
    for (int idx = threadIdx.x; idx < in_h * TW; idx += NT) {
        const int r = idx / TW, c = idx - r * TW;
        float a[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
        for (int k = 0; k < d.ww; ++k) {
            const float w = gw[k];
#pragma unroll
            for (int j = 0; j < 5; ++j) a[j] = fmaf(w, sm[j][r][c + d.ww - 1 - k], a[j]);
        }
#pragma unroll
        for (int j = 0; j < 5; ++j) hz[j][r][c] = a[j];
    }
    __syncthreads();
    const float cf = coef[plane];
    for (int idx = threadIdx.x; idx < TH * TW; idx += NT) {
        const int r = idx / TW, c = idx - r * TW;
        const int y = r0 + r, x = c0 + c;
        if (y >= d.H || x >= d.W) continue;
        float T[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
        for (int k = 0; k < d.wh; ++k) {
            const float w = gh[k];
#pragma unroll
            for (int j = 0; j < 5; ++j) T[j] = fmaf(w, hz[j][r + d.wh - 1 - k][c], T[j]);
        }
        const size_t o = (static_cast<size_t>(plane) * d.H + y) * d.W + x;
        const float xv = X[o], yv = Y[o];
        const float gx = cf * (T[0] + 2.f * xv * T[2] + yv * T[4]);
        const float gy = cf * (T[1] + 2.f * yv * T[3] + xv * T[4]);
        dX[o] = accumulate ? dX[o] + gx : gx;
        dY[o] = accumulate ? dY[o] + gy : gy;
    }
}

// avg_pool2d(kernel 2, stride 2, padding (ph, pw), count_include_pad=True) — ssim.py:215-216
__global__ void avgpool2_fwd_kernel(const float* __restrict__ in, int planes, int H, int W, int ph, int pw, int OH, int OW,
                                    float* __restrict__ out) {
    const long long idx = blockIdx.x * 1LL * blockDim.x + threadIdx.x;
    if (idx >= 1LL * planes * OH * OW) return;
    const int ox = static_cast<int>(idx % OW);
    const int oy = static_cast<int>((idx / OW) % OH);
    const long long pl = idx / (1LL * OW * OH);
    const float* p = in + pl * H * W;
    float s = 0.f;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const int y = 2 * oy - ph + a, x = 2 * ox - pw + b;
            if (y >= 0 && y < H && x >= 0 && x < W) s += p[static_cast<size_t>(y) * W + x];
        }
    out[idx] = 0.25f * s;
}
__global__ void avgpool2_bwd_kernel(const float* __restrict__ dout, int planes, int H, int W, int ph, int pw, int OH, int OW,
                                    float* __restrict__ din, int accumulate) {
    const long long idx = blockIdx.x * 1LL * blockDim.x + threadIdx.x;
    if (idx >= 1LL * planes * H * W) return;
    const int x = static_cast<int>(idx % W);
    const int y = static_cast<int>((idx / W) % H);
    const long long pl = idx / (1LL * W * H);
    const int oy = (y + ph) >> 1, ox = (x + pw) >> 1;
    float g = 0.f;
    if (oy < OH && ox < OW) g = 0.25f * dout[(pl * OH + oy) * OW + ox];
    din[idx] = accumulate ? din[idx] + g : g;
}

// MS-SSIM combine (ssim.py:207-225) / single-scale SSIM tail (ssim.py:143-150).
//   sums [levels][2][planes] (ssim sums, cs sums); counts[levels] = OH*OW of the level.
//   level l < levels-1 uses cs, the last level uses ssim; v_l = relu(mean) (relu optional for levels == 1);
//   prod = PI v_l^w_l;  out[0] = mean over planes (size_average) or out[b] = mean over the C planes of image b.
__global__ void msssim_combine_fwd_kernel(const double* sums, const double* counts, const float* weights, int levels,
                                          int planes, int C, int size_average, int use_relu, float* prod, float* out) {
    // one block; planes is small (B*C)
    for (int p = threadIdx.x; p < planes; p += blockDim.x) {
        float pr = 1.f;
        for (int l = 0; l < levels; ++l) {
            const double s = sums[(static_cast<size_t>(l) * 2 + (l == levels - 1 ? 0 : 1)) * planes + p];
            float v = static_cast<float>(s / counts[l]);
            if (use_relu && v < 0.f) v = 0.f;
            pr *= (levels == 1) ? v : powf(v, weights[l]);
        }
        prod[p] = pr;
    }
    __syncthreads();
    if (size_average) {
        __shared__ float red[NT / 32];
        float s = 0.f;
        for (int p = threadIdx.x; p < planes; p += blockDim.x) s += prod[p];
        s = block_sum(s, red);
        if (threadIdx.x == 0) out[0] = s / planes;
    } else {
        const int B = planes / C;
        for (int b = threadIdx.x; b < B; b += blockDim.x) {
            float s = 0.f;
            for (int c = 0; c < C; ++c) s += prod[b * C + c];
            out[b] = s / C;
        }
    }
}
// coef[l][p] = d out / d (spatial SUM of level l's map of plane p)
__global__ void msssim_combine_bwd_kernel(const double* sums, const double* counts, const float* weights, int levels,
                                          int planes, int C, int size_average, int use_relu, const float* prod,
                                          const float* gout, float* coef) {
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < planes; p += gridDim.x * blockDim.x) {
        const float gp = size_average ? gout[0] / planes : gout[p / C] / C;
        for (int l = 0; l < levels; ++l) {
            const double s = sums[(static_cast<size_t>(l) * 2 + (l == levels - 1 ? 0 : 1)) * planes + p];
            const float mean = static_cast<float>(s / counts[l]);
            float dv;  // d prod / d mean_l
            if (levels == 1) {
                dv = (use_relu && !(mean > 0.f)) ? 0.f : 1.f;
            } else if (!(mean > 0.f)) {
                dv = 0.f;  // relu masks the (infinite) power derivative, like torch's threshold_backward
            } else {
                dv = weights[l] * prod[p] / mean;
            }
            coef[static_cast<size_t>(l) * planes + p] = gp * dv / static_cast<float>(counts[l]);
        }
    }
}

inline unsigned blocks_for(long long n) { return static_cast<unsigned>((n + NT - 1) / NT); }
inline unsigned chunks_for(long long n, int B) {
    long long c = (n + NT * 4 - 1) / (NT * 4);
    const long long cap = (8LL * sm_count() + B - 1) / B;
    if (c > cap) c = cap;
    return static_cast<unsigned>(c < 1 ? 1 : c);
}

}  // namespace
}  // namespace fcd

using namespace fcd;

extern "C" {

int fcd_masked_recon_fwd(const float* t, const float* g, const float* cmap, int B, int C, int H, int W, int kind,
                         double* sums, float* out2, float* tm, float* gm, void* stream) {
    FCD_CHECK_ARG(t && g && cmap && sums && out2 && ((tm == nullptr) == (gm == nullptr)), "fcd_masked_recon_fwd: bad pointers");
    FCD_CHECK_ARG(kind == FCD_LOSS_L1 || kind == FCD_LOSS_MSE, "fcd_masked_recon_fwd: kind");
    cudaStream_t s = as_stream(stream);
    const long long HW = 1LL * H * W;
    FCD_CUDA_OK(cudaMemsetAsync(sums, 0, sizeof(double) * 3 * B, s));
    masked_recon_fwd_kernel<<<dim3(chunks_for(HW, B), B), NT, 0, s>>>(t, g, cmap, B, C, HW, kind, sums, tm, gm);
    masked_recon_finalize_kernel<<<1, 32, 0, s>>>(sums, B, C, HW, kind, out2);
    FCD_LAUNCH_OK();
    return FCD_OK;
}

int fcd_masked_recon_bwd(const float* t, const float* g, const float* cmap, int B, int C, int H, int W, int kind,
                         const double* sums, const float* g_gen, const float* g_l1, const float* dtm, const float* dgm,
                         float* dt, float* dg, float* dcmap, void* stream) {
    FCD_CHECK_ARG(t && g && cmap && sums && ((dtm == nullptr) == (dgm == nullptr)), "fcd_masked_recon_bwd: bad pointers");
    const long long HW = 1LL * H * W;
    masked_recon_bwd_kernel<<<dim3(chunks_for(HW, B), B), NT, 0, as_stream(stream)>>>(t, g, cmap, B, C, HW, kind, sums, g_gen,
                                                                                       g_l1, dtm, dgm, dt, dg, dcmap);
    FCD_LAUNCH_OK();
    return FCD_OK;
}

int fcd_region_loss_fwd(const float* cmap, const float* region, int B, long long n, long long HW, int kind, double* sums,
                        float* out, void* stream) {
    FCD_CHECK_ARG(cmap && region && sums && out, "fcd_region_loss_fwd: null pointer");
    cudaStream_t s = as_stream(stream);
    FCD_CUDA_OK(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * B, s));
    region_fwd_kernel<<<dim3(chunks_for(n, B), B), NT, 0, s>>>(cmap, region, B, n, kind, sums);
    region_finalize_kernel<<<1, 32, 0, s>>>(sums, B, n, HW, out);
    FCD_LAUNCH_OK();
    return FCD_OK;
}

int fcd_region_loss_bwd(const float* cmap, const float* region, int B, long long n, long long HW, int kind,
                        const double* sums, const float* gout, float* dcmap, void* stream) {
    FCD_CHECK_ARG(cmap && region && sums && gout && dcmap, "fcd_region_loss_bwd: null pointer");
    region_bwd_kernel<<<dim3(chunks_for(n, B), B), NT, 0, as_stream(stream)>>>(cmap, region, B, n, HW, kind, sums, gout, dcmap);
    FCD_LAUNCH_OK();
    return FCD_OK;
}

int fcd_mean_fwd(const float* x, long long n, int mode, double* acc, float* out, void* stream) {
    FCD_CHECK_ARG(x && acc && out && n > 0 && mode >= 0 && mode <= 2, "fcd_mean_fwd: bad arguments");
    cudaStream_t s = as_stream(stream);
    FCD_CUDA_OK(cudaMemsetAsync(acc, 0, sizeof(double), s));
    mean_fwd_kernel<<<chunks_for(n, 1), NT, 0, s>>>(x, n, mode, acc);
    mean_finalize_kernel<<<1, 32, 0, s>>>(acc, n, out);
    FCD_LAUNCH_OK();
    return FCD_OK;
}

int fcd_mean_bwd(const float* x, long long n, int mode, const float* gout, float* dx, void* stream) {
    FCD_CHECK_ARG(x && gout && dx, "fcd_mean_bwd: null pointer");
    mean_bwd_kernel<<<chunks_for(n, 1), NT, 0, as_stream(stream)>>>(x, n, mode, gout, dx);
    FCD_LAUNCH_OK();
    return FCD_OK;
}

static int ssim_dims(int H, int W, int win_size, float C1, float C2, SsimDims* d) {
    FCD_CHECK_ARG(win_size >= 1 && (win_size & 1), "ssim: window size must be odd");
    if (win_size > MAX_WS) {
        set_error(FCD_ERR_UNSUPPORTED, "ssim: win_size %d > %d is not supported by the fused kernel", win_size, MAX_WS);
        return FCD_ERR_UNSUPPORTED;
    }
    d->H = H; d->W = W;
    d->wh = H >= win_size ? win_size : 1;   // gaussian_filter skips a dimension smaller than the window
    d->ww = W >= win_size ? win_size : 1;
    d->OH = H - d->wh + 1; d->OW = W - d->ww + 1;
    d->C1 = C1; d->C2 = C2;
    return FCD_OK;
}

int fcd_ssim_level_fwd(const float* X, const float* Y, int planes, int H, int W, const float* win, int win_size, float C1,
                       float C2, double* sums, float* dmaps, int which, void* stream) {
    FCD_CHECK_ARG(X && Y && win && planes > 0 && (sums || dmaps), "fcd_ssim_level_fwd: bad arguments");
    SsimDims d;
    int rc = ssim_dims(H, W, win_size, C1, C2, &d);
    if (rc) return rc;
    cudaStream_t s = as_stream(stream);
    if (sums) FCD_CUDA_OK(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * planes, s));
    dim3 grid((d.OW + TW - 1) / TW, (d.OH + TH - 1) / TH, planes);
    ssim_fwd_kernel<<<grid, NT, 0, s>>>(X, Y, d, win, win, planes, sums, dmaps, which);
    FCD_LAUNCH_OK();
    return FCD_OK;
}

int fcd_ssim_level_bwd(const float* dmaps, const float* X, const float* Y, int planes, int H, int W, const float* win,
                       int win_size, const float* coef, float* dX, float* dY, int accumulate, void* stream) {
    FCD_CHECK_ARG(dmaps && X && Y && win && coef && dX && dY, "fcd_ssim_level_bwd: null pointer");
    SsimDims d;
    int rc = ssim_dims(H, W, win_size, 0.f, 0.f, &d);
    if (rc) return rc;
    dim3 grid((W + TW - 1) / TW, (H + TH - 1) / TH, planes);
    ssim_bwd_kernel<<<grid, NT, 0, as_stream(stream)>>>(dmaps, X, Y, d, win, win, planes, coef, dX, dY, accumulate);
    FCD_LAUNCH_OK();
    return FCD_OK;
}

int fcd_avgpool2_fwd(const float* in, int planes, int H, int W, int pad_h, int pad_w, float* out, void* stream) {
    FCD_CHECK_ARG(in && out && pad_h >= 0 && pad_h <= 1 && pad_w >= 0 && pad_w <= 1, "fcd_avgpool2_fwd: bad arguments");
    const int OH = (H + 2 * pad_h - 2) / 2 + 1, OW = (W + 2 * pad_w - 2) / 2 + 1;
    avgpool2_fwd_kernel<<<blocks_for(1LL * planes * OH * OW), NT, 0, as_stream(stream)>>>(in, planes, H, W, pad_h, pad_w, OH,
                                                                                           OW, out);
    FCD_LAUNCH_OK();
    return FCD_OK;
}

int fcd_avgpool2_bwd(const float* dout, int planes, int H, int W, int pad_h, int pad_w, float* din, int accumulate,
                     void* stream) {
    FCD_CHECK_ARG(dout && din, "fcd_avgpool2_bwd: null pointer");
    const int OH = (H + 2 * pad_h - 2) / 2 + 1, OW = (W + 2 * pad_w - 2) / 2 + 1;
    avgpool2_bwd_kernel<<<blocks_for(1LL * planes * H * W), NT, 0, as_stream(stream)>>>(dout, planes, H, W, pad_h, pad_w, OH,
                                                                                         OW, din, accumulate);
    FCD_LAUNCH_OK();
    return FCD_OK;
}

int fcd_msssim_combine_fwd(const double* sums, const double* counts, const float* weights, int levels, int planes, int C,
                           int size_average, int use_relu, float* prod, float* out, void* stream) {
    FCD_CHECK_ARG(sums && counts && weights && prod && out && levels >= 1 && planes % C == 0, "fcd_msssim_combine_fwd: bad arguments");
    msssim_combine_fwd_kernel<<<1, NT, 0, as_stream(stream)>>>(sums, counts, weights, levels, planes, C, size_average, use_relu,
                                                               prod, out);
    FCD_LAUNCH_OK();
    return FCD_OK;
}

int fcd_msssim_combine_bwd(const double* sums, const double* counts, const float* weights, int levels, int planes, int C,
                           int size_average, int use_relu, const float* prod, const float* gout, float* coef, void* stream) {
    FCD_CHECK_ARG(sums && counts && weights && prod && gout && coef, "fcd_msssim_combine_bwd: null pointer");
    msssim_combine_bwd_kernel<<<(planes + NT - 1) / NT, NT, 0, as_stream(stream)>>>(sums, counts, weights, levels, planes, C,
                                                                                    size_average, use_relu, prod, gout, coef);
    FCD_LAUNCH_OK();
    return FCD_OK;
}

}  // extern "C"
