// heads.cu — the small heads and the inline masking arithmetic of the training loops:
//   * OutConv: Conv2d 1x1 (128 -> n_out) + Sigmoid producing the change-density map   (Module.py:82-90)
//   * Discriminator classifier: AdaptiveAvgPool2d(1) of (fx - fy), two 1x1 convs (= dense layers) with
//     LeakyReLU(0.2) between, sigmoid                                                 (Module.py:212-223)
//   * soft masking x*(1-cmap) and the fake-unchanged synthesis y*(1-region)+x*region
//                                                       (Demo_RSSS.py:290-300, Demo_WSSS.py:261-279)
// All are HBM-bound: one pass over the activations, warp-shuffle reductions, double atomics for the
// few parameter-gradient sums.
#include "fcd_common.cuh"

namespace fcd {
namespace {

constexpr int NT = 256;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// ---- OutConv -------------------------------------------------------------------------------------
// one warp per pixel; lane handles channels lane*4 .. +3 (+128 per round); n_out <= 4
__global__ void outconv_fwd_kernel(const __nv_bfloat16* x_hi, const __nv_bfloat16* x_lo, int x_ld, int Cin,
                                   const float* __restrict__ w, const float* __restrict__ b, int n_out, long long npix,
                                   long long HW, float* __restrict__ out /* NCHW (N,n_out,H,W) */) {
    const int lane = threadIdx.x & 31;
    const long long warp = (blockIdx.x * 1LL * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = (gridDim.x * 1LL * blockDim.x) >> 5;
    for (long long pix = warp; pix < npix; pix += nwarps) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int c = lane * 4; c < Cin; c += 128) {
            const size_t off = static_cast<size_t>(pix) * x_ld + c;
            float v[4];
            const uint2 h = *reinterpret_cast<const uint2*>(x_hi + off);
            const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&h);
            float2 a = __bfloat1622float2(hp[0]), bb = __bfloat1622float2(hp[1]);
            v[0] = a.x; v[1] = a.y; v[2] = bb.x; v[3] = bb.y;
            if (x_lo) {
                const uint2 l = *reinterpret_cast<const uint2*>(x_lo + off);
                const __nv_bfloat162* lp = reinterpret_cast<const __nv_bfloat162*>(&l);
                a = __bfloat1622float2(lp[0]); bb = __bfloat1622float2(lp[1]);
                v[0] += a.x; v[1] += a.y; v[2] += bb.x; v[3] += bb.y;
            }
            for (int o = 0; o < n_out; ++o) {
                const float4 ww = *reinterpret_cast<const float4*>(w + o * Cin + c);
                acc[o] += v[0] * ww.x + v[1] * ww.y + v[2] * ww.z + v[3] * ww.w;
            }
        }
        for (int o = 0; o < n_out; ++o) acc[o] = warp_sum(acc[o]);
        if (lane == 0) {
            const long long n = pix / HW, p = pix - n * HW;
            for (int o = 0; o < n_out; ++o) out[(n * n_out + o) * HW + p] = sigmoidf_(acc[o] + b[o]);
        }
    }
}

// dlogit = dout * s * (1 - s);  dx[pix][c] = sum_o dlogit_o w[o][c];  dw[o][c] += sum_pix dlogit_o x[pix][c];  db[o] += sum dlogit_o
__global__ void outconv_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ out, const __nv_bfloat16* x_hi,
                                   const __nv_bfloat16* x_lo, int x_ld, int Cin, const float* __restrict__ w, int n_out,
                                   long long npix, long long HW, float* __restrict__ dx, int dx_ld, double* dw, double* db) {
    // block: 256 threads = 8 warps; thread owns channels lane*4.. (Cin <= 128) ; accumulates dw privately over its pixels
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const long long warp = blockIdx.x * 8LL + wid, nwarps = gridDim.x * 8LL;
    float dwacc[4][4];
    float dbacc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int o = 0; o < 4; ++o)
#pragma unroll
        for (int j = 0; j < 4; ++j) dwacc[o][j] = 0.f;
    const int c = lane * 4;
    float wv[4][4];
#pragma unroll
    for (int o = 0; o < 4; ++o)
#pragma unroll
        for (int j = 0; j < 4; ++j) wv[o][j] = (o < n_out && c + j < Cin) ? w[o * Cin + c + j] : 0.f;
    for (long long pix = warp; pix < npix; pix += nwarps) {
        const long long n = pix / HW, p = pix - n * HW;
        float dl[4] = {0.f, 0.f, 0.f, 0.f};
        for (int o = 0; o < n_out; ++o) {
            const float s = out[(n * n_out + o) * HW + p];
            dl[o] = dout[(n * n_out + o) * HW + p] * s * (1.f - s);
        }
        if (c < Cin) {
            const size_t off = static_cast<size_t>(pix) * x_ld + c;
            float v[4];
            const uint2 h = *reinterpret_cast<const uint2*>(x_hi + off);
            const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&h);
            float2 a = __bfloat1622float2(hp[0]), bb = __bfloat1622float2(hp[1]);
            v[0] = a.x; v[1] = a.y; v[2] = bb.x; v[3] = bb.y;
            if (x_lo) {
                const uint2 l = *reinterpret_cast<const uint2*>(x_lo + off);
                const __nv_bfloat162* lp = reinterpret_cast<const __nv_bfloat162*>(&l);
                a = __bfloat1622float2(lp[0]); bb = __bfloat1622float2(lp[1]);
                v[0] += a.x; v[1] += a.y; v[2] += bb.x; v[3] += bb.y;
            }
            float g[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int o = 0; o < 4; ++o) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    g[j] = fmaf(dl[o], wv[o][j], g[j]);
                    dwacc[o][j] = fmaf(dl[o], v[j], dwacc[o][j]);
                }
            }
            *reinterpret_cast<float4*>(dx + static_cast<size_t>(pix) * dx_ld + c) = make_float4(g[0], g[1], g[2], g[3]);
        }
        if (lane == 0)
            for (int o = 0; o < n_out; ++o) dbacc[o] += dl[o];
    }
    // combine the 8 warps of the block, then one double atomic per (o, c)
    __shared__ float red[8][4][128];
#pragma unroll
    for (int o = 0; o < 4; ++o)
#pragma unroll
        for (int j = 0; j < 4; ++j) red[wid][o][c + j] = dwacc[o][j];
    __shared__ float redb[8][4];
    if (lane == 0)
        for (int o = 0; o < 4; ++o) redb[wid][o] = dbacc[o];
    __syncthreads();
    for (int t = threadIdx.x; t < n_out * Cin; t += blockDim.x) {
        const int o = t / Cin, cc = t % Cin;
        float s = 0.f;
        for (int k = 0; k < 8; ++k) s += red[k][o][cc];
        atomicAdd(dw + o * Cin + cc, static_cast<double>(s));
    }
    if (threadIdx.x < n_out) {
        float s = 0.f;
        for (int k = 0; k < 8; ++k) s += redb[k][threadIdx.x];
        atomicAdd(db + threadIdx.x, static_cast<double>(s));
    }
}

// ---- OutConv, wide form: L = Cin / 8 lanes per pixel (16-byte loads of both planes), 32 / L pixels per warp pass -----------
// The one-warp-per-pixel kernels above keep only Cin / 4 lanes busy and pay a 64-bit division per pixel; for the Segmentor's
// Cin = 64 (Module.py:139) they ran at 1/8 of the HBM rate (r02 event tables: 0.59 ms forward, 1.45 ms backward at 32 x 256^2).
__device__ __forceinline__ void ld_split8_regs(const __nv_bfloat16* hi, const __nv_bfloat16* lo, size_t off, float v[8]) {
    const uint4 u = *reinterpret_cast<const uint4*>(hi + off);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(h[i]);
        v[2 * i] = f.x;
        v[2 * i + 1] = f.y;
    }
    if (lo) {
        const uint4 ul = *reinterpret_cast<const uint4*>(lo + off);
        const __nv_bfloat162* l = reinterpret_cast<const __nv_bfloat162*>(&ul);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 f = __bfloat1622float2(l[i]);
            v[2 * i] += f.x;
            v[2 * i + 1] += f.y;
        }
    }
}

template <int L>
__global__ void outconv_fwd_wide_kernel(const __nv_bfloat16* x_hi, const __nv_bfloat16* x_lo, int x_ld, int Cin,
                                        const float* __restrict__ w, const float* __restrict__ b, int n_out, unsigned npix,
                                        unsigned HW, float* __restrict__ out /* NCHW (N,n_out,H,W) */) {
    constexpr int PPW = 32 / L;
    const int lane = threadIdx.x & 31, sub = lane / L, l = lane % L;
    const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    float wv[4][8];
#pragma unroll
    for (int o = 0; o < 4; ++o)
#pragma unroll
        for (int j = 0; j < 8; ++j) wv[o][j] = o < n_out ? w[o * Cin + l * 8 + j] : 0.f;
    for (unsigned p0 = warp * PPW; p0 < npix; p0 += nwarps * PPW) {
        const unsigned pix = p0 + sub;
        const bool valid = pix < npix;
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (valid) ld_split8_regs(x_hi, x_lo, static_cast<size_t>(pix) * x_ld + l * 8, v);
        float acc[4];
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            float a = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) a = fmaf(v[j], wv[o][j], a);
#pragma unroll
            for (int d = L / 2; d > 0; d >>= 1) a += __shfl_xor_sync(0xffffffffu, a, d);
            acc[o] = a;
        }
        if (valid && l == 0) {
            const unsigned n = pix / HW, p = pix - n * HW;
            for (int o = 0; o < n_out; ++o) out[(static_cast<size_t>(n) * n_out + o) * HW + p] = sigmoidf_(acc[o] + b[o]);
        }
    }
}

template <int L>
__global__ void outconv_bwd_wide_kernel(const float* __restrict__ dout, const float* __restrict__ out, const __nv_bfloat16* x_hi,
                                        const __nv_bfloat16* x_lo, int x_ld, int Cin, const float* __restrict__ w, int n_out,
                                        unsigned npix, unsigned HW, float* __restrict__ dx, int dx_ld, double* dw, double* db) {
    constexpr int PPW = 32 / L;
    const int lane = threadIdx.x & 31, sub = lane / L, l = lane % L;
    const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const int c = l * 8;
    float wv[4][8], dwacc[4][8], dbacc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int o = 0; o < 4; ++o)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            wv[o][j] = o < n_out ? w[o * Cin + c + j] : 0.f;
            dwacc[o][j] = 0.f;
        }
    for (unsigned p0 = warp * PPW; p0 < npix; p0 += nwarps * PPW) {
        const unsigned pix = p0 + sub;
        if (pix >= npix) continue;
        const unsigned n = pix / HW, p = pix - n * HW;
        float dl[4] = {0.f, 0.f, 0.f, 0.f};
        for (int o = 0; o < n_out; ++o) {
            const size_t i = (static_cast<size_t>(n) * n_out + o) * HW + p;
            const float sg = out[i];
            dl[o] = dout[i] * sg * (1.f - sg);
        }
        float v[8];
        ld_split8_regs(x_hi, x_lo, static_cast<size_t>(pix) * x_ld + c, v);
        float g[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int o = 0; o < 4; ++o)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                g[j] = fmaf(dl[o], wv[o][j], g[j]);
                dwacc[o][j] = fmaf(dl[o], v[j], dwacc[o][j]);
            }
        float* d = dx + static_cast<size_t>(pix) * dx_ld + c;
        *reinterpret_cast<float4*>(d) = make_float4(g[0], g[1], g[2], g[3]);
        *reinterpret_cast<float4*>(d + 4) = make_float4(g[4], g[5], g[6], g[7]);
        if (l == 0)
            for (int o = 0; o < n_out; ++o) dbacc[o] += dl[o];
    }
    // deterministic combination: the 32 / L pixel sub-groups of a warp by shuffles, the 8 warps of the block in shared memory in
    // a fixed order, then one double atomic per (o, c) and per o
#pragma unroll
    for (int o = 0; o < 4; ++o) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int d = L; d < 32; d <<= 1) dwacc[o][j] += __shfl_xor_sync(0xffffffffu, dwacc[o][j], d);
#pragma unroll
        for (int d = L; d < 32; d <<= 1) dbacc[o] += __shfl_xor_sync(0xffffffffu, dbacc[o], d);
    }
    __shared__ float red[8][4][128];
    __shared__ float redb[8][4];
    const int wid = threadIdx.x >> 5;
    if (sub == 0) {
#pragma unroll
        for (int o = 0; o < 4; ++o)
#pragma unroll
            for (int j = 0; j < 8; ++j) red[wid][o][c + j] = dwacc[o][j];
        if (l == 0)
            for (int o = 0; o < 4; ++o) redb[wid][o] = dbacc[o];
    }
    __syncthreads();
    for (int t = threadIdx.x; t < n_out * Cin; t += blockDim.x) {
        const int o = t / Cin, cc = t % Cin;
        float sum = 0.f;
        for (int k = 0; k < 8; ++k) sum += red[k][o][cc];
        atomicAdd(dw + o * Cin + cc, static_cast<double>(sum));
    }
    if (threadIdx.x < n_out) {
        float sum = 0.f;
        for (int k = 0; k < 8; ++k) sum += redb[k][threadIdx.x];
        atomicAdd(db + threadIdx.x, static_cast<double>(sum));
    }
}

__global__ void double_to_float_kernel(const double* src, float* dst, int n, int accumulate) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = (accumulate ? dst[i] : 0.f) + static_cast<float>(src[i]);
}

// ---- discriminator head ----------------------------------------------------------------------------
// pooled[n][c] = mean_{h,w}(fx - fy); grid (N, C/8 groups / per-block), simple strided loop
__global__ void gap_diff_fwd_kernel(const __nv_bfloat16* x_hi, const __nv_bfloat16* x_lo, const __nv_bfloat16* y_hi,
                                    const __nv_bfloat16* y_lo, int ld, int HW, int C, float* pooled) {
    const int n = blockIdx.y;
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    const int py = threadIdx.x >> 5;  // 8 pixel lanes
    float acc = 0.f;
    if (c < C) {
        SplitCPtr X{x_hi, x_lo}, Y{y_hi, y_lo};
        for (int p = py; p < HW; p += 8) {
            const size_t off = (static_cast<size_t>(n) * HW + p) * ld + c;
            acc += load_split(X, off) - load_split(Y, off);
        }
    }
    __shared__ float red[8][32];
    red[py][threadIdx.x & 31] = acc;
    __syncthreads();
    if (py == 0 && c < C) {
        float s = 0.f;
        for (int k = 0; k < 8; ++k) s += red[k][threadIdx.x];
        pooled[n * C + c] = s / HW;
    }
}

// d_fx[n,p,c] = dpooled[n][c]/HW, d_fy = -d_fx  (fp32 NHWC)
__global__ void gap_diff_bwd_kernel(const float* __restrict__ dpooled, int HW, int C, long long total, float* dfx,
                                    float* dfy, int ld) {
    const long long idx = blockIdx.x * 1LL * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c = static_cast<int>(idx % C);
    const long long pix = idx / C;
    const long long n = pix / HW;
    const float g = dpooled[n * C + c] / HW;
    dfx[pix * ld + c] = g;
    dfy[pix * ld + c] = -g;
}

// dense layer: out[n][o] = act(b[o] + sum_i in[n][i] w[o][i]); one warp per (n, o).  act: 0 none, 3 leaky(0.2), 4 sigmoid
__global__ void fc_fwd_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ b, int N,
                              int I, int O, int act, float* __restrict__ pre, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long warp = (blockIdx.x * 1LL * blockDim.x + threadIdx.x) >> 5;
    if (warp >= 1LL * N * O) return;
    const int n = static_cast<int>(warp / O), o = static_cast<int>(warp % O);
    float acc = 0.f;
    for (int i = lane; i < I; i += 32) acc = fmaf(in[n * I + i], w[static_cast<size_t>(o) * I + i], acc);
    acc = warp_sum(acc);
    if (lane == 0) {
        const float u = acc + b[o];
        if (pre) pre[n * O + o] = u;
        out[n * O + o] = act == 3 ? (u > 0.f ? u : 0.2f * u) : act == 4 ? sigmoidf_(u) : u;
    }
}

// du[n][o] = dout[n][o] * act'(.)   (for sigmoid uses out; for leaky uses pre)
__global__ void fc_bwd_act_kernel(const float* dout, const float* pre, const float* out, int total, int act, float* du) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    float g = dout[i];
    if (act == 3) g *= (pre[i] > 0.f ? 1.f : 0.2f);
    else if (act == 4) g *= out[i] * (1.f - out[i]);
    du[i] = g;
}
// din[n][i] = sum_o du[n][o] w[o][i]
// (blockIdx.y splits the O loop into chunks of FC_OCHUNK so that the 16 x 512 outputs of the discriminator head keep more than
// 32 blocks busy; partial sums meet in `din`, which the wrapper zeroes first)
constexpr int FC_OCHUNK = 64;
__global__ void fc_bwd_input_kernel(const float* __restrict__ du, const float* __restrict__ w, int N, int I, int O,
                                    float* __restrict__ din) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * I) return;
    const int n = idx / I, i = idx % I;
    const int o0 = blockIdx.y * FC_OCHUNK, o1 = o0 + FC_OCHUNK < O ? o0 + FC_OCHUNK : O;
    float acc = 0.f;
    for (int o = o0; o < o1; ++o) acc = fmaf(du[n * O + o], w[static_cast<size_t>(o) * I + i], acc);
    atomicAdd(din + idx, acc);
}
// dw[o][i] (+)= sum_n du[n][o] in[n][i];  db[o] (+)= sum_n du[n][o]
__global__ void fc_bwd_weight_kernel(const float* __restrict__ du, const float* __restrict__ in, int N, int I, int O,
                                     float* __restrict__ dw, float* __restrict__ db, int accumulate) {
    const long long idx = blockIdx.x * 1LL * blockDim.x + threadIdx.x;
    if (idx >= 1LL * O * I) return;
    const int o = static_cast<int>(idx / I), i = static_cast<int>(idx % I);
    float acc = 0.f;
    for (int n = 0; n < N; ++n) acc = fmaf(du[n * O + o], in[n * I + i], acc);
    dw[idx] = (accumulate ? dw[idx] : 0.f) + acc;
    if (i == 0) {
        float s = 0.f;
        for (int n = 0; n < N; ++n) s += du[n * O + o];
        db[o] = (accumulate ? db[o] : 0.f) + s;
    }
}

// ---- masking (NCHW fp32, boundary level) -------------------------------------------------------------
// out = (a*(1-r) + b*r) * (1 - m)     with r = region (optional, b optional), m = mask (N,1,H,W)
__global__ void mask_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ region,
                                const float* __restrict__ m, int C, long long HW, long long total, float* __restrict__ out) {
    const long long idx = blockIdx.x * 1LL * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const long long p = idx % HW;
    const long long n = idx / (HW * C);
    float v = a[idx];
    if (region) {
        const float r = region[n * HW + p];
        v = v * (1.f - r) + b[idx] * r;
    }
    out[idx] = v * (1.f - m[n * HW + p]);
}
// The same on 4 consecutive pixels of one (n, c) plane per thread (float4 accesses, no 64-bit divisions): blockIdx.y = n*C + c.
// Needs HW % 4 == 0 and 16-byte aligned tensors; the arithmetic per element is identical to mask_fwd_kernel's.
__global__ void mask_fwd_vec4_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ region,
                                     const float* __restrict__ m, int C, long long HW, float* __restrict__ out) {
    const long long p = (blockIdx.x * 1LL * blockDim.x + threadIdx.x) * 4;
    if (p >= HW) return;
    const int plane = blockIdx.y, n = plane / C;
    const long long idx = plane * HW + p, pm = n * HW + p;
    float4 v = *reinterpret_cast<const float4*>(a + idx);
    if (region) {
        const float4 r = *reinterpret_cast<const float4*>(region + pm);
        const float4 w = *reinterpret_cast<const float4*>(b + idx);
        v.x = v.x * (1.f - r.x) + w.x * r.x;
        v.y = v.y * (1.f - r.y) + w.y * r.y;
        v.z = v.z * (1.f - r.z) + w.z * r.z;
        v.w = v.w * (1.f - r.w) + w.w * r.w;
    }
    const float4 k = *reinterpret_cast<const float4*>(m + pm);
    *reinterpret_cast<float4*>(out + idx) = make_float4(v.x * (1.f - k.x), v.y * (1.f - k.y), v.z * (1.f - k.z), v.w * (1.f - k.w));
}
// dm[n,p] (+)= -sum_c dout * (a*(1-r)+b*r);   one thread per pixel
__global__ void mask_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ a, const float* __restrict__ b,
                                const float* __restrict__ region, int C, long long HW, long long npix,
                                float* __restrict__ dm, int accumulate) {
    const long long pix = blockIdx.x * 1LL * blockDim.x + threadIdx.x;
    if (pix >= npix) return;
    const long long n = pix / HW, p = pix - n * HW;
    const float r = region ? region[pix] : 0.f;
    float acc = 0.f;
    for (int c = 0; c < C; ++c) {
        const long long i = (n * C + c) * HW + p;
        float v = a[i];
        if (region) v = v * (1.f - r) + b[i] * r;
        acc = fmaf(dout[i], v, acc);
    }
    dm[pix] = (accumulate ? dm[pix] : 0.f) - acc;
}

inline unsigned blocks_for(long long n) { return static_cast<unsigned>((n + NT - 1) / NT); }

}  // namespace
}  // namespace fcd

using namespace fcd;
#define CBF(p) reinterpret_cast<const __nv_bfloat16*>(p)

extern "C" {

int fcd_outconv_sigmoid_fwd(const void* x_hi, const void* x_lo, int x_ld, int Cin, const float* w, const float* b,
                            int n_out, int N, int H, int W, float* out_nchw, void* stream) {
    FCD_CHECK_ARG(x_hi && w && b && out_nchw, "fcd_outconv_sigmoid_fwd: null pointer");
    FCD_CHECK_ARG(n_out >= 1 && n_out <= 4 && Cin % 4 == 0 && x_ld % 4 == 0, "fcd_outconv_sigmoid_fwd: n_out in 1..4, Cin %% 4");
    const long long npix = 1LL * N * H * W;
    const bool wide = (Cin == 32 || Cin == 64 || Cin == 128) && x_ld % 8 == 0 && npix < (1LL << 31) &&
                      (reinterpret_cast<uintptr_t>(x_hi) & 15) == 0 && (!x_lo || (reinterpret_cast<uintptr_t>(x_lo) & 15) == 0);
    if (wide) {
        const int L = Cin / 8, ppb = (NT / 32) * (32 / L);       // pixels per block pass
        long long wb = (npix + ppb - 1) / ppb;
        if (wb > 16LL * sm_count()) wb = 16LL * sm_count();
        const unsigned g = static_cast<unsigned>(wb), np = static_cast<unsigned>(npix), hw = static_cast<unsigned>(1LL * H * W);
        cudaStream_t st = as_stream(stream);
        if (L == 4) outconv_fwd_wide_kernel<4><<<g, NT, 0, st>>>(CBF(x_hi), CBF(x_lo), x_ld, Cin, w, b, n_out, np, hw, out_nchw);
        else if (L == 8) outconv_fwd_wide_kernel<8><<<g, NT, 0, st>>>(CBF(x_hi), CBF(x_lo), x_ld, Cin, w, b, n_out, np, hw, out_nchw);
        else outconv_fwd_wide_kernel<16><<<g, NT, 0, st>>>(CBF(x_hi), CBF(x_lo), x_ld, Cin, w, b, n_out, np, hw, out_nchw);
        FCD_LAUNCH_OK();
        return FCD_OK;
    }
    long long blocks = (npix + 7) / 8;
    if (blocks > 16LL * sm_count()) blocks = 16LL * sm_count();
    outconv_fwd_kernel<<<static_cast<unsigned>(blocks), NT, 0, as_stream(stream)>>>(CBF(x_hi), CBF(x_lo), x_ld, Cin, w, b,
                                                                                    n_out, npix, 1LL * H * W, out_nchw);
    FCD_LAUNCH_OK();
    return FCD_OK;
}

// scratch: double[n_out*Cin + n_out], zeroed here.
int fcd_outconv_sigmoid_bwd(const float* dout_nchw, const float* out_nchw, const void* x_hi, const void* x_lo, int x_ld,
                            int Cin, const float* w, int n_out, int N, int H, int W, float* dx, int dx_ld, float* dw,
                            float* db, int accumulate, double* scratch, void* stream) {
    FCD_CHECK_ARG(dout_nchw && out_nchw && x_hi && w && dx && dw && db && scratch, "fcd_outconv_sigmoid_bwd: null pointer");
    FCD_CHECK_ARG(n_out >= 1 && n_out <= 4 && Cin % 4 == 0 && Cin <= 128 && dx_ld % 4 == 0,
                  "fcd_outconv_sigmoid_bwd: n_out in 1..4, Cin <= 128");
    cudaStream_t s = as_stream(stream);
    FCD_CUDA_OK(cudaMemsetAsync(scratch, 0, sizeof(double) * (n_out * Cin + n_out), s));
    const long long npix = 1LL * N * H * W;
    const bool wide = (Cin == 32 || Cin == 64 || Cin == 128) && x_ld % 8 == 0 && npix < (1LL << 31) &&
                      (reinterpret_cast<uintptr_t>(x_hi) & 15) == 0 && (!x_lo || (reinterpret_cast<uintptr_t>(x_lo) & 15) == 0) &&
                      (reinterpret_cast<uintptr_t>(dx) & 15) == 0;
    if (wide) {
        const int L = Cin / 8, ppb = (NT / 32) * (32 / L);
        long long wb = (npix + 8LL * ppb - 1) / (8LL * ppb);      // >= 8 pixels per thread before the block-level reduction
        if (wb > 8LL * sm_count()) wb = 8LL * sm_count();
        if (wb < 1) wb = 1;
        const unsigned g = static_cast<unsigned>(wb), np = static_cast<unsigned>(npix), hw = static_cast<unsigned>(1LL * H * W);
        double* sdb = scratch + n_out * Cin;
        if (L == 4) outconv_bwd_wide_kernel<4><<<g, NT, 0, s>>>(dout_nchw, out_nchw, CBF(x_hi), CBF(x_lo), x_ld, Cin, w, n_out, np, hw, dx, dx_ld, scratch, sdb);
        else if (L == 8) outconv_bwd_wide_kernel<8><<<g, NT, 0, s>>>(dout_nchw, out_nchw, CBF(x_hi), CBF(x_lo), x_ld, Cin, w, n_out, np, hw, dx, dx_ld, scratch, sdb);
        else outconv_bwd_wide_kernel<16><<<g, NT, 0, s>>>(dout_nchw, out_nchw, CBF(x_hi), CBF(x_lo), x_ld, Cin, w, n_out, np, hw, dx, dx_ld, scratch, sdb);
    } else {
        long long blocks = (npix + 63) / 64;
        if (blocks > 4LL * sm_count()) blocks = 4LL * sm_count();
        outconv_bwd_kernel<<<static_cast<unsigned>(blocks), NT, 0, s>>>(dout_nchw, out_nchw, CBF(x_hi), CBF(x_lo), x_ld, Cin, w,
                                                                        n_out, npix, 1LL * H * W, dx, dx_ld, scratch,
                                                                        scratch + n_out * Cin);
    }
    FCD_LAUNCH_OK();
    double_to_float_kernel<<<(n_out * Cin + 127) / 128, 128, 0, s>>>(scratch, dw, n_out * Cin, accumulate);
    double_to_float_kernel<<<1, 32, 0, s>>>(scratch + n_out * Cin, db, n_out, accumulate);
    FCD_LAUNCH_OK();
    return FCD_OK;
}

int fcd_gap_diff_fwd(const void* x_hi, const void* x_lo, const void* y_hi, const void* y_lo, int ld, int N, int HW, int C,
                     float* pooled, void* stream) {
    FCD_CHECK_ARG(x_hi && y_hi && pooled, "fcd_gap_diff_fwd: null pointer");
    gap_diff_fwd_kernel<<<dim3((C + 31) / 32, N), NT, 0, as_stream(stream)>>>(CBF(x_hi), CBF(x_lo), CBF(y_hi), CBF(y_lo), ld,
                                                                              HW, C, pooled);
    FCD_LAUNCH_OK();
    return FCD_OK;
}

int fcd_gap_diff_bwd(const float* dpooled, int N, int HW, int C, float* dfx, float* dfy, int ld, void* stream) {
    FCD_CHECK_ARG(dpooled && dfx && dfy, "fcd_gap_diff_bwd: null pointer");
    const long long total = 1LL * N * HW * C;
    gap_diff_bwd_kernel<<<blocks_for(total), NT, 0, as_stream(stream)>>>(dpooled, HW, C, total, dfx, dfy, ld);
    FCD_LAUNCH_OK();
    return FCD_OK;
}

int fcd_fc_fwd(const float* in, const float* w, const float* b, int N, int I, int O, int act, float* pre, float* out,
               void* stream) {
    FCD_CHECK_ARG(in && w && b && out, "fcd_fc_fwd: null pointer");
    fc_fwd_kernel<<<blocks_for(32LL * N * O), NT, 0, as_stream(stream)>>>(in, w, b, N, I, O, act, pre, out);
    FCD_LAUNCH_OK();
    return FCD_OK;
}

// du is caller scratch of N*O floats
int fcd_fc_bwd(const float* dout, const float* pre, const float* out, const float* in, const float* w, int N, int I, int O,
               int act, float* du, float* din, float* dw, float* db, int accumulate, void* stream) {
    FCD_CHECK_ARG(dout && in && w && du && dw && db, "fcd_fc_bwd: null pointer");
    cudaStream_t s = as_stream(stream);
    fc_bwd_act_kernel<<<(N * O + NT - 1) / NT, NT, 0, s>>>(dout, pre, out, N * O, act, du);
    if (din) {
        FCD_CUDA_OK(cudaMemsetAsync(din, 0, sizeof(float) * N * I, s));
        fc_bwd_input_kernel<<<dim3((N * I + NT - 1) / NT, (O + FC_OCHUNK - 1) / FC_OCHUNK), NT, 0, s>>>(du, w, N, I, O, din);
    }
    fc_bwd_weight_kernel<<<blocks_for(1LL * O * I), NT, 0, s>>>(du, in, N, I, O, dw, db, accumulate);
    FCD_LAUNCH_OK();
    return FCD_OK;
}

int fcd_mask_fwd(const float* a, const float* b, const float* region, const float* mask, int N, int C, int H, int W,
                 float* out, void* stream) {
    FCD_CHECK_ARG(a && mask && out && ((region == nullptr) == (b == nullptr)), "fcd_mask_fwd: bad arguments");
    const long long total = 1LL * N * C * H * W, HW = 1LL * H * W;
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    if (HW % 4 == 0 && 1LL * N * C <= 65535 && al16(a) && al16(mask) && al16(out) && al16(b) && al16(region)) {
        mask_fwd_vec4_kernel<<<dim3(blocks_for(HW / 4), N * C), NT, 0, as_stream(stream)>>>(a, b, region, mask, C, HW, out);
        FCD_LAUNCH_OK();
        return FCD_OK;
    }
    mask_fwd_kernel<<<blocks_for(total), NT, 0, as_stream(stream)>>>(a, b, region, mask, C, HW, total, out);
    FCD_LAUNCH_OK();
    return FCD_OK;
}

int fcd_mask_bwd(const float* dout, const float* a, const float* b, const float* region, int N, int C, int H, int W,
                 float* dmask, int accumulate, void* stream) {
    FCD_CHECK_ARG(dout && a && dmask, "fcd_mask_bwd: null pointer");
    const long long npix = 1LL * N * H * W;
    mask_bwd_kernel<<<blocks_for(npix), NT, 0, as_stream(stream)>>>(dout, a, b, region, C, 1LL * H * W, npix, dmask,
                                                                    accumulate);
    FCD_LAUNCH_OK();
    return FCD_OK;
}

}  // extern "C"
