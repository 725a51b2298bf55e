// conv_simt.cu — fp32 CUDA-core implicit-GEMM convolution kernels (any filter size / stride / channel
// count).  They serve the shapes the tcgen05 engine does not take (13-band heads, 1x1 heads, stride-2
// discriminator layers) and are the on-device cross-check for it.  Same operand conventions as
// conv_tc.cu: split-bf16 NHWC inputs, packed [tap][rows][cols] weights, fp32 NHWC outputs.
//
// Replaces nn.Conv2d forward/backward at Module.py:26,29,146,155,158,177,180,196-207.
#include "fcd_common.cuh"

namespace fcd {
int channel_sum_split(const void* hi, const void* lo, int ld, long long npix, int C, float* out, int accumulate,
                      cudaStream_t stream);
namespace {

constexpr int TM = 64;   // pixels per block tile
constexpr int TN = 64;   // output columns per block tile
constexpr int TK = 16;   // reduction chunk
constexpr int NT = 256;  // threads

__device__ __forceinline__ void load4_split(const __nv_bfloat16* hi, const __nv_bfloat16* lo, size_t off, float (&v)[4]) {
    // 4 consecutive bf16 (8 bytes) from each plane
    uint2 h = *reinterpret_cast<const uint2*>(hi + off);
    const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&h);
    float2 a = __bfloat1622float2(hp[0]), b = __bfloat1622float2(hp[1]);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    if (lo) {
        uint2 l = *reinterpret_cast<const uint2*>(lo + off);
        const __nv_bfloat162* lp = reinterpret_cast<const __nv_bfloat162*>(&l);
        float2 c = __bfloat1622float2(lp[0]), d = __bfloat1622float2(lp[1]);
        v[0] += c.x; v[1] += c.y; v[2] += d.x; v[3] += d.y;
    }
}

struct ConvDims {
    int N, H, W;       // input tensor spatial dims (for dgrad-strided: the dz dims)
    int OH, OW;        // output tensor spatial dims
    int Cin_p, Cout_p; // K-side channels and output-column channels of THIS gemm
    int KH, KW, stride, pad;
};

// MODE 0: forward conv (gather input pixel = out*stride + tap - pad)
// MODE 1: strided dgrad (gather dz pixel = (out + pad - tap)/stride when divisible); weights are the forward
//         packing [tap][Cout_p(K side)][Cin_p(columns)], i.e. B is read k-major-by-row.
template <int MODE>
__global__ void __launch_bounds__(NT)
conv_simt_kernel(SplitCPtr x, int x_ld, SplitCPtr w, const float* __restrict__ bias, const float* __restrict__ addend,
                 int addend_ld, float* __restrict__ z, int z_ld, ConvDims d, double* stat_sum, double* stat_sqsum) {
    __shared__ float As[TK][TM + 4];
    __shared__ float Bs[TK][TN + 4];

    const int tid = threadIdx.x;
    const int tiles_w = (d.OW + 7) / 8, tiles_h = (d.OH + 7) / 8;
    int bt = blockIdx.x;
    const int tw = bt % tiles_w; bt /= tiles_w;
    const int th = bt % tiles_h;
    const int n = bt / tiles_h;
    const int col0 = blockIdx.y * TN;

    // loader mapping: row = tid / 4 (0..63), k-quad = tid % 4
    const int lrow = tid >> 2, lk = (tid & 3) * 4;
    const int a_oh = th * 8 + (lrow >> 3), a_ow = tw * 8 + (lrow & 7);
    // compute mapping: 4 pixels x 4 columns
    const int ty = tid >> 4, tx = tid & 15;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int kchunks = d.Cin_p / TK;
    for (int tap = 0; tap < d.KH * d.KW; ++tap) {
        const int r = tap / d.KW, s = tap - r * d.KW;
        int ih, iw;
        bool ok;
        if (MODE == 0) {
            ih = a_oh * d.stride + r - d.pad;
            iw = a_ow * d.stride + s - d.pad;
            ok = (a_oh < d.OH) && (a_ow < d.OW) && ih >= 0 && ih < d.H && iw >= 0 && iw < d.W;
        } else {
            const int nh = a_oh + d.pad - r, nw = a_ow + d.pad - s;
            ih = nh / d.stride;
            iw = nw / d.stride;
            ok = (a_oh < d.OH) && (a_ow < d.OW) && nh >= 0 && nw >= 0 && (nh % d.stride == 0) &&
                 (nw % d.stride == 0) && ih < d.H && iw < d.W;
        }
        const size_t a_base = ((static_cast<size_t>(n) * d.H + ih) * d.W + iw) * x_ld;
        for (int kc = 0; kc < kchunks; ++kc) {
            float av[4] = {0.f, 0.f, 0.f, 0.f};
            if (ok) load4_split(x.hi, x.lo, a_base + kc * TK + lk, av);
            float bv[4] = {0.f, 0.f, 0.f, 0.f};
            if (MODE == 0) {
                // B[k][col] = w[tap][col0+col][kc*TK + k]  (rows = output channels, cols = input channels)
                const int co = col0 + lrow;
                if (co < d.Cout_p)
                    load4_split(w.hi, w.lo, (static_cast<size_t>(tap) * d.Cout_p + co) * d.Cin_p + kc * TK + lk, bv);
#pragma unroll
                for (int j = 0; j < 4; ++j) Bs[lk + j][lrow] = bv[j];
            } else {
                // B[k][col] = w[tap][kc*TK + k][col0+col]  (rows = K side, cols contiguous)
                const int krow = tid >> 4, cq = (tid & 15) * 4;
                if (col0 + cq < d.Cout_p)
                    load4_split(w.hi, w.lo, (static_cast<size_t>(tap) * d.Cin_p + kc * TK + krow) * d.Cout_p + col0 + cq, bv);
#pragma unroll
                for (int j = 0; j < 4; ++j) Bs[krow][cq + j] = bv[j];
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) As[lk + j][lrow] = av[j];
            __syncthreads();
#pragma unroll
            for (int k = 0; k < TK; ++k) {
                const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
                const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
                const float a[4] = {a4.x, a4.y, a4.z, a4.w};
                const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
            __syncthreads();
        }
    }

    const int c = col0 + tx * 4;
    float ssum[4] = {0.f, 0.f, 0.f, 0.f}, ssq[4] = {0.f, 0.f, 0.f, 0.f};
    if (c < d.Cout_p) {
        float b4[4] = {0.f, 0.f, 0.f, 0.f};
        if (bias) {
#pragma unroll
            for (int j = 0; j < 4; ++j) b4[j] = bias[c + j];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int prow = ty * 4 + i;
            const int oh = th * 8 + (prow >> 3), ow = tw * 8 + (prow & 7);
            if (oh < d.OH && ow < d.OW) {
                float4 o = make_float4(acc[i][0] + b4[0], acc[i][1] + b4[1], acc[i][2] + b4[2], acc[i][3] + b4[3]);
                const size_t pix = (static_cast<size_t>(n) * d.OH + oh) * d.OW + ow;
                if (addend) {
                    const float4 a = *reinterpret_cast<const float4*>(addend + pix * addend_ld + c);
                    o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
                }
                *reinterpret_cast<float4*>(z + pix * z_ld + c) = o;
                ssum[0] += o.x; ssum[1] += o.y; ssum[2] += o.z; ssum[3] += o.w;
                ssq[0] += o.x * o.x; ssq[1] += o.y * o.y; ssq[2] += o.z * o.z; ssq[3] += o.w * o.w;
            }
        }
    }
    if (stat_sum) {
        // reduce the 16 pixel-groups (ty) that share a column quad through shared memory
        __syncthreads();
        float* red = &As[0][0];  // 16 * 64 * 2 floats needed; As+Bs are contiguous enough? use As only: 16*68 = 1088
        // layout red[ty][tx*4+j] for sums, then reuse Bs for squares
        float* red2 = &Bs[0][0];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            red[ty * TN + tx * 4 + j] = ssum[j];
            red2[ty * TN + tx * 4 + j] = ssq[j];
        }
        __syncthreads();
        if (tid < TN && col0 + tid < d.Cout_p) {
            float a = 0.f, b = 0.f;
#pragma unroll
            for (int y = 0; y < 16; ++y) {
                a += red[y * TN + tid];
                b += red2[y * TN + tid];
            }
            atomicAdd(stat_sum + col0 + tid, static_cast<double>(a));
            atomicAdd(stat_sqsum + col0 + tid, static_cast<double>(b));
        }
    }
}

// wgrad: dw[co][ci][r][s] += sum_pix dz[pix][co] * x[pix*stride + tap - pad][ci]
// grid: x = taps * co_tiles * ci_tiles, y = split over pixel chunks
__global__ void __launch_bounds__(NT)
wgrad_simt_kernel(SplitCPtr x, int x_ld, SplitCPtr dz, int dz_ld, float* __restrict__ dw, ConvDims d, int Cin,
                  int Cout, int chunks_per_block) {
    __shared__ float As[TK][TM + 4];  // dz  : [pixel k][co]
    __shared__ float Bs[TK][TN + 4];  // x   : [pixel k][ci]
    const int tid = threadIdx.x;
    const int co_tiles = (d.Cout_p + TM - 1) / TM, ci_tiles = (d.Cin_p + TN - 1) / TN;
    int b = blockIdx.x;
    const int cit = b % ci_tiles; b /= ci_tiles;
    const int cot = b % co_tiles;
    const int tap = b / co_tiles;
    const int r = tap / d.KW, s = tap - r * d.KW;
    const int co0 = cot * TM, ci0 = cit * TN;

    const long long npix = 1LL * d.N * d.OH * d.OW;
    const long long pix_begin = 1LL * blockIdx.y * chunks_per_block * TK;
    long long pix_end = pix_begin + 1LL * chunks_per_block * TK;
    if (pix_end > npix) pix_end = npix;

    const int krow = tid >> 4, cq = (tid & 15) * 4;  // loader: pixel row within chunk, column quad
    const int ty = tid >> 4, tx = tid & 15;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (long long p0 = pix_begin; p0 < pix_end; p0 += TK) {
        const long long pix = p0 + krow;
        float av[4] = {0.f, 0.f, 0.f, 0.f}, bv[4] = {0.f, 0.f, 0.f, 0.f};
        if (pix < pix_end) {
            const int ow = static_cast<int>(pix % d.OW);
            const long long t = pix / d.OW;
            const int oh = static_cast<int>(t % d.OH);
            const int n = static_cast<int>(t / d.OH);
            if (co0 + cq < d.Cout_p) load4_split(dz.hi, dz.lo, static_cast<size_t>(pix) * dz_ld + co0 + cq, av);
            const int ih = oh * d.stride + r - d.pad, iw = ow * d.stride + s - d.pad;
            if (ih >= 0 && ih < d.H && iw >= 0 && iw < d.W && ci0 + cq < d.Cin_p)
                load4_split(x.hi, x.lo, ((static_cast<size_t>(n) * d.H + ih) * d.W + iw) * x_ld + ci0 + cq, bv);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            As[krow][cq + j] = av[j];
            Bs[krow][cq + j] = bv[j];
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < TK; ++k) {
            const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float a[4] = {a4.x, a4.y, a4.z, a4.w};
            const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int co = co0 + ty * 4 + i;
        if (co >= Cout) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int ci = ci0 + tx * 4 + j;
            if (ci < Cin) atomicAdd(dw + ((static_cast<size_t>(co) * Cin + ci) * d.KH + r) * d.KW + s, acc[i][j]);
        }
    }
}

__global__ void pack_weight_kernel(const float* __restrict__ w, int Cout, int Cin, int KH, int KW, int Cout_p,
                                   int Cin_p, int mode, __nv_bfloat16* hi, __nv_bfloat16* lo) {
    const long long total = 1LL * KH * KW * Cout_p * Cin_p;
    for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
        long long t = i;
        int co, ci;
        if (mode == 0) {
            ci = static_cast<int>(t % Cin_p); t /= Cin_p;
            co = static_cast<int>(t % Cout_p); t /= Cout_p;
        } else {
            co = static_cast<int>(t % Cout_p); t /= Cout_p;
            ci = static_cast<int>(t % Cin_p); t /= Cin_p;
        }
        int tap = static_cast<int>(t);
        int r = tap / KW, s = tap % KW;
        if (mode == 1) {
            r = KH - 1 - r;
            s = KW - 1 - s;
        }
        float v = 0.f;
        if (co < Cout && ci < Cin) v = w[((static_cast<size_t>(co) * Cin + ci) * KH + r) * KW + s];
        __nv_bfloat16 h = __float2bfloat16_rn(v);
        hi[i] = h;
        if (lo) lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
}

}  // namespace

int conv2d_fwd_simt(const void* x_hi, const void* x_lo, int x_ld, const void* w_hi, const void* w_lo,
                    const float* bias, const float* addend, int addend_ld, float* z, int z_ld, int N, int H, int W,
                    int Cin_p, int Cout_p, int KH, int KW, int stride, int pad, double* stat_sum, double* stat_sqsum,
                    cudaStream_t stream) {
    ConvDims d;
    d.N = N; d.H = H; d.W = W;
    d.OH = (H + 2 * pad - KH) / stride + 1;
    d.OW = (W + 2 * pad - KW) / stride + 1;
    d.Cin_p = Cin_p; d.Cout_p = Cout_p; d.KH = KH; d.KW = KW; d.stride = stride; d.pad = pad;
    FCD_CHECK_ARG(d.OH > 0 && d.OW > 0, "conv2d_fwd_simt: empty output");
    FCD_CHECK_ARG(Cin_p % 16 == 0 && Cout_p % 4 == 0 && z_ld % 4 == 0 && x_ld % 4 == 0,
                  "conv2d_fwd_simt: Cin_p %% 16, Cout_p %% 4, pitches %% 4 required (Cin_p=%d Cout_p=%d)", Cin_p, Cout_p);
    dim3 grid(N * ((d.OH + 7) / 8) * ((d.OW + 7) / 8), (Cout_p + TN - 1) / TN);
    SplitCPtr x{(const __nv_bfloat16*)x_hi, (const __nv_bfloat16*)x_lo};
    SplitCPtr w{(const __nv_bfloat16*)w_hi, (const __nv_bfloat16*)w_lo};
    conv_simt_kernel<0><<<grid, NT, 0, stream>>>(x, x_ld, w, bias, addend, addend_ld, z, z_ld, d, stat_sum, stat_sqsum);
    FCD_LAUNCH_OK();
    return FCD_OK;
}

int conv2d_dgrad_strided_simt(const void* dz_hi, const void* dz_lo, int dz_ld, const void* w_hi, const void* w_lo,
                              const float* addend, int addend_ld, float* dx, int dx_ld, int N, int H, int W, int Cin_p, int Cout_p, int KH, int KW,
                              int stride, int pad, cudaStream_t stream) {
    // H, W are the INPUT (dx) dims; dz has dims OH x OW.
    ConvDims d;
    d.N = N;
    d.H = (H + 2 * pad - KH) / stride + 1;  // gathered tensor = dz
    d.W = (W + 2 * pad - KW) / stride + 1;
    d.OH = H; d.OW = W;                     // produced tensor = dx
    d.Cin_p = Cout_p;                       // K side = output channels of the forward conv
    d.Cout_p = Cin_p;                       // columns = input channels of the forward conv
    d.KH = KH; d.KW = KW; d.stride = stride; d.pad = pad;
    FCD_CHECK_ARG(Cout_p % 16 == 0 && Cin_p % 4 == 0 && dx_ld % 4 == 0 && dz_ld % 4 == 0,
                  "conv2d_dgrad_strided: channel padding requirements violated");
    dim3 grid(N * ((H + 7) / 8) * ((W + 7) / 8), (Cin_p + TN - 1) / TN);
    SplitCPtr g{(const __nv_bfloat16*)dz_hi, (const __nv_bfloat16*)dz_lo};
    SplitCPtr w{(const __nv_bfloat16*)w_hi, (const __nv_bfloat16*)w_lo};
    conv_simt_kernel<1><<<grid, NT, 0, stream>>>(g, dz_ld, w, nullptr, addend, addend_ld, dx, dx_ld, d, nullptr, nullptr);
    FCD_LAUNCH_OK();
    return FCD_OK;
}

int conv2d_wgrad_simt(const void* x_hi, const void* x_lo, int x_ld, const void* dz_hi, const void* dz_lo, int dz_ld,
                      float* dw, float* db, int N, int H, int W, int Cin, int Cin_p, int Cout, int Cout_p, int KH,
                      int KW, int stride, int pad, int accumulate, cudaStream_t stream) {
    ConvDims d;
    d.N = N; d.H = H; d.W = W;
    d.OH = (H + 2 * pad - KH) / stride + 1;
    d.OW = (W + 2 * pad - KW) / stride + 1;
    d.Cin_p = Cin_p; d.Cout_p = Cout_p; d.KH = KH; d.KW = KW; d.stride = stride; d.pad = pad;
    FCD_CHECK_ARG(Cin_p % 4 == 0 && Cout_p % 4 == 0 && x_ld % 4 == 0 && dz_ld % 4 == 0, "conv2d_wgrad_simt: padding");
    if (!accumulate)
        FCD_CUDA_OK(cudaMemsetAsync(dw, 0, sizeof(float) * static_cast<size_t>(Cout) * Cin * KH * KW, stream));
    const long long npix = 1LL * N * d.OH * d.OW;
    const int co_tiles = (Cout_p + TM - 1) / TM, ci_tiles = (Cin_p + TN - 1) / TN;
    const int gx = KH * KW * co_tiles * ci_tiles;
    const long long chunks = (npix + TK - 1) / TK;
    // aim for ~8 blocks per SM overall
    long long want_y = (8LL * sm_count() + gx - 1) / gx;
    if (want_y < 1) want_y = 1;
    if (want_y > chunks) want_y = chunks;
    const int chunks_per_block = static_cast<int>((chunks + want_y - 1) / want_y);
    const int gy = static_cast<int>((chunks + chunks_per_block - 1) / chunks_per_block);
    SplitCPtr x{(const __nv_bfloat16*)x_hi, (const __nv_bfloat16*)x_lo};
    SplitCPtr g{(const __nv_bfloat16*)dz_hi, (const __nv_bfloat16*)dz_lo};
    wgrad_simt_kernel<<<dim3(gx, gy), NT, 0, stream>>>(x, x_ld, g, dz_ld, dw, d, Cin, Cout, chunks_per_block);
    FCD_LAUNCH_OK();
    if (db) return channel_sum_split(dz_hi, dz_lo, dz_ld, npix, Cout, db, accumulate, stream);
    return FCD_OK;
}

int pack_conv_weight(const float* w, int Cout, int Cin, int KH, int KW, int Cout_p, int Cin_p, int mode, void* hi,
                     void* lo, cudaStream_t stream) {
    const long long total = 1LL * KH * KW * Cout_p * Cin_p;
    const int blocks = static_cast<int>((total + 255) / 256 > 4096 ? 4096 : (total + 255) / 256);
    pack_weight_kernel<<<blocks, 256, 0, stream>>>(w, Cout, Cin, KH, KW, Cout_p, Cin_p, mode, (__nv_bfloat16*)hi,
                                                   (__nv_bfloat16*)lo);
    FCD_LAUNCH_OK();
    return FCD_OK;
}

}  // namespace fcd
