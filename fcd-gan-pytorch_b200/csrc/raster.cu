// raster.cu — the data formats either side of the hot path: raster -> normalised tile batch,
// tile batch -> stitched change-density raster, thresholded tiles -> confusion matrix.  All HBM-bound byte/integer work:
// one coalesced pass each, no shared-memory staging needed (every element is touched once).
//
//   fcd_tiles_gather         GDALDataset.__getitem__ (data_utils.py:94-123) + NORMALIZE.forward (CommonFunc.py:208-224)
//   fcd_tiles_moments        Dataset_mean / Dataset_std per-tile reductions (CommonFunc.py:436-499)
//   fcd_tiles_scatter        GDALDataset.GDALwriteDefault (data_utils.py:178-213)
//   fcd_confusion_accumulate Demo_USSS.py:349-362 + Evaluator._generate_matrix_bymap (metrics.py:74-80)
#include "fcd_common.cuh"

namespace fcd {
namespace {

constexpr int NT = 256;

template <typename T>
__device__ __forceinline__ double load_as_double(const void* base, size_t i) {
    return static_cast<double>(reinterpret_cast<const T*>(base)[i]);
}

__device__ __forceinline__ double raster_value(const void* base, int dtype, size_t i) {
    switch (dtype) {
        case FCD_RASTER_U16: return load_as_double<uint16_t>(base, i);
        case FCD_RASTER_I16: return load_as_double<int16_t>(base, i);
        case FCD_RASTER_U8: return load_as_double<uint8_t>(base, i);
        default: return load_as_double<float>(base, i);
    }
}

// float((d - mean) / std) with the division done in float64 and ONE rounding to float32, exactly like the reference's numpy
// code, but without paying for a float64 division per element: q = (d - mean) * (1 / std) differs from the correctly rounded
// quotient by at most 2 ulp(float64), so float(q) can only differ from float(quotient) when q lies within a few float64 ulps
// of a float32 rounding boundary (a midpoint between two adjacent floats: low 29 mantissa bits == 0x10000000).  Those rare
// elements (about 1 in 2^26), and anything non-finite or in the float32 subnormal range, take the exact division.
__device__ __forceinline__ float normalise_exact(double d, double mean, double stdv, double inv) {
    const double num = d - mean;
    const double q = num * inv;
    const unsigned long long bits = static_cast<unsigned long long>(__double_as_longlong(q));
    const unsigned int low = static_cast<unsigned int>(bits) & 0x1FFFFFFFu;
    const unsigned int dist = low > 0x10000000u ? low - 0x10000000u : 0x10000000u - low;
    const double aq = fabs(q);
    if (dist <= 8u || !(aq > 1e-30 && aq < 1e30) || !(fabs(inv) < 1e300)) return static_cast<float>(num / stdv);
    return static_cast<float>(q);
}

// geom[t] = {read_x, read_y, read_w, read_h, write_x, write_y} for tile t of the grid; items[b] selects the tile of batch
// slot b (null = identity): the raster window [read_y, read_y+read_h) x [read_x, ..) lands at (write_y, write_x) of the
// zero-initialised patch (slice_read / slice_write of data_utils.py:151-176).
// grid = (quads of one plane, C, B): a thread produces 4 consecutive pixels of one row -> one 16-byte store.
template <int VEC>
__global__ void tiles_gather_kernel(const void* __restrict__ raster, int dtype, int C, int H, int W,
                                    const int* __restrict__ geom, const int* __restrict__ items, int pw, int ph,
                                    const double* __restrict__ mean, const double* __restrict__ stdv,
                                    float* __restrict__ out) {
    const int c = blockIdx.y, b = blockIdx.z;
    const int* g = geom + (items ? items[b] : b) * 6;
    const int rx = g[0], ry = g[1], rw = g[2], rh = g[3], wx = g[4], wy = g[5];
    const bool norm = mean != nullptr;
    const double m = norm ? mean[c] : 0.0, sd = norm ? stdv[c] : 1.0, inv = 1.0 / sd;
    const int qpr = pw / VEC;                              // quads per row
    const int nq = qpr * ph;
    float* ob = out + (static_cast<size_t>(b) * C + c) * ph * pw;
    const size_t plane = static_cast<size_t>(c) * H * W;
    for (int q = blockIdx.x * NT + threadIdx.x; q < nq; q += gridDim.x * NT) {
        const int py = q / qpr, px0 = (q - py * qpr) * VEC;
        const int dy = py - wy;
        float v[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) v[j] = 0.f;
        if (dy >= 0 && dy < rh) {
            const size_t row = plane + static_cast<size_t>(ry + dy) * W + rx;
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                const int dx = px0 + j - wx;
                if (dx >= 0 && dx < rw) {
                    const double d = raster_value(raster, dtype, row + dx);
                    v[j] = norm ? normalise_exact(d, m, sd, inv) : static_cast<float>(d);
                }
            }
        }
        if (VEC == 4) {
            *reinterpret_cast<float4*>(ob + static_cast<size_t>(py) * pw + px0) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
#pragma unroll
            for (int j = 0; j < VEC; ++j) ob[static_cast<size_t>(py) * pw + px0 + j] = v[j];
        }
    }
}

// Per tile b: valid = pixels whose channel sum of x is non-zero (fp32 sequential sum over bands, CommonFunc.py:446,479);
// pass 1 (centre == nullptr): sums[b][0][c] = sum_valid x[c], sums[b][1][c] = sum_valid y[c], counts[b] = #valid;
// pass 2: the same with (v - centre)^2, centre = [meanX | meanY].
__global__ void tiles_moments_kernel(const float* __restrict__ x, const float* __restrict__ y, int C, int npix,
                                     const double* __restrict__ centre, double* __restrict__ sums,
                                     long long* __restrict__ counts) {
    const int b = blockIdx.y;
    const float* xb = x + static_cast<size_t>(b) * C * npix;
    const float* yb = y + static_cast<size_t>(b) * C * npix;
    __shared__ double red[NT / 32];
    long long cnt = 0;
    // thread-private accumulators for up to 16 bands per sweep
    for (int c0 = 0; c0 < C; c0 += 8) {
        double ax[8], ay[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) ax[j] = ay[j] = 0.0;
        for (int p = blockIdx.x * NT + threadIdx.x; p < npix; p += gridDim.x * NT) {
            float s = 0.f;
            for (int c = 0; c < C; ++c) s += xb[static_cast<size_t>(c) * npix + p];
            if (s != 0.f) {
                if (c0 == 0) ++cnt;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (c0 + j < C) {
                        double vx = xb[static_cast<size_t>(c0 + j) * npix + p], vy = yb[static_cast<size_t>(c0 + j) * npix + p];
                        if (centre) {
                            vx -= centre[c0 + j];
                            vy -= centre[C + c0 + j];
                            vx *= vx;
                            vy *= vy;
                        }
                        ax[j] += vx;
                        ay[j] += vy;
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (c0 + j >= C) break;
            for (int which = 0; which < 2; ++which) {
                double v = warp_sum(which ? ay[j] : ax[j]);
                if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
                __syncthreads();
                if (threadIdx.x == 0) {
                    double t = 0.0;
                    for (int w = 0; w < NT / 32; ++w) t += red[w];
                    atomicAdd(sums + (static_cast<size_t>(b) * 2 + which) * C + c0 + j, t);
                }
                __syncthreads();
            }
        }
    }
    if (counts) {
        // block-level count, one atomic per block
        __shared__ long long cred[NT / 32];
        long long v = cnt;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) cred[threadIdx.x >> 5] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
            long long t = 0;
            for (int w = 0; w < NT / 32; ++w) t += cred[w];
            atomicAdd(reinterpret_cast<unsigned long long*>(counts + b), static_cast<unsigned long long>(t));
        }
    }
}

// geom[t] = {pad_x, pad_y, slice_x, slice_y, slice_w, slice_h} (tile t = items[b]): tile[pad_y : pad_y+slice_h, pad_x : pad_x+slice_w] is written
// at (slice_y, slice_x) of the full raster (data_utils.py:212-213).  Centre crops of different tiles never overlap.
__global__ void tiles_scatter_kernel(const float* __restrict__ tiles, const int* __restrict__ geom,
                                     const int* __restrict__ items, int pw, int ph, float* __restrict__ raster, int H, int W) {
    const int b = blockIdx.y;
    const int* g = geom + (items ? items[b] : b) * 6;
    const int sw = g[4], sh = g[5];
    const float* tb = tiles + static_cast<size_t>(b) * ph * pw;
    for (int i = blockIdx.x * NT + threadIdx.x; i < sw * sh; i += gridDim.x * NT) {
        const int dy = i / sw, dx = i - dy * sw;
        const int ty = g[1] + dy, tx = g[0] + dx, ry = g[3] + dy, rx = g[2] + dx;
        if (ty < ph && tx < pw && ry < H && rx < W) raster[static_cast<size_t>(ry) * W + rx] = tb[static_cast<size_t>(ty) * pw + tx];
    }
}

// counts[i*2 + j] += #{pixels in the tile's centre crop : int16(ref) == gt_map[i] and (cmap > thresh) == pre_map[j]}
__global__ void confusion_kernel(const float* __restrict__ cmap, const float* __restrict__ ref, const int* __restrict__ geom,
                                 const int* __restrict__ items, int pw, int ph, float thresh, int gt0, int gt1, int pre0,
                                 int pre1, unsigned long long* __restrict__ counts) {
    const int b = blockIdx.y;
    const int* g = geom + (items ? items[b] : b) * 6;
    const int sw = g[4], sh = g[5];
    const float* cb = cmap + static_cast<size_t>(b) * ph * pw;
    const float* rb = ref + static_cast<size_t>(b) * ph * pw;
    unsigned int c[4] = {0, 0, 0, 0};
    for (int i = blockIdx.x * NT + threadIdx.x; i < sw * sh; i += gridDim.x * NT) {
        const int dy = i / sw, dx = i - dy * sw;
        const int ty = g[1] + dy, tx = g[0] + dx;
        if (ty >= ph || tx >= pw) continue;
        const size_t k = static_cast<size_t>(ty) * pw + tx;
        const int r = static_cast<int>(static_cast<short>(rb[k]));        // .astype(np.int16): truncation toward zero
        const int m = cb[k] > thresh ? 1 : 0;                               // cmask[cmap > prob_thresh] = 1
        const int gi = r == gt0 ? 0 : (r == gt1 ? 1 : -1);
        const int pj = m == pre0 ? 0 : (m == pre1 ? 1 : -1);
        if (gi >= 0 && pj >= 0) ++c[gi * 2 + pj];
    }
    __shared__ unsigned int red[4][NT / 32];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        unsigned int v = c[q];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) red[q][threadIdx.x >> 5] = v;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        unsigned long long t = 0;
        for (int w = 0; w < NT / 32; ++w) t += red[threadIdx.x][w];
        if (t) atomicAdd(counts + threadIdx.x, t);
    }
}

}  // namespace
}  // namespace fcd

using namespace fcd;

extern "C" {

int fcd_tiles_gather(const void* raster, int dtype, int C, int H, int W, const int* geom, const int* items, int B,
                     int patch_w, int patch_h, const double* mean, const double* stdv, float* out, void* stream) {
    FCD_CHECK_ARG(raster && geom && out, "fcd_tiles_gather: null pointer");
    FCD_CHECK_ARG(dtype >= FCD_RASTER_F32 && dtype <= FCD_RASTER_U8, "fcd_tiles_gather: unknown raster dtype %d", dtype);
    FCD_CHECK_ARG(C > 0 && H > 0 && W > 0 && B > 0 && patch_w > 0 && patch_h > 0, "fcd_tiles_gather: bad dims");
    FCD_CHECK_ARG(C <= 65535 && B <= 65535, "fcd_tiles_gather: at most 65535 bands / tiles per launch");
    FCD_CHECK_ARG((mean == nullptr) == (stdv == nullptr), "fcd_tiles_gather: mean and std go together");
    const bool vec = patch_w % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    const int nq = (vec ? patch_w / 4 : patch_w) * patch_h;
    int gx = (nq + NT - 1) / NT;
    if (gx > 64) gx = 64;
    const dim3 grid(gx, C, B);
    if (vec)
        tiles_gather_kernel<4><<<grid, NT, 0, as_stream(stream)>>>(raster, dtype, C, H, W, geom, items, patch_w, patch_h, mean,
                                                                   stdv, out);
    else
        tiles_gather_kernel<1><<<grid, NT, 0, as_stream(stream)>>>(raster, dtype, C, H, W, geom, items, patch_w, patch_h, mean,
                                                                   stdv, out);
    FCD_LAUNCH_OK();
    return FCD_OK;
}

int fcd_tiles_moments(const float* x_tiles, const float* y_tiles, int B, int C, int npix_per_tile, const double* centre,
                      double* sums, long long* counts, void* stream) {
    FCD_CHECK_ARG(x_tiles && y_tiles && sums, "fcd_tiles_moments: null pointer");
    FCD_CHECK_ARG(B > 0 && C > 0 && npix_per_tile > 0, "fcd_tiles_moments: bad dims");
    int gx = (npix_per_tile + NT * 4 - 1) / (NT * 4);
    if (gx < 1) gx = 1;
    if (gx > 64) gx = 64;
    tiles_moments_kernel<<<dim3(gx, B), NT, 0, as_stream(stream)>>>(x_tiles, y_tiles, C, npix_per_tile, centre, sums, counts);
    FCD_LAUNCH_OK();
    return FCD_OK;
}

int fcd_tiles_scatter(const float* tiles, const int* geom, const int* items, int B, int patch_w, int patch_h, float* raster,
                      int H, int W, void* stream) {
    FCD_CHECK_ARG(tiles && geom && raster, "fcd_tiles_scatter: null pointer");
    FCD_CHECK_ARG(B > 0 && patch_w > 0 && patch_h > 0 && H > 0 && W > 0, "fcd_tiles_scatter: bad dims");
    int gx = (patch_w * patch_h + NT - 1) / NT;
    tiles_scatter_kernel<<<dim3(gx, B), NT, 0, as_stream(stream)>>>(tiles, geom, items, patch_w, patch_h, raster, H, W);
    FCD_LAUNCH_OK();
    return FCD_OK;
}

int fcd_confusion_accumulate(const float* cmap, const float* ref, const int* geom, const int* items, int B, int patch_w,
                             int patch_h, float thresh, int gt0, int gt1, int pre0, int pre1, long long* counts, void* stream) {
    FCD_CHECK_ARG(cmap && ref && geom && counts, "fcd_confusion_accumulate: null pointer");
    FCD_CHECK_ARG(B > 0 && patch_w > 0 && patch_h > 0, "fcd_confusion_accumulate: bad dims");
    int gx = (patch_w * patch_h + NT * 4 - 1) / (NT * 4);
    confusion_kernel<<<dim3(gx, B), NT, 0, as_stream(stream)>>>(cmap, ref, geom, items, patch_w, patch_h, thresh, gt0, gt1, pre0,
                                                                pre1, reinterpret_cast<unsigned long long*>(counts));
    FCD_LAUNCH_OK();
    return FCD_OK;
}

}  // extern "C"
