// wgrad_tc.cu — tcgen05 weight-gradient GEMM for stride-1 convolutions (nn.Conv2d backward w.r.t. weight,
// i.e. autograd of Module.py:26,29,155,177,180).
//
//   dW[tap][ci][co] = sum_{pixels} x[pixel + tap][ci] * dz[pixel][co]
//
// GEMM view: M = (tap, ci) "row groups" of 64 input channels, two per 128-row UMMA; N = co;
// K = output pixels.  Both operands are NHWC so the reduction dimension (pixels) is the strided one:
// the UMMA descriptors are MN-major, the TMA boxes {64 ch, 16 w, 4 h} land as [64 pixel rows][128 B].
// Each CTA owns one (M-block, N-block) accumulator in TMEM over a contiguous range of pixel tiles
// (split-K) and adds it into an fp32 workspace [tap][ci][co]; a small kernel then scatters the
// workspace into the OIHW gradient torch expects.
#include <cuda.h>

#include "fcd_common.cuh"
#include "fcd_tc.cuh"

namespace fcd {
using namespace tc;

int make_act_tmap(CUtensorMap* m, const void* base, int C, int W, int H, int N, int ld, int box_c, int box_w,
                  int box_h, int estride_w, int estride_h);

bool g_wgrad_halo_enabled = true;    // FCD_WGRAD_HALO=0 (read by fcd_set_option) selects the per-tap-pair kernel

namespace {

constexpr int PT_H = 4, PT_W = 16;        // pixel tile = 64 pixels = K per stage
constexpr int SUB_BYTES = 64 * 128;       // one [64 px][64 ch] bf16 sub-tile
constexpr int NUM_THREADS = 192;

template <int BLOCK_N, bool SPLIT>
struct WCfg {
    static constexpr int PLANES = SPLIT ? 2 : 1;
    static constexpr int NB = BLOCK_N / 64;
    static constexpr int A_BYTES = 2 * SUB_BYTES;            // per plane
    static constexpr int B_BYTES = NB * SUB_BYTES;           // per plane
    static constexpr int STAGE_BYTES = (A_BYTES + B_BYTES) * PLANES;
    static constexpr int STAGES = (192 * 1024) / STAGE_BYTES > 8 ? 8 : (192 * 1024) / STAGE_BYTES;
    static constexpr int TMEM_COLS = BLOCK_N;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
};

struct WgradParams {
    float* ws;            // [R*64][Cout_p] fp32, zero-initialised by the host wrapper
    int N, OH, OW;
    int Cin_p, Cout_p, KH, KW, stride;   // KH x KW = the tap grid iterated (n_r x n_s for the generic tap-list form)
    int h_off, w_off, s_step;           // x pixel of tap (r, s) for output pixel (h, w): (h*stride + r + h_off, w*stride + s*s_step + w_off)
    int cchunks, R;       // row groups = taps * cchunks
    int m_blocks, n_blocks, ksplit;
    int tiles_h, tiles_w;
    long long total_pt, pt_per_split;
};

template <int BLOCK_N, bool SPLIT>
__global__ void __launch_bounds__(NUM_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_x_hi, const __grid_constant__ CUtensorMap map_x_lo,
                const __grid_constant__ CUtensorMap map_g_hi, const __grid_constant__ CUtensorMap map_g_lo,
                const WgradParams p) {
    using C = WCfg<BLOCK_N, SPLIT>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
    uint64_t* empty_bar = full_bar + C::STAGES;
    uint64_t* done_bar = empty_bar + C::STAGES;
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(done_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    int item = blockIdx.x;
    const int ks = item % p.ksplit; item /= p.ksplit;
    const int nb = item % p.n_blocks;
    const int mb = item / p.n_blocks;
    const long long pt_begin = ks * p.pt_per_split;
    long long pt_end = pt_begin + p.pt_per_split;
    if (pt_end > p.total_pt) pt_end = p.total_pt;
    const int rg0 = 2 * mb, rg1 = (2 * mb + 1 < p.R) ? 2 * mb + 1 : 2 * mb;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_x_hi);
        tma_prefetch_desc(&map_g_hi);
        for (int i = 0; i < C::STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        mbar_init(done_bar, 1);
        mbar_fence_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_holder, C::TMEM_COLS < 32 ? 32 : C::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    if (warp == 0) {
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            int tapv[2], cbv[2];
            tapv[0] = rg0 / p.cchunks; cbv[0] = rg0 % p.cchunks;
            tapv[1] = rg1 / p.cchunks; cbv[1] = rg1 % p.cchunks;
            for (long long pt = pt_begin; pt < pt_end; ++pt) {
                long long t = pt;
                const int tw = static_cast<int>(t % p.tiles_w); t /= p.tiles_w;
                const int th = static_cast<int>(t % p.tiles_h);
                const int n = static_cast<int>(t / p.tiles_h);
                const int h0 = th * PT_H, w0 = tw * PT_W;
                mbar_wait(&empty_bar[stage], phase ^ 1u);
                uint8_t* st = smem + stage * C::STAGE_BYTES;
                mbar_expect_tx(&full_bar[stage], C::STAGE_BYTES);
#pragma unroll
                for (int sub = 0; sub < 2; ++sub) {
                    const int r = tapv[sub] / p.KW, s = tapv[sub] % p.KW;
                    // stride 2: the x tensor map has elementStrides 2, so the box picks every other pixel
                    const int xw = w0 * p.stride + s * p.s_step + p.w_off, xh = h0 * p.stride + r + p.h_off;
                    tma_load_4d(st + sub * SUB_BYTES, &map_x_hi, &full_bar[stage], cbv[sub] * 64, xw, xh, n);
                    if (SPLIT)
                        tma_load_4d(st + C::A_BYTES + sub * SUB_BYTES, &map_x_lo, &full_bar[stage], cbv[sub] * 64, xw, xh, n);
                }
                uint8_t* sb = st + C::A_BYTES * C::PLANES;
#pragma unroll
                for (int blk = 0; blk < C::NB; ++blk) {
                    tma_load_4d(sb + blk * SUB_BYTES, &map_g_hi, &full_bar[stage], nb * BLOCK_N + blk * 64, w0, h0, n);
                    if (SPLIT)
                        tma_load_4d(sb + C::B_BYTES + blk * SUB_BYTES, &map_g_lo, &full_bar[stage],
                                    nb * BLOCK_N + blk * 64, w0, h0, n);
                }
                if (++stage == C::STAGES) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc_bf16(128, BLOCK_N, 1, 1);
            const uint64_t dbase = desc_base(SUB_BYTES, 1024, kSwizzle128);   // MN-major: LBO = distance between 64-wide blocks
            int stage = 0;
            uint32_t phase = 0;
            uint32_t first = 1;
            for (long long pt = pt_begin; pt < pt_end; ++pt) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint32_t a_hi = smem_u32(smem + stage * C::STAGE_BYTES);
                const uint64_t da_hi = dbase + (a_hi >> 4);
                const uint64_t da_lo = da_hi + (C::A_BYTES >> 4);
                const uint64_t db_hi = da_hi + ((C::A_BYTES * C::PLANES) >> 4);
                const uint64_t db_lo = db_hi + (C::B_BYTES >> 4);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    constexpr int KSTEP = 2048 >> 4;          // 16 pixel rows of 128 bytes
                    umma_f16(tmem_base, da_hi + k * KSTEP, db_hi + k * KSTEP, idesc, (first && k == 0) ? 0u : 1u);
                    if (SPLIT) {
                        umma_f16(tmem_base, da_lo + k * KSTEP, db_hi + k * KSTEP, idesc, 1u);
                        umma_f16(tmem_base, da_hi + k * KSTEP, db_lo + k * KSTEP, idesc, 1u);
                    }
                }
                first = 0;
                umma_commit(&empty_bar[stage]);
                if (++stage == C::STAGES) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
            umma_commit(done_bar);
        }
    } else if (pt_end > pt_begin) {
        const int q = warp & 3;
        const int row = q * 32 + lane;           // M row: sub-tile = row / 64, ci = row % 64
        const int sub = row >> 6;
        const int rg = sub == 0 ? rg0 : rg1;
        const bool live = (sub == 0) || (2 * mb + 1 < p.R);
        mbar_wait(done_bar, 0);
        tc_fence_after();
        float* dst = p.ws + (static_cast<size_t>(rg) * 64 + (row & 63)) * p.Cout_p + nb * BLOCK_N;
#pragma unroll 1
        for (int c = 0; c < BLOCK_N; c += 32) {
            uint32_t v[32];
            tmem_ld_32x32(tmem_base + c + (static_cast<uint32_t>(q * 32) << 16), v);
            tmem_ld_wait();
            if (live) {
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    atomicAdd(reinterpret_cast<float4*>(dst + c + j),
                              make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                          __uint_as_float(v[j + 3])));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, C::TMEM_COLS < 32 ? 32 : C::TMEM_COLS);
    }
}

// ws[(tap*cchunks + cb)*64 + cil][co]  ->  dw[co][ci][r][s]
__global__ void wgrad_scatter_kernel(const float* __restrict__ ws, float* __restrict__ dw, int Cin, int Cout, int Cin_p,
                                     int Cout_p, int KH, int KW, int accumulate) {
    const long long total = 1LL * Cout * Cin * KH * KW;
    for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
        long long t = i;
        const int s = static_cast<int>(t % KW); t /= KW;
        const int r = static_cast<int>(t % KH); t /= KH;
        const int ci = static_cast<int>(t % Cin);
        const int co = static_cast<int>(t / Cin);
        const int tap = r * KW + s;
        const float v = ws[(static_cast<size_t>(tap) * Cin_p + ci) * Cout_p + co];
        dw[i] = accumulate ? dw[i] + v : v;
    }
}

__global__ void channel_sum_kernel2(SplitCPtr v, int ld, long long npix, int C, float* out, int accumulate) {
    // grid.x = channel groups of 32, grid.y = pixel slices; partial sums combined with float atomics after
    // an in-block double reduction.  `out` must be pre-zeroed when accumulate == 0 (done by the wrapper).
    __shared__ double red[8][32];
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    const int py = threadIdx.x >> 5;
    double acc = 0.0;
    if (c < C)
        for (long long p = blockIdx.y * 8LL + py; p < npix; p += 8LL * gridDim.y)
            acc += load_split(v, static_cast<size_t>(p) * ld + c);
    red[py][threadIdx.x & 31] = acc;
    __syncthreads();
    if (py == 0 && c < C) {
        double t = 0.0;
        for (int y = 0; y < 8; ++y) t += red[y][threadIdx.x];
        atomicAdd(out + c, static_cast<float>(t));
    }
}

}  // namespace

bool wgrad_tc_supported(int Cin_p, int Cout_p, int KH, int KW, int stride) {
    return (stride == 1 || stride == 2) && Cin_p % 64 == 0 && Cout_p % 64 == 0 && KH <= 9 && KW <= 9;
}

size_t wgrad_tc_workspace(int N, int H, int W, int Cin_p, int Cout_p, int KH, int KW, int pad) {
    (void)N; (void)H; (void)W; (void)pad;
    return sizeof(float) * static_cast<size_t>(KH) * KW * Cin_p * Cout_p;
}

// Vectorised form: a thread owns 8 consecutive channels (16-byte loads of both planes) and strides over the pixels; the
// 256 / (C/8) pixel lanes of a block meet in shared memory, one float atomic per (block, channel).
__global__ void channel_sum_vec_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo, int ld,
                                       long long npix, int C, int cg, float* out) {
    __shared__ float red[256 * 8];
    const int g = threadIdx.x % cg, lane = threadIdx.x / cg, lanes = blockDim.x / cg;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (lane < lanes) {
        for (long long p = blockIdx.x * 1LL * lanes + lane; p < npix; p += 1LL * gridDim.x * lanes) {
            const size_t off = static_cast<size_t>(p) * ld + g * 8;
            const uint4 a = *reinterpret_cast<const uint4*>(hi + off);
            const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float2 f = __bfloat1622float2(pa[k]);
                acc[2 * k] += f.x;
                acc[2 * k + 1] += f.y;
            }
            if (lo) {
                const uint4 b = *reinterpret_cast<const uint4*>(lo + off);
                const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float2 f = __bfloat1622float2(pb[k]);
                    acc[2 * k] += f.x;
                    acc[2 * k + 1] += f.y;
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) red[threadIdx.x * 8 + k] = (lane < lanes) ? acc[k] : 0.f;
    __syncthreads();
    if (threadIdx.x < cg * 8) {
        const int c = threadIdx.x;               // channel = group * 8 + k
        const int gg = c >> 3, k = c & 7;
        float t = 0.f;
        for (int l = 0; l < lanes; ++l) t += red[(l * cg + gg) * 8 + k];
        if (c < C) atomicAdd(out + c, t);
    }
}

int channel_sum_split(const void* hi, const void* lo, int ld, long long npix, int C, float* out, int accumulate,
                      cudaStream_t stream) {
    if (!accumulate) FCD_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(float) * C, stream));
    const int cg = (C + 7) / 8;
    const bool vec = ld % 8 == 0 && cg * 8 <= ld && cg <= 32 && (reinterpret_cast<uintptr_t>(hi) & 15) == 0 &&
                     (!lo || (reinterpret_cast<uintptr_t>(lo) & 15) == 0);
    if (vec) {
        const int lanes = 256 / cg;
        long long blocks = (npix + lanes - 1) / lanes;
        const long long cap = 4LL * sm_count();
        if (blocks > cap) blocks = cap;
        if (blocks < 1) blocks = 1;
        channel_sum_vec_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>((const __nv_bfloat16*)hi, (const __nv_bfloat16*)lo,
                                                                                  ld, npix, C, cg, out);
        FCD_LAUNCH_OK();
        return FCD_OK;
    }
    int gy = static_cast<int>(npix / 2048);
    if (gy < 1) gy = 1;
    if (gy > 1024) gy = 1024;
    SplitCPtr v{(const __nv_bfloat16*)hi, (const __nv_bfloat16*)lo};
    channel_sum_kernel2<<<dim3((C + 31) / 32, gy), 256, 0, stream>>>(v, ld, npix, C, out, accumulate);
    FCD_LAUNCH_OK();
    return FCD_OK;
}

template <int BLOCK_N, bool SPLIT>
static int launch_wgrad(const CUtensorMap& mxh, const CUtensorMap& mxl, const CUtensorMap& mgh, const CUtensorMap& mgl,
                        const WgradParams& p, cudaStream_t stream) {
    using C = WCfg<BLOCK_N, SPLIT>;
    static PerDeviceOnce attr_once;
    if (!attr_once()) {
        FCD_CUDA_OK(cudaFuncSetAttribute(wgrad_tc_kernel<BLOCK_N, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         C::SMEM_BYTES));
        attr_once() = true;
    }
    const unsigned grid = static_cast<unsigned>(p.m_blocks) * p.n_blocks * p.ksplit;
    wgrad_tc_kernel<BLOCK_N, SPLIT><<<grid, NUM_THREADS, C::SMEM_BYTES, stream>>>(mxh, mxl, mgh, mgl, p);
    FCD_LAUNCH_OK();
    return FCD_OK;
}

// =====================================================================================================
// Halo-reuse weight gradient (stride 1): one CTA owns (64 input channels) x (64 output channels) x (up to 16 taps).
// Per K tile of 8 x 8 output pixels it stages the x tile ONCE with its halo ({64 ch, 8 + (n_s-1)*s_step, 8 + n_r-1} TMA
// box) and the dz tile once; every tap is a ROW-SHIFTED VIEW of the staged x tile (MN-major UMMA descriptor whose start
// address is moved by (r * halo_w + s * s_step) pixels, stride-byte-offset = halo_w * 128 so the 8-pixel groups follow
// the halo pitch).  Two taps form one M = 128 operand: the descriptor's leading-byte-offset is the distance between the
// two views.  All tap accumulators live in TMEM (64 columns per tap pair).  Versus wgrad_tc_kernel this moves 3.6x
// fewer bytes L2 -> shared memory for a 3x3 layer (the x tile is no longer re-fetched per tap pair).
// Descriptor semantics (SBO not a multiple of 1024, small LBO, shifted start) are pinned by scripts/probe_halo.py.
// =====================================================================================================
constexpr int HP = 8;                 // K tile = HP x HP output pixels
constexpr int H_MAX_TAPS = 16;        // 8 accumulators x 64 TMEM columns
constexpr int H_SMEM_BUDGET = 200 * 1024;
constexpr int H_MAX_GROUPS = 12;

struct WHaloParams {
    float* ws;
    int N, GH, GW;
    int Cin_p, Cout_p;
    int n_r, n_s, s_step, h_off, w_off;
    int halo_w, halo_h;
    int x_bytes;              // bytes of one staged x plane (halo_w * halo_h * 128), slot rounded up to 1024
    int x_slot, stage_bytes, stages;
    int taps_per_group, tap_groups;
    int cchunks, n_blocks;
    int tiles_h, tiles_w;
    long long total_pt;
    // tap groups differ in their UMMA count (an odd group pads to a whole tap pair), so each group gets its own split-K
    // factor, proportional to its pair count: every CTA then issues about the same number of UMMAs.
    int ksplit_g[H_MAX_GROUPS], cta_start[H_MAX_GROUPS + 1];
};

template <bool SPLIT>
__global__ void __launch_bounds__(NUM_THREADS, 1)
wgrad_halo_kernel(const __grid_constant__ CUtensorMap map_x_hi, const __grid_constant__ CUtensorMap map_x_lo,
                  const __grid_constant__ CUtensorMap map_g_hi, const __grid_constant__ CUtensorMap map_g_lo,
                  const WHaloParams p) {
    constexpr int PLANES = SPLIT ? 2 : 1;
    constexpr int G_BYTES = 64 * 128;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.stages * p.stage_bytes);
    uint64_t* empty_bar = full_bar + 8;
    uint64_t* done_bar = empty_bar + 8;
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(done_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int tg = 0;
    while (tg + 1 < p.tap_groups && static_cast<int>(blockIdx.x) >= p.cta_start[tg + 1]) ++tg;
    int item = blockIdx.x - p.cta_start[tg];
    const int ksplit = p.ksplit_g[tg];
    const int ks = item % ksplit; item /= ksplit;
    const int nb = item % p.n_blocks;
    const int cb = item / p.n_blocks;
    const int total_taps = p.n_r * p.n_s;
    const int tap0 = tg * p.taps_per_group;
    const int ntaps = (total_taps - tap0 < p.taps_per_group) ? total_taps - tap0 : p.taps_per_group;
    const int npairs = (ntaps + 1) >> 1;
    // Split-K is ROUND-ROBIN over the K tiles (tile = ks, ks + ksplit, ...), not a contiguous range per CTA: every tap group
    // then sweeps the tensor in the same global order at the same rate (the split factors are proportional to the groups'
    // UMMA counts), so the x / dz tiles one group pulls from HBM are still in L2 when the other group asks for them
    // (ncu: 1128 MB of DRAM reads per launch with contiguous ranges = both groups streaming the operands from HBM).
    const long long pt_begin = ks, pt_end = p.total_pt, pt_step = ksplit;
    // split precision: [dz_hi | dz_lo] is ONE N = 128 operand, so a tap pair owns 128 accumulator columns
    constexpr int PAIR_COLS = SPLIT ? 128 : 64;
    const uint32_t need_cols = npairs * PAIR_COLS;
    const uint32_t tmem_cols = need_cols <= 64 ? 64 : need_cols <= 128 ? 128 : need_cols <= 256 ? 256 : 512;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_x_hi);
        tma_prefetch_desc(&map_g_hi);
        for (int i = 0; i < p.stages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        mbar_init(done_bar, 1);
        mbar_fence_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_holder, tmem_cols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    const uint32_t stage_tx = static_cast<uint32_t>((p.x_bytes + G_BYTES) * PLANES);

    if (warp == 0) {
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (long long pt = pt_begin; pt < pt_end; pt += pt_step) {
                long long t = pt;
                const int tw = static_cast<int>(t % p.tiles_w); t /= p.tiles_w;
                const int th = static_cast<int>(t % p.tiles_h);
                const int n = static_cast<int>(t / p.tiles_h);
                const int h0 = th * HP, w0 = tw * HP;
                mbar_wait(&empty_bar[stage], phase ^ 1u);
                uint8_t* st = smem + stage * p.stage_bytes;
                mbar_expect_tx(&full_bar[stage], stage_tx);
                tma_load_4d(st, &map_x_hi, &full_bar[stage], cb * 64, w0 + p.w_off, h0 + p.h_off, n);
                if (SPLIT) tma_load_4d(st + p.x_slot, &map_x_lo, &full_bar[stage], cb * 64, w0 + p.w_off, h0 + p.h_off, n);
                uint8_t* sg = st + p.x_slot * PLANES;
                tma_load_4d(sg, &map_g_hi, &full_bar[stage], nb * 64, w0, h0, n);
                if (SPLIT) tma_load_4d(sg + G_BYTES, &map_g_lo, &full_bar[stage], nb * 64, w0, h0, n);
                if (++stage == p.stages) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc_bf16(128, 64, 1, 1);
            constexpr uint32_t idesc_wide = make_idesc_bf16(128, 128, 1, 1);
            const uint32_t sbo_a = static_cast<uint32_t>(p.halo_w) * 128u;   // 8-pixel groups follow the halo row pitch
            const uint32_t kstep_a = (2u * sbo_a) >> 4;                     // UMMA K = 16 pixels = two tile rows (16-byte units)
            // dz_hi and dz_lo tiles are adjacent (LBO = G_BYTES): with N = 128 one UMMA yields x_hi*dz_hi in columns [0,64)
            // and x_hi*dz_lo in [64,128); the epilogue adds the halves.
            const uint64_t dbase_b = desc_base(G_BYTES, 1024, kSwizzle128);
            // per tap pair: the A descriptor (LBO = distance between the two tap views) relative to the staged x tile
            uint64_t da_pair[H_MAX_TAPS / 2];
#pragma unroll
            for (int pr = 0; pr < H_MAX_TAPS / 2; ++pr) {
                da_pair[pr] = 0;
                if (pr >= npairs) continue;
                int ta = tap0 + 2 * pr, tb = ta + 1;
                if (tb >= tap0 + ntaps) {      // odd tap count: the last pair re-uses the previous tap as its first half
                    tb = ta;
                    ta = ta - 1;
                }
                const uint32_t off_a = static_cast<uint32_t>((ta / p.n_s) * p.halo_w + (ta % p.n_s) * p.s_step) * 128u;
                const uint32_t off_b = static_cast<uint32_t>((tb / p.n_s) * p.halo_w + (tb % p.n_s) * p.s_step) * 128u;
                da_pair[pr] = desc_base(off_b - off_a, sbo_a, kSwizzle128) + (off_a >> 4);
            }
            int stage = 0;
            uint32_t phase = 0;
            uint32_t first = 1;
            for (long long pt = pt_begin; pt < pt_end; pt += pt_step) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint32_t x_hi = smem_u32(smem + stage * p.stage_bytes);
                const uint32_t xs = x_hi >> 4, xl = (x_hi + p.x_slot) >> 4;
                const uint64_t db_hi = dbase_b + ((x_hi + p.x_slot * PLANES) >> 4);
#pragma unroll
                for (int pr = 0; pr < H_MAX_TAPS / 2; ++pr) {
                    if (pr >= npairs) break;
                    const uint32_t d_tmem = tmem_base + pr * PAIR_COLS;
                    const uint64_t da_hi = da_pair[pr] + xs, da_lo = da_pair[pr] + xl;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (!SPLIT) {
                            umma_f16(d_tmem, da_hi + k * kstep_a, db_hi + k * (2048 >> 4), idesc, (first && k == 0) ? 0u : 1u);
                        } else {
                            umma_f16(d_tmem, da_hi + k * kstep_a, db_hi + k * (2048 >> 4), idesc_wide, (first && k == 0) ? 0u : 1u);
                            umma_f16(d_tmem, da_lo + k * kstep_a, db_hi + k * (2048 >> 4), idesc, 1u);
                        }
                    }
                }
                first = 0;
                umma_commit(&empty_bar[stage]);
                if (++stage == p.stages) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
            umma_commit(done_bar);
        }
    } else if (pt_end > pt_begin) {
        const int q = warp & 3;
        const int row = q * 32 + lane;      // accumulator row: half = row / 64 (first / second tap of the pair), ci = row % 64
        const int half = row >> 6;
        mbar_wait(done_bar, 0);
        tc_fence_after();
        for (int pr = 0; pr < npairs; ++pr) {
            int ta = tap0 + 2 * pr, tb = ta + 1;
            bool live = true;
            if (tb >= tap0 + ntaps) {
                tb = ta;
                ta = ta - 1;
                live = half == 1;            // the duplicated first half is discarded
            }
            const int tap = half ? tb : ta;
            float* dst = p.ws + (static_cast<size_t>(tap * p.cchunks + cb) * 64 + (row & 63)) * p.Cout_p + nb * 64;
#pragma unroll 1
            for (int c = 0; c < 64; c += 32) {
                uint32_t v[32];
                float f[32];
                tmem_ld_32x32(tmem_base + pr * PAIR_COLS + c + (static_cast<uint32_t>(q * 32) << 16), v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                if (SPLIT) {
                    tmem_ld_32x32(tmem_base + pr * PAIR_COLS + 64 + c + (static_cast<uint32_t>(q * 32) << 16), v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] += __uint_as_float(v[j]);
                }
                if (live) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        atomicAdd(reinterpret_cast<float4*>(dst + c + j), make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

static bool wgrad_halo_fits(int n_r, int n_s, int s_step, bool split, WHaloParams* out) {
    const int halo_w = HP + (n_s - 1) * s_step, halo_h = HP + n_r - 1;
    if (n_r * n_s < 2 || halo_w > 256 || halo_h > 256) return false;
    const int x_bytes = halo_w * halo_h * 128;
    const int x_slot = (x_bytes + 1023) & ~1023;
    const int stage_bytes = (x_slot + 64 * 128) * (split ? 2 : 1);
    int stages = H_SMEM_BUDGET / stage_bytes;
    if (stages < 2) return false;
    if (stages > 6) stages = 6;
    if (out) {
        out->halo_w = halo_w; out->halo_h = halo_h; out->x_bytes = x_bytes; out->x_slot = x_slot;
        out->stage_bytes = stage_bytes; out->stages = stages;
    }
    return true;
}

template <bool SPLIT>
static int launch_wgrad_halo(const CUtensorMap& mxh, const CUtensorMap& mxl, const CUtensorMap& mgh, const CUtensorMap& mgl,
                             const WHaloParams& p, unsigned grid, cudaStream_t stream) {
    static PerDeviceOnce attr_once;
    const int smem_bytes = H_SMEM_BUDGET + 1024 + 256;
    if (!attr_once()) {
        FCD_CUDA_OK(cudaFuncSetAttribute(wgrad_halo_kernel<SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        attr_once() = true;
    }
    wgrad_halo_kernel<SPLIT><<<grid, NUM_THREADS, smem_bytes, stream>>>(mxh, mxl, mgh, mgl, p);
    FCD_LAUNCH_OK();
    return FCD_OK;
}

// 4D NHWC bf16 map with an arbitrary box (halo tiles); defined in conv_tc.cu
int make_act_tmap(CUtensorMap* m, const void* base, int C, int W, int H, int N, int ld, int box_c, int box_w, int box_h,
                  int estride_w, int estride_h);

// Core: dW[(r, s)][ci][co] (+)= sum_{n, h < GH, w < GW} x[n, h*stride + r + h_off, w*stride + s*s_step + w_off, ci] * dz[n, h, w, co]
// for r < n_r, s < n_s (x out of bounds = 0), written as fp32 OIHW [Cout][Cin][n_r][n_s].
static int wgrad_tc_core(const void* x_hi, const void* x_lo, int x_ld, int XH, int XW, const void* dz_hi, const void* dz_lo,
                         int dz_ld, int GH, int GW, float* dw, int N, int Cin, int Cin_p, int Cout, int Cout_p, int n_r,
                         int n_s, int stride, int h_off, int w_off, int s_step, int accumulate, void* workspace,
                         size_t workspace_bytes, cudaStream_t stream) {
    const size_t need = sizeof(float) * static_cast<size_t>(n_r) * n_s * Cin_p * Cout_p;
    FCD_CHECK_ARG(workspace && workspace_bytes >= need, "conv2d_wgrad_tc: workspace too small (%zu < %zu)",
                  workspace_bytes, need);
    FCD_CHECK_ARG(x_ld % 8 == 0 && dz_ld % 8 == 0, "conv2d_wgrad_tc: pitches must be multiples of 8");
    const bool split = x_lo && dz_lo;
    WHaloParams hp;
    if (stride == 1 && g_wgrad_halo_enabled && wgrad_halo_fits(n_r, n_s, s_step, split, &hp)) {
        hp.ws = static_cast<float*>(workspace);
        hp.N = N; hp.GH = GH; hp.GW = GW; hp.Cin_p = Cin_p; hp.Cout_p = Cout_p;
        hp.n_r = n_r; hp.n_s = n_s; hp.s_step = s_step; hp.h_off = h_off; hp.w_off = w_off;
        {   // tap groups: at most 512 TMEM columns per CTA (16 taps; 8 in split precision), groups balanced, each >= 2 taps
            const int total = n_r * n_s, cap = split ? H_MAX_TAPS / 2 : H_MAX_TAPS;
            hp.tap_groups = (total + cap - 1) / cap;
            FCD_CHECK_ARG(hp.tap_groups <= H_MAX_GROUPS, "conv2d_wgrad_tc: too many taps (%d)", total);
            hp.taps_per_group = (total + hp.tap_groups - 1) / hp.tap_groups;
        }
        hp.cchunks = Cin_p / 64; hp.n_blocks = Cout_p / 64;
        hp.tiles_h = ceil_div(GH, HP); hp.tiles_w = ceil_div(GW, HP);
        hp.total_pt = 1LL * N * hp.tiles_h * hp.tiles_w;
        // CTA budget: one wave (1 CTA per SM is resident).  Each (cb, nb) block gets `slots` CTAs, shared between the tap
        // groups in proportion to their tap-pair counts.
        const long long mn_blocks = 1LL * hp.cchunks * hp.n_blocks;
        const long long sms = sm_count();
        int pairs_g[H_MAX_GROUPS], pairs_total = 0;
        for (int g = 0; g < hp.tap_groups; ++g) {
            const int t0 = g * hp.taps_per_group;
            const int nt = (n_r * n_s - t0 < hp.taps_per_group) ? n_r * n_s - t0 : hp.taps_per_group;
            pairs_g[g] = (nt + 1) / 2;
            pairs_total += pairs_g[g];
        }
        long long slots = sms / mn_blocks;
        if (slots < hp.tap_groups) slots = hp.tap_groups;
        unsigned grid = 0;
        for (int g = 0; g < hp.tap_groups; ++g) {
            long long ks = (slots * pairs_g[g] + pairs_total / 2) / pairs_total;
            if (ks < 1) ks = 1;
            if (ks > hp.total_pt) ks = hp.total_pt;
            const long long per = (hp.total_pt + ks - 1) / ks;
            hp.ksplit_g[g] = static_cast<int>((hp.total_pt + per - 1) / per);
            hp.cta_start[g] = static_cast<int>(grid);
            grid += static_cast<unsigned>(hp.ksplit_g[g] * mn_blocks);
        }
        hp.cta_start[hp.tap_groups] = static_cast<int>(grid);
        CUtensorMap mxh, mxl, mgh, mgl;
        int rc;
        if ((rc = make_act_tmap(&mxh, x_hi, Cin_p, XW, XH, N, x_ld, 64, hp.halo_w, hp.halo_h, 1, 1))) return rc;
        if ((rc = make_act_tmap(&mgh, dz_hi, Cout_p, GW, GH, N, dz_ld, 64, HP, HP, 1, 1))) return rc;
        if (split) {
            if ((rc = make_act_tmap(&mxl, x_lo, Cin_p, XW, XH, N, x_ld, 64, hp.halo_w, hp.halo_h, 1, 1))) return rc;
            if ((rc = make_act_tmap(&mgl, dz_lo, Cout_p, GW, GH, N, dz_ld, 64, HP, HP, 1, 1))) return rc;
        } else {
            mxl = mxh;
            mgl = mgh;
        }
        FCD_CUDA_OK(cudaMemsetAsync(workspace, 0, need, stream));
        rc = split ? launch_wgrad_halo<true>(mxh, mxl, mgh, mgl, hp, grid, stream)
                   : launch_wgrad_halo<false>(mxh, mxl, mgh, mgl, hp, grid, stream);
        if (rc) return rc;
        const long long total = 1LL * Cout * Cin * n_r * n_s;
        const int blocks = static_cast<int>((total + 255) / 256 > 2048 ? 2048 : (total + 255) / 256);
        wgrad_scatter_kernel<<<blocks, 256, 0, stream>>>(hp.ws, dw, Cin, Cout, Cin_p, Cout_p, n_r, n_s, accumulate);
        FCD_LAUNCH_OK();
        return FCD_OK;
    }
    const int block_n = (Cout_p % 128 == 0) ? 128 : 64;
    WgradParams p;
    p.ws = static_cast<float*>(workspace);
    p.N = N; p.OH = GH; p.OW = GW;
    p.Cin_p = Cin_p; p.Cout_p = Cout_p; p.KH = n_r; p.KW = n_s; p.stride = stride;
    p.h_off = h_off; p.w_off = w_off; p.s_step = s_step;
    p.cchunks = Cin_p / 64;
    p.R = n_r * n_s * p.cchunks;
    p.m_blocks = (p.R + 1) / 2;
    p.n_blocks = Cout_p / block_n;
    p.tiles_h = ceil_div(GH, PT_H);
    p.tiles_w = ceil_div(GW, PT_W);
    p.total_pt = 1LL * N * p.tiles_h * p.tiles_w;
    // split-K so that the grid fills whole waves: one CTA per SM is resident (192 KB of shared memory), so the CTA count
    // should be just BELOW a multiple of the SM count (a 300-CTA grid on 148 SMs runs three waves, the last one with 4 CTAs).
    const long long mn = 1LL * p.m_blocks * p.n_blocks;
    const long long sms = sm_count();
    long long waves = (mn + sms - 1) / sms;          // waves needed without any split
    if (waves < 2) waves = 2;
    long long ks = (waves * sms) / mn;               // floor: ks * mn <= waves * sms
    if (ks < 1) ks = 1;
    if (ks > p.total_pt) ks = p.total_pt;
    p.pt_per_split = (p.total_pt + ks - 1) / ks;
    p.ksplit = static_cast<int>((p.total_pt + p.pt_per_split - 1) / p.pt_per_split);

    CUtensorMap mxh, mxl, mgh, mgl;
    int rc;
    if ((rc = make_act_tmap(&mxh, x_hi, Cin_p, XW, XH, N, x_ld, 64, PT_W, PT_H, stride, stride))) return rc;
    if ((rc = make_act_tmap(&mgh, dz_hi, Cout_p, GW, GH, N, dz_ld, 64, PT_W, PT_H, 1, 1))) return rc;
    if (split) {
        if ((rc = make_act_tmap(&mxl, x_lo, Cin_p, XW, XH, N, x_ld, 64, PT_W, PT_H, stride, stride))) return rc;
        if ((rc = make_act_tmap(&mgl, dz_lo, Cout_p, GW, GH, N, dz_ld, 64, PT_W, PT_H, 1, 1))) return rc;
    } else {
        mxl = mxh;
        mgl = mgh;
    }
    FCD_CUDA_OK(cudaMemsetAsync(workspace, 0, need, stream));
    if (block_n == 128)
        rc = split ? launch_wgrad<128, true>(mxh, mxl, mgh, mgl, p, stream)
                   : launch_wgrad<128, false>(mxh, mxl, mgh, mgl, p, stream);
    else
        rc = split ? launch_wgrad<64, true>(mxh, mxl, mgh, mgl, p, stream)
                   : launch_wgrad<64, false>(mxh, mxl, mgh, mgl, p, stream);
    if (rc) return rc;
    const long long total = 1LL * Cout * Cin * n_r * n_s;
    const int blocks = static_cast<int>((total + 255) / 256 > 2048 ? 2048 : (total + 255) / 256);
    wgrad_scatter_kernel<<<blocks, 256, 0, stream>>>(p.ws, dw, Cin, Cout, Cin_p, Cout_p, n_r, n_s, accumulate);
    FCD_LAUNCH_OK();
    return FCD_OK;
}

int conv2d_wgrad_tc(const void* x_hi, const void* x_lo, int x_ld, const void* dz_hi, const void* dz_lo, int dz_ld,
                    float* dw, int N, int H, int W, int Cin, int Cin_p, int Cout, int Cout_p, int KH, int KW, int stride,
                    int pad, int accumulate, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    const int OH = (H + 2 * pad - KH) / stride + 1, OW = (W + 2 * pad - KW) / stride + 1;
    return wgrad_tc_core(x_hi, x_lo, x_ld, H, W, dz_hi, dz_lo, dz_ld, OH, OW, dw, N, Cin, Cin_p, Cout, Cout_p, KH, KW, stride,
                         -pad, -pad, 1, accumulate, workspace, workspace_bytes, stream);
}

// Generic tap-list weight gradient (stride 1): see wgrad_tc_core; serves the 4-pixel channel-packed 13-band layers.
int conv2d_taps_wgrad_tc(const void* x_hi, const void* x_lo, int x_ld, int XH, int XW, const void* dz_hi, const void* dz_lo,
                         int dz_ld, int GH, int GW, float* dw, int N, int Cin, int Cin_p, int Cout, int Cout_p, int n_r,
                         int n_s, int dh0, int dw0, int dw_step, int accumulate, void* workspace, size_t workspace_bytes,
                         cudaStream_t stream) {
    FCD_CHECK_ARG(Cin_p % 64 == 0 && Cout_p % 64 == 0 && n_r > 0 && n_s > 0, "conv2d_taps_wgrad_tc: bad dims");
    return wgrad_tc_core(x_hi, x_lo, x_ld, XH, XW, dz_hi, dz_lo, dz_ld, GH, GW, dw, N, Cin, Cin_p, Cout, Cout_p, n_r, n_s, 1, dh0,
                         dw0, dw_step, accumulate, workspace, workspace_bytes, stream);
}

}  // namespace fcd
