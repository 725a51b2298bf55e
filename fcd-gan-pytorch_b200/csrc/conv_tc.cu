// conv_tc.cu — tcgen05 + TMA implicit-GEMM convolution (forward / stride-1 dgrad) for sm_100a.
//
// Replaces the nn.Conv2d calls of the reference networks (Module.py:26,29,155,177,180 and, through
// dgrad, their autograd backward).  GEMM view:  M = output pixels, N = output channels,
// K = taps x input channels.
//
//   * One CTA tile = 8 x 16 output pixels of one image (M = 128) x BLOCK_N output channels.
//   * For every (tap, 64-channel chunk) k-block a TMA box {64 ch, 16 w, 8 h, 1 n} of the NHWC input,
//     shifted by the tap offset, lands in 128B-swizzled shared memory: exactly the K-major A operand
//     of a 128 x 64 UMMA tile.  Out-of-bounds rows/columns are zero-filled by TMA = the conv padding.
//   * Accumulators live in TMEM (double buffered, BLOCK_N fp32 columns each); one elected thread
//     issues tcgen05.mma; with split-bf16 operands three MMAs per k-step (hi*hi + lo*hi + hi*lo)
//     recover fp32-class products.
//   * Warp roles: warp 0 TMA producer, warp 1 MMA issuer (+TMEM alloc), warps 2-5 epilogue
//     (tcgen05.ld -> +bias -> fp32 NHWC store, optional per-channel sum / sum-of-squares for BN).
//   * Persistent: grid = #SMs, static round-robin over tiles.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <initializer_list>

#include "fcd_common.cuh"
#include "fcd_tc.cuh"

namespace fcd {

using namespace tc;

namespace {

constexpr int TILE_H = 8;
constexpr int TILE_W = 16;
constexpr int BLOCK_M = TILE_H * TILE_W;  // 128
constexpr int BLOCK_K = 64;               // bf16 elements = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 192;
constexpr int EPI_WARP0 = 2;

template <int BLOCK_N, bool SPLIT>
struct Cfg {
    static constexpr int PLANES = SPLIT ? 2 : 1;
    static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
    static constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
    static constexpr int STAGE_BYTES = (A_BYTES + B_BYTES) * PLANES;
    static constexpr int SMEM_BUDGET = 200 * 1024;
    static constexpr int STAGES_RAW = SMEM_BUDGET / STAGE_BYTES;
    static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
    // split precision stacks [B_hi ; B_lo] into ONE UMMA of N = 2*BLOCK_N (x_hi*w_hi and x_hi*w_lo land in adjacent column
    // ranges, the epilogue adds them), so an accumulator is ACC_COLS = 2*BLOCK_N wide; two accumulators are in flight.
    static constexpr int ACC_COLS = SPLIT ? 2 * BLOCK_N : BLOCK_N;
    static constexpr int TMEM_COLS = 2 * ACC_COLS < 32 ? 32 : 2 * ACC_COLS;  // power of two for 32..512
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

struct ConvTcParams {
    const float* bias;
    const float* addend;  // optional fp32 NHWC tensor added in the epilogue (gradient accumulation / residual)
    int addend_ld;
    float* z;
    double* stat_sum;
    double* stat_sqsum;
    int z_ld;
    int N, OH, OW;       // dims of the OUTPUT tensor (addressing)
    int TOH, TOW;        // extent of the output-pixel grid this launch covers (== OH, OW except for strided-dgrad classes)
    int out_sh, out_oh, out_sw, out_ow;   // output pixel = (oh * out_sh + out_oh, ow * out_sw + out_ow)
    int Cin_p, Cout_p;
    int n_r, n_s;        // taps iterated in this launch: i in [0, n_r), j in [0, n_s)
    int csh, csw;        // TMA coordinate scale per dim (2 for a stride-2 forward conv: the tensor map has elementStrides 2)
    int dh0, dh_step, dw0, dw_step;       // input coordinate of tap (i, j): (oh0*cs + dh0 + i*dh_step, ow0*cs + dw0 + j*dw_step)
    int w_r0, w_rstep, w_s0, w_sstep, KW; // packed-weight tap index = (w_r0 + i*w_rstep) * KW + (w_s0 + j*w_sstep)
    int tiles_h, tiles_w, n_blocks;
    long long total_tiles;
};

template <int BLOCK_N, bool SPLIT>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_x_hi, const __grid_constant__ CUtensorMap map_x_lo,
               const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
               const ConvTcParams p) {
    using C = Cfg<BLOCK_N, SPLIT>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
    uint64_t* empty_bar = full_bar + C::STAGES;
    uint64_t* tmem_full = empty_bar + C::STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_x_hi);
        tma_prefetch_desc(&map_w_hi);
        if (SPLIT) {
            tma_prefetch_desc(&map_x_lo);
            tma_prefetch_desc(&map_w_lo);
        }
        for (int i = 0; i < C::STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 128);
        }
        mbar_fence_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_holder, C::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    const int cchunks = p.Cin_p / BLOCK_K;
    const int num_kb = p.n_r * p.n_s * cchunks;

    if (warp == 0) {
        // ===================== TMA producer =====================
        // elect_one(), not `lane == 0`: with an elect.sync predicate ptxas issues the uniform-datapath instructions (UBLKCP,
        // UTCHMMA) straight; a lane test makes it wrap each of them in an ELECT / BRA.U.ANY loop over the active lanes.
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                long long t = tile;
                const int nblk = static_cast<int>(t % p.n_blocks);
                t /= p.n_blocks;
                const int tw = static_cast<int>(t % p.tiles_w);
                t /= p.tiles_w;
                const int th = static_cast<int>(t % p.tiles_h);
                const int n = static_cast<int>(t / p.tiles_h);
                const int h0 = th * TILE_H * p.csh + p.dh0, w0 = tw * TILE_W * p.csw + p.dw0;
                for (int kb = 0; kb < num_kb; ++kb) {
                    const int tapi = kb / cchunks, cc = kb - tapi * cchunks;
                    const int ti = tapi / p.n_s, tj = tapi - ti * p.n_s;
                    const int r = ti * p.dh_step, s = tj * p.dw_step;   // input offsets of this tap
                    const int tap = (p.w_r0 + ti * p.w_rstep) * p.KW + (p.w_s0 + tj * p.w_sstep);
                    mbar_wait(&empty_bar[stage], phase ^ 1u);
                    uint8_t* st = smem + stage * C::STAGE_BYTES;
                    mbar_expect_tx(&full_bar[stage], C::STAGE_BYTES);
                    tma_load_4d(st, &map_x_hi, &full_bar[stage], cc * BLOCK_K, w0 + s, h0 + r, n);
                    tma_load_3d(st + C::A_BYTES * C::PLANES, &map_w_hi, &full_bar[stage], cc * BLOCK_K,
                                nblk * BLOCK_N, tap);
                    if (SPLIT) {
                        tma_load_4d(st + C::A_BYTES, &map_x_lo, &full_bar[stage], cc * BLOCK_K, w0 + s, h0 + r, n);
                        tma_load_3d(st + C::A_BYTES * 2 + C::B_BYTES, &map_w_lo, &full_bar[stage], cc * BLOCK_K,
                                    nblk * BLOCK_N, tap);
                    }
                    if (++stage == C::STAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc_bf16(BLOCK_M, BLOCK_N, 0, 0);
            constexpr uint32_t idesc_wide = make_idesc_bf16(BLOCK_M, 2 * BLOCK_N, 0, 0);
            const uint64_t dbase = desc_base(16, 1024, kSwizzle128);    // K-major, 128B swizzle, 8-row groups 1024 B apart
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * C::ACC_COLS;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t a_hi = smem_u32(smem + stage * C::STAGE_BYTES);
                    const uint64_t da_hi = dbase + (a_hi >> 4);
                    const uint64_t da_lo = da_hi + (C::A_BYTES >> 4);
                    const uint64_t db_hi = da_hi + ((C::A_BYTES * C::PLANES) >> 4);
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        constexpr int KSTEP = (UMMA_K * 2) >> 4;      // 32 bytes along the 128-byte row, in 16-byte units
                        if (!SPLIT) {
                            umma_f16(d_tmem, da_hi + k * KSTEP, db_hi + k * KSTEP, idesc, (kb | k) != 0 ? 1u : 0u);
                        } else {
                            // B_hi and B_lo are adjacent in the stage: the same descriptor with N = 2*BLOCK_N reads both.
                            //   cols [0, N)   += x_hi * w_hi  (+ x_lo * w_hi from the second, N-wide UMMA)
                            //   cols [N, 2N)  += x_hi * w_lo
                            // 2 operand fetches of the A tile instead of 3 for the same three products.
                            umma_f16(d_tmem, da_hi + k * KSTEP, db_hi + k * KSTEP, idesc_wide, (kb | k) != 0 ? 1u : 0u);
                            umma_f16(d_tmem, da_lo + k * KSTEP, db_hi + k * KSTEP, idesc, 1u);
                        }
                    }
                    umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
                    if (++stage == C::STAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                umma_commit(&tmem_full[acc]);
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1u;
                }
            }
        }
    } else {
        // ===================== epilogue: TMEM -> registers -> global =====================
        const int q = warp & 3;  // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;
        const int lh = row / TILE_W, lw = row - lh * TILE_W;
        int acc = 0;
        uint32_t acc_phase = 0;
        __shared__ float s_red[4][2][32];  // per-epilogue-warp staging for the BN statistics
        for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
            long long t = tile;
            const int nblk = static_cast<int>(t % p.n_blocks);
            t /= p.n_blocks;
            const int tw = static_cast<int>(t % p.tiles_w);
            t /= p.tiles_w;
            const int th = static_cast<int>(t % p.tiles_h);
            const int n = static_cast<int>(t / p.tiles_h);
            const int oh_l = th * TILE_H + lh, ow_l = tw * TILE_W + lw;
            const bool valid = (oh_l < p.TOH) && (ow_l < p.TOW);
            const int oh = oh_l * p.out_sh + p.out_oh, ow = ow_l * p.out_sw + p.out_ow;
            const size_t pix = static_cast<size_t>(n) * p.OH * p.OW + static_cast<size_t>(oh) * p.OW + ow;
            float* zrow = p.z + pix * p.z_ld + nblk * BLOCK_N;
            const float* arow = p.addend ? p.addend + pix * p.addend_ld + nblk * BLOCK_N : nullptr;
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < BLOCK_N; c += 32) {
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + acc * C::ACC_COLS + c + (static_cast<uint32_t>(q * 32) << 16), v);
                tmem_ld_wait();
                float f[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                if (SPLIT) {   // add the x_hi * w_lo half of the stacked accumulator
                    tmem_ld_32x32(tmem_base + acc * C::ACC_COLS + BLOCK_N + c + (static_cast<uint32_t>(q * 32) << 16), v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] += __uint_as_float(v[j]);
                }
                if (p.bias) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] += __ldg(p.bias + nblk * BLOCK_N + c + j);
                }
                if (valid) {
                    if (arow) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 a = *reinterpret_cast<const float4*>(arow + c + j);
                            f[j] += a.x; f[j + 1] += a.y; f[j + 2] += a.z; f[j + 3] += a.w;
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<float4*>(zrow + c + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
                }
                if (p.stat_sum) {
                    // column sums over this warp's 32 pixels, then one double atomic per channel
                    const int ew = warp - EPI_WARP0;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        float a = valid ? f[j] : 0.f;
                        float b = a * a;
                        a = warp_sum(a);
                        b = warp_sum(b);
                        if (lane == j) {
                            s_red[ew][0][j] = a;
                            s_red[ew][1][j] = b;
                        }
                    }
                    __syncwarp();
                    atomicAdd(p.stat_sum + nblk * BLOCK_N + c + lane, static_cast<double>(s_red[ew][0][lane]));
                    atomicAdd(p.stat_sqsum + nblk * BLOCK_N + c + lane, static_cast<double>(s_red[ew][1][lane]));
                    __syncwarp();
                }
            }
            tc_fence_before();
            mbar_arrive(&tmem_empty[acc]);
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1u;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
}


// =====================================================================================================================
// Halo-reuse convolution (stride-1 reads): conv_tc_kernel re-fetches the input tile once per tap (9 x 32 KB per 128 output
// pixels of a 3x3 layer in split precision) and is bound by the L2 -> SM fabric (ncu: lts 59 %, tensor pipe 41 %).  Here
//   * the packed weights of the CTA's output-channel block ([tap][k-chunk][w_hi ; w_lo], <= ~150 KB) are loaded into shared
//     memory ONCE per CTA and stay resident for all of its pixel tiles;
//   * per pixel tile (16 rows x 8 pixels = M 128) the input is staged once WITH its halo (TMA box {64 ch, 8 + (n_s-1)|dw|,
//     16 + (n_r-1)|dh|}); every tap's A operand is a ROW-SHIFTED VIEW of that staged tile: K-major descriptor, start address
//     moved by (r * halo_w + s) pixels, stride-byte-offset = halo_w * 128 so the 8-pixel groups follow the halo row pitch
//     (semantics pinned on the device by scripts/probe_halo.py);
//   * split precision stages the hi and lo planes as separate ring items: all taps of x_hi * [w_hi ; w_lo] (N = 2*BLOCK_N),
//     then all taps of x_lo * w_hi (N = BLOCK_N); two plane slots (one multiplied, one in flight) keep the MMA warp fed.
// L2 -> shared traffic per pixel tile drops from taps * 48 KB to 2 * 23 KB for a 3x3 64->64 layer.
// The BatchNorm statistics are accumulated in registers across the CTA's tiles (butterfly column sums, fp32 per 32 pixels,
// double across tiles) and flushed with one double atomic per (warp, channel) at the end.
// =====================================================================================================================
constexpr int HT_H = 16, HT_W = 8;      // pixel tile of the halo kernel (M = 128 = 16 groups of 8 pixels)
// warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 and 6-9 TWO epilogue warpgroups: group g drains TMEM accumulator g, i.e.
// every other pixel tile.  One group needs ~3300 cycles per tile (tcgen05.ld -> bias -> staging -> TMA store -> statistics, a
// chain of latencies), the MMA of a tile 1728 (single-pass bf16) / 4032 (split) cycles: with one group the single-pass kernel
// was epilogue-bound at 24 % tensor-pipe activity (ncu r02).
constexpr int HALO_EPI_GROUPS = 2;
constexpr int HALO_THREADS = 64 + HALO_EPI_GROUPS * 128;

struct ConvHaloParams {
    ConvTcParams c;
    int halo_w, halo_h;      // staged box, pixels
    int x_bytes, slot_bytes; // bytes of one staged plane; slot = x_bytes rounded up to 1024
    int nslots, w_bytes;     // x ring depth; resident weight bytes (taps * cchunks * PLANES * BLOCK_N * 128)
    int box_dh, box_dw;      // box origin relative to the tile's first output pixel
    int off_h0, off_w0;      // in-box position of tap (0, 0)
    long long pix_tiles;     // N * tiles_h * tiles_w
    int tma_out;             // 1: the epilogue stages 32 pixels x 16 channels per warp in shared memory and TMA-stores them
    int tma_reduce;          //    (.add form when the addend IS the output buffer: in-place gradient accumulation)
    int out_bufs;            // staging buffers per epilogue warp (1, 2 or 4 x 2 KB): more = fewer waits on the TMA unit's reads
};

// column sums over the 32 lanes of a warp: on return lane l holds sum_over_lanes v[l]  (31 shuffles; v is destroyed)
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            const float send = upper ? v[i] : v[i + off];
            const float keep = upper ? v[i + off] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return v[0];
}

template <int BLOCK_N, bool SPLIT>
__global__ void __launch_bounds__(HALO_THREADS, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap map_x_hi, const __grid_constant__ CUtensorMap map_x_lo,
                 const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
                 const __grid_constant__ CUtensorMap map_z, const ConvHaloParams hp) {
    constexpr int PLANES = SPLIT ? 2 : 1;
    constexpr int WT_BYTES = BLOCK_N * 128;                  // one (tap, k-chunk) weight tile of one plane
    constexpr int ACC_COLS = SPLIT ? 2 * BLOCK_N : BLOCK_N;
    constexpr int TMEM_COLS = 2 * ACC_COLS;
    constexpr int MAX_SLOTS = 8;
    const ConvTcParams& p = hp.c;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    uint8_t* wsm = smem;
    uint8_t* xsm = smem + hp.w_bytes;                        // w_bytes is a multiple of 1024
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(xsm + hp.nslots * hp.slot_bytes);
    uint64_t* empty_bar = full_bar + MAX_SLOTS;
    uint64_t* tmem_full = empty_bar + MAX_SLOTS;
    uint64_t* tmem_empty = tmem_full + 2;
    uint64_t* w_bar = tmem_empty + 2;
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(w_bar + 1);
    uint8_t* out_stage = xsm + hp.nslots * hp.slot_bytes + 1024;     // 8 epilogue warps x out_bufs x 2 KB (32 pixels x 16 fp32 channels)

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int nblk = blockIdx.x % p.n_blocks;
    const long long tile0 = blockIdx.x / p.n_blocks, tile_step = gridDim.x / p.n_blocks;
    const int cchunks = p.Cin_p / BLOCK_K;
    const int ntaps = p.n_r * p.n_s;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_x_hi);
        tma_prefetch_desc(&map_w_hi);
        if (SPLIT) {
            tma_prefetch_desc(&map_x_lo);
            tma_prefetch_desc(&map_w_lo);
        }
        if (hp.tma_out) tma_prefetch_desc(&map_z);
        for (int i = 0; i < hp.nslots; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 128);
        }
        mbar_init(w_bar, 1);
        mbar_fence_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_holder, TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            // resident weights: [tap][k-chunk][plane][BLOCK_N rows x 128 B]
            mbar_expect_tx(w_bar, static_cast<uint32_t>(hp.w_bytes));
            for (int t = 0; t < ntaps; ++t) {
                const int ti = t / p.n_s, tj = t - ti * p.n_s;
                const int tap = (p.w_r0 + ti * p.w_rstep) * p.KW + (p.w_s0 + tj * p.w_sstep);
                for (int cc = 0; cc < cchunks; ++cc) {
                    uint8_t* dst = wsm + static_cast<size_t>(t * cchunks + cc) * PLANES * WT_BYTES;
                    tma_load_3d(dst, &map_w_hi, w_bar, cc * BLOCK_K, nblk * BLOCK_N, tap);
                    if (SPLIT) tma_load_3d(dst + WT_BYTES, &map_w_lo, w_bar, cc * BLOCK_K, nblk * BLOCK_N, tap);
                }
            }
            int slot = 0;
            uint32_t phase = 0;
            for (long long tile = tile0; tile < hp.pix_tiles; tile += tile_step) {
                unsigned t = static_cast<unsigned>(tile);          // 32-bit: four 64-bit divisions per tile were ~9 % of the
                const int tw = static_cast<int>(t % p.tiles_w);    // fast-precision tile period (ncu r02); pix_tiles < 2^31 is
                t /= p.tiles_w;                                    // checked on the host
                const int th = static_cast<int>(t % p.tiles_h);
                const int n = static_cast<int>(t / p.tiles_h);
                const int h0 = th * HT_H + hp.box_dh, w0 = tw * HT_W + hp.box_dw;
                for (int cc = 0; cc < cchunks; ++cc) {
#pragma unroll
                    for (int pl = 0; pl < PLANES; ++pl) {
                        mbar_wait(&empty_bar[slot], phase ^ 1u);
                        mbar_expect_tx(&full_bar[slot], static_cast<uint32_t>(hp.x_bytes));
                        tma_load_4d(xsm + slot * hp.slot_bytes, pl ? &map_x_lo : &map_x_hi, &full_bar[slot], cc * BLOCK_K, w0,
                                    h0, n);
                        if (++slot == hp.nslots) {
                            slot = 0;
                            phase ^= 1u;
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc_bf16(BLOCK_M, BLOCK_N, 0, 0);
            constexpr uint32_t idesc_wide = make_idesc_bf16(BLOCK_M, 2 * BLOCK_N, 0, 0);
            constexpr int KSTEP = (UMMA_K * 2) >> 4;
            const uint64_t dbase_a = desc_base(16, static_cast<uint32_t>(hp.halo_w) * 128u, kSwizzle128);
            const uint64_t dbase_b = desc_base(16, 1024, kSwizzle128) + (smem_u32(wsm) >> 4);
            const uint32_t xs0 = smem_u32(xsm) >> 4;
            mbar_wait(w_bar, 0);
            int slot = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (long long tile = tile0; tile < hp.pix_tiles; tile += tile_step) {
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * ACC_COLS;
                for (int cc = 0; cc < cchunks; ++cc) {
#pragma unroll
                    for (int pl = 0; pl < PLANES; ++pl) {
                        mbar_wait(&full_bar[slot], phase);
                        tc_fence_after();
                        const uint64_t da0 = dbase_a + (xs0 + ((slot * hp.slot_bytes) >> 4));
                        uint32_t toff_row = static_cast<uint32_t>(hp.off_h0 * hp.halo_w + hp.off_w0) * 8u;   // 128 B = 8 units
                        uint32_t woff = static_cast<uint32_t>(cc * PLANES * WT_BYTES) >> 4;
                        for (int ti = 0; ti < p.n_r; ++ti) {
                            uint32_t toff = toff_row;
                            for (int tj = 0; tj < p.n_s; ++tj) {
                                const uint64_t da = da0 + toff;
                                const uint64_t db = dbase_b + woff;
#pragma unroll
                                for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                                    if (pl == 0)
                                        umma_f16(d_tmem, da + k * KSTEP, db + k * KSTEP, SPLIT ? idesc_wide : idesc,
                                                 (cc | ti | tj | k) != 0 ? 1u : 0u);
                                    else
                                        umma_f16(d_tmem, da + k * KSTEP, db + k * KSTEP, idesc, 1u);
                                }
                                toff += static_cast<uint32_t>(p.dw_step * 8);
                                woff += static_cast<uint32_t>(cchunks * PLANES * WT_BYTES) >> 4;
                            }
                            toff_row += static_cast<uint32_t>(p.dh_step * hp.halo_w * 8);
                        }
                        umma_commit(&empty_bar[slot]);      // frees the plane slot when these MMAs retire
                        if (++slot == hp.nslots) {
                            slot = 0;
                            phase ^= 1u;
                        }
                    }
                }
                umma_commit(&tmem_full[acc]);
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1u;
                }
            }
        }
    } else {
        // ===================== epilogue: TMEM -> registers -> global =====================
        const int q = warp & 3;  // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;
        const int lh = row / HT_W, lw = row - lh * HT_W;
        const int egroup = (warp - EPI_WARP0) >> 2;     // this warp's epilogue group = the TMEM accumulator it drains
        const int acc = egroup;
        uint32_t acc_phase = 0;
        int it = 0;
        // per-lane statistics accumulators: [32-column chunk] in the per-thread-store epilogue (lane = column),
        // [16-column half] in the TMA-store epilogue (lane % 16 = column)
        // Compensated (Kahan) fp32 pairs, not doubles: the F2F.F64 + DADD per half sat on the FP64 pipe with its long latency
        // in the middle of the epilogue's dependency chain (ncu r02: 13-18 % of the kernel's stall samples were
        // math_pipe_throttle on exactly those DADDs).  The pair carries ~48 bits, converted to double once, at the end.
        KahanF st_sum[BLOCK_N / 16], st_sq[BLOCK_N / 16];
        const float* bias = p.bias ? p.bias + nblk * BLOCK_N : nullptr;
        int obuf = 0;
        for (long long tile = tile0; tile < hp.pix_tiles; tile += tile_step, ++it) {
            if ((it & 1) != egroup) continue;               // the other group's tile
            unsigned t = static_cast<unsigned>(tile);
            const int tw = static_cast<int>(t % p.tiles_w);
            t /= p.tiles_w;
            const int th = static_cast<int>(t % p.tiles_h);
            const int n = static_cast<int>(t / p.tiles_h);
            const int oh_l = th * HT_H + lh, ow_l = tw * HT_W + lw;
            const bool valid = (oh_l < p.TOH) && (ow_l < p.TOW);
            const int oh = oh_l * p.out_sh + p.out_oh, ow = ow_l * p.out_sw + p.out_ow;
            const size_t pix = static_cast<size_t>(n) * p.OH * p.OW + static_cast<size_t>(oh) * p.OW + ow;
            float* zrow = p.z + pix * p.z_ld + nblk * BLOCK_N;
            const float* arow = p.addend ? p.addend + pix * p.addend_ld + nblk * BLOCK_N : nullptr;
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            if (hp.tma_out) {
                // ---- TMA-store epilogue: the accumulator rows of this warp are 4 image rows x 8 pixels; 16 channels at a time
                // go through a 2 KB staging buffer (64B-swizzled rows, conflict-free 16-byte writes) and leave as ONE bulk
                // tensor store {16 ch, 8 w, 4 h}: full 64-byte segments instead of 32 scattered 16-byte stores per instruction,
                // out-of-image pixels clipped by the TMA unit, in-place accumulation as a reduce-add.
                uint8_t* stage_ring = out_stage + (egroup * 4 + q) * hp.out_bufs * 2048;
                const int oh0 = th * HT_H + q * 4, ow0 = tw * HT_W;
#pragma unroll
                for (int ci = 0; ci < BLOCK_N / 32; ++ci) {
                    const int c = ci * 32;
                    uint32_t v[32];
                    tmem_ld_32x32(tmem_base + acc * ACC_COLS + c + (static_cast<uint32_t>(q * 32) << 16), v);
                    tmem_ld_wait();
                    float f[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                    if (SPLIT) {   // add the x_hi * w_lo half of the stacked accumulator
                        tmem_ld_32x32(tmem_base + acc * ACC_COLS + BLOCK_N + c + (static_cast<uint32_t>(q * 32) << 16), v);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] += __uint_as_float(v[j]);
                    }
                    if (ci == BLOCK_N / 32 - 1) {      // every accumulator column is in registers: hand the buffer back
                        tc_fence_before();
                        mbar_arrive(&tmem_empty[acc]);
                    }
                    if (bias) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] += __ldg(bias + c + j);
                    }
                    if (!valid) {       // out-of-image pixels: clipped by the TMA store, and they must not enter the statistics
#pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] = 0.f;
                    }
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        uint8_t* stage = stage_ring + obuf * 2048;
                        if (++obuf == hp.out_bufs) obuf = 0;
                        // the store that last used THIS buffer has finished reading it (out_bufs - 1 younger ones may be in flight)
                        if (lane == 0) tma_store_wait_read_pending(hp.out_bufs - 1);
                        __syncwarp();
#pragma unroll
                        for (int qd = 0; qd < 4; ++qd) {
                            const int j = half * 16 + qd * 4;
                            *reinterpret_cast<float4*>(stage + lane * 64 + ((qd ^ ((lane >> 1) & 3)) << 4)) =
                                make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
                        }
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            const int c0 = nblk * BLOCK_N + c + half * 16;
                            if (hp.tma_reduce) tma_reduce_add_4d(&map_z, stage, c0, ow0, oh0, n);
                            else tma_store_4d(&map_z, stage, c0, ow0, oh0, n);
                            tma_store_commit();
                        }
                        if (p.stat_sum) {
                            // BatchNorm statistics from the staged tile: lane -> (column lane % 16, rows of parity lane / 16);
                            // two adjacent 64-byte rows cover all 32 banks, so every LDS is conflict-free.
                            const int col = lane & 15, rpar = lane >> 4;
                            const int cq = col >> 2, cw = (col & 3) << 2;
                            float a = 0.f, b = 0.f;
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                const int r = 2 * i + rpar;
                                const float v1 = *reinterpret_cast<const float*>(stage + r * 64 + ((cq ^ ((r >> 1) & 3)) << 4) + cw);
                                a += v1;
                                b = fmaf(v1, v1, b);
                            }
                            a += __shfl_xor_sync(0xffffffffu, a, 16);
                            b += __shfl_xor_sync(0xffffffffu, b, 16);
                            st_sum[ci * 2 + half].add(a);
                            st_sq[ci * 2 + half].add(b);
                        }
                    }
                }
                acc_phase ^= 1u;
                continue;
            }
#pragma unroll
            for (int ci = 0; ci < BLOCK_N / 32; ++ci) {
                const int c = ci * 32;
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + acc * ACC_COLS + c + (static_cast<uint32_t>(q * 32) << 16), v);
                tmem_ld_wait();
                float f[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                if (SPLIT) {   // add the x_hi * w_lo half of the stacked accumulator
                    tmem_ld_32x32(tmem_base + acc * ACC_COLS + BLOCK_N + c + (static_cast<uint32_t>(q * 32) << 16), v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] += __uint_as_float(v[j]);
                }
                if (bias) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] += __ldg(bias + c + j);
                }
                if (valid) {
                    if (arow) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 a = *reinterpret_cast<const float4*>(arow + c + j);
                            f[j] += a.x; f[j + 1] += a.y; f[j + 2] += a.z; f[j + 3] += a.w;
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<float4*>(zrow + c + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
                }
                if (p.stat_sum) {
                    float sq[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        f[j] = valid ? f[j] : 0.f;
                        sq[j] = f[j] * f[j];
                    }
                    st_sum[ci].add(warp_colsum32(f, lane));
                    st_sq[ci].add(warp_colsum32(sq, lane));
                }
            }
            tc_fence_before();
            mbar_arrive(&tmem_empty[acc]);
            acc_phase ^= 1u;
        }
        if (hp.tma_out && lane == 0) tma_store_wait_all();
        if (p.stat_sum && hp.tma_out) {
            if (lane < 16) {
#pragma unroll
                for (int h16 = 0; h16 < BLOCK_N / 16; ++h16) {
                    atomicAdd(p.stat_sum + nblk * BLOCK_N + h16 * 16 + lane, st_sum[h16].value());
                    atomicAdd(p.stat_sqsum + nblk * BLOCK_N + h16 * 16 + lane, st_sq[h16].value());
                }
            }
        } else if (p.stat_sum) {
#pragma unroll
            for (int ci = 0; ci < BLOCK_N / 32; ++ci) {
                atomicAdd(p.stat_sum + nblk * BLOCK_N + ci * 32 + lane, st_sum[ci].value());
                atomicAdd(p.stat_sqsum + nblk * BLOCK_N + ci * 32 + lane, st_sq[ci].value());
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ---- host side ------------------------------------------------------------------------------
PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
    }
    return fn;
}

}  // namespace

// 4D NHWC bf16 activation map: dims {C, W, H, N}, box {64, TILE_W, TILE_H, 1}, 128B swizzle, zero fill.
int make_act_tmap(CUtensorMap* m, const void* base, int C, int W, int H, int N, int ld, int box_c, int box_w,
                  int box_h, int estride_w, int estride_h) {
    auto enc = get_encode_fn();
    if (!enc) {
        set_error(FCD_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
        return FCD_ERR_CUDA;
    }
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)W * ld * 2, (cuuint64_t)H * W * ld * 2};
    // with elementStrides e the box spans box*e coordinates and transfers ceil(box*e / e) = box elements per dim
    cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)(box_w * estride_w), (cuuint32_t)(box_h * estride_h), 1};
    cuuint32_t estr[4] = {1, (cuuint32_t)estride_w, (cuuint32_t)estride_h, 1};
    CUtensorMapSwizzle sw = box_c * 2 >= 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                            : box_c * 2 == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                              : CU_TENSOR_MAP_SWIZZLE_32B;
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error(FCD_ERR_CUDA, "cuTensorMapEncodeTiled(act) failed: %d (C=%d W=%d H=%d N=%d ld=%d)", (int)r, C, W, H,
                  N, ld);
        return FCD_ERR_CUDA;
    }
    return FCD_OK;
}

// 3D packed weight map: dims {cols, rows, taps}, box {64, box_rows, 1}.
int make_wgt_tmap(CUtensorMap* m, const void* base, int cols, int rows, int taps, int box_cols, int box_rows) {
    auto enc = get_encode_fn();
    if (!enc) {
        set_error(FCD_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
        return FCD_ERR_CUDA;
    }
    cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)taps};
    cuuint64_t strides[2] = {(cuuint64_t)cols * 2, (cuuint64_t)rows * cols * 2};
    cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUtensorMapSwizzle sw = box_cols * 2 >= 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                            : box_cols * 2 == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                 : CU_TENSOR_MAP_SWIZZLE_32B;
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error(FCD_ERR_CUDA, "cuTensorMapEncodeTiled(weight) failed: %d (cols=%d rows=%d taps=%d)", (int)r, cols,
                  rows, taps);
        return FCD_ERR_CUDA;
    }
    return FCD_OK;
}

// 4D NHWC fp32 output map for the TMA-store epilogue: dims {C, W, H, N}, box {16 ch, box_w, box_h, 1}, 64B swizzle.
int make_out_tmap(CUtensorMap* m, float* base, int C, int W, int H, int N, int ld, int box_w, int box_h) {
    auto enc = get_encode_fn();
    if (!enc) {
        set_error(FCD_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
        return FCD_ERR_CUDA;
    }
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)ld * 4, (cuuint64_t)W * ld * 4, (cuuint64_t)H * W * ld * 4};
    cuuint32_t box[4] = {16, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error(FCD_ERR_CUDA, "cuTensorMapEncodeTiled(out) failed: %d (C=%d W=%d H=%d N=%d ld=%d)", (int)r, C, W, H, N, ld);
        return FCD_ERR_CUDA;
    }
    return FCD_OK;
}

bool conv_tc_supported(int Cin_p, int Cout_p, int KH, int KW, int stride) {
    return (stride == 1 || stride == 2) && Cin_p % 64 == 0 && Cout_p % 64 == 0 && KH >= 1 && KW >= 1 && KH <= 9 && KW <= 9;
}


bool g_conv_halo_enabled = true;    // fcd_set_option("conv_halo", 0) selects the per-tap kernel everywhere
bool g_conv_tma_out = true;         // fcd_set_option("conv_tma_out", 0): per-thread stores in the halo kernel's epilogue
int g_conv_halo_slots = 0;          // fcd_set_option("conv_halo_slots", n): cap of the x-plane ring (0 = default); what the ring
                                    // does not take deepens the epilogue's output staging ring

namespace {
constexpr int HALO_SMEM_MAX = 227 * 1024 - 1024;   // opt-in limit minus the kernel's static shared memory
}

template <int BLOCK_N, bool SPLIT>
static int launch_conv_halo(const CUtensorMap& mxh, const CUtensorMap& mxl, const CUtensorMap& mwh, const CUtensorMap& mwl,
                            const CUtensorMap& mz, const ConvHaloParams& hp, int smem_bytes, cudaStream_t stream) {
    static PerDeviceOnce attr_once;
    if (!attr_once()) {
        FCD_CUDA_OK(cudaFuncSetAttribute(conv_halo_kernel<BLOCK_N, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         HALO_SMEM_MAX));
        attr_once() = true;
    }
    long long grid = sm_count() / hp.c.n_blocks * hp.c.n_blocks;
    if (grid < hp.c.n_blocks) grid = hp.c.n_blocks;
    if (grid > hp.pix_tiles * hp.c.n_blocks) grid = hp.pix_tiles * hp.c.n_blocks;
    conv_halo_kernel<BLOCK_N, SPLIT><<<(unsigned)grid, HALO_THREADS, smem_bytes, stream>>>(mxh, mxl, mwh, mwl, mz, hp);
    FCD_LAUNCH_OK();
    return FCD_OK;
}

// Fills hp and returns true when the halo kernel takes this launch: stride-1 reads, every tap view inside one TMA box, and the
// CTA's weight block resident in shared memory next to >= 2 plane slots and the two epilogue groups' staging rings.
static bool conv_halo_plan(const ConvTcParams& p, bool split, int csh, int csw, int* block_n_out, ConvHaloParams* hp,
                           int* smem_bytes) {
    if (!g_conv_halo_enabled || csh != 1 || csw != 1 || p.n_r * p.n_s < 1) return false;
    const int span_h = (p.n_r - 1) * (p.dh_step < 0 ? -p.dh_step : p.dh_step);
    const int span_w = (p.n_s - 1) * (p.dw_step < 0 ? -p.dw_step : p.dw_step);
    const int halo_h = HT_H + span_h, halo_w = HT_W + span_w;
    if (halo_h > 256 || halo_w > 256) return false;
    const int x_bytes = halo_h * halo_w * 128;
    const int slot = (x_bytes + 1023) & ~1023;
    const int planes = split ? 2 : 1;
    const int cchunks = p.Cin_p / BLOCK_K;
    const int min_slots = 2;      // hi / lo planes alternate: one being multiplied, one in flight (a third slot measured no gain)
    int block_n = 0, nslots = 0, w_bytes = 0;
    for (int bn : {128, 64}) {
        if (p.Cout_p % bn) continue;
        if (split && bn == 128) continue;                      // 2 * (2 * 128) accumulator columns would need all of TMEM
        const long long wb = 1LL * p.n_r * p.n_s * cchunks * planes * bn * 128;
        const long long room = HALO_SMEM_MAX - 1024 - 1024 - HALO_EPI_GROUPS * 4 * 2048 - wb;    // alignment slack, barriers, minimal output staging
        if (room < 1LL * min_slots * slot) continue;
        block_n = bn; w_bytes = static_cast<int>(wb);
        nslots = static_cast<int>(room / slot);
        if (nslots > 4) nslots = 4;        // a one-tile look-ahead is all the MMA warp can use; the rest goes to the epilogue
        if (g_conv_halo_slots >= 2 && nslots > g_conv_halo_slots) nslots = g_conv_halo_slots;
        break;
    }
    if (!block_n) return false;
    hp->c = p;
    hp->halo_w = halo_w; hp->halo_h = halo_h; hp->x_bytes = x_bytes; hp->slot_bytes = slot; hp->nslots = nslots;
    hp->w_bytes = w_bytes;
    const int min_dh = p.dh_step < 0 ? (p.n_r - 1) * p.dh_step : 0, min_dw = p.dw_step < 0 ? (p.n_s - 1) * p.dw_step : 0;
    hp->box_dh = p.dh0 + min_dh; hp->box_dw = p.dw0 + min_dw;
    hp->off_h0 = -min_dh; hp->off_w0 = -min_dw;
    *block_n_out = block_n;
    // whatever is left after the weights and the x ring deepens the epilogue's staging ring (8 warps x out_bufs x 2 KB)
    const long long left = HALO_SMEM_MAX - 1024 - 1024 - w_bytes - 1LL * nslots * slot;
    constexpr int EW = HALO_EPI_GROUPS * 4;
    hp->out_bufs = left >= 4 * EW * 2048 ? 4 : left >= 2 * EW * 2048 ? 2 : 1;
    *smem_bytes = w_bytes + nslots * slot + 1024 + 1024 + EW * hp->out_bufs * 2048;
    return true;
}

template <int BLOCK_N, bool SPLIT>
static int launch_conv_tc(const CUtensorMap& mxh, const CUtensorMap& mxl, const CUtensorMap& mwh,
                          const CUtensorMap& mwl, const ConvTcParams& p, cudaStream_t stream) {
    using C = Cfg<BLOCK_N, SPLIT>;
    static PerDeviceOnce attr_once;
    if (!attr_once()) {
        FCD_CUDA_OK(cudaFuncSetAttribute(conv_tc_kernel<BLOCK_N, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         C::SMEM_BYTES));
        attr_once() = true;
    }
    long long grid = p.total_tiles < sm_count() ? p.total_tiles : sm_count();
    conv_tc_kernel<BLOCK_N, SPLIT><<<(unsigned)grid, NUM_THREADS, C::SMEM_BYTES, stream>>>(mxh, mxl, mwh, mwl, p);
    FCD_LAUNCH_OK();
    return FCD_OK;
}

static int run_conv_tc(const void* x_hi, const void* x_lo, int x_ld, int XH, int XW, const void* w_hi, const void* w_lo,
                       int w_rows, int w_cols, int w_taps, ConvTcParams p, int N, int csh, int csw, cudaStream_t stream) {
    const bool split = (x_lo != nullptr) && (w_lo != nullptr);
    CUtensorMap mxh, mxl, mwh, mwl;
    int rc;
    {
        ConvHaloParams hp;
        int hbn = 0, hsmem = 0;
        if (conv_halo_plan(p, split, csh, csw, &hbn, &hp, &hsmem)) {
            if ((rc = make_act_tmap(&mxh, x_hi, p.Cin_p, XW, XH, N, x_ld, BLOCK_K, hp.halo_w, hp.halo_h, 1, 1))) return rc;
            if ((rc = make_wgt_tmap(&mwh, w_hi, w_cols, w_rows, w_taps, BLOCK_K, hbn))) return rc;
            if (split) {
                if ((rc = make_act_tmap(&mxl, x_lo, p.Cin_p, XW, XH, N, x_ld, BLOCK_K, hp.halo_w, hp.halo_h, 1, 1))) return rc;
                if ((rc = make_wgt_tmap(&mwl, w_lo, w_cols, w_rows, w_taps, BLOCK_K, hbn))) return rc;
            } else {
                mxl = mxh;
                mwl = mwh;
            }
            hp.c.N = N; hp.c.csh = 1; hp.c.csw = 1;
            hp.c.tiles_h = ceil_div(p.TOH, HT_H);
            hp.c.tiles_w = ceil_div(p.TOW, HT_W);
            hp.c.n_blocks = p.Cout_p / hbn;
            hp.pix_tiles = 1LL * N * hp.c.tiles_h * hp.c.tiles_w;
            FCD_CHECK_ARG(hp.pix_tiles < (1LL << 31), "conv2d (halo kernel): more than 2^31 pixel tiles");
            hp.c.total_tiles = hp.pix_tiles * hp.c.n_blocks;
            // TMA-store epilogue: dense output pixels, 16-byte aligned rows; the addend (if any) must be the output itself
            CUtensorMap mz = mxh;
            hp.tma_out = g_conv_tma_out && p.out_sh == 1 && p.out_sw == 1 && p.out_oh == 0 && p.out_ow == 0 &&
                         (reinterpret_cast<uintptr_t>(p.z) & 15) == 0 && p.z_ld % 4 == 0 &&
                         (p.addend == nullptr || (p.addend == p.z && p.addend_ld == p.z_ld));
            hp.tma_reduce = hp.tma_out && p.addend != nullptr;
            if (hp.tma_out && (rc = make_out_tmap(&mz, p.z, p.Cout_p, p.OW, p.OH, N, p.z_ld, HT_W, 4))) return rc;
            if (hbn == 128) return launch_conv_halo<128, false>(mxh, mxl, mwh, mwl, mz, hp, hsmem, stream);
            return split ? launch_conv_halo<64, true>(mxh, mxl, mwh, mwl, mz, hp, hsmem, stream)
                         : launch_conv_halo<64, false>(mxh, mxl, mwh, mwl, mz, hp, hsmem, stream);
        }
    }
    const int block_n = (p.Cout_p % 128 == 0) ? 128 : 64;
    // the per-tap kernel's in-epilogue statistics cost two double atomics per (warp, channel, tile): run the dedicated
    // reduction kernel after it instead
    double* stat_sum = p.stat_sum;
    double* stat_sqsum = p.stat_sqsum;
    p.stat_sum = nullptr;
    p.stat_sqsum = nullptr;
    if ((rc = make_act_tmap(&mxh, x_hi, p.Cin_p, XW, XH, N, x_ld, BLOCK_K, TILE_W, TILE_H, csw, csh))) return rc;
    if ((rc = make_wgt_tmap(&mwh, w_hi, w_cols, w_rows, w_taps, BLOCK_K, block_n))) return rc;
    if (split) {
        if ((rc = make_act_tmap(&mxl, x_lo, p.Cin_p, XW, XH, N, x_ld, BLOCK_K, TILE_W, TILE_H, csw, csh))) return rc;
        if ((rc = make_wgt_tmap(&mwl, w_lo, w_cols, w_rows, w_taps, BLOCK_K, block_n))) return rc;
    } else {
        mxl = mxh;
        mwl = mwh;
    }
    p.N = N;
    p.csh = csh;
    p.csw = csw;
    p.tiles_h = ceil_div(p.TOH, TILE_H);
    p.tiles_w = ceil_div(p.TOW, TILE_W);
    p.n_blocks = p.Cout_p / block_n;
    p.total_tiles = 1LL * N * p.tiles_h * p.tiles_w * p.n_blocks;
    if (block_n == 128)
        rc = split ? launch_conv_tc<128, true>(mxh, mxl, mwh, mwl, p, stream)
                   : launch_conv_tc<128, false>(mxh, mxl, mwh, mwl, p, stream);
    else
        rc = split ? launch_conv_tc<64, true>(mxh, mxl, mwh, mwl, p, stream)
                   : launch_conv_tc<64, false>(mxh, mxl, mwh, mwl, p, stream);
    if (rc) return rc;
    if (stat_sum)
        return fcd_bn_stats(p.z, p.z_ld, 1LL * N * p.OH * p.OW, p.Cout_p, stat_sum, stat_sqsum, stream);
    return FCD_OK;
}

int conv2d_fwd_tc(const void* x_hi, const void* x_lo, int x_ld, const void* w_hi, const void* w_lo,
                  const float* bias, const float* addend, int addend_ld, float* z, int z_ld, int N, int H, int W,
                  int Cin_p, int Cout_p, int KH, int KW, int stride, int pad, double* stat_sum, double* stat_sqsum,
                  cudaStream_t stream) {
    const int OH = (H + 2 * pad - KH) / stride + 1, OW = (W + 2 * pad - KW) / stride + 1;
    FCD_CHECK_ARG(OH > 0 && OW > 0, "conv2d_fwd_tc: empty output");
    FCD_CHECK_ARG(z_ld % 4 == 0 && x_ld % 8 == 0 && addend_ld % 4 == 0, "conv2d_fwd_tc: pitches must keep 16-byte alignment");
    ConvTcParams p{};
    p.bias = bias; p.addend = addend; p.addend_ld = addend_ld; p.z = z;
    p.stat_sum = stat_sum; p.stat_sqsum = stat_sqsum; p.z_ld = z_ld;
    p.OH = OH; p.OW = OW; p.TOH = OH; p.TOW = OW;
    p.out_sh = 1; p.out_oh = 0; p.out_sw = 1; p.out_ow = 0;
    p.Cin_p = Cin_p; p.Cout_p = Cout_p;
    p.n_r = KH; p.n_s = KW;
    p.dh0 = -pad; p.dh_step = 1; p.dw0 = -pad; p.dw_step = 1;
    p.w_r0 = 0; p.w_rstep = 1; p.w_s0 = 0; p.w_sstep = 1; p.KW = KW;
    return run_conv_tc(x_hi, x_lo, x_ld, H, W, w_hi, w_lo, Cout_p, Cin_p, KH * KW, p, N, stride, stride, stream);
}

// Generic "tap list" convolution: z[n,oh,ow,:] = bias + addend + sum_{i<n_r, j<n_s} x[n, oh*csh + dh0 + i*dh_step,
// ow*csw + dw0 + j*dw_step, :] . w[i*n_s + j]      (x out of bounds = 0).
// Serves the 4-pixel channel-packed forms of the 13-band convolutions (engine.py: conv_small_in / conv_small_out).
int conv2d_taps_tc(const void* x_hi, const void* x_lo, int x_ld, int XH, int XW, const void* w_hi, const void* w_lo,
                   const float* bias, const float* addend, int addend_ld, float* z, int z_ld, int N, int OH, int OW,
                   int Cin_p, int Cout_p, int n_r, int n_s, int dh0, int dh_step, int dw0, int dw_step, int csh, int csw,
                   cudaStream_t stream) {
    FCD_CHECK_ARG(OH > 0 && OW > 0 && n_r > 0 && n_s > 0 && csh >= 1 && csw >= 1, "conv2d_taps_tc: bad geometry");
    FCD_CHECK_ARG(Cin_p % 64 == 0 && Cout_p % 64 == 0, "conv2d_taps_tc: channel counts must be multiples of 64");
    FCD_CHECK_ARG(z_ld % 4 == 0 && x_ld % 8 == 0 && addend_ld % 4 == 0, "conv2d_taps_tc: pitches must keep 16-byte alignment");
    FCD_CHECK_ARG(TILE_W * csw <= 256 && TILE_H * csh <= 256, "conv2d_taps_tc: coordinate scale too large for a TMA box");
    ConvTcParams p{};
    p.bias = bias; p.addend = addend; p.addend_ld = addend_ld; p.z = z;
    p.stat_sum = nullptr; p.stat_sqsum = nullptr; p.z_ld = z_ld;
    p.OH = OH; p.OW = OW; p.TOH = OH; p.TOW = OW;
    p.out_sh = 1; p.out_oh = 0; p.out_sw = 1; p.out_ow = 0;
    p.Cin_p = Cin_p; p.Cout_p = Cout_p;
    p.n_r = n_r; p.n_s = n_s;
    p.dh0 = dh0; p.dh_step = dh_step; p.dw0 = dw0; p.dw_step = dw_step;
    p.w_r0 = 0; p.w_rstep = 1; p.w_s0 = 0; p.w_sstep = 1; p.KW = n_s;
    return run_conv_tc(x_hi, x_lo, x_ld, XH, XW, w_hi, w_lo, Cout_p, Cin_p, n_r * n_s, p, N, csh, csw, stream);
}

// Stride-2 dgrad on the tcgen05 engine: the input-gradient pixels split into stride*stride parity classes; each class
// is a stride-1 implicit GEMM over dz with the subset of taps whose parity matches, scattered to every stride-th pixel.
//   dx[n, h, w, ci] = sum_{r,s: (h+pad-r), (w+pad-s) divisible by stride} dz[n, (h+pad-r)/stride, (w+pad-s)/stride, :] . w[r,s][:, ci]
// wT are MODE-1 packed weights: wT[(KH-1-r)*KW + (KW-1-s)][ci][co].
int conv2d_dgrad_strided_tc(const void* dz_hi, const void* dz_lo, int dz_ld, const void* wT_hi, const void* wT_lo,
                            const float* addend, int addend_ld, float* dx, int dx_ld, int N, int H, int W, int Cin_p,
                            int Cout_p, int KH, int KW, int stride, int pad, cudaStream_t stream) {
    const int OH = (H + 2 * pad - KH) / stride + 1, OW = (W + 2 * pad - KW) / stride + 1;   // dz dims
    FCD_CHECK_ARG(dx_ld % 4 == 0 && dz_ld % 8 == 0 && addend_ld % 4 == 0, "conv2d_dgrad_strided_tc: pitches");
    for (int ph = 0; ph < stride; ++ph) {
        for (int pw = 0; pw < stride; ++pw) {
            const int r0 = (ph + pad) % stride, s0 = (pw + pad) % stride;
            const int n_r = r0 < KH ? (KH - r0 + stride - 1) / stride : 0;
            const int n_s = s0 < KW ? (KW - s0 + stride - 1) / stride : 0;
            const int TOH = (H - ph + stride - 1) / stride, TOW = (W - pw + stride - 1) / stride;
            if (TOH <= 0 || TOW <= 0) continue;
            if (n_r == 0 || n_s == 0) {
                set_error(FCD_ERR_UNSUPPORTED, "conv2d_dgrad_strided_tc: a parity class has no taps (K=%dx%d stride %d)", KH, KW, stride);
                return FCD_ERR_UNSUPPORTED;
            }
            ConvTcParams p{};
            p.bias = nullptr; p.addend = addend; p.addend_ld = addend_ld; p.z = dx;
            p.stat_sum = nullptr; p.stat_sqsum = nullptr; p.z_ld = dx_ld;
            p.OH = H; p.OW = W; p.TOH = TOH; p.TOW = TOW;
            p.out_sh = stride; p.out_oh = ph; p.out_sw = stride; p.out_ow = pw;
            p.Cin_p = Cout_p;   // K side = dz channels
            p.Cout_p = Cin_p;   // N side = dx channels
            p.n_r = n_r; p.n_s = n_s;
            // tap i <-> r = r0 + i*stride reads dz row  (h + pad - r)/stride = oh_l + (ph + pad - r0)/stride - i
            p.dh0 = (ph + pad - r0) / stride; p.dh_step = -1;
            p.dw0 = (pw + pad - s0) / stride; p.dw_step = -1;
            p.w_r0 = KH - 1 - r0; p.w_rstep = -stride; p.w_s0 = KW - 1 - s0; p.w_sstep = -stride; p.KW = KW;
            int rc = run_conv_tc(dz_hi, dz_lo, dz_ld, OH, OW, wT_hi, wT_lo, Cin_p, Cout_p, KH * KW, p, N, 1, 1, stream);
            if (rc) return rc;
        }
    }
    return FCD_OK;
}

}  // namespace fcd
