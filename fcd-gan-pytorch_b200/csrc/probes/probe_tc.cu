// probe_tc.cu — single-CTA tcgen05 probe used by scripts/probe_umma.py and tests/test_umma_probe.py.
// It fills shared memory with the 128B-swizzle pattern TMA would produce, issues UMMAs with a
// caller-chosen start-address shift / base_offset / major-ness, and writes the fp32 accumulator back,
// so the descriptor semantics the convolution kernels rely on (row-shifted "halo" views of one staged
// tile; MN-major operands for wgrad) are pinned by measurement rather than by reading of the ISA text.
#include "../fcd_common.cuh"
#include "../fcd_tc.cuh"
#include "../../../include/fcd_b200_probes.h"

namespace fcd {
using namespace tc;
namespace {

// smem image: rows of 128 bytes (64 bf16), 16-byte chunk c of row r stored at chunk (c ^ (r & 7)).
__device__ void fill_swizzled(uint8_t* dst, const __nv_bfloat16* src, int rows) {
    for (int i = threadIdx.x; i < rows * 8; i += blockDim.x) {
        const int r = i >> 3, c = i & 7;
        const uint4 v = *reinterpret_cast<const uint4*>(src + static_cast<size_t>(r) * 64 + c * 8);
        *reinterpret_cast<uint4*>(dst + r * 128 + ((c ^ (r & 7)) << 4)) = v;
    }
}

struct ProbeParams {
    const __nv_bfloat16* a;  // [a_blocks][a_rows][64]
    const __nv_bfloat16* b;  // [b_blocks][b_rows][64]
    float* d;                // [128][n]
    int a_rows, b_rows, a_blocks, b_blocks;
    int mn_major;            // 0: K-major A and B; 1: MN-major A and B
    int n;                   // UMMA N (multiple of 16, <= 256)
    int ksteps;              // number of K=16 UMMAs
    int a_shift_rows, a_base_offset;
    int b_shift_rows, b_base_offset;
    int a_sbo, b_sbo;        // stride-byte-offset overrides (0 = 1024)
    int a_lbo, a_kstep;      // leading-byte-offset / per-UMMA start advance overrides for A (0 = defaults)
};

__global__ void __launch_bounds__(128, 1) umma_probe_kernel(ProbeParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_holder;
    const int warp = threadIdx.x >> 5;

    uint8_t* sa = smem;
    const int a_block_bytes = p.a_rows * 128;
    uint8_t* sb = sa + ((p.a_blocks * a_block_bytes + 1023) & ~1023);
    const int b_block_bytes = p.b_rows * 128;
    for (int blk = 0; blk < p.a_blocks; ++blk)
        fill_swizzled(sa + blk * a_block_bytes, p.a + static_cast<size_t>(blk) * p.a_rows * 64, p.a_rows);
    for (int blk = 0; blk < p.b_blocks; ++blk)
        fill_swizzled(sb + blk * b_block_bytes, p.b + static_cast<size_t>(blk) * p.b_rows * 64, p.b_rows);
    // generic-proxy writes must be visible to the async (tensor-core) proxy
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");

    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    if (warp == 0) {
        tmem_alloc(&tmem_holder, 256);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_holder;

    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_bf16(128, p.n, p.mn_major, p.mn_major);
        const uint32_t a0 = smem_u32(sa) + p.a_shift_rows * 128;
        const uint32_t b0 = smem_u32(sb) + p.b_shift_rows * 128;
        const uint32_t a_sbo = p.a_sbo ? p.a_sbo : 1024, b_sbo = p.b_sbo ? p.b_sbo : 1024;
        for (int k = 0; k < p.ksteps; ++k) {
            uint64_t da, db;
            if (!p.mn_major) {
                // K-major: 16 bf16 = 32 bytes further along the 128-byte row
                da = make_smem_desc(a0 + k * (p.a_kstep ? p.a_kstep : 32), 16, a_sbo, kSwizzle128, p.a_base_offset);
                db = make_smem_desc(b0 + k * 32, 16, b_sbo, kSwizzle128, p.b_base_offset);
            } else {
                // MN-major: K runs over rows; 16 rows = 2048 bytes; LBO = distance between 64-wide MN blocks
                da = make_smem_desc(a0 + k * (p.a_kstep ? p.a_kstep : 2048), p.a_lbo ? p.a_lbo : a_block_bytes, a_sbo,
                                    kSwizzle128, p.a_base_offset);
                db = make_smem_desc(b0 + k * 2048, b_block_bytes, b_sbo, kSwizzle128, p.b_base_offset);
            }
            umma_f16(tmem_base, da, db, idesc, k != 0 ? 1u : 0u);
        }
        umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    tc_fence_after();

    const int row = warp * 32 + (threadIdx.x & 31);
    for (int c = 0; c < p.n; c += 16) {
        uint32_t v[16];
        tmem_ld_32x16(tmem_base + c + (static_cast<uint32_t>(warp * 32) << 16), v);
        tmem_ld_wait();
        for (int j = 0; j < 16; ++j) p.d[static_cast<size_t>(row) * p.n + c + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}


// ---- UMMA issue-rate micro-benchmark (scripts/umma_bench.py): how many cycles does a stream of tcgen05.mma with a given
// shape / operand major-ness / descriptor geometry take when the operands are already in shared memory?
struct BenchParams {
    int mn_major, n1, n2;        // n2 == 0: one UMMA per k-step; else a second one (N = n2) with A at +a2_off
    int a_sbo, a_lbo, a_kstep, a_shift, a2_off;
    int b_sbo, b_lbo, b_kstep;
    int stage_stride, stages, b_off;
    int iters;
    int a_tmem;                  // 1: A operand of the first UMMA comes from TMEM (columns 256..) instead of shared memory
    long long* cycles;
};

__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n"
        :
        : "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__global__ void __launch_bounds__(128, 1) umma_bench_kernel(BenchParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_holder;
    const int warp = threadIdx.x >> 5;
    const int total = p.stage_stride * p.stages;
    for (int i = threadIdx.x; i < total / 4; i += blockDim.x)
        reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + (i & 0xff);       // small finite bf16 pairs
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    if (warp == 0) {
        tmem_alloc(&tmem_holder, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_holder;
    if (warp == 0 && elect_one()) {
        const uint32_t idesc1 = make_idesc_bf16(128, p.n1, p.mn_major, p.mn_major);
        const uint32_t idesc2 = make_idesc_bf16(128, p.n2 ? p.n2 : 64, p.mn_major, p.mn_major);
        const uint32_t base = smem_u32(smem);
        // descriptors are built ONCE; the loop only adds 16-byte-unit offsets to the low word (address field)
        const uint64_t da0 = make_smem_desc(base + p.a_shift * 128, p.a_lbo, p.a_sbo, kSwizzle128);
        const uint64_t db0 = make_smem_desc(base + p.b_off, p.b_lbo, p.b_sbo, kSwizzle128);
        const uint32_t a_k = p.a_kstep >> 4, b_k = p.b_kstep >> 4, a2 = p.a2_off >> 4, st = p.stage_stride >> 4;
        const bool two = p.n2 != 0, ts = p.a_tmem != 0;
        const long long t0 = clock64();
        int stage = 0;
        for (int it = 0; it < p.iters; ++it) {
            const uint64_t da = da0 + static_cast<uint64_t>(stage * st);
            const uint64_t db = db0 + static_cast<uint64_t>(stage * st);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (ts) umma_f16_ts(tmem_base, tmem_base + 256 + k * 8, db + k * b_k, idesc1, 1u);
                else umma_f16(tmem_base, da + k * a_k, db + k * b_k, idesc1, 1u);
                if (two) umma_f16(tmem_base, da + a2 + k * a_k, db + k * b_k, idesc2, 1u);
            }
            if (++stage == p.stages) stage = 0;
        }
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        const long long t1 = clock64();
        p.cycles[blockIdx.x] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}


// ---- UMMA "program" benchmark (scripts/umma_bench2.py): per k-step issue a fixed, COMPILE-TIME list of UMMAs (entry = N,
// A offset, B offset, accumulator column), fully unrolled with immediate descriptor offsets, so the issue thread spends
// ~2 instructions per UMMA and the measured cycles are the tensor pipe's, not the issue loop's.
struct PEntry { int n, a, b, d; };
constexpr int PKB = 1024;
constexpr int PB0 = 128 * PKB, PB1 = 160 * PKB;      // B tiles (32 KB each, up to N = 256); A tiles are 16 KB apart from 0
#define PA(i) ((i) * 16 * PKB)
constexpr int NPROG = 20;
__device__ constexpr int PCNT[NPROG] = {1, 2, 2, 4, 2, 2, 2, 2, 4, 4, 8, 1, 2, 3, 4, 1, 4, 2, 4, 2};
__device__ constexpr PEntry PROGS[NPROG][8] = {
    /* 0 */ {{128, PA(0), PB0, 0}},
    /* 1 */ {{128, PA(0), PB0, 0}, {128, PA(1), PB0, 0}},                                   // same B, same accumulator
    /* 2 */ {{128, PA(0), PB0, 0}, {128, PA(1), PB0, 128}},                                 // same B, separate accumulators
    /* 3 */ {{128, PA(0), PB0, 0}, {128, PA(1), PB0, 128}, {128, PA(2), PB0, 256}, {128, PA(3), PB0, 384}},
    /* 4 */ {{128, PA(0), PB0, 0}, {128, PA(0), PB1, 128}},                                 // same A
    /* 5 */ {{128, PA(0), PB0, 0}, {128, PA(1), PB1, 128}},                                 // nothing shared
    /* 6 */ {{128, PA(0), PB0, 0}, {64, PA(1), PB0, 0}},                                    // split pair, current order
    /* 7 */ {{128, PA(0), PB0, 0}, {64, PA(1), PB0, 128}},
    /* 8 */ {{128, PA(0), PB0, 0}, {128, PA(2), PB0, 128}, {64, PA(1), PB0, 0}, {64, PA(3), PB0, 128}},   // 2 tiles hi,hi,lo,lo
    /* 9 */ {{128, PA(0), PB0, 0}, {64, PA(1), PB0, 0}, {128, PA(2), PB0, 128}, {64, PA(3), PB0, 128}},   // 2 tiles hi,lo,hi,lo
    /*10 */ {{128, PA(0), PB0, 0}, {128, PA(2), PB0, 128}, {128, PA(4), PB0, 256}, {128, PA(6), PB0, 384},
             {64, PA(1), PB0, 0}, {64, PA(3), PB0, 128}, {64, PA(5), PB0, 256}, {64, PA(7), PB0, 384}},
    /*11 */ {{256, PA(0), PB0, 0}},
    /*12 */ {{256, PA(0), PB0, 0}, {256, PA(1), PB0, 256}},
    /*13 */ {{64, PA(0), PB0, 0}, {64, PA(1), PB0, 0}, {64, PA(0), PB0 + 8 * PKB, 0}},       // unstacked split
    /*14 */ {{64, PA(0), PB0, 0}, {64, PA(1), PB0, 64}, {64, PA(2), PB0, 128}, {64, PA(3), PB0, 192}},
    /*15 */ {{64, PA(0), PB0, 0}},
    /*16 */ {{256, PA(0), PB0, 0}, {256, PA(1), PB0, 256}, {256, PA(2), PB0, 0}, {256, PA(3), PB0, 256}},
    /*17 */ {{64, PA(0), PB0, 0}, {64, PA(1), PB1, 64}},                                    // N=64, nothing shared
    /*18 */ {{128, PA(0), PB0, 0}, {128, PA(1), PB1, 128}, {128, PA(2), PB0, 256}, {128, PA(3), PB1, 384}},   // nothing shared x4
    /*19 */ {{192, PA(0), PB0, 0}, {192, PA(1), PB0, 192}},
};

template <int P>
__device__ __forceinline__ void run_prog(uint32_t base16, uint32_t tmem_base, uint32_t hi_word, int iters) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
#pragma unroll
            for (int i = 0; i < PCNT[P]; ++i) {
                const uint32_t a_lo = base16 + ((PROGS[P][i].a + k * 32) >> 4) + (1u << 16);    // LBO field = 1 (16 bytes)
                const uint32_t b_lo = base16 + ((PROGS[P][i].b + k * 32) >> 4) + (1u << 16);
                const uint64_t da = (static_cast<uint64_t>(hi_word) << 32) | a_lo;
                const uint64_t db = (static_cast<uint64_t>(hi_word) << 32) | b_lo;
                umma_f16(tmem_base + PROGS[P][i].d, da, db, make_idesc_bf16(128, PROGS[P][i].n, 0, 0), 1u);
            }
        }
    }
}

__global__ void __launch_bounds__(128, 1) umma_prog_kernel(int prog, int iters, int a_sbo, long long* cycles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_holder;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 192 * PKB / 4; i += blockDim.x)
        reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + (i & 0xff);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    if (warp == 0) {
        tmem_alloc(&tmem_holder, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_holder;
    if (warp == 0 && elect_one()) {
        const uint32_t base16 = smem_u32(smem) >> 4;
        const uint32_t hi_word = static_cast<uint32_t>(make_smem_desc(0, 0, a_sbo, kSwizzle128) >> 32);
        const long long t0 = clock64();
        switch (prog) {
#define PCASE(P) case P: run_prog<P>(base16, tmem_base, hi_word, iters); break;
            PCASE(0) PCASE(1) PCASE(2) PCASE(3) PCASE(4) PCASE(5) PCASE(6) PCASE(7) PCASE(8) PCASE(9) PCASE(10) PCASE(11)
            PCASE(12) PCASE(13) PCASE(14) PCASE(15) PCASE(16) PCASE(17) PCASE(18) PCASE(19)
#undef PCASE
            default: break;
        }
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        const long long t1 = clock64();
        cycles[blockIdx.x] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace
}  // namespace fcd

using namespace fcd;

extern "C" int fcd_debug_umma_probe(const void* a, const void* b, float* d, int a_rows, int b_rows, int a_blocks,
                                    int b_blocks, int mn_major, int n, int ksteps, int a_shift_rows,
                                    int a_base_offset, int b_shift_rows, int b_base_offset, int a_sbo, int b_sbo,
                                    int a_lbo, int a_kstep, void* stream) {
    FCD_CHECK_ARG(a && b && d, "umma_probe: null pointer");
    FCD_CHECK_ARG(n % 16 == 0 && n >= 16 && n <= 256, "umma_probe: bad N");
    ProbeParams p;
    p.a = (const __nv_bfloat16*)a;
    p.b = (const __nv_bfloat16*)b;
    p.d = d;
    p.a_rows = a_rows; p.b_rows = b_rows; p.a_blocks = a_blocks; p.b_blocks = b_blocks;
    p.mn_major = mn_major; p.n = n; p.ksteps = ksteps;
    p.a_shift_rows = a_shift_rows; p.a_base_offset = a_base_offset;
    p.b_shift_rows = b_shift_rows; p.b_base_offset = b_base_offset;
    p.a_sbo = a_sbo; p.b_sbo = b_sbo;
    p.a_lbo = a_lbo; p.a_kstep = a_kstep;
    const int smem = ((a_blocks * a_rows * 128 + 1023) & ~1023) + b_blocks * b_rows * 128 + 2048;
    FCD_CHECK_ARG(smem <= 200 * 1024, "umma_probe: operands do not fit shared memory");
    FCD_CUDA_OK(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    umma_probe_kernel<<<1, 128, smem, as_stream(stream)>>>(p);
    FCD_LAUNCH_OK();
    return FCD_OK;
}

extern "C" int fcd_debug_umma_bench(int mn_major, int n1, int n2, int a_sbo, int a_lbo, int a_kstep, int a_shift, int a2_off,
                                    int b_sbo, int b_lbo, int b_kstep, int stage_stride, int stages, int b_off, int iters,
                                    int a_tmem, int grid, long long* cycles, void* stream) {
    FCD_CHECK_ARG(cycles && grid > 0 && iters > 0 && stages > 0, "umma_bench: bad arguments");
    FCD_CHECK_ARG(stage_stride * stages <= 200 * 1024, "umma_bench: operands do not fit shared memory");
    BenchParams p;
    p.mn_major = mn_major; p.n1 = n1; p.n2 = n2;
    p.a_sbo = a_sbo; p.a_lbo = a_lbo; p.a_kstep = a_kstep; p.a_shift = a_shift; p.a2_off = a2_off;
    p.b_sbo = b_sbo; p.b_lbo = b_lbo; p.b_kstep = b_kstep;
    p.stage_stride = stage_stride; p.stages = stages; p.b_off = b_off; p.iters = iters; p.a_tmem = a_tmem; p.cycles = cycles;
    FCD_CUDA_OK(cudaFuncSetAttribute(umma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 204 * 1024));
    umma_bench_kernel<<<grid, 128, stage_stride * stages + 2048, as_stream(stream)>>>(p);
    FCD_LAUNCH_OK();
    return FCD_OK;
}

extern "C" int fcd_debug_umma_prog(int prog, int iters, int a_sbo, int grid, long long* cycles, void* stream) {
    FCD_CHECK_ARG(prog >= 0 && prog < NPROG && iters > 0 && grid > 0 && cycles, "umma_prog: bad arguments");
    FCD_CUDA_OK(cudaFuncSetAttribute(umma_prog_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 204 * 1024));
    umma_prog_kernel<<<grid, 128, 194 * 1024, as_stream(stream)>>>(prog, iters, a_sbo, cycles);
    FCD_LAUNCH_OK();
    return FCD_OK;
}
