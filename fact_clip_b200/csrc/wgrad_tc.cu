// Weight gradient on the 5th-generation tensor cores:  dW[n,k] (+)= alpha * sum_t dZ[b,t,n] * A[b,t+off,k]   (bf16 operands)
//
// The contraction runs over the ROWS of two row-major activations, so both UMMA operands are "MN-major": a TMA box of
// [64 rows x 64 columns] bf16 (128-byte rows, 128B swizzle) IS the canonical MN-major SWIZZLE_128B layout -- 64 contiguous
// M (or N) elements per 128-byte line, 8 consecutive K (= frame) lines per 1024-byte swizzle atom (SBO = 1024), the next 64
// M/N elements in the next box (LBO = one box = 8192 bytes) -- and the instruction descriptor marks A and B as MN-major
// (bits 15, 16).  No transposed copy of dZ or A ever exists.
//
// One CTA = one [128 x BN] tile of dW for one chunk of rows of one video: warp 0 streams the boxes through a 4-stage TMA /
// mbarrier ring (tap offsets are a row coordinate; rows outside [0, slot) arrive as zeros), warp 1 issues tcgen05.mma
// (M = 128, N = BN, K = 16 frames per instruction) into a TMEM accumulator, warps 2-5 read it back with tcgen05.ld and write
// the fp32 partial tile; partial_reduce_kernel (train.cu) sums the chunks in a fixed order.  Rows >= len[b] must be zero in
// dZ (the training engine zero-fills gradient rows it never writes).
#include "common.cuh"
#include "tc_common.cuh"

namespace factk {

struct WgParams {
    alignas(64) CUtensorMap zmap;
    alignas(64) CUtensorMap amap;
    int N, K, slot, row_off, rows_per_chunk, nchunk, tiles_k;
    const int32_t* len;
    float* ws;
};

constexpr int WG_BOX = 64 * 128;          // bytes of one [64 rows x 64 cols] bf16 box
constexpr int WG_THREADS = 192;

template <int BN>
struct WgCfg {
    static constexpr int STAGE_Z = 2 * WG_BOX;
    static constexpr int STAGE_A = (BN / 64) * WG_BOX;
    static constexpr int STAGE = STAGE_Z + STAGE_A;
    static constexpr int NSTAGE = 4;
    static constexpr int SMEM = NSTAGE * STAGE + 256 + 1024;
};

// MN-major SWIZZLE_128B operand: start >> 4 | LBO (8192 B: next 64 M/N elements) | SBO (1024 B: next 8 K lines) | version 1 | layout 2
__device__ __forceinline__ uint64_t smem_desc_mn_sw128(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(WG_BOX >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}

template <int BN>
__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_tc_kernel(const __grid_constant__ WgParams p) {
    using Cfg = WgCfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::NSTAGE * Cfg::STAGE);
    uint64_t* empty = full + Cfg::NSTAGE;
    uint64_t* tfull = empty + Cfg::NSTAGE;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tn = blockIdx.x / p.tiles_k, tk = blockIdx.x % p.tiles_k;
    const int chunk = blockIdx.y, b = blockIdx.z;
    const int len_b = p.len ? min(p.len[b], p.slot) : p.slot;
    const int r0 = chunk * p.rows_per_chunk;
    if (r0 >= len_b) return;                                  // uniform per CTA: nothing allocated yet
    const int r1 = min(r0 + p.rows_per_chunk, p.slot);
    const int nsteps = (min(r1, len_b) - r0 + 63) / 64;       // 64-row stages (rows in [len, slot) are zero in dZ)

    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&p.zmap);
        tc::tma_prefetch_desc(&p.amap);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int i = 0; i < Cfg::NSTAGE; ++i) {
                tc::mbar_init(&full[i], 1);
                tc::mbar_init(&empty[i], 1);
            }
            tc::mbar_init(tfull, 1);
            tc::fence_barrier_init();
        }
        __syncwarp();
        tc::tmem_alloc(tmem_slot, BN);
        tc::tmem_relinquish();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int s = 0; s < nsteps; ++s) {
                tc::mbar_wait(&empty[stage], phase ^ 1);
                tc::mbar_arrive_expect_tx(&full[stage], Cfg::STAGE);
                uint8_t* st = smem + stage * Cfg::STAGE;
                const int row = r0 + s * 64;
#pragma unroll
                for (int j = 0; j < 2; ++j) tc::tma_load_3d(st + j * WG_BOX, &p.zmap, &full[stage], tn * 128 + j * 64, row, b);
#pragma unroll
                for (int j = 0; j < BN / 64; ++j)
                    tc::tma_load_3d(st + Cfg::STAGE_Z + j * WG_BOX, &p.amap, &full[stage], tk * BN + j * 64, row + p.row_off, b);
                if (++stage == Cfg::NSTAGE) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // A and B MN-major: bits 15 and 16
            constexpr uint32_t idesc = tc::instr_desc(128, BN, false) | (1u << 15) | (1u << 16);
            int stage = 0;
            uint32_t phase = 0, accumulate = 0;
            for (int s = 0; s < nsteps; ++s) {
                tc::mbar_wait(&full[stage], phase);
                tc::tc_fence_after();
                const uint32_t z_addr = tc::smem_u32(smem + stage * Cfg::STAGE);
                const uint32_t a_addr = z_addr + Cfg::STAGE_Z;
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) {          // 16 frames = two 8-line swizzle atoms per instruction
                    tc::umma<false>(tmem_base, smem_desc_mn_sw128(z_addr + k4 * 2048), smem_desc_mn_sw128(a_addr + k4 * 2048), idesc,
                                    accumulate);
                    accumulate = 1;
                }
                tc::umma_commit(&empty[stage]);
                if (++stage == Cfg::NSTAGE) { stage = 0; phase ^= 1; }
            }
            tc::umma_commit(tfull);
        }
    } else {
        const int q = warp & 3;                               // TMEM lane quarter this warp may read
        tc::mbar_wait(tfull, 0);
        tc::tc_fence_after();
        const int n = tn * 128 + q * 32 + lane;
        float* out = p.ws + (size_t)(b * p.nchunk + chunk) * (size_t)p.N * (size_t)p.K + (size_t)n * p.K + (size_t)tk * BN;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
            float v[32];
            tc::tmem_ld32(tmem_base + c * 32 + ((uint32_t)(q * 32) << 16), v);
            tc::tmem_ld_wait();
            if (n < p.N) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const int k = tk * BN + c * 32 + j;
                    if (k + 4 <= p.K) *reinterpret_cast<float4*>(out + c * 32 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    else
                        for (int jj = 0; jj < 4 && k + jj < p.K; ++jj) out[c * 32 + j + jj] = v[j + jj];
                }
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem_base, BN);
    }
}

}  // namespace factk

using namespace factk;

/* 1 when factk_wgrad_tc accepts the operands (bf16 dZ and A, 16-byte aligned rows, N and K multiples of 64, K-tile 64/128/256). */
extern "C" int factk_wgrad_tc_supported(int dz_dtype, int lddz, int a_dtype, int lda, int N, int K, int slot) {
    if (dz_dtype != FACTK_BF16 || a_dtype != FACTK_BF16) return 0;
    if (N % 64 || K % 64 || lddz % 8 || lda % 8 || slot % 64) return 0;
    return 1;
}

static int wg_rows_per_chunk(int B, int slot, int tiles) {
    int rc = 2048;
    while (rc > 256 && (long)B * ((slot + rc - 1) / rc) * tiles < 2 * 148) rc >>= 1;
    return rc;
}

extern "C" size_t factk_wgrad_tc_ws_floats(int B, int slot, int N, int K) {
    // sized for the finest chunking the launcher may choose
    return (size_t)B * ((slot + 255) / 256) * (size_t)N * (size_t)K;
}

extern "C" int factk_wgrad_tc(const void* dZ, int lddz, const void* A, int lda, int a_slot, int row_off, int N, int K, float* dW,
                              int lddw, long long dw_bstride, float alpha, int accumulate, int B, int slot, const int32_t* len,
                              float* ws, void* stream) {
    FACTK_REQUIRE(dZ && A && dW && ws && B > 0 && slot > 0, "factk_wgrad_tc: bad args");
    FACTK_REQUIRE(factk_wgrad_tc_supported(FACTK_BF16, lddz, FACTK_BF16, lda, N, K, slot) && a_slot > 0 && aligned16(dZ) && aligned16(A),
                  "factk_wgrad_tc: unsupported operands (N=%d K=%d lddz=%d lda=%d slot=%d)", N, K, lddz, lda, slot);
    const int BN = (K % 256 == 0) ? 256 : ((K % 128 == 0) ? 128 : 64);
    WgParams p;
    if (!tc_get_map(&p.zmap, dZ, 2, (uint64_t)N, (uint64_t)slot, (uint64_t)B, (uint64_t)lddz, (uint64_t)slot * lddz, 64)) return FACTK_ERR_CUDA;
    if (!tc_get_map(&p.amap, A, 2, (uint64_t)K, (uint64_t)slot, (uint64_t)B, (uint64_t)lda, (uint64_t)a_slot * lda, 64)) return FACTK_ERR_CUDA;
    const int tiles_n = (N + 127) / 128;
    p.tiles_k = K / BN;
    p.N = N; p.K = K; p.slot = slot; p.row_off = row_off; p.len = len; p.ws = ws;
    p.rows_per_chunk = wg_rows_per_chunk(B, slot, tiles_n * p.tiles_k);
    p.nchunk = (slot + p.rows_per_chunk - 1) / p.rows_per_chunk;
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid(tiles_n * p.tiles_k, p.nchunk, B);
    static unsigned long long devs = 0;
    if (first_use_on_device(devs)) {
        cudaFuncSetAttribute(wgrad_tc_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, WgCfg<256>::SMEM);
        cudaFuncSetAttribute(wgrad_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, WgCfg<128>::SMEM);
        cudaFuncSetAttribute(wgrad_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, WgCfg<64>::SMEM);
    }
    if (BN == 256) wgrad_tc_kernel<256><<<grid, WG_THREADS, WgCfg<256>::SMEM, st>>>(p);
    else if (BN == 128) wgrad_tc_kernel<128><<<grid, WG_THREADS, WgCfg<128>::SMEM, st>>>(p);
    else wgrad_tc_kernel<64><<<grid, WG_THREADS, WgCfg<64>::SMEM, st>>>(p);
    launch_partial_reduce(ws, (size_t)N * K, N * K, K, dW, lddw, dw_bstride, B, slot, len, p.nchunk, p.rows_per_chunk, alpha, accumulate, st);
    return check_launch("factk_wgrad_tc");
}
