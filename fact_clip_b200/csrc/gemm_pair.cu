// Frame-side bf16 GEMM on CTA pairs:  y[b,t,:N] = act(x[b,t,:K] W^T + bias + pre[b, idx(t)])   (bf16 in, bf16 out)
//
// The frame / segment side Linears and 1x1 convolutions that are not part of a dilated residual layer (conv_out
// basic.py:182,213; the SCA key/value projection basic.py:465,513; sf_merge blocks.py:414,445; the CLIP projection
// blocks.py:153-159; seg_combine blocks.py:402) are plain [rows x K] x [K x N] products with K, N in {256, 512}.  This
// kernel applies the recipe measured on tcn_layer_kernel (profiles/r1_fused_summary.md) to them:
//   * tcgen05.mma.cta_group::2, M = 256 per CTA pair: each CTA loads its own 128 rows and HALF of the weight tile, which
//     halves the weight stream from L2 and the shared-memory fill traffic per MMA;
//   * two TMEM accumulators: the MMAs of tile i+1 run under the epilogue of tile i;
//   * epilogue: tcgen05.ld -> bf16 -> staged thread = row in shared memory (conflict-free 16-byte stores), accumulator
//     released at once, then a coalesced copy-out (lane = 16-byte column segment) adds bias / pre-activation addend and
//     applies ReLU in fp32 -- every global access is a full 16-byte-per-lane access, the shared-memory data pipe is the
//     scarce resource.
// Warp roles (320 threads): warp 0 TMA producer, warp 1 MMA issuer (leader CTA), warps 2..9 epilogue.
#include <cstring>

#include "common.cuh"
#include "tc_common.cuh"

namespace factk {

struct PairParams {
    alignas(64) CUtensorMap xmap;    // x [B][a_slot][lda] bf16, box 64 x 128 x 1
    alignas(64) CUtensorMap wmap;    // W [N][ldw] bf16, box 64 x BN/2 x 1
    __nv_bfloat16* y;
    const float* bias;               // [N] or null
    const float* pre;                // optional fp32 addend rows pre[b*pre_bstride + idx*ldpre + n]
    const int32_t* pre_idx;          // optional [B][slot] row index into pre (default: the row itself)
    long long pre_bstride;
    const int32_t* len;
    int ldpre, ldy, relu;
    int B, slot, N, kchunks, tiles_m, tiles_n, total;
};

constexpr int GP_EPI_WARPS = 8;
constexpr int GP_THREADS = 64 + GP_EPI_WARPS * 32;

template <int BN>
struct PairCfg {
    static constexpr int STAGE_A = 128 * 128;
    static constexpr int STAGE_B = (BN / 2) * 128;
    static constexpr int STAGE = STAGE_A + STAGE_B;
    static constexpr int OBYTES = 128 * BN * 2;        // bf16 staging of one output tile (this CTA's 128 rows)
    static constexpr int BAR_BYTES = 256;
    static constexpr int LEN_CACHE = 320;
    static constexpr int NSTAGE_RAW = (232448 - 1024 - BAR_BYTES - LEN_CACHE * 4 - OBYTES) / STAGE;
    static constexpr int NSTAGE = NSTAGE_RAW > 8 ? 8 : NSTAGE_RAW;
    static constexpr int SMEM = 1024 + NSTAGE * STAGE + OBYTES + BAR_BYTES + LEN_CACHE * 4;
    static_assert(NSTAGE >= 4, "pipeline too shallow");
};

// tiles in (video, 256-row super tile, n tile) order, n fastest: the A rows of a super tile are re-used from L2
struct PairTileIter {
    int st, step, tiles_m, tiles_n, total, slot;
    const int32_t* len;
    __device__ PairTileIter(const PairParams& p, int unit, int nunits, const int32_t* len_)
        : st(unit - nunits), step(nunits), tiles_m(p.tiles_m), tiles_n(p.tiles_n), total(p.total), slot(p.slot), len(len_) {}
    __device__ bool next(int& b, int& t0s, int& nt, int& len_b) {
        while (true) {
            st += step;
            if (st >= total) return false;
            nt = st % tiles_n;
            const int r = st / tiles_n;
            b = r / tiles_m;
            t0s = (r - b * tiles_m) * 256;
            len_b = len ? min(len[b], slot) : slot;
            if (t0s < len_b) return true;
        }
    }
};

template <int BN>
__global__ void __launch_bounds__(GP_THREADS, 1) gemm_pair_kernel(const __grid_constant__ PairParams p) {
    using Cfg = PairCfg<BN>;
    constexpr int NS = Cfg::NSTAGE;
    constexpr int CG = 2;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* obuf = smem + NS * Cfg::STAGE;
    uint64_t* full = reinterpret_cast<uint64_t*>(obuf + Cfg::OBYTES);   // [NS] leader: both CTAs' TMA bytes
    uint64_t* empty = full + NS;                                        // [NS] each CTA: stage consumed
    uint64_t* tfull = empty + NS;                                       // [2]  each CTA: accumulator ready
    uint64_t* tempty = tfull + 2;                                       // [2]  leader: accumulator drained
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    int32_t* slen = reinterpret_cast<int32_t*>(obuf + Cfg::OBYTES + Cfg::BAR_BYTES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = tc::cluster_ctarank();
    const bool leader = rank == 0;
    const int unit = blockIdx.x / CG, nunits = gridDim.x / CG;

    const bool len_cached = p.len != nullptr && p.B <= Cfg::LEN_CACHE;
    if (len_cached)
        for (int i = threadIdx.x; i < p.B; i += GP_THREADS) slen[i] = p.len[i];
    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&p.xmap);
        tc::tma_prefetch_desc(&p.wmap);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int i = 0; i < NS; ++i) {
                tc::mbar_init(&full[i], 1);
                tc::mbar_init(&empty[i], 1);
            }
            for (int i = 0; i < 2; ++i) {
                tc::mbar_init(&tfull[i], 1);
                tc::mbar_init(&tempty[i], CG * GP_EPI_WARPS);
            }
            tc::fence_barrier_init();
        }
        __syncwarp();
        tc::tmem_alloc_cg<CG>(tmem_slot, 2 * BN);
        tc::tmem_relinquish_cg<CG>();
    }
    tc::tc_fence_before();
    tc::cluster_sync_all();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int32_t* lenp = len_cached ? slen : p.len;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t full0 = tc::mapa_u32(tc::smem_u32(&full[0]), 0);
            PairTileIter iter(p, unit, nunits, lenp);
            int b, t0s, nt, len_b;
            while (iter.next(b, t0s, nt, len_b)) {
                const int t0 = t0s + (int)rank * 128;
#pragma unroll 1
                for (int kc = 0; kc < p.kchunks; ++kc) {
                    tc::mbar_wait(&empty[stage], phase ^ 1);
                    if (leader) tc::mbar_arrive_expect_tx(&full[stage], CG * Cfg::STAGE);
                    const uint32_t st = tc::smem_u32(smem + stage * Cfg::STAGE);
                    tc::tma_load_3d_cg<CG>(st, &p.xmap, full0 + stage * 8, kc * 64, t0, b);
                    tc::tma_load_3d_cg<CG>(st + Cfg::STAGE_A, &p.wmap, full0 + stage * 8, kc * 64, nt * BN + (int)rank * (BN / 2), 0);
                    if (++stage == NS) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (leader CTA)
        if (leader && lane == 0) {
            constexpr uint32_t idesc = tc::instr_desc(256, BN, false);
            int stage = 0, it = 0;
            uint32_t phase = 0;
            PairTileIter iter(p, unit, nunits, lenp);
            int b, t0s, nt, len_b;
            while (iter.next(b, t0s, nt, len_b)) {
                const int acc = it & 1;
                if (it >= 2) {
                    tc::mbar_wait(&tempty[acc], (uint32_t)(((it >> 1) - 1) & 1));
                    tc::tc_fence_after();
                }
                ++it;
                const uint32_t d = tmem_base + acc * BN;
                uint32_t accumulate = 0;
#pragma unroll 1
                for (int kc = 0; kc < p.kchunks; ++kc) {
                    tc::mbar_wait(&full[stage], phase);
                    tc::tc_fence_after();
                    const uint32_t a_addr = tc::smem_u32(smem + stage * Cfg::STAGE);
                    const uint32_t b_addr = a_addr + Cfg::STAGE_A;
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) {
                        tc::umma_bf16_cg<CG>(d, tc::smem_desc_sw128(a_addr + k4 * 32), tc::smem_desc_sw128(b_addr + k4 * 32), idesc, accumulate);
                        accumulate = 1;
                    }
                    tc::umma_commit_cg<CG>(&empty[stage]);
                    if (++stage == NS) { stage = 0; phase ^= 1; }
                }
                tc::umma_commit_cg<CG>(&tfull[acc]);
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue warps
        const int ew = warp - 2;
        const int q = warp & 3;              // TMEM lane quarter this warp may access
        const int half = ew >> 2;            // column half of the tile
        constexpr int HALF_COLS = BN / 2;
        constexpr int NG = HALF_COLS / 32;   // 32-column groups per warp
        constexpr int NPC = HALF_COLS / 64;  // 128-byte chunks of a staged row owned by this warp
        const int r = q * 32 + lane;
        const uint32_t o_u32 = tc::smem_u32(obuf);
        const uint32_t tempty_l = tc::mapa_u32(tc::smem_u32(&tempty[0]), 0);
        // staging layout: [BN/64 chunks][128 rows][128 B], 16-byte groups XOR-swizzled by the row (conflict-free both ways)
        auto sw_addr = [&](int row, int col) -> uint32_t {
            return o_u32 + (uint32_t)(col >> 6) * 16384u + (uint32_t)row * 128u + (uint32_t)((((col & 63) >> 3) ^ (row & 7)) << 4);
        };
        const uint32_t my_row_addr = o_u32 + (uint32_t)((half * HALF_COLS) >> 6) * 16384u + (uint32_t)r * 128u;
        const uint32_t my_rx = (uint32_t)(r & 7) << 4;
        auto row_addr = [&](int c) -> uint32_t { return my_row_addr + (uint32_t)(c >> 3) * 16384u + ((uint32_t)((c & 7) << 4) ^ my_rx); };
        constexpr int LPR = 8 * NPC, RPI = 32 / LPR, NIT = 32 / RPI;
        const int lr = lane / LPR, lc = lane % LPR;
        const int ccol = half * HALF_COLS + lc * 8;                    // first of this lane's 8 columns inside the tile
        uint32_t tfc0 = 0u, tfc1 = 0u;
        int it = 0;
        PairTileIter iter(p, unit, nunits, lenp);
        int b, t0s, nt, len_b;
        while (iter.next(b, t0s, nt, len_b)) {
            const int acc = it & 1;
            ++it;
            const int t0 = t0s + (int)rank * 128;
            const int rows_valid = len_b - (t0 + q * 32);
            const size_t rowbase = (size_t)b * p.slot + t0 + q * 32;
            const int ncol = nt * BN + ccol;                           // global output column of this lane
            float bb[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) bb[j] = 0.f;
            if (p.bias) {
                const float4 t0_ = __ldg(reinterpret_cast<const float4*>(p.bias + ncol)), t1_ = __ldg(reinterpret_cast<const float4*>(p.bias + ncol + 4));
                bb[0] = t0_.x; bb[1] = t0_.y; bb[2] = t0_.z; bb[3] = t0_.w; bb[4] = t1_.x; bb[5] = t1_.y; bb[6] = t1_.z; bb[7] = t1_.w;
            }
            {
                const uint32_t par = (acc ? tfc1 : tfc0) & 1u;
                tc::mbar_wait(&tfull[acc], par);
                if (acc) ++tfc1; else ++tfc0;
                tc::tc_fence_after();
            }
            const uint32_t taddr = tmem_base + acc * BN + half * HALF_COLS + ((uint32_t)(q * 32) << 16);
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                float v[32];
                tc::tmem_ld32(taddr + g * 32, v);
                tc::tmem_ld_wait();
                if (g + 1 == NG) {
                    tc::tc_fence_before();    // accumulator fully read: hand it back to the MMA issuer
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive_cluster(tempty_l + (uint32_t)acc * 8u);
                }
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    tc::sts_v4(row_addr(g * 4 + c), tc::pack_bf16x2(v[c * 8], v[c * 8 + 1]), tc::pack_bf16x2(v[c * 8 + 2], v[c * 8 + 3]),
                               tc::pack_bf16x2(v[c * 8 + 4], v[c * 8 + 5]), tc::pack_bf16x2(v[c * 8 + 6], v[c * 8 + 7]));
            }
            __syncwarp();
#pragma unroll 4
            for (int i = 0; i < NIT; ++i) {
                const int rr = i * RPI + lr;
                uint4 o;
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(o.x), "=r"(o.y), "=r"(o.z), "=r"(o.w)
                             : "r"(sw_addr(q * 32 + rr, ccol)) : "memory");
                if (rr >= rows_valid) continue;
                const size_t grow = rowbase + rr;
                float x[8];
                {
                    const uint32_t ow[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float2 d = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ow[k]));
                        x[2 * k] = d.x + bb[2 * k];
                        x[2 * k + 1] = d.y + bb[2 * k + 1];
                    }
                }
                if (p.pre) {
                    const int prow = p.pre_idx ? p.pre_idx[grow] : (t0 + q * 32 + rr);
                    const float* pp = p.pre + (size_t)b * (size_t)p.pre_bstride + (size_t)prow * (size_t)p.ldpre + ncol;
                    const float4 a0 = __ldg(reinterpret_cast<const float4*>(pp)), a1 = __ldg(reinterpret_cast<const float4*>(pp + 4));
                    x[0] += a0.x; x[1] += a0.y; x[2] += a0.z; x[3] += a0.w; x[4] += a1.x; x[5] += a1.y; x[6] += a1.z; x[7] += a1.w;
                }
                if (p.relu) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) x[k] = fmaxf(x[k], 0.f);
                }
                *reinterpret_cast<uint4*>(p.y + grow * (size_t)p.ldy + ncol) =
                    make_uint4(tc::pack_bf16x2(x[0], x[1]), tc::pack_bf16x2(x[2], x[3]), tc::pack_bf16x2(x[4], x[5]), tc::pack_bf16x2(x[6], x[7]));
            }
            __syncwarp();                     // staged rows are read before the next tile overwrites them
        }
    }
    __syncwarp();

    tc::tc_fence_before();
    tc::cluster_sync_all();
    if (warp == 1) {
        tc::tc_fence_after();
        tc::tmem_dealloc_cg<CG>(tmem_base, 2 * BN);
    }
}

template <int BN>
static int launch_pair(const PairParams& p, int sms, cudaStream_t st) {
    using Cfg = PairCfg<BN>;
    static unsigned long long attr_devs = 0;
    if (first_use_on_device(attr_devs)) {
        cudaError_t e = cudaFuncSetAttribute(gemm_pair_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
        if (e != cudaSuccess) { set_error("factk_gemm_pair: smem attribute: %s", cudaGetErrorString(e)); return FACTK_ERR_CUDA; }
    }
    int units = sms / 2;
    if (units > p.total) units = p.total;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(units * 2, 1, 1);
    cfg.blockDim = dim3(GP_THREADS, 1, 1);
    cfg.dynamicSmemBytes = Cfg::SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_pair_kernel<BN>, p);
    if (e != cudaSuccess) { set_error("factk_gemm_pair: launch: %s", cudaGetErrorString(e)); return FACTK_ERR_CUDA; }
    return check_launch("factk_gemm_pair");
}

}  // namespace factk

using namespace factk;

extern "C" int factk_gemm_pair_supported(int K, int N) { return (K > 0 && K % 64 == 0 && N > 0 && N % 128 == 0) ? 1 : 0; }

extern "C" int factk_gemm_pair(const void* x, int lda, int a_slot, const void* W, int ldw, int K, int N, const float* bias,
                               const float* pre, int ldpre, long long pre_bstride, const int32_t* pre_idx, int relu, void* y,
                               int ldy, int B, int slot, const int32_t* len, void* stream) {
    FACTK_REQUIRE(x && W && y && B > 0 && slot > 0, "factk_gemm_pair: bad args");
    FACTK_REQUIRE(factk_gemm_pair_supported(K, N), "factk_gemm_pair: K must be a multiple of 64 and N of 128 (got K=%d N=%d)", K, N);
    FACTK_REQUIRE(aligned16(x) && aligned16(W) && aligned16(y) && (lda % 8) == 0 && (ldw % 8) == 0 && (ldy % 8) == 0 && lda >= K &&
                      ldw >= K && ldy >= N && a_slot >= slot,
                  "factk_gemm_pair: alignment / leading dimensions");
    FACTK_REQUIRE(!bias || aligned16(bias), "factk_gemm_pair: bias alignment");
    FACTK_REQUIRE(!pre || (aligned16(pre) && (ldpre % 4) == 0 && (pre_bstride % 4) == 0), "factk_gemm_pair: pre alignment");
    const int BN = (N % 256 == 0) ? 256 : 128;
    PairParams p;
    memset(&p, 0, sizeof(p));
    if (!tc_get_map(&p.xmap, x, 2, K, a_slot, B, lda, (uint64_t)a_slot * lda, 128)) return FACTK_ERR_CUDA;
    if (!tc_get_map(&p.wmap, W, 2, K, N, 1, ldw, (uint64_t)N * ldw, BN / 2)) return FACTK_ERR_CUDA;
    p.y = reinterpret_cast<__nv_bfloat16*>(y);
    p.bias = bias; p.pre = pre; p.pre_idx = pre_idx; p.pre_bstride = pre_bstride; p.ldpre = ldpre; p.ldy = ldy; p.relu = relu;
    p.len = len; p.B = B; p.slot = slot; p.N = N; p.kchunks = K / 64;
    p.tiles_m = (slot + 255) / 256;
    p.tiles_n = N / BN;
    p.total = B * p.tiles_m * p.tiles_n;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaStream_t st = (cudaStream_t)stream;
    return BN == 256 ? launch_pair<256>(p, sms, st) : launch_pair<128>(p, sms, st);
}
