// X2Y_map in the f2a direction (models/basic.py:349-389 with X = frame / segment rows, Y = action tokens; blocks.py:346,458):
// every token attends ALL rows of its video -- softmax over the rows, then the weighted row sum -- as ONE tcgen05 / TMEM
// kernel, flash-attention style with the tokens as queries and the rows as keys AND values (one head of H = 256 channels):
//
//   S[128 x 64]   = qt . rows^T                 qt = alpha Wk^T Y_Q(tokens + pos) (token-side fold, engine.py); the per-token
//                                               bias alpha yq . bk is constant along the rows and cancels in the softmax
//   P             = exp2(S log2e - m)           tcgen05.ld: one thread = one token row of S in registers; running max / sum per
//                                               token; the output accumulator is rescaled only when a row maximum grew by more
//                                               than 2^8 (the stale maximum stays the reference otherwise: P <= 256, exact in
//                                               fp32 sums, harmless in bf16)
//   O[128 x 256] += P . rows                    P as bf16 in the 128B-swizzled K-major A layout in shared memory; the SAME TMA
//                                               tile of the rows is the K-major B operand of S and the MN-major B operand of O
//
// The rows are read ONCE (512 B per row); the fp32 logits [B, slot, M] that the unfused chain (tcgen05 logit GEMM -> column
// statistics -> mma.sync apply) wrote and read twice never exist.  One CTA = one 2048-row split of one video: warp 0 = TMA
// producer (4-stage ring of [64 rows x 256 channels]), warp 1 issues S = Q X^T, warp 2 issues O += P X, warps 4-11 = softmax /
// epilogue (TMEM lane quarter = warp % 4; the two warps of a quarter split the 64 columns of a tile).  Every split writes (max, sum) and the unnormalised O per token;
// f2a_combine_kernel merges the splits in a fixed order (bit-reproducible, batch invariant).
// This token-major kernel serves hid_dim 256 (with more than 80 tokens by default); hid_dim 512 -- every shipped configuration --
// runs the transposed kernel of f2a_fused_t.cu (rows on the M side), whose header explains why.  factk_f2a_fused dispatches.
#include <cstdlib>

#include "common.cuh"
#include "tc_common.cuh"

namespace factk {

constexpr int FF_SPLIT = 2048;                       // rows per CTA: 64 videos x 4096 rows = 128 CTAs, one round on 148 SMs.  A constant
                                                     // (not a function of B), so results do not depend on the batch a video is in
constexpr int FF_TILE = 64;                          // rows per pipeline stage
constexpr int FF_H = 256;
constexpr int FF_QBYTES = 4 * 128 * 128;             // 4 boxes [128 tokens x 64 channels] bf16
constexpr int FF_XBYTES = 4 * FF_TILE * 128;         // 4 boxes [64 rows x 64 channels] bf16
constexpr int FF_NSTAGE = 4;                         // with 3 stages the S issuer waited ~1100 cycles for rows every third tile
constexpr int FF_PBYTES = 128 * 128;                 // [128 tokens x 64 rows] bf16
constexpr int FF_MXBYTES = 3 * 2 * 128 * 4;          // row maxima exchanged between the two column halves: [tile parity][half][token]; [2] = final sums
constexpr int FF_SMEM = FF_QBYTES + FF_NSTAGE * FF_XBYTES + FF_PBYTES + FF_MXBYTES + 256 + 1024;
constexpr int FF_THREADS = 128 + 256;                // TMA, S issuer, P X issuer, (idle), 8 softmax warps
constexpr float FF_RESCALE = 8.f;                    // log2 units: rescale O only when the maximum grew by more than this
static_assert(FF_SMEM <= 232448, "dynamic shared memory limit");

struct FfParams {
    alignas(64) CUtensorMap qmap;    // qt   [B][M][H]    bf16, box 64 x 128 x 1 (token rows >= M arrive as zeros)
    alignas(64) CUtensorMap xmap;    // rows [B][slot][H] bf16, box 64 x 64 x 1
    int M, slot, nsplit;
    const int32_t* len;
    float* stats;                    // [B][nsplit][M][2] = (max in log2 units, sum)
    float* part;                     // [B][nsplit][M][H] unnormalised weighted row sums
    long long* dbg;                  // optional clock64 timeline of CTA (0,0), softmax warp 4 lane 0 (development aid)
};

// descriptors as (low word, constant high word): see attn_tc.cu -- the issuing lane does 32-bit adds per MMA
constexpr uint32_t FF_DESC_HI = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);       // SBO | version 1 | SWIZZLE_128B
__device__ __forceinline__ uint32_t ff_desc_lo_k(uint32_t saddr) { return ((saddr >> 4) & 0x3FFFu) | (1u << 16); }
__device__ __forceinline__ uint32_t ff_desc_lo_mn(uint32_t saddr) { return ((saddr >> 4) & 0x3FFFu) | ((uint32_t)(8192 >> 4) << 16); }
__device__ __forceinline__ void ff_umma(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\tsetp.ne.b32 p, %5, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
        ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(FF_DESC_HI), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void ff_tmem_st32(uint32_t taddr, const float v[32]) {
    const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
          "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
          "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
          "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ float ff_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__global__ void __launch_bounds__(FF_THREADS, 1) f2a_fused_kernel(const __grid_constant__ FfParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* qs = smem;
    uint8_t* xs = qs + FF_QBYTES;
    uint8_t* ps = xs + FF_NSTAGE * FF_XBYTES;
    float* mxs = reinterpret_cast<float*>(ps + FF_PBYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(ps + FF_PBYTES + FF_MXBYTES);
    uint64_t *x_full = bars, *x_empty = bars + FF_NSTAGE, *s_full = bars + 2 * FF_NSTAGE, *s_empty = s_full + 2, *p_full = s_full + 4,
             *p_empty = s_full + 5, *q_full = s_full + 6, *o_full = s_full + 7;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long t_entry = clock64();
    const int split = blockIdx.x, b = blockIdx.y;
    const int len_b = p.len ? min(p.len[b], p.slot) : p.slot;
    const int r0 = split * FF_SPLIT;
    if (r0 >= len_b) return;                             // uniform per CTA: nothing allocated yet
    const int nrows = min(r0 + FF_SPLIT, len_b) - r0;
    const int ntile = (nrows + FF_TILE - 1) / FF_TILE;

    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&p.qmap);
        tc::tma_prefetch_desc(&p.xmap);
        for (int i = 0; i < FF_NSTAGE; ++i) {
            tc::mbar_init(&x_full[i], 1);
            tc::mbar_init(&x_empty[i], 2);               // the S issuer and the P X issuer both release a stage
        }
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(&s_full[i], 1);
            tc::mbar_init(&s_empty[i], 8);
        }
        tc::mbar_init(p_full, 8);
        tc::mbar_init(p_empty, 1);
        tc::mbar_init(q_full, 1);
        tc::mbar_init(o_full, 1);
        tc::fence_barrier_init();
    }
    __syncthreads();                                     // barriers initialised; the TMA warp runs ahead from here
    uint32_t tmem_base = 0;
    if (warp >= 1) {
        if (warp == 1) {
            tc::tmem_alloc(tmem_slot, 512);
            tc::tmem_relinquish();
        }
        tc::tc_fence_before();
        asm volatile("bar.sync 1, %0;" ::"n"(FF_THREADS - 32) : "memory");
        tc::tc_fence_after();
        tmem_base = *tmem_slot;
    }
    const uint32_t tmem_o = tmem_base + 2 * FF_TILE;     // S buffers: columns [0, 128); O: [128, 384)

    if (warp == 0) {
        if (lane == 0) {
            tc::mbar_arrive_expect_tx(q_full, FF_QBYTES);
#pragma unroll
            for (int j = 0; j < 4; ++j) tc::tma_load_3d(qs + j * 16384, &p.qmap, q_full, j * 64, 0, b);
            for (int t = 0; t < ntile; ++t) {
                const int st = t % FF_NSTAGE;
                tc::mbar_wait(&x_empty[st], ((t / FF_NSTAGE) & 1) ^ 1);
                tc::mbar_arrive_expect_tx(&x_full[st], FF_XBYTES);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    tc::tma_load_3d(xs + st * FF_XBYTES + j * (FF_TILE * 128), &p.xmap, &x_full[st], j * 64, r0 + t * FF_TILE, b);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {                                  // S[t & 1] = Q X_t^T: 4 channel boxes x 4 k-steps of 16
            constexpr uint32_t idesc_s = tc::instr_desc(128, FF_TILE, false);
            const uint32_t q_lo = ff_desc_lo_k(tc::smem_u32(qs)), x_lo = ff_desc_lo_k(tc::smem_u32(xs));
            tc::mbar_wait(q_full, 0);
            for (int t = 0; t < ntile; ++t) {
                const int st = t % FF_NSTAGE, sb = t & 1;
                tc::mbar_wait(&x_full[st], (t / FF_NSTAGE) & 1);
                tc::mbar_wait(&s_empty[sb], ((t >> 1) & 1) ^ 1);
                tc::tc_fence_after();
                const uint32_t xa = x_lo + (uint32_t)((st * FF_XBYTES) >> 4);
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4)
                        ff_umma(tmem_base + sb * FF_TILE, q_lo + (uint32_t)((j * 16384 + k4 * 32) >> 4),
                                xa + (uint32_t)((j * FF_TILE * 128 + k4 * 32) >> 4), idesc_s, (j | k4) ? 1u : 0u);
                tc::umma_commit(&s_full[sb]);
                tc::umma_commit(&x_empty[st]);
            }
        }
    } else if (warp == 2) {
        if (lane == 0) {                                  // O += P_t X_t: N = 256 channels, K = 64 rows in 4 steps of 16
            constexpr uint32_t idesc_o = tc::instr_desc(128, FF_H, false) | (1u << 16);        // B (= rows) MN-major
            const uint32_t p_lo = ff_desc_lo_k(tc::smem_u32(ps)), x_lo = ff_desc_lo_mn(tc::smem_u32(xs));
            for (int t = 0; t < ntile; ++t) {
                const int st = t % FF_NSTAGE;
                tc::mbar_wait(&x_full[st], (t / FF_NSTAGE) & 1);
                tc::mbar_wait(p_full, t & 1);
                tc::tc_fence_after();
                const uint32_t xa = x_lo + (uint32_t)((st * FF_XBYTES) >> 4);
#pragma unroll
                for (int k = 0; k < FF_TILE / 16; ++k)
                    ff_umma(tmem_o, p_lo + k * 2, xa + k * 128, idesc_o, (t > 0 || k > 0) ? 1u : 0u);      // + 32 B / + 2048 B
                tc::umma_commit(p_empty);
                tc::umma_commit(&x_empty[st]);
            }
            tc::umma_commit(o_full);
        }
    } else if (warp >= 4) {
        // Two warps per TMEM lane quarter: warp 4 + q reads columns [0, 32) of a tile of S, warp 8 + q columns [32, 64) of the
        // SAME 32 token rows; the pair exchanges its row maxima through shared memory (named barrier 2 + q, 64 threads), so
        // both halves use one reference and every thread does half of the exp2 work.  Two warps per scheduler also hide each
        // other's latencies (one warp per scheduler ran at 1530 cycles per tile, of which ~450 were barrier round trips).
        const int q = warp & 3, half = (warp - 4) >> 2;
        const int row = q * 32 + lane;                     // token
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const uint32_t prow = tc::smem_u32(ps) + (uint32_t)row * 128u;
        constexpr float LOG2E = 1.4426950408889634f;
        constexpr int HT = FF_TILE / 2;                    // columns of S per thread
        float mx = -INFINITY, ls = 0.f;                    // reference maximum (log2 units), running sum of this half's columns
        const bool dbg_on = p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && warp == 4 && lane == 0;
        if (dbg_on) { p.dbg[0] = t_entry; p.dbg[1] = clock64(); }
        for (int t = 0; t < ntile; ++t) {
            const int sb = t & 1;
            const int valid = min(FF_TILE, nrows - t * FF_TILE) - half * HT;      // columns of this half below len[b] (may be <= 0)
            if (dbg_on && t < 12) p.dbg[8 + t * 4 + 0] = clock64();
            tc::mbar_wait(&s_full[sb], (t >> 1) & 1);
            tc::tc_fence_after();
            if (dbg_on && t < 12) p.dbg[8 + t * 4 + 1] = clock64();
            float s[HT];
            tc::tmem_ld32(tmem_base + sb * FF_TILE + half * HT + lane_off, s);
            tc::tmem_ld_wait();
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&s_empty[sb]);
            if (valid < HT) {                              // warp-uniform: only the last tile of a video is ragged
#pragma unroll
                for (int j = 0; j < HT; ++j)
                    if (j >= valid) s[j] = -INFINITY;
            }
            float r4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
            for (int j = 0; j < HT; j += 4) {
                r4[0] = fmaxf(r4[0], s[j]); r4[1] = fmaxf(r4[1], s[j + 1]); r4[2] = fmaxf(r4[2], s[j + 2]); r4[3] = fmaxf(r4[3], s[j + 3]);
            }
            const float hmax = fmaxf(fmaxf(r4[0], r4[1]), fmaxf(r4[2], r4[3]));
            float* mxt = mxs + (t & 1) * 256;              // parity-double-buffered: the partner reads the slot of tile t while
            mxt[half * 128 + row] = hmax;                  // this thread may already write the one of tile t + 1
            asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");
            const float tmax = fmaxf(hmax, mxt[(half ^ 1) * 128 + row]) * LOG2E;
            const bool grow = tmax > mx + FF_RESCALE;      // (always true on the first tile: mx = -inf); identical in both halves
            const float mnew = grow ? tmax : mx;
            float a4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < HT; j += 4) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float e = ff_ex2(fmaf(s[j + u], LOG2E, -mnew));
                    s[j + u] = e;
                    a4[u] += e;
                }
            }
            const float rsum = (a4[0] + a4[1]) + (a4[2] + a4[3]);
            // the previous O += P X (and every earlier MMA) has completed once the P buffer is free again
            if (dbg_on && t < 12) p.dbg[8 + t * 4 + 2] = clock64();
            tc::mbar_wait(p_empty, (t & 1) ^ 1);
            tc::tc_fence_after();
            if (dbg_on && t < 12) p.dbg[8 + t * 4 + 3] = clock64();
            const float sc = ff_ex2(mx - mnew);            // 1 when the reference did not move, 0 on the first tile
            if (t > 0 && __any_sync(0xffffffffu, grow)) {  // this half rescales its 128 columns of O
#pragma unroll 1
                for (int c = half * (FF_H / 2); c < (half + 1) * (FF_H / 2); c += 32) {
                    float o[32];
                    tc::tmem_ld32(tmem_o + c + lane_off, o);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) o[j] *= sc;
                    ff_tmem_st32(tmem_o + c + lane_off, o);
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            }
            ls = ls * sc + rsum;
            mx = mnew;
            // P row -> bf16, this half's 4 chunks of 16 bytes, K-major 128B swizzle (chunk ^ (row % 8))
#pragma unroll
            for (int c = 0; c < 4; ++c)
                tc::sts_v4(prow + (uint32_t)(((half * 4 + c) ^ (row & 7)) << 4), tc::pack_bf16x2(s[c * 8], s[c * 8 + 1]),
                           tc::pack_bf16x2(s[c * 8 + 2], s[c * 8 + 3]), tc::pack_bf16x2(s[c * 8 + 4], s[c * 8 + 5]),
                           tc::pack_bf16x2(s[c * 8 + 6], s[c * 8 + 7]));
            const int valid_t = min(FF_TILE, nrows - t * FF_TILE);
            if (valid_t < FF_TILE) {
                // rows in [len, slot) of the tile may hold anything (0 x NaN would poison O): zero them in shared memory.  S of
                // this tile has completed, so the stage is idle until the P X issuer sees p_full.  A row is 128 bytes in each
                // of the 4 channel boxes; the swizzle only permutes 16-byte chunks inside a row.
                const uint32_t xst = tc::smem_u32(xs) + (uint32_t)((t % FF_NSTAGE) * FF_XBYTES);
                const int nz = (FF_TILE - valid_t) * 32;    // 16-byte chunks to clear: rows x 4 boxes x 8 chunks
                for (int i = (warp - 4) * 32 + lane; i < nz; i += 256) {
                    const int r = valid_t + (i >> 5), j = (i >> 3) & 3, c = i & 7;
                    tc::sts_v4(xst + (uint32_t)(j * FF_TILE * 128 + r * 128 + c * 16), 0u, 0u, 0u, 0u);
                }
            }
            tc::fence_proxy_async_smem();
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(p_full);
        }
        // epilogue: (max, sum) and the unnormalised O of every token of this split
        if (dbg_on) p.dbg[2] = clock64();
        mxs[512 + half * 128 + row] = ls;
        asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");
        const float lsum = ls + mxs[512 + (half ^ 1) * 128 + row];
        tc::mbar_wait(o_full, 0);
        tc::tc_fence_after();
        if (dbg_on) p.dbg[3] = clock64();
        const size_t prt0 = ((size_t)b * p.nsplit + split) * p.M;          // first token row of this split's partials
        if (half == 0 && row < p.M) *reinterpret_cast<float2*>(p.stats + (prt0 + row) * 2) = make_float2(mx, lsum);
        // every MMA and TMA load has completed: the row stages serve as per-warp [32][33] transposition buffers, so that the 32
        // floats a warp stores at a time are consecutive channels of ONE token (128-byte lines instead of 32 scattered 16-byte pieces)
        float* stg = reinterpret_cast<float*>(xs) + (warp - 4) * (32 * 33);
        const int nval = max(0, min(32, p.M - q * 32));
#pragma unroll 1
        for (int c = half * (FF_H / 2); c < (half + 1) * (FF_H / 2); c += 32) {
            float o[32];
            tc::tmem_ld32(tmem_o + c + lane_off, o);
            tc::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) stg[lane * 33 + j] = o[j];
            __syncwarp();
            float* dst = p.part + (prt0 + q * 32) * FF_H + c + lane;
            for (int r = 0; r < nval; ++r) dst[(size_t)r * FF_H] = stg[r * 33 + lane];
            __syncwarp();
        }
    }
    if (p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 128) p.dbg[4] = clock64();
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem_base, 512);
    }
}

long long* g_f2a_dbg = nullptr;

// out[b][m][e] = sum_s w_s O_s[m][e] / sum_s w_s l_s,  w_s = 2^(max_s - max over the splits); fixed split order.
__global__ void __launch_bounds__(256) f2a_combine_kernel(const float* __restrict__ stats, const float* __restrict__ part,
                                                          float* __restrict__ out, int ldo, int slot, const int32_t* __restrict__ len,
                                                          int M, int H, int nsplit) {
    const int b = blockIdx.y;
    const int len_b = len ? min(len[b], slot) : slot;
    const int ns = (len_b + FF_SPLIT - 1) / FF_SPLIT;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * H) return;
    const int m = i / H, e = i % H;
    float gm = -INFINITY;
    for (int s = 0; s < ns; ++s) gm = fmaxf(gm, stats[(((size_t)b * nsplit + s) * M + m) * 2]);
    float l = 0.f, a = 0.f;
    for (int s = 0; s < ns; ++s) {
        const size_t r = ((size_t)b * nsplit + s) * M + m;
        const float w = exp2f(stats[r * 2] - gm);
        l = fmaf(stats[r * 2 + 1], w, l);
        a = fmaf(part[r * H + e], w, a);
    }
    out[((size_t)b * M + m) * ldo + e] = ns > 0 ? a / l : 0.f;
}

}  // namespace factk

using namespace factk;

namespace factk {
// f2a_fused_t.cu: the transposed kernel (rows on the M side of both products): H = 512 and H = 256
bool f2a_fused_t_ok(int M, int H, int slot);
int f2a_fused_t_launch(const void* X, int ldx, const void* Qt, int ldq, long long qt_bstride, int B, int slot, const int32_t* len, int M, int H,
                       float* stats, float* part, cudaStream_t st);
}  // namespace factk

// FACTK_F2A_KERNEL=token forces the token-major kernel of this file where both serve the shape (H == 256), =rows the transposed one
static bool f2a_use_rows_kernel(int M, int H, int slot) {
    static const int mode = [] { const char* e = getenv("FACTK_F2A_KERNEL"); return !e ? 0 : (e[0] == 't' ? 1 : 2); }();
    const bool t_ok = f2a_fused_t_ok(M, H, slot), v_ok = H == FF_H && M <= 128 && (slot % FF_TILE) == 0;
    if (t_ok && v_ok) return mode == 2 || (mode == 0 && M <= 80);      // (more than 80 token columns spill registers in the rows kernel)
    return t_ok;
}

/* 1 when factk_f2a_fused serves the shapes: H == 256 with M <= 128, or H == 512 with M <= 80; slot a multiple of 128. */
extern "C" int factk_f2a_fused_supported(int M, int H, int slot) {
    return M >= 1 && (slot % 128) == 0 && ((H == FF_H && M <= 128) || f2a_fused_t_ok(M, H, slot));
}

extern "C" size_t factk_f2a_fused_ws_floats(int B, int slot, int M, int H) {
    const size_t ns = (size_t)(slot + FF_SPLIT - 1) / FF_SPLIT;
    return (size_t)B * ns * M * 2 + 4 + (size_t)B * ns * M * H;
}

extern "C" int factk_f2a_fused(const void* X, int ldx, const void* Qt, int ldq, long long qt_bstride, float* out, int ldo, int B, int slot,
                               const int32_t* len, int M, int H, float* ws, void* stream) {
    FACTK_REQUIRE(X && Qt && out && ws && B > 0, "factk_f2a_fused: bad args");
    FACTK_REQUIRE(factk_f2a_fused_supported(M, H, slot), "factk_f2a_fused: unsupported shape M=%d H=%d slot=%d", M, H, slot);
    FACTK_REQUIRE(ldx % 8 == 0 && ldq % 8 == 0 && qt_bstride % 8 == 0 && aligned16(X) && aligned16(Qt) && aligned16(ws),
                  "factk_f2a_fused: operands must be 16-byte aligned");
    const int ns = (slot + FF_SPLIT - 1) / FF_SPLIT;
    float* stats = ws;
    float* part = ws + (((size_t)B * ns * M * 2 + 3) & ~(size_t)3);
    cudaStream_t st = (cudaStream_t)stream;
    if (f2a_use_rows_kernel(M, H, slot)) {
        const int rc = f2a_fused_t_launch(X, ldx, Qt, ldq, qt_bstride, B, slot, len, M, H, stats, part, st);
        if (rc) return rc;
    } else {
        FfParams p;
        if (!tc_get_map(&p.qmap, Qt, 2, (uint64_t)H, (uint64_t)M, (uint64_t)B, (uint64_t)ldq, (uint64_t)qt_bstride, 128)) return FACTK_ERR_CUDA;
        if (!tc_get_map(&p.xmap, X, 2, (uint64_t)H, (uint64_t)slot, (uint64_t)B, (uint64_t)ldx, (uint64_t)slot * ldx, FF_TILE)) return FACTK_ERR_CUDA;
        p.M = M; p.slot = slot; p.nsplit = ns; p.len = len; p.dbg = g_f2a_dbg;
        p.stats = stats;
        p.part = part;
        static unsigned long long devs = 0;
        if (first_use_on_device(devs)) cudaFuncSetAttribute(f2a_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FF_SMEM);
        f2a_fused_kernel<<<dim3(ns, B), FF_THREADS, FF_SMEM, st>>>(p);
    }
    f2a_combine_kernel<<<dim3((M * H + 255) / 256, B), 256, 0, st>>>(stats, part, out, ldo, slot, len, M, H, ns);
    return check_launch("factk_f2a_fused");
}

/* development aid: clock64 timeline buffer (>= 8 + 4 * tiles int64) for the next launches; NULL switches it off */
extern "C" int factk_f2a_debug(long long* dbg) {
    factk::g_f2a_dbg = dbg;
    return FACTK_OK;
}
