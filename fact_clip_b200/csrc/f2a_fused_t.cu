// X2Y_map, f2a direction (models/basic.py:349-389 with X = frame / segment rows, Y = action tokens) for hid_dim = 512 -- every
// shipped configuration -- as ONE tcgen05 / TMEM kernel in the TRANSPOSED orientation: the rows are the M side of both products,
// the few tokens are the N side (N = 80 for 75 tokens: no padded MMA lanes, no padded exp2 work).
//
//   S^T[128 rows x N]        = rows . qt^T                     A = the row tile, K-major (TMA boxes of 64 channels); B = qt, K-major
//   P^T                      = exp2(S^T log2e - ref[token])    one thread = one ROW; the softmax runs over the rows, i.e. ACROSS
//                                                              threads, so the reference of a token is shared by the CTA and lazy:
//                                                              it only moves when some logit exceeds it by more than 2^8 (a CTA-wide
//                                                              vote, bar.red.or); then the tile maximum is reduced (shuffles + shared
//                                                              memory) and the accumulators are rescaled per token column
//   O^T[H channels x N]     += rows^T . P^T                    A = the SAME rows MN-major (M = 128 channels per tile, K = rows),
//                                                              B = P^T MN-major (written by the softmax threads as bf16)
//
// O^T (H / 128 tiles of N columns) and two S^T buffers fit the 512 TMEM columns for H = 512, N <= 80 -- the token-major
// orientation of f2a_fused.cu needs 512 columns for O alone and a 128 KB query tile.  The row tile (128 x H bf16 = 128 KB at
// H = 512) does not stay resident: it streams through a 3-stage ring of [128 rows x 128 channels] twice per tile, once for
// S^T (from HBM) and once for O^T (from L2).  The normaliser is a per-thread running sum per token, reduced once at the end.
// Warp 0 = TMA producer, warp 1 = MMA issuer (order S(0) S(1) | O(0) S(2) | O(1) S(3) ...), warps 4-11 = softmax / epilogue: the
// two warps of a TMEM lane quarter split the token columns.  Rows in [len[b], slot) may hold anything: their logits are masked,
// and before the O^T pass of a ragged last tile the softmax warps zero them in the shared-memory stages (0 x NaN would poison O^T).
// Output: the same per-split partials as f2a_fused.cu (max in log2 units, sum, unnormalised O), merged by f2a_combine_kernel.
#include "common.cuh"
#include "tc_common.cuh"

namespace factk {

constexpr int FT_SPLIT = 2048;                       // rows per CTA (== FF_SPLIT of f2a_fused.cu: shared combine kernel)
constexpr int FT_TILE = 128;                         // rows per tile
constexpr int FT_BOX = FT_TILE * 128;                // one TMA box: [128 rows x 64 channels] bf16
constexpr int FT_STAGE = 2 * FT_BOX;                 // ring stage: 128 channels
constexpr int FT_NSTAGE = 3;
constexpr int FT_PBYTES = 2 * FT_BOX;                // P^T: [128 rows] x 2 atom columns of 64 tokens
constexpr int FT_SMALL = 6 * 128 * 4;                // ref[128], scale[128], per-quarter maxima / sums [4][128]
constexpr int FT_THREADS = 128 + 256;
constexpr float FT_RESCALE = 8.f;

struct FtParams {
    alignas(64) CUtensorMap qmap;    // qt   [B][M][H]    bf16, box 64 x N x 1 (token rows >= M arrive as zeros)
    alignas(64) CUtensorMap xmap;    // rows [B][slot][H] bf16, box 64 x 128 x 1
    int M, H, slot, nsplit;
    const int32_t* len;
    float* stats;                    // [B][nsplit][M][2] = (reference in log2 units, sum)
    float* part;                     // [B][nsplit][M][H] unnormalised weighted row sums
    long long* dbg;                  // optional clock64 timeline of CTA (0,0), softmax warp 4 lane 0 (development aid)
};

constexpr uint32_t FT_DESC_HI = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);       // SBO 1024 | version 1 | SWIZZLE_128B
__device__ __forceinline__ uint32_t ft_desc_lo_k(uint32_t saddr) { return ((saddr >> 4) & 0x3FFFu) | (1u << 16); }
__device__ __forceinline__ uint32_t ft_desc_lo_mn(uint32_t saddr) { return ((saddr >> 4) & 0x3FFFu) | ((uint32_t)(FT_BOX >> 4) << 16); }
__device__ __forceinline__ void ft_umma(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\tsetp.ne.b32 p, %5, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
        ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(FT_DESC_HI), "r"(idesc), "r"(acc)
        : "memory");
}
// 32 lanes x 8 consecutive fp32 columns
__device__ __forceinline__ void ft_tmem_ld8(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void ft_tmem_st8(uint32_t taddr, const float* v) {
    const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ float ft_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// CTA-wide OR over the 256 softmax threads (named barrier 2)
__device__ __forceinline__ bool ft_any256(bool pred) {
    uint32_t r;
    asm volatile(
        "{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %1, 0;\n\t"
        "bar.red.or.pred q, 2, 256, p;\n\tselp.u32 %0, 1, 0, q;\n\t}"
        : "=r"(r)
        : "r"((uint32_t)pred)
        : "memory");
    return r != 0;
}
__device__ __forceinline__ void ft_bar256() { asm volatile("bar.sync 3, 256;" ::: "memory"); }

constexpr int ft_smem_bytes(int N, int H) { return (H / 64) * N * 128 + FT_NSTAGE * FT_STAGE + FT_PBYTES + FT_SMALL + 256 + 1024; }

template <int N>     // token columns (a multiple of 16): 75 tokens -> 80
__global__ void __launch_bounds__(FT_THREADS, 1) f2a_fused_t_kernel(const __grid_constant__ FtParams p) {
    constexpr int NH = N / 2;                            // token columns per softmax thread
    static_assert(N % 16 == 0 && N <= 128, "token columns");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int KCH = p.H / 64, MT = p.H / 128;
    uint8_t* qs = smem;                                  // KCH boxes [N tokens x 64 channels], K-major
    uint8_t* ring = qs + KCH * N * 128;
    uint8_t* pt = ring + FT_NSTAGE * FT_STAGE;
    float* ref = reinterpret_cast<float*>(pt + FT_PBYTES);          // reference maximum per token (log2 units)
    float* scs = ref + 128;                                          // rescale factors of the current slow path
    float* wq = scs + 128;                                           // [4 lane quarters][128]: tile maxima, final sums
    uint64_t* bars = reinterpret_cast<uint64_t*>(pt + FT_PBYTES + FT_SMALL);
    uint64_t *full = bars, *empty = bars + FT_NSTAGE, *s_full = bars + 2 * FT_NSTAGE, *s_empty = s_full + 2, *p_full = s_full + 4,
             *p_empty = s_full + 5, *q_full = s_full + 6, *o_full = s_full + 7, *clean = s_full + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 12);      // (clean[0..3]: one single-use barrier per stage of the pass)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long t_entry = clock64();
    const int split = blockIdx.x, b = blockIdx.y;
    const int len_b = p.len ? min(p.len[b], p.slot) : p.slot;
    const int r0 = split * FT_SPLIT;
    if (r0 >= len_b) return;                             // uniform per CTA: nothing allocated yet
    const int nrows = min(r0 + FT_SPLIT, len_b) - r0;
    const int ntile = (nrows + FT_TILE - 1) / FT_TILE;
    const int last_valid = nrows - (ntile - 1) * FT_TILE;      // rows of the last tile below len[b]; < 128: the tile is ragged

    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&p.qmap);
        tc::tma_prefetch_desc(&p.xmap);
        for (int i = 0; i < FT_NSTAGE; ++i) {
            tc::mbar_init(&full[i], 1);
            tc::mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(&s_full[i], 1);
            tc::mbar_init(&s_empty[i], 8);
        }
        tc::mbar_init(p_full, 8);
        tc::mbar_init(p_empty, 1);
        tc::mbar_init(q_full, 1);
        tc::mbar_init(o_full, 1);
        for (int i = 0; i < 4; ++i) tc::mbar_init(&clean[i], 8);
        tc::fence_barrier_init();
    }
    if (threadIdx.x >= 128 && threadIdx.x < 256) ref[threadIdx.x - 128] = -INFINITY;
    __syncthreads();                                     // barriers initialised; the TMA warp runs ahead from here
    uint32_t tmem_base = 0;
    if (warp >= 1) {
        if (warp == 1) {
            tc::tmem_alloc(tmem_slot, 512);
            tc::tmem_relinquish();
        }
        tc::tc_fence_before();
        asm volatile("bar.sync 1, %0;" ::"n"(FT_THREADS - 32) : "memory");
        tc::tc_fence_after();
        tmem_base = *tmem_slot;
    }
    const uint32_t tmem_o = tmem_base + 2 * N;           // S^T buffers: columns [0, 2N); O^T tile j: 2N + jN

    if (warp == 0) {
        if (lane == 0) {
            tc::mbar_arrive_expect_tx(q_full, (uint32_t)(KCH * N * 128));
            for (int j = 0; j < KCH; ++j) tc::tma_load_3d(qs + j * (N * 128), &p.qmap, q_full, j * 64, 0, b);
            int stage = 0;
            uint32_t phase = 0;
            auto load_tile = [&](int t) {                // the MT stages of tile t (for S^T or for O^T: the same boxes)
                for (int j = 0; j < MT; ++j) {
                    tc::mbar_wait(&empty[stage], phase ^ 1);
                    tc::mbar_arrive_expect_tx(&full[stage], FT_STAGE);
                    uint8_t* st = ring + stage * FT_STAGE;
                    tc::tma_load_3d(st, &p.xmap, &full[stage], j * 128, r0 + t * FT_TILE, b);
                    tc::tma_load_3d(st + FT_BOX, &p.xmap, &full[stage], j * 128 + 64, r0 + t * FT_TILE, b);
                    if (++stage == FT_NSTAGE) { stage = 0; phase ^= 1; }
                }
            };
            load_tile(0);
            if (ntile > 1) load_tile(1);
            for (int t = 0; t < ntile; ++t) {
                load_tile(t);
                if (t + 2 < ntile) load_tile(t + 2);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc_s = tc::instr_desc(128, N, false);
            constexpr uint32_t idesc_o = tc::instr_desc(128, N, false) | (1u << 15) | (1u << 16);      // A and B MN-major
            const uint32_t q_lo = ft_desc_lo_k(tc::smem_u32(qs)), xk_lo = ft_desc_lo_k(tc::smem_u32(ring));
            const uint32_t xm_lo = ft_desc_lo_mn(tc::smem_u32(ring)), p_lo = ft_desc_lo_mn(tc::smem_u32(pt));
            int stage = 0;
            uint32_t phase = 0;
            auto s_tile = [&](int t) {                   // S^T[t & 1] = X_t qt^T over the MT stages (2 boxes x 4 k-steps each)
                const int sb = t & 1;
                tc::mbar_wait(&s_empty[sb], ((t >> 1) & 1) ^ 1);
                tc::tc_fence_after();
                for (int j = 0; j < MT; ++j) {
                    tc::mbar_wait(&full[stage], phase);
                    tc::tc_fence_after();
                    const uint32_t xa = xk_lo + (uint32_t)((stage * FT_STAGE) >> 4);
#pragma unroll
                    for (int bx = 0; bx < 2; ++bx)
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4)
                            ft_umma(tmem_base + sb * N, xa + (uint32_t)((bx * FT_BOX + k4 * 32) >> 4),
                                    q_lo + (uint32_t)(((2 * j + bx) * (N * 128) + k4 * 32) >> 4), idesc_s, (j | bx | k4) ? 1u : 0u);
                    tc::umma_commit(&empty[stage]);
                    if (++stage == FT_NSTAGE) { stage = 0; phase ^= 1; }
                }
                tc::umma_commit(&s_full[sb]);
            };
            auto o_tile = [&](int t) {                   // O^T_j += X_t[:, 128 j ..]^T P^T: 8 k-steps of 16 rows per channel tile
                tc::mbar_wait(p_full, t & 1);
                tc::tc_fence_after();
                const bool ragged = t == ntile - 1 && last_valid < FT_TILE;
                for (int j = 0; j < MT; ++j) {
                    tc::mbar_wait(&full[stage], phase);
                    if (ragged) tc::mbar_wait(&clean[j], 0);           // rows past len[b] of this stage zeroed by the softmax warps
                    tc::tc_fence_after();
                    const uint32_t xa = xm_lo + (uint32_t)((stage * FT_STAGE) >> 4);
#pragma unroll
                    for (int k = 0; k < FT_TILE / 16; ++k)
                        ft_umma(tmem_o + j * N, xa + k * 128, p_lo + k * 128, idesc_o, (t > 0 || k > 0) ? 1u : 0u);      // + 2048 B
                    tc::umma_commit(&empty[stage]);
                    if (++stage == FT_NSTAGE) { stage = 0; phase ^= 1; }
                }
                tc::umma_commit(p_empty);
            };
            tc::mbar_wait(q_full, 0);
            s_tile(0);
            if (ntile > 1) s_tile(1);
            for (int t = 0; t < ntile; ++t) {
                o_tile(t);
                if (t + 2 < ntile) s_tile(t + 2);
            }
            tc::umma_commit(o_full);
        }
    } else if (warp >= 4) {
        const int q = warp & 3, half = (warp - 4) >> 2;
        const int f = q * 32 + lane;                       // row of the tile (TMEM lane)
        const int c0 = half * NH;                          // first token column of this thread
        const int tsm = threadIdx.x - 128;                 // 0..255 among the softmax threads
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const uint32_t prow = tc::smem_u32(pt) + (uint32_t)(f >> 3) * 1024u + (uint32_t)(f & 7) * 128u;
        constexpr float LOG2E = 1.4426950408889634f;
        float l[NH];                                       // running sums of this thread's rows, per token of its half
#pragma unroll
        for (int c = 0; c < NH; ++c) l[c] = 0.f;
        const bool dbg_on = p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && warp == 4 && lane == 0;
        if (dbg_on) { p.dbg[0] = t_entry; p.dbg[1] = clock64(); }
        for (int t = 0; t < ntile; ++t) {
            const int sb = t & 1;
            const bool live = t * FT_TILE + f < nrows;
            if (dbg_on && t < 12) p.dbg[8 + t * 4 + 0] = clock64();
            tc::mbar_wait(&s_full[sb], (t >> 1) & 1);
            tc::tc_fence_after();
            if (dbg_on && t < 12) p.dbg[8 + t * 4 + 1] = clock64();
            float x[NH];
#pragma unroll
            for (int g = 0; g < NH / 8; ++g) ft_tmem_ld8(tmem_base + sb * N + c0 + g * 8 + lane_off, x + g * 8);
            tc::tmem_ld_wait();
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&s_empty[sb]);
            float ex = -INFINITY;                          // largest excess over the reference
#pragma unroll
            for (int c = 0; c < NH; c += 4) {
                const float4 r4 = *reinterpret_cast<const float4*>(ref + c0 + c);      // broadcast reads
                x[c] = live ? x[c] * LOG2E : -INFINITY;
                x[c + 1] = live ? x[c + 1] * LOG2E : -INFINITY;
                x[c + 2] = live ? x[c + 2] * LOG2E : -INFINITY;
                x[c + 3] = live ? x[c + 3] * LOG2E : -INFINITY;
                // (reference -inf on the first tile: +inf for live rows; -inf - -inf = NaN for dead ones, which fmaxf drops)
                ex = fmaxf(fmaxf(ex, x[c] - r4.x), fmaxf(x[c + 1] - r4.y, fmaxf(x[c + 2] - r4.z, x[c + 3] - r4.w)));
            }
            const bool slow = ft_any256(ex > FT_RESCALE);
            if (slow) {                                    // CTA-uniform: move the references to the maxima seen so far
#pragma unroll
                for (int c = 0; c < NH; ++c) {
                    const float m = warp_max(x[c]);
                    if (lane == 0) wq[q * 128 + c0 + c] = m;
                }
                ft_bar256();
                if (tsm < N) {
                    const float old = ref[tsm];
                    const float mt = fmaxf(fmaxf(wq[tsm], wq[128 + tsm]), fmaxf(wq[256 + tsm], wq[384 + tsm]));
                    const float mnew = fmaxf(old, mt);
                    scs[tsm] = ft_ex2(old - mnew);         // 0 on the first tile
                    ref[tsm] = mnew;
                }
                ft_bar256();
#pragma unroll
                for (int c = 0; c < NH; ++c) l[c] *= scs[c0 + c];
            }
#pragma unroll
            for (int c = 0; c < NH; c += 4) {
                const float4 r4 = *reinterpret_cast<const float4*>(ref + c0 + c);
                x[c] = ft_ex2(x[c] - r4.x);                // rows past the end: exp2(-inf) = 0
                x[c + 1] = ft_ex2(x[c + 1] - r4.y);
                x[c + 2] = ft_ex2(x[c + 2] - r4.z);
                x[c + 3] = ft_ex2(x[c + 3] - r4.w);
                l[c] += x[c]; l[c + 1] += x[c + 1]; l[c + 2] += x[c + 2]; l[c + 3] += x[c + 3];
            }
            // the previous O^T += X^T P^T (and every earlier MMA) has completed once the P^T buffer is free again
            if (dbg_on && t < 12) p.dbg[8 + t * 4 + 2] = clock64();
            tc::mbar_wait(p_empty, (t & 1) ^ 1);
            tc::tc_fence_after();
            if (dbg_on && t < 12) p.dbg[8 + t * 4 + 3] = clock64();
            if (slow && t > 0) {                           // rescale this thread's share of O^T: lane quarter q, its token columns
                for (int j = 0; j < MT; ++j) {
#pragma unroll
                    for (int g = 0; g < NH / 8; ++g) {
                        float o[8];
                        ft_tmem_ld8(tmem_o + j * N + c0 + g * 8 + lane_off, o);
                        tc::tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 8; ++i) o[i] *= scs[c0 + g * 8 + i];
                        ft_tmem_st8(tmem_o + j * N + c0 + g * 8 + lane_off, o);
                    }
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            }
            // P^T row -> bf16, MN-major 128B swizzle: 16-byte chunk G = 8 tokens, atom column G / 8, position (G % 8) ^ (row % 8)
#pragma unroll
            for (int g = 0; g < NH / 8; ++g) {
                const int G = c0 / 8 + g;
                tc::sts_v4(prow + (uint32_t)(G >> 3) * (uint32_t)FT_BOX + (uint32_t)(((G & 7) ^ (f & 7)) << 4),
                           tc::pack_bf16x2(x[g * 8], x[g * 8 + 1]), tc::pack_bf16x2(x[g * 8 + 2], x[g * 8 + 3]),
                           tc::pack_bf16x2(x[g * 8 + 4], x[g * 8 + 5]), tc::pack_bf16x2(x[g * 8 + 6], x[g * 8 + 7]));
            }
            tc::fence_proxy_async_smem();
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(p_full);
        }
        if (dbg_on) p.dbg[2] = clock64();
        if (last_valid < FT_TILE) {
            // Ragged last tile: its rows in [len[b], slot) have exact-zero probabilities, but they may hold anything (0 x NaN would
            // poison O^T), so the stages of the LAST O^T pass are cleaned in shared memory between the TMA arrival and the MMAs.
            // That pass is ring pass number 2 ntile - 1 (all S^T passes and the other O^T passes come before it).
            const int first = (2 * ntile - 1) * MT;
            for (int j = 0; j < MT; ++j) {
                const int g = first + j, stage = g % FT_NSTAGE;
                tc::mbar_wait(&full[stage], (uint32_t)(g / FT_NSTAGE) & 1u);
                const uint32_t st = tc::smem_u32(ring) + (uint32_t)(stage * FT_STAGE);
                const int nz = (FT_TILE - last_valid) * 16;             // 16-byte chunks: rows x 2 boxes x 8 (a row is a 128-byte line)
                for (int i = tsm; i < nz; i += 256) {
                    const int r = last_valid + (i >> 4), bx = (i >> 3) & 1, c = i & 7;
                    tc::sts_v4(st + (uint32_t)(bx * FT_BOX + r * 128 + c * 16), 0u, 0u, 0u, 0u);
                }
                tc::fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&clean[j]);      // (one barrier per stage: a phase-counted one could run two phases ahead of the issuer)
            }
        }
        // normaliser: sum over the rows = over lanes, quarters (fixed order)
#pragma unroll
        for (int c = 0; c < NH; ++c) {
            const float v = warp_sum(l[c]);
            if (lane == 0) wq[q * 128 + c0 + c] = v;
        }
        ft_bar256();
        const size_t prt0 = ((size_t)b * p.nsplit + split) * p.M;          // first token row of this split's partials
        if (tsm < p.M)
            *reinterpret_cast<float2*>(p.stats + (prt0 + tsm) * 2) = make_float2(ref[tsm], ((wq[tsm] + wq[128 + tsm]) + wq[256 + tsm]) + wq[384 + tsm]);
        // O^T: thread = channel, columns = tokens: the 32 lanes of a warp write 32 consecutive channels of one token
        tc::mbar_wait(o_full, 0);
        tc::tc_fence_after();
        if (dbg_on) p.dbg[3] = clock64();
        for (int j = 0; j < MT; ++j) {
            float o[NH];
#pragma unroll
            for (int g = 0; g < NH / 8; ++g) ft_tmem_ld8(tmem_o + j * N + c0 + g * 8 + lane_off, o + g * 8);
            tc::tmem_ld_wait();
            float* dst = p.part + prt0 * p.H + j * 128 + f;
#pragma unroll
            for (int c = 0; c < NH; ++c)
                if (c0 + c < p.M) dst[(size_t)(c0 + c) * p.H] = o[c];
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

extern long long* g_f2a_dbg;     // f2a_fused.cu (factk_f2a_debug)

static int ft_cols(int M) {
    const int opts[] = {32, 64, 80, 96, 128};
    for (int n : opts)
        if (M <= n) return n;
    return 0;
}

// 1 when the transposed kernel serves the shape: H in {256, 512}, token columns N with (H / 128 + 2) N <= 512 TMEM columns.
bool f2a_fused_t_ok(int M, int H, int slot) {
    const int N = ft_cols(M);
    return (H == 256 || H == 512) && M >= 1 && N > 0 && (H / 128 + 2) * N <= 512 && (slot % 128) == 0 && ft_smem_bytes(N, H) <= 232448;
}

template <int N>
static void ft_launch(const FtParams& p, int B, cudaStream_t st) {
    const int smem = ft_smem_bytes(N, p.H);
    static unsigned long long devs = 0;
    if (first_use_on_device(devs)) cudaFuncSetAttribute(f2a_fused_t_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    f2a_fused_t_kernel<N><<<dim3(p.nsplit, B), FT_THREADS, smem, st>>>(p);
}

int f2a_fused_t_launch(const void* X, int ldx, const void* Qt, int ldq, long long qt_bstride, int B, int slot, const int32_t* len, int M, int H,
                       float* stats, float* part, cudaStream_t st) {
    FtParams p;
    const int N = ft_cols(M);
    const int ns = (slot + FT_SPLIT - 1) / FT_SPLIT;
    if (!tc_get_map(&p.qmap, Qt, 2, (uint64_t)H, (uint64_t)M, (uint64_t)B, (uint64_t)ldq, (uint64_t)qt_bstride, (uint32_t)N)) return FACTK_ERR_CUDA;
    if (!tc_get_map(&p.xmap, X, 2, (uint64_t)H, (uint64_t)slot, (uint64_t)B, (uint64_t)ldx, (uint64_t)slot * ldx, FT_TILE)) return FACTK_ERR_CUDA;
    p.M = M; p.H = H; p.slot = slot; p.nsplit = ns; p.len = len; p.stats = stats; p.part = part; p.dbg = g_f2a_dbg;
    switch (N) {
        case 32: ft_launch<32>(p, B, st); break;
        case 64: ft_launch<64>(p, B, st); break;
        case 80: ft_launch<80>(p, B, st); break;
        case 96: ft_launch<96>(p, B, st); break;
        default: ft_launch<128>(p, B, st); break;
    }
    return FACTK_OK;
}

}  // namespace factk
