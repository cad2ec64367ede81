// Fused dilated residual layer of the frame branch (models/basic.py:154-171, DilatedResidualLayer):
//
//     y[t] = x[t] + W1 . relu( sum_{k=0..2} W3[k] . x[t + (k-1) d] + b3 ) + b1          (eval: dropout = id)
//
// ONE persistent kernel per layer, no intermediate in HBM.  Per 128-frame tile:
//   GEMM 1  (K = 3F)  TMA boxes of the three shifted x tiles and of W3 -> tcgen05.mma -> D1 in TMEM
//   epi  1            tcgen05.ld D1, + b3, ReLU, bf16 -> shared memory in the UMMA 128B-swizzled K-major
//                     layout (the A operand of GEMM 2; never leaves the SM)
//   GEMM 2  (K = F)   A = that tile, B = W1 via TMA -> D2 in TMEM (re-uses D1's columns)
//   epi  2            tcgen05.ld D2, + b1, + x (residual, coalesced re-read, L2-hot), bf16 -> y
// Two accumulators ping-pong between consecutive tiles so GEMM 1 of tile i+1 runs under the epilogues of
// tile i.  CG = 2 runs the MMAs on a CTA pair (tcgen05 cta_group::2, M = 256): each CTA loads its own 128
// frames and HALF of every weight tile, which halves the weight traffic from L2 (the binding resource: the
// weights are re-streamed per tile) and the shared-memory footprint per stage.
//
// Warp roles (320 threads): warp 0 TMA producer, warp 1 MMA issuer (leader CTA only), warps 2..9 epilogue.
// All synchronisation is mbarrier based; waits are bounded (trap instead of hang).
#include <cstring>

#include "common.cuh"
#include "tc_common.cuh"

namespace factk {

struct TcnParams {
    alignas(64) CUtensorMap xmap;    // x  [B][slot][F] bf16, box 64 x 128 x 1
    alignas(64) CUtensorMap w3map;   // W3 [3][F][F]   bf16, box 64 x F/CG x 1
    alignas(64) CUtensorMap w1map;   // W1 [1][F][F]   bf16, box 64 x F/CG x 1
    const __nv_bfloat16* x;
    __nv_bfloat16* y;
    const float* b3;
    const float* b1;
    const int32_t* len;
    int B, slot, dil, tiles_m, total;
    long long* dbg;   // optional timeline of CTA 0 (tools/tcn_timeline.py): [tile][16] clock64 stamps
};

constexpr int TCN_EPI_WARPS = 8;
constexpr int TCN_THREADS = 64 + TCN_EPI_WARPS * 32;

template <int F, int CG>
struct TcnCfg {
    static constexpr int KCH = F / 64;            // 128-byte K chunks per tap (and of the 1x1)
    static constexpr int BROWS = F / CG;          // weight rows this CTA loads per chunk
    static constexpr int STAGE_A = 128 * 128;
    static constexpr int STAGE_B = BROWS * 128;
    static constexpr int STAGE = STAGE_A + STAGE_B;
    static constexpr int HBYTES = KCH * 16384;    // relu(conv3) tile: KCH chunks of [128 rows][128 B]
    static constexpr int BAR_BYTES = 256;
    static constexpr int LEN_CACHE = 320;         // video lengths cached in shared memory (ints)
    static constexpr int NSTAGE_RAW = (232448 - 1024 - BAR_BYTES - LEN_CACHE * 4 - HBYTES) / STAGE;
    static constexpr int NSTAGE = NSTAGE_RAW > 8 ? 8 : NSTAGE_RAW;
    static constexpr int SMEM = 1024 + NSTAGE * STAGE + HBYTES + BAR_BYTES + LEN_CACHE * 4;
    static_assert(NSTAGE >= 3, "pipeline too shallow");
    static_assert(2 * NSTAGE + 2 + 2 + 1 <= BAR_BYTES / 8 - 1, "barrier area");
};


// Walks the super tiles (CG x 128 frames of one video) owned by this CTA group, skipping tiles past the video end.
struct TcnTileIter {
    int st, step, tiles_m, total, slot, rows_per;
    const int32_t* len;
    __device__ TcnTileIter(const TcnParams& p, int unit, int nunits, int cg, const int32_t* len_)
        : st(unit - nunits), step(nunits), tiles_m(p.tiles_m), total(p.total), slot(p.slot), rows_per(128 * cg), len(len_) {}
    __device__ bool next(int& b, int& t0s, int& len_b) {
        while (true) {
            st += step;
            if (st >= total) return false;
            b = st / tiles_m;
            t0s = (st - b * tiles_m) * rows_per;
            len_b = len ? min(len[b], slot) : slot;
            if (t0s < len_b) return true;
        }
    }
};

template <int F, int CG>
__global__ void __launch_bounds__(TCN_THREADS, 1) tcn_layer_kernel(const __grid_constant__ TcnParams p) {
    using Cfg = TcnCfg<F, CG>;
    constexpr int KCH = Cfg::KCH;
    constexpr int NS = Cfg::NSTAGE;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* hbuf = smem + NS * Cfg::STAGE;
    uint64_t* full = reinterpret_cast<uint64_t*>(hbuf + Cfg::HBYTES);   // [NS]  leader: both CTAs' TMA bytes
    uint64_t* empty = full + NS;                                        // [NS]  each CTA: stage consumed
    uint64_t* tfull = empty + NS;                                       // [2]   each CTA: accumulator ready (D1, then D2)
    uint64_t* tempty = tfull + 2;                                       // [2]   leader: accumulator drained by epilogue 2
    uint64_t* hready = tempty + 2;                                      // [1]   leader: relu tile written, D1 drained
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(hready + 1);
    int32_t* slen = reinterpret_cast<int32_t*>(hbuf + Cfg::HBYTES + Cfg::BAR_BYTES);   // len[] cache: no global loads in the role loops

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = CG == 2 ? tc::cluster_ctarank() : 0u;
    const bool leader = rank == 0;
    const int unit = blockIdx.x / CG, nunits = gridDim.x / CG;

    const bool len_cached = p.len != nullptr && p.B <= Cfg::LEN_CACHE;
    if (len_cached)
        for (int i = threadIdx.x; i < p.B; i += TCN_THREADS) slen[i] = p.len[i];
    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&p.xmap);
        tc::tma_prefetch_desc(&p.w3map);
        tc::tma_prefetch_desc(&p.w1map);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int i = 0; i < NS; ++i) {
                tc::mbar_init(&full[i], 1);
                tc::mbar_init(&empty[i], 1);
            }
            for (int i = 0; i < 2; ++i) {
                tc::mbar_init(&tfull[i], 1);
                tc::mbar_init(&tempty[i], CG * TCN_EPI_WARPS);
            }
            tc::mbar_init(hready, CG * TCN_EPI_WARPS);
            tc::fence_barrier_init();
        }
        __syncwarp();
        tc::tmem_alloc_cg<CG>(tmem_slot, 2 * F);
        tc::tmem_relinquish_cg<CG>();
    }
    tc::tc_fence_before();
    if constexpr (CG == 2) tc::cluster_sync_all(); else __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t full0 = tc::mapa_u32(tc::smem_u32(&full[0]), 0);   // the leader's barriers
            auto load_g1 = [&](int b, int t0s) {
                const int t0 = t0s + (int)rank * 128;
#pragma unroll 1
                for (int tap = 0; tap < 3; ++tap) {
#pragma unroll 1
                    for (int kc = 0; kc < KCH; ++kc) {
                        tc::mbar_wait(&empty[stage], phase ^ 1);
                        if (leader) tc::mbar_arrive_expect_tx(&full[stage], CG * Cfg::STAGE);
                        const uint32_t st = tc::smem_u32(smem + stage * Cfg::STAGE);
                        tc::tma_load_3d_cg<CG>(st, &p.xmap, full0 + stage * 8, kc * 64, t0 + (tap - 1) * p.dil, b);
                        tc::tma_load_3d_cg<CG>(st + Cfg::STAGE_A, &p.w3map, full0 + stage * 8, kc * 64, (int)rank * Cfg::BROWS, tap);
                        if (++stage == NS) { stage = 0; phase ^= 1; }
                    }
                }
            };
            auto load_g2 = [&]() {
#pragma unroll 1
                for (int kc = 0; kc < KCH; ++kc) {
                    tc::mbar_wait(&empty[stage], phase ^ 1);
                    if (leader) tc::mbar_arrive_expect_tx(&full[stage], CG * Cfg::STAGE_B);
                    const uint32_t st = tc::smem_u32(smem + stage * Cfg::STAGE);
                    tc::tma_load_3d_cg<CG>(st + Cfg::STAGE_A, &p.w1map, full0 + stage * 8, kc * 64, (int)rank * Cfg::BROWS, 0);
                    if (++stage == NS) { stage = 0; phase ^= 1; }
                }
            };
            // tiles are processed in pairs (a, b): GEMM1(a) GEMM1(b) GEMM2(a) GEMM2(b)
            TcnTileIter iter(p, unit, nunits, CG, len_cached ? slen : p.len);
            int ba, ta, la, bb, tb, lb;
            while (iter.next(ba, ta, la)) {
                const bool two = iter.next(bb, tb, lb);
                load_g1(ba, ta);
                if (two) load_g1(bb, tb);
                load_g2();
                if (two) load_g2();
                if (!two) break;
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (leader CTA)
        if (leader && lane == 0) {
            constexpr uint32_t idesc = tc::instr_desc(128 * CG, F, false);
            int stage = 0;
            uint32_t phase = 0;
            const bool dbg = p.dbg != nullptr && blockIdx.x == 0;
            auto stamp = [&](int it, int slot_) { if (dbg && it < 64) p.dbg[it * 16 + slot_] = clock64(); };
            int n_h = 0;                      // relu tiles consumed so far (parity of the hready wait)
            auto gemm1 = [&](int it, int acc, int pair) {
                stamp(it, 0);
                if (pair >= 1) {
                    tc::mbar_wait(&tempty[acc], (uint32_t)((pair - 1) & 1));
                    tc::tc_fence_after();
                }
                const uint32_t d = tmem_base + acc * F;
                uint32_t accumulate = 0;
                stamp(it, 1);
#pragma unroll 1
                for (int c = 0; c < 3 * KCH; ++c) {
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
                stamp(it, 2);
            };
            auto gemm2 = [&](int it, int acc) {
                stamp(it, 3);
                tc::mbar_wait(hready, (uint32_t)(n_h & 1));
                ++n_h;
                tc::tc_fence_after();
                stamp(it, 4);
                const uint32_t d = tmem_base + acc * F;
                const uint32_t h_addr = tc::smem_u32(hbuf);
                uint32_t accumulate = 0;
#pragma unroll 1
                for (int kc = 0; kc < KCH; ++kc) {
                    tc::mbar_wait(&full[stage], phase);
                    tc::tc_fence_after();
                    const uint32_t a_addr = h_addr + kc * 16384;
                    const uint32_t b_addr = tc::smem_u32(smem + stage * Cfg::STAGE) + Cfg::STAGE_A;
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) {
                        tc::umma_bf16_cg<CG>(d, tc::smem_desc_sw128(a_addr + k4 * 32), tc::smem_desc_sw128(b_addr + k4 * 32), idesc, accumulate);
                        accumulate = 1;
                    }
                    tc::umma_commit_cg<CG>(&empty[stage]);
                    if (++stage == NS) { stage = 0; phase ^= 1; }
                }
                tc::umma_commit_cg<CG>(&tfull[acc]);
                stamp(it, 5);
            };
            TcnTileIter iter(p, unit, nunits, CG, len_cached ? slen : p.len);
            int b, t0s, len_b;
            int pair = 0;
            while (iter.next(b, t0s, len_b)) {
                const bool two = iter.next(b, t0s, len_b);
                gemm1(2 * pair, 0, pair);
                if (two) gemm1(2 * pair + 1, 1, pair);
                gemm2(2 * pair, 0);
                if (two) gemm2(2 * pair + 1, 1);
                ++pair;
                if (!two) break;
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue warps
        const int ew = warp - 2;
        const int q = warp & 3;              // TMEM lane quarter this warp may access
        const int half = ew >> 2;            // column half
        constexpr int HALF_COLS = F / 2;
        constexpr int NG = HALF_COLS / 32;   // 32-column groups per warp
        constexpr int NPC = HALF_COLS / 64;  // 128-byte chunks of a row owned by this warp (1 or 2)
        const int r = q * 32 + lane;         // tile row of this thread in the thread = row phases
        const uint32_t h_u32 = tc::smem_u32(hbuf);
        const uint32_t hready_l = tc::mapa_u32(tc::smem_u32(hready), 0);
        const uint32_t tempty_l = tc::mapa_u32(tc::smem_u32(&tempty[0]), 0);
        // shared-memory address of (row, column group): the UMMA K-major 128B-swizzled layout, chunk = 64 columns
        auto sw_addr = [&](int row, int col) -> uint32_t {
            return h_u32 + (uint32_t)(col >> 6) * 16384u + (uint32_t)row * 128u + (uint32_t)((((col & 63) >> 3) ^ (row & 7)) << 4);
        };
        const float* b3h = p.b3 + half * HALF_COLS;
        const bool dbg = p.dbg != nullptr && blockIdx.x == 0 && ew == 0 && lane == 0;
        auto stamp = [&](int it, int slot_) { if (dbg && it < 64) p.dbg[it * 16 + slot_] = clock64(); };
        // thread = row phases: address of 16-byte column group c (0 .. HALF_COLS/8-1) of this warp's half in row r
        const uint32_t my_row_addr = h_u32 + (uint32_t)((half * HALF_COLS) >> 6) * 16384u + (uint32_t)r * 128u;
        const uint32_t my_rx = (uint32_t)(r & 7) << 4;
        auto row_addr = [&](int c) -> uint32_t { return my_row_addr + (uint32_t)(c >> 3) * 16384u + ((uint32_t)((c & 7) << 4) ^ my_rx); };
        uint32_t tfc0 = 0u, tfc1 = 0u;       // uses of tfull[0] / tfull[1] so far (parity of the next wait)
        auto wait_tfull = [&](int acc) {
            const uint32_t par = (acc ? tfc1 : tfc0) & 1u;
            tc::mbar_wait(&tfull[acc], par);
            if (acc) ++tfc1; else ++tfc0;
            tc::tc_fence_after();
        };

        // epilogue 1: D1 -> + b3 -> ReLU -> bf16 (packed pairs, thread = row) in registers
        auto drain1 = [&](int it, int acc, uint32_t (&pk)[HALF_COLS / 2]) {
            stamp(it, 6);
            wait_tfull(acc);
            stamp(it, 7);
            const uint32_t taddr = tmem_base + acc * F + half * HALF_COLS + ((uint32_t)(q * 32) << 16);
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                const int col0 = half * HALF_COLS + g * 32;
                float v[32];
                tc::tmem_ld32(taddr + g * 32, v);
                float4 bq[8];
                const float* bp = b3h + g * 32;
                bq[0] = tc::ldg_v4f_ordered<0>(bp);  bq[1] = tc::ldg_v4f_ordered<16>(bp); bq[2] = tc::ldg_v4f_ordered<32>(bp);
                bq[3] = tc::ldg_v4f_ordered<48>(bp); bq[4] = tc::ldg_v4f_ordered<64>(bp); bq[5] = tc::ldg_v4f_ordered<80>(bp);
                bq[6] = tc::ldg_v4f_ordered<96>(bp); bq[7] = tc::ldg_v4f_ordered<112>(bp);
                tc::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    pk[g * 16 + 2 * j] = tc::pack_bf16x2(fmaxf(v[4 * j] + bq[j].x, 0.f), fmaxf(v[4 * j + 1] + bq[j].y, 0.f));
                    pk[g * 16 + 2 * j + 1] = tc::pack_bf16x2(fmaxf(v[4 * j + 2] + bq[j].z, 0.f), fmaxf(v[4 * j + 3] + bq[j].w, 0.f));
                }
            }
            tc::tc_fence_before();            // D1 fully read (GEMM 2 overwrites these columns)
        };
        // ... -> the A operand of GEMM 2 in shared memory (conflict-free 16-byte stores), then signal the MMA issuer
        auto store1 = [&](int it, const uint32_t (&pk)[HALF_COLS / 2]) {
#pragma unroll
            for (int c = 0; c < HALF_COLS / 8; ++c)
                tc::sts_v4(row_addr(c), pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
            tc::fence_proxy_async_smem();     // generic-proxy writes -> visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) tc::mbar_arrive_cluster(hready_l);
            stamp(it, 8);
        };

        // epilogue 2: D2 -> bf16 -> staged (thread = row) in this warp's own rows of the (dead) relu tile; the accumulator is
        // released as soon as it is drained.  Copy-out (lane = 16-byte column segment, coalesced): + b1 + x (residual) in
        // fp32 -> bf16 -> y.  The shared-memory data pipe is the scarce resource of this kernel (tensor-core operand
        // reads + TMA fills already use ~90% of it), so every global access here is fully coalesced and the staging is bf16.
        constexpr int LPR = 8 * NPC, RPI = 32 / LPR, NIT = 32 / RPI;   // lanes per row, rows per instruction, iterations
        constexpr int NBATCH = NIT / 8;
        const int lr = lane / LPR, lc = lane % LPR;
        const int ccol = half * HALF_COLS + lc * 8;                    // first of this lane's 8 columns in the copy-out
        float b1r[8];
        {
            const float4 t0_ = __ldg(reinterpret_cast<const float4*>(p.b1 + ccol)), t1_ = __ldg(reinterpret_cast<const float4*>(p.b1 + ccol + 4));
            b1r[0] = t0_.x; b1r[1] = t0_.y; b1r[2] = t0_.z; b1r[3] = t0_.w; b1r[4] = t1_.x; b1r[5] = t1_.y; b1r[6] = t1_.z; b1r[7] = t1_.w;
        }
        auto epi2 = [&](int it, int acc, int b, int t0s, int len_b, int wait_acc) {
            const int t0 = t0s + (int)rank * 128;
            const int rows_valid = len_b - (t0 + q * 32);
            const size_t rowbase = (size_t)b * p.slot + t0 + q * 32;
            const __nv_bfloat16* xcol = p.x + rowbase * F + ccol;
            __nv_bfloat16* ycol = p.y + rowbase * F + ccol;
            uint4 rres[8];
            auto fetch_res = [&](int batch) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int rr = (batch * 8 + i) * RPI + lr;
                    rres[i] = make_uint4(0u, 0u, 0u, 0u);
                    if (rr < rows_valid) {
                        const float4 t = tc::ldg_v4f_ordered<0>(xcol + (size_t)rr * F);
                        rres[i] = make_uint4(__float_as_uint(t.x), __float_as_uint(t.y), __float_as_uint(t.z), __float_as_uint(t.w));
                    }
                }
            };
            fetch_res(0);
            stamp(it, 9);
            if (wait_acc >= 0) wait_tfull(wait_acc);
            stamp(it, 10);
            const uint32_t taddr = tmem_base + acc * F + half * HALF_COLS + ((uint32_t)(q * 32) << 16);
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                float v[32];
                tc::tmem_ld32(taddr + g * 32, v);
                tc::tmem_ld_wait();
                if (g + 1 == NG) {
                    tc::tc_fence_before();    // accumulator fully read: hand it back to the MMA issuer
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive_cluster(tempty_l + (uint32_t)acc * 8u);
                    stamp(it, 11);
                }
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    tc::sts_v4(row_addr(g * 4 + c), tc::pack_bf16x2(v[c * 8], v[c * 8 + 1]), tc::pack_bf16x2(v[c * 8 + 2], v[c * 8 + 3]),
                               tc::pack_bf16x2(v[c * 8 + 4], v[c * 8 + 5]), tc::pack_bf16x2(v[c * 8 + 6], v[c * 8 + 7]));
            }
            __syncwarp();
#pragma unroll
            for (int batch = 0; batch < NBATCH; ++batch) {
                uint4 rcur[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) rcur[i] = rres[i];
                if (batch + 1 < NBATCH) fetch_res(batch + 1);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int rr = (batch * 8 + i) * RPI + lr;
                    uint4 o;
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(o.x), "=r"(o.y), "=r"(o.z), "=r"(o.w)
                                 : "r"(sw_addr(q * 32 + rr, ccol)) : "memory");
                    const uint32_t ow[4] = {o.x, o.y, o.z, o.w};
                    const uint32_t rw[4] = {rcur[i].x, rcur[i].y, rcur[i].z, rcur[i].w};
                    uint32_t yw[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float2 d = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ow[k]));
                        const float2 x2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&rw[k]));
                        yw[k] = tc::pack_bf16x2(d.x + b1r[2 * k] + x2.x, d.y + b1r[2 * k + 1] + x2.y);
                    }
                    if (rr < rows_valid) *reinterpret_cast<uint4*>(ycol + (size_t)rr * F) = make_uint4(yw[0], yw[1], yw[2], yw[3]);
                }
            }
            __syncwarp();                     // staged rows are read before the next relu tile overwrites them
            stamp(it, 12);
        };

        // Pair schedule (matches the MMA issuer): GEMM1(a) GEMM1(b) GEMM2(a) GEMM2(b).
        TcnTileIter iter(p, unit, nunits, CG, len_cached ? slen : p.len);
        int ba, ta, la, bb, tb, lb;
        int pair = 0;
        while (iter.next(ba, ta, la)) {
            const bool two = iter.next(bb, tb, lb);
            uint32_t pk[HALF_COLS / 2];
            drain1(2 * pair, 0, pk);          // D1(a): runs under GEMM1(b); the relu tile area is free (previous pair finished)
            store1(2 * pair, pk);
            if (two) drain1(2 * pair + 1, 1, pk);   // D1(b) -> registers while GEMM2(a) still reads the relu tile of a
            wait_tfull(0);                    // D2(a) complete: GEMM2(a) no longer reads shared memory
            if (two) store1(2 * pair + 1, pk);
            // the staging area (= relu tile area) is busy until GEMM2(b) completes
            epi2(2 * pair, 0, ba, ta, la, two ? 1 : -1);
            if (two) epi2(2 * pair + 1, 1, bb, tb, lb, -1);
            ++pair;
            if (!two) break;
        }
    }
    __syncwarp();

    tc::tc_fence_before();
    if constexpr (CG == 2) tc::cluster_sync_all(); else __syncthreads();
    if (warp == 1) {
        tc::tc_fence_after();
        tc::tmem_dealloc_cg<CG>(tmem_base, 2 * F);
    }
}

template <int F, int CG>
static int launch_tcn(const TcnParams& p, int sms, cudaStream_t st) {
    using Cfg = TcnCfg<F, CG>;
    static unsigned long long attr_devs = 0;
    if (first_use_on_device(attr_devs)) {
        cudaError_t e = cudaFuncSetAttribute(tcn_layer_kernel<F, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
        if (e != cudaSuccess) { set_error("factk_tcn_layer: smem attribute: %s", cudaGetErrorString(e)); return FACTK_ERR_CUDA; }
    }
    int units = sms / CG;
    if (units > p.total) units = p.total;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(units * CG, 1, 1);
    cfg.blockDim = dim3(TCN_THREADS, 1, 1);
    cfg.dynamicSmemBytes = Cfg::SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, tcn_layer_kernel<F, CG>, p);
    if (e != cudaSuccess) { set_error("factk_tcn_layer: launch: %s", cudaGetErrorString(e)); return FACTK_ERR_CUDA; }
    return check_launch("factk_tcn_layer");
}

}  // namespace factk

using namespace factk;

extern "C" int factk_tcn_layer_dbg(const void*, void*, const void*, const float*, const void*, const float*, int, int, int, int,
                                   const int32_t*, int, long long*, void*);
extern "C" int factk_tcn_layer_supported(int F) { return (F == 128 || F == 256) ? 1 : 0; }

extern "C" int factk_tcn_layer(const void* x, void* y, const void* w3, const float* b3, const void* w1, const float* b1, int B,
                               int slot, int F, int dilation, const int32_t* len, int cta_group, void* stream) {
    return factk_tcn_layer_dbg(x, y, w3, b3, w1, b1, B, slot, F, dilation, len, cta_group, nullptr, stream);
}

extern "C" int factk_tcn_layer_dbg(const void* x, void* y, const void* w3, const float* b3, const void* w1, const float* b1, int B,
                                   int slot, int F, int dilation, const int32_t* len, int cta_group, long long* dbg, void* stream) {
    FACTK_REQUIRE(x && y && w3 && b3 && w1 && b1 && B > 0 && slot > 0 && dilation > 0, "factk_tcn_layer: bad args");
    FACTK_REQUIRE(x != y, "factk_tcn_layer: in-place is not possible (the taps read neighbouring rows)");
    FACTK_REQUIRE(F == 128 || F == 256, "factk_tcn_layer: f_dim %d unsupported by the fused kernel (128 or 256)", F);
    FACTK_REQUIRE(aligned16(x) && aligned16(y) && aligned16(w3) && aligned16(w1) && aligned16(b3) && aligned16(b1),
                  "factk_tcn_layer: 16-byte alignment required");
    const int CG = cta_group == 1 ? 1 : 2;
    TcnParams p;
    memset(&p, 0, sizeof(p));
    if (!tc_get_map(&p.xmap, x, 2, F, slot, B, F, (uint64_t)slot * F, 128)) return FACTK_ERR_CUDA;
    if (!tc_get_map(&p.w3map, w3, 2, F, F, 3, F, (uint64_t)F * F, F / CG)) return FACTK_ERR_CUDA;
    if (!tc_get_map(&p.w1map, w1, 2, F, F, 1, F, (uint64_t)F * F, F / CG)) return FACTK_ERR_CUDA;
    p.x = reinterpret_cast<const __nv_bfloat16*>(x);
    p.y = reinterpret_cast<__nv_bfloat16*>(y);
    p.b3 = b3; p.b1 = b1; p.len = len;
    p.B = B; p.slot = slot; p.dil = dilation; p.dbg = dbg;
    p.tiles_m = (slot + 128 * CG - 1) / (128 * CG);
    p.total = B * p.tiles_m;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaStream_t st = (cudaStream_t)stream;
    if (F == 256) return CG == 2 ? launch_tcn<256, 2>(p, sms, st) : launch_tcn<256, 1>(p, sms, st);
    return CG == 2 ? launch_tcn<128, 2>(p, sms, st) : launch_tcn<128, 1>(p, sms, st);
}
