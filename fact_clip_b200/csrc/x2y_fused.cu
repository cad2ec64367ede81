// X2Y_map in the a2f direction (models/basic.py:349-389 with X = action tokens, Y = frame / segment rows; blocks.py:352,463)
// as ONE tcgen05 / TMEM kernel per 128-row tile -- "fused attention" with the rows as queries and the few tokens as keys/values:
//
//   S[128 x M]    = rows . kt^T + cb            kt = alpha Wq^T X_K(tokens + pos) (token-side fold, engine.py), cb = alpha xk . bq
//   P             = softmax_M(S)                one thread = one row of S in registers (tcgen05.ld), masking of padded tokens by -inf
//   out[128 x F]  = rows . Wy^T + P . vt^T + b  vt = (Wa Wv) tokens^T (token-side fold): Y_W(cat[Y, attn . X_V(tokens)])
//
// The rows are read ONCE: every 64-channel chunk of the tile feeds two MMAs, S += X_k kt_k^T (N = M padded to 16) and
// D += X_k Wy_k^T (N = F = 256), both accumulating in tensor memory; after the last chunk the four epilogue warps turn S into
// probabilities, write them as bf16 into the 128B-swizzled K-major A-operand layout in shared memory (and, when asked, the
// fp32 logits and attention rows to HBM for the loss / eval fusion), the MMA warp adds P . vt^T to D, and the epilogue writes the
// bf16 output rows.  Nothing else touches HBM: the unfused path wrote and re-read the fp32 logits, the fp32 attention and a
// bf16 copy of it ([B, slot, M] each) in three launches.
//
// Warp 0 = TMA producer (3-stage ring: X chunk 16 KB + kt chunk <= 16 KB + Wy chunk 32 KB; the two vt boxes reuse ring slots),
// warp 1 = MMA issuer, warps 2-5 = epilogue (TMEM lane quarter = warp % 4).  Persistent over tiles.
#include "common.cuh"
#include "tc_common.cuh"

namespace factk {

constexpr int XF_F = 256;                         // output width (f_dim)
constexpr int XF_A = 128 * 128, XF_B1 = 128 * 128, XF_B2 = 256 * 128;
constexpr int XF_STAGE = XF_A + XF_B1 + XF_B2;    // 64 KB
constexpr int XF_NSTAGE = 3;
constexpr int XF_PBYTES = 2 * 128 * 128;          // P tile: 2 boxes [128 rows x 64 tokens] bf16
constexpr int XF_SMEM = XF_NSTAGE * XF_STAGE + XF_PBYTES + 1024 + 256 + 1024;
constexpr int XF_THREADS = 192;
static_assert(XF_SMEM <= 232448, "dynamic shared memory limit");

struct XfParams {
    alignas(64) CUtensorMap xmap, kmap, wmap, vmap;
    const float* cb;
    long long cb_bstride;
    const float* bias;
    __nv_bfloat16* out;
    int ldo;
    float* logit;
    float* attn;
    int ldl;
    int B, slot, M, N1, kchunks, pboxes, tiles_m, total_tiles;
    const int32_t* len;
};

__global__ void __launch_bounds__(XF_THREADS, 1) a2f_fused_kernel(const __grid_constant__ XfParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* ring = smem;
    uint8_t* ptile = ring + XF_NSTAGE * XF_STAGE;
    float* cbs = reinterpret_cast<float*>(ptile + XF_PBYTES);            // 128 floats (+ 128 bias... kept in registers) -> 1 KB
    uint64_t* bars = reinterpret_cast<uint64_t*>(ptile + XF_PBYTES + 1024);
    uint64_t *full = bars, *empty = bars + XF_NSTAGE, *s_full = bars + 2 * XF_NSTAGE, *p_full = s_full + 1, *d_full = s_full + 2,
             *s_free = s_full + 3, *d_free = s_full + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 5);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&p.xmap); tc::tma_prefetch_desc(&p.kmap); tc::tma_prefetch_desc(&p.wmap); tc::tma_prefetch_desc(&p.vmap);
        for (int i = 0; i < XF_NSTAGE; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
        tc::mbar_init(s_full, 1); tc::mbar_init(p_full, 4); tc::mbar_init(d_full, 1);
        tc::mbar_init(s_free, 4); tc::mbar_init(d_free, 4);
        tc::fence_barrier_init();
    }
    __syncthreads();
    if (warp == 1) {
        tc::tmem_alloc(tmem_slot, 512);
        tc::tmem_relinquish();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_s = tmem_base, tmem_d = tmem_base + 128;      // S: columns [0, 128); D: [128, 384)

    auto tile_coords = [&](int tile, int& b, int& t0, int& len_b) {
        b = tile / p.tiles_m;
        t0 = (tile % p.tiles_m) * 128;
        len_b = p.len ? min(p.len[b], p.slot) : p.slot;
        return t0 < len_b;
    };
    const int nst = p.kchunks + p.pboxes;          // ring stages per tile: K chunks, then the vt boxes

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                int b, t0, len_b;
                if (!tile_coords(tile, b, t0, len_b)) continue;
                for (int s = 0; s < nst; ++s) {
                    tc::mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* st = ring + stage * XF_STAGE;
                    if (s < p.kchunks) {
                        tc::mbar_arrive_expect_tx(&full[stage], XF_A + p.N1 * 128 + XF_B2);
                        tc::tma_load_3d(st, &p.xmap, &full[stage], s * 64, t0, b);
                        tc::tma_load_3d(st + XF_A, &p.kmap, &full[stage], s * 64, 0, b);
                        tc::tma_load_3d(st + XF_A + XF_B1, &p.wmap, &full[stage], s * 64, 0, 0);
                    } else {
                        tc::mbar_arrive_expect_tx(&full[stage], XF_B2);
                        tc::tma_load_3d(st + XF_A + XF_B1, &p.vmap, &full[stage], (s - p.kchunks) * 64, 0, b);
                    }
                    if (++stage == XF_NSTAGE) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc_s = tc::instr_desc(128, p.N1, false);
            constexpr uint32_t idesc_d = tc::instr_desc(128, XF_F, false);
            int stage = 0, it = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                int b, t0, len_b;
                if (!tile_coords(tile, b, t0, len_b)) continue;
                const uint32_t tpar = it & 1;
                ++it;
                tc::mbar_wait(s_free, tpar ^ 1);             // the previous tile's logits have been read out of TMEM
                tc::mbar_wait(d_free, tpar ^ 1);             // ... and its output accumulator
                tc::tc_fence_after();
                for (int s = 0; s < p.kchunks; ++s) {
                    tc::mbar_wait(&full[stage], phase);
                    tc::tc_fence_after();
                    const uint32_t a_addr = tc::smem_u32(ring + stage * XF_STAGE);
                    const uint32_t b1 = a_addr + XF_A, b2 = b1 + XF_B1;
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) {
                        const uint32_t acc = (s > 0 || k4 > 0) ? 1u : 0u;
                        tc::umma<false>(tmem_s, tc::smem_desc_sw128(a_addr + k4 * 32), tc::smem_desc_sw128(b1 + k4 * 32), idesc_s, acc);
                        tc::umma<false>(tmem_d, tc::smem_desc_sw128(a_addr + k4 * 32), tc::smem_desc_sw128(b2 + k4 * 32), idesc_d, acc);
                    }
                    tc::umma_commit(&empty[stage]);
                    if (s == p.kchunks - 1) tc::umma_commit(s_full);
                    if (++stage == XF_NSTAGE) { stage = 0; phase ^= 1; }
                }
                tc::mbar_wait(p_full, tpar);                 // probabilities in shared memory
                tc::tc_fence_after();
                const uint32_t p_addr = tc::smem_u32(ptile);
                for (int j = 0; j < p.pboxes; ++j) {
                    tc::mbar_wait(&full[stage], phase);
                    tc::tc_fence_after();
                    const uint32_t v_addr = tc::smem_u32(ring + stage * XF_STAGE) + XF_A + XF_B1;
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4)
                        tc::umma<false>(tmem_d, tc::smem_desc_sw128(p_addr + j * 16384 + k4 * 32), tc::smem_desc_sw128(v_addr + k4 * 32), idesc_d, 1u);
                    tc::umma_commit(&empty[stage]);
                    if (++stage == XF_NSTAGE) { stage = 0; phase ^= 1; }
                }
                tc::umma_commit(d_full);
            }
        }
    } else {
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const uint32_t p_addr = tc::smem_u32(ptile);
        int it = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
            int b, t0, len_b;
            if (!tile_coords(tile, b, t0, len_b)) continue;
            const uint32_t tpar = it & 1;
            ++it;
            const int t = t0 + row;
            const bool live = t < len_b;
            const size_t grow = (size_t)b * p.slot + t;
            // per-token logit bias of this video (the epilogue warps only synchronise among themselves: named barrier 1)
            asm volatile("bar.sync 1, 128;" ::: "memory");
            for (int i = threadIdx.x - 64; i < 128; i += 128) cbs[i] = i < p.M ? p.cb[(size_t)b * p.cb_bstride + i] : 0.f;
            asm volatile("bar.sync 1, 128;" ::: "memory");
            // ---------------- epilogue 1: softmax over the tokens, thread = row
            tc::mbar_wait(s_full, tpar);
            tc::tc_fence_after();
            float mx = -INFINITY, sum = 0.f;
            // pass 1: row maximum (logits stay in TMEM; they are read again below -- 128 fp32 per thread would not fit beside the rest)
            for (int c = 0; c < p.N1; c += 32) {
                float v[32];
                tc::tmem_ld32(tmem_s + c + lane_off, v);
                tc::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (c + j < p.M) mx = fmaxf(mx, v[j] + cbs[c + j]);
            }
            for (int c = 0; c < p.N1; c += 32) {
                float v[32];
                tc::tmem_ld32(tmem_s + c + lane_off, v);
                tc::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (c + j < p.M) sum += __expf(v[j] + cbs[c + j] - mx);
            }
            const float inv = 1.f / sum;
            for (int c = 0; c < 64 * p.pboxes; c += 32) {
                float v[32];
                if (c < p.N1) {
                    tc::tmem_ld32(tmem_s + c + lane_off, v);
                    tc::tmem_ld_wait();
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = 0.f;
                }
                float pr[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float l = v[j] + cbs[(c + j) & 127];
                    v[j] = l;
                    pr[j] = (c + j < p.M) ? __expf(l - mx) * inv : 0.f;
                }
                if (live && p.logit != nullptr) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        if (c + j + 4 <= p.ldl) *reinterpret_cast<float4*>(p.logit + grow * p.ldl + c + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                }
                if (live && p.attn != nullptr) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        if (c + j + 4 <= p.ldl) *reinterpret_cast<float4*>(p.attn + grow * p.ldl + c + j) = make_float4(pr[j], pr[j + 1], pr[j + 2], pr[j + 3]);
                }
                // bf16 probabilities -> K-major 128B-swizzled boxes of 64 tokens (rows past len[b] contribute to rows that are never written)
                const uint32_t base = p_addr + (uint32_t)(c >> 6) * 16384u + (uint32_t)row * 128u;
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const int chunk = ((c & 63) >> 3) + g;
                    tc::sts_v4(base + (uint32_t)((chunk ^ (row & 7)) << 4), tc::pack_bf16x2(pr[g * 8], pr[g * 8 + 1]),
                               tc::pack_bf16x2(pr[g * 8 + 2], pr[g * 8 + 3]), tc::pack_bf16x2(pr[g * 8 + 4], pr[g * 8 + 5]),
                               tc::pack_bf16x2(pr[g * 8 + 6], pr[g * 8 + 7]));
                }
            }
            tc::fence_proxy_async_smem();
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) { tc::mbar_arrive(p_full); tc::mbar_arrive(s_free); }
            // ---------------- epilogue 2: out = D + bias -> bf16 rows
            tc::mbar_wait(d_full, tpar);
            tc::tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < XF_F; c += 32) {
                float v[32];
                tc::tmem_ld32(tmem_d + c + lane_off, v);
                tc::tmem_ld_wait();
                if (live) {
                    __nv_bfloat16* o = p.out + grow * (size_t)p.ldo + c;
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + c + j)), b1 = __ldg(reinterpret_cast<const float4*>(p.bias + c + j + 4));
                        *reinterpret_cast<uint4*>(o + j) = make_uint4(tc::pack_bf16x2(v[j] + b0.x, v[j + 1] + b0.y), tc::pack_bf16x2(v[j + 2] + b0.z, v[j + 3] + b0.w),
                                                                     tc::pack_bf16x2(v[j + 4] + b1.x, v[j + 5] + b1.y), tc::pack_bf16x2(v[j + 6] + b1.z, v[j + 7] + b1.w));
                    }
                }
            }
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(d_free);
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace factk

using namespace factk;

/* 1 when factk_a2f_fused serves the shapes: H a multiple of 64, F == 256, M <= 128, slot a multiple of 128. */
extern "C" int factk_a2f_fused_supported(int M, int H, int F, int slot) {
    return (H % 64 == 0) && F == XF_F && M >= 1 && M <= 128 && (slot % 128 == 0);
}

extern "C" int factk_a2f_fused(const void* X, int ldx, const void* Kt, int ldkt, long long kt_bstride, const float* cb, long long cb_bstride,
                               const void* Wy, int ldwy, const void* Vt, int ldvt, long long vt_bstride, const float* bias, void* out, int ldo,
                               float* logit, float* attn, int ldl, int B, int slot, const int32_t* len, int M, int H, int F, void* stream) {
    FACTK_REQUIRE(X && Kt && cb && Wy && Vt && bias && out && B > 0, "factk_a2f_fused: bad args");
    FACTK_REQUIRE(factk_a2f_fused_supported(M, H, F, slot), "factk_a2f_fused: unsupported shape M=%d H=%d F=%d slot=%d", M, H, F, slot);
    FACTK_REQUIRE(ldx % 8 == 0 && ldkt % 8 == 0 && ldwy % 8 == 0 && ldvt % 8 == 0 && ldo % 8 == 0 && kt_bstride % 8 == 0 && vt_bstride % 8 == 0 &&
                      aligned16(X) && aligned16(Kt) && aligned16(Wy) && aligned16(Vt) && aligned16(out) && aligned16(bias),
                  "factk_a2f_fused: operands must be 16-byte aligned");
    FACTK_REQUIRE((!logit && !attn) || (ldl % 4 == 0 && ldl >= M && (!logit || aligned16(logit)) && (!attn || aligned16(attn))),
                  "factk_a2f_fused: logit / attn rows must be 16-byte aligned");
    XfParams p;
    const int pboxes = (M + 63) / 64;
    p.N1 = (M + 15) & ~15;
    if (!tc_get_map(&p.xmap, X, 2, (uint64_t)H, (uint64_t)slot, (uint64_t)B, (uint64_t)ldx, (uint64_t)slot * ldx, 128)) return FACTK_ERR_CUDA;
    if (!tc_get_map(&p.kmap, Kt, 2, (uint64_t)H, (uint64_t)M, (uint64_t)B, (uint64_t)ldkt, (uint64_t)kt_bstride, (uint32_t)p.N1)) return FACTK_ERR_CUDA;
    if (!tc_get_map(&p.wmap, Wy, 2, (uint64_t)H, (uint64_t)F, 1, (uint64_t)ldwy, (uint64_t)F * ldwy, 256)) return FACTK_ERR_CUDA;
    if (!tc_get_map(&p.vmap, Vt, 2, (uint64_t)(64 * pboxes), (uint64_t)F, (uint64_t)B, (uint64_t)ldvt, (uint64_t)vt_bstride, 256)) return FACTK_ERR_CUDA;
    p.cb = cb; p.cb_bstride = cb_bstride; p.bias = bias; p.out = reinterpret_cast<__nv_bfloat16*>(out); p.ldo = ldo;
    p.logit = logit; p.attn = attn; p.ldl = ldl;
    p.B = B; p.slot = slot; p.M = M; p.kchunks = H / 64; p.pboxes = pboxes; p.tiles_m = slot / 128; p.total_tiles = B * p.tiles_m; p.len = len;
    static unsigned long long devs = 0;
    if (first_use_on_device(devs)) cudaFuncSetAttribute(a2f_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, XF_SMEM);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = p.total_tiles < sms ? p.total_tiles : sms;
    a2f_fused_kernel<<<grid, XF_THREADS, XF_SMEM, (cudaStream_t)stream>>>(p);
    return check_launch("factk_a2f_fused");
}
