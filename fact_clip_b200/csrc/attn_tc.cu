// Tokens-attend-frames attention of the SCALayer (models/basic.py:507-514, nn.MultiheadAttention core) on the 5th-generation
// tensor cores: few queries (the action tokens), very long key / value sequences (the frames), 8 heads of 32 channels.
//
//   S_h = Q_h K_h^T          tcgen05.mma  M = 128 token rows (one query block), N = 64 frames, K = 32    -> TMEM
//   softmax                  tcgen05.ld: one thread = one token row of S_h in registers; running max / sum per (token, head) in
//                            registers, exp2 on pre-scaled logits, masking of frames >= len[b] by -inf; the rare rescale of the
//                            output accumulator (the row maximum grew) is a tcgen05.ld / tcgen05.st round trip of 32 columns
//   O_h += P_h V_h           P_h written as bf16 into the 128B-swizzled K-major A-operand layout in shared memory,
//                            V_h consumed MN-major straight from its row-major TMA tile; fp32 accumulators of all 8 heads in TMEM
//
// One CTA = one 512-frame split of one video x one block of 128 tokens, all 8 heads: a K tile and a V tile of [64 frames x
// 256 channels] per pipeline stage arrive by TMA (2 stages), warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-5 and 6-9 = two
// softmax / epilogue warpgroups (TMEM lane quarter = warp % 4): group g owns the heads of parity g, S buffer g in TMEM and P
// buffer g in shared memory, so the softmax of head h+1 runs beside that of head h (two warps per scheduler hide each other's
// latencies -- with one group the kernel was issue-latency bound at 5.5 cycles per instruction, profiles/r2_attn_tc.md) while the
// tensor core alternates S_{h+2} = Q K^T and O_h += P V.
// Every split writes (max, sum, unnormalised O) per (head, token); attn_rows_combine_kernel (attn.cu) merges the splits in a
// fixed order -- same partial format as the mma.sync kernel it replaces.
//
// Rows of K / V in [len[b], slot) must hold finite values (the engine zero-initialises the buffer): their probabilities are
// exactly zero, but 0 * NaN would poison the accumulator.
#include "common.cuh"
#include "tc_common.cuh"

namespace factk {

constexpr int AT_SPLIT = 512;            // frames per CTA (== SPLIT_ROWS of attn.cu)
constexpr int AT_TILE = 64;              // frames per pipeline stage
constexpr int AT_A = 256, AT_DH = 32, AT_NH = 8;
constexpr int AT_QBYTES = 4 * 128 * 128;             // 4 boxes [128 tokens x 64 channels] bf16
constexpr int AT_KVBYTES = 4 * AT_TILE * 128;        // 4 boxes [64 frames x 64 channels] bf16
constexpr int AT_PBYTES = 128 * 128;                 // [128 tokens x 64 frames] bf16
constexpr int AT_SMEM = AT_QBYTES + 2 * 2 * AT_KVBYTES + 2 * AT_PBYTES + 256 + 1024;
constexpr int AT_THREADS = 128 + 256;    // TMA warp, S = Q K^T issuer, O += P V issuer, (idle), two softmax warpgroups (even / odd heads)
static_assert(AT_SMEM <= 232448, "dynamic shared memory limit");

struct AtParams {
    alignas(64) CUtensorMap kmap;
    alignas(64) CUtensorMap vmap;
    const float* Q;
    int ldq, M, slot, nsplit, nqb;
    const int32_t* len;
    float* ws;
    float qscale;       // log2(e) / sqrt(dh)
    long long* dbg;     // optional clock64 timeline of CTA (0,0), softmax group 0 warp 0 (development aid)
};

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float v[32]) {
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
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// MN-major SWIZZLE_128B operand (see wgrad_tc.cu): SBO = 1024 B between 8-frame groups; LBO unused for N <= 64
__device__ __forceinline__ uint64_t at_desc_mn(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(8192 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}

// tcgen05.mma with descriptors given as (low word, constant high word): the issuing thread is ONE lane whose instruction
// stream is latency bound (~5 cycles per dependent instruction), so per MMA it does two 32-bit adds instead of rebuilding two
// 64-bit descriptors (measured: 130 -> ~30 cycles per issued MMA, profiles/r2_attn_tc.md)
constexpr uint32_t AT_DESC_HI = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);       // SBO | version 1 | SWIZZLE_128B
__device__ __forceinline__ uint32_t at_desc_lo_k(uint32_t saddr) { return ((saddr >> 4) & 0x3FFFu) | (1u << 16); }
__device__ __forceinline__ uint32_t at_desc_lo_mn(uint32_t saddr) { return ((saddr >> 4) & 0x3FFFu) | ((uint32_t)(8192 >> 4) << 16); }
template <bool ACC>
__device__ __forceinline__ void at_umma(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\tsetp.ne.b32 p, %5, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
        ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(AT_DESC_HI), "r"(idesc), "r"(ACC ? 1u : 0u)
        : "memory");
}

__global__ void __launch_bounds__(AT_THREADS, 1) attn_rows_tc_kernel(const __grid_constant__ AtParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* qs = smem;
    uint8_t* kvs = qs + AT_QBYTES;                         // stage s: K at kvs + s * 2 * KVBYTES, V right behind it
    uint8_t* ps = kvs + 2 * 2 * AT_KVBYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(ps + 2 * AT_PBYTES);
    uint64_t *kv_full = bars, *kv_empty = bars + 2, *s_full = bars + 4, *s_empty = bars + 6, *p_full = bars + 8, *p_empty = bars + 10,
             *o_full = bars + 12;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long t_entry = clock64();
    const int split = blockIdx.x / p.nqb, qb = blockIdx.x % p.nqb, b = blockIdx.y;
    const int len_b = p.len ? min(p.len[b], p.slot) : p.slot;
    const int r0 = split * AT_SPLIT;
    if (r0 >= len_b) return;
    const int nrows = min(r0 + AT_SPLIT, len_b) - r0;
    const int ntile = (nrows + AT_TILE - 1) / AT_TILE;
    const int nit = ntile * AT_NH;

    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&p.kmap);
        tc::tma_prefetch_desc(&p.vmap);
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(&kv_full[i], 1);
            tc::mbar_init(&kv_empty[i], 2);            // the S issuer (K tile) and the P V issuer (V tile) both release a stage
            tc::mbar_init(&s_full[i], 1);
            tc::mbar_init(&s_empty[i], 4);
            tc::mbar_init(&p_full[i], 4);
            tc::mbar_init(&p_empty[i], 1);
        }
        tc::mbar_init(o_full, 1);
        tc::fence_barrier_init();
    }
    __syncthreads();                                     // barriers initialised; the TMA warp runs ahead from here
    if (warp == 1) {
        tc::tmem_alloc(tmem_slot, 512);
        tc::tmem_relinquish();
    }
    if (warp >= 4) {
        // stage the query block: thread = token row, scaled by log2(e) / sqrt(dh), bf16, 128B-swizzled K-major boxes of 64 channels
        const int row = (warp & 3) * 32 + lane, m = qb * 128 + row;
        const float* q = p.Q + ((size_t)b * p.M + m) * (size_t)p.ldq;
        const uint32_t qbase = tc::smem_u32(qs);
        const int c8_0 = ((warp - 4) >> 2) * (AT_A / 16);           // each group stages half of the channels
        const bool vec = (p.ldq & 3) == 0 && (reinterpret_cast<uintptr_t>(p.Q) & 15u) == 0;
#pragma unroll
        for (int half = 0; half < 2; ++half) {                       // 8 chunks of 8 channels per batch: 16 loads in flight
            float4 v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int c = (c8_0 + half * 8) * 8 + i * 4;
                if (m >= p.M) v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                else if (vec) v[i] = __ldg(reinterpret_cast<const float4*>(q + c));
                else v[i] = make_float4(q[c], q[c + 1], q[c + 2], q[c + 3]);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int c8 = c8_0 + half * 8 + i;
                const float4 a = v[2 * i], c = v[2 * i + 1];
                const uint32_t addr = qbase + (uint32_t)(c8 >> 3) * 16384u + (uint32_t)row * 128u + (uint32_t)(((c8 & 7) ^ (row & 7)) << 4);
                tc::sts_v4(addr, tc::pack_bf16x2(a.x * p.qscale, a.y * p.qscale), tc::pack_bf16x2(a.z * p.qscale, a.w * p.qscale),
                           tc::pack_bf16x2(c.x * p.qscale, c.y * p.qscale), tc::pack_bf16x2(c.z * p.qscale, c.w * p.qscale));
            }
        }
        tc::fence_proxy_async_smem();
    }
    uint32_t tmem_base = 0;
    if (warp >= 1) {                                     // Q staged + TMEM allocated: everyone but the TMA warp meets here
        tc::tc_fence_before();
        asm volatile("bar.sync 1, %0;" ::"n"(AT_THREADS - 32) : "memory");
        tc::tc_fence_after();
        tmem_base = *tmem_slot;
    }
    const uint32_t tmem_o = tmem_base + 128;               // S buffers: columns [0, 128); O_h: 128 + 32 h

    if (warp == 0) {
        if (lane == 0) {
            for (int t = 0; t < ntile; ++t) {
                const int st = t & 1;
                tc::mbar_wait(&kv_empty[st], ((t >> 1) & 1) ^ 1);
                tc::mbar_arrive_expect_tx(&kv_full[st], 2 * AT_KVBYTES);
                uint8_t* kd = kvs + st * 2 * AT_KVBYTES;
                const int row = r0 + t * AT_TILE;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    tc::tma_load_3d(kd + j * (AT_TILE * 128), &p.kmap, &kv_full[st], j * 64, row, b);
                    tc::tma_load_3d(kd + AT_KVBYTES + j * (AT_TILE * 128), &p.vmap, &kv_full[st], j * 64, row, b);
                }
            }
        }
    } else if (warp == 1 || warp == 2) {
        // Two issuing threads: warp 1 keeps the S buffers full (it only ever waits for a free S buffer and for K tiles), warp 2
        // issues O_h += P V as probabilities arrive.  With ONE in-order issuer the wait for P of one softmax group delayed
        // the next S of the other: every iteration then paid two ~500-cycle mbarrier hand-offs in series (profiles/r2_attn_tc.md).
        if (lane == 0) {
            constexpr uint32_t idesc_s = tc::instr_desc(128, AT_TILE, false);
            constexpr uint32_t idesc_o = tc::instr_desc(128, AT_DH, false) | (1u << 16);        // B (= V) MN-major
            const uint32_t q_lo = at_desc_lo_k(tc::smem_u32(qs)), k_lo = at_desc_lo_k(tc::smem_u32(kvs)), p_lo = at_desc_lo_k(tc::smem_u32(ps));
            const uint32_t v_lo = at_desc_lo_mn(tc::smem_u32(kvs) + AT_KVBYTES);
            // descriptor low-word offsets (bytes >> 4): head h -> box h/2 (+ 64 B for odd heads); stage -> 2 * KVBYTES
            auto head_off = [](int h, int box_bytes) { return (uint32_t)(((h >> 1) * box_bytes + (h & 1) * 64) >> 4); };
            if (warp == 1) {
                for (int it = 0; it < nit; ++it) {         // S[it & 1] = Q_h K_h^T
                    const int t = it >> 3, h = it & 7, sb = it & 1;
                    if (h == 0) {
                        tc::mbar_wait(&kv_full[t & 1], (t >> 1) & 1);
                        tc::tc_fence_after();
                    }
                    tc::mbar_wait(&s_empty[sb], ((it >> 1) & 1) ^ 1);
                    tc::tc_fence_after();
                    const uint32_t qa = q_lo + head_off(h, 16384);
                    const uint32_t ka = k_lo + (uint32_t)(((t & 1) * 2 * AT_KVBYTES) >> 4) + head_off(h, AT_TILE * 128);
                    at_umma<false>(tmem_base + sb * AT_TILE, qa, ka, idesc_s);
                    at_umma<true>(tmem_base + sb * AT_TILE, qa + 2, ka + 2, idesc_s);
                    tc::umma_commit(&s_full[sb]);
                    if (h == AT_NH - 1) tc::umma_commit(&kv_empty[t & 1]);
                }
            } else {
                for (int it = 0; it < nit; ++it) {         // O_h += P V_h
                    const int t = it >> 3, h = it & 7, pb = it & 1;
                    if (h == 0) {
                        tc::mbar_wait(&kv_full[t & 1], (t >> 1) & 1);
                        tc::tc_fence_after();
                    }
                    tc::mbar_wait(&p_full[pb], (it >> 1) & 1);
                    tc::tc_fence_after();
                    const uint32_t pa = p_lo + (uint32_t)((pb * AT_PBYTES) >> 4);
                    const uint32_t va = v_lo + (uint32_t)(((t & 1) * 2 * AT_KVBYTES) >> 4) + head_off(h, AT_TILE * 128);
                    const uint32_t d = tmem_o + h * AT_DH;
                    if (t > 0) at_umma<true>(d, pa, va, idesc_o);
                    else at_umma<false>(d, pa, va, idesc_o);
#pragma unroll
                    for (int k = 1; k < AT_TILE / 16; ++k) at_umma<true>(d, pa + k * 2, va + k * 128, idesc_o);     // + 32 B / + 2048 B
                    tc::umma_commit(&p_empty[pb]);
                    if (h == AT_NH - 1) tc::umma_commit(&kv_empty[t & 1]);
                }
                tc::umma_commit(o_full);
            }
        }
    } else if (warp == 3) {
        // idle (keeps the softmax warps at warp % 4 == TMEM lane quarter)
    } else {
        const int q = warp & 3, grp = (warp - 4) >> 2;             // group = head parity = S / P buffer index
        const int row = q * 32 + lane, m = qb * 128 + row;
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        float mx[AT_NH / 2], ls[AT_NH / 2];
#pragma unroll
        for (int i = 0; i < AT_NH / 2; ++i) { mx[i] = -INFINITY; ls[i] = 0.f; }
        const uint32_t s_tmem = tmem_base + grp * AT_TILE + lane_off;
        const uint32_t prow = tc::smem_u32(ps) + grp * AT_PBYTES + (uint32_t)row * 128u;
        int n = 0;                                                 // uses of this group's buffers so far
        const bool dbg_on = p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && warp == 4 && lane == 0;
        if (dbg_on) { p.dbg[0] = clock64(); p.dbg[500] = t_entry; }
        for (int t = 0; t < ntile; ++t) {
            const int valid = min(AT_TILE, nrows - t * AT_TILE);   // frames of this tile below len[b]
#pragma unroll
            for (int i = 0; i < AT_NH / 2; ++i, ++n) {
                const int h = 2 * i + grp;
                if (dbg_on) p.dbg[1 + n * 4 + 0] = clock64();
                tc::mbar_wait(&s_full[grp], n & 1);
                tc::tc_fence_after();
                if (dbg_on) p.dbg[1 + n * 4 + 1] = clock64();
                float s[AT_TILE];
                tc::tmem_ld32(s_tmem, s);
                tc::tmem_ld32(s_tmem + 32, s + 32);
                tc::tmem_ld_wait();
                tc::tc_fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&s_empty[grp]);
                if (valid < AT_TILE) {                             // warp-uniform: only the last tile of a video is ragged
#pragma unroll
                    for (int j = 0; j < AT_TILE; ++j)
                        if (j >= valid) s[j] = -INFINITY;
                }
                float r4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};        // 4 independent chains
#pragma unroll
                for (int j = 0; j < AT_TILE; j += 4) {
                    r4[0] = fmaxf(r4[0], s[j]); r4[1] = fmaxf(r4[1], s[j + 1]); r4[2] = fmaxf(r4[2], s[j + 2]); r4[3] = fmaxf(r4[3], s[j + 3]);
                }
                const float mnew = fmaxf(mx[i], fmaxf(fmaxf(r4[0], r4[1]), fmaxf(r4[2], r4[3])));
                float a4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int j = 0; j < AT_TILE; j += 4) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        float e;
                        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(s[j + u] - mnew));
                        s[j + u] = e;
                        a4[u] += e;
                    }
                }
                const float rsum = (a4[0] + a4[1]) + (a4[2] + a4[3]);
                // the previous P V of this head (and every earlier MMA) has completed once this group's P buffer is free again
                if (dbg_on) p.dbg[1 + n * 4 + 2] = clock64();
                tc::mbar_wait(&p_empty[grp], (n & 1) ^ 1);
                tc::tc_fence_after();
                if (dbg_on) p.dbg[1 + n * 4 + 3] = clock64();
                float sc;
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(sc) : "f"(mx[i] - mnew));
                if (t > 0 && __any_sync(0xffffffffu, mnew > mx[i])) {      // rescale O_h: a row maximum moved
                    float o[AT_DH];
                    tc::tmem_ld32(tmem_o + h * AT_DH + lane_off, o);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < AT_DH; ++j) o[j] *= sc;
                    tmem_st32(tmem_o + h * AT_DH + lane_off, o);
                    tmem_st_wait();
                }
                ls[i] = ls[i] * sc + rsum;
                mx[i] = mnew;
                // P row -> bf16, 8 chunks of 16 bytes, K-major 128B swizzle (chunk ^ (row % 8))
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    tc::sts_v4(prow + (uint32_t)((c ^ (row & 7)) << 4), tc::pack_bf16x2(s[c * 8], s[c * 8 + 1]),
                               tc::pack_bf16x2(s[c * 8 + 2], s[c * 8 + 3]), tc::pack_bf16x2(s[c * 8 + 4], s[c * 8 + 5]),
                               tc::pack_bf16x2(s[c * 8 + 6], s[c * 8 + 7]));
                tc::fence_proxy_async_smem();
                tc::tc_fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&p_full[grp]);
            }
        }
        // epilogue: (max, sum, O) of every (head, token) of this split
        if (dbg_on) p.dbg[501] = clock64();
        tc::mbar_wait(o_full, 0);
        tc::tc_fence_after();
        if (dbg_on) p.dbg[502] = clock64();
        // every MMA and TMA load has completed: the K / V stages serve as staging so that each warp's 32 rows x 34 floats, which
        // are contiguous in the partial workspace, leave as coalesced 128-byte stores
        float* stg = reinterpret_cast<float*>(kvs) + (warp - 4) * (32 * (AT_DH + 2));
        const int m0 = qb * 128 + q * 32;
        const int nval = max(0, min(32, p.M - m0));
#pragma unroll
        for (int i = 0; i < AT_NH / 2; ++i) {
            const int h = 2 * i + grp;
            float o[AT_DH];
            tc::tmem_ld32(tmem_o + h * AT_DH + lane_off, o);
            tc::tmem_ld_wait();
            float* r = stg + lane * (AT_DH + 2);
            r[0] = mx[i] * 0.6931471805599453f;                // back to natural-log units for the combine kernel
            r[1] = ls[i];
#pragma unroll
            for (int j = 0; j < AT_DH; j += 2) *reinterpret_cast<float2*>(r + 2 + j) = make_float2(o[j], o[j + 1]);
            __syncwarp();
            float* out = p.ws + ((((size_t)b * AT_NH + h) * p.nsplit + split) * p.M + m0) * (AT_DH + 2);
            for (int e = lane; e < nval * (AT_DH + 2); e += 32) out[e] = stg[e];
            __syncwarp();
        }
    }
    if (p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 128) p.dbg[503] = clock64();
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem_base, 512);
    }
}

// 1 when the tcgen05 kernel serves these operands (bf16 rows, 8 heads of 32 channels, 16-byte aligned, slot % 64 == 0).
bool attn_rows_tc_ok(const void* Kx, const void* Vx, int kv_dtype, int ldkv, int slot, int nhead, int dh) {
    return kv_dtype == FACTK_BF16 && nhead == AT_NH && dh == AT_DH && (ldkv % 8) == 0 && (slot % AT_TILE) == 0 && aligned16(Kx) && aligned16(Vx);
}

long long* g_attn_dbg = nullptr;

int attn_rows_tc_launch(const float* Q, int ldq, const void* Kx, const void* Vx, int ldkv, int B, int slot, const int32_t* len, int M,
                        int nsplit, float* ws, cudaStream_t st) {
    AtParams p;
    p.dbg = g_attn_dbg;
    if (!tc_get_map(&p.kmap, Kx, 2, AT_A, (uint64_t)slot, (uint64_t)B, (uint64_t)ldkv, (uint64_t)slot * ldkv, AT_TILE)) return FACTK_ERR_CUDA;
    if (!tc_get_map(&p.vmap, Vx, 2, AT_A, (uint64_t)slot, (uint64_t)B, (uint64_t)ldkv, (uint64_t)slot * ldkv, AT_TILE)) return FACTK_ERR_CUDA;
    p.Q = Q; p.ldq = ldq; p.M = M; p.slot = slot; p.nsplit = nsplit; p.nqb = (M + 127) / 128; p.len = len; p.ws = ws;
    p.qscale = 1.4426950408889634f / sqrtf((float)AT_DH);
    static unsigned long long devs = 0;
    if (first_use_on_device(devs)) cudaFuncSetAttribute(attn_rows_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM);
    attn_rows_tc_kernel<<<dim3(nsplit * p.nqb, B), AT_THREADS, AT_SMEM, st>>>(p);
    return FACTK_OK;
}

}  // namespace factk

/* development aid: clock64 timeline buffer (>= 1 + 4 * 32 int64) for the next attention launches; NULL switches it off */
extern "C" int factk_attn_tc_debug(long long* dbg) {
    factk::g_attn_dbg = dbg;
    return FACTK_OK;
}
