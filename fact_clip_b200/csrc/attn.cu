// Attention cores that are not plain GEMMs.
//  * mha_tokens: token self-attention (M <= ~300 tokens), one CTA per (head, video).
//  * attn_rows: tokens attend the T frames of their video (SCALayer cross attention), split over T
//    with a deterministic log-sum-exp combine.
//  * col_softmax_apply: softmax over the frame/segment axis of a logit matrix and the weighted row sum
//    (the f2a direction of X2Y_map), again split over rows and combined in fixed order.
// All softmax state lives in registers; fp32 throughout.
#include <cstdlib>

#include "common.cuh"

namespace factk {

constexpr int SPLIT_ROWS = 512;   // rows of a video handled by one CTA in the split kernels

// ------------------------------------------------------------------------------------------------
// nn.MultiheadAttention core for self attention among tokens (models/basic.py:437,500).
template <int DH>
__global__ void __launch_bounds__(128) mha_tokens_kernel(const float* __restrict__ Q, const float* __restrict__ K,
                                                         const float* __restrict__ V, int ld, float* __restrict__ O,
                                                         int ldo, int M) {
    extern __shared__ float sm[];
    float* Ks = sm;
    float* Vs = sm + (size_t)M * DH;
    const int h = blockIdx.x, b = blockIdx.y;
    const size_t base = (size_t)b * M;
    for (int i = threadIdx.x; i < M * DH; i += blockDim.x) {
        const int m = i / DH, d = i % DH;
        Ks[i] = K[(base + m) * ld + h * DH + d];
        Vs[i] = V[(base + m) * ld + h * DH + d];
    }
    __syncthreads();
    const float scale = rsqrtf((float)DH);
    for (int m = threadIdx.x; m < M; m += blockDim.x) {
        float q[DH], acc[DH];
#pragma unroll
        for (int d = 0; d < DH; ++d) { q[d] = Q[(base + m) * ld + h * DH + d] * scale; acc[d] = 0.f; }
        float mx = -INFINITY, l = 0.f;
        for (int j = 0; j < M; ++j) {
            float s = 0.f;
#pragma unroll
            for (int d = 0; d < DH; ++d) s = fmaf(q[d], Ks[j * DH + d], s);
            const float nm = fmaxf(mx, s);
            const float corr = __expf(mx - nm), p = __expf(s - nm);
            l = l * corr + p;
#pragma unroll
            for (int d = 0; d < DH; ++d) acc[d] = fmaf(p, Vs[j * DH + d], acc[d] * corr);
            mx = nm;
        }
        const float inv = 1.f / l;
#pragma unroll
        for (int d = 0; d < DH; ++d) O[(base + m) * ldo + h * DH + d] = acc[d] * inv;
    }
}

// ------------------------------------------------------------------------------------------------
// The same on the tensor cores (tf32 mma.sync m16n8k8, fp32 accumulate), bf16 compute mode: one CTA per (head, video),
// one warp per 16-query tile, keys in chunks of 80 with a running max / sum (flash style), K and V of the head staged in
// shared memory as tf32.  The thread-per-query kernel above is a 75-iteration dependent chain on 75 of 128 threads
// (43 us per launch at 64 videos x 8 heads); the tensor-core form is a few hundred MMAs per CTA.
//   QK^T: A = Q tile (row-major, straight from global, pre-scaled), B[k = dh][n = key] = Ks[key][dh]
//   P V : A = P in the accumulator layout of QK^T with the key index inside each 8-block relabelled (k = t <-> key 2t,
//         k = t+4 <-> key 2t+1: a sum over keys does not care), B[k][n = dh] = Vs[key(k)][dh]  -> no shuffles
// Row strides of 36 floats make both B-fragment patterns bank-conflict free.
__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

constexpr int MT_WARPS = 5, MT_KCH = 80, MT_LD = 36 + 0;     // key chunk, smem row stride (floats) for DH = 32

template <int DH>
__global__ void __launch_bounds__(MT_WARPS * 32) mha_tokens_mma_kernel(const float* __restrict__ Q, const float* __restrict__ K,
                                                                       const float* __restrict__ V, int ld, float* __restrict__ O,
                                                                       int ldo, int M) {
    constexpr int LDS_ = DH + 4;          // 4 * key + dh and 8 * t + g bank patterns are conflict free for DH % 32 == 0 or 16
    extern __shared__ uint32_t smu[];
    const int Mp = (M + 7) & ~7;
    uint32_t* Ks = smu;                   // [Mp][LDS_] tf32
    uint32_t* Vs = smu + (size_t)Mp * LDS_;
    const int h = blockIdx.x, b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const size_t base = (size_t)b * M;
    for (int i = threadIdx.x; i < Mp * DH; i += blockDim.x) {
        const int m = i / DH, d = i % DH;
        const bool ok = m < M;
        Ks[m * LDS_ + d] = to_tf32(ok ? K[(base + m) * ld + h * DH + d] : 0.f);
        Vs[m * LDS_ + d] = to_tf32(ok ? V[(base + m) * ld + h * DH + d] : 0.f);
    }
    __syncthreads();
    const float scale = rsqrtf((float)DH);
    for (int m0 = warp * 16; m0 < M; m0 += MT_WARPS * 16) {
        // Q tile fragments: a0 = Q[g][t], a1 = Q[g+8][t], a2 = Q[g][t+4], a3 = Q[g+8][t+4] per k-step of 8
        uint32_t qa[DH / 8][4];
        const int r0 = m0 + g, r1 = m0 + g + 8;
#pragma unroll
        for (int kk = 0; kk < DH / 8; ++kk) {
            const int d = h * DH + kk * 8 + t;
            qa[kk][0] = to_tf32(r0 < M ? Q[(base + r0) * ld + d] * scale : 0.f);
            qa[kk][1] = to_tf32(r1 < M ? Q[(base + r1) * ld + d] * scale : 0.f);
            qa[kk][2] = to_tf32(r0 < M ? Q[(base + r0) * ld + d + 4] * scale : 0.f);
            qa[kk][3] = to_tf32(r1 < M ? Q[(base + r1) * ld + d + 4] * scale : 0.f);
        }
        float acc[DH / 8][4];
#pragma unroll
        for (int n = 0; n < DH / 8; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;
        float mx0 = -INFINITY, mx1 = -INFINITY, l0 = 0.f, l1 = 0.f;
        for (int k0 = 0; k0 < Mp; k0 += MT_KCH) {
            const int nkt = min(MT_KCH, Mp - k0) / 8;           // 8-key tiles in this chunk
            float sc[MT_KCH / 8][4];
#pragma unroll
            for (int j = 0; j < MT_KCH / 8; ++j) {
                sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = 0.f;
                if (j < nkt) {
                    const uint32_t* kr = Ks + (size_t)(k0 + j * 8 + g) * LDS_;
#pragma unroll
                    for (int kk = 0; kk < DH / 8; ++kk) mma_tf32(sc[j], qa[kk], kr[kk * 8 + t], kr[kk * 8 + t + 4]);
                }
            }
            // running softmax over the chunk: thread holds rows g (c0,c1) and g+8 (c2,c3), keys 2t, 2t+1 of each 8-tile
            float cm0 = -INFINITY, cm1 = -INFINITY;
#pragma unroll
            for (int j = 0; j < MT_KCH / 8; ++j) {
                if (j < nkt) {
                    const int key = k0 + j * 8 + 2 * t;
                    if (key >= M) sc[j][0] = sc[j][2] = -INFINITY;
                    if (key + 1 >= M) sc[j][1] = sc[j][3] = -INFINITY;
                    cm0 = fmaxf(cm0, fmaxf(sc[j][0], sc[j][1]));
                    cm1 = fmaxf(cm1, fmaxf(sc[j][2], sc[j][3]));
                }
            }
            cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 1));
            cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 2));
            cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 1));
            cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 2));
            const float n0 = fmaxf(mx0, cm0), n1 = fmaxf(mx1, cm1);
            const float f0 = __expf(mx0 - n0), f1 = __expf(mx1 - n1);
            mx0 = n0; mx1 = n1;
            l0 *= f0; l1 *= f1;
#pragma unroll
            for (int n = 0; n < DH / 8; ++n) { acc[n][0] *= f0; acc[n][1] *= f0; acc[n][2] *= f1; acc[n][3] *= f1; }
#pragma unroll
            for (int j = 0; j < MT_KCH / 8; ++j) {
                if (j < nkt) {
                    const float p0 = __expf(sc[j][0] - n0), p1 = __expf(sc[j][1] - n0);
                    const float p2 = __expf(sc[j][2] - n1), p3 = __expf(sc[j][3] - n1);
                    l0 += p0 + p1;
                    l1 += p2 + p3;
                    // A fragment of P with the relabelled key order: k = t <-> key 2t, k = t+4 <-> key 2t+1
                    const uint32_t pa[4] = {to_tf32(p0), to_tf32(p2), to_tf32(p1), to_tf32(p3)};
                    const uint32_t* v0 = Vs + (size_t)(k0 + j * 8 + 2 * t) * LDS_;
#pragma unroll
                    for (int n = 0; n < DH / 8; ++n) mma_tf32(acc[n], pa, v0[n * 8 + g], v0[LDS_ + n * 8 + g]);
                }
            }
        }
        l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
        l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
        const float i0 = 1.f / l0, i1 = 1.f / l1;
#pragma unroll
        for (int n = 0; n < DH / 8; ++n) {
            const int d = h * DH + n * 8 + 2 * t;
            if (r0 < M) *reinterpret_cast<float2*>(O + (base + r0) * ldo + d) = make_float2(acc[n][0] * i0, acc[n][1] * i0);
            if (r1 < M) *reinterpret_cast<float2*>(O + (base + r1) * ldo + d) = make_float2(acc[n][2] * i1, acc[n][3] * i1);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// SCALayer cross attention core (models/basic.py:507-514): partial over one split of rows.
// partial layout: [B][nhead][nsplit][M][DH+2] = (running max, running sum, acc[DH]).
template <int DH>
__global__ void __launch_bounds__(320) attn_rows_partial_kernel(const float* __restrict__ Q, int ldq,
                                                                const void* __restrict__ Kx, const void* __restrict__ Vx,
                                                                int kv_dtype, int ldkv, float* __restrict__ part,
                                                                int slot, const int32_t* __restrict__ len, int M,
                                                                int nhead, int nsplit) {
    constexpr int TR = 32;   // rows staged per iteration
    __shared__ __align__(16) float Ks[TR][DH];
    __shared__ __align__(16) float Vs[TR][DH];
    const int sp = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int len_b = len ? min(len[b], slot) : slot;
    const int r0 = sp * SPLIT_ROWS;
    if (r0 >= len_b) return;
    const int r1 = min(r0 + SPLIT_ROWS, len_b);
    const int m = threadIdx.x;
    const bool active = m < M;
    const float scale = rsqrtf((float)DH);
    float q[DH], acc[DH];
#pragma unroll
    for (int d = 0; d < DH; ++d) {
        q[d] = active ? Q[((size_t)b * M + m) * ldq + h * DH + d] * scale : 0.f;
        acc[d] = 0.f;
    }
    float mx = -INFINITY, l = 0.f;
    for (int t0 = r0; t0 < r1; t0 += TR) {
        const int nr = min(TR, r1 - t0);
        __syncthreads();
        for (int i = threadIdx.x; i < TR * DH; i += blockDim.x) {
            const int r = i / DH, d = i % DH;
            float kv = 0.f, vv = 0.f;
            if (r < nr) {
                const size_t off = ((size_t)b * slot + t0 + r) * ldkv + h * DH + d;
                kv = ld_elem(Kx, kv_dtype, off);
                vv = ld_elem(Vx, kv_dtype, off);
            }
            Ks[r][d] = kv;
            Vs[r][d] = vv;
        }
        __syncthreads();
        if (!active) continue;
#pragma unroll 1
        for (int r8 = 0; r8 < nr; r8 += 8) {
            float s[8];
            float bm = mx;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float a = 0.f;
#pragma unroll
                for (int d = 0; d < DH; d += 4) {
                    const float4 k4 = *reinterpret_cast<const float4*>(&Ks[(r8 + j) & (TR - 1)][d]);
                    a = fmaf(q[d], k4.x, a); a = fmaf(q[d + 1], k4.y, a);
                    a = fmaf(q[d + 2], k4.z, a); a = fmaf(q[d + 3], k4.w, a);
                }
                s[j] = (r8 + j < nr) ? a : -INFINITY;
                bm = fmaxf(bm, s[j]);
            }
            const float corr = __expf(mx - bm);
            l *= corr;
#pragma unroll
            for (int d = 0; d < DH; ++d) acc[d] *= corr;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float p = __expf(s[j] - bm);   // exp(-inf) = 0 for padded rows
                l += p;
#pragma unroll
                for (int d = 0; d < DH; d += 4) {
                    const float4 v4 = *reinterpret_cast<const float4*>(&Vs[(r8 + j) & (TR - 1)][d]);
                    acc[d] = fmaf(p, v4.x, acc[d]); acc[d + 1] = fmaf(p, v4.y, acc[d + 1]);
                    acc[d + 2] = fmaf(p, v4.z, acc[d + 2]); acc[d + 3] = fmaf(p, v4.w, acc[d + 3]);
                }
            }
            mx = bm;
        }
    }
    if (active) {
        float* o = part + ((((size_t)b * nhead + h) * nsplit + sp) * M + m) * (DH + 2);
        o[0] = mx;
        o[1] = l;
#pragma unroll
        for (int d = 0; d < DH; ++d) o[2 + d] = acc[d];
    }
}

// Same computation on the warp-level tensor-core path for bf16 keys / values (DH a multiple of 16): each warp owns
// 16 query tokens; K/V tiles of 64 frames are staged in shared memory (register-prefetched one tile ahead),
// S = Q K^T and O += P V run as mma.sync m16n8k16 (bf16 in, fp32 accumulate), the online softmax stays in the
// accumulator registers with quad shuffles.  Writes the same partial layout as attn_rows_partial_kernel.
__device__ __forceinline__ void mma_bf16_16816(float c[4], const uint32_t a[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int DH, bool QSPLIT>
__global__ void __launch_bounds__(256) attn_rows_mma_kernel(const float* __restrict__ Q, int ldq,
                                                            const __nv_bfloat16* __restrict__ Kx,
                                                            const __nv_bfloat16* __restrict__ Vx, int ldkv,
                                                            float* __restrict__ part, int slot,
                                                            const int32_t* __restrict__ len, int M, int nhead, int nsplit) {
    constexpr int TR = 64;              // frames per tile
    constexpr int LDT = DH + 8;         // padded smem row (bf16 elements): conflict-free fragment loads, 16B-aligned rows
    constexpr int KS = DH / 16;         // k-steps of the QK^T product
    constexpr int NT = DH / 8;          // n-tiles of the output
    constexpr int CH = DH / 8;          // 16-byte chunks per row
    __shared__ __align__(16) __nv_bfloat16 Ksm[2][TR * LDT];      // double-buffered by cp.async: the copy of tile i+1 runs
    __shared__ __align__(16) __nv_bfloat16 Vsm[2][TR * LDT];      // under the MMAs of tile i without holding registers
    const int sp = blockIdx.x % nsplit, qb = blockIdx.x / nsplit, h = blockIdx.y, b = blockIdx.z;
    const int len_b = len ? min(len[b], slot) : slot;
    const int r0 = sp * SPLIT_ROWS;
    if (r0 >= len_b) return;
    const int r1 = min(r0 + SPLIT_ROWS, len_b);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tig = lane & 3;
    const int m_lo = qb * 128 + warp * 16 + g, m_hi = m_lo + 8;   // up to 8 warps x 16 queries per CTA
    // logits in log2 units (log2(e) folded into the query scale): one ex2.approx per probability, no multiply
    const float scale = rsqrtf((float)DH) * 1.4426950408889634f;

    // Q fragments (A operand, row-major 16 x DH), scaled, bf16
    // (queries are fp32: split q = hi + lo into two bf16 fragments so the logits only carry the keys' rounding)
    uint32_t qa[KS][4], qb_lo[KS][4];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
        const int d = ks * 16 + tig * 2;
        const float* ql = Q + ((size_t)b * M + m_lo) * ldq + h * DH + d;
        const float* qh = Q + ((size_t)b * M + m_hi) * ldq + h * DH + d;
        const bool vl = m_lo < M, vh = m_hi < M;
        float v[8] = {vl ? ql[0] * scale : 0.f, vl ? ql[1] * scale : 0.f, vh ? qh[0] * scale : 0.f, vh ? qh[1] * scale : 0.f,
                      vl ? ql[8] * scale : 0.f, vl ? ql[9] * scale : 0.f, vh ? qh[8] * scale : 0.f, vh ? qh[9] * scale : 0.f};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            qa[ks][i] = pack_bf16(v[2 * i], v[2 * i + 1]);
            const float2 hi = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&qa[ks][i]));
            qb_lo[ks][i] = pack_bf16(v[2 * i] - hi.x, v[2 * i + 1] - hi.y);
        }
    }
    float o[NT][4];
#pragma unroll
    for (int i = 0; i < NT; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
    float mx_lo = -INFINITY, mx_hi = -INFINITY, l_lo = 0.f, l_hi = 0.f;

    // cooperative tile load: chunk id -> (row, 16B chunk), 16-byte cp.async (rows beyond the split are zero-filled)
    constexpr int NCHUNK = TR * CH;
    const int nthr = blockDim.x;
    constexpr int MAXPF = 4;            // blockDim.x >= 128 and NCHUNK <= 512, so 4 chunks per thread cover a tile
    auto issue_tile = [&](int t0, int buf) {
#pragma unroll
        for (int i = 0; i < MAXPF; ++i) {
            const int c = threadIdx.x + i * nthr;
            if (c < NCHUNK) {
                const int r = c / CH, ch = c % CH;
                const bool ok = t0 + r < r1;
                const size_t off = ((size_t)b * slot + (ok ? t0 + r : r0)) * ldkv + h * DH + ch * 8;
                const uint32_t dk = (uint32_t)__cvta_generic_to_shared(&Ksm[buf][r * LDT + ch * 8]);
                const uint32_t dv = (uint32_t)__cvta_generic_to_shared(&Vsm[buf][r * LDT + ch * 8]);
                const int nbytes = ok ? 16 : 0;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dk), "l"(Kx + off), "r"(nbytes) : "memory");
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dv), "l"(Vx + off), "r"(nbytes) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    issue_tile(r0, 0);
    int buf = 0;
    for (int t0 = r0; t0 < r1; t0 += TR, buf ^= 1) {
        if (t0 + TR < r1) {
            issue_tile(t0 + TR, buf ^ 1);        // its previous contents were consumed before the barrier that ended tile i-1
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        const __nv_bfloat16* Ks = Ksm[buf];
        const __nv_bfloat16* Vs = Vsm[buf];
        // S = Q K^T  (16 queries x 64 frames)
        float sacc[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            sacc[j][0] = sacc[j][1] = sacc[j][2] = sacc[j][3] = 0.f;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                const __nv_bfloat16* kr = &Ks[(j * 8 + g) * LDT + ks * 16 + tig * 2];
                const uint32_t b0 = *reinterpret_cast<const uint32_t*>(kr);
                const uint32_t b1 = *reinterpret_cast<const uint32_t*>(kr + 8);
                mma_bf16_16816(sacc[j], qa[ks], b0, b1);
                if constexpr (QSPLIT) mma_bf16_16816(sacc[j], qb_lo[ks], b0, b1);
            }
        }
        // mask frames beyond the split (only its last tile can be ragged), online softmax (rows g and g+8 of this warp's tile)
        if (t0 + TR > r1) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int t = t0 + j * 8 + tig * 2;
                if (t >= r1) { sacc[j][0] = -INFINITY; sacc[j][2] = -INFINITY; }
                if (t + 1 >= r1) { sacc[j][1] = -INFINITY; sacc[j][3] = -INFINITY; }
            }
        }
        float nm_lo = mx_lo, nm_hi = mx_hi;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            nm_lo = fmaxf(nm_lo, fmaxf(sacc[j][0], sacc[j][1]));
            nm_hi = fmaxf(nm_hi, fmaxf(sacc[j][2], sacc[j][3]));
        }
        nm_lo = fmaxf(nm_lo, __shfl_xor_sync(0xffffffffu, nm_lo, 1));
        nm_lo = fmaxf(nm_lo, __shfl_xor_sync(0xffffffffu, nm_lo, 2));
        nm_hi = fmaxf(nm_hi, __shfl_xor_sync(0xffffffffu, nm_hi, 1));
        nm_hi = fmaxf(nm_hi, __shfl_xor_sync(0xffffffffu, nm_hi, 2));
        if (__any_sync(0xffffffffu, nm_lo != mx_lo || nm_hi != mx_hi)) {           // a running maximum moved: rescale (rare after the
            const float c_lo = ex2_approx(mx_lo - nm_lo), c_hi = ex2_approx(mx_hi - nm_hi);   // first tiles); 2^(-inf) = 0 on the first tile
            l_lo *= c_lo; l_hi *= c_hi;
#pragma unroll
            for (int i = 0; i < NT; ++i) { o[i][0] *= c_lo; o[i][1] *= c_lo; o[i][2] *= c_hi; o[i][3] *= c_hi; }
            mx_lo = nm_lo; mx_hi = nm_hi;
        }
        // P = exp(S - max) packed straight into A fragments; O += P V
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            uint32_t pa[4];
            {
                const float p00 = ex2_approx(sacc[2 * kk][0] - nm_lo), p01 = ex2_approx(sacc[2 * kk][1] - nm_lo);
                const float p10 = ex2_approx(sacc[2 * kk][2] - nm_hi), p11 = ex2_approx(sacc[2 * kk][3] - nm_hi);
                const float q00 = ex2_approx(sacc[2 * kk + 1][0] - nm_lo), q01 = ex2_approx(sacc[2 * kk + 1][1] - nm_lo);
                const float q10 = ex2_approx(sacc[2 * kk + 1][2] - nm_hi), q11 = ex2_approx(sacc[2 * kk + 1][3] - nm_hi);
                l_lo += p00 + p01 + q00 + q01;
                l_hi += p10 + p11 + q10 + q11;
                pa[0] = pack_bf16(p00, p01); pa[1] = pack_bf16(p10, p11);
                pa[2] = pack_bf16(q00, q01); pa[3] = pack_bf16(q10, q11);
            }
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                // B fragment (k = frame, n = channel) from row-major V via ldmatrix.trans
                const int r = kk * 16 + (lane & 15);
                const uint32_t addr = (uint32_t)__cvta_generic_to_shared(&Vs[r * LDT + nt * 8]);
                uint32_t b0, b1;
                asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];" : "=r"(b0), "=r"(b1) : "r"(addr));
                mma_bf16_16816(o[nt], pa, b0, b1);
            }
        }
        __syncthreads();                 // tile fully consumed: the next iteration refills the other buffer's partner
    }
    // partial results: (running max, running sum, acc[DH]) per query
    l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1); l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
    l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1); l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
    float* base = part + (((size_t)b * nhead + h) * nsplit + sp) * M * (DH + 2);
    if (m_lo < M) {
        float* p = base + (size_t)m_lo * (DH + 2);
        if (tig == 0) { p[0] = mx_lo * 0.6931471805599453f; p[1] = l_lo; }      // back to natural-log units for the combine kernel
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) { p[2 + nt * 8 + tig * 2] = o[nt][0]; p[2 + nt * 8 + tig * 2 + 1] = o[nt][1]; }
    }
    if (m_hi < M) {
        float* p = base + (size_t)m_hi * (DH + 2);
        if (tig == 0) { p[0] = mx_hi * 0.6931471805599453f; p[1] = l_hi; }
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) { p[2 + nt * 8 + tig * 2] = o[nt][2]; p[2 + nt * 8 + tig * 2 + 1] = o[nt][3]; }
    }
}

template <int DH>
__global__ void attn_rows_combine_kernel(const float* __restrict__ part, float* __restrict__ O, int ldo, int slot,
                                         const int32_t* __restrict__ len, int M, int nhead, int nsplit) {
    const int h = blockIdx.x, b = blockIdx.y;
    const int len_b = len ? min(len[b], slot) : slot;
    const int ns = (len_b + SPLIT_ROWS - 1) / SPLIT_ROWS;
    for (int i = threadIdx.x; i < M * DH; i += blockDim.x) {
        const int m = i / DH, d = i % DH;
        const float* p0 = part + (((size_t)b * nhead + h) * nsplit * M + m) * (DH + 2);
        float gm = -INFINITY;
#pragma unroll 8
        for (int s = 0; s < ns; ++s) gm = fmaxf(gm, p0[(size_t)s * M * (DH + 2)]);      // (unrolled: the loads of the splits are independent)
        float l = 0.f, a = 0.f;
#pragma unroll 8
        for (int s = 0; s < ns; ++s) {
            const float* p = p0 + (size_t)s * M * (DH + 2);
            const float w = __expf(p[0] - gm);
            l = fmaf(p[1], w, l);
            a = fmaf(p[2 + d], w, a);
        }
        O[((size_t)b * M + m) * ldo + h * DH + d] = a / l;
    }
}

// ------------------------------------------------------------------------------------------------
// f2a direction of X2Y_map (models/basic.py:373-379): softmax over rows + weighted row sum.
// stats partial: [B][nsplit][M][2]; stats: [B][M][2] = (max, 1/sum).
__global__ void __launch_bounds__(256) col_stats_partial_kernel(const float* __restrict__ L, int ldl,
                                                                float* __restrict__ sp_out, int slot,
                                                                const int32_t* __restrict__ len, int M, int nsplit) {
    const int sp = blockIdx.x, b = blockIdx.y;
    const int len_b = len ? min(len[b], slot) : slot;
    const int r0 = sp * SPLIT_ROWS;
    if (r0 >= len_b) return;
    const int r1 = min(r0 + SPLIT_ROWS, len_b);
    for (int m = threadIdx.x; m < M; m += blockDim.x) {
        float mx = -INFINITY, l = 0.f;
        for (int t = r0; t < r1; ++t) {
            const float v = L[((size_t)b * slot + t) * ldl + m];
            const float nm = fmaxf(mx, v);
            l = l * __expf(mx - nm) + __expf(v - nm);
            mx = nm;
        }
        float* o = sp_out + (((size_t)b * nsplit + sp) * M + m) * 2;
        o[0] = mx;
        o[1] = l;
    }
}

__global__ void col_stats_combine_kernel(const float* __restrict__ sp_in, float* __restrict__ stats, int slot,
                                         const int32_t* __restrict__ len, int M, int nsplit) {
    const int b = blockIdx.x;
    const int len_b = len ? min(len[b], slot) : slot;
    const int ns = (len_b + SPLIT_ROWS - 1) / SPLIT_ROWS;
    for (int m = threadIdx.x; m < M; m += blockDim.x) {
        float gm = -INFINITY;
        for (int s = 0; s < ns; ++s) gm = fmaxf(gm, sp_in[(((size_t)b * nsplit + s) * M + m) * 2]);
        float l = 0.f;
        for (int s = 0; s < ns; ++s) {
            const float* p = sp_in + (((size_t)b * nsplit + s) * M + m) * 2;
            l = fmaf(p[1], __expf(p[0] - gm), l);
        }
        stats[((size_t)b * M + m) * 2] = gm;
        stats[((size_t)b * M + m) * 2 + 1] = 1.f / l;
    }
}

// partial weighted sums: part[b][split][m][e] = sum_{t in split} p[t,m] X[t,e]; 32 tokens x 128 channels per CTA.
__global__ void __launch_bounds__(256) col_apply_partial_kernel(const float* __restrict__ L, int ldl,
                                                                const float* __restrict__ stats,
                                                                const void* __restrict__ X, int x_dtype, int ldx,
                                                                float* __restrict__ part, float* __restrict__ P, int ldp,
                                                                int slot, const int32_t* __restrict__ len, int M, int E,
                                                                int nsplit, int etiles) {
    constexpr int TR = 32, TM = 32, TE = 128;
    __shared__ float Ps[TR][TM + 1];
    __shared__ __align__(16) float Xs[TR][TE];
    const int sp = blockIdx.x, b = blockIdx.z;
    const int mt = blockIdx.y / etiles, et = blockIdx.y % etiles;
    const int len_b = len ? min(len[b], slot) : slot;
    const int r0 = sp * SPLIT_ROWS;
    if (r0 >= len_b) return;
    const int r1 = min(r0 + SPLIT_ROWS, len_b);
    const int tid = threadIdx.x;
    const int tm = tid >> 5, te = tid & 31;   // 8 groups of 4 tokens, 32 groups of 4 channels
    const int m0 = mt * TM, e0 = et * TE;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int t0 = r0; t0 < r1; t0 += TR) {
        const int nr = min(TR, r1 - t0);
        __syncthreads();
        for (int i = tid; i < TR * TM; i += 256) {
            const int r = i / TM, mm = i % TM;
            float p = 0.f;
            if (r < nr && m0 + mm < M) {
                const size_t row = (size_t)b * slot + t0 + r;
                const float* st = stats + ((size_t)b * M + m0 + mm) * 2;
                p = __expf(L[row * ldl + m0 + mm] - st[0]) * st[1];
                if (P != nullptr && et == 0) P[row * ldp + m0 + mm] = p;
            }
            Ps[r][mm] = p;
        }
        for (int i = tid; i < TR * TE; i += 256) {
            const int r = i / TE, ee = i % TE;
            Xs[r][ee] = (r < nr && e0 + ee < E) ? ld_elem(X, x_dtype, ((size_t)b * slot + t0 + r) * ldx + e0 + ee) : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int r = 0; r < TR; ++r) {
            const float4 x4 = *reinterpret_cast<const float4*>(&Xs[r][te * 4]);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float p = Ps[r][tm * 4 + i];
                acc[i][0] = fmaf(p, x4.x, acc[i][0]); acc[i][1] = fmaf(p, x4.y, acc[i][1]);
                acc[i][2] = fmaf(p, x4.z, acc[i][2]); acc[i][3] = fmaf(p, x4.w, acc[i][3]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + tm * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int e = e0 + te * 4 + j;
            if (e < E) part[(((size_t)b * nsplit + sp) * M + m) * E + e] = acc[i][j];
        }
    }
}


// Tensor-core version of col_apply_partial_kernel for bf16 rows: part[b][split][m][e] = sum_{t in split} p[t,m] X[t,e] as
// mma.sync m16n8k16 (M = tokens, N = channels, K = rows of the split).  One CTA = 80 tokens x 256 channels x one split;
// the probabilities are formed per 64-row stage as bf16 in shared memory (token-major, the A operand), the rows tile is
// staged as is (row-major [t][e], read through ldmatrix.trans as the B operand); fp32 accumulators in registers.
constexpr int CAM_MT = 5, CAM_TOK = CAM_MT * 16, CAM_E = 256, CAM_TR = 64;
constexpr int CAM_XS = CAM_E * 2 + 16;      // bytes per staged row (padded: conflict-free ldmatrix)
constexpr int CAM_PS = CAM_TR * 2 + 16;     // bytes per token row of the probability tile
constexpr int CAM_SMEM = 2 * CAM_TR * CAM_XS + CAM_TOK * CAM_PS + 2 * CAM_TOK * 4;
__global__ void __launch_bounds__(256) col_apply_mma_kernel(const float* __restrict__ L, int ldl, const float* __restrict__ stats,
                                                            const __nv_bfloat16* __restrict__ X, int ldx, float* __restrict__ part,
                                                            float* __restrict__ P, int ldp, int slot,
                                                            const int32_t* __restrict__ len, int M, int E, int nsplit, int etiles) {
    // The rows tile is double-buffered with cp.async and the logits of the next stage are prefetched into registers, so the
    // global loads of stage s+1 are in flight under the MMAs of stage s (ncu: the single-buffered version stalled 67 % of its
    // issue slots on the loads, at 2 CTAs per SM).
    extern __shared__ __align__(16) uint8_t cam_smem[];
    uint8_t* Xs = cam_smem;                                      // [2][64 rows][CAM_XS]
    uint8_t* Pt = cam_smem + 2 * CAM_TR * CAM_XS;                // [80 tokens][CAM_PS]
    float* st_mx = reinterpret_cast<float*>(Pt + CAM_TOK * CAM_PS);
    float* st_inv = st_mx + CAM_TOK;
    const int sp = blockIdx.x, b = blockIdx.z;
    const int mc = blockIdx.y / etiles, et = blockIdx.y % etiles;
    const int len_b = len ? min(len[b], slot) : slot;
    const int r0 = sp * SPLIT_ROWS;
    if (r0 >= len_b) return;
    const int r1 = min(r0 + SPLIT_ROWS, len_b);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int m0 = mc * CAM_TOK, e0 = et * CAM_E;
    if (tid < CAM_TOK) {
        const bool ok = m0 + tid < M;
        st_mx[tid] = ok ? stats[((size_t)b * M + m0 + tid) * 2] : 0.f;
        st_inv[tid] = ok ? stats[((size_t)b * M + m0 + tid) * 2 + 1] : 0.f;
    }
    float acc[CAM_MT][4][4];
#pragma unroll
    for (int i = 0; i < CAM_MT; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.f;
    const uint32_t xs_u32 = (uint32_t)__cvta_generic_to_shared(Xs), pt_u32 = (uint32_t)__cvta_generic_to_shared(Pt);
    // ldmatrix lane addresses: A (tokens x rows) from Pt, B (rows x channels, transposed on load) from Xs
    const uint32_t a_lane = pt_u32 + (uint32_t)((lane & 7) + ((lane >> 3) & 1) * 8) * CAM_PS + (uint32_t)(lane >> 4) * 16u;
    const uint32_t b_lane = xs_u32 + (uint32_t)((lane & 7) + ((lane >> 3) & 1) * 8) * CAM_XS + (uint32_t)(w * 32 + (lane >> 4) * 8) * 2u;

    // logits: thread -> one token (tid % 80; threads 240..255 idle) and rows tid / 80 + 3 k: cheap incremental addressing
    constexpr int LPT = (CAM_TR + 2) / 3;                        // 22 rows per thread and stage
    const int lmm = tid % CAM_TOK, lr0 = tid / CAM_TOK;
    const bool lactive = tid < 3 * CAM_TOK && m0 + lmm < M;
    float lreg[LPT];
    auto load_logits = [&](int t0) {
        const int nr = min(CAM_TR, r1 - t0);
        const float* lp = L + ((size_t)b * slot + t0 + lr0) * ldl + m0 + lmm;
#pragma unroll
        for (int k = 0; k < LPT; ++k) lreg[k] = (lactive && lr0 + 3 * k < nr) ? lp[(size_t)(3 * k) * ldl] : -INFINITY;
    };
    auto issue_rows = [&](int t0, int buf) {                     // 64 rows x 256 channels, 16-byte cp.async; rows past the end zero-filled
        const int nr = min(CAM_TR, r1 - t0);
#pragma unroll
        for (int k = 0; k < CAM_TR * (CAM_E / 8) / 256; ++k) {
            const int i = tid + k * 256, r = i / (CAM_E / 8), c = i % (CAM_E / 8);
            const bool ok = r < nr && e0 + c * 8 < E;
            const __nv_bfloat16* src = X + ((size_t)b * slot + t0 + (ok ? r : 0)) * ldx + (ok ? e0 + c * 8 : 0);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(xs_u32 + (uint32_t)(buf * CAM_TR * CAM_XS + r * CAM_XS + c * 16)),
                         "l"(src), "r"(ok ? 16 : 0) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    issue_rows(r0, 0);
    load_logits(r0);
    __syncthreads();                                             // stats in shared memory
    int buf = 0;
    for (int t0 = r0; t0 < r1; t0 += CAM_TR, buf ^= 1) {
        const int nr = min(CAM_TR, r1 - t0);
        // probabilities of this stage, bf16, token-major (the previous stage's MMAs are done: trailing barrier)
        if (tid < 3 * CAM_TOK) {
            const float smx = st_mx[lmm], sinv = st_inv[lmm];
#pragma unroll
            for (int k = 0; k < LPT; ++k) {
                const int r = lr0 + 3 * k;
                if (r < CAM_TR) {
                    const float p = __expf(lreg[k] - smx) * sinv;            // exp(-inf) = 0 past the end / beyond M
                    if (P != nullptr && et == 0 && r < nr && lactive) P[((size_t)b * slot + t0 + r) * ldp + m0 + lmm] = p;
                    *reinterpret_cast<__nv_bfloat16*>(Pt + lmm * CAM_PS + r * 2) = __float2bfloat16_rn(p);
                }
            }
        }
        const bool more = t0 + CAM_TR < r1;
        if (more) {
            issue_rows(t0 + CAM_TR, buf ^ 1);
            load_logits(t0 + CAM_TR);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        const uint32_t b_cur = b_lane + (uint32_t)(buf * CAM_TR * CAM_XS);
#pragma unroll
        for (int ks = 0; ks < CAM_TR / 16; ++ks) {
            uint32_t bf[4][2];
#pragma unroll
            for (int jj = 0; jj < 2; ++jj)
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                             : "=r"(bf[2 * jj][0]), "=r"(bf[2 * jj][1]), "=r"(bf[2 * jj + 1][0]), "=r"(bf[2 * jj + 1][1])
                             : "r"(b_cur + (uint32_t)(ks * 16) * CAM_XS + (uint32_t)jj * 32u));
#pragma unroll
            for (int i = 0; i < CAM_MT; ++i) {
                uint32_t a[4];
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                             : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3])
                             : "r"(a_lane + (uint32_t)(i * 16) * CAM_PS + (uint32_t)ks * 32u));
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                                 : "+f"(acc[i][j][0]), "+f"(acc[i][j][1]), "+f"(acc[i][j][2]), "+f"(acc[i][j][3])
                                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(bf[j][0]), "r"(bf[j][1]));
            }
        }
        __syncthreads();                                         // Pt and the other rows buffer are free again
    }
#pragma unroll
    for (int i = 0; i < CAM_MT; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int e = e0 + w * 32 + j * 8 + (lane & 3) * 2;
            if (e >= E) continue;
#pragma unroll
            for (int hrow = 0; hrow < 2; ++hrow) {
                const int m = m0 + i * 16 + (lane >> 2) + hrow * 8;
                if (m < M)
                    *reinterpret_cast<float2*>(part + (((size_t)b * nsplit + sp) * M + m) * E + e) =
                        make_float2(acc[i][j][2 * hrow], acc[i][j][2 * hrow + 1]);
            }
        }
}

__global__ void col_apply_combine_kernel(const float* __restrict__ part, float* __restrict__ out, int ldo, int slot,
                                         const int32_t* __restrict__ len, int M, int E, int nsplit) {
    const int b = blockIdx.y;
    const int len_b = len ? min(len[b], slot) : slot;
    const int ns = (len_b + SPLIT_ROWS - 1) / SPLIT_ROWS;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * E) return;
    const int m = i / E, e = i % E;
    float a = 0.f;
    for (int s = 0; s < ns; ++s) a += part[(((size_t)b * nsplit + s) * M + m) * E + e];
    out[((size_t)b * M + m) * ldo + e] = a;
}

static inline int nsplit_of(int slot) { return (slot + SPLIT_ROWS - 1) / SPLIT_ROWS; }

// attn_tc.cu
bool attn_rows_tc_ok(const void* Kx, const void* Vx, int kv_dtype, int ldkv, int slot, int nhead, int dh);
int attn_rows_tc_launch(const float* Q, int ldq, const void* Kx, const void* Vx, int ldkv, int B, int slot, const int32_t* len, int M,
                        int nsplit, float* ws, cudaStream_t st);

}  // namespace factk

using namespace factk;

extern "C" int factk_mha_tokens(const float* Q, const float* K, const float* V, int ld, float* O, int ldo, int B, int M,
                                int nhead, int dh, int tf32, void* stream) {
    FACTK_REQUIRE(Q && K && V && O && B > 0 && M > 0 && nhead > 0, "factk_mha_tokens: bad args");
    if (tf32 && (dh == 16 || dh == 32 || dh == 64) && (ldo % 2) == 0 && (reinterpret_cast<uintptr_t>(O) & 7u) == 0) {
        const int Mp = (M + 7) & ~7;
        const size_t smem_t = (size_t)2 * Mp * (dh + 4) * sizeof(float);
        FACTK_REQUIRE(smem_t <= 200 * 1024, "factk_mha_tokens: M*dh too large for shared memory (%zu B)", smem_t);
        const dim3 tgrid(nhead, B);
#define LAUNCH_T(DH)                                                                                                \
    do {                                                                                                            \
        cudaFuncSetAttribute(mha_tokens_mma_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);   \
        mha_tokens_mma_kernel<DH><<<tgrid, MT_WARPS * 32, smem_t, (cudaStream_t)stream>>>(Q, K, V, ld, O, ldo, M);  \
    } while (0)
        if (dh == 16) LAUNCH_T(16);
        else if (dh == 32) LAUNCH_T(32);
        else LAUNCH_T(64);
#undef LAUNCH_T
        return check_launch("factk_mha_tokens(tf32)");
    }
    const size_t smem = (size_t)2 * M * dh * sizeof(float);
    FACTK_REQUIRE(smem <= 200 * 1024, "factk_mha_tokens: M*dh too large for shared memory (%zu B)", smem);
    dim3 grid(nhead, B);
#define LAUNCH(DH)                                                                                              \
    do {                                                                                                        \
        cudaFuncSetAttribute(mha_tokens_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);   \
        mha_tokens_kernel<DH><<<grid, 128, smem, (cudaStream_t)stream>>>(Q, K, V, ld, O, ldo, M);               \
    } while (0)
    switch (dh) {
        case 4: LAUNCH(4); break;
        case 8: LAUNCH(8); break;
        case 16: LAUNCH(16); break;
        case 32: LAUNCH(32); break;
        case 64: LAUNCH(64); break;
        default: FACTK_REQUIRE(false, "factk_mha_tokens: head dim %d unsupported (4/8/16/32/64)", dh);
    }
#undef LAUNCH
    return check_launch("factk_mha_tokens");
}

extern "C" size_t factk_attn_rows_ws_floats(int B, int slot, int M, int nhead, int dh) {
    return (size_t)B * nhead * nsplit_of(slot) * M * (dh + 2);
}

extern "C" int factk_attn_rows(const float* Q, int ldq, const void* Kx, const void* Vx, int kv_dtype, int ldkv, float* O,
                               int ldo, int B, int slot, const int32_t* len, int M, int nhead, int dh, float* ws,
                               void* stream) {
    FACTK_REQUIRE(Q && Kx && Vx && O && ws && B > 0 && slot > 0 && M > 0, "factk_attn_rows: bad args");
    FACTK_REQUIRE(M <= 320, "factk_attn_rows: at most 320 tokens (got %d)", M);
    const int ns = nsplit_of(slot);
    dim3 grid(ns, nhead, B), cgrid(nhead, B);
    const int threads = ((M + 31) / 32) * 32;
    cudaStream_t st_ = (cudaStream_t)stream;
    const bool kv16 = ((reinterpret_cast<uintptr_t>(Kx) | reinterpret_cast<uintptr_t>(Vx)) & 15u) == 0 && (ldkv % 8) == 0;
    // tcgen05 / TMEM kernel (attn_tc.cu): 8 heads of 32 channels, bf16 rows.  Measured on B200 (profiles/r2_attn_tc.md): with
    // M = 300 tokens (three full 128-row query blocks) it matches the mma.sync kernel below (585 vs 598 us at 8 x 16384 frames);
    // with M = 75 the 128-row MMA wastes 41 % of the softmax lanes -- the kernel is bound by exp2 throughput and by the latency
    // of two warps per scheduler, not by the tensor pipe -- and it is slower (264 vs 175 us at 64 x 4096), so it is the default
    // for M > 128 only.  FACTK_ATTN_TC=1 forces it, FACTK_ATTN_TC=0 disables it.
    static const int tc_mode = [] { const char* e = getenv("FACTK_ATTN_TC"); return e ? atoi(e) : -1; }();
    const bool use_tc = tc_mode == 1 || (tc_mode < 0 && M > 128);
    if (use_tc && attn_rows_tc_ok(Kx, Vx, kv_dtype, ldkv, slot, nhead, dh)) {
        const int rc = attn_rows_tc_launch(Q, ldq, Kx, Vx, ldkv, B, slot, len, M, ns, ws, st_);
        if (rc) return rc;
        attn_rows_combine_kernel<32><<<cgrid, 256, 0, st_>>>(ws, O, ldo, slot, len, M, nhead, ns);
        return check_launch("factk_attn_rows(tcgen05)");
    }
    if (kv_dtype == FACTK_BF16 && kv16 && (dh == 16 || dh == 32 || dh == 64)) {
        const int nqb = (M + 127) / 128;
        int qwarps = nqb > 1 ? 8 : (M + 15) / 16;    // a query block is 128 tokens = 8 warps
        if (qwarps < 4) qwarps = 4;                  // extra warps only help staging the K/V tiles
        const int wthreads = qwarps * 32;
        const dim3 mgrid(ns * nqb, nhead, B);
        const __nv_bfloat16* K16 = reinterpret_cast<const __nv_bfloat16*>(Kx);
        const __nv_bfloat16* V16 = reinterpret_cast<const __nv_bfloat16*>(Vx);
        // FACTK_ATTN_QSPLIT=1: queries as hi + lo bf16 pairs (logits carry only the keys' rounding) at +50 % QK^T tensor work
        static const bool qsplit = [] { const char* e = getenv("FACTK_ATTN_QSPLIT"); return e && atoi(e) != 0; }();
#define LAUNCH_MMA(DH)                                                                                                  \
    do {                                                                                                                \
        if (qsplit) attn_rows_mma_kernel<DH, true><<<mgrid, wthreads, 0, st_>>>(Q, ldq, K16, V16, ldkv, ws, slot, len, M, nhead, ns);   \
        else attn_rows_mma_kernel<DH, false><<<mgrid, wthreads, 0, st_>>>(Q, ldq, K16, V16, ldkv, ws, slot, len, M, nhead, ns);             \
        attn_rows_combine_kernel<DH><<<cgrid, 256, 0, st_>>>(ws, O, ldo, slot, len, M, nhead, ns);                       \
    } while (0)
        if (dh == 16) LAUNCH_MMA(16); else if (dh == 32) LAUNCH_MMA(32); else LAUNCH_MMA(64);
#undef LAUNCH_MMA
        return check_launch("factk_attn_rows");
    }
#define LAUNCH(DH)                                                                                                     \
    do {                                                                                                               \
        attn_rows_partial_kernel<DH><<<grid, threads, 0, (cudaStream_t)stream>>>(Q, ldq, Kx, Vx, kv_dtype, ldkv, ws,    \
                                                                                 slot, len, M, nhead, ns);             \
        attn_rows_combine_kernel<DH><<<cgrid, 256, 0, (cudaStream_t)stream>>>(ws, O, ldo, slot, len, M, nhead, ns);     \
    } while (0)
    switch (dh) {
        case 4: LAUNCH(4); break;
        case 8: LAUNCH(8); break;
        case 16: LAUNCH(16); break;
        case 32: LAUNCH(32); break;
        case 64: LAUNCH(64); break;
        default: FACTK_REQUIRE(false, "factk_attn_rows: head dim %d unsupported (4/8/16/32/64)", dh);
    }
#undef LAUNCH
    return check_launch("factk_attn_rows");
}

extern "C" size_t factk_col_softmax_ws_floats(int B, int slot, int M, int E) {
    const size_t ns = nsplit_of(slot);
    return (size_t)B * ns * M * 2 + (size_t)B * M * 2 + (size_t)B * ns * M * E;
}

extern "C" int factk_col_softmax_apply(const float* L, int ldl, const void* X, int x_dtype, int ldx, float* out, int ldo,
                                       float* P, int ldp, int B, int slot, const int32_t* len, int M, int E, float* ws,
                                       void* stream) {
    FACTK_REQUIRE(L && X && out && ws && B > 0 && slot > 0 && M > 0 && E > 0, "factk_col_softmax_apply: bad args");
    const int ns = nsplit_of(slot);
    float* sp = ws;
    float* stats = sp + (size_t)B * ns * M * 2;
    float* part = stats + (size_t)B * M * 2;
    cudaStream_t st = (cudaStream_t)stream;
    col_stats_partial_kernel<<<dim3(ns, B), 256, 0, st>>>(L, ldl, sp, slot, len, M, ns);
    col_stats_combine_kernel<<<B, 256, 0, st>>>(sp, stats, slot, len, M, ns);
    if (x_dtype == FACTK_BF16 && (E % 8) == 0 && (ldx % 8) == 0 && aligned16(X)) {
        const int etiles = (E + CAM_E - 1) / CAM_E, mchunks = (M + CAM_TOK - 1) / CAM_TOK;
        static unsigned long long cam_devs = 0;
        if (first_use_on_device(cam_devs))
            cudaFuncSetAttribute(col_apply_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CAM_SMEM);
        col_apply_mma_kernel<<<dim3(ns, mchunks * etiles, B), 256, CAM_SMEM, st>>>(L, ldl, stats, reinterpret_cast<const __nv_bfloat16*>(X), ldx,
                                                                           part, P, ldp, slot, len, M, E, ns, etiles);
    } else {
        const int etiles = (E + 127) / 128, mtiles = (M + 31) / 32;
        col_apply_partial_kernel<<<dim3(ns, mtiles * etiles, B), 256, 0, st>>>(L, ldl, stats, X, x_dtype, ldx, part, P, ldp,
                                                                               slot, len, M, E, ns, etiles);
    }
    col_apply_combine_kernel<<<dim3((M * E + 255) / 256, B), 256, 0, st>>>(part, out, ldo, slot, len, M, E, ns);
    return check_launch("factk_col_softmax_apply");
}
