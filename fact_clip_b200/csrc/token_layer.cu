// One launch per token-side decoder layer (half-layer for SCALayer, whose cross attention over the frames sits in the middle):
//   [self-attention (packed in_proj -> per-head softmax(q k^T) v) ->] out_proj + residual -> LayerNorm
//   [-> cross-attention query projection]  [-> FFN (linear1, ReLU, linear2) + residual -> LayerNorm]
// replacing the 8-11 dependent launches per layer of the unfused path (models/basic.py:429-452 SALayer.forward,
// 494-523 SCALayer.forward).  One 2-CTA cluster per video: the <= 80 x A token matrix never leaves the two SMs between the
// stages -- fp32 pre-norm rows and bf16 GEMM operands in shared memory, every CTA holding the full rows -- and each CTA
// computes half of the output columns of every GEMM (its half of the heads in the attention), writing them into its own and,
// through distributed shared memory, its peer's copy; a cluster barrier separates the stages.  The weights stream from L2
// once per CTA pair, pre-packed on the host side in mma.m16n8k16 B-fragment order (one coalesced 16-byte load per lane and
// k-step, prefetched seven k-steps ahead; no shared-memory staging of weights).  bf16 operands, fp32 accumulate, fp32
// residual, LayerNorm and softmax.  The token side is latency-bound (75 rows per video, measured: the legacy tensor pipe
// retires one m16n8k16 per ~4.4 cycles per SM, which is what bounds a stage): the point of this kernel is the launch count,
// the L2 round trips between dependent launches, and spreading 64 videos over 128 SMs.
#include "common.cuh"

namespace factk {

__device__ long long* tl_dbg = nullptr;       // factk_token_layer_debug: clock64 at the stage boundaries of CTA 0
#define TL_MARK(i) do { if (tl_dbg != nullptr && blockIdx.x == 0 && threadIdx.x == 0) tl_dbg[i] = clock64(); } while (0)

constexpr int TL_NT = 256, TL_WARPS = TL_NT / 32, TL_MAXM = 80, TL_MAXMT = TL_MAXM / 16;

__device__ __forceinline__ void tl_ldsm4(uint32_t (&r)[4], const void* p) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void tl_ldsm4t(uint32_t (&r)[4], const void* p) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void tl_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t tl_saddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t tl_cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t tl_mapa(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void tl_st_peer_b32(uint32_t addr, uint32_t v) { asm volatile("st.shared::cluster.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void tl_st_peer_f2(uint32_t addr, float a, float b) {
    asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(a), "f"(b) : "memory");
}
// stage boundary: every thread of both CTAs; orders the (distributed) shared-memory writes of the stage before the reads of the next
__device__ __forceinline__ void tl_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float tl_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t tl_pack(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

// C[m][n] = sum_k A[m][k] W[n][k] for the MT 16-row tiles of the bf16 operand in shared memory and the columns
// sec * sec_stride + col0 + [0, ncols) of nsec sections (this CTA's share); W packed per (32-column group, k-step of 16, lane)
// as 2 x 4 registers, a warp task being 16 columns (one of the two uint4).  Each warp owns the tasks warp, warp + 8, ...  The
// epilogue is two functors over the two adjacent columns (col, col + 1) of one row: add(row, col) loads the addend (bias,
// position table, residual) -- all of a task's addends are fetched before its k loop, so their (global memory) latencies hide
// under the MMAs instead of serialising behind possibly-aliasing stores -- and put(row, col, v0, v1) stores.
// K % 64 == 0, col0 % 16 == 0, ncols % 16 == 0.
template <int RING, typename Add, typename Put>
__device__ __forceinline__ void tl_gemm_r(const __nv_bfloat16* Ab, int lda, int MT, int K, const uint4* __restrict__ Wp, int nsec, int sec_stride,
                                          int col0, int ncols, Add add, Put put) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int KS = K >> 4, tps = ncols >> 4, ntask = nsec * tps;
    const __nv_bfloat16* arow = Ab + (size_t)(lane & 15) * lda + (lane >> 4) * 8;
    for (int task = warp; task < ntask; task += TL_WARPS) {
        const int cbase = (task / tps) * sec_stride + col0 + (task % tps) * 16;
        float acc[TL_MAXMT][2][4];
#pragma unroll
        for (int mt = 0; mt < TL_MAXMT; ++mt)
#pragma unroll
            for (int t = 0; t < 2; ++t) acc[mt][t][0] = acc[mt][t][1] = acc[mt][t][2] = acc[mt][t][3] = 0.f;
        const uint4* wp = Wp + ((size_t)(cbase >> 5) * KS * 32 + lane) * 2 + ((cbase >> 4) & 1);          // + 64 uint4 per k-step
        uint4 q[RING];               // weight fragments RING - 1 k-steps ahead (L2 latency ~ 800 cycles, a k-step ~ 100-200)
#pragma unroll
        for (int s = 0; s < RING - 1; ++s) q[s] = __ldg(wp + s * 64);
        // the epilogue's addends (bias, tables, residual) are fetched now: their latency hides under the k loop
        const int g = lane >> 2, c2 = (lane & 3) * 2;
        float2 ad[TL_MAXMT][2][2];
#pragma unroll
        for (int mt = 0; mt < TL_MAXMT; ++mt) {
            if (mt < MT) {
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    ad[mt][t][0] = add(mt * 16 + g, cbase + t * 8 + c2);
                    ad[mt][t][1] = add(mt * 16 + g + 8, cbase + t * 8 + c2);
                }
            }
        }
        for (int ks0 = 0; ks0 < KS; ks0 += RING) {
#pragma unroll
            for (int s = 0; s < RING; ++s) {
                const int ks = ks0 + s;
                if (ks + RING - 1 < KS) q[(s + RING - 1) & (RING - 1)] = __ldg(wp + (ks + RING - 1) * 64);
                const uint32_t bf[4] = {q[s].x, q[s].y, q[s].z, q[s].w};
                uint32_t a[TL_MAXMT][4];          // every A fragment of the k-step first: the ldmatrix latencies overlap
#pragma unroll
                for (int mt = 0; mt < TL_MAXMT; ++mt)
                    if (mt < MT) tl_ldsm4(a[mt], arow + (size_t)mt * 16 * lda + ks * 16);
#pragma unroll
                for (int mt = 0; mt < TL_MAXMT; ++mt) {
                    if (mt < MT) {
                        tl_mma(acc[mt][0], a[mt], bf[0], bf[1]);
                        tl_mma(acc[mt][1], a[mt], bf[2], bf[3]);
                    }
                }
            }
        }
#pragma unroll
        for (int mt = 0; mt < TL_MAXMT; ++mt) {
            if (mt < MT) {
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    put(mt * 16 + g, cbase + t * 8 + c2, acc[mt][t][0] + ad[mt][t][0].x, acc[mt][t][1] + ad[mt][t][0].y);
                    put(mt * 16 + g + 8, cbase + t * 8 + c2, acc[mt][t][2] + ad[mt][t][1].x, acc[mt][t][3] + ad[mt][t][1].y);
                }
            }
        }
    }
}

template <typename Add, typename Put>
__device__ __forceinline__ void tl_gemm(const __nv_bfloat16* Ab, int lda, int MT, int K, const uint4* __restrict__ Wp, int nsec, int sec_stride,
                                        int col0, int ncols, Add add, Put put) {
    if ((K & 127) == 0) tl_gemm_r<8>(Ab, lda, MT, K, Wp, nsec, sec_stride, col0, ncols, add, put);      // K / 16 % 8 == 0
    else tl_gemm_r<4>(Ab, lda, MT, K, Wp, nsec, sec_stride, col0, ncols, add, put);
}

// Per-head softmax(q k^T / sqrt(dh)) v for the heads h0 .. h0 + nh - 1, q | k | v the column bands of the bf16 rows in shared
// memory; the (head, 16-query tile) units are dealt to the warps; the output rows go to this CTA's and the peer's copy.
template <int DH>
__device__ __forceinline__ void tl_attention(const __nv_bfloat16* qkv, int ldq, int M, int MT, int A, int h0, int nh, __nv_bfloat16* ob, int ldo,
                                             uint32_t peer_off) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, c2 = (lane & 3) * 2;
    const float sl2 = rsqrtf((float)DH) * 1.4426950408889634f;
    for (int unit = warp; unit < nh * MT; unit += TL_WARPS) {
        const int h = h0 + unit / MT, mt = unit % MT;
        const __nv_bfloat16* Q = qkv + h * DH;
        const __nv_bfloat16* Kp = qkv + A + h * DH;
        const __nv_bfloat16* V = qkv + 2 * A + h * DH;
        {
            float s[2 * TL_MAXMT][4];
#pragma unroll
            for (int j = 0; j < 2 * TL_MAXMT; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
            for (int kd = 0; kd < DH / 16; ++kd) {
                uint32_t a[4];
                tl_ldsm4(a, Q + (size_t)(mt * 16 + (lane & 15)) * ldq + kd * 16 + (lane >> 4) * 8);
#pragma unroll
                for (int nt = 0; nt < TL_MAXMT; ++nt) {
                    if (nt < MT) {
                        // keys nt*16 .. +15: matrices (keys 0-7, k 0-7), (keys 0-7, k 8-15), (keys 8-15, k 0-7), (keys 8-15, k 8-15)
                        uint32_t r[4];
                        tl_ldsm4(r, Kp + (size_t)(nt * 16 + ((lane >> 4) << 3) + (lane & 7)) * ldq + kd * 16 + ((lane >> 3) & 1) * 8);
                        tl_mma(s[2 * nt], a, r[0], r[1]);
                        tl_mma(s[2 * nt + 1], a, r[2], r[3]);
                    }
                }
            }
            float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
            for (int j = 0; j < 2 * TL_MAXMT; ++j) {
                if (j < 2 * MT) {
                    const int key = j * 8 + c2;
                    if (key >= M) s[j][0] = s[j][2] = -INFINITY;
                    if (key + 1 >= M) s[j][1] = s[j][3] = -INFINITY;
                    m0 = fmaxf(m0, fmaxf(s[j][0], s[j][1]));
                    m1 = fmaxf(m1, fmaxf(s[j][2], s[j][3]));
                }
            }
            m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
            m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
            m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
            m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
            float l0 = 0.f, l1 = 0.f;
#pragma unroll
            for (int j = 0; j < 2 * TL_MAXMT; ++j) {
                if (j < 2 * MT) {
                    s[j][0] = tl_ex2((s[j][0] - m0) * sl2); s[j][1] = tl_ex2((s[j][1] - m0) * sl2);
                    s[j][2] = tl_ex2((s[j][2] - m1) * sl2); s[j][3] = tl_ex2((s[j][3] - m1) * sl2);
                    l0 += s[j][0] + s[j][1];
                    l1 += s[j][2] + s[j][3];
                }
            }
            l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
            l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
            l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
            l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
            float o[DH / 8][4];
#pragma unroll
            for (int n = 0; n < DH / 8; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
#pragma unroll
            for (int kt = 0; kt < TL_MAXMT; ++kt) {
                if (kt < MT) {
                    const uint32_t pa[4] = {tl_pack(s[2 * kt][0], s[2 * kt][1]), tl_pack(s[2 * kt][2], s[2 * kt][3]),
                                            tl_pack(s[2 * kt + 1][0], s[2 * kt + 1][1]), tl_pack(s[2 * kt + 1][2], s[2 * kt + 1][3])};
#pragma unroll
                    for (int dt = 0; dt < DH / 16; ++dt) {
                        // values of tokens kt*16 .. +15, channels dt*16 .. +15, transposed on load: (tok 0-7, d 0-7), (tok 8-15, d 0-7),
                        // (tok 0-7, d 8-15), (tok 8-15, d 8-15)
                        uint32_t r[4];
                        tl_ldsm4t(r, V + (size_t)(kt * 16 + ((lane >> 3) & 1) * 8 + (lane & 7)) * ldq + dt * 16 + (lane >> 4) * 8);
                        tl_mma(o[2 * dt], pa, r[0], r[1]);
                        tl_mma(o[2 * dt + 1], pa, r[2], r[3]);
                    }
                }
            }
            const float i0 = 1.f / l0, i1 = 1.f / l1;
            __nv_bfloat16* o0 = ob + (size_t)(mt * 16 + g) * ldo + h * DH + c2;
            __nv_bfloat16* o1 = o0 + (size_t)8 * ldo;
#pragma unroll
            for (int n = 0; n < DH / 8; ++n) {
                const uint32_t v0 = tl_pack(o[n][0] * i0, o[n][1] * i0), v1 = tl_pack(o[n][2] * i1, o[n][3] * i1);
                *reinterpret_cast<uint32_t*>(o0 + n * 8) = v0;
                *reinterpret_cast<uint32_t*>(o1 + n * 8) = v1;
                tl_st_peer_b32(tl_saddr(o0 + n * 8) + peer_off, v0);
                tl_st_peer_b32(tl_saddr(o1 + n * 8) + peer_off, v1);
            }
        }
    }
}

// LayerNorm of the fp32 rows in shared memory (in place), plus a bf16 copy as the next GEMM operand and / or the fp32 rows
// to global memory.  The rows are shared out between the two CTAs (m % 2 == rank), each result goes into both copies (16-byte
// distributed-shared-memory stores); one warp per row, a lane owns eight adjacent columns, two rows in flight per warp;
// two-pass mean / variance like rowops.cu's layernorm.  A cluster barrier must follow when `both`.
__device__ __forceinline__ void tl_layernorm(float* ts, int ldt, int M, int A, const float* __restrict__ w, const float* __restrict__ bia,
                                             float eps, __nv_bfloat16* ab, int ldab, float* gout, int rank, uint32_t peer_off, bool both) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, c0 = lane * 8;
    const bool on = c0 < A;                      // A % 64 == 0: a lane holds eight columns or none
    float wv[8], bv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        wv[i] = on ? __ldg(w + c0 + i) : 0.f;
        bv[i] = on ? __ldg(bia + c0 + i) : 0.f;
    }
    const float inv = 1.f / (float)A;
    for (int m0 = rank + 2 * warp; m0 < M; m0 += 4 * TL_WARPS) {
        float v[2][8], mean[2], rstd[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int m = min(m0 + 2 * TL_WARPS * u, M - 1);
            float4 x0 = make_float4(0.f, 0.f, 0.f, 0.f), x1 = x0;
            if (on) {
                x0 = *reinterpret_cast<const float4*>(ts + (size_t)m * ldt + c0);
                x1 = *reinterpret_cast<const float4*>(ts + (size_t)m * ldt + c0 + 4);
            }
            v[u][0] = x0.x; v[u][1] = x0.y; v[u][2] = x0.z; v[u][3] = x0.w;
            v[u][4] = x1.x; v[u][5] = x1.y; v[u][6] = x1.z; v[u][7] = x1.w;
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            float sum = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) sum += v[u][i];
            mean[u] = warp_sum(sum) * inv;
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            float sq = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float d = on ? v[u][i] - mean[u] : 0.f;
                sq += d * d;
            }
            rstd[u] = rsqrtf(warp_sum(sq) * inv + eps);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int m = m0 + 2 * TL_WARPS * u;
            if (m >= M || !on) continue;
            float y[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) y[i] = (v[u][i] - mean[u]) * rstd[u] * wv[i] + bv[i];
            if (both) {          // the last LayerNorm of a launch only feeds global memory (and the peer may have exited)
                float* t = ts + (size_t)m * ldt + c0;
                *reinterpret_cast<float4*>(t) = make_float4(y[0], y[1], y[2], y[3]);
                *reinterpret_cast<float4*>(t + 4) = make_float4(y[4], y[5], y[6], y[7]);
                const uint32_t ra = tl_saddr(t) + peer_off;
                asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(ra), "f"(y[0]), "f"(y[1]), "f"(y[2]), "f"(y[3]) : "memory");
                asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(ra + 16), "f"(y[4]), "f"(y[5]), "f"(y[6]), "f"(y[7]) : "memory");
                if (ab) {
                    const uint32_t p0 = tl_pack(y[0], y[1]), p1 = tl_pack(y[2], y[3]), p2 = tl_pack(y[4], y[5]), p3 = tl_pack(y[6], y[7]);
                    __nv_bfloat16* a = ab + (size_t)m * ldab + c0;
                    *reinterpret_cast<uint4*>(a) = make_uint4(p0, p1, p2, p3);
                    asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tl_saddr(a) + peer_off), "r"(p0), "r"(p1), "r"(p2), "r"(p3) : "memory");
                }
            }
            if (gout) {
                *reinterpret_cast<float4*>(gout + (size_t)m * A + c0) = make_float4(y[0], y[1], y[2], y[3]);
                *reinterpret_cast<float4*>(gout + (size_t)m * A + c0 + 4) = make_float4(y[4], y[5], y[6], y[7]);
            }
        }
    }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TL_NT, 1) token_layer_kernel(const __grid_constant__ factk_token_layer_t p) {
    extern __shared__ __align__(16) uint8_t tl_smem[];
    const int b = blockIdx.x >> 1, tid = threadIdx.x;
    const int rank = (int)tl_cluster_rank();
    const uint32_t peer_off = tl_mapa(tl_saddr(tl_smem), (uint32_t)(rank ^ 1)) - tl_saddr(tl_smem);
    const int M = p.M, A = p.A, ff = p.ff, MT = (M + 15) >> 4, rows = MT * 16, Ah = A >> 1;
    const int ldab = A + 8, ldq = 3 * A + 8, ldh = ff + 8;
    __nv_bfloat16* ab = reinterpret_cast<__nv_bfloat16*>(tl_smem);
    uint8_t* R = tl_smem + (size_t)TL_MAXM * ldab * 2;
    __nv_bfloat16* qkv = reinterpret_cast<__nv_bfloat16*>(R);
    __nv_bfloat16* ob = reinterpret_cast<__nv_bfloat16*>(R + (size_t)TL_MAXM * ldq * 2);
    float* ts = reinterpret_cast<float*>(R);                                          // pre-norm rows (after the attention: q|k|v are dead)
    __nv_bfloat16* hb = reinterpret_cast<__nv_bfloat16*>(R + (size_t)TL_MAXM * A * 4);   // FFN hidden rows (the attention output is dead)
    float* xg = p.x + (size_t)b * M * A;
    const bool has_self = p.w_in != nullptr;

    TL_MARK(0);
    // stage 0: the GEMM operand rows (tokens, or the cross attention's output for the tail half of an SCALayer) as bf16, padding rows
    // zero.  Local copy only (each CTA reads all rows: a distributed-shared-memory store needs the peer resident, i.e. the barrier below).
    {
        const float* src = has_self ? xg : p.o_in + (size_t)b * M * A;
        const int nv = rows * (A / 4);
        for (int i0 = tid; i0 < nv; i0 += 8 * TL_NT) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = i0 + u * TL_NT, m = i / (A / 4), c = (i % (A / 4)) * 4;
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (i < nv && m < M) v[u] = __ldg(reinterpret_cast<const float4*>(src + (size_t)m * A + c));
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = i0 + u * TL_NT, m = i / (A / 4), c = (i % (A / 4)) * 4;
                if (i < nv) {
                    uint2 w;
                    w.x = tl_pack(v[u].x, v[u].y);
                    w.y = tl_pack(v[u].z, v[u].w);
                    *reinterpret_cast<uint2*>(ab + (size_t)m * ldab + c) = w;
                }
            }
        }
    }
    tl_cluster_sync();              // also: the peer CTA is resident before the first distributed-shared-memory store
    TL_MARK(1);
    if (has_self) {
        // q | k | v of this CTA's half of the heads (three sections of A/2 columns): needed by this CTA's attention only
        const float* __restrict__ bin = p.b_in;
        const float* __restrict__ pre = p.pre_qk;
        tl_gemm(ab, ldab, MT, A, reinterpret_cast<const uint4*>(p.w_in), 3, A, rank * Ah, Ah,
                [&](int m, int n) {
                    float2 v = __ldg(reinterpret_cast<const float2*>(bin + n));
                    if (pre != nullptr && n < 2 * A && m < M) {
                        const float2 t = __ldg(reinterpret_cast<const float2*>(pre + (size_t)m * 2 * A + n));
                        v.x += t.x;
                        v.y += t.y;
                    }
                    return v;
                },
                [&](int m, int n, float v0, float v1) { *reinterpret_cast<uint32_t*>(qkv + (size_t)m * ldq + n) = tl_pack(v0, v1); });
        __syncthreads();
        TL_MARK(2);
        const int dh = A / p.nhead, nh = p.nhead >> 1;
        if (dh == 16) tl_attention<16>(qkv, ldq, M, MT, A, rank * nh, nh, ob, ldab, peer_off);
        else if (dh == 32) tl_attention<32>(qkv, ldq, M, MT, A, rank * nh, nh, ob, ldab, peer_off);
        else tl_attention<64>(qkv, ldq, M, MT, A, rank * nh, nh, ob, ldab, peer_off);
        tl_cluster_sync();
        TL_MARK(3);
    }
    // out_proj + residual -> pre-norm rows (this CTA's half of the columns, into both copies)
    {
        const float* __restrict__ bo = p.b_o;
        tl_gemm(has_self ? ob : ab, ldab, MT, A, reinterpret_cast<const uint4*>(p.w_o), 1, 0, rank * Ah, Ah,
                [&](int m, int n) {
                    float2 v = __ldg(reinterpret_cast<const float2*>(bo + n));
                    if (m < M) {
                        const float2 r = __ldg(reinterpret_cast<const float2*>(xg + (size_t)m * A + n));
                        v.x += r.x;
                        v.y += r.y;
                    }
                    return v;
                },
                [&](int m, int n, float v0, float v1) {
                    float* t = ts + (size_t)m * A + n;
                    *reinterpret_cast<float2*>(t) = make_float2(v0, v1);
                    tl_st_peer_f2(tl_saddr(t) + peer_off, v0, v1);
                });
    }
    tl_cluster_sync();
    TL_MARK(4);
    const bool has_ffn = p.w_1 != nullptr;
    tl_layernorm(ts, A, M, A, p.ln1_w, p.ln1_b, p.eps, ab, ldab, has_ffn ? nullptr : xg, rank, peer_off, true);
    tl_cluster_sync();
    TL_MARK(5);
    if (p.w_q != nullptr) {         // the cross attention's query projection of the normalised tokens (+ query positions)
        const float* __restrict__ bq = p.b_q;
        const float* __restrict__ pre = p.pre_q;
        float* cq = p.cq_out + (size_t)b * M * A;
        tl_gemm(ab, ldab, MT, A, reinterpret_cast<const uint4*>(p.w_q), 1, 0, rank * Ah, Ah,
                [&](int m, int n) {
                    float2 v = __ldg(reinterpret_cast<const float2*>(bq + n));
                    if (pre != nullptr && m < M) {
                        const float2 t = __ldg(reinterpret_cast<const float2*>(pre + (size_t)m * A + n));
                        v.x += t.x;
                        v.y += t.y;
                    }
                    return v;
                },
                [&](int m, int n, float v0, float v1) {
                    if (m < M) *reinterpret_cast<float2*>(cq + (size_t)m * A + n) = make_float2(v0, v1);
                });
    }
    TL_MARK(6);
    if (!has_ffn) return;
    {
        const float* __restrict__ b1 = p.b_1;
        tl_gemm(ab, ldab, MT, A, reinterpret_cast<const uint4*>(p.w_1), 1, 0, rank * (ff >> 1), ff >> 1,
                [&](int m, int n) { return __ldg(reinterpret_cast<const float2*>(b1 + n)); },
                [&](int m, int n, float v0, float v1) {
                    __nv_bfloat16* hp = hb + (size_t)m * ldh + n;
                    const uint32_t v = tl_pack(fmaxf(v0, 0.f), fmaxf(v1, 0.f));
                    *reinterpret_cast<uint32_t*>(hp) = v;
                    tl_st_peer_b32(tl_saddr(hp) + peer_off, v);
                });
    }
    tl_cluster_sync();
    TL_MARK(7);
    {
        const float* __restrict__ b2 = p.b_2;
        tl_gemm(hb, ldh, MT, ff, reinterpret_cast<const uint4*>(p.w_2), 1, 0, rank * Ah, Ah,
                [&](int m, int n) {
                    float2 v = __ldg(reinterpret_cast<const float2*>(b2 + n));
                    const float2 r = *reinterpret_cast<const float2*>(ts + (size_t)m * A + n);
                    v.x += r.x;
                    v.y += r.y;
                    return v;
                },
                [&](int m, int n, float v0, float v1) {
                    float* t = ts + (size_t)m * A + n;
                    *reinterpret_cast<float2*>(t) = make_float2(v0, v1);
                    tl_st_peer_f2(tl_saddr(t) + peer_off, v0, v1);
                });
    }
    tl_cluster_sync();
    TL_MARK(8);
    tl_layernorm(ts, A, M, A, p.ln2_w, p.ln2_b, p.eps, nullptr, 0, xg, rank, peer_off, false);
    TL_MARK(9);
}

static size_t tl_smem_bytes(int A, int ff) {
    const size_t ab = (size_t)TL_MAXM * (A + 8) * 2;
    const size_t r1 = (size_t)TL_MAXM * (3 * A + 8) * 2 + (size_t)TL_MAXM * (A + 8) * 2;
    const size_t r2 = (size_t)TL_MAXM * A * 4 + (size_t)TL_MAXM * (ff + 8) * 2;
    return ab + (r1 > r2 ? r1 : r2);
}

}  // namespace factk

/* Debug: clock64 of CTA 0 at the ten stage boundaries of the next launches into buf[10] (device memory); NULL turns it off. */
extern "C" int factk_token_layer_debug(void* buf) {
    long long* p = reinterpret_cast<long long*>(buf);
    cudaMemcpyToSymbol(factk::tl_dbg, &p, sizeof(p));
    return factk::check_launch("factk_token_layer_debug");
}

extern "C" int factk_token_layer_supported(int M, int A, int nhead, int ff) {
    using namespace factk;
    if (M < 1 || M > TL_MAXM || A < 64 || A > 256 || A % 64 != 0 || nhead < 2 || nhead % 2 != 0 || A % nhead != 0) return 0;
    const int dh = A / nhead;
    if (dh != 16 && dh != 32 && dh != 64) return 0;
    if (ff < 0 || ff % 64 != 0) return 0;
    return tl_smem_bytes(A, ff) <= 227 * 1024 ? 1 : 0;
}

extern "C" int factk_token_layer(const factk_token_layer_t* p, void* stream) {
    using namespace factk;
    FACTK_REQUIRE(p && p->x && p->B > 0, "factk_token_layer: bad args");
    FACTK_REQUIRE(factk_token_layer_supported(p->M, p->A, p->nhead, p->w_1 ? p->ff : 0), "factk_token_layer: unsupported shape M=%d A=%d heads=%d ff=%d",
                  p->M, p->A, p->nhead, p->ff);
    FACTK_REQUIRE(p->w_o && p->b_o && p->ln1_w && p->ln1_b, "factk_token_layer: out_proj / norm parameters missing");
    FACTK_REQUIRE((p->w_in != nullptr) != (p->o_in != nullptr), "factk_token_layer: exactly one of the self-attention weights and o_in");
    FACTK_REQUIRE(!p->w_in || p->b_in, "factk_token_layer: in_proj bias missing");
    FACTK_REQUIRE(!p->w_q || (p->b_q && p->cq_out), "factk_token_layer: query projection needs its bias and output");
    FACTK_REQUIRE(!p->w_1 || (p->b_1 && p->w_2 && p->b_2 && p->ln2_w && p->ln2_b && p->ff > 0), "factk_token_layer: FFN parameters missing");
    const size_t smem = tl_smem_bytes(p->A, p->w_1 ? p->ff : 0);
    static unsigned long long configured = 0;
    if (first_use_on_device(configured))
        cudaFuncSetAttribute(token_layer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    token_layer_kernel<<<2 * p->B, TL_NT, smem, (cudaStream_t)stream>>>(*p);      // one 2-CTA cluster per video
    return check_launch("factk_token_layer");
}
