// Generalised multi-source GEMM on CUDA cores (fp32 accumulate):
//   y[b,t,:N] = act(alpha * sum_s A_s[b, idx_s(t)] W_s[b]^T + bias) + res
// One kernel serves every Conv1d tap / Linear / concat-Linear of the FACT forward in fp32 mode and
// the small (token-side) products in bf16 mode.  Replaces the cuDNN/cuBLAS calls behind
// models/basic.py:138-139,158,177,182,343-347 and models/blocks.py:153-159,402,414.
// Tile 128x128x16, 256 threads, 8x8 register micro-tile, register prefetch + 2 smem buffers.
#include "common.cuh"

namespace factk {

constexpr int BM = 128, BN = 128, BK = 16, NT = 256, LDS = BM + 4;

__device__ __forceinline__ void load_a8(const factk_gemm_t& g, const factk_src_t& s, int b, int r_out, int len_b,
                                        int k, float a[8]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = 0.f;
    if (r_out >= len_b) return;
    long srow;
    if (s.gather) {
        srow = s.gather[(size_t)b * g.slot + r_out];
    } else {
        srow = (long)r_out + s.row_off;
        if (srow < 0 || srow >= len_b) return;
    }
    const size_t base = ((size_t)b * s.a_slot + (size_t)srow) * (size_t)s.lda;
    const bool vec = ((reinterpret_cast<uintptr_t>(s.A) & 15u) == 0) && ((s.lda & 3) == 0);
    if (vec && k + 8 <= s.K) {
        float4 v0 = ld_vec4(s.A, s.a_dtype, base + k), v1 = ld_vec4(s.A, s.a_dtype, base + k + 4);
        a[0] = v0.x; a[1] = v0.y; a[2] = v0.z; a[3] = v0.w;
        a[4] = v1.x; a[5] = v1.y; a[6] = v1.z; a[7] = v1.w;
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (k + j < s.K) a[j] = ld_elem(s.A, s.a_dtype, base + k + j);
    }
    if (s.pos != nullptr && k < s.pos_d) {
        const size_t pidx = s.pos_idx ? (size_t)s.pos_idx[(size_t)b * g.slot + r_out] : (size_t)r_out;
        const float* pr = s.pos + pidx * (size_t)s.pos_ld;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (k + j < s.pos_d && k + j < s.K) a[j] += pr[k + j];
    }
}

__device__ __forceinline__ void load_w8(const factk_gemm_t& g, const factk_src_t& s, int b, int n, int k, float w[8]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) w[j] = 0.f;
    if (n >= g.N) return;
    const float* wr = reinterpret_cast<const float*>(s.W) + (size_t)b * (size_t)s.w_bstride + (size_t)n * (size_t)s.ldw;
    const bool vec = ((reinterpret_cast<uintptr_t>(s.W) & 15u) == 0) && ((s.ldw & 3) == 0) && ((s.w_bstride & 3) == 0);
    if (vec && k + 8 <= s.K) {
        float4 v0 = *reinterpret_cast<const float4*>(wr + k), v1 = *reinterpret_cast<const float4*>(wr + k + 4);
        w[0] = v0.x; w[1] = v0.y; w[2] = v0.z; w[3] = v0.w;
        w[4] = v1.x; w[5] = v1.y; w[6] = v1.z; w[7] = v1.w;
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (k + j < s.K) w[j] = wr[k + j];
    }
}

__global__ void __launch_bounds__(NT) gemm_simt_kernel(const __grid_constant__ factk_gemm_t g, int tiles_per_video) {
    __shared__ __align__(16) float As[2][BK][LDS];
    __shared__ __align__(16) float Bs[2][BK][LDS];

    const int b = blockIdx.x / tiles_per_video;
    const int t0 = (blockIdx.x % tiles_per_video) * BM;
    const int n0 = blockIdx.y * BN;
    const int len_b = g.len ? min(g.len[b], g.slot) : g.slot;
    if (t0 >= len_b) return;

    const int tid = threadIdx.x;
    const int lrow = tid & 127, kseg = (tid >> 7) * 8;
    const int tx = tid & 15, ty = tid >> 4;

    int nchunk_total = 0;
    for (int s = 0; s < g.nsrc; ++s) nchunk_total += (g.src[s].K + BK - 1) / BK;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    float ra[8], rw[8];
    int cur_s = 0, cur_k = 0;   // source / k offset of the chunk being fetched
    auto fetch = [&]() {
        const factk_src_t& s = g.src[cur_s];
        load_a8(g, s, b, t0 + lrow, len_b, cur_k + kseg, ra);
        load_w8(g, s, b, n0 + lrow, cur_k + kseg, rw);
        cur_k += BK;
        if (cur_k >= s.K) { cur_k = 0; ++cur_s; }
    };
    auto stash = [&](int buf) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            As[buf][kseg + j][lrow] = ra[j];
            Bs[buf][kseg + j][lrow] = rw[j];
        }
    };

    fetch();
    stash(0);
    __syncthreads();
    for (int it = 0; it < nchunk_total; ++it) {
        const int buf = it & 1;
        const bool more = it + 1 < nchunk_total;
        if (more) fetch();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (more) stash(buf ^ 1);
        __syncthreads();
    }

    // epilogue
    const bool yvec = ((reinterpret_cast<uintptr_t>(g.Y) & 15u) == 0) && ((g.ldy & 3) == 0);
    const bool rvec = g.res && ((reinterpret_cast<uintptr_t>(g.res) & 15u) == 0) && ((g.ldres & 3) == 0);
    const float* bias = g.bias ? g.bias + (size_t)b * (size_t)g.bias_bstride : nullptr;
    const size_t pre_base = (size_t)b * (size_t)g.pre_bstride;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = t0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (r >= len_b) continue;
        const size_t yrow = ((size_t)b * g.slot + r) * (size_t)g.ldy;
        const size_t rrow = ((size_t)b * g.slot + r) * (size_t)g.ldres;
        const size_t prow = g.pre ? pre_base + (size_t)(g.pre_idx ? g.pre_idx[(size_t)b * g.slot + r] : r) * (size_t)g.ldpre : 0;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int c = n0 + h * 64 + tx * 4;
            if (c >= g.N) continue;
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float x = acc[i][h * 4 + j] * g.alpha;
                if (bias && c + j < g.N) x += bias[c + j];
                if (g.pre && c + j < g.N) x += ld_elem(g.pre, g.pre_dtype, prow + c + j);
                if (g.relu) x = fmaxf(x, 0.f);
                v[j] = x;
            }
            if (c + 4 <= g.N) {
                if (g.res) {
                    if (rvec) {
                        float4 rr = ld_vec4(g.res, g.res_dtype, rrow + c);
                        v[0] += rr.x; v[1] += rr.y; v[2] += rr.z; v[3] += rr.w;
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) v[j] += ld_elem(g.res, g.res_dtype, rrow + c + j);
                    }
                }
                if (yvec) {
                    st_vec4(g.Y, g.y_dtype, yrow + c, make_float4(v[0], v[1], v[2], v[3]));
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) st_elem(g.Y, g.y_dtype, yrow + c + j, v[j]);
                }
            } else {
                for (int j = 0; j < 4 && c + j < g.N; ++j) {
                    float x = v[j];
                    if (g.res) x += ld_elem(g.res, g.res_dtype, rrow + c + j);
                    st_elem(g.Y, g.y_dtype, yrow + c + j, x);
                }
            }
        }
    }
}

}  // namespace factk

extern "C" int factk_gemm(const factk_gemm_t* g, void* stream) {
    using namespace factk;
    FACTK_REQUIRE(g != nullptr, "factk_gemm: null descriptor");
    FACTK_REQUIRE(g->B > 0 && g->slot > 0 && g->N > 0, "factk_gemm: bad shape B=%d slot=%d N=%d", g->B, g->slot, g->N);
    FACTK_REQUIRE(g->nsrc >= 1 && g->nsrc <= FACTK_MAX_SRC, "factk_gemm: nsrc=%d out of range", g->nsrc);
    FACTK_REQUIRE(g->Y != nullptr, "factk_gemm: null output");
    for (int s = 0; s < g->nsrc; ++s) {
        const factk_src_t& x = g->src[s];
        FACTK_REQUIRE(x.A && x.W && x.K > 0 && x.lda >= x.K && x.ldw >= x.K, "factk_gemm: bad source %d", s);
        FACTK_REQUIRE(x.a_dtype == FACTK_F32 || x.a_dtype == FACTK_BF16, "factk_gemm: bad a_dtype");
        FACTK_REQUIRE(x.w_dtype == FACTK_F32, "factk_gemm: weights must be fp32 (source %d)", s);
    }
    const int tpv = (g->slot + BM - 1) / BM;
    dim3 grid((unsigned)(tpv * g->B), (unsigned)((g->N + BN - 1) / BN));
    gemm_simt_kernel<<<grid, NT, 0, (cudaStream_t)stream>>>(*g, tpv);
    return check_launch("factk_gemm");
}
