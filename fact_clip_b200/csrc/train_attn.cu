// Head-batched products of the training step's attention cores (one launch for all heads of all videos):
//   C[b][h][m][n] (+)= alpha * sum_k A[b][h](m, k) * Bm[b][h](n, k)
// with every operand a column band of a rows tensor (head h starts at column h * hstride) and each operand either
// k-contiguous (element (m, k) at row m, column k) or k-major (element (m, k) at row k, column m).  The six products of
// nn.MultiheadAttention's forward / backward per attention (logits, apply, dP, dV, dQ, dK; models/basic.py:437,500,507-514)
// are three operand layouts of this one kernel; the reductions over the frames of a video (k = frame) split k across CTAs
// and combine the partials in a fixed order (bit-reproducible, CUDA-graph safe).
// fp32 accumulate on the CUDA cores: per head the inner dimension is 32 (logits) or the output is 32 wide (apply), the
// [frames x heads x tokens] probability tensor is the HBM traffic that bounds these launches.
#include "common.cuh"

namespace factk {

constexpr int HB_BK = 16, HB_NT = 256;

struct HeadsOperand {
    const void* p;
    int dtype, ld, kmajor;
    size_t base;      // element offset of (video, head)
    int lim;          // valid extent of the non-k index
};

// Fetch the 16-wide k chunk of a [ROWS x 16] operand tile into registers: EPT = ROWS / 16 elements per thread.
template <int ROWS>
__device__ __forceinline__ void hb_fetch(const HeadsOperand& o, int r0, int k0, int k_end, float (&r)[ROWS / 16]) {
    constexpr int EPT = ROWS / 16, G = HB_NT / ROWS;
    const int tid = threadIdx.x;
#pragma unroll
    for (int j = 0; j < EPT; ++j) r[j] = 0.f;
    if (o.kmajor) {
        const int m = r0 + tid % ROWS, kq = tid / ROWS;
        if (m >= o.lim) return;
#pragma unroll
        for (int j = 0; j < EPT; ++j) {
            const int k = k0 + kq + j * G;
            if (k < k_end) r[j] = ld_elem(o.p, o.dtype, o.base + (size_t)k * o.ld + m);
        }
    } else {
        const int m = r0 + tid % ROWS, k = k0 + (tid / ROWS) * EPT;
        if (m >= o.lim) return;
        const size_t e = o.base + (size_t)m * o.ld + k;
        const uintptr_t addr = reinterpret_cast<uintptr_t>(o.p) + e * (o.dtype == FACTK_BF16 ? 2 : 4);
        if (EPT >= 4 && k + EPT <= k_end && (addr & (o.dtype == FACTK_BF16 ? 7u : 15u)) == 0) {
#pragma unroll
            for (int j = 0; j < EPT / 4; ++j) {
                const float4 v = ld_vec4(o.p, o.dtype, e + 4 * j);
                r[4 * j] = v.x; r[4 * j + 1] = v.y; r[4 * j + 2] = v.z; r[4 * j + 3] = v.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < EPT; ++j)
                if (k + j < k_end) r[j] = ld_elem(o.p, o.dtype, e + j);
        }
    }
}

template <int ROWS>
__device__ __forceinline__ void hb_stash(int kmajor, float (*S)[ROWS + 4], const float (&r)[ROWS / 16]) {
    constexpr int EPT = ROWS / 16, G = HB_NT / ROWS;
    const int tid = threadIdx.x, m = tid % ROWS, q = tid / ROWS;
#pragma unroll
    for (int j = 0; j < EPT; ++j) {
        if (kmajor) S[q + j * G][m] = r[j];
        else S[q * EPT + j][m] = r[j];
    }
}

// BM x BN tile of C per CTA, TM x TN micro-tile per thread ((BM / TM) * (BN / TN) == 256; TM, TN in {4, 8}).
template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(HB_NT, 2) heads_mm_kernel(const __grid_constant__ factk_heads_mm_t g, int ntn, int ksplit, int kchunk) {
    static_assert((BM / TM) * (BN / TN) == HB_NT && (TM == 4 || TM == 8) && (TN == 4 || TN == 8), "thread layout");
    __shared__ __align__(16) float As[2][HB_BK][BM + 4];
    __shared__ __align__(16) float Bs[2][HB_BK][BN + 4];

    const int b = blockIdx.z / g.nhead, h = blockIdx.z % g.nhead;
    const int m0 = (blockIdx.x / ntn) * BM, n0 = (blockIdx.x % ntn) * BN;
    const int split = blockIdx.y;
    const int len_b = g.len ? g.len[b] : 0x7fffffff;
    const int Mb = g.len_mode == 1 ? min(len_b, g.M) : g.M;
    const int Kb = g.len_mode == 2 ? min(len_b, g.K) : g.K;
    const int k_begin = split * kchunk, k_end = min(Kb, k_begin + kchunk);
    if (m0 >= Mb || (ksplit > 1 && k_begin >= k_end)) return;

    HeadsOperand oa{g.A, g.a_dtype, g.lda, g.a_kmajor, (size_t)b * (size_t)g.a_bstride + (size_t)h * g.a_hstride, Mb};
    HeadsOperand ob{g.Bm, g.b_dtype, g.ldb, g.b_kmajor, (size_t)b * (size_t)g.b_bstride + (size_t)h * g.b_hstride, g.N};

    const int tid = threadIdx.x;
    constexpr int TX = BN / TN;
    const int tx = tid % TX, ty = tid / TX;

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    float ra[BM / 16], rb[BN / 16];
    const int nchunk = (k_end - k_begin + HB_BK - 1) / HB_BK;
    if (nchunk > 0) {
        hb_fetch<BM>(oa, m0, k_begin, k_end, ra);
        hb_fetch<BN>(ob, n0, k_begin, k_end, rb);
        hb_stash<BM>(g.a_kmajor, As[0], ra);
        hb_stash<BN>(g.b_kmajor, Bs[0], rb);
    }
    __syncthreads();
    for (int it = 0; it < nchunk; ++it) {
        const int buf = it & 1;
        const bool more = it + 1 < nchunk;
        if (more) {
            hb_fetch<BM>(oa, m0, k_begin + (it + 1) * HB_BK, k_end, ra);
            hb_fetch<BN>(ob, n0, k_begin + (it + 1) * HB_BK, k_end, rb);
        }
#pragma unroll
        for (int k = 0; k < HB_BK; ++k) {
            float av[TM], bv[TN];
            {
                const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
                av[0] = a0.x; av[1] = a0.y; av[2] = a0.z; av[3] = a0.w;
                if (TM == 8) {
                    const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][BM / 2 + ty * 4]);
                    av[TM - 4] = a1.x; av[TM - 3] = a1.y; av[TM - 2] = a1.z; av[TM - 1] = a1.w;
                }
                const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
                bv[0] = b0.x; bv[1] = b0.y; bv[2] = b0.z; bv[3] = b0.w;
                if (TN == 8) {
                    const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][BN / 2 + tx * 4]);
                    bv[TN - 4] = b1.x; bv[TN - 3] = b1.y; bv[TN - 2] = b1.z; bv[TN - 1] = b1.w;
                }
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (more) {
            hb_stash<BM>(g.a_kmajor, As[buf ^ 1], ra);
            hb_stash<BN>(g.b_kmajor, Bs[buf ^ 1], rb);
        }
        __syncthreads();
    }

    if (ksplit > 1) {       // raw partial; heads_mm_reduce applies alpha / accumulate
        float* w = g.ws + ((size_t)blockIdx.z * ksplit + split) * (size_t)g.M * (size_t)g.N;
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            const int r = m0 + (i < 4 ? ty * 4 + i : BM / 2 + ty * 4 + (i - 4));
            if (r >= Mb) continue;
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                const int c = n0 + (j < 4 ? tx * 4 + j : BN / 2 + tx * 4 + (j - 4));
                if (c < g.N) w[(size_t)r * g.N + c] = acc[i][j];
            }
        }
        return;
    }
    const size_t cbase = (size_t)b * (size_t)g.c_bstride + (size_t)h * g.c_hstride;
    const int esz = g.c_dtype == FACTK_BF16 ? 2 : 4;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int r = m0 + (i < 4 ? ty * 4 + i : BM / 2 + ty * 4 + (i - 4));
        if (r >= Mb) continue;
#pragma unroll
        for (int q = 0; q < TN / 4; ++q) {
            const int c = n0 + (q == 0 ? tx * 4 : BN / 2 + tx * 4);
            if (c >= g.N) continue;
            const size_t e = cbase + (size_t)r * g.ldc + c;
            const bool vec = c + 4 <= g.N && ((reinterpret_cast<uintptr_t>(g.C) + e * esz) & (esz * 4 - 1)) == 0;
            if (vec) {
                float4 v = make_float4(acc[i][q * 4] * g.alpha, acc[i][q * 4 + 1] * g.alpha, acc[i][q * 4 + 2] * g.alpha, acc[i][q * 4 + 3] * g.alpha);
                if (g.accumulate) {
                    const float4 o = ld_vec4(g.C, g.c_dtype, e);
                    v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
                }
                st_vec4(g.C, g.c_dtype, e, v);
            } else {
                for (int j = 0; j < 4 && c + j < g.N; ++j) {
                    float v = acc[i][q * 4 + j] * g.alpha;
                    if (g.accumulate) v += ld_elem(g.C, g.c_dtype, e + j);
                    st_elem(g.C, g.c_dtype, e + j, v);
                }
            }
        }
    }
}

// C = alpha * (sum of the live k-split partials, in split order) (+ C)
__global__ void __launch_bounds__(256) heads_mm_reduce_kernel(const __grid_constant__ factk_heads_mm_t g, int ksplit, int kchunk) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.M * g.N) return;
    const int b = blockIdx.y / g.nhead, h = blockIdx.y % g.nhead;
    const int len_b = g.len ? g.len[b] : 0x7fffffff;
    const int Kb = g.len_mode == 2 ? min(len_b, g.K) : g.K;
    const int r = i / g.N, c = i % g.N;
    if (g.len_mode == 1 && r >= min(len_b, g.M)) return;
    const float* w = g.ws + (size_t)blockIdx.y * ksplit * (size_t)g.M * (size_t)g.N + i;
    float s = 0.f;
    for (int sp = 0; sp < ksplit && sp * kchunk < Kb; ++sp) s += w[(size_t)sp * g.M * g.N];
    const size_t e = (size_t)b * (size_t)g.c_bstride + (size_t)h * g.c_hstride + (size_t)r * g.ldc + c;
    float v = g.alpha * s;
    if (g.accumulate) v += ld_elem(g.C, g.c_dtype, e);
    st_elem(g.C, g.c_dtype, e, v);
}

struct HeadsPlan {
    int bm, bn, ntm, ntn, ksplit, kchunk;
};

static HeadsPlan heads_plan(int batch, int nhead, int M, int N, int K) {
    HeadsPlan p;
    if (N <= 32) p.bn = 32;
    else p.bn = ((N + 63) / 64) * 64 < ((N + 127) / 128) * 128 ? 64 : 128;
    p.bm = (p.bn == 32 && M >= 2048) ? 256 : 128;
    p.ntm = (M + p.bm - 1) / p.bm;
    p.ntn = (N + p.bn - 1) / p.bn;
    p.ksplit = 1;
    p.kchunk = ((K + HB_BK - 1) / HB_BK) * HB_BK;
    const long ctas = (long)p.ntm * p.ntn * batch * nhead;
    if (K >= 2048 && ctas < 4 * 148) {          // a reduction over the frames with a token-sized output: split k
        int want = (int)((4 * 148 + ctas - 1) / ctas);
        int chunk = (K + want - 1) / want;
        chunk = ((chunk + 255) / 256) * 256;    // >= 256 rows per partial, a multiple of the k step
        p.kchunk = chunk;
        p.ksplit = (K + chunk - 1) / chunk;
    }
    return p;
}

}  // namespace factk

extern "C" size_t factk_heads_mm_ws_floats(int batch, int nhead, int M, int N, int K) {
    using namespace factk;
    const HeadsPlan p = heads_plan(batch, nhead, M, N, K);
    return p.ksplit > 1 ? (size_t)batch * nhead * p.ksplit * (size_t)M * (size_t)N : 0;
}

extern "C" int factk_heads_mm(const factk_heads_mm_t* g, void* stream) {
    using namespace factk;
    FACTK_REQUIRE(g && g->A && g->Bm && g->C, "factk_heads_mm: null operand");
    FACTK_REQUIRE(g->batch > 0 && g->nhead > 0 && g->M > 0 && g->N > 0 && g->K > 0, "factk_heads_mm: bad shape");
    FACTK_REQUIRE(g->len_mode >= 0 && g->len_mode <= 2 && (g->len_mode == 0 || g->len), "factk_heads_mm: len_mode %d without len", g->len_mode);
    FACTK_REQUIRE((g->a_dtype == FACTK_F32 || g->a_dtype == FACTK_BF16) && (g->b_dtype == FACTK_F32 || g->b_dtype == FACTK_BF16) &&
                      (g->c_dtype == FACTK_F32 || g->c_dtype == FACTK_BF16), "factk_heads_mm: bad dtype");
    const HeadsPlan p = heads_plan(g->batch, g->nhead, g->M, g->N, g->K);
    FACTK_REQUIRE(p.ksplit == 1 || g->ws, "factk_heads_mm: this shape needs the k-split workspace (factk_heads_mm_ws_floats)");
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 grid((unsigned)(p.ntm * p.ntn), (unsigned)p.ksplit, (unsigned)(g->batch * g->nhead));
    if (p.bn == 32 && p.bm == 256) heads_mm_kernel<256, 32, 8, 4><<<grid, HB_NT, 0, st>>>(*g, p.ntn, p.ksplit, p.kchunk);
    else if (p.bn == 32) heads_mm_kernel<128, 32, 4, 4><<<grid, HB_NT, 0, st>>>(*g, p.ntn, p.ksplit, p.kchunk);
    else if (p.bn == 64) heads_mm_kernel<128, 64, 8, 4><<<grid, HB_NT, 0, st>>>(*g, p.ntn, p.ksplit, p.kchunk);
    else heads_mm_kernel<128, 128, 8, 8><<<grid, HB_NT, 0, st>>>(*g, p.ntn, p.ksplit, p.kchunk);
    if (p.ksplit > 1)
        heads_mm_reduce_kernel<<<dim3((unsigned)((g->M * g->N + 255) / 256), (unsigned)(g->batch * g->nhead)), 256, 0, st>>>(*g, p.ksplit, p.kchunk);
    return check_launch("factk_heads_mm");
}
