// Backward-pass primitives of the FACT / FACT_CLIP training step (scripts/train.py:262-268: loss.backward()).
// Everything the reference leaves to torch autograd over eager ops is rebuilt from a handful of kernels on the "rows"
// layout ([B][slot][ld], len[b] valid rows):
//   * wgrad:      dW[n,k] += sum_rows dZ[row,n] * A[row+off,k]      (weight gradient of every Conv1d tap / Linear, and the
//                 rows^T x rows contraction behind "attention^T @ values" in the softmax-over-frames direction)
//   * dgrad is the forward GEMM itself with transposed weights and negated tap offsets (factk_gemm / factk_gemm_tc)
//   * colsum:     bias / LayerNorm-affine gradients
//   * row kernels: ReLU mask, dropout, softmax-splice, row softmax, LayerNorm, L2-normalise backward
//   * column softmax forward (normalised attention kept for the backward) and backward
//   * segment reduce / expand (segment mean backward, gathered pre-activation backward)
// Reductions are two-stage with a fixed summation order: bit-reproducible gradients, no atomics.
#include "common.cuh"

namespace factk {

// rows per partial-sum chunk: short (token-side) tensors get small chunks so that more than a handful of CTAs share the rows
__host__ __device__ inline int wg_rc(int slot) { return slot <= 4096 ? 128 : 1024; }

__host__ __device__ inline int wg_nchunk(int slot) { return (slot + wg_rc(slot) - 1) / wg_rc(slot); }

// ------------------------------------------------------------------------------------------------ wgrad (CUDA cores)
// 64 x 64 tile of dW per CTA, 256 threads, 4 x 4 micro-tile, 16 rows per shared-memory stage.
__global__ void __launch_bounds__(256) wgrad_partial_kernel(const void* __restrict__ dZ, int dz_dtype, int lddz,
                                                            const void* __restrict__ A, int a_dtype, int lda, int a_slot,
                                                            int row_off, const float* __restrict__ pos, int pos_ld, int pos_d,
                                                            const int32_t* __restrict__ pos_idx, int N, int K, float* __restrict__ ws,
                                                            int slot, const int32_t* __restrict__ len, int nchunk, int ktiles) {
    const int tn = blockIdx.x / ktiles, tk = blockIdx.x % ktiles;
    const int chunk = blockIdx.y, b = blockIdx.z;
    const int len_b = len ? min(len[b], slot) : slot;
    const int r0 = chunk * wg_rc(slot);
    if (r0 >= len_b) return;
    const int r1 = min(r0 + wg_rc(slot), len_b);
    __shared__ __align__(16) float Zs[16][68];
    __shared__ __align__(16) float As[16][68];
    const int tid = threadIdx.x, lr = tid >> 4, lc = (tid & 15) * 4, ty = tid >> 4, tx = tid & 15;
    const bool zvec = ((reinterpret_cast<uintptr_t>(dZ) & 15u) == 0) && ((lddz & 3) == 0);
    const bool avec = ((reinterpret_cast<uintptr_t>(A) & 15u) == 0) && ((lda & 3) == 0);
    const int n_ld = tn * 64 + lc, k_ld = tk * 64 + lc;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    float z[4], a[4];
    auto fetch = [&](int rb) {           // rows rb .. rb + 15 of the chunk into registers (one row of 4 + 4 values per thread)
        const int r = rb + lr;
#pragma unroll
        for (int j = 0; j < 4; ++j) z[j] = a[j] = 0.f;
        if (r >= r1) return;
        const size_t zb = ((size_t)b * slot + r) * (size_t)lddz;
        if (zvec && n_ld + 4 <= N) {
            const float4 v = ld_vec4(dZ, dz_dtype, zb + n_ld);
            z[0] = v.x; z[1] = v.y; z[2] = v.z; z[3] = v.w;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (n_ld + j < N) z[j] = ld_elem(dZ, dz_dtype, zb + n_ld + j);
        }
        const int sr = r + row_off;
        if (sr >= 0 && sr < len_b) {
            const size_t ab = ((size_t)b * a_slot + sr) * (size_t)lda;
            if (avec && k_ld + 4 <= K) {
                const float4 v = ld_vec4(A, a_dtype, ab + k_ld);
                a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (k_ld + j < K) a[j] = ld_elem(A, a_dtype, ab + k_ld + j);
            }
            if (pos != nullptr && k_ld < pos_d) {
                const size_t pi = pos_idx ? (size_t)pos_idx[(size_t)b * slot + r] : (size_t)r;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (k_ld + j < pos_d && k_ld + j < K) a[j] += pos[pi * (size_t)pos_ld + k_ld + j];
            }
        }
    };
    fetch(r0);
    for (int rb = r0; rb < r1; rb += 16) {
        *reinterpret_cast<float4*>(&Zs[lr][lc]) = make_float4(z[0], z[1], z[2], z[3]);
        *reinterpret_cast<float4*>(&As[lr][lc]) = make_float4(a[0], a[1], a[2], a[3]);
        __syncthreads();
        if (rb + 16 < r1) fetch(rb + 16);        // the next stage's loads fly under this stage's FMAs
#pragma unroll
        for (int rr = 0; rr < 16; ++rr) {
            const float4 zv = *reinterpret_cast<const float4*>(&Zs[rr][ty * 4]);
            const float4 av = *reinterpret_cast<const float4*>(&As[rr][tx * 4]);
            const float zz[4] = {zv.x, zv.y, zv.z, zv.w}, aa[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(zz[i], aa[j], acc[i][j]);
        }
        __syncthreads();
    }
    float* w = ws + (size_t)(b * nchunk + chunk) * (size_t)N * (size_t)K;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int n = tn * 64 + ty * 4 + i;
        if (n >= N) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = tk * 64 + tx * 4 + j;
            if (k < K) w[(size_t)n * K + k] = acc[i][j];
        }
    }
}

// out[bo][i / K][i % K] = alpha * sum over the valid (video, chunk) partials (+ out): fixed order.  Partials of `psz`
// floats spaced `pstride` apart; per_video: one output per video (bo = b), else the videos are summed too.
__global__ void __launch_bounds__(256) partial_reduce_kernel(const float* __restrict__ ws, size_t pstride, int psz, int K,
                                                             float* __restrict__ out, int ldo, long long out_bstride, int B, int slot,
                                                             const int32_t* __restrict__ len, int nchunk, int rows_per_chunk, float alpha,
                                                             int accumulate, int per_video) {
    // 32 outputs x 8 chunk groups per CTA: group q sums the partials q, q + 8, ... of every video, the eight sums combine in
    // group order -- the order of additions depends on the shape only
    __shared__ float part[8][32];
    const int o = threadIdx.x & 31, q = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + o;
    const int b0 = per_video ? blockIdx.y : 0, b1 = per_video ? blockIdx.y + 1 : B;
    float s = 0.f;
    if (i < psz) {
        for (int b = b0; b < b1; ++b) {
            const int len_b = len ? min(len[b], slot) : slot;
            const int live = min(nchunk, (len_b + rows_per_chunk - 1) / rows_per_chunk);
            const float* w = ws + (size_t)b * nchunk * pstride + i;
#pragma unroll 4
            for (int c = q; c < live; c += 8) s += w[(size_t)c * pstride];
        }
    }
    part[q][o] = s;
    __syncthreads();
    if (q != 0 || i >= psz) return;
#pragma unroll
    for (int j = 1; j < 8; ++j) s += part[j][o];
    float* op = out + (per_video ? (size_t)blockIdx.y * (size_t)out_bstride : 0) + (size_t)(i / K) * ldo + (i % K);
    *op = alpha * s + (accumulate ? *op : 0.f);
}

void launch_partial_reduce(const float* ws, size_t pstride, int psz, int K, float* out, int ldo, long long out_bstride, int B, int slot,
                           const int32_t* len, int nchunk, int rows_per_chunk, float alpha, int accumulate, cudaStream_t st) {
    const int per_video = out_bstride != 0;
    partial_reduce_kernel<<<dim3((psz + 31) / 32, per_video ? B : 1), 256, 0, st>>>(ws, pstride, psz, K, out, ldo, out_bstride, B, slot, len,
                                                                                     nchunk, rows_per_chunk, alpha, accumulate, per_video);
}

// ------------------------------------------------------------------------------------------------ column sums
// ws[(b, chunk)][n] = sum over the chunk's valid rows of X[b,t,n] (* Y[b,t,n] when Y != NULL)
constexpr int CSUM_RC = 128;      // rows per partial of the column sums / LayerNorm affine gradients
__host__ __device__ inline int csum_nchunk(int slot) { return (slot + CSUM_RC - 1) / CSUM_RC; }

// 64 columns x 4 row groups per CTA: thread (c, g) sums rows r0 + g, r0 + g + 4, ...; the four partials combine in a fixed order
__global__ void __launch_bounds__(256) colsum_partial_kernel(const void* __restrict__ X, int x_dtype, int ldx, const void* __restrict__ Y,
                                                             int y_dtype, int ldy, int N, float* __restrict__ ws, int slot,
                                                             const int32_t* __restrict__ len, int nchunk) {
    __shared__ float part[4][64];
    const int c = threadIdx.x & 63, g = threadIdx.x >> 6;
    const int n = blockIdx.x * 64 + c;
    const int chunk = blockIdx.y, b = blockIdx.z;
    const int len_b = len ? min(len[b], slot) : slot;
    const int r0 = chunk * CSUM_RC;
    if (r0 >= len_b) return;
    const int r1 = min(r0 + CSUM_RC, len_b);
    float s = 0.f;
    if (n < N) {
#pragma unroll 4
        for (int r = r0 + g; r < r1; r += 4) {
            float v = ld_elem(X, x_dtype, ((size_t)b * slot + r) * (size_t)ldx + n);
            if (Y) v *= ld_elem(Y, y_dtype, ((size_t)b * slot + r) * (size_t)ldy + n);
            s += v;
        }
    }
    part[g][c] = s;
    __syncthreads();
    if (g == 0 && n < N) ws[(size_t)(b * nchunk + chunk) * N + n] = (part[0][c] + part[1][c]) + (part[2][c] + part[3][c]);
}

// ------------------------------------------------------------------------------------------------ elementwise on rows
enum { EW_RELU_BWD = 0, EW_AXPY = 1, EW_DROPOUT = 2, EW_DROPOUT_CH = 3, EW_COPY = 4, EW_MUL = 5, EW_ADD = 6, EW_RELU = 7, EW_ROWSCALE = 8, EW_ADDTAB = 9 };
constexpr int EW_ROWS = 8;      // rows per CTA

// counter-based uniform in [0,1): a 64-bit mix of (seed, site, element index) -- the same value in the forward and the
// backward pass, nothing stored
__device__ __forceinline__ float hash_uniform(unsigned long long seed, unsigned site, unsigned long long idx) {
    unsigned long long x = seed ^ (0x9E3779B97F4A7C15ull * (unsigned long long)(site + 1)) ^ (idx * 0xBF58476D1CE4E5B9ull);
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27; x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    return (float)(x >> 40) * (1.0f / 16777216.0f);
}

// Y[b,t,n] = op(X[b,t,n], Y or Y2 ...) for t < len[b]:
//   RELU_BWD:   Y = X * (R > 0)              (R = the layer's ReLU output)
//   AXPY:       Y = Y + alpha * X
//   DROPOUT:    Y = X * keep(b,t,n) / (1-p)  (keep from the hash; element index = (b*slot + t) * N + n)
//   DROPOUT_CH: Y = X * keep(b,n) / (1-p)    (nn.Dropout2d on (1,D,T): whole feature channels, blocks.py:614-617)
//   COPY:       Y = alpha * X
//   MUL:        Y = X * R
//   ADD:        Y = X + R
//   RELU:       Y = max(X, 0)
//   ROWSCALE:   Y = X * R[row]               (R: one value per row, e.g. the time mask of basic.time_mask)
//   ADDTAB:     Y = X + R[ridx ? ridx[b,t] : t][n]   (R: one table shared by the videos: add_positional_encoding with the
//                                                     sinusoid table, rows picked by frame index or by segment centre)
// seed_ptr (optional, device): added to ``seed`` -- lets a captured CUDA graph draw new masks on every replay
// x_slot: row slots per video of X (0 broadcasts one [slot][ldx] table over the videos: positional / query tables)
__global__ void rows_elementwise_kernel(int op, const void* X, int x_dtype, int ldx, const void* R,
                                        int r_dtype, int ldr, void* Y, int y_dtype, int ldy, int N, int slot,
                                        const int32_t* __restrict__ len, float alpha, float p, unsigned long long seed,
                                        unsigned site, int x_slot, const unsigned long long* __restrict__ seed_ptr,
                                        const int32_t* __restrict__ ridx) {
    const int b = blockIdx.z;
    const int len_b = len ? min(len[b], slot) : slot;
    if (seed_ptr) seed += *seed_ptr;
    const float keep_scale = (op == EW_DROPOUT || op == EW_DROPOUT_CH) ? 1.f / (1.f - p) : 0.f;
    // 4 consecutive elements per thread when every operand allows 8 / 16-byte accesses (the common case: channel counts and
    // leading dimensions are multiples of 4); the scalar loop below serves the rest
    const bool v4 = (N & 3) == 0 && (ldx & 3) == 0 && (ldy & 3) == 0 && (!R || (ldr & 3) == 0 || op == EW_ROWSCALE) &&
                    (reinterpret_cast<uintptr_t>(X) & 15u) == 0 && (reinterpret_cast<uintptr_t>(Y) & 15u) == 0 &&
                    (!R || (reinterpret_cast<uintptr_t>(R) & 15u) == 0);
    for (int t = blockIdx.y * EW_ROWS; t < min((int)(blockIdx.y + 1) * EW_ROWS, len_b); ++t) {
    const size_t row = (size_t)b * slot + t, xrow = (size_t)b * x_slot + t;
    if (v4) {
        const size_t rrow = (op == EW_ADDTAB) ? (size_t)(ridx ? ridx[row] : t) : row;
        for (int n = (blockIdx.x * blockDim.x + threadIdx.x) * 4; n < N; n += gridDim.x * blockDim.x * 4) {
            const float4 xv = ld_vec4(X, x_dtype, xrow * ldx + n);
            float x[4] = {xv.x, xv.y, xv.z, xv.w}, r[4] = {0.f, 0.f, 0.f, 0.f}, y0[4] = {0.f, 0.f, 0.f, 0.f}, y[4];
            if (op == EW_RELU_BWD || op == EW_MUL || op == EW_ADD || op == EW_ADDTAB) {
                const float4 rv = ld_vec4(R, r_dtype, rrow * ldr + n);
                r[0] = rv.x; r[1] = rv.y; r[2] = rv.z; r[3] = rv.w;
            } else if (op == EW_ROWSCALE) {
                r[0] = r[1] = r[2] = r[3] = ld_elem(R, r_dtype, row * ldr);
            }
            if (op == EW_AXPY) {
                const float4 yv = ld_vec4(Y, y_dtype, row * ldy + n);
                y0[0] = yv.x; y0[1] = yv.y; y0[2] = yv.z; y0[3] = yv.w;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                switch (op) {
                    case EW_RELU_BWD: y[j] = r[j] > 0.f ? x[j] : 0.f; break;
                    case EW_AXPY: y[j] = y0[j] + alpha * x[j]; break;
                    case EW_DROPOUT: y[j] = hash_uniform(seed, site, row * (size_t)N + n + j) >= p ? x[j] * keep_scale : 0.f; break;
                    case EW_DROPOUT_CH: y[j] = hash_uniform(seed, site, (size_t)b * N + n + j) >= p ? x[j] * keep_scale : 0.f; break;
                    case EW_MUL: case EW_ROWSCALE: y[j] = x[j] * r[j]; break;
                    case EW_ADD: case EW_ADDTAB: y[j] = x[j] + r[j]; break;
                    case EW_RELU: y[j] = fmaxf(x[j], 0.f); break;
                    default: y[j] = alpha * x[j]; break;
                }
            }
            st_vec4(Y, y_dtype, row * ldy + n, make_float4(y[0], y[1], y[2], y[3]));
        }
        continue;
    }
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
        const float x = ld_elem(X, x_dtype, xrow * ldx + n);
        float y;
        switch (op) {
            case EW_RELU_BWD: y = ld_elem(R, r_dtype, row * ldr + n) > 0.f ? x : 0.f; break;
            case EW_AXPY: y = ld_elem(Y, y_dtype, row * ldy + n) + alpha * x; break;
            case EW_DROPOUT: y = hash_uniform(seed, site, row * (size_t)N + n) >= p ? x * keep_scale : 0.f; break;
            case EW_DROPOUT_CH: y = hash_uniform(seed, site, (size_t)b * N + n) >= p ? x * keep_scale : 0.f; break;
            case EW_MUL: y = x * ld_elem(R, r_dtype, row * ldr + n); break;
            case EW_ADD: y = x + ld_elem(R, r_dtype, row * ldr + n); break;
            case EW_RELU: y = fmaxf(x, 0.f); break;
            case EW_ROWSCALE: y = x * ld_elem(R, r_dtype, row * ldr); break;
            case EW_ADDTAB: y = x + ld_elem(R, r_dtype, (size_t)(ridx ? ridx[row] : t) * ldr + n); break;
            default: y = alpha * x; break;
        }
        st_elem(Y, y_dtype, row * ldy + n, y);
    }
    }
}

// dst[b][c][r] = src[b][r][c] (fp32; 32 x 32 tiles through shared memory)
__global__ void transpose_kernel(const float* __restrict__ src, int lds, long long src_bstride, float* __restrict__ dst, int ldd,
                                 long long dst_bstride, int R, int Ccols) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z, r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    const float* s = src + (size_t)b * src_bstride;
    float* d = dst + (size_t)b * dst_bstride;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < R && c < Ccols) ? s[(size_t)r * lds + c] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        if (r < R && c < Ccols) d[(size_t)c * ldd + r] = tile[threadIdx.x][i];
    }
}

// ------------------------------------------------------------------------------------------------ row kernels (one warp per row)
// process_feature backward (blocks.py:195-202): Y rows = [feat (H-C), softmax(logits) (C)], dY their gradient, dCl the
// gradient of the raw logits returned beside it -> dX = [dY feat, p * (dYp - sum p dYp) + dCl]
__global__ void splice_bwd_kernel(const void* __restrict__ Y, int y_dtype, int ldy, const void* __restrict__ dY, int dy_dtype, int lddy,
                                  const float* __restrict__ dCl, int lddc, void* __restrict__ dX, int dx_dtype, int lddx, int H, int C,
                                  int slot, const int32_t* __restrict__ len) {
    const int warps = blockDim.x >> 5, lane = threadIdx.x & 31;
    const int t = blockIdx.x * warps + (threadIdx.x >> 5), b = blockIdx.y;
    const int len_b = len ? min(len[b], slot) : slot;
    if (t >= len_b) return;
    const size_t row = (size_t)b * slot + t;
    const int F = H - C;
    for (int c = lane; c < F; c += 32) st_elem(dX, dx_dtype, row * lddx + c, dY ? ld_elem(dY, dy_dtype, row * lddy + c) : 0.f);
    float dot = 0.f;
    if (dY)
        for (int c = lane; c < C; c += 32) dot += ld_elem(Y, y_dtype, row * ldy + F + c) * ld_elem(dY, dy_dtype, row * lddy + F + c);
    dot = warp_sum(dot);
    for (int c = lane; c < C; c += 32) {
        float g = dCl ? dCl[row * lddc + c] : 0.f;
        if (dY) g += ld_elem(Y, y_dtype, row * ldy + F + c) * (ld_elem(dY, dy_dtype, row * lddy + F + c) - dot);
        st_elem(dX, dx_dtype, row * lddx + F + c, g);
    }
}

// softmax over the first M columns of a row: dL = P * (dP - sum P dP) (+ dL when accumulate)
__global__ void row_softmax_bwd_kernel(const float* __restrict__ P, int ldp, const float* __restrict__ dP, int lddp, float* __restrict__ dL,
                                       int lddl, int M, int slot, const int32_t* __restrict__ len, int accumulate) {
    const int warps = blockDim.x >> 5, lane = threadIdx.x & 31;
    const int t = blockIdx.x * warps + (threadIdx.x >> 5), b = blockIdx.y;
    const int len_b = len ? min(len[b], slot) : slot;
    if (t >= len_b) return;
    const size_t row = (size_t)b * slot + t;
    float dot = 0.f;
    for (int c = lane; c < M; c += 32) dot += P[row * ldp + c] * dP[row * lddp + c];
    dot = warp_sum(dot);
    for (int c = lane; c < M; c += 32) {
        const float g = P[row * ldp + c] * (dP[row * lddp + c] - dot);
        dL[row * lddl + c] = g + (accumulate ? dL[row * lddl + c] : 0.f);
    }
}

constexpr int LN_RC = 32;          // rows per CTA of the LayerNorm backward (8 warps x 4 rows)
__host__ __device__ inline int ln_nchunk(int slot) { return (slot + LN_RC - 1) / LN_RC; }

// LayerNorm backward: y = relu?((v - mu) * rstd * w + b), v = x (+ r).  dV (the gradient of v) is written (or
// accumulated); per-chunk partial sums of dgamma = dy * xhat and dbeta = dy go to ws[(b,chunk)][2][E].
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const void* __restrict__ X, int x_dtype, int ldx, const void* __restrict__ R,
                                                            int r_dtype, int ldr, const float* __restrict__ w, const float* __restrict__ bias,
                                                            float eps, int relu, const void* __restrict__ dY, int dy_dtype, int lddy,
                                                            void* __restrict__ dV, int dv_dtype, int lddv, int accumulate,
                                                            float* __restrict__ ws, int E, int slot, const int32_t* __restrict__ len, int nchunk) {
    extern __shared__ float sm[];          // [8 warps][2][E]
    const int chunk = blockIdx.x, b = blockIdx.y;
    const int len_b = len ? min(len[b], slot) : slot;
    const int r0 = chunk * LN_RC;
    if (r0 >= len_b) return;
    const int r1 = min(r0 + LN_RC, len_b);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    float* mg = sm + (size_t)warp * 2 * E;
    float* mb = mg + E;
    for (int e = lane; e < E; e += 32) { mg[e] = 0.f; mb[e] = 0.f; }
    for (int t = r0 + warp; t < r1; t += nw) {
        const size_t row = (size_t)b * slot + t;
        float s = 0.f, ss = 0.f;
        for (int e = lane; e < E; e += 32) {
            float v = ld_elem(X, x_dtype, row * ldx + e);
            if (R) v += ld_elem(R, r_dtype, row * ldr + e);
            s += v;
        }
        const float mu = warp_sum(s) / E;
        for (int e = lane; e < E; e += 32) {
            float v = ld_elem(X, x_dtype, row * ldx + e);
            if (R) v += ld_elem(R, r_dtype, row * ldr + e);
            ss += (v - mu) * (v - mu);
        }
        const float rstd = rsqrtf(warp_sum(ss) / E + eps);
        float a = 0.f, c = 0.f;        // sum g, sum g * xhat with g = dy * w
        for (int e = lane; e < E; e += 32) {
            float v = ld_elem(X, x_dtype, row * ldx + e);
            if (R) v += ld_elem(R, r_dtype, row * ldr + e);
            const float xh = (v - mu) * rstd;
            float dy = ld_elem(dY, dy_dtype, row * lddy + e);
            if (relu && xh * w[e] + bias[e] <= 0.f) dy = 0.f;
            mg[e] += dy * xh;
            mb[e] += dy;
            const float g = dy * w[e];
            a += g;
            c += g * xh;
        }
        a = warp_sum(a) / E;
        c = warp_sum(c) / E;
        for (int e = lane; e < E; e += 32) {
            float v = ld_elem(X, x_dtype, row * ldx + e);
            if (R) v += ld_elem(R, r_dtype, row * ldr + e);
            const float xh = (v - mu) * rstd;
            float dy = ld_elem(dY, dy_dtype, row * lddy + e);
            if (relu && xh * w[e] + bias[e] <= 0.f) dy = 0.f;
            float g = rstd * (dy * w[e] - a - xh * c);
            if (accumulate) g += ld_elem(dV, dv_dtype, row * lddv + e);
            st_elem(dV, dv_dtype, row * lddv + e, g);
        }
    }
    __syncthreads();
    float* o = ws + (size_t)(b * nchunk + chunk) * 2 * E;
    for (int i = threadIdx.x; i < 2 * E; i += blockDim.x) {
        float s = 0.f;
        for (int wv = 0; wv < nw; ++wv) s += sm[(size_t)wv * 2 * E + i];
        o[i] = s;
    }
}

// F.normalize backward (blocks.py:174): y = x / max(|x|, eps) -> dx = (dy - y (y . dy)) / max(|x|, eps)
__global__ void l2norm_bwd_kernel(const void* __restrict__ X, int x_dtype, int ldx, const void* __restrict__ dY, int dy_dtype, int lddy,
                                  void* __restrict__ dX, int dx_dtype, int lddx, int E, float eps, int slot, const int32_t* __restrict__ len) {
    const int warps = blockDim.x >> 5, lane = threadIdx.x & 31;
    const int t = blockIdx.x * warps + (threadIdx.x >> 5), b = blockIdx.y;
    const int len_b = len ? min(len[b], slot) : slot;
    if (t >= len_b) return;
    const size_t row = (size_t)b * slot + t;
    float ss = 0.f, dot = 0.f;
    for (int e = lane; e < E; e += 32) {
        const float x = ld_elem(X, x_dtype, row * ldx + e);
        ss += x * x;
        dot += x * ld_elem(dY, dy_dtype, row * lddy + e);
    }
    const float nrm = sqrtf(warp_sum(ss));
    const float d = fmaxf(nrm, eps);
    dot = warp_sum(dot);
    // below the clamp the map is x / eps: a plain scaling
    const float k = nrm > eps ? dot / (d * d * d) : 0.f;
    for (int e = lane; e < E; e += 32) {
        const float x = ld_elem(X, x_dtype, row * ldx + e);
        st_elem(dX, dx_dtype, row * lddx + e, ld_elem(dY, dy_dtype, row * lddy + e) / d - x * k);
    }
}

// ------------------------------------------------------------------------------------------------ column softmax (softmax over rows)
constexpr int CS_ROWS = 256;      // rows per partial chunk of the column statistics
__host__ __device__ inline int cs_nchunk(int slot) { return (slot + CS_ROWS - 1) / CS_ROWS; }

// ws[(b, chunk)][m] = (max, sum exp(x - max)) over the chunk's rows.  Thread per column, coalesced over m.
__global__ void colsm_stats_kernel(const void* __restrict__ L, int l_dtype, int ldl, float* __restrict__ ws, int M, int slot,
                                   const int32_t* __restrict__ len, int nchunk, float scale) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x, chunk = blockIdx.y, b = blockIdx.z;
    const int len_b = len ? min(len[b], slot) : slot;
    const int r0 = chunk * CS_ROWS;
    if (m >= M || r0 >= len_b) return;
    const int r1 = min(r0 + CS_ROWS, len_b);
    float mx = -INFINITY, s = 0.f;
#pragma unroll 4
    for (int r = r0; r < r1; ++r) {
        const float x = scale * ld_elem(L, l_dtype, ((size_t)b * slot + r) * ldl + m);
        if (x > mx) { s = s * __expf(mx - x) + 1.f; mx = x; }
        else s += __expf(x - mx);
    }
    float* o = ws + ((size_t)(b * nchunk + chunk) * M + m) * 2;
    o[0] = mx; o[1] = s;
}

// stats[b][m] = (max, 1 / sum) over all valid rows, combined in chunk order
__global__ void colsm_combine_kernel(const float* __restrict__ ws, float* __restrict__ stats, int M, int slot,
                                     const int32_t* __restrict__ len, int nchunk) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (m >= M) return;
    const int len_b = len ? min(len[b], slot) : slot;
    float mx = -INFINITY, s = 0.f;
    for (int c = 0; c < nchunk && c * CS_ROWS < len_b; ++c) {
        const float* p = ws + ((size_t)(b * nchunk + c) * M + m) * 2;
        const float nm = fmaxf(mx, p[0]);
        s = s * __expf(mx - nm) + p[1] * __expf(p[0] - nm);
        mx = nm;
    }
    stats[((size_t)b * M + m) * 2] = mx;
    stats[((size_t)b * M + m) * 2 + 1] = 1.f / s;
}

// Four adjacent columns per thread, CS_TR rows per CTA (vector loads / stores when the rows allow it).
constexpr int CS_TR = 8;
__global__ void __launch_bounds__(128) colsm_normalize_kernel(const void* __restrict__ L, int l_dtype, int ldl, const float* __restrict__ stats,
                                                              void* __restrict__ P, int p_dtype, int ldp, int M, int slot,
                                                              const int32_t* __restrict__ len, float scale, int vec) {
    const int b = blockIdx.z, m = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int len_b = len ? min(len[b], slot) : slot;
    if (m >= M) return;
    float mx[4], is[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float* st = stats + ((size_t)b * M + min(m + j, M - 1)) * 2;
        mx[j] = st[0]; is[j] = st[1];
    }
    const int t1 = min((int)(blockIdx.y + 1) * CS_TR, len_b);
    for (int t = blockIdx.y * CS_TR; t < t1; ++t) {
        const size_t row = (size_t)b * slot + t;
        if (vec && m + 4 <= M) {
            const float4 x = ld_vec4(L, l_dtype, row * ldl + m);
            st_vec4(P, p_dtype, row * ldp + m, make_float4(__expf(scale * x.x - mx[0]) * is[0], __expf(scale * x.y - mx[1]) * is[1],
                                                           __expf(scale * x.z - mx[2]) * is[2], __expf(scale * x.w - mx[3]) * is[3]));
        } else {
            for (int j = 0; j < 4 && m + j < M; ++j)
                st_elem(P, p_dtype, row * ldp + m + j, __expf(scale * ld_elem(L, l_dtype, row * ldl + m + j) - mx[j]) * is[j]);
        }
    }
}

// dL[t,m] (+)= scale * P[t,m] * (dP[t,m] - c[m]),  c[m] = sum_t P[t,m] dP[t,m]  (cvec from the colsum of P * dP)
__global__ void __launch_bounds__(128) colsm_bwd_kernel(const void* __restrict__ P, int p_dtype, int ldp, const void* __restrict__ dP, int dp_dtype,
                                                        int lddp, const float* __restrict__ cvec, void* __restrict__ dL, int dl_dtype, int lddl,
                                                        int M, int slot, const int32_t* __restrict__ len, float scale, int accumulate, int vec) {
    const int b = blockIdx.z, m = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int len_b = len ? min(len[b], slot) : slot;
    if (m >= M) return;
    float c[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) c[j] = cvec[(size_t)b * M + min(m + j, M - 1)];
    const int t1 = min((int)(blockIdx.y + 1) * CS_TR, len_b);
    for (int t = blockIdx.y * CS_TR; t < t1; ++t) {
        const size_t row = (size_t)b * slot + t;
        if (vec && m + 4 <= M) {
            const float4 p = ld_vec4(P, p_dtype, row * ldp + m), d = ld_vec4(dP, dp_dtype, row * lddp + m);
            float4 g = make_float4(scale * p.x * (d.x - c[0]), scale * p.y * (d.y - c[1]), scale * p.z * (d.z - c[2]), scale * p.w * (d.w - c[3]));
            if (accumulate) {
                const float4 o = ld_vec4(dL, dl_dtype, row * lddl + m);
                g.x += o.x; g.y += o.y; g.z += o.z; g.w += o.w;
            }
            st_vec4(dL, dl_dtype, row * lddl + m, g);
        } else {
            for (int j = 0; j < 4 && m + j < M; ++j) {
                float g = scale * ld_elem(P, p_dtype, row * ldp + m + j) * (ld_elem(dP, dp_dtype, row * lddp + m + j) - c[j]);
                if (accumulate) g += ld_elem(dL, dl_dtype, row * lddl + m + j);
                st_elem(dL, dl_dtype, row * lddl + m + j, g);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ segments
// out[b][s][:E] (+)= (mean ? 1/len : 1) * sum over the segment's frames of X[b][t][:E]   (segments are contiguous runs)
__global__ void segment_reduce_kernel(const void* __restrict__ X, int x_dtype, int ldx, void* __restrict__ out, int o_dtype, int ldo,
                                      const int32_t* __restrict__ seg_start, const int32_t* __restrict__ seg_len,
                                      const int32_t* __restrict__ nseg, int slot, int E, int mean, int accumulate) {
    const int s = blockIdx.x, b = blockIdx.y;
    if (s >= nseg[b]) return;
    const int t0 = seg_start[(size_t)b * slot + s], n = seg_len[(size_t)b * slot + s];
    const float sc = mean ? 1.f / (float)n : 1.f;
    for (int e = threadIdx.x; e < E; e += blockDim.x) {
        float a = 0.f;
        for (int t = t0; t < t0 + n; ++t) a += ld_elem(X, x_dtype, ((size_t)b * slot + t) * ldx + e);
        const size_t o = ((size_t)b * slot + s) * ldo + e;
        st_elem(out, o_dtype, o, a * sc + (accumulate ? ld_elem(out, o_dtype, o) : 0.f));
    }
}

// out[b][t][:E] (+)= (inv_len ? 1/len[label] : 1) * seg[b][label[b][t]][:E]
__global__ void segment_expand_kernel(const void* __restrict__ seg, int s_dtype, int lds, const int32_t* __restrict__ seg_label,
                                      const int32_t* __restrict__ seg_len, void* __restrict__ out, int o_dtype, int ldo, int slot,
                                      const int32_t* __restrict__ len, int E, int inv_len, int accumulate) {
    const int b = blockIdx.z, t = blockIdx.y;
    const int len_b = len ? min(len[b], slot) : slot;
    if (t >= len_b) return;
    const int s = seg_label[(size_t)b * slot + t];
    const float sc = inv_len ? 1.f / (float)seg_len[(size_t)b * slot + s] : 1.f;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x) {
        const size_t o = ((size_t)b * slot + t) * ldo + e;
        const float v = sc * ld_elem(seg, s_dtype, ((size_t)b * slot + s) * lds + e);
        st_elem(out, o_dtype, o, v + (accumulate ? ld_elem(out, o_dtype, o) : 0.f));
    }
}

}  // namespace factk

using namespace factk;

extern "C" size_t factk_wgrad_ws_floats(int B, int slot, int N, int K) {
    return (size_t)B * wg_nchunk(slot) * (size_t)N * (size_t)K;
}

extern "C" int factk_wgrad(const void* dZ, int dz_dtype, int lddz, const void* A, int a_dtype, int lda, int a_slot, int row_off,
                           const float* pos, int pos_ld, int pos_d, const int32_t* pos_idx, int N, int K, float* dW, int lddw,
                           long long dw_bstride, float alpha, int accumulate, int B, int slot, const int32_t* len, float* ws,
                           void* stream) {
    FACTK_REQUIRE(dZ && A && dW && ws && N > 0 && K > 0 && B > 0 && slot > 0, "factk_wgrad: bad args");
    FACTK_REQUIRE(lddz >= N && lda >= K && lddw >= K, "factk_wgrad: leading dimensions too small");
    const int nchunk = wg_nchunk(slot), ntn = (N + 63) / 64, ntk = (K + 63) / 64;
    cudaStream_t st = (cudaStream_t)stream;
    wgrad_partial_kernel<<<dim3(ntn * ntk, nchunk, B), 256, 0, st>>>(dZ, dz_dtype, lddz, A, a_dtype, lda, a_slot, row_off, pos, pos_ld,
                                                                     pos_d, pos_idx, N, K, ws, slot, len, nchunk, ntk);
    const int psz = N * K, per_video = dw_bstride != 0;
    partial_reduce_kernel<<<dim3((psz + 31) / 32, per_video ? B : 1), 256, 0, st>>>(ws, (size_t)psz, psz, K, dW, lddw, dw_bstride, B, slot,
                                                                                     len, nchunk, wg_rc(slot), alpha, accumulate, per_video);
    return check_launch("factk_wgrad");
}

extern "C" size_t factk_colsum_ws_floats(int B, int slot, int N) { return (size_t)B * csum_nchunk(slot) * (size_t)N; }

/* out[(b)][n] (+)= alpha * sum over valid rows of X[b,t,n] (* Y[b,t,n]) */
extern "C" int factk_colsum(const void* X, int x_dtype, int ldx, const void* Y, int y_dtype, int ldy, int N, float* out,
                            long long out_bstride, float alpha, int accumulate, int B, int slot, const int32_t* len, float* ws,
                            void* stream) {
    FACTK_REQUIRE(X && out && ws && N > 0 && B > 0 && slot > 0, "factk_colsum: bad args");
    const int nchunk = csum_nchunk(slot), per_video = out_bstride != 0;
    cudaStream_t st = (cudaStream_t)stream;
    colsum_partial_kernel<<<dim3((N + 63) / 64, nchunk, B), 256, 0, st>>>(X, x_dtype, ldx, Y, y_dtype, ldy, N, ws, slot, len, nchunk);
    partial_reduce_kernel<<<dim3((N + 31) / 32, per_video ? B : 1), 256, 0, st>>>(ws, (size_t)N, N, N, out, N, out_bstride, B, slot, len,
                                                                                 nchunk, CSUM_RC, alpha, accumulate, per_video);
    return check_launch("factk_colsum");
}

extern "C" int factk_rows_elementwise(int op, const void* X, int x_dtype, int ldx, const void* R, int r_dtype, int ldr, void* Y,
                                      int y_dtype, int ldy, int N, int B, int slot, const int32_t* len, float alpha, float p,
                                      unsigned long long seed, unsigned site, int x_slot, const unsigned long long* seed_ptr,
                                      const int32_t* ridx, void* stream) {
    FACTK_REQUIRE(X && Y && N > 0 && B > 0 && slot > 0 && op >= 0 && op <= EW_ADDTAB, "factk_rows_elementwise: bad args");
    FACTK_REQUIRE(!(op == EW_RELU_BWD || op == EW_MUL || op == EW_ADD || op == EW_ROWSCALE || op == EW_ADDTAB) || R,
                  "factk_rows_elementwise: op %d needs R", op);
    FACTK_REQUIRE(p >= 0.f && p < 1.f, "factk_rows_elementwise: p = %f", p);
    const int threads = N >= 1024 ? 256 : (N >= 512 ? 128 : 64);
    const int gx = (N + threads * 4 - 1) / (threads * 4);
    rows_elementwise_kernel<<<dim3(gx > 0 ? gx : 1, (slot + EW_ROWS - 1) / EW_ROWS, B), threads, 0, (cudaStream_t)stream>>>(
        op, X, x_dtype, ldx, R, r_dtype, ldr, Y, y_dtype, ldy, N, slot, len, alpha, p, seed, site, x_slot < 0 ? slot : x_slot, seed_ptr, ridx);
    return check_launch("factk_rows_elementwise");
}

extern "C" int factk_transpose(const float* src, int lds, long long src_bstride, float* dst, int ldd, long long dst_bstride, int R,
                               int Ccols, int B, void* stream) {
    FACTK_REQUIRE(src && dst && R > 0 && Ccols > 0 && B > 0, "factk_transpose: bad args");
    transpose_kernel<<<dim3((Ccols + 31) / 32, (R + 31) / 32, B), dim3(32, 8), 0, (cudaStream_t)stream>>>(src, lds, src_bstride, dst, ldd,
                                                                                                      dst_bstride, R, Ccols);
    return check_launch("factk_transpose");
}

extern "C" int factk_splice_bwd(const void* Y, int y_dtype, int ldy, const void* dY, int dy_dtype, int lddy, const float* dCl,
                                int lddc, void* dX, int dx_dtype, int lddx, int H, int C, int B, int slot, const int32_t* len,
                                void* stream) {
    FACTK_REQUIRE(Y && dX && H > 0 && C > 0 && C <= H && (dY || dCl), "factk_splice_bwd: bad args");
    splice_bwd_kernel<<<dim3((slot + 7) / 8, B), 256, 0, (cudaStream_t)stream>>>(Y, y_dtype, ldy, dY, dy_dtype, lddy, dCl, lddc, dX, dx_dtype,
                                                                               lddx, H, C, slot, len);
    return check_launch("factk_splice_bwd");
}

extern "C" int factk_row_softmax_bwd(const float* P, int ldp, const float* dP, int lddp, float* dL, int lddl, int M, int accumulate,
                                     int B, int slot, const int32_t* len, void* stream) {
    FACTK_REQUIRE(P && dP && dL && M > 0, "factk_row_softmax_bwd: bad args");
    row_softmax_bwd_kernel<<<dim3((slot + 7) / 8, B), 256, 0, (cudaStream_t)stream>>>(P, ldp, dP, lddp, dL, lddl, M, slot, len, accumulate);
    return check_launch("factk_row_softmax_bwd");
}

extern "C" size_t factk_layernorm_bwd_ws_floats(int B, int slot, int E) { return (size_t)B * ln_nchunk(slot) * 2 * (size_t)E; }

extern "C" int factk_layernorm_bwd(const void* X, int x_dtype, int ldx, const void* R, int r_dtype, int ldr, const float* w,
                                   const float* b, float eps, int relu, const void* dY, int dy_dtype, int lddy, void* dV, int dv_dtype,
                                   int lddv, int accumulate, float* dw, float* db, int B, int slot, const int32_t* len, int E,
                                   float* ws, void* stream) {
    FACTK_REQUIRE(X && w && b && dY && dV && dw && db && ws && E > 0, "factk_layernorm_bwd: bad args");
    const int nchunk = ln_nchunk(slot);
    const size_t smem = (size_t)8 * 2 * E * sizeof(float);
    FACTK_REQUIRE(smem <= 200 * 1024, "factk_layernorm_bwd: E = %d too wide", E);
    cudaStream_t st = (cudaStream_t)stream;
    static unsigned long long devs = 0;
    if (first_use_on_device(devs)) cudaFuncSetAttribute(layernorm_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    layernorm_bwd_kernel<<<dim3(nchunk, B), 256, smem, st>>>(X, x_dtype, ldx, R, r_dtype, ldr, w, b, eps, relu, dY, dy_dtype, lddy, dV,
                                                           dv_dtype, lddv, accumulate, ws, E, slot, len, nchunk);
    partial_reduce_kernel<<<dim3((E + 31) / 32, 1), 256, 0, st>>>(ws, (size_t)2 * E, E, E, dw, E, 0, B, slot, len, nchunk, LN_RC, 1.f, 1, 0);
    partial_reduce_kernel<<<dim3((E + 31) / 32, 1), 256, 0, st>>>(ws + E, (size_t)2 * E, E, E, db, E, 0, B, slot, len, nchunk, LN_RC, 1.f, 1, 0);
    return check_launch("factk_layernorm_bwd");
}

extern "C" int factk_l2norm_bwd(const void* X, int x_dtype, int ldx, const void* dY, int dy_dtype, int lddy, void* dX, int dx_dtype,
                                int lddx, int B, int slot, const int32_t* len, int E, float eps, void* stream) {
    FACTK_REQUIRE(X && dY && dX && E > 0, "factk_l2norm_bwd: bad args");
    l2norm_bwd_kernel<<<dim3((slot + 7) / 8, B), 256, 0, (cudaStream_t)stream>>>(X, x_dtype, ldx, dY, dy_dtype, lddy, dX, dx_dtype, lddx, E,
                                                                               eps, slot, len);
    return check_launch("factk_l2norm_bwd");
}

extern "C" size_t factk_col_softmax_train_ws_floats(int B, int slot, int M) {
    return (size_t)B * cs_nchunk(slot) * M * 2 + (size_t)B * M * 2;
}

static bool rows_vec4(const void* p, int dtype, int ld) {
    return (reinterpret_cast<uintptr_t>(p) & (dtype == FACTK_BF16 ? 7u : 15u)) == 0 && (ld & 3) == 0;
}

/* P[b,t,m] = softmax over the valid rows t of scale * L[b,t,m]  (normalised attention kept for the backward pass); L, P fp32 or bf16 */
extern "C" int factk_col_softmax(const void* L, int l_dtype, int ldl, void* P, int p_dtype, int ldp, int M, float scale, int B, int slot,
                                 const int32_t* len, float* ws, void* stream) {
    FACTK_REQUIRE(L && P && ws && M > 0, "factk_col_softmax: bad args");
    const int nchunk = cs_nchunk(slot);
    float* stats = ws + (size_t)B * nchunk * M * 2;
    cudaStream_t st = (cudaStream_t)stream;
    colsm_stats_kernel<<<dim3((M + 63) / 64, nchunk, B), 64, 0, st>>>(L, l_dtype, ldl, ws, M, slot, len, nchunk, scale);
    colsm_combine_kernel<<<dim3((M + 63) / 64, B), 64, 0, st>>>(ws, stats, M, slot, len, nchunk);
    const int vec = rows_vec4(L, l_dtype, ldl) && rows_vec4(P, p_dtype, ldp);
    colsm_normalize_kernel<<<dim3((M + 511) / 512, (slot + CS_TR - 1) / CS_TR, B), 128, 0, st>>>(L, l_dtype, ldl, stats, P, p_dtype, ldp, M, slot,
                                                                                                len, scale, vec);
    return check_launch("factk_col_softmax");
}

/* dL (+)= scale * P * (dP - colsum(P * dP)); P, dP, dL fp32 or bf16; ws: factk_colsum_ws_floats(B, slot, M) + B*M floats */
extern "C" int factk_col_softmax_bwd(const void* P, int p_dtype, int ldp, const void* dP, int dp_dtype, int lddp, void* dL, int dl_dtype, int lddl,
                                     int M, float scale, int accumulate, int B, int slot, const int32_t* len, float* ws, void* stream) {
    FACTK_REQUIRE(P && dP && dL && ws && M > 0, "factk_col_softmax_bwd: bad args");
    float* cvec = ws + factk_colsum_ws_floats(B, slot, M);
    int rc = factk_colsum(P, p_dtype, ldp, dP, dp_dtype, lddp, M, cvec, M, 1.f, 0, B, slot, len, ws, stream);
    if (rc) return rc;
    const int vec = rows_vec4(P, p_dtype, ldp) && rows_vec4(dP, dp_dtype, lddp) && rows_vec4(dL, dl_dtype, lddl);
    colsm_bwd_kernel<<<dim3((M + 511) / 512, (slot + CS_TR - 1) / CS_TR, B), 128, 0, (cudaStream_t)stream>>>(
        P, p_dtype, ldp, dP, dp_dtype, lddp, cvec, dL, dl_dtype, lddl, M, slot, len, scale, accumulate, vec);
    return check_launch("factk_col_softmax_bwd");
}

extern "C" int factk_segment_reduce(const void* X, int x_dtype, int ldx, void* out, int o_dtype, int ldo, const int32_t* seg_start,
                                    const int32_t* seg_len, const int32_t* nseg, int B, int slot, int E, int mean, int accumulate,
                                    void* stream) {
    FACTK_REQUIRE(X && out && seg_start && seg_len && nseg && E > 0, "factk_segment_reduce: bad args");
    const int threads = E >= 256 ? 256 : 128;
    segment_reduce_kernel<<<dim3(slot, B), threads, 0, (cudaStream_t)stream>>>(X, x_dtype, ldx, out, o_dtype, ldo, seg_start, seg_len, nseg,
                                                                             slot, E, mean, accumulate);
    return check_launch("factk_segment_reduce");
}

extern "C" int factk_segment_expand(const void* seg, int s_dtype, int lds, const int32_t* seg_label, const int32_t* seg_len, void* out,
                                    int o_dtype, int ldo, int B, int slot, const int32_t* len, int E, int inv_len, int accumulate,
                                    void* stream) {
    FACTK_REQUIRE(seg && seg_label && out && E > 0 && (!inv_len || seg_len), "factk_segment_expand: bad args");
    const int threads = E >= 256 ? 256 : 128;
    segment_expand_kernel<<<dim3((E + threads - 1) / threads, slot, B), threads, 0, (cudaStream_t)stream>>>(
        seg, s_dtype, lds, seg_label, seg_len, out, o_dtype, ldo, slot, len, E, inv_len, accumulate);
    return check_launch("factk_segment_expand");
}
