// Backward pass of the bidirectional GRU recurrence (nn.GRU(H, H/2, layers, bidirectional=True), blocks.py:401,432) --
// back-propagation through time over the segments of each video, one 4-CTA cluster per (video, direction).
//
// What is sequential is only  dh_{t-1} = z_t * dh_t + W_hh^T dgh_t : the gate pre-activations are NOT recomputed step by
// step -- gi = x W_ih^T + b_ih is kept from the forward, gh = h_{t-1} W_hh^T + b_hh for ALL steps is one GEMM over the
// saved hidden states with a row offset of -1 / +1 (factk_gemm), and everything that is not sequential leaves the kernel as
// plain rows tensors: dgi (the gradient of the input projection's output) and dgh, from which the generic kernels form
// dW_ih, dx (GEMMs), dW_hh = sum_t dgh_t h_{t-1}^T (factk_wgrad with the same row offset) and the bias gradients (column sums).
//
// CTA r of a cluster owns the hidden units [r*U, (r+1)*U), U = Hh/4: it keeps W_hh[:, units]^T (3Hh x U fp32, 192 KB at
// Hh = 256) in shared memory for the whole kernel, computes the gate backward of its units, broadcasts their three dgh values
// to the four CTAs through distributed shared memory (double buffered: one cluster barrier per step), and reduces the
// 3Hh-long dot products of its units.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace factk {

constexpr int GB_THREADS = 256;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void __launch_bounds__(GB_THREADS) gru_bwd_kernel(const float* __restrict__ gi, const float* __restrict__ gh,
                                                             const void* __restrict__ hout, int h_dtype, int ldh,
                                                             const void* __restrict__ dout, int do_dtype, int lddo,
                                                             const float* __restrict__ w_hh_f, const float* __restrict__ w_hh_b, int Hh,
                                                             float* __restrict__ dgi, float* __restrict__ dgh, int slot,
                                                             const int32_t* __restrict__ nseg) {
    extern __shared__ __align__(16) float smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int dir = blockIdx.y, b = blockIdx.z;
    const int U = Hh / 4, G = 3 * Hh, Q = GB_THREADS / U;
    float* Wt = smem;                          // [G][U]
    float* dfull = Wt + (size_t)G * U;         // [2][G]
    float* part = dfull + 2 * G;               // [Q][U]
    float* carry = part + GB_THREADS;          // [U]
    const float* W = dir ? w_hh_b : w_hh_f;
    const int tid = threadIdx.x;
    for (int i = tid; i < G * U; i += GB_THREADS) Wt[i] = W[(size_t)(i / U) * Hh + rank * U + (i % U)];
    if (tid < U) carry[tid] = 0.f;
    const int n = min(nseg[b], slot);
    float* peer[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) peer[r] = cluster.map_shared_rank(dfull, r);
    cluster.sync();

    const int jj = tid, j = rank * U + jj;
    const bool unit = tid < U;
    const size_t gbase = (size_t)b * slot;
    const int goff = dir * G;                  // column offset of this direction in gi / gh / dgi / dgh
    const int hoff = dir * Hh;                 // ... in the hidden-state rows
    // registers of the step being processed (prefetched one step ahead)
    float c_gi[3], c_gh[3], c_hp = 0.f, c_do = 0.f;
    auto fetch = [&](int t) {
        if (!unit) return;
        const size_t row = gbase + t;
#pragma unroll
        for (int g = 0; g < 3; ++g) {
            c_gi[g] = gi[row * (6 * (size_t)Hh) + goff + g * Hh + j];
            c_gh[g] = gh[row * (6 * (size_t)Hh) + goff + g * Hh + j];
        }
        const int tp = dir ? t + 1 : t - 1;
        c_hp = (tp >= 0 && tp < n) ? ld_elem(hout, h_dtype, (gbase + tp) * (size_t)ldh + hoff + j) : 0.f;
        c_do = ld_elem(dout, do_dtype, row * (size_t)lddo + hoff + j);
    };
    if (n > 0) fetch(dir ? 0 : n - 1);
    for (int s = 0; s < n; ++s) {
        const int t = dir ? s : n - 1 - s;
        const int buf = s & 1;
        float dh_dir = 0.f;
        if (unit) {
            const float r = sigmoidf_(c_gi[0] + c_gh[0]), z = sigmoidf_(c_gi[1] + c_gh[1]);
            const float nn = tanhf(c_gi[2] + r * c_gh[2]);
            const float dh = c_do + carry[jj];
            const float dnp = dh * (1.f - z) * (1.f - nn * nn);
            const float dzp = dh * (c_hp - nn) * z * (1.f - z);
            const float drp = dnp * c_gh[2] * r * (1.f - r);
            dh_dir = dh * z;
            const float gg[3] = {drp, dzp, dnp * r};
            const size_t row = (gbase + t) * (6 * (size_t)Hh) + goff;
            dgi[row + j] = drp; dgi[row + Hh + j] = dzp; dgi[row + 2 * Hh + j] = dnp;
#pragma unroll
            for (int g = 0; g < 3; ++g) {
                dgh[row + g * Hh + j] = gg[g];
#pragma unroll
                for (int rr = 0; rr < 4; ++rr) peer[rr][buf * G + g * Hh + j] = gg[g];
            }
        }
        if (s + 1 < n) fetch(dir ? s + 1 : n - 2 - s);       // next step's operands: in flight across the barrier
        cluster.sync();
        // dot products: thread (q, u) sums W_hh[i][unit u] * dgh[i] over its slice of i
        {
            const int u = tid % U, q = tid / U;
            const int per = (G + Q - 1) / Q, i0 = q * per, i1 = min(i0 + per, G);
            const float* dv = dfull + buf * G;
            float a = 0.f;
            for (int i = i0; i < i1; ++i) a = fmaf(Wt[(size_t)i * U + u], dv[i], a);
            part[q * U + u] = a;
        }
        __syncthreads();
        if (unit) {
            float a = dh_dir;
            for (int q = 0; q < Q; ++q) a += part[q * U + jj];
            carry[jj] = a;
        }
        __syncthreads();
    }
    cluster.sync();      // no CTA exits while a peer may still write into its shared memory
}

}  // namespace factk

using namespace factk;

/* BPTT of factk_gru_bidir (one layer).  gi, gh fp32 [B][slot][6*Hh] (gh = h_{t-1} W_hh^T + b_hh for every step, both directions);
 * hout [B][slot][ldh] = the layer's output BEFORE any ReLU (cat[fwd, bwd]); dout its gradient; -> dgi, dgh fp32 [B][slot][6*Hh]
 * (rows < nseg[b] written).  Hh in {32, 64, 128, 256}. */
extern "C" int factk_gru_bwd(const float* gi, const float* gh, const void* hout, int h_dtype, int ldh, const void* dout, int do_dtype,
                             int lddo, const float* w_hh_f, const float* w_hh_b, int Hh, float* dgi, float* dgh, int B, int slot,
                             const int32_t* nseg, void* stream) {
    FACTK_REQUIRE(gi && gh && hout && dout && w_hh_f && w_hh_b && dgi && dgh && nseg && B > 0 && slot > 0, "factk_gru_bwd: bad args");
    FACTK_REQUIRE(Hh == 32 || Hh == 64 || Hh == 128 || Hh == 256, "factk_gru_bwd: Hh = %d unsupported (32/64/128/256)", Hh);
    const int U = Hh / 4, G = 3 * Hh;
    const size_t smem = ((size_t)G * U + 2 * G + GB_THREADS + U) * sizeof(float);
    static unsigned long long devs = 0;
    if (first_use_on_device(devs)) cudaFuncSetAttribute(gru_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(4, 2, B);
    cfg.blockDim = dim3(GB_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 4; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, gru_bwd_kernel, gi, gh, hout, h_dtype, ldh, dout, do_dtype, lddo, w_hh_f, w_hh_b, Hh, dgi, dgh,
                                       slot, nseg);
    if (e != cudaSuccess) {
        set_error("factk_gru_bwd: %s", cudaGetErrorString(e));
        return FACTK_ERR_CUDA;
    }
    return check_launch("factk_gru_bwd");
}
