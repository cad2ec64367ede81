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

// ---------------------------------------------------------------------------------------------------------------------------
// Hh = 256 (every shipped configuration): W_hh^T in REGISTERS.  The shared-memory version above streams its 192 KB weight slice
// through the LSU every step (>= 1536 cycles of shared-memory bandwidth, measured 4.2 us per step with the cluster barrier);
// here thread (unit u, slice q) of 512 keeps the 96 weights W_hh[96 q .. 96 q + 95][unit] in registers for the whole kernel,
// the received dgh vector is read as broadcast 128-bit loads, and the exchange is plain 16-byte distributed-shared-memory
// stores with the arrival flag IN the data (the protocol of the tensor-core forward kernel, gru_mma.cu): a word that has not
// arrived holds a NaN pattern no arithmetic produces; the 64 threads of a K slice poll two words each, meet at a named
// barrier, read the slice, and re-mark their words after the CTA barrier that ends the step.  (st.async completing a
// transaction mbarrier cost ~1000 cycles per step here, the flagged stores cost one remote-store latency: 1.9 -> ~1 us.)
constexpr int GB2_THREADS = 512, GB2_Q = 8, GB2_KS = 96;      // 768 = 8 x 96
__device__ long long* gb2_dbg = nullptr;     // factk_gru_bwd_debug: clock64 of (cluster 0, rank 0, thread 0) at six points of the first 64 steps
#define GB2_MARK(i) do { if (dbgp != nullptr && s < 64) dbgp[s * 8 + (i)] = clock64(); } while (0)
constexpr uint32_t GB2_EMPTY = 0xFFFFFFFFu;                     // "not arrived yet" marker of an exchanged word

__device__ __forceinline__ void gb_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
    uint32_t ok = 0, spins = 0;
    while (true) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (ok) break;
        if (++spins > (1u << 24)) __trap();
    }
}

__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(GB2_THREADS, 1)
gru_bwd256_kernel(const float* __restrict__ gi, const float* __restrict__ gh, const void* __restrict__ hout, int h_dtype, int ldh,
                  const void* __restrict__ dout, int do_dtype, int lddo, const float* __restrict__ w_hh_f,
                  const float* __restrict__ w_hh_b, float* __restrict__ dgi, float* __restrict__ dgh, int slot,
                  const int32_t* __restrict__ nseg) {
    constexpr int Hh = 256, U = 64, G = 768;
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int dir = blockIdx.y, b = blockIdx.z;
    __shared__ __align__(16) float dfull[2][G];
    __shared__ __align__(16) float mine[3 * U];
    __shared__ float part[GB2_Q][U];
    __shared__ float carry[U];
    long long* const dbgp = (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0) ? gb2_dbg : nullptr;
    const float* W = dir ? w_hh_b : w_hh_f;
    const int tid = threadIdx.x, lane = tid & 31;
    const int u = tid & (U - 1), q = tid >> 6;
    float w[GB2_KS];
#pragma unroll
    for (int i = 0; i < GB2_KS; ++i) w[i] = W[(size_t)(q * GB2_KS + i) * Hh + rank * U + u];
    if (tid < U) carry[tid] = 0.f;
    for (int i = tid; i < 2 * G; i += GB2_THREADS) reinterpret_cast<uint32_t*>(&dfull[0][0])[i] = GB2_EMPTY;
    const int n = min(nseg[b], slot);
    const bool unit = tid < U;
    const int j = rank * U + u;
    const size_t gbase = (size_t)b * slot;
    const int goff = dir * G, hoff = dir * Hh;
    float c_gi[3] = {0.f, 0.f, 0.f}, c_gh[3] = {0.f, 0.f, 0.f}, c_hp = 0.f, c_do = 0.f;
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
    const uint32_t dfull_u32 = (uint32_t)__cvta_generic_to_shared(&dfull[0][0]);
    // polling role: thread p < 48 of slice q owns words 2p, 2p + 1 of the slice's 96
    const int pl = tid & 63;
    const uint32_t poll_u32 = dfull_u32 + (uint32_t)(q * GB2_KS + 2 * pl) * 4u;
    __syncthreads();
    cluster.sync();

    for (int s = 0; s < n; ++s) {
        const int t = dir ? s : n - 1 - s;
        const int buf = s & 1;
        float dh_dir = 0.f;
        GB2_MARK(0);
        if (unit) {
            const float r = sigmoidf_(c_gi[0] + c_gh[0]), z = sigmoidf_(c_gi[1] + c_gh[1]);
            const float nn = tanhf(c_gi[2] + r * c_gh[2]);
            const float dh = c_do + carry[u];
            const float dnp = dh * (1.f - z) * (1.f - nn * nn);
            const float dzp = dh * (c_hp - nn) * z * (1.f - z);
            const float drp = dnp * c_gh[2] * r * (1.f - r);
            dh_dir = dh * z;
            const size_t row = (gbase + t) * (6 * (size_t)Hh) + goff;
            dgi[row + j] = drp; dgi[row + Hh + j] = dzp; dgi[row + 2 * Hh + j] = dnp;
            dgh[row + j] = drp; dgh[row + Hh + j] = dzp; dgh[row + 2 * Hh + j] = dnp * r;
            mine[u] = drp; mine[U + u] = dzp; mine[2 * U + u] = dnp * r;
        }
        __syncthreads();
        GB2_MARK(1);
        if (tid < 192) {                    // 48 x 16 bytes to each of the four CTAs (this one included)
            const int dest = tid / 48, c = tid % 48, g = c >> 4, u4 = (c & 15) * 4;
            const float4 v = *reinterpret_cast<const float4*>(&mine[g * U + u4]);
            const uint32_t o = dfull_u32 + (uint32_t)(buf * G + g * Hh + rank * U + u4) * 4u;
            uint32_t ra;
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(o), "r"(dest));
            uint32_t x[4] = {__float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w)};
#pragma unroll
            for (int i = 0; i < 4; ++i) x[i] = (x[i] == GB2_EMPTY) ? 0x7FC00000u : x[i];      // a NaN stays a NaN, never "not arrived"
            asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ra), "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]) : "memory");
        }
        GB2_MARK(2);
        if (s + 1 < n) fetch(dir ? s + 1 : n - 2 - s);       // next step's operands: in flight across the exchange
        GB2_MARK(3);
        if (pl < GB2_KS / 2) {              // wait for this slice's words (two per polling thread)
            uint32_t a0, a1, spins = 0;
            const uint32_t pa = poll_u32 + (uint32_t)(buf * G) * 4u;
            while (true) {
                asm volatile("ld.volatile.shared.v2.b32 {%0, %1}, [%2];" : "=r"(a0), "=r"(a1) : "r"(pa) : "memory");
                if (a0 != GB2_EMPTY && a1 != GB2_EMPTY) break;
                if (++spins > (1u << 24)) __trap();     // protocol bug: fail loudly instead of hanging the GPU
            }
        }
        asm volatile("bar.sync %0, 64;" ::"r"(q + 1) : "memory");           // the two warps of slice q
        GB2_MARK(4);
        {
            const float* dv = &dfull[buf][q * GB2_KS];
            float a4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int i = 0; i < GB2_KS; i += 4) {
                const float4 d4 = *reinterpret_cast<const float4*>(dv + i);
                a4[0] = fmaf(w[i], d4.x, a4[0]); a4[1] = fmaf(w[i + 1], d4.y, a4[1]);
                a4[2] = fmaf(w[i + 2], d4.z, a4[2]); a4[3] = fmaf(w[i + 3], d4.w, a4[3]);
            }
            part[q][u] = (a4[0] + a4[1]) + (a4[2] + a4[3]);
        }
        GB2_MARK(5);
        __syncthreads();
        GB2_MARK(6);
        if (pl < GB2_KS / 2) {              // everyone of this CTA has read dfull[buf]: mark it "not arrived" for step s + 2
            const uint32_t pa = poll_u32 + (uint32_t)(buf * G) * 4u;
            asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(pa), "r"(GB2_EMPTY), "r"(GB2_EMPTY) : "memory");
        }
        if (unit) {
            float a = dh_dir;
#pragma unroll
            for (int qq = 0; qq < GB2_Q; ++qq) a += part[qq][u];
            carry[u] = a;
        }
    }
    cluster.sync();      // no CTA exits while a peer may still write into its shared memory
}

}  // namespace factk

using namespace factk;

/* Debug aid: clock64 at seven points of the first 64 steps of the Hh = 256 kernel (cluster 0, rank 0) into buf[64][8] (int64, device); NULL: off. */
extern "C" int factk_gru_bwd_debug(void* buf) {
    long long* p = reinterpret_cast<long long*>(buf);
    cudaMemcpyToSymbol(factk::gb2_dbg, &p, sizeof(p));
    return factk::check_launch("factk_gru_bwd_debug");
}

/* BPTT of factk_gru_bidir (one layer).  gi, gh fp32 [B][slot][6*Hh] (gh = h_{t-1} W_hh^T + b_hh for every step, both directions);
 * hout [B][slot][ldh] = the layer's output BEFORE any ReLU (cat[fwd, bwd]); dout its gradient; -> dgi, dgh fp32 [B][slot][6*Hh]
 * (rows < nseg[b] written).  Hh in {32, 64, 128, 256}. */
extern "C" int factk_gru_bwd(const float* gi, const float* gh, const void* hout, int h_dtype, int ldh, const void* dout, int do_dtype,
                             int lddo, const float* w_hh_f, const float* w_hh_b, int Hh, float* dgi, float* dgh, int B, int slot,
                             const int32_t* nseg, void* stream) {
    FACTK_REQUIRE(gi && gh && hout && dout && w_hh_f && w_hh_b && dgi && dgh && nseg && B > 0 && slot > 0, "factk_gru_bwd: bad args");
    FACTK_REQUIRE(Hh == 32 || Hh == 64 || Hh == 128 || Hh == 256, "factk_gru_bwd: Hh = %d unsupported (32/64/128/256)", Hh);
    if (Hh == 256) {
        gru_bwd256_kernel<<<dim3(4, 2, B), GB2_THREADS, 0, (cudaStream_t)stream>>>(gi, gh, hout, h_dtype, ldh, dout, do_dtype, lddo, w_hh_f,
                                                                                 w_hh_b, dgi, dgh, slot, nseg);
        return check_launch("factk_gru_bwd");
    }
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
