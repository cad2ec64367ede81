// Temporal down/up-sampling entirely on device (no host round trip):
//  * tdu_segment  : run-length segmentation of the per-frame argmax (utils/utils.py:25-48,
//                   models/basic.py:597-607, models/blocks.py:454) -- block scan per video.
//  * segment_mean : deterministic mean over each contiguous run (models/basic.py:615-625).
//  * gru_bidir    : the bidirectional GRU recurrence over segments (models/blocks.py:401,432) in fp32 (the fp32 compute
//                   mode and hidden sizes other than 256): a persistent kernel on 4-CTA clusters, each CTA keeps the
//                   W_hh rows of a quarter of the hidden units in registers and the new hidden state is exchanged through
//                   distributed shared memory once per step.  The bf16 mode uses gru_mma.cu.
#include <cooperative_groups.h>

#include <cstdlib>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace factk {

__global__ void __launch_bounds__(1024) tdu_segment_kernel(const int32_t* __restrict__ pred, int slot,
                                                           const int32_t* __restrict__ len, int32_t* seg_label,
                                                           int32_t* seg_start, int32_t* seg_len, int32_t* seg_center,
                                                           int32_t* nseg) {
    __shared__ int warp_tot[32];
    __shared__ int carry_s;
    const int b = blockIdx.x;
    const int T = len ? min(len[b], slot) : slot;
    const size_t base = (size_t)b * slot;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int t0 = 0; t0 < T; t0 += 1024) {
        const int t = t0 + tid;
        int flag = 0;
        if (t < T) flag = (t == 0) || (pred[base + t] != pred[base + t - 1]);
        int v = flag;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += n;
        }
        if (lane == 31) warp_tot[w] = v;
        __syncthreads();
        if (w == 0) {
            int x = warp_tot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int n = __shfl_up_sync(0xffffffffu, x, o);
                if (lane >= o) x += n;
            }
            warp_tot[lane] = x;   // inclusive totals
        }
        __syncthreads();
        const int carry = carry_s;
        const int incl = carry + v + (w > 0 ? warp_tot[w - 1] : 0);
        if (t < T) {
            seg_label[base + t] = incl - 1;
            if (flag) seg_start[base + incl - 1] = t;
        }
        __syncthreads();
        if (tid == 1023) carry_s = incl;
        __syncthreads();
    }
    const int S = carry_s;
    if (tid == 0) nseg[b] = S;
    for (int s = tid; s < S; s += 1024) {
        const int st = seg_start[base + s];
        const int en = (s + 1 < S) ? seg_start[base + s + 1] : T;   // exclusive
        seg_len[base + s] = en - st;
        seg_center[base + s] = (st + en - 1) / 2;
    }
}

// Frame-parallel: one CTA owns the segments that START inside its chunk of SEGM_CHUNK frames (found through seg_label, two
// dependent loads), so consecutive CTAs stream consecutive frames and every frame is read exactly once; a segment that
// extends past the chunk is finished by the CTA that started it.  1024 threads = 8 row lanes x 128 channel threads: short
// segments (<= 4 rows) are summed by lane 0 alone in row order; longer ones (random-init predictions produce videos that
// are ONE 4096-frame segment) are split over the eight lanes with eight row loads in flight each and combined in fixed
// lane order -- deterministic, no atomics.
constexpr int SEGM_CHUNK = 32;
constexpr int SEGM_LANES = 8;
__global__ void __launch_bounds__(128 * SEGM_LANES) segment_mean_kernel(const void* __restrict__ X, int x_dtype, int ldx, void* seg,
                                                                        int s_dtype, int lds, const int32_t* __restrict__ seg_label,
                                                                        const int32_t* __restrict__ seg_start,
                                                                        const int32_t* __restrict__ seg_len,
                                                                        const int32_t* __restrict__ nseg, int slot, int E) {
    __shared__ float4 red[SEGM_LANES][128];
    const int b = blockIdx.y;
    const int S = min(nseg[b], slot);
    if (S <= 0) return;
    const size_t base = (size_t)b * slot;
    const int T = seg_start[base + S - 1] + seg_len[base + S - 1];
    const int f0 = blockIdx.x * SEGM_CHUNK, f1 = f0 + SEGM_CHUNK;
    if (f0 >= T) return;
    const int l0 = seg_label[base + f0];
    const int s_first = l0 + (seg_start[base + l0] != f0 ? 1 : 0);
    int s_end = S;
    if (f1 < T) {
        const int l1 = seg_label[base + f1];
        s_end = l1 + (seg_start[base + l1] != f1 ? 1 : 0);
    }
    const int cl = threadIdx.x & 127, rl = threadIdx.x >> 7;
    const bool vec = ((reinterpret_cast<uintptr_t>(X) & 15u) == 0) && ((ldx & 3) == 0) && ((E & 3) == 0) &&
                     ((reinterpret_cast<uintptr_t>(seg) & 15u) == 0) && ((lds & 3) == 0);
    // (start, length) of the chunk's segments -> shared memory once: no dependent global loads inside the loops
    __shared__ int s_st[SEGM_CHUNK], s_n[SEGM_CHUNK];
    const int nsc = s_end - s_first;            // <= SEGM_CHUNK: every segment starts on a distinct frame of the chunk
    if ((int)threadIdx.x < nsc) {
        s_st[threadIdx.x] = seg_start[base + s_first + threadIdx.x];
        s_n[threadIdx.x] = seg_len[base + s_first + threadIdx.x];
    }
    __syncthreads();
    if (!vec) {
        for (int i = rl; i < nsc; i += SEGM_LANES) {
            const int st = s_st[i], n = s_n[i];
            for (int c = cl; c < E; c += 128) {
                float a = 0.f;
                for (int r = 0; r < n; ++r) a += ld_elem(X, x_dtype, (base + st + r) * (size_t)ldx + c);
                st_elem(seg, s_dtype, (base + s_first + i) * (size_t)lds + c, a / (float)n);
            }
        }
        return;
    }
    // short segments (the common case): one row lane per segment, eight segments in flight per CTA, rows summed in order
    for (int i = rl; i < nsc; i += SEGM_LANES) {
        const int st = s_st[i], n = s_n[i];
        if (n > 4) continue;
        const float inv = 1.f / (float)n;
        for (int c = cl * 4; c < E; c += 512) {
            float4 v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                v[j] = (j < n) ? ld_vec4(X, x_dtype, (base + st + j) * (size_t)ldx + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            float4 a = v[0];
#pragma unroll
            for (int j = 1; j < 4; ++j)
                if (j < n) { a.x += v[j].x; a.y += v[j].y; a.z += v[j].z; a.w += v[j].w; }
            st_vec4(seg, s_dtype, (base + s_first + i) * (size_t)lds + c, make_float4(a.x * inv, a.y * inv, a.z * inv, a.w * inv));
        }
    }
    // long segments: all row lanes cooperate on one segment (uniform control flow: the barriers are reached by every thread)
    for (int i = 0; i < nsc; ++i) {
        const int st = s_st[i], n = s_n[i];
        if (n <= 4) continue;
        const float inv = 1.f / (float)n;
        for (int c0 = 0; c0 < E; c0 += 512) {
            const int c = c0 + cl * 4;
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c < E)
                for (int r = rl; r < n; r += 8 * SEGM_LANES) {
                    float4 v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        v[j] = (r + j * SEGM_LANES < n) ? ld_vec4(X, x_dtype, (base + st + r + j * SEGM_LANES) * (size_t)ldx + c)
                                                        : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int j = 0; j < 8; ++j) { a.x += v[j].x; a.y += v[j].y; a.z += v[j].z; a.w += v[j].w; }
                }
            __syncthreads();
            red[rl][cl] = a;
            __syncthreads();
            if (rl == 0 && c < E) {
                float4 t = red[0][cl];
#pragma unroll
                for (int j = 1; j < SEGM_LANES; ++j) { t.x += red[j][cl].x; t.y += red[j][cl].y; t.z += red[j][cl].z; t.w += red[j][cl].w; }
                st_vec4(seg, s_dtype, (base + s_first + i) * (size_t)lds + c, make_float4(t.x * inv, t.y * inv, t.z * inv, t.w * inv));
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Streaming segment mean in two passes (rows with 16-byte aligned row starts).
//  pass 1, one CTA per 32-frame chunk: the 32 rows are fetched with fully independent coalesced loads into shared memory
//          while the chunk's labels are read; the chunk is cut into "pieces" at the label changes, each piece is summed in
//          row order by one of eight row lanes.  A piece that is a whole segment is finished here; a piece whose segment
//          started in an earlier chunk goes to head[b][chunk], one whose segment continues past the chunk to tail[b][chunk].
//  pass 2, one CTA per chunk: if a segment starting in this chunk continues past it, total = tail[chunk] + head[chunk+1]
//          + ... in chunk order.
// Every frame is read exactly once, nothing but a short index chain depends on the segmentation, and the summation order
// is fixed (deterministic, no atomics -- SURVEY D5): the time no longer depends on how pathological the segmentation is
// (S ~ 0.6 T in one block, S = 1 in the next under random-init weights).
constexpr int SM2_CHUNK = 32, SM2_LANES = 2;      // 256-thread CTAs: several resident per SM (occupancy hides the index chain)
__global__ void __launch_bounds__(128 * SM2_LANES) segment_mean_pass1_kernel(const void* __restrict__ X, int x_dtype, int ldx,
                                                                              void* seg, int s_dtype, int lds,
                                                                              const int32_t* __restrict__ seg_label,
                                                                              const int32_t* __restrict__ seg_start,
                                                                              const int32_t* __restrict__ seg_len,
                                                                              const int32_t* __restrict__ nseg, int slot, int E,
                                                                              float* __restrict__ head, float* __restrict__ tail) {
    extern __shared__ __align__(16) uint8_t sm2_rows[];          // [32 rows][E] in the rows' dtype
    __shared__ int lab[SM2_CHUNK + 2];                           // labels of frames f0-1 .. f0+32 (-1 outside the video)
    __shared__ int p_begin[SM2_CHUNK + 1], n_piece_s;
    const int b = blockIdx.y, c = blockIdx.x, nchunks = gridDim.x;
    const int S = min(nseg[b], slot);
    if (S <= 0) return;
    const size_t base = (size_t)b * slot;
    const int f0 = c * SM2_CHUNK;
    const int es = x_dtype == FACTK_BF16 ? 2 : 4;
    const int row_bytes = E * es, chunks16 = row_bytes / 16;
    // rows -> shared memory: independent loads, issued before anything that depends on the segmentation
    for (int i = threadIdx.x; i < SM2_CHUNK * chunks16; i += blockDim.x) {
        const int r = i / chunks16, k = i % chunks16;
        if (f0 + r < slot)
            *reinterpret_cast<uint4*>(sm2_rows + (size_t)r * row_bytes + k * 16) =
                *reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(X) + ((base + f0 + r) * (size_t)ldx) * es + k * 16);
    }
    const int T = seg_start[base + S - 1] + seg_len[base + S - 1];   // frames of this video (labels exist for t < T)
    if (f0 >= T) return;                                             // uniform
    if (threadIdx.x < SM2_CHUNK + 2) {
        const int t = f0 - 1 + (int)threadIdx.x;
        lab[threadIdx.x] = (t >= 0 && t < T) ? seg_label[base + t] : -1;
    }
    __syncthreads();
    const int nrows = min(SM2_CHUNK, T - f0);
    if (threadIdx.x < 32) {                                      // piece boundaries: one ballot instead of a serial scan
        const int r = threadIdx.x;
        const bool flag = r < nrows && (r == 0 || lab[r + 1] != lab[r]);
        const unsigned mask = __ballot_sync(0xffffffffu, flag);
        if (flag) p_begin[__popc(mask & ((2u << r) - 1u)) - 1] = r;
        if (r == 0) {
            p_begin[__popc(mask)] = nrows;
            n_piece_s = __popc(mask);
        }
    }
    __syncthreads();
    const int np = n_piece_s;
    const int cl = threadIdx.x & 127, rl = threadIdx.x >> 7;
    for (int k = rl; k < np; k += SM2_LANES) {
        const int r0 = p_begin[k], r1 = p_begin[k + 1];
        const int s = lab[r0 + 1];
        const bool is_head = (k == 0) && (lab[0] == s);              // the segment started in an earlier chunk
        const bool is_tail = (k == np - 1) && (lab[nrows + 1] == s);  // the segment continues in the next chunk
        for (int c4 = cl * 4; c4 < E; c4 += 512) {
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int r = r0; r < r1; ++r) {
                const float4 v = ld_vec4(sm2_rows + (size_t)r * row_bytes, x_dtype, c4);
                a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
            }
            if (is_head) {
                *reinterpret_cast<float4*>(head + ((size_t)b * nchunks + c) * E + c4) = a;
            } else if (is_tail) {
                *reinterpret_cast<float4*>(tail + ((size_t)b * nchunks + c) * E + c4) = a;
            } else {
                const float inv = 1.f / (float)(r1 - r0);
                st_vec4(seg, s_dtype, (base + s) * (size_t)lds + c4, make_float4(a.x * inv, a.y * inv, a.z * inv, a.w * inv));
            }
        }
    }
}

__global__ void __launch_bounds__(128) segment_mean_pass2_kernel(void* seg, int s_dtype, int lds,
                                                                 const int32_t* __restrict__ seg_label,
                                                                 const int32_t* __restrict__ seg_start,
                                                                 const int32_t* __restrict__ seg_len,
                                                                 const int32_t* __restrict__ nseg, int slot, int E,
                                                                 const float* __restrict__ head, const float* __restrict__ tail) {
    const int b = blockIdx.y, c = blockIdx.x, nchunks = gridDim.x;
    const int S = min(nseg[b], slot);
    if (S <= 0) return;
    const size_t base = (size_t)b * slot;
    const int f1 = (c + 1) * SM2_CHUNK;
    const int T = seg_start[base + S - 1] + seg_len[base + S - 1];
    if (f1 >= T) return;                                   // nothing continues past the last chunk of the video
    const int s = seg_label[base + f1 - 1];
    if (seg_label[base + f1] != s) return;                 // no segment crosses this chunk's end
    const int st = seg_start[base + s];
    if (st < f1 - SM2_CHUNK) return;                       // it started in an earlier chunk: that chunk's CTA sums it
    const int n = seg_len[base + s];
    const int c_end = (st + n - 1) / SM2_CHUNK;            // last chunk the segment touches
    const float inv = 1.f / (float)n;
    for (int c4 = threadIdx.x * 4; c4 < E; c4 += 512) {
        float4 a = *reinterpret_cast<const float4*>(tail + ((size_t)b * nchunks + c) * E + c4);
        for (int cc = c + 1; cc <= c_end; cc += 4) {       // four independent loads in flight, added in chunk order
            float4 v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                v[j] = (cc + j <= c_end) ? *reinterpret_cast<const float4*>(head + ((size_t)b * nchunks + cc + j) * E + c4)
                                         : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < 4; ++j) { a.x += v[j].x; a.y += v[j].y; a.z += v[j].z; a.w += v[j].w; }
        }
        st_vec4(seg, s_dtype, (base + s) * (size_t)lds + c4, make_float4(a.x * inv, a.y * inv, a.z * inv, a.w * inv));
    }
}

// ------------------------------------------------------------------------------------------------
// Bidirectional GRU recurrence.  One 4-CTA cluster advances NB chains (videos) of one direction; CTA `rank`
// owns hidden units [rank*U, rank*U+U).  The W_hh rows of those units live in REGISTERS (thread (kq, j) keeps
// the KS = Hh/4 weights of gate row j for K slice kq), so a step reads only the hidden state from shared memory.
// Per step: (1) every thread accumulates its K slice for the NB chains -> partials in shared memory,
// (2) one thread per (chain, unit) sums the partials, applies the gates and stores the new hidden value into
// the next-step buffer of all four CTAs with st.async (a distributed-shared-memory store that completes
// transaction bytes on the destination CTA's "hidden state ready" mbarrier), (3) consumers wait on their local
// barrier.  No fence and no cluster-wide barrier inside the loop; the global store of the output and the
// prefetch of the next input gates come after the exchange, off the critical path.
constexpr int GRU_CS = 4;   // CTAs per cluster
constexpr int GRU_KQ = 4;   // K split of the mat-vec inside a CTA

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

__device__ __forceinline__ uint32_t cluster_map_u32(const void* local_smem, int rank) {
    uint32_t laddr = (uint32_t)__cvta_generic_to_shared(local_smem), raddr;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(laddr), "r"(rank));
    return raddr;
}
__device__ __forceinline__ void mbar_arrive_remote_release(uint32_t raddr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
// st.async data is visible to a thread that observes the phase completion; CTA-scope acquire avoids the L1
// invalidate (CCTL.IVALL) a cluster-scope acquire would emit every step.
__device__ __forceinline__ void mbar_wait_acquire_cluster(uint64_t* bar, uint32_t parity) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
    uint32_t ok = 0, spins = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (ok) break;
        if (++spins > (1u << 24)) __trap();     // protocol bug: fail loudly instead of hanging the GPU
    }
}

template <int NB, int KS, int NACC>
__global__ void __cluster_dims__(GRU_CS, 1, 1) __launch_bounds__(GRU_KQ * 3 * KS, 1)
gru_cluster_kernel(const float* __restrict__ gi, const float* __restrict__ whh_f, const float* __restrict__ bhh_f,
                   const float* __restrict__ whh_b, const float* __restrict__ bhh_b, void* out, int o_dtype,
                   int ldo, int relu, int B, int slot, const int32_t* __restrict__ nseg) {
    constexpr int Hh = GRU_KQ * KS, U = KS, R = 3 * U;
    constexpr int GATE_T = NB * U, GATE_W = (GATE_T + 31) / 32;
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int cid = blockIdx.x / GRU_CS;
    const int dir = cid & 1, grp = cid >> 1;
    __shared__ __align__(16) float hb[2 * NB * Hh];          // double-buffered hidden state
    __shared__ float part[NB * GRU_KQ * R];                  // partial dot products
    __shared__ __align__(8) uint64_t hready[2];
    const float* whh = dir ? whh_b : whh_f;
    const float* bhh = dir ? bhh_b : bhh_f;
    const int tid = threadIdx.x, lane = tid & 31;
    const int kq = tid / R, j = tid % R;                     // blockDim.x == KQ * R exactly

    float w[KS];
    {
        const int g = j / U, u = j % U;
        const float* wr = whh + ((size_t)g * Hh + rank * U + u) * Hh + kq * KS;
#pragma unroll
        for (int k = 0; k < KS; ++k) w[k] = wr[k];
    }
    for (int i = tid; i < 2 * NB * Hh; i += blockDim.x) hb[i] = 0.f;
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            const uint32_t a = (uint32_t)__cvta_generic_to_shared(&hready[i]);
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(1) : "memory");
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int i = 0; i < 2; ++i) {   // arm both phases: every step delivers NB*Hh floats from the four CTAs
            const uint32_t a = (uint32_t)__cvta_generic_to_shared(&hready[i]);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(NB * Hh * 4) : "memory");
        }
    }

    int S[NB], vb[NB], maxS = 0;
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
        vb[nb] = grp * NB + nb;
        S[nb] = (vb[nb] < B) ? min(nseg[vb[nb]], slot) : 0;
        maxS = max(maxS, S[nb]);
    }
    const bool gate = tid < GATE_T;
    const int gnb = gate ? tid / U : 0, gu = gate ? tid % U : 0, unit = rank * U + gu;
    int myS = 0, myv = 0;
#pragma unroll
    for (int nb = 0; nb < NB; ++nb)
        if (nb == gnb) { myS = S[nb]; myv = vb[nb]; }
    if (!gate) myS = 0;
    float b_r = 0.f, b_z = 0.f, b_n = 0.f;
    if (gate) { b_r = bhh[unit]; b_z = bhh[Hh + unit]; b_n = bhh[2 * Hh + unit]; }
    const size_t gstride = (size_t)6 * Hh;
    float g_r = 0.f, g_z = 0.f, g_n = 0.f;
    auto load_gi = [&](int t) {
        if (t < myS) {
            const int s = dir ? myS - 1 - t : t;
            const float* p = gi + ((size_t)myv * slot + s) * gstride + (size_t)dir * 3 * Hh + unit;
            g_r = __ldg(p); g_z = __ldg(p + Hh); g_n = __ldg(p + 2 * Hh);
        }
    };
    load_gi(0);
    const uint32_t hb_u32 = (uint32_t)__cvta_generic_to_shared(hb);
    const uint32_t bar_u32 = (uint32_t)__cvta_generic_to_shared(&hready[0]);
    __syncthreads();
    cluster.sync();

    for (int t = 0; t < maxS; ++t) {
        const int cur = t & 1;
        if (t > 0) {
            if (lane == 0) mbar_wait_acquire_cluster(&hready[cur], ((t - 1) >> 1) & 1);
            __syncwarp();
            if (tid == 0) {   // re-arm for this buffer's next use (exchange t+1)
                const uint32_t a = (uint32_t)__cvta_generic_to_shared(&hready[cur]);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(NB * Hh * 4) : "memory");
            }
        }
        const float* h = hb + cur * NB * Hh;
        {
            float acc[NB][NACC];         // NACC independent chains per video shorten the FMA dependency chain
#pragma unroll
            for (int nb = 0; nb < NB; ++nb)
#pragma unroll
                for (int a = 0; a < NACC; ++a) acc[nb][a] = 0.f;
            const float* hk = h + kq * KS;
#pragma unroll
            for (int k = 0; k < KS; k += 4) {
#pragma unroll
                for (int nb = 0; nb < NB; ++nb) {
                    const float4 h4 = *reinterpret_cast<const float4*>(&hk[nb * Hh + k]);
                    acc[nb][0 % NACC] = fmaf(w[k], h4.x, acc[nb][0 % NACC]);
                    acc[nb][1 % NACC] = fmaf(w[k + 1], h4.y, acc[nb][1 % NACC]);
                    acc[nb][2 % NACC] = fmaf(w[k + 2], h4.z, acc[nb][2 % NACC]);
                    acc[nb][3 % NACC] = fmaf(w[k + 3], h4.w, acc[nb][3 % NACC]);
                }
            }
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) {
                float t = acc[nb][0];
#pragma unroll
                for (int a = 1; a < NACC; ++a) t += acc[nb][a];
                part[(nb * GRU_KQ + kq) * R + j] = t;
            }
        }
        __syncthreads();
        if (gate) {
            const bool live = t < myS;
            float hn = h[gnb * Hh + unit];      // finished chains re-send their last state: byte counts stay constant
            if (live) {
                const float* pp = part + gnb * GRU_KQ * R;
                float a_r = b_r, a_z = b_z, a_n = b_n;
#pragma unroll
                for (int q = 0; q < GRU_KQ; ++q) {
                    a_r += pp[q * R + gu]; a_z += pp[q * R + U + gu]; a_n += pp[q * R + 2 * U + gu];
                }
                const float r = sigmoidf_(g_r + a_r), z = sigmoidf_(g_z + a_z);
                const float n = tanhf(g_n + r * a_n);
                hn = (1.f - z) * n + z * hn;
            }
            const uint32_t o = (uint32_t)((cur ^ 1) * NB * Hh + gnb * Hh + unit) * 4u;
#pragma unroll
            for (int rr = 0; rr < GRU_CS; ++rr) {
                uint32_t ra, rb;
                asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(hb_u32 + o), "r"(rr));
                asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rb) : "r"(bar_u32 + (uint32_t)(cur ^ 1) * 8u), "r"(rr));
                asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];"
                             ::"r"(ra), "r"(__float_as_uint(hn)), "r"(rb) : "memory");
            }
            if (live) {
                const int s = dir ? myS - 1 - t : t;
                st_elem(out, o_dtype, ((size_t)myv * slot + s) * (size_t)ldo + (size_t)dir * Hh + unit, relu ? fmaxf(hn, 0.f) : hn);
            }
            load_gi(t + 1);
        }
    }
    cluster.sync();     // nobody may exit while peers can still write into its shared memory
}

}  // namespace factk

using namespace factk;

extern "C" int factk_tdu_segment(const int32_t* pred, int B, int slot, const int32_t* len, int32_t* seg_label,
                                 int32_t* seg_start, int32_t* seg_len, int32_t* seg_center, int32_t* nseg, void* stream) {
    FACTK_REQUIRE(pred && seg_label && seg_start && seg_len && seg_center && nseg && B > 0 && slot > 0,
                  "factk_tdu_segment: bad args");
    tdu_segment_kernel<<<B, 1024, 0, (cudaStream_t)stream>>>(pred, slot, len, seg_label, seg_start, seg_len, seg_center, nseg);
    return check_launch("factk_tdu_segment");
}

extern "C" size_t factk_segment_mean_ws_floats(int B, int slot, int E) {
    const size_t nchunks = (size_t)(slot + SM2_CHUNK - 1) / SM2_CHUNK;
    return 2 * (size_t)B * nchunks * (size_t)E;
}

extern "C" int factk_segment_mean(const void* X, int x_dtype, int ldx, void* seg, int s_dtype, int lds,
                                  const int32_t* seg_label, const int32_t* seg_start, const int32_t* seg_len, const int32_t* nseg,
                                  int B, int slot, int E, float* ws, void* stream) {
    FACTK_REQUIRE(X && seg && seg_label && seg_start && seg_len && nseg && B > 0 && slot > 0 && E > 0, "factk_segment_mean: bad args");
    FACTK_REQUIRE(B <= 65535, "factk_segment_mean: B too large");
    cudaStream_t st = (cudaStream_t)stream;
    const int es = x_dtype == FACTK_BF16 ? 2 : 4, ss = s_dtype == FACTK_BF16 ? 2 : 4;
    const size_t smem = (size_t)SM2_CHUNK * E * es;
    const bool stream_ok = ws != nullptr && aligned16(X) && aligned16(seg) && aligned16(ws) && (E % 8) == 0 && ((size_t)ldx * es) % 16 == 0 &&
                           ((size_t)lds * ss) % 16 == 0 && smem <= 160 * 1024;
    if (stream_ok) {
        const int nchunks = (slot + SM2_CHUNK - 1) / SM2_CHUNK;
        float* head = ws;
        float* tail = ws + (size_t)B * nchunks * E;
        static unsigned long long attr_devs = 0;
        if (first_use_on_device(attr_devs)) {
            cudaFuncSetAttribute(segment_mean_pass1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        }
        segment_mean_pass1_kernel<<<dim3(nchunks, B), 128 * SM2_LANES, smem, st>>>(X, x_dtype, ldx, seg, s_dtype, lds, seg_label, seg_start,
                                                                                   seg_len, nseg, slot, E, head, tail);
        segment_mean_pass2_kernel<<<dim3(nchunks, B), 128, 0, st>>>(seg, s_dtype, lds, seg_label, seg_start, seg_len, nseg, slot, E, head, tail);
        return check_launch("factk_segment_mean");
    }
    segment_mean_kernel<<<dim3((slot + SEGM_CHUNK - 1) / SEGM_CHUNK, B), 128 * SEGM_LANES, 0, st>>>(X, x_dtype, ldx, seg, s_dtype, lds, seg_label,
                                                                                                   seg_start, seg_len, nseg, slot, E);
    return check_launch("factk_segment_mean");
}

extern "C" int factk_gru_bidir(const float* gi, const float* w_hh_f, const float* b_hh_f, const float* w_hh_b,
                               const float* b_hh_b, int Hh, void* out, int o_dtype, int ldo, int relu, int B, int slot,
                               const int32_t* nseg, void* stream) {
    FACTK_REQUIRE(gi && w_hh_f && b_hh_f && w_hh_b && b_hh_b && out && nseg && B > 0 && slot > 0, "factk_gru_bidir: bad args");
    FACTK_REQUIRE(Hh == 32 || Hh == 64 || Hh == 128 || Hh == 256,
                  "factk_gru_bidir: hidden size per direction %d unsupported (32/64/128/256, i.e. hid_dim 64..512)", Hh);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // chains advanced per cluster: as few as possible while all clusters are co-resident (one CTA per SM)
    int NB = 1;
    while (NB < 4 && ((B + NB - 1) / NB) * 2 * GRU_CS > sms - 16) NB *= 2;
    const int groups = (B + NB - 1) / NB;
    const int KS = Hh / GRU_KQ;
    const int threads = GRU_KQ * 3 * KS;
    cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH(N_, K_)                                                                                                  \
    gru_cluster_kernel<N_, K_, 1><<<groups * 2 * GRU_CS, threads, 0, st>>>(gi, w_hh_f, b_hh_f, w_hh_b, b_hh_b, out, o_dtype, ldo, relu, B, \
                                                                             slot, nseg)
#define LAUNCH_K(N_)                                                \
    do {                                                            \
        if (KS == 8) LAUNCH(N_, 8);                                 \
        else if (KS == 16) LAUNCH(N_, 16);                          \
        else if (KS == 32) LAUNCH(N_, 32);                          \
        else LAUNCH(N_, 64);                                        \
    } while (0)
    if (NB == 1) LAUNCH_K(1); else if (NB == 2) LAUNCH_K(2); else LAUNCH_K(4);
#undef LAUNCH_K
#undef LAUNCH
    return check_launch("factk_gru_bidir");
}
