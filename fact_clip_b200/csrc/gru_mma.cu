// Bidirectional GRU recurrence over segments (nn.GRU(H, H/2, bidirectional=True), models/blocks.py:401,432) batched
// over videos on the tensor cores -- the bf16-mode replacement of gru_cluster_kernel (tdu.cu, fp32 FFMA).
//
// The recurrence h_t = GRU(gi_t, h_{t-1}) is sequential in t, but the SAME W_hh serves every video of a direction,
// so one step of 8 videos is a [768 x 256] x [256 x 8] product: mma.sync m16n8k16 (bf16 in, fp32 accumulate) with
// M = gate rows, N = videos, K = hidden units.  One 8-CTA cluster advances 8 videos of one direction:
//   * CTA `rank` owns hidden units [32 rank, 32 rank + 32): 96 gate rows (r, z, n) = 6 m-tiles; its W_hh rows live in
//     REGISTERS as A fragments for the whole kernel (warp w keeps the K slice [64 w, 64 w + 64): 6 x 4 x 4 registers);
//   * per step every warp reads its K slice of the hidden state (bf16 [video][unit] in shared memory, conflict-free
//     32-bit loads = B fragments), issues 24 HMMAs, and the four K slices are reduced through shared memory;
//   * one thread per (video, unit pair) applies the gates (fp32 state carried in registers, MUFU.TANH based
//     sigmoid / tanh), and the new hidden values travel to all eight CTAs as plain 16-byte distributed-shared-memory
//     stores into the next step's buffer;
//   * there is NO barrier of any kind in the exchange: every hidden-state word is written exactly once per step and read
//     by exactly one lane, so the word itself carries the arrival flag -- the reader spins on its B-fragment loads until
//     none of them holds the "empty" marker (a bf16 NaN pair that is never produced as data) and re-marks them empty.
//     (st.async / bulk copies completing transaction bytes on a remote mbarrier were measured at ~1000 cycles per step
//     for this exchange, the flag-in-data stores at one remote-store latency.)
//   * output stores and the prefetch of the next input gates are off the critical path.
// Hidden size per direction is fixed at 256 (hid_dim 512: every shipped configuration, SURVEY.md note N3).
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace factk {

constexpr int GM_CS = 8;                    // CTAs per cluster
constexpr int GM_NV = 8;                    // videos per cluster (MMA N)
constexpr int GM_HH = 256;                  // hidden units per direction
constexpr int GM_U = GM_HH / GM_CS;         // 32 units per CTA
constexpr int GM_ROWS = 3 * GM_U;           // 96 gate rows per CTA
constexpr int GM_HSTRIDE = GM_HH * 2 + 16;  // bytes per video row of the hidden-state buffer (padded: conflict-free B loads)
constexpr int GM_HBUF = GM_NV * GM_HSTRIDE; // 4224 bytes per buffer
constexpr int GM_PSTRIDE = GM_ROWS + 4;     // floats per video row of the partial-sum buffer
constexpr int GM_THREADS = 128;
#ifndef GM_TAIL_FIRST
#define GM_TAIL_FIRST 0      // 1: output store + input-gate prefetch BEFORE the exchange stores.  Measured: the step's tail
                             // shrinks from ~350 to 20 cycles but the send is delayed by as much: 0.79 instead of 0.715 us per step
                             // up to 8 clusters, 0.98 instead of 1.01 at 16 -- kept off
#endif
#ifndef GM_SPIN_SLEEP
#define GM_SPIN_SLEEP 0
#endif
constexpr uint32_t GM_EMPTY = 0xFFFFFFFFu;  // "not written yet" marker of a hidden-state word

__device__ __forceinline__ uint32_t gm_pack(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float gm_tanh(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float gm_sigmoid(float x) { return fmaf(0.5f, gm_tanh(0.5f * x), 0.5f); }
__device__ __forceinline__ void gm_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t gm_mapa(uint32_t local_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}

constexpr int GM_PART = 4 * GM_NV * GM_PSTRIDE;                                  // floats per partial-sum buffer
constexpr int GM_GDEPTH = 6;                                                      // input-gate ring: loads issued 5 steps ahead
constexpr int GM_GRING = GM_GDEPTH * 3 * GM_THREADS * 8;                          // bytes per chain: [depth][gate][thread] float2
constexpr int gm_smem_bytes(int nc) { return nc * 2 * GM_HBUF + nc * 2 * GM_PART * 4 + nc * GM_GRING; }   // dynamic shared memory

// GM_NC = independent 8-video groups (chains) interleaved on one cluster.  With two chains, while the hidden state of chain A travels through distributed shared memory (~700 cycles), the
// cluster computes the step of chain B (~600 cycles), so a pair of steps costs about what one step costs alone.
template <int GM_NC>
__global__ void __launch_bounds__(GM_THREADS, 1)
gru_mma_kernel(const float* __restrict__ gi, const float* __restrict__ whh_f, const float* __restrict__ bhh_f,
               const float* __restrict__ whh_b, const float* __restrict__ bhh_b, void* out, int o_dtype, int ldo, int relu,
               int B, int slot, const int32_t* __restrict__ nseg, const int32_t* __restrict__ order, long long* dbg) {
    extern __shared__ __align__(16) uint8_t gm_smem[];
    uint8_t* hb = gm_smem;                                                       // [chain][parity] hidden state, bf16 [video][unit]
    // K-slice partial sums [chain][step parity][warp][video][gate row]: parity-double-buffered because a warp only waits for
    // the two CTAs that own its K slice, so it may start the next step while other warps still read the partials of this one
    float* part = reinterpret_cast<float*>(gm_smem + GM_NC * 2 * GM_HBUF);
    uint8_t* gring = gm_smem + GM_NC * 2 * GM_HBUF + GM_NC * 2 * GM_PART * 4;        // [chain][depth][gate][thread] float2

    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int cid = blockIdx.x / GM_CS;
    const int dir = cid & 1, grp = cid >> 1;
    const float* whh = dir ? whh_b : whh_f;
    const float* bhh = dir ? bhh_b : bhh_f;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;

    // ---- W_hh rows of this CTA as mma A fragments (bf16), resident in registers
    uint32_t A[6][4][4];
#pragma unroll
    for (int t6 = 0; t6 < 6; ++t6) {
        const int g = t6 >> 1, uh = t6 & 1;
        const float* r0 = whh + (size_t)(g * GM_HH + rank * GM_U + uh * 16 + (lane >> 2)) * GM_HH;
        const float* r1 = r0 + 8 * GM_HH;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            const int k0 = 64 * w + 16 * s + (lane & 3) * 2;
            const float2 x0 = __ldg(reinterpret_cast<const float2*>(r0 + k0)), x1 = __ldg(reinterpret_cast<const float2*>(r1 + k0));
            const float2 x2 = __ldg(reinterpret_cast<const float2*>(r0 + k0 + 8)), x3 = __ldg(reinterpret_cast<const float2*>(r1 + k0 + 8));
            A[t6][s][0] = gm_pack(x0.x, x0.y);
            A[t6][s][1] = gm_pack(x1.x, x1.y);
            A[t6][s][2] = gm_pack(x2.x, x2.y);
            A[t6][s][3] = gm_pack(x3.x, x3.y);
        }
    }
    // per chain: buffer 0 = h_0 = 0; buffer 1 = "empty" (filled by the exchange of step 0)
    for (int i = tid; i < GM_NC * 2 * GM_HBUF / 4; i += GM_THREADS)
        reinterpret_cast<uint32_t*>(hb)[i] = ((i / (GM_HBUF / 4)) & 1) == 0 ? 0u : GM_EMPTY;
    const uint32_t hb_u32 = (uint32_t)__cvta_generic_to_shared(hb);

    // ---- gate-phase role of this thread: video n of each chain, units (2 up, 2 up + 1) of this CTA
    const int n = tid & 7, up = tid >> 3;
    const int unit = rank * GM_U + 2 * up;
    const float2 b_r = *reinterpret_cast<const float2*>(bhh + unit);
    const float2 b_z = *reinterpret_cast<const float2*>(bhh + GM_HH + unit);
    const float2 b_n = *reinterpret_cast<const float2*>(bhh + 2 * GM_HH + unit);
    const size_t gstride = (size_t)6 * GM_HH;
    int vb[GM_NC], myS[GM_NC], maxS[GM_NC];
    float2 hprev[GM_NC], g_r[GM_NC], g_z[GM_NC], g_n[GM_NC];
    int maxAll = 0;
#pragma unroll
    for (int c = 0; c < GM_NC; ++c) {
        // chain slot -> video: by index, or through `order` (videos sorted by segment count, so that chains of similar
        // length share a cluster and whole clusters retire early: fewer co-running clusters = shorter steps for the rest)
        auto video_of = [&](int sidx) { return sidx < B ? (order ? order[sidx] : sidx) : B; };
        vb[c] = video_of((grp * GM_NC + c) * GM_NV + n);
        myS[c] = vb[c] < B ? min(nseg[vb[c]], slot) : 0;
        maxS[c] = 0;
        for (int i = 0; i < GM_NV; ++i) {
            const int v = video_of((grp * GM_NC + c) * GM_NV + i);
            maxS[c] = max(maxS[c], v < B ? min(nseg[v], slot) : 0);
        }
        maxAll = max(maxAll, maxS[c]);
        hprev[c] = g_r[c] = g_z[c] = g_n[c] = make_float2(0.f, 0.f);
    }
    // The input gates stream from HBM (B x S x 6 Hh floats).  They are fetched GM_GDEPTH-1 steps ahead with cp.async into a
    // per-thread ring in shared memory: a real load cannot be dropped the way an L2 prefetch hint is under load (with 16
    // clusters the hinted version lost 0.3 us per step to HBM misses), and five steps (~4 us) cover the HBM latency.
    // Running pointers (one add per step) keep the address arithmetic off the step's critical path.
    const long long gstep = dir ? -(long long)gstride : (long long)gstride;     // segment order: forward / reverse
    const long long ostep = dir ? -(long long)ldo : (long long)ldo;
    const float* gp[GM_NC];          // input gates of the NEXT step to fetch (this thread's video / units)
    long long oidx[GM_NC];           // output element index of step t
    int gfetched[GM_NC];             // steps fetched so far
    const uint32_t gring_u32 = (uint32_t)__cvta_generic_to_shared(gring) + (uint32_t)tid * 8u;
#pragma unroll
    for (int c = 0; c < GM_NC; ++c) {
        const int s0 = dir ? myS[c] - 1 : 0;
        gp[c] = gi + ((size_t)vb[c] * slot + (myS[c] > 0 ? s0 : 0)) * gstride + (size_t)dir * 3 * GM_HH + unit;
        oidx[c] = ((long long)vb[c] * slot + (myS[c] > 0 ? s0 : 0)) * (long long)ldo + (long long)dir * GM_HH + unit;
        gfetched[c] = 0;
    }
    // fetch the gates of step gfetched[c] into ring slot gfetched[c] % depth; one commit group per call (possibly empty)
    auto fetch_gi = [&](int c) {
        if (gfetched[c] < myS[c]) {
            const uint32_t dst = gring_u32 + (uint32_t)(c * GM_GRING + (gfetched[c] % GM_GDEPTH) * 3 * GM_THREADS * 8);
#pragma unroll
            for (int g = 0; g < 3; ++g)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + (uint32_t)(g * GM_THREADS * 8)), "l"(gp[c] + g * GM_HH) : "memory");
            gp[c] += gstep;
        }
        ++gfetched[c];
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // the chains share one cp.async group sequence: prologue fetches GM_GDEPTH-1 steps of every chain, chain-major per step
    for (int d = 0; d < GM_GDEPTH - 1; ++d)
#pragma unroll
        for (int c = 0; c < GM_NC; ++c) fetch_gi(c);
    // exchange: lanes (n, n+8, n+16, n+24) of a warp hold 8 consecutive units of video n; lane group j = lane / 8 stores that
    // 16-byte packet into the next-step buffer of CTAs 2j and 2j+1
    const uint32_t xoff = (uint32_t)n * GM_HSTRIDE + (uint32_t)(rank * GM_U + 8 * w) * 2u;
    uint32_t dst_h[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) dst_h[i] = gm_mapa(hb_u32, 2 * (lane >> 3) + i) + xoff;
    const uint32_t bsrc = hb_u32 + (uint32_t)(lane >> 2) * GM_HSTRIDE + (uint32_t)(64 * w + (lane & 3) * 2) * 2u;   // B fragments
    float* pw = part + (size_t)w * GM_NV * GM_PSTRIDE + (size_t)((lane & 3) * 2) * GM_PSTRIDE + (lane >> 2);        // partial stores
    const float* pr = part + (size_t)n * GM_PSTRIDE + 2 * up;                                                      // partial loads

    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");

    const bool dbg_on = dbg != nullptr && blockIdx.x == 0 && tid == 0;
    for (int t = 0; t < maxAll; ++t) {
        const int cur = t & 1;
#pragma unroll
        for (int c = 0; c < GM_NC; ++c) {
            if (t >= maxS[c]) {                  // uniform over the cluster: this chain has finished everywhere
                asm volatile("cp.async.commit_group;" ::: "memory");   // keep one group per (step, chain)
                continue;
            }
            if (dbg_on && t < 64) dbg[t * 16 + c * 8 + 0] = clock64();
            // ---- this step's input gates out of the ring (fetched GM_GDEPTH-1 steps ago: the wait never stalls); done BEFORE
            // waiting for the hidden state so that it is off the critical path of the exchange
            const bool live = t < myS[c];
            // groups are committed chain-major, one per (step, chain): everything up to (t, c) must have landed
            if (GM_NC == 1) asm volatile("cp.async.wait_group %0;" ::"n"(GM_GDEPTH - 2) : "memory");
            else if (c == 0) asm volatile("cp.async.wait_group %0;" ::"n"(2 * (GM_GDEPTH - 2) + 1) : "memory");
            else asm volatile("cp.async.wait_group %0;" ::"n"(2 * (GM_GDEPTH - 2)) : "memory");
            if (live) {
                const uint8_t* gsrc = gring + (size_t)c * GM_GRING + (size_t)(t % GM_GDEPTH) * 3 * GM_THREADS * 8 + (size_t)tid * 8;
                g_r[c] = *reinterpret_cast<const float2*>(gsrc);
                g_z[c] = *reinterpret_cast<const float2*>(gsrc + GM_THREADS * 8);
                g_n[c] = *reinterpret_cast<const float2*>(gsrc + 2 * GM_THREADS * 8);
            }
            // ---- B fragments of this warp's K slice: spin until every word has landed, then mark the words empty again
            uint32_t b[4][2];
            {
                const uint32_t ba = bsrc + (uint32_t)(c * 2 + cur) * GM_HBUF;
                uint32_t spins = 0;
                while (true) {
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                        asm volatile("ld.volatile.shared.b32 %0, [%1];" : "=r"(b[s][0]) : "r"(ba + 32u * s) : "memory");
                        asm volatile("ld.volatile.shared.b32 %0, [%1];" : "=r"(b[s][1]) : "r"(ba + 32u * s + 16u) : "memory");
                    }
                    bool ok = true;
#pragma unroll
                    for (int s = 0; s < 4; ++s) ok = ok && (b[s][0] != GM_EMPTY) && (b[s][1] != GM_EMPTY);
                    if (__all_sync(0xffffffffu, ok)) break;
                    if (GM_SPIN_SLEEP) __nanosleep(GM_SPIN_SLEEP);
                    if (++spins > (1u << 22)) __trap();   // protocol bug: fail loudly instead of hanging the GPU
                }
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    asm volatile("st.shared.b32 [%0], %1;" ::"r"(ba + 32u * s), "r"(GM_EMPTY) : "memory");
                    asm volatile("st.shared.b32 [%0], %1;" ::"r"(ba + 32u * s + 16u), "r"(GM_EMPTY) : "memory");
                }
            }
            if (dbg_on && t < 64) dbg[t * 16 + c * 8 + 1] = clock64();
            // ---- [96 x 64] x [64 x 8] per warp on the tensor cores
            float acc[6][4];
#pragma unroll
            for (int t6 = 0; t6 < 6; ++t6) acc[t6][0] = acc[t6][1] = acc[t6][2] = acc[t6][3] = 0.f;
#pragma unroll
            for (int s = 0; s < 4; ++s)
#pragma unroll
                for (int t6 = 0; t6 < 6; ++t6) gm_mma(acc[t6], A[t6][s], b[s][0], b[s][1]);
            float* pwc = pw + (size_t)(c * 2 + cur) * GM_PART;
#pragma unroll
            for (int t6 = 0; t6 < 6; ++t6) {
                pwc[16 * t6] = acc[t6][0];
                pwc[GM_PSTRIDE + 16 * t6] = acc[t6][1];
                pwc[16 * t6 + 8] = acc[t6][2];
                pwc[GM_PSTRIDE + 16 * t6 + 8] = acc[t6][3];
            }
            if (dbg_on && t < 64) dbg[t * 16 + c * 8 + 2] = clock64();
            __syncthreads();
            if (dbg_on && t < 64) dbg[t * 16 + c * 8 + 3] = clock64();
            // ---- gates for (video n, units 2up, 2up+1)
            float2 a_r = b_r, a_z = b_z, a_n = b_n;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float* pq = pr + (size_t)(c * 2 + cur) * GM_PART + (size_t)q * GM_NV * GM_PSTRIDE;
                const float2 x0 = *reinterpret_cast<const float2*>(pq), x1 = *reinterpret_cast<const float2*>(pq + GM_U),
                             x2 = *reinterpret_cast<const float2*>(pq + 2 * GM_U);
                a_r.x += x0.x; a_r.y += x0.y; a_z.x += x1.x; a_z.y += x1.y; a_n.x += x2.x; a_n.y += x2.y;
            }
            float2 hn = hprev[c];               // finished videos re-send their last state: every word is written every step
            if (live) {
                const float r0 = gm_sigmoid(g_r[c].x + a_r.x), r1 = gm_sigmoid(g_r[c].y + a_r.y);
                const float z0 = gm_sigmoid(g_z[c].x + a_z.x), z1 = gm_sigmoid(g_z[c].y + a_z.y);
                const float n0 = gm_tanh(fmaf(r0, a_n.x, g_n[c].x)), n1 = gm_tanh(fmaf(r1, a_n.y, g_n[c].y));
                hn.x = fmaf(z0, hprev[c].x - n0, n0);
                hn.y = fmaf(z1, hprev[c].y - n1, n1);
                hprev[c] = hn;
            }
#if GM_TAIL_FIRST
            if (live) {
                const float y0 = relu ? fmaxf(hn.x, 0.f) : hn.x, y1 = relu ? fmaxf(hn.y, 0.f) : hn.y;
                if (o_dtype == FACTK_BF16) *reinterpret_cast<uint32_t*>(reinterpret_cast<__nv_bfloat16*>(out) + oidx[c]) = gm_pack(y0, y1);
                else *reinterpret_cast<float2*>(reinterpret_cast<float*>(out) + oidx[c]) = make_float2(y0, y1);
                oidx[c] += ostep;
            }
            fetch_gi(c);                        // step t + GM_GDEPTH - 1 of this chain
#endif
            // The data word doubles as the arrival flag, so it must never equal GM_EMPTY.  cvt.rn.bf16x2 canonicalises NaN to
            // 0x7FFF and finite values never produce 0xFFFF; the select below makes that explicit (two NaN inputs with any
            // payload still exchange as canonical NaNs instead of looking "not written yet" and spinning into the trap).
            // Ordering assumption of the protocol (documented, not enforced by fences): a reader re-marks a word EMPTY with a
            // plain st.shared only AFTER it has consumed the peer's value, and the peer overwrites the same word two steps
            // later, after it has itself consumed a value this CTA produced after the re-mark (causality through the
            // exchanged data); sm_100 executes same-address shared-memory stores of one SM in arrival order.
            uint32_t word = gm_pack(hn.x, hn.y);
            word = (word == GM_EMPTY) ? 0x7FFF7FFFu : word;
            const uint32_t w0 = __shfl_sync(0xffffffffu, word, lane & 7), w1 = __shfl_sync(0xffffffffu, word, (lane & 7) + 8),
                           w2 = __shfl_sync(0xffffffffu, word, (lane & 7) + 16), w3 = __shfl_sync(0xffffffffu, word, (lane & 7) + 24);
            const uint32_t nxt = (uint32_t)(c * 2 + (cur ^ 1));
#pragma unroll
            for (int i = 0; i < 2; ++i)
                asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst_h[i] + nxt * GM_HBUF), "r"(w0), "r"(w1), "r"(w2), "r"(w3) : "memory");
            if (dbg_on && t < 64) dbg[t * 16 + c * 8 + 4] = clock64();
#if !GM_TAIL_FIRST
            if (live) {
                const float y0 = relu ? fmaxf(hn.x, 0.f) : hn.x, y1 = relu ? fmaxf(hn.y, 0.f) : hn.y;
                if (o_dtype == FACTK_BF16) *reinterpret_cast<uint32_t*>(reinterpret_cast<__nv_bfloat16*>(out) + oidx[c]) = gm_pack(y0, y1);
                else *reinterpret_cast<float2*>(reinterpret_cast<float*>(out) + oidx[c]) = make_float2(y0, y1);
                oidx[c] += ostep;
            }
            fetch_gi(c);                        // step t + GM_GDEPTH - 1 of this chain
#endif
            if (dbg_on && t < 64) dbg[t * 16 + c * 8 + 5] = clock64();
        }
    }
    // nobody may exit while peers can still write into its shared memory
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

}  // namespace factk

using namespace factk;

// order[r] = index of the video with the r-th largest segment count (ties by index): one small CTA, O(B^2 / 256).
__global__ void gru_rank_kernel(const int32_t* __restrict__ nseg, int B, int32_t* __restrict__ order) {
    for (int i = threadIdx.x; i < B; i += blockDim.x) {
        const int si = nseg[i];
        int r = 0;
        for (int j = 0; j < B; ++j) {
            const int sj = nseg[j];
            r += (sj > si) || (sj == si && j < i);
        }
        order[r] = i;
    }
}

static int gru_mma_launch(const float* gi, const float* w_hh_f, const float* b_hh_f, const float* w_hh_b, const float* b_hh_b, int Hh,
                          void* out, int o_dtype, int ldo, int relu, int B, int slot, const int32_t* nseg, const int32_t* order,
                          long long* dbg, void* stream);

extern "C" int factk_gru_bidir_mma(const float* gi, const float* w_hh_f, const float* b_hh_f, const float* w_hh_b,
                                   const float* b_hh_b, int Hh, void* out, int o_dtype, int ldo, int relu, int B, int slot,
                                   const int32_t* nseg, void* stream) {
    return gru_mma_launch(gi, w_hh_f, b_hh_f, w_hh_b, b_hh_b, Hh, out, o_dtype, ldo, relu, B, slot, nseg, nullptr, nullptr, stream);
}

extern "C" int factk_gru_bidir_mma_sorted(const float* gi, const float* w_hh_f, const float* b_hh_f, const float* w_hh_b,
                                          const float* b_hh_b, int Hh, void* out, int o_dtype, int ldo, int relu, int B, int slot,
                                          const int32_t* nseg, int32_t* order_ws, void* stream) {
    FACTK_REQUIRE(nseg && order_ws && B > 0, "factk_gru_bidir_mma_sorted: bad args");
    gru_rank_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(nseg, B, order_ws);
    int rc = check_launch("factk_gru_bidir_mma_sorted(rank)");
    if (rc) return rc;
    return gru_mma_launch(gi, w_hh_f, b_hh_f, w_hh_b, b_hh_b, Hh, out, o_dtype, ldo, relu, B, slot, nseg, order_ws, nullptr, stream);
}

extern "C" int factk_gru_bidir_mma_dbg(const float* gi, const float* w_hh_f, const float* b_hh_f, const float* w_hh_b,
                                       const float* b_hh_b, int Hh, void* out, int o_dtype, int ldo, int relu, int B, int slot,
                                       const int32_t* nseg, long long* dbg, void* stream) {
    return gru_mma_launch(gi, w_hh_f, b_hh_f, w_hh_b, b_hh_b, Hh, out, o_dtype, ldo, relu, B, slot, nseg, nullptr, dbg, stream);
}

static int gru_mma_launch(const float* gi, const float* w_hh_f, const float* b_hh_f, const float* w_hh_b, const float* b_hh_b, int Hh,
                          void* out, int o_dtype, int ldo, int relu, int B, int slot, const int32_t* nseg, const int32_t* order,
                          long long* dbg, void* stream) {
    FACTK_REQUIRE(gi && w_hh_f && b_hh_f && w_hh_b && b_hh_b && out && nseg && B > 0 && slot > 0, "factk_gru_bidir_mma: bad args");
    FACTK_REQUIRE(Hh == GM_HH, "factk_gru_bidir_mma: hidden size per direction must be %d (got %d)", GM_HH, Hh);
    FACTK_REQUIRE((ldo % 2) == 0 && aligned16(gi) && aligned16(b_hh_f) && aligned16(b_hh_b) &&
                      (reinterpret_cast<uintptr_t>(out) & 7u) == 0,
                  "factk_gru_bidir_mma: alignment");
    // one chain per cluster while all clusters are co-resident (lowest latency per step); two interleaved chains beyond that
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    static const int nc_env = [] { const char* e = getenv("FACTK_GRU_CHAINS"); return e ? atoi(e) : 0; }();
    const int groups8 = (B + GM_NV - 1) / GM_NV;
    const int NC = nc_env == 1 || nc_env == 2 ? nc_env : (groups8 * 2 * GM_CS <= sms ? 1 : 2);
    const int groups = (groups8 + NC - 1) / NC;
    static unsigned long long attr_devs = 0;
    if (first_use_on_device(attr_devs)) {
        cudaError_t ea = cudaFuncSetAttribute(gru_mma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (ea == cudaSuccess) ea = cudaFuncSetAttribute(gru_mma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (ea != cudaSuccess) { set_error("factk_gru_bidir_mma: smem attribute: %s", cudaGetErrorString(ea)); return FACTK_ERR_CUDA; }
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(groups * 2 * GM_CS, 1, 1);
    cfg.blockDim = dim3(GM_THREADS, 1, 1);
    cfg.dynamicSmemBytes = gm_smem_bytes(NC);
    // Two of these CTAs fit one SM, and with 16 clusters (64 videos) the scheduler does co-locate CTAs of different clusters:
    // 0.71 us per step up to 8 clusters, 0.80 up to 14, 1.01 at 16 (tools/bench_gru_scale.py).  Forbidding the co-location
    // (FACTK_GRU_SMEM_KB=120: more than half of an SM's shared memory) is WORSE at 16 clusters -- only 15 exclusive 8-CTA
    // clusters fit the 148 SMs, the 16th becomes a second wave (1.43 us per step) -- so it is off by default; grouping the
    // videos by segment count (factk_gru_bidir_mma_sorted) is what recovers most of the loss.
    static const int excl_kb = [] { const char* e = getenv("FACTK_GRU_SMEM_KB"); return e ? atoi(e) : 0; }();
    if (excl_kb * 1024 > (int)cfg.dynamicSmemBytes) cfg.dynamicSmemBytes = (size_t)excl_kb * 1024;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = GM_CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = NC == 1 ? cudaLaunchKernelEx(&cfg, gru_mma_kernel<1>, gi, w_hh_f, b_hh_f, w_hh_b, b_hh_b, out, o_dtype, ldo, relu, B, slot, nseg, order, dbg)
                            : cudaLaunchKernelEx(&cfg, gru_mma_kernel<2>, gi, w_hh_f, b_hh_f, w_hh_b, b_hh_b, out, o_dtype, ldo, relu, B, slot, nseg, order, dbg);
    if (e != cudaSuccess) { set_error("factk_gru_bidir_mma: launch: %s", cudaGetErrorString(e)); return FACTK_ERR_CUDA; }
    return check_launch("factk_gru_bidir_mma");
}

// Diagnostic: how many 8-CTA clusters of gru_mma_kernel<1> can be co-resident on the current device.
extern "C" int factk_gru_max_clusters(void) {
    cudaFuncSetAttribute(gru_mma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, gm_smem_bytes(1));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(32 * GM_CS, 1, 1);
    cfg.blockDim = dim3(GM_THREADS, 1, 1);
    cfg.dynamicSmemBytes = gm_smem_bytes(1);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = GM_CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = -1;
    if (cudaOccupancyMaxActiveClusters(&n, gru_mma_kernel<1>, &cfg) != cudaSuccess) { cudaGetLastError(); return -1; }
    return n;
}
