// Micro-benchmark: all-to-all exchange inside a thread-block cluster through distributed shared memory with the
// arrival flag carried in the data (the exchange pattern of gru_mma_kernel), cycles per round.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dsmem_pingpong dsmem_pingpong.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#include <stdint.h>

constexpr uint32_t EMPTY = 0xFFFFFFFFu;
constexpr int HBUF = 4224, HSTRIDE = 528;

__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t r) {
    uint32_t o;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(a), "r"(r));
    return o;
}

// work: extra dependent FMAs per round emulating the compute between receive and send
template <int CS, int MODE>
__global__ void __launch_bounds__(128, 1) pingpong(int rounds, int work, long long* out, float* sink) {
    __shared__ __align__(16) uint8_t hb[2 * HBUF];
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    for (int i = tid; i < 2 * HBUF / 4; i += 128) reinterpret_cast<uint32_t*>(hb)[i] = i < HBUF / 4 ? 0u : EMPTY;
    const uint32_t hb_u32 = (uint32_t)__cvta_generic_to_shared(hb);
    const int n = tid & 7;
    // units per CTA = 256 / CS; this thread's packet: 8 units of video n starting at unit rank*(256/CS) + 8*(tid>>5)... (CS=8 layout)
    constexpr int U = 256 / CS;
    constexpr int PK = U / 8;                 // 16-byte packets per (video, CTA)
    // each of the 4 lane groups (lane>>3) sends to CS/4 destinations (CS=8: 2, CS=4: 1, CS=2: lanes groups 0,1 only)
    const int grp = lane >> 3;
    uint32_t dst[8];
    int ndst = 0;
    for (int d = 0; d < CS; ++d)
        if (d % 4 == grp) dst[ndst++] = mapa(hb_u32, d);
    const uint32_t bsrc = hb_u32 + (uint32_t)(lane >> 2) * HSTRIDE + (uint32_t)(64 * w + (lane & 3) * 2) * 2u;
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    float acc = (float)tid;
    long long t0 = clock64();
    for (int t = 0; t < rounds; ++t) {
        const int cur = t & 1;
        uint32_t b[8];
        const uint32_t ba = bsrc + cur * HBUF;
        uint32_t spins = 0;
        while (true) {
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                asm volatile("ld.volatile.shared.b32 %0, [%1];" : "=r"(b[2 * s]) : "r"(ba + 32u * s) : "memory");
                asm volatile("ld.volatile.shared.b32 %0, [%1];" : "=r"(b[2 * s + 1]) : "r"(ba + 32u * s + 16u) : "memory");
            }
            bool ok = true;
#pragma unroll
            for (int s = 0; s < 8; ++s) ok = ok && b[s] != EMPTY;
            if (MODE == 3 || __all_sync(0xffffffffu, ok)) break;
            if (++spins > (1u << 22)) __trap();
        }
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(ba + 32u * s), "r"(EMPTY) : "memory");
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(ba + 32u * s + 16u), "r"(EMPTY) : "memory");
        }
        for (int i = 0; i < work; ++i) acc = fmaf(acc, 1.0001f, (float)b[i & 7] * 1e-30f);
        __syncthreads();
        const uint32_t word = (uint32_t)(t + 1) & 0x7FFF7FFFu;
        const uint32_t nxt = cur ^ 1;
        // this CTA owns units [rank*U, rank*U+U): 8 videos x PK packets, each to CS destinations = 8*PK*CS stores over 128 threads
        for (int j = tid; j < 8 * PK * CS; j += 128) {
            const int d = j % CS, pk = (j / CS) % PK, nn = j / (CS * PK);
            const uint32_t off = (uint32_t)nn * HSTRIDE + (uint32_t)(rank * U + 8 * pk) * 2u + nxt * HBUF;
            const uint32_t a = mapa(hb_u32, d) + off;
            if constexpr (MODE == 0 || MODE == 3) asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(a), "r"(word) : "memory");
            if constexpr (MODE == 1) asm volatile("st.volatile.shared::cluster.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(a), "r"(word) : "memory");
            if constexpr (MODE == 2) asm volatile("st.relaxed.cluster.shared::cluster.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(a), "r"(word) : "memory");
        }
        if constexpr (MODE == 3) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    long long t1 = clock64();
    if (tid == 0) out[blockIdx.x] = t1 - t0;
    if (acc == 12345.678f) sink[0] = acc;
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int CS, int MODE>
void run(int nclusters, int rounds, int work) {
    long long* out;
    float* sink;
    cudaMalloc(&out, sizeof(long long) * nclusters * CS);
    cudaMalloc(&sink, 4);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(nclusters * CS);
    cfg.blockDim = dim3(128);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    for (int rep = 0; rep < 2; ++rep) {
        cudaError_t e = cudaLaunchKernelEx(&cfg, pingpong<CS, MODE>, rounds, work, out, sink);
        if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); return; }
        e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return; }
    }
    long long h[256];
    cudaMemcpy(h, out, sizeof(long long) * nclusters * CS, cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < nclusters * CS; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("mode=%d cluster=%d nclusters=%2d work=%4d : %7.1f cycles/round\n", MODE, CS, nclusters, work, (double)mx / rounds);
    cudaFree(out); cudaFree(sink);
}

// Pull variant: every CTA keeps its own slice locally as {word, tag} pairs; readers fetch their K-slice words from the
// owners with remote loads and retry until the tag matches the round.
template <int CS>
__global__ void __launch_bounds__(128, 1) pull_kernel(int rounds, int work, long long* out, float* sink) {
    constexpr int U = 256 / CS;
    __shared__ __align__(16) uint2 own[2][8][U / 2 + 1];     // [parity][video][unit pair] {word, tag}
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    for (int i = tid; i < 2 * 8 * (U / 2 + 1); i += 128) (&own[0][0][0])[i] = make_uint2(0u, 0u);
    const uint32_t own_u32 = (uint32_t)__cvta_generic_to_shared(&own[0][0][0]);
    // reader lane: video nb = lane>>2, unit pairs kp = (64 w + 16 s + (lane&3)*2 (+8)) / 2
    uint32_t src[8];
    for (int s = 0; s < 4; ++s)
        for (int h = 0; h < 2; ++h) {
            const int k = 64 * w + 16 * s + (lane & 3) * 2 + 8 * h;
            const int owner = k / U, kp = (k % U) / 2;
            src[2 * s + h] = mapa(own_u32 + (uint32_t)(((lane >> 2) * (U / 2 + 1) + kp) * 8), owner);
        }
    const uint32_t pbytes = 8 * (U / 2 + 1) * 8;
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    float acc = (float)tid;
    long long t0 = clock64();
    for (int t = 0; t < rounds; ++t) {
        const int cur = t & 1;
        uint32_t b[8], tg[8];
        uint32_t spins = 0;
        while (true) {
#pragma unroll
            for (int s = 0; s < 8; ++s)
                asm volatile("ld.relaxed.cluster.shared::cluster.v2.b32 {%0, %1}, [%2];" : "=r"(b[s]), "=r"(tg[s]) : "r"(src[s] + cur * pbytes) : "memory");
            bool ok = true;
#pragma unroll
            for (int s = 0; s < 8; ++s) ok = ok && tg[s] == (uint32_t)t;
            if (__all_sync(0xffffffffu, ok)) break;
            if (++spins > (1u << 22)) __trap();
        }
        for (int i = 0; i < work; ++i) acc = fmaf(acc, 1.0001f, (float)b[i & 7] * 1e-30f);
        __syncthreads();
        // owner: 8 videos x U/2 pairs = 4U words; threads tid < 4U write one pair each
        if (tid < 4 * U) {
            const int n = tid & 7, kp = tid >> 3;
            own[cur ^ 1][n][kp] = make_uint2((uint32_t)(t + 1) & 0x7FFF7FFFu, (uint32_t)(t + 1));
        }
    }
    long long t1 = clock64();
    if (tid == 0) out[blockIdx.x] = t1 - t0;
    if (acc == 12345.678f) sink[0] = acc;
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int CS>
void run_pull(int nclusters, int rounds, int work) {
    long long* out;
    float* sink;
    cudaMalloc(&out, sizeof(long long) * nclusters * CS);
    cudaMalloc(&sink, 4);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(nclusters * CS);
    cfg.blockDim = dim3(128);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    for (int rep = 0; rep < 2; ++rep) {
        cudaError_t e = cudaLaunchKernelEx(&cfg, pull_kernel<CS>, rounds, work, out, sink);
        if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); return; }
        e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return; }
    }
    long long h[256];
    cudaMemcpy(h, out, sizeof(long long) * nclusters * CS, cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < nclusters * CS; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("pull   cluster=%d nclusters=%2d work=%4d : %7.1f cycles/round\n", CS, nclusters, work, (double)mx / rounds);
    cudaFree(out); cudaFree(sink);
}

int main() {
    const int rounds = 2000;
    for (int work : {0}) {
        for (int nc : {1, 4}) run<8, 0>(nc, rounds, work);
        for (int nc : {1, 4}) run<8, 3>(nc, rounds, work);
        for (int nc : {1, 4}) run<4, 0>(nc, rounds, work);
        for (int nc : {1, 4}) run<4, 3>(nc, rounds, work);
        for (int nc : {1, 4}) run<2, 0>(nc, rounds, work);
        for (int nc : {1, 4}) run<2, 3>(nc, rounds, work);
    }
    return 0;
}
