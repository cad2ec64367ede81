// Generalised multi-source GEMM on the 5th-generation tensor cores (same contract as gemm_simt.cu):
//   y[b,t,:N] = act(alpha * sum_s A_s[b, t + off_s] W_s[b]^T + bias + pre) + res
//
// Persistent, warp-specialised kernel (one CTA per SM, 192 threads):
//   warp 0      TMA producer  : cp.async.bulk.tensor 3-D boxes [128 B of K x 128 rows] of A (one box per
//                               convolution tap, shifted rows, hardware zero fill outside the video slot) and
//                               [128 B of K x BN rows] of W into a ring of 128B-swizzled shared-memory stages
//   warp 1      MMA issuer    : one thread issues tcgen05.mma (M=128, N=BN, K=16 bf16 / 8 tf32) from the
//                               shared-memory descriptors into one of two TMEM accumulators (2 x BN columns)
//   warps 2..5  epilogue      : tcgen05.ld (thread = output row, 32 columns at a time), bias / pre-activation
//                               addend / ReLU / residual in registers, vectorised row stores; overlaps the
//                               next tile's MMAs through the second accumulator
// Synchronisation is mbarrier-only (full/empty per stage, full/empty per accumulator).
// Replaces cuDNN implicit-GEMM Conv1d / cuBLAS Linear behind models/basic.py:138-139,158-164,177,182,
// 237-250 and models/blocks.py:153-159,402,414.
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>

#include "common.cuh"
#include "tc_common.cuh"

namespace factk {

struct TcParams {
    alignas(64) CUtensorMap amap[FACTK_MAX_SRC];
    alignas(64) CUtensorMap wmap[FACTK_MAX_SRC];
    int kchunks[FACTK_MAX_SRC];
    int row_off[FACTK_MAX_SRC];
    int w_batched[FACTK_MAX_SRC];
    int nsrc, B, slot, N;
    const int32_t* len;
    const float* bias;
    long long bias_bstride;
    float alpha;
    int relu;
    const void* pre;
    const int32_t* pre_idx;
    long long pre_bstride;
    int pre_dtype, ldpre;
    const void* res;
    int res_dtype, ldres;
    void* Y;
    int y_dtype, ldy;
    int tiles_m, tiles_n, total_tiles;
    int vec_io;   // rows of Y / res / pre are 16-byte aligned -> vector path
};

template <int BN>
struct TcCfg {
    static constexpr int STAGE_A = 128 * 128;
    static constexpr int STAGE_B = BN * 128;
    static constexpr int STAGE = STAGE_A + STAGE_B;
    static constexpr int NSTAGE = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
    static constexpr int STG = 8 * 32 * 20 * 4;   // per-epilogue-warp staging tile [32 rows][16 cols + 4 pad] fp32
    static constexpr int SMEM = NSTAGE * STAGE + 256 + STG + 1024;
};

__device__ __forceinline__ void ld_row32(const void* base, int dtype, size_t off, bool vec, float o[32]) {
    if (vec) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            const float4 v = ld_vec4(base, dtype, off + j);
            o[j] = v.x; o[j + 1] = v.y; o[j + 2] = v.z; o[j + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) o[j] = ld_elem(base, dtype, off + j);
    }
}

static_assert(TcCfg<256>::SMEM <= 232448 && TcCfg<128>::SMEM <= 232448 && TcCfg<64>::SMEM <= 232448, "dynamic smem limit");
constexpr int TC_EPI_WARPS = 8;
constexpr int TC_THREADS = 64 + TC_EPI_WARPS * 32;

__device__ __forceinline__ void sts128(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}

// YBF: output dtype is bf16 (else fp32).  RES: 0 = no residual, 1 = bf16 residual, 2 = fp32 residual.
template <int BN, bool TF32, bool YBF, int RES>
__global__ void __launch_bounds__(TC_THREADS, 1) gemm_tc_kernel(const __grid_constant__ TcParams p) {
    using Cfg = TcCfg<BN>;
    constexpr int EPC = TF32 ? 32 : 64;   // elements per 128-byte K chunk
    constexpr int NC = BN / 32;           // 32-column chunks per tile
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::NSTAGE * Cfg::STAGE);
    uint64_t* empty = full + Cfg::NSTAGE;
    uint64_t* tfull = empty + Cfg::NSTAGE;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < p.nsrc; ++s) {
            tc::tma_prefetch_desc(&p.amap[s]);
            tc::tma_prefetch_desc(&p.wmap[s]);
        }
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int i = 0; i < Cfg::NSTAGE; ++i) {
                tc::mbar_init(&full[i], 1);
                tc::mbar_init(&empty[i], 1);
            }
            for (int i = 0; i < 2; ++i) {
                tc::mbar_init(&tfull[i], 1);
                tc::mbar_init(&tempty[i], TC_EPI_WARPS * 32);
            }
            tc::fence_barrier_init();
        }
        __syncwarp();
        tc::tmem_alloc(tmem_slot, 2 * BN);
        tc::tmem_relinquish();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    auto tile_coords = [&](int tile, int& b, int& t0, int& nt, int& len_b) {
        nt = tile % p.tiles_n;
        const int rest = tile / p.tiles_n;
        const int mt = rest % p.tiles_m;
        b = rest / p.tiles_m;
        t0 = mt * 128;
        len_b = p.len ? min(p.len[b], p.slot) : p.slot;
        return t0 < len_b;
    };

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                int b, t0, nt, len_b;
                if (!tile_coords(tile, b, t0, nt, len_b)) continue;
                for (int s = 0; s < p.nsrc; ++s) {
                    for (int kc = 0; kc < p.kchunks[s]; ++kc) {
                        tc::mbar_wait(&empty[stage], phase ^ 1);
                        tc::mbar_arrive_expect_tx(&full[stage], Cfg::STAGE);
                        uint8_t* st = smem + stage * Cfg::STAGE;
                        tc::tma_load_3d(st, &p.amap[s], &full[stage], kc * EPC, t0 + p.row_off[s], b);
                        tc::tma_load_3d(st + Cfg::STAGE_A, &p.wmap[s], &full[stage], kc * EPC, nt * BN, p.w_batched[s] ? b : 0);
                        if (++stage == Cfg::NSTAGE) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = tc::instr_desc(128, BN, TF32);
            int stage = 0, it = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                int b, t0, nt, len_b;
                if (!tile_coords(tile, b, t0, nt, len_b)) continue;
                const int acc = it & 1;
                const uint32_t aphase = (it >> 1) & 1;
                ++it;
                tc::mbar_wait(&tempty[acc], aphase ^ 1);
                tc::tc_fence_after();
                const uint32_t d = tmem_base + acc * BN;
                uint32_t accumulate = 0;
                for (int s = 0; s < p.nsrc; ++s) {
                    for (int kc = 0; kc < p.kchunks[s]; ++kc) {
                        tc::mbar_wait(&full[stage], phase);
                        tc::tc_fence_after();
                        const uint32_t a_addr = tc::smem_u32(smem + stage * Cfg::STAGE);
                        const uint32_t b_addr = a_addr + Cfg::STAGE_A;
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4) {
                            tc::umma<TF32>(d, tc::smem_desc_sw128(a_addr + k4 * 32), tc::smem_desc_sw128(b_addr + k4 * 32),
                                           idesc, accumulate);
                            accumulate = 1;
                        }
                        tc::umma_commit(&empty[stage]);
                        if (++stage == Cfg::NSTAGE) { stage = 0; phase ^= 1; }
                    }
                }
                tc::umma_commit(&tfull[acc]);
            }
        }
    } else {
        // ---------------- epilogue: 8 warps; warp -> (TMEM lane quarter q, column half) ----------------
        const int ew = warp - 2;
        const int q = warp & 3;
        const int half = ew >> 2;
        constexpr int CPW = NC / 2;                       // chunks per warp
        const uint32_t stg = tc::smem_u32(smem + Cfg::NSTAGE * Cfg::STAGE + 256) + (uint32_t)ew * (32 * 20 * 4);
        const int pr = lane >> 2, cg = (lane & 3) * 4;   // phase 2: 4 lanes x 4 columns per row, 8 rows per pass, 16-column halves
        int it = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
            int b, t0, nt, len_b;
            if (!tile_coords(tile, b, t0, nt, len_b)) continue;
            const int acc = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            ++it;
            const size_t rowbase = (size_t)b * p.slot + t0 + q * 32;
            const int rows_valid = len_b - (t0 + q * 32);     // rows r < rows_valid of this warp's 32 are real
            const float* bias = p.bias ? p.bias + (size_t)b * (size_t)p.bias_bstride : nullptr;
            const int c_begin = half * CPW;

            uint4 rnext[8];
            auto fetch_res = [&](int c) {
                if constexpr (RES != 0) {
#pragma unroll
                    for (int pass = 0; pass < 8; ++pass) {
                        const int n = nt * BN + c * 32 + (pass >> 2) * 16 + cg;
                        const int r = (pass & 3) * 8 + pr;
                        rnext[pass] = make_uint4(0u, 0u, 0u, 0u);
                        if (p.vec_io && r < rows_valid && n + 4 <= p.N) {
                            const size_t off = (rowbase + r) * (size_t)p.ldres + n;
                            if constexpr (RES == 1) {
                                const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(p.res) + off);
                                rnext[pass].x = u.x; rnext[pass].y = u.y;
                            } else {
                                rnext[pass] = *reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(p.res) + off);
                            }
                        }
                    }
                }
            };
            // residual rows do not depend on the MMAs: fetch the first chunk before waiting for the accumulator
            if (nt * BN + c_begin * 32 < p.N) fetch_res(c_begin);
            tc::mbar_wait(&tfull[acc], aphase);
            tc::tc_fence_after();
#pragma unroll 1
            for (int ci = 0; ci < CPW; ++ci) {
                const int c = c_begin + ci;
                const int n0 = nt * BN + c * 32;
                if (n0 >= p.N) break;
                uint4 rcur[8];
                if constexpr (RES != 0) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) rcur[i] = rnext[i];
                    if (ci + 1 < CPW && n0 + 32 < p.N) fetch_res(c + 1);
                }
                float v[32];
                __syncwarp();
                tc::tmem_ld32(tmem_base + acc * BN + c * 32 + ((uint32_t)(q * 32) << 16), v);
                tc::tmem_ld_wait();
#pragma unroll
                for (int sub = 0; sub < 2; ++sub) {
                    // phase 1: thread = accumulator row -> staging tile [32 rows][16 cols (+4 pad)], conflict-free 128-bit stores
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 16; j += 4)
                        sts128(stg + (uint32_t)(lane * 20 + j) * 4u, v[sub * 16 + j], v[sub * 16 + j + 1], v[sub * 16 + j + 2], v[sub * 16 + j + 3]);
                    __syncwarp();
                    // phase 2: contiguous 4-column groups per lane -> coalesced row segments
                    const int n = n0 + sub * 16 + cg;
                    if (n >= p.N) continue;
                    const bool colv = p.vec_io && (n + 4 <= p.N);
                    float b4[4] = {0.f, 0.f, 0.f, 0.f};
                    if (bias) {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (n + j < p.N) b4[j] = __ldg(bias + n + j);
                    }
#pragma unroll
                    for (int ps = 0; ps < 4; ++ps) {
                        const int r = ps * 8 + pr;
                        if (r >= rows_valid) continue;
                        const int pass = sub * 4 + ps;
                        const size_t grow = rowbase + r;
                        const float4 a4 = lds128(stg + (uint32_t)(r * 20 + cg) * 4u);
                        float x[4] = {fmaf(a4.x, p.alpha, b4[0]), fmaf(a4.y, p.alpha, b4[1]), fmaf(a4.z, p.alpha, b4[2]),
                                      fmaf(a4.w, p.alpha, b4[3])};
                        if (p.pre) {
                            const size_t prow = (size_t)b * (size_t)p.pre_bstride +
                                                (size_t)(p.pre_idx ? p.pre_idx[grow] : (int)(t0 + q * 32 + r)) * (size_t)p.ldpre + n;
                            for (int j = 0; j < 4 && n + j < p.N; ++j) x[j] += ld_elem(p.pre, p.pre_dtype, prow + j);
                        }
                        if (p.relu) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) x[j] = fmaxf(x[j], 0.f);
                        }
                        if constexpr (RES != 0) {
                            if (colv) {
                                if constexpr (RES == 1) {
                                    const float2 lo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&rcur[pass].x));
                                    const float2 hi = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&rcur[pass].y));
                                    x[0] += lo.x; x[1] += lo.y; x[2] += hi.x; x[3] += hi.y;
                                } else {
                                    x[0] += __uint_as_float(rcur[pass].x); x[1] += __uint_as_float(rcur[pass].y);
                                    x[2] += __uint_as_float(rcur[pass].z); x[3] += __uint_as_float(rcur[pass].w);
                                }
                            } else {
                                for (int j = 0; j < 4 && n + j < p.N; ++j)
                                    x[j] += ld_elem(p.res, RES == 1 ? FACTK_BF16 : FACTK_F32, grow * (size_t)p.ldres + n + j);
                            }
                        }
                        const size_t yoff = grow * (size_t)p.ldy + n;
                        if (colv) {
                            st_vec4(p.Y, YBF ? FACTK_BF16 : FACTK_F32, yoff, make_float4(x[0], x[1], x[2], x[3]));
                        } else {
                            for (int j = 0; j < 4 && n + j < p.N; ++j) st_elem(p.Y, YBF ? FACTK_BF16 : FACTK_F32, yoff + j, x[j]);
                        }
                    }
                }
            }
            tc::tc_fence_before();
            tc::mbar_arrive(&tempty[acc]);
        }
    }

    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem_base, 2 * BN);
    }
}

// ------------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess) f = nullptr;
        return reinterpret_cast<EncodeTiledFn>(f);
    }();
    return fn;
}

// 3-D map {inner K, rows, batch} with a [128 B x box_rows x 1] box, 128B swizzle, zero OOB fill.  Cached.
bool tc_get_map(CUtensorMap* out, const void* ptr, int es, uint64_t k, uint64_t rows, uint64_t batch, uint64_t row_stride_el,
                    uint64_t batch_stride_el, uint32_t box_rows) {
    struct Key { const void* p; uint64_t a[7]; };
    Key key{ptr, {(uint64_t)es, k, rows, batch, row_stride_el, batch_stride_el, box_rows}};
    static std::mutex mu;
    static std::unordered_map<std::string, CUtensorMap> cache;
    std::string ks(reinterpret_cast<const char*>(&key), sizeof(key));
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(ks);
    if (it != cache.end()) { *out = it->second; return true; }
    EncodeTiledFn fn = encode_fn();
    if (!fn) { set_error("factk_gemm_tc: cuTensorMapEncodeTiled not available"); return false; }
    cuuint64_t dims[3] = {k, rows, batch};
    cuuint64_t strides[2] = {row_stride_el * es, (batch > 1 ? batch_stride_el : rows * row_stride_el) * es};
    cuuint32_t box[3] = {(cuuint32_t)(128 / es), box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUtensorMap m;
    CUresult r = fn(&m, es == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(ptr), dims,
                    strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("factk_gemm_tc: cuTensorMapEncodeTiled failed (%d) k=%llu rows=%llu batch=%llu stride=%llu", (int)r,
                  (unsigned long long)k, (unsigned long long)rows, (unsigned long long)batch, (unsigned long long)row_stride_el);
        return false;
    }
    if (cache.size() > 4096) cache.clear();
    cache.emplace(ks, m);
    *out = m;
    return true;
}

static const char* tc_unsupported_reason(const factk_gemm_t* g) {
    if (!g || g->nsrc < 1 || g->nsrc > FACTK_MAX_SRC) return "nsrc";
    if (g->B <= 0 || g->slot <= 0 || g->N <= 0 || !g->Y) return "shape";
    const int dt = g->src[0].a_dtype;
    if (dt != FACTK_BF16 && dt != FACTK_F32) return "dtype";
    const int es = dt == FACTK_BF16 ? 2 : 4;
    for (int s = 0; s < g->nsrc; ++s) {
        const factk_src_t& x = g->src[s];
        if (x.a_dtype != dt || x.w_dtype != dt) return "all sources must share one dtype (A and W)";
        if (!x.A || !x.W || x.K <= 0 || (x.K * es) % 128 != 0) return "K must be a multiple of 128 bytes";
        if (x.gather || x.pos) return "gather/pos not supported on tensor-core sources";
        if (x.a_slot <= 0) return "shared A rows";
        if ((x.lda * es) % 16 != 0 || (reinterpret_cast<uintptr_t>(x.A) & 15u)) return "A alignment";
        if ((x.ldw * es) % 16 != 0 || (reinterpret_cast<uintptr_t>(x.W) & 15u) || (x.w_bstride * es) % 16 != 0) return "W alignment";
        if (x.lda < x.K || x.ldw < x.K) return "leading dimension";
    }
    if (g->y_dtype != FACTK_BF16 && g->y_dtype != FACTK_F32) return "y dtype";
    return nullptr;
}

template <int BN, bool TF32, bool YBF, int RES>
static int launch_tc4(const TcParams& p, int grid, cudaStream_t st) {
    static unsigned long long attr_devs = 0;
    if (first_use_on_device(attr_devs)) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN, TF32, YBF, RES>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<BN>::SMEM);
        if (e != cudaSuccess) { set_error("factk_gemm_tc: smem attribute: %s", cudaGetErrorString(e)); return FACTK_ERR_CUDA; }
    }
    gemm_tc_kernel<BN, TF32, YBF, RES><<<grid, TC_THREADS, TcCfg<BN>::SMEM, st>>>(p);
    return check_launch("factk_gemm_tc");
}

template <int BN, bool TF32>
static int launch_tc(const TcParams& p, int grid, cudaStream_t st) {
    const bool ybf = p.y_dtype == FACTK_BF16;
    const int res = p.res == nullptr ? 0 : (p.res_dtype == FACTK_BF16 ? 1 : 2);
    if (ybf) {
        if (res == 0) return launch_tc4<BN, TF32, true, 0>(p, grid, st);
        if (res == 1) return launch_tc4<BN, TF32, true, 1>(p, grid, st);
        return launch_tc4<BN, TF32, true, 2>(p, grid, st);
    }
    if (res == 0) return launch_tc4<BN, TF32, false, 0>(p, grid, st);
    if (res == 1) return launch_tc4<BN, TF32, false, 1>(p, grid, st);
    return launch_tc4<BN, TF32, false, 2>(p, grid, st);
}

}  // namespace factk

using namespace factk;

extern "C" int factk_gemm_tc_supported(const factk_gemm_t* g) { return tc_unsupported_reason(g) == nullptr ? 1 : 0; }

extern "C" int factk_gemm_tc(const factk_gemm_t* g, void* stream) {
    const char* why = tc_unsupported_reason(g);
    if (why) { set_error("factk_gemm_tc: unsupported descriptor (%s)", why); return FACTK_ERR_UNSUPPORTED; }
    const bool tf32 = g->src[0].a_dtype == FACTK_F32;
    const int es = tf32 ? 4 : 2;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // N tile: the widest tile that covers N re-reads A least, but small problems (token rows: one 128-row tile per video)
    // do not fill the GPU with it -- pick the width that minimises waves x bytes loaded per tile and K chunk
    int BN = g->N <= 64 ? 64 : (g->N <= 128 ? 128 : 256);
    {
        const long long tm = (long long)g->B * ((g->slot + 127) / 128);
        long long best = -1;
        for (int bn = BN; bn >= 64; bn >>= 1) {
            const long long tiles = tm * ((g->N + bn - 1) / bn);
            const long long cost = ((tiles + sms - 1) / sms) * (16384 + bn * 128);
            if (best < 0 || cost < best) { best = cost; BN = bn; }
        }
    }
    TcParams p;
    memset(&p, 0, sizeof(p));
    for (int s = 0; s < g->nsrc; ++s) {
        const factk_src_t& x = g->src[s];
        if (!tc_get_map(&p.amap[s], x.A, es, x.K, x.a_slot, g->B, x.lda, (uint64_t)x.a_slot * x.lda, 128)) return FACTK_ERR_CUDA;
        const bool batched = x.w_bstride != 0;
        if (!tc_get_map(&p.wmap[s], x.W, es, x.K, g->N, batched ? g->B : 1, x.ldw, batched ? (uint64_t)x.w_bstride : (uint64_t)g->N * x.ldw, BN))
            return FACTK_ERR_CUDA;
        p.kchunks[s] = x.K * es / 128;
        p.row_off[s] = x.row_off;
        p.w_batched[s] = batched;
    }
    p.nsrc = g->nsrc; p.B = g->B; p.slot = g->slot; p.N = g->N; p.len = g->len;
    p.bias = g->bias; p.bias_bstride = g->bias_bstride; p.alpha = g->alpha; p.relu = g->relu;
    p.pre = g->pre; p.pre_idx = g->pre_idx; p.pre_bstride = g->pre_bstride; p.pre_dtype = g->pre_dtype; p.ldpre = g->ldpre;
    p.res = g->res; p.res_dtype = g->res_dtype; p.ldres = g->ldres;
    p.Y = g->Y; p.y_dtype = g->y_dtype; p.ldy = g->ldy;
    p.tiles_m = (g->slot + 127) / 128;
    p.tiles_n = (g->N + BN - 1) / BN;
    p.total_tiles = g->B * p.tiles_m * p.tiles_n;
    auto al = [](const void* ptr, int ld, int dtype) {
        const int e = dtype == FACTK_BF16 ? 2 : 4;
        return ptr == nullptr || (((reinterpret_cast<uintptr_t>(ptr) & 15u) == 0) && ((ld * e) % 16 == 0));
    };
    p.vec_io = al(g->Y, g->ldy, g->y_dtype) && al(g->res, g->ldres, g->res_dtype) && al(g->pre, g->ldpre, g->pre_dtype) &&
               (g->pre == nullptr || (g->pre_bstride * (g->pre_dtype == FACTK_BF16 ? 2 : 4)) % 16 == 0);
    const int grid = p.total_tiles < sms ? p.total_tiles : sms;
    cudaStream_t st = (cudaStream_t)stream;
    if (tf32) {
        if (BN == 64) return launch_tc<64, true>(p, grid, st);
        if (BN == 128) return launch_tc<128, true>(p, grid, st);
        return launch_tc<256, true>(p, grid, st);
    }
    if (BN == 64) return launch_tc<64, false>(p, grid, st);
    if (BN == 128) return launch_tc<128, false>(p, grid, st);
    return launch_tc<256, false>(p, grid, st);
}
