// conv.cu — output-stationary gather-GEMM for sparse convolution (fp32 CUDA-core path).
//
// Replaces spconv v1.2 `indice_conv` / `indice_conv_backward` (per offset: gather kernel -> cuBLAS SGEMM ->
// scatter-add kernel).  One launch per conv instead of 1 + 26*3:
//   * a CTA owns BM output rows x BN output channels and loops over the kernel offsets that have at
//     least one neighbour in the tile; neighbour rows are gathered with 16-byte cp.async (zero-filled
//     for missing neighbours) straight into a 3-stage shared-memory ring, W[k] chunks ride in the same ring;
//   * every output row is written exactly once (float4 stores) -> no scatter atomics, no [P, C]
//     gather/scatter buffers crossing HBM;
//   * dgrad is the same kernel on the mirrored/transposed weights; wgrad walks the canonical pair lists.
#include "common.cuh"
#include <algorithm>
#include <stdlib.h>
#include <string.h>

namespace b200sp {

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

struct GGParams {
    const float* in;
    const float* W;       // [K][Cin][Cout]
    const int* tab;       // TAB mode: [n_rows][K] input rows (or NULL with K==1: identity)
    const int* orow;      // TAB mode: output row of table row r (NULL: identity)
    const int* pin;       // PAIRS mode: [K][pstride] input rows
    const int* pout;      // PAIRS mode: [K][pstride] output rows
    const int* pairnum;   // PAIRS mode: device [K]
    float* out;
    int64_t n_rows;       // TAB mode: number of output rows
    int64_t pstride;
    int Cin, Cout, K;
    int accumulate;
    int pairs_mode;
};

constexpr int GG_BK = 16;
constexpr int GG_STAGES = 3;
constexpr int GG_MAXK = 32;

template <int BM, int BN, int TM, int TN>
struct GGSmem {
    float A[GG_STAGES][BM][GG_BK + 4];
    float Wt[GG_STAGES][GG_BK][BN];
    int idx[BM * GG_MAXK];  // TAB: [BM][K]; PAIRS: [BM]
    int orow[BM];
    int klist[GG_MAXK];
    int kflag[GG_MAXK];
    int nk;
};

template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN)) k_gather_gemm(GGParams p) {
    constexpr int NT = (BM / TM) * (BN / TN);
    constexpr int RT = BM / TM;  // row-threads
    constexpr int CT = BN / TN;  // col-threads
    static_assert(TN % 4 == 0, "TN must be a multiple of 4");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    GGSmem<BM, BN, TM, TN>& sm = *reinterpret_cast<GGSmem<BM, BN, TM, TN>*>(smem_raw);

    const int tid = threadIdx.x;
    const int tx = tid % CT, ty = tid / CT;
    const int n0 = blockIdx.y * BN;
    const int K = p.K;
    int64_t row0 = (int64_t)blockIdx.x * BM;
    int rows;
    int kfixed = 0;
    const int KT = p.pairs_mode ? 1 : K;  // idx columns held per row

    if (p.pairs_mode) {
        kfixed = blockIdx.z;
        int n = p.pairnum[kfixed];
        if (row0 >= n) return;
        rows = (int)min((int64_t)BM, n - row0);
        for (int r = tid; r < BM; r += NT) {
            bool ok = r < rows;
            sm.idx[r] = ok ? p.pin[(int64_t)kfixed * p.pstride + row0 + r] : -1;
            sm.orow[r] = ok ? p.pout[(int64_t)kfixed * p.pstride + row0 + r] : -1;
        }
        if (tid == 0) {
            sm.nk = 1;
            sm.klist[0] = 0;
        }
    } else {
        rows = (int)min((int64_t)BM, p.n_rows - row0);
        if (tid < GG_MAXK) sm.kflag[tid] = 0;
        __syncthreads();
        if (p.tab) {
            for (int i = tid; i < BM * K; i += NT) {
                int v = (i < rows * K) ? p.tab[row0 * K + i] : -1;
                sm.idx[i] = v;
                if (v >= 0) sm.kflag[i % K] = 1;
            }
        } else {  // identity rows (1x1 conv / dense GEMM)
            for (int r = tid; r < BM; r += NT) sm.idx[r] = r < rows ? (int)(row0 + r) : -1;
            if (tid == 0) sm.kflag[0] = 1;
        }
        for (int r = tid; r < BM; r += NT) sm.orow[r] = r < rows ? (p.orow ? p.orow[row0 + r] : (int)(row0 + r)) : -1;
        __syncthreads();
        if (tid == 0) {
            int nk = 0;
            for (int k = 0; k < K; ++k)
                if (sm.kflag[k]) sm.klist[nk++] = k;
            sm.nk = nk;
        }
    }
    __syncthreads();

    const int nk = sm.nk;
    const int Cin = p.Cin, Cout = p.Cout;
    const int nck = (Cin + GG_BK - 1) / GG_BK;
    const int nsteps = nk * nck;
    const bool vecA = (Cin % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.in) & 15) == 0);
    const bool vecW = (Cout % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.W) & 15) == 0);

    auto load_step = [&](int step, int stage) {
        int kk = step / nck;
        int c0 = (step - kk * nck) * GG_BK;
        int kcol = sm.klist[kk];             // column in idx
        int kw = p.pairs_mode ? kfixed : kcol;  // weight slice
        // A: BM rows x 4 chunks of 16 B
        for (int i = tid; i < BM * (GG_BK / 4); i += NT) {
            int r = i >> 2, c4 = i & 3;
            int src = sm.idx[r * KT + kcol];
            int col = c0 + c4 * 4;
            float* dst = &sm.A[stage][r][c4 * 4];
            if (vecA) {
                bool ok = (src >= 0) && (col < Cin);
                const float* g = ok ? p.in + (int64_t)src * Cin + col : p.in;
                cp_async16(dst, g, ok ? 16 : 0);
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    bool ok = (src >= 0) && (col + e < Cin);
                    const float* g = ok ? p.in + (int64_t)src * Cin + col + e : p.in;
                    cp_async4(dst + e, g, ok ? 4 : 0);
                }
            }
        }
        // W: BK rows x BN/4 chunks
        const float* Wk = p.W + (int64_t)kw * Cin * Cout;
        for (int i = tid; i < GG_BK * (BN / 4); i += NT) {
            int r = i / (BN / 4), c4 = i % (BN / 4);
            int ci = c0 + r, col = n0 + c4 * 4;
            float* dst = &sm.Wt[stage][r][c4 * 4];
            if (vecW) {
                bool ok = (ci < Cin) && (col < Cout);
                const float* g = ok ? Wk + (int64_t)ci * Cout + col : p.W;
                cp_async16(dst, g, ok ? 16 : 0);
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    bool ok = (ci < Cin) && (col + e < Cout);
                    const float* g = ok ? Wk + (int64_t)ci * Cout + col + e : p.W;
                    cp_async4(dst + e, g, ok ? 4 : 0);
                }
            }
        }
    };

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

#pragma unroll
    for (int s = 0; s < GG_STAGES - 1; ++s) {
        if (s < nsteps) load_step(s, s);
        cp_async_commit();
    }

    for (int step = 0; step < nsteps; ++step) {
        cp_async_wait<GG_STAGES - 2>();
        __syncthreads();
        {
            int nxt = step + GG_STAGES - 1;
            if (nxt < nsteps) load_step(nxt, nxt % GG_STAGES);
            cp_async_commit();
        }
        const int stage = step % GG_STAGES;
#pragma unroll
        for (int kk = 0; kk < GG_BK; kk += 4) {
            float4 a[TM];
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] = *reinterpret_cast<const float4*>(&sm.A[stage][ty + i * RT][kk]);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                float w[TN];
#pragma unroll
                for (int q = 0; q < TN / 4; ++q) {
                    float4 w4 = *reinterpret_cast<const float4*>(&sm.Wt[stage][kk + r][tx * 4 + q * (CT * 4)]);
                    w[q * 4 + 0] = w4.x;
                    w[q * 4 + 1] = w4.y;
                    w[q * 4 + 2] = w4.z;
                    w[q * 4 + 3] = w4.w;
                }
#pragma unroll
                for (int i = 0; i < TM; ++i) {
                    float av = r == 0 ? a[i].x : r == 1 ? a[i].y : r == 2 ? a[i].z : a[i].w;
#pragma unroll
                    for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av, w[j], acc[i][j]);
                }
            }
        }
    }
    cp_async_wait<0>();

    // epilogue: thread owns rows ty + i*RT, column groups tx*4 + q*(CT*4)
    const bool vecO = (Cout % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0);
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int orow = sm.orow[ty + i * RT];
        if (orow < 0) continue;
        float* o = p.out + (int64_t)orow * Cout;
#pragma unroll
        for (int q = 0; q < TN / 4; ++q) {
            int col = n0 + tx * 4 + q * (CT * 4);
            if (col >= Cout) continue;
            if (vecO) {
                float4 v = make_float4(acc[i][q * 4 + 0], acc[i][q * 4 + 1], acc[i][q * 4 + 2], acc[i][q * 4 + 3]);
                if (p.accumulate) {
                    float4 old = *reinterpret_cast<const float4*>(o + col);
                    v.x += old.x; v.y += old.y; v.z += old.z; v.w += old.w;
                }
                *reinterpret_cast<float4*>(o + col) = v;
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    if (col + e < Cout) {
                        float v = acc[i][q * 4 + e];
                        if (p.accumulate) v += o[col + e];
                        o[col + e] = v;
                    }
                }
            }
        }
    }
}

template <int BM, int BN, int TM, int TN>
static int launch_gg(const GGParams& p, int64_t tiles_rows, int zdim, cudaStream_t st) {
    constexpr int NT = (BM / TM) * (BN / TN);
    size_t smem = sizeof(GGSmem<BM, BN, TM, TN>);
    static bool attr_done = false;
    if (!attr_done) {
        B200SP_CUDA(cudaFuncSetAttribute(k_gather_gemm<BM, BN, TM, TN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem));
        attr_done = true;
    }
    dim3 grid((unsigned)cdiv(tiles_rows, BM), (unsigned)cdiv(p.Cout, BN), (unsigned)zdim);
    k_gather_gemm<BM, BN, TM, TN><<<grid, NT, smem, st>>>(p);
    B200SP_LAUNCH_CHECK();
    return B200SP_OK;
}

static int dispatch_gg(const GGParams& p, int64_t tile_rows, int zdim, cudaStream_t st) {
    // tile choice: wide row tiles when there are enough rows to fill 148 SMs, small tiles for the deep levels
    const int64_t big_tiles = cdiv(tile_rows, 128) * cdiv(p.Cout, 64) * zdim;
    if (p.Cout <= 16) {
        if (tile_rows >= 128 * 148) return launch_gg<128, 16, 4, 4>(p, tile_rows, zdim, st);
        return launch_gg<32, 16, 1, 4>(p, tile_rows, zdim, st);
    }
    if (p.Cout <= 32) {
        if (tile_rows >= 128 * 148) return launch_gg<128, 32, 8, 4>(p, tile_rows, zdim, st);
        return launch_gg<32, 32, 2, 4>(p, tile_rows, zdim, st);
    }
    if (big_tiles >= 148) return launch_gg<128, 64, 8, 4>(p, tile_rows, zdim, st);
    return launch_gg<32, 32, 2, 4>(p, tile_rows, zdim, st);
}

// ---------------- wgrad over canonical pair lists ----------------
struct WGParams {
    const float* a;   // [*, Ca]
    const float* b;   // [*, Cb]
    const int* pa;    // [K][pstride] or NULL (identity)
    const int* pb;
    const int* pairnum;  // device [K] or NULL (identity: n_rows)
    float* dW;        // [K][Ca][Cb]
    int64_t pstride;
    int64_t n_rows;
    int Ca, Cb, K;
    int ta, tb;       // tile extents (multiples of 4, <= 64)
    int ppb;          // pairs per block
};

__global__ void __launch_bounds__(256) k_wgrad(WGParams p) {
    __shared__ float s_red[256 * 16];
    const int k = blockIdx.y;
    const int64_t n = p.pairnum ? p.pairnum[k] : p.n_rows;
    const int64_t p0 = (int64_t)blockIdx.x * p.ppb;
    if (p0 >= n) return;
    const int64_t p1 = min(n, p0 + p.ppb);
    const int tilesB = (p.Cb + p.tb - 1) / p.tb;
    const int a0 = (blockIdx.z / tilesB) * p.ta, b0 = (blockIdx.z % tilesB) * p.tb;
    const int nB = p.tb / 4, RT = (p.ta / 4) * nB;
    const int slices = 256 / RT;
    const int tid = threadIdx.x;
    const int slice = tid / RT, within = tid % RT;
    const int ia = within / nB, ib = within % nB;
    const int ca = a0 + ia * 4, cb = b0 + ib * 4;
    const bool active = slice < slices;
    const bool vecA = (p.Ca % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.a) & 15) == 0);
    const bool vecB = (p.Cb % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.b) & 15) == 0);
    const int* pa = p.pa ? p.pa + (int64_t)k * p.pstride : nullptr;
    const int* pb = p.pb ? p.pb + (int64_t)k * p.pstride : nullptr;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    auto ld4 = [](const float* base, int64_t row, int C, int c, bool vec) -> float4 {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        const float* q = base + row * C + c;
        if (vec) {
            if (c < C) v = __ldg(reinterpret_cast<const float4*>(q));
        } else {
            if (c + 0 < C) v.x = __ldg(q + 0);
            if (c + 1 < C) v.y = __ldg(q + 1);
            if (c + 2 < C) v.z = __ldg(q + 2);
            if (c + 3 < C) v.w = __ldg(q + 3);
        }
        return v;
    };

    if (active) {
        constexpr int U = 4;
        for (int64_t i = p0 + slice; i < p1; i += (int64_t)slices * U) {
            float4 av[U], bv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                int64_t ii = i + (int64_t)u * slices;
                if (ii < p1) {
                    int64_t ra = pa ? pa[ii] : ii;
                    int64_t rb = pb ? pb[ii] : ii;
                    av[u] = ld4(p.a, ra, p.Ca, ca, vecA);
                    bv[u] = ld4(p.b, rb, p.Cb, cb, vecB);
                } else {
                    av[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    bv[u] = av[u];
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                float a4[4] = {av[u].x, av[u].y, av[u].z, av[u].w};
                float b4[4] = {bv[u].x, bv[u].y, bv[u].z, bv[u].w};
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y) acc[x][y] = fmaf(a4[x], b4[y], acc[x][y]);
            }
        }
    }
    // cross-slice reduction through shared memory, then one atomic per element per block
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) s_red[(x * 4 + y) * 256 + tid] = active ? acc[x][y] : 0.f;
    __syncthreads();
    if (tid < RT) {
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y) {
                float s = 0.f;
                for (int sl = 0; sl < slices; ++sl) s += s_red[(x * 4 + y) * 256 + sl * RT + tid];
                if (ca + x < p.Ca && cb + y < p.Cb)
                    atomicAdd(&p.dW[((int64_t)k * p.Ca + ca + x) * p.Cb + cb + y], s);
            }
    }
}

__global__ void k_weight_transpose(const float* __restrict__ W, int K, int Ci, int Co, int mirror,
                                   float* __restrict__ out) {
    // out[k'][co][ci] = W[k][ci][co], k' = mirror ? K-1-k : k
    int64_t n = (int64_t)K * Ci * Co;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int ci = (int)(i % Ci);
        int64_t t = i / Ci;
        int co = (int)(t % Co);
        int kp = (int)(t / Co);
        int k = mirror ? K - 1 - kp : kp;
        out[i] = W[((int64_t)k * Ci + ci) * Co + co];
    }
}

}  // namespace b200sp

using namespace b200sp;

namespace b200sp {
int conv_tc_run(const float* in, int Cin, const float* W, int Ci_w, int Co_w, int wflags, const int* tab, const int* orow,
                const int* rowmask, const int* pin,
                const int* pout, const int* pairnum, int64_t n_rows, int64_t pstride, int K, float* out, int Cout,
                int accumulate, int pairs_mode, void* ws, int64_t ws_bytes, cudaStream_t st, int64_t n_in,
                const float* res = nullptr);
int64_t conv_tc_ws_bytes(int K, int Cin, int Cout);
bool conv_direct_covers(int K, int Cin, int Cout);
void conv_direct_set(int on);
int conv_direct_run(const float* in, int Cin, const float* W, int wflags, const int* tab, const int* orow,
                    const int* rowmask, long long n_rows, int K, float* out, int Cout, int accumulate, cudaStream_t st,
                    const float* res = nullptr);
int conv_tc_prep_batch(const int64_t* desc_host, int n, void* desc_dev, int64_t desc_dev_bytes, cudaStream_t st);
bool wgrad_direct_covers(int K, int Ca, int Cb);
int wgrad_direct_run(const float* a, int Ca, const float* g, int Cb, const int* tab, const int* orow, const int* rowmask,
                     long long n_rows, int K, float* dW, cudaStream_t st);
int wgrad_tc_run(const float* a, int Ca, const float* b, int Cb, const int* pa, const int* pb, const int* pairnum,
                 int64_t n_upper, int K, int64_t pstride, float* dW, cudaStream_t st);
bool wgrad_os_covers(int K, int Ca, int Cb);
int wgrad_os_run(const float* a, int Ca, const float* g, int Cb, const int* tab, const int* orow, const int* rowmask,
                 long long n_rows, int K, float* dW, cudaStream_t st);

// 0 = tensor-core path (tcgen05 3xTF32) whenever the shape is covered, 1 = fp32 CUDA-core kernel only
static int g_conv_impl = -1;
static int conv_impl() {
    if (g_conv_impl < 0) {
        const char* e = getenv("B200SP_CONV_IMPL");
        g_conv_impl = (e && strcmp(e, "fp32") == 0) ? 1 : 0;
    }
    return g_conv_impl;
}

// fp32 kernel wants W as [K][Cin][Cout]; wflags bit0 = W is stored [K][Cout][Cin], bit1 = mirrored offsets
static int fp32_weights(const float* W, int K, int Cin, int Cout, int wflags, void* ws, int64_t ws_bytes,
                        cudaStream_t st, const float** Wuse) {
    if (wflags == 0) {
        *Wuse = W;
        return B200SP_OK;
    }
    const int64_t need = (int64_t)K * Cin * Cout * 4;
    if (!ws || ws_bytes < need) {
        set_error("gather_gemm: workspace too small for the transposed weights (%lld < %lld)", (long long)ws_bytes,
                  (long long)need);
        return B200SP_ENOMEM;
    }
    B200SP_CHECK_ARG(wflags & 1, "gather_gemm: mirror without transpose is not used");
    int64_t n = need / 4;
    // stored [K][Cout][Cin] (Ci_w = Cout, Co_w = Cin) -> [K][Cin][Cout]
    k_weight_transpose<<<(unsigned)std::min<int64_t>(cdiv(n, 256), 2048), 256, 0, st>>>(W, K, Cout, Cin, (wflags >> 1) & 1,
                                                                                      static_cast<float*>(ws));
    B200SP_LAUNCH_CHECK();
    *Wuse = static_cast<const float*>(ws);
    return B200SP_OK;
}
}  // namespace b200sp

extern "C" int b200sp_set_conv_impl(int impl) {
    B200SP_CHECK_ARG(impl == 0 || impl == 1, "set_conv_impl: 0 = tensor cores, 1 = fp32 CUDA cores");
    b200sp::g_conv_impl = impl;
    return B200SP_OK;
}

extern "C" int64_t b200sp_conv_prepared_bytes(int K, int Cin, int Cout) {
    if (b200sp::conv_direct_covers(K, Cin, Cout)) return 0;  // the register-gather kernel reads the raw weights
    return b200sp::conv_tc_ws_bytes(K, Cin, Cout);
}

namespace b200sp {
void conv_tc_set_tma(int on);
}
extern "C" int b200sp_set_conv_tma(int on) {
    b200sp::conv_tc_set_tma(on);
    return B200SP_OK;
}

extern "C" int b200sp_set_conv_direct(int on) {
    b200sp::conv_direct_set(on);
    return B200SP_OK;
}

extern "C" int b200sp_conv_direct_covers(int K, int Cin, int Cout) { return b200sp::conv_direct_covers(K, Cin, Cout) ? 1 : 0; }

extern "C" int b200sp_prep_weights_batch(const int64_t* desc_host, int n, void* desc_dev, int64_t desc_dev_bytes,
                                         void* stream) {
    B200SP_CHECK_ARG(desc_host && n >= 0 && desc_dev, "prep_weights_batch: bad arguments");
    return b200sp::conv_tc_prep_batch(desc_host, n, desc_dev, desc_dev_bytes, (cudaStream_t)stream);
}

extern "C" int64_t b200sp_conv_ws_bytes(int K, int Cin, int Cout) {
    int64_t a = b200sp::conv_tc_ws_bytes(K, Cin, Cout);
    int64_t b = (int64_t)K * Cin * Cout * 4;
    return align_up(a > b ? a : b, 256);
}

namespace b200sp {
namespace {
__global__ void __launch_bounds__(256) k_add_rows(float* __restrict__ a, const float* __restrict__ b, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) a[i] += __ldg(b + i);
}
}  // namespace

// res != NULL: out = conv + res (res [n_out][Cout], may not alias out); accumulate: out += conv
int gather_gemm_impl(const float* in, int64_t n_in, int Cin, const float* W, int wflags, const int32_t* tab,
                     const int32_t* orow, const int32_t* rowmask, int K, float* out, int64_t n_out, int Cout,
                     int accumulate, const float* res, void* ws, int64_t ws_bytes, void* stream) {
    B200SP_CHECK_ARG(Cin >= 1 && Cout >= 1 && K >= 1, "gather_gemm: bad Cin/Cout/K");
    B200SP_CHECK_ARG(K <= GG_MAXK, "gather_gemm: K=%d > %d not supported by this build", K, GG_MAXK);
    B200SP_CHECK_ARG(tab || K == 1, "gather_gemm: tab==NULL requires K==1");
    B200SP_CHECK_ARG(!(res && accumulate), "gather_gemm: a residual and accumulate exclude each other");
    if (n_out == 0) return B200SP_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (conv_impl() == 0) {
        const int Ci_w = (wflags & 1) ? Cout : Cin, Co_w = (wflags & 1) ? Cin : Cout;
        int rc = conv_direct_run(in, Cin, W, wflags, tab, orow, rowmask, n_out, K, out, Cout, accumulate, st, res);
        if (rc != B200SP_EUNSUP) return rc;
        rc = conv_tc_run(in, Cin, W, Ci_w, Co_w, wflags, tab, orow, rowmask, nullptr, nullptr, nullptr, n_out, 0, K, out, Cout,
                             accumulate, 0, ws, ws_bytes, st, tab ? n_in : n_out, res);
        if (rc != B200SP_EUNSUP) return rc;
    }
    B200SP_CHECK_ARG(!(wflags & 4), "gather_gemm: a prepared weight image needs the tensor path");
    const float* Wuse = nullptr;
    int rc = fp32_weights(W, K, Cin, Cout, wflags, ws, ws_bytes, st, &Wuse);
    if (rc) return rc;
    GGParams p{};
    p.in = in; p.W = Wuse; p.tab = tab; p.orow = orow; p.out = out;
    p.n_rows = n_out; p.Cin = Cin; p.Cout = Cout; p.K = K;
    p.accumulate = accumulate; p.pairs_mode = 0;
    note_kernel("k_gather_gemm");
    rc = dispatch_gg(p, n_out, 1, st);
    if (rc || !res) return rc;
    const int64_t n = n_out * Cout;  // the CUDA-core fallback adds the residual in a second pass
    k_add_rows<<<(unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)num_sms() * 8), 256, 0, st>>>(out, res, n);
    B200SP_LAUNCH_CHECK();
    return B200SP_OK;
}
}  // namespace b200sp

extern "C" int b200sp_gather_gemm(const float* in, int64_t n_in, int Cin, const float* W, int wflags, const int32_t* tab,
                                  const int32_t* orow, const int32_t* rowmask, int K, float* out, int64_t n_out, int Cout,
                                  int accumulate, void* ws, int64_t ws_bytes, void* stream) {
    return gather_gemm_impl(in, n_in, Cin, W, wflags, tab, orow, rowmask, K, out, n_out, Cout, accumulate, nullptr, ws, ws_bytes,
                            stream);
}

extern "C" int b200sp_gather_gemm_res(const float* in, int64_t n_in, int Cin, const float* W, int wflags, const int32_t* tab,
                                      const int32_t* orow, const int32_t* rowmask, int K, float* out, int64_t n_out, int Cout,
                                      const float* res, void* ws, int64_t ws_bytes, void* stream) {
    return gather_gemm_impl(in, n_in, Cin, W, wflags, tab, orow, rowmask, K, out, n_out, Cout, 0, res, ws, ws_bytes, stream);
}

extern "C" int b200sp_gather_gemm_pairs(const float* in, int Cin, const float* W, int wflags, const int32_t* pin,
                                        const int32_t* pout, const int32_t* pairnum_dev, int64_t n_upper, int K,
                                        int64_t pstride, float* out, int Cout, int accumulate, void* ws,
                                        int64_t ws_bytes, void* stream) {
    B200SP_CHECK_ARG(Cin >= 1 && Cout >= 1 && K >= 1 && K <= GG_MAXK, "gather_gemm_pairs: bad Cin/Cout/K");
    B200SP_CHECK_ARG(pin && pout && pairnum_dev, "gather_gemm_pairs: null pair lists");
    if (n_upper <= 0) return B200SP_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (conv_impl() == 0) {
        const int Ci_w = (wflags & 1) ? Cout : Cin, Co_w = (wflags & 1) ? Cin : Cout;
        int rc = conv_tc_run(in, Cin, W, Ci_w, Co_w, wflags, nullptr, nullptr, nullptr, pin, pout, pairnum_dev, n_upper, pstride, K,
                             out, Cout, accumulate, 1, ws, ws_bytes, st, 0);
        if (rc != B200SP_EUNSUP) return rc;
    }
    const float* Wuse = nullptr;
    int rc = fp32_weights(W, K, Cin, Cout, wflags, ws, ws_bytes, st, &Wuse);
    if (rc) return rc;
    GGParams p{};
    p.in = in; p.W = Wuse; p.pin = pin; p.pout = pout; p.pairnum = pairnum_dev; p.out = out;
    p.pstride = pstride; p.Cin = Cin; p.Cout = Cout; p.K = K;
    p.accumulate = accumulate; p.pairs_mode = 1;
    note_kernel("k_gather_gemm");
    return dispatch_gg(p, n_upper, K, st);
}

extern "C" int b200sp_wgrad(const float* a, int Ca, const float* b, int Cb, const int32_t* pa, const int32_t* pb,
                            const int32_t* pairnum_dev, int64_t n_upper, int K, int64_t pstride, float* dW,
                            void* stream) {
    B200SP_CHECK_ARG(Ca >= 1 && Cb >= 1 && K >= 1, "wgrad: bad Ca/Cb/K");
    B200SP_CHECK_ARG(!(pa || pb) || pairnum_dev, "wgrad: pair lists need pairnum_dev");
    const int64_t nmax = n_upper;
    if (nmax <= 0) return B200SP_OK;
    if (conv_impl() == 0) {
        int rc = wgrad_tc_run(a, Ca, b, Cb, pa, pb, pairnum_dev, n_upper, K, pstride, dW, (cudaStream_t)stream);
        if (rc != B200SP_EUNSUP) return rc;
    }
    WGParams p{};
    p.a = a; p.b = b; p.pa = pa; p.pb = pb; p.pairnum = (pa || pb) ? pairnum_dev : nullptr; p.dW = dW;
    p.pstride = pstride; p.n_rows = n_upper; p.Ca = Ca; p.Cb = Cb; p.K = K;
    auto tile = [](int C) { int t = (C + 3) / 4 * 4; return t > 64 ? 64 : t; };
    p.ta = tile(Ca); p.tb = tile(Cb);
    // enough blocks to fill the machine, few enough that the final atomics stay cheap
    int tiles = (int)(cdiv(Ca, p.ta) * cdiv(Cb, p.tb));
    int64_t ppb = 4096;
    while (ppb > 256 && cdiv(nmax, ppb) * K * tiles < 2 * 148) ppb >>= 1;
    p.ppb = (int)ppb;
    note_kernel("k_wgrad");
    dim3 grid((unsigned)cdiv(nmax, ppb), K, tiles);
    k_wgrad<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
    B200SP_LAUNCH_CHECK();
    return B200SP_OK;
}

// which table-form kernel takes a layer: 1 = register-gather (wgrad_direct.cu: Ca, Cb in {16, 32}, not both 32),
// 2 = out-stationary tcgen05 (wgrad_os.cu: channels % 8 == 0 up to 256), 0 = neither (the pair-list kernels).
// n_rows < 0: shape only.  Measured (tools/dev_wgrad_os.py, us): 300 k rows 16x16 direct 76 / os 118; 32x16 176 / 160;
// 16x32 167 / 133; 118 k rows 32x32 os 72 / pair-list 170; 64x32 os ~150 / pair-list 259; 26.5 k rows 48x48 os 68 /
// pair-list 46; 223 rows 96x96 os 13 / pair-list 20.
static int wgrad_table_kernel(int K, int Ca, int Cb, int64_t n_rows) {
    if (conv_impl() != 0) return 0;
    B200SP_ENV_INT(env_os, "B200SP_WGRAD_OS", 1);
    B200SP_ENV_INT(env_lo, "B200SP_WGOS_MIN_ROWS", 50000);
    B200SP_ENV_INT(env_tiny, "B200SP_WGOS_TINY_ROWS", 512);
    const bool direct = b200sp::wgrad_direct_covers(K, Ca, Cb);
    const bool os = env_os && b200sp::wgrad_os_covers(K, Ca, Cb);
    if (env_os == 2 && os) return 2;
    if (direct && (Ca == 16 && Cb == 16)) return 1;
    if (os) {
        const int MB = Ca <= 64 ? 1 : (Ca + 127) / 128, Npad = (Cb + 15) / 16 * 16;
        const bool pairlist_tc = MB * Npad <= 512 && MB <= 4;  // wgrad_tc_run's coverage
        if (n_rows < 0) return direct ? 1 : 2;
        if (n_rows >= env_lo || (n_rows <= env_tiny && !direct) || !pairlist_tc) return 2;
    }
    return direct ? 1 : 0;
}

extern "C" int b200sp_wgrad_table_covers(int K, int Ca, int Cb) { return wgrad_table_kernel(K, Ca, Cb, -1) ? 1 : 0; }

// Is the table form also the FASTER one for n_rows rows?  (the dispatch rule of the layer executor)
extern "C" int b200sp_wgrad_table_prefers(int K, int Ca, int Cb, int64_t n_rows) {
    return wgrad_table_kernel(K, Ca, Cb, n_rows < 0 ? 0 : n_rows) ? 1 : 0;
}

extern "C" int b200sp_wgrad_table(const float* a, int Ca, const float* g, int Cb, const int32_t* tab, const int32_t* orow,
                                  const int32_t* rowmask, int64_t n_rows, int K, float* dW, void* stream) {
    B200SP_CHECK_ARG(Ca >= 1 && Cb >= 1 && K >= 1 && n_rows >= 0, "wgrad_table: bad sizes");
    B200SP_CHECK_ARG(tab || K == 1, "wgrad_table: tab == NULL requires K == 1");
    int which = wgrad_table_kernel(K, Ca, Cb, n_rows);
    if (!which) which = wgrad_table_kernel(K, Ca, Cb, -1);  // covered, just not preferred at this size: still runs
    B200SP_CHECK_ARG(which, "wgrad_table: shape K=%d %dx%d is not covered (ask b200sp_wgrad_table_covers; use b200sp_wgrad)", K, Ca, Cb);
    if (which == 2) return wgrad_os_run(a, Ca, g, Cb, tab, orow, rowmask, n_rows, K, dW, (cudaStream_t)stream);
    return wgrad_direct_run(a, Ca, g, Cb, tab, orow, rowmask, n_rows, K, dW, (cudaStream_t)stream);
}

extern "C" int b200sp_weight_transpose(const float* W, int K, int Cin, int Cout, int mirror, float* out,
                                       void* stream) {
    B200SP_CHECK_ARG(K >= 1 && Cin >= 1 && Cout >= 1, "weight_transpose: bad sizes");
    int64_t n = (int64_t)K * Cin * Cout;
    k_weight_transpose<<<(unsigned)std::min<int64_t>(cdiv(n, 256), 2048), 256, 0, (cudaStream_t)stream>>>(W, K, Cin, Cout,
                                                                                                   mirror, out);
    B200SP_LAUNCH_CHECK();
    return B200SP_OK;
}
