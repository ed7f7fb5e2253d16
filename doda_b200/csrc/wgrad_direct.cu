// wgrad_direct.cu — out-stationary weight gradient for the narrow layers (Ca, Cb in {16, 32}), table form:
//     dW[k][ca][cb] += sum_r a[tab[r][k]][ca] * g[orow[r]][cb]
// Same idea as conv_direct.cu: a warp owns 16 consecutive rows of the (mask-sorted) processing order, loads its 16
// rows of g ONCE as B fragments (hi/lo split kept in registers), and for every offset present in the tile gathers the
// 16 neighbour rows of a straight into A fragments (transposed: M = channels of a, K = rows) of a warp-level
// mma.m16n8k8 3xTF32.  The 16-row partial product is added to a per-block accumulator in shared memory laid out
// [offset][fragment register][lane] (conflict-free), and the block adds its accumulator to dW once at the end with
// vector atomics.  Compared with the pair-list kernels (wgrad_tc.cu) g is read once per row instead of once per pair
// and nothing is staged through shared memory.
#include "common.cuh"
#include <stdlib.h>

namespace b200sp {
namespace {

struct WDParams {
    const float* a;      // [*, Ca]   gathered operand (layer input)
    const float* g;      // [*, Cb]   stationary operand (gradient of the layer output)
    const int* tab;      // [n_rows][K] row of a per (processing row, offset), -1 = none; nullptr: K == 1, identity
    const int* orow;     // [n_rows] row of g per processing row, nullptr = identity
    const int* rowmask;  // [n_rows] bit k set <=> tab[r][k] >= 0, nullptr = unknown
    float* dW;           // [K][Ca][Cb], accumulated into
    long long n_rows;
    int K, Ca, Cb;
    int total_warps;
};

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void split(float v, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(v) & 0xFFFFE000u;
    lo = __float_as_uint(v - __uint_as_float(hi));
}
__device__ __forceinline__ void red_add_v4(float* addr, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}
// 2 * W consecutive floats (W = 1: float2, W = 2: float4) of a row, streamed past L1
template <int W>
__device__ __forceinline__ void ld_chunk(const float* p, float (&v)[2 * W]) {
    if constexpr (W == 1) {
        asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];\n" : "=f"(v[0]), "=f"(v[1]) : "l"(p));
    } else {
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];\n"
                     : "=f"(v[0]), "=f"(v[1]), "=f"(v[2 * W - 2]), "=f"(v[2 * W - 1])
                     : "l"(p));
    }
}

// CA = Ca / 16 (m-tiles), CB = Cb / 16 (pairs of n-tiles).  Channel maps (g = lane / 4, t = lane % 4):
//   m-tile mt, m = g + 8h   <-> channel of a: 2*CA*g + 2*mt + h         (the lane's chunk of a row: 2*CA floats at 2*CA*g)
//   n-tile j,  n            <-> channel of g: 2*CB*n + j                 (the lane's chunk of a g row: 2*CB floats at 2*CB*g)
template <int CA, int CB>
__global__ void __launch_bounds__(256) k_wgrad_direct(const WDParams p) {
    constexpr int NT = 2 * CB;            // n-tiles
    constexpr int REGS = CA * NT * 4;     // accumulator registers per lane per offset
    extern __shared__ float acc_s[];      // [K][REGS][32]
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int K = p.K, Ca = p.Ca, Cb = p.Cb;
    for (int i = threadIdx.x; i < K * REGS * 32; i += blockDim.x) acc_s[i] = 0.f;
    __syncthreads();
    pdl_wait();  // the accumulator clear above overlaps the predecessor's tail
    const long long n_tiles = (p.n_rows + 15) >> 4;

    for (long long tile = warp; tile < n_tiles; tile += p.total_warps) {
        const long long row0 = tile * 16;
        long long r[4];
        bool ok[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            r[q] = row0 + t + 4 * q;
            ok[q] = r[q] < p.n_rows;
        }
        unsigned mask;
        if (p.rowmask) {
            unsigned m = 0;
            if (lane < 16 && row0 + lane < p.n_rows) m = (unsigned)__ldg(p.rowmask + row0 + lane);
            mask = __reduce_or_sync(0xFFFFFFFFu, m);
        } else {
            mask = K >= 32 ? 0xFFFFFFFFu : ((1u << K) - 1u);
        }
        // the tile's 16 rows of g as B fragments: chunk c (rows 8c..8c+7): b0 = row 8c+t, b1 = row 8c+t+4
        uint32_t bh[2][NT][2], bl[2][NT][2];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float v[2 * CB];
#pragma unroll
            for (int j = 0; j < 2 * CB; ++j) v[j] = 0.f;
            if (ok[q]) {
                const long long o = p.orow ? (long long)__ldg(p.orow + r[q]) : r[q];
                ld_chunk<CB>(p.g + o * Cb + 2 * CB * g, v);
            }
#pragma unroll
            for (int j = 0; j < NT; ++j) split(v[j], bh[q >> 1][j][q & 1], bl[q >> 1][j][q & 1]);
        }
        while (mask) {
            const int k = __ffs(mask) - 1;
            mask &= mask - 1;
            int idx[4];
            bool any = false;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                idx[q] = ok[q] ? (p.tab ? __ldg(p.tab + r[q] * K + k) : (int)r[q]) : -1;
                any |= idx[q] >= 0;
            }
            if (!p.rowmask && !__any_sync(0xFFFFFFFFu, any)) continue;
            float av[4][2 * CA];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
#pragma unroll
                for (int j = 0; j < 2 * CA; ++j) av[q][j] = 0.f;
                if (idx[q] >= 0) ld_chunk<CA>(p.a + (size_t)idx[q] * Ca + 2 * CA * g, av[q]);
            }
            float acc[CA][NT][4];
#pragma unroll
            for (int mt = 0; mt < CA; ++mt)
#pragma unroll
                for (int j = 0; j < NT; ++j)
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc[mt][j][i] = 0.f;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
#pragma unroll
                for (int mt = 0; mt < CA; ++mt) {
                    uint32_t ahi[4], alo[4];
                    split(av[2 * c][2 * mt], ahi[0], alo[0]);          // m = g,     k = t
                    split(av[2 * c][2 * mt + 1], ahi[1], alo[1]);      // m = g + 8, k = t
                    split(av[2 * c + 1][2 * mt], ahi[2], alo[2]);      // m = g,     k = t + 4
                    split(av[2 * c + 1][2 * mt + 1], ahi[3], alo[3]);  // m = g + 8, k = t + 4
#pragma unroll
                    for (int j = 0; j < NT; ++j) {
                        mma_tf32(acc[mt][j], alo, bh[c][j][0], bh[c][j][1]);
                        mma_tf32(acc[mt][j], ahi, bl[c][j][0], bl[c][j][1]);
                        mma_tf32(acc[mt][j], ahi, bh[c][j][0], bh[c][j][1]);
                    }
                }
            }
            float* s = acc_s + (size_t)k * (REGS * 32) + lane;
#pragma unroll
            for (int mt = 0; mt < CA; ++mt)
#pragma unroll
                for (int j = 0; j < NT; ++j)
#pragma unroll
                    for (int i = 0; i < 4; ++i) atomicAdd(s + ((mt * NT + j) * 4 + i) * 32, acc[mt][j][i]);
        }
    }
    __syncthreads();
    pdl_trigger();  // late trigger (the block holds up to 55 KB of shared memory)
    // c-fragment: i = 2h + e  <->  m = g + 8h, n = 2t + e.  The lane owns, for each of its 2*CA channels of a, the
    // 4*CB consecutive channels of g starting at 4*CB*t.
    for (int idx = threadIdx.x; idx < K * 32; idx += blockDim.x) {
        const int k = idx >> 5, ln = idx & 31, gg = ln >> 2, tt = ln & 3;
        const float* s = acc_s + (size_t)k * (REGS * 32) + ln;
#pragma unroll
        for (int mt = 0; mt < CA; ++mt)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int ca = 2 * CA * gg + 2 * mt + h;
                float* dst = p.dW + ((size_t)k * Ca + ca) * Cb + 4 * CB * tt;
                auto R = [&](int j, int i) { return s[((mt * NT + j) * 4 + i) * 32]; };
                if constexpr (CB == 1) {  // n = 2t + e  ->  channels 4t + 2e + j
                    red_add_v4(dst, make_float4(R(0, 2 * h), R(1, 2 * h), R(0, 2 * h + 1), R(1, 2 * h + 1)));
                } else {                  // n = 2t + e  ->  channels 8t + 4e + j
#pragma unroll
                    for (int e = 0; e < 2; ++e)
                        red_add_v4(dst + 4 * e, make_float4(R(0, 2 * h + e), R(1, 2 * h + e), R(2, 2 * h + e), R(3, 2 * h + e)));
                }
            }
    }
}

template <int CA, int CB>
int launch(const WDParams& p0, cudaStream_t st) {
    WDParams p = p0;
    const size_t smem = (size_t)p.K * CA * 2 * CB * 4 * 32 * sizeof(float);
    static bool attr_done = false;
    if (!attr_done) {
        B200SP_CUDA(cudaFuncSetAttribute(k_wgrad_direct<CA, CB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_done = true;
    }
    int per_sm = (int)((220 * 1024) / (smem + 1024));
    if (per_sm > 4) per_sm = 4;
    if (per_sm < 1) return B200SP_EUNSUP;
    const long long n_tiles = (p.n_rows + 15) / 16;
    long long blocks = (n_tiles + 7) / 8;
    const long long cap = (long long)num_sms() * per_sm;
    if (blocks > cap) blocks = cap;
    p.total_warps = (int)blocks * 8;
    B200SP_CUDA(launch_pdl(k_wgrad_direct<CA, CB>, dim3((unsigned)blocks), dim3(256), smem, st, p));
    B200SP_LAUNCH_CHECK();
    return B200SP_OK;
}

}  // namespace

bool conv_direct_enabled();

bool wgrad_direct_covers(int K, int Ca, int Cb) {
    // 32 x 32 is built but not used: 48 MMAs per 16-row tile-offset and 116 registers (two blocks per SM) make it
    // 205 us inside the U-Net step on 118 k rows against 182 us for the tcgen05 pair-list kernel; 16 x 16 is 86 vs
    // 200 us, 32 x 16 160 vs 290 us, 16 x 32 70 vs 140 us (tools/gpu_timeline.py, tools/dev_wgrad_direct.py)
    B200SP_ENV_INT(env_all, "B200SP_WGRAD_DIRECT_ALL", 0);
    auto ok = [](int c) { return c == 16 || c == 32; };
    return conv_direct_enabled() && K >= 1 && K <= 32 && ok(Ca) && ok(Cb) && (env_all || Ca == 16 || Cb == 16);
}

int wgrad_direct_run(const float* a, int Ca, const float* g, int Cb, const int* tab, const int* orow, const int* rowmask,
                     long long n_rows, int K, float* dW, cudaStream_t st) {
    if (!wgrad_direct_covers(K, Ca, Cb)) return B200SP_EUNSUP;
    if (!tab && K != 1) return B200SP_EUNSUP;
    B200SP_CHECK_ARG((((uintptr_t)a | (uintptr_t)g | (uintptr_t)dW) & 15) == 0, "wgrad_direct: pointers must be 16-byte aligned");
    if (n_rows <= 0) return B200SP_OK;
    WDParams p{};
    note_kernel("k_wgrad_direct");
    p.a = a; p.g = g; p.tab = tab; p.orow = orow; p.rowmask = rowmask; p.dW = dW;
    p.n_rows = n_rows; p.K = K; p.Ca = Ca; p.Cb = Cb;
    if (Ca == 16) return Cb == 16 ? launch<1, 1>(p, st) : launch<1, 2>(p, st);
    return Cb == 16 ? launch<2, 1>(p, st) : launch<2, 2>(p, st);
}

}  // namespace b200sp
