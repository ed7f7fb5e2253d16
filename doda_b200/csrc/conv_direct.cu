// conv_direct.cu — register-gather sparse conv for the narrow layers (Cin, Cout in {16, 32}; built up to 64).
//
// The top levels of DODA's U-Net (model/unet_block.py:62-85 with m=16: 16/32/48/64 channels on 300 k / 118 k / 27 k /
// 6 k rows) move ~70 MB per layer and do < 3 GFLOP: they are bound by how many gathered rows are in flight, not by the
// contraction.  The tcgen05 pipeline of conv_tc.cu (gather -> smem stage -> transform -> MMA -> commit, a handful of
// stages in flight per SM) is latency-bound on them.  Here every warp owns 16 consecutive rows of the (mask-sorted)
// processing order and gathers its neighbour rows STRAIGHT INTO the A fragments of a warp-level
// mma.m16n8k8 (tf32, fp32 accumulate): no shared memory, no barriers, 16-48 warps per SM each with its own loads in
// flight.  3xTF32 (a_lo*b_hi + a_hi*b_lo + a_hi*b_hi) keeps fp32-level accuracy like the tcgen05 path.
//
// Fragment mapping (g = lane / 4, t = lane % 4).  The K index of a 16-channel group is permuted so that lane t's four
// A values are the float4 at channels 4t..4t+3 of its row (one LDG.128 per row per 16 channels); the N index is
// permuted so that the lane's four outputs per row are channels 4t..4t+3 (one STG.128), and so that the two B values it
// needs for an n-tile pair are adjacent in the raw weight tensor (one LDG.64, forward or transposed) -- the kernel
// reads the raw [K][Ci_w][Co_w] weights through L1, no prepared image.
#include "common.cuh"
#include <stdlib.h>

namespace b200sp {
namespace {

struct DirectParams {
    const float* in;
    const float* W;       // raw weights [K][Ci_w][Co_w]
    const int* tab;       // [n_rows][K] input row per (processing row, offset), -1 = none; nullptr: K == 1, identity
    const int* orow;      // [n_rows] output row of each processing row, nullptr = identity
    const int* rowmask;   // [n_rows] bit k set <=> tab[r][k] >= 0, nullptr = not available
    float* out;
    long long n_rows;
    int K, Cin, Cout;
    int transposed;       // Weff[k][ci][co] = W[k'][co][ci]
    int mirror;           // k' = K-1-k
    int accumulate;
    const float* res;     // added to the result: out itself when accumulating, a residual branch, or nullptr
    int total_warps;
};

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float4 ldg_stream4(const float* p) {  // gathered rows: do not displace the weights in L1
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];\n"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ void split(float v, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(v) & 0xFFFFE000u;
    lo = __float_as_uint(v - __uint_as_float(hi));
}

// rows per warp = 16 * MT: the B fragments of an offset (loaded through L1 and split into hi/lo) are shared by MT
// m-tiles -- at MT = 1 those loads and splits, not the MMAs or the gathers, take most of the issue slots
__host__ __device__ constexpr int direct_mt(int CG, int NG) { return 1; }

// CG = Cin / 16, NG = 16-column output groups per warp (the warp's columns start at blockIdx.y * NG * 16)
template <int CG, int NG, bool TRANSPOSED>
__global__ void __launch_bounds__(256) k_conv_direct(const DirectParams p) {
    constexpr int MT = direct_mt(CG, NG);
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int col0 = blockIdx.y * NG * 16;
    const int K = p.K, Cin = p.Cin, Cout = p.Cout;
    const long long n_tiles = (p.n_rows + 16 * MT - 1) / (16 * MT);
    const int Ci_w = TRANSPOSED ? Cout : Cin, Co_w = TRANSPOSED ? Cin : Cout;
    const size_t wk_stride = (size_t)Ci_w * Co_w;

    for (long long tile = warp; tile < n_tiles; tile += p.total_warps) {
        const long long row0 = tile * (16 * MT);
        unsigned mask;
        if (p.rowmask) {
            unsigned m = 0;
#pragma unroll
            for (int i = 0; i < (16 * MT + 31) / 32; ++i) {
                const int r = i * 32 + lane;
                if (r < 16 * MT && row0 + r < p.n_rows) m |= (unsigned)__ldg(p.rowmask + row0 + r);
            }
            mask = __reduce_or_sync(0xFFFFFFFFu, m);
        } else {
            mask = K >= 32 ? 0xFFFFFFFFu : ((1u << K) - 1u);
        }
        float acc[MT][NG][2][4];
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
            for (int G = 0; G < NG; ++G)
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc[m][G][h][i] = 0.f;

        while (mask) {
            const int k = __ffs(mask) - 1;
            mask &= mask - 1;
            int ia[MT], ib[MT];
            bool any = false;
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                const long long ra = row0 + m * 16 + g, rb = ra + 8;
                ia[m] = ra < p.n_rows ? (p.tab ? __ldg(p.tab + ra * K + k) : (int)ra) : -1;
                ib[m] = rb < p.n_rows ? (p.tab ? __ldg(p.tab + rb * K + k) : (int)rb) : -1;
                any |= (ia[m] >= 0) | (ib[m] >= 0);
            }
            if (!p.rowmask && !__any_sync(0xFFFFFFFFu, any)) continue;
            float4 xa[MT][CG], xb[MT][CG];
#pragma unroll
            for (int m = 0; m < MT; ++m)
#pragma unroll
                for (int c = 0; c < CG; ++c) {
                    xa[m][c] = ia[m] >= 0 ? ldg_stream4(p.in + (size_t)ia[m] * Cin + c * 16 + 4 * t) : make_float4(0.f, 0.f, 0.f, 0.f);
                    xb[m][c] = ib[m] >= 0 ? ldg_stream4(p.in + (size_t)ib[m] * Cin + c * 16 + 4 * t) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            const float* Wk = p.W + (size_t)(p.mirror ? K - 1 - k : k) * wk_stride;
#pragma unroll
            for (int c = 0; c < CG; ++c) {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    // k-chunk j of channel group c: logical k = t <-> channel 4t+2j, logical k = t+4 <-> channel 4t+2j+1
                    const int ci = c * 16 + 4 * t + 2 * j;  // channel of b0; b1 is ci + 1
                    uint32_t bh[NG][4], bl[NG][4];           // [n-tile 0: b0, b1; n-tile 1: b0, b1]
#pragma unroll
                    for (int G = 0; G < NG; ++G) {
                        const int co = col0 + G * 16 + 2 * g;  // n-tile 0 <-> column co, n-tile 1 <-> column co + 1
                        float b00, b01, b10, b11;
                        if (TRANSPOSED) {                       // Weff[ci][co] = W[co][ci]: (ci, ci+1) adjacent
                            const float2 w0 = __ldg((const float2*)(Wk + (size_t)co * Co_w + ci));
                            const float2 w1 = __ldg((const float2*)(Wk + (size_t)(co + 1) * Co_w + ci));
                            b00 = w0.x; b01 = w0.y; b10 = w1.x; b11 = w1.y;
                        } else {                                // Weff[ci][co] = W[ci][co]: (co, co+1) adjacent
                            const float2 w0 = __ldg((const float2*)(Wk + (size_t)ci * Co_w + co));
                            const float2 w1 = __ldg((const float2*)(Wk + (size_t)(ci + 1) * Co_w + co));
                            b00 = w0.x; b10 = w0.y; b01 = w1.x; b11 = w1.y;
                        }
                        split(b00, bh[G][0], bl[G][0]); split(b01, bh[G][1], bl[G][1]);
                        split(b10, bh[G][2], bl[G][2]); split(b11, bh[G][3], bl[G][3]);
                    }
#pragma unroll
                    for (int m = 0; m < MT; ++m) {
                        uint32_t ahi[4], alo[4];
                        split(j ? xa[m][c].z : xa[m][c].x, ahi[0], alo[0]);
                        split(j ? xb[m][c].z : xb[m][c].x, ahi[1], alo[1]);
                        split(j ? xa[m][c].w : xa[m][c].y, ahi[2], alo[2]);
                        split(j ? xb[m][c].w : xb[m][c].y, ahi[3], alo[3]);
#pragma unroll
                        for (int G = 0; G < NG; ++G) {
                            mma_tf32(acc[m][G][0], alo, bh[G][0], bh[G][1]);
                            mma_tf32(acc[m][G][1], alo, bh[G][2], bh[G][3]);
                            mma_tf32(acc[m][G][0], ahi, bl[G][0], bl[G][1]);
                            mma_tf32(acc[m][G][1], ahi, bl[G][2], bl[G][3]);
                            mma_tf32(acc[m][G][0], ahi, bh[G][0], bh[G][1]);
                            mma_tf32(acc[m][G][1], ahi, bh[G][2], bh[G][3]);
                        }
                    }
                }
            }
        }
        // c-fragment: c0/c1 = row g, logical cols 2t / 2t+1; c2/c3 = row g+8.  n-tile 0 logical col n <-> column 2n,
        // n-tile 1 <-> 2n+1  =>  the lane owns columns 4t..4t+3 = (nt0.c0, nt1.c0, nt0.c1, nt1.c1)
#pragma unroll
        for (int m = 0; m < MT; ++m) {
            const long long ra = row0 + m * 16 + g, rb = ra + 8;
            const bool va_ok = ra < p.n_rows, vb_ok = rb < p.n_rows;
            long long oa = ra, ob = rb;
            if (p.orow) {
                if (va_ok) oa = __ldg(p.orow + ra);
                if (vb_ok) ob = __ldg(p.orow + rb);
            }
#pragma unroll
            for (int G = 0; G < NG; ++G) {
                const int c = col0 + G * 16 + 4 * t;
                if (va_ok) {
                    float4 v = make_float4(acc[m][G][0][0], acc[m][G][1][0], acc[m][G][0][1], acc[m][G][1][1]);
                    float4* dst = (float4*)(p.out + (size_t)oa * Cout + c);
                    if (p.res) { const float4 o = *(const float4*)(p.res + (size_t)oa * Cout + c); v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
                    *dst = v;
                }
                if (vb_ok) {
                    float4 v = make_float4(acc[m][G][0][2], acc[m][G][1][2], acc[m][G][0][3], acc[m][G][1][3]);
                    float4* dst = (float4*)(p.out + (size_t)ob * Cout + c);
                    if (p.res) { const float4 o = *(const float4*)(p.res + (size_t)ob * Cout + c); v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
                    *dst = v;
                }
            }
        }
    }
}

int g_direct = -1;  // 1 = use the register-gather kernel where it applies (default), 0 = never (B200SP_DIRECT=0)

template <int CG, int NG>
int launch(const DirectParams& p, dim3 grid, cudaStream_t st) {
    if (p.transposed)
        B200SP_CUDA(launch_pdl(k_conv_direct<CG, NG, true>, grid, dim3(256), 0, st, p));
    else
        B200SP_CUDA(launch_pdl(k_conv_direct<CG, NG, false>, grid, dim3(256), 0, st, p));
    B200SP_LAUNCH_CHECK();
    return B200SP_OK;
}
template <int CG>
int launch_ng(int ng, const DirectParams& p, dim3 grid, cudaStream_t st) {
    switch (ng) {
        case 1: return launch<CG, 1>(p, grid, st);
        case 2: return launch<CG, 2>(p, grid, st);
        default: return launch<CG, 3>(p, grid, st);
    }
}

}  // namespace

bool conv_direct_enabled() {
    if (g_direct < 0) {
        const char* e = getenv("B200SP_DIRECT");
        g_direct = (e && e[0] == '0') ? 0 : 1;
    }
    return g_direct != 0;
}
void conv_direct_set(int on) { g_direct = on ? 1 : 0; }

bool conv_direct_covers(int K, int Cin, int Cout) {
    // measured inside the U-Net step (tools/gpu_timeline.py): 16 -> 16 on 300 k rows 56 us vs 81 us for the tcgen05
    // kernel, 32 -> 32 on 118 k rows 89 vs 99 us, 48 -> 48 on 27 k rows 75 vs 75 us, 64 -> 64 on 6 k rows 64-92 vs
    // 26 us: the legacy warp-level MMA runs ~20 cycles per m16n8k8 per SM sub-partition, so from 48 channels on the
    // contraction, not the gather, is the bound and tcgen05 wins.  B200SP_DIRECT_MAXC=64 widens it for experiments.
    static int maxc = -1;
    if (maxc < 0) {
        const char* e = getenv("B200SP_DIRECT_MAXC");
        maxc = e ? atoi(e) : 32;
    }
    auto ok = [&](int c) { return (c == 16 || c == 32 || c == 48 || c == 64) && c <= maxc; };
    return conv_direct_enabled() && K >= 1 && K <= 32 && ok(Cin) && ok(Cout);
}

// table mode only (tab [n_rows][K] or identity); B200SP_EUNSUP when the shape is not covered
int conv_direct_run(const float* in, int Cin, const float* W, int wflags, const int* tab, const int* orow,
                    const int* rowmask, long long n_rows, int K, float* out, int Cout, int accumulate, cudaStream_t st, const float* res) {
    if (!conv_direct_covers(K, Cin, Cout) || (wflags & 4)) return B200SP_EUNSUP;
    if (!tab && K != 1) return B200SP_EUNSUP;
    B200SP_CHECK_ARG((((uintptr_t)in | (uintptr_t)out | (uintptr_t)W) & 15) == 0, "conv_direct: pointers must be 16-byte aligned");
    DirectParams p{};
    note_kernel("k_conv_direct");
    p.in = in; p.W = W; p.tab = tab; p.orow = orow; p.rowmask = rowmask; p.out = out;
    p.n_rows = n_rows; p.K = K; p.Cin = Cin; p.Cout = Cout;
    p.transposed = wflags & 1; p.mirror = (wflags >> 1) & 1; p.accumulate = accumulate;
    p.res = res ? res : (accumulate ? out : nullptr);
    B200SP_CHECK_ARG(((uintptr_t)p.res & 15) == 0, "conv_direct: residual pointer must be 16-byte aligned");
    const int groups = Cout / 16;
    const int ng = groups == 4 ? 2 : groups;  // 64 columns = two warps of 32
    const int ysplit = groups / ng;
    const int rows_per_warp = 16 * direct_mt(Cin / 16, ng);
    const long long n_tiles = (n_rows + rows_per_warp - 1) / rows_per_warp;
    const int sms = num_sms();
    long long blocks = (n_tiles + 7) / 8;
    const long long cap = (long long)sms * 6;
    if (blocks > cap) blocks = cap;
    p.total_warps = (int)blocks * 8;
    dim3 grid((unsigned)blocks, (unsigned)ysplit);
    switch (Cin / 16) {
        case 1: return launch_ng<1>(ng, p, grid, st);
        case 2: return launch_ng<2>(ng, p, grid, st);
        case 3: return launch_ng<3>(ng, p, grid, st);
        default: return launch_ng<4>(ng, p, grid, st);
    }
}

}  // namespace b200sp
