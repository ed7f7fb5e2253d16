// wgrad_os.cu — OUT-STATIONARY sparse-conv weight gradient on the Blackwell tensor cores (tcgen05 + TMEM), table form.
//
// Replaces the wgrad half of spconv v1.2 `indice_conv_backward` (per offset: two gather kernels -> cuBLAS SGEMM over the
// pair index) for every layer from 32 channels up:
//
//     dW[k][ca][cb] += sum_r  a[tab[r][k]][ca] * g[orow[r]][cb]          r = rows of the (mask-sorted) processing order
//
// The pair-list kernel (wgrad_tc.cu) gathers BOTH operands once per pair (8.9x the compulsory bytes through L2, 3 % of
// the HBM roofline in round 1).  Here a persistent CTA walks tiles of 64 processing rows and
//   * loads the tile's 64 rows of g ONCE (gathered through `orow`), converts them to bf16 hi/lo and keeps them in shared
//     memory as the MN-major B operand for every offset of the tile;
//   * per GROUP of offsets present in the tile (rowmask) gathers the neighbour rows of `a` -- `opm` offsets side by side
//     along the M dimension of one UMMA (Ca = 32: 4 offsets x 32 channels = 128 rows of M), so one
//     tcgen05.mma M=128 x N=Cb x K=16 serves up to 8 offsets;
//   * keeps the dW accumulators of ALL its offset groups resident in TMEM across all its tiles ((group, M-block) ->
//     Npad columns) and adds them to dW once at the end with vector reductions (one flush per CTA, not per tile).
// Offsets that do not fit the 512 TMEM columns are dealt to `npass` CTAs per tile range (blockIdx.y), which also gives
// the deep levels (a handful of tiles) enough CTAs.
//
// Roles (13 warps): warp 0 tile metadata (table rows by bulk copy, offset-group list from the row masks), warps 1-4 MMA
// issuers (one tcgen05.mma issue costs ~215 cycles per issuing thread on this part whatever its shape, so stages are
// dealt round-robin to `ni` issuers, each with its own accumulator set, summed in the epilogue), warps 5-12 two loader
// teams: LDG.128 x 16 in flight per thread straight to registers, bf16 hi/lo split (v = hi + lo, products
// hi*hi + lo*hi + hi*lo: bf16x3, ~5e-6 relative), STS into the canonical un-swizzled MN-major layout (core matrix =
// 8 rows x 16 B; the channel-chunk stride is padded to 144 B so that a quarter-warp's stores hit 32 distinct banks).
#include "common.cuh"
#include "tc_common.cuh"
#include <cuda_bf16.h>
#include <algorithm>
#include <stdlib.h>

namespace b200sp {

using namespace tc;

namespace {

constexpr int OS_TR = 64;           // rows per tile
constexpr int OS_MAXK = 32;
constexpr int OS_MAX_ISSUERS = 4;
constexpr int OS_TEAM_WARPS = 4;    // loader team = 128 threads
constexpr int OS_TEAM = OS_TEAM_WARPS * 32;
constexpr int OS_W_META = 0;
constexpr int OS_W_MMA = 1;         // warps 1..4
constexpr int OS_W_LOAD = 5;        // warps 5..20: up to 4 loader teams (team t: warps 5+4t..8+4t); team 0 also runs the epilogue
constexpr int OS_MAX_TEAMS = 4;
constexpr int OS_THREADS = (1 + OS_MAX_ISSUERS + OS_MAX_TEAMS * OS_TEAM_WARPS) * 32;
constexpr uint32_t OS_SBO = 144;    // bytes between 8-channel chunks (128 + 16 pad: conflict-free stores)

struct OSParams {
    const float* a;
    const float* g;
    const int* tab;      // [n_rows][K] or NULL (K == 1: row r of a)
    const int* orow;     // [n_rows] or NULL
    const int* rowmask;  // [n_rows] or NULL
    float* dW;           // [K][Ca][Cb]
    long long n_rows;
    int K, Ca, Cb;
    int ca8;             // Ca / 8
    int cb8;             // Npad / 8 (chunks of the B tile), cbv = chunks that hold real channels
    int cbv;
    int opm;             // offsets side by side along M (MB == 1)
    int MB;              // 128-row M blocks per offset (Ca > 128)
    int CH;              // 16 * MB chunks per row of the A tile
    int Npad;
    int G;               // offset groups = ceil(K / opm)
    int gpp;             // groups per pass (accumulators resident per CTA and issuer)
    int ni, S;
    int nteams;          // loader teams in use: stage s belongs to team s % nteams (nteams divides S)
    uint32_t gsA, tileA, gsB, tileB;  // 8-row group stride / bytes of one of {hi, lo}
    uint32_t tmem_cols;
    int ntiles;
    int meta_ints;
};

// read-only path, L1 allocating: the two 16-byte halves of a thread's 32-byte chunk (and its neighbours' chunks of the
// same row) share 32-byte sectors / a 128-byte line -- without L1 every sector would be fetched from L2 twice
__device__ __forceinline__ float4 ldg_nc4(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void red_add_v4(float* addr, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// 8 fp32 -> 8 bf16 hi + 8 bf16 lo.  hi = the value TRUNCATED to bf16 (its top 16 bits: one PRMT packs two of them),
// lo = bf16_rn(v - hi) (the difference is exact in fp32; one packed cvt per pair): v = hi + lo + O(2^-15 |v|).
__device__ __forceinline__ uint32_t pack_hi(float a, float b) {  // {a.top16, b.top16} -> bf16x2 (a in the low half)
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, 0x7632;\n" : "=r"(r) : "r"(__float_as_uint(a)), "r"(__float_as_uint(b)));
    return r;
}
__device__ __forceinline__ uint32_t pack_lo(float a, float b) {  // bf16x2 round-to-nearest, a in the low half
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;\n" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
__device__ __forceinline__ void split8(const float4& x, const float4& y, uint4& hi, uint4& lo) {
    const float v[8] = {x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w};
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float a = v[2 * i], b = v[2 * i + 1];
        const float ah = __uint_as_float(__float_as_uint(a) & 0xFFFF0000u), bh = __uint_as_float(__float_as_uint(b) & 0xFFFF0000u);
        h[i] = pack_hi(a, b);
        l[i] = pack_lo(a - ah, b - bh);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// smem carve-up
struct OSSmem {
    uint32_t offB, offMeta, offBars, total;
};
__host__ __device__ inline OSSmem os_layout(const OSParams& p) {
    OSSmem L;
    L.offB = (uint32_t)p.S * 2u * p.tileA;
    L.offMeta = L.offB + 2u * 2u * p.tileB;
    L.offBars = (L.offMeta + 2u * (uint32_t)p.meta_ints * 4u + 15u) & ~15u;
    L.total = L.offBars + (uint32_t)(2 * p.S + 2 + 2 + 2 + 2 + 2 + 1) * 8u + 64u;
    return L;
}

__global__ void __launch_bounds__(OS_THREADS, 1) k_wgrad_os(const OSParams p) {  // 672 threads: <= 96 registers
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const OSSmem L = os_layout(p);
    const int S = p.S, K = p.K, ni = p.ni;
    unsigned char* sA = smem;
    unsigned char* sB = smem + L.offB;
    int* s_meta = reinterpret_cast<int*>(smem + L.offMeta);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + L.offBars);
    uint64_t* empty = full + S;
    uint64_t* bfull = empty + S;    // [2] g tile landed
    uint64_t* bfree = bfull + 2;    // [2] every issuer is done with the g tile
    uint64_t* tready = bfree + 2;   // [2] tile metadata published
    uint64_t* tfree = tready + 2;   // [2] every reader is done with the metadata
    uint64_t* tload = tfree + 2;    // [2] bulk copy of the table rows landed
    uint64_t* accdone = tload + 2;  // all MMAs of the CTA complete
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(accdone + 1);
    uint32_t* s_touched = s_tmem + 1;  // [OS_MAX_ISSUERS] groups each issuer accumulated into

    const int pass = blockIdx.y;
    const int g_first = pass * p.gpp;
    const int g_count = min(p.gpp, p.G - g_first);  // groups of this pass
    const int ntiles = p.ntiles > (int)blockIdx.x ? (p.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int KT = p.tab ? K : 1;

#define OS_IDX(b) (s_meta + (b) * p.meta_ints)
#define OS_OROW(b) (OS_IDX(b) + OS_TR * KT)
#define OS_GLIST(b) (OS_OROW(b) + OS_TR)
#define OS_NG(b) (OS_GLIST(b)[OS_MAXK])
#define OS_ROWS(b) (OS_GLIST(b)[OS_MAXK + 1])

    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full[s], OS_TEAM_WARPS);
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&bfull[b], OS_TEAM_WARPS);
            mbar_init(&bfree[b], ni);
            mbar_init(&tready[b], 1);
            mbar_init(&tfree[b], p.nteams * OS_TEAM_WARPS + ni);
            mbar_init(&tload[b], 1);
        }
        mbar_init(accdone, ni);
        mbar_fence_init();
    }
    if (warp == OS_W_MMA) tmem_alloc(s_tmem, p.tmem_cols);
    if (tid < OS_MAX_ISSUERS) s_touched[tid] = 0u;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s_tmem;
    pdl_wait();  // barrier init and TMEM allocation above overlap the predecessor's tail; no global access before here

    if (warp == OS_W_META) {
        // ========== tile metadata, one tile ahead ==========
        for (int i = 0; i < ntiles; ++i) {
            const int b = i & 1;
            const uint32_t ph = (uint32_t)(i >> 1) & 1u;
            mbar_wait(&tfree[b], ph ^ 1u);
            const long long tile = (long long)blockIdx.x + (long long)i * gridDim.x;
            const long long row0 = tile * OS_TR;
            const int rows = (int)min((long long)OS_TR, p.n_rows - row0);
            int* idx = OS_IDX(b);
            int* orow = OS_OROW(b);
            int* glist = OS_GLIST(b);
            unsigned mask = 0;
            for (int r = lane; r < OS_TR; r += 32) {
                orow[r] = r < rows ? (p.orow ? __ldg(p.orow + row0 + r) : (int)(row0 + r)) : -1;
                if (p.rowmask && r < rows) mask |= (unsigned)__ldg(p.rowmask + row0 + r);
            }
            if (p.tab) {
                const int* t = p.tab + row0 * K;
                const int tot = rows * K;
                const uint32_t bytes = (uint32_t)tot * 4u;
                if (rows == OS_TR && (bytes & 15u) == 0 && ((reinterpret_cast<uintptr_t>(t) & 15) == 0)) {
                    if (lane == 0) {
                        mbar_arrive_expect_tx(&tload[b], bytes);
                        bulk_g2s(idx, t, bytes, &tload[b]);
                    }
                    mbar_wait(&tload[b], ph);
                } else {
                    for (int e = lane; e < OS_TR * K; e += 32) idx[e] = e < tot ? __ldg(t + e) : -1;
                    if (lane == 0) mbar_arrive(&tload[b]);  // keep this buffer's barrier phase in step
                    __syncwarp();
                }
                if (!p.rowmask) {
                    int kk = lane % K;
                    const int step = 32 % K;
                    for (int e = lane; e < OS_TR * K; e += 32) {
                        if (idx[e] >= 0) mask |= 1u << kk;
                        kk += step;
                        if (kk >= K) kk -= K;
                    }
                }
            } else {
                mask = rows > 0 ? 1u : 0u;
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) mask |= __shfl_xor_sync(0xffffffffu, mask, o);
            if (lane == 0) {
                int ng = 0;
                const unsigned gm = p.opm >= 32 ? 0xFFFFFFFFu : ((1u << p.opm) - 1u);
                for (int gl = 0; gl < g_count; ++gl) {
                    const int g = g_first + gl;
                    if ((mask >> (g * p.opm)) & gm) glist[ng++] = gl;
                }
                OS_NG(b) = ng;
                OS_ROWS(b) = rows;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&tready[b]);
        }
    } else if (warp >= OS_W_MMA && warp < OS_W_MMA + OS_MAX_ISSUERS) {
        // ========== MMA issuers: stage s belongs to issuer s % ni, which accumulates into ITS OWN accumulator set ==========
        const int w = warp - OS_W_MMA;
        if (w < ni) {
            const uint32_t idesc = make_idesc_bf16(128, p.Npad, 1, 1);  // both operands MN-major
            const uint64_t dA0 = make_desc(smem_u32(sA), p.gsA, OS_SBO);
            const uint64_t dB0 = make_desc(smem_u32(sB), p.gsB, OS_SBO);
            const uint32_t slotA16 = (2u * p.tileA) >> 4, loA16 = p.tileA >> 4, kA16 = (2u * p.gsA) >> 4, mbA16 = (16u * OS_SBO) >> 4;
            const uint32_t bufB16 = (2u * p.tileB) >> 4, loB16 = p.tileB >> 4, kB16 = (2u * p.gsB) >> 4;
            const uint32_t cols = (uint32_t)(p.MB * p.Npad);
            const uint32_t set0 = tmem + (uint32_t)(w * p.gpp) * cols;
            uint32_t touched = 0;
            int s0 = 0, nb = 0;
            for (int i = 0; i < ntiles; ++i) {
                const int b = i & 1;
                mbar_wait(&tready[b], (uint32_t)(i >> 1) & 1u);
                const int ng = OS_NG(b);
                const int* glist = OS_GLIST(b);
                const int bb = nb & 1;
                bool mine = false;
                for (int e = 0; e < ng; ++e) {
                    const int s = s0 + e;
                    if (s % ni != w) continue;
                    if (!mine) {
                        mbar_wait(&bfull[bb], (uint32_t)(nb >> 1) & 1u);
                        mine = true;
                    }
                    const int slot = s % S;
                    mbar_wait(&full[slot], (uint32_t)(s / S) & 1u);
                    tc_fence_after();
                    const int gl = glist[e];
                    const uint64_t da = dA0 + (uint64_t)((uint32_t)slot * slotA16);
                    const uint64_t db = dB0 + (uint64_t)((uint32_t)bb * bufB16);
                    const uint32_t dcol = set0 + (uint32_t)gl * cols;
                    const uint32_t seen = (touched >> gl) & 1u;
                    if (elect_one()) {
#pragma unroll 1
                        for (int t = 0; t < OS_TR / 16; ++t) {
                            for (int mb = 0; mb < p.MB; ++mb) {
                                const uint64_t a = da + (uint32_t)t * kA16 + (uint32_t)mb * mbA16;
                                const uint64_t bq = db + (uint32_t)t * kB16;
                                const uint32_t d = dcol + (uint32_t)(mb * p.Npad);
                                mma_bf16_ss(d, a, bq, idesc, (seen | (uint32_t)(t > 0)) ? 1u : 0u);
                                mma_bf16_ss(d, a + loA16, bq, idesc, 1u);
                                mma_bf16_ss(d, a, bq + loB16, idesc, 1u);
                            }
                        }
                        mma_commit(&empty[slot]);
                    }
                    __syncwarp();
                    touched |= 1u << gl;
                }
                if (ng > 0) {
                    // the g tile may be overwritten once every issuer's MMAs on it have completed
                    if (mine) {
                        if (elect_one()) mma_commit(&bfree[bb]);
                    } else if (lane == 0) {
                        mbar_arrive(&bfree[bb]);
                    }
                    __syncwarp();
                    ++nb;
                }
                s0 += ng;
                __syncwarp();
                if (lane == 0) mbar_arrive(&tfree[b]);
            }
            if (lane == 0) s_touched[w] = touched;
            __threadfence_block();
            __syncwarp();
            if (elect_one()) mma_commit(accdone);
            __syncwarp();
        }
    } else {
        // ========== loader teams: stage s belongs to team s % nteams ==========
        const int team = (warp - OS_W_LOAD) / OS_TEAM_WARPS;
        if (team < p.nteams) {
        const int lt = tid - (OS_W_LOAD * 32 + team * OS_TEAM);  // 0..127
        const int CH = p.CH;
        const int c = lt % CH;               // chunk of the A row this thread fills: constant over the kernel
        const int rstep = OS_TEAM / CH, r_first = lt / CH;
        // which (offset slot, 8-channel chunk) of the group that is
        int j, cc;
        bool chunk_ok;
        if (p.MB == 1) {
            j = c / p.ca8;
            cc = c - j * p.ca8;
            chunk_ok = j < p.opm;
        } else {
            j = 0;
            cc = c;
            chunk_ok = c < p.ca8;
        }
        int s0 = 0, nb = 0;
        for (int i = 0; i < ntiles; ++i) {
            const int b = i & 1;
            mbar_wait(&tready[b], (uint32_t)(i >> 1) & 1u);
            const int ng = OS_NG(b);
            const int rows = OS_ROWS(b);
            const int* idx = OS_IDX(b);
            const int* orow = OS_OROW(b);
            const int* glist = OS_GLIST(b);
            const long long row0 = ((long long)blockIdx.x + (long long)i * gridDim.x) * OS_TR;
            const int bb = nb & 1;
            for (int e = 0; e < ng; ++e) {
                const int s = s0 + e;
                if (s % p.nteams != team) continue;
                const int slot = s % S;
                mbar_wait(&empty[slot], ((uint32_t)(s / S) & 1u) ^ 1u);
                if (e == 0) {
                    // ---- the tile's rows of g, once: fp32 -> bf16 hi / lo, MN-major B tile ----
                    mbar_wait(&bfree[bb], ((uint32_t)(nb >> 1) & 1u) ^ 1u);
                    unsigned char* b_hi = sB + (size_t)bb * 2u * p.tileB;
                    unsigned char* b_lo = b_hi + p.tileB;
                    const int items = OS_TR * p.cb8;
                    for (int it0 = lt; it0 < items; it0 += 4 * OS_TEAM) {
                        float4 x[4], y[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int it = it0 + u * OS_TEAM;
                            x[u] = y[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (it < items) {
                                const int r = it / p.cb8, cq = it - r * p.cb8;
                                const int o = r < rows ? orow[r] : -1;
                                if (o >= 0 && cq < p.cbv) {
                                    const float* src = p.g + (size_t)o * p.Cb + cq * 8;
                                    x[u] = ldg_nc4(src);
                                    y[u] = ldg_nc4(src + 4);
                                }
                            }
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int it = it0 + u * OS_TEAM;
                            if (it < items) {
                                const int r = it / p.cb8, cq = it - r * p.cb8;
                                uint4 hi, lo;
                                split8(x[u], y[u], hi, lo);
                                const uint32_t off = (uint32_t)(r >> 3) * p.gsB + (uint32_t)cq * OS_SBO + (uint32_t)(r & 7) * 16u;
                                *reinterpret_cast<uint4*>(b_hi + off) = hi;
                                *reinterpret_cast<uint4*>(b_lo + off) = lo;
                            }
                        }
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bfull[bb]);
                }
                // ---- the neighbour rows of a for the offsets of group glist[e] ----
                const int g = g_first + glist[e];
                const int k = p.MB == 1 ? g * p.opm + j : g;
                const bool kok = chunk_ok && k < K;
                unsigned char* a_hi = sA + (size_t)slot * 2u * p.tileA;
                unsigned char* a_lo = a_hi + p.tileA;
                const int nq = OS_TR / rstep;
                for (int q0 = 0; q0 < nq; q0 += 4) {
                    float4 x[4], y[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int r = r_first + (q0 + u) * rstep;
                        x[u] = y[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (q0 + u < nq && kok && r < rows) {
                            const int src = p.tab ? idx[r * KT + k] : (int)(row0 + r);
                            if (src >= 0) {
                                const float* sp = p.a + (size_t)src * p.Ca + cc * 8;
                                x[u] = ldg_nc4(sp);
                                y[u] = ldg_nc4(sp + 4);
                            }
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (q0 + u < nq) {
                            const int r = r_first + (q0 + u) * rstep;
                            uint4 hi, lo;
                            split8(x[u], y[u], hi, lo);
                            const uint32_t off = (uint32_t)(r >> 3) * p.gsA + (uint32_t)c * OS_SBO + (uint32_t)(r & 7) * 16u;
                            *reinterpret_cast<uint4*>(a_hi + off) = hi;
                            *reinterpret_cast<uint4*>(a_lo + off) = lo;
                        }
                    }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&full[slot]);
            }
            if (ng > 0) ++nb;
            s0 += ng;
            __syncwarp();
            if (lane == 0) mbar_arrive(&tfree[b]);
        }
        // ========== epilogue (team 0 = warps 5-8 = TMEM lane quarters 1, 2, 3, 0): one flush of the accumulators ==========
        if (team == 0) {
            const int q4 = warp & 3;
            mbar_wait(accdone, 0);
            tc_fence_after();
            uint32_t touched[OS_MAX_ISSUERS];
            uint32_t any = 0;
#pragma unroll
            for (int w = 0; w < OS_MAX_ISSUERS; ++w) {
                touched[w] = w < ni ? *reinterpret_cast<volatile uint32_t*>(&s_touched[w]) : 0u;
                any |= touched[w];
            }
            const uint32_t cols = (uint32_t)(p.MB * p.Npad);
            const int m = q4 * 32 + lane;  // accumulator row (TMEM lane) of this thread
            for (int gl = 0; gl < g_count; ++gl) {
                if (!((any >> gl) & 1u)) continue;
                const int g = g_first + gl;
                for (int mb = 0; mb < p.MB; ++mb) {
                    int k, ch;
                    bool ok;
                    if (p.MB == 1) {
                        const int cq = m >> 3, jj = cq / p.ca8;
                        ch = (cq - jj * p.ca8) * 8 + (m & 7);
                        k = g * p.opm + jj;
                        ok = jj < p.opm && k < K;
                    } else {
                        ch = mb * 128 + m;
                        k = g;
                        ok = ch < p.Ca;
                    }
                    for (int c16 = 0; c16 * 16 < p.Npad; ++c16) {
                        float v[16];
#pragma unroll
                        for (int e = 0; e < 16; ++e) v[e] = 0.f;
                        for (int w = 0; w < ni; ++w) {
                            if (!((touched[w] >> gl) & 1u)) continue;  // warp-uniform
                            float t[16];
                            tmem_ld16(tmem + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(w * p.gpp + gl) * cols + (uint32_t)(mb * p.Npad + c16 * 16), t);
#pragma unroll
                            for (int e = 0; e < 16; ++e) v[e] += t[e];
                        }
                        if (!ok) continue;
                        float* o = p.dW + ((size_t)k * p.Ca + ch) * p.Cb + c16 * 16;
#pragma unroll
                        for (int g4 = 0; g4 < 4; ++g4) {
                            if (c16 * 16 + g4 * 4 >= p.Cb) break;
                            red_add_v4(o + g4 * 4, make_float4(v[g4 * 4], v[g4 * 4 + 1], v[g4 * 4 + 2], v[g4 * 4 + 3]));
                        }
                    }
                }
            }
        }
        }  // team < nteams
    }
    pdl_trigger();  // late trigger: see conv_tc.cu
#undef OS_IDX
#undef OS_OROW
#undef OS_GLIST
#undef OS_NG
#undef OS_ROWS
    tc_fence_before();
    __syncthreads();
    if (warp == OS_W_MMA) tmem_dealloc(tmem, p.tmem_cols);
}

bool os_plan(int K, int Ca, int Cb, long long n_rows, OSParams& p) {
    if (K < 1 || K > OS_MAXK || Ca < 8 || Cb < 8 || Ca > 256 || Cb > 256 || (Ca & 7) || (Cb & 7)) return false;
    p.K = K; p.Ca = Ca; p.Cb = Cb;
    p.ca8 = Ca / 8;
    p.Npad = (Cb + 15) / 16 * 16;
    p.cb8 = p.Npad / 8;
    p.cbv = Cb / 8;
    p.MB = (Ca + 127) / 128;
    p.CH = 16 * p.MB;
    p.opm = p.MB == 1 ? std::min(16 / p.ca8, K) : 1;
    if (p.opm < 1) p.opm = 1;
    p.G = (K + p.opm - 1) / p.opm;
    p.gsA = (uint32_t)p.CH * OS_SBO;
    p.tileA = (OS_TR / 8) * p.gsA;
    p.gsB = (uint32_t)p.cb8 * OS_SBO;
    p.tileB = (OS_TR / 8) * p.gsB;
    p.meta_ints = OS_TR * K + OS_TR + OS_MAXK + 4;
    const int cols = p.MB * p.Npad;
    if (cols > 512) return false;
    // ring depth: 4 slots when they fit next to the two g tiles, else 2
    auto fits = [&](int S) {
        OSParams q = p;
        q.S = S;
        return os_layout(q).total <= 224u * 1024u;
    };
    p.S = fits(4) ? 4 : 2;
    if (!fits(p.S)) return false;
    {
        B200SP_ENV_INT(env_teams, "B200SP_WGOS_TEAMS", 0);
        p.nteams = p.S == 4 ? 4 : 2;
        if (env_teams == 1 || env_teams == 2 || (env_teams == 4 && p.S == 4)) p.nteams = env_teams;
    }
    // issuers x groups per pass.  Every pass re-reads the g tiles and walks all tiles again, so on the big levels the
    // FEWEST passes win and issuers come second (118 k rows, 32 x 32: one pass x 2 issuers 72 us, two passes x 4 issuers
    // 99 us); the deep levels (a handful of tiles) want CTAs, i.e. many passes with 4 issuers each.
    B200SP_ENV_INT(env_ni, "B200SP_WGOS_ISSUERS", 0);
    int ni = 0, best_pass = 1 << 30;
    for (int cand = (p.S == 4 ? 4 : 2); cand >= 1; cand >>= 1) {
        if (env_ni && cand != env_ni) continue;
        const int gpp = std::min(p.G, 512 / (cand * cols));
        if (gpp < 1) continue;
        const int npass = (p.G + gpp - 1) / gpp;
        if (n_rows <= 8192 && !env_ni) {  // small layer: the first (largest) issuer count that fits
            ni = cand;
            p.gpp = gpp;
            break;
        }
        if (npass < best_pass) {
            best_pass = npass;
            ni = cand;
            p.gpp = gpp;
        }
    }
    if (ni == 0) return false;
    p.ni = ni;
    uint32_t tc = 32;
    while (tc < (uint32_t)(p.ni * p.gpp * cols)) tc <<= 1;
    p.tmem_cols = tc;
    return true;
}

}  // namespace

bool wgrad_os_covers(int K, int Ca, int Cb) {
    B200SP_ENV_INT(env_os, "B200SP_WGRAD_OS", 1);  // 0: off, 1: shapes the register-gather kernel does not take, 2: all
    if (!env_os) return false;
    OSParams p{};
    return os_plan(K, Ca, Cb, 1 << 20, p);
}

int wgrad_os_run(const float* a, int Ca, const float* g, int Cb, const int* tab, const int* orow, const int* rowmask,
                 long long n_rows, int K, float* dW, cudaStream_t st) {
    OSParams p{};
    if (!os_plan(K, Ca, Cb, n_rows, p)) return B200SP_EUNSUP;
    if (!tab && K != 1) return B200SP_EUNSUP;
    B200SP_CHECK_ARG((((uintptr_t)a | (uintptr_t)g | (uintptr_t)dW) & 15) == 0, "wgrad_os: pointers must be 16-byte aligned");
    if (n_rows <= 0) return B200SP_OK;
    p.a = a; p.g = g; p.tab = tab; p.orow = orow; p.rowmask = rowmask; p.dW = dW; p.n_rows = n_rows;
    if (n_rows > ((long long)1 << 30) * OS_TR / 2) return B200SP_EUNSUP;
    p.ntiles = (int)((n_rows + OS_TR - 1) / OS_TR);
    const int npass = (p.G + p.gpp - 1) / p.gpp;
    const uint32_t smem = os_layout(p).total;
    static uint32_t attr_smem = 0;
    if (smem > attr_smem) {
        B200SP_CUDA(cudaFuncSetAttribute(k_wgrad_os, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_smem = smem;
    }
    int gx = std::max(1, num_sms() / npass);
    gx = std::min(gx, p.ntiles);
    note_kernel("k_wgrad_os");
    dim3 grid((unsigned)gx, (unsigned)npass);
    B200SP_CUDA(launch_pdl(k_wgrad_os, grid, dim3(OS_THREADS), smem, st, p));
    B200SP_LAUNCH_CHECK();
    return B200SP_OK;
}

}  // namespace b200sp
