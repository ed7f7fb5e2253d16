// conv_tc.cu — sparse convolution forward / dgrad on the Blackwell tensor cores (tcgen05 + TMEM), fp32-accurate.
//
// Replaces spconv v1.2 `indice_conv` / the dgrad half of `indice_conv_backward` (per offset: gather kernel ->
// cuBLAS SGEMM -> scatter-add).  One launch per conv, output-stationary:
//
//   out[r, :] = sum_k  in[tab[r,k], :] @ W[k]                 r in a 128-row tile owned by one CTA
//
//   * warps 8-11 (loaders, one thread per row): gather the neighbour rows of the tile for one (offset, channel-chunk)
//     stage with 16-byte cp.async (zero-fill for missing neighbours) straight into shared memory in the un-swizzled
//     canonical K-major UMMA layout (core matrix = 8 rows x 16 B); no registers, completion lands on an mbarrier,
//     so up to `nslots` stages of gathers stay in flight per CTA;
//   * warps 0-7 (transform): read the landed fp32 values back, write lo = v - tf32(v) into the twin tile (the tensor
//     core itself uses tf32(v) = the top 19 bits of the raw copy as the high part);
//   * warp 13: streams the matching pre-split weight block (prepared once per call by k_prep_weights, already in the
//     canonical layout) with one cp.async.bulk per stage;
//   * warp 12, one thread: issues tcgen05.mma kind::tf32 M=128 x N=Cout x K=8 — three products per K step
//     (hi*hi + lo*hi + hi*lo, i.e. 3xTF32: fp32-level accuracy, |err| ~ 2^-21) accumulating in TMEM over ALL offsets
//     and chunks, and releases each smem slot with tcgen05.commit;
//   * epilogue (warps 0-7): tcgen05.ld the 128 x Cout accumulator and write every output row exactly once
//     (no scatter atomics, no [P, C] gather/scatter buffers in HBM).
//
// dgrad is the same kernel on the transposed (SubM: mirrored) weights; the non-overlapping k2/s2 inverse conv and
// strided dgrad use the pair-grouped mode (one offset per CTA, rows scattered through the pair list).
#include "common.cuh"
#include "tc_common.cuh"
#include <map>
#include <mutex>
#include <cuda.h>  // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)
#include <algorithm>
#include <vector>
#include <stdlib.h>
#include <string.h>

namespace b200sp {

using namespace tc;

constexpr int TC_BM = 128;
constexpr int TC_MAXK = 32;
// warp roles of the persistent CTA (18 warps)
constexpr int TC_W_XFORM = 0;     // warps 0-3  : transform (lo = v - tf32(v)), thread <-> row
constexpr int TC_W_EPI = 4;       // warps 4-7  : epilogue, warp (4+q) reads TMEM lane quarter q
constexpr int TC_W_LOAD = 8;      // warps 8-11 : cp.async row gather straight into the canonical A layout
constexpr int TC_W_MMA = 12;      // warps 12-15: MMA issuers.  One thread sustains only ~1 tcgen05.mma per ~215 cycles
constexpr int TC_MAX_ISSUERS = 4; //   whatever its shape (tools/mma_bench) but issuers run concurrently: stages are dealt
                                  //   round-robin to ni warps, each with its own TMEM accumulator
constexpr int TC_W_WEIGHT = 16;   // warp 16    : weight blocks, one cp.async.bulk per stage
constexpr int TC_W_TILE = 17;     // warp 17    : tile prefetcher (table rows, active-offset list) one tile ahead
constexpr int TC_THREADS = 18 * 32;
constexpr int TC_XFORM_THREADS = 128;
constexpr int TC_LOADERS = 128;
constexpr int TC_NBUF = 2;        // tile-metadata buffers and TMEM accumulator sets

struct TCParams {
    const float* in;
    const float* Wp;     // prepared weights: [K][nchunks]{hi block, lo block}, block = canonical [KC/4][Cout_pad/8][8][4] fp32
    const int* tab;      // [n_rows][K] or NULL (identity, K == 1)
    const int* orow;     // output row of table row r (NULL: identity)
    const int* rowmask;  // K-bit mask of table row r (NULL: scan the table)
    const int* pin;      // pairs mode: [K][pstride]
    const int* pout;
    const int* pairnum;  // device [K]
    float* out;
    int64_t n_rows;
    int64_t pstride;
    int Cin, Cout, K;
    int nchunks, Cout_pad;
    int accumulate, pairs_mode;
    const float* res;       // added to the result (same [row][Cout] layout as out): out itself when accumulating, a residual
                            // branch for out = conv + res (model/unet_block.py:37), or nullptr
    int nslots;             // A ring (gathered rows, hi + lo tiles)
    int nslots_b;           // B ring (weight blocks): deeper, so weight copies run far ahead of the A slots
    uint32_t stageB_bytes;  // 2 * Cout_pad * KC * 4
    uint32_t tmem_cols;
    int ni;                 // MMA issuer warps in use (accumulators per set)
    int row_tiles;          // ceil(n_rows / 128)
    int total_tiles;        // row_tiles * (pairs_mode ? K : 1)
    int nsplit;             // table mode, few row tiles: the active offsets of a row tile are dealt to the nsplit CTAs of
                            // one thread-block CLUSTER; their partial tiles meet through distributed shared memory and
                            // are added in split order 0, 1, ... (deterministic: no float atomics on the output)
    int meta_bufs;          // tile-metadata buffers in shared memory: 2 (one tile prefetched ahead), or 1 when no CTA gets more
                            // than one tile -- the 14 KB that frees buy a fourth stage slot at two CTAs per SM
    int use_tma;            // gathers by cp.async.bulk.tensor tile::gather4 (tensor map = kernel parameter) instead of LDGSTS
    int dbg;                // dev only: 1 no MMAs, 2 no gathers, 4 no transform, 8 no epilogue data, 16 no table copy/mask, 32 no weight copies
};

template <int KC>
struct TCLayout {
    // K-adjacent core matrices of A are LBO bytes apart.  The pad (128/CPR bytes) staggers the banks so that the
    // loaders' quarter-warps (8/CPR rows x CPR chunks of 16 B) write 128 distinct bytes' worth of banks.
    static constexpr uint32_t CPR = KC / 4;                       // 16-byte chunks per row per stage
    static constexpr uint32_t LBO = TC_BM * 16u + 128u / CPR;
    static constexpr uint32_t A_BYTES = (CPR * LBO + 1023u) & ~1023u;  // one of {hi, lo}; 1024-aligned: the TMA path's
                                                                         // 128-byte-swizzled tile needs it
    static constexpr uint32_t ROWB = KC * 4u;                          // TMA path: bytes per row of the (swizzled) K-major tile
    __host__ __device__ static uint32_t meta_ints(int KT) { return (uint32_t)(TC_BM * KT + TC_BM + TC_MAXK + 4); }
    __host__ __device__ static uint32_t offB(int nslots) { return (uint32_t)nslots * 2u * A_BYTES; }
    __host__ __device__ static uint32_t offMeta(int nslots, int nslots_b, uint32_t stageB) {
        return offB(nslots) + (uint32_t)nslots_b * stageB;
    }
    __host__ __device__ static uint32_t offBars(int nslots, int nslots_b, uint32_t stageB, int KT, int mbufs) {
        uint32_t o = offMeta(nslots, nslots_b, stageB) + (uint32_t)mbufs * meta_ints(KT) * 4u;
        return (o + 15u) & ~15u;
    }
    __host__ __device__ static uint32_t total(int nslots, int nslots_b, uint32_t stageB, int KT, int mbufs) {
        return offBars(nslots, nslots_b, stageB, KT, mbufs) + (uint32_t)(3 * nslots + 2 * nslots_b + 5 * TC_NBUF) * 8u + 16u;
    }
};

__device__ __forceinline__ void cp_async16_zfill(uint32_t smem_dst, const void* gsrc, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_dst), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async4_zfill(uint32_t smem_dst, const void* gsrc, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(smem_dst), "l"(gsrc), "r"(src_bytes) : "memory");
}
// arrive on the mbarrier once every cp.async issued so far by this thread has landed (does not bump the pending count)
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}

// Persistent CTA: walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...  Every role keeps its own running counters
// (tile sequence number i, global stage number g); the smem stage ring and the two TMEM accumulator sets run
// seamlessly across tiles, so gathers of tile t+1 are in flight while tile t is multiplied and tile t-1 is stored.
template <int KC>
__global__ void __launch_bounds__(TC_THREADS, 2) k_conv_tc(const TCParams p, const __grid_constant__ CUtensorMap tmap) {  // 2 CTAs per SM: <= 56 registers
    using L = TCLayout<KC>;
    constexpr int NV = KC / 4;  // 16-byte chunks per row per stage
    extern __shared__ __align__(1024) unsigned char smem[];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int S = p.nslots;
    const int K = p.K;
    const int KT = p.pairs_mode ? 1 : K;
    const int ni = p.ni;
    unsigned char* sA = smem;
    unsigned char* sB = smem + L::offB(S);
    const int SB = p.nslots_b;
    int* s_meta = reinterpret_cast<int*>(smem + L::offMeta(S, SB, p.stageB_bytes));
    const int meta_ints = (int)L::meta_ints(KT);
    // per buffer: idx[128*KT] | orow[128] | klist[32] | nk | kfixed | pad
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::offBars(S, SB, p.stageB_bytes, KT, p.meta_bufs));  // hi+lo tiles ready -> MMA
    uint64_t* empty = full + S;                                                                // MMA done -> A slot reusable
    uint64_t* raw = empty + S;                                                                 // gathered rows landed -> transform
    uint64_t* fullb = raw + S;            // [SB] weight block landed -> MMA
    uint64_t* emptyb = fullb + SB;        // [SB] MMA done -> B slot reusable
    uint64_t* tready = emptyb + SB;       // [NBUF] tile metadata published
    uint64_t* tfree = tready + TC_NBUF;   // [NBUF] every reader is done with the tile metadata
    uint64_t* accf = tfree + TC_NBUF;     // [NBUF] accumulator set complete -> epilogue
    uint64_t* acce = accf + TC_NBUF;      // [NBUF] accumulator set drained -> issuers
    uint64_t* tload = acce + TC_NBUF;     // [NBUF] bulk copy of the tile's table rows landed (prefetcher only)
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(tload + TC_NBUF);

    const int ntiles = p.total_tiles > (int)blockIdx.x ? (p.total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    // barrier init spread over threads of three different warps (one thread doing all ~40 inits was ~1 us of every launch)
    if (tid < S) {
        mbar_init(&full[tid], TC_XFORM_THREADS / 32);
        mbar_init(&empty[tid], 1);
        mbar_init(&raw[tid], p.use_tma ? 1 : TC_LOADERS);
        mbar_fence_init();
    } else if (tid >= 32 && tid < 32 + SB) {
        mbar_init(&fullb[tid - 32], 1);
        mbar_init(&emptyb[tid - 32], 1);
        mbar_fence_init();
    } else if (tid >= 64 && tid < 64 + TC_NBUF) {
        const int b = tid - 64;
        mbar_init(&tready[b], 1);
        mbar_init(&tfree[b], (p.use_tma ? 1 : TC_LOADERS) + 4 /*xform warps*/ + 4 /*epilogue warps*/ + ni + 1 /*weights*/);
        mbar_init(&accf[b], ni);
        mbar_init(&acce[b], 4);
        mbar_init(&tload[b], 1);
        mbar_fence_init();
    }
    if (warp == TC_W_MMA) tmem_alloc(s_tmem, p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    const uint32_t tmem = *s_tmem;
    const int nchunks = p.nchunks;
    const int Cin = p.Cin, Cout = p.Cout;
    pdl_wait();  // barrier init and TMEM allocation above overlap the predecessor's tail; no global access before here

#define TC_META(b) (s_meta + (b) * meta_ints)
#define TC_IDX(b) (TC_META(b))
#define TC_OROW(b) (TC_META(b) + TC_BM * KT)
#define TC_KLIST(b) (TC_META(b) + TC_BM * KT + TC_BM)
#define TC_NK(b) (TC_META(b)[TC_BM * KT + TC_BM + TC_MAXK])
#define TC_KFIX(b) (TC_META(b)[TC_BM * KT + TC_BM + TC_MAXK + 1])
#define TC_NPOS(b) (TC_META(b)[TC_BM * KT + TC_BM + TC_MAXK + 2])  // active offsets of the row tile (all splits)

    if (warp == TC_W_TILE) {
        // ========== tile prefetcher ==========
        for (int i = 0; i < ntiles; ++i) {
            const int b = i & 1;
            const uint32_t ph = (uint32_t)(i >> 1) & 1u;
            mbar_wait(&tfree[b], ph ^ 1u);
            const int tile = (int)blockIdx.x + i * (int)gridDim.x;
            int* idx = TC_IDX(b);
            int* orow = TC_OROW(b);
            int* klist = TC_KLIST(b);
            if (p.pairs_mode) {
                const int kf = tile / p.row_tiles;
                const int64_t row0 = (int64_t)(tile - kf * p.row_tiles) * TC_BM;
                const int n = __ldg(p.pairnum + kf);
                const int rows = (int)max((int64_t)0, min((int64_t)TC_BM, (int64_t)n - row0));
                for (int r = lane; r < TC_BM; r += 32) {
                    const bool ok = r < rows;
                    idx[r] = ok ? __ldg(p.pin + (int64_t)kf * p.pstride + row0 + r) : -1;
                    orow[r] = ok ? __ldg(p.pout + (int64_t)kf * p.pstride + row0 + r) : -1;
                }
                if (lane == 0) {
                    klist[0] = 0;
                    TC_NK(b) = rows > 0 ? 1 : 0;
                    TC_KFIX(b) = kf;
                }
            } else {
                const int rt = tile / p.nsplit, split = tile - rt * p.nsplit;
                const int64_t row0 = (int64_t)rt * TC_BM;
                const int rows = (int)min((int64_t)TC_BM, p.n_rows - row0);
                unsigned mask = 0, mrow = 0;
                // the table rows first (one bulk copy in flight), the tile's output rows / row masks behind it: a CTA with
                // a single tile pays these global round trips back to back before its first gather can start
                const int* t = p.tab ? p.tab + row0 * K : nullptr;
                const int tot = rows * K;
                const uint32_t bytes = (uint32_t)tot * 4u;
                const bool bulk = p.tab && !(p.dbg & 16) && rows == TC_BM && (bytes & 15u) == 0 && ((reinterpret_cast<uintptr_t>(t) & 15) == 0);
                if (bulk && lane == 0) {
                    mbar_arrive_expect_tx(&tload[b], bytes);
                    bulk_g2s(idx, t, bytes, &tload[b]);
                }
                {
                    int ov[TC_BM / 32];
#pragma unroll
                    for (int q = 0; q < TC_BM / 32; ++q) {
                        const int r = lane + 32 * q;
                        ov[q] = r < rows ? (p.orow ? __ldg(p.orow + row0 + r) : (int)(row0 + r)) : -1;
                        if (p.rowmask && r < rows) mrow |= (unsigned)__ldg(p.rowmask + row0 + r);
                    }
#pragma unroll
                    for (int q = 0; q < TC_BM / 32; ++q) orow[lane + 32 * q] = ov[q];
                }
                if (p.tab) {
                    if (p.dbg & 16) {
                        if (lane == 0) mbar_arrive(&tload[b]);
                        __syncwarp();
                    } else if (bulk) {
                        mbar_wait(&tload[b], ph);
                    } else {
                        for (int e = lane; e < TC_BM * K; e += 32) idx[e] = e < tot ? __ldg(t + e) : -1;
                        if (lane == 0) mbar_arrive(&tload[b]);  // keep the phase of this buffer's barrier in step
                        __syncwarp();
                    }
                    if (p.rowmask) {
                        mask = mrow;  // per-row masks from the rulebook builder: 4 words per lane instead of 4*K
                    } else {
                        int kk = lane % K;  // column of element e = lane, lane+32, ...
                        const int step = 32 % K;
                        for (int e = lane; e < ((p.dbg & 16) ? 0 : TC_BM * K); e += 32) {
                            if (idx[e] >= 0) mask |= 1u << kk;
                            kk += step;
                            if (kk >= K) kk -= K;
                        }
                    }
                } else {
                    for (int r = lane; r < TC_BM; r += 32) idx[r] = r < rows ? (int)(row0 + r) : -1;
                    mask = rows > 0 ? 1u : 0u;
                }
                if (p.dbg & 16) mask = 0x1FFu;  // pretend 9 offsets
#pragma unroll
                for (int o = 16; o; o >>= 1) mask |= __shfl_xor_sync(0xffffffffu, mask, o);
                if (lane == 0) {
                    int nk = 0, pos = 0;
                    for (int k = 0; k < K; ++k)
                        if (mask >> k & 1u) {
                            if (pos % p.nsplit == split) klist[nk++] = k;  // this CTA's share of the active offsets
                            ++pos;
                        }
                    TC_NK(b) = nk;
                    TC_KFIX(b) = 0;
                    TC_NPOS(b) = pos;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&tready[b]);
        }
    } else if (warp >= TC_W_LOAD && warp < TC_W_LOAD + 4) {
        // ========== loaders: 16-byte cp.async (zero-fill for missing neighbours) straight into the canonical K-major
        // layout.  Consecutive lanes take consecutive 16-byte chunks of the SAME row, so one warp instruction touches
        // 32*16/(KC*4) rows = that many cache lines (not 32); completion is signalled on raw[slot] by the copy engine ==========
        const int lt = tid - TC_W_LOAD * 32;  // 0..127
        constexpr int CPR = KC / 4;
        const bool vec = (Cin % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.in) & 15) == 0);
        int slot = 0;
        uint32_t sph = 0;
        if (p.use_tma) {
            // ========== TMA gather: ONE warp, one cp.async.bulk.tensor tile::gather4 per lane per stage: lane l hands the
            // TMA engine the four table entries of rows 4l..4l+3 as row coordinates (a missing neighbour, -1, is out of
            // bounds and arrives as zeros), the engine generates the addresses, writes the 4 rows in the 128/64/32-byte
            // swizzled K-major layout the UMMA descriptor expects and completes the bytes on raw[slot] ==========
            if (warp == TC_W_LOAD) {
                for (int i = 0; i < ntiles; ++i) {
                    const int b = i & 1;
                    mbar_wait(&tready[b], (uint32_t)(i >> 1) & 1u);
                    const int* idx = TC_IDX(b) + 4 * lane * KT;
                    const int* klist = TC_KLIST(b);
                    const int nk = TC_NK(b);
                    for (int kk = 0; kk < nk; ++kk) {
                        const int kcol = p.pairs_mode ? 0 : klist[kk];
                        const int r0 = idx[kcol], r1 = idx[KT + kcol], r2 = idx[2 * KT + kcol], r3 = idx[3 * KT + kcol];
                        for (int c = 0; c < nchunks; ++c) {
                            mbar_wait(&empty[slot], sph ^ 1u);
                            if (lane == 0) mbar_arrive_expect_tx(&raw[slot], (uint32_t)TC_BM * L::ROWB);
                            __syncwarp();
                            if (!(p.dbg & 2)) {
                                const uint32_t dst = smem_u32(sA + (size_t)slot * 2 * L::A_BYTES) + (uint32_t)lane * 4u * L::ROWB;
                                asm volatile(
                                    "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes "
                                    "[%0], [%1, {%2, %3, %4, %5, %6}], [%7];\n" ::"r"(dst),
                                    "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(c * KC), "r"(r0), "r"(r1), "r"(r2), "r"(r3),
                                    "r"(smem_u32(&raw[slot]))
                                    : "memory");
                            } else if (lane == 0) {
                                asm volatile("mbarrier.complete_tx.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(&raw[slot])), "r"((uint32_t)TC_BM * L::ROWB) : "memory");
                            }
                            if (++slot == S) {
                                slot = 0;
                                sph ^= 1u;
                            }
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tfree[b]);
                }
            }
        } else
        for (int i = 0; i < ntiles; ++i) {
            const int b = i & 1;
            mbar_wait(&tready[b], (uint32_t)(i >> 1) & 1u);
            const int* idx = TC_IDX(b);
            const int* klist = TC_KLIST(b);
            const int nk = TC_NK(b);
            for (int kk = 0; kk < nk; ++kk) {
                const int kcol = p.pairs_mode ? 0 : klist[kk];
                for (int c = 0; c < nchunks; ++c) {
                    const int c0 = c * KC;
                    const uint32_t dst0 = smem_u32(sA + (size_t)slot * 2 * L::A_BYTES);
                    mbar_wait(&empty[slot], sph ^ 1u);
                    if (p.dbg & 2) {
                    } else if (vec) {
                        // thread-constant: chunk j and row phase; q only advances the 8-row group
                        const int j = lt % CPR, rowb = lt / CPR;
                        constexpr int RPQ = TC_LOADERS / CPR;  // rows covered per pass (multiple of 8)
                        const bool colok = c0 + 4 * j < Cin;
                        const uint32_t dstb = dst0 + (uint32_t)j * L::LBO + (uint32_t)(rowb >> 3) * 128u + (uint32_t)(rowb & 7) * 16u;
                        const int* ip = idx + rowb * KT + kcol;
                        // all table entries first: the cp.async asm is a compiler barrier ("memory"), interleaved with
                        // it every index load would sit in its own LDS -> compare -> address -> LDGSTS chain
                        int srcs[CPR];
#pragma unroll
                        for (int q = 0; q < CPR; ++q) srcs[q] = ip[q * RPQ * KT];
#pragma unroll
                        for (int q = 0; q < CPR; ++q) {
                            const int src = srcs[q];
                            const bool ok = (src >= 0) && colok;
                            const float* g = p.in + ((int64_t)(ok ? src : 0) * Cin + (ok ? c0 + 4 * j : 0));
                            cp_async16_zfill(dstb + (uint32_t)q * (RPQ / 8) * 128u, g, ok ? 16 : 0);
                        }
                    } else {
#pragma unroll 4
                        for (int q = 0; q < KC; ++q) {
                            const int e0 = q * TC_LOADERS + lt;
                            const int row = e0 / KC, e = e0 % KC;
                            const int src = idx[row * KT + kcol];
                            const bool ok = (src >= 0) && (c0 + e < Cin);
                            const float* g = p.in + (int64_t)(ok ? src : 0) * Cin + (ok ? c0 + e : 0);
                            cp_async4_zfill(dst0 + (uint32_t)(e >> 2) * L::LBO + (uint32_t)(row >> 3) * 128u +
                                                (uint32_t)(row & 7) * 16u + 4u * (uint32_t)(e & 3),
                                            g, ok ? 4 : 0);
                        }
                    }
                    cp_async_arrive_noinc(&raw[slot]);
                    if (++slot == S) {
                        slot = 0;
                        sph ^= 1u;
                    }
                }
            }
            mbar_arrive(&tfree[b]);
        }
        asm volatile("cp.async.wait_all;\n" ::: "memory");
    } else if (warp < TC_W_XFORM + 4) {
        // ========== transform: lo = v - tf32(v) next to the raw rows (the tensor core reads tf32(v) from the raw copy) ==========
        const int row = tid;  // 0..127
        const uint32_t soff = (uint32_t)(row >> 3) * 128u + (uint32_t)(row & 7) * 16u;
        int slot = 0;
        uint32_t sph = 0;
        for (int i = 0; i < ntiles; ++i) {
            const int b = i & 1;
            mbar_wait(&tready[b], (uint32_t)(i >> 1) & 1u);
            const int nit = TC_NK(b) * nchunks;
            __syncwarp();
            if (lane == 0) mbar_arrive(&tfree[b]);
            for (int it = 0; it < nit; ++it) {
                unsigned char* a_hi = sA + (size_t)slot * 2 * L::A_BYTES;
                unsigned char* a_lo = a_hi + L::A_BYTES;
                mbar_wait(&raw[slot], sph);
#pragma unroll
                for (int q = 0; q < ((p.dbg & 4) ? 0 : NV); ++q) {
                    // legacy layout: this thread's row, chunk q; TMA layout: any bijection will do (lo sits at the same
                    // swizzled position as hi), so consecutive threads take consecutive 16-byte pieces
                    const uint32_t off = p.use_tma ? (uint32_t)(q * TC_XFORM_THREADS + tid) * 16u : (uint32_t)q * L::LBO + soff;
                    const float4 v = *reinterpret_cast<const float4*>(a_hi + off);
                    float4 h, l;
                    split_tf32(v.x, h.x, l.x);
                    split_tf32(v.y, h.y, l.y);
                    split_tf32(v.z, h.z, l.z);
                    split_tf32(v.w, h.w, l.w);
                    *reinterpret_cast<float4*>(a_lo + off) = l;
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&full[slot]);
                if (++slot == S) {
                    slot = 0;
                    sph ^= 1u;
                }
            }
        }
    } else if (warp >= TC_W_EPI && warp < TC_W_EPI + 4) {
        // ========== epilogue: TMEM -> registers -> global (each output row written once) ==========
        const int q4 = warp - TC_W_EPI;
        const bool vecO = (Cout % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0);
        int g0 = 0;  // global stage number of the tile's first stage
        uint32_t aph[TC_NBUF] = {0, 0};  // phase of each accumulator set = number of NON-EMPTY tiles it served (mod 2)
        for (int i = 0; i < ntiles; ++i) {
            const int b = i & 1;
            const uint32_t ph = (uint32_t)(i >> 1) & 1u;
            mbar_wait(&tready[b], ph);
            const int nit = TC_NK(b) * nchunks;
            const int orow = TC_OROW(b)[q4 * 32 + lane];
            __syncwarp();
            if (lane == 0) mbar_arrive(&tfree[b]);
            if (nit > 0) {
                mbar_wait(&accf[b], aph[b]);
                aph[b] ^= 1u;
                tc_fence_after();
            }
            const int nused = min(ni, nit);
            const uint32_t set_col = tmem + (uint32_t)(b * ni * p.Cout_pad);
            const bool split_mode = p.nsplit > 1;
            const int tile = (int)blockIdx.x + i * (int)gridDim.x;
            const int rt = split_mode ? tile / p.nsplit : 0, split = split_mode ? tile - rt * p.nsplit : 0;
            // split mode: this CTA's partial sums of the tile go to its OWN shared memory (the stage rings are idle once
            // the accumulators are complete), row r at s_part[r][:] with a 4-float pad against bank conflicts
            float* part = split_mode ? reinterpret_cast<float*>(smem) + (size_t)(q4 * 32 + lane) * (p.Cout_pad + 4) : nullptr;
            (void)rt; (void)split;
            for (int ch = 0; ch * 16 < (((split_mode && nit == 0) || (p.dbg & 8)) ? 0 : p.Cout_pad); ++ch) {
                float v[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) v[e] = 0.f;
                for (int u = 0; u < nused; ++u) {
                    const int w = (g0 + u) % ni;  // issuers that received a stage of this tile
                    float t[16];
                    tmem_ld16(set_col + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(w * p.Cout_pad + ch * 16), t);
#pragma unroll
                    for (int e = 0; e < 16; ++e) v[e] += t[e];
                }
                if (split_mode) {
#pragma unroll
                    for (int g4 = 0; g4 < 4; ++g4)
                        *reinterpret_cast<float4*>(part + ch * 16 + g4 * 4) =
                            make_float4(v[g4 * 4 + 0], v[g4 * 4 + 1], v[g4 * 4 + 2], v[g4 * 4 + 3]);
                    continue;
                }
                if (orow < 0) continue;
                float* o = p.out + (int64_t)orow * Cout + ch * 16;
                const float* rs = p.res ? p.res + (int64_t)orow * Cout + ch * 16 : nullptr;
#pragma unroll
                for (int g4 = 0; g4 < 4; ++g4) {
                    const int col = ch * 16 + g4 * 4;
                    if (col >= Cout) break;
                    if (vecO) {
                        float4 wv = make_float4(v[g4 * 4 + 0], v[g4 * 4 + 1], v[g4 * 4 + 2], v[g4 * 4 + 3]);
                        if (rs) {
                            const float4 old = *reinterpret_cast<const float4*>(rs + g4 * 4);
                            wv.x += old.x; wv.y += old.y; wv.z += old.z; wv.w += old.w;
                        }
                        *reinterpret_cast<float4*>(o + g4 * 4) = wv;
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            if (col + e < Cout) {
                                float wv = v[g4 * 4 + e];
                                if (rs) wv += rs[g4 * 4 + e];
                                o[g4 * 4 + e] = wv;
                            }
                        }
                    }
                }
            }
            if (nit > 0) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acce[b]);
            }
            g0 += nit;
        }
    } else if (warp >= TC_W_MMA && warp < TC_W_MMA + TC_MAX_ISSUERS) {
        // ========== MMA issuers: warp w owns the global stages g = w (mod ni) and accumulator w of each set.  The whole
        // warp walks the loops (uniform control flow), one elected lane issues; descriptors advance by constant adds ==========
        const int w = warp - TC_W_MMA;
        if (w < ni) {
            const uint32_t idesc = make_idesc_tf32(TC_BM, p.Cout_pad, 0, 0);
            const uint32_t lboB = (uint32_t)p.Cout_pad * 16u;  // K-adjacent core matrices of B
            // descriptors of slot 0 (hi tiles); lo tiles / other slots / K steps are constant offsets in 16-byte units
            // A: legacy = un-swizzled canonical K-major (LBO between 16-byte K chunks, 128 B between 8-row groups);
            // TMA = swizzled K-major rows of ROWB bytes (SBO = 8 rows, layout code 2 / 4 / 6 = SWIZZLE_128B / 64B / 32B)
            const uint64_t dA0 = p.use_tma
                                     ? (make_desc(smem_u32(sA), 16u, 8u * L::ROWB) |
                                        ((uint64_t)(KC == 32 ? 2 : (KC == 16 ? 4 : 6)) << 61))
                                     : make_desc(smem_u32(sA), L::LBO, 128u);
            const uint64_t dB0 = make_desc(smem_u32(sB), lboB, 128u);
            const uint32_t slotA16 = (2u * L::A_BYTES) >> 4, loA16 = L::A_BYTES >> 4, kA16 = p.use_tma ? 2u : (2u * L::LBO) >> 4;
            const uint32_t slotB16 = p.stageB_bytes >> 4, loB16 = p.stageB_bytes >> 5, kB16 = (2u * lboB) >> 4;
            int g0 = 0;
            uint32_t aph[TC_NBUF] = {0, 0};  // as in the epilogue: empty tiles do not touch the accumulator barriers
            for (int i = 0; i < ntiles; ++i) {
                const int b = i & 1;
                const uint32_t ph = (uint32_t)(i >> 1) & 1u;
                mbar_wait(&tready[b], ph);
                const int nit = TC_NK(b) * nchunks;
                __syncwarp();
                if (lane == 0) mbar_arrive(&tfree[b]);
                if (nit > 0) {
                    mbar_wait(&acce[b], aph[b] ^ 1u);  // the epilogue drained this accumulator set
                    aph[b] ^= 1u;
                    tc_fence_after();
                    const uint32_t dcol = tmem + (uint32_t)((b * ni + w) * p.Cout_pad);
                    uint32_t acc = 0;
                    // first global stage >= g0 owned by this issuer
                    int g = g0 + ((w - g0 % ni) + ni) % ni;
                    for (; g < g0 + nit; g += ni) {
                        const int slot = g % S, sb = g % SB;
                        const uint32_t sph = (uint32_t)(g / S) & 1u, bph = (uint32_t)(g / SB) & 1u;
                        mbar_wait(&fullb[sb], bph);
                        mbar_wait(&full[slot], sph);
                        tc_fence_after();
                        const uint64_t da = dA0 + (uint64_t)((uint32_t)slot * slotA16);
                        const uint64_t db = dB0 + (uint64_t)((uint32_t)sb * slotB16);
                        if (elect_one()) {
#pragma unroll
                            for (int j = 0; j < ((p.dbg & 1) ? 0 : KC / 8); ++j) {
                                mma_tf32_ss(dcol, da + j * kA16, db + j * kB16, idesc, acc);
                                mma_tf32_ss(dcol, da + loA16 + j * kA16, db + j * kB16, idesc, 1u);
                                mma_tf32_ss(dcol, da + j * kA16, db + loB16 + j * kB16, idesc, 1u);
                                acc = 1u;
                            }
                            mma_commit(&empty[slot]);
                            mma_commit(&emptyb[sb]);
                        }
                        acc = 1u;
                        __syncwarp();
                    }
                    if (elect_one()) mma_commit(&accf[b]);
                    __syncwarp();
                }
                g0 += nit;
            }
        }
    } else if (warp == TC_W_WEIGHT) {
        // ========== weight loader (one thread): one bulk copy per stage ==========
        if (lane == 0) {
            int slot = 0;
            uint32_t sph = 0;
            for (int i = 0; i < ntiles; ++i) {
                const int b = i & 1;
                mbar_wait(&tready[b], (uint32_t)(i >> 1) & 1u);
                const int nk = TC_NK(b);
                const int kfixed = TC_KFIX(b);
                const int* klist = TC_KLIST(b);
                for (int kk = 0; kk < nk; ++kk) {
                    const int kw = p.pairs_mode ? kfixed : klist[kk];
                    for (int c = 0; c < nchunks; ++c) {
                        mbar_wait(&emptyb[slot], sph ^ 1u);
                        if (p.dbg & 32) {
                            mbar_arrive(&fullb[slot]);
                        } else {
                        mbar_arrive_expect_tx(&fullb[slot], p.stageB_bytes);
                        bulk_g2s(sB + (size_t)slot * p.stageB_bytes,
                                 reinterpret_cast<const unsigned char*>(p.Wp) + ((size_t)kw * nchunks + c) * p.stageB_bytes,
                                 p.stageB_bytes, &fullb[slot]);
                        }
                        if (++slot == SB) {
                            slot = 0;
                            sph ^= 1u;
                        }
                    }
                }
                mbar_arrive(&tfree[b]);
            }
        }
    }
    pdl_trigger();  // this CTA's work is issued: the next kernel of the stream may start its prologue (TMEM-heavy kernels
                    // trigger late -- an early-resident successor would sit on shared memory / TMEM for the whole run)
    if (p.nsplit > 1) {
        // ---- split mode: one tile per CTA, the nsplit CTAs of the row tile form one cluster ----
        // every thread of the cluster: partial tiles are complete and visible cluster-wide
        asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
        if (warp >= TC_W_EPI && warp < TC_W_EPI + 4 && !(p.dbg & 8)) {
            const int et = (warp - TC_W_EPI) * 32 + lane;
            const int rt = (int)blockIdx.x / p.nsplit, split = (int)blockIdx.x - rt * p.nsplit;
            const int nact = min(p.nsplit, TC_NPOS(0));  // splits that were dealt at least one offset (same in every CTA)
            const int cp4 = p.Cout_pad >> 2, pstr4 = cp4 + 1;
            const int rows_valid = (int)min((int64_t)TC_BM, p.n_rows - (int64_t)rt * TC_BM);
            const int total = rows_valid * cp4;
            const int per = (total + p.nsplit - 1) / p.nsplit;
            const int e_end = min(total, (split + 1) * per);
            const bool vecO = (Cout % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0);
            const int* orow_s = TC_OROW(0);
            const uint32_t base = smem_u32(smem);
            // this CTA adds ITS slice of the tile's [rows][Cout_pad] block over the peers' partial tiles, in split order
            for (int e = split * per + et; e < e_end; e += 128) {
                const int row = e / cp4, c4 = e - row * cp4;
                const uint32_t laddr = base + (uint32_t)(row * pstr4 + c4) * 16u;
                // all peer loads in flight first (distributed shared memory is a ~200-cycle round trip), then the adds in
                // split order
                float4 tp[8];
#pragma unroll
                for (int sp = 0; sp < 8; ++sp) {
                    tp[sp] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (sp < nact) {
                        uint32_t raddr;
                        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(raddr) : "r"(laddr), "r"(sp));
                        asm volatile("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4];\n"
                                     : "=f"(tp[sp].x), "=f"(tp[sp].y), "=f"(tp[sp].z), "=f"(tp[sp].w)
                                     : "r"(raddr));
                    }
                }
                float4 acc = tp[0];
#pragma unroll
                for (int sp = 1; sp < 8; ++sp) {
                    if (sp < nact) { acc.x += tp[sp].x; acc.y += tp[sp].y; acc.z += tp[sp].z; acc.w += tp[sp].w; }
                }
                if (c4 * 4 >= Cout) continue;
                const int orw = orow_s[row];
                if (orw < 0) continue;
                float* o = p.out + (int64_t)orw * Cout + c4 * 4;
                const float* rs = p.res ? p.res + (int64_t)orw * Cout + c4 * 4 : nullptr;
                const float av[4] = {acc.x, acc.y, acc.z, acc.w};
                if (vecO) {
                    float4 wv = acc;
                    if (rs) {
                        const float4 old = *reinterpret_cast<const float4*>(rs);
                        wv.x += old.x; wv.y += old.y; wv.z += old.z; wv.w += old.w;
                    }
                    *reinterpret_cast<float4*>(o) = wv;
                } else {
#pragma unroll
                    for (int e2 = 0; e2 < 4; ++e2)
                        if (c4 * 4 + e2 < Cout) o[e2] = rs ? rs[e2] + av[e2] : av[e2];
                }
            }
        }
        // nobody leaves (and frees its shared memory) while a peer may still read it
        asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
    }
#undef TC_META
#undef TC_IDX
#undef TC_OROW
#undef TC_KLIST
#undef TC_NK
#undef TC_KFIX
#undef TC_NPOS
    tc_fence_before();
    __syncthreads();
    if (warp == TC_W_MMA) tmem_dealloc(tmem, p.tmem_cols);
}

// W [K][Ci_w][Co_w] fp32 -> per (offset, chunk) {hi, lo} blocks in the canonical K-major B layout.
// transposed == 0: B_k[n][kin] = W[k][kin][n]        (forward: kin over Ci_w, n over Co_w)
// transposed == 1: B_k[n][kin] = W[k][n][kin]        (dgrad:   kin over Co_w, n over Ci_w)
// mirror: use W[K-1-k] for block k (SubM dgrad)
__global__ void k_prep_weights(const float* __restrict__ W, int K, int Ci_w, int Co_w, int transposed, int mirror,
                               int KC, int nchunks, int Cout_pad, float* __restrict__ Wp) {
    const int Cin = transposed ? Co_w : Ci_w;
    const int Cout = transposed ? Ci_w : Co_w;
    const int64_t per_block = (int64_t)Cout_pad * KC;
    const int64_t total = (int64_t)K * nchunks * per_block;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        // i enumerates (k, c, kin_local, n) with n fastest so that global reads of W are coalesced for the forward case
        const int n = (int)(i % Cout_pad);
        int64_t t = i / Cout_pad;
        const int kl = (int)(t % KC);
        t /= KC;
        const int c = (int)(t % nchunks);
        const int k = (int)(t / nchunks);
        const int kin = c * KC + kl;
        float v = 0.f;
        if (kin < Cin && n < Cout) {
            const int ks = mirror ? K - 1 - k : k;
            v = transposed ? W[((int64_t)ks * Ci_w + n) * Co_w + kin] : W[((int64_t)ks * Ci_w + kin) * Co_w + n];
        }
        float hi, lo;
        tc::split_tf32(v, hi, lo);
        const int64_t base = ((int64_t)k * nchunks + c) * 2 * per_block;
        const int64_t off = ((int64_t)(kl >> 2) * (Cout_pad >> 3) + (n >> 3)) * 32 + (n & 7) * 4 + (kl & 3);
        Wp[base + off] = hi;
        Wp[base + per_block + off] = lo;
    }
}

// one launch for MANY weight tensors (all conv layers of a model): blockIdx.y = tensor
struct PrepItem {
    const float* W;
    float* Wp;
    int K, Ci_w, Co_w, transposed, mirror, KC, nchunks, Cout_pad;
};
__global__ void k_prep_weights_batch(const PrepItem* __restrict__ items) {
    const PrepItem it = items[blockIdx.y];
    const int Cin = it.transposed ? it.Co_w : it.Ci_w;
    const int Cout = it.transposed ? it.Ci_w : it.Co_w;
    const int64_t per_block = (int64_t)it.Cout_pad * it.KC;
    const int64_t total = (int64_t)it.K * it.nchunks * per_block;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int n = (int)(i % it.Cout_pad);
        int64_t t = i / it.Cout_pad;
        const int kl = (int)(t % it.KC);
        t /= it.KC;
        const int c = (int)(t % it.nchunks);
        const int k = (int)(t / it.nchunks);
        const int kin = c * it.KC + kl;
        float v = 0.f;
        if (kin < Cin && n < Cout) {
            const int ks = it.mirror ? it.K - 1 - k : k;
            v = it.transposed ? it.W[((int64_t)ks * it.Ci_w + n) * it.Co_w + kin] : it.W[((int64_t)ks * it.Ci_w + kin) * it.Co_w + n];
        }
        float hi, lo;
        tc::split_tf32(v, hi, lo);
        const int64_t base = ((int64_t)k * it.nchunks + c) * 2 * per_block;
        const int64_t off = ((int64_t)(kl >> 2) * (it.Cout_pad >> 3) + (n >> 3)) * 32 + (n & 7) * 4 + (kl & 3);
        it.Wp[base + off] = hi;
        it.Wp[base + per_block + off] = lo;
    }
}

struct TCPlan {
    int KC, nchunks, Cout_pad, nslots, nslots_b, ni, meta_bufs;
    uint32_t stageB, tmem_cols, smem;
    int64_t wp_bytes;
};

// two CTAs per SM: 228 KB of shared memory per SM, 1 KB reserved per resident CTA
constexpr uint32_t TC_TWO_CTA_SMEM = (233472u - 2u * 1024u) / 2u - 256u;

static bool tc_plan(int K, int Cin, int Cout, int KT, TCPlan& pl, int64_t n_tiles = 0, int force_meta_bufs = 0) {
    if (K < 1 || K > TC_MAXK || Cin < 1 || Cout < 1) return false;
    const int Cin_pad = (Cin + 7) / 8 * 8;
    pl.KC = (Cin_pad % 32 == 0) ? 32 : (Cin_pad % 16 == 0) ? 16 : 8;
    pl.nchunks = Cin_pad / pl.KC;
    pl.Cout_pad = (Cout + 15) / 16 * 16;
    if (pl.Cout_pad > 256) return false;
    pl.stageB = 2u * (uint32_t)pl.Cout_pad * pl.KC * 4u;
    pl.wp_bytes = (int64_t)K * pl.nchunks * pl.stageB;
    const uint32_t cpr = (uint32_t)pl.KC / 4u;
    const uint32_t a_bytes = 2u * ((cpr * (TC_BM * 16u + 128u / cpr) + 1023u) & ~1023u);
    // one metadata buffer is enough when no CTA can get a second tile (n_tiles <= resident CTAs, checked again below and
    // in launch_tc); the double buffer only exists to prefetch the NEXT tile's table rows
    B200SP_ENV_INT(env_mb, "B200SP_TC_META_BUFS", 0);  // dev knob: 2 = always double-buffered (round-2 behaviour before)
    const int want_bufs = force_meta_bufs ? force_meta_bufs : (env_mb ? env_mb : ((n_tiles > 0 && n_tiles <= 2 * (int64_t)num_sms()) ? 1 : 2));
    pl.meta_bufs = want_bufs;
    const uint32_t fixed = (uint32_t)want_bufs * (uint32_t)(TC_BM * KT + TC_BM + TC_MAXK + 4) * 4u + 1536u;
    // slots: 4 under ~110 KB lets two CTAs share an SM; big stages fall back to fewer slots / one CTA per SM
    int best = 0;
    int smax = 4;
    uint32_t two_cta_budget = TC_TWO_CTA_SMEM;
    {
        B200SP_ENV_INT(env_slots, "B200SP_TC_SLOTS", 0);  // dev knob: > 4 trades the second CTA per SM for a deeper ring
        if (env_slots >= 2) {
            smax = env_slots;
            if (smax > 4) two_cta_budget = 0;
        } else if (n_tiles > 0 && n_tiles * 2 <= num_sms()) {
            // few row tiles (split mode, levels 4-7 of DODA's net): even split 8 ways the layer has at most one CTA
            // per SM, so the second CTA's shared memory buys a deeper ring instead (6 148 rows x 64: 28.4 -> 26.2 us)
            smax = 6;
            two_cta_budget = 0;
        }
    }
    uint32_t budget = 0;
    for (int s = smax; s >= 2; --s) {
        const uint32_t tot = (uint32_t)s * (a_bytes + pl.stageB) + fixed;
        if (tot <= two_cta_budget || ((s <= 3 || smax > 4) && tot <= 220u * 1024u)) {
            best = s;
            budget = tot <= two_cta_budget ? two_cta_budget : 220u * 1024u;
            break;
        }
    }
    if (best == 0) return false;
    pl.nslots = best;
    // weight ring: as deep as the remaining budget allows (multiple of the A ring, at most 4x), because a weight
    // block can only be requested once its slot is free and its L2 latency would otherwise sit in every A slot's chain
    {
        const uint32_t used = (uint32_t)best * a_bytes + fixed;
        int mult = 4;
        B200SP_ENV_INT(env_bmult, "B200SP_TC_BMULT", 0);
        if (env_bmult >= 1) mult = env_bmult;
        while (mult > 1 && used + (uint32_t)(best * mult) * (pl.stageB + 16u) > budget) --mult;
        pl.nslots_b = best * mult;
    }
    // issuer warps: each smem slot must belong to exactly ONE issuer (its mbarrier waits are parity waits, an issuer
    // running a fill ahead of a slot it shares would alias phases), so ni divides nslots; two accumulator sets.
    // TMEM decides the residency: with more tiles than SMs two CTAs per SM (<= 256 columns each) beat one CTA with
    // twice the issuers -- the layer then runs in one round instead of two (ncu, level 3 of DODA's net: 207 tiles on
    // 148 single CTAs took 72 us, the stage ring being latency-bound at ~0.55 us per stage per CTA).
    {
        B200SP_ENV_INT(env_issuers, "B200SP_TC_ISSUERS", TC_MAX_ISSUERS);
        const int want = std::max(1, std::min(env_issuers, TC_MAX_ISSUERS));
        const bool two_cta_smem = budget == two_cta_budget && two_cta_budget > 0;
        const uint32_t col_cap = (two_cta_smem && n_tiles > num_sms()) ? 256u : 512u;
        pl.ni = 0;
        for (int pass = 0; pass < 2 && pl.ni == 0; ++pass) {
            const uint32_t cap = pass == 0 ? col_cap : 512u;
            for (int ni = want; ni >= 1; --ni)
                if (best % ni == 0 && (uint32_t)(TC_NBUF * ni * pl.Cout_pad) <= cap) {
                    pl.ni = ni;
                    break;
                }
        }
        if (pl.ni == 0) return false;
    }
    uint32_t cols = 32;
    while (cols < (uint32_t)(TC_NBUF * pl.ni * pl.Cout_pad)) cols <<= 1;
    pl.tmem_cols = cols;
    if (pl.meta_bufs == 1 && !force_meta_bufs) {
        // residency this plan will get (launch_tc computes the same): a CTA must never see a second tile
        int occ = budget == two_cta_budget && two_cta_budget > 0 ? 2 : 1;
        while (occ > 1 && (uint32_t)occ * cols > 512u) --occ;
        if (n_tiles > (int64_t)num_sms() * occ) return tc_plan(K, Cin, Cout, KT, pl, n_tiles, 2);
    }
    return true;
}

static int g_tc_tma = -1;
int conv_tc_tma_enabled() {
    if (g_tc_tma < 0) {
        const char* e = getenv("B200SP_TC_TMA");
        g_tc_tma = (e && e[0] == '1') ? 1 : 0;
    }
    return g_tc_tma;
}
void conv_tc_set_tma(int on) { g_tc_tma = on ? 1 : 0; }

// cuTensorMapEncodeTiled through the runtime's driver entry point (the library links cudart only)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            f = nullptr;
        return reinterpret_cast<EncodeTiledFn>(f);
    }();
    return fn;
}

// 2-D map of the feature matrix [rows][Cin] fp32 for tile::gather4: box = {KC columns, 1 row} (the instruction names 4
// rows), swizzle = the row size (128 / 64 / 32 B), out-of-bounds rows / columns read as zeros
static bool make_gather_map(const float* in, int64_t rows, int Cin, int KC, CUtensorMap* map) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return false;
    cuuint64_t gdim[2] = {(cuuint64_t)Cin, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)Cin * 4u};
    cuuint32_t box[2] = {(cuuint32_t)KC, 1u};
    cuuint32_t estr[2] = {1u, 1u};
    const CUtensorMapSwizzle sw = KC == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : (KC == 16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(in), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int KC>
static int launch_tc(const TCParams& p0, int KT, int64_t n_in, cudaStream_t st) {
    using L = TCLayout<KC>;
    TCParams p = p0;
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    {
        // TMA gather is possible whenever the rows are 16-byte aligned multiples of 16 bytes (everything but the 3-channel
        // input conv).  It is OPT-IN (B200SP_TC_TMA=1 / b200sp_set_conv_tma): measured per launch, gather4 vs LDGSTS:
        // 26.5 k rows x 48 ch 115 vs 65 us, 6 148 x 64 58 vs 47, 1 380 x 80 43 vs 29, 222 x 96 29 vs 25 -- with 64..128-byte
        // rows one gather4 moves 256..512 bytes and the 32 instructions of a stage are bound by the TMA unit's issue
        // rate, while 128 LDGSTS threads issue in parallel.
        const int env_tma = conv_tc_tma_enabled();
        p.use_tma = 0;
        if (env_tma && p.Cin % 4 == 0 && (reinterpret_cast<uintptr_t>(p.in) & 15) == 0 && p.Cin >= KC / 2)
            p.use_tma = make_gather_map(p.in, n_in > 0 ? n_in : ((int64_t)1 << 31) - 1, p.Cin, KC, &tmap) ? 1 : 0;
    }
    const uint32_t smem = L::total(p.nslots, p.nslots_b, p.stageB_bytes, KT, p.meta_bufs);
    static uint32_t attr_smem = 0;
    if (smem > attr_smem) {
        B200SP_CUDA(cudaFuncSetAttribute(k_conv_tc<KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        // two CTAs of up to 113 KB each only fit with the whole 228 KB carved out as shared memory
        B200SP_CUDA(cudaFuncSetAttribute(k_conv_tc<KC>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
        attr_smem = smem;
    }
    // persistent grid: as many CTAs as fit on the machine (TMEM: 512 columns per SM), never more than tiles
    int occ = smem <= TC_TWO_CTA_SMEM ? 2 : 1;
    while (occ > 1 && (uint32_t)occ * p.tmem_cols > 512u) --occ;
    if (p.nsplit > 1) {
        // split mode: p.nsplit is the WISH.  One tile per CTA, the nsplit CTAs of a row tile are one cluster (co-scheduled
        // by the hardware, so they may wait for each other); their partial tiles must fit into the idle stage rings
        p.nsplit = std::min(std::min(p.nsplit, 8), num_sms() * occ / p.row_tiles);
        // ... and ALL clusters must be resident at once: a cluster lives inside one GPC, so e.g. 49 clusters of 3 CTAs do
        // not fit a machine whose GPCs hold 6 such clusters each (2 SMs per GPC stay empty) and the stragglers run as a
        // second wave (measured 6 148 rows x 64 ch: 3-way 39.1 us, 2-way 31.6 us).  Ask the driver, per (smem, n), once.
        {
            static std::map<uint64_t, int> s_fit;
            static std::mutex s_fit_mu;
            std::lock_guard<std::mutex> lk(s_fit_mu);
            B200SP_ENV_INT(env_fit, "B200SP_TC_FITCHECK", 1);  // dev knob: 0 = the pre-check behaviour
            while (env_fit && p.nsplit > 1) {
                const uint64_t key = ((uint64_t)smem << 8) | (uint64_t)p.nsplit;
                auto it = s_fit.find(key);
                int fit = 0;
                if (it == s_fit.end()) {
                    cudaLaunchConfig_t q{};
                    q.gridDim = dim3((unsigned)(p.nsplit * 64));
                    q.blockDim = dim3(TC_THREADS);
                    q.dynamicSmemBytes = smem;
                    cudaLaunchAttribute qa[1];
                    qa[0].id = cudaLaunchAttributeClusterDimension;
                    qa[0].val.clusterDim.x = (unsigned)p.nsplit;
                    qa[0].val.clusterDim.y = 1;
                    qa[0].val.clusterDim.z = 1;
                    q.attrs = qa;
                    q.numAttrs = 1;
                    if (cudaOccupancyMaxActiveClusters(&fit, k_conv_tc<KC>, &q) != cudaSuccess) {
                        (void)cudaGetLastError();
                        fit = num_sms() * occ / p.nsplit;
                    }
                    if (occ == 1) fit = std::min(fit, num_sms() / p.nsplit);  // TMEM (unknown to the driver) allows one CTA per SM
                    s_fit[key] = fit;
                } else {
                    fit = it->second;
                }
                if (fit >= p.row_tiles) break;
                --p.nsplit;
            }
        }
        const uint32_t part_bytes = (uint32_t)TC_BM * (uint32_t)(p.Cout_pad + 4) * 4u;
        if (part_bytes > L::offMeta(p.nslots, p.nslots_b, p.stageB_bytes)) p.nsplit = 1;
        if (p.nsplit < 2) p.nsplit = 1;
        p.total_tiles = p.row_tiles * p.nsplit;
        if (p.nsplit > 1) {
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3((unsigned)p.total_tiles);
            cfg.blockDim = dim3(TC_THREADS);
            cfg.dynamicSmemBytes = smem;
            cfg.stream = st;
            cudaLaunchAttribute attr[2];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = (unsigned)p.nsplit;
            attr[0].val.clusterDim.y = 1;
            attr[0].val.clusterDim.z = 1;
            attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[1].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attr;
            cfg.numAttrs = pdl_enabled() ? 2 : 1;
            B200SP_CUDA(cudaLaunchKernelEx(&cfg, k_conv_tc<KC>, p, tmap));
            B200SP_LAUNCH_CHECK();
            return B200SP_OK;
        }
    }
    const int grid = std::min(p.total_tiles, num_sms() * occ);
    B200SP_CHECK_ARG(p.meta_bufs == 2 || p.total_tiles <= grid, "conv_tc: single metadata buffer with %d tiles on %d CTAs", p.total_tiles, grid);
    B200SP_CUDA(launch_pdl(k_conv_tc<KC>, dim3((unsigned)grid), dim3(TC_THREADS), smem, st, p, tmap));
    B200SP_LAUNCH_CHECK();
    return B200SP_OK;
}

// returns B200SP_EUNSUP when the shape is outside what the tensor path covers (caller falls back to the fp32 kernel)
int conv_tc_run(const float* in, int Cin, const float* W, int Ci_w, int Co_w, int wflags, const int* tab, const int* orow,
                const int* rowmask, const int* pin, const int* pout, const int* pairnum, int64_t n_rows, int64_t pstride, int K, float* out,
                int Cout, int accumulate, int pairs_mode, void* ws, int64_t ws_bytes, cudaStream_t st, int64_t n_in, const float* res) {
    TCPlan pl;
    const int KT = pairs_mode ? 1 : K;
    if (!tc_plan(K, Cin, Cout, KT, pl, cdiv(n_rows, TC_BM) * (pairs_mode ? K : 1))) return B200SP_EUNSUP;
    if (!(wflags & 4) && (!ws || ws_bytes < pl.wp_bytes)) {
        set_error("conv_tc: workspace too small (%lld < %lld bytes)", (long long)ws_bytes, (long long)pl.wp_bytes);
        return B200SP_ENOMEM;
    }
    if (cdiv(n_rows, TC_BM) * (pairs_mode ? K : 1) > (int64_t)1 << 30) return B200SP_EUNSUP;
    float* Wp = static_cast<float*>(ws);
    if (wflags & 4) {
        Wp = const_cast<float*>(W);  // W already IS the prepared image (b200sp_prep_weights_batch)
    } else {
        const int64_t total = (int64_t)K * pl.nchunks * pl.Cout_pad * pl.KC;
        const unsigned blocks = (unsigned)std::min<int64_t>(cdiv(total, 256), 4 * 148);
        k_prep_weights<<<blocks, 256, 0, st>>>(W, K, Ci_w, Co_w, wflags & 1, (wflags >> 1) & 1, pl.KC, pl.nchunks,
                                               pl.Cout_pad, Wp);
        B200SP_LAUNCH_CHECK();
    }
    note_kernel("k_conv_tc");
    TCParams p{};
    p.in = in; p.Wp = Wp; p.tab = tab; p.orow = orow; p.rowmask = rowmask; p.pin = pin; p.pout = pout; p.pairnum = pairnum; p.out = out;
    p.n_rows = n_rows; p.pstride = pstride; p.Cin = Cin; p.Cout = Cout; p.K = K;
    p.nchunks = pl.nchunks; p.Cout_pad = pl.Cout_pad; p.accumulate = accumulate; p.pairs_mode = pairs_mode;
    p.res = res ? res : (accumulate ? out : nullptr);
    if (((uintptr_t)p.res & 15) != 0) return B200SP_EUNSUP;  // the epilogue reads it as float4
    p.nslots = pl.nslots; p.nslots_b = pl.nslots_b; p.stageB_bytes = pl.stageB; p.tmem_cols = pl.tmem_cols; p.ni = pl.ni;
    p.meta_bufs = pl.meta_bufs;
    {
        B200SP_ENV_INT(env_dbg, "B200SP_TC_DEBUG", 0);
        p.dbg = env_dbg;
    }
    p.row_tiles = (int)cdiv(n_rows, TC_BM);
    // few row tiles (deep U-Net levels): deal each tile's active offsets to nsplit CTAs so the machine is not idle
    // behind two or three serial tiles; the partial tiles are added in split order through the cluster's shared memory
    p.nsplit = 1;
    if (!pairs_mode && tab && K > 1 && p.row_tiles * 2 <= num_sms()) {
        B200SP_ENV_INT(env_nosplit, "B200SP_TC_NOSPLIT", 0);
        // up to 8 ways = the portable cluster size (B200SP_TC_MAXSPLIT: dev knob)
        B200SP_ENV_INT(env_maxsplit, "B200SP_TC_MAXSPLIT", 8);
        if (!env_nosplit) p.nsplit = std::max(1, std::min(std::min(K, env_maxsplit), 2 * num_sms() / p.row_tiles));
    }
    p.total_tiles = p.row_tiles * (pairs_mode ? K : p.nsplit);  // split mode: finalised by launch_tc (residency)
    if (pl.KC == 32) return launch_tc<32>(p, KT, n_in, st);
    if (pl.KC == 16) return launch_tc<16>(p, KT, n_in, st);
    return launch_tc<8>(p, KT, n_in, st);
}

// rows of desc_host: {W ptr, Wp ptr, K, Ci_w, Co_w, wflags}; desc_dev: device scratch of n * sizeof(PrepItem) bytes
int conv_tc_prep_batch(const int64_t* desc_host, int n, void* desc_dev, int64_t desc_dev_bytes, cudaStream_t st) {
    if (n <= 0) return B200SP_OK;
    if (desc_dev_bytes < (int64_t)n * (int64_t)sizeof(PrepItem)) {
        set_error("prep_weights_batch: descriptor scratch too small");
        return B200SP_ENOMEM;
    }
    std::vector<PrepItem> items((size_t)n);
    int64_t max_total = 1;
    for (int i = 0; i < n; ++i) {
        const int64_t* d = desc_host + (size_t)i * 6;
        PrepItem& it = items[(size_t)i];
        it.W = reinterpret_cast<const float*>(d[0]);
        it.Wp = reinterpret_cast<float*>(d[1]);
        it.K = (int)d[2]; it.Ci_w = (int)d[3]; it.Co_w = (int)d[4];
        it.transposed = (int)(d[5] & 1); it.mirror = (int)((d[5] >> 1) & 1);
        const int Cin = it.transposed ? it.Co_w : it.Ci_w, Cout = it.transposed ? it.Ci_w : it.Co_w;
        TCPlan pl;
        if (!tc_plan(it.K, Cin, Cout, it.K, pl)) {
            set_error("prep_weights_batch: shape K=%d Cin=%d Cout=%d is not covered by the tensor path", it.K, Cin, Cout);
            return B200SP_EUNSUP;
        }
        it.KC = pl.KC; it.nchunks = pl.nchunks; it.Cout_pad = pl.Cout_pad;
        max_total = std::max<int64_t>(max_total, (int64_t)it.K * pl.nchunks * pl.Cout_pad * pl.KC);
    }
    B200SP_CUDA(cudaMemcpyAsync(desc_dev, items.data(), (size_t)n * sizeof(PrepItem), cudaMemcpyHostToDevice, st));
    dim3 grid((unsigned)std::min<int64_t>(cdiv(max_total, 256), 64), (unsigned)n);
    k_prep_weights_batch<<<grid, 256, 0, st>>>(static_cast<const PrepItem*>(desc_dev));
    B200SP_LAUNCH_CHECK();
    return B200SP_OK;
}

int64_t conv_tc_ws_bytes(int K, int Cin, int Cout) {
    TCPlan pl;
    if (!tc_plan(K, Cin, Cout, K, pl)) return 0;
    return pl.wp_bytes;
}

}  // namespace b200sp
