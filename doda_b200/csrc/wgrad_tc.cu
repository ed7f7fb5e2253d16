// wgrad_tc.cu — sparse-conv weight gradient on the Blackwell tensor cores (tcgen05 + TMEM).
//
// Replaces the wgrad half of spconv v1.2 `indice_conv_backward` (per offset: two gather kernels -> cuBLAS SGEMM with
// the pair index as the contraction dimension):
//
//   dW[k][ci][co] += sum_i  a[pa[k][i]][ci] * b[pb[k][i]][co]          (pa/pb NULL: identity rows, the 1x1 conv)
//
// The contraction runs over PAIRS, so both operands are MN-major for the tensor core (channels contiguous in memory).
// kind::tf32 only transposes through the 32-byte-base swizzle, so the operands are split into bf16 pairs instead
// (v = hi + lo + O(2^-17 |v|)) and the product is hi*hi + lo*hi + hi*lo with fp32 accumulation (bf16x3: relative
// error ~1e-5 per term, far inside the 1e-4 bound the weight gradient is held to).
//
// One CTA owns offset k and a range of `ppb` pairs and walks it in stages of PS pairs:
//   * warps 8-11 (loaders): 16-byte cp.async gathers of the PS rows of `a` and `b` into fp32 staging rows;
//   * warps 0-7 (transform): fp32 staging -> bf16 hi / lo tiles in the un-swizzled canonical MN-major UMMA layout
//     [group of 8 pairs][chunk of 8 channels][8 pairs][16 B];
//   * warp 12, one thread: per 16 pairs three tcgen05.mma kind::f16, A = a-tile^T (M = 64 or 128 channel rows),
//     B = b-tile (N = Cb), accumulating the whole pair range in TMEM.  Channel rows beyond Ca alias whatever
//     follows in shared memory: they only produce accumulator rows that are never read;
//   * epilogue (warps 0-3): tcgen05.ld the Ca x Cb block and add it to dW[k] with vector atomics (one set per CTA).
#include "common.cuh"
#include "tc_common.cuh"
#include <cuda_bf16.h>
#include <algorithm>
#include <stdlib.h>

namespace b200sp {

using namespace tc;

constexpr int WG_XFORM = 256;
constexpr int WG_LOADERS = 128;
constexpr int WG_WARP_MMA = 12;    // warps 12..15: MMA issuers (one thread sustains only ~1 tcgen05.mma per ~215 cycles,
constexpr int WG_MAX_ISSUERS = 4;  // but issuers run concurrently: MMAs are dealt round-robin, one accumulator per issuer)
constexpr int WG_THREADS = WG_XFORM + WG_LOADERS + 32 * WG_MAX_ISSUERS;

struct WGTParams {
    const float* a;
    const float* b;
    const int* pa;       // [K][pstride] or NULL
    const int* pb;
    const int* pairnum;  // device [K] or NULL (identity: n_rows)
    float* dW;           // [K][Ca][Cb]
    int64_t pstride, n_rows;
    int Ca, Cb, K;
    int cprA, cprB;      // 16-byte fp32 chunks per staged row
    int rsA, rsB;        // staging row stride (bytes)
    int c8A, c8B;        // chunks of 8 channels
    int PS, psh, ppb, nslots;  // pairs per stage (power of two >= 16), log2(PS)
    int Mmma, MB, Npad;
    int dbg;             // dev only: 1 = skip the MMAs, 2 = skip the gathers, 4 = skip the transform
    int nacc;            // MMA issuer warps = independent TMEM accumulators (summed in the epilogue)
    uint32_t stgA, stgB;    // staging bytes per slot
    uint32_t tileA, tileB;  // bytes of ONE of {hi, lo}, slack for the aliased rows included
    uint32_t tmem_cols;
    uint32_t magicA, magicB;  // ceil(2^32 / cpr) for the item -> (row, chunk) split (0: cpr == 1)
};

__device__ __forceinline__ void wg_cp16(uint32_t dst, const void* src, int bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void wg_cp4(uint32_t dst, const void* src, int bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void wg_cp_arrive(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}

// 8 fp32 -> 8 bf16 hi + 8 bf16 lo (lo = bf16(v - hi))
__device__ __forceinline__ void split_bf16x8(const float* v, uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * i]), h1 = __float2bfloat16_rn(v[2 * i + 1]);
        const __nv_bfloat16 l0 = __float2bfloat16_rn(v[2 * i] - __bfloat162float(h0));
        const __nv_bfloat16 l1 = __float2bfloat16_rn(v[2 * i + 1] - __bfloat162float(h1));
        h[i] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
        l[i] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// staging rows [PS][rs] fp32 -> hi / lo tiles; lanes walk consecutive pairs (conflict-free on both sides)
__device__ __forceinline__ void wg_transform(const unsigned char* stg, int rs, int C, int c8, int PS, int psh,
                                             unsigned char* t_hi, unsigned char* t_lo, int tid) {
    const int items = PS * c8;
    const uint32_t gs = (uint32_t)c8 * 128u;
    for (int i = tid; i < items; i += WG_XFORM) {
        const int pr = i & (PS - 1), j8 = i >> psh;
        const float* src = reinterpret_cast<const float*>(stg + (size_t)pr * rs) + j8 * 8;
        float v[8];
        const float4 x = *reinterpret_cast<const float4*>(src);
        v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
        if (j8 * 8 + 4 < C) {
            const float4 y = *reinterpret_cast<const float4*>(src + 4);
            v[4] = y.x; v[5] = y.y; v[6] = y.z; v[7] = y.w;
        } else {
            v[4] = v[5] = v[6] = v[7] = 0.f;
        }
        uint4 hi, lo;
        split_bf16x8(v, hi, lo);
        const uint32_t off = (uint32_t)(pr >> 3) * gs + (uint32_t)j8 * 128u + (uint32_t)(pr & 7) * 16u;
        *reinterpret_cast<uint4*>(t_hi + off) = hi;
        *reinterpret_cast<uint4*>(t_lo + off) = lo;
    }
}

__device__ __forceinline__ void wg_gather(const float* base, int C, int cpr, uint32_t magic, int rs, bool vec,
                                          const int* idx, int PS, uint32_t stg, int lt) {
    const int items = PS * cpr;
#pragma unroll 4
    for (int i = lt; i < items; i += WG_LOADERS) {
        const int row = magic ? (int)__umulhi((unsigned)i, magic) : i;
        const int j = i - row * cpr;
        const int src = idx[row];
        const uint32_t dst = stg + (uint32_t)row * (uint32_t)rs + (uint32_t)j * 16u;
        if (vec) {
            const bool ok = src >= 0;
            wg_cp16(dst, ok ? base + (int64_t)src * C + 4 * j : base, ok ? 16 : 0);
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const bool ok = src >= 0 && 4 * j + e < C;
                wg_cp4(dst + 4u * e, ok ? base + (int64_t)src * C + 4 * j + e : base, ok ? 4 : 0);
            }
        }
    }
}

__global__ void __launch_bounds__(WG_THREADS) k_wgrad_tc(WGTParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int S = p.nslots;
    const int k = blockIdx.y;
    pdl_wait();
    const int64_t n = p.pairnum ? (int64_t)p.pairnum[k] : p.n_rows;
    const int64_t p0 = (int64_t)blockIdx.x * p.ppb;
    if (p0 >= n) return;
    const int np = (int)min((int64_t)p.ppb, n - p0);  // pairs of this CTA
    const int nst = (np + p.PS - 1) / p.PS;

    // slot: [stgA | stgB | A_hi | A_lo | B_hi | B_lo]
    const uint32_t slot_bytes = p.stgA + p.stgB + 2u * (p.tileA + p.tileB);
    unsigned char* s_slots = smem;
    int* s_ia = reinterpret_cast<int*>(smem + (size_t)S * slot_bytes);
    int* s_ib = s_ia + p.ppb;
    uint64_t* full = reinterpret_cast<uint64_t*>(s_ib + p.ppb);
    uint64_t* empty = full + S;
    uint64_t* raw = empty + S;
    uint64_t* accum = raw + S;
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(accum + 1);

    // pair lists of the whole range -> smem (coalesced)
    for (int i = tid; i < p.ppb; i += WG_THREADS) {
        const bool ok = i < np;
        s_ia[i] = ok ? (p.pa ? __ldg(p.pa + (int64_t)k * p.pstride + p0 + i) : (int)(p0 + i)) : -1;
        s_ib[i] = ok ? (p.pb ? __ldg(p.pb + (int64_t)k * p.pstride + p0 + i) : (int)(p0 + i)) : -1;
    }
    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full[s], WG_XFORM / 32);
            mbar_init(&empty[s], 1);
            mbar_init(&raw[s], WG_LOADERS);
        }
        mbar_init(accum, p.nacc);
        mbar_fence_init();
    }
    if (warp == WG_WARP_MMA) tmem_alloc(s_tmem, p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s_tmem;

    if (warp < WG_XFORM / 32) {
        // ========== transform: fp32 staging -> bf16 hi / lo tiles ==========
        int slot = 0;
        uint32_t ph = 0;
        for (int st = 0; st < nst; ++st) {
            unsigned char* sa = s_slots + (size_t)slot * slot_bytes;
            unsigned char* sb = sa + p.stgA;
            unsigned char* a_hi = sb + p.stgB;
            unsigned char* a_lo = a_hi + p.tileA;
            unsigned char* b_hi = a_lo + p.tileA;
            unsigned char* b_lo = b_hi + p.tileB;
            mbar_wait(&raw[slot], ph);
            if (!(p.dbg & 4)) {
                wg_transform(sa, p.rsA, p.Ca, p.c8A, p.PS, p.psh, a_hi, a_lo, tid);
                wg_transform(sb, p.rsB, p.Cb, p.c8B, p.PS, p.psh, b_hi, b_lo, tid);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[slot]);
            if (++slot == S) { slot = 0; ph ^= 1u; }
        }
        // ========== epilogue (warps 0-3): Ca x Cb block of TMEM -> vector atomics on dW[k] ==========
        if (warp < 4) {
            mbar_wait(accum, 0);
            tc_fence_after();
            const bool vec = (p.Cb % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.dW) & 15) == 0);
            const int nused = min(p.nacc, nst);
            for (int mb = 0; mb < p.MB; ++mb) {
                // accumulator row of this thread: M=128 -> lane 32*warp+lane; M=64 -> lanes 0..15 of each quarter
                const int m = p.Mmma == 128 ? warp * 32 + lane : (lane < 16 ? warp * 16 + lane : -1);
                const int ci = mb * 128 + m;
                for (int ch = 0; ch * 16 < p.Npad; ++ch) {
                    float v[16];
                    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(mb * p.Npad + ch * 16), v);
                    for (int ac = 1; ac < nused; ++ac) {
                        float w[16];
                        tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(ac * p.MB * p.Npad + mb * p.Npad + ch * 16), w);
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] += w[i];
                    }
                    if (m < 0 || ci >= p.Ca) continue;
                    float* o = p.dW + ((int64_t)k * p.Ca + ci) * p.Cb + ch * 16;
#pragma unroll
                    for (int g4 = 0; g4 < 4; ++g4) {
                        const int col = ch * 16 + g4 * 4;
                        if (col >= p.Cb) break;
                        if (vec) {
                            atomicAdd(reinterpret_cast<float4*>(o + g4 * 4),
                                      make_float4(v[g4 * 4], v[g4 * 4 + 1], v[g4 * 4 + 2], v[g4 * 4 + 3]));
                        } else {
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                if (col + e < p.Cb) atomicAdd(o + g4 * 4 + e, v[g4 * 4 + e]);
                        }
                    }
                }
            }
        }
    } else if (warp < (WG_XFORM + WG_LOADERS) / 32) {
        // ========== loaders ==========
        const int lt = tid - WG_XFORM;
        const bool vecA = (p.Ca % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.a) & 15) == 0);
        const bool vecB = (p.Cb % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.b) & 15) == 0);
        int slot = 0;
        uint32_t ph = 0;
        for (int st = 0; st < nst; ++st) {
            const uint32_t sa = smem_u32(s_slots + (size_t)slot * slot_bytes);
            const uint32_t sb = sa + p.stgA;
            mbar_wait(&empty[slot], ph ^ 1u);
            if (!(p.dbg & 2)) {
                wg_gather(p.a, p.Ca, p.cprA, p.magicA, p.rsA, vecA, s_ia + st * p.PS, p.PS, sa, lt);
                wg_gather(p.b, p.Cb, p.cprB, p.magicB, p.rsB, vecB, s_ib + st * p.PS, p.PS, sb, lt);
            }
            wg_cp_arrive(&raw[slot]);
            if (++slot == S) { slot = 0; ph ^= 1u; }
        }
        asm volatile("cp.async.wait_all;\n" ::: "memory");
    } else {
        // ========== MMA issuers: warp w owns the stages st = w, w+nacc, ... and accumulator w.  The whole warp walks
        // the loop (uniform control flow), one elected lane issues; descriptors are advanced by constant adds ==========
        const int w = warp - WG_WARP_MMA;
        if (w < p.nacc) {
            const uint32_t idesc = make_idesc_bf16(p.Mmma, p.Npad, 1, 1);  // both operands MN-major
            const uint32_t gsA = (uint32_t)p.c8A * 128u, gsB = (uint32_t)p.c8B * 128u;
            const int nk16 = (p.dbg & 1) ? 0 : (p.PS >> 4);
            const uint32_t acc_col = tmem + (uint32_t)(w * p.MB * p.Npad);
            // K = 16 pairs = two 8-pair groups (LBO = group stride); channel chunks are 128 B apart (SBO)
            const uint64_t dA0 = make_desc(smem_u32(s_slots) + p.stgA + p.stgB, gsA, 128u);
            const uint64_t dB0 = make_desc(smem_u32(s_slots) + p.stgA + p.stgB + 2u * p.tileA, gsB, 128u);
            const uint32_t slot16 = slot_bytes >> 4, loA16 = p.tileA >> 4, loB16 = p.tileB >> 4;
            const uint32_t kA16 = (2u * gsA) >> 4, kB16 = (2u * gsB) >> 4, mb16 = (16u * 128u) >> 4;  // 128 rows = 16 chunks
            int slot = w % S;
            uint32_t ph = (uint32_t)(w / S) & 1u;
            uint32_t acc = 0;
            for (int st = w; st < nst; st += p.nacc) {
                mbar_wait(&full[slot], ph);
                tc_fence_after();
                const uint64_t da = dA0 + (uint64_t)((uint32_t)slot * slot16);
                const uint64_t db = dB0 + (uint64_t)((uint32_t)slot * slot16);
                if (elect_one()) {
                    for (int t = 0; t < nk16; ++t) {
                        for (int mb = 0; mb < p.MB; ++mb) {
                            const uint64_t a = da + (uint32_t)t * kA16 + (uint32_t)mb * mb16;
                            const uint64_t b = db + (uint32_t)t * kB16;
                            const uint32_t d = acc_col + (uint32_t)(mb * p.Npad);
                            mma_bf16_ss(d, a, b, idesc, acc);
                            mma_bf16_ss(d, a + loA16, b, idesc, 1u);
                            mma_bf16_ss(d, a, b + loB16, idesc, 1u);
                        }
                        acc = 1u;
                    }
                    mma_commit(&empty[slot]);
                }
                acc = 1u;
                __syncwarp();
                slot += p.nacc;
                while (slot >= S) { slot -= S; ph ^= 1u; }
            }
            if (elect_one()) mma_commit(accum);
            __syncwarp();
        }
    }
    pdl_trigger();  // late trigger: see conv_tc.cu (CTAs that returned early count as triggered)
    tc_fence_before();
    __syncthreads();
    if (warp == WG_WARP_MMA) tmem_dealloc(tmem, p.tmem_cols);
}

// returns B200SP_EUNSUP when the shape is outside what the tensor path covers
int wgrad_tc_run(const float* a, int Ca, const float* b, int Cb, const int* pa, const int* pb, const int* pairnum,
                 int64_t n_upper, int K, int64_t pstride, float* dW, cudaStream_t st) {
    if (Ca < 1 || Cb < 1 || Ca > 256 || Cb > 256) return B200SP_EUNSUP;
    WGTParams p{};
    p.a = a; p.b = b; p.pa = pa; p.pb = pb; p.pairnum = (pa || pb) ? pairnum : nullptr; p.dW = dW;
    p.pstride = pstride; p.n_rows = n_upper; p.Ca = Ca; p.Cb = Cb; p.K = K;
    p.cprA = (Ca + 3) / 4; p.cprB = (Cb + 3) / 4;
    p.c8A = (Ca + 7) / 8; p.c8B = (Cb + 7) / 8;
    p.rsA = p.c8A * 32 + 16; p.rsB = p.c8B * 32 + 16;  // +16 B: consecutive rows start 4 (or 20) banks apart
    p.Mmma = Ca <= 64 ? 64 : 128;
    p.MB = p.Mmma == 64 ? 1 : (Ca + 127) / 128;
    p.Npad = (Cb + 15) / 16 * 16;
    if (p.MB * p.Npad > 512 || p.MB > 4) return B200SP_EUNSUP;
    {
        B200SP_ENV_INT(env_dbg, "B200SP_WG_DEBUG", 0);
        p.dbg = env_dbg;
    }
    p.magicA = p.cprA == 1 ? 0u : (uint32_t)(((1ull << 32) + p.cprA - 1) / p.cprA);
    p.magicB = p.cprB == 1 ? 0u : (uint32_t)(((1ull << 32) + p.cprB - 1) / p.cprB);
    // pairs per stage: biggest of {128, 64, 32, 16} with 4 slots under ~170 KB
    const int slots = 4;
    int PS = 128;
    {
        B200SP_ENV_INT(env_ps, "B200SP_WG_PS", 0);  // dev knob: cap on pairs per stage
        if (env_ps >= 16) PS = env_ps;
    }
    uint32_t slot = 0;
    for (; PS >= 16; PS >>= 1) {
        p.stgA = (uint32_t)PS * p.rsA;
        p.stgB = (uint32_t)PS * p.rsB;
        // slack: the M-padded operand reads up to Mmma*MB/8 (N: Npad/8) chunks from the start of the last pair group
        p.tileA = ((uint32_t)(PS / 8) * p.c8A * 128u + (uint32_t)(p.MB * p.Mmma / 8) * 128u + 127u) & ~127u;
        p.tileB = ((uint32_t)(PS / 8) * p.c8B * 128u + (uint32_t)(p.Npad / 8) * 128u + 127u) & ~127u;
        p.stgA = (p.stgA + 127u) & ~127u;
        p.stgB = (p.stgB + 127u) & ~127u;
        slot = p.stgA + p.stgB + 2u * (p.tileA + p.tileB);
        if ((uint64_t)slots * slot <= 170u * 1024u) break;
    }
    if (PS < 16) return B200SP_EUNSUP;
    p.PS = PS; p.nslots = slots;
    // issuer warps: each smem slot belongs to exactly one issuer (parity waits must not run a fill ahead of a shared
    // slot), so nacc divides nslots; one accumulator set (MB x Npad columns) per issuer
    {
        B200SP_ENV_INT(env_nacc, "B200SP_WG_NACC", WG_MAX_ISSUERS);
        const int want = std::max(1, std::min(env_nacc, WG_MAX_ISSUERS));
        p.nacc = 1;
        for (int ni = want; ni >= 1; --ni)
            if (slots % ni == 0 && ni * p.MB * p.Npad <= 512) {
                p.nacc = ni;
                break;
            }
    }
    uint32_t cols = 32;
    while (cols < (uint32_t)(p.nacc * p.MB * p.Npad)) cols <<= 1;
    p.tmem_cols = cols;
    p.psh = 0;
    while ((1 << p.psh) < PS) ++p.psh;
    // pairs per block: enough CTAs to fill the machine, few enough that the final atomics stay cheap
    int64_t ppb = 4096;
    {
        B200SP_ENV_INT(env_ppb, "B200SP_WG_PPB", 0);  // dev knob: cap on pairs per CTA
        if (env_ppb >= 256) ppb = env_ppb;
    }
    while (ppb > 2 * PS && ppb > 256 && cdiv(n_upper, ppb) * K < 3 * 148) ppb >>= 1;
    if (ppb < PS) ppb = PS;
    p.ppb = (int)ppb;
    const uint32_t smem = (uint32_t)slots * slot + (uint32_t)ppb * 8u + (uint32_t)(3 * slots + 1) * 8u + 32u;
    static uint32_t attr_smem = 0;
    if (smem > attr_smem) {
        B200SP_CUDA(cudaFuncSetAttribute(k_wgrad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_smem = smem;
    }
    note_kernel("k_wgrad_tc");
    dim3 grid((unsigned)cdiv(n_upper, ppb), (unsigned)K);
    B200SP_CUDA(launch_pdl(k_wgrad_tc, grid, dim3(WG_THREADS), smem, st, p));
    B200SP_LAUNCH_CHECK();
    return B200SP_OK;
}

}  // namespace b200sp
