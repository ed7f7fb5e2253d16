// elementwise.cu — HBM-bound row kernels: BatchNorm(+ReLU) on active sites, point<->voxel transfer,
// row gather / scatter-add.  All are streaming kernels over [rows, C] fp32 with float4 accesses.
//
// Replaces nn.BatchNorm1d/DSNorm + nn.ReLU (model/unet.py:28,43; model/dsnorm.py:79-84),
// voxelize_fp/bp (lib/pointgroup_ops/src/voxelize/voxelize.cu:10-52) and the devoxelize gather
// (model/unet.py:62) of the reference.
#include "common.cuh"

namespace b200sp {

constexpr int BN_THREADS = 256;
constexpr int BN_MAXGRID = 296;  // 2 CTAs x 148 SMs (the last block folds this many partial rows per channel)

// ---------------------------------------------------------------------------------------------
// per-channel reduction of up to two quantities over rows.  Layout: thread (rl, cg) with cg the float4
// column group; a block walks row chunks with a grid stride, then reduces over rl in shared memory and
// writes one partial row per block: partial[block][q][C].
// MODE 0: q0 = sum x, q1 = sum x^2                       (BN forward statistics)
// MODE 1: q0 = sum g, q1 = sum g*xhat,  g = dy * relu'   (BN backward)
// ---------------------------------------------------------------------------------------------
// everything the LAST block of the reduction needs to finish the layer's statistics in the same launch
struct BNFinal {
    unsigned* ticket;  // zero before the launch; the last block resets it
    int64_t M;
    const float* w;
    const float* b;
    float eps, momentum;
    float* mean;
    float* invstd;
    float* running_mean;
    float* running_var;
    long long* num_batches_tracked;
    float* scale;
    float* shift;
    float* dw;
    float* db;
    float* c_g;
    float* c_mean_g;
    float* c_mean_gx;
};

// the forward's fused shift, spelled once so that the forward's ReLU and every backward's gate test round identically
__device__ __forceinline__ float bn_shift(float b, float mu, float sc) { return fmaf(-mu, sc, b); }

// single-launch form for small layers (defined at the end of this file); *done = false when the layer does not qualify
int bn_cluster_fwd(const float* x, int64_t M, int C, int relu, float* y, const BNFinal& fin, cudaStream_t st, bool* done);
int bn_cluster_bwd(const float* x, const float* dy, int64_t M, int C, int relu, float* dx, const BNFinal& fin, const float* mean,
                   const float* invstd, cudaStream_t st, bool* done);

template <int VEC, int MODE>
__global__ void __launch_bounds__(BN_THREADS) k_bn_reduce(const float* __restrict__ x, const float* __restrict__ dy,
                                                          int64_t M, int C, const float* __restrict__ scale,
                                                          const float* __restrict__ shift,
                                                          const float* __restrict__ mean,
                                                          const float* __restrict__ invstd, int relu,
                                                          float* __restrict__ partial, BNFinal fin) {
    extern __shared__ float s_part[];  // [2][rpb][C]
    __shared__ int s_last;
    pdl_trigger();
    pdl_wait();
    const int CG = C / VEC;
    const int rpb = BN_THREADS / CG;
    const int tid = threadIdx.x;
    const int cg = tid % CG, rl = tid / CG;
    const bool active = rl < rpb;
    float a0[VEC], a1[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) a0[v] = a1[v] = 0.f;
    float sc[VEC], sh[VEC], mu[VEC], is[VEC];
    if (MODE == 0 && active) {
        // statistics are accumulated around row 0's values: var = E[(x-k)^2] - (E[x-k])^2 does not cancel
        // catastrophically when |mean| >> std (the single-pass E[x^2] - mean^2 loses ~(mean/std)^2 ulps)
#pragma unroll
        for (int v = 0; v < VEC; ++v) mu[v] = __ldg(x + cg * VEC + v);
    }
    if (MODE == 1 && active) {
        (void)scale;
        (void)shift;
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            int c = cg * VEC + v;
            mu[v] = mean[c];
            is[v] = invstd[c];
            sc[v] = (fin.w ? fin.w[c] : 1.f) * is[v];  // the forward's fused scale / shift, recomputed per block
            sh[v] = bn_shift(fin.b ? fin.b[c] : 0.f, mu[v], sc[v]);
        }
    }
    if (active) {
        // UNR rows per trip with all loads issued before the first use: a block is ~8 warps and the grid ~2 blocks
        // per SM, so memory-level parallelism has to come from inside the thread
        constexpr int UNR = 4;
        const int64_t stride = (int64_t)gridDim.x * rpb;
        for (int64_t r0 = (int64_t)blockIdx.x * rpb + rl; r0 < M; r0 += UNR * stride) {
            float xv[UNR][VEC], gv[UNR][VEC];
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const int64_t r = r0 + u * stride;
                if (r < M) {
                    if (VEC == 4) {
                        float4 t = __ldg(reinterpret_cast<const float4*>(x + r * C) + cg);
                        xv[u][0] = t.x; xv[u][1] = t.y; xv[u][2] = t.z; xv[u][3] = t.w;
                        if (MODE == 1) {
                            float4 g = __ldg(reinterpret_cast<const float4*>(dy + r * C) + cg);
                            gv[u][0] = g.x; gv[u][1] = g.y; gv[u][2] = g.z; gv[u][3] = g.w;
                        }
                    } else {
                        xv[u][0] = __ldg(x + r * C + cg);
                        if (MODE == 1) gv[u][0] = __ldg(dy + r * C + cg);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                if (r0 + u * stride >= M) break;
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    if (MODE == 0) {
                        const float d = xv[u][v] - mu[v];
                        a0[v] += d;
                        a1[v] = fmaf(d, d, a1[v]);
                    } else {
                        float g = gv[u][v];
                        if (relu && fmaf(xv[u][v], sc[v], sh[v]) <= 0.f) g = 0.f;
                        a0[v] += g;
                        a1[v] = fmaf(g, (xv[u][v] - mu[v]) * is[v], a1[v]);
                    }
                }
            }
        }
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            s_part[rl * C + cg * VEC + v] = a0[v];
            s_part[(rpb + rl) * C + cg * VEC + v] = a1[v];
        }
    }
    __syncthreads();
    for (int i = tid; i < 2 * C; i += BN_THREADS) {
        int q = i / C, c = i % C;
        float s = 0.f;
        for (int r = 0; r < rpb; ++r) s += s_part[(q * rpb + r) * C + c];
        partial[((int64_t)blockIdx.x * 2 + q) * C + c] = s;
    }
    // ---- last block to finish folds the per-block partials (fp64) and finalises ----
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(fin.ticket, 1u) == gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // all threads fold the G per-block partials: item (slice, entry) sums every SL-th partial row of one of the 2C
    // entries (coalesced across threads), slices meet in shared memory, then one thread per channel finalises
    __shared__ double s_fold[512];
    const int G = gridDim.x;
    const int E = 2 * C;
    // eight loads in flight per thread: with two, a 96-entry layer walked 74 dependent L2 round trips here and the
    // fold, not the reduction, set the kernel's duration (25 us for 27 k rows x 48 channels)
    auto fold = [&](int e, int first, int step) {
        double a[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        int g = first;
        for (; g + 7 * step < G; g += 8 * step) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __ldcg(&partial[(int64_t)(g + u * step) * E + e]);
#pragma unroll
            for (int u = 0; u < 8; ++u) a[u] += (double)v[u];
        }
        for (; g < G; g += step) a[0] += (double)__ldcg(&partial[(int64_t)g * E + e]);
        return ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
    };
    // narrow layers (E <= 256 entries): SL slices of the partial rows per entry so that all threads work, slices meet
    // in shared memory (E * SL <= 256 doubles).  Wide layers (C > 128, up to 1024): one thread per entry folds all
    // rows itself and nothing is staged -- s_fold is never indexed past E * SL <= 512.
    const bool sliced = E <= BN_THREADS;
    const int SL = sliced ? BN_THREADS / E : 1;
    if (sliced) {
        for (int i = tid; i < E * SL; i += BN_THREADS) {
            const int sl = i / E, e = i - sl * E;
            s_fold[i] = fold(e, sl, SL);
        }
    }
    __syncthreads();
    for (int c = tid; c < C; c += BN_THREADS) {
        double s0 = 0.0, s1 = 0.0;
        if (sliced) {
            for (int sl = 0; sl < SL; ++sl) {
                s0 += s_fold[sl * E + c];
                s1 += s_fold[sl * E + C + c];
            }
        } else {
            s0 = fold(c, 0, 1);
            s1 = fold(C + c, 0, 1);
        }
        if (MODE == 0) {
            const double dm = s0 / (double)fin.M;  // mean of (x - k), k = row 0's value of this channel
            const double mu = (double)__ldg(x + c) + dm;
            double var = s1 / (double)fin.M - dm * dm;
            if (var < 0.0) var = 0.0;
            const float is = (float)(1.0 / sqrt(var + (double)fin.eps));
            fin.mean[c] = (float)mu;
            fin.invstd[c] = is;
            // running statistics exactly as F.batch_norm: unbiased variance, exponential average
            if (fin.running_mean) fin.running_mean[c] = (1.f - fin.momentum) * fin.running_mean[c] + fin.momentum * (float)mu;
            if (fin.running_var) {
                const float vu = (float)(fin.M > 1 ? var * (double)fin.M / (double)(fin.M - 1) : var);
                fin.running_var[c] = (1.f - fin.momentum) * fin.running_var[c] + fin.momentum * vu;
            }
            const float wv = fin.w ? fin.w[c] : 1.f, bv = fin.b ? fin.b[c] : 0.f;
            const float scv = wv * is;
            fin.scale[c] = scv;
            fin.shift[c] = bn_shift(bv, (float)mu, scv);
        } else {
            if (fin.dw) fin.dw[c] = (float)s1;
            if (fin.db) fin.db[c] = (float)s0;
            const float wv = fin.w ? fin.w[c] : 1.f;
            const float isv = invstd[c];
            fin.c_g[c] = wv * isv;
            fin.c_mean_g[c] = (float)(s0 / (double)fin.M);
            fin.c_mean_gx[c] = (float)(s1 / (double)fin.M);
            const float scv = wv * isv;
            fin.scale[c] = scv;
            fin.shift[c] = bn_shift(fin.b ? fin.b[c] : 0.f, mean[c], scv);
        }
    }
    if (tid == 0) {
        *fin.ticket = 0u;  // ready for the next launch on this workspace
        if (MODE == 0 && fin.num_batches_tracked) *fin.num_batches_tracked += 1;
    }
}

// y = [relu](x*scale + shift)
template <int VEC>
__global__ void __launch_bounds__(256) k_affine_relu(const float* __restrict__ x, int64_t M, int C,
                                                     const float* __restrict__ scale,
                                                     const float* __restrict__ shift, int relu,
                                                     float* __restrict__ y) {
    pdl_trigger();
    pdl_wait();
    const int CG = C / VEC;
    const int64_t n = M * CG;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int cg = (int)(i % CG);
        if (VEC == 4) {
            float4 t = __ldg(reinterpret_cast<const float4*>(x) + i);
            float4 sc = __ldg(reinterpret_cast<const float4*>(scale) + cg);
            float4 sh = __ldg(reinterpret_cast<const float4*>(shift) + cg);
            float4 o;
            o.x = fmaf(t.x, sc.x, sh.x); o.y = fmaf(t.y, sc.y, sh.y);
            o.z = fmaf(t.z, sc.z, sh.z); o.w = fmaf(t.w, sc.w, sh.w);
            if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
            reinterpret_cast<float4*>(y)[i] = o;
        } else {
            float o = fmaf(__ldg(x + i), scale[cg], shift[cg]);
            y[i] = relu ? fmaxf(o, 0.f) : o;
        }
    }
}

// backward finalize: dw = sum g*xhat, db = sum g; coefficient rows for the apply pass
// dx = w*invstd * (g - mean(g) - xhat*mean(g*xhat))
template <int VEC>
__global__ void __launch_bounds__(256) k_bn_bwd_apply(const float* __restrict__ x, const float* __restrict__ dy,
                                                      int64_t M, int C, const float* __restrict__ scale,
                                                      const float* __restrict__ shift,
                                                      const float* __restrict__ mean,
                                                      const float* __restrict__ invstd,
                                                      const float* __restrict__ c_g,
                                                      const float* __restrict__ c_mean_g,
                                                      const float* __restrict__ c_mean_gx, int relu,
                                                      float* __restrict__ dx, const float* __restrict__ add) {
    pdl_trigger();
    pdl_wait();
    const int CG = C / VEC;
    const int64_t n = M * CG;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int cg = (int)(i % CG);
        float xv[VEC], gv[VEC], o[VEC];
        if (VEC == 4) {
            float4 t = __ldg(reinterpret_cast<const float4*>(x) + i);
            float4 g = __ldg(reinterpret_cast<const float4*>(dy) + i);
            xv[0] = t.x; xv[1] = t.y; xv[2] = t.z; xv[3] = t.w;
            gv[0] = g.x; gv[1] = g.y; gv[2] = g.z; gv[3] = g.w;
        } else {
            xv[0] = __ldg(x + i);
            gv[0] = __ldg(dy + i);
        }
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            int c = cg * VEC + v;
            float g = gv[v];
            if (relu && fmaf(xv[v], scale[c], shift[c]) <= 0.f) g = 0.f;
            float xhat = (xv[v] - mean[c]) * invstd[c];
            o[v] = __fmul_rn(c_g[c], g - c_mean_g[c] - xhat * c_mean_gx[c]);
        }
        if (add) {  // gradient of a second consumer of x (residual skip, U-Net skip): a separate rounded add, as if the
                    // sum were formed by a following kernel (no contraction with the product above)
            if (VEC == 4) {
                const float4 a4 = __ldg(reinterpret_cast<const float4*>(add) + i);
                o[0] = __fadd_rn(o[0], a4.x); o[1] = __fadd_rn(o[1], a4.y); o[2] = __fadd_rn(o[2], a4.z); o[3] = __fadd_rn(o[3], a4.w);
            } else {
                o[0] = __fadd_rn(o[0], __ldg(add + i));
            }
        }
        if (VEC == 4) reinterpret_cast<float4*>(dx)[i] = make_float4(o[0], o[1], o[2], o[3]);
        else dx[i] = o[0];
    }
}

// ---------------------------------------------------------------------------------------------
// point <-> voxel
// ---------------------------------------------------------------------------------------------
__global__ void k_voxelize_fp(const float* __restrict__ feats, float* __restrict__ out, const int* __restrict__ map,
                              int average, int64_t M, int A, int C) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * C) return;
    int64_t v = i / C;
    int c = (int)(i - v * C);
    const int* r = map + v * (A + 1);
    int cnt = r[0];
    float s = 0.f;
    for (int j = 1; j <= cnt; ++j) s += __ldg(feats + (int64_t)r[j] * C + c);
    float mult = (average && cnt > 0) ? 1.f / (float)cnt : 1.f;
    out[i] += mult * s;
}

__global__ void k_voxelize_bp(const float* __restrict__ dout, float* __restrict__ dfeats,
                              const int* __restrict__ map, int average, int64_t M, int A, int C) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * C) return;
    int64_t v = i / C;
    int c = (int)(i - v * C);
    const int* r = map + v * (A + 1);
    int cnt = r[0];
    float mult = (average && cnt > 0) ? 1.f / (float)cnt : 1.f;
    float g = mult * dout[i];
    for (int j = 1; j <= cnt; ++j) atomicAdd(dfeats + (int64_t)r[j] * C + c, g);
}

template <typename IdxT, int VEC>
__global__ void __launch_bounds__(256) k_gather_rows(const float* __restrict__ src, const IdxT* __restrict__ idx,
                                                     int64_t n, int C, float* __restrict__ out) {
    const int CG = C / VEC;
    const int64_t total = n * CG;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i / CG;
        int cg = (int)(i - r * CG);
        int64_t s = (int64_t)idx[r];
        if (VEC == 4)
            reinterpret_cast<float4*>(out)[i] = __ldg(reinterpret_cast<const float4*>(src + s * C) + cg);
        else
            out[i] = __ldg(src + s * C + cg);
    }
}

__device__ __forceinline__ void red_add_v4(float* addr, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}

template <typename IdxT, int VEC>
__global__ void __launch_bounds__(256) k_scatter_add_rows(const float* __restrict__ src,
                                                          const IdxT* __restrict__ idx, int64_t n, int C,
                                                          float* __restrict__ dst) {
    const int CG = C / VEC;
    const int64_t total = n * CG;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i / CG;
        int cg = (int)(i - r * CG);
        int64_t d = (int64_t)idx[r];
        if (VEC == 4) red_add_v4(dst + d * C + cg * 4, __ldg(reinterpret_cast<const float4*>(src) + i));
        else atomicAdd(dst + d * C + cg, __ldg(src + i));
    }
}

static inline unsigned stream_grid(int64_t n_items, int threads) {
    int64_t g = cdiv(n_items, threads);
    int64_t cap = (int64_t)148 * 16;
    return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

static inline bool vec4_ok(int C, const void* a, const void* b = nullptr, const void* c = nullptr) {
    auto al = [](const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    return C % 4 == 0 && al(a) && al(b) && al(c);
}

}  // namespace b200sp

using namespace b200sp;

extern "C" int64_t b200sp_bn_ws_bytes(int64_t M, int C) {
    (void)M;
    // ticket (first 256 bytes, same place for every C) | partials | scale, shift, c_g, c_mean_g, c_mean_gx.  The caller
    // zero-fills the workspace ONCE; every launch leaves the ticket at zero again
    return (int64_t)sizeof(float) * ((int64_t)BN_MAXGRID * 2 * C + 8 * (int64_t)C) + 1024 + 256;
}

static int bn_grid(int64_t M, int C, int VEC) {
    int CG = C / VEC;
    int rpb = BN_THREADS / CG;
    int64_t g = cdiv(M, (int64_t)rpb * 16);  // >= 16 rows per thread: fewer partial rows for the last block to fold
    if (g < 1) g = 1;
    if (g > BN_MAXGRID) g = BN_MAXGRID;
    return (int)g;
}

extern "C" int b200sp_bn_fwd_train(const float* x, int64_t M, int C, const float* w, const float* b, float eps,
                                   int relu, float* y, float* mean, float* invstd, float* running_mean,
                                   float* running_var, float momentum, int64_t* num_batches_tracked, void* ws,
                                   int64_t ws_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    B200SP_CHECK_ARG(M >= 1 && C >= 1, "bn_fwd_train: need M>=1, C>=1");
    B200SP_CHECK_ARG(ws_bytes >= b200sp_bn_ws_bytes(M, C), "bn_fwd_train: workspace too small");
    const bool v4 = vec4_ok(C, x, y) && C / 4 <= BN_THREADS;
    B200SP_CHECK_ARG(v4 || C <= BN_THREADS, "bn_fwd_train: C=%d unsupported", C);
    float* partial = (float*)ws + 64;
    float* scale = partial + (int64_t)BN_MAXGRID * 2 * C;
    float* shift = scale + C;
    int VEC = v4 ? 4 : 1;
    int G = bn_grid(M, C, VEC);
    int rpb = BN_THREADS / (C / VEC);
    size_t smem = sizeof(float) * 2 * rpb * C;
    BNFinal fin{};
    fin.ticket = reinterpret_cast<unsigned*>(ws);
    fin.M = M; fin.w = w; fin.b = b; fin.eps = eps; fin.momentum = momentum; fin.mean = mean; fin.invstd = invstd;
    fin.running_mean = running_mean; fin.running_var = running_var;
    fin.num_batches_tracked = (long long*)num_batches_tracked; fin.scale = scale; fin.shift = shift;
    if (v4) {
        bool done = false;
        int rc = bn_cluster_fwd(x, M, C, relu, y, fin, st, &done);
        if (rc != B200SP_OK || done) return rc;
    }
    if (v4)
        B200SP_CUDA(launch_pdl(k_bn_reduce<4, 0>, dim3(G), dim3(BN_THREADS), smem, st, x, nullptr, M, C, nullptr, nullptr, nullptr, nullptr, 0, partial, fin));
    else
        B200SP_CUDA(launch_pdl(k_bn_reduce<1, 0>, dim3(G), dim3(BN_THREADS), smem, st, x, nullptr, M, C, nullptr, nullptr, nullptr, nullptr, 0, partial, fin));
    if (v4)
        B200SP_CUDA(launch_pdl(k_affine_relu<4>, dim3(stream_grid(M * (C / 4), 256)), dim3(256), 0, st, x, M, C, scale, shift, relu, y));
    else
        B200SP_CUDA(launch_pdl(k_affine_relu<1>, dim3(stream_grid(M * C, 256)), dim3(256), 0, st, x, M, C, scale, shift, relu, y));
    B200SP_LAUNCH_CHECK_N(2);
    return B200SP_OK;
}

extern "C" int b200sp_affine_relu(const float* x, int64_t M, int C, const float* scale, const float* shift, int relu,
                                  float* y, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    B200SP_CHECK_ARG(M >= 0 && C >= 1, "affine_relu: bad sizes");
    if (M == 0) return B200SP_OK;
    if (vec4_ok(C, x, y, scale) && vec4_ok(C, shift))
        B200SP_CUDA(launch_pdl(k_affine_relu<4>, dim3(stream_grid(M * (C / 4), 256)), dim3(256), 0, st, x, M, C, scale, shift, relu, y));
    else
        B200SP_CUDA(launch_pdl(k_affine_relu<1>, dim3(stream_grid(M * C, 256)), dim3(256), 0, st, x, M, C, scale, shift, relu, y));
    B200SP_LAUNCH_CHECK();
    return B200SP_OK;
}

extern "C" int b200sp_bn_bwd_add(const float* x, const float* dy, int64_t M, int C, const float* w, const float* b,
                                 const float* mean, const float* invstd, int relu, float* dx, float* dw, float* db,
                                 const float* add, void* ws, int64_t ws_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    B200SP_CHECK_ARG(M >= 1 && C >= 1, "bn_bwd: need M>=1, C>=1");
    B200SP_CHECK_ARG(!add || ((uintptr_t)add & 15) == 0 || !(vec4_ok(C, x, dy, dx) && C / 4 <= BN_THREADS), "bn_bwd: the added gradient must be 16-byte aligned");
    B200SP_CHECK_ARG(ws_bytes >= b200sp_bn_ws_bytes(M, C), "bn_bwd: workspace too small");
    const bool v4 = vec4_ok(C, x, dy, dx) && C / 4 <= BN_THREADS;
    B200SP_CHECK_ARG(v4 || C <= BN_THREADS, "bn_bwd: C=%d unsupported", C);
    float* partial = (float*)ws + 64;
    float* scale = partial + (int64_t)BN_MAXGRID * 2 * C;
    float* shift = scale + C;
    float* c_g = shift + C;
    float* c_mg = c_g + C;
    float* c_mgx = c_mg + C;
    int VEC = v4 ? 4 : 1;
    int G = bn_grid(M, C, VEC);
    int rpb = BN_THREADS / (C / VEC);
    size_t smem = sizeof(float) * 2 * rpb * C;
    BNFinal fin{};
    fin.ticket = reinterpret_cast<unsigned*>(ws);
    fin.M = M; fin.w = w; fin.b = b; fin.scale = scale; fin.shift = shift; fin.dw = dw; fin.db = db;
    fin.c_g = c_g; fin.c_mean_g = c_mg; fin.c_mean_gx = c_mgx;
    if (v4 && !add) {
        bool done = false;
        int rc = bn_cluster_bwd(x, dy, M, C, relu, dx, fin, mean, invstd, st, &done);
        if (rc != B200SP_OK || done) return rc;
    }
    if (v4)
        B200SP_CUDA(launch_pdl(k_bn_reduce<4, 1>, dim3(G), dim3(BN_THREADS), smem, st, x, dy, M, C, nullptr, nullptr, mean, invstd, relu, partial, fin));
    else
        B200SP_CUDA(launch_pdl(k_bn_reduce<1, 1>, dim3(G), dim3(BN_THREADS), smem, st, x, dy, M, C, nullptr, nullptr, mean, invstd, relu, partial, fin));
    if (v4)
        B200SP_CUDA(launch_pdl(k_bn_bwd_apply<4>, dim3(stream_grid(M * (C / 4), 256)), dim3(256), 0, st, x, dy, M, C, scale, shift,
                               mean, invstd, c_g, c_mg, c_mgx, relu, dx, v4 ? add : nullptr));
    else
        B200SP_CUDA(launch_pdl(k_bn_bwd_apply<1>, dim3(stream_grid(M * C, 256)), dim3(256), 0, st, x, dy, M, C, scale, shift, mean,
                               invstd, c_g, c_mg, c_mgx, relu, dx, v4 ? nullptr : add));
    B200SP_LAUNCH_CHECK_N(2);
    return B200SP_OK;
}

extern "C" int b200sp_bn_bwd(const float* x, const float* dy, int64_t M, int C, const float* w, const float* b,
                             const float* mean, const float* invstd, int relu, float* dx, float* dw, float* db,
                             void* ws, int64_t ws_bytes, void* stream) {
    return b200sp_bn_bwd_add(x, dy, M, C, w, b, mean, invstd, relu, dx, dw, db, nullptr, ws, ws_bytes, stream);
}

extern "C" int b200sp_copy_cols(const float* src, int64_t rows, int src_stride, int src_col0, int ncols, float* dst,
                                int dst_stride, int dst_col0, void* stream) {
    B200SP_CHECK_ARG(rows >= 0 && ncols >= 1 && src_col0 >= 0 && dst_col0 >= 0 && src_col0 + ncols <= src_stride &&
                         dst_col0 + ncols <= dst_stride,
                     "copy_cols: column range outside the row");
    if (rows == 0) return B200SP_OK;
    B200SP_CUDA(cudaMemcpy2DAsync(dst + dst_col0, (size_t)dst_stride * sizeof(float), src + src_col0, (size_t)src_stride * sizeof(float),
                                  (size_t)ncols * sizeof(float), (size_t)rows, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return B200SP_OK;
}

extern "C" int b200sp_voxelize_fp(const float* feats, float* out, const int32_t* map, int average, int64_t M, int A,
                                  int C, void* stream) {
    B200SP_CHECK_ARG(M >= 0 && A >= 0 && C >= 1, "voxelize_fp: bad sizes");
    if (M == 0) return B200SP_OK;
    k_voxelize_fp<<<(unsigned)cdiv(M * C, 256), 256, 0, (cudaStream_t)stream>>>(feats, out, map, average, M, A, C);
    B200SP_LAUNCH_CHECK();
    return B200SP_OK;
}

extern "C" int b200sp_voxelize_bp(const float* dout, float* dfeats, const int32_t* map, int average, int64_t M,
                                  int A, int C, void* stream) {
    B200SP_CHECK_ARG(M >= 0 && A >= 0 && C >= 1, "voxelize_bp: bad sizes");
    if (M == 0) return B200SP_OK;
    k_voxelize_bp<<<(unsigned)cdiv(M * C, 256), 256, 0, (cudaStream_t)stream>>>(dout, dfeats, map, average, M, A, C);
    B200SP_LAUNCH_CHECK();
    return B200SP_OK;
}

extern "C" int b200sp_gather_rows(const float* src, const void* idx, int idx_is_i64, int64_t n, int C, float* out,
                                  void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    B200SP_CHECK_ARG(n >= 0 && C >= 1, "gather_rows: bad sizes");
    if (n == 0) return B200SP_OK;
    bool v4 = vec4_ok(C, src, out);
    unsigned grid = stream_grid(n * (v4 ? C / 4 : C), 256);
    if (idx_is_i64) {
        if (v4) k_gather_rows<int64_t, 4><<<grid, 256, 0, st>>>(src, (const int64_t*)idx, n, C, out);
        else k_gather_rows<int64_t, 1><<<grid, 256, 0, st>>>(src, (const int64_t*)idx, n, C, out);
    } else {
        if (v4) k_gather_rows<int32_t, 4><<<grid, 256, 0, st>>>(src, (const int32_t*)idx, n, C, out);
        else k_gather_rows<int32_t, 1><<<grid, 256, 0, st>>>(src, (const int32_t*)idx, n, C, out);
    }
    B200SP_LAUNCH_CHECK();
    return B200SP_OK;
}

extern "C" int b200sp_scatter_add_rows(const float* src, const void* idx, int idx_is_i64, int64_t n, int C,
                                       float* dst, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    B200SP_CHECK_ARG(n >= 0 && C >= 1, "scatter_add_rows: bad sizes");
    if (n == 0) return B200SP_OK;
    bool v4 = vec4_ok(C, src, dst);
    unsigned grid = stream_grid(n * (v4 ? C / 4 : C), 256);
    if (idx_is_i64) {
        if (v4) k_scatter_add_rows<int64_t, 4><<<grid, 256, 0, st>>>(src, (const int64_t*)idx, n, C, dst);
        else k_scatter_add_rows<int64_t, 1><<<grid, 256, 0, st>>>(src, (const int64_t*)idx, n, C, dst);
    } else {
        if (v4) k_scatter_add_rows<int32_t, 4><<<grid, 256, 0, st>>>(src, (const int32_t*)idx, n, C, dst);
        else k_scatter_add_rows<int32_t, 1><<<grid, 256, 0, st>>>(src, (const int32_t*)idx, n, C, dst);
    }
    B200SP_LAUNCH_CHECK();
    return B200SP_OK;
}

// ---------------------------------------------------------------------------------------------
// BatchNorm(+ReLU) for SMALL layers in ONE launch: the rows are dealt to the CTAs of one thread-block cluster, every
// CTA keeps its slice in shared memory, the per-channel partial sums meet through distributed shared memory (added in
// rank order: deterministic), and the normalised rows are written from shared memory -- x (and dy) cross L2 once
// instead of twice and the deep U-Net levels (45 .. 6 k rows: a 7-12 us floor per launch in the two-kernel form) pay
// one launch instead of two.
// ---------------------------------------------------------------------------------------------
namespace b200sp {

constexpr int BNC_THREADS = 256;
constexpr int BNC_MAX_CLUSTER = 16;  // > 8 is the opt-in (non-portable) cluster size; a B200 GPC holds it

__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ float ld_dsmem(const float* local_smem_ptr, int rank) {
    uint32_t laddr = (uint32_t)__cvta_generic_to_shared(local_smem_ptr), raddr;
    float v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(raddr) : "r"(laddr), "r"(rank));
    asm volatile("ld.shared::cluster.f32 %0, [%1];\n" : "=f"(v) : "r"(raddr) : "memory");
    return v;
}

// MODE 0: forward (statistics + apply);  MODE 1: backward (sum g, sum g*xhat, dx)
// smem: xs [rows_cta][C] (| gs [rows_cta][C] for MODE 1) | part [2][C] | coef [4][C]
template <int MODE>
__global__ void __launch_bounds__(BNC_THREADS) k_bn_cluster(const float* __restrict__ x, const float* __restrict__ dy,
                                                            int64_t M, int C, int rows_cta, int relu, float* __restrict__ out,
                                                            BNFinal fin, const float* __restrict__ mean_in,
                                                            const float* __restrict__ invstd_in) {
    extern __shared__ __align__(16) float s_bnc[];
    pdl_trigger();
    pdl_wait();
    unsigned rank, nranks;
    asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(rank));
    asm volatile("mov.u32 %0, %%cluster_nctarank;\n" : "=r"(nranks));
    const int tid = threadIdx.x;
    const int C4 = C >> 2;
    const int64_t r0 = (int64_t)rank * rows_cta;
    const int rows = (int)max((int64_t)0, min((int64_t)rows_cta, M - r0));
    float* xs = s_bnc;
    float* gs = xs + (size_t)rows_cta * C;
    float* part = (MODE == 1 ? gs + (size_t)rows_cta * C : gs);
    float* coef = part + 2 * C;
    // ---- load the slice (coalesced float4) ----
    const int n4 = rows * C4;
#pragma unroll 4
    for (int i = tid; i < n4; i += BNC_THREADS) {
        reinterpret_cast<float4*>(xs)[i] = __ldg(reinterpret_cast<const float4*>(x + r0 * C) + i);
        if (MODE == 1) reinterpret_cast<float4*>(gs)[i] = __ldg(reinterpret_cast<const float4*>(dy + r0 * C) + i);
    }
    __syncthreads();
    // ---- per-channel partial sums over this CTA's rows: thread (rl, c) with c the channel; rows strided ----
    const int rpb = BNC_THREADS / C > 0 ? BNC_THREADS / C : 1;  // row lanes per channel (C <= 256)
    const int c = tid % C, rl = tid / C;
    float a0 = 0.f, a1 = 0.f;
    float shift = 0.f, mu = 0.f, is = 0.f, sc = 0.f, sh = 0.f;
    if (rl < rpb && tid < rpb * C) {
        if (MODE == 0) {
            shift = __ldg(x + c);  // row 0 of the whole layer: the same shift in every CTA
            for (int r = rl; r < rows; r += rpb) {
                const float d = xs[r * C + c] - shift;
                a0 += d;
                a1 = fmaf(d, d, a1);
            }
        } else {
            mu = mean_in[c];
            is = invstd_in[c];
            sc = (fin.w ? fin.w[c] : 1.f) * is;
            sh = bn_shift(fin.b ? fin.b[c] : 0.f, mu, sc);
            for (int r = rl; r < rows; r += rpb) {
                const float xv = xs[r * C + c];
                float g = gs[r * C + c];
                if (relu && fmaf(xv, sc, sh) <= 0.f) g = 0.f;
                gs[r * C + c] = g;  // masked gradient, reused by the apply pass
                a0 += g;
                a1 = fmaf(g, (xv - mu) * is, a1);
            }
        }
    }
    // the row lanes of one channel meet in shared memory
    __shared__ float s_red[2 * BNC_THREADS];
    s_red[tid] = a0;
    s_red[BNC_THREADS + tid] = a1;
    __syncthreads();
    if (tid < C) {
        float s0 = 0.f, s1 = 0.f;
        for (int l = 0; l < rpb; ++l) {
            s0 += s_red[l * C + tid];
            s1 += s_red[BNC_THREADS + l * C + tid];
        }
        part[tid] = s0;
        part[C + tid] = s1;
    }
    cluster_sync_all();
    // ---- every CTA adds the partials of all ranks in rank order (fp64) and derives the coefficients ----
    if (tid < C) {
        double s0 = 0.0, s1 = 0.0;
        for (unsigned rk = 0; rk < nranks; ++rk) {
            s0 += (double)ld_dsmem(part + tid, (int)rk);
            s1 += (double)ld_dsmem(part + C + tid, (int)rk);
        }
        if (MODE == 0) {
            const double dm = s0 / (double)M;
            const double mud = (double)__ldg(x + tid) + dm;
            double var = s1 / (double)M - dm * dm;
            if (var < 0.0) var = 0.0;
            const float isv = (float)(1.0 / sqrt(var + (double)fin.eps));
            const float wv = fin.w ? fin.w[tid] : 1.f, bv = fin.b ? fin.b[tid] : 0.f;
            const float scv = wv * isv;
            coef[tid] = scv;
            coef[C + tid] = bn_shift(bv, (float)mud, scv);
            if (rank == 0) {
                if (fin.scale) fin.scale[tid] = scv;
                if (fin.shift) fin.shift[tid] = coef[C + tid];
                fin.mean[tid] = (float)mud;
                fin.invstd[tid] = isv;
                if (fin.running_mean) fin.running_mean[tid] = (1.f - fin.momentum) * fin.running_mean[tid] + fin.momentum * (float)mud;
                if (fin.running_var) {
                    const float vu = (float)(M > 1 ? var * (double)M / (double)(M - 1) : var);
                    fin.running_var[tid] = (1.f - fin.momentum) * fin.running_var[tid] + fin.momentum * vu;
                }
            }
        } else {
            const float wv = fin.w ? fin.w[tid] : 1.f;
            const float isv = invstd_in[tid];
            coef[tid] = wv * isv;                       // c_g
            coef[C + tid] = (float)(s0 / (double)M);    // mean g
            coef[2 * C + tid] = (float)(s1 / (double)M);  // mean g*xhat
            coef[3 * C + tid] = mean_in[tid];
            if (rank == 0) {
                if (fin.dw) fin.dw[tid] = (float)s1;
                if (fin.db) fin.db[tid] = (float)s0;
            }
        }
    }
    if (MODE == 0 && rank == 0 && tid == 0 && fin.num_batches_tracked) *fin.num_batches_tracked += 1;
    __syncthreads();
    // ---- apply from shared memory ----
    for (int i = tid; i < n4; i += BNC_THREADS) {
        const int c4 = (i % C4) * 4;
        const float4 xv = reinterpret_cast<const float4*>(xs)[i];
        float4 o;
        if (MODE == 0) {
            o.x = fmaf(xv.x, coef[c4], coef[C + c4]);
            o.y = fmaf(xv.y, coef[c4 + 1], coef[C + c4 + 1]);
            o.z = fmaf(xv.z, coef[c4 + 2], coef[C + c4 + 2]);
            o.w = fmaf(xv.w, coef[c4 + 3], coef[C + c4 + 3]);
            if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
        } else {
            const float4 g = reinterpret_cast<const float4*>(gs)[i];
            const float xa[4] = {xv.x, xv.y, xv.z, xv.w}, ga[4] = {g.x, g.y, g.z, g.w};
            float r[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int cc = c4 + e;
                const float xhat = (xa[e] - coef[3 * C + cc]) * invstd_in[cc];
                r[e] = coef[cc] * (ga[e] - coef[C + cc] - xhat * coef[2 * C + cc]);
            }
            o = make_float4(r[0], r[1], r[2], r[3]);
        }
        reinterpret_cast<float4*>(out + r0 * C)[i] = o;
    }
    cluster_sync_all();  // nobody exits while a peer may still read its partial sums
}

// cluster size and rows per CTA for a small layer, or 0 when the layer does not fit / is not worth it
static int bn_cluster_plan(int64_t M, int C, int mode, int* rows_cta, size_t* smem) {
    B200SP_ENV_INT(env_max, "B200SP_BN_CLUSTER", 0);        // largest cluster (0 = never; 16 = opt-in non-portable size)
    B200SP_ENV_INT(env_kb, "B200SP_BN_CLUSTER_KB", 64);     // largest slice per CTA
    if (env_max <= 0 || C % 4 != 0 || C > BNC_THREADS || M < 1) return 0;
    const size_t per_row = (size_t)C * 4u * (mode == 1 ? 2u : 1u);
    const size_t fixed = (size_t)6 * C * 4u + 64u;
    // one SM streams its slice at ~0.15 TB/s: past ~64 KB per CTA the two-kernel form (whole-GPU grids) is faster
    // (measured: 6149 x 64 in 8 slices of 196 KB, 16.6 us against 13.4 us)
    const size_t cap = (size_t)(env_kb > 200 ? 200 : env_kb) * 1024u;
    const int nmax = env_max > BNC_MAX_CLUSTER ? BNC_MAX_CLUSTER : env_max;
    for (int n = 1; n <= nmax; n <<= 1) {
        const int64_t rc = (M + n - 1) / n;
        const size_t need = (size_t)rc * per_row + fixed;
        if (need <= cap && (need <= 32u * 1024u || n * 2 > nmax)) {
            *rows_cta = (int)rc;
            *smem = need;
            return n;
        }
    }
    return 0;
}

template <int MODE>
static int bn_cluster_launch(int nclu, int rows_cta, size_t smem, const float* x, const float* dy, int64_t M, int C, int relu,
                             float* out, const BNFinal& fin, const float* mean, const float* invstd, cudaStream_t st) {
    static size_t attr_smem = 0;
    if (smem > attr_smem) {
        B200SP_CUDA(cudaFuncSetAttribute(k_bn_cluster<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024)));
        B200SP_CUDA(cudaFuncSetAttribute(k_bn_cluster<MODE>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        attr_smem = 200 * 1024;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)nclu);
    cfg.blockDim = dim3(BNC_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)nclu;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    B200SP_CUDA(cudaLaunchKernelEx(&cfg, k_bn_cluster<MODE>, x, dy, M, C, rows_cta, relu, out, fin, mean, invstd));
    B200SP_LAUNCH_CHECK();
    return B200SP_OK;
}

int bn_cluster_fwd(const float* x, int64_t M, int C, int relu, float* y, const BNFinal& fin, cudaStream_t st, bool* done) {
    int rows_cta = 0;
    size_t smem = 0;
    const int n = bn_cluster_plan(M, C, 0, &rows_cta, &smem);
    *done = n > 0;
    if (!n) return B200SP_OK;
    return bn_cluster_launch<0>(n, rows_cta, smem, x, nullptr, M, C, relu, y, fin, nullptr, nullptr, st);
}
int bn_cluster_bwd(const float* x, const float* dy, int64_t M, int C, int relu, float* dx, const BNFinal& fin, const float* mean,
                   const float* invstd, cudaStream_t st, bool* done) {
    int rows_cta = 0;
    size_t smem = 0;
    const int n = bn_cluster_plan(M, C, 1, &rows_cta, &smem);
    *done = n > 0;
    if (!n) return B200SP_OK;
    return bn_cluster_launch<1>(n, rows_cta, smem, x, dy, M, C, relu, dx, fin, mean, invstd, st);
}

}  // namespace b200sp
