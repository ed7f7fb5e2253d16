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
            sh[v] = (fin.b ? fin.b[c] : 0.f) - mu[v] * sc[v];
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
            fin.shift[c] = bv - (float)mu * scv;
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
            fin.shift[c] = (fin.b ? fin.b[c] : 0.f) - mean[c] * scv;
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
                                                      float* __restrict__ dx) {
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
            o[v] = c_g[c] * (g - c_mean_g[c] - xhat * c_mean_gx[c]);
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

extern "C" int b200sp_bn_bwd(const float* x, const float* dy, int64_t M, int C, const float* w, const float* b,
                             const float* mean, const float* invstd, int relu, float* dx, float* dw, float* db,
                             void* ws, int64_t ws_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    B200SP_CHECK_ARG(M >= 1 && C >= 1, "bn_bwd: need M>=1, C>=1");
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
    if (v4)
        B200SP_CUDA(launch_pdl(k_bn_reduce<4, 1>, dim3(G), dim3(BN_THREADS), smem, st, x, dy, M, C, nullptr, nullptr, mean, invstd, relu, partial, fin));
    else
        B200SP_CUDA(launch_pdl(k_bn_reduce<1, 1>, dim3(G), dim3(BN_THREADS), smem, st, x, dy, M, C, nullptr, nullptr, mean, invstd, relu, partial, fin));
    if (v4)
        B200SP_CUDA(launch_pdl(k_bn_bwd_apply<4>, dim3(stream_grid(M * (C / 4), 256)), dim3(256), 0, st, x, dy, M, C, scale, shift,
                               mean, invstd, c_g, c_mg, c_mgx, relu, dx));
    else
        B200SP_CUDA(launch_pdl(k_bn_bwd_apply<1>, dim3(stream_grid(M * C, 256)), dim3(256), 0, st, x, dy, M, C, scale, shift, mean,
                               invstd, c_g, c_mg, c_mgx, relu, dx));
    B200SP_LAUNCH_CHECK_N(2);
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
