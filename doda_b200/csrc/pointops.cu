// pointops.cu — pointops2_cuda replacements (lib/pointops2/src/pointops_api.cpp:13-23).
// `offset` arrays are END offsets without the leading zero (lib/pointops2/functions/pointops2.py:47,66).
// All kernels take the caller's stream (the reference launches everything on the legacy default stream).
#include "common.cuh"

namespace b200sp {

__device__ __forceinline__ int batch_of(int i, const int* __restrict__ end_off) {
    int b = 0;
    while (i >= end_off[b]) ++b;  // knnquery_cuda_kernel.cu:51-62
    return b;
}

// max-heap of the nsample best candidates kept in local memory, heap-sorted ascending at the end
// (knnquery_cuda_kernel.cu:17-49, 65-108).  Missing neighbours keep idx=start, dist2=1e10.
__device__ __forceinline__ void reheap(float* d, int* ix, int k) {
    int root = 0, child = 1;
    while (child < k) {
        if (child + 1 < k && d[child + 1] > d[child]) ++child;
        if (d[root] > d[child]) return;
        float td = d[root]; d[root] = d[child]; d[child] = td;
        int ti = ix[root]; ix[root] = ix[child]; ix[child] = ti;
        root = child;
        child = root * 2 + 1;
    }
}

__global__ void k_knnquery(int m, int ns, const float* __restrict__ xyz, const float* __restrict__ nxyz,
                           const int* __restrict__ off, const int* __restrict__ noff, int* __restrict__ idx,
                           float* __restrict__ dist2) {
    int pt = blockIdx.x * blockDim.x + threadIdx.x;
    if (pt >= m) return;
    int b = batch_of(pt, noff);
    int s = b == 0 ? 0 : off[b - 1], e = off[b];
    float qx = nxyz[pt * 3], qy = nxyz[pt * 3 + 1], qz = nxyz[pt * 3 + 2];
    float bd[100];
    int bi[100];
    for (int i = 0; i < ns; ++i) {
        bd[i] = 1e10f;
        bi[i] = s;
    }
    for (int i = s; i < e; ++i) {
        float x = __ldg(xyz + i * 3), y = __ldg(xyz + i * 3 + 1), z = __ldg(xyz + i * 3 + 2);
        float d = (qx - x) * (qx - x) + (qy - y) * (qy - y) + (qz - z) * (qz - z);
        if (d < bd[0]) {
            bd[0] = d;
            bi[0] = i;
            reheap(bd, bi, ns);
        }
    }
    for (int i = ns - 1; i > 0; --i) {
        float td = bd[0]; bd[0] = bd[i]; bd[i] = td;
        int ti = bi[0]; bi[0] = bi[i]; bi[i] = ti;
        reheap(bd, bi, i);
    }
    for (int i = 0; i < ns; ++i) {
        idx[pt * ns + i] = bi[i];
        dist2[pt * ns + i] = bd[i];
    }
}

// nsample == 1 fast path (the only call DODA makes: model/unet.py:135-138): plain running minimum
__global__ void k_knnquery1(int m, const float* __restrict__ xyz, const float* __restrict__ nxyz,
                            const int* __restrict__ off, const int* __restrict__ noff, int* __restrict__ idx,
                            float* __restrict__ dist2) {
    int pt = blockIdx.x * blockDim.x + threadIdx.x;
    if (pt >= m) return;
    int b = batch_of(pt, noff);
    int s = b == 0 ? 0 : off[b - 1], e = off[b];
    float qx = nxyz[pt * 3], qy = nxyz[pt * 3 + 1], qz = nxyz[pt * 3 + 2];
    float bd = 1e10f;
    int bi = s;
    for (int i = s; i < e; ++i) {
        float x = __ldg(xyz + i * 3), y = __ldg(xyz + i * 3 + 1), z = __ldg(xyz + i * 3 + 2);
        float d = (qx - x) * (qx - x) + (qy - y) * (qy - y) + (qz - z) * (qz - z);
        if (d < bd) {
            bd = d;
            bi = i;
        }
    }
    idx[pt] = bi;
    dist2[pt] = bd;
}

// furthest point sampling: one block per batch element (sampling_dim_cuda_kernel.cu:15-130); the block size and
// the (value, lower slot wins on ties) tree reduction follow the reference so that ties resolve identically.
__global__ void k_fps(int dim, const float* __restrict__ xyz, const int* __restrict__ off,
                      const int* __restrict__ noff, float* __restrict__ tmp, int* __restrict__ idx) {
    extern __shared__ unsigned char s_raw[];
    float* dists = reinterpret_cast<float*>(s_raw);
    int* dists_i = reinterpret_cast<int*>(dists + blockDim.x);
    int bid = blockIdx.x, tid = threadIdx.x;
    int start_n = bid == 0 ? 0 : off[bid - 1], end_n = off[bid];
    int start_m = bid == 0 ? 0 : noff[bid - 1], end_m = noff[bid];
    int old = start_n;
    if (tid == 0 && start_m < end_m) idx[start_m] = start_n;
    __syncthreads();
    for (int j = start_m + 1; j < end_m; ++j) {
        int besti = start_n;
        float best = -1.f;
        const float* e = xyz + (int64_t)old * dim;
        for (int k = start_n + tid; k < end_n; k += blockDim.x) {
            float d = 0.f;
            for (int kk = 0; kk < dim; ++kk) {
                float t = xyz[(int64_t)k * dim + kk] - e[kk];
                d += t * t;
            }
            float d2 = fminf(d, tmp[k]);
            tmp[k] = d2;
            besti = d2 > best ? k : besti;
            best = d2 > best ? d2 : best;
        }
        dists[tid] = best;
        dists_i[tid] = besti;
        __syncthreads();
        for (int s = blockDim.x >> 1; s > 0; s >>= 1) {
            if (tid < s) {
                float v1 = dists[tid], v2 = dists[tid + s];
                int i1 = dists_i[tid], i2 = dists_i[tid + s];
                dists[tid] = fmaxf(v1, v2);
                dists_i[tid] = v2 > v1 ? i2 : i1;
            }
            __syncthreads();
        }
        old = dists_i[0];
        if (tid == 0) idx[j] = old;
        __syncthreads();
    }
}

__global__ void k_grouping_fwd(int64_t total, int ns, int c, const float* __restrict__ in,
                               const int* __restrict__ idx, float* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int ci = (int)(i % c);
    int64_t ms = i / c;  // m_idx * ns + s_idx
    out[i] = __ldg(in + (int64_t)idx[ms] * c + ci);
}
__global__ void k_grouping_bwd(int64_t total, int ns, int c, const float* __restrict__ dout,
                               const int* __restrict__ idx, float* __restrict__ din) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int ci = (int)(i % c);
    int64_t ms = i / c;
    atomicAdd(din + (int64_t)idx[ms] * c + ci, dout[i]);
}
__global__ void k_interp_fwd(int n, int c, int k, const float* __restrict__ in, const int* __restrict__ idx,
                             const float* __restrict__ w, float* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)n * c) return;
    int ci = (int)(i % c);
    int64_t ni = i / c;
    float acc = out[i];  // the reference accumulates into the caller-zeroed output
    for (int j = 0; j < k; ++j) acc += __ldg(in + (int64_t)idx[ni * k + j] * c + ci) * w[ni * k + j];
    out[i] = acc;
}
__global__ void k_interp_bwd(int n, int c, int k, const float* __restrict__ dout, const int* __restrict__ idx,
                             const float* __restrict__ w, float* __restrict__ din) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)n * c) return;
    int ci = (int)(i % c);
    int64_t ni = i / c;
    float g = dout[i];
    for (int j = 0; j < k; ++j) atomicAdd(din + (int64_t)idx[ni * k + j] * c + ci, g * w[ni * k + j]);
}
__global__ void k_sub_fwd(int64_t total, int ns, int c, const float* __restrict__ in1, const float* __restrict__ in2,
                          const int* __restrict__ idx, float* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int ci = (int)(i % c);
    int64_t nsx = i / c;
    int64_t ni = nsx / ns;
    out[i] = in1[ni * c + ci] - __ldg(in2 + (int64_t)idx[nsx] * c + ci);
}
__global__ void k_sub_bwd(int64_t total, int ns, int c, const int* __restrict__ idx, const float* __restrict__ dout,
                          float* __restrict__ d1, float* __restrict__ d2) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int ci = (int)(i % c);
    int64_t nsx = i / c;
    int64_t ni = nsx / ns;
    float g = dout[i];
    atomicAdd(d1 + ni * c + ci, g);
    atomicAdd(d2 + (int64_t)idx[nsx] * c + ci, -g);
}
__global__ void k_agg_fwd(int n, int ns, int c, int wc, const float* __restrict__ in, const float* __restrict__ pos,
                          const float* __restrict__ w, const int* __restrict__ idx, float* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)n * c) return;
    int ci = (int)(i % c);
    int64_t ni = i / c;
    int wi = ci % wc;
    float acc = out[i];
    for (int s = 0; s < ns; ++s) {
        int64_t ii = ni * ns + s;
        acc += (__ldg(in + (int64_t)idx[ii] * c + ci) + pos[ii * c + ci]) * w[ii * wc + wi];
    }
    out[i] = acc;
}
__global__ void k_agg_bwd(int n, int ns, int c, int wc, const float* __restrict__ in, const float* __restrict__ pos,
                          const float* __restrict__ w, const int* __restrict__ idx, const float* __restrict__ dout,
                          float* __restrict__ din, float* __restrict__ dpos, float* __restrict__ dw) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)n * c) return;
    int ci = (int)(i % c);
    int64_t ni = i / c;
    int wi = ci % wc;
    float g = dout[i];
    for (int s = 0; s < ns; ++s) {
        int64_t ii = ni * ns + s;
        int64_t src = (int64_t)idx[ii] * c + ci;
        float wv = w[ii * wc + wi];
        atomicAdd(din + src, g * wv);
        dpos[ii * c + ci] = g * wv;
        atomicAdd(dw + ii * wc + wi, g * (in[src] + pos[ii * c + ci]));
    }
}

}  // namespace b200sp

using namespace b200sp;

#define PO_GRID(total) (unsigned)cdiv((int64_t)(total), 256), 256, 0, (cudaStream_t)stream

extern "C" int b200sp_knnquery(int m, int nsample, const float* xyz, const float* new_xyz, const int32_t* offset,
                               const int32_t* new_offset, int32_t* idx, float* dist2, void* stream) {
    B200SP_CHECK_ARG(m >= 0 && nsample >= 1 && nsample <= 100, "knnquery: nsample must be in 1..100");
    if (m == 0) return B200SP_OK;
    if (nsample == 1)
        k_knnquery1<<<(unsigned)cdiv(m, 128), 128, 0, (cudaStream_t)stream>>>(m, xyz, new_xyz, offset, new_offset, idx, dist2);
    else
        k_knnquery<<<(unsigned)cdiv(m, 128), 128, 0, (cudaStream_t)stream>>>(m, nsample, xyz, new_xyz, offset, new_offset,
                                                                           idx, dist2);
    B200SP_LAUNCH_CHECK();
    return B200SP_OK;
}

extern "C" int b200sp_furthestsampling(int b, int n_max, int dim, const float* xyz, const int32_t* offset,
                                       const int32_t* new_offset, float* tmp, int32_t* idx, void* stream) {
    B200SP_CHECK_ARG(b >= 0 && dim >= 1, "furthestsampling: bad sizes");
    if (b == 0) return B200SP_OK;
    int threads = 1;
    while (threads * 2 <= n_max && threads < 1024) threads *= 2;  // opt_n_threads (lib/pointops2/src/cuda_utils.h)
    k_fps<<<b, threads, threads * 8, (cudaStream_t)stream>>>(dim, xyz, offset, new_offset, tmp, idx);
    B200SP_LAUNCH_CHECK();
    return B200SP_OK;
}

extern "C" int b200sp_grouping_fwd(int m, int nsample, int c, const float* in, const int32_t* idx, float* out,
                                   void* stream) {
    int64_t total = (int64_t)m * nsample * c;
    if (total == 0) return B200SP_OK;
    k_grouping_fwd<<<PO_GRID(total)>>>(total, nsample, c, in, idx, out);
    B200SP_LAUNCH_CHECK();
    return B200SP_OK;
}
extern "C" int b200sp_grouping_bwd(int m, int nsample, int c, const float* dout, const int32_t* idx, float* din,
                                   void* stream) {
    int64_t total = (int64_t)m * nsample * c;
    if (total == 0) return B200SP_OK;
    k_grouping_bwd<<<PO_GRID(total)>>>(total, nsample, c, dout, idx, din);
    B200SP_LAUNCH_CHECK();
    return B200SP_OK;
}
extern "C" int b200sp_interpolation_fwd(int n, int c, int k, const float* in, const int32_t* idx, const float* weight,
                                        float* out, void* stream) {
    if ((int64_t)n * c == 0) return B200SP_OK;
    k_interp_fwd<<<PO_GRID((int64_t)n * c)>>>(n, c, k, in, idx, weight, out);
    B200SP_LAUNCH_CHECK();
    return B200SP_OK;
}
extern "C" int b200sp_interpolation_bwd(int n, int c, int k, const float* dout, const int32_t* idx,
                                        const float* weight, float* din, void* stream) {
    if ((int64_t)n * c == 0) return B200SP_OK;
    k_interp_bwd<<<PO_GRID((int64_t)n * c)>>>(n, c, k, dout, idx, weight, din);
    B200SP_LAUNCH_CHECK();
    return B200SP_OK;
}
extern "C" int b200sp_subtraction_fwd(int n, int nsample, int c, const float* in1, const float* in2,
                                      const int32_t* idx, float* out, void* stream) {
    int64_t total = (int64_t)n * nsample * c;
    if (total == 0) return B200SP_OK;
    k_sub_fwd<<<PO_GRID(total)>>>(total, nsample, c, in1, in2, idx, out);
    B200SP_LAUNCH_CHECK();
    return B200SP_OK;
}
extern "C" int b200sp_subtraction_bwd(int n, int nsample, int c, const int32_t* idx, const float* dout, float* din1,
                                      float* din2, void* stream) {
    int64_t total = (int64_t)n * nsample * c;
    if (total == 0) return B200SP_OK;
    k_sub_bwd<<<PO_GRID(total)>>>(total, nsample, c, idx, dout, din1, din2);
    B200SP_LAUNCH_CHECK();
    return B200SP_OK;
}
extern "C" int b200sp_aggregation_fwd(int n, int nsample, int c, int w_c, const float* in, const float* pos,
                                      const float* w, const int32_t* idx, float* out, void* stream) {
    B200SP_CHECK_ARG(w_c >= 1, "aggregation_fwd: w_c must be >= 1");
    if ((int64_t)n * c == 0) return B200SP_OK;
    k_agg_fwd<<<PO_GRID((int64_t)n * c)>>>(n, nsample, c, w_c, in, pos, w, idx, out);
    B200SP_LAUNCH_CHECK();
    return B200SP_OK;
}
extern "C" int b200sp_aggregation_bwd(int n, int nsample, int c, int w_c, const float* in, const float* pos,
                                      const float* w, const int32_t* idx, const float* dout, float* din, float* dpos,
                                      float* dw, void* stream) {
    B200SP_CHECK_ARG(w_c >= 1, "aggregation_bwd: w_c must be >= 1");
    if ((int64_t)n * c == 0) return B200SP_OK;
    k_agg_bwd<<<PO_GRID((int64_t)n * c)>>>(n, nsample, c, w_c, in, pos, w, idx, dout, din, dpos, dw);
    B200SP_LAUNCH_CHECK();
    return B200SP_OK;
}
