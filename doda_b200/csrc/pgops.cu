// pgops.cu — the remaining PG_OP kernels (lib/pointgroup_ops/src/pointgroup_ops_api.cpp:7-26):
// CSR-segment reductions, ROI max-pool, IoU, ball query and batched kNN.
// Layout choices for B200: a warp walks a segment with lanes across channels (row reads are coalesced),
// ball query is count -> cub scan -> fill (deterministic CSR, no device malloc, no blocking memcpy pair),
// everything takes the caller's stream (the reference launches on the legacy default stream).
#include "common.cuh"
#include <cub/device/device_scan.cuh>
#include <float.h>

namespace b200sp {

enum SecOp { SEC_MEAN = 0, SEC_MIN = 1, SEC_MAX = 2, SEC_ARGMAX = 3 };

// block = (32 channels) x (8 segments); grid.y tiles the channels
template <int OP>
__global__ void k_sec_reduce(const float* __restrict__ inp, const int* __restrict__ offsets, float* __restrict__ out,
                             int* __restrict__ argidx, int P, int C) {
    int c = blockIdx.y * 32 + threadIdx.x;
    for (int p = blockIdx.x * blockDim.y + threadIdx.y; p < P; p += gridDim.x * blockDim.y) {
        if (c >= C) continue;
        int s = offsets[p], e = offsets[p + 1];
        if (OP == SEC_MEAN) {
            float cnt = (float)(e - s), m = 0.f;
            for (int i = s; i < e; ++i) m += __ldg(inp + (int64_t)i * C + c) / cnt;  // sec_mean.cu:22
            out[(int64_t)p * C + c] = m;
        } else if (OP == SEC_MIN) {
            float v = INFINITY;  // 1e50 as float literal -> +inf (sec_mean.cu:68)
            for (int i = s; i < e; ++i) v = fminf(v, __ldg(inp + (int64_t)i * C + c));
            out[(int64_t)p * C + c] = v;
        } else {
            float v = -INFINITY;
            int a = -1;
            for (int i = s; i < e; ++i) {
                float t = __ldg(inp + (int64_t)i * C + c);
                if (t > v) {
                    v = t;
                    a = i;
                }
            }
            out[(int64_t)p * C + c] = v;
            if (OP == SEC_ARGMAX) argidx[(int64_t)p * C + c] = a;
        }
    }
}

__global__ void k_sec_mean_bp(const float* __restrict__ dout, const int* __restrict__ offsets,
                              float* __restrict__ dinp, int P, int C) {
    int c = blockIdx.y * 32 + threadIdx.x;
    for (int p = blockIdx.x * blockDim.y + threadIdx.y; p < P; p += gridDim.x * blockDim.y) {
        if (c >= C) continue;
        int s = offsets[p], e = offsets[p + 1];
        float g = dout[(int64_t)p * C + c] / (float)(e - s);
        for (int i = s; i < e; ++i) dinp[(int64_t)i * C + c] = g;
    }
}

__global__ void k_roipool_bp(const float* __restrict__ dout, const int* __restrict__ maxidx,
                             float* __restrict__ dfeats, int P, int C) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)P * C) return;
    int a = maxidx[i];
    if (a >= 0) atomicAdd(dfeats + (int64_t)a * C + (int)(i % C), dout[i]);
}

__global__ void k_get_iou(const int* __restrict__ pidx, const int* __restrict__ poff,
                          const int64_t* __restrict__ labels, const int* __restrict__ pointnum,
                          float* __restrict__ iou, int P, int I) {
    // one block per proposal: histogram the proposal's instance labels in shared memory once
    extern __shared__ int s_hist[];
    for (int p = blockIdx.x; p < P; p += gridDim.x) {
        for (int i = threadIdx.x; i < I; i += blockDim.x) s_hist[i] = 0;
        __syncthreads();
        int s = poff[p], e = poff[p + 1];
        for (int i = s + threadIdx.x; i < e; i += blockDim.x) {
            int l = (int)labels[pidx[i]];
            if (l >= 0 && l < I) atomicAdd(&s_hist[l], 1);
        }
        __syncthreads();
        int total = e - s;
        for (int i = threadIdx.x; i < I; i += blockDim.x) {
            int inter = s_hist[i];
            iou[(int64_t)p * I + i] = (float)inter / ((float)(total + pointnum[i] - inter) + 1e-5f);
        }
        __syncthreads();
    }
}

// ---- ball query: points of one batch segment are streamed through shared memory tiles ----
constexpr int BQ_TILE = 256;
constexpr int BQ_CAP = 1000;  // the reference keeps at most 1000 neighbours per point (bfs_cluster.cu:20,38-45)

template <bool FILL>
__global__ void __launch_bounds__(BQ_TILE) k_ballquery(const float* __restrict__ xyz, const int* __restrict__ bidx,
                                                       const int* __restrict__ boff, int n, float r2,
                                                       int* __restrict__ counts, const int* __restrict__ starts,
                                                       int* __restrict__ idx, int* __restrict__ start_len,
                                                       int64_t limit) {
    int pt = blockIdx.x * blockDim.x + threadIdx.x;
    bool live = pt < n;
    float ox = 0, oy = 0, oz = 0;
    int s = 0, e = 0;
    if (live) {
        ox = xyz[pt * 3 + 0]; oy = xyz[pt * 3 + 1]; oz = xyz[pt * 3 + 2];
        int b = bidx[pt];
        s = boff[b]; e = boff[b + 1];
    }
    int cnt = 0;
    int64_t base = 0;
    if (FILL && live) base = starts[pt];
    for (int k = s; k < e; ++k) {
        float x = __ldg(xyz + k * 3 + 0), y = __ldg(xyz + k * 3 + 1), z = __ldg(xyz + k * 3 + 2);
        float d2 = (ox - x) * (ox - x) + (oy - y) * (oy - y) + (oz - z) * (oz - z);
        if (d2 < r2) {
            if (cnt >= BQ_CAP) break;
            if (FILL && base + cnt < limit) idx[base + cnt] = k;
            ++cnt;
        }
    }
    if (live) {
        if (!FILL) counts[pt] = cnt;
        else {
            start_len[pt * 2 + 0] = (int)base;
            start_len[pt * 2 + 1] = cnt;
        }
    }
}

// ---- batched kNN (knn.cu:7-50): insertion into a sorted list of k<=40 candidates ----
__global__ void k_knn_batch(int n, int k, const float* __restrict__ xyz, const float* __restrict__ q,
                            const int* __restrict__ bidx, const int* __restrict__ qoff, int* __restrict__ idx) {
    int pt = blockIdx.x * blockDim.x + threadIdx.x;
    if (pt >= n) return;
    float ox = xyz[pt * 3], oy = xyz[pt * 3 + 1], oz = xyz[pt * 3 + 2];
    float best[40];
    int besti[40];
    for (int i = 0; i < k; ++i) {
        best[i] = 1e20f;
        besti[i] = 0;
    }
    int b = bidx[pt];
    int s = qoff[b], e = qoff[b + 1];
    for (int i = s; i < e; ++i) {
        float x = __ldg(q + i * 3), y = __ldg(q + i * 3 + 1), z = __ldg(q + i * 3 + 2);
        float d2 = (ox - x) * (ox - x) + (oy - y) * (oy - y) + (oz - z) * (oz - z);
        if (d2 < best[k - 1]) {
            int p = k - 1;
            while (p > 0 && d2 < best[p - 1]) {
                best[p] = best[p - 1];
                besti[p] = besti[p - 1];
                --p;
            }
            best[p] = d2;
            besti[p] = i;
        }
    }
    for (int i = 0; i < k; ++i) idx[pt * k + i] = besti[i];
}

static dim3 sec_grid(int P, int C) {
    int gx = (int)cdiv(P, 8);
    if (gx > 148 * 32) gx = 148 * 32;
    if (gx < 1) gx = 1;
    return dim3(gx, (unsigned)cdiv(C, 32));
}

}  // namespace b200sp

using namespace b200sp;

#define SEC_ENTRY(NAME, OP)                                                                                      \
    extern "C" int NAME(const float* inp, const int32_t* offsets, float* out, int P, int C, void* stream) {      \
        B200SP_CHECK_ARG(P >= 0 && C >= 1, #NAME ": bad sizes");                                                 \
        if (P == 0) return B200SP_OK;                                                                            \
        k_sec_reduce<OP><<<sec_grid(P, C), dim3(32, 8), 0, (cudaStream_t)stream>>>(inp, offsets, out, nullptr, P, C); \
        B200SP_LAUNCH_CHECK();                                                                                   \
        return B200SP_OK;                                                                                        \
    }
SEC_ENTRY(b200sp_sec_mean, SEC_MEAN)
SEC_ENTRY(b200sp_sec_min, SEC_MIN)
SEC_ENTRY(b200sp_sec_max, SEC_MAX)

extern "C" int b200sp_sec_mean_bp(const float* dout, const int32_t* offsets, float* dinp, int P, int C, void* stream) {
    B200SP_CHECK_ARG(P >= 0 && C >= 1, "sec_mean_bp: bad sizes");
    if (P == 0) return B200SP_OK;
    k_sec_mean_bp<<<sec_grid(P, C), dim3(32, 8), 0, (cudaStream_t)stream>>>(dout, offsets, dinp, P, C);
    B200SP_LAUNCH_CHECK();
    return B200SP_OK;
}

extern "C" int b200sp_roipool_fp(const float* feats, const int32_t* offsets, float* out, int32_t* maxidx, int P, int C,
                                 void* stream) {
    B200SP_CHECK_ARG(P >= 0 && C >= 1 && maxidx, "roipool_fp: bad arguments");
    if (P == 0) return B200SP_OK;
    k_sec_reduce<SEC_ARGMAX><<<sec_grid(P, C), dim3(32, 8), 0, (cudaStream_t)stream>>>(feats, offsets, out, maxidx, P, C);
    B200SP_LAUNCH_CHECK();
    return B200SP_OK;
}

extern "C" int b200sp_roipool_bp(const float* dout, const int32_t* maxidx, float* dfeats, int P, int C, void* stream) {
    B200SP_CHECK_ARG(P >= 0 && C >= 1, "roipool_bp: bad sizes");
    if (P == 0) return B200SP_OK;
    k_roipool_bp<<<(unsigned)cdiv((int64_t)P * C, 256), 256, 0, (cudaStream_t)stream>>>(dout, maxidx, dfeats, P, C);
    B200SP_LAUNCH_CHECK();
    return B200SP_OK;
}

extern "C" int b200sp_get_iou(const int32_t* pidx, const int32_t* poff, const int64_t* labels, const int32_t* pointnum,
                              float* iou, int P, int I, void* stream) {
    B200SP_CHECK_ARG(P >= 0 && I >= 0, "get_iou: bad sizes");
    B200SP_CHECK_ARG((size_t)I * 4 <= 48 * 1024, "get_iou: more than 12288 instances not supported");
    if (P == 0 || I == 0) return B200SP_OK;
    k_get_iou<<<P < 148 * 8 ? P : 148 * 8, 256, sizeof(int) * I, (cudaStream_t)stream>>>(pidx, poff, labels, pointnum,
                                                                                     iou, P, I);
    B200SP_LAUNCH_CHECK();
    return B200SP_OK;
}

extern "C" int64_t b200sp_ballquery_ws_bytes(int n) {
    size_t cub_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, (int*)nullptr, (int*)nullptr, n > 0 ? n : 1);
    return align_up((int64_t)cub_bytes, 256) + 2 * align_up((int64_t)(n + 1) * 4, 256) + 512;
}

extern "C" int b200sp_ballquery_batch_p(const float* xyz, const int32_t* bidx, const int32_t* boff, int32_t* idx,
                                        int32_t* start_len, int n, int mean_active, float radius, void* ws,
                                        int64_t ws_bytes, int32_t* n_active_host, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    B200SP_CHECK_ARG(n >= 0 && mean_active >= 0 && n_active_host, "ballquery: bad arguments");
    *n_active_host = 0;
    if (n == 0) return B200SP_OK;
    B200SP_CHECK_ARG(ws && ws_bytes >= b200sp_ballquery_ws_bytes(n), "ballquery: workspace too small");
    char* p = (char*)ws;
    int* counts = (int*)p;
    p += align_up((int64_t)(n + 1) * 4, 256);
    int* starts = (int*)p;
    p += align_up((int64_t)(n + 1) * 4, 256);
    size_t cub_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, counts, starts, n + 1);
    unsigned grid = (unsigned)cdiv(n, BQ_TILE);
    float r2 = radius * radius;
    B200SP_CUDA(cudaMemsetAsync(counts + n, 0, 4, st));
    k_ballquery<false><<<grid, BQ_TILE, 0, st>>>(xyz, bidx, boff, n, r2, counts, nullptr, nullptr, nullptr, 0);
    B200SP_CUDA(cub::DeviceScan::ExclusiveSum(p, cub_bytes, counts, starts, n + 1, st));
    int64_t limit = (int64_t)n * mean_active;
    k_ballquery<true><<<grid, BQ_TILE, 0, st>>>(xyz, bidx, boff, n, r2, nullptr, starts, idx, start_len, limit);
    B200SP_LAUNCH_CHECK_N(2 + 2 /* cub scan */);
    int total = 0;
    B200SP_CUDA(cudaMemcpyAsync(&total, starts + n, 4, cudaMemcpyDeviceToHost, st));
    B200SP_CUDA(cudaStreamSynchronize(st));
    *n_active_host = total;
    return B200SP_OK;
}

extern "C" int b200sp_knn_batch(const float* xyz, const float* query, const int32_t* bidx, const int32_t* qoff,
                                int32_t* idx, int n, int m, int k, void* stream) {
    (void)m;
    B200SP_CHECK_ARG(n >= 0 && k >= 1 && k <= 40, "knn_batch: k must be in 1..40 (knn.cu:18)");
    if (n == 0) return B200SP_OK;
    k_knn_batch<<<(unsigned)cdiv(n, 128), 128, 0, (cudaStream_t)stream>>>(n, k, xyz, query, bidx, qoff, idx);
    B200SP_LAUNCH_CHECK();
    return B200SP_OK;
}
