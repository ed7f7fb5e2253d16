// loss.cu — softmax cross-entropy over the point logits (model/unet.py:168-170: nn.CrossEntropyLoss(ignore_index,
// optional class weights) on scores [N, ncls]).  torch's nll_loss mean reduction walks the N rows in ONE block
// (0.46 ms + 0.27 ms per step at N = 450 k); here one pass over the logits per direction, HBM-bound:
//   fwd: per row lse = log sum exp; loss = sum_i w[y_i] (lse_i - x_i[y_i]) / sum_i w[y_i] over rows with y_i != ignore
//   bwd: dx_i[c] = (softmax_i[c] - [c == y_i]) * w[y_i] * dloss / sum w   (0 for ignored rows), lse recomputed
#include "common.cuh"

namespace b200sp {
namespace {

constexpr int CE_THREADS = 256;
constexpr int CE_MAXGRID = 296;
constexpr int CE_MAXC = 64;

template <int MAXC>
__device__ __forceinline__ void row_lse(const float* __restrict__ x, int C, float (&v)[MAXC], float& m, float& s) {
    m = -INFINITY;
#pragma unroll
    for (int c = 0; c < MAXC; ++c)
        if (c < C) {
            v[c] = __ldg(x + c);
            m = fmaxf(m, v[c]);
        }
    s = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c)
        if (c < C) s += __expf(v[c] - m);
}

// ws: [0] ticket (zero before launch, reset by the last block) | +64 B: per-block partial (loss sum, weight sum) doubles
template <int MAXC>
__global__ void __launch_bounds__(CE_THREADS) k_ce_fwd(const float* __restrict__ logits, const long long* __restrict__ labels,
                                                       const float* __restrict__ weight, long long N, int C,
                                                       long long ignore_index, float* __restrict__ out2,
                                                       unsigned* ticket, double* partial, int* bad) {
    double lsum = 0.0, wsum = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
        const long long y = labels[i];
        if (y == ignore_index) continue;
        if (y < 0 || y >= C) {
            *bad = 1;
            continue;
        }
        float v[MAXC], m, s;
        row_lse<MAXC>(logits + i * C, C, v, m, s);
        const float w = weight ? __ldg(weight + y) : 1.f;
        const float xy = __ldg(logits + i * C + y);
        lsum += (double)(w * (m + logf(s) - xy));
        wsum += (double)w;
    }
    __shared__ double sl[CE_THREADS / 32], sw[CE_THREADS / 32];
    __shared__ int s_last;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lsum += __shfl_xor_sync(0xFFFFFFFFu, lsum, o);
        wsum += __shfl_xor_sync(0xFFFFFFFFu, wsum, o);
    }
    if ((threadIdx.x & 31) == 0) {
        sl[threadIdx.x >> 5] = lsum;
        sw[threadIdx.x >> 5] = wsum;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int i = 0; i < CE_THREADS / 32; ++i) {
            a += sl[i];
            b += sw[i];
        }
        partial[2 * blockIdx.x] = a;
        partial[2 * blockIdx.x + 1] = b;
        __threadfence();
        s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1 : 0;
    }
    __syncthreads();
    if (s_last && threadIdx.x == 0) {  // fixed fold order: the loss is reproducible run to run
        __threadfence();
        double a = 0.0, b = 0.0;
        for (unsigned i = 0; i < gridDim.x; ++i) {
            a += ((volatile double*)partial)[2 * i];
            b += ((volatile double*)partial)[2 * i + 1];
        }
        out2[0] = (float)(a / b);  // 0 / 0 = nan when every row is ignored, like torch
        out2[1] = (float)b;
        *ticket = 0u;
    }
}

template <int MAXC>
__global__ void __launch_bounds__(CE_THREADS) k_ce_bwd(const float* __restrict__ logits, const long long* __restrict__ labels,
                                                       const float* __restrict__ weight, long long N, int C,
                                                       long long ignore_index, const float* __restrict__ out2,
                                                       const float* __restrict__ dloss, float* __restrict__ dlogits) {
    const float scale = dloss[0] / out2[1];
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
        const long long y = labels[i];
        float* d = dlogits + i * C;
        if (y == ignore_index || y < 0 || y >= C) {
#pragma unroll
            for (int c = 0; c < MAXC; ++c)
                if (c < C) d[c] = 0.f;
            continue;
        }
        float v[MAXC], m, s;
        row_lse<MAXC>(logits + i * C, C, v, m, s);
        const float k = (weight ? __ldg(weight + y) : 1.f) * scale;
        const float inv = 1.f / s;
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
            if (c < C) d[c] = k * (__expf(v[c] - m) * inv - (c == (int)y ? 1.f : 0.f));
    }
}

int ce_grid(long long N) {
    long long g = cdiv(N, CE_THREADS * 4);
    if (g < 1) g = 1;
    if (g > CE_MAXGRID) g = CE_MAXGRID;
    return (int)g;
}

}  // namespace
}  // namespace b200sp

using namespace b200sp;

extern "C" int64_t b200sp_cross_entropy_ws_bytes(void) { return 64 + (int64_t)CE_MAXGRID * 16 + 64; }

extern "C" int b200sp_cross_entropy_fwd(const float* logits, const int64_t* labels, const float* weight, int64_t N, int C,
                                        int64_t ignore_index, float* out2, void* ws, int64_t ws_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    B200SP_CHECK_ARG(N >= 1 && C >= 1 && C <= CE_MAXC, "cross_entropy_fwd: need N >= 1 and 1 <= C <= %d", CE_MAXC);
    B200SP_CHECK_ARG(logits && labels && out2 && ws && ws_bytes >= b200sp_cross_entropy_ws_bytes(),
                     "cross_entropy_fwd: null pointer or workspace too small");
    unsigned* ticket = (unsigned*)ws;
    int* bad = (int*)ws + 1;  // sticky: a label outside [0, C) that is not ignore_index
    double* partial = (double*)((char*)ws + 64);
    const int G = ce_grid(N);
    if (C <= 16)
        k_ce_fwd<16><<<G, CE_THREADS, 0, st>>>(logits, (const long long*)labels, weight, N, C, ignore_index, out2, ticket, partial, bad);
    else
        k_ce_fwd<CE_MAXC><<<G, CE_THREADS, 0, st>>>(logits, (const long long*)labels, weight, N, C, ignore_index, out2, ticket, partial, bad);
    B200SP_LAUNCH_CHECK();
    return B200SP_OK;
}

extern "C" int b200sp_cross_entropy_bwd(const float* logits, const int64_t* labels, const float* weight, int64_t N, int C,
                                        int64_t ignore_index, const float* out2, const float* dloss, float* dlogits,
                                        void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    B200SP_CHECK_ARG(N >= 1 && C >= 1 && C <= CE_MAXC, "cross_entropy_bwd: need N >= 1 and 1 <= C <= %d", CE_MAXC);
    B200SP_CHECK_ARG(logits && labels && out2 && dloss && dlogits, "cross_entropy_bwd: null pointer");
    const int G = ce_grid(N);
    if (C <= 16)
        k_ce_bwd<16><<<G, CE_THREADS, 0, st>>>(logits, (const long long*)labels, weight, N, C, ignore_index, out2, dloss, dlogits);
    else
        k_ce_bwd<CE_MAXC><<<G, CE_THREADS, 0, st>>>(logits, (const long long*)labels, weight, N, C, ignore_index, out2, dloss, dlogits);
    B200SP_LAUNCH_CHECK();
    return B200SP_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Metric epilogue: per-class intersection / union / target counts of util/common_utils.py:233-247
// (intersectionAndUnionGPU: three torch.histc calls on CPU copies = three device->host syncs per iteration,
// tool/train.py:113-118).  One pass over (prediction, label), block-private histograms in shared memory, integer
// atomics to a [3][K] device buffer, counts returned as float like histc does.  Rows whose label is ignore_index are
// skipped entirely; predictions / labels outside [0, K) fall outside the histogram range, as with histc.
// ---------------------------------------------------------------------------------------------------------------
namespace b200sp {
namespace {
constexpr int IOU_MAXK = 1024;

__global__ void __launch_bounds__(256) k_iou_hist(const long long* __restrict__ pred, const long long* __restrict__ label,
                                                  long long N, int K, long long ignore_index, int* __restrict__ acc) {
    extern __shared__ int s_h[];  // [3][K]: intersection, prediction, target
    for (int i = threadIdx.x; i < 3 * K; i += blockDim.x) s_h[i] = 0;
    __syncthreads();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
        const long long t = label[i];
        if (t == ignore_index) continue;
        const long long o = pred[i];
        if (o >= 0 && o < K) atomicAdd(&s_h[K + (int)o], 1);
        if (t >= 0 && t < K) {
            atomicAdd(&s_h[2 * K + (int)t], 1);
            if (o == t) atomicAdd(&s_h[(int)t], 1);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * K; i += blockDim.x)
        if (s_h[i]) atomicAdd(&acc[i], s_h[i]);
}

// out[0] = intersection, out[1] = union = prediction + target - intersection, out[2] = target
__global__ void k_iou_finish(const int* __restrict__ acc, int K, float* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= K) return;
    const int inter = acc[c], po = acc[K + c], ta = acc[2 * K + c];
    out[c] = (float)inter;
    out[K + c] = (float)(po + ta - inter);
    out[2 * K + c] = (float)ta;
}
}  // namespace
}  // namespace b200sp

extern "C" int b200sp_intersection_union(const int64_t* pred, const int64_t* label, int64_t N, int K, int64_t ignore_index,
                                         float* out3k, void* ws, int64_t ws_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    B200SP_CHECK_ARG(N >= 0 && K >= 1 && K <= IOU_MAXK, "intersection_union: need N >= 0 and 1 <= K <= %d", IOU_MAXK);
    B200SP_CHECK_ARG(out3k && ws && ws_bytes >= (int64_t)3 * K * 4, "intersection_union: null output or workspace < 12 K bytes");
    B200SP_CHECK_ARG(N == 0 || (pred && label), "intersection_union: null input");
    int* acc = (int*)ws;
    B200SP_CUDA(cudaMemsetAsync(acc, 0, (size_t)3 * K * 4, st));
    if (N > 0) {
        long long g = cdiv(N, 256 * 8);
        if (g > 296) g = 296;
        k_iou_hist<<<(unsigned)g, 256, (size_t)3 * K * sizeof(int), st>>>((const long long*)pred, (const long long*)label, N, K,
                                                                        ignore_index, acc);
        B200SP_LAUNCH_CHECK();
    }
    k_iou_finish<<<(unsigned)cdiv(K, 128), 128, 0, st>>>(acc, K, out3k);
    B200SP_LAUNCH_CHECK();
    return B200SP_OK;
}
