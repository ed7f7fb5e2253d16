// layer.cu — one C-ABI call per sparse-conv LAYER and direction (the host-side executor of the hot path).
//
// A training step of DODA's U-Net is ~850 kernel launches for ~13 ms of GPU work; driven layer by layer from Python
// (autograd node -> 5 wrapper functions -> 3-5 C calls, each with its own allocations and argument marshalling) the
// host needs about as long to ENQUEUE a step as the GPU needs to run it.  These two entry points run everything one
// [BatchNorm -> ReLU ->] sparse conv layer needs in one call:
//
//   b200sp_conv_layer_fwd : batch statistics + normalise + ReLU (k_bn_reduce, k_affine_relu), then the conv with the
//                           kernel the shape calls for (k_conv_direct / k_conv_tc table or pair mode / fp32 fallback)
//   b200sp_conv_layer_bwd : fork the side stream, weight gradient there (k_wgrad_direct / k_wgrad_tc / k_wgrad),
//                           dgrad on the main stream (same dispatch as the forward, transposed / mirrored weights),
//                           [+ the extra gradient of the exposed activation], BN backward (k_bn_reduce<1>,
//                           k_bn_bwd_apply), join
//
// i.e. the dispatch that doda_b200/ops.py used to do per call in Python (conv_forward_raw / _conv_dgrad / _conv_wgrad)
// -- spconv v1.2's `indice_conv` / `indice_conv_backward` behind SubMConv3d / SparseConv3d /
// SparseInverseConv3d (model/unet.py:36, model/unet_block.py:20,26,29,48,70,78) and the nn.BatchNorm1d + nn.ReLU in
// front of them (model/unet_block.py:24-28).  Arguments arrive as one int64 vector (pointers, sizes; doubles as their
// bit pattern): no torch types, no allocation inside, every buffer comes from the caller.
#include "common.cuh"
#include <string.h>
#include <algorithm>

using namespace b200sp;

namespace {

inline double as_f64(int64_t v) {
    double d;
    memcpy(&d, &v, sizeof(d));
    return d;
}
template <typename T>
inline T* as_ptr(int64_t v) {
    return reinterpret_cast<T*>(static_cast<uintptr_t>(v));
}

// host-side rulebook descriptor: B200SP_RB_WORDS x int64 filled once per rulebook by the host mirror
struct RB {
    const int32_t *nbr, *nbr_perm, *order, *rowmask, *fwd, *bwd, *pairs_in, *pairs_out, *pairnum;
    int64_t n_fine, n_coarse, pstride;
    int K, nonoverlap;
};
RB read_rb(const int64_t* d) {
    RB r{};
    if (!d) return r;
    r.nbr = as_ptr<const int32_t>(d[B200SP_RB_NBR]);
    r.nbr_perm = as_ptr<const int32_t>(d[B200SP_RB_NBR_PERM]);
    r.order = as_ptr<const int32_t>(d[B200SP_RB_ORDER]);
    r.rowmask = as_ptr<const int32_t>(d[B200SP_RB_ROWMASK]);
    r.fwd = as_ptr<const int32_t>(d[B200SP_RB_FWD]);
    r.bwd = as_ptr<const int32_t>(d[B200SP_RB_BWD]);
    r.pairs_in = as_ptr<const int32_t>(d[B200SP_RB_PAIRS_IN]);
    r.pairs_out = as_ptr<const int32_t>(d[B200SP_RB_PAIRS_OUT]);
    r.pairnum = as_ptr<const int32_t>(d[B200SP_RB_PAIRNUM]);
    r.n_fine = d[B200SP_RB_N_FINE];
    r.n_coarse = d[B200SP_RB_N_COARSE];
    r.pstride = d[B200SP_RB_PSTRIDE];
    r.K = (int)d[B200SP_RB_K];
    r.nonoverlap = (int)d[B200SP_RB_NONOVERLAP];
    return r;
}

enum { KIND_SUBM = 0, KIND_DENSE = 1, KIND_CONV = 2, KIND_INVERSE = 3 };
constexpr int W_T = 1, W_T_MIRROR = 3, W_PREP = 4;

struct Weights {
    const float* raw;   // module weight [K][Ci_w][Co_w]
    const float* img;   // prepared tensor-core image for the requested flags, or NULL
    void* ws;           // scratch for a per-call weight pre-pass when img == NULL
    int64_t ws_bytes;
};

// res != nullptr: out = conv + res (the residual / skip branch that meets this layer's output, or, in a dgrad, the
// gradient of a second consumer of this layer's input)
int gg(const float* in, int64_t n_in, int Cin, const Weights& w, int wflags, const int32_t* tab, const int32_t* orow,
       const int32_t* rowmask, int K, float* out, int64_t n_out, int Cout, void* st, const float* res = nullptr) {
    if (res) {
        if (w.img) return b200sp_gather_gemm_res(in, n_in, Cin, w.img, wflags | W_PREP, tab, orow, rowmask, K, out, n_out, Cout, res, nullptr, 0, st);
        return b200sp_gather_gemm_res(in, n_in, Cin, w.raw, wflags, tab, orow, rowmask, K, out, n_out, Cout, res, w.ws, w.ws_bytes, st);
    }
    if (w.img) return b200sp_gather_gemm(in, n_in, Cin, w.img, wflags | W_PREP, tab, orow, rowmask, K, out, n_out, Cout, 0, nullptr, 0, st);
    return b200sp_gather_gemm(in, n_in, Cin, w.raw, wflags, tab, orow, rowmask, K, out, n_out, Cout, 0, w.ws, w.ws_bytes, st);
}
int gg_pairs(const float* in, int Cin, const Weights& w, int wflags, const RB& rb, const int32_t* pin, const int32_t* pout,
             int64_t n_upper, float* out, int64_t n_out, int Cout, void* st) {
    // rows no pair reaches stay zero
    B200SP_CUDA(cudaMemsetAsync(out, 0, (size_t)n_out * Cout * sizeof(float), (cudaStream_t)st));
    if (w.img) return b200sp_gather_gemm_pairs(in, Cin, w.img, wflags | W_PREP, pin, pout, rb.pairnum, n_upper, rb.K, rb.pstride, out, Cout, 0, nullptr, 0, st);
    return b200sp_gather_gemm_pairs(in, Cin, w.raw, wflags, pin, pout, rb.pairnum, n_upper, rb.K, rb.pstride, out, Cout, 0, w.ws, w.ws_bytes, st);
}

// forward of one sparse conv: out [n_out, Cout] from feat [M, Cin]
int conv_forward(int kind, const RB& rb, const float* feat, int64_t M, int Cin, const Weights& w, int Kw, float* out,
                 int Cout, void* st, const float* res = nullptr) {
    B200SP_CHECK_ARG(!res || kind == KIND_SUBM || kind == KIND_DENSE, "conv_layer_fwd: a residual needs a conv that keeps the sites");
    switch (kind) {
        case KIND_SUBM:
            if (rb.nbr_perm) return gg(feat, M, Cin, w, 0, rb.nbr_perm, rb.order, rb.rowmask, Kw, out, M, Cout, st, res);
            return gg(feat, M, Cin, w, 0, rb.nbr, nullptr, nullptr, Kw, out, M, Cout, st, res);
        case KIND_DENSE:
            return gg(feat, M, Cin, w, 0, nullptr, nullptr, nullptr, 1, out, M, Cout, st, res);
        case KIND_CONV:
            return gg(feat, M, Cin, w, 0, rb.bwd, nullptr, nullptr, Kw, out, rb.n_coarse, Cout, st);
        default:  // inverse: coarse -> fine through the strided conv's rulebook with the roles swapped
            if (rb.nonoverlap && !b200sp_conv_direct_covers(Kw, Cin, Cout) && rb.pairs_in)
                return gg_pairs(feat, Cin, w, 0, rb, rb.pairs_out, rb.pairs_in, rb.n_fine, out, rb.n_fine, Cout, st);
            return gg(feat, M, Cin, w, 0, rb.fwd, nullptr, nullptr, Kw, out, rb.n_fine, Cout, st);
    }
}

// dgrad: din [M, Cin] from g [n_g, Cout] (the conv maps Cin -> Cout; wflags transpose the weights)
int conv_dgrad(int kind, const RB& rb, const float* g, int64_t n_g, int Cin, const Weights& w, int Kw, float* din, int64_t M,
               int Cout, void* st, const float* res = nullptr) {
    B200SP_CHECK_ARG(!res || kind == KIND_SUBM || kind == KIND_DENSE, "conv_layer_bwd: dx_add without BatchNorm needs a conv that keeps the sites");
    switch (kind) {
        case KIND_SUBM:
            if (rb.nbr_perm) return gg(g, n_g, Cout, w, W_T_MIRROR, rb.nbr_perm, rb.order, rb.rowmask, Kw, din, M, Cin, st, res);
            return gg(g, n_g, Cout, w, W_T_MIRROR, rb.nbr, nullptr, nullptr, Kw, din, M, Cin, st, res);
        case KIND_DENSE:
            return gg(g, n_g, Cout, w, W_T, nullptr, nullptr, nullptr, 1, din, M, Cin, st, res);
        case KIND_CONV:
            if (rb.nonoverlap && !b200sp_conv_direct_covers(Kw, Cout, Cin) && rb.pairs_in)
                return gg_pairs(g, Cout, w, W_T, rb, rb.pairs_out, rb.pairs_in, M, din, M, Cin, st);
            return gg(g, n_g, Cout, w, W_T, rb.fwd, nullptr, nullptr, Kw, din, M, Cin, st);
        default:  // inverse
            return gg(g, n_g, Cout, w, W_T, rb.bwd, nullptr, nullptr, Kw, din, M, Cin, st);
    }
}

// weight gradient dW [Kw][Cin][Cout] (zero-filled by the caller) from a = conv input [M, Cin], g = output gradient
int conv_wgrad(int kind, const RB& rb, const float* a, int64_t M, int Cin, const float* g, int64_t n_g, int Cout, int Kw,
               float* dW, void* st) {
    // table rows of this layer's weight gradient: the output sites (SubM, 1x1: M; strided: coarse; inverse: fine)
    const int64_t n_tab = kind == KIND_CONV ? n_g : (kind == KIND_INVERSE ? rb.n_fine : M);
    const bool have_pairs = kind == KIND_DENSE || rb.pairs_in != nullptr;
    const bool table = have_pairs ? b200sp_wgrad_table_prefers(Kw, Cin, Cout, n_tab) != 0
                                  : b200sp_wgrad_table_covers(Kw, Cin, Cout) != 0;
    switch (kind) {
        case KIND_SUBM:
            if (rb.nbr_perm && table) return b200sp_wgrad_table(a, Cin, g, Cout, rb.nbr_perm, rb.order, rb.rowmask, M, Kw, dW, st);
            B200SP_CHECK_ARG(rb.pairs_in, "conv_layer_bwd: the SubM rulebook was built without pair lists");
            return b200sp_wgrad(a, Cin, g, Cout, rb.pairs_in, rb.pairs_out, rb.pairnum, M, Kw, rb.pstride, dW, st);
        case KIND_DENSE:
            if (table) return b200sp_wgrad_table(a, Cin, g, Cout, nullptr, nullptr, nullptr, M, 1, dW, st);
            return b200sp_wgrad(a, Cin, g, Cout, nullptr, nullptr, nullptr, M, 1, 0, dW, st);
        case KIND_CONV:
            if (table) return b200sp_wgrad_table(a, Cin, g, Cout, rb.bwd, nullptr, nullptr, n_g, Kw, dW, st);
            B200SP_CHECK_ARG(rb.pairs_in, "conv_layer_bwd: the strided rulebook was built without pair lists");
            return b200sp_wgrad(a, Cin, g, Cout, rb.pairs_in, rb.pairs_out, rb.pairnum, M, Kw, rb.pstride, dW, st);
        default:  // inverse: a lives on the coarse sites, g on the fine ones
            if (table) return b200sp_wgrad_table(a, Cin, g, Cout, rb.fwd, nullptr, nullptr, rb.n_fine, Kw, dW, st);
            B200SP_CHECK_ARG(rb.pairs_in, "conv_layer_bwd: the strided rulebook was built without pair lists");
            return b200sp_wgrad(a, Cin, g, Cout, rb.pairs_out, rb.pairs_in, rb.pairnum, rb.n_fine, Kw, rb.pstride, dW, st);
    }
}

__global__ void __launch_bounds__(256) k_add_inplace(float* __restrict__ a, const float* __restrict__ b, int64_t n4) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 x = reinterpret_cast<float4*>(a)[i];
        const float4 y = __ldg(reinterpret_cast<const float4*>(b) + i);
        x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w;
        reinterpret_cast<float4*>(a)[i] = x;
    }
}

}  // namespace

// args (int64 each; see doda_b200/ops.py:_layer_fwd_args):
//  0 kind  1 rb desc (host ptr, 0 for 1x1)  2 x  3 M  4 Cin  5 W  6 W image (fwd) or 0  7 Kw  8 Cout  9 out  10 n_out
//  11 has_bn  12 bn_w  13 bn_b  14 eps (f64 bits)  15 momentum (f64 bits)  16 running_mean  17 running_var
//  18 num_batches_tracked  19 y  20 stats [2][Cin]  21 bn_ws  22 bn_ws_bytes  23 conv_ws  24 conv_ws_bytes  25 stream
//  26 (optional) residual [n_out][Cout]: out = conv + residual (SubM / 1x1 only)
extern "C" int b200sp_conv_layer_fwd(const int64_t* a, int n) {
    B200SP_CHECK_ARG(a && n >= 26, "conv_layer_fwd: expected 26 arguments, got %d", n);
    const int kind = (int)a[0];
    const RB rb = read_rb(as_ptr<const int64_t>(a[1]));
    const float* x = as_ptr<const float>(a[2]);
    const int64_t M = a[3];
    const int Cin = (int)a[4], Kw = (int)a[7], Cout = (int)a[8];
    Weights w{as_ptr<const float>(a[5]), as_ptr<const float>(a[6]), as_ptr<void>(a[23]), a[24]};
    float* out = as_ptr<float>(a[9]);
    void* st = as_ptr<void>(a[25]);
    B200SP_CHECK_ARG(kind >= 0 && kind <= 3 && (kind == KIND_DENSE || a[1]), "conv_layer_fwd: bad kind / missing rulebook");
    if (M == 0) return B200SP_OK;
    const float* feat = x;
    if (a[11]) {
        float* y = as_ptr<float>(a[19]);
        float* stats = as_ptr<float>(a[20]);
        int rc = b200sp_bn_fwd_train(x, M, Cin, as_ptr<const float>(a[12]), as_ptr<const float>(a[13]), (float)as_f64(a[14]), 1,
                                     y, stats, stats + Cin, as_ptr<float>(a[16]), as_ptr<float>(a[17]), (float)as_f64(a[15]),
                                     as_ptr<int64_t>(a[18]), as_ptr<void>(a[21]), a[22], st);
        if (rc) return rc;
        feat = y;
    }
    return conv_forward(kind, rb, feat, M, Cin, w, Kw, out, Cout, st, n >= 27 ? as_ptr<const float>(a[26]) : nullptr);
}

// args (see doda_b200/ops.py:_layer_bwd_args):
//  0 kind  1 rb desc  2 x (BN input; unused without BN)  3 M  4 Cin  5 W  6 W image (dgrad) or 0  7 Kw  8 Cout
//  9 grad_out  10 n_g (rows of grad_out)  11 has_bn  12 bn_w  13 bn_b  14 stats  15 a (conv input: y with BN, x without)
//  16 bn_ws  17 bn_ws_bytes  18 conv_ws  19 conv_ws_bytes  20 need_din  21 need_dw  22 dW (zero-filled)
//  23 dy [M][Cin] (gradient of the conv input)  24 dx [M][Cin] (gradient of the BN input)  25 dwb [2][Cin]
//  26 main stream  27 side stream (0: everything on main)  28 fork event  29 join event
//  30 grad_y (extra gradient of the exposed BN+ReLU activation, or 0)  31 defer_join (caller joins later)
//  32 (optional) dx_add [M][Cin]: added to this layer's input gradient (dx with BN, dy without) -- the gradient of a
//     second consumer of the layer's input (residual skip, U-Net skip connection)
extern "C" int b200sp_conv_layer_bwd(const int64_t* a, int n) {
    B200SP_CHECK_ARG(a && n >= 32, "conv_layer_bwd: expected 32 arguments, got %d", n);
    const int kind = (int)a[0];
    const RB rb = read_rb(as_ptr<const int64_t>(a[1]));
    const int64_t M = a[3], n_g = a[10];
    const int Cin = (int)a[4], Kw = (int)a[7], Cout = (int)a[8];
    Weights w{as_ptr<const float>(a[5]), as_ptr<const float>(a[6]), as_ptr<void>(a[18]), a[19]};
    const float* g = as_ptr<const float>(a[9]);
    const float* act = as_ptr<const float>(a[15]);
    const bool has_bn = a[11] != 0, need_din = a[20] != 0 || has_bn, need_dw = a[21] != 0;
    void* main_st = as_ptr<void>(a[26]);
    void* side_st = as_ptr<void>(a[27]);
    B200SP_CHECK_ARG(kind >= 0 && kind <= 3 && (kind == KIND_DENSE || a[1]), "conv_layer_bwd: bad kind / missing rulebook");
    if (M == 0) return B200SP_OK;
    const bool forked = need_dw && need_din && side_st != nullptr;
    int rc;
    if (need_dw) {
        void* wst = main_st;
        if (forked) {
            rc = b200sp_stream_fork(main_st, side_st, as_ptr<void>(a[28]));
            if (rc) return rc;
            wst = side_st;
        }
        rc = conv_wgrad(kind, rb, act, M, Cin, g, n_g, Cout, Kw, as_ptr<float>(a[22]), wst);
        if (rc) return rc;
    }
    if (need_din) {
        float* dy = as_ptr<float>(a[23]);
        const float* dx_add = n >= 33 ? as_ptr<const float>(a[32]) : nullptr;
        rc = conv_dgrad(kind, rb, g, n_g, Cin, w, Kw, dy, M, Cout, main_st, has_bn ? nullptr : dx_add);
        if (rc) return rc;
        if (a[30]) {
            const int64_t n4 = M * Cin / 4;
            B200SP_CHECK_ARG((M * Cin) % 4 == 0, "conv_layer_bwd: M * Cin must be a multiple of 4 to add grad_y");
            k_add_inplace<<<(unsigned)std::min<int64_t>(cdiv(n4, 256), 148 * 8), 256, 0, (cudaStream_t)main_st>>>(
                dy, as_ptr<const float>(a[30]), n4);
            B200SP_LAUNCH_CHECK();
        }
        if (has_bn) {
            const float* stats = as_ptr<const float>(a[14]);
            float* dwb = as_ptr<float>(a[25]);
            rc = b200sp_bn_bwd_add(as_ptr<const float>(a[2]), dy, M, Cin, as_ptr<const float>(a[12]), as_ptr<const float>(a[13]),
                                   stats, stats + Cin, 1, as_ptr<float>(a[24]), dwb, dwb + Cin, dx_add, as_ptr<void>(a[16]), a[17],
                                   main_st);
            if (rc) return rc;
        }
    }
    if (forked && !a[31]) return b200sp_stream_join(main_st, side_st, as_ptr<void>(a[29]));
    return B200SP_OK;
}
