// Augmentation hot spots of the data pipeline on the device: elastic distortion, the crop masks
// and the 3x3 scene transform -- what dataset/augmentor/augmentor_utils.py:61-80 (elastic), 85-104 (scene_aug's final
// matmul) and 449-472 (crop) compute with numpy / scipy on one host core per worker.
//
// The arithmetic type follows the reference: scipy.ndimage.convolve accumulates in double and stores float32 per pass,
// RegularGridInterpolator evaluates in double, and `x + g(x) * mag` is a float64 array -- so the kernels below
// accumulate in double and the distorted coordinates are double.  All of it is HBM-bound streaming work on small
// arrays (noise grids of ~30^3 .. 90^3 cells, 1e5 .. 1e6 points); nothing here belongs on tensor cores.
#include "common.cuh"

namespace b200sp {

// one pass of the reference's separable blur: out = convolve(in, ones(3)/3 along `axis`, mode='constant', cval=0)
// (augmentor_utils.py:62-64, 68-73).  The 1/3 weight is the float32 the reference builds, widened to double.
__global__ void __launch_bounds__(256) k_box3_axis(const float* __restrict__ in, float* __restrict__ out, int64_t total, int nx,
                                                   int ny, int nz, int axis) {
    const double w = (double)(1.0f / 3.0f);
    const int64_t cell = (int64_t)nx * ny * nz;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i % cell;
        int pos, n;
        int64_t stride;
        if (axis == 0) {
            pos = (int)(r / ((int64_t)ny * nz)); n = nx; stride = (int64_t)ny * nz;
        } else if (axis == 1) {
            pos = (int)((r / nz) % ny); n = ny; stride = nz;
        } else {
            pos = (int)(r % nz); n = nz; stride = 1;
        }
        // convolve walks the mirrored footprint: +1, 0, -1
        double acc = 0.0;
        acc = __dadd_rn(acc, __dmul_rn(w, pos + 1 < n ? (double)__ldg(in + i + stride) : 0.0));
        acc = __dadd_rn(acc, __dmul_rn(w, (double)__ldg(in + i)));
        acc = __dadd_rn(acc, __dmul_rn(w, pos > 0 ? (double)__ldg(in + i - stride) : 0.0));
        out[i] = (float)acc;
    }
}

struct ElasticGrid {
    int n[3];
    double start[3], step[3], stop[3];  // np.linspace: axis[i] = i * step + start (unfused), axis[n - 1] = stop
    __device__ __forceinline__ double at(int d, int i) const {
        return i == n[d] - 1 ? stop[d] : __dadd_rn(__dmul_rn((double)i, step[d]), start[d]);
    }
};

// out[p, :] = x[p, :] + mag * (trilinear interpolation of the three smoothed noise grids at x[p, :])
// (RegularGridInterpolator(method='linear', bounds_error=0, fill_value=0), augmentor_utils.py:74-80)
template <typename T>
__global__ void __launch_bounds__(256) k_elastic_apply(const T* __restrict__ x, const float* __restrict__ noise, ElasticGrid g,
                                                       double mag, int64_t N, double* __restrict__ out) {
    const int64_t cell = (int64_t)g.n[0] * g.n[1] * g.n[2];
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < N; p += (int64_t)gridDim.x * blockDim.x) {
        double xv[3], t[3];
        int idx[3];
        bool inside = true;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            xv[d] = (double)x[p * 3 + d];
            if (xv[d] < g.start[d] || xv[d] > g.stop[d] || xv[d] != xv[d]) inside = false;
            // searchsorted(grid, x) - 1, clipped to [0, n - 2]: grid[i] < x <= grid[i + 1]
            int i = (int)ceil((xv[d] - g.start[d]) / g.step[d]) - 1;
            i = max(0, min(g.n[d] - 2, i));
            while (i > 0 && !(g.at(d, i) < xv[d])) --i;
            while (i < g.n[d] - 2 && g.at(d, i + 1) < xv[d]) ++i;
            const double lo = g.at(d, i), hi = g.at(d, i + 1);
            idx[d] = i;
            t[d] = (xv[d] - lo) / (hi - lo);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            double v = 0.0;
            if (inside) {
                const float* gr = noise + c * cell;
#pragma unroll
                for (int corner = 0; corner < 8; ++corner) {
                    const int e0 = corner >> 2, e1 = (corner >> 1) & 1, e2 = corner & 1;
                    const double wgt = (e0 ? t[0] : 1.0 - t[0]) * (e1 ? t[1] : 1.0 - t[1]) * (e2 ? t[2] : 1.0 - t[2]);
                    const int64_t off = ((int64_t)(idx[0] + e0) * g.n[1] + (idx[1] + e1)) * g.n[2] + (idx[2] + e2);
                    v += wgt * (double)__ldg(gr + off);
                }
            }
            out[p * 3 + c] = xv[c] + v * mag;
        }
    }
}

// crop (augmentor_utils.py:449-472): xyz_offset = xyz + offset; valid &= all(xyz_offset >= 0) & all(xyz_offset < full_scale);
// *count += number of valid points (integer atomics: exact and order-free)
__global__ void __launch_bounds__(256) k_crop_mask(const double* __restrict__ xyz, int64_t N, double o0, double o1, double o2,
                                                   double f0, double f1, double f2, unsigned char* __restrict__ valid,
                                                   double* __restrict__ xyz_offset, int* __restrict__ count) {
    int local = 0;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < N; p += (int64_t)gridDim.x * blockDim.x) {
        const double a = xyz[p * 3] + o0, b = xyz[p * 3 + 1] + o1, c = xyz[p * 3 + 2] + o2;
        if (xyz_offset) {
            xyz_offset[p * 3] = a;
            xyz_offset[p * 3 + 1] = b;
            xyz_offset[p * 3 + 2] = c;
        }
        const bool ok = valid[p] && fmin(a, fmin(b, c)) >= 0.0 && a < f0 && b < f1 && c < f2;
        valid[p] = ok ? 1 : 0;
        local += ok ? 1 : 0;
    }
    local = __reduce_add_sync(0xffffffffu, local);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(count, local);
}

// out = xyz @ m  (scene_aug's flip / jitter / rotation matrix, augmentor_utils.py:103)
template <typename T>
__global__ void __launch_bounds__(256) k_affine3(const T* __restrict__ xyz, int64_t N, double m00, double m01, double m02,
                                                 double m10, double m11, double m12, double m20, double m21, double m22,
                                                 double* __restrict__ out) {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < N; p += (int64_t)gridDim.x * blockDim.x) {
        const double a = (double)xyz[p * 3], b = (double)xyz[p * 3 + 1], c = (double)xyz[p * 3 + 2];
        // np.matmul row by row: x0*m0j + x1*m1j + x2*m2j, unfused
        out[p * 3] = __dadd_rn(__dadd_rn(__dmul_rn(a, m00), __dmul_rn(b, m10)), __dmul_rn(c, m20));
        out[p * 3 + 1] = __dadd_rn(__dadd_rn(__dmul_rn(a, m01), __dmul_rn(b, m11)), __dmul_rn(c, m21));
        out[p * 3 + 2] = __dadd_rn(__dadd_rn(__dmul_rn(a, m02), __dmul_rn(b, m12)), __dmul_rn(c, m22));
    }
}

static inline unsigned aug_grid(int64_t n) {
    int64_t g = (n + 255) / 256;
    const int64_t cap = (int64_t)num_sms() * 8;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (unsigned)g;
}

}  // namespace b200sp

using namespace b200sp;

extern "C" int b200sp_elastic_blur(float* noise_dev, float* scratch_dev, int nx, int ny, int nz, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    B200SP_CHECK_ARG(nx >= 1 && ny >= 1 && nz >= 1, "elastic_blur: empty grid");
    const int64_t total = 3ll * nx * ny * nz;
    // blur0, blur1, blur2, blur0, blur1, blur2 -- the reference's order; six ping-pong passes end in noise_dev
    float* a = noise_dev;
    float* b = scratch_dev;
    for (int pass = 0; pass < 6; ++pass) {
        k_box3_axis<<<aug_grid(total), 256, 0, st>>>(a, b, total, nx, ny, nz, pass % 3);
        float* t = a; a = b; b = t;
    }
    note_kernel("k_box3_axis");
    B200SP_LAUNCH_CHECK();
    return B200SP_OK;
}

extern "C" int b200sp_elastic_apply(const void* xyz_dev, int xyz_is_f64, int64_t N, const float* noise_dev, int nx, int ny, int nz,
                                    double gran, double mag, double* out_dev, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    B200SP_CHECK_ARG(N >= 0 && nx >= 2 && ny >= 2 && nz >= 2, "elastic_apply: need N>=0 and a grid of >= 2 cells per axis");
    if (N == 0) return B200SP_OK;
    ElasticGrid g;
    const int n[3] = {nx, ny, nz};
    for (int d = 0; d < 3; ++d) {
        g.n[d] = n[d];
        // np.linspace(-(b - 1) * gran, (b - 1) * gran, b): step = (stop - start) / (b - 1)
        const double stop = (double)(n[d] - 1) * gran;
        g.start[d] = -stop;
        g.stop[d] = stop;
        g.step[d] = (stop - g.start[d]) / (double)(n[d] - 1);
    }
    if (xyz_is_f64)
        k_elastic_apply<double><<<aug_grid(N), 256, 0, st>>>((const double*)xyz_dev, noise_dev, g, mag, N, out_dev);
    else
        k_elastic_apply<float><<<aug_grid(N), 256, 0, st>>>((const float*)xyz_dev, noise_dev, g, mag, N, out_dev);
    note_kernel("k_elastic_apply");
    B200SP_LAUNCH_CHECK();
    return B200SP_OK;
}

extern "C" int b200sp_crop_mask(const double* xyz_dev, int64_t N, const double* offset3_host, const double* full_scale3_host,
                                void* valid_u8_dev, double* xyz_offset_dev, int32_t* count_dev, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    B200SP_CHECK_ARG(N >= 0 && offset3_host && full_scale3_host && valid_u8_dev && count_dev, "crop_mask: null argument");
    B200SP_CUDA(cudaMemsetAsync(count_dev, 0, sizeof(int32_t), st));
    if (N == 0) return B200SP_OK;
    k_crop_mask<<<aug_grid(N), 256, 0, st>>>(xyz_dev, N, offset3_host[0], offset3_host[1], offset3_host[2], full_scale3_host[0],
                                             full_scale3_host[1], full_scale3_host[2], (unsigned char*)valid_u8_dev, xyz_offset_dev,
                                             count_dev);
    note_kernel("k_crop_mask");
    B200SP_LAUNCH_CHECK();
    return B200SP_OK;
}

extern "C" int b200sp_affine3(const void* xyz_dev, int xyz_is_f64, int64_t N, const double* m9_host, double* out_dev, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    B200SP_CHECK_ARG(N >= 0 && m9_host, "affine3: null matrix");
    if (N == 0) return B200SP_OK;
    const double* m = m9_host;
    if (xyz_is_f64)
        k_affine3<double><<<aug_grid(N), 256, 0, st>>>((const double*)xyz_dev, N, m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], m[8], out_dev);
    else
        k_affine3<float><<<aug_grid(N), 256, 0, st>>>((const float*)xyz_dev, N, m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], m[8], out_dev);
    note_kernel("k_affine3");
    B200SP_LAUNCH_CHECK();
    return B200SP_OK;
}
