// voxelize_gpu.cu — PG_OP.voxelize_idx on the device (batch assembly without the serial CPU hash map).
//
// Same results as lib/pointgroup_ops/src/voxelize/voxelize.cpp:62-155 / b200sp_voxelize_idx_cpu, bit for bit:
//   * voxel ids in FIRST-TOUCH order: voxel v is the v-th distinct coordinate met while scanning the points in order;
//   * input_map[p] = voxel of point p;  out_coords[v] = coordinates of the voxel (of its first point);
//   * output_map[v] = [count, points of v in ascending order ..., -1 padding] for modes 3/4 (maxActive = largest count),
//     [1, first point] for modes 0/1, [1, last point] for mode 2.
// How: hash the packed coordinate with atomicMin of the point index -> every voxel knows
// [packing: a range pass finds the largest coordinate; x, y, z take cb = bits(max) each (<= 20) and the batch index the
//  remaining 64 - 3 cb bits, so a 1024^3 grid leaves 34 bits for the batch index and only a 2^20-wide grid is held to
//  batch < 15; the all-ones key is the table's empty marker and stays unreachable]
// its first point; an exclusive scan over the "I am a first point" flags numbers the voxels in first-touch order; a
// stable radix sort of (voxel id, point) groups each voxel's points in ascending order.  Two calls (the host has to
// size out_coords / output_map from M and maxActive): _begin ... sync ... _finish, like the strided rulebook.
#include <algorithm>
#include <cub/cub.cuh>
#include "common.cuh"

namespace b200sp {
namespace {

struct WsCarverV {
    char* p;
    int64_t left;
    void* take(int64_t bytes) {
        bytes = align_up(bytes, 256);
        if (bytes > left) return nullptr;
        void* r = p;
        p += bytes;
        left -= bytes;
        return r;
    }
};

struct VoxWs {
    HashTab t;
    int* first;      // [N] first point of p's voxel
    int* flag;       // [N] 1 if p is the first point of its voxel
    int* scan;       // [N] exclusive scan of flag
    int* vid;        // [N] voxel id of p (= input_map)
    int* cnt;        // [N] points per voxel (first M used)
    int* start;      // [N] exclusive scan of cnt
    int* sorted_vid; // [N]
    int* sorted_p;   // [N]
    int* iota;       // [N]
    int* info;       // [4]: M, maxActive, bad-coordinate flag, duplicate flag
    void* cub_ws;
    size_t cub_bytes;
};

size_t vox_cub_bytes(int64_t N) {
    size_t a = 0, b = 0, c = 0;
    const int n = (int)(N > 0 ? N : 1);
    cub::DeviceScan::ExclusiveSum(nullptr, a, (int*)nullptr, (int*)nullptr, n);
    cub::DeviceRadixSort::SortPairs(nullptr, b, (int*)nullptr, (int*)nullptr, (int*)nullptr, (int*)nullptr, n);
    cub::DeviceReduce::Max(nullptr, c, (int*)nullptr, (int*)nullptr, n);
    return std::max(a, std::max(b, c));
}

int carve(VoxWs& w, int64_t N, void* ws, int64_t ws_bytes) {
    WsCarverV c{(char*)ws, ws_bytes};
    const int64_t n = N > 0 ? N : 1;
    const int64_t cap = hash_capacity(n);
    w.t.keys = (unsigned long long*)c.take(cap * 8);
    w.t.vals = (int*)c.take(cap * 4);
    w.t.mask = (uint32_t)(cap - 1);
    w.first = (int*)c.take(n * 4);
    w.flag = (int*)c.take(n * 4);
    w.scan = (int*)c.take(n * 4);
    w.vid = (int*)c.take(n * 4);
    w.cnt = (int*)c.take(n * 4);
    w.start = (int*)c.take(n * 4);
    w.sorted_vid = (int*)c.take(n * 4);
    w.sorted_p = (int*)c.take(n * 4);
    w.iota = (int*)c.take(n * 4);
    w.info = (int*)c.take(256);
    w.cub_bytes = vox_cub_bytes(N);
    w.cub_ws = c.take((int64_t)w.cub_bytes);
    if (!w.t.keys || !w.t.vals || !w.first || !w.flag || !w.scan || !w.vid || !w.cnt || !w.start || !w.sorted_vid ||
        !w.sorted_p || !w.iota || !w.info || (!w.cub_ws && w.cub_bytes)) {
        set_error("voxelize_idx_gpu: workspace too small (%lld bytes)", (long long)ws_bytes);
        return B200SP_ENOMEM;
    }
    return B200SP_OK;
}

// info[4] = largest coordinate, info[5] = largest batch index (k_vox_range); -> bits per coordinate, or -1 when the
// batch does not fit into what the coordinates leave of the 64-bit key
__device__ __forceinline__ int key_layout(const int* info) {
    const int cmax = info[4], bmax = info[5];
    const int cb = 32 - __clz(cmax);  // 0 for an all-zero grid
    if (cb > 20) return -1;
    const int bb = 64 - 3 * cb;
    if (bb < 32 && (long long)bmax >= (1ll << bb) - 1) return -1;
    return cb;
}

__device__ __forceinline__ bool pack_key(const long long* c, int ncol, int cb, unsigned long long& key) {
    const long long b = ncol == 4 ? c[0] : 0;
    const long long x = c[ncol - 3], y = c[ncol - 2], z = c[ncol - 1];
    if (cb < 0 || b < 0 || x < 0 || y < 0 || z < 0) return false;
    key = ((unsigned long long)b << (3 * cb)) | ((unsigned long long)x << (2 * cb)) | ((unsigned long long)y << cb) | (unsigned long long)z;
    return true;
}

// largest coordinate / batch index of the call; negative or >= 2^31 values raise the bad-coordinate flag
__global__ void k_vox_range(const long long* __restrict__ coords, long long N, int ncol, int* info) {
    const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    int cm = 0, bm = 0, bad = 0;
    if (p < N) {
        const long long* c = coords + p * ncol;
        const long long b = ncol == 4 ? c[0] : 0;
        const long long x = c[ncol - 3], y = c[ncol - 2], z = c[ncol - 1];
        const long long hi = max(x, max(y, z)), lo = min(min(x, y), min(z, b));
        if (lo < 0 || hi >= (1ll << 31) || b >= (1ll << 31)) bad = 1;
        else { cm = (int)hi; bm = (int)b; }
    }
    cm = __reduce_max_sync(0xffffffffu, cm);
    bm = __reduce_max_sync(0xffffffffu, bm);
    bad = __reduce_max_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0) {
        if (cm > info[4]) atomicMax(&info[4], cm);
        if (bm > info[5]) atomicMax(&info[5], bm);
        if (bad) info[2] = 1;
    }
}

__global__ void k_vox_insert(const long long* __restrict__ coords, long long N, int ncol, HashTab t, int* info) {
    const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (p >= N) return;
    unsigned long long key;
    if (!pack_key(coords + p * ncol, ncol, key_layout(info), key)) {
        info[2] = 1;
        return;
    }
    hash_insert(t, key, (int)p);  // vals[slot] = min point index of the voxel
}

__global__ void k_vox_first(const long long* __restrict__ coords, long long N, int ncol, HashTab t, int* __restrict__ first,
                            int* __restrict__ flag, int* __restrict__ iota, const int* __restrict__ info) {
    const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (p >= N) return;
    unsigned long long key;
    int f = (int)p;
    if (pack_key(coords + p * ncol, ncol, key_layout(info), key)) {
        const int v = hash_lookup(t, key);
        if (v >= 0) f = v;
    }
    first[p] = f;
    flag[p] = f == (int)p ? 1 : 0;
    iota[p] = (int)p;
}

// vid[p] = scan[first[p]]; count the points of each voxel; M = scan[N-1] + flag[N-1]
__global__ void k_vox_ids(long long N, const int* __restrict__ first, const int* __restrict__ flag,
                          const int* __restrict__ scan, int* __restrict__ vid, int* __restrict__ cnt, int* info) {
    const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (p >= N) return;
    const int v = scan[first[p]];
    vid[p] = v;
    if (atomicAdd(&cnt[v], 1) >= 1) info[3] = 1;  // some voxel holds more than one point (mode 0 forbids it)
    if (p == N - 1) info[0] = scan[p] + flag[p];
}

// sorted (voxel, point) pairs, points ascending inside a voxel -> output_map rows and the voxel's coordinates
__global__ void k_vox_fill(long long N, int ncol, int mode, int W, const long long* __restrict__ coords,
                           const int* __restrict__ sorted_vid, const int* __restrict__ sorted_p,
                           const int* __restrict__ start, const int* __restrict__ cnt, long long* __restrict__ out_coords,
                           int* __restrict__ output_map) {
    const long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (q >= N) return;
    const int v = sorted_vid[q], p = sorted_p[q];
    const int rank = (int)q - start[v], n = cnt[v];
    int* row = output_map + (long long)v * W;
    if (mode == 3 || mode == 4) {
        if (rank == 0) row[0] = n;
        row[1 + rank] = p;
    } else if (mode == 2) {  // back(): the last point
        if (rank == n - 1) { row[0] = 1; row[1] = p; }
    } else {                 // 0, 1: front(): the first point
        if (rank == 0) { row[0] = 1; row[1] = p; }
    }
    const bool rep = (mode == 2) ? (rank == n - 1) : (rank == 0);  // out_coords[v] = coords[output_map[v][1]]
    if (rep)
        for (int j = 0; j < ncol; ++j) out_coords[(long long)v * ncol + j] = coords[(long long)p * ncol + j];
}

}  // namespace
}  // namespace b200sp

using namespace b200sp;

extern "C" int64_t b200sp_voxelize_idx_gpu_ws_bytes(int64_t N) {
    const int64_t n = N > 0 ? N : 1;
    const int64_t cap = hash_capacity(n);
    return align_up(cap * 8, 256) + align_up(cap * 4, 256) + 9 * align_up(n * 4, 256) + 256 +
           align_up((int64_t)vox_cub_bytes(N), 256) + 4096;
}

// first half: voxel ids (input_map), M and maxActive; starts an async copy of info[4] = {M, maxActive, bad, dup} to
// info_host (pinned).  The caller synchronises, sizes out_coords [M, ncol] and output_map [M, 1 + maxActive], and
// calls _finish with the SAME workspace.
extern "C" int b200sp_voxelize_idx_gpu_begin(const int64_t* coords, int64_t N, int ncol, int mode, int32_t* input_map,
                                             int32_t* info_host, void* ws, int64_t ws_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    B200SP_CHECK_ARG(ncol == 3 || ncol == 4, "voxelize_idx_gpu: coords must have 3 or 4 columns (got %d)", ncol);
    B200SP_CHECK_ARG(mode >= 0 && mode <= 4, "voxelize_idx_gpu: mode %d not in 0..4", mode);
    B200SP_CHECK_ARG(N >= 0 && N < (1ll << 31) && info_host, "voxelize_idx_gpu: bad arguments");
    info_host[0] = 0; info_host[1] = 1; info_host[2] = 0; info_host[3] = 0;
    if (N == 0) return B200SP_OK;
    B200SP_CHECK_ARG(coords && input_map, "voxelize_idx_gpu: null pointer");
    VoxWs w;
    int rc = carve(w, N, ws, ws_bytes);
    if (rc) return rc;
    const int64_t cap = (int64_t)w.t.mask + 1;
    B200SP_CUDA(cudaMemsetAsync(w.t.keys, 0xFF, cap * 8, st));
    B200SP_CUDA(cudaMemsetAsync(w.t.vals, 0x7F, cap * 4, st));
    B200SP_CUDA(cudaMemsetAsync(w.cnt, 0, N * 4, st));
    B200SP_CUDA(cudaMemsetAsync(w.info, 0, 32, st));
    const unsigned grid = (unsigned)cdiv(N, 256);
    k_vox_range<<<grid, 256, 0, st>>>((const long long*)coords, N, ncol, w.info);
    k_vox_insert<<<grid, 256, 0, st>>>((const long long*)coords, N, ncol, w.t, w.info);
    k_vox_first<<<grid, 256, 0, st>>>((const long long*)coords, N, ncol, w.t, w.first, w.flag, w.iota, w.info);
    B200SP_CUDA(cub::DeviceScan::ExclusiveSum(w.cub_ws, w.cub_bytes, w.flag, w.scan, (int)N, st));
    k_vox_ids<<<grid, 256, 0, st>>>(N, w.first, w.flag, w.scan, w.vid, w.cnt, w.info);
    // maxActive = max count (modes 3/4 only; 1 otherwise, as the reference sizes output_map)
    B200SP_CUDA(cub::DeviceReduce::Max(w.cub_ws, w.cub_bytes, w.cnt, w.info + 1, (int)N, st));
    B200SP_LAUNCH_CHECK_N(4 + 2);
    B200SP_CUDA(cudaMemcpyAsync(input_map, w.vid, N * 4, cudaMemcpyDeviceToDevice, st));
    B200SP_CUDA(cudaMemcpyAsync(info_host, w.info, 16, cudaMemcpyDeviceToHost, st));
    return B200SP_OK;
}

extern "C" int b200sp_voxelize_idx_gpu_finish(const int64_t* coords, int64_t N, int ncol, int mode, int64_t M,
                                              int max_active, int64_t* out_coords, int32_t* output_map, void* ws,
                                              int64_t ws_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    B200SP_CHECK_ARG(ncol == 3 || ncol == 4, "voxelize_idx_gpu: coords must have 3 or 4 columns (got %d)", ncol);
    B200SP_CHECK_ARG(mode >= 0 && mode <= 4, "voxelize_idx_gpu: mode %d not in 0..4", mode);
    B200SP_CHECK_ARG(N >= 0 && M >= 0 && M <= N && max_active >= 1, "voxelize_idx_gpu_finish: bad sizes");
    if (N == 0 || M == 0) return B200SP_OK;
    B200SP_CHECK_ARG(coords && out_coords && output_map, "voxelize_idx_gpu_finish: null pointer");
    VoxWs w;
    int rc = carve(w, N, ws, ws_bytes);
    if (rc) return rc;
    const int W = max_active + 1;
    B200SP_CUDA(cudaMemsetAsync(output_map, 0xFF, (size_t)M * W * 4, st));  // -1 padding
    int bits = 1;
    while (bits < 31 && (M >> bits)) ++bits;
    // stable sort by voxel id: the points of a voxel stay in ascending order
    B200SP_CUDA(cub::DeviceRadixSort::SortPairs(w.cub_ws, w.cub_bytes, w.vid, w.sorted_vid, w.iota, w.sorted_p, (int)N, 0,
                                                bits, st));
    B200SP_CUDA(cub::DeviceScan::ExclusiveSum(w.cub_ws, w.cub_bytes, w.cnt, w.start, (int)M, st));
    k_vox_fill<<<(unsigned)cdiv(N, 256), 256, 0, st>>>(N, ncol, mode, W, (const long long*)coords, w.sorted_vid, w.sorted_p,
                                                      w.start, w.cnt, (long long*)out_coords, output_map);
    B200SP_LAUNCH_CHECK_N(1 + 4);
    return B200SP_OK;
}
