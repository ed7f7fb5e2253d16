// core.cu — error plumbing + the CPU-side entry points (voxelize_idx, bfs_cluster).
// The CPU functions never touch the CUDA runtime: they run inside forked DataLoader workers
// (dataset/dataset.py:182 of the reference calls voxelization_idx from collate_fn).
#include "common.cuh"
#include <string.h>
#include <unordered_map>
#include <vector>
#include <queue>
#include <atomic>

namespace b200sp {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static thread_local const char* g_kernel = "";
void note_kernel(const char* name) { g_kernel = name; }

static std::atomic<long long> g_launches{0};
void add_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
long long get_launches() { return g_launches.load(std::memory_order_relaxed); }

bool pdl_enabled() {
    static int on = -1;
    if (on < 0) {
        // ON by default (B200SP_PDL=0 turns it off).  While the step was host-bound it measured as nothing or a loss
        // (12.95-13.45 ms against 12.80 ms); with the taped U-Net (doda_b200/tape.py) the step is GPU-bound and the
        // ~1.5 us between dependent kernels is worth 12.01 -> 11.76 ms per step over ~610 launches
        const char* e = getenv("B200SP_PDL");
        on = (e && e[0] == '0') ? 0 : 1;
    }
    return on != 0;
}

int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
            n = 148;
    }
    return n;
}

struct Key4 {
    int64_t b, x, y, z;
    bool operator==(const Key4& o) const { return b == o.b && x == o.x && y == o.y && z == o.z; }
};
struct Key4Hash {
    size_t operator()(const Key4& k) const {
        uint64_t h = 1469598103934665603ull;
        auto mix = [&](int64_t v) {
            h ^= (uint64_t)v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
        };
        mix(k.b); mix(k.x); mix(k.y); mix(k.z);
        return (size_t)h;
    }
};

}  // namespace b200sp

using namespace b200sp;

extern "C" const char* b200sp_last_error(void) { return g_err; }
extern "C" const char* b200sp_last_kernel(void) { return g_kernel; }
extern "C" int b200sp_version(void) { return 200; }
extern "C" int64_t b200sp_launch_count(void) { return (int64_t)b200sp::get_launches(); }

// Fork / join of a side stream around independent kernels of one layer (the weight gradient next to dgrad + BN
// backward): side waits for everything queued on main so far / main waits for everything queued on side so far.
// The event handle belongs to the caller (one per direction is enough: a wait captures the record it follows).
extern "C" int b200sp_event_create(void** event_out) {
    B200SP_CHECK_ARG(event_out, "event_create: null");
    cudaEvent_t ev;
    B200SP_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    *event_out = (void*)ev;
    return B200SP_OK;
}
extern "C" int b200sp_stream_fork(void* main_stream, void* side_stream, void* event) {
    B200SP_CHECK_ARG(event, "stream_fork: null event");
    B200SP_CUDA(cudaEventRecord((cudaEvent_t)event, (cudaStream_t)main_stream));
    B200SP_CUDA(cudaStreamWaitEvent((cudaStream_t)side_stream, (cudaEvent_t)event, 0));
    return B200SP_OK;
}
extern "C" int b200sp_stream_join(void* main_stream, void* side_stream, void* event) {
    B200SP_CHECK_ARG(event, "stream_join: null event");
    B200SP_CUDA(cudaEventRecord((cudaEvent_t)event, (cudaStream_t)side_stream));
    B200SP_CUDA(cudaStreamWaitEvent((cudaStream_t)main_stream, (cudaEvent_t)event, 0));
    return B200SP_OK;
}

// output_map rows [count, points ..., -1 pad] and the voxel coordinates from the point -> voxel map
static void vox_fill_cpu(const int64_t* coords, int64_t N, int ncol, int mode, const int32_t* p2v, int64_t M, int maxActive,
                         int64_t* out_coords, int32_t* output_map) {
    const int W = maxActive + 1;
    for (int64_t v = 0; v < M; ++v) {
        int32_t* r = output_map + v * W;
        r[0] = 0;
        for (int j = 1; j < W; ++j) r[j] = -1;
    }
    for (int64_t i = 0; i < N; ++i) {
        int32_t* r = output_map + (int64_t)p2v[i] * W;
        if (mode == 3 || mode == 4) {
            r[++r[0]] = (int32_t)i;
        } else if (mode == 2) {  // back(): last point wins
            r[0] = 1;
            r[1] = (int32_t)i;
        } else {  // 0, 1: front(): first point
            if (r[0] == 0) {
                r[0] = 1;
                r[1] = (int32_t)i;
            }
        }
    }
    for (int64_t v = 0; v < M; ++v) {
        const int64_t* c = coords + (int64_t)output_map[v * W + 1] * ncol;
        for (int j = 0; j < ncol; ++j) out_coords[v * ncol + j] = c[j];
    }
}

// Semantics follow lib/pointgroup_ops/src/voxelize/voxelize.cpp:62-155 (first-touch voxel ids while scanning
// points in order; per-batch maps; output_map rows = [count, pt..., -1 pad]; coords of the first point) and
// voxelize_outputmap 35-52.  Modes: 0 unique, 1 first point, 2 last point, 3 sum, 4 mean (the code, not the
// swapped comment at voxelize.cpp:54).  Note the reference zero-fills output_map before writing and pads with -1
// only up to maxActive (voxelize.cpp:20-21,36-41).
extern "C" int b200sp_voxelize_idx_cpu(const int64_t* coords, int64_t N, int ncol, int batch_size, int mode,
                                       int64_t* out_coords, int32_t* input_map, int32_t* output_map, int64_t* M_out,
                                       int32_t* max_active_out) {
    (void)batch_size;
    B200SP_CHECK_ARG(ncol == 3 || ncol == 4, "voxelize_idx: coords must have 3 or 4 columns (got %d)", ncol);
    B200SP_CHECK_ARG(mode >= 0 && mode <= 4, "voxelize_idx: mode %d not in 0..4", mode);
    B200SP_CHECK_ARG(N >= 0 && M_out && max_active_out, "voxelize_idx: bad arguments");
    std::vector<int32_t> p2v((size_t)N);
    std::vector<int32_t> count;
    // Fast path (what DODA feeds: batch < 2^15, voxel coordinates < 2^16): the four columns pack into one 64-bit key
    // and go through a flat open-addressing table (linear probing, load <= 0.5) -- ~15x faster than a node-based
    // map at 450 k points.  Anything else (negative or huge coordinates) takes the general map below.
    bool packable = true;
    for (int64_t i = 0; i < N && packable; ++i) {
        const int64_t* c = coords + i * ncol;
        for (int j = 0; j < ncol; ++j)
            if (c[j] < 0 || c[j] >= (ncol == 4 && j == 0 ? (1 << 15) : (1 << 16))) packable = false;
    }
    if (packable) {
        uint64_t cap = 1024;
        while (cap < (uint64_t)N * 2) cap <<= 1;
        const uint64_t mask = cap - 1, kEmpty = ~0ull;
        std::vector<uint64_t> keys((size_t)cap, kEmpty);
        std::vector<int32_t> vals((size_t)cap);
        count.reserve((size_t)N / 2 + 16);
        for (int64_t i = 0; i < N; ++i) {
            const int64_t* c = coords + i * ncol;
            const uint64_t key = ncol == 4 ? ((uint64_t)c[0] << 48) | ((uint64_t)c[1] << 32) | ((uint64_t)c[2] << 16) | (uint64_t)c[3]
                                           : ((uint64_t)c[0] << 32) | ((uint64_t)c[1] << 16) | (uint64_t)c[2];
            uint64_t h = key;
            h ^= h >> 33; h *= 0xff51afd7ed558ccdull; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ull; h ^= h >> 33;
            uint64_t s = h & mask;
            int32_t v;
            while (true) {
                if (keys[s] == key) { v = vals[s]; break; }
                if (keys[s] == kEmpty) {
                    v = (int32_t)count.size();
                    keys[s] = key; vals[s] = v;
                    count.push_back(0);
                    break;
                }
                s = (s + 1) & mask;
            }
            count[v]++;
            p2v[i] = v;
        }
    } else {
        std::unordered_map<Key4, int32_t, Key4Hash> mp;
        mp.reserve((size_t)N);
        for (int64_t i = 0; i < N; ++i) {
            const int64_t* c = coords + i * ncol;
            Key4 k = ncol == 4 ? Key4{c[0], c[1], c[2], c[3]} : Key4{0, c[0], c[1], c[2]};
            auto it = mp.find(k);
            int32_t v;
            if (it == mp.end()) {
                v = (int32_t)count.size();
                mp.emplace(k, v);
                count.push_back(0);
            } else {
                v = it->second;
            }
            count[v]++;
            p2v[i] = v;
        }
    }
    const int64_t M = (int64_t)count.size();
    int32_t maxActive = 1;
    if (mode == 3 || mode == 4)
        for (int32_t c : count) maxActive = c > maxActive ? c : maxActive;
    if (mode == 0) {
        for (int32_t c : count)
            if (c != 1) {
                set_error("voxelize_idx: mode 0 requires unique coordinates");
                return B200SP_EINVAL;
            }
    }
    *M_out = M;
    *max_active_out = maxActive;
    if (!out_coords && !output_map) {  // size query; with input_map given it also returns the point -> voxel map,
        if (input_map)                 // and b200sp_voxelize_idx_cpu_fill finishes without hashing again
            for (int64_t i = 0; i < N; ++i) input_map[i] = p2v[i];
        return B200SP_OK;
    }
    B200SP_CHECK_ARG(out_coords && input_map && output_map, "voxelize_idx: all three outputs are required");
    for (int64_t i = 0; i < N; ++i) input_map[i] = p2v[i];
    vox_fill_cpu(coords, N, ncol, mode, input_map, M, maxActive, out_coords, output_map);
    return B200SP_OK;
}

// second half of the two-call protocol when the first call was given input_map: builds output_map / out_coords from
// the point -> voxel map alone (counting, no hash)
extern "C" int b200sp_voxelize_idx_cpu_fill(const int64_t* coords, int64_t N, int ncol, int mode, const int32_t* input_map,
                                            int64_t M, int32_t max_active, int64_t* out_coords, int32_t* output_map) {
    B200SP_CHECK_ARG(ncol == 3 || ncol == 4, "voxelize_idx: coords must have 3 or 4 columns (got %d)", ncol);
    B200SP_CHECK_ARG(mode >= 0 && mode <= 4, "voxelize_idx: mode %d not in 0..4", mode);
    B200SP_CHECK_ARG(N >= 0 && M >= 0 && M <= N && max_active >= 1, "voxelize_idx_fill: bad sizes");
    if (N == 0) return B200SP_OK;
    B200SP_CHECK_ARG(coords && input_map && out_coords && output_map, "voxelize_idx_fill: null pointer");
    for (int64_t i = 0; i < N; ++i)
        B200SP_CHECK_ARG(input_map[i] >= 0 && input_map[i] < M, "voxelize_idx_fill: input_map[%lld] out of range", (long long)i);
    vox_fill_cpu(coords, N, ncol, mode, input_map, M, max_active, out_coords, output_map);
    return B200SP_OK;
}

// Semantics follow lib/pointgroup_ops/src/bfs_cluster/bfs_cluster.cpp:28-111: BFS over the ball-query
// adjacency restricted to equal semantic labels, points in discovery order, clusters kept when
// size >= threshold (bfs_cluster.cpp:67); output cluster_idxs [sumNPoint, 2] = (cluster id, point id),
// cluster_offsets [nCluster+1].
extern "C" int b200sp_bfs_cluster_cpu(const int32_t* sem, const int32_t* idx, const int32_t* start_len, int N,
                                      int threshold, int32_t* cluster_idxs, int32_t* cluster_offsets,
                                      int64_t* n_idx_out, int64_t* n_cluster_out) {
    B200SP_CHECK_ARG(N >= 0 && n_idx_out && n_cluster_out, "bfs_cluster: bad arguments");
    std::vector<char> visited((size_t)N, 0);
    std::vector<std::vector<int32_t>> clusters;
    int64_t total = 0;
    for (int i = 0; i < N; ++i) {
        if (visited[i]) continue;
        std::vector<int32_t> cl;
        std::queue<int32_t> q;
        q.push(i);
        visited[i] = 1;
        const int32_t label = sem[i];
        while (!q.empty()) {
            int32_t a = q.front();
            q.pop();
            cl.push_back(a);
            int32_t s = start_len[a * 2], l = start_len[a * 2 + 1];
            for (int32_t t = s; t < s + l; ++t) {
                int32_t nb = idx[t];
                if (sem[nb] != label) continue;
                if (!visited[nb]) {
                    visited[nb] = 1;
                    q.push(nb);
                }
            }
        }
        if ((int)cl.size() >= threshold) {
            total += (int64_t)cl.size();
            clusters.push_back(std::move(cl));
        }
    }
    *n_idx_out = total;
    *n_cluster_out = (int64_t)clusters.size();
    if (!cluster_idxs && !cluster_offsets) return B200SP_OK;
    B200SP_CHECK_ARG(cluster_idxs && cluster_offsets, "bfs_cluster: both outputs are required");
    int64_t pos = 0;
    cluster_offsets[0] = 0;
    for (size_t c = 0; c < clusters.size(); ++c) {
        for (int32_t pt : clusters[c]) {
            cluster_idxs[pos * 2 + 0] = (int32_t)c;
            cluster_idxs[pos * 2 + 1] = pt;
            ++pos;
        }
        cluster_offsets[c + 1] = (int32_t)pos;
    }
    return B200SP_OK;
}
