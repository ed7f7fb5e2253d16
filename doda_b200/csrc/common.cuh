// common.cuh — shared helpers for libb200sparse (error plumbing, hashing, small device utils).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdarg.h>
#include "../../include/b200sparse.h"

namespace b200sp {

void set_error(const char* fmt, ...);
void note_kernel(const char* name);  // b200sp_last_kernel(): which kernel family a dispatching entry point chose

#define B200SP_CHECK_ARG(cond, ...)                  \
    do {                                             \
        if (!(cond)) {                               \
            b200sp::set_error(__VA_ARGS__);          \
            return B200SP_EINVAL;                    \
        }                                            \
    } while (0)

#define B200SP_CUDA(call)                                                                     \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            b200sp::set_error("%s:%d CUDA error: %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
            return B200SP_ECUDA;                                                              \
        }                                                                                     \
    } while (0)

// every kernel launch of this library is counted (bench.py reports it as gpu_launches)
void add_launches(int n);
#define B200SP_LAUNCH_CHECK_N(n)            \
    do {                                    \
        b200sp::add_launches(n);            \
        B200SP_CUDA(cudaGetLastError());    \
    } while (0)
#define B200SP_LAUNCH_CHECK() B200SP_LAUNCH_CHECK_N(1)

// developer knobs are read from the environment ONCE per call site (getenv walks the whole environment; the conv
// launch path used to do that seven times per launch)
#define B200SP_ENV_INT(var, name, dflt)                  \
    static const int var = [] {                          \
        const char* e__ = getenv(name);                  \
        return e__ ? atoi(e__) : (dflt);                 \
    }()

// ---- programmatic dependent launch (PDL) ----
// One training step is ~800 dependent launches of 5-100 us kernels: the ~2-4 us between the end of one kernel and the
// first useful instruction of the next (launch latency + prologue: barrier init, TMEM allocation) is a tenth of the
// step.  Kernels of the hot path therefore (a) call pdl_trigger() first thing -- the next kernel of the stream may
// become resident as soon as every CTA of this one has started -- and (b) call pdl_wait() before their FIRST global
// memory access (read or write): it returns once the preceding kernel of the stream has completed and its writes are
// visible.  Everything before pdl_wait() (shared-memory carve-up, mbarrier init, tcgen05.alloc) overlaps the
// predecessor's tail.  Launched without the attribute (B200SP_PDL=0) both calls are no-ops.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;\n" ::: "memory"); }

bool pdl_enabled();

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

static inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }
static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

int num_sms();

// ---- 64-bit key hash table: open addressing, linear probing ----
constexpr unsigned long long kEmptyKey = 0xFFFFFFFFFFFFFFFFull;
constexpr int kEmptyVal = 0x7F7F7F7F;  // memset(0x7F) pattern, > any row id

struct HashTab {
    unsigned long long* keys;
    int* vals;
    uint32_t mask;  // capacity - 1 (capacity is a power of two)
};

__device__ __forceinline__ uint32_t hash64(unsigned long long k) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdull;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ull;
    k ^= k >> 33;
    return (uint32_t)k;
}

// returns true if this call created the slot. vals[slot] = min(vals[slot], val) (deterministic on duplicates)
__device__ __forceinline__ bool hash_insert(const HashTab& t, unsigned long long key, int val, uint32_t* slot_out = nullptr) {
    uint32_t s = hash64(key) & t.mask;
    while (true) {
        unsigned long long prev = atomicCAS(&t.keys[s], kEmptyKey, key);
        if (prev == kEmptyKey || prev == key) {
            if (val >= 0) atomicMin(&t.vals[s], val);
            if (slot_out) *slot_out = s;
            return prev == kEmptyKey;
        }
        s = (s + 1) & t.mask;
    }
}

__device__ __forceinline__ int hash_find_slot(const HashTab& t, unsigned long long key) {
    uint32_t s = hash64(key) & t.mask;
    while (true) {
        unsigned long long k = __ldg(&t.keys[s]);
        if (k == key) return (int)s;
        if (k == kEmptyKey) return -1;
        s = (s + 1) & t.mask;
    }
}

__device__ __forceinline__ int hash_lookup(const HashTab& t, unsigned long long key) {
    int s = hash_find_slot(t, key);
    return s < 0 ? -1 : __ldg(&t.vals[s]);
}

static inline uint32_t hash_capacity(int64_t n) {
    uint64_t c = 1024;
    while (c < (uint64_t)n * 2) c <<= 1;
    return (uint32_t)c;
}

}  // namespace b200sp
