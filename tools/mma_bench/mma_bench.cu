// dev microbenchmark: cycles per tcgen05.mma for small shapes / operand layouts (one CTA per SM, one issuing thread).
// usage: mma_bench kind(0 tf32,1 bf16) M N amaj bmaj layout(0,2,4,6) lboA sboA lboB sboB ts(0/1) nacc reps
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../doda_b200/csrc/tc_common.cuh"
using namespace b200sp::tc;

struct Cfg { int kind, M, N, amaj, bmaj, layout, lboA, sboA, lboB, sboB, ts, nacc, reps; int elect; int smem_kb; int nwarps; int tcols; };
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, %1;\n\t@px mov.s32 %0, 1;\n\t}\n" : "+r"(pred) : "r"(0xFFFFFFFFu));
    return pred != 0;
}

__device__ __forceinline__ uint64_t desc_l(uint32_t addr, uint32_t lbo, uint32_t sbo, int layout) {
    return make_desc(addr, lbo, sbo) | ((uint64_t)layout << 61);
}
__device__ __forceinline__ void mma_ts(int kind, uint32_t d, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    if (kind == 0)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
                     ::"r"(d), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
                     ::"r"(d), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}

__global__ void __launch_bounds__(128) k_bench(Cfg c, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bars[4];
    uint64_t& bar = bars[threadIdx.x >> 5];
    __shared__ uint32_t s_tmem;
    for (int i = threadIdx.x; i < c.smem_kb * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.f;
    if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1); mbar_fence_init(); }
    if (threadIdx.x < 32) tmem_alloc(&s_tmem, c.tcols);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    if (c.elect && threadIdx.x < 32 * c.nwarps) {
        const int w = threadIdx.x >> 5;
        const uint32_t idesc = c.kind == 0 ? make_idesc_tf32(c.M, c.N, c.amaj, c.bmaj) : make_idesc_bf16(c.M, c.N, c.amaj, c.bmaj);
        const uint32_t a = smem_u32(smem), b = smem_u32(smem + c.smem_kb * 512);
        const uint64_t da = desc_l(a, c.lboA, c.sboA, c.layout), db = desc_l(b, c.lboB, c.sboB, c.layout);
        long long t0 = clock64();
        for (int r = 0; r < c.reps; ++r) {
            const uint32_t d = tmem + (uint32_t)(w * c.nacc * c.N + (r % c.nacc) * c.N);
            if (elect_one()) {
                if (c.ts) mma_ts(c.kind, d, tmem + (uint32_t)(c.tcols / 2), db, idesc, 1u);
                else if (c.kind == 0) mma_tf32_ss(d, da, db, idesc, 1u);
                else mma_bf16_ss(d, da, db, idesc, 1u);
            }
            __syncwarp();
        }
        if (elect_one()) mma_commit(&bar);
        __syncwarp();
        mbar_wait(&bar, 0);
        long long t1 = clock64();
        if ((threadIdx.x & 31) == 0) atomicMax((unsigned long long*)&out[blockIdx.x], (unsigned long long)(t1 - t0));
    } else if (!c.elect && threadIdx.x == 0) {
        const uint32_t idesc = c.kind == 0 ? make_idesc_tf32(c.M, c.N, c.amaj, c.bmaj) : make_idesc_bf16(c.M, c.N, c.amaj, c.bmaj);
        const uint32_t a = smem_u32(smem), b = smem_u32(smem + 96 * 1024);
        const uint64_t da = desc_l(a, c.lboA, c.sboA, c.layout), db = desc_l(b, c.lboB, c.sboB, c.layout);
        long long t0 = clock64();
        for (int r = 0; r < c.reps; ++r) {
            const uint32_t d = tmem + (uint32_t)((r % c.nacc) * c.N);
            if (c.ts) mma_ts(c.kind, d, tmem + 256u, db, idesc, 1u);
            else if (c.kind == 0) mma_tf32_ss(d, da, db, idesc, 1u);
            else mma_bf16_ss(d, da, db, idesc, 1u);
        }
        mma_commit(&bar);
        mbar_wait(&bar, 0);
        long long t1 = clock64();
        out[blockIdx.x] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem, c.tcols);
}

int main(int argc, char** argv) {
    if (argc < 14) { printf("args\n"); return 1; }
    Cfg c;
    int* f = &c.kind;
    for (int i = 0; i < 13; ++i) f[i] = atoi(argv[1 + i]);
    c.elect = argc > 15 ? atoi(argv[15]) : 0;
    int cps = argc > 16 ? atoi(argv[16]) : 1;
    c.nwarps = argc > 17 ? atoi(argv[17]) : 1;
    c.smem_kb = 160 / cps; c.tcols = 512 / cps;
    long long* out;
    cudaMalloc(&out, 1024 * sizeof(long long));
    cudaMemset(out, 0, 1024 * sizeof(long long));
    cudaFuncSetAttribute(k_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, c.smem_kb * 1024);
    int grid = (argc > 14 ? atoi(argv[14]) : 148) * cps;
    k_bench<<<grid, 128, c.smem_kb * 1024>>>(c, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("ERR %s\n", cudaGetErrorString(e)); return 2; }
    long long h[1024];
    cudaMemcpy(h, out, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    double s = 0;
    for (int i = 0; i < grid; ++i) s += h[i];
    printf("kind %d M %3d N %3d maj %d%d layout %d lboA %5d sboA %5d lboB %5d sboB %5d ts %d nacc %d elect %d cta/sm %d warps %d : %7.1f cyc/MMA/issuer\n", c.kind, c.M, c.N,
           c.amaj, c.bmaj, c.layout, c.lboA, c.sboA, c.lboB, c.sboB, c.ts, c.nacc, c.elect, cps, c.nwarps, s / grid / c.reps);
    return 0;
}
