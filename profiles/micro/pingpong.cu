// Micro-benchmark: latency of an L2-mediated hand-off between two SMs with the same 16-byte relaxed.gpu accesses the
// dataflow contact sweep uses (st.relaxed.gpu.b128 -> polled ld.relaxed.gpu.b128), and the cost of grid.sync().
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;
__device__ __forceinline__ float4 ld128(const float4* p) {
    float4 v;
    asm volatile("{\n\t.reg .b128 t;\n\tld.relaxed.gpu.global.b128 t, [%4];\n\tmov.b128 {%0, %1, %2, %3}, t;\n\t}"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st128(float4* p, float4 v) {
    asm volatile("{\n\t.reg .b128 t;\n\tmov.b128 t, {%1, %2, %3, %4};\n\tst.relaxed.gpu.global.b128 [%0], t;\n\t}" ::"l"(p),
                 "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// block 0 and block `peer` bounce a counter `iters` times (one thread each)
__global__ void pingpong(float4* word, int iters, int peer, long long* cycles) {
    if (threadIdx.x != 0) return;
    if (blockIdx.x != 0 && blockIdx.x != peer) return;
    const int me = blockIdx.x == 0 ? 0 : 1;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        const unsigned want = 2u * i + me;  // I may write when version == want
        float4 v;
        do { v = ld128(word); } while (__float_as_uint(v.w) != want);
        v.x += 1.0f;
        v.w = __uint_as_float(want + 1u);
        st128(word, v);
    }
    if (me == 0) *cycles = clock64() - t0;
}
__global__ void gridsync(int iters, long long* cycles) {
    cg::grid_group g = cg::this_grid();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) g.sync();
    if (blockIdx.x == 0 && threadIdx.x == 0) *cycles = clock64() - t0;
}
int main() {
    float4* w; long long* c; cudaMalloc(&w, 64); cudaMalloc(&c, 8);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int iters = 20000;
    for (int peer : {1, 2, 37, 74, 100, 147}) {
        cudaMemset(w, 0, 64);
        pingpong<<<148, 32>>>(w, iters, peer, c);
        long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        printf("pingpong block0<->block%-3d: %.0f cycles per hand-off (%.3f us at %d MHz nominal)\n", peer, (double)h / (2.0 * iters), (double)h / (2.0 * iters) / (clk / 1e3), clk / 1000);
    }
    for (int bps : {1, 2, 4}) for (int tpb : {256, 1024}) {
        if (bps * tpb > 2048) continue;
        void* args[] = {(void*)&iters, (void*)&c};
        int it2 = 2000; args[0] = &it2;
        cudaError_t e = cudaLaunchCooperativeKernel((void*)gridsync, dim3(148 * bps), dim3(tpb), args, 0, 0);
        long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        printf("grid.sync %d blocks x %d thr: %.0f cycles (%.3f us)  [%s]\n", 148 * bps, tpb, (double)h / it2, (double)h / it2 / (clk / 1e3), cudaGetErrorString(e));
    }
    return 0;
}
