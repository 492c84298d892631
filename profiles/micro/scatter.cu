// Micro-benchmark: chip-wide rate of SCATTERED 16-byte accesses to an L2-resident array (the body-word pattern of the
// contact sweep: 2 loads + 2 stores per manifold), for weak/.cg/relaxed.gpu flavours, plus a coalesced stream for scale.
#include <cstdio>
#include <cuda_runtime.h>
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
__device__ __forceinline__ unsigned hash(unsigned x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }
// MODE 0: plain ld/st, 1: ld.cg + st, 2: relaxed.gpu b128, 3: coalesced plain (index = consecutive)
template <int MODE>
__global__ void k(float4* a, unsigned n, int iters, float4* sink) {
    const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    float4 acc = make_float4(0, 0, 0, 0);
    for (int it = 0; it < iters; ++it) {
        const unsigned i = MODE == 3 ? (tid + it * nth) % n : hash(tid * 2654435761u + it) % n;
        float4 v;
        if (MODE == 0 || MODE == 3) v = a[i];
        else if (MODE == 1) v = __ldcg(a + i);
        else v = ld128(a + i);
        v.x += 1.0f;
        if (MODE == 2) st128(a + i, v); else a[i] = v;
        acc.x += v.y;
    }
    if (acc.x == 123.456f) *sink = acc;
}
template <int MODE> void run(const char* name, float4* a, unsigned n, float4* sink, int blocks, int tpb) {
    const int iters = 64;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, tpb>>>(a, n, iters, sink);
    cudaEventRecord(e0);
    k<MODE><<<blocks, tpb>>>(a, n, iters, sink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double pairs = (double)blocks * tpb * iters;
    printf("%-28s grid %4d x %4d: %7.1f ld+st pairs/ns  (%6.2f TB/s of 16 B payload, %.1f us)\n", name, blocks, tpb, pairs / (ms * 1e6), pairs * 32 / (ms * 1e9) , ms * 1e3);
}
int main() {
    const unsigned n = 100003;  // 1.6 MB: the body array of pile100k
    float4 *a, *sink; cudaMalloc(&a, (size_t)n * 16); cudaMalloc(&sink, 16); cudaMemset(a, 0, (size_t)n * 16);
    for (int bps : {1, 4, 8}) {
        run<0>("scattered plain ld/st", a, n, sink, 148 * bps, 256);
        run<1>("scattered ld.cg / st", a, n, sink, 148 * bps, 256);
        run<2>("scattered relaxed.gpu b128", a, n, sink, 148 * bps, 256);
        run<3>("coalesced plain ld/st", a, n, sink, 148 * bps, 256);
    }
    return 0;
}
