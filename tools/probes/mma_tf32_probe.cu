// Issue rate of the legacy tensor-core path on sm_100a for the shapes a K2 (Q = phi . W) variant would use:
// mma.sync.aligned.m16n8k8 tf32 (fp32 accumulate) and m16n8k16 bf16, 4 independent accumulator tiles per warp.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes/mma_tf32_probe tools/probes/mma_tf32_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <int KIND>
__global__ void k(float *out, int iters) {
    float c[4][4] = {};
    uint32_t a[4] = {0x3f800000u + threadIdx.x, 0x3f000000u, 0x3e800000u, 0x3f400000u};
    uint32_t b[2] = {0x3f800000u, 0x3f000000u + threadIdx.x};
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (KIND == 0) mma_tf32(c[j], a, b); else mma_bf16(c[j], a, b);
        }
    }
    float s = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float *out; cudaMalloc(&out, 148 * 1024 * sizeof(float));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int kind = 0; kind < 2; ++kind) {
        for (int wps : {1, 2, 4}) {
            int threads = 128 * wps, blocks = 148;
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                if (kind == 0) k<0><<<blocks, threads>>>(out, iters); else k<1><<<blocks, threads>>>(out, iters);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
            }
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double mmas = (double)blocks * (threads / 32) * iters * 4.0;
            double macs = mmas * (kind == 0 ? 16 * 8 * 8 : 16 * 8 * 16);
            printf("%s warps/scheduler %d: %.3f ms, %.2f MMA/clk/SM, %.0f MAC/clk/SM, %.1f dense TFLOP/s (at 1.965 GHz)\n",
                   kind == 0 ? "m16n8k8 tf32 " : "m16n8k16 bf16", wps, ms, mmas / (ms * 1e-3) / 148 / 1.965e9,
                   macs / (ms * 1e-3) / 148 / 1.965e9, 2 * macs / ms / 1e9);
        }
    }
    return 0;
}
