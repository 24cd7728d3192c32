// FP32 FMA issue rate on sm_100a: scalar FFMA against the packed FFMA2 (fma.rn.f32x2), 8 independent chains per thread.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes/ffma2_probe tools/probes/ffma2_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}
template <bool PACKED>
__global__ void k(float *out, int iters, float m, float a) {
    float2 acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = make_float2(threadIdx.x + j, threadIdx.x - j);
    float2 mm = make_float2(m, m * 1.0001f), aa = make_float2(a, a * 0.999f);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (PACKED) acc[j] = ffma2(acc[j], mm, aa);
            else { acc[j].x = fmaf(acc[j].x, mm.x, aa.x); acc[j].y = fmaf(acc[j].y, mm.y, aa.y); }
        }
    }
    float s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += acc[j].x + acc[j].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float *out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int packed = 0; packed < 2; ++packed) {
        for (int wps : {1, 2, 4, 8}) {       // warps per scheduler: 4 schedulers per SM
            int threads = 128 * wps > 1024 ? 1024 : 128 * wps, blocks = 148 * (128 * wps / threads);
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                if (packed) k<true><<<blocks, threads>>>(out, iters, 0.9999f, 0.5f); else k<false><<<blocks, threads>>>(out, iters, 0.9999f, 0.5f);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
            }
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double fma = (double)blocks * threads * iters * 16.0;
            printf("%s warps/scheduler %d: %.3f ms, %.2f TFLOP/s, %.1f FMA/clk/SM at 1.965 GHz\n", packed ? "FFMA2" : "FFMA ", wps, ms,
                   2 * fma / ms / 1e9, fma / (ms * 1e-3) / 148 / 1.965e9);
        }
    }
    return 0;
}
