// micro-benchmark: legacy mma.sync m16n8k16 (HMMA) issue rate on sm_100a, to size the network kernel's ceiling.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
template <int ILP>
__global__ void k(float *out, int iters)
{
    float c[ILP][4];
    uint32_t a[4] = {threadIdx.x, threadIdx.x * 3u, 7u, 9u}, b0 = threadIdx.x, b1 = 5u;
#pragma unroll
    for (int i = 0; i < ILP; i++) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main()
{
    float *out;
    cudaMalloc(&out, 148 * 1024 * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int warps : {4, 8, 16, 32}) {
        k<12><<<148, warps * 32>>>(out, 100);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        k<12><<<148, warps * 32>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        double mma = 148.0 * warps * 12.0 * iters;
        double tflops = mma * 2048 * 2 / (ms * 1e-3) / 1e12;
        printf("warps/SM %2d: %.3f ms, %.1f TFLOP/s dense f16 via mma.sync, %.2f MMA/us/SM\n", warps, ms, tflops,
               mma / 148 / (ms * 1e3));
    }
    return 0;
}
