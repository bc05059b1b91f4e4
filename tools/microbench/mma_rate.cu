// Micro-benchmark: issue rate of legacy warp-level mma.sync on sm_100a (TF32 m16n8k8, BF16 m16n8k16)
// and of plain FFMA, to decide which pipe the small bi-interaction GEMMs should use.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

template <int MODE>
__global__ void k(float* out, int iters) {
    float d[8][4];
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) d[i][j] = threadIdx.x * 1e-3f + i + j;
    uint32_t a[4] = {threadIdx.x, threadIdx.x + 1u, threadIdx.x + 2u, threadIdx.x + 3u}, b[2] = {threadIdx.x * 3u, threadIdx.x * 5u};
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
            else if (MODE == 1)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
            else {
#pragma unroll
                for (int j = 0; j < 4; ++j) d[i][j] = fmaf(d[i][j], 1.0001f, 0.5f);
            }
        }
    }
    float s = 0;
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += d[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, double macs_per_inst) {
    float* out;
    cudaMalloc(&out, 148 * 8 * 256 * 4);
    const int iters = 20000;
    k<MODE><<<148 * 4, 256>>>(out, 10);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<148 * 4, 256>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double warps = 148.0 * 4 * 8, inst = warps * iters * 8.0;
    double per_sm_clk = inst * macs_per_inst / (ms * 1e-3) / 148 / 1.9e9;
    printf("%-10s %8.3f ms  %8.2f T-inst/s  ~%7.1f MAC/clk/SM (at 1.9 GHz)  %8.1f TFLOP/s\n", name, ms, inst / ms / 1e9, per_sm_clk, 2 * inst * macs_per_inst / ms / 1e9);
    cudaFree(out);
}

int main() {
    run<0>("tf32 k8", 16 * 8 * 8);
    run<1>("bf16 k16", 16 * 8 * 16);
    run<2>("ffma x4", 32 * 4);
    return 0;
}
