// Micro-benchmark for DESIGN.md section 9 item 2: can a cluster-wide hot-row cache in distributed shared memory take
// load off the L2 -> SM fabric that bounds the SpMM gather (12.5 TB/s at the C3 shape)?
//
// One warp gathers 256-byte rows (64 fp32) by index and accumulates them, like spmm_task_kernel.  A fraction `hot` of
// the indices points into a hot set that every 8-CTA cluster holds once in its distributed shared memory
// (rows striped over the 8 CTAs); the rest points into a 40 MB table that lives in L2.  Reported: gathered GB/s for
// hot = 0 (all L2), 0.2, 0.4, 0.6 and 1.0 (all DSMEM), with the hot rows read either through DSMEM or -- control --
// through global memory as well.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o dsmem_gather dsmem_gather.cu && ./dsmem_gather
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>

namespace cg = cooperative_groups;

constexpr int kD = 64;                 // floats per row
constexpr int kCluster = 8;
constexpr int kHotPerCta = 768;        // rows per CTA: 768 x 256 B = 192 KB
constexpr int kHot = kHotPerCta * kCluster;
constexpr int kThreads = 256;

template <bool USE_DSMEM>
__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(kThreads, 1)
gather_kernel(const float* __restrict__ table, const int32_t* __restrict__ idx, int64_t n_idx_per_warp, float* __restrict__ out) {
    extern __shared__ __align__(16) float hot[];  // [kHotPerCta][kD]: hot rows r with r % kCluster == cluster rank
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = cluster.block_rank();
    for (int i = threadIdx.x; i < kHotPerCta * (kD / 4); i += kThreads) {
        const int r = i / (kD / 4), q = i % (kD / 4);
        reinterpret_cast<float4*>(hot)[i] = __ldg(reinterpret_cast<const float4*>(table + (int64_t)(r * kCluster + rank) * kD) + q);
    }
    cluster.sync();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t gwarp = (int64_t)blockIdx.x * (kThreads / 32) + warp;
    const int32_t* my = idx + gwarp * n_idx_per_warp;
    const int half = lane >> 4, q = lane & 15;  // two rows per warp instruction, 16 lanes x 16 B per row
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t i = 0; i < n_idx_per_warp; i += 2) {
        const int32_t r = my[i + half];
        float4 v;
        if (USE_DSMEM && r < kHot) {
            const float* remote = cluster.map_shared_rank(hot, r % kCluster);
            v = *reinterpret_cast<const float4*>(remote + (r / kCluster) * kD + q * 4);
        } else {
            v = __ldg(reinterpret_cast<const float4*>(table + (int64_t)r * kD) + q);
        }
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    out[(int64_t)blockIdx.x * kThreads + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
    cluster.sync();  // nobody leaves while a peer may still read its shared memory
}

int main() {
    const int n_rows = 159251;
    const int n_ctas = 144;  // 18 clusters of 8 (148 SMs: 4 stay idle, as they would in the real kernel)
    const int64_t per_warp = 4096;
    const int64_t n_idx = (int64_t)n_ctas * (kThreads / 32) * per_warp;
    float* table; int32_t* idx; float* out;
    cudaMalloc(&table, (size_t)n_rows * kD * 4);
    cudaMalloc(&idx, n_idx * 4);
    cudaMalloc(&out, (size_t)n_ctas * kThreads * 4);
    cudaMemset(table, 0, (size_t)n_rows * kD * 4);
    const size_t smem = (size_t)kHotPerCta * kD * 4;
    cudaFuncSetAttribute(gather_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(gather_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    std::vector<int32_t> h(n_idx);
    const double fracs[] = {0.0, 0.2, 0.4, 0.6, 1.0};
    for (double f : fracs) {
        uint64_t s = 88172645463325252ull;
        auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
        for (int64_t i = 0; i < n_idx; ++i) {
            const bool is_hot = (rnd() % 1000) < (uint64_t)(f * 1000);
            h[i] = is_hot ? (int32_t)(rnd() % kHot) : (int32_t)(kHot + rnd() % (n_rows - kHot));
        }
        cudaMemcpy(idx, h.data(), n_idx * 4, cudaMemcpyHostToDevice);
        for (int mode = 0; mode < 2; ++mode) {
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0); cudaEventCreate(&e1);
            float best = 1e30f;
            for (int rep = 0; rep < 5; ++rep) {
                cudaEventRecord(e0);
                if (mode == 0) gather_kernel<true><<<n_ctas, kThreads, smem>>>(table, idx, per_warp, out);
                else gather_kernel<false><<<n_ctas, kThreads, smem>>>(table, idx, per_warp, out);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (ms < best) best = ms;
            }
            const cudaError_t err = cudaGetLastError();
            printf("hot fraction %.1f  hot rows via %-6s : %8.1f us  %7.1f GB/s gathered%s\n", f, mode == 0 ? "DSMEM" : "global", best * 1e3,
                   (double)n_idx * kD * 4 / (best * 1e-3) / 1e9, err == cudaSuccess ? "" : cudaGetErrorString(err));
        }
    }
    printf("(the hot-set fill -- %d rows per CTA -- is inside the timed region: %.1f MB per launch)\n", kHotPerCta,
           (double)n_ctas * kHotPerCta * kD * 4 / 1e6);
    return 0;
}
