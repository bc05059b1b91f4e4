// Microbenchmark: the B200's L2 -> SM peak for a RANDOM 256-byte ROW GATHER out of an L2-resident table -- the access
// pattern of the attentive SpMM (csrc/spmm.cu) at the Amazon-book shape (table 159,251 x 64 fp32 = 41 MB < 126 MB of L2).
// No arithmetic beyond what keeps the loads alive, no output besides one float per warp, every SM busy.  The best
// configuration found is the denominator of bench.py's `roofline_l2` block ("make the SpMM bound
// measurable").
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/microbench/l2_gather tools/microbench/l2_gather.cu
//   tools/microbench/l2_gather [--json] [rows=159251] [d=64] [gathers_per_row_of_index=40]
//
// Variants: lanes per row = d / 4 (one LDG.128 per lane: 16 lanes fetch a 256 B row, a warp instruction fetches two rows);
// U independent loads in flight per lane (4 .. 16); CTAs x threads chosen to sweep occupancy; read-only path with and
// without L1 allocation.  Indices are uniform random (no hot rows: what L1 can add on a skewed graph is not part of the
// L2 figure), read coalesced, 32 per warp step like the kernel's staging slab.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e = (x);                                                           \
        if (e != cudaSuccess) {                                                        \
            fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); \
            exit(1);                                                                   \
        }                                                                              \
    } while (0)

template <bool L1>
__device__ __forceinline__ float4 ld4(const float* p) {
    float4 r;
    if (L1)
        asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    else
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

// each warp walks its slice of the index list: 32 indices per step, EPW = 32 / LPE rows per load instruction, U in flight
template <int D, int U, bool L1>
__global__ void gather_kernel(const float* __restrict__ table, const int* __restrict__ idx, long long n_idx, float* __restrict__ sink) {
    constexpr int LPE = D / 4, EPW = 32 / LPE, EPI = EPW * U;
    const int lane = threadIdx.x & 31, sub = lane % LPE, slot = lane / LPE;
    const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long per = (n_idx / 32 + n_warps - 1) / n_warps * 32;
    const long long begin = warp * per, end = begin + per < n_idx ? begin + per : n_idx;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const char* base = reinterpret_cast<const char*>(table) + sub * 16;
    for (long long b = begin; b + 32 <= end; b += 32) {
        const unsigned off = (unsigned)idx[b + lane] * (unsigned)(D * 4);
#pragma unroll
        for (int j = 0; j < 32; j += EPI) {
            float4 x[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const unsigned o = __shfl_sync(0xffffffffu, off, j + u * EPW + slot);
                x[u] = ld4<L1>(reinterpret_cast<const float*>(base + o));
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                acc.x += x[u].x; acc.y += x[u].y; acc.z += x[u].z; acc.w += x[u].w;
            }
        }
    }
    if (acc.x + acc.y + acc.z + acc.w == 123.456f) sink[warp] = acc.x;  // keeps the loads alive, (almost) never stores
}

struct Result {
    int u, threads, ctas_per_sm;
    bool l1;
    double gbs;
};

template <int D, int U, bool L1>
double run(const float* table, const int* idx, long long n_idx, float* sink, int sms, int threads, int ctas_per_sm) {
    const int grid = sms * ctas_per_sm;
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    for (int w = 0; w < 2; ++w) gather_kernel<D, U, L1><<<grid, threads>>>(table, idx, n_idx, sink);
    CK(cudaEventRecord(a));
    const int reps = 5;
    for (int r = 0; r < reps; ++r) gather_kernel<D, U, L1><<<grid, threads>>>(table, idx, n_idx, sink);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, a, b));
    CK(cudaGetLastError());
    const long long per_warp = (n_idx / 32 + (long long)grid * threads / 32 - 1) / ((long long)grid * threads / 32) * 32;
    const double rows = (double)std::min<long long>(n_idx, per_warp * ((long long)grid * threads / 32)) / 32 * 32;
    return rows * D * 4.0 / (ms / reps * 1e-3) / 1e9;
}

int main(int argc, char** argv) {
    bool json = false;
    long long rows = 159251, per_row = 40;
    int d = 64;
    int pos = 0;
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "--json")) { json = true; continue; }
        if (pos == 0) rows = atoll(argv[i]);
        if (pos == 1) d = atoi(argv[i]);
        if (pos == 2) per_row = atoll(argv[i]);
        ++pos;
    }
    if (d != 64 && d != 32 && d != 128) { fprintf(stderr, "d must be 32, 64 or 128\n"); return 1; }
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    const long long n_idx = rows * per_row / 32 * 32;
    std::vector<int> h_idx(n_idx);
    std::mt19937_64 rng(2024);
    for (auto& v : h_idx) v = (int)(rng() % (unsigned long long)rows);
    float* table;
    int* idx;
    float* sink;
    CK(cudaMalloc(&table, rows * d * 4));
    CK(cudaMemset(table, 0, rows * d * 4));
    CK(cudaMalloc(&idx, n_idx * 4));
    CK(cudaMemcpy(idx, h_idx.data(), n_idx * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&sink, 1 << 22));
    std::vector<Result> res;
#define SWEEP(D_, U_)                                                                                         \
    for (int l1 = 0; l1 < 2; ++l1)                                                                             \
        for (int threads : {128, 256, 512})                                                                    \
            for (int c : {2, 4, 8, 16}) {                                                                      \
                if (threads * c > 2048) continue;                                                              \
                const double g = l1 ? run<D_, U_, true>(table, idx, n_idx, sink, sms, threads, c)              \
                                    : run<D_, U_, false>(table, idx, n_idx, sink, sms, threads, c);            \
                res.push_back({U_, threads, c, (bool)l1, g});                                                  \
            }
    if (d == 64) { SWEEP(64, 4) SWEEP(64, 8) SWEEP(64, 16) }
    if (d == 32) { SWEEP(32, 4) SWEEP(32, 8) }
    if (d == 128) { SWEEP(128, 2) SWEEP(128, 4) SWEEP(128, 8) }
    std::sort(res.begin(), res.end(), [](const Result& a, const Result& b) { return a.gbs > b.gbs; });
    if (json) {
        printf("{\"gather_peak_gbs\": %.1f, \"rows\": %lld, \"d\": %d, \"table_mb\": %.1f, \"gathers\": %lld, \"sms\": %d, "
               "\"best\": {\"loads_in_flight\": %d, \"threads\": %d, \"ctas_per_sm\": %d, \"l1_allocate\": %s}}\n",
               res[0].gbs, rows, d, rows * d * 4 / 1e6, n_idx, sms, res[0].u, res[0].threads, res[0].ctas_per_sm, res[0].l1 ? "true" : "false");
    } else {
        printf("L2 -> SM random %d-byte row gather, table %.1f MB (%lld rows), %lld gathers, %d SMs (%s)\n", d * 4, rows * d * 4 / 1e6, rows,
               n_idx, sms, prop.name);
        for (const auto& r : res)
            printf("  U=%2d threads=%3d ctas/sm=%2d l1=%d  %8.1f GB/s\n", r.u, r.threads, r.ctas_per_sm, (int)r.l1, r.gbs);
    }
    return 0;
}
