// Graph containers: key grouping (COO coalescing), CSR/CSC decoding, deterministic duplicate merge.
// Replaces torch.sparse's implicit coalesce (reference model.py:359-364) and scipy's coo->csr
// (reference preprocess.py:629).  One-off work per graph / per edge list: CUB does the radix sort
// and the scan, the glue kernels are ours.
#include <cub/cub.cuh>

#include <string>

#include "common.cuh"

namespace kgat {

static thread_local std::string g_last_error;
void set_cuda_error(cudaError_t e) { g_last_error = cudaGetErrorString(e); }

namespace {

__global__ void iota_kernel(int32_t* p, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) p[i] = (int32_t)i;
}

__global__ void head_flag_kernel(const uint64_t* __restrict__ keys, int64_t n, int32_t* __restrict__ flag) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) flag[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

__global__ void scatter_groups_kernel(const uint64_t* __restrict__ keys, const int32_t* __restrict__ order,
                                      const int32_t* __restrict__ rank_incl, int64_t n, int32_t* __restrict__ group_of,
                                      int32_t* __restrict__ group_ptr, uint64_t* __restrict__ unique_keys) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    int32_t g = rank_incl[i] - 1;
    group_of[order[i]] = g;
    if (i == 0 || keys[i] != keys[i - 1]) {
        group_ptr[g] = (int32_t)i;
        unique_keys[g] = keys[i];
    }
    if (i == n - 1) group_ptr[g + 1] = (int32_t)n;
}

__global__ void decode_keys_kernel(const uint64_t* __restrict__ keys, int64_t n_keys, int64_t n_major, uint64_t n_minor,
                                   int32_t* __restrict__ major_ptr, int32_t* __restrict__ minor_idx) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n_keys) minor_idx[i] = (int32_t)(keys[i] % n_minor);
    if (i <= n_major) {
        // lower_bound(keys, i * n_minor)
        uint64_t target = (uint64_t)i * n_minor;
        int64_t lo = 0, hi = n_keys;
        while (lo < hi) {
            int64_t mid = (lo + hi) >> 1;
            if (keys[mid] < target) lo = mid + 1; else hi = mid;
        }
        major_ptr[i] = (int32_t)lo;
    }
}

__global__ void segment_sum_kernel(const float* __restrict__ in, const int32_t* __restrict__ order,
                                   const int32_t* __restrict__ group_ptr, int64_t n_groups, float* __restrict__ out) {
    int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    int32_t b = group_ptr[g], e = group_ptr[g + 1];
    float s = in[order[b]];
    for (int32_t p = b + 1; p < e; ++p) s += in[order[p]];
    out[g] = s;
}

__global__ void gather_kernel(const float* __restrict__ in, const int32_t* __restrict__ index, int64_t n, float* __restrict__ out) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[index[i]];
}

__global__ void ids_to_i32_kernel(const int64_t* __restrict__ in, int64_t n, int64_t bound, int32_t* __restrict__ out,
                                  int32_t* __restrict__ bad) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t v = in[i];
    if (v < 0 || v >= bound) {
        atomicAdd(bad, 1);
        v = 0;
    }
    out[i] = (int32_t)v;
}

__global__ void fill_kernel(float* p, int64_t n, float v) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = v;
}

// dst[0 .. elems) = src[(counter % n_batches) * elems ...]: picks the current step's pre-sampled batch
// from a device-resident epoch array inside a captured CUDA graph (the counter is the optimiser's
// device step counter, so a replayed graph walks through the epoch with no host work).
__global__ void select_batch_kernel(const int64_t* __restrict__ src, int64_t n_batches, int64_t elems,
                                    const int64_t* __restrict__ counter, int64_t* __restrict__ dst) {
    const int64_t b = counter[0] % n_batches;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < elems; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = src[b * elems + i];
}

// select_batch + kgat_adam_advance in one single-CTA launch (the first node of a captured KG step): every thread reads the
// counter, the batch is copied, and only then is the counter advanced and the step's Adam scalars written
__global__ void __launch_bounds__(1024) step_begin_kernel(const int64_t* __restrict__ src, int64_t n_batches, int64_t elems,
                                                          int64_t* __restrict__ step, int64_t* __restrict__ dst, double lr, double b1,
                                                          double b2, double eps, float* __restrict__ hyper) {
    const int64_t s_old = step[0];
    const int64_t b = s_old % n_batches;
    for (int64_t i = threadIdx.x; i < elems; i += blockDim.x) dst[i] = src[b * elems + i];
    __syncthreads();
    if (threadIdx.x == 0) {
        const int64_t s = s_old + 1;
        step[0] = s;
        const double bc1 = 1.0 - pow(b1, (double)s);
        const double bc2 = 1.0 - pow(b2, (double)s);
        hyper[0] = (float)(1.0 - b1);
        hyper[1] = (float)b2;
        hyper[2] = (float)(1.0 - b2);
        hyper[3] = (float)(lr / bc1);
        hyper[4] = (float)(1.0 / sqrt(bc2));
        hyper[5] = (float)eps;
    }
}

inline int64_t align256(int64_t x) { return (x + 255) & ~(int64_t)255; }
inline unsigned blocks_for(int64_t n, int threads = 256) { return (unsigned)((n + threads - 1) / threads); }

size_t cub_temp_bytes(int64_t n) {
    size_t sort_bytes = 0, scan_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const int32_t*)nullptr,
                                    (int32_t*)nullptr, (int)n, 0, 64, (cudaStream_t)0);
    cub::DeviceScan::InclusiveSum(nullptr, scan_bytes, (const int32_t*)nullptr, (int32_t*)nullptr, (int)n, (cudaStream_t)0);
    return sort_bytes > scan_bytes ? sort_bytes : scan_bytes;
}

}  // namespace
}  // namespace kgat

using namespace kgat;

extern "C" {

int kgat_abi_version(void) { return KGAT_ABI_VERSION; }

const char* kgat_error_string(int code) {
    switch (code) {
        case KGAT_OK: return "ok";
        case KGAT_ERR_INVALID_ARGUMENT: return "invalid argument";
        case KGAT_ERR_CUDA: return "CUDA runtime error (see kgat_last_cuda_error)";
        case KGAT_ERR_UNSUPPORTED: return "unsupported shape or configuration";
        case KGAT_ERR_WORKSPACE: return "workspace too small";
        default: return "unknown error";
    }
}

const char* kgat_last_cuda_error(void) { return g_last_error.c_str(); }

int kgat_device_info(int* sm, int* cc_major, int* cc_minor, int64_t* l2_bytes) {
    int dev = 0;
    KGAT_CUDA_TRY(cudaGetDevice(&dev));
    cudaDeviceProp p;
    KGAT_CUDA_TRY(cudaGetDeviceProperties(&p, dev));
    if (sm) *sm = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (l2_bytes) *l2_bytes = p.l2CacheSize;
    return KGAT_OK;
}

int64_t kgat_group_by_key_workspace_bytes(int64_t n) {
    if (n < 0 || n > 0x7fffffff) return -1;
    int64_t m = n > 0 ? n : 1;
    return align256(m * 8) + 2 * align256(m * 4) + align256((int64_t)cub_temp_bytes(m)) + 256;
}

int kgat_group_by_key(const uint64_t* keys, int64_t n, int key_bits, void* workspace, int64_t workspace_bytes,
                      int32_t* order, int32_t* group_of, int32_t* group_ptr, uint64_t* unique_keys, int64_t* n_groups_host,
                      void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n < 0 || n > 0x7fffffff || !n_groups_host || key_bits <= 0 || key_bits > 64) return KGAT_ERR_INVALID_ARGUMENT;
    if (n == 0) {
        *n_groups_host = 0;
        int32_t zero = 0;
        KGAT_CUDA_TRY(cudaMemcpyAsync(group_ptr, &zero, sizeof(zero), cudaMemcpyHostToDevice, stream));
        KGAT_CUDA_TRY(cudaStreamSynchronize(stream));
        return KGAT_OK;
    }
    if (workspace_bytes < kgat_group_by_key_workspace_bytes(n)) return KGAT_ERR_WORKSPACE;
    char* ws = (char*)workspace;
    uint64_t* keys_sorted = (uint64_t*)ws;
    ws += align256(n * 8);
    int32_t* iota = (int32_t*)ws;
    ws += align256(n * 4);
    int32_t* rank = (int32_t*)ws;
    ws += align256(n * 4);
    size_t temp_bytes = cub_temp_bytes(n);
    void* temp = ws;

    iota_kernel<<<blocks_for(n), 256, 0, stream>>>(iota, n);
    KGAT_CUDA_TRY(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys, keys_sorted, iota, order, (int)n, 0, key_bits, stream));
    int32_t* flag = iota;  // iota is dead after the sort
    head_flag_kernel<<<blocks_for(n), 256, 0, stream>>>(keys_sorted, n, flag);
    KGAT_CUDA_TRY(cub::DeviceScan::InclusiveSum(temp, temp_bytes, flag, rank, (int)n, stream));
    scatter_groups_kernel<<<blocks_for(n), 256, 0, stream>>>(keys_sorted, order, rank, n, group_of, group_ptr, unique_keys);
    int32_t last = 0;
    KGAT_CUDA_TRY(cudaMemcpyAsync(&last, rank + (n - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    KGAT_CUDA_TRY(cudaStreamSynchronize(stream));
    *n_groups_host = last;
    return check_launch();
}

int kgat_decode_sorted_keys(const uint64_t* unique_keys, int64_t n_keys, int64_t n_major, int64_t n_minor, int32_t* major_ptr,
                            int32_t* minor_idx, void* stream) {
    if (n_keys < 0 || n_major < 0 || n_minor <= 0) return KGAT_ERR_INVALID_ARGUMENT;
    int64_t work = n_keys > n_major + 1 ? n_keys : n_major + 1;
    decode_keys_kernel<<<blocks_for(work), 256, 0, (cudaStream_t)stream>>>(unique_keys, n_keys, n_major, (uint64_t)n_minor,
                                                                          major_ptr, minor_idx);
    return check_launch();
}

int kgat_segment_sum_f32(const float* in, const int32_t* order, const int32_t* group_ptr, int64_t n_groups, float* out,
                         void* stream) {
    if (n_groups < 0) return KGAT_ERR_INVALID_ARGUMENT;
    if (n_groups == 0) return KGAT_OK;
    segment_sum_kernel<<<blocks_for(n_groups), 256, 0, (cudaStream_t)stream>>>(in, order, group_ptr, n_groups, out);
    return check_launch();
}

int kgat_gather_f32(const float* in, const int32_t* index, int64_t n, float* out, void* stream) {
    if (n < 0) return KGAT_ERR_INVALID_ARGUMENT;
    if (n == 0) return KGAT_OK;
    gather_kernel<<<blocks_for(n), 256, 0, (cudaStream_t)stream>>>(in, index, n, out);
    return check_launch();
}

int kgat_ids64_to_i32(const int64_t* in, int64_t n, int64_t bound, int32_t* out, int32_t* bad_count_dev, void* stream) {
    if (n < 0 || bound <= 0 || bound > 0x7fffffff) return KGAT_ERR_INVALID_ARGUMENT;
    if (n == 0) return KGAT_OK;
    ids_to_i32_kernel<<<blocks_for(n), 256, 0, (cudaStream_t)stream>>>(in, n, bound, out, bad_count_dev);
    return check_launch();
}

int kgat_select_batch_i64(const int64_t* src, int64_t n_batches, int64_t elems, const int64_t* counter_dev, int64_t* dst,
                          void* stream) {
    if (n_batches <= 0 || elems <= 0 || !counter_dev) return KGAT_ERR_INVALID_ARGUMENT;
    select_batch_kernel<<<(unsigned)((elems + 255) / 256 < 64 ? (elems + 255) / 256 : 64), 256, 0, (cudaStream_t)stream>>>(
        src, n_batches, elems, counter_dev, dst);
    return check_launch();
}

int kgat_step_begin_i64(const int64_t* src, int64_t n_batches, int64_t elems, int64_t* step_dev, int64_t* dst, double lr, double beta1,
                        double beta2, double eps, float* hyper_dev, void* stream) {
    if (!src || n_batches <= 0 || elems <= 0 || !step_dev || !dst || !hyper_dev) return KGAT_ERR_INVALID_ARGUMENT;
    step_begin_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(src, n_batches, elems, step_dev, dst, lr, beta1, beta2, eps, hyper_dev);
    return check_launch();
}

int kgat_fill_f32(float* p, int64_t n, float value, void* stream) {
    if (n < 0) return KGAT_ERR_INVALID_ARGUMENT;
    if (n == 0) return KGAT_OK;
    int64_t blocks = (n + 255) / 256;
    int64_t cap = (int64_t)sm_count() * 16;
    fill_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(p, n, value);
    return check_launch();
}

}  // extern "C"
