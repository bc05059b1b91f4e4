// Row exchange between the ranks of a row-sharded propagation over NVLink peer memory.
//
// The reference is single-device; the sharded propagation needs, after every layer, the rows the other ranks
// have just produced.  Instead of an all-gather collective, every rank keeps its tables in one IPC-exportable
// allocation (`kgat_peer_alloc`), maps the allocations of its peers (`kgat_peer_import`) and
//   * WRITES the rows it produces straight into every peer's copy of the table -- either from the producing
//     kernel's own epilogue (bi-interaction forward / backward take a `peer_out` pointer list) or with
//     `kgat_peer_push` for rows produced elsewhere (the embedding rows after Adam);
//   * then raises a per-channel sequence number in every peer's flag pad and waits until every peer has raised
//     its own (`kgat_peer_signal_wait`): one launch, ~one NVLink round trip.
// Stores to peer memory issued by a kernel are complete when that kernel completes, and kernels of one stream
// run in order, so "producer kernel(s) ; signal_wait" on every rank means: after signal_wait returns, every
// peer's rows for this channel are in local HBM.  Everything is stream-ordered device work (no host
// synchronisation), so a whole sharded training step, exchanges included, is captured as one CUDA graph.
//
// A peer that never arrives must not wedge the GPU: the wait gives up after `timeout_cycles` and records the
// failure in `status` (checked by the host at the end of the epoch).
#include "common.cuh"

namespace kgat {
namespace {

__device__ __forceinline__ void st_release_sys(int32_t* p, int32_t v) {
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int32_t ld_acquire_sys(const int32_t* p) {
    int32_t v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// copy n_vec float4 from src to the same offset of every peer table
__global__ void __launch_bounds__(256) peer_push_kernel(const float4* __restrict__ src, float4* const* __restrict__ peer_dst, int n_peers,
                                                        int64_t n_vec) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_vec; i += stride) {
        const float4 v = __ldg(src + i);
        for (int q = 0; q < n_peers; ++q) peer_dst[q][i] = v;
    }
}

// the listed rows (node ids, device-side count) of a table -> the same rows of every peer's table; d / 4 lanes per row
__global__ void __launch_bounds__(256) peer_push_rows_kernel(const float* __restrict__ table, float* const* __restrict__ peer_tables, int n_peers,
                                                             const int32_t* __restrict__ rows, const int32_t* __restrict__ cnt_dev, int d4,
                                                             int64_t ld) {
    const int64_t total = (int64_t)cnt_dev[0] * d4;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t off = (int64_t)__ldg(rows + i / d4) * ld + (i % d4) * 4;
        const float4 v = __ldg(reinterpret_cast<const float4*>(table + off));
        for (int q = 0; q < n_peers; ++q) *reinterpret_cast<float4*>(peer_tables[q] + off) = v;
    }
}

// thread q: seq = ++my sequence number (thread 0 publishes it), raise it at peer q, wait for peer q's
__global__ void peer_signal_wait_kernel(int32_t* const* __restrict__ peer_flags /* my slot in peer q's pad */,
                                        const int32_t* __restrict__ my_flags /* slot q of my pad: written by peer q */, int n_peers,
                                        int32_t* __restrict__ seq, int32_t* __restrict__ status, long long timeout_cycles) {
    __shared__ int32_t s_seq;
    if (threadIdx.x == 0) s_seq = seq[0] + 1;
    __syncthreads();
    const int32_t want = s_seq;
    const int q = threadIdx.x;
    if (q < n_peers) {
        __threadfence_system();
        st_release_sys(peer_flags[q], want);
        const long long t0 = clock64();
        const bool failed_before = *reinterpret_cast<volatile int32_t*>(status) != 0;  // one time-out is enough: do not wait again
        while (!failed_before && ld_acquire_sys(my_flags + q) < want) {
            if (clock64() - t0 > timeout_cycles) {
                atomicExch(status, 1);
                break;
            }
            __nanosleep(64);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) seq[0] = want;
}

}  // namespace

int peer_push_launch(const float* src, float* const* peer_dst, int n_peers, int64_t n_floats, cudaStream_t stream, int max_ctas) {
    if (n_peers < 0 || n_floats < 0 || (n_floats & 3) || (reinterpret_cast<uintptr_t>(src) & 15)) return KGAT_ERR_INVALID_ARGUMENT;
    if (n_peers == 0 || n_floats == 0) return KGAT_OK;
    const int64_t n_vec = n_floats / 4;
    const int64_t want = (n_vec + 255) / 256;
    const int64_t cap = max_ctas > 0 ? max_ctas : (int64_t)sm_count() * 8;
    const int grid = (int)(want < cap ? want : cap);
    peer_push_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const float4*>(src), reinterpret_cast<float4* const*>(peer_dst), n_peers,
                                               n_vec);
    return check_launch();
}

}  // namespace kgat

using namespace kgat;

extern "C" {

int kgat_peer_alloc(int64_t bytes, void** ptr) {
    if (bytes <= 0 || !ptr) return KGAT_ERR_INVALID_ARGUMENT;
    KGAT_CUDA_TRY(cudaMalloc(ptr, (size_t)bytes));
    KGAT_CUDA_TRY(cudaMemset(*ptr, 0, (size_t)bytes));
    return KGAT_OK;
}

int kgat_peer_free(void* ptr) {
    KGAT_CUDA_TRY(cudaFree(ptr));
    return KGAT_OK;
}

int kgat_peer_export(const void* ptr, void* handle64) {
    static_assert(sizeof(cudaIpcMemHandle_t) == KGAT_PEER_HANDLE_BYTES, "handle size");
    if (!ptr || !handle64) return KGAT_ERR_INVALID_ARGUMENT;
    KGAT_CUDA_TRY(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64), const_cast<void*>(ptr)));
    return KGAT_OK;
}

int kgat_peer_import(const void* handle64, void** ptr) {
    if (!ptr || !handle64) return KGAT_ERR_INVALID_ARGUMENT;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    KGAT_CUDA_TRY(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return KGAT_OK;
}

int kgat_peer_close(void* ptr) {
    KGAT_CUDA_TRY(cudaIpcCloseMemHandle(ptr));
    return KGAT_OK;
}

int kgat_peer_push(const float* src, float* const* peer_dst, int32_t n_peers, int64_t n_floats, int32_t max_ctas, void* stream) {
    return peer_push_launch(src, peer_dst, n_peers, n_floats, (cudaStream_t)stream, max_ctas);
}

int kgat_peer_push_rows(const float* table, float* const* peer_tables, int32_t n_peers, const int32_t* rows, const int32_t* count_dev,
                        int64_t max_rows, int32_t d, int64_t ld, void* stream) {
    if (!table || !rows || !count_dev || n_peers < 0 || max_rows < 0 || d <= 0 || (d & 3) || (ld & 3) || (n_peers && !peer_tables))
        return KGAT_ERR_INVALID_ARGUMENT;
    if (n_peers == 0 || max_rows == 0) return KGAT_OK;
    const int64_t want = (max_rows * (d / 4) + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 8;
    peer_push_rows_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>(table, peer_tables, n_peers, rows, count_dev,
                                                                                                 d / 4, ld);
    return check_launch();
}

int kgat_peer_copy(void* dst, const void* src, int64_t bytes, void* stream) {
    if (bytes < 0 || (bytes && (!dst || !src))) return KGAT_ERR_INVALID_ARGUMENT;
    if (bytes == 0) return KGAT_OK;
    KGAT_CUDA_TRY(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDefault, (cudaStream_t)stream));
    return KGAT_OK;
}

int kgat_peer_signal_wait(int32_t* const* peer_flags, const int32_t* my_flags, int32_t n_peers, int32_t* seq, int32_t* status,
                          int64_t timeout_cycles, void* stream) {
    if (n_peers < 0 || n_peers > KGAT_MAX_PEERS || !seq || !status) return KGAT_ERR_INVALID_ARGUMENT;
    if (n_peers == 0) return KGAT_OK;
    peer_signal_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(peer_flags, my_flags, n_peers, seq, status, (long long)timeout_cycles);
    return check_launch();
}

}  // extern "C"
