// Host-latency helpers of the reference-facing API fast path (functions.py: GraphedStep / LazyLoss).
//
// The reference driver runs loss = model(...); loss.backward(); model.update_*_weights(); loss.item() per step
// (main.py:306-314, 334-343).  A KG step is ~60 us of GPU work, so the API path is bound by host time per call:
//   * kgat_step_submit     one C call = host->device (or device->device) copy of the step's ids + cudaGraphLaunch of the
//                          step's captured graph (no Python-side replay wrapper, no second library call);
//   * kgat_publish_loss    a graph node that writes (serial, loss) as ONE 8-byte word into a ring in mapped pinned host
//                          memory right after the loss kernel, so loss.item() is a host-side poll of that word and does
//                          not wait for the backward / Adam kernels queued behind it.
#include "common.cuh"

namespace kgat {
namespace {

// ring[serial % n_slots] = (serial << 32) | bits(loss); serial is the device-side count of published losses (1-based)
__global__ void publish_loss_kernel(const float* __restrict__ loss, unsigned long long* __restrict__ serial_dev,
                                    volatile unsigned long long* __restrict__ ring_host, int n_slots) {
    const unsigned long long s = serial_dev[0] + 1ull;
    serial_dev[0] = s;
    const unsigned long long word = (s << 32) | (unsigned long long)__float_as_uint(loss[0]);
    ring_host[s % (unsigned long long)n_slots] = word;  // one aligned 8-byte store: the host sees both halves or neither
    __threadfence_system();
}

}  // namespace
}  // namespace kgat

using namespace kgat;

extern "C" {

int kgat_publish_loss(const float* loss, uint64_t* serial_dev, uint64_t* ring_host_mapped, int32_t n_slots, void* stream) {
    if (!loss || !serial_dev || !ring_host_mapped || n_slots <= 0) return KGAT_ERR_INVALID_ARGUMENT;
    publish_loss_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(loss, reinterpret_cast<unsigned long long*>(serial_dev),
                                                           reinterpret_cast<volatile unsigned long long*>(ring_host_mapped), n_slots);
    return check_launch();
}

int kgat_graph_launch(void* graph_exec, void* stream) {
    if (!graph_exec) return KGAT_ERR_INVALID_ARGUMENT;
    KGAT_CUDA_TRY(cudaGraphLaunch((cudaGraphExec_t)graph_exec, (cudaStream_t)stream));
    return KGAT_OK;
}

int kgat_step_submit(void* dst_dev, const void* src, int64_t n_bytes, void* graph_exec, void* stream) {
    if (!graph_exec || n_bytes < 0 || (n_bytes > 0 && (!dst_dev || !src))) return KGAT_ERR_INVALID_ARGUMENT;
    if (n_bytes > 0) KGAT_CUDA_TRY(cudaMemcpyAsync(dst_dev, src, (size_t)n_bytes, cudaMemcpyDefault, (cudaStream_t)stream));
    KGAT_CUDA_TRY(cudaGraphLaunch((cudaGraphExec_t)graph_exec, (cudaStream_t)stream));
    return KGAT_OK;
}

}  // extern "C"
