// K1: attentive SpMM  Y = A * X (+ Z) over the CSR attentive matrix.
// Replaces torch.matmul(sparse_coo, dense) (reference aggregator.py:54) and, run on the transposed
// container, its autograd backward A^T * g.
//
// HBM/L2-bound gather: one warp per task (a whole short row, or a <=chunk slice of a long row).
// A feature row of D floats is fetched by D/4 lanes with one 128-bit read-only load each, so a warp
// covers 32/(D/4) edges per step and keeps several steps in flight (4 independent 128-bit loads per
// lane).  (col, val) of 32 consecutive edges are read coalesced once and broadcast by shuffle.
// Long rows are split by the host-side plan into chunks whose partial sums are reduced in a fixed
// order by a second tiny kernel: deterministic, no atomics.
#include <stdlib.h>

#include "common.cuh"

namespace kgat {
namespace {

__device__ __forceinline__ int ld_stream_i32(const int32_t* p) {
    int r;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ float ld_stream_f32(const float* p) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}

// A long row is split into chunks (one task each) whose partial sums land in `partials`.  The warp that
// finishes the LAST chunk of a row (device-wide arrival counter in heavy[h].w) adds the partials up in chunk
// order -- deterministic, no float atomics -- and resets the counter for the next launch.
template <int D>
__device__ __forceinline__ void finish_heavy_row(int slot, float* __restrict__ partials, int4* __restrict__ heavy, int n_heavy,
                                                 float* __restrict__ Y, int64_t ldy, const float* __restrict__ Z, int64_t ldz, int lane) {
    int lo = 0, hi = n_heavy - 1;  // heavy rows are sorted by first_slot: find the one owning `slot`
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (heavy[mid].y <= slot) lo = mid; else hi = mid - 1;
    }
    const int row = heavy[lo].x, first = heavy[lo].y, n_chunks = heavy[lo].z;
    __threadfence();  // this warp's partial is visible device-wide before it is counted
    int prev = 0;
    if (lane == 0) prev = atomicAdd(&heavy[lo].w, 1);
    prev = __shfl_sync(kFull, prev, 0);
    if (prev != n_chunks - 1) return;
    __threadfence();
    for (int f = lane * 4; f < D; f += 128) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int c = 0; c < n_chunks; ++c) {
            const float4 p = __ldcg(reinterpret_cast<const float4*>(partials + (int64_t)(first + c) * D + f));
            a.x += p.x; a.y += p.y; a.z += p.z; a.w += p.w;
        }
        if (Z != nullptr) {
            const float4 z = ld_stream4(Z + (int64_t)row * ldz + f);
            a.x += z.x; a.y += z.y; a.z += z.z; a.w += z.w;
        }
        *reinterpret_cast<float4*>(Y + (int64_t)row * ldy + f) = a;
    }
    if (lane == 0) heavy[lo].w = 0;
}

// One warp per task.  The (byte offset, weight) pairs of 32 consecutive edges are staged in a
// per-warp slab of shared memory (one coalesced global read + one STS.64 per lane), so the hot loop
// per edge group is: LDS.64 (broadcast), LDG.128 (gather of the neighbour row), 4 FFMA -- no
// shuffles, no predicates (tail lanes carry weight 0 and a valid, already-cached address).
// WIDE = false: the table is < 4 GiB so a 32-bit byte offset addresses it.
// MASKED (needed-row pruning of a CF step, frontier.cu): `edge_mask` (bitmap over columns, NULL = all; may point to a
// shared-memory copy) drops the edges whose source row holds no valid data -- the surviving edges are compacted in the
// staging slab, so the hot loop is unchanged -- and gates the addend Z the same way (Z[row] counts only when row is in
// `edge_mask`).  Row masking (skipping the tasks of rows nobody reads) is done by the callers.
template <int D, int U, bool WIDE, bool MASKED>
__device__ __forceinline__ void spmm_do_task(const int4 t, int2* __restrict__ eb, const int32_t* __restrict__ col_idx,
                                             const float* __restrict__ vals, const float* __restrict__ X, int64_t ldx,
                                             float* __restrict__ Y, int64_t ldy, const float* __restrict__ Z, int64_t ldz,
                                             float* __restrict__ partials, int4* __restrict__ heavy, int n_heavy,
                                             const uint32_t* __restrict__ edge_mask, const uint32_t* __restrict__ edge_mask_global) {
    constexpr int LPE = D / 4;        // lanes per edge
    constexpr int EPW = 32 / LPE;     // edges per warp step
    constexpr int EPI = EPW * U;      // edges per unrolled iteration (divides 32)
    static_assert(32 % EPI == 0, "unroll must divide the 32-edge batch");
    const int lane = threadIdx.x & 31;
    if (MASKED && Z != nullptr && edge_mask_global != nullptr && !((__ldg(edge_mask_global + (t.x >> 5)) >> (t.x & 31)) & 1u)) Z = nullptr;
    const int sub = lane % LPE;
    const int slot = lane / LPE;
    const char* xb = reinterpret_cast<const char*>(X) + sub * 16;
    const uint32_t row_bytes32 = (uint32_t)(ldx * 4);
    const int64_t row_bytes64 = ldx * 4;

    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int base = t.y;
    int c = 0;
    float v = 0.f;
    if (base < t.z) {
        const int k = base + lane;
        c = ld_stream_i32(col_idx + min(k, t.z - 1));
        v = k < t.z ? ld_stream_f32(vals + k) : 0.f;
    }
    while (base < t.z) {
        int cnt;
        if (MASKED && edge_mask != nullptr) {
            // keep the edges whose source row is live, packed to the front of the slab; pad to a whole unrolled step
            // (padding entries carry weight 0 and the address of a LIVE row: dead rows may hold NaN garbage)
            const bool live = (base + lane < t.z) && ((edge_mask[c >> 5] >> (c & 31)) & 1u);
            const unsigned m = __ballot_sync(kFull, live);
            cnt = __popc(m);
            const int off = WIDE ? c : (int)((uint32_t)c * row_bytes32);
            if (live) eb[__popc(m & ((1u << lane) - 1u))] = make_int2(off, __float_as_int(v));
            if (cnt > 0) {
                const int live_off = __shfl_sync(kFull, off, __ffs(m) - 1);
                const int pad = (EPI - (cnt % EPI)) % EPI;
                if (lane < pad) eb[cnt + lane] = make_int2(live_off, 0);
            }
        } else {
            eb[lane] = make_int2(WIDE ? c : (int)((uint32_t)c * row_bytes32), __float_as_int(v));
            cnt = min(32, t.z - base);
        }
        __syncwarp();
        base += 32;
        if (base < t.z) {  // prefetch the next 32 (col, val) pairs while this batch is gathered
            const int k = base + lane;
            c = ld_stream_i32(col_idx + min(k, t.z - 1));
            v = k < t.z ? ld_stream_f32(vals + k) : 0.f;
        }
        for (int j = 0; j < cnt; j += EPI) {
            float4 x[U];
            float w[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int2 ev = eb[j + u * EPW + slot];
                w[u] = __int_as_float(ev.y);
                const char* src = WIDE ? xb + (int64_t)ev.x * row_bytes64 : xb + (uint32_t)ev.x;
                x[u] = __ldg(reinterpret_cast<const float4*>(src));
            }
#pragma unroll
            for (int u = 0; u < U; ++u) fma4(acc, w[u], x[u]);
        }
        __syncwarp();
    }
#pragma unroll
    for (int o = LPE; o < 32; o <<= 1) {
        acc.x += __shfl_xor_sync(kFull, acc.x, o);
        acc.y += __shfl_xor_sync(kFull, acc.y, o);
        acc.z += __shfl_xor_sync(kFull, acc.z, o);
        acc.w += __shfl_xor_sync(kFull, acc.w, o);
    }
    if (slot == 0) {
        if (t.w < 0) {
            if (Z != nullptr) {
                const float4 z = ld_stream4(Z + (int64_t)t.x * ldz + sub * 4);
                acc.x += z.x; acc.y += z.y; acc.z += z.z; acc.w += z.w;
            }
            *reinterpret_cast<float4*>(Y + (int64_t)t.x * ldy + sub * 4) = acc;
        } else {
            *reinterpret_cast<float4*>(partials + (int64_t)t.w * D + sub * 4) = acc;
        }
    }
    if (t.w >= 0) finish_heavy_row<D>(t.w, partials, heavy, n_heavy, Y, ldy, Z, ldz, lane);
}

// grid-per-task launch: one warp per task of the plan (`row_mask`: bitmap over output rows, NULL = all; the warps of
// other rows exit at once and their Y rows stay untouched)
template <int D, int U, bool WIDE, bool MASKED>
__global__ void __launch_bounds__(128) spmm_task_kernel(const int4* __restrict__ tasks, int64_t n_tasks,
                                                        const int32_t* __restrict__ col_idx, const float* __restrict__ vals,
                                                        const float* __restrict__ X, int64_t ldx, float* __restrict__ Y,
                                                        int64_t ldy, const float* __restrict__ Z, int64_t ldz,
                                                        float* __restrict__ partials, int4* __restrict__ heavy, int n_heavy,
                                                        const uint32_t* __restrict__ row_mask, const uint32_t* __restrict__ edge_mask) {
    __shared__ int2 ebuf[4][32];
    const int64_t task_id = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (task_id >= n_tasks) return;
    const int4 t = __ldg(tasks + task_id);  // {row, begin, end, partial_slot}
    if (MASKED && row_mask != nullptr && !((__ldg(row_mask + (t.x >> 5)) >> (t.x & 31)) & 1u)) return;
    spmm_do_task<D, U, WIDE, MASKED>(t, ebuf[threadIdx.x >> 5], col_idx, vals, X, ldx, Y, ldy, Z, ldz, partials, heavy, n_heavy, edge_mask,
                                     edge_mask);
}

// Persistent launch over a needed-row list (frontier.cu): the warps of a fixed grid stride over
//   [0, n_heavy_tasks)            the chunk tasks of the heavy rows (first in the plan), filtered by `row_mask`, then
//   rows[0 .. *n_rows_dev)        the listed rows: a light row owns exactly one task, n_heavy_tasks + light_rank[row]
// (rows == NULL: every task of the plan -- the dense-output backward, which only masks edges).  No warp is launched for
// a row outside the frontier -- the grid-per-task kernel spends ~30 us on 40 k empty CTAs when 3 % of the rows are live --
// and the edge bitmap is staged in shared memory once per CTA when it fits, so the per-edge test is an LDS.
constexpr int kRowsThreads = 256;
template <int D, int U, bool WIDE, bool MASKED, int CTAS>
__global__ void __launch_bounds__(kRowsThreads, CTAS) spmm_rows_kernel(const int4* __restrict__ tasks, int n_tasks, int n_heavy_tasks,
                                                                const int32_t* __restrict__ light_rank, const int32_t* __restrict__ rows,
                                                                const int32_t* __restrict__ n_rows_dev, const int32_t* __restrict__ col_idx,
                                                                const float* __restrict__ vals, const float* __restrict__ X, int64_t ldx,
                                                                float* __restrict__ Y, int64_t ldy, const float* __restrict__ Z, int64_t ldz,
                                                                float* __restrict__ partials, int4* __restrict__ heavy, int n_heavy,
                                                                const uint32_t* __restrict__ row_mask, const uint32_t* __restrict__ edge_mask,
                                                                int mask_words_smem) {
    extern __shared__ __align__(16) uint32_t smem_mask[];
    __shared__ int2 ebuf[kRowsThreads / 32][32];
    const uint32_t* emask = edge_mask;
    if (MASKED && mask_words_smem > 0) {
        for (int i = threadIdx.x; i < mask_words_smem; i += kRowsThreads) smem_mask[i] = __ldg(edge_mask + i);
        __syncthreads();
        emask = smem_mask;
    }
    const int n_warps = (gridDim.x * kRowsThreads) >> 5;
    const int total = rows != nullptr ? n_heavy_tasks + n_rows_dev[0] : n_tasks;
    for (int i = (blockIdx.x * kRowsThreads + threadIdx.x) >> 5; i < total; i += n_warps) {
        int4 t;
        if (rows == nullptr) {
            t = __ldg(tasks + i);
        } else if (i < n_heavy_tasks) {
            t = __ldg(tasks + i);
            if (row_mask != nullptr && !((__ldg(row_mask + (t.x >> 5)) >> (t.x & 31)) & 1u)) continue;
        } else {
            const int lr = __ldg(light_rank + __ldg(rows + (i - n_heavy_tasks)));
            if (lr < 0) continue;  // a heavy row: its chunks were taken above
            t = __ldg(tasks + n_heavy_tasks + lr);
        }
        spmm_do_task<D, U, WIDE, MASKED>(t, ebuf[threadIdx.x >> 5], col_idx, vals, X, ldx, Y, ldy, Z, ldz, partials, heavy, n_heavy,
                                         MASKED ? emask : nullptr, MASKED ? edge_mask : nullptr);
    }
}

// Transposed product as a SCATTER over the rows of a (small) needed-row list:  Y[c] += A[r, c] * G[r] for every listed row r
// and every edge (r, c), plus Y[r] += Z[r].  Used for the backward of the sparse upper layers of a pruned CF step, where the
// gradient sources are a few thousand rows: the gather formulation (spmm_rows_kernel over A^T) has to stream and test every
// edge of every destination row (3.7 M edge tests for 0.7 M live edges at the Amazon-book shape, 76 us), the scatter touches
// the live edges only.  128-bit vector reductions (red.global.add.v4.f32); the destination rows are zeroed beforehand
// (kgat_frontier_zero_rows).  Summation order is not fixed (fp32 atomics), like the BPR / TransR gradient scatters.
__device__ __forceinline__ void red_add4(float* p, float a, float b, float c, float d) {
    asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int D>
__global__ void __launch_bounds__(256) spmm_scatter_rows_kernel(const int4* __restrict__ tasks, int n_heavy_tasks,
                                                               const int32_t* __restrict__ light_rank, const int32_t* __restrict__ rows,
                                                               const int32_t* __restrict__ n_rows_dev, const uint32_t* __restrict__ row_mask,
                                                               const int32_t* __restrict__ col_idx, const float* __restrict__ vals,
                                                               const float* __restrict__ G, int64_t ldg, const float* __restrict__ Z,
                                                               int64_t ldz, float* __restrict__ Y, int64_t ldy) {
    constexpr int LPE = D / 4;     // lanes per edge
    constexpr int EPW = 32 / LPE;  // edges per warp step
    const int lane = threadIdx.x & 31;
    const int sub = lane % LPE, slot = lane / LPE;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    const int total = n_heavy_tasks + n_rows_dev[0];
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < total; i += n_warps) {
        int4 t;
        bool first_chunk = true;  // the task that also carries the row's direct term Z[r]
        if (i < n_heavy_tasks) {
            t = __ldg(tasks + i);
            if (!((__ldg(row_mask + (t.x >> 5)) >> (t.x & 31)) & 1u)) continue;
            first_chunk = (i == 0) || (__ldg(tasks + i - 1).x != t.x);
        } else {
            const int lr = __ldg(light_rank + __ldg(rows + (i - n_heavy_tasks)));
            if (lr < 0) continue;
            t = __ldg(tasks + n_heavy_tasks + lr);
        }
        const float4 g = ldg4(G + (int64_t)t.x * ldg + sub * 4);
        if (Z != nullptr && first_chunk && slot == 0) {
            const float4 z = ldg4(Z + (int64_t)t.x * ldz + sub * 4);
            red_add4(Y + (int64_t)t.x * ldy + sub * 4, z.x, z.y, z.z, z.w);
        }
        for (int k = t.y + slot; k < t.z; k += EPW) {
            const int c = __ldg(col_idx + k);
            const float a = __ldg(vals + k);
            red_add4(Y + (int64_t)c * ldy + sub * 4, a * g.x, a * g.y, a * g.z, a * g.w);
        }
    }
}

// any d % 4 == 0, d <= 256: a whole warp per edge, up to two float4 per lane
__global__ void __launch_bounds__(128) spmm_task_kernel_generic(const int4* __restrict__ tasks, int64_t n_tasks,
                                                                const int32_t* __restrict__ col_idx,
                                                                const float* __restrict__ vals, const float* __restrict__ X,
                                                                int64_t ldx, float* __restrict__ Y, int64_t ldy,
                                                                const float* __restrict__ Z, int64_t ldz,
                                                                float* __restrict__ partials, int d) {
    const int lane = threadIdx.x & 31;
    const int64_t task_id = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (task_id >= n_tasks) return;
    const int4 t = __ldg(tasks + task_id);
    const bool has0 = lane * 4 < d, has1 = (lane + 32) * 4 < d;
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
    for (int base = t.y; base < t.z; base += 32) {
        const int k = base + lane;
        int c = 0;
        float v = 0.f;
        if (k < t.z) {
            c = col_idx[k];
            v = vals[k];
        }
        const int cnt = min(32, t.z - base);
        for (int j = 0; j < cnt; ++j) {
            const int cc = __shfl_sync(kFull, c, j);
            const float w = __shfl_sync(kFull, v, j);
            const float* xr = X + (int64_t)cc * ldx;
            if (has0) fma4(a0, w, ldg4(xr + lane * 4));
            if (has1) fma4(a1, w, ldg4(xr + (lane + 32) * 4));
        }
    }
    float* dst;
    int64_t zoff = -1;
    if (t.w < 0) {
        dst = Y + (int64_t)t.x * ldy;
        if (Z != nullptr) zoff = (int64_t)t.x * ldz;
    } else {
        dst = partials + (int64_t)t.w * d;
    }
    if (has0) {
        if (zoff >= 0) { const float4 z = ldg4(Z + zoff + lane * 4); a0.x += z.x; a0.y += z.y; a0.z += z.z; a0.w += z.w; }
        *reinterpret_cast<float4*>(dst + lane * 4) = a0;
    }
    if (has1) {
        if (zoff >= 0) { const float4 z = ldg4(Z + zoff + (lane + 32) * 4); a1.x += z.x; a1.y += z.y; a1.z += z.z; a1.w += z.w; }
        *reinterpret_cast<float4*>(dst + (lane + 32) * 4) = a1;
    }
}

// heavy rows: sum the chunk partials in chunk order (one warp per heavy row)
__global__ void __launch_bounds__(128) spmm_heavy_reduce_kernel(const int4* __restrict__ heavy, int64_t n_heavy,
                                                                const float* __restrict__ partials, float* __restrict__ Y,
                                                                int64_t ldy, const float* __restrict__ Z, int64_t ldz, int d) {
    const int lane = threadIdx.x & 31;
    const int64_t h = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (h >= n_heavy) return;
    const int4 r = __ldg(heavy + h);  // {row, first_slot, n_chunks, 0}
    for (int f = lane * 4; f < d; f += 128) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int c = 0; c < r.z; ++c) {
            const float4 p = ld_stream4(partials + (int64_t)(r.y + c) * d + f);
            a.x += p.x; a.y += p.y; a.z += p.z; a.w += p.w;
        }
        if (Z != nullptr) {
            const float4 z = ld_stream4(Z + (int64_t)r.x * ldz + f);
            a.x += z.x; a.y += z.y; a.z += z.z; a.w += z.w;
        }
        *reinterpret_cast<float4*>(Y + (int64_t)r.x * ldy + f) = a;
    }
}

}  // namespace
}  // namespace kgat

using namespace kgat;

// CTAs per SM of the unmasked row-list kernel (KGAT_SPMM_CTAS = 5 | 6; A/B switch)
static int spmm_plain_ctas() {
    static int c = 0;
    if (c == 0) {
        const char* e = getenv("KGAT_SPMM_CTAS");
        c = (e != nullptr && atoi(e) == 6) ? 6 : 5;  // 6 CTAs (40 registers) spills and measured 2 % slower
    }
    return c;
}

// loads in flight per lane for d = 64 (KGAT_SPMM_U = 4 | 8; A/B switch for the profiling runs).  Measured on captured CF steps
// (tools/prof_cf.py, same box): depth 8 everywhere 647 us, depth 8 for the edge-masked instance only 658 us, depth 4 everywhere
// 663 us.  (Per-kernel eager timings suggested the opposite for the plain forward instance; the captured step is what counts.)
static int spmm_unroll64() {
    static int u = 0;
    if (u == 0) {
        const char* e = getenv("KGAT_SPMM_U");
        u = (e != nullptr && atoi(e) == 4) ? 4 : 8;
    }
    return u;
}

static int spmm_launch(const int32_t* tasks, int64_t n_tasks, int32_t* heavy_rows, int64_t n_heavy, const int32_t* col_idx,
                       const float* vals, const float* X, int64_t n_cols, int64_t ldx, float* Y, int64_t ldy, const float* Z, int64_t ldz,
                       int32_t d, float* partials, const uint32_t* row_mask, const uint32_t* edge_mask, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_tasks < 0 || n_heavy < 0 || d <= 0 || (d & 3) || d > 256) return KGAT_ERR_INVALID_ARGUMENT;
    if ((ldx & 3) || (ldy & 3) || (Z && (ldz & 3))) return KGAT_ERR_INVALID_ARGUMENT;
    if (n_heavy > 0 && partials == nullptr) return KGAT_ERR_INVALID_ARGUMENT;
    if (n_tasks == 0) return KGAT_OK;
    if (n_cols <= 0) return KGAT_ERR_INVALID_ARGUMENT;
    const bool use_wide = n_cols * ldx * 4 >= ((int64_t)1 << 32);  // 32-bit byte offsets cover tables < 4 GiB
    const bool masked = row_mask != nullptr || edge_mask != nullptr;
    const int threads = 128;
    const unsigned blocks = (unsigned)((n_tasks * 32 + threads - 1) / threads);
    const int4* t4 = reinterpret_cast<const int4*>(tasks);
    int4* h4 = reinterpret_cast<int4*>(heavy_rows);
    bool fused_reduce = false;  // the templated kernels reduce heavy rows themselves (last chunk to arrive)
#define KGAT_SPMM_ARGS t4, n_tasks, col_idx, vals, X, ldx, Y, ldy, Z, ldz, partials, h4, (int)n_heavy, row_mask, edge_mask
#define KGAT_SPMM_LAUNCH(DD, UU)                                                                                       \
    do {                                                                                                               \
        if (masked) {                                                                                                  \
            if (use_wide) spmm_task_kernel<DD, UU, true, true><<<blocks, threads, 0, stream>>>(KGAT_SPMM_ARGS);        \
            else spmm_task_kernel<DD, UU, false, true><<<blocks, threads, 0, stream>>>(KGAT_SPMM_ARGS);                \
        } else {                                                                                                       \
            if (use_wide) spmm_task_kernel<DD, UU, true, false><<<blocks, threads, 0, stream>>>(KGAT_SPMM_ARGS);       \
            else spmm_task_kernel<DD, UU, false, false><<<blocks, threads, 0, stream>>>(KGAT_SPMM_ARGS);               \
        }                                                                                                              \
        fused_reduce = true;                                                                                           \
    } while (0)
    switch (d) {
        case 16: KGAT_SPMM_LAUNCH(16, 2); break;
        case 32: KGAT_SPMM_LAUNCH(32, 4); break;
        case 64:
            if (spmm_unroll64() == 8) KGAT_SPMM_LAUNCH(64, 8);
            else KGAT_SPMM_LAUNCH(64, 4);
            break;
        case 128: KGAT_SPMM_LAUNCH(128, 4); break;
        default:
            if (masked) return KGAT_ERR_UNSUPPORTED;
            spmm_task_kernel_generic<<<blocks, threads, 0, stream>>>(t4, n_tasks, col_idx, vals, X, ldx, Y, ldy, Z, ldz, partials, d);
    }
#undef KGAT_SPMM_LAUNCH
#undef KGAT_SPMM_ARGS
    if (n_heavy > 0 && !fused_reduce) {
        const unsigned hb = (unsigned)((n_heavy * 32 + threads - 1) / threads);
        spmm_heavy_reduce_kernel<<<hb, threads, 0, stream>>>(reinterpret_cast<const int4*>(heavy_rows), n_heavy, partials, Y, ldy, Z,
                                                             ldz, d);
    }
    return check_launch();
}

extern "C" int kgat_spmm_csr(const int32_t* tasks, int64_t n_tasks, int32_t* heavy_rows, int64_t n_heavy,
                             const int32_t* col_idx, const float* vals, const float* X, int64_t n_cols, int64_t ldx, float* Y, int64_t ldy,
                             const float* Z, int64_t ldz, int32_t d, float* partials, void* stream_) {
    return spmm_launch(tasks, n_tasks, heavy_rows, n_heavy, col_idx, vals, X, n_cols, ldx, Y, ldy, Z, ldz, d, partials, nullptr, nullptr,
                       stream_);
}

extern "C" int kgat_spmm_csr_masked(const int32_t* tasks, int64_t n_tasks, int32_t* heavy_rows, int64_t n_heavy,
                                    const int32_t* col_idx, const float* vals, const float* X, int64_t n_cols, int64_t ldx, float* Y,
                                    int64_t ldy, const float* Z, int64_t ldz, int32_t d, float* partials, const uint32_t* row_mask,
                                    const uint32_t* edge_mask, void* stream_) {
    return spmm_launch(tasks, n_tasks, heavy_rows, n_heavy, col_idx, vals, X, n_cols, ldx, Y, ldy, Z, ldz, d, partials, row_mask, edge_mask,
                       stream_);
}

extern "C" int kgat_spmm_csr_rows(const int32_t* tasks, int64_t n_tasks, int64_t n_heavy_tasks, const int32_t* light_rank,
                                  int32_t* heavy_rows, int64_t n_heavy, const int32_t* col_idx, const float* vals, const float* X,
                                  int64_t n_cols, int64_t ldx, float* Y, int64_t ldy, const float* Z, int64_t ldz, int32_t d, float* partials,
                                  const int32_t* rows, const int32_t* n_rows_dev, const uint32_t* row_mask, const uint32_t* edge_mask,
                                  int64_t n_mask_bits, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_tasks < 0 || n_heavy < 0 || n_heavy_tasks < 0 || n_heavy_tasks > n_tasks || n_tasks >= ((int64_t)1 << 31)) return KGAT_ERR_INVALID_ARGUMENT;
    if ((ldx & 3) || (ldy & 3) || (Z && (ldz & 3))) return KGAT_ERR_INVALID_ARGUMENT;
    if (n_heavy > 0 && partials == nullptr) return KGAT_ERR_INVALID_ARGUMENT;
    if ((rows == nullptr) != (n_rows_dev == nullptr) || (rows != nullptr && light_rank == nullptr)) return KGAT_ERR_INVALID_ARGUMENT;
    if (rows != nullptr && n_heavy_tasks > 0 && row_mask == nullptr) return KGAT_ERR_INVALID_ARGUMENT;
    if (n_tasks == 0) return KGAT_OK;
    if (n_cols <= 0) return KGAT_ERR_INVALID_ARGUMENT;
    if (d != 16 && d != 32 && d != 64 && d != 128) return KGAT_ERR_UNSUPPORTED;
    const bool use_wide = n_cols * ldx * 4 >= ((int64_t)1 << 32);
    int mask_words = 0;
    if (edge_mask != nullptr) {
        const int64_t words = (n_mask_bits + 31) / 32;
        if (words > 0 && words * 4 <= 32 * 1024) mask_words = (int)words;  // 1 M nodes: a 32 KB bitmap per CTA still leaves 6 CTAs / SM
    }
    const size_t smem = (size_t)mask_words * 4;
    // edge-masked (backward) instance: 48 registers x 256 threads -> 5 CTAs / SM (5 x (32 KB bitmap + 2 KB slabs) of shared memory
    // fit as well); the unmasked (forward) instance is compiled on its own and runs kCtasPlain CTAs / SM
    const bool em = edge_mask != nullptr;
    const unsigned blocks = (unsigned)(sm_count() * (em ? 5 : spmm_plain_ctas()));
    const int4* t4 = reinterpret_cast<const int4*>(tasks);
    int4* h4 = reinterpret_cast<int4*>(heavy_rows);
#define KGAT_ROWS_ARGS t4, (int)n_tasks, (int)n_heavy_tasks, light_rank, rows, n_rows_dev, col_idx, vals, X, ldx, Y, ldy, Z, ldz, partials, h4, \
                       (int)n_heavy, row_mask, edge_mask, mask_words
#define KGAT_ROWS_LAUNCH(DD, UU)                                                                                           \
    do {                                                                                                                   \
        if (em) {                                                                                                          \
            if (use_wide) spmm_rows_kernel<DD, UU, true, true, 5><<<blocks, kRowsThreads, smem, stream>>>(KGAT_ROWS_ARGS);  \
            else spmm_rows_kernel<DD, UU, false, true, 5><<<blocks, kRowsThreads, smem, stream>>>(KGAT_ROWS_ARGS);          \
        } else if (spmm_plain_ctas() == 6) {                                                                               \
            if (use_wide) spmm_rows_kernel<DD, UU, true, false, 6><<<blocks, kRowsThreads, smem, stream>>>(KGAT_ROWS_ARGS); \
            else spmm_rows_kernel<DD, UU, false, false, 6><<<blocks, kRowsThreads, smem, stream>>>(KGAT_ROWS_ARGS);         \
        } else {                                                                                                           \
            if (use_wide) spmm_rows_kernel<DD, UU, true, false, 5><<<blocks, kRowsThreads, smem, stream>>>(KGAT_ROWS_ARGS); \
            else spmm_rows_kernel<DD, UU, false, false, 5><<<blocks, kRowsThreads, smem, stream>>>(KGAT_ROWS_ARGS);         \
        }                                                                                                                  \
    } while (0)
    switch (d) {
        case 16: KGAT_ROWS_LAUNCH(16, 2); break;
        case 32: KGAT_ROWS_LAUNCH(32, 4); break;
        case 64:
            if (spmm_unroll64() == 8) KGAT_ROWS_LAUNCH(64, 8);
            else KGAT_ROWS_LAUNCH(64, 4);
            break;
        default: KGAT_ROWS_LAUNCH(128, 4); break;
    }
#undef KGAT_ROWS_LAUNCH
#undef KGAT_ROWS_ARGS
    return check_launch();
}

extern "C" int kgat_spmm_scatter_rows(const int32_t* tasks, int64_t n_heavy_tasks, const int32_t* light_rank, const int32_t* rows,
                                      const int32_t* n_rows_dev, int64_t max_rows, const uint32_t* row_mask, const int32_t* col_idx,
                                      const float* vals, const float* G, int64_t ldg, const float* Z, int64_t ldz, float* Y, int64_t ldy,
                                      int32_t d, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!tasks || !light_rank || !rows || !n_rows_dev || !col_idx || !vals || !G || !Y || max_rows <= 0 || n_heavy_tasks < 0 ||
        n_heavy_tasks >= ((int64_t)1 << 30) || (n_heavy_tasks > 0 && !row_mask) || (ldg & 3) || (ldy & 3) || (Z && (ldz & 3)))
        return KGAT_ERR_INVALID_ARGUMENT;
    int64_t ctas = (max_rows + n_heavy_tasks + 7) / 8;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (ctas > cap) ctas = cap;
    const int4* t4 = reinterpret_cast<const int4*>(tasks);
#define KGAT_SCATTER(DD)                                                                                                               \
    spmm_scatter_rows_kernel<DD><<<(unsigned)ctas, 256, 0, stream>>>(t4, (int)n_heavy_tasks, light_rank, rows, n_rows_dev, row_mask, col_idx, \
                                                                     vals, G, ldg, Z, ldz, Y, ldy)
    switch (d) {
        case 16: KGAT_SCATTER(16); break;
        case 32: KGAT_SCATTER(32); break;
        case 64: KGAT_SCATTER(64); break;
        case 128: KGAT_SCATTER(128); break;
        default: return KGAT_ERR_UNSUPPORTED;
    }
#undef KGAT_SCATTER
    return check_launch();
}
