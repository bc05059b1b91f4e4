// P5: BPR / KG batch samplers on the device (reference preprocess.py:328-530).
//
// Same sampling *semantics* as the reference's per-sample Python loops:
//   CF batch: B distinct users (with replacement only if fewer than B users have interactions), one uniformly
//             drawn positive item of the user, one uniformly drawn item the user has NOT interacted with (rejection);
//   KG batch: B distinct heads, one uniformly drawn (relation, tail) of the head, one uniformly drawn node that is
//             not a tail of (head, relation) (rejection).
// The reference draws from an unseeded numpy Generator, so parity is distributional (validity + uniformity are
// tested; the RNG-stream-exact replay of the reference sampler lives in the oracle).  Randomness: Philox keyed by
// (seed, step counter in device memory, sample index), so a captured CUDA graph draws a fresh batch every replay.
// "Distinct" is obtained without a shuffle: sample i takes element perm(i) of a keyed pseudo-random permutation
// of [0, n) (4-round Feistel network on the next power of four, cycle-walking).
#include "common.cuh"

namespace kgat {
namespace {

__device__ __forceinline__ uint32_t feistel_perm(uint32_t i, uint32_t n, uint64_t key) {
    int bits = 2;
    while ((1u << bits) < n) bits += 2;  // even number of bits: two equal halves
    const int half = bits >> 1;
    const uint32_t mask = (1u << half) - 1u;
    uint32_t x = i;
    do {
        uint32_t l = x >> half, r = x & mask;
#pragma unroll
        for (int round = 0; round < 4; ++round) {
            uint32_t f = (r + (uint32_t)(key >> (round * 16))) * 0x9E3779B1u;
            f ^= f >> 15;
            f *= 0x85EBCA77u;
            f ^= f >> 13;
            const uint32_t nl = r;
            r = (l ^ f) & mask;
            l = nl;
        }
        x = (l << half) | r;
    } while (x >= n);  // cycle-walk back into [0, n)
    return x;
}

// uniform integer in [0, n) from 32 random bits (multiply-shift; bias < n / 2^32)
__device__ __forceinline__ uint32_t bounded(uint32_t bits, uint32_t n) { return (uint32_t)(((uint64_t)bits * n) >> 32); }

__device__ __forceinline__ bool sorted_contains(const int32_t* __restrict__ a, int begin, int end, int32_t v) {
    int lo = begin, hi = end;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo < end && a[lo] == v;
}

__global__ void sample_cf_kernel(const int32_t* __restrict__ user_ptr, const int32_t* __restrict__ user_items,
                                 const int32_t* __restrict__ active_users, int n_active, int item_num, int batch, uint64_t seed,
                                 const int64_t* __restrict__ step, int64_t* __restrict__ out /* [3][batch] */) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= batch) return;
    const uint64_t s = (uint64_t)step[0];
    const uint64_t key = seed * 0x9E3779B97F4A7C15ull + s * 0xD1B54A32D192ED03ull;
    uint4 rnd = philox4x32(seed, (s << 20) + (uint64_t)i * 64);
    const int slot = batch <= n_active ? (int)feistel_perm((uint32_t)i, (uint32_t)n_active, key) : (int)bounded(rnd.x, n_active);
    const int u = active_users[slot];
    const int b = user_ptr[u], e = user_ptr[u + 1];
    const int pos = user_items[b + bounded(rnd.y, e - b)];
    int neg = (int)bounded(rnd.z, item_num);
    for (int tries = 1; tries < 64 && sorted_contains(user_items, b, e, neg); ++tries) {
        rnd = philox4x32(seed, (s << 20) + (uint64_t)i * 64 + tries);
        neg = (int)bounded(rnd.x, item_num);
    }
    out[i] = u;
    out[batch + i] = pos;
    out[2 * batch + i] = neg;
}

// edges of a head are sorted by tail: is (rel, tail) among them?
__device__ __forceinline__ bool head_has(const int32_t* __restrict__ tails, const int32_t* __restrict__ rels, int lo, int hi, int rel,
                                         int tail) {
    int a = lo, b = hi;
    while (a < b) {
        const int mid = (a + b) >> 1;
        if (tails[mid] < tail) a = mid + 1; else b = mid;
    }
    for (; a < hi && tails[a] == tail; ++a)
        if (rels[a] == rel) return true;
    return false;
}

__global__ void sample_kg_kernel(const int32_t* __restrict__ head_ptr, const int32_t* __restrict__ edge_rel,
                                 const int32_t* __restrict__ edge_tail, const int32_t* __restrict__ active_heads, int n_active,
                                 int node_num, int batch, uint64_t seed, const int64_t* __restrict__ step,
                                 int64_t* __restrict__ out /* [4][batch] */) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= batch) return;
    const uint64_t s = (uint64_t)step[0];
    const uint64_t key = seed * 0xC2B2AE3D27D4EB4Full + s * 0x9E3779B97F4A7C15ull;
    uint4 rnd = philox4x32(seed ^ 0x5bd1e995u, (s << 20) + (uint64_t)i * 64);
    const int slot = batch <= n_active ? (int)feistel_perm((uint32_t)i, (uint32_t)n_active, key) : (int)bounded(rnd.x, n_active);
    const int h = active_heads[slot];
    const int b = head_ptr[h], e = head_ptr[h + 1];
    const int pick = b + (int)bounded(rnd.y, e - b);
    const int rel = edge_rel[pick], tail = edge_tail[pick];
    int neg = (int)bounded(rnd.z, node_num);
    for (int tries = 1; tries < 64 && head_has(edge_tail, edge_rel, b, e, rel, neg); ++tries) {
        rnd = philox4x32(seed ^ 0x5bd1e995u, (s << 20) + (uint64_t)i * 64 + tries);
        neg = (int)bounded(rnd.x, node_num);
    }
    out[i] = h;
    out[batch + i] = rel;
    out[2 * batch + i] = tail;
    out[3 * batch + i] = neg;
}

}  // namespace
}  // namespace kgat

using namespace kgat;

extern "C" {

int kgat_sample_cf_batch(const int32_t* user_ptr, const int32_t* user_items, const int32_t* active_users, int32_t n_active,
                         int32_t item_num, int32_t batch, uint64_t seed, const int64_t* step_dev, int64_t* out, void* stream) {
    if (n_active <= 0 || item_num <= 0 || batch <= 0 || !step_dev) return KGAT_ERR_INVALID_ARGUMENT;
    sample_cf_kernel<<<(batch + 127) / 128, 128, 0, (cudaStream_t)stream>>>(user_ptr, user_items, active_users, n_active, item_num, batch,
                                                                          seed, step_dev, out);
    return check_launch();
}

int kgat_sample_kg_batch(const int32_t* head_ptr, const int32_t* edge_rel, const int32_t* edge_tail, const int32_t* active_heads,
                         int32_t n_active, int32_t node_num, int32_t batch, uint64_t seed, const int64_t* step_dev, int64_t* out,
                         void* stream) {
    if (n_active <= 0 || node_num <= 0 || batch <= 0 || !step_dev) return KGAT_ERR_INVALID_ARGUMENT;
    sample_kg_kernel<<<(batch + 127) / 128, 128, 0, (cudaStream_t)stream>>>(head_ptr, edge_rel, edge_tail, active_heads, n_active, node_num,
                                                                          batch, seed, step_dev, out);
    return check_launch();
}

}  // extern "C"
