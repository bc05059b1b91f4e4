// K4: fused BPR loss over the per-layer embedding tables (reference model.py:189-202, 142-163)
// K5: fused TransR loss without materialising W_r[rel] (reference model.py:204-261)
// Both with hand-written backward kernels.  One warp per sample; the per-sample terms go to a small
// scratch and are reduced in a fixed order by a single-CTA kernel (deterministic loss value).
#include "common.cuh"

namespace kgat {
namespace {

// loss = mean_b(-logsigmoid(margin_b)) + reg * mean_b(l2_b);   scratch = [margin (B)][l2 (B)]
// `pub` (optional): the loss is also published to the host -- ring[s % n_slots] = (s << 32) | bits(loss) with s = ++serial, one
// aligned 8-byte store into mapped pinned memory (what kgat_publish_loss does as a launch of its own)
struct Publish {
    unsigned long long* serial_dev;
    volatile unsigned long long* ring_host;
    int n_slots;
};

__global__ void __launch_bounds__(256) loss_reduce_kernel(const float* __restrict__ scratch, int batch, float reg,
                                                          float* __restrict__ loss, float* __restrict__ loss_sum = nullptr,
                                                          Publish pub = Publish{nullptr, nullptr, 0}) {
    __shared__ float sh_a[8], sh_b[8];
    float a = 0.f, b = 0.f;
    for (int i = threadIdx.x; i < batch; i += 256) {
        a += -log_sigmoid(scratch[i]);
        b += scratch[batch + i];
    }
    a = warp_sum(a);
    b = warp_sum(b);
    if ((threadIdx.x & 31) == 0) {
        sh_a[threadIdx.x >> 5] = a;
        sh_b[threadIdx.x >> 5] = b;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float ta = 0.f, tb = 0.f;
        for (int w = 0; w < 8; ++w) {
            ta += sh_a[w];
            tb += sh_b[w];
        }
        const float l = ta / (float)batch + reg * (tb / (float)batch);
        loss[0] = l;
        if (loss_sum != nullptr) loss_sum[0] += l;
        if (pub.serial_dev != nullptr) {
            const unsigned long long sn = pub.serial_dev[0] + 1ull;
            pub.serial_dev[0] = sn;
            pub.ring_host[sn % (unsigned long long)pub.n_slots] = (sn << 32) | (unsigned long long)__float_as_uint(l);
            __threadfence_system();
        }
    }
}

inline Publish make_publish(const kgat_publish_t* p) {
    if (p == nullptr || p->serial_dev == nullptr || p->ring_host_mapped == nullptr || p->n_slots <= 0) return Publish{nullptr, nullptr, 0};
    return Publish{reinterpret_cast<unsigned long long*>(p->serial_dev), reinterpret_cast<volatile unsigned long long*>(p->ring_host_mapped),
                   p->n_slots};
}

// ---------------------------------------------------------------------------------------------
// BPR
// ---------------------------------------------------------------------------------------------
struct Tables {
    int n;
    int q[KGAT_MAX_LAYERS];  // float4 per row
    const float* p[KGAT_MAX_LAYERS];
    int64_t ld[KGAT_MAX_LAYERS];
};
struct GradTables {
    float* p[KGAT_MAX_LAYERS];
    int64_t ld[KGAT_MAX_LAYERS];
};

__global__ void __launch_bounds__(256) bpr_fwd_kernel(Tables T, const int64_t* __restrict__ users, const int64_t* __restrict__ pos,
                                                      const int64_t* __restrict__ neg, int batch, float* __restrict__ scratch) {
    const int lane = threadIdx.x & 31;
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (b >= batch) return;
    const int64_t u = users[b], p = pos[b], n = neg[b];
    float sp = 0.f, sn = 0.f, l2 = 0.f;
    for (int t = 0; t < T.n; ++t) {
        for (int f = lane; f < T.q[t]; f += 32) {
            const float4 a = ldg4(T.p[t] + u * T.ld[t] + f * 4);
            const float4 x = ldg4(T.p[t] + p * T.ld[t] + f * 4);
            const float4 y = ldg4(T.p[t] + n * T.ld[t] + f * 4);
            sp += a.x * x.x + a.y * x.y + a.z * x.z + a.w * x.w;
            sn += a.x * y.x + a.y * y.y + a.z * y.z + a.w * y.w;
            l2 += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w + x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w + y.x * y.x +
                  y.y * y.y + y.z * y.z + y.w * y.w;
        }
    }
    sp = warp_sum(sp);
    sn = warp_sum(sn);
    l2 = warp_sum(l2);
    if (lane == 0) {
        scratch[b] = sp - sn;
        scratch[batch + b] = 0.5f * l2;
    }
}

__global__ void __launch_bounds__(256) bpr_bwd_kernel(Tables T, GradTables G, const int64_t* __restrict__ users,
                                                      const int64_t* __restrict__ pos, const int64_t* __restrict__ neg, int batch,
                                                      float reg, const float* __restrict__ scratch,
                                                      const float* __restrict__ g_loss) {
    const int lane = threadIdx.x & 31;
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (b >= batch) return;
    const int64_t u = users[b], p = pos[b], n = neg[b];
    const float g = g_loss[0] / (float)batch;
    const float s = sigmoidf_(-scratch[b]) * g;  // d(-logsigmoid(m))/dm = -sigmoid(-m)
    const float lam = reg * g;
    for (int t = 0; t < T.n; ++t) {
        if (G.p[t] == nullptr) continue;
        for (int f = lane; f < T.q[t]; f += 32) {
            const float4 a = ldg4(T.p[t] + u * T.ld[t] + f * 4);
            const float4 x = ldg4(T.p[t] + p * T.ld[t] + f * 4);
            const float4 y = ldg4(T.p[t] + n * T.ld[t] + f * 4);
            float* gu = G.p[t] + u * G.ld[t] + f * 4;
            float* gp = G.p[t] + p * G.ld[t] + f * 4;
            float* gn = G.p[t] + n * G.ld[t] + f * 4;
            atomicAdd(gu + 0, s * (y.x - x.x) + lam * a.x);
            atomicAdd(gu + 1, s * (y.y - x.y) + lam * a.y);
            atomicAdd(gu + 2, s * (y.z - x.z) + lam * a.z);
            atomicAdd(gu + 3, s * (y.w - x.w) + lam * a.w);
            atomicAdd(gp + 0, -s * a.x + lam * x.x);
            atomicAdd(gp + 1, -s * a.y + lam * x.y);
            atomicAdd(gp + 2, -s * a.z + lam * x.z);
            atomicAdd(gp + 3, -s * a.w + lam * x.w);
            atomicAdd(gn + 0, s * a.x + lam * y.x);
            atomicAdd(gn + 1, s * a.y + lam * y.y);
            atomicAdd(gn + 2, s * a.z + lam * y.z);
            atomicAdd(gn + 3, s * a.w + lam * y.w);
        }
    }
}

int pack_tables(const kgat_tables_t* t, Tables* out) {
    if (!t || t->n_tables <= 0 || t->n_tables > KGAT_MAX_LAYERS) return KGAT_ERR_INVALID_ARGUMENT;
    out->n = t->n_tables;
    for (int i = 0; i < t->n_tables; ++i) {
        if (t->dims[i] <= 0 || (t->dims[i] & 3) || (t->lds[i] & 3) || !t->tables[i]) return KGAT_ERR_INVALID_ARGUMENT;
        out->q[i] = t->dims[i] / 4;
        out->p[i] = t->tables[i];
        out->ld[i] = t->lds[i];
    }
    return KGAT_OK;
}

// ---------------------------------------------------------------------------------------------
// TransR:  x = e W_r  (row-vector convention).  One CTA of 4 warps per sample: the kernels are latency
// bound (512 samples, a 64-step dependent mat-vec each), so the rows j of W_r are split over the 4 warps
// and the partial projections are combined through shared memory.  Lane owns the KM consecutive output columns from KM * lane (ld_cols / red_cols).
// ---------------------------------------------------------------------------------------------
constexpr int kTrWarps = 4;

// A lane owns the KM CONSECUTIVE output columns [KM lane, KM lane + KM): its slice of a row of W_r / e_r is one 4 KM-byte load
// and its slice of a gradient row ONE vector reduction (red.global.add.v2/.v4.f32) -- the W_r gradient is 2.1 M scalar float
// atomics per 512-sample batch otherwise, which is what bounds the kernel.
template <int KM>
__device__ __forceinline__ void ld_cols(const float* __restrict__ p, int lane, float (&x)[KM]) {
    if constexpr (KM == 1) {
        x[0] = __ldg(p + lane);
    } else if constexpr (KM == 2) {
        const float2 v = __ldg(reinterpret_cast<const float2*>(p) + lane);
        x[0] = v.x; x[1] = v.y;
    } else {
        static_assert(KM == 4, "KM is 1, 2 or 4");
        const float4 v = __ldg(reinterpret_cast<const float4*>(p) + lane);
        x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
    }
}
template <int KM>
__device__ __forceinline__ void red_cols(float* __restrict__ p, int lane, const float (&x)[KM]) {
    if constexpr (KM == 1) {
        atomicAdd(p + lane, x[0]);
    } else if constexpr (KM == 2) {
        asm volatile("red.relaxed.gpu.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p + 2 * lane), "f"(x[0]), "f"(x[1]) : "memory");
    } else {
        asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p + 4 * lane), "f"(x[0]), "f"(x[1]), "f"(x[2]), "f"(x[3])
                     : "memory");
    }
}

// partial projections of (e_h, e_p, e_n) over this warp's rows [j0, j0 + JW) of W_r
template <int D, int KM>
__device__ __forceinline__ void transr_partial(const float* __restrict__ Wr, const float* __restrict__ emb, int64_t h, int64_t p, int64_t n,
                                               int warp, int lane, float& eh, float& ep, float& en, float (&xh)[KM], float (&xp)[KM],
                                               float (&xn)[KM]) {
    constexpr int K = KM * 32;
    constexpr int JW = D / kTrWarps;  // rows of W_r per warp (8, 16 or 32)
    const int j0 = warp * JW;
    eh = ep = en = 0.f;
    if (lane < JW) {
        eh = __ldg(emb + h * D + j0 + lane);
        ep = __ldg(emb + p * D + j0 + lane);
        en = __ldg(emb + n * D + j0 + lane);
    }
#pragma unroll
    for (int m = 0; m < KM; ++m) xh[m] = xp[m] = xn[m] = 0.f;
#pragma unroll 8
    for (int jj = 0; jj < JW; ++jj) {
        const float a = __shfl_sync(kFull, eh, jj);
        const float b = __shfl_sync(kFull, ep, jj);
        const float c = __shfl_sync(kFull, en, jj);
        float w[KM];
        ld_cols<KM>(Wr + (j0 + jj) * K, lane, w);
#pragma unroll
        for (int m = 0; m < KM; ++m) {
            xh[m] = fmaf(a, w[m], xh[m]);
            xp[m] = fmaf(b, w[m], xp[m]);
            xn[m] = fmaf(c, w[m], xn[m]);
        }
    }
}

// combine the 4 warps' partials: every warp returns with the full projections
template <int KM>
__device__ __forceinline__ void transr_combine(float (*part)[3][KM * 32], int warp, int lane, float (&xh)[KM], float (&xp)[KM],
                                               float (&xn)[KM]) {
#pragma unroll
    for (int m = 0; m < KM; ++m) {
        part[warp][0][lane + 32 * m] = xh[m];
        part[warp][1][lane + 32 * m] = xp[m];
        part[warp][2][lane + 32 * m] = xn[m];
    }
    __syncthreads();
#pragma unroll
    for (int m = 0; m < KM; ++m) {
        float a = 0.f, b = 0.f, c = 0.f;
#pragma unroll
        for (int w = 0; w < kTrWarps; ++w) {
            a += part[w][0][lane + 32 * m];
            b += part[w][1][lane + 32 * m];
            c += part[w][2][lane + 32 * m];
        }
        xh[m] = a;
        xp[m] = b;
        xn[m] = c;
    }
}

template <int DM, int KM>
__global__ void __launch_bounds__(128) transr_fwd_kernel(const float* __restrict__ emb, const float* __restrict__ rel_emb,
                                                         const float* __restrict__ W, const int64_t* __restrict__ heads,
                                                         const int64_t* __restrict__ rels, const int64_t* __restrict__ pt,
                                                         const int64_t* __restrict__ nt, int batch, float* __restrict__ scratch) {
    constexpr int D = DM * 32, K = KM * 32;
    __shared__ float part[kTrWarps][3][K];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.x;
    const int64_t h = heads[b], r = rels[b], p = pt[b], n = nt[b];
    float eh, ep, en, xh[KM], xp[KM], xn[KM];
    transr_partial<D, KM>(W + r * (int64_t)D * K, emb, h, p, n, warp, lane, eh, ep, en, xh, xp, xn);
    transr_combine<KM>(part, warp, lane, xh, xp, xn);
    if (warp != 0) return;
    float ps = 0.f, ns = 0.f, l2 = 0.f;
    float er[KM];
    ld_cols<KM>(rel_emb + r * K, lane, er);
#pragma unroll
    for (int m = 0; m < KM; ++m) {
        const float dp = xh[m] + er[m] - xp[m];
        const float dn = xh[m] + er[m] - xn[m];
        ps = fmaf(dp, dp, ps);
        ns = fmaf(dn, dn, ns);
        l2 += xh[m] * xh[m] + er[m] * er[m] + xp[m] * xp[m] + xn[m] * xn[m];
    }
    ps = warp_sum(ps);
    ns = warp_sum(ns);
    l2 = warp_sum(l2);
    if (lane == 0) {
        scratch[b] = ns - ps;
        scratch[batch + b] = 0.5f * l2;
    }
}

// FUSED: forward and backward in one pass (the backward recomputes the projections anyway): the kernel derives the
// sample's margin itself, records it (and the L2 term) in `scratch` for the loss reduction that follows, and uses
// d loss / d batch-loss = 1.
template <int DM, int KM, bool FUSED = false>
__global__ void __launch_bounds__(128) transr_bwd_kernel(const float* __restrict__ emb, const float* __restrict__ rel_emb,
                                                         const float* __restrict__ W, const int64_t* __restrict__ heads,
                                                         const int64_t* __restrict__ rels, const int64_t* __restrict__ pt,
                                                         const int64_t* __restrict__ nt, int batch, float reg,
                                                         float* __restrict__ scratch, const float* __restrict__ g_loss,
                                                         float* __restrict__ g_emb, float* __restrict__ g_rel,
                                                         float* __restrict__ g_W, const int32_t* __restrict__ row_slot) {
    constexpr int D = DM * 32, K = KM * 32;
    constexpr int JW = D / kTrWarps;
    __shared__ float part[kTrWarps][3][K];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.x;
    const int64_t h = heads[b], r = rels[b], p = pt[b], n = nt[b];
    const float* Wr = W + r * (int64_t)D * K;
    float eh, ep, en, xh[KM], xp[KM], xn[KM];
    transr_partial<D, KM>(Wr, emb, h, p, n, warp, lane, eh, ep, en, xh, xp, xn);
    transr_combine<KM>(part, warp, lane, xh, xp, xn);

    float er[KM];
    ld_cols<KM>(rel_emb + r * K, lane, er);
    float margin;
    if (FUSED) {
        float ps = 0.f, ns = 0.f, l2 = 0.f;
#pragma unroll
        for (int m = 0; m < KM; ++m) {
            const float dp = xh[m] + er[m] - xp[m];
            const float dn = xh[m] + er[m] - xn[m];
            ps = fmaf(dp, dp, ps);
            ns = fmaf(dn, dn, ns);
            l2 += xh[m] * xh[m] + er[m] * er[m] + xp[m] * xp[m] + xn[m] * xn[m];
        }
        ps = warp_sum(ps);
        ns = warp_sum(ns);
        l2 = warp_sum(l2);
        margin = ns - ps;
        if (warp == 0 && lane == 0) {
            scratch[b] = margin;
            scratch[batch + b] = 0.5f * l2;
        }
    } else {
        margin = scratch[b];
    }
    const float g = (FUSED ? 1.f : g_loss[0]) / (float)batch;
    const float s2 = 2.f * sigmoidf_(-margin) * g;  // loss = -logsigmoid(ns - ps)
    const float lam = reg * g;
    float gxh[KM], gxp[KM], gxn[KM], ger[KM];
#pragma unroll
    for (int m = 0; m < KM; ++m) {
        const float dp = xh[m] + er[m] - xp[m];
        const float dn = xh[m] + er[m] - xn[m];
        const float gdp = s2 * dp;   // dL/d dpos
        const float gdn = -s2 * dn;  // dL/d dneg
        gxh[m] = gdp + gdn + lam * xh[m];
        gxp[m] = -gdp + lam * xp[m];
        gxn[m] = -gdn + lam * xn[m];
        ger[m] = gdp + gdn + lam * er[m];
    }
    if (warp == 0) red_cols<KM>(g_rel + r * K, lane, ger);
    // this warp's rows of W_r: d e[j] = <g_x, W_r[j, :]>, d W_r[j, :] += e[j] * g_x
    const int j0 = warp * JW;
    float* gWr = g_W + r * (int64_t)D * K;
    float geh = 0.f, gep = 0.f, gen = 0.f;
#pragma unroll 4
    for (int jj = 0; jj < JW; ++jj) {
        const int j = j0 + jj;
        const float a = __shfl_sync(kFull, eh, jj);
        const float bb = __shfl_sync(kFull, ep, jj);
        const float c = __shfl_sync(kFull, en, jj);
        float ph = 0.f, pp = 0.f, pn = 0.f;
        float w[KM], gw[KM];
        ld_cols<KM>(Wr + j * K, lane, w);
#pragma unroll
        for (int m = 0; m < KM; ++m) {
            ph = fmaf(gxh[m], w[m], ph);
            pp = fmaf(gxp[m], w[m], pp);
            pn = fmaf(gxn[m], w[m], pn);
            gw[m] = a * gxh[m] + bb * gxp[m] + c * gxn[m];
        }
        red_cols<KM>(gWr + j * K, lane, gw);
        ph = warp_sum(ph);
        pp = warp_sum(pp);
        pn = warp_sum(pn);
        if (lane == jj) {
            geh = ph;
            gep = pp;
            gen = pn;
        }
    }
    if (lane < JW) {
        // dense gradient table indexed by node, or compact rows indexed through row_slot (kgat_transr_claim_rows)
        const int64_t gh = row_slot ? row_slot[h] : h, gp = row_slot ? row_slot[p] : p, gn = row_slot ? row_slot[n] : n;
        atomicAdd(g_emb + gh * D + j0 + lane, geh);
        atomicAdd(g_emb + gp * D + j0 + lane, gep);
        atomicAdd(g_emb + gn * D + j0 + lane, gen);
    }
}

// Compact gradient rows for the embedding table: the 3B ids of a batch claim a slot per DISTINCT node
// (row_slot[node] = index of the first claimant, -1 = untouched; the Adam kernel resets the claims), and the
// 3B x d gradient rows are zeroed -- instead of zeroing and re-reading an n_nodes x d gradient table per step.
__global__ void transr_claim_rows_kernel(const int64_t* __restrict__ heads, const int64_t* __restrict__ pt, const int64_t* __restrict__ nt,
                                         int batch, int d, int32_t* __restrict__ row_slot, float4* __restrict__ g_rows,
                                         float4* __restrict__ zero_a = nullptr, int n_a = 0, float4* __restrict__ zero_b = nullptr,
                                         int n_b = 0) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 3 * batch * (d / 4)) g_rows[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < n_a) zero_a[i] = make_float4(0.f, 0.f, 0.f, 0.f);  // the step's other gradient buffers, zeroed in the same launch
    if (i < n_b) zero_b[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < 3 * batch) {
        const int64_t id = i < batch ? heads[i] : (i < 2 * batch ? pt[i - batch] : nt[i - 2 * batch]);
        atomicCAS(row_slot + id, -1, i);
    }
}

// T[ids[i], :] = 0: re-zero the few rows a sparse scatter touched instead of memsetting the whole table
__global__ void zero_rows_i64_kernel(float* __restrict__ T, int64_t ld, int d4, const int64_t* __restrict__ ids, int64_t n_ids,
                                     int64_t n_rows) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_ids * d4) return;
    const int64_t r = ids[i / d4];
    if (r < 0 || r >= n_rows) return;
    reinterpret_cast<float4*>(T + r * ld)[i % d4] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// dense[id, :] = g_rows[slot, :] for every batch id whose claim this entry won (row_slot[id] == its index): the dense
// embedding gradient autograd would hand out, restricted to the <= 3B rows that are not zero
__global__ void transr_rows_to_dense_kernel(const float4* __restrict__ g_rows, const int32_t* __restrict__ row_slot,
                                            const int64_t* __restrict__ heads, const int64_t* __restrict__ pt, const int64_t* __restrict__ nt,
                                            int batch, int d4, float* __restrict__ dense, int64_t ld, int64_t* __restrict__ keep_h,
                                            int64_t* __restrict__ keep_pt, int64_t* __restrict__ keep_nt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 3 * batch * d4) return;
    const int e = i / d4, q = i % d4;
    const int64_t id = e < batch ? heads[e] : (e < 2 * batch ? pt[e - batch] : nt[e - 2 * batch]);
    if (keep_h != nullptr && q == 0) {  // remember whose rows the dense view now holds (cleared before the next batch's rows go in)
        if (e < batch) keep_h[e] = id;
        else if (e < 2 * batch) keep_pt[e - batch] = id;
        else keep_nt[e - 2 * batch] = id;
    }
    if (row_slot[id] != e) return;
    reinterpret_cast<float4*>(dense + id * ld)[q] = g_rows[(int64_t)e * d4 + q];
}

// dense[ids[i], :] = 0 and row_slot[ids[i]] = -1: undo what the previous batch left behind
__global__ void transr_release_rows_kernel(float* __restrict__ dense, int64_t ld, int d4, const int64_t* __restrict__ ids, int64_t n_ids,
                                           int64_t n_rows, int32_t* __restrict__ row_slot) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_ids * d4) return;
    const int64_t r = ids[i / d4];
    if (r < 0 || r >= n_rows) return;
    reinterpret_cast<float4*>(dense + r * ld)[i % d4] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i % d4 == 0) row_slot[r] = -1;
}

}  // namespace
}  // namespace kgat

using namespace kgat;

extern "C" {

int kgat_transr_rows_to_dense(const float* g_rows, const int32_t* row_slot, const int64_t* heads, const int64_t* pos_tails,
                              const int64_t* neg_tails, int32_t batch, int32_t d, float* dense, int64_t ld, int64_t* keep_heads,
                              int64_t* keep_pos_tails, int64_t* keep_neg_tails, void* stream) {
    if ((keep_heads != nullptr) != (keep_pos_tails != nullptr) || (keep_heads != nullptr) != (keep_neg_tails != nullptr)) return KGAT_ERR_INVALID_ARGUMENT;
    if (!g_rows || !row_slot || !heads || !pos_tails || !neg_tails || !dense || batch <= 0 || d <= 0 || (d & 3) || (ld & 3))
        return KGAT_ERR_INVALID_ARGUMENT;
    const int total = 3 * batch * (d / 4);
    transr_rows_to_dense_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(g_rows), row_slot, heads,
                                                                                    pos_tails, neg_tails, batch, d / 4, dense, ld, keep_heads,
                                                                                    keep_pos_tails, keep_neg_tails);
    return check_launch();
}

int kgat_transr_release_rows(float* dense, int64_t n_rows, int64_t ld, int32_t d, const int64_t* ids64, int64_t n_ids, int32_t* row_slot,
                             void* stream) {
    if (!dense || !ids64 || !row_slot || n_rows <= 0 || d <= 0 || (d & 3) || (ld & 3) || n_ids < 0) return KGAT_ERR_INVALID_ARGUMENT;
    if (n_ids == 0) return KGAT_OK;
    const int64_t total = n_ids * (d / 4);
    transr_release_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(dense, ld, d / 4, ids64, n_ids, n_rows, row_slot);
    return check_launch();
}

int kgat_zero_rows_i64(float* T, int64_t n_rows, int64_t ld, int32_t d, const int64_t* ids64, int64_t n_ids, void* stream) {
    if (!T || !ids64 || n_rows <= 0 || d <= 0 || (d & 3) || (ld & 3) || n_ids < 0) return KGAT_ERR_INVALID_ARGUMENT;
    if (n_ids == 0) return KGAT_OK;
    const int64_t total = n_ids * (d / 4);
    zero_rows_i64_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(T, ld, d / 4, ids64, n_ids, n_rows);
    return check_launch();
}

int kgat_bpr_forward(const kgat_tables_t* tables, const int64_t* users, const int64_t* pos, const int64_t* neg, int32_t batch,
                     float reg, float* loss, float* loss_sum, float* margin, const kgat_publish_t* publish, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    Tables T;
    int rc = pack_tables(tables, &T);
    if (rc != KGAT_OK) return rc;
    if (batch <= 0) return KGAT_ERR_INVALID_ARGUMENT;
    bpr_fwd_kernel<<<(batch * 32 + 255) / 256, 256, 0, stream>>>(T, users, pos, neg, batch, margin);
    loss_reduce_kernel<<<1, 256, 0, stream>>>(margin, batch, reg, loss, loss_sum, make_publish(publish));
    return check_launch();
}

int kgat_bpr_backward(const kgat_tables_t* tables, const kgat_grad_tables_t* grads, const int64_t* users, const int64_t* pos,
                      const int64_t* neg, int32_t batch, float reg, const float* margin, const float* g_loss, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    Tables T;
    int rc = pack_tables(tables, &T);
    if (rc != KGAT_OK) return rc;
    if (batch <= 0 || !grads || grads->n_tables != tables->n_tables) return KGAT_ERR_INVALID_ARGUMENT;
    GradTables G;
    for (int i = 0; i < T.n; ++i) {
        if (grads->tables[i] && (grads->dims[i] != tables->dims[i] || (grads->lds[i] & 3))) return KGAT_ERR_INVALID_ARGUMENT;
        G.p[i] = grads->tables[i];
        G.ld[i] = grads->lds[i];
    }
    bpr_bwd_kernel<<<(batch * 32 + 255) / 256, 256, 0, stream>>>(T, G, users, pos, neg, batch, reg, margin, g_loss);
    return check_launch();
}

#define KGAT_TRANSR_DISPATCH(CALL)                                            \
    if (d == 32 && k == 32) { constexpr int DM = 1, KM = 1; CALL; }           \
    else if (d == 64 && k == 64) { constexpr int DM = 2, KM = 2; CALL; }      \
    else if (d == 128 && k == 128) { constexpr int DM = 4, KM = 4; CALL; }    \
    else return KGAT_ERR_UNSUPPORTED;

int kgat_transr_forward(const float* emb, const float* rel_emb, const float* W, int32_t d, int32_t k, const int64_t* heads,
                        const int64_t* rels, const int64_t* pos_tails, const int64_t* neg_tails, int32_t batch, float reg, float* loss,
                        float* margin, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (batch <= 0) return KGAT_ERR_INVALID_ARGUMENT;
    const unsigned blocks = (unsigned)batch;  // one 4-warp CTA per sample
    KGAT_TRANSR_DISPATCH((transr_fwd_kernel<DM, KM><<<blocks, 128, 0, stream>>>(emb, rel_emb, W, heads, rels, pos_tails, neg_tails,
                                                                               batch, margin)));
    loss_reduce_kernel<<<1, 256, 0, stream>>>(margin, batch, reg, loss);
    return check_launch();
}

int kgat_transr_backward(const float* emb, const float* rel_emb, const float* W, int32_t d, int32_t k, const int64_t* heads,
                         const int64_t* rels, const int64_t* pos_tails, const int64_t* neg_tails, int32_t batch, float reg,
                         const float* margin, const float* g_loss, float* g_emb, float* g_rel_emb, float* g_W, const int32_t* row_slot,
                         void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (batch <= 0) return KGAT_ERR_INVALID_ARGUMENT;
    const unsigned blocks = (unsigned)batch;  // one 4-warp CTA per sample
    KGAT_TRANSR_DISPATCH((transr_bwd_kernel<DM, KM><<<blocks, 128, 0, stream>>>(emb, rel_emb, W, heads, rels, pos_tails, neg_tails,
                                                                               batch, reg, const_cast<float*>(margin), g_loss, g_emb,
                                                                               g_rel_emb, g_W, row_slot)));
    return check_launch();
}

int kgat_transr_step(const float* emb, const float* rel_emb, const float* W, int32_t d, int32_t k, int32_t n_rel, const int64_t* heads,
                     const int64_t* rels, const int64_t* pos_tails, const int64_t* neg_tails, int32_t batch, float reg, float* loss,
                     float* loss_sum, float* margin, int32_t* row_slot, float* g_rows, float* g_rel_emb, float* g_W,
                     const kgat_publish_t* publish, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (batch <= 0 || d <= 0 || (d & 3) || (k & 3) || n_rel <= 0 || !row_slot || !g_rows || !g_rel_emb || !g_W || !loss || !margin)
        return KGAT_ERR_INVALID_ARGUMENT;
    const int n_rows4 = 3 * batch * (d / 4), n_rel4 = n_rel * k / 4, n_w4 = n_rel * d * (k / 4);
    const int n = n_rows4 > n_w4 ? (n_rows4 > n_rel4 ? n_rows4 : n_rel4) : (n_w4 > n_rel4 ? n_w4 : n_rel4);
    transr_claim_rows_kernel<<<(n + 255) / 256, 256, 0, stream>>>(heads, pos_tails, neg_tails, batch, d, row_slot,
                                                                 reinterpret_cast<float4*>(g_rows), reinterpret_cast<float4*>(g_rel_emb),
                                                                 n_rel4, reinterpret_cast<float4*>(g_W), n_w4);
    const unsigned blocks = (unsigned)batch;
    KGAT_TRANSR_DISPATCH((transr_bwd_kernel<DM, KM, true><<<blocks, 128, 0, stream>>>(emb, rel_emb, W, heads, rels, pos_tails, neg_tails,
                                                                                     batch, reg, margin, nullptr, g_rows, g_rel_emb, g_W,
                                                                                     row_slot)));
    loss_reduce_kernel<<<1, 256, 0, stream>>>(margin, batch, reg, loss, loss_sum, make_publish(publish));
    return check_launch();
}

/* kgat_transr_step without its first launch: the compact rows were claimed and the gradient buffers zeroed by the caller
 * (kgat_adam_rolling_prepare does both while it brings the batch rows up to date) */
int kgat_transr_step_claimed(const float* emb, const float* rel_emb, const float* W, int32_t d, int32_t k, const int64_t* heads,
                             const int64_t* rels, const int64_t* pos_tails, const int64_t* neg_tails, int32_t batch, float reg, float* loss,
                             float* loss_sum, float* margin, const int32_t* row_slot, float* g_rows, float* g_rel_emb, float* g_W,
                             const kgat_publish_t* publish, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (batch <= 0 || d <= 0 || !row_slot || !g_rows || !g_rel_emb || !g_W || !loss || !margin) return KGAT_ERR_INVALID_ARGUMENT;
    const unsigned blocks = (unsigned)batch;
    KGAT_TRANSR_DISPATCH((transr_bwd_kernel<DM, KM, true><<<blocks, 128, 0, stream>>>(emb, rel_emb, W, heads, rels, pos_tails, neg_tails,
                                                                                     batch, reg, margin, nullptr, g_rows, g_rel_emb, g_W,
                                                                                     row_slot)));
    loss_reduce_kernel<<<1, 256, 0, stream>>>(margin, batch, reg, loss, loss_sum, make_publish(publish));
    return check_launch();
}

int kgat_transr_claim_rows(const int64_t* heads, const int64_t* pos_tails, const int64_t* neg_tails, int32_t batch, int32_t d,
                           int32_t* row_slot, float* g_rows, void* stream_) {
    if (batch <= 0 || d <= 0 || (d & 3) || !row_slot || !g_rows) return KGAT_ERR_INVALID_ARGUMENT;
    const int n = 3 * batch * (d / 4);
    transr_claim_rows_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream_>>>(heads, pos_tails, neg_tails, batch, d, row_slot,
                                                                              reinterpret_cast<float4*>(g_rows));
    return check_launch();
}

}  // extern "C"
