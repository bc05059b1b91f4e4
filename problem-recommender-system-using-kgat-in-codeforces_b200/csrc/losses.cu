// K4: fused BPR loss over the per-layer embedding tables (reference model.py:189-202, 142-163)
// K5: fused TransR loss without materialising W_r[rel] (reference model.py:204-261)
// Both with hand-written backward kernels.  One warp per sample; the per-sample terms go to a small
// scratch and are reduced in a fixed order by a single-CTA kernel (deterministic loss value).
#include "common.cuh"

namespace kgat {
namespace {

// loss = mean_b(-logsigmoid(margin_b)) + reg * mean_b(l2_b);   scratch = [margin (B)][l2 (B)]
__global__ void __launch_bounds__(256) loss_reduce_kernel(const float* __restrict__ scratch, int batch, float reg,
                                                          float* __restrict__ loss) {
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
        loss[0] = ta / (float)batch + reg * (tb / (float)batch);
    }
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
// TransR:  x = e W_r  (row-vector convention), lane owns output columns lane + 32 m
// ---------------------------------------------------------------------------------------------
template <int DM, int KM>
__device__ __forceinline__ void transr_project(const float* __restrict__ Wr, const float (&eh)[DM], const float (&ep)[DM],
                                               const float (&en)[DM], int lane, float (&xh)[KM], float (&xp)[KM], float (&xn)[KM]) {
    constexpr int K = KM * 32;
#pragma unroll
    for (int m = 0; m < KM; ++m) xh[m] = xp[m] = xn[m] = 0.f;
#pragma unroll
    for (int jm = 0; jm < DM; ++jm) {
#pragma unroll 8
        for (int jj = 0; jj < 32; ++jj) {
            const float a = __shfl_sync(kFull, eh[jm], jj);
            const float b = __shfl_sync(kFull, ep[jm], jj);
            const float c = __shfl_sync(kFull, en[jm], jj);
            const float* wrow = Wr + (jm * 32 + jj) * K + lane;
#pragma unroll
            for (int m = 0; m < KM; ++m) {
                const float w = __ldg(wrow + 32 * m);
                xh[m] = fmaf(a, w, xh[m]);
                xp[m] = fmaf(b, w, xp[m]);
                xn[m] = fmaf(c, w, xn[m]);
            }
        }
    }
}

template <int DM, int KM>
__global__ void __launch_bounds__(128) transr_fwd_kernel(const float* __restrict__ emb, const float* __restrict__ rel_emb,
                                                         const float* __restrict__ W, const int64_t* __restrict__ heads,
                                                         const int64_t* __restrict__ rels, const int64_t* __restrict__ pt,
                                                         const int64_t* __restrict__ nt, int batch, float* __restrict__ scratch) {
    constexpr int D = DM * 32, K = KM * 32;
    const int lane = threadIdx.x & 31;
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (b >= batch) return;
    const int64_t h = heads[b], r = rels[b], p = pt[b], n = nt[b];
    float eh[DM], ep[DM], en[DM], er[KM], xh[KM], xp[KM], xn[KM];
#pragma unroll
    for (int m = 0; m < DM; ++m) {
        eh[m] = __ldg(emb + h * D + lane + 32 * m);
        ep[m] = __ldg(emb + p * D + lane + 32 * m);
        en[m] = __ldg(emb + n * D + lane + 32 * m);
    }
#pragma unroll
    for (int m = 0; m < KM; ++m) er[m] = __ldg(rel_emb + r * K + lane + 32 * m);
    transr_project<DM, KM>(W + r * (int64_t)D * K, eh, ep, en, lane, xh, xp, xn);
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
    if (lane == 0) {
        scratch[b] = ns - ps;
        scratch[batch + b] = 0.5f * l2;
    }
}

template <int DM, int KM>
__global__ void __launch_bounds__(128) transr_bwd_kernel(const float* __restrict__ emb, const float* __restrict__ rel_emb,
                                                         const float* __restrict__ W, const int64_t* __restrict__ heads,
                                                         const int64_t* __restrict__ rels, const int64_t* __restrict__ pt,
                                                         const int64_t* __restrict__ nt, int batch, float reg,
                                                         const float* __restrict__ scratch, const float* __restrict__ g_loss,
                                                         float* __restrict__ g_emb, float* __restrict__ g_rel,
                                                         float* __restrict__ g_W) {
    constexpr int D = DM * 32, K = KM * 32;
    const int lane = threadIdx.x & 31;
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (b >= batch) return;
    const int64_t h = heads[b], r = rels[b], p = pt[b], n = nt[b];
    float eh[DM], ep[DM], en[DM], er[KM], xh[KM], xp[KM], xn[KM];
#pragma unroll
    for (int m = 0; m < DM; ++m) {
        eh[m] = __ldg(emb + h * D + lane + 32 * m);
        ep[m] = __ldg(emb + p * D + lane + 32 * m);
        en[m] = __ldg(emb + n * D + lane + 32 * m);
    }
#pragma unroll
    for (int m = 0; m < KM; ++m) er[m] = __ldg(rel_emb + r * K + lane + 32 * m);
    const float* Wr = W + r * (int64_t)D * K;
    transr_project<DM, KM>(Wr, eh, ep, en, lane, xh, xp, xn);

    const float g = g_loss[0] / (float)batch;
    const float s2 = 2.f * sigmoidf_(-scratch[b]) * g;  // loss = -logsigmoid(ns - ps)
    const float lam = reg * g;
    float gxh[KM], gxp[KM], gxn[KM];
#pragma unroll
    for (int m = 0; m < KM; ++m) {
        const float dp = xh[m] + er[m] - xp[m];
        const float dn = xh[m] + er[m] - xn[m];
        const float gdp = s2 * dp;   // dL/d dpos
        const float gdn = -s2 * dn;  // dL/d dneg
        gxh[m] = gdp + gdn + lam * xh[m];
        gxp[m] = -gdp + lam * xp[m];
        gxn[m] = -gdn + lam * xn[m];
        atomicAdd(g_rel + r * K + lane + 32 * m, gdp + gdn + lam * er[m]);
    }
    float geh[DM], gep[DM], gen[DM];
    float* gWr = g_W + r * (int64_t)D * K;
#pragma unroll
    for (int jm = 0; jm < DM; ++jm) {
        geh[jm] = gep[jm] = gen[jm] = 0.f;
#pragma unroll 4
        for (int jj = 0; jj < 32; ++jj) {
            const int j = jm * 32 + jj;
            const float a = __shfl_sync(kFull, eh[jm], jj);
            const float bb = __shfl_sync(kFull, ep[jm], jj);
            const float c = __shfl_sync(kFull, en[jm], jj);
            float ph = 0.f, pp = 0.f, pn = 0.f;
#pragma unroll
            for (int m = 0; m < KM; ++m) {
                const float w = __ldg(Wr + j * K + lane + 32 * m);
                ph = fmaf(gxh[m], w, ph);
                pp = fmaf(gxp[m], w, pp);
                pn = fmaf(gxn[m], w, pn);
                atomicAdd(gWr + j * K + lane + 32 * m, a * gxh[m] + bb * gxp[m] + c * gxn[m]);
            }
            ph = warp_sum(ph);
            pp = warp_sum(pp);
            pn = warp_sum(pn);
            if (lane == jj) {
                geh[jm] = ph;
                gep[jm] = pp;
                gen[jm] = pn;
            }
        }
    }
#pragma unroll
    for (int m = 0; m < DM; ++m) {
        atomicAdd(g_emb + h * D + lane + 32 * m, geh[m]);
        atomicAdd(g_emb + p * D + lane + 32 * m, gep[m]);
        atomicAdd(g_emb + n * D + lane + 32 * m, gen[m]);
    }
}

}  // namespace
}  // namespace kgat

using namespace kgat;

extern "C" {

int kgat_bpr_forward(const kgat_tables_t* tables, const int64_t* users, const int64_t* pos, const int64_t* neg, int32_t batch,
                     float reg, float* loss, float* margin, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    Tables T;
    int rc = pack_tables(tables, &T);
    if (rc != KGAT_OK) return rc;
    if (batch <= 0) return KGAT_ERR_INVALID_ARGUMENT;
    bpr_fwd_kernel<<<(batch * 32 + 255) / 256, 256, 0, stream>>>(T, users, pos, neg, batch, margin);
    loss_reduce_kernel<<<1, 256, 0, stream>>>(margin, batch, reg, loss);
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
    const unsigned blocks = (batch * 32 + 127) / 128;
    KGAT_TRANSR_DISPATCH((transr_fwd_kernel<DM, KM><<<blocks, 128, 0, stream>>>(emb, rel_emb, W, heads, rels, pos_tails, neg_tails,
                                                                               batch, margin)));
    loss_reduce_kernel<<<1, 256, 0, stream>>>(margin, batch, reg, loss);
    return check_launch();
}

int kgat_transr_backward(const float* emb, const float* rel_emb, const float* W, int32_t d, int32_t k, const int64_t* heads,
                         const int64_t* rels, const int64_t* pos_tails, const int64_t* neg_tails, int32_t batch, float reg,
                         const float* margin, const float* g_loss, float* g_emb, float* g_rel_emb, float* g_W, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (batch <= 0) return KGAT_ERR_INVALID_ARGUMENT;
    const unsigned blocks = (batch * 32 + 127) / 128;
    KGAT_TRANSR_DISPATCH((transr_bwd_kernel<DM, KM><<<blocks, 128, 0, stream>>>(emb, rel_emb, W, heads, rels, pos_tails, neg_tails,
                                                                               batch, reg, margin, g_loss, g_emb, g_rel_emb, g_W)));
    return check_launch();
}

}  // extern "C"
