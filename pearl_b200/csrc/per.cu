// per.cu — K8: proportional prioritized replay (sum tree + min tree in HBM, stratified sampling,
// importance weights, priority updates).  The reference has NO prioritized replay; the
// specification is oracle/per_oracle.py (Schaul et al. 2016), restated there and mirrored here:
//   leaves = physical ring slots; node i = node 2i (+ | min) node 2i+1 in fp32, fixed pairwise order, so the
//   device tree is bit-identical to the oracle's given the same leaves; draws u_k = (k + U_k) * (total / B)
//   with U_k from Philox4x32-10(counter = (k, step), key = seed); w_i = (p_min / p_i)^beta.
// HBM-bound integer/float work: a sampled batch touches B * log2(C) nodes (top 11 levels from shared
// memory), an update rewrites B root paths level by level.
#include <math.h>
#include <stdarg.h>

#include <new>

#include "common.cuh"

using namespace prl;

struct prl_per {
    prl_per_cfg cfg;
    int64_t C2;          // leaves (power of two >= capacity)
    int levels;          // log2(C2)
    float *sum, *mn;     // device, 2*C2 each (node 1 = root, leaves at [C2, 2*C2))
    float *max_priority; // device scalar
    uint64_t draws;      // sample() calls so far = Philox step
};

namespace {

constexpr int kTopLevels = 11;            // nodes 1 .. 2^11-1 cached in shared memory by the sampler
constexpr int kTopNodes = 1 << kTopLevels;

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t &o0) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    o0 = c0;
}

__global__ void k_per_init(float *sum, float *mn, int64_t n2, float *max_priority) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n2) { sum[i] = 0.f; mn[i] = INFINITY; }
    if (i == 0) *max_priority = 1.f;
}

// leaves [first, first + count) <- *max_priority   (contiguous, no wrap)
__global__ void k_per_fill_leaves(float *sum, float *mn, int64_t C2, int64_t first, int64_t count, const float *max_priority) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= count) return;
    const float p = *max_priority;
    sum[C2 + first + i] = p;
    mn[C2 + first + i] = p;
}
// recompute the parents [lo, hi] of one level from their children
__global__ void k_per_fix_level(float *sum, float *mn, int64_t lo, int64_t hi) {
    const int64_t i = lo + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i > hi) return;
    sum[i] = __fadd_rn(sum[2 * i], sum[2 * i + 1]);
    mn[i] = fminf(mn[2 * i], mn[2 * i + 1]);
}

// priorities of k sampled slots from their TD errors, then the k root paths level by level (one CTA)
__global__ void __launch_bounds__(1024, 1)
k_per_update(float *sum, float *mn, int64_t C2, int levels, const int32_t *__restrict__ slots, const float *__restrict__ td,
             int k, float alpha, float eps, float *max_priority, float *out_priority) {
    const int t = threadIdx.x;
    int64_t idx = 0;
    if (t < k) {
        const float p = powf(fabsf(td[t]) + eps, alpha);
        idx = C2 + slots[t];
        volatile float *vs = sum, *vm = mn;
        vs[idx] = p; vm[idx] = p;
        atomicMax(reinterpret_cast<int *>(max_priority), __float_as_int(p));   // p > 0: int order == float order
        if (out_priority) out_priority[t] = p;
    }
    for (int l = 0; l < levels; l++) {
        __syncthreads();
        if (t < k) {
            idx >>= 1;
            volatile float *vs = sum, *vm = mn;
            vs[idx] = __fadd_rn(vs[2 * idx], vs[2 * idx + 1]);   // duplicates of a parent write the same value
            vm[idx] = fminf(vm[2 * idx], vm[2 * idx + 1]);
        }
    }
}

__global__ void __launch_bounds__(1024, 1)
k_per_sample(const float *__restrict__ sum, const float *__restrict__ mn, int64_t C2, int levels, int k, uint64_t step,
             uint32_t key0, uint32_t key1, float beta, int32_t *__restrict__ out_slots, float *__restrict__ out_w) {
    __shared__ float top[kTopNodes];
    const int t = threadIdx.x;
    const int cached = (int)((int64_t)kTopNodes < 2 * C2 ? kTopNodes : 2 * C2);
    for (int i = t; i < cached; i += blockDim.x) top[i] = sum[i];
    __syncthreads();
    if (t >= k) return;
    const float total = top[1];
    const float seg = __fdiv_rn(total, (float)k);
    uint32_t x0;
    philox4x32_10((uint32_t)t, (uint32_t)step, (uint32_t)(step >> 32), 0u, key0, key1, x0);
    const float U = __fmul_rn((float)(x0 >> 8), 5.9604644775390625e-08f);   // 2^-24
    float u = __fmul_rn(__fadd_rn((float)t, U), seg);
    int64_t idx = 1;
    for (int l = 0; l < levels; l++) {
        const int64_t c = 2 * idx;
        const float left = c < cached ? top[c] : __ldg(sum + c);
        const float right = c + 1 < cached ? top[c + 1] : __ldg(sum + c + 1);
        if (u < left || right == 0.f) idx = c;
        else { u = __fsub_rn(u, left); idx = c + 1; }
    }
    const float p = __ldg(sum + idx);
    out_slots[t] = (int32_t)(idx - C2);
    out_w[t] = powf(__fdiv_rn(mn[1], p), beta);
}

}  // namespace

extern "C" int64_t prl_per_tree_floats(int64_t capacity) {
    if (capacity <= 0) return -1;
    int64_t c2 = 1;
    while (c2 < capacity) c2 <<= 1;
    return 2 * c2;
}

extern "C" int prl_per_create(prl_per **out, const prl_per_cfg *cfg, float *sum_tree_dev, float *min_tree_dev,
                              float *max_priority_dev, void *stream) {
    PRL_REQUIRE(out && cfg && sum_tree_dev && min_tree_dev && max_priority_dev, "null argument");
    PRL_REQUIRE(cfg->capacity > 0 && cfg->capacity < (1ll << 30), "capacity out of range");
    PRL_REQUIRE(cfg->alpha >= 0 && cfg->beta >= 0 && cfg->eps > 0, "alpha, beta >= 0 and eps > 0 required");
    prl_per *p = new (std::nothrow) prl_per();
    if (!p) return fail(PRL_ENOMEM, "out of host memory");
    p->cfg = *cfg;
    p->C2 = 1; p->levels = 0;
    while (p->C2 < cfg->capacity) { p->C2 <<= 1; p->levels++; }
    p->sum = sum_tree_dev; p->mn = min_tree_dev; p->max_priority = max_priority_dev;
    p->draws = 0;
    const int64_t n2 = 2 * p->C2;
    k_per_init<<<(unsigned)((n2 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p->sum, p->mn, n2, p->max_priority);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { delete p; return fail(PRL_ECUDA, "prl_per_create: %s", cudaGetErrorString(e)); }
    *out = p;
    return PRL_OK;
}
extern "C" int prl_per_destroy(prl_per *p) { delete p; return PRL_OK; }
extern "C" int prl_per_set_beta(prl_per *p, double beta) {
    PRL_REQUIRE(p && beta >= 0, "bad argument");
    p->cfg.beta = beta;
    return PRL_OK;
}
extern "C" int64_t prl_per_draws(const prl_per *p) { return p ? (int64_t)p->draws : -1; }

static int fix_range(prl_per *p, int64_t first, int64_t count, cudaStream_t stream) {
    int64_t lo = p->C2 + first, hi = p->C2 + first + count - 1;
    for (int l = 0; l < p->levels; l++) {
        lo >>= 1; hi >>= 1;
        const int64_t n = hi - lo + 1;
        k_per_fix_level<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(p->sum, p->mn, lo, hi);
    }
    PRL_CUDA(cudaGetLastError());
    return PRL_OK;
}

extern "C" int prl_per_push(prl_per *p, int64_t first_slot, int64_t count, void *stream_) {
    PRL_REQUIRE(p, "null handle");
    PRL_REQUIRE(first_slot >= 0 && first_slot < p->cfg.capacity && count >= 0, "bad slot range");
    cudaStream_t stream = (cudaStream_t)stream_;
    if (count > p->cfg.capacity) { first_slot = (first_slot + count) % p->cfg.capacity; count = p->cfg.capacity; }
    while (count > 0) {   // at most two contiguous pieces (ring wrap)
        const int64_t m = count < p->cfg.capacity - first_slot ? count : p->cfg.capacity - first_slot;
        k_per_fill_leaves<<<(unsigned)((m + 255) / 256), 256, 0, stream>>>(p->sum, p->mn, p->C2, first_slot, m, p->max_priority);
        int rc = fix_range(p, first_slot, m, stream);
        if (rc) return rc;
        count -= m;
        first_slot = 0;
    }
    return PRL_OK;
}

extern "C" int prl_per_set_priorities(prl_per *p, const int32_t *slots_dev, const float *td_dev, int k,
                                      float *out_priority_dev, void *stream) {
    PRL_REQUIRE(p && slots_dev && td_dev, "null argument");
    PRL_REQUIRE(k > 0 && k <= 1024, "1 <= k <= 1024 priorities per update");
    k_per_update<<<1, 1024, 0, (cudaStream_t)stream>>>(p->sum, p->mn, p->C2, p->levels, slots_dev, td_dev, k,
                                                        (float)p->cfg.alpha, (float)p->cfg.eps, p->max_priority,
                                                        out_priority_dev);
    PRL_CUDA(cudaGetLastError());
    return PRL_OK;
}

extern "C" int prl_per_sample(prl_per *p, int k, int32_t *out_slots_dev, float *out_weights_dev, void *stream) {
    PRL_REQUIRE(p && out_slots_dev && out_weights_dev, "null argument");
    PRL_REQUIRE(k > 0 && k <= 1024, "1 <= k <= 1024 draws per sample");
    k_per_sample<<<1, 1024, 0, (cudaStream_t)stream>>>(p->sum, p->mn, p->C2, p->levels, k, p->draws,
                                                        (uint32_t)(p->cfg.seed & 0xffffffffu), (uint32_t)(p->cfg.seed >> 32),
                                                        (float)p->cfg.beta, out_slots_dev, out_weights_dev);
    PRL_CUDA(cudaGetLastError());
    p->draws++;
    return PRL_OK;
}
