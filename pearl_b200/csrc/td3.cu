// td3.cu — the deterministic actor-critic learners TD3 and DDPG (TD3.learn_batch / ActorCriticBase.learn_batch driven by
// PolicyLearner.learn), replacing
//   policy_learners/sequential_decision_making/td3.py:106-202         delayed actor + target updates, clipped target noise
//   policy_learners/sequential_decision_making/ddpg.py:105-157        actor loss -mean Q1(s, pi(s)), twin critic loss
//   policy_learners/sequential_decision_making/actor_critic_base.py:309-366  step order, soft target updates
//   neural_networks/sequential_decision_making/actor_networks.py:29-51,448-485  VanillaContinuousActorNetwork, action_scaling
//   neural_networks/sequential_decision_making/twin_critic.py:75-91, utils/functional_utils/learning/critic_utils.py:103-122,170-203
// DDPG is the same step with actor_update_freq = 1 and no target noise (this reference trains a twin critic for DDPG too).
// Same launch structure as the SAC learner (sac.cu): one round = a fixed sequence of launches of the tiled contraction
// kernel (gemm.cuh) plus small elementwise kernels, everything round-dependent read on the device through a per-call
// block; two CUDA graphs (round with / without the actor update) are captured once and replayed.
#include <math.h>
#include <stdarg.h>

#include <new>

#include "common.cuh"
#include "gemm.cuh"

using namespace prl;

namespace {

struct Td3Call {
    const float *noise;      // [rounds][B][A] target-policy noise (torch.normal draws), or null (DDPG)
    const int32_t *slots;    // [rounds][B]
    float *out_actor, *out_critic;
};

__global__ void k_td3_gather(const uint32_t *__restrict__ records, prl_buf_layout L, int obs, int act, const Td3Call *__restrict__ call,
                             const int *__restrict__ round_idx, int B, float *__restrict__ S, float *__restrict__ A, float *__restrict__ R,
                             float *__restrict__ S2, float *__restrict__ T) {
    const int lane = threadIdx.x & 31, w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= B) return;
    const int32_t *slots = call->slots + (size_t)(*round_idx) * B;
    const uint32_t *r = records + (size_t)slots[w] * L.record_words;
    for (int p = lane; p < obs; p += 32) {
        S[(size_t)w * obs + p] = __uint_as_float(r[L.off_state + p]);
        S2[(size_t)w * obs + p] = __uint_as_float(r[L.off_next_state + p]);
    }
    for (int p = lane; p < act; p += 32) A[(size_t)w * act + p] = __uint_as_float(r[L.off_action + p]);
    if (lane == 0) { R[w] = __uint_as_float(r[L.off_reward]); T[w] = (r[L.off_flags] & 1u) ? 1.f : 0.f; }
}

// VanillaContinuousActorNetwork.sample_action: tanh head, action_scaling (actor_networks.py:29-51,475-485);
// target policy (td3.py:150-175): + clamp(noise, +-clip) * (high - low) / 2, clamped to the box
__global__ void k_td3_act(int B, int A, const float *__restrict__ pre, const float *__restrict__ low, const float *__restrict__ high,
                          const Td3Call *__restrict__ call, const int *__restrict__ round_idx, int with_noise, float clip,
                          float *__restrict__ action, float *__restrict__ na_out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= B * A) return;
    const int d = e % A;
    const float lo = low[d], hi = high[d];
    const float na = tanhf(pre[e]);
    float a = (((hi - lo) * (na + 1.0f)) / 2.f) + lo;
    if (with_noise && call->noise) {
        float nz = call->noise[(size_t)(*round_idx) * B * A + e];
        nz = fminf(fmaxf(nz, -clip), clip) * (hi - lo) / 2.f;
        a = fminf(fmaxf(a + nz, lo), hi);
    }
    action[e] = a;
    if (na_out) na_out[e] = na;
}

// actor loss = -mean(q1); dq1 = -1/B
__global__ void k_td3_actor_loss(int B, const float *__restrict__ q1, float *__restrict__ dq, const Td3Call *__restrict__ call,
                                 const int *__restrict__ round_idx, float *__restrict__ last_actor_loss) {
    __shared__ float red[256];
    float s = 0.f;
    const float ib = 1.f / (float)B;
    for (int b = threadIdx.x; b < B; b += blockDim.x) { s -= q1[b]; dq[b] = -ib; }
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) { *last_actor_loss = red[0] * ib; call->out_actor[*round_idx] = red[0] * ib; }
}
// rounds without an actor update report the last actor loss again (td3.py:128)
__global__ void k_td3_repeat_actor_loss(const Td3Call *__restrict__ call, const int *__restrict__ round_idx, const float *__restrict__ last) {
    call->out_actor[*round_idx] = *last;
}
// d(pre) = d(action) * (high - low) / 2 * (1 - tanh^2)
__global__ void k_td3_head_grad(int B, int A, const float *__restrict__ da, const float *__restrict__ na, const float *__restrict__ low,
                                const float *__restrict__ high, float *__restrict__ dpre) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= B * A) return;
    const int d = e % A;
    const float n = na[e];
    dpre[e] = da[e] * ((high[d] - low[d]) * 0.5f) * (1.f - n * n);
}
// y = min(q1t, q2t) * gamma * (1 - terminated) + reward
__global__ void k_td3_target(int B, const float *__restrict__ qt, float gamma, const float *__restrict__ term, const float *__restrict__ rew,
                             float *__restrict__ y) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    y[b] = __fadd_rn(__fmul_rn(__fmul_rn(fminf(qt[b], qt[B + b]), gamma), 1.f - term[b]), rew[b]);
}
__global__ void k_td3_critic_loss(int B, const float *__restrict__ q, const float *__restrict__ y, float *__restrict__ dq,
                                  const Td3Call *__restrict__ call, const int *__restrict__ round_idx) {
    __shared__ float red[256];
    float s = 0.f;
    const float ib = 1.f / (float)B;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        const float e1 = q[b] - y[b], e2 = q[B + b] - y[b];
        s += e1 * e1 + e2 * e2;
        dq[b] = e1 * ib;            // d/dq1 of (mse1 + mse2) / 2
        dq[B + b] = e2 * ib;
    }
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) call->out_critic[*round_idx] = red[0] * ib * 0.5f;
}
__global__ void k_td3_soft_update(int n, float *__restrict__ target, const float *__restrict__ src, float tau, float omtau) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) target[i] = __fadd_rn(__fmul_rn(tau, src[i]), __fmul_rn(omtau, target[i]));
}
__global__ void k_td3_bump(int *round_idx, int *actor_round_idx, int actor_updated) {
    *round_idx += 1;
    if (actor_updated) *actor_round_idx += 1;
}

}  // namespace

struct prl_td3 {
    prl_td3_cfg cfg;
    int Pa, Pc;                        // actor parameters; parameters of ONE critic
    int aW1, ab1, aW2, ab2, aW3, ab3;
    int cW1, cb1, cW2, cb2, cW3, cb3;
    float *actor, *actor_m, *actor_v, *actor_x, *actor_t;
    float *critic, *critic_m, *critic_v, *critic_x, *critic_t;
    const float *low, *high;
    int64_t actor_step, critic_step;
    float *S, *A, *R, *S2, *T, *h1, *h2, *pre, *act_s, *na, *c1, *c2, *q, *qt, *dq, *dc2, *dc1, *da, *dpre, *dh2, *dh1, *y, *g_actor,
        *g_critic, *last_actor_loss;
    int32_t *slots, *logical;
    float2 *scal_a, *scal_c;
    Td3Call *call;
    int *round_idx, *actor_round_idx;
    bool use_graph;
    cudaGraphExec_t graph_exec[2];     // [0] round without, [1] with the actor update
    int graph_batch;
    const uint32_t *graph_buf;
    int launches_per_round[2];
    float2 *scal_host[2];
    cudaEvent_t scal_done[2];
    int scal_next;
    int64_t last_launches;
};

static int64_t al64_(int64_t x) { return (x + 255) / 256 * 256; }

static void td3_layout(prl_td3 *s) {
    const prl_td3_cfg &c = s->cfg;
    int o = 0;
    s->aW1 = o; o += c.actor_h1 * c.obs_dim; s->ab1 = o; o += c.actor_h1;
    s->aW2 = o; o += c.actor_h2 * c.actor_h1; s->ab2 = o; o += c.actor_h2;
    s->aW3 = o; o += c.act_dim * c.actor_h2; s->ab3 = o; o += c.act_dim;
    s->Pa = o;
    const int D = c.obs_dim + c.act_dim;
    o = 0;
    s->cW1 = o; o += c.critic_h1 * D; s->cb1 = o; o += c.critic_h1;
    s->cW2 = o; o += c.critic_h2 * c.critic_h1; s->cb2 = o; o += c.critic_h2;
    s->cW3 = o; o += c.critic_h2; s->cb3 = o; o += 1;
    s->Pc = o;
}
static int td3_check(const prl_td3_cfg *c) {
    PRL_REQUIRE(c, "null cfg");
    PRL_REQUIRE(c->obs_dim > 0 && c->act_dim > 0 && c->actor_h1 > 0 && c->actor_h2 > 0 && c->critic_h1 > 0 && c->critic_h2 > 0,
                "dimensions must be positive");
    PRL_REQUIRE(c->max_batch > 0 && c->max_rounds > 0 && c->actor_update_freq >= 1, "max_batch / max_rounds / actor_update_freq must be positive");
    return PRL_OK;
}
extern "C" int64_t prl_td3_actor_param_count(const prl_td3_cfg *c) {
    if (td3_check(c)) return -1;
    prl_td3 t; t.cfg = *c; td3_layout(&t);
    return t.Pa;
}
extern "C" int64_t prl_td3_critic_param_count(const prl_td3_cfg *c) {   // ONE critic; the twin vector holds two
    if (td3_check(c)) return -1;
    prl_td3 t; t.cfg = *c; td3_layout(&t);
    return t.Pc;
}
struct Td3Ws { int64_t off[40]; int64_t total; };
static Td3Ws td3_ws(const prl_td3_cfg *c, int Pa, int Pc) {
    Td3Ws w; int64_t o = 0; int k = 0;
    const int64_t B = c->max_batch, A = c->act_dim, O = c->obs_dim;
    auto add = [&](int64_t floats) { w.off[k++] = o; o = al64_(o + floats * 4); };
    add(B * O); add(B * A); add(B); add(B * O); add(B);                                  // S A R S2 T
    add(B * c->actor_h1); add(B * c->actor_h2); add(B * A); add(B * A); add(B * A);       // h1 h2 pre act_s na
    add(2 * B * c->critic_h1); add(2 * B * c->critic_h2); add(2 * B); add(2 * B);         // c1 c2 q qt
    add(2 * B); add(2 * B * c->critic_h2); add(2 * B * c->critic_h1); add(2 * B * A);     // dq dc2 dc1 da
    add(B * A); add(B * c->actor_h2); add(B * c->actor_h1); add(B);                      // dpre dh2 dh1 y
    add(Pa); add(2 * (int64_t)Pc); add(4);                                                // g_actor g_critic last_actor_loss
    add((int64_t)c->max_rounds * B); add((int64_t)c->max_rounds * B);                    // slots logical (int32)
    add(4 * (int64_t)c->max_rounds + 64);                                                 // scal_a | scal_c | call | round_idx | actor_round_idx
    w.total = o;
    return w;
}
extern "C" int64_t prl_td3_workspace_bytes(const prl_td3_cfg *c) {
    if (td3_check(c)) return -1;
    prl_td3 t; t.cfg = *c; td3_layout(&t);
    return td3_ws(c, t.Pa, t.Pc).total;
}

extern "C" int prl_td3_create(prl_td3 **out, const prl_td3_cfg *cfg, float *actor_w, float *actor_m, float *actor_v, float *actor_vmax,
                              float *actor_target_w, float *critic_w, float *critic_m, float *critic_v, float *critic_vmax,
                              float *critic_target_w, const float *low_dev, const float *high_dev, int64_t actor_adam_step,
                              int64_t critic_adam_step, void *workspace) {
    PRL_REQUIRE(out && actor_w && actor_m && actor_v && actor_vmax && actor_target_w && critic_w && critic_m && critic_v && critic_vmax &&
                    critic_target_w && low_dev && high_dev && workspace, "null argument");
    int rc = td3_check(cfg);
    if (rc) return rc;
    prl_td3 *s = new (std::nothrow) prl_td3();
    if (!s) return fail(PRL_ENOMEM, "out of host memory");
    s->cfg = *cfg;
    td3_layout(s);
    s->actor = actor_w; s->actor_m = actor_m; s->actor_v = actor_v; s->actor_x = actor_vmax; s->actor_t = actor_target_w;
    s->critic = critic_w; s->critic_m = critic_m; s->critic_v = critic_v; s->critic_x = critic_vmax; s->critic_t = critic_target_w;
    s->low = low_dev; s->high = high_dev;
    s->actor_step = actor_adam_step; s->critic_step = critic_adam_step;
    Td3Ws w = td3_ws(cfg, s->Pa, s->Pc);
    char *b = (char *)workspace;
    float **f[] = {&s->S, &s->A, &s->R, &s->S2, &s->T, &s->h1, &s->h2, &s->pre, &s->act_s, &s->na, &s->c1, &s->c2, &s->q, &s->qt, &s->dq,
                   &s->dc2, &s->dc1, &s->da, &s->dpre, &s->dh2, &s->dh1, &s->y, &s->g_actor, &s->g_critic, &s->last_actor_loss};
    int k = 0;
    for (auto p : f) *p = (float *)(b + w.off[k++]);
    s->slots = (int32_t *)(b + w.off[k++]); s->logical = (int32_t *)(b + w.off[k++]);
    s->scal_a = (float2 *)(b + w.off[k++]); s->scal_c = s->scal_a + cfg->max_rounds;
    s->call = (Td3Call *)(s->scal_c + cfg->max_rounds); s->round_idx = (int *)(s->call + 1); s->actor_round_idx = s->round_idx + 1;
    static_assert(sizeof(Td3Call) + 8 <= 64 * 4, "call block fits the reserved tail");
    s->scal_next = 0; s->use_graph = true; s->graph_exec[0] = s->graph_exec[1] = nullptr; s->graph_batch = 0; s->graph_buf = nullptr;
    s->last_launches = 0;
    cudaError_t e = cudaMemset(s->last_actor_loss, 0, 16);
    for (int i = 0; i < 2 && e == cudaSuccess; i++) {
        e = cudaHostAlloc((void **)&s->scal_host[i], (size_t)cfg->max_rounds * 16 + 256, cudaHostAllocDefault);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->scal_done[i], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) { delete s; return fail(PRL_ECUDA, "prl_td3_create: %s", cudaGetErrorString(e)); }
    *out = s;
    return PRL_OK;
}
extern "C" int prl_td3_destroy(prl_td3 *s) {
    if (!s) return PRL_OK;
    for (int i = 0; i < 2; i++) { cudaEventSynchronize(s->scal_done[i]); cudaEventDestroy(s->scal_done[i]); cudaFreeHost(s->scal_host[i]); }
    for (int i = 0; i < 2; i++) if (s->graph_exec[i]) cudaGraphExecDestroy(s->graph_exec[i]);
    delete s;
    return PRL_OK;
}
extern "C" int64_t prl_td3_actor_adam_step(const prl_td3 *s) { return s ? s->actor_step : -1; }
extern "C" int64_t prl_td3_critic_adam_step(const prl_td3 *s) { return s ? s->critic_step : -1; }

// one learner round, launched (or captured) on `st`
static int td3_round(prl_td3 *s, prl_buf *buf, int B, bool update_actor, cudaStream_t st) {
    const prl_td3_cfg &c = s->cfg;
    const int O = c.obs_dim, A = c.act_dim, D = O + A;
    const int H1 = c.actor_h1, H2 = c.actor_h2, C1 = c.critic_h1, C2 = c.critic_h2;
    const long long Pc = s->Pc;
    AdamHp ha = adam_hp(c.actor_lr, c.beta1, c.beta2, c.eps, c.weight_decay), hc = adam_hp(c.critic_lr, c.beta1, c.beta2, c.eps, c.weight_decay);
    GemmLauncher L; L.st = st;
    const float *cw = s->critic, *ct = s->critic_t;
    const long long sC1 = (long long)B * C1, sC2 = (long long)B * C2;
    auto actor_forward = [&](const float *net, const float *X) {
        L.fwd(mat(X, O), B, net + s->aW1, O, 0, net + s->ab1, 0, H1, O, true, s->h1, H1, 0);
        L.fwd(mat(s->h1, H1), B, net + s->aW2, H1, 0, net + s->ab2, 0, H2, H1, true, s->h2, H2, 0);
        L.fwd(mat(s->h2, H2), B, net + s->aW3, H2, 0, net + s->ab3, 0, A, H2, false, s->pre, A, 0);
    };
    auto critic_forward = [&](const float *net, const float *X, const float *Act, float *qout, int nets) {
        L.fwd(mat2(X, O, O, Act, A), B, net + s->cW1, D, Pc, net + s->cb1, Pc, C1, D, true, s->c1, C1, sC1, nets);
        L.fwd(mat(s->c1, C1, sC1), B, net + s->cW2, C1, Pc, net + s->cb2, Pc, C2, C1, true, s->c2, C2, sC2, nets);
        L.fwd(mat(s->c2, C2, sC2), B, net + s->cW3, C2, Pc, net + s->cb3, Pc, 1, C2, false, qout, 1, B, nets);
    };
    const int eb = 256;
    int small = 0;
    k_td3_gather<<<(B * 32 + eb - 1) / eb, eb, 0, st>>>(buf->records, buf->lay, O, A, s->call, s->round_idx, B, s->S, s->A, s->R, s->S2, s->T);
    small++;
    if (update_actor) {
        // ---------------- actor step: maximise Q1(s, pi(s))   (ddpg.py:105-121)
        actor_forward(s->actor, s->S);
        k_td3_act<<<(B * A + eb - 1) / eb, eb, 0, st>>>(B, A, s->pre, s->low, s->high, s->call, s->round_idx, 0, 0.f, s->act_s, s->na);
        critic_forward(cw, s->S, s->act_s, s->q, 1);
        k_td3_actor_loss<<<1, 256, 0, st>>>(B, s->q, s->dq, s->call, s->round_idx, s->last_actor_loss);
        dim3 g1((B * C2 + eb - 1) / eb, 1, 1);
        k_head_bwd<<<g1, eb, 0, st>>>(B, C2, s->dq, cw + s->cW3, Pc, s->c2, s->dc2);
        L.bwd_x(s->dc2, C2, sC2, B, C2, cw + s->cW2, C1, Pc, 0, C1, s->dc1, C1, sC1, s->c1, C1, sC1, false, 1);
        L.bwd_x(s->dc1, C1, sC1, B, C1, cw + s->cW1, D, Pc, O, A, s->da, A, (long long)B * A, nullptr, 0, 0, false, 1);
        k_td3_head_grad<<<(B * A + eb - 1) / eb, eb, 0, st>>>(B, A, s->da, s->na, s->low, s->high, s->dpre);
        float *ga = s->g_actor;
        const float *aw = s->actor;
        L.bwd_w(s->dpre, A, 0, B, A, mat(s->h2, H2), H2, ga + s->aW3, H2, 0, ga + s->ab3, 0);
        L.bwd_x(s->dpre, A, 0, B, A, aw + s->aW3, H2, 0, 0, H2, s->dh2, H2, 0, s->h2, H2, 0, false);
        L.bwd_w(s->dh2, H2, 0, B, H2, mat(s->h1, H1), H1, ga + s->aW2, H1, 0, ga + s->ab2, 0);
        L.bwd_x(s->dh2, H2, 0, B, H2, aw + s->aW2, H1, 0, 0, H1, s->dh1, H1, 0, s->h1, H1, 0, false);
        L.bwd_w(s->dh1, H1, 0, B, H1, mat(s->S, O), O, ga + s->aW1, O, 0, ga + s->ab1, 0);
        k_adamw<<<(s->Pa + eb - 1) / eb, eb, 0, st>>>(s->Pa, s->actor, s->actor_m, s->actor_v, s->actor_x, ga, ha, s->scal_a, s->actor_round_idx,
                                                     nullptr, 0.f, 0.f);
        small += 5;
    } else {
        k_td3_repeat_actor_loss<<<1, 1, 0, st>>>(s->call, s->round_idx, s->last_actor_loss);
        small++;
    }
    // ---------------- critic step (td3.py:150-202 / ddpg.py:123-157): target action from the TARGET actor (+ clipped noise)
    actor_forward(s->actor_t, s->S2);
    k_td3_act<<<(B * A + eb - 1) / eb, eb, 0, st>>>(B, A, s->pre, s->low, s->high, s->call, s->round_idx, 1, (float)c.noise_clip, s->act_s, nullptr);
    critic_forward(ct, s->S2, s->act_s, s->qt, 2);
    k_td3_target<<<(B + eb - 1) / eb, eb, 0, st>>>(B, s->qt, (float)c.gamma, s->T, s->R, s->y);
    critic_forward(cw, s->S, s->A, s->q, 2);
    k_td3_critic_loss<<<1, 256, 0, st>>>(B, s->q, s->y, s->dq, s->call, s->round_idx);
    {
        float *gc = s->g_critic;
        L.bwd_w(s->dq, 1, B, B, 1, mat(s->c2, C2, sC2), C2, gc + s->cW3, C2, Pc, gc + s->cb3, Pc, 2);
        dim3 g2((B * C2 + eb - 1) / eb, 1, 2);
        k_head_bwd<<<g2, eb, 0, st>>>(B, C2, s->dq, cw + s->cW3, Pc, s->c2, s->dc2);
        L.bwd_w(s->dc2, C2, sC2, B, C2, mat(s->c1, C1, sC1), C1, gc + s->cW2, C1, Pc, gc + s->cb2, Pc, 2);
        L.bwd_x(s->dc2, C2, sC2, B, C2, cw + s->cW2, C1, Pc, 0, C1, s->dc1, C1, sC1, s->c1, C1, sC1, false, 2);
        L.bwd_w(s->dc1, C1, sC1, B, C1, mat2(s->S, O, O, s->A, A), D, gc + s->cW1, D, Pc, gc + s->cb1, Pc, 2);
        const int n2p = 2 * s->Pc;
        // the critic targets follow only on rounds with an actor update (td3.py:136-147); DDPG: every round
        k_adamw<<<(n2p + eb - 1) / eb, eb, 0, st>>>(n2p, s->critic, s->critic_m, s->critic_v, s->critic_x, gc, hc, s->scal_c, s->round_idx,
                                                  update_actor ? s->critic_t : nullptr, (float)c.critic_tau, (float)(1.0 - c.critic_tau));
    }
    small += 6;
    if (update_actor) {
        k_td3_soft_update<<<(s->Pa + eb - 1) / eb, eb, 0, st>>>(s->Pa, s->actor_t, s->actor, (float)c.actor_tau, (float)(1.0 - c.actor_tau));
        small++;
    }
    k_td3_bump<<<1, 1, 0, st>>>(s->round_idx, s->actor_round_idx, update_actor ? 1 : 0);
    small++;
    s->launches_per_round[update_actor ? 1 : 0] = L.count + small;
    return PRL_OK;
}

extern "C" int prl_td3_learn(prl_td3 *s, prl_buf *buf, int rounds, int batch, int64_t training_steps0, const float *noise_dev,
                             float *out_actor_loss, float *out_critic_loss, int32_t *out_logical, void *stream_) {
    PRL_REQUIRE(s && buf && out_actor_loss && out_critic_loss, "null argument");
    const prl_td3_cfg &c = s->cfg;
    PRL_REQUIRE(rounds > 0 && rounds <= c.max_rounds && batch > 0 && batch <= c.max_batch, "rounds / batch outside the configured maxima");
    PRL_REQUIRE((buf->desc.flags & PRL_BUF_CONTINUOUS) && buf->desc.obs_dim == c.obs_dim && buf->desc.act_dim == c.act_dim,
                "TD3 / DDPG need a continuous-action buffer with matching dimensions");
    cudaStream_t st = (cudaStream_t)stream_;
    int rc = prl_buf_sample_indices(buf, rounds, batch, out_logical ? out_logical : s->logical, s->slots, stream_);
    if (rc) return rc;
    // which rounds update the actor: PolicyLearner.learn increments _training_steps before learn_batch (policy_learner.py:183),
    // TD3 tests `_training_steps % actor_update_freq == 0` (td3.py:121)
    auto updates = [&](int r) { return c.actor_update_freq <= 1 || (training_steps0 + r + 1) % c.actor_update_freq == 0; };
    const int sb = s->scal_next; s->scal_next ^= 1;
    PRL_CUDA(cudaEventSynchronize(s->scal_done[sb]));
    float2 *hs = s->scal_host[sb];
    int n_actor = 0;
    for (int r = 0; r < rounds; r++) {
        const double cstep = (double)(s->critic_step + r + 1);
        hs[c.max_rounds + r] = make_float2((float)(c.critic_lr / (1.0 - pow(c.beta1, cstep))), (float)sqrt(1.0 - pow(c.beta2, cstep)));
        if (updates(r)) {     // the actor optimizer's own step count: it only advances on update rounds
            const double astep = (double)(s->actor_step + n_actor + 1);
            hs[n_actor] = make_float2((float)(c.actor_lr / (1.0 - pow(c.beta1, astep))), (float)sqrt(1.0 - pow(c.beta2, astep)));
            n_actor++;
        }
    }
    Td3Call *hc = reinterpret_cast<Td3Call *>(hs + 2 * (size_t)c.max_rounds);
    hc->noise = noise_dev; hc->slots = s->slots; hc->out_actor = out_actor_loss; hc->out_critic = out_critic_loss;
    int *hround = reinterpret_cast<int *>(hc + 1);
    hround[0] = 0; hround[1] = 0;
    PRL_CUDA(cudaMemcpyAsync(s->scal_a, hs, 2 * (size_t)c.max_rounds * 8 + sizeof(Td3Call) + 8, cudaMemcpyHostToDevice, st));
    PRL_CUDA(cudaEventRecord(s->scal_done[sb], st));

    if (s->use_graph) {
        if (!s->graph_exec[0] || s->graph_batch != batch || s->graph_buf != buf->records) {
            for (int u = 0; u < 2; u++) {
                if (s->graph_exec[u]) { cudaGraphExecDestroy(s->graph_exec[u]); s->graph_exec[u] = nullptr; }
                cudaStream_t cs;
                PRL_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
                cudaGraph_t graph = nullptr;
                cudaError_t e = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal);
                if (e == cudaSuccess) {
                    td3_round(s, buf, batch, u == 1, cs);
                    e = cudaStreamEndCapture(cs, &graph);
                }
                if (e == cudaSuccess) e = cudaGraphInstantiate(&s->graph_exec[u], graph, 0);
                if (graph) cudaGraphDestroy(graph);
                cudaStreamDestroy(cs);
                if (e != cudaSuccess) { s->graph_exec[u] = nullptr; return fail(PRL_ECUDA, "prl_td3_learn: graph capture failed: %s", cudaGetErrorString(e)); }
            }
            s->graph_batch = batch; s->graph_buf = buf->records;
        }
        for (int r = 0; r < rounds; r++) PRL_CUDA(cudaGraphLaunch(s->graph_exec[updates(r) ? 1 : 0], st));
    } else {
        for (int r = 0; r < rounds; r++) {
            rc = td3_round(s, buf, batch, updates(r), st);
            if (rc) return rc;
        }
    }
    PRL_CUDA(cudaGetLastError());
    s->critic_step += rounds;
    s->actor_step += n_actor;
    s->last_launches = 0;
    for (int r = 0; r < rounds; r++) s->last_launches += s->launches_per_round[updates(r) ? 1 : 0];
    return PRL_OK;
}
extern "C" int prl_td3_set_graph(prl_td3 *s, int enable) {
    PRL_REQUIRE(s, "null handle");
    s->use_graph = enable != 0;
    return PRL_OK;
}
extern "C" int64_t prl_td3_last_launches(const prl_td3 *s) { return s ? s->last_launches : -1; }
