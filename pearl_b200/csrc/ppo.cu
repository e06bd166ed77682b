// ppo.cu — K7: generalized advantage estimation + truncated lambda returns over a rollout.
//
// Replaces the per-transition Python loop of ProximalPolicyOptimization.preprocess_replay_buffer
// (pearl/policy_learners/sequential_decision_making/ppo.py:271-293), which walks the stored
// transitions newest -> oldest:
//     td    = reward + gamma * next_value * (~terminated) - V[i]
//     gae   = td + gamma * lambda * (not (terminated or truncated)) * gae
//     lam_return = gae + V[i];   next_value = V[i]
// (next_value starts as the critic's value of the LAST stored next_state; note that V(s_{t+1}) is
// taken to be the value of the next STORED transition's state — a reference quirk that is mirrored.)
//
// td depends only on the transition and its newer neighbour, and the gae chain restarts at every
// terminated / truncated transition, so each chain (one episode) is walked sequentially by one
// thread with exactly the reference's fp32 operation order => results are BIT-IDENTICAL to the
// reference loop, while the episodes of the rollout run in parallel.  HBM traffic: 10 bytes read +
// 8 bytes written per transition.
#include <stdarg.h>

#include "common.cuh"

using namespace prl;

namespace {

// arrays are in TIME order (index 0 = oldest stored transition)
__global__ void k_ppo_gae(int n, const float *__restrict__ values, float last_next_value,
                          const float *__restrict__ reward, const uint8_t *__restrict__ terminated,
                          const uint8_t *__restrict__ truncated, float gamma, float c_live,
                          float *__restrict__ out_gae, float *__restrict__ out_lam_return) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const bool head = (t == n - 1) || terminated[t] || truncated[t];   // newest element of its chain
    if (!head) return;
    float gae = 0.f;
    for (int s = t; s >= 0; s--) {
        const bool term = terminated[s] != 0, cut = term || truncated[s] != 0;
        if (s != t && cut) break;                                       // the next chain's head
        const float nv = (s == n - 1) ? last_next_value : values[s + 1];
        const float v = values[s];
        // reward + gamma * next_value * (~terminated) - V[i]   (left to right, fp32)
        const float td = __fsub_rn(__fadd_rn(reward[s], __fmul_rn(__fmul_rn(gamma, nv), term ? 0.f : 1.f)), v);
        // td + (gamma * lambda * mask) * gae ; the scalar product is evaluated in double by Python
        gae = __fadd_rn(td, __fmul_rn(cut ? 0.f : c_live, gae));
        out_gae[s] = gae;
        out_lam_return[s] = __fadd_rn(gae, v);
    }
}

}  // namespace

extern "C" int prl_ppo_gae(int n, const float *values_dev, float last_next_value, const float *reward_dev,
                           const uint8_t *terminated_dev, const uint8_t *truncated_dev, double gamma, double lam,
                           float *out_gae_dev, float *out_lam_return_dev, void *stream) {
    PRL_REQUIRE(n >= 0, "negative length");
    if (n == 0) return PRL_OK;
    PRL_REQUIRE(values_dev && reward_dev && terminated_dev && truncated_dev && out_gae_dev && out_lam_return_dev,
                "null argument");
    const int threads = 256;
    k_ppo_gae<<<(n + threads - 1) / threads, threads, 0, (cudaStream_t)stream>>>(
        n, values_dev, last_next_value, reward_dev, terminated_dev, truncated_dev, (float)gamma, (float)(gamma * lam),
        out_gae_dev, out_lam_return_dev);
    PRL_CUDA(cudaGetLastError());
    return PRL_OK;
}
