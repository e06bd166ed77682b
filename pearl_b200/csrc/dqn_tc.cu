// dqn_tc.cu — tensor-core DQN learner: ONE SM (one CTA) runs one complete learner, every dense
// contraction of the step on tcgen05 (3xTF32 UMMA, fp32 accumulators in TMEM), so that a launch
// with L CTAs trains L independent learners (seeds / agents) concurrently — the aggregate mode that
// fills a B200 — and a single learner needs one SM instead of 65.
//
// Same semantics as the cooperative SIMT kernel in dqn.cu (DeepQLearning.learn: sample -> Q(s,a) ->
// max_a' Q_target(s',a') -> MSE -> backward -> AdamW(amsgrad) -> scheduled soft target update; reference
// call sites in include/pearl_b200.h).  Supported shape class: two hidden layers of 64, obs % 8 == 0,
// obs <= 128, n_actions in {1,2,4,8,16}, batch in {128, 256}; everything else stays on the SIMT kernel.
//
// Per gradient step (256 threads, thread (m = tid % 128, h = tid / 128) owns batch row m of the
// current 128-row tile = TMEM lane m, and column half h):
//   phase T  target net: T1 = S' W1s^T per row tile (K chunks of 64 staged through smem); then one
//            128 x 64 x 64 product per (row tile, action slot): A tile row m = relu(T1[m] + W1a[:,a] + b1)
//            built from registers, two thread groups ping-pong (own A buffer + own accumulator) so the
//            tensor pipe always has the other group's tile; running max over action slots in registers.
//   phase O  online net forward (same tiles), loss gradient, dZ2, dH1 = dZ2 W2 (B operand = W2^T tile),
//            and the weight gradients as tensor-core products over the batch dimension:
//            dW2 = dZ2^T H1, dW1s = dZ1^T S, and [db | dW1a] = dZ^T E with E = [1 | onehot(action)];
//            their operands are explicitly transposed tiles written with a 144-byte chunk pitch
//            (bank-conflict-free column scatter).  Accumulators for all of these live in TMEM.
//   AdamW    gradients TMEM -> shared staging (M = 64 lane map), then one coalesced 16-byte sweep over the flat parameter
//            vector (parameters / moments in L2) that also writes the operand-layout weight tiles.
//
// Round 2: the weights are kept in global memory a second time IN THE UMMA OPERAND LAYOUT (hi tile = the fp32 values —
// the tensor core truncates them to TF32 — and lo tile = x - trunc_tf32(x)), written by AdamW / the soft target update
// next to the flat torch-order vectors, so staging a network's B operands is three TMA bulk copies
// (cp.async.bulk + mbarrier complete_tx) issued by one thread instead of a SIMT split-and-scatter pass of all 256
// threads (18.5 k of 172 k clk per round in round 1); the online W1 tiles travel while the all-actions products run.
// A warp-specialised rewrite of this kernel (row groups + issuer / loader / tile-builder warps) was built and measured
// in round 2 as well — bit-correct but slower; see profiles/r2_k_dqn_tc_summary.md.  What a source-level stall profile of
// THIS kernel found instead (profiles/r2b_k_dqn_tc_stalls.md): the MMA issue code must sit under elect.sync
// (umma::elect_one), the target network's small vectors are cached in shared memory between soft updates, AdamW walks one
// flat list with two groups of loads in flight, the soft target update is applied to the tiles in tile order.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>

#include <new>

#include "dqn_common.cuh"
#include "sampler.cuh"
#include "umma.cuh"

using namespace prl;

int prl_sampler_params(const prl_buf *b, int k, prl::SamplerParams *sp, size_t *smem_bytes);

namespace {

constexpr int NTH = 256;
constexpr int HID = 64;
constexpr int MAX_B = 256;
// shared memory regions (bytes)
constexpr int REG1 = 0, REG2 = 65536, REG3 = 131072, MISC_OFF = 196608;
constexpr int HALF = 32768;          // hi tile at region base, lo tile at base + HALF
// transposed-tile arena (regions 1+2) for the weight-gradient products, 64 batch rows per pass
constexpr int TL = 144;              // chunk pitch of transposed tiles
constexpr int AR_A_HI = 0, AR_A_LO = 18432, AR_B_HI = 36864, AR_B_LO = 73728, AR_E = 110592;
// TMEM columns
constexpr int TM_T1 = 0, TM_ACC0 = 128, TM_ACC1 = 192, TM_DW2 = 256, TM_DW1 = 320, TM_DE2 = 448, TM_DE1 = 480;
// A operands held in tensor memory (TS products) while the weight-gradient accumulators are not live:
// group g: hi at TM_A + 128 g, lo 64 columns further
constexpr int TM_A = 256;

struct TcLearner {            // one per CTA, in global memory
    const uint32_t *records;
    float *tiles;             // operand-layout weight tiles: online net, then target net (net_tile_floats each)
    const int32_t *slots;     // [rounds][B]
    float *w, *wt, *m, *v, *vmax;
    const float2 *scal;       // [rounds]
    float *out_mae, *out_q, *out_y;
    long long steps0;
    int buf_flags;
    int pad_;
};

struct TcArgs {
    const TcLearner *learners;
    prl_buf_layout lay;
    Dims d;
    int B, rounds, freq;
    int round0;        // this launch runs rounds [round0, round0 + rounds) of the call (chunked launches)
    float decay, omb1, beta2, omb2, eps, gamma, tau, omtau, inv_b2;
    long long *prof;   // optional [rounds][16] SM-clock stamps of one CTA
    int prof_cta;      // which CTA writes them (PRL_TC_PROF_CTA, default 0: the learners do not all run at the same speed)
};

#define TC_STAMP(idx)                                                                          \
    do {                                                                                       \
        if (a.prof && blockIdx.x == a.prof_cta && threadIdx.x == 0) a.prof[(size_t)round * 16 + (idx)] = clock64(); \
    } while (0)

struct Misc {
    // small fp32 vectors of a network: W1[:, obs + a] + b1, b2, w3, b3.  [0] online, [1] target (the target's only change at
    // a scheduled soft update, so they are loaded then and in the prologue, not every round)
    struct Smalls { float watb[16][HID]; float b2[HID], w3[HID]; float b3, pad0[3]; } sm[2];
    float y[MAX_B];
    float vpart[128][2];
    float vmax2[2][128];
    int act[MAX_B], cnt[MAX_B], slot[MAX_B];
    float rew[MAX_B], term[MAX_B];
    float redw[8][32];
    float redmae[8], reddb3[8];
    unsigned long long bar[10];   // 1,2: per-group MMA; 3,4,5: phase-O products; 6: target tiles; 7: online W1 tiles; 8: online W2 tiles
    uint32_t tmem_base;
};

__device__ __forceinline__ void cp_async16_zfill_tc(void *smem_dst, const void *gmem_src, int src_bytes) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(sa), "l"(gmem_src), "r"(src_bytes) : "memory");
}

__device__ __forceinline__ void group_sync(int g) { asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "r"(128) : "memory"); }

__device__ __forceinline__ void st_split4(float *hi_base, float *lo_base, int idx, float4 v) {
    float4 h, l;
    umma::split_tf32(v.x, h.x, l.x); umma::split_tf32(v.y, h.y, l.y);
    umma::split_tf32(v.z, h.z, l.z); umma::split_tf32(v.w, h.w, l.w);
    *reinterpret_cast<float4 *>(hi_base + idx) = h;
    *reinterpret_cast<float4 *>(lo_base + idx) = l;
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(umma::smem_u32(bar)), "r"(bytes) : "memory");
}
// TMA bulk copy global -> shared, completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(umma::smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(umma::smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ float tf32_lo(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

// Operand-layout copies of one network's matrices in global memory (floats):
//   [W1 hi: 64 x obs][W1 lo][W2 hi: 64 x 64][W2 lo]      (the shared-memory images, tile by tile; the W2^T tiles of the
//   online network are transposed out of the W2 tiles in shared memory: keeping them in global memory too cost AdamW eight
//   scattered 4-byte stores per 16-byte group, more than the transposition)
struct NetTiles { float *w1hi, *w1lo, *w2; };   // w2: the 2 x 4096-float block W2 hi | W2 lo
__host__ __device__ inline int net_tile_floats(int obs) { return 128 * obs + 2 * HID * HID; }
__device__ __forceinline__ NetTiles net_tiles(float *base, int obs) { return NetTiles{base, base + 64 * obs, base + 128 * obs}; }
// all tiles of one network from its flat parameter vector (kernel prologue, soft target update)
__device__ void rebuild_tiles(const float *__restrict__ net, const Dims &d, const NetTiles &t, int tid) {
    for (int e = tid; e < HID * d.obs; e += NTH) {
        const int j = e / d.obs, k = e - j * d.obs;
        const float x = __ldcg(net + d.oW1 + (size_t)j * d.D + k);
        const int idx = umma::tile_index(j, k, d.obs);
        t.w1hi[idx] = x; t.w1lo[idx] = tf32_lo(x);
    }
    for (int e = tid; e < HID * HID; e += NTH) {
        const int j = e >> 6, k = e & 63;
        const float x = __ldcg(net + d.oW2 + e);
        const int i1 = umma::tile_index(j, k, HID);
        t.w2[i1] = x; t.w2[4096 + i1] = tf32_lo(x);
    }
}
// the small fp32 vectors of a network (action columns + b1, b2, w3, b3); all loads of a thread issued before the first use
__device__ void load_smalls(const float *__restrict__ net, const Dims &d, Misc::Smalls &mi) {
    const int tid = threadIdx.x;
    float wa[4], bb[4];                                           // A * 64 <= 1024 elements: <= 4 per thread
#pragma unroll
    for (int u = 0; u < 4; u++) {
        const int e = tid + u * NTH;
        if (e < d.A * HID) {
            const int j = e / d.A, a = e - j * d.A;
            wa[u] = __ldcg(net + d.oW1 + (size_t)j * d.D + d.obs + a);
            bb[u] = __ldcg(net + d.ob1 + j);
        }
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
        const int e = tid + u * NTH;
        if (e < d.A * HID) { const int j = e / d.A, a = e - j * d.A; mi.watb[a][j] = wa[u] + bb[u]; }
    }
    if (tid < HID) { mi.b2[tid] = __ldcg(net + d.ob2 + tid); mi.w3[tid] = __ldcg(net + d.oW3 + tid); }
    if (tid == 0) mi.b3 = __ldcg(net + d.ob3);
}

__device__ __forceinline__ float fast_sqrt(float x) {
    float r;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
struct AdamScalarsTc : AdamScalars { float inv_bc2_sqrt; };

// AdamW on parameters [i0, i0 + valid) with gradients g[OFF .. OFF + COUNT): COUNT is a compile-time
// multiple of 16 so that g stays in registers; `valid` <= COUNT masks the tail.  Per 16-parameter batch:
// all loads first (one L2 round trip), then arithmetic, then stores.
__device__ __forceinline__ float adam_math(float w, float &m, float &v, float &x, float g, const AdamScalarsTc &hs) {
    float p = __fmul_rn(w, hs.decay);
    m = fmaf(hs.omb1, g - m, m);
    v = __fadd_rn(__fmul_rn(v, hs.beta2), __fmul_rn(__fmul_rn(hs.omb2, g), g));
    x = fmaxf(x, v);
    // one SM updates all 13.5k parameters: MUFU sqrt / divide (<= 2 ulp) instead of the IEEE
    // software sequences; well inside the 1e-4 parity budget
    const float denom = __fadd_rn(__fmul_rn(fast_sqrt(x), hs.inv_bc2_sqrt), hs.eps);
    return __fadd_rn(p, __fdividef(__fmul_rn(-hs.step_size, m), denom));
}
// T1[tile] = X[tile rows] W1s^T, X = state (online) or next_state (target).  The two 128-thread groups
// take alternate row tiles; each thread streams its own row from global memory, splits it and writes it
// straight into tensor memory (TS product: no shared-memory staging of A), 64 columns of K per pass.
__device__ void layer1_all_tiles(const TcArgs &a, const TcLearner &L, Misc &mi, char *smem, int field_off, uint32_t tm,
                                 uint32_t tlane, uint32_t &parg, int round = 0, bool stamp = false) {
    const int m = threadIdx.x & 127, g = threadIdx.x >> 7;
    const int ntiles = a.B >> 7;
    const umma::Tile B_hi = umma::make_tile(smem + REG1, a.d.obs, 128), B_lo = umma::make_tile(smem + REG1 + HALF, a.d.obs, 128);
    const uint32_t a_hi = TM_A + 128 * g, a_lo = a_hi + 64;
    uint64_t *bar = reinterpret_cast<uint64_t *>(mi.bar);
    // staging buffer of this group in region 2: [128 rows][64 floats], 16-byte chunk c of row r stored at
    // chunk position c ^ (r & 7): rows arrive by coalesced 16-byte cp.async (16 lanes per row), and the
    // row-owning thread reads its row back with conflict-free 128-bit loads.
    float *sbuf = reinterpret_cast<float *>(smem + REG2 + g * 32768);
    for (int t = g; t < ntiles; t += 2) {
        for (int kc = 0; kc * 64 < a.d.obs; kc++) {
            if (stamp && kc == 1) TC_STAMP(10);
#pragma unroll 4
            for (int i = 0; i < 16; i++) {
                const int item = i * 128 + m, rr = item >> 4, c = item & 15, k = kc * 64 + c * 4;
                const float *src = reinterpret_cast<const float *>(L.records + (size_t)mi.slot[t * 128 + rr] * a.lay.record_words) +
                                   field_off + k;
                const int nb = k < a.d.obs ? 16 : 0;
                cp_async16_zfill_tc(sbuf + rr * 64 + ((c ^ (rr & 7)) << 2), nb ? src : reinterpret_cast<const float *>(L.records), nb);
            }
            cp_async_commit();
            cp_async_wait<0>();
            group_sync(g);
#pragma unroll
            for (int half = 0; half < 2; half++) {
                float hi[32], lo[32];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int c = half * 8 + i;
                    const float4 v = *reinterpret_cast<const float4 *>(sbuf + m * 64 + ((c ^ (m & 7)) << 2));
                    umma::split_tf32(v.x, hi[4 * i + 0], lo[4 * i + 0]); umma::split_tf32(v.y, hi[4 * i + 1], lo[4 * i + 1]);
                    umma::split_tf32(v.z, hi[4 * i + 2], lo[4 * i + 2]); umma::split_tf32(v.w, hi[4 * i + 3], lo[4 * i + 3]);
                }
                umma::tmem_st32(tlane + a_hi + half * 32, hi);
                umma::tmem_st32(tlane + a_lo + half * 32, lo);
            }
            umma::tmem_st_wait();
            if (stamp && kc == 1) TC_STAMP(11);
            umma::fence_before_thread_sync();
            group_sync(g);
            if (stamp && kc == 1) TC_STAMP(12);
            if (m < 32 && umma::elect_one()) {
                umma::fence_after_thread_sync();
                const int keff = min(64, a.d.obs - kc * 64);
                umma::gemm3_ts(tm + TM_T1 + t * 64, tm + a_hi, tm + a_lo, B_hi.shifted(kc * 2048), B_lo.shifted(kc * 2048), 128, HID,
                               keff, kc > 0);
                umma::mma_commit(&bar[1 + g]);
            }
            if (stamp && kc == 1) TC_STAMP(13);
            umma::mbar_wait(&bar[1 + g], parg);
            parg ^= 1;
            umma::fence_after_thread_sync();
            if (stamp && kc == 1) TC_STAMP(14);
        }
    }
    if (stamp) TC_STAMP(15);
    umma::fence_before_thread_sync();
    __syncthreads();
    umma::fence_after_thread_sync();
}

// v[c] = this lane's (row's) value of column c; returns, in lane c, the sum of column c over the warp's
// 32 rows: reduce-scatter butterfly, 31 shuffles instead of 32 x 5.
__device__ __forceinline__ float warp_colsum32(float *v, int lane) {
#pragma unroll
    for (int n = 16; n >= 1; n >>= 1) {
        const bool up = (lane & n) != 0;
#pragma unroll
        for (int i = 0; i < n; i++) {
            const float send = up ? v[i] : v[i + n];
            const float keep = up ? v[i + n] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, n);
        }
    }
    return v[0];
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(NTH, 1) k_dqn_tc(const TcArgs a) {
    extern __shared__ __align__(1024) char smem[];
    Misc &mi = *reinterpret_cast<Misc *>(smem + MISC_OFF);
    TcLearner L = a.learners[blockIdx.x];
    if (a.round0) {   // a later chunk of the same call: shift every per-round array once
        L.slots += (size_t)a.round0 * a.B;
        L.scal += a.round0;
        L.out_mae += a.round0;
        if (L.out_q) L.out_q += (size_t)a.round0 * a.B;
        if (L.out_y) L.out_y += (size_t)a.round0 * a.B;
        L.steps0 += a.round0;
    }
    const Dims &d = a.d;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m = tid & 127, h = tid >> 7;
    const int ntiles = a.B >> 7;
    const int W = a.lay.record_words;

    // De-phase the learners: identical CTAs started together run in lock-step and hit L2 / HBM with their AdamW sweeps
    // (432 KB each) and row gathers all at once; a start offset of up to ~50 us spreads those bursts over the round.
    if (tid == 0 && gridDim.x > 1) {
        const long long t0 = clock64(), wait = (long long)(blockIdx.x % 48) * 2000;
        while (clock64() - t0 < wait) __nanosleep(100);
    }
    if (warp == 0) umma::tmem_alloc(&mi.tmem_base, 512);
    if (tid == 0)
        for (int i = 0; i < 10; i++) umma::mbar_init(reinterpret_cast<uint64_t *>(&mi.bar[i]), 1);
    // the operand-layout tiles follow the flat parameters (which the host may have changed between calls)
    const NetTiles To = net_tiles(L.tiles, d.obs), Tt = net_tiles(L.tiles + net_tile_floats(d.obs), d.obs);
    rebuild_tiles(L.w, d, To, tid);
    rebuild_tiles(L.wt, d, Tt, tid);
    fence_proxy_async_all();
    umma::fence_before_thread_sync();
    __syncthreads();
    umma::fence_after_thread_sync();
    const uint32_t w1_bytes = 64u * d.obs * 4, w2_bytes = 2u * HID * HID * 4;
    // B-operand tiles of one network by TMA: W1 hi / lo into region 1, the W2 block into region 3
    auto tma_w1 = [&](const NetTiles &t, uint64_t *b) {
        mbar_expect_tx(b, 2 * w1_bytes);
        bulk_g2s(smem + REG1, t.w1hi, w1_bytes, b);
        bulk_g2s(smem + REG1 + HALF, t.w1lo, w1_bytes, b);
    };
    auto tma_w2 = [&](const NetTiles &t, uint64_t *b) {
        mbar_expect_tx(b, w2_bytes);
        bulk_g2s(smem + REG3, t.w2, w2_bytes, b);
    };
    const uint32_t tm = mi.tmem_base;
    const uint32_t tlane = tm + ((uint32_t)((warp & 3) * 32) << 16);   // this warp's 32 TMEM lanes
    uint32_t parg = 0, par3 = 0, par4 = 0, par5 = 0;                  // mbarrier phase parities
    uint64_t *bar = reinterpret_cast<uint64_t *>(mi.bar);

    for (int round = 0; round < a.rounds; round++) {
        TC_STAMP(0);
        // ---- per-round row scalars + L2 prefetch of the NEXT round's transitions
        if (tid < a.B) {
            const int slot = L.slots[(size_t)round * a.B + tid];
            const uint32_t *r = L.records + (size_t)slot * W;
            mi.slot[tid] = slot;
            mi.act[tid] = (int)r[a.lay.off_action];
            mi.rew[tid] = __uint_as_float(r[a.lay.off_reward]);
            const uint32_t fl = r[a.lay.off_flags];
            mi.term[tid] = (fl & 1u) ? 1.f : 0.f;
            mi.cnt[tid] = (int)((fl >> 8) & 0xffffu);
            if (round + 1 < a.rounds) {
                const char *nx = reinterpret_cast<const char *>(L.records + (size_t)L.slots[(size_t)(round + 1) * a.B + tid] * W);
                for (int o = 0; o < W * 4; o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + o));
            }
        }
        // ---- scheduled soft target update happens BEFORE this round's gradient step
        //      (deep_td_learning.py:283-284: (training_steps + 1) % freq == 0)
        const bool soft_upd = (L.steps0 + round + 2) % a.freq == 0;
        if (soft_upd) {
            for (int i = tid; i < d.P; i += NTH) L.wt[i] = soft_update(__ldcg(L.w + i), __ldcg(L.wt + i), a.tau, a.omtau);
            // The operand-layout tiles get the same elementwise update IN TILE ORDER (the hi tiles hold exactly the fp32 values of
            // the flat vectors, permuted; lo = residual of the new value): coalesced 16-byte accesses instead of rebuilding the
            // target tiles from the flat vector with scattered 4-byte stores.  Same function of the same inputs: bit-identical.
            auto update_tile = [&](const float *on_hi, float *tg_hi, float *tg_lo, int n) {
                for (int i = tid * 4; i < n; i += NTH * 4) {
                    const float4 w = __ldcg(reinterpret_cast<const float4 *>(on_hi + i)), t = __ldcg(reinterpret_cast<const float4 *>(tg_hi + i));
                    const float4 r = make_float4(soft_update(w.x, t.x, a.tau, a.omtau), soft_update(w.y, t.y, a.tau, a.omtau),
                                                 soft_update(w.z, t.z, a.tau, a.omtau), soft_update(w.w, t.w, a.tau, a.omtau));
                    *reinterpret_cast<float4 *>(tg_hi + i) = r;
                    *reinterpret_cast<float4 *>(tg_lo + i) = make_float4(tf32_lo(r.x), tf32_lo(r.y), tf32_lo(r.z), tf32_lo(r.w));
                }
            };
            update_tile(To.w1hi, Tt.w1hi, Tt.w1lo, HID * d.obs);
            update_tile(To.w2, Tt.w2, Tt.w2 + HID * HID, HID * HID);
            fence_proxy_async_all();
        }
        __syncthreads();

        // ================= phase T: y = max_a' Q_target(s', a') * gamma * (1 - term) + r =================
        TC_STAMP(1);
        if (tid == 0) {   // target tiles: three TMA bulk copies (regions 1 and 3 are free: every product of the last round was waited for)
            mbar_expect_tx(bar + 6, 2 * w1_bytes + w2_bytes);
            bulk_g2s(smem + REG1, Tt.w1hi, w1_bytes, bar + 6);
            bulk_g2s(smem + REG1 + HALF, Tt.w1lo, w1_bytes, bar + 6);
            bulk_g2s(smem + REG3, Tt.w2, w2_bytes, bar + 6);
        }
        if (soft_upd || round == 0) load_smalls(L.wt, d, mi.sm[1]);   // the target's small vectors only change at a soft update
        umma::mbar_wait(bar + 6, round & 1);
        __syncthreads();
        TC_STAMP(2);
        layer1_all_tiles(a, L, mi, smem, a.lay.off_next_state, tm, tlane, parg);
        if (tid == 0) tma_w1(To, bar + 7);      // region 1 is free again: the online W1 tiles travel while the all-actions products run
        TC_STAMP(3);
        {
            // group g = h: own A operand in tensor memory and own accumulator; the two groups ping-pong
            const umma::Tile B_hi = umma::make_tile(smem + REG3, 64, 128), B_lo = umma::make_tile(smem + REG3 + 16384, 64, 128);
            const uint32_t acc_col = h ? TM_ACC1 : TM_ACC0;
            const uint32_t a_hi = TM_A + 128 * h, a_lo = a_hi + 64;
            for (int t = 0; t < ntiles; t++) {
                const int row = t * 128 + m;
                float t1[64];
                umma::tmem_ld32(tlane + TM_T1 + t * 64, t1);
                umma::tmem_ld32(tlane + TM_T1 + t * 64 + 32, t1 + 32);
                const int cnt = mi.cnt[row];
                const uint8_t *ids = reinterpret_cast<const uint8_t *>(L.records + (size_t)mi.slot[row] * W + a.lay.off_avail);
                float best = -INFINITY;
                for (int act = h; act < d.A; act += 2) {
                    int id = act;
                    if ((L.buf_flags & PRL_BUF_DYNAMIC_ACTIONS) && act < cnt) id = ids[act];
#pragma unroll
                    for (int half = 0; half < 2; half++) {
                        float hi[32], lo[32];
#pragma unroll
                        for (int c4 = 0; c4 < 8; c4++) {
                            const float4 wv = *reinterpret_cast<const float4 *>(&mi.sm[1].watb[id][half * 32 + 4 * c4]);
                            const int c = half * 32 + 4 * c4;
                            umma::split_tf32(fmaxf(t1[c + 0] + wv.x, 0.f), hi[4 * c4 + 0], lo[4 * c4 + 0]);
                            umma::split_tf32(fmaxf(t1[c + 1] + wv.y, 0.f), hi[4 * c4 + 1], lo[4 * c4 + 1]);
                            umma::split_tf32(fmaxf(t1[c + 2] + wv.z, 0.f), hi[4 * c4 + 2], lo[4 * c4 + 2]);
                            umma::split_tf32(fmaxf(t1[c + 3] + wv.w, 0.f), hi[4 * c4 + 3], lo[4 * c4 + 3]);
                        }
                        umma::tmem_st32(tlane + a_hi + half * 32, hi);
                        umma::tmem_st32(tlane + a_lo + half * 32, lo);
                    }
                    umma::tmem_st_wait();
                    
                    umma::fence_before_thread_sync();
                    group_sync(h);
                    
                    if (m < 32 && umma::elect_one()) {
                        umma::fence_after_thread_sync();
                        umma::gemm3_ts(tm + acc_col, tm + a_hi, tm + a_lo, B_hi, B_lo, 128, HID, HID, false);
                        umma::mma_commit(&bar[1 + h]);
                    }
                    
                    umma::mbar_wait(&bar[1 + h], parg);
                    parg ^= 1;
                    umma::fence_after_thread_sync();
                    
                    float acc[64];
                    umma::tmem_ld32(tlane + acc_col, acc);
                    umma::tmem_ld32(tlane + acc_col + 32, acc + 32);
                    float q = 0.f;
#pragma unroll
                    for (int c4 = 0; c4 < 16; c4++) {
                        const float4 w3v = *reinterpret_cast<const float4 *>(&mi.sm[1].w3[4 * c4]);
                        const float4 b2v = *reinterpret_cast<const float4 *>(&mi.sm[1].b2[4 * c4]);
                        q = fmaf(w3v.x, fmaxf(acc[4 * c4 + 0] + b2v.x, 0.f), q);
                        q = fmaf(w3v.y, fmaxf(acc[4 * c4 + 1] + b2v.y, 0.f), q);
                        q = fmaf(w3v.z, fmaxf(acc[4 * c4 + 2] + b2v.z, 0.f), q);
                        q = fmaf(w3v.w, fmaxf(acc[4 * c4 + 3] + b2v.w, 0.f), q);
                    }
                    q += mi.sm[1].b3;
                    if (act >= cnt) q = -INFINITY;   // next_state_action_values[mask] = -inf
                    best = fmaxf(best, q);
                    
                    
                }
                mi.vmax2[h][m] = best;
                umma::fence_before_thread_sync();
                __syncthreads();
                if (h == 0) {
                    const float v = fmaxf(mi.vmax2[0][m], mi.vmax2[1][m]);
                    mi.y[row] = __fadd_rn(__fmul_rn(__fmul_rn(v, a.gamma), 1.f - mi.term[row]), mi.rew[row]);
                }
                __syncthreads();
            }
        }

        // ================= phase O: online forward, loss, backward =================
        TC_STAMP(4);
        if (tid == 0) tma_w2(To, bar + 8);         // region 3 held the target W2 until the last all-actions product
        load_smalls(L.w, d, mi.sm[0]);
        umma::mbar_wait(bar + 8, round & 1);
        {   // W2^T tiles (B operand of dH1 = dZ2 W2) out of the W2 tiles: element (k, j) <- (j, k)
            const float *w2 = reinterpret_cast<const float *>(smem + REG3);
            float *w2t = reinterpret_cast<float *>(smem + REG3 + 32768);
            for (int e = tid; e < HID * HID; e += NTH) {
                const int j = e >> 6, k = e & 63, i1 = umma::tile_index(j, k, HID), i2 = umma::tile_index(k, j, HID);
                w2t[i2] = w2[i1]; w2t[4096 + i2] = w2[4096 + i1];
            }
            umma::fence_async_smem();
        }
        umma::mbar_wait(bar + 7, round & 1);
        __syncthreads();
        TC_STAMP(5);
        layer1_all_tiles(a, L, mi, smem, a.lay.off_state, tm, tlane, parg, round, true);
        TC_STAMP(6);
        float dw3_acc = 0.f, mae_acc = 0.f, db3_acc = 0.f;
        const umma::Tile W2_hi = umma::make_tile(smem + REG3, 64, 128), W2_lo = umma::make_tile(smem + REG3 + 16384, 64, 128);
        const umma::Tile W2T_hi = umma::make_tile(smem + REG3 + 32768, 64, 128), W2T_lo = umma::make_tile(smem + REG3 + 49152, 64, 128);
        float *r2hi = reinterpret_cast<float *>(smem + REG2), *r2lo = r2hi + HALF / 4;
        const umma::Tile R2_hi = umma::make_tile(r2hi, 64, 128), R2_lo = umma::make_tile(r2lo, 64, 128);
        for (int t = 0; t < ntiles; t++) {
            const int row = t * 128 + m, c0 = h * 32;
            float h1[32];
            umma::tmem_ld32(tlane + TM_T1 + t * 64 + c0, h1);
            {
                const int ai = mi.act[row];
#pragma unroll
                for (int c = 0; c < 32; c++) h1[c] = fmaxf(h1[c] + mi.sm[0].watb[ai][c0 + c], 0.f);
#pragma unroll
                for (int c4 = 0; c4 < 8; c4++)
                    st_split4(r2hi, r2lo, umma::tile_index(m, c0 + 4 * c4, 64),
                              make_float4(h1[4 * c4], h1[4 * c4 + 1], h1[4 * c4 + 2], h1[4 * c4 + 3]));
            }
            umma::fence_async_smem();
            umma::fence_before_thread_sync();
            __syncthreads();
            if (warp == 0 && umma::elect_one()) {
                umma::fence_after_thread_sync();
                umma::gemm3(tm + TM_ACC0, R2_hi, R2_lo, W2_hi, W2_lo, 128, HID, HID, false);
                umma::mma_commit(&bar[4]);
            }
            umma::mbar_wait(&bar[4], par4);
            par4 ^= 1;
            umma::fence_after_thread_sync();
            float z[32];   // h2, then dz2
            umma::tmem_ld32(tlane + TM_ACC0 + c0, z);
            float part = 0.f;
#pragma unroll
            for (int c = 0; c < 32; c++) { z[c] = fmaxf(z[c] + mi.sm[0].b2[c0 + c], 0.f); part = fmaf(mi.sm[0].w3[c0 + c], z[c], part); }
            mi.vpart[m][h] = part;
            umma::fence_before_thread_sync();
            __syncthreads();
            const float q = mi.vpart[m][0] + mi.vpart[m][1] + mi.sm[0].b3;
            const float y = mi.y[row];
            const float dq = (q - y) * a.inv_b2;
            if (h == 0) {
                if (L.out_q) L.out_q[(size_t)round * a.B + row] = q;
                if (L.out_y) L.out_y[(size_t)round * a.B + row] = y;
                mae_acc += fabsf(q - y);
                db3_acc += dq;
            }
            // dW3[c0 + c] += sum_rows dq * h2 : reduce over this warp's 32 rows, lane c keeps column c
            {
                float cs[32];
#pragma unroll
                for (int c = 0; c < 32; c++) cs[c] = dq * z[c];
                dw3_acc += warp_colsum32(cs, lane);   // lane c receives the sum of column c over the 32 rows
            }
#pragma unroll
            for (int c = 0; c < 32; c++) z[c] = (z[c] > 0.f) ? dq * mi.sm[0].w3[c0 + c] : 0.f;   // dZ2
#pragma unroll
            for (int c4 = 0; c4 < 8; c4++)
                st_split4(r2hi, r2lo, umma::tile_index(m, c0 + 4 * c4, 64),
                          make_float4(z[4 * c4], z[4 * c4 + 1], z[4 * c4 + 2], z[4 * c4 + 3]));
            umma::fence_async_smem();
            umma::fence_before_thread_sync();
            __syncthreads();
            if (warp == 0 && umma::elect_one()) {
                umma::fence_after_thread_sync();
                umma::gemm3(tm + TM_ACC1, R2_hi, R2_lo, W2T_hi, W2T_lo, 128, HID, HID, false);   // dH1 = dZ2 W2
                umma::mma_commit(&bar[5]);
            }
            umma::mbar_wait(&bar[5], par5);
            par5 ^= 1;
            umma::fence_after_thread_sync();
            float dz1[32];
            umma::tmem_ld32(tlane + TM_ACC1 + c0, dz1);
            if (t == 0) TC_STAMP(7);
#pragma unroll
            for (int c = 0; c < 32; c++) dz1[c] = (h1[c] > 0.f) ? dz1[c] : 0.f;

            // ---- weight gradients: contractions over the batch rows, 64 rows per pass
            float *arena = reinterpret_cast<float *>(smem);
            const umma::Tile TA_hi = umma::make_tile(smem + AR_A_HI, 64, TL), TA_lo = umma::make_tile(smem + AR_A_LO, 64, TL);
            const umma::Tile TB_hi = umma::make_tile(smem + AR_B_HI, 64, TL), TB_lo = umma::make_tile(smem + AR_B_LO, 64, TL);
            const umma::Tile TE = umma::make_tile(smem + AR_E, 64, TL);
            for (int hf = 0; hf < 2; hf++) {
                const bool mine = (m >> 6) == hf;
                const int rr = m & 63;
                const bool first = (t == 0 && hf == 0);
                umma::fence_before_thread_sync();
                __syncthreads();   // previous products have been waited for: the arena is free
                if (mine) {
#pragma unroll
                    for (int c = 0; c < 32; c++) {
                        const int idx = umma::tile_index2(c0 + c, rr, 64, TL);
                        float hi, lo;
                        umma::split_tf32(z[c], hi, lo);            // dZ2^T
                        arena[AR_A_HI / 4 + idx] = hi; arena[AR_A_LO / 4 + idx] = lo;
                        umma::split_tf32(h1[c], hi, lo);           // H1^T
                        arena[AR_B_HI / 4 + idx] = hi; arena[AR_B_LO / 4 + idx] = lo;
                    }
                    const int ai = mi.act[row];
#pragma unroll
                    for (int e = 0; e < 16; e++) {                 // E^T: [1 | onehot(action)]
                        const int er = h * 16 + e;
                        arena[AR_E / 4 + umma::tile_index2(er, rr, 64, TL)] = (er == 0 || er == ai + 1) ? 1.f : 0.f;
                    }
                }
                umma::fence_async_smem();
                umma::fence_before_thread_sync();
                __syncthreads();
                if (warp == 0 && umma::elect_one()) {
                    umma::fence_after_thread_sync();
                    umma::gemm3(tm + TM_DW2, TA_hi, TA_lo, TB_hi, TB_lo, 64, HID, 64, !first);
                    umma::gemm3(tm + TM_DE2, TA_hi, TA_lo, TE, TE, 64, 32, 64, !first, false, true);
                    umma::mma_commit(&bar[3]);
                }
                // S^T: every warp reads whole state rows of this half coalesced (lane = k) and scatters them into column rq of
                // the transposed tile; the first four rows are requested BEFORE waiting for the dW2 product
                float sv[4][4];
                auto st_rows_load = [&](int r0) {
#pragma unroll
                    for (int ri = 0; ri < 4; ri++) {
                        const int rq = warp + 8 * (r0 + ri);
                        const float *src = reinterpret_cast<const float *>(L.records + (size_t)mi.slot[t * 128 + hf * 64 + rq] * W) + a.lay.off_state;
#pragma unroll
                        for (int jj = 0; jj < 4; jj++) sv[ri][jj] = (lane + 32 * jj < d.obs) ? __ldg(src + lane + 32 * jj) : 0.f;
                    }
                };
                auto st_rows_scatter = [&](int r0) {
#pragma unroll
                    for (int ri = 0; ri < 4; ri++) {
                        const int rq = warp + 8 * (r0 + ri);
#pragma unroll
                        for (int jj = 0; jj < 4; jj++)
                            if (lane + 32 * jj < d.obs) {
                                float hi, lo;
                                umma::split_tf32(sv[ri][jj], hi, lo);
                                const int idx = umma::tile_index2(lane + 32 * jj, rq, 64, TL);
                                arena[AR_B_HI / 4 + idx] = hi; arena[AR_B_LO / 4 + idx] = lo;
                            }
                    }
                };
                st_rows_load(0);
                umma::mbar_wait(&bar[3], par3);
                par3 ^= 1;
                umma::fence_after_thread_sync();
                umma::fence_before_thread_sync();
                __syncthreads();
                if (mine) {
#pragma unroll
                    for (int c = 0; c < 32; c++) {
                        const int idx = umma::tile_index2(c0 + c, rr, 64, TL);
                        float hi, lo;
                        umma::split_tf32(dz1[c], hi, lo);          // dZ1^T
                        arena[AR_A_HI / 4 + idx] = hi; arena[AR_A_LO / 4 + idx] = lo;
                    }
                }
                st_rows_scatter(0);
                st_rows_load(4);
                st_rows_scatter(4);
                umma::fence_async_smem();
                umma::fence_before_thread_sync();
                __syncthreads();
                if (warp == 0 && umma::elect_one()) {
                    umma::fence_after_thread_sync();
                    umma::gemm3(tm + TM_DW1, TA_hi, TA_lo, TB_hi, TB_lo, 64, d.obs, 64, !first);
                    umma::gemm3(tm + TM_DE1, TA_hi, TA_lo, TE, TE, 64, 32, 64, !first, false, true);
                    umma::mma_commit(&bar[3]);
                }
                umma::mbar_wait(&bar[3], par3);
                par3 ^= 1;
                umma::fence_after_thread_sync();
            }
        }

        TC_STAMP(8);
        // ================= AdamW (gradients straight from TMEM; M = 64 rows live in lanes 32q + (0..15)) ==========
        mi.redw[warp][lane] = dw3_acc;
        if (lane == 0) { mi.redmae[warp] = 0.f; mi.reddb3[warp] = 0.f; }
        {
            const float ms = warp_sum(mae_acc), ds = warp_sum(db3_acc);
            if (lane == 0) { mi.redmae[warp] = ms; mi.reddb3[warp] = ds; }
        }
        umma::fence_before_thread_sync();
        __syncthreads();
        umma::fence_after_thread_sync();
        {
            // 1) gradients TMEM -> shared staging in region 2 (the arena is free).  Row
            //    pitches are odd, so the 16 live lanes of a warp (rows 16q .. 16q+15) hit distinct banks.
            float *gs_w2 = reinterpret_cast<float *>(smem + REG2);     // [64][65]
            const int pitch1 = d.D | 1;
            float *gs_w1 = gs_w2 + 64 * 65;                            // [64][pitch1]   (state cols, then action cols)
            float *gs_b1 = gs_w1 + 64 * pitch1, *gs_b2 = gs_b1 + 64;
            const int j = (warp & 3) * 16 + (lane & 15);               // M = 64 lane map: lanes >= 16 hold nothing
            const bool live = lane < 16;
            float g[32];
            umma::tmem_ld32(tlane + TM_DW2 + h * 32, g);
            if (live)
#pragma unroll
                for (int c = 0; c < 32; c++) gs_w2[j * 65 + h * 32 + c] = g[c];
            for (int cc = 0; cc < 2; cc++) {
                const int kbase = h * 64 + cc * 32;
                if (kbase < d.obs) {   // warp-uniform
                    umma::tmem_ld32(tlane + TM_DW1 + kbase, g);
                    if (live)
#pragma unroll
                        for (int c = 0; c < 32; c++)
                            if (kbase + c < d.obs) gs_w1[j * pitch1 + kbase + c] = g[c];
                }
            }
            umma::tmem_ld32(tlane + (h ? TM_DE1 : TM_DE2), g);
            if (live) {
                if (h == 0) gs_b2[j] = g[0];
                else {
                    gs_b1[j] = g[0];
#pragma unroll
                    for (int k = 0; k < 16; k++)
                        if (k < d.A) gs_w1[j * pitch1 + d.obs + k] = g[1 + k];
                }
            }
            umma::fence_before_thread_sync();
            __syncthreads();
            // 2) AdamW over the flat parameter vector, all 256 threads, fully coalesced 16-byte accesses.  The sweep is bound by
            //    L2 round trips, not bytes: W1 | W2 are walked as ONE flat list of 16-byte groups (no 4-lane tail passes per W1
            //    row) in batches of two groups per thread whose 8 loads are issued before the first use, and the loads of the
            //    small vectors (b1 | b2 | W3 | b3, one parameter per thread) are in flight across the whole sweep.
            const float2 sc = L.scal[round];
            AdamScalarsTc hs;
            hs.decay = a.decay; hs.omb1 = a.omb1; hs.beta2 = a.beta2; hs.omb2 = a.omb2; hs.eps = a.eps;
            hs.step_size = sc.x; hs.bc2_sqrt = sc.y; hs.inv_bc2_sqrt = 1.0f / sc.y;
            int si = -1;                // this thread's small parameter
            float sg = 0.f, sw = 0.f, sm = 0.f, sv2 = 0.f, sx = 0.f;
            if (tid < 64) { si = d.ob1 + tid; sg = gs_b1[tid]; }
            else if (tid < 128) { si = d.ob2 + tid - 64; sg = gs_b2[tid - 64]; }
            else if (tid < 192) {       // W3: sum the four row-quarter partials of this column in fixed order
                const int col = tid - 128, hh = col >> 5, c = col & 31;
                sg = ((mi.redw[hh * 4 + 0][c] + mi.redw[hh * 4 + 1][c]) + mi.redw[hh * 4 + 2][c]) + mi.redw[hh * 4 + 3][c];
                si = d.oW3 + col;
            } else if (tid == 192) {
                sg = ((mi.reddb3[0] + mi.reddb3[1]) + mi.reddb3[2]) + mi.reddb3[3];
                si = d.ob3;
                const float e = ((mi.redmae[0] + mi.redmae[1]) + mi.redmae[2]) + mi.redmae[3];
                L.out_mae[round] = e / (float)a.B;   // reported "loss": mean |q - y|
            }
            if (si >= 0) { sw = __ldcg(L.w + si); sm = __ldcg(L.m + si); sv2 = __ldcg(L.v + si); sx = __ldcg(L.vmax + si); }
            const bool vecD = ((d.D & 3) == 0) && ((d.oW2 & 3) == 0);
            if (vecD) {
                constexpr int U = 2;
                const int D4 = d.D >> 2, n1 = HID * D4, ntot = n1 + HID * HID / 4;
                for (int q0 = tid; q0 < ntot; q0 += U * NTH) {
                    float4 w4[U], m4[U], v4[U], x4[U];
#pragma unroll
                    for (int u = 0; u < U; u++) {
                        const int q = q0 + u * NTH;
                        if (q < ntot) {
                            const int i = q < n1 ? d.oW1 + q * 4 : d.oW2 + (q - n1) * 4;   // W1 rows are contiguous: row * D + c4 * 4 == q * 4
                            w4[u] = __ldcg(reinterpret_cast<const float4 *>(L.w + i)); m4[u] = __ldcg(reinterpret_cast<const float4 *>(L.m + i));
                            v4[u] = __ldcg(reinterpret_cast<const float4 *>(L.v + i)); x4[u] = __ldcg(reinterpret_cast<const float4 *>(L.vmax + i));
                        }
                    }
#pragma unroll
                    for (int u = 0; u < U; u++) {
                        const int q = q0 + u * NTH;
                        if (q < ntot) {
                            int i;
                            const float *gp;
                            float *thi, *tlo;             // the same 16-byte chunk of the operand-layout tiles (hi = the value, lo = residual)
                            if (q < n1) {
                                const int row = q / D4, c = (q - row * D4) * 4;
                                i = d.oW1 + q * 4;
                                gp = gs_w1 + row * pitch1 + c;
                                const int ti = umma::tile_index(row, c, d.obs);
                                thi = c < d.obs ? To.w1hi + ti : nullptr; tlo = To.w1lo + ti;
                            } else {
                                const int q4 = q - n1, row = q4 >> 4, c = (q4 & 15) * 4;
                                i = d.oW2 + q4 * 4;
                                gp = gs_w2 + row * 65 + c;
                                const int ti = umma::tile_index(row, c, HID);
                                thi = To.w2 + ti; tlo = To.w2 + 4096 + ti;
                            }
                            float4 w = w4[u], mm = m4[u], vv = v4[u], xx = x4[u];
                            w.x = adam_math(w.x, mm.x, vv.x, xx.x, gp[0], hs); w.y = adam_math(w.y, mm.y, vv.y, xx.y, gp[1], hs);
                            w.z = adam_math(w.z, mm.z, vv.z, xx.z, gp[2], hs); w.w = adam_math(w.w, mm.w, vv.w, xx.w, gp[3], hs);
                            *reinterpret_cast<float4 *>(L.w + i) = w; *reinterpret_cast<float4 *>(L.m + i) = mm;
                            *reinterpret_cast<float4 *>(L.v + i) = vv; *reinterpret_cast<float4 *>(L.vmax + i) = xx;
                            if (thi) {
                                *reinterpret_cast<float4 *>(thi) = w;
                                *reinterpret_cast<float4 *>(tlo) = make_float4(tf32_lo(w.x), tf32_lo(w.y), tf32_lo(w.z), tf32_lo(w.w));
                            }
                        }
                    }
                }
            } else {
                for (int row = warp; row < HID; row += 8)
                    for (int c = lane; c < d.D; c += 32) {
                        const int i = d.oW1 + row * d.D + c;
                        float mm = __ldcg(L.m + i), vv = __ldcg(L.v + i), xx = __ldcg(L.vmax + i);
                        L.w[i] = adam_math(__ldcg(L.w + i), mm, vv, xx, gs_w1[row * pitch1 + c], hs);
                        L.m[i] = mm; L.v[i] = vv; L.vmax[i] = xx;
                    }
                for (int e = tid; e < HID * HID; e += NTH) {
                    const int i = d.oW2 + e;
                    float mm = __ldcg(L.m + i), vv = __ldcg(L.v + i), xx = __ldcg(L.vmax + i);
                    L.w[i] = adam_math(__ldcg(L.w + i), mm, vv, xx, gs_w2[(e >> 6) * 65 + (e & 63)], hs);
                    L.m[i] = mm; L.v[i] = vv; L.vmax[i] = xx;
                }
                __syncthreads();
                rebuild_tiles(L.w, d, To, tid);         // unaligned shapes (D % 4 != 0): tiles from the flat vector
            }
            fence_proxy_async_all();                    // the tile writes are read by TMA in the next round
            if (si >= 0) {
                L.w[si] = adam_math(sw, sm, sv2, sx, sg, hs);
                L.m[si] = sm; L.v[si] = sv2; L.vmax[si] = sx;
            }
        }
        umma::fence_before_thread_sync();
        __syncthreads();
        umma::fence_after_thread_sync();
        TC_STAMP(9);
    }
    if (warp == 0) umma::tmem_dealloc(tm, 512);
}

// a few microseconds of nothing: lets the learner CTAs of the main stream take their SMs before the next chunk's
// index producers (side stream) become ready
__global__ void k_delay(unsigned ns) {
    const long long t0 = clock64();
    while (clock64() - t0 < (long long)ns * 2) __nanosleep(200);
}

// all learners' index streams in one launch: CTA i draws `rounds` samples for learner i
struct MultiSampler {
    uint32_t *mt_state;
    SamplerParams sp;
};
__global__ void __launch_bounds__(kSamplerThreads, 1) k_sample_indices_multi(const MultiSampler *items, int rounds, int round0) {
    extern __shared__ __align__(16) unsigned char dyn[];
    __shared__ SamplerState S;
    MultiSampler it = items[blockIdx.x];
    if (it.sp.out_slot) it.sp.out_slot += (size_t)round0 * it.sp.k;
    if (it.sp.out_logical) it.sp.out_logical += (size_t)round0 * it.sp.k;
    sampler_init(S, it.mt_state, dyn, it.sp);
    sampler_advance(S, dyn, it.sp, rounds);
    __syncthreads();
    sampler_store(S, it.mt_state);
}

}  // namespace

static bool tc_eligible(const prl_dqn_cfg &c, int batch) {
    const bool a_ok = c.n_actions == 1 || c.n_actions == 2 || c.n_actions == 4 || c.n_actions == 8 || c.n_actions == 16;
    return c.hidden1 == HID && c.hidden2 == HID && c.obs_dim % 8 == 0 && c.obs_dim >= 8 && c.obs_dim <= 128 && a_ok &&
           !c.double_dqn && (batch == 128 || batch == 256);
}

extern "C" int prl_dqn_tc_supported(const prl_dqn *q, int batch) { return q && q->tc_tiles && tc_eligible(q->cfg, batch) ? 1 : 0; }

extern "C" int prl_dqn_learn_multi(prl_dqn *const *dqns, prl_buf *const *bufs, int count, int rounds, int batch,
                                   const int64_t *training_steps0, float *const *out_mae, float *const *out_q,
                                   float *const *out_y, int32_t *const *out_logical, void *stream_) {
    PRL_REQUIRE(dqns && bufs && training_steps0 && out_mae && count > 0, "null / empty argument");
    cudaStream_t stream = (cudaStream_t)stream_;
    prl_dqn *q0 = dqns[0];
    const prl_dqn_cfg &c = q0->cfg;
    if (!tc_eligible(c, batch))
        return fail(PRL_EUNSUPPORTED,
                    "tensor-core learner needs hidden [64,64], obs %% 8 == 0 <= 128, actions in {1,2,4,8,16}, DQN, batch 128/256");
    PRL_REQUIRE(rounds > 0 && rounds <= c.max_rounds && batch <= c.max_batch, "rounds / batch outside the configured maxima");
    PRL_REQUIRE((size_t)count * (sizeof(TcLearner) + sizeof(MultiSampler)) <= 64 * 1024, "too many learners in one group");
    PRL_REQUIRE(count <= q0->sm_count, "more learners than SMs in one launch");
    // per-learner descriptors are staged in pinned memory of learner 0 and copied to its workspace
    static thread_local TcLearner *h_learn = nullptr;
    static thread_local MultiSampler *h_samp = nullptr;
    static thread_local cudaEvent_t h_done = nullptr;
    if (!h_learn) {
        PRL_CUDA(cudaHostAlloc((void **)&h_learn, 32 * 1024, cudaHostAllocDefault));
        PRL_CUDA(cudaHostAlloc((void **)&h_samp, 32 * 1024, cudaHostAllocDefault));
        PRL_CUDA(cudaEventCreateWithFlags(&h_done, cudaEventDisableTiming));
    }
    PRL_CUDA(cudaEventSynchronize(h_done));
    size_t samp_smem = 0;
    for (int i = 0; i < count; i++) {
        prl_dqn *q = dqns[i];
        prl_buf *b = bufs[i];
        PRL_REQUIRE(q && b, "null learner / buffer");
        const prl_dqn_cfg &ci = q->cfg;
        PRL_REQUIRE(ci.obs_dim == c.obs_dim && ci.n_actions == c.n_actions && ci.hidden1 == c.hidden1 &&
                        ci.hidden2 == c.hidden2 && ci.double_dqn == c.double_dqn &&
                        ci.target_update_freq == c.target_update_freq && ci.lr == c.lr && ci.beta1 == c.beta1 &&
                        ci.beta2 == c.beta2 && ci.eps == c.eps && ci.weight_decay == c.weight_decay &&
                        ci.gamma == c.gamma && ci.tau == c.tau && rounds <= ci.max_rounds && batch <= ci.max_batch,
                    "all learners of a group must share one configuration");
        PRL_REQUIRE(q->tc_tiles, "learner %d has no operand-tile workspace", i);
        PRL_REQUIRE((b->desc.flags & PRL_BUF_DISCRETE) && b->desc.obs_dim == c.obs_dim && b->desc.n_actions == c.n_actions,
                    "buffer %d does not match the learner", i);
        PRL_REQUIRE(b->lay.record_words == bufs[0]->lay.record_words && b->lay.off_avail == bufs[0]->lay.off_avail,
                    "buffers must share one record layout");
        // the per-round AdamW scalars (two double pow() per round, as torch evaluates them) depend only on the shared
        // configuration and the step count: learners at the same step read learner 0's copy
        const bool shared_scal = i > 0 && q->adam_step == q0->adam_step;
        int rc = shared_scal ? PRL_OK : prl_dqn_stage_scalars(q, rounds, stream);
        if (rc) return rc;
        size_t sbytes = 0;
        rc = prl_sampler_params(b, batch, &h_samp[i].sp, &sbytes);
        if (rc) return rc;
        if (sbytes > samp_smem) samp_smem = sbytes;
        h_samp[i].mt_state = b->mt_state;
        h_samp[i].sp.out_slot = q->slots;
        h_samp[i].sp.out_logical = out_logical ? out_logical[i] : nullptr;
        TcLearner &L = h_learn[i];
        L.records = b->records; L.slots = q->slots;
        L.tiles = q->tc_tiles;
        L.w = q->w; L.wt = q->wt; L.m = q->m; L.v = q->v; L.vmax = q->vmax;
        L.scal = shared_scal ? q0->scal_dev : q->scal_dev;
        L.out_mae = out_mae[i]; L.out_q = out_q ? out_q[i] : nullptr; L.out_y = out_y ? out_y[i] : nullptr;
        L.steps0 = training_steps0[i];
        L.buf_flags = b->desc.flags;
        L.pad_ = 0;
    }
    TcLearner *d_learn = reinterpret_cast<TcLearner *>(q0->multi_dev);
    MultiSampler *d_samp = reinterpret_cast<MultiSampler *>(reinterpret_cast<char *>(q0->multi_dev) + 32 * 1024);
    PRL_CUDA(cudaMemcpyAsync(d_learn, h_learn, (size_t)count * sizeof(TcLearner), cudaMemcpyHostToDevice, stream));
    PRL_CUDA(cudaMemcpyAsync(d_samp, h_samp, (size_t)count * sizeof(MultiSampler), cudaMemcpyHostToDevice, stream));
    PRL_CUDA(cudaEventRecord(h_done, stream));

    PRL_CUDA(cudaFuncSetAttribute(k_sample_indices_multi, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));

    TcArgs a;
    a.learners = d_learn;
    a.lay = bufs[0]->lay;
    a.d = q0->d;
    a.B = batch; a.rounds = rounds; a.freq = c.target_update_freq; a.round0 = 0;
    a.decay = (float)(1.0 - c.lr * c.weight_decay);
    a.omb1 = (float)(1.0 - c.beta1);
    a.beta2 = (float)c.beta2;
    a.omb2 = (float)(1.0 - c.beta2);
    a.eps = (float)c.eps;
    a.gamma = (float)c.gamma;
    a.tau = (float)c.tau;
    a.omtau = (float)(1.0 - c.tau);
    a.inv_b2 = 2.0f / (float)batch;
    a.prof = q0->prof;
    { static const int c = [] { const char *e = getenv("PRL_TC_PROF_CTA"); return e ? atoi(e) : 0; }(); a.prof_cta = c; }
    const size_t smem = MISC_OFF + sizeof(Misc);
    PRL_REQUIRE(smem <= (size_t)q0->max_smem, "tensor-core learner needs %zu B of shared memory", smem);
    PRL_CUDA(cudaFuncSetAttribute(k_dqn_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));

    // The index streams are produced in chunks of rounds: chunk 0 ahead of the first learner launch, chunk c + 1 on a
    // side stream WHILE the learners run chunk c (one learner CTA fills an SM, so the producers of the next chunk run
    // on the SMs the group leaves free; with no spare SM they simply run between the learner launches).
    const int spare = q0->sm_count - count;
    int nchunks = 1;
    if (!q0->prof && spare >= 2 && rounds >= 64) nchunks = rounds >= 256 ? 4 : 2;
    static thread_local cudaStream_t side = nullptr;
    static thread_local cudaEvent_t ev_idx[8], ev_fork = nullptr;
    if (nchunks > 1 && !side) {
        PRL_CUDA(cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking));
        for (auto &e : ev_idx) PRL_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        PRL_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
    }
    // chunk 0 is the only one whose index production is NOT hidden behind a learner launch: keep it short (32 rounds)
    const int first = nchunks > 1 ? (rounds / nchunks < 32 ? rounds / nchunks : 32) : rounds;
    auto chunk_begin = [&](int cix) {
        return cix == 0 ? 0 : first + (int)((long long)(rounds - first) * (cix - 1) / (nchunks > 1 ? nchunks - 1 : 1));
    };
    k_sample_indices_multi<<<count, kSamplerThreads, samp_smem, stream>>>(d_samp, chunk_begin(1), 0);
    PRL_CUDA(cudaGetLastError());
    if (nchunks > 1) {
        PRL_CUDA(cudaEventRecord(ev_fork, stream));
        PRL_CUDA(cudaStreamWaitEvent(side, ev_fork, 0));
    }
    if (q0->timing) PRL_CUDA(cudaEventRecord(q0->t0, stream));
    for (int cix = 0; cix < nchunks; cix++) {
        const int r0 = chunk_begin(cix), r1 = chunk_begin(cix + 1);
        if (cix > 0) PRL_CUDA(cudaStreamWaitEvent(stream, ev_idx[cix], 0));
        a.round0 = r0; a.rounds = r1 - r0;
        k_dqn_tc<<<count, NTH, smem, stream>>>(a);
        PRL_CUDA(cudaGetLastError());
        if (cix + 1 < nchunks) {
            const int n0 = r1, n1 = chunk_begin(cix + 2);
            k_delay<<<1, 1, 0, side>>>(30000u);
            k_sample_indices_multi<<<count, kSamplerThreads, samp_smem, side>>>(d_samp, n1 - n0, n0);
            PRL_CUDA(cudaGetLastError());
            PRL_CUDA(cudaEventRecord(ev_idx[cix + 1], side));
        }
    }
    if (q0->timing) PRL_CUDA(cudaEventRecord(q0->t1, stream));
    for (int i = 0; i < count; i++) {
        dqns[i]->adam_step += rounds;
        dqns[i]->last_launches = 3 * nchunks - 1;   // index producers + learner launches + the side-stream delays
        dqns[i]->last_ctas = 1;
        dqns[i]->last_rows = batch;
    }
    return PRL_OK;
}
