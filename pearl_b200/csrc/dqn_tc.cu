// dqn_tc.cu — tensor-core DQN learner: ONE SM (one CTA) runs one complete learner, every dense
// contraction of the step on tcgen05 (3xTF32 UMMA, fp32 accumulators in TMEM), so that a launch
// with L CTAs trains L independent learners (seeds / agents) concurrently — the aggregate mode that
// fills a B200.
//
// Same semantics as the cooperative SIMT kernel in dqn.cu (DeepQLearning.learn: sample -> Q(s,a) ->
// max_a' Q_target(s',a') -> MSE -> backward -> AdamW(amsgrad) -> scheduled soft target update; reference
// call sites in include/pearl_b200.h).  Shape class: two hidden layers of 64, obs % 8 == 0, obs <= 128,
// n_actions <= 16, batch in {128, 256}; everything else stays on the SIMT kernel.
//
// Warp-specialised CTA of 9 warps:
//   warps 0-3 / 4-7  two ROW GROUPS: group g owns batch rows [128 g, 128 g + 128) for the whole step; thread
//                    (g, m) is batch row m of that tile = TMEM lane m.  The groups are independent pipelines
//                    (own A operand + own accumulator in tensor memory, own mbarriers) that meet only at the
//                    weight-gradient phase, so the tensor pipe always has the other group's products to run
//                    while one group is in SIMT code.  The elected lane of a group issues that group's MMAs.
//   warp 8           LOADER: TMA bulk copies (cp.async.bulk + mbarrier complete_tx) of the weight tiles and of
//                    the raw state rows of the weight-gradient passes.
// Data movement is TMA throughout: every sampled record row reaches shared memory as 128-byte bulk copies
// issued by the lane that owns the row into a warp-private double buffer (issued one phase ahead, the next
// round's first chunks during AdamW), and the weights are kept in global memory a second time IN THE UMMA
// OPERAND LAYOUT (hi = the fp32 value, lo = x - trunc_tf32(x), stacked [hi ; lo] along N), written by AdamW /
// the soft target update, so staging a network is three bulk copies instead of a SIMT pass.
//
// Per gradient step:
//   phase T  T1 = S' W1s_t^T (rows -> TMEM, TS products); then for every action slot one 128 x 64 x 64 product
//            whose A row m is relu(T1[m] + W1_t[:, obs + a] + b1) built in registers.  B is the STACKED tile
//            [W_hi ; W_lo] (N = 128): A_hi x [W_hi ; W_lo] is one N = 128 instruction (the N = 64 form is
//            issue-bound at 53 clk against a 32-clk pipe floor, N = 128 runs at the floor), A_lo x W_hi one
//            N = 64 instruction: 2 instead of 3 issues per K step; the epilogue adds the two accumulator halves.
//   phase O  online forward on the group's tile, MSE gradient, dZ2, dH1 = dZ2 W2 (B = stacked W2^T);
//            h1 and dZ1 are parked in the group's TMEM columns.
//   grads    contractions over the batch as M = 128 products on TRANSPOSED tiles (144-byte chunk pitch =
//            bank-conflict-free column scatter), 32 batch rows (one warp) per pass:
//              G1 = S^T dZ1                        -> dW1[:, :obs]^T
//              G2 = [E ; H1 ; 0]^T [dZ2 | dZ1]     -> dW2^T, db2, db1, dW1[:, obs:]^T   (E = [1 | onehot(action)])
//            The pass's warp builds the tiles from its own TMEM lanes and issues the products itself; two
//            arena halves (G2 operands / G1 operands) are chained through mbarriers so that pass p + 1 is
//            being built while pass p multiplies.
//   AdamW    straight from the TMEM accumulators (lane = input index: coalesced L2 accesses), new weights
//            written to the flat vector AND to the operand-layout tiles.
#include <math.h>
#include <stdarg.h>

#include <new>

#include "dqn_common.cuh"
#include "sampler.cuh"
#include "umma.cuh"

using namespace prl;

int prl_sampler_params(const prl_buf *b, int k, prl::SamplerParams *sp, size_t *smem_bytes);

namespace {

constexpr int NROW = 256;            // row threads: 2 groups x 4 warps
constexpr int NTH = 384;             // 8 row warps + one helper warpgroup: loader, two MMA issuers, tile builder
constexpr int HID = 64;
// ---- shared memory map (bytes)
constexpr int R_W1 = 0;              // 64 KB: stacked W1[:, :obs] tile [128][obs] (target, then online); grads: G1 operands
constexpr int R_W2 = 65536;          // 32 KB: stacked online W2 [128][64];                          grads: raw rows, buffer 0
constexpr int STAGE = 98304;         // 72 KB: 8 warps x 2 x [32 rows][36 floats] row chunks;       grads: G2 operands
constexpr int R_W2T = 172032;        // 32 KB: stacked target W2 (phase T) / online W2^T (phase O); grads: raw rows, buffer 1
constexpr int MISC_OFF = 204800;
constexpr int SROW = 36;             // floats per staged row chunk (32 + 4: conflict-free 128-bit reads)
constexpr int SBUF = 32 * SROW * 4;  // 4608 B per (warp, buffer)
// transposed tiles of the weight-gradient passes: 128 (M or N) x 32 batch rows, chunk pitch 144 B
constexpr int TL = 144, TSBO = 8 * TL, TTILE = 16 * TSBO;   // 18432 B
constexpr int H1_EHT_HI = STAGE, H1_EHT_LO = STAGE + TTILE, H1_DZ_HI = STAGE + 2 * TTILE, H1_DZ_LO = STAGE + 3 * TTILE;
constexpr int H2_ST_HI = R_W1, H2_ST_LO = R_W1 + TTILE;
constexpr int RAWP = 132;            // floats per raw row (128 + 4)
// ---- tensor memory columns
constexpr int TM_A0 = 0, TM_S0 = 256;          // group g: A operand at 128 g (hi 64 | lo 64), accumulators / parking at 256 + 128 g
constexpr int TM_G1 = 0, TM_G2 = 128;          // weight-gradient accumulators (the A regions are free by then)

// mbarrier indices
enum {
    B_W1 = 0, B_W2, B_W2T, B_MMA /*2*/ = 3, B_W1FREE = 5, B_ACTDONE, B_RFULL /*2*/ = 7, B_RFREE /*2*/ = 9,
    // one barrier PER PASS for the two arena chains: a parity wait only tells adjacent phases apart, and pass p + 2 may
    // reach its wait before pass p has even committed
    B_H1FREE /*8*/ = 11, B_H2FREE /*8*/ = 19, B_READY /*2*/ = 27, B_DZREADY /*8*/ = 29, B_COUNT = 37
};

struct TcLearner {            // one per CTA, in global memory
    const uint32_t *records;
    const int32_t *slots;     // [rounds][B]
    float *w, *wt, *m, *v, *vmax;
    float *tiles;             // operand-layout copies: online [W1s | W2 | W2^T], target [W1s | W2]
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
    long long *prof;   // optional [rounds][16] SM-clock stamps of CTA 0
};

#define TC_STAMP(idx)                                                                          \
    do {                                                                                       \
        if (a.prof && blockIdx.x == 0 && threadIdx.x == 0) a.prof[(size_t)round * 16 + (idx)] = clock64(); \
    } while (0)

struct Smalls {              // the small fp32 vectors of one network
    float watb[16][68];      // W1[:, obs + a] + b1, row pitch 68: rows of different action ids fall in different banks
    float b2[HID], w3[HID];
    float b3, pad[3];
};
struct Misc {
    Smalls t, o;
    float redw[8][HID];
    float redmae[8], reddb3[8];
    unsigned long long rowptr[8][32];   // this round's record row of every row thread (the warp's lanes copy each other's rows)
    unsigned long long bar[B_COUNT + 1];
    uint32_t tmem_base;
};

// ------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(umma::smem_u32(bar)) : "memory");
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
__device__ __forceinline__ void group_sync(int g) { asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "r"(128) : "memory"); }
__device__ __forceinline__ void rows_sync() { asm volatile("bar.sync 3, 256;" ::: "memory"); }
__device__ __forceinline__ void cta_sync() { __syncthreads(); }

__device__ __forceinline__ float tf32_lo(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

__device__ __forceinline__ float fast_sqrt(float x) {
    float r;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
struct AdamScalarsTc : AdamScalars { float inv_bc2_sqrt; };
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

// ------------------------------------------------------------------------------------------ operand-layout weight tiles
// float offsets inside TcLearner::tiles
struct TileOff { int oW1, oW2, oW2T, tW1, tW2, total; };
__host__ __device__ inline TileOff tile_offsets(int obs) {
    TileOff t;
    t.oW1 = 0; t.oW2 = 128 * obs; t.oW2T = t.oW2 + 128 * HID; t.tW1 = t.oW2T + 128 * HID; t.tW2 = t.tW1 + 128 * obs;
    t.total = t.tW2 + 128 * HID;
    return t;
}
// stacked tile: rows [0, 64) hold the value (the tensor core truncates it to TF32), rows [64, 128) the residual
__device__ __forceinline__ void put_stacked(float *tile, int r, int k, int K, float x) {
    tile[umma::tile_index(r, k, K)] = x;
    tile[umma::tile_index(64 + r, k, K)] = tf32_lo(x);
}
// all tiles of one network from its flat parameter vector (kernel prologue, soft target update)
__device__ void rebuild_tiles(const float *__restrict__ net, const Dims &d, float *tW1, float *tW2, float *tW2T, int tid) {
    for (int e = tid; e < HID * d.obs; e += NROW) {
        const int j = e / d.obs, k = e - j * d.obs;
        put_stacked(tW1, j, k, d.obs, __ldcg(net + d.oW1 + (size_t)j * d.D + k));
    }
    for (int e = tid; e < HID * HID; e += NROW) {
        const int j = e >> 6, k = e & 63;
        const float x = __ldcg(net + d.oW2 + e);
        put_stacked(tW2, j, k, HID, x);
        if (tW2T) put_stacked(tW2T, k, j, HID, x);
    }
}
__device__ void load_smalls(const float *__restrict__ net, const Dims &d, Smalls &s, int tid) {
    for (int e = tid; e < d.A * HID; e += NROW) {
        const int j = e / d.A, ac = e - j * d.A;
        s.watb[ac][j] = __ldcg(net + d.oW1 + (size_t)j * d.D + d.obs + ac) + __ldcg(net + d.ob1 + j);
    }
    if (tid < HID) { s.b2[tid] = __ldcg(net + d.ob2 + tid); s.w3[tid] = __ldcg(net + d.oW3 + tid); }
    if (tid == 0) s.b3 = __ldcg(net + d.ob3);
}

// ------------------------------------------------------------------------------------------ per-thread pipeline state
struct RowCtx {
    int g, q, lane, warp, m;       // group, warp within the group (TMEM lane quarter), lane, warp id, row within the tile
    uint32_t tm, tlane;            // TMEM base, this warp's lane window
    uint32_t a_col, s_col;         // the group's A-operand and accumulator columns
    uint32_t n_chunk;              // staged row chunks consumed so far by this warp (running, all phases and rounds)
    uint32_t n_issued;             // staged row chunks requested so far
    uint32_t n_mma;                // MMA batches committed so far on the group's barrier
    bool elected;                  // issues the group's MMAs
    float *sbuf;                   // this warp's two staging buffers
    uint64_t *bar;
    char *smem;
};

// Chunk `cc` (32 columns) of the field at `field_off` of this warp's 32 rows -> buffer (n & 1): 16-byte cp.async, 8 lanes
// per row (one full 128-byte line), 4 rows per instruction; one cp.async group per chunk.  (Per-row TMA bulk copies were
// measured first: ~2300 128-byte operations per round serialise in the SM's TMA unit and tripled the round time; TMA is
// kept for what it is good at here, the 16 - 64 KB weight tiles.)
__device__ __forceinline__ void issue_chunk(RowCtx &c, uint32_t n, const unsigned long long *rowptr, int field_off, int cc, int obs) {
    const int kc = min(32, obs - 32 * cc);
    float *dst = c.sbuf + (n & 1) * (SBUF / 4);
    const int c16 = c.lane & 7, rsub = c.lane >> 3;
    if (4 * c16 < kc) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int row = 4 * i + rsub;
            cp_async16(dst + row * SROW + 4 * c16, reinterpret_cast<const float *>(rowptr[row]) + field_off + 32 * cc + 4 * c16);
        }
    }
    cp_async_commit();
    c.n_issued = n + 1;
}
// chunk n of this warp has landed (at most one younger chunk may still be in flight)
__device__ __forceinline__ void wait_chunk(const RowCtx &c, uint32_t n) {
    if (c.n_issued > n + 1) cp_async_wait<1>(); else cp_async_wait<0>();
    __syncwarp();
}

// hand the A operand (or any TMEM state) of this warp's rows to the group's issuer: 4 warp arrivals complete a phase
__device__ __forceinline__ void publish_a(const RowCtx &c) {
    umma::tmem_st_wait();
    umma::fence_before_thread_sync();
    __syncwarp();
    if (c.lane == 0) mbar_arrive(c.bar + B_READY + c.g);
}
__device__ __forceinline__ void wait_group_mma(RowCtx &c) {
    umma::mbar_wait(c.bar + B_MMA + c.g, (c.n_mma - 1) & 1);
    umma::fence_after_thread_sync();
}

// Layer 1 of the group's tile: acc[128 rows x 128 stacked columns] = X W1s^T, X = the staged field of the sampled rows.
// Chunks of 32 columns arrive by TMA in the warp's private double buffer; the row thread splits its chunk and
// writes it to tensor memory (TS product: A never exists in shared memory in operand layout).  Two chunks (K = 64)
// per MMA batch.  `request_ahead(n)`: issue the copies of the warp's chunk number n (two ahead of the one just read).
template <typename NextFn>
__device__ __forceinline__ void layer1(RowCtx &c, int obs, NextFn request_ahead) {
    const int nch = (obs + 31) >> 5;
    bool pending = false;
    for (int cc = 0; cc < nch; cc++) {
        const uint32_t n = c.n_chunk++;
        const int kc = min(32, obs - 32 * cc);
        wait_chunk(c, n);
        const float *row = c.sbuf + (n & 1) * (SBUF / 4) + c.lane * SROW;
        float hi[32];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (4 * i < kc) v = *reinterpret_cast<const float4 *>(row + 4 * i);
            hi[4 * i + 0] = v.x; hi[4 * i + 1] = v.y; hi[4 * i + 2] = v.z; hi[4 * i + 3] = v.w;
        }
        __syncwarp();
        request_ahead(n + 2);                       // this buffer is free again
        if (pending && (cc & 1) == 0) {             // the previous batch still reads the A operand
            wait_group_mma(c);
            pending = false;
        }
        {
            float lo[32];
#pragma unroll
            for (int i = 0; i < 32; i++) lo[i] = tf32_lo(hi[i]);
            umma::tmem_st32(c.tlane + c.a_col + 32 * (cc & 1), hi);
            umma::tmem_st32(c.tlane + c.a_col + 64 + 32 * (cc & 1), lo);
        }
        if ((cc & 1) || cc == nch - 1) {
            publish_a(c);                           // the group's issuer multiplies this K batch (stacked W1 tile)
            c.n_mma++;
            pending = true;
        }
    }
    wait_group_mma(c);
}

// ---- issuer side (one lane of a helper warp per row group) -------------------------------------------------------------
struct IssueCtx { uint32_t tm, a_col; char *smem; uint64_t *done; };
// K batch `b` (64 columns) of layer 1: acc[128 stacked columns] (+)= A [W_hi ; W_lo]^T : hi x stacked (N = 128), lo x W_hi (N = 64)
__device__ __forceinline__ void issue_layer1(const IssueCtx &x, int obs, int b, uint32_t acc_col) {
    const umma::Tile W1 = umma::make_tile(x.smem + R_W1, obs, 128);
    const int ksteps = min(64, obs - 64 * b) >> 3;
    const uint32_t i128 = umma::make_idesc_tf32(128, 128), i64 = umma::make_idesc_tf32(128, 64);
    uint64_t bd = W1.shifted((uint32_t)(16 * b) * 128).desc(0);
    for (int ks = 0; ks < ksteps; ks++) {
        umma::mma_tf32_ts(x.tm + acc_col, x.tm + x.a_col + ks * 8, bd, i128, b > 0 || ks > 0);
        umma::mma_tf32_ts(x.tm + acc_col, x.tm + x.a_col + 64 + ks * 8, bd, i64, true);
        bd += (uint64_t)((2 * 128) >> 4);
    }
    umma::mma_commit(x.done);
}
// one 128 x 64 x 64 product, A (hi | lo) in the group's TMEM A columns, B = the stacked [hi ; lo] tile at `region`:
// acc[64 columns at acc_col] = A_lo W_hi + A_hi W_lo + A_hi W_hi
__device__ __forceinline__ void issue_hidden(const IssueCtx &x, int region, uint32_t acc_col) {
    const umma::Tile Wt = umma::make_tile(x.smem + region, HID, 128);
    umma::gemm3_ts(x.tm + acc_col, x.tm + x.a_col, x.tm + x.a_col + 64, Wt, Wt.rows_from(64), 128, HID, HID, false);
    umma::mma_commit(x.done);
}

// the same product against the stacked tile as ONE N = 128 issue (A_hi x [W_hi ; W_lo]) plus one N = 64 issue (A_lo x W_hi)
// per K step: acc[128 columns] = [A_hi W_hi + A_lo W_hi | A_hi W_lo]
__device__ __forceinline__ void issue_hidden_stacked(const IssueCtx &x, int region, uint32_t acc_col) {
    const umma::Tile Wt = umma::make_tile(x.smem + region, HID, 128);
    const uint32_t i128 = umma::make_idesc_tf32(128, 128), i64 = umma::make_idesc_tf32(128, 64);
    uint64_t bd = Wt.desc(0);
#pragma unroll
    for (int ks = 0; ks < 8; ks++) {
        umma::mma_tf32_ts(x.tm + acc_col, x.tm + x.a_col + ks * 8, bd, i128, ks > 0);
        umma::mma_tf32_ts(x.tm + acc_col, x.tm + x.a_col + 64 + ks * 8, bd, i64, true);
        bd += (uint64_t)((2 * 128) >> 4);
    }
    umma::mma_commit(x.done);
}

// stacked K-major operand tile [hi ; lo] of W[rows = 64 outputs][K] in shared memory from the flat fp32 weights
// (row pitch `ld` floats); `transpose`: the tile of W^T (K = 64 outputs of W, rows = inputs).  One warp; lane = (row % 8,
// chunk % 4): conflict-free 16-byte shared stores, 64-byte global segments.
__device__ void build_tile_warp(float *tile, const float *__restrict__ W, int ld, int K, bool vec, int lane) {
    if (vec) {
        const int rl = lane & 7, cl = lane >> 3, cg = (K + 15) >> 4;     // groups of 4 chunks along K
        for (int it = 0; it < 8 * cg; it += 4) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int i2 = it + u, r = (i2 / cg) * 8 + rl, k = ((i2 % cg) * 4 + cl) * 4;
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (i2 < 8 * cg && k < K) v[u] = __ldcg(reinterpret_cast<const float4 *>(W + (size_t)r * ld + k));
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int i2 = it + u, r = (i2 / cg) * 8 + rl, k = ((i2 % cg) * 4 + cl) * 4;
                if (i2 < 8 * cg && k < K) {
                    *reinterpret_cast<float4 *>(tile + umma::tile_index(r, k, K)) = v[u];
                    *reinterpret_cast<float4 *>(tile + umma::tile_index(64 + r, k, K)) =
                        make_float4(tf32_lo(v[u].x), tf32_lo(v[u].y), tf32_lo(v[u].z), tf32_lo(v[u].w));
                }
            }
        }
    } else {
        for (int e = lane; e < 64 * K; e += 32) {
            const int r = e / K, k = e - r * K;
            const float x = __ldcg(W + (size_t)r * ld + k);
            tile[umma::tile_index(r, k, K)] = x;
            tile[umma::tile_index(64 + r, k, K)] = tf32_lo(x);
        }
    }
}
// the tile of W2^T: element (r = input k, column = output j) = W2[j][k]; coalesced global reads along k, scattered stores
__device__ void build_w2t_warp(float *tile, const float *__restrict__ W2, int lane) {
    for (int e0 = lane; e0 < HID * HID; e0 += 32 * 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) v[u] = __ldcg(W2 + e0 + 32 * u);
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int e = e0 + 32 * u, j = e >> 6, k = e & 63;
            tile[umma::tile_index(k, j, HID)] = v[u];
            tile[umma::tile_index(64 + k, j, HID)] = tf32_lo(v[u]);
        }
    }
}

// registers -> the group's A operand (hi | lo), 32 columns at `col`
__device__ __forceinline__ void put_a32(const RowCtx &c, int col, const float *hi) {
    float lo[32];
#pragma unroll
    for (int i = 0; i < 32; i++) lo[i] = tf32_lo(hi[i]);
    umma::tmem_st32(c.tlane + c.a_col + col, hi);
    umma::tmem_st32(c.tlane + c.a_col + 64 + col, lo);
}

// float offset of element (r, column = batch row `col` of the pass) inside a transposed 128 x 32 tile
__device__ __forceinline__ int tt_off(int r, int colbase) { return (r >> 3) * (TSBO / 4) + (r & 7) * 4 + colbase; }

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
    const int ntiles = a.B >> 7;
    const int W = a.lay.record_words;
    const int obs = d.obs, nch = (obs + 31) >> 5;
    const TileOff to = tile_offsets(obs);
    uint64_t *bar = reinterpret_cast<uint64_t *>(mi.bar);
    const uint32_t w1_bytes = 128u * obs * 4, w2_bytes = 128u * HID * 4;

    // De-phase the learners: identical CTAs started together would run in lock-step and hit L2 / HBM with their AdamW
    // sweeps (432 KB each) and row gathers all at once; a start offset of up to ~50 us spreads those bursts over the round.
    if (tid == 0 && gridDim.x > 1) {
        const long long t0 = clock64(), wait = (long long)(blockIdx.x % 48) * 2000;
        while (clock64() - t0 < wait) __nanosleep(100);
    }
    if (warp == 0) umma::tmem_alloc(&mi.tmem_base, 512);
    if (tid == 0)
        for (int i = 0; i < B_COUNT; i++)
            umma::mbar_init(bar + i, (i == B_W1FREE || i == B_ACTDONE) ? ntiles : (i == B_RFULL || i == B_RFULL + 1) ? 32
                                    : (i == B_READY || i == B_READY + 1) ? 4 : (i >= B_H1FREE && i < B_H1FREE + 8) ? 2 : 1);
    // prologue: the operand-layout tiles and the small vectors follow the flat parameters (which the host may have
    // changed between calls)
    if (warp < 8) {
        rebuild_tiles(L.wt, d, L.tiles + to.tW1, L.tiles + to.tW2, nullptr, tid);
        load_smalls(L.wt, d, mi.t, tid);
        load_smalls(L.w, d, mi.o, tid);
        mi.redw[warp][lane] = 0.f; mi.redw[warp][32 + lane] = 0.f;
        if (lane == 0) { mi.redmae[warp] = 0.f; mi.reddb3[warp] = 0.f; }
        fence_proxy_async_all();
    }
    umma::fence_before_thread_sync();
    __syncthreads();
    umma::fence_after_thread_sync();

    const int np = 4 * ntiles;                                   // weight-gradient passes per round (32 rows each)
    uint32_t n_pass = 0;                                         // running pass counter (identical in every thread)

    if (warp >= 8) {
        // ================================================================== helper warpgroup
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        auto soft_due = [&](int round) { return (L.steps0 + round + 2) % a.freq == 0; };
        if (warp == 8) {
            // ---- LOADER: target weight tiles by TMA, raw state rows of the weight-gradient passes by cp.async
            for (int round = 0; round < a.rounds; round++) {
                cta_sync();   // (A)
                int pslot[8];
#pragma unroll
                for (int p = 0; p < 8; p++) pslot[p] = p < np ? L.slots[(size_t)round * a.B + p * 32 + lane] : 0;
                if (lane == 0) {
                    mbar_expect_tx(bar + B_W1, w1_bytes);  bulk_g2s(smem + R_W1, L.tiles + to.tW1, w1_bytes, bar + B_W1);
                    mbar_expect_tx(bar + B_W2T, w2_bytes); bulk_g2s(smem + R_W2T, L.tiles + to.tW2, w2_bytes, bar + B_W2T);
                }
                __syncwarp();
                cta_sync();   // (B) phase O done: the weight regions become the arena of the weight-gradient passes
#pragma unroll
                for (int p = 0; p < 8; p++) {
                    if (p < np) {
                        const uint32_t n = n_pass + p, b = n & 1;
                        umma::mbar_wait(bar + B_RFREE + b, ((n >> 1) & 1) ^ 1);
                        float *raw = reinterpret_cast<float *>(smem + (b ? R_W2T : R_W2));
                        const int q4 = obs >> 2;                       // 16-byte chunks per row
                        for (int idx = lane; idx < 32 * q4; idx += 32) {
                            const int row = idx / q4, ch = idx - row * q4;
                            const int slot = __shfl_sync(0xffffffffu, pslot[p], row);
                            cp_async16(raw + row * RAWP + 4 * ch, reinterpret_cast<const float *>(L.records + (size_t)slot * W) + a.lay.off_state + 4 * ch);
                        }
                        // the barrier's 32 arrivals fire as each lane's copies land
                        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(umma::smem_u32(bar + B_RFULL + b)) : "memory");
                    }
                }
                n_pass += np;
            }
        } else if (warp == 11) {
            // ---- TILE BUILDER: the online network's operand tiles straight from the flat weights AdamW just wrote
            const bool vec1 = (d.D & 3) == 0 && (reinterpret_cast<uintptr_t>(L.w) & 15) == 0, vec2 = (d.oW2 & 3) == 0 && vec1;
            for (int round = 0; round < a.rounds; round++) {
                cta_sync();   // (A): AdamW's writes are visible
                build_tile_warp(reinterpret_cast<float *>(smem + R_W2), L.w + d.oW2, HID, HID, vec2, lane);
                umma::fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar + B_W2);
                umma::mbar_wait(bar + B_W1FREE, round & 1);       // every group is through target layer 1
                build_tile_warp(reinterpret_cast<float *>(smem + R_W1), L.w + d.oW1, d.D, obs, vec1, lane);
                umma::fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar + B_W1);
                umma::mbar_wait(bar + B_ACTDONE, round & 1);      // ... and through the all-actions products (target W2 is free)
                build_w2t_warp(reinterpret_cast<float *>(smem + R_W2T), L.w + d.oW2, lane);
                umma::fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar + B_W2T);
                cta_sync();   // (B)
            }
        } else {
            // ---- ISSUER of row group g: the group's products in program order, one lane; the row warps only publish
            //      operands (B_READY) and wait for results (B_MMA), so the ~50-60 clk per issued MMA never blocks a row warp
            const int g = warp - 9;
            IssueCtx x;
            x.tm = mi.tmem_base; x.a_col = TM_A0 + 128 * g; x.smem = smem; x.done = bar + B_MMA + g;
            const uint32_t s_col = TM_S0 + 128 * g;
            uint32_t n_ready = 0;
            const int nb = (nch + 1) >> 1;
            for (int round = 0; round < a.rounds; round++) {
                cta_sync();   // (A)
                if (g < ntiles && lane == 0) {
                    auto ready = [&]() { umma::mbar_wait(bar + B_READY + g, n_ready & 1); n_ready++; umma::fence_after_thread_sync(); };
                    for (int b = 0; b < nb; b++) {                       // target layer 1
                        ready();
                        if (b == 0) umma::mbar_wait(bar + B_W1, 0);
                        issue_layer1(x, obs, b, s_col);
                    }
                    for (int ac = 0; ac < d.A; ac++) {                   // all-actions products (stacked: 2 issues per K step)
                        ready();
                        if (ac == 0) umma::mbar_wait(bar + B_W2T, 0);
                        issue_hidden_stacked(x, R_W2T, s_col);
                    }
                    for (int b = 0; b < nb; b++) {                       // online layer 1
                        ready();
                        if (b == 0) umma::mbar_wait(bar + B_W1, 1);
                        issue_layer1(x, obs, b, s_col);
                    }
                    ready();                                             // layer 2
                    umma::mbar_wait(bar + B_W2, round & 1);
                    issue_hidden(x, R_W2, s_col + 64);
                    ready();                                             // dH1 = dZ2 W2
                    umma::mbar_wait(bar + B_W2T, 1);
                    issue_hidden(x, R_W2T, s_col + 64);
                }
                __syncwarp();
                cta_sync();   // (B)
            }
        }
        cta_sync();       // end: everything is done before warp 0 frees the tensor memory
        return;
    }
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");

    RowCtx c;
    c.warp = warp; c.lane = lane; c.g = (warp >> 2) & 1; c.q = warp & 3; c.m = tid & 127;
    c.tm = mi.tmem_base;
    c.tlane = c.tm + ((uint32_t)(c.q * 32) << 16);
    c.a_col = TM_A0 + 128 * c.g; c.s_col = TM_S0 + 128 * c.g;
    c.n_chunk = 0; c.n_issued = 0; c.n_mma = 0;
    c.elected = c.q == 0 && lane == 0;
    c.sbuf = reinterpret_cast<float *>(smem + STAGE + (warp & 7) * 2 * SBUF);
    c.bar = bar; c.smem = smem;
    const bool active = c.g < ntiles;
    const bool dyn = (L.buf_flags & PRL_BUF_DYNAMIC_ACTIONS) != 0;

    // this round's row: slot and scalars are fetched one round ahead
    const uint32_t *rec = L.records;
    int act = 0, cnt = 0;
    float rew = 0.f, term = 0.f;
    uint32_t ids0 = 0, ids1 = 0, ids2 = 0, ids3 = 0;
    auto fetch_row = [&](int slot) {
        rec = L.records + (size_t)slot * W;
        act = (int)rec[a.lay.off_action];
        rew = __uint_as_float(rec[a.lay.off_reward]);
        const uint32_t fl = rec[a.lay.off_flags];
        term = (fl & 1u) ? 1.f : 0.f;
        cnt = (int)((fl >> 8) & 0xffffu);
        if (dyn) {
            ids0 = rec[a.lay.off_avail];
            if (d.A > 4) ids1 = rec[a.lay.off_avail + 1];
            if (d.A > 8) { ids2 = rec[a.lay.off_avail + 2]; ids3 = rec[a.lay.off_avail + 3]; }
        }
        mi.rowptr[warp][lane] = (unsigned long long)rec;
        __syncwarp();
    };
    auto first_chunks = [&](int round) {     // the two chunks every warp requests before a round starts
        const uint32_t n0 = (uint32_t)round * 2 * nch;
        issue_chunk(c, n0, mi.rowptr[warp], a.lay.off_next_state, 0, obs);
        issue_chunk(c, n0 + 1, mi.rowptr[warp], nch > 1 ? a.lay.off_next_state : a.lay.off_state, nch > 1 ? 1 : 0, obs);
    };
    if (active) {
        fetch_row(L.slots[c.g * 128 + c.m]);
        first_chunks(0);
    }

    for (int round = 0; round < a.rounds; round++) {
        TC_STAMP(0);
        // ---- scheduled soft target update happens BEFORE this round's gradient step
        //      (deep_td_learning.py:283-284: (training_steps + 1) % freq == 0)
        if ((L.steps0 + round + 2) % a.freq == 0) {
            for (int i = tid; i < d.P; i += NROW) L.wt[i] = soft_update(__ldcg(L.w + i), __ldcg(L.wt + i), a.tau, a.omtau);
            rows_sync();
            rebuild_tiles(L.wt, d, L.tiles + to.tW1, L.tiles + to.tW2, nullptr, tid);
            load_smalls(L.wt, d, mi.t, tid);
            fence_proxy_async_all();
        }
        umma::fence_before_thread_sync();
        cta_sync();        // (A) parameters, tiles and TMEM of the previous round are settled; the weight regions are free
        umma::fence_after_thread_sync();
        TC_STAMP(1);

        // ---------------------------------------------------------------------- ROW GROUPS
        float dq = 0.f;
        uint32_t mask2a = 0u, mask2b = 0u;
        const int act_now = act;
        if (active) {
            auto ahead = [&](uint32_t n) {   // chunks of a round in order: next_state 0..nch-1, state 0..nch-1
                const int s = (int)(n - (uint32_t)round * 2 * nch);
                if (s < 2 * nch)
                    issue_chunk(c, n, mi.rowptr[warp], s < nch ? a.lay.off_next_state : a.lay.off_state, s < nch ? s : s - nch, obs);
            };
            // ================= phase T: y = max_a' Q_target(s', a') * gamma * (1 - term) + r =================
            layer1(c, obs, ahead);
            float yv;
            {
                float t1[64];
                {
                    float d2[32];
                    umma::tmem_ld32(c.tlane + c.s_col, t1);
                    umma::tmem_ld32(c.tlane + c.s_col + 64, d2);
#pragma unroll
                    for (int i = 0; i < 32; i++) t1[i] += d2[i];
                    umma::tmem_ld32(c.tlane + c.s_col + 32, t1 + 32);
                    umma::tmem_ld32(c.tlane + c.s_col + 96, d2);
#pragma unroll
                    for (int i = 0; i < 32; i++) t1[32 + i] += d2[i];
                }
                if (c.elected) mbar_arrive(bar + B_W1FREE);
                TC_STAMP(2);
                float best = -INFINITY;
                float h[64];                      // the NEXT action slot's layer-1 activations, built while a product runs
                auto build_h = [&](int ac) {
                    int id = ac;
                    if (dyn && ac < cnt) {
                        const uint32_t wsel = ac < 4 ? ids0 : ac < 8 ? ids1 : ac < 12 ? ids2 : ids3;
                        id = (int)((wsel >> (8 * (ac & 3))) & 0xffu);
                    }
#pragma unroll
                    for (int c4 = 0; c4 < 16; c4++) {
                        const float4 wv = *reinterpret_cast<const float4 *>(&mi.t.watb[id][4 * c4]);
                        h[4 * c4 + 0] = fmaxf(t1[4 * c4 + 0] + wv.x, 0.f);
                        h[4 * c4 + 1] = fmaxf(t1[4 * c4 + 1] + wv.y, 0.f);
                        h[4 * c4 + 2] = fmaxf(t1[4 * c4 + 2] + wv.z, 0.f);
                        h[4 * c4 + 3] = fmaxf(t1[4 * c4 + 3] + wv.w, 0.f);
                    }
                };
                auto finish = [&](const float *z, int ap) {     // z = layer-2 pre-activations of slot ap (accumulator halves summed)
                    float qv = 0.f;
#pragma unroll
                    for (int c4 = 0; c4 < 16; c4++) {
                        const float4 w3v = *reinterpret_cast<const float4 *>(&mi.t.w3[4 * c4]);
                        const float4 b2v = *reinterpret_cast<const float4 *>(&mi.t.b2[4 * c4]);
                        qv = fmaf(w3v.x, fmaxf(z[4 * c4 + 0] + b2v.x, 0.f), qv);
                        qv = fmaf(w3v.y, fmaxf(z[4 * c4 + 1] + b2v.y, 0.f), qv);
                        qv = fmaf(w3v.z, fmaxf(z[4 * c4 + 2] + b2v.z, 0.f), qv);
                        qv = fmaf(w3v.w, fmaxf(z[4 * c4 + 3] + b2v.w, 0.f), qv);
                    }
                    qv += mi.t.b3;
                    if (ap >= cnt) qv = -INFINITY;   // next_state_action_values[mask] = -inf
                    best = fmaxf(best, qv);
                };
                auto read_acc = [&](float *z) {
                    float d2[32];
                    umma::tmem_ld32(c.tlane + c.s_col, z);
                    umma::tmem_ld32(c.tlane + c.s_col + 64, d2);
#pragma unroll
                    for (int i = 0; i < 32; i++) z[i] += d2[i];
                    umma::tmem_ld32(c.tlane + c.s_col + 32, z + 32);
                    umma::tmem_ld32(c.tlane + c.s_col + 96, d2);
#pragma unroll
                    for (int i = 0; i < 32; i++) z[32 + i] += d2[i];
                };
                // Per slot the only serial work between two products is: operand registers -> TMEM, accumulator -> registers,
                // publish.  The epilogue arithmetic of slot a - 1 and the operand of slot a + 1 are computed while product a runs.
                build_h(0);
                for (int ac = 0; ac < d.A; ac++) {
                    if (ac == 2) TC_STAMP(13);
                    if (ac > 0) wait_group_mma(c);      // product ac - 1 is done: its A operand is free, its result is ready
                    if (ac == 2) TC_STAMP(14);
                    put_a32(c, 0, h);
                    put_a32(c, 32, h + 32);
                    float z[64];
                    if (ac > 0) read_acc(z);            // before product ac overwrites the accumulator
                    publish_a(c);
                    if (ac == 2) TC_STAMP(15);
                    c.n_mma++;
                    if (ac > 0) finish(z, ac - 1);
                    if (ac + 1 < d.A) build_h(ac + 1);
                }
                wait_group_mma(c);
                {
                    float z[64];
                    read_acc(z);
                    finish(z, d.A - 1);
                }
                if (c.elected) mbar_arrive(bar + B_ACTDONE);
                yv = __fadd_rn(__fmul_rn(__fmul_rn(best, a.gamma), 1.f - term), rew);
            }
            TC_STAMP(3);

            // ================= phase O: online forward, loss, backward through the hidden layers =================
            layer1(c, obs, ahead);
            uint32_t mask1a = 0u, mask1b = 0u;
            {
                float h1[64];
                {
                    float d2[32];
                    umma::tmem_ld32(c.tlane + c.s_col, h1);
                    umma::tmem_ld32(c.tlane + c.s_col + 64, d2);
#pragma unroll
                    for (int i = 0; i < 32; i++) h1[i] += d2[i];
                    umma::tmem_ld32(c.tlane + c.s_col + 32, h1 + 32);
                    umma::tmem_ld32(c.tlane + c.s_col + 96, d2);
#pragma unroll
                    for (int i = 0; i < 32; i++) h1[32 + i] += d2[i];
                }
#pragma unroll
                for (int j = 0; j < 64; j++) {
                    h1[j] = fmaxf(h1[j] + mi.o.watb[act_now][j], 0.f);
                    if (h1[j] > 0.f) { if (j < 32) mask1a |= 1u << j; else mask1b |= 1u << (j - 32); }
                }
                umma::tmem_st32(c.tlane + c.s_col, h1);          // parked for the weight-gradient pass
                umma::tmem_st32(c.tlane + c.s_col + 32, h1 + 32);
                put_a32(c, 0, h1);
                put_a32(c, 32, h1 + 32);
            }
            TC_STAMP(4);
            publish_a(c);
            c.n_mma++;
            wait_group_mma(c);
            {
                float z[64];   // h2, then dZ2, then dZ1
                umma::tmem_ld32(c.tlane + c.s_col + 64, z);
                umma::tmem_ld32(c.tlane + c.s_col + 96, z + 32);
                float part = 0.f;
#pragma unroll
                for (int j = 0; j < 64; j++) { z[j] = fmaxf(z[j] + mi.o.b2[j], 0.f); part = fmaf(mi.o.w3[j], z[j], part); }
                const float qv = part + mi.o.b3;
                dq = (qv - yv) * a.inv_b2;
                const int row = c.g * 128 + c.m;
                if (L.out_q) L.out_q[(size_t)round * a.B + row] = qv;
                if (L.out_y) L.out_y[(size_t)round * a.B + row] = yv;
                {   // dW3[j] = sum_rows dq h2[j]: reduce over this warp's 32 rows, lane j keeps columns j and 32 + j
                    float cs[32];
#pragma unroll
                    for (int j = 0; j < 32; j++) cs[j] = dq * z[j];
                    const float s0 = warp_colsum32(cs, lane);
#pragma unroll
                    for (int j = 0; j < 32; j++) cs[j] = dq * z[32 + j];
                    const float s1 = warp_colsum32(cs, lane);
                    mi.redw[warp][lane] = s0; mi.redw[warp][32 + lane] = s1;
                    const float ms = warp_sum(fabsf(qv - yv)), ds = warp_sum(dq);
                    if (lane == 0) { mi.redmae[warp] = ms; mi.reddb3[warp] = ds; }
                }
#pragma unroll
                for (int j = 0; j < 64; j++) {
                    const bool on = z[j] > 0.f;
                    if (on) { if (j < 32) mask2a |= 1u << j; else mask2b |= 1u << (j - 32); }
                    z[j] = on ? dq * mi.o.w3[j] : 0.f;      // dZ2
                }
                put_a32(c, 0, z);
                put_a32(c, 32, z + 32);
                publish_a(c);                               // dH1 = dZ2 W2
                c.n_mma++;
                wait_group_mma(c);
                umma::tmem_ld32(c.tlane + c.s_col + 64, z);
                umma::tmem_ld32(c.tlane + c.s_col + 96, z + 32);
#pragma unroll
                for (int j = 0; j < 64; j++) {
                    const bool on = j < 32 ? ((mask1a >> j) & 1u) : ((mask1b >> (j - 32)) & 1u);
                    z[j] = on ? z[j] : 0.f;                 // dZ1
                }
                umma::tmem_st32(c.tlane + c.s_col + 64, z);   // parked for the weight-gradient pass
                umma::tmem_st32(c.tlane + c.s_col + 96, z + 32);
            }
        }
        TC_STAMP(5);
        umma::tmem_st_wait();
        umma::fence_before_thread_sync();
        cta_sync();        // (B) every product of phase O is complete: A columns -> gradient accumulators, weight regions -> arena
        umma::fence_after_thread_sync();
        {   // constant-zero parts of the transposed operand tiles (the arena was weights / row chunks until now)
            const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
            auto zero = [&](int byte_off, int bytes) {
                for (int i = tid * 16; i < bytes; i += NROW * 16) *reinterpret_cast<float4 *>(smem + byte_off + i) = z4;
            };
            zero(H1_EHT_HI + 2 * TSBO, 2 * TSBO);      // E rows 16..31 (row 16 is rewritten by the passes)
            zero(H1_EHT_HI + 12 * TSBO, 4 * TSBO);     // rows 96..127
            zero(H1_EHT_LO, 4 * TSBO);                 // E is exact: no residual
            zero(H1_EHT_LO + 12 * TSBO, 4 * TSBO);
            if (obs < 128) {
                zero(H2_ST_HI + (obs >> 3) * TSBO, (16 - (obs >> 3)) * TSBO);
                zero(H2_ST_LO + (obs >> 3) * TSBO, (16 - (obs >> 3)) * TSBO);
            }
            umma::fence_async_smem();
            rows_sync();
        }
        TC_STAMP(6);

        // the next round's slot travels during the passes (its record fields are requested before AdamW)
        int next_slot = 0;
        if (active && round + 1 < a.rounds) next_slot = L.slots[(size_t)(round + 1) * a.B + c.g * 128 + c.m];
        // ================= weight gradients: pass p = the 32 batch rows of warp p =================
        // The OWNER of pass p (warp p: the rows are its TMEM lanes and its registers hold dq, the ReLU mask and the action)
        // builds [E ; H1 ; 0]^T and [dZ2 | dZ1]^T and issues G2.  The S^T tiles need no TMEM (the loader re-fetched the raw
        // rows), so they are built by warp p + 6 mod 8 — on ANOTHER sub-partition than the owner (a TMEM lane quarter is
        // tied to the sub-partition of its warps, so splitting the owner's own work would not add issue slots) — which then
        // issues G1 = S^T dZ1 against the dZ1 rows of the owner's tile.  Arena hand-over between passes: one mbarrier per pass
        // (tcgen05.commit arrives): the G2 arena needs the commits of G2(p) and G1(p), the S^T arena that of G1(p).
        {
            const int colbase = (lane >> 2) * (TL / 4) + (lane & 3);
            float *dz_hi = reinterpret_cast<float *>(smem + H1_DZ_HI), *dz_lo = reinterpret_cast<float *>(smem + H1_DZ_LO);
            auto owner_half = [&](int p) {
                float *eh_hi = reinterpret_cast<float *>(smem + H1_EHT_HI), *eh_lo = reinterpret_cast<float *>(smem + H1_EHT_LO);
                float v[64];
                umma::tmem_ld32(c.tlane + c.s_col, v);            // h1
                umma::tmem_ld32(c.tlane + c.s_col + 32, v + 32);
                if (p > 0) umma::mbar_wait(bar + B_H1FREE + p - 1, round & 1);      // the previous pass's G2 and G1 are done
                eh_hi[tt_off(0, colbase)] = 1.f;
#pragma unroll
                for (int e = 0; e < 16; e++) eh_hi[tt_off(1 + e, colbase)] = (e == act_now) ? 1.f : 0.f;
#pragma unroll
                for (int j = 0; j < 64; j++) {
                    eh_hi[tt_off(32 + j, colbase)] = v[j];
                    eh_lo[tt_off(32 + j, colbase)] = tf32_lo(v[j]);
                }
                umma::tmem_ld32(c.tlane + c.s_col + 64, v);       // dZ1
                umma::tmem_ld32(c.tlane + c.s_col + 96, v + 32);
#pragma unroll
                for (int j = 0; j < 64; j++) {
                    dz_hi[tt_off(64 + j, colbase)] = v[j];
                    dz_lo[tt_off(64 + j, colbase)] = tf32_lo(v[j]);
                }
                umma::fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar + B_DZREADY + p);  // the S^T builder may multiply against dZ1^T
#pragma unroll
                for (int j = 0; j < 64; j++) {
                    const bool on = j < 32 ? ((mask2a >> j) & 1u) : ((mask2b >> (j - 32)) & 1u);
                    const float x = on ? dq * mi.o.w3[j] : 0.f;   // dZ2 again (cheaper than parking it)
                    dz_hi[tt_off(j, colbase)] = x;
                    dz_lo[tt_off(j, colbase)] = tf32_lo(x);
                }
                umma::fence_async_smem();
                umma::fence_before_thread_sync();
                __syncwarp();
                if (lane == 0) {
                    umma::fence_after_thread_sync();
                    const umma::Tile EH{umma::smem_u32(eh_hi), TL, TSBO}, EL{umma::smem_u32(eh_lo), TL, TSBO};
                    const umma::Tile DH{umma::smem_u32(dz_hi), TL, TSBO}, DL{umma::smem_u32(dz_lo), TL, TSBO};
                    const uint32_t i128 = umma::make_idesc_tf32(128, 128);
#pragma unroll
                    for (int ks = 0; ks < 4; ks++) {
                        umma::mma_tf32(c.tm + TM_G2, EL.desc(ks), DH.desc(ks), i128, p > 0 || ks > 0);
                        umma::mma_tf32(c.tm + TM_G2, EH.desc(ks), DL.desc(ks), i128, true);
                        umma::mma_tf32(c.tm + TM_G2, EH.desc(ks), DH.desc(ks), i128, true);
                    }
                    umma::mma_commit(bar + B_H1FREE + p);
                }
                __syncwarp();
            };
            auto st_half = [&](int p) {
                float *st_hi = reinterpret_cast<float *>(smem + H2_ST_HI), *st_lo = reinterpret_cast<float *>(smem + H2_ST_LO);
                const uint32_t n = n_pass + p;
                if (p > 0) umma::mbar_wait(bar + B_H2FREE + p - 1, round & 1);   // first: orders this pass behind the fill it waits for next
                umma::mbar_wait(bar + B_RFULL + (n & 1), (n >> 1) & 1);
                const float *raw = reinterpret_cast<const float *>(smem + ((n & 1) ? R_W2T : R_W2)) + lane * RAWP;
#pragma unroll 4
                for (int k4 = 0; k4 < (obs >> 2); k4++) {
                    const float4 x = *reinterpret_cast<const float4 *>(raw + 4 * k4);
                    const int o = tt_off(4 * k4, colbase);      // rows 4 k4 .. 4 k4 + 3 share one 8-row group
                    st_hi[o] = x.x; st_hi[o + 4] = x.y; st_hi[o + 8] = x.z; st_hi[o + 12] = x.w;
                    st_lo[o] = tf32_lo(x.x); st_lo[o + 4] = tf32_lo(x.y); st_lo[o + 8] = tf32_lo(x.z); st_lo[o + 12] = tf32_lo(x.w);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(bar + B_RFREE + (n & 1));
                umma::fence_async_smem();
                umma::fence_before_thread_sync();
                __syncwarp();
                if (lane == 0) {
                    umma::mbar_wait(bar + B_DZREADY + p, round & 1);             // the owner's dZ1^T rows are in place
                    umma::fence_after_thread_sync();
                    const umma::Tile SH{umma::smem_u32(st_hi), TL, TSBO}, SL{umma::smem_u32(st_lo), TL, TSBO};
                    const umma::Tile D1H = umma::Tile{umma::smem_u32(dz_hi), TL, TSBO}.rows_from(64);
                    const umma::Tile D1L = umma::Tile{umma::smem_u32(dz_lo), TL, TSBO}.rows_from(64);
                    const uint32_t i64 = umma::make_idesc_tf32(128, 64);
#pragma unroll
                    for (int ks = 0; ks < 4; ks++) {
                        umma::mma_tf32(c.tm + TM_G1, SL.desc(ks), D1H.desc(ks), i64, p > 0 || ks > 0);
                        umma::mma_tf32(c.tm + TM_G1, SH.desc(ks), D1L.desc(ks), i64, true);
                        umma::mma_tf32(c.tm + TM_G1, SH.desc(ks), D1H.desc(ks), i64, true);
                    }
                    umma::mma_commit(bar + B_H2FREE + p);
                    umma::mma_commit(bar + B_H1FREE + p);
                }
                __syncwarp();
            };
            const int p_own = warp, p_st = (warp + 2) & 7;
            if (p_st < p_own) { if (p_st < np) st_half(p_st); if (p_own < np && active) owner_half(p_own); }
            else              { if (p_own < np && active) owner_half(p_own); if (p_st < np) st_half(p_st); }
        }
        umma::mbar_wait(bar + B_H1FREE + np - 1, round & 1);
        umma::mbar_wait(bar + B_H2FREE + np - 1, round & 1);
        umma::fence_after_thread_sync();
        n_pass += np;
        TC_STAMP(7);

        // the next round's row and its first two chunks travel while AdamW runs (the staging buffers are free again)
        if (active && round + 1 < a.rounds) {
            fetch_row(next_slot);
            first_chunks(round + 1);
        }
        TC_STAMP(9);

        // ================= AdamW =================
        // (1) TMEM accumulators -> the gradient in the flat (torch) parameter order in shared memory (the G1 arena is free):
        //     lane = input index, so a warp's 32 values of one output row are 128 contiguous bytes (conflict-free).
        // (2) all 256 row threads sweep the flat vectors with 16-byte accesses, two float4 groups in flight per thread.
        {
            const float2 sc = L.scal[round];
            AdamScalarsTc hs;
            hs.decay = a.decay; hs.omb1 = a.omb1; hs.beta2 = a.beta2; hs.omb2 = a.omb2; hs.eps = a.eps;
            hs.step_size = sc.x; hs.bc2_sqrt = sc.y; hs.inv_bc2_sqrt = 1.0f / sc.y;
            float *gsm = reinterpret_cast<float *>(smem + R_W1);
            const int gs = c.g, j0 = 32 * gs;
            {   // G1[lane k][col j] = dW1[j][k]
                float g1[32];
                umma::tmem_ld32(c.tlane + TM_G1 + j0, g1);
                const int k = 32 * c.q + lane;
                if (k < obs)
#pragma unroll
                    for (int jj = 0; jj < 32; jj++) gsm[d.oW1 + (j0 + jj) * d.D + k] = g1[jj];
            }
            if (c.q == 1 || c.q == 2) {   // G2 lanes 32 + k2: dW2[j][k2] in the dZ2 columns
                float g2[32];
                umma::tmem_ld32(c.tlane + TM_G2 + j0, g2);
                const int k2 = 32 * (c.q - 1) + lane;
#pragma unroll
                for (int jj = 0; jj < 32; jj++) gsm[d.oW2 + (j0 + jj) * HID + k2] = g2[jj];
            } else if (c.q == 0) {        // G2 lane 0: db2 (dZ2 columns), db1 (dZ1 columns); lanes 1..A: dW1[:, obs + a]
                float gz2[32], gz1[32];
                umma::tmem_ld32(c.tlane + TM_G2 + j0, gz2);
                umma::tmem_ld32(c.tlane + TM_G2 + 64 + j0, gz1);
                if (lane == 0) {
#pragma unroll
                    for (int jj = 0; jj < 32; jj++) { gsm[d.ob1 + j0 + jj] = gz1[jj]; gsm[d.ob2 + j0 + jj] = gz2[jj]; }
                } else if (lane <= d.A) {
#pragma unroll
                    for (int jj = 0; jj < 32; jj++) gsm[d.oW1 + (j0 + jj) * d.D + obs + lane - 1] = gz1[jj];
                }
            } else {                      // W3 and b3 from the shuffle reductions of phase O, fixed summation order
                const int j = j0 + lane;
                float gsum = 0.f;
#pragma unroll
                for (int w8 = 0; w8 < 8; w8++) gsum += mi.redw[w8][j];
                gsm[d.oW3 + j] = gsum;
                if (gs == 1 && lane == 0) {
                    float ds = 0.f, ms = 0.f;
#pragma unroll
                    for (int w8 = 0; w8 < 8; w8++) { ds += mi.reddb3[w8]; ms += mi.redmae[w8]; }
                    gsm[d.ob3] = ds;
                    L.out_mae[round] = ms / (float)a.B;   // reported "loss": mean |q - y|
                }
            }
            TC_STAMP(10);
            umma::fence_before_thread_sync();
            rows_sync();
            TC_STAMP(11);
            const int P = d.P;
            if ((reinterpret_cast<uintptr_t>(L.w) | reinterpret_cast<uintptr_t>(L.m) | reinterpret_cast<uintptr_t>(L.v) |
                 reinterpret_cast<uintptr_t>(L.vmax)) & 15) {
                for (int i = tid; i < P; i += NROW) {      // unaligned vectors: scalar sweep
                    float mm = __ldcg(L.m + i), vv = __ldcg(L.v + i), xx = __ldcg(L.vmax + i);
                    L.w[i] = adam_math(__ldcg(L.w + i), mm, vv, xx, gsm[i], hs);
                    L.m[i] = mm; L.v[i] = vv; L.vmax[i] = xx;
                }
            } else {
                const int P4 = P >> 2;
                for (int q0 = tid; q0 < P4; q0 += 2 * NROW) {
                    float4 w4[2], m4[2], v4[2], x4[2];
#pragma unroll
                    for (int u = 0; u < 2; u++) {
                        const int qq = q0 + u * NROW;
                        if (qq < P4) {
                            w4[u] = __ldcg(reinterpret_cast<const float4 *>(L.w) + qq); m4[u] = __ldcg(reinterpret_cast<const float4 *>(L.m) + qq);
                            v4[u] = __ldcg(reinterpret_cast<const float4 *>(L.v) + qq); x4[u] = __ldcg(reinterpret_cast<const float4 *>(L.vmax) + qq);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 2; u++) {
                        const int qq = q0 + u * NROW;
                        if (qq < P4) {
                            const float4 g4 = *reinterpret_cast<const float4 *>(gsm + 4 * qq);
                            w4[u].x = adam_math(w4[u].x, m4[u].x, v4[u].x, x4[u].x, g4.x, hs);
                            w4[u].y = adam_math(w4[u].y, m4[u].y, v4[u].y, x4[u].y, g4.y, hs);
                            w4[u].z = adam_math(w4[u].z, m4[u].z, v4[u].z, x4[u].z, g4.z, hs);
                            w4[u].w = adam_math(w4[u].w, m4[u].w, v4[u].w, x4[u].w, g4.w, hs);
                            reinterpret_cast<float4 *>(L.w)[qq] = w4[u]; reinterpret_cast<float4 *>(L.m)[qq] = m4[u];
                            reinterpret_cast<float4 *>(L.v)[qq] = v4[u]; reinterpret_cast<float4 *>(L.vmax)[qq] = x4[u];
                        }
                    }
                }
                for (int i = 4 * P4 + tid; i < P; i += NROW) {
                    float mm = __ldcg(L.m + i), vv = __ldcg(L.v + i), xx = __ldcg(L.vmax + i);
                    L.w[i] = adam_math(__ldcg(L.w + i), mm, vv, xx, gsm[i], hs);
                    L.m[i] = mm; L.v[i] = vv; L.vmax[i] = xx;
                }
            }
            umma::fence_before_thread_sync();
            rows_sync();
            TC_STAMP(12);
            load_smalls(L.w, d, mi.o, tid);
        }
        TC_STAMP(8);
    }
    umma::fence_before_thread_sync();
    cta_sync();
    if (warp == 0) umma::tmem_dealloc(mi.tmem_base, 512);
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
    PRL_REQUIRE((size_t)count * sizeof(TcLearner) <= 32 * 1024 && (size_t)count * sizeof(MultiSampler) <= 32 * 1024,
                "too many learners in one group");
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
        PRL_REQUIRE((b->lay.record_words & 3) == 0 && (b->lay.off_state & 3) == 0 && (b->lay.off_next_state & 3) == 0 &&
                        ((uintptr_t)b->records & 15) == 0,
                    "record rows must be 16-byte aligned for the TMA row copies");
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
        L.w = q->w; L.wt = q->wt; L.m = q->m; L.v = q->v; L.vmax = q->vmax;
        L.tiles = q->tc_tiles;
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
    auto chunk_begin = [&](int cix) { return (int)((long long)rounds * cix / nchunks); };
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
