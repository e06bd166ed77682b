// gemm_tc.cu — the actor-critic learners' contraction (gemm.cuh: C[i][j] = sum_c A(i, c) * B(j, c)) on the 5th-generation
// tensor cores: tcgen05.mma kind::tf32 with the 3xTF32 operand split of umma.cuh (hi*hi + hi*lo + lo*hi accumulated in
// fp32 in tensor memory), so the products keep fp32 parity with the reference's torch matmuls (1e-4 gate, ~1e-6 measured).
//
// One CTA computes a 128 x TN tile of C (TN = 64 or 32).  The contraction axis is walked in chunks of 32:
//   * all 8 warps are loaders: they read the chunk's A (128 x 32) and B (TN x 32) elements from global memory exactly as
//     gemm.cuh's `Mat` describes them (two concatenated sources, a ones column, stacked networks; either the row or the
//     feature axis of the Mat is the contraction axis), TWO chunks ahead of the one being multiplied (three register
//     sets), split them into hi / lo and store them as 16-byte chunks into K-major no-swizzle UMMA tiles with a 144-byte
//     chunk pitch — with that pitch both global layouts give conflict-free STS.128 and fully coalesced global loads;
//   * three shared-memory stages; one thread issues the 12 MMAs of a chunk (4 K-steps x 3 split terms) and commits them
//     to the stage's mbarrier, which is what the loaders wait on before they overwrite the stage three chunks later;
//   * epilogue: tcgen05.ld of the accumulator (warp w: lanes 32(w%4).., columns 32(w/4)..), a per-warp transpose through
//     shared memory so that bias / mask / accumulate reads and the C stores are coalesced along j.
// What bounds it (profiles/r2_gemm_tc.md): not the tensor pipe (12 % active) and not shared memory (the TS form below, which
// keeps the 128-row operand out of shared memory altogether, runs at the same speed) but the operand loads: a CTA moves
// 24 KB from L2 per 32-deep chunk of its 128 x 64 tile, ~1500 clk per chunk alone and ~2800 clk with every SM busy, i.e.
// 21 FLOP per byte at the ~2.5 TB/s the chip delivers to this access pattern = the measured 52 TFLOP/s on 65536 x 256 x
// 256.  That is 2.1x the SIMT tiles on large products and a tie (or a loss) on the learners' own batch-256 / 512 products,
// which is how gemm_tc_launch chooses.
#include "gemm.cuh"
#include "umma.cuh"

namespace prl {

int g_contraction_engine = 1;   // 1: tcgen05 tiles, 0: SIMT tiles (prl_set_contraction_engine)

namespace {

constexpr int TM = 128, GKT = 32, LBO = 144, SBO = 8 * LBO, NSTAGE = 3, NSET = 3, NLOAD = 256, NTHR = NLOAD + 32;

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(umma::smem_u32(bar)) : "memory");
}

// which (row of the tile, 16-byte K chunk) a thread's slot u covers.  XO: the Mat's rows are the tile's rows (features are
// contracted, contiguous along k): 8 lanes read one row's 128 contiguous bytes.  !XO: the Mat's rows are contracted
// (contiguous along the tile's row index): 32 lanes read 32 consecutive tile rows of one Mat row.
template <int ROWS, bool XO>
__device__ __forceinline__ void slot_coords(int u, int w, int lane, int &r, int &kc) {
    if (XO) {
        r = w * (ROWS / 8) + u * 4 + (lane >> 3);
        kc = lane & 7;
    } else {
        constexpr int NRB = ROWS / 32;
        r = (w % NRB) * 32 + lane;
        kc = (w / NRB) * NRB + u;
    }
}

// One operand (ROWS x 32 per chunk) as seen by one loader thread.  Everything that does not change from chunk to chunk
// is computed once.  Loads never need a predicate: indices are clamped into the operand (a tile row past the operand's
// extent repeats its last row — its products land in rows / columns of C that the epilogue never stores — and an index past
// the end of the contraction axis repeats the last one and is zeroed when the tile is stored).  The loop body is kept
// small on purpose: the first version inlined four paths per operand and ran out of the instruction cache at ~10 clocks
// per instruction.
template <int ROWS, bool XO>
struct Operand {
    static constexpr int S = ROWS / 32;
    const Mat &m;
    int Kc, kc4;                // contraction length; XO: this thread's 16-byte chunk (floats) inside a 32-deep chunk
    const float *zb1, *zb2;     // part bases with the network offset applied (zb2 also with -split)
    int off1[S], off2[S];       // XO: float offset of the slot's row in part 1 / part 2
    int kcu[S];                 // !XO: first contraction index of slot u inside a chunk
    int soff[S];                // byte offset of the slot's 16-byte chunk inside the UMMA tile
    const float *fbase;         // !XO: this thread's feature column
    int fld;                    // !XO: its row pitch
    bool isone, vec1, vec2;

    __device__ __forceinline__ Operand(const Mat &m_, int z, int out0, int out_lim, int Kc_, int w, int lane) : m(m_), Kc(Kc_) {
        zb1 = m.p1 + z * m.net_stride1;
        zb2 = m.p2 + z * m.net_stride2 - m.split;
        fbase = zb1; fld = 0; isone = false; kc4 = 0; vec1 = vec2 = false;
#pragma unroll
        for (int u = 0; u < S; u++) {
            int r, kc;
            slot_coords<ROWS, XO>(u, w, lane, r, kc);
            soff[u] = (r >> 3) * SBO + kc * LBO + (r & 7) * 16;
            kcu[u] = kc * 4;
            const int row = min(out0 + r, out_lim - 1);
            off1[u] = XO ? row * m.ld1 : 0;
            off2[u] = XO ? row * m.ld2 : 0;
            if (XO) kc4 = kc * 4;
        }
        if (XO) {
            vec1 = ((reinterpret_cast<uintptr_t>(zb1) & 15) | (m.ld1 & 3)) == 0;
            vec2 = m.p2 != nullptr && ((reinterpret_cast<uintptr_t>(zb2) & 15) | (m.ld2 & 3)) == 0;
        } else {
            int r, kc;
            slot_coords<ROWS, XO>(0, w, lane, r, kc);
            const int out = out0 + r;
            isone = out == m.ones_at && out < out_lim;
            int f = min(out, out_lim - 1);
            if (f == m.ones_at) f = 0;                      // any valid address: the value is replaced by 1.0 / never used
            fbase = f < m.split ? zb1 + f : zb2 + f;
            fld = f < m.split ? m.ld1 : m.ld2;
        }
    }
    // XO: chunk c lies inside one source, away from the end of the contraction axis and the ones column
    __device__ __forceinline__ bool interior(int f0) const {
        return f0 + GKT <= Kc && (f0 + GKT <= m.split || f0 >= m.split) && !(m.ones_at >= f0 && m.ones_at < f0 + GKT);
    }
    __device__ __forceinline__ void fetch(int c, float (&reg)[S][4]) const {
        const int f0 = c * GKT;
        if (XO) {
            const bool second = f0 >= m.split;
            if (interior(f0) && (second ? vec2 : vec1)) {
                const float *base = (second ? zb2 : zb1) + f0 + kc4;
#pragma unroll
                for (int u = 0; u < S; u++) {
                    const float4 q = __ldg(reinterpret_cast<const float4 *>(base + (second ? off2[u] : off1[u])));
                    reg[u][0] = q.x; reg[u][1] = q.y; reg[u][2] = q.z; reg[u][3] = q.w;
                }
            } else {
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    int f = min(f0 + kc4 + e, Kc - 1);
                    if (f == m.ones_at) f = 0;
                    const bool sec = f >= m.split;
                    const float *base = (sec ? zb2 : zb1) + f;
#pragma unroll
                    for (int u = 0; u < S; u++) reg[u][e] = __ldg(base + (sec ? off2[u] : off1[u]));
                }
            }
        } else if (isone) {                      // the ones column (bias gradient): no load, the stored value is 1
#pragma unroll
            for (int u = 0; u < S; u++)
#pragma unroll
                for (int e = 0; e < 4; e++) reg[u][e] = 1.f;
        } else if (f0 + GKT <= Kc) {
#pragma unroll
            for (int u = 0; u < S; u++) {
                const float *q = fbase + (size_t)(f0 + kcu[u]) * fld;
#pragma unroll
                for (int e = 0; e < 4; e++) reg[u][e] = __ldg(q + (size_t)e * fld);
            }
        } else {
#pragma unroll
            for (int u = 0; u < S; u++)
#pragma unroll
                for (int e = 0; e < 4; e++) reg[u][e] = __ldg(fbase + (size_t)min(f0 + kcu[u] + e, Kc - 1) * fld);
        }
    }
    __device__ __forceinline__ bool edge(int c) const { return XO ? !interior(c * GKT) : c * GKT + GKT > Kc; }
    // the value of element (slot u, e) of an EDGE chunk as the product sees it (ones column, zero past the contraction axis);
    // in every other chunk it is the loaded value itself
    __device__ __forceinline__ float edge_elem(int c, int u, int e, float loaded) const {
        const int con = c * GKT + (XO ? kc4 : kcu[u]) + e;
        float v = loaded;
        if (XO && con == m.ones_at) v = 1.f;
        if (con >= Kc) v = 0.f;
        return v;
    }
    // split into hi (the value itself: the tensor core truncates) / lo and store as 16-byte chunks of the UMMA tiles
    __device__ __forceinline__ void store(int c, const float (&reg)[S][4], unsigned char *hi, unsigned char *lo) const {
        if (edge(c)) {                               // uniform branch, taken for at most two chunks of a product
#pragma unroll
            for (int u = 0; u < S; u++) {
                float v[4], l[4];
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    v[e] = edge_elem(c, u, e, reg[u][e]);
                    l[e] = v[e] - __uint_as_float(__float_as_uint(v[e]) & 0xffffe000u);
                }
                *reinterpret_cast<float4 *>(hi + soff[u]) = make_float4(v[0], v[1], v[2], v[3]);
                *reinterpret_cast<float4 *>(lo + soff[u]) = make_float4(l[0], l[1], l[2], l[3]);
            }
            return;
        }
#pragma unroll
        for (int u = 0; u < S; u++) {
            float l[4];
#pragma unroll
            for (int e = 0; e < 4; e++) l[e] = reg[u][e] - __uint_as_float(__float_as_uint(reg[u][e]) & 0xffffe000u);
            *reinterpret_cast<float4 *>(hi + soff[u]) = make_float4(reg[u][0], reg[u][1], reg[u][2], reg[u][3]);
            *reinterpret_cast<float4 *>(lo + soff[u]) = make_float4(l[0], l[1], l[2], l[3]);
        }
    }
    // ROWS = 128, !XO only: this thread's 16 values of chunk c (tile row = 32 (w % 4) + lane, k = 16 (w / 4) .. + 16) go to
    // tensor memory as the A operand of the TS-form MMA: hi at column a_col + 16 (w / 4), lo 32 columns further
    __device__ __forceinline__ void store_tmem(int c, const float (&reg)[S][4], uint32_t taddr_hi) const {
        static_assert(!XO || ROWS != 128, "TMEM operands are loaded with the lanes along the tile rows");
        const bool is_edge = edge(c);
        float v[16], l[16];
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int e = 0; e < 4; e++) v[u * 4 + e] = reg[u][e];
        if (is_edge) {
#pragma unroll
            for (int u = 0; u < 4; u++)
#pragma unroll
                for (int e = 0; e < 4; e++) v[u * 4 + e] = edge_elem(c, u, e, v[u * 4 + e]);
        }
#pragma unroll
        for (int i = 0; i < 16; i++) l[i] = v[i] - __uint_as_float(__float_as_uint(v[i]) & 0xffffe000u);
        umma::tmem_st16(taddr_hi, v);
        umma::tmem_st16(taddr_hi + 32, l);
    }
};

// warps 0-7: loaders + epilogue; warp 8: TMEM owner, one lane issues the MMAs.  288 threads put three warps on one
// sub-partition, which caps the kernel at 168 registers per thread.
template <int TN, bool AO, bool BO>
__global__ void __maxnreg__(168) k_gemm_tc(const GemmArgs g) {
    extern __shared__ __align__(128) unsigned char dsm[];
    __shared__ __align__(8) uint64_t full[NSTAGE], empty[NSTAGE], done;
    __shared__ uint32_t tmem_slot;
    constexpr int A_B = (TM / 8) * SBO, B_B = (TN / 8) * SBO, STAGE_B = 2 * A_B + 2 * B_B;
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31, z = blockIdx.z;
    const int i0 = blockIdx.x * TM, j0 = blockIdx.y * TN;
    const int nch = (g.Kc + GKT - 1) / GKT;
    const bool prof = g.stamps != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && z == 0 && lane == 0 && (w == 0 || w == 8);

    if (w == 8) {
        umma::tmem_alloc(&tmem_slot, TN);
        if (lane == 0) {
            for (int b = 0; b < NSTAGE; b++) { umma::mbar_init(&full[b], NLOAD / 32); umma::mbar_init(&empty[b], 1); }
            umma::mbar_init(&done, 1);
        }
        umma::fence_before_thread_sync();
        __syncthreads();
        umma::fence_after_thread_sync();
        if (umma::elect_one()) {   // one lane of the converged issuer warp: back-to-back UTCHMMA, descriptor math on the uniform datapath
            const uint32_t tmem = tmem_slot, idesc = umma::make_idesc_tf32(TM, TN);
            const uint64_t dproto = umma::make_desc2(0, LBO, SBO);
            const uint32_t base = umma::smem_u32(dsm);
            for (int c = 0; c < nch; c++) {
                const int s = c % NSTAGE;
                umma::mbar_wait(&full[s], (c / NSTAGE) & 1);
                if (prof && c < 32) g.stamps[c * 8 + 5] = clock64();
                umma::fence_after_thread_sync();
                const uint32_t a = (base + s * STAGE_B) >> 4;
                const uint64_t ah = dproto + a, al = ah + (A_B >> 4), bh = al + (A_B >> 4), bl = bh + (B_B >> 4);
#pragma unroll
                for (int ks = 0; ks < GKT / 8; ks++) {
                    const uint64_t o = (uint64_t)(ks * ((2 * LBO) >> 4));
                    umma::mma_tf32(tmem, al + o, bh + o, idesc, c > 0 || ks > 0);   // small terms first
                    umma::mma_tf32(tmem, ah + o, bl + o, idesc, true);
                    umma::mma_tf32(tmem, ah + o, bh + o, idesc, true);
                }
                umma::mma_commit(&empty[s]);      // arrives when the MMAs that read stage s have completed
                if (prof && c < 32) g.stamps[c * 8 + 6] = clock64();
            }
            umma::mma_commit(&done);
        }
        __syncwarp();
    } else {
        const Operand<TM, AO> opA(g.A, z, i0, g.Mo, g.Kc, w, lane);
        const Operand<TN, BO> opB(g.B, z, j0, g.No, g.Kc, w, lane);
        float ra[NSET][TM / 32][4], rb[NSET][TN / 32][4];
#pragma unroll
        for (int p = 0; p < NSET - 1; p++)
            if (p < nch) { opA.fetch(p, ra[p]); opB.fetch(p, rb[p]); }
        if (prof) g.stamps[32 * 8 + 2] = clock64();
        __syncthreads();   // barriers initialised, TMEM allocated
        if (prof) g.stamps[32 * 8 + 3] = clock64();
        static_assert(NSET == NSTAGE, "stage index == register set index");
        for (int c0 = 0; c0 < nch; c0 += NSET) {
#pragma unroll
            for (int u = 0; u < NSET; u++) {
                const int c = c0 + u;
                if (c < nch) {
                    if (prof && c < 32) g.stamps[c * 8 + 0] = clock64();
                    if (c + NSET - 1 < nch) {       // two chunks ahead of the one being stored
                        opA.fetch(c + NSET - 1, ra[(u + NSET - 1) % NSET]);
                        opB.fetch(c + NSET - 1, rb[(u + NSET - 1) % NSET]);
                    }
                    if (prof && c < 32) g.stamps[c * 8 + 1] = clock64();
                    if (c >= NSTAGE) umma::mbar_wait(&empty[u], ((c / NSTAGE) - 1) & 1);
                    if (prof && c < 32) g.stamps[c * 8 + 2] = clock64();
                    unsigned char *st = dsm + u * STAGE_B;
                    opA.store(c, ra[u], st, st + A_B);
                    opB.store(c, rb[u], st + 2 * A_B, st + 2 * A_B + B_B);
                    if (prof && c < 32) g.stamps[c * 8 + 4] = clock64();
                    umma::fence_async_smem();       // generic-proxy writes -> visible to the tensor core's async proxy
                    __syncwarp();
                    if (prof && c < 32) g.stamps[c * 8 + 3] = clock64();
                    if (lane == 0) mbar_arrive(&full[u]);
                }
            }
        }
        umma::mbar_wait(&done, 0);
        if (prof) g.stamps[32 * 8 + 0] = clock64();
        umma::fence_after_thread_sync();

        const uint32_t tmem = tmem_slot;
        const int q = w & 3, h = w >> 2;
        if (h * 32 < TN) {
            float v[32];
            umma::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * 32), v);
            float *tb = reinterpret_cast<float *>(dsm) + w * (32 * 33);   // the stages are free: every MMA has completed
#pragma unroll
            for (int t = 0; t < 32; t++) tb[lane * 33 + t] = v[t];
            __syncwarp();
            const int j = j0 + h * 32 + lane;
            if (j < g.No) {
                const float bj = g.bias ? __ldg(g.bias + z * g.bias_net_stride + j) : 0.f;
                const bool tail = g.C_tail && j == g.tail_col;
#pragma unroll 4
                for (int rr = 0; rr < 32; rr++) {
                    const int i = i0 + q * 32 + rr;
                    if (i >= g.Mo) break;
                    float x = tb[rr * 33 + lane];
                    if (tail) { g.C_tail[z * g.tail_net_stride + i] = x; continue; }
                    float *dst = g.C + z * g.c_net_stride + (size_t)i * g.ldc + j;
                    if (g.bias) x += bj;
                    if (g.accumulate) x += *dst;
                    if (g.relu) x = fmaxf(x, 0.f);
                    if (g.mask && !(__ldg(g.mask + z * g.mask_net_stride + (size_t)i * g.ldm + j) > 0.f)) x = 0.f;
                    *dst = x;
                }
            }
        }
        umma::fence_before_thread_sync();
        if (prof) g.stamps[32 * 8 + 1] = clock64();
    }
    __syncthreads();
    if (w == 8) {
        umma::fence_after_thread_sync();
        umma::tmem_dealloc(tmem_slot, TN);
    }
}

// ---- TS form: the A operand lives in tensor memory -----------------------------------------------------------------
// The SS kernel above moves 120 KB through shared memory per 32-deep chunk of a 128 x 64 tile (48 KB of hi / lo stores and
// 72 KB of operand reads by the 12 MMAs).  This form was written to test whether that is what bounds it — it is not: the two
// run within 10 % of each other.  Here the 128-row operand never touches shared memory: it must be the operand whose Mat
// has the tile rows along the contiguous axis ("rows contracted"), so that lane l of warp w loads row 32 (w % 4) + l for
// 16 consecutive contraction indices with coalesced loads and writes hi / lo straight into its own TMEM lane
// (tcgen05.st).  Shared memory only carries the TN-row operand: 16 KB of stores + 24 KB of MMA reads per chunk.
//   nn.Linear forward     D[n][m] = sum_k Wt[k][n] x[m][k]     A = a transposed copy of W kept by the learner, SWAP
//   backward-data         D[k][m] = sum_n W[n][k] dy[m][n]      A = W as stored, SWAP
//   backward-weight       D[n][k] = sum_m dy[m][n] x[m][k]      A = dy as stored
// SWAP: the accumulator is C transposed (lanes = C's column index): the epilogue stores straight from registers, coalesced.
constexpr int TS_ACC_COLS = 64, TS_STAGE_COLS = 64, TS_TMEM_COLS = 256;

template <int TN, bool BO, bool SWAP>
__global__ void __maxnreg__(168) k_gemm_ts(const GemmArgs g) {
    extern __shared__ __align__(128) unsigned char dsm[];
    __shared__ __align__(8) uint64_t full[NSTAGE], empty[NSTAGE], done;
    __shared__ uint32_t tmem_slot;
    constexpr int B_B = (TN / 8) * SBO, STAGE_B = 2 * B_B;
    static_assert(TN <= TS_ACC_COLS && TS_ACC_COLS + NSTAGE * TS_STAGE_COLS <= TS_TMEM_COLS, "TMEM column budget");
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31, z = blockIdx.z;
    const int a0 = blockIdx.x * TM, b0 = blockIdx.y * TN;
    const int nch = (g.Kc + GKT - 1) / GKT;

    if (w == 8) {
        umma::tmem_alloc(&tmem_slot, TS_TMEM_COLS);
        if (lane == 0) {
            for (int b = 0; b < NSTAGE; b++) { umma::mbar_init(&full[b], NLOAD / 32); umma::mbar_init(&empty[b], 1); }
            umma::mbar_init(&done, 1);
        }
        umma::fence_before_thread_sync();
        __syncthreads();
        umma::fence_after_thread_sync();
        if (umma::elect_one()) {   // one lane of the converged issuer warp: back-to-back UTCHMMA, descriptor math on the uniform datapath
            const uint32_t tmem = tmem_slot, idesc = umma::make_idesc_tf32(TM, TN);
            const uint64_t dproto = umma::make_desc2(0, LBO, SBO);
            const uint32_t base = umma::smem_u32(dsm);
            for (int c = 0; c < nch; c++) {
                const int s = c % NSTAGE;
                umma::mbar_wait(&full[s], (c / NSTAGE) & 1);
                umma::fence_after_thread_sync();
                const uint64_t bh = dproto + ((base + s * STAGE_B) >> 4), bl = bh + (B_B >> 4);
                const uint32_t ah = tmem + TS_ACC_COLS + s * TS_STAGE_COLS, al = ah + 32;
#pragma unroll
                for (int ks = 0; ks < GKT / 8; ks++) {
                    const uint64_t o = (uint64_t)(ks * ((2 * LBO) >> 4));
                    umma::mma_tf32_ts(tmem, al + ks * 8, bh + o, idesc, c > 0 || ks > 0);   // small terms first
                    umma::mma_tf32_ts(tmem, ah + ks * 8, bl + o, idesc, true);
                    umma::mma_tf32_ts(tmem, ah + ks * 8, bh + o, idesc, true);
                }
                umma::mma_commit(&empty[s]);
            }
            umma::mma_commit(&done);
        }
        __syncwarp();
    } else {
        const Operand<TM, false> opA(g.A, z, a0, g.Mo, g.Kc, w, lane);
        const Operand<TN, BO> opB(g.B, z, b0, g.No, g.Kc, w, lane);
        float ra[NSET][TM / 32][4], rb[NSET][TN / 32][4];
#pragma unroll
        for (int p = 0; p < NSET - 1; p++)
            if (p < nch) { opA.fetch(p, ra[p]); opB.fetch(p, rb[p]); }
        __syncthreads();   // barriers initialised, TMEM allocated
        const uint32_t tmem = tmem_slot;
        const uint32_t my_a = tmem + ((uint32_t)((w & 3) * 32) << 16) + TS_ACC_COLS + (w >> 2) * 16;   // this warp's lanes, its half of the chunk
        for (int c0 = 0; c0 < nch; c0 += NSET) {
#pragma unroll
            for (int u = 0; u < NSET; u++) {
                const int c = c0 + u;
                if (c < nch) {
                    if (c + NSET - 1 < nch) {
                        opA.fetch(c + NSET - 1, ra[(u + NSET - 1) % NSET]);
                        opB.fetch(c + NSET - 1, rb[(u + NSET - 1) % NSET]);
                    }
                    if (c >= NSTAGE) {
                        umma::mbar_wait(&empty[u], ((c / NSTAGE) - 1) & 1);
                        umma::fence_after_thread_sync();
                    }
                    unsigned char *st = dsm + u * STAGE_B;
                    opA.store_tmem(c, ra[u], my_a + u * TS_STAGE_COLS);
                    opB.store(c, rb[u], st, st + B_B);
                    umma::tmem_st_wait();
                    umma::fence_async_smem();
                    umma::fence_before_thread_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&full[u]);
                }
            }
        }
        umma::mbar_wait(&done, 0);
        umma::fence_after_thread_sync();

        const int q = w & 3, h = w >> 2;
        if (h * 32 < TN) {
            float v[32];
            umma::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * 32), v);
            if (SWAP) {      // lane = C's column j, accumulator column = C's row i: every store instruction covers 32 consecutive j
                const int j = a0 + q * 32 + lane;
                if (j < g.Mo) {
#pragma unroll
                    for (int t = 0; t < 32; t++) {
                        const int i = b0 + h * 32 + t;
                        if (i < g.No) gemm_emit(g, z, i, j, v[t]);
                    }
                }
            } else {         // lane = C's row: transpose through shared memory (the stages are free) for coalesced stores
                float *tb = reinterpret_cast<float *>(dsm) + w * (32 * 33);
#pragma unroll
                for (int t = 0; t < 32; t++) tb[lane * 33 + t] = v[t];
                __syncwarp();
                const int j = b0 + h * 32 + lane;
                if (j < g.No) {
#pragma unroll 4
                    for (int rr = 0; rr < 32; rr++) {
                        const int i = a0 + q * 32 + rr;
                        if (i < g.Mo) gemm_emit(g, z, i, j, tb[rr * 33 + lane]);
                    }
                }
            }
        }
        umma::fence_before_thread_sync();
    }
    __syncthreads();
    if (w == 8) {
        umma::fence_after_thread_sync();
        umma::tmem_dealloc(tmem_slot, TS_TMEM_COLS);
    }
}

template <int TN>
constexpr int ts_smem_bytes() {
    constexpr int stages = NSTAGE * 2 * (TN / 8) * SBO, transpose = 8 * 32 * 33 * 4;
    return stages > transpose ? stages : transpose;
}
template <int TN, bool BO, bool SWAP>
void launch_ts(const GemmArgs &g, int nets, cudaStream_t st) {
    dim3 grid((g.Mo + TM - 1) / TM, (g.No + TN - 1) / TN, nets);
    k_gemm_ts<TN, BO, SWAP><<<grid, NTHR, ts_smem_bytes<TN>(), st>>>(g);
}

template <int TN>
constexpr int smem_bytes() { return NSTAGE * (2 * (TM / 8) * SBO + 2 * (TN / 8) * SBO); }

template <int TN, bool AO, bool BO>
cudaError_t prepare_one() {
    return cudaFuncSetAttribute(k_gemm_tc<TN, AO, BO>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<TN>());
}

template <int TN, bool AO, bool BO>
void launch_one(const GemmArgs &g, int nets, cudaStream_t st) {
    dim3 grid((g.Mo + TM - 1) / TM, (g.No + TN - 1) / TN, nets);
    k_gemm_tc<TN, AO, BO><<<grid, NTHR, smem_bytes<TN>(), st>>>(g);
}

template <bool AO, bool BO>
void launch_tn(const GemmArgs &g, int nets, cudaStream_t st, int tn) {
    if (tn == 64) launch_one<64, AO, BO>(g, nets, st);
    else launch_one<32, AO, BO>(g, nets, st);
}

}  // namespace

// >48 KB of dynamic shared memory needs the attribute; set once per device from prl_init (never inside a stream capture)
cudaError_t gemm_tc_prepare() {
    cudaError_t e;
    if ((e = prepare_one<64, true, true>()) != cudaSuccess) return e;
    if ((e = prepare_one<32, true, true>()) != cudaSuccess) return e;
    if ((e = prepare_one<64, true, false>()) != cudaSuccess) return e;
    if ((e = prepare_one<32, true, false>()) != cudaSuccess) return e;
    if ((e = prepare_one<64, false, false>()) != cudaSuccess) return e;
    if ((e = prepare_one<32, false, false>()) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_gemm_ts<64, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ts_smem_bytes<64>())) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_gemm_ts<32, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ts_smem_bytes<32>())) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_gemm_ts<64, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ts_smem_bytes<64>())) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_gemm_ts<32, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ts_smem_bytes<32>())) != cudaSuccess) return e;
    return cudaSuccess;
}

// engine: -1 = the library default (g_contraction_engine); 0 = SIMT tiles; 1 = automatic: tcgen05 tiles from 4096 output
// rows on, where they are the faster of the two (see the header of this file), SIMT tiles below; 2 = tcgen05 always (callers
// whose results must not depend on how many rows a launch sees: PPO's rollout passes); 64 / 32 = tcgen05 with that tile width.
bool gemm_tc_launch(const GemmArgs &g, int nets, bool ao, bool bo, cudaStream_t st, int engine) {
    if (engine < 0) engine = g_contraction_engine;
    if (engine == 0 || g.Mo <= 0 || g.No <= 0 || g.Kc <= 0) return false;
    if (!ao && bo) return false;   // not a shape the learners use
    if (engine == 1 && g.Mo < 4096) return false;
    if (engine == 164 || engine == 132) engine = 2;    // the TS form declined the shape
    const int tn = engine == 64 || engine == 32 ? engine : (g.No > 32 ? 64 : 32);
    if (ao && bo) launch_tn<true, true>(g, nets, st, tn);
    else if (ao) launch_tn<true, false>(g, nets, st, tn);
    else launch_tn<false, false>(g, nets, st, tn);
    return true;
}

// TS form.  gs is already in the kernel's orientation: gs.A = the 128-row operand (its Mat rows are contracted), gs.Mo its
// extent, gs.B / gs.No the other operand; swap = the accumulator is C transposed.  Only on request (engine 164 / 132 = this
// form with tile width 64 / 32): measured on a B200 it is within +-10 % of the SS form on every shape of the learners
// (profiles/r2_gemm_tc.md) — neither form is bound by shared-memory bandwidth or by the tensor pipe, both wait for the
// global loads of their operands — so the automatic choice never needs it and no learner has to keep transposed weights.
bool gemm_ts_launch(const GemmArgs &gs, int nets, bool bo, bool swap, cudaStream_t st, int engine) {
    if (engine < 0) engine = g_contraction_engine;
    if ((engine != 164 && engine != 132) || gs.Mo <= 0 || gs.No <= 0 || gs.Kc <= 0) return false;
    if (bo != swap) return false;   // the two combinations the learners use
    const int tn = engine == 164 ? 64 : 32;
    if (swap) { if (tn == 64) launch_ts<64, true, true>(gs, nets, st); else launch_ts<32, true, true>(gs, nets, st); }
    else { if (tn == 64) launch_ts<64, false, false>(gs, nets, st); else launch_ts<32, false, false>(gs, nets, st); }
    return true;
}

// dst[z][c][r] = src[z][r][c]  (the transposed weight copies the TS-form forward reads)
__global__ void k_transpose(int rows, int cols, const float *__restrict__ src, float *__restrict__ dst) {
    __shared__ float t[32][33];
    const size_t zo = (size_t)blockIdx.z * rows * cols;
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int y = threadIdx.y; y < 32; y += blockDim.y)
        if (r0 + y < rows && c0 + threadIdx.x < cols) t[y][threadIdx.x] = src[zo + (size_t)(r0 + y) * cols + c0 + threadIdx.x];
    __syncthreads();
    for (int y = threadIdx.y; y < 32; y += blockDim.y)
        if (c0 + y < cols && r0 + threadIdx.x < rows) dst[zo + (size_t)(c0 + y) * rows + r0 + threadIdx.x] = t[threadIdx.x][y];
}
void transpose_weights(int rows, int cols, int nets, const float *src, float *dst, cudaStream_t st) {
    dim3 grid((cols + 31) / 32, (rows + 31) / 32, nets), block(32, 8);
    k_transpose<<<grid, block, 0, st>>>(rows, cols, src, dst);
}

}  // namespace prl

extern "C" int prl_set_contraction_engine(int engine) {
    PRL_REQUIRE(engine >= 0 && engine <= 2, "engine must be 0 (SIMT tiles), 1 (automatic) or 2 (tcgen05 tiles always)");
    prl::g_contraction_engine = engine;
    return PRL_OK;
}
extern "C" int prl_get_contraction_engine(void) { return prl::g_contraction_engine; }

static long long *g_test_stamps = nullptr;
/* developer profiling: device int64[33][8] receiving SM-clock stamps of CTA 0 of the next prl_test_contraction calls */
extern "C" int prl_test_contraction_stamps(long long *stamps_dev) { g_test_stamps = stamps_dev; return PRL_OK; }

// Test hook: one contraction of the learners' three kinds through GemmLauncher.
//   op 0  y[M x N]  = act(x W^T + b)          a = x [M x K] (or [M x split] with a2 = [M x (K - split)]), b = W [N x K]
//   op 1  dx[M x K] (+)= dy W  (masked)       a = dy [M x N], b = W [N x K], mask [M x K]
//   op 2  dW[N x K] = dy^T x, db = dy^T 1     a = dy [M x N], b = x [M x K] (or split with a2), c_tail = db [N]
// `nets` stacked problems are laid out contiguously in every operand.
extern "C" int prl_test_contraction(int op, int engine, int M, int N, int K, const float *a, const float *b, const float *a2, int split,
                                    const float *bias, const float *mask, int relu, int accumulate, float *c, float *c_tail, int nets,
                                    void *stream) {
    using namespace prl;
    PRL_REQUIRE(op >= 0 && op <= 2 && M > 0 && N > 0 && K > 0 && a && b && c && nets >= 1, "bad argument");
    GemmLauncher L;   // prl_init has prepared the kernels
    L.st = (cudaStream_t)stream;
    L.engine = engine;
    L.stamps = g_test_stamps;
    const long long MK = (long long)M * K, MN = (long long)M * N, NK = (long long)N * K;
    if (op == 0) {
        Mat X = a2 ? mat2(a, split, split, a2, K - split, (long long)M * split, (long long)M * (K - split)) : mat(a, K, MK);
        float *wt = nullptr;
        if (engine == 164 || engine == 132) {          // a learner using this form would keep the copy up to date itself
            PRL_CUDA(cudaMallocAsync((void **)&wt, sizeof(float) * NK * nets, L.st));
            transpose_weights(N, K, nets, b, wt, L.st);
        }
        L.fwd(X, M, b, K, NK, bias, N, N, K, relu != 0, c, N, MN, nets, wt, N, NK);
        if (wt) PRL_CUDA(cudaFreeAsync(wt, L.st));
    } else if (op == 1) {
        L.bwd_x(a, N, MN, M, N, b, K, NK, 0, K, c, K, MK, mask, K, MK, accumulate != 0, nets);
    } else {
        PRL_REQUIRE(c_tail, "op 2 needs c_tail");
        Mat X = a2 ? mat2(b, split, split, a2, K - split, (long long)M * split, (long long)M * (K - split)) : mat(b, K, MK);
        L.bwd_w(a, N, MN, M, N, X, K, c, K, NK, c_tail, N, nets);
    }
    PRL_CUDA(cudaGetLastError());
    return PRL_OK;
}
