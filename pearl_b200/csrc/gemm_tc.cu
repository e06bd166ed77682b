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
// The launch is latency-bound for the learners' shapes (batch 256-512, widths 64-256): what it buys over the SIMT
// tiles is ~0.35 us per 32-deep chunk instead of ~1.2 us, see profiles/r2_gemm_tc.md.
#include "gemm.cuh"
#include "umma.cuh"

namespace prl {

int g_contraction_engine = 1;   // 1: tcgen05 tiles, 0: SIMT tiles (prl_set_contraction_engine)

namespace {

constexpr int TM = 128, GKT = 32, LBO = 144, SBO = 8 * LBO, NSTAGE = 3, NSET = 3, NTHR = 256;

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

template <int ROWS, bool XO>
__device__ __forceinline__ void fetch_op(const Mat &m, int z, int out0, int out_lim, int Kc, int c, int w, int lane,
                                         float (&reg)[ROWS / 32][4]) {
#pragma unroll
    for (int u = 0; u < ROWS / 32; u++) {
        int r, kc;
        slot_coords<ROWS, XO>(u, w, lane, r, kc);
        const int out = out0 + r;
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const int con = c * GKT + kc * 4 + e;
            const int row = XO ? out : con, f = XO ? con : out;
            const bool ld = out < out_lim && con < Kc && f != m.ones_at;
            reg[u][e] = __ldg(ld ? m.addr(row, f, z) : m.p1);   // unconditional loads: all of a chunk's are in flight together
        }
    }
}

template <int ROWS, bool XO>
__device__ __forceinline__ void store_op(const Mat &m, int out0, int out_lim, int Kc, int c, int w, int lane,
                                         const float (&reg)[ROWS / 32][4], unsigned char *hi, unsigned char *lo) {
#pragma unroll
    for (int u = 0; u < ROWS / 32; u++) {
        int r, kc;
        slot_coords<ROWS, XO>(u, w, lane, r, kc);
        const int out = out0 + r;
        float v[4], l[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const int con = c * GKT + kc * 4 + e;
            const int f = XO ? con : out;
            const bool ok = out < out_lim && con < Kc;
            v[e] = ok ? (f == m.ones_at ? 1.f : reg[u][e]) : 0.f;
            l[e] = v[e] - __uint_as_float(__float_as_uint(v[e]) & 0xffffe000u);   // the tensor core truncates hi itself
        }
        const int off = (r >> 3) * SBO + kc * LBO + (r & 7) * 16;
        *reinterpret_cast<float4 *>(hi + off) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4 *>(lo + off) = make_float4(l[0], l[1], l[2], l[3]);
    }
}

template <int TN, bool AO, bool BO>
__global__ void __launch_bounds__(NTHR, 1) k_gemm_tc(const GemmArgs g) {
    extern __shared__ __align__(128) unsigned char dsm[];
    __shared__ __align__(8) uint64_t bar[NSTAGE + 1];
    __shared__ uint32_t tmem_slot;
    constexpr int A_B = (TM / 8) * SBO, B_B = (TN / 8) * SBO, STAGE_B = 2 * A_B + 2 * B_B;
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31, z = blockIdx.z;
    const int i0 = blockIdx.x * TM, j0 = blockIdx.y * TN;

    if (w == 0) umma::tmem_alloc(&tmem_slot, TN);
    if (tid == 32)
        for (int b = 0; b <= NSTAGE; b++) umma::mbar_init(&bar[b], 1);
    const int nch = (g.Kc + GKT - 1) / GKT;
    float ra[NSET][TM / 32][4], rb[NSET][TN / 32][4];
#pragma unroll
    for (int p = 0; p < NSET - 1; p++)
        if (p < nch) {
            fetch_op<TM, AO>(g.A, z, i0, g.Mo, g.Kc, p, w, lane, ra[p]);
            fetch_op<TN, BO>(g.B, z, j0, g.No, g.Kc, p, w, lane, rb[p]);
        }
    umma::fence_before_thread_sync();
    __syncthreads();
    umma::fence_after_thread_sync();
    const uint32_t tmem = tmem_slot;

    for (int c0 = 0; c0 < nch; c0 += NSET) {
#pragma unroll
        for (int u = 0; u < NSET; u++) {
            const int c = c0 + u;
            if (c < nch) {
            if (c + NSET - 1 < nch) {
                fetch_op<TM, AO>(g.A, z, i0, g.Mo, g.Kc, c + NSET - 1, w, lane, ra[(u + NSET - 1) % NSET]);
                fetch_op<TN, BO>(g.B, z, j0, g.No, g.Kc, c + NSET - 1, w, lane, rb[(u + NSET - 1) % NSET]);
            }
            static_assert(NSET == NSTAGE, "stage index == register set index");
            if (c >= NSTAGE) umma::mbar_wait(&bar[u], ((c / NSTAGE) - 1) & 1);   // the MMAs that read this stage are done
            unsigned char *st = dsm + u * STAGE_B;
            store_op<TM, AO>(g.A, i0, g.Mo, g.Kc, c, w, lane, ra[u], st, st + A_B);
            store_op<TN, BO>(g.B, j0, g.No, g.Kc, c, w, lane, rb[u], st + 2 * A_B, st + 2 * A_B + B_B);
            umma::fence_async_smem();
            __syncthreads();
            if (tid == 0) {
                umma::fence_after_thread_sync();
                const uint32_t a = umma::smem_u32(st);
                const umma::Tile ah{a, LBO, SBO}, al{a + A_B, LBO, SBO}, bh{a + 2 * A_B, LBO, SBO}, bl{a + 2 * A_B + B_B, LBO, SBO};
                umma::gemm3(tmem, ah, al, bh, bl, TM, TN, GKT, c > 0);
                umma::mma_commit(&bar[u]);
            }
            }
        }
    }
    if (tid == 0) umma::mma_commit(&bar[NSTAGE]);
    umma::mbar_wait(&bar[NSTAGE], 0);
    umma::fence_after_thread_sync();

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
    __syncthreads();
    if (w == 0) umma::tmem_dealloc(tmem, TN);
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
    return cudaSuccess;
}

// engine: -1 = the library default (g_contraction_engine), 0 = SIMT, 1 = tcgen05 (TN by shape), 64 / 32 = tcgen05 with that TN
bool gemm_tc_launch(const GemmArgs &g, int nets, bool ao, bool bo, cudaStream_t st, int engine) {
    if (engine < 0) engine = g_contraction_engine;
    if (engine == 0 || g.Mo <= 0 || g.No <= 0 || g.Kc <= 0) return false;
    if (!ao && bo) return false;   // not a shape the learners use
    const int tn = engine == 64 || engine == 32 ? engine : (g.No > 32 ? 64 : 32);
    if (ao && bo) launch_tn<true, true>(g, nets, st, tn);
    else if (ao) launch_tn<true, false>(g, nets, st, tn);
    else launch_tn<false, false>(g, nets, st, tn);
    return true;
}

}  // namespace prl

extern "C" int prl_set_contraction_engine(int engine) {
    PRL_REQUIRE(engine == 0 || engine == 1, "engine must be 0 (SIMT tiles) or 1 (tcgen05 tiles)");
    prl::g_contraction_engine = engine;
    return PRL_OK;
}
extern "C" int prl_get_contraction_engine(void) { return prl::g_contraction_engine; }

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
    PRL_CUDA(gemm_tc_prepare());
    GemmLauncher L;
    L.st = (cudaStream_t)stream;
    L.engine = engine;
    const long long MK = (long long)M * K, MN = (long long)M * N, NK = (long long)N * K;
    if (op == 0) {
        Mat X = a2 ? mat2(a, split, split, a2, K - split, (long long)M * split, (long long)M * (K - split)) : mat(a, K, MK);
        L.fwd(X, M, b, K, NK, bias, N, N, K, relu != 0, c, N, MN, nets);
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
