// umma.cuh — thin inline-PTX layer over Blackwell's 5th-generation tensor cores
// (tcgen05.mma with TMEM accumulators) for the dense contractions of the learner.
//
// fp32 parity needs more than one TF32 pass: every fp32 operand x is split
//   x = hi + lo,  hi = trunc_tf32(x) (done by the hardware),  lo = x - hi
// and a product is accumulated as hi*hi + hi*lo + lo*hi in the fp32 TMEM
// accumulator ("3xTF32"): relative error ~2^-21 per product instead of 2^-11.
//
// Operand tiles live in shared memory in the canonical K-major NO-SWIZZLE
// ("interleaved") UMMA layout, written directly by SIMT code:
//   16-byte chunk c = k/4 of row r sits at  (r/8)*SBO + c*LBO + (r%8)*16
//   (8 rows x 16 B = one 128-byte core matrix; LBO = 128 B so the core matrices
//    of one 8-row group are contiguous along K; SBO = (K/4)*128 B)
// One tcgen05.mma.kind::tf32 consumes K = 8 (two chunks); the next K step
// advances the descriptor start address by 2*LBO.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- operand layout -------------------------------------------------------
constexpr int LBO = 128;  // bytes between the two 16-byte K chunks of one MMA / consecutive core matrices
__host__ __device__ constexpr int tile_sbo(int K) { return (K / 4) * 128; }           // bytes per 8-row group
__host__ __device__ constexpr int tile_bytes(int rows, int K) { return (rows / 8) * tile_sbo(K); }
// float index of element (r, k) inside a tile of K columns
__device__ __forceinline__ int tile_index(int r, int k, int K) {
    return (r >> 3) * (tile_sbo(K) >> 2) + (k >> 2) * 32 + (r & 7) * 4 + (k & 3);
}

__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}
// The tensor core TRUNCATES fp32 operands to TF32 (measured, tests/test_gpu_umma.py), so the "hi" operand
// can be x itself and lo = x - trunc_tf32(x) (exact in fp32; its own truncation costs 2^-21 relative).
__device__ __forceinline__ void split_tf32(float x, float &hi, float &lo) {
    hi = x;
    lo = x - __uint_as_float(__float_as_uint(x) & 0xffffe000u);
}

// ---- descriptors ------------------------------------------------------------
// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout): start address>>4 [0,14),
// LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout type [61,64) = 0 (no swizzle)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, int sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
    d |= (uint64_t)((LBO >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// instruction descriptor for kind::tf32, fp32 accumulate, both operands K-major
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]^T, M x N x 8
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}

// same with the A operand in tensor memory (lanes = rows, one 32-bit column per K element): no shared
// memory traffic for A, which at N = 64 is what bounds the SS form (6 KB of operands per MMA at 128 B/clk)
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t b_desc, uint32_t idesc,
                                            bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "r"(tmem_a), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}

// generic descriptor: explicit LBO / SBO (bytes)
__device__ __forceinline__ uint64_t make_desc2(uint32_t smem_addr, int lbo_bytes, int sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// K-major operand tile with a free chunk pitch.  Element (r, k) of a tile with K columns lives at byte
//   (r/8)*sbo + (k/4)*lbo + (r%8)*16 + (k%4)*4,   sbo = (K/4)*lbo.
// lbo = 128: dense (rows written by their owner thread as 16-byte chunks);
// lbo = 144: "transposed-write friendly": when 32 lanes write the SAME row r and consecutive k
//            (a thread that owns batch row k scatters its values into column k of the tile), the
//            addresses (k/4)*144 + (k%4)*4 hit 32 distinct banks.
// MN-major tf32 operands are NOT usable with the no-swizzle layout (measured: garbage), so the backward
// pass builds explicitly transposed tiles instead.
struct Tile {
    uint32_t addr;  // shared-memory byte address
    int lbo, sbo;   // bytes
    __device__ __forceinline__ uint64_t desc(int kstep) const { return make_desc2(addr + (uint32_t)kstep * 2 * lbo, lbo, sbo); }
    __device__ __forceinline__ Tile rows_from(int r) const { return Tile{addr + (uint32_t)(r >> 3) * sbo, lbo, sbo}; }  // r % 8 == 0
    __device__ __forceinline__ Tile shifted(uint32_t bytes) const { return Tile{addr + bytes, lbo, sbo}; }
};
__host__ __device__ constexpr int tile_bytes2(int rows, int K, int lbo) { return (rows / 8) * (K / 4) * lbo; }
__device__ __forceinline__ int tile_index2(int r, int k, int K, int lbo) {
    return (r >> 3) * ((K >> 2) * (lbo >> 2)) + (k >> 2) * (lbo >> 2) + (r & 7) * 4 + (k & 3);
}
__device__ __forceinline__ Tile make_tile(const void *p, int K, int lbo) { return Tile{smem_u32(p), lbo, (K >> 2) * lbo}; }

// One lane of a CONVERGED warp (call it under a warp-uniform condition, all 32 lanes executing).  MMA issue code must sit
// under `if (elect_one())`, not under `if (lane == 0)`: tcgen05.mma is a uniform-datapath instruction (UTCHMMA, operands in
// uniform registers).  Behind a thread-index predicate nvcc wraps EVERY mma in a serialisation loop over the active lanes
// (ELECT + 4 R2UR.BROADCAST + PLOP3 + BRA.U.ANY: ~13 dependent instructions, 53 clk per 128 x 64 x 8 product against a
// tensor-pipe floor of 32); behind elect.sync it knows exactly one lane runs and emits back-to-back UTCHMMA with the
// descriptor arithmetic on the uniform datapath (ncu source view: profiles/r2_k_dqn_tc_stalls.md).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
    return pred != 0;
}

// 3xTF32: D[M x N] (+)= A[M x K] * B[N x K]^T; issued by ONE thread.  *_exact: the operand is exactly
// representable in TF32 (0/1 indicators), its lo tile is not needed.  Descriptors are built once; a K
// step of 8 only adds (2*lbo)>>4 to the 14-bit start-address field.
__device__ __forceinline__ void gemm3(uint32_t tmem_d, Tile a_hi, Tile a_lo, Tile b_hi, Tile b_lo, int M, int N, int K,
                                      bool accumulate_first, bool a_exact = false, bool b_exact = false) {
    const uint32_t idesc = make_idesc_tf32(M, N);
    uint64_t ah = a_hi.desc(0), al = a_lo.desc(0), bh = b_hi.desc(0), bl = b_lo.desc(0);
    const uint64_t da = (uint64_t)((2 * a_hi.lbo) >> 4), db = (uint64_t)((2 * b_hi.lbo) >> 4);
    bool acc = accumulate_first;
    for (int ks = 0; ks < (K >> 3); ks++) {
        if (!a_exact) { mma_tf32(tmem_d, al, bh, idesc, acc); acc = true; }
        if (!b_exact) { mma_tf32(tmem_d, ah, bl, idesc, acc); acc = true; }
        mma_tf32(tmem_d, ah, bh, idesc, acc);
        acc = true;
        ah += da; al += da; bh += db; bl += db;
    }
}

// 3xTF32 with A (hi / lo) in tensor memory: a_hi / a_lo are TMEM column addresses of [M lanes][K columns]
__device__ __forceinline__ void gemm3_ts(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, Tile b_hi, Tile b_lo, int M, int N,
                                         int K, bool accumulate_first) {
    const uint32_t idesc = make_idesc_tf32(M, N);
    uint64_t bh = b_hi.desc(0), bl = b_lo.desc(0);
    const uint64_t db = (uint64_t)((2 * b_hi.lbo) >> 4);
    bool acc = accumulate_first;
    for (int ks = 0; ks < (K >> 3); ks++) {
        mma_tf32_ts(tmem_d, a_lo + ks * 8, bh, idesc, acc);
        mma_tf32_ts(tmem_d, a_hi + ks * 8, bl, idesc, true);
        mma_tf32_ts(tmem_d, a_hi + ks * 8, bh, idesc, true);
        acc = true;
        bh += db; bl += db;
    }
}

// 3xTF32 product of one [M x K] A tile pair and one [N x K] B tile pair into a TMEM accumulator.
// Issued by ONE thread.  a_hi/a_lo/b_hi/b_lo: shared addresses of the tiles.
__device__ __forceinline__ void gemm_3xtf32(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi,
                                            uint32_t b_lo, int M, int N, int K, bool accumulate_first) {
    const uint32_t idesc = make_idesc_tf32(M, N);
    const int sbo = tile_sbo(K);
    bool acc = accumulate_first;
    for (int k = 0; k < K; k += 8) {
        const uint32_t o = (uint32_t)(k >> 2) * LBO;
        const uint64_t ah = make_desc(a_hi + o, sbo), al = make_desc(a_lo + o, sbo);
        const uint64_t bh = make_desc(b_hi + o, sbo), bl = make_desc(b_lo + o, sbo);
        mma_tf32(tmem_d, al, bh, idesc, acc);   // small terms first
        mma_tf32(tmem_d, ah, bl, idesc, true);
        mma_tf32(tmem_d, ah, bh, idesc, true);
        acc = true;
    }
}

// ---- TMEM management ----------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_result, uint32_t ncols) {  // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // the allocating warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_thread_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_thread_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// make generic-proxy shared-memory writes visible to the tensor core (async proxy)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- mbarrier (completion of committed MMAs) -------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t *bar) {  // arrives on `bar` when all prior MMAs of this thread finish
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// Bounded spin: a wait that cannot complete (protocol bug, lost TMA transaction) ends the kernel with a trap and a
// message instead of hanging the device — over a second of polling, orders of magnitude beyond any legitimate wait.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    uint32_t done;
    long long t0 = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}\n"
            : "=r"(done)
            : "r"(a), "r"(parity)
            : "memory");
        if (!done) {
            if (t0 == 0) t0 = clock64();
            if (clock64() - t0 < 3000000000ll) continue;   // ~1.5 s of SM clocks
            printf("pearl_b200: mbarrier wait timed out: block %d thread %d barrier@%u parity %u\n", (int)blockIdx.x,
                   (int)threadIdx.x, a, parity);
            __trap();
        }
    } while (!done);
}

// ---- accumulator read-back: this warp's 32 lanes (rows) x 32 consecutive fp32 columns -------------
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float *v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}

// registers -> tensor memory: this warp's 32 lanes (rows) x 32 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float *v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
          "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])),
          "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])),
          "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
          "r"(__float_as_uint(v[15])), "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])),
          "r"(__float_as_uint(v[19])), "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])),
          "r"(__float_as_uint(v[23])), "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])),
          "r"(__float_as_uint(v[27])), "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])),
          "r"(__float_as_uint(v[31]))
        : "memory");
}
// registers -> tensor memory: this warp's 32 lanes (rows) x 16 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float *v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
          "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])),
          "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])),
          "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
          "r"(__float_as_uint(v[15]))
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

}  // namespace umma
