// common.cuh — error plumbing and small device helpers shared by the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "../../include/pearl_b200.h"

namespace prl {

extern thread_local char g_err[512];

inline int fail(int code, const char *fmt, ...) __attribute__((format(printf, 2, 3)));
inline int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define PRL_CUDA(expr)                                                                        \
    do {                                                                                      \
        cudaError_t e__ = (expr);                                                             \
        if (e__ != cudaSuccess)                                                               \
            return prl::fail(PRL_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                             __FILE__, __LINE__);                                             \
    } while (0)

#define PRL_REQUIRE(cond, ...)                                  \
    do {                                                        \
        if (!(cond)) return prl::fail(PRL_EINVAL, __VA_ARGS__); \
    } while (0)

__host__ __device__ inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// 16-byte async copy global -> shared, bypassing L1 (data written by other SMs
// in the previous phase lives in L2).
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

__device__ __forceinline__ float4 ldcg4(const float *p) { return __ldcg(reinterpret_cast<const float4 *>(p)); }

}  // namespace prl

// ---- handle definitions (host side) ---------------------------------------
struct prl_buf {
    prl_buf_desc desc;
    prl_buf_layout lay;
    uint32_t *records;     // device, capacity * record_words
    uint32_t *mt_state;    // device uint32[625]
    int64_t len;           // number of valid records
    int64_t write_pos;     // physical slot the next push writes
    // pinned staging for host pushes (double buffered)
    uint32_t *stage[2];
    cudaEvent_t stage_done[2];
    int64_t stage_records;  // records per staging buffer
    int stage_next;
    int device;
    // shard of a replay distributed over `shard_world` ranks (prl_buf_set_shard; 0 / 1 = not sharded)
    int shard_rank, shard_world;
    int64_t g_pushed;      // transitions pushed to the LOGICAL buffer so far (global write counter)
};

// multi-GPU communicator (see include/pearl_b200.h)
struct prl_comm {
    int rank, world;
    int64_t slot_floats;       // floats per (parity, rank) inbox slot
    float *inbox;              // local: [2 parities][world][slot_floats] of (value f32, sequence u32)
    unsigned int *flags;       // local: [kCommFlags] monotonically increasing arrival counters
    float *peer_inbox[16];     // peer_inbox[p] = rank p's inbox as mapped here (self: local)
    unsigned int *peer_flags[16];
    unsigned long long exchanges;  // exchanges completed so far (identical on all ranks)
    bool opened;
};
static const int kCommFlags = 256;
