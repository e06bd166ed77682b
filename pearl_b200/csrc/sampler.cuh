// sampler.cuh — K2: MT19937-exact sampler (CPython random.sample, both branches)
// as CTA-level device routines, shared by the stand-alone k_sample_indices kernel
// (replay_buffer.cu) and the producer CTA of the persistent learner kernels.
//
// Reference call site: pearl/replay_buffers/tensor_based_replay_buffer.py:276
// `random.sample(self.memory, batch_size)`; algorithm: CPython Lib/random.py:242-250
// (_randbelow_with_getrandbits), :359-452 (sample) and Modules/_randommodule.c
// (MT19937, getrandbits(k<=32) = genrand_uint32() >> (32-k)).
#pragma once
#include <stdint.h>

namespace prl {

constexpr int MT_N = 624, MT_M = 397;
constexpr int kSamplerThreads = 256;

struct SamplerState {       // lives in shared memory of the sampling CTA
    uint32_t mt[2][MT_N];   // double-buffered MT19937 block
    int cur;                // which mt[] holds the live block
    int pos;                // next unread word, 0..624
    int done;               // completed samples
    int cnt;                // accepted draws in the current sample
    uint32_t epoch;         // tags hash-table entries of the current sample
    int pool_ready;
};

struct SamplerParams {
    uint32_t n;             // population = len(buffer)
    int k;                  // sample size
    int use_pool;           // n <= setsize: pool branch
    uint32_t table_cap;     // set branch: hash table capacity (power of two >= 2k)
    long long head, capacity;
    int32_t *out_logical, *out_slot;   // [rounds][k], either may be null
};

__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}
__device__ __forceinline__ uint32_t mt_mix(uint32_t a, uint32_t b) {
    uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
    return (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}

// One MT19937 block regeneration ("twist") by the whole CTA: old -> nw.
// new[i] depends on old[i], old[i+1] and on old[i+397] (i<227) or new[i-227].
__device__ inline void mt_twist_cta(const uint32_t *old, uint32_t *nw) {
    const int t = threadIdx.x;
    if (t < MT_N - MT_M) nw[t] = old[t + MT_M] ^ mt_mix(old[t], old[t + 1]);
    __syncthreads();
    if (t < MT_N - MT_M) {
        int i = t + (MT_N - MT_M);  // 227..453
        nw[i] = nw[i - (MT_N - MT_M)] ^ mt_mix(old[i], old[i + 1]);
    }
    __syncthreads();
    if (t < MT_N - 1 - 2 * (MT_N - MT_M)) {
        int i = t + 2 * (MT_N - MT_M);  // 454..622
        nw[i] = nw[i - (MT_N - MT_M)] ^ mt_mix(old[i], old[i + 1]);
    }
    if (t == kSamplerThreads - 1) nw[MT_N - 1] = nw[MT_M - 1] ^ mt_mix(old[MT_N - 1], nw[0]);
    __syncthreads();
}

__device__ __forceinline__ int bit_length_u32(uint32_t n) { return 32 - __clz(n); }

__device__ __forceinline__ int32_t slot_of(uint32_t j, long long head, long long capacity) {
    long long s = head + (long long)j;
    if (s >= capacity) s -= capacity;
    return (int32_t)s;
}

// dyn: set branch -> uint64 table[table_cap]; pool branch -> int32 pool[n]
__device__ inline void sampler_init(SamplerState &S, const uint32_t *__restrict__ mt_state, void *dyn,
                                    const SamplerParams &p) {
    const int tid = threadIdx.x;
    for (int i = tid; i < MT_N; i += kSamplerThreads) S.mt[0][i] = mt_state[i];
    if (!p.use_pool) {
        unsigned long long *table = reinterpret_cast<unsigned long long *>(dyn);
        for (uint32_t i = tid; i < p.table_cap; i += kSamplerThreads) table[i] = 0ull;
    }
    if (tid == 0) {
        S.cur = 0;
        S.pos = (int)mt_state[MT_N];
        S.done = 0;
        S.cnt = 0;
        S.epoch = 1;
        S.pool_ready = 0;
    }
    __syncthreads();
}

// hand the advanced state back (same layout as random.getstate()[1])
__device__ inline void sampler_store(const SamplerState &S, uint32_t *__restrict__ mt_state) {
    const int tid = threadIdx.x;
    for (int i = tid; i < MT_N; i += kSamplerThreads) mt_state[i] = S.mt[S.cur][i];
    if (tid == 0) mt_state[MT_N] = (uint32_t)S.pos;
}

// Generate samples until S.done >= target (all threads of the CTA call this).
__device__ inline void sampler_advance(SamplerState &S, void *dyn, const SamplerParams &p, int target) {
    unsigned long long *table = reinterpret_cast<unsigned long long *>(dyn);
    int32_t *pool = reinterpret_cast<int32_t *>(dyn);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t n = p.n;
    const int k = p.k;
    if (k == 0) return;  // random.sample(.., 0) draws nothing
    __syncthreads();
    if (S.done >= target) return;
    const int shift = 32 - bit_length_u32(n);  // set branch: getrandbits(n.bit_length())
    while (true) {
        if (p.use_pool && !S.pool_ready)
            for (uint32_t i = tid; i < n; i += kSamplerThreads) pool[i] = (int32_t)i;
        if (S.pos >= MT_N) {  // uniform: S.pos only changes between barriers
            mt_twist_cta(S.mt[S.cur], S.mt[S.cur ^ 1]);
            if (tid == 0) { S.cur ^= 1; S.pos = 0; }
        }
        __syncthreads();
        if (p.use_pool) {
            // n <= setsize: partial Fisher-Yates on a pool copy, inherently
            // sequential (Lib/random.py:435-442); only small buffers get here.
            if (tid == 0) {
                S.pool_ready = 1;
                const uint32_t *mt = S.mt[S.cur];
                int pos = S.pos, cnt = S.cnt, done = S.done;
                while (pos < MT_N && done < target) {
                    uint32_t m = n - (uint32_t)cnt;  // randbelow(n - i)
                    uint32_t r = mt_temper(mt[pos++]) >> (32 - bit_length_u32(m));
                    if (r >= m) continue;
                    int32_t v = pool[r];
                    pool[r] = pool[m - 1];
                    size_t o = (size_t)done * k + cnt;
                    if (p.out_logical) p.out_logical[o] = v;
                    if (p.out_slot) p.out_slot[o] = slot_of((uint32_t)v, p.head, p.capacity);
                    if (++cnt == k) { cnt = 0; done++; S.pool_ready = 0; break; }
                }
                S.pos = pos; S.cnt = cnt; S.done = done;
            }
        } else if (warp == 0) {
            // n > setsize: j = randbelow(n) until j not yet selected
            // (Lib/random.py:443-451).  A word is consumed per draw whether it
            // is accepted, out of range or a duplicate, so the accepted
            // sequence is an order-preserving compaction of the word stream.
            const uint32_t *mt = S.mt[S.cur];
            int pos = S.pos, cnt = S.cnt, done = S.done;
            uint32_t epoch = S.epoch;
            const unsigned lt = (1u << lane) - 1u;
            while (pos < MT_N && done < target) {
                const int w = pos + lane;
                const bool inb = w < MT_N;
                const uint32_t r = inb ? (mt_temper(mt[w]) >> shift) : 0xffffffffu;
                bool cand = inb && r < n;
                uint32_t h = (r * 2654435761u) & (p.table_cap - 1);
                if (cand) {  // already selected in an earlier chunk of this sample?
                    while (true) {
                        unsigned long long e = table[h];
                        if ((uint32_t)(e >> 32) != epoch) break;
                        if ((uint32_t)e == r) { cand = false; break; }
                        h = (h + 1) & (p.table_cap - 1);
                    }
                }
                // duplicates inside the chunk: the earliest word wins.  31 shuffles
                // (match.any would serialise over the 32 distinct values).
                const uint32_t key = cand ? r : 0xffffffffu;
                bool first = cand;
#pragma unroll
                for (int dd = 1; dd < 32; dd++) {
                    const uint32_t o = __shfl_up_sync(0xffffffffu, key, dd);
                    if (lane >= dd && o == key) first = false;
                }
                first = first && cand;
                const unsigned am = __ballot_sync(0xffffffffu, first);
                const int total = __popc(am), need = k - cnt;
                int consumed;
                unsigned take;
                bool finished = false;
                if (total >= need) {
                    // lane of the need-th accepted word: strip the need-1 lowest set bits
                    unsigned m2 = am;
                    for (int q = 1; q < need; q++) m2 &= m2 - 1;
                    const int last = __ffs(m2) - 1;
                    take = am & ((last == 31) ? 0xffffffffu : ((2u << last) - 1u));
                    consumed = last + 1;
                    finished = true;
                } else {
                    take = am;
                    consumed = min(32, MT_N - pos);
                }
                if ((take >> lane) & 1u) {
                    const int rank = __popc(take & lt);
                    size_t o = (size_t)done * k + cnt + rank;
                    if (p.out_logical) p.out_logical[o] = (int32_t)r;
                    if (p.out_slot) p.out_slot[o] = slot_of(r, p.head, p.capacity);
                    if (!finished) {  // remember it for the rest of this sample
                        const unsigned long long mine = ((unsigned long long)epoch << 32) | r;
                        while (true) {
                            unsigned long long e = table[h];
                            if ((uint32_t)(e >> 32) == epoch) { h = (h + 1) & (p.table_cap - 1); continue; }
                            if (atomicCAS(&table[h], e, mine) == e) break;
                        }
                    }
                }
                __syncwarp();
                pos += consumed;
                if (finished) { done++; cnt = 0; epoch++; } else { cnt += __popc(take); }
            }
            if (lane == 0) { S.pos = pos; S.cnt = cnt; S.done = done; S.epoch = epoch; }
        }
        __syncthreads();
        if (S.done >= target) break;
    }
}

}  // namespace prl
