"""oracle/per_oracle.py — CPU restatement of the prioritized replay sampler (TEST INFRASTRUCTURE ONLY).

The reference (facebookresearch/Pearl @ 48f1fbb) contains NO prioritized replay:
**parity is unpinned by the reference**.  This file is the specification the CUDA implementation
(pearl_b200/csrc/per.cu) is measured against — proportional prioritization (Schaul et al. 2016):

  * leaf priority p_i = (|td_i| + eps)^alpha, fp32; new transitions enter at the running maximum.
  * sum tree / min tree over C2 = next power of two >= capacity leaves, node i = fp32(node 2i (+|min) node 2i+1):
    fixed left-to-right pairwise order, so CPU and GPU trees are bit-identical given the same leaves.
  * stratified draws: u_k = (k + U_k) * (total / B), U_k = (philox4x32-10(counter=(k, step, 0, 0),
    key=(seed_lo, seed_hi))[0] >> 8) * 2^-24; prefix descent `u < left ? left : (u -= left, right)`
    (never into an empty right subtree).
  * importance weights w_i = (p_min / p_i)^beta  (= (N P(i))^-beta / max_j (N P(j))^-beta), delivered in
    TransitionBatch.weight (pearl/replay_buffers/transition.py:128) and applied to the squared TD error.
Leaf index = physical ring slot of the transition.
"""
from __future__ import annotations

import numpy as np

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85


def philox4x32_10(counter, key):
    c = [int(x) & 0xFFFFFFFF for x in counter]
    k = [int(x) & 0xFFFFFFFF for x in key]
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        hi0, lo0, hi1, lo1 = p0 >> 32, p0 & 0xFFFFFFFF, p1 >> 32, p1 & 0xFFFFFFFF
        c = [hi1 ^ c[1] ^ k[0], lo1, hi0 ^ c[3] ^ k[1], lo0]
        k = [(k[0] + W0) & 0xFFFFFFFF, (k[1] + W1) & 0xFFFFFFFF]
    return c


class PerOracle:
    def __init__(self, capacity, alpha=0.6, beta=0.4, eps=1e-6, seed=0):
        self.capacity = capacity
        self.C2 = 1
        while self.C2 < capacity:
            self.C2 *= 2
        self.alpha, self.beta, self.eps = np.float32(alpha), np.float32(beta), np.float32(eps)
        self.key = (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
        self.sum = np.zeros(2 * self.C2, dtype=np.float32)
        self.min = np.full(2 * self.C2, np.inf, dtype=np.float32)
        self.max_priority = np.float32(1.0)

    def _fix(self, leaves):
        idx = np.unique(np.asarray(leaves, dtype=np.int64) + self.C2)
        while idx[0] > 1:
            idx = np.unique(idx >> 1)
            self.sum[idx] = self.sum[2 * idx] + self.sum[2 * idx + 1]
            self.min[idx] = np.minimum(self.min[2 * idx], self.min[2 * idx + 1])

    def set_leaves(self, slots, priorities):
        slots = np.asarray(slots, dtype=np.int64)
        p = np.asarray(priorities, dtype=np.float32)
        self.sum[slots + self.C2] = p
        self.min[slots + self.C2] = p
        self.max_priority = np.float32(max(self.max_priority, p.max()))
        self._fix(slots)

    def push(self, slots):
        self.set_leaves(slots, np.full(len(slots), self.max_priority, dtype=np.float32))

    def priority_of(self, td):
        return np.power(np.abs(np.asarray(td, dtype=np.float32)) + self.eps, self.alpha, dtype=np.float32)

    def sample(self, batch, step):
        total = self.sum[1]
        seg = np.float32(total / np.float32(batch))
        slots = np.zeros(batch, dtype=np.int64)
        for k in range(batch):
            x0 = philox4x32_10((k, step & 0xFFFFFFFF, (step >> 32) & 0xFFFFFFFF, 0), self.key)[0]
            U = np.float32(x0 >> 8) * np.float32(2.0 ** -24)
            u = np.float32((np.float32(k) + U) * seg)
            idx = 1
            while idx < self.C2:
                left = self.sum[2 * idx]
                if u < left or self.sum[2 * idx + 1] == 0:
                    idx = 2 * idx
                else:
                    u = np.float32(u - left)
                    idx = 2 * idx + 1
            slots[k] = idx - self.C2
        p = self.sum[slots + self.C2]
        w = np.power(self.min[1] / p, self.beta, dtype=np.float32)
        return slots, w
