"""Deterministic synthetic transitions (TEST / BENCH INFRASTRUCTURE, not product).

Shapes and distributions follow BASELINE.md §3: state, next_state ~ N(0,1),
action = i mod A, reward ~ N(0,1), terminated ~ Bernoulli(0.02), truncated =
False.  Values are rounded to a 1/256 grid so that (a) a fixture can store them
as int16 and (b) they are exactly representable in fp32 on every machine.
Pure numpy; no reference import.
"""
from __future__ import annotations

import numpy as np

GRID = 256.0


def make_transitions(n: int, obs: int, n_act: int, seed: int, dynamic: bool = False,
                     p_term: float = 0.02) -> dict:
    rng = np.random.Generator(np.random.PCG64(seed))
    state_q = np.clip(np.rint(rng.standard_normal((n, obs)) * GRID), -32000, 32000).astype(np.int16)
    next_q = np.clip(np.rint(rng.standard_normal((n, obs)) * GRID), -32000, 32000).astype(np.int16)
    reward_q = np.clip(np.rint(rng.standard_normal(n) * GRID), -32000, 32000).astype(np.int16)
    terminated = rng.random(n) < p_term
    out = dict(
        state_q=state_q, next_state_q=next_q, reward_q=reward_q,
        state=(state_q.astype(np.float32) / GRID), next_state=(next_q.astype(np.float32) / GRID),
        reward=(reward_q.astype(np.float32) / GRID),
        action=(np.arange(n) % n_act).astype(np.int64),
        terminated=terminated, truncated=np.zeros(n, dtype=bool),
    )
    if dynamic:
        # a random non-empty, ordered subset of actions is available in the next state
        ids = np.zeros((n, n_act), dtype=np.int64)
        cnt = np.zeros(n, dtype=np.int64)
        for i in range(n):
            m = int(rng.integers(1, n_act + 1))
            sub = np.sort(rng.choice(n_act, size=m, replace=False))
            ids[i, :m] = sub
            cnt[i] = m
        out["next_avail_ids"] = ids
        out["next_avail_n"] = cnt
    return out


def from_fixture(fx) -> dict:
    """Rebuild the float transitions from a tests/golden/*.npz fixture."""
    out = dict(
        state=fx["state_q"].astype(np.float32) / GRID,
        next_state=fx["next_state_q"].astype(np.float32) / GRID,
        reward=fx["reward_q"].astype(np.float32) / GRID,
        action=fx["action"].astype(np.int64),
        terminated=fx["terminated"].astype(bool),
        truncated=fx["truncated"].astype(bool),
    )
    if "next_avail_ids" in fx:
        out["next_avail_ids"] = fx["next_avail_ids"].astype(np.int64)
        out["next_avail_n"] = fx["next_avail_n"].astype(np.int64)
    return out
