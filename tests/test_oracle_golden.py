"""The oracle (oracle/) against fixtures recorded from the real reference.

CPU only.  Pins (a) the C restatement of CPython's MT19937 / random.sample
against known answers produced by CPython itself and (b) the torch restatement
of DeepQLearning / DoubleDQN `learn()` against q, y, loss, parameters and AdamW
state recorded from /root/reference by oracle/gen_golden.py.
"""
import glob
import json
import os
import random

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import c_oracle
from oracle.pearl_oracle import OracleDQN, OracleReplayBuffer, flat
from oracle.synth import from_fixture

KAT = json.load(open(os.path.join(GOLDEN, "random_sample_kat.json")))["cases"]
CASES = sorted(os.path.basename(p)[:-4] for pat in ("dqn_*.npz", "ddqn_*.npz") for p in glob.glob(os.path.join(GOLDEN, pat)))


@pytest.mark.parametrize("case", KAT, ids=lambda c: f"seed{c['seed']}_n{c['n']}_k{c['k']}")
def test_c_mt19937_sample_known_answers(case):
    mt = c_oracle.MT(seed=case["seed"])
    assert int(np.bitwise_xor.reduce(mt.st[:624])) == case["state_before_xor"]
    assert [mt.getrandbits32() for _ in range(4)] == case["first_words"]
    mt = c_oracle.MT(seed=case["seed"])
    for want in case["samples"]:
        got = mt.sample(case["n"], case["k"])
        assert got.tolist() == want
    assert int(mt.st[624]) == case["state_after_index"]
    assert int(np.bitwise_xor.reduce(mt.st[:624])) == case["state_after_xor"]
    assert mt.getrandbits32() == case["next_word_after"]


def test_c_sample_matches_this_interpreter():
    """Live cross-check against the CPython running the tests."""
    for seed, n, k in [(3, 1000, 100), (4, 50, 50), (5, 2_000_003, 300), (6, 17, 0)]:
        random.seed(seed)
        want = random.sample(range(n), k)
        mt = c_oracle.MT(seed=seed)
        assert mt.sample(n, k).tolist() == want
        st = np.asarray(random.getstate()[1], dtype=np.uint64).astype(np.uint32)
        assert np.array_equal(st, mt.st)


def test_c_sample_rejects_oversized():
    with pytest.raises(ValueError):
        c_oracle.MT(seed=1).sample(5, 6)


def test_setsize_thresholds():
    L = c_oracle.lib()
    assert [L.orc_sample_setsize(k) for k in (5, 6, 32, 256, 512)] == [21, 85, 277, 1045, 4117]


def replay_case(name):
    fx = np.load(os.path.join(GOLDEN, name + ".npz"))
    cfg = json.loads(bytes(fx["config"]).decode())
    data = from_fixture(fx)
    buf = OracleReplayBuffer(cfg["capacity"], cfg["n_act"])
    for i in range(cfg["n_push"]):
        ids = None
        if cfg["dynamic"]:
            ids = data["next_avail_ids"][i, : data["next_avail_n"][i]]
        buf.push(torch.from_numpy(data["state"][i]), int(data["action"][i]), float(data["reward"][i]),
                 bool(data["terminated"][i]), bool(data["truncated"][i]),
                 torch.from_numpy(data["next_state"][i]), ids)
    dqn = OracleDQN(cfg["obs"], cfg["n_act"], cfg["hidden"], lr=cfg["lr"], gamma=cfg["gamma"],
                    batch_size=cfg["batch"], training_rounds=cfg["rounds"],
                    target_update_freq=cfg["target_update_freq"], tau=cfg["tau"],
                    double=cfg["double"], init_q=fx["init_q"], init_q_target=fx["init_q_target"])
    return fx, cfg, buf, dqn


@pytest.mark.parametrize("name", CASES)
def test_torch_oracle_reproduces_reference(name):
    torch.set_num_threads(1)
    fx, cfg, buf, dqn = replay_case(name)
    st = fx["mt_state_before"]
    random.setstate((3, tuple(int(x) for x in st), None))
    dqn.trace = {"q": [], "y": []}
    snaps = {}
    mae = []
    # one learn() per reference learn() call, snapshotting after each round
    orig = dqn.learn_batch

    def spy(b):
        out = orig(b)
        r = len(mae) + 1
        if r in cfg["snap_rounds"]:
            snaps[r] = (flat(dqn.Q).numpy().copy(), flat(dqn.Qt).numpy().copy())
        mae.append(out)
        return out

    dqn.learn_batch = spy
    for _ in range(cfg["learn_calls"]):
        dqn.learn(buf)
    after = np.asarray(random.getstate()[1], dtype=np.uint64).astype(np.uint32)
    assert np.array_equal(after, fx["mt_state_after"]), "index stream diverged"
    tol = dict(rtol=2e-6, atol=2e-7)
    np.testing.assert_allclose(torch.stack(dqn.trace["q"]).numpy(), fx["q"], **tol)
    np.testing.assert_allclose(torch.stack(dqn.trace["y"]).numpy(), fx["y"], **tol)
    np.testing.assert_allclose(np.asarray(mae), fx["mae"], rtol=2e-6)
    for r in cfg["snap_rounds"]:
        np.testing.assert_allclose(snaps[r][0], fx[f"q_after_{r}"], **tol)
        np.testing.assert_allclose(snaps[r][1], fx[f"qt_after_{r}"], **tol)
    st_ = dqn.opt.state_dict()["state"]
    for key in ("exp_avg", "exp_avg_sq", "max_exp_avg_sq"):
        got = np.concatenate([st_[i][key].numpy().ravel() for i in range(len(st_))])
        np.testing.assert_allclose(got, fx[key], rtol=2e-6, atol=1e-12)


@pytest.mark.parametrize("name", CASES)
def test_c_sampler_reproduces_reference_indices(name):
    """C MT19937/sample restatement vs the indices the reference's learn() drew."""
    fx = np.load(os.path.join(GOLDEN, name + ".npz"))
    cfg = json.loads(bytes(fx["config"]).decode())
    n = min(cfg["n_push"], cfg["capacity"])
    k = min(cfg["batch"], n)
    mt = c_oracle.MT(state=fx["mt_state_before"])
    for r in range(fx["idx"].shape[0]):
        assert mt.sample(n, k).tolist() == fx["idx"][r].tolist()
    assert np.array_equal(mt.st, fx["mt_state_after"])
