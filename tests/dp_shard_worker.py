"""Worker for the sharded-replay data-parallel parity test (torch.distributed.run, one process per GPU).
SURVEY.md 8e: ONE logical replay buffer sharded by interleaved global write counter; every rank runs the SAME
MT19937 stream and so draws the same B global indices as a single GPU; it works on the rows it owns and the
in-kernel NVLink exchange sums the unnormalised partial gradients.  Rank 0 additionally holds the whole buffer
and runs the ordinary single-GPU learner: indices must be bit-identical, parameters within 1e-4 (and, against the
torch oracle on the very same batches, likewise); all ranks end with bit-identical parameters."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import pearl_b200
    from oracle import c_oracle
    from oracle.pearl_oracle import OracleDQN, flat, load_flat
    from oracle.synth import make_transitions

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    obs, A, hidden, B, rounds = 24, 6, (32, 24), 96, 14
    double = bool(int(os.environ.get("DP_DOUBLE", "0")))
    cap_local = 700
    cap = cap_local * world
    n_push = cap + 37 * world + 3          # the logical ring wraps, and the last push is not a multiple of world

    class Space:
        def __init__(self, k):
            self.n = k
            self.actions = [torch.tensor([i]) for i in range(k)]
            self.actions_batch = torch.arange(k).view(k, 1)

    d = make_transitions(n_push, obs, A, seed=2024)      # the same global stream on every rank
    t = lambda k, a, b: torch.from_numpy(d[k][a:b])
    keys = ("state", "action", "reward", "next_state", "terminated", "truncated")
    shard = pearl_b200.B200ReplayBuffer(cap_local, device=dev, rng="device")
    full = pearl_b200.B200ReplayBuffer(cap, device=dev, rng="device") if rank == 0 else None
    for a0, b0 in ((0, 501), (501, cap + 11), (cap + 11, n_push)):          # uneven pushes
        shard.push_batch_sharded(rank, world, *(t(k, a0, b0) for k in keys), max_number_actions=A)
        if full is not None:
            full.push_batch(*(t(k, a0, b0) for k in keys), max_number_actions=A)
    assert len(shard) == cap
    shard.seed(4242)
    cls = pearl_b200.B200DoubleDQN if double else pearl_b200.B200DeepQLearning
    mk = lambda: cls(state_dim=obs, action_space=Space(A), hidden_dims=list(hidden), training_rounds=rounds, batch_size=B,
                     target_update_freq=5, soft_update_tau=0.6,
                     action_representation_module=pearl_b200.OneHotActionTensorRepresentationModule(A)).to(dev)
    torch.manual_seed(5)                    # identical initial weights on every rank
    learner = mk()
    init_q, init_qt = flat(learner._Q).cpu().numpy(), flat(learner._Q_target).cpu().numpy()
    comm = pearl_b200.B200Communicator(learner.flat_parameters.numel() + 1, dev)
    learner.set_communicator(comm)
    rep = learner.learn(shard, trace=True)
    rep2 = learner.learn(shard, trace=True)            # exchange counters and stream carry over between launches
    idx = np.concatenate([rep["idx"].cpu().numpy(), rep2["idx"].cpu().numpy()])
    loss = np.asarray(rep["loss"] + rep2["loss"])
    payload = dict(idx=idx, params=learner.flat_parameters.cpu().numpy(), target=learner.flat_target_parameters.cpu().numpy(),
                   loss=loss, q=np.concatenate([rep["q"].cpu().numpy(), rep2["q"].cpu().numpy()]))
    gathered = [None] * world
    dist.all_gather_object(gathered, payload)
    if rank == 0:
        for r in range(1, world):
            assert np.array_equal(gathered[r]["idx"], idx), "ranks drew different indices"
            assert np.array_equal(gathered[r]["params"], gathered[0]["params"]), "ranks diverged"
            assert np.array_equal(gathered[r]["target"], gathered[0]["target"])
            assert np.array_equal(gathered[r]["loss"], loss)
        # (1) the index stream is the single-GPU one: CPython's random.sample over the whole logical buffer
        mt = c_oracle.MT(seed=4242)
        assert np.array_equal(idx, np.stack([mt.sample(cap, B) for _ in range(2 * rounds)])), "not the 1-GPU index stream"
        # (2) the ordinary single-GPU learner on the whole buffer
        full.seed(4242)
        solo = mk()
        load_flat(solo._Q, init_q)
        load_flat(solo._Q_target, init_qt)
        s1 = solo.learn(full, trace=True)
        s2 = solo.learn(full, trace=True)
        assert np.array_equal(np.concatenate([s1["idx"].cpu().numpy(), s2["idx"].cpu().numpy()]), idx)
        want = solo.flat_parameters.cpu().numpy()
        err = float(np.max(np.abs(gathered[0]["params"] - want) / (np.abs(want) + 1e-2)))
        print(f"sharded dp vs 1 GPU: world={world} double={double} max rel err {err:.3e}")
        np.testing.assert_allclose(gathered[0]["params"], want, rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(gathered[0]["target"], solo.flat_target_parameters.cpu().numpy(), rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(loss, np.asarray(s1["loss"] + s2["loss"]), rtol=1e-4)
        # each row's q was written by exactly one rank (its owner); the union is the solo q
        qsum = sum(g["q"] for g in gathered)   # non-owners leave zeros
        # (3) the torch oracle on the same batches (logical index j = j-th oldest of the last `cap` pushes)
        orc = OracleDQN(obs, A, hidden, batch_size=B, target_update_freq=5, tau=0.6, double=double, init_q=init_q, init_q_target=init_qt)
        eye = torch.eye(A).unsqueeze(0).expand(B, A, A)
        base = n_push - cap
        for r in range(2 * rounds):
            rows = base + idx[r]
            orc.training_steps += 1
            b = dict(state=torch.from_numpy(d["state"][rows]), action=orc._one_hot(torch.from_numpy(d["action"][rows])),
                     reward=torch.from_numpy(d["reward"][rows]), terminated=torch.from_numpy(d["terminated"][rows]),
                     next_state=torch.from_numpy(d["next_state"][rows]), next_available_actions=eye,
                     next_unavailable_actions_mask=torch.zeros((B, A), dtype=torch.bool))
            orc.learn_batch(b)
        np.testing.assert_allclose(gathered[0]["params"], flat(orc.Q).numpy(), rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(gathered[0]["target"], flat(orc.Qt).numpy(), rtol=1e-4, atol=1e-6)
        print("DP_SHARD_OK")
    dist.barrier()
    comm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
