"""Worker for the multi-GPU data-parallel parity test (launched with torch.distributed.run,
one process per GPU).  Each rank: own replay shard + own MT19937 stream, same initial weights;
learn() with the in-kernel NVLink gradient exchange; rank 0 checks against the oracle run on the
CONCATENATED batches (batch W*B) and that all ranks hold bit-identical parameters."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import pearl_b200
    from oracle.pearl_oracle import OracleDQN, flat, load_flat
    from oracle.synth import make_transitions
    from oracle import c_oracle

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    obs, A, hidden, B, n, rounds = 24, 6, (32, 24), 64, 3000, 12
    double = bool(int(os.environ.get("DP_DOUBLE", "0")))

    class Space:
        def __init__(self, k):
            self.n = k
            self.actions = [torch.tensor([i]) for i in range(k)]
            self.actions_batch = torch.arange(k).view(k, 1)

    d = make_transitions(n, obs, A, seed=1000 + rank)
    buf = pearl_b200.B200ReplayBuffer(n, device=dev, rng="device")
    buf.push_batch(*(torch.from_numpy(d[k]) for k in ("state", "action", "reward", "next_state", "terminated", "truncated")),
                   max_number_actions=A)
    buf.seed(77 + rank)
    torch.manual_seed(5)  # identical initial weights on every rank
    cls = pearl_b200.B200DoubleDQN if double else pearl_b200.B200DeepQLearning
    learner = cls(state_dim=obs, action_space=Space(A), hidden_dims=list(hidden), training_rounds=rounds, batch_size=B,
                  target_update_freq=5, soft_update_tau=0.6,
                  action_representation_module=pearl_b200.OneHotActionTensorRepresentationModule(A)).to(dev)
    init_q, init_qt = flat(learner._Q).cpu().numpy(), flat(learner._Q_target).cpu().numpy()
    comm = pearl_b200.B200Communicator(learner.flat_parameters.numel(), dev)
    learner.set_communicator(comm)
    rep = learner.learn(buf, trace=True)
    idx = rep["idx"].cpu().numpy()
    # the per-rank index stream is still CPython's random.sample for that rank's seed
    mt = c_oracle.MT(seed=77 + rank)
    assert np.array_equal(idx, np.stack([mt.sample(n, B) for _ in range(rounds)]))
    # second call: exchange counters and parity carry over between launches
    rep2 = learner.learn(buf, trace=True)
    idx = np.concatenate([idx, rep2["idx"].cpu().numpy()])
    params = learner.flat_parameters.cpu().numpy()
    payload = dict(idx=idx, params=params, target=learner.flat_target_parameters.cpu().numpy(),
                   loss=np.asarray(rep["loss"] + rep2["loss"]), data={k: d[k] for k in ("state", "action", "reward", "next_state", "terminated")})
    gathered = [None] * world
    dist.all_gather_object(gathered, payload)
    ok = True
    if rank == 0:
        for r in range(1, world):
            assert np.array_equal(gathered[r]["params"], gathered[0]["params"]), "ranks diverged"
            assert np.array_equal(gathered[r]["target"], gathered[0]["target"])
        orc = OracleDQN(obs, A, hidden, batch_size=B * world, target_update_freq=5, tau=0.6, double=double,
                        init_q=init_q, init_q_target=init_qt)
        eye = torch.eye(A).unsqueeze(0).expand(B * world, A, A)
        for t in range(2 * rounds):
            cat = lambda key: torch.cat([torch.from_numpy(g["data"][key][g["idx"][t]]) for g in gathered])
            orc.training_steps += 1
            b = dict(state=cat("state"), action=orc._one_hot(cat("action")), reward=cat("reward"),
                     terminated=cat("terminated"), next_state=cat("next_state"), next_available_actions=eye,
                     next_unavailable_actions_mask=torch.zeros((B * world, A), dtype=torch.bool))
            orc.learn_batch(b)
        want = flat(orc.Q).numpy()
        err = float(np.max(np.abs(gathered[0]["params"] - want) / (np.abs(want) + 1e-2)))
        print(f"dp parity: world={world} double={double} max rel err {err:.3e}")
        np.testing.assert_allclose(gathered[0]["params"], want, rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(gathered[0]["target"], flat(orc.Qt).numpy(), rtol=1e-4, atol=1e-6)
        print("DP_OK")
    dist.barrier()
    comm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
