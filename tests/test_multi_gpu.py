"""N>1: host-side logic on CPU with gloo (world_size 2), and — on a box with >= 2 GPUs — the
data-parallel learner with the in-kernel NVLink gradient exchange against the oracle."""
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT


def _run_workers(script, nproc, extra_env=None, timeout=600):
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1", **(extra_env or {}))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 400), script]
    return subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=timeout, cwd=ROOT)


GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, %r)
import torch, torch.distributed as dist
dist.init_process_group("gloo")
from pearl_b200.dist import all_gather_bytes, shard_owner
rank, world = dist.get_rank(), dist.get_world_size()
blob = bytes([rank * 16 + i %% 16 for i in range(128)])
got = all_gather_bytes(blob)
assert len(got) == world and all(got[r] == bytes([r * 16 + i %% 16 for i in range(128)]) for r in range(world))
# interleaved shard ownership: balanced, FIFO-consistent
own = [shard_owner(g, world) for g in range(10)]
assert [o[0] for o in own] == [g %% world for g in range(10)] and own[5] == (5 %% world, 5 // world)
# max-over-ranks reduction used by bench.py
t = torch.tensor([float(rank + 1)])
dist.all_reduce(t, op=dist.ReduceOp.MAX)
assert t.item() == world
if rank == 0: print("GLOO_OK")
dist.destroy_process_group()
'''


def test_host_logic_world2_gloo(tmp_path):
    script = tmp_path / "gloo_worker.py"
    script.write_text(GLOO_WORKER % ROOT)
    out = _run_workers(str(script), 2, timeout=300)
    assert out.returncode == 0 and "GLOO_OK" in out.stdout, out.stderr[-3000:]


@pytest.mark.gpu
@pytest.mark.parametrize("double", [0, 1])
def test_data_parallel_learner_matches_oracle_on_concatenated_batch(double):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    out = _run_workers(os.path.join(ROOT, "tests", "dp_worker.py"), min(n, 4), {"DP_DOUBLE": str(double)})
    assert out.returncode == 0 and "DP_OK" in out.stdout, (out.stdout[-2000:], out.stderr[-4000:])
