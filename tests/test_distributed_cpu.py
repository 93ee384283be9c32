"""world_size-2 gloo test of the multi-rank host logic (sharding, barrier, max-over-ranks, gather) on CPU."""

import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import sys, torch
    sys.path.insert(0, %r)
    from progressive_stable_diffusion_b200 import parallel
    rank, local, world = parallel.init_from_env(backend="gloo")
    units = [(p, l) for p in range(3) for l in range(13)]          # 3 patients x 13 MES levels
    mine = parallel.shard_indices(len(units), rank, world)
    parallel.barrier()
    total = parallel.sum_over_ranks(float(len(mine)))
    slowest = parallel.max_over_ranks(1.0 + rank)
    payload = torch.full((2, 3), float(rank))
    gathered = parallel.gather_to_rank0(payload)
    if rank == 0:
        assert gathered.shape == (4, 3) and gathered[:2].eq(0).all() and gathered[2:].eq(1).all()
        print("OK", int(total), slowest, len(mine))
    else:
        assert gathered is None
""") % ROOT


def test_two_rank_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29617", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "OK 39 2.0 20" in outs[0], outs
