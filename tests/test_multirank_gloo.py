"""N > 1 host logic on CPU: world_size-2 gloo.  Members are independent, so ranks take contiguous blocks of
members (the same partition libmbm_b200's n_devices sharding uses), no data-path collective; only the
node-step count (SUM) and the timing (MAX) are reduced, as bench.py does."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, str(ROOT))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import load
    from mbm_b200 import drivers as d, ensembles as E
    ora = load.load()
    pars = E.c4_mb_growth_resorption(6, res=48, rank=0, max_steps=200)       # the same 6 members on every rank
    per = (len(pars) + world - 1) // world
    mine = pars[rank * per:(rank + 1) * per]                                  # contiguous block
    ens, _ = d.build_couple_ensemble(mine, ora, cap_left=128, cap_right=160)
    assert ens.run(ora, max_steps=40) == 0
    ns = torch.tensor([float(ens.node_steps.sum())], dtype=torch.float64)
    tmax = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(ns, op=dist.ReduceOp.SUM)
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    np.savez(Path(out_dir) / f"rank{rank}.npz", C_left=ens.C_left, C_right=ens.C_right, res=ens.res, it=ens.it,
             ns=ns.numpy(), tmax=tmax.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process(tmp_path, oracle):
    from mbm_b200 import drivers as d, ensembles as E
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    pars = E.c4_mb_growth_resorption(6, res=48, rank=0, max_steps=200)
    ens, _ = d.build_couple_ensemble(pars, oracle, cap_left=128, cap_right=160)
    assert ens.run(oracle, max_steps=40) == 0
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    assert np.array_equal(np.concatenate([r0["C_left"], r1["C_left"]]), ens.C_left)
    assert np.array_equal(np.concatenate([r0["C_right"], r1["C_right"]]), ens.C_right)
    assert np.array_equal(np.concatenate([r0["res"], r1["res"]]), ens.res)
    assert np.array_equal(np.concatenate([r0["it"], r1["it"]]), ens.it)
    assert r0["ns"][0] == r1["ns"][0] == float(ens.node_steps.sum())           # SUM over ranks
    assert r0["tmax"][0] == r1["tmax"][0] == 2.0                               # MAX over ranks


def test_reference_arm_under_torchrun_uses_the_host_cores():
    """`bench.py --impl reference` launched as the driver launches it for N > 1: rank 0 alone prints ONE line, and it runs on
    all the host threads although torchrun exports OMP_NUM_THREADS=1 to its ranks (`cores` is what the OpenMP runtime confirms)."""
    import json
    import os
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    env = dict(os.environ); env.pop("OMP_NUM_THREADS", None)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                          "127.0.0.1", "--master-port", "29541", str(root / "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, env=env, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0)) and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
