"""world_size-2 (or more) CPU worker of tests/test_dist_cpu.py, launched by torch.distributed.run with the gloo backend.

Checks the host side of the x-slab sharding without a GPU:
  * every rank's plan (ogn_shard_plan: the same arithmetic the layers use) tiles the hidden rows exactly, its learn range
    is the set of visible rows its slab's fields touch, and its R replicas are every hidden row covering that range;
  * the exchange contract of ComputeSystem::allGather — buffer [world][count] int32, own slab in place, complete on every
    rank afterwards — reproduces on each rank the CSDR the CPU oracle computes unsharded, when each rank contributes only
    the rows it owns.
"""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    import ogmaneo2_b200 as og
    from ogmaneo2_b200 import configs
    from ogmaneo2_b200.flat_api import FlatHierarchy
    from ogmaneo2_b200.hierarchy import _bind_product
    from common import make_inputs
    from oracle import loader

    flat = _bind_product(og.lib())
    H, V, r = 32, 32, 4
    plan = (C.c_int * 6)()
    assert flat.shard_plan(H, V, r, rank, world, plan) == 0
    mine = list(plan)
    plans = [None] * world
    dist.all_gather_object(plans, mine)
    # owned hidden rows tile [0, H)
    edges = sorted((p[0], p[1]) for p in plans)
    assert edges[0][0] == 0 and edges[-1][1] == H and all(a[1] == b[0] for a, b in zip(edges, edges[1:])), plans
    x0, x1, vx0, vx1, rx0, rx1 = mine
    centre = lambda o: int(np.float32(o) * (np.float32(V) / np.float32(H)) + np.float32(0.5))
    touched = sorted({v for o in range(x0, x1) for v in range(max(0, centre(o) - r), min(V - 1, centre(o) + r) + 1)})
    assert touched == list(range(vx0, vx1)), (mine, touched)
    cover = sorted({o for o in range(H) for v in range(vx0, vx1) if max(0, centre(o) - r) <= v <= min(V - 1, centre(o) + r)})
    assert cover == list(range(rx0, rx1)), (mine, cover)
    assert rx0 <= x0 and rx1 >= x1
    bad = (C.c_int * 6)()
    assert flat.shard_plan(H + 1, V, r, rank, world, bad) != 0 or world == 1  # rows must divide evenly

    # exchange contract on real CSDRs: the oracle's hidden state, each rank contributing only its slab
    cfg = configs.c4_large(32)
    chk, _ = loader.best()
    h = FlatHierarchy(chk, cfg["inputSizes"], cfg["inputTypes"], cfg["layerDescs"], cfg["seed"])
    for t in range(3):
        h.step(make_inputs(cfg, t), True)
        full = np.asarray(h.scHiddenCs(0), dtype=np.int32).copy()
        count = (x1 - x0) * 32
        buf = torch.full((world * count,), -1, dtype=torch.int32)
        buf[rank * count:(rank + 1) * count] = torch.from_numpy(full[x0 * 32:x1 * 32])
        dist.all_gather_into_tensor(buf, buf[rank * count:(rank + 1) * count].clone())
        assert np.array_equal(buf.numpy(), full), f"rank {rank}: gathered CSDR differs from the unsharded one at step {t}"
    h.close()
    dist.barrier()
    dist.destroy_process_group()
    print(f"rank {rank}/{world} ok")


if __name__ == "__main__":
    main()
