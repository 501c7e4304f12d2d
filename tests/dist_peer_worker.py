"""One rank per GPU (torch.distributed.run, >= 2 GPUs): x-slab sharding with the library's peer-memory all-gather kernel
must reproduce the unsharded run on the same GPU bit for bit (predictions every step, hidden CSDRs, owned weights)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    dist.init_process_group("gloo")  # only carries the 64-byte IPC handles
    rank, world = dist.get_rank(), dist.get_world_size()
    torch.cuda.set_device(rank)
    import ogmaneo2_b200 as og
    from ogmaneo2_b200 import configs
    from ogmaneo2_b200.flat_api import LayerDesc
    from common import make_inputs

    cfg = configs.c4_large(32)
    cfg["layerDescs"].append(LayerDesc(hiddenSize=(16, 16, 16), ffRadius=2, pRadius=2, temporalHorizon=2))
    steps = int(os.environ.get("OGN_PEER_STEPS", "25"))
    whole = og.Hierarchy(cfg["inputSizes"], cfg["inputTypes"], cfg["layerDescs"], seed=9,
                         options=og.CreateOptions(device=rank, deviceInit=True, replayRng=False))
    part = og.Hierarchy(cfg["inputSizes"], cfg["inputTypes"], cfg["layerDescs"], seed=9,
                        options=og.CreateOptions(device=rank, deviceInit=True, replayRng=False, shardRank=rank, shardWorld=world,
                                                 peerExchange=True))
    handles = [None] * world
    dist.all_gather_object(handles, part.peerExport())
    part.peerAttach(handles)
    dist.barrier()
    for t in range(steps):
        ins = make_inputs(cfg, t)
        whole.step(ins, True)
        part.step(ins, True)
        if t % 3 == 0 or t == steps - 1:  # let the ranks drift apart in between: the arrival flags must order them
            a, b = part.getPredictionCs(0), whole.getPredictionCs(0)
            assert np.array_equal(a, b), f"rank {rank}: prediction differs at step {t} ({int((a != b).sum())} columns)"
    for l in range(whole.getNumLayers()):
        assert np.array_equal(part.scHiddenCs(l), whole.scHiddenCs(l)), f"rank {rank}: layer {l} hidden CSDR differs"
        w, full = part.scWeights(l, 0), whole.scWeights(l, 0)
        own = w != 0
        assert own.any() and np.array_equal(w[own].view(np.int32), full[own].view(np.int32)), f"rank {rank}: layer {l} weights differ"
    # per-rank checkpoint of the shard: restored into a fresh sharded hierarchy, it continues like the one that wrote it
    import tempfile
    path = os.path.join(tempfile.gettempdir(), f"ogn_peer_ckpt_{os.getpid()}_{rank}.ckpt")
    part.checkpointSave(path)
    again = og.Hierarchy.checkpointLoad(path, cfg["inputSizes"], options=og.CreateOptions(device=rank, replayRng=False, shardRank=rank,
                                                                                       shardWorld=world, peerExchange=True))
    os.remove(path)
    handles = [None] * world
    dist.all_gather_object(handles, again.peerExport())
    again.peerAttach(handles)
    dist.barrier()
    for t in range(steps, steps + 6):
        ins = make_inputs(cfg, t)
        whole.step(ins, True)
        again.step(ins, True)
        a, b = again.getPredictionCs(0), whole.getPredictionCs(0)
        assert np.array_equal(a, b), f"rank {rank}: restored shard differs at step {t}"
    again.synchronize()
    part.synchronize()
    dist.barrier()
    print(f"rank {rank}/{world} ok")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
