#!/usr/bin/env python
"""bench.py — hierarchy steps/s (learning on) of the B200 path, its HBM roofline, and the reference CPU path beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c4|c2|c1|c3|c5] [--impl reference]

One JSON line on stdout (rank 0). A "step" is one Hierarchy::step with learning on over one synthetic input CSDR.
  value   whole-job steps/s with the inputs already resident in HBM (ogn_step_device, CUDA events on the launch stream)
  e2e     the same through the public host API: Hierarchy.step(host CSDR) + getPredictionCs() every step, i.e. the H2D
          copy of the input from pinned memory and the D2H read of the prediction are inside the timed region
  roofline  dominant kernel: algorithmic bytes per launch (SURVEY.md §8d formulas) / its mean launch time measured with
          CUDA events inside the timed region, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the reference's own CPU implementation (oracle/_ref when it was compiled, else the oracle port) timed on
          this box's host cores on a bounded sample of the same workload

Default workload: BASELINE config 4 (the large-layer config north_star's target is stated on): 512x512x16 input ->
one layer of 512x512x32 hidden columns, radius 4, temporalHorizon 1, 2%-noisy drifting stream. With --gpus N > 1
(launched under torch.distributed.run, one rank per GPU) the same layer is sharded by hidden-column x-slab across the N
GPUs (strong scaling); the only exchange is an all-gather of int32 CSDR slabs over NCCL after the SparseCoder and after
the Predictor. PyTorch is used for plumbing only (events on the library's stream, NCCL, process group).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from ogmaneo2_b200 import configs, streams  # noqa: E402
from ogmaneo2_b200.flat_api import FlatHierarchy, InputType  # noqa: E402

METRIC = "hierarchy_steps_per_s_learn_on"
# dram__bytes_read.sum + dram__bytes_write.sum per launch of each kernel at the default workload (C4 512x512), from the
# `ncu --set full` capture summarised in profiles/README.md (filled in from the capture; None = not captured)
TRAFFIC_CSV = os.path.join(ROOT, "profiles", "ncu_c4_512_traffic.csv")
UNIT = "steps/s"


def load_traffic():
    """kernel kind -> DRAM bytes per launch at the default workload, read from the committed summary of the latest
    `ncu --set full` capture (profiles/ncu_c4_512_traffic.csv: kind,kernel,dram_bytes_read,dram_bytes_write,source)."""
    out, src = {}, {}
    try:
        with open(TRAFFIC_CSV) as f:
            for line in f:
                p = [x.strip() for x in line.split(",")]
                if len(p) >= 5 and p[0] != "kind" and not p[0].startswith("#"):
                    out[p[0]] = float(p[2]) + float(p[3])
                    src[p[0]] = p[4]
    except (OSError, ValueError):
        pass
    return out, src


# ----------------------------------------------------------------------------------------------------------------------
# workloads
def workload(name, n_gpus, size=None):
    if name == "c4":
        n = size or 512
        cfg = configs.c4_large(n)
        cfg["desc"] = f"C4: {n}x{n}x16 input CSDR -> 1 layer {n}x{n}x32 hidden columns, r=4, temporalHorizon=1, noisy(2%) stream"
        cfg["sample"] = configs.c4_large(128)  # SURVEY.md's oracle size for this config: ~20 s of reference init + ~10 s of steps
        cfg["device_init"] = True  # 2.2e10 weights: the reference cannot even construct this size
    elif name == "c2":
        cfg = configs.c2_video(32, 3)
        cfg["desc"] = "C2: 32x32x16 input CSDR -> 3 layers 32x32x32, r=3, prediction on every layer, noisy(2%) stream"
        cfg["sample"] = configs.c2_video(32, 3)
        cfg["device_init"] = False
    elif name == "c1":
        cfg = configs.c1_sine()
        cfg["desc"] = "C1: sine 1x1x32 -> 4 layers 4x4x16, r=2"
        cfg["sample"] = configs.c1_sine()
        cfg["device_init"] = False
    elif name == "c3":
        cfg = configs.c3_actor(8)
        cfg["desc"] = "C3: 8x8x16 observation + 2x2x4 action -> 3 layers 8x8x16, Actor head"
        cfg["sample"] = configs.c3_actor(8)
        cfg["device_init"] = False
    elif name == "c5":
        cfg = configs.c5_unit(16)
        cfg["desc"] = ("C5: ensemble of independent hierarchies stepped in lockstep, each 16x16x16 input -> 4 layers 16x16x16, r=2 "
                       "(member k: seed 1234+k, stream shifted by 17k)")
        cfg["sample"] = configs.c5_unit(16)
        cfg["device_init"] = True
    else:
        raise SystemExit(f"unknown workload {name}")
    cfg["key"] = name
    return cfg


# ----------------------------------------------------------------------------------------------------------------------
# algorithmic bytes: pairs P of a receptive-field geometry and per-kernel-kind bytes per step
def rf_pairs(hidden, visible, r):
    def axis(H, V):
        s = np.float32(V) / np.float32(H)
        tot = 0
        for o in range(H):
            c = int(np.float32(o) * s + np.float32(0.5))
            tot += max(0, min(V - 1, c + r) - max(0, c - r) + 1)
        return tot
    return axis(hidden[0], visible[0]) * axis(hidden[1], visible[1])


def bytes_per_launch(cfg):
    """[(layer, kind, bytes per launch excluding the data-dependent SC writes, bytes per updated pair)] per launch site."""
    L = len(cfg["layerDescs"])
    sites = []
    for l, ld in enumerate(cfg["layerDescs"]):
        H = ld.hiddenSize
        if l == 0:
            vis = [s for s in cfg["inputSizes"] for _ in range(ld.temporalHorizon)]
        else:
            vis = [cfg["layerDescs"][l - 1].hiddenSize] * ld.temporalHorizon
        fwd = sum(4 * rf_pairs(H, v, ld.ffRadius) * H[2] for v in vis)
        rec = sum(4 * rf_pairs(H, v, ld.ffRadius) * v[2] for v in vis)
        sites.append((l, "sc_forward", fwd, 0))
        sites.append((l, "sc_learn", rec, 4 * vis[0][2]))
        nvl = 2 if l < L - 1 else 1
        if l == 0:
            preds = [s for s, t in zip(cfg["inputSizes"], cfg["inputTypes"]) if t == InputType.prediction]
        else:
            preds = [cfg["layerDescs"][l - 1].hiddenSize] * ld.ticksPerUpdate
        for ph in preds:
            sites.append((l, "pred_step", 12 * rf_pairs(ph, H, ld.pRadius) * ph[2] * nvl, 0))
    return sites


def tick_counts(cfg, steps):
    """how many times each layer updates in `steps` steps from a fresh hierarchy (Hierarchy.cpp:219-249)"""
    L = len(cfg["layerDescs"])
    tpu = [1] + [ld.ticksPerUpdate for ld in cfg["layerDescs"][1:]]
    ticks = [0] * L
    counts = np.zeros((steps, L), dtype=np.int64)
    for s in range(steps):
        ticks[0] = 0
        for l in range(L):
            if l == 0 or ticks[l] >= tpu[l]:
                ticks[l] = 0
                counts[s, l] = 1
                if l < L - 1:
                    ticks[l + 1] += 1
    return counts


# ----------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        mhz, mx, reasons, power = [], None, set(), []
        for ts, line in self.rows:
            p = [x.strip() for x in line.split(",")]
            if len(p) < 7:
                continue
            try:
                if t0 - 0.05 <= ts <= t1 + 0.15:
                    mhz.append(float(p[0]))
                    power.append(float(p[2]))
                mx = float(p[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[3:7]):
                if v.lower().startswith("active") and t0 - 0.05 <= ts <= t1 + 0.15:
                    reasons.add(name)
        return {"sm_mhz": float(np.median(mhz)) if mhz else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "power_w_max": max(power) if power else None, "samples": len(mhz)}


class _DevArray:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (ptr, False), "version": 2}


# ----------------------------------------------------------------------------------------------------------------------
def time_cpu_members(members, steps, warmup, budget_s):
    """C5: `members` independent hierarchies (seed 1234 + k, stream shifted by 17 k) stepped one after the other by the
    reference with all host threads; returns (member-steps/s, info)."""
    from oracle import loader
    flat, kind = loader.best()
    cores = os.cpu_count() or 1
    flat.set_num_threads(cores)
    total_steps, total_t, t_init = 0, 0.0, 0.0
    for k in range(members):
        c = configs.c5_unit(16, k)
        t0 = time.perf_counter()
        h = FlatHierarchy(flat, c["inputSizes"], c["inputTypes"], c["layerDescs"], c["seed"])
        t_init += time.perf_counter() - t0
        ins = [[streams.make("noisy", t, c["inputSizes"][0], seed=c["noise_seed"], shift=c["shift"])] for t in range(warmup + steps)]
        for t in range(warmup):
            h.step(ins[t], True)
        t0 = time.perf_counter()
        done = 0
        for t in range(warmup, warmup + steps):
            h.step(ins[t], True)
            done += 1
            if time.perf_counter() - t0 > budget_s / members:
                break
        total_t += time.perf_counter() - t0
        total_steps += done
        h.close()
    v = total_steps / total_t
    return v, {"kind": kind, "cores": cores, "sample": f"C5: {members} hierarchies (seeds 1234+k) stepped one after the other, {total_steps} "
               f"member-steps in {total_t:.2f}s after {warmup} warm-up steps each (init {t_init:.1f}s); an ensemble of B members costs B x this",
               "sample_steps_per_s": v}


def time_cpu(cfg_sample, steps, warmup, full_cols, budget_s=25.0, want_trace=False):
    """The reference's CPU implementation on a bounded sample; returns (full-workload steps/s, info dict[, trace])."""
    from oracle import loader
    flat, kind = loader.best()
    cores = os.cpu_count() or 1
    has_actor = any(t == InputType.action for t in cfg_sample["inputTypes"])
    threads = 1 if has_actor else cores  # the reference's Actor path is only deterministic single-threaded
    flat.set_num_threads(threads)
    t0 = time.perf_counter()
    h = FlatHierarchy(flat, cfg_sample["inputSizes"], cfg_sample["inputTypes"], cfg_sample["layerDescs"], cfg_sample["seed"])
    t_init = time.perf_counter() - t0
    ins = lambda t: [streams.make(k if k != "action" else "iid", t, s) for s, k in zip(cfg_sample["inputSizes"], cfg_sample["streams"])]
    pred_input = next((i for i, t in enumerate(cfg_sample["inputTypes"]) if t == InputType.prediction), None)
    trace = {"inputs": ins, "pred": [], "pred_input": pred_input}
    keep = want_trace and pred_input is not None and not has_actor
    for t in range(warmup):
        h.step(ins(t), True)
        if keep:
            trace["pred"].append(h.getPredictionCs(pred_input))
    done, t0 = 0, time.perf_counter()
    pre = [ins(warmup + t) for t in range(steps)]
    untimed = 0.0
    t0 = time.perf_counter()
    for t in range(steps):
        h.step(pre[t], True)
        done += 1
        if keep:  # reading the prediction back is not part of the timed work
            u0 = time.perf_counter()
            trace["pred"].append(h.getPredictionCs(pred_input))
            untimed += time.perf_counter() - u0
        if time.perf_counter() - t0 - untimed > budget_s:
            break
    dt = time.perf_counter() - t0 - untimed
    if keep:
        trace["hidden0"] = h.scHiddenCs(0)
    sample_cols = sum(ld.hiddenSize[0] * ld.hiddenSize[1] for ld in cfg_sample["layerDescs"][:1])
    sample_sps = done / dt
    scale = sample_cols / full_cols  # column-steps/s is the size-independent rate for the single-layer config
    h.close()
    info = {"kind": kind, "cores": threads, "sample": f"{cfg_sample['name']}: {done} steps in {dt:.2f}s after {warmup} warm-up "
            f"(init {t_init:.1f}s), scaled by hidden columns {sample_cols}/{full_cols}", "sample_steps_per_s": sample_sps}
    if want_trace:
        return sample_sps * scale, info, (trace if keep else None)
    return sample_sps * scale, info


def state_hash(h, cfg, pred_idx, rank, world, dist, ens):
    """crc32 over the prediction CSDRs, every SparseCoder's hidden CSDR and a sample of weight columns (SparseCoder and
    Predictor, 8 columns per layer spread over the x-slabs, fetched from whichever rank stores them), taken after the last
    step of the run. The run is deterministic, so the hash must not depend on --gpus: N = 8 equal to N = 1 is the
    sharded-equals-unsharded check of SURVEY.md 8e on the full-size layer. Ensembles hash rank 0's members."""
    import ctypes as C
    import zlib
    lib = h.f.lib
    lib.ogn_debug_column_weights.restype = C.c_int64
    lib.ogn_debug_column_weights.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_int64]
    crc = 0
    for i in pred_idx:
        crc = zlib.crc32(np.ascontiguousarray(h.getPredictionCs(i)).tobytes(), crc)
    L = len(cfg["layerDescs"])
    for l in range(L):
        crc = zlib.crc32(np.ascontiguousarray(h.scHiddenCs(l)).tobytes(), crc)
    mine = {}
    if not ens:
        for l, ld in enumerate(cfg["layerDescs"]):
            Hx, Hy = ld.hiddenSize[0], ld.hiddenSize[1]
            # (kind, layer, grid x extent, grid y extent): the SparseCoder, and predictor 0 of the layer (layer 0: first
            # prediction input; its hidden grid is the input grid)
            targets = [(0, l, Hx, Hy)]
            if l == 0 and pred_idx and cfg["inputTypes"][pred_idx[0]] == InputType.prediction and pred_idx[0] == 0:
                targets.append((1, l, cfg["inputSizes"][0][0], cfg["inputSizes"][0][1]))
            elif l > 0:
                targets.append((1, l, cfg["layerDescs"][l - 1].hiddenSize[0], cfg["layerDescs"][l - 1].hiddenSize[1]))
            for kind, ll, gx, gy in targets:
                for k in range(8):
                    ox, oy = min(gx - 1, (2 * k + 1) * gx // 16), (37 * k + 5) % gy
                    owner = ox * world // gx if world > 1 and gx % world == 0 else 0
                    if owner != rank:
                        continue
                    n = lib.ogn_debug_column_weights(h.h, kind, ll, 0, ox, oy, None, 0)
                    if n <= 0:
                        raise RuntimeError(f"state_hash: column weights unavailable: {h.f.error()}")
                    buf = np.empty(n, dtype=np.float32)
                    lib.ogn_debug_column_weights(h.h, kind, ll, 0, ox, oy, buf.ctypes.data_as(C.POINTER(C.c_float)), n)
                    mine[(kind, ll, k)] = zlib.crc32(buf.tobytes())
    allw = [mine]
    if dist is not None and not ens:
        allw = [None] * world
        dist.all_gather_object(allw, mine)
    merged = {}
    for d in allw:
        merged.update(d)
    for key in sorted(merged):
        crc = zlib.crc32(np.array([key[0], key[1], key[2], merged[key]], dtype=np.int64).tobytes(), crc)
    return f"{crc:08x}", len(merged)


def parity_sample(og, cfg_sample, ref_trace, device):
    """The CPU sample the cpu_baseline leg just timed, replayed on the GPU at the SAME size from the same seed and inputs
    (std::mt19937 init in the reference's order): every step's prediction CSDR and the final hidden CSDR must be bit-equal."""
    opt = og.CreateOptions(device=device, deviceInit=False, replayRng=False)
    g = og.Hierarchy(cfg_sample["inputSizes"], cfg_sample["inputTypes"], cfg_sample["layerDescs"], seed=cfg_sample["seed"], options=opt)
    ok, first_bad = True, None
    for t, want in enumerate(ref_trace["pred"]):
        g.step(ref_trace["inputs"](t), True)
        if not np.array_equal(g.getPredictionCs(ref_trace["pred_input"]), want):
            ok, first_bad = False, t
            break
    if ok and not np.array_equal(g.scHiddenCs(0), ref_trace["hidden0"]):
        ok, first_bad = False, "final hidden CSDR"
    g.close()
    return {"config": cfg_sample["name"], "steps_compared": len(ref_trace["pred"]), "bit_identical": ok, "first_difference": first_bad,
            "what": "prediction CSDR after every step + final layer-0 hidden CSDR, GPU vs the CPU run that was just timed"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--workload", default="c4")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--stream", default="noisy", help="input stream kind for prediction inputs: noisy|coherent|iid")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--burnin", type=int, default=None,
                    help="c4: untimed training steps BEFORE the W warm-up steps so that the hierarchy is always measured in the "
                         "regime SURVEY.md §8d prescribes (200 steps of learning on the stream before the timed region); "
                         "default max(0, 200 - W)")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N > 1: CSDR slab exchange by the library's peer-memory kernel (default) or NCCL all-gather")
    ap.add_argument("--ensemble", type=int, default=512, help="c5 only: hierarchies per GPU (weak scaling; 512 = 80 GB)")
    ap.add_argument("--size", type=int, default=None, help="c4 only: grid edge (default 512); smaller sizes are for profiling runs")
    ap.add_argument("--cpu-size", type=int, default=None,
                    help="c4 only: grid edge of the CPU sample the reference is timed on (default 128; 224 is the largest size the "
                         "reference can construct and needs ~45 GiB of host memory and two minutes of init)")
    ap.add_argument("--cpu-members", type=int, default=8, help="c5 only: hierarchies the reference steps for the CPU sample")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = workload(args.workload, args.gpus, args.size)
    if args.cpu_size and args.workload == "c4":
        cfg["sample"] = configs.c4_large(args.cpu_size)
    if args.workload == "c5":
        cfg["sample_members"] = max(1, args.cpu_members)
    full_cols = cfg["layerDescs"][0].hiddenSize[0] * cfg["layerDescs"][0].hiddenSize[1]
    config = {"workload": cfg["desc"], "stream": args.stream, "weights": "fp32, random init", "learn": True}

    # ------------------------------------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        K = args.steps if args.steps is not None else 10
        W = args.warmup if args.warmup is not None else 3
        if args.workload == "c5":
            value, info = time_cpu_members(cfg["sample_members"], max(K, 20), W, budget_s=60.0)
        else:
            value, info = time_cpu(cfg["sample"], max(K, 40) if args.workload == "c4" else K, W, full_cols, budget_s=60.0)
        info["value"] = value
        info["unit"] = UNIT
        same = cfg["sample"]["name"] == cfg["name"]
        config["same_config"] = same
        config["extrapolated"] = not same
        if not same:  # the reference cannot construct the full size: say what actually ran
            config["workload"] = (f"{cfg['sample']['name']} sample of [{cfg['desc']}]: the reference ran {info['sample']}; value = "
                                  "sample steps/s x sample hidden columns / full hidden columns")
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": K,
                          "warmup": W, "ms_per_step": 1000.0 / value, "higher_is_better": True, "scaling": "strong",
                          "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config, "cpu_baseline": info,
                          "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    # ------------------------------------------------------------------------------------------------ B200 arm
    # defaults: the step counts SURVEY.md §8d prescribes for each config (C4: 200 warm-up + 300 timed)
    K = args.steps if args.steps is not None else {"c4": 300, "c5": 300}.get(args.workload, 2000)
    W = args.warmup if args.warmup is not None else {"c4": 200, "c5": 200}.get(args.workload, 300)
    W = max(W, 3)
    burn = args.burnin if args.burnin is not None else (max(0, 200 - W) if args.workload == "c4" else 0)
    import torch
    import ogmaneo2_b200 as og
    from ogmaneo2_b200 import build
    if build.stale():
        build.build()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if args.gpus != world:
        raise SystemExit(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; launch with torch.distributed.run")

    ens = args.workload == "c5"
    B = args.ensemble if ens else 1
    peer = world > 1 and args.exchange == "peer" and not ens
    if ens:  # members shard across GPUs with no exchange at all: rank r steps members [r*B, (r+1)*B)
        opt = og.CreateOptions(device=local_rank, deviceInit=True, replayRng=False,
                               ensembleSeeds=[1234 + rank * B + k for k in range(B)])
    else:
        opt = og.CreateOptions(device=local_rank, deviceInit=cfg["device_init"], replayRng=False, shardRank=rank, shardWorld=world,
                               peerExchange=peer)
    ext = {}

    def make_hierarchy(use_peer):
        if not ens:
            opt.peerExchange = use_peer
            opt.allGather = None
        if world > 1 and not use_peer and not ens:
            cache = {}

            def all_gather(buf, count, stream_ptr):
                key = (buf, count)
                if key not in cache:
                    full = torch.as_tensor(_DevArray(buf, count * world), device=f"cuda:{local_rank}")
                    cache[key] = (full, full[rank * count:(rank + 1) * count])
                full, mine = cache[key]
                with torch.cuda.stream(ext["s"]):
                    dist.all_gather_into_tensor(full, mine)
            opt.allGather = all_gather
        hh = og.Hierarchy(cfg["inputSizes"], cfg["inputTypes"], cfg["layerDescs"], seed=cfg["seed"], options=opt)
        ext["s"] = torch.cuda.ExternalStream(hh.stream(), device=local_rank)
        return hh

    h = make_hierarchy(peer)
    if peer:  # one CUDA IPC handle per rank, gathered over the process group; after this the step path never calls NCCL
        ok = 1
        try:
            handles = [None] * world
            dist.all_gather_object(handles, h.peerExport())
            h.peerAttach(handles)
        except Exception as e:  # CUDA IPC unavailable (e.g. no shared IPC namespace): fall back to NCCL on every rank
            print(f"bench.py: rank {rank}: peer-memory exchange unavailable ({e}); using NCCL all-gather", file=sys.stderr)
            ok = 0
        t = torch.tensor([ok], device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        if int(t.item()) == 0:
            h.close()
            del h
            peer = False
            h = make_hierarchy(False)
        dist.barrier()
    h.synchronize()

    # inputs: every prediction/none input gets the chosen stream; generated once, resident in HBM before timing
    kinds = [args.stream if k in ("noisy", "coherent", "iid") else k for k in cfg["streams"]]
    has_action = "action" in kinds
    n_in = len(cfg["inputSizes"])
    Ke = K
    total = burn + W + K + Ke  # the end-to-end leg gets its own, fresh window of the stream (re-feeding the timed window would
    # show the encoder inputs it has just adapted to: fewer weight rewrites, a faster step than the device-resident leg)
    if ens:
        gk = [rank * B + k for k in range(B)]
        host_in = [np.stack([np.concatenate([streams.make(kinds[0], t, cfg["inputSizes"][0], seed=7 + k, shift=17 * k) for k in gk])
                             for t in range(total)])]
    else:
        host_in = [np.stack([streams.make(k if k != "action" else "iid", t, s) for t in range(total)])
                   for s, k in zip(cfg["inputSizes"], kinds)]
    dev_in = [torch.from_numpy(a).to(f"cuda:{local_rank}") for a in host_in]
    torch.cuda.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def run_device(t0, n):
        for t in range(t0, t0 + n):
            h.stepDevice([d[t].data_ptr() for d in dev_in], True)

    run_device(0, burn)  # state preparation (the data-dependent SparseCoder write rate falls from 30 % to 3 % of the pairs)
    run_device(burn, W)
    barrier()
    for l in range(len(cfg["layerDescs"])):
        h.takeSCUpdateCount(l)
    # per-kernel CUDA events on every PROF_PERIOD-th step of the timed region (every step would cost ~8% of a sharded step):
    # at least 16 bracketed steps whatever K is
    PROF_PERIOD = max(1, K // 16)
    h.profileEnable(os.environ.get("OGN_BENCH_NOPROF") != "1", period=PROF_PERIOD)
    if peer:
        h.peerTakeWaitNs()
    clocks = ClockSampler(local_rank) if rank == 0 else None
    time.sleep(0.25 if rank == 0 else 0.0)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wall0 = time.time()
    e0.record(ext["s"])
    hq0 = time.perf_counter()
    run_device(burn + W, K)
    host_enqueue_ms = (time.perf_counter() - hq0) * 1e3 / K  # host time to queue one step (no waiting unless a queue fills up)
    e1.record(ext["s"])
    barrier()
    wall1 = time.time()
    ms = e0.elapsed_time(e1)
    clk = clocks.stop(wall0, wall1) if clocks else None
    prof_t = h.profileRead(timed=True)
    prof = {k: (v[0] * (v[1] / v[2]) if v[2] else 0.0, v[1]) for k, v in prof_t.items()}
    h.profileEnable(False)
    upd = [h.takeSCUpdateCount(l) for l in range(len(cfg["layerDescs"]))]
    ranks_info = None
    if dist is not None:
        t = torch.tensor([ms], device=f"cuda:{local_rank}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        mine = {"step_ms": ms / K, "kernel_ms_per_step": sum(v[0] for v in prof.values()) / K,
                "sc_pairs_rewritten_per_step": sum(upd) / K, "host_enqueue_ms_per_step": host_enqueue_ms,
                "kernel_ms": {k: round(v[0] / K, 4) for k, v in prof.items() if v[1]},
                "exchange_wait_us_per_step": (h.peerTakeWaitNs() / 1e3 / K) if peer else None}
        ms = float(t.item())
        ranks_info = [None] * world
        dist.all_gather_object(ranks_info, mine)
    value = K / (ms / 1000.0) * (B * world if ens else 1)  # c5: hierarchy steps/s summed over every member on every GPU

    # ------------------------------------------------------------------------------------------------ e2e (host API)
    barrier()
    pred_idx = [i for i, t in enumerate(cfg["inputTypes"]) if t != InputType.none]
    # streaming callers collect the prediction of step t after queueing step t + 1 (ogn_prediction_readback_queue/finish:
    # one step of pipelining, pinned slots on a copy stream); a shard reads back ITS slab of the prediction (1/N of the
    # result per rank, like the input rows it uploads). Hierarchies with an Actor or several prediction inputs use the
    # plain step() + getPredictionCs() loop.
    pipelined = len(pred_idx) == 1 and cfg["inputTypes"][pred_idx[0]] == InputType.prediction
    rx0 = rx1 = 0
    if pipelined and world > 1 and not ens:
        X = cfg["inputSizes"][pred_idx[0]][0]
        rx0, rx1 = rank * X // world, (rank + 1) * X // world
    if pipelined:
        pi = pred_idx[0]
        n_out = (cfg["inputSizes"][pi][0] if rx1 <= rx0 else rx1 - rx0) * cfg["inputSizes"][pi][1] * B
        out_buf = np.empty(n_out, dtype=np.int32)
        for _ in range(2):  # one-time set-up outside the timed region: copy stream, events and the two pinned slots
            h.predictionReadbackFinish(h.predictionReadbackQueue(pi, rx0, rx1), out_buf)
        barrier()
    w0 = time.perf_counter()
    if pipelined:
        ticket = None
        trace = [] if os.environ.get("OGN_BENCH_E2E_TRACE") == "1" else None
        for t in range(burn + W + K, burn + W + K + Ke):
            t_a = time.perf_counter()
            h.step([a[t] for a in host_in], True)
            t_b = time.perf_counter()
            nxt = h.predictionReadbackQueue(pi, rx0, rx1)
            t_c = time.perf_counter()
            if ticket is not None:
                h.predictionReadbackFinish(ticket, out_buf)
            ticket = nxt
            if trace is not None:
                trace.append((round((t_b - t_a) * 1e3, 3), round((t_c - t_b) * 1e3, 3), round((time.perf_counter() - t_c) * 1e3, 3)))
        h.predictionReadbackFinish(ticket, out_buf)
        if trace is not None and rank == 0:
            print("e2e trace (ms: step, queue, finish):", trace[:12], "...", trace[-4:], file=sys.stderr)
    else:
        for t in range(burn + W + K, burn + W + K + Ke):
            h.step([a[t] for a in host_in], True)
            for i in pred_idx:
                h.getPredictionCs(i)
    barrier()
    e2e_s = time.perf_counter() - w0
    if dist is not None:
        t = torch.tensor([e2e_s], device=f"cuda:{local_rank}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    h2d = sum(int(a[0].nbytes) for a in host_in)
    if world > 1 and not ens:  # a shard uploads only the input rows its slab reads (fields of its hidden rows + its Predictor rows)
        import ctypes as C
        plan = (C.c_int * 6)()
        ld0, s0 = cfg["layerDescs"][0], cfg["inputSizes"][0]
        h.f.shard_plan(ld0.hiddenSize[0], s0[0], ld0.ffRadius, rank, world, plan)
        x0, x1 = rank * s0[0] // world, (rank + 1) * s0[0] // world
        h2d = (max(plan[3], x1) - min(plan[2], x0)) * s0[1] * 4
    d2h = sum(cfg["inputSizes"][i][0] * cfg["inputSizes"][i][1] * 4 * B for i in pred_idx)
    if pipelined and rx1 > rx0:
        d2h = (rx1 - rx0) * cfg["inputSizes"][pred_idx[0]][1] * 4
    try:
        shash, scols = state_hash(h, cfg, pred_idx, rank, world, dist, ens)
    except Exception as e:  # never take the timing down with it
        shash, scols = f"unavailable: {e!r}", 0

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ------------------------------------------------------------------------------------------------ roofline
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    traffic, traffic_src = load_traffic()
    sites = bytes_per_launch(cfg)
    counts = tick_counts(cfg, burn + W + K)[burn + W:].sum(axis=0)  # updates per layer inside the timed region
    alg = {}
    for (l, kind, b, per_pair) in sites:
        a = alg.setdefault(kind, 0.0)
        # bytes one rank moves: x-slab sharding splits every launch; an ensemble multiplies it by the members per GPU
        alg[kind] = a + (b * counts[l] * B if ens else b * counts[l] / max(world, 1))
    # data-dependent SC writes (4 B x Vz per rewritten pair): the update kernel's bytes when learn ran as two kernels
    upd_bytes = sum(u * 4 * cfg["inputSizes"][0][2] if l == 0 else u * 4 * cfg["layerDescs"][l - 1].hiddenSize[2]
                    for l, u in enumerate(upd))
    if prof.get("sc_update", (0.0, 0))[1] > 0:
        alg["sc_update"] = upd_bytes
    elif "sc_learn" in alg:
        alg["sc_learn"] += upd_bytes
    if prof.get("hier_small", (0.0, 0))[1] > 0:  # the whole step ran as ONE kernel: its bytes are the step's bytes
        alg = {"hier_small": sum(alg.values())}
    kernels = {}
    for kind, (kms, n) in prof.items():
        if n == 0:
            continue
        ab = alg.get(kind, 0.0)
        kernels[kind] = {"launches": int(n), "timed_launches": int(prof_t[kind][2]), "ms_per_launch": kms / n,
                         "alg_bytes_per_launch": ab / n if ab else None,
                         "achieved_gbs": (ab / n) / (kms / n * 1e-3) / 1e9 if ab and kms > 0 else None,
                         "share_of_step": kms / ms}
    cand = [k for k in kernels if kernels[k]["achieved_gbs"]]
    dom = max(cand, key=lambda k: kernels[k]["ms_per_launch"] * kernels[k]["launches"]) if cand else None
    total_alg = sum(alg.values())
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["achieved_gbs"] if dom else None, "peak": peak, "unit": "GB/s",
                "frac": kernels[dom]["achieved_gbs"] / peak if dom else None,
                "traffic": traffic.get(dom) if args.workload == "c4" and world == 1 and not args.size else None,
                "traffic_source": (f"profiles/ncu_c4_512_traffic.csv <- {traffic_src.get(dom)} (ncu --set full, dram__bytes_read.sum + "
                                   "dram__bytes_write.sum per launch)") if traffic.get(dom) else None,
                "peak_source": peak_src,
                "step_achieved_gbs": total_alg / (ms * 1e-3) / 1e9, "step_frac": total_alg / (ms * 1e-3) / 1e9 / peak,
                "alg_bytes_per_step_per_gpu": total_alg / K, "kernels": kernels,
                "sc_pairs_rewritten_per_step": [u / K for u in upd]}

    cpu = None
    if not args.no_cpu and args.gpus == 1:
        try:
            if ens:
                v, info = time_cpu_members(cfg["sample_members"], 10, 1, budget_s=12.0)
                trace = None
            else:
                v, info, trace = time_cpu(cfg["sample"], 60 if args.workload == "c4" else 10, 1, full_cols, budget_s=12.0, want_trace=True)
            info["value"], info["unit"] = v, UNIT
            cpu = info
            if trace is not None and not ens:
                try:
                    cpu["gpu_parity_on_sample"] = parity_sample(og, cfg["sample"], trace, local_rank)
                except Exception as e:
                    cpu["gpu_parity_on_sample"] = {"bit_identical": None, "error": repr(e)}
        except Exception as e:  # the baseline must never take the GPU result down with it
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "unavailable", "sample": repr(e)}

    config.update({"parallelism": (f"{B} ensemble members per GPU x {world} GPU(s), no exchange" if ens else
                                   (f"x-slab sharding over {world} GPUs, CSDR slabs exchanged by "
                                    f"{'a peer-memory (NVLink) all-gather kernel' if peer else 'NCCL all-gather'}")
                                   if world > 1 else "single GPU"),
                   "l2": "per-step weight traffic (GBs) far exceeds the 126 MB L2; no flush needed" if args.workload in ("c4", "c5")
                   else "weights exceed L2" if args.workload == "c2" else "L2-resident (latency-bound config)",
                   "init": "device counter-hash" if cfg["device_init"] else "std::mt19937 in reference order",
                   "burn_in_steps": burn,
                   "kernel_timing": f"CUDA events around every kernel launch of every {PROF_PERIOD}th step of the timed region "
                                    "(timed_launches per kernel in roofline.kernels)"})
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": ms / K,
           "higher_is_better": True, "scaling": "weak" if ens else "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": config, "roofline": roofline, "cpu_baseline": cpu,
           "e2e": {"value": Ke / e2e_s * (B * world if ens else 1), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": Ke,
                   "per": "rank" if world > 1 else "job",
                   "api": ("Hierarchy.step(host CSDR) + prediction readback pipelined one step deep (ogn_prediction_readback_queue/finish)"
                           + ("; every rank uploads the input rows its shard reads and reads back its own slab of the prediction" if world > 1 and not ens else ""))
                   if pipelined else "Hierarchy.step(host CSDR) + getPredictionCs() every step"},
           "gpu_launches": int(sum(n for _, n in prof.values())), "clocks": clk,
           "state_hash": {"crc32": shash, "weight_columns": scols, "after_steps": burn + W + K + Ke,
                          "covers": "prediction CSDRs, SparseCoder hidden CSDRs, 8 SparseCoder + 8 Predictor weight columns per layer; "
                                    "independent of --gpus for the same --steps/--warmup"},
           "ranks": ranks_info, "host_enqueue_ms_per_step": host_enqueue_ms,
           "column_steps_per_s": value * sum(ld.hiddenSize[0] * ld.hiddenSize[1] * c for ld, c in zip(cfg["layerDescs"], counts)) / K}
    if args.workload == "c1" and world == 1:
        # a fresh hierarchy run for the contract's 10 000 steps; mispredicted steps among the last 1000
        # (19 for the reference on this stream; tests/test_oracle.py::test_c1_learns_the_sine_wave pins it). Untimed.
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from common import c1_last1000_errors
        fresh = og.Hierarchy(cfg["inputSizes"], cfg["inputTypes"], cfg["layerDescs"], seed=cfg["seed"])
        out["c1_last1000_prediction_errors"] = {"value": c1_last1000_errors(fresh, cfg), "reference": 19, "steps": 10000}
    print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
