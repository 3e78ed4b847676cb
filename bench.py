#!/usr/bin/env python
"""bench.py -- throughput of the MCTF analysis hot path on B200.

Metric (BASELINE.json): 1080p MCTF analysis frames/s (+ ME SAD Gops/s).
A "step" is one full temporal analysis (all TRLs-1 levels: split ->
motion_estimate -> decorrelate -> update) of one synthetic clip per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                  [--workload cfg3|cfg2|cfg1|cfg4]

N > 1: launched under torchrun, one rank per GPU.  The job is ONE clip of N times the workload's
GOPs, GOP-sharded (qsvc_b200/shard.py): every rank analyses its GOP range, the prediction tail of
pictures whose height is not a multiple of the block size is handed from shard to shard, and the
end-to-end leg gathers every rank's results into one set of sub-band files in the reference's layout
(weak scaling; NCCL carries the barrier, the max-over-ranks time and the point-to-point tail state).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: X, Y, GOPs, TRLs(reference flag), block, search range, subpixel, always_B
    "cfg1": dict(X=352, Y=288, GOPs=2, TRLs=5, bs=16, sr=4, a=0, always_B=0,
                 desc="CIF 352x288, 33 frames, GOP 16, block 16, search 4, integer-pel"),
    "cfg2": dict(X=704, Y=576, GOPs=4, TRLs=5, bs=16, sr=8, a=1, always_B=0,
                 desc="4CIF 704x576, 65 frames, GOP 16, block 16, search 8, half-pel"),
    "cfg3": dict(X=1920, Y=1080, GOPs=4, TRLs=6, bs=16, sr=16, a=2, always_B=1,
                 desc="1080p 1920x1080, 129 frames, GOP 32, block 16, search 16, quarter-pel"),
    "cfg4": dict(X=3840, Y=2160, GOPs=8, TRLs=6, bs=16, sr=16, a=0, always_B=1,
                 desc="2160p 3840x2160, 257 frames, GOP 32, block 16, search 16, integer-pel"),
}
SEEDS = {"cfg1": 1, "cfg2": 11, "cfg3": 2, "cfg4": 13}


def n_frames(w):
    return w["GOPs"] * 2 ** (w["TRLs"] - 1) + 1


def desp(x, l):
    for _ in range(l):
        x = (x + 1) // 2
    return x


def sad_ops_total(w):
    """18 * [sum_l bs^2 * blocks(l) + sum_subpel (bs<<l)^2 * blocks] per pair."""
    import math
    BY, BX = w["Y"] // w["bs"], w["X"] // w["bs"]
    pictures, sr, total = n_frames(w), w["sr"], 0.0
    for _ in range(1, w["TRLs"]):
        L = max(0, int(round(math.log2(sr))) - 1)
        per_pair = sum(w["bs"] ** 2 * desp(BY, l) * desp(BX, l) for l in range(L + 1))
        per_pair += sum((w["bs"] << l) ** 2 * BY * BX for l in range(1, w["a"] + 1))
        total += 18.0 * per_pair * (pictures // 2)
        pictures = (pictures + 1) // 2
        sr = min(2 * sr, 128)
    return total


def mc_bytes_total(w):
    """every input frame of a level read once, every output frame
    written once (high + low), plus the motion fields read and written."""
    fb = w["X"] * w["Y"] * 3 // 2
    BY, BX = w["Y"] // w["bs"], w["X"] // w["bs"]
    pictures, total = n_frames(w), 0
    for _ in range(1, w["TRLs"]):
        pairs = pictures // 2
        total += fb * (pictures + pairs + pairs + 1) + 16 * BY * BX * pairs
        pictures = (pictures + 1) // 2
    return total


# ------------------------------------------------------------------ clocks

class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.samples, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.QUERY}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), line.strip()))

    def stop(self, t_a=None, t_b=None):
        """Samples read between t_a and t_b (the timed region); when the region is shorter than
        the sampling period the samples of the whole loaded phase (warm-up + timed) are used."""
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [s for (t, s) in self.samples if t_a is None or (t_a <= t <= t_b + 0.15)]
        window = "timed region"
        if len(inside) < 3:
            inside, window = [s for (_, s) in self.samples], "warm-up + timed region"
        self.window = window
        for s in inside:
            f = [x.strip() for x in s.split(",")]
            try:
                sm.append(float(f[0]))
                mx = max(mx, float(f[1]))
                for name, v in zip(names, f[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


# --------------------------------------------------------- reference (CPU) arm

def cpu_reference_sample(w, procs, passes=1):
    """Times the reference's own CPU tools (oracle/_ref, unmodified) on a bounded sample of
    the workload and extrapolates to the whole clip (the full cfg3 analysis is ~20 minutes
    of one core).  The sample: `procs` concurrent chains, one per host core, spread round-
    robin over the workload's temporal levels; a chain is one frame pair of the workload's
    shape through split, motion_estimate, decorrelate and update with THAT level's search
    range, on frames 2^(t-1) apart of the synthetic pan (the motion a level-t pair really
    sees).  The reference is single-threaded and its cost is linear in the number of pairs
    of a level, so
        cpu_seconds(clip) = sum_t pairs_t * seconds_per_pair_t
    and frames/s = frames / cpu_seconds on one core ("as shipped"), frames / (cpu_seconds /
    procs) with one chain per core (our parallelisation of the reference; the per-pair
    seconds are measured with all `procs` cores busy, so contention is priced in)."""
    from oracle import run_ref
    from qsvc_b200 import yuv
    if not run_ref.build():
        return None
    X, Y = w["X"], w["Y"]
    levels, pictures, sr = [], n_frames(w), w["sr"]
    for t in range(1, w["TRLs"]):
        levels.append(dict(t=t, sr=sr, pairs=pictures // 2))
        pictures = (pictures + 1) // 2
        sr = min(2 * sr, 128)
    procs = max(procs, len(levels))
    span = 2 ** (w["TRLs"] - 1)
    clip = yuv.synthetic_clip(X, Y, span + 1, 99, max_shift=min(48, 3 * w["sr"]))
    tmp = tempfile.mkdtemp(prefix="qsvc_ref_")
    chains = []
    for p in range(procs):
        lv = levels[p % len(levels)]
        d = os.path.join(tmp, f"p{p}")
        os.makedirs(d)
        step = 2 ** (lv["t"] - 1)
        yuv.write_frames(os.path.join(d, "low_0"), clip[[0, step, 2 * step]])
        chains.append((d, lv))
    errs, secs = [], {lv["t"]: [] for lv in levels}

    def chain(d, lv):
        try:
            t0 = time.perf_counter()
            # one pair = 3 pictures; the chain's files are named as temporal_subband 1
            run_ref.analyze_step(d, 1, 3, X, Y, w["bs"], lv["sr"], w["a"], 0.0, w["always_B"])
            secs[lv["t"]].append(time.perf_counter() - t0)
        except Exception as e:  # noqa: BLE001
            errs.append(str(e))

    t0 = time.perf_counter()
    for _ in range(passes):
        th = [threading.Thread(target=chain, args=c) for c in chains]
        for t in th:
            t.start()
        for t in th:
            t.join()
    dt = time.perf_counter() - t0
    subprocess.call(["rm", "-rf", tmp])
    if errs:
        raise RuntimeError("reference chain failed: " + errs[0])
    per_level = {t: sum(v) / len(v) for t, v in secs.items()}
    cpu_seconds = sum(lv["pairs"] * per_level[lv["t"]] for lv in levels)
    pairs_total = sum(lv["pairs"] for lv in levels)
    return dict(value=n_frames(w) / (cpu_seconds / procs), value_1core=n_frames(w) / cpu_seconds,
                seconds=dt, procs=procs, extrapolated=True, cpu_seconds_clip=cpu_seconds,
                seconds_per_pair={f"level_{t}": round(v, 2) for t, v in per_level.items()},
                sample=f"{procs} concurrent single-pair chains of {X}x{Y} (a={w['a']}, bs={w['bs']}) spread over "
                       f"temporal levels 1..{len(levels)} (sr {levels[0]['sr']}..{levels[-1]['sr']}), "
                       f"{passes} pass(es), {dt:.1f} s wall; EXTRAPOLATED per level to the clip's "
                       f"{pairs_total} pairs = {cpu_seconds:.0f} core-seconds; reference as shipped "
                       f"(1 core) = {n_frames(w) / cpu_seconds:.3f} frames/s")


def _pictures_per_level(w):
    out, p = [], n_frames(w)
    for _ in range(1, w["TRLs"]):
        out.append(p)
        p = (p + 1) // 2
    return out


def bind_to_gpu_numa(index):
    """Pins this process to the CPUs of the NUMA node its GPU hangs off (and thereby, by first touch,
    its pinned host buffers to that node's memory): with several ranks per box the host<->device copies
    of the end-to-end path otherwise cross the socket interconnect.  Best effort; returns a note."""
    try:
        import torch
        props = torch.cuda.get_device_properties(index)
        bdf = None
        if hasattr(props, "pci_bus_id") and hasattr(props, "pci_device_id"):
            bdf = f"{getattr(props, 'pci_domain_id', 0):04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        if bdf is None:
            out = subprocess.run(["nvidia-smi", "-i", str(index), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                                 capture_output=True, text=True).stdout.strip()
            bdf = out[-12:].lower() if out else None
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return "numa: unknown node"
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return f"numa: node {node} has no allowed CPUs"
        os.sched_setaffinity(0, cpus)
        return f"numa: node {node}, {len(cpus)} CPUs"
    except Exception as e:  # noqa: BLE001
        return f"numa: not bound ({type(e).__name__})"


# ----------------------------------------------------------------------- main

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: weak = every GPU takes the workload's GOPs of an N times longer clip (default); "
                         "strong = the workload's own GOPs are divided among the GPUs (north_star config 4)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    wname = args.workload or "cfg3"
    w = WORKLOADS[wname]
    frames = n_frames(w)
    cfg = {"workload": f"{wname}: {w['desc']} (reference flags --TRLs={w['TRLs']} --GOPs={w['GOPs']}"
                       f" --update_factor=0 --always_B={w['always_B']})",
           "frames_per_gpu": frames, "sharding": "whole GOPs per GPU, no collective",
           "l2": f"inputs {frames * w['X'] * w['Y'] * 3 // 2 / 1e6:.0f} MB per step exceed the 126 MB L2"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        procs = min(os.cpu_count() or 1, 32)
        # a "step" of this arm is one bounded sample pass (~20-30 s of wall clock on every host
        # core); K timed steps would repeat identical CPU work, so at most two passes are run
        passes = 2 if args.steps >= 2 else 1
        r = cpu_reference_sample(w, procs, passes=passes)
        if r is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built"}))
            return 0
        # --steps/--warmup: the sample is one bounded pass; repeating it K times would
        # only repeat the same CPU work, so K is honoured as min(K, 1) passes.
        line = {"metric": "1080p MCTF analysis frames/s" if wname == "cfg3" else f"{wname} MCTF analysis frames/s",
                "value": r["value"], "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": frames / r["value"] * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16",
                "data": "synthetic", "config": cfg, "impl": "reference",
                "cpu_baseline": {"value": r["value"], "unit": "frames/s", "cores": r["procs"],
                                 "kind": "reference", "sample": r["sample"], "extrapolated": True,
                                 "value_1core": r["value_1core"], "sample_seconds": r["seconds"],
                                 "seconds_per_pair": r["seconds_per_pair"]},
                "extrapolated": True, "passes": passes,
                "e2e": {"value": r["value"], "unit": "frames/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import zlib

    import torch
    import torch.distributed as dist
    from qsvc_b200 import shard, yuv
    from qsvc_b200.mctf import Context, level_schedule

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (qsvc_b200 has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    numa_note = bind_to_gpu_numa(local_rank) if world > 1 else "numa: single rank, not bound"
    if world > 1:
        # NCCL's own log lines (version banner, NCCL_DEBUG=INFO) must not share stdout with the JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    X, Y, T, bs = w["X"], w["Y"], w["TRLs"], w["bs"]
    G = 2 ** (T - 1)
    strong = args.scaling == "strong" and world > 1
    # ONE clip, GOP-sharded: rank r takes GOPs [g0, g1), i.e. frames [g0*G, g1*G] inclusive.
    # weak (default): the job is world times the workload's GOPs, every rank takes the workload's count.  The
    #   long clip is the seeded base clip played forwards and backwards in turn, so that neighbouring shards
    #   agree on the frame they share and every rank can build its shard without the others'.
    # strong: the workload's own GOPs are divided among the ranks (cfg4: 8 GOPs over 2 / 4 / 8 GPUs).
    base = yuv.synthetic_clip(X, Y, frames, SEEDS[wname], max_shift=min(48, 3 * w["sr"]))
    if strong:
        if w["GOPs"] < world:
            raise SystemExit(f"bench.py: --scaling strong needs at least one GOP per GPU ({w['GOPs']} GOPs, {world} GPUs)")
        job_gops = w["GOPs"]
        ranges = shard.partition(job_gops, world)
        g0, g1 = ranges[rank]
        clip = base[g0 * G:g1 * G + 1]
    else:
        job_gops = world * w["GOPs"]
        ranges = shard.partition(job_gops, world)
        g0, g1 = ranges[rank]
        clip = base if rank % 2 == 0 else base[::-1]
    GOPs = g1 - g0                      # this rank's GOPs
    my_frames = GOPs * G + 1
    job_frames = n_frames(w) if strong else world * frames  # units the whole job processes
    pinned = torch.empty(clip.shape, dtype=torch.uint8).pin_memory()
    pinned.numpy()[...] = clip
    clip_pinned = pinned.numpy()
    cfg["frames_per_gpu"] = my_frames
    if world > 1:
        cfg["sharding"] = (f"one clip of {job_gops} GOPs ({job_gops * G + 1} frames), whole GOPs per GPU "
                           f"({'/'.join(str(b - a) for a, b in ranges)}), no data-path collective"
                           + (f"; the prediction tail of the {Y}-line picture is handed from shard to shard GPU to GPU "
                              "(NCCL point-to-point, once per level)" if Y % bs else ""))

    ctx = Context(local_rank)
    kw = dict(TRLs=T, block_size=bs, search_range=w["sr"], subpixel_accuracy=w["a"],
              update_factor=0.0, always_B=w["always_B"], block_size_min=bs)
    sched = level_schedule(GOPs, T, bs, w["sr"], bs)
    relay = shard.TailRelay(rank, ranges) if shard.needs_tail_exchange(Y, bs, world) else None
    if relay is not None:
        ctx.set_tail_exchange(relay, device=relay.device)
    first_global = g0 == 0

    # ---- device-resident throughput: inputs already in HBM when the timed region starts
    ctx.resident_load(clip_pinned, X, Y)
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        ctx.resident_analyze(first_global=first_global, **kw)
    barrier()
    l0 = ctx.launches
    t_a = time.time()
    ctx.timer_start()
    for _ in range(args.steps):
        ctx.resident_analyze(first_global=first_global, **kw)
    ms = ctx.timer_stop()
    barrier()
    clocks = sampler.stop(t_a, time.time())
    launches = ctx.launches - l0
    # checksums of the resident results: the end-to-end leg below must reproduce them
    crc_res = {}
    for s_ in sched:
        got = ctx.resident_fetch(s_["t"], s_["pairs"], s_["block_size"], want=("high", "motion"))
        crc_res[s_["t"]] = (zlib.crc32(got["high"]), zlib.crc32(got["motion"]))
    # per-class device times: separate, untimed steps (two events per launch perturb the step), with the
    # two compute lanes serialised so that every kernel is timed alone
    ctx.set_overlap(False)
    ctx.profile_enable(True)
    for _ in range(args.steps):
        ctx.resident_analyze(first_global=first_global, **kw)
    prof = ctx.profile_read()
    ctx.profile_enable(False)
    ctx.set_overlap(True)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    ms_per_step = ms_max / args.steps
    value = job_frames / (ms_per_step * 1e-3)

    # ---- end to end through the public API with host buffers (H2D + analysis + D2H + gather)
    fb = clip.shape[1]
    h2d = clip.nbytes
    # host-side gather (north_star): every rank's results land in ONE set of sub-band files laid out like
    # the reference's (high_t, motion_t, motion_filtered_t of all GOPs in order; low_{T-1}).  The files are
    # a shared mapping that every rank page-locks, and qsvc_analyze's device-to-host copies write the rank's
    # slices in place: the gather costs no extra pass over the data.
    gather_note, out_bufs, maps = "single GPU: results in pinned buffers of the context", None, {}
    if world > 1:
        shm_dir = f"/dev/shm/qsvc_bench_{os.environ.get('MASTER_PORT', '0')}"
        need = 0
        shapes = {}
        for s_ in sched:
            t_, b = s_["t"], s_["block_size"]
            n = job_gops * (G >> t_)  # pairs of the whole job at this level
            shapes[f"high_{t_}"] = ((n, fb), np.uint8)
            shapes[f"motion_{t_}"] = ((n, 4, Y // b, X // b), np.int16)
            shapes[f"motion_filtered_{t_}"] = ((n, 4, Y // b, X // b), np.int16)
        shapes[f"low_{T-1}"] = ((job_gops + 1, fb), np.uint8)
        need = sum(int(np.prod(sh)) * np.dtype(dt).itemsize for sh, dt in shapes.values())
        ok = torch.zeros(1, dtype=torch.int32, device="cuda")
        if rank == 0:
            try:
                os.makedirs(shm_dir, exist_ok=True)
                st = os.statvfs(shm_dir)
                if st.f_bavail * st.f_frsize > need + (64 << 20):
                    for k, (sh, dt) in shapes.items():
                        np.memmap(os.path.join(shm_dir, k), dtype=dt, mode="w+", shape=sh).flush()
                    ok[0] = 1
            except OSError:
                pass
        dist.broadcast(ok, 0)
        if int(ok.item()):
            for k, (sh, dt) in shapes.items():
                maps[k] = np.memmap(os.path.join(shm_dir, k), dtype=dt, mode="r+", shape=sh)
            out_bufs = {}
            for s_ in sched:
                t_, ppg = s_["t"], G >> s_["t"]  # pairs per GOP at this level
                for name in ("high", "motion", "motion_filtered"):
                    out_bufs[f"{name}_{t_}"] = maps[f"{name}_{t_}"][g0 * ppg:g1 * ppg]
            out_bufs[f"low_{T-1}"] = maps[f"low_{T-1}"][g0:g1 + 1]
            # first touch: every rank (already bound to its GPU's NUMA node) faults in ITS slices before anyone
            # page-locks the files, so a rank's device-to-host copies land in memory next to its GPU instead
            # of wherever the first registering rank put the whole file
            for k, v in out_bufs.items():
                if k != f"low_{T-1}" or rank == 0:
                    v[...] = 0
                else:
                    v[1:] = 0  # frame g0 belongs to the previous shard
            barrier()
            registered = True
            for k in shapes:
                registered = ctx.host_register(maps[k]) and registered
            gather_note = (f"rank slices written in place into shared sub-band files in {shm_dir} "
                           f"({need / 1e6:.0f} MB, reference file layout), "
                           + ("page-locked by every rank: the device-to-host copies are the gather"
                              if registered else "NOT page-locked (cudaHostRegister refused): pageable copies"))
        else:
            gather_note = f"no shared memory for the gathered files ({need / 1e6:.0f} MB needed): results stay per rank"

    outs = {}

    def e2e_step():
        # the public API call: host clip in, host sub-bands out (pinned buffers reused)
        outs.update(ctx.analyze(clip_pinned, X, Y, GOPs, reuse_buffers=True, first_global=first_global,
                                out=out_bufs, **kw))

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    d2h = int(ctx.last_d2h_bytes)
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = job_frames * args.steps / float(t.item())
    e2e_ok = all((zlib.crc32(outs[f"high_{t_}"]), zlib.crc32(outs[f"motion_{t_}"])) == crc_res[t_] for t_ in crc_res)
    tt = torch.tensor([1 if e2e_ok else 0], dtype=torch.int32, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MIN)
    e2e_ok = bool(int(tt.item()))

    # sharded == whole (N = 2: the 2-shard clip once more on rank 0 alone, untimed, against the gathered files)
    sharded_equals_whole = None
    if world == 2 and maps:
        barrier()
        if rank == 0:
            ctx.set_tail_exchange(None)
            long_clip = base if strong else np.concatenate([base, base[::-1][1:]], axis=0)
            whole = ctx.analyze(long_clip, X, Y, job_gops, first_global=True, **kw)
            sharded_equals_whole = all(np.array_equal(whole[k], maps[k]) for k in maps)
            del whole, long_clip
        barrier()
    if relay is not None:
        ctx.set_tail_exchange(None)

    if rank != 0:
        for m_ in maps.values():
            ctx.host_unregister(m_)
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- rooflines
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6650 GB/s"
    u8_peak, i32_peak = ctx.int_peak()
    steps = args.steps
    cls_ms = {k: v[0] / steps for k, v in prof.items()}
    cls_n = {k: v[1] // steps for k, v in prof.items()}
    share_of_workload = GOPs / w["GOPs"]  # strong scaling: rank 0 holds a part of the workload's GOPs
    sad = sad_ops_total(w) * share_of_workload
    mc_total = mc_bytes_total(w) * share_of_workload
    search_ms = cls_ms["search"] + cls_ms.get("search_exact", 0.0)
    # everything motion estimation launches: pyramid DWT, byte planes and their interpolations, searches
    # (the image class also holds decorrelate's reference up-sampling: an upper bound for ME)
    me_all_ms = search_ms + cls_ms["dwt_rows"] + cls_ms["dwt_cols"] + cls_ms["image"]
    dominant = max(cls_ms, key=cls_ms.get)
    rooflines = {
        "me_search": {"bound": "int-sad", "achieved": sad / (search_ms * 1e-3) / 1e9 if search_ms else None,
                      "peak": u8_peak / 1e9, "peak_i32": i32_peak / 1e9, "unit": "G SAD-op/s",
                      "frac": (sad / (search_ms * 1e-3)) / u8_peak if search_ms else None,
                      "frac_all_me_kernels": (sad / (me_all_ms * 1e-3)) / u8_peak if me_all_ms else None,
                      "frac_whole_step": (sad / (ms_per_step * 1e-3)) / u8_peak,
                      "note": "peak = measured VABSDIFF4.U8.ACC issue rate x4 (qsvc_int_peak: >95% of the loop's "
                              "instructions are SADs, profiles/r2_int_peak_sass.txt; 64 thread-instr/clk/SM, "
                              "profiles/r2_pipe_probe.txt); algorithmic SAD-ops of SURVEY 8(d); frac = against the "
                              "time in the search kernels, frac_all_me_kernels = against search + pyramid DWT + "
                              "plane preparation, frac_whole_step = against the whole analysis step"},
        "mc_path": {"bound": "hbm", "achieved": mc_total / (ms_per_step * 1e-3) / 1e9,
                    "peak": hbm_peak, "unit": "GB/s",
                    "frac": mc_total / (ms_per_step * 1e-3) / 1e9 / hbm_peak,
                    "note": "algorithmic frame+motion bytes of all levels / whole step time"},
    }
    # the dominant kernel class of the step, against the bound that applies to it
    share = cls_ms[dominant] / max(1e-9, sum(cls_ms.values()))
    pairs_total = sum(s_["pairs"] for s_ in sched)
    fbytes = X * Y * 3 // 2
    field_bytes = 8 * (Y // bs) * (X // bs)
    # algorithmic bytes of decorrelate per pair = two reference frames and the odd
    # frame read, the high frame written, one motion field read and written (12.4 MB at 1080p)
    mc_pair_bytes = 4 * fbytes + 2 * field_bytes
    kernel_of = {"residue": "k_mc_march", "image": "k_upsample_chain/k_upsample2x (+ plane loads, border fills)",
                 "predict": "k_predict_u8/k_tail_state", "dwt_rows": "k_dwt_rows", "dwt_cols": "k_dwt_cols"}
    # dram__bytes_read+write per step of the class's kernels, from the committed ncu capture
    # (profiles/r*_ncu.json, written by profiles/summarize.py; cfg3 only)
    traffic = {}
    for tag in ("r2", "r1"):
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", f"{tag}_ncu.json"))).get("dram_bytes_per_step", {})
            break
        except (OSError, ValueError):
            pass
    if dominant in ("search", "search_exact"):
        roof = {"bound": "int-sad", "kernel": "k_subpel_tma/k_subpel_strip/k_subpel_exact/k_search16",
                "achieved": rooflines["me_search"]["achieved"], "peak": u8_peak / 1e9, "unit": "G SAD-op/s",
                "frac": rooflines["me_search"]["frac"], "traffic": None, "share_of_step": share}
    else:
        a_bytes = mc_pair_bytes * pairs_total if dominant in ("residue", "predict") else mc_total
        ach = a_bytes / (cls_ms[dominant] * 1e-3) / 1e9
        tr = traffic.get(dominant) if wname == "cfg3" else None
        roof = {"bound": "hbm", "kernel": kernel_of.get(dominant, dominant), "achieved": ach, "peak": hbm_peak,
                "unit": "GB/s", "frac": ach / hbm_peak, "traffic": tr,
                "algorithmic_bytes": a_bytes, "launches": cls_n[dominant], "ms": cls_ms[dominant],
                "share_of_step": share, "peak_source": hbm_src,
                "note": "achieved = SURVEY 8(d) algorithmic bytes of the class's launches of one step / their "
                        "summed device time (CUDA events on the library's stream); traffic = ncu dram bytes "
                        "of the same launches.  k_mc_march reads the up-sampled reference planes "
                        "(16x the frame bytes) and is bound by instruction issue, not by HBM: see "
                        "profiles/r2_summary.md"}

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        try:
            r = cpu_reference_sample(w, min(os.cpu_count() or 1, 32))
            if r:
                cpu = {"value": r["value"], "unit": "frames/s", "cores": r["procs"],
                       "kind": "reference", "sample": r["sample"], "extrapolated": True,
                       "value_1core": r["value_1core"], "sample_seconds": r["seconds"],
                       "seconds_per_pair": r["seconds_per_pair"]}
        except Exception as e:  # noqa: BLE001
            cpu = {"value": None, "unit": "frames/s", "cores": 0, "kind": "reference",
                   "sample": f"failed: {e}"}

    extras = None
    if world == 1 and not args.no_extras:
        try:
            extras = measure_extras(ctx, w, wname, clip_pinned, outs, kw)
        except Exception as e:  # noqa: BLE001
            extras = {"failed": f"{type(e).__name__}: {e}"}

    line = {
        "metric": "1080p MCTF analysis frames/s" if wname == "cfg3" else f"{wname} MCTF analysis frames/s",
        "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "int16", "data": "synthetic",
        "config": dict(cfg, host=numa_note), "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "matches_resident_run": e2e_ok, "gather": gather_note,
                "sharded_equals_whole": sharded_equals_whole,
                "note": "update_factor 0: low_t of the lower levels are frames of the input clip and are not "
                        "copied back (views of the caller's array); high_t, both motion fields, frame types "
                        "and low_{T-1} are"},
        "gpu_launches": int(launches),
        "roofline": roof, "rooflines": rooflines,
        "me_sad_gops": sad / (ms_per_step * 1e-3) / 1e9,
        "kernel_ms_per_step": cls_ms, "kernel_launches_per_step": cls_n,
        "kernel_ms_note": "per-class device time of separate, untimed steps with the two compute lanes serialised "
                          "(every kernel alone); the timed steps overlap motion estimation of level t+1 with the "
                          "decorrelate of level t, so the classes sum to more than ms_per_step",
        "cpu_baseline": cpu,
        "extras": extras,
    }
    print(json.dumps(line))
    for m_ in maps.values():
        ctx.host_unregister(m_)
    if world > 1:
        try:
            import shutil
            shutil.rmtree(f"/dev/shm/qsvc_bench_{os.environ.get('MASTER_PORT', '0')}", ignore_errors=True)
        except OSError:
            pass
        dist.destroy_process_group()
    return 0


def measure_extras(ctx, w, wname, clip_pinned, outs, kw):
    """What production runs besides the headline (not part of `value`): the workload with the codec's
    default --update_factor=0.25 (compress.py:101), its synthesis (north_star config 5: the decode path),
    and the cfg2 analyze + synthesize round trip (config 2).  A few steps each."""
    import zlib
    from qsvc_b200 import yuv
    X, Y, GOPs, T, bs = w["X"], w["Y"], w["GOPs"], w["TRLs"], w["bs"]
    frames = n_frames(w)
    ex = {}
    # (a) update_factor 0.25, device-resident
    kw25 = dict(kw, update_factor=0.25)
    ctx.resident_load(clip_pinned, X, Y)
    ctx.resident_analyze(**kw25)
    ctx.timer_start()
    for _ in range(2):
        ctx.resident_analyze(**kw25)
    ms = ctx.timer_stop() / 2
    ctx.set_overlap(False)  # per-class times: every kernel alone on one stream
    ctx.profile_enable(True)
    ctx.resident_analyze(**kw25)
    prof = ctx.profile_read()
    ctx.profile_enable(False)
    ctx.set_overlap(True)
    ex["update_factor_0.25"] = {"ms_per_step": ms, "frames_per_s": frames / (ms * 1e-3),
                                "kernel_ms": {k: v[0] for k, v in prof.items()},
                                "dominant": max(prof, key=lambda k: prof[k][0]),
                                "note": "device-resident analysis; the levels run in order (level t+1 needs low_t), "
                                        "motion estimation beside the decorrelate's plane preparation; kernel_ms "
                                        "from a separate step with one stream"}
    # (b) synthesis of the headline analysis (update_factor 0), host sub-bands in, host frames out
    sub = {f"low_{T-1}": outs[f"low_{T-1}"]}
    for t in range(1, T):
        sub[f"high_{t}"], sub[f"motion_{t}"] = outs[f"high_{t}"], outs[f"motion_filtered_{t}"]
        sub[f"frame_types_{t}"] = outs[f"frame_types_{t}"]
    rec = ctx.host_alloc(clip_pinned.shape)
    skw = dict(block_size=bs, search_range=w["sr"], subpixel_accuracy=w["a"], update_factor=0.0)
    ctx.synthesize(sub, X, Y, GOPs, T, out=rec, **skw)
    t0 = time.perf_counter()
    dev_ms = 0.0
    for _ in range(2):
        ctx.synthesize(sub, X, Y, GOPs, T, out=rec, **skw)
        dev_ms += ctx.resident_stats()["total_ms"]
    e2e_s = (time.perf_counter() - t0) / 2
    G = 2 ** (T - 1)
    ex["synthesis"] = {"e2e_frames_per_s": frames / e2e_s, "device_ms_per_step": dev_ms / 2,
                       "device_frames_per_s": frames / (dev_ms / 2 * 1e-3),
                       "even_frames_of_top_level_restored": bool(np.array_equal(rec[0::G], clip_pinned[0::G])),
                       "crc32_low_0": zlib.crc32(rec),
                       "note": "inverse MCTF (un_update, correlate, merge per level) of the sub-bands the e2e leg "
                               "produced; e2e = pushes + synthesis + download of low_0"}
    # (c) cfg2 round trip
    if wname != "cfg2":
        w2 = WORKLOADS["cfg2"]
        c2 = yuv.synthetic_clip(w2["X"], w2["Y"], n_frames(w2), SEEDS["cfg2"], max_shift=min(48, 3 * w2["sr"]))
        k2 = dict(TRLs=w2["TRLs"], block_size=w2["bs"], search_range=w2["sr"], subpixel_accuracy=w2["a"],
                  update_factor=0.0, always_B=w2["always_B"], block_size_min=w2["bs"])
        p2 = ctx.host_alloc(c2.shape)
        p2[...] = c2
        r2 = ctx.host_alloc(c2.shape)

        def trip():
            o = ctx.analyze(p2, w2["X"], w2["Y"], w2["GOPs"], reuse_buffers=True, **k2)
            sb = {f"low_{w2['TRLs']-1}": o[f"low_{w2['TRLs']-1}"]}
            for t in range(1, w2["TRLs"]):
                sb[f"high_{t}"], sb[f"motion_{t}"] = o[f"high_{t}"], o[f"motion_filtered_{t}"]
                sb[f"frame_types_{t}"] = o[f"frame_types_{t}"]
            ta = time.perf_counter()
            ctx.synthesize(sb, w2["X"], w2["Y"], w2["GOPs"], w2["TRLs"], block_size=w2["bs"],
                           search_range=w2["sr"], subpixel_accuracy=w2["a"], update_factor=0.0, out=r2)
            return time.perf_counter() - ta

        trip()
        t0 = time.perf_counter()
        syn = sum(trip() for _ in range(3))
        tot = time.perf_counter() - t0
        f2 = n_frames(w2)
        ex["cfg2_round_trip"] = {"analyze_frames_per_s": 3 * f2 / (tot - syn), "synthesize_frames_per_s": 3 * f2 / syn,
                                 "round_trip_frames_per_s": 3 * f2 / tot,
                                 "psnr_db": float(10 * np.log10(255.0 ** 2 / max(1e-9, ((r2.astype(np.float64) - c2) ** 2).mean()))),
                                 "note": "704x576, 65 frames, GOP 16, half-pel: analyze + synthesize through host buffers"}
    return ex


if __name__ == "__main__":
    sys.exit(main())
