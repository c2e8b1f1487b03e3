#!/usr/bin/env python
"""bench.py — radar frames/s polar->pose on B200 (BASELINE.json metric) for the fused front end.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]; configs[2] with --mds 1): a seeded synthetic Oxford-shaped
sequence (400 azimuths x 3768 range bins uint8 + 11 metadata bytes, 0.0438 m/bin) of
--frames scans per GPU.  One step = one pass of the hot path over that batch: scan decode +
polar->Cartesian + u8 pyramid for every frame, then pyramidal LK, distance-consistency clique
rejection, Kabsch (and the motion-distortion solve) for every consecutive pair -> one pose per
pair.  `value` = poses/s with the scans resident in HBM; `e2e` = the same through the C ABI with
host buffers (pinned H2D of every scan + D2H of the poses inside the timed region).
Multi-GPU: every rank owns an independent sequence (weak scaling, no data-path collective);
the poses are gathered to rank 0 over NCCL once per step.

--impl reference times the reference's own CPU path (the third-party calls the reference
makes: cv2.warpPolar, cv2.calcOpticalFlowPyrLK, scipy cdist, networkx find_cliques, numpy SVD,
scipy least_squares — oracle/ref_pipeline.py) on the host cores for the same workload.
"""
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

L2_BYTES = 126e6


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200, help="timed steps (the pipeline drains inside the timed region: ~6 ms of clique tail after the last step)")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=256, help="scans per GPU per step")
    ap.add_argument("--features", type=int, default=200)
    ap.add_argument("--res", type=float, default=0.0438, help="range resolution m/bin (BASELINE: 0.0438; reference: 0.0432)")
    ap.add_argument("--mds", type=int, default=0, help="1 = motion-distortion solve enabled (configs[2])")
    ap.add_argument("--write-f32", type=int, default=0, help="also materialise the f32 Cartesian image per frame")
    ap.add_argument("--cpu-pairs", type=int, default=0, help="pairs in the bounded CPU sample (0 = 8 x cores: about 15-20 s of CPU work)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--batches", type=int, default=5, help="batches in flight on one handle (pipeline depth; the clique + MDS tail of a batch lasts about three steps)")
    ap.add_argument("--no-numa", action="store_true", help="do not bind ranks to their GPU's NUMA node (multi-GPU runs)")
    ap.add_argument("--no-gather", action="store_true", help="diagnostic: skip the per-step NCCL pose gather")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------
def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index, disabled=False, n_gpus=1):
        self.index = index
        self.disabled = disabled
        self.n_gpus = n_gpus
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        if self.disabled:
            return
        self._start()

    def _start(self):
        """index = None samples every GPU of the box from ONE nvidia-smi process (multi-GPU runs: rank 0 only, so that
        eight samplers do not poll the driver while eight ranks are being timed)."""
        try:
            sel = [] if self.index is None else ["-i", str(self.index)]
            self.p = subprocess.Popen(["nvidia-smi"] + sel + [f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                                               "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable" if not self.disabled else "sampled by rank 0"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.count(",") >= 8]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                if self.index is None and int(r[0]) >= self.n_gpus:
                    continue                      # a GPU of the box this job does not use
                sm.append(float(r[1])); mx.append(float(r[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower() == "active":
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def workload(args, rank):
    """Seeded synthetic sequence for this rank: raw scans, pairs, features, ground-truth poses."""
    from radarslampy_b200 import synthetic as S
    rb = int(87.5 / args.res)
    world = S.World(seed=1234 + rank)
    # configs[2] (--mds 1): the same sequence rendered with intra-scan motion distortion
    raw, poses = S.make_sequence(args.frames, res_m=args.res, world=world, first=0, distort=bool(args.mds))
    kmax = max(64, (args.features + 63) // 64 * 64)
    pair_idx, feats, counts = S.sequence_pairs(args.frames, world, poses, args.res, rb, k=args.features, max_features=kmax)
    return rb, kmax, raw, poses, pair_idx, feats, counts


def stage_bytes(cfgd, S_frames, P, K_total, write_f32, levels, fused):
    """ALGORITHMIC bytes per launch group: every input read once, every mandated output
    written once."""
    A, W, n = cfgd["azimuths"], cfgd["range_bins"], cfgd["n"]
    lv = [(n, n)]
    for _ in range(1, levels):
        lv.append(((lv[-1][0] + 1) // 2, (lv[-1][1] + 1) // 2))
    px = [a * b for a, b in lv]
    b = {}
    if fused:
        # the interleave is a layout pass (read the used bins, write them back 4 frames per word)
        b["polar2cart"] = S_frames * 2 * A * W
        # scan -> level 0 + level 1: read the used polar bins once, write both levels once
        b["scan_to_l0l1"] = S_frames * (A * W + px[0] + px[1])
        b["pyr_down"] = S_frames * sum(px[l - 1] + px[l] for l in range(2, levels))
    else:
        b["polar2cart"] = S_frames * (A * W + px[0] + (4 * px[0] if write_f32 else 0))
        b["scan_to_l0l1"] = 0
        b["pyr_down"] = S_frames * sum(px[l - 1] + px[l] for l in range(1, levels))
    # one track: per level an 18x18 previous-image patch + at least one 16x16 next-image window, + the err pass
    b["klt"] = K_total * (levels * (324 + 256) + 256) + K_total * (8 + 8 + 1 + 4)
    b["compact"] = K_total * (8 + 8 + 1) + K_total * (8 + 8 + 4)
    b["reject"] = K_total * 16 + K_total
    b["kabsch"] = K_total * 17 + P * 48
    b["mds"] = K_total * 17 + P * (48 + 24 + 48)
    b["finish"] = K_total * 6 + P * 120
    return b


# ---------------------------------------------------------------------------------------
def cpu_pairs_worker(job):
    """One contiguous chunk of pairs through the reference's library calls (frames converted once)."""
    import cv2
    cv2.setNumThreads(1)
    from oracle import ref_pipeline as P
    raw, feats, counts, poses, res_m, with_mds = job
    carts = [P.polar_to_cart(P.extract_polar(r, res_m)) for r in raw]
    out = []
    for p in range(len(raw) - 1):
        o = P.track_pair(None, None, feats[p, :counts[p]], prev_pose=poses[p], with_mds=with_mds, range_res_m=res_m,
                         carts=(carts[p], carts[p + 1]))
        out.append((o["h"], o["R"], o["n_inliers"]))
    return out


def cpu_reference_rate(args, raw, feats, counts, poses, n_pairs, repeats=1, pool=None):
    """poses/s of the CPU path over `n_pairs` consecutive pairs, all host cores (one process per core,
    OpenCV single-threaded inside each).  `pool`: a warmed multiprocessing pool to reuse across calls."""
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    n_pairs = max(1, min(n_pairs, len(raw) - 1))
    workers = min(cores, n_pairs)
    per = -(-n_pairs // workers)
    jobs = []
    for w in range(workers):
        a, bnd = w * per, min(n_pairs, (w + 1) * per)
        if a >= bnd:
            break
        jobs.append((raw[a:bnd + 1], feats[a:bnd], counts[a:bnd], poses[a:bnd], args.res, bool(args.mds)))
    best = None

    def timed(pl):
        nonlocal best
        for _ in range(repeats):
            t0 = time.perf_counter()
            pl.map(cpu_pairs_worker, jobs, chunksize=1)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)

    if pool is not None:
        timed(pool)
    else:
        with mp.get_context("fork").Pool(len(jobs)) as pl:
            pl.map(cpu_pairs_worker, [(j[0][:2], j[1][:1], j[2][:1], j[3][:1], j[4], j[5]) for j in jobs])   # warm imports
            timed(pl)
    return n_pairs / best, len(jobs), n_pairs, best


def warmed_pool(args, raw, feats, counts, poses):
    """One worker process per host core with the libraries imported (kept across the steps of --impl reference)."""
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    pool = mp.get_context("fork").Pool(cores)
    job = (raw[:2], feats[:1], counts[:1], poses[:1], args.res, bool(args.mds))
    pool.map(cpu_pairs_worker, [job] * cores, chunksize=1)
    return pool


def run_reference(args, rank, world):
    if rank != 0:
        return
    rb, kmax, raw, poses, pair_idx, feats, counts = workload(args, 0)
    cores = os.cpu_count() or 1
    n_pairs = args.cpu_pairs
    pool = warmed_pool(args, raw, feats, counts, poses)
    if not n_pairs:
        # bounded sample per step, sized so that warmup + steps finish in about two minutes whatever K is
        rate0, _, _, _ = cpu_reference_rate(args, raw, feats, counts, poses, min(args.frames - 1, 2 * cores), pool=pool)
        per_step_s = 120.0 / max(1, args.warmup + args.steps)
        n_pairs = int(min(args.frames - 1, max(cores, rate0 * per_step_s)))
        n_pairs -= n_pairs % min(cores, n_pairs)            # equal chunks for the worker processes
    times = []
    for i in range(args.warmup + args.steps):
        rate, workers, n_used, dt = cpu_reference_rate(args, raw, feats, counts, poses, n_pairs, pool=pool)
        if i >= args.warmup:
            times.append(dt)
    pool.close()
    pool.join()
    t = float(np.mean(times))
    value = n_used / t
    sample = (f"{n_used} consecutive pairs of the {args.frames}-frame synthetic sequence per step, split over {workers} worker "
              f"processes (cv2 single-threaded in each), frames converted once per worker")
    line = {
        "impl": "reference", "metric": "radar frames/sec polar->pose", "value": value, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8/f32/f64", "data": "synthetic",
        "config": workload_config(args, rb, kmax),
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": workers, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(args, rb, kmax):
    return {"workload": f"synthetic Oxford-shaped sequence, {args.frames} scans/GPU/step (400x3768 u8 + 11 metadata bytes, "
                        f"{args.res} m/bin -> {rb} used bins, {2 * (rb // 2)}^2 Cartesian), {args.features} features/pair given, "
                        f"KLT 15x15 x 4 levels, clique rejection, Kabsch" + (", scans rendered with intra-scan motion distortion, motion-distortion LM" if args.mds else ""),
            "frames_per_step_per_gpu": args.frames, "pairs_per_step_per_gpu": args.frames - 1, "features_per_pair": args.features,
            "range_res_m": args.res, "mds": bool(args.mds), "write_cart_f32": bool(args.write_f32),
            "l2": "inputs larger than L2: each step streams >= 1.3 GB of scans + pyramids per GPU (L2 = 126 MB)"}


# ---------------------------------------------------------------------------------------
def bind_to_gpu_numa(index: int):
    """Pin this process (and so the pinned staging buffers it allocates next: first touch) to the host cores of the
    NUMA node GPU `index` hangs off.  With several ranks on one box every rank otherwise stages through whichever
    socket the launcher left it on, and the H2D streams of the far-socket GPUs cross the inter-socket link."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.split(":", 1)
        dev = f"{int(dom, 16):04x}:{rest.lower()}"
        cpus = set()
        for part in open(f"/sys/bus/pci/devices/{dev}/local_cpulist").read().strip().split(","):
            if part:
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        node = int(open(f"/sys/bus/pci/devices/{dev}/numa_node").read().strip())
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return {"node": node, "cpus": len(cpus), "pci": dev}
    except Exception as e:      # no NVML / sysfs (containers): leave the affinity alone
        return {"node": None, "cpus": None, "error": type(e).__name__}


def emit(line: dict):
    """The ONE JSON line goes to the process's original stdout (see main)."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def ncu_traffic(kernel: str, frames: int):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu
    `--set full` capture (profiles/ncu_traffic.json, written by tools/ncu_summary.py --traffic); scaled when the
    capture used a different frame count (traffic is linear in frames).  None if there is no capture."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None
    rec = json.load(open(p)).get(kernel)
    if not rec:
        return None
    return float(rec["dram_bytes_per_launch"]) * frames / float(rec["frames"])


def main():
    global _REAL_STDOUT
    # libraries (NCCL's version banner, torch warnings) write to fd 1: keep the JSON line alone on stdout by
    # pointing fd 1 at stderr for the rest of the process and writing the result to a saved copy
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    numa = bind_to_gpu_numa(local_rank) if (world > 1 and not args.no_numa) else {"node": None, "cpus": None}
    rb, kmax, raw_np, poses, pair_idx, feats, counts = workload(args, rank)
    cpu_line = None
    if not args.no_cpu_baseline and rank == 0 and world == 1:
        # timed BEFORE the CUDA context exists (fork-safe), on a bounded sample of the same workload
        cores = os.cpu_count() or 1
        n_pairs = args.cpu_pairs or min(args.frames - 1, 8 * cores)
        rate, workers, n_used, dt = cpu_reference_rate(args, raw_np, feats, counts, poses, n_pairs, repeats=2)
        cpu_line = {"value": rate, "unit": "frames/s", "cores": workers, "kind": "port",
                    "sample": f"first {n_used} pairs of the same sequence through oracle/ref_pipeline.py (the reference's own "
                              f"cv2/scipy/networkx/numpy calls), {workers} processes, best of 2 ({dt:.2f} s)"}

    import torch
    import torch.distributed as dist
    from radarslampy_b200 import _ffi, _shard

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - this framework has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    S, P = args.frames, args.frames - 1
    cfg = _ffi.default_config()
    cfg.range_bins = rb
    cfg.cart_res_m = 2 * args.res
    cfg.dist_thr_px = 0.5 / (2 * args.res)
    cfg.max_frames, cfg.max_pairs, cfg.max_features = S, max(P, 1), kmax
    cfg.write_cart_f32 = args.write_f32
    stream = torch.cuda.Stream()
    fe = _ffi.RadarFE(cfg, device=local_rank, stream=stream.cuda_stream)
    # NB batches used in rotation: the upload of one overlaps the image + KLT kernels of the previous one and the
    # latency-bound clique search of the one before that.  With two, an upload has to wait for the
    # tail of the same batch's previous run; three keep all three streams busy.
    NB = max(1, args.batches)
    batches = [fe.new_batch() for _ in range(NB)]
    # pinned host staging (what a caller streaming scans from disk would fill); everything that is uploaded
    # asynchronously must be page-locked or the copy call blocks the host
    def pin(a, dt):
        out = _ffi.pinned_empty(a.shape, dt); out[...] = a
        return out
    raw = pin(raw_np, np.uint8)
    feats_p = pin(feats, np.float32)
    pair_idx_p = pin(np.asarray(pair_idx, np.int32), np.int32)
    counts_p = pin(np.asarray(counts, np.int32), np.int32)
    prev_pose = pin(poses[:-1], np.float64)
    outs = [b.alloc_outputs(pinned=True) for b in batches]
    gatherer = _shard.PoseGatherer(world * P, world, rank, device="cuda") if world > 1 else None
    with_mds = bool(args.mds)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        fe.sync()

    def gather_poses(res_host):
        """trajectory concatenation: poses of every rank to rank 0 over NCCL"""
        if world > 1 and not args.no_gather:
            gatherer.gather(_shard.pack_records(res_host))

    def upload(b):
        b.upload(raw, pair_idx_p, feats_p, counts_p, prev_pose=prev_pose, sync=False)

    # ---- device-resident arm: scans already in HBM ------------------------------------------
    for b in batches:
        upload(b)
        b.set_profiling(True)
    fe.sync()
    for i in range(max(args.warmup, NB)):
        batches[i % NB].run_async(with_mds=with_mds)
    fe.sync()
    for b in batches:
        b.stage_times()
    # one GPU: sample it; several: rank 0 samples all of them (the GPUs the job's ranks run on)
    sampler = ClockSampler(local_rank if world == 1 else None, disabled=(world > 1 and rank != 0), n_gpus=world)
    barrier()
    sampler.start()
    n0 = fe.launch_count()
    fe.timer_start()
    for i in range(args.steps):
        batches[i % NB].run_async(with_mds=with_mds)
    ms = fe.timer_stop_ms()          # joins the tail streams: every step's poses are complete
    launches = fe.launch_count() - n0
    barrier()
    stage_ms, runs = {}, 0
    for b in batches:
        sm, r = b.stage_times()
        runs += r
        for k, v in sm.items():
            stage_ms[k] = stage_ms.get(k, 0.0) + v
    res, nxt, corr = batches[0].download(outs[0])
    fe.sync()
    gather_poses(res)
    barrier()

    # ---- end-to-end arm: host buffers through the C ABI, copies inside the timed region ------
    for b in batches:
        b.set_profiling(False)

    def e2e_loop(n):
        # NB steps in flight: a batch's poses are read (and gathered) right before that batch is re-used
        for i in range(n):
            k = i % NB
            b = batches[k]
            if i >= NB:
                b.wait()
                gather_poses(outs[k][0][:P])
            upload(b)                                   # H2D of every scan of the step (copy stream)
            b.run_async(with_mds=with_mds)
            b.download(outs[k], sync=False, want_tracks=False)   # D2H of the poses (tail stream)
        for i in range(max(0, n - NB), n):
            batches[i % NB].wait()
            gather_poses(outs[i % NB][0][:P])

    e2e_loop(max(args.warmup, NB))
    barrier()
    t0 = time.perf_counter()
    fe.timer_start()
    e2e_loop(args.steps)
    ms_e2e = fe.timer_stop_ms()
    barrier()
    ms_e2e = max(ms_e2e, 0.0)
    if world > 1:                                 # the NCCL gather runs on torch's stream: bound it by the host clock
        ms_e2e = max(ms_e2e, (time.perf_counter() - t0) * 1e3)
    clocks = sampler.stop()

    # ---- reduce over ranks (max time) ---------------------------------------------------------
    if world > 1:
        t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    K_total = int(counts.sum())
    cfgd = {"azimuths": cfg.azimuths, "range_bins": rb, "n": fe.n}
    sb = stage_bytes(cfgd, S, P, K_total, bool(args.write_f32), 4, fused=not args.write_f32)
    peak, peak_src = peaks()
    stages = {}
    for name, tot in stage_ms.items():
        per = tot / max(runs, 1)
        stages[name] = {"ms": per, "alg_bytes": sb[name], "gbs": (sb[name] / (per * 1e-3) / 1e9) if per > 0 else None}
    # dominant kernel = the HBM-streaming stage with the largest share of the step (the clique search is a
    # latency-bound combinatorial kernel with ~3 KB of traffic per pair; it has no bandwidth roofline and is
    # reported in `stages` and `clique`)
    streaming = [k for k in ("polar2cart", "scan_to_l0l1", "pyr_down", "klt") if stages[k]["ms"] > 0]
    dom = max(streaming, key=lambda k: stages[k]["ms"])
    achieved = stages[dom]["gbs"] or 0.0
    lv = [fe.n]
    for _ in range(3):
        lv.append((lv[-1] + 1) // 2)
    survey_bytes = S * (cfg.azimuths * rb + 4 * fe.n * fe.n + sum(v * v for v in lv))      # per launch of the conversion
    value = world * P * args.steps / (ms * 1e-3)
    e2e = world * P * args.steps / (ms_e2e * 1e-3)
    h2d = S * cfg.azimuths * (cfg.meta_bytes + rb) + P * (8 + kmax * 8 + 4 + 24)
    d2h = P * 120
    line = {
        "metric": "radar frames/sec polar->pose", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8/f32/f64", "data": "synthetic", "config": dict(workload_config(args, rb, kmax), pipelining=f"{NB} batches alternate on one handle (copy / image+KLT / rejection+solve streams)"),
        "klt_tracks_per_s": world * K_total * args.steps / (ms * 1e-3),
        "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": int(h2d) * world, "d2h_bytes_per_step": int(d2h) * world,
                "ms_per_step": ms_e2e / args.steps, "note": "bytes are the whole job's (all ranks) per step"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": ncu_traffic(dom, S), "peak_source": peak_src,
                     "alg_bytes_per_launch": sb[dom], "ms_per_launch": stages[dom]["ms"],
                     # SURVEY.md §8(d) counts the reference's materialised f32 Cartesian image (4 n^2 per frame) plus the
                     # whole u8 pyramid as mandated outputs of the conversion; this kernel never writes the f32 image and
                     # produces levels 0-1 only, so `achieved` / `frac` above use the smaller, actual figure
                     "note": "dominant = largest main-stream (step-time-determining) kernel; k_clique has longer launches "
                             "(stages.reject) but is one warp per pair on ~2 % of the machine, latency-bound with ~3 KB of traffic "
                             "per pair, and runs on per-batch tail streams underneath the image kernels of the following batches",
                     "survey_8d_bytes_per_launch": survey_bytes, "survey_8d_frac": (survey_bytes / (stages[dom]["ms"] * 1e-3) / 1e9 / peak)
                     if dom == "scan_to_l0l1" and stages[dom]["ms"] > 0 else None},
        "stages": stages,
        "clocks": clocks,
        "numa": numa,
        "pose_check": {"median_dtheta_rad": float(np.median(np.arctan2(res["R"][:, 2], res["R"][:, 0]))),
                       "expected_dtheta_rad": 0.025, "median_inliers": float(np.median(res["n_inliers"])),
                       "worklimit_pairs": int((res["status"] != 0).sum())},
        "clique": {"nodes_median": float(np.median(res["clique_nodes"])), "nodes_p90": float(np.percentile(res["clique_nodes"], 90)),
                   "nodes_max": int(res["clique_nodes"].max()), "good_median": float(np.median(res["n_good"]))},
    }
    if cpu_line is not None:
        line["cpu_baseline"] = cpu_line
    emit(line)
    for b in batches:
        b.close()
    fe.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
