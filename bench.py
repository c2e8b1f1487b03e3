#!/usr/bin/env python
"""bench.py — radar frames/s polar->pose on B200 (BASELINE.json metric) for the fused front end.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline workload (BASELINE.json configs[1] data; configs[2] with --mds 1): a seeded synthetic Oxford-shaped
sequence (400 azimuths x 3768 range bins uint8 + 11 metadata bytes, 0.0438 m/bin) of
--frames scans per GPU.  One pass = the hot path over that device batch: scan decode +
polar->Cartesian + u8 pyramid for every frame, then pyramidal LK, distance-consistency clique
rejection, Kabsch (and the motion-distortion solve) for every consecutive pair -> one pose per
pair; one STEP = --passes-per-step (4) such passes, pipelined (1024 scans, 1020 poses per GPU), so that the
K steps the driver asks for are long against the ~6 ms drain of the last passes' clique tail, which is inside the
timed region; FEATURES ARE HANDED IN and the pairs are INDEPENDENT (operationally configs[3]'s pair batches).
`value` = poses/s with the scans resident in HBM; `e2e` = the same through the C ABI with
host buffers (pinned H2D of every scan + D2H of the poses inside the timed region).
Multi-GPU: every rank owns an independent sequence (weak scaling, no data-path collective);
the poses are gathered to rank 0 over NCCL once per step.

Further legs inside the same JSON line (--legs, default all):
  chained     the REAL chained odometry (RawROAMSystem.run): 256 sequences in lock step on the device (rf_seq):
              features carried over, on-device re-detection + SSC, keyframes, MDS, one CUDA graph per step;
              frames/s resident and e2e, single-sequence latency, and the CPU loop (oracle/ref_system.py) beside it
  strong      configs[3]: 4096 independent pairs from 16 sequences (seeds 1000..1015) in contiguous blocks over the
              ranks (_shard.block_range), poses gathered over NCCL — strong scaling
  mds         configs[2]: motion-distorted scans + the motion-distortion solve (value and e2e)
  stress      configs[4]: dense 2000^2 grid, 10 k SSC features, 4-level KLT -> tracks/s
  detect_ssc / fmt / peaks   on-device detection + SSC bisection, FMT rotation prior, polar peak extraction
  clique_real rejectOutliers on the real data/tiny graphs and the reference's outlier_test.npz, one at a time
  parity_sample   GPU poses / inlier counts of the first pairs against what the cpu_baseline leg just computed

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
    ap.add_argument("--steps", type=int, default=50, help="timed steps (the pipeline drains inside the timed region: ~6 ms of clique tail after the last step)")
    ap.add_argument("--passes-per-step", type=int, default=4,
                    help="device batches (passes of --frames scans through the pipeline) that make up one step: a step is "
                         "passes x frames scans per GPU.  One 256-scan pass lasts ~2 ms, so 20 one-pass steps would be a 41 ms "
                         "timed region of which the ~6 ms drain of the last passes' clique tail is 13 %; four passes per step "
                         "make the same 20 steps 164 ms")
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
    ap.add_argument("--legs", default="all", help="comma list of extra legs: chained,strong,mds,stress,detect_ssc,fmt,peaks,clique_real (all | none)")
    ap.add_argument("--pairs", type=int, default=4096, help="strong leg: total independent pairs (BASELINE configs[3])")
    ap.add_argument("--chain-seq", type=int, default=256, help="chained leg: sequences per GPU in lock step")
    ap.add_argument("--chain-steps", type=int, default=12)
    ap.add_argument("--chain-seq-large", type=int, default=512,
                    help="chained leg: a second resident measurement with this many sequences per GPU (0 = off): the 256-sequence "
                         "step is bound by per-runner launch latencies, twice the sequences show the machine's throughput")
    ap.add_argument("--gather-every", type=int, default=8, help="multi-GPU: steps whose pose records travel in one NCCL gather")
    ap.add_argument("--chain-runners", type=int, default=16, help="chained leg: rf_seq runners (streams) the sequences are split over")
    return ap.parse_args()


ALL_LEGS = ("chained", "strong", "mds", "stress", "detect_ssc", "fmt", "peaks", "clique_real")
CHAIN_WARMUP = 3


def wanted_legs(args):
    if args.legs in ("none", ""):
        return ()
    if args.legs == "all":
        return ALL_LEGS
    return tuple(l for l in args.legs.split(",") if l in ALL_LEGS)


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


def workload(args, rank, extra_frames=0, distort=None, seed=None, workers=None):
    """Seeded synthetic sequence for this rank: raw scans, pairs, features, ground-truth poses.  Rendered on all host
    cores (identical to the serial generator: every frame has its own seeds).  `extra_frames` more scans are rendered
    past --frames for the chained leg (sequence s = frames s, s + 1, ...)."""
    from radarslampy_b200 import synthetic as S
    rb = int(87.5 / args.res)
    seed = 1234 + rank if seed is None else seed
    world = S.World(seed=seed)
    distort = bool(args.mds) if distort is None else distort
    # configs[2] (--mds 1): the same sequence rendered with intra-scan motion distortion
    raw, poses = S.make_sequence_parallel(args.frames + extra_frames, res_m=args.res, seed=seed, first=0, distort=distort, workers=workers)
    kmax = max(64, (args.features + 63) // 64 * 64)
    pair_idx, feats, counts = S.sequence_pairs(args.frames, world, poses[:args.frames], args.res, rb, k=args.features, max_features=kmax)
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
    outputs = []

    def timed(pl):
        nonlocal best
        for _ in range(repeats):
            t0 = time.perf_counter()
            res = pl.map(cpu_pairs_worker, jobs, chunksize=1)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
            outputs[:] = [o for chunk in res for o in chunk]     # (h [2,1] m, R [2,2], n_inliers) per pair, in pair order

    if pool is not None:
        timed(pool)
    else:
        with mp.get_context("fork").Pool(len(jobs)) as pl:
            pl.map(cpu_pairs_worker, [(j[0][:2], j[1][:1], j[2][:1], j[3][:1], j[4], j[5]) for j in jobs])   # warm imports
            timed(pl)
    cpu_reference_rate.last_outputs = outputs
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
        per_step_s = 180.0 / max(1, args.warmup + args.steps)
        n_pairs = int(min(args.frames - 1, max(cores, rate0 * per_step_s)))
        if n_pairs < args.frames - 1:                       # a bounded sample: equal chunks for the worker processes
            n_pairs -= n_pairs % min(cores, n_pairs)        # (else: every pair of the step, exactly the GPU arm's work)
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
    M = max(1, args.passes_per_step)
    return {"workload": f"BASELINE configs[1] data as independent pair batches: synthetic Oxford-shaped sequence, {M * args.frames} scans/GPU/step "
                        f"as {M} pipelined device batches of {args.frames} "
                        f"(400x3768 u8 + 11 metadata bytes, {args.res} m/bin -> {rb} used bins, {2 * (rb // 2)}^2 Cartesian), "
                        f"{args.features} features/pair HANDED IN (ground-truth scatterers), consecutive pairs treated as INDEPENDENT, "
                        f"KLT 15x15 x 4 levels, clique rejection, Kabsch" + (", scans rendered with intra-scan motion distortion, motion-distortion LM" if args.mds else "")
                        + "; the chained odometry with on-device detection is the `chained` leg",
            "frames_per_step_per_gpu": M * args.frames, "pairs_per_step_per_gpu": M * (args.frames - 1), "passes_per_step": M,
            "frames_per_pass": args.frames, "features_per_pair": args.features,
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

    t_setup0 = time.perf_counter()
    numa = bind_to_gpu_numa(local_rank) if (world > 1 and not args.no_numa) else {"node": None, "cpus": None}
    legs = wanted_legs(args)
    gen_workers = max(1, (os.cpu_count() or 1) // world)
    chain_need = (max(args.chain_seq, args.chain_seq_large) + CHAIN_WARMUP + args.chain_steps + 2) if "chained" in legs else 0
    # headline drive (configs[1]; rendered with motion distortion under --mds 1), and the motion-distorted drive of the
    # same world / poses that the `chained` and `mds` legs read (configs[2]); one render when they coincide
    rb, kmax, raw_long, poses_long, pair_idx, feats, counts = workload(args, rank, extra_frames=max(0, chain_need - args.frames) if args.mds else 0,
                                                                       workers=gen_workers)
    raw_np, poses = raw_long[:args.frames], poses_long[:args.frames]
    pre = {}
    if args.mds:
        dist_long = raw_long
    else:
        n_dist = max(chain_need, args.frames if ("mds" in legs and world == 1 and args.frames >= 4) else 0)
        dist_long = None
        if n_dist:
            from radarslampy_b200 import synthetic as SY
            dist_long, _ = SY.make_sequence_parallel(n_dist, res_m=args.res, seed=1234 + rank, first=0, distort=True, workers=gen_workers)
            if "mds" in legs and world == 1 and args.frames >= 4:
                pre["mds"] = (rb, kmax, dist_long[:args.frames], poses, pair_idx, feats, counts)
    if "strong" in legs:
        pre["strong"] = strong_data(args, rank, world, rb, kmax, gen_workers)
    cpu_line, cpu_out = None, None
    if not args.no_cpu_baseline and rank == 0 and world == 1:
        # timed BEFORE the CUDA context exists (fork-safe), on a bounded sample of the same workload
        cores = os.cpu_count() or 1
        n_pairs = args.cpu_pairs or min(args.frames - 1, 8 * cores)
        rate, workers, n_used, dt = cpu_reference_rate(args, raw_np, feats, counts, poses, n_pairs, repeats=2)
        cpu_out = list(cpu_reference_rate.last_outputs)
        cpu_line = {"value": rate, "unit": "frames/s", "cores": workers, "kind": "port",
                    "sample": f"first {n_used} pairs of the same sequence through oracle/ref_pipeline.py (the reference's own "
                              f"cv2/scipy/networkx/numpy calls), {workers} processes, best of 2 ({dt:.2f} s)"}
        if "chained" in legs:
            pre["chained_cpu"] = chained_cpu_baseline(args, dist_long, cores)
    setup_s = time.perf_counter() - t_setup0

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - this framework has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = dict(args=args, rank=rank, local_rank=local_rank, world=world, rb=rb, kmax=kmax, torch=torch, dist=dist)
    line = weak_leg(ctx, raw_np, poses, pair_idx, feats, counts, cpu_line, cpu_out, numa, setup_s)
    extra_legs = {}
    for name in legs:
        t0 = time.perf_counter()
        try:
            if name == "chained":
                out = leg_chained(ctx, dist_long, pre.get("chained_cpu"))
            elif name == "strong":
                out = leg_strong(ctx, pre["strong"])
            elif world > 1 or rank != 0:
                continue                       # the remaining legs are single-GPU properties: measured at N = 1 only
            elif name == "mds":
                out = leg_mds(ctx, pre.get("mds"))
            elif name == "stress":
                out = leg_stress(ctx)
            elif name == "detect_ssc":
                out = leg_detect_ssc(ctx, raw_np)
            elif name == "fmt":
                out = leg_fmt(ctx, raw_np)
            elif name == "peaks":
                out = leg_peaks(ctx, raw_np)
            elif name == "clique_real":
                out = leg_clique_real(ctx)
        except Exception as e:                 # a failing leg must not take the headline down with it
            out = {"error": f"{type(e).__name__}: {e}"}
        if out is not None:
            out["leg_s"] = time.perf_counter() - t0
            extra_legs[name] = out
    if rank == 0:
        line.update(extra_legs)
        line["legs_s"] = {k: v.get("leg_s") for k, v in extra_legs.items()}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def weak_leg(ctx, raw_np, poses, pair_idx, feats, counts, cpu_line, cpu_out, numa, setup_s):
    """The headline measurement (docstring of this file): resident `value`, `e2e`, roofline of the dominant kernel."""
    args, rank, local_rank, world, rb, kmax = (ctx[k] for k in ("args", "rank", "local_rank", "world", "rb", "kmax"))
    torch, dist = ctx["torch"], ctx["dist"]
    from radarslampy_b200 import _ffi, _shard

    S, P = args.frames, args.frames - 1
    M = max(1, args.passes_per_step)          # passes (device batches of S scans) per step
    n_pass = args.steps * M                    # passes inside the timed region
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
    # trajectory concatenation over NCCL: the records of --gather-every steps travel in one dist.gather, queued without
    # blocking the host (_shard.PoseGatherer.gather_async).  A blocking gather per step staged its records through a
    # pageable H2D copy that waits behind every queued scan upload, and cost a quarter of the 2-GPU e2e rate (99 k vs
    # 129 k frames/s without any gather; box ceiling 131 k).
    G_EVERY = max(1, args.gather_every)
    gatherer = _shard.PoseGatherer(world * P * G_EVERY, world, rank, device="cuda") if world > 1 else None
    gather_buf = np.zeros((P * G_EVERY, _shard.RECORD_WIDTH), np.float64)
    gather_fill = [0]
    with_mds = bool(args.mds)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        fe.sync()

    def gather_poses(res_host, flush=False):
        """trajectory concatenation: poses of every rank to rank 0 over NCCL, G_EVERY steps per message"""
        if world > 1 and not args.no_gather:
            if res_host is not None:
                j = gather_fill[0]
                gather_buf[j * P:(j + 1) * P] = _shard.pack_records(res_host)
                gather_fill[0] = j + 1
            if gather_fill[0] == G_EVERY or (flush and gather_fill[0] > 0):
                gatherer.gather_async(gather_buf)       # queued; the host goes on feeding uploads
                gather_fill[0] = 0
            if flush:
                gatherer.result()                       # every pose has reached rank 0

    def upload(b):
        b.upload(raw, pair_idx_p, feats_p, counts_p, prev_pose=prev_pose, sync=False)

    # ---- device-resident arm: scans already in HBM ------------------------------------------
    for b in batches:
        upload(b)
        b.set_profiling(True)
    fe.sync()
    for i in range(max(args.warmup * M, NB)):
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
    for i in range(n_pass):
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
    gather_poses(res, flush=True)
    barrier()
    # the same kernels timed ALONE: one batch at a time with a full sync between passes, so nothing of another batch
    # (clique / solve tail, uploads) shares the SMs -- the in-step figures above carry that interference
    alone_ms, alone_runs = {}, 0
    for _ in range(3):
        batches[0].run_async(with_mds=with_mds)
        fe.sync()
    batches[0].stage_times()
    for _ in range(8):
        batches[0].run_async(with_mds=with_mds)
        fe.sync()
    sm_a, alone_runs = batches[0].stage_times()
    alone_ms = {k: v / max(alone_runs, 1) for k, v in sm_a.items()}
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
        gather_poses(None, flush=True)                  # the last partial message: every pose has reached rank 0

    e2e_loop(max(args.warmup * M, NB))
    barrier()
    t0 = time.perf_counter()
    fe.timer_start()
    e2e_loop(n_pass)
    ms_e2e = fe.timer_stop_ms()
    barrier()
    ms_e2e = max(ms_e2e, 0.0)
    if world > 1:                                 # the NCCL gather runs on torch's stream: bound it by the host clock
        ms_e2e = max(ms_e2e, (time.perf_counter() - t0) * 1e3)
    clocks = sampler.stop()
    # H2D ceiling of the e2e arm: the same uploads (2-D used-column copies from the same pinned buffers, all ranks at once)
    # with nothing else queued -- what the box's PCIe / host-memory system gives this job
    n_up = max(8, min(n_pass, 32))
    for i in range(NB):
        upload(batches[i])
    barrier()
    t0 = time.perf_counter()
    for i in range(n_up):
        upload(batches[i % NB])
    fe.sync()
    ms_h2d = (time.perf_counter() - t0) * 1e3 / n_up          # per pass
    barrier()

    # ---- reduce over ranks (max time) ---------------------------------------------------------
    if world > 1:
        t = torch.tensor([ms, ms_e2e, ms_h2d], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e, ms_h2d = float(t[0]), float(t[1]), float(t[2])
    res = {k: np.array(res[k]) for k in res.dtype.names}      # detach from the pinned output buffers
    for b in batches:
        b.close()
    fe_n = fe.n
    fe.close()
    if rank != 0:
        return None

    K_total = int(counts.sum())
    cfgd = {"azimuths": cfg.azimuths, "range_bins": rb, "n": fe_n}
    sb = stage_bytes(cfgd, S, P, K_total, bool(args.write_f32), 4, fused=not args.write_f32)
    peak, peak_src = peaks()
    stages = {}
    for name, tot in stage_ms.items():
        per = tot / max(runs, 1)
        stages[name] = {"ms": per, "alg_bytes": sb[name], "gbs": (sb[name] / (per * 1e-3) / 1e9) if per > 0 else None}
        am = alone_ms.get(name, 0.0)
        if am > 0:
            stages[name]["ms_alone"] = am
            stages[name]["gbs_alone"] = sb[name] / (am * 1e-3) / 1e9
    # dominant kernel = the HBM-streaming stage with the largest share of the step (the clique search is a
    # latency-bound combinatorial kernel with ~3 KB of traffic per pair; it has no bandwidth roofline and is
    # reported in `stages` and `clique`)
    streaming = [k for k in ("polar2cart", "scan_to_l0l1", "pyr_down", "klt") if stages[k]["ms"] > 0]
    dom = max(streaming, key=lambda k: stages[k]["ms"])
    achieved = stages[dom]["gbs"] or 0.0
    lv = [fe_n]
    for _ in range(3):
        lv.append((lv[-1] + 1) // 2)
    survey_bytes = S * (cfg.azimuths * rb + 4 * fe_n * fe_n + sum(v * v for v in lv))      # per launch of the conversion
    value = world * P * n_pass / (ms * 1e-3)
    e2e = world * P * n_pass / (ms_e2e * 1e-3)
    h2d = S * cfg.azimuths * (cfg.meta_bytes + rb) + P * (8 + kmax * 8 + 4 + 24)      # per pass
    d2h = P * 120
    line = {
        "metric": "radar frames/sec polar->pose", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8/f32/f64", "data": "synthetic", "config": workload_config(args, rb, kmax),
        "pipelining": f"{NB} batches alternate on one handle (copy / image+KLT / rejection+solve streams)",
        "passes_per_step": M, "ms_per_pass": ms / n_pass,
        "klt_tracks_per_s": world * K_total * n_pass / (ms * 1e-3),
        "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": int(h2d) * world * M, "d2h_bytes_per_step": int(d2h) * world * M,
                "ms_per_step": ms_e2e / args.steps, "ms_per_pass": ms_e2e / n_pass,
                "note": "bytes are the whole job's (all ranks) per step",
                # the uploads alone (same pinned buffers, same 2-D copies, every rank at once): the H2D ceiling of this job on this box
                "h2d_only": {"ms_per_step": ms_h2d * M, "gbs": int(h2d) * world / (ms_h2d * 1e-3) / 1e9,
                             "frames_per_s_ceiling": world * P / (ms_h2d * 1e-3),
                             "e2e_fraction_of_ceiling": (e2e / (world * P / (ms_h2d * 1e-3))) if ms_h2d > 0 else None}},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": ncu_traffic(dom, S), "peak_source": peak_src,
                     "alg_bytes_per_launch": sb[dom], "ms_per_launch": stages[dom]["ms"],
                     # `achieved` / `frac` are in-step (other batches' clique / solve kernels and uploads share the SMs: the
                     # sustained figure); the same launch with the device to itself:
                     "alone": {"ms_per_launch": stages[dom].get("ms_alone"), "achieved": stages[dom].get("gbs_alone"),
                               "frac": (stages[dom]["gbs_alone"] / peak) if stages[dom].get("gbs_alone") else None},
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
    if cpu_out:
        line["parity_sample"] = parity_sample(res, cpu_out, with_mds, [0.5 * fe_n * cfg.cart_res_m] * 2)
    line["setup_s"] = setup_s
    return line


def parity_sample(res, cpu_out, with_mds, centre_m):
    """GPU batch results of the first pairs against what the cpu_baseline leg computed for the same pairs
    (oracle/ref_pipeline.track_pair: the reference's own cv2 / scipy / networkx / numpy calls).

    Tracker.getTransform's h is the translation of the IMAGE CORNER frame (pixel coordinates scaled to metres, Tracker.py:108-127),
    so a rotation difference d_theta shows up in it multiplied by the lever arm to the sensor (|centre| ~ 124 m here:
    1e-6 rad <-> 1.2e-4 m).  The pose the reference integrates is the sensor's (coordinates centred on RADAR_CART_CENTER,
    RawROAMSystem.py:153,199,266): t_centre = h + (R - I) c.  Both differences are reported; the north_star tolerance
    (1e-4 m / 1e-5 rad) is applied to the sensor pose."""
    n = min(len(cpu_out), len(res["h"]))
    dm, dc, drad, same = 0.0, 0.0, 0.0, 0
    c = np.asarray(centre_m, np.float64)
    for p in range(n):
        h, R, n_in = cpu_out[p]
        h = np.asarray(h, np.float64).ravel()
        Rg = np.asarray(res["R"][p], np.float64).reshape(2, 2)
        hg = np.asarray(res["h"][p], np.float64)
        dm = max(dm, float(np.abs(hg - h).max()))
        dc = max(dc, float(np.abs((hg + (Rg - np.eye(2)) @ c) - (h + (np.asarray(R, np.float64) - np.eye(2)) @ c)).max()))
        d = np.arctan2(Rg[1, 0], Rg[0, 0]) - np.arctan2(R[1, 0], R[0, 0])
        drad = max(drad, float(abs((d + np.pi) % (2 * np.pi) - np.pi)))
        same += int(res["n_inliers"][p] == n_in)
    return {"pairs": n, "max_abs_dt_sensor_m": dc, "max_abs_dtheta_rad": drad, "inlier_count_equal": same,
            "max_abs_dh_corner_frame_m": dm, "lever_arm_m": float(np.hypot(*c)),
            "tolerance": "1e-4 m / 1e-5 rad (north_star) on the sensor pose t = h + (R - I) c, c = image centre; h itself is the "
                         "corner-frame translation and carries d_theta x lever_arm",
            "pass": bool(dc <= 1e-4 and drad <= 1e-5 and same == n)}


# ---------------------------------------------------------------------------------------
# further legs (see the module docstring)
def _fe_config(args, rb, max_frames, max_pairs, max_features):
    from radarslampy_b200 import _ffi
    cfg = _ffi.default_config()
    cfg.range_bins = rb
    cfg.cart_res_m = 2 * args.res
    cfg.dist_thr_px = 0.5 / (2 * args.res)
    cfg.max_frames, cfg.max_pairs, cfg.max_features = max_frames, max(max_pairs, 1), max_features
    cfg.write_cart_f32 = args.write_f32      # 0: the fused image path of the headline (the library default materialises the f32 image)
    return cfg


def _pin(a, dt=None):
    from radarslampy_b200 import _ffi
    a = np.asarray(a)
    out = _ffi.pinned_empty(a.shape, dt or a.dtype)
    out[...] = a
    return out


def _reduce_max(ctx, *vals):
    if ctx["world"] == 1:
        return vals
    torch, dist = ctx["torch"], ctx["dist"]
    t = torch.tensor(list(vals), dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return tuple(float(v) for v in t)


def _barrier(ctx, fe):
    if ctx["world"] > 1:
        ctx["dist"].barrier()
    ctx["torch"].cuda.synchronize()
    fe.sync()


def measure_pair_chunks(ctx, chunks, with_mds, steps, warmup, max_frames, max_pairs, gather_total=None, nb=None):
    """Pair batches through rf_batch: `chunks` = [(raw [F,A,W] u8, pair_idx, feats, counts, prev_pose)], one pass over all
    of them per step.  Returns (ms resident per step, ms e2e per step, results of every chunk of the last pass)."""
    from radarslampy_b200 import _ffi, _shard
    args, world, rank = ctx["args"], ctx["world"], ctx["rank"]
    torch = ctx["torch"]
    cfg = _fe_config(args, ctx["rb"], max_frames, max_pairs, ctx["kmax"])
    stream = torch.cuda.Stream()
    fe = _ffi.RadarFE(cfg, device=ctx["local_rank"], stream=stream.cuda_stream)
    pinned = [tuple(_pin(x) for x in (c[0], np.asarray(c[1], np.int32), np.asarray(c[2], np.float32), np.asarray(c[3], np.int32),
                                      np.asarray(c[4], np.float64))) for c in chunks]
    n_local = sum(len(c[1]) for c in chunks)
    gatherer = _shard.PoseGatherer(gather_total, world, rank, device="cuda") if (world > 1 and gather_total) else None
    # resident: one batch object per chunk (at least NB of them, cycling the chunks), scans uploaded once; at most NB
    # passes are in flight: more only pile up one-warp clique CTAs that hold shared memory the image kernels need
    NB = max(1, nb or args.batches)
    n_obj = max(len(chunks), min(NB, 8))
    res_batches = [fe.new_batch() for _ in range(n_obj)]
    for k, b in enumerate(res_batches):
        c = pinned[k % len(pinned)]
        b.upload(c[0], c[1], c[2], c[3], prev_pose=c[4], sync=False)
    fe.sync()

    def resident_passes(n_chunks_to_run):
        for i in range(n_chunks_to_run):
            if i >= NB:
                res_batches[(i - NB) % n_obj].wait()
            res_batches[i % n_obj].run_async(with_mds=with_mds)

    resident_passes(max(1, warmup) * len(chunks))
    _barrier(ctx, fe)
    n0 = fe.launch_count()
    fe.timer_start()
    resident_passes(steps * len(chunks))
    ms = fe.timer_stop_ms()
    launches = fe.launch_count() - n0
    _barrier(ctx, fe)
    results = []
    for b in res_batches[:len(chunks)]:
        r, _, _ = b.download()
        results.append(r.copy())
    fe.sync()
    for b in res_batches:
        b.close()
    # e2e: NB rotating batches, every chunk's scans H2D + poses D2H inside the timed region, then the NCCL gather
    NB = min(NB, max(2, len(chunks)), 8)
    rot = [fe.new_batch() for _ in range(NB)]
    outs = [b.alloc_outputs(pinned=True) for b in rot]
    host_res = np.zeros(n_local, _ffi.PAIR_RESULT_DTYPE)
    offs = np.concatenate([[0], np.cumsum([len(c[1]) for c in chunks])])

    def e2e_run(n_passes):
        """n_passes passes over the chunks with at most NB of them in flight (a pass's gather runs when its last chunk retires)"""
        inflight, state = [], {"retired": 0, "gathered": None}

        def retire():
            j, cj = inflight.pop(0)
            rot[j].wait()
            host_res[offs[cj]:offs[cj + 1]] = outs[j][0][:offs[cj + 1] - offs[cj]]
            state["retired"] += 1
            if state["retired"] % len(pinned) == 0 and gatherer is not None:
                state["gathered"] = gatherer.gather(_shard.pack_records(host_res))

        for it in range(n_passes * len(pinned)):
            ci, k = it % len(pinned), it % NB
            if len(inflight) >= NB:
                retire()
            c = pinned[ci]
            rot[k].upload(c[0], c[1], c[2], c[3], prev_pose=c[4], sync=False)
            rot[k].run_async(with_mds=with_mds)
            rot[k].download(outs[k], sync=False, want_tracks=False)
            inflight.append((k, ci))
        while inflight:
            retire()
        return state["gathered"]

    e2e_run(max(1, min(warmup, 2)))
    _barrier(ctx, fe)
    t0 = time.perf_counter()
    gathered = e2e_run(steps)
    torch.cuda.synchronize()
    ms_e2e = (time.perf_counter() - t0) * 1e3
    _barrier(ctx, fe)
    for b in rot:
        b.close()
    fe.close()
    ms, ms_e2e = _reduce_max(ctx, ms, ms_e2e)
    h2d = sum(c[0].shape[0] * cfg.azimuths * (cfg.meta_bytes + ctx["rb"]) + len(c[1]) * (8 + ctx["kmax"] * 8 + 4 + 24) for c in chunks)
    return {"ms": ms / steps, "ms_e2e": ms_e2e / steps, "results": np.concatenate(results) if results else None,
            "launches_per_step": launches / steps, "h2d_bytes": int(h2d), "d2h_bytes": int(n_local * 120), "gathered": gathered}


# ---- configs[3]: strong scaling over a fixed set of independent pairs ------------------------------------------------
STRONG_SEQ_PAIRS = 256


def strong_data(args, rank, world, rb, kmax, workers):
    """BASELINE configs[3]: --pairs independent frame pairs = consecutive pairs of pairs/256 synthetic drives (seeds
    1000 + q), split into contiguous blocks over the ranks (_shard.shard_pairs).  Every rank renders only the scans its
    block touches.  Returns the block as chunks of <= 256 pairs (one rf_batch each)."""
    from radarslampy_b200 import _shard, synthetic as S
    P_total = args.pairs
    F = STRONG_SEQ_PAIRS + 1
    q = np.arange(P_total) // STRONG_SEQ_PAIRS
    i = np.arange(P_total) % STRONG_SEQ_PAIRS
    pair_idx = np.stack([q * F + i, q * F + i + 1], 1)
    lo, hi, frame_ids, local_idx = _shard.shard_pairs(pair_idx, world, rank)
    chunks = []
    for a in range(0, hi - lo, STRONG_SEQ_PAIRS):
        b = min(hi - lo, a + STRONG_SEQ_PAIRS)
        ids, inv = np.unique(local_idx[a:b].ravel(), return_inverse=True)
        gids = frame_ids[ids]
        raw = np.empty((len(gids), S.A, S.RAW_WIDTH), np.uint8)
        for seq in np.unique(gids // F):
            sel = np.flatnonzero(gids // F == seq)
            ks = gids[sel] % F
            seq_raw, _ = S.make_sequence_parallel(int(ks.max() - ks.min() + 1), res_m=args.res, seed=1000 + int(seq), first=int(ks.min()),
                                                  workers=workers)
            raw[sel] = seq_raw[ks - ks.min()]
        feats = np.zeros((b - a, kmax, 2), np.float32)
        counts = np.zeros(b - a, np.int32)
        prev = np.zeros((b - a, 3))
        worlds = {}
        for j in range(b - a):
            g = int(gids[inv.reshape(-1, 2)[j, 0]])
            seq, k = g // F, g % F
            w = worlds.setdefault(seq, S.World(seed=1000 + seq))
            prev[j] = S.twist_pose(k)
            f = S.scatterer_features(w, prev[j], args.res, rb, k=args.features)
            counts[j] = len(f)
            feats[j, :len(f)] = f
        chunks.append((raw, inv.reshape(-1, 2).astype(np.int32), feats, counts, prev))
    return {"lo": lo, "hi": hi, "chunks": chunks, "pairs_total": P_total}


def leg_strong(ctx, data):
    args, world = ctx["args"], ctx["world"]
    m = measure_pair_chunks(ctx, data["chunks"], bool(args.mds), steps=3, warmup=1, max_frames=STRONG_SEQ_PAIRS + 2,
                            max_pairs=STRONG_SEQ_PAIRS, gather_total=data["pairs_total"])
    if ctx["rank"] != 0:
        return None
    P = data["pairs_total"]
    out = {"workload": f"BASELINE configs[3]: {P} independent frame pairs ({P // STRONG_SEQ_PAIRS} synthetic drives x {STRONG_SEQ_PAIRS} consecutive pairs, "
                       f"features handed in) in contiguous blocks over {world} rank(s) (_shard.shard_pairs), poses gathered to rank 0 over NCCL",
           "scaling": "strong", "pairs": P, "n_gpus": world, "value": P / (m["ms"] * 1e-3), "unit": "pairs/s", "ms_per_pass": m["ms"],
           "e2e": {"value": P / (m["ms_e2e"] * 1e-3), "unit": "pairs/s", "ms_per_pass": m["ms_e2e"],
                   "h2d_bytes_per_pass_rank0": m["h2d_bytes"], "d2h_bytes_per_pass_rank0": m["d2h_bytes"]},
           "rank0_block": [int(data["lo"]), int(data["hi"])], "passes": 3}
    if m["gathered"] is not None:
        from radarslampy_b200 import _shard
        g = _shard.unpack_records(m["gathered"])
        out["gathered_pairs"] = int(len(g["status"]))
        out["median_dtheta_rad_all_ranks"] = float(np.median(np.arctan2(g["R"][:, 1, 0], g["R"][:, 0, 0])))
        out["status_nonzero"] = int((g["status"] != 0).sum())
    else:
        r = m["results"]
        out["median_dtheta_rad_all_ranks"] = float(np.median(np.arctan2(r["R"][:, 2], r["R"][:, 0])))
        out["status_nonzero"] = int((r["status"] != 0).sum())
    return out


# ---- configs[2] -----------------------------------------------------------------------------------------------------
def leg_mds(ctx, data):
    if data is None:
        return {"skipped": "the headline already runs with --mds 1"}
    rb, kmax, raw, poses, pair_idx, feats, counts = data
    args = ctx["args"]
    m = measure_pair_chunks(ctx, [(raw, pair_idx, feats, counts, poses[:-1])], True, steps=max(5, min(args.steps * max(1, args.passes_per_step), 80)), warmup=3,
                            max_frames=args.frames, max_pairs=args.frames - 1)
    P = args.frames - 1
    r = m["results"]
    return {"workload": "BASELINE configs[2]: the same synthetic drive rendered with intra-scan motion distortion, motion-distortion LM solve "
                        "per pair (features handed in, pairs independent), pipelined like the headline",
            "value": P / (m["ms"] * 1e-3), "unit": "frames/s", "ms_per_step": m["ms"],
            "e2e": {"value": P / (m["ms_e2e"] * 1e-3), "unit": "frames/s", "ms_per_step": m["ms_e2e"], "h2d_bytes_per_step": m["h2d_bytes"],
                    "d2h_bytes_per_step": m["d2h_bytes"]},
            "median_mds_iters": float(np.median(r["mds_iters"])), "median_v_mps": float(np.median(np.hypot(r["mds_x"][:, 0], r["mds_x"][:, 1]))),
            "expected_v_mps": 10.0, "status_nonzero": int((r["status"] != 0).sum()), "launches_per_step": m["launches_per_step"]}


# ---- chained odometry (the real configs[1] / configs[2] loop) ----------------------------------------------------------
CHAIN_CPU_FRAMES = 5


def chained_cpu_worker(job):
    import cv2
    cv2.setNumThreads(1)
    from oracle import ref_detect, ref_system
    raw, res_m, with_mds = job
    t0 = time.perf_counter()
    ref_system.run_odometry(raw[:1], ref_detect.detect_min_eig, with_mds=with_mds, range_res_m=res_m)
    t_init = time.perf_counter() - t0
    t0 = time.perf_counter()
    r = ref_system.run_odometry(raw, ref_detect.detect_min_eig, with_mds=with_mds, range_res_m=res_m)
    t_all = time.perf_counter() - t0
    return {"poses": r["poses"], "n_tracked": r["n_tracked"], "retrack": r["retrack"], "n_features_in": r["n_features_in"],
            "t_steps": max(t_all - t_init, 1e-9)}


def chained_cpu_baseline(args, raw_long, cores):
    """The reference's system loop on the CPU (oracle/ref_system.run_odometry: cv2 / scipy / networkx calls + the keyframe
    bookkeeping; detector = oracle/ref_detect.detect_min_eig), one independent sequence per host core: sequence c =
    frames c .. c + CHAIN_CPU_FRAMES - 1 of the drive (the GPU leg's sequence c).  Steps only (the first-frame detection
    is timed separately and subtracted, as the GPU leg times steps after the reset)."""
    import multiprocessing as mp
    T = CHAIN_CPU_FRAMES
    n = max(1, min(cores, len(raw_long) - T))
    jobs = [(raw_long[c:c + T], args.res, True) for c in range(n)]
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(n) as pl:
        outs = pl.map(chained_cpu_worker, jobs, chunksize=1)
    wall = time.perf_counter() - t0
    t_steps = max(o["t_steps"] for o in outs)
    return {"value": n * (T - 1) / t_steps, "unit": "frames/s", "cores": n, "kind": "port",
            "sample": f"{n} sequences x {T - 1} chained frames through oracle/ref_system.py (reference's cv2/scipy/networkx calls, cv2.cornerMinEigenVal "
                      f"detector + the reference's SSC), one process per sequence, slowest sequence {t_steps:.2f} s of steps ({wall:.1f} s incl. first-frame detection)",
            "outs": outs}


def leg_chained(ctx, raw_long, cpu):
    from radarslampy_b200 import _ffi
    args, world, rank, rb = ctx["args"], ctx["world"], ctx["rank"], ctx["rb"]
    S, NG, K, W = args.chain_seq, max(1, args.chain_runners), args.chain_steps, CHAIN_WARMUP
    while S % NG:
        NG -= 1
    per, T = S // NG, W + K
    cfg = _fe_config(args, rb, 2, 2, 320)
    fe = _ffi.RadarFE(cfg, device=ctx["local_rank"])
    with_mds = True
    out = {"workload": f"BASELINE configs[1]/[2] as the reference runs them (RawROAMSystem.run): {S} independent sequences per GPU advance one frame per "
                       f"step in lock step on the device (rf_seq): features carried over from the previous frame's KLT survivors, on-device re-detection "
                       f"(structure-tensor response, NMS, SSC bisection) when <= 60 survive, keyframe bookkeeping, motion-distortion solve, one CUDA graph "
                       f"per step per runner; sequence s = frames s, s+1, ... of the synthetic drive (rendered with intra-scan motion distortion)",
           "sequences_per_gpu": S, "runners": NG, "steps": K, "warmup": W, "n_gpus": world, "mds": with_mds}
    # ---- resident: the drive is uploaded once, sequence s reads frame s + t --------------------------------------------
    runners = [fe.new_sequences(per, per + T + 1) for _ in range(NG)]
    for g, r in enumerate(runners):
        r.upload(0, raw_long[g * per:g * per + per + T + 1])
        r.reset(0, 1)
    fe.sync()
    for t in range(1, W + 1):
        for r in runners:
            r.step(t, 1, with_mds=with_mds, graph=True)
    _barrier(ctx, fe)
    n0 = fe.launch_count()
    fe.timer_start()
    for t in range(W + 1, T + 1):
        for r in runners:
            r.step(t, 1, with_mds=with_mds, graph=True)
    ms = fe.timer_stop_ms()
    launches = fe.launch_count() - n0
    _barrier(ctx, fe)
    ring = runners[0].ring
    recs = np.concatenate([np.stack([r.results(t) for t in range(max(W + 1, T + 1 - ring), T + 1)], 1) for r in runners])
    for r in runners:
        r.close()
    # ---- e2e: every step's fresh scans (one per sequence) H2D from pinned memory, results D2H, inside the timed region ----
    SLOTS = 3
    host = [_pin(raw_long[g * per:g * per + per + T + 1]) for g in range(NG)]
    runners = [fe.new_sequences(per, SLOTS * per) for _ in range(NG)]
    res_host = [[_ffi.pinned_empty((per,), _ffi.SEQ_RESULT_DTYPE) for _ in range(SLOTS)] for _ in range(NG)]
    for g, r in enumerate(runners):
        r.upload(0, host[g][0:per])
        r.reset(0, 1)
    fe.sync()

    def e2e_steps(t_from, t_to):
        for t in range(t_from, t_to):
            slot = t % SLOTS
            for g, r in enumerate(runners):
                r.upload(slot * per, host[g][t:t + per])        # frame s + t of every sequence s of this runner
                r.step(slot * per, 1, with_mds=with_mds, graph=True)
                r.results(t, out=res_host[g][slot], sync=False)
        for r in runners:
            r.sync()

    e2e_steps(1, W + 1)
    _barrier(ctx, fe)
    t0 = time.perf_counter()
    e2e_steps(W + 1, T + 1)
    ms_e2e = (time.perf_counter() - t0) * 1e3
    _barrier(ctx, fe)
    e2e_last = np.concatenate([res_host[g][T % SLOTS] for g in range(NG)])
    for r in runners:
        r.close()
    # ---- the same resident measurement with more sequences per GPU (throughput rather than per-step latency) -------------
    large = None
    S2 = args.chain_seq_large
    if S2 > S and len(raw_long) >= S2 + T + 1:
        NG2 = max(1, S2 // per)
        while S2 % NG2:
            NG2 -= 1
        per2, K2 = S2 // NG2, max(4, K // 2)
        T2 = W + K2
        runners = [fe.new_sequences(per2, per2 + T2 + 1) for _ in range(NG2)]
        for g, r in enumerate(runners):
            r.upload(0, raw_long[g * per2:g * per2 + per2 + T2 + 1])
            r.reset(0, 1)
        fe.sync()
        for t in range(1, W + 1):
            for r in runners:
                r.step(t, 1, with_mds=with_mds, graph=True)
        _barrier(ctx, fe)
        fe.timer_start()
        for t in range(W + 1, T2 + 1):
            for r in runners:
                r.step(t, 1, with_mds=with_mds, graph=True)
        ms2 = fe.timer_stop_ms()
        _barrier(ctx, fe)
        rec2 = np.concatenate([r.results(T2) for r in runners])
        for r in runners:
            r.close()
        (ms2,) = _reduce_max(ctx, ms2)
        large = {"sequences_per_gpu": S2, "runners": NG2, "steps": K2, "value": world * S2 * K2 / (ms2 * 1e-3), "unit": "frames/s",
                 "ms_per_step": ms2 / K2, "status_nonzero": int((rec2["status"] != 0).sum()),
                 "note": "resident; sequence s = frames s, s + 1, ... of the same drive (every sequence a distinct window)"}
    # ---- single-sequence latency: one sequence, graph replay, no host sync between steps -------------------------------
    one = fe.new_sequences(1, T + 2)
    one.upload(0, raw_long[:T + 2])
    one.reset(0, 1)
    for t in range(1, W + 1):
        one.step(t, 1, with_mds=with_mds, graph=True)
    fe.sync()
    fe.timer_start()
    for t in range(W + 1, T + 1):
        one.step(t, 1, with_mds=with_mds, graph=True)
    ms1 = fe.timer_stop_ms()
    l1 = one.launches_per_step
    one.close()
    # ---- parity against the CPU loop on the sequences the cpu leg ran ---------------------------------------------------
    parity = None
    if cpu is not None and cpu.get("outs"):
        outs = cpu["outs"]
        n, Tc = len(outs), CHAIN_CPU_FRAMES
        pr = fe.new_sequences(n, n + Tc)
        pr.upload(0, raw_long[:n + Tc])
        pr.reset(0, 1)
        steps = []
        for t in range(1, Tc):
            pr.step(t, 1, with_mds=True, graph=True)
            steps.append(pr.results(t).copy())
        pr.close()
        g = np.stack(steps, 1)                                                    # [n, Tc - 1]
        want = np.stack([o["poses"][1:] for o in outs])                           # [n, Tc - 1, 3]
        d = g["pose"] - want
        dth = np.abs((d[..., 2] + np.pi) % (2 * np.pi) - np.pi)
        same = np.array([g["n_tracked"][c].tolist() == list(outs[c]["n_tracked"]) for c in range(n)])
        parity = {"sequences": n, "frames_each": Tc - 1, "max_abs_dxy_m": float(np.abs(d[..., :2]).max()), "max_abs_dtheta_rad": float(dth.max()),
                  "n_tracked_equal_sequences": int(same.sum()),
                  "retrack_equal_sequences": int(sum(g["retrack"][c].astype(bool).tolist() == list(outs[c]["retrack"]) for c in range(n))),
                  "tolerance": f"absolute pose after k chained frames within k x (1e-4 m, 1e-5 rad); here k <= {Tc - 1}",
                  "note": "the CPU side detects on cv2.cornerMinEigenVal's response, the device on its own (pinned to it within 2e-6): "
                          "the selection is bit-exact only for identical responses (tests/test_gpu_seq.py hands the device response to both)",
                  "pass": bool(np.abs(d[..., :2]).max() <= (Tc - 1) * 1e-4 and dth.max() <= (Tc - 1) * 1e-5 and same.all())}
    fe.close()
    ms, ms_e2e = _reduce_max(ctx, ms, ms_e2e)
    if rank != 0:
        return None
    cfgA, metaB = 400, 11
    out.update({
        "value": world * S * K / (ms * 1e-3), "unit": "frames/s", "ms_per_step": ms / K, "launches_per_step": launches / K,
        "gpu_launches": int(launches),
        "e2e": {"value": world * S * K / (ms_e2e * 1e-3), "unit": "frames/s", "ms_per_step": ms_e2e / K,
                "h2d_bytes_per_step": int(world * S * cfgA * rb), "d2h_bytes_per_step": int(world * S * _ffi.SEQ_RESULT_DTYPE.itemsize),
                "note": "one fresh scan per sequence per step (used bins only, 2-D copy from pinned host memory) + the step's result records"},
        "single_sequence": {"ms_per_frame": ms1 / K, "launches_per_step": l1, "note": "one sequence, one CUDA graph launch per frame, device-timed"},
        "retrack_fraction": float(recs["retrack"].mean()), "median_tracked": float(np.median(recs["n_tracked"])),
        "median_features_in": float(np.median(recs["n_features_in"])), "status_nonzero": int((recs["status"] != 0).sum()),
        "median_step_m": float(np.median(np.hypot(recs["h"][..., 0], recs["h"][..., 1]))), "expected_step_m": 2.5,
        "median_dtheta_rad": float(np.median(np.arctan2(recs["R"][..., 2], recs["R"][..., 0]))), "expected_dtheta_rad": 0.025,
        "e2e_status_nonzero": int((e2e_last["status"] != 0).sum()),
    })
    if large is not None:
        out["large"] = large
    if cpu is not None:
        out["cpu_baseline"] = {k: v for k, v in cpu.items() if k != "outs"}
    if parity is not None:
        out["parity_sample"] = parity
    return out


# ---- configs[4] and the single-stage legs ------------------------------------------------------------------------------
def leg_stress(ctx):
    """BASELINE configs[4]: 2000^2 Cartesian grid, 10 k SSC-selected features, 4-level KLT -> tracks/s (tools/klt_stress.py)."""
    import cv2
    from radarslampy_b200 import _ffi, synthetic as S
    N = 2000
    cfg = _ffi.default_config()
    cfg.range_bins, cfg.downsample = N + 1, 2
    cfg.raw_width = cfg.meta_bytes + cfg.range_bins
    cfg.max_frames, cfg.max_pairs, cfg.max_features = 2, 1, 2048
    fe = _ffi.RadarFE(cfg, device=ctx["local_rank"])
    a, b = S.dense_scene(5, n=N), S.dense_scene(5, shift=(1.7, -0.9), n=N)
    fa, fb = fe.frame_from_cart(a), fe.frame_from_cart(b)
    cand, n = fe.detect(fa, -0.01, cap=200000)
    fe.ssc(cand, 10000, 0.1, N, N)
    t0 = time.perf_counter()
    sel = fe.ssc(cand, 10000, 0.1, N, N)
    t_ssc = time.perf_counter() - t0
    t0 = time.perf_counter()
    fe.detect(fa, -0.01, cap=200000)
    t_det = time.perf_counter() - t0
    pts = np.ascontiguousarray(cand[sel][:, [1, 0]], np.float32)
    fe.klt(fa, fb, pts, apply_err_gate=True)
    reps = 20
    t0 = time.perf_counter()
    fe.timer_start()
    for _ in range(reps):
        nxt, st, err = fe.klt(fa, fb, pts, apply_err_gate=True)
    ms_dev = fe.timer_stop_ms() / reps
    dt = (time.perf_counter() - t0) / reps
    u8a, u8b = fa.download(1), fb.download(1)
    cv2.setNumThreads(0)
    t0 = time.perf_counter()
    cv_nxt, cv_st, _ = cv2.calcOpticalFlowPyrLK(u8a, u8b, pts, None, winSize=(15, 15), maxLevel=3, criteria=(3, 10, 0.03))
    dt_cv = time.perf_counter() - t0
    g = (st.ravel() > 0) & (cv_st.ravel() > 0)
    K = len(pts)
    fe.close()
    return {"workload": "BASELINE configs[4]: dense 2000^2 Cartesian grid (60 k Gaussian scatterers + speckle), SSC -> 10 k features, KLT 15x15 x 4 levels",
            "features": K, "candidates": int(n), "value": K / (ms_dev * 1e-3), "unit": "tracks/s", "klt_ms_device": ms_dev,
            "e2e": {"value": K / dt, "unit": "tracks/s", "note": "rf_klt with host points in / host results out (H2D + kernel + D2H per call)",
                    "h2d_bytes_per_step": K * 8, "d2h_bytes_per_step": K * 13},
            "ssc_ms": 1e3 * t_ssc, "detect_ms": 1e3 * t_det, "tracked_fraction": float(st.mean()),
            "cpu_baseline": {"value": K / dt_cv, "unit": "tracks/s", "cores": cv2.getNumThreads(), "kind": "port",
                             "sample": "the same 10 k points through cv2.calcOpticalFlowPyrLK once, OpenCV's own thread pool"},
            "parity_sample": {"status_equal": int((st.ravel() == cv_st.ravel()).sum()), "of": K,
                              "max_abs_dpx": float(np.abs(nxt[g] - cv_nxt.reshape(-1, 2)[g]).max()) if g.any() else None, "tolerance_px": 0.02}}


def leg_detect_ssc(ctx, raw_long):
    """a9 / a10 on the device: first-frame detection of 64 sequences in one reset (scan -> pyramid -> response -> NMS ->
    candidate sort -> SSC bisection -> append), no host synchronisation; and the single-frame host-API path."""
    from radarslampy_b200 import _ffi
    args, rb = ctx["args"], ctx["rb"]
    cfg = _fe_config(args, rb, 2, 2, 320)
    fe = _ffi.RadarFE(cfg, device=ctx["local_rank"])
    S = min(64, len(raw_long))
    r = fe.new_sequences(S, S)
    r.upload(0, raw_long[:S])
    r.reset(0, 1)
    fe.sync()
    reps = 5
    fe.timer_start()
    for _ in range(reps):
        r.reset(0, 1)
    ms = fe.timer_stop_ms() / reps
    rec = r.results(0)
    r.close()
    frame, cart = fe.polar_to_cart(raw=raw_long[0])
    fe.detect(frame, -0.01)
    t0 = time.perf_counter()
    cand, n = fe.detect(frame, -0.01)
    t_det = time.perf_counter() - t0
    fe.ssc(cand, 200, 0.1, fe.n, fe.n)
    t0 = time.perf_counter()
    sel = fe.ssc(cand, 200, 0.1, fe.n, fe.n)
    t_ssc = time.perf_counter() - t0
    fe.close()
    return {"workload": f"first-frame detection of {S} sequences in one rf_seq reset (scan -> u8 pyramid + f32 image, min-eigenvalue response, 3x3 NMS, "
                        f"candidate sort, SSC bisection to 200 +- 10 %, append), device-timed; single frame through rf_detect / rf_ssc (host arrays)",
            "value": S / (ms * 1e-3), "unit": "frames/s", "ms_per_frame_batched": ms / S, "frames": S,
            "median_features": float(np.median(rec["n_features_out"])), "median_candidates": float(np.median(rec["n_candidates"])),
            "single_frame": {"detect_ms": 1e3 * t_det, "ssc_ms": 1e3 * t_ssc, "candidates": int(n), "selected": int(len(sel))}}


def leg_fmt(ctx, raw_np):
    """N1: FMT rotation prior of every consecutive pair of the resident scans (rf_batch_fmt)."""
    from radarslampy_b200 import _ffi
    args, rb = ctx["args"], ctx["rb"]
    F = min(len(raw_np), 256)
    cfg = _fe_config(args, rb, F, F - 1, 64)
    fe = _ffi.RadarFE(cfg, device=ctx["local_rank"])
    b = fe.new_batch()
    pairs = np.stack([np.arange(F - 1), np.arange(1, F)], 1).astype(np.int32)
    b.upload(raw_np[:F], pairs, np.zeros((F - 1, 64, 2), np.float32), np.zeros(F - 1, np.int32))
    clip = int(87.5 / (2 * args.res))
    b.fmt_rotation(10, clip)
    reps = 5
    n0 = fe.launch_count()
    t0 = time.perf_counter()
    for _ in range(reps):
        ang, sc, resp, sh = b.fmt_rotation(10, clip)
    dt = (time.perf_counter() - t0) / reps
    launches = (fe.launch_count() - n0) // reps
    b.close()
    fe.close()
    return {"workload": f"FMT rotation prior (Tracker.py:62-63, FMT.py:13-90) of {F - 1} consecutive pairs from the resident u8 scans, synchronous call incl. D2H",
            "value": (F - 1) / dt, "unit": "pairs/s", "ms_per_batch": 1e3 * dt, "launches_per_batch": int(launches),
            "median_angle_rad": float(np.median(ang)), "note": "the reference's prior is taken between log-polar images of the raw polar scans"}


def leg_peaks(ctx, raw_np):
    """a12: getPointCloudPolarInd on one scan (host f32 polar in, int64 index pairs out)."""
    from radarslampy_b200 import _ffi
    args, rb = ctx["args"], ctx["rb"]
    cfg = _fe_config(args, rb, 2, 1, 64)
    fe = _ffi.RadarFE(cfg, device=ctx["local_rank"])
    polar = fe.extract_polar(raw_np[0])[0]
    fe.polar_peaks(polar)
    reps = 10
    t0 = time.perf_counter()
    for _ in range(reps):
        pk = fe.polar_peaks(polar)
    dt = (time.perf_counter() - t0) / reps
    fe.close()
    return {"workload": f"getPointCloudPolarInd (getPointCloud.py:10-60) on one {polar.shape[0]} x {polar.shape[1]} f32 polar scan, host in / host out",
            "value": 1.0 / dt, "unit": "scans/s", "ms_per_scan": 1e3 * dt, "peaks": int(len(pk))}


def leg_clique_real(ctx):
    """a6 on REAL graphs: the 10 data/tiny pairs (edge density 0.24-0.82, up to 36 k maximal cliques) and the reference's
    outlier_test.npz fixture (139 nodes, 8 097 maximal cliques), each through rf_reject_outliers alone on the device (CUDA events:
    H2D of the points, k_adjacency, k_maxclique, k_clique, D2H of the mask), masks against the goldens recorded from the
    unmodified reference (tests/golden).  The synthetic batches of the headline are inlier-dominated; these are not."""
    from radarslampy_b200 import _ffi
    g = os.path.join(ROOT, "tests", "golden")
    st, fx = np.load(os.path.join(g, "tiny_stages.npz")), np.load(os.path.join(g, "clique_fixture.npz"))
    cases = [(f"tiny pair {i}", st[f"klt_good_old_{i}"], st[f"klt_good_new_{i}"], st[f"rej_mask_{i}"]) for i in range(10)]
    cases.append(("outlier_test.npz", fx["prev"], fx["new"], fx["mask"]))
    fe = _ffi.RadarFE(device=ctx["local_rank"])        # the reference's own constants (0.0432 m/bin: threshold 5.79 px)
    rows = []
    for name, prev, new, want in cases:
        fe.reject_outliers(prev, new)
        best = 1e9
        for _ in range(5):
            fe.timer_start()
            mask, n_in, nodes = fe.reject_outliers(prev, new)
            best = min(best, fe.timer_stop_ms())
        rows.append({"case": name, "nodes": int(len(prev)), "clique_size": int(n_in), "descents": int(nodes), "ms": round(best, 4),
                     "mask_exact": bool(np.array_equal(mask, want.astype(bool)))})
    fe.close()
    ms = [r["ms"] for r in rows]
    return {"workload": "rejectOutliers (outlierRejection.py:16-95) on the 10 data/tiny pairs and the reference's outlier_test.npz, one graph "
                        "at a time, device-timed",
            "value": 1e3 / float(np.mean(ms)), "unit": "graphs/s (one at a time)", "worst_tiny_ms": max(ms[:10]), "median_ms": float(np.median(ms)),
            "fixture_ms": ms[-1], "all_masks_exact": all(r["mask_exact"] for r in rows), "rows": rows,
            "networkx_ms": "22-729 ms per graph on one host core (profiles/r02f_clique_table.json)"}


if __name__ == "__main__":
    main()
