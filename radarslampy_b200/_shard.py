"""Multi-GPU sharding of the front end: independent frame pairs — or whole
sequences — are split into contiguous blocks, one block per rank (one process per GPU); the only
exchange is the gather of the fixed-width pose records to rank 0, which concatenates them in pair
order ("trajectory concatenation").  Pose chaining inside one sequence is sequential
(RawROAMSystem.py:296-298), so a single sequence does not shard: replicas only.

The transport is torch.distributed: NCCL over NVLink on the GPU box, gloo in the CPU tests
(tests/test_shard_gloo.py).  Nothing here touches the kernels; every rank calls
_ffi.Batch.track on its own block."""
import numpy as np

# one gathered record per pair: R(4) h(2) mds_x(6) n_features n_good n_inliers status  -> 16 f64 = 128 B
RECORD_WIDTH = 16


def block_range(n_items: int, world: int, rank: int):
    """Contiguous block [lo, hi) of `n_items` owned by `rank`: item p goes to rank p * world // n_items
    up to rounding, i.e. the first (n_items % world) ranks hold one extra item."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank {rank} / world {world}")
    base, extra = divmod(int(n_items), world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_pairs(pair_idx: np.ndarray, world: int, rank: int):
    """Block of a global [P,2] pair list for `rank`, re-indexed to the frames it actually touches.
    Returns (lo, hi, frame_ids, local_pair_idx): frame_ids are the global frame indices this rank must
    hold (sorted, unique) and local_pair_idx indexes into them."""
    pair_idx = np.asarray(pair_idx, np.int64).reshape(-1, 2)
    lo, hi = block_range(pair_idx.shape[0], world, rank)
    blk = pair_idx[lo:hi]
    frame_ids, inv = np.unique(blk.ravel(), return_inverse=True)
    return lo, hi, frame_ids, inv.reshape(-1, 2).astype(np.int32)


def pack_records(res: np.ndarray) -> np.ndarray:
    """rf_pair_result structured array -> [P, RECORD_WIDTH] f64 wire format."""
    P = res.shape[0]
    out = np.empty((P, RECORD_WIDTH), np.float64)
    out[:, 0:4] = res["R"]
    out[:, 4:6] = res["h"]
    out[:, 6:12] = res["mds_x"]
    out[:, 12] = res["n_features"]
    out[:, 13] = res["n_good"]
    out[:, 14] = res["n_inliers"]
    out[:, 15] = res["status"]
    return out


def unpack_records(rec: np.ndarray) -> dict:
    rec = np.asarray(rec, np.float64).reshape(-1, RECORD_WIDTH)
    return {"R": rec[:, 0:4].reshape(-1, 2, 2), "h": rec[:, 4:6], "mds_x": rec[:, 6:12],
            "n_features": rec[:, 12].astype(np.int32), "n_good": rec[:, 13].astype(np.int32),
            "n_inliers": rec[:, 14].astype(np.int32), "status": rec[:, 15].astype(np.int32)}


class PoseGatherer:
    """Gathers per-rank record blocks of (possibly different) known sizes to rank 0.

    The block sizes follow from block_range, so no size exchange is needed: every rank pads its block
    to the largest block and one dist.gather moves it.  Buffers are allocated once."""

    def __init__(self, n_pairs_total: int, world: int, rank: int, device="cpu", group=None):
        import torch
        self.torch = torch
        self.world, self.rank, self.group = world, rank, group
        self.sizes = [block_range(n_pairs_total, world, r) for r in range(world)]
        self.max_block = max(hi - lo for lo, hi in self.sizes) if world else 0
        self.n_total = int(n_pairs_total)
        self.send = torch.zeros((self.max_block, RECORD_WIDTH), dtype=torch.float64, device=device)
        self.recv = ([torch.zeros_like(self.send) for _ in range(world)] if (rank == 0 and world > 1) else None)
        # asynchronous path (device tensors only): pinned staging on both sides, so that neither the H2D of the records nor
        # the read-back on rank 0 blocks the host thread that feeds the upload / kernel pipeline
        self._cuda = str(device).startswith("cuda")
        self._pin = torch.zeros((self.max_block, RECORD_WIDTH), dtype=torch.float64, pin_memory=True) if self._cuda else None
        self._recv_host = ([torch.zeros((self.max_block, RECORD_WIDTH), dtype=torch.float64, pin_memory=True) for _ in range(world)]
                           if (self._cuda and rank == 0 and world > 1) else None)
        self._ev = torch.cuda.Event() if self._cuda else None
        self._pending = False

    def gather(self, records: np.ndarray):
        """records: this rank's [P_local, RECORD_WIDTH].  Returns the concatenated [P_total, RECORD_WIDTH]
        array on rank 0, None elsewhere."""
        torch = self.torch
        lo, hi = self.sizes[self.rank]
        if records.shape != (hi - lo, RECORD_WIDTH):
            raise ValueError(f"rank {self.rank} owns pairs [{lo},{hi}) but got records of shape {records.shape}")
        if self.world == 1:
            return np.array(records, np.float64)
        self.send[:hi - lo].copy_(torch.from_numpy(np.ascontiguousarray(records)), non_blocking=True)
        import torch.distributed as dist
        dist.gather(self.send, self.recv, dst=0, group=self.group)
        if self.rank != 0:
            return None
        out = np.empty((self.n_total, RECORD_WIDTH), np.float64)
        for r, (a, b) in enumerate(self.sizes):
            out[a:b] = self.recv[r][:b - a].cpu().numpy()
        return out


    def gather_async(self, records: np.ndarray):
        """Queue one gather without blocking the host: records -> pinned staging -> device -> dist.gather -> (rank 0) pinned
        read-back, all on torch's current stream.  `result()` waits for it.  A small pageable H2D copy would block the
        caller behind every scan upload already queued on the copy engine (measured: 1-4 ms per gather at 2 GPUs)."""
        torch = self.torch
        if self.world == 1 or not self._cuda:
            self._last = self.gather(records)
            return
        lo, hi = self.sizes[self.rank]
        if records.shape != (hi - lo, RECORD_WIDTH):
            raise ValueError(f"rank {self.rank} owns pairs [{lo},{hi}) but got records of shape {records.shape}")
        if self._pending:
            self._ev.synchronize()          # the staging buffers are reused: the previous gather must have drained
        self._pin[:hi - lo].copy_(torch.from_numpy(np.ascontiguousarray(records)))
        self.send[:hi - lo].copy_(self._pin[:hi - lo], non_blocking=True)
        import torch.distributed as dist
        dist.gather(self.send, self.recv, dst=0, group=self.group)
        if self.rank == 0:
            for r in range(self.world):
                self._recv_host[r].copy_(self.recv[r], non_blocking=True)
        self._ev.record()
        self._pending = True

    def result(self):
        """Wait for the last gather_async; the concatenated [P_total, RECORD_WIDTH] array on rank 0, None elsewhere."""
        if self.world == 1 or not self._cuda:
            return getattr(self, "_last", None)
        if not self._pending:
            return None
        self._ev.synchronize()
        self._pending = False
        if self.rank != 0:
            return None
        out = np.empty((self.n_total, RECORD_WIDTH), np.float64)
        for r, (a, b) in enumerate(self.sizes):
            out[a:b] = self._recv_host[r][:b - a].numpy()
        return out


def chain_poses(R: np.ndarray, h: np.ndarray, start=None) -> np.ndarray:
    """Trajectory from per-pair relative transforms: T_{k+1} = T_k @ [R_k h_k; 0 1]
    (RawROAMSystem.py:201, trajectoryPlotting.py:116-123).  Returns [P+1,3] (x, y, theta)."""
    P = R.shape[0]
    T = np.eye(3) if start is None else np.array(start, np.float64)
    out = np.empty((P + 1, 3))
    out[0] = (T[0, 2], T[1, 2], np.arctan2(T[1, 0], T[0, 0]))
    for k in range(P):
        A = np.eye(3)
        A[:2, :2] = np.asarray(R[k]).reshape(2, 2)
        A[:2, 2] = np.asarray(h[k]).ravel()
        T = T @ A
        out[k + 1] = (T[0, 2], T[1, 2], np.arctan2(T[1, 0], T[0, 0]))
    return out
