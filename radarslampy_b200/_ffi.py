"""ctypes binding of libradarfe.so (include/radarfe.h).  The ONLY module that touches the
shared library; every drop-in module goes through `RadarFE`.

There is no CPU fallback: if the library is missing, or no sm_100 GPU is visible,
construction raises."""
import ctypes as C
import os
import threading

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RADARFE_LIB") or os.path.join(_PKG, "libradarfe.so")   # RADARFE_LIB: A/B builds of the same library

RF_OK, RF_E_BADARG, RF_E_CAPACITY, RF_E_CUDA, RF_E_WORKLIMIT, RF_E_NOMEM = 0, -1, -2, -3, -4, -5


class RfConfig(C.Structure):
    _fields_ = [
        ("azimuths", C.c_int32), ("raw_width", C.c_int32), ("meta_bytes", C.c_int32), ("range_bins", C.c_int32),
        ("downsample", C.c_int32), ("max_features", C.c_int32), ("max_pairs", C.c_int32), ("max_frames", C.c_int32),
        ("klt_win", C.c_int32), ("klt_max_level", C.c_int32), ("klt_max_iters", C.c_int32),
        ("klt_eps", C.c_float), ("klt_min_eig", C.c_float), ("klt_err_thr", C.c_float),
        ("dist_thr_px", C.c_double), ("cart_res_m", C.c_double), ("mds_period", C.c_double),
        ("mds_sigma_p", C.c_double * 2), ("mds_sigma_v", C.c_double * 3),
        ("clique_node_limit", C.c_int64), ("write_cart_f32", C.c_int32), ("retrack_threshold", C.c_int32),
        ("ssc_num_ret", C.c_int32), ("doh_num_sigma", C.c_int32), ("ssc_tolerance", C.c_double),
        ("kf_rot_thr", C.c_double), ("kf_trans_thr", C.c_double), ("detect_quality", C.c_double),
        ("doh_min_sigma", C.c_double), ("doh_max_sigma", C.c_double), ("doh_threshold", C.c_double),
    ]


class RfPairResult(C.Structure):
    _fields_ = [
        ("R", C.c_double * 4), ("h", C.c_double * 2), ("mds_x", C.c_double * 6),
        ("n_features", C.c_int32), ("n_good", C.c_int32), ("n_inliers", C.c_int32), ("mds_iters", C.c_int32),
        ("status", C.c_int32), ("clique_nodes", C.c_int32),
    ]


PAIR_RESULT_DTYPE = np.dtype([
    ("R", np.float64, (4,)), ("h", np.float64, (2,)), ("mds_x", np.float64, (6,)),
    ("n_features", np.int32), ("n_good", np.int32), ("n_inliers", np.int32), ("mds_iters", np.int32),
    ("status", np.int32), ("clique_nodes", np.int32)], align=True)
assert PAIR_RESULT_DTYPE.itemsize == C.sizeof(RfPairResult)

SEQ_RESULT_DTYPE = np.dtype([
    ("pose", np.float64, (3,)), ("R", np.float64, (4,)), ("h", np.float64, (2,)), ("mds_x", np.float64, (6,)),
    ("kab_R", np.float64, (4,)), ("kab_h", np.float64, (2,)),
    ("n_features_in", np.int32), ("n_good", np.int32), ("n_tracked", np.int32), ("retrack", np.int32),
    ("keyframe_added", np.int32), ("n_keyframes", np.int32), ("n_features_out", np.int32), ("n_candidates", np.int32),
    ("mds_iters", np.int32), ("clique_nodes", np.int32), ("status", np.int32), ("reserved", np.int32)], align=True)
assert SEQ_RESULT_DTYPE.itemsize == 21 * 8 + 12 * 4
RF_SEQ_MDS, RF_SEQ_GRAPH = 1, 2

# every symbol include/radarfe.h declares (tests check the library exports all of them)
SYMBOLS = [
    "rf_default_config", "rf_create", "rf_destroy", "rf_last_error", "rf_version", "rf_cart_size", "rf_stream",
    "rf_timer_start", "rf_timer_stop_ms", "rf_launch_count", "rf_extract_polar", "rf_frame_create",
    "rf_frame_destroy", "rf_polar_to_cart", "rf_polar_to_cart_log", "rf_frame_from_cart", "rf_frame_download", "rf_klt",
    "rf_reject_outliers", "rf_reject_outliers_f64", "rf_consistency_adjacency", "rf_consistency_adjacency_f64", "rf_clique_search", "rf_kabsch", "rf_kabsch_f64", "rf_mds_solve", "rf_mds_undistort", "rf_mds_undistort_times", "rf_ssc",
    "rf_detect", "rf_detect_doh", "rf_doh_response", "rf_corner_response", "rf_nms_select", "rf_polar_peaks", "rf_batch_create", "rf_batch_destroy", "rf_batch_upload",
    "rf_batch_run_async", "rf_sync", "rf_batch_download", "rf_track_batch", "rf_track_pair", "rf_batch_upload_async",
    "rf_batch_download_async", "rf_batch_klt_status", "rf_batch_frame_download", "rf_batch_set_profiling",
    "rf_batch_stage_times", "rf_host_alloc", "rf_host_free", "rf_batch_wait", "rf_fmt_rotation", "rf_fmt_rotation_frames", "rf_fmt_log_polar", "rf_cart_to_polar",
    "rf_phase_correlate", "rf_batch_fmt", "rf_chain_poses", "rf_png_info", "rf_ingest_png",
    "rf_seq_create", "rf_seq_destroy", "rf_seq_upload_async", "rf_seq_reset_async", "rf_seq_step_async", "rf_seq_results",
    "rf_seq_results_async", "rf_seq_ring", "rf_seq_steps_done", "rf_seq_features", "rf_seq_sync", "rf_seq_launches_per_step",
]
STAGES = ("polar2cart", "scan_to_l0l1", "pyr_down", "klt", "compact", "reject", "kabsch", "mds", "finish")

_lib = None
_lib_lock = threading.Lock()


def load_library():
    """dlopen libradarfe.so (built by radarslampy_b200._build); raises if absent."""
    global _lib
    with _lib_lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"{LIB_PATH} not found: build it with `python -m radarslampy_b200._build` "
                    "(this package has no CPU fallback)")
            L = C.CDLL(LIB_PATH)
            L.rf_last_error.restype = C.c_char_p
            L.rf_last_error.argtypes = [C.c_void_p]
            L.rf_stream.restype = C.c_void_p
            L.rf_stream.argtypes = [C.c_void_p]
            L.rf_launch_count.restype = C.c_int64
            L.rf_launch_count.argtypes = [C.c_void_p]
            L.rf_destroy.restype = None
            L.rf_destroy.argtypes = [C.c_void_p]
            L.rf_frame_destroy.restype = None
            L.rf_frame_destroy.argtypes = [C.c_void_p, C.c_void_p]
            L.rf_batch_destroy.restype = None
            L.rf_batch_destroy.argtypes = [C.c_void_p, C.c_void_p]
            L.rf_default_config.restype = None
            L.rf_host_alloc.argtypes = [C.c_size_t, C.POINTER(C.c_void_p)]
            L.rf_host_free.restype = None
            L.rf_host_free.argtypes = [C.c_void_p]
            L.rf_seq_destroy.restype = None
            L.rf_seq_destroy.argtypes = [C.c_void_p, C.c_void_p]
            L.rf_seq_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
            L.rf_seq_upload_async.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
            L.rf_seq_reset_async.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
            L.rf_seq_step_async.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
            L.rf_seq_results.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
            L.rf_seq_results_async.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
            L.rf_seq_features.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
            L.rf_seq_sync.argtypes = [C.c_void_p, C.c_void_p]
            for fn in (L.rf_seq_ring, L.rf_seq_steps_done, L.rf_seq_launches_per_step):
                fn.argtypes = [C.c_void_p]
            _lib = L
    return _lib


def default_config() -> RfConfig:
    cfg = RfConfig()
    load_library().rf_default_config(C.byref(cfg))
    return cfg


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def png_info(path):
    """(rows, cols) of one scan file (rf_png_info)."""
    L = load_library()
    r, c = C.c_int(0), C.c_int(0)
    rc = L.rf_png_info(os.fsencode(path), C.byref(r), C.byref(c))
    if rc:
        raise ValueError(L.rf_last_error(None).decode())
    return r.value, c.value


def ingest_png(paths, threads=0, out=None, pinned=False):
    """Decode scan files in parallel (rf_ingest_png) -> u8 [n, rows, cols]; `out` may be a pinned staging array."""
    L = load_library()
    paths = [os.fsencode(p) for p in paths]
    n = len(paths)
    if n == 0:
        return np.empty((0, 0, 0), np.uint8)
    if out is None:
        rows, cols = png_info(paths[0])
        out = pinned_empty((n, rows, cols), np.uint8) if pinned else np.empty((n, rows, cols), np.uint8)
    if out.dtype != np.uint8 or out.ndim != 3 or out.shape[0] < n or not out.flags.c_contiguous:
        raise ValueError("ingest_png: `out` must be a C-contiguous u8 [>= n, rows, cols] array")
    arr = (C.c_char_p * n)(*paths)
    rc = L.rf_ingest_png(arr, n, out.shape[1], out.shape[2], int(threads), _ptr(out))
    if rc:
        raise ValueError(L.rf_last_error(None).decode())
    return out[:n]


class _PinnedOwner:
    def __init__(self, lib, ptr):
        self.lib, self.ptr = lib, ptr

    def __del__(self):
        try:
            if self.ptr:
                self.lib.rf_host_free(self.ptr)
                self.ptr = None
        except Exception:
            pass


def pinned_empty(shape, dtype):
    """NumPy array in page-locked host memory (for the *_async batch calls)."""
    import weakref
    lib = load_library()
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape)) * dtype.itemsize
    p = C.c_void_p()
    rc = lib.rf_host_alloc(C.c_size_t(max(nbytes, 1)), C.byref(p))
    if rc != RF_OK:
        msg = lib.rf_last_error(None)
        raise MemoryError(f"rf_host_alloc failed ({rc}): {msg.decode() if msg else ''}")
    buf = (C.c_uint8 * max(nbytes, 1)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    owner = _PinnedOwner(lib, p)
    weakref.finalize(buf, lambda o=owner: o.__del__())
    return arr


class _PinnedImagePool:
    """Recycled page-locked host buffers for the images the drop-in API hands back as NumPy arrays (16 MB per
    Cartesian image): a fresh np.empty pays page faults and a pageable D2H every frame.  A buffer returns to the pool
    when the last array referring to it is garbage-collected (arrays keep their base alive, so a live view can never
    see recycled memory).  At most MAX_OUT buffers are handed out at a time; callers that hoard images beyond that get
    ordinary pageable arrays, so pinned memory stays bounded."""
    MAX_OUT = 6

    def __init__(self):
        self.free = {}          # nbytes -> [pointer values]
        self.out = 0
        self.lock = threading.Lock()

    def empty(self, shape, dtype):
        import weakref
        dtype = np.dtype(dtype)
        count = int(np.prod(shape))
        nbytes = max(count * dtype.itemsize, 1)
        with self.lock:
            if self.out >= self.MAX_OUT:
                return np.empty(shape, dtype)
            lst = self.free.get(nbytes)
            ptr = lst.pop() if lst else None
            self.out += 1
        if ptr is None:
            lib = load_library()
            p = C.c_void_p()
            if lib.rf_host_alloc(C.c_size_t(nbytes), C.byref(p)) != RF_OK:
                with self.lock:
                    self.out -= 1
                return np.empty(shape, dtype)
            ptr = p.value
        buf = (C.c_uint8 * nbytes).from_address(ptr)
        weakref.finalize(buf, self._release, nbytes, ptr)
        return np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)

    def _release(self, nbytes, ptr):
        with self.lock:
            self.out -= 1
            self.free.setdefault(nbytes, []).append(ptr)


_image_pool = _PinnedImagePool()


class Batch:
    """Device-resident batch of independent frame pairs (rf_batch)."""

    def __init__(self, fe):
        self.fe = fe
        p = C.c_void_p()
        fe._check(fe.lib.rf_batch_create(fe.h, C.byref(p)))
        self.p = p
        self.n_pairs = 0
        self._keep = None

    def close(self):
        if self.p and self.fe.h:
            self.fe.lib.rf_batch_destroy(self.fe.h, self.p)
        self.p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _marshal(self, raw, pair_idx, feats, counts, prev_pose):
        fe = self.fe
        cfg = fe.cfg
        raw = np.ascontiguousarray(raw, np.uint8)
        if raw.ndim != 3 or raw.shape[1:] != (cfg.azimuths, cfg.raw_width):
            raise ValueError(f"raw scans must be [n_frames, {cfg.azimuths}, {cfg.raw_width}] uint8, got {raw.shape}")
        pair_idx = np.ascontiguousarray(pair_idx, np.int32).reshape(-1, 2)
        P = pair_idx.shape[0]
        feats = np.ascontiguousarray(feats, np.float32)
        if feats.shape != (P, cfg.max_features, 2):
            raise ValueError(f"feats must be [n_pairs, max_features={cfg.max_features}, 2] float32, got {feats.shape}")
        counts = np.ascontiguousarray(counts, np.int32).reshape(-1)
        if counts.shape[0] != P:
            raise ValueError("feat_counts must have one entry per pair")
        if prev_pose is not None:
            prev_pose = np.ascontiguousarray(prev_pose, np.float64).reshape(P, 3)
        return raw, pair_idx, feats, counts, prev_pose

    def upload(self, raw, pair_idx, feats, counts, prev_pose=None, sync=True):
        fe = self.fe
        raw, pair_idx, feats, counts, prev_pose = self._marshal(raw, pair_idx, feats, counts, prev_pose)
        fn = fe.lib.rf_batch_upload if sync else fe.lib.rf_batch_upload_async
        fe._check(fn(fe.h, self.p, _ptr(raw), raw.shape[0], _ptr(pair_idx), pair_idx.shape[0], _ptr(feats), _ptr(counts),
                     _ptr(prev_pose)))
        self.n_pairs = pair_idx.shape[0]
        self._keep = (raw, pair_idx, feats, counts, prev_pose)   # async: keep host buffers alive

    def run_async(self, with_mds=False):
        self.fe._check(self.fe.lib.rf_batch_run_async(self.fe.h, self.p, int(with_mds)))

    def wait(self):
        """Block until the last run + download queued for THIS batch has finished (rf_batch_wait)."""
        self.fe._check(self.fe.lib.rf_batch_wait(self.fe.h, self.p))

    def alloc_outputs(self, pinned=False):
        P, K = self.fe.cfg.max_pairs, self.fe.cfg.max_features
        mk = pinned_empty if pinned else (lambda shape, dt: np.empty(shape, dt))
        return mk((P,), PAIR_RESULT_DTYPE), mk((P, K, 2), np.float32), mk((P, K), np.uint8)

    def download(self, out=None, sync=True, want_tracks=True):
        fe = self.fe
        P = self.n_pairs
        res, nxt, st = out if out is not None else self.alloc_outputs()
        fn = fe.lib.rf_batch_download if sync else fe.lib.rf_batch_download_async
        fe._check(fn(fe.h, self.p, _ptr(res), _ptr(nxt) if want_tracks else None, _ptr(st) if want_tracks else None))
        return res[:P], nxt[:P], st[:P]

    def klt_status(self):
        fe = self.fe
        P, K = self.n_pairs, fe.cfg.max_features
        st = np.zeros((max(P, 1), K), np.uint8)
        err = np.zeros((max(P, 1), K), np.float32)
        fe._check(fe.lib.rf_batch_klt_status(fe.h, self.p, _ptr(st), _ptr(err)))
        return st[:P], err[:P]

    def fmt_rotation(self, downsample=10, clip_px=0):
        """FMT rotation prior (Tracker.py:62-63) of every pair of the uploaded batch, from the resident scans
        -> (angle_rad [P], scale [P], response [P], shift_xy [P, 2])."""
        fe, P = self.fe, self.n_pairs
        ang, sc, resp, sh = np.zeros(P), np.zeros(P), np.zeros(P), np.zeros((P, 2))
        fe._check(fe.lib.rf_batch_fmt(fe.h, self.p, int(downsample), int(clip_px), _ptr(ang), _ptr(sc), _ptr(resp), _ptr(sh)))
        return ang, sc, resp, sh

    def frame(self, idx, what=1):
        fe = self.fe
        r, c = C.c_int(0), C.c_int(0)
        fe._check(fe.lib.rf_batch_frame_download(fe.h, self.p, idx, what, None, C.byref(r), C.byref(c)))
        out = np.empty((r.value, c.value), np.float32 if what == 0 else np.uint8)
        fe._check(fe.lib.rf_batch_frame_download(fe.h, self.p, idx, what, _ptr(out), C.byref(r), C.byref(c)))
        return out

    def set_profiling(self, on=True):
        self.fe._check(self.fe.lib.rf_batch_set_profiling(self.fe.h, self.p, int(on)))

    def stage_times(self):
        """({stage: total ms}, n_runs) over the runs recorded since the last call."""
        ms = (C.c_float * len(STAGES))()
        n = C.c_int(0)
        self.fe._check(self.fe.lib.rf_batch_stage_times(self.fe.h, self.p, ms, len(STAGES), C.byref(n)))
        return dict(zip(STAGES, [float(v) for v in ms])), n.value

    def track(self, raw, pair_idx, feats, counts, prev_pose=None, with_mds=False):
        """upload + run + download (rf_track_batch)."""
        fe = self.fe
        raw, pair_idx, feats, counts, prev_pose = self._marshal(raw, pair_idx, feats, counts, prev_pose)
        P = pair_idx.shape[0]
        res, nxt, st = self.alloc_outputs()
        fe._check(fe.lib.rf_track_batch(fe.h, self.p, _ptr(raw), raw.shape[0], _ptr(pair_idx), P, _ptr(feats), _ptr(counts),
                                        _ptr(prev_pose), int(with_mds), _ptr(res), _ptr(nxt), _ptr(st)))
        self.n_pairs = P
        return res[:P], nxt[:P], st[:P]


class Sequences:
    """Lock-step runner of n_seq device-resident odometry chains (rf_seq, include/radarfe.h)."""

    def __init__(self, fe, n_seq, arena_frames, detector_mode=0):
        self.fe, self.n_seq, self.arena_frames = fe, int(n_seq), int(arena_frames)
        p = C.c_void_p()
        fe._check(fe.lib.rf_seq_create(fe.h, self.n_seq, self.arena_frames, int(detector_mode), C.byref(p)))
        self.p = p
        self._keep = []

    def close(self):
        if self.p and self.fe.h:
            self.fe.lib.rf_seq_destroy(self.fe.h, self.p)
        self.p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload(self, first_frame, raw):
        """raw u8 [n_frames, A, raw_width] -> arena frames [first_frame, first_frame + n_frames) (asynchronous: `raw` is
        kept alive until the next sync())."""
        cfg = self.fe.cfg
        raw = np.ascontiguousarray(raw, np.uint8)
        if raw.ndim != 3 or raw.shape[1:] != (cfg.azimuths, cfg.raw_width):
            raise ValueError(f"raw scans must be [n_frames, {cfg.azimuths}, {cfg.raw_width}] uint8, got {raw.shape}")
        self.fe._check(self.fe.lib.rf_seq_upload_async(self.fe.h, self.p, int(first_frame), raw.shape[0], _ptr(raw)))
        self._keep.append(raw)

    def reset(self, base=0, stride=1, init_pose=None):
        ip = None if init_pose is None else _c(np.asarray(init_pose, np.float64).reshape(self.n_seq, 3), np.float64)
        self.fe._check(self.fe.lib.rf_seq_reset_async(self.fe.h, self.p, int(base), int(stride), _ptr(ip)))

    def step(self, base, stride=1, with_mds=True, graph=True):
        flags = (RF_SEQ_MDS if with_mds else 0) | (RF_SEQ_GRAPH if graph else 0)
        self.fe._check(self.fe.lib.rf_seq_step_async(self.fe.h, self.p, int(base), int(stride), flags))

    @property
    def steps_done(self):
        return int(self.fe.lib.rf_seq_steps_done(self.p))

    @property
    def ring(self):
        return int(self.fe.lib.rf_seq_ring(self.p))

    @property
    def launches_per_step(self):
        return int(self.fe.lib.rf_seq_launches_per_step(self.p))

    def results(self, step, out=None, sync=True):
        out = np.zeros(self.n_seq, SEQ_RESULT_DTYPE) if out is None else out
        fn = self.fe.lib.rf_seq_results if sync else self.fe.lib.rf_seq_results_async
        self.fe._check(fn(self.fe.h, self.p, int(step), _ptr(out)))
        return out

    def features(self):
        K = self.fe.cfg.max_features
        feats = np.zeros((self.n_seq, K, 2), np.float32)
        counts = np.zeros(self.n_seq, np.int32)
        self.fe._check(self.fe.lib.rf_seq_features(self.fe.h, self.p, _ptr(feats), _ptr(counts)))
        return feats, counts

    def sync(self):
        self.fe._check(self.fe.lib.rf_seq_sync(self.fe.h, self.p))
        self._keep = []


class Frame:
    """Device-resident scan (f32 Cartesian image + u8 LK pyramid)."""

    POOL_MAX = 8    # released frames kept per handle: a sequential caller allocates device memory only for the first few scans

    def __init__(self, fe):
        self.fe = fe
        pool = fe._frame_pool
        if pool:
            self.p = pool.pop()          # every frame of a handle has the same geometry
            return
        p = C.c_void_p()
        fe._check(fe.lib.rf_frame_create(fe.h, C.byref(p)))
        self.p = p

    def close(self):
        """Release the frame: back to the handle's pool (cudaMalloc / cudaFree cost milliseconds per frame), or
        destroyed when the pool is full.  Every entry point that touches a frame is synchronous, so no work is pending."""
        if self.p and self.fe.h:
            if len(self.fe._frame_pool) < self.POOL_MAX:
                self.fe._frame_pool.append(self.p)
            else:
                self.fe.lib.rf_frame_destroy(self.fe.h, self.p)
        self.p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def download(self, what=0):
        fe = self.fe
        r, c = C.c_int(0), C.c_int(0)
        fe._check(fe.lib.rf_frame_download(fe.h, self.p, what, None, C.byref(r), C.byref(c)))
        out = np.empty((r.value, c.value), np.float32 if what == 0 else np.uint8)
        fe._check(fe.lib.rf_frame_download(fe.h, self.p, what, _ptr(out), C.byref(r), C.byref(c)))
        return out


class RadarFE:
    """One rf_handle: a CUDA stream, the geometry table and the workspaces on one GPU."""

    def __init__(self, cfg: RfConfig = None, device: int = 0, stream: int = None):
        self._frame_pool = []
        self.lib = load_library()
        self.cfg = cfg if cfg is not None else default_config()
        h = C.c_void_p()
        rc = self.lib.rf_create(C.byref(self.cfg), int(device), C.c_void_p(stream) if stream else None, C.byref(h))
        if rc != RF_OK:
            msg = self.lib.rf_last_error(None)
            raise RuntimeError(f"rf_create failed ({rc}): {msg.decode() if msg else ''}")
        self.h = h
        self.device = device
        self.n = self.lib.rf_cart_size(self.h)

    # -- plumbing ---------------------------------------------------------------
    def close(self):
        if getattr(self, "h", None):
            for p in self._frame_pool:
                self.lib.rf_frame_destroy(self.h, p)
            self._frame_pool = []
            self.lib.rf_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc == RF_OK:
            return
        msg = self.lib.rf_last_error(self.h)
        msg = msg.decode() if msg else ""
        if rc in (RF_E_BADARG, RF_E_CAPACITY):
            raise ValueError(f"libradarfe ({rc}): {msg}")
        raise RuntimeError(f"libradarfe ({rc}): {msg}")

    def launch_count(self) -> int:
        return int(self.lib.rf_launch_count(self.h))

    def timer_start(self):
        self._check(self.lib.rf_timer_start(self.h))

    def timer_stop_ms(self) -> float:
        ms = C.c_float(0)
        self._check(self.lib.rf_timer_stop_ms(self.h, C.byref(ms)))
        return ms.value

    def sync(self):
        self._check(self.lib.rf_sync(self.h))

    def new_frame(self) -> Frame:
        return Frame(self)

    def new_batch(self) -> Batch:
        return Batch(self)

    def new_sequences(self, n_seq, arena_frames, detector_mode=0) -> Sequences:
        return Sequences(self, n_seq, arena_frames, detector_mode)

    # -- a1 ---------------------------------------------------------------------
    def extract_polar(self, raw):
        raw = _c(raw, np.uint8)
        A, W = self.cfg.azimuths, self.cfg.range_bins
        if raw.shape != (A, self.cfg.raw_width):
            raise ValueError(f"raw scan must be {(A, self.cfg.raw_width)}, got {raw.shape}")
        polar = np.empty((A, W), np.float32)
        ts = np.empty((A, 1), np.int64)
        az = np.empty((A, 1), np.float32)
        valid = np.empty((A, 1), np.uint8)
        self._check(self.lib.rf_extract_polar(self.h, _ptr(raw), _ptr(polar), _ptr(ts), _ptr(az), _ptr(valid)))
        return polar, ts, az, valid.astype(bool)

    # -- a2 / a3 ----------------------------------------------------------------
    def polar_to_cart(self, raw=None, polar=None, frame: Frame = None, want_host=True):
        frame = frame or self.new_frame()
        out = _image_pool.empty((self.n, self.n), np.float32) if want_host else None
        if raw is not None:
            raw = _c(raw, np.uint8)
            if raw.shape != (self.cfg.azimuths, self.cfg.raw_width):
                raise ValueError("raw scan has the wrong shape")
            self._check(self.lib.rf_polar_to_cart(self.h, _ptr(raw), None, frame.p, _ptr(out)))
        else:
            polar = _c(polar, np.float32)
            if polar.shape != (self.cfg.azimuths, self.cfg.range_bins):
                raise ValueError(f"polar image must be {(self.cfg.azimuths, self.cfg.range_bins)}, got {polar.shape}")
            self._check(self.lib.rf_polar_to_cart(self.h, None, _ptr(polar), frame.p, _ptr(out)))
        return frame, out

    def polar_to_cart_log(self, polar):
        """convertPolarImageToCartesian(polar, logPolarMode=True) -> f32 [n, n]."""
        polar = _c(polar, np.float32)
        if polar.shape != (self.cfg.azimuths, self.cfg.range_bins):
            raise ValueError(f"polar image must be {(self.cfg.azimuths, self.cfg.range_bins)}, got {polar.shape}")
        out = np.empty((self.n, self.n), np.float32)
        self._check(self.lib.rf_polar_to_cart_log(self.h, _ptr(polar), _ptr(out)))
        return out

    def frame_from_cart(self, cart, frame: Frame = None):
        cart = _c(cart, np.float32)
        if cart.ndim != 2 or cart.shape[0] != cart.shape[1]:
            raise ValueError("Cartesian image must be square")
        frame = frame or self.new_frame()
        self._check(self.lib.rf_frame_from_cart(self.h, _ptr(cart), cart.shape[0], frame.p))
        return frame

    # -- a4 / a5 ----------------------------------------------------------------
    def klt(self, prev: Frame, nxt: Frame, pts, apply_err_gate=True):
        pts = _c(pts, np.float32).reshape(-1, 2)
        K = pts.shape[0]
        out = np.zeros((K, 2), np.float32)
        st = np.zeros((K, 1), np.uint8)
        err = np.zeros((K, 1), np.float32)
        if K:
            self._check(self.lib.rf_klt(self.h, prev.p, nxt.p, _ptr(pts), K, int(apply_err_gate), _ptr(out), _ptr(st),
                                        _ptr(err)))
        return out, st, err

    # -- a6 ---------------------------------------------------------------------
    @staticmethod
    def _coord_dtype(*arrays):
        """float64 if any input holds float64 coordinates (the reference's cdist / SVD then run on them in full precision),
        else float32 (what cv2.calcOpticalFlowPyrLK returns and the batch path carries)"""
        return np.float64 if any(np.asarray(a).dtype == np.float64 for a in arrays) else np.float32

    def reject_outliers(self, prev_xy, new_xy):
        dt = self._coord_dtype(prev_xy, new_xy)
        a = _c(prev_xy, dt).reshape(-1, 2)
        b = _c(new_xy, dt).reshape(-1, 2)
        if a.shape != b.shape:
            raise ValueError("Coordinates should be the same shape")
        K = a.shape[0]
        mask = np.zeros(K, np.uint8)
        n_in, nodes = C.c_int(0), C.c_int(0)
        if K:
            fn = self.lib.rf_reject_outliers_f64 if dt == np.float64 else self.lib.rf_reject_outliers
            rc = fn(self.h, _ptr(a), _ptr(b), K, _ptr(mask), C.byref(n_in), C.byref(nodes))
            if rc == RF_E_WORKLIMIT:
                # the reference always completes (in seconds to minutes on such graphs); keep the odometry loop alive with
                # the largest clique found within cfg.clique_node_limit search nodes
                import warnings
                warnings.warn(f"rejectOutliers: clique search stopped after {nodes.value} nodes; using the best clique so far "
                              f"({n_in.value} of {K} points)", RuntimeWarning, stacklevel=3)
            else:
                self._check(rc)
        return mask.astype(bool), n_in.value, nodes.value

    def consistency_adjacency(self, prev_xy, new_xy):
        dt = self._coord_dtype(prev_xy, new_xy)
        a = _c(prev_xy, dt).reshape(-1, 2)
        b = _c(new_xy, dt).reshape(-1, 2)
        K = a.shape[0]
        adj = np.zeros((K, K), np.uint8)
        fn = self.lib.rf_consistency_adjacency_f64 if dt == np.float64 else self.lib.rf_consistency_adjacency
        self._check(fn(self.h, _ptr(a), _ptr(b), K, _ptr(adj)))
        return adj

    def clique_search(self, adj, prune=True):
        """Test hook: (mask, size, n_yields, order_hash, nodes) for a K x K adjacency matrix."""
        adj = _c(adj, np.uint8)
        K = adj.shape[0]
        mask = np.zeros(max(K, 1), np.int32)
        size, ny, hs, nodes = C.c_int(0), C.c_int64(0), C.c_uint64(0), C.c_int64(0)
        self._check(self.lib.rf_clique_search(self.h, _ptr(adj), K, int(prune), _ptr(mask), C.byref(size), C.byref(ny),
                                              C.byref(hs), C.byref(nodes)))
        return mask[:K].astype(bool), size.value, ny.value, hs.value, nodes.value

    # -- a7 ---------------------------------------------------------------------
    def kabsch(self, src_xy, tgt_xy):
        dt = self._coord_dtype(src_xy, tgt_xy)
        s = _c(src_xy, dt).reshape(-1, 2)
        t = _c(tgt_xy, dt).reshape(-1, 2)
        if s.shape != t.shape:
            raise ValueError("point sets must have the same shape")
        R = np.zeros((2, 2), np.float64)
        h = np.zeros((2, 1), np.float64)
        fn = self.lib.rf_kabsch_f64 if dt == np.float64 else self.lib.rf_kabsch
        self._check(fn(self.h, _ptr(s), _ptr(t), s.shape[0], _ptr(R), _ptr(h)))
        return R, h

    # -- a8 ---------------------------------------------------------------------
    def mds_solve(self, T_wj0, p_w, p_jt, T_wj, sigma_p=None, sigma_v=None, period=0.0):
        T0 = _c(T_wj0, np.float64).reshape(3, 3)
        T1 = _c(T_wj, np.float64).reshape(3, 3)
        pw = _c(np.asarray(p_w)[:, :2], np.float64)
        pj = _c(np.asarray(p_jt)[:, :2], np.float64)
        if pw.shape != pj.shape:
            raise ValueError("p_w and p_jt must have the same shape")
        sp = None if sigma_p is None else _c(sigma_p, np.float64).reshape(2)
        sv = None if sigma_v is None else _c(sigma_v, np.float64).reshape(3)
        x = np.zeros(6, np.float64)
        it = C.c_int(0)
        cost = C.c_double(0)
        self._check(self.lib.rf_mds_solve(self.h, _ptr(T0), _ptr(pw), _ptr(pj), pw.shape[0], _ptr(T1), _ptr(sp), _ptr(sv),
                                          C.c_double(period), _ptr(x), C.byref(it), C.byref(cost)))
        return x, it.value, cost.value

    def mds_undistort(self, v, pts_xy, period, times=None):
        v = _c(v, np.float64).reshape(3)
        p = _c(np.asarray(pts_xy)[:, :2], np.float64)
        out = np.zeros_like(p)
        if p.shape[0] and times is not None:
            t = _c(np.broadcast_to(np.asarray(times, np.float64), (p.shape[0],)), np.float64)
            self._check(self.lib.rf_mds_undistort_times(self.h, _ptr(v), _ptr(p), _ptr(t), p.shape[0], _ptr(out)))
        elif p.shape[0]:
            self._check(self.lib.rf_mds_undistort(self.h, _ptr(v), _ptr(p), p.shape[0], C.c_double(period), _ptr(out)))
        return out

    # -- a9 ---------------------------------------------------------------------
    def ssc(self, keypoints, num_ret_points, tolerance, cols, rows):
        kp = _c(keypoints, np.float64).reshape(-1, 3)
        n = kp.shape[0]
        sel = np.zeros(max(n, 1), np.int32)
        m = C.c_int(0)
        self._check(self.lib.rf_ssc(self.h, _ptr(kp), n, int(num_ret_points), C.c_double(tolerance), int(cols), int(rows),
                                    _ptr(sel), C.byref(m)))
        return sel[:m.value].copy()

    # -- a10 --------------------------------------------------------------------
    def detect(self, frame: Frame, threshold, cap=65536, mode=0):
        out = np.zeros((cap, 3), np.float64)
        n = C.c_int(0)
        self._check(self.lib.rf_detect(self.h, frame.p, mode, C.c_float(threshold), _ptr(out), cap, C.byref(n)))
        return out[:min(n.value, cap)].copy(), n.value

    def detect_doh(self, frame: Frame, min_sigma=1, max_sigma=30, num_sigma=10, threshold=0.01, overlap=0.5, cap=8192):
        """skimage.feature.blob_doh on the frame's f32 Cartesian image -> f64 [K, 3] (row, col, sigma)."""
        out = np.zeros((cap, 3), np.float64)
        n = C.c_int(0)
        self._check(self.lib.rf_detect_doh(self.h, frame.p, C.c_double(min_sigma), C.c_double(max_sigma), int(num_sigma),
                                           C.c_double(threshold), C.c_double(overlap), _ptr(out), cap, C.byref(n)))
        return out[:min(n.value, cap)].copy()

    def doh_response(self, frame: Frame, min_sigma, max_sigma, num_sigma, sigma_index):
        """Test hook: float64 integral image (sigma_index < 0) or one Hessian-determinant plane."""
        out = np.zeros((self.n, self.n), np.float64)
        self._check(self.lib.rf_doh_response(self.h, frame.p, C.c_double(min_sigma), C.c_double(max_sigma), int(num_sigma),
                                             int(sigma_index), _ptr(out)))
        return out

    def nms_select(self, resp, threshold, cap=None):
        resp = _c(resp, np.float32)
        rows, cols = resp.shape
        cap = rows * cols if cap is None else cap
        out = np.zeros((max(cap, 1), 3), np.float64)
        n = C.c_int(0)
        self._check(self.lib.rf_nms_select(self.h, _ptr(resp), rows, cols, C.c_float(threshold), _ptr(out), cap, C.byref(n)))
        return out[:min(n.value, cap)].copy(), n.value

    def corner_response(self, frame: Frame, mode=0):
        out = np.zeros((self.n, self.n), np.float32)
        self._check(self.lib.rf_corner_response(self.h, frame.p, mode, _ptr(out)))
        return out

    # -- a12 --------------------------------------------------------------------
    def polar_peaks(self, polar):
        polar = _c(polar, np.float32)
        A, W = polar.shape
        cap = A * (W // 2 + 1)
        out = np.empty((cap, 2), np.int64)
        n = C.c_int64(0)
        self._check(self.lib.rf_polar_peaks(self.h, _ptr(polar), A, W, _ptr(out), C.c_int64(cap), C.byref(n)))
        return out[:n.value].copy()


    # -- a11 one pair (rf_track_pair) ----------------------------------------------
    def track_pair(self, raw_prev, raw_next, feats_xy, prev_pose=None, with_mds=False):
        """Two raw scans + features on the first -> (rf_pair_result record, next_xy [K, 2], corrStatus u8 [K])."""
        a, b = _c(raw_prev, np.uint8), _c(raw_next, np.uint8)
        want = (self.cfg.azimuths, self.cfg.raw_width)
        if a.shape != want or b.shape != want:
            raise ValueError(f"raw scans must be {want}")
        f = _c(np.asarray(feats_xy, np.float32).reshape(-1, 2), np.float32)
        K = f.shape[0]
        pose = None if prev_pose is None else _c(np.asarray(prev_pose, np.float64).reshape(3), np.float64)
        res = np.zeros(1, PAIR_RESULT_DTYPE)
        nxt, st = np.zeros((K, 2), np.float32), np.zeros(K, np.uint8)
        self._check(self.lib.rf_track_pair(self.h, _ptr(a), _ptr(b), _ptr(f), K, _ptr(pose), int(with_mds), _ptr(res),
                                           _ptr(nxt), _ptr(st)))
        return res[0], nxt, st

    # -- N1 FMT rotation prior (FMT.py:13-90) ---------------------------------------
    def fmt_rotation(self, polar, pairs=((0, 1),), downsample=10, clip_px=0):
        """polar [F, A, W] f32, pairs [P, 2] -> (angle_rad [P], scale [P], response [P], shift_xy [P, 2])."""
        pairs = _c(np.asarray(pairs, np.int32).reshape(-1, 2), np.int32)
        P = pairs.shape[0]
        ang, sc, resp, sh = np.zeros(P), np.zeros(P), np.zeros(P), np.zeros((P, 2))
        if isinstance(polar, (list, tuple)):          # separately allocated images: one pointer each, no stacking
            frames = [_c(f, np.float32) for f in polar]
            if not frames or any(f.ndim != 2 or f.shape != frames[0].shape for f in frames):
                raise ValueError("fmt_rotation: the images need the same 2-D shape")
            A, W = frames[0].shape
            ptrs = (C.c_void_p * len(frames))(*[f.ctypes.data for f in frames])
            self._check(self.lib.rf_fmt_rotation_frames(self.h, ptrs, len(frames), A, W, _ptr(pairs), P, int(downsample),
                                                        int(clip_px), _ptr(ang), _ptr(sc), _ptr(resp), _ptr(sh)))
            return ang, sc, resp, sh
        polar = _c(polar, np.float32)
        if polar.ndim != 3:
            raise ValueError(f"expected [frames, azimuths, bins], got {polar.shape}")
        F, A, W = polar.shape
        self._check(self.lib.rf_fmt_rotation(self.h, _ptr(polar), F, A, W, _ptr(pairs), P, int(downsample), int(clip_px),
                                             _ptr(ang), _ptr(sc), _ptr(resp), _ptr(sh)))
        return ang, sc, resp, sh

    def fmt_log_polar(self, polar, downsample=10, clip_px=0):
        """parseData.convertPolarImgToLogPolar(cv2.resize(polar[:, :clip_px], ...)) -> f32 [h_lp, w_lp]."""
        polar = _c(polar, np.float32)
        A, W = polar.shape
        h_lp, w_lp = C.c_int(0), C.c_int(0)
        self._check(self.lib.rf_fmt_log_polar(self.h, _ptr(polar), A, W, int(downsample), int(clip_px), None, C.c_int64(0),
                                              C.byref(h_lp), C.byref(w_lp)))
        out = np.empty((h_lp.value, w_lp.value), np.float32)
        self._check(self.lib.rf_fmt_log_polar(self.h, _ptr(polar), A, W, int(downsample), int(clip_px), _ptr(out),
                                              C.c_int64(out.size), C.byref(h_lp), C.byref(w_lp)))
        return out

    def cart_to_polar(self, cart, log_mode=False, shape_hw=None):
        """cv2.warpPolar forward map of a square f32 image (parseData.convertCartesianImageToPolar) -> f32 [rows, cols]."""
        cart = _c(cart, np.float32)
        if cart.ndim != 2 or cart.shape[0] != cart.shape[1]:
            raise AssertionError("Should be a square Cartesian image")
        ro, co = (0, 0) if shape_hw is None else (int(shape_hw[0]), int(shape_hw[1]))
        r, c = C.c_int(0), C.c_int(0)
        self._check(self.lib.rf_cart_to_polar(self.h, _ptr(cart), cart.shape[0], int(bool(log_mode)), ro, co, None, C.c_int64(0),
                                              C.byref(r), C.byref(c)))
        out = np.empty((r.value, c.value), np.float32)
        self._check(self.lib.rf_cart_to_polar(self.h, _ptr(cart), cart.shape[0], int(bool(log_mode)), ro, co, _ptr(out),
                                              C.c_int64(out.size), C.byref(r), C.byref(c)))
        return out

    def phase_correlate(self, a, b):
        """cv2.phaseCorrelate(a, b, hanning) -> ((dx, dy), response)."""
        a, b = _c(a, np.float32), _c(b, np.float32)
        if a.shape != b.shape or a.ndim != 2:
            raise ValueError("phase_correlate: images need the same 2-D shape")
        dx, dy, r = C.c_double(0), C.c_double(0), C.c_double(0)
        self._check(self.lib.rf_phase_correlate(self.h, _ptr(a), _ptr(b), a.shape[0], a.shape[1], C.byref(dx), C.byref(dy),
                                                C.byref(r)))
        return (dx.value, dy.value), r.value


    # -- N3 trajectory chaining (trajectoryPlotting.py:27-60) -------------------------
    def chain_poses(self, R, h, start_pose=None, left_multiply=True):
        """R [P, 2, 2] (or [P, 4]), h [P, 2] -> poses [P + 1, 3] (x, y, theta)."""
        R = _c(np.asarray(R, np.float64).reshape(-1, 4), np.float64)
        h = _c(np.asarray(h, np.float64).reshape(-1, 2), np.float64)
        if R.shape[0] != h.shape[0]:
            raise ValueError("chain_poses: R and h need the same length")
        P = R.shape[0]
        sp = None if start_pose is None else _c(np.asarray(start_pose, np.float64).reshape(3), np.float64)
        out = np.empty((P + 1, 3), np.float64)
        self._check(self.lib.rf_chain_poses(self.h, _ptr(R), _ptr(h), P, _ptr(sp), int(bool(left_multiply)), _ptr(out)))
        return out


_default = {}
_default_lock = threading.Lock()


def default_engine(device: int = 0) -> RadarFE:
    """Process-wide engine used by the drop-in modules (one per device)."""
    with _default_lock:
        fe = _default.get(device)
        if fe is None:
            fe = RadarFE(device=device)
            _default[device] = fe
        return fe
