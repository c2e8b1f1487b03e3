"""Engine cache and device-backed images shared by the drop-in modules.

The reference's functions take and return plain NumPy arrays.  The drop-in
keeps that, and additionally remembers which device frame (f32 Cartesian image + u8 LK
pyramid, an rf_frame) an image it returned lives in, so a later getTrackedPointsKLT call on the
same array does not upload and rebuild the pyramid again.  One rf_handle exists per scan
geometry (azimuths, range bins, downsample factor) and device."""
import threading

import numpy as np

from . import _ffi

_engines = {}
_lock = threading.Lock()
_device = 0


def set_device(device: int):
    """Select the CUDA device the drop-in modules run on (default 0)."""
    global _device
    _device = int(device)


def engine(range_bins: int = None, azimuths: int = None, downsample: int = None, raw_width: int = None) -> _ffi.RadarFE:
    """rf_handle for a scan geometry; created on first use.  Raises if there is no GPU."""
    cfg = _ffi.default_config()
    if range_bins is not None:
        cfg.range_bins = int(range_bins)
    if azimuths is not None:
        cfg.azimuths = int(azimuths)
    if downsample is not None:
        cfg.downsample = int(downsample)
    if raw_width is not None:
        cfg.raw_width = int(raw_width)
    cfg.raw_width = max(cfg.raw_width, cfg.meta_bytes + cfg.range_bins)
    key = (_device, cfg.azimuths, cfg.range_bins, cfg.downsample, cfg.raw_width)
    with _lock:
        fe = _engines.get(key)
        if fe is None:
            fe = _ffi.RadarFE(cfg, device=_device)
            _engines[key] = fe
        return fe


def engine_for_cart(n: int) -> _ffi.RadarFE:
    """Engine whose Cartesian size is n x n (any cached one, else the geometry 2R = n, W = n + 1)."""
    with _lock:
        for fe in _engines.values():
            if fe.n == n and fe.device == _device:
                return fe
    if n % 2:
        raise ValueError(f"Cartesian images must have an even size, got {n}")
    return engine(range_bins=n + 1)


class DeviceImage(np.ndarray):
    """ndarray that remembers the device frame it was downloaded from.  Read-only: the device copy
    and the host copy must not diverge (np.array(img) gives a private writable copy)."""
    _rf_frame = None
    _rf_engine = None

    def __array_finalize__(self, obj):
        # views, slices and arithmetic results are NOT the device image any more
        self._rf_frame = None
        self._rf_engine = None


def wrap(host: np.ndarray, fe, frame) -> DeviceImage:
    out = host.view(DeviceImage)
    out._rf_frame, out._rf_engine = frame, fe
    out.flags.writeable = False
    return out


def frame_of(img, fe=None):
    """(engine, rf_frame) for an image: reuse the device frame it came from, else upload it."""
    f = getattr(img, "_rf_frame", None)
    if f is not None and f.p:
        return img._rf_engine, f
    img = np.asarray(img)
    if img.ndim != 2 or img.shape[0] != img.shape[1]:
        raise ValueError(f"expected a square Cartesian image, got {img.shape}")
    fe = fe or engine_for_cart(img.shape[0])
    return fe, fe.frame_from_cart(img)
