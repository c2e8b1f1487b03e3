"""radarslampy_b200 — B200-native (sm_100a) drop-in for the per-frame radar-odometry front
end of Samleo8/RadarSLAMPy.  The modules parseData / getFeatures / ANMS / getTransformKLT /
outlierRejection / motionDistortion / getPointCloud / Tracker / FMT / Mapping / trajectoryPlotting keep the
reference's names and call signatures and run on hand-written CUDA kernels behind the C ABI in
include/radarfe.h.  There is no CPU fallback.

    import radarslampy_b200
    radarslampy_b200.install()            # the reference's bare module names now resolve to this package
    from RawROAMSystem import RawROAMSystem   # unmodified reference driver (its directory on sys.path)
"""
import importlib
import sys

__version__ = "0.2.0"

# the reference's flat module names (RawROAMSystem.py:6-13, Tracker.py:8-12, Mapping.py:5-10) this package replaces
HOT_PATH_MODULES = ("parseData", "getFeatures", "ANMS", "getTransformKLT", "outlierRejection", "motionDistortion",
                    "getPointCloud", "FMT", "Tracker", "Mapping")
# numeric half of trajectoryPlotting.py; not installed by default because the reference driver also imports its
# plotting function (plotGtAndEstTrajectory) from the same module
OPTIONAL_MODULES = ("trajectoryPlotting",)


def install(names=None, device: int = None):
    """Register the drop-in modules under the reference's bare module names in sys.modules, so that the unmodified
    reference files (`from parseData import ...`, `from Tracker import Tracker`) bind to this package.  Call it
    BEFORE importing any reference module.  Returns the list of names installed."""
    if device is not None:
        from . import _engine
        _engine.set_device(device)
    names = tuple(names) if names is not None else HOT_PATH_MODULES
    done = []
    for n in names:
        if n not in HOT_PATH_MODULES + OPTIONAL_MODULES:
            raise ValueError(f"{n!r} is not a module this package replaces")
        loaded = sys.modules.get(n)
        mod = importlib.import_module(f"{__name__}.{n}")
        if loaded is not None and loaded is not mod:
            raise RuntimeError(f"module {n!r} was already imported from {getattr(loaded, '__file__', '?')}: "
                               "call radarslampy_b200.install() before importing the reference")
        sys.modules[n] = mod
        done.append(n)
    return done


def uninstall():
    """Remove the aliases install() created."""
    for n in HOT_PATH_MODULES + OPTIONAL_MODULES:
        m = sys.modules.get(n)
        if m is not None and getattr(m, "__name__", "").startswith(__name__ + "."):
            del sys.modules[n]
