"""Drop-in for the reference's getPointCloud.py: per-azimuth peak extraction with mean + std
pruning (rf_polar_peaks, csrc/k_features.cu), exact against getPointCloud.py:11-54."""
import numpy as np

from . import _engine


def getPointCloudPolarInd(polarImage: np.ndarray, peakDistance: float = None, peakProminence: float = None) -> np.ndarray:
    """-> int [K, 2] rows [azimuth index, range index], azimuth-major."""
    if peakDistance is not None or peakProminence is not None:
        raise NotImplementedError("the reference never passes peakDistance/peakProminence (Mapping.py:62)")
    polar = np.ascontiguousarray(polarImage, dtype=np.float32)
    return _engine.engine().polar_peaks(polar).astype(int)
