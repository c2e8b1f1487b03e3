"""Drop-in for the reference's FMT.py: Fourier-Mellin rotation prior on libradarfe.so.
Same names, arguments and return values as FMT.py:13-90; the arithmetic runs in
rf_fmt_rotation / rf_phase_correlate (csrc/k_fmt.cu).  The plotting helpers (FMT.py:103-165) import matplotlib lazily;
rotateImg (cv2.warpAffine, only used by those plots) is not provided."""
from typing import Tuple

import numpy as np

from . import _engine
from .parseData import RANGE_RESOLUTION_CART_M

FMT_DOWNSAMPLE_FACTOR = 10   # FMT.py:10
FMT_RANGE_CLIP_M = 87.5      # FMT.py:11


def getTranslationUsingPhaseCorrelation(srcImg: np.ndarray, targetImg: np.ndarray) -> Tuple[Tuple[float, float], float]:
    """FMT.py:13-33: Hann-windowed cv2.phaseCorrelate -> ((dx, dy), response)."""
    fe = _engine.engine()
    return fe.phase_correlate(srcImg, targetImg)


def getRotationUsingFMT(srcPolarImg: np.ndarray, targetPolarImg: np.ndarray, downsampleFactor: int = FMT_DOWNSAMPLE_FACTOR,
                        maxRangeClipM=FMT_RANGE_CLIP_M) -> Tuple[float, float, float]:
    """FMT.py:36-90 -> (angleRad with R(angleRad) @ src = target, scale, response)."""
    assert srcPolarImg.shape == targetPolarImg.shape, "Images need to have the same shape!"
    clip_px = int(maxRangeClipM / RANGE_RESOLUTION_CART_M) if maxRangeClipM > 0 else 0   # FMT.py:55-58
    fe = _engine.engine()
    ang, sc, resp, _ = fe.fmt_rotation([srcPolarImg, targetPolarImg], ((0, 1),), downsample=int(downsampleFactor), clip_px=clip_px)
    return float(ang[0]), float(sc[0]), float(resp[0])


def plotCartPolar(prevImgPolar, currImgPolar, prevImgCart, currImgCart):
    """FMT.py:103-133: 2 x 2 panel of the polar / Cartesian images."""
    from matplotlib import pyplot as plt
    panels = ((prevImgPolar, "Prev Image Polar"), (currImgPolar, "Curr Image Polar"),
              (prevImgCart, "Prev Image Cartesian"), (currImgCart, "Curr Image Cartesian"))
    for i, (img, title) in enumerate(panels, start=1):
        plt.subplot(2, 2, i)
        if img is not None:
            plt.imshow(img)
            plt.title(title)


def plotCartPolarWithRotation(prevImgCart, currImgCart, rotRad):
    """FMT.py:136-165 without the rotation-corrected panels (they need rotateImg)."""
    from matplotlib import pyplot as plt
    for i, (img, title) in enumerate(((prevImgCart, "Prev Image"), (currImgCart, "Curr Image")), start=1):
        plt.subplot(1, 2, i)
        if img is not None:
            plt.imshow(img)
            plt.axis("off")
            plt.title(title)
