"""Drop-in for the reference's parseData.py (scan decode and polar -> Cartesian conversion) on
libradarfe.so.  Same names, arguments and return values as parseData.py:17-53,100-135,160-226;
the arithmetic runs in rf_extract_polar / rf_polar_to_cart (csrc/k_image.cu)."""
import os
from typing import List, Tuple

import numpy as np

from . import _engine

RANGE_RESOLUTION_M = 0.0432            # parseData.py:9
DOWNSAMPLE_FACTOR = 2                  # parseData.py:10
RANGE_RESOLUTION_CART_M = RANGE_RESOLUTION_M * DOWNSAMPLE_FACTOR
MAX_RANGE_CLIP_DEFAULT = 87.5          # parseData.py:14
META_BYTES = 11


def extractDataFromRadarImage(polarImgData: np.ndarray, maxRangeClipM: float = MAX_RANGE_CLIP_DEFAULT
                              ) -> Tuple[np.ndarray, np.ndarray, float, float, np.ndarray, np.ndarray]:
    """parseData.py:17-53 -> (range_azimuth_data f32 [A, W], azimuths f32 [A,1], range_resolution,
    azimuth_resolution, valid bool [A,1], timestamps i64 [A,1])."""
    raw = np.ascontiguousarray(polarImgData, dtype=np.uint8)
    A, wtot = raw.shape
    bins = wtot - META_BYTES
    if maxRangeClipM > 0:
        bins = min(bins, int(maxRangeClipM / RANGE_RESOLUTION_M))
    fe = _engine.engine(range_bins=bins, azimuths=A, raw_width=wtot)
    polar, timestamps, azimuths, valid = fe.extract_polar(raw)
    azimuth_resolution = azimuths[1] - azimuths[0]
    return polar, azimuths, RANGE_RESOLUTION_M, azimuth_resolution, valid, timestamps


def convertPolarImageToCartesian(imgPolar: np.ndarray, logPolarMode: bool = False,
                                 downsampleFactor: int = DOWNSAMPLE_FACTOR,
                                 changeGlobalRangeResolution: bool = False) -> np.ndarray:
    """parseData.py:100-135 (cv2.warpPolar, inverse linear map) -> f32 [2R, 2R].  The returned array
    also keeps its device frame (u8 image + LK pyramid) for getTrackedPointsKLT."""
    imgPolar = np.ascontiguousarray(imgPolar, dtype=np.float32)
    A, W = imgPolar.shape
    if changeGlobalRangeResolution:
        global RANGE_RESOLUTION_CART_M
        RANGE_RESOLUTION_CART_M = RANGE_RESOLUTION_M * downsampleFactor
    fe = _engine.engine(range_bins=W, azimuths=A, downsample=max(int(downsampleFactor), 1))
    if logPolarMode:
        # parseData.py:131-133 (flags += cv2.WARP_POLAR_LOG): never taken on the reference's live path; a plain array
        # (no device frame), equal to cv2's within the tolerance of its f32 log
        return fe.polar_to_cart_log(imgPolar)
    frame, cart = fe.polar_to_cart(polar=imgPolar)
    return _engine.wrap(cart, fe, frame)


def convertCartesianImageToPolar(imgCart: np.ndarray, logPolarMode: bool = False, shapeHW: Tuple[int, int] = None) -> np.ndarray:
    """parseData.py:69-97 (cv2.warpPolar forward map, linear or semi-log) -> f32 [round(pi n / 2), round(n / 2)] or shapeHW."""
    imgCart = np.ascontiguousarray(imgCart, dtype=np.float32)
    h, w = imgCart.shape
    assert w == h, "Should be a square Cartesian image"
    return _engine.engine().cart_to_polar(imgCart, log_mode=logPolarMode, shape_hw=shapeHW)


def drawCVPoint(img: np.ndarray, point, point_color: Tuple[int, int, int] = (0, 0, 255)):
    """parseData.py:56-66: cv2.circle(img, point, radius=0, thickness=-1) paints exactly the pixel (x, y), in place."""
    if hasattr(point, "asTuple"):
        point = point.asTuple()
    x, y = int(point[0]), int(point[1])
    if 0 <= y < img.shape[0] and 0 <= x < img.shape[1]:
        img[y, x] = point_color if img.ndim == 3 else point_color[0]
    return img


def convertPolarImgToLogPolar(imgPolar: np.ndarray) -> np.ndarray:
    """parseData.py:138-157: polar -> Cartesian (no down-sampling) -> semi-log polar, f32 [round(pi W), W]."""
    imgPolar = np.ascontiguousarray(imgPolar, dtype=np.float32)
    return _engine.engine().fmt_log_polar(imgPolar, downsample=1, clip_px=0)


def convertRawScanToCartesian(polarImgData: np.ndarray, maxRangeClipM: float = MAX_RANGE_CLIP_DEFAULT) -> np.ndarray:
    """Fused extractDataFromRadarImage + convertPolarImageToCartesian straight from the raw u8 scan
    (no f32 polar image on the host); bit-identical to calling the two in sequence."""
    raw = np.ascontiguousarray(polarImgData, dtype=np.uint8)
    A, wtot = raw.shape
    bins = wtot - META_BYTES
    if maxRangeClipM > 0:
        bins = min(bins, int(maxRangeClipM / RANGE_RESOLUTION_M))
    fe = _engine.engine(range_bins=bins, azimuths=A, raw_width=wtot)
    frame, cart = fe.polar_to_cart(raw=raw)
    return _engine.wrap(cart, fe, frame)


# ---- file access (parseData.py:160-226): PNG decode on the host cores (csrc/ingest.cu) -----------
def readRadarScans(imgPathArr: List[str], threads: int = 0, pinned: bool = False) -> np.ndarray:
    """All the scans of a sequence at once, decoded in parallel on the host cores (rf_ingest_png): u8 [n, A, 3779]."""
    from . import _ffi
    return _ffi.ingest_png(imgPathArr, threads=threads, pinned=pinned)


def getDataFromImgPathsByIndex(imgPathArr: List[str], index: int):
    """parseData.py:160-179 (cv2.imread(..., IMREAD_GRAYSCALE) replaced by the library's own decoder)."""
    from . import _ffi
    if not os.path.exists(imgPathArr[index]):
        raise FileNotFoundError(imgPathArr[index])
    return extractDataFromRadarImage(_ffi.ingest_png([imgPathArr[index]], threads=1)[0])


def getPolarImageFromImgPaths(imgPathArr: List[str], index: int) -> np.ndarray:
    return getDataFromImgPathsByIndex(imgPathArr, index)[0]


def getCartImageFromImgPaths(imgPathArr: List[str], index: int) -> np.ndarray:
    return convertPolarImageToCartesian(getPolarImageFromImgPaths(imgPathArr, index))


def getRadarImgPaths(dataPath: str, timestampPath: str) -> List[str]:
    imgPathArr = []
    with open(timestampPath, "r") as f:
        for line in f:
            stamp, valid = line.strip().split(" ")
            if valid:                                  # parseData.py:221 (a non-empty string, so "0" passes too)
                imgPathArr.append(os.path.join(dataPath, stamp + ".png"))
    return imgPathArr


def getRadarStreamPolar(dataPath: str, timestampPath: str) -> np.ndarray:
    """parseData.py:229-259 -> f32 [A, W, N] stack of the polar images of a sequence (scans decoded in parallel)."""
    imgPathArr = getRadarImgPaths(dataPath, timestampPath)
    raws = readRadarScans(imgPathArr)
    streamArr = None
    for i in range(len(imgPathArr)):
        imgPolar = extractDataFromRadarImage(raws[i])[0]
        if streamArr is None:
            streamArr = np.empty(imgPolar.shape + (len(imgPathArr),), dtype=imgPolar.dtype)
        streamArr[:, :, i] = imgPolar
    return streamArr
