"""TEST INFRASTRUCTURE ONLY — generates tests/golden/fmt_fixture.npz by running the UNMODIFIED
reference's FMT.getRotationUsingFMT / parseData.convertPolarImgToLogPolar (imported from
/root/reference through oracle/shims) on the data/tiny scans stored in tiny_frames.npz.

    python -m oracle.gen_golden_fmt

Recorded (all produced by the reference's own functions):
  fmt_angle, fmt_scale, fmt_response   getRotationUsingFMT(polar_i, polar_j) for the pairs in `pairs`
  fmt_shift                            the cv2.phaseCorrelate deltas of the same pairs (dx, dy)
  logpolar_0                           convertPolarImgToLogPolar(cv2.resize(polar_0[:, :1012], (101, 400)))
  roll_angle                           getRotationUsingFMT(polar_0, np.roll(polar_0, k, axis=0)) for k in roll_k
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")


def main():
    sys.path.insert(0, ROOT)
    from oracle import ref_import as ri
    ri.activate()
    import cv2
    with ri.quiet():
        import FMT
        import parseData
    fr = np.load(os.path.join(GOLD, "tiny_frames.npz"))
    polar = [parseData.extractDataFromRadarImage(fr[f"raw_{i}"])[0] for i in range(3)]
    pairs = np.array([[0, 1], [1, 2], [0, 2], [1, 0]], np.int32)
    ang, sc, resp, shift = [], [], [], []
    for i, j in pairs:
        a, s, r = FMT.getRotationUsingFMT(polar[i], polar[j])
        ang.append(a); sc.append(s); resp.append(r)
        clip = int(FMT.FMT_RANGE_CLIP_M / parseData.RANGE_RESOLUTION_CART_M)
        lp = [parseData.convertPolarImgToLogPolar(cv2.resize(polar[k][:, :clip], (clip // FMT.FMT_DOWNSAMPLE_FACTOR, polar[k].shape[0])))
              for k in (i, j)]
        d, _ = FMT.getTranslationUsingPhaseCorrelation(lp[0], lp[1])
        shift.append(d)
    clip = int(FMT.FMT_RANGE_CLIP_M / parseData.RANGE_RESOLUTION_CART_M)
    lp0 = parseData.convertPolarImgToLogPolar(cv2.resize(polar[0][:, :clip], (clip // FMT.FMT_DOWNSAMPLE_FACTOR, polar[0].shape[0])))
    roll_k = np.array([1, 3, 10, -7, 40], np.int32)
    roll_angle = [FMT.getRotationUsingFMT(polar[0], np.roll(polar[0], int(k), axis=0))[0] for k in roll_k]
    np.savez_compressed(os.path.join(GOLD, "fmt_fixture.npz"), pairs=pairs, fmt_angle=np.array(ang, np.float64),
                        fmt_scale=np.array(sc, np.float64), fmt_response=np.array(resp, np.float64),
                        fmt_shift=np.array(shift, np.float64), logpolar_0=lp0.astype(np.float32), roll_k=roll_k,
                        roll_angle=np.array(roll_angle, np.float64))
    print("fmt_fixture.npz:", dict(angle=ang, scale=sc, response=resp, roll=roll_angle))


if __name__ == "__main__":
    main()
