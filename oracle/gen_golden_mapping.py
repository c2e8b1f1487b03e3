"""TEST INFRASTRUCTURE ONLY — generates tests/golden/mapping_fixture.npz by running the UNMODIFIED reference's
Mapping.Keyframe / Mapping.Map (imported from /root/reference through oracle/shims) on data/tiny frame 0 and the
feature points recorded in tiny_stages.npz.

    python -m oracle.gen_golden_mapping
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
    with ri.quiet():
        import Mapping as M
        import parseData
    fr = np.load(os.path.join(GOLD, "tiny_frames.npz"))
    st = np.load(os.path.join(GOLD, "tiny_stages.npz"))
    polar = parseData.extractDataFromRadarImage(fr["raw_0"])[0]
    feats_px = st["feat_in_0"].astype(np.float64)
    centered = (feats_px - 1012.0) * parseData.RANGE_RESOLUTION_CART_M          # RawROAMSystem.py:197-198
    pose = np.array([12.5, -3.0, 0.4])
    vel = np.array([9.0, 0.5, 0.08])
    with ri.quiet():
        kf = M.Keyframe(pose, centered, polar, vel)
    corr = (np.arange(len(centered)) % 3 != 0).astype(np.uint8).reshape(-1, 1)
    und = kf.featurePointsLocalUndistorted.copy()
    glob_all = kf.getPrunedFeaturesGlobalPosition().copy()
    kf.pruneFeaturePoints(corr)
    glob_pruned = kf.getPrunedFeaturesGlobalPosition().copy()
    l2g = kf.convertFeaturesLocalToGlobal(feats_px.copy())
    mp = M.Map("tiny", None, ["a"], {})
    mp.addKeyframe(kf)
    cands = np.array([[12.5, -3.0, 0.4], [12.5, -3.0, 0.61], [14.0, -2.0, 0.4], [14.5, -3.1, 0.45], [13.9, -3.0, 0.59]])
    good = []
    for c in cands:
        with ri.quiet():
            k2 = M.Keyframe(c, centered[:5], polar, vel)
        good.append(mp.isGoodKeyframe(k2))
    np.savez_compressed(os.path.join(GOLD, "mapping_fixture.npz"), pose=pose, velocity=vel, centered=centered, corr=corr,
                        undistorted=und, global_all=glob_all, global_pruned=glob_pruned, local_to_global=l2g,
                        center=np.asarray(M.RADAR_CART_CENTER, np.float64), point_cloud=np.asarray(kf.pointCloud, np.int64),
                        cands=cands, good=np.array(good, bool))
    print("mapping_fixture.npz:", und.shape, glob_pruned.shape, kf.pointCloud.shape, good)


if __name__ == "__main__":
    main()
