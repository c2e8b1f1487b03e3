"""GPU: the Mapping drop-in (keyframe bookkeeping) against goldens recorded from the unmodified
reference's Mapping.Keyframe / Map on data/tiny (tests/golden/mapping_fixture.npz, oracle/gen_golden_mapping.py)."""
import numpy as np
import pytest

from oracle import restate as R

pytestmark = pytest.mark.gpu


def test_keyframe_and_map(golden):
    from radarslampy_b200 import Mapping as M
    g = golden["mapping_fixture"]
    polar = R.extract_polar(golden["tiny_frames"]["raw_0"], 2025)
    kf = M.Keyframe(g["pose"], g["centered"], polar, g["velocity"])
    assert np.array_equal(M.RADAR_CART_CENTER, g["center"])
    assert np.array_equal(kf.pointCloud, g["point_cloud"])                         # bit-exact peak extraction
    assert np.abs(kf.featurePointsLocalUndistorted - g["undistorted"]).max() <= 1e-9
    assert np.abs(kf.getPrunedFeaturesGlobalPosition() - g["global_all"]).max() <= 1e-9
    kf.pruneFeaturePoints(g["corr"])
    assert kf.prunedFeaturePoints.shape[0] == int(g["corr"].sum())
    assert np.abs(kf.getPrunedFeaturesGlobalPosition() - g["global_pruned"]).max() <= 1e-9
    feats_px = g["centered"] / 0.0864 + 1012.0
    assert np.abs(kf.convertFeaturesLocalToGlobal(feats_px) - g["local_to_global"]).max() <= 1e-9
    mp = M.Map("tiny", None, ["a"], {})
    mp.addKeyframe(kf)
    good = [mp.isGoodKeyframe(M.Keyframe(c, g["centered"][:5], polar, g["velocity"])) for c in g["cands"]]
    assert good == [bool(v) for v in g["good"]]
    assert np.abs(mp.getMapPointsGlobal() - g["global_pruned"]).max() <= 1e-9
    kf2 = M.Keyframe(g["pose"], g["centered"][:3], polar, g["velocity"])
    kf2.copyFromOtherKeyframe(kf)
    assert kf2.featurePointsLocal.shape == kf.featurePointsLocal.shape
