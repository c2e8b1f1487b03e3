import sys, numpy as np
sys.path.insert(0, '.')
from oracle import ref_system
from radarslampy_b200 import odometry, synthetic as S
from radarslampy_b200.getFeatures import appendNewFeatures
raw, _ = S.make_sequence(24, res_m=0.0432, world=S.World(seed=4321))
got = odometry.run_odometry(raw)
want = ref_system.run_odometry(raw, lambda cart, old: appendNewFeatures(cart, old)[0])
gt, wt = np.asarray(got["n_tracked"]), np.asarray(want["n_tracked"])
print("n_tracked got ", gt.tolist())
print("n_tracked want", wt.tolist())
gi, wi = np.asarray(got["n_features_in"]), np.asarray(want["n_features_in"])
print("feat_in got ", gi.tolist())
print("feat_in want", wi.tolist())
d = got["traj"].poses - want["poses"]
print("pose diff per frame", np.abs(d).max(axis=1).round(6).tolist())
print("retrack got", np.asarray(got["retrack"]).astype(int).tolist())
print("retrack want", np.asarray(want["retrack"]).astype(int).tolist())
