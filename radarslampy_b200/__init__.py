"""radarslampy_b200 — B200-native (sm_100a) drop-in for the per-frame radar-odometry front
end of Samleo8/RadarSLAMPy.  The modules parseData / getFeatures / ANMS / getTransformKLT /
outlierRejection / motionDistortion / getPointCloud / Tracker keep the reference's names and
call signatures and run on hand-written CUDA kernels behind the C ABI in include/radarfe.h.
There is no CPU fallback."""
__version__ = "0.1.0"
