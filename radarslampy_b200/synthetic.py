"""Synthetic Oxford-shaped radar sequences (the role genFakeData.py plays in the reference,
for whole scans instead of correspondences).  SURVEY.md §8(d) config 2:

  raw uint8 [400, 3779] per frame, 11 metadata bytes per azimuth exactly as
  parseData.py:39-42 reads them (int64 us timestamp, uint16 encoder 14*i+13, valid = 255),
  3768 power bins.  World = point scatterers uniform in a square, rendered with a Gaussian
  PSF (sigma 2.5 range bins, 0.6 azimuth bins) plus exponential speckle; ego motion is a
  constant twist sampled at 4 Hz.  Everything is seeded.

Host-side NumPy only (input generation is not part of the hot path)."""
import numpy as np

A = 400
RAW_WIDTH = 3779
META = 11
BINS = RAW_WIDTH - META


class World:
    def __init__(self, n_scatterers=4000, extent_m=250.0, seed=1234):
        rng = np.random.default_rng(seed)
        self.xy = rng.uniform(-extent_m / 2, extent_m / 2, (n_scatterers, 2))
        self.amp = rng.uniform(80.0, 255.0, n_scatterers)


def twist_pose(k, v=10.0, w=0.10, hz=4.0):
    """Pose (x, y, theta) after k frames of a constant twist (v m/s forward, w rad/s)."""
    t = k / hz
    th = w * t
    if abs(w) < 1e-12:
        return np.array([v * t, 0.0, 0.0])
    return np.array([v / w * np.sin(th), v / w * (1 - np.cos(th)), th])


def sensor_points(world_xy, pose):
    """World points in the sensor frame of `pose` (x, y, theta)."""
    c, s = np.cos(pose[2]), np.sin(pose[2])
    d = world_xy - pose[:2]
    return np.column_stack([c * d[:, 0] + s * d[:, 1], -s * d[:, 0] + c * d[:, 1]])


def twist_poses(k, v=10.0, w=0.10, hz=4.0):
    """twist_pose for an array of (fractional) frame indices -> [n, 3]."""
    t = np.asarray(k, np.float64) / hz
    th = w * t
    if abs(w) < 1e-12:
        return np.column_stack([v * t, np.zeros_like(t), np.zeros_like(t)])
    return np.column_stack([v / w * np.sin(th), v / w * (1 - np.cos(th)), th])


def distorted_sensor_points(world_xy, frame_idx, v, w):
    """Intra-scan motion distortion (BASELINE configs[2]): every scatterer is seen from the pose the sensor has when the
    beam sweeps over it.  Timing follows the model the reference's solver assumes (motionDistortion.py:107-124):
    the frame's pose is the mid-scan pose and the beam at azimuth row a fires (a / 400 - 0.5) scan periods later."""
    ps = sensor_points(world_xy, twist_pose(frame_idx, v, w))
    for _ in range(3):                                  # fixed point: the firing time depends on the observed azimuth
        frac = (np.arctan2(ps[:, 1], ps[:, 0]) % (2 * np.pi)) / (2 * np.pi) - 0.5
        pose = twist_poses(frame_idx + frac, v, w)
        c, s = np.cos(pose[:, 2]), np.sin(pose[:, 2])
        d = world_xy - pose[:, :2]
        ps = np.column_stack([c * d[:, 0] + s * d[:, 1], -s * d[:, 0] + c * d[:, 1]])
    return ps


def render_scan(world: World, pose, frame_idx, res_m=0.0438, t0_us=1_547_131_046_000_000, speckle_mean=8.0,
                sigma_r=2.5, sigma_a=0.6, seed=5678, distort=None):
    """One raw scan uint8 [400, 3779].  distort = (v, w) renders intra-scan motion distortion for that twist."""
    ps = sensor_points(world.xy, pose) if distort is None else distorted_sensor_points(world.xy, frame_idx, *distort)
    rb = np.hypot(ps[:, 0], ps[:, 1]) / res_m                       # range in bins
    az = (np.arctan2(ps[:, 1], ps[:, 0]) % (2 * np.pi)) / (2 * np.pi) * A   # azimuth in rows
    keep = rb < BINS + 8
    rb, az, amp = rb[keep], az[keep], world.amp[keep]
    img = np.zeros((A, BINS), np.float32)
    dr = np.arange(-8, 9)
    da = np.arange(-2, 3)
    r0 = np.round(rb).astype(np.int64)
    a0 = np.round(az).astype(np.int64)
    R = r0[:, None, None] + dr[None, None, :]                        # [n, 1, 17]
    Arow = a0[:, None, None] + da[None, :, None]                     # [n, 5, 1]
    val = amp[:, None, None] * np.exp(-0.5 * (((R - rb[:, None, None]) / sigma_r) ** 2 +
                                              ((Arow - az[:, None, None]) / sigma_a) ** 2))
    R = np.broadcast_to(R, val.shape)
    Arow = np.broadcast_to(Arow % A, val.shape)
    ok = (R >= 0) & (R < BINS)
    np.add.at(img, (Arow[ok], R[ok]), val[ok].astype(np.float32))
    rng = np.random.default_rng(seed + frame_idx)
    img += rng.exponential(speckle_mean, img.shape).astype(np.float32)
    raw = np.empty((A, RAW_WIDTH), np.uint8)
    raw[:, META:] = np.clip(img, 0, 255).astype(np.uint8)
    ts = (t0_us + frame_idx * 250_000 + (np.arange(A) * 250_000) // A).astype("<i8")
    raw[:, 0:8] = ts.view(np.uint8).reshape(A, 8)
    enc = (14 * np.arange(A) + 13).astype("<u2")
    raw[:, 8:10] = enc.view(np.uint8).reshape(A, 2)
    raw[:, 10] = 255
    return raw


def make_sequence(n_frames, res_m=0.0438, world: World = None, v=10.0, w=0.10, first=0, out=None, distort=False):
    """raw [n_frames, 400, 3779] uint8 and the ground-truth (mid-scan) poses [n_frames, 3]."""
    world = world or World()
    raw = out if out is not None else np.empty((n_frames, A, RAW_WIDTH), np.uint8)
    poses = np.zeros((n_frames, 3))
    for k in range(n_frames):
        poses[k] = twist_pose(first + k, v, w)
        raw[k] = render_scan(world, poses[k], first + k, res_m=res_m, distort=(v, w) if distort else None)
    return raw, poses


def scatterer_features(world: World, pose, res_m, range_bins, k=200, margin_px=24):
    """Pixel coordinates (x, y) f32 [<=k, 2] of the k strongest scatterers visible in the
    Cartesian image of a frame at `pose` — the 'features given' input of BASELINE config 4."""
    R = range_bins // 2
    px_per_m = (R / range_bins) / res_m            # cart pixel = bin * R / W
    ps = sensor_points(world.xy, pose) * px_per_m + R
    ok = (ps[:, 0] > margin_px) & (ps[:, 0] < 2 * R - margin_px) & (ps[:, 1] > margin_px) & (ps[:, 1] < 2 * R - margin_px)
    ok &= np.hypot(ps[:, 0] - R, ps[:, 1] - R) < R - margin_px
    idx = np.flatnonzero(ok)
    idx = idx[np.argsort(-world.amp[idx], kind="stable")][:k]
    return ps[idx].astype(np.float32)


def sequence_pairs(n_frames, world, poses, res_m, range_bins, k=200, max_features=256):
    """Consecutive pairs (i, i+1) with features on frame i: pair_idx [P,2] i32,
    feats [P, max_features, 2] f32, counts [P] i32."""
    P = n_frames - 1
    pair_idx = np.column_stack([np.arange(P), np.arange(1, P + 1)]).astype(np.int32)
    feats = np.zeros((P, max_features, 2), np.float32)
    counts = np.zeros(P, np.int32)
    for p in range(P):
        f = scatterer_features(world, poses[p], res_m, range_bins, k)
        counts[p] = len(f)
        feats[p, :len(f)] = f
    return pair_idx, feats, counts


def _render_job(job):
    seed, n_scat, extent, k, first, res_m, v, w, distort = job
    world = World(n_scat, extent, seed)
    pose = twist_pose(first + k, v, w)
    return render_scan(world, pose, first + k, res_m=res_m, distort=(v, w) if distort else None)


def make_sequence_parallel(n_frames, res_m=0.0438, seed=1234, v=10.0, w=0.10, first=0, out=None, distort=False, workers=None,
                           n_scatterers=4000, extent_m=250.0):
    """make_sequence over a process pool (identical output: every frame is rendered from its own seeds).  Call it
    BEFORE the CUDA context exists (fork)."""
    import multiprocessing as mp
    import os
    workers = max(1, min(workers or (os.cpu_count() or 1), n_frames))
    raw = out if out is not None else np.empty((n_frames, A, RAW_WIDTH), np.uint8)
    poses = np.stack([twist_pose(first + k, v, w) for k in range(n_frames)]) if n_frames else np.zeros((0, 3))
    jobs = [(seed, n_scatterers, extent_m, k, first, res_m, v, w, distort) for k in range(n_frames)]
    if workers == 1:
        for k, j in enumerate(jobs):
            raw[k] = _render_job(j)
        return raw, poses
    with mp.get_context("fork").Pool(workers) as pool:
        for k, scan in enumerate(pool.imap(_render_job, jobs, chunksize=max(1, n_frames // (4 * workers)))):
            raw[k] = scan
    return raw, poses


def dense_scene(seed, shift=(0.0, 0.0), n=2000, n_points=60000):
    """BASELINE configs[4] input: dense point-scatterer scene on an n x n Cartesian grid (Gaussian blobs, sigma 1.5 px)
    plus weak speckle; `shift` moves every scatterer (the second frame of a tracking pair).  f32 in [0, 1]."""
    rng = np.random.default_rng(seed)
    pts = rng.uniform(8, n - 8, (n_points, 2))
    amp = rng.uniform(0.3, 1.0, len(pts))
    img = np.zeros((n, n), np.float32)
    x, y = pts[:, 0] + shift[0], pts[:, 1] + shift[1]
    ix, iy = np.floor(x).astype(int), np.floor(y).astype(int)
    for dy in range(-4, 6):
        for dx in range(-4, 6):
            xx, yy = ix + dx, iy + dy
            ok = (xx >= 0) & (xx < n) & (yy >= 0) & (yy < n)
            w = amp * np.exp(-((xx - x) ** 2 + (yy - y) ** 2) / (2 * 1.5 ** 2))
            np.add.at(img, (yy[ok], xx[ok]), w[ok].astype(np.float32))
    img += np.random.default_rng(99).exponential(0.01, img.shape).astype(np.float32)
    return np.clip(img, 0, 1).astype(np.float32)
