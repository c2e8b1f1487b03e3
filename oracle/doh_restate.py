"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's feature detector,
skimage.feature.blob_doh as getFeatures.getBlobsFromCart calls it (getFeatures.py:13-18,22-53):

    blob_doh(img.astype(np.double), min_sigma=.01, max_sigma=10, num_sigma=3, threshold=.0005)

PARITY UNPINNED.  scikit-image 0.19.2 (requirements.txt:5) is a third-party dependency that is not vendored in
/root/reference, is not installed in this image and cannot be installed (no index).  The reference holds no golden
vector for the detector.  What follows restates the PUBLISHED algorithm of scikit-image 0.19.2
  skimage/feature/blob.py            blob_doh, _prune_blobs, _blob_overlap, _compute_disk_overlap
  skimage/transform/integral.py      integral_image
  skimage/feature/_hessian_det_appx.pyx   _hessian_matrix_det, _integ, _clip     (cdivision=True)
  skimage/feature/peak.py            peak_local_max, _get_peak_mask, _get_high_intensity_peaks
  skimage/_shared/coord.py           ensure_spacing (a no-op here: integer coordinates, spacing 1, p = inf)
and runs the SciPy routines those functions bottom out in (scipy.ndimage.maximum_filter, scipy.spatial.cKDTree) where
SciPy is the arbiter of the result.  Three points of the published algorithm are implementation-defined and are
written down here (DESIGN.md §3 repeats them):

 (1) NaN plane.  With min_sigma = 0.01 the first scale has box size int(3 * 0.01) = 0, so w_i = 1.0 / 0 / 0 = inf
     (C division) and dxy = -0.0 * inf = NaN: the whole sigma = 0.01 plane is NaN.  peak_local_max runs
     scipy.ndimage.maximum_filter(cube, footprint=ones((3, 3, 3)), mode='nearest'); an all-true footprint takes
     SciPy's separable path (maximum_filter1d per axis, ni_filters.c NI_MinOrMaximumFilter1D: a ring buffer whose
     comparisons are all false for NaN).  Along the scale axis the line [NaN, M1, M2] then filters to
     [NaN, NaN, max(M1, M2)]: plane 0 and plane 1 can never equal their filtered value, so every blob comes from the
     LAST scale (sigma = 10), compared against the 3 x 3 neighbourhoods of the sigma = 5.005 and sigma = 10 planes.
     `scale_axis_max` restates that ring buffer; tests/test_doh_oracle.py checks it against the SciPy in this image
     (1.18.1; the reference pins 1.7.3, whose ni_filters.c has the same routine).  skimage < 0.18 used
     mode='constant', under which plane 1 would also yield blobs.
 (2) Order of equal responses.  _get_high_intensity_peaks sorts with np.argsort(-intensities) (unstable).  Here:
     stable, i.e. ties keep np.nonzero's C order (row, column, scale).
 (3) Pruning order.  _prune_blobs walks `list(tree.query_pairs(d))` — a Python set of (i, j) tuples filled in the
     traversal order of a C++ k-d tree — and zeroes the sigma of one blob of every overlapping pair using the
     sigmas as they are AT THAT MOMENT, so chains of overlapping blobs (A~B, B~C) prune differently depending on
     that order, which changes with the SciPy build and the CPython hash-table layout.  `prune_blobs(order="sorted")`
     (the product's definition, csrc/k_doh.cu) walks the pairs in ascending (i, j); `order="scipy"` walks whatever
     this image's SciPy / CPython produce.  The two agree whenever no blob is both the loser of one pair and the
     winner of another.

Everything is float64 and evaluated in the operand order of the C / Python sources (no fused multiply-add)."""
import math

import numpy as np


def integral_image(img):
    """skimage.transform.integral_image: S = img.cumsum(0).cumsum(1) (float64, sequential sums)."""
    S = np.asarray(img, np.float64)
    for ax in range(S.ndim):
        S = S.cumsum(axis=ax)
    return S


def _integ(ii, r, c, rl, cl):
    """_hessian_det_appx.pyx:_integ for index ARRAYS r, c (window lengths rl, cl are scalars)."""
    H, W = ii.shape
    r = np.clip(r, 0, H - 1)
    c = np.clip(c, 0, W - 1)
    r2 = np.clip(r + rl, 0, H - 1)
    c2 = np.clip(c + cl, 0, W - 1)
    R, Cc = r[:, None], c[None, :]
    R2, C2 = r2[:, None], c2[None, :]
    ans = ((ii[R, Cc] + ii[R2, C2]) - ii[R, C2]) - ii[R2, Cc]
    return np.where(ans > 0, ans, 0.0)                      # Cython max(0, ans): (ans > 0) ? ans : 0


def box_geometry(sigma):
    """(size, l, b, w, w_i) of _hessian_matrix_det; C integer division truncates towards zero (cdivision=True)."""
    size = int(3 * sigma)
    b = int((size - 1) / 2)                                 # (size - 1) / 2 in C: (-1) / 2 == 0
    l = size // 3
    w = size
    with np.errstate(divide="ignore", invalid="ignore"):
        w_i = np.float64(1.0) / np.float64(size) / np.float64(size)
    return size, l, b, w, float(w_i)


def hessian_det(ii, sigma):
    """_hessian_matrix_det(integral image, sigma) -> float64 plane."""
    H, W = ii.shape
    size, l, b, w, w_i = box_geometry(sigma)
    r = np.arange(H)
    c = np.arange(W)
    with np.errstate(invalid="ignore", over="ignore"):
        tl = _integ(ii, r - l, c - l, l, l)
        br = _integ(ii, r + 1, c + 1, l, l)
        bl = _integ(ii, r - l, c + 1, l, l)
        tr = _integ(ii, r + 1, c - l, l, l)
        dxy = ((bl + tr) - tl) - br
        dxy = (-dxy) * w_i
        mid = _integ(ii, r - l + 1, c - b, 2 * l - 1, w)
        side = _integ(ii, r - l + 1, c - l // 2, 2 * l - 1, l)
        dxx = mid - 3 * side
        dxx = (-dxx) * w_i
        mid = _integ(ii, r - b, c - l + 1, w, 2 * l - 1)
        side = _integ(ii, r - l // 2, c - l + 1, l, 2 * l - 1)
        dyy = mid - 3 * side
        dyy = (-dyy) * w_i
        return dxx * dyy - 0.81 * (dxy * dxy)


def sigma_list(min_sigma, max_sigma, num_sigma):
    return np.linspace(min_sigma, max_sigma, num_sigma)


def hessian_cube(img, min_sigma, max_sigma, num_sigma):
    ii = integral_image(np.asarray(img).astype(np.float64))
    return np.dstack([hessian_det(ii, s) for s in sigma_list(min_sigma, max_sigma, num_sigma)])


def ring_max_1d(line, size=3):
    """NI_MinOrMaximumFilter1D (maximum) on an already-extended line: out[i] = filtered value of line[i + size // 2]."""
    n = len(line) - (size - 1)
    ring = [[0.0, 0] for _ in range(size)]
    mp = last = 0
    ring[0] = [line[0], size]
    out = []
    for ll in range(1, size + n - 1):
        val = line[ll]
        if ring[mp][1] == ll:
            mp = (mp + 1) % size
        if val >= ring[mp][0]:
            ring[mp] = [val, ll + size]
            last = mp
        else:
            while ring[last][0] <= val:
                last = (last - 1) % size
            last = (last + 1) % size
            ring[last] = [val, ll + size]
        if ll >= size - 1:
            out.append(ring[mp][0])
    return out


def scale_axis_max(M):
    """Maximum filter of size 3, mode 'nearest', along the last axis of M [H, W, n] with SciPy's NaN behaviour."""
    H, W, n = M.shape
    ext = np.concatenate([M[..., :1], M, M[..., -1:]], axis=-1)
    if not np.isnan(M).any():
        return np.maximum(np.maximum(ext[..., :-2], ext[..., 1:-1]), ext[..., 2:])
    out = np.empty_like(M)
    nan_planes = tuple(bool(np.isnan(M[..., k]).all()) for k in range(n))
    if all(np.isnan(M[..., k]).any() == nan_planes[k] for k in range(n)):
        # planes are NaN either everywhere or nowhere: the comparison outcomes involving a NaN are the same for every
        # pixel, but those between finite planes are not — run the ring per distinct ordering of the finite values
        flat = ext.reshape(-1, n + 2)
        res = np.empty((flat.shape[0], n))
        for i in range(flat.shape[0]):
            res[i] = ring_max_1d(list(flat[i]))
        return res.reshape(H, W, n)
    flat = ext.reshape(-1, n + 2)
    res = np.empty((flat.shape[0], n))
    for i in range(flat.shape[0]):
        res[i] = ring_max_1d(list(flat[i]))
    return res.reshape(H, W, n)


def peak_local_max_3d(cube, threshold, use_scipy=True):
    """peak_local_max(cube, threshold_abs=threshold, exclude_border=False, footprint=ones((3, 3, 3))) -> int [K, 3]
    (row, col, scale index), highest response first (ties: C order)."""
    if use_scipy:
        from scipy import ndimage as ndi
        image_max = ndi.maximum_filter(cube, footprint=np.ones((3, 3, 3)), mode="nearest")
    else:
        # in-plane 3 x 3 maximum (planes are NaN everywhere or nowhere), then the scale axis with the ring buffer
        H, W, n = cube.shape
        p = np.pad(cube, ((1, 1), (1, 1), (0, 0)), mode="edge")
        m = p[1:-1, 1:-1]
        for dy in range(3):
            for dx in range(3):
                with np.errstate(invalid="ignore"):
                    m = np.where(p[dy:dy + H, dx:dx + W] > m, p[dy:dy + H, dx:dx + W], m)
        image_max = scale_axis_max(m)
    with np.errstate(invalid="ignore"):
        out = cube == image_max
        if np.all(out):                                        # "no peak for a trivial image"
            out[:] = False
        out &= cube > threshold
    coord = np.nonzero(out)
    intensities = cube[coord]
    order = np.argsort(-intensities, kind="stable")
    return np.transpose(coord)[order]


def _compute_disk_overlap(d, r1, r2):
    ratio1 = (d ** 2 + r1 ** 2 - r2 ** 2) / (2 * d * r1)
    ratio1 = np.clip(ratio1, -1, 1)
    acos1 = math.acos(ratio1)
    ratio2 = (d ** 2 + r2 ** 2 - r1 ** 2) / (2 * d * r2)
    ratio2 = np.clip(ratio2, -1, 1)
    acos2 = math.acos(ratio2)
    a = -d + r2 + r1
    b = d - r2 + r1
    c = d + r2 - r1
    d = d + r2 + r1
    area = (r1 ** 2 * acos1 + r2 ** 2 * acos2 - 0.5 * math.sqrt(abs(a * b * c * d)))
    return area / (math.pi * (min(r1, r2) ** 2))


def _blob_overlap(blob1, blob2):
    root_ndim = math.sqrt(2)
    if blob1[-1] == blob2[-1] == 0:
        return 0.0
    elif blob1[-1] > blob2[-1]:
        max_sigma = blob1[-1:]
        r1 = 1
        r2 = blob2[-1] / blob1[-1]
    else:
        max_sigma = blob2[-1:]
        r2 = 1
        r1 = blob1[-1] / blob2[-1]
    pos1 = blob1[:2] / (max_sigma * root_ndim)
    pos2 = blob2[:2] / (max_sigma * root_ndim)
    d = np.sqrt(np.sum((pos2 - pos1) ** 2))
    if d > r1 + r2:
        return 0.0
    if d <= abs(r1 - r2):
        return 1.0
    return _compute_disk_overlap(d, r1, r2)


def prune_blobs(blobs_array, overlap=0.5, order="sorted"):
    """skimage.feature.blob._prune_blobs (sigma_dim = 1).  See point (3) of the module docstring for `order`."""
    from scipy import spatial
    blobs_array = np.array(blobs_array, np.float64)
    if len(blobs_array) == 0:
        return blobs_array.reshape(0, 3)
    sigma = blobs_array[:, -1:].max()
    distance = 2 * sigma * math.sqrt(blobs_array.shape[1] - 1)
    tree = spatial.cKDTree(blobs_array[:, :-1])
    pairs = list(tree.query_pairs(distance))
    if order == "sorted":
        pairs = sorted(pairs)
    pairs = np.array(pairs)
    if len(pairs) == 0:
        return blobs_array
    for (i, j) in pairs:
        blob1, blob2 = blobs_array[i], blobs_array[j]
        if _blob_overlap(blob1, blob2) > overlap:
            if blob1[-1] > blob2[-1]:
                blob2[-1] = 0
            else:
                blob1[-1] = 0
    keep = [b for b in blobs_array if b[-1] > 0]
    return np.stack(keep) if keep else np.empty((0, 3))


def blob_doh(image, min_sigma=1, max_sigma=30, num_sigma=10, threshold=0.01, overlap=0.5, prune_order="sorted",
             use_scipy_filter=True):
    """skimage.feature.blob_doh(image, ...) -> float64 [K, 3] rows (row, col, sigma)."""
    sl = sigma_list(min_sigma, max_sigma, num_sigma)
    cube = hessian_cube(image, min_sigma, max_sigma, num_sigma)
    local_maxima = peak_local_max_3d(cube, threshold, use_scipy=use_scipy_filter)
    if local_maxima.size == 0:
        return np.empty((0, 3))
    lm = local_maxima.astype(np.float64)
    lm[:, -1] = sl[local_maxima[:, -1]]
    return prune_blobs(lm, overlap, order=prune_order)


def get_features(img, old_xy, min_sigma=0.01, max_sigma=10, num_sigma=3, threshold=0.0005, num_ret=200, tol=0.1,
                 prune_order="sorted"):
    """getFeatures.appendNewFeatures (getFeatures.py:66-118) on top of the restated detector; np.argsort(blobs[:, 2])
    (unstable in the reference, implementation-defined for equal sigmas) is taken stable."""
    from . import restate as R
    blobs = blob_doh(np.asarray(img).astype(np.double), min_sigma, max_sigma, num_sigma, threshold, prune_order=prune_order)
    kp = blobs[np.argsort(blobs[:, 2], kind="stable")]
    sel = R.ssc(kp, num_ret, tol, img.shape[1], img.shape[0]) if len(kp) else np.zeros(0, int)
    new = np.fliplr(kp[sel][:, :2])
    pts = np.vstack((np.asarray(old_xy, np.float64).reshape(-1, 2), new))
    _, idx = np.unique(pts, axis=0, return_index=True)
    return np.ascontiguousarray(pts[np.sort(idx)]).astype(np.float32)
