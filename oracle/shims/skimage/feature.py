_impl = None


def set_blob_doh(fn):
    global _impl
    _impl = fn


def blob_doh(image, **kw):
    if _impl is None:
        raise RuntimeError("skimage is not installed; install a stand-in with skimage.feature.set_blob_doh")
    return _impl(image, **kw)


def blob_dog(*a, **k):
    raise NotImplementedError


def blob_log(*a, **k):
    raise NotImplementedError
