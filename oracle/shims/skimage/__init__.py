"""Import shim (test infrastructure): scikit-image is absent from this image.  The
reference only needs `skimage.feature.blob_doh`; the stand-in raises unless a test
installs a detector via `set_blob_doh` (detector parity is unpinned)."""
