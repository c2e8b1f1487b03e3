"""CPU tests of the drop-in boundary: libradarfe.so loads, exports every symbol that
include/radarfe.h declares, its POD structs match the ctypes mirrors, and it fails loudly
(no CPU fallback) when no sm_100 GPU is present.  No compute calls."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "radarfe.h")).read()
    return sorted(set(re.findall(r"RF_API\s+[\w\s\*]+?\b(rf_\w+)\s*\(", txt)))


@pytest.fixture(scope="module")
def lib():
    from radarslampy_b200 import _build, _ffi
    _build.build()
    return _ffi.load_library()


def test_every_declared_symbol_is_exported(lib):
    from radarslampy_b200 import _ffi
    declared = _header_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"libradarfe.so does not export {name}"
    assert sorted(_ffi.SYMBOLS) == declared, "radarslampy_b200/_ffi.py SYMBOLS out of sync with include/radarfe.h"


def test_no_unintended_exports():
    import subprocess
    from radarslampy_b200 import _ffi
    out = subprocess.run(["nm", "-D", "--defined-only", _ffi.LIB_PATH], capture_output=True, text=True).stdout
    ours = [l.split()[-1] for l in out.splitlines() if " T " in l and l.split()[-1].startswith("rf_")]
    assert sorted(ours) == _header_symbols()


def test_default_config_is_the_reference_constants(lib):
    from radarslampy_b200 import _ffi
    cfg = _ffi.default_config()
    assert (cfg.azimuths, cfg.raw_width, cfg.meta_bytes, cfg.range_bins, cfg.downsample) == (400, 3779, 11, 2025, 2)
    assert (cfg.klt_win, cfg.klt_max_level, cfg.klt_max_iters) == (15, 3, 10)
    assert abs(cfg.klt_eps - 0.03) < 1e-7 and abs(cfg.klt_err_thr - 10) < 1e-7
    assert cfg.dist_thr_px == 0.5 / 0.0864            # outlierRejection.py:10-11
    assert cfg.cart_res_m == 0.0864 and cfg.mds_period == 0.25
    assert list(cfg.mds_sigma_p) == [4.0, 4.0]
    assert list(cfg.mds_sigma_v) == [1.0, 1.0, (5 * np.pi / 180) ** 2]
    assert lib.rf_version() == 100


def test_pair_result_layout():
    from radarslampy_b200 import _ffi
    assert C.sizeof(_ffi.RfPairResult) == _ffi.PAIR_RESULT_DTYPE.itemsize == 120
    assert _ffi.PAIR_RESULT_DTYPE.fields["mds_x"][1] == 48


def test_create_fails_loudly_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    from radarslampy_b200 import _ffi
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        _ffi.RadarFE(device=0)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under radarslampy_b200/ may reference it."""
    pkg = os.path.join(ROOT, "radarslampy_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M), f"{f} imports oracle"
                assert "liboracle" not in txt, f"{f} links the oracle"
