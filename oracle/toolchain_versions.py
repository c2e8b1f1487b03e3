"""TEST INFRASTRUCTURE (oracle/): records which interpreter and third-party wheels produced tests/golden/*.

The clique goldens (outlierRejection.py:62-78) depend on the iteration order of CPython sets and on networkx's
find_cliques; the KLT / warpPolar goldens on the OpenCV build; SSC / peaks / MDS on NumPy and SciPy.  Goldens
recorded under another interpreter can diverge silently on tied maximum cliques, so the generators write this
record (tests/golden/VERSIONS.json) and tests/test_oracle_pins.py::test_golden_toolchain_versions compares it
with the running interpreter.

    python -m oracle.toolchain_versions          # rewrite the record (run together with the gen_golden*.py scripts)
"""
import json
import os
import sys

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
PATH = os.path.join(GOLD, "VERSIONS.json")


def current() -> dict:
    import cv2
    import networkx
    import numpy
    import scipy
    return {"cpython": "%d.%d.%d" % sys.version_info[:3], "networkx": networkx.__version__, "scipy": scipy.__version__,
            "opencv": cv2.__version__, "numpy": numpy.__version__}


def write() -> dict:
    v = current()
    with open(PATH, "w") as f:
        json.dump(v, f, indent=1, sort_keys=True)
        f.write("\n")
    return v


if __name__ == "__main__":
    print(write())
