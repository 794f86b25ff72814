import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def lib():
    from qex_b200 import _lib

    return _lib.load(build_if_missing=True)


def pytest_sessionfinish(session, exitstatus):
    """Writes the max-norm AND element-wise relative error of every comparison the session made (worst per test)
    to gpurun_out/parity_report.json, so both figures are on record next to the pass/fail verdict."""
    import json

    from tests import _util

    if not _util.REPORT:
        return
    worst = {}
    for tid, r, e in _util.REPORT:
        w = worst.setdefault(tid, [0.0, 0.0, 0])
        w[0], w[1], w[2] = max(w[0], r), max(w[1], e), w[2] + 1
    out = {"elementwise_floor": f"denominator max(|ref_i|, {_util.ELEM_FLOOR:g} * max|ref|)",
           "tests": {k: {"maxnorm_rel": v[0], "elementwise_rel": v[1], "comparisons": v[2]} for k, v in worst.items()}}
    try:
        d = os.path.join(ROOT, "gpurun_out")
        os.makedirs(d, exist_ok=True)
        gpu = "gpu" if any("gpu" in k for k in worst) else "cpu"
        with open(os.path.join(d, f"parity_report_{gpu}.json"), "w") as f:
            json.dump(out, f, indent=1, sort_keys=True)
    except OSError:
        pass
