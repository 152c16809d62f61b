import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


# Achieved parity errors (max relative error per case / team / path) collected by the GPU parity tests; written to
# gpurun_out/parity_errors.json at the end of a `-m gpu` session so the measured numbers (not just the tolerances)
# can be quoted (profiles/r2_parity_errors.json is a committed copy).
PARITY_ERRORS = {}


def record_parity_error(key, **errs):
    PARITY_ERRORS[key] = {k: float(v) for k, v in errs.items()}


def pytest_sessionfinish(session, exitstatus):
    if not PARITY_ERRORS:
        return
    import json
    out = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_errors.json"), "w") as f:
            json.dump(PARITY_ERRORS, f, indent=1, sort_keys=True)
    except OSError:
        pass
