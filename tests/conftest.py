import glob
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_NAMES = sorted(n for n in (os.path.splitext(os.path.basename(p))[0]
                                  for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
                      if not n.startswith(("encoder", "pixel_decoder")))       # op-level fixtures only

# tolerances of BASELINE.json's north_star
FWD_ABS_TOL = 1e-5      # fp32 forward, max abs error vs the fp64 oracle, value ~ N(0,1)
GRAD_REL_TOL = 1e-4     # per tensor: max|g - g_ref| / max|g_ref|


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def rel_err(got, ref):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    denom = np.abs(ref).max()
    num = np.abs(got - ref).max()
    return float(num / denom) if denom > 0 else float(num)


@pytest.fixture(scope="session")
def pkg():
    from __graft_entry__ import load_package
    return load_package()


@pytest.fixture(scope="session")
def oracle():
    from __graft_entry__ import load_oracle
    return load_oracle()
