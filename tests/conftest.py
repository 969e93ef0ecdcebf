import ast
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


def load_golden(name):
    """-> dict of numpy arrays; 'sd/…' entries are regrouped into a torch state_dict under 'sd'."""
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    out, sd, sd0, sd1 = {}, {}, {}, {}
    for k in z.files:
        if k.startswith("sd/"):
            sd[k[3:]] = torch.from_numpy(z[k])
        elif k.startswith("sd0/"):
            sd0[k[4:]] = torch.from_numpy(z[k])
        elif k.startswith("sd1/"):
            sd1[k[4:]] = torch.from_numpy(z[k])
        else:
            out[k] = z[k]
    out["sd"], out["sd0"], out["sd1"] = sd, sd0, sd1
    if "cfg" in out:
        out["cfg"] = ast.literal_eval(str(out["cfg"]))
    return out


def rel_l2(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return ((a - b).norm() / b.norm()).item()


def rel_max(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return ((a - b).abs().max() / b.abs().max()).item()


@pytest.fixture(scope="session")
def golden():
    return load_golden
