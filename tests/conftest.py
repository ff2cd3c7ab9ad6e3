import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def net_from(g, prefix=""):
    Ws, bs = [], []
    i = 0
    while f"{prefix}W{i}" in g:
        Ws.append(g[f"{prefix}W{i}"]); bs.append(g[f"{prefix}b{i}"]); i += 1
    return Ws, bs


def grads_from(g, prefix=""):
    gWs, gbs = [], []
    i = 0
    while f"{prefix}gW{i}" in g:
        gWs.append(g[f"{prefix}gW{i}"]); gbs.append(g[f"{prefix}gb{i}"]); i += 1
    return gWs, gbs


def flat(gWs, gbs):
    """Flatten in nn.Module.parameters() order: W0, b0, W1, b1, ..."""
    return np.concatenate([np.concatenate([np.asarray(W).reshape(-1), np.asarray(b).reshape(-1)]) for W, b in zip(gWs, gbs)])


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    den = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (den if den > 0 else 1.0)


def rel_max(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    den = np.max(np.abs(b))
    return np.max(np.abs(a - b)) / (den if den > 0 else 1.0)


def assert_grads_close(got, want, tol, what=""):
    """Per-parameter-tensor relative L2 and global max-abs criteria."""
    (gW, gb), (wW, wb) = got, want
    scale = max(max(np.max(np.abs(w)) for w in wW), max(np.max(np.abs(b)) for b in wb))
    for i, (a, b) in enumerate(zip(list(gW) + list(gb), list(wW) + list(wb))):
        a = np.asarray(a, dtype=np.float64).reshape(np.asarray(b).shape)
        nb = np.linalg.norm(b)
        if nb > 1e-3 * scale * np.sqrt(b.size):
            assert np.linalg.norm(a - b) / nb <= tol, f"{what} tensor {i}: rel l2 {np.linalg.norm(a - b) / nb:.3e} > {tol}"
        assert np.max(np.abs(a - b)) <= tol * scale, f"{what} tensor {i}: max abs {np.max(np.abs(a - b)):.3e} > {tol}*{scale:.3e}"


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
