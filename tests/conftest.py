import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_layers():
    import torch
    return torch.load(os.path.join(ROOT, "tests", "golden", "layers.pt"))


@pytest.fixture(scope="session")
def golden_models():
    import torch
    return torch.load(os.path.join(ROOT, "tests", "golden", "models.pt"))


@pytest.fixture(scope="session")
def golden_next():
    import torch
    return torch.load(os.path.join(ROOT, "tests", "golden", "next.pt"))


@pytest.fixture(scope="session")
def golden_nnconv():
    import torch
    return torch.load(os.path.join(ROOT, "tests", "golden", "nnconv.pt"))


@pytest.fixture(params=["fp32", "tf32"])
def math_mode(request):
    """Runs a GPU test once with exact-fp32 projections and once with the TF32 tensor-core projections."""
    import helpers
    from glam_b200 import _lib
    _lib.set_math_mode(request.param)
    helpers.MATH_MODE["mode"] = request.param
    yield request.param
    _lib.set_math_mode("tf32")
    helpers.MATH_MODE["mode"] = "fp32"


def pytest_terminal_summary(terminalreporter):
    """Realised parity errors of the run (tests/helpers.tol_check): the largest error / tensor-scale per math mode, next to the
    bound that was applied — so a loose bound cannot hide a large error."""
    try:
        import helpers
    except ImportError:
        return
    if not helpers.REALISED:
        return
    tr = terminalreporter
    for mode in ("fp32", "tf32", "fp32-abs", "tf32-abs"):
        rows = sorted((r for r in helpers.REALISED if r[1] == mode), key=lambda r: -r[2])
        if not rows:
            continue
        tr.write_line(f"parity, {mode} mode: {len(rows)} tensors checked; largest realised |err| / scale:")
        for name, _, e, b in rows[:8]:
            tr.write_line(f"    {e:9.2e}  (bound {b:8.2e})  {name}")
