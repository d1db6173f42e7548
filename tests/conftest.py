"""Shared test plumbing.  `-m gpu` tests need a B200; everything else runs on CPU.

sys.path: the product package lives in lsqfakequantize-pytorch_b200/ (directory name is not an
identifier, so it is put on sys.path and imported as `torchlsq`); the oracle is imported as
`oracle.lsq_oracle` from the repo root (tests only).
"""
import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
PKG = ROOT / "lsqfakequantize-pytorch_b200"
for p in (str(PKG), str(ROOT)):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")
    # a fresh checkout has no built artefacts (they are git-ignored): build the two in-tree libraries once, before any test imports
    # the package (nvcc cross-compiles sm_100a without a GPU; same command as __graft_entry__.build())
    pkg = PKG / "torchlsq"
    if not (pkg / "libtorchlsq_b200.so").exists() or not (pkg / "_C.so").exists():
        import subprocess
        subprocess.check_call(["make", "-C", str(PKG / "csrc"), "-j8"], stdout=subprocess.DEVNULL)


def pytest_sessionfinish(session, exitstatus):
    """GPU sessions: dump the achieved parity margins (tests/gpu_util.py::assert_grads_close) next to the other evidence."""
    gu = sys.modules.get("gpu_util")
    margins = getattr(gu, "MARGINS", None) if gu is not None else None
    if not margins:
        return
    worst = {}
    for rec in margins:
        key = (rec["dtype"], rec["rel_bound"])
        w = worst.setdefault(key, dict(dtype=rec["dtype"], rel_bound=rec["rel_bound"], calls=0, max_plain_rel_err=0.0,
                                       max_plain_rel_err_where_cancellation_le_16=0.0,
                                       max_plain_rel_err_where_cancellation_le_100=0.0, max_cancellation_ratio=0.0))
        w["calls"] += 1
        for k in ("max_plain_rel_err", "max_plain_rel_err_where_cancellation_le_16", "max_plain_rel_err_where_cancellation_le_100",
                  "max_cancellation_ratio"):
            w[k] = max(w[k], rec[k])
    out = ROOT / "gpurun_out"
    try:
        out.mkdir(exist_ok=True)
        (out / "parity_margins.json").write_text(json.dumps({"summary": list(worst.values()), "calls": len(margins)}, indent=1))
    except OSError:
        pass


@pytest.fixture(scope="session")
def golden_ops():
    data = np.load(GOLDEN / "ref_cpu_ops.npz")
    meta = json.loads((GOLDEN / "ref_module_trace.json").read_text())["ops_meta"]
    cases = {}
    for name, m in meta.items():
        cases[name] = dict(meta=m, **{k: data[f"{name}/{k}"] for k in ("x", "g", "scale", "shift", "y", "dx", "ds", "db")})
    return cases


@pytest.fixture(scope="session")
def golden_module():
    return json.loads((GOLDEN / "ref_module_trace.json").read_text())


@pytest.fixture(scope="session")
def native_lib():
    """Build (if needed) and load the C-ABI library; used by CPU symbol tests and GPU tests."""
    import subprocess
    so = PKG / "torchlsq" / "libtorchlsq_b200.so"
    if not so.exists():
        subprocess.check_call(["make", "-C", str(PKG / "csrc"), "-j8"], stdout=subprocess.DEVNULL)
    from torchlsq import _cabi
    return _cabi.load()


def geometry(shape, axis):
    outer = int(np.prod(shape[:axis], dtype=np.int64))
    inner = int(np.prod(shape[axis + 1:], dtype=np.int64))
    return outer, shape[axis], inner
