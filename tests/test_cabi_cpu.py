"""CPU: the C-ABI library loads and exports every symbol include/gpkl.h declares; descriptor checks and
workspace sizing run without a GPU (no compute calls)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def _lib():
    import __graft_entry__ as ge
    ge.build()
    import gpkl
    return gpkl


def test_header_symbols_exported():
    gpkl = _lib()
    hdr = open(os.path.join(ROOT, "include", "gpkl.h")).read()
    declared = set(re.findall(r"\b(gpkl_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(gpkl._lib.SYMBOLS), declared ^ set(gpkl._lib.SYMBOLS)
    L = gpkl._lib.lib()
    for s in declared:
        assert getattr(L, s) is not None
    assert L.gpkl_version() == 1
    assert b"workspace" in L.gpkl_strerror(4)


def test_descriptor_validation_and_workspace():
    gpkl = _lib()
    from gpkl.ops import _make_desc, workspace_bytes
    L = gpkl._lib.lib()
    d = _make_desc(256, 35, 48, 1, 256 * 48, "rbf", "gp", 1e-3)
    n = workspace_bytes(d)
    assert n >= (256 + 1) * 8 + 2 * 256 * 35 * 4
    assert L.gpkl_step_host_bytes(ctypes.byref(d)) > n
    # large T needs per-CTA matrix slots in the workspace
    big = _make_desc(8, 4, 512, 1, 8 * 512, "cauchy", "gp", 1e-3)
    assert workspace_bytes(big) > workspace_bytes(_make_desc(8, 4, 64, 1, 8 * 64, "cauchy", "gp", 1e-3))
    bad = _make_desc(4, 0, 8, 1, 32, "rbf", "gp", 1e-3)
    assert L.gpkl_workspace_bytes(ctypes.byref(bad)) == 0
    # error paths return codes, never touch the device: NULL pointers
    rc = L.gpkl_forward(ctypes.byref(d), *([None] * 12), None, 0, None)
    assert rc == 1
    rc = L.gpkl_forward(ctypes.byref(bad), *([None] * 12), None, 0, None)
    assert rc == 2


def test_ops_refuse_cpu_tensors():
    """There is no CPU fallback: the op refuses host tensors loudly."""
    gpkl = _lib()
    import torch
    with pytest.raises(RuntimeError, match="no CPU implementation"):
        gpkl.gp_prior_kl_forward(torch.zeros(4, 2), torch.zeros(1, 4), torch.tensor([4], dtype=torch.int32),
                                 torch.ones(2), torch.ones(2), torch.zeros(1, 2, 1, 4))
