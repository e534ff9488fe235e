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


def test_adjacent_rows_argument_checks():
    """Error paths of the rows either side of the path (recog sampler, ragged batch producer) return codes before any
    device work: bad descriptors, NULL pointers, short workspaces."""
    gpkl = _lib()
    from gpkl.ops import _make_desc
    L = gpkl._lib.lib()
    d = _make_desc(5, 100, 20, 1, 100, "rbf", "gp", 1e-3)
    n = L.gpkl_recog_workspace_bytes(ctypes.byref(d))
    assert n > L.gpkl_workspace_bytes(ctypes.byref(d)) > 0          # inner workspace + scratch KL + offsets
    bad = _make_desc(5, 0, 20, 1, 100, "rbf", "gp", 1e-3)
    assert L.gpkl_recog_workspace_bytes(ctypes.byref(bad)) == 0
    assert L.gpkl_recog_forward(ctypes.byref(bad), *([None] * 10), None, 0, None) == 2     # GPKL_ERR_DESC
    assert L.gpkl_recog_forward(ctypes.byref(d), *([None] * 10), None, 0, None) == 1       # kl_sum NULL
    assert L.gpkl_recog_backward(ctypes.byref(d), *([None] * 13), None, 0, None) == 1
    assert L.gpkl_collate_workspace_bytes(8, 45) >= 9 * 8
    assert L.gpkl_collate_workspace_bytes(-1, 45) == 0
    assert L.gpkl_collate(4, 0, 45, 2, 45, *([None] * 7), None, 0, None) == 2              # F = 0
    assert L.gpkl_collate(4, 15, 45, 5, 45, *([None] * 7), None, 0, None) == 2             # B > N without an index
    assert L.gpkl_collate(4, 15, 45, 2, 45, *([None] * 7), None, 0, None) == 1             # NULL data
    assert L.gpkl_collate(4, 15, 45, 0, 45, *([None] * 7), None, 0, None) == 0             # empty batch: nothing to do


def test_reference_named_shims_exist():
    """The host-side mirrors of the reference's call sites (names and argument order of the reference functions)."""
    gpkl = _lib()
    import inspect
    for cls, names in ((gpkl.GPPriorPath, ("prior_kernels", "approx_kernels", "gp_vae_sample", "calc_gp_kl")),
                       (gpkl.GPRecogPath, ("approx_kernels", "gp_vae_sample", "standard_vae_kl")),
                       (gpkl.SyntheticDataHandlerGPU, ("data_batch",))):
        for n in names:
            assert callable(getattr(cls, n))
    sig = list(inspect.signature(gpkl.GPRecogPath.approx_kernels).parameters)
    assert sig[1:7] == ["sequences", "sequence_sizes", "latent_size", "batch_size", "number_samples", "encode_log_var"]
    sig = list(inspect.signature(gpkl.GPPriorPath.calc_gp_kl).parameters)
    assert sig[1:] == ["mean", "sequence_sizes", "approx_linear_kernel", "prior_kernel", "batch_size", "latent_size"]
