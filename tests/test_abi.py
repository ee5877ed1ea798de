"""CPU tests of the drop-in boundary: the shared library builds/loads without a GPU and exports exactly the symbols
include/scenerf_b200.h declares; argument validation works without touching the device."""
import ctypes as C
import os
import re

import pytest

from scenerf_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "scenerf_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return set(re.findall(r"\b(srf_[a-z0-9_]+)\s*\(", txt))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    declared = _declared_symbols()
    assert declared, "header parse found nothing"
    assert declared == set(_lib.SYMBOLS), "binding and header disagree: %s" % (declared ^ set(_lib.SYMBOLS))
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.srf_abi_version() == 1


def test_struct_layouts_match_header_sizes():
    # the library reports sizeof() of each ABI struct; the ctypes mirror must agree (guards against binding drift)
    lib = _lib.load()
    for which, st in enumerate((_lib.Config, _lib.Pyramid, _lib.MlpWeights, _lib.Outputs)):
        assert lib.srf_sizeof(which) == C.sizeof(st), st.__name__
    assert C.sizeof(_lib.Outputs) == 15 * 8


def test_argument_validation_without_gpu():
    lib = _lib.load()
    cfg = _lib.Config()
    cfg.n_gaussians = 99
    cfg.n_pts_uni = 32
    cfg.n_pts_per_gaussian = 8
    out = _lib.Outputs()
    rc = lib.srf_render_rays(C.byref(cfg), None, None, None, None, 4, None, None, C.byref(out), None, 0, None)
    assert rc == 1 and b"n_gaussians" in lib.srf_last_error()
    with pytest.raises(ValueError):
        _lib.check(rc)
    cfg.n_gaussians = 4
    cfg.sphere_W, cfg.sphere_H = 300, 90
    cfg.precision = 7
    rc = lib.srf_render_rays(C.byref(cfg), None, None, None, None, 4, None, None, C.byref(out), None, 0, None)
    assert rc == 1 and b"precision" in lib.srf_last_error()
    cfg.precision = 0
    assert lib.srf_render_workspace_bytes(C.byref(cfg), 1024) > 1024 * 64 * 4 * 9


def test_renderer_refuses_cpu():
    import torch
    from scenerf_b200.renderer import B200Renderer
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        B200Renderer({}, {}, {}, device="cpu")
