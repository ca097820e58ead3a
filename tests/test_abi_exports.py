"""
CPU-side checks of the drop-in boundary: the libraries load without a GPU,
export every symbol include/simwave_cuda.h declares, and the product path
fails loudly (no CPU fallback) when no device is present.
"""
import ctypes
import os
import re

import numpy as np
import pytest

import simwave_b200 as api
from simwave_b200.kernel.backend.compiler import prebuilt_library, LIB_DIR

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(REPO, "include", "simwave_cuda.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(simwave_(?:cuda|plan)_\w+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    names = declared_symbols()
    for dim in (2, 3):
        for dens in ("constant", "variable"):
            for prec in ("f32", "f64"):
                assert "simwave_cuda_forward_%dd_%s_%s" % (dim, dens, prec) in names
    for extra in ("simwave_cuda_last_error", "simwave_plan_create",
                  "simwave_plan_run", "simwave_plan_download",
                  "simwave_plan_destroy", "simwave_cuda_set_device"):
        assert extra in names


def test_core_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(os.path.join(LIB_DIR, "libsimwave_b200.so"))
    for name in declared_symbols():
        assert hasattr(lib, name), name


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("dens", ["constant_density", "variable_density"])
@pytest.mark.parametrize("prec", ["-DFLOAT", "-DDOUBLE"])
def test_every_shim_exports_forward(dim, dens, prec):
    path = prebuilt_library(dim, dens, prec)
    lib = ctypes.CDLL(path)
    assert hasattr(lib, "forward")
    # side exports are reachable through the shim as well (same process image)
    assert hasattr(lib, "simwave_cuda_last_error")


def test_compiler_returns_prebuilt_library_for_cuda():
    comp = api.Compiler(language="cuda")
    path = comp.compile(dimension=3, density="constant_density",
                        float_precision="-DFLOAT", operator="forward")
    assert path.endswith("libsimwave_cuda_3d_constant_f32.so")
    assert os.path.exists(path)


def test_no_cpu_fallback(monkeypatch):
    """No language ever selects a CPU kernel: without a custom kernel file the
    reference's other languages are served by the prebuilt CUDA library (with
    a warning, so the reference's examples run unmodified), or refused under
    SIMWAVE_B200_STRICT_LANGUAGE=1."""
    with pytest.warns(RuntimeWarning, match="CUDA"):
        path = api.Compiler(language="c").compile(2, "constant_density", "-DFLOAT",
                                                  "forward")
    assert path.endswith("libsimwave_cuda_2d_constant_f32.so")
    path = api.Compiler(cc="gcc", language="cpu_openmp", cflags="-O3 -fPIC").compile(
        3, "variable_density", "-DDOUBLE", "forward")
    assert path.endswith("libsimwave_cuda_3d_variable_f64.so")
    monkeypatch.setenv("SIMWAVE_B200_STRICT_LANGUAGE", "1")
    with pytest.raises(NotImplementedError):
        api.Compiler(language="cpu_openmp").compile(3, "variable_density",
                                                    "-DDOUBLE", "forward")


def _has_gpu():
    lib = ctypes.CDLL(os.path.join(LIB_DIR, "libsimwave_b200.so"))
    return lib.simwave_cuda_device_count() > 0


@pytest.mark.skipif(_has_gpu(), reason="only meaningful without a device")
def test_forward_fails_loudly_without_a_device():
    vel = np.full((20, 20), 1500.0, dtype=np.float32)
    sm = api.SpaceModel((0, 190, 0, 190), (10, 10), vel, space_order=2)
    tm = api.TimeModel(sm, tf=0.01)
    solver = api.Solver(sm, tm, api.Source(sm, [(90, 90)], 1),
                        api.Receiver(sm, [(90, 90)], 1),
                        api.RickerWavelet(10.0, tm))   # default compiler: cuda
    with pytest.raises(RuntimeError, match="forward failed"):
        solver.forward()
