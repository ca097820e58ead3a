"""
simwave_b200 -- B200-native (sm_100a) backend for simwave's acoustic forward
modelling time loop, behind simwave's own Python API.

The public names are the ones simwave/__init__.py:26-41 exports.  Everything
on the hot path (Solver.forward -> Middleware -> ``forward`` C-ABI -> CUDA
time loop) lives here; plotting and SEG-Y ingest are out of scope and only
kept importable (their optional dependencies are imported on first use).
"""
from simwave_b200.kernel import (
    Compiler, Middleware, SpaceModel, TimeModel, Source, Receiver, Wavelet,
    RickerWavelet, MultiWavelet, Solver
)
from simwave_b200.extras import (
    read_2D_segy, plot_wavefield, plot_shotrecord, plot_velocity_model,
    plot_wavelet
)

__version__ = "0.1.0"

# The front end's bit-exactness claims (tables equal to the reference's) hold
# under NumPy >= 2 promotion rules (NEP 50): `float32 array - float64 scalar`
# stays float32 there as it did for Python scalars before; the batched table
# builder subtracts an array of positions and would promote differently from
# the one-by-one path under NumPy 1.x value-based casting.
import numpy as _np
if int(_np.__version__.split(".")[0]) < 2:          # pragma: no cover
    import warnings as _warnings
    _warnings.warn("simwave_b200 is validated with numpy >= 2; interpolation "
                   "tables of float64 models may differ in the last bit from "
                   "simwave's under numpy %s" % _np.__version__)

__all__ = [
    "Compiler", "SpaceModel", "TimeModel", "Source", "Receiver", "Wavelet",
    "RickerWavelet", "MultiWavelet", "Solver", "plot_wavefield",
    "plot_shotrecord", "plot_velocity_model", "plot_wavelet", "read_2D_segy"
]
