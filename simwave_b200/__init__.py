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

__all__ = [
    "Compiler", "SpaceModel", "TimeModel", "Source", "Receiver", "Wavelet",
    "RickerWavelet", "MultiWavelet", "Solver", "plot_wavefield",
    "plot_shotrecord", "plot_velocity_model", "plot_wavelet", "read_2D_segy"
]
