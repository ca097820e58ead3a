from simwave_b200.kernel.backend import Compiler, Middleware
from simwave_b200.kernel.frontend import (
    SpaceModel, TimeModel, Source, Receiver, Wavelet, RickerWavelet,
    MultiWavelet, Solver
)

__all__ = [
    "Compiler", "Middleware", "SpaceModel", "TimeModel", "Source",
    "Receiver", "Wavelet", "RickerWavelet", "MultiWavelet", "Solver",
]
