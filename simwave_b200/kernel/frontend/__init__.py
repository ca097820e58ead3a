from simwave_b200.kernel.frontend.model import SpaceModel, TimeModel
from simwave_b200.kernel.frontend.source import (
    Source, Receiver, Wavelet, RickerWavelet, MultiWavelet
)
from simwave_b200.kernel.frontend.solver import Solver

__all__ = [
    "SpaceModel", "TimeModel", "Source", "Receiver", "Wavelet",
    "RickerWavelet", "MultiWavelet", "Solver"
]
