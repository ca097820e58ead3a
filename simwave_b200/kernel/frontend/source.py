"""
Sources, receivers and wavelets.

Host-side mirror of simwave/kernel/frontend/source.py: same class names,
constructor signatures, properties and error behaviour.  The coordinate ->
grid-position -> interpolation-table pipeline is vectorised over all
sources, but keeps the reference's arithmetic (same dtype per operation), so
``interpolated_points_and_values`` is bit-identical to the reference in the
same NumPy/SciPy environment.
"""
import numpy as np

from simwave_b200.kernel.frontend import kws


class Source:
    """
    A set of point sources (or receivers) placed at physical coordinates.

    Parameters
    ----------
    space_model : SpaceModel
        Space model object.
    coordinates : list of tuple, list of list or ndarray
        Physical coordinates (in meters), one row per source.
    window_radius : int, optional
        Half-width of the Kaiser-windowed sinc. Default is 4.
    """
    def __init__(self, space_model, coordinates, window_radius=4):
        self._space_model = space_model
        self._window_radius = window_radius

        if not isinstance(coordinates, (list, tuple, np.ndarray)):
            raise ValueError("Source/Receiver coordinates must be "
                             "represented as lists, tuples or ndarrays.")

        coords = np.asarray(coordinates, dtype=space_model.dtype)
        if coords.ndim == 1:
            # a single position given as a flat sequence
            coords = coords.reshape((1,) + coords.shape)
        elif coords.ndim != 2 or coords.shape[1] not in (2, 3):
            raise ValueError("Invalid source/receiver coordinates format.")
        self._coordinates = coords

    @property
    def space_model(self):
        """Corresponding space model."""
        return self._space_model

    @property
    def coordinates(self):
        """Physical coordinates (in meters), shape (count, dimension)."""
        return self._coordinates

    @property
    def window_radius(self):
        """Half-width of the Kaiser-windowed sinc."""
        return self._window_radius

    @property
    def count(self):
        """Number of sources/receivers."""
        return self.coordinates.shape[0]

    @property
    def grid_positions(self):
        """Positions in grid points relative to the bounding-box origin
        (reference source.py:63-103)."""
        dtype = self.space_model.dtype
        coords = self.coordinates
        ndim = coords.shape[1]
        if ndim not in (2, 3):
            raise Exception("Dimension %d not supported." % ndim)

        bbox = self.space_model.bounding_box
        lower = np.array(bbox[0:2 * ndim:2], dtype=dtype)
        upper = np.array(bbox[1:2 * ndim:2], dtype=dtype)
        spacing = np.array(self.space_model.grid_spacing[:ndim], dtype=dtype)

        outside = np.any((coords < lower) | (coords > upper), axis=1)
        if outside.any():
            bad = coords[int(np.argmax(outside))]
            raise Exception("Coordinates %s out of bounds." % bad)

        # same two float ops per element as the reference: (c - min) / h
        return np.asarray((coords - lower) / spacing, dtype=dtype)

    @property
    def adjusted_grid_positions(self):
        """Grid positions shifted by the damping layer and the halo on the
        'before' side of every axis (reference source.py:106-121)."""
        dtype = self.space_model.dtype
        before_nbl = self.space_model.nbl[::2]
        before_halo = self.space_model.halo_size[::2]
        origin = np.array([n + h for n, h in zip(before_nbl, before_halo)],
                          dtype=dtype)
        return np.asarray(self.grid_positions + origin, dtype=dtype)

    @property
    def interpolated_points_and_values(self):
        """
        Interpolation tables consumed by the kernel
        (reference source.py:124-159).

        Returns
        ----------
        ndarray
            uint64 ``[b_axis1, e_axis1, .., b_axisN, e_axisN]`` per source.
        ndarray
            Weights ``[axis1 values, .., axisN values]`` per source, in the
            space model's dtype.
        ndarray
            uint64 running offsets into the weights, length count + 1.
        """
        points, weights, offsets = kws.get_source_points_batch(
            self.space_model.extended_shape, self.adjusted_grid_positions,
            self.window_radius)
        return points, weights.astype(self.space_model.dtype), offsets


# a receiver is positioned and interpolated exactly like a source
Receiver = Source


class Wavelet:
    """
    A source time function given as a callable.

    Parameters
    ----------
    function : object
        Function (expression) that creates the wavelet.
    kwargs : dict
        key word arguments of the function.
    """
    def __init__(self, function, **kwargs):
        self._function = function
        self._kwargs = kwargs

    @property
    def function(self):
        """Function (expression) that creates the wavelet."""
        return self._function

    @property
    def kwargs(self):
        """key word arguments of the function."""
        return self._kwargs

    @property
    def values(self):
        """Wavelet samples, one per timestep."""
        return self.function(**self.kwargs)

    @property
    def num_sources(self):
        """Number of independent wavelets (1: shared by all sources)."""
        return 1

    @property
    def timesteps(self):
        """Number of timesteps."""
        return len(self.values)


class RickerWavelet(Wavelet):
    """
    Ricker wavelet sampled on a time model (reference source.py:212-257).

    Parameters
    ----------
    peak_frequency : float
        Peak frequency for the wavelet in Hz.
    time_model: TimeModel
        Time model object.
    amplitude : float, optional
        Amplitude of the wavelet. Default is 1.0.
    """
    def __init__(self, peak_frequency, time_model, amplitude=1):
        self._peak_frequency = peak_frequency
        self._time_model = time_model
        self._amplitude = amplitude
        super().__init__(
            self._ricker,
            peak_frequency=peak_frequency,
            time_model=time_model,
            amplitude=amplitude
        )

    @property
    def peak_frequency(self):
        """Peak frequency of the wavelet in Hz."""
        return self._peak_frequency

    @property
    def time_model(self):
        """Corresponding time model."""
        return self._time_model

    @property
    def amplitude(self):
        """Amplitude of the wavelet."""
        return self._amplitude

    @staticmethod
    def _ricker(peak_frequency, time_model, amplitude):
        delay = 1 / peak_frequency
        arg = np.pi * peak_frequency * (time_model.time_values - delay)
        return amplitude * (1 - 2.0 * arg**2) * np.exp(-arg**2)


class MultiWavelet(Wavelet):
    """
    One wavelet per source (reference source.py:260-302).

    Parameters
    ----------
    values : ndarray
        Numpy array [timesteps][sources]
    time_model: TimeModel
        Time model object.
    """
    def __init__(self, values, time_model):
        self._values = values
        self._time_model = time_model

        if self.timesteps != self.time_model.timesteps:
            # the reference builds this message with a misplaced .format and
            # dies with AttributeError (source.py:275-278); the intended
            # ValueError is raised here
            raise ValueError("Wavelet must have {} timesteps.".format(
                self.time_model.timesteps
            ))

    @property
    def values(self):
        """Wavelet samples, shape (timesteps, sources), C-contiguous."""
        return np.ascontiguousarray(self._values, dtype=self.dtype)

    @property
    def num_sources(self):
        """Number of sources."""
        return self.values.shape[1]

    @property
    def timesteps(self):
        """Number of timesteps."""
        return self.values.shape[0]

    @property
    def time_model(self):
        """Corresponding time model."""
        return self._time_model

    @property
    def dtype(self):
        return self.time_model.dtype
