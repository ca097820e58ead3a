"""
Acoustic forward solver front end.

Host-side mirror of simwave/kernel/frontend/solver.py: ``Solver`` keeps the
constructor signature and ``forward()`` contract; it gathers the kernel
arguments (same keyword names the reference hands to its Middleware,
solver.py:129-161) and strips the time and space halos from the result.
"""
import numpy as np

from simwave_b200.kernel.backend.middleware import Middleware


class Solver:
    """
    Acoustic solver for the simulation.

    Parameters
    ----------
    space_model : SpaceModel
        Space model object.
    time_model: TimeModel
        Time model object.
    sources : Source
        Source object.
    receivers : Receiver
        Receiver object.
    wavelet : Wavelet
        Wavelet object.
    compiler : Compiler
        Backend compiler object. ``None`` selects the prebuilt CUDA backend.
    """
    def __init__(self, space_model, time_model, sources,
                 receivers, wavelet, compiler=None):
        self._space_model = space_model
        self._time_model = time_model
        self._sources = sources
        self._receivers = receivers
        self._wavelet = wavelet
        self._compiler = compiler
        self._middleware = Middleware(compiler=compiler)

    @property
    def space_model(self):
        """Space model object."""
        return self._space_model

    @property
    def time_model(self):
        """Time model object."""
        return self._time_model

    @property
    def sources(self):
        """Source object."""
        return self._sources

    @property
    def receivers(self):
        """Receiver object."""
        return self._receivers

    @property
    def wavelet(self):
        """Wavelet object."""
        return self._wavelet

    @property
    def compiler(self):
        """Compiler object."""
        return self._compiler

    @property
    def snapshot_indexes(self):
        """Time indexes of the wavefields that are kept
        (reference solver.py:68-83)."""
        stride = self.time_model.saving_stride
        if stride == 0:
            return [self.time_model.time_indexes[-1]]
        first = self.time_model.time_indexes[0]
        return list(range(first, self.time_model.timesteps, stride))

    @property
    def num_snapshots(self):
        """Number of wavefields that are kept."""
        return len(self.snapshot_indexes)

    @property
    def shot_record(self):
        """Fresh zero shot record (timesteps, receivers)."""
        return np.zeros(
            shape=(self.time_model.timesteps, self.receivers.count),
            dtype=self.space_model.dtype
        )

    @property
    def u_full(self):
        """Fresh zero slot array (snapshots + 2, nz, nx [, ny]); the two
        extra slots are the time halo of the 2nd-order scheme."""
        shape = (self.num_snapshots + 2,) + self.space_model.extended_shape
        return np.zeros(shape, dtype=self.space_model.dtype)

    def forward(self):
        """
        Run the forward propagator.

        Returns
        ----------
        ndarray
            Wavefield snapshots without time and space halos.
        ndarray
            Shot record.
        """
        space, time = self.space_model, self.time_model

        src_points, src_values, src_offsets = \
            self.sources.interpolated_points_and_values
        rec_points, rec_values, rec_offsets = \
            self.receivers.interpolated_points_and_values

        u_full = self.u_full

        u_full, recv = self._middleware.exec(
            operator='forward',
            u_full=u_full,
            velocity_model=space.extended_velocity_model,
            density_model=space.extended_density_model,
            damping_mask=space.damping_mask,
            wavelet=self.wavelet.values,
            wavelet_size=self.wavelet.timesteps,
            wavelet_count=self.wavelet.num_sources,
            second_order_fd_coefficients=space.fd_coefficients(2),
            first_order_fd_coefficients=space.fd_coefficients(1),
            boundary_condition=space.boundary_condition,
            src_points_interval=src_points,
            src_points_interval_size=len(src_points),
            src_points_values=src_values,
            src_points_values_offset=src_offsets,
            src_points_values_size=len(src_values),
            rec_points_interval=rec_points,
            rec_points_interval_size=len(rec_points),
            rec_points_values=rec_values,
            rec_points_values_offset=rec_offsets,
            rec_points_values_size=len(rec_values),
            shot_record=self.shot_record,
            num_sources=self.sources.count,
            num_receivers=self.receivers.count,
            grid_spacing=space.grid_spacing,
            saving_stride=time.saving_stride,
            dt=time.dt,
            begin_timestep=1,
            end_timestep=time.timesteps,
            space_order=space.space_order,
            num_snapshots=u_full.shape[0]
        )

        u_full = time.remove_time_halo_region(u_full)
        u_full = space.remove_halo_region(u_full)
        return u_full, recv
