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

    @staticmethod
    def _table_arguments(prefix, acquisition):
        """The five kernel arguments that describe the interpolation windows
        of sources ('src') or receivers ('rec'): interval table, weights,
        per-point offsets and the two sizes (reference solver.py:143-152)."""
        points, values, offsets = acquisition.interpolated_points_and_values
        return {
            prefix + '_points_interval': points,
            prefix + '_points_interval_size': len(points),
            prefix + '_points_values': values,
            prefix + '_points_values_offset': offsets,
            prefix + '_points_values_size': len(values),
        }

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
        return self._run('forward')

    def adjoint(self, shot_record):
        """
        Apply the adjoint of the forward operator to a shot record (new; the
        reference stops at ``forward``).  With F the linear map from the source
        wavelets to the shot record of ``forward()``, returns F^T applied to
        ``shot_record``: what a gradient computation correlates with the
        forward wavefield.  Constant density, ``saving_stride == 0``.

        Parameters
        ----------
        shot_record : ndarray
            Data of shape (timesteps, receivers), e.g. a residual.

        Returns
        ----------
        ndarray
            Last adjoint wavefield without time and space halos.
        ndarray
            Adjoint source of shape (timesteps, sources), or (timesteps,)
            for a single source.
        """
        expected = (self.time_model.timesteps, self.receivers.count)
        data = np.ascontiguousarray(shot_record, dtype=self.space_model.dtype)
        if data.shape != expected:
            raise ValueError("shot_record must have shape {}".format(expected))
        if self.time_model.saving_stride != 0:
            raise ValueError("adjoint needs saving_stride == 0")
        return self._run('adjoint', data)

    def _run(self, operator, shot_record=None):
        space, time = self.space_model, self.time_model
        u_full = self.u_full
        if operator == 'adjoint':
            count = self.sources.count
            wavelet = np.zeros((time.timesteps, count) if count > 1
                               else (time.timesteps,), dtype=space.dtype)
            wavelet_count = count
            wavelet_size = time.timesteps
        else:
            shot_record = self.shot_record
            wavelet = self.wavelet.values
            wavelet_count = self.wavelet.num_sources
            wavelet_size = self.wavelet.timesteps

        # keyword names are the ones Middleware.exec unpacks into the ABI
        # argument list (middleware.py, _keys_in_order)
        arguments = {
            'u_full': u_full,
            'shot_record': shot_record,
            'num_snapshots': u_full.shape[0],
            # model
            'velocity_model': space.extended_velocity_model,
            'density_model': space.extended_density_model,
            'damping_mask': space.damping_mask,
            'boundary_condition': space.boundary_condition,
            'grid_spacing': space.grid_spacing,
            'space_order': space.space_order,
            'second_order_fd_coefficients': space.fd_coefficients(2),
            'first_order_fd_coefficients': space.fd_coefficients(1),
            # time axis
            'dt': time.dt,
            'saving_stride': time.saving_stride,
            'begin_timestep': 1,
            'end_timestep': time.timesteps,
            # acquisition
            'wavelet': wavelet,
            'wavelet_size': wavelet_size,
            'wavelet_count': wavelet_count,
            'num_sources': self.sources.count,
            'num_receivers': self.receivers.count,
        }
        arguments.update(self._table_arguments('src', self.sources))
        arguments.update(self._table_arguments('rec', self.receivers))

        # What this call knows about its own arrays, for the CUDA backend's
        # data path (include/simwave_cuda.h, simwave_cuda_set_hint): u_full is
        # freshly allocated zeros; with saving_stride == 0 only slot
        # timesteps % 3 is handed back below; the extended model arrays are
        # the SpaceModel's read-only ones, unchanged while its token is.
        self._middleware.hints = {
            'wavefield_in_zero': 1,
            'wavefield_out': 1 if time.saving_stride == 0 else 0,
            'model_resident': space.model_token,
        }

        u_full, recv = self._middleware.exec(operator=operator, **arguments)

        u_full = time.remove_time_halo_region(u_full)
        u_full = space.remove_halo_region(u_full)
        return u_full, recv


def _read_only(attribute, doc):
    return property(lambda self: getattr(self, '_' + attribute), doc=doc)


# the constructor arguments, readable under the reference's names
for _name, _doc in (('space_model', 'Space model object.'),
                    ('time_model', 'Time model object.'),
                    ('sources', 'Source object.'),
                    ('receivers', 'Receiver object.'),
                    ('wavelet', 'Wavelet object.'),
                    ('compiler', 'Compiler object.')):
    setattr(Solver, _name, _read_only(_name, _doc))
del _name, _doc
