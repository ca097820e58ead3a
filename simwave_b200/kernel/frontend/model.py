"""
Space and time models.

Host-side mirror of simwave/kernel/frontend/model.py: ``SpaceModel`` and
``TimeModel`` keep the reference's constructor signatures, properties and
validation.  They define every array the kernel receives (padded velocity /
density, damping mask, FD weights, dt, number of timesteps), so each value is
produced with the same NumPy operations on the same dtypes as the reference.
The code is written once for any dimension instead of per-2D/3D branches.
"""
import itertools

import numpy as np
from scipy.interpolate import RegularGridInterpolator

from simwave_b200.kernel.frontend import fd

_BC_NAMES = ('none', 'null_dirichlet', 'null_neumann')

# tokens of the extended-model arrays handed to the kernel (see
# SpaceModel.model_token); unique per process, never 0
_MODEL_TOKENS = itertools.count(1)


class SpaceModel:
    """
    Spatial model of the simulation.

    Parameters
    ----------
    bounding_box : tuple of float
        Minimum and maximum coordinates in meters of domain corners.
        e.g., (z_min, z_max, x_min, x_max [, y_min, y_max]).
    grid_spacing : tuple of float
        Grid spacing in meters in each axis (z, x [, y]).
    velocity_model : ndarray
        Numpy n-dimensional array with P wave velocity (m/s) profile.
    density_model : ndarray, optional
        Numpy n-dimensional array with the density profile.
    space_order : int, optional
        Spatial order of the stencil. Accepts even orders.
        Default is 2.
    dtype : data-type, optional
        Numpy array float data-type (numpy.float32 or numpy.float64).
        Default is numpy.float32.
    """
    def __init__(self, bounding_box, grid_spacing, velocity_model,
                 density_model=None, space_order=2, dtype=np.float32):
        self._dtype = dtype
        self._bounding_box = tuple(dtype(v) for v in bounding_box)
        self._grid_spacing = tuple(dtype(v) for v in grid_spacing)
        self._space_order = space_order

        if space_order % 2 != 0:
            raise ValueError(
                "Odd space order {} not supported".format(space_order)
            )
        if not 2 <= space_order <= 20:
            raise ValueError("Space order limited from 2 to 20.")

        self._dimension = len(velocity_model.shape)

        self._velocity_model = self.interpolate(velocity_model)
        self._density_model = None if density_model is None \
            else self.interpolate(density_model)
        # The resampled models belong to this object and never change after
        # construction (the reference has no setter either): read-only, so
        # that the extended arrays derived from them can be kept (below).
        self._velocity_model.flags.writeable = False
        if self._density_model is not None:
            self._density_model.flags.writeable = False
        self._extended = {}

    # ---- plain attributes -------------------------------------------------
    @property
    def bounding_box(self):
        return self._bounding_box

    @property
    def grid_spacing(self):
        return self._grid_spacing

    @property
    def velocity_model(self):
        return self._velocity_model

    @property
    def density_model(self):
        return self._density_model

    @property
    def space_order(self):
        return self._space_order

    @property
    def dimension(self):
        return self._dimension

    @property
    def dtype(self):
        return self._dtype

    # ---- grid geometry ----------------------------------------------------
    def _axis_bounds(self):
        """[(min, max)] per axis, from the flat bounding box."""
        b = self.bounding_box
        return [(b[2 * a], b[2 * a + 1]) for a in range(self.dimension)]

    @property
    def shape(self):
        """Grid points per axis: int((max - min) / h) + 1
        (reference model.py:94-112)."""
        return tuple(
            int((hi - lo) / h) + 1
            for (lo, hi), h in zip(self._axis_bounds(), self.grid_spacing)
        )

    @property
    def halo_size(self):
        """Halo width (stencil radius) on both edges of every axis."""
        return (self.space_order // 2,) * self.dimension * 2

    @property
    def nbl(self):
        """Damping-layer width in grid points on the edges of each axis."""
        return getattr(self, '_nbl', (0,) * self.dimension * 2)

    @property
    def damping_length(self):
        """Damping-layer length in meters on the edges of each axis."""
        return getattr(self, '_damping_length', (0.0,) * self.dimension * 2)

    @property
    def boundary_condition(self):
        """Boundary condition name on the edges of each axis."""
        return getattr(self, '_boundary_condition',
                       ('none',) * self.dimension * 2)

    @property
    def damping_polynomial_degree(self):
        """Degree of the polynomial damping profile."""
        return getattr(self, '_damping_polynomial_degree', 3)

    @property
    def damping_alpha(self):
        """Scale of the polynomial damping profile."""
        return getattr(self, '_damping_alpha', 0.001)

    @property
    def nbl_pad_width(self):
        """Damping-layer widths in ``numpy.pad`` format."""
        return tuple(zip(self.nbl[::2], self.nbl[1::2]))

    @property
    def halo_pad_width(self):
        """Halo widths in ``numpy.pad`` format."""
        return tuple(zip(self.halo_size[::2], self.halo_size[1::2]))

    @property
    def extended_shape(self):
        """Grid shape including damping layers and halos
        (reference model.py:115-133)."""
        return tuple(
            n + nb + na + hb + ha
            for n, (nb, na), (hb, ha) in zip(self.shape, self.nbl_pad_width,
                                             self.halo_pad_width)
        )

    @property
    def grid(self):
        """Zero grid of the un-extended shape."""
        return np.zeros(self.shape, dtype=self.dtype)

    @property
    def extended_grid(self):
        """Zero grid of the extended shape."""
        return self._pad(self.grid, mode="constant")

    def fd_coefficients(self, derivative_order):
        """Centre + right-half FD weights in the model dtype
        (reference model.py:192-208)."""
        return self.dtype(
            fd.half_coefficients(derivative_order, self.space_order)
        )

    # ---- model preparation ------------------------------------------------
    def interpolate(self, data):
        """
        Linearly resample ``data`` (velocity or density) onto the grid that
        covers the bounding box with ``grid_spacing``
        (reference model.py:210-261: RegularGridInterpolator on a meshgrid).
        """
        if tuple(data.shape) == self.shape:
            # The model already lives on the target grid: both sets of
            # coordinates are the same linspace, every target point is a node
            # of the source grid, and linear interpolation at a node returns
            # the node's value exactly (weights 1 and 0).  Same bits as the
            # interpolant below, without its O(ndim * N) float64 passes
            # (minutes at 1024^3).
            return np.array(data, dtype=self.dtype)
        bounds = self._axis_bounds()
        source_axes = tuple(
            np.linspace(lo, hi, n) for (lo, hi), n in zip(bounds, data.shape)
        )
        interpolant = RegularGridInterpolator(source_axes, data)

        target_axes = [
            np.linspace(lo, hi, n) for (lo, hi), n in zip(bounds, self.shape)
        ]
        # The interpolant is evaluated point by point, so the target grid can
        # be walked in blocks of leading-axis planes with identical results:
        # the float64 coordinate meshes of a whole 1024^3 grid (what the
        # reference builds, and why it runs out of memory there, SURVEY.md
        # section 7.3) never exist.
        out = np.empty(self.shape, dtype=self.dtype)
        plane = int(np.prod(self.shape[1:]))
        rows = max(1, self._interp_block_bytes // (plane * 8 * (self.dimension + 4)))
        for lo in range(0, self.shape[0], rows):
            hi = min(self.shape[0], lo + rows)
            mesh = np.meshgrid(target_axes[0][lo:hi], *target_axes[1:],
                               indexing='ij')
            out[lo:hi] = interpolant(tuple(mesh))
        return out

    # working-set bound of one interpolation block (float64 meshes + result)
    _interp_block_bytes = 256 << 20

    def config_boundary(self, damping_length=0.0, boundary_condition="none",
                        damping_polynomial_degree=3, damping_alpha=0.001):
        """
        Configure the absorbing layers and the boundary conditions
        (reference model.py:271-349).

        Parameters
        ----------
        damping_length : float or tuple of float, optional
            Layer length in meters on the edges of each axis, e.g.
            (z_before, z_after, x_before, x_after [, y_before, y_after]);
            a scalar applies to every edge. Default is 0.
        boundary_condition : str or tuple of str
            none, null_dirichlet or null_neumann per edge (same order);
            a str applies to every edge. Default is none.
        damping_polynomial_degree : int, optional
            Degree of the damping polynomial. Default is 3.
        damping_alpha : float, optional
            Scale of the damping polynomial. Default is 0.001.
        """
        edges = self.dimension * 2
        self._damping_polynomial_degree = damping_polynomial_degree
        self._damping_alpha = damping_alpha

        if isinstance(damping_length, (float, int)):
            self._damping_length = (self.dtype(damping_length),) * edges
        else:
            self._damping_length = tuple(
                self.dtype(v) for v in damping_length
            )

        if isinstance(boundary_condition, str):
            self._boundary_condition = (boundary_condition,) * edges
        else:
            self._boundary_condition = boundary_condition

        for bc in self.boundary_condition:
            if bc not in _BC_NAMES:
                raise ValueError(
                    'Boundary condition {} not available.'.format(
                        self.boundary_condition
                    )
                )

        # meters -> grid points, truncating, with the axis' own spacing
        spacing_per_edge = [h for h in self.grid_spacing for _ in (0, 1)]
        lengths = tuple(self._damping_length)
        if len(lengths) != edges:
            raise ValueError(
                "not enough values to unpack (expected {}, got {})".format(
                    edges, len(lengths))
            )
        self._nbl = tuple(
            int(length / h) for length, h in zip(lengths, spacing_per_edge)
        )

    def _pad(self, array, mode):
        """Damping-layer pad followed by halo pad with the same mode
        ('edge' or 'constant'): padding twice with one of these modes equals
        padding once by the summed widths, which makes one copy of the
        extended array instead of two (4.5 GB each at 1040^3)."""
        widths = tuple(
            (nb + hb, na + ha) for (nb, na), (hb, ha)
            in zip(self.nbl_pad_width, self.halo_pad_width)
        )
        return np.pad(array=array, pad_width=widths, mode=mode)

    def _kept(self, name, build):
        """The extended array ``name``, built once per boundary configuration.

        The reference rebuilds these arrays on every access (twice per
        ``Solver.forward`` for the velocity of a 10^8-point model: seconds of
        ``numpy.pad``).  They only depend on the (read-only) models and on
        what ``config_boundary`` set, so they are kept, read-only, until that
        changes; ``model_token`` changes with them."""
        signature = (self.nbl, self.halo_size, self.damping_polynomial_degree,
                     self.damping_alpha)
        entry = self._extended.get(name)
        if entry is None or entry[0] != signature:
            array = build()
            if array is not None:
                array.flags.writeable = False
            self._extended[name] = entry = (signature, array)
            self._model_token = next(_MODEL_TOKENS)
        return entry[1]

    @property
    def model_token(self):
        """Identifies the current set of extended arrays (velocity, density,
        damping mask): a new value whenever one of them is rebuilt.  The CUDA
        backend keeps the preprocessed model of the last ``forward`` on the
        device under this token (SIMWAVE_HINT_MODEL_RESIDENT), so a survey of
        many shots over one SpaceModel uploads the model once."""
        self.extended_velocity_model, self.extended_density_model, self.damping_mask
        return self._model_token

    @property
    def damping_mask(self):
        """
        Damping coefficient per extended grid point: zero in the physical
        domain and in the halo, ``alpha * d**degree`` in the layers, d being
        the distance in grid points from the physical domain
        (reference model.py:378-406).  Read-only; kept between accesses.
        """
        def build():
            # The reference pads a zero array of the whole grid with linear
            # ramps, axis after axis, then raises it to the degree: several
            # passes over 10^9 points for a 1024^3 model.  A ramp value only
            # depends on how deep the point sits in the layer of every axis
            # (numpy ramps from the end value to the EDGE value, which itself
            # only depends on the depths along the axes padded before), not
            # on the size of the physical domain.  So the same numpy calls on
            # a ONE-cell domain give every distinct value, bit for bit, and
            # the full mask is a gather from that small array.
            widths = self.nbl_pad_width
            small = np.pad(
                array=np.zeros((1,) * self.dimension, dtype=self.dtype),
                pad_width=widths,
                mode="linear_ramp",
                end_values=widths
            )
            small = (small ** self.damping_polynomial_degree) * self.damping_alpha
            # index into `small` of every grid index along each axis
            index = []
            for n, (before, after) in zip(self.shape, widths):
                i = np.full(before + n + after, before, dtype=np.intp)
                i[:before] = np.arange(before)
                i[before + n:] = before + 1 + np.arange(after)
                index.append(i)
            halo = self.halo_pad_width
            shape = tuple(len(i) + hb + ha for i, (hb, ha) in zip(index, halo))
            mask = np.zeros(shape, dtype=small.dtype)
            core = mask[tuple(slice(hb, s - ha) for s, (hb, ha) in zip(shape, halo))]
            # slowest axis: every index of the physical domain shares one
            # hyperplane, the layer indices have their own
            rest = np.ix_(*index[1:]) if self.dimension > 1 else ()
            before0, n0 = widths[0][0], self.shape[0]
            core[before0:before0 + n0] = small[(before0,) + rest]
            for k in list(range(before0)) + list(range(before0 + n0, len(index[0]))):
                core[k] = small[(index[0][k],) + rest]
            return mask
        return self._kept('damping_mask', build)

    @property
    def extended_velocity_model(self):
        """Velocity edge-padded over layers and halo
        (reference model.py:421-441).  Read-only; kept between accesses."""
        return self._kept('velocity',
                          lambda: self._pad(self.velocity_model, mode="edge"))

    @property
    def extended_density_model(self):
        """Density edge-padded over layers and halo, or None
        (reference model.py:444-466).  Read-only; kept between accesses."""
        if self.density_model is None:
            return None
        return self._kept('density',
                          lambda: self._pad(self.density_model, mode="edge"))

    # ---- output trimming --------------------------------------------------
    def _trim(self, u, widths):
        """Drop ``widths`` = ((before, after), ..) points from the spatial
        axes of ``u`` (axis 0 is the snapshot axis).  A zero 'after' width
        yields an empty axis, as the reference's ``-0`` slice does."""
        if self.dimension not in (2, 3):
            raise Exception("Wavefield dimension not supported.")
        index = (slice(None),) + tuple(
            slice(before, -after) for before, after in widths
        )
        return u[index]

    def remove_halo_region(self, u):
        """Strip the stencil halo from a wavefield with snapshots
        (reference model.py:468-492)."""
        halo = self.halo_size[0]
        return self._trim(u, ((halo, halo),) * self.dimension)

    def remove_nbl(self, u):
        """Strip the damping layers from a wavefield with snapshots
        (reference model.py:494-515)."""
        return self._trim(u, self.nbl_pad_width)


class TimeModel:
    """
    Time axis of the simulation.

    Parameters
    ----------
    space_model : object
        Space model object.
    tf : float
        End time in seconds.
    dt : float. optional
        Timestep variation in seconds.
    t0 : float, optional
        Start time in seconds. Default is 0.0.
    saving_stride : int
        Skipping factor when saving the wavefields.
        If saving_stride is 0, only the last wavefield is saved. Default is 0.
    """
    def __init__(self, space_model, tf, dt=None, t0=0.0, saving_stride=0):
        self._space_model = space_model
        self._tf = space_model.dtype(tf)
        self._t0 = space_model.dtype(t0)
        self._saving_stride = saving_stride

        # CFL limit first; a user dt may only lower it
        self._dt = fd.calculate_dt(
            dimension=space_model.dimension,
            space_order=space_model.space_order,
            grid_spacing=space_model.grid_spacing,
            velocity_model=space_model.velocity_model
        )
        if dt is not None:
            self.dt = dt

        if not (0 <= self.saving_stride <= self.timesteps):
            raise Exception(
                "Saving jumps can not be less than zero or "
                "greater than the number of timesteps."
            )

    @property
    def space_model(self):
        """Corresponding space model."""
        return self._space_model

    @property
    def tf(self):
        """End time value in seconds."""
        return self._tf

    @property
    def t0(self):
        """Initial time value in seconds."""
        return self._t0

    @property
    def saving_stride(self):
        """Skipping factor when saving the wavefields."""
        return self._saving_stride

    @property
    def dtype(self):
        return self.space_model.dtype

    @property
    def dt(self):
        """Time step in seconds, in the model dtype."""
        return self.dtype(self._dt)

    @dt.setter
    def dt(self, value):
        if value < 0:
            raise ValueError("Time step cannot be negative.")
        if value > self.dt:
            raise ValueError("Time step value violates CFL condition.")
        self._dt = value

    @property
    def timesteps(self):
        """Number of timesteps: ceil((tf - t0 + dt) / dt), then raised until
        ``timesteps % saving_stride == 1`` when the stride is above 1
        (reference model.py:600-609)."""
        count = int(np.ceil((self.tf - self.t0 + self.dt) / self.dt))
        stride = self.saving_stride
        if 1 < stride <= count:
            count += (1 - count) % stride
        return count

    @property
    def time_indexes(self):
        """Time indexes 0 .. timesteps-1."""
        return np.linspace(0, self.timesteps - 1, self.timesteps,
                           dtype=np.uint)

    @property
    def time_values(self):
        """Time values from t0 to tf, one per timestep."""
        return np.linspace(self.t0, self.tf, self.timesteps, dtype=self.dtype)

    def remove_time_halo_region(self, u):
        """
        Keep only the saved snapshots of the slot array returned by the
        kernel (reference model.py:623-643): with ``saving_stride == 0`` the
        last field lives in slot ``timesteps % 3``; otherwise the first and
        last slots are time halos.
        """
        if self.saving_stride == 0:
            last = self.timesteps % 3
            return u[last:last + 1]
        return u[1:-1]
