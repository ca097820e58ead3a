"""
Python side of the ``forward`` C-ABI.

Host-side mirror of simwave/kernel/backend/middleware.py: converts the
boundary-condition names to codes, unpacks shape and spacing, orders the
arguments the way the C function expects them (middleware.py:166-202), maps
Python/NumPy values to ctypes (middleware.py:261-285) and calls ``forward``
from the library chosen by the ``Compiler``.

Additions for the CUDA backend: ``forward`` returns a negative number on
failure and the library exports ``simwave_cuda_last_error()``; that is turned
into a Python exception here instead of the reference CUDA path's ``exit()``
(constant_density/3d/cuda/wave.cu:13-20).
"""
import ctypes

import numpy as np
from numpy.ctypeslib import ndpointer

from simwave_b200.kernel.backend.compiler import Compiler

# argument order of every `forward` variant; absent (None) entries are
# skipped, which is how one list serves 2D/3D and constant/variable density
ARGUMENT_ORDER = (
    'u_full', 'velocity_model', 'density_model', 'damping_mask',
    'wavelet', 'wavelet_size', 'wavelet_count',
    'second_order_fd_coefficients', 'first_order_fd_coefficients',
    'boundary_condition',
    'src_points_interval', 'src_points_interval_size',
    'src_points_values', 'src_points_values_size', 'src_points_values_offset',
    'rec_points_interval', 'rec_points_interval_size',
    'rec_points_values', 'rec_points_values_size', 'rec_points_values_offset',
    'shot_record', 'num_sources', 'num_receivers',
    'nz', 'nx', 'ny', 'dz', 'dx', 'dy',
    'saving_stride', 'dt', 'begin_timestep', 'end_timestep',
    'space_order', 'num_snapshots',
)

_BC_CODE = {'none': 0, 'null_dirichlet': 1, 'null_neumann': 2}

_CTYPE = {
    'int': ctypes.c_size_t,
    'float': ctypes.c_float,
    'float32': ctypes.c_float,
    'float64': ctypes.c_double,
    'np(uint64)': ndpointer(ctypes.c_size_t, flags="C_CONTIGUOUS"),
    'np(float32)': ndpointer(ctypes.c_float, flags="C_CONTIGUOUS"),
    'np(float64)': ndpointer(ctypes.c_double, flags="C_CONTIGUOUS"),
}

_PRECISION_MACRO = {'float32': '-DFLOAT', 'float64': '-DDOUBLE'}

# include/simwave_cuda.h: SIMWAVE_HINT_*
_HINT_CODE = {'wavefield_in_zero': 1, 'wavefield_out': 2, 'model_resident': 3}


class Middleware:
    """
    Communication interface between frontend and backend.

    Parameters
    ----------
    compiler : Compiler
        Compiler object. ``None`` selects ``Compiler(language='cuda')``,
        the only backend this package ships.
    """
    def __init__(self, compiler):
        self._compiler = Compiler(language='cuda') if compiler is None \
            else compiler
        # promises of the caller about the arrays of the next exec(), by name
        # (see _HINT_CODE); applied around the call and withdrawn after it
        self.hints = {}

    @property
    def compiler(self):
        return self._compiler

    def library(self, dimension, density, dtype, operator="forward"):
        """Load and return the library that exports the operator."""
        shared_object = self.compiler.compile(
            dimension=dimension,
            density="constant_density" if density is None
            else "variable_density",
            float_precision=_PRECISION_MACRO[str(dtype)],
            operator=operator
        )
        return ctypes.cdll.LoadLibrary(shared_object)

    def exec(self, operator, **kwargs):
        """
        Run an operator.

        Parameters
        ----------
        operator : str
            operator to be executed.
        kwargs : dict
            List of keyword arguments.

        Returns
        ----------
        tuple
            The operation results
        """
        kwargs['boundary_condition'] = self._convert_boundary_condition(
            kwargs.get('boundary_condition')
        )

        # constant density: the kernel takes neither the density model nor
        # the first-derivative weights
        if kwargs.get('density_model') is None:
            kwargs.pop('density_model', None)
            kwargs.pop('first_order_fd_coefficients', None)

        shape = kwargs['velocity_model'].shape
        spacing = kwargs.pop('grid_spacing')
        if len(shape) not in (2, 3) or len(spacing) != len(shape):
            raise ValueError("Grid must be 2D or 3D with one spacing per axis.")
        kwargs.update(zip(('nz', 'nx', 'ny'), shape))
        kwargs.update(zip(('dz', 'dx', 'dy'), spacing))

        if operator == 'forward':
            return self._exec_forward(**kwargs)
        if operator == 'adjoint':
            return self._exec_forward(_symbol='adjoint', **kwargs)
        raise ValueError("Operator {} not available.".format(operator))

    def _exec_forward(self, _symbol='forward', **kwargs):
        """
        Run the forward operator; returns (u_full, shot_record), the same
        arrays that were passed in, updated in place by the kernel.

        ``_symbol='adjoint'`` runs the adjoint operator through the same
        argument list (include/simwave_cuda.h section 1b): ``shot_record`` is
        then the input and ``wavelet`` is updated in place; returns
        (u_full, wavelet).
        """
        velocity = kwargs['velocity_model']
        lib = self.library(
            dimension=velocity.ndim,
            density=kwargs.get('density_model'),
            dtype=velocity.dtype,
            operator=_symbol
        )

        types = self._argtypes(**kwargs)
        keys = [k for k in ARGUMENT_ORDER if kwargs.get(k) is not None]

        forward = getattr(lib, _symbol)
        forward.restype = ctypes.c_double
        forward.argtypes = [types[k] for k in keys]

        hinted = self._apply_hints(lib, self.hints)
        try:
            exec_time = forward(*[kwargs[k] for k in keys])
        finally:
            self._apply_hints(lib, dict.fromkeys(hinted, 0))
            self.hints = {}

        if exec_time < 0:
            raise RuntimeError(
                "{} failed: {}".format(_symbol, self._last_error(lib))
            )

        print('Run %s in %f seconds.' % (_symbol, exec_time))

        if _symbol == 'adjoint':
            return kwargs.get('u_full'), kwargs.get('wavelet')
        return kwargs.get('u_full'), kwargs.get('shot_record')

    @staticmethod
    def _apply_hints(lib, hints):
        """Hand data-path hints to a library that takes them (the prebuilt
        CUDA backend); a custom kernel built from ``cfile`` has no such entry
        point and is simply called as the reference would.  Returns the names
        that were set."""
        if not hints:
            return []
        try:
            setter = lib.simwave_cuda_set_hint
        except AttributeError:
            return []
        setter.restype = ctypes.c_int
        setter.argtypes = [ctypes.c_int, ctypes.c_longlong]
        done = []
        for name, value in hints.items():
            if setter(_HINT_CODE[name], int(value)) == 0:
                done.append(name)
        return done

    @staticmethod
    def _last_error(lib):
        """Message of the last failure, if the library keeps one."""
        try:
            getter = lib.simwave_cuda_last_error
        except AttributeError:
            return "unknown error (library has no simwave_cuda_last_error)"
        getter.restype = ctypes.c_char_p
        getter.argtypes = []
        message = getter()
        return message.decode(errors='replace') if message else "unknown error"

    @property
    def _keys_in_order(self):
        """All possible arg keys in the order the C function expects."""
        return list(ARGUMENT_ORDER)

    def _argtypes(self, **kwargs):
        """ctypes argtype for each keyword argument (by value type)."""
        types = {}
        for key, value in kwargs.items():
            if isinstance(value, np.ndarray):
                name = 'np({})'.format(str(value.dtype))
            else:
                name = type(value).__name__
            types[key] = self._convert_type_to_ctypes(name)
        return types

    def _convert_boundary_condition(self, boundary_condition):
        """(none: 0, null_dirichlet: 1, null_neumann: 2) as a uint64 array."""
        return np.uint([_BC_CODE[name] for name in boundary_condition])

    def _convert_type_to_ctypes(self, type):
        """Python / NumPy type name -> ctypes argtype."""
        return _CTYPE[type]
