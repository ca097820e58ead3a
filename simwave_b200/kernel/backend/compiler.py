"""
Backend selection: which shared library provides ``forward``.

Host-side mirror of simwave/kernel/backend/compiler.py.  ``Compiler`` keeps
the constructor signature, the attribute validation and the cflags handling
pinned by the reference's tests/test_compiler.py.  The difference is what
``compile()`` does for ``language='cuda'``: the reference shells out to nvcc
on its cuda/wave.cu (compiler.py:150-152, 209-227); here it returns the path
of a *prebuilt* sm_100a library from simwave_b200/lib/ and never runs a
compiler.  This package ships no CPU kernels: with a user-supplied ``cfile``
(the reference's custom-kernel hook, compiler.py:157-159) the other languages
compile and cache that file the way the reference does; without one they run
on the prebuilt CUDA backend too, with a warning (so that the reference's
examples and its default ``Compiler()`` work unmodified), or are refused with
SIMWAVE_B200_STRICT_LANGUAGE=1.  Nothing ever runs the time loop on the CPU.
"""
import os
import subprocess
import warnings
from hashlib import sha1

LANGUAGES = ('c', 'cpu_openmp', 'gpu_openmp', 'gpu_openacc', 'cuda')
OPERATORS = ('forward', 'adjoint')

_DEFAULT_CFLAGS = '-O3 -fPIC -Wall -std=c99 -shared'
_OPENMP_FLAG = {'gcc': '-fopenmp', 'icc': '-openmp',
                'pgcc': '-mp', 'clang': '-fopenmp'}
_LANGUAGE_MACRO = {'cpu_openmp': '-DCPU_OPENMP', 'gpu_openmp': '-DGPU_OPENMP',
                   'gpu_openacc': '-DGPU_OPENACC'}
_PRECISION_TAG = {'-DFLOAT': 'f32', '-DDOUBLE': 'f64'}

_WARNED_LANGUAGES = set()

LIB_DIR = os.path.join(
    os.path.dirname(os.path.dirname(os.path.dirname(
        os.path.realpath(__file__)))),
    'lib'
)
# development aid: experimental builds of the same libraries (make LIB=...)
LIB_DIR = os.environ.get('SIMWAVE_B200_LIB_DIR', LIB_DIR)


def prebuilt_library(dimension, density, float_precision):
    """Path of the prebuilt sm_100a library for one kernel variant.

    ``density`` is 'constant_density' or 'variable_density';
    ``float_precision`` is '-DFLOAT' or '-DDOUBLE' (the reference's own
    selectors, middleware.py:31-48)."""
    return os.path.join(LIB_DIR, 'libsimwave_cuda_{}d_{}_{}.so'.format(
        dimension, density.replace('_density', ''),
        _PRECISION_TAG[float_precision.strip()]
    ))


class Compiler:
    """
    Backend selector / runtime compiler.

    Parameters
    ----------
    cc : str, optional
        C compiler. Default is gcc. Ignored for language='cuda'.
    language: str, optional
        c, cpu_openmp, gpu_openmp, gpu_openacc or cuda. Default is c.
        Only cuda is shipped prebuilt; the others need ``cfile``.
    cflags : str, optional
        C compiler flags.
        Default is '-O3 -fPIC -Wall -std=c99 -shared'.
    cfile : str, optional
        Path to a file with a custom C kernel implementation.
    """
    def __init__(self, cc='gcc', language='c', cflags=None, cfile=None):
        self.cc = cc
        self.language = language
        self.cflags = cflags
        self.cfile = cfile

    @property
    def cc(self):
        return self._cc

    @cc.setter
    def cc(self, value):
        if not isinstance(value, str):
            raise TypeError("Compiler.cc attribute must be str.")
        self._cc = value

    @property
    def language(self):
        return self._language

    @language.setter
    def language(self, value):
        if not isinstance(value, str):
            raise TypeError("Compiler.language attribute must be str.")
        if value not in LANGUAGES:
            raise ValueError(
                "Compiler.language {} not implemented.".format(value)
            )
        self._language = value

    @property
    def cflags(self):
        return self._cflags

    @cflags.setter
    def cflags(self, value):
        if value is None:
            value = _DEFAULT_CFLAGS
        if not isinstance(value, str):
            raise TypeError("Compiler.cflags attribute must be str.")

        if '-shared' not in value:
            value += ' -shared'

        if self.language in ('cpu_openmp', 'gpu_openmp'):
            omp_flag = self.get_openmp_flag()
            if omp_flag is None:
                print("WARNING: make sure OpenMP flag is provided in cflags.")
            elif omp_flag not in value:
                value += ' {}'.format(omp_flag)

        self._cflags = value

    @property
    def cfile(self):
        return self._cfile

    @cfile.setter
    def cfile(self, value):
        self._cfile = value

    def get_openmp_flag(self):
        """OpenMP flag of the configured compiler, None if unknown."""
        return _OPENMP_FLAG.get(self.cc)

    def compile(self, dimension, density, float_precision, operator):
        """
        Return the path of the shared object that exports ``forward``.

        Parameters
        ----------
        dimension : int
            Grid dimension. 2D (2) or 3D (3).
        density : str
            'constant_density' or 'variable_density'.
        float_precision : str
            '-DFLOAT' or '-DDOUBLE'.
        operator : str
            'forward', or 'adjoint' (constant density; exported by the same
            prebuilt library, include/simwave_cuda.h section 1b).
        """
        if operator not in OPERATORS:
            raise ValueError("Operator {} not available.".format(operator))

        if self.cfile is not None:
            return self._compile_custom(dimension, density, float_precision,
                                        operator)

        if self.language != 'cuda':
            # The reference's examples, benchmarks and its default
            # Compiler() ask for 'c' / 'cpu_openmp' / 'gpu_*' kernels, which
            # this package does not ship.  They still run: on the prebuilt
            # CUDA backend, never on the CPU, and say so once per language.
            if os.environ.get('SIMWAVE_B200_STRICT_LANGUAGE') == '1':
                raise NotImplementedError(
                    "simwave_b200 ships only the prebuilt CUDA (sm_100a) "
                    "backend: use Compiler(language='cuda'), or pass cfile= to "
                    "build a custom kernel with language={!r}."
                    .format(self.language)
                )
            if self.language not in _WARNED_LANGUAGES:
                _WARNED_LANGUAGES.add(self.language)
                warnings.warn(
                    "simwave_b200 has no {!r} kernels: running on the prebuilt "
                    "CUDA (sm_100a) backend instead; cc and cflags are "
                    "ignored.".format(self.language), RuntimeWarning,
                    stacklevel=2
                )

        path = prebuilt_library(dimension, density, float_precision)
        if not os.path.exists(path):
            raise FileNotFoundError(
                "Prebuilt CUDA library {} is missing; build it with "
                "`python -c \"import __graft_entry__ as g; g.build()\"` "
                "from the repository root. There is no CPU fallback."
                .format(path)
            )
        return path

    def _compile_custom(self, dimension, density, float_precision, operator):
        """Compile ``cfile`` into ./tmp/<sha1>.so unless cached; the cache
        key covers compiler, flags, variant and source text, like the
        reference's (compiler.py:180-202)."""
        with open(self.cfile, 'r', encoding='utf-8') as f:
            source = f.read()

        macro = _LANGUAGE_MACRO.get(self.language, '')
        key = '\n'.join([
            'wave', self.language, self.cc, self.cflags,
            '{}d'.format(dimension), operator, density, float_precision,
            macro, source
        ])
        object_dir = os.path.join(os.getcwd(), 'tmp')
        object_path = os.path.join(
            object_dir, sha1(key.encode()).hexdigest() + '.so'
        )

        if os.path.exists(object_path):
            print("Shared object already compiled in:", object_path)
            return object_path

        command = [self.cc, self.cfile]
        command += self.cflags.split()
        command += [flag for flag in (float_precision.strip(), macro) if flag]
        command += ['-o', object_path]
        print("Compilation command:", ' '.join(command))

        os.makedirs(object_dir, exist_ok=True)
        if subprocess.run(command).returncode != 0:
            raise Exception("Compilation failed")
        return object_path
