"""
ctypes binding of ``libcpfftlog.so`` (C ABI declared in ``include/cpfftlog.h``).

The library is built in-tree by :func:`build` (``nvcc -gencode arch=compute_100a,code=sm_100a``) and loaded lazily.
There is deliberately NO CPU fallback: if the shared library or a CUDA device is missing, every compute entry point
raises (``ImportError`` / ``RuntimeError``) instead of silently computing something else.
"""

import os
import ctypes
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, 'csrc')
LIB_PATH = os.path.join(_HERE, 'libcpfftlog.so')
HEADER_PATH = os.path.join(os.path.dirname(_HERE), 'include', 'cpfftlog.h')

CPF_OK, CPF_EINVAL, CPF_ECUDA, CPF_ENOMEM, CPF_EUNSUPPORTED = range(5)
EXTRAP_CONST, EXTRAP_EDGE, EXTRAP_LOG = range(3)
MAX_N = 8192

_c_double_p = ctypes.POINTER(ctypes.c_double)
_c_int32_p = ctypes.POINTER(ctypes.c_int32)
_vp = ctypes.c_void_p
_i, _i64, _d = ctypes.c_int, ctypes.c_int64, ctypes.c_double

# name -> (restype, argtypes); must list every symbol of include/cpfftlog.h (tests/test_abi.py checks it)
SIGNATURES = {
    'cpf_version': (_i, []),
    'cpf_last_error': (ctypes.c_char_p, []),
    'cpf_device_count': (_i, [ctypes.POINTER(_i)]),
    'cpf_trim': (_i, [_i]),
    'cpf_counter': (_i64, [_i]),
    'cpf_plan_create': (_i, [ctypes.POINTER(_vp), _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _i]),
    'cpf_plan_destroy': (_i, [_vp]),
    'cpf_plan_kernel_family': (_i, [_vp, _i, _d, _i, _d, _i]),
    'cpf_fftlog': (_i, [_vp, _vp, _i64, _i, _i, _d, _i, _d, _i, _vp, _i, _i, _vp]),
    'cpf_rfft': (_i, [_i, _vp, _i64, _vp, _i, _i, _i, _vp]),
    'cpf_irfft_conj': (_i, [_i, _vp, _i64, _vp, _i, _i, _i, _vp]),
    'cpf_spline_create': (_i, [ctypes.POINTER(_vp), _vp, _vp, _i, _i64, _i, _i, _i, _i, _i, _i, _vp]),
    'cpf_spline_eval': (_i, [_vp, _vp, _i, _i, _vp, _i, _vp]),
    'cpf_spline_eval_t': (_i, [_vp, _vp, _i, _i, _vp, _i, _vp]),
    'cpf_spline_destroy': (_i, [_vp]),
    'cpf_spline_create_padlog': (_i, [ctypes.POINTER(_vp), _vp, _vp, _i, _i64, _i, _vp, _i, _i, _vp]),
    'cpf_column_nan_flags': (_i, [_vp, _i, _i64, _i, _vp, _i, _vp]),
    'cpf_spline_eval_rows': (_i, [_vp, _vp, _i, _i64, _vp, _i, _i, _i, _i, _vp, _i, _i, _vp]),
    'cpf_dst': (_i, [_i, _vp, _i, _i64, _vp, _i, _i, _vp]),
    'cpf_wallish2018': (_i, [_vp, _vp, _i, _vp, _vp, _i, _i64, _vp, _vp, _i, _i, _vp]),
    'cpf_wallish2018_rows': (_i, [_vp, _vp, _i, _vp, _vp, _i, _i64, _vp, _vp, _i, _i, _vp]),
    'cpf_eh_pk': (_i, [_vp, _vp, _i64, _i, _vp, _i, _d, _d, _d, _i, _vp, _vp, _i, _i, _vp]),
    'cpf_measure_fp64_peak': (_i, [_i, ctypes.POINTER(_d)]),
}

NVCC_FLAGS = ['-O3', '-std=c++17', '--threads', '4', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo',
              '-Xcompiler', '-fPIC', '-shared']

_lock = threading.Lock()
_lib = None


def sources():
    return sorted(os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith('.cu'))


def needs_build():
    if not os.path.isfile(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(_CSRC, f) for f in os.listdir(_CSRC)] + [HEADER_PATH]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile ``csrc/*.cu`` into ``libcpfftlog.so`` for sm_100a (cross-compiles without a GPU)."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = os.environ.get('NVCC', 'nvcc')
    cmd = [nvcc] + NVCC_FLAGS + ['-o', LIB_PATH] + sources()
    if verbose:
        print(' '.join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed:\n{}\n{}'.format(res.stdout, res.stderr))
    return LIB_PATH


def load():
    """Return the loaded library (ctypes.CDLL with prototypes set); raise ImportError if it cannot be loaded."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.isfile(LIB_PATH):
            raise ImportError('{} not found: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                              '(cosmoprimo_b200 has no CPU fallback)'.format(LIB_PATH))
        lib = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(lib, name)   # AttributeError if the ABI lost a symbol
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib


class CudaError(RuntimeError):
    pass


def check(rc):
    """Map a C status code to the exception type the reference raises in the same situation."""
    if rc == CPF_OK:
        return
    msg = load().cpf_last_error().decode('utf-8', 'replace')
    if rc == CPF_EINVAL:
        raise ValueError(msg)
    if rc == CPF_EUNSUPPORTED:
        raise NotImplementedError(msg)
    if rc == CPF_ENOMEM:
        raise MemoryError(msg)
    raise CudaError(msg)


def device_count():
    lib = load()
    n = ctypes.c_int(0)
    rc = lib.cpf_device_count(ctypes.byref(n))
    if rc != CPF_OK:
        return 0
    return n.value


def require_device():
    if device_count() < 1:
        raise CudaError('no CUDA device visible: cosmoprimo_b200 computes on the GPU only ({})'.format(
            load().cpf_last_error().decode('utf-8', 'replace')))
