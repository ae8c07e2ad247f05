"""The C-ABI shared library loads and exports every symbol include/cpfftlog.h declares; argument validation that
does not need a GPU.  CPU only."""
import re
import ctypes

import numpy as np

from cosmoprimo_b200 import _lib


def header_symbols():
    src = open(_lib.HEADER_PATH).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(cpf_[a-z0-9_]+)\s*\(', src)))


def test_every_declared_symbol_is_exported():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = header_symbols()
    assert len(names) >= 12
    for name in names:
        assert hasattr(lib, name), 'libcpfftlog.so does not export {}'.format(name)
    assert sorted(_lib.SIGNATURES) == names, 'ctypes prototypes out of sync with the header'


def test_version_and_errors():
    lib = _lib.load()
    assert lib.cpf_version() == 100
    handle = ctypes.c_void_p()
    dummy = np.zeros(16)
    # bad sizes are rejected before any CUDA call
    rc = lib.cpf_plan_create(ctypes.byref(handle), 4, 12, 1, 4, 4, dummy.ctypes.data, dummy.ctypes.data, dummy.ctypes.data, None, 0)
    assert rc == _lib.CPF_EINVAL and b'power of two' in lib.cpf_last_error()
    rc = lib.cpf_plan_create(ctypes.byref(handle), 4, 1 << 20, 1, 4, 4, dummy.ctypes.data, dummy.ctypes.data, dummy.ctypes.data, None, 0)
    assert rc == _lib.CPF_EUNSUPPORTED
    rc = lib.cpf_plan_create(ctypes.byref(handle), 9, 8, 1, 0, 0, dummy.ctypes.data, dummy.ctypes.data, dummy.ctypes.data, None, 0)
    assert rc == _lib.CPF_EINVAL
    assert lib.cpf_plan_destroy(None) == _lib.CPF_OK
    assert lib.cpf_fftlog(None, None, 1, 1, 0, 0., 0, 0., 0, None, 0, 0, None) == _lib.CPF_EINVAL
    for rc, exc in [(_lib.CPF_EINVAL, ValueError), (_lib.CPF_EUNSUPPORTED, NotImplementedError), (_lib.CPF_ECUDA, RuntimeError)]:
        try:
            _lib.check(rc)
        except exc:
            pass
        else:
            raise AssertionError('status {} not mapped to {}'.format(rc, exc))


def test_no_cpu_fallback_without_gpu():
    """Without a device the product path must fail loudly, not compute on the CPU."""
    if _lib.device_count() > 0:
        return
    from cosmoprimo_b200.fftlog import PowerToCorrelation
    k = np.geomspace(1e-3, 1e1, 64)
    try:
        PowerToCorrelation(k)(np.ones(64))
    except RuntimeError:
        pass
    else:
        raise AssertionError('call succeeded without a CUDA device')
