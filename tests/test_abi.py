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


def test_argument_validation_of_the_other_entry_points():
    """Bad arguments are rejected with CPF_EINVAL / CPF_EUNSUPPORTED and a message before any CUDA work."""
    lib = _lib.load()
    d = np.zeros(16)
    p = d.ctypes.data
    handle = ctypes.c_void_p()
    E, U = _lib.CPF_EINVAL, _lib.CPF_EUNSUPPORTED
    assert lib.cpf_spline_create(ctypes.byref(handle), p, p, 1, 1, 0, 0, 0, 0, 0, 0, None) == E           # one knot
    assert lib.cpf_spline_create(ctypes.byref(handle), p, p, 8, 1, 7, 0, 0, 0, 0, 0, None) == E           # unknown end condition
    assert lib.cpf_spline_create(ctypes.byref(handle), p, p, 3, 1, 2, 0, 0, 0, 0, 0, None) == U           # not-a-knot needs 4 knots
    assert b'4 knots' in lib.cpf_last_error()
    assert lib.cpf_spline_eval(None, p, 1, 0, p, 0, None) == E and lib.cpf_spline_eval_t(None, p, 1, 0, p, 0, None) == E
    assert lib.cpf_spline_destroy(None) == _lib.CPF_OK
    assert lib.cpf_spline_eval_rows(p, p, 1, 1, p, 1, 0, 0, 0, p, 0, 0, None) == E                        # one knot
    assert lib.cpf_spline_eval_rows(p, p, 8, 1, p, 1, 5, 0, 0, p, 0, 0, None) == E                        # unknown end condition
    assert lib.cpf_spline_eval_rows(p, p, 8, 1, p, 1, 0, -1, 0, p, 0, 0, None) == E                       # negative window
    assert lib.cpf_spline_eval_rows(p, p, 8, 0, p, 1, 0, 0, 0, p, 0, 0, None) == _lib.CPF_OK              # no rows: nothing to do
    assert lib.cpf_dst(4, p, 4096, 1, p, 0, 0, None) == E and lib.cpf_dst(2, p, 1024, 1, p, 0, 0, None) == U
    assert lib.cpf_wallish2018(p, p, 1024, p, p, 16, 1, p, None, 0, 0, None) == U                         # the grid of bao_filter.py:364 only
    assert lib.cpf_eh_pk(p, p, -1, 1, p, 4, 2.7255, 4e-5, 0.05, 0, p, None, 0, 0, None) == E
    assert lib.cpf_eh_pk(p, p, 1, 0, p, 4, 2.7255, 4e-5, 0.05, 0, p, None, 0, 0, None) == E               # nz < 1
    assert lib.cpf_eh_pk(p, None, 1, 3, p, 4, 2.7255, 4e-5, 0.05, 0, p, None, 0, 0, None) == E            # z = NULL needs nz = 1
    assert lib.cpf_eh_pk(p, p, 1, 1, p, 4, -1., 4e-5, 0.05, 0, p, None, 0, 0, None) == E                  # T_cmb <= 0
    assert lib.cpf_eh_pk(p, p, 0, 1, p, 4, 2.7255, 4e-5, 0.05, 0, p, None, 0, 0, None) == _lib.CPF_OK     # empty batch
    assert lib.cpf_rfft(12, p, 1, p, 0, 0, 0, None) == E


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
