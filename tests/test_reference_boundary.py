"""CPU side of the reference-facing boundary: the unmodified reference (baseline/_ref) accepts our plug-ins by construction.  No compute calls
(those are in test_reference_boundary_gpu.py)."""
import numpy as np

from conftest import reference
from cosmoprimo_b200 import fftlog as F
from cosmoprimo_b200 import bao_filter as B


def test_engine_instance_passes_through_reference_factory():
    """ref fftlog.py:641-663: a non-string engine is returned unchanged, so FFTlog(..., engine=CudaFFTEngine(...)) needs no patch."""
    ref = reference('fftlog')
    eng = F.CudaFFTEngine(2048, nparallel=3)
    assert ref.get_fft_engine(eng, size=2048, nparallel=3) is eng
    k = np.geomspace(1e-5, 1e2, 1024)
    obj = ref.PowerToCorrelation(k, ell=[0, 2, 4], engine=eng)
    assert obj._engine is eng and obj.padded_size == eng.size and obj.nparallel == eng.nparallel
    # our host plan tables are the reference's (same _setup arithmetic)
    ours = F.PowerToCorrelation(k, ell=[0, 2, 4], engine=eng)
    for name in ['padded_prefactor', 'padded_postfactor', 'padded_u', 'y']:
        np.testing.assert_allclose(getattr(ours, name), getattr(obj, name), rtol=1e-12, atol=0)


def test_filter_registers_in_reference_registry():
    """ref bao_filter.py:22-31, 912-921: the factory looks the engine name up in the metaclass registry."""
    reference()
    refb = B.register_in_reference()
    reg = refb.RegisteredPowerSpectrumBAOFilter._registry
    assert reg['wallish2018_cuda'] is B.Wallish2018PowerSpectrumBAOFilter
    assert reg['wallish2018'].__module__ == 'cosmoprimo.bao_filter'      # the reference's own entry is untouched
    # same constructor contract as the reference's base class (ref:39-64)
    import inspect
    ref_args = list(inspect.signature(refb.BasePowerSpectrumBAOFilter.__init__).parameters)
    our_args = list(inspect.signature(B.BasePowerSpectrumBAOFilter.__init__).parameters)
    assert ref_args[:4] == our_args[:4] == ['self', 'pk_interpolator', 'cosmo', 'cosmo_fid']
