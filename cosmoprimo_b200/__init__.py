"""
cosmoprimo_b200: B200-native (sm_100a) engine for cosmoprimo's FFTLog / Wallish2018 / cubic-spline hot path.

Only the hot path is here: ``fftlog`` mirrors ``cosmoprimo.fftlog`` with ``engine='cuda'``.
"""

from .fftlog import (FFTlog, HankelTransform, PowerToCorrelation, CorrelationToPower, TophatVariance, GaussianVariance,
                     CudaFFTEngine, get_fft_engine, pad)

__version__ = '0.1.0'
