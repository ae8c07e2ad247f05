"""CPU thread-emulation of the register FFT passes the CUDA kernels are built from (tests/emul/emul_fft.cpp)."""
import os
import subprocess
import tempfile


def test_three_pass_fft_emulation():
    here = os.path.dirname(os.path.abspath(__file__))
    with tempfile.TemporaryDirectory() as tmp:
        exe = os.path.join(tmp, 'emul_fft')
        subprocess.run(['g++', '-O2', '-std=c++17', '-o', exe, os.path.join(here, 'emul', 'emul_fft.cpp')], check=True)
        res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout
    assert 'OK' in res.stdout


def test_stream_kernel_flow_emulation():
    """Data flow of the stream kernel (cpf_stream_core.h): slot ownership, warp-local exchanges, twiddle tables."""
    here = os.path.dirname(os.path.abspath(__file__))
    with tempfile.TemporaryDirectory() as tmp:
        exe = os.path.join(tmp, 'emul_stream')
        subprocess.run(['g++', '-O2', '-std=c++17', '-o', exe, os.path.join(here, 'emul', 'emul_stream.cpp')], check=True)
        res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout
    assert 'OK' in res.stdout


def test_spline_window_weights_emulation():
    """Windowed spline weights behind cpf_spline_eval_rows (cpf_spline_core.h) against a long-double full solve."""
    here = os.path.dirname(os.path.abspath(__file__))
    with tempfile.TemporaryDirectory() as tmp:
        exe = os.path.join(tmp, 'emul_spline_window')
        subprocess.run(['g++', '-O2', '-std=c++17', '-o', exe, os.path.join(here, 'emul', 'emul_spline_window.cpp')], check=True)
        res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout
    assert 'OK' in res.stdout


def test_fastmath_against_libm():
    """cpf_fastmath.h (fast_log / fast_exp of the Wallish2018 kernel, fast_log10 / fast_exp10 of the spline kernels) against long-double libm:
    every function < 2 ulp over 2 M arguments, special values as libm."""
    here = os.path.dirname(os.path.abspath(__file__))
    with tempfile.TemporaryDirectory() as tmp:
        exe = os.path.join(tmp, 'emul_fastmath')
        subprocess.run(['g++', '-O2', '-std=c++17', '-o', exe, os.path.join(here, 'emul', 'emul_fastmath.cpp')], check=True)
        res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout
    assert 'OK' in res.stdout
