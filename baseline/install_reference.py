"""
Install the UNMODIFIED reference (cosmodesi/cosmoprimo, pure Python) into ``baseline/_ref`` so that it travels to the GPU box
(``baseline/_ref`` is git-ignored, not gpurun-ignored) and can be driven there through its own public API:

* ``bench.py --impl reference`` times ``cosmoprimo.fftlog.PowerToCorrelation(k, ell=[0, 2, 4], engine='numpy')``;
* the ``-m gpu`` tests hand ``engine=CudaFFTEngine(...)`` / ``engine='wallish2018_cuda'`` to the reference's own classes.

    python baseline/install_reference.py [--force]

The command is the base contract's: ``pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse
--target baseline/_ref <copy of /root/reference>`` (``--no-deps``: numpy / scipy are already in the image, the wheelhouse has no
numpy wheel; a copy under /tmp because the build writes an egg-info into the source tree and /root/reference is read-only).
No reference source is committed: only this recipe.
"""

import os
import sys
import shutil
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
TARGET = os.path.join(HERE, '_ref')
SOURCE = os.environ.get('CPF_REFERENCE_SOURCE', '/root/reference')


def installed():
    return os.path.isfile(os.path.join(TARGET, 'cosmoprimo', 'fftlog.py'))


def install(force=False, verbose=True):
    """Returns True if baseline/_ref holds the reference afterwards."""
    if installed() and not force:
        return True
    if not os.path.isdir(os.path.join(SOURCE, 'cosmoprimo')):
        return False          # the GPU box: only the prebuilt copy is used
    tmp = tempfile.mkdtemp(prefix='cpf_ref_')
    try:
        src = os.path.join(tmp, 'reference')
        shutil.copytree(SOURCE, src, ignore=shutil.ignore_patterns('.git', '__pycache__'))
        if os.path.isdir(TARGET):
            shutil.rmtree(TARGET)
        cmd = [sys.executable, '-m', 'pip', 'install', '--no-index', '--no-build-isolation', '--no-deps', '--find-links', '/opt/wheelhouse',
               '--target', TARGET, src]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if verbose:
            print(' '.join(cmd))
            print(res.stdout[-400:] + res.stderr[-400:])
        return res.returncode == 0 and installed()
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def activate():
    """Put baseline/_ref on sys.path (or, in the build container only, /root/reference + the dist-info shim). Returns the path used or None."""
    if installed():
        path = TARGET
    elif os.path.isdir(os.path.join(SOURCE, 'cosmoprimo')):
        shim = os.path.join(os.path.dirname(HERE), 'tools', 'refshim')
        if shim not in sys.path:
            sys.path.insert(0, shim)
        path = SOURCE
    else:
        return None
    if path not in sys.path:
        sys.path.insert(0, path)
    return path


if __name__ == '__main__':
    ok = install(force='--force' in sys.argv)
    print('baseline/_ref {}'.format('ready' if ok else 'NOT installed'))
    sys.exit(0 if ok else 1)
