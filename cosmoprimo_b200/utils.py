"""
Small dense least squares behind the polynomial BAO filters (``cosmoprimo/utils.py:144-272``, cited as ``ref:LINE``).

The model is linear, ``model = params . gradient``, so the best fit is a LINEAR map of the data: ``params = data . projector``.  The
projector depends on the gradient, the precision and the constraints only; it is built once on the host (``numpy.linalg.solve`` on the
bordered normal matrix, as the reference does at every call when ``compute_inverse=False``, ref:253) and applied to all spectra of a
batch with one matrix product -- numpy for host arrays, a cuBLAS fp64 GEMM through torch for CUDA tensors (a plain library GEMM of a
(batch, ndata) x (ndata, nparams) product; nothing here is worth a hand-written kernel).
"""

import numpy as np


class LeastSquareSolver(object):
    r"""
    Solve :math:`d\chi^{2}/d\mathbf{p} = 0` with :math:`\chi^{2} = (\delta - \mathbf{p} \cdot \mathbf{grad})^{T} \mathbf{F} (\delta - \mathbf{p} \cdot \mathbf{grad})`,
    optionally under linear constraints (Lagrange multipliers; ref:144-272, same constructor and methods).
    """

    def __init__(self, gradient, precision=1., constraint_gradient=None, compute_inverse=True):
        self.gradient = np.atleast_1d(np.asarray(gradient, dtype='f8'))
        self.isscalar = self.gradient.ndim == 1
        if self.isscalar:
            self.gradient = self.gradient[None, :]
        elif self.gradient.ndim != 2:
            raise ValueError('gradient must be at most 2D')
        self.precision = np.asarray(precision, dtype='f8')
        if self.precision.ndim <= 1:
            hv = self.gradient * self.precision                                      # ref:199-200
        else:
            hv = self.gradient.dot(self.precision)                                   # ref:202
        invfisher = hv.dot(self.gradient.T)                                          # ref:203
        nparams, ndata = self.gradient.shape
        if constraint_gradient is None:
            self.nconstraints = 0
        else:
            cg = np.atleast_2d(np.asarray(constraint_gradient, dtype='f8'))
            self.nconstraints = cg.shape[-1]
            if cg.ndim != 2 or cg.shape[0] != nparams:
                raise ValueError('constraint_gradient must be 2D, of first dimension the number of model parameters (gradient first dimension)')
            nc = self.nconstraints                                                   # bordered system, ref:213-216
            invfisher = np.block([[invfisher, -cg], [cg.T, np.zeros((nc, nc))]])
            hv = np.block([[hv, np.zeros((nparams, nc))], [np.zeros((nc, ndata)), np.eye(nc)]])
        self.inverse_fisher = invfisher
        self.gradient_precision = hv
        self.compute_inverse = bool(compute_inverse)
        if compute_inverse:
            fisher = np.linalg.inv(invfisher)                                        # ref:221
            if not np.allclose(fisher.dot(invfisher), np.eye(invfisher.shape[0]), rtol=1e-4, atol=1e-4):
                import warnings
                warnings.warn('Numerically inaccurate inverse matrix')
            self.projector = fisher.dot(hv).T                                        # ref:234
        else:
            # the reference solves the system at every call (ref:253); the data enter linearly, so one solve for the operator is the same map
            self.projector = np.linalg.solve(invfisher, hv).T
        self._device_projector = {}

    def _projector_like(self, delta):
        if isinstance(delta, np.ndarray):
            return self.projector
        key = (str(delta.device), delta.dtype)
        if key not in self._device_projector:
            import torch
            self._device_projector[key] = (torch.as_tensor(self.projector, device=delta.device, dtype=delta.dtype),
                                           torch.as_tensor(self.gradient, device=delta.device, dtype=delta.dtype))
        return self._device_projector[key][0]

    def compute(self, delta, constraint=None):
        """Best-fit parameters of ``delta`` (..., ndata) (numpy array or CUDA tensor), ref:245-255."""
        is_numpy = not hasattr(delta, 'device') or isinstance(delta, np.ndarray)
        if is_numpy:
            delta = np.atleast_1d(np.asarray(delta, dtype='f8'))
            self.delta = delta
            if constraint is not None:
                delta = np.concatenate([delta, np.atleast_1d(np.asarray(constraint, dtype='f8'))], axis=-1)
            if self.compute_inverse:
                params = delta.dot(self.projector)
            else:
                # host data (the one-off fiducial fits of the filters): the reference's own operation order (ref:253), so that quantities
                # derived with thresholds from the fit -- the peak positions of the peak-average filter -- come out identical
                params = np.linalg.solve(self.inverse_fisher, self.gradient_precision.dot(delta.T)).T
        else:
            import torch
            self.delta = delta
            if constraint is not None:
                delta = torch.cat([delta, constraint], dim=-1)
            params = delta @ self._projector_like(delta)
        self.params = params[..., :self.gradient.shape[0]]

    def __call__(self, delta, constraint=None):
        self.compute(delta, constraint=constraint)
        if self.isscalar: return self.params[..., 0]
        return self.params

    def model(self):
        """Model at the best fit (ref:263-265)."""
        if isinstance(self.params, np.ndarray):
            return self.params.dot(self.gradient)
        self._projector_like(self.params)
        return self.params @ self._device_projector[(str(self.params.device), self.params.dtype)][1]

    def chi2(self):
        """chi2 at the best fit (ref:267-272); host arrays only."""
        delta = np.asarray(self.delta) - np.asarray(self.model())
        if self.precision.ndim <= 1:
            return ((delta * self.precision) * delta).sum(axis=-1)
        return (delta.dot(self.precision) * delta).sum(axis=-1)
