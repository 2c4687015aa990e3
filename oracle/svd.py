"""Safe batched SVD with the reference's hand-written VJP (test infrastructure).

Follows core/engine/svd_safe_batch.py:19-102.  Forward = LAPACK SVD
(torch.linalg.svd on CPU, the same gesdd family jnp.linalg.svd uses on CPU).
Backward = the formula at svd_safe_batch.py:65-102, evaluated as written
(including the terms that vanish analytically for real square input).
"""
import torch

DEFAULT_EPS = 1e-12  # svd_safe_batch.py:10


def _safe_inverse(x, eps):
    return x / (x ** 2 + eps)  # svd_safe_batch.py:54-55


class SafeSVD(torch.autograd.Function):
    @staticmethod
    def forward(ctx, A):
        U, S, Vh = torch.linalg.svd(A, full_matrices=False)
        ctx.save_for_backward(U, S, Vh)
        return U, S, Vh

    @staticmethod
    def backward(ctx, dU, dS, dVh):
        U, S, Vh = ctx.saved_tensors
        eps = DEFAULT_EPS
        if dU is None:
            dU = torch.zeros_like(U)
        if dS is None:
            dS = torch.zeros_like(S)
        if dVh is None:
            dVh = torch.zeros_like(Vh)
        Ut = U.transpose(1, 2)
        Vt = Vh                                 # svd_safe_batch.py:75 (conj only)
        Vt_dV = Vt @ dVh.transpose(1, 2)        # :76
        S_squared = S ** 2
        S_inv = _safe_inverse(S, eps)
        k = S.shape[-1]
        I = torch.eye(k, dtype=A_dtype(U)).expand(S.shape[0], k, k)
        F = _safe_inverse(S_squared[:, None, :] - S_squared[..., None], eps)  # :82
        F = F - I * F
        J = F * (Ut @ dU)
        K = F * Vt_dV
        L = I * Vt_dV
        Pc_U_perp = I - U @ Ut
        Pc_V_perp = I - Vh.transpose(1, 2) @ Vt
        S_, dS_, S_inv_ = S[:, None, :], dS[:, None, :], S_inv[:, None, :]
        dA = (U * dS_) @ Vt \
            + U @ ((J + J.transpose(1, 2)) * S_) @ Vt \
            + (U * S_) @ (K + K.transpose(1, 2)) @ Vt \
            + 0.5 * ((U * S_inv_) @ (L - L.transpose(1, 2)) @ Vt) \
            + Pc_U_perp @ (dU * S_inv_) @ Vt \
            + (U * S_inv_) @ dVh @ Pc_V_perp
        return dA


def A_dtype(t):
    return t.dtype


def svd(A):
    """A: (n,3,3) -> U (n,3,3), S (n,3) descending >=0, Vh (n,3,3)."""
    return SafeSVD.apply(A)
