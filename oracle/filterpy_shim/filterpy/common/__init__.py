"""filterpy.common restatement (test infrastructure only, see ../__init__.py)."""
import numpy as np


def Q_discrete_white_noise(dim, dt=1.0, var=1.0, block_size=1, order_by_dim=True):
    """Discrete white-noise process covariance, filterpy 1.4.5 semantics.

    The reference only calls it with dim=3, block_size=1 (constants.py:212-214,
    241-242); other shapes are restated for completeness of the published
    formula.
    """
    if dim not in (2, 3, 4):
        raise ValueError("dim must be between 2 and 4")
    if dim == 2:
        Q = [[0.25 * dt**4, 0.5 * dt**3],
             [0.5 * dt**3, dt**2]]
    elif dim == 3:
        Q = [[0.25 * dt**4, 0.5 * dt**3, 0.5 * dt**2],
             [0.5 * dt**3, dt**2, dt],
             [0.5 * dt**2, dt, 1]]
    else:
        Q = [[dt**6 / 36, dt**5 / 12, dt**4 / 6, dt**3 / 6],
             [dt**5 / 12, dt**4 / 4, dt**3 / 2, dt**2 / 2],
             [dt**4 / 6, dt**3 / 2, dt**2, dt],
             [dt**3 / 6, dt**2 / 2, dt, 1.0]]
    Q = np.array(Q, dtype=float)
    if block_size > 1:
        # order_by_dim layouts are never used by the reference
        from scipy.linalg import block_diag
        if order_by_dim:
            return block_diag(*[Q] * block_size) * var
        raise NotImplementedError("order_by_dim=False is outside the restated surface")
    return Q * var
