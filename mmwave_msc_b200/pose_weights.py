"""Weights of the MARS-derived keypoint regressor (reference train.py:33-106).

The trained ``MARS.h5`` is not shipped with the reference (constants.py:14), so
benchmarks and parity tests use seeded random-initialised weights of exactly that
architecture.  The flat fp32 blob the C ABI takes (``mmw_load_pose_weights``) is
``numpy.concatenate([w.ravel() for w in model.get_weights()])`` of the Keras
model, i.e. in this order:

  conv1 kernel (k.., cin, cout), conv1 bias, conv2 kernel, conv2 bias,
  BN1 gamma, beta, moving_mean, moving_variance,
  dense1 kernel (in, out), dense1 bias, BN2 gamma, beta, moving_mean, moving_variance,
  dense2 kernel (in, out), dense2 bias
"""
from __future__ import annotations

from typing import List

import numpy as np

VARIANT_2D = 0      # define_CNN     : (8,8,5)   -> 57   (FB_FRAMES_BATCH == 0)
VARIANT_3D = 1      # define_CNN_3D  : (3,8,8,5) -> 57   (default constants)
N_KEYPOINTS = 57


def weight_shapes(variant: int):
    if variant == VARIANT_2D:
        k1, k2, flat, hid = (3, 3, 5, 16), (3, 3, 16, 32), 8 * 8 * 32, 512
    elif variant == VARIANT_3D:
        k1, k2, flat, hid = (3, 3, 3, 5, 16), (3, 3, 3, 16, 32), 3 * 8 * 8 * 32, 3 * 512
    else:
        raise ValueError("variant must be VARIANT_2D or VARIANT_3D")
    return [k1, (16,), k2, (32,), (32,), (32,), (32,), (32,),
            (flat, hid), (hid,), (hid,), (hid,), (hid,), (hid,),
            (hid, N_KEYPOINTS), (N_KEYPOINTS,)]


def blob_size(variant: int) -> int:
    return int(sum(int(np.prod(s)) for s in weight_shapes(variant)))


def make_pose_weights(variant: int, seed: int = 7) -> List[np.ndarray]:
    """He-initialised seeded weights: conv/dense ~ N(0, sqrt(2/fan_in)),
    biases ~ N(0, 0.05), BN gamma ~ U(0.5,1.5), beta ~ N(0,0.1), mean ~ N(0,0.1), var ~ U(0.5,1.5)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    sh = weight_shapes(variant)
    out: List[np.ndarray] = []

    def he(shape):
        fan_in = int(np.prod(shape[:-1]))
        return rng.normal(0, np.sqrt(2.0 / fan_in), size=shape)

    def bn(n):
        return [rng.uniform(0.5, 1.5, size=n), rng.normal(0, 0.1, size=n),
                rng.normal(0, 0.1, size=n), rng.uniform(0.5, 1.5, size=n)]

    out += [he(sh[0]), rng.normal(0, 0.05, size=sh[1])]
    out += [he(sh[2]), rng.normal(0, 0.05, size=sh[3])]
    out += bn(32)
    out += [he(sh[8]), rng.normal(0, 0.05, size=sh[9])]
    out += bn(sh[9][0])
    out += [he(sh[14]), rng.normal(0, 0.05, size=sh[15])]
    return [np.ascontiguousarray(a, dtype=np.float32) for a in out]


def pack_blob(weights: List[np.ndarray]) -> np.ndarray:
    return np.ascontiguousarray(np.concatenate([np.asarray(w, dtype=np.float32).ravel() for w in weights]))


def unpack_blob(blob: np.ndarray, variant: int) -> List[np.ndarray]:
    out, o = [], 0
    for s in weight_shapes(variant):
        n = int(np.prod(s))
        out.append(np.asarray(blob[o:o + n], dtype=np.float32).reshape(s))
        o += n
    if o != blob.size:
        raise ValueError("blob has %d floats, variant needs %d" % (blob.size, o))
    return out
