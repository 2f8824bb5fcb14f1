"""Kernel-point dispositions for KPConv.

Mirror of `load_kernels` (/root/reference/Predator_APR/kernels/kernel_points.py:388-470) for the case every shipped
config uses: the stored K=15, 'center', 3-D disposition (kernels/dispositions/k_015_center_3D.ply; its 15x3 float64
coordinates are kept in apr_b200/data/k_015_center_3D.npy). Like the reference it draws a random z-rotation
(:436-445) and N(0, 0.01) noise (:462) from numpy's GLOBAL RNG, scales by the radius (:465), rotates (:468) and
returns float32. Other dispositions would need the reference's offline optimiser (:66-385), which is out of scope.
"""
import os

import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def load_kernels(radius, num_kpoints, dimension, fixed, lloyd=False):
    path = os.path.join(_DATA, 'k_{:03d}_{:s}_{:d}D.npy'.format(num_kpoints, fixed, dimension))
    if not os.path.exists(path):
        raise NotImplementedError(f"kernel disposition {os.path.basename(path)} is not shipped (only K=15 'center' 3D)")
    kernel_points = np.load(path)
    # same draws, in the same order, as the reference (np.random.rand for theta, np.random.normal for the noise)
    theta = np.random.rand() * 2 * np.pi
    c, s = np.cos(theta), np.sin(theta)
    R = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], dtype=np.float32)
    kernel_points = kernel_points + np.random.normal(scale=0.01, size=kernel_points.shape)
    kernel_points = radius * kernel_points
    kernel_points = np.matmul(kernel_points, R)
    return kernel_points.astype(np.float32)
