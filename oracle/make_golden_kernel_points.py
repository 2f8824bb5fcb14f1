"""TEST INFRASTRUCTURE — freezes tests/golden/kernel_points.npz from the REAL reference's load_kernels
(/root/reference/Predator_APR/kernels/kernel_points.py:388-470) under fixed numpy global seeds, so that
apr_b200.kernel_points.load_kernels can be checked bit for bit where /root/reference does not exist.
Run from the repo root:  PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_kernel_points.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/Predator_APR"
sys.dont_write_bytecode = True


def main():
    os.chdir(REF)                                            # load_kernels opens the relative path kernels/dispositions
    sys.path.insert(0, REF)
    from kernels.kernel_points import load_kernels
    out = {}
    for seed in (0, 1, 12345):
        np.random.seed(seed)
        for i, radius in enumerate((1.275, 2.55, 5.1, 10.2)):   # consecutive draws from ONE global stream, like a model build
            out[f"s{seed}_{i}"] = load_kernels(radius, 15, dimension=3, fixed='center')
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "kernel_points.npz"), **out)
    print("wrote", len(out), "kernel-point sets")


if __name__ == "__main__":
    main()
