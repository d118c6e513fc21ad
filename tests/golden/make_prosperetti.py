"""Copies the reference's known-answer data for the capillary-wave case into a fixture.

Source: /root/reference/test/small_test/multiphase/capillary_wave/prosperetti.csv -- Prosperetti's analytic solution
of the viscous capillary wave (non-dimensional time omega_0 t, maximum interface amplitude), the curve the reference's
postpro.py:92-101 measures its result against.  It is data, not code; /root/reference does not exist on the GPU box,
so the 738 points travel as tests/golden/prosperetti_capillary.npz.

Usage (in the build container, where /root/reference is mounted):  python tests/golden/make_prosperetti.py
"""
import os

import numpy as np

SRC = "/root/reference/test/small_test/multiphase/capillary_wave/prosperetti.csv"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "prosperetti_capillary.npz")

if __name__ == "__main__":
    curve = np.genfromtxt(SRC, delimiter=",")
    assert curve.shape == (738, 2)
    np.savez_compressed(OUT, curve=curve)
    print(OUT, curve.shape)
