"""fen_b200 -- B200-native drop-in for FEN's fractional-step Navier-Stokes hot path.

The product is ``libfen_gpu.so`` (hand-written sm_100a CUDA behind the C ABI of
``include/fen_gpu.h``); this package is the thin host-side mirror of the reference's solver API."""
from .api import (FenError, MultiphaseSolver, PoissonSolver, Solver, VoF, center_to_face, curl, divergence,  # noqa: F401
                  face_to_center, gradient, grid, laplacian, scalar, vector)
