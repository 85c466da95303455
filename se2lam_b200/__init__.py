"""se2lam_b200 — B200-native (sm_100a CUDA) implementation of se2lam's two data-parallel hot paths:
the ORB front-end and the SE(2)-XYZ local bundle adjustment, behind the C ABI of include/se2gpu.h.

Host-side mirrors of the reference interfaces: `orb.ORBextractor`, `matcher.ORBmatcher`, `ba.SlamOptimizer`.
"""
__version__ = "0.1.0"
