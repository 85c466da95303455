"""Host-side mirror of se2lam::ORBmatcher (reference include/se2lam/ORBmatcher.h:40-81) over the C ABI.

The reference's Frame / KeyFrame / MapPoint objects are flattened to arrays (see include/se2gpu.h);
`FrameView` carries what the matcher reads from a Frame: keyPointsUn, descriptors and the grid bounds.
"""
from __future__ import annotations

import ctypes as C
import dataclasses

import numpy as np

from . import _capi
from ._capi import BowKF, GridParams, KP_DTYPE, check, lib, ptr

FRAME_GRID_ROWS, FRAME_GRID_COLS = 48, 64   # Frame.h:26-27


@dataclasses.dataclass
class FrameView:
    keyPointsUn: np.ndarray      # KP_DTYPE [N]
    descriptors: np.ndarray      # [N,32] uint8
    minXUn: float = 0.0
    maxXUn: float = 640.0
    minYUn: float = 0.0
    maxYUn: float = 480.0

    @property
    def N(self):
        return len(self.keyPointsUn)

    def grid(self) -> GridParams:   # Frame.cpp:37-40
        f = np.float32
        return GridParams(f(self.minXUn), f(self.minYUn), f(f(FRAME_GRID_COLS) / f(f(self.maxXUn) - f(self.minXUn))),
                          f(f(FRAME_GRID_ROWS) / f(f(self.maxYUn) - f(self.minYUn))))


class ORBmatcher:
    TH_HIGH, TH_LOW, HISTO_LENGTH = 100, 75, 30   # ORBmatcher.cpp:45-47

    def __init__(self, nnratio=0.6, checkOri=True, device=0):
        self.mfNNratio, self.mbCheckOrientation, self.device = float(nnratio), bool(checkOri), device

    @staticmethod
    def DescriptorDistance(a, b, device=0):
        a = np.ascontiguousarray(a, np.uint8).reshape(-1, 32); b = np.ascontiguousarray(b, np.uint8).reshape(-1, 32)
        out = np.zeros(len(a), np.int32)
        check(lib().se2gpu_hamming_distance(ptr(a), ptr(b), len(a), ptr(out), device), "se2gpu_hamming_distance")
        return int(out[0]) if len(out) == 1 else out

    def MatchByWindow(self, frame1: FrameView, frame2: FrameView, vbPrevMatched, winSize, levelOffset=1, minLevel=0, maxLevel=8):
        """Returns (nmatches, vnMatches12); vbPrevMatched [N1,2] float32 is updated in place."""
        kp1 = np.ascontiguousarray(frame1.keyPointsUn, KP_DTYPE); kp2 = np.ascontiguousarray(frame2.keyPointsUn, KP_DTYPE)
        d1 = np.ascontiguousarray(frame1.descriptors, np.uint8); d2 = np.ascontiguousarray(frame2.descriptors, np.uint8)
        assert vbPrevMatched.dtype == np.float32 and vbPrevMatched.flags.c_contiguous
        m = np.full(len(kp1), -1, np.int32)
        n = check(lib().se2gpu_match_by_window(ptr(kp1), ptr(d1), len(kp1), ptr(kp2), ptr(d2), len(kp2), ptr(vbPrevMatched),
                                               frame2.grid(), int(winSize), levelOffset, minLevel, maxLevel, self.mfNNratio,
                                               ptr(m), self.device), "se2gpu_match_by_window")
        return n, m

    def MatchByProjection(self, kf: FrameView, kf_observed, mp_valid, mp_uv, mp_octave, mp_desc, winSize, levelOffset):
        kp = np.ascontiguousarray(kf.keyPointsUn, KP_DTYPE); d = np.ascontiguousarray(kf.descriptors, np.uint8)
        obs = np.ascontiguousarray(kf_observed, np.uint8); val = np.ascontiguousarray(mp_valid, np.uint8)
        uv = np.ascontiguousarray(mp_uv, np.float32); octv = np.ascontiguousarray(mp_octave, np.int32)
        md = np.ascontiguousarray(mp_desc, np.uint8)
        m = np.full(len(kp), -1, np.int32)
        n = check(lib().se2gpu_match_by_projection(ptr(kp), ptr(d), len(kp), ptr(obs), ptr(val), ptr(uv), len(val), ptr(octv),
                                                   ptr(md), kf.grid(), int(winSize), int(levelOffset), self.mfNNratio, ptr(m),
                                                   self.device), "se2gpu_match_by_projection")
        return n, m

    def SearchByBoW(self, kf1: dict, kf2: dict, bIfMPOnly=True):
        """kf = dict(angle, desc, has_mp, node (ascending), ptr, feat). Returns (nmatches, matches12 [-1 = none])."""
        keep = []

        def pack(k):
            a = [np.ascontiguousarray(k["angle"], np.float32), np.ascontiguousarray(k["desc"], np.uint8),
                 np.ascontiguousarray(k["has_mp"], np.uint8), np.ascontiguousarray(k["node"], np.int32),
                 np.ascontiguousarray(k["ptr"], np.int32), np.ascontiguousarray(k["feat"], np.int32)]
            keep.append(a)
            return BowKF(ptr(a[0]).value, ptr(a[1]).value, ptr(a[2]).value, len(a[0]), ptr(a[3]).value, len(a[3]),
                         ptr(a[4]).value, ptr(a[5]).value)
        b1, b2 = pack(kf1), pack(kf2)
        m = np.full(b1.n, -1, np.int32)
        n = check(lib().se2gpu_search_by_bow(C.byref(b1), C.byref(b2), int(bIfMPOnly), self.mfNNratio, int(self.mbCheckOrientation),
                                             ptr(m), self.device), "se2gpu_search_by_bow")
        return n, m
