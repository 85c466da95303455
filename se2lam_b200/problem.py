"""SoA description of one SE(2)-XYZ local-BA window: the argument bundle of se2gpu_ba_set_problem."""
from __future__ import annotations

import dataclasses

import numpy as np


@dataclasses.dataclass
class BAProblem:
    """SoA description of one SE(2)-XYZ local-BA window == the arguments of se2gpu_ba_set_problem.

    Vertex/edge semantics follow Map::loadLocalGraph (reference src/Map.cpp:891-1053):
    poses are VertexSE2 (x,y,theta) of Twb; points are marginalised VertexSBAPointXYZ;
    EdgeSE2XYZ carries uv + a full symmetric 2x2 information (stored xx,xy,yy);
    PreEdgeSE2 carries a 3-vector measurement and a full symmetric 3x3 information
    (stored row-major upper: 00,01,02,11,12,22).
    """
    poses: np.ndarray        # [P,3] f64
    fixed: np.ndarray        # [P] u8
    points: np.ndarray       # [L,3] f64
    edge_pose: np.ndarray    # [E] i32
    edge_point: np.ndarray   # [E] i32
    uv: np.ndarray           # [E,2] f64
    info: np.ndarray         # [E,3] f64 (xx, xy, yy)
    odo_i: np.ndarray        # [O] i32
    odo_j: np.ndarray        # [O] i32
    odo_meas: np.ndarray     # [O,3] f64
    odo_info: np.ndarray     # [O,6] f64
    fx: float
    cx: float
    cy: float
    Tcb: np.ndarray          # [12] f64: row-major 3x3 Rcb then tcb
    huber_delta: float
    # ground truth (not part of the problem; for self-consistency tests)
    gt_poses: np.ndarray | None = None
    gt_points: np.ndarray | None = None

    @property
    def P(self): return self.poses.shape[0]
    @property
    def L(self): return self.points.shape[0]
    @property
    def E(self): return self.edge_pose.shape[0]
    @property
    def O(self): return self.odo_i.shape[0]
