// TEST STAND-IN for the project's own include/se2lam/EdgeSE2XYZ.h when it is compiled against the g2o mock: same classes
// and setters as reference include/se2lam/EdgeSE2XYZ.h:32-102 (the real header needs Eigen + g2o and stays untouched in
// the reference tree; computeError / linearizeOplus live on the GPU once the graph is handed to se2gpu::G2oGpuLevenberg).
#pragma once
#include <g2o/core/base_binary_edge.h>
#include <g2o/types/sba/types_six_dof_expmap.h>
#include <g2o/types/slam2d/vertex_se2.h>
namespace g2o {
class EdgeSE2XYZ : public BaseBinaryEdgeT<2, Vector2D> {
public:
    inline void setCameraParameter(g2o::CameraParameters* _cam) { cam = _cam; }
    inline void setExtParameter(const g2o::SE3Quat& _Tbc) { Tbc = _Tbc; Tcb = Tbc.inverse(); }
private:
    g2o::SE3Quat Tbc, Tcb;
    g2o::CameraParameters* cam = 0;
};
class PreEdgeSE2 : public BaseBinaryEdgeT<3, Vector3D> {};
}
