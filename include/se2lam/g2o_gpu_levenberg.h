// se2gpu::G2oGpuLevenberg - the GPU local BA as a g2o::OptimizationAlgorithm (option i).
//
// For a se2lam build that HAS g2o: keep g2o for graph construction, for the SE(3) graphs of GlobalMapper / Localizer and
// for removeOutlierChi2, and swap only the algorithm typedef of include/se2lam/optimizer.h (reference :32):
//     typedef g2o::OptimizationAlgorithmLevenberg SlamAlgorithm;   ->   typedef se2gpu::G2oGpuLevenberg SlamAlgorithm;
// LocalMapper::localBA (src/LocalMapper.cpp:239-260) then runs unchanged: `new SlamAlgorithm(blockSolver)`,
// `optimizer.setAlgorithm(solver)`, `initializeOptimization(0)`, `optimize(Config::LOCAL_ITER)`. g2o's optimize() calls
// solve(iteration) once per LM iteration; at iteration 0 this class inspects the ACTIVE graph: if it consists only of
// VertexSE2 / VertexSBAPointXYZ (marginalised, free) vertices and EdgeSE2XYZ / PreEdgeSE2 edges it is exported to SoA and
// every solve(iteration) becomes se2gpu_ba_optimize_from(h, iteration, 1, ...) - one LM iteration with all its lambda
// trials, lambda / nu continued across calls exactly like OptimizationAlgorithmLevenberg's members - followed by writing
// the estimates back into the vertices. Any other graph is handed to the stock OptimizationAlgorithmLevenberg::solve.
//
// Two things g2o's public API does not expose are taken from where Map::loadLocalGraph takes them (src/Map.cpp:897, 1048):
// the camera (CameraParameters with id 0, registered by addCamPara) and the body-camera extrinsic Config::bTc, which
// EdgeSE2XYZ keeps private - pass it once with setExtrinsic(toSE3Quat(Config::bTc)).
// Compiles against real g2o (tag 20160424_git) and against tests/native/mock_g2o (this container has no g2o).
#ifndef SE2GPU_G2O_GPU_LEVENBERG_H
#define SE2GPU_G2O_GPU_LEVENBERG_H

#include <map>
#include <vector>

#include <g2o/core/optimization_algorithm_levenberg.h>
#include <g2o/core/robust_kernel_impl.h>
#include <g2o/core/sparse_optimizer.h>
#include <g2o/types/sba/types_six_dof_expmap.h>
#include <g2o/types/slam2d/vertex_se2.h>

#include <se2lam/EdgeSE2XYZ.h>      // with g2o present: forwards to the project's own header (include_next)
#include "../se2gpu.h"

namespace se2gpu {

class G2oGpuLevenberg : public g2o::OptimizationAlgorithmLevenberg {
public:
    explicit G2oGpuLevenberg(g2o::Solver* solver) : g2o::OptimizationAlgorithmLevenberg(solver), h_(0), gpu_(false), haveTbc_(false) {}
    virtual ~G2oGpuLevenberg() { if (h_) se2gpu_ba_destroy(h_); }

    void setExtrinsic(const g2o::SE3Quat& Tbc) { Tbc_ = Tbc; haveTbc_ = true; }
    bool onGpu() const { return gpu_; }
    const se2gpu_ba_iter_stats& lastStats() const { return st_; }

    virtual SolverResult solve(int iteration, bool online = false) {
        if (iteration == 0) gpu_ = exportGraph();
        if (!gpu_) return g2o::OptimizationAlgorithmLevenberg::solve(iteration, online);
        const int n = se2gpu_ba_optimize_from(h_, iteration, 1, (const volatile unsigned char*)_optimizer->forceStopFlag(), &st_, 0, 0);
        if (n < 0) return Fail;
        if (n == 0) return Terminate;                        // abort flag raised before the iteration started
        writeBack();
        _levenbergIterations = st_.trials;
        return st_.terminate ? Terminate : OK;               // 10 failed trials or rho == 0, like the stock LM
    }

private:
    // true when the active graph is the SE(2)-XYZ local-BA graph of Map::loadLocalGraph and has been loaded on the device
    bool exportGraph() {
        poses_.clear(); points_.clear();
        if (!haveTbc_) return false;
        const g2o::CameraParameters* cam = dynamic_cast<const g2o::CameraParameters*>(_optimizer->parameter(0));
        if (!cam) return false;
        std::map<int, int> pidx, lidx;
        std::vector<double> xp, xl, uv, info, om, oinf;
        std::vector<unsigned char> fixed;
        std::vector<int> ep, el, oi, oj;
        const g2o::OptimizableGraph::VertexContainer& av = _optimizer->activeVertices();
        for (size_t i = 0; i < av.size(); ++i) {             // ascending id: g2o's index mapping puts poses before points the same way
            if (g2o::VertexSE2* v = dynamic_cast<g2o::VertexSE2*>(av[i])) {
                pidx[v->id()] = (int)poses_.size(); poses_.push_back(v);
                xp.push_back(v->estimate().translation()[0]); xp.push_back(v->estimate().translation()[1]); xp.push_back(v->estimate().rotation().angle());
                fixed.push_back(v->fixed() ? 1 : 0);
            } else if (g2o::VertexSBAPointXYZ* p = dynamic_cast<g2o::VertexSBAPointXYZ*>(av[i])) {
                if (!p->marginalized() || p->fixed()) return false;
                lidx[p->id()] = (int)points_.size(); points_.push_back(p);
                for (int k = 0; k < 3; ++k) xl.push_back(p->estimate()[k]);
            } else {
                return false;                                // VertexSE3Expmap etc.: not this algorithm's graph
            }
        }
        double delta = -1.0;
        const g2o::OptimizableGraph::EdgeContainer& ae = _optimizer->activeEdges();
        for (size_t i = 0; i < ae.size(); ++i) {
            if (g2o::EdgeSE2XYZ* e = dynamic_cast<g2o::EdgeSE2XYZ*>(ae[i])) {
                const g2o::RobustKernelHuber* rk = dynamic_cast<const g2o::RobustKernelHuber*>(e->robustKernel());
                if (!rk) return false;                       // addEdgeSE2XYZ always sets a Huber kernel (optimizer.cpp:27-29)
                if (delta < 0) delta = rk->delta(); else if (delta != rk->delta()) return false;
                ep.push_back(pidx.at(e->vertices()[0]->id())); el.push_back(lidx.at(e->vertices()[1]->id()));
                uv.push_back(e->measurement()[0]); uv.push_back(e->measurement()[1]);
                info.push_back(e->information()(0, 0)); info.push_back(0.5 * (e->information()(0, 1) + e->information()(1, 0))); info.push_back(e->information()(1, 1));
            } else if (g2o::PreEdgeSE2* o = dynamic_cast<g2o::PreEdgeSE2*>(ae[i])) {
                if (o->robustKernel()) return false;
                oi.push_back(pidx.at(o->vertices()[0]->id())); oj.push_back(pidx.at(o->vertices()[1]->id()));
                for (int k = 0; k < 3; ++k) om.push_back(o->measurement()[k]);
                oinf.push_back(o->information()(0, 0)); oinf.push_back(o->information()(0, 1)); oinf.push_back(o->information()(0, 2));
                oinf.push_back(o->information()(1, 1)); oinf.push_back(o->information()(1, 2)); oinf.push_back(o->information()(2, 2));
            } else {
                return false;
            }
        }
        if (poses_.empty()) return false;
        if (delta < 0) delta = 1.0;
        const g2o::SE3Quat Tcb = Tbc_.inverse();             // EdgeSE2XYZ::setExtParameter (EdgeSE2XYZ.h:52)
        double T[12];
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) T[3 * r + c] = Tcb.rotation().toRotationMatrix()(r, c);
        for (int r = 0; r < 3; ++r) T[9 + r] = Tcb.translation()[r];
        const int P = (int)poses_.size(), L = (int)points_.size(), E = (int)ep.size(), O = (int)oi.size();
        if (h_) { se2gpu_ba_destroy(h_); h_ = 0; }
        h_ = se2gpu_ba_create(P, L > 0 ? L : 1, E > 0 ? E : 1, O > 0 ? O : 1, 0);
        if (!h_) return false;
        static const double zero3[3] = {0, 0, 0};
        static const int zero1[1] = {0};
        return se2gpu_ba_set_problem(h_, P, L, E, O, &xp[0], &fixed[0], L ? &xl[0] : zero3, E ? &ep[0] : zero1, E ? &el[0] : zero1, E ? &uv[0] : zero3,
                                     E ? &info[0] : zero3, O ? &oi[0] : zero1, O ? &oj[0] : zero1, O ? &om[0] : zero3, O ? &oinf[0] : zero3,
                                     cam->focal_length, cam->principle_point[0], cam->principle_point[1], T, delta) == SE2GPU_OK;
    }

    void writeBack() {                                        // SparseOptimizer::update -> oplus, done on the device; estimates back into the graph
        std::vector<double> xp(3 * poses_.size()), xl(3 * points_.size() + 3);
        se2gpu_ba_get(h_, &xp[0], &xl[0]);
        for (size_t i = 0; i < poses_.size(); ++i) poses_[i]->setEstimate(g2o::SE2(xp[3 * i], xp[3 * i + 1], xp[3 * i + 2]));
        for (size_t j = 0; j < points_.size(); ++j) { g2o::Vector3D p; for (int k = 0; k < 3; ++k) p[k] = xl[3 * j + k]; points_[j]->setEstimate(p); }
    }

    se2gpu_ba* h_;
    bool gpu_, haveTbc_;
    g2o::SE3Quat Tbc_;
    se2gpu_ba_iter_stats st_;
    std::vector<g2o::VertexSE2*> poses_;
    std::vector<g2o::VertexSBAPointXYZ*> points_;
};

}  // namespace se2gpu

#endif
