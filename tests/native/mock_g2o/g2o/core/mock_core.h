// TEST MOCK of the subset of g2o (tag 20160424_git) that include/se2lam/g2o_gpu_levenberg.h touches. g2o is an
// un-vendored dependency that is absent from this build container; this mock reproduces the NAMES and CALL SHAPES of the
// real API (SparseOptimizer::optimize driving OptimizationAlgorithm::init / solve(iteration), active vertex / edge sets,
// vertex and edge accessors) so that the binding compiles and runs here exactly as written for real g2o. No solver inside:
// OptimizationAlgorithmLevenberg::solve of the mock fails loudly (the stock CPU LM is what real g2o provides).
#pragma once
#include <cstdio>
#include <map>
#include <set>
#include <vector>

namespace g2o {

template <int N> struct VecN {
    double d[N];
    VecN() { for (int i = 0; i < N; ++i) d[i] = 0; }
    double& operator[](int i) { return d[i]; } double operator[](int i) const { return d[i]; }
    double& operator()(int i) { return d[i]; } double operator()(int i) const { return d[i]; }
};
template <int R, int C> struct MatRC {
    double d[R * C];
    MatRC() { for (int i = 0; i < R * C; ++i) d[i] = 0; }
    double& operator()(int r, int c) { return d[r * C + c]; } double operator()(int r, int c) const { return d[r * C + c]; }
};
typedef VecN<2> Vector2D; typedef VecN<3> Vector3D; typedef MatRC<2, 2> Matrix2D; typedef MatRC<3, 3> Matrix3D;

struct Rotation2Dd { double a; double angle() const { return a; } };
class SE2 {
public:
    SE2() : th_(0) {}
    SE2(double x, double y, double theta) : th_(theta) { t_[0] = x; t_[1] = y; }
    const Vector2D& translation() const { return t_; }
    Rotation2Dd rotation() const { return Rotation2Dd{th_}; }
    Vector3D toVector() const { Vector3D v; v[0] = t_[0]; v[1] = t_[1]; v[2] = th_; return v; }
private:
    Vector2D t_; double th_;
};
struct SE3Quat {
    Matrix3D R; Vector3D t;
    SE3Quat() { R(0, 0) = R(1, 1) = R(2, 2) = 1; }
    SE3Quat(const Matrix3D& R_, const Vector3D& t_) : R(R_), t(t_) {}
    const Vector3D& translation() const { return t; }
    struct Rot { const Matrix3D* m; Matrix3D toRotationMatrix() const { return *m; } };
    Rot rotation() const { return Rot{&R}; }
    SE3Quat inverse() const {
        SE3Quat o;
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) o.R(i, j) = R(j, i);
        for (int i = 0; i < 3; ++i) o.t[i] = -(o.R(i, 0) * t[0] + o.R(i, 1) * t[1] + o.R(i, 2) * t[2]);
        return o;
    }
};

struct HyperGraph {
    struct Vertex { virtual ~Vertex() {} int id() const { return _id; } void setId(int i) { _id = i; } int _id = -1; };
    struct Edge { virtual ~Edge() {} std::vector<Vertex*>& vertices() { return _vertices; } const std::vector<Vertex*>& vertices() const { return _vertices; }
                  std::vector<Vertex*> _vertices = std::vector<Vertex*>(2, (Vertex*)0); };
};
struct RobustKernel { virtual ~RobustKernel() {} double delta() const { return _delta; } void setDelta(double d) { _delta = d; } double _delta = 1.0; };
struct RobustKernelHuber : RobustKernel {};
struct Parameter { virtual ~Parameter() {} void setId(int i) { _id = i; } int id() const { return _id; } int _id = 0; };

struct OptimizableGraph : HyperGraph {
    struct Vertex : HyperGraph::Vertex { bool fixed() const { return _fixed; } void setFixed(bool f) { _fixed = f; } bool marginalized() const { return _marg; } void setMarginalized(bool m) { _marg = m; }
                                         bool _fixed = false, _marg = false; };
    struct Edge : HyperGraph::Edge { int level() const { return _level; } void setLevel(int l) { _level = l; } RobustKernel* robustKernel() const { return _rk; }
                                     void setRobustKernel(RobustKernel* k) { delete _rk; _rk = k; } void setVertex(int i, HyperGraph::Vertex* v) { _vertices[i] = v; }
                                     ~Edge() { delete _rk; } int _level = 0; RobustKernel* _rk = 0; };
    typedef std::vector<Vertex*> VertexContainer;
    typedef std::vector<Edge*> EdgeContainer;
};

template <class E> struct BaseVertexT : OptimizableGraph::Vertex { const E& estimate() const { return _est; } void setEstimate(const E& e) { _est = e; } E _est; };
template <int D, class M> struct BaseBinaryEdgeT : OptimizableGraph::Edge {
    const M& measurement() const { return _meas; } void setMeasurement(const M& m) { _meas = m; }
    const MatRC<D, D>& information() const { return _info; } void setInformation(const MatRC<D, D>& i) { _info = i; }
    M _meas; MatRC<D, D> _info;
};

class SparseOptimizer;
struct Solver {};
struct OptimizationAlgorithm {
    enum SolverResult { Terminate = 2, OK = 1, Fail = -1 };
    virtual ~OptimizationAlgorithm() {}
    virtual bool init(bool online = false) = 0;
    virtual SolverResult solve(int iteration, bool online = false) = 0;
    void setOptimizer(SparseOptimizer* o) { _optimizer = o; }
    SparseOptimizer* _optimizer = 0;
};
struct OptimizationAlgorithmLevenberg : OptimizationAlgorithm {
    explicit OptimizationAlgorithmLevenberg(Solver* s) : _solver(s) {}
    ~OptimizationAlgorithmLevenberg() { delete _solver; }
    virtual bool init(bool = false) { return true; }
    virtual SolverResult solve(int, bool = false) { std::fprintf(stderr, "mock g2o: the stock CPU Levenberg-Marquardt is not part of the mock\n"); return Fail; }
    int levenbergIteration() const { return _levenbergIterations; }
    Solver* _solver; int _levenbergIterations = 0;
};

class SparseOptimizer : public OptimizableGraph {
public:
    ~SparseOptimizer() { for (auto& kv : _vmap) delete kv.second; for (auto* e : _edges) delete e; for (auto& kv : _params) delete kv.second; delete _algorithm; }
    void setAlgorithm(OptimizationAlgorithm* a) { delete _algorithm; _algorithm = a; if (a) a->setOptimizer(this); }
    void setVerbose(bool v) { _verbose = v; }
    void setForceStopFlag(bool* f) { _forceStopFlag = f; }
    bool* forceStopFlag() const { return _forceStopFlag; }
    bool terminate() { return _forceStopFlag ? *_forceStopFlag : false; }
    OptimizableGraph::Vertex* vertex(int id) { auto it = _vmap.find(id); return it == _vmap.end() ? 0 : it->second; }
    bool addVertex(OptimizableGraph::Vertex* v) { if (_vmap.count(v->id())) return false; _vmap[v->id()] = v; return true; }
    bool addEdge(OptimizableGraph::Edge* e) { _edges.push_back(e); return true; }
    bool addParameter(Parameter* p) { _params[p->id()] = p; return true; }
    Parameter* parameter(int id) { auto it = _params.find(id); return it == _params.end() ? 0 : it->second; }
    bool initializeOptimization(int level = 0) {          // active sets: the edges of `level` and their vertices, by id
        _activeEdges.clear(); _activeVertices.clear();
        std::set<int> seen;
        for (auto* e : _edges) if (e->level() == level) { _activeEdges.push_back(e); for (auto* v : e->vertices()) seen.insert(v->id()); }
        for (int id : seen) _activeVertices.push_back(_vmap[id]);
        return !_activeVertices.empty();
    }
    const VertexContainer& activeVertices() const { return _activeVertices; }
    const EdgeContainer& activeEdges() const { return _activeEdges; }
    int optimize(int iterations, bool online = false) {   // SparseOptimizer::optimize (sparse_optimizer.cpp): init, then solve(i) until Terminate
        if (_activeVertices.empty()) { std::fprintf(stderr, "optimize: 0 vertices to optimize, maybe forgot to call initializeOptimization()\n"); return -1; }
        if (!_algorithm->init(online)) return -1;
        int cjIterations = 0;
        bool ok = true;
        for (int i = 0; i < iterations && !terminate() && ok; i++) {
            OptimizationAlgorithm::SolverResult result = _algorithm->solve(i, online);
            ok = (result == OptimizationAlgorithm::OK);
            ++cjIterations;
        }
        if (!ok && cjIterations == 1 && false) return 0;
        return cjIterations;
    }
private:
    std::map<int, OptimizableGraph::Vertex*> _vmap;
    std::vector<OptimizableGraph::Edge*> _edges;
    std::map<int, Parameter*> _params;
    VertexContainer _activeVertices; EdgeContainer _activeEdges;
    OptimizationAlgorithm* _algorithm = 0;
    bool* _forceStopFlag = 0; bool _verbose = false;
};

struct VertexSE2 : BaseVertexT<SE2> {};
struct VertexSBAPointXYZ : BaseVertexT<Vector3D> {};
struct CameraParameters : Parameter {
    CameraParameters(double f, const Vector2D& pp, double b) : focal_length(f), principle_point(pp), baseline(b) {}
    double focal_length; Vector2D principle_point; double baseline;
};

}  // namespace g2o
