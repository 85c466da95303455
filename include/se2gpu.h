/* se2gpu — C ABI of the B200-native hot paths of se2lam (ORB front-end, SE(2)-XYZ local BA).
 *
 * Plain C, plain pointers and sizes, int status codes, no exceptions across the boundary.
 * The reference (izhengfan/se2lam) has no FFI layer of its own: the seam is the C++ symbols
 * listed below, which the header shims in include/se2lam/ forward to these entry points
 * (INTEGRATION.md shows the binding).  Every function is implemented by hand-written sm_100a
 * CUDA in se2lam_b200/csrc; there is no CPU fallback — a call without a usable CUDA device
 * returns SE2GPU_ERR_NO_DEVICE.
 *
 *   entry point                      replaces (reference file:line)
 *   -------------------------------  -------------------------------------------------------------
 *   se2gpu_orb_create                se2lam::ORBextractor::ORBextractor       src/ORBextractor.cpp:463-520
 *   se2gpu_orb_extract[_device]      se2lam::ORBextractor::operator()         src/ORBextractor.cpp:727-788
 *                                    (ComputePyramid :790-831, ComputeKeyPoints :531-716,
 *                                     IC_Angle :130-157, computeOrbDescriptor :160-200)
 *   se2gpu_hamming_distance          se2lam::ORBmatcher::DescriptorDistance   src/ORBmatcher.cpp:110-126
 *   se2gpu_matcher_create            se2lam::ORBmatcher::ORBmatcher           src/ORBmatcher.cpp:49-51
 *   se2gpu_[matcher_]match_by_window[_device]      ORBmatcher::MatchByWindow  src/ORBmatcher.cpp:278-381
 *                                    (+ Frame grid / GetFeaturesInArea        src/Frame.cpp:64-77, 209-286)
 *   se2gpu_[matcher_]match_by_projection[_device]  ORBmatcher::MatchByProjection  src/ORBmatcher.cpp:383-454
 *   se2gpu_[matcher_]search_by_bow   se2lam::ORBmatcher::SearchByBoW          src/ORBmatcher.cpp:128-276
 *   se2gpu_ba_set_problem            Map::loadLocalGraph -> addVertexSE2 / addEdgeSE2 / addVertexSBAXYZ /
 *                                    addEdgeSE2XYZ / addCamPara               src/Map.cpp:891-1053, src/optimizer.cpp:17-62,207-215,316-324
 *                                    + SparseOptimizer::initializeOptimization(0)   src/LocalMapper.cpp:259
 *   se2gpu_ba_optimize               SlamOptimizer::optimize(Config::LOCAL_ITER)    src/LocalMapper.cpp:260
 *                                    (EdgeSE2XYZ::computeError/linearizeOplus src/EdgeSE2XYZ.cpp:61-106,
 *                                     PreEdgeSE2 include/se2lam/EdgeSE2XYZ.h:62-102, g2o LM/Schur/Cholesky/Huber [upstream])
 *   se2gpu_ba_get[_f32]              estimateVertexSE2 / estimateVertexSBAXYZ src/optimizer.cpp:45-50, 549-554
 *                                    (+ the float write-back of Map::optimizeLocalGraph  src/Map.cpp:768-779)
 *   se2gpu_ba_build_information      per-edge Omega of Map::loadLocalGraph    src/Map.cpp:1024-1049
 *   se2gpu_voc_create / _transform   DBoW2 TemplatedVocabulary::transform     Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1220-1262
 *   se2gpu_median_descriptor         MapPoint::updateMainKFandDescriptor      src/MapPoint.cpp:228-272
 */
#ifndef SE2GPU_H
#define SE2GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SE2GPU_OK 0
#define SE2GPU_ERR_NO_DEVICE -1
#define SE2GPU_ERR_CUDA -2
#define SE2GPU_ERR_INVALID -3
#define SE2GPU_ERR_CAPACITY -4

/* ------------------------------------------------------------------------------------------ common */
int se2gpu_device_count(void);
/* last error message of the calling thread ("" if none) */
const char* se2gpu_last_error(void);
/* number of kernel launches issued by this library in this process so far (bench.py's gpu_launches) */
unsigned long long se2gpu_launch_count(void);

/* ------------------------------------------------------------------------------------------ ORB */
typedef struct se2gpu_orb se2gpu_orb;

/* bit-identical to cv::KeyPoint (28 bytes) so results can be copied straight into a
 * std::vector<cv::KeyPoint> */
typedef struct se2gpu_keypoint {
    float x, y;      /* pt, in level-0 pixel coordinates */
    float size;      /* PATCH_SIZE * scale(level), truncated to int */
    float angle;     /* degrees, [0,360) */
    float response;  /* FAST score */
    int octave;      /* pyramid level */
    int class_id;    /* -1 */
} se2gpu_keypoint;

/* ORBextractor(nfeatures, scaleFactor, nlevels, FAST_SCORE, fastTh) for frames up to max_w x max_h,
 * batches of up to max_batch frames, on CUDA device `device`. Returns NULL on failure. */
se2gpu_orb* se2gpu_orb_create(int nfeatures, float scale_factor, int nlevels, int fast_th, int max_w, int max_h,
                              int max_batch, int device);
void se2gpu_orb_destroy(se2gpu_orb* h);

/* operator() over a batch of n frames held in HOST memory (CV_8UC1, row stride `stride` bytes, frame i at
 * imgs + i*frame_stride). Synchronous: copies frames in, runs the extractor, copies results out.
 *   kps    [n * nfeatures]      frame i's keypoints start at kps + i*nfeatures
 *   desc   [n * nfeatures * 32] 256-bit rBRIEF, row i*nfeatures + k belongs to kps[i*nfeatures + k]
 *   counts [n]                  number of keypoints of each frame (<= nfeatures)
 * An empty image (w<=0 || h<=0 || imgs==NULL) returns SE2GPU_OK with all counts 0 (the reference returns
 * silently, ORBextractor.cpp:730-731). */
int se2gpu_orb_extract(se2gpu_orb* h, const uint8_t* imgs, int n, int w, int hgt, int stride, size_t frame_stride,
                       se2gpu_keypoint* kps, uint8_t* desc, int* counts);

/* Asynchronous, two-deep form of se2gpu_orb_extract for streams of batches: _submit enqueues the copies and kernels of one
 * batch and returns at once; _wait blocks until the OLDEST submitted batch is complete and its results are in the buffers
 * handed to _submit. Up to two batches may be in flight, so the host->device copy of batch k+1 and the device->host copy of
 * batch k-1 overlap the kernels of batch k (a second internal context is created on first use). The buffers of a submitted
 * batch must stay valid and untouched until its _wait returns; page-locked buffers make the copies true DMA transfers. */
int se2gpu_orb_submit(se2gpu_orb* h, const uint8_t* imgs, int n, int w, int hgt, int stride, size_t frame_stride,
                      se2gpu_keypoint* kps, uint8_t* desc, int* counts);
int se2gpu_orb_wait(se2gpu_orb* h);

/* Same with DEVICE buffers (frames already resident in HBM, results left in HBM); asynchronous on
 * `stream` (a cudaStream_t passed as void*; NULL = the default stream). */
int se2gpu_orb_extract_device(se2gpu_orb* h, const uint8_t* d_imgs, int n, int w, int hgt, int stride,
                              size_t frame_stride, se2gpu_keypoint* d_kps, uint8_t* d_desc, int* d_counts, void* stream);

/* parity/debug: geometry and contents of pyramid level `level` of frame `frame` from the last extract.
 * out must hold (h+32)*pitch bytes where pitch/h come from se2gpu_orb_level_dims. blurred!=0 returns the
 * Gaussian-blurred plane the descriptors were sampled from. */
int se2gpu_orb_level_dims(se2gpu_orb* h, int level, int* w, int* hgt, int* pitch);
int se2gpu_orb_get_level(se2gpu_orb* h, int frame, int level, int blurred, uint8_t* out);

/* per-kernel device timing (CUDA events on the launching stream) for bench.py's roofline line.
 * groups: 0 pyramid (orb_pyr0 + orb_resize), 1 orb_fast_cells, 2 orb_select, 3 orb_blur, 4 orb_orient_describe.
 * enable!=0 starts/restarts accumulation; read returns accumulated milliseconds and launch counts per group
 * (synchronises the events it reads). */
#define SE2GPU_ORB_PROFILE_GROUPS 5
int se2gpu_orb_profile(se2gpu_orb* h, int enable);
int se2gpu_orb_profile_read(se2gpu_orb* h, double* ms, int* launches);

/* Folds the lens undistortion that Frame::Frame applies right before the extractor (reference src/Frame.cpp:22:
 * cv::undistort(im, img, Config::Kcam, Config::Dcam)) into the pyramid's level 0: after this call the frames handed to
 * se2gpu_orb_extract / _extract_device are RAW frames and the keypoints/descriptors are those of the undistorted frame,
 * bit-identical to undistort-then-extract. K: 3x3 row-major float32 camera matrix (Config::Kcam), dist: 0/4/5/8/12
 * float32 coefficients (k1 k2 p1 p2 [k3 [k4 k5 k6 [s1 s2 s3 s4]]], Config::Dcam). K == NULL switches it off. */
int se2gpu_orb_set_undistort(se2gpu_orb* h, const float* K, const float* dist, int ndist);
/* Test hook (host only, no GPU needed): the fixed-point map se2gpu_orb_set_undistort builds for a w x h frame;
 * m1 [h*w*2] int16 integer source coordinates (x, y), m2 [h*w] uint16 fraction index (fy*32 + fx). */
int se2gpu_orb_debug_undistort_map(const float* K, const float* dist, int ndist, int w, int h, int16_t* m1, uint16_t* m2);

/* Test hook for the device selection primitive behind KeyPointsFilter::retainBest (ORBextractor.cpp:692, :708):
 * runs the warp-cooperative std::nth_element on `count` independent lists of packed records (score in bits 31..24)
 * stored back to back in HOST memory; list k is values[offsets[k] .. offsets[k+1]) and is permuted in place exactly
 * as std::nth_element(begin, begin + nth[k], end, score-greater) of libstdc++ would. */
int se2gpu_orb_debug_nth_element(uint32_t* values, const int* offsets, const int* nth, int count, int device);

/* ------------------------------------------------------------------------------------------ matcher */
/* DescriptorDistance for n pairs of 32-byte descriptors in HOST memory: out[i] = popcount(a_i ^ b_i) */
int se2gpu_hamming_distance(const uint8_t* a, const uint8_t* b, int n, int* out, int device);

/* Keypoint grid of Frame (64 x 48 cells over [minX,maxX) x [minY,maxY)): inv_w = 64/(maxX-minX) etc. */
typedef struct se2gpu_grid_params {
    float min_x, min_y, inv_w, inv_h;
} se2gpu_grid_params;

/* Matcher context: owns every device buffer the matchers need for up to max_queries query items (frame-1 keypoints / map
 * points / KF1 features) against up to max_db database keypoints (candidate table max_queries x max_db x 8 bytes), a
 * stream and page-locked staging for the host entry points. No allocation happens per call. One context per calling
 * thread (like the reference's stack-allocated ORBmatcher objects, the calls are not re-entrant on one context). */
typedef struct se2gpu_matcher se2gpu_matcher;
se2gpu_matcher* se2gpu_matcher_create(int max_queries, int max_db, int device);
void se2gpu_matcher_destroy(se2gpu_matcher* m);

/* MatchByWindow(frame1, frame2, vbPrevMatched, winSize, vnMatches12, levelOffset, minLevel, maxLevel) with nnratio =
 * ORBmatcher::mfNNratio on DEVICE buffers, asynchronous on `stream` (cudaStream_t as void*, NULL = default stream): the
 * keypoint / descriptor buffers are the ones se2gpu_orb_extract_device wrote (reference call chain Track.cpp:129-132 runs
 * the extractor and MatchByWindow back to back), nothing crosses PCIe. n1 / n2 are capacities; d_n1 / d_n2 (may be NULL)
 * point to the actual counts in device memory (e.g. the extractor's d_counts entries). d_prev [n1*2] is vbPrevMatched,
 * updated in place; d_matches12 [n1]; d_nmatches (may be NULL) receives the match count. */
int se2gpu_match_by_window_device(se2gpu_matcher* m, const se2gpu_keypoint* d_kp1, const uint8_t* d_desc1, int n1, const int* d_n1,
                                  const se2gpu_keypoint* d_kp2, const uint8_t* d_desc2, int n2, const int* d_n2, float* d_prev,
                                  se2gpu_grid_params grid, int win_size, int level_offset, int min_level, int max_level,
                                  float nnratio, int* d_matches12, int* d_nmatches, void* stream);
/* vbPrevMatched initialisation (Track.cpp:113-116: the reference frame's keypoint positions): d_xy[2i..] = d_kp[i].pt */
int se2gpu_keypoints_to_points_device(const se2gpu_keypoint* d_kp, int n, const int* d_n, float* d_xy, void* stream);

/* MatchByProjection on DEVICE buffers (same flattening of the object graph as se2gpu_match_by_projection below). */
int se2gpu_match_by_projection_device(se2gpu_matcher* m, const se2gpu_keypoint* d_kf_kp, const uint8_t* d_kf_desc, int n_kf,
                                      const int* d_n_kf, const uint8_t* d_kf_observed, const uint8_t* d_mp_valid,
                                      const float* d_mp_uv, int n_mp, const int* d_mp_octave, const uint8_t* d_mp_desc,
                                      se2gpu_grid_params grid, int win_size, int level_offset, float nnratio,
                                      int* d_matches_idx_mp, int* d_nmatches, void* stream);

/* HOST-buffer entry points on an explicit context (synchronous; one stream synchronisation per call).
 * MatchByWindow: prev [n1*2] is vbPrevMatched, updated in place. matches12 [n1]. Returns the number of matches (>=0)
 * or a negative error. */
int se2gpu_matcher_match_by_window(se2gpu_matcher* m, const se2gpu_keypoint* kp1, const uint8_t* desc1, int n1,
                                   const se2gpu_keypoint* kp2, const uint8_t* desc2, int n2, float* prev,
                                   se2gpu_grid_params grid, int win_size, int level_offset, int min_level, int max_level,
                                   float nnratio, int* matches12);
/* MatchByProjection(pNewKF, localMPs, winSize, levelOffset, vMatchesIdxMP), object graph flattened:
 *   mp_valid[i]  = !isNull && isGoodPrl && !pNewKF->hasObservation(pMP) && inImgBound(predictUV)
 *   mp_uv[i]     = predictUV;  mp_octave[i] = mMainOctave;  mp_desc = mMainDescriptor
 *   kf_observed[k] = pNewKF->hasObservation(k)
 * matches_idx_mp [n_kf]. Returns the number of matches. */
int se2gpu_matcher_match_by_projection(se2gpu_matcher* m, const se2gpu_keypoint* kf_kp, const uint8_t* kf_desc, int n_kf,
                                       const uint8_t* kf_observed, const uint8_t* mp_valid, const float* mp_uv, int n_mp,
                                       const int* mp_octave, const uint8_t* mp_desc, se2gpu_grid_params grid, int win_size,
                                       int level_offset, float nnratio, int* matches_idx_mp);
/* SearchByBoW(pKF1, pKF2, mapMatches12, bIfMPOnly); each DBoW2::FeatureVector flattened to ascending node
 * ids + CSR feature lists (ptr has n_node+1 entries). matches12 [n1], -1 = unmatched. */
typedef struct se2gpu_bow_kf {
    const float* angle;     /* keyPointsUn[i].angle */
    const uint8_t* desc;    /* [n*32] */
    const uint8_t* has_mp;  /* GetMapPointMatches()[i] && !isNull */
    int n;
    const int* node;        /* ascending node ids */
    int n_node;
    const int* ptr;         /* [n_node+1] */
    const int* feat;        /* feature indices */
} se2gpu_bow_kf;
int se2gpu_matcher_search_by_bow(se2gpu_matcher* m, const se2gpu_bow_kf* kf1, const se2gpu_bow_kf* kf2, int mp_only,
                                 float nnratio, int check_orientation, int* matches12);

/* The same three with a context per device created on first use and kept for the life of the process. */
int se2gpu_match_by_window(const se2gpu_keypoint* kp1, const uint8_t* desc1, int n1, const se2gpu_keypoint* kp2,
                           const uint8_t* desc2, int n2, float* prev, se2gpu_grid_params grid, int win_size,
                           int level_offset, int min_level, int max_level, float nnratio, int* matches12, int device);
int se2gpu_match_by_projection(const se2gpu_keypoint* kf_kp, const uint8_t* kf_desc, int n_kf,
                               const uint8_t* kf_observed, const uint8_t* mp_valid, const float* mp_uv, int n_mp,
                               const int* mp_octave, const uint8_t* mp_desc, se2gpu_grid_params grid, int win_size,
                               int level_offset, float nnratio, int* matches_idx_mp, int device);
int se2gpu_search_by_bow(const se2gpu_bow_kf* kf1, const se2gpu_bow_kf* kf2, int mp_only, float nnratio,
                         int check_orientation, int* matches12, int device);

/* per-kernel device timing for bench.py; groups: 0 k_grid_build, 1 k_candidates, 2 k_resolve, 3 k_fallback_* */
#define SE2GPU_MATCHER_PROFILE_GROUPS 4
int se2gpu_matcher_profile(se2gpu_matcher* m, int enable);
int se2gpu_matcher_profile_read(se2gpu_matcher* m, double* ms, int* launches);
/* diagnostics of the last resolve on this context: speculative rounds it took, and whether the sequential fallback ran */
int se2gpu_matcher_last_rounds(se2gpu_matcher* m, int* rounds, int* used_fallback);

/* ------------------------------------------------------------------------------------------ bag of words */
/* DBoW2 vocabulary tree (TemplatedVocabulary<FORB::TDescriptor, FORB>, reference Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h)
 * flattened: node 0 is the root (m_nodes[0]); node i has the 32-byte descriptor node_desc[32 i], the children
 * children[child_ptr[i] .. child_ptr[i+1]) in m_nodes[i].children order, and - for a leaf - word_id[i] >= 0 and its weight
 * (inner nodes: word_id -1). levels = m_L. */
typedef struct se2gpu_voc se2gpu_voc;
se2gpu_voc* se2gpu_voc_create(int n_nodes, const uint8_t* node_desc, const int* child_ptr, const int* children,
                              const int* word_id, const double* weight, int levels, int device);
void se2gpu_voc_destroy(se2gpu_voc* v);
/* transform(feature, word_id, weight, &nid, levelsup) (TemplatedVocabulary.h:1220-1262) for n descriptors [n*32]:
 * word_id [n], weight [n], node_id [n] (may be NULL) = the node at level m_L - levelsup on the feature's path (0 = root
 * when that level is <= 0; -1 when the leaf is shallower - the reference leaves *nid unset there). The BowVector /
 * FeatureVector assembly of transform(features, v, fv, levelsup) (:1150-1216: v.addWeight(id, w) in feature order,
 * fv.addFeature(nid, i), L1 normalisation) is a sorted-map accumulation over these n triples and stays with the caller.
 * KeyFrame::ComputeBoW (src/KeyFrame.cpp:244-254) uses levelsup = 4. HOST buffers, synchronous. */
int se2gpu_voc_transform(se2gpu_voc* v, const uint8_t* desc, int n, int levelsup, int* word_id, double* weight, int* node_id);
/* same on DEVICE buffers (e.g. the extractor's d_desc), asynchronous on `stream` */
int se2gpu_voc_transform_device(se2gpu_voc* v, const uint8_t* d_desc, int n, int levelsup, int* d_word_id, double* d_weight,
                                int* d_node_id, void* stream);

/* MapPoint::updateMainKFandDescriptor (reference src/MapPoint.cpp:228-272) for M map points at once: the descriptors of map
 * point m's observations are desc rows ptr[m] .. ptr[m+1]); best_idx [m] = index (within the point's list) of the descriptor
 * with the least median Hamming distance to the others - median = element int(0.5*(N-1)) of the sorted distances incl. the
 * zero self-distance, first index wins ties - and best_median [m] (may be NULL) that median. HOST buffers. */
int se2gpu_median_descriptor(const uint8_t* desc, const int* ptr, int M, int* best_idx, int* best_median, int device);

/* ------------------------------------------------------------------------------------------ local BA */
typedef struct se2gpu_ba se2gpu_ba;

/* one entry per completed LM iteration (OptimizationAlgorithmLevenberg::solve call) */
typedef struct se2gpu_ba_iter_stats {
    double chi2_before; /* activeRobustChi2 at the start of the iteration */
    double chi2_after;  /* after the last accepted trial (== chi2_before if none) */
    double lambda;      /* damping after the iteration */
    double rho;         /* gain ratio of the last trial */
    int trials;         /* lambda trials used (1..10) */
    int accepted;       /* 1 if a trial was accepted */
    int terminate;      /* 1 if LM returned Terminate (10 failed trials or rho==0) */
    int pad;
} se2gpu_ba_iter_stats;

/* capacity-sized solver context on CUDA device `device` */
se2gpu_ba* se2gpu_ba_create(int max_poses, int max_points, int max_edges, int max_odo, int device);
void se2gpu_ba_destroy(se2gpu_ba* h);

/* Loads one local-BA window (what Map::loadLocalGraph + initializeOptimization(0) build):
 *   poses   [P*3]  VertexSE2 estimates (x,y,theta) of Twb; vertex id = index
 *   fixed   [P]    setFixed flags
 *   points  [L*3]  VertexSBAPointXYZ estimates (marginalised); landmarks without edges stay untouched
 *   edge_pose/edge_point [E], uv [E*2], info [E*3] (xx,xy,yy of the symmetric 2x2 information),
 *   odo_i/odo_j [O] PreEdgeSE2 vertices (error = Ri^T(rj-ri)-m), odo_meas [O*3], odo_info [O*6] (00,01,02,11,12,22)
 *   fx,cx,cy       CamPara (single focal length), Tcb [12] = row-major Rcb then tcb (SE3 of setExtParameter, inverted)
 *   huber_delta    RobustKernelHuber delta of every EdgeSE2XYZ
 * All HOST pointers; copies and re-indexes synchronously. */
int se2gpu_ba_set_problem(se2gpu_ba* h, int P, int L, int E, int O, const double* poses, const uint8_t* fixed,
                          const double* points, const int* edge_pose, const int* edge_point, const double* uv,
                          const double* info, const int* odo_i, const int* odo_j, const double* odo_meas,
                          const double* odo_info, double fx, double cx, double cy, const double* Tcb,
                          double huber_delta);

/* optimize(max_iters): returns the number of LM iterations performed (like SparseOptimizer::optimize) or a
 * negative error. stop_flag (may be NULL) is polled between trials (setForceStopFlag). stats (may be NULL)
 * receives one entry per iteration. trace_poses [max_iters*P*3] / trace_points [max_iters*L*3] (may be NULL)
 * receive the estimates after each iteration (parity tests). */
int se2gpu_ba_optimize(se2gpu_ba* h, int max_iters, const volatile unsigned char* stop_flag,
                       se2gpu_ba_iter_stats* stats, double* trace_poses, double* trace_points);

/* The same for ONE g2o `solve(iteration)`-style slice of an optimisation: runs LM iterations first_iteration ..
 * first_iteration + max_iters - 1. lambda is initialised (1e-5 max|diag H|) at iteration 0 only; a call with first_iteration > 0
 * continues the lambda / nu schedule where the previous call on this context stopped, so optimize_from(0, 1), (1, 1), ... (9, 1)
 * is bit-identical to optimize(10). This is what a g2o::OptimizationAlgorithm subclass needs (option i). */
int se2gpu_ba_optimize_from(se2gpu_ba* h, int first_iteration, int max_iters, const volatile unsigned char* stop_flag,
                            se2gpu_ba_iter_stats* stats, double* trace_poses, double* trace_points);

/* restore the estimates loaded by the last se2gpu_ba_set_problem (device-side copy; lets a caller re-run
 * optimize on the same window without re-uploading it) */
int se2gpu_ba_reset(se2gpu_ba* h);

/* current estimates -> host (poses [P*3], points [L*3]) */
int se2gpu_ba_get(se2gpu_ba* h, double* poses, double* points);

/* Map::optimizeLocalGraph's write-back (reference src/Map.cpp:768-779) in the reference's storage type: poses [P*3] as
 * Se2(float x, float y, float theta) - theta narrowed to float, then normalised like Se2::Se2 (Config.cpp:194-195) - and
 * points [L*3] as cv::Point3f (toCvPt3f). Narrowed on the device; either pointer may be NULL. */
int se2gpu_ba_get_f32(se2gpu_ba* h, float* poses, float* points);

/* Map::loadLocalGraph's per-edge information matrix (reference src/Map.cpp:1024-1049), evaluated on the device from the
 * reference's own float data:
 *   Omega_e = (sigma_rot J_rotxy J_rotxy^T + sigma_z J_z J_z^T + mvLevelSigma2[octave_e] I)^-1,
 *   J_pi from view_mp[e] = pKF->mViewMPs[ftrIdx] and fx = Config::fxCam, Rcw = rows of pKF->Tcw(0:3,0:3),
 *   J_rotxy = (J_pi Rcw skew(lw - (Twb.x, Twb.y, 0)))[:, 0:2], J_z = -(J_pi Rcw)[:, 2],
 *   sigma_rot = 1/xrot_info, sigma_z = 1/z_info as float (Config::PLANEMOTION_XROT_INFO / _Z_INFO).
 * kf_Rcw [P*9], kf_twb_xy [P*2], mp_pos [L*3], view_mp [E*3], octave [E], level_sigma2 [nlevels]; info [E*3] receives
 * (xx, xy, yy) in double - the `info` argument of se2gpu_ba_set_problem. HOST pointers. */
int se2gpu_ba_build_information(int P, int L, int E, const float* view_mp, const int* edge_pose, const int* edge_point,
                                const int* octave, const float* kf_Rcw, const float* kf_twb_xy, const float* mp_pos,
                                const float* level_sigma2, int nlevels, float fx, float xrot_info, float z_info, double* info,
                                int device);

/* Multi-GPU: this context owns the landmarks j with j % world == rank (call before set_problem); the reduced
 * pose system [S | b | chi2 | scale] is summed over ranks once per LM trial through `allreduce`, which must
 * sum (op 0) or max (op 1) `count` doubles at device pointer `buf` in place across ranks, ordered on `stream`. */
typedef int (*se2gpu_allreduce_fn)(void* user, double* buf, size_t count, int op, void* stream);
int se2gpu_ba_set_shard(se2gpu_ba* h, int rank, int world, se2gpu_allreduce_fn allreduce, void* user);
/* Sharded runs on one NVLink node without any collective library on the data path: when the reduced system fits one CTA's
 * shared memory (<= 52 free poses) the whole optimize() of every rank is ONE persistent cooperative kernel, and the ranks'
 * kernels exchange through peer memory - twice per lambda-trial: the partial reduced systems [S | b] (every CTA of every rank
 * sums the ranks' buffers slice by slice over NVLink) and the scalars [chi2, scale, abort] - with flag words in peer memory
 * as the only synchronisation. The callback of se2gpu_ba_set_shard is then unused (it stays the path for larger windows).
 * One process per GPU: after se2gpu_ba_set_shard every rank calls _peer_export, the handles (SE2GPU_BA_PEER_HANDLE_BYTES
 * each, CUDA IPC) are all-gathered by the caller in rank order and passed to _peer_import. One process driving several
 * contexts (one per GPU, or several on one GPU for tests): _peer_attach_local on the array of contexts in rank order; every
 * context then needs its own host thread and stream, since the ranks' optimize() calls must run concurrently.
 * se2gpu_ba_optimize is a COLLECTIVE call in a sharded run (same arguments on every rank); stop_flag may differ per rank -
 * the abort decision is OR-ed over the ranks, so all of them stop after the same trial. A rank that does not show up within
 * SE2GPU_BA_PEER_TIMEOUT_S (default 10 s) makes the others return SE2GPU_ERR_CUDA instead of hanging. */
#define SE2GPU_BA_PEER_HANDLE_BYTES 128
int se2gpu_ba_peer_export(se2gpu_ba* h, void* handle_out);
int se2gpu_ba_peer_import(se2gpu_ba* h, const void* handles, int world);
int se2gpu_ba_peer_attach_local(se2gpu_ba** contexts, int world);
/* stream all BA work is enqueued on (cudaStream_t as void*); NULL = default stream */
int se2gpu_ba_set_stream(se2gpu_ba* h, void* stream);

/* parity/debug: linearise at the current estimate, Schur-reduce with damping `lambda`, solve; copy out whatever
 * is non-NULL. Hpp, S: [n*n] row-major (lower triangle valid), n = 3*#free poses; bp, bs, dx_p: [n];
 * Hll [L*9], bl [L*3], dx_l [L*3]; Hpl [E*9] (3x3 per edge, rows = pose, cols = point; original edge order).
 * Returns n (>=0) or a negative error. Does not change the estimates. */
/* execution mode of se2gpu_ba_optimize: 0 = auto (persistent cooperative kernel when available: single GPU, reduced system
 * fits one CTA's shared memory), 1 = one kernel per phase (always used for sharded runs), 2 = persistent or fail */
int se2gpu_ba_set_mode(se2gpu_ba* h, int mode);

/* per-kernel device timing for bench.py's roofline line; groups: 0 ba_linearize (+chi2 evaluation), 1 ba_pose_reduce,
 * 2 ba_lm_prep, 3 ba_schur, 4 ba_chol_solve, 5 ba_backsub_update, 6 ba_iter_begin/ba_decide, 7 ba_persistent (whole optimize),
 * 8 TMA staging of S inside the persistent kernel. In persistent mode groups 0-6 and 8 are in-kernel phase times of CTA 0. */
#define SE2GPU_BA_PROFILE_GROUPS 9
int se2gpu_ba_profile(se2gpu_ba* h, int enable);
int se2gpu_ba_profile_read(se2gpu_ba* h, double* ms, int* launches);

int se2gpu_ba_debug_system(se2gpu_ba* h, double lambda, double* chi2, double* Hpp, double* bp, double* Hll, double* bl,
                           double* Hpl, double* S, double* bs, double* dx_p, double* dx_l);

#ifdef __cplusplus
}
#endif
#endif /* SE2GPU_H */
