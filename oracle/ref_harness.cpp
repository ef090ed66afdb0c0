// TEST INFRASTRUCTURE — NOT PART OF THE PRODUCT PATH.
//
// C-ABI harness around the UNMODIFIED reference sources (compiled from /root/reference/src by
// oracle/Makefile into oracle/_ref/libsmplpp_ref.so).  It exists so that tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs can (1) pin the Python restatement in
// oracle/smpl_oracle.py against the real libtorch implementation, and (2) time the reference's own CPU path.
//
// What is the reference and what is restated here:
//   * forward pass, normals, IkTask geometry, VPoser decoder: the reference's own classes are called
//     (smplpp::SMPL::launch src/SMPL.cpp:671-737, smplpp::IkTask src/IkTask.cpp, smplpp::VPoserDecoder
//     src/VPoser.cpp, the four pipeline modules).
//   * the IK iteration lives inline in the reference's ROS executable (node/node.cpp:645-1002), which cannot
//     be compiled here (ROS, Eigen, QpSolverCollection, igl, ezc3d are absent).  ref_ik_iteration() restates
//     node/node.cpp:753-968 step by step on top of the compiled reference objects: theta assembly (:761-776),
//     launch (:777), tangents/re-weighting (:803-804), residual (:807-820), Jacobian rows by one-hot
//     Tensor::backward (:823-873), fp64 normal equations + damping + VPoser prior (:884-904), solve
//     (:907-939) and update (:946-968).  The QP backend (QLD via QpSolverCollection) is third-party and not
//     vendored; the objective is strictly convex, so the unique minimiser is computed with an fp64 primal
//     active-set box-QP below ("parity unpinned" at that boundary).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <string>
#include <vector>

#include <torch/torch.h>

#include <smplpp/BlendShape.h>
#include <smplpp/IkTask.h>
#include <smplpp/JointRegression.h>
#include <smplpp/LinearBlendSkinning.h>
#include <smplpp/SMPL.h>
#include <smplpp/VPoser.h>
#include <smplpp/WorldTransformation.h>
#include <smplpp/definition/def.h>

namespace smplpp
{
// include/smplpp/toolbox/GeometryUtils.h:42-52 — defined (non-inline) in that header, compiled into IkTask.o.
torch::Tensor calcTriangleVertexWeights(const torch::Tensor & pos, const torch::Tensor & vertices);
} // namespace smplpp

namespace
{
thread_local std::string g_lastError;

struct SmplHandle
{
  std::shared_ptr<smplpp::SMPL> smpl;
};

struct VposerHandle
{
  smplpp::VPoserDecoder vposer{nullptr};
};

torch::Tensor fromF32(const float * p, std::vector<int64_t> shape)
{
  return torch::from_blob(const_cast<float *>(p), shape, torch::kFloat32).clone();
}

void toF32(const torch::Tensor & t, float * out)
{
  torch::Tensor c = t.detach().to(torch::kCPU).to(torch::kFloat32).contiguous();
  std::memcpy(out, c.data_ptr<float>(), sizeof(float) * c.numel());
}

// ---- fp64 dense helpers for the restated node.cpp:884-939 (Eigen is not available) ----

// In-place lower Cholesky of the n×n SPD matrix a (row-major); returns false on a non-positive pivot
// (Eigen::LLT NumericalIssue, node.cpp:934-937).
bool cholesky(std::vector<double> & a, int n)
{
  for(int j = 0; j < n; j++)
  {
    double d = a[j * n + j];
    for(int k = 0; k < j; k++) d -= a[j * n + k] * a[j * n + k];
    if(!(d > 0.0)) return false;
    d = std::sqrt(d);
    a[j * n + j] = d;
    for(int i = j + 1; i < n; i++)
    {
      double s = a[i * n + j];
      for(int k = 0; k < j; k++) s -= a[i * n + k] * a[j * n + k];
      a[i * n + j] = s / d;
    }
  }
  return true;
}

void choleskySolve(const std::vector<double> & l, int n, std::vector<double> & x)
{
  for(int i = 0; i < n; i++)
  {
    double s = x[i];
    for(int k = 0; k < i; k++) s -= l[i * n + k] * x[k];
    x[i] = s / l[i * n + i];
  }
  for(int i = n - 1; i >= 0; i--)
  {
    double s = x[i];
    for(int k = i + 1; k < n; k++) s -= l[k * n + i] * x[k];
    x[i] = s / l[i * n + i];
  }
}

// min 1/2 x'Ax + b'x  s.t. lo <= x <= hi, A SPD: primal active-set, exact in finitely many steps.
// Returns 0 on success, 1 on Cholesky failure, 2 when the iteration cap is hit.
int solveBoxQp(const std::vector<double> & A,
               const std::vector<double> & b,
               const std::vector<double> & lo,
               const std::vector<double> & hi,
               int n,
               std::vector<double> & x)
{
  x.assign(n, 0.0);
  std::vector<int> fixed(n, 0); // 0 free, -1 at lo, +1 at hi, 2 pinned (lo == hi)
  for(int i = 0; i < n; i++)
  {
    x[i] = std::min(std::max(0.0, lo[i]), hi[i]);
    if(lo[i] == hi[i]) fixed[i] = 2;
  }
  std::vector<double> g(n);
  // atMinimiser: the last step was a full, unblocked Newton step on the free variables, i.e. x minimises the
  // objective on the current face up to rounding.  Re-solving there yields a step of rounding-noise size whose
  // norm need not fall under any fixed threshold (cond(A) ~ 1e5 near convergence), so the multiplier test follows
  // directly instead.
  bool atMinimiser = false;
  for(int iter = 0; iter < 20 * n + 50; iter++)
  {
    for(int i = 0; i < n; i++)
    {
      double s = b[i];
      for(int k = 0; k < n; k++) s += A[i * n + k] * x[k];
      g[i] = s;
    }
    std::vector<int> freeIdx;
    for(int i = 0; i < n; i++)
      if(fixed[i] == 0) freeIdx.push_back(i);
    int nf = static_cast<int>(freeIdx.size());
    std::vector<double> d(n, 0.0);
    double dmax = 0.0;
    if(nf > 0 && !atMinimiser)
    {
      std::vector<double> Aff(static_cast<size_t>(nf) * nf), rhs(nf);
      for(int r = 0; r < nf; r++)
      {
        rhs[r] = -g[freeIdx[r]];
        for(int c = 0; c < nf; c++) Aff[r * nf + c] = A[freeIdx[r] * n + freeIdx[c]];
      }
      if(!cholesky(Aff, nf)) return 1;
      choleskySolve(Aff, nf, rhs);
      for(int r = 0; r < nf; r++)
      {
        d[freeIdx[r]] = rhs[r];
        dmax = std::max(dmax, std::abs(rhs[r]));
      }
    }
    double xscale = 1.0;
    for(int i = 0; i < n; i++) xscale = std::max(xscale, std::abs(x[i]));
    if(atMinimiser || dmax <= 1e-14 * xscale)
    {
      // stationary on the current face: check multipliers of the bound-active variables
      int worst = -1;
      double worstVal = 1e-12;
      for(int i = 0; i < n; i++)
      {
        double viol = 0.0;
        if(fixed[i] == -1) viol = -g[i]; // at lower bound: g < 0 wants to increase x -> release
        if(fixed[i] == 1) viol = g[i]; // at upper bound: g > 0 wants to decrease x -> release
        if(viol > worstVal)
        {
          worstVal = viol;
          worst = i;
        }
      }
      if(worst < 0) return 0;
      fixed[worst] = 0;
      atMinimiser = false;
      continue;
    }
    double alpha = 1.0;
    int block = -1, blockSide = 0;
    for(int i : freeIdx)
    {
      if(d[i] > 0.0 && std::isfinite(hi[i]))
      {
        double a = (hi[i] - x[i]) / d[i];
        if(a < alpha)
        {
          alpha = a;
          block = i;
          blockSide = 1;
        }
      }
      else if(d[i] < 0.0 && std::isfinite(lo[i]))
      {
        double a = (lo[i] - x[i]) / d[i];
        if(a < alpha)
        {
          alpha = a;
          block = i;
          blockSide = -1;
        }
      }
    }
    for(int i : freeIdx) x[i] += alpha * d[i];
    atMinimiser = block < 0;
    if(block >= 0)
    {
      x[block] = blockSide > 0 ? hi[block] : lo[block];
      fixed[block] = blockSide;
    }
  }
  return 2;
}
} // namespace

#define REF_TRY try {
#define REF_CATCH                         \
  }                                       \
  catch(const std::exception & ex)        \
  {                                       \
    g_lastError = ex.what();              \
    return -1;                            \
  }                                       \
  catch(...)                              \
  {                                       \
    g_lastError = "unknown exception";    \
    return -1;                            \
  }

extern "C"
{

const char * ref_last_error()
{
  return g_lastError.c_str();
}

void ref_set_num_threads(int n)
{
  at::set_num_threads(n);
}

int ref_get_num_threads()
{
  return at::get_num_threads();
}

void ref_set_shape(int64_t batchSize, int64_t vertexNum)
{
  smplpp::BATCH_SIZE = batchSize;
  smplpp::VERTEX_NUM = vertexNum;
}

// ---------------------------------------------------------------------------------------------------------
// smplpp::SMPL (src/SMPL.cpp)
// ---------------------------------------------------------------------------------------------------------

int ref_smpl_create(const char * jsonPath, void ** out)
{
  REF_TRY
  smplpp::BATCH_SIZE = 1;
  smplpp::VERTEX_NUM = 6890;
  auto h = std::make_unique<SmplHandle>();
  h->smpl = std::make_shared<smplpp::SMPL>();
  torch::Device device(torch::kCPU, 0);
  h->smpl->setDevice(device);
  h->smpl->setModelPath(std::string(jsonPath));
  h->smpl->init();
  *out = h.release();
  return 0;
  REF_CATCH
}

void ref_smpl_destroy(void * handle)
{
  delete static_cast<SmplHandle *>(handle);
}

// SMPL::launch on `batch` frames.  batched == 0: the stock path, BATCH_SIZE = 1, one launch per frame
// (def.h:8).  batched != 0: BATCH_SIZE = batch, ONE launch over all frames (same unmodified module sources).
// noGrad != 0 wraps the call in torch::NoGradGuard (the reference itself records the autograd graph in IK mode).
int ref_smpl_forward(void * handle,
                     int64_t batch,
                     int batched,
                     int noGrad,
                     const float * beta, // (batch, 10)
                     const float * theta, // (batch, 25, 3): row 0 translation, rows 1..24 axis-angle
                     float * vertices, // (batch, 6890, 3) or null
                     float * joints, // (batch, 24, 3) or null
                     float * restShape) // (batch, 6890, 3) or null
{
  REF_TRY
  auto & smpl = *static_cast<SmplHandle *>(handle)->smpl;
  std::unique_ptr<torch::NoGradGuard> guard;
  if(noGrad) guard = std::make_unique<torch::NoGradGuard>();
  const int64_t V = 6890;
  smplpp::VERTEX_NUM = V;
  if(batched)
  {
    smplpp::BATCH_SIZE = batch;
    torch::Tensor b = fromF32(beta, {batch, 10});
    torch::Tensor t = fromF32(theta, {batch, 25, 3});
    smpl.launch(b, t);
    if(vertices) toF32(smpl.getVertex(), vertices);
    if(joints) toF32(smpl.getRestJoint(), joints);
    if(restShape) toF32(smpl.getRestShape(), restShape);
    smplpp::BATCH_SIZE = 1;
  }
  else
  {
    smplpp::BATCH_SIZE = 1;
    for(int64_t i = 0; i < batch; i++)
    {
      torch::Tensor b = fromF32(beta + i * 10, {1, 10});
      torch::Tensor t = fromF32(theta + i * 75, {1, 25, 3});
      smpl.launch(b, t);
      if(vertices) toF32(smpl.getVertex(), vertices + i * V * 3);
      if(joints) toF32(smpl.getRestJoint(), joints + i * 72);
      if(restShape) toF32(smpl.getRestShape(), restShape + i * V * 3);
    }
  }
  return 0;
  REF_CATCH
}

// Face and vertex normals of the LAST launched mesh (SMPL::calcNormal / calcVertexNormal, SMPL.cpp:518-535).
int ref_smpl_normals(void * handle,
                     int64_t nFaces,
                     const int64_t * faceIdx,
                     float * faceNormals,
                     int64_t nVerts,
                     const int64_t * vertIdx,
                     float * vertNormals)
{
  REF_TRY
  auto & smpl = *static_cast<SmplHandle *>(handle)->smpl;
  torch::NoGradGuard guard;
  for(int64_t i = 0; i < nFaces; i++) toF32(smpl.calcNormal(faceIdx[i]), faceNormals + 3 * i);
  for(int64_t i = 0; i < nVerts; i++) toF32(smpl.calcVertexNormal(vertIdx[i]), vertNormals + 3 * i);
  return 0;
  REF_CATCH
}

// ---------------------------------------------------------------------------------------------------------
// The four pipeline modules driven directly (known-answer vectors of src/toolbox/Tester.cpp use tiny V).
// ---------------------------------------------------------------------------------------------------------

int ref_blend_shape(int64_t batch,
                    int64_t V,
                    const float * beta, // (batch, 10)
                    const float * theta, // (batch, 24, 3)
                    const float * shapeBasis, // (V, 3, 10)
                    const float * poseBasis, // (V, 3, 207)
                    float * shapeBlend, // (batch, V, 3)
                    float * poseBlend, // (batch, V, 3)
                    float * poseRot) // (batch, 24, 3, 3)
{
  REF_TRY
  torch::NoGradGuard guard;
  smplpp::BATCH_SIZE = batch;
  smplpp::VERTEX_NUM = V;
  torch::Device device(torch::kCPU, 0);
  smplpp::BlendShape m;
  m.setDevice(device);
  m.setBeta(fromF32(beta, {batch, 10}));
  m.setTheta(fromF32(theta, {batch, 24, 3}));
  m.setShapeBlendBasis(fromF32(shapeBasis, {V, 3, 10}));
  m.setPoseBlendBasis(fromF32(poseBasis, {V, 3, 207}));
  m.blend();
  toF32(m.getShapeBlendShape(), shapeBlend);
  toF32(m.getPoseBlendShape(), poseBlend);
  toF32(m.getPoseRotation(), poseRot);
  smplpp::BATCH_SIZE = 1;
  smplpp::VERTEX_NUM = 6890;
  return 0;
  REF_CATCH
}

int ref_joint_regression(int64_t batch,
                         int64_t V,
                         const float * templ, // (V, 3)
                         const float * jointRegressor, // (24, V)
                         const float * shapeBlend, // (batch, V, 3)
                         const float * poseBlend, // (batch, V, 3)
                         float * restShape, // (batch, V, 3)
                         float * joints) // (batch, 24, 3)
{
  REF_TRY
  torch::NoGradGuard guard;
  smplpp::BATCH_SIZE = batch;
  smplpp::VERTEX_NUM = V;
  torch::Device device(torch::kCPU, 0);
  smplpp::JointRegression m;
  m.setDevice(device);
  m.setShapeBlendShape(fromF32(shapeBlend, {batch, V, 3}));
  m.setPoseBlendShape(fromF32(poseBlend, {batch, V, 3}));
  m.setTemplateRestShape(fromF32(templ, {V, 3}));
  m.setJointRegressor(fromF32(jointRegressor, {24, V}));
  m.regress();
  toF32(m.getRestShape(), restShape);
  toF32(m.getJoint(), joints);
  smplpp::BATCH_SIZE = 1;
  smplpp::VERTEX_NUM = 6890;
  return 0;
  REF_CATCH
}

int ref_world_transformation(int64_t batch,
                             const int64_t * kineTree, // (2, 24)
                             const float * joints, // (batch, 24, 3)
                             const float * poseRot, // (batch, 24, 3, 3)
                             float * transforms) // (batch, 24, 4, 4)
{
  REF_TRY
  torch::NoGradGuard guard;
  smplpp::BATCH_SIZE = batch;
  torch::Device device(torch::kCPU, 0);
  smplpp::WorldTransformation m;
  m.setDevice(device);
  m.setKinematicTree(torch::from_blob(const_cast<int64_t *>(kineTree), {2, 24}, torch::kInt64).clone());
  m.setJoint(fromF32(joints, {batch, 24, 3}));
  m.setPoseRotation(fromF32(poseRot, {batch, 24, 3, 3}));
  m.transform();
  toF32(m.getTransformation(), transforms);
  smplpp::BATCH_SIZE = 1;
  return 0;
  REF_CATCH
}

int ref_linear_blend_skinning(int64_t batch,
                              int64_t V,
                              const float * weights, // (V, 24)
                              const float * restShape, // (batch, V, 3)
                              const float * transforms, // (batch, 24, 4, 4)
                              const float * rootPos, // (batch, 1, 3)
                              float * vertices) // (batch, V, 3)
{
  REF_TRY
  torch::NoGradGuard guard;
  smplpp::BATCH_SIZE = batch;
  smplpp::VERTEX_NUM = V;
  torch::Device device(torch::kCPU, 0);
  smplpp::LinearBlendSkinning m;
  m.setDevice(device);
  m.setWeight(fromF32(weights, {V, 24}));
  m.setRestShape(fromF32(restShape, {batch, V, 3}));
  m.setTransformation(fromF32(transforms, {batch, 24, 4, 4}));
  m.setRootPos(fromF32(rootPos, {batch, 1, 3}));
  m.skinning();
  toF32(m.getVertex(), vertices);
  smplpp::BATCH_SIZE = 1;
  smplpp::VERTEX_NUM = 6890;
  return 0;
  REF_CATCH
}

// ---------------------------------------------------------------------------------------------------------
// VPoser (src/VPoser.cpp) and geometry helpers
// ---------------------------------------------------------------------------------------------------------

int ref_vposer_create(const char * jsonPath, void ** out)
{
  REF_TRY
  auto h = std::make_unique<VposerHandle>();
  h->vposer = smplpp::VPoserDecoder();
  h->vposer->loadParamsFromJson(std::string(jsonPath));
  h->vposer->eval();
  *out = h.release();
  return 0;
  REF_CATCH
}

void ref_vposer_destroy(void * handle)
{
  delete static_cast<VposerHandle *>(handle);
}

// forward: latent (batch, 32) -> axis-angle (batch, 21, 3); jac (optional): (batch, 63, 32) by autograd rows.
int ref_vposer_forward(void * handle, int64_t batch, const float * latent, float * axisAngle, float * jac)
{
  REF_TRY
  auto & vposer = static_cast<VposerHandle *>(handle)->vposer;
  for(int64_t i = 0; i < batch; i++)
  {
    torch::Tensor z = fromF32(latent + 32 * i, {1, 32});
    z.set_requires_grad(jac != nullptr);
    torch::Tensor out = vposer->forward(z); // (1, 21, 3)
    toF32(out, axisAngle + 63 * i);
    if(jac)
    {
      torch::Tensor flat = out.view({63});
      for(int64_t r = 0; r < 63; r++)
      {
        if(z.grad().defined()) z.mutable_grad().zero_();
        torch::Tensor sel = torch::zeros({63});
        sel.index_put_({r}, 1);
        flat.backward(sel, true);
        toF32(z.grad().view({32}), jac + (i * 63 + r) * 32);
      }
    }
  }
  return 0;
  REF_CATCH
}

// convertRotMatToAxisAngle (VPoser.cpp:25-120) on n matrices; grad (optional) = d(sum of outputs)/dR (n, 3, 3)
// — used for the NaN-free-gradient property of tests/src/TestVPoser.cpp:45-70.
int ref_rotmat_to_axis_angle(int64_t n, const float * rotMat, float * axisAngle, float * grad)
{
  REF_TRY
  torch::Tensor r = fromF32(rotMat, {n, 3, 3});
  r.set_requires_grad(grad != nullptr);
  torch::Tensor aa = smplpp::convertRotMatToAxisAngle(r);
  toF32(aa, axisAngle);
  if(grad)
  {
    aa.sum().backward();
    toF32(r.grad(), grad);
  }
  return 0;
  REF_CATCH
}

// calcTriangleVertexWeights (GeometryUtils.h:42-52)
int ref_triangle_vertex_weights(const float * pos, const float * triangle, float * weights)
{
  REF_TRY
  torch::NoGradGuard guard;
  toF32(smplpp::calcTriangleVertexWeights(fromF32(pos, {3}), fromF32(triangle, {3, 3})), weights);
  return 0;
  REF_CATCH
}

// ---------------------------------------------------------------------------------------------------------
// One IK iteration — restatement of node/node.cpp:705-968 (see the header comment).
// Tasks are given in the order the caller wants (the node iterates its std::map alphabetically).
// ---------------------------------------------------------------------------------------------------------

struct RefIkOptions
{
  int32_t enableVposer; // node.cpp:316-322; thetaDim = 44 instead of 75
  int32_t optimizeBeta; // node.cpp:652-656
  int32_t enableQp; // node.cpp:907 (else LLT, :933-938)
  int32_t nTasks;
  int32_t updateState; // apply node.cpp:946-968 to theta / beta
  int32_t reserved;
};

// Returns 0 ok, 1 skipped (too few markers, node.cpp:785), 2 LLT/QP numerical issue, -1 exception.
int ref_ik_iteration(void * smplHandle,
                     void * vposerHandle, // may be null when !enableVposer
                     const RefIkOptions * opt,
                     float * thetaState, // (thetaDim) in/out: g_theta
                     float * betaState, // (10) in/out: g_beta
                     const int64_t * faceIdx, // (n) IkTask::faceIdx_
                     float * vertexWeights, // (n, 3) in/out: IkTask::vertexWeights_ (re-weighted at :803-804)
                     const float * targetPos, // (n, 3)
                     const float * targetNormal, // (n, 3)
                     const double * posTaskWeight, // (n)
                     const double * normalTaskWeight, // (n)
                     const double * phiLimit, // (n)
                     const double * normalOffset, // (n)
                     double * eOut, // (4n) or null
                     double * jOut, // (4n, dim) row-major or null, dim = thetaDim + 2n + (optimizeBeta ? 10 : 0)
                     double * aOut, // (dim, dim) or null
                     double * bOut, // (dim) or null
                     double * deltaOut, // (dim) or null
                     float * actualPosOut) // (n, 3) or null: calcActualPos()+tangents*dphi of :955-958
{
  REF_TRY
  auto smpl = static_cast<SmplHandle *>(smplHandle)->smpl;
  const int n = opt->nTasks;
  const bool enableVposer = opt->enableVposer != 0;
  const bool optimizeBeta = opt->optimizeBeta != 0;
  smplpp::BATCH_SIZE = 1;
  smplpp::VERTEX_NUM = 6890;

  const int thetaDim = enableVposer ? (smplpp::LATENT_DIM + 12) : 3 * (smplpp::JOINT_NUM + 1);
  const int phiDim = 2 * n;
  const int betaDim = optimizeBeta ? static_cast<int>(smplpp::SHAPE_BASIS_DIM) : 0;
  const int dim = thetaDim + phiDim + betaDim;

  torch::Tensor gTheta =
      enableVposer ? fromF32(thetaState, {thetaDim}) : fromF32(thetaState, {smplpp::JOINT_NUM + 1, 3});
  torch::Tensor gBeta = fromF32(betaState, {smplpp::SHAPE_BASIS_DIM});

  std::vector<smplpp::IkTask> tasks;
  tasks.reserve(n);
  int validNum = 0;
  for(int i = 0; i < n; i++)
  {
    tasks.emplace_back(smpl, faceIdx[i], fromF32(targetPos + 3 * i, {3}), fromF32(targetNormal + 3 * i, {3}));
    auto & t = tasks.back();
    t.posTaskWeight_ = posTaskWeight[i];
    t.normalTaskWeight_ = normalTaskWeight[i];
    t.phiLimit_ = phiLimit[i];
    t.normalOffset_ = normalOffset[i];
    t.vertexWeights_ = fromF32(vertexWeights + 3 * i, {3});
    if(posTaskWeight[i] > 0.0) validNum++;
  }

  // node.cpp:705-748
  gTheta.set_requires_grad(true);
  for(auto & t : tasks) t.phi_.set_requires_grad(t.phiLimit_ > 0.0);
  gBeta.set_requires_grad(optimizeBeta);

  // node.cpp:753-777
  torch::Tensor theta;
  if(enableVposer)
  {
    auto & vposer = static_cast<VposerHandle *>(vposerHandle)->vposer;
    using at::indexing::Slice;
    theta = torch::empty({smplpp::JOINT_NUM + 1, 3});
    theta.index_put_({0}, gTheta.index({Slice(0, 3)}));
    theta.index_put_({1}, gTheta.index({Slice(3, 6)}));
    torch::Tensor vposerOut = vposer->forward(gTheta.index({Slice(6, smplpp::LATENT_DIM + 6)}).view({1, -1})).index({0});
    theta.index_put_({Slice(2, 2 + 21)}, vposerOut);
    theta.index_put_({23}, gTheta.index({Slice(smplpp::LATENT_DIM + 6, smplpp::LATENT_DIM + 9)}));
    theta.index_put_({24}, gTheta.index({Slice(smplpp::LATENT_DIM + 9, smplpp::LATENT_DIM + 12)}));
  }
  else
  {
    theta = gTheta;
  }
  smpl->launch(gBeta.view({1, -1}), theta.view({1, theta.size(0), theta.size(1)}));

  // node.cpp:785 — in motion mode frames with fewer than half of the markers are skipped; reported to the
  // caller, which decides (the body/interactive modes never skip).
  const bool tooFew = validNum < n / 2;

  std::vector<double> e(4 * n, 0.0), J(static_cast<size_t>(4 * n) * dim, 0.0);
  auto copyRow = [&](const torch::Tensor & grad, int row, int col, int len) {
    torch::Tensor g = grad.detach().to(torch::kCPU).to(torch::kFloat32).contiguous().view({len});
    const float * p = g.data_ptr<float>();
    for(int k = 0; k < len; k++) J[static_cast<size_t>(row) * dim + col + k] = static_cast<double>(p[k]);
  };
  auto zeroGrad = [](torch::Tensor & t) {
    if(t.grad().defined()) t.mutable_grad().zero_();
  };

  int rowIdx = 0;
  for(int ti = 0; ti < n; ti++)
  {
    auto & task = tasks[ti];
    // node.cpp:803-804
    task.calcTangents();
    task.calcVertexWeights(task.calcActualPos().to(torch::kCPU).clone().detach());
    toF32(task.vertexWeights_, vertexWeights + 3 * ti);

    // node.cpp:807-820
    torch::Tensor posError = task.posTaskWeight_ * (task.calcActualPos() - task.targetPos_).to(torch::kCPU);
    {
      torch::Tensor pe = posError.detach().contiguous();
      for(int k = 0; k < 3; k++) e[rowIdx + k] = static_cast<double>(pe.data_ptr<float>()[k]);
    }
    torch::Tensor normalError;
    if(task.normalTaskWeight_ > 0.0)
    {
      normalError =
          task.normalTaskWeight_ * (at::dot(task.calcActualNormal(), task.targetNormal_).to(torch::kCPU) + 1.0);
      e[rowIdx + 3] = static_cast<double>(normalError.detach().item<float>());
    }

    // node.cpp:823-873
    auto harvest = [&](int row) {
      copyRow(gTheta.grad(), row, 0, thetaDim);
      zeroGrad(gTheta);
      if(task.phiLimit_ > 0.0)
      {
        copyRow(task.phi_.grad(), row, thetaDim + 2 * ti, 2);
        zeroGrad(task.phi_);
      }
      if(optimizeBeta)
      {
        copyRow(gBeta.grad(), row, thetaDim + phiDim, betaDim);
        zeroGrad(gBeta);
      }
    };
    for(int i = 0; i < 3; i++)
    {
      torch::Tensor select = torch::zeros({3});
      select.index_put_({i}, 1);
      posError.backward(select, true);
      harvest(rowIdx + i);
    }
    if(task.normalTaskWeight_ > 0.0)
    {
      normalError.backward({}, true);
      harvest(rowIdx + 3);
    }
    rowIdx += 4;
  }
  if(eOut) std::memcpy(eOut, e.data(), sizeof(double) * e.size());
  if(jOut) std::memcpy(jOut, J.data(), sizeof(double) * J.size());

  // node.cpp:884-904
  std::vector<double> A(static_cast<size_t>(dim) * dim, 0.0), b(dim, 0.0);
  for(int r = 0; r < 4 * n; r++)
  {
    const double * jr = &J[static_cast<size_t>(r) * dim];
    for(int i = 0; i < dim; i++)
    {
      if(jr[i] == 0.0) continue;
      b[i] += jr[i] * e[r];
      for(int k = 0; k < dim; k++) A[static_cast<size_t>(i) * dim + k] += jr[i] * jr[k];
    }
  }
  double eSq = 0.0;
  for(double v : e) eSq += v * v;
  for(int i = 0; i < dim; i++)
  {
    double reg = i < thetaDim ? 1e-3 : (i < thetaDim + phiDim ? 1e-1 : 1e-3);
    A[static_cast<size_t>(i) * dim + i] += reg + eSq;
  }
  if(enableVposer)
  {
    torch::Tensor th = gTheta.detach().contiguous();
    for(int i = 0; i < thetaDim; i++)
    {
      double w = i < 6 ? 0.0 : (i >= thetaDim - 6 ? 1e3 : 1e-5);
      A[static_cast<size_t>(i) * dim + i] += w;
      b[i] += w * static_cast<double>(th.data_ptr<float>()[i]);
    }
  }
  if(aOut) std::memcpy(aOut, A.data(), sizeof(double) * A.size());
  if(bOut) std::memcpy(bOut, b.data(), sizeof(double) * b.size());

  // node.cpp:907-939
  std::vector<double> delta(dim, 0.0);
  int status = 0;
  if(opt->enableQp)
  {
    const double inf = std::numeric_limits<double>::infinity();
    std::vector<double> lo(dim, -inf), hi(dim, inf);
    for(int ti = 0; ti < n; ti++)
    {
      for(int k = 0; k < 2; k++)
      {
        lo[thetaDim + 2 * ti + k] = -phiLimit[ti];
        hi[thetaDim + 2 * ti + k] = phiLimit[ti];
      }
    }
    for(int i = 0; i < betaDim; i++)
    {
      lo[thetaDim + phiDim + i] = -0.5;
      hi[thetaDim + phiDim + i] = 0.5;
    }
    if(solveBoxQp(A, b, lo, hi, dim, delta) != 0) status = 2;
  }
  else
  {
    std::vector<double> L = A;
    if(!cholesky(L, dim))
    {
      status = 2;
    }
    else
    {
      for(int i = 0; i < dim; i++) delta[i] = -b[i];
      choleskySolve(L, dim, delta);
    }
  }
  if(deltaOut) std::memcpy(deltaOut, delta.data(), sizeof(double) * delta.size());

  // node.cpp:955-958 (point that the node re-projects onto the mesh; evaluated on the pre-update mesh)
  if(actualPosOut)
  {
    torch::NoGradGuard guard;
    for(int ti = 0; ti < n; ti++)
    {
      auto & task = tasks[ti];
      task.phi_.set_requires_grad(false);
      torch::Tensor phi = torch::empty({2});
      phi.index_put_({0}, static_cast<float>(delta[thetaDim + 2 * ti]));
      phi.index_put_({1}, static_cast<float>(delta[thetaDim + 2 * ti + 1]));
      toF32(task.calcActualPos().to(torch::kCPU) + torch::matmul(task.tangents_, phi), actualPosOut + 3 * ti);
    }
  }

  // node.cpp:946-968
  if(opt->updateState && status == 0 && !tooFew)
  {
    for(int i = 0; i < thetaDim; i++) thetaState[i] += static_cast<float>(delta[i]);
    if(optimizeBeta)
      for(int i = 0; i < betaDim; i++) betaState[i] += static_cast<float>(delta[thetaDim + phiDim + i]);
  }
  if(status != 0) return status;
  return tooFew ? 1 : 0;
  REF_CATCH
}

} // extern "C"
