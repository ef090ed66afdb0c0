// Parity test of the header-only C++ facade (include/smplpp_b200/smplpp.hpp) against tensors produced by the COMPILED
// REFERENCE (tests/golden/ref_forward.npz, ref_ik.npz, ref_vposer.npz, dumped as raw little-endian files by
// tests/test_facade_gpu.py).  It drives the classes the way node/node.cpp and the reference's gtests drive libsmplpp:
//   smplpp::SMPL            setModelPath / init / launch / getVertex / getRestJoint / getRestShape / getVertexRaw /
//                           getFaceIndexRaw / calcNormal / calcVertexNormal / getAdjacentFaces
//   the four modules        setters -> blend / regress / transform / skinning -> getters, chained like SMPL::launch
//   smplpp::IkTask          calcActualPos / calcActualNormal / calcTangents / calcVertexWeights on public fields
//   smplpp::IkTaskSet       step + getError / getJacobian (the rows the reference takes from Tensor::backward)
//   smplpp::VPoserDecoder   loadParamsFromJson / eval / forward
// Tolerances (north_star): vertices 1e-5 m, Jacobians 1e-4 relative.   usage: facade_parity <dir>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>

#include "smplpp_b200/smplpp.hpp"

using namespace smplpp;

static std::string g_dir;
template<typename T>
static std::vector<T> load(const std::string & name)
{
  std::ifstream f(g_dir + "/" + name, std::ios::binary | std::ios::ate);
  if(!f) throw Exception("facade_parity: cannot open " + name);
  const size_t bytes = static_cast<size_t>(f.tellg());
  std::vector<T> v(bytes / sizeof(T));
  f.seekg(0);
  f.read(reinterpret_cast<char *>(v.data()), static_cast<std::streamsize>(bytes));
  return v;
}
static Array arr(const std::string & name, std::vector<int64_t> shape)
{
  Array a(std::move(shape));
  const std::vector<float> v = load<float>(name);
  if(v.size() != a.data.size()) throw Exception("facade_parity: unexpected size of " + name);
  a.data = v;
  return a;
}
static double max_abs(const float * a, const float * b, size_t n)
{
  double m = 0;
  for(size_t i = 0; i < n; i++) m = std::fmax(m, std::fabs(static_cast<double>(a[i]) - b[i]));
  return m;
}
static int g_failed = 0;
static void expect(const char * what, double got, double tol)
{
  std::printf("  %-58s %.3g (tol %.1g) %s\n", what, got, tol, got <= tol ? "ok" : "FAIL");
  if(!(got <= tol)) g_failed++;
}

int main(int argc, char ** argv)
{
  if(argc < 2) return 2;
  g_dir = argv[1];
  try
  {
    const int64_t V = VERTEX_NUM;
    auto smpl = std::make_shared<SMPL>();
    smpl->setModelPath(g_dir + "/model.npz");
    smpl->init();
    // ---- forward pass (SMPL::launch + getters) against the compiled reference ----
    const std::vector<float> beta_v = load<float>("fwd_beta.f32");
    const int64_t B = static_cast<int64_t>(beta_v.size() / 10);
    Array beta = arr("fwd_beta.f32", {B, 10}), theta = arr("fwd_theta.f32", {B, 25, 3});
    smpl->launch(beta, theta);
    const Array gv = arr("fwd_vertices.f32", {B, V, 3}), gj = arr("fwd_joints.f32", {B, 24, 3}), gr = arr("fwd_rest_shape.f32", {B, V, 3});
    std::puts("SMPL::launch");
    expect("getVertex vs reference [m]", max_abs(smpl->getVertex().ptr(), gv.ptr(), gv.data.size()), 1e-5);
    expect("getRestJoint vs reference [m]", max_abs(smpl->getRestJoint().ptr(), gj.ptr(), gj.data.size()), 1e-5);
    expect("getRestShape vs reference [m]", max_abs(smpl->getRestShape().ptr(), gr.ptr(), gr.data.size()), 1e-5);
    {
      const Array raw = smpl->getVertexRaw(std::vector<int64_t>{0, 17, V - 1});
      const float * v0 = gv.ptr();
      double m = std::fmax(max_abs(raw.ptr(), v0, 3), std::fmax(max_abs(raw.ptr() + 3, v0 + 17 * 3, 3), max_abs(raw.ptr() + 6, v0 + (V - 1) * 3, 3)));
      expect("getVertexRaw (batch element 0)", m, 1e-5);
      const std::array<int32_t, 3> f5 = smpl->getFaceIndexRaw(5);
      const std::vector<int32_t> & fi = smpl->getFaceIndex();
      expect("getFaceIndexRaw (1-based ids)", (f5[0] == fi[15] && f5[1] == fi[16] && f5[2] == fi[17] && f5[0] >= 1) ? 0.0 : 1.0, 0.0);
      const auto adj = smpl->getAdjacentFaces(f5[0] - 1);
      double wsum = 0;
      bool has5 = false;
      for(const auto & kv : adj) wsum += kv.second, has5 |= kv.first == 5;
      expect("getAdjacentFaces: weights 1/deg, contains the face", has5 ? std::fabs(wsum - 1.0) : 1.0, 1e-6);
    }
    // ---- normals on frame 0 (SMPL::calcNormal / calcVertexNormal); 1 / edge-length conditioned like the reference's ----
    {
      const std::vector<int64_t> nf = load<int64_t>("normal_face_idx.i64"), nv = load<int64_t>("normal_vert_idx.i64");
      const Array gfn = arr("face_normals.f32", {static_cast<int64_t>(nf.size()), 3}), gvn = arr("vertex_normals.f32", {static_cast<int64_t>(nv.size()), 3});
      double mf = 0, mv = 0;
      for(size_t i = 0; i < nf.size(); i += 5) mf = std::fmax(mf, max_abs(smpl->calcNormal(nf[i]).ptr(), gfn.ptr() + 3 * i, 3));
      for(size_t i = 0; i < nv.size(); i += 11) mv = std::fmax(mv, max_abs(smpl->calcVertexNormal(nv[i]).ptr(), gvn.ptr() + 3 * i, 3));
      expect("calcNormal vs reference", mf, 2e-4);
      expect("calcVertexNormal vs reference", mv, 2e-4);
    }
    // ---- the four modules chained like SMPL::launch (SMPL.cpp:684-727) ----
    {
      std::puts("BlendShape -> JointRegression -> WorldTransformation -> LinearBlendSkinning");
      Array thetaBody({B, 24, 3}), rootPos({B, 3});
      for(int64_t b = 0; b < B; b++)
      {
        std::copy(theta.ptr() + b * 75 + 3, theta.ptr() + (b + 1) * 75, thetaBody.ptr() + b * 72);
        std::copy(theta.ptr() + b * 75, theta.ptr() + b * 75 + 3, rootPos.ptr() + b * 3);
      }
      BlendShape blender;
      blender.setBeta(beta);
      blender.setTheta(thetaBody);
      blender.setShapeBlendBasis(arr("shape_blend_shapes.f32", {V, 3, 10}));
      blender.setPoseBlendBasis(arr("pose_blend_shapes.f32", {V, 3, 207}));
      blender.blend();
      JointRegression regressor;
      regressor.setTemplateRestShape(arr("vertices_template.f32", {V, 3}));
      regressor.setJointRegressor(arr("joint_regressor.f32", {24, V}));
      regressor.setShapeBlendShape(blender.getShapeBlendShape());
      regressor.setPoseBlendShape(blender.getPoseBlendShape());
      regressor.regress();
      expect("JointRegression::getRestShape [m]", max_abs(regressor.getRestShape().ptr(), gr.ptr(), gr.data.size()), 1e-5);
      expect("JointRegression::getJoint [m]", max_abs(regressor.getJoint().ptr(), gj.ptr(), gj.data.size()), 1e-5);
      WorldTransformation transformer;
      transformer.setKinematicTree(load<int64_t>("kinematic_tree.i64"));
      transformer.setJoint(regressor.getJoint());
      transformer.setPoseRotation(blender.getPoseRotation());
      transformer.transform();
      LinearBlendSkinning skinner;
      skinner.setWeight(arr("weights.f32", {V, 24}));
      skinner.setRestShape(regressor.getRestShape());
      skinner.setTransformation(transformer.getTransformation());
      skinner.setRootPos(rootPos);
      skinner.skinning();
      expect("LinearBlendSkinning::getVertex [m]", max_abs(skinner.getVertex().ptr(), gv.ptr(), gv.data.size()), 1e-5);
      try
      {
        blender.setBeta(Array({B, 9}));
        g_failed++;
      }
      catch(const Exception & e)
      {
        expect("setBeta with 9 betas throws the reference's text", std::strstr(e.what(), "BlendShape Error: Failed to set beta!") ? 0 : 1, 0);
      }
    }
    // ---- IkTask object and one IK iteration at the golden state ----
    {
      std::puts("IkTask / IkTaskSet::step (motion mode: 15 mm normal offset, fixed beta)");
      const std::vector<int64_t> faces = load<int64_t>("ik_face_idx.i64");
      const int64_t n = static_cast<int64_t>(faces.size());
      Array th = arr("ik_theta_in.f32", {1, 75}), be = arr("ik_beta_in.f32", {1, 10}), vw = arr("ik_vertex_weights_in.f32", {1, n, 3});
      Array th25 = th;
      th25.shape = {1, 25, 3};
      smpl->launch(be, th25);
      const Array actual = arr("ik_motion_actual_pos.f32", {n, 3}), vwOut = arr("ik_motion_vertex_weights_out.f32", {n, 3});
      double mp = 0, mw = 0;
      for(int64_t m = 0; m < n; m += 3)
      {
        IkTask task(smpl, faces[m]);
        task.normalOffset_ = 0.015;
        std::copy(vw.ptr() + 3 * m, vw.ptr() + 3 * m + 3, task.vertexWeights_.ptr());
        // node.cpp:803-804: tangents, then the weights of the (offset) actual position; :807 the position with the new weights
        task.calcTangents();
        task.calcVertexWeights(task.calcActualPos());
        mw = std::fmax(mw, max_abs(task.vertexWeights_.ptr(), vwOut.ptr() + 3 * m, 3));
        mp = std::fmax(mp, max_abs(task.calcActualPos().ptr(), actual.ptr() + 3 * m, 3));
        const Array nrm = task.calcActualNormal();
        const double len = std::sqrt(nrm.data[0] * nrm.data[0] + nrm.data[1] * nrm.data[1] + nrm.data[2] * nrm.data[2]);
        if(std::fabs(len - 1.0) > 1e-5) g_failed++;
      }
      expect("IkTask::calcVertexWeights vs reference", mw, 1e-4);
      expect("IkTask::calcActualPos vs reference [m]", mp, 1e-5);
      IkTaskSet set(*smpl, faces);
      smplpp_ik_options opt = IkTaskSet::defaultOptions();
      opt.skip_if_too_few = 0;
      const Array tgt = arr("ik_motion_target.f32", {1, n, 3}), pw = arr("ik_motion_pos_task_weight.f32", {1, n});
      {
        // the getter alone (smplpp_ik_jacobian): same rows, state untouched
        Array vwLin = vw;
        const Array thBefore = th;
        set.linearize(*smpl, nullptr, opt, th, be, vwLin, tgt, pw);
        const Array Jlin = set.getJacobian(), JrefLin = arr("ik_motion_J.f32", {1, 4 * n, 75 + 2 * n});
        double jm = 0;
        for(float x : JrefLin.data) jm = std::fmax(jm, std::fabs(x));
        expect("linearize: Jacobian vs reference autograd rows (relative)", max_abs(Jlin.ptr(), JrefLin.ptr(), Jlin.data.size()) / jm, 1e-4);
        expect("linearize leaves theta unchanged", max_abs(th.ptr(), thBefore.ptr(), 75), 0.0);
      }
      const std::vector<int32_t> status = set.step(*smpl, nullptr, opt, th, be, vw, tgt, pw);
      const std::vector<double> eRef = load<double>("ik_motion_e.f64");
      const Array e = set.getError();
      double me = 0;
      for(size_t i = 0; i < eRef.size(); i++) me = std::fmax(me, std::fabs(e.data[i] - eRef[i]));
      const Array J = set.getJacobian(), Jref = arr("ik_motion_J.f32", {1, 4 * n, 75 + 2 * n});
      double jmax = 0;
      for(float x : Jref.data) jmax = std::fmax(jmax, std::fabs(x));
      expect("step status", status[0], 0);
      expect("getError vs reference [m]", me, 1e-5);
      expect("getJacobian vs reference autograd rows (relative)", max_abs(J.ptr(), Jref.ptr(), J.data.size()) / jmax, 1e-4);
      expect("updated theta vs reference", max_abs(th.ptr(), arr("ik_motion_theta_out.f32", {1, 75}).ptr(), 75), 2e-4);
    }
    // ---- VPoser decoder ----
    {
      std::puts("VPoserDecoder");
      VPoserDecoder vposer;
      vposer.loadParamsFromJson(g_dir + "/vposer.json");
      vposer.eval();
      const std::vector<float> lat = load<float>("vposer_latent.f32");
      const int64_t L = static_cast<int64_t>(lat.size() / 32);
      const Array aa = vposer.forward(arr("vposer_latent.f32", {L, 32}));
      expect("forward vs reference [rad]", max_abs(aa.ptr(), arr("vposer_axis_angle.f32", {L, 21, 3}).ptr(), aa.data.size()), 2e-5);
    }
  }
  catch(const Exception & e)
  {
    std::printf("facade parity: %s\n", e.what());
    return std::strstr(e.what(), "CUDA") ? 3 : 1;
  }
  std::printf("facade parity: %s\n", g_failed ? "FAILED" : "OK");
  return g_failed ? 1 : 0;
}
