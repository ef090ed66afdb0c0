// Smoke test of the header-only C++ facade (include/smplpp_b200/smplpp.hpp): drives smplpp::SMPL the way
// node/node.cpp:412-415, 777, 1114-1117 does (init, launch, getVertex, getRestJoint) on a small procedural model and
// checks finiteness, translation equivariance and the reference's exception texts.  Needs a CUDA device at run time;
// on a machine without one it must fail with the library's "no CPU fallback" error (exit code 3).
#include <cmath>
#include <cstdio>
#include <cstring>

#include "smplpp_b200/smplpp.hpp"

static uint32_t g_seed = 12345u;
static float frand()
{
  g_seed = g_seed * 1664525u + 1013904223u;
  return static_cast<float>((g_seed >> 8) & 0xFFFFFF) / 16777216.f - 0.5f;
}

int main(int argc, char ** argv)
{
  using namespace smplpp;
  const int64_t V = 300; // even, >= 128: exercises the tcgen05 path
  ModelParams p;
  p.vertex_num = V;
  p.shape_blend_shapes.resize(V * 3 * SHAPE_BASIS_DIM);
  p.pose_blend_shapes.resize(V * 3 * POSE_BASIS_DIM);
  p.vertices_template.resize(V * 3);
  p.joint_regressor.assign(JOINT_NUM * V, 0.f);
  p.weights.assign(V * JOINT_NUM, 0.f);
  for(auto & x : p.shape_blend_shapes) x = 0.02f * frand();
  for(auto & x : p.pose_blend_shapes) x = 0.004f * frand();
  for(auto & x : p.vertices_template) x = frand();
  for(int64_t j = 0; j < JOINT_NUM; j++)
    for(int64_t k = 0; k < 8; k++) p.joint_regressor[j * V + (j * 11 + k * 7) % V] = 0.125f;
  for(int64_t v = 0; v < V; v++)
  {
    p.weights[v * JOINT_NUM + v % JOINT_NUM] = 0.75f;
    p.weights[v * JOINT_NUM + (v / 3 + 5) % JOINT_NUM] += 0.25f;
  }
  const int64_t parents[24] = {-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21};
  p.kinematic_tree.resize(48);
  for(int j = 0; j < 24; j++)
  {
    p.kinematic_tree[j] = parents[j] < 0 ? 4294967295ll : parents[j];
    p.kinematic_tree[24 + j] = j;
  }
  for(int f = 0; f < 100; f++)
    for(int k = 0; k < 3; k++) p.face_indices.push_back(1 + (f * 3 + k * 17) % V);

  try
  {
    SMPL smpl;
    try
    {
      smpl.launch(Array({1, 10}), Array({1, 25, 3}));
      std::puts("FAIL: launch before init did not throw");
      return 1;
    }
    catch(const Exception & e)
    {
      if(!std::strstr(e.what(), "Cannot launch a SMPL model!")) return 1;
    }
    try
    {
      smpl.setModelPath("/nonexistent/smpl_male.json");
      std::puts("FAIL: setModelPath on a missing file did not throw");
      return 1;
    }
    catch(const Exception & e)
    {
      if(!std::strstr(e.what(), "SMPL Error: Failed to initialize model path!")) return 1;
    }
    try
    {
      VPoserDecoder missing;
      missing.loadParamsFromJson("/nonexistent/vposer.json");
      return 1;
    }
    catch(const Exception & e)
    {
      if(!std::strstr(e.what(), "VPoser Error: Cannot find a JSON file!")) return 1;
    }
    if(argc > 3)
    {
      // C3D reader (node.cpp:572-595): host only
      C3d c3d(argv[3]);
      std::vector<float> xyz;
      std::vector<uint8_t> valid;
      c3d.read(0, c3d.nbFrames(), xyz, valid);
      std::printf("facade smoke: c3d %lld frames x %lld points at %.0f Hz, first label %s\n", (long long)c3d.nbFrames(),
                  (long long)c3d.nbPoints(), c3d.frameRate(), c3d.label(0).c_str());
      if(c3d.findLabel("no such marker") != c3d.nbPoints()) return 1;
    }
    smpl.init(p);
    if(argc > 2)
    {
      // setModelPath + init() and VPoserDecoder::loadParamsFromJson through the library's JSON reader
      SMPL fromJson;
      fromJson.setModelPath(argv[1]);
      fromJson.init();
      VPoserDecoder vposer;
      vposer.loadParamsFromJson(argv[2]);
      Array b1({1, 10}), t1({1, 25, 3});
      fromJson.launch(b1, t1);
      std::printf("facade smoke: model from %s: V=%lld, %zu faces\n", argv[1], (long long)fromJson.getVertexNum(),
                  fromJson.getFaceIndex().size() / 3);
      if(fromJson.getVertex().size(1) != fromJson.getVertexNum() || !vposer.handle()) return 1;
    }
    const int64_t N = 37;
    Array beta({N, 10}), theta({N, 25, 3});
    for(auto & x : beta.data) x = 2.f * frand();
    for(auto & x : theta.data) x = 0.6f * frand();
    smpl.launch(beta, theta);
    Array v0 = smpl.getVertex();
    if(v0.size(0) != N || v0.size(1) != V || smpl.getRestJoint().size(1) != JOINT_NUM) return 1;
    for(float x : v0.data)
      if(!std::isfinite(x)) return 1;
    // moving the root translation (theta row 0) moves every vertex by the same vector (SMPL.cpp:726-727)
    Array theta2 = theta;
    for(int64_t n = 0; n < N; n++) theta2.data[n * 75 + 1] += 0.5f;
    smpl.launch(beta, theta2);
    const Array & v1 = smpl.getVertex();
    double worst = 0;
    for(size_t i = 0; i < v0.data.size(); i++)
      worst = std::fmax(worst, std::fabs(double(v1.data[i]) - v0.data[i] - (i % 3 == 1 ? 0.5 : 0.0)));
    std::printf("facade smoke: N=%lld V=%lld, translation equivariance max err %.3g\n", (long long)N, (long long)V, worst);
    if(worst > 2e-6) return 1;
    try
    {
      smpl.launch(Array({N, 9}), theta);
      return 1;
    }
    catch(const Exception & e)
    {
      if(!std::strstr(e.what(), "BlendShape Error: Failed to set beta!")) return 1;
    }
    std::puts("facade smoke: OK");
  }
  catch(const Exception & e)
  {
    std::printf("facade smoke: %s\n", e.what());
    return std::strstr(e.what(), "CUDA") ? 3 : 1;
  }
  return 0;
}
