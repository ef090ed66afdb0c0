// smplpp_b200 — header-only C++ facade over the C ABI (include/smplpp_b200.h) that keeps the class and method names
// of the reference's libsmplpp (include/smplpp/SMPL.h:210-270, IkTask.h:13-85, VPoser.h:33-90,
// toolbox/Exception.h:49) so that a caller such as node/node.cpp keeps its shape once libtorch is removed from the
// path.  torch::Tensor is replaced by plain row-major host arrays (smplpp::Array) for the object-level API; callers
// that keep everything on the device use the C entry points directly.  Errors are rethrown as smplpp::Exception with
// the reference's message text ("<module> Error: <msg>", src/toolbox/Exception.cpp:77-91).
// There is NO CPU fallback: every call fails with "CUDA Error: ..." without a device.
#pragma once

#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../smplpp_b200.h"

namespace smplpp
{
constexpr int64_t VERTEX_NUM = SMPLPP_VERTEX_NUM;   // definition/def.h:8-14
constexpr int64_t JOINT_NUM = SMPLPP_JOINT_NUM;
constexpr int64_t SHAPE_BASIS_DIM = SMPLPP_SHAPE_DIM;
constexpr int64_t POSE_BASIS_DIM = SMPLPP_POSE_DIM;
constexpr int64_t FACE_INDEX_NUM = SMPLPP_FACE_NUM;
constexpr int64_t LATENT_DIM = SMPLPP_LATENT_DIM;

/// Counterpart of smplpp::Exception (toolbox/Exception.h:122): what() = "<module> Error: <text>".
class Exception : public std::runtime_error
{
public:
  explicit Exception(const std::string & what) : std::runtime_error(what) {}
};

inline void check(int rc)
{
  if(rc != SMPLPP_OK)
  {
    const char * msg = smplpp_last_error();
    throw Exception(msg && *msg ? msg : "SMPL Error: smplpp_b200 call failed");
  }
}

/// Dense row-major float array with a shape (the stand-in for the torch::Tensor values the reference returns).
struct Array
{
  std::vector<int64_t> shape;
  std::vector<float> data;
  Array() = default;
  explicit Array(std::vector<int64_t> s) : shape(std::move(s))
  {
    size_t n = 1;
    for(int64_t d : shape) n *= static_cast<size_t>(d);
    data.assign(n, 0.f);
  }
  int64_t size(size_t dim) const { return shape.at(dim); }
  float * ptr() { return data.data(); }
  const float * ptr() const { return data.data(); }
};

/// The tensors of the model file (keys of src/SMPL.cpp:573-611).
struct ModelParams
{
  int64_t vertex_num = VERTEX_NUM;
  std::vector<int32_t> face_indices;      // (F,3) 1-based
  std::vector<float> shape_blend_shapes;  // (V,3,10)
  std::vector<float> pose_blend_shapes;   // (V,3,207)
  std::vector<float> vertices_template;   // (V,3)
  std::vector<float> joint_regressor;     // (24,V)
  std::vector<int64_t> kinematic_tree;    // (2,24)
  std::vector<float> weights;             // (V,24)
};

/// smplpp::SMPL (src/SMPL.cpp): init / launch / getters with host arrays.
class SMPL
{
public:
  SMPL() = default;
  explicit SMPL(const ModelParams & params) { init(params); }
  SMPL(const SMPL &) = delete;
  SMPL & operator=(const SMPL &) = delete;
  ~SMPL()
  {
    if(m__model) smplpp_model_destroy(m__model);
  }

  /// SMPL::setModelPath (SMPL.cpp:325-339) + SMPL::init() (SMPL.cpp:560-643): the model JSON is read by the library
  /// (smplpp_model_load_json); the messages are the reference's.
  void setModelPath(const std::string & modelPath)
  {
    if(FILE * f = fopen(modelPath.c_str(), "rb"))
      fclose(f);
    else
      throw Exception("SMPL Error: Failed to initialize model path!"); // SMPL.cpp:337
    m__modelPath = modelPath;
  }
  void init()
  {
    if(m__model) smplpp_model_destroy(m__model);
    m__model = nullptr;
    check(smplpp_model_load_json(m__modelPath.c_str(), &m__model));
    m__vertexNum = smplpp_model_vertex_num(m__model);
    smplpp_json_t * j = nullptr;
    check(smplpp_json_open(m__modelPath.c_str(), &j));
    int32_t ndim = 0;
    int64_t shape[8];
    const double * data = nullptr;
    const int rc = smplpp_json_array(j, "face_indices", &ndim, shape, &data);
    if(rc == SMPLPP_OK && ndim == 2)
    {
      m__faceIndices.resize(static_cast<size_t>(shape[0] * shape[1]));
      for(size_t i = 0; i < m__faceIndices.size(); i++) m__faceIndices[i] = static_cast<int32_t>(data[i]);
    }
    smplpp_json_close(j);
    check(rc);
  }

  /// SMPL::init from already-parsed arrays
  void init(const ModelParams & p)
  {
    const size_t V = static_cast<size_t>(p.vertex_num);
    if(p.shape_blend_shapes.size() != V * 3 * SHAPE_BASIS_DIM)
      throw Exception("SMPL Error: Shape parameter dimensions are invalid!"); // SMPL.cpp:581
    if(p.pose_blend_shapes.size() != V * 3 * POSE_BASIS_DIM)
      throw Exception("SMPL Error: Pose parameter dimensions are invalid!"); // SMPL.cpp:588
    if(p.vertices_template.size() != V * 3 || p.joint_regressor.size() != JOINT_NUM * V
       || p.kinematic_tree.size() != 2 * JOINT_NUM || p.weights.size() != V * JOINT_NUM || p.face_indices.size() % 3)
      throw Exception("SMPL Error: Cannot initialize a SMPL model!"); // SMPL.cpp:616
    smplpp_model_desc d;
    d.vertex_num = p.vertex_num;
    d.face_num = static_cast<int64_t>(p.face_indices.size() / 3);
    d.face_indices = p.face_indices.data();
    d.shape_blend_shapes = p.shape_blend_shapes.data();
    d.pose_blend_shapes = p.pose_blend_shapes.data();
    d.vertices_template = p.vertices_template.data();
    d.joint_regressor = p.joint_regressor.data();
    d.kinematic_tree = p.kinematic_tree.data();
    d.weights = p.weights.data();
    if(m__model) smplpp_model_destroy(m__model);
    m__model = nullptr;
    check(smplpp_model_create(&d, &m__model));
    m__faceIndices = p.face_indices;
    m__vertexNum = p.vertex_num;
  }

  /// SMPL::launch (SMPL.cpp:671-737): beta (N,10) or (1,10) shared, theta (N,25,3) with row 0 = root translation.
  void launch(const Array & beta, const Array & theta)
  {
    if(!m__model || theta.shape.size() != 3 || theta.shape[1] != JOINT_NUM + 1 || theta.shape[2] != 3)
      throw Exception("SMPL Error: Cannot launch a SMPL model!"); // SMPL.cpp:676
    const int64_t n = theta.shape[0];
    if(beta.shape.size() != 2 || beta.shape[1] != SHAPE_BASIS_DIM || (beta.shape[0] != n && beta.shape[0] != 1))
      throw Exception("BlendShape Error: Failed to set beta!"); // BlendShape.cpp:340
    m__vertices = Array({n, m__vertexNum, 3});
    m__joints = Array({n, JOINT_NUM, 3});
    const int64_t stride = (beta.shape[0] == 1 && n > 1) ? 0 : SHAPE_BASIS_DIM;
    check(smplpp_forward_host(m__model, n, beta.ptr(), stride, theta.ptr(), m__vertices.ptr(), m__joints.ptr()));
    m__launched = true;
  }

  /// SMPL::getVertex (SMPL.cpp:446-461): (N,V,3)
  const Array & getVertex() const
  {
    if(!m__launched) throw Exception("LinearBlendSknning Error: Failed to get vertices of new pose!"); // LinearBlendSkinning.cpp:409
    return m__vertices;
  }
  /// SMPL::getRestJoint (SMPL.cpp:425-440): (N,24,3)
  const Array & getRestJoint() const
  {
    if(!m__launched) throw Exception("JointRegression Error: Failed to get joints!");
    return m__joints;
  }
  /// SMPL::setVertPath + SMPL::out (SMPL.cpp:341-355, 757-790): mesh `index` of the batch as Wavefront OBJ
  void setVertPath(const std::string & vertexPath) { m__vertPath = vertexPath; }
  void out(int64_t index) const
  {
    if(!m__launched || index < 0 || index >= m__vertices.size(0) || m__vertPath.empty())
      throw Exception("SMPL Error: Cannot export the deformed mesh!"); // SMPL.cpp:785
    check(smplpp_write_obj(m__vertPath.c_str(), m__vertexNum, m__vertices.ptr() + index * m__vertexNum * 3,
                           static_cast<int64_t>(m__faceIndices.size() / 3), m__faceIndices.data()));
  }
  /// SMPL::getFaceIndex (SMPL.cpp:386-405): (F,3), 1-based like the stored tensor
  const std::vector<int32_t> & getFaceIndex() const { return m__faceIndices; }
  int64_t getVertexNum() const { return m__vertexNum; }
  /// the C handle, for device-pointer calls (smplpp_forward, smplpp_ik_step, ...)
  smplpp_model_t * handle() const { return m__model; }

private:
  smplpp_model_t * m__model = nullptr;
  std::string m__modelPath, m__vertPath;
  std::vector<int32_t> m__faceIndices;
  int64_t m__vertexNum = 0;
  Array m__vertices, m__joints;
  bool m__launched = false;
};

/// smplpp::VPoserDecoder (src/VPoser.cpp:143-238): the six decoder_net tensors, row-major (out, in).
class VPoserDecoder
{
public:
  VPoserDecoder(const std::vector<float> & w0, const std::vector<float> & b0, const std::vector<float> & w3,
                const std::vector<float> & b3, const std::vector<float> & w5, const std::vector<float> & b5)
  {
    if(w0.size() != 512 * 32 || b0.size() != 512 || w3.size() != 512 * 512 || b3.size() != 512 || w5.size() != 126 * 512
       || b5.size() != 126)
      throw Exception("VPoser Error: invalid dimension of decoder parameters!"); // VPoser.cpp:190
    smplpp_vposer_desc d{w0.data(), b0.data(), w3.data(), b3.data(), w5.data(), b5.data()};
    check(smplpp_vposer_create(&d, &vposer_));
  }
  /// VPoserDecoderImpl::loadParamsFromJson (VPoser.cpp:169-238)
  VPoserDecoder() = default;
  void loadParamsFromJson(const std::string & jsonPath)
  {
    if(vposer_) smplpp_vposer_destroy(vposer_);
    vposer_ = nullptr;
    check(smplpp_vposer_load_json(jsonPath.c_str(), &vposer_));
  }
  VPoserDecoder(const VPoserDecoder &) = delete;
  VPoserDecoder & operator=(const VPoserDecoder &) = delete;
  ~VPoserDecoder()
  {
    if(vposer_) smplpp_vposer_destroy(vposer_);
  }
  smplpp_vposer_t * handle() const { return vposer_; }

private:
  smplpp_vposer_t * vposer_ = nullptr;
};

/// The C3D file of the mocap modes as node/node.cpp:572-595, 667-691 uses it (there: ezc3d::c3d).
class C3d
{
public:
  explicit C3d(const std::string & path) { check(smplpp_c3d_open(path.c_str(), &c3d_)); }
  C3d(const C3d &) = delete;
  C3d & operator=(const C3d &) = delete;
  ~C3d()
  {
    if(c3d_) smplpp_c3d_close(c3d_);
  }
  int64_t nbFrames() const { return smplpp_c3d_frame_count(c3d_); }
  int64_t nbPoints() const { return smplpp_c3d_point_count(c3d_); }
  double frameRate() const { return smplpp_c3d_frame_rate(c3d_); }
  std::string label(int64_t i) const { return smplpp_c3d_label(c3d_, i); }
  /// first label ending with the task name (node.cpp:580-594); nbPoints() when there is none
  int64_t findLabel(const std::string & taskName) const { return smplpp_c3d_find_label(c3d_, taskName.c_str()); }
  /// frames [first, first + count): xyz (count, points, 3), valid (count, points); a missing point is zero / 0
  void read(int64_t first, int64_t count, std::vector<float> & xyz, std::vector<uint8_t> & valid) const
  {
    xyz.resize(static_cast<size_t>(count * nbPoints() * 3));
    valid.resize(static_cast<size_t>(count * nbPoints()));
    check(smplpp_c3d_read(c3d_, first, count, xyz.data(), valid.data()));
  }

private:
  smplpp_c3d_t * c3d_ = nullptr;
};

/// n smplpp::IkTask objects (IkTask.h:13-85) handled as one batch: the attachment faces are fixed here, the per-frame
/// fields (targetPos_, posTaskWeight_, vertexWeights_) are device arrays passed to smplpp_ik_step.
class IkTaskSet
{
public:
  IkTaskSet(const SMPL & smpl, const std::vector<int64_t> & faceIdx)
  {
    check(smplpp_tasks_create(smpl.handle(), static_cast<int32_t>(faceIdx.size()), faceIdx.data(), &tasks_));
  }
  IkTaskSet(const IkTaskSet &) = delete;
  IkTaskSet & operator=(const IkTaskSet &) = delete;
  ~IkTaskSet()
  {
    if(tasks_) smplpp_tasks_destroy(tasks_);
  }
  int32_t size() const { return smplpp_tasks_count(tasks_); }
  smplpp_tasks_t * handle() const { return tasks_; }
  /// IkTask defaults (IkTask.h:59-84) + the constants of node/node.cpp:884-929
  static smplpp_ik_options defaultOptions()
  {
    smplpp_ik_options o;
    smplpp_ik_options_default(&o);
    return o;
  }

private:
  smplpp_tasks_t * tasks_ = nullptr;
};
} // namespace smplpp
