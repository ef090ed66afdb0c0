// Host-buffer entry points of the IK path: what a caller of the reference's mocap modes does per frame
// (node/node.cpp:645-1002: targets from the C3D frame -> K iterations of the IK step -> theta) for a whole batch of
// frames at once, with HOST arrays in and out.  bench.py's `ik.e2e` times smplpp_ik_solve_host.
#include <dlfcn.h>

#include <algorithm>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "common.cuh"
#include "tasks.cuh"
#include "vposer.cuh"

using namespace sb;

namespace
{
// mean over the valid markers of |e_m| (the first three rows of every task), one thread per frame
__global__ void marker_residual_kernel(int B, int n, const float * __restrict__ e, const float * __restrict__ posw,
                                       float * __restrict__ out)
{
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if(f >= B) return;
  float acc = 0.f;
  int cnt = 0;
  for(int m = 0; m < n; m++)
  {
    if(posw && !(posw[static_cast<size_t>(f) * n + m] > 0.f)) continue;
    const float * em = e + (static_cast<size_t>(f) * n + m) * 4;
    acc += sqrtf(em[0] * em[0] + em[1] * em[1] + em[2] * em[2]);
    cnt++;
  }
  out[f] = cnt ? acc / static_cast<float>(cnt) : 0.f;
}

int grow(void ** p, size_t * have, size_t need)
{
  if(*have >= need) return SMPLPP_OK;
  if(*p) cudaFree(*p);
  *p = nullptr, *have = 0;
  SB_CUDA(cudaMalloc(p, need));
  *have = need;
  return SMPLPP_OK;
}
} // namespace

void sb_release_host_solve(smplpp_tasks * t)
{
  smplpp_tasks::HostSolve & hs = t->host;
  if(hs.buf) cudaFree(hs.buf);
  if(hs.ws) cudaFree(hs.ws);
  if(hs.stream) cudaStreamDestroy(hs.stream);
  hs = smplpp_tasks::HostSolve();
}

extern "C" int smplpp_ik_solve_host(const smplpp_model_t * model, const smplpp_vposer_t * vposer, smplpp_tasks_t * tasks,
                                    const smplpp_ik_options * opt, int64_t batch, int32_t iterations,
                                    float * theta_state_host, float * beta_host, int64_t beta_stride,
                                    float * vertex_weights_host, const float * target_pos_host,
                                    const float * pos_task_weight_host, int32_t * status_host, float * residual_host)
{
  if(!model || !tasks || !opt || batch < 1 || iterations < 1 || !theta_state_host || !beta_host || !vertex_weights_host
     || !target_pos_host || !status_host)
    return fail(SMPLPP_ERR_INVALID, "IkTask", "invalid IK solve arguments!");
  if(beta_stride != 0 && beta_stride < kShapeDim) return fail(SMPLPP_ERR_INVALID, "BlendShape", "Failed to set beta!");
  const int n = tasks->d.n;
  const int theta_dim = smplpp_ik_theta_dim(opt);
  const size_t B = static_cast<size_t>(batch);
  smplpp_tasks::HostSolve & hs = tasks->host;
  if(!hs.stream) SB_CUDA(cudaStreamCreateWithFlags(&hs.stream, cudaStreamNonBlocking));
  // one device block: theta | beta | vertex weights | targets | marker weights | status | e | residual
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += align_up(bytes);
    return o;
  };
  const size_t n_beta = beta_stride == 0 ? kShapeDim : (B - 1) * beta_stride + kShapeDim;
  const size_t o_theta = take(B * theta_dim * sizeof(float)), o_beta = take(n_beta * sizeof(float)),
               o_vw = take(B * n * 3 * sizeof(float)), o_tgt = take(B * n * 3 * sizeof(float)),
               o_pw = take(B * n * sizeof(float)), o_status = take(B * sizeof(int32_t)),
               o_e = take(B * 4 * n * sizeof(float)), o_res = take(B * sizeof(float));
  int rc = grow(&hs.buf, &hs.buf_bytes, off);
  if(rc != SMPLPP_OK) return rc;
  const size_t ws_need = smplpp_ik_workspace_bytes(tasks, opt, batch);
  rc = grow(&hs.ws, &hs.ws_bytes, ws_need + 512);
  if(rc != SMPLPP_OK) return rc;
  char * d = static_cast<char *>(hs.buf);
  float * d_theta = reinterpret_cast<float *>(d + o_theta);
  float * d_beta = reinterpret_cast<float *>(d + o_beta);
  float * d_vw = reinterpret_cast<float *>(d + o_vw);
  float * d_tgt = reinterpret_cast<float *>(d + o_tgt);
  float * d_pw = pos_task_weight_host ? reinterpret_cast<float *>(d + o_pw) : nullptr;
  int32_t * d_status = reinterpret_cast<int32_t *>(d + o_status);
  float * d_e = residual_host ? reinterpret_cast<float *>(d + o_e) : nullptr;
  float * d_res = reinterpret_cast<float *>(d + o_res);
  cudaStream_t st = hs.stream;
  SB_CUDA(cudaMemcpyAsync(d_theta, theta_state_host, B * theta_dim * sizeof(float), cudaMemcpyHostToDevice, st));
  SB_CUDA(cudaMemcpyAsync(d_beta, beta_host, n_beta * sizeof(float), cudaMemcpyHostToDevice, st));
  SB_CUDA(cudaMemcpyAsync(d_vw, vertex_weights_host, B * n * 3 * sizeof(float), cudaMemcpyHostToDevice, st));
  SB_CUDA(cudaMemcpyAsync(d_tgt, target_pos_host, B * n * 3 * sizeof(float), cudaMemcpyHostToDevice, st));
  if(d_pw) SB_CUDA(cudaMemcpyAsync(d_pw, pos_task_weight_host, B * n * sizeof(float), cudaMemcpyHostToDevice, st));
  for(int32_t k = 0; k < iterations; k++)
  {
    const bool last = k + 1 == iterations;
    rc = smplpp_ik_step(model, vposer, tasks, opt, st, batch, d_theta, d_beta, beta_stride, d_vw, d_tgt, nullptr, d_pw,
                        d_status, last ? d_e : nullptr, nullptr, nullptr, nullptr, nullptr, hs.ws, hs.ws_bytes);
    if(rc != SMPLPP_OK)
    {
      cudaStreamSynchronize(st);
      return rc;
    }
  }
  if(residual_host)
  {
    marker_residual_kernel<<<static_cast<unsigned>((batch + 127) / 128), 128, 0, st>>>(static_cast<int>(batch), n, d_e, d_pw, d_res);
    SB_LAUNCHED();
    SB_CUDA(cudaMemcpyAsync(residual_host, d_res, B * sizeof(float), cudaMemcpyDeviceToHost, st));
  }
  SB_CUDA(cudaMemcpyAsync(theta_state_host, d_theta, B * theta_dim * sizeof(float), cudaMemcpyDeviceToHost, st));
  if(opt->optimize_beta)
    SB_CUDA(cudaMemcpyAsync(beta_host, d_beta, n_beta * sizeof(float), cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaMemcpyAsync(vertex_weights_host, d_vw, B * n * 3 * sizeof(float), cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaMemcpyAsync(status_host, d_status, B * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaStreamSynchronize(st));
  return SMPLPP_OK;
}

// ------------------------------------------------------------------------------------------------------------
// device-memory helpers for callers that link only the C ABI (the header-only C++ facade has no CUDA headers)
// ------------------------------------------------------------------------------------------------------------
extern "C" int smplpp_device_alloc(void ** out, size_t bytes)
{
  if(!out || bytes == 0) return fail(SMPLPP_ERR_INVALID, "SMPL", "smplpp_device_alloc: bad arguments");
  *out = nullptr;
  SB_CUDA(cudaMalloc(out, bytes));
  return SMPLPP_OK;
}

extern "C" void smplpp_device_free(void * ptr)
{
  if(ptr) cudaFree(ptr);
}

extern "C" int smplpp_copy_to_device(void * dst_dev, const void * src_host, size_t bytes, void * stream)
{
  if(!dst_dev || !src_host) return fail(SMPLPP_ERR_INVALID, "SMPL", "smplpp_copy_to_device: bad arguments");
  SB_CUDA(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, as_stream(stream)));
  return SMPLPP_OK;
}

extern "C" int smplpp_copy_to_host(void * dst_host, const void * src_dev, size_t bytes, void * stream)
{
  if(!dst_host || !src_dev) return fail(SMPLPP_ERR_INVALID, "SMPL", "smplpp_copy_to_host: bad arguments");
  SB_CUDA(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, as_stream(stream)));
  SB_CUDA(cudaStreamSynchronize(as_stream(stream)));
  return SMPLPP_OK;
}

extern "C" int smplpp_stream_synchronize(void * stream)
{
  SB_CUDA(cudaStreamSynchronize(as_stream(stream)));
  return SMPLPP_OK;
}

// ------------------------------------------------------------------------------------------------------------
// shared-beta stage as ONE call with a communicator: reduce -> ncclAllReduce(111 doubles) -> apply.
// NCCL is resolved at run time (the symbol of the process when a framework already loaded it, else libnccl.so.2):
// the library itself links nothing but the CUDA runtime.
// ------------------------------------------------------------------------------------------------------------
namespace
{
using nccl_all_reduce_fn = int (*)(const void *, void *, size_t, int /*ncclDataType_t*/, int /*ncclRedOp_t*/, void * /*comm*/,
                                   cudaStream_t);
nccl_all_reduce_fn resolve_nccl_all_reduce()
{
  static nccl_all_reduce_fn fn = nullptr;
  static bool tried = false;
  if(tried) return fn;
  tried = true;
  void * sym = dlsym(RTLD_DEFAULT, "ncclAllReduce");
  if(!sym)
  {
    const char * env = getenv("SMPLPP_NCCL_LIB");
    for(const char * name : {env ? env : "libnccl.so.2", "libnccl.so.2", "libnccl.so"})
    {
      if(void * h = dlopen(name, RTLD_NOW | RTLD_GLOBAL))
      {
        sym = dlsym(h, "ncclAllReduce");
        if(sym) break;
      }
    }
  }
  fn = reinterpret_cast<nccl_all_reduce_fn>(sym);
  return fn;
}
} // namespace

extern "C" int smplpp_ik_shared_beta_step(const smplpp_model_t * model, const smplpp_vposer_t * vposer,
                                          const smplpp_tasks_t * tasks, const smplpp_ik_options * opt, void * stream,
                                          int64_t batch, float * theta_state, float * shared_beta, float * vertex_weights,
                                          const float * target_pos, const float * pos_task_weight, int32_t * status,
                                          double * reduced, void * nccl_comm, void * workspace, size_t workspace_bytes)
{
  if(!reduced) return fail(SMPLPP_ERR_INVALID, "IkTask", "invalid shared-beta arguments!");
  int rc = smplpp_ik_shared_beta_reduce(model, vposer, tasks, opt, stream, batch, theta_state, shared_beta, vertex_weights,
                                        target_pos, pos_task_weight, status, reduced, workspace, workspace_bytes);
  if(rc != SMPLPP_OK) return rc;
  if(nccl_comm)
  {
    nccl_all_reduce_fn all_reduce = resolve_nccl_all_reduce();
    if(!all_reduce) return fail(SMPLPP_ERR_INVALID, "IkTask", "ncclAllReduce is not available in this process (libnccl.so.2)");
    // ncclFloat64 = 8, ncclSum = 0 (nccl.h); in place, on the caller's stream: no host round trip
    const int nrc = all_reduce(reduced, reduced, 111, 8, 0, nccl_comm, as_stream(stream));
    if(nrc != 0) return fail(SMPLPP_ERR_CUDA, "NCCL", "ncclAllReduce failed with code " + std::to_string(nrc));
  }
  return smplpp_ik_shared_beta_apply(tasks, opt, stream, batch, theta_state, shared_beta, status, reduced, workspace,
                                     workspace_bytes);
}

// ------------------------------------------------------------------------------------------------------------
// The motion stage of the mocap mode as one call (node/node.cpp: attachments and beta from MocapBody.yaml :509-535,
// marker <-> task matching by label suffix :571-595, per-frame targets with missing markers -> weight 0 :667-691, the
// "fewer than half of the markers" skip :785, the loop :1369-1407, theta of every frame -> motion text,
// scripts/convertRosbagToText.py:13-19).
// The reference walks the frames serially, ONE iteration per frame, warm-started from the previous frame after 31 warm-up
// iterations on the first one.  Here all frames are solved at once: `warmup_iterations` on the first frame give the
// common start, then every frame gets `iterations` steps (with `reproject` the full loop body: projection onto the
// pre-update mesh and re-seated attachments per frame).
// ------------------------------------------------------------------------------------------------------------
namespace
{
struct DevBuf
{
  void * p = nullptr;
  ~DevBuf()
  {
    if(p) cudaFree(p);
  }
  int alloc(size_t bytes)
  {
    SB_CUDA(cudaMalloc(&p, std::max<size_t>(bytes, 16)));
    return SMPLPP_OK;
  }
  template<typename T>
  T * as() const
  {
    return static_cast<T *>(p);
  }
};

__global__ void broadcast_rows_kernel(long long rows, int cols, const float * __restrict__ src, float * __restrict__ dst)
{
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if(i < rows * cols) dst[i] = src[i % cols];
}
__global__ void broadcast_rows_i32_kernel(long long rows, int cols, const int32_t * __restrict__ src, int32_t * __restrict__ dst)
{
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if(i < rows * cols) dst[i] = src[i % cols];
}
// (B, 44) VPoser state -> theta rows 0, 1, 23, 24 of (B, 75); rows 2..22 come from the decoder
__global__ void state_to_theta_kernel(int B, const float * __restrict__ state, float * __restrict__ theta)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= B * 12) return;
  const int b = i / 12, k = i % 12;
  theta[static_cast<size_t>(b) * 75 + (k < 6 ? k : 63 + k)] = state[static_cast<size_t>(b) * 44 + (k < 6 ? k : 32 + k)];
}
} // namespace

extern "C" int smplpp_solve_mocap_motion(const smplpp_model_t * model, const smplpp_vposer_t * vposer, const char * c3d_path,
                                         const char * mocap_body_yaml_path, const smplpp_ik_options * opt_in,
                                         int32_t warmup_iterations, int32_t iterations, int32_t reproject,
                                         const float * initial_state_host, int64_t first_frame, int64_t frame_count,
                                         float * theta75_out_host, int32_t * status_out_host, float * residual_out_host,
                                         const char * motion_text_path, smplpp_mocap_summary * summary)
{
  if(!model || !c3d_path || !mocap_body_yaml_path || !opt_in || iterations < 1 || warmup_iterations < 0 || !initial_state_host)
    return fail(SMPLPP_ERR_INVALID, "node", "invalid mocap motion arguments!");
  if(opt_in->enable_vposer && !vposer) return fail(SMPLPP_ERR_INVALID, "VPoser", "VPoser decoder is required!");
  // ---- MocapBody.yaml: beta, task names, faces, weights (node.cpp:509-535) ----
  smplpp_mocap_body_t * body = nullptr;
  int rc = smplpp_mocap_body_open(mocap_body_yaml_path, &body);
  if(rc != SMPLPP_OK) return rc;
  std::unique_ptr<smplpp_mocap_body_t, void (*)(smplpp_mocap_body_t *)> body_guard(body, smplpp_mocap_body_close);
  const int n = smplpp_mocap_body_task_count(body);
  if(n < 1) return fail(SMPLPP_ERR_IO, "node", "MocapBody.yaml holds no IK task");
  std::vector<float> beta(kShapeDim), vw0(static_cast<size_t>(n) * 3);
  std::vector<int64_t> faces(n);
  rc = smplpp_mocap_body_get(body, beta.data(), faces.data(), vw0.data());
  if(rc != SMPLPP_OK) return rc;
  // the node keeps its tasks in a std::map: alphabetical order of the names (node.cpp:47, 798)
  std::vector<int> order(n);
  for(int i = 0; i < n; i++) order[i] = i;
  std::sort(order.begin(), order.end(), [&](int a, int b) {
    return std::string(smplpp_mocap_body_task_name(body, a)) < std::string(smplpp_mocap_body_task_name(body, b));
  });
  // ---- C3D: marker index of every task by label suffix (node.cpp:571-595) ----
  smplpp_c3d_t * c3d = nullptr;
  rc = smplpp_c3d_open(c3d_path, &c3d);
  if(rc != SMPLPP_OK) return rc;
  std::unique_ptr<smplpp_c3d_t, void (*)(smplpp_c3d_t *)> c3d_guard(c3d, smplpp_c3d_close);
  const int64_t points = smplpp_c3d_point_count(c3d), total_frames = smplpp_c3d_frame_count(c3d);
  if(first_frame < 0 || first_frame >= total_frames) return fail(SMPLPP_ERR_INVALID, "node", "first mocap frame outside the C3D file");
  const int64_t F = frame_count > 0 ? std::min(frame_count, total_frames - first_frame) : total_frames - first_frame;
  std::vector<int64_t> marker(n), face_sorted(n);
  std::vector<float> vw_sorted(static_cast<size_t>(n) * 3);
  for(int i = 0; i < n; i++)
  {
    const int src = order[i];
    const char * name = smplpp_mocap_body_task_name(body, src);
    marker[i] = smplpp_c3d_find_label(c3d, name);
    if(marker[i] >= points)
      return fail(SMPLPP_ERR_IO, "node", std::string("no C3D point label ends with the IK task name ") + name);
    face_sorted[i] = faces[src];
    for(int k = 0; k < 3; k++) vw_sorted[3 * i + k] = vw0[3 * src + k];
  }
  smplpp_tasks_t * tasks = nullptr;
  rc = smplpp_tasks_create(model, n, face_sorted.data(), &tasks);
  if(rc != SMPLPP_OK) return rc;
  std::unique_ptr<smplpp_tasks_t, void (*)(smplpp_tasks_t *)> tasks_guard(tasks, smplpp_tasks_destroy);

  // ---- targets of every frame (node.cpp:667-691): missing marker -> weight 0, target zeroed ----
  std::vector<float> xyz(static_cast<size_t>(F) * points * 3), target(static_cast<size_t>(F) * n * 3), posw(static_cast<size_t>(F) * n);
  std::vector<uint8_t> ok(static_cast<size_t>(F) * points);
  rc = smplpp_c3d_read(c3d, first_frame, F, xyz.data(), ok.data());
  if(rc != SMPLPP_OK) return rc;
  for(int64_t f = 0; f < F; f++)
    for(int i = 0; i < n; i++)
    {
      const size_t src = static_cast<size_t>(f) * points + marker[i];
      const bool have = ok[src] != 0;
      posw[static_cast<size_t>(f) * n + i] = have ? 1.f : 0.f;
      for(int k = 0; k < 3; k++) target[(static_cast<size_t>(f) * n + i) * 3 + k] = have ? xyz[3 * src + k] : 0.f;
    }

  // ---- options of the motion stage (node.cpp:558-560, 699): phi pinned, fixed beta, skip rule on ----
  smplpp_ik_options opt = *opt_in;
  opt.optimize_beta = 0, opt.enable_phi = 0, opt.phi_limit = 0.f, opt.update_state = 1;
  const int theta_dim = smplpp_ik_theta_dim(&opt);
  cudaStream_t st = nullptr;
  SB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  struct StreamGuard
  {
    cudaStream_t s;
    ~StreamGuard() { cudaStreamDestroy(s); }
  } st_guard{st};
  const size_t B = static_cast<size_t>(F);
  DevBuf d_theta, d_beta, d_vw, d_tgt, d_pw, d_status, d_e, d_res, d_face, d_ws, d_init, d_vw0, d_face0, d_dphi, d_pre, d_th75;
  const size_t ws_step = reproject ? smplpp_ik_faces_workspace_bytes(tasks, &opt, F) : smplpp_ik_workspace_bytes(tasks, &opt, F);
  const size_t ws_rep = reproject ? smplpp_ik_reproject_workspace_bytes(model, tasks, F) : 0;
  if((rc = d_theta.alloc(B * theta_dim * 4)) || (rc = d_beta.alloc(kShapeDim * 4)) || (rc = d_vw.alloc(B * n * 12))
     || (rc = d_tgt.alloc(B * n * 12)) || (rc = d_pw.alloc(B * n * 4)) || (rc = d_status.alloc(B * 4)) || (rc = d_e.alloc(B * n * 16))
     || (rc = d_res.alloc(B * 4)) || (rc = d_face.alloc(B * n * 4)) || (rc = d_ws.alloc(std::max(ws_step, ws_rep) + 512))
     || (rc = d_init.alloc(theta_dim * 4)) || (rc = d_vw0.alloc(n * 12)) || (rc = d_face0.alloc(n * 4)) || (rc = d_dphi.alloc(B * n * 8))
     || (rc = d_pre.alloc(B * theta_dim * 4)) || (rc = d_th75.alloc(B * 75 * 4)))
    return rc;
  std::vector<int32_t> face32(n);
  for(int i = 0; i < n; i++) face32[i] = static_cast<int32_t>(face_sorted[i]);
  SB_CUDA(cudaMemcpyAsync(d_beta.p, beta.data(), kShapeDim * 4, cudaMemcpyHostToDevice, st));
  SB_CUDA(cudaMemcpyAsync(d_tgt.p, target.data(), B * n * 12, cudaMemcpyHostToDevice, st));
  SB_CUDA(cudaMemcpyAsync(d_pw.p, posw.data(), B * n * 4, cudaMemcpyHostToDevice, st));
  SB_CUDA(cudaMemcpyAsync(d_init.p, initial_state_host, theta_dim * 4, cudaMemcpyHostToDevice, st));
  SB_CUDA(cudaMemcpyAsync(d_vw0.p, vw_sorted.data(), n * 12, cudaMemcpyHostToDevice, st));
  SB_CUDA(cudaMemcpyAsync(d_face0.p, face32.data(), n * 4, cudaMemcpyHostToDevice, st));
  SB_CUDA(cudaMemsetAsync(d_dphi.p, 0, B * n * 8, st));
  auto bcast = [&](long long rows) -> int {
    broadcast_rows_kernel<<<static_cast<unsigned>((rows * theta_dim + 255) / 256), 256, 0, st>>>(rows, theta_dim, d_init.as<float>(), d_theta.as<float>());
    SB_LAUNCHED();
    broadcast_rows_kernel<<<static_cast<unsigned>((rows * n * 3 + 255) / 256), 256, 0, st>>>(rows, n * 3, d_vw0.as<float>(), d_vw.as<float>());
    SB_LAUNCHED();
    broadcast_rows_i32_kernel<<<static_cast<unsigned>((rows * n + 255) / 256), 256, 0, st>>>(rows, n, d_face0.as<int32_t>(), d_face.as<int32_t>());
    SB_LAUNCHED();
    return SMPLPP_OK;
  };
  // one iteration of the loop body for `rows` frames starting at row 0 of the device arrays
  auto iterate = [&](int64_t rows, bool last) -> int {
    float * e_out = last ? d_e.as<float>() : nullptr;
    if(!reproject)
      return smplpp_ik_step(model, vposer, tasks, &opt, st, rows, d_theta.as<float>(), d_beta.as<float>(), 0, d_vw.as<float>(),
                            d_tgt.as<float>(), nullptr, d_pw.as<float>(), d_status.as<int32_t>(), e_out, nullptr, nullptr, nullptr,
                            nullptr, d_ws.p, ws_step + 256);
    SB_CUDA(cudaMemcpyAsync(d_pre.p, d_theta.p, static_cast<size_t>(rows) * theta_dim * 4, cudaMemcpyDeviceToDevice, st));
    int r2 = smplpp_ik_step_faces(model, vposer, tasks, &opt, st, rows, d_theta.as<float>(), d_beta.as<float>(), 0, d_vw.as<float>(),
                                  d_face.as<int32_t>(), d_tgt.as<float>(), nullptr, d_pw.as<float>(), d_status.as<int32_t>(), e_out,
                                  nullptr, nullptr, nullptr, nullptr, d_dphi.as<float>(), d_ws.p, ws_step + 256);
    if(r2 != SMPLPP_OK) return r2;
    return smplpp_ik_reproject(model, vposer, tasks, &opt, st, rows, d_pre.as<float>(), d_beta.as<float>(), 0, d_vw.as<float>(),
                               d_face.as<int32_t>(), d_dphi.as<float>(), nullptr, d_ws.p, ws_rep + 256);
  };
  // ---- warm-up on the first frame (node.cpp:1369: the first 31 iterations stay on frame 0), its state starts all frames ----
  if((rc = bcast(1)) != SMPLPP_OK) return rc;
  for(int k = 0; k < warmup_iterations; k++)
    if((rc = iterate(1, false)) != SMPLPP_OK) return rc;
  if(warmup_iterations > 0)
  {
    SB_CUDA(cudaMemcpyAsync(d_init.p, d_theta.p, theta_dim * 4, cudaMemcpyDeviceToDevice, st));
    SB_CUDA(cudaMemcpyAsync(d_vw0.p, d_vw.p, n * 12, cudaMemcpyDeviceToDevice, st));
    SB_CUDA(cudaMemcpyAsync(d_face0.p, d_face.p, n * 4, cudaMemcpyDeviceToDevice, st));
  }
  if((rc = bcast(F)) != SMPLPP_OK) return rc;
  for(int k = 0; k < iterations; k++)
    if((rc = iterate(F, k + 1 == iterations)) != SMPLPP_OK) return rc;
  marker_residual_kernel<<<static_cast<unsigned>((F + 127) / 128), 128, 0, st>>>(static_cast<int>(F), n, d_e.as<float>(), d_pw.as<float>(),
                                                                             d_res.as<float>());
  SB_LAUNCHED();
  // ---- theta (25, 3) of every frame (node.cpp:1374-1391: the latent state goes through the decoder once more) ----
  const float * theta75_dev = d_theta.as<float>();
  if(opt.enable_vposer)
  {
    state_to_theta_kernel<<<static_cast<unsigned>((F * 12 + 127) / 128), 128, 0, st>>>(static_cast<int>(F), d_theta.as<float>(), d_th75.as<float>());
    SB_LAUNCHED();
    rc = launch_vposer_decode(vposer, st, static_cast<int>(F), d_theta.as<float>() + 6, 44, d_th75.as<float>() + 6, 75, nullptr);
    if(rc != SMPLPP_OK) return rc;
    theta75_dev = d_th75.as<float>();
  }
  std::vector<float> theta_h(B * 75), res_h(B);
  std::vector<int32_t> status_h(B);
  SB_CUDA(cudaMemcpyAsync(theta_h.data(), theta75_dev, B * 75 * 4, cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaMemcpyAsync(res_h.data(), d_res.p, B * 4, cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaMemcpyAsync(status_h.data(), d_status.p, B * 4, cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaStreamSynchronize(st));
  if(theta75_out_host) std::memcpy(theta75_out_host, theta_h.data(), B * 75 * 4);
  if(status_out_host) std::memcpy(status_out_host, status_h.data(), B * 4);
  if(residual_out_host) std::memcpy(residual_out_host, res_h.data(), B * 4);
  if(summary)
  {
    *summary = smplpp_mocap_summary{};
    summary->frames = F, summary->markers = n;
    double acc = 0.0;
    for(size_t f = 0; f < B; f++)
    {
      if(status_h[f] == 0)
      {
        summary->solved++;
        acc += res_h[f];
        summary->max_residual = std::max<double>(summary->max_residual, res_h[f]);
      }
      else if(status_h[f] == 1)
        summary->skipped++;
      else
        summary->failed++;
    }
    summary->mean_residual = summary->solved ? acc / static_cast<double>(summary->solved) : 0.0;
  }
  if(motion_text_path) return smplpp_write_motion_text(motion_text_path, F, theta_h.data());
  return SMPLPP_OK;
}
