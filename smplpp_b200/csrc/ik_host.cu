// Host-buffer entry points of the IK path: what a caller of the reference's mocap modes does per frame
// (node/node.cpp:645-1002: targets from the C3D frame -> K iterations of the IK step -> theta) for a whole batch of
// frames at once, with HOST arrays in and out.  bench.py's `ik.e2e` times smplpp_ik_solve_host.
#include <algorithm>
#include <cstring>

#include "common.cuh"
#include "tasks.cuh"

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
