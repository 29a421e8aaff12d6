// Target distributions that are not Gaussian mixtures: the planar-robot density of BASELINE
// config C2.  The Student-t mixture (C1 / C4) reuses the Gaussian log-density and mixture-gradient kernels.
#include "common.cuh"

namespace gvi {

constexpr int PR_MAX_LINKS = 64;
constexpr int PR_MAX_GOALS = 8;

// experiments/target_distributions/planar_robot.py:29-66.  One thread per sample (D = #links is 10 in the reference's
// experiments): c_i = theta_0 + ... + theta_i, end effector (sum_i l_i cos c_i, sum_i l_i sin c_i),
//   log p = N(theta; 0, diag(prior_std^2)) + max_g N(pos; goal_g, lik_std^2 I)                      (:49-53, :65-66)
// and, when grad != nullptr, its gradient through the arg-max goal (the reference back-propagates through reduce_max):
//   d/dtheta_i = -theta_i / s_i^2 - [(x - gx) dx_i + (y - gy) dy_i] / lik_std^2,
//   dx_i = -sum_{m >= i} l_m sin c_m,  dy_i = sum_{m >= i} l_m cos c_m.
__global__ void __launch_bounds__(128)
planar_robot_kernel(const float* __restrict__ theta, int N, int D, const float* __restrict__ prior_stds,
                    const float* __restrict__ link_lengths, const float* __restrict__ goals, int G, float lik_std,
                    float* __restrict__ lnpdf, float* __restrict__ grad) {
  __shared__ float s_std[PR_MAX_LINKS], s_len[PR_MAX_LINKS], s_goal[2 * PR_MAX_GOALS];
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    s_std[i] = prior_stds[i];
    s_len[i] = link_lengths ? link_lengths[i] : 1.f;
  }
  for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) s_goal[i] = goals[i];
  __syncthreads();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float* t = theta + (long long)n * D;
  float sn[PR_MAX_LINKS], cs[PR_MAX_LINKS];
  float c = 0.f, x = 0.f, y = 0.f, prior = 0.f, logdet = 0.f;
  for (int i = 0; i < D; ++i) {
    const float th = __ldg(t + i);
    c += th;
    float s_, c_;
    sincosf(c, &s_, &c_);
    sn[i] = s_len[i] * s_;
    cs[i] = s_len[i] * c_;
    x += cs[i];
    y += sn[i];
    const float z = th / s_std[i];
    prior = fmaf(z, z, prior);
    logdet += logf(s_std[i]);
  }
  float best = -INFINITY, bx = 0.f, by = 0.f;
  for (int g = 0; g < G; ++g) {             // first maximum wins, like tf.reduce_max's gradient on ties of distinct goals
    const float dx = x - s_goal[2 * g], dy = y - s_goal[2 * g + 1];
    const float zx = dx / lik_std, zy = dy / lik_std;
    const float l = -0.5f * (zx * zx + zy * zy);
    if (l > best) { best = l; bx = dx; by = dy; }
  }
  lnpdf[n] = (-0.5f * prior - logdet - 0.5f * (float)D * kLog2Pi) + (best - 2.f * logf(lik_std) - kLog2Pi);
  if (grad == nullptr) return;
  float* gr = grad + (long long)n * D;
  const float inv_var = 1.f / (lik_std * lik_std);
  float ssum = 0.f, csum = 0.f;
  for (int i = D - 1; i >= 0; --i) {
    ssum += sn[i];
    csum += cs[i];
    const float th = __ldg(t + i);
    gr[i] = -th / (s_std[i] * s_std[i]) - (by * csum - bx * ssum) * inv_var;
  }
}

}  // namespace gvi

extern "C" int gvi_planar_robot_f32(const float* theta, int N, int D, const float* prior_stds, const float* link_lengths,
                                    const float* goals, int G, float likelihood_std, float* lnpdf, float* grad,
                                    void* stream) {
  using namespace gvi;
  GVI_REQUIRE(N >= 0 && D > 0 && G > 0, "gvi_planar_robot_f32: bad sizes");
  GVI_REQUIRE(D <= PR_MAX_LINKS && G <= PR_MAX_GOALS, "gvi_planar_robot_f32: at most %d links and %d goals", PR_MAX_LINKS,
              PR_MAX_GOALS);
  GVI_REQUIRE(likelihood_std > 0.f, "gvi_planar_robot_f32: likelihood_std must be positive");
  if (N == 0) return GVI_OK;
  GVI_REQUIRE(theta && prior_stds && goals && lnpdf, "gvi_planar_robot_f32: null pointer");
  planar_robot_kernel<<<ceil_div(N, 128), 128, 0, (cudaStream_t)stream>>>(theta, N, D, prior_stds, link_lengths, goals, G,
                                                                          likelihood_std, lnpdf, grad);
  return check_launch("planar_robot_kernel");
}
