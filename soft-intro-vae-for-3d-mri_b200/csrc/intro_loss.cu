// intro_loss.cu -- the introspective loss assembly of one Soft-IntroVAE iteration as two tiny fused kernels per phase
// (SURVEY.md section 8 rows a9 / a10): the [B]-vector arithmetic that utils/my_trainer.py:260-284 (lossE) and :301-321
// (lossD) express as ~25 torch ops each (and autograd as ~60 more backward launches).
//
//   lossE = 10 * ( s * (b_rec * mean(r_real) + b_kl * mean(k_real))
//                  + 0.5 * ( mean_b exp(-2 s (b_rec * r_fake[b] + b_neg * k_fake[b]))
//                          + mean_b exp(-2 s (b_rec * r_rec[b]  + b_neg * k_rec[b])) ) )
//   lossD = 10 * s * ( b_rec * mean(r_real) + 0.5 b_kl (mean(k_rec) + mean(k_fake))
//                      + g_r * 0.5 b_rec (mean(r_rec_rec) + mean(r_fake_rec)) )
// r_* = per-sample squared reconstruction errors, k_* = per-sample KL terms (both [B], produced by the mse / kl kernels).
// One warp; fp32, batch means as torch computes them (sum / B).
#include "sivae_common.cuh"

namespace sivae {

// out[0..4] = lossE, mean(r_real), mean(k_real), exp_elbo_fake, exp_elbo_rec
__global__ void intro_loss_e_fwd_kernel(const float* __restrict__ r_real, const float* __restrict__ k_real,
                                        const float* __restrict__ r_fake, const float* __restrict__ k_fake,
                                        const float* __restrict__ r_rec, const float* __restrict__ k_rec, int B,
                                        float scale, float b_rec, float b_kl, float b_neg, float* __restrict__ out) {
  float s_rr = 0.f, s_kr = 0.f, s_ef = 0.f, s_er = 0.f;
  for (int b = threadIdx.x; b < B; b += 32) {
    s_rr += r_real[b];
    s_kr += k_real[b];
    s_ef += expf(-2.f * scale * (b_rec * r_fake[b] + b_neg * k_fake[b]));
    s_er += expf(-2.f * scale * (b_rec * r_rec[b] + b_neg * k_rec[b]));
  }
  s_rr = warp_sum(s_rr); s_kr = warp_sum(s_kr); s_ef = warp_sum(s_ef); s_er = warp_sum(s_er);
  if (threadIdx.x == 0) {
    const float inv = 1.f / (float)B;
    const float m_rr = s_rr * inv, m_kr = s_kr * inv, ef = s_ef * inv, er = s_er * inv;
    out[0] = 10.f * (scale * (b_rec * m_rr + b_kl * m_kr) + 0.5f * (ef + er));
    out[1] = m_rr; out[2] = m_kr; out[3] = ef; out[4] = er;
  }
}

// gradients of lossE * g[0] with respect to the six [B] vectors (any output may be NULL)
__global__ void intro_loss_e_bwd_kernel(const float* __restrict__ r_fake, const float* __restrict__ k_fake,
                                        const float* __restrict__ r_rec, const float* __restrict__ k_rec,
                                        const float* __restrict__ g, int B, float scale, float b_rec, float b_kl,
                                        float b_neg, float* d_r_real, float* d_k_real, float* d_r_fake, float* d_k_fake,
                                        float* d_r_rec, float* d_k_rec) {
  const float go = 10.f * g[0] / (float)B;
  for (int b = threadIdx.x; b < B; b += 32) {
    if (d_r_real) d_r_real[b] = go * scale * b_rec;
    if (d_k_real) d_k_real[b] = go * scale * b_kl;
    const float ef = expf(-2.f * scale * (b_rec * r_fake[b] + b_neg * k_fake[b]));
    const float er = expf(-2.f * scale * (b_rec * r_rec[b] + b_neg * k_rec[b]));
    if (d_r_fake) d_r_fake[b] = go * 0.5f * ef * (-2.f * scale * b_rec);
    if (d_k_fake) d_k_fake[b] = go * 0.5f * ef * (-2.f * scale * b_neg);
    if (d_r_rec) d_r_rec[b] = go * 0.5f * er * (-2.f * scale * b_rec);
    if (d_k_rec) d_k_rec[b] = go * 0.5f * er * (-2.f * scale * b_neg);
  }
}

// out[0..5] = lossD, mean(r_real), mean(k_rec), mean(k_fake), mean(r_rec_rec), mean(r_fake_rec)
__global__ void intro_loss_d_fwd_kernel(const float* __restrict__ r_real, const float* __restrict__ k_rec,
                                        const float* __restrict__ k_fake, const float* __restrict__ r_rec_rec,
                                        const float* __restrict__ r_fake_rec, int B, float scale, float b_rec, float b_kl,
                                        float gamma_r, float* __restrict__ out) {
  float s[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  for (int b = threadIdx.x; b < B; b += 32) {
    s[0] += r_real[b]; s[1] += k_rec[b]; s[2] += k_fake[b]; s[3] += r_rec_rec[b]; s[4] += r_fake_rec[b];
  }
#pragma unroll
  for (int i = 0; i < 5; ++i) s[i] = warp_sum(s[i]) / (float)B;
  if (threadIdx.x == 0) {
    out[0] = 10.f * (scale * (b_rec * s[0] + 0.5f * b_kl * (s[1] + s[2]) + gamma_r * 0.5f * b_rec * (s[3] + s[4])));
#pragma unroll
    for (int i = 0; i < 5; ++i) out[1 + i] = s[i];
  }
}

__global__ void intro_loss_d_bwd_kernel(const float* __restrict__ g, int B, float scale, float b_rec, float b_kl,
                                        float gamma_r, float* d_r_real, float* d_k_rec, float* d_k_fake,
                                        float* d_r_rec_rec, float* d_r_fake_rec) {
  const float go = 10.f * g[0] * scale / (float)B;
  for (int b = threadIdx.x; b < B; b += 32) {
    if (d_r_real) d_r_real[b] = go * b_rec;
    if (d_k_rec) d_k_rec[b] = go * 0.5f * b_kl;
    if (d_k_fake) d_k_fake[b] = go * 0.5f * b_kl;
    if (d_r_rec_rec) d_r_rec_rec[b] = go * gamma_r * 0.5f * b_rec;
    if (d_r_fake_rec) d_r_fake_rec[b] = go * gamma_r * 0.5f * b_rec;
  }
}

int intro_loss_e_fwd(const float* r_real, const float* k_real, const float* r_fake, const float* k_fake,
                     const float* r_rec, const float* k_rec, int B, float scale, float b_rec, float b_kl, float b_neg,
                     float* out, cudaStream_t st) {
  SIVAE_CHECK(B > 0 && r_real && k_real && r_fake && k_fake && r_rec && k_rec && out, "intro_loss_e_fwd: bad arguments");
  intro_loss_e_fwd_kernel<<<1, 32, 0, st>>>(r_real, k_real, r_fake, k_fake, r_rec, k_rec, B, scale, b_rec, b_kl, b_neg, out);
  SIVAE_LAUNCH_OK("intro_loss_e_fwd_kernel");
  return 0;
}
int intro_loss_e_bwd(const float* r_fake, const float* k_fake, const float* r_rec, const float* k_rec, const float* g,
                     int B, float scale, float b_rec, float b_kl, float b_neg, float* d_r_real, float* d_k_real,
                     float* d_r_fake, float* d_k_fake, float* d_r_rec, float* d_k_rec, cudaStream_t st) {
  SIVAE_CHECK(B > 0 && r_fake && k_fake && r_rec && k_rec && g, "intro_loss_e_bwd: bad arguments");
  intro_loss_e_bwd_kernel<<<1, 32, 0, st>>>(r_fake, k_fake, r_rec, k_rec, g, B, scale, b_rec, b_kl, b_neg, d_r_real,
                                             d_k_real, d_r_fake, d_k_fake, d_r_rec, d_k_rec);
  SIVAE_LAUNCH_OK("intro_loss_e_bwd_kernel");
  return 0;
}
int intro_loss_d_fwd(const float* r_real, const float* k_rec, const float* k_fake, const float* r_rec_rec,
                     const float* r_fake_rec, int B, float scale, float b_rec, float b_kl, float gamma_r, float* out,
                     cudaStream_t st) {
  SIVAE_CHECK(B > 0 && r_real && k_rec && k_fake && r_rec_rec && r_fake_rec && out, "intro_loss_d_fwd: bad arguments");
  intro_loss_d_fwd_kernel<<<1, 32, 0, st>>>(r_real, k_rec, k_fake, r_rec_rec, r_fake_rec, B, scale, b_rec, b_kl, gamma_r, out);
  SIVAE_LAUNCH_OK("intro_loss_d_fwd_kernel");
  return 0;
}
int intro_loss_d_bwd(const float* g, int B, float scale, float b_rec, float b_kl, float gamma_r, float* d_r_real,
                     float* d_k_rec, float* d_k_fake, float* d_r_rec_rec, float* d_r_fake_rec, cudaStream_t st) {
  SIVAE_CHECK(B > 0 && g, "intro_loss_d_bwd: bad arguments");
  intro_loss_d_bwd_kernel<<<1, 32, 0, st>>>(g, B, scale, b_rec, b_kl, gamma_r, d_r_real, d_k_rec, d_k_fake, d_r_rec_rec,
                                             d_r_fake_rec);
  SIVAE_LAUNCH_OK("intro_loss_d_bwd_kernel");
  return 0;
}

}  // namespace sivae
