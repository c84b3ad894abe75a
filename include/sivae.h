/*
 * sivae.h -- C ABI of libsivae.so: hand-written sm_100a kernels for the training hot path of
 * M-hayatooo/Soft-intro-VAE-for-3D-MRI (encoder/decoder forward+backward and the introspective loss).
 *
 * The reference has no FFI / plugin layer: every operation below is an *implicit library call*
 * the reference makes through torch.nn (cuDNN / ATen).  Each entry point cites the reference
 * call site (path:line relative to the reference checkout) whose computation it replaces.
 * INTEGRATION.md shows the ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *   - plain pointers + sizes only; every pointer is a DEVICE pointer unless marked "host".
 *   - activations are NDHWC ("channels-last-3d") bf16: x[n][d][h][w][c]; 1-channel model inputs /
 *     outputs / latents are fp32 [n][d][h][w] (identical bytes to the reference's NCDHW with C=1).
 *   - C (channels) of bf16 activations must be a multiple of 8; the tcgen05 convolutions require
 *     Cin, Cout in {64, 128, 256} (the headline network; other widths are channel-padded by the host).
 *   - the caller owns every buffer including workspaces; no allocation, no host sync inside.
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it.
 *   - return value: 0 on success, negative on error; sivae_last_error() gives the message
 *     (thread-local).
 */
#ifndef SIVAE_H_
#define SIVAE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SIVAE_ABI_VERSION 1

/* resample modes of the fused BN/activation pass */
#define SIVAE_RESAMPLE_NONE 0
#define SIVAE_RESAMPLE_AVGPOOL2 1 /* nn.AvgPool3d(kernel_size=2)           models/models.py:20 */
#define SIVAE_RESAMPLE_UPSAMPLE2 2 /* nn.Upsample(scale_factor=2), nearest  models/models.py:58 */

const char* sivae_last_error(void);
int sivae_abi_version(void);
/* 0 if the current device is compute capability 10.x, negative otherwise. */
int sivae_device_check(void);
/* number of kernels launched by this library since load (diagnostic for bench.py "gpu_launches"). */
long long sivae_launch_count(void);

/* Dropout epoch counter (nn.Dropout draws a fresh mask on every call, models/models.py:95,122,140).  Every Philox
 * dropout call is keyed by (seed argument, *device_counter): registering a device-resident uint64 and advancing it
 * once per training step lets a captured CUDA graph, whose seed arguments are baked in, draw new masks on each replay.
 * Pass NULL to unregister (the seed argument alone keys the masks). */
int sivae_set_seed_counter(unsigned long long* device_counter);
int sivae_advance_seed_counter(void* stream);

/* ------------------------------------------------------------------------------------------------
 * 3x3x3 convolutions, stride 1, pad 1, no bias (nn.Conv3d in BuildingBlock / UpsampleBuildingkBlock,
 * models/models.py:17,21,55,59) and their autograd (lossE.backward(), utils/my_trainer.py:287,323).
 * ---------------------------------------------------------------------------------------------- */

/* fp32 [Cout][Cin][3][3][3] (torch layout) -> two bf16 packs:
 *   wf[tap][Cout][Cin]            forward operand
 *   wd[26-tap][Cin][Cout]         data-gradient operand (spatially flipped, channel-transposed)
 * either output may be NULL. */
int sivae_pack_conv3_weights(const float* w, int Cout, int Cin, void* wf_bf16, void* wd_bf16, void* stream);

/* y[n,d,h,w,co] = sum_{tap,ci} x[n,d+kd-1,h+kh-1,w+kw-1,ci] * wpack[tap][co][ci]   (zero padding)
 * Implicit GEMM on tcgen05 (TMA-staged NDHWC tiles, fp32 accumulation in TMEM, bf16 store).
 * Forward uses wf; the data gradient is the same call on dy with wd (Cin/Cout swapped). */
int sivae_conv3_igemm(const void* x_bf16, const void* wpack_bf16, void* y_bf16,
                      int N, int D, int H, int W, int Cin, int Cout, void* stream);

/* dw[co][ci][tap] = sum_{n,d,h,w} dy[n,d,h,w,co] * x[n,d+kd-1,h+kh-1,w+kw-1,ci]   (fp32, torch layout)
 * Split-K tcgen05 GEMM over voxels + deterministic second-stage reduction. */
/* Conv3d(k=3,p=1,bias=False) followed by the train-mode statistics of BatchNorm3d on its output (models/models.py:17-18,
 * :21-22, :55-56): y as sivae_conv3_igemm, then mean / invstd / scale = gamma*invstd / shift = beta - mean*scale and the
 * running-stat update exactly as sivae_bn_train_coeffs.  When the persistent convolution kernel takes the shape the
 * per-channel sums come out of its epilogue (no extra pass over y); otherwise the separate statistics pass runs.
 * workspace: sivae_bn_workspace_bytes(Cout). */
int sivae_conv3_igemm_bn(const void* x_bf16, const void* wpack_bf16, void* y_bf16,
                         int N, int D, int H, int W, int Cin, int Cout,
                         const float* gamma, const float* beta, float* running_mean, float* running_var,
                         long long* num_batches_tracked, float momentum, float eps,
                         float* mean, float* invstd, float* scale, float* shift,
                         void* workspace, size_t workspace_bytes, void* stream);
size_t sivae_conv3_wgrad_workspace_bytes(int N, int D, int H, int W, int Cin, int Cout);
int sivae_conv3_wgrad(const void* x_bf16, const void* dy_bf16, float* dw,
                      void* workspace, size_t workspace_bytes,
                      int N, int D, int H, int W, int Cin, int Cout, void* stream);

/* nn.Upsample(scale_factor=2, nearest) followed by the 3x3x3 convolution (UpsampleBuildingkBlock, models/models.py:58-59)
 * WITHOUT materialising the upsampled tensor: an output voxel (2d+pd, 2h+ph, 2w+pw) sees only 2x2x2 distinct low-res
 * inputs, so each of the 8 output parities is an 8-tap convolution on the low-res grid with pre-summed weights
 * (27 -> 8 multiply-accumulates per output; exact in real arithmetic, weights are summed in fp32 before the bf16 cast).
 *   x_lo  [N][D][H][W][Cin],  y_hi / dy_hi [N][2D][2H][2W][Cout]   (D, H, W are the LOW-resolution extents)
 *   wup  bf16 [64 = parity*8 + abc][Cout][Cin],  wupT bf16 [64][Cin][Cout]   (from sivae_pack_upconv3_weights)
 * fprop:  y_hi = conv3(upsample2(x_lo), w)
 * dgrad:  dx_lo = upsample2^T(conv3^T(dy_hi))       (what autograd computes through Upsample + Conv3d)
 * wgrad:  dw fp32 [Cout][Cin][3][3][3] = conv3 weight gradient with input upsample2(x_lo) and output gradient dy_hi */
int sivae_pack_upconv3_weights(const float* w, int Cout, int Cin, void* wup_bf16, void* wupT_bf16, void* stream);
int sivae_upconv3_fprop(const void* x_lo_bf16, const void* wup_bf16, void* y_hi_bf16,
                        int N, int D, int H, int W, int Cin, int Cout, void* stream);
/* fprop + train-mode BatchNorm3d coefficients of the high-res output (models/models.py:58-60), as sivae_conv3_igemm_bn:
 * statistics from the convolution epilogue when the persistent kernel takes the shape.  workspace:
 * sivae_bn_workspace_bytes(Cout). */
int sivae_upconv3_fprop_bn(const void* x_lo_bf16, const void* wup_bf16, void* y_hi_bf16,
                           int N, int D, int H, int W, int Cin, int Cout,
                           const float* gamma, const float* beta, float* running_mean, float* running_var,
                           long long* num_batches_tracked, float momentum, float eps,
                           float* mean, float* invstd, float* scale, float* shift,
                           void* workspace, size_t workspace_bytes, void* stream);
int sivae_upconv3_dgrad(const void* dy_hi_bf16, const void* wupT_bf16, void* dx_lo_bf16,
                        int N, int D, int H, int W, int Cin, int Cout, void* stream);
size_t sivae_upconv3_wgrad_workspace_bytes(int N, int D, int H, int W, int Cin, int Cout);
int sivae_upconv3_wgrad(const void* x_lo_bf16, const void* dy_hi_bf16, float* dw,
                        void* workspace, size_t workspace_bytes,
                        int N, int D, int H, int W, int Cin, int Cout, void* stream);

/* Single-output-channel 3x3x3 convolution on tcgen05 (N=16 accumulator tile, column 0 used), fp32 output with the
 * epilogue fused:   y[v] = act( bias[0] + sum_{tap,c} w[c][tap'] * x[v + delta(tap)][c] ),  tap' = flip ? 26-tap : tap.
 * Decoder tail Conv3d(C,1,3)+ReLU+Dropout(.35) (models/models.py:137-140; act=1) and the input gradient of the
 * encoder stem Conv3d(1,C,3) (models/models.py:92; flip=1, act=0).  w is the torch weight as it lies in memory
 * ([1][C][3][3][3] or [C][1][3][3][3] = w[c][tap]); dropout as in sivae_bn_act_fwd with mask [N][D][H][W]. */
size_t sivae_conv3_to1_workspace_bytes(int C);
int sivae_conv3_to1(const void* x_bf16, const float* w, const float* bias, float* y,
                    int N, int D, int H, int W, int C, int flip, int act,
                    const uint8_t* mask, float p, unsigned long long seed,
                    void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * BatchNorm3d (train mode) + (Leaky)ReLU + residual + AvgPool/Upsample + Dropout, fused
 * (nn.BatchNorm3d models/models.py:18,22,56,60,93,119; LeakyReLU :15,19,94,121; residual :39-41;
 *  AvgPool3d :20; Upsample :58; Dropout :95,122).
 * ---------------------------------------------------------------------------------------------- */

size_t sivae_bn_workspace_bytes(int C);

/* Batch statistics of y (bf16 [nvox][C]) and the affine coefficients of the normalisation:
 *   mean[c], invstd[c] = 1/sqrt(var_biased+eps), scale[c] = gamma*invstd, shift[c] = beta-mean*scale
 * Updates running_mean/running_var (momentum, unbiased variance) and num_batches_tracked (int64)
 * in place when they are non-NULL, exactly as torch's native_batch_norm in training mode. */
int sivae_bn_train_coeffs(const void* y_bf16, long long nvox, int C,
                          const float* gamma, const float* beta,
                          float* running_mean, float* running_var, long long* num_batches_tracked,
                          float momentum, float eps,
                          float* mean, float* invstd, float* scale, float* shift,
                          void* workspace, size_t workspace_bytes, void* stream);

/* Train-mode BatchNorm3d statistics + coefficients + apply (+ residual) + (Leaky)ReLU in ONE call, no resampling and no
 * dropout (models/models.py:18-19, :22 + :39-41): out = act(bn(y) (+ res)); mean / invstd / scale / shift and the running
 * statistics as sivae_bn_train_coeffs.  Runs as statistics / finalize / apply launches; with SIVAE_BN_CLUSTER=1 in the
 * environment tensors up to 3 Mi elements run as a single launch of one 16-CTA thread-block cluster (partials exchanged
 * through distributed shared memory) -- an experiment that measured slower inside the training step.
 * workspace: sivae_bn_workspace_bytes(C). */
int sivae_bn_train_act_fwd(const void* y_bf16, const void* res_bf16, void* out_bf16, int N, int D, int H, int W, int C,
                           const float* gamma, const float* beta, float* running_mean, float* running_var,
                           long long* num_batches_tracked, float momentum, float eps, float slope,
                           float* mean, float* invstd, float* scale, float* shift,
                           void* workspace, size_t workspace_bytes, void* stream);
/* out = resample( dropout( act( y*scale[c]+shift[c] (+ res) ) ) )
 *   act(t) = t>0 ? t : slope*t        (slope 0.2 = LeakyReLU(0.2), slope 0 = ReLU)
 *   dropout: mask != NULL, seed == 0: caller-provided keep-mask (uint8 NDHWC, 0/1), read;
 *            mask != NULL, seed != 0, p > 0 (resample = none only): KEEP-BIT STORE -- the kernel draws Philox(seed) and
 *              WRITES its decisions to mask as uint8 [N*D*H*W*C/8], bit k of byte i = element 8i+k kept;
 *              sivae_bn_act_bwd with the same (mask, seed, p) reads them instead of regenerating Philox;
 *            mask == NULL: Philox(seed) if p > 0, else none.   Kept values are scaled by 1/(1-p).
 * y, res: [N][D][H][W][C] bf16; out: [N][D'][H'][W'][C] with D' = D/2, D or 2D per `resample`. */
int sivae_bn_act_fwd(const void* y_bf16, const float* scale, const float* shift, const void* res_bf16,
                     void* out_bf16, int N, int D, int H, int W, int C, float slope, int resample,
                     uint8_t* mask, float p, unsigned long long seed, void* stream);

/* Backward of sivae_bn_train_coeffs + sivae_bn_act_fwd given g = dLoss/d(out) (bf16, out's shape):
 *   dconv = gradient w.r.t. y (bf16), dres = gradient w.r.t. res (bf16, may be NULL),
 *   dgamma, dbeta fp32 [C] (may be NULL).  Full train-mode BN backward:
 *   dy = gamma*invstd*(dt - mean(dt) - xhat*mean(dt*xhat)).                                       */
int sivae_bn_act_bwd(const void* g_bf16, const void* y_bf16, const void* res_bf16,
                     const float* mean, const float* invstd, const float* gamma, const float* beta,
                     void* dconv_bf16, void* dres_bf16, float* dgamma, float* dbeta,
                     int N, int D, int H, int W, int C, float slope, int resample,
                     const uint8_t* mask, float p, unsigned long long seed,
                     void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Thin convolutions with a single channel on one side (direct, HBM-bound; no tensor cores):
 *   encoder stem Conv3d(1,in_ch,3)        models/models.py:92
 *   decoder tail Conv3d(in_ch,1,3)+ReLU+Dropout(.35)   models/models.py:137-140
 *   mu / var heads Conv3d(C,1,1)          models/models.py:216-217
 *   decoder stem Conv3d(1,C,1)            models/models.py:118
 * T is the tap count: 27 (3x3x3, pad 1) or 1 (1x1x1).  Weights are the torch tensors as they lie in
 * memory: both [C][1][k][k][k] and [1][C][k][k][k] are w[c][t].  flip=1 uses tap 26-t (transposed conv).
 * ---------------------------------------------------------------------------------------------- */

/* y[v][c] (+)= bias[c] + sum_t w[c][t] * x1[v + delta(t)]           x1 fp32 [N][D][H][W], y bf16 NDHWC */
int sivae_c1_to_cn(const float* x1, const float* w, const float* bias, void* y_bf16,
                   int N, int D, int H, int W, int C, int T, int flip, int accumulate, void* stream);

/* The 3x3x3, C = 64 case of sivae_c1_to_cn on tensor cores (encoder stem Conv3d(1,64,3), models/models.py:92, and the
 * input gradient of the decoder tail): the im2col of the one-channel fp32 input is built in shared memory and multiplied
 * on tcgen05 with a bf16 hi/lo split of both operands (~fp32 products), bf16 NDHWC output.  No accumulate mode. */
size_t sivae_c1_to_c64_workspace_bytes(void);
int sivae_c1_to_c64(const float* x1, const float* w, const float* bias, void* y_bf16,
                    int N, int D, int H, int W, int flip, void* workspace, size_t workspace_bytes, void* stream);

/* Encoder stem Conv3d(1,64,3,bias) + train-mode BatchNorm3d coefficients of its output (models/models.py:92-93) in one
 * call: y as sivae_c1_to_c64, coefficients / running stats as sivae_bn_train_coeffs, the channel sums taken in the
 * convolution epilogue.  workspaces: sivae_c1_to_c64_workspace_bytes() and sivae_bn_workspace_bytes(64). */
int sivae_c1_to_c64_bn(const float* x1, const float* w, const float* bias, void* y_bf16,
                       int N, int D, int H, int W, int flip,
                       const float* gamma, const float* beta, float* running_mean, float* running_var,
                       long long* num_batches_tracked, float momentum, float eps,
                       float* mean, float* invstd, float* scale, float* shift,
                       void* pack_workspace, size_t pack_workspace_bytes,
                       void* bn_workspace, size_t bn_workspace_bytes, void* stream);
/* y[v] = act( bias[0] + sum_{t,c} w[c][t] * x[v + delta(t)][c] )     x bf16 NDHWC, y fp32 [N][D][H][W]
 *   act = 0: identity; act = 1: ReLU followed by dropout (mask / Philox(seed) / p as above).        */
int sivae_cn_to_c1(const void* x_bf16, const float* w, const float* bias, float* y,
                   int N, int D, int H, int W, int C, int T, int flip, int act,
                   const uint8_t* mask, float p, unsigned long long seed, void* stream);

/* dw[c][t] = sum_v xc[v][c] * x1[v + delta(t)];  sum_c[c] = sum_v xc[v][c];  sum_1[0] = sum_v x1[v]
 * (sum_c, sum_1 may be NULL).  Deterministic two-stage reduction. */
size_t sivae_wgrad_c1_workspace_bytes(int N, int D, int H, int W, int C, int T);
int sivae_wgrad_c1(const void* xc_bf16, const float* x1, float* dw, float* sum_c, float* sum_1,
                   int N, int D, int H, int W, int C, int T, int flip,
                   void* workspace, size_t workspace_bytes, void* stream);

/* The 3x3x3, C = 64 case of sivae_wgrad_c1 on tensor cores (split-K GEMM over voxels; the one-channel operand's im2col
 * is built in shared memory with a bf16 hi/lo split, the 64-channel operand streams through TMA).  dw fp32 [64][27]. */
size_t sivae_wgrad_c64_workspace_bytes(void);
int sivae_wgrad_c64(const void* xc_bf16, const float* x1, float* dw, float* sum_c, float* sum_1,
                    int N, int D, int H, int W, int flip, void* workspace, size_t workspace_bytes, void* stream);

/* backward of ReLU+Dropout on the decoder output: dy[i] = out[i] > 0 ? g[i]/(1-p) : 0 */
int sivae_relu_drop_bwd(const float* g, const float* out, float* dy, long long n, float p, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Latent / loss kernels (fp32, coalesced float4 + warp-shuffle reductions)
 *   reparameterize           models/models.py:263-271
 *   calc_kl                  utils/my_trainer.py:38-48   (= lossf.kld_loss, models/lossf.py:14-18)
 *   calc_reconstruction_loss utils/my_trainer.py:62-78   (= lossf.mse_loss, models/lossf.py:5-12)
 * ---------------------------------------------------------------------------------------------- */

/* z = mu + eps*exp(0.5*logvar), three separately rounded fp32 operations (bit-equal to torch).
 * eps == NULL -> the scalar eps_const is used (validation path, eps = 0.1). */
int sivae_reparam_fwd(const float* mu, const float* logvar, const float* eps, float eps_const,
                      float* z, long long n, void* stream);
/* The sampler of the training path (models/models.py:265-266, eps = torch.randn_like(std)): eps ~ N(0,1) is drawn IN the
 * kernel (Philox4x32-10 keyed by `seed` and the registered dropout epoch counter -> a captured graph draws fresh noise on
 * every replay; Box-Muller), written to eps_out [n] for sivae_reparam_bwd, and z = mu + eps*exp(0.5*logvar) as above. */
int sivae_reparam_draw_fwd(const float* mu, const float* logvar, float* eps_out, float* z, long long n,
                           unsigned long long seed, void* stream);
/* dmu (+)= dz ; dlogvar (+)= dz*eps*0.5*exp(0.5*logvar).  accumulate=1 adds into dmu/dlogvar. */
int sivae_reparam_bwd(const float* dz, const float* logvar, const float* eps, float eps_const,
                      float* dmu, float* dlogvar, long long n, int accumulate, void* stream);

/* kl[b] = -0.5 * sum_j (1 + logvar - mu^2 - exp(logvar))     mu, logvar: [B][n] */
int sivae_kl_persample_fwd(const float* mu, const float* logvar, float* kl, int B, long long n, void* stream);
/* dmu (+)= g[b]*mu ; dlogvar (+)= g[b]*0.5*(exp(logvar)-1) */
int sivae_kl_persample_bwd(const float* mu, const float* logvar, const float* g, float* dmu, float* dlogvar,
                           int B, long long n, int accumulate, void* stream);

/* r[b] = sum_j (x[b][j]-y[b][j])^2 */
size_t sivae_mse_workspace_bytes(int B, long long n);
int sivae_mse_persample_fwd(const float* x, const float* y, float* r, int B, long long n,
                            void* workspace, size_t workspace_bytes, void* stream);
/* dx = 2*(x-y)*g[b], dy = -dx; either output may be NULL (Q13: both operands can need a gradient). */
int sivae_mse_persample_bwd(const float* x, const float* y, const float* g, float* dx, float* dy,
                            int B, long long n, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Introspective loss assembly (utils/my_trainer.py:260-284 lossE, :301-321 lossD) on the per-sample [B] vectors
 * r_* = sum of squared errors (sivae_mse_persample_fwd) and k_* = KL (sivae_kl_persample_fwd).
 *   lossE = 10*( s*(b_rec*mean(r_real) + b_kl*mean(k_real))
 *                + 0.5*(mean_b exp(-2s(b_rec*r_fake[b] + b_neg*k_fake[b])) + mean_b exp(-2s(b_rec*r_rec[b] + b_neg*k_rec[b]))) )
 *   out[5] = {lossE, mean(r_real), mean(k_real), exp_elbo_fake, exp_elbo_rec}
 *   lossD = 10*s*( b_rec*mean(r_real) + 0.5*b_kl*(mean(k_rec)+mean(k_fake)) + gamma_r*0.5*b_rec*(mean(r_rec_rec)+mean(r_fake_rec)) )
 *   out[6] = {lossD, mean(r_real), mean(k_rec), mean(k_fake), mean(r_rec_rec), mean(r_fake_rec)}
 * The backward entry points write d(loss * g[0]) / d(vector); any gradient pointer may be NULL.
 * ---------------------------------------------------------------------------------------------- */
int sivae_intro_loss_e_fwd(const float* r_real, const float* k_real, const float* r_fake, const float* k_fake,
                           const float* r_rec, const float* k_rec, int B, float scale, float beta_rec, float beta_kl,
                           float beta_neg, float* out, void* stream);
int sivae_intro_loss_e_bwd(const float* r_fake, const float* k_fake, const float* r_rec, const float* k_rec,
                           const float* g, int B, float scale, float beta_rec, float beta_kl, float beta_neg,
                           float* d_r_real, float* d_k_real, float* d_r_fake, float* d_k_fake, float* d_r_rec,
                           float* d_k_rec, void* stream);
int sivae_intro_loss_d_fwd(const float* r_real, const float* k_rec, const float* k_fake, const float* r_rec_rec,
                           const float* r_fake_rec, int B, float scale, float beta_rec, float beta_kl, float gamma_r,
                           float* out, void* stream);
int sivae_intro_loss_d_bwd(const float* g, int B, float scale, float beta_rec, float beta_kl, float gamma_r,
                           float* d_r_real, float* d_k_rec, float* d_k_fake, float* d_r_rec_rec, float* d_r_fake_rec,
                           void* stream);

/* ------------------------------------------------------------------------------------------------
 * Layout helpers at the module boundary (used by parity tests and for non-unit channel inputs)
 * ---------------------------------------------------------------------------------------------- */
/* fp32 NCDHW -> bf16 NDHWC and back */
int sivae_ncdhw_f32_to_ndhwc_bf16(const float* src, void* dst_bf16, int N, int C, long long vox, void* stream);
int sivae_ndhwc_bf16_to_ncdhw_f32(const void* src_bf16, float* dst, int N, int C, long long vox, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Optimiser step fused with the weight re-pack (SURVEY.md section 8f NEXT-2).
 * Replaces optimizer_e.step() / optimizer_d.step() (utils/my_trainer.py:288,:324; torch.optim.Adam constructed at
 * :183-184 with lr 2e-4 and default betas / eps) and the pack_conv3_weights call per updated 3x3x3 weight.
 * ---------------------------------------------------------------------------------------------- */
typedef struct sivae_adam_tensor {
  float* param;            /* fp32 parameter, updated in place */
  const float* grad;       /* fp32 gradient */
  float* exp_avg;          /* Adam first moment */
  float* exp_avg_sq;       /* Adam second moment */
  long long numel;
  void* pack_fwd;          /* optional: bf16 wf[27][Cout][Cin] of a Conv3d(k=3) weight [Cout][Cin][27], refreshed */
  void* pack_dgrad;        /* optional: bf16 wd[27][Cin][Cout] (flipped + transposed) */
  int cout, cin;           /* only read when pack_fwd != NULL */
} sivae_adam_tensor;
/* `tensors` is a HOST array (descriptors travel as kernel arguments); `lr` and `step` are DEVICE scalars: the update
 * uses t = *step + 1 for the bias corrections and *step is incremented once at the end of the call. */
int sivae_adam_step(const sivae_adam_tensor* tensors, int ntensors, const float* lr, float beta1, float beta2,
                    float eps, long long* step, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Latent retrieval (SURVEY.md section 8f NEXT-4): top-k similarity search over encoder latents, the content-based
 * image retrieval the reference states as its goal (README.md:4-11; latent extraction loop: logistic1.ipynb cell 7).
 * q [nq][dim], db [nd][dim] fp32 row-major; metric 0 = cosine similarity, 1 = negative squared L2 distance (larger is
 * more similar for both); out_scores / out_index [nq][k], best first, ties broken by the lower database index;
 * 1 <= k <= 32.  workspace: sivae_similarity_workspace_bytes(nq, nd).
 * ---------------------------------------------------------------------------------------------- */
size_t sivae_similarity_workspace_bytes(int nq, int nd);
int sivae_similarity_topk(const float* q, const float* db, int nq, int nd, int dim, int metric, int k,
                          float* out_scores, int* out_index, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * On-GPU input pipeline (SURVEY.md section 8f NEXT-3).  x, y: fp32 [B][n] / [B][D][H][W] on the device.
 * ---------------------------------------------------------------------------------------------- */
/* stats[b] = {mean, population std, min, max} of volume b.  workspace: sivae_volume_stats_workspace_bytes(B). */
size_t sivae_volume_stats_workspace_bytes(int B);
int sivae_volume_stats(const float* x, int B, long long n, float* stats, void* workspace, size_t workspace_bytes,
                       void* stream);
/* BrainDataset._preprocess (utils/data_load.py:25-30): y = minmax(clip(x, 0, cut_range * std(x))) per volume
 * (cut_range = 4 in the reference); also returns the raw-volume stats.  In place (y == x) is allowed. */
int sivae_preprocess_clip_minmax(const float* x, float* y, int B, long long n, float cut_range, float* stats,
                                 void* workspace, size_t workspace_bytes, void* stream);
/* Trilinear resampling through per-volume 3x4 matrices mats[b] mapping OUTPUT voxel (d,h,w,1) to INPUT voxel
 * coordinates (tio.RandomAffine, aug-z-1200main.py:114; identity rows leave a volume unchanged).  Samples outside the
 * volume take pad[b]; with pad == NULL the per-volume minimum stats[b][2] (torchio default_pad_value='minimum'), or 0
 * when stats is NULL as well.  y must not alias x. */
int sivae_affine_resample(const float* x, float* y, int B, int D, int H, int W, const float* mats, const float* pad,
                          const float* stats, void* stream);

/* ------------------------------------------------------------------------------------------------
 * FC-latent variant (SURVEY.md section 8f NEXT-1): the Linear heads of models/mymodel.py (:125 fc, :150-153 dfc+ReLU)
 * and the layout changes around them.  x [B][K], y / dy [B][J], W / dW [J][K] (torch nn.Linear layout), bias / db [J],
 * all fp32 row-major on the device.  Weight-streaming kernels (HBM-bound for B <= 8), fp32 FMA accumulation,
 * deterministic split reductions.  workspace: sivae_linear_workspace_bytes(B, K, J) for fwd and dgrad.
 * ---------------------------------------------------------------------------------------------- */
size_t sivae_linear_workspace_bytes(int B, int K, int J);
/* y = act(x W^T + bias); act 0 = none (mymodel.py:141), 1 = ReLU (mymodel.py:152); bias may be NULL. */
int sivae_linear_fwd(const float* x, const float* W, const float* bias, float* y, int B, int K, int J, int act,
                     void* workspace, size_t workspace_bytes, void* stream);
/* dx = dy W */
int sivae_linear_dgrad(const float* dy, const float* W, float* dx, int B, int K, int J, void* workspace,
                       size_t workspace_bytes, void* stream);
/* dW = dy^T x, db = column sums of dy (db may be NULL) */
int sivae_linear_wgrad(const float* x, const float* dy, float* dW, float* db, int B, int K, int J, void* stream);
/* x.view(B, -1) of an NCDHW tensor (mymodel.py:140) from the NDHWC bf16 trunk: dst[b][c*S + s] = src[b][s][c], c < C <= Cp
 * (Cp = padded channel count); with gate != NULL entries whose gate[b][c*S + s] <= 0 are zeroed (ReLU gradient). */
int sivae_ndhwc_to_flat(const void* src, float* dst, int B, int S, int C, int Cp, const float* gate, void* stream);
/* y.view(B, C, d, h, w) (mymodel.py:219) into NDHWC bf16 with channels zero-padded to Cp. */
int sivae_flat_to_ndhwc(const float* src, void* dst, int B, int S, int C, int Cp, void* stream);
/* out = LeakyReLU_slope(a + b) on n bf16 elements (n % 8 == 0) -- mymodel.py:136, where the skip branch carries its own
 * activation; and dz = g * (out > 0 ? 1 : slope), the gradient with respect to both a and b. */
int sivae_add_act_fwd(const void* a, const void* b, void* out, long long n, float slope, void* stream);
int sivae_add_act_bwd(const void* g, const void* out, void* dz, long long n, float slope, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SIVAE_H_ */
