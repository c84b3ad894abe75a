// optim.cu -- multi-tensor Adam step fused with the bf16 weight re-pack (NEXT-2).
//
// Replaces the two torch.optim.Adam.step() calls of one training iteration (utils/my_trainer.py:183-184 construct
// them, :288 and :324 step them; defaults betas (0.9, 0.999), eps 1e-8, no weight decay, no amsgrad) and the
// fp32 -> bf16 tap-major re-pack of every 3x3x3 convolution weight that follows an update (pack_conv3_weights).
// One launch covers up to kAdamMax parameter tensors: the descriptors travel as kernel arguments, so a captured CUDA
// graph replays them without any pointer table in device memory.  blockIdx.y = tensor, blockIdx.x strides over its
// elements.  lr and the step counter live in device memory (LR schedulers and graph replays keep working).
//
// Update (same operation order as torch/optim/adam.py _single_tensor_adam):
//   m  = m + (g - m) * (1 - beta1)
//   v  = v * beta2 + (1 - beta2) * g * g
//   p -= (lr / (1 - beta1^t)) * m / (sqrt(v) / sqrt(1 - beta2^t) + eps)
#include "sivae_common.cuh"

namespace sivae {

static constexpr int kAdamMax = 32;

struct AdamBatch {
  sivae_adam_tensor t[kAdamMax];
};

__global__ void __launch_bounds__(256)
adam_multi_kernel(const __grid_constant__ AdamBatch batch, const float* __restrict__ lr_p, float beta1, float beta2,
                  float eps, const long long* __restrict__ step_p) {
  const sivae_adam_tensor& t = batch.t[blockIdx.y];
  const double step = (double)(*step_p + 1);
  const float bc1 = (float)(1.0 - pow((double)beta1, step));
  const float bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, step));
  const float step_size = *lr_p / bc1;
  const float w1 = 1.f - beta1, w2 = 1.f - beta2;
  float* __restrict__ p = t.param;
  const float* __restrict__ g = t.grad;
  float* __restrict__ m = t.exp_avg;
  float* __restrict__ v = t.exp_avg_sq;
  __nv_bfloat16* wf = reinterpret_cast<__nv_bfloat16*>(t.pack_fwd);
  __nv_bfloat16* wd = reinterpret_cast<__nv_bfloat16*>(t.pack_dgrad);
  const int cin = t.cin, cout = t.cout;
  const long long n = t.numel;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i];
    float mi = m[i], vi = v[i];
    mi = __fmaf_rn(gi - mi, w1, mi);
    vi = __fmaf_rn(w2 * gi, gi, vi * beta2);
    const float denom = __fsqrt_rn(vi) / bc2_sqrt + eps;
    const float pi = p[i] - step_size * (mi / denom);
    m[i] = mi;
    v[i] = vi;
    p[i] = pi;
    if (wf != nullptr) {   // Conv3d weight [Cout][Cin][27] -> wf[tap][co][ci], wd[26-tap][ci][co]
      const int tap = (int)(i % 27);
      const int ci = (int)((i / 27) % cin);
      const int co = (int)(i / (27ll * cin));
      const __nv_bfloat16 b = __float2bfloat16_rn(pi);
      wf[((long long)tap * cout + co) * cin + ci] = b;
      if (wd != nullptr) wd[((long long)(26 - tap) * cin + ci) * cout + co] = b;
    }
  }
}

__global__ void adam_advance_step_kernel(long long* step) { *step += 1; }

int adam_step(const sivae_adam_tensor* tensors, int ntensors, const float* lr, float beta1, float beta2, float eps,
              long long* step, cudaStream_t st) {
  SIVAE_CHECK(ntensors >= 0 && (ntensors == 0 || tensors != nullptr), "adam_step: bad tensor list");
  SIVAE_CHECK(lr != nullptr && step != nullptr, "adam_step: lr and step must be device pointers");
  SIVAE_CHECK(beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && eps >= 0.f, "adam_step: bad hyper-parameters");
  for (int i = 0; i < ntensors; ++i) {
    const sivae_adam_tensor& t = tensors[i];
    SIVAE_CHECK(t.param && t.grad && t.exp_avg && t.exp_avg_sq && t.numel > 0, "adam_step: tensor %d incomplete", i);
    if (t.pack_fwd != nullptr)
      SIVAE_CHECK(t.cin > 0 && t.cout > 0 && t.numel == 27ll * t.cin * t.cout,
                  "adam_step: tensor %d: pack requested but numel %lld != 27*%d*%d", i, t.numel, t.cin, t.cout);
  }
  for (int base = 0; base < ntensors; base += kAdamMax) {
    AdamBatch b;
    memset(&b, 0, sizeof(b));
    const int nb = ntensors - base < kAdamMax ? ntensors - base : kAdamMax;
    long long biggest = 0;
    for (int i = 0; i < nb; ++i) {
      b.t[i] = tensors[base + i];
      if (b.t[i].numel > biggest) biggest = b.t[i].numel;
    }
    int gx = (int)((biggest + 256 * 4 - 1) / (256 * 4));
    if (gx > 148) gx = 148;
    if (gx < 1) gx = 1;
    adam_multi_kernel<<<dim3((unsigned)gx, (unsigned)nb), 256, 0, st>>>(b, lr, beta1, beta2, eps, step);
    SIVAE_LAUNCH_OK("adam_multi_kernel");
  }
  adam_advance_step_kernel<<<1, 1, 0, st>>>(step);
  SIVAE_LAUNCH_OK("adam_advance_step_kernel");
  return 0;
}

}  // namespace sivae
