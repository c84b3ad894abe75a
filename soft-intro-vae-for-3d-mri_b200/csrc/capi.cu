// capi.cu -- the extern "C" surface of libsivae.so (declared in include/sivae.h) and host utilities.
#include <mutex>

#include "sivae_common.cuh"

namespace sivae {

static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static std::atomic<unsigned long long*> g_seed_counter{nullptr};
SeedRef make_seed_ref(unsigned long long seed) { return SeedRef{seed, g_seed_counter.load()}; }

__global__ void advance_seed_counter_kernel(unsigned long long* ctr) { *ctr += 1; }

EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tmap_bf16(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return -1;
  }
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i < rank - 1) gstr[i] = strides_bytes[i];
  }
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr,
                         bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed: CUresult %d (rank %d, base %p, dims %llu/%llu/%llu, box %u/%u/%u)", (int)r,
              rank, base, (unsigned long long)dims[0], (unsigned long long)dims[1],
              (unsigned long long)(rank > 2 ? dims[2] : 0), box[0], box[1], rank > 2 ? box[2] : 0);
    return -1;
  }
  return 0;
}

// implemented in the other translation units
int conv3_igemm(const void*, const void*, void*, int, int, int, int, int, int, cudaStream_t);
size_t conv3_wgrad_workspace_bytes(int, int, int, int, int, int);
int conv3_wgrad(const void*, const void*, float*, void*, size_t, int, int, int, int, int, int, cudaStream_t);
int pack_conv3_weights(const float*, int, int, void*, void*, cudaStream_t);
int upconv3_fprop(const void*, const void*, void*, int, int, int, int, int, int, cudaStream_t);
int upconv3_dgrad(const void*, const void*, void*, int, int, int, int, int, int, cudaStream_t);
size_t upconv3_wgrad_workspace_bytes(int, int, int, int, int, int);
int upconv3_wgrad(const void*, const void*, float*, void*, size_t, int, int, int, int, int, int, cudaStream_t);
int pack_upconv3_weights(const float*, int, int, void*, void*, cudaStream_t);
size_t wgrad_c1_tc_workspace_bytes();
int wgrad_c1_tc(const void*, const float*, float*, float*, float*, int, int, int, int, int, void*, size_t, cudaStream_t);
size_t c1_to_c64_workspace_bytes();
int c1_to_c64_tc(const float*, const float*, const float*, void*, int, int, int, int, int, void*, size_t, cudaStream_t);
size_t conv3_to1_workspace_bytes(int);
int conv3_to1(const void*, const float*, const float*, float*, int, int, int, int, int, int, int, const uint8_t*, float,
              unsigned long long, void*, size_t, cudaStream_t);
size_t bn_workspace_bytes(int);
int bn_train_coeffs(const void*, long long, int, const float*, const float*, float*, float*, long long*, float, float,
                    float*, float*, float*, float*, void*, size_t, cudaStream_t);
int bn_act_fwd(const void*, const float*, const float*, const void*, void*, int, int, int, int, int, float, int,
               const uint8_t*, float, unsigned long long, cudaStream_t);
int bn_act_bwd(const void*, const void*, const void*, const float*, const float*, const float*, const float*, void*,
               void*, float*, float*, int, int, int, int, int, float, int, const uint8_t*, float, unsigned long long,
               void*, size_t, cudaStream_t);
int ncdhw_f32_to_ndhwc_bf16(const float*, void*, int, int, long long, cudaStream_t);
int ndhwc_bf16_to_ncdhw_f32(const void*, float*, int, int, long long, cudaStream_t);
int c1_to_cn(const float*, const float*, const float*, void*, int, int, int, int, int, int, int, int, cudaStream_t);
int cn_to_c1(const void*, const float*, const float*, float*, int, int, int, int, int, int, int, int, const uint8_t*,
             float, unsigned long long, cudaStream_t);
size_t wgrad_c1_workspace_bytes(int, int, int, int, int, int);
int wgrad_c1(const void*, const float*, float*, float*, float*, int, int, int, int, int, int, int, void*, size_t,
             cudaStream_t);
int relu_drop_bwd(const float*, const float*, float*, long long, float, cudaStream_t);
int reparam_fwd(const float*, const float*, const float*, float, float*, long long, cudaStream_t);
int reparam_bwd(const float*, const float*, const float*, float, float*, float*, long long, int, cudaStream_t);
int reparam_draw_fwd(const float*, const float*, float*, float*, long long, unsigned long long, cudaStream_t);
int kl_persample_fwd(const float*, const float*, float*, int, long long, cudaStream_t);
int kl_persample_bwd(const float*, const float*, const float*, float*, float*, int, long long, int, cudaStream_t);
size_t mse_workspace_bytes(int, long long);
int mse_persample_fwd(const float*, const float*, float*, int, long long, void*, size_t, cudaStream_t);
int mse_persample_bwd(const float*, const float*, const float*, float*, float*, int, long long, cudaStream_t);
int adam_step(const sivae_adam_tensor*, int, const float*, float, float, float, long long*, cudaStream_t);
int bn_train_act_fwd(const void*, const void*, void*, int, int, int, int, int, const float*, const float*, float*, float*,
                     long long*, float, float, float, float*, float*, float*, float*, void*, size_t, cudaStream_t);
int intro_loss_e_fwd(const float*, const float*, const float*, const float*, const float*, const float*, int, float, float,
                     float, float, float*, cudaStream_t);
int intro_loss_e_bwd(const float*, const float*, const float*, const float*, const float*, int, float, float, float, float,
                     float*, float*, float*, float*, float*, float*, cudaStream_t);
int intro_loss_d_fwd(const float*, const float*, const float*, const float*, const float*, int, float, float, float, float,
                     float*, cudaStream_t);
int intro_loss_d_bwd(const float*, int, float, float, float, float, float*, float*, float*, float*, float*, cudaStream_t);
size_t volume_stats_workspace_bytes(int);
int volume_stats(const float*, int, long long, float*, void*, size_t, cudaStream_t);
int preprocess_clip_minmax(const float*, float*, int, long long, float, float*, void*, size_t, cudaStream_t);
int affine_resample(const float*, float*, int, int, int, int, const float*, const float*, const float*, cudaStream_t);
int upconv3_fprop_bn(const void*, const void*, void*, int, int, int, int, int, int, const float*, const float*, float*,
                     float*, long long*, float, float, float*, float*, float*, float*, void*, size_t, cudaStream_t);
size_t linear_workspace_bytes(int, int, int);
int linear_fwd(const float*, const float*, const float*, float*, int, int, int, int, void*, size_t, cudaStream_t);
int linear_dgrad(const float*, const float*, float*, int, int, int, void*, size_t, cudaStream_t);
int linear_wgrad(const float*, const float*, float*, float*, int, int, int, cudaStream_t);
int ndhwc_to_flat(const void*, float*, int, int, int, int, const float*, cudaStream_t);
int flat_to_ndhwc(const float*, void*, int, int, int, int, cudaStream_t);
int add_act_fwd(const void*, const void*, void*, long long, float, cudaStream_t);
int add_act_bwd(const void*, const void*, void*, long long, float, cudaStream_t);
size_t similarity_workspace_bytes(int, int);
int similarity_topk(const float*, const float*, int, int, int, int, int, float*, int*, void*, size_t, cudaStream_t);
int c1_to_c64_bn(const float*, const float*, const float*, void*, int, int, int, int, int, const float*, const float*,
                 float*, float*, long long*, float, float, float*, float*, float*, float*, void*, size_t, void*, size_t,
                 cudaStream_t);
int conv3_igemm_bn(const void*, const void*, void*, int, int, int, int, int, int, const float*, const float*, float*,
                   float*, long long*, float, float, float*, float*, float*, float*, void*, size_t, cudaStream_t);

}  // namespace sivae

using namespace sivae;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" {

const char* sivae_last_error(void) { return g_err; }
int sivae_abi_version(void) { return SIVAE_ABI_VERSION; }
long long sivae_launch_count(void) { return g_launches.load(); }

int sivae_set_seed_counter(unsigned long long* device_counter) {
  g_seed_counter.store(device_counter);
  return 0;
}
int sivae_advance_seed_counter(void* stream) {
  unsigned long long* c = g_seed_counter.load();
  SIVAE_CHECK(c != nullptr, "sivae_advance_seed_counter: no counter registered");
  advance_seed_counter_kernel<<<1, 1, 0, ST(stream)>>>(c);
  SIVAE_LAUNCH_OK("advance_seed_counter_kernel");
  return 0;
}

int sivae_device_check(void) {
  int dev = 0;
  if (check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return -1;
  int major = 0;
  if (check_cuda(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev), "cudaDeviceGetAttribute"))
    return -1;
  if (major != 10) {
    set_error("libsivae requires a compute-capability 10.x (Blackwell, sm_100a) device, found major=%d", major);
    return -3;
  }
  return 0;
}

int sivae_pack_conv3_weights(const float* w, int Cout, int Cin, void* wf, void* wd, void* stream) {
  return pack_conv3_weights(w, Cout, Cin, wf, wd, ST(stream));
}
int sivae_conv3_igemm(const void* x, const void* wpack, void* y, int N, int D, int H, int W, int Cin, int Cout,
                      void* stream) {
  return conv3_igemm(x, wpack, y, N, D, H, W, Cin, Cout, ST(stream));
}
int sivae_conv3_igemm_bn(const void* x, const void* wpack, void* y, int N, int D, int H, int W, int Cin, int Cout,
                         const float* gamma, const float* beta, float* rm, float* rv, long long* nbt, float momentum,
                         float eps, float* mean, float* invstd, float* scale, float* shift, void* ws, size_t ws_bytes,
                         void* stream) {
  return conv3_igemm_bn(x, wpack, y, N, D, H, W, Cin, Cout, gamma, beta, rm, rv, nbt, momentum, eps, mean, invstd, scale,
                        shift, ws, ws_bytes, ST(stream));
}
size_t sivae_conv3_wgrad_workspace_bytes(int N, int D, int H, int W, int Cin, int Cout) {
  return conv3_wgrad_workspace_bytes(N, D, H, W, Cin, Cout);
}
int sivae_conv3_wgrad(const void* x, const void* dy, float* dw, void* ws, size_t ws_bytes, int N, int D, int H, int W,
                      int Cin, int Cout, void* stream) {
  return conv3_wgrad(x, dy, dw, ws, ws_bytes, N, D, H, W, Cin, Cout, ST(stream));
}
int sivae_pack_upconv3_weights(const float* w, int Cout, int Cin, void* wup, void* wupT, void* stream) {
  return pack_upconv3_weights(w, Cout, Cin, wup, wupT, ST(stream));
}
int sivae_upconv3_fprop(const void* x_lo, const void* wup, void* y_hi, int N, int D, int H, int W, int Cin, int Cout,
                        void* stream) {
  return upconv3_fprop(x_lo, wup, y_hi, N, D, H, W, Cin, Cout, ST(stream));
}
int sivae_upconv3_dgrad(const void* dy_hi, const void* wupT, void* dx_lo, int N, int D, int H, int W, int Cin, int Cout,
                        void* stream) {
  return upconv3_dgrad(dy_hi, wupT, dx_lo, N, D, H, W, Cin, Cout, ST(stream));
}
size_t sivae_upconv3_wgrad_workspace_bytes(int N, int D, int H, int W, int Cin, int Cout) {
  return upconv3_wgrad_workspace_bytes(N, D, H, W, Cin, Cout);
}
int sivae_upconv3_wgrad(const void* x_lo, const void* dy_hi, float* dw, void* ws, size_t ws_bytes, int N, int D, int H,
                        int W, int Cin, int Cout, void* stream) {
  return upconv3_wgrad(x_lo, dy_hi, dw, ws, ws_bytes, N, D, H, W, Cin, Cout, ST(stream));
}
size_t sivae_wgrad_c64_workspace_bytes(void) { return wgrad_c1_tc_workspace_bytes(); }
int sivae_wgrad_c64(const void* xc, const float* x1, float* dw, float* sum_c, float* sum_1, int N, int D, int H, int W,
                    int flip, void* ws, size_t ws_bytes, void* stream) {
  return wgrad_c1_tc(xc, x1, dw, sum_c, sum_1, N, D, H, W, flip, ws, ws_bytes, ST(stream));
}
size_t sivae_c1_to_c64_workspace_bytes(void) { return c1_to_c64_workspace_bytes(); }
int sivae_c1_to_c64(const float* x1, const float* w, const float* bias, void* y, int N, int D, int H, int W, int flip,
                    void* ws, size_t ws_bytes, void* stream) {
  return c1_to_c64_tc(x1, w, bias, y, N, D, H, W, flip, ws, ws_bytes, ST(stream));
}
size_t sivae_conv3_to1_workspace_bytes(int C) { return conv3_to1_workspace_bytes(C); }
int sivae_conv3_to1(const void* x, const float* w, const float* bias, float* y, int N, int D, int H, int W, int C,
                    int flip, int act, const uint8_t* mask, float p, unsigned long long seed, void* ws, size_t ws_bytes,
                    void* stream) {
  return conv3_to1(x, w, bias, y, N, D, H, W, C, flip, act, mask, p, seed, ws, ws_bytes, ST(stream));
}
size_t sivae_bn_workspace_bytes(int C) { return bn_workspace_bytes(C); }
int sivae_bn_train_coeffs(const void* y, long long nvox, int C, const float* gamma, const float* beta,
                          float* running_mean, float* running_var, long long* nbt, float momentum, float eps,
                          float* mean, float* invstd, float* scale, float* shift, void* ws, size_t ws_bytes,
                          void* stream) {
  return bn_train_coeffs(y, nvox, C, gamma, beta, running_mean, running_var, nbt, momentum, eps, mean, invstd, scale,
                         shift, ws, ws_bytes, ST(stream));
}
int sivae_bn_act_fwd(const void* y, const float* scale, const float* shift, const void* res, void* out, int N, int D,
                     int H, int W, int C, float slope, int resample, uint8_t* mask, float p,
                     unsigned long long seed, void* stream) {
  return bn_act_fwd(y, scale, shift, res, out, N, D, H, W, C, slope, resample, mask, p, seed, ST(stream));
}
int sivae_bn_act_bwd(const void* g, const void* y, const void* res, const float* mean, const float* invstd,
                     const float* gamma, const float* beta, void* dconv, void* dres, float* dgamma, float* dbeta, int N,
                     int D, int H, int W, int C, float slope, int resample, const uint8_t* mask, float p,
                     unsigned long long seed, void* ws, size_t ws_bytes, void* stream) {
  return bn_act_bwd(g, y, res, mean, invstd, gamma, beta, dconv, dres, dgamma, dbeta, N, D, H, W, C, slope, resample,
                    mask, p, seed, ws, ws_bytes, ST(stream));
}
int sivae_c1_to_cn(const float* x1, const float* w, const float* bias, void* y, int N, int D, int H, int W, int C,
                   int T, int flip, int accumulate, void* stream) {
  return c1_to_cn(x1, w, bias, y, N, D, H, W, C, T, flip, accumulate, ST(stream));
}
int sivae_cn_to_c1(const void* x, const float* w, const float* bias, float* y, int N, int D, int H, int W, int C, int T,
                   int flip, int act, const uint8_t* mask, float p, unsigned long long seed, void* stream) {
  return cn_to_c1(x, w, bias, y, N, D, H, W, C, T, flip, act, mask, p, seed, ST(stream));
}
size_t sivae_wgrad_c1_workspace_bytes(int N, int D, int H, int W, int C, int T) {
  return wgrad_c1_workspace_bytes(N, D, H, W, C, T);
}
int sivae_wgrad_c1(const void* xc, const float* x1, float* dw, float* sum_c, float* sum_1, int N, int D, int H, int W,
                   int C, int T, int flip, void* ws, size_t ws_bytes, void* stream) {
  return wgrad_c1(xc, x1, dw, sum_c, sum_1, N, D, H, W, C, T, flip, ws, ws_bytes, ST(stream));
}
int sivae_relu_drop_bwd(const float* g, const float* out, float* dy, long long n, float p, void* stream) {
  return relu_drop_bwd(g, out, dy, n, p, ST(stream));
}
int sivae_reparam_fwd(const float* mu, const float* logvar, const float* eps, float eps_const, float* z, long long n,
                      void* stream) {
  return reparam_fwd(mu, logvar, eps, eps_const, z, n, ST(stream));
}
int sivae_reparam_draw_fwd(const float* mu, const float* logvar, float* eps_out, float* z, long long n,
                           unsigned long long seed, void* stream) {
  return reparam_draw_fwd(mu, logvar, eps_out, z, n, seed, ST(stream));
}
int sivae_reparam_bwd(const float* dz, const float* logvar, const float* eps, float eps_const, float* dmu,
                      float* dlogvar, long long n, int accumulate, void* stream) {
  return reparam_bwd(dz, logvar, eps, eps_const, dmu, dlogvar, n, accumulate, ST(stream));
}
int sivae_kl_persample_fwd(const float* mu, const float* logvar, float* kl, int B, long long n, void* stream) {
  return kl_persample_fwd(mu, logvar, kl, B, n, ST(stream));
}
int sivae_kl_persample_bwd(const float* mu, const float* logvar, const float* g, float* dmu, float* dlogvar, int B,
                           long long n, int accumulate, void* stream) {
  return kl_persample_bwd(mu, logvar, g, dmu, dlogvar, B, n, accumulate, ST(stream));
}
size_t sivae_mse_workspace_bytes(int B, long long n) { return mse_workspace_bytes(B, n); }
int sivae_mse_persample_fwd(const float* x, const float* y, float* r, int B, long long n, void* ws, size_t ws_bytes,
                            void* stream) {
  return mse_persample_fwd(x, y, r, B, n, ws, ws_bytes, ST(stream));
}
int sivae_mse_persample_bwd(const float* x, const float* y, const float* g, float* dx, float* dy, int B, long long n,
                            void* stream) {
  return mse_persample_bwd(x, y, g, dx, dy, B, n, ST(stream));
}
int sivae_c1_to_c64_bn(const float* x1, const float* w, const float* bias, void* y, int N, int D, int H, int W, int flip,
                       const float* gamma, const float* beta, float* rm, float* rv, long long* nbt, float momentum,
                       float eps, float* mean, float* invstd, float* scale, float* shift, void* ws_pack,
                       size_t ws_pack_bytes, void* ws_bn, size_t ws_bn_bytes, void* stream) {
  return c1_to_c64_bn(x1, w, bias, y, N, D, H, W, flip, gamma, beta, rm, rv, nbt, momentum, eps, mean, invstd, scale, shift,
                      ws_pack, ws_pack_bytes, ws_bn, ws_bn_bytes, ST(stream));
}
size_t sivae_linear_workspace_bytes(int B, int K, int J) { return linear_workspace_bytes(B, K, J); }
int sivae_linear_fwd(const float* x, const float* W, const float* bias, float* y, int B, int K, int J, int act, void* ws,
                     size_t ws_bytes, void* stream) {
  return linear_fwd(x, W, bias, y, B, K, J, act, ws, ws_bytes, ST(stream));
}
int sivae_linear_dgrad(const float* dy, const float* W, float* dx, int B, int K, int J, void* ws, size_t ws_bytes,
                       void* stream) {
  return linear_dgrad(dy, W, dx, B, K, J, ws, ws_bytes, ST(stream));
}
int sivae_linear_wgrad(const float* x, const float* dy, float* dW, float* db, int B, int K, int J, void* stream) {
  return linear_wgrad(x, dy, dW, db, B, K, J, ST(stream));
}
int sivae_ndhwc_to_flat(const void* src, float* dst, int B, int S, int C, int Cp, const float* gate, void* stream) {
  return ndhwc_to_flat(src, dst, B, S, C, Cp, gate, ST(stream));
}
int sivae_flat_to_ndhwc(const float* src, void* dst, int B, int S, int C, int Cp, void* stream) {
  return flat_to_ndhwc(src, dst, B, S, C, Cp, ST(stream));
}
int sivae_add_act_fwd(const void* a, const void* b, void* out, long long n, float slope, void* stream) {
  return add_act_fwd(a, b, out, n, slope, ST(stream));
}
int sivae_add_act_bwd(const void* g, const void* out, void* dz, long long n, float slope, void* stream) {
  return add_act_bwd(g, out, dz, n, slope, ST(stream));
}
size_t sivae_similarity_workspace_bytes(int nq, int nd) { return similarity_workspace_bytes(nq, nd); }
int sivae_similarity_topk(const float* q, const float* db, int nq, int nd, int dim, int metric, int k, float* out_scores,
                          int* out_index, void* ws, size_t ws_bytes, void* stream) {
  return similarity_topk(q, db, nq, nd, dim, metric, k, out_scores, out_index, ws, ws_bytes, ST(stream));
}
int sivae_upconv3_fprop_bn(const void* x, const void* wup, void* y, int N, int D, int H, int W, int Cin, int Cout,
                           const float* gamma, const float* beta, float* rm, float* rv, long long* nbt, float momentum,
                           float eps, float* mean, float* invstd, float* scale, float* shift, void* ws, size_t ws_bytes,
                           void* stream) {
  return upconv3_fprop_bn(x, wup, y, N, D, H, W, Cin, Cout, gamma, beta, rm, rv, nbt, momentum, eps, mean, invstd, scale,
                          shift, ws, ws_bytes, ST(stream));
}
size_t sivae_volume_stats_workspace_bytes(int B) { return volume_stats_workspace_bytes(B); }
int sivae_volume_stats(const float* x, int B, long long n, float* stats, void* ws, size_t ws_bytes, void* stream) {
  return volume_stats(x, B, n, stats, ws, ws_bytes, ST(stream));
}
int sivae_preprocess_clip_minmax(const float* x, float* y, int B, long long n, float cut_range, float* stats, void* ws,
                                 size_t ws_bytes, void* stream) {
  return preprocess_clip_minmax(x, y, B, n, cut_range, stats, ws, ws_bytes, ST(stream));
}
int sivae_affine_resample(const float* x, float* y, int B, int D, int H, int W, const float* mats, const float* pad,
                          const float* stats, void* stream) {
  return affine_resample(x, y, B, D, H, W, mats, pad, stats, ST(stream));
}
int sivae_intro_loss_e_fwd(const float* r_real, const float* k_real, const float* r_fake, const float* k_fake,
                           const float* r_rec, const float* k_rec, int B, float scale, float b_rec, float b_kl, float b_neg,
                           float* out, void* stream) {
  return intro_loss_e_fwd(r_real, k_real, r_fake, k_fake, r_rec, k_rec, B, scale, b_rec, b_kl, b_neg, out, ST(stream));
}
int sivae_intro_loss_e_bwd(const float* r_fake, const float* k_fake, const float* r_rec, const float* k_rec, const float* g,
                           int B, float scale, float b_rec, float b_kl, float b_neg, float* d_r_real, float* d_k_real,
                           float* d_r_fake, float* d_k_fake, float* d_r_rec, float* d_k_rec, void* stream) {
  return intro_loss_e_bwd(r_fake, k_fake, r_rec, k_rec, g, B, scale, b_rec, b_kl, b_neg, d_r_real, d_k_real, d_r_fake,
                          d_k_fake, d_r_rec, d_k_rec, ST(stream));
}
int sivae_intro_loss_d_fwd(const float* r_real, const float* k_rec, const float* k_fake, const float* r_rec_rec,
                           const float* r_fake_rec, int B, float scale, float b_rec, float b_kl, float gamma_r, float* out,
                           void* stream) {
  return intro_loss_d_fwd(r_real, k_rec, k_fake, r_rec_rec, r_fake_rec, B, scale, b_rec, b_kl, gamma_r, out, ST(stream));
}
int sivae_intro_loss_d_bwd(const float* g, int B, float scale, float b_rec, float b_kl, float gamma_r, float* d_r_real,
                           float* d_k_rec, float* d_k_fake, float* d_r_rec_rec, float* d_r_fake_rec, void* stream) {
  return intro_loss_d_bwd(g, B, scale, b_rec, b_kl, gamma_r, d_r_real, d_k_rec, d_k_fake, d_r_rec_rec, d_r_fake_rec,
                          ST(stream));
}
int sivae_bn_train_act_fwd(const void* y, const void* res, void* out, int N, int D, int H, int W, int C, const float* gamma,
                           const float* beta, float* rm, float* rv, long long* nbt, float momentum, float eps, float slope,
                           float* mean, float* invstd, float* scale, float* shift, void* ws, size_t ws_bytes, void* stream) {
  return bn_train_act_fwd(y, res, out, N, D, H, W, C, gamma, beta, rm, rv, nbt, momentum, eps, slope, mean, invstd, scale,
                          shift, ws, ws_bytes, ST(stream));
}
int sivae_adam_step(const sivae_adam_tensor* tensors, int ntensors, const float* lr, float beta1, float beta2,
                    float eps, long long* step, void* stream) {
  return adam_step(tensors, ntensors, lr, beta1, beta2, eps, step, ST(stream));
}
int sivae_ncdhw_f32_to_ndhwc_bf16(const float* src, void* dst, int N, int C, long long vox, void* stream) {
  return ncdhw_f32_to_ndhwc_bf16(src, dst, N, C, vox, ST(stream));
}
int sivae_ndhwc_bf16_to_ncdhw_f32(const void* src, float* dst, int N, int C, long long vox, void* stream) {
  return ndhwc_bf16_to_ncdhw_f32(src, dst, N, C, vox, ST(stream));
}

}  // extern "C"
