// C-ABI entry points of libqnnb200.so (declared in include/qnnb200.h): argument validation,
// kernel selection, error plumbing.  No persistent state, no allocation.
#include "common.cuh"

#include <string.h>

namespace qnnb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  return QNNB_ECUDA;
}

int validate_epilogue(const qnnb_epilogue& e, bool allow_pool, bool allow_residual) {
  QNNB_CHECK_ARG(e.act >= QNNB_ACT_NONE && e.act <= QNNB_ACT_SIGN_I8, "epilogue: bad act %d", e.act);
  QNNB_CHECK_ARG(e.act != QNNB_ACT_QUANT || (e.abits >= 2 && e.abits <= 8), "epilogue: abits=%d outside 2..8", e.abits);
  QNNB_CHECK_ARG((e.bn_inv == nullptr) == (e.bn_shift == nullptr), "epilogue: bn_inv and bn_shift must come together");
  QNNB_CHECK_ARG(e.pool == 0 || (allow_pool && e.pool == 2), "epilogue: pool=%d not supported here", e.pool);
  QNNB_CHECK_ARG(e.res_kind == QNNB_KIND_NONE || (allow_residual && (e.res_kind == QNNB_KIND_I8 || e.res_kind == QNNB_KIND_F32)),
                 "epilogue: res_kind=%d not supported here", e.res_kind);
  QNNB_CHECK_ARG(e.res_kind == QNNB_KIND_NONE || e.residual != nullptr, "epilogue: residual pointer is null");
  QNNB_CHECK_ARG(e.res_kind == QNNB_KIND_NONE || e.pool == 0, "epilogue: residual and pool cannot be combined");
  QNNB_CHECK_ARG(e.act != QNNB_ACT_LEAKY || e.leaky_alpha >= 0.f, "epilogue: leaky_alpha must be >= 0 (monotone pooling)");
  QNNB_CHECK_ARG(e.acc_scale > 0.f, "epilogue: acc_scale must be > 0");
  return QNNB_OK;
}

static int validate_conv(const qnnb_conv_desc& d) {
  QNNB_CHECK_ARG(d.n >= 0 && d.h > 0 && d.w > 0 && d.cin > 0 && d.cout > 0, "conv2d: bad shape n=%d h=%d w=%d cin=%d cout=%d", d.n, d.h, d.w, d.cin, d.cout);
  QNNB_CHECK_ARG(d.kh >= 1 && d.kh <= 3 && d.kw >= 1 && d.kw <= 3, "conv2d: kernel %dx%d outside 1..3", d.kh, d.kw);
  QNNB_CHECK_ARG(d.stride == 1 || d.stride == 2, "conv2d: stride %d not in {1,2}", d.stride);
  QNNB_CHECK_ARG(d.in_kind == QNNB_KIND_U8 || d.in_kind == QNNB_KIND_I8 || d.in_kind == QNNB_KIND_B1 || d.in_kind == QNNB_KIND_F32,
                 "conv2d: bad in_kind %d", d.in_kind);
  QNNB_CHECK_ARG(d.w_f32 == 0 || (d.w_f32 == 1 && d.in_kind == QNNB_KIND_F32), "conv2d: an fp32 kernel (w_f32) needs fp32 activations");
  QNNB_CHECK_ARG(d.max_ctas >= 0, "conv2d: max_ctas=%d must be >= 0", d.max_ctas);
  int rc = validate_epilogue(d.epi, true, true);
  if (rc) return rc;
  if (d.epi.pool) {
    int oh, pt, ow, pl;
    same_pad(d.h, d.kh, d.stride, &oh, &pt);
    same_pad(d.w, d.kw, d.stride, &ow, &pl);
    QNNB_CHECK_ARG(oh >= 2 && ow >= 2, "conv2d: pooled output would be empty");
  }
  return QNNB_OK;
}

}  // namespace qnnb

using namespace qnnb;

extern "C" {

int qnnb_version(void) { return QNNB_VERSION; }

const char* qnnb_last_error(void) { return g_err; }

int qnnb_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor) {
  int dev = 0;
  QNNB_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  QNNB_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return QNNB_OK;
}

int64_t qnnb_packed_weight_bytes(int32_t wfmt, int32_t kh, int32_t kw, int32_t cin, int32_t cout) {
  if (kh <= 0 || kw <= 0 || cin <= 0 || cout <= 0) return 0;
  if (wfmt == QNNB_WFMT_I8) return (int64_t)cout * kh * kw * ((cin + 3) / 4 * 4);
  if (wfmt == QNNB_WFMT_B1) return (int64_t)cout * kh * kw * ((cin + 31) / 32) * 4;
  if (wfmt == QNNB_WFMT_F32) return (int64_t)cout * kh * kw * ((cin + 3) / 4 * 4) * 4;
  return 0;
}

int qnnb_pack_weights(int32_t mode, int32_t nb, float H, const float* w_hwio, int32_t kh, int32_t kw, int32_t cin,
                      int32_t cout, int32_t wfmt, void* out, float* scratch, void* stream) {
  return launch_pack_weights(mode, nb, H, w_hwio, kh, kw, cin, cout, wfmt, out, scratch, (cudaStream_t)stream);
}

int qnnb_conv2d_tc_supported(const qnnb_conv_desc* d) {
  if (!d || validate_conv(*d) != QNNB_OK) return 0;
  const char* why = "";
  if (d->w_f32) return 0;                  /* fp32 kernels: CUDA-core FFMA path only */
  if (d->in_kind == QNNB_KIND_F32) return conv_f32_tc_supported(*d, &why) ? 1 : 0;
  return conv_tc_supported(*d, &why) ? 1 : 0;
}

int qnnb_conv2d_out_shape(const qnnb_conv_desc* d, int32_t* oh, int32_t* ow) {
  QNNB_CHECK_ARG(d && oh && ow, "conv2d_out_shape: null pointer");
  int rc = validate_conv(*d);
  if (rc) return rc;
  int o1, o2, p;
  same_pad(d->h, d->kh, d->stride, &o1, &p);
  same_pad(d->w, d->kw, d->stride, &o2, &p);
  if (d->epi.pool) { o1 /= 2; o2 /= 2; }
  *oh = o1; *ow = o2;
  return QNNB_OK;
}

int qnnb_conv2d(const qnnb_conv_desc* d, const void* x, const void* w, void* y, void* stream) {
  QNNB_CHECK_ARG(d, "conv2d: null descriptor");
  int rc = validate_conv(*d);
  if (rc) return rc;
  if (d->n == 0) return QNNB_OK;           /* empty batch: nothing to launch, pointers may be null */
  QNNB_CHECK_ARG(x && w && y, "conv2d: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const char* why = "";
  if (d->w_f32) {
    // 'float' networks: fp32 kernel x fp32 activations on the FFMA path (K4's bf16 split is exact for LEVELS only)
    if (d->impl == QNNB_IMPL_TCGEN05 || d->impl == QNNB_IMPL_TCGEN05_V1) { set_error("conv2d: fp32 kernels have no tcgen05 path"); return QNNB_EUNSUPPORTED; }
    return launch_conv_generic(*d, x, w, y, st);
  }
  if (d->in_kind == QNNB_KIND_F32 && (d->impl == QNNB_IMPL_AUTO || d->impl == QNNB_IMPL_TCGEN05)) {
    // fp32 activations: bf16-split tensor-core kernel where the shape allows it
    if (conv_f32_tc_supported(*d, &why)) return launch_conv_f32_tc(*d, x, w, y, st);
    if (d->impl == QNNB_IMPL_TCGEN05) { set_error("conv2d: tcgen05 path does not cover this shape: %s", why); return QNNB_EUNSUPPORTED; }
    return launch_conv_generic(*d, x, w, y, st);
  }
  const bool tc_ok = conv_tc_supported(*d, &why);
  if (d->impl == QNNB_IMPL_TCGEN05_V1) {
    if (!tc_ok || !conv_tc_v1_supported(*d)) { set_error("conv2d: tcgen05 v1 kernel does not cover this shape"); return QNNB_EUNSUPPORTED; }
    return launch_conv_tc(*d, x, w, y, st);
  }
  if (d->impl == QNNB_IMPL_TCGEN05) {
    if (!tc_ok) { set_error("conv2d: tcgen05 path does not cover this shape: %s", why); return QNNB_EUNSUPPORTED; }
    return launch_conv_tc(*d, x, w, y, st);
  }
  if (d->impl == QNNB_IMPL_AUTO && tc_ok) return launch_conv_tc(*d, x, w, y, st);
  QNNB_CHECK_ARG(d->impl == QNNB_IMPL_AUTO || d->impl == QNNB_IMPL_GENERIC, "conv2d: bad impl %d", d->impl);
  return launch_conv_generic(*d, x, w, y, st);
}

int qnnb_debug_set_trace(void* buf, int64_t nwords) {
  set_trace_buffer((unsigned long long*)buf, (int)nwords);
  return QNNB_OK;
}

int qnnb_dense(const qnnb_dense_desc* d, const void* x, const void* w, float* y, float* logits, void* stream) {
  QNNB_CHECK_ARG(d, "dense: null descriptor");
  QNNB_CHECK_ARG(d->n >= 0 && d->fin > 0 && d->units > 0, "dense: bad shape n=%d fin=%d units=%d", d->n, d->fin, d->units);
  QNNB_CHECK_ARG(d->n == 0 || (x && w && y), "dense: null pointer");
  QNNB_CHECK_ARG(d->in_kind == QNNB_KIND_I8 || d->in_kind == QNNB_KIND_B1 || d->in_kind == QNNB_KIND_F32, "dense: bad in_kind %d", d->in_kind);
  int rc = validate_epilogue(d->epi, false, false);
  if (rc) return rc;
  QNNB_CHECK_ARG(d->epi.act == QNNB_ACT_NONE, "dense: fused activation not supported (act=%d)", d->epi.act);
  QNNB_CHECK_ARG(d->w_f32 == 0 || (d->w_f32 == 1 && d->in_kind == QNNB_KIND_F32), "dense: an fp32 kernel (w_f32) needs fp32 input");
  QNNB_CHECK_ARG(d->max_ctas >= 0, "dense: max_ctas=%d must be >= 0", d->max_ctas);
  QNNB_CHECK_ARG(d->avg_positions >= 0 && (d->avg_positions <= 1 || d->in_kind == QNNB_KIND_F32),
                 "dense: avg_positions=%d needs fp32 input", d->avg_positions);
  if (d->n == 0) return QNNB_OK;
  return launch_dense(*d, x, w, y, logits, (cudaStream_t)stream);
}

int qnnb_vgg_forward_supported(const qnnb_vgg_desc* d) {
  if (!d) return 0;
  const char* why = "";
  return vgg_fused_supported(*d, &why) ? 1 : 0;
}

static int validate_vgg(const qnnb_vgg_desc* d, const char* who) {
  QNNB_CHECK_ARG(d, "%s: null descriptor", who);
  QNNB_CHECK_ARG(d->n >= 0, "%s: bad batch size %d", who, d->n);
  QNNB_CHECK_ARG(d->nconv >= 1 && d->nconv <= QNNB_NET_MAX_CONVS, "%s: nconv=%d outside 1..%d", who, d->nconv, QNNB_NET_MAX_CONVS);
  QNNB_CHECK_ARG(d->max_ctas >= 0, "%s: max_ctas=%d must be >= 0", who, d->max_ctas);
  for (int l = 0; l < d->nconv; ++l) {
    int rc = validate_epilogue(d->conv[l].epi, true, false);
    if (rc) return rc;
  }
  int rc = validate_epilogue(d->dense_epi, false, false);
  if (rc) return rc;
  QNNB_CHECK_ARG(d->dense_epi.act == QNNB_ACT_NONE, "%s: the dense head has no activation (act=%d)", who, d->dense_epi.act);
  return QNNB_OK;
}

int64_t qnnb_vgg_blob_bytes(const qnnb_vgg_desc* d) {
  if (!d || validate_vgg(d, "vgg_blob_bytes") != QNNB_OK) return 0;
  return (int64_t)vgg_fused_blob_bytes(*d);
}

int qnnb_vgg_pack(const qnnb_vgg_desc* d, void* blob, void* stream) {
  int rc = validate_vgg(d, "vgg_pack");
  if (rc) return rc;
  return launch_vgg_pack(*d, blob, (cudaStream_t)stream);
}

int qnnb_vgg_forward(const qnnb_vgg_desc* d, const void* blob, const void* x, float* y, void* stream) {
  int rc = validate_vgg(d, "vgg_forward");
  if (rc) return rc;
  const char* why = "";
  if (!vgg_fused_supported(*d, &why)) { set_error("vgg_forward: outside the whole-network kernel's scope (%s)", why); return QNNB_EUNSUPPORTED; }
  if (d->n == 0) return QNNB_OK;
  QNNB_CHECK_ARG(blob && x && y, "vgg_forward: null pointer");
  return launch_vgg_fused(*d, blob, x, y, (cudaStream_t)stream);
}

}  // extern "C"
