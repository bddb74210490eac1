// Shared host/device helpers for libqnnb200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <math.h>
#include <stdlib.h>

#include "../../include/qnnb200.h"

namespace qnnb {

// ------------------------------------------------------------------ error plumbing
void set_error(const char* fmt, ...);
int  cuda_fail(cudaError_t e, const char* what);

#define QNNB_CHECK_ARG(cond, ...)                                   \
  do {                                                              \
    if (!(cond)) { ::qnnb::set_error(__VA_ARGS__); return QNNB_EINVAL; } \
  } while (0)

#define QNNB_CUDA(call)                                             \
  do {                                                              \
    cudaError_t e__ = (call);                                       \
    if (e__ != cudaSuccess) return ::qnnb::cuda_fail(e__, #call);   \
  } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// TensorFlow SAME padding: out = ceil(in/stride); total = max((out-1)*stride + k - in, 0);
// before = total/2 (so 32 -> 16 under stride 2 with k=3 pads 0 before, 1 after).
static inline void same_pad(int size, int k, int stride, int* out, int* before) {
  int o = (size + stride - 1) / stride;
  int total = (o - 1) * stride + k - size;
  if (total < 0) total = 0;
  *out = o;
  *before = total / 2;
}

// ------------------------------------------------------------------ fused epilogue
// Device-side copy of qnnb_epilogue plus derived constants.
struct Epi {
  float acc_scale;
  const float* bias;
  const float* bn_inv;
  const float* bn_shift;
  int   res_kind;
  const void* residual;
  float res_scale;
  float res_mul;
  int   act;
  float qm;          // 2^(abits-1)
  float leaky_alpha;
  int   pool;
};

static inline Epi make_epi(const qnnb_epilogue& e) {
  Epi d;
  d.acc_scale = e.acc_scale;
  d.bias = e.bias;
  d.bn_inv = e.bn_inv;
  d.bn_shift = e.bn_shift;
  d.res_kind = e.res_kind;
  d.residual = e.residual;
  d.res_scale = e.res_scale;
  d.res_mul = e.res_mul;
  d.act = e.act;
  d.qm = (float)(1 << ((e.abits > 0 ? e.abits : 1) - 1));
  d.leaky_alpha = e.leaky_alpha;
  d.pool = e.pool;
  return d;
}

int validate_epilogue(const qnnb_epilogue& e, bool allow_pool, bool allow_residual);

// true when x is a finite power of two whose folded constants stay far from the subnormal range
static inline bool is_pow2_scale(float x) {
  int ex;
  float m = frexpf(x, &ex);
  return x > 0.f && m == 0.5f && ex > -40 && ex < 40;
}

#ifdef __CUDACC__
// ---- programmatic dependent launch (PDL) -----------------------------------------------------------------
// Consecutive layers are separate kernels on one stream.  Every kernel of the path (a) lets the NEXT kernel's CTAs
// start as soon as SM resources free up (launch_dependents at entry) and (b) runs its own prologue -- barrier init,
// TMEM allocation, tensor-map prefetch, staging of the layer's kernel weights, none of which depends on the previous
// layer -- before griddepcontrol.wait, which returns once the previous kernel has completed and its writes are
// visible.  Launch latency and prologue of layer L+1 thus overlap the tail of layer L.  Without the launch attribute
// (QNNB_PDL=0) both instructions are no-ops.
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Kernel launch with the programmatic-stream-serialization attribute (captured as a programmatic edge in CUDA graphs).
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("QNNB_PDL"); enabled = (e && e[0] == '0') ? 0 : 1; }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = enabled ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// Per-channel constants of the affine part (steps 1-3 of the fixed order).
struct ChanConst {
  float scale, bias, inv, shift;
  bool  has_bias, has_bn;
};

__device__ __forceinline__ ChanConst load_chan(const Epi& e, int ch, bool valid) {
  ChanConst c;
  c.scale = e.acc_scale;
  c.has_bias = (e.bias != nullptr);
  c.has_bn = (e.bn_inv != nullptr);
  c.bias = (c.has_bias && valid) ? __ldg(e.bias + ch) : 0.f;
  c.inv = (c.has_bn && valid) ? __ldg(e.bn_inv + ch) : 1.f;
  c.shift = (c.has_bn && valid) ? __ldg(e.bn_shift + ch) : 0.f;
  return c;
}

// y = ((float(acc) * s) + bias) * inv + shift, each step a separate RN op (no FMA contraction).
__device__ __forceinline__ float affine(float accf, const ChanConst& c) {
  float v = __fmul_rn(accf, c.scale);
  if (c.has_bias) v = __fadd_rn(v, c.bias);
  if (c.has_bn) { v = __fmul_rn(v, c.inv); v = __fadd_rn(v, c.shift); }
  return v;
}

// the affine map is non-decreasing in acc unless the BN slope is negative
__device__ __forceinline__ bool decreasing(const ChanConst& c) { return c.has_bn && c.inv < 0.f; }

__device__ __forceinline__ float add_residual(float y, float shortcut, float res_mul) {
  return __fmul_rn(__fadd_rn(shortcut, y), res_mul);
}

// quantized_tanh -> integer level (rintf = round-half-to-even like tf.round)
__device__ __forceinline__ int act_quant(float z, float qm) {
  float q = rintf(__fmul_rn(z, qm));
  q = fminf(fmaxf(q, -qm), qm - 1.f);
  return (int)q;
}

// binary_tanh(z) == +1  <=>  z > 2^-24   
__device__ __forceinline__ bool act_sign(float z) { return z > 5.9604644775390625e-08f; }

// ---- pre-scaled form of the same pipeline for the tensor-core epilogues ---------------------------------
// q = clamp(rint(z * qm)) with z = ((f*s + bias)*inv + shift).  qm is a power of two, so it folds into inv and
// shift without changing any rounding (scaling by 2^k commutes with round-to-nearest as long as nothing is
// subnormal); when s is a power of two as well (int8 x int8 layers) it folds too:
//   FOLD : zq = ((f + bias/s) * (inv*s*qm)) + shift*qm          general: zq = ((f*s + bias) * (inv*qm)) + shift*qm
// Absent bias / BN are the neutral constants 0 / 1 / 0 (adding 0 and multiplying by a power of two are exact).
struct QConst { float s, a, b, c; };

template <bool FOLD>
__device__ __forceinline__ QConst make_qconst(const Epi& e, int ch, bool valid) {
  const float bias = (e.bias != nullptr && valid) ? __ldg(e.bias + ch) : 0.f;
  const float inv = (e.bn_inv != nullptr && valid) ? __ldg(e.bn_inv + ch) : 1.f;
  const float shift = (e.bn_inv != nullptr && valid) ? __ldg(e.bn_shift + ch) : 0.f;
  QConst q;
  if (FOLD) {
    q.s = 1.f;
    q.a = __fmul_rn(bias, __frcp_rn(e.acc_scale));
    q.b = __fmul_rn(__fmul_rn(inv, e.acc_scale), e.qm);
  } else {
    q.s = e.acc_scale;
    q.a = bias;
    q.b = __fmul_rn(inv, e.qm);
  }
  q.c = __fmul_rn(shift, e.qm);
  return q;
}

template <bool FOLD>
__device__ __forceinline__ float qaffine(int acc, const QConst& q) {
  float f = (float)acc;                         // cvt.rn
  if (!FOLD) f = __fmul_rn(f, q.s);
  f = __fadd_rn(f, q.a);
  f = __fmul_rn(f, q.b);
  return __fadd_rn(f, q.c);                     // == z * qm
}

// clamp bounds are integers, so clamping before the round-to-nearest-even conversion is identical to after.
// The conversion itself is one RN add of 1.5 * 2^23: for |v| <= 2^22 the sum's ulp is 1, so the add rounds v to the
// nearest integer (ties to even, the magic constant being even) and leaves it, in two's complement, in the low
// mantissa bits -- same result as cvt.rni.s32.f32 but on the FMA pipe instead of the quarter-rate conversion unit.
// Only the LOW BYTE of the returned word is the level (callers store it as int8).
__device__ __forceinline__ int quant_scaled(float zq, float qm) {
  return __float_as_int(__fadd_rn(fminf(fmaxf(zq, -qm), qm - 1.f), 12582912.f));
}

__device__ __forceinline__ float act_leaky(float z, float alpha) { return z > 0.f ? z : __fmul_rn(alpha, z); }
#endif  // __CUDACC__

// ------------------------------------------------------------------ kernel launchers (one per .cu)
int launch_pack_weights(int mode, int nb, float H, const float* w, int kh, int kw, int cin, int cout,
                        int wfmt, void* out, float* scratch, cudaStream_t st);
int launch_conv_generic(const qnnb_conv_desc& d, const void* x, const void* w, void* y, cudaStream_t st);
bool conv_tc_supported(const qnnb_conv_desc& d, const char** why);
bool conv_tc_v1_supported(const qnnb_conv_desc& d);
int launch_conv_tc(const qnnb_conv_desc& d, const void* x, const void* w, void* y, cudaStream_t st);
bool conv_first_tc_shape(const qnnb_conv_desc& d);
int launch_conv_first_tc(const qnnb_conv_desc& d, const void* x, const void* w, void* y, cudaStream_t st);
bool conv_f32_tc_supported(const qnnb_conv_desc& d, const char** why);
int launch_conv_f32_tc(const qnnb_conv_desc& d, const void* x, const void* w, void* y, cudaStream_t st);
void set_trace_buffer(unsigned long long* buf, int cap);
unsigned long long* get_trace_buffer();
bool vgg_fused_supported(const qnnb_vgg_desc& d, const char** why);
long long vgg_fused_blob_bytes(const qnnb_vgg_desc& d);
int launch_vgg_pack(const qnnb_vgg_desc& d, void* blob, cudaStream_t st);
int launch_vgg_fused(const qnnb_vgg_desc& d, const void* blob, const void* x, float* y, cudaStream_t st);
int launch_dense(const qnnb_dense_desc& d, const void* x, const void* w, float* y, float* logits, cudaStream_t st);

}  // namespace qnnb
