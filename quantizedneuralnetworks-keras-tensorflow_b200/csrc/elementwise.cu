// Stand-alone (un-fused) fp32 layer ops: used when a layer object is called on its own rather
// than through a fused plan.  Memory-bound, one element (or one packed word) per thread,
// grid-stride, 128-bit accesses where the shape allows.
//   quantize_act : quantized_tanh (layers/quantized_ops.py:87-100) / binary_tanh (layers/binary_ops.py:37-51)
//   batchnorm    : keras BatchNormalization inference, x*inv + shift
//   maxpool2     : MaxPooling2D(2,2) valid (models/vgg.py:23)
//   leaky        : LeakyReLU (models/model_factory.py:27)
#include "common.cuh"

namespace qnnb {

namespace {

__global__ void quant_act_kernel(const float* __restrict__ x, long long count, float qm, int8_t* __restrict__ y) {
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += stride)
    y[i] = (int8_t)act_quant(x[i], qm);
}

// one thread per output word; rows = count / channels
__global__ void sign_act_kernel(const float* __restrict__ x, long long rows, int channels, int words, uint32_t* __restrict__ y) {
  long long total = rows * words;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += stride) {
    long long r = i / words;
    int wd = (int)(i % words);
    uint32_t bits = 0;
    for (int b = 0; b < 32; ++b) {
      int c = wd * 32 + b;
      if (c < channels && act_sign(x[r * channels + c])) bits |= 1u << b;
    }
    y[i] = bits;
  }
}

__global__ void bn_kernel(const float* __restrict__ x, long long total, int ch, const float* __restrict__ inv,
                          const float* __restrict__ shift, float* __restrict__ y) {
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += stride) {
    int c = (int)(i % ch);
    y[i] = __fadd_rn(__fmul_rn(x[i], __ldg(inv + c)), __ldg(shift + c));
  }
}

__global__ void maxpool_kernel(const float* __restrict__ x, int n, int h, int w, int c, float* __restrict__ y) {
  int oh = h / 2, ow = w / 2;
  long long total = (long long)n * oh * ow * c;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += stride) {
    int cc = (int)(i % c);
    long long r = i / c;
    int ox = (int)(r % ow); r /= ow;
    int oy = (int)(r % oh);
    long long img = r / oh;
    const float* b = x + ((img * h + 2 * oy) * w + 2 * ox) * c + cc;
    float m = fmaxf(fmaxf(b[0], b[c]), fmaxf(b[(long long)w * c], b[(long long)w * c + c]));
    y[i] = m;
  }
}

__global__ void leaky_kernel(const float* __restrict__ x, long long count, float alpha, float* __restrict__ y) {
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += stride)
    y[i] = act_leaky(x[i], alpha);
}

__global__ void round_kernel(const float* __restrict__ x, long long count, float* __restrict__ y) {
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += stride) y[i] = rintf(x[i]);
}

__global__ void dequant_kernel(int kind, const void* __restrict__ x, long long count, int channels, float scale, float* __restrict__ y) {
  long long stride = (long long)gridDim.x * blockDim.x;
  const int words = (channels + 31) / 32;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += stride) {
    float v;
    if (kind == QNNB_KIND_I8) v = __fmul_rn((float)((const int8_t*)x)[i], scale);
    else if (kind == QNNB_KIND_U8) v = __fmul_rn((float)((const uint8_t*)x)[i], scale);
    else {
      long long r = i / channels;
      int c = (int)(i % channels);
      uint32_t wv = ((const uint32_t*)x)[r * words + (c >> 5)];
      v = ((wv >> (c & 31)) & 1u) ? 1.f : -1.f;
    }
    y[i] = v;
  }
}

int grid_for(long long total) {
  long long b = (total + 255) / 256;
  if (b > 148 * 16) b = 148 * 16;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace
}  // namespace qnnb

using namespace qnnb;

extern "C" {

int qnnb_quantize_act(int32_t act, int32_t abits, const float* x, int64_t count, int32_t channels, void* y, void* stream) {
  QNNB_CHECK_ARG(x && y && count >= 0, "quantize_act: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (count == 0) return QNNB_OK;
  if (act == QNNB_ACT_QUANT) {
    QNNB_CHECK_ARG(abits >= 2 && abits <= 8, "quantize_act: abits=%d outside 2..8", abits);
    quant_act_kernel<<<grid_for(count), 256, 0, st>>>(x, count, (float)(1 << (abits - 1)), (int8_t*)y);
  } else if (act == QNNB_ACT_SIGN) {
    QNNB_CHECK_ARG(channels > 0 && count % channels == 0, "quantize_act: count %% channels != 0");
    int words = (channels + 31) / 32;
    long long rows = count / channels;
    sign_act_kernel<<<grid_for(rows * words), 256, 0, st>>>(x, rows, channels, words, (uint32_t*)y);
  } else {
    set_error("quantize_act: act must be QUANT or SIGN");
    return QNNB_EINVAL;
  }
  QNNB_CUDA(cudaGetLastError());
  return QNNB_OK;
}

int qnnb_batchnorm_f32(const float* x, int64_t rows, int32_t ch, const float* inv, const float* shift, float* y, void* stream) {
  QNNB_CHECK_ARG(x && y && inv && shift && rows >= 0 && ch > 0, "batchnorm: bad arguments");
  if (rows == 0) return QNNB_OK;
  bn_kernel<<<grid_for(rows * ch), 256, 0, (cudaStream_t)stream>>>(x, rows * ch, ch, inv, shift, y);
  QNNB_CUDA(cudaGetLastError());
  return QNNB_OK;
}

int qnnb_maxpool2_f32(const float* x, int32_t n, int32_t h, int32_t w, int32_t c, float* y, void* stream) {
  QNNB_CHECK_ARG(x && y && n >= 0 && h >= 2 && w >= 2 && c > 0, "maxpool: bad arguments");
  if (n == 0) return QNNB_OK;
  maxpool_kernel<<<grid_for((long long)n * (h / 2) * (w / 2) * c), 256, 0, (cudaStream_t)stream>>>(x, n, h, w, c, y);
  QNNB_CUDA(cudaGetLastError());
  return QNNB_OK;
}

int qnnb_leaky_f32(const float* x, int64_t count, float alpha, float* y, void* stream) {
  QNNB_CHECK_ARG(x && y && count >= 0, "leaky: bad arguments");
  if (count == 0) return QNNB_OK;
  leaky_kernel<<<grid_for(count), 256, 0, (cudaStream_t)stream>>>(x, count, alpha, y);
  QNNB_CUDA(cudaGetLastError());
  return QNNB_OK;
}

int qnnb_round_f32(const float* x, int64_t count, float* y, void* stream) {
  QNNB_CHECK_ARG(x && y && count >= 0, "round: bad arguments");
  if (count == 0) return QNNB_OK;
  round_kernel<<<grid_for(count), 256, 0, (cudaStream_t)stream>>>(x, count, y);
  QNNB_CUDA(cudaGetLastError());
  return QNNB_OK;
}

int qnnb_dequantize(int32_t kind, const void* x, int64_t count, int32_t channels, float scale, float* y, void* stream) {
  QNNB_CHECK_ARG(x && y && count >= 0, "dequantize: bad arguments");
  QNNB_CHECK_ARG(kind == QNNB_KIND_I8 || kind == QNNB_KIND_U8 || kind == QNNB_KIND_B1, "dequantize: bad kind %d", kind);
  QNNB_CHECK_ARG(kind != QNNB_KIND_B1 || (channels > 0 && count % channels == 0), "dequantize: count %% channels != 0");
  if (count == 0) return QNNB_OK;
  dequant_kernel<<<grid_for(count), 256, 0, (cudaStream_t)stream>>>(kind, x, count, channels > 0 ? channels : 1, scale, y);
  QNNB_CUDA(cudaGetLastError());
  return QNNB_OK;
}

}  // extern "C"
