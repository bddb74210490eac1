// K0 -- weight quantisers + packers, run once at load time.
//
// Restates quantize (layers/quantized_ops.py:49-66), binarize (layers/binary_ops.py:54-64) and
// _ternarize (layers/ternary_ops.py:15-30), which the reference re-executes on every forward
// (layers/quantized_layers.py:80,165; binary_layers.py:79,161; ternary_layers.py:78,157).
// Input is the Keras HWIO fp32 kernel; output is K-major ([cout][kh][kw][cin]) so that the
// conv / dense kernels (and the TMA descriptors of the tcgen05 path) read contiguous K.
#include "common.cuh"

namespace qnnb {

// mean(|W/H|) in the fixed order the oracle uses: 1024 strided float64 partial sums (element i
// -> lane i % 1024, increasing i), halving tree, divide by count in float64, round to fp32.
// cutoff = fl32(0.7f * mean) is written to scratch[0].
__global__ void __launch_bounds__(1024, 1)
ternary_cutoff_kernel(const float* __restrict__ w, long long count, float H, float* __restrict__ scratch) {
  __shared__ double part[1024];
  const int t = threadIdx.x;
  double acc = 0.0;
  for (long long i = t; i < count; i += 1024) {
    float x = __fdiv_rn(w[i], H);
    acc = __dadd_rn(acc, (double)fabsf(x));
  }
  part[t] = acc;
  __syncthreads();
  for (int width = 512; width >= 1; width >>= 1) {
    if (t < width) part[t] = __dadd_rn(part[t], part[t + width]);
    __syncthreads();
  }
  if (t == 0) {
    float mean = (float)(part[0] / (double)count);
    scratch[0] = __fmul_rn(0.7f, mean);
  }
}

__device__ __forceinline__ int weight_level(int mode, float v, float qm, float H, float cutoff) {
  if (mode == QNNB_W_QUANT) {
    float q = rintf(__fmul_rn(v, qm));                  // tf.round: half-to-even
    q = fminf(fmaxf(q, -qm), qm - 1.f);
    return (int)q;
  } else if (mode == QNNB_W_BINARY) {
    // H * binary_tanh(W/H): +1 iff round(clip(0.5*x+0.5, 0, 1)) == 1
    float x = __fdiv_rn(v, H);
    float hs = fminf(fmaxf(__fadd_rn(__fmul_rn(0.5f, x), 0.5f), 0.f), 1.f);
    return rintf(hs) > 0.5f ? 1 : -1;
  } else {
    float x = __fdiv_rn(v, H);
    return x > cutoff ? 1 : (x <= -cutoff ? -1 : 0);
  }
}

// one thread per packed int8 element: out[co][tap][ci_pad]
__global__ void pack_i8_kernel(int mode, float qm, float H, const float* __restrict__ w,
                               int taps, int cin, int cin_pad, int cout,
                               int8_t* __restrict__ out, const float* __restrict__ scratch) {
  long long total = (long long)cout * taps * cin_pad;
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  int ci = (int)(i % cin_pad);
  long long r = i / cin_pad;
  int tap = (int)(r % taps);
  int co = (int)(r / taps);
  int lvl = 0;
  if (ci < cin) {
    float cutoff = (mode == QNNB_W_TERNARY) ? scratch[0] : 0.f;
    float v = w[((long long)tap * cin + ci) * cout + co];
    lvl = weight_level(mode, v, qm, H, cutoff);
  }
  out[i] = (int8_t)lvl;
}

// one thread per packed uint32 word: out[co][tap][word]
__global__ void pack_b1_kernel(int mode, float qm, float H, const float* __restrict__ w,
                               int taps, int cin, int words, int cout,
                               uint32_t* __restrict__ out, const float* __restrict__ scratch) {
  long long total = (long long)cout * taps * words;
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  int wd = (int)(i % words);
  long long r = i / words;
  int tap = (int)(r % taps);
  int co = (int)(r / taps);
  float cutoff = (mode == QNNB_W_TERNARY) ? scratch[0] : 0.f;
  uint32_t bits = 0;
  for (int b = 0; b < 32; ++b) {
    int ci = wd * 32 + b;
    if (ci < cin) {
      float v = w[((long long)tap * cin + ci) * cout + co];
      if (weight_level(mode, v, qm, H, cutoff) > 0) bits |= (1u << b);
    }
  }
  out[i] = bits;
}

// 'float' networks (plain Conv2D / Dense, models/model_factory.py:24-27): no quantiser, the kernel values re-laid out
// K-major like the level formats: out[co][tap][ci_pad], zero filled
__global__ void pack_f32_kernel(const float* __restrict__ w, int taps, int cin, int cin_pad, int cout, float* __restrict__ out) {
  long long total = (long long)cout * taps * cin_pad;
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  int ci = (int)(i % cin_pad);
  long long r = i / cin_pad;
  int tap = (int)(r % taps);
  int co = (int)(r / taps);
  out[i] = ci < cin ? w[((long long)tap * cin + ci) * cout + co] : 0.f;
}

int launch_pack_weights(int mode, int nb, float H, const float* w, int kh, int kw, int cin, int cout,
                        int wfmt, void* out, float* scratch, cudaStream_t st) {
  QNNB_CHECK_ARG(mode == QNNB_W_QUANT || mode == QNNB_W_BINARY || mode == QNNB_W_TERNARY || mode == QNNB_W_FLOAT, "pack_weights: bad mode %d", mode);
  QNNB_CHECK_ARG(wfmt == QNNB_WFMT_I8 || wfmt == QNNB_WFMT_B1 || wfmt == QNNB_WFMT_F32, "pack_weights: bad wfmt %d", wfmt);
  QNNB_CHECK_ARG((mode == QNNB_W_FLOAT) == (wfmt == QNNB_WFMT_F32), "pack_weights: QNNB_W_FLOAT and QNNB_WFMT_F32 come together");
  QNNB_CHECK_ARG(w && out, "pack_weights: null pointer");
  QNNB_CHECK_ARG(kh > 0 && kw > 0 && cin > 0 && cout > 0, "pack_weights: bad shape %dx%dx%dx%d", kh, kw, cin, cout);
  if (mode == QNNB_W_FLOAT) {
    const int cin_pad = (cin + 3) / 4 * 4;
    const long long total = (long long)cout * kh * kw * cin_pad;
    pack_f32_kernel<<<(int)((total + 255) / 256), 256, 0, st>>>(w, kh * kw, cin, cin_pad, cout, (float*)out);
    QNNB_CUDA(cudaGetLastError());
    return QNNB_OK;
  }
  QNNB_CHECK_ARG(mode != QNNB_W_QUANT || (nb >= 2 && nb <= 8), "pack_weights: nb=%d outside 2..8 (int8 levels)", nb);
  QNNB_CHECK_ARG(wfmt != QNNB_WFMT_B1 || mode == QNNB_W_BINARY, "pack_weights: 1-bit format needs binary weights");
  QNNB_CHECK_ARG(H > 0.f, "pack_weights: H must be > 0");
  QNNB_CHECK_ARG(mode != QNNB_W_TERNARY || scratch, "pack_weights: ternary mode needs scratch");
  const int taps = kh * kw;
  const float qm = (float)(1 << ((mode == QNNB_W_QUANT ? nb : 1) - 1));
  if (mode == QNNB_W_TERNARY) {
    ternary_cutoff_kernel<<<1, 1024, 0, st>>>(w, (long long)taps * cin * cout, H, scratch);
    QNNB_CUDA(cudaGetLastError());
  }
  if (wfmt == QNNB_WFMT_I8) {
    int cin_pad = (cin + 3) / 4 * 4;
    long long total = (long long)cout * taps * cin_pad;
    int blocks = (int)((total + 255) / 256);
    pack_i8_kernel<<<blocks, 256, 0, st>>>(mode, qm, H, w, taps, cin, cin_pad, cout, (int8_t*)out, scratch);
  } else {
    int words = (cin + 31) / 32;
    long long total = (long long)cout * taps * words;
    int blocks = (int)((total + 255) / 256);
    pack_b1_kernel<<<blocks, 256, 0, st>>>(mode, qm, H, w, taps, cin, words, cout, (uint32_t*)out, scratch);
  }
  QNNB_CUDA(cudaGetLastError());
  return QNNB_OK;
}

}  // namespace qnnb
