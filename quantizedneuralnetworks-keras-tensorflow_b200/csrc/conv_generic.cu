// Generic fused convolution on CUDA cores (dp4a / xor-popc / FFMA): every shape the layers
// accept (kernel <= 3x3, stride 1|2, any Cin/Cout, any H/W), all four input kinds.  It is the
// shape-complete path; conv_tc.cu (tcgen05) takes over for the tensor-core-friendly shapes.
//
// Stands in for K.conv2d + bias_add (+ the BatchNormalization / add / Activation / MaxPooling2D
// that follow it in models/vgg.py:15-37 and models/resnet.py:57-129) of
//   QuantizedConv2D.call layers/quantized_layers.py:164-194
//   BinaryConv2D.call    layers/binary_layers.py:160-187
//   TernaryConv2D.call   layers/ternary_layers.py:156-174
//
// Tiling: one CTA = TPY x TPX output pixels of one image x TC output channels; 256 threads, thread
// (ty, tx) owns the 2x2 pixel window ty and the 4 consecutive channels 4*tx..4*tx+3, so the
// 2x2 max-pool is an in-register reduction.  Three shapes keep all 256 threads busy on narrow
// layers: 8x8 px x 64 ch, 8x16 px x 32 ch (Cout <= 32), 16x16 px x 16 ch (Cout <= 16).  K is streamed through shared memory in chunks of
// CKW words per tap (a word = 4 int8 channels, 32 binary channels or 1 float channel).
// Because every epilogue step is monotone in the accumulator for a fixed channel, pooling is
// done on the raw accumulators (max, or min when the BN slope is negative) BEFORE the fp32
// epilogue -- bit-identical to pooling the activations.
#include "common.cuh"

namespace qnnb {

namespace {

constexpr int CKW = 8;
constexpr int MAXK = 3;

// tile shape for a channel tile of TC: (TC / 4) channel quads x (256 / (TC / 4)) pixel windows
template <int TC> struct Tile {
  static constexpr int NQ = TC / 4;                         // threads along the channel axis
  static constexpr int WIN = 256 / NQ;                      // 2x2 pixel windows per CTA
  static constexpr int WX = (TC == 64) ? 4 : 8;             // windows per row
  static constexpr int WY = WIN / WX;
  static constexpr int TPX = 2 * WX, TPY = 2 * WY;          // output pixels per CTA
  static constexpr int MAX_IX = (TPX - 1) * 2 + MAXK;       // halo extent under stride 2
  static constexpr int MAX_IY = (TPY - 1) * 2 + MAXK;
};

struct ConvP {
  int n, h, w, cin, cout, kh, kw, stride, pad_t, pad_l, oh, ow;
  int tiles_y, tiles_x;
  int kwords;        // K words per tap (int8: cin_pad/4, b1: ceil(cin/32), f32: cin)
  int cin_pad;       // int8 weight row pitch in bytes
  int w_f32;         // KIND_F32: the packed kernel holds fp32 values (QNNB_WFMT_F32, row pitch cin_pad floats) instead of int8 levels
  const void* x;
  const void* wts;
  void* y;
  Epi epi;
};

__device__ __forceinline__ int dp4a_us(unsigned a, int b, int c) {
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

template <int KIND> struct AccT { typedef int type; };
template <> struct AccT<QNNB_KIND_F32> { typedef float type; };

template <int KIND, int TC>
__global__ void __launch_bounds__(256)
conv_generic_kernel(const ConvP p) {
  typedef typename AccT<KIND>::type acc_t;
  typedef Tile<TC> T;
  constexpr int TPX = T::TPX, TPY = T::TPY;
  if (threadIdx.x == 0) griddep_launch_dependents();
  griddep_wait();
  __shared__ __align__(16) uint32_t sa[T::MAX_IY * T::MAX_IX * CKW];
  __shared__ __align__(16) uint32_t sw[MAXK * MAXK * CKW * TC];

  const int tid = threadIdx.x;
  const int tx = tid % T::NQ;
  const int ty = tid / T::NQ;
  const int wy = ty / T::WX, wx = ty % T::WX;

  int tile = blockIdx.x;
  const int tile_x = tile % p.tiles_x; tile /= p.tiles_x;
  const int tile_y = tile % p.tiles_y; tile /= p.tiles_y;
  const int img = tile;
  const int oy0 = tile_y * TPY, ox0 = tile_x * TPX;
  const int c_base = blockIdx.y * TC;
  const int taps = p.kh * p.kw;

  // halo origin / extent in input coordinates
  const int iy0 = oy0 * p.stride - p.pad_t;
  const int ix0 = ox0 * p.stride - p.pad_l;
  const int IH = (TPY - 1) * p.stride + p.kh;
  const int IW = (TPX - 1) * p.stride + p.kw;

  acc_t acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0;

  // B1: which (pixel, tap) pairs read a real input pixel (zero padding contributes 0, not -1)
  unsigned long long vmask = 0ull;
  int nvalid[4] = {0, 0, 0, 0};
  if constexpr (KIND == QNNB_KIND_B1) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int oy = oy0 + 2 * wy + (q >> 1), ox = ox0 + 2 * wx + (q & 1);
      for (int t = 0; t < taps; ++t) {
        int iy = oy * p.stride - p.pad_t + t / p.kw;
        int ix = ox * p.stride - p.pad_l + t % p.kw;
        if (iy >= 0 && iy < p.h && ix >= 0 && ix < p.w) { vmask |= 1ull << (q * 9 + t); nvalid[q]++; }
      }
    }
  }

  for (int k0 = 0; k0 < p.kwords; k0 += CKW) {
    const int kc = min(CKW, p.kwords - k0);
    __syncthreads();
    // ---- stage the input halo: sa[(iy*IW + ix)*CKW + k]
    // Fast path (full chunk, 16-byte aligned rows): one item = one halo pixel = its 8 K words as two 128-bit loads
    // (independent, two items in flight per thread) instead of eight scalar loads with per-word index arithmetic.
    bool a_vec;
    if constexpr (KIND == QNNB_KIND_F32) a_vec = (kc == CKW) && (p.cin & 3) == 0;
    else if constexpr (KIND == QNNB_KIND_B1) a_vec = (kc == CKW) && (p.kwords & 3) == 0;
    else a_vec = (kc == CKW) && (p.cin & 15) == 0;
    if (a_vec) {
      const int npix = IH * IW;
#pragma unroll 2
      for (int pix = tid; pix < npix; pix += 256) {
        const int iy = pix / IW, ix = pix - iy * IW;
        const int gy = iy0 + iy, gx = ix0 + ix;
        uint4 v0 = make_uint4(0, 0, 0, 0), v1 = v0;
        if (gy >= 0 && gy < p.h && gx >= 0 && gx < p.w) {
          const long long pixel = ((long long)img * p.h + gy) * p.w + gx;
          const uint4* src;
          if constexpr (KIND == QNNB_KIND_F32) src = reinterpret_cast<const uint4*>((const float*)p.x + pixel * p.cin + k0);
          else if constexpr (KIND == QNNB_KIND_B1) src = reinterpret_cast<const uint4*>((const uint32_t*)p.x + pixel * p.kwords + k0);
          else src = reinterpret_cast<const uint4*>((const uint8_t*)p.x + pixel * p.cin + k0 * 4);
          v0 = __ldg(src);
          v1 = __ldg(src + 1);
        }
        *reinterpret_cast<uint4*>(&sa[pix * CKW]) = v0;
        *reinterpret_cast<uint4*>(&sa[pix * CKW + 4]) = v1;
      }
    } else
    for (int i = tid; i < IH * IW * CKW; i += 256) {
      int k = i % CKW;
      int pix = i / CKW;
      int ix = pix % IW, iy = pix / IW;
      int gy = iy0 + iy, gx = ix0 + ix;
      uint32_t v = 0;
      if (k < kc && gy >= 0 && gy < p.h && gx >= 0 && gx < p.w) {
        long long pixel = ((long long)img * p.h + gy) * p.w + gx;
        if constexpr (KIND == QNNB_KIND_F32) {
          v = __float_as_uint(__ldg((const float*)p.x + pixel * p.cin + (k0 + k)));
        } else if constexpr (KIND == QNNB_KIND_B1) {
          v = __ldg((const uint32_t*)p.x + pixel * p.kwords + (k0 + k));
        } else {
          const uint8_t* src = (const uint8_t*)p.x + pixel * p.cin + (k0 + k) * 4;
          if ((p.cin & 3) == 0) {
            v = __ldg((const uint32_t*)src);
          } else {
            int c = (k0 + k) * 4;
#pragma unroll
            for (int b = 0; b < 4; ++b)
              if (c + b < p.cin) v |= (uint32_t)__ldg(src + b) << (8 * b);
          }
        }
      }
      sa[i] = v;
    }
    // ---- stage the weights: sw[(tap*CKW + k)*TC + c]
    // Fast path: one item = (tap, channel) = 8 consecutive K words of the packed kernel as vector loads
    bool w_vec;
    if constexpr (KIND == QNNB_KIND_F32) w_vec = (kc == CKW) && (p.cin_pad & 7) == 0;
    else w_vec = (kc == CKW) && (p.kwords & 3) == 0;
    if (w_vec) {
#pragma unroll 2
      for (int item = tid; item < taps * TC; item += 256) {
        const int c = item % TC, t = item / TC;
        const int co = c_base + c;
        uint32_t wd[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (co < p.cout) {
          if constexpr (KIND == QNNB_KIND_F32) {
            if (p.w_f32) {
              const uint4* src = reinterpret_cast<const uint4*>((const float*)p.wts + ((long long)co * taps + t) * p.cin_pad + k0);
              const uint4 a = __ldg(src), b = __ldg(src + 1);
              wd[0] = a.x; wd[1] = a.y; wd[2] = a.z; wd[3] = a.w; wd[4] = b.x; wd[5] = b.y; wd[6] = b.z; wd[7] = b.w;
            } else {
              const uint2 raw = __ldg(reinterpret_cast<const uint2*>((const int8_t*)p.wts + ((long long)co * taps + t) * p.cin_pad + k0));
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const uint32_t word = k < 4 ? raw.x : raw.y;
                wd[k] = __float_as_uint((float)(int)(int8_t)(word >> (8 * (k & 3))));
              }
            }
          } else {
            const uint4* src = reinterpret_cast<const uint4*>((const uint32_t*)p.wts + ((long long)co * taps + t) * p.kwords + k0);
            const uint4 a = __ldg(src), b = __ldg(src + 1);
            wd[0] = a.x; wd[1] = a.y; wd[2] = a.z; wd[3] = a.w; wd[4] = b.x; wd[5] = b.y; wd[6] = b.z; wd[7] = b.w;
          }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) sw[(t * CKW + k) * TC + c] = wd[k];
      }
    } else
    for (int i = tid; i < taps * CKW * TC; i += 256) {
      int k = i % CKW;
      int r = i / CKW;
      int t = r % taps;
      int c = r / taps;
      int co = c_base + c;
      uint32_t v = 0;
      if (k < kc && co < p.cout) {
        if constexpr (KIND == QNNB_KIND_F32) {
          const long long wi = ((long long)co * taps + t) * p.cin_pad + (k0 + k);
          v = p.w_f32 ? __float_as_uint(__ldg((const float*)p.wts + wi)) : __float_as_uint((float)__ldg((const int8_t*)p.wts + wi));
        } else {
          v = __ldg((const uint32_t*)p.wts + ((long long)co * taps + t) * p.kwords + (k0 + k));
        }
      }
      sw[(t * CKW + k) * TC + c] = v;
    }
    __syncthreads();
    // ---- multiply-accumulate
    for (int t = 0; t < taps; ++t) {
      const int r = t / p.kw, s = t % p.kw;
      int aoff[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        int iy = (2 * wy + (q >> 1)) * p.stride + r;
        int ix = (2 * wx + (q & 1)) * p.stride + s;
        aoff[q] = (iy * IW + ix) * CKW;
      }
      for (int k = 0; k < kc; ++k) {
        const uint4 wv = *reinterpret_cast<const uint4*>(&sw[(t * CKW + k) * TC + tx * 4]);
        const uint32_t wj[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t av = sa[aoff[q] + k];
          if constexpr (KIND == QNNB_KIND_B1) {
            if ((vmask >> (q * 9 + t)) & 1ull) {
#pragma unroll
              for (int j = 0; j < 4; ++j) acc[q][j] += __popc(av ^ wj[j]);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if constexpr (KIND == QNNB_KIND_F32) acc[q][j] = fmaf(__uint_as_float(av), __uint_as_float(wj[j]), acc[q][j]);
              else if constexpr (KIND == QNNB_KIND_U8) acc[q][j] = dp4a_us(av, (int)wj[j], acc[q][j]);
              else acc[q][j] = __dp4a((int)av, (int)wj[j], acc[q][j]);
            }
          }
        }
      }
    }
  }
  if constexpr (KIND == QNNB_KIND_B1) {
    // sum of +-1 products over the valid taps = valid*cin - 2*popc(xor)
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[q][j] = nvalid[q] * p.cin - 2 * acc[q][j];
  }

  // ------------------------------------------------------------------ fused epilogue
  const Epi& e = p.epi;
  const int ch0 = c_base + tx * 4;
  ChanConst cc[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) cc[j] = load_chan(e, ch0 + j, ch0 + j < p.cout);

  const int npix = e.pool ? 1 : 4;
  const int out_h = e.pool ? p.oh / 2 : p.oh;
  const int out_w = e.pool ? p.ow / 2 : p.ow;
  if (e.pool) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      acc_t mx = acc[0][j], mn = acc[0][j];
#pragma unroll
      for (int q = 1; q < 4; ++q) { mx = max(mx, acc[q][j]); mn = min(mn, acc[q][j]); }
      acc[0][j] = decreasing(cc[j]) ? mn : mx;
    }
  }
  for (int q = 0; q < npix; ++q) {
    int oy, ox;
    if (e.pool) { oy = (oy0 >> 1) + wy; ox = (ox0 >> 1) + wx; }
    else { oy = oy0 + 2 * wy + (q >> 1); ox = ox0 + 2 * wx + (q & 1); }
    const bool pvalid = (oy < out_h) && (ox < out_w);
    const long long opix = ((long long)img * out_h + oy) * out_w + ox;
    float z[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v = affine((float)acc[q][j], cc[j]);      // int -> float is cvt.rn
      if (e.res_kind != QNNB_KIND_NONE && pvalid && ch0 + j < p.cout) {
        float sc;
        if (e.res_kind == QNNB_KIND_I8) sc = __fmul_rn((float)__ldg((const int8_t*)e.residual + opix * p.cout + ch0 + j), e.res_scale);
        else sc = __ldg((const float*)e.residual + opix * p.cout + ch0 + j);
        v = add_residual(v, sc, e.res_mul);
      }
      z[j] = v;
    }
    if (e.act == QNNB_ACT_QUANT || e.act == QNNB_ACT_SIGN_I8) {
      uint32_t packed = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int lvl = (e.act == QNNB_ACT_QUANT) ? act_quant(z[j], e.qm) : (act_sign(z[j]) ? 1 : -1);
        packed |= ((uint32_t)(lvl & 0xff)) << (8 * j);
      }
      if (pvalid) {
        int8_t* dst = (int8_t*)p.y + opix * p.cout + ch0;
        if ((p.cout & 3) == 0) { if (ch0 < p.cout) *reinterpret_cast<uint32_t*>(dst) = packed; }
        else {
#pragma unroll
          for (int j = 0; j < 4; ++j) if (ch0 + j < p.cout) dst[j] = (int8_t)(packed >> (8 * j));
        }
      }
    } else if (e.act == QNNB_ACT_SIGN) {
      uint32_t nib = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) if (ch0 + j < p.cout && act_sign(z[j])) nib |= 1u << j;
      uint32_t word = nib << (4 * (tx & 7));
      word |= __shfl_xor_sync(0xffffffffu, word, 1);
      word |= __shfl_xor_sync(0xffffffffu, word, 2);
      word |= __shfl_xor_sync(0xffffffffu, word, 4);
      const int cwords = (p.cout + 31) / 32;
      const int widx = blockIdx.y * (TC / 32) + (tx >> 3);
      if (pvalid && (tx & 7) == 0 && widx < cwords) ((uint32_t*)p.y)[opix * cwords + widx] = word;
    } else {
      if (e.act == QNNB_ACT_LEAKY) {
#pragma unroll
        for (int j = 0; j < 4; ++j) z[j] = act_leaky(z[j], e.leaky_alpha);
      }
      if (pvalid) {
        float* dst = (float*)p.y + opix * p.cout + ch0;
        if ((p.cout & 3) == 0) { if (ch0 < p.cout) *reinterpret_cast<float4*>(dst) = make_float4(z[0], z[1], z[2], z[3]); }
        else {
#pragma unroll
          for (int j = 0; j < 4; ++j) if (ch0 + j < p.cout) dst[j] = z[j];
        }
      }
    }
  }
}

}  // namespace

int launch_conv_generic(const qnnb_conv_desc& d, const void* x, const void* w, void* y, cudaStream_t st) {
  ConvP p;
  p.n = d.n; p.h = d.h; p.w = d.w; p.cin = d.cin; p.cout = d.cout; p.kh = d.kh; p.kw = d.kw; p.stride = d.stride;
  same_pad(d.h, d.kh, d.stride, &p.oh, &p.pad_t);
  same_pad(d.w, d.kw, d.stride, &p.ow, &p.pad_l);
  // channel tile: the widest one that does not leave most threads idle (the bit-packing epilogue needs >= 32 channels)
  const int tc = (d.cout <= 16 && d.epi.act != QNNB_ACT_SIGN) ? 16 : (d.cout <= 32 ? 32 : 64);
  p.tiles_y = ceil_div(p.oh, tc == 16 ? Tile<16>::TPY : (tc == 32 ? Tile<32>::TPY : Tile<64>::TPY));
  p.tiles_x = ceil_div(p.ow, tc == 16 ? Tile<16>::TPX : (tc == 32 ? Tile<32>::TPX : Tile<64>::TPX));
  p.cin_pad = (d.cin + 3) / 4 * 4;
  p.w_f32 = d.w_f32;
  if (d.in_kind == QNNB_KIND_F32) p.kwords = d.cin;
  else if (d.in_kind == QNNB_KIND_B1) p.kwords = (d.cin + 31) / 32;
  else p.kwords = p.cin_pad / 4;
  p.x = x; p.wts = w; p.y = y;
  p.epi = make_epi(d.epi);
  long long gx = (long long)d.n * p.tiles_y * p.tiles_x;
  QNNB_CHECK_ARG(gx > 0 && gx < 2147483647LL, "conv2d: grid too large");
  dim3 grid((unsigned)gx, (unsigned)ceil_div(d.cout, tc));
#define QNNB_GENERIC_LAUNCH(KIND)                                                          \
  do {                                                                                     \
    cudaError_t e__;                                                                       \
    if (tc == 16) e__ = launch_pdl(conv_generic_kernel<KIND, 16>, grid, dim3(256), (size_t)0, st, p);      \
    else if (tc == 32) e__ = launch_pdl(conv_generic_kernel<KIND, 32>, grid, dim3(256), (size_t)0, st, p); \
    else e__ = launch_pdl(conv_generic_kernel<KIND, 64>, grid, dim3(256), (size_t)0, st, p);               \
    if (e__ != cudaSuccess) return cuda_fail(e__, "conv_generic launch");                  \
  } while (0)
  switch (d.in_kind) {
    case QNNB_KIND_U8: QNNB_GENERIC_LAUNCH(QNNB_KIND_U8); break;
    case QNNB_KIND_I8: QNNB_GENERIC_LAUNCH(QNNB_KIND_I8); break;
    case QNNB_KIND_B1: QNNB_GENERIC_LAUNCH(QNNB_KIND_B1); break;
    case QNNB_KIND_F32: QNNB_GENERIC_LAUNCH(QNNB_KIND_F32); break;
    default: set_error("conv2d: bad in_kind %d", d.in_kind); return QNNB_EINVAL;
  }
#undef QNNB_GENERIC_LAUNCH
  QNNB_CUDA(cudaGetLastError());
  return QNNB_OK;
}

}  // namespace qnnb
