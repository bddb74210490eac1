// Kf -- whole-network kernel for the small VGG nets: ONE launch runs every layer of models/vgg.py:15-42
//   [QuantizedConv2D | BinaryConv2D] -> BatchNormalization -> Activation [-> MaxPooling2D] ... -> Flatten -> Fc -> BN
// (layers/quantized_layers.py:164-194, layers/binary_layers.py:160-187 for the convolutions; :79-88 / :78-85 for the
// dense head) for the configurations whose packed kernels fit in one SM's shared memory (<= 64 filters per layer:
// BASELINE config 1, the MNIST 28x28x1 64/64/64 net, and the CIFAR-10 64/64/64 variant of config/config_CIFAR-10.py).
//
// Why: as separate layer kernels such a net is pure launch latency and fill / drain -- cfg1 took 61 us for 100 images
// (four launches of ~15 us on 100 x 9.5 M MACs), 0.5 % of any roofline.  Here one persistent CTA takes an IMAGE through
// the whole net: the activations never leave the SM, the kernels of all layers stay resident as UMMA A operands, and
// the only global traffic is the image in (one bulk copy) and `units` floats out.
//
//   * Activations live in shared memory in the un-swizzled K-major core-matrix layout, one "raster" per layer:
//     [channel / 16][pixel of the zero-haloed map, pitch W + 2][16 channels] -- pixel p of a plane sits at byte 16 p,
//     which is exactly the row pitch of an 8 x 16 B core matrix, so the B operand of filter tap (r, s) is THE SAME
//     raster seen through a descriptor whose start address is shifted by (r * pitch + s) pixels (no im2col, any map
//     width; the two halo columns between image rows produce two garbage accumulator columns per row, never read).
//   * conv l >= 1: per block of image rows 9 taps x Cin/32 tcgen05.mma.kind::i8 (M 128 of which `cout` rows are kernel
//     rows, N = rows x pitch <= 256, K 32), int32 accumulators in TMEM, two 256-column buffers (MMA of block b + 1
//     overlaps the epilogue of block b).  First layer (Cin 1 or 3, K = 9 Cin <= 27): an explicit 32-byte im2col row
//     per pixel, ONE MMA per block, u8 x s8.
//   * Epilogue: thread = output channel (TMEM lane), 8 warps (the lane quarters that hold the <= 64 kernel rows), four
//     warp pairs take alternate image-row pairs of a block.  2x2 max-pool on the raw accumulators (min where the BN slope
//     is negative), then the fixed fp32 op order of common.cuh (qaffine / quant_scaled, or the binary_tanh threshold),
//     and the level byte goes straight into the NEXT layer's raster (or the flat HWC feature vector of the dense head).
//   * Pooled first layer (the bulk of the accumulators: 50 k of cfg1's 66 k per image): operands swapped -- M = 128 POOLED
//     pixels, N = the kernel's 64 rows, and one MMA per 2x2 window position into its own 64 accumulator columns (the
//     im2col rows are written position-major).  Thread = pooled pixel: all four SM sub-partitions work, the pool is a
//     max over four columns of the same lane, a warp handles 16 channels and its 16 level bytes leave as ONE 16-byte
//     store into the next raster (per-channel constants are broadcast shared-memory loads).
//   * Warp roles (544 threads): warp 16 -> one elected MMA-issuing thread; warps 0..15: (id % 4) < 2 -> epilogue of
//     layers 1.., the others -> image fetch (bulk copy, double buffered), first-layer im2col of the NEXT image while
//     this one is in flight, and the dense head (dp4a + warp reduce, fp32 affine) of the PREVIOUS image; all sixteen
//     share the pooled first layer's epilogue.
#include "common.cuh"
#include "tc_ptx.cuh"

#include <string.h>

namespace qnnb {

namespace {

using namespace tcx;

constexpr int NF_THREADS = 544;           // 16 epilogue / worker warps + the MMA-issuing warp 16
constexpr int NF_MAXL = QNNB_NET_MAX_CONVS;
constexpr int NF_ABLK = 2048;            // one A block: 64 kernel rows x 32 K bytes (the MMA reads 128 rows: the next block)
constexpr int NF_WORKERS = 8;            // worker warps: ids with (id % 4) >= 2 among the first 16
constexpr int NF_SMEM_MAX = 232448;
constexpr int NF_MMA_WARP = 16;

struct NfLayer {
  int h, w, cin, cout;
  int wp;                  // input pitch in pixels (layer 0: w, the im2col rows; else w + 2)
  int rb, nblk;            // image rows per block, blocks
  int pool, sign;
  int in_off, in_plane;    // bytes: operand base, distance between 16-channel planes (K chunks)
  int out_off, out_plane, out_wp;      // next raster; out_plane == 0: flat [pixel][cout] feature vector
  int a_off;               // A blocks of this layer
  int cin_pad;             // channel pitch of the packed kernel
  int posmajor;            // pooled first layer: pixels on M, one MMA per 2x2 window position (see nf_first_pooled)
  int ppw, npp, mtiles;    // posmajor: pooled width, pooled pixels, 128-row M tiles
  float qm, acc_scale;
  const int8_t* wpk;
  const float *bias, *bn_inv, *bn_shift;
};

struct NfParams {
  const uint8_t* x;
  float* y;
  int n, nconv;
  int img_bytes, raw_off, raw_stride;
  int feat_off, feat_stride, fin;
  int dw_off, units;
  int cst_off, dcst_off;
  int blob_bytes;          // resident image: [A blocks | dense kernel | constants] at shared-memory offset 0
  int zero_off, zero_bytes;
  int bar_off;
  const uint8_t* blob;
  const int8_t* dense_w;
  unsigned long long* tr;  // TRACE builds: event buffer (qnnb_debug_set_trace), else NULL
  Epi dense_epi;
  NfLayer L[NF_MAXL];
};

// device-side timeline of CTA 0 (make TRACE=1 + qnnb_debug_set_trace; tools/net_trace.py): per-warp event regions
#ifdef QNNB_TRACE
#define nftrace(p, tag, idx)                                                                         \
  do {                                                                                               \
    if ((p).tr != nullptr && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && trk__ < 510) {            \
      const unsigned long long t__ = (unsigned long long)clock64();                                  \
      *reinterpret_cast<ulonglong2*>((p).tr + (threadIdx.x >> 5) * 1024 + 2 + 2 * trk__) =           \
          make_ulonglong2(((unsigned long long)(tag) << 32) | (unsigned)(idx), t__);                 \
      ++trk__;                                                                                       \
    }                                                                                                \
  } while (0)
#else
#define nftrace(p, tag, idx) do { } while (0)
#endif

__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, int* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait_dep8x2(int (&a)[8], int (&b)[8]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]), "+r"(b[0]),
                 "+r"(b[1]), "+r"(b[2]), "+r"(b[3]), "+r"(b[4]), "+r"(b[5]), "+r"(b[6]), "+r"(b[7])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait_dep8(int (&a)[8]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7])
               :
               : "memory");
}

// level byte of one accumulator: the fixed pipeline in its pre-scaled form (common.cuh make_qconst<false> / qaffine)
template <bool SIGN>
__device__ __forceinline__ int nf_level(int acc, const QConst& q, float qm) {
  const float zq = qaffine<false>(acc, q);
  if (SIGN) return act_sign(zq) ? 1 : -1;          // qm == 1 for sign layers: zq == z
  return quant_scaled(zq, qm);
}

// Epilogue of one accumulator block for one warp: rows (row pairs when pooling) u = pair, pair + 4, ... of the block.
// taddr: TMEM address of the block's column 0 in this warp's lane quarter; the output pixel (oh, ow) of this thread's
// channel lives at obase + oh * orow + ow * opix (next layer's raster, or the flat feature vector).  TMEM loads run one
// chunk of 8 columns ahead of the arithmetic.  DEC: some lane of the warp has a decreasing channel (BN slope < 0): the
// pool is then a min for those lanes, computed as ~max(~x).
#ifdef QNNB_TRACE
#define NF_TRARGS , const NfParams& p, int& trk__
#define NF_TRPASS , p, trk__
#else
#define NF_TRARGS
#define NF_TRPASS
#endif

// Epilogue of one accumulator block for one warp: rows (row pairs when pooling) u = pair, pair + 4, ... of the block.
// taddr: TMEM address of the block's column 0 in this warp's lane quarter; the output pixel (oh, ow) of this thread's
// channel lives at obase + oh * orow + ow * opix (next layer's raster: OPIX = 16; the flat feature vector: OPIX = 0,
// run-time pitch).  TMEM loads run one chunk of 8 columns ahead of the arithmetic (issued as soon as the pool has
// consumed the registers).  DEC: some lane of the warp has a decreasing channel (BN slope < 0): the pool is then a
// min for those lanes, ~max(~x).
template <bool POOL, bool SIGN, bool DEC, int OPIX>
__device__ __forceinline__ void nf_epi_block(uint32_t taddr, int pair, int rows, int h0, int w, int wp, uint8_t* obase, int orow,
                                             int opix, const QConst& qc, float qm, int flip NF_TRARGS) {
  const int nch = (w + 7) >> 3;
  const int step = OPIX ? OPIX : opix;
  if (POOL) {
    const int units = rows >> 1;
    const int ow = w >> 1;
    for (int u = pair; u < units; u += 4) {
      uint8_t* o = obase + ((h0 >> 1) + u) * orow;
      const uint32_t ta = taddr + (uint32_t)(2 * u * wp);
      int a[8], b[8];
      __syncwarp();
      tmem_ld8_nowait(ta, a);
      tmem_ld8_nowait(ta + wp, b);
      for (int ch = 0; ch < nch; ++ch, o += 4 * step) {
        tmem_ld_wait_dep8x2(a, b);
        nftrace(p, 22, ch);
        int m[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (DEC) m[i] = max(max(a[2 * i] ^ flip, a[2 * i + 1] ^ flip), max(b[2 * i] ^ flip, b[2 * i + 1] ^ flip)) ^ flip;
          else m[i] = max(max(a[2 * i], a[2 * i + 1]), max(b[2 * i], b[2 * i + 1]));
        }
        if (ch + 1 < nch) {                       // warp-uniform
          __syncwarp();
          tmem_ld8_nowait(ta + 8 * (ch + 1), a);
          tmem_ld8_nowait(ta + wp + 8 * (ch + 1), b);
        }
        int lv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) lv[i] = nf_level<SIGN>(m[i], qc, qm);
        const int lim = ow - 4 * ch;              // outputs still inside the map (< 4 only in a row's last chunk)
        if (lim >= 4) {
#pragma unroll
          for (int i = 0; i < 4; ++i) o[OPIX ? i * OPIX : i * opix] = (uint8_t)lv[i];
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (i < lim) o[OPIX ? i * OPIX : i * opix] = (uint8_t)lv[i];
        }
        nftrace(p, 23, ch);
      }
    }
  } else {
    for (int u = pair; u < rows; u += 4) {
      uint8_t* o = obase + (h0 + u) * orow;
      const uint32_t ta = taddr + (uint32_t)(u * wp);
      int a[8];
      __syncwarp();
      tmem_ld8_nowait(ta, a);
      for (int ch = 0; ch < nch; ++ch, o += 8 * step) {
        tmem_ld_wait_dep8(a);
        int m[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) m[i] = a[i];
        if (ch + 1 < nch) {
          __syncwarp();
          tmem_ld8_nowait(ta + 8 * (ch + 1), a);
        }
        int lv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) lv[i] = nf_level<SIGN>(m[i], qc, qm);
        const int lim = w - 8 * ch;
        if (lim >= 8) {
#pragma unroll
          for (int i = 0; i < 8; ++i) o[OPIX ? i * OPIX : i * opix] = (uint8_t)lv[i];
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (i < lim) o[OPIX ? i * OPIX : i * opix] = (uint8_t)lv[i];
        }
      }
    }
  }
}

// Pooled first layer with pixels on M (see the file header): this warp's lane quarter (32 pooled pixels of each M tile) x
// the 16 channels of group g = warp / 4.  Accumulator columns of M tile mt: [mt * 256 + pos * 64 + channel], pos = the 2x2
// window position.  cst: the layer's per-channel constants {s, bias, inv * qm, shift * qm} (broadcast loads); the 16 level
// bytes of a pooled pixel leave as one 16-byte store: next raster plane g, or the flat [pixel][cout] feature vector.
template <bool SIGN>
__device__ __forceinline__ void nf_first_pooled(uint32_t tmem_base, int warp, int lane, int mtiles, int npp, int ppw, int cout, float qm,
                                                const float4* cst, uint8_t* out, int out_plane, int out_wp) {
  const int quarter = warp & 3, g = warp >> 2;
  if (16 * g >= cout) return;                                // warp-uniform
  const uint32_t rcp = (65536u + (uint32_t)ppw - 1u) / (uint32_t)ppw;   // pp / ppw for pp < 2048, ppw <= 16
  for (int mt = 0; mt < mtiles; ++mt) {
    const int pp = mt * 128 + quarter * 32 + lane;
    const uint32_t ta = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(mt * 256 + 16 * g);
    int p0[16], p1[16], p2[16], p3[16];
    __syncwarp();
    tmem_ld16_nowait(ta, p0);
    tmem_ld16_nowait(ta + 64, p1);
    tmem_ld16_nowait(ta + 128, p2);
    tmem_ld16_nowait(ta + 192, p3);
    tmem_ld_wait_dep16x4(p0, p1, p2, p3);
    uint32_t wd[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float4 kc = cst[16 * g + j];
      QConst qc;
      qc.s = kc.x; qc.a = kc.y; qc.b = kc.z; qc.c = kc.w;
      const bool dec = kc.z < 0.f;                           // uniform: one channel per j for the whole warp
      const int m = dec ? min(min(p0[j], p1[j]), min(p2[j], p3[j])) : max(max(p0[j], p1[j]), max(p2[j], p3[j]));
      const int lv = nf_level<SIGN>(m, qc, qm);
      wd[j >> 2] |= (uint32_t)(lv & 0xFF) << (8 * (j & 3));
    }
    if (pp < npp) {
      uint8_t* dst;
      if (out_plane == 0) {
        dst = out + pp * cout + 16 * g;
      } else {
        const int ph = (int)(((uint32_t)pp * rcp) >> 16), pw = pp - ph * ppw;
        dst = out + g * out_plane + (((ph + 1) * out_wp + pw + 1) << 4);
      }
      *reinterpret_cast<uint4*>(dst) = make_uint4(wd[0], wd[1], wd[2], wd[3]);
    }
  }
}

// first-layer im2col: pixel (h, w) -> 32 K bytes (tap-major, channel-minor; zero outside the image) at row h * WP + w of
// the two 16-byte K planes.  The 3 * CIN bytes a filter row needs are contiguous in the raw image: they are cut out of
// aligned 32-bit words with funnel shifts (byte loads cost ~20 instructions per tap).
// POS (pooled first layer with pixels on M): four tiles, one per 2x2 window position, row = pooled pixel; tile = 2 planes.
template <int CIN, bool POS>
__device__ __forceinline__ void nf_im2col(const uint8_t* raw, uint8_t* dst, int plane, int H, int W, int WP, int wt, int nthreads) {
  const int npix = H * W;
  const uint32_t rcp = (65536u + (uint32_t)W - 1u) / (uint32_t)W;       // n / W == (n * rcp) >> 16 for n < 2048, W <= 32
  const uint32_t* words = reinterpret_cast<const uint32_t*>(raw);       // raw is 16-byte aligned
  for (int n = wt; n < npix; n += nthreads) {
    const int h = (int)(((uint32_t)n * rcp) >> 16), w = n - h * W;
    if (POS && (h >= (H & ~1) || w >= (W & ~1))) continue;     // the odd last row / column is dropped by the valid pooling
    const bool left = w == 0, right = w == W - 1;
    uint32_t v[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int ih = h + r - 1;
      const bool rok = ih >= 0 && ih < H;
      const int o = rok ? (ih * W + w - 1) * CIN : 0;                    // byte offset of the row's first tap (may be -CIN at w = 0)
      const int wa = o >> 2;                                             // arithmetic shift: floor
      const uint32_t sh = (uint32_t)(o & 3) * 8u;
      const uint32_t x0 = words[wa], x1 = words[wa + 1];
      if (CIN == 1) {
        uint32_t t = __funnelshift_r(x0, x1, sh) & 0x00FFFFFFu;
        if (left) t &= 0x00FFFF00u;
        if (right) t &= 0x0000FFFFu;
        v[r][0] = rok ? t : 0u;
        v[r][1] = v[r][2] = 0u;
      } else {
        const uint32_t x2 = words[wa + 2], x3 = words[wa + 3];
        uint32_t t0 = __funnelshift_r(x0, x1, sh), t1 = __funnelshift_r(x1, x2, sh), t2 = __funnelshift_r(x2, x3, sh) & 0xFFu;
        if (left) t0 &= 0xFF000000u;                                     // bytes 0..2: the pixel left of the image
        if (right) { t1 &= 0x0000FFFFu; t2 = 0u; }                       // bytes 6..8: the pixel right of the image
        v[r][0] = rok ? t0 : 0u; v[r][1] = rok ? t1 : 0u; v[r][2] = rok ? t2 : 0u;
      }
    }
    uint32_t wd[8];
    if (CIN == 1) {                                                      // K bytes 0..2 | 3..5 | 6..8
      wd[0] = v[0][0] | (v[1][0] << 24);
      wd[1] = (v[1][0] >> 8) | (v[2][0] << 16);
      wd[2] = v[2][0] >> 16;
      wd[3] = wd[4] = wd[5] = wd[6] = wd[7] = 0u;
    } else {                                                             // K bytes 0..8 | 9..17 | 18..26
      wd[0] = v[0][0];
      wd[1] = v[0][1];
      wd[2] = v[0][2] | (v[1][0] << 8);
      wd[3] = (v[1][0] >> 24) | (v[1][1] << 8);
      wd[4] = (v[1][1] >> 24) | (v[1][2] << 8) | (v[2][0] << 16);
      wd[5] = (v[2][0] >> 16) | (v[2][1] << 16);
      wd[6] = (v[2][1] >> 16) | (v[2][2] << 16);
      wd[7] = 0u;
    }
    // row of the im2col tile: natural order (pitch WP >= W), or [window position][pooled pixel] (WP = pooled width)
    uint8_t* d = POS ? dst + ((h & 1) * 2 + (w & 1)) * 2 * plane + ((h >> 1) * WP + (w >> 1)) * 16 : dst + (h * WP + w) * 16;
    *reinterpret_cast<uint4*>(d) = make_uint4(wd[0], wd[1], wd[2], wd[3]);
    if (9 * CIN > 16) *reinterpret_cast<uint4*>(d + plane) = make_uint4(wd[4], wd[5], wd[6], wd[7]);
  }
}

// ---- K0f: the net's resident image (A operands of every layer in UMMA core-matrix order, the dense kernel, the
// per-channel constants), built once per set of weights; the forward kernel brings it in with bulk copies.
__global__ void __launch_bounds__(256)
vgg_pack_kernel(const NfParams p, uint8_t* blob) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int nthr = gridDim.x * blockDim.x;
  // A operands: kernel rows in the K-major core-matrix layout (8-row groups 256 B apart, the two 16-byte K chunks of a
  // group 128 B apart); rows >= cout stay zero (the blob is cleared first)
  {
    const NfLayer& L0 = p.L[0];
    for (int i = tid; i < 64 * 2; i += nthr) {
      const int row = i >> 1, chunk = i & 1;
      uint32_t wd[4] = {0u, 0u, 0u, 0u};
      if (row < L0.cout) {
#pragma unroll
        for (int b = 0; b < 16; ++b) {
          const int k = chunk * 16 + b;                 // K index = (tap * cin + channel)
          if (k < 9 * L0.cin) {
            const int t = k / L0.cin, ci = k - t * L0.cin;
            const uint32_t v = (uint8_t)__ldg(L0.wpk + (row * 9 + t) * L0.cin_pad + ci);
            wd[b >> 2] |= v << (8 * (b & 3));
          }
        }
      }
      *reinterpret_cast<uint4*>(blob + L0.a_off + (row >> 3) * 256 + chunk * 128 + (row & 7) * 16) = make_uint4(wd[0], wd[1], wd[2], wd[3]);
    }
  }
  for (int l = 1; l < p.nconv; ++l) {
    const NfLayer& Ly = p.L[l];
    const int nhalf = Ly.cin >> 5;
    const int items = Ly.cout * 9 * nhalf * 2;
    const uint4* src = reinterpret_cast<const uint4*>(Ly.wpk);
    for (int i = tid; i < items; i += nthr) {
      const int chunk = i & 1;
      const int j = i >> 1;
      const int hf = j % nhalf;
      const int rt = j / nhalf;                          // row * 9 + tap
      const int row = rt / 9, t = rt - row * 9;
      // packed [cout][9][cin]: 16-byte pieces in exactly this order
      *reinterpret_cast<uint4*>(blob + Ly.a_off + (t * nhalf + hf) * NF_ABLK + (row >> 3) * 256 + chunk * 128 + (row & 7) * 16) = __ldg(src + i);
    }
  }
  for (int i = tid; i < ((p.units * p.fin) >> 4); i += nthr)
    *reinterpret_cast<uint4*>(blob + p.dw_off + i * 16) = __ldg(reinterpret_cast<const uint4*>(p.dense_w) + i);
  for (int i = tid; i < p.nconv * 64; i += nthr) {
    const int l = i >> 6, c = i & 63;
    const NfLayer& Ly = p.L[l];
    float4 k = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < Ly.cout) {
      const float bias = Ly.bias ? __ldg(Ly.bias + c) : 0.f;
      const float inv = Ly.bn_inv ? __ldg(Ly.bn_inv + c) : 1.f;
      const float shift = Ly.bn_inv ? __ldg(Ly.bn_shift + c) : 0.f;
      k = make_float4(Ly.acc_scale, bias, __fmul_rn(inv, Ly.qm), __fmul_rn(shift, Ly.qm));
    }
    *reinterpret_cast<float4*>(blob + p.cst_off + i * 16) = k;
  }
  for (int u = tid; u < p.units; u += nthr) {
    const Epi& e = p.dense_epi;
    *reinterpret_cast<float4*>(blob + p.dcst_off + u * 16) =
        make_float4(e.acc_scale, e.bias ? __ldg(e.bias + u) : 0.f, e.bn_inv ? __ldg(e.bn_inv + u) : 1.f, e.bn_inv ? __ldg(e.bn_shift + u) : 0.f);
  }
}

__global__ void __launch_bounds__(NF_THREADS, 1)
vgg_fused_kernel(const __grid_constant__ NfParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sg = smem_raw + (smem_base - smem_u32(smem_raw));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int G = (int)gridDim.x;
  const uint32_t bar_base = smem_base + p.bar_off;
  const uint32_t b_imfull = bar_base, b_imfree = bar_base + 8;
  auto tfull = [&](int b) { return bar_base + 16u + 8u * b; };
  auto tempty = [&](int b) { return bar_base + 32u + 8u * b; };
  auto rawfull = [&](int b) { return bar_base + 48u + 8u * b; };
  auto featfull = [&](int b) { return bar_base + 64u + 8u * b; };
  auto actfull = [&](int l) { return bar_base + 80u + 8u * l; };
  const uint32_t b_blob = bar_base + 136u;
  const uint32_t b_c0full = bar_base + 144u;        // pooled first layer: its MMAs have completed
  const uint32_t tmem_slot = bar_base + 160u;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(sg + p.bar_off + 160);

#ifdef QNNB_TRACE
  int trk__ = 0;
#endif
  nftrace(p, 1, 0);
  // ------------------------------------------------------------------ prologue (independent of the previous kernel)
  if (tid == 0) {
    griddep_launch_dependents();
    mbar_init(b_imfull, NF_WORKERS);
    mbar_init(b_imfree, 1);
    mbar_init(b_blob, 1);
    mbar_init(b_c0full, 1);
    const int pm = p.L[0].posmajor;                 // then all 16 warps write the first layer's output
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull(b), 1);
      mbar_init(tempty(b), 8);
      mbar_init(rawfull(b), 1);
      mbar_init(featfull(b), (pm && p.nconv == 1) ? 16 : 8);
    }
    for (int l = 0; l < NF_MAXL; ++l) mbar_init(actfull(l), (pm && l == 0) ? 16 : 8);
    fence_barrier_init();
    // the net's resident image: weights only, so it does not wait for the previous kernel
    mbar_expect_tx(b_blob, (uint32_t)p.blob_bytes);
    for (int off = 0; off < p.blob_bytes; off += 32768) {
      const int sz = min(32768, p.blob_bytes - off);
      bulk_load_1d(smem_base + off, p.blob + off, (uint32_t)sz, b_blob);
    }
  }
  if (warp == NF_MMA_WARP) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  // halo pixels of every raster must read as zero (SAME padding); interiors are rewritten per image
  for (int i = tid; i < (p.zero_bytes >> 4); i += NF_THREADS)
    *reinterpret_cast<uint4*>(sg + p.zero_off + i * 16) = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async();                   // zero fills were written through the generic proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  nftrace(p, 2, 0);

  const int nloc = p.n > (int)blockIdx.x ? (p.n - (int)blockIdx.x + G - 1) / G : 0;     // images of this CTA

  // the pooled first layer's epilogue, shared by all 16 warps (warp = lane quarter x channel group)
  auto first_pooled = [&](int k) {
    const NfLayer& L0 = p.L[0];
    mbar_wait_parked(b_c0full, (uint32_t)k & 1u);
    tc_fence_after();
    const float4* cst = reinterpret_cast<const float4*>(sg + p.cst_off);
    uint8_t* out = L0.out_plane == 0 ? sg + p.feat_off + (k & 1) * p.feat_stride : sg + L0.out_off;
    if (L0.sign) nf_first_pooled<true>(tmem_base, warp, lane, L0.mtiles, L0.npp, L0.ppw, L0.cout, L0.qm, cst, out, L0.out_plane, L0.out_wp);
    else nf_first_pooled<false>(tmem_base, warp, lane, L0.mtiles, L0.npp, L0.ppw, L0.cout, L0.qm, cst, out, L0.out_plane, L0.out_wp);
    tc_fence_before();
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) mbar_arrive(p.nconv > 1 ? actfull(0) : featfull(k & 1));
  };
  const bool posmajor = p.L[0].posmajor != 0;

  if (warp == NF_MMA_WARP) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      mbar_wait(b_blob, 0u);
      uint32_t q = 0;
      for (int k = 0; k < nloc; ++k) {
        for (int l = 0; l < p.nconv; ++l) {
          const NfLayer& Ly = p.L[l];
          if (l == 0 && posmajor) {
            // pixels on M: A = the position-major im2col tiles (u8), B = the kernel block (64 rows, s8); one MMA per
            // (M tile, window position) into its own 64 columns.  Both accumulator buffers are used: every reader of
            // the previous image must be done (its last layer's warps have arrived on featfull).
            mbar_wait(b_imfull, (uint32_t)k & 1u);
            if (k > 0) mbar_wait(featfull((k - 1) & 1), (uint32_t)((k - 1) >> 1) & 1u);
            tc_fence_after();
            nftrace(p, 10, 0);
            const uint32_t idesc = make_idesc_i8(128, 64, /*a signed*/ false, /*b signed*/ true);
            const uint64_t w_desc = make_smem_desc_interleaved(smem_base + Ly.a_off, 128, 256);
            const uint64_t x_desc = make_smem_desc_interleaved(smem_base + Ly.in_off, Ly.in_plane, 128);
            for (int mt = 0; mt < Ly.mtiles; ++mt)
              for (int pos = 0; pos < 4; ++pos)
                umma_i8(tmem_base + (uint32_t)(mt * 256 + pos * 64), x_desc + (uint64_t)((pos * 2 * Ly.in_plane + mt * 2048) >> 4), w_desc, idesc, 0u);
            umma_commit(b_c0full);
            umma_commit(b_imfree);
            nftrace(p, 12, 0);
            continue;
          }
          const int nblk = Ly.nblk, rb = Ly.rb, lh = Ly.h, wp = Ly.wp, nhalf = Ly.cin >> 5;
          // descriptors advance by (bytes >> 4) in their 14-bit address field
          const uint64_t a_desc0 = make_smem_desc_interleaved(smem_base + Ly.a_off, 128, 256);
          const uint64_t b_desc0 = make_smem_desc_interleaved(smem_base + Ly.in_off, Ly.in_plane, 128);
          const uint32_t plane16 = (uint32_t)(Ly.in_plane >> 4);
          if (l == 0) mbar_wait(b_imfull, (uint32_t)k & 1u);
          else mbar_wait(actfull(l - 1), (uint32_t)k & 1u);
          tc_fence_after();
          nftrace(p, 10, l);
          for (int blk = 0; blk < nblk; ++blk, ++q) {
            const int buf = (int)(q & 1u);
            mbar_wait(tempty(buf), ((q >> 1) & 1u) ^ 1u);
            tc_fence_after();
            nftrace(p, 11, l * 16 + blk);
            const int h0 = blk * rb;
            const int rows = min(rb, lh - h0);
            const int N = (rows * wp + 15) & ~15;
            const uint32_t d_tmem = tmem_base + (uint32_t)(buf * 256);
            if (l == 0) {
              const uint32_t idesc = make_idesc_i8(128, N, /*a signed*/ true, /*b unsigned*/ false);
              umma_i8(d_tmem, a_desc0, b_desc0 + (uint64_t)(h0 * wp), idesc, 0u);
            } else {
              const uint32_t idesc = make_idesc_i8(128, N, true, true);
              uint64_t a_desc = a_desc0;
              uint32_t acc = 0u;
#pragma unroll
              for (int r = 0; r < 3; ++r) {
                const uint64_t b_row = b_desc0 + (uint64_t)((h0 + r) * wp);
#pragma unroll
                for (int s = 0; s < 3; ++s) {
                  for (int hf = 0; hf < nhalf; ++hf) {
                    umma_i8(d_tmem, a_desc, b_row + (uint64_t)(s + hf * 2 * plane16), idesc, acc);
                    a_desc += NF_ABLK >> 4;
                    acc = 1u;
                  }
                }
              }
            }
            umma_commit(tfull(buf));
            nftrace(p, 12, l * 16 + blk);
          }
          if (l == 0) umma_commit(b_imfree);       // the im2col tile may be rebuilt once these MMAs have completed
        }
      }
    }
  } else if ((warp & 3) < 2) {
    // ===================== epilogue (8 warps: TMEM lanes 0..63 = kernel rows) =====================
    const int pair = warp >> 2;                    // 0..3: takes row pairs (rows) pair, pair + 4, ... of a block
    const int half = warp & 3;                     // lanes 32 * half ..
    const int c = half * 32 + lane;
    mbar_wait_parked(b_blob, 0u);
    griddep_wait();                                // (nothing of the previous kernel is read here; keeps the exit ordered)
    uint32_t q = 0;
    for (int k = 0; k < nloc; ++k) {
      for (int l = 0; l < p.nconv; ++l) {
        if (l == 0 && posmajor) { first_pooled(k); nftrace(p, 21, 0); continue; }
        const NfLayer& Ly = p.L[l];
        const int nblk = Ly.nblk, rb = Ly.rb, lh = Ly.h, lw = Ly.w, wp = Ly.wp, pool = Ly.pool, sign = Ly.sign, cout = Ly.cout;
        const float qm = Ly.qm;
        const bool active = half * 32 < cout;      // warp-uniform
        const float4 kc = *reinterpret_cast<const float4*>(sg + p.cst_off + (l * 64 + c) * 16);
        QConst qc;
        qc.s = kc.x; qc.a = kc.y; qc.b = kc.z; qc.c = kc.w;
        const int flip = kc.z < 0.f ? -1 : 0;
        const bool any_dec = pool && __any_sync(0xffffffffu, flip != 0);
        // output pixel (oh, ow) of channel c: next raster [c / 16][(oh + 1) * pitch + ow + 1][c % 16], or flat [oh][ow][c]
        uint8_t* obase;
        int orow, opix;
        const bool flat = Ly.out_plane == 0;
        if (flat) {
          opix = cout; orow = (pool ? (lw >> 1) : lw) * cout;
          obase = sg + p.feat_off + (k & 1) * p.feat_stride + c;
        } else {
          opix = 16; orow = Ly.out_wp * 16;
          obase = sg + Ly.out_off + (c >> 4) * Ly.out_plane + (c & 15) + orow + 16;
        }
        for (int blk = 0; blk < nblk; ++blk, ++q) {
          const int buf = (int)(q & 1u);
          mbar_wait_parked(tfull(buf), (q >> 1) & 1u);
          tc_fence_after();
          nftrace(p, 20, l * 16 + blk);
          const int h0 = blk * rb;
          const int rows = min(rb, lh - h0);
          const uint32_t taddr = tmem_base + ((uint32_t)(half * 32) << 16) + (uint32_t)(buf * 256);
          if (active) {
#define NF_CALL(POOL_, SIGN_, DEC_)                                                                                            \
  do {                                                                                                                         \
    if (flat) nf_epi_block<POOL_, SIGN_, DEC_, 0>(taddr, pair, rows, h0, lw, wp, obase, orow, opix, qc, qm, flip NF_TRPASS);  \
    else nf_epi_block<POOL_, SIGN_, DEC_, 16>(taddr, pair, rows, h0, lw, wp, obase, orow, opix, qc, qm, flip NF_TRPASS);       \
  } while (0)
            if (pool) {
              if (sign) { if (any_dec) NF_CALL(true, true, true); else NF_CALL(true, true, false); }
              else { if (any_dec) NF_CALL(true, false, true); else NF_CALL(true, false, false); }
            } else {
              if (sign) NF_CALL(false, true, false); else NF_CALL(false, false, false);
            }
#undef NF_CALL
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty(buf));
          nftrace(p, 21, l * 16 + blk);
        }
        // this warp's part of the layer output is in shared memory: visible to the tensor core (async proxy) / the dense warps
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(l + 1 < p.nconv ? actfull(l) : featfull(k & 1));
      }
    }
  } else {
    // ===================== workers: image fetch, first-layer im2col, dense head =====================
    const int wi = (warp >> 2) * 2 + (warp & 3) - 2;          // 0..7
    const int wt = wi * 32 + lane;
    const int l0h = p.L[0].h, l0w = p.L[0].w, l0wp = p.L[0].wp, l0cin = p.L[0].cin, im_off = p.L[0].in_off, im_plane = p.L[0].in_plane;
    auto fetch = [&](int k) {
      if (k >= nloc) return;
      const long long img = (long long)blockIdx.x + (long long)k * G;
      mbar_expect_tx(rawfull(k & 1), (uint32_t)p.img_bytes);
      bulk_load_1d(smem_base + p.raw_off + (k & 1) * p.raw_stride, p.x + img * p.img_bytes, (uint32_t)p.img_bytes, rawfull(k & 1));
    };
    const int units = p.units, fin = p.fin;
    auto dense = [&](int kk) {
      mbar_wait_parked(featfull(kk & 1), (uint32_t)(kk >> 1) & 1u);
      const long long img = (long long)blockIdx.x + (long long)kk * G;
      const int* fw = reinterpret_cast<const int*>(sg + p.feat_off + (kk & 1) * p.feat_stride);
      const int nw = fin >> 2;
      for (int u = wi; u < units; u += NF_WORKERS) {          // warp-uniform
        const int* ww = reinterpret_cast<const int*>(sg + p.dw_off + u * fin);
        int acc = 0;
        for (int i = lane; i < nw; i += 32) acc = __dp4a(fw[i], ww[i], acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) {
          const float4 kc = *reinterpret_cast<const float4*>(sg + p.dcst_off + u * 16);
          ChanConst cc;
          cc.scale = kc.x; cc.bias = kc.y; cc.inv = kc.z; cc.shift = kc.w;
          cc.has_bias = p.dense_epi.bias != nullptr; cc.has_bn = p.dense_epi.bn_inv != nullptr;
          p.y[img * units + u] = affine((float)acc, cc);
        }
      }
    };
    griddep_wait();                        // the images / the output buffer only after the previous kernel has completed
    nftrace(p, 3, 0);
    if (wt == 0) fetch(0);
    for (int k = 0; k < nloc; ++k) {
      named_bar_sync(1, NF_WORKERS * 32);                     // every worker is done with raw buffer (k + 1) & 1
      if (wt == 0) fetch(k + 1);
      mbar_wait_parked(rawfull(k & 1), (uint32_t)(k >> 1) & 1u);
      nftrace(p, 30, k);
      if (k >= 1) mbar_wait_parked(b_imfree, (uint32_t)(k - 1) & 1u);
      const uint8_t* raw = sg + p.raw_off + (k & 1) * p.raw_stride;
      if (posmajor) {
        if (l0cin == 1) nf_im2col<1, true>(raw, sg + im_off, im_plane, l0h, l0w, p.L[0].ppw, wt, NF_WORKERS * 32);
        else nf_im2col<3, true>(raw, sg + im_off, im_plane, l0h, l0w, p.L[0].ppw, wt, NF_WORKERS * 32);
      } else {
        if (l0cin == 1) nf_im2col<1, false>(raw, sg + im_off, im_plane, l0h, l0w, l0wp, wt, NF_WORKERS * 32);
        else nf_im2col<3, false>(raw, sg + im_off, im_plane, l0h, l0w, l0wp, wt, NF_WORKERS * 32);
      }
      fence_proxy_async();                                    // im2col rows -> visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(b_imfull);
      nftrace(p, 31, k);
      if (k == 0) mbar_wait_parked(b_blob, 0u);               // the dense kernel / constants are part of the resident image
      if (posmajor) first_pooled(k);
      if (k >= 1) dense(k - 1);
      nftrace(p, 33, k);
    }
    if (nloc >= 1) dense(nloc - 1);
    nftrace(p, 34, 0);
  }

  nftrace(p, 40, 0);
  tc_fence_before();
  __syncthreads();
  if (warp == NF_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
#ifdef QNNB_TRACE
  if (p.tr != nullptr && blockIdx.x == 0 && lane == 0) p.tr[warp * 1024] = (unsigned long long)trk__;
#endif
}

// host-side geometry; returns a message when the net is outside the kernel's scope
const char* nf_plan(const qnnb_vgg_desc& d, NfParams& p, int& smem_bytes) {
  memset(&p, 0, sizeof(p));
  if (d.nconv < 1 || d.nconv > NF_MAXL) return "1..6 convolutions";
  if (d.h < 4 || d.w < 4 || d.h > 32 || d.w > 32 || (d.cin != 1 && d.cin != 3)) return "images up to 32x32 with 1 or 3 channels";
  if (d.units < 1 || d.units > 32) return "at most 32 dense units";
  if ((d.h * d.w * d.cin) % 16 != 0) return "image bytes must be a multiple of 16";
  int off = 0;
  // ---- the resident image (same offsets in the packed blob and in shared memory): A blocks, dense kernel, constants
  int h = d.h, w = d.w, cin = d.cin;
  for (int l = 0; l < d.nconv; ++l) {
    const qnnb_net_conv& c = d.conv[l];
    NfLayer& L = p.L[l];
    if (c.cout != 32 && c.cout != 64) return "32 or 64 filters per layer";
    if (c.epi.act != QNNB_ACT_QUANT && c.epi.act != QNNB_ACT_SIGN_I8) return "quantized_tanh / binary_tanh (int8 levels) activations";
    if (c.epi.act == QNNB_ACT_QUANT && (c.epi.abits < 2 || c.epi.abits > 8)) return "abits 2..8";
    if (c.epi.res_kind != QNNB_KIND_NONE) return "no residual";
    if (c.pool != 0 && c.pool != 2) return "pool 0 or 2";
    if (c.pool && (h < 2 || w < 2)) return "pooled map would be empty";
    L.h = h; L.w = w; L.cin = cin; L.cout = c.cout;
    L.pool = c.pool ? 1 : 0;
    L.sign = c.epi.act == QNNB_ACT_SIGN_I8;
    L.qm = L.sign ? 1.f : (float)(1 << (c.epi.abits - 1));
    L.acc_scale = c.epi.acc_scale;
    L.wpk = (const int8_t*)c.w; L.bias = c.epi.bias; L.bn_inv = c.epi.bn_inv; L.bn_shift = c.epi.bn_shift;
    L.cin_pad = (cin + 3) & ~3;
    L.wp = l == 0 ? w : w + 2;
    L.a_off = off;
    off += (l == 0 ? 1 : 9 * (cin / 32)) * NF_ABLK;
    // rows per block: N = rows * pitch (+ the 8-column loads' overhang past the last row) <= 256 accumulator columns
    const int over = ((w + 7) / 8) * 8 - L.wp;
    const int limit = 256 - (over > 0 ? over : 0);
    int rb = limit / L.wp;
    if (L.pool) rb &= ~1;
    if (rb < (L.pool ? 2 : 1)) return "map too wide";
    if (rb > h) rb = h;
    int nblk = (h + rb - 1) / rb;
    if (nblk == 1 && h >= 8) nblk = 2;                 // two blocks: the second block's MMAs overlap the first one's epilogue
    rb = (h + nblk - 1) / nblk;
    if (L.pool) rb = (rb + 1) & ~1;
    nblk = (h + rb - 1) / rb;
    L.rb = rb; L.nblk = nblk;
    if (l == 0 && L.pool) {
      // pooled first layer: pixels on M, position-major im2col (see nf_first_pooled)
      L.posmajor = 1;
      L.ppw = w >> 1;
      L.npp = (h >> 1) * (w >> 1);
      L.mtiles = (L.npp + 127) / 128;
      if (L.mtiles > 2) return "first layer: more than 256 pooled pixels";
    }
    if (L.pool) { h >>= 1; w >>= 1; }
    cin = c.cout;
  }
  off += NF_ABLK;                                       // the last block's rows 64..127 are read from here (zeros)
  if (h < 1 || w < 1) return "empty feature map";
  const int fin = h * w * cin;
  p.fin = fin;
  p.dw_off = off; off += (d.units * fin + 15) & ~15;
  p.cst_off = off; off += d.nconv * 64 * 16;
  p.dcst_off = off; off += d.units * 16;
  p.blob_bytes = off;
  // ---- per-image state
  {
    NfLayer& L = p.L[0];                                // first-layer im2col: two 16-byte K planes of (pixels + 16) rows
    L.in_off = off;
    if (L.posmajor) {
      L.in_plane = L.mtiles * 128 * 16;                 // [position][plane][pooled pixel] x 16 bytes
      off += 4 * 2 * L.in_plane;
    } else {
      L.in_plane = (L.h * L.wp + 16) * 16;
      off += 2 * L.in_plane;
    }
  }
  p.zero_off = off;                                     // rasters (inputs of layers 1..): zero-haloed maps, cin / 16 planes
  for (int l = 1; l < d.nconv; ++l) {
    NfLayer& L = p.L[l];
    L.in_off = off;
    L.in_plane = ((L.h + 3) * L.wp + 32) * 16;
    off += (L.cin / 16) * L.in_plane;
    p.L[l - 1].out_off = L.in_off; p.L[l - 1].out_plane = L.in_plane; p.L[l - 1].out_wp = L.wp;
  }
  p.zero_bytes = off - p.zero_off;
  p.L[d.nconv - 1].out_plane = 0;
  p.feat_stride = (fin + 15) & ~15;
  p.feat_off = off; off += 2 * p.feat_stride;
  p.img_bytes = d.h * d.w * d.cin;
  p.raw_stride = p.img_bytes;
  p.raw_off = off; off += 2 * p.raw_stride;
  p.bar_off = (off + 15) & ~15; off = p.bar_off + 256;
  smem_bytes = off + 1024;
  if (smem_bytes > NF_SMEM_MAX) return "kernels + activation maps exceed one SM's shared memory";
  p.n = d.n; p.nconv = d.nconv; p.units = d.units;
  p.dense_w = (const int8_t*)d.dense_w;
  p.dense_epi = make_epi(d.dense_epi);
  return nullptr;
}

}  // namespace

bool vgg_fused_supported(const qnnb_vgg_desc& d, const char** why) {
  NfParams p;
  int smem = 0;
  const char* msg = nf_plan(d, p, smem);
  if (why) *why = msg ? msg : "";
  return msg == nullptr;
}

long long vgg_fused_blob_bytes(const qnnb_vgg_desc& d) {
  NfParams p;
  int smem = 0;
  return nf_plan(d, p, smem) ? 0 : (long long)p.blob_bytes;
}

int launch_vgg_pack(const qnnb_vgg_desc& d, void* blob, cudaStream_t st) {
  NfParams p;
  int smem = 0;
  const char* msg = nf_plan(d, p, smem);
  if (msg) { set_error("vgg_pack: outside the whole-network kernel's scope (%s)", msg); return QNNB_EUNSUPPORTED; }
  for (int l = 0; l < d.nconv; ++l)
    if (!d.conv[l].w) { set_error("vgg_pack: conv[%d].w is null", l); return QNNB_EINVAL; }
  if (!d.dense_w || !blob) { set_error("vgg_pack: null pointer"); return QNNB_EINVAL; }
  if (((uintptr_t)blob & 15u) != 0) { set_error("vgg_pack: the blob must be 16-byte aligned"); return QNNB_EINVAL; }
  QNNB_CUDA(cudaMemsetAsync(blob, 0, (size_t)p.blob_bytes, st));
  vgg_pack_kernel<<<32, 256, 0, st>>>(p, (uint8_t*)blob);
  QNNB_CUDA(cudaGetLastError());
  return QNNB_OK;
}

int launch_vgg_fused(const qnnb_vgg_desc& d, const void* blob, const void* x, float* y, cudaStream_t st) {
  NfParams p;
  int smem = 0;
  const char* msg = nf_plan(d, p, smem);
  if (msg) { set_error("vgg_forward: outside the whole-network kernel's scope (%s)", msg); return QNNB_EUNSUPPORTED; }
  if (((uintptr_t)x & 15u) != 0 || ((uintptr_t)blob & 15u) != 0) { set_error("vgg_forward: the image batch and the blob must be 16-byte aligned"); return QNNB_EINVAL; }
  p.x = (const uint8_t*)x; p.y = y; p.blob = (const uint8_t*)blob;
  p.tr = get_trace_buffer();
  static SmemConfigured once;
  QNNB_CUDA(once.ensure(vgg_fused_kernel, smem));
  const int grid = d.n < grid_sms(d.max_ctas) ? d.n : grid_sms(d.max_ctas);
  QNNB_CUDA(launch_pdl(vgg_fused_kernel, dim3(grid), dim3(NF_THREADS), (size_t)smem, st, p));
  return QNNB_OK;
}

}  // namespace qnnb
