// K4 -- 3x3 stride-1 SAME convolution of FP32 activations with small-integer kernels on the tensor cores.
//
// Stands in for K.conv2d + bias_add + BatchNormalization (+ add + Lambda(x*0.5)) + LeakyReLU of the `qnn`, `bnn` and
// `tnn` network types (models/model_factory.py:24-58: quantised / binarised / ternarised kernels, fp32 LeakyReLU
// activations), i.e. QuantizedConv2D.call (layers/quantized_layers.py:164-194), BinaryConv2D.call
// (layers/binary_layers.py:160-187), TernaryConv2D.call (layers/ternary_layers.py:156-174) as chained by
// models/resnet.py:57-129.  These maps are narrow (16 / 32 / 64 channels) and every layer is HBM bound
// (SURVEY.md 8d, config 5), so the job of the kernel is to stream activations once at memory speed.
//
// Arithmetic.  The kernel levels are exact in bf16; an fp32 activation is split EXACTLY into three bf16 terms
// x = hi + mid + lo (8 mantissa bits each), so  sum x*w = sum hi*w + sum mid*w + sum lo*w  is three bf16 MMAs whose
// products are exact and whose fp32 accumulation happens in TMEM -- fp32-grade results (the tolerance class of this
// path: <= 1e-4 relative, tests/) at tensor-core speed, instead of FFMA loops on the CUDA cores.
//
// Orientation: D[pixel][cout] = A[pixel][cin] * W[cout][cin]^T per filter tap: PIXELS are the MMA's M (128 = 16
// groups of 8 consecutive pixels of one image row), output channels its N (16..64), so an epilogue thread owns one
// pixel and reads / writes its cout contiguous floats (NHWC rows) straight from / to global memory.
//
// Pipeline per (tile, 16-channel chunk), all stages mbarrier rings:
//   warp 0      TMA: one 4-D box {16 ch, 10 px, TN images, TH+2 rows} of the fp32 NHWC tensor, i.e. the pixel tile with
//               its halo; out-of-bounds rows / columns are zero-filled by the TMA unit (= SAME padding)
//   warps 8-15  split every halo pixel into the three bf16 planes, stored chunk-major [8-ch chunk][halo pixel][16 B]
//               (un-swizzled K-major core matrices: 8 consecutive pixels = 8 rows x 16 B)
//   warp 1      27 MMAs (3 planes x 9 taps, kind::f16, M=128, N=cout, K=16): tap (r,s) is the SAME plane viewed
//               through a descriptor shifted by (r*TN*10 + s) pixels, group stride = one halo row (as in K1 v2);
//               all taps' kernels stay resident in shared memory as bf16
//   warps 4-7   epilogue: tcgen05.ld 16 columns at a time, scale / bias / BN / residual / LeakyReLU in the fixed
//               fp32 op order of common.cuh, 128-bit stores
#include "common.cuh"
#include "tc_ptx.cuh"

#include <cuda.h>
#include <cuda_bf16.h>

namespace qnnb {

namespace {

using namespace tcx;

constexpr int F_THREADS = 512;
constexpr int F_CVT_WARPS = 8;
constexpr int F_CVT_THREADS = F_CVT_WARPS * 32;
constexpr int F_EPI_WARPS = 4;
constexpr int F_KC = 16;             // channels per pipeline unit (= one bf16 MMA K step)
constexpr int F_FSTAGES = 4;         // fp32 halo ring
constexpr int F_PSTAGES = 3;         // bf16 plane ring
constexpr int F_ACCS = 4;            // TMEM accumulators, 64 columns each
constexpr int F_MAXC = 64;           // channel limit (Cin and Cout)
constexpr int F_W_BYTES = 9 * F_MAXC * F_MAXC * 2;

struct F32Params {
  int n, h, w, cin, cout;
  int tiles_w, tiles_h, num_tiles, kchunks;
  FastDiv fd_w, fd_h;
  const int8_t* wpk;                 // packed kernel levels [cout][3][3][cin]
  float* y;
  Epi epi;
};

template <int TH>
struct F32Smem {
  static constexpr int TN = 16 / TH;
  static constexpr int HALO_PX = (TH + 2) * TN * 10;
  static constexpr int F_BYTES = HALO_PX * F_KC * 4;
  // 16-byte chunk planes: the two chunks of a K step must fall into different bank halves (stride = 64 mod 128)
  static constexpr int K8_BYTES = HALO_PX * 16 + ((HALO_PX * 16) % 128 == 64 ? 0 : 64);
  static constexpr int PLANE_BYTES = 2 * K8_BYTES;
  static constexpr int PSTAGE_BYTES = 3 * PLANE_BYTES;
  static constexpr int F_OFF = 0;
  static constexpr int P_OFF = (F_OFF + F_FSTAGES * F_BYTES + 127) / 128 * 128;
  static constexpr int W_OFF = (P_OFF + F_PSTAGES * PSTAGE_BYTES + 127) / 128 * 128;
  static constexpr int C_OFF = W_OFF + F_W_BYTES;            // bias[64], inv[64], shift[64]
  static constexpr int BAR_OFF = C_OFF + 3 * F_MAXC * 4;
  static constexpr int TOTAL = BAR_OFF + 256 + 128;
  static_assert(F_BYTES % 128 == 0, "TMA destination alignment");
};

// D[tmem] (+)= A[smem] * B[smem], bf16 x bf16 -> fp32
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// 32 lanes x 16 consecutive columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// x0, x1 -> three bf16x2 words (low half = x0): hi = bf16(x), mid = bf16(x - hi), lo = bf16(x - hi - mid).
// Both subtractions are exact in fp32, so hi + mid + lo reproduces x to 24 bits.
__device__ __forceinline__ void split3(float x0, float x1, uint32_t& hi, uint32_t& mid, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
  const float r0 = __fsub_rn(x0, __low2float(h)), r1 = __fsub_rn(x1, __high2float(h));
  const __nv_bfloat162 m = __floats2bfloat162_rn(r0, r1);
  const float q0 = __fsub_rn(r0, __low2float(m)), q1 = __fsub_rn(r1, __high2float(m));
  const __nv_bfloat162 l = __floats2bfloat162_rn(q0, q1);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  mid = *reinterpret_cast<const uint32_t*>(&m);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

template <int TH>
__global__ void __launch_bounds__(F_THREADS, 1)
conv3x3_f32_tc_kernel(const __grid_constant__ CUtensorMap map_x, const F32Params p) {
  using SL = F32Smem<TH>;
  constexpr int TN = SL::TN;
  constexpr int HALO_PX = SL::HALO_PX;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 127u) & ~127u;
  uint8_t* sg = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + SL::BAR_OFF;
  auto ffull = [&](int s) { return bar_base + 8u * s; };
  auto fempty = [&](int s) { return bar_base + 8u * (F_FSTAGES + s); };
  auto pfull = [&](int b) { return bar_base + 8u * (2 * F_FSTAGES + b); };
  auto pempty = [&](int b) { return bar_base + 8u * (2 * F_FSTAGES + F_PSTAGES + b); };
  auto tfull = [&](int a) { return bar_base + 8u * (2 * F_FSTAGES + 2 * F_PSTAGES + a); };
  auto tempty = [&](int a) { return bar_base + 8u * (2 * F_FSTAGES + 2 * F_PSTAGES + F_ACCS + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * F_FSTAGES + 2 * F_PSTAGES + 2 * F_ACCS);
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(sg + SL::BAR_OFF + 8 * (2 * F_FSTAGES + 2 * F_PSTAGES + 2 * F_ACCS));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cout = p.cout;
  const int kchunks = p.kchunks;

  // ---- resident kernel: bf16, [(tap, chunk, 8-ch half)][cout][16 B] (un-swizzled K-major core matrices)
  {
    const int items = 9 * kchunks * 2 * cout;
    for (int i = threadIdx.x; i < items; i += F_THREADS) {
      const int co = i % cout;
      int r = i / cout;
      const int k8 = r & 1; r >>= 1;
      const int kc = r % kchunks;
      const int tap = r / kchunks;
      const int8_t* src = p.wpk + ((long long)co * 9 + tap) * p.cin + kc * F_KC + k8 * 8;
      const uint2 raw = __ldg(reinterpret_cast<const uint2*>(src));
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t word = (j < 2) ? raw.x : raw.y;
        const int b0 = (int)(int8_t)(word >> (16 * (j & 1)));
        const int b1 = (int)(int8_t)(word >> (16 * (j & 1) + 8));
        const __nv_bfloat162 v = __floats2bfloat162_rn((float)b0, (float)b1);
        o[j] = *reinterpret_cast<const uint32_t*>(&v);
      }
      *reinterpret_cast<uint4*>(sg + SL::W_OFF + (size_t)i * 16) = make_uint4(o[0], o[1], o[2], o[3]);
    }
    float* cst = reinterpret_cast<float*>(sg + SL::C_OFF);
    for (int c = threadIdx.x; c < F_MAXC; c += F_THREADS) {
      const bool ok = c < cout;
      cst[c] = (p.epi.bias != nullptr && ok) ? __ldg(p.epi.bias + c) : 0.f;
      cst[F_MAXC + c] = (p.epi.bn_inv != nullptr && ok) ? __ldg(p.epi.bn_inv + c) : 1.f;
      cst[2 * F_MAXC + c] = (p.epi.bn_inv != nullptr && ok) ? __ldg(p.epi.bn_shift + c) : 0.f;
    }
  }
  if (warp == 0 && lane == 0) tma_prefetch_desc(&map_x);
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < F_FSTAGES; ++s) { mbar_init(ffull(s), 1); mbar_init(fempty(s), F_CVT_WARPS); }
    for (int b = 0; b < F_PSTAGES; ++b) { mbar_init(pfull(b), F_CVT_WARPS); mbar_init(pempty(b), 1); }
    for (int a = 0; a < F_ACCS; ++a) { mbar_init(tfull(a), 1); mbar_init(tempty(a), F_EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  fence_proxy_async();                   // resident kernel written through the generic proxy, read by the tensor core
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  auto decode = [&](int tile, int& n0, int& h0, int& w0) {
    const int q = fdiv(tile, p.fd_w);
    w0 = (tile - q * p.tiles_w) * 8;
    const int q2 = fdiv(q, p.fd_h);
    h0 = (q - q2 * p.tiles_h) * TH;
    n0 = q2 * TN;
  };

  if (warp == 0) {
    // ===================== TMA: fp32 halo tile per (tile, chunk) =====================
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        int n0, h0, w0;
        decode(tile, n0, h0, w0);
        for (int kc = 0; kc < kchunks; ++kc) {
          mbar_wait(fempty(s), ph ^ 1u);
          mbar_expect_tx(ffull(s), SL::F_BYTES);
          tma_load_4d(smem_base + SL::F_OFF + s * SL::F_BYTES, &map_x, ffull(s), kc * F_KC, w0 - 1, n0, h0 - 1);
          if (++s == F_FSTAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // instruction descriptor: dense, F32 accumulate, BF16 x BF16, both K-major, N = cout, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(cout >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t w_lbo = (uint32_t)cout * 16u;
      int b = 0; uint32_t pph = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const int acc = it % F_ACCS;
        const uint32_t acc_phase = (uint32_t)(it / F_ACCS) & 1u;
        mbar_wait(tempty(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 64);
        for (int kc = 0; kc < kchunks; ++kc) {
          mbar_wait(pfull(b), pph);
          tc_fence_after();
          const uint32_t planes = smem_base + SL::P_OFF + b * SL::PSTAGE_BYTES;
#pragma unroll 1
          for (int split = 0; split < 3; ++split) {
            const uint32_t plane = planes + split * SL::PLANE_BYTES;
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              const int r = tap / 3, s = tap - 3 * r;
              // A: 16 groups of 8 pixels, one halo row (10 px x 16 B) apart; the two 16-byte K chunks K8_BYTES apart
              const uint64_t a_desc = make_smem_desc_interleaved(plane + (uint32_t)((r * TN * 10 + s) * 16), SL::K8_BYTES, 160);
              const uint64_t b_desc = make_smem_desc_interleaved(smem_base + SL::W_OFF + (uint32_t)((tap * kchunks + kc) * 2) * w_lbo, w_lbo, 128);
              umma_bf16(d_tmem, a_desc, b_desc, idesc, (kc > 0 || split > 0 || tap > 0) ? 1u : 0u);
            }
          }
          umma_commit(pempty(b));
          if (++b == F_PSTAGES) { b = 0; pph ^= 1u; }
        }
        umma_commit(tfull(acc));
      }
    }
  } else if (warp >= 4 && warp < 4 + F_EPI_WARPS) {
    // ===================== epilogue: thread = pixel =====================
    const int quarter = warp & 3;
    const int L = quarter * 32 + lane;               // TMEM lane = pixel of the tile
    const int g = L >> 3, px = L & 7;
    const int row = g / TN, img = g % TN;
    const Epi& e = p.epi;
    const float* cst = reinterpret_cast<const float*>(sg + SL::C_OFF);
    const bool has_bias = e.bias != nullptr, has_bn = e.bn_inv != nullptr;
    const bool has_res = e.res_kind == QNNB_KIND_F32;
    const bool leaky = e.act == QNNB_ACT_LEAKY;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      int n0, h0, w0;
      decode(tile, n0, h0, w0);
      const int acc = it % F_ACCS;
      const uint32_t acc_phase = (uint32_t)(it / F_ACCS) & 1u;
      const int nimg = n0 + img;
      const bool valid = nimg < p.n;
      const long long pix = ((long long)nimg * p.h + (h0 + row)) * p.w + (w0 + px);
      float* yrow = p.y + pix * cout;
      const float* rrow = reinterpret_cast<const float*>(e.residual) + pix * cout;
      mbar_wait_parked(tfull(acc), acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * 64);
      for (int c0 = 0; c0 < cout; c0 += 16) {
        float4 rs[4];
        if (has_res && valid) {
#pragma unroll
          for (int j = 0; j < 4; ++j) rs[j] = __ldg(reinterpret_cast<const float4*>(rrow + c0) + j);
        }
        float v[16];
        __syncwarp();
        tmem_ld16(taddr + (uint32_t)c0, v);
        if (c0 + 16 >= cout) {
          // last TMEM read of this tile: hand the accumulator back before the arithmetic
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty(acc));
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float z = __fmul_rn(v[j], e.acc_scale);
          if (has_bias) z = __fadd_rn(z, cst[c0 + j]);
          if (has_bn) { z = __fmul_rn(z, cst[F_MAXC + c0 + j]); z = __fadd_rn(z, cst[2 * F_MAXC + c0 + j]); }
          if (has_res) {
            const float sc = (j & 3) == 0 ? rs[j >> 2].x : ((j & 3) == 1 ? rs[j >> 2].y : ((j & 3) == 2 ? rs[j >> 2].z : rs[j >> 2].w));
            z = add_residual(z, sc, e.res_mul);
          }
          if (leaky) z = act_leaky(z, e.leaky_alpha);
          v[j] = z;
        }
        if (valid) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            reinterpret_cast<float4*>(yrow + c0)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
      }
    }
  } else if (warp >= 8) {
    // ===================== converters: fp32 halo -> three bf16 planes =====================
    const int t = threadIdx.x - 8 * 32;
    int s = 0; uint32_t fph = 0;
    int b = 0; uint32_t pph = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      for (int kc = 0; kc < kchunks; ++kc) {
        mbar_wait_parked(ffull(s), fph);
        mbar_wait_parked(pempty(b), pph ^ 1u);
        const uint8_t* fsrc = sg + SL::F_OFF + s * SL::F_BYTES;
        uint8_t* pdst = sg + SL::P_OFF + b * SL::PSTAGE_BYTES;
        for (int item = t; item < HALO_PX * 2; item += F_CVT_THREADS) {
          const int hp = item >> 1, half = item & 1;
          const float4 a = *reinterpret_cast<const float4*>(fsrc + hp * 64 + half * 32);
          const float4 c = *reinterpret_cast<const float4*>(fsrc + hp * 64 + half * 32 + 16);
          uint32_t hi[4], mid[4], lo[4];
          split3(a.x, a.y, hi[0], mid[0], lo[0]);
          split3(a.z, a.w, hi[1], mid[1], lo[1]);
          split3(c.x, c.y, hi[2], mid[2], lo[2]);
          split3(c.z, c.w, hi[3], mid[3], lo[3]);
          uint8_t* d = pdst + half * SL::K8_BYTES + hp * 16;
          *reinterpret_cast<uint4*>(d) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(d + SL::PLANE_BYTES) = make_uint4(mid[0], mid[1], mid[2], mid[3]);
          *reinterpret_cast<uint4*>(d + 2 * SL::PLANE_BYTES) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
        fence_proxy_async();                         // generic-proxy writes -> visible to the tensor core
        __syncwarp();
        if (lane == 0) { mbar_arrive(pfull(b)); mbar_arrive(fempty(s)); }
        if (++s == F_FSTAGES) { s = 0; fph ^= 1u; }
        if (++b == F_PSTAGES) { b = 0; pph ^= 1u; }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

template <int TH>
int launch_th(const CUtensorMap& mx, const F32Params& p, int grid, cudaStream_t st) {
  auto kern = conv3x3_f32_tc_kernel<TH>;
  constexpr int smem = F32Smem<TH>::TOTAL;
  static_assert(smem <= 232448, "shared memory budget");
  static bool configured = false;
  if (!configured) {
    QNNB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  kern<<<grid, F_THREADS, smem, st>>>(mx, p);
  QNNB_CUDA(cudaGetLastError());
  return QNNB_OK;
}

}  // namespace

bool conv_f32_tc_supported(const qnnb_conv_desc& d, const char** why) {
  if (d.in_kind != QNNB_KIND_F32) { *why = "input is not fp32"; return false; }
  if (d.kh != 3 || d.kw != 3 || d.stride != 1) { *why = "only 3x3 stride 1"; return false; }
  if (d.cin % 16 != 0 || d.cin > F_MAXC || d.cout % 16 != 0 || d.cout > F_MAXC) { *why = "Cin and Cout must be 16, 32, 48 or 64"; return false; }
  if (d.h % 8 != 0 || d.w % 8 != 0) { *why = "spatial size must be a multiple of 8"; return false; }
  if (d.epi.pool != 0) { *why = "no pooling on the fp32 tensor-core path"; return false; }
  if (d.epi.act != QNNB_ACT_NONE && d.epi.act != QNNB_ACT_LEAKY) { *why = "activation must be none or LeakyReLU"; return false; }
  if (d.epi.res_kind != QNNB_KIND_NONE && d.epi.res_kind != QNNB_KIND_F32) { *why = "residual must be fp32"; return false; }
  if ((long long)d.n * (d.h / 8) * (d.w / 8) >= (1ll << 24)) { *why = "batch too large for one launch"; return false; }
  return true;
}

int launch_conv_f32_tc(const qnnb_conv_desc& d, const void* x, const void* w, void* y, cudaStream_t st) {
  EncodeTiledFn encode = get_encode();
  if (!encode) { set_error("conv2d: cuTensorMapEncodeTiled is not available from the driver"); return QNNB_ECUDA; }
  const int TH = (d.h % 16 == 0) ? 16 : 8;
  const int TN = 16 / TH;
  CUtensorMap mx;
  {
    // dimension order (C, W, N, H): the box lands as [row][image][10 px][16 ch]
    cuuint64_t dims[4] = {(cuuint64_t)d.cin, (cuuint64_t)d.w, (cuuint64_t)d.n, (cuuint64_t)d.h};
    cuuint64_t strides[3] = {(cuuint64_t)d.cin * 4, (cuuint64_t)d.h * d.w * d.cin * 4, (cuuint64_t)d.w * d.cin * 4};
    cuuint32_t box[4] = {(cuuint32_t)F_KC, 10u, (cuuint32_t)TN, (cuuint32_t)(TH + 2)};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(&mx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(x), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("conv2d: cuTensorMapEncodeTiled(fp32 halo) failed with %d", (int)r); return QNNB_ECUDA; }
  }
  F32Params p;
  p.n = d.n; p.h = d.h; p.w = d.w; p.cin = d.cin; p.cout = d.cout;
  p.tiles_w = d.w / 8;
  p.tiles_h = d.h / TH;
  p.num_tiles = p.tiles_w * p.tiles_h * ceil_div(d.n, TN);
  p.kchunks = d.cin / F_KC;
  p.fd_w = make_fastdiv(p.tiles_w);
  p.fd_h = make_fastdiv(p.tiles_h);
  p.wpk = (const int8_t*)w;
  p.y = (float*)y;
  p.epi = make_epi(d.epi);
  const int grid = p.num_tiles < sm_count() ? p.num_tiles : sm_count();
  if (TH == 16) return launch_th<16>(mx, p, grid, st);
  return launch_th<8>(mx, p, grid, st);
}

}  // namespace qnnb
