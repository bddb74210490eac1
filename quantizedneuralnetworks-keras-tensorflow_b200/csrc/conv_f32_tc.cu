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
// x = hi + mid + lo (8 significant bits each, by truncation), so  sum x*w = sum hi*w + sum mid*w + sum lo*w  is three bf16 MMAs whose
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
//               all taps' kernels stay resident in shared memory as bf16.  With N this small an MMA occupies the
//               tensor pipe for a few cycles but its accumulate latency is ~100, so a chain of MMAs into ONE
//               accumulator runs at latency (measured: ~110 cycles per MMA for N = 16, 32 and 64 alike).  The MMAs
//               of a tile therefore rotate over NP independent partial accumulators (NP x cout TMEM columns), which
//               the epilogue adds up.
//   warps 4-7   epilogue: tcgen05.ld 16 columns at a time, scale / bias / BN / residual / LeakyReLU in the fixed
//               fp32 op order of common.cuh.  The residual tile of the shortcut branch is TMA-loaded into a ring of
//               [pixel][channel] tiles (64B/128B swizzle: a thread reads its own pixel row without bank conflicts);
//               the result overwrites it in place and leaves through one TMA store per tile, so the epilogue
//               never waits on a global-memory round trip and HBM sees full 128-byte rows.
#include "common.cuh"
#include "tc_ptx.cuh"

#include <cuda.h>
#include <cuda_bf16.h>
#include <string.h>
#include <stdlib.h>

namespace qnnb {

namespace {

using namespace tcx;

// warp roles: 0 = halo TMA, 1 and 3 = MMA issuers (even / odd tiles), 2 = TMEM allocator + shortcut-tile TMA,
// 4..11 = two epilogue groups of four warps (even / odd tiles), 12..17 = converters
constexpr int F_CVT_WARPS = 8;
constexpr int F_CVT_THREADS = F_CVT_WARPS * 32;
constexpr int F_EPI_WARPS = 4;       // per group
constexpr int F_EPI_GROUPS = 2;
constexpr int F_CVT_WARP0 = 4 + F_EPI_GROUPS * F_EPI_WARPS;
constexpr int F_THREADS = (F_CVT_WARP0 + F_CVT_WARPS) * 32;
constexpr int F_KC = 16;             // channels per pipeline unit (= one bf16 MMA K step)
constexpr int F_FSTAGES = 8;         // fp32 halo ring (at most; the host picks the depth that fits)
constexpr int F_RSTAGES = 6;         // residual / output tile ring (at most)
constexpr int F_PSTAGES = 4;         // bf16 plane ring (at most; 4 or 2)
constexpr int F_CVT_GROUPS = 2;      // converter groups (alternate units)
constexpr int F_CVT_GWARPS = F_CVT_WARPS / F_CVT_GROUPS;
constexpr int F_EPI_BAR = 2;         // named barriers 2, 3: one per epilogue group
constexpr int F_ACCS = 4;            // TMEM accumulator slots in flight (at most)
constexpr int F_MAXC = 64;           // channel limit (Cin and Cout)
constexpr int F_W_BYTES = 9 * F_MAXC * F_MAXC * 2;

struct F32Params {
  int n, h, w, cin, cout;
  int tiles_w, tiles_h, num_tiles, kchunks;
  int f_stages, r_stages, p_stages;  // ring depths chosen by the host (all even: a slot always belongs to the same group)
  int in_merged;                     // Cin == 16: the halo box is a 3-D box over (W*C, N, H) -- rows of 640 B instead of 64 B
  unsigned long long* dbg;           // QNNB_TRACE builds: CTA 0 timeline (region = role * 1024 words: [count, (tag<<32|idx, ns)...])
  int exp_mode;                      // diagnostics only (QNNB_K4_EXP): 1 = issue one plane's MMAs, 2 = converters skip the split, 3 = epilogue skips math
  int np, accs, acc_cols;            // partial accumulators per tile, tiles in flight in TMEM, columns per tile (np * cout)
  int off_p, off_w, off_r, off_c, off_bar;   // shared-memory offsets (bytes)
  int r_bytes;                       // one residual / output tile: 128 px x cout floats
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
  static_assert(F_BYTES % 128 == 0, "TMA destination alignment");
};

__device__ __forceinline__ void ftrace(const F32Params& p, int role, int tag, int idx) {
#ifdef QNNB_TRACE
  if (p.dbg != nullptr && blockIdx.x == 0) {
    unsigned long long* reg = p.dbg + role * 1024;
    const unsigned long long k = reg[0];
    if (k < 500) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
      reg[2 + 2 * k] = ((unsigned long long)tag << 32) | (unsigned)idx;
      reg[3 + 2 * k] = t;
      reg[0] = k + 1;
    }
  }
#endif
}

// D[tmem] (+)= A[smem] * B[smem], bf16 x bf16 -> fp32.  The descriptors are passed as (low, high) words: the issuing
// thread is a single thread whose instruction count per MMA is what bounds this kernel (N is tiny), so the loop keeps
// the high words constant and only adds small offsets to the low words (start address field, 16-byte units).
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
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

// x0, x1 -> three bf16x2 words (low half = x0).  Each term is the TOP 16 BITS of what is left (truncation, one PRMT per
// pair): hi = top(x), mid = top(x - hi), lo = top(x - hi - mid).  A bf16 keeps 8 significant bits and every subtraction
// is exact in fp32, so the three terms peel off 8 + 8 + 8 = all 24 significant bits: hi + mid + lo == x exactly
// (for |x| >= 2^-100; below that the residuals become subnormal and the absolute error is < 2^-126 --
// tests/test_kernel_arithmetic_cpu.py).
// (Integer byte-permutes and ANDs instead of cvt.rn.bf16x2: the conversion pipe was the converter warps' bottleneck.)
__device__ __forceinline__ void split3(float x0, float x1, uint32_t& hi, uint32_t& mid, uint32_t& lo) {
  const uint32_t u0 = __float_as_uint(x0), u1 = __float_as_uint(x1);
  hi = __byte_perm(u0, u1, 0x7632);                                  // {top16(x1), top16(x0)}
  const float r0 = __fsub_rn(x0, __uint_as_float(u0 & 0xFFFF0000u));
  const float r1 = __fsub_rn(x1, __uint_as_float(u1 & 0xFFFF0000u));
  const uint32_t v0 = __float_as_uint(r0), v1 = __float_as_uint(r1);
  mid = __byte_perm(v0, v1, 0x7632);
  const float q0 = __fsub_rn(r0, __uint_as_float(v0 & 0xFFFF0000u));
  const float q1 = __fsub_rn(r1, __uint_as_float(v1 & 0xFFFF0000u));
  lo = __byte_perm(__float_as_uint(q0), __float_as_uint(q1), 0x7632);
}

template <int TH>
__global__ void __launch_bounds__(F_THREADS, 1)
conv3x3_f32_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_r,
                      const __grid_constant__ CUtensorMap map_y, const F32Params p) {
  using SL = F32Smem<TH>;
  constexpr int TN = SL::TN;
  constexpr int HALO_PX = SL::HALO_PX;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;        // swizzled residual tiles need 1024 B
  uint8_t* sg = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + p.off_bar;
  auto ffull = [&](int s) { return bar_base + 8u * s; };
  auto fempty = [&](int s) { return bar_base + 8u * (F_FSTAGES + s); };
  // "planes ready" has one barrier array PER MMA ISSUER (a unit is published to the issuer that will consume it): a
  // parity wait can only tell adjacent phases apart, so every waiter must see consecutive phases of its barriers; the
  // two issuers take alternate tiles and would otherwise each see only some of a slot's phases
  auto pfull = [&](int issuer, int b) { return bar_base + 8u * (2 * F_FSTAGES + issuer * F_PSTAGES + b); };
  auto pempty = [&](int b) { return bar_base + 8u * (2 * F_FSTAGES + 2 * F_PSTAGES + b); };
  auto tfull = [&](int a) { return bar_base + 8u * (2 * F_FSTAGES + 3 * F_PSTAGES + a); };
  auto tempty = [&](int a) { return bar_base + 8u * (2 * F_FSTAGES + 3 * F_PSTAGES + F_ACCS + a); };
  auto rfull = [&](int r) { return bar_base + 8u * (2 * F_FSTAGES + 3 * F_PSTAGES + 2 * F_ACCS + r); };
  auto rempty = [&](int r) { return bar_base + 8u * (2 * F_FSTAGES + 3 * F_PSTAGES + 2 * F_ACCS + F_RSTAGES + r); };
  constexpr int NBARS = 2 * F_FSTAGES + 3 * F_PSTAGES + 2 * F_ACCS + 2 * F_RSTAGES;
  const uint32_t tmem_slot = bar_base + 8u * NBARS;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(sg + p.off_bar + 8 * NBARS);
  const bool has_res = p.epi.res_kind == QNNB_KIND_F32;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cout = p.cout;
  const int kchunks = p.kchunks;

  if (threadIdx.x == 0) griddep_launch_dependents();
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
      *reinterpret_cast<uint4*>(sg + p.off_w + (size_t)i * 16) = make_uint4(o[0], o[1], o[2], o[3]);
    }
    float* cst = reinterpret_cast<float*>(sg + p.off_c);
    for (int c = threadIdx.x; c < F_MAXC; c += F_THREADS) {
      const bool ok = c < cout;
      cst[c] = (p.epi.bias != nullptr && ok) ? __ldg(p.epi.bias + c) : 0.f;
      cst[F_MAXC + c] = (p.epi.bn_inv != nullptr && ok) ? __ldg(p.epi.bn_inv + c) : 1.f;
      cst[2 * F_MAXC + c] = (p.epi.bn_inv != nullptr && ok) ? __ldg(p.epi.bn_shift + c) : 0.f;
    }
  }
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&map_x); tma_prefetch_desc(&map_r); tma_prefetch_desc(&map_y); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < F_FSTAGES; ++s) { mbar_init(ffull(s), 1); mbar_init(fempty(s), F_CVT_GWARPS); }
    for (int r = 0; r < F_RSTAGES; ++r) { mbar_init(rfull(r), 1); mbar_init(rempty(r), 1); }
    for (int b = 0; b < F_PSTAGES; ++b) { mbar_init(pfull(0, b), F_CVT_GWARPS); mbar_init(pfull(1, b), F_CVT_GWARPS); mbar_init(pempty(b), 1); }
    for (int a = 0; a < F_ACCS; ++a) { mbar_init(tfull(a), 1); mbar_init(tempty(a), F_EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async();                   // resident kernel written through the generic proxy, read by the tensor core
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  griddep_wait();                        // activations / shortcut / output buffers only after the previous kernel completed

  auto decode = [&](int tile, int& n0, int& h0, int& w0) {
    const int q = fdiv(tile, p.fd_w);
    w0 = (tile - q * p.tiles_w) * 8;
    const int q2 = fdiv(q, p.fd_h);
    h0 = (q - q2 * p.tiles_h) * TH;
    n0 = q2 * TN;
  };

  if (warp == 0) {
    // ===================== TMA: fp32 halo tile per (tile, chunk) =====================
    if (elect_one()) {
      int s = 0; uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        int n0, h0, w0;
        decode(tile, n0, h0, w0);
        for (int kc = 0; kc < kchunks; ++kc) {
          mbar_wait(fempty(s), ph ^ 1u);
          ftrace(p, 0, 1, tile);                      // TMA: halo slot free, load issued
          mbar_expect_tx(ffull(s), SL::F_BYTES);
          if (p.in_merged) tma_load_3d(smem_base + 0 + s * SL::F_BYTES, &map_x, ffull(s), (w0 - 1) * F_KC, n0, h0 - 1);
          else tma_load_4d(smem_base + 0 + s * SL::F_BYTES, &map_x, ffull(s), kc * F_KC, w0 - 1, n0, h0 - 1);
          if (++s == p.f_stages) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 2) {
    // ===================== TMA: shortcut (residual) / output tile ring =====================
    // its own thread: a single-threaded loop pays a few hundred ns per barrier wait, and this wait depends on the
    // epilogue (buffer release), which must not hold back the halo loads of the tiles behind it
    if (elect_one()) {
      int r = 0; uint32_t rph = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        int n0, h0, w0;
        decode(tile, n0, h0, w0);
        // acquire the buffer, then either fill it or just hand it over
        mbar_wait(rempty(r), rph ^ 1u);
        ftrace(p, 0, 2, tile);                        // TMA: residual slot free
        if (has_res) {
          mbar_expect_tx(rfull(r), p.r_bytes);
          if (cout == 16) tma_load_4d(smem_base + p.off_r + r * p.r_bytes, &map_r, rfull(r), 0, w0 >> 1, n0, h0);
          else
            for (int sub = 0; sub * 32 < cout; ++sub)
              tma_load_4d(smem_base + p.off_r + r * p.r_bytes + sub * (128 * 128), &map_r, rfull(r), sub * 32, w0, n0, h0);
        } else {
          mbar_arrive(rfull(r));
        }
        if (++r == p.r_stages) { r = 0; rph ^= 1u; }
      }
    }
  } else if (warp == 1 || warp == 3) {
    // ===================== MMA issuers =====================
    // The issue loop is single-threaded and, with MMAs this small, it IS the critical path (timeline trace: ~70 cycles
    // per MMA plus ~0.4 us of barrier waits per tile).  Two warps therefore issue alternate tiles: warp 1 the even
    // iterations, warp 3 the odd ones; each walks the whole unit sequence to keep its ring indices in step.
    const int my_par = warp == 1 ? 0 : 1;
    if (elect_one()) {
      // instruction descriptor: dense, F32 accumulate, BF16 x BF16, both K-major, N = cout, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(cout >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      // un-swizzled K-major descriptors: low word = start address >> 4 | (K-chunk stride >> 4) << 16,
      //                                  high word = 8-row group stride >> 4 | version 1 (bit 46)
      // A: 16 groups of 8 pixels, one halo row (10 px x 16 B) apart; the two 16-byte K chunks K8_BYTES apart
      // B: 8-channel groups 128 B apart; the two K chunks cout * 16 B apart
      const uint32_t a_hi = (160u >> 4) | (1u << 14);
      const uint32_t b_hi = (128u >> 4) | (1u << 14);
      const uint32_t a_lo0 = (((smem_base + p.off_p) & 0x3FFFFu) >> 4) | ((uint32_t)(SL::K8_BYTES >> 4) << 16);
      const uint32_t b_lo0 = (((smem_base + p.off_w) & 0x3FFFFu) >> 4) | ((uint32_t)cout << 16);
      const uint32_t w_tap = (uint32_t)(kchunks * 2 * cout);      // 16-byte units between consecutive taps of the kernel
      const bool np3 = p.np == 3;
      uint32_t uses = 0;                               // four 8-bit counters: own uses of each plane-ring slot
      int it = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        if ((it & 1) != my_par) continue;              // the other issuer's tile
        const int acc = it % p.accs;
        const uint32_t acc_phase = (uint32_t)(it / p.accs) & 1u;
        mbar_wait(tempty(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.acc_cols);
        for (int kc = 0; kc < kchunks; ++kc) {
          const int b = (it * kchunks + kc) % p.p_stages;              // ring slot of this unit
          mbar_wait(pfull(my_par, b), (uses >> (8 * b)) & 1u);         // parity = how often WE have used this slot
          uses += 1u << (8 * b);
          ftrace(p, 2, 6, tile);                      // MMA: planes ready
          tc_fence_after();
          const uint32_t a_unit = a_lo0 + (uint32_t)(b * (SL::PSTAGE_BYTES >> 4));
          const uint32_t b_unit = b_lo0 + (uint32_t)(kc * 2 * cout);
          const uint32_t later = kc > 0 ? 1u : 0u;
#pragma unroll
          for (int split = 0; split < 3; ++split) {
            if (p.exp_mode == 1 && split > 0) break;
            // np == 3: one partial accumulator per bf16 plane (three independent MMA chains)
            const uint32_t d = d_tmem + (np3 ? (uint32_t)(split * cout) : 0u);
            const uint32_t first = (split == 0 || np3) ? later : 1u;
            uint32_t b_lo = b_unit;
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              const int r = tap / 3, sft = tap - 3 * r;
              umma_bf16(d, a_unit + (uint32_t)(split * (SL::PLANE_BYTES >> 4) + r * TN * 10 + sft), a_hi, b_lo, b_hi, idesc,
                        tap == 0 ? first : 1u);
              b_lo += w_tap;
            }
          }
          umma_commit(pempty(b));
        }
        umma_commit(tfull(acc));
        ftrace(p, 2, 7, tile);                        // MMA: tile issued
      }
    }
  } else if (warp >= 4 && warp < F_CVT_WARP0) {
    // ===================== epilogue: thread = pixel =====================
    const int quarter = warp & 3;
    const int L = quarter * 32 + lane;               // TMEM lane = pixel of the tile = row of the residual tile
    const Epi& e = p.epi;
    const float* cst = reinterpret_cast<const float*>(sg + p.off_c);
    const bool has_bias = e.bias != nullptr, has_bn = e.bn_inv != nullptr;
    const bool leaky = e.act == QNNB_ACT_LEAKY;
    const int egrp = (warp - 4) >> 2;                // epilogue group: tiles with (it & 1) == egrp
    const bool leader = (quarter == 0 && lane == 0);
    // residual / output tile: 128-byte rows, SWIZZLE_128B.  Cout >= 32: row = pixel, 32 channels per sub-tile;
    // Cout = 16: row = two neighbouring pixels (the tensor map views the image as [W/2][32 floats])
    const bool two_px = cout == 16;
    const int trow = two_px ? (L >> 1) : L;
    const uint32_t xr = (uint32_t)(trow & 7);        // 16-byte chunk XOR of this row
    int it = 0;
    int r = 0; uint32_t rph = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      if ((it & 1) != egrp) {                        // the other group's tile (r_stages is even: its buffers are never ours)
        if (++r == p.r_stages) { r = 0; rph ^= 1u; }
        continue;
      }
      int n0, h0, w0;
      decode(tile, n0, h0, w0);
      const int acc = it % p.accs;
      const uint32_t acc_phase = (uint32_t)(it / p.accs) & 1u;
      uint8_t* rbuf = sg + p.off_r + r * p.r_bytes;
      mbar_wait_parked(rfull(r), rph);               // shortcut tile landed (or: buffer is ours)
      if (leader) ftrace(p, 3, 8, tile);
      mbar_wait_parked(tfull(acc), acc_phase);
      if (leader) ftrace(p, 3, 9, tile);             // epilogue: accumulator complete
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * p.acc_cols);
      for (int c0 = 0; c0 < cout; c0 += 16) {
        if (p.exp_mode == 3) { if (c0 == 0) { tc_fence_before(); __syncwarp(); if (lane == 0) mbar_arrive(tempty(acc)); } continue; }
        uint8_t* rowp = rbuf + (c0 >> 5) * (128 * 128) + trow * 128;
        const uint32_t ch0 = two_px ? (uint32_t)((L & 1) * 4) : (uint32_t)((c0 & 31) >> 2);   // first 16-byte chunk of this block
        float4 rs[4];
        if (has_res) {
#pragma unroll
          for (int j = 0; j < 4; ++j) rs[j] = *reinterpret_cast<const float4*>(rowp + (((ch0 + j) ^ xr) << 4));
        }
        float v[16];
        __syncwarp();
        tmem_ld16(taddr + (uint32_t)c0, v);
        for (int part = 1; part < p.np; ++part) {      // add the other partial accumulators
          float u[16];
          tmem_ld16(taddr + (uint32_t)(part * cout + c0), u);
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __fadd_rn(v[j], u[j]);
        }
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
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<float4*>(rowp + (((ch0 + j) ^ xr) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
      if (leader) ftrace(p, 3, 13, tile);            // epilogue: math done
      fence_proxy_async();                           // result tile -> visible to the TMA store
      named_bar_sync(F_EPI_BAR + egrp, F_EPI_WARPS * 32);
      if (leader) {
        if (two_px) tma_store_4d(&map_y, smem_u32(rbuf), 0, w0 >> 1, n0, h0);
        else
          for (int sub = 0; sub * 32 < cout; ++sub)
            tma_store_4d(&map_y, smem_u32(rbuf + sub * (128 * 128)), sub * 32, w0, n0, h0);
        tma_store_commit();
        ftrace(p, 3, 10, tile);                       // epilogue: tile stored
        // release the buffer as soon as the store has read it (this group has a whole tile of slack before it is needed)
        tma_store_wait_read();
        mbar_arrive(rempty(r));
      }
      if (++r == p.r_stages) { r = 0; rph ^= 1u; }
    }
    if (leader) tma_store_wait_all();
  } else if (warp >= F_CVT_WARP0) {
    // ===================== converters: fp32 halo -> three bf16 planes =====================
    // Two groups of four warps convert alternate units: the per-unit fixed cost (two barrier waits, the proxy fence,
    // the arrivals) is paid once per 2.8 items per thread instead of 1.4, and two units are in flight.  All ring depths
    // are even, so a ring slot always belongs to the same group and every group sees consecutive phases of its slots.
    const int cg = (warp - F_CVT_WARP0) / F_CVT_GWARPS;
    const int t = threadIdx.x - (F_CVT_WARP0 + cg * F_CVT_GWARPS) * 32;       // 0..127 within the group
    constexpr int GT = F_CVT_GWARPS * 32;
    int unit = 0, it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      for (int kc = 0; kc < kchunks; ++kc, ++unit) {
        if ((unit & 1) != cg) continue;
        const int s = unit % p.f_stages, b = unit % p.p_stages;
        const uint32_t fph = (uint32_t)(unit / p.f_stages) & 1u, pph = (uint32_t)(unit / p.p_stages) & 1u;
        mbar_wait_parked(ffull(s), fph);
        if (t == 0) ftrace(p, 1, 3, tile);            // converter: halo landed
        mbar_wait_parked(pempty(b), pph ^ 1u);
        if (t == 0) ftrace(p, 1, 4, tile);            // converter: plane slot free
        const uint8_t* fsrc = sg + 0 + s * SL::F_BYTES;
        uint8_t* pdst = sg + p.off_p + b * SL::PSTAGE_BYTES;
        // (pixel, 8-channel half) items, two per pass: their shared-memory loads are issued before the first conversion
        constexpr int PASSES = (HALO_PX * 2 + 2 * GT - 1) / (2 * GT);
#pragma unroll 1
        for (int ps = 0; ps < PASSES; ++ps) {
          float4 va[2], vc[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int item = t + (2 * ps + u) * GT;
            if (item < HALO_PX * 2) {
              const int hp = item >> 1, half = item & 1;
              va[u] = *reinterpret_cast<const float4*>(fsrc + hp * 64 + half * 32);
              vc[u] = *reinterpret_cast<const float4*>(fsrc + hp * 64 + half * 32 + 16);
            }
          }
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int item = t + (2 * ps + u) * GT;
            if (item < HALO_PX * 2 && p.exp_mode != 2) {
              const int hp = item >> 1, half = item & 1;
              uint32_t hi[4], mid[4], lo[4];
              split3(va[u].x, va[u].y, hi[0], mid[0], lo[0]);
              split3(va[u].z, va[u].w, hi[1], mid[1], lo[1]);
              split3(vc[u].x, vc[u].y, hi[2], mid[2], lo[2]);
              split3(vc[u].z, vc[u].w, hi[3], mid[3], lo[3]);
              uint8_t* d = pdst + half * SL::K8_BYTES + hp * 16;
              *reinterpret_cast<uint4*>(d) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
              *reinterpret_cast<uint4*>(d + SL::PLANE_BYTES) = make_uint4(mid[0], mid[1], mid[2], mid[3]);
              *reinterpret_cast<uint4*>(d + 2 * SL::PLANE_BYTES) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
          }
        }
        if (t == 0) ftrace(p, 1, 12, tile);           // converter: planes written
        fence_proxy_async();                         // generic-proxy writes -> visible to the tensor core
        __syncwarp();
        if (lane == 0) { mbar_arrive(pfull(it & 1, b)); mbar_arrive(fempty(s)); }
        if (t == 0) ftrace(p, 1, 5, tile);            // converter: planes published
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int TH>
int launch_th(const CUtensorMap& mx, const CUtensorMap& mr, const CUtensorMap& my, F32Params p, int grid, cudaStream_t st) {
  using SL = F32Smem<TH>;
  auto kern = conv3x3_f32_tc_kernel<TH>;
  // shared-memory plan: bf16 planes (fixed depth), resident kernel, then as many fp32 halo and residual / output stages
  // as fit (both rings need >= 2)
  const int w_bytes = 9 * p.cin * p.cout * 2;
  // plane ring: 4 slots (two per converter group) when the minimum fp32 / residual rings still fit, else 2
  const int other = w_bytes + 3 * F_MAXC * 4 + 512 + 1024 + 3 * 1024;                                     // + alignment slack
  const int r_one = p.cout == 16 ? 64 * 128 : ceil_div(p.cout, 32) * 128 * 128;
  p.p_stages = (232448 - other - 4 * SL::PSTAGE_BYTES >= 2 * SL::F_BYTES + 2 * r_one) ? 4 : 2;
  const int fixed = p.p_stages * SL::PSTAGE_BYTES + other;
  const int budget = 232448 - fixed;
  {
    p.r_bytes = p.cout == 16 ? 64 * 128 : ceil_div(p.cout, 32) * 128 * 128;   // whole swizzled sub-tiles (Cout = 48: the second is half empty)
  }
  p.f_stages = getenv("QNNB_K4_FST") ? atoi(getenv("QNNB_K4_FST")) : 6;
  p.r_stages = getenv("QNNB_K4_RST") ? atoi(getenv("QNNB_K4_RST")) : 6;
  if (p.f_stages < 2 || p.f_stages > F_FSTAGES) p.f_stages = 6;
  if (p.r_stages < 2 || p.r_stages > F_RSTAGES) p.r_stages = 6;
  p.r_stages &= ~1;                                // even: a buffer always belongs to the same epilogue group
  p.f_stages &= ~1;                                // even: a slot always belongs to the same converter group
  while (p.f_stages * SL::F_BYTES + p.r_stages * p.r_bytes > budget) {
    if (p.r_stages > 2 && p.r_stages * p.r_bytes >= p.f_stages * SL::F_BYTES) p.r_stages -= 2;
    else if (p.f_stages > 2) p.f_stages -= 2;
    else if (p.r_stages > 2) p.r_stages -= 2;
    else { set_error("conv2d: fp32 tensor-core kernel does not fit shared memory (cin=%d cout=%d)", p.cin, p.cout); return QNNB_EINVAL; }
  }
  // partial accumulators: as many independent MMA chains as TMEM holds with >= 2 tiles in flight (NP x cout columns each)
  {
    p.dbg = get_trace_buffer();
    p.exp_mode = getenv("QNNB_K4_EXP") ? atoi(getenv("QNNB_K4_EXP")) : 0;
    int np = 1;
    if (const char* e = getenv("QNNB_K4_NP")) np = atoi(e) == 3 ? 3 : 1;
    p.np = np;
    p.acc_cols = np * p.cout;
    p.accs = 4 * p.acc_cols <= 512 ? 4 : 2;         // even: each accumulator slot always belongs to the same issuer warp
  }
  auto up = [](int v, int a) { return (v + a - 1) / a * a; };
  p.off_p = up(p.f_stages * SL::F_BYTES, 128);
  p.off_w = up(p.off_p + p.p_stages * SL::PSTAGE_BYTES, 128);
  p.off_r = up(p.off_w + w_bytes, 1024);
  p.off_c = p.off_r + p.r_stages * p.r_bytes;
  p.off_bar = up(p.off_c + 3 * F_MAXC * 4, 16);
  const int smem = p.off_bar + 512 + 1024;
  if (smem > 232448) { set_error("conv2d: fp32 tensor-core kernel shared-memory plan overflow (%d B)", smem); return QNNB_EINVAL; }
  static tcx::SmemConfigured once;                // largest size this instantiation was configured for, per device
  QNNB_CUDA(once.ensure(kern, smem));
  QNNB_CUDA(launch_pdl(kern, dim3(grid), dim3(F_THREADS), (size_t)smem, st, mx, mr, my, p));
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
  const bool in_merged = d.cin == F_KC;
  if (in_merged) {
    // Cin = 16: a pixel IS one 64-byte chunk, so W and C merge into one dimension and the box has 640-byte rows
    // (10 pixels); the TMA engine's cost is per box row, and 64-byte rows left it the bottleneck of the kernel
    cuuint64_t dims[3] = {(cuuint64_t)d.cin * d.w, (cuuint64_t)d.n, (cuuint64_t)d.h};
    cuuint64_t strides[2] = {(cuuint64_t)d.h * d.w * d.cin * 4, (cuuint64_t)d.w * d.cin * 4};
    cuuint32_t box[3] = {(cuuint32_t)(10 * F_KC), (cuuint32_t)TN, (cuuint32_t)(TH + 2)};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(&mx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(x), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("conv2d: cuTensorMapEncodeTiled(fp32 halo, merged) failed with %d", (int)r); return QNNB_ECUDA; }
  } else {
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
  // residual (shortcut) and output tiles: 128-byte rows, SWIZZLE_128B.  Cout >= 32: [pixel][32 channels] sub-tiles;
  // Cout = 16: the image is viewed as [W/2][32 floats] so that a row holds two pixels
  const bool two_px = d.cout == 16;
  CUtensorMap mr, my;
  memset(&mr, 0, sizeof(mr));
  for (int which = 0; which < 2; ++which) {
    void* base = which == 0 ? const_cast<void*>(d.epi.residual) : y;
    if (which == 0 && d.epi.res_kind != QNNB_KIND_F32) continue;
    const int wd = two_px ? d.w / 2 : d.w, cd = two_px ? 32 : d.cout;
    cuuint64_t dims[4] = {(cuuint64_t)cd, (cuuint64_t)wd, (cuuint64_t)d.n, (cuuint64_t)d.h};
    cuuint64_t strides[3] = {(cuuint64_t)cd * 4, (cuuint64_t)d.h * d.w * d.cout * 4, (cuuint64_t)d.w * d.cout * 4};
    cuuint32_t box[4] = {32u, two_px ? 4u : 8u, (cuuint32_t)TN, (cuuint32_t)TH};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(which == 0 ? &mr : &my, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("conv2d: cuTensorMapEncodeTiled(fp32 %s) failed with %d", which == 0 ? "residual" : "output", (int)r); return QNNB_ECUDA; }
  }
  F32Params p;
  p.n = d.n; p.h = d.h; p.w = d.w; p.cin = d.cin; p.cout = d.cout;
  p.tiles_w = d.w / 8;
  p.tiles_h = d.h / TH;
  p.num_tiles = p.tiles_w * p.tiles_h * ceil_div(d.n, TN);
  p.kchunks = d.cin / F_KC;
  p.in_merged = in_merged ? 1 : 0;
  p.fd_w = make_fastdiv(p.tiles_w);
  p.fd_h = make_fastdiv(p.tiles_h);
  p.wpk = (const int8_t*)w;
  p.y = (float*)y;
  p.epi = make_epi(d.epi);
  const int grid = p.num_tiles < grid_sms(d.max_ctas) ? p.num_tiles : grid_sms(d.max_ctas);
  if (TH == 16) return launch_th<16>(mx, mr, my, p, grid, st);
  return launch_th<8>(mx, mr, my, p, grid, st);
}

}  // namespace qnnb
