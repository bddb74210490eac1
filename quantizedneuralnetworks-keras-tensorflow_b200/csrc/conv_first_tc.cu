// K5 -- first layer: 3x3 stride-1 SAME convolution of uint8 RGB images (Cin = 3, 32 pixels wide) on the 5th-gen
// tensor cores, fused with bias + BatchNormalization + quantized_tanh / binary_tanh (+ 2x2 MaxPooling2D).
//
// Stands in for QuantizedConv2D.call / BinaryConv2D.call (layers/quantized_layers.py:164-194,
// layers/binary_layers.py:160-187) on the network input, followed by the BN / Activation / MaxPooling2D of
// models/vgg.py:15-23 (first conv of block A) -- the layer that reads the image (utils/load_data.py:40: uint8 / 255).
//
// Why a special kernel: three channel bytes are too narrow for a TMA box or a UMMA K step, and an explicit im2col
// tile costs more instructions than the arithmetic it feeds (round 1: the epilogue warps waited on the im2col
// producers a third of the time).  Here NO im2col tile is ever built:
//
//   * The rows of the image tile (with a one-pixel halo, zero outside the image = SAME padding) are expanded
//     RGB -> RGBX once into shared memory, pixel p of a row at byte 4*(p+1): a "block" of four output pixels
//     4j..4j+3 reads the 32 bytes of pixels 4j-1..4j+6, and consecutive blocks start 16 bytes apart -- exactly the
//     row pitch of an un-swizzled K-major UMMA core matrix.  So the B operand of the MMA (N = blocks) IS the staged
//     halo tile, addressed through a shared-memory descriptor: 8 blocks of an image row form one 8-row core-matrix
//     group, the next group is the next image row (SBO = row pitch), the second 16-byte K chunk comes from a copy of
//     the tile shifted by 16 bytes (LBO = plane size), and filter row r just shifts the start address by r rows.
//   * The horizontal taps move into the A operand: for output position i (0..3) inside a block, channel c and filter
//     row r the 32-byte A row holds w[c][r][s] at byte 4*(i+s) and zeros elsewhere (a Toeplitz expansion of the
//     2.3 KB kernel, built once per CTA).  M = 64 channels x 4 positions = two 128-row MMA tiles: tile E holds the
//     even positions, tile O the odd ones, with identical lane order (lane = 64*(i/2) + c), so the two horizontal
//     neighbours of a 2x2 pooling window are the SAME lane of two accumulators and the vertical ones are 8 columns
//     apart: pooling stays an in-register max.
//   * One tile = 16 image rows = N 128 blocks; per 64-channel group ("sub-tile") 2 x 3 tcgen05.mma.kind::i8
//     (M 128, N 128, K 32), u8 activations x s8 kernel levels -> int32 in TMEM, 2 x 128 columns, double buffered.
//
// Warp roles (round 2: the first version ran staging and epilogue on the same 16 warps in lockstep on one tile; the
// device timeline showed ~0.8 us of serialised barrier / fence / wait latencies per tile with nothing to overlap them):
//   warp 0        one elected thread: 1-D bulk copies of the raw image rows (18 contiguous rows per tile) into a ring
//                 six tiles deep, and the MMAs;
//   warps 1..16   epilogue, in two groups of eight that work on ALTERNATE sub-tiles (group g owns accumulator buffer
//                 g), so the waits of one group sit under the arithmetic of the other; inside a group a warp owns a
//                 lane quarter (hardware: warp id % 4) and one half of the tile's rows, processed as two passes of
//                 four image rows; the four warps of a row half share staging buffers and one TMA store per pass;
//   warps 17..19  halo staging: raw RGB rows -> RGBX planes, three halo buffers, running ahead of the MMAs.
//
// Epilogue arithmetic: thread = one (channel, position pair).  Fixed fp32 op order of common.cuh on PAIRS of
// accumulators with packed FFMA2 (tc_ptx.cuh); where the BN slope of a channel is negative the kernel levels of that
// channel are negated while the A tile is built (acc -> -acc, s -> -s: float(acc)*s is unchanged bit for bit), so
// pooling is always ONE max tree (a kernel level of -128 in such a channel cannot be negated: the CTA then takes a
// general min / max path).  Levels are staged as [pixel][64 channels] bytes and leave by TMA.
#include "common.cuh"
#include "tc_ptx.cuh"

#include <cuda.h>
#include <string.h>
#include <type_traits>

namespace qnnb {

namespace {

using namespace tcx;

constexpr int F_EPI_WARPS = 16;                      // two groups of eight
constexpr int F_STAGE_WARPS = 3;
constexpr int F_THREADS = (1 + F_EPI_WARPS + F_STAGE_WARPS) * 32;   // 640
constexpr int F_ROWS = 16;                           // image rows per tile
constexpr int F_N = 128;                             // UMMA N: 16 rows x 8 blocks of 4 pixels
constexpr int F_CG = 64;                             // channels per group (two 128-lane MMA tiles: even / odd positions)
constexpr int F_PITCH = 160;                         // bytes per staged halo row: pixel p (-1..38) at byte 4*(p+1)
constexpr int F_HROWS = F_ROWS + 2;
constexpr int F_PLANE = F_HROWS * F_PITCH;           // plane 0; plane 1 is the same tile shifted by 16 bytes
constexpr int F_HALO = 2 * F_PLANE;
constexpr int F_NHALO = 3;                           // halo buffers (staging runs ahead of the MMAs)
constexpr int F_ABLK = 128 * 32;                     // one A block: 128 rows x 32 K bytes (per group, parity, filter row)
constexpr int F_MAX_GROUPS = 4;                      // Cout <= 256
constexpr int F_RAW = F_HROWS * 96;                  // raw RGB rows of one tile (contiguous in the image)
constexpr int F_RING = 6;                            // raw tiles in flight (1-D bulk copies, issued by the MMA thread)
constexpr int F_ITEMS = F_HROWS * 8;                 // staging items per tile: (halo row, group of 4 pixels)

struct FirstParams {
  const uint8_t* x;
  const int8_t* wpk;         // packed kernel [cout][3][3][4] (K0, QNNB_WFMT_I8 with cin_pad = 4)
  void* y;
  int n, h, cout, groups, tiles_h, num_tiles, out_pitch;
  FastDiv fd_h, fd_g;
  int exp;                   // TRACE builds: diagnostic bit mask (QNNB_K5_EXP, tools/k5_probe.py): 1 no epilogue math, 2 no MMA, 4 no TMA store, 8 no TMEM loads
  unsigned long long* tr;    // TRACE builds: event buffer (qnnb_debug_set_trace), else NULL
  float neg_zero, one;       // -0.0f and 1.0f as RUN-TIME values (see qaffine2: keeps ptxas from re-fusing the fma chain)
  float qlo, qhi;            // clamp bounds -qm, qm - 1 (constant-bank operands of the min / max)
  Epi epi;
};

template <bool POOL, bool OUT_F32>
struct FirstSmem {
  // one staging buffer = one TMA store of one (group, row half): POOL: the row half's 4 pooled rows of a sub-tile;
  // otherwise one pass (4 image rows); 64 channel bytes per pixel
  static constexpr int UNIT_BYTES = (POOL ? 4 * 16 : 4 * 32) * F_CG;
  // 2 groups x 2 row halves x 2 alternating buffers; the region also holds the prologue's copy of the kernel words
  static constexpr int STG_BYTES = OUT_F32 ? F_MAX_GROUPS * F_CG * 9 * 4 : 2 * 2 * 2 * UNIT_BYTES;
  static_assert(STG_BYTES >= F_MAX_GROUPS * F_CG * 9 * 4, "kernel words alias the staging buffers");
};

// device-side timeline of CTA 0 (make TRACE=1 + qnnb_debug_set_trace; tools/k5_trace.py): per-warp event regions, event
// counter in a register (fire-and-forget stores)
#ifdef QNNB_TRACE
#define ftrace(p, tag, idx)                                                                          \
  do {                                                                                               \
    if ((p).tr != nullptr && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && trk__ < 510) {            \
      const unsigned long long t__ = (unsigned long long)clock64();   /* SM cycles: cheap (CS2R) */   \
      *reinterpret_cast<ulonglong2*>((p).tr + (threadIdx.x >> 5) * 1024 + 2 + 2 * trk__) =           \
          make_ulonglong2(((unsigned long long)(tag) << 32) | (unsigned)(idx), t__);                 \
      ++trk__;                                                                                       \
    }                                                                                                \
  } while (0)
#define F_EXP(p, bit) (((p).exp & (bit)) != 0)
#else
#define ftrace(p, tag, idx) do { } while (0)
#define F_EXP(p, bit) false
#endif

// the fixed pipeline on a pair of accumulators: ((float(acc) * s + bias) * (inv*qm)) + shift*qm, every step one
// IEEE round-to-nearest operation (see tc_ptx.cuh fma2).  nz / one are -0 / 1 pairs read from the kernel parameters.
struct PairConst { uint64_t s, a, b, c, nz, one; };

// NP pairs at once, stage by stage (every stage NP independent packed operations): written this way because a pair-at-
// a-time loop compiles to ONE dependent chain per thread (~12 instructions deep, ~60 cycles per pair), which is what the
// epilogue warps then spend their time on.  in: (m0[i], m1[i]) accumulator pairs; out: levels (low byte significant).
template <bool SIGN, int NP>
__device__ __forceinline__ void affine_levels(const int (&m0)[NP], const int (&m1)[NP], int (&l0)[NP], int (&l1)[NP],
                                              const PairConst& q, float qlo, float qhi) {
  uint64_t f[NP];
#pragma unroll
  for (int i = 0; i < NP; ++i) f[i] = pack2((float)m0[i], (float)m1[i]);       // cvt.rn
#pragma unroll
  for (int i = 0; i < NP; ++i) f[i] = fma2(f[i], q.s, q.nz);                   // * s
#pragma unroll
  for (int i = 0; i < NP; ++i) f[i] = fma2(f[i], q.one, q.a);                  // + bias
#pragma unroll
  for (int i = 0; i < NP; ++i) f[i] = fma2(f[i], q.b, q.nz);                   // * inv * qm
#pragma unroll
  for (int i = 0; i < NP; ++i) f[i] = fma2(f[i], q.one, q.c);                  // + shift * qm     == z * qm
  if constexpr (SIGN) {
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      float z0, z1;
      unpack2(f[i], z0, z1);
      l0[i] = act_sign(z0) ? 1 : -1;
      l1[i] = act_sign(z1) ? 1 : -1;
    }
  } else {
    // clamp, then round to nearest even by one add of 1.5 * 2^23 (common.cuh quant_scaled); low byte = level
    const uint64_t magic = pack2(12582912.f, 12582912.f);
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      float z0, z1;
      unpack2(f[i], z0, z1);
      f[i] = pack2(fminf(fmaxf(z0, qlo), qhi), fminf(fmaxf(z1, qlo), qhi));
    }
#pragma unroll
    for (int i = 0; i < NP; ++i) f[i] = fma2(f[i], q.one, magic);
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      float z0, z1;
      unpack2(f[i], z0, z1);
      l0[i] = __float_as_int(z0);
      l1[i] = __float_as_int(z1);
    }
  }
}

// PITCH: bytes per staged pixel = channels per TMA-store box (64, or 32 for a 32-channel layer); a compile-time
// constant so that every staging store is base + immediate
template <bool POOL, bool OUT_F32, bool SIGN, int PITCH>
__global__ void __launch_bounds__(F_THREADS, 1)
conv3x3_c3_direct_tc_kernel(const __grid_constant__ CUtensorMap map_y, const FirstParams p) {
  using SL = FirstSmem<POOL, OUT_F32>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sg = smem_raw + (smem_base - smem_u32(smem_raw));
  const int a_bytes = p.groups * 6 * F_ABLK;
  const int halo_off = a_bytes;                                    // F_NHALO halo buffers (each: plane 0 + plane 1)
  const int cst_off = halo_off + F_NHALO * F_HALO;                 // per-channel constants: float4 {s, a, b, c}
  const int raw_off = cst_off + p.groups * F_CG * 16;              // ring of raw RGB tiles (bulk-copy destinations)
  const int stg_off = (raw_off + F_RING * F_RAW + 1023) & ~1023;
  const int bar_off = stg_off + SL::STG_BYTES;
  const uint32_t bar_base = smem_base + bar_off;
  auto hfull = [&](int b) { return bar_base + 8u * b; };           // 0..2
  auto hempty = [&](int b) { return bar_base + 8u * (3 + b); };    // 3..5
  auto tfull = [&](int a) { return bar_base + 8u * (6 + a); };     // 6, 7
  auto tempty = [&](int a) { return bar_base + 8u * (8 + a); };    // 8, 9
  auto rfull = [&](int s) { return bar_base + 8u * (10 + s); };    // 10..15
  const uint32_t tmem_slot = bar_base + 128u;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(sg + bar_off + 128);
  int* flag_gen = reinterpret_cast<int*>(sg + bar_off + 136);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const Epi& e = p.epi;
  const int G = (int)gridDim.x;

#ifdef QNNB_TRACE
  int trk__ = 0;
#endif
  ftrace(p, 1, 0);
  // local tile k of this CTA is tile blockIdx.x + k * G
  const int ntl = (p.num_tiles - (int)blockIdx.x + G - 1) / G;
  // raw rows of local tile k -> ring slot k % F_RING: the 18 image rows a tile needs are contiguous in the image, so
  // one 1-D bulk copy (clipped at the top / bottom edge) brings them in (issued by the elected thread of warp 0)
  auto fetch = [&](int k) {
    if (k >= ntl || F_EXP(p, 256)) return;
    const int tile = blockIdx.x + k * G;
    const int img = fdiv(tile, p.fd_h);
    const int top = (tile - img * p.tiles_h) * F_ROWS - 1;
    const int lo = max(top, 0), hi = min(top + F_HROWS, p.h);
    const int slot = k % F_RING;
    const uint32_t bytes = (uint32_t)(hi - lo) * 96u;
    mbar_expect_tx(rfull(slot), bytes);
    bulk_load_1d(smem_base + raw_off + slot * F_RAW + (lo - top) * 96, p.x + ((long long)img * p.h + lo) * 96, bytes, rfull(slot));
  };

  // ---- prologue.  Warp 0 sets up the barriers and TMEM, then -- as soon as the previous kernel has completed -- starts
  // the first raw-row copies; meanwhile warps 1..19 build the A operand and the per-channel constants (independent of
  // the previous kernel), synchronising among themselves with a named barrier so that nobody waits for warp 0.
  constexpr int F_WORKERS = F_THREADS - 32;
  uint32_t* wsm = reinterpret_cast<uint32_t*>(sg + stg_off);     // the packed kernel words, [cout][9]
  bool neg_ok = false;
  if (warp == 0) {
    if (lane == 0) {
      griddep_launch_dependents();
      tma_prefetch_desc(&map_y);
      for (int b = 0; b < F_NHALO; ++b) { mbar_init(hfull(b), F_STAGE_WARPS); mbar_init(hempty(b), 1); }
      for (int a = 0; a < 2; ++a) { mbar_init(tfull(a), 1); mbar_init(tempty(a), F_EPI_WARPS / 2); }
      for (int s = 0; s < F_RING; ++s) mbar_init(rfull(s), 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
    if (lane == 0) {
      griddep_wait();                    // the input images only after the previous kernel has completed
      for (int k = 0; k < F_RING; ++k) fetch(k);
    }
    __syncwarp();
  } else {
    const int wt = threadIdx.x - 32;
    // one round of global loads, all in flight together: kernel words and per-channel constants
    const uint32_t* w32 = reinterpret_cast<const uint32_t*>(p.wpk);
    const int nw = p.cout * 9;
    constexpr int WPT = (F_MAX_GROUPS * F_CG * 9 + F_WORKERS - 1) / F_WORKERS;
    uint32_t wv[WPT];
#pragma unroll
    for (int j = 0; j < WPT; ++j) {
      const int i = wt + j * F_WORKERS;
      wv[j] = i < nw ? __ldg(w32 + i) : 0u;
    }
    float4 k0 = make_float4(0.f, 0.f, 1.f, 0.f);
    if (wt < p.cout) {
      k0.y = e.bias ? __ldg(e.bias + wt) : 0.f;
      k0.z = e.bn_inv ? __ldg(e.bn_inv + wt) : 1.f;
      k0.w = e.bn_inv ? __ldg(e.bn_shift + wt) : 0.f;
    }
    // halo planes: everything the staging warps never write must be zero (pixel -1, pixels 32.., and rows outside the
    // image, which are rewritten per tile)
    for (int i = wt; i < F_NHALO * F_HALO / 16; i += F_WORKERS)
      *reinterpret_cast<uint4*>(sg + halo_off + i * 16) = make_uint4(0u, 0u, 0u, 0u);
    if (wt == 0) *flag_gen = 0;
#pragma unroll
    for (int j = 0; j < WPT; ++j) {
      const int i = wt + j * F_WORKERS;
      if (i < nw) wsm[i] = wv[j];
    }
    if (wt < p.groups * F_CG) *reinterpret_cast<float4*>(sg + cst_off + wt * 16) = k0;    // raw {-, bias, inv, shift}
    named_bar_sync(5, F_WORKERS);
    // A channel whose affine map is decreasing (BN slope < 0) gets its kernel levels negated, unless a level of -128
    // (8-bit kernels) makes that impossible somewhere: then the whole CTA falls back to a per-channel min / max choice.
    if (POOL && e.bn_inv != nullptr) {
      int bad = 0;
      for (int i = wt; i < nw; i += F_WORKERS) {
        const int ch = i / 9;
        if (reinterpret_cast<const float4*>(sg + cst_off)[ch].z < 0.f && __vcmpeq4(wsm[i], 0x80808080u) != 0u) bad = 1;
      }
      if (bad) atomicOr(flag_gen, 1);
      named_bar_sync(5, F_WORKERS);
    }
    neg_ok = POOL && (*reinterpret_cast<volatile int*>(flag_gen) == 0);
    ftrace(p, 2, 0);
    for (int i = wt; i < p.groups * 6 * 128 * 2; i += F_WORKERS) {
      const int chunk = i & 1;
      const int row = (i >> 1) & 127;
      const int blk = i >> 8;                       // (group * 2 + parity) * 3 + r
      const int r = blk % 3;
      const int par = (blk / 3) & 1;
      const int cg = blk / 6;
      const int ch = cg * F_CG + (row & 63);
      const int pos = 2 * (row >> 6) + par;         // output position inside the block of four pixels
      uint32_t wd[4] = {0u, 0u, 0u, 0u};
      if (ch < p.cout) {
        const bool neg = neg_ok && reinterpret_cast<const float4*>(sg + cst_off)[ch].z < 0.f;
#pragma unroll
        for (int b4 = 0; b4 < 4; ++b4) {
          const int s = chunk * 4 + b4 - pos;       // horizontal tap served by K slot (chunk*4 + b4)
          if (s >= 0 && s <= 2) {
            const uint32_t v = wsm[ch * 9 + r * 3 + s];
            wd[b4] = neg ? __vneg4(v) : v;
          }
        }
      }
      *reinterpret_cast<uint4*>(sg + blk * F_ABLK + (row >> 3) * 256 + chunk * 128 + (row & 7) * 16) =
          make_uint4(wd[0], wd[1], wd[2], wd[3]);
    }
    named_bar_sync(5, F_WORKERS);                   // every reader of the raw constants is done
    const float qm = SIGN ? 1.f : e.qm;
    if (wt < p.groups * F_CG) {
      float4 k = *reinterpret_cast<const float4*>(sg + cst_off + wt * 16);
      const bool neg = neg_ok && k.z < 0.f;
      k.x = neg ? -e.acc_scale : e.acc_scale;
      if (!OUT_F32) { k.z = __fmul_rn(k.z, qm); k.w = __fmul_rn(k.w, qm); }
      if (wt >= p.cout) k = make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(sg + cst_off + wt * 16) = k;
    }
    ftrace(p, 3, 0);
  }
  fence_proxy_async();                   // A blocks / zero fills were written through the generic proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  ftrace(p, 4, 0);
  if (warp != 0) griddep_wait();         // output buffer (and the staged rows' source) only after the previous kernel
  ftrace(p, 5, 0);

  if (warp == 0) {
    // ===================== bulk fetch + MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_i8(128, F_N, /*a signed*/ true, /*b unsigned*/ false);
      int it = 0;
      for (int k = 0; k < ntl; ++k) {
        const int hb = k % F_NHALO;
        mbar_wait(hfull(hb), (uint32_t)(k / F_NHALO) & 1u);
        ftrace(p, 10, k);
        tc_fence_after();
        fetch(k + F_RING);                // the staging warps have consumed ring slot k % F_RING
        const uint32_t halo = smem_base + halo_off + hb * F_HALO;
        for (int cg = 0; cg < p.groups; ++cg, ++it) {
          const int buf = it & 1;
          mbar_wait(tempty(buf), ((uint32_t)(it >> 1) & 1u) ^ 1u);
          ftrace(p, 11, it);
          tc_fence_after();
#pragma unroll
          for (int par = 0; par < 2; ++par) {
            const uint32_t d_tmem = tmem_base + (uint32_t)(buf * 256 + par * F_N);
#pragma unroll
            for (int r = 0; r < 3; ++r) {
              // A: 8-row groups 256 B apart, the two 16-byte K chunks 128 B apart
              const uint64_t a_desc = make_smem_desc_interleaved(smem_base + ((cg * 2 + par) * 3 + r) * F_ABLK, 128, 256);
              // B: image row (h + r) of the halo tile; 8 blocks = one 8-row group, next group = next image row,
              // second K chunk = the shifted plane
              const uint64_t b_desc = make_smem_desc_interleaved(halo + r * F_PITCH, F_PLANE, F_PITCH);
              if (!F_EXP(p, 2)) umma_i8(d_tmem, a_desc, b_desc, idesc, r > 0 ? 1u : 0u);
            }
          }
          umma_commit(tfull(buf));
          ftrace(p, 12, it);
        }
        umma_commit(hempty(hb));          // all MMAs that read this halo buffer have completed when this arrives
      }
    }
  } else if (warp > F_EPI_WARPS) {
    // ===================== halo staging (3 warps) =====================
    // item (halo row, group of four pixels): 12 RGB bytes from the raw ring -> 4 RGBX words, both planes
    const int st = threadIdx.x - (1 + F_EPI_WARPS) * 32;      // 0..95
    for (int k = 0; k < ntl; ++k) {
      const int hb = k % F_NHALO;
      if (k >= F_NHALO) mbar_wait(hempty(hb), (uint32_t)(k / F_NHALO - 1) & 1u);
      ftrace(p, 7, k);
      if (!F_EXP(p, 256)) mbar_wait(rfull(k % F_RING), (uint32_t)(k / F_RING) & 1u);
      ftrace(p, 8, k);
      const int tile = blockIdx.x + k * G;
      const int img = fdiv(tile, p.fd_h);
      const int top = (tile - img * p.tiles_h) * F_ROWS - 1;
#pragma unroll
      for (int rnd = 0; rnd < (F_ITEMS + F_STAGE_WARPS * 32 - 1) / (F_STAGE_WARPS * 32); ++rnd) {
        const int item = st + rnd * F_STAGE_WARPS * 32;
        if (item < F_ITEMS && !F_EXP(p, 128)) {
          const int row = item >> 3, col = item & 7;
          const int gh = top + row;
          uint32_t g0 = 0u, g1 = 0u, g2 = 0u;
          if (gh >= 0 && gh < p.h) {
            const uint32_t* src = reinterpret_cast<const uint32_t*>(sg + raw_off + (k % F_RING) * F_RAW + row * 96) + col * 3;
            g0 = src[0]; g1 = src[1]; g2 = src[2];
          }
          const uint32_t p0 = g0 & 0x00FFFFFFu;
          const uint32_t p1 = (g0 >> 24) | ((g1 & 0x0000FFFFu) << 8);
          const uint32_t p2 = (g1 >> 16) | ((g2 & 0x000000FFu) << 16);
          const uint32_t p3 = g2 >> 8;
          // pixels 4c..4c+3 live at bytes 16c+4 .. 16c+19 of the row: word, two words (8-byte aligned), word
          uint8_t* d0 = sg + halo_off + hb * F_HALO + row * F_PITCH + col * 16 + 4;
          *reinterpret_cast<uint32_t*>(d0) = p0;
          *reinterpret_cast<uint2*>(d0 + 4) = make_uint2(p1, p2);
          *reinterpret_cast<uint32_t*>(d0 + 12) = p3;
          if (col > 0) {                               // plane 1 starts at pixel 3: the first item only contributes p3
            uint8_t* d1 = d0 + F_PLANE - 16;
            *reinterpret_cast<uint32_t*>(d1) = p0;
            *reinterpret_cast<uint2*>(d1 + 4) = make_uint2(p1, p2);
            *reinterpret_cast<uint32_t*>(d1 + 12) = p3;
          } else {
            *reinterpret_cast<uint32_t*>(d0 + F_PLANE - 16 + 12) = p3;
          }
        }
      }
      fence_proxy_async();                         // generic-proxy writes -> visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(hfull(hb));
      ftrace(p, 6, k);
    }
  } else {
    // ===================== epilogue (2 groups x 8 warps, alternate sub-tiles) =====================
    // warp = (group, row half, lane quarter): TMEM lanes 32*quarter.. (hardware: warp id % 4), image rows 8*rh .. +7 of
    // the tile in two passes of four rows.  The accumulator buffer goes back to the MMA warp right after the last
    // TMEM load of the sub-tile.  The four warps of a (group, row half) share two alternating staging buffers and
    // leave through their own TMA store per pass (one 128-thread named barrier, no CTA-wide one).
    const int ew = warp - 1;
    const int grp = ew >> 3;
    const int rh = (ew >> 2) & 1;
    const int quarter = warp & 3;
    const int epair = quarter >> 1;               // position pair: output pixels 4j + 2*epair (+1)
    const int c_local = (quarter & 1) * 32 + lane;
    const bool leader = (ew & 3) == 0 && lane == 0;
    const int bar_id = 1 + (ew >> 2);             // 1..4
    const bool neg_mode = POOL && (*flag_gen == 0);
    constexpr int pitch = PITCH;
    uint8_t* stg_base = sg + stg_off + (ew >> 2) * 2 * SL::UNIT_BYTES;

    PairConst q;
    q.nz = pack2(p.neg_zero, p.neg_zero);
    q.one = pack2(p.one, p.one);
    q.s = q.a = q.b = q.c = 0ull;
    bool dec = false;
    const int nsub = ntl * p.groups;
    int unit_no = 0;                              // alternates this (group, row half)'s two staging buffers
    for (int it = grp; it < nsub; it += 2) {
      const int k = fdiv(it, p.fd_g);
      const int cg = it - k * p.groups;
      const int tile = blockIdx.x + k * G;
      const int img = fdiv(tile, p.fd_h);
      const int h0 = (tile - img * p.tiles_h) * F_ROWS;
      const int ch = cg * F_CG + c_local;
      const bool ch_ok = ch < p.cout;
      const bool warp_active = (cg * F_CG + (quarter & 1) * 32) < p.cout;      // warp-uniform
      if (p.groups > 1 || it == grp) {
        const float4 kc = *reinterpret_cast<const float4*>(sg + cst_off + ch * 16);
        q.s = pack2(kc.x, kc.x); q.a = pack2(kc.y, kc.y); q.b = pack2(kc.z, kc.z); q.c = pack2(kc.w, kc.w);
        dec = POOL && !neg_mode && kc.z < 0.f;
      }
      ftrace(p, 20, it);
      if (F_EXP(p, 64)) mbar_wait(tfull(grp), (uint32_t)(it >> 1) & 1u);
      else mbar_wait_parked(tfull(grp), (uint32_t)(it >> 1) & 1u);
      tc_fence_after();
      ftrace(p, 21, it);
      if constexpr (POOL && !OUT_F32) {
        // Pooled path: the sub-tile's four row pairs (chunks) are software-pipelined through ONE set of 32 registers:
        // as soon as the max tree has reduced a chunk to 8 values, the TMEM loads of the next chunk are issued, so
        // their latency sits under the affine arithmetic of the current one.
        uint8_t* stg = stg_base + (unit_no & 1) * SL::UNIT_BYTES;
        int a0[16], b0[16];
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(grp * 256 + rh * 64);
        const bool loads = warp_active && !F_EXP(p, 8);
        const bool math = warp_active && !F_EXP(p, 1);
        if (loads) {
          __syncwarp();
          tmem_ld16_nowait(taddr, a0);
          tmem_ld16_nowait(taddr + F_N, b0);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {              // pooled row c of the row half: image rows 8*rh + 2*c, +1
          if (loads) tmem_ld_wait_dep16x2(a0, b0);
          // pooled pixel (4*rh + c, 2j + epair) = max over {a, b} x {upper, lower image row} of block j
          int m0[4], m1[4], l0[4], l1[4];
          if (dec) {                               // rare: a decreasing channel whose kernel could not be negated
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              m0[i] = min(min(a0[2 * i], b0[2 * i]), min(a0[8 + 2 * i], b0[8 + 2 * i]));
              m1[i] = min(min(a0[2 * i + 1], b0[2 * i + 1]), min(a0[9 + 2 * i], b0[9 + 2 * i]));
            }
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              m0[i] = max(max(a0[2 * i], b0[2 * i]), max(a0[8 + 2 * i], b0[8 + 2 * i]));
              m1[i] = max(max(a0[2 * i + 1], b0[2 * i + 1]), max(a0[9 + 2 * i], b0[9 + 2 * i]));
            }
          }
          if (c < 3) {
            if (loads) {
              __syncwarp();
              tmem_ld16_nowait(taddr + 16 * (c + 1), a0);
              tmem_ld16_nowait(taddr + F_N + 16 * (c + 1), b0);
            }
          } else {
            // all TMEM reads of this warp for this sub-tile are done: hand the accumulator back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty(grp));
          }
          if (math) {
            affine_levels<SIGN, 4>(m0, m1, l0, l1, q, p.qlo, p.qhi);
            uint8_t* srow = stg + ((c * 16 + epair) * pitch) + c_local;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              srow[(4 * i) * pitch] = (uint8_t)l0[i];
              srow[(4 * i + 2) * pitch] = (uint8_t)l1[i];
            }
          }
        }
        ftrace(p, 24, it);
        if (!F_EXP(p, 16)) fence_proxy_async();    // staging writes -> visible to the TMA (async proxy)
        // the store this row half issued one sub-tile ago (other buffer) has had this whole sub-tile to drain; once the
        // leader has confirmed that, the barrier also tells the four warps that the OTHER buffer may be overwritten
        if (leader) tma_store_wait_read();
        if (!F_EXP(p, 32)) named_bar_sync(bar_id, 128);
        if (leader && !F_EXP(p, 4)) {
          tma_store_4d(&map_y, smem_u32(stg), cg * F_CG, 0, (h0 >> 1) + 4 * rh, img);   // the row half's 4 pooled rows
          tma_store_commit();
        }
        ++unit_no;
        ftrace(p, 26, it);
      } else {
#pragma unroll 1
      for (int ps = 0; ps < 2; ++ps) {
        const int rowq = 2 * rh + ps;             // image rows 4*rowq .. 4*rowq + 3 of the tile
        uint8_t* stg = stg_base + (unit_no & 1) * SL::UNIT_BYTES;
        // two row pairs per pass, each through the same 32 registers (more live accumulators leave the compiler no
        // room to interleave independent pairs in the arithmetic)
        int a0[16], b0[16];
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(grp * 256 + rowq * 32);
        const bool loads = warp_active && !F_EXP(p, 8);
        const bool math = warp_active && !F_EXP(p, 1);
#pragma unroll
        for (int pr = 0; pr < 2; ++pr) {
          if (loads) {
            __syncwarp();
            tmem_ld16_nowait(taddr + 16 * pr, a0);
            tmem_ld16_nowait(taddr + F_N + 16 * pr, b0);
            tmem_ld_wait_dep16x2(a0, b0);
          }
          if (pr == 1 && ps == 1) {
            // all TMEM reads of this warp for this sub-tile are done: hand the accumulator back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty(grp));
          }
          if (!math) continue;
          // pr = row pair inside the pass (image rows 4*rowq + 2*pr, +1)
          if constexpr (OUT_F32) {
            // plain fp32 output (no activation): y = affine(acc); pixels (h, 4j + 2*epair) and the next one
            const float4 kc = *reinterpret_cast<const float4*>(sg + cst_off + ch * 16);
            ChanConst cc;
            cc.scale = kc.x; cc.bias = kc.y; cc.inv = kc.z; cc.shift = kc.w;
            cc.has_bias = (e.bias != nullptr); cc.has_bn = (e.bn_inv != nullptr);
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
              const int hh = h0 + 4 * rowq + 2 * pr + rr;
              if (hh < p.h && ch_ok) {
                float* dst = (float*)p.y + (((long long)img * p.h + hh) * 32) * p.cout + ch;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  dst[(long long)(4 * j + 2 * epair) * p.cout] = affine((float)a0[rr * 8 + j], cc);
                  dst[(long long)(4 * j + 2 * epair + 1) * p.cout] = affine((float)b0[rr * 8 + j], cc);
                }
              }
            }
          } else if constexpr (POOL) {
            // pooled pixel (2*rowq + pr, 2j + epair) = max over {a, b} x {upper, lower image row} of block j
            int m0[4], m1[4], l0[4], l1[4];
            if (dec) {                             // rare: a decreasing channel whose kernel could not be negated
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                m0[i] = min(min(a0[2 * i], b0[2 * i]), min(a0[8 + 2 * i], b0[8 + 2 * i]));
                m1[i] = min(min(a0[2 * i + 1], b0[2 * i + 1]), min(a0[9 + 2 * i], b0[9 + 2 * i]));
              }
            } else {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                m0[i] = max(max(a0[2 * i], b0[2 * i]), max(a0[8 + 2 * i], b0[8 + 2 * i]));
                m1[i] = max(max(a0[2 * i + 1], b0[2 * i + 1]), max(a0[9 + 2 * i], b0[9 + 2 * i]));
              }
            }
            affine_levels<SIGN, 4>(m0, m1, l0, l1, q, p.qlo, p.qhi);
            uint8_t* srow = stg + (((2 * ps + pr) * 16 + epair) * pitch) + c_local;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              srow[(4 * i) * pitch] = (uint8_t)l0[i];
              srow[(4 * i + 2) * pitch] = (uint8_t)l1[i];
            }
          } else {
            // pixels (row, 4j + 2*epair) = a, (row, 4j + 2*epair + 1) = b
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
              int m0[8], m1[8], l0[8], l1[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) { m0[j] = a0[rr * 8 + j]; m1[j] = b0[rr * 8 + j]; }
              affine_levels<SIGN, 8>(m0, m1, l0, l1, q, p.qlo, p.qhi);
              uint8_t* srow = stg + (((2 * pr + rr) * 32 + 2 * epair) * pitch) + c_local;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                srow[(4 * j) * pitch] = (uint8_t)l0[j];
                srow[(4 * j + 1) * pitch] = (uint8_t)l1[j];
              }
            }
          }
        }
        ftrace(p, 24, it);
        if constexpr (!OUT_F32) {
          if (!POOL || ps == 1) {
            if (!F_EXP(p, 16)) fence_proxy_async();    // staging writes -> visible to the TMA (async proxy)
            // the store this row half issued one unit ago (other buffer) has had this whole unit to drain; once the
            // leader has confirmed that, the barrier also tells the four warps that the OTHER buffer may be overwritten
            if (leader) tma_store_wait_read();
            if (!F_EXP(p, 32)) named_bar_sync(bar_id, 128);
            if (leader && !F_EXP(p, 4)) {
              // POOL: the row half's 4 pooled rows of this sub-tile; otherwise one pass of 4 image rows
              if constexpr (POOL) tma_store_4d(&map_y, smem_u32(stg), cg * F_CG, 0, (h0 >> 1) + 4 * rh, img);
              else tma_store_4d(&map_y, smem_u32(stg), cg * F_CG, 0, h0 + 4 * rowq, img);
              tma_store_commit();
            }
            ++unit_no;
          }
        }
        ftrace(p, 26, it);
      }
      }
    }
    if constexpr (!OUT_F32) {
      if (leader) tma_store_wait_all();
    }
  }

  ftrace(p, 30, 0);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
  ftrace(p, 31, 0);
#ifdef QNNB_TRACE
  if (p.tr != nullptr && blockIdx.x == 0 && lane == 0) p.tr[warp * 1024] = (unsigned long long)trk__;
#endif
}

}  // namespace

bool conv_first_tc_shape(const qnnb_conv_desc& d) {
  return d.in_kind == QNNB_KIND_U8 && d.cin == 3 && d.kh == 3 && d.kw == 3 && d.stride == 1 && d.w == 32 && d.h % 8 == 0 &&
         d.cout <= F_MAX_GROUPS * F_CG && d.cout % 32 == 0;
}

int launch_conv_first_tc(const qnnb_conv_desc& d, const void* x, const void* w, void* y, cudaStream_t st) {
  EncodeTiledFn encode = get_encode();
  if (!encode) { set_error("conv2d: cuTensorMapEncodeTiled is not available from the driver"); return QNNB_ECUDA; }
  const bool pool = d.epi.pool == 2;
  const bool f32 = d.epi.act == QNNB_ACT_NONE;
  const bool sign = d.epi.act == QNNB_ACT_SIGN_I8;
  if (((uintptr_t)x & 15u) != 0) {
    // the raw image rows arrive by 16-byte bulk copies
    if (d.impl == QNNB_IMPL_AUTO) return launch_conv_generic(d, x, w, y, st);
    set_error("conv2d: the first-layer tensor-core kernel needs a 16-byte aligned input pointer");
    return QNNB_EUNSUPPORTED;
  }
  FirstParams p;
  memset(&p, 0, sizeof(p));
  p.x = (const uint8_t*)x; p.wpk = (const int8_t*)w; p.y = y;
  p.n = d.n; p.h = d.h; p.cout = d.cout;
  p.groups = ceil_div(d.cout, F_CG);
  p.tiles_h = ceil_div(d.h, F_ROWS);
  p.num_tiles = p.tiles_h * d.n;
  p.fd_h = make_fastdiv(p.tiles_h);
  p.fd_g = make_fastdiv(p.groups);
  p.out_pitch = d.cout < F_CG ? d.cout : F_CG;
  p.neg_zero = -0.0f; p.one = 1.0f;
  p.tr = get_trace_buffer();
  { const char* ev = getenv("QNNB_K5_EXP"); p.exp = ev ? atoi(ev) : 0; }
  p.epi = make_epi(d.epi);
  p.qlo = -p.epi.qm; p.qhi = p.epi.qm - 1.f;
  const int grid = p.num_tiles < grid_sms(d.max_ctas) ? p.num_tiles : grid_sms(d.max_ctas);
  CUtensorMap my;
  memset(&my, 0, sizeof(my));
  if (!f32) {
    const int oh = pool ? d.h / 2 : d.h, ow = pool ? 16 : 32;
    cuuint64_t dims[4] = {(cuuint64_t)d.cout, (cuuint64_t)ow, (cuuint64_t)oh, (cuuint64_t)d.n};
    cuuint64_t strides[3] = {(cuuint64_t)d.cout, (cuuint64_t)ow * d.cout, (cuuint64_t)oh * ow * d.cout};
    cuuint32_t box[4] = {(cuuint32_t)p.out_pitch, (cuuint32_t)ow, 4u, 1u};       // one store unit: 4 (pooled) rows
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(&my, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, y, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("conv2d: cuTensorMapEncodeTiled(first-layer output) failed with %d", (int)r); return QNNB_ECUDA; }
  }
  auto go = [&](auto kern, auto sl, SmemConfigured& once) -> int {
    using SLT = decltype(sl);
    const int smem = (((p.groups * 6 * F_ABLK + F_NHALO * F_HALO + p.groups * F_CG * 16 + F_RING * F_RAW + 1023) & ~1023) + SLT::STG_BYTES) + 256 + 1024;
    if (smem > 232448) { set_error("conv2d: first-layer kernel needs %d bytes of shared memory", smem); return QNNB_EINVAL; }
    QNNB_CUDA(once.ensure(kern, smem));
    QNNB_CUDA(launch_pdl(kern, dim3(grid), dim3(F_THREADS), (size_t)smem, st, my, p));
    return QNNB_OK;
  };
  static SmemConfigured once[9];
  if (f32) return go(conv3x3_c3_direct_tc_kernel<false, true, false, 64>, FirstSmem<false, true>{}, once[0]);
  const bool narrow = p.out_pitch == 32;
  if (pool) {
    if (sign) return narrow ? go(conv3x3_c3_direct_tc_kernel<true, false, true, 32>, FirstSmem<true, false>{}, once[1])
                            : go(conv3x3_c3_direct_tc_kernel<true, false, true, 64>, FirstSmem<true, false>{}, once[2]);
    return narrow ? go(conv3x3_c3_direct_tc_kernel<true, false, false, 32>, FirstSmem<true, false>{}, once[3])
                  : go(conv3x3_c3_direct_tc_kernel<true, false, false, 64>, FirstSmem<true, false>{}, once[4]);
  }
  if (sign) return narrow ? go(conv3x3_c3_direct_tc_kernel<false, false, true, 32>, FirstSmem<false, false>{}, once[5])
                          : go(conv3x3_c3_direct_tc_kernel<false, false, true, 64>, FirstSmem<false, false>{}, once[6]);
  return narrow ? go(conv3x3_c3_direct_tc_kernel<false, false, false, 32>, FirstSmem<false, false>{}, once[7])
                : go(conv3x3_c3_direct_tc_kernel<false, false, false, 64>, FirstSmem<false, false>{}, once[8]);
}

}  // namespace qnnb
