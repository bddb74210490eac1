// K1 -- 3x3 stride-1 SAME int8 convolution as an implicit GEMM on the 5th-gen tensor cores:
// TMA-staged NHWC tiles -> tcgen05.mma kind::i8 -> int32 accumulators in TMEM -> fused epilogue.
//
// Stands in for K.conv2d + bias_add + BatchNormalization + quantized_tanh (+ MaxPooling2D) of
// QuantizedConv2D.call (layers/quantized_layers.py:164-194) as chained by models/vgg.py:15-37.
//
// Orientation: D[channel][pixel] = W[channel][K] * X[pixel][K]^T, i.e. the WEIGHTS are the MMA's A
// operand (M = 128 output channels) and the ACTIVATIONS the B operand (N = 256 output pixels), both
// K-major.  TMEM lane = output channel, TMEM column = pixel, so an epilogue thread owns ONE channel:
// its bias / BN constants live in registers and the 2x2 max-pool is an in-register reduction over
// columns (done on the raw accumulators: every epilogue step is monotone, see conv_generic.cu).
//
// Implicit GEMM: K = 9 taps x Cin.  For tap (r,s) and channel chunk c0 the B tile is ONE 4-D TMA box
// {KC, TW, TH, TN} of the NHWC tensor at (c0, w0+s-1, h0+r-1, n0): out-of-bounds rows/columns are
// zero-filled by the TMA unit, which IS the SAME padding.  The box lands as 256 rows of KC bytes with
// the 64B/128B swizzle the UMMA shared-memory descriptor expects.  The A tile is a 2-D box
// {KC, 128} of the packed kernel [Cout][9*Cin].
//
// Output path: the epilogue threads write their int8 levels into a [pixel][128 channel] staging tile in
// shared memory (byte stores, 32 consecutive bytes per warp instruction), then one thread issues a single
// 4-D TMA store of the whole tile ({128, TW', TH', TN} box of the NHWC output; out-of-range images are
// clipped by the TMA unit), so HBM sees full 128-byte rows.
//
// Warp roles (384 threads; every single-threaded role runs under elect.sync): warp 0 = TMA producer (v2: halo tiles),
// warp 1 = MMA issuer, warp 2 = TMEM allocator (v2: weight-stage watcher), warp 3 = (v2) weight-tile TMA producer,
// warps 4..11 = epilogue: lock-step (two warps per TMEM lane quarter, each half of the columns) or, where a second
// staging tile fits, two independent groups of four warps draining alternate tiles.
// Pipelines: smem rings full/empty (TMA <-> MMA), 2 TMEM accumulators full/empty (MMA <-> epilogue), persistent static
// tile schedule (tile index decoded by multiply-shift).  Layers chain through programmatic dependent launch: the next
// kernel's prologue and weight loads run while this one drains (griddepcontrol, common.cuh).
#include "common.cuh"
#include "tc_ptx.cuh"

#include <cuda.h>

namespace qnnb {

namespace {

constexpr int TILE_M = 128;   // output channels per tile (UMMA M)
constexpr int TILE_N = 256;   // output pixels per tile   (UMMA N)
constexpr int UMMA_K = 32;    // int8
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = 128 + NUM_EPI_WARPS * 32;

using namespace tcx;

// Optional device-side timeline (debug/profiling only): when a buffer is registered with qnnb_debug_set_trace(),
// CTA 0 records globaltimer stamps of its pipeline events as (tag, value) pairs.  NULL in production.
struct Trace { unsigned long long* buf; int cap; };
static Trace g_trace = {nullptr, 0};

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// compiled in only with -DQNNB_TRACE (make TRACE=1): even a never-taken branch in the single-thread MMA issue loop
// costs measurable time (v1 16x16 layer: 22 -> 32 us)
__device__ __forceinline__ void trace(const Trace& tr, int tag, int idx) {
#ifdef QNNB_TRACE
  // lock-free: every recording thread owns a 512-event region (chosen by its warp) and a private counter kept in
  // the first word of that region; two fire-and-forget stores per event
  if (tr.buf != nullptr && blockIdx.x == 0) {
    unsigned long long* reg = tr.buf + (threadIdx.x >> 5) * 1024;
    const unsigned long long k = reg[0];
    if (k < 510) {
      reg[2 + 2 * k] = ((unsigned long long)tag << 32) | (unsigned)idx;
      reg[3 + 2 * k] = gtime();
      reg[0] = k + 1;
    }
  }
#endif
}

// tile -> (channel tile, column tile, row tile, image group)
struct TileCoord { int mt, tw_i, th_i, pt; };

struct TcParams {
  Trace tr;
  int n, h, w, cin, cout;
  int tiles_w, tiles_h, tiles_n, m_tiles, num_tiles;
  FastDiv fd_m, fd_w, fd_h;
  int kchunks;            // cin / KC
  int resident;           // v2: the CTA's nine weight tiles are loaded once and reused by every tile (Cin = 64)
  int out_pitch;          // bytes per staged output row = channels per TMA-store box (<= 128)
  void* y;
  Epi epi;
};

// two decoupled epilogue warp groups (see epilogue_role_n) where a second staging tile fits; -DQNNB_NO_SPLIT for A/B runs
#ifdef QNNB_NO_SPLIT
constexpr bool SPLIT_EPILOGUE = false;
#else
constexpr bool SPLIT_EPILOGUE = true;
#endif
__device__ __forceinline__ TileCoord decode_tile(const TcParams& p, int tile) {
  TileCoord t;
  int q = fdiv(tile, p.fd_m);
  t.mt = tile - q * p.m_tiles;
  int q2 = fdiv(q, p.fd_w);
  t.tw_i = q - q2 * p.tiles_w;
  t.pt = fdiv(q2, p.fd_h);
  t.th_i = q2 - t.pt * p.tiles_h;
  return t;
}

constexpr int EPI_BAR_ID = 2;                      // named barriers 2 and 3
constexpr int EPI_THREADS = NUM_EPI_WARPS * 32;

template <bool POOL, bool OUT_F32>
struct StageTile {
  static constexpr int ROWS = OUT_F32 ? 0 : (POOL ? TILE_N / 4 : TILE_N);
  static constexpr int BYTES = ROWS * TILE_M;
};

// ------------------------------------------------------------------ epilogue role (shared by K1 and K5)
// Warp `warp` (4..11) drains TMEM lanes 32*(warp%4).. of both accumulators: thread = one output channel,
// columns = the 256 pixels of the tile in {TN, TH, TW} order.  Fixed fp32 op order of common.cuh.
// FOLD: acc_scale is a power of two (see QConst); PITCH: bytes per staged row (0 = runtime p.out_pitch).
// GROUPS (K5 only, TW = 32): when Cout <= 64 the 128 TMEM lanes hold GROUPS = 2 pixel groups of 64 channels --
// lane L is channel L % 64 of the pixels 8*(L / 64) rows further down, so one tile covers 16 image rows.
// NSTG staging tiles (stg, stg + stg_bytes): with 2, the TMA store of tile t drains while tile t+1 is computed.
// PREFETCH: issue both 64-column TMEM loads up front (128 accumulator registers; fine at 384 threads per CTA, too
// many for the 512-thread first-layer kernel).
// ILV (v2 kernel, TW = 8): the accumulator columns are ordered [row][image][8 px] (image-interleaved rows) instead of
// [image][row][8 px]; the output tensor map has its N and H dimensions swapped to match.
// SPLIT: the eight warps form two independent groups of four (one warp per TMEM lane quarter); group g owns
// accumulator g, i.e. every second tile of the CTA, drains all 256 columns of it and has its own staging tile, named
// barrier and TMA-store leader.  The two groups run out of phase, so the TMEM-load latency, barrier and store of one
// tile overlap the arithmetic of the next (lock-step, every tile pays them in sequence).  tempty barriers then count 4.
// SIGN: the activation is binary_tanh instead of quantized_tanh (level +1 / -1; qm is 1, so qaffine yields z itself).
template <int TW, bool POOL, bool OUT_F32, bool FOLD, int PITCH, int GROUPS = 1, int TH_ = 0, int NSTG = 1, bool PREFETCH = true,
          bool ILV = false, bool SIGN = false, bool SPLIT = false>
__device__ __forceinline__ void epilogue_role_n(const TcParams& p, const CUtensorMap* map_y, uint32_t tmem_base, uint32_t tfull0,
                                                uint32_t tempty0, uint8_t* stg0, int stg_bytes, int warp, int lane);

template <int TW, bool POOL, bool OUT_F32, bool FOLD, int PITCH, int GROUPS = 1, int TH_ = 0>
__device__ __forceinline__ void epilogue_role(const TcParams& p, const CUtensorMap* map_y, uint32_t tmem_base, uint32_t tfull0,
                                              uint32_t tempty0, uint8_t* stg, int warp, int lane) {
  epilogue_role_n<TW, POOL, OUT_F32, FOLD, PITCH, GROUPS, TH_, 1>(p, map_y, tmem_base, tfull0, tempty0, stg, 0, warp, lane);
}

template <bool SIGN>
__device__ __forceinline__ int out_level(float zq, float qm) {
  if constexpr (SIGN) return act_sign(zq) ? 1 : -1;
  else return quant_scaled(zq, qm);
}

template <int TW, bool POOL, bool OUT_F32, bool FOLD, int PITCH, int GROUPS, int TH_, int NSTG, bool PREFETCH, bool ILV, bool SIGN, bool SPLIT>
__device__ __forceinline__ void epilogue_role_n(const TcParams& p, const CUtensorMap* map_y, uint32_t tmem_base, uint32_t tfull0,
                                                uint32_t tempty0, uint8_t* stg0, int stg_bytes, int warp, int lane) {
  constexpr int TH = TH_ ? TH_ : ((TW == 32) ? 8 : (TW == 16 ? 16 : 8));
  constexpr int TN = TILE_N / (TW * TH);
  static_assert(GROUPS == 1 || TW == 32, "pixel groups only exist for the first-layer geometry");
  const int quarter = warp & 3;                 // TMEM lanes 32*quarter .. +31 (hardware restriction: warp_id % 4)
  const int egrp = (warp - 4) >> 2;             // epilogue warp group
  const int half = SPLIT ? 0 : egrp;            // lock-step: which 128 columns of the accumulator
  const int group = (GROUPS == 2) ? (quarter >> 1) : 0;
  const int grow = group * TH;                  // row offset of this lane's pixel group inside the tile
  const int ch_in_tile = (GROUPS == 2) ? ((quarter & 1) * 32 + lane) : (quarter * 32 + lane);
  const bool leader = SPLIT ? (quarter == 0 && lane == 0) : (warp == 4 && lane == 0);
  const int bar_id = SPLIT ? EPI_BAR_ID + egrp : EPI_BAR_ID;
  constexpr int BAR_THREADS = SPLIT ? EPI_THREADS / 2 : EPI_THREADS;
  constexpr int IT_STEP = SPLIT ? 2 : 1;
  static_assert(!SPLIT || OUT_F32 || NSTG == 2, "split epilogue groups need one staging tile each");
  const Epi& e = p.epi;
  int cur_mt = -1;
  ChanConst cc = {};
  QConst qc = {1.f, 0.f, 1.f, 0.f};
  int it = SPLIT ? egrp : 0;
  for (int tile = blockIdx.x + it * (int)gridDim.x; tile < p.num_tiles; tile += IT_STEP * (int)gridDim.x, it += IT_STEP) {
    const TileCoord tc = decode_tile(p, tile);
    const int mt = tc.mt;
    const int n0 = tc.pt * TN, h0 = tc.th_i * (TH * GROUPS), w0 = tc.tw_i * TW;
    const int ch = mt * TILE_M + ch_in_tile;
    const bool ch_ok = ch < p.cout;
    const bool warp_active = (mt * TILE_M + ch_in_tile - lane) < p.cout; // warp-uniform
    // per-channel constants: global loads, so only when the channel tile changes (with an even grid stride a
    // persistent CTA keeps the same channel tile for its whole life)
    if (mt != cur_mt) {
      cur_mt = mt;
      if constexpr (OUT_F32) cc = load_chan(e, ch, ch_ok);
      else qc = make_qconst<FOLD>(e, ch, ch_ok);
    }
    // Pooling picks max(acc) where the channel's map is increasing and min(acc) where it is decreasing (BN slope
    // < 0).  The integer min/max instruction takes the choice as a predicate operand, so selecting per 2-/3-input
    // step (pick2 / pick3 below) costs ONE reduction tree -- not both extrema plus a select.
    const bool dec = qc.b < 0.f;
    auto pick4 = [dec](int a, int b, int c, int d) {
      const int lo = dec ? min(a, b) : max(a, b);
      return dec ? min(min(c, d), lo) : max(max(c, d), lo);
    };
    const int pitch = PITCH ? PITCH : p.out_pitch;
    const float qm = e.qm;
    const int acc = it & 1;
    const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
    uint8_t* stg = stg0 + ((NSTG == 2) ? (it & 1) * stg_bytes : 0);
    if (leader) trace(p.tr, 11, tile);                    // epilogue: loop top
    if constexpr (!OUT_F32 && (NSTG == 1 || SPLIT)) {
      // single staging tile (per group): the previous TMA store must have finished READING it before it is overwritten
      if (leader) tma_store_wait_read();
      named_bar_sync(bar_id, BAR_THREADS);
    }
    mbar_wait_parked(tfull0 + 8u * acc, acc_phase);
    if (leader) trace(p.tr, 7, tile);                     // epilogue: accumulator complete
    tc_fence_after();
    // one 64-column chunk (rows [row0, row0 + 64/TW) of image n0 + img) -> output
    auto process = [&](int (&v)[64], int col0) {
      if constexpr (ILV) {
        // 64 columns = 8 groups of 8 pixels; group g = row (g / TN) of image (g % TN)
        static_assert(!ILV || TW == 8, "interleaved order is defined for 8-pixel-wide tiles");
        constexpr int RPC = 8 / TN;                  // image rows covered by one chunk
        const int row0 = (col0 >> 3) / TN;
        if constexpr (OUT_F32) {
#pragma unroll
          for (int gi = 0; gi < 8; ++gi) {
            const int hh = row0 + gi / TN, img = gi % TN;
            if (n0 + img < p.n) {
              const long long pix0 = ((long long)(n0 + img) * p.h + (h0 + hh)) * p.w + w0;
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                const float z = affine((float)v[gi * 8 + c], cc);
                if (ch_ok) ((float*)p.y)[(pix0 + c) * p.cout + ch] = z;
              }
            }
          }
        } else if constexpr (POOL) {
          constexpr int PRC = RPC / 2;
          uint8_t* srow = stg + ((row0 >> 1) * TN * 4) * pitch + ch_in_tile;
#pragma unroll
          for (int pr = 0; pr < PRC; ++pr) {
#pragma unroll
            for (int img = 0; img < TN; ++img) {
#pragma unroll
              for (int pc = 0; pc < 4; ++pc) {
                const int i00 = ((2 * pr) * TN + img) * 8 + 2 * pc;
                constexpr int VS = TN * 8;             // columns between vertically adjacent pixels
                const int mx = pick4(v[i00], v[i00 + 1], v[i00 + VS], v[i00 + VS + 1]);
                srow[((pr * TN + img) * 4 + pc) * pitch] = (uint8_t)out_level<SIGN>(qaffine<FOLD>(mx, qc), qm);
              }
            }
          }
        } else {
          uint8_t* srow = stg + col0 * pitch + ch_in_tile;
#pragma unroll
          for (int c = 0; c < 64; ++c) srow[c * pitch] = (uint8_t)out_level<SIGN>(qaffine<FOLD>(v[c], qc), qm);
        }
        return;
      }
      const int img = col0 / (TH * TW);
      const int row0 = (col0 % (TH * TW)) / TW;
      if constexpr (OUT_F32) {
        const int nimg = n0 + img;
        if (nimg >= p.n) return;                   // warp-uniform (ragged last image group)
        constexpr int R = 64 / TW;
        const long long pix0 = ((long long)nimg * p.h + (h0 + grow + row0)) * p.w + w0;
#pragma unroll
        for (int rr = 0; rr < R; ++rr) {
#pragma unroll
          for (int c = 0; c < TW; ++c) {
            const float z = affine((float)v[rr * TW + c], cc);
            if (ch_ok) ((float*)p.y)[(pix0 + (long long)rr * p.w + c) * p.cout + ch] = z;
          }
        }
      } else if constexpr (POOL) {
        constexpr int PR = 64 / TW / 2, PC = TW / 2;
        // every lane of an active warp owns a real channel (Cout % 32 == 0 on this path)
        uint8_t* srow = stg + (img * (TH / 2) * PC + ((grow + row0) >> 1) * PC) * pitch + ch_in_tile;
#pragma unroll
        for (int pr = 0; pr < PR; ++pr) {
#pragma unroll
          for (int pc = 0; pc < PC; ++pc) {
            const int i00 = (2 * pr) * TW + 2 * pc;
            const int mx = pick4(v[i00], v[i00 + 1], v[i00 + TW], v[i00 + TW + 1]);
            srow[(pr * PC + pc) * pitch] = (uint8_t)out_level<SIGN>(qaffine<FOLD>(mx, qc), qm);
          }
        }
      } else {
        uint8_t* srow = stg + (grow * TW + col0) * pitch + ch_in_tile;
#pragma unroll
        for (int c = 0; c < 64; ++c) srow[c * pitch] = (uint8_t)out_level<SIGN>(qaffine<FOLD>(v[c], qc), qm);
      }
    };
    if (warp_active) {
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * TILE_N + half * 128);
      if constexpr (SPLIT && PREFETCH) {
        // four 64-column chunks, software pipelined: the load of chunk j+1 is issued after the wait for chunk j
        // (tcgen05.wait::ld covers every load issued so far) and is in flight while chunk j is processed
        int va[64], vb[64];
        __syncwarp();
        tmem_ld64(taddr, va);
        tmem_ld_wait_dep(va);
        __syncwarp();
        tmem_ld64(taddr + 64, vb);
        process(va, 0);
        tmem_ld_wait_dep(vb);
        __syncwarp();
        tmem_ld64(taddr + 128, va);
        process(vb, 64);
        tmem_ld_wait_dep(va);
        __syncwarp();
        tmem_ld64(taddr + 192, vb);
        process(va, 128);
        tmem_ld_wait_dep(vb);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty0 + 8u * acc);
        process(vb, 192);
      } else if constexpr (SPLIT && POOL && TW == 32 && !OUT_F32 && !ILV) {
        // first-layer geometry (8 rows x 32 px per pixel group), 512-thread kernel: 128 registers per thread rule out
        // two 64-column buffers, so the accumulator is drained in eight 2 x 16-column pieces -- 16 pixels of image
        // rows 2pr and 2pr+1, i.e. eight pooled outputs -- double-buffered: the loads of piece k+1 are in flight
        // while piece k is pooled, scaled and re-quantised
        constexpr int PC = TW / 2;
        int va[32], vb[32];
        auto ld_piece = [&](int k, int (&v)[32]) {
          const uint32_t a = taddr + (uint32_t)((k >> 1) * 2 * TW + (k & 1) * 16);
          __syncwarp();
          tmem_ld16_nowait(a, &v[0]);
          tmem_ld16_nowait(a + TW, &v[16]);
        };
        auto do_piece = [&](int k, int (&v)[32]) {
          uint8_t* srow = stg + (((grow >> 1) + (k >> 1)) * PC + (k & 1) * 8) * pitch + ch_in_tile;
#pragma unroll
          for (int pc = 0; pc < 8; ++pc) {
            const int mx = pick4(v[2 * pc], v[2 * pc + 1], v[16 + 2 * pc], v[16 + 2 * pc + 1]);
            srow[pc * pitch] = (uint8_t)out_level<SIGN>(qaffine<FOLD>(mx, qc), qm);
          }
        };
        ld_piece(0, va);
        tmem_ld_wait_dep32(va);
#pragma unroll
        for (int k = 0; k < 8; k += 2) {
          ld_piece(k + 1, vb);
          do_piece(k, va);
          tmem_ld_wait_dep32(vb);
          if (k + 2 < 8) {
            ld_piece(k + 2, va);
          } else {
            tc_fence_before();                       // all TMEM reads of this tile done: release the accumulator
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty0 + 8u * acc);
          }
          do_piece(k + 1, vb);
          if (k + 2 < 8) tmem_ld_wait_dep32(va);
        }
      } else if constexpr (SPLIT) {
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
          int v[64];
          __syncwarp();
          tmem_ld64(taddr + 64 * j, v);
          tmem_ld_wait_dep(v);
          if (leader) trace(p.tr, 20 + j, tile);
          if (j == 3) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty0 + 8u * acc);
          }
          process(v, 64 * j);
        }
      } else if constexpr (PREFETCH) {
        int va[64], vb[64];
        __syncwarp();                              // tcgen05.ld is warp-collective (.sync.aligned)
        tmem_ld64(taddr, va);
        tmem_ld64(taddr + 64, vb);                 // second chunk in flight while the first is processed
        tmem_ld_wait_dep(va);
        if (leader) trace(p.tr, 12, tile);
        process(va, half * 128);
        __syncwarp();
        tmem_ld_wait_dep(vb);
        // all TMEM reads of this warp for this tile are done: hand the accumulator back to the MMA warp
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty0 + 8u * acc);
        if (leader) trace(p.tr, 13, tile);
        process(vb, half * 128 + 64);
      } else {
#pragma unroll 1
        for (int j = 0; j < 2; ++j) {
          int v[64];
          __syncwarp();
          tmem_ld64(taddr + 64 * j, v);
          tmem_ld_wait_dep(v);
          if (leader) trace(p.tr, 12 + j, tile);
          if (j == 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty0 + 8u * acc);
          }
          process(v, half * 128 + 64 * j);
        }
      }
    } else {
      // this lane quarter holds only padding channels: nothing to read
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8u * acc);
    }
    if (leader) trace(p.tr, 14, tile);                    // epilogue: math done
    if constexpr (!OUT_F32) {
      fence_proxy_async();                         // staging writes -> visible to the TMA (async proxy)
      // two staging tiles: the store issued one tile ago has had this whole tile to drain; once the leader has
      // confirmed that, the barrier below also tells everyone that the OTHER tile may be overwritten next
      if (NSTG == 2 && !SPLIT && leader) tma_store_wait_read();
      named_bar_sync(bar_id, BAR_THREADS);
      if (leader) {
        if constexpr (ILV) {                       // output map dimensions are (C, W, N, H)
          if constexpr (POOL) tma_store_4d(map_y, smem_u32(stg), mt * TILE_M, w0 >> 1, n0, h0 >> 1);
          else tma_store_4d(map_y, smem_u32(stg), mt * TILE_M, w0, n0, h0);
        } else if constexpr (POOL) tma_store_4d(map_y, smem_u32(stg), mt * TILE_M, w0 >> 1, h0 >> 1, n0);
        else tma_store_4d(map_y, smem_u32(stg), mt * TILE_M, w0, h0, n0);
        tma_store_commit();
        trace(p.tr, 8, tile);                             // epilogue: tile stored
      }
    }
  }
  if constexpr (!OUT_F32) {
    if (leader) tma_store_wait_all();
  }
}

template <int KC, int STAGES, bool POOL, bool OUT_F32>
struct SmemLayout {
  static constexpr int A_BYTES = TILE_M * KC;
  static constexpr int B_BYTES = TILE_N * KC;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STG_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int BAR_OFFSET = STG_OFFSET + StageTile<POOL, OUT_F32>::BYTES;
  static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;   // barriers + alignment slack
};

#ifdef QNNB_WITH_V1
// ------------------------------------------------------------------ K1: the int8 implicit-GEMM kernel
// TW in {32, 16, 8} selects the pixel-tile geometry {TH, TW, TN}: {8,32,1}, {16,16,1}, {8,8,4}.
template <int KC, int STAGES, int TW, bool POOL, bool OUT_F32>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv3x3_i8_tc_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_x,
                     const __grid_constant__ CUtensorMap map_y, const TcParams p) {
  constexpr int TH = (TW == 32) ? 8 : (TW == 16 ? 16 : 8);
  constexpr int TN = (TW == 8) ? 4 : 1;
  static_assert(TH * TW * TN == TILE_N, "tile geometry");
  using SL = SmemLayout<KC, STAGES, POOL, OUT_F32>;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;     // swizzle atoms need 1024 B alignment
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + SL::BAR_OFFSET;
  // barrier slots (8 B each): full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2]; then the TMEM base word
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + SL::BAR_OFFSET + 8 * (2 * STAGES + 4));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_w);
    tma_prefetch_desc(&map_x);
    tma_prefetch_desc(&map_y);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), NUM_EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  if (threadIdx.x == 0) trace(p.tr, 1, 0);          // kernel entry (after barrier init / alloc issue)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  if (threadIdx.x == 0) trace(p.tr, 2, 0);          // setup done

  const int ksteps = 9 * p.kchunks;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {                               // single-threaded role (see tc_ptx.cuh)
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const TileCoord tc = decode_tile(p, tile);
        const int mt = tc.mt;
        const int n0 = tc.pt * TN, h0 = tc.th_i * TH, w0 = tc.tw_i * TW;
        for (int ks = 0; ks < ksteps; ++ks) {
          const int tap = ks / p.kchunks, kc = ks % p.kchunks;
          const int r = tap / 3, s = tap % 3;
          mbar_wait(empty_bar(stage), phase ^ 1u);
          if (ks == 0) trace(p.tr, 3, tile);                                 // producer: first load of a tile issued
          const uint32_t a_dst = smem_base + stage * SL::STAGE_BYTES;
          const uint32_t b_dst = a_dst + SL::A_BYTES;
          mbar_expect_tx(full_bar(stage), SL::STAGE_BYTES);
          tma_load_2d(a_dst, &map_w, full_bar(stage), tap * p.cin + kc * KC, mt * TILE_M);
          tma_load_4d(b_dst, &map_x, full_bar(stage), kc * KC, w0 + s - 1, h0 + r - 1, n0);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {                               // single-threaded role (see tc_ptx.cuh)
      constexpr uint32_t idesc = make_idesc_i8(TILE_M, TILE_N, true, true);
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);       // epilogue has drained this accumulator
        trace(p.tr, 4, tile);                             // MMA: accumulator available
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * TILE_N);
        for (int ks = 0; ks < ksteps; ++ks) {
          mbar_wait(full_bar(stage), phase);
          if (ks == 0) trace(p.tr, 5, tile);              // MMA: first operands of the tile landed
          tc_fence_after();
          const uint32_t a_addr = smem_base + stage * SL::STAGE_BYTES;
          const uint32_t b_addr = a_addr + SL::A_BYTES;
          const uint64_t a_desc = make_smem_desc<KC>(a_addr);
          const uint64_t b_desc = make_smem_desc<KC>(b_addr);
#pragma unroll
          for (int k = 0; k < KC / UMMA_K; ++k) {
            // advance both descriptors by k*32 bytes inside the swizzle atom (start-address field is in 16 B units)
            umma_i8(d_tmem, a_desc + (uint64_t)(k * (UMMA_K >> 4)), b_desc + (uint64_t)(k * (UMMA_K >> 4)), idesc,
                    (ks > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));                  // frees the smem slot when these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit(tfull_bar(acc));                      // accumulator complete
        trace(p.tr, 6, tile);                             // MMA: all MMAs of the tile issued
      }
    }
  } else if (warp >= 4 && warp < 4 + NUM_EPI_WARPS) {
    epilogue_role<TW, POOL, OUT_F32, /*FOLD*/ true, /*PITCH*/ TILE_M>(p, &map_y, tmem_base, tfull_bar(0), tempty_bar(0),
                                                                        smem_gen + SL::STG_OFFSET, warp, lane);
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) trace(p.tr, 9, 0);          // teardown
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

#endif  // QNNB_WITH_V1

// ------------------------------------------------------------------ K1 v2: halo-resident implicit GEMM
// v1 re-fetches the pixel tile once per filter tap (9x the input bytes through L2 -> SMEM), which the timeline
// trace showed to be the limiter (TMA supply, not the tensor pipe).  v2 loads each pixel tile ONCE per channel
// chunk WITH its one-pixel halo -- a single 4-D TMA box {KC, 10, TH+2, TN} at (c0, w0-1, h0-1, n0), again using
// the TMA zero fill for the SAME padding -- and feeds all nine taps from it: the B operand of tap (r,s) is the
// same shared-memory tile viewed through a descriptor whose start address is shifted by (r*10 + s) pixel rows.
// Tiles are 8 pixels wide so that the 8 rows of every UMMA core-matrix group are contiguous halo rows and the
// group stride (SBO) is one halo row of 10 pixels; TH x 8 pixels per image, TN images per tile (TH*TN = 32).
// The halo tile is stored [row][image][10 px] (the tensor map lists N before H), so that the 8-row groups of ALL
// TN images are uniformly strided and one N = 256 MMA per tap covers the whole tile (per-image N = 64/128 MMAs
// re-read the weight tile from shared memory and were SMEM-bandwidth bound).  Accumulator columns and the
// staged output are therefore in [row][image][px] order; the output tensor map is permuted the same way.
// Only the weights are streamed per tap (ring of A stages).
template <int KC, int TH, bool POOL, bool OUT_F32>
struct Smem2 {
  static constexpr int TN = 32 / TH;
  static constexpr int HALO_ROWS = TN * (TH + 2) * 10;
  static constexpr int HALO_BYTES = (HALO_ROWS * KC + 1023) / 1024 * 1024;
  static constexpr int HBUFS = 2;
  static constexpr int A_BYTES = TILE_M * KC;
  static constexpr int STG_BYTES = StageTile<POOL, OUT_F32>::BYTES;
  // pooled tiles stage 8 KB, so a second tile (one per epilogue group) is affordable; un-pooled ones (32 KB) keep the
  // weight ring deep instead and run the epilogue in lock-step
  static constexpr bool SPLIT = SPLIT_EPILOGUE && POOL && !OUT_F32;
  static constexpr int NSTG = SPLIT ? 2 : 1;
  static constexpr int BUDGET = 232448 - 1024 - 512 - NSTG * STG_BYTES - HBUFS * HALO_BYTES;
  // KC = 64: nine stages hold ALL taps of a 64-channel layer, which then stay resident (TcParams::resident)
  static constexpr int ACAP = (KC == 64) ? 9 : 8;
  static constexpr int ASTAGES = (BUDGET / A_BYTES) > ACAP ? ACAP : (BUDGET / A_BYTES);
  static constexpr int HALO_OFFSET = 0;
  static constexpr int A_OFFSET = HBUFS * HALO_BYTES;
  static constexpr int STG_OFFSET = A_OFFSET + ASTAGES * A_BYTES;
  static constexpr int BAR_OFFSET = STG_OFFSET + NSTG * STG_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + 512 + 1024;
  static_assert(ASTAGES >= 3, "not enough shared memory for the weight ring");
  static_assert(KC != 64 || ASTAGES == 9, "a 64-channel layer must be able to keep its nine weight tiles resident");
};

template <int KC, int TH, bool POOL, bool OUT_F32, bool SIGN = false>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv3x3_i8_tc2_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_x,
                      const __grid_constant__ CUtensorMap map_y, const TcParams p) {
  using SL = Smem2<KC, TH, POOL, OUT_F32>;
  constexpr int TN = SL::TN;
  constexpr int AS = SL::ASTAGES, HB = SL::HBUFS;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + SL::BAR_OFFSET;
  auto afull = [&](int s) { return bar_base + 8u * s; };
  auto aempty = [&](int s) { return bar_base + 8u * (AS + s); };
  auto hfull = [&](int b) { return bar_base + 8u * (2 * AS + b); };
  auto hempty = [&](int b) { return bar_base + 8u * (2 * AS + HB + b); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * AS + 2 * HB + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * AS + 2 * HB + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * AS + 2 * HB + 4);
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + SL::BAR_OFFSET + 8 * (2 * AS + 2 * HB + 4));
  if (threadIdx.x == 0) griddep_launch_dependents();
  // weight stages made ready so far (monotonic): written by the watcher thread, polled by the MMA issuer
  const uint32_t a_ready = tmem_slot + 8u;
  if (threadIdx.x == 0) *reinterpret_cast<volatile uint32_t*>(smem_gen + SL::BAR_OFFSET + 8 * (2 * AS + 2 * HB + 4) + 8) = 0u;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_w);
    tma_prefetch_desc(&map_x);
    tma_prefetch_desc(&map_y);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < AS; ++s) { mbar_init(afull(s), 1); mbar_init(aempty(s), 1); }
    for (int b = 0; b < HB; ++b) { mbar_init(hfull(b), 1); mbar_init(hempty(b), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), SL::SPLIT ? NUM_EPI_WARPS / 2 : NUM_EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  auto decode = [&](int tile, int& mt, int& n0, int& h0, int& w0) {
    const TileCoord tc = decode_tile(p, tile);
    mt = tc.mt;
    n0 = tc.pt * TN; h0 = tc.th_i * TH; w0 = tc.tw_i * 8;
  };

  if (warp == 0) {
    // ===================== halo producer: one box per (tile, channel chunk) =====================
    if (elect_one()) {                               // single-threaded role (see tc_ptx.cuh)
      griddep_wait();                                // the previous layer's output is complete and visible from here on;
                                                     // the weight ring (warp 3) starts filling before that
      int hb = 0; uint32_t hphase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        int mt, n0, h0, w0;
        decode(tile, mt, n0, h0, w0);
        for (int kc = 0; kc < p.kchunks; ++kc) {
          mbar_wait(hempty(hb), hphase ^ 1u);
          mbar_expect_tx(hfull(hb), SL::HALO_ROWS * KC);
          tma_load_4d(smem_base + SL::HALO_OFFSET + hb * SL::HALO_BYTES, &map_x, hfull(hb), kc * KC, w0 - 1, n0, h0 - 1);
          if (++hb == HB) { hb = 0; hphase ^= 1u; }
        }
      }
    }
  } else if (warp == 3) {
    // ===================== weight producer: one {KC, 128} box per (tile, chunk, tap) =====================
    if (elect_one()) {                               // single-threaded role (see tc_ptx.cuh)
      int as = 0; uint32_t aphase = 0;
      int issued = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        if (p.resident && tile != (int)blockIdx.x) break;         // resident weights: one pass fills the nine stages for good
        const int mt = tile - fdiv(tile, p.fd_m) * p.m_tiles;
        for (int kc = 0; kc < p.kchunks; ++kc) {
          for (int tap = 0; tap < 9; ++tap) {
            // the ring is filled once ahead of the previous kernel's completion (PDL); any wait that can actually block
            // happens after it, so the spin limit in mbar_wait never measures another kernel's run time
            if (issued == AS) griddep_wait();
            ++issued;
            mbar_wait(aempty(as), aphase ^ 1u);
            mbar_expect_tx(afull(as), SL::A_BYTES);
            tma_load_2d(smem_base + SL::A_OFFSET + as * SL::A_BYTES, &map_w, afull(as), tap * p.cin + kc * KC, mt * TILE_M);
            if (++as == AS) { as = 0; aphase ^= 1u; }
          }
        }
      }
    }
#ifndef QNNB_NO_WATCHER
  } else if (warp == 2) {
    // ===================== weight-stage watcher =====================
    // An mbarrier try_wait costs the waiting thread ~180 cycles even when the phase is already complete, and the
    // tcgen05.mma issue is itself blocking, so nine such waits per chunk in the MMA issuer left the tensor pipe idle
    // ~30 % of the time (cycle accounting, profiles/).  This otherwise idle thread absorbs the waits and publishes a
    // running count of landed weight stages; the issuer only polls that word (one shared-memory load).
    if (elect_one()) {                               // single-threaded role (see tc_ptx.cuh)
      int as = 0; uint32_t aphase = 0, count = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        if (p.resident && tile != (int)blockIdx.x) break;
        for (int ks = 0; ks < 9 * p.kchunks; ++ks) {
          mbar_wait(afull(as), aphase);
          st_release_shared(a_ready, ++count);
          if (++as == AS) { as = 0; aphase ^= 1u; }
        }
      }
    }
#endif
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {                               // single-threaded role (see tc_ptx.cuh)
      constexpr uint32_t idesc = make_idesc_i8(TILE_M, TILE_N, true, true);
      griddep_wait();                                // no global access here: turns the first halo wait into a hardware wait
      uint32_t a_need = 0, a_seen = 0;
      (void)a_need; (void)a_seen;
      // B view: rows of KC bytes, 8-row groups one halo row (10 pixels) apart
      constexpr uint64_t b_hi = ((uint64_t)((10 * KC) >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)(KC == 128 ? 2 : 4) << 61) | ((uint64_t)1 << 16);
      int as = 0; uint32_t aphase = 0;
      int hb = 0; uint32_t hphase = 0;
      int it = 0;
#ifdef QNNB_TRACE
      // cycle accounting of this thread (TRACE builds): where does the issue loop wait?
      long long c_tempty = 0, c_hfull = 0, c_afull = 0, c_t0 = clock64();
#define QNNB_CLK(var, stmt) { long long c0__ = clock64(); stmt; var += clock64() - c0__; }
#else
#define QNNB_CLK(var, stmt) { stmt; }
#endif
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
        QNNB_CLK(c_tempty, mbar_wait(tempty_bar(acc), acc_phase ^ 1u));
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * TILE_N);
        for (int kc = 0; kc < p.kchunks; ++kc) {
          QNNB_CLK(c_hfull, mbar_wait(hfull(hb), hphase));
          tc_fence_after();
          const uint32_t halo = smem_base + SL::HALO_OFFSET + hb * SL::HALO_BYTES;
          for (int tap = 0; tap < 9; ++tap) {
            const int r = tap / 3, s = tap - 3 * r;
#ifndef QNNB_NO_WATCHER
            if (!p.resident || it == 0) {
              ++a_need;
              QNNB_CLK(c_afull, { uint32_t spins = 0; while ((int)(a_seen - a_need) < 0) { a_seen = ld_acquire_shared(a_ready); if (++spins > (1u << 26)) __trap(); } });
            }
#else
            QNNB_CLK(c_afull, mbar_wait(afull(as), aphase));
#endif
            tc_fence_after();
            const uint64_t a_desc = make_smem_desc<KC>(smem_base + SL::A_OFFSET + as * SL::A_BYTES);
            // halo row (h + r) of every image starts r*TN halo rows further down; pixel shift s within the row
            const uint32_t b_addr = halo + (uint32_t)((r * TN * 10 + s) * KC);
            const uint64_t b_desc = b_hi | (uint64_t)((b_addr & 0x3FFFF) >> 4);
#pragma unroll
            for (int k = 0; k < KC / UMMA_K; ++k)
              umma_i8(d_tmem, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc, (kc > 0 || tap > 0 || k > 0) ? 1u : 0u);
            if (!p.resident) umma_commit(aempty(as));
            if (++as == AS) { as = 0; aphase ^= 1u; }
          }
          umma_commit(hempty(hb));                       // halo buffer free once these MMAs retire
          if (++hb == HB) { hb = 0; hphase ^= 1u; }
        }
        umma_commit(tfull_bar(acc));
      }
#ifdef QNNB_TRACE
      if (p.tr.buf != nullptr && blockIdx.x == 0) {
        unsigned long long* o = p.tr.buf + 15 * 1024;
        o[0] = 5; o[2] = (unsigned long long)c_tempty; o[3] = (unsigned long long)c_hfull; o[4] = (unsigned long long)c_afull;
        o[5] = 0; o[6] = (unsigned long long)(clock64() - c_t0); o[7] = (unsigned long long)it;
      }
#endif
    }
  } else if (warp >= 4 && warp < 4 + NUM_EPI_WARPS) {
    griddep_wait();                                  // (its stores follow the halo loads anyway; this makes it explicit)
    epilogue_role_n<8, POOL, OUT_F32, /*FOLD*/ true, /*PITCH*/ TILE_M, 1, TH, SL::NSTG, true, /*ILV*/ true, SIGN, SL::SPLIT>(
        p, &map_y, tmem_base, tfull_bar(0), tempty_bar(0), smem_gen + SL::STG_OFFSET, SL::STG_BYTES, warp, lane);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------ host side
struct Geometry { int tw, th, tn; };

#ifdef QNNB_WITH_V1
bool pick_geometry(int h, int w, Geometry* g) {
  if (w == 32 && h % 8 == 0) { *g = {32, 8, 1}; return true; }
  if (w == 16 && h % 16 == 0) { *g = {16, 16, 1}; return true; }
  if (w == 8 && h == 8) { *g = {8, 8, 4}; return true; }
  return false;
}

// tensor map of the int8 NHWC output for the epilogue's TMA store (box = one staged tile)
int make_output_map(EncodeTiledFn encode, CUtensorMap* my, void* y, int n, int oh, int ow, int cout, int pitch, const Geometry& g, bool pool) {
  cuuint64_t dims[4] = {(cuuint64_t)cout, (cuuint64_t)ow, (cuuint64_t)oh, (cuuint64_t)n};
  cuuint64_t strides[3] = {(cuuint64_t)cout, (cuuint64_t)ow * cout, (cuuint64_t)oh * ow * cout};
  cuuint32_t box[4] = {(cuuint32_t)pitch, (cuuint32_t)(pool ? g.tw / 2 : g.tw), (cuuint32_t)(pool ? g.th / 2 : g.th), (cuuint32_t)g.tn};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = encode(my, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, y, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("conv2d: cuTensorMapEncodeTiled(output) failed with %d", (int)r); return QNNB_ECUDA; }
  return QNNB_OK;
}

template <int KC, int STAGES, int TW, bool POOL, bool OUT_F32>
int launch_variant(const CUtensorMap& mw, const CUtensorMap& mx, const CUtensorMap& my, const TcParams& p, int grid, cudaStream_t st) {
  auto kern = conv3x3_i8_tc_kernel<KC, STAGES, TW, POOL, OUT_F32>;
  constexpr int smem = SmemLayout<KC, STAGES, POOL, OUT_F32>::TOTAL;
  static_assert(smem <= 232448, "shared memory budget");
  static SmemConfigured once;                     // per template instantiation, keyed by device inside
  QNNB_CUDA(once.ensure(kern, smem));
  kern<<<grid, NUM_THREADS, smem, st>>>(mw, mx, my, p);
  QNNB_CUDA(cudaGetLastError());
  return QNNB_OK;
}

template <int KC, int STAGES, int TW>
int launch_tw(const CUtensorMap& mw, const CUtensorMap& mx, const CUtensorMap& my, const TcParams& p, int grid, bool pool, bool f32, cudaStream_t st) {
  if (f32) return launch_variant<KC, STAGES, TW, false, true>(mw, mx, my, p, grid, st);
  if (pool) return launch_variant<KC, STAGES, TW, true, false>(mw, mx, my, p, grid, st);
  return launch_variant<KC, STAGES, TW, false, false>(mw, mx, my, p, grid, st);
}

template <int KC, int STAGES>
int launch_kc(const CUtensorMap& mw, const CUtensorMap& mx, const CUtensorMap& my, const TcParams& p, int grid, int tw, bool pool, bool f32, cudaStream_t st) {
  if (tw == 32) return launch_tw<KC, STAGES, 32>(mw, mx, my, p, grid, pool, f32, st);
  if (tw == 16) return launch_tw<KC, STAGES, 16>(mw, mx, my, p, grid, pool, f32, st);
  return launch_tw<KC, STAGES, 8>(mw, mx, my, p, grid, pool, f32, st);
}

#endif  // QNNB_WITH_V1

bool epilogue_ok(const qnnb_conv_desc& d, const char** why) {
  if (d.epi.res_kind != QNNB_KIND_NONE) { *why = "residual epilogue not on the tensor-core path"; return false; }
  if (d.epi.act == QNNB_ACT_QUANT || d.epi.act == QNNB_ACT_SIGN_I8) return true;
  if (d.epi.act == QNNB_ACT_NONE && d.epi.pool == 0) return true;
  *why = "epilogue must be quantized_tanh / binary_tanh to int8 levels (optionally pooled) or plain fp32";
  return false;
}

// ---- v2 (halo-resident) launch
template <int KC, int TH, bool POOL, bool OUT_F32, bool SIGN = false>
int launch_v2_variant(const CUtensorMap& mw, const CUtensorMap& mx, const CUtensorMap& my, const TcParams& p, int grid, cudaStream_t st) {
  auto kern = conv3x3_i8_tc2_kernel<KC, TH, POOL, OUT_F32, SIGN>;
  constexpr int smem = Smem2<KC, TH, POOL, OUT_F32>::TOTAL;
  static_assert(smem <= 232448, "shared memory budget");
  static SmemConfigured once;                     // per template instantiation, keyed by device inside
  QNNB_CUDA(once.ensure(kern, smem));
  QNNB_CUDA(launch_pdl(kern, dim3(grid), dim3(NUM_THREADS), (size_t)smem, st, mw, mx, my, p));
  return QNNB_OK;
}

template <int KC, int TH>
int launch_v2_th(const CUtensorMap& mw, const CUtensorMap& mx, const CUtensorMap& my, const TcParams& p, int grid, bool pool, bool f32, cudaStream_t st) {
  if (f32) return launch_v2_variant<KC, TH, false, true>(mw, mx, my, p, grid, st);
  if (p.epi.act == QNNB_ACT_SIGN_I8) {
    if (pool) return launch_v2_variant<KC, TH, true, false, true>(mw, mx, my, p, grid, st);
    return launch_v2_variant<KC, TH, false, false, true>(mw, mx, my, p, grid, st);
  }
  if (pool) return launch_v2_variant<KC, TH, true, false>(mw, mx, my, p, grid, st);
  return launch_v2_variant<KC, TH, false, false>(mw, mx, my, p, grid, st);
}

template <int KC>
int launch_v2_kc(const CUtensorMap& mw, const CUtensorMap& mx, const CUtensorMap& my, const TcParams& p, int grid, int th, bool pool, bool f32, cudaStream_t st) {
  if (th == 32) return launch_v2_th<KC, 32>(mw, mx, my, p, grid, pool, f32, st);
  if (th == 16) return launch_v2_th<KC, 16>(mw, mx, my, p, grid, pool, f32, st);
  return launch_v2_th<KC, 8>(mw, mx, my, p, grid, pool, f32, st);
}

bool pick_geometry_v2(int h, int w, Geometry* g) {
  if (w % 8 != 0) return false;
  if (h % 32 == 0) { *g = {8, 32, 1}; return true; }
  if (h % 16 == 0) { *g = {8, 16, 2}; return true; }
  if (h % 8 == 0) { *g = {8, 8, 4}; return true; }
  return false;
}

int launch_conv_tc_v2(const qnnb_conv_desc& d, const void* x, const void* w, void* y, cudaStream_t st) {
  EncodeTiledFn encode = get_encode();
  if (!encode) { set_error("conv2d: cuTensorMapEncodeTiled is not available from the driver"); return QNNB_ECUDA; }
  Geometry g;
  pick_geometry_v2(d.h, d.w, &g);
  const int KC = (d.cin % 128 == 0) ? 128 : 64;
  const CUtensorMapSwizzle swz = (KC == 128) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  const bool pool = d.epi.pool == 2;
  const bool f32 = d.epi.act == QNNB_ACT_NONE;
  CUtensorMap mw, mx, my;
  memset(&my, 0, sizeof(my));
  {
    cuuint64_t dims[2] = {(cuuint64_t)9 * d.cin, (cuuint64_t)d.cout};
    cuuint64_t strides[1] = {(cuuint64_t)9 * d.cin};
    cuuint32_t box[2] = {(cuuint32_t)KC, (cuuint32_t)TILE_M};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&mw, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(w), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("conv2d: cuTensorMapEncodeTiled(weights) failed with %d", (int)r); return QNNB_ECUDA; }
  }
  {
    // dimension order (C, W, N, H): the box lands as [row][image][10 px][KC]
    cuuint64_t dims[4] = {(cuuint64_t)d.cin, (cuuint64_t)d.w, (cuuint64_t)d.n, (cuuint64_t)d.h};
    cuuint64_t strides[3] = {(cuuint64_t)d.cin, (cuuint64_t)d.h * d.w * d.cin, (cuuint64_t)d.w * d.cin};
    cuuint32_t box[4] = {(cuuint32_t)KC, 10u, (cuuint32_t)g.tn, (cuuint32_t)(g.th + 2)};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(&mx, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<void*>(x), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("conv2d: cuTensorMapEncodeTiled(halo) failed with %d", (int)r); return QNNB_ECUDA; }
  }
  if (!f32) {
    const int oh = pool ? d.h / 2 : d.h, ow = pool ? d.w / 2 : d.w;
    cuuint64_t dims[4] = {(cuuint64_t)d.cout, (cuuint64_t)ow, (cuuint64_t)d.n, (cuuint64_t)oh};
    cuuint64_t strides[3] = {(cuuint64_t)d.cout, (cuuint64_t)oh * ow * d.cout, (cuuint64_t)ow * d.cout};
    cuuint32_t box[4] = {(cuuint32_t)TILE_M, (cuuint32_t)(pool ? 4 : 8), (cuuint32_t)g.tn, (cuuint32_t)(pool ? g.th / 2 : g.th)};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(&my, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, y, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("conv2d: cuTensorMapEncodeTiled(output) failed with %d", (int)r); return QNNB_ECUDA; }
  }
  TcParams p;
  p.tr = g_trace;
  p.n = d.n; p.h = d.h; p.w = d.w; p.cin = d.cin; p.cout = d.cout;
  p.tiles_w = d.w / 8;
  p.tiles_h = d.h / g.th;
  p.tiles_n = ceil_div(d.n, g.tn);
  p.m_tiles = d.cout / TILE_M;
  p.num_tiles = p.tiles_w * p.tiles_h * p.tiles_n * p.m_tiles;
  p.fd_m = make_fastdiv(p.m_tiles); p.fd_w = make_fastdiv(p.tiles_w); p.fd_h = make_fastdiv(p.tiles_h);
  p.kchunks = d.cin / KC;
  p.out_pitch = TILE_M;
  p.y = y;
  p.epi = make_epi(d.epi);
  const int grid = p.num_tiles < grid_sms(d.max_ctas) ? p.num_tiles : grid_sms(d.max_ctas);
  // Cin = 64: all nine taps (72 KB) fit the nine-stage weight ring; if every tile of a CTA uses the same channel tile they
  // are loaded once and stay (no per-tap barrier traffic afterwards)
  p.resident = (KC == 64 && p.kchunks == 1 && (p.m_tiles == 1 || grid % p.m_tiles == 0) && getenv("QNNB_NO_RESIDENT") == nullptr) ? 1 : 0;
  if (KC == 128) return launch_v2_kc<128>(mw, mx, my, p, grid, g.th, pool, f32, st);
  return launch_v2_kc<64>(mw, mx, my, p, grid, g.th, pool, f32, st);
}

}  // namespace

void set_trace_buffer(unsigned long long* buf, int cap) { g_trace.buf = buf; g_trace.cap = cap; }
unsigned long long* get_trace_buffer() { return g_trace.buf; }

bool conv_tc_v1_supported(const qnnb_conv_desc& d) {
#ifdef QNNB_WITH_V1
  Geometry g;
  return d.in_kind == QNNB_KIND_I8 && d.epi.act != QNNB_ACT_SIGN_I8 && pick_geometry(d.h, d.w, &g);
#else
  (void)d;
  return false;
#endif
}

bool conv_tc_supported(const qnnb_conv_desc& d, const char** why) {
  Geometry g;
  // tile indices must stay below 2^24 (FastDiv): at most 8 tiles per image on any supported shape
  if ((long long)d.n * ((d.h + 7) / 8) * ((d.w + 7) / 8) * ((d.cout + 127) / 128) >= (1ll << 24)) { *why = "batch too large for one launch"; return false; }
  if (conv_first_tc_shape(d)) return epilogue_ok(d, why);
  if (d.in_kind != QNNB_KIND_I8) { *why = "input must be int8 levels (or uint8 32-wide RGB for the first layer)"; return false; }
  if (d.kh != 3 || d.kw != 3 || d.stride != 1) { *why = "only 3x3 stride 1"; return false; }
  if (d.cin % 64 != 0 || d.cin > 256) { *why = "Cin must be 64, 128, 192 or 256"; return false; }
  if (d.cout % 128 != 0) { *why = "Cout must be a multiple of 128"; return false; }
  if (!pick_geometry_v2(d.h, d.w, &g)) { *why = "spatial size must be a multiple of 8 in both directions"; return false; }
  if ((d.epi.act == QNNB_ACT_QUANT || d.epi.act == QNNB_ACT_SIGN_I8) && !is_pow2_scale(d.epi.acc_scale)) { *why = "acc_scale must be a power of two"; return false; }
  return epilogue_ok(d, why);
}

int launch_conv_tc(const qnnb_conv_desc& d, const void* x, const void* w, void* y, cudaStream_t st) {
  if (conv_first_tc_shape(d)) return launch_conv_first_tc(d, x, w, y, st);
#ifndef QNNB_WITH_V1
  // production build: the halo-resident kernel serves every supported shape (the first-generation kernel -- one TMA
  // box per filter tap -- is only compiled with -DQNNB_WITH_V1 for A/B profiling)
  if (d.impl == QNNB_IMPL_TCGEN05_V1) { set_error("conv2d: this build does not contain the v1 kernel (make EXTRA=-DQNNB_WITH_V1)"); return QNNB_EUNSUPPORTED; }
  return launch_conv_tc_v2(d, x, w, y, st);
#else
  // Kernel choice (measured on B200, profiles/): the halo-resident kernel wins when a tile is one 32-row block
  // (one N = 256 MMA per tap: 32-row maps, -8..-20 %); with 16- or 8-row maps it needs N = 128 / 64 MMAs that re-read
  // the weight tile from shared memory per image and becomes SMEM-bandwidth bound, so v1 keeps those shapes.
  Geometry g1;
  const bool v1_ok = pick_geometry(d.h, d.w, &g1);
  const bool want_v2 = (d.impl != QNNB_IMPL_TCGEN05_V1);
  if (want_v2 || !v1_ok) return launch_conv_tc_v2(d, x, w, y, st);
  EncodeTiledFn encode = get_encode();
  if (!encode) { set_error("conv2d: cuTensorMapEncodeTiled is not available from the driver"); return QNNB_ECUDA; }
  Geometry g;
  pick_geometry(d.h, d.w, &g);
  const int KC = (d.cin % 128 == 0) ? 128 : 64;
  const CUtensorMapSwizzle swz = (KC == 128) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  const bool pool = d.epi.pool == 2;
  const bool f32 = d.epi.act == QNNB_ACT_NONE;

  CUtensorMap mw, mx, my;
  memset(&my, 0, sizeof(my));
  {
    cuuint64_t dims[2] = {(cuuint64_t)9 * d.cin, (cuuint64_t)d.cout};
    cuuint64_t strides[1] = {(cuuint64_t)9 * d.cin};
    cuuint32_t box[2] = {(cuuint32_t)KC, (cuuint32_t)TILE_M};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&mw, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(w), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("conv2d: cuTensorMapEncodeTiled(weights) failed with %d", (int)r); return QNNB_ECUDA; }
  }
  {
    cuuint64_t dims[4] = {(cuuint64_t)d.cin, (cuuint64_t)d.w, (cuuint64_t)d.h, (cuuint64_t)d.n};
    cuuint64_t strides[3] = {(cuuint64_t)d.cin, (cuuint64_t)d.w * d.cin, (cuuint64_t)d.h * d.w * d.cin};
    cuuint32_t box[4] = {(cuuint32_t)KC, (cuuint32_t)g.tw, (cuuint32_t)g.th, (cuuint32_t)g.tn};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(&mx, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<void*>(x), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("conv2d: cuTensorMapEncodeTiled(activations) failed with %d", (int)r); return QNNB_ECUDA; }
  }
  if (!f32) {
    int rc = make_output_map(encode, &my, y, d.n, pool ? d.h / 2 : d.h, pool ? d.w / 2 : d.w, d.cout, TILE_M, g, pool);
    if (rc) return rc;
  }

  TcParams p;
  p.tr = g_trace;
  p.n = d.n; p.h = d.h; p.w = d.w; p.cin = d.cin; p.cout = d.cout;
  p.tiles_w = d.w / g.tw;
  p.tiles_h = d.h / g.th;
  p.tiles_n = ceil_div(d.n, g.tn);
  p.m_tiles = d.cout / TILE_M;
  p.num_tiles = p.tiles_w * p.tiles_h * p.tiles_n * p.m_tiles;
  p.fd_m = make_fastdiv(p.m_tiles); p.fd_w = make_fastdiv(p.tiles_w); p.fd_h = make_fastdiv(p.tiles_h);
  p.kchunks = d.cin / KC;
  p.resident = 0;
  p.out_pitch = TILE_M;
  p.y = y;
  p.epi = make_epi(d.epi);

  const int grid = p.num_tiles < grid_sms(d.max_ctas) ? p.num_tiles : grid_sms(d.max_ctas);
  if (KC == 128) return launch_kc<128, 4>(mw, mx, my, p, grid, g.tw, pool, f32, st);
  return launch_kc<64, 6>(mw, mx, my, p, grid, g.tw, pool, f32, st);
#endif
}

}  // namespace qnnb
