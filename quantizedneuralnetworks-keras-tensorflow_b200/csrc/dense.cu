// K2 -- classifier head: Dense + bias (+ BatchNormalization | softmax), one warp per image.
//
// Stands in for QuantizedDense.call (layers/quantized_layers.py:79-88), BinaryDense.call
// (layers/binary_layers.py:78-85), TernaryDense.call (layers/ternary_layers.py:77-84) followed by
// the BatchNormalization of models/vgg.py:42 or the softmax of models/resnet.py:137.  The
// reference's AveragePooling2D(8)+Flatten (models/resnet.py:134-135) is folded in by the host:
// the packed kernel is replicated over the 64 pooled pixels and acc_scale carries the 1/64.
//
// HBM-bound (reads fin bytes per image, writes 40): x is read once, coalesced, 128 B per warp
// per step; the packed kernel (<= 40 KB) stays in L1/L2.
#include "common.cuh"
#include "tc_ptx.cuh"

namespace qnnb {

namespace {

constexpr int UG = 16;    // units accumulated per pass
constexpr int MAX_UNITS = 32;

struct DenseP {
  int n, fin, units, kwords, fin_pad;
  const void* x;
  const void* w;
  float* y;
  float* logits;
  int softmax;
  int w_f32;         // fp32 input: the packed kernel holds fp32 values (QNNB_WFMT_F32)
  Epi epi;
};

template <int KIND>
__global__ void __launch_bounds__(256)
dense_kernel(const DenseP p) {
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int warps_per_grid = gridDim.x * 8;
  for (int img = blockIdx.x * 8 + warp; img < p.n; img += warps_per_grid) {
    float z = 0.f;                       // lane u < units ends up with unit u's pre-activation
    for (int u0 = 0; u0 < p.units; u0 += UG) {
      const int ug = min(UG, p.units - u0);
      if constexpr (KIND == QNNB_KIND_F32) {
        float acc[UG];
#pragma unroll
        for (int u = 0; u < UG; ++u) acc[u] = 0.f;
        const float* xr = (const float*)p.x + (long long)img * p.fin;
        for (int k = lane; k < p.fin; k += 32) {
          const float xv = __ldg(xr + k);
#pragma unroll
          for (int u = 0; u < UG; ++u)
            if (u < ug) {
              const long long wi = (long long)(u0 + u) * p.fin_pad + k;
              acc[u] = fmaf(xv, p.w_f32 ? __ldg((const float*)p.w + wi) : (float)__ldg((const int8_t*)p.w + wi), acc[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < UG; ++u) {
          float v = acc[u];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
          if (u < ug && lane == u0 + u) z = v;
        }
      } else {
        int acc[UG];
#pragma unroll
        for (int u = 0; u < UG; ++u) acc[u] = 0;
        const uint32_t* xr = (const uint32_t*)p.x + (long long)img * p.kwords;
        for (int k = lane; k < p.kwords; k += 32) {
          const uint32_t xv = __ldg(xr + k);
#pragma unroll
          for (int u = 0; u < UG; ++u) {
            if (u < ug) {
              const uint32_t wv = __ldg((const uint32_t*)p.w + (long long)(u0 + u) * p.kwords + k);
              if constexpr (KIND == QNNB_KIND_B1) acc[u] += __popc(xv ^ wv);
              else acc[u] = __dp4a((int)xv, (int)wv, acc[u]);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < UG; ++u) {
          int v = acc[u];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
          if constexpr (KIND == QNNB_KIND_B1) v = p.fin - 2 * v;
          if (u < ug && lane == u0 + u) z = (float)v;      // cvt.rn
        }
      }
    }
    const bool active = lane < p.units;
    ChanConst cc = load_chan(p.epi, lane, active);
    z = affine(z, cc);
    if (p.softmax) {
      if (active && p.logits) p.logits[(long long)img * p.units + lane] = z;
      float m = active ? z : -INFINITY;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      float ex = active ? expf(z - m) : 0.f;
      float s = ex;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      z = ex / s;
    }
    if (active) p.y[(long long)img * p.units + lane] = z;
  }
}

// Fast path (int8 / bit-packed inputs, units <= 16, K a multiple of 16 bytes): the packed kernel is staged
// once per CTA into shared memory by ONE bulk-async copy (cp.async.bulk + mbarrier) while every lane already has
// its first XB 16-byte vectors of the image in flight (512 B per warp per load, fully coalesced), so the kernel
// pays one memory latency, and HBM traffic is exactly one read of x.
__device__ __forceinline__ uint32_t dsmem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int KIND>
__global__ void __launch_bounds__(256)
dense_smem_kernel(const DenseP p) {
  extern __shared__ __align__(128) uint4 swv[];
  __shared__ __align__(8) uint64_t wbar;
  const int kvec = p.kwords >> 2;
  const uint32_t bar = dsmem_u32(&wbar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    griddep_launch_dependents();
    const uint32_t bytes = (uint32_t)(p.units * kvec) * 16u;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    constexpr uint32_t CH = 16384;
    for (uint32_t off = 0; off < bytes; off += CH) {
      const uint32_t sz = bytes - off < CH ? bytes - off : CH;
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(dsmem_u32(swv) + off), "l"((const char*)p.w + off), "r"(sz), "r"(bar) : "memory");
    }
  }
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const bool active = lane < p.units;
  const ChanConst cc = load_chan(p.epi, lane, active);
  constexpr int XB = 8;
  bool weights_ready = false;
  griddep_wait();                        // x is the previous layer's output; the kernel weights above are not
  for (int img = blockIdx.x * 8 + warp; img < p.n; img += gridDim.x * 8) {
    int acc[UG];
#pragma unroll
    for (int u = 0; u < UG; ++u) acc[u] = 0;
    const uint4* xr = reinterpret_cast<const uint4*>(p.x) + (long long)img * kvec;
    for (int kv0 = lane; kv0 < kvec; kv0 += 32 * XB) {
      uint4 xb[XB];
#pragma unroll
      for (int j = 0; j < XB; ++j) {
        const int kv = kv0 + 32 * j;
        xb[j] = (kv < kvec) ? __ldg(xr + kv) : make_uint4(0, 0, 0, 0);
      }
      if (!weights_ready) {                      // warp-uniform: first pass only
        uint32_t done = 0;
        while (!done)
          asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                       : "=r"(done) : "r"(bar), "r"(0u) : "memory");
        weights_ready = true;
      }
#pragma unroll
      for (int j = 0; j < XB; ++j) {
        const int kv = kv0 + 32 * j;
        if (kv < kvec) {
          const uint4 xv = xb[j];
#pragma unroll
          for (int u = 0; u < UG; ++u) {
            if (u < p.units) {
              const uint4 wv = swv[u * kvec + kv];
              if constexpr (KIND == QNNB_KIND_B1) {
                acc[u] += __popc(xv.x ^ wv.x) + __popc(xv.y ^ wv.y) + __popc(xv.z ^ wv.z) + __popc(xv.w ^ wv.w);
              } else {
                acc[u] = __dp4a((int)xv.x, (int)wv.x, acc[u]);
                acc[u] = __dp4a((int)xv.y, (int)wv.y, acc[u]);
                acc[u] = __dp4a((int)xv.z, (int)wv.z, acc[u]);
                acc[u] = __dp4a((int)xv.w, (int)wv.w, acc[u]);
              }
            }
          }
        }
      }
    }
    float z = 0.f;
#pragma unroll
    for (int u = 0; u < UG; ++u) {
      int v = acc[u];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if constexpr (KIND == QNNB_KIND_B1) v = p.fin - 2 * v;
      if (lane == u) z = (float)v;
    }
    z = affine(z, cc);
    if (p.softmax) {
      if (active && p.logits) p.logits[(long long)img * p.units + lane] = z;
      float m = active ? z : -INFINITY;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      float ex = active ? expf(z - m) : 0.f;
      float sum = ex;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      z = ex / sum;
    }
    if (active) p.y[(long long)img * p.units + lane] = z;
  }
}

// Global average pooling + dense on fp32 maps (models/resnet.py:134-140): one warp per image sums the P positions
// per channel (coalesced 128-byte rows, eight independent loads in flight), then takes the [units][fin] dot products.
// HBM-bound: reads P * fin floats per image exactly once.
__global__ void __launch_bounds__(256)
dense_avgpool_f32_kernel(const DenseP p, int positions) {
  if (threadIdx.x == 0) griddep_launch_dependents();
  griddep_wait();
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  constexpr int MAXJ = 8;                                    // fin <= 256
  const int nj = (p.fin + 31) / 32;
  const bool active = lane < p.units;
  const ChanConst cc = load_chan(p.epi, lane, active);
  for (int img = blockIdx.x * 8 + warp; img < p.n; img += gridDim.x * 8) {
    float s[MAXJ];
#pragma unroll
    for (int j = 0; j < MAXJ; ++j) s[j] = 0.f;
    const float* xr = (const float*)p.x + (long long)img * positions * p.fin;
    for (int pos0 = 0; pos0 < positions; pos0 += 8) {
#pragma unroll
      for (int j = 0; j < MAXJ; ++j) {
        if (j < nj) {
          const int c = lane + 32 * j;
          float v[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) v[q] = (c < p.fin && pos0 + q < positions) ? __ldg(xr + (long long)(pos0 + q) * p.fin + c) : 0.f;
#pragma unroll
          for (int q = 0; q < 8; ++q) s[j] += v[q];
        }
      }
    }
    float z = 0.f;
    for (int u = 0; u < p.units; ++u) {
      float part = 0.f;
#pragma unroll
      for (int j = 0; j < MAXJ; ++j) {
        const int c = lane + 32 * j;
        if (j < nj && c < p.fin) {
          const long long wi = (long long)u * p.fin_pad + c;
          part = fmaf(s[j], p.w_f32 ? __ldg((const float*)p.w + wi) : (float)__ldg((const int8_t*)p.w + wi), part);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      if (lane == u) z = part;
    }
    z = affine(z, cc);
    if (p.softmax) {
      if (active && p.logits) p.logits[(long long)img * p.units + lane] = z;
      float m = active ? z : -INFINITY;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      float ex = active ? expf(z - m) : 0.f;
      float sum = ex;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      z = ex / sum;
    }
    if (active) p.y[(long long)img * p.units + lane] = z;
  }
}

}  // namespace

int launch_dense(const qnnb_dense_desc& d, const void* x, const void* w, float* y, float* logits, cudaStream_t st) {
  QNNB_CHECK_ARG(d.units >= 1 && d.units <= MAX_UNITS, "dense: units=%d outside 1..%d", d.units, MAX_UNITS);
  DenseP p;
  p.n = d.n; p.fin = d.fin; p.units = d.units;
  p.fin_pad = (d.fin + 3) / 4 * 4;
  if (d.in_kind == QNNB_KIND_B1) p.kwords = (d.fin + 31) / 32;
  else p.kwords = p.fin_pad / 4;
  QNNB_CHECK_ARG(d.in_kind != QNNB_KIND_I8 || (d.fin & 3) == 0, "dense: int8 input needs fin %% 4 == 0 (got %d)", d.fin);
  p.x = x; p.w = w; p.y = y; p.logits = logits; p.softmax = d.softmax; p.w_f32 = d.w_f32;
  p.epi = make_epi(d.epi);
  int blocks = ceil_div(d.n, 8);
  if (blocks < 1) blocks = 1;
  if (d.avg_positions > 1) {
    QNNB_CHECK_ARG(d.in_kind == QNNB_KIND_F32 && d.fin <= 256, "dense: avg_positions needs fp32 input with fin <= 256 (got kind %d, fin %d)", d.in_kind, d.fin);
    if (blocks > tcx::grid_sms(d.max_ctas) * 8) blocks = tcx::grid_sms(d.max_ctas) * 8;
    QNNB_CUDA(launch_pdl(dense_avgpool_f32_kernel, dim3(blocks), dim3(256), (size_t)0, st, p, (int)d.avg_positions));
    return QNNB_OK;
  }
  const size_t wbytes = (size_t)d.units * p.kwords * 4;
  if (d.in_kind != QNNB_KIND_F32 && d.units <= UG && (p.kwords & 3) == 0 && wbytes <= 160 * 1024) {
    const int cap = tcx::grid_sms(d.max_ctas) * 2;
    int fb = blocks > cap ? cap : blocks;
    if (d.in_kind == QNNB_KIND_I8) {
      QNNB_CUDA(cudaFuncSetAttribute(dense_smem_kernel<QNNB_KIND_I8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wbytes));
      QNNB_CUDA(launch_pdl(dense_smem_kernel<QNNB_KIND_I8>, dim3(fb), dim3(256), wbytes, st, p));
    } else {
      QNNB_CUDA(cudaFuncSetAttribute(dense_smem_kernel<QNNB_KIND_B1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wbytes));
      QNNB_CUDA(launch_pdl(dense_smem_kernel<QNNB_KIND_B1>, dim3(fb), dim3(256), wbytes, st, p));
    }
    QNNB_CUDA(cudaGetLastError());
    return QNNB_OK;
  }
  if (blocks > tcx::grid_sms(d.max_ctas) * 8) blocks = tcx::grid_sms(d.max_ctas) * 8;
  switch (d.in_kind) {
    case QNNB_KIND_I8: dense_kernel<QNNB_KIND_I8><<<blocks, 256, 0, st>>>(p); break;
    case QNNB_KIND_B1: dense_kernel<QNNB_KIND_B1><<<blocks, 256, 0, st>>>(p); break;
    case QNNB_KIND_F32: dense_kernel<QNNB_KIND_F32><<<blocks, 256, 0, st>>>(p); break;
    default: set_error("dense: bad in_kind %d", d.in_kind); return QNNB_EINVAL;
  }
  QNNB_CUDA(cudaGetLastError());
  return QNNB_OK;
}

}  // namespace qnnb
