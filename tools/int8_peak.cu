// Micro-benchmark: dense int8 tensor-core ceiling of this GPU for the exact instruction the conv kernels
// issue (tcgen05.mma.cta_group::1.kind::i8, M=128, N=256, K=32, operands resident in shared memory with the
// 128B swizzle, int32 accumulators in TMEM).  One CTA per SM, one thread issues back-to-back MMAs on random
// operand bytes (data-dependent power), no loads, no epilogue.  Prints a JSON line; bench.py uses
// profiles/int8_peak.json as the roofline denominator for the tensor-bound layers.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// SWIZZLE_64B K-major (64-byte rows, 512-byte atoms) and un-swizzled interleaved (8x16 B core matrices) variants,
// to compare the tensor pipe's operand-fetch rate per layout
__device__ __forceinline__ uint64_t desc_sw64(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}
__device__ __forceinline__ uint64_t desc_none(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// 0 = SWIZZLE_128B, 1 = SWIZZLE_64B, 2 = no swizzle (interleaved core matrices);
// 3/4/5 = the same three layouts with the B operand read through a SHIFTED view (start one pixel row in, 8-row groups
// 10 rows apart), i.e. what the halo-resident conv kernel issues for its off-centre filter taps
template <int MODE>
__global__ void __launch_bounds__(128, 1) peak_kernel(int groups, uint32_t seed) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* gen = raw + (base - smem_u32(raw));
  __shared__ uint64_t bar;
  __shared__ uint64_t bar2;          // completed once at start: parity-0 waits on it succeed immediately
  __shared__ uint32_t tslot;
  // A: 128 x 128 B, B: 256 x 128 B, two copies each (ping-pong like a real pipeline)
  const int total = 2 * (128 + 256) * 128;
  uint32_t s = seed ^ (blockIdx.x * 2654435761u) ^ (threadIdx.x * 40503u);
  for (int i = threadIdx.x * 4; i < total; i += 128 * 4) {
    s = s * 1664525u + 1013904223u;
    *reinterpret_cast<uint32_t*>(gen + i) = s;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2)) : "memory");
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar2)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tslot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tslot;
  if (warp == 0 && lane == 0) {
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
    uint32_t phase = 0;
    for (int g = 0; g < groups; ++g) {
      // one "k-step" group = 16 MMAs (K = 512) into alternating accumulators
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (MODE >= 6) {
          const int every = (MODE == 7) ? 4 : 2;
          if (i % every == 0) {
            uint32_t done = 0;
            while (!done) {
              if (MODE == 9)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(smem_u32(&bar2)), "r"(0u) : "memory");
              else
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(smem_u32(&bar2)), "r"(0u) : "memory");
            }
            if (MODE != 8) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          }
        }
        const uint32_t buf = (i >> 2) & 1;
        uint64_t a, b;
        if (MODE == 0 || MODE >= 6) {
          a = desc_sw128(base + buf * (128 * 128)) + (uint64_t)((i & 3) * 2);
          b = desc_sw128(base + 2 * 128 * 128 + buf * (256 * 128)) + (uint64_t)((i & 3) * 2);
        } else if (MODE == 1) {
          a = desc_sw64(base + buf * (128 * 128) + ((i >> 1) & 1) * (128 * 64)) + (uint64_t)((i & 1) * 2);
          b = desc_sw64(base + 2 * 128 * 128 + buf * (256 * 128) + ((i >> 1) & 1) * (256 * 64)) + (uint64_t)((i & 1) * 2);
        } else if (MODE == 2) {
          a = desc_none(base + buf * (128 * 128) + (i & 3) * 2 * 2048, 2048, 128);
          b = desc_none(base + 2 * 128 * 128 + buf * (256 * 128) + (i & 3) * 2 * 4096, 4096, 128);
        } else if (MODE == 3) {
          // B: 340 halo rows of 128 B (43.5 KB at base + 32 KB), view shifted by (i % 3) rows, group stride 10 rows
          a = desc_sw128(base + buf * (128 * 128)) + (uint64_t)((i & 3) * 2);
          uint64_t d = (uint64_t)(((base + 2 * 128 * 128 + (i % 3) * 128 + (i % 5) * 1280) & 0x3FFFF) >> 4);
          d |= (uint64_t)1 << 16; d |= (uint64_t)(1280 >> 4) << 32; d |= (uint64_t)1 << 46; d |= (uint64_t)2 << 61;
          b = d + (uint64_t)((i & 3) * 2);
        } else if (MODE == 4) {
          a = desc_sw64(base + buf * (128 * 128) + ((i >> 1) & 1) * (128 * 64)) + (uint64_t)((i & 1) * 2);
          uint64_t d = (uint64_t)(((base + 2 * 128 * 128 + (i % 3) * 64 + (i % 5) * 640) & 0x3FFFF) >> 4);
          d |= (uint64_t)1 << 16; d |= (uint64_t)(640 >> 4) << 32; d |= (uint64_t)1 << 46; d |= (uint64_t)4 << 61;
          b = d + (uint64_t)((i & 1) * 2);
        } else {
          // chunk-major un-swizzled halo: [16-byte chunk][340 px rows][16 B]; pixel shift = 16 B, group stride 10 px = 160 B
          a = desc_none(base + buf * (128 * 128) + (i & 3) * 2 * 2048, 2048, 128);
          b = desc_none(base + 2 * 128 * 128 + (i & 3) * 2 * 5440 + (i % 3) * 16 + (i % 5) * 160, 5440, 160);
        }
        const uint32_t d = tmem + ((g & 1) ? 256u : 0u);
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"((uint32_t)(i > 0))
            : "memory");
      }
      if (MODE >= 6) {
        // per-"tap" bookkeeping of the conv kernels: wait on an (already complete) mbarrier + tcgen05 fence,
        // MODE 6 every 2 MMAs, MODE 7 every 4 MMAs, MODE 8 = wait only (no fence), MODE 9 = test_wait instead of try_wait
      }
      if ((g & 7) == 7 || g == groups - 1) {
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t done = 0;
        while (!done) {
          asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                       : "=r"(done) : "r"(smem_u32(&bar)), "r"(phase) : "memory");
        }
        phase ^= 1u;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

template <int MODE>
void run(const char* name, int groups, int reps, int sms, int smem) {
  cudaFuncSetAttribute(peak_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  peak_kernel<MODE><<<sms, 128, smem>>>(groups, 1u);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(cudaGetLastError())); exit(1); }
  double best = 1e30;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(e0);
    peak_kernel<MODE><<<sms, 128, smem>>>(groups, 2u + r);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  cudaEventRecord(e0);
  for (int r = 0; r < reps * 10; ++r) peak_kernel<MODE><<<sms, 128, smem>>>(groups, 100u + r);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms_all; cudaEventElapsedTime(&ms_all, e0, e1);
  const double total = ms_all / (reps * 10);
  const double ops = (double)sms * groups * 16.0 * 2.0 * 128 * 256 * 32;
  if (MODE == 0)
    printf("{\"int8_tops\": %.1f, \"int8_tops_sustained\": %.1f, \"sms\": %d, \"kernel_ms_best\": %.4f, \"kernel_ms_sustained\": %.4f, "
           "\"how\": \"tcgen05.mma.cta_group::1.kind::i8 M=128 N=256 K=32 from swizzled smem, random operand bytes, %d MMAs per SM per launch; burst = best of %d launches, sustained = %d back-to-back launches\"}\n",
           ops / (best * 1e-3) / 1e12, ops / (total * 1e-3) / 1e12, sms, best, total, groups * 16, reps, reps * 10);
  else
    fprintf(stderr, "layout %s: burst %.1f TOP/s, sustained %.1f TOP/s\n", name, ops / (best * 1e-3) / 1e12, ops / (total * 1e-3) / 1e12);
}

// In-process entry (bench.py loads tools/libint8peak.so with ctypes): the SWIZZLE_128B measurement only, on the current
// device and its primary context.  Returns 0 on success.
extern "C" int qnnb_int8_peak(int groups, int reps, double* burst_tops, double* sustained_tops) {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 1;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int smem = 2 * (128 + 256) * 128 + 1024;
  if (cudaFuncSetAttribute(peak_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return 2;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  peak_kernel<0><<<sms, 128, smem>>>(groups, 1u);
  if (cudaDeviceSynchronize() != cudaSuccess) return 3;
  double best = 1e30;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(e0);
    peak_kernel<0><<<sms, 128, smem>>>(groups, 2u + r);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  cudaEventRecord(e0);
  for (int r = 0; r < reps * 4; ++r) peak_kernel<0><<<sms, 128, smem>>>(groups, 100u + r);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms_all; cudaEventElapsedTime(&ms_all, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (cudaGetLastError() != cudaSuccess) return 4;
  const double ops = (double)sms * groups * 16.0 * 2.0 * 128 * 256 * 32;
  if (burst_tops) *burst_tops = ops / (best * 1e-3) / 1e12;
  if (sustained_tops) *sustained_tops = ops / ((double)ms_all / (reps * 4) * 1e-3) / 1e12;
  return 0;
}

#ifndef QNNB_PEAK_LIB
int main(int argc, char** argv) {
  int groups = argc > 1 ? atoi(argv[1]) : 4000;
  int reps = argc > 2 ? atoi(argv[2]) : 20;
  int dev = 0, sms = 0;
  cudaSetDevice(dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int smem = 2 * (128 + 256) * 128 + 1024;
  run<3>("SWIZZLE_128B, shifted halo view", groups, reps / 4 + 1, sms, smem);
  run<4>("SWIZZLE_64B, shifted halo view", groups, reps / 4 + 1, sms, smem);
  run<5>("SWIZZLE_NONE chunk-major, shifted halo view", groups, reps / 4 + 1, sms, smem);
  run<6>("SW128 + try_wait+fence every 2 MMAs", groups, reps / 4 + 1, sms, smem);
  run<7>("SW128 + try_wait+fence every 4 MMAs", groups, reps / 4 + 1, sms, smem);
  run<8>("SW128 + try_wait (no fence) every 2 MMAs", groups, reps / 4 + 1, sms, smem);
  run<9>("SW128 + test_wait+fence every 2 MMAs", groups, reps / 4 + 1, sms, smem);
  run<1>("SWIZZLE_64B", groups, reps / 4 + 1, sms, smem);
  run<2>("SWIZZLE_NONE (interleaved)", groups, reps / 4 + 1, sms, smem);
  run<0>("SWIZZLE_128B", groups, reps, sms, smem);
  return 0;
}
#endif
