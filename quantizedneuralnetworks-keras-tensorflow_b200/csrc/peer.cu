// NVLink logit path: a device buffer of one process mapped into its sibling processes on the
// same box with CUDA IPC, so that the final dense kernel of every batch shard stores its logits straight into the
// gathering rank's memory over NVLink / NVSwitch -- no per-step collective.  Plumbing only: allocation and mapping.
#include "common.cuh"

#include <string.h>

using namespace qnnb;

extern "C" {

int qnnb_peer_alloc(int64_t bytes, void** ptr, void* handle) {
  QNNB_CHECK_ARG(bytes > 0 && ptr && handle, "peer_alloc: bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == QNNB_PEER_HANDLE_BYTES, "handle size");
  void* p = nullptr;
  QNNB_CUDA(cudaMalloc(&p, (size_t)bytes));
  cudaError_t e = cudaMemset(p, 0, (size_t)bytes);
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) { cudaFree(p); return cuda_fail(e, "peer_alloc"); }
  memcpy(handle, &h, sizeof(h));
  *ptr = p;
  return QNNB_OK;
}

int qnnb_peer_open(const void* handle, void** ptr) {
  QNNB_CHECK_ARG(handle && ptr, "peer_open: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  // the caller's current device is the GPU that will write: peer access to the owner is enabled lazily
  QNNB_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return QNNB_OK;
}

int qnnb_peer_close(void* ptr) {
  if (ptr) QNNB_CUDA(cudaIpcCloseMemHandle(ptr));
  return QNNB_OK;
}

int qnnb_peer_free(void* ptr) {
  if (ptr) QNNB_CUDA(cudaFree(ptr));
  return QNNB_OK;
}

}  // extern "C"
