// Peer-mappable device memory for the mailboxes of the sharded batch-wide criterion (include/admm_b200.h,
// "batch-wide criterion of a batch sharded over the GPUs of one box").  Plain CUDA IPC: the owner allocates
// with cudaMalloc and exports a 64-byte handle, every other process of the box maps it (peer access over
// NVLink is enabled lazily by the driver).  Host-synchronous set-up calls; nothing here runs per iteration.
#include "common.cuh"

using namespace admm;

extern "C" {

int admm_peer_alloc(size_t bytes, void** devptr, unsigned char* handle_host) {
  ADMM_REQUIRE(devptr != nullptr && handle_host != nullptr && bytes > 0, ADMM_EINVAL, "admm_peer_alloc: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e == cudaSuccess) e = cudaMemset(p, 0, bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    if (p) cudaFree(p);
    cudaGetLastError();
    set_error("admm_peer_alloc: %s", cudaGetErrorString(e));
    return ADMM_ECUDA;
  }
  memcpy(handle_host, &h, sizeof(h));
  *devptr = p;
  return ADMM_OK;
}

int admm_peer_open(const unsigned char* handle_host, void** devptr) {
  ADMM_REQUIRE(devptr != nullptr && handle_host != nullptr, ADMM_EINVAL, "admm_peer_open: bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle_host, sizeof(h));
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("admm_peer_open: %s (are the ranks on one box with peer access?)", cudaGetErrorString(e));
    return ADMM_ECUDA;
  }
  *devptr = p;
  return ADMM_OK;
}

int admm_peer_close(void* devptr) {
  if (devptr == nullptr) return ADMM_OK;
  cudaError_t e = cudaIpcCloseMemHandle(devptr);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("admm_peer_close: %s", cudaGetErrorString(e));
    return ADMM_ECUDA;
  }
  return ADMM_OK;
}

int admm_peer_free(void* devptr) {
  if (devptr == nullptr) return ADMM_OK;
  cudaError_t e = cudaFree(devptr);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("admm_peer_free: %s", cudaGetErrorString(e));
    return ADMM_ECUDA;
  }
  return ADMM_OK;
}

}  // extern "C"
