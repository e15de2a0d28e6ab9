// peer.cu — NVLink peer memory for the lattice-T split (SURVEY §8e secondary partitioning, BASELINE config 5).
//
// The reference moves a t-halo per hop and per eigenvector through QUDA's exchangeGhost on host-staged MPI buffers
// (/root/reference/lib/contract_wrappers.cu:166-174) and an extended gauge field with exchangeExtendedGhost
// (lib/displace.cpp:104-134).  Here every rank keeps its eigenvector slabs in the extended layout
// [vector][parity][t = 0 .. Tl+2H)[V3/2 sites][12 complex] in ONE device allocation that its two time neighbours map
// through CUDA IPC (one process per GPU, NVSwitch: every peer at full NVLink bandwidth), and the boundary slices of a
// whole eigenvector batch go STRAIGHT into the neighbour's halo slices:
//   mode 0: one strided 2-D copy on the copy engines (rows = (vector, parity) blocks, equal pitch) - no SM, no staging
//           buffer, no pack/unpack pass; this is what overlaps best with the FP64-bound fused kernel;
//   mode 1: an SM push kernel with 128-bit peer stores (for comparison and for drivers without P2P DMA).
// Ordering across ranks (the neighbour's kernels may read a halo only after the push has landed) is the caller's:
// mugiq_b200/tsplit.py puts a one-element NCCL all-reduce behind the push on the same stream.
#include <algorithm>
#include <cstring>

#include "kernels.cuh"

namespace mugiq_b200 {

__global__ void __launch_bounds__(256)
halo_push_kernel(uint4 *__restrict__ dst, const uint4 *__restrict__ src, const long long pitch16, const long long width16,
                 const long long total16) {
  // element i -> (row, column) of the 2-D block; consecutive threads store consecutive 16-byte words of a row
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total16; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / width16, col = i - row * width16;
    dst[row * pitch16 + col] = __ldg(src + row * pitch16 + col);
  }
}

}  // namespace mugiq_b200

using namespace mugiq_b200;

extern "C" {

int mugiq_b200_peer_alloc(void **ptr_d, long long bytes, void *handle64) {
  const char *who = "mugiq_b200_peer_alloc";
  if (!ptr_d || !handle64 || bytes <= 0) return set_error(MUGIQ_B200_EINVAL, "%s: bad argument", who);
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  *ptr_d = nullptr;
  if (cudaMalloc(ptr_d, (size_t)bytes) != cudaSuccess) {
    cudaGetLastError();
    return set_error(MUGIQ_B200_ENOMEM, "%s: cannot allocate %lld bytes", who, bytes);
  }
  cudaIpcMemHandle_t h;
  MUGIQ_CUDA_CHECK(cudaIpcGetMemHandle(&h, *ptr_d));
  memcpy(handle64, &h, sizeof(h));
  return MUGIQ_B200_OK;
}

int mugiq_b200_peer_open(void **ptr_d, const void *handle64) {
  if (!ptr_d || !handle64) return set_error(MUGIQ_B200_EINVAL, "mugiq_b200_peer_open: bad argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  MUGIQ_CUDA_CHECK(cudaIpcOpenMemHandle(ptr_d, h, cudaIpcMemLazyEnablePeerAccess));
  return MUGIQ_B200_OK;
}

int mugiq_b200_peer_close(void *ptr_d) {
  if (ptr_d) MUGIQ_CUDA_CHECK(cudaIpcCloseMemHandle(ptr_d));
  return MUGIQ_B200_OK;
}

int mugiq_b200_peer_free(void *ptr_d) {
  if (ptr_d) MUGIQ_CUDA_CHECK(cudaFree(ptr_d));
  return MUGIQ_B200_OK;
}

int mugiq_b200_halo_push_t(void *dst_slabs_d, const void *src_slabs_d, int first_vec, int nvec, int Lt_ext, long long V3h,
                           int site_bytes, int src_t, int dst_t, int nslices, int mode, void *stream) {
  const char *who = "mugiq_b200_halo_push_t";
  if (!dst_slabs_d || !src_slabs_d) return set_error(MUGIQ_B200_EINVAL, "%s: NULL slab pointer", who);
  if (first_vec < 0 || nvec < 1 || Lt_ext < 1 || V3h < 1 || site_bytes < 16 || site_bytes % 16 || nslices < 1 || src_t < 0 ||
      dst_t < 0 || src_t + nslices > Lt_ext || dst_t + nslices > Lt_ext)
    return set_error(MUGIQ_B200_EINVAL, "%s: bad geometry (vectors %d+%d, Lt_ext %d, slices %d -> %d x %d)", who, first_vec,
                     nvec, Lt_ext, src_t, dst_t, nslices);
  const size_t slice = (size_t)V3h * site_bytes;
  const size_t pitch = (size_t)Lt_ext * slice;  // distance between consecutive (vector, parity) blocks
  const size_t width = (size_t)nslices * slice;
  const size_t rows = (size_t)2 * nvec;
  const char *src = static_cast<const char *>(src_slabs_d) + (size_t)2 * first_vec * pitch + (size_t)src_t * slice;
  char *dst = static_cast<char *>(dst_slabs_d) + (size_t)2 * first_vec * pitch + (size_t)dst_t * slice;
  ProfScope prof(K_HALO_PUSH, (cudaStream_t)stream, 2.0 * (double)rows * (double)width);
  if (mode == 0) {
    MUGIQ_CUDA_CHECK(cudaMemcpy2DAsync(dst, pitch, src, pitch, width, rows, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  } else {
    const long long total16 = (long long)(rows * width / 16);
    const int blocks = (int)std::min<long long>((total16 + 255) / 256, 148LL * 8);
    halo_push_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<uint4 *>(dst), reinterpret_cast<const uint4 *>(src),
                                                               (long long)(pitch / 16), (long long)(width / 16), total16);
    MUGIQ_LAUNCH_CHECK();
  }
  return MUGIQ_B200_OK;
}

}  // extern "C"
