// tmap_probe.cu — does cuTensorMapEncodeTiled accept a tensor whose strides are NOT nested (components before 8-site
// chunks), and where does a (16 doubles, 12 components, NB chunks, 1 parity) box of a QUDA FLOAT2 spinor land in shared memory?
// Prints one JSON line.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tools/tmap_probe tools/tmap_probe.cu -lcudart
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void probe(const __grid_constant__ CUtensorMap tm, int chunk0, int parity, int nb, double *out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem);
  double *dst = reinterpret_cast<double *>(smem + 1024);
  const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(bar), dst_a = (uint32_t)__cvta_generic_to_shared(dst);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t bytes = 16 * 8 * 12 * nb;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst_a),
        "l"(&tm), "r"(0), "r"(0), "r"(chunk0), "r"(parity), "r"(bar_a)
        : "memory");
  }
  asm volatile(
      "{\n.reg .pred P1;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n@P1 bra D;\nbra W;\nD:\n}" ::"r"(bar_a)
      : "memory");
  for (int i = threadIdx.x; i < 16 * 12 * nb; i += blockDim.x) out[i] = dst[i];
}

int main() {
  EncodeFn encode = nullptr;
  cudaDriverEntryPointQueryResult qr;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&encode, cudaEnableDefault, &qr) != cudaSuccess || !encode) {
    printf("{\"error\": \"no cuTensorMapEncodeTiled\"}\n");
    return 1;
  }
  const int volumeCB = 4096;  // sites per parity; value of component c of site x of parity p: p*1e6 + c*1e4 + x (re), negative (im)
  std::vector<double> h((size_t)2 * 12 * volumeCB * 2);
  for (int p = 0; p < 2; p++)
    for (int c = 0; c < 12; c++)
      for (int x = 0; x < volumeCB; x++) {
        h[(((size_t)p * 12 + c) * volumeCB + x) * 2] = p * 1e6 + c * 1e4 + x;
        h[(((size_t)p * 12 + c) * volumeCB + x) * 2 + 1] = -(p * 1e6 + c * 1e4 + x);
      }
  double *d, *out;
  cudaMalloc(&d, h.size() * 8);
  cudaMalloc(&out, 16 * 12 * 4 * 8);
  cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
  printf("{");
  for (int variant = 0; variant < 2; variant++) {
    // variant 0: dims (16 doubles, 12 comps, chunks, parity) - strides NOT nested; variant 1: (16, chunks, 12 comps, parity)
    for (int nb = 1; nb <= 4; nb += 3) {
      CUtensorMap tm;
      cuuint64_t dim[4], str[3];
      cuuint32_t box[4], es[4] = {1, 1, 1, 1};
      if (variant == 0) {
        dim[0] = 16; dim[1] = 12; dim[2] = volumeCB / 8; dim[3] = 2;
        str[0] = (cuuint64_t)volumeCB * 16; str[1] = 128; str[2] = (cuuint64_t)12 * volumeCB * 16;
        box[0] = 16; box[1] = 12; box[2] = nb; box[3] = 1;
      } else {
        dim[0] = 16; dim[1] = volumeCB / 8; dim[2] = 12; dim[3] = 2;
        str[0] = 128; str[1] = (cuuint64_t)volumeCB * 16; str[2] = (cuuint64_t)12 * volumeCB * 16;
        box[0] = 16; box[1] = nb; box[2] = 12; box[3] = 1;
      }
      CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, d, dim, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      printf("\"v%d_nb%d_encode\": %d, ", variant, nb, (int)r);
      if (r != CUDA_SUCCESS) continue;
      cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
      const int chunk0 = 5, parity = 1;
      if (variant == 0)
        probe<<<1, 128, 1024 + 16 * 8 * 12 * 4>>>(tm, chunk0, parity, nb, out);
      else {
        // coordinates follow the dim order: (0, chunk0, 0, parity)
        CUtensorMap tm2 = tm;
        (void)tm2;
        probe<<<1, 128, 1024 + 16 * 8 * 12 * 4>>>(tm, 0, parity, nb, out);  // see note: variant 1 passes chunk as coord 1 below
      }
      cudaError_t e = cudaDeviceSynchronize();
      std::vector<double> o(16 * 12 * nb);
      cudaMemcpy(o.data(), out, o.size() * 8, cudaMemcpyDeviceToHost);
      // expected layout for variant 0: [chunk][comp][8 sites][re,im]
      int bad = 0;
      if (variant == 0)
        for (int k = 0; k < nb; k++)
          for (int c = 0; c < 12; c++)
            for (int s = 0; s < 8; s++) {
              const double want = parity * 1e6 + c * 1e4 + (chunk0 + k) * 8 + s;
              if (o[((k * 12 + c) * 8 + s) * 2] != want || o[((k * 12 + c) * 8 + s) * 2 + 1] != -want) bad++;
            }
      printf("\"v%d_nb%d_run\": \"%s\", \"v%d_nb%d_mismatches\": %d, ", variant, nb, cudaGetErrorString(e), variant, nb,
             variant == 0 ? bad : -1);
    }
  }
  printf("\"done\": 1}\n");
  return 0;
}
