// kernels.cuh — internal (C++) interface between the C-ABI layer (cabi.cu) and the per-stage kernels.
#pragma once
#include "common.cuh"
#include "prof.cuh"

namespace mugiq_b200 {

constexpr int kMaxBatch = 96;  // eigenvector pairs per launch of the batched kernels (kernel-parameter budget)

// Pointer table handed to kernels by value (no device-side argument struct to cudaMalloc/cudaMemcpy/
// cudaFree per call as in /root/reference/lib/contract_wrappers.cu:93-114).
struct VecBatch {
  const void *vL[kMaxBatch];
  const void *vR[kMaxBatch];
  double inv_sigma[kMaxBatch];
  int nvec;
};

// sigma is cast to Float before inversion (lib/loop_mugiq.cpp:479) and inverted in double
// (include/contract_util.cuh:133).
inline double inv_sigma_of(double sigma, int precision) {
  const double s = (precision == MUGIQ_B200_PREC_SINGLE) ? (double)(float)sigma : sigma;
  return 1.0 / s;
}

inline size_t prec_bytes(int precision) { return precision == MUGIQ_B200_PREC_DOUBLE ? 8 : 4; }

int check_entries(const mugiq_b200_disp_entry_t *entries, int nentries, const char *who);  // cabi.cu

// stage 1
int contract_batch(void *loop_d, const void *const *vL, const void *const *vR, const double *sigma, int nvec,
                   int accumulate, const LatGeom &g, int precision, cudaStream_t stream);
// stage 2
int displace(void *dst_d, const void *src_d, const void *gauge_d, int dir, int sign, const LatGeom &g,
             int precision, cudaStream_t stream);
int displace_batch(void *const *dst_d, const void *const *src_d, int nvec, const void *gauge_d, int dir, int sign,
                   const LatGeom &g, int precision, cudaStream_t stream);
// stages 1 and 2 on fields in QUDA's native FLOAT2 / FLOAT4 orders (native_order.cu)
int contract_batch_native(void *loop_d, const void *const *vL, const void *const *vR, const double *sigma, int nvec, int order,
                          int accumulate, const LatGeom &g, int precision, cudaStream_t stream);
int displace_batch_native(void *const *dst_d, const void *const *src_d, int nvec, const void *gauge_d, int dir, int sign,
                          int order, const LatGeom &g, int precision, cudaStream_t stream);
// stages 1+2 fused
long long loop_workspace_bytes(const LatGeom &g, int precision, int nvec, const mugiq_b200_disp_entry_t *entries,
                               int nentries);
int loop_accumulate(void *dataPos_d, const void *const *evec_d, const double *sigma_h, int nvec, const void *gauge_d,
                    const mugiq_b200_disp_entry_t *entries, int nentries, int accumulate, void *workspace_d,
                    const LatGeom &g, int precision, cudaStream_t stream);
// stage 3
int reorder_mapgamma(void *out_d, const void *in_d, int nLoop, const LatGeom &g, int precision, cudaStream_t stream);
// stage 4
int phase_matrix(void *phase_d, const int *mom_h, int Nmom, int ftsign, const int localL[4], const int totalL[4],
                 const int commCoord[4], int precision, cudaStream_t stream);
long long momproj_workspace_bytes(long long M, int N, long long K, int precision);
int momproj(void *mom_d, const void *posMP_d, const void *phase_d, long long M, int N, long long K, int precision,
            void *workspace_d, cudaStream_t stream);
// stages 3+4 fused (momproj_pos.cu)
int phase_matrix_eo(void *phase_d, const int *mom_h, int Nmom, int ftsign, const int localL[4], const int totalL[4],
                    const int commCoord[4], int precision, cudaStream_t stream);
long long momproj_pos_workspace_bytes(const LatGeom &g, int nLoop, int N, int precision);
int momproj_pos(void *mom_d, const void *pos_d, const void *phase_eo_d, int nLoop, int N, const LatGeom &g, int precision,
                void *workspace_d, cudaStream_t stream);
// layout conversion
int convert_spinor(void *dst_d, const void *src_d, int order, bool to_site, const LatGeom &g, int precision,
                   cudaStream_t stream);
int convert_spinor_batch(void *const *dst_d, const void *const *src_d, int n, int order, bool to_site, const LatGeom &g,
                         int precision, cudaStream_t stream);

}  // namespace mugiq_b200
