// cabi.cu — the extern "C" boundary declared in include/mugiq_b200.h: argument validation, error
// reporting and dispatch into the per-stage kernels.  No torch types, no C++ types in any signature.
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "kernels.cuh"

namespace mugiq_b200 {

static thread_local char g_last_error[512] = "";

int set_error(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
  return code;
}

int check_geom(const mugiq_b200_geom_t *geom, const char *who) {
  if (!geom) return set_error(MUGIQ_B200_EINVAL, "%s: geom is NULL", who);
  for (int i = 0; i < 4; i++)
    if (geom->L[i] < 1) return set_error(MUGIQ_B200_EINVAL, "%s: L[%d] = %d must be positive", who, i, geom->L[i]);
  if (geom->L[0] & 1)
    return set_error(MUGIQ_B200_EINVAL, "%s: L[0] = %d must be even (even/odd site order)", who, geom->L[0]);
  if (geom->precision != MUGIQ_B200_PREC_SINGLE && geom->precision != MUGIQ_B200_PREC_DOUBLE)
    return set_error(MUGIQ_B200_EINVAL, "%s: precision %d not supported (4 = single, 8 = double)", who,
                     geom->precision);
  const long long v = (long long)geom->L[0] * geom->L[1] * geom->L[2] * geom->L[3];
  if (v > 0x3fffffffLL) return set_error(MUGIQ_B200_EINVAL, "%s: local volume %lld too large", who, v);
  return MUGIQ_B200_OK;
}

// Displacements need every extent even: the neighbour of a site is looked up as (checkerboard index, opposite parity),
// which breaks across the periodic wrap of an odd extent (the wrapped site keeps its parity).  QUDA, and with it the
// reference, only runs on even local extents.  Contraction, reorder and projection are site-local and take any Ly, Lz, Lt.
int check_geom_even(const mugiq_b200_geom_t *geom, const char *who) {
  for (int i = 1; i < 4; i++)
    if (geom->L[i] & 1)
      return set_error(MUGIQ_B200_EINVAL, "%s: L[%d] = %d must be even for displacements (even/odd neighbour lookup)", who, i,
                       geom->L[i]);
  return MUGIQ_B200_OK;
}

static int check_dir_sign(int dir, int sign, const char *who) {
  // Displace::setupDisplacement rejects anything else (lib/displace.cpp:214-222)
  if (dir < 0 || dir > 3) return set_error(MUGIQ_B200_EINVAL, "%s: invalid displacement direction %d", who, dir);
  if (sign != 0 && sign != 1) return set_error(MUGIQ_B200_EINVAL, "%s: invalid displacement sign %d", who, sign);
  return MUGIQ_B200_OK;
}

int check_entries(const mugiq_b200_disp_entry_t *entries, int nentries, const char *who) {
  if (nentries < 0 || nentries > MUGIQ_B200_MAX_ENTRIES)
    return set_error(MUGIQ_B200_EINVAL, "%s: nentries = %d out of range [0,%d]", who, nentries, MUGIQ_B200_MAX_ENTRIES);
  if (nentries > 0 && !entries) return set_error(MUGIQ_B200_EINVAL, "%s: entries is NULL", who);
  for (int e = 0; e < nentries; e++) {
    int rc = check_dir_sign(entries[e].dir, entries[e].sign, who);
    if (rc) return rc;
    // LoopComputeParam swaps start/stop with a warning before the loop runs (include/loop_mugiq.h:234-239);
    // at this level the entry must already be ordered
    if (entries[e].start > entries[e].stop)
      return set_error(MUGIQ_B200_EINVAL, "%s: entry %d has start %d > stop %d", who, e, entries[e].start,
                       entries[e].stop);
  }
  return MUGIQ_B200_OK;
}

}  // namespace mugiq_b200

using namespace mugiq_b200;

#define REQUIRE_PTR(p, who) \
  if (!(p)) return set_error(MUGIQ_B200_EINVAL, "%s: %s is NULL", who, #p)

extern "C" {

int mugiq_b200_version(void) { return MUGIQ_B200_VERSION; }

const char *mugiq_b200_last_error(void) { return g_last_error; }

int mugiq_b200_device_info(char *name, int name_len, int *sm_count, int *cc, long long *free_bytes,
                           long long *total_bytes) {
  int dev = 0;
  MUGIQ_CUDA_CHECK(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  MUGIQ_CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
  if (name && name_len > 0) {
    strncpy(name, prop.name, name_len - 1);
    name[name_len - 1] = 0;
  }
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc) *cc = prop.major * 10 + prop.minor;
  size_t f = 0, t = 0;
  MUGIQ_CUDA_CHECK(cudaMemGetInfo(&f, &t));
  if (free_bytes) *free_bytes = (long long)f;
  if (total_bytes) *total_bytes = (long long)t;
  return MUGIQ_B200_OK;
}

int mugiq_b200_gamma_tables(double *row_value, int *column_index, double *map_sign, int *map_index) {
  constexpr GammaTables gt = gamma_tables();
  static const double re[4] = {1, 0, -1, 0}, im[4] = {0, 1, 0, -1};
  for (int G = 0; G < 16; G++) {
    for (int s = 0; s < 4; s++) {
      if (row_value) {
        row_value[(G * 4 + s) * 2 + 0] = re[gt.ipow[G][s]];
        row_value[(G * 4 + s) * 2 + 1] = im[gt.ipow[G][s]];
      }
      if (column_index) column_index[G * 4 + s] = gt.col[G][s];
    }
    if (map_sign) map_sign[G] = gt.map_sign[G];
    if (map_index) map_index[G] = gt.map_index[G];
  }
  return MUGIQ_B200_OK;
}

static int check_order(int order, int precision, const char *who) {
  if (order == MUGIQ_B200_ORDER_FLOAT2 || order == MUGIQ_B200_ORDER_FLOAT4 || order == MUGIQ_B200_ORDER_SITE)
    return MUGIQ_B200_OK;
  (void)precision;
  return set_error(MUGIQ_B200_EINVAL, "%s: unknown colour-spinor order %d", who, order);
}

int mugiq_b200_ingest_spinor(void *dst_site_d, const void *src_d, int src_order, const mugiq_b200_geom_t *geom,
                             void *stream) {
  const char *who = "mugiq_b200_ingest_spinor";
  int rc = check_geom(geom, who);
  if (rc) return rc;
  REQUIRE_PTR(dst_site_d, who);
  REQUIRE_PTR(src_d, who);
  if ((rc = check_order(src_order, geom->precision, who))) return rc;
  const LatGeom g = make_geom(geom->L);
  if (src_order == MUGIQ_B200_ORDER_SITE) {
    MUGIQ_CUDA_CHECK(cudaMemcpyAsync(dst_site_d, src_d, (size_t)g.volume * 24 * prec_bytes(geom->precision),
                                     cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return MUGIQ_B200_OK;
  }
  return convert_spinor(dst_site_d, src_d, src_order, true, g, geom->precision, (cudaStream_t)stream);
}

int mugiq_b200_ingest_spinor_batch(void *const *dst_site_d, const void *const *src_d, int nfields, int src_order,
                                   const mugiq_b200_geom_t *geom, void *stream) {
  const char *who = "mugiq_b200_ingest_spinor_batch";
  int rc = check_geom(geom, who);
  if (rc) return rc;
  REQUIRE_PTR(dst_site_d, who);
  REQUIRE_PTR(src_d, who);
  if (nfields < 0) return set_error(MUGIQ_B200_EINVAL, "%s: nfields = %d", who, nfields);
  if (src_order != MUGIQ_B200_ORDER_FLOAT2 && src_order != MUGIQ_B200_ORDER_FLOAT4)
    return set_error(MUGIQ_B200_EINVAL, "%s: source order must be FLOAT2 or FLOAT4 (got %d)", who, src_order);
  for (int i = 0; i < nfields; i++)
    if (!dst_site_d[i] || !src_d[i]) return set_error(MUGIQ_B200_EINVAL, "%s: field %d is NULL", who, i);
  return convert_spinor_batch(dst_site_d, src_d, nfields, src_order, true, make_geom(geom->L), geom->precision,
                              (cudaStream_t)stream);
}

int mugiq_b200_export_spinor(void *dst_d, int dst_order, const void *src_site_d, const mugiq_b200_geom_t *geom,
                             void *stream) {
  const char *who = "mugiq_b200_export_spinor";
  int rc = check_geom(geom, who);
  if (rc) return rc;
  REQUIRE_PTR(dst_d, who);
  REQUIRE_PTR(src_site_d, who);
  if ((rc = check_order(dst_order, geom->precision, who))) return rc;
  const LatGeom g = make_geom(geom->L);
  if (dst_order == MUGIQ_B200_ORDER_SITE) {
    MUGIQ_CUDA_CHECK(cudaMemcpyAsync(dst_d, src_site_d, (size_t)g.volume * 24 * prec_bytes(geom->precision),
                                     cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return MUGIQ_B200_OK;
  }
  return convert_spinor(dst_d, src_site_d, dst_order, false, g, geom->precision, (cudaStream_t)stream);
}

int mugiq_b200_gauge_upload(void *gauge_d, const void *const gauge_h[4], const mugiq_b200_geom_t *geom,
                            void *stream) {
  const char *who = "mugiq_b200_gauge_upload";
  int rc = check_geom(geom, who);
  if (rc) return rc;
  REQUIRE_PTR(gauge_d, who);
  REQUIRE_PTR(gauge_h, who);
  const LatGeom g = make_geom(geom->L);
  const size_t dir_bytes = (size_t)g.volume * kLinkLen * 2 * prec_bytes(geom->precision);
  for (int mu = 0; mu < 4; mu++) {
    if (!gauge_h[mu]) return set_error(MUGIQ_B200_EINVAL, "%s: gauge_h[%d] is NULL", who, mu);
    MUGIQ_CUDA_CHECK(cudaMemcpyAsync(static_cast<char *>(gauge_d) + mu * dir_bytes, gauge_h[mu], dir_bytes,
                                     cudaMemcpyHostToDevice, (cudaStream_t)stream));
  }
  MUGIQ_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
  return MUGIQ_B200_OK;
}

int mugiq_b200_contract(void *loop_d, const void *vL_d, const void *vR_d, double sigma,
                        const mugiq_b200_geom_t *geom, void *stream) {
  const char *who = "mugiq_b200_contract";
  int rc = check_geom(geom, who);
  if (rc) return rc;
  REQUIRE_PTR(loop_d, who);
  REQUIRE_PTR(vL_d, who);
  REQUIRE_PTR(vR_d, who);
  const LatGeom g = make_geom(geom->L);
  // vL == vR is the ultra-local call of the reference (fineEvecR is a copy of fineEvecL,
  // lib/loop_mugiq.cpp:501-502): read the field once.
  return contract_batch(loop_d, &vL_d, (vL_d == vR_d) ? nullptr : &vR_d, &sigma, 1, 1, g, geom->precision,
                        (cudaStream_t)stream);
}

int mugiq_b200_contract_batch(void *loop_d, const void *const *vL_d, const void *const *vR_d, const double *sigma_h,
                              int nvec, int accumulate, const mugiq_b200_geom_t *geom, void *stream) {
  const char *who = "mugiq_b200_contract_batch";
  int rc = check_geom(geom, who);
  if (rc) return rc;
  REQUIRE_PTR(loop_d, who);
  REQUIRE_PTR(vL_d, who);
  REQUIRE_PTR(sigma_h, who);
  if (nvec < 0) return set_error(MUGIQ_B200_EINVAL, "%s: nvec = %d", who, nvec);
  const LatGeom g = make_geom(geom->L);
  if (nvec == 0) {
    if (!accumulate)
      MUGIQ_CUDA_CHECK(cudaMemsetAsync(loop_d, 0, (size_t)16 * g.volume * 2 * prec_bytes(geom->precision),
                                       (cudaStream_t)stream));
    return MUGIQ_B200_OK;
  }
  for (int i = 0; i < nvec; i++) {
    if (!vL_d[i] || (vR_d && !vR_d[i])) return set_error(MUGIQ_B200_EINVAL, "%s: eigenvector %d is NULL", who, i);
  }
  return contract_batch(loop_d, vL_d, vR_d, sigma_h, nvec, accumulate, g, geom->precision, (cudaStream_t)stream);
}

int mugiq_b200_displace(void *dst_d, const void *src_d, const void *gauge_d, int dir, int sign,
                        const mugiq_b200_geom_t *geom, void *stream) {
  const char *who = "mugiq_b200_displace";
  int rc = check_geom(geom, who);
  if (rc) return rc;
  if ((rc = check_geom_even(geom, who))) return rc;
  REQUIRE_PTR(dst_d, who);
  REQUIRE_PTR(src_d, who);
  REQUIRE_PTR(gauge_d, who);
  if ((rc = check_dir_sign(dir, sign, who))) return rc;
  if (dst_d == src_d) return set_error(MUGIQ_B200_EINVAL, "%s: dst and src must differ", who);
  const LatGeom g = make_geom(geom->L);
  return displace(dst_d, src_d, gauge_d, dir, sign, g, geom->precision, (cudaStream_t)stream);
}

int mugiq_b200_displace_batch(void *const *dst_d, const void *const *src_d, int nvec, const void *gauge_d, int dir,
                              int sign, const mugiq_b200_geom_t *geom, void *stream) {
  const char *who = "mugiq_b200_displace_batch";
  int rc = check_geom(geom, who);
  if (rc) return rc;
  if ((rc = check_geom_even(geom, who))) return rc;
  REQUIRE_PTR(dst_d, who);
  REQUIRE_PTR(src_d, who);
  REQUIRE_PTR(gauge_d, who);
  if ((rc = check_dir_sign(dir, sign, who))) return rc;
  if (nvec < 0) return set_error(MUGIQ_B200_EINVAL, "%s: nvec = %d", who, nvec);
  for (int i = 0; i < nvec; i++) {
    if (!dst_d[i] || !src_d[i]) return set_error(MUGIQ_B200_EINVAL, "%s: field %d is NULL", who, i);
    if (dst_d[i] == src_d[i]) return set_error(MUGIQ_B200_EINVAL, "%s: dst and src of field %d must differ", who, i);
  }
  const LatGeom g = make_geom(geom->L);
  return displace_batch(dst_d, src_d, nvec, gauge_d, dir, sign, g, geom->precision, (cudaStream_t)stream);
}

static int check_native_order(int order, const char *who) {
  if (order != MUGIQ_B200_ORDER_FLOAT2 && order != MUGIQ_B200_ORDER_FLOAT4)
    return set_error(MUGIQ_B200_EINVAL, "%s: order = %d is not a QUDA native order (FLOAT2 = 2, FLOAT4 = 4)", who, order);
  return MUGIQ_B200_OK;
}

int mugiq_b200_contract_native(void *loop_d, const void *const *vL_d, const void *const *vR_d, const double *sigma_h, int nvec,
                               int order, int accumulate, const mugiq_b200_geom_t *geom, void *stream) {
  const char *who = "mugiq_b200_contract_native";
  int rc = check_geom(geom, who);
  if (rc) return rc;
  if ((rc = check_native_order(order, who))) return rc;
  REQUIRE_PTR(loop_d, who);
  REQUIRE_PTR(vL_d, who);
  REQUIRE_PTR(sigma_h, who);
  if (nvec < 1) return set_error(MUGIQ_B200_EINVAL, "%s: nvec = %d must be positive", who, nvec);
  for (int i = 0; i < nvec; i++)
    if (!vL_d[i] || (vR_d && !vR_d[i])) return set_error(MUGIQ_B200_EINVAL, "%s: field %d is NULL", who, i);
  return contract_batch_native(loop_d, vL_d, vR_d, sigma_h, nvec, order, accumulate, make_geom(geom->L), geom->precision,
                               (cudaStream_t)stream);
}

int mugiq_b200_displace_native(void *const *dst_d, const void *const *src_d, int nvec, const void *gauge_d, int dir, int sign,
                               int order, const mugiq_b200_geom_t *geom, void *stream) {
  const char *who = "mugiq_b200_displace_native";
  int rc = check_geom(geom, who);
  if (rc) return rc;
  if ((rc = check_geom_even(geom, who))) return rc;
  if ((rc = check_native_order(order, who))) return rc;
  REQUIRE_PTR(dst_d, who);
  REQUIRE_PTR(src_d, who);
  REQUIRE_PTR(gauge_d, who);
  if ((rc = check_dir_sign(dir, sign, who))) return rc;
  if (nvec < 1) return set_error(MUGIQ_B200_EINVAL, "%s: nvec = %d must be positive", who, nvec);
  for (int i = 0; i < nvec; i++) {
    if (!dst_d[i] || !src_d[i]) return set_error(MUGIQ_B200_EINVAL, "%s: field %d is NULL", who, i);
    if (dst_d[i] == src_d[i]) return set_error(MUGIQ_B200_EINVAL, "%s: dst and src of field %d must differ", who, i);
  }
  return displace_batch_native(dst_d, src_d, nvec, gauge_d, dir, sign, order, make_geom(geom->L), geom->precision,
                               (cudaStream_t)stream);
}

long long mugiq_b200_loop_workspace_bytes(const mugiq_b200_geom_t *geom, int nvec,
                                          const mugiq_b200_disp_entry_t *entries, int nentries) {
  const char *who = "mugiq_b200_loop_workspace_bytes";
  int rc = check_geom(geom, who);
  if (rc) return rc;
  if ((rc = check_entries(entries, nentries, who))) return rc;
  return loop_workspace_bytes(make_geom(geom->L), geom->precision, nvec, entries, nentries);
}

int mugiq_b200_loop_accumulate(void *dataPos_d, const void *const *evec_d, const double *sigma_h, int nvec,
                               const void *gauge_d, const mugiq_b200_disp_entry_t *entries, int nentries,
                               int accumulate, void *workspace_d, const mugiq_b200_geom_t *geom, void *stream) {
  const char *who = "mugiq_b200_loop_accumulate";
  int rc = check_geom(geom, who);
  if (rc) return rc;
  REQUIRE_PTR(dataPos_d, who);
  REQUIRE_PTR(evec_d, who);
  REQUIRE_PTR(sigma_h, who);
  if (nvec < 1) return set_error(MUGIQ_B200_EINVAL, "%s: nvec = %d must be positive", who, nvec);
  if ((rc = check_entries(entries, nentries, who))) return rc;
  if (nentries > 0 && (rc = check_geom_even(geom, who))) return rc;
  if (nentries > 0) REQUIRE_PTR(gauge_d, who);
  for (int i = 0; i < nvec; i++)
    if (!evec_d[i]) return set_error(MUGIQ_B200_EINVAL, "%s: eigenvector %d is NULL", who, i);
  return loop_accumulate(dataPos_d, evec_d, sigma_h, nvec, gauge_d, entries, nentries, accumulate, workspace_d,
                         make_geom(geom->L), geom->precision, (cudaStream_t)stream);
}

int mugiq_b200_reorder_mapgamma(void *out_d, const void *in_d, int nData, int nLoop, const mugiq_b200_geom_t *geom,
                                void *stream) {
  const char *who = "mugiq_b200_reorder_mapgamma";
  int rc = check_geom(geom, who);
  if (rc) return rc;
  REQUIRE_PTR(out_d, who);
  REQUIRE_PTR(in_d, who);
  // lib/contract_wrappers.cu:138
  if (nLoop < 1 || nData != nLoop * 16)
    return set_error(MUGIQ_B200_EINVAL, "%s: This function assumes that nData = nLoop * NGamma (got %d, %d)", who,
                     nData, nLoop);
  if (out_d == in_d) return set_error(MUGIQ_B200_EINVAL, "%s: out and in must differ", who);
  return reorder_mapgamma(out_d, in_d, nLoop, make_geom(geom->L), geom->precision, (cudaStream_t)stream);
}

int mugiq_b200_phase_matrix(void *phase_d, const int *mom_h, int Nmom, int ftsign, const int localL[4],
                            const int totalL[4], const int commCoord[4], int precision, void *stream) {
  const char *who = "mugiq_b200_phase_matrix";
  REQUIRE_PTR(phase_d, who);
  REQUIRE_PTR(mom_h, who);
  REQUIRE_PTR(localL, who);
  REQUIRE_PTR(totalL, who);
  if (Nmom < 1) return set_error(MUGIQ_B200_EINVAL, "%s: Nmom = %d must be positive", who, Nmom);
  if (ftsign != 1 && ftsign != -1)  // LoopFTSign, include/enum_mugiq.h:30-35
    return set_error(MUGIQ_B200_EINVAL, "%s: FTSign must be +1 or -1 (got %d)", who, ftsign);
  if (precision != MUGIQ_B200_PREC_SINGLE && precision != MUGIQ_B200_PREC_DOUBLE)
    return set_error(MUGIQ_B200_EINVAL, "%s: precision %d not supported", who, precision);
  for (int i = 0; i < 3; i++)
    if (localL[i] < 1 || totalL[i] < localL[i])
      return set_error(MUGIQ_B200_EINVAL, "%s: bad extents in dimension %d (local %d, total %d)", who, i, localL[i],
                       totalL[i]);
  return phase_matrix(phase_d, mom_h, Nmom, ftsign, localL, totalL, commCoord, precision, (cudaStream_t)stream);
}

int mugiq_b200_phase_matrix_eo(void *phase_eo_d, const int *mom_h, int Nmom, int ftsign, const int localL[4],
                               const int totalL[4], const int commCoord[4], int precision, void *stream) {
  const char *who = "mugiq_b200_phase_matrix_eo";
  REQUIRE_PTR(phase_eo_d, who);
  REQUIRE_PTR(mom_h, who);
  REQUIRE_PTR(localL, who);
  REQUIRE_PTR(totalL, who);
  if (Nmom < 1) return set_error(MUGIQ_B200_EINVAL, "%s: Nmom = %d must be positive", who, Nmom);
  if (ftsign != 1 && ftsign != -1) return set_error(MUGIQ_B200_EINVAL, "%s: FTSign must be +1 or -1 (got %d)", who, ftsign);
  if (precision != MUGIQ_B200_PREC_SINGLE && precision != MUGIQ_B200_PREC_DOUBLE)
    return set_error(MUGIQ_B200_EINVAL, "%s: precision %d not supported", who, precision);
  if (localL[0] & 1) return set_error(MUGIQ_B200_EINVAL, "%s: localL[0] = %d must be even", who, localL[0]);
  for (int i = 0; i < 3; i++)
    if (localL[i] < 1 || totalL[i] < localL[i])
      return set_error(MUGIQ_B200_EINVAL, "%s: bad extents in dimension %d (local %d, total %d)", who, i, localL[i], totalL[i]);
  return phase_matrix_eo(phase_eo_d, mom_h, Nmom, ftsign, localL, totalL, commCoord, precision, (cudaStream_t)stream);
}

long long mugiq_b200_momproj_pos_workspace_bytes(const mugiq_b200_geom_t *geom, int nLoop, int Nmom) {
  const char *who = "mugiq_b200_momproj_pos_workspace_bytes";
  int rc = check_geom(geom, who);
  if (rc) return rc;
  if (nLoop < 1 || Nmom < 1) return set_error(MUGIQ_B200_EINVAL, "%s: nLoop = %d, Nmom = %d must be positive", who, nLoop, Nmom);
  return momproj_pos_workspace_bytes(make_geom(geom->L), nLoop, Nmom, geom->precision);
}

int mugiq_b200_momproj_pos(void *mom_d, const void *dataPos_d, const void *phase_eo_d, int nLoop, int Nmom,
                           const mugiq_b200_geom_t *geom, void *workspace_d, void *stream) {
  const char *who = "mugiq_b200_momproj_pos";
  int rc = check_geom(geom, who);
  if (rc) return rc;
  REQUIRE_PTR(mom_d, who);
  REQUIRE_PTR(dataPos_d, who);
  REQUIRE_PTR(phase_eo_d, who);
  REQUIRE_PTR(workspace_d, who);
  if (nLoop < 1 || Nmom < 1) return set_error(MUGIQ_B200_EINVAL, "%s: nLoop = %d, Nmom = %d must be positive", who, nLoop, Nmom);
  return momproj_pos(mom_d, dataPos_d, phase_eo_d, nLoop, Nmom, make_geom(geom->L), geom->precision, workspace_d,
                     (cudaStream_t)stream);
}

long long mugiq_b200_momproj_workspace_bytes(long long M, int N, long long K, int precision) {
  if (M < 1 || N < 1 || K < 1)
    return set_error(MUGIQ_B200_EINVAL, "mugiq_b200_momproj_workspace_bytes: bad shape %lld x %d x %lld", M, N, K);
  if (precision != MUGIQ_B200_PREC_SINGLE && precision != MUGIQ_B200_PREC_DOUBLE)
    return set_error(MUGIQ_B200_EINVAL, "mugiq_b200_momproj_workspace_bytes: precision %d not supported", precision);
  return momproj_workspace_bytes(M, N, K, precision);
}

int mugiq_b200_momproj(void *mom_d, const void *posMP_d, const void *phase_d, long long M, int N, long long K,
                       int precision, void *workspace_d, void *stream) {
  const char *who = "mugiq_b200_momproj";
  REQUIRE_PTR(mom_d, who);
  REQUIRE_PTR(posMP_d, who);
  REQUIRE_PTR(phase_d, who);
  if (M < 1 || N < 1 || K < 1) return set_error(MUGIQ_B200_EINVAL, "%s: bad shape %lld x %d x %lld", who, M, N, K);
  if (precision != MUGIQ_B200_PREC_SINGLE && precision != MUGIQ_B200_PREC_DOUBLE)
    return set_error(MUGIQ_B200_EINVAL, "%s: precision %d not supported", who, precision);
  return momproj(mom_d, posMP_d, phase_d, M, N, K, precision, workspace_d, (cudaStream_t)stream);
}

}  // extern "C"
