// plan.cuh — the loop plan (loop_fused.cu) as seen by the other translation units of the library (comm.cu).
#pragma once
#include <vector>

#include "fused.cuh"

namespace mugiq_b200 {

struct LoopPlan {
  struct Comp {  // a loop computed by the fused kernel
    int dir, sign, len;
    int iL;         // slot of dataPos it is written to
    const void *W;  // Wilson line (device), nullptr for the ultra-local loop
  };
  struct Derive {  // a slot filled after the eigenvector sum
    int kind;      // 0 = copy of slot src, 1 = minus from plus
    int dst, src, dir, len;
  };
  LatGeom g;
  int precision;
  int nLoop;
  bool symmetric;
  int evec_order = MUGIQ_B200_ORDER_SITE;  // layout of the eigenvectors accumulate() is given (set_evec_order)
  int t_begin = 0, t_end = -1;  // time-slices the fused kernels compute (-1: all): interior of a lattice-T split slab
  std::vector<Comp> comps;
  std::vector<Derive> derives;
  std::vector<int> zero_slots;  // slots no hop reaches (start < 1): stay zero, as in the reference
  std::vector<FusedGroup> groups;  // displaced loops; the ultra-local loop rides in groups[0]
  // Wilson-line storage
  struct WField {
    int dir, sign, len;
    size_t index;  // field index inside wbuf
  };
  std::vector<WField> wfields;
  size_t nW = 0;
  void *wbuf = nullptr;
  bool own_wbuf = false;

  size_t link_field_bytes() const { return (size_t)g.volume * kLinkLen * 2 * prec_bytes(precision); }
  size_t loop_bytes() const { return (size_t)16 * g.volume * 2 * prec_bytes(precision); }
  const void *wptr(size_t index) const { return static_cast<const char *>(wbuf) + index * link_field_bytes(); }
};

// the plan behind the opaque C-ABI handle
const LoopPlan &plan_of(const mugiq_b200_loop_plan_t *plan);
// Contribution of the given eigenvectors to every loop the plan computes, on the time-slices [t0, t1) only (t1 < 0: up
// to Lt).  zero_unreached: also clear the slots no hop reaches (once per call, not once per chunk).
// evec_order: layout of the fields (MUGIQ_B200_ORDER_SITE or _FLOAT2; < 0: the plan's own setting).
int plan_accumulate_range(const LoopPlan &pl, void *dataPos_d, const void *const *evec_d, const double *sigma_h, int nvec,
                          int accumulate, int t0, int t1, bool zero_unreached, cudaStream_t stream, int evec_order = -1);

}  // namespace mugiq_b200
