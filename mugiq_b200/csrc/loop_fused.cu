// loop_fused.cu — schedule of the eigenvector / displacement loop nest of Loop_Mugiq::computeCoarseLoop
// (/root/reference/lib/loop_mugiq.cpp:455-509) on top of the fused kernel (fused_kernel.cu).
//
// The reference runs, per displacement entry and per eigenvector, `stop` one-link hops (each: zero + kernel +
// two field copies, lib/displace.cpp:47-67) and a contraction per requested length, re-traversing all
// eigenvectors once per entry.  Here a LoopPlan is built once per (gauge field, entry list):
//   1. the set of distinct (direction, sign, length) loops the entries request is collected;
//   2. Wilson lines W_k are built from the links (wilson.cu), so each loop is one 3x3 multiply per site;
//   3. a minus-direction loop whose plus-direction partner is also requested is not computed at all: it is
//      derived from the partner's finished loop buffer (loop_minus_from_plus, exact identity);
//      repeated requests of the same loop are copied;
//   4. the remaining loops + the ultra-local loop are packed into launch groups of the fused kernel, which
//      reads every eigenvector once per group and writes the loop buffer once per batch.
// LoopPlan::accumulate() may be called for any number of eigenvector batches (shards, host-streamed batches);
// LoopPlan::finalize() then fills the derived slots.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "fused.cuh"
#include "plan.cuh"

namespace mugiq_b200 {

static bool env_flag(const char *name) {
  const char *e = getenv(name);
  return e && e[0] && e[0] != '0';
}

// Collects requested loops, decides what is computed / derived, sizes the Wilson-line storage.
static const LoopPlan::WField *find_w(const LoopPlan &pl, int dir, int sign, int len) {
  for (const LoopPlan::WField &w : pl.wfields)
    if (w.dir == dir && w.sign == sign && w.len == len) return &w;
  return nullptr;
}

// derive_ok: slots may be filled after the eigenvector sum (minus from plus, copies of repeated requests); false
// when the call accumulates onto partial sums it did not produce itself, then every requested loop is computed.
static void plan_layout(LoopPlan &pl, const LatGeom &g, int precision, const mugiq_b200_disp_entry_t *entries, int nentries,
                        bool derive_ok) {
  pl.g = g;
  pl.precision = precision;
  pl.symmetric = derive_ok && !env_flag("MUGIQ_B200_NO_PM_SYMMETRY");
  pl.comps.clear();
  pl.derives.clear();
  pl.zero_slots.clear();
  pl.wfields.clear();
  struct Req {
    int dir, sign, len, iL;
  };
  std::vector<Req> reqs;
  int iL = 1;
  for (int e = 0; e < nentries; e++) {
    const mugiq_b200_disp_entry_t &en = entries[e];
    const int nL = en.stop - en.start + 1;  // nLoopPerEntry, include/loop_mugiq.h:241
    const int first = en.start < 1 ? 1 : en.start;
    for (int s = 0; s < nL; s++) {
      const int len = first + s;  // slot s holds the s-th contraction of the hop loop (dispCount, lib/loop_mugiq.cpp:492-495)
      if (len <= en.stop)
        reqs.push_back({en.dir, en.sign, len, iL + s});
      else
        pl.zero_slots.push_back(iL + s);
    }
    iL += nL;
  }
  pl.nLoop = iL;

  auto find_comp = [&](int dir, int sign, int len) -> int {
    for (size_t i = 0; i < pl.comps.size(); i++)
      if (pl.comps[i].dir == dir && pl.comps[i].sign == sign && pl.comps[i].len == len) return (int)i;
    return -1;
  };
  pl.comps.push_back({-1, 0, 0, 0, nullptr});  // ultra-local loop, slot 0
  // pass 1: plus loops (and, without the symmetry, minus loops) are computed
  for (const Req &r : reqs) {
    if (r.sign == MUGIQ_B200_SIGN_MINUS && pl.symmetric) continue;
    const int c = derive_ok ? find_comp(r.dir, r.sign, r.len) : -1;
    if (c < 0)
      pl.comps.push_back({r.dir, r.sign, r.len, r.iL, nullptr});
    else
      pl.derives.push_back({0, r.iL, pl.comps[c].iL, 0, 0});
  }
  // pass 2 (symmetric): a minus loop is derived from its plus partner if that is computed, else computed itself
  if (pl.symmetric) {
    for (const Req &r : reqs) {
      if (r.sign != MUGIQ_B200_SIGN_MINUS) continue;
      const int cp = find_comp(r.dir, MUGIQ_B200_SIGN_PLUS, r.len);
      if (cp >= 0) {
        pl.derives.push_back({1, r.iL, pl.comps[cp].iL, r.dir, r.len});
        continue;
      }
      const int c = find_comp(r.dir, r.sign, r.len);
      if (c < 0)
        pl.comps.push_back({r.dir, r.sign, r.len, r.iL, nullptr});
      else
        pl.derives.push_back({0, r.iL, pl.comps[c].iL, 0, 0});
    }
  }
  // Wilson lines: per direction the plus chain 2..kmax (k = 1 is the gauge field itself) and one field per
  // computed minus loop
  pl.nW = 0;
  for (int dir = 0; dir < 4; dir++) {
    int kmax = 0;
    for (const LoopPlan::Comp &c : pl.comps)
      if (c.dir == dir) kmax = std::max(kmax, c.len);
    for (int k = 2; k <= kmax; k++) pl.wfields.push_back({dir, MUGIQ_B200_SIGN_PLUS, k, pl.nW++});
    for (const LoopPlan::Comp &c : pl.comps)
      if (c.dir == dir && c.sign == MUGIQ_B200_SIGN_MINUS && !find_w(pl, dir, MUGIQ_B200_SIGN_MINUS, c.len))
        pl.wfields.push_back({dir, MUGIQ_B200_SIGN_MINUS, c.len, pl.nW++});
  }
}

// Launch groups (host only): the ultra-local loop rides in the first group; the displaced loops are sorted by
// (length, direction) so that a group mixes directions.  What a group stages per eigenvector is the union of its loops'
// shifted images of the tile: an x loop adds a few sites, a y loop a row or two, a z or t loop a whole copy of the tile.
// Groups of one direction each (x:1..4 | y:1..4 | z:1..4 | t:1..4, the round-1 order) move the same total but leave the
// z and t groups with 5x the tile per stage - bound by the L2 -> shared-memory path instead of the FP64 pipe - while
// mixed groups all sit near 3.5x.  MUGIQ_B200_GROUP_BY_DIR=1 restores the per-direction order (for measurements).
static int plan_make_groups(LoopPlan &pl) {
  const int maxl = fused_max_loops_per_group(pl.g, pl.precision);
  if (maxl < 0 || (maxl < 1 && pl.comps.size() > 1))
    return set_error(MUGIQ_B200_EINVAL, "loop plan: lattice %dx%dx%dx%d does not fit the fused kernel's tile", pl.g.L[0],
                     pl.g.L[1], pl.g.L[2], pl.g.L[3]);
  std::vector<LoopPlan::Comp> order(pl.comps.begin(), pl.comps.end());
  const bool by_dir = env_flag("MUGIQ_B200_GROUP_BY_DIR");
  std::stable_sort(order.begin() + 1, order.end(), [by_dir](const LoopPlan::Comp &a, const LoopPlan::Comp &b) {
    if (by_dir) {
      if (a.dir != b.dir) return a.dir < b.dir;
      if (a.sign != b.sign) return a.sign > b.sign;
      return a.len < b.len;
    }
    if (a.len != b.len) return a.len < b.len;
    if (a.dir != b.dir) return a.dir < b.dir;
    return a.sign > b.sign;
  });
  pl.groups.clear();
  const size_t per_loop = (size_t)16 * pl.g.volume;
  size_t i = 1;  // order[0] is the ultra-local loop
  do {
    FusedGroup grp;
    grp.nloops = maxl > 0 ? (int)std::min<size_t>(maxl, order.size() - i) : 0;
    for (int j = 0; j < grp.nloops; j++) {
      const LoopPlan::Comp &c = order[i + j];
      grp.loop[j].W = c.W;
      grp.loop[j].out_off = (long long)(per_loop * c.iL);
      grp.loop[j].dir = c.dir;
      grp.loop[j].sign = c.sign == MUGIQ_B200_SIGN_PLUS ? +1 : -1;
      grp.loop[j].len = c.len;
      grp.loop[j].pad_ = 0;
    }
    pl.groups.push_back(grp);
    i += grp.nloops;
  } while (i < order.size());
  return MUGIQ_B200_OK;
}

// Builds the Wilson lines into pl.wbuf and the launch groups.
static int plan_build(LoopPlan &pl, const void *gauge_d, cudaStream_t stream) {
  const size_t lfb = pl.link_field_bytes();
  auto plus_ptr = [&](int dir, int len) -> const void * {
    if (len == 1) return static_cast<const char *>(gauge_d) + (size_t)dir * lfb;
    const LoopPlan::WField *w = find_w(pl, dir, MUGIQ_B200_SIGN_PLUS, len);
    return w ? pl.wptr(w->index) : nullptr;
  };
  for (const LoopPlan::WField &w : pl.wfields) {  // ordered: plus chain ascending, then minus fields, per direction
    void *dst = static_cast<char *>(pl.wbuf) + w.index * lfb;
    int rc;
    if (w.sign == MUGIQ_B200_SIGN_PLUS)  // W+_k(x) = W+_{k-1}(x) U(x + (k-1) mu)
      rc = wilson_extend(dst, plus_ptr(w.dir, w.len - 1), gauge_d, w.dir, w.len - 1, pl.g, pl.precision, stream);
    else  // W-_k(x) = [W+_k(x - k mu)]^dag
      rc = wilson_minus_from_plus(dst, plus_ptr(w.dir, w.len), w.dir, w.len, pl.g, pl.precision, stream);
    if (rc) return rc;
  }
  for (LoopPlan::Comp &c : pl.comps) {
    if (c.dir < 0) continue;
    if (c.sign == MUGIQ_B200_SIGN_PLUS)
      c.W = plus_ptr(c.dir, c.len);
    else
      c.W = pl.wptr(find_w(pl, c.dir, MUGIQ_B200_SIGN_MINUS, c.len)->index);
  }
  return plan_make_groups(pl);
}

// Contribution of the given eigenvectors to every loop the plan computes, on the time-slices [t0, t1) only.
int plan_accumulate_range(const LoopPlan &pl, void *dataPos_d, const void *const *evec_d, const double *sigma_h, int nvec,
                          int accumulate, int t0, int t1, bool zero_unreached, cudaStream_t stream, int evec_order) {
  if (evec_order < 0) evec_order = pl.evec_order;
  char *pos = static_cast<char *>(dataPos_d);
  if (!accumulate && zero_unreached)
    for (int z : pl.zero_slots) MUGIQ_CUDA_CHECK(cudaMemsetAsync(pos + (size_t)z * pl.loop_bytes(), 0, pl.loop_bytes(), stream));
  for (int done = 0; done < nvec; done += kFusedMaxVec) {
    FusedVecTable vt;
    vt.nvec = std::min(kFusedMaxVec, nvec - done);
    vt.native = evec_order == MUGIQ_B200_ORDER_SITE ? 0 : evec_order;
    for (int i = 0; i < vt.nvec; i++) {
      vt.evec[i] = evec_d[done + i];
      vt.inv_sigma[i] = inv_sigma_of(sigma_h[done + i], pl.precision);
    }
    if (vt.native) {  // QUDA FLOAT2 fields are staged through their tensor maps (encoded once per field, cached)
      int rc = fused_native_tmaps(vt.evec, evec_d + done, vt.nvec, pl.g, pl.precision, stream);
      if (rc) return rc;
    }
    for (size_t gi = 0; gi < pl.groups.size(); gi++) {
      // slot 0 (ultra-local) is accumulated by the first group
      int rc = fused_group_launch(dataPos_d, pl.groups[gi], gi == 0 ? 0 : -1, vt, accumulate || done > 0, pl.g,
                                  pl.precision, stream, t0, t1);
      if (rc) return rc;
    }
  }
  return MUGIQ_B200_OK;
}

static int plan_accumulate(const LoopPlan &pl, void *dataPos_d, const void *const *evec_d, const double *sigma_h, int nvec,
                           int accumulate, cudaStream_t stream) {
  return plan_accumulate_range(pl, dataPos_d, evec_d, sigma_h, nvec, accumulate, pl.t_begin, pl.t_end, true, stream);
}

static int plan_finalize(const LoopPlan &pl, void *dataPos_d, int accumulate, cudaStream_t stream) {
  char *pos = static_cast<char *>(dataPos_d);
  const size_t lb = pl.loop_bytes();
  MinusBatch mb;
  mb.n = 0;
  for (const LoopPlan::Derive &d : pl.derives) {
    if (d.kind != 1) continue;
    mb.item[mb.n++] = {d.dst, d.src, d.dir, d.len};
    if (mb.n == kMinusBatch) {
      int rc = loop_minus_from_plus(dataPos_d, mb, accumulate, pl.g, pl.precision, stream);
      if (rc) return rc;
      mb.n = 0;
    }
  }
  if (int rc = loop_minus_from_plus(dataPos_d, mb, accumulate, pl.g, pl.precision, stream)) return rc;
  // copies last: a copy source may itself be a derived slot only through `repeated` requests of computed loops,
  // which are resolved against computed slots above, so order does not matter beyond kind
  for (const LoopPlan::Derive &d : pl.derives) {
    if (d.kind == 0) {
      if (accumulate)
        return set_error(MUGIQ_B200_EINVAL, "loop plan: repeated entries cannot be finalized in accumulate mode");
      MUGIQ_CUDA_CHECK(cudaMemcpyAsync(pos + (size_t)d.dst * lb, pos + (size_t)d.src * lb, lb, cudaMemcpyDeviceToDevice, stream));
    }
  }
  return MUGIQ_B200_OK;
}

// ---- one-shot entry points (kernels.cuh) ---------------------------------------------------------------------
long long loop_workspace_bytes(const LatGeom &g, int precision, int nvec, const mugiq_b200_disp_entry_t *entries,
                               int nentries) {
  (void)nvec;
  if (nentries <= 0) return 0;
  // the non-symmetric layout needs the most Wilson lines; the one-shot call may use either
  LoopPlan a, b;
  plan_layout(a, g, precision, entries, nentries, true);
  plan_layout(b, g, precision, entries, nentries, false);
  return (long long)(std::max(a.nW, b.nW) * a.link_field_bytes());
}

int loop_accumulate(void *dataPos_d, const void *const *evec_d, const double *sigma_h, int nvec, const void *gauge_d,
                    const mugiq_b200_disp_entry_t *entries, int nentries, int accumulate, void *workspace_d,
                    const LatGeom &g, int precision, cudaStream_t stream) {
  LoopPlan pl;
  // accumulating onto partial sums: the minus slots cannot be derived from accumulated plus slots, compute them
  plan_layout(pl, g, precision, entries, nentries, accumulate == 0);
  if (pl.nW > 0 && !workspace_d) return set_error(MUGIQ_B200_EINVAL, "loop_accumulate: workspace_d is NULL");
  pl.wbuf = workspace_d;
  int rc = plan_build(pl, gauge_d, stream);
  if (rc) return rc;
  if ((rc = plan_accumulate(pl, dataPos_d, evec_d, sigma_h, nvec, accumulate, stream))) return rc;
  return plan_finalize(pl, dataPos_d, 0, stream);
}

}  // namespace mugiq_b200

// ---- plan C-ABI (include/mugiq_b200.h) ---------------------------------------------------------------------------
using namespace mugiq_b200;

struct mugiq_b200_loop_plan_s {
  LoopPlan pl;
};

namespace mugiq_b200 {
const LoopPlan &plan_of(const mugiq_b200_loop_plan_t *plan) { return plan->pl; }
}  // namespace mugiq_b200

extern "C" {

int mugiq_b200_loop_plan_create(mugiq_b200_loop_plan_t **plan, const void *gauge_d, const mugiq_b200_disp_entry_t *entries,
                                int nentries, const mugiq_b200_geom_t *geom, void *stream) {
  const char *who = "mugiq_b200_loop_plan_create";
  if (!plan) return set_error(MUGIQ_B200_EINVAL, "%s: plan is NULL", who);
  *plan = nullptr;
  int rc = check_geom(geom, who);
  if (rc) return rc;
  if ((rc = check_entries(entries, nentries, who))) return rc;
  if (nentries > 0 && (rc = check_geom_even(geom, who))) return rc;
  if (nentries > 0 && !gauge_d) return set_error(MUGIQ_B200_EINVAL, "%s: gauge_d is NULL", who);
  mugiq_b200_loop_plan_s *p = new mugiq_b200_loop_plan_s;
  plan_layout(p->pl, make_geom(geom->L), geom->precision, entries, nentries, true);
  if (p->pl.nW > 0) {
    if (cudaMalloc(&p->pl.wbuf, p->pl.nW * p->pl.link_field_bytes()) != cudaSuccess) {
      cudaGetLastError();
      const size_t want = p->pl.nW * p->pl.link_field_bytes();
      delete p;
      return set_error(MUGIQ_B200_ENOMEM, "%s: cannot allocate %zu bytes of Wilson-line storage", who, want);
    }
    p->pl.own_wbuf = true;
  }
  rc = plan_build(p->pl, gauge_d, (cudaStream_t)stream);
  if (rc) {
    if (p->pl.own_wbuf) cudaFree(p->pl.wbuf);
    delete p;
    return rc;
  }
  *plan = p;
  return MUGIQ_B200_OK;
}

int mugiq_b200_fused_tiling_check(const mugiq_b200_disp_entry_t *entries, int nentries, const mugiq_b200_geom_t *geom, int t_begin,
                                  int t_end, int group, int evec_order, long long out[8]) {
  const char *who = "mugiq_b200_fused_tiling_check";
  int rc = check_geom(geom, who);
  if (rc) return rc;
  if ((rc = check_entries(entries, nentries, who))) return rc;
  if (nentries > 0 && (rc = check_geom_even(geom, who))) return rc;
  if (!out) return set_error(MUGIQ_B200_EINVAL, "%s: out is NULL", who);
  LoopPlan pl;
  plan_layout(pl, make_geom(geom->L), geom->precision, entries, nentries, true);
  if ((rc = plan_make_groups(pl))) return rc;
  if (group < 0) return (int)pl.groups.size();
  if (group >= (int)pl.groups.size()) return set_error(MUGIQ_B200_EINVAL, "%s: the plan has %zu groups", who, pl.groups.size());
  if (t_end < 0) t_end = geom->L[3];
  if (t_begin < 0 || t_begin >= t_end || t_end > geom->L[3]) return set_error(MUGIQ_B200_EINVAL, "%s: bad time-slice range", who);
  // the ultra-local loop rides in the first group (plan_accumulate_range)
  return fused_tiling_check(pl.groups[group], pl.g, pl.precision, t_begin, t_end, evec_order == MUGIQ_B200_ORDER_FLOAT2, group == 0, out);
}

int mugiq_b200_loop_plan_destroy(mugiq_b200_loop_plan_t *plan) {
  if (!plan) return MUGIQ_B200_OK;
  if (plan->pl.own_wbuf && plan->pl.wbuf) cudaFree(plan->pl.wbuf);
  delete plan;
  return MUGIQ_B200_OK;
}

int mugiq_b200_loop_plan_nloop(const mugiq_b200_loop_plan_t *plan) {
  if (!plan) return set_error(MUGIQ_B200_EINVAL, "mugiq_b200_loop_plan_nloop: plan is NULL");
  return plan->pl.nLoop;
}

int mugiq_b200_loop_plan_info(const mugiq_b200_loop_plan_t *plan, int *ncomputed, int *nderived, int *ngroups,
                              long long *wilson_bytes) {
  if (!plan) return set_error(MUGIQ_B200_EINVAL, "mugiq_b200_loop_plan_info: plan is NULL");
  if (ncomputed) *ncomputed = (int)plan->pl.comps.size();
  if (nderived) *nderived = (int)plan->pl.derives.size();
  if (ngroups) *ngroups = (int)plan->pl.groups.size();
  if (wilson_bytes) *wilson_bytes = (long long)(plan->pl.nW * plan->pl.link_field_bytes());
  return MUGIQ_B200_OK;
}

int mugiq_b200_loop_plan_computed_slots(const mugiq_b200_loop_plan_t *plan, int *slots, int max_slots) {
  if (!plan) return set_error(MUGIQ_B200_EINVAL, "mugiq_b200_loop_plan_computed_slots: plan is NULL");
  const int n = (int)plan->pl.comps.size();
  for (int i = 0; slots && i < n && i < max_slots; i++) slots[i] = plan->pl.comps[i].iL;
  return n;
}

int mugiq_b200_loop_plan_set_evec_order(mugiq_b200_loop_plan_t *plan, int order) {
  const char *who = "mugiq_b200_loop_plan_set_evec_order";
  if (!plan) return set_error(MUGIQ_B200_EINVAL, "%s: plan is NULL", who);
  if (order == MUGIQ_B200_ORDER_FLOAT4)
    return set_error(MUGIQ_B200_EINVAL, "%s: FLOAT4 fields are not staged directly (convert them with mugiq_b200_ingest_spinor_batch, "
                     "or feed them through mugiq_b200_loop_feed_*)", who);
  if (order != MUGIQ_B200_ORDER_SITE && order != MUGIQ_B200_ORDER_FLOAT2)
    return set_error(MUGIQ_B200_EINVAL, "%s: unknown field order %d", who, order);
  if (order == MUGIQ_B200_ORDER_FLOAT2 && plan->pl.g.volumeCB % 8)
    return set_error(MUGIQ_B200_EINVAL, "%s: FLOAT2 staging needs volumeCB = %d to be a multiple of 8", who, plan->pl.g.volumeCB);
  plan->pl.evec_order = order;
  return MUGIQ_B200_OK;
}

int mugiq_b200_loop_plan_set_t_range(mugiq_b200_loop_plan_t *plan, int t_begin, int t_end) {
  const char *who = "mugiq_b200_loop_plan_set_t_range";
  if (!plan) return set_error(MUGIQ_B200_EINVAL, "%s: plan is NULL", who);
  if (t_begin < 0 || t_begin >= t_end || t_end > plan->pl.g.L[3])
    return set_error(MUGIQ_B200_EINVAL, "%s: bad range [%d, %d) on Lt = %d", who, t_begin, t_end, plan->pl.g.L[3]);
  plan->pl.t_begin = t_begin;
  plan->pl.t_end = t_end;
  return MUGIQ_B200_OK;
}

int mugiq_b200_loop_plan_t_halo(const mugiq_b200_loop_plan_t *plan, int *evec_lower, int *evec_upper, int *loop_lower) {
  if (!plan) return set_error(MUGIQ_B200_EINVAL, "mugiq_b200_loop_plan_t_halo: plan is NULL");
  int lo = 0, up = 0, ll = 0;
  for (const LoopPlan::Comp &c : plan->pl.comps) {
    if (c.dir != MUGIQ_B200_DIR_T) continue;
    if (c.sign == MUGIQ_B200_SIGN_PLUS) up = std::max(up, c.len);
    if (c.sign == MUGIQ_B200_SIGN_MINUS) lo = std::max(lo, c.len);
  }
  for (const LoopPlan::Derive &d : plan->pl.derives)
    if (d.kind == 1 && d.dir == MUGIQ_B200_DIR_T) ll = std::max(ll, d.len);
  if (evec_lower) *evec_lower = lo;
  if (evec_upper) *evec_upper = up;
  if (loop_lower) *loop_lower = ll;
  return MUGIQ_B200_OK;
}

int mugiq_b200_loop_plan_accumulate(const mugiq_b200_loop_plan_t *plan, void *dataPos_d, const void *const *evec_d,
                                    const double *sigma_h, int nvec, int accumulate, void *stream) {
  const char *who = "mugiq_b200_loop_plan_accumulate";
  if (!plan) return set_error(MUGIQ_B200_EINVAL, "%s: plan is NULL", who);
  if (!dataPos_d || !evec_d || !sigma_h) return set_error(MUGIQ_B200_EINVAL, "%s: NULL argument", who);
  if (nvec < 1) return set_error(MUGIQ_B200_EINVAL, "%s: nvec = %d must be positive", who, nvec);
  for (int i = 0; i < nvec; i++)
    if (!evec_d[i]) return set_error(MUGIQ_B200_EINVAL, "%s: eigenvector %d is NULL", who, i);
  return plan_accumulate(plan->pl, dataPos_d, evec_d, sigma_h, nvec, accumulate, (cudaStream_t)stream);
}

int mugiq_b200_loop_plan_finalize(const mugiq_b200_loop_plan_t *plan, void *dataPos_d, int accumulate, void *stream) {
  const char *who = "mugiq_b200_loop_plan_finalize";
  if (!plan) return set_error(MUGIQ_B200_EINVAL, "%s: plan is NULL", who);
  if (!dataPos_d) return set_error(MUGIQ_B200_EINVAL, "%s: dataPos_d is NULL", who);
  return plan_finalize(plan->pl, dataPos_d, accumulate, (cudaStream_t)stream);
}

}  // extern "C"
