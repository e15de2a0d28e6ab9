// fused_kernel.cu — stages 1+2 fused: for one tile of lattice sites and one group of loops, stream every
// eigenvector of the batch through shared memory once and accumulate, in registers,
//     M_l(x)[be][al] += (1/sigma_n) sum_c conj(v_n(x)[be,c]) * [W_l(x) v_n(x + d_l)][al,c]
// for every displaced loop l of the group, plus the ultra-local matrix (W = 1, d = 0), then project on the 16
// gammas and write the loop buffer once.
//
// Replaces, for the whole eigenvector x displacement loop nest of Loop_Mugiq::computeCoarseLoop
// (/root/reference/lib/loop_mugiq.cpp:455-509): performCovariantDisplacementVector + the kernel
// lib/mugiq_displace_kernels.cu:156-185, Displace::doVectorDisplacement's zero/copy/copy
// (lib/displace.cpp:47-67), performLoopContraction + loopContract_kernel (lib/mugiq_contract_kernels.cu:45-122).
// A displacement of length k is applied as one Wilson-line multiplication W_k(x) v(x + k mu) (wilson.cu
// builds W_k from the links once per gauge field), so no displaced eigenvector is ever written to memory.
//
// Design ("kernels"; the numbers quoted are measured on B200, profiles/):
//  * The kernel is FP64-FMA bound (B200: 34 TFLOP/s measured, DMMA shares the pipe): 360 DFMA/DMUL per
//    (eigvec, site, displaced loop) + ~30 for the share of the ultra-local matrix, against 192 B of compulsory
//    HBM traffic per (eigvec, site).  Everything else is arranged so that the FP64 pipe is the only busy unit:
//    on this chip every other instruction takes FP64 issue slots (one IMAD per DFMA halves the DFMA rate,
//    tools/microbench.cu), so the loop body is ~390 FP64 + ~175 other instructions per eigenvector.
//  * CTA tile = NR (<= 4) lattice rows (all x at fixed y,z,t), preferably consecutive in y: in the even/odd
//    site-major layout a row is two contiguous half-rows (one per parity) of Lx/2 sites x 192 B and consecutive-y
//    rows are contiguous, so the tile and its shifted copies are fetched with a handful of multi-KB bulk-TMA
//    copies (cp.async.bulk + mbarrier complete_tx; small copies cost ~100-500 cycles each, tools/tma_bench.cu).
//    Per eigenvector a stage holds the tile's own rows plus the rows shifted by every displacement of the group,
//    de-duplicated; x-displacements stay inside the row (periodic wrap).  Layout in shared memory is dense:
//    [parity][slot][Lx/2 sites][12 complex].
//  * 8 warps (2 per SM sub-partition -> 255 registers per thread, no spills): warp = (loop, parity-half of the
//    tile), thread = (site, loop) and owns the 4x4 complex spin matrix M (32 doubles), the 3x3 link W (18) and its
//    share of the Hermitian ultra-local matrix for the whole batch.  Every warp issues its share of the TMA copies
//    S-2 stages ahead (full/empty mbarrier ring), so there is no dedicated producer warp and no block barrier in
//    the eigenvector loop.
//  * Bank conflicts: 192-B site stride means lanes reading the same component hit only two 16-B bank groups.
//    Instead of padding (which would forbid multi-row bulk copies) each lane reads its site with the spin index
//    rotated by k = (site_index/2) mod 4, i.e. it keeps spin (b+k) mod 4 in register slot b.  The 8 lanes of a
//    quarter-warp then touch 8 distinct bank groups (conflict-free 128-bit reads); the rotation only relabels
//    M[be][al] and is undone once in the epilogue.
#include <cstdlib>

#include "fused.cuh"

namespace mugiq_b200 {

// ---- PTX helpers: mbarrier + bulk TMA -------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

// ---- staged-row ("slot") table of a tile: shared by host (sizing) and device -------------------------------
constexpr int kMaxCopies = 2 * kFusedMaxSlots;
struct SlotTable {
  int row[kFusedMaxSlots];                 // lexicographic row index y + Ly*(z + Lz*t) of each slot
  int nbr[kFusedMaxLoops][kFusedMaxRows];  // slot holding the shifted row of own row i for loop j
  int nslots;
  // bulk copies of one stage: runs of slots whose rows are consecutive in memory, per parity
  int ncopies;
  int cp_soff[kMaxCopies];    // byte offset inside the stage
  int cp_goff16[kMaxCopies];  // offset inside the eigenvector, in units of 16 B
  int cp_bytes[kMaxCopies];
};

__host__ __device__ inline int wrap(int a, int n) {
  a %= n;
  return a < 0 ? a + n : a;
}

__host__ __device__ inline void build_slots(SlotTable &st, const FusedGroup &grp, const FusedTiling &tl, const LatGeom &g,
                                            int site_bytes, int y0, int z0, int t0) {
  int nslots = tl.NR;
  for (int i = 0; i < tl.NR; i++) {
    const int a = i % tl.TY, b = (i / tl.TY) % tl.TZ, c = i / (tl.TY * tl.TZ);
    st.row[i] = (y0 + a) + g.L[1] * ((z0 + b) + g.L[2] * (t0 + c));
  }
  for (int j = 0; j < grp.nloops; j++) {
    const FusedLoop &lp = grp.loop[j];
    for (int i = 0; i < tl.NR; i++) {
      if (lp.dir == 0) {  // x-displacement: the neighbour lives in the same row
        st.nbr[j][i] = i;
        continue;
      }
      const int a = i % tl.TY, b = (i / tl.TY) % tl.TZ, c = i / (tl.TY * tl.TZ);
      int y = y0 + a, z = z0 + b, t = t0 + c;
      const int sh = lp.sign * lp.len;
      if (lp.dir == 1) y = wrap(y + sh, g.L[1]);
      if (lp.dir == 2) z = wrap(z + sh, g.L[2]);
      if (lp.dir == 3) t = wrap(t + sh, g.L[3]);
      const int r = y + g.L[1] * (z + g.L[2] * t);
      int found = -1;
      for (int k = 0; k < nslots; k++)
        if (st.row[k] == r) found = k;
      if (found < 0) {
        found = nslots++;
        st.row[found] = r;
      }
      st.nbr[j][i] = found;
    }
  }
  st.nslots = nslots;
  // merge slots with consecutive rows into one copy per parity
  const int hrb = g.Lh * site_bytes;
  int nc = 0;
  for (int k = 0; k < nslots;) {
    int len = 1;
    while (k + len < nslots && st.row[k + len] == st.row[k] + len) len++;
    for (int p = 0; p < 2; p++) {
      st.cp_soff[nc] = (p * nslots + k) * hrb;
      st.cp_goff16[nc] = (int)((((long long)p * g.volumeCB + (long long)st.row[k] * g.Lh) * site_bytes) >> 4);
      st.cp_bytes[nc] = len * hrb;
      nc++;
    }
    k += len;
  }
  st.ncopies = nc;
}

template <typename F> struct FusedArgs {
  LatGeom g;
  FusedTiling tl;
  FusedGroup grp;      // displaced loops only (0..kFusedMaxLoops)
  FusedVecTable vt;
  F *dataPos;
  long long ul_off;    // complex offset of the ultra-local loop's block in dataPos, < 0: not in this launch
  int accumulate;
  int tt0;             // first tile in t of this launch (time-slice range of a lattice-T split; 0 = whole lattice)
};

constexpr int kSmemHeader = 2048;  // barriers + slot table

template <typename F> __device__ __forceinline__ Cplx<F> lds_c(const char *p) {
  using V = typename vec2_of<F>::type;
  const V v = *reinterpret_cast<const V *>(p);
  return make_c<F>(v.x, v.y);
}

// The ultra-local spin matrix M0 = sum_n (1/sigma_n) v_n(x)^dag (x) v_n(x) is Hermitian: 4 real diagonal
// entries (index 0..3) and 6 complex entries be < al (index 4..9).  They are shared out among the threads
// that work on the same site for the displaced loops of the group (balanced to +-1 entry), so that the
// ultra-local loop costs no warp of its own.
__host__ __device__ constexpr int ul_pair_be(int e) { return e < 7 ? 0 : (e < 9 ? 1 : 2); }
__host__ __device__ constexpr int ul_pair_al(int e) { return e == 4 ? 1 : e == 5 ? 2 : e == 6 ? 3 : e == 7 ? 2 : 3; }
__host__ __device__ constexpr int ul_pair_index(int be, int al) {  // be < al
  return be == 0 ? 3 + al : be == 1 ? 5 + al : 9;
}

// rotate the first (kRow) or second index of a 4x4 matrix back: out[(b+K)&3][a] = in[b][a]
template <typename F, int K, bool kRow> __device__ __forceinline__ void unrotate(Cplx<F> M[4][4]) {
  Cplx<F> T[4][4];
#pragma unroll
  for (int b = 0; b < 4; b++)
#pragma unroll
    for (int a = 0; a < 4; a++) {
      if (kRow)
        T[(b + K) & 3][a] = M[b][a];
      else
        T[b][(a + K) & 3] = M[b][a];
    }
#pragma unroll
  for (int b = 0; b < 4; b++)
#pragma unroll
    for (int a = 0; a < 4; a++) M[b][a] = T[b][a];
}
template <typename F, bool kRow> __device__ __forceinline__ void unrotate_rt(Cplx<F> M[4][4], int k) {
  if (k == 1) unrotate<F, 1, kRow>(M);
  if (k == 2) unrotate<F, 2, kRow>(M);
  if (k == 3) unrotate<F, 3, kRow>(M);
}

// Everything a thread needs inside the eigenvector loop.
template <typename F> struct ThreadCtx {
  const char *stages;
  uint64_t *full, *empty;
  const SlotTable *st;
  int S, stage_bytes, nvec, ahead, nActive, warp, lane;
  int own_sp[4], nbr_sp[4];  // byte offsets (inside a stage) of the 4 rotated spin blocks of v(x) and v(x+d)
};

// Share of the ultra-local matrix a thread accumulates (compile-time: only the needed FMAs are issued).
//   UL_NONE : nothing
//   UL_ALL  : all 10 entries (groups with fewer than 4 displaced loops: role 0 does it alone)
//   UL_ROT  : groups with 4 displaced loops.  Role j reads v(x) with its spin labels rotated by j on top of the
//             bank rotation, and every role runs the SAME code: diagonal entry 0, pair (0,1), and - roles 0 and 1
//             only - pair (0,2), in its own labels.  Over j = 0..3 that is d0..d3, the four "adjacent" pairs
//             (0,1) (1,2) (2,3) (3,0) and the two "opposite" pairs (0,2) (1,3): all 10 entries exactly once,
//             with one loop body in the instruction cache instead of four (no_instruction stalls were 12%).
enum { UL_NONE = 0, UL_ALL = 1, UL_ROT = 2 };

// ---- hot-loop primitives on 32-bit shared addresses.  On B200 every non-FP64 instruction costs FP64 issue slots
// (tools/microbench.cu: one IMAD per DFMA drops the DFMA rate from 33.5 to 18.3 TFLOP/s at 8 warps per SM), so the loop
// body avoids branches (predicated mbarrier / TMA instructions instead of `if (lane == 0)` blocks), address
// arithmetic (running per-thread addresses, immediates for the colour offset) and integer division.
__device__ __forceinline__ void mbar_wait_u32(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_if(uint32_t bar, int pred) {
  asm volatile("{\n.reg .pred q;\nsetp.ne.b32 q, %1, 0;\n@q mbarrier.arrive.shared::cta.b64 _, [%0];\n}" ::"r"(bar), "r"(pred) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_if(uint32_t bar, uint32_t bytes, int pred) {
  asm volatile("{\n.reg .pred q;\nsetp.ne.b32 q, %2, 0;\n@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n}" ::"r"(bar),
               "r"(bytes), "r"(pred)
               : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s_if(uint32_t dst, const void *src_gmem, uint32_t bytes, uint32_t bar, int pred) {
  asm volatile(
      "{\n.reg .pred q;\nsetp.ne.b32 q, %4, 0;\n"
      "@q cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n}" ::"r"(dst),
      "l"(src_gmem), "r"(bytes), "r"(bar), "r"(pred)
      : "memory");
}
// The eigenvector loop of one role: ND displaced loops in the group, this thread works on one of them and on its
// share of the ultra-local entries.
template <typename F, int ND, int UL>
__device__ __forceinline__ void evec_loop(const FusedArgs<F> &A, const ThreadCtx<F> &c, const Cplx<F> (&W)[3][3],
                                          Cplx<F> (&M)[4][4], F (&Md)[4], Cplx<F> (&Mo)[6], const bool opposite) {
  // producer side: the warps take turns - warp (m mod nActive) issues ALL bulk copies of eigenvector m's stage (lane i
  // issues copy i), the others only advance their cursors.  Issuing costs ~100 non-FP64 instructions, which on this
  // chip are paid in FP64 issue slots; spread over the warps in turn it is ~15 per warp and eigenvector instead of ~65
  // when every warp issued its own share every time.
  uint32_t total_tx = 0;
  for (int i = 0; i < c.st->ncopies; i++) total_tx += (uint32_t)c.st->cp_bytes[i];
  const int lead = c.lane == 0;
  const uint32_t stages_u32 = smem_u32(c.stages);
  const uint32_t full_u32 = smem_u32(c.full), empty_u32 = smem_u32(c.empty);
  const int ring_bytes = c.S * c.stage_bytes;
  constexpr int kC = 2 * (int)sizeof(F);

  uint32_t p_full = full_u32, p_empty = empty_u32, p_dst = stages_u32;  // producer cursor (stage of the next issue)
  int p_left = c.S;                                                      // stages until the cursor wraps
  uint32_t p_par = 1;  // parity to wait for on the empty barrier; the first pass over the ring does not wait
  int turn = c.warp;   // issues when it reaches 0
  auto issue_share = [&](int m, bool wait) {
    if (turn == 0) {  // warp-uniform
      if (wait) mbar_wait_u32(p_empty, p_par);
      const char *ev = static_cast<const char *>(A.vt.evec[m]);
      mbar_expect_tx_if(p_full, total_tx, lead);
      for (int i = c.lane; i < c.st->ncopies; i += 32)
        tma_bulk_g2s_if(p_dst + (uint32_t)c.st->cp_soff[i], ev + ((size_t)c.st->cp_goff16[i] << 4), (uint32_t)c.st->cp_bytes[i],
                        p_full, 1);
      turn = c.nActive;
    }
    turn--;
    p_full += 8;
    p_empty += 8;
    p_dst += (uint32_t)c.stage_bytes;
    if (--p_left == 0) {
      p_left = c.S;
      p_full = full_u32;
      p_empty = empty_u32;
      p_dst = stages_u32;
      p_par ^= 1u;
    }
  };
  for (int m = 0; m < c.ahead && m < c.nvec; m++) issue_share(m, false);  // ahead < S: no tenant to wait for

  // consumer cursor: absolute shared addresses of the 4 rotated spin blocks of v(x) and v(x+d) in the current stage
  const char *a_own[4], *a_nbr[4];
#pragma unroll
  for (int b = 0; b < 4; b++) {
    a_own[b] = c.stages + c.own_sp[b];
    a_nbr[b] = c.stages + c.nbr_sp[b];
  }
  uint32_t c_full = full_u32, c_empty = empty_u32;
  int c_left = c.S;
  uint32_t c_par = 0;
  const int n_issue = c.nvec - c.ahead;  // iterations that still have a stage to issue
  for (int n = 0; n < c.nvec; n++) {
    if (n < n_issue) issue_share(n + c.ahead, n + c.ahead >= c.S);
    mbar_wait_u32(c_full, c_par);
    const F is = (F)A.vt.inv_sigma[n];
    Cplx<F> vp[12];
    if (ND > 0) {
#pragma unroll
      for (int al = 0; al < 4; al++) {
        vp[al * 3 + 0] = lds_c<F>(a_nbr[al]);
        vp[al * 3 + 1] = lds_c<F>(a_nbr[al] + kC);
        vp[al * 3 + 2] = lds_c<F>(a_nbr[al] + 2 * kC);
      }
    }
#pragma unroll
    for (int cc = 0; cc < 3; cc++) {
      Cplx<F> lc[4];
#pragma unroll
      for (int be = 0; be < 4; be++) lc[be] = lds_c<F>(a_own[be] + cc * kC);
      if (cc == 2) {  // last shared-memory read of this stage: hand it back before the remaining FMAs
        __syncwarp();
        mbar_arrive_if(c_empty, lead);
      }
      // (1/sigma) v(x): one scaling serves the displaced and the ultra-local accumulation
      Cplx<F> ls[4];
#pragma unroll
      for (int be = 0; be < 4; be++) ls[be] = make_c<F>(lc[be].re * is, lc[be].im * is);
      if (ND > 0) {
        Cplx<F> Rc[4];
#pragma unroll
        for (int al = 0; al < 4; al++) {
          Rc[al] = cmul(W[cc][0], vp[al * 3 + 0]);
          cmac(Rc[al], W[cc][1], vp[al * 3 + 1]);
          cmac(Rc[al], W[cc][2], vp[al * 3 + 2]);
        }
#pragma unroll
        for (int be = 0; be < 4; be++)
#pragma unroll
          for (int al = 0; al < 4; al++) cmac_conj(M[be][al], ls[be], Rc[al]);
      }
      if (UL == UL_ALL) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
          Md[k] = fma(ls[k].re, lc[k].re, Md[k]);
          Md[k] = fma(ls[k].im, lc[k].im, Md[k]);
        }
#pragma unroll
        for (int k = 4; k < 10; k++) cmac_conj(Mo[k - 4], ls[ul_pair_be(k)], lc[ul_pair_al(k)]);
      }
      if (UL == UL_ROT) {
        Md[0] = fma(ls[0].re, lc[0].re, Md[0]);
        Md[0] = fma(ls[0].im, lc[0].im, Md[0]);
        cmac_conj(Mo[0], ls[0], lc[1]);
        cmac_conj(Mo[1], ls[0], lc[2]);  // needed from roles 0 and 1 only; computing it everywhere keeps the code uniform
      }
    }
    // next stage: the eight running addresses move on (one add each instead of recomputing base + offset)
    c_full += 8;
    c_empty += 8;
#pragma unroll
    for (int b = 0; b < 4; b++) {
      a_own[b] += c.stage_bytes;
      a_nbr[b] += c.stage_bytes;
    }
    if (--c_left == 0) {
      c_left = c.S;
      c_full = full_u32;
      c_empty = empty_u32;
      c_par ^= 1u;
#pragma unroll
      for (int b = 0; b < 4; b++) {
        a_own[b] -= ring_bytes;
        a_nbr[b] -= ring_bytes;
      }
    }
  }
}

template <typename F, int ND>
__global__ void __launch_bounds__(kFusedThreads, 1) loop_fused_kernel(const __grid_constant__ FusedArgs<F> A) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t *full = reinterpret_cast<uint64_t *>(smem);        // [nstages]
  uint64_t *empty = reinterpret_cast<uint64_t *>(smem + 64);  // [nstages]
  SlotTable &st = *reinterpret_cast<SlotTable *>(smem + 128);
  const LatGeom &g = A.g;
  const FusedTiling &tl = A.tl;
  F *xch = reinterpret_cast<F *>(smem + kSmemHeader);  // [units*32][16] ultra-local entries, true spin labels
  char *stages = reinterpret_cast<char *>(smem + kSmemHeader + tl.units * 32 * 16 * (int)sizeof(F));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool has_ul = A.ul_off >= 0;
  constexpr int nrole = ND > 0 ? ND : 1;  // a launch without displaced loops runs one pure ultra-local role
  const int nActive = nrole * tl.units;   // compute warps in use
  constexpr int kSite = 24 * (int)sizeof(F);
  const int Lh = g.Lh;

  const int bid = blockIdx.x;
  const int y0 = (bid % tl.nTy) * tl.TY;
  const int z0 = ((bid / tl.nTy) % tl.nTz) * tl.TZ;
  const int t0 = (bid / (tl.nTy * tl.nTz) + A.tt0) * tl.TT;

  if (threadIdx.x == 0) {
    build_slots(st, A.grp, tl, g, kSite, y0, z0, t0);
    for (int s = 0; s < tl.nstages; s++) {
      mbar_init(&full[s], 1);        // one arrive.expect_tx by the warp whose turn it is
      mbar_init(&empty[s], nActive);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const bool active = warp < nActive;
  // ---- role of this thread: displaced loop j (if any) on site q of parity p ------------------------------------
  const int j = active ? warp / tl.units : 0, u = active ? warp % tl.units : 0;
  const int upp = tl.units >> 1;  // warps per parity
  const int nsite = tl.NR * Lh;   // sites of one parity in the tile
  const int p = u / upp;
  int q = (u % upp) * 32 + lane;
  const bool valid = active && q < nsite;
  if (!valid) q = 0;                      // park on a real site; nothing is stored
  const int i = q / Lh, sx = q - i * Lh;  // own row, position in the half-row
  const int ya = y0 + i % tl.TY, za = z0 + (i / tl.TY) % tl.TZ, ta = t0 + i / (tl.TY * tl.TZ);
  const int x = 2 * sx + ((ya + za + ta + p) & 1);
  const size_t x_eo = (size_t)p * g.volumeCB + (size_t)st.row[i] * Lh + sx;
  const int hrb = Lh * kSite;
  const bool ul_rot = has_ul && ND == 4;              // see UL_ROT
  const int k_own = (((q >> 1) & 3) + (ul_rot ? j : 0)) & 3;  // spin-label rotation of v(x): bank rotation + role rotation
  int k_nbr = 0;

  ThreadCtx<F> c;
  c.stages = stages;
  c.full = full;
  c.empty = empty;
  c.st = &st;
  c.S = tl.nstages;
  c.stage_bytes = tl.stage_bytes;
  c.nvec = A.vt.nvec;
  c.ahead = c.S > 2 ? c.S - 2 : 1;  // stages in flight beyond the one being consumed
  c.nActive = nActive;
  c.warp = warp;
  c.lane = lane;
  {
    const int off = p * st.nslots * hrb + q * kSite;
#pragma unroll
    for (int b = 0; b < 4; b++) c.own_sp[b] = off + ((b + k_own) & 3) * (kSite / 4);
  }
  const FusedLoop lp = A.grp.loop[ND > 0 ? j : 0];
  if (ND > 0) {
    const int pn = (p + lp.len) & 1;
    int slot = i, sn = sx;
    if (lp.dir == 0)
      sn = wrap(x + lp.sign * lp.len, g.L[0]) >> 1;
    else
      slot = st.nbr[j][i];
    const int qn = slot * Lh + sn;
    k_nbr = (qn >> 1) & 3;
    const int off = pn * st.nslots * hrb + qn * kSite;
#pragma unroll
    for (int b = 0; b < 4; b++) c.nbr_sp[b] = off + ((b + k_nbr) & 3) * (kSite / 4);
  } else {
#pragma unroll
    for (int b = 0; b < 4; b++) c.nbr_sp[b] = c.own_sp[b];
  }

  Cplx<F> M[4][4];  // displaced loop, rotated labels: M[b][a] = true M[(b+k_own)&3][(a+k_nbr)&3]
  F Md[4];          // ultra-local diagonal, rotated labels
  Cplx<F> Mo[6];    // ultra-local be < al, rotated labels
#pragma unroll
  for (int be = 0; be < 4; be++) {
    Md[be] = 0;
#pragma unroll
    for (int al = 0; al < 4; al++) M[be][al] = make_c<F>(0, 0);
  }
#pragma unroll
  for (int k = 0; k < 6; k++) Mo[k] = make_c<F>(0, 0);

  Cplx<F> W[3][3];
#pragma unroll
  for (int k = 0; k < 9; k++) W[k / 3][k % 3] = make_c<F>(0, 0);
  if (ND > 0 && active) {
    const F *pw = static_cast<const F *>(lp.W) + x_eo * (2 * kLinkLen);
#pragma unroll
    for (int k = 0; k < 9; k++) W[k / 3][k % 3] = ldg_c<F>(pw + 2 * k);
  }

  const int ul_mode = !has_ul ? UL_NONE : (ND == 4 ? UL_ROT : (j == 0 ? UL_ALL : UL_NONE));
  if (active) {
    if (ul_mode == UL_NONE)
      evec_loop<F, ND, UL_NONE>(A, c, W, M, Md, Mo, false);
    else if (ul_mode == UL_ALL)
      evec_loop<F, ND, UL_ALL>(A, c, W, M, Md, Mo, false);
    else
      evec_loop<F, ND, UL_ROT>(A, c, W, M, Md, Mo, j < 2);
  }

  // ---- epilogue: undo the spin rotation, gamma projection (adds/swaps only), one write of the loop buffer ---------
  const int nid = u * 32 + lane;
  if (has_ul && valid && ul_mode != UL_NONE) {  // publish this thread's share under the true spin labels
    F *px = xch + (size_t)nid * 16;
    auto put_pair = [&](int bl, int al, const Cplx<F> z) {  // entry (bl, al) in this thread's labels
      const int b = (bl + k_own) & 3, a = (al + k_own) & 3;
      const int e = b < a ? ul_pair_index(b, a) : ul_pair_index(a, b);
      px[4 + 2 * (e - 4)] = z.re;
      px[5 + 2 * (e - 4)] = b < a ? z.im : -z.im;
    };
    if (ul_mode == UL_ALL) {
#pragma unroll
      for (int k = 0; k < 4; k++) px[(k + k_own) & 3] = Md[k];
#pragma unroll
      for (int k = 4; k < 10; k++) put_pair(ul_pair_be(k), ul_pair_al(k), Mo[k - 4]);
    } else {
      px[k_own & 3] = Md[0];
      put_pair(0, 1, Mo[0]);
      if (j < 2) put_pair(0, 2, Mo[1]);
    }
  }
  __syncthreads();
  if (!valid) return;
  if (ND > 0) {
    unrotate_rt<F, true>(M, k_own);
    unrotate_rt<F, false>(M, k_nbr);
    Cplx<F> T[16];
    gamma_project(T, M);
    F *out = A.dataPos + 2 * ((size_t)lp.out_off + x_eo);
#pragma unroll
    for (int G = 0; G < 16; G++) {
      F *po = out + 2 * (size_t)g.volume * G;
      Cplx<F> o = T[G];
      if (A.accumulate) {
        const Cplx<F> old = ldg_c<F>(po);
        o.re += old.re;
        o.im += old.im;
      }
      st_c<F>(po, o);
    }
  }
  if (has_ul && j == 0) {  // role 0 of every site gathers the Hermitian matrix and writes the ultra-local loop
    const F *px = xch + (size_t)nid * 16;
    Cplx<F> M0[4][4];
#pragma unroll
    for (int k = 0; k < 4; k++) M0[k][k] = make_c<F>(px[k], 0);
#pragma unroll
    for (int k = 4; k < 10; k++) {
      const int be = ul_pair_be(k), al = ul_pair_al(k);
      const F re = px[4 + 2 * (k - 4)], im = px[5 + 2 * (k - 4)];
      M0[be][al] = make_c<F>(re, im);
      M0[al][be] = make_c<F>(re, -im);
    }
    Cplx<F> T[16];
    gamma_project(T, M0);
    F *out = A.dataPos + 2 * ((size_t)A.ul_off + x_eo);
#pragma unroll
    for (int G = 0; G < 16; G++) {
      F *po = out + 2 * (size_t)g.volume * G;
      Cplx<F> o = T[G];
      if (A.accumulate) {
        const Cplx<F> old = ldg_c<F>(po);
        o.re += old.re;
        o.im += old.im;
      }
      st_c<F>(po, o);
    }
  }
}

// ---- host side: tiling and launch ----------------------------------------------------------------------------
static bool choose_tiling(FusedTiling &tl, const FusedGroup &grp, const LatGeom &g, int precision, int smem_limit) {
  // consecutive-y rows first: they are contiguous in memory, so own rows and shifted rows arrive as few large copies
  static const int cand[][3] = {{4, 1, 1}, {2, 2, 1}, {2, 1, 2}, {1, 2, 2}, {1, 4, 1}, {1, 1, 4},
                                {2, 1, 1}, {1, 2, 1}, {1, 1, 2}, {1, 1, 1}};
  const int site = 24 * (int)prec_bytes(precision);
  const int nrole = grp.nloops > 0 ? grp.nloops : 1;
  for (const auto &c : cand) {
    if (g.L[1] % c[0] || g.L[2] % c[1] || g.L[3] % c[2]) continue;
    tl.TY = c[0];
    tl.TZ = c[1];
    tl.TT = c[2];
    tl.NR = c[0] * c[1] * c[2];
    tl.nTy = g.L[1] / c[0];
    tl.nTz = g.L[2] / c[1];
    tl.nTt = g.L[3] / c[2];
    tl.units = 2 * ((tl.NR * g.Lh + 31) / 32);
    if (tl.units * nrole > kFusedComputeWarps) continue;
    SlotTable st;
    build_slots(st, grp, tl, g, site, 0, 0, 0);
    tl.nslots = st.nslots;
    tl.stage_bytes = (tl.nslots * 2 * g.Lh * site + 127) / 128 * 128;
    tl.nstages = (smem_limit - kSmemHeader - tl.units * 32 * 16 * (int)prec_bytes(precision)) / tl.stage_bytes;
    if (tl.nstages > 8) tl.nstages = 8;
    if (const char *e = getenv("MUGIQ_B200_FUSED_STAGES")) {
      const int want = atoi(e);
      if (want >= 2 && want < tl.nstages) tl.nstages = want;
    }
    if (tl.nstages >= 2) return true;
  }
  return false;
}

static int smem_limit_bytes() {
  static int limit = -1;
  if (limit < 0) {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) == cudaSuccess)
      limit = v;
    else
      limit = 227 * 1024;
  }
  return limit;
}

int fused_max_loops_per_group(const LatGeom &g, int precision) {
  // worst case: every displaced loop of the group shifts rows in y, z or t and nothing de-duplicates
  for (int nl = kFusedMaxLoops; nl >= 0; nl--) {
    FusedGroup grp;
    grp.nloops = nl;
    for (int j = 0; j < nl; j++) {
      grp.loop[j].dir = 3;
      grp.loop[j].sign = 1;
      grp.loop[j].len = 2 * (j + 1) + 1;  // distinct far shifts
      grp.loop[j].W = nullptr;
      grp.loop[j].out_off = 0;
    }
    FusedTiling tl;
    if (choose_tiling(tl, grp, g, precision, smem_limit_bytes())) return nl;
  }
  return -1;
}

template <typename F, int ND> static int launch_fused_nd(const FusedArgs<F> &args, size_t smem, int grid, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    MUGIQ_CUDA_CHECK(cudaFuncSetAttribute(loop_fused_kernel<F, ND>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          smem_limit_bytes()));
    attr_set = true;
  }
  loop_fused_kernel<F, ND><<<grid, kFusedThreads, smem, stream>>>(args);
  MUGIQ_LAUNCH_CHECK();
  return MUGIQ_B200_OK;
}

template <typename F>
static int launch_fused(void *dataPos_d, const FusedGroup &grp, long long ul_off, const FusedVecTable &vt, int accumulate,
                        const LatGeom &g, int precision, cudaStream_t stream, int t_begin, int t_end) {
  FusedArgs<F> args;
  args.g = g;
  args.grp = grp;
  args.vt = vt;
  args.ul_off = ul_off;
  args.dataPos = static_cast<F *>(dataPos_d);
  args.accumulate = accumulate;
  if (!choose_tiling(args.tl, grp, g, precision, smem_limit_bytes()))
    return set_error(MUGIQ_B200_EINVAL, "loop_fused: no tiling fits %d loops on a %dx%dx%dx%d lattice", grp.nloops, g.L[0],
                     g.L[1], g.L[2], g.L[3]);
  // time-slice range -> whole tiles in t
  const int tt0 = t_begin / args.tl.TT, tt1 = (t_end + args.tl.TT - 1) / args.tl.TT;
  args.tt0 = tt0;
  const int grid = args.tl.nTy * args.tl.nTz * (tt1 - tt0);
  const double frac = (double)(tt1 - tt0) / (double)args.tl.nTt;  // share of the lattice this launch computes
  const size_t smem =
      kSmemHeader + (size_t)args.tl.units * 32 * 16 * sizeof(F) + (size_t)args.tl.nstages * args.tl.stage_bytes;
  // algorithmic (compulsory) bytes: every eigenvector site once, every link once, the accumulators once
  const double S = 24.0 * sizeof(F), U = 18.0 * sizeof(F), Acc = 32.0 * sizeof(F);
  const int nl = grp.nloops + (ul_off >= 0 ? 1 : 0);
  // FP64 work per (eigvec, site): 336 DFMA + 24 DMUL per displaced loop (W v: 144, scale: 24, colour trace: 192) and
  // 96 DFMA for the Hermitian ultra-local matrix (+24 DMUL when it runs alone); FMA = 2 flop
  const double flop_site = grp.nloops * (336.0 * 2 + 24.0) + (ul_off >= 0 ? 96.0 * 2 + (grp.nloops == 0 ? 24.0 : 0.0) : 0.0);
  ProfScope prof(K_LOOP_FUSED, stream, frac * (double)g.volume * (vt.nvec * S + grp.nloops * U + nl * Acc * (accumulate ? 2 : 1)),
                 frac * (double)g.volume * vt.nvec * flop_site);
  switch (grp.nloops) {
    case 0: return launch_fused_nd<F, 0>(args, smem, grid, stream);
    case 1: return launch_fused_nd<F, 1>(args, smem, grid, stream);
    case 2: return launch_fused_nd<F, 2>(args, smem, grid, stream);
    case 3: return launch_fused_nd<F, 3>(args, smem, grid, stream);
    default: return launch_fused_nd<F, 4>(args, smem, grid, stream);
  }
}

int fused_group_launch(void *dataPos_d, const FusedGroup &grp, long long ul_off, const FusedVecTable &vt, int accumulate,
                       const LatGeom &g, int precision, cudaStream_t stream, int t_begin, int t_end) {
  if (t_end < 0) t_end = g.L[3];
  if (t_begin < 0 || t_begin >= t_end || t_end > g.L[3])
    return set_error(MUGIQ_B200_EINVAL, "loop_fused: bad time-slice range [%d, %d) on Lt = %d", t_begin, t_end, g.L[3]);
  if (grp.nloops < 0 || grp.nloops > kFusedMaxLoops || (grp.nloops == 0 && ul_off < 0) || vt.nvec < 1 ||
      vt.nvec > kFusedMaxVec)
    return set_error(MUGIQ_B200_EINVAL, "loop_fused: bad group (%d loops, %d eigenvectors)", grp.nloops, vt.nvec);
  return precision == MUGIQ_B200_PREC_DOUBLE
             ? launch_fused<double>(dataPos_d, grp, ul_off, vt, accumulate, g, precision, stream, t_begin, t_end)
             : launch_fused<float>(dataPos_d, grp, ul_off, vt, accumulate, g, precision, stream, t_begin, t_end);
}

}  // namespace mugiq_b200
