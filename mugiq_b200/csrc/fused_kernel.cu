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
// Design (the numbers quoted are measured on B200, profiles/):
//  * The kernel is FP64-pipe bound: 360 DFMA/DMUL per (eigvec, site, displaced loop) + ~30 for the share of the ultra-local
//    matrix, against 192 B of compulsory HBM traffic per (eigvec, site).  What the pipe delivers depends on the operands: a
//    chained stream (two operands shared by all instructions) issues one DFMA per 2.22 cycles and SM sub-partition, the
//    complex 4x4 outer products of this kernel (two fresh 64-bit register operands per instruction) one per 2.52 cycles
//    (tools/dfma_bench.cu); the eigenvector loop runs at 2.6.  Integer / uniform instructions issue in the shadow of the
//    DFMAs and are almost free, an LDS.128 costs ~4 cycles of its warp, a barrier wait whose result is consumed at once ~50
//    (tools/issue_bench.cu) - which is why the compute warps do nothing but load, multiply and release (see below).
//  * CTA tile = a RUN of 32*k consecutive checkerboard sites [c0, c0 + run) of BOTH parities (k = 1 with 3 or 4
//    displaced loops in the group).  In the even/odd site-major layout a lattice row (all x at fixed y,z,t) is two
//    contiguous half-rows (one per parity) of Lx/2 sites x 192 B and rows consecutive in y are contiguous, so for
//    Lx/2 = 8, 16, 32 a run is 4, 2, 1 whole rows; for Lx/2 = 12, 24 (24^3x48, 48^3x96) it is a fractional number of
//    rows and still fills every lane of every warp.  What a stage holds per eigenvector is a union of INTERVALS of
//    checkerboard-index space: the run itself, and for every loop of the group the image of each row piece of the run
//    under the shift.  Overlapping and adjoining intervals are merged, so every interval is ONE bulk-TMA copy
//    (cp.async.bulk + mbarrier complete_tx) and shared sites are fetched once.  Layout in shared memory is dense:
//    [interval][site][12 complex].  The stage maps of all CTAs of a launch shape are built once on the device
//    (stage_maps_kernel) and cached; a CTA copies its 1.8 KB.
//  * 12 warps: 8 compute warps (2 per SM sub-partition; 232 registers each after setmaxnreg) + a producer warp group of
//    which one warp issues every bulk copy of every stage and waits for the empty barriers (40 registers).  Compute warp =
//    (loop, parity-half of the tile), thread = (site, loop); it owns the 4x4 complex spin matrix M (32 doubles), the 3x3 link
//    W (18) and its share of the Hermitian ultra-local matrix for the whole batch.  No block barrier in the eigenvector loop.
//  * Bank conflicts: 192-B site stride means lanes reading the same component hit only two 16-B bank groups.
//    Instead of padding (which would forbid multi-row bulk copies) each lane reads its site with the spin index
//    rotated by k = (lane/2) mod 4, i.e. it keeps spin (b+k) mod 4 in register slot b.  The 8 lanes of a
//    quarter-warp then touch 8 distinct bank groups (conflict-free 128-bit reads); the rotation only relabels
//    M[be][al] and is undone once in the epilogue.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <map>
#include <mutex>
#include <tuple>
#include <utility>
#include <vector>

#include "fused.cuh"
#include "fused_stage.cuh"

namespace mugiq_b200 {

// per-CTA timeline (mugiq_b200_prof_fused_trace): thread 0 stamps %globaltimer at the marks of the kernel
static long long *g_fused_trace = nullptr;
static int g_fused_trace_ctas = 0;
void fused_set_trace(long long *trace_d, long long capacity_ctas) {
  g_fused_trace = trace_d;
  g_fused_trace_ctas = trace_d ? (int)std::min<long long>(capacity_ctas, 1 << 30) : 0;
}
template <typename F> __device__ __forceinline__ void trace_mark(const FusedArgs<F> &A, int slot) {
  if (A.trace != nullptr && threadIdx.x == 0 && (int)blockIdx.x < A.trace_ctas) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    A.trace[16 * (size_t)blockIdx.x + slot] = (long long)t;
    A.trace[16 * (size_t)blockIdx.x + 8 + slot] = clock64();
    if (slot == 1) {
      unsigned sm;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
      A.trace[16 * (size_t)blockIdx.x] = sm;
    }
  }
}

// ---- the eigenvector loop, warp-specialised ----------------------------------------------------------------------------------
// Measured (tools/issue_bench.cu, tools/dfma_bench.cu, tools/fused_trace.py, ncu instruction samples): with
// two warps per SM sub-partition the FP64 pipe is busy only while BOTH are in FP64 code, so everything else a compute warp
// does is paid in pipe time.  In the round-2 loop the compute warps took turns issuing the bulk copies: ~200-260 cycles per
// eigenvector and warp (integer division for the ring position, a blocking wait on the empty barrier, the copy list), plus a
// full-barrier try_wait whose result the next instruction needed (~50 cycles).  Now a PRODUCER warp (its own warp group, its
// registers handed to the compute warps with setmaxnreg) issues every copy; a compute warp tests the full barrier of the NEXT
// stage half an eigenvector ahead (non-blocking mbarrier.test_wait) and only spins if that test failed.  Loop body: 390 FP64
// + 24 LDS.128 + ~45 integer / control instructions, 2040 cycles per eigenvector at configs[1] (1560 would be one FP64
// instruction per 2 cycles; the bare outer-product stream needs 1960).
__device__ __forceinline__ uint32_t mbar_test_u32(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred P1;\nmbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}"
               : "=r"(ok)
               : "r"(bar), "r"(parity)
               : "memory");
  return ok;
}

template <typename F, int NATIVE>
__device__ __forceinline__ void producer_loop(const FusedArgs<F> &A, const StageMap &st, uint32_t stages_u32, uint32_t bar_u32, int S,
                                              int stage_bytes, int lane) {
  uint32_t total_tx = 0;
  for (int i = 0; i < st.ncp; i++) total_tx += (uint32_t)st.cp_bytes[i];
  const int lead = lane == 0;
  const int nvec = A.vt.nvec;
  int sm = 0;           // stage of eigenvector m
  uint32_t par = 1;     // parity of the empty barrier's phase BEFORE its first completion: the first S waits pass at once
  for (int m = 0; m < nvec; m++) {
    const uint32_t p_full = bar_u32 + 8u * (uint32_t)sm, p_dst = stages_u32 + (uint32_t)(sm * stage_bytes);
    mbar_wait_u32(p_full + 64, par);  // the stage's previous tenant has been consumed by every compute warp
    const char *ev = static_cast<const char *>(A.vt.evec[m]);
    mbar_expect_tx_if(p_full, total_tx, lead);
    if (NATIVE) {  // ev: this eigenvector's three tensor maps (boxes of 1, 2, 4 chunks) in device memory
      for (int i = lane; i < st.ncp; i += 32) {
        const int d = st.cp_goff16[i];
        tma_tensor4_g2s(p_dst + (uint32_t)st.cp_soff[i], ev + ((d >> 29) & 3) * 128, d & 0x0fffffff, (d >> 28) & 1, p_full);
      }
    } else {
      for (int i = lane; i < st.ncp; i += 32)
        tma_bulk_g2s_if(p_dst + (uint32_t)st.cp_soff[i], ev + ((size_t)st.cp_goff16[i] << 4), (uint32_t)st.cp_bytes[i], p_full, 1);
    }
    if (++sm == S) {
      sm = 0;
      par ^= 1u;
    }
  }
}

template <typename F, int ND, int UL, int NATIVE>
__device__ __forceinline__ void consumer_loop(const FusedArgs<F> &A, const ThreadCtx<F> &c, const Cplx<F> (&W)[3][3],
                                              Cplx<F> (&M)[4][4], F (&Md)[4], Cplx<F> (&Mo)[6]) {
  const int lead = c.lane == 0;
  const uint32_t bar_u32 = smem_u32(c.full);  // full[s] at bar_u32 + 8 s, empty[s] 64 bytes further
  const int ring_bytes = c.S * c.stage_bytes;
  constexpr int kC = NATIVE ? kChunk * 2 * (int)sizeof(F) : 2 * (int)sizeof(F);
  const char *a_own[4], *a_nbr[4];
#pragma unroll
  for (int b = 0; b < 4; b++) {
    a_own[b] = c.stages + c.own_sp[b];
    a_nbr[b] = c.stages + c.nbr_sp[b];
  }
  uint32_t c_bar = bar_u32, c_par = 0;
  int c_left = c.S;
  uint32_t ok = mbar_test_u32(c_bar, c_par);
  for (int n = 0; n < c.nvec; n++) {
    if (!ok) mbar_wait_u32(c_bar, c_par);
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
    const uint32_t bar_n = c_bar;
#pragma unroll
    for (int cc = 0; cc < 3; cc++) {
      Cplx<F> lc[4];
#pragma unroll
      for (int be = 0; be < 4; be++) lc[be] = lds_c<F>(a_own[be] + cc * kC);
      if (cc == 1) {  // look at the next stage's barrier now, use the answer at the top of the next iteration
        c_bar += 8;
        if (--c_left == 0) {
          c_left = c.S;
          c_bar = bar_u32;
          c_par ^= 1u;
        }
        ok = mbar_test_u32(c_bar, c_par);
      }
      if (cc == 2) {  // last shared-memory read of this stage: hand it back before the remaining FMAs
        __syncwarp();
        mbar_arrive_if(bar_n + 64, lead);
      }
      // 1/sigma rides on the link row (6 multiplications per colour instead of 8 on v(x)); the ultra-local share scales the
      // entries it needs itself (as they are used: no scaled copy of v(x) is kept, the registers are needed elsewhere)
      constexpr bool kScaleOwn = ND == 0;
      Cplx<F> ls[4];
      if (kScaleOwn) {
#pragma unroll
        for (int be = 0; be < 4; be++) ls[be] = make_c<F>(lc[be].re * is, lc[be].im * is);
      }
      if (ND > 0) {
        Cplx<F> Ws[3], Rc[4];
#pragma unroll
        for (int k = 0; k < 3; k++) Ws[k] = kScaleOwn ? W[cc][k] : make_c<F>(W[cc][k].re * is, W[cc][k].im * is);
#pragma unroll
        for (int al = 0; al < 4; al++) {
          Rc[al] = cmul(Ws[0], vp[al * 3 + 0]);
          cmac(Rc[al], Ws[1], vp[al * 3 + 1]);
          cmac(Rc[al], Ws[2], vp[al * 3 + 2]);
        }
#pragma unroll
        for (int be = 0; be < 4; be++)
#pragma unroll
          for (int al = 0; al < 4; al++) cmac_conj(M[be][al], kScaleOwn ? ls[be] : lc[be], Rc[al]);
      }
      if (UL == UL_ALL) {
#pragma unroll
        for (int be = 0; be < 4; be++) {
          const Cplx<F> l = kScaleOwn ? ls[be] : make_c<F>(lc[be].re * is, lc[be].im * is);
          Md[be] = fma(l.re, lc[be].re, Md[be]);
          Md[be] = fma(l.im, lc[be].im, Md[be]);
#pragma unroll
          for (int k = 4; k < 10; k++)
            if (ul_pair_be(k) == be) cmac_conj(Mo[k - 4], l, lc[ul_pair_al(k)]);
        }
      }
      if (UL == UL_ROT) {
        const Cplx<F> l0 = make_c<F>(lc[0].re * is, lc[0].im * is);
        Md[0] = fma(l0.re, lc[0].re, Md[0]);
        Md[0] = fma(l0.im, lc[0].im, Md[0]);
        cmac_conj(Mo[0], l0, lc[1]);
        cmac_conj(Mo[1], l0, lc[2]);  // needed from roles 0 and 1 only; computing it everywhere keeps the code uniform
      }
    }
    // next stage: the eight running addresses move on; c_left already counts the NEXT stage
#pragma unroll
    for (int b = 0; b < 4; b++) {
      a_own[b] += c.stage_bytes;
      a_nbr[b] += c.stage_bytes;
    }
    if (c_left == c.S) {
#pragma unroll
      for (int b = 0; b < 4; b++) {
        a_own[b] -= ring_bytes;
        a_nbr[b] -= ring_bytes;
      }
    }
  }
}

template <typename F, int ND, int NATIVE>
__global__ void __launch_bounds__(kFusedThreadsWS, 1) loop_fused_kernel(const __grid_constant__ FusedArgs<F> A) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t *full = reinterpret_cast<uint64_t *>(smem);        // [nstages]
  uint64_t *empty = reinterpret_cast<uint64_t *>(smem + 64);  // [nstages]
  StageMap &st = *reinterpret_cast<StageMap *>(smem + 128);
  const LatGeom &g = A.g;
  const FusedTiling &tl = A.tl;
  F *xch = reinterpret_cast<F *>(smem + kSmemHeader);  // [units*32][16] ultra-local entries, true spin labels
  char *stages = reinterpret_cast<char *>(smem + kSmemHeader + tl.units * 32 * 16 * (int)sizeof(F));

  trace_mark(A, 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool has_ul = A.ul_off >= 0;
  // roles = threads working on one site: one per displaced loop; the ultra-local matrix rides with them when there are four
  // (UL_ROT: shared out, same code in every warp), is the only role when there are none, and otherwise gets a role of its
  // own - a warp that only reads v(x) - instead of loading role 0 with all ten entries (474 against 354 FP64 instructions
  // per eigenvector and more registers than the compute warps have)
  const bool ul_role = has_ul && ND >= 1 && ND <= 3;
  const int nrole = (ND > 0 ? ND : 1) + (ul_role ? 1 : 0);
  const int nActive = nrole * tl.units;   // compute warps in use
  constexpr int kSite = 24 * (int)sizeof(F);

  // this CTA's run of checkerboard sites (both parities)
  const int c0 = A.c_begin + (int)blockIdx.x * tl.run;
  const int c1 = min(c0 + tl.run, A.c_end);

  {  // this CTA's stage map, worked out once per (lattice, group, range) by stage_maps_kernel
    const int4 *src = reinterpret_cast<const int4 *>(A.maps + blockIdx.x);
    int4 *dst = reinterpret_cast<int4 *>(&st);
    for (int i = threadIdx.x; i < (int)(sizeof(StageMap) / sizeof(int4)); i += blockDim.x) dst[i] = __ldg(src + i);
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < tl.nstages; s++) {
      mbar_init(&full[s], 1);        // one arrive.expect_tx by the producer warp
      mbar_init(&empty[s], nActive);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  trace_mark(A, 2);
  if (warp >= kFusedComputeWarps) {  // producer warp group: keeps 40 registers per thread, one warp works
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == kFusedComputeWarps)
      producer_loop<F, NATIVE>(A, st, smem_u32(stages), smem_u32(full), tl.nstages, tl.stage_bytes, lane);
    return;
  }
  // 8 x 32 x 232 + 4 x 32 x 40 = 64512 of the 65536 registers of an SM
  asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");

  const bool active = warp < nActive;
  // ---- role of this thread: displaced loop j (if any) on site c0 + q of parity p -------------------------------
  const int j = active ? warp / tl.units : 0, u = active ? warp % tl.units : 0;
  const int upp = tl.units >> 1;  // warps per parity
  const int p = u / upp;
  int q = (u % upp) * 32 + lane;
  const bool valid = active && c0 + q < c1;
  if (!valid) q = 0;  // park on a real site; nothing is stored
  const int cb = c0 + q;
  const size_t x_eo = (size_t)p * g.volumeCB + (size_t)cb;
  const bool ul_rot = has_ul && ND == 4;  // see UL_ROT
  const bool ul_thread = ul_role && j == ND;  // this warp accumulates the ultra-local matrix and nothing else
  const int s_own = max(stage_site(st, p, cb), 0);
  // Spin-label rotation against bank conflicts: the 8 lanes of a quarter-warp read 16 B each at a 192-byte site stride,
  // i.e. bank group (12 s + 3 k + c) mod 8 for stage position s and rotation k; with k = (lane / 2) mod 4 the groups are
  // distinct whenever the two lanes of a pair sit on stage positions of different parity - consecutive sites, also across
  // the jump between two row pieces of a run (Lx/2 = 12, 24), where a position-based rotation collided.
  // A FLOAT2 stage needs none: consecutive sites are consecutive 16-byte words of a component row.
  const int k_bank = NATIVE ? 0 : (lane >> 1) & 3;
  const int k_own = (k_bank + (ul_rot ? j : 0)) & 3;  // bank rotation + role rotation of the ultra-local share
  int k_nbr = 0;

  ThreadCtx<F> c;
  c.stages = stages;
  c.full = full;
  c.empty = empty;
  c.st = &st;
  c.S = tl.nstages;
  c.stage_bytes = tl.stage_bytes;
  c.nvec = A.vt.nvec;
  c.nActive = nActive;
  c.warp = warp;
  c.lane = lane;
  constexpr int kSpin = NATIVE ? 3 * kChunk * 2 * (int)sizeof(F) : kSite / 4;  // distance between the spin blocks of a site
  {
    const int off = stage_site_bytes<NATIVE>(s_own, kSite);
#pragma unroll
    for (int b = 0; b < 4; b++) c.own_sp[b] = off + ((b + k_own) & 3) * kSpin;
  }
  const FusedLoop lp = A.grp.loop[ND > 0 ? min(j, ND - 1) : 0];
  if (ND > 0 && !ul_thread) {
    // neighbour x + sign*len*dir and its place in the stage
    const int s_nbr = max(stage_site(st, (p + lp.len) & 1, neighbour_cb(g, lp, p, cb)), 0);
    k_nbr = k_bank;
    const int off = stage_site_bytes<NATIVE>(s_nbr, kSite);
#pragma unroll
    for (int b = 0; b < 4; b++) c.nbr_sp[b] = off + ((b + k_nbr) & 3) * kSpin;
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
  if (ND > 0 && active && !ul_thread) {
    const F *pw = static_cast<const F *>(lp.W) + x_eo * (2 * kLinkLen);
#pragma unroll
    for (int k = 0; k < 9; k++) W[k / 3][k % 3] = ldg_c<F>(pw + 2 * k);
  }

  const int ul_mode = !has_ul ? UL_NONE : (ND == 4 ? UL_ROT : ((ND == 0 || ul_thread) ? UL_ALL : UL_NONE));
  trace_mark(A, 3);
  if (active) {
    if (ul_mode == UL_NONE)
      consumer_loop<F, ND, UL_NONE, NATIVE>(A, c, W, M, Md, Mo);
    else if (ul_mode == UL_ALL)
      consumer_loop<F, 0, UL_ALL, NATIVE>(A, c, W, M, Md, Mo);
    else
      consumer_loop<F, ND, UL_ROT, NATIVE>(A, c, W, M, Md, Mo);
  }

  trace_mark(A, 4);
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
  asm volatile("bar.sync 1, %0;" ::"n"(kFusedThreads) : "memory");  // the compute warps only (the producer group has left)
  if (!valid) {
    trace_mark(A, 5);
    return;
  }
  if (ND > 0 && !ul_thread) {
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
  if (has_ul && j == (ul_role ? ND : 0)) {  // one role of every site gathers the Hermitian matrix and writes the ultra-local loop
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
  trace_mark(A, 5);
}

// ---- host side: tiling and launch ----------------------------------------------------------------------------
static int fused_smem_limit_bytes() {
  // the opt-in limit is a per-device attribute: one cached value per device ordinal (one process may drive several GPUs)
  static int limit[64];
  static bool known[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 227 * 1024;
  if (!known[dev]) {
    int v = 0;
    limit[dev] = cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) == cudaSuccess ? v : 227 * 1024;
    known[dev] = true;
  }
  return limit[dev];
}

// Largest stage (in sites) over the CTAs of a launch, exact: the merged intervals of every run are evaluated once per
// (lattice, group, run, range) and remembered - 10^4 runs of a few dozen interval insertions each.
static int max_stage_sites(const FusedGroup &grp, const LatGeom &g, int run, int c_begin, int c_end, int align, bool *overflow) {
  static std::mutex mu;
  static std::map<std::vector<int>, std::pair<int, bool>> cache;
  std::vector<int> key = {g.L[0], g.L[1], g.L[2], g.L[3], run, c_begin, c_end, align, grp.nloops};
  for (int j = 0; j < grp.nloops; j++) {
    key.push_back(grp.loop[j].dir);
    key.push_back(grp.loop[j].sign);
    key.push_back(grp.loop[j].len);
  }
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(key);
  if (it == cache.end()) {
    int best = 0;
    bool ovf = false;
    StageMap m;
    for (int c0 = c_begin; c0 < c_end; c0 += run) {
      build_stage_map(m, grp, g, 1, c0, std::min(c0 + run, c_end), align);
      best = std::max(best, m.sites);
      ovf = ovf || m.overflow;
    }
    it = cache.emplace(key, std::make_pair(best, ovf)).first;
  }
  *overflow = it->second.second;
  return it->second.first;
}

// ---- stage maps of a launch, built once on the device ----------------------------------------------------------------------
// Thread 0 of a CTA needed ~7 us (3 % of a CTA's life at BASELINE configs[1], tools/fused_trace.py) to work out its stage map;
// the maps depend on (lattice, group geometry, range, run) only, so one small kernel writes them all to device memory the
// first time a launch shape is seen and every later CTA just copies its 1.8 KB.
__global__ void stage_maps_kernel(StageMap *out, FusedGroup grp, LatGeom g, int site_bytes, int c_begin, int c_end, int run, int align,
                                  int ncta) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= ncta) return;
  const int c0 = c_begin + b * run;
  build_stage_map(out[b], grp, g, site_bytes, c0, min(c0 + run, c_end), align);
}

namespace {
struct StageMapsEntry {
  StageMap *dev = nullptr;
  cudaEvent_t built = nullptr;
  bool ready = false;
  size_t bytes = 0;
};
struct StageMapsCache {
  std::mutex mu;
  std::map<std::vector<int>, StageMapsEntry> entries;
  size_t bytes = 0;
};
StageMapsCache &stage_maps_cache(int dev) {
  static StageMapsCache c[64];
  return c[(dev >= 0 && dev < 64) ? dev : 0];
}
}  // namespace

static int stage_maps_get(const StageMap **maps, const FusedGroup &grp, const LatGeom &g, int site_bytes, int c_begin, int c_end,
                          int run, int align, int ncta, cudaStream_t stream) {
  static_assert(sizeof(StageMap) % sizeof(int4) == 0, "StageMap is copied as int4");
  int dev = 0;
  MUGIQ_CUDA_CHECK(cudaGetDevice(&dev));
  StageMapsCache &c = stage_maps_cache(dev);
  std::vector<int> key = {g.L[0], g.L[1], g.L[2], g.L[3], site_bytes, run, c_begin, c_end, align, grp.nloops};
  for (int j = 0; j < grp.nloops; j++) {
    key.push_back(grp.loop[j].dir);
    key.push_back(grp.loop[j].sign);
    key.push_back(grp.loop[j].len);
  }
  std::lock_guard<std::mutex> lock(c.mu);
  auto it = c.entries.find(key);
  if (it == c.entries.end()) {
    const size_t bytes = (size_t)ncta * sizeof(StageMap);
    if (c.bytes + bytes > ((size_t)2 << 30)) {  // start over: nothing in flight may still read the old maps
      MUGIQ_CUDA_CHECK(cudaDeviceSynchronize());
      for (auto &e : c.entries) {
        cudaFree(e.second.dev);
        cudaEventDestroy(e.second.built);
      }
      c.entries.clear();
      c.bytes = 0;
    }
    StageMapsEntry e;
    e.bytes = bytes;
    MUGIQ_CUDA_CHECK(cudaMalloc((void **)&e.dev, bytes));
    MUGIQ_CUDA_CHECK(cudaEventCreateWithFlags(&e.built, cudaEventDisableTiming));
    stage_maps_kernel<<<(ncta + 63) / 64, 64, 0, stream>>>(e.dev, grp, g, site_bytes, c_begin, c_end, run, align, ncta);
    MUGIQ_LAUNCH_CHECK();
    MUGIQ_CUDA_CHECK(cudaEventRecord(e.built, stream));
    c.bytes += bytes;
    it = c.entries.emplace(key, e).first;
  }
  StageMapsEntry &e = it->second;
  if (!e.ready) {  // a launch on another stream must not overtake the builder
    if (cudaEventQuery(e.built) == cudaSuccess)
      e.ready = true;
    else
      MUGIQ_CUDA_CHECK(cudaStreamWaitEvent(stream, e.built, 0));
  }
  *maps = e.dev;
  return MUGIQ_B200_OK;
}

// Run length: the CTA's warps = (roles: threads working on one site) x (warps per role), a warp = 32 consecutive sites of
// one parity.  v1 (8 warps, one loop per thread): a group of 3 or 4 displaced loops gets runs of 32 sites per parity,
// 2 loops 64, 1 loop or the ultra-local loop alone 128; shorter if the shared-memory ring would otherwise have fewer than
// `want_stages` stages.  smem_avail: dynamic shared memory minus what is not ring (header, exchange buffer per unit).
static bool fused_choose_tiling(FusedTiling &tl, const FusedGroup &grp, const LatGeom &g, int precision, int c_begin, int c_end,
                                int warps, int roles, int smem_avail, int smem_per_unit, int want_stages, int align) {
  const int site = 24 * (int)prec_bytes(precision);
  int max_stages = 8;
  if (const char *e = getenv("MUGIQ_B200_FUSED_STAGES")) {
    const int want = atoi(e);
    if (want >= 2 && want < max_stages) max_stages = want;
  }
  FusedTiling best;
  best.nstages = 0;
  for (int units = (warps / roles) & ~1; units >= 2; units -= 2) {
    FusedTiling t;
    t.units = units;
    t.run = 16 * units;
    bool overflow = false;
    const int sites = max_stage_sites(grp, g, t.run, c_begin, c_end, align, &overflow);
    if (overflow) continue;
    t.stage_bytes = (sites * site + 127) / 128 * 128;
    t.nstages = std::min(max_stages, (smem_avail - t.units * smem_per_unit) / t.stage_bytes);
    if (t.nstages > best.nstages) best = t;
    if (best.nstages >= std::min(want_stages, max_stages)) break;
  }
  if (best.nstages < 2) return false;
  tl = best;
  return true;
}

static bool choose_tiling(FusedTiling &tl, const FusedGroup &grp, const LatGeom &g, int precision, int smem_limit, int c_begin,
                          int c_end, int native, bool with_ul) {
  // roles of the kernel (loop_fused_kernel): the ultra-local matrix has its own role beside 1..3 displaced loops
  const int nrole = (grp.nloops > 0 ? grp.nloops : 1) + ((with_ul && grp.nloops >= 1 && grp.nloops <= 3) ? 1 : 0);
  return fused_choose_tiling(tl, grp, g, precision, c_begin, c_end, kFusedComputeWarps, nrole, smem_limit - kSmemHeader,
                             32 * 16 * (int)prec_bytes(precision), 3, native ? kChunk : 1);
}

// Host-only self-check of the tiling (no GPU needed; exported as mugiq_b200_fused_tiling_check for the CPU tests): for
// every CTA of a launch, the stage map must be sorted, disjoint and within the sized stage, and every thread's own and
// neighbour site must lie in it.
int fused_tiling_check(const FusedGroup &grp, const LatGeom &g, int precision, int t_begin, int t_end, int native, bool with_ul,
                       long long out[8]) {
  FusedTiling tl;
  const int V3h = g.V3 / 2, c_begin = t_begin * V3h, c_end = t_end * V3h;
  if (native && g.volumeCB % kChunk)
    return set_error(MUGIQ_B200_EINVAL, "fused_tiling_check: QUDA-ordered eigenvectors need volumeCB to be a multiple of %d", kChunk);
  if (!choose_tiling(tl, grp, g, precision, 227 * 1024, c_begin, c_end, native, with_ul))
    return set_error(MUGIQ_B200_EINVAL, "fused_tiling_check: no tiling fits");
  const int site = 24 * (int)prec_bytes(precision);
  long long misses = 0, bad_maps = 0, stage_sites = 0, ctas = 0;
  int max_copies = 0;
  StageMap m;
  for (int c0 = c_begin; c0 < c_end; c0 += tl.run) {
    const int c1 = std::min(c0 + tl.run, c_end);
    build_stage_map(m, grp, g, site, c0, c1, native ? kChunk : 1);
    ctas++;
    stage_sites += m.sites;
    max_copies = std::max(max_copies, m.ncp);
    if (native) {  // copies: whole chunks, inside the stage and the lattice, covering every interval exactly once
      long long covered = 0;
      for (int i = 0; i < m.ncp; i++) {
        const int d = m.cp_goff16[i], nb = 1 << ((d >> 29) & 3), chunk = d & 0x0fffffff;
        if (m.cp_bytes[i] != nb * kChunk * site || (m.cp_soff[i] % (kChunk * site)) || (chunk + nb) * kChunk > g.volumeCB ||
            m.cp_soff[i] + m.cp_bytes[i] > tl.stage_bytes)
          bad_maps++;
        covered += nb * kChunk;
      }
      if (covered != m.sites) bad_maps++;
      for (int i = 0; i < m.n; i++)
        if ((m.lo[i] % kChunk) || (m.hi[i] % kChunk)) bad_maps++;
    }
    if (m.overflow || m.sites * site > tl.stage_bytes) bad_maps++;
    for (int i = 0; i + 1 < m.n; i++)
      if (m.par[i] > m.par[i + 1] || (m.par[i] == m.par[i + 1] && m.hi[i] >= m.lo[i + 1])) bad_maps++;
    for (int i = 0; i < m.n; i++)
      if (m.lo[i] < 0 || m.hi[i] > g.volumeCB || m.lo[i] >= m.hi[i] || (m.cp_bytes[i] & 15) || (m.cp_soff[i] & 15)) bad_maps++;
    for (int p = 0; p < 2; p++)
      for (int cb = c0; cb < c1; cb++) {
        if (stage_site(m, p, cb) < 0) misses++;
        for (int j = 0; j < grp.nloops; j++)
          if (stage_site(m, (p + grp.loop[j].len) & 1, neighbour_cb(g, grp.loop[j], p, cb)) < 0) misses++;
      }
  }
  out[0] = tl.run;
  out[1] = tl.units;
  out[2] = tl.nstages;
  out[3] = tl.stage_bytes;
  out[4] = max_copies;
  out[5] = ctas ? stage_sites / ctas : 0;  // mean sites staged per CTA and eigenvector
  out[6] = misses;
  out[7] = bad_maps;
  return MUGIQ_B200_OK;
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
    if (choose_tiling(tl, grp, g, precision, fused_smem_limit_bytes(), 0, g.volumeCB, 0, false)) return nl;
  }
  return -1;
}

template <typename F, int ND, int NATIVE>
static int launch_fused_nd(const FusedArgs<F> &args, size_t smem, int grid, cudaStream_t stream) {
  // the shared-memory opt-in is a per-device function attribute
  static bool attr_set[64];
  int dev = 0;
  MUGIQ_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    MUGIQ_CUDA_CHECK(cudaFuncSetAttribute(loop_fused_kernel<F, ND, NATIVE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          fused_smem_limit_bytes()));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  loop_fused_kernel<F, ND, NATIVE><<<grid, kFusedThreadsWS, smem, stream>>>(args);
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
  args.trace = g_fused_trace;
  args.trace_ctas = g_fused_trace_ctas;
  // time-slice range -> range of checkerboard indices (a time-slice is V3/2 consecutive sites of each parity)
  const int V3h = g.V3 / 2;
  args.c_begin = t_begin * V3h;
  args.c_end = t_end * V3h;
  if (!choose_tiling(args.tl, grp, g, precision, fused_smem_limit_bytes(), args.c_begin, args.c_end, vt.native, ul_off >= 0))
    return set_error(MUGIQ_B200_EINVAL, "loop_fused: no tiling fits %d loops on a %dx%dx%dx%d lattice", grp.nloops, g.L[0],
                     g.L[1], g.L[2], g.L[3]);
  const int grid = (args.c_end - args.c_begin + args.tl.run - 1) / args.tl.run;
  if (const int rc = stage_maps_get(&args.maps, grp, g, 24 * (int)sizeof(F), args.c_begin, args.c_end, args.tl.run,
                                    vt.native ? kChunk : 1, grid, stream))
    return rc;
  const double frac = (double)(t_end - t_begin) / (double)g.L[3];  // share of the lattice this launch computes
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
  if (vt.native) {  // eigenvectors in QUDA FLOAT2 order, staged as tensor boxes (vt.evec[] = their tensor maps)
    switch (grp.nloops) {
      case 0: return launch_fused_nd<F, 0, 1>(args, smem, grid, stream);
      case 1: return launch_fused_nd<F, 1, 1>(args, smem, grid, stream);
      case 2: return launch_fused_nd<F, 2, 1>(args, smem, grid, stream);
      case 3: return launch_fused_nd<F, 3, 1>(args, smem, grid, stream);
      default: return launch_fused_nd<F, 4, 1>(args, smem, grid, stream);
    }
  }
  switch (grp.nloops) {
    case 0: return launch_fused_nd<F, 0, 0>(args, smem, grid, stream);
    case 1: return launch_fused_nd<F, 1, 0>(args, smem, grid, stream);
    case 2: return launch_fused_nd<F, 2, 0>(args, smem, grid, stream);
    case 3: return launch_fused_nd<F, 3, 0>(args, smem, grid, stream);
    default: return launch_fused_nd<F, 4, 0>(args, smem, grid, stream);
  }
}

// ---- tensor maps of QUDA-ordered eigenvectors ------------------------------------------------------------------------------
// A FLOAT2 field [parity][component][x_cb] is a 4-D tensor (16 reals = 8 sites, 12 components, volumeCB/8 chunks, 2 parities):
// a box of (16, 12, NB, 1) lands in shared memory as NB x [component][8 sites], the layout the kernel's stage map expects.
// Three maps per eigenvector (NB = 1, 2, 4), encoded on the host (cuTensorMapEncodeTiled, bound at run time: the library does
// not link libcuda) and kept in device memory; the encoding depends on (address, lattice, precision) only, so it is cached.
namespace {
typedef CUresult (*TmapEncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
struct TmapCache {
  std::mutex mu;
  std::map<std::tuple<const void *, int, int>, int> slot;  // (field, volumeCB, precision) -> index
  char *dev = nullptr;
  int used = 0;
  static constexpr int kCapacity = 16384;  // eigenvectors
  static constexpr int kBytes = 3 * 128;   // three CUtensorMap objects
};
TmapCache &tmap_cache(int dev) {
  static TmapCache c[64];
  return c[(dev >= 0 && dev < 64) ? dev : 0];
}
TmapEncodeFn tmap_encoder() {
  static TmapEncodeFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<TmapEncodeFn>(p);
  });
  return fn;
}
}  // namespace

int fused_native_tmaps(const void **tmap_d, const void *const *evec_d, int nvec, const LatGeom &g, int precision,
                       cudaStream_t stream) {
  if (g.volumeCB % kChunk)
    return set_error(MUGIQ_B200_EINVAL, "loop plan: QUDA-ordered eigenvectors need volumeCB = %d to be a multiple of %d", g.volumeCB, kChunk);
  TmapEncodeFn encode = tmap_encoder();
  if (!encode) return set_error(MUGIQ_B200_ESTATE, "loop plan: cuTensorMapEncodeTiled is not available from this driver");
  int dev = 0;
  MUGIQ_CUDA_CHECK(cudaGetDevice(&dev));
  TmapCache &c = tmap_cache(dev);
  std::lock_guard<std::mutex> lock(c.mu);
  if (!c.dev) MUGIQ_CUDA_CHECK(cudaMalloc((void **)&c.dev, (size_t)TmapCache::kCapacity * TmapCache::kBytes));
  const size_t pb = prec_bytes(precision);
  for (int i = 0; i < nvec; i++) {
    const auto key = std::make_tuple(evec_d[i], g.volumeCB, precision);
    auto it = c.slot.find(key);
    if (it == c.slot.end()) {
      if (c.used == TmapCache::kCapacity) {  // start over: nothing in flight may still read the old entries
        MUGIQ_CUDA_CHECK(cudaDeviceSynchronize());
        c.slot.clear();
        c.used = 0;
      }
      if ((uintptr_t)evec_d[i] & 15) return set_error(MUGIQ_B200_EINVAL, "loop plan: eigenvector %d is not 16-byte aligned", i);
      alignas(64) CUtensorMap tm[3];
      const cuuint64_t dim[4] = {16, 12, (cuuint64_t)g.volumeCB / kChunk, 2};
      const cuuint64_t str[3] = {(cuuint64_t)g.volumeCB * 2 * pb, (cuuint64_t)kChunk * 2 * pb, (cuuint64_t)12 * g.volumeCB * 2 * pb};
      const cuuint32_t es[4] = {1, 1, 1, 1};
      for (int k = 0; k < 3; k++) {
        const cuuint32_t box[4] = {16, 12, (cuuint32_t)(1 << k), 1};
        const CUresult r = encode(&tm[k], pb == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4,
                                  const_cast<void *>(evec_d[i]), dim, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return set_error(MUGIQ_B200_ECUDA, "loop plan: cuTensorMapEncodeTiled failed with %d", (int)r);
      }
      static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap size");
      MUGIQ_CUDA_CHECK(cudaMemcpyAsync(c.dev + (size_t)c.used * TmapCache::kBytes, tm, TmapCache::kBytes, cudaMemcpyHostToDevice, stream));
      it = c.slot.emplace(key, c.used++).first;
    }
    tmap_d[i] = c.dev + (size_t)it->second * TmapCache::kBytes;
  }
  return MUGIQ_B200_OK;
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
