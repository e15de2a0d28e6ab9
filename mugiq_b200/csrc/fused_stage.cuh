// fused_stage.cuh — building blocks of the fused displace+contract kernel (fused_kernel.cu): the stage map of a tile
// (host + device), the predicated mbarrier / bulk-TMA primitives of the eigenvector loop, the spin-rotation helpers of the
// epilogue.
#pragma once
#include "fused.cuh"

namespace mugiq_b200 {

// ---- PTX helpers: mbarrier + bulk TMA -------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

// ---- stage map of a tile: the merged intervals of checkerboard-index space one eigenvector stage holds ------------
// shared by host (sizing) and device (thread 0 of every CTA builds its own)
constexpr int kMaxIv = kFusedMaxIv;
constexpr int kChunk = 8;  // sites per chunk of a natively ordered (QUDA FLOAT2) stage
struct StageMap {
  int n;         // merged intervals
  int ncp;       // bulk copies per stage (site-major: one per interval; FLOAT2: tensor boxes of 4, 2 or 1 chunks)
  int sites;     // sites per stage
  int overflow;  // more than kMaxIv intervals or copies (the host checks this before launching)
  int par[kMaxIv], lo[kMaxIv], hi[kMaxIv];  // interval [lo, hi) of checkerboard indices of parity par, sorted by (par, lo)
  int soff[kMaxIv];                          // first site of the interval inside the stage
  // the bulk copies, precomputed for the issuing lanes
  int cp_soff[kMaxIv];    // byte offset inside the stage
  int cp_goff16[kMaxIv];  // site-major: offset inside the eigenvector in units of 16 B;
                          // FLOAT2: first chunk | parity << 28 | box code << 29 (0, 1, 2: boxes of 1, 2, 4 chunks)
  int cp_bytes[kMaxIv];
};

__host__ __device__ inline int wrap(int a, int n) {
  a %= n;
  return a < 0 ? a + n : a;
}

// insert [lo, hi) into the sorted list of disjoint intervals m.lo/m.hi[0..m.n), merging what overlaps or adjoins
__host__ __device__ inline void iv_insert(StageMap &m, int lo, int hi) {
  if (lo >= hi) return;
  int i = 0;
  while (i < m.n && m.hi[i] < lo) i++;
  if (i < m.n && m.lo[i] <= hi) {  // touches interval i: grow it, swallow the followers it reaches
    if (lo < m.lo[i]) m.lo[i] = lo;
    if (hi > m.hi[i]) m.hi[i] = hi;
    int k = i + 1;
    while (k < m.n && m.lo[k] <= m.hi[i]) {
      if (m.hi[k] > m.hi[i]) m.hi[i] = m.hi[k];
      k++;
    }
    if (k > i + 1) {
      for (int d = i + 1, s = k; s < m.n; d++, s++) {
        m.lo[d] = m.lo[s];
        m.hi[d] = m.hi[s];
      }
      m.n -= k - (i + 1);
    }
    return;
  }
  if (2 * (m.n + 1) > kMaxIv) {  // the list is replicated for the second parity at the end
    m.overflow = 1;
    return;
  }
  for (int d = m.n; d > i; d--) {
    m.lo[d] = m.lo[d - 1];
    m.hi[d] = m.hi[d - 1];
  }
  m.lo[i] = lo;
  m.hi[i] = hi;
  m.n++;
}

// Stage of the run [c0, c1) (both parities) for the loops of grp: the run itself plus, per loop, the image of every row
// piece of the run.  A y/z/t shift maps the piece [a, b) of row r to the same positions of the shifted row; an x shift
// of length k keeps the row and moves the half-row index by at most ceil(k/2) (periodic inside the row).  Both
// parities are staged for every image (a site's neighbour has parity p ^ (k & 1) and both own parities are in the tile),
// so the interval list is built once and replicated.  Images of consecutive row pieces usually adjoin: they are joined
// before they are inserted.
//
// align = 1: site-major eigenvectors, every interval is one linear bulk copy.  align = kChunk (QUDA FLOAT2 order,
// [parity][spin*3+colour][x_cb]): intervals are widened to whole chunks of 8 sites, the stage holds them as
// [chunk][component][8 sites] (1536 B per chunk in FP64) and they arrive as 4-D tensor boxes of 4, 2 or 1 chunks.
__host__ __device__ inline void iv_insert_aligned(StageMap &m, int lo, int hi, int align, int limit) {
  if (lo >= hi) return;
  if (align > 1) {
    lo = lo / align * align;
    hi = (hi + align - 1) / align * align;
    if (hi > limit) hi = limit;
  }
  iv_insert(m, lo, hi);
}

__host__ __device__ inline void build_stage_map(StageMap &m, const FusedGroup &grp, const LatGeom &g, int site_bytes, int c0,
                                                int c1, int align = 1) {
  m.n = 0;
  m.overflow = 0;
  const int Lh = g.Lh;
  const int vcb = g.volumeCB;
  iv_insert_aligned(m, c0, c1, align, vcb);
  for (int j = 0; j < grp.nloops; j++) {
    const FusedLoop &lp = grp.loop[j];
    const int sh = lp.sign * lp.len;
    int plo = 0, phi = 0;  // pending image interval
    for (int c = c0; c < c1;) {
      const int row = c / Lh, a = c - row * Lh;
      int b = a + (c1 - c);
      if (b > Lh) b = Lh;
      const int base = row * Lh;
      int lo, hi;
      if (lp.dir == 0) {
        const int h = (lp.len + 1) >> 1;
        lo = a - h;
        hi = b + h;
        if (hi - lo >= Lh) {
          lo = 0;
          hi = Lh;
        }
        if (lo < 0) {
          iv_insert_aligned(m, base + lo + Lh, base + Lh, align, vcb);
          lo = 0;
        }
        if (hi > Lh) {
          iv_insert_aligned(m, base, base + hi - Lh, align, vcb);
          hi = Lh;
        }
        lo += base;
        hi += base;
      } else {
        int y = row % g.L[1], z = (row / g.L[1]) % g.L[2], t = row / (g.L[1] * g.L[2]);
        if (lp.dir == 1) y = wrap(y + sh, g.L[1]);
        if (lp.dir == 2) z = wrap(z + sh, g.L[2]);
        if (lp.dir == 3) t = wrap(t + sh, g.L[3]);
        const int nb = (y + g.L[1] * (z + g.L[2] * t)) * Lh;
        lo = nb + a;
        hi = nb + b;
      }
      if (lo <= phi && plo <= hi && phi > plo) {  // adjoins or overlaps the pending image: join
        if (lo < plo) plo = lo;
        if (hi > phi) phi = hi;
      } else {
        iv_insert_aligned(m, plo, phi, align, vcb);
        plo = lo;
        phi = hi;
      }
      c += b - a;
    }
    iv_insert_aligned(m, plo, phi, align, vcb);
  }
  // parity 0 block, then the same intervals for parity 1
  const int n = m.n;
  int off = 0;
  for (int p = 0; p < 2; p++)
    for (int i = 0; i < n; i++) {
      const int k = p * n + i;
      m.par[k] = p;
      m.lo[k] = m.lo[i];
      m.hi[k] = m.hi[i];
      m.soff[k] = off;
      off += m.hi[i] - m.lo[i];
    }
  m.n = 2 * n;
  m.sites = off;
  if (align == 1) {
    for (int k = 0; k < m.n; k++) {
      m.cp_soff[k] = m.soff[k] * site_bytes;
      m.cp_goff16[k] = (int)((((long long)m.par[k] * g.volumeCB + m.lo[k]) * site_bytes) >> 4);
      m.cp_bytes[k] = (m.hi[k] - m.lo[k]) * site_bytes;
    }
    m.ncp = m.n;
  } else {
    int nc = 0;
    for (int k = 0; k < m.n; k++) {
      int chunk = m.lo[k] / kChunk, left = (m.hi[k] - m.lo[k]) / kChunk, pos = m.soff[k];
      while (left > 0) {
        const int code = left >= 4 ? 2 : (left >= 2 ? 1 : 0), nb = 1 << code;
        if (nc == kMaxIv) {
          m.overflow = 1;
          break;
        }
        m.cp_soff[nc] = pos * site_bytes;  // a chunk holds kChunk sites of all 12 components: kChunk * site_bytes
        m.cp_goff16[nc] = chunk | (m.par[k] << 28) | (code << 29);
        m.cp_bytes[nc] = nb * kChunk * site_bytes;
        nc++;
        chunk += nb;
        left -= nb;
        pos += nb * kChunk;
      }
    }
    m.ncp = nc;
  }
}

// byte offset, inside a stage, of component 0 of the site at stage position pos
template <int NATIVE> __host__ __device__ inline int stage_site_bytes(int pos, int site_bytes) {
  // FLOAT2 stage: [chunk][component][8 sites]; a complex number is site_bytes / 12 bytes
  return NATIVE ? (pos / kChunk) * (kChunk * site_bytes) + (pos % kChunk) * (site_bytes / 12) : pos * site_bytes;
}

// position (in sites) of checkerboard site cb of parity par inside the stage; -1 if the stage does not hold it
__host__ __device__ inline int stage_site(const StageMap &m, int par, int cb) {
  const int n = m.n >> 1;
  for (int i = par * n; i < (par + 1) * n; i++)
    if (m.lo[i] <= cb && cb < m.hi[i]) return m.soff[i] + cb - m.lo[i];
  return -1;
}

// checkerboard index of the site x + sign*len*dir a loop reads for the own site (parity p, checkerboard index cb);
// its parity is p ^ (len & 1).  Neighbour selection of lib/mugiq_displace_kernels.cu:116-151 for a hop of `len` links.
__host__ __device__ inline int neighbour_cb(const LatGeom &g, const FusedLoop &lp, int p, int cb) {
  const int Lh = g.Lh;
  const int row = cb / Lh, sx = cb - row * Lh;
  const int ya = row % g.L[1], za = (row / g.L[1]) % g.L[2], ta = row / (g.L[1] * g.L[2]);
  const int sh = lp.sign * lp.len;
  if (lp.dir == 0) {
    const int x = 2 * sx + ((ya + za + ta + p) & 1);
    return row * Lh + (wrap(x + sh, g.L[0]) >> 1);
  }
  int yn = ya, zn = za, tn = ta;
  if (lp.dir == 1) yn = wrap(ya + sh, g.L[1]);
  if (lp.dir == 2) zn = wrap(za + sh, g.L[2]);
  if (lp.dir == 3) tn = wrap(ta + sh, g.L[3]);
  return (yn + g.L[1] * (zn + g.L[2] * tn)) * Lh + sx;
}

template <typename F> struct FusedArgs {
  LatGeom g;
  FusedTiling tl;
  FusedGroup grp;      // displaced loops only (0..kFusedMaxLoops)
  FusedVecTable vt;
  F *dataPos;
  long long ul_off;    // complex offset of the ultra-local loop's block in dataPos, < 0: not in this launch
  int accumulate;
  const StageMap *maps;  // stage map of every CTA of the launch (device memory, stage_maps_kernel)
  long long *trace;    // per-CTA timeline (mugiq_b200_prof_fused_trace), normally NULL
  int trace_ctas;
  int c_begin, c_end;  // checkerboard-index range [c_begin, c_end) of both parities this launch computes: the time-slices
                       // [t_begin, t_end) of a lattice-T split slab, or the whole lattice
};

constexpr int kSmemHeader = 3072;  // barriers + stage map

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
  const StageMap *st;
  int S, stage_bytes, nvec, nActive, warp, lane;
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

// ---- hot-loop primitives on 32-bit shared addresses: predicated mbarrier / TMA instructions instead of `if (lane == 0)`
// blocks, so that the loops stay branch-free.
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
// 4-D tensor box of a QUDA FLOAT2 eigenvector: coordinates (0, 0, chunk, parity) of the tensor (16 reals, 12 components,
// volumeCB / 8 chunks, 2 parities); lands as [chunk][component][8 sites]
__device__ __forceinline__ void tma_tensor4_g2s(uint32_t dst, const void *tmap, int chunk, int parity, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
      "l"(tmap), "r"(0), "r"(0), "r"(chunk), "r"(parity), "r"(bar)
      : "memory");
}
}  // namespace mugiq_b200
