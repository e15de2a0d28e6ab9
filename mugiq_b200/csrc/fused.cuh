// fused.cuh — interface of the fused displace+contract kernel (fused_kernel.cu) and the gauge-only helper
// kernels (wilson.cu) used by the loop schedule in loop_fused.cu.
#pragma once
#include "kernels.cuh"

namespace mugiq_b200 {

constexpr int kFusedMaxVec = 256;    // eigenvectors per launch (pointer + 1/sigma table travels as a kernel parameter)
constexpr int kFusedMaxLoops = 4;    // displaced loops per launch group (the ultra-local loop rides along for free)
constexpr int kFusedMaxIv = 64;      // merged intervals (= bulk copies) of one eigenvector stage
// 8 warps = 2 per SM sub-partition: each sub-partition owns 16384 registers, so 2 warps may use up to 255 registers
// per thread (a third warp would cap the kernel at 168 and spill the 4x4 spin matrix + link + operands)
constexpr int kFusedComputeWarps = 8;
constexpr int kFusedThreads = kFusedComputeWarps * 32;
// warp-specialised form: one more warp group (4 warps, setmaxnreg works per warp group) whose first warp is the TMA producer
constexpr int kFusedThreadsWS = kFusedThreads + 128;

// One displaced loop of a launch group: the kernel computes  M(x) += (1/sigma) conj(v(x)) (x) [W(x) v(x + sign*len*dir)]
// where W is the Wilson line of `len` links (already daggered / shifted for minus), stored like one direction
// of the gauge field: [x_eo][row][col] complex.
struct FusedLoop {
  const void *W;
  long long out_off;  // complex offset of this loop's 16*V4 block in dataPos
  int dir, sign, len;
  int pad_;
};

struct FusedGroup {
  FusedLoop loop[kFusedMaxLoops];
  int nloops;
};

// Tile geometry chosen on the host (same for every CTA): a CTA works on `run` consecutive checkerboard sites of both
// parities; what a stage holds is worked out per CTA (fused_kernel.cu, StageMap).
struct FusedTiling {
  int run;          // consecutive checkerboard sites per parity and CTA (multiple of 32)
  int units;        // warps per loop of the group: 2 * run / 32
  int stage_bytes;  // largest stage over the CTAs of the launch (merged intervals), rounded up to 128 B
  int nstages;
};

struct FusedVecTable {
  const void *evec[kFusedMaxVec];  // site-major fields; native != 0: each eigenvector's tensor maps (fused_native_tmaps)
  double inv_sigma[kFusedMaxVec];
  int nvec;
  int native;  // 0: canonical site-major fields; MUGIQ_B200_ORDER_FLOAT2: QUDA FLOAT2 fields staged as tensor boxes
};

// maximum number of displaced loops per group for this lattice / precision (warp and shared-memory budget)
int fused_max_loops_per_group(const LatGeom &g, int precision);
// Launches one group (0..kFusedMaxLoops displaced loops, plus the ultra-local loop if ul_off >= 0) for up to kFusedMaxVec
// eigenvectors over the sites of the time-slices [t_begin, t_end) (the whole lattice for 0, Lt): a rank of a lattice-T split computes its interior only and merely READS the halo slices.
int fused_group_launch(void *dataPos_d, const FusedGroup &grp, long long ul_off, const FusedVecTable &vt, int accumulate,
                       const LatGeom &g, int precision, cudaStream_t stream, int t_begin = 0, int t_end = -1);

// per-CTA timeline of the following launches (diagnostics; nullptr switches it off)
void fused_set_trace(long long *trace_d, long long capacity_ctas);

// host-only self-check of the tiling of one launch group (fused_kernel.cu); out = {run, units, nstages, stage_bytes,
// max copies per stage, mean sites staged per CTA, sites not found in their stage, malformed stage maps}
int fused_tiling_check(const FusedGroup &grp, const LatGeom &g, int precision, int t_begin, int t_end, int native, bool with_ul,
                       long long out[8]);
// device addresses of the tensor maps (three per eigenvector: boxes of 1, 2, 4 chunks of 8 sites) of QUDA FLOAT2 fields,
// encoded on first use and cached per (address, lattice, precision); new entries are uploaded on `stream`
int fused_native_tmaps(const void **tmap_d, const void *const *evec_d, int nvec, const LatGeom &g, int precision,
                       cudaStream_t stream);

// ---- gauge-only helpers (wilson.cu) ---------------------------------------------------------------------
// Wout(x) = Win(x) * U_dir(x + shift*dir)       (extends a plus-direction Wilson line by one link)
int wilson_extend(void *Wout_d, const void *Win_d, const void *gauge_d, int dir, int shift, const LatGeom &g,
                  int precision, cudaStream_t stream);
// Wminus(x) = [Wplus(x - len*dir)]^dagger
int wilson_minus_from_plus(void *Wminus_d, const void *Wplus_d, int dir, int len, const LatGeom &g, int precision,
                           cudaStream_t stream);
// loopMinus[G][x] (+)= herm(G) * conj(loopPlus[G][x - len*dir])   (Gamma_G^dagger = herm(G) Gamma_G), for up to
// kMinusBatch (plus slot, minus slot) pairs of the loop buffer in one launch
constexpr int kMinusBatch = 32;
struct MinusBatch {
  struct Item {
    int dst, src, dir, len;  // loop slots of dataPos
  };
  Item item[kMinusBatch];
  int n;
};
int loop_minus_from_plus(void *dataPos_d, const MinusBatch &batch, int accumulate, const LatGeom &g, int precision,
                         cudaStream_t stream);

}  // namespace mugiq_b200
