// momproj_pos.cu — stages 3+4 fused: momentum projection straight from the position-space loop buffer.
//
//   dataMom[t + Lt*((15-G) + 16*iL) + Lt*nData*im] = sum_{v3} sign[G] * dataPos[x_eo(v3,t) + V4*(G + 16*iL)] * phase(v3, im)
//
// Replaces convertIdxOrder_mapGamma (/root/reference/lib/contract_wrappers.cu:133-156, lib/mugiq_util_kernels.cu:59-99)
// AND the cuBLAS Zgemm/Cgemm that follows it (lib/loop_mugiq.cpp:364-377): the reorder pass exists only to bring
// dataPos into the operand order of a column-major GEMM; it moves 2 x 16*V4*nLoop complex through HBM (5.6 GB each
// way at 24^3x48 with 33 loops) and needs a second buffer of that size.  Here the GEMM reads dataPos in place:
//  * for fixed (G, iL, t, parity) the V3/2 sites of a time-slice are CONTIGUOUS in dataPos (even/odd order), so that
//    run is the K dimension of a row of the A operand, read with 128-bit loads, every byte of every sector used;
//  * the spatial index behind position i of such a run depends on (t + parity) & 1 only (QUDA getCoords: x = 2*(i % Lh)
//    + ((y + z + t + parity) & 1)), so the phase matrix is stored once in the same even/odd order for both values
//    of s = (t + parity) & 1:  phase_eo[s][im][i]   (mugiq_b200_phase_matrix_eo);
//  * the gamma map (G -> 15-G, sign) is applied when the result is written.
// FP64: DMMA m8n8k4 tiles, complex product embedded in real tiles exactly as in momproj.cu (C^T = P_emb^T A^T); a warp
// owns the 16 gammas of one (t, iL) - they share t, hence s and the phase operand - for all momenta over a K chunk.
// FP32: SIMT.  Split-K partial sums are reduced in a fixed order (deterministic).
#include <algorithm>

#include "kernels.cuh"

namespace mugiq_b200 {

struct PosGeom {
  int Lt, V3h, nLoop, N;
  long long Vh, V4;
  long long M;  // Lt * 16 * nLoop
  int kchunk, nchunk;  // chunk of the V3/2 run handled by one task (multiple of kPosKT); chunks per run
  // DMMA kernel: row groups (t, iL) are binned by class c = 2*parity + s, s = (t + parity) & 1, so that the warps of a
  // CTA share one phase operand; class c owns CTAs [blk0[c], blk0[c+1]) x nchunk
  int nT[2];    // time-slices with t & 1 == 0 / 1
  int blk0[5];
};

// GammaMap as a bit mask / closed form, for kernels that index it with a runtime G (a constexpr table indexed at run
// time lands in local memory); verified against the literal tables of common.cuh at compile time.
constexpr unsigned kMapMinusMask = (1u << 3) | (1u << 6) | (1u << 9) | (1u << 11) | (1u << 12) | (1u << 14);
constexpr bool map_mask_matches_tables() {
  constexpr GammaTables t = gamma_tables();
  for (int G = 0; G < 16; G++)
    if (t.map_index[G] != 15 - G || (t.map_sign[G] < 0) != (((kMapMinusMask >> G) & 1) != 0)) return false;
  return true;
}
static_assert(map_mask_matches_tables(), "gamma map mask out of sync with gamma_tables()");

constexpr int kPosKT = 16;     // sites of the K run per shared-memory phase tile (4 DMMA k-steps)
constexpr int kPosWarps = 4;   // warps (= row groups of 16 gammas) per CTA

__device__ __forceinline__ void dmma_m8n8k4_pos(double &c0, double &c1, const double a, const double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

// CTA = kPosWarps row groups (t, iL) of one class (same parity, same s => same phase rows) x one K chunk; warp = the 16
// gammas of one row group for NT*4 momenta.  The phase operand of a K tile is staged in shared memory once per CTA,
// already in the real embedding the A fragment needs ((re, -im) for the real rows of the result, (im, re) for the
// imaginary ones: one conflict-free 128-bit read per 4 DMMAs, no selects); the dataPos operand has no reuse and goes
// straight from global memory into the B fragments, one K tile ahead of the tensor pipe.
template <int NT>
__global__ void __launch_bounds__(kPosWarps * 32, (NT <= 4 ? 3 : 2))
momproj_pos_dmma_kernel(double *__restrict__ partial, const double *__restrict__ pos, const double *__restrict__ P,
                        const PosGeom pg, const int n0) {
  constexpr int kThreads = kPosWarps * 32;
  constexpr int kElems = NT * 4 * kPosKT;                  // complex phases per tile
  constexpr int kPer = (kElems + kThreads - 1) / kThreads;  // per thread
  __shared__ __align__(128) double ph_s[2][NT * 4][kPosKT][4];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int b = blockIdx.x;
  const int chunk = b % pg.nchunk;
  b /= pg.nchunk;
  const int c = (b >= pg.blk0[1]) + (b >= pg.blk0[2]) + (b >= pg.blk0[3]);
  const int p = c >> 1, s = c & 1, tbit = s ^ p;
  const int nT = tbit ? pg.nT[1] : pg.nT[0];
  const int cb0 = c == 0 ? pg.blk0[0] : c == 1 ? pg.blk0[1] : c == 2 ? pg.blk0[2] : pg.blk0[3];
  const int r = (b - cb0) * kPosWarps + warp;
  const bool active = r < nT * pg.nLoop;  // warp-uniform
  const int iL = active ? r / nT : 0;
  const int t = active ? 2 * (r - iL * nT) + tbit : tbit;
  const int gi = lane >> 2, j = lane & 3, comp = gi & 1;
  const int tile0 = chunk * (pg.kchunk / kPosKT);
  const int tile1 = min((pg.V3h + kPosKT - 1) / kPosKT, tile0 + pg.kchunk / kPosKT);

  double acc[2][NT][2];
#pragma unroll
  for (int mt = 0; mt < 2; mt++)
#pragma unroll
    for (int nt = 0; nt < NT; nt++) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;

  // row of gamma G = mt*8 + gi: the sites of (t, parity) start at p*Vh + t*V3h
  const double2 *arow[2];
#pragma unroll
  for (int mt = 0; mt < 2; mt++)
    arow[mt] = reinterpret_cast<const double2 *>(pos) + ((long long)p * pg.Vh + (long long)t * pg.V3h) +
               pg.V4 * ((mt * 8 + gi) + 16LL * iL);
  const double2 *P2 = reinterpret_cast<const double2 *>(P) + (long long)s * pg.N * pg.V3h;

  double2 phr[kPer];
  auto ph_load = [&](int tile) {
#pragma unroll
    for (int i = 0; i < kPer; i++) {
      const int e = tid + i * kThreads;
      const int n = n0 + e / kPosKT, k = tile * kPosKT + (e % kPosKT);
      phr[i] = (e < kElems && n < pg.N && k < pg.V3h) ? __ldg(P2 + (long long)n * pg.V3h + k) : make_double2(0.0, 0.0);
    }
  };
  auto ph_store = [&](int buf) {
#pragma unroll
    for (int i = 0; i < kPer; i++) {
      const int e = tid + i * kThreads;
      if (e < kElems) {
        double2 *d = reinterpret_cast<double2 *>(&ph_s[buf][e / kPosKT][e % kPosKT][0]);
        d[0] = make_double2(phr[i].x, -phr[i].y);
        d[1] = make_double2(phr[i].y, phr[i].x);
      }
    }
  };
  double2 aT[4][2], aN[4][2];
  auto a_load = [&](double2 (&a)[4][2], int tile) {
#pragma unroll
    for (int ks = 0; ks < 4; ks++) {
      const int k = tile * kPosKT + ks * 4 + j;
#pragma unroll
      for (int mt = 0; mt < 2; mt++) a[ks][mt] = (active && k < pg.V3h) ? __ldg(arow[mt] + k) : make_double2(0.0, 0.0);
    }
  };

  // the dataPos operand runs kAhead tiles ahead of the tensor pipe: two for few momenta (HBM bound, registers to spare:
  // more bytes in flight per warp), one for many (the accumulators need the registers)
  constexpr int kAhead = NT <= 4 ? 2 : 1;
  double2 aNN[4][2];
  ph_load(tile0);
  a_load(aT, tile0);
  if (kAhead == 2 && tile0 + 1 < tile1) a_load(aN, tile0 + 1);
  ph_store(0);
  __syncthreads();
  for (int tile = tile0; tile < tile1; tile++) {
    const int buf = (tile - tile0) & 1;
    const bool more = tile + 1 < tile1;
    if (more) ph_load(tile + 1);
    if (kAhead == 2) {
      if (tile + 2 < tile1) a_load(aNN, tile + 2);
    } else if (more) {
      a_load(aN, tile + 1);
    }
    if (active) {
#pragma unroll
      for (int ks = 0; ks < 4; ks++)
#pragma unroll
        for (int nt = 0; nt < NT; nt++) {
          const double2 a = *reinterpret_cast<const double2 *>(&ph_s[buf][nt * 4 + (gi >> 1)][ks * 4 + j][comp * 2]);
          dmma_m8n8k4_pos(acc[0][nt][0], acc[0][nt][1], a.x, aT[ks][0].x);
          dmma_m8n8k4_pos(acc[1][nt][0], acc[1][nt][1], a.x, aT[ks][1].x);
          dmma_m8n8k4_pos(acc[0][nt][0], acc[0][nt][1], a.y, aT[ks][0].y);
          dmma_m8n8k4_pos(acc[1][nt][0], acc[1][nt][1], a.y, aT[ks][1].y);
        }
    }
    if (more) ph_store(buf ^ 1);
    __syncthreads();
#pragma unroll
    for (int ks = 0; ks < 4; ks++)
#pragma unroll
      for (int mt = 0; mt < 2; mt++) {
        aT[ks][mt] = aN[ks][mt];
        if (kAhead == 2) aN[ks][mt] = aNN[ks][mt];
      }
  }
  if (!active) return;
  // c-fragment: row gi -> (n, comp); columns 2j, 2j+1 -> gamma G = mt*8 + 2j + e.  Written under the mapped index.
  double *out = partial + 2 * (size_t)(p * pg.nchunk + chunk) * (size_t)pg.M * pg.N;
#pragma unroll
  for (int mt = 0; mt < 2; mt++)
#pragma unroll
    for (int nt = 0; nt < NT; nt++) {
      const int n = n0 + nt * 4 + (gi >> 1);
      if (n < pg.N) {
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int G = mt * 8 + 2 * j + e;  // runtime index: use the bit-mask form of the map tables (checked below)
          const long long m = t + (long long)pg.Lt * ((15 - G) + 16 * iL);
          out[2 * (m + pg.M * n) + comp] = ((kMapMinusMask >> G) & 1) ? -acc[mt][nt][e] : acc[mt][nt][e];
        }
      }
    }
}

// SIMT version (FP32; the reference has cublasCgemm, lib/loop_mugiq.cpp:371-377): block = (t, iL, parity, chunk).  The
// threads stride the K chunk (coalesced); each holds the 16 gammas x 4 momenta of its sites in registers, so the dataPos
// operand is read once per group of four momenta; the 128 partial results of a group are reduced across lanes by shuffles
// and across the four warps by 128 threads in parallel.
template <typename F>
__global__ void __launch_bounds__(128)
momproj_pos_simt_kernel(F *__restrict__ partial, const F *__restrict__ pos, const F *__restrict__ P, const PosGeom pg) {
  constexpr GammaTables gt = gamma_tables();
  constexpr int NT = 4;
  __shared__ F red[4][16 * NT * 2];
  const long long task = blockIdx.x;
  const int chunk = (int)(task % pg.nchunk);
  const int p = (int)((task / pg.nchunk) & 1);
  const long long pair = task / (2 * pg.nchunk);
  const int t = (int)(pair % pg.Lt), iL = (int)(pair / pg.Lt);
  const int s = (t + p) & 1;
  const int kbeg = chunk * pg.kchunk, kend = min(pg.V3h, kbeg + pg.kchunk);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  F *out = partial + 2 * (size_t)(p * pg.nchunk + chunk) * (size_t)pg.M * pg.N;
  const F *arow0 = pos + 2 * (((long long)p * pg.Vh + (long long)t * pg.V3h) + pg.V4 * (16LL * iL));
  for (int n0 = 0; n0 < pg.N; n0 += NT) {
    Cplx<F> acc[16][NT];
#pragma unroll
    for (int G = 0; G < 16; G++)
#pragma unroll
      for (int j = 0; j < NT; j++) acc[G][j] = make_c<F>(0, 0);
    for (int k = kbeg + threadIdx.x; k < kend; k += blockDim.x) {
      Cplx<F> ph[NT];
#pragma unroll
      for (int j = 0; j < NT; j++)
        ph[j] = n0 + j < pg.N ? ldg_c<F>(P + 2 * (((long long)s * pg.N + n0 + j) * pg.V3h + k)) : make_c<F>(0, 0);
#pragma unroll
      for (int G = 0; G < 16; G++) {
        const Cplx<F> a = ldg_c<F>(arow0 + 2 * (pg.V4 * (long long)G + k));
#pragma unroll
        for (int j = 0; j < NT; j++) cmac(acc[G][j], a, ph[j]);
      }
    }
#pragma unroll
    for (int G = 0; G < 16; G++)
#pragma unroll
      for (int j = 0; j < NT; j++) {
        Cplx<F> v = acc[G][j];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          v.re += __shfl_xor_sync(0xffffffffu, v.re, off);
          v.im += __shfl_xor_sync(0xffffffffu, v.im, off);
        }
        if (lane == 0) {
          red[warp][(G * NT + j) * 2] = v.re;
          red[warp][(G * NT + j) * 2 + 1] = v.im;
        }
      }
    __syncthreads();
    {  // one output number per thread: (G, j, re/im) summed over the four warps, gamma map applied
      const int o = threadIdx.x, G = o / (NT * 2), j = (o >> 1) % NT, comp = o & 1;
      if (n0 + j < pg.N) {
        const F v = red[0][o] + red[1][o] + red[2][o] + red[3][o];
        const long long m = t + (long long)pg.Lt * ((15 - G) + 16 * iL);
        out[2 * (m + pg.M * (n0 + j)) + comp] = ((kMapMinusMask >> G) & 1) ? -v : v;
      }
    }
    __syncthreads();
  }
  (void)gt;
}

template <typename F>
__global__ void __launch_bounds__(256)
splitk_reduce_pos_kernel(F *__restrict__ out, const F *__restrict__ partial, const long long nreal, const int ksplit) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nreal) return;
  F s = 0;
  for (int ks = 0; ks < ksplit; ks++) s += partial[(size_t)ks * nreal + i];
  out[i] = s;
}

// phase_eo[s][im][i]: the phase of the site behind position i of a (t, parity) run with (t + parity) & 1 == s
struct PhaseEoArg {
  int localL[3], totalL[3], commCoord[3];
  int V3h, Lh, Nmom, ftsign;
};
template <typename F>
__global__ void __launch_bounds__(256) phase_matrix_eo_kernel(F *__restrict__ phase, const int *__restrict__ mom, const PhaseEoArg a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int im = blockIdx.y, s = blockIdx.z;
  if (i >= a.V3h) return;
  const int xh = i % a.Lh, yz = i / a.Lh;
  const int y = yz % a.localL[1], z = yz / a.localL[1];
  const int x = 2 * xh + ((y + z + s) & 1);
  const int gx = x + a.commCoord[0] * a.localL[0], gy = y + a.commCoord[1] * a.localL[1], gz = z + a.commCoord[2] * a.localL[2];
  const double phi = (double)(mom[0 + 3 * im] * gx) / (double)a.totalL[0] + (double)(mom[1 + 3 * im] * gy) / (double)a.totalL[1] +
                     (double)(mom[2 + 3 * im] * gz) / (double)a.totalL[2];
  double sn, cs;
  sincospi(2.0 * phi, &sn, &cs);
  st_c<F>(phase + 2 * (((size_t)s * a.Nmom + im) * a.V3h + i), make_c<F>((F)cs, (F)((double)a.ftsign * sn)));
}

int phase_matrix_eo(void *phase_d, const int *mom_h, int Nmom, int ftsign, const int localL[4], const int totalL[4],
                    const int commCoord[4], int precision, cudaStream_t stream) {
  PhaseEoArg a;
  for (int i = 0; i < 3; i++) {
    a.localL[i] = localL[i];
    a.totalL[i] = totalL[i];
    a.commCoord[i] = commCoord ? commCoord[i] : 0;
  }
  a.Lh = localL[0] / 2;
  a.V3h = a.Lh * localL[1] * localL[2];
  a.Nmom = Nmom;
  a.ftsign = ftsign;
  int *mom_d = nullptr;
  MUGIQ_CUDA_CHECK(cudaMallocAsync((void **)&mom_d, sizeof(int) * 3 * Nmom, stream));
  MUGIQ_CUDA_CHECK(cudaMemcpyAsync(mom_d, mom_h, sizeof(int) * 3 * Nmom, cudaMemcpyHostToDevice, stream));
  const dim3 grid((a.V3h + 255) / 256, Nmom, 2);
  {
    ProfScope prof(K_PHASE, stream, 2.0 * a.V3h * Nmom * 2.0 * prec_bytes(precision));
    if (precision == MUGIQ_B200_PREC_DOUBLE)
      phase_matrix_eo_kernel<double><<<grid, 256, 0, stream>>>((double *)phase_d, mom_d, a);
    else
      phase_matrix_eo_kernel<float><<<grid, 256, 0, stream>>>((float *)phase_d, mom_d, a);
    MUGIQ_LAUNCH_CHECK();
  }
  MUGIQ_CUDA_CHECK(cudaFreeAsync(mom_d, stream));
  MUGIQ_CUDA_CHECK(cudaStreamSynchronize(stream));  // mom_h may be a temporary of the caller
  return MUGIQ_B200_OK;
}

static PosGeom make_pos_geom(const LatGeom &g, int nLoop, int N) {
  PosGeom pg;
  pg.Lt = g.L[3];
  pg.V3h = g.V3 / 2;
  pg.nLoop = nLoop;
  pg.N = N;
  pg.Vh = g.volumeCB;
  pg.V4 = g.volume;
  pg.M = (long long)g.L[3] * 16 * nLoop;
  // CTAs of the DMMA kernel: classes c = 2*parity + s hold the time-slices with t & 1 == s ^ parity
  pg.nT[0] = (pg.Lt + 1) / 2;
  pg.nT[1] = pg.Lt / 2;
  pg.blk0[0] = 0;
  for (int c = 0; c < 4; c++) {
    const int nR = pg.nT[(c & 1) ^ (c >> 1)] * nLoop;
    pg.blk0[c + 1] = pg.blk0[c] + (nR + kPosWarps - 1) / kPosWarps;
  }
  // split of the V3/2 run: whole tiles, at least 4 tiles per chunk, at most 16 chunks; the count that fills the 148 SMs
  // most evenly (CTAs resident per SM: 3 for <= 16 momenta, 2 above), fewer chunks preferred (less split-K traffic)
  const int ntiles = (pg.V3h + kPosKT - 1) / kPosKT;
  const int slots = 148 * (std::min(N, 36) <= 16 ? 3 : 2);
  int best = 1;
  double best_score = -1.0;
  for (int nc = 1; nc <= 16 && nc * 4 <= std::max(ntiles, 4); nc++) {
    const int tpc = (ntiles + nc - 1) / nc;
    const int real_nc = (ntiles + tpc - 1) / tpc;
    const long long ctas = (long long)pg.blk0[4] * real_nc;
    const long long waves = (ctas + slots - 1) / slots;
    const double score = (double)ctas / (double)(waves * slots) - 0.004 * real_nc;
    if (score > best_score + 1e-9) {
      best_score = score;
      best = real_nc;
    }
  }
  const int tpc = (ntiles + best - 1) / best;
  pg.kchunk = tpc * kPosKT;
  pg.nchunk = (ntiles + tpc - 1) / tpc;
  return pg;
}

long long momproj_pos_workspace_bytes(const LatGeom &g, int nLoop, int N, int precision) {
  const PosGeom pg = make_pos_geom(g, nLoop, N);
  return (long long)2 * pg.nchunk * pg.M * N * 2 * (long long)prec_bytes(precision);
}

int momproj_pos(void *mom_d, const void *pos_d, const void *phase_eo_d, int nLoop, int N, const LatGeom &g, int precision,
                void *workspace_d, cudaStream_t stream) {
  const PosGeom pg = make_pos_geom(g, nLoop, N);
  const int ksplit = 2 * pg.nchunk;
  const long long ntask = (long long)pg.Lt * nLoop * ksplit;
  const double bytes = 2.0 * prec_bytes(precision) * ((double)pg.M * g.V3 + 2.0 * (double)g.V3 * N + (double)pg.M * N * (ksplit + 1));
  {
    ProfScope prof(K_MOMPROJ, stream, bytes, 8.0 * (double)pg.M * N * (double)g.V3);
    if (precision == MUGIQ_B200_PREC_DOUBLE) {
      const int blocks = pg.blk0[4] * pg.nchunk;
      const double *A = (const double *)pos_d, *P = (const double *)phase_eo_d;
      double *ws = (double *)workspace_d;
      // momenta in passes of up to 36 (9 tiles of 4): A is re-read once per pass
      for (int n0 = 0; n0 < N; n0 += 36) {
        const int nt = (std::min(N - n0, 36) + 3) / 4;
        switch (nt) {
#define MUGIQ_NT_CASE(k) \
  case k: momproj_pos_dmma_kernel<k><<<blocks, kPosWarps * 32, 0, stream>>>(ws, A, P, pg, n0); break;
          MUGIQ_NT_CASE(9) MUGIQ_NT_CASE(8) MUGIQ_NT_CASE(7) MUGIQ_NT_CASE(6) MUGIQ_NT_CASE(5)
          MUGIQ_NT_CASE(4) MUGIQ_NT_CASE(3) MUGIQ_NT_CASE(2)
          default: momproj_pos_dmma_kernel<1><<<blocks, kPosWarps * 32, 0, stream>>>(ws, A, P, pg, n0); break;
#undef MUGIQ_NT_CASE
        }
        MUGIQ_LAUNCH_CHECK();
      }
    } else {
      momproj_pos_simt_kernel<float><<<(unsigned)ntask, 128, 0, stream>>>((float *)workspace_d, (const float *)pos_d,
                                                                          (const float *)phase_eo_d, pg);
      MUGIQ_LAUNCH_CHECK();
    }
  }
  const long long nreal = 2 * pg.M * N;
  const int rblocks = (int)((nreal + 255) / 256);
  ProfScope prof2(K_SPLITK_REDUCE, stream, (double)nreal * (ksplit + 1) * prec_bytes(precision));
  if (precision == MUGIQ_B200_PREC_DOUBLE)
    splitk_reduce_pos_kernel<double><<<rblocks, 256, 0, stream>>>((double *)mom_d, (const double *)workspace_d, nreal, ksplit);
  else
    splitk_reduce_pos_kernel<float><<<rblocks, 256, 0, stream>>>((float *)mom_d, (const float *)workspace_d, nreal, ksplit);
  MUGIQ_LAUNCH_CHECK();
  return MUGIQ_B200_OK;
}

}  // namespace mugiq_b200
