/* mugiq_b200.h — C-ABI of the B200-native disconnected-loop hot path.
 *
 * Plain C, POD arguments only (raw device/host pointers, ints, doubles): this is the boundary a
 * maintainer of ckallidonis/mugiq binds instead of the reference's CUDA translation units
 * (lib/contract_wrappers.cu, lib/mugiq_{contract,displace,util}_kernels.cu) and the cuBLAS call in
 * lib/loop_mugiq.cpp:358-387.  The reference itself has no true C ABI (MugiqLoopParam holds
 * std::vector/std::string inside extern "C", include/mugiq.h:16-47); every entry point below cites
 * the reference function it replaces.  INTEGRATION.md shows the reference-side call sites.
 *
 * Conventions
 *  - Every function returns 0 on success and a negative MUGIQ_B200_E* code on failure;
 *    mugiq_b200_last_error() returns a human-readable message for the calling thread.
 *    (The reference aborts through errorQuda; the C++ mirror in mugiq_b200/host translates a
 *    non-zero status into the same abort.)
 *  - "_d" pointers are device pointers, "_h" pointers are host pointers.
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Calls are
 *    asynchronous with respect to the host unless stated otherwise; nothing here calls
 *    cudaDeviceSynchronize (the reference synchronises after every launch,
 *    lib/contract_wrappers.cu:71,110,151,192).
 *  - Complex numbers are interleaved (re, im) pairs of the field precision.
 *
 * Lattice geometry and memory layouts
 *  - Sites use QUDA's even/odd (checkerboard) order: parity = (x+y+z+t)&1, x_cb = lexicographic>>1,
 *    lexicographic = x + Lx*(y + Ly*(z + Lz*t));  full-site index x_eo = x_cb + parity*volumeCB
 *    (this is `tid` of lib/mugiq_contract_kernels.cu:52).  L[0] must be even; every entry point that displaces
 *    (displace*, loop_accumulate / loop_plan_create with entries) needs all four extents even, as QUDA does.
 *  - Colour-spinor fields (12 complex per site, component index = colour + 3*spin,
 *    include/util_mugiq.h:19):
 *      MUGIQ_B200_ORDER_SITE   [parity][x_cb][spin][colour]           canonical, site-major (192 B/site FP64)
 *      MUGIQ_B200_ORDER_FLOAT2 [parity][spin*3+colour][x_cb]          QUDA FLOAT2 native order
 *      MUGIQ_B200_ORDER_FLOAT4 [parity][j=0..5][x_cb][2 complex]      QUDA FLOAT4 native order
 *    All compute kernels consume the canonical site-major order; mugiq_b200_ingest_spinor /
 *    mugiq_b200_export_spinor convert from/to the QUDA orders.
 *  - Gauge field on the device: [mu][parity][x_cb][row][col] complex (the reference's host QDP order,
 *    lib/displace.cpp:70-100, kept on the device), 144 B/link FP64.
 *  - Position-space loop buffer ("dataPos"): complex index x_eo + V4*(G + 16*iL)
 *    (lib/mugiq_contract_kernels.cu:120, lib/loop_mugiq.cpp:468,492).
 *  - Momentum-projection input ("dataPosMP"): t + Lt*(G' + 16*iL) + Lt*nData*v3,
 *    v3 = x + Lx*y + Lx*Ly*z  (lib/mugiq_util_kernels.cu:88-97).
 *  - Momentum-space result ("dataMom"): t + Lt*(G' + 16*iL) + Lt*nData*im (lib/loop_mugiq.cpp:415-418).
 */
#ifndef MUGIQ_B200_H
#define MUGIQ_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define MUGIQ_B200_VERSION 100 /* 0.1.0 */

/* error codes */
#define MUGIQ_B200_OK 0
#define MUGIQ_B200_EINVAL (-1)  /* bad argument (what errorQuda reports for bad params) */
#define MUGIQ_B200_ECUDA (-2)   /* CUDA runtime error (what checkCudaError() reports) */
#define MUGIQ_B200_ENOMEM (-3)  /* allocation failed */
#define MUGIQ_B200_ESTATE (-4)  /* wrong call order */

/* precision tags: numerically equal to QudaPrecision (4 = single, 8 = double) */
#define MUGIQ_B200_PREC_SINGLE 4
#define MUGIQ_B200_PREC_DOUBLE 8

/* colour-spinor memory orders (see header comment) */
#define MUGIQ_B200_ORDER_SITE 0
#define MUGIQ_B200_ORDER_FLOAT2 2
#define MUGIQ_B200_ORDER_FLOAT4 4

/* displacement direction / sign: numerically equal to DisplaceDir / DisplaceSign (include/enum_mugiq.h:72-85) */
#define MUGIQ_B200_DIR_X 0
#define MUGIQ_B200_DIR_Y 1
#define MUGIQ_B200_DIR_Z 2
#define MUGIQ_B200_DIR_T 3
#define MUGIQ_B200_SIGN_MINUS 0
#define MUGIQ_B200_SIGN_PLUS 1

#define MUGIQ_B200_NGAMMA 16
#define MUGIQ_B200_MAX_ENTRIES 64 /* displacement entries per call of the fused path */

/* Local lattice geometry.  Replaces ArgGeom (include/contract_util.cuh:71-120): dims, volumeCB and
 * volume are derived from L; only full-site-subset fields are supported, as in the reference
 * (lib/contract_wrappers.cu:100,185). */
typedef struct mugiq_b200_geom_s {
  int L[4];      /* local lattice extents x,y,z,t; L[0] even */
  int precision; /* MUGIQ_B200_PREC_* of spinors, links and loop buffers */
} mugiq_b200_geom_t;

/* One displacement entry "<sign><dir>:<start>[,<stop>]" of the reference's --displace-entry-string
 * (tests/loop.cpp:607-718; include/loop_mugiq.h:221-250). */
typedef struct mugiq_b200_disp_entry_s {
  int dir;   /* MUGIQ_B200_DIR_* */
  int sign;  /* MUGIQ_B200_SIGN_* */
  int start; /* first displacement length contracted (>= 1) */
  int stop;  /* last displacement length contracted (>= start) */
} mugiq_b200_disp_entry_t;

/* ---- library -------------------------------------------------------------------------------- */
int mugiq_b200_version(void);
const char *mugiq_b200_last_error(void);
/* Device query: fills name (<= name_len), SM count, compute capability major*10+minor, free and
 * total device memory.  Replaces printGPUMemInfo (lib/util_mugiq.cpp:22-33). */
int mugiq_b200_device_info(char *name, int name_len, int *sm_count, int *cc, long long *free_bytes,
                           long long *total_bytes);

/* ---- gamma tables ---------------------------------------------------------------------------- */
/* Host copy of the tables the kernels have compiled in.  Replaces copyGammaCoeffStructToSymbol and
 * copyGammaMapStructToSymbol (lib/contract_wrappers.cu:6-47; data include/gamma.h:32-109).
 * row_value[16][4][2] (re,im), column_index[16][4], map_sign[16], map_index[16]; any pointer may be NULL. */
int mugiq_b200_gamma_tables(double *row_value, int *column_index, double *map_sign, int *map_index);

/* ---- layout conversion ----------------------------------------------------------------------- */
/* QUDA FLOAT2/FLOAT4 -> canonical site-major and back (SURVEY §7 "layout contract"). */
int mugiq_b200_ingest_spinor(void *dst_site_d, const void *src_d, int src_order,
                             const mugiq_b200_geom_t *geom, void *stream);
int mugiq_b200_export_spinor(void *dst_d, int dst_order, const void *src_site_d,
                             const mugiq_b200_geom_t *geom, void *stream);
/* nfields conversions in one launch (HOST arrays of device pointers): what a QUDA-backed caller uses per eigenvector
 * batch, one 50 MB field per launch leaves the GPU mostly idle. */
int mugiq_b200_ingest_spinor_batch(void *const *dst_site_d, const void *const *src_d, int nfields, int src_order,
                                   const mugiq_b200_geom_t *geom, void *stream);
/* Host QDP-order gauge (void* gauge[4], one pointer per direction, MugiqLoopParam::gauge,
 * include/mugiq.h:43) -> device [mu][parity][x_cb][3][3].  Replaces Displace::createCudaGaugeField /
 * createExtendedCudaGaugeField for an unpartitioned lattice (lib/displace.cpp:70-134).  Synchronous. */
int mugiq_b200_gauge_upload(void *gauge_d, const void *const gauge_h[4], const mugiq_b200_geom_t *geom,
                            void *stream);

/* ---- stage 1: loop contraction --------------------------------------------------------------- */
/* loop_d[x_eo + V4*G] += (1/sigma) * sum_{s2} rowval[G][s2] * sum_c conj(vL[s2,c]) * vR[col[G][s2],c]
 * for the 16 G.  Replaces performLoopContraction (lib/contract_wrappers.cu:88-115) and
 * loopContract_kernel (lib/mugiq_contract_kernels.cu:45-122); inv_sigma = 1.0/sigma as in
 * LoopContractArg (include/contract_util.cuh:133). */
int mugiq_b200_contract(void *loop_d, const void *vL_d, const void *vR_d, double sigma,
                        const mugiq_b200_geom_t *geom, void *stream);
/* Same sum for nvec eigenvector pairs in one launch with register accumulation and a single
 * read-modify-write of loop_d (accumulate != 0) or a plain store (accumulate == 0).
 * vL_d / vR_d are HOST arrays of nvec device pointers; vR_d == NULL means vR = vL (ultra-local). */
int mugiq_b200_contract_batch(void *loop_d, const void *const *vL_d, const void *const *vR_d,
                              const double *sigma_h, int nvec, int accumulate,
                              const mugiq_b200_geom_t *geom, void *stream);

/* ---- stage 2: covariant displacement ---------------------------------------------------------- */
/* sign = PLUS : dst(x) = U_dir(x) * src(x + dir)
 * sign = MINUS: dst(x) = U_dir(x - dir)^dagger * src(x - dir),  periodic wrap, dst != src.
 * Replaces performCovariantDisplacementVector (lib/contract_wrappers.cu:171-198) and
 * covariantDisplacementVector_kernel (lib/mugiq_displace_kernels.cu:156-185). */
int mugiq_b200_displace(void *dst_d, const void *src_d, const void *gauge_d, int dir, int sign,
                        const mugiq_b200_geom_t *geom, void *stream);
/* The same hop for nvec fields in one launch (HOST arrays of device pointers): one link tile in shared memory serves
 * the whole batch, 2S + U/nvec bytes per eigvec*site instead of 2S + U.  What Displace::doVectorDisplacement
 * (lib/displace.cpp:55-67) does per eigenvector inside the loop nest lib/loop_mugiq.cpp:478-507, hoisted over n. */
int mugiq_b200_displace_batch(void *const *dst_d, const void *const *src_d, int nvec, const void *gauge_d, int dir,
                              int sign, const mugiq_b200_geom_t *geom, void *stream);

/* ---- stages 1 and 2 directly on QUDA-native fields ------------------------------------------------------------ */
/* The same two operations on fields in QUDA's FLOAT2 / FLOAT4 orders (order = MUGIQ_B200_ORDER_FLOAT2 / _FLOAT4), which is
 * how the reference's kernels see them through FieldOrderCB (lib/mugiq_contract_kernels.cu:82-83,
 * lib/mugiq_displace_kernels.cu:85-113): no layout conversion and no scratch field for the reference-shaped single
 * calls performLoopContraction / performCovariantDisplacementVector (lib/contract_wrappers.cu:88-115,171-198); source
 * AND destination of the displacement are native-order fields.  vL_d / vR_d / dst_d / src_d are HOST arrays of nvec
 * device pointers; vR_d == NULL means vR = vL; accumulate as in mugiq_b200_contract_batch. */
int mugiq_b200_contract_native(void *loop_d, const void *const *vL_d, const void *const *vR_d, const double *sigma_h, int nvec,
                               int order, int accumulate, const mugiq_b200_geom_t *geom, void *stream);
int mugiq_b200_displace_native(void *const *dst_d, const void *const *src_d, int nvec, const void *gauge_d, int dir, int sign,
                               int order, const mugiq_b200_geom_t *geom, void *stream);

/* ---- stages 1+2 fused: the eigenvector loop of Loop_Mugiq::computeCoarseLoop -------------------- */
/* For every eigenvector n and every loop iL (0 = ultra-local, then the entries in order, lengths
 * start..stop):  dataPos[x_eo + V4*(G + 16*iL)] (+)= (1/sigma_n) v_n(x)^dag Gamma_G (D^k v_n)(x).
 * Replaces the body of the eigenvector/displacement loop nest lib/loop_mugiq.cpp:455-509 (including
 * Displace::doVectorDisplacement, lib/displace.cpp:55-67) without materialising displaced vectors
 * for one-hop entries.  evec_d is a HOST array of nvec device pointers (canonical order).
 * workspace_d: device scratch of mugiq_b200_loop_workspace_bytes() bytes (may be NULL if that is 0).
 * accumulate == 0 overwrites dataPos_d, != 0 adds to it (eigenvector batches / shards). */
long long mugiq_b200_loop_workspace_bytes(const mugiq_b200_geom_t *geom, int nvec,
                                          const mugiq_b200_disp_entry_t *entries, int nentries);
int mugiq_b200_loop_accumulate(void *dataPos_d, const void *const *evec_d, const double *sigma_h, int nvec,
                               const void *gauge_d, const mugiq_b200_disp_entry_t *entries, int nentries,
                               int accumulate, void *workspace_d, const mugiq_b200_geom_t *geom,
                               void *stream);

/* Plan form of the same loop nest, for callers that feed the eigenvectors in several batches (host-streamed
 * batches, shards): the plan owns what depends on the gauge field and the entry list only — the Wilson lines
 * W_k(x) = U(x) U(x+mu) ... U(x+(k-1)mu) (a displacement of length k is then ONE 3x3 multiply per site instead of
 * k hops of lib/displace.cpp:55-67) and the launch schedule.  Plays the role of the Displace object
 * (lib/displace.cpp:4-37) for the fused path.
 *   create     : builds the Wilson lines on `stream` (allocates device memory for them)
 *   accumulate : dataPos (+)= contribution of the given eigenvectors, for every loop the plan COMPUTES
 *   finalize   : fills the slots the plan DERIVES after the eigenvector sum — the minus-direction loop of a
 *                (+mu,-mu) entry pair from its plus partner, T-_G(x) = h_G conj(T+_G(x - k mu)) with
 *                Gamma_G^dag = h_G Gamma_G (exact identity), and copies for repeated requests.  Call once, after
 *                the last accumulate (and after any cross-rank reduction is fine too: it is linear).
 * MUGIQ_B200_NO_PM_SYMMETRY=1 in the environment makes the plan compute every requested loop. */
typedef struct mugiq_b200_loop_plan_s mugiq_b200_loop_plan_t;
int mugiq_b200_loop_plan_create(mugiq_b200_loop_plan_t **plan, const void *gauge_d, const mugiq_b200_disp_entry_t *entries,
                                int nentries, const mugiq_b200_geom_t *geom, void *stream);
int mugiq_b200_loop_plan_destroy(mugiq_b200_loop_plan_t *plan);
int mugiq_b200_loop_plan_nloop(const mugiq_b200_loop_plan_t *plan);
int mugiq_b200_loop_plan_info(const mugiq_b200_loop_plan_t *plan, int *ncomputed, int *nderived, int *ngroups,
                              long long *wilson_bytes);
/* The dataPos slots (iL) accumulate() writes - what a cross-rank sum has to cover; the others are filled by finalize().
 * Writes at most max_slots of them, returns how many there are. */
int mugiq_b200_loop_plan_computed_slots(const mugiq_b200_loop_plan_t *plan, int *slots, int max_slots);
/* Eigenvectors in QUDA's native FLOAT2 order ([parity][spin*3+colour][x_cb], what FieldOrderCB hands the reference's kernels,
 * lib/mugiq_contract_kernels.cu:82-83; interface_mugiq.cpp:226-235 dispatches on it): after set_evec_order(plan,
 * MUGIQ_B200_ORDER_FLOAT2) every accumulate form of the plan (accumulate, accumulate_allreduce, the feed) takes FLOAT2
 * fields and the fused kernel stages them itself - each field is a 4-D tensor (8 sites, 12 components, volumeCB/8 chunks,
 * 2 parities) and the pieces of a stage arrive as TMA tensor boxes - so no layout-conversion pass and no site-major copy
 * exist.  Needs volumeCB % 8 == 0.  MUGIQ_B200_ORDER_SITE (default) restores the canonical order. */
int mugiq_b200_loop_plan_set_evec_order(mugiq_b200_loop_plan_t *plan, int order);
/* Lattice-T split (BASELINE config 5): a rank runs the plan on its time slab EXTENDED by halo slices.
 *   set_t_range : accumulate() computes dataPos only on the time-slices [t_begin, t_end) of the plan's lattice (the
 *                 rank's interior) and merely reads the others; finalize() still spans the whole lattice.
 *   t_halo      : what the plan reads across a slab boundary - eigenvector slices below the interior (directly
 *                 computed minus-t loops), above it (plus-t loops), and loop-buffer slices below it (minus-t loops
 *                 DERIVED from their plus partner: the caller fetches the partner's top slices from the rank below
 *                 instead of computing them, once per run instead of once per eigenvector).
 * Replaces the per-hop exchangeGhost of lib/contract_wrappers.cu:166-174 and the extended gauge field of
 * lib/displace.cpp:104-134 for a partitioned t direction. */
int mugiq_b200_loop_plan_set_t_range(mugiq_b200_loop_plan_t *plan, int t_begin, int t_end);
int mugiq_b200_loop_plan_t_halo(const mugiq_b200_loop_plan_t *plan, int *evec_lower, int *evec_upper, int *loop_lower);
int mugiq_b200_loop_plan_accumulate(const mugiq_b200_loop_plan_t *plan, void *dataPos_d, const void *const *evec_d,
                                    const double *sigma_h, int nvec, int accumulate, void *stream);
int mugiq_b200_loop_plan_finalize(const mugiq_b200_loop_plan_t *plan, void *dataPos_d, int accumulate, void *stream);

/* Host-only diagnostic (no GPU needed): the fused kernel's tiling for launch group `group` of the plan these entries
 * produce, over the time-slices [t_begin, t_end) (t_end < 0: all), checked CTA by CTA - every thread's own and neighbour
 * site must lie in the merged intervals its CTA stages per eigenvector.  group < 0 returns the number of launch groups.
 * out = {run (checkerboard sites per parity and CTA), warps per loop, ring stages, stage bytes, most bulk copies per
 * stage, mean sites staged per CTA and eigenvector, sites NOT found in their stage (must be 0), malformed stage maps
 * (must be 0)}.  evec_order: MUGIQ_B200_ORDER_SITE or _FLOAT2 (the tiling of natively ordered eigenvectors, see
 * mugiq_b200_loop_plan_set_evec_order). */
int mugiq_b200_fused_tiling_check(const mugiq_b200_disp_entry_t *entries, int nentries, const mugiq_b200_geom_t *geom,
                                  int t_begin, int t_end, int group, int evec_order, long long out[8]);

/* ---- streamed eigenvector feed of a loop plan --------------------------------------------------------------------- */
/* The producer/consumer form of the eigenvector loop: the reference makes the fine eigenvector right before it is used,
 * `prolongateEvec(fineEvecL, eVecs[n])` through the multigrid transfer operators or a field copy (lib/loop_mugiq.cpp:276-319,
 * 478-483).  1000-2000 fine eigenvectors of the BASELINE lattices do not fit a GPU, so the fused path takes them as a
 * stream of batches: the feed owns `nbuf` device staging batches of `batch` fields each (layout `order`: SITE or FLOAT2,
 * which the kernels stage directly, or FLOAT4, which is converted per batch); a producer fills batch b+1 on ITS stream while the loop kernels consume
 * batch b on the feed's compute stream.
 *   create    : `stream` is the compute stream; accumulate != 0: the first batch adds to dataPos_d instead of overwriting
 *   acquire   : n <= batch device field pointers of the next staging batch; `producer_stream` is made to wait until the
 *               kernels that last read that batch have finished (stream-ordered, the host does not block)
 *   commit    : the producer's writes on `producer_stream` are complete in stream order; enqueues (conversion +) the plan's
 *               kernels for these n eigenvectors on the compute stream
 *   push_host : acquire + cudaMemcpyAsync from (pinned) host fields on the feed's copy stream + commit, batch by batch:
 *               the H2D copies of batch b+1 overlap the kernels of batch b; consecutive host fields travel as one copy
 *   set_plan  : between runs, point the feed at another plan (e.g. rebuilt for a new gauge field) / loop buffer of the same
 *               lattice and precision; the staging batches are kept
 *   finish    : returns the number of eigenvectors consumed and re-arms the feed; work enqueued on the compute stream
 *               afterwards (loop_plan_finalize, the cross-rank sum, the projection) sees the complete sum */
typedef struct mugiq_b200_loop_feed_s mugiq_b200_loop_feed_t;
int mugiq_b200_loop_feed_create(mugiq_b200_loop_feed_t **feed, const mugiq_b200_loop_plan_t *plan, void *dataPos_d, int batch,
                                int nbuf, int order, int accumulate, void *stream);
int mugiq_b200_loop_feed_destroy(mugiq_b200_loop_feed_t *feed);
int mugiq_b200_loop_feed_set_plan(mugiq_b200_loop_feed_t *feed, const mugiq_b200_loop_plan_t *plan, void *dataPos_d);
int mugiq_b200_loop_feed_acquire(mugiq_b200_loop_feed_t *feed, void **field_d, int n, void *producer_stream);
int mugiq_b200_loop_feed_commit(mugiq_b200_loop_feed_t *feed, const double *sigma_h, int n, void *producer_stream);
int mugiq_b200_loop_feed_push_host(mugiq_b200_loop_feed_t *feed, const void *const *evec_h, const double *sigma_h, int n);
int mugiq_b200_loop_feed_finish(mugiq_b200_loop_feed_t *feed, long long *nvec_total);

/* ---- stage 3: gamma-basis / time-slice reorder -------------------------------------------------- */
/* out[t + Lt*((15-G) + 16*iL) + Lt*nData*v3] = sign[G] * in[x_eo + V4*(G + 16*iL)].
 * Replaces convertIdxOrder_mapGamma (lib/contract_wrappers.cu:133-156,
 * lib/mugiq_util_kernels.cu:59-99).  nData must equal 16*nLoop (lib/contract_wrappers.cu:138). */
int mugiq_b200_reorder_mapgamma(void *out_d, const void *in_d, int nData, int nLoop,
                                const mugiq_b200_geom_t *geom, void *stream);

/* ---- stage 4: momentum projection --------------------------------------------------------------- */
/* phase_d[v3 + V3*im] = cos(2 pi phi) + i*ftsign*sin(2 pi phi),
 * phi = sum_d mom[d + 3*im]*(x_d + commCoord_d*localL_d)/totalL_d.
 * Replaces createPhaseMatrixGPU (lib/contract_wrappers.cu:50-77, lib/mugiq_util_kernels.cu:3-35).
 * mom_h: host int[3*Nmom] in MOM_MATRIX_IDX order (include/util_mugiq.h:24). */
int mugiq_b200_phase_matrix(void *phase_d, const int *mom_h, int Nmom, int ftsign, const int localL[4],
                            const int totalL[4], const int commCoord[4], int precision, void *stream);
/* dataMom(M x N) = dataPosMP(M x K) * phase(K x N), all column-major, alpha = 1, beta = 0.
 * Replaces cublasZgemm/cublasCgemm of lib/loop_mugiq.cpp:364-377.  workspace_d: device scratch of
 * mugiq_b200_momproj_workspace_bytes(M, N, K, precision) bytes (split-K partial sums). */
long long mugiq_b200_momproj_workspace_bytes(long long M, int N, long long K, int precision);
int mugiq_b200_momproj(void *mom_d, const void *posMP_d, const void *phase_d, long long M, int N,
                       long long K, int precision, void *workspace_d, void *stream);

/* ---- stages 3+4 fused: momentum projection straight from dataPos ------------------------------------------------ */
/* dataMom[t + Lt*((15-G) + 16*iL) + Lt*nData*im] = sum_v3 sign[G] * dataPos[x_eo(v3,t) + V4*(G + 16*iL)] * phase(v3, im).
 * One call replaces convertIdxOrder_mapGamma (lib/contract_wrappers.cu:133-156) AND cublasZgemm/Cgemm
 * (lib/loop_mugiq.cpp:364-377): the GEMM reads the position-space buffer in place (for fixed G, iL, t, parity the
 * V3/2 sites are contiguous), so the reorder pass, its 2 x 16*V4*nLoop complex of HBM traffic and the dataPosMP buffer
 * disappear.  phase_eo_d holds the phase matrix in the matching even/odd order, phase_eo[s][im][i] with
 * s = (t + parity) & 1, 2*Nmom*V3/2 complex, built by mugiq_b200_phase_matrix_eo (same arguments as
 * mugiq_b200_phase_matrix; replaces createPhaseMatrixGPU, lib/contract_wrappers.cu:50-77, for this path). */
int mugiq_b200_phase_matrix_eo(void *phase_eo_d, const int *mom_h, int Nmom, int ftsign, const int localL[4],
                               const int totalL[4], const int commCoord[4], int precision, void *stream);
long long mugiq_b200_momproj_pos_workspace_bytes(const mugiq_b200_geom_t *geom, int nLoop, int Nmom);
int mugiq_b200_momproj_pos(void *mom_d, const void *dataPos_d, const void *phase_eo_d, int nLoop, int Nmom,
                           const mugiq_b200_geom_t *geom, void *workspace_d, void *stream);

/* ---- eigenvector shards: the cross-GPU sum of the loop buffer --------------------------------------------------- */
/* One process per GPU; every rank runs the loop plan on its shard of the eigenvectors (the gauge field is replicated),
 * then the loop buffer is summed over the ranks with NCCL over NVLink / NVSwitch.  Replaces the reference's D2H copy +
 * MPI_Reduce over COMM_SPACE + MPI_Gather over COMM_TIME + MPI_Bcast on host buffers (lib/loop_mugiq.cpp:386-424) and
 * the MPI_Comm_split pair of Loop_Mugiq::setupComms (:62-88).  NCCL is bound at run time (the process's own
 * libnccl.so.2); the caller moves the 128-byte id from rank 0 to the other ranks by whatever transport it has (the
 * Python front end: its process group; the C++ front end: a file).
 *   comm_unique_id : rank 0 makes the id (ncclGetUniqueId)
 *   comm_create    : collective over all `size` ranks, on the calling thread's current CUDA device
 *   allreduce      : in-place sum of `count` REAL numbers of the given precision, asynchronous on `stream`
 *   allgather      : every rank contributes `bytes` bytes, rank-major result (the COMM_TIME gather of a lattice-T split)
 *   allreduce_pos  : in-place sum of the time-slices [t_begin, t_end) (t_end < 0: all) of the loop slots slots_h[0..nslots)
 *                    of a position-space buffer - the "per-time-slice loop buffer" - as one grouped NCCL launch
 *   loop_plan_accumulate_allreduce : mugiq_b200_loop_plan_accumulate for the LAST (or only) eigenvector batch of a
 *                    sharded run, fused with the sum of every slot the plan computes: the lattice is processed in
 *                    time-slice chunks - each half of what is left, down to Lt / (2 nchunks) slices; nchunks <= 1: one -
 *                    and the all-reduce of chunk k runs on a high-priority side stream while the kernels of chunk k+1
 *                    compute (only the last, smallest chunk's sum is exposed).  On return,
 *                    work enqueued on `stream` sees the summed buffer; call loop_plan_finalize afterwards (the derived
 *                    slots are linear in the computed ones, so they need no sum of their own). */
#define MUGIQ_B200_COMM_ID_BYTES 128
typedef struct mugiq_b200_comm_s mugiq_b200_comm_t;
int mugiq_b200_comm_unique_id(void *id128);
int mugiq_b200_comm_create(mugiq_b200_comm_t **comm, const void *id128, int rank, int size);
int mugiq_b200_comm_destroy(mugiq_b200_comm_t *comm);
int mugiq_b200_comm_info(const mugiq_b200_comm_t *comm, int *rank, int *size, int *nccl_version);
int mugiq_b200_allreduce(void *buf_d, long long count, int precision, mugiq_b200_comm_t *comm, void *stream);
int mugiq_b200_allgather(void *recv_d, const void *send_d, long long bytes, mugiq_b200_comm_t *comm, void *stream);
int mugiq_b200_allreduce_pos(void *dataPos_d, const int *slots_h, int nslots, int t_begin, int t_end,
                             const mugiq_b200_geom_t *geom, mugiq_b200_comm_t *comm, void *stream);
/* Peer transport of the overlapped position-space sum: every rank maps all ranks' position-space buffers and one staging
 * area per rank (mugiq_b200_peer_alloc / _open; own pointers at index `rank`) and attaches the tables.  While attached,
 * loop_plan_accumulate_allreduce on the attached buffer moves the chunks with the copy engines (reduce-scatter into the
 * owners' staging areas, a small summing kernel, all-gather into the peers' buffers; one-element NCCL all-reduces order
 * the rounds) instead of NCCL's all-reduce kernels, which compete with the FP64-bound loop kernels for SMs.
 *   comm_stage_bytes  : staging bytes per rank for the plan's computed loops summed in `nchunks` chunks over `size` ranks
 *   comm_attach_peers : tables of `size` device pointers each (256-byte aligned); NULL, NULL detaches */
long long mugiq_b200_comm_stage_bytes(const mugiq_b200_loop_plan_t *plan, int nchunks, int size);
int mugiq_b200_comm_attach_peers(mugiq_b200_comm_t *comm, void *const *peer_pos_d, void *const *peer_stage_d, long long stage_bytes);
int mugiq_b200_loop_plan_accumulate_allreduce(const mugiq_b200_loop_plan_t *plan, void *dataPos_d, const void *const *evec_d,
                                              const double *sigma_h, int nvec, int accumulate, mugiq_b200_comm_t *comm,
                                              int nchunks, void *stream);

/* ---- lattice-T split: halo slices over NVLink peer memory ------------------------------------------------------ */
/* Replaces, for a partitioned t direction, ColorSpinorField::exchangeGhost (lib/contract_wrappers.cu:166-174: one
 * host-staged nFace = 1 halo per hop and eigenvector).  Every rank (one process per GPU) stores its eigenvector slabs
 * in the extended layout [vector][parity][t = 0 .. Lt_ext)[V3/2][12 complex] in one allocation made by peer_alloc; the
 * 64-byte handle travels to the two time neighbours (any transport: the Python front end sends it over its process group),
 * which map the allocation with peer_open.  halo_push_t then writes nslices boundary time-slices of a whole batch of
 * vectors straight into the neighbour's halo slices over NVLink: mode 0 = one strided 2-D copy on the copy engines
 * (no SM, no staging buffer, no pack / unpack), mode 1 = SM kernel with 128-bit peer stores.  The call is asynchronous
 * on `stream`; making the neighbour wait for the data (e.g. a one-element all-reduce on the same stream) is the
 * caller's job.  site_bytes = 192 (FP64) or 96 (FP32). */
int mugiq_b200_peer_alloc(void **ptr_d, long long bytes, void *handle64);
int mugiq_b200_peer_open(void **ptr_d, const void *handle64);
int mugiq_b200_peer_close(void *ptr_d);
int mugiq_b200_peer_free(void *ptr_d);
int mugiq_b200_halo_push_t(void *dst_slabs_d, const void *src_slabs_d, int first_vec, int nvec, int Lt_ext, long long V3h,
                           int site_bytes, int src_t, int dst_t, int nslices, int mode, void *stream);

/* ---- instrumentation ---------------------------------------------------------------------------- */
/* Per-kernel launch counters (always on) and CUDA-event timers (while enabled) around every kernel launch
 * of the library, recorded on the stream the kernel is launched on.  The reference only brackets
 * init/total/free with QUDA TimeProfile (lib/interface_mugiq.cpp:36-47,193-244).  prof_query synchronises
 * on the recorded events; alg_bytes_total / alg_flops_total are the ALGORITHMIC byte and FP64 flop counts of the
 * launches (FMA = 2 flop). */
int mugiq_b200_prof_enable(int on);
int mugiq_b200_prof_reset(void);
int mugiq_b200_prof_num_kernels(void);
const char *mugiq_b200_prof_name(int kernel_id);
int mugiq_b200_prof_query(int kernel_id, long long *launches, long long *timed_launches, double *ms_total,
                          double *alg_bytes_total, double *alg_flops_total);
/* Per-CTA timeline of the fused kernel (diagnostics): while trace_d != NULL every CTA b < capacity_ctas of
 * the following fused launches writes 16 long long to trace_d[16 b]: SM id, then %globaltimer (ns) at CTA start, stage map
 * ready, eigenvector loop entered, loop left, CTA end, two spare; the SM's clock64 at the same marks in [9..13].  NULL
 * switches it off (one predicated store per mark). */
int mugiq_b200_prof_fused_trace(void *trace_d, long long capacity_ctas);

#ifdef __cplusplus
}
#endif
#endif /* MUGIQ_B200_H */
