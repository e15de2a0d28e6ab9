// feed.cu — streamed eigenvector feed of the loop plan (include/mugiq_b200.h, mugiq_b200_loop_feed_*).
//
// The reference produces the eigenvector the loop nest works on one at a time, right before it is used:
// `prolongateEvec(fineEvecL, eVecs[n])` through the multigrid transfer operators when the eigenvectors are coarse, a field
// copy otherwise (/root/reference/lib/loop_mugiq.cpp:276-319, 478-483) - and repeats that once per displacement entry.
// 1000-2000 fine eigenvectors of a 32^3x64 or 48^3x96 lattice do not fit a GPU (403 GB / 4.08 TB), so the fused path needs
// the same producer/consumer shape, batched: the feed owns a ring of device staging batches; a PRODUCER (QUDA's
// prolongator, an eigensolver writing Ritz vectors, or the copy engines moving fields from pinned host memory) fills batch
// b+1 on its own stream while the loop kernels consume batch b on the compute stream.  Stream-ordered throughout: acquire
// makes the producer stream wait until the kernels that read the staging batch have finished, commit makes the compute
// stream wait for the producer's writes; the host never blocks except in finish.
#include <algorithm>
#include <vector>

#include "fused.cuh"
#include "plan.cuh"

using namespace mugiq_b200;

struct mugiq_b200_loop_feed_s {
  const LoopPlan *pl = nullptr;
  void *dataPos_d = nullptr;
  int batch = 0, nbuf = 0, order = MUGIQ_B200_ORDER_SITE;
  size_t field_bytes = 0;
  int volume = 0, precision = 0;  // of the plan the staging ring was sized for (kept here: a plan may be destroyed before set_plan)
  cudaStream_t compute = nullptr, copy = nullptr;
  std::vector<char *> stage;       // nbuf staging batches of `batch` fields, in the producer's order
  char *site = nullptr;            // one site-major batch: target of the layout conversion (native orders only)
  std::vector<cudaEvent_t> freed;  // staging batch b may be overwritten (recorded on the compute stream)
  std::vector<bool> used;
  cudaEvent_t filled = nullptr;
  int next = 0;          // staging batch the next acquire hands out
  int acquired = -1;     // staging batch handed out and not yet committed
  int acquired_n = 0;
  long long total = 0;   // eigenvectors consumed so far
  int accumulate0 = 0;   // the first batch adds to dataPos instead of overwriting it
};

static int feed_acquire(mugiq_b200_loop_feed_s *f, void **field_d, int n, cudaStream_t producer, const char *who) {
  if (f->acquired >= 0) return set_error(MUGIQ_B200_ESTATE, "%s: the previous batch was acquired but not committed", who);
  if (n < 1 || n > f->batch) return set_error(MUGIQ_B200_EINVAL, "%s: n = %d not in [1, %d]", who, n, f->batch);
  const int b = f->next;
  if (f->used[b]) MUGIQ_CUDA_CHECK(cudaStreamWaitEvent(producer, f->freed[b], 0));
  for (int i = 0; i < n; i++) field_d[i] = f->stage[b] + (size_t)i * f->field_bytes;
  f->acquired = b;
  f->acquired_n = n;
  return MUGIQ_B200_OK;
}

static int feed_commit(mugiq_b200_loop_feed_s *f, const double *sigma_h, int n, cudaStream_t producer, const char *who) {
  if (f->acquired < 0) return set_error(MUGIQ_B200_ESTATE, "%s: no batch was acquired", who);
  if (n < 1 || n > f->acquired_n) return set_error(MUGIQ_B200_EINVAL, "%s: n = %d, but %d fields were acquired", who, n, f->acquired_n);
  const int b = f->acquired;
  MUGIQ_CUDA_CHECK(cudaEventRecord(f->filled, producer));
  MUGIQ_CUDA_CHECK(cudaStreamWaitEvent(f->compute, f->filled, 0));
  std::vector<const void *> ptr(n);
  int rc = MUGIQ_B200_OK;
  // FLOAT2 staging batches go to the kernels as they are (tensor-box staging); FLOAT4 ones are converted first
  const bool convert = f->site != nullptr;
  if (convert) {  // producer's native order -> canonical site-major, one launch for the batch
    std::vector<void *> dst(n);
    std::vector<const void *> src(n);
    for (int i = 0; i < n; i++) {
      src[i] = f->stage[b] + (size_t)i * f->field_bytes;
      dst[i] = f->site + (size_t)i * f->field_bytes;
      ptr[i] = dst[i];
    }
    rc = convert_spinor_batch(dst.data(), src.data(), n, f->order, true, f->pl->g, f->pl->precision, f->compute);
    if (rc) return rc;
    MUGIQ_CUDA_CHECK(cudaEventRecord(f->freed[b], f->compute));  // the staging batch is free once it has been converted
  } else {
    for (int i = 0; i < n; i++) ptr[i] = f->stage[b] + (size_t)i * f->field_bytes;
  }
  rc = plan_accumulate_range(*f->pl, f->dataPos_d, ptr.data(), sigma_h, n, f->accumulate0 || f->total > 0, f->pl->t_begin,
                             f->pl->t_end, f->total == 0, f->compute, convert ? MUGIQ_B200_ORDER_SITE : f->order);
  if (rc) return rc;
  if (!convert) MUGIQ_CUDA_CHECK(cudaEventRecord(f->freed[b], f->compute));
  f->used[b] = true;
  f->total += n;
  f->next = (b + 1) % f->nbuf;
  f->acquired = -1;
  return MUGIQ_B200_OK;
}

extern "C" {

int mugiq_b200_loop_feed_create(mugiq_b200_loop_feed_t **feed, const mugiq_b200_loop_plan_t *plan, void *dataPos_d, int batch,
                                int nbuf, int order, int accumulate, void *stream) {
  const char *who = "mugiq_b200_loop_feed_create";
  if (!feed) return set_error(MUGIQ_B200_EINVAL, "%s: feed is NULL", who);
  *feed = nullptr;
  if (!plan || !dataPos_d) return set_error(MUGIQ_B200_EINVAL, "%s: NULL argument", who);
  if (batch < 1 || batch > kFusedMaxVec) return set_error(MUGIQ_B200_EINVAL, "%s: batch = %d not in [1, %d]", who, batch, kFusedMaxVec);
  if (nbuf < 2 || nbuf > 8) return set_error(MUGIQ_B200_EINVAL, "%s: nbuf = %d not in [2, 8]", who, nbuf);
  if (order != MUGIQ_B200_ORDER_SITE && order != MUGIQ_B200_ORDER_FLOAT2 && order != MUGIQ_B200_ORDER_FLOAT4)
    return set_error(MUGIQ_B200_EINVAL, "%s: unknown field order %d", who, order);
  mugiq_b200_loop_feed_s *f = new mugiq_b200_loop_feed_s;
  f->pl = &plan_of(plan);
  f->dataPos_d = dataPos_d;
  f->batch = batch;
  f->nbuf = nbuf;
  f->order = order;
  f->accumulate0 = accumulate != 0;
  f->compute = (cudaStream_t)stream;
  f->volume = f->pl->g.volume;
  f->precision = f->pl->precision;
  f->field_bytes = (size_t)f->volume * kSpinorLen * 2 * prec_bytes(f->precision);
  auto fail = [&](int code, const char *what) {
    mugiq_b200_loop_feed_destroy(f);
    return set_error(code, "%s: %s", who, what);
  };
  f->stage.assign(nbuf, nullptr);
  f->freed.assign(nbuf, nullptr);
  f->used.assign(nbuf, false);
  for (int b = 0; b < nbuf; b++) {
    if (cudaMalloc((void **)&f->stage[b], f->field_bytes * batch) != cudaSuccess) {
      cudaGetLastError();
      return fail(MUGIQ_B200_ENOMEM, "cannot allocate the staging batches");
    }
    if (cudaEventCreateWithFlags(&f->freed[b], cudaEventDisableTiming) != cudaSuccess) return fail(MUGIQ_B200_ECUDA, "cudaEventCreate failed");
  }
  const bool direct = order == MUGIQ_B200_ORDER_SITE || (order == MUGIQ_B200_ORDER_FLOAT2 && f->pl->g.volumeCB % 8 == 0);
  if (!direct && cudaMalloc((void **)&f->site, f->field_bytes * batch) != cudaSuccess) {
    cudaGetLastError();
    return fail(MUGIQ_B200_ENOMEM, "cannot allocate the conversion batch");
  }
  if (cudaEventCreateWithFlags(&f->filled, cudaEventDisableTiming) != cudaSuccess) return fail(MUGIQ_B200_ECUDA, "cudaEventCreate failed");
  if (cudaStreamCreateWithFlags(&f->copy, cudaStreamNonBlocking) != cudaSuccess) return fail(MUGIQ_B200_ECUDA, "cudaStreamCreate failed");
  *feed = f;
  return MUGIQ_B200_OK;
}

int mugiq_b200_loop_feed_set_plan(mugiq_b200_loop_feed_t *feed, const mugiq_b200_loop_plan_t *plan, void *dataPos_d) {
  const char *who = "mugiq_b200_loop_feed_set_plan";
  if (!feed || !plan || !dataPos_d) return set_error(MUGIQ_B200_EINVAL, "%s: NULL argument", who);
  if (feed->acquired >= 0 || feed->total > 0) return set_error(MUGIQ_B200_ESTATE, "%s: a run is in progress (call finish first)", who);
  const LoopPlan &pl = plan_of(plan);
  if (pl.g.volume != feed->volume || pl.precision != feed->precision)
    return set_error(MUGIQ_B200_EINVAL, "%s: the new plan lives on another lattice or precision", who);
  feed->pl = &pl;
  feed->dataPos_d = dataPos_d;
  return MUGIQ_B200_OK;
}

int mugiq_b200_loop_feed_destroy(mugiq_b200_loop_feed_t *f) {
  if (!f) return MUGIQ_B200_OK;
  if (f->compute || f->copy) cudaDeviceSynchronize();
  for (char *p : f->stage)
    if (p) cudaFree(p);
  if (f->site) cudaFree(f->site);
  for (cudaEvent_t e : f->freed)
    if (e) cudaEventDestroy(e);
  if (f->filled) cudaEventDestroy(f->filled);
  if (f->copy) cudaStreamDestroy(f->copy);
  delete f;
  return MUGIQ_B200_OK;
}

int mugiq_b200_loop_feed_acquire(mugiq_b200_loop_feed_t *feed, void **field_d, int n, void *producer_stream) {
  const char *who = "mugiq_b200_loop_feed_acquire";
  if (!feed || !field_d) return set_error(MUGIQ_B200_EINVAL, "%s: NULL argument", who);
  return feed_acquire(feed, field_d, n, (cudaStream_t)producer_stream, who);
}

int mugiq_b200_loop_feed_commit(mugiq_b200_loop_feed_t *feed, const double *sigma_h, int n, void *producer_stream) {
  const char *who = "mugiq_b200_loop_feed_commit";
  if (!feed || !sigma_h) return set_error(MUGIQ_B200_EINVAL, "%s: NULL argument", who);
  return feed_commit(feed, sigma_h, n, (cudaStream_t)producer_stream, who);
}

int mugiq_b200_loop_feed_push_host(mugiq_b200_loop_feed_t *feed, const void *const *evec_h, const double *sigma_h, int n) {
  const char *who = "mugiq_b200_loop_feed_push_host";
  if (!feed || !evec_h || !sigma_h) return set_error(MUGIQ_B200_EINVAL, "%s: NULL argument", who);
  if (n < 1) return set_error(MUGIQ_B200_EINVAL, "%s: n = %d must be positive", who, n);
  std::vector<void *> dst(feed->batch);
  for (int done = 0; done < n; done += feed->batch) {
    const int nb = std::min(feed->batch, n - done);
    int rc = feed_acquire(feed, dst.data(), nb, feed->copy, who);
    if (rc) return rc;
    for (int i = 0; i < nb; i++) {
      if (!evec_h[done + i]) return set_error(MUGIQ_B200_EINVAL, "%s: eigenvector %d is NULL", who, done + i);
      // fields that follow one another in host memory travel as one copy
      int run = 1;
      while (i + run < nb && static_cast<const char *>(evec_h[done + i + run]) ==
                                 static_cast<const char *>(evec_h[done + i]) + (size_t)run * feed->field_bytes)
        run++;
      MUGIQ_CUDA_CHECK(cudaMemcpyAsync(dst[i], evec_h[done + i], feed->field_bytes * run, cudaMemcpyHostToDevice, feed->copy));
      i += run - 1;
    }
    if ((rc = feed_commit(feed, sigma_h + done, nb, feed->copy, who))) return rc;
  }
  return MUGIQ_B200_OK;
}

int mugiq_b200_loop_feed_finish(mugiq_b200_loop_feed_t *feed, long long *nvec_total) {
  const char *who = "mugiq_b200_loop_feed_finish";
  if (!feed) return set_error(MUGIQ_B200_EINVAL, "%s: feed is NULL", who);
  if (feed->acquired >= 0) return set_error(MUGIQ_B200_ESTATE, "%s: a batch was acquired but not committed", who);
  if (nvec_total) *nvec_total = feed->total;
  // everything the feed enqueued is ordered on the compute stream; the staging batches may be reused by a new run
  feed->total = 0;
  return MUGIQ_B200_OK;
}

}  // extern "C"
