// prof.cuh — per-kernel launch counters and (optional) CUDA-event timers around every kernel launch of the
// library.  The reference only brackets init/total/free with QUDA TimeProfile
// (/root/reference/lib/interface_mugiq.cpp:36-47,193-244); here every launch site is wrapped in a
// ProfScope so that bench.py can report the dominant kernel's average launch duration measured live, on the
// stream the kernel is launched on, over the timed region (mugiq_b200_prof_* in include/mugiq_b200.h).
#pragma once
#include <cuda_runtime.h>

namespace mugiq_b200 {

enum KernelId {
  K_CONTRACT = 0,     // contract_batch_kernel
  K_DISPLACE,         // displace_kernel
  K_LOOP_FUSED,       // loop_fused_kernel (all loops of an eigenvector batch in one pass)
  K_WILSON_LINE,      // wilson_line_kernel (gauge-only products of links)
  K_MINUS_FROM_PLUS,  // minus_from_plus_kernel
  K_REORDER,          // reorder_mapgamma_kernel
  K_PHASE,            // phase_matrix_kernel
  K_MOMPROJ,          // momproj_*_kernel
  K_SPLITK_REDUCE,    // splitk_reduce_kernel
  K_CONVERT,          // convert_spinor_kernel
  K_HALO_PUSH,        // halo_push_kernel / copy-engine 2-D copy into a time neighbour's slabs (peer.cu)
  K_ALLREDUCE,        // NCCL all-reduce of (a time-slice range of) the loop buffer (comm.cu); NCCL's kernel, not ours
  K_COUNT
};

const char *kernel_name(int id);
// Called by launch sites.  Always counts the launch; records events only while profiling is enabled.
void prof_begin(int id, cudaStream_t stream, double alg_bytes, double alg_flops);
void prof_end(int id, cudaStream_t stream);

struct ProfScope {
  int id;
  cudaStream_t stream;
  ProfScope(int id_, cudaStream_t s, double alg_bytes = 0.0, double alg_flops = 0.0) : id(id_), stream(s) {
    prof_begin(id, s, alg_bytes, alg_flops);
  }
  ~ProfScope() { prof_end(id, stream); }
};

}  // namespace mugiq_b200
