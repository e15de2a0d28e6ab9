// mugiq.h — the public header of the mirror, under the name of the reference's (/root/reference/include/mugiq.h): enums,
// macros, gamma names, the parameter struct and computeLoop of MuGiq's loop interface, kept numerically and by name
// compatible with the reference so that its drivers compile against this mirror.  One header holds what the reference
// spreads over mugiq.h, enum_mugiq.h, util_mugiq.h, gamma.h and interface_mugiq.h:
//   enums                 /root/reference/include/enum_mugiq.h:12-92
//   MugiqLoopParam        /root/reference/include/mugiq.h:28-47
//   size / index macros   /root/reference/include/util_mugiq.h:7-27
//   GammaName             /root/reference/include/gamma.h:11-20  (the tables themselves live in the CUDA library,
//                          mugiq_b200_gamma_tables())
//   checkGauge            /root/reference/include/interface_mugiq.h:9 (forward declaration of QUDA's)
#ifndef MUGIQ_B200_HOST_MUGIQ_H
#define MUGIQ_B200_HOST_MUGIQ_H

#include <climits>
#include <string>
#include <vector>

#include "quda_shim.h"

#define MUGIQ_INVALID_ENUM INT_MIN

typedef enum MuGiqBool_s { MUGIQ_BOOL_FALSE = 0, MUGIQ_BOOL_TRUE = 1, MUGIQ_BOOL_INVALID = MUGIQ_INVALID_ENUM } MuGiqBool;
typedef enum MuGiqTask_s {
  MUGIQ_COMPUTE_EVECS_QUDA,
  MUGIQ_COMPUTE_EVECS_MUGIQ,
  MUGIQ_COMPUTE_LOOP,
  MUGIQ_TASK_INVALID = MUGIQ_INVALID_ENUM
} MuGiqTask;
typedef enum MuGiqEigOperator_s {
  MUGIQ_EIG_OPERATOR_M,
  MUGIQ_EIG_OPERATOR_Mdag,
  MUGIQ_EIG_OPERATOR_MdagM,
  MUGIQ_EIG_OPERATOR_MMdag,
  MUGIQ_EIG_OPERATOR_INVALID = MUGIQ_INVALID_ENUM
} MuGiqEigOperator;
typedef enum LoopFTSign_s { LOOP_FT_SIGN_MINUS = -1, LOOP_FT_SIGN_PLUS = 1, LOOP_FT_SIGN_INVALID = MUGIQ_INVALID_ENUM } LoopFTSign;
typedef enum LoopCalcType_s {
  LOOP_CALC_TYPE_BLAS,
  LOOP_CALC_TYPE_OPT_KERNEL,
  LOOP_CALC_TYPE_BASIC_KERNEL,
  LOOP_CALC_TYPE_INVALID = MUGIQ_INVALID_ENUM
} LoopCalcType;
typedef enum DisplaceType_s { DISPLACE_TYPE_COVARIANT = 0, DISPLACE_TYPE_INVALID = MUGIQ_INVALID_ENUM } DisplaceType;
// "+x" "-x" "+y" "-y" "+z" "-z" "+t" "-t" in this order: flag = 2*dir + (sign == minus)
typedef enum DisplaceFlag_s {
  DispFlagNone = MUGIQ_INVALID_ENUM,
  DispFlag_X = 0, DispFlag_x = 1, DispFlag_Y = 2, DispFlag_y = 3, DispFlag_Z = 4, DispFlag_z = 5, DispFlag_T = 6, DispFlag_t = 7
} DisplaceFlag;
typedef enum DisplaceDir_s { DispDirNone = MUGIQ_INVALID_ENUM, DispDir_x = 0, DispDir_y = 1, DispDir_z = 2, DispDir_t = 3 } DisplaceDir;
typedef enum DisplaceSign_s { DispSignNone = MUGIQ_INVALID_ENUM, DispSignMinus = 0, DispSignPlus = 1 } DisplaceSign;
typedef enum MuGiqBoundaryDirection_s {
  MUGIQ_BOUNDARY_BACKWARD = 0,
  MUGIQ_BOUNDARY_FORWARD = 1,
  MUGIQ_BOUNDARY_INVALID = MUGIQ_INVALID_ENUM
} MuGiqBoundaryDirection;

// sizes and index orders of the loop buffers
#define N_DIM_ 4
#define MOM_DIM_ 3
#define N_SPIN_ 4
#define N_COLOR_ 3
#define N_GAMMA_ 16
#define N_DISPLACE_TYPES 1
#define N_DISPLACE_SIGNS 2
#define SPINOR_SITE_LEN_ (N_SPIN_ * N_COLOR_)
#define GAMMA_LEN_ (N_SPIN_ * N_SPIN_)
#define SPINOR_SITE_IDX(s, c) ((c) + N_COLOR_ * (s))
#define GAMMA_MAT_IDX(r, c) ((c) + N_SPIN_ * (r))
#define MOM_MATRIX_IDX(id, im) ((id) + MOM_DIM_ * (im))

// name of the current each of the 16 gamma indices is stored under (output order, i.e. after the g5*Gamma map)
inline const std::vector<std::string> &GammaName() {
  static const std::vector<std::string> names{"1",  "g1",   "g2",   "g1g2", "g3",   "g1g3", "g2g3", "g5g4",
                                              "g4", "g1g4", "g2g4", "g5g3", "g3g4", "g5g2", "g5g1", "g5"};
  return names;
}

// Parameters of a loop calculation.  Passed by value to computeLoop and by pointer to the constructors, which copy
// what they need; `gauge` (host links in QDP order, one pointer per direction) and `gauge_param` are borrowed.
typedef struct MugiqLoopParam_s {
  int Nmom = 0;
  std::vector<std::vector<int>> momMatrix;  // [Nmom][3]
  LoopFTSign FTSign = LOOP_FT_SIGN_MINUS;
  LoopCalcType calcType = LOOP_CALC_TYPE_OPT_KERNEL;
  MuGiqBool writeMomSpaceHDF5 = MUGIQ_BOOL_FALSE;
  MuGiqBool writePosSpaceHDF5 = MUGIQ_BOOL_FALSE;
  MuGiqBool doMomProj = MUGIQ_BOOL_FALSE;
  MuGiqBool doNonLocal = MUGIQ_BOOL_FALSE;
  std::vector<std::string> disp_entry;  // e.g. "+z:1,8"
  std::vector<std::string> disp_str;    // e.g. "+z"
  std::string fname_mom_h5;
  std::string fname_pos_h5;
  std::vector<int> disp_start;
  std::vector<int> disp_stop;
  void *gauge[4] = {nullptr, nullptr, nullptr, nullptr};
  QudaGaugeParam *gauge_param = nullptr;
} MugiqLoopParam;

// The reference's public entry point (include/mugiq.h:79-81).  QUDA's eigensolver is an external input of the hot
// path: here the eigenpairs come from the Eigsolve_Mugiq object registered with setExternalEigsolve()
// (eigsolve_mugiq.h) instead of being computed from mgParams / eigParams.
template <typename Float>
void computeLoop(QudaMultigridParam mgParams, QudaEigParam eigParams, MugiqLoopParam loopParams, MuGiqBool computeCoarse,
                 MuGiqBool useMG);

// Driver-side parsers of the reference executable (tests/loop.cpp:607-746): "--displace-entry-string" grammar
// "<+-dir>:<start>[,<stop>];..." and the momenta text file (three ints per line).
void parseDisplaceEntryString(MugiqLoopParam &prm, const std::string &entries);
void readMomentaFile(MugiqLoopParam &prm, const std::string &filename);

void printMemoryInfo();

namespace quda {
bool checkGauge(const QudaGaugeParam *param);  // extents positive and even in x, precision single or double
}

#endif
