// interface_mugiq.cpp — public entry point computeLoop<Float> (/root/reference/lib/interface_mugiq.cpp:158-253), the
// driver-side parsers of the reference executable and the external-eigenpair registry.
#include <sys/sysinfo.h>

#include <fstream>
#include <sstream>

#include "host_util.h"
#include "loop_mugiq.h"

using namespace quda;

Eigsolve_Mugiq::Eigsolve_Mugiq(MugiqEigParam *eigParams_, const std::vector<ColorSpinorField *> &evecs, const std::vector<double> &sigma)
    : eigParams(eigParams_), eVecs(evecs), eVals_sigma(sigma.empty() ? nullptr : new std::vector<double>(sigma)) {
  if (!eigParams) errorQuda("Eigsolve_Mugiq: eigParams is NULL");
  // with a producer (setEvecProducer) one field suffices; Loop_Mugiq::computeCoarseLoop checks the count when it runs
  if (eVecs.empty()) errorQuda("Eigsolve_Mugiq: no eigenvector field given (at least the geometry reference is needed)");
}
Eigsolve_Mugiq::~Eigsolve_Mugiq() { delete eVals_sigma; }
void Eigsolve_Mugiq::printInfo() {
  printfQuda("Eigsolve_Mugiq: %d external eigenpairs, %s\n", eigParams->nEv, eVals_sigma ? "singular values present" : "no singular values");
}

static Eigsolve_Mugiq *g_external = nullptr;
void setExternalEigsolve(Eigsolve_Mugiq *e) { g_external = e; }
Eigsolve_Mugiq *getExternalEigsolve() { return g_external; }

// lib/interface_mugiq.cpp:158-172
template <typename Float, QudaFieldOrder fieldOrder> static void computeLoopOrdered(MugiqLoopParam loopParams, Eigsolve_Mugiq *eigsolve) {
  Loop_Mugiq<Float, fieldOrder> *loop = new Loop_Mugiq<Float, fieldOrder>(&loopParams, eigsolve);
  loop->computeCoarseLoop();
  if (loopParams.writeMomSpaceHDF5 || loopParams.writePosSpaceHDF5)
    loop->writeLoopsHDF5();
  else
    warningQuda("%s: Will NOT write output data!\n", __func__);
  delete loop;
}

template <typename Float>
void computeLoop(QudaMultigridParam, QudaEigParam eigParams, MugiqLoopParam loopParams, MuGiqBool computeCoarse, MuGiqBool useMG) {
  if (computeCoarse || useMG)
    errorQuda("%s: multigrid / coarse eigenvectors are external inputs of this build (QUDA's MG transfer is not part of the hot path)", __func__);
  Eigsolve_Mugiq *eigsolve = getExternalEigsolve();
  if (!eigsolve) errorQuda("%s: no eigenpairs registered (setExternalEigsolve): QUDA's eigensolver is an external input", __func__);
  if (eigParams.nEv > 0 && eigParams.nEv != eigsolve->getEigParams()->nEv)
    errorQuda("%s: eigParams.nEv = %d does not match the registered eigenpairs (%d)", __func__, eigParams.nEv, eigsolve->getEigParams()->nEv);
  eigsolve->printInfo();
  const QudaPrecision ePrec = eigsolve->getEvecs()[0]->Precision();
  if ((ePrec == QUDA_DOUBLE_PRECISION && sizeof(Float) != 8) || (ePrec == QUDA_SINGLE_PRECISION && sizeof(Float) != 4) ||
      (ePrec != QUDA_DOUBLE_PRECISION && ePrec != QUDA_SINGLE_PRECISION))
    errorQuda("%s: Incompatible precision between the eigenvectors (%d) and the template (%zu bytes)", __func__, (int)ePrec, sizeof(Float));
  switch (eigsolve->getEvecs()[0]->FieldOrder()) {
    case QUDA_FLOAT2_FIELD_ORDER: computeLoopOrdered<Float, QUDA_FLOAT2_FIELD_ORDER>(loopParams, eigsolve); break;
    case QUDA_FLOAT4_FIELD_ORDER: computeLoopOrdered<Float, QUDA_FLOAT4_FIELD_ORDER>(loopParams, eigsolve); break;
    case QUDA_SPACE_SPIN_COLOR_FIELD_ORDER: computeLoopOrdered<Float, QUDA_SPACE_SPIN_COLOR_FIELD_ORDER>(loopParams, eigsolve); break;
    default: errorQuda("%s: Unsupported Field order %d\n", __func__, (int)eigsolve->getEvecs()[0]->FieldOrder());
  }
}
template void computeLoop<double>(QudaMultigridParam, QudaEigParam, MugiqLoopParam, MuGiqBool, MuGiqBool);
template void computeLoop<float>(QudaMultigridParam, QudaEigParam, MugiqLoopParam, MuGiqBool, MuGiqBool);

// "--displace-entry-string" of the reference driver (tests/loop.cpp:607-718): entries separated by ';', each
// "<+-dir>:<start>[,<stop>]"
void parseDisplaceEntryString(MugiqLoopParam &prm, const std::string &entries) {
  if (entries.empty()) errorQuda("Got option '--loop-do-nonlocal yes' but option --displace-entry-string is not set!\n");
  prm.disp_entry.clear();
  prm.disp_str.clear();
  prm.disp_start.clear();
  prm.disp_stop.clear();
  std::stringstream all(entries);
  std::string ent;
  int idx = 0;
  while (std::getline(all, ent, ';')) {
    const size_t colon = ent.find(':');
    if (colon == std::string::npos || ent.find(':', colon + 1) != std::string::npos)
      errorQuda("Displacement entry %d has the Wrong format. Example of good entries: +z:1,8 , +x:3\n", idx);
    const std::string dir = ent.substr(0, colon), lim = ent.substr(colon + 1);
    std::vector<int> v;
    std::stringstream ls(lim);
    std::string tok;
    while (std::getline(ls, tok, ',')) {
      char *endp = nullptr;
      const long val = strtol(tok.c_str(), &endp, 10);
      if (tok.empty() || *endp != 0) errorQuda("Wrong format of displacement entry %d. Example of good entries: +z:1,8 , +x:3\n", idx);
      v.push_back((int)val);
    }
    if (v.empty() || v.size() > 2) errorQuda("Wrong format of displacement entry %d. Example of good entries: +z:1,8 , +x:3\n", idx);
    prm.disp_entry.push_back(ent);
    prm.disp_str.push_back(dir);
    prm.disp_start.push_back(v[0]);
    prm.disp_stop.push_back(v.size() == 2 ? v[1] : v[0]);
    idx++;
  }
  prm.doNonLocal = MUGIQ_BOOL_TRUE;
}

// momenta text file of the reference driver (tests/loop.cpp:723-746): three integers per line
void readMomentaFile(MugiqLoopParam &prm, const std::string &filename) {
  std::ifstream in(filename);
  if (!in) errorQuda("Cannot open momenta file %s\n", filename.c_str());
  prm.momMatrix.clear();
  std::string line;
  int n = 0;
  while (std::getline(in, line)) {
    if (line.find_first_not_of(" \t\r") == std::string::npos) continue;
    std::stringstream ls(line);
    int p[3];
    if (!(ls >> p[0] >> p[1] >> p[2])) errorQuda("Incorrect file format in Line %d\n", n);
    prm.momMatrix.push_back({p[0], p[1], p[2]});
    n++;
  }
  prm.Nmom = (int)prm.momMatrix.size();
  prm.doMomProj = prm.Nmom > 0 ? MUGIQ_BOOL_TRUE : MUGIQ_BOOL_FALSE;
}

// lib/util_mugiq.cpp:6-40
void printMemoryInfo() {
  struct sysinfo si;
  if (sysinfo(&si) == 0)
    printfQuda("CPU memory: total %.2f GB, free %.2f GB\n", si.totalram * (double)si.mem_unit / 1e9, si.freeram * (double)si.mem_unit / 1e9);
  long long fr = 0, tot = 0;
  if (mugiq_b200_device_info(nullptr, 0, nullptr, nullptr, &fr, &tot) == 0)
    printfQuda("GPU memory: total %.2f GB, free %.2f GB\n", tot / 1e9, fr / 1e9);
}
