// loop_driver.cpp — stand-in for the reference's `loop` executable (/root/reference/tests/loop.cpp:751-961) on top of
// the host mirror: same loop-related options (tests/test_params_mugiq.cpp:77-112), but eigenpairs and links are read
// from raw files (QUDA's eigensolver and gauge I/O are external inputs) instead of being computed / loaded by QUDA.
//
//   loop_driver --dim 4 4 4 8 --prec double --n-ev 16 --evecs-file ev.bin --sigma-file sig.bin --gauge-file u.bin
//               --loop-do-nonlocal yes --displace-entry-string "+z:1,3;-x:2" --loop-do-momproj yes
//               --momenta-filename mom.txt --loop-ft-sign minus --loop-write-mom-space yes
//               --loop-mom-space-filename loops.dat [--field-order site|float2|float4] [--dump-pos f] [--dump-mom f]
//   loop_driver --parse-only ...   parses and prints the loop parameters, touches no GPU (used by the CPU tests)
//   loop_driver --bench --dim 16 16 16 32 --n-ev 200 --displace-entry-string "+x:1;..." --bench-p2max 1 --bench-steps 3
//               end-to-end timing of the C++ front end on synthetic data: eigenvectors in PINNED HOST memory streamed
//               through Eigsolve_Mugiq::setEvecProducer (H2D inside the timed region), links uploaded every step, dataPos
//               and dataMom copied back; prints one JSON line with the fields of bench.py's `e2e` object
// File formats: evecs [nEv][V4 (even/odd order)][12] complex, gauge [4][V4][3][3] complex, sigma nEv doubles.
#include <cuda_runtime.h>

#include <cmath>
#include <complex>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>

#include "comm_mugiq.h"
#include "host_util.h"
#include "loop_mugiq.h"

using namespace quda;

static std::vector<char> slurp(const std::string &f, size_t expect) {
  std::ifstream in(f, std::ios::binary);
  if (!in) errorQuda("cannot open %s", f.c_str());
  std::vector<char> buf((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
  if (expect && buf.size() != expect) errorQuda("%s holds %zu bytes, expected %zu", f.c_str(), buf.size(), expect);
  return buf;
}
static bool yes(const std::string &s) { return s == "yes" || s == "true" || s == "1"; }

// site-major host field -> QUDA native order (for exercising the FLOAT2 / FLOAT4 ingest path)
template <typename Float> static void to_native(std::vector<char> &field, QudaFieldOrder order, size_t volumeCB) {
  if (order == QUDA_SPACE_SPIN_COLOR_FIELD_ORDER) return;
  const Float *src = reinterpret_cast<const Float *>(field.data());
  std::vector<char> out(field.size());
  Float *dst = reinterpret_cast<Float *>(out.data());
  const int N = order == QUDA_FLOAT2_FIELD_ORDER ? 2 : 4;
  for (int pty = 0; pty < 2; pty++)
    for (size_t x = 0; x < volumeCB; x++)
      for (int k = 0; k < 24; k++)
        dst[pty * volumeCB * 24 + ((size_t)(k / N) * volumeCB + x) * N + k % N] = src[(pty * volumeCB + x) * 24 + k];
  field.swap(out);
}

// ---- eigenvectors that live in host memory: producer for Eigsolve_Mugiq::setEvecProducer -------------------------
struct HostEvecs {
  const char *base;
  size_t fieldBytes;
};
static void copy_from_host(void *ctx, int n, ColorSpinorField *fine, void *stream) {
  const HostEvecs *h = static_cast<const HostEvecs *>(ctx);
  HOST_CUDA(cudaMemcpyAsync(fine->V(), h->base + (size_t)n * h->fieldBytes, h->fieldBytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
}

template <typename Float> static int run(std::map<std::string, std::string> &opt, const int Xg[4], MugiqLoopParam &prm, int nEv,
                                         QudaFieldOrder order) {
  // --tsplit N: the files hold the GLOBAL lattice Xg; this rank (--comm-rank, 0 for N = 1) keeps the time slab
  // [rank*T/N, (rank+1)*T/N) of every eigenvector and the whole (replicated) gauge field
  const int tsN = opt.count("--tsplit") ? atoi(opt["--tsplit"].c_str()) : 0;
  const int tsRank = tsN > 1 ? mugiqCommRank(getLoopComm()) : 0;
  int X[4] = {Xg[0], Xg[1], Xg[2], Xg[3]};
  if (tsN > 0) {
    if (order != QUDA_SPACE_SPIN_COLOR_FIELD_ORDER) errorQuda("--tsplit takes site-major eigenvectors");
    if (Xg[3] % tsN) errorQuda("T = %d is not divisible by %d time ranks", Xg[3], tsN);
    if (tsN != mugiqCommSize(getLoopComm())) errorQuda("--tsplit %d needs --comm-size %d", tsN, tsN);
    X[3] = Xg[3] / tsN;
    setTimePartition(tsN, tsRank);
    setLoopTSplit(true);
  }
  const size_t V4g = (size_t)Xg[0] * Xg[1] * Xg[2] * Xg[3];
  const size_t V4 = (size_t)X[0] * X[1] * X[2] * X[3];
  const QudaPrecision prec = precision_of<Float>();
  const size_t fieldBytesG = V4g * 24 * sizeof(Float);
  const size_t fieldBytes = V4 * 24 * sizeof(Float);
  std::vector<char> ev = slurp(opt["--evecs-file"], fieldBytesG * nEv);
  std::vector<char> sg = slurp(opt["--sigma-file"], sizeof(double) * nEv);
  std::vector<double> sigma(nEv);
  memcpy(sigma.data(), sg.data(), sg.size());
  std::vector<char> gauge;
  QudaGaugeParam gp;
  if (prm.doNonLocal) {
    gauge = slurp(opt["--gauge-file"], 4 * V4g * 18 * sizeof(Float));
    for (int i = 0; i < 4; i++) {
      gp.X[i] = X[i];
      prm.gauge[i] = gauge.data() + (size_t)i * V4g * 18 * sizeof(Float);
    }
    gp.cpu_prec = gp.cuda_prec = prec;
    gp.gauge_order = QUDA_QDP_GAUGE_ORDER;
    prm.gauge_param = &gp;
  }
  // eigenvector shards: this rank keeps [lo, hi) of the file's eigenpairs
  int lo = 0, hi = nEv;
  if (getLoopComm() && tsN == 0) mugiqCommShard(nEv, mugiqCommRank(getLoopComm()), mugiqCommSize(getLoopComm()), &lo, &hi);
  if (hi <= lo) errorQuda("rank %d got an empty eigenvector shard (%d eigenvectors over %d ranks)", mugiqCommRank(getLoopComm()), nEv, mugiqCommSize(getLoopComm()));
  {
    std::vector<double> shard(sigma.begin() + lo, sigma.begin() + hi);
    sigma.swap(shard);
  }
  // --stream-evecs yes: the eigenvectors stay in HOST memory and reach the loop through the producer hook
  // (Eigsolve_Mugiq::setEvecProducer, the stand-in for prolongateEvec): one device field only, as geometry reference
  const bool streamEvecs = opt.count("--stream-evecs") && yes(opt["--stream-evecs"]);
  std::vector<ColorSpinorField *> fields;
  std::vector<char> hostFields;
  ColorSpinorParam cs;
  for (int i = 0; i < 4; i++) cs.x[i] = X[i];
  cs.precision = prec;
  cs.fieldOrder = order;
  for (int n = lo; n < hi; n++) {
    std::vector<char> one;
    if (tsN > 0) {  // this rank's time-slices of both parities (a time-slice is one contiguous block per parity)
      const size_t V3h = (size_t)X[0] * X[1] * X[2] / 2, slice = V3h * 24 * sizeof(Float);
      one.resize(fieldBytes);
      for (int p = 0; p < 2; p++)
        memcpy(one.data() + (size_t)p * X[3] * slice, ev.data() + n * fieldBytesG + ((size_t)p * Xg[3] + (size_t)tsRank * X[3]) * slice,
               (size_t)X[3] * slice);
    } else {
      one.assign(ev.begin() + n * fieldBytes, ev.begin() + (n + 1) * fieldBytes);
    }
    to_native<Float>(one, order, V4 / 2);
    if (streamEvecs) {
      hostFields.insert(hostFields.end(), one.begin(), one.end());
      if (fields.empty()) fields.push_back(ColorSpinorField::Create(cs));
      continue;
    }
    fields.push_back(ColorSpinorField::Create(cs));
    HOST_CUDA(cudaMemcpy(fields.back()->V(), one.data(), fieldBytes, cudaMemcpyHostToDevice));
  }
  QudaEigParam qe;
  qe.nEv = hi - lo;
  MugiqEigParam ep(&qe);
  Eigsolve_Mugiq eigsolve(&ep, fields, sigma);
  HostEvecs he{hostFields.data(), fieldBytes};
  if (streamEvecs) {
    if (tsN > 0) errorQuda("--stream-evecs is not available with --tsplit");
    eigsolve.setEvecProducer(copy_from_host, &he, opt.count("--stream-batch") ? atoi(opt["--stream-batch"].c_str()) : 4);
  }

  if (opt.count("--dump-pos") || opt.count("--dump-mom")) {
    // class-level use (what computeLoop<Float,order> does), keeping the object to read its buffers
    auto dump = [&](auto *loop) {
      loop->computeCoarseLoop();
      if (prm.writeMomSpaceHDF5 || prm.writePosSpaceHDF5) loop->writeLoopsHDF5();
      if (opt.count("--dump-pos")) std::ofstream(opt["--dump-pos"], std::ios::binary).write((const char *)loop->hostDataPos(), loop->numElemPos() * 2 * sizeof(Float));
      if (opt.count("--dump-mom") && prm.doMomProj) std::ofstream(opt["--dump-mom"], std::ios::binary).write((const char *)loop->hostDataMom(), loop->numElemMom() * 2 * sizeof(Float));
      delete loop;
    };
    if (order == QUDA_FLOAT2_FIELD_ORDER) dump(new Loop_Mugiq<Float, QUDA_FLOAT2_FIELD_ORDER>(&prm, &eigsolve));
    else if (order == QUDA_FLOAT4_FIELD_ORDER) dump(new Loop_Mugiq<Float, QUDA_FLOAT4_FIELD_ORDER>(&prm, &eigsolve));
    else dump(new Loop_Mugiq<Float, QUDA_SPACE_SPIN_COLOR_FIELD_ORDER>(&prm, &eigsolve));
  } else {
    setExternalEigsolve(&eigsolve);
    QudaMultigridParam mg;
    mg.n_level = 0;
    computeLoop<Float>(mg, qe, prm, MUGIQ_BOOL_FALSE, MUGIQ_BOOL_FALSE);
    setExternalEigsolve(nullptr);
  }
  for (ColorSpinorField *f : fields) delete f;
  return 0;
}

// ---- --bench -----------------------------------------------------------------------------------------------------------
static int run_bench(std::map<std::string, std::string> &opt, const int X[4], MugiqLoopParam &prm, int nEv) {
  typedef double Float;
  const size_t V4 = (size_t)X[0] * X[1] * X[2] * X[3];
  const size_t fieldBytes = V4 * 24 * sizeof(Float);
  const int steps = opt.count("--bench-steps") ? atoi(opt["--bench-steps"].c_str()) : 3;
  const int batch = opt.count("--bench-batch") ? atoi(opt["--bench-batch"].c_str()) : 16;
  const int p2max = opt.count("--bench-p2max") ? atoi(opt["--bench-p2max"].c_str()) : 0;
  // momenta |p|^2 <= p2max in lexicographic order (bench.py: momenta_up_to)
  prm.momMatrix.clear();
  const int pm = (int)std::ceil(std::sqrt((double)p2max));
  for (int px = -pm; px <= pm; px++)
    for (int py = -pm; py <= pm; py++)
      for (int pz = -pm; pz <= pm; pz++)
        if (px * px + py * py + pz * pz <= p2max) prm.momMatrix.push_back({px, py, pz});
  prm.Nmom = (int)prm.momMatrix.size();
  prm.doMomProj = MUGIQ_BOOL_TRUE;
  prm.FTSign = LOOP_FT_SIGN_MINUS;
  // synthetic inputs: eigenvectors (unit norm) and links in pinned host memory
  char *ev_h = nullptr;
  HOST_CUDA(cudaMallocHost((void **)&ev_h, fieldBytes * nEv));
  std::vector<double> sigma(nEv);
  double sumInvSigma = 0;
  {
    uint64_t s = 0x6d75676971ULL;
    auto rnd = [&s]() {
      s = s * 6364136223846793005ULL + 1442695040888963407ULL;
      return (double)((s >> 11) & ((1ULL << 53) - 1)) / (double)(1ULL << 53) - 0.5;
    };
    for (int n = 0; n < nEv; n++) {
      Float *v = reinterpret_cast<Float *>(ev_h + (size_t)n * fieldBytes);
      double nrm = 0;
      for (size_t i = 0; i < V4 * 24; i++) {
        v[i] = rnd();
        nrm += v[i] * v[i];
      }
      const double inv = 1.0 / std::sqrt(nrm);
      for (size_t i = 0; i < V4 * 24; i++) v[i] *= inv;
      sigma[n] = 0.01 + 0.001 * n;
      sumInvSigma += 1.0 / sigma[n];
    }
  }
  Float *gauge = nullptr;  // pinned as well: a staged copy from pageable memory costs 5 ms per step
  const size_t gaugeLen = 4 * V4 * 18;
  QudaGaugeParam gp;
  if (prm.doNonLocal) {
    HOST_CUDA(cudaMallocHost((void **)&gauge, gaugeLen * sizeof(Float)));
    uint64_t s = 12345;
    for (size_t k = 0; k < gaugeLen; k++) {
      s = s * 6364136223846793005ULL + 1442695040888963407ULL;
      gauge[k] = (double)((s >> 11) & ((1ULL << 53) - 1)) / (double)(1ULL << 53) - 0.5;
    }
    for (int i = 0; i < 4; i++) {
      gp.X[i] = X[i];
      prm.gauge[i] = gauge + (size_t)i * V4 * 18;
    }
    gp.cpu_prec = gp.cuda_prec = QUDA_DOUBLE_PRECISION;
    gp.gauge_order = QUDA_QDP_GAUGE_ORDER;
    prm.gauge_param = &gp;
  }
  ColorSpinorParam cs;
  for (int i = 0; i < 4; i++) cs.x[i] = X[i];
  cs.precision = QUDA_DOUBLE_PRECISION;
  cs.fieldOrder = QUDA_SPACE_SPIN_COLOR_FIELD_ORDER;
  std::vector<ColorSpinorField *> fields{ColorSpinorField::Create(cs)};  // the geometry reference (lib/loop_mugiq.cpp:42-43)
  QudaEigParam qe;
  qe.nEv = nEv;
  MugiqEigParam ep(&qe);
  Eigsolve_Mugiq eigsolve(&ep, fields, sigma);
  HostEvecs he{ev_h, fieldBytes};
  eigsolve.setEvecProducer(copy_from_host, &he, batch);
  Loop_Mugiq<Float, QUDA_SPACE_SPIN_COLOR_FIELD_ORDER> loop(&prm, &eigsolve);
  cudaEvent_t e0, e1;
  HOST_CUDA(cudaEventCreate(&e0));
  HOST_CUDA(cudaEventCreate(&e1));
  auto step = [&]() {
    loop.resetRun();           // H2D of the links
    loop.computeCoarseLoop();  // H2D of every eigenvector (streamed), kernels, D2H of dataPos and dataMom
  };
  step();  // warm-up
  HOST_CUDA(cudaDeviceSynchronize());
  HOST_CUDA(cudaEventRecord(e0, nullptr));
  for (int i = 0; i < steps; i++) step();
  HOST_CUDA(cudaEventRecord(e1, nullptr));
  HOST_CUDA(cudaDeviceSynchronize());
  float ms = 0;
  HOST_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  ms /= steps;
  // checksum of the result: sum_x T_1(x) = sum_n |v_n|^2 / sigma_n
  std::complex<double> tr = 0;
  for (size_t x = 0; x < V4; x++) tr += loop.hostDataPos()[x];
  const double chk = std::abs(tr - sumInvSigma) / sumInvSigma;
  const double h2d = (double)fieldBytes * nEv + (prm.doNonLocal ? (double)gaugeLen * sizeof(Float) : 0.0);
  const double d2h = (double)(loop.numElemPos() + loop.numElemMom()) * 2 * sizeof(Float);
  const double units = (double)nEv * V4 * loop.nLoop();
  printf("{\"value\": %.6e, \"unit\": \"eigvec*site*loop contractions/s (16 gamma each)\", \"ms_per_step\": %.4f, \"steps\": %d, "
         "\"h2d_bytes_per_step\": %.0f, \"d2h_bytes_per_step\": %.0f, \"h2d_GBps\": %.2f, \"stream_batch\": %d, \"nLoop\": %d, \"Nmom\": %d, "
         "\"checksum_rel_err\": %.3e, \"front_end\": \"C++ host mirror (Loop_Mugiq<double> + Eigsolve_Mugiq::setEvecProducer over "
         "mugiq_b200_loop_feed_*), eigenvectors in pinned host memory\"}\n",
         units / (ms * 1e-3), ms, steps, h2d, d2h, h2d / (ms * 1e-3) / 1e9, batch, loop.nLoop(), prm.Nmom, chk);
  fflush(stdout);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  delete fields[0];
  cudaFreeHost(ev_h);
  return chk < 1e-10 ? 0 : 2;
}

int main(int argc, char **argv) {
  std::map<std::string, std::string> opt;
  int X[4] = {0, 0, 0, 0};
  bool parseOnly = false, bench = false;
  for (int i = 1; i < argc; i++) {
    const std::string a = argv[i];
    if (a == "--parse-only") {
      parseOnly = true;
    } else if (a == "--bench") {
      bench = true;
    } else if (a == "--dim") {
      if (i + 4 >= argc) errorQuda("--dim needs four extents");
      for (int d = 0; d < 4; d++) X[d] = atoi(argv[++i]);
    } else if (a.rfind("--", 0) == 0) {
      if (i + 1 >= argc) errorQuda("option %s needs a value", a.c_str());
      opt[a] = argv[++i];
    } else {
      errorQuda("unexpected argument %s", a.c_str());
    }
  }
  if (opt.count("--verbosity") && opt["--verbosity"] == "verbose") setVerbosityQuda(QUDA_VERBOSE);
  MugiqLoopParam prm;
  if (bench && opt.count("--displace-entry-string") && !opt["--displace-entry-string"].empty()) opt["--loop-do-nonlocal"] = "yes";
  if (opt.count("--loop-do-nonlocal") && yes(opt["--loop-do-nonlocal"])) parseDisplaceEntryString(prm, opt["--displace-entry-string"]);
  if (opt.count("--loop-do-momproj") && yes(opt["--loop-do-momproj"])) {
    if (!opt.count("--momenta-filename")) errorQuda("Got option '--loop-do-momproj yes' but option --momenta-filename is not set!\n");
    readMomentaFile(prm, opt["--momenta-filename"]);
  }
  if (opt.count("--loop-ft-sign")) {
    const std::string s = opt["--loop-ft-sign"];
    if (s == "plus") prm.FTSign = LOOP_FT_SIGN_PLUS;
    else if (s == "minus") prm.FTSign = LOOP_FT_SIGN_MINUS;
    else errorQuda("Unknown --loop-ft-sign %s (plus/minus)", s.c_str());
  }
  if (opt.count("--loop-write-mom-space")) prm.writeMomSpaceHDF5 = yes(opt["--loop-write-mom-space"]) ? MUGIQ_BOOL_TRUE : MUGIQ_BOOL_FALSE;
  if (opt.count("--loop-write-pos-space")) prm.writePosSpaceHDF5 = yes(opt["--loop-write-pos-space"]) ? MUGIQ_BOOL_TRUE : MUGIQ_BOOL_FALSE;
  if (opt.count("--loop-mom-space-filename")) prm.fname_mom_h5 = opt["--loop-mom-space-filename"];
  if (opt.count("--loop-pos-space-filename")) prm.fname_pos_h5 = opt["--loop-pos-space-filename"];
  if (parseOnly) {
    printf("nonlocal %d entries %zu\n", (int)prm.doNonLocal, prm.disp_str.size());
    for (size_t i = 0; i < prm.disp_str.size(); i++) printf("entry %zu %s %d %d\n", i, prm.disp_str[i].c_str(), prm.disp_start[i], prm.disp_stop[i]);
    printf("momproj %d Nmom %d ftsign %d\n", (int)prm.doMomProj, prm.Nmom, (int)prm.FTSign);
    for (auto &p : prm.momMatrix) printf("mom %d %d %d\n", p[0], p[1], p[2]);
    return 0;
  }
  const int nEv = opt.count("--n-ev") ? atoi(opt["--n-ev"].c_str()) : 0;
  if (nEv < 1) errorQuda("--n-ev must be positive");
  if (bench) {
    if (opt.count("--device")) HOST_CUDA(cudaSetDevice(atoi(opt["--device"].c_str())));
    setVerbosityQuda(QUDA_SILENT);
    return run_bench(opt, X, prm, nEv);
  }
  QudaFieldOrder order = QUDA_SPACE_SPIN_COLOR_FIELD_ORDER;
  if (opt.count("--field-order")) {
    if (opt["--field-order"] == "float2") order = QUDA_FLOAT2_FIELD_ORDER;
    else if (opt["--field-order"] == "float4") order = QUDA_FLOAT4_FIELD_ORDER;
    else if (opt["--field-order"] != "site") errorQuda("Unknown --field-order %s", opt["--field-order"].c_str());
  }
  // one process per GPU: --comm-size N --comm-rank r --comm-id-file f [--device d]
  MugiqComm *comm = nullptr;
  if (opt.count("--comm-size") && atoi(opt["--comm-size"].c_str()) > 1) {
    if (!opt.count("--comm-rank") || !opt.count("--comm-id-file")) errorQuda("--comm-size needs --comm-rank and --comm-id-file");
    const int rank = atoi(opt["--comm-rank"].c_str());
    comm = mugiqCommInit(rank, atoi(opt["--comm-size"].c_str()), opt.count("--device") ? atoi(opt["--device"].c_str()) : rank,
                         opt["--comm-id-file"].c_str());
    setLoopComm(comm);
    // --peer-reduce: the overlapped position-space sum moves its chunks over peer-mapped buffers (copy engines)
    if (opt.count("--peer-reduce") && yes(opt["--peer-reduce"])) setenv("MUGIQ_B200_PEER_REDUCE", "1", 1);
  } else if (opt.count("--device")) {
    HOST_CUDA(cudaSetDevice(atoi(opt["--device"].c_str())));
  }
  const std::string prec = opt.count("--prec") ? opt["--prec"] : "double";
  int rc = 1;
  if (prec == "double") rc = run<double>(opt, X, prm, nEv, order);
  else if (prec == "single") rc = run<float>(opt, X, prm, nEv, order);
  else errorQuda("Unknown --prec %s (double/single)", prec.c_str());
  setLoopTSplit(false);
  setLoopComm(nullptr);
  mugiqCommFinalize(comm);
  return rc;
}
