// C ABI of the B200-native ABC-OCT reconstruction path (see include/abcoct.h for the reference citations).
// Host side only: parameter / .ini handling, the bit-exact lambda->k tables, calibration state, device
// buffers, the pinned ingest ring on CUDA streams and the multi-GPU sharding by B-scan.
#include "../../include/abcoct.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <thread>
#include <vector>

#include "kernels.h"

using namespace abcoct;

namespace {

thread_local std::string g_create_error;

constexpr int kTimedChunks = 512;  // timing-event pool size (chunks timed between two abcoct_timing_reset calls)
constexpr int kSlots = 3;  // pinned-ring depth per GPU (>= 3 streams per GPU)

// the images one call can produce (abcoct_outputs), bytes per pixel of each
enum { O_U8 = 0, O_DB, O_LIN, O_BGR, O_JSUB, O_JBGR, O_COUNT };
constexpr size_t kOutBpp[O_COUNT] = {1, 4, 4, 3, 1, 3};
struct OutPtrs {
  void* p[O_COUNT] = {};
  unsigned mask() const {
    unsigned m = 0;
    for (int k = 0; k < O_COUNT; ++k) m |= p[k] ? 1u << k : 0u;
    return m;
  }
};

struct GpuState {
  int dev = 0;
  int sm_count = 0;
  cudaStream_t stream[kSlots] = {};
  cudaEvent_t slot_done[kSlots] = {};
  std::vector<cudaEvent_t> tev;  // timing-event pool: 3 events per timed chunk (before recon, between, after normalise)
  size_t tev_used = 0;
  unsigned char* d_tables = nullptr;
  float* d_gain = nullptr;
  float* d_subg = nullptr;
  size_t cal_floats = 0;  // capacity of d_gain / d_subg
  // general pre-processing path (prep_kernels.cu): unswizzled f32 calibration, window, FFT twiddles, per-slot work buffers
  float *d_yb = nullptr, *d_yp = nullptr, *d_yd = nullptr, *d_win = nullptr;
  float2 *d_twW = nullptr, *d_twM = nullptr;
  void* d_med[kSlots] = {};
  void* d_bin[kSlots] = {};
  float* d_rows[kSlots] = {};
  float* d_fmm[kSlots] = {};
  // normalised-calibration regime (single_row_regime): calibration, window and lerp weights as doubles, f64 rows per slot
  double *d_yb64 = nullptr, *d_yp64 = nullptr, *d_yd64 = nullptr, *d_win64 = nullptr, *d_gwq64 = nullptr;
  double* d_rows64[kSlots] = {};
  float* d_pre32[kSlots] = {};
  long long* d_fmm64[kSlots] = {};
  int* d_gidx = nullptr;  // generic kernel: gather indices, weights, exp(+2 pi i k / N)
  float* d_gwq = nullptr;
  float2* d_twN = nullptr;
  size_t prep_frames[kSlots] = {};
  // per-slot device + pinned staging for the host-buffer API
  uint8_t* d_in[kSlots] = {};
  uint8_t* h_in[kSlots] = {};
  void* d_o[O_COUNT][kSlots] = {};  // device / pinned staging of every requested output image
  void* h_o[O_COUNT][kSlots] = {};
  // J0 lock-in: the reference B-scan, per-B-scan min/max words, and temporaries of the device entry point
  float* d_jscan = nullptr;
  int* d_jmm[kSlots] = {};
  float* d_dc01[kSlots] = {};   // dB of bins 0, 1 before the DC-row mask, [chunk B-scans][oph][2]
  void* d_tmp[2][kSlots] = {};  // [0] subtracted u8, [1] dB image when needed but not asked for by the caller
  size_t tmp_bytes[2][kSlots] = {};
  float* d_scratch[kSlots] = {};
  int* d_sched[kSlots] = {};
  size_t scratch_bscans[kSlots] = {};
  size_t slot_in_bytes = 0, slot_out_px = 0;  // capacity of the per-slot staging buffers
  unsigned slot_mask = 0;                      // which output images the slots hold
};

}  // namespace

struct abcoct_ctx {
  abcoct_params p{};
  int opw = 0, oph = 0, M = 0, N = 0, D = 0, A = 1;
  std::vector<int32_t> nk;
  std::vector<double> frac, win;
  std::vector<double> yb, yp, yd, yr, ys;
  std::vector<float> jscan;  // jscansave (BscanFFT.cpp:572, 1294): D x oph linear B-scan, empty = lock-in off
  bool jscan_dirty = false;
  bool have_yr = false, have_ys = false;
  bool have_yb = false, have_yp = false, have_yd = false, cal_dirty = true, has_sub = false;
  bool general = false;  // any optional pre-processing stage is on: frames go through prep_kernels.cu first
  // BscanFFTspinjnt.cpp:1856-1862 with bscanbinx = bscanbiny = binvaluey = 1 < binvaluex: both resizes are copies, what is left is
  // bscan *= multiplyfactor (= binvaluex, :835) before the log, i.e. + (20 / 2.303) ln(multiplyfactor) dB on the dB image; the display
  // image sees the threshold and the clamp value lowered by the same amount (min-max normalisation is shift invariant)
  float db_shift = 0.f;
  bool generic = false;  // no fused plan for this N / row width / D: generic_recon_kernel on the prepared rows (implies general)
  std::vector<int> radN;
  // warp-per-A-scan kernel (wrow_kernel.cuh): eligible configurations and the plan in use (nullptr: recon_kernel.cuh)
  bool wrow_eligible = false;
  const WPlanEntry* wplan = nullptr;
  std::vector<unsigned char> blob1, wblob, rblob;  // table blobs of the group kernel, the scratch warp kernel, the resident-row kernel
  const std::vector<unsigned char>* blob_loaded = nullptr;  // which blob d_tables holds
  std::vector<int> gidx;      // the kernel's remapped gather indices / weights (debug tap)
  std::vector<float> gwq;
  bool single = false;        // single_row_regime(p): f64 row preparation + f64 lerp, one row per transform
  int px_bytes = 2;      // bytes per pixel of the CALLER's frames: 1, 2, or 3 (interleaved BGR, channelnum >= 3)
  bool bgr = false;      // BscanFFTwebcam.cpp:1021-1037: sum of the three channels * 0.00130718954
  float px_scale = 1.f;  // factor between the integer pixel (sum) and data_y
  std::vector<int> radW, radM;
  const PlanEntry* plan = nullptr;
  int G = 1, smem = 0, regs = 0;
  std::vector<GpuState> gpus;
  std::string err;
  std::atomic<uint64_t> launches{0};  // bumped by the per-GPU feeder threads of a multi-GPU context
  double last_recon_ms = 0, last_norm_ms = 0;
};

namespace {

int fail(abcoct_ctx* c, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  static std::mutex err_mutex;  // the feeder threads of a multi-GPU context may fail at the same time
  std::lock_guard<std::mutex> lock(err_mutex);
  if (c)
    c->err = buf;
  else
    g_create_error = buf;
  return code;
}

#define CU(c, call)                                                                                      \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess) return fail(c, ABCOCT_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
  } while (0)

// ------------------------------------------------------------------------------------------ reference tables
// BscanFFT.cpp:615-698, same expression order, plain IEEE doubles (this file is built with -ffp-contract=off).
void build_ref_tables(int opw, int m, int N, double lmin, double lmax, std::vector<int32_t>& nk, std::vector<double>& frac) {
  const double pi = 3.141592653589793;  // BscanFFT.cpp:609
  const int M = m * opw;
  const double deltalambda = (lmax - lmin) / opw;  // :615
  std::vector<double> lambdas(M), k(M), klinear(N), diffk(M);
  for (unsigned i = 0; i < (unsigned)M; ++i) lambdas[i] = lmin + i * deltalambda / (unsigned)m;  // :641
  for (int i = 0; i < M; ++i) k[i] = (2 * pi) / lambdas[i];                                       // :644
  const double kmin = 2 * pi / (lmax - deltalambda);                                              // :645
  const double kmax = 2 * pi / lmin;                                                              // :646
  const double deltak = (kmax - kmin) / N;                                                        // :647
  for (unsigned f = 0; f < (unsigned)N; ++f) klinear[f] = kmin + (f + 1) * deltak;                // :652
  for (int i = 1; i < M; ++i) diffk[i] = k[i - 1] - k[i];                                         // :667
  diffk[0] = diffk[1];                                                                            // :671
  nk.assign(N, 0);
  frac.assign(N, 0.0);
  for (int f = 0; f < N; ++f)  // :673-690
    for (int i = 0; i < M; ++i)
      if (k[i] < klinear[f]) {
        nk[f] = i;
        break;
      }
  for (int f = 0; f < N; ++f) frac[f] = (klinear[f] - k[nk[f]]) / diffk[nk[f]];  // :695
}

// BscanFFT.cpp:936-944: x = float(p)/float(opw-1) in f32, everything else in f64.
void build_window(int opw, std::vector<double>& win) {
  const double pi = 3.141592653589793;
  win.resize(opw);
  for (unsigned p = 0; p < (unsigned)opw; ++p) {
    float nn = p;
    float NN = opw - 1;
    win[p] = 0.62 - 0.48 * std::abs(nn / NN - 0.5) + 0.38 * std::cos(2 * pi * (nn / NN - 0.5));
  }
}

// Configurations the fused kernels are not compiled for - a transform length without a plan, rows that are not a multiple of
// 8 samples, display rows above N / 2 - run on the generic kernel (prep_kernels.cu: generic_recon_kernel); cv::dft and colRange
// take all of them (BscanFFT.cpp:1185, 1193).
// Also there: the normalised-calibration regime (rowwisenormalize / !donotnormalize, BscanFFT.cpp:1126-1129, 1044-1047).  The
// captured data_yb is stretched to [1e-4, 1], 1 / data_yb spans four decades and neighbouring rows differ by orders of magnitude;
// the fused kernels pack two rows into one complex transform, whose f32 rounding noise then follows the LARGER row (measured
// 1.9e-4 of the smaller one against the reference's 0.7e-4).  The generic kernel transforms one row at a time there, so every
// A-scan keeps its own noise floor like cv::dft's rows do and the 1e-4 bound holds.
bool single_row_regime(const abcoct_params& p) { return p.rowwisenormalize || !p.donotnormalize; }
bool needs_generic(const abcoct_params& p) {
  return !find_plan((int)p.numfftpoints) || (p.w / p.binx) % 8 != 0 || p.numdisplaypoints > p.numfftpoints / 2 || single_row_regime(p);
}

int validate(const abcoct_params& p, std::string& why, int& code) {
  code = ABCOCT_ERR_INVALID;
  auto bad = [&](const char* s) {
    why = s;
    return 1;
  };
  if (p.w == 0 || p.h == 0) return bad("w and h must be positive");
  if (p.binx == 0 || p.biny == 0) return bad("binning factors must be >= 1");
  if (p.bpp != 8 && p.bpp != 16) return bad("bpp must be 8 or 16");
  if (p.averages == 0) return bad("averages must be >= 1");
  if (p.fft_multiplier == 0) return bad("fft_multiplier must be >= 1");
  if (!(p.lambdamax > p.lambdamin) || !(p.lambdamin > 0)) return bad("need 0 < lambdamin < lambdamax");
  const unsigned opw = p.w / p.binx, oph = p.h / p.biny;
  if (opw < 8 || oph < 1) return bad("binned frame too small");
  if (p.numfftpoints < p.fft_multiplier * opw)
    return bad("numfftpoints < fft_multiplier * (w / binx): the reference reads past fractionalk (BscanFFT.cpp:1170)");
  if (p.numdisplaypoints > p.numfftpoints) return bad("numdisplaypoints > numfftpoints (colRange(0, numdisplaypoints) throws, BscanFFT.cpp:1193)");
  if (p.numdisplaypoints < 6) return bad("numdisplaypoints < 6 (rows 4 and 5 are addressed, BscanFFT.cpp:1239, 1252)");
  if (p.clampupper && oph < 6) return bad("clampupper needs at least 6 A-scans (element (5,5), BscanFFT.cpp:1252)");
  if (p.variant > 1 || p.weight_mode > 1) return bad("variant / weight_mode out of range");
  if (p.binx > 1 && p.w % p.binx) return bad("w is not a multiple of the x binning factor (cv::resize would round the size)");
  if (p.biny > 1 && p.h % p.biny) return bad("h is not a multiple of the y binning factor (cv::resize would round the size)");
  if (p.mediann < 0 || p.movavgn < 0) return bad("mediann / movavgn must be >= 0");
  if (p.channelnum >= 3 && p.bpp != 8) return bad("channelnum >= 3 (sum of the BGR channels) needs 8-bit frames");
  if (p.channelnum >= 3 && p.mediann != 0)
    return bad("channelnum >= 3 with mediann > 0: cv::medianBlur throws on the CV_64F channel sum in the reference (BscanFFTwebcam.cpp:1045)");
  code = ABCOCT_ERR_UNSUPPORTED;
  if (p.mediann != 0 && p.mediann != 3 && p.mediann != 5)
    return bad("medianBlur kernel sizes other than 3 and 5 are not built (OpenCV itself only takes 3 / 5 for 16-bit frames)");
  if (p.movavgn > 64) return bad("movavgn > 64 is not built");
  if (p.channelnum >= 3 && (p.binx > 1 || p.biny > 1))
    return bad("channelnum >= 3 with binning (INTER_AREA on the CV_64F channel sum, BscanFFTwebcam.cpp:1049) is not built");
  if (p.output_rebin && (p.bscanbinx > 1 || p.bscanbiny > 1 || p.biny > 1))
    return bad("BscanFFTspinjnt's output re-binning of the linear B-scan with a real resampling step (INTER_AREA down by bscanbinx/y, "
               "INTER_CUBIC up by bscanbinx * binvaluey and bscanbiny; BscanFFTspinjnt.cpp:1856-1862) is not built: the bicubic "
               "overshoot goes negative next to every bright A-scan and the reference's own log() turns that into NaN; only the "
               "shipped shape (bscanbinx = bscanbiny = binvaluey = 1, where both resizes are copies and the block multiplies the "
               "B-scan by multiplyfactor) is supported");
  if (p.fft_multiplier > 1) {
    unsigned r = opw;
    for (unsigned f : {2u, 3u, 5u})
      while (r % f == 0) r /= f;
    if (r != 1 || (opw & 1)) return bad("increasefftpointsmultiplier > 1 needs w / binx = 2^a 3^b 5^c, even");
    r = p.fft_multiplier;
    for (unsigned f : {2u, 3u, 5u})
      while (r % f == 0) r /= f;
    if (r != 1) return bad("increasefftpointsmultiplier must be of the form 2^a 3^b 5^c");
  }
  if (needs_generic(p)) {  // any N = 2^a 3^b 5^c, any row width, D up to N: the shared-memory Stockham kernel (prep_kernels.cu)
    unsigned r = p.numfftpoints;
    for (unsigned f : {2u, 3u, 5u})
      while (r % f == 0) r /= f;
    if (r != 1) {
      why = "numfftpoints must be of the form 2^a 3^b 5^c (cv::getOptimalDFTSize lengths); fused plans exist for:";
      int ns[64];
      int n = list_plans(ns, 64);
      for (int i = 0; i < n && i < 64; ++i) why += " " + std::to_string(ns[i]);
      return 1;
    }
    if (generic_smem_bytes((int)p.numfftpoints, (int)p.numdisplaypoints) > 227 * 1024)
      return bad("numfftpoints too long for the shared-memory transform (two complex rows and two accumulators must fit in 227 KB)");
  }
  code = 0;
  return 0;
}

std::vector<int> factor_radices(int n) {  // greedy, largest first; every entry has a register butterfly in prep_kernels.cu
  std::vector<int> r;
  for (int f : {16, 15, 12, 10, 9, 8, 6, 5, 4, 3, 2})
    while (n % f == 0) {
      r.push_back(f);
      n /= f;
    }
  return r;
}

int upload_calibration(abcoct_ctx* c) {
  if (!c->have_yb) return fail(c, ABCOCT_ERR_STATE, "no background set: data_yb is all zeros in the reference until key 'b' (BscanFFT.cpp:562)");
  const size_t n = (size_t)c->oph * c->opw;
  std::vector<float> gain(n), subg(n);
  bool has_sub = false;
  const bool dark = c->p.variant == 1 && c->have_yd;
  for (size_t i = 0; i < n; ++i) {
    const double sub = (dark ? c->yd[i] : 0.0) + (c->have_yp ? c->yp[i] : 0.0);
    gain[i] = (float)(1.0 / c->yb[i]);
    subg[i] = (float)(sub / c->yb[i] + 1.0);  // + 1: the kernel stages t - 1 (see phase_pre in recon_kernel.cuh)
    has_sub |= (sub != 0.0);
  }
  c->has_sub = has_sub && !c->general;
  {  // unswizzled f32 copies for rowprep_kernel (general path and debug tap)
    std::vector<float> fb(n), fp(n), fd(n);
    for (size_t i = 0; i < n; ++i) {
      fb[i] = (float)c->yb[i];
      fp[i] = c->have_yp ? (float)c->yp[i] : 0.f;
      fd[i] = dark ? (float)c->yd[i] : 0.f;
    }
    for (GpuState& g : c->gpus) {
      CU(c, cudaSetDevice(g.dev));
      if (!g.d_yb) CU(c, cudaMalloc(&g.d_yb, n * 4));
      if (!g.d_yp) CU(c, cudaMalloc(&g.d_yp, n * 4));
      if (!g.d_yd) CU(c, cudaMalloc(&g.d_yd, n * 4));
      CU(c, cudaMemcpy(g.d_yb, fb.data(), n * 4, cudaMemcpyHostToDevice));
      CU(c, cudaMemcpy(g.d_yp, fp.data(), n * 4, cudaMemcpyHostToDevice));
      CU(c, cudaMemcpy(g.d_yd, fd.data(), n * 4, cudaMemcpyHostToDevice));
      if (c->single) {  // the f64 row preparation reads the calibration frames as they are
        std::vector<double> zero(n, 0.0);
        if (!g.d_yb64) CU(c, cudaMalloc(&g.d_yb64, n * 8));
        if (!g.d_yp64) CU(c, cudaMalloc(&g.d_yp64, n * 8));
        if (!g.d_yd64) CU(c, cudaMalloc(&g.d_yd64, n * 8));
        CU(c, cudaMemcpy(g.d_yb64, c->yb.data(), n * 8, cudaMemcpyHostToDevice));
        CU(c, cudaMemcpy(g.d_yp64, c->have_yp ? c->yp.data() : zero.data(), n * 8, cudaMemcpyHostToDevice));
        CU(c, cudaMemcpy(g.d_yd64, dark ? c->yd.data() : zero.data(), n * 8, cudaMemcpyHostToDevice));
      }
    }
  }
  if (c->generic) {  // generic_recon_kernel reads the prepared rows only: no swizzled calibration, no table blob, no plan attributes
    c->smem = (int)generic_smem_bytes(c->N, c->D);
    c->G = 1;
    c->regs = 0;
    c->cal_dirty = false;
    return ABCOCT_OK;
  }
  // which kernel: the warp-per-A-scan kernel wherever it applies (wrow_kernel.cuh).  For A/B measurements ABCOCT_KERNEL=1 forces
  // the group-per-row-pair kernel (recon_kernel.cuh, the fallback for every other transform length and for the general path) and
  // ABCOCT_KERNEL=3 the resident-row kernel (wres_kernel.cuh: dB rows parked in tensor memory, no global scratch - 1.3x instead
  // of 2.9x the algorithmic DRAM traffic, but measured slower: its 4-byte stores into the depth-major image are partial-sector
  // writes); ABCOCT_WROW_NW / ABCOCT_WRES_NW pick another occupancy point of a plan.
  const WPlanEntry* wp = nullptr;
  if (c->wrow_eligible) {
    int force = 0, nw = 0, lm = -1, rnw = 0;
    if (const char* e = getenv("ABCOCT_KERNEL")) force = atoi(e);
    if (const char* e = getenv("ABCOCT_WROW_NW")) nw = atoi(e);
    if (const char* e = getenv("ABCOCT_WROW_LM")) lm = atoi(e);
    if (const char* e = getenv("ABCOCT_WRES_NW")) rnw = atoi(e);
    if (force == 3 && !c->rblob.empty()) {  // opt-in: measured slower than the scratch kernel
      wp = find_rplan(c->N, rnw);
      if (!wp) wp = find_rplan(c->N, 0);
      // progress of the slot protocol: a team's rounds k - K and k must never lie in the same B-scan (wres_kernel.cuh)
      const long long nbb = (c->oph + 3) / 4;
      if (wp && nbb > (long long)wp->slots * wp->teams * c->gpus[0].sm_count) wp = nullptr;
    }
    if (!wp && force != 1) {  // 0 / 2
      wp = find_wplan(c->N, nw, lm);
      if (!wp) wp = find_wplan(c->N, 0, -1);
    }
  }
  c->wplan = wp;
  const size_t pitch = wp ? (size_t)wp->wmax : (size_t)c->opw;
  {  // calibration rows in the layout the chosen kernel reads (wrow_permute_cal_row / cal_swizzle_row)
    std::vector<float> pg((size_t)c->oph * pitch, 0.f), ps((size_t)c->oph * pitch, 0.f);
    for (int r = 0; r < c->oph; ++r) {
      if (wp) {
        wp->permute_cal_row(&gain[(size_t)r * c->opw], c->opw, &pg[(size_t)r * pitch]);
        wp->permute_cal_row(&subg[(size_t)r * c->opw], c->opw, &ps[(size_t)r * pitch]);
      } else {
        cal_swizzle_row(&gain[(size_t)r * c->opw], &pg[(size_t)r * pitch], c->opw);
        cal_swizzle_row(&subg[(size_t)r * c->opw], &ps[(size_t)r * pitch], c->opw);
      }
    }
    for (GpuState& g : c->gpus) {
      CU(c, cudaSetDevice(g.dev));
      CU(c, cudaDeviceSynchronize());  // no launch may still be reading the old calibration (async device entry point)
      if (g.cal_floats < pg.size()) {
        cudaFree(g.d_gain);
        cudaFree(g.d_subg);
        g.d_gain = g.d_subg = nullptr;
        g.cal_floats = 0;
        CU(c, cudaMalloc(&g.d_gain, pg.size() * 4));
        CU(c, cudaMalloc(&g.d_subg, pg.size() * 4));
        g.cal_floats = pg.size();
      }
      CU(c, cudaMemcpy(g.d_gain, pg.data(), pg.size() * 4, cudaMemcpyHostToDevice));
      CU(c, cudaMemcpy(g.d_subg, ps.data(), ps.size() * 4, cudaMemcpyHostToDevice));
    }
  }
  const std::vector<unsigned char>& blob = wp ? (wp->resident ? c->rblob : c->wblob) : c->blob1;
  if (c->blob_loaded != &blob || c->gpus[0].d_tables == nullptr) {
    for (GpuState& g : c->gpus) {
      CU(c, cudaSetDevice(g.dev));
      cudaFree(g.d_tables);
      g.d_tables = nullptr;
      CU(c, cudaMalloc(&g.d_tables, blob.size()));
      CU(c, cudaMemcpy(g.d_tables, blob.data(), blob.size(), cudaMemcpyHostToDevice));
    }
    c->blob_loaded = &blob;
  }
  if (wp) {
    c->G = wp->nw;
    c->smem = wp->smem_bytes;
    for (GpuState& g : c->gpus) {
      CU(c, cudaSetDevice(g.dev));
      CU(c, wp->attrs(c->has_sub, c->A == 1, c->D == c->N / 2, &c->regs));
    }
  } else {
    const int G = c->plan->groups(c->has_sub);  // compile-time choice of the plan
    c->G = G;
    c->smem = c->plan->smem_bytes(c->general ? c->M : c->opw, c->has_sub, G);
    if (c->smem > 227 * 1024) return fail(c, ABCOCT_ERR_UNSUPPORTED, "shared memory budget exceeded (%d bytes)", c->smem);
    for (GpuState& g : c->gpus) {
      CU(c, cudaSetDevice(g.dev));
      CU(c, c->plan->attrs(c->has_sub, c->A == 1, c->general, c->smem, &c->regs));
    }
  }
  c->cal_dirty = false;
  return ABCOCT_OK;
}

int scratch_pitch(const abcoct_ctx* c) { return (c->D + 31) / 32 * 32; }  // whole 128-byte lines per A-scan

int ensure_scratch(abcoct_ctx* c, GpuState& g, int slot, size_t nB) {
  if (g.scratch_bscans[slot] >= nB) return ABCOCT_OK;
  if (g.d_scratch[slot]) cudaFree(g.d_scratch[slot]);
  if (g.d_sched[slot]) cudaFree(g.d_sched[slot]);
  if (g.d_jmm[slot]) cudaFree(g.d_jmm[slot]);
  if (g.d_dc01[slot]) cudaFree(g.d_dc01[slot]);
  g.d_dc01[slot] = nullptr;
  g.d_scratch[slot] = nullptr;
  g.d_sched[slot] = nullptr;
  g.d_jmm[slot] = nullptr;
  g.scratch_bscans[slot] = 0;
  const bool resident = c->wplan && c->wplan->resident;  // the resident-row kernel keeps the dB rows in shared memory: no scratch
  CU(c, cudaMalloc(&g.d_scratch[slot], resident ? 256 : nB * c->oph * (size_t)scratch_pitch(c) * sizeof(float)));
  CU(c, cudaMalloc(&g.d_sched[slot], sched_ints((int)nB) * sizeof(int)));
  CU(c, cudaMalloc(&g.d_jmm[slot], 2 * nB * sizeof(int)));
  CU(c, cudaMalloc(&g.d_dc01[slot], 2 * nB * c->oph * sizeof(float)));
  g.scratch_bscans[slot] = nB;
  return ABCOCT_OK;
}

size_t scratch_chunk_bscans(const abcoct_ctx* c) {
  size_t mb = 8192;  // address space only: the scratch lives in L2 between the two halves of the kernel
  if (const char* e = getenv("ABCOCT_SCRATCH_MB")) mb = (size_t)std::max(1L, atol(e));
  const size_t per = (size_t)c->oph * scratch_pitch(c) * sizeof(float);
  return std::max<size_t>(1, mb * 1024 * 1024 / per);
}

int ensure_prep(abcoct_ctx* c, GpuState& g, int slot, size_t nframes) {
  if (g.prep_frames[slot] >= nframes) return ABCOCT_OK;
  cudaFree(g.d_med[slot]);
  cudaFree(g.d_bin[slot]);
  cudaFree(g.d_rows[slot]);
  cudaFree(g.d_fmm[slot]);
  cudaFree(g.d_rows64[slot]);
  cudaFree(g.d_pre32[slot]);
  cudaFree(g.d_fmm64[slot]);
  g.d_rows64[slot] = nullptr;
  g.d_pre32[slot] = nullptr;
  g.d_fmm64[slot] = nullptr;
  g.d_med[slot] = g.d_bin[slot] = nullptr;
  g.d_rows[slot] = g.d_fmm[slot] = nullptr;
  g.prep_frames[slot] = 0;
  if (c->bgr) {  // the channel sums (<= 765) as 16-bit frames; reuses the median buffer slot (mediann is 0 with channelnum >= 3)
    CU(c, cudaMalloc(&g.d_med[slot], nframes * c->p.h * c->p.w * 2));
  }
  if (c->p.mediann > 0) CU(c, cudaMalloc(&g.d_med[slot], nframes * c->p.h * c->p.w * c->px_bytes));
  if (c->p.binx > 1 || c->p.biny > 1) CU(c, cudaMalloc(&g.d_bin[slot], nframes * c->oph * c->opw * c->px_bytes));
  CU(c, cudaMalloc(&g.d_rows[slot], nframes * c->oph * (size_t)c->M * sizeof(float)));
  CU(c, cudaMalloc(&g.d_fmm[slot], nframes * 2 * sizeof(float)));
  if (c->single) {
    if (c->p.fft_multiplier > 1)
      CU(c, cudaMalloc(&g.d_pre32[slot], nframes * c->oph * (size_t)c->opw * sizeof(float)));
    else
      CU(c, cudaMalloc(&g.d_rows64[slot], nframes * c->oph * (size_t)c->opw * sizeof(double)));
    CU(c, cudaMalloc(&g.d_fmm64[slot], nframes * 2 * sizeof(long long)));
  }
  g.prep_frames[slot] = nframes;
  return ABCOCT_OK;
}

// General path: median -> binning -> row preparation (+ Fourier upsample) for `nframes` frames into g.d_rows[slot].
int run_prep(abcoct_ctx* c, GpuState& g, int slot, const uint8_t* d_frames, size_t nframes, size_t row_stride, size_t frame_stride,
             cudaStream_t st, int* launches) {
  const int pb = c->px_bytes;
  if (row_stride % pb || frame_stride % pb) return fail(c, ABCOCT_ERR_INVALID, "strides must be multiples of the pixel size");
  const void* src = d_frames;
  size_t rs = row_stride / pb, fs = frame_stride / pb;
  int n = 0;
  int bpp = (int)c->p.bpp;
  if (c->bgr) {  // BscanFFTwebcam.cpp:1021-1037
    CU(c, launch_bgr_sum(d_frames, static_cast<uint16_t*>(g.d_med[slot]), (int)c->p.w, (int)c->p.h, row_stride, frame_stride, (int)nframes, st));
    src = g.d_med[slot];
    rs = c->p.w;
    fs = (size_t)c->p.w * c->p.h;
    bpp = 16;
    ++n;
  }
  if (c->p.mediann > 0) {
    CU(c, launch_median(src, g.d_med[slot], (int)c->p.bpp, c->p.mediann, (int)c->p.w, (int)c->p.h, rs, fs, (int)nframes, st));
    src = g.d_med[slot];
    rs = c->p.w;
    fs = (size_t)c->p.w * c->p.h;
    ++n;
  }
  if (c->p.binx > 1 || c->p.biny > 1) {
    CU(c, launch_bin(src, g.d_bin[slot], (int)c->p.bpp, c->opw, c->oph, (int)c->p.binx, (int)c->p.biny, rs, fs, (int)nframes, st));
    src = g.d_bin[slot];
    rs = c->opw;
    fs = (size_t)c->opw * c->oph;
    ++n;
  }
  int n64 = 0;
  if (c->single) {  // normalised-calibration regime: every stage up to the apodised row in f64 (rowprep64_kernel)
    PrepArgs64Host q{};
    q.binned = src;
    q.row_stride = rs;
    q.frame_stride = fs;
    q.bpp = bpp;
    q.opw = c->opw;
    q.oph = c->oph;
    q.nframes = (int)nframes;
    q.movavgn = c->p.movavgn;
    q.px_scale = c->bgr ? 0.00130718954 : 1.0;
    q.yd = (c->p.variant == 1 && c->have_yd) ? g.d_yd64 : nullptr;
    q.yb = g.d_yb64;
    q.yp = c->have_yp ? g.d_yp64 : nullptr;
    q.win = g.d_win64;
    q.rowwise = c->p.rowwisenormalize;
    q.global_norm = c->p.donotnormalize ? 0 : 1;
    q.frame_minmax = g.d_fmm64[slot];
    q.out64 = g.d_rows64[slot];
    q.out32 = g.d_pre32[slot];
    CU(c, launch_rowprep64(q, st, &n64));
    if (c->p.fft_multiplier <= 1) {
      *launches = n + n64;
      return ABCOCT_OK;
    }
  }
  PrepArgsHost h{};
  h.pre = c->single ? g.d_pre32[slot] : nullptr;
  h.binned = src;  // without binning the row kernel reads the caller's (or the median's) frames through their strides
  h.row_stride = rs;
  h.frame_stride = fs;
  h.bpp = bpp;
  h.px_scale = c->px_scale;
  h.opw = c->opw;
  h.oph = c->oph;
  h.nframes = (int)nframes;
  h.movavgn = c->p.movavgn;
  h.yd = (c->p.variant == 1 && c->have_yd) ? g.d_yd : nullptr;
  h.rowwise = c->p.rowwisenormalize;
  h.global_norm = c->p.donotnormalize ? 0 : 1;
  h.frame_minmax = g.d_fmm[slot];
  h.yb = g.d_yb;
  h.yp = c->have_yp ? g.d_yp : nullptr;
  h.win = g.d_win;
  h.m = (int)c->p.fft_multiplier;
  h.M = c->M;
  h.bandpass = c->p.bandpassfilter;
  h.nradW = (int)c->radW.size();
  h.nradM = (int)c->radM.size();
  for (size_t i = 0; i < c->radW.size(); ++i) h.radW[i] = c->radW[i];
  for (size_t i = 0; i < c->radM.size(); ++i) h.radM[i] = c->radM[i];
  h.twW = g.d_twW;
  h.twM = g.d_twM;
  h.out = g.d_rows[slot];
  int nl = 0;
  CU(c, launch_rowprep(h, st, &nl));
  *launches = n + nl + n64;
  return ABCOCT_OK;
}

int ensure_tmp(abcoct_ctx* c, GpuState& g, int which, int slot, size_t bytes) {
  if (g.tmp_bytes[which][slot] >= bytes) return ABCOCT_OK;
  if (g.d_tmp[which][slot]) cudaFree(g.d_tmp[which][slot]);
  g.d_tmp[which][slot] = nullptr;
  g.tmp_bytes[which][slot] = 0;
  CU(c, cudaMalloc(&g.d_tmp[which][slot], bytes));
  g.tmp_bytes[which][slot] = bytes;
  return ABCOCT_OK;
}

int upload_jscan(abcoct_ctx* c) {
  for (GpuState& g : c->gpus) {
    CU(c, cudaSetDevice(g.dev));
    CU(c, cudaDeviceSynchronize());  // no launch may still be reading the old reference
    if (!g.d_jscan) CU(c, cudaMalloc(&g.d_jscan, (size_t)c->D * c->oph * sizeof(float)));
    CU(c, cudaMemcpy(g.d_jscan, c->jscan.data(), c->jscan.size() * sizeof(float), cudaMemcpyHostToDevice));
  }
  c->jscan_dirty = false;
  return ABCOCT_OK;
}

// Enqueue the kernels for nB B-scans resident on device g; everything on `st`.  `o` holds DEVICE pointers.
int enqueue_device(abcoct_ctx* c, GpuState& g, int slot, const uint8_t* d_frames, size_t nB, size_t row_stride, size_t frame_stride,
                   const OutPtrs& o_in, cudaStream_t st, bool time_it) {
  OutPtrs o = o_in;
  const size_t px = (size_t)c->D * c->oph;
  const bool want_j = o.p[O_JSUB] || o.p[O_JBGR];
  if (want_j && c->db_shift != 0.f)
    return fail(c, ABCOCT_ERR_UNSUPPORTED, "the J0 lock-in outputs together with BscanFFTspinjnt's multiplyfactor are not built");
  if (want_j) {
    if (c->jscan.empty()) return fail(c, ABCOCT_ERR_STATE, "jsub output requested but no jscan is set (abcoct_set_jscan)");
    if (c->jscan_dirty) {
      int rc = upload_jscan(c);
      if (rc) return rc;
      CU(c, cudaSetDevice(g.dev));
    }
    if (!o.p[O_JSUB]) {
      int rc = ensure_tmp(c, g, 0, slot, nB * px);
      if (rc) return rc;
      o.p[O_JSUB] = g.d_tmp[0][slot];
    }
  }
  // the linear image - stored on request, or evaluated on the fly by the lock-in kernels - is derived from the dB image and
  // the unmasked DC rows
  const bool need_db = o.p[O_LIN] || want_j;
  if (need_db && !o.p[O_DB]) {
    int rc = ensure_tmp(c, g, 1, slot, nB * px * sizeof(float));
    if (rc) return rc;
    o.p[O_DB] = g.d_tmp[1][slot];
  }
  uint8_t* d_out8 = static_cast<uint8_t*>(o.p[O_U8]);
  float* d_outdb = static_cast<float*>(o.p[O_DB]);
  float* d_outlin = static_cast<float*>(o.p[O_LIN]);
  size_t chunkB = std::min(nB, scratch_chunk_bscans(c));
  if (c->general) {  // the pre-processed f32 rows of a chunk stay below ~1 GiB
    const size_t per_bscan = (size_t)c->A * c->oph * c->M * (c->single ? sizeof(double) : sizeof(float));
    chunkB = std::min(chunkB, std::max<size_t>(1, ((size_t)1 << 30) / per_bscan));
  }
  if (c->generic) chunkB = std::min<size_t>(chunkB, 65535);  // one grid row / plane per B-scan
  int rc = ensure_scratch(c, g, slot, chunkB);
  if (rc) return rc;
  if (c->general) {
    rc = ensure_prep(c, g, slot, chunkB * c->A);
    if (rc) return rc;
  }
  for (size_t b0 = 0; b0 < nB; b0 += chunkB) {
    const size_t nb = std::min(chunkB, nB - b0);
    ReconArgs a{};
    a.frames = d_frames + b0 * c->A * frame_stride;
    a.frame_stride = frame_stride;
    a.row_stride = row_stride;
    a.W = c->opw;
    if (c->general) {
      int nl = 0;
      rc = run_prep(c, g, slot, a.frames, nb * c->A, row_stride, frame_stride, st, &nl);
      if (rc) return rc;
      c->launches += nl;
      a.frames = reinterpret_cast<const uint8_t*>(g.d_rows[slot]);
      a.W = c->M;
    }
    a.oph = c->oph;
    a.D = c->D;
    a.Dp = scratch_pitch(c);
    a.A = c->A;
    a.nB = (int)nb;
    a.npairs = (c->oph + 1) / 2;
    if ((size_t)a.npairs * nb > 0x7fff0000u) return fail(c, ABCOCT_ERR_INVALID, "too many A-scan pairs in one chunk");
    a.nitems = a.npairs * (int)nb;
    if (c->generic) {  // any transform length / row width / D up to N: two launches on the prepared rows
      GenericHost gh{};
      gh.rows = g.d_rows[slot];
      gh.rows64 = (c->single && c->p.fft_multiplier <= 1) ? g.d_rows64[slot] : nullptr;
      gh.wq64 = g.d_gwq64;
      gh.M = c->M; gh.N = c->N; gh.D = c->D; gh.Dp = a.Dp; gh.oph = c->oph; gh.A = c->A; gh.nB = (int)nb;
      gh.idx = g.d_gidx;
      gh.wq = g.d_gwq;
      gh.nrad = (int)c->radN.size();
      for (size_t i = 0; i < c->radN.size(); ++i) gh.rad[i] = c->radN[i];
      gh.tw = g.d_twN;
      gh.scratch = g.d_scratch[slot];
      gh.minv = g.d_jmm[slot];  // free until the lock-in kernels run (they reset it themselves)
      gh.maxv = g.d_jmm[slot] + nb;
      gh.dc01 = need_db ? g.d_dc01[slot] : nullptr;
      gh.out8 = d_out8 + b0 * c->D * c->oph;
      gh.outdb = d_outdb ? d_outdb + b0 * c->D * c->oph : nullptr;
      gh.out_scale = 0.5f / (float)c->A;
      gh.db_scale_ln = (float)(20.0 * (1.0 / 2.303));
      gh.thr = (float)c->p.bscanthreshold - c->db_shift;
      gh.clamp_db = (float)c->p.clamp_db - c->db_shift;
      gh.clamp55 = c->p.clampupper ? 1 : 0;
      gh.single_row = single_row_regime(c->p) ? 1 : 0;
      a.out8 = gh.out8;
      a.outdb = gh.outdb;
      a.dc01 = gh.dc01;
      a.db_scale = (float)(0.6931471805599453 * (20.0 * (1.0 / 2.303)));
      a.thr = gh.thr;
      const bool timed = time_it && g.tev_used + 3 <= g.tev.size();
      if (timed) CU(c, cudaEventRecord(g.tev[g.tev_used], st));
      int nl = 0;
      CU(c, launch_generic(gh, st, &nl));
      if (timed) {
        CU(c, cudaEventRecord(g.tev[g.tev_used + 1], st));
        CU(c, cudaEventRecord(g.tev[g.tev_used + 2], st));
        g.tev_used += 3;
      }
      c->launches += nl;
    } else {
    a.nparts = (c->oph + c->plan->d.T - 1) / c->plan->d.T;
    int grid = 0;
    if (c->wplan) {  // one item per row, normalisation parts of 32 A-scans split into depth-tile ranges for short launches
      if ((size_t)c->oph * nb > 0x7fff0000u) return fail(c, ABCOCT_ERR_INVALID, "too many A-scans in one chunk");
      a.nitems = c->oph * (int)nb;
      a.nparts = (c->oph + 31) / 32;
      a.calpitch = c->wplan->wmax;
      const int workers = c->wplan->nw - 1;  // the last warp of a CTA is its service warp
      grid = std::min(g.sm_count, (a.nitems + workers - 1) / workers);
      if (c->wplan->resident) {  // static schedule: blocks of 4 A-scans over teams of 4 warps
        const long long blocks = (long long)((c->oph + 3) / 4) * (long long)nb;
        grid = (int)std::min<long long>(g.sm_count, (blocks + c->wplan->teams - 1) / c->wplan->teams);
      }
      const int ntiles = (c->D + 31) / 32;
      const long long parts = (long long)a.nparts * (long long)nb;
      // normalisation jobs (32 A-scans x a range of depth tiles) are handed out dynamically as soon as a B-scan is complete:
      // about eight tiles (32 KB of dB scratch) per job keeps them short, so the scratch is consumed while it is still in
      // L2; short launches get at least two jobs per CTA
      long long split = std::max<long long>((ntiles + 7) / 8, (2LL * grid + parts - 1) / parts);
      if (const char* e = getenv("ABCOCT_NSPLIT")) split = atol(e);
      a.nsplit = (int)std::max<long long>(1, std::min<long long>(ntiles, split));
      a.hints = 0;
      if (const char* e = getenv("ABCOCT_HINTS")) a.hints = atoi(e);
      // EXPERIMENT, only in a build with -DABC_WROW_RING (ABCOCT_BUILD_RING=1): the dB scratch as a ring reused round-robin
      a.ringB = 0;
#ifdef ABC_WROW_RING
      if (!c->wplan->resident) {
        size_t ring_mb = 0;
        if (const char* e = getenv("ABCOCT_RING_MB")) ring_mb = (size_t)std::max(0L, atol(e));
        const size_t per = (size_t)c->oph * a.Dp * sizeof(float);
        const size_t ring = std::max<size_t>(8, ring_mb * 1024 * 1024 / per);
        if (ring_mb > 0 && ring < nb) a.ringB = (int)ring;
      }
#endif
    }
    a.gain = g.d_gain;
    a.subg = g.d_subg;
    a.idxT = reinterpret_cast<const uint32_t*>(g.d_tables);
    a.scratch = g.d_scratch[slot];
    a.sched = g.d_sched[slot];
    a.out8 = d_out8 + b0 * c->D * c->oph;
    a.outdb = d_outdb ? d_outdb + b0 * c->D * c->oph : nullptr;
    a.dc01 = need_db ? g.d_dc01[slot] : nullptr;
    a.inv_W = 1.0f / (float)c->opw;
    a.out_scale = 0.5f / (float)c->A;
    a.db_scale = (float)(0.6931471805599453 * (20.0 * (1.0 / 2.303)));
    a.thr = (float)c->p.bscanthreshold - c->db_shift;
    a.clamp_db = (float)c->p.clamp_db - c->db_shift;
    a.clamp55 = c->p.clampupper ? 1 : 0;
    if (!c->wplan) grid = std::min(g.sm_count, (a.nitems + c->G - 1) / c->G);
    CU(c, launch_sched_init(a.sched, (int)nb, st));
    const bool timed = time_it && g.tev_used + 3 <= g.tev.size();
    if (timed) CU(c, cudaEventRecord(g.tev[g.tev_used], st));
    if (c->wplan)
      CU(c, c->wplan->launch(a, c->has_sub, grid, st));
    else
      CU(c, c->plan->launch(a, c->has_sub, c->general, grid, st));
    if (timed) {
      CU(c, cudaEventRecord(g.tev[g.tev_used + 1], st));
      CU(c, cudaEventRecord(g.tev[g.tev_used + 2], st));
      g.tev_used += 3;
    }
    c->launches += 2;
    }
    if (c->db_shift != 0.f && (a.outdb || a.dc01)) {  // the dB image (and the unmasked DC rows) of bscan * multiplyfactor
      int nl = 0;
      CU(c, launch_add_const(a.outdb, nb * px, a.dc01, 2 * nb * (size_t)c->oph, c->db_shift, g.sm_count, st, &nl));
      c->launches += nl;
    }
    // consumers of the finished B-scans (post_kernels.cu), same stream
    float* lin = d_outlin ? d_outlin + b0 * px : nullptr;
    const float inv = (float)(1.0 / (0.6931471805599453 * (20.0 * (1.0 / 2.303))));
    if (lin) {
      int nl = 0;
      CU(c, launch_lin_from_db(a.outdb, a.dc01, lin, c->oph, px, (int)nb, inv, g.sm_count, st, &nl));
      c->launches += nl;
    }
    if (o.p[O_BGR]) {
      CU(c, launch_jet(a.out8, static_cast<uint8_t*>(o.p[O_BGR]) + 3 * b0 * px, nb * px, g.sm_count, st));
      c->launches += 1;
    }
    if (want_j) {
      int nl = 0;
      uint8_t* jsub = static_cast<uint8_t*>(o.p[O_JSUB]) + b0 * px;
      CU(c, launch_jsub(lin, a.outdb, a.dc01, g.d_jscan, g.d_jmm[slot], jsub, c->oph, px, (int)nb, a.db_scale, inv, a.thr, g.sm_count,
                        st, &nl));
      c->launches += nl;
      if (o.p[O_JBGR]) {
        CU(c, launch_jet(jsub, static_cast<uint8_t*>(o.p[O_JBGR]) + 3 * b0 * px, nb * px, g.sm_count, st));
        c->launches += 1;
      }
    }
  }
  return ABCOCT_OK;
}

int ensure_slots(abcoct_ctx* c, GpuState& g, size_t slotB, unsigned want) {
  const size_t in_bytes = slotB * c->A * (size_t)c->p.h * c->p.w * c->px_bytes;
  const size_t out_px = slotB * (size_t)c->D * c->oph;
  if (g.slot_in_bytes >= in_bytes && g.slot_out_px >= out_px && (g.slot_mask & want) == want) return ABCOCT_OK;
  CU(c, cudaSetDevice(g.dev));
  want |= g.slot_mask;
  g.slot_in_bytes = g.slot_out_px = 0;  // nothing is valid until every buffer below exists
  g.slot_mask = 0;
  for (int s = 0; s < kSlots; ++s) {
    if (g.d_in[s]) cudaFree(g.d_in[s]);
    if (g.h_in[s]) cudaFreeHost(g.h_in[s]);
    g.d_in[s] = g.h_in[s] = nullptr;
    for (int k = 0; k < O_COUNT; ++k) {
      if (g.d_o[k][s]) cudaFree(g.d_o[k][s]);
      if (g.h_o[k][s]) cudaFreeHost(g.h_o[k][s]);
      g.d_o[k][s] = g.h_o[k][s] = nullptr;
    }
    CU(c, cudaMalloc(&g.d_in[s], in_bytes));
    CU(c, cudaMallocHost(&g.h_in[s], in_bytes));
    for (int k = 0; k < O_COUNT; ++k)
      if (want & (1u << k)) {
        CU(c, cudaMalloc(&g.d_o[k][s], out_px * kOutBpp[k]));
        CU(c, cudaMallocHost(&g.h_o[k][s], out_px * kOutBpp[k]));
      }
  }
  g.slot_in_bytes = in_bytes;
  g.slot_out_px = out_px;
  g.slot_mask = want;
  return ABCOCT_OK;
}

// Staging copies between pageable caller memory and the pinned ring.  One core moves ~10 GB/s, a PCIe 5 x16 link takes 50+:
// large copies are split over a few threads (the slot is 64 MiB; thread start-up is ~50 us each).
void staging_copy(void* dst, const void* src, size_t bytes) {
  constexpr size_t kMinPerThread = 4u << 20;
  unsigned hw = std::thread::hardware_concurrency();
  size_t nt = std::min<size_t>({bytes / kMinPerThread, hw ? hw / 2 : 1, 8});
  if (const char* e = getenv("ABCOCT_COPY_THREADS")) nt = std::min<size_t>((size_t)std::max(1L, atol(e)), 64);
  if (nt <= 1) {
    memcpy(dst, src, bytes);
    return;
  }
  const size_t per = ((bytes + nt - 1) / nt + 63) & ~(size_t)63;
  std::vector<std::thread> th;
  th.reserve(nt - 1);
  size_t done_to = std::min(per, bytes);  // [0, done_to) is copied by this thread, the rest by the helpers that could be started
  for (size_t t = 1; t < nt; ++t) {
    const size_t o = t * per;
    if (o >= bytes) break;
    const size_t n = std::min(per, bytes - o);
    try {
      th.emplace_back([=] { memcpy(static_cast<char*>(dst) + o, static_cast<const char*>(src) + o, n); });
    } catch (...) {  // no thread to be had (resource limit): this thread takes the remainder, nothing crosses the C ABI
      memcpy(static_cast<char*>(dst) + o, static_cast<const char*>(src) + o, bytes - o);
      break;
    }
  }
  memcpy(dst, src, done_to);
  for (std::thread& t : th) t.join();
}

bool is_pinned(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeHost;
}

}  // namespace

// ================================================================================================ C ABI
extern "C" {

void abcoct_params_default(abcoct_params* o) {
  if (!o) return;
  memset(o, 0, sizeof *o);
  o->w = 640;  // BscanFFT.cpp:389
  o->h = 480;  // :390
  o->bpp = 8;  // :357
  o->binx = o->biny = 1;
  o->averages = 1;
  o->numfftpoints = 1024;     // :368
  o->numdisplaypoints = 512;  // :369
  o->lambdamin = 816e-9;      // :381
  o->lambdamax = 884e-9;      // :382
  o->mediann = 5;             // :383
  o->movavgn = 0;             // :373
  o->fft_multiplier = 1;      // :384
  o->rowwisenormalize = 0;    // :386
  o->donotnormalize = 1;      // :387
  o->variant = 0;
  o->weight_mode = 0;
  o->bscanthreshold = -30.0;  // :385
  o->clampupper = 0;          // :374
  o->clamp_db = 50.0;         // :1252
  o->bscanbinx = o->bscanbiny = 1;  // BscanFFTspinjnt.cpp:707
}

int abcoct_params_from_ini(const char* path, int flavour, abcoct_params* o) {
  if (!o || !path) return ABCOCT_ERR_INVALID;
  if (flavour < ABCOCT_INI_BSCANFFT || flavour > ABCOCT_INI_SIM) return ABCOCT_ERR_INVALID;
  abcoct_params_default(o);
  if (flavour == ABCOCT_INI_DARK) o->variant = 1;
  if (flavour == ABCOCT_INI_SPINJNT) {
    o->clamp_db = 30.0;   // BscanFFTspinjnt.cpp:1886
    o->output_rebin = 1;  // BscanFFTspinjnt.cpp:1856-1862
  }
  std::ifstream in(path);
  if (!in.is_open()) return ABCOCT_ERR_IO;  // "Unable to open ini file, using defaults." BscanFFT.cpp:484
  enum F { SKIP, BPP, W, H, BIN, BINX, BINY, OBINX, OBINY, AVG, NFFT, MOVAVG, NDISP, LMIN, LMAX, MEDIAN, MULT, ROWNORM, NONORM, BANDPASS, LOWPASS, CHANNEL };
  std::vector<F> order = {SKIP /*camgain*/, SKIP /*camtime*/, BPP, W, H};
  const bool offsets = flavour == ABCOCT_INI_BSCANFFT || flavour == ABCOCT_INI_SPINJ || flavour == ABCOCT_INI_SPINJNT;
  if (offsets) {
    order.push_back(SKIP);  // offsetx
    order.push_back(SKIP);  // offsety
  }
  for (int i = 0; i < 4; ++i) order.push_back(SKIP);  // camspeed cambinx cambiny usbtraffic
  if (flavour == ABCOCT_INI_SPINJNT) {
    order.insert(order.end(), {BINX, BINY, OBINX, OBINY});
  } else {
    order.push_back(BIN);
  }
  order.insert(order.end(), {SKIP /*dirdescr*/, AVG, NFFT, SKIP /*saveframes*/, SKIP /*manualaveraging*/, SKIP /*manualaverages*/,
                             SKIP /*saveinterferograms*/, MOVAVG, NDISP, LMIN, LMAX, MEDIAN, MULT});
  if (flavour != ABCOCT_INI_SIM) order.insert(order.end(), {ROWNORM, NONORM});
  if (flavour == ABCOCT_INI_DARK) order.insert(order.end(), {BANDPASS, LOWPASS});
  if (flavour == ABCOCT_INI_WEBCAM) order.push_back(CHANNEL);  // BscanFFTwebcam.cpp:508
  std::string tok;
  for (int i = 0; i < 3; ++i)  // "first three lines of ini file are comments" BscanFFT.cpp:420-423
    if (!(in >> tok)) return ABCOCT_OK;
  bool first = true;
  for (F f : order) {
    if (!first && !(in >> tok)) break;  // the comment token between values
    first = false;
    if (!(in >> tok)) break;            // the value token
    char* end = nullptr;
    const long iv = strtol(tok.c_str(), &end, 10);
    const bool numeric = end != tok.c_str();
    const double dv = atof(tok.c_str());  // BscanFFT.cpp:479-480 (atof on the lambda strings)
    bool stop = false;
    switch (f) {
      case SKIP: break;
      case BPP: o->bpp = (uint32_t)iv; stop = !numeric; break;
      case W: o->w = (uint32_t)iv; stop = !numeric; break;
      case H: o->h = (uint32_t)iv; stop = !numeric; break;
      case BIN: o->binx = o->biny = (uint32_t)iv; stop = !numeric; break;
      case BINX: o->binx = (uint32_t)iv; stop = !numeric; break;
      case BINY: o->biny = (uint32_t)iv; stop = !numeric; break;
      case OBINX: o->bscanbinx = (uint8_t)std::min(255L, std::max(0L, iv)); stop = !numeric; break;
      case OBINY: o->bscanbiny = (uint8_t)std::min(255L, std::max(0L, iv)); stop = !numeric; break;
      case AVG: o->averages = (uint32_t)iv; stop = !numeric; break;
      case NFFT: o->numfftpoints = (uint32_t)iv; stop = !numeric; break;
      case MOVAVG: o->movavgn = (int32_t)iv; stop = !numeric; break;
      case NDISP: o->numdisplaypoints = (uint32_t)iv; stop = !numeric; break;
      case LMIN: o->lambdamin = dv; break;
      case LMAX: o->lambdamax = dv; break;
      case MEDIAN: o->mediann = (int32_t)iv; stop = !numeric; break;
      case MULT: o->fft_multiplier = (uint32_t)iv; stop = !numeric; break;
      case ROWNORM: o->rowwisenormalize = iv != 0; stop = !numeric; break;
      case NONORM: o->donotnormalize = iv != 0; stop = !numeric; break;
      case BANDPASS: o->bandpassfilter = iv != 0; stop = !numeric; break;
      case LOWPASS: o->lowpassfilter = iv != 0; stop = !numeric; break;
      case CHANNEL: o->channelnum = (uint8_t)std::min(255L, std::max(0L, iv)); stop = !numeric; break;
    }
    if (stop) break;  // an istream in the fail state ignores every later extraction
  }
  return ABCOCT_OK;
}

const char* abcoct_last_error(const abcoct_ctx* c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int abcoct_create(const abcoct_params* params, const int* gpu_ids, int ngpu, abcoct_ctx** out) {
  if (!params || !out) return fail(nullptr, ABCOCT_ERR_INVALID, "null argument");
  *out = nullptr;
  std::string why;
  int code = 0;
  if (validate(*params, why, code)) return fail(nullptr, code, "%s", why.c_str());
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(nullptr, ABCOCT_ERR_CUDA, "no CUDA device: this library has no CPU fallback");
  }
  if (ngpu < 1) ngpu = 1;
  if (ngpu > ndev && !gpu_ids) return fail(nullptr, ABCOCT_ERR_INVALID, "ngpu = %d but only %d devices are visible", ngpu, ndev);
  abcoct_ctx* c = new abcoct_ctx();
  c->p = *params;
  c->opw = params->w / params->binx;
  c->oph = params->h / params->biny;
  c->M = params->fft_multiplier * c->opw;
  c->N = params->numfftpoints;
  c->D = params->numdisplaypoints;
  c->A = params->averages;
  c->plan = find_plan(c->N);
  c->px_bytes = params->bpp == 8 ? 1 : 2;
  c->bgr = params->channelnum >= 3;
  if (c->bgr) {
    c->px_bytes = 3;
    c->px_scale = (float)0.00130718954;  // 1 / 255 / 3, BscanFFTwebcam.cpp:1036
  }
  c->general = params->bpp == 8 || params->binx > 1 || params->biny > 1 || params->mediann > 0 || params->movavgn > 0 ||
               params->fft_multiplier > 1 || params->rowwisenormalize || !params->donotnormalize;  // (8-bit covers channelnum >= 3)
  if (params->output_rebin && params->binx > 1)
    c->db_shift = (float)(20.0 * (1.0 / 2.303) * std::log((double)(params->bscanbinx * params->bscanbiny * params->binx * params->biny)));
  c->generic = needs_generic(*params);
  if (c->generic) {
    c->general = true;  // rowprep_kernel prepares f32 rows for every configuration, generic_recon_kernel consumes them
    c->plan = nullptr;
    c->radN = factor_radices(c->N);
  }
  if (params->fft_multiplier > 1) {
    c->radW = factor_radices(c->opw);
    c->radM = factor_radices(c->M);
  }
  build_ref_tables(c->opw, params->fft_multiplier, c->N, params->lambdamin, params->lambdamax, c->nk, c->frac);
  build_window(c->opw, c->win);
  // gather tables for the kernel: end points q = 0, N-1 are never written in the reference (BscanFFT.cpp:1164)
  std::vector<int> idx(c->N);
  std::vector<float> wq(c->N), winf(c->opw);
  std::vector<double> wq64(c->N);
  c->single = single_row_regime(*params);
  for (int q = 0; q < c->N; ++q) {
    int i = c->nk[q];
    double w = params->weight_mode == 0 ? c->frac[c->nk[q]] : c->frac[q];  // :1170 quirk vs corrected
    if (q == 0 || q == c->N - 1) {
      i = c->M;  // sentinel slot holding zero
      w = 0.0;
    } else if (i == 0) {  // slopes[0] = slopes[1] (:1161): y0 + w (y1 - y0) == y1 + (w - 1)(y1 - y0)
      i = 1;
      w = w - 1.0;
    }
    idx[q] = i;
    wq[q] = (float)w;
    wq64[q] = w;
  }
  for (int i = 0; i < c->opw; ++i) winf[i] = (float)c->win[i];
  c->gidx = idx;
  c->gwq = wq;
  std::vector<unsigned char>& blob = c->blob1;
  if (c->generic) {
    // gidx / gwq go to the device as they are
  } else if (c->general) {  // the fused kernel sees rows of M apodised samples: no window / mean term left to apply there
    std::vector<float> zero(c->M, 0.f);
    c->plan->build_blob(c->M, idx.data(), wq.data(), zero.data(), blob);
  } else {
    c->plan->build_blob(c->opw, idx.data(), wq.data(), winf.data(), blob);
  }
  // warp-per-A-scan kernel: 16-bit frames straight into the transform (no optional pre-processing stage), the reference's
  // source-indexed resampling weight (BscanFFT.cpp:1170), a transform length it has a plan for, and no gather from sample 0
  // (whose slope the reference copies from sample 1, :1161 - the fallback kernel handles that corner)
  c->wrow_eligible = !c->general && !c->generic && params->weight_mode == 0 && params->bpp == 16 && find_wplan(c->N, 0, -1) != nullptr;
  for (int q = 1; q + 1 < c->N && c->wrow_eligible; ++q) c->wrow_eligible = c->nk[q] >= 1 && c->nk[q] < c->opw;
  if (c->wrow_eligible) {
    std::vector<int> widx(c->N);
    for (int q = 0; q < c->N; ++q) widx[q] = (q == 0 || q == c->N - 1) ? -1 : c->nk[q];
    WrowTablesHost wt{c->opw, widx.data(), c->frac.data(), c->win.data()};
    find_wplan(c->N, 0, -1)->build_blob(wt, c->wblob);  // the blob does not depend on the warps per CTA
    if (const WPlanEntry* rp = find_rplan(c->N, 0)) rp->build_blob(wt, c->rblob);  // nor on the warps / slots of a resident plan
  }
  std::vector<float2> twW, twM;
  if (params->fft_multiplier > 1) {
    const double tau = 6.283185307179586476925286766559;
    twW.resize(c->opw);
    twM.resize(c->M);
    for (int k = 0; k < c->opw; ++k) twW[k] = make_float2((float)std::cos(tau * k / c->opw), (float)-std::sin(tau * k / c->opw));
    for (int k = 0; k < c->M; ++k) twM[k] = make_float2((float)std::cos(tau * k / c->M), (float)std::sin(tau * k / c->M));
  }
  std::vector<float2> twN;
  if (c->generic) {
    const double tau = 6.283185307179586476925286766559;
    twN.resize(c->N);
    for (int k = 0; k < c->N; ++k) twN[k] = make_float2((float)std::cos(tau * k / c->N), (float)std::sin(tau * k / c->N));
  }

  c->gpus.resize(ngpu);
  for (int i = 0; i < ngpu; ++i) {
    GpuState& g = c->gpus[i];
    g.dev = gpu_ids ? gpu_ids[i] : i;
    cudaError_t e = cudaSetDevice(g.dev);
    cudaDeviceProp prop{};
    if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, g.dev);
    if (e != cudaSuccess) {
      fail(nullptr, ABCOCT_ERR_CUDA, "cudaSetDevice(%d): %s", g.dev, cudaGetErrorString(e));
      abcoct_destroy(c);
      return ABCOCT_ERR_CUDA;
    }
    if (prop.major != 10) {
      fail(nullptr, ABCOCT_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", g.dev, prop.major, prop.minor);
      abcoct_destroy(c);
      return ABCOCT_ERR_CUDA;
    }
    g.sm_count = prop.multiProcessorCount;
    bool ok = true;
    for (int s = 0; s < kSlots && ok; ++s) {
      ok = ok && cudaStreamCreateWithFlags(&g.stream[s], cudaStreamNonBlocking) == cudaSuccess;
      ok = ok && cudaEventCreateWithFlags(&g.slot_done[s], cudaEventDisableTiming) == cudaSuccess;
    }
    g.tev.assign(3 * kTimedChunks, nullptr);
    for (size_t k = 0; k < g.tev.size() && ok; ++k) ok = ok && cudaEventCreate(&g.tev[k]) == cudaSuccess;
    ok = ok && post_init_device() == cudaSuccess;
    {
      ok = ok && cudaMalloc(&g.d_win, winf.size() * 4) == cudaSuccess;
      ok = ok && cudaMemcpy(g.d_win, winf.data(), winf.size() * 4, cudaMemcpyHostToDevice) == cudaSuccess;
      if (!twW.empty()) {
        ok = ok && cudaMalloc(&g.d_twW, twW.size() * 8) == cudaSuccess && cudaMalloc(&g.d_twM, twM.size() * 8) == cudaSuccess;
        ok = ok && cudaMemcpy(g.d_twW, twW.data(), twW.size() * 8, cudaMemcpyHostToDevice) == cudaSuccess;
        ok = ok && cudaMemcpy(g.d_twM, twM.data(), twM.size() * 8, cudaMemcpyHostToDevice) == cudaSuccess;
      }
      if (c->single) {
        ok = ok && cudaMalloc(&g.d_gwq64, wq64.size() * 8) == cudaSuccess && cudaMalloc(&g.d_win64, c->win.size() * 8) == cudaSuccess;
        ok = ok && cudaMemcpy(g.d_gwq64, wq64.data(), wq64.size() * 8, cudaMemcpyHostToDevice) == cudaSuccess;
        ok = ok && cudaMemcpy(g.d_win64, c->win.data(), c->win.size() * 8, cudaMemcpyHostToDevice) == cudaSuccess;
      }
      if (c->generic) {
        ok = ok && cudaMalloc(&g.d_gidx, idx.size() * 4) == cudaSuccess && cudaMalloc(&g.d_gwq, wq.size() * 4) == cudaSuccess &&
             cudaMalloc(&g.d_twN, twN.size() * 8) == cudaSuccess;
        ok = ok && cudaMemcpy(g.d_gidx, idx.data(), idx.size() * 4, cudaMemcpyHostToDevice) == cudaSuccess;
        ok = ok && cudaMemcpy(g.d_gwq, wq.data(), wq.size() * 4, cudaMemcpyHostToDevice) == cudaSuccess;
        ok = ok && cudaMemcpy(g.d_twN, twN.data(), twN.size() * 8, cudaMemcpyHostToDevice) == cudaSuccess;
      }
    }
    if (!ok) {
      fail(nullptr, ABCOCT_ERR_CUDA, "device %d setup failed: %s", g.dev, cudaGetErrorString(cudaGetLastError()));
      abcoct_destroy(c);
      return ABCOCT_ERR_CUDA;
    }
  }
  *out = c;
  return ABCOCT_OK;
}

void abcoct_destroy(abcoct_ctx* c) {
  if (!c) return;
  for (GpuState& g : c->gpus) {
    cudaSetDevice(g.dev);
    cudaDeviceSynchronize();
    for (int s = 0; s < kSlots; ++s) {
      if (g.stream[s]) cudaStreamDestroy(g.stream[s]);
      if (g.slot_done[s]) cudaEventDestroy(g.slot_done[s]);
      cudaFree(g.d_in[s]);
      for (int k = 0; k < O_COUNT; ++k) {
        cudaFree(g.d_o[k][s]);
        if (g.h_o[k][s]) cudaFreeHost(g.h_o[k][s]);
      }
      cudaFree(g.d_jmm[s]);
      if (s == 0) cudaFree(g.d_jscan);
      cudaFree(g.d_tmp[0][s]);
      cudaFree(g.d_tmp[1][s]);
      cudaFree(g.d_dc01[s]);
      cudaFree(g.d_scratch[s]);
      cudaFree(g.d_sched[s]);
      if (g.h_in[s]) cudaFreeHost(g.h_in[s]);
    }
    for (cudaEvent_t e : g.tev)
      if (e) cudaEventDestroy(e);
    cudaFree(g.d_tables);
    cudaFree(g.d_gain);
    cudaFree(g.d_subg);
    cudaFree(g.d_yb);
    cudaFree(g.d_yp);
    cudaFree(g.d_yd);
    cudaFree(g.d_win);
    cudaFree(g.d_twW);
    cudaFree(g.d_twM);
    cudaFree(g.d_yb64);
    cudaFree(g.d_yp64);
    cudaFree(g.d_yd64);
    cudaFree(g.d_win64);
    cudaFree(g.d_gwq64);
    for (int s2 = 0; s2 < kSlots; ++s2) {
      cudaFree(g.d_rows64[s2]);
      cudaFree(g.d_pre32[s2]);
      cudaFree(g.d_fmm64[s2]);
    }
    cudaFree(g.d_gidx);
    cudaFree(g.d_gwq);
    cudaFree(g.d_twN);
    for (int s2 = 0; s2 < kSlots; ++s2) {
      cudaFree(g.d_med[s2]);
      cudaFree(g.d_bin[s2]);
      cudaFree(g.d_rows[s2]);
      cudaFree(g.d_fmm[s2]);
    }
  }
  cudaGetLastError();
  delete c;
}

int abcoct_set_threshold(abcoct_ctx* c, double thr) {
  if (!c) return ABCOCT_ERR_INVALID;
  c->p.bscanthreshold = thr;
  return ABCOCT_OK;
}
int abcoct_set_clampupper(abcoct_ctx* c, int on) {
  if (!c) return ABCOCT_ERR_INVALID;
  if (on && c->oph < 6) return fail(c, ABCOCT_ERR_INVALID, "clampupper needs at least 6 A-scans (element (5,5), BscanFFT.cpp:1252)");
  c->p.clampupper = on ? 1 : 0;
  return ABCOCT_OK;
}
int abcoct_set_averages(abcoct_ctx* c, uint32_t averages) {
  if (!c) return ABCOCT_ERR_INVALID;
  if (averages == 0) return fail(c, ABCOCT_ERR_INVALID, "averages must be >= 1");
  c->p.averages = averages;
  c->A = (int)averages;
  c->cal_dirty = true;  // the averages == 1 kernel variant has its own attributes
  return ABCOCT_OK;
}

static int set_cal(abcoct_ctx* c, std::vector<double>& dst, bool& have, const double* src, size_t ld) {
  if (!c) return ABCOCT_ERR_INVALID;
  const size_t n = (size_t)c->oph * c->opw;
  if (!src) {
    dst.clear();
    have = false;
  } else {
    if (ld == 0) ld = c->opw;
    if (ld < (size_t)c->opw) return fail(c, ABCOCT_ERR_INVALID, "ld < opw");
    dst.resize(n);
    for (int r = 0; r < c->oph; ++r) memcpy(&dst[(size_t)r * c->opw], src + (size_t)r * ld, (size_t)c->opw * sizeof(double));
    have = true;
  }
  c->cal_dirty = true;
  return ABCOCT_OK;
}
int abcoct_set_background(abcoct_ctx* c, const double* yb, size_t ld) {
  if (c && !yb) return fail(c, ABCOCT_ERR_INVALID, "background must not be NULL");
  return c ? set_cal(c, c->yb, c->have_yb, yb, ld) : ABCOCT_ERR_INVALID;
}
int abcoct_set_pishift(abcoct_ctx* c, const double* yp, size_t ld) { return c ? set_cal(c, c->yp, c->have_yp, yp, ld) : ABCOCT_ERR_INVALID; }
int abcoct_set_dark(abcoct_ctx* c, const double* yd, size_t ld) { return c ? set_cal(c, c->yd, c->have_yd, yd, ld) : ABCOCT_ERR_INVALID; }

int abcoct_set_calibration_from_frames(abcoct_ctx* c, int which, const void* frames, size_t nframes, size_t stride_bytes) {
  if (!c || !frames || nframes == 0 || which < 0 || which > 4) return c ? fail(c, ABCOCT_ERR_INVALID, "bad argument") : ABCOCT_ERR_INVALID;
  const int pb = c->px_bytes, w = (int)c->p.w, h = (int)c->p.h;
  if (stride_bytes == 0) stride_bytes = (size_t)w * pb;
  if (stride_bytes % pb) return fail(c, ABCOCT_ERR_INVALID, "stride_bytes must be a multiple of the pixel size");
  if (which == 1 && nframes != 1) return fail(c, ABCOCT_ERR_INVALID, "the pi-shifted frame is a copy of ONE frame (BscanFFT.cpp:1081)");
  const size_t n = (size_t)c->oph * c->opw;
  std::vector<double> acc(n, 0.0);
  {
    // accumulate(data_y, baccum) over the frames (BscanFFT.cpp:1041-1046) on the GPU, with the very kernels of the processing path
    // for the integer stages (channel sum, medianBlur, INTER_AREA binning) and an f64 kernel for convertTo + smoothmovavg + the sum
    GpuState& g = c->gpus[0];
    CU(c, cudaSetDevice(g.dev));
    cudaStream_t st = g.stream[0];
    const size_t in_bytes = nframes * (size_t)h * stride_bytes;
    const int ipb = c->bgr ? 2 : pb;  // bytes per pixel after the channel sum
    uint8_t *d_in = nullptr, *d_a = nullptr, *d_b = nullptr;
    double *d_acc = nullptr, *d_cs = nullptr, *d_sn = nullptr;
    long long* d_mm = nullptr;
    auto cleanup = [&]() {
      cudaFree(d_in);
      cudaFree(d_a);
      cudaFree(d_b);
      cudaFree(d_acc);
      cudaFree(d_cs);
      cudaFree(d_sn);
      cudaFree(d_mm);
    };
    cudaError_t e = cudaMalloc(&d_in, in_bytes);
    if (e == cudaSuccess) e = cudaMalloc(&d_acc, n * sizeof(double));
    if (e == cudaSuccess && (c->bgr || c->p.mediann > 0)) e = cudaMalloc(&d_a, nframes * (size_t)w * h * ipb);
    if (e == cudaSuccess && (c->p.binx > 1 || c->p.biny > 1)) e = cudaMalloc(&d_b, nframes * n * ipb);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_in, frames, in_bytes, cudaMemcpyHostToDevice, st);
    const void* src = d_in;
    size_t rs = stride_bytes / pb, fs = (size_t)h * stride_bytes / pb;
    int bpp = (int)c->p.bpp;
    if (e == cudaSuccess && c->bgr) {  // BscanFFTwebcam.cpp:1021-1037 (mediann == 0 and no binning, validated at create)
      e = launch_bgr_sum(d_in, reinterpret_cast<uint16_t*>(d_a), w, h, stride_bytes, (size_t)h * stride_bytes, (int)nframes, st);
      src = d_a;
      rs = w;
      fs = (size_t)w * h;
      bpp = 16;
    }
    if (e == cudaSuccess && c->p.mediann > 0) {  // medianBlur while the numbers are still integers, BscanFFT.cpp:953-954
      e = launch_median(src, d_a, bpp, c->p.mediann, w, h, rs, fs, (int)nframes, st);
      src = d_a;
      rs = w;
      fs = (size_t)w * h;
    }
    if (e == cudaSuccess && (c->p.binx > 1 || c->p.biny > 1)) {  // resize(..., INTER_AREA), BscanFFT.cpp:958
      e = launch_bin(src, d_b, bpp, c->opw, c->oph, (int)c->p.binx, (int)c->p.biny, rs, fs, (int)nframes, st);
      src = d_b;
      rs = c->opw;
      fs = n;
    }
    if (e == cudaSuccess)  // convertTo + smoothmovavg (BscanFFT.cpp:987-991) + accumulate, f64
      e = launch_cal_accum(src, bpp, rs, fs, (int)nframes, c->opw, c->oph, c->p.movavgn, c->bgr ? 0.00130718954 : 1.0, d_acc, st);
    // the once-per-capture tail, on the device too (cal_*_kernel, f64, the reference's operation order):
    //   keys b / o / r / t, BscanFFT.cpp:1050-1057: if (rowwisenormalize) normalizerows(.., 0.0001, 1); if (!donotnormalize)
    //     normalize(.., 0.0001, 1) else / n (Mat / double multiplies by the reciprocal); BscanDark.cpp:1070-1074, 1145-1149, 1218-1222:
    //     lpfilter on the dark / reference / sample captures;
    //   key p, BscanFFT.cpp:1092-1096: the copy is normalised to [0, 1] like data_y is (:1126-1129) - no 0.0001 floor, no division
    CalTailHost t{};
    t.x = d_acc;
    t.rows = c->oph;
    t.cols = c->opw;
    t.rowwise = c->p.rowwisenormalize ? 1 : 0;
    t.global_norm = c->p.donotnormalize ? 0 : 1;
    t.lo = which == 1 ? 0.0 : 0.0001;
    t.inv_n = which == 1 ? 1.0 : 1.0 / (double)nframes;
    t.lowpass = (which >= 2 && c->p.lowpassfilter) ? 1 : 0;
    if (e == cudaSuccess) e = cudaMalloc(&d_mm, 2 * sizeof(long long));
    t.mm = d_mm;
    if (e == cudaSuccess && t.lowpass) {
      const double tau = 6.283185307179586476925286766559;
      std::vector<double> cs(c->opw), sn(c->opw);
      for (int k = 0; k < c->opw; ++k) {
        cs[k] = std::cos(tau * k / c->opw);
        sn[k] = std::sin(tau * k / c->opw);
      }
      e = cudaMalloc(&d_cs, cs.size() * 8);
      if (e == cudaSuccess) e = cudaMalloc(&d_sn, sn.size() * 8);
      if (e == cudaSuccess) e = cudaMemcpy(d_cs, cs.data(), cs.size() * 8, cudaMemcpyHostToDevice);
      if (e == cudaSuccess) e = cudaMemcpy(d_sn, sn.data(), sn.size() * 8, cudaMemcpyHostToDevice);
      t.cs = d_cs;
      t.sn = d_sn;
    }
    int ntail = 0;
    if (e == cudaSuccess) e = launch_cal_tail(t, st, &ntail);
    if (e == cudaSuccess) e = cudaMemcpyAsync(acc.data(), d_acc, n * sizeof(double), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cleanup();
    if (e != cudaSuccess) return fail(c, ABCOCT_ERR_CUDA, "calibration capture: %s", cudaGetErrorString(e));
    c->launches += 1 + ntail + (c->bgr ? 1 : 0) + (c->p.mediann > 0 ? 1 : 0) + ((c->p.binx > 1 || c->p.biny > 1) ? 1 : 0);
  }
  std::vector<double>& dst = which == 0 ? c->yb : which == 1 ? c->yp : which == 2 ? c->yd : which == 3 ? c->yr : c->ys;
  bool& have = which == 0 ? c->have_yb : which == 1 ? c->have_yp : which == 2 ? c->have_yd : which == 3 ? c->have_yr : c->have_ys;
  dst.swap(acc);
  have = true;
  if (which <= 2) c->cal_dirty = true;
  return ABCOCT_OK;
}

int abcoct_get_calibration(const abcoct_ctx* c, int which, double* out, size_t ld) {
  if (!c || !out || which < 0 || which > 4) return ABCOCT_ERR_INVALID;
  const std::vector<double>& src = which == 0 ? c->yb : which == 1 ? c->yp : which == 2 ? c->yd : which == 3 ? c->yr : c->ys;
  const bool have = which == 0 ? c->have_yb : which == 1 ? c->have_yp : which == 2 ? c->have_yd : which == 3 ? c->have_yr : c->have_ys;
  if (!have) return ABCOCT_ERR_STATE;
  if (ld == 0) ld = c->opw;
  if (ld < (size_t)c->opw) return ABCOCT_ERR_INVALID;
  for (int r = 0; r < c->oph; ++r) memcpy(out + (size_t)r * ld, &src[(size_t)r * c->opw], (size_t)c->opw * sizeof(double));
  return ABCOCT_OK;
}

int abcoct_compose_dark_background(abcoct_ctx* c) {
  if (!c) return ABCOCT_ERR_INVALID;
  if (!c->have_yr || !c->have_ys || !c->have_yd)
    return fail(c, ABCOCT_ERR_STATE, "need the dark, reference-arm and sample-arm captures first (BscanDark.cpp:996)");
  const size_t n = (size_t)c->oph * c->opw;
  c->yb.resize(n);
  for (size_t i = 0; i < n; ++i) c->yb[i] = (c->yr[i] - c->yd[i]) + (c->ys[i] - c->yd[i]);
  c->have_yb = true;
  c->cal_dirty = true;
  return ABCOCT_OK;
}

int abcoct_build_tables(const abcoct_params* p, int32_t* nk, double* frac, double* win) {
  if (!p || p->binx == 0 || p->w / p->binx == 0 || p->fft_multiplier == 0 || p->numfftpoints == 0) return ABCOCT_ERR_INVALID;
  const int opw = p->w / p->binx;
  if (p->numfftpoints < p->fft_multiplier * (uint32_t)opw) return ABCOCT_ERR_INVALID;
  std::vector<int32_t> vnk;
  std::vector<double> vfrac, vwin;
  build_ref_tables(opw, p->fft_multiplier, p->numfftpoints, p->lambdamin, p->lambdamax, vnk, vfrac);
  build_window(opw, vwin);
  if (nk) memcpy(nk, vnk.data(), vnk.size() * sizeof(int32_t));
  if (frac) memcpy(frac, vfrac.data(), vfrac.size() * sizeof(double));
  if (win) memcpy(win, vwin.data(), vwin.size() * sizeof(double));
  return ABCOCT_OK;
}

int abcoct_get_tables(const abcoct_ctx* c, int32_t* nk, double* frac) {
  if (!c) return ABCOCT_ERR_INVALID;
  if (nk) memcpy(nk, c->nk.data(), c->nk.size() * sizeof(int32_t));
  if (frac) memcpy(frac, c->frac.data(), c->frac.size() * sizeof(double));
  return ABCOCT_OK;
}
int abcoct_get_window(const abcoct_ctx* c, double* w) {
  if (!c || !w) return ABCOCT_ERR_INVALID;
  memcpy(w, c->win.data(), c->win.size() * sizeof(double));
  return ABCOCT_OK;
}

int abcoct_process_bscans_device_ex(abcoct_ctx* c, int gi, const void* d_frames, size_t nframes, size_t stride_bytes,
                                    const abcoct_outputs* d_out, void* cuda_stream) {
  if (!c) return ABCOCT_ERR_INVALID;
  if (gi < 0 || gi >= (int)c->gpus.size()) return fail(c, ABCOCT_ERR_INVALID, "gpu_index out of range");
  if (!d_frames || !d_out || !d_out->bscan_u8) return fail(c, ABCOCT_ERR_INVALID, "null buffer");
  if (d_out->reserved[0] || d_out->reserved[1]) return fail(c, ABCOCT_ERR_INVALID, "abcoct_outputs.reserved must be NULL");
  if (nframes == 0 || nframes % c->A) return fail(c, ABCOCT_ERR_INVALID, "nframes must be a positive multiple of averages (%d)", c->A);
  if (stride_bytes == 0) stride_bytes = (size_t)c->p.w * c->px_bytes;
  if (!c->general && (stride_bytes % 16 || (reinterpret_cast<uintptr_t>(d_frames) & 15)))
    return fail(c, ABCOCT_ERR_INVALID, "device frames must be 16-byte aligned with a 16-byte multiple row stride");
  if (c->cal_dirty) {
    int rc = upload_calibration(c);
    if (rc) return rc;
  }
  GpuState& g = c->gpus[gi];
  CU(c, cudaSetDevice(g.dev));
  cudaStream_t st = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : g.stream[0];
  OutPtrs o;
  o.p[O_U8] = d_out->bscan_u8;
  o.p[O_DB] = d_out->bscan_db;
  o.p[O_LIN] = d_out->bscan_lin;
  o.p[O_BGR] = d_out->bscan_bgr;
  o.p[O_JSUB] = d_out->jsub_u8;
  o.p[O_JBGR] = d_out->jsub_bgr;
  int rc = enqueue_device(c, g, 0, static_cast<const uint8_t*>(d_frames), nframes / c->A, stride_bytes, stride_bytes * c->p.h, o, st, true);
  if (rc) return rc;
  if (!cuda_stream) CU(c, cudaStreamSynchronize(st));
  return ABCOCT_OK;
}

int abcoct_process_bscans_device(abcoct_ctx* c, int gi, const void* d_frames, size_t nframes, size_t stride_bytes, uint8_t* d_u8,
                                 float* d_db, void* cuda_stream) {
  abcoct_outputs o{};
  o.bscan_u8 = d_u8;
  o.bscan_db = d_db;
  return abcoct_process_bscans_device_ex(c, gi, d_frames, nframes, stride_bytes, &o, cuda_stream);
}

int abcoct_set_jscan(abcoct_ctx* c, const float* jscan, size_t ld) {
  if (!c) return ABCOCT_ERR_INVALID;
  if (!jscan) {  // key 'c': lock-in off
    c->jscan.clear();
    c->jscan_dirty = false;
    return ABCOCT_OK;
  }
  if (ld == 0) ld = (size_t)c->oph;
  if (ld < (size_t)c->oph) return fail(c, ABCOCT_ERR_INVALID, "ld < oph");
  c->jscan.resize((size_t)c->D * c->oph);
  for (int d = 0; d < c->D; ++d) memcpy(&c->jscan[(size_t)d * c->oph], jscan + (size_t)d * ld, (size_t)c->oph * sizeof(float));
  c->jscan_dirty = true;
  return ABCOCT_OK;
}

int abcoct_timing_reset(abcoct_ctx* c) {
  if (!c) return ABCOCT_ERR_INVALID;
  for (GpuState& g : c->gpus) g.tev_used = 0;
  return ABCOCT_OK;
}

int abcoct_timing_read(abcoct_ctx* c, int gi, uint32_t* nchunks, double* recon_ms, double* norm_ms) {
  if (!c) return ABCOCT_ERR_INVALID;
  if (gi < 0 || gi >= (int)c->gpus.size()) return fail(c, ABCOCT_ERR_INVALID, "gpu_index out of range");
  GpuState& g = c->gpus[gi];
  CU(c, cudaSetDevice(g.dev));
  double r = 0, n = 0;
  for (size_t k = 0; k + 3 <= g.tev_used; k += 3) {
    float a = 0, b = 0;
    CU(c, cudaEventSynchronize(g.tev[k + 2]));
    CU(c, cudaEventElapsedTime(&a, g.tev[k], g.tev[k + 1]));
    CU(c, cudaEventElapsedTime(&b, g.tev[k + 1], g.tev[k + 2]));
    r += a;
    n += b;
  }
  if (nchunks) *nchunks = (uint32_t)(g.tev_used / 3);
  if (recon_ms) *recon_ms = r;
  if (norm_ms) *norm_ms = n;
  if (g.tev_used) {
    c->last_recon_ms = r / (g.tev_used / 3);
    c->last_norm_ms = n / (g.tev_used / 3);
  }
  return ABCOCT_OK;
}

int abcoct_process_bscans(abcoct_ctx* c, const void* frames, size_t nframes, size_t stride_bytes, uint8_t* out8, float* outdb) {
  abcoct_outputs o{};
  o.bscan_u8 = out8;
  o.bscan_db = outdb;
  return abcoct_process_bscans_ex(c, frames, nframes, stride_bytes, &o);
}

int abcoct_process_bscans_ex(abcoct_ctx* c, const void* frames, size_t nframes, size_t stride_bytes, const abcoct_outputs* out) {
  if (!c) return ABCOCT_ERR_INVALID;
  if (!frames || !out || !out->bscan_u8) return fail(c, ABCOCT_ERR_INVALID, "null buffer");
  if (out->reserved[0] || out->reserved[1]) return fail(c, ABCOCT_ERR_INVALID, "abcoct_outputs.reserved must be NULL");
  if (nframes == 0 || nframes % c->A) return fail(c, ABCOCT_ERR_INVALID, "nframes must be a positive multiple of averages (%d)", c->A);
  OutPtrs host;
  host.p[O_U8] = out->bscan_u8;
  host.p[O_DB] = out->bscan_db;
  host.p[O_LIN] = out->bscan_lin;
  host.p[O_BGR] = out->bscan_bgr;
  host.p[O_JSUB] = out->jsub_u8;
  host.p[O_JBGR] = out->jsub_bgr;
  const unsigned want = host.mask();
  if ((want & ((1u << O_JSUB) | (1u << O_JBGR))) && c->jscan.empty())
    return fail(c, ABCOCT_ERR_STATE, "jsub output requested but no jscan is set (abcoct_set_jscan)");
  const size_t dense = (size_t)c->p.w * c->px_bytes;
  if (stride_bytes == 0) stride_bytes = dense;
  if (stride_bytes < dense) return fail(c, ABCOCT_ERR_INVALID, "stride_bytes < w * bytes per pixel");
  if (c->cal_dirty) {
    int rc = upload_calibration(c);
    if (rc) return rc;
  }
  if (c->jscan_dirty) {
    int rc = upload_jscan(c);
    if (rc) return rc;
  }
  const size_t nB = nframes / c->A;
  const size_t ngpu = c->gpus.size();
  const size_t frame_in = stride_bytes * c->p.h;          // caller layout
  const size_t frame_dev = dense * c->p.h;                // dense on the device
  const size_t bscan_in_dev = frame_dev * c->A;
  const size_t out_px = (size_t)c->D * c->oph;
  // slot size: about 64 MiB of input per slot, at least one B-scan, and enough slots to cover all GPUs
  size_t slotB = std::max<size_t>(1, (64u << 20) / bscan_in_dev);
  slotB = std::min(slotB, std::max<size_t>(1, (nB + ngpu * kSlots - 1) / (ngpu * kSlots)));
  for (GpuState& g : c->gpus) {
    int rc = ensure_slots(c, g, slotB, want);
    if (rc) return rc;
  }
  const bool in_pinned = is_pinned(frames) && stride_bytes == dense;
  bool out_pinned = true;
  for (int k = 0; k < O_COUNT; ++k) out_pinned = out_pinned && (!host.p[k] || is_pinned(host.p[k]));
  struct Pending {
    size_t b0 = 0, nb = 0;
    bool busy = false;
  };
  std::vector<Pending> pend(ngpu * kSlots);
  auto drain = [&](size_t gi, int s) -> int {
    Pending& pd = pend[gi * kSlots + s];
    if (!pd.busy) return ABCOCT_OK;
    GpuState& g = c->gpus[gi];
    CU(c, cudaSetDevice(g.dev));
    CU(c, cudaEventSynchronize(g.slot_done[s]));
    if (!out_pinned)
      for (int k = 0; k < O_COUNT; ++k)
        if (host.p[k]) staging_copy(static_cast<uint8_t*>(host.p[k]) + pd.b0 * out_px * kOutBpp[k], g.h_o[k][s], pd.nb * out_px * kOutBpp[k]);
    pd.busy = false;
    return ABCOCT_OK;
  };
  // on any error: wait for everything already enqueued - the header promises that inputs may be freed on return
  auto bail = [&](int rc) -> int {
    for (GpuState& g : c->gpus) {
      cudaSetDevice(g.dev);
      for (int s = 0; s < kSlots; ++s) cudaStreamSynchronize(g.stream[s]);
    }
    cudaGetLastError();
    return rc;
  };
  // One feeder per GPU: chunks gi, gi + ngpu, ... go through that GPU's slot ring (copy in, kernels, copy out, all on the slot's
  // stream).  With several GPUs every feeder is its own host thread - a single loop would serialise the staging copies and the
  // enqueue calls of all GPUs (round 1: 2.5x at 8 GPUs).
  auto feed = [&](size_t gi) -> int {
    GpuState& g = c->gpus[gi];
    CU(c, cudaSetDevice(g.dev));
    size_t k = 0;
    for (size_t chunk = gi; chunk * slotB < nB; chunk += ngpu, ++k) {
      const size_t b0 = chunk * slotB;
      const size_t nb = std::min(slotB, nB - b0);
      const int s = (int)(k % kSlots);
      int rc = drain(gi, s);
      if (rc) return rc;
      cudaStream_t st = g.stream[s];
      const uint8_t* src = static_cast<const uint8_t*>(frames) + b0 * c->A * frame_in;
      const size_t nfr = nb * c->A;
      if (in_pinned) {
        CU(c, cudaMemcpyAsync(g.d_in[s], src, nfr * frame_dev, cudaMemcpyHostToDevice, st));
      } else {
        // pageable (or pitched) caller memory: stage through the pinned ring
        if (stride_bytes == dense) {
          staging_copy(g.h_in[s], src, nfr * frame_dev);
        } else {
          for (size_t r = 0; r < nfr * c->p.h; ++r) memcpy(g.h_in[s] + r * dense, src + r * stride_bytes, dense);
        }
        CU(c, cudaMemcpyAsync(g.d_in[s], g.h_in[s], nfr * frame_dev, cudaMemcpyHostToDevice, st));
      }
      OutPtrs dev;
      for (int q = 0; q < O_COUNT; ++q) dev.p[q] = host.p[q] ? g.d_o[q][s] : nullptr;
      rc = enqueue_device(c, g, s, g.d_in[s], nb, dense, frame_dev, dev, st, false);
      if (rc) return rc;
      for (int q = 0; q < O_COUNT; ++q)
        if (host.p[q]) {
          void* dst = out_pinned ? static_cast<void*>(static_cast<uint8_t*>(host.p[q]) + b0 * out_px * kOutBpp[q]) : g.h_o[q][s];
          CU(c, cudaMemcpyAsync(dst, g.d_o[q][s], nb * out_px * kOutBpp[q], cudaMemcpyDeviceToHost, st));
        }
      CU(c, cudaEventRecord(g.slot_done[s], st));
      pend[gi * kSlots + s] = Pending{b0, nb, true};
    }
    for (int s = 0; s < kSlots; ++s) {
      int rc = drain(gi, s);
      if (rc) return rc;
    }
    return ABCOCT_OK;
  };
  int rc_all = ABCOCT_OK;
  if (ngpu == 1) {
    rc_all = feed(0);
  } else {
    std::vector<int> rcs(ngpu, ABCOCT_OK);
    std::vector<std::thread> th;
    for (size_t gi = 1; gi < ngpu; ++gi) {
      try {
        th.emplace_back([&, gi] { rcs[gi] = feed(gi); });
      } catch (...) {  // no thread to be had: this thread feeds that GPU after its own
        rcs[gi] = -12345;
      }
    }
    rcs[0] = feed(0);
    for (std::thread& t : th) t.join();
    for (size_t gi = 1; gi < ngpu; ++gi)
      if (rcs[gi] == -12345) rcs[gi] = feed(gi);
    for (int r : rcs)
      if (r && !rc_all) rc_all = r;
  }
  if (rc_all) return bail(rc_all);
  return ABCOCT_OK;
}

int abcoct_debug_linearised(abcoct_ctx* c, const void* frame, size_t stride_bytes, float* ylin) {
  if (!c) return ABCOCT_ERR_INVALID;
  if (!frame || !ylin) return fail(c, ABCOCT_ERR_INVALID, "null buffer");
  const size_t dense = (size_t)c->p.w * c->px_bytes;
  if (stride_bytes == 0) stride_bytes = dense;
  if (stride_bytes < dense) return fail(c, ABCOCT_ERR_INVALID, "stride_bytes < w * bytes per pixel");
  if (c->cal_dirty) {
    int rc = upload_calibration(c);
    if (rc) return rc;
  }
  GpuState& g = c->gpus[0];
  CU(c, cudaSetDevice(g.dev));
  cudaStream_t st = g.stream[0];
  // stage the frame densely, then run the general-path kernels for this one frame and the lerp on top
  std::vector<uint8_t> host((size_t)c->p.h * dense);
  for (uint32_t r = 0; r < c->p.h; ++r) memcpy(&host[r * dense], static_cast<const uint8_t*>(frame) + r * stride_bytes, dense);
  uint8_t* d_frame = nullptr;
  int* d_idx = nullptr;
  float *d_wq = nullptr, *d_ylin = nullptr;
  auto cleanup = [&]() {
    cudaFree(d_frame);
    cudaFree(d_idx);
    cudaFree(d_wq);
    cudaFree(d_ylin);
  };
  int rc = ensure_prep(c, g, 0, 1);
  if (rc) return rc;
  cudaError_t e = cudaMalloc(&d_frame, host.size());
  if (e == cudaSuccess) e = cudaMalloc(&d_idx, c->N * sizeof(int));
  if (e == cudaSuccess) e = cudaMalloc(&d_wq, c->N * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&d_ylin, (size_t)c->oph * c->N * sizeof(float));
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_frame, host.data(), host.size(), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_idx, c->gidx.data(), c->N * sizeof(int), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_wq, c->gwq.data(), c->N * sizeof(float), cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) {
    cleanup();
    return fail(c, ABCOCT_ERR_CUDA, "debug tap setup: %s", cudaGetErrorString(e));
  }
  int nl = 0;
  rc = run_prep(c, g, 0, d_frame, 1, dense, dense * c->p.h, st, &nl);
  if (rc == ABCOCT_OK) {
    if (c->single)
      e = launch_lerp_rows64(g.d_rows[0], c->p.fft_multiplier <= 1 ? g.d_rows64[0] : nullptr, d_idx, g.d_gwq64, d_ylin, c->M, c->N, c->oph, st);
    else
      e = launch_lerp_rows(g.d_rows[0], d_idx, d_wq, d_ylin, c->M, c->N, c->oph, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(ylin, d_ylin, (size_t)c->oph * c->N * sizeof(float), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) rc = fail(c, ABCOCT_ERR_CUDA, "debug tap: %s", cudaGetErrorString(e));
    c->launches += nl + 1;
  }
  cleanup();
  return rc;
}

void* abcoct_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return p;
}
void abcoct_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

int abcoct_get_info(const abcoct_ctx* c, abcoct_info* o) {
  if (!c || !o) return ABCOCT_ERR_INVALID;
  memset(o, 0, sizeof *o);
  o->opw = c->opw;
  o->oph = c->oph;
  o->M = c->M;
  o->N = c->N;
  o->D = c->D;
  o->averages = c->A;
  if (c->generic) {  // run-time radix list of the shared-memory Stockham transform (first three passes reported)
    o->fft_threads = 256;
    for (size_t i = 0; i < 3; ++i) o->fft_radix[i] = i < c->radN.size() ? c->radN[i] : 1;
  } else {
    o->fft_threads = c->wplan ? 32 : c->plan->d.T;
    o->fft_radix[0] = c->wplan ? c->wplan->R : c->plan->d.R0;
    o->fft_radix[1] = c->wplan ? 1 : c->plan->d.R1;
    o->fft_radix[2] = c->wplan ? 32 : c->plan->d.RL;
  }
  o->groups_per_cta = c->G;
  o->ctas_per_sm = 1;
  o->smem_bytes = c->smem;
  o->regs_per_thread = c->regs;
  o->ngpu = (uint32_t)c->gpus.size();
  o->sm_count = c->gpus.empty() ? 0 : c->gpus[0].sm_count;
  o->kernel_launches = c->launches;
  o->last_recon_ms = c->last_recon_ms;
  o->last_norm_ms = c->last_norm_ms;
  o->kernel_kind = c->generic ? 3u : c->wplan ? (c->wplan->resident ? 2u : 1u) : 0u;
  o->slots_per_warp = c->wplan ? (uint32_t)c->wplan->slots : 0u;
  return ABCOCT_OK;
}

}  // extern "C"
