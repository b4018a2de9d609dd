// Host-visible interface of the CUDA translation units (plan registry + launchers).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <vector>

#include "plan.h"
#include "recon_kernel.cuh"
#include "wrow_kernel.cuh"
#include "wres_kernel.cuh"

namespace abcoct {

// One compiled FFT plan: how to build its table blob and how to launch it.
struct PlanEntry {
  PlanDesc d;
  int (*groups)(bool has_sub);  // thread groups (packed A-scan pairs in flight) per CTA, fixed at compile time
  // bytes of dynamic shared memory for G groups
  int (*smem_bytes)(int W, bool has_sub, int G);
  int (*table_bytes)(int W);
  // pack idx (N entries, already sentinel-remapped, values in [1, M]), weights (N), window (W) and the
  // inter-pass twiddles into the blob the kernel copies to shared memory
  void (*build_blob)(int W, const int* idx, const float* wq, const float* win, std::vector<unsigned char>& blob);
  // in_f32: the frames are pre-processed f32 rows from the general path (no calibration, no prefetch); picks the averages == 1 variant itself
  cudaError_t (*launch)(const ReconArgs& a, bool has_sub, bool in_f32, int grid, cudaStream_t st);
  cudaError_t (*attrs)(bool has_sub, bool a1, bool in_f32, int smem, int* regs);  // opt in to large smem, report registers/thread
};

// One compiled plan of the warp-per-A-scan kernel (wrow_kernel.cuh).
struct WPlanEntry {
  int N, R, nw, lm, wmax, smem_bytes;  // lm: WPlan::LM, how pixel / gain rows reach the registers
  // resident-row kernel (wres_kernel.cuh): dB rows stay in shared memory (`slots` per warp, `teams` of 4 warps per CTA), no scratch
  int resident = 0, slots = 0, teams = 0;
  void (*build_blob)(const WrowTablesHost& t, std::vector<unsigned char>& blob);
  void (*permute_cal_row)(const float* in, int W, float* out);  // one calibration row into the kernel's layout (wmax floats)
  cudaError_t (*launch)(const ReconArgs& a, bool has_sub, int grid, cudaStream_t st);  // picks A1 / FULLD from the arguments
  cudaError_t (*attrs)(bool has_sub, bool a1, bool fulld, int* regs);
};
const WPlanEntry* find_wplan(int N, int nw, int lm);  // nw == 0 / lm < 0: the default plan of that length
const WPlanEntry* find_rplan(int N, int nw);         // resident-row plan of that length (nw == 0: the default), or nullptr

const PlanEntry* find_plan(int N);
int list_plans(int* out, int cap);

cudaError_t launch_sched_init(int* sched, int nB, cudaStream_t st);

// ---- general pre-processing path (prep_kernels.cu)
struct PrepArgsHost {
  const void* binned;  // integer pixels after median + binning: frames of oph rows of opw pixels
  size_t row_stride, frame_stride;  // of `binned`, in pixels
  int bpp, opw, oph, nframes, movavgn;
  float px_scale;      // data_y = pixel * px_scale (1, or 1 / 765 for the webcam channel sum)
  const float* yd;     // nullable (DARK variant)
  int rowwise, global_norm;
  float* frame_minmax;  // [nframes][2], needed when global_norm
  const float *yb, *yp, *win;
  int m, M, bandpass;
  int nradW, nradM;
  int radW[12], radM[12];
  const float2 *twW, *twM;  // exp(-2 pi i k / opw), exp(+2 pi i k / M); needed when m > 1
  float* out;               // [nframes][oph][M]
  const float* pre;         // nullable: rows already apodised in f64 (launch_rowprep64); only the Fourier upsample runs
};
// the same stages in f64, one row at a time (normalised-calibration regime, see prep_kernels.cu)
struct PrepArgs64Host {
  const void* binned;
  size_t row_stride, frame_stride;
  int bpp, opw, oph, nframes, movavgn;
  double px_scale;
  const double *yd, *yb, *yp, *win;
  int rowwise, global_norm;
  long long* frame_minmax;  // [nframes][2]
  double* out64;            // [nframes][oph][opw], or NULL and
  float* out32;             // the apodised rows as floats for the Fourier upsample
};
cudaError_t launch_rowprep64(const PrepArgs64Host& h, cudaStream_t st, int* launched);
cudaError_t launch_median(const void* in, void* out, int bpp, int k, int w, int h, size_t row_stride_elems, size_t frame_stride_elems,
                          int nframes, cudaStream_t st);
// BscanFFTwebcam.cpp:1021-1037: interleaved 8-bit BGR frames (strides in bytes) -> dense 16-bit frames of channel sums
cudaError_t launch_bgr_sum(const void* in, uint16_t* out, int w, int h, size_t row_stride_bytes, size_t frame_stride_bytes, int nframes,
                           cudaStream_t st);
cudaError_t launch_bin(const void* in, void* out, int bpp, int opw, int oph, int bx, int by, size_t row_stride_elems,
                       size_t frame_stride_elems, int nframes, cudaStream_t st);
// calibration captures: sum over the frames of the binned pixels as f64 (+ smoothmovavg), BscanFFT.cpp:1041-1046
cudaError_t launch_cal_accum(const void* px, int bpp, size_t row_stride_elems, size_t frame_stride_elems, int nframes, int opw, int oph,
                             int movavgn, double px_scale, double* acc, cudaStream_t st);
// the once-per-capture tail on the accumulated frame x (rows x cols doubles, in place): row-wise / global min-max normalise to
// [lo, 1] (BscanFFT.cpp:1050-1055, 1092-1096), else x *= inv_n (:1057), then lpfilter (BscanDark.cpp:119-167, 1070-1074)
struct CalTailHost {
  double* x;
  int rows, cols;
  int rowwise, global_norm, lowpass;
  double lo, inv_n;
  long long* mm;          // two words of scratch (global_norm)
  const double *cs, *sn;  // cos / sin(2 pi k / cols), k < cols (lowpass)
};
cudaError_t launch_cal_tail(const CalTailHost& h, cudaStream_t st, int* launched);
cudaError_t launch_rowprep(const PrepArgsHost& h, cudaStream_t st, int* launched);
size_t rowprep_smem_bytes(int opw, int M, int m, int movavgn);
cudaError_t launch_lerp_rows64(const float* rows, const double* rows64, const int* idx, const double* wq, float* ylin, int M, int N, int oph,
                               cudaStream_t st);
cudaError_t launch_lerp_rows(const float* rows, const int* idx, const float* wq, float* ylin, int M, int N, int oph, cudaStream_t st);

// ---- any transform length N = 2^a 3^b 5^c, any row width, D up to N (prep_kernels.cu): gather-lerp + Stockham DFT + magnitude +
// accumulate + dB into the scratch, then normalise + transpose; runs on the rows prepared by launch_rowprep
struct GenericHost {
  const float* rows;
  const double* rows64;  // single_row without Fourier upsample: the prepared rows as doubles
  const double* wq64;    // single_row: lerp weights in f64
  int M, N, D, Dp, oph, A, nB;
  const int* idx;
  const float* wq;
  int nrad, rad[12];
  const float2* tw;  // exp(+2 pi i k / N)
  float* scratch;
  int *minv, *maxv;
  float* dc01;
  uint8_t* out8;
  float* outdb;
  float out_scale, db_scale_ln, thr, clamp_db;
  int clamp55;
  int single_row;  // one A-scan per transform instead of two packed ones (normalised-calibration regime)
};
size_t generic_smem_bytes(int N, int D);
cudaError_t launch_generic(const GenericHost& h, cudaStream_t st, int* launched);

// ---- consumers of a finished B-scan (post_kernels.cu)
cudaError_t post_init_device();  // once per device: the JET table
// linear bscan (BscanFFT.cpp:1220-1222) from the dB image [nB][px] and the unmasked DC rows dc01 [nB][oph][2]
cudaError_t launch_lin_from_db(const float* db, const float* dc01, float* lin, int oph, size_t px, int nB, float inv_db_scale,
                               int sm_count, cudaStream_t st, int* launched);
// 'Bscan subtracted' display of the J0 lock-in (BscanFFT.cpp:1225-1231, 1257-1267).  Source: lin [nB][px] when the linear image
// exists, else (lin == nullptr) the dB image + dc01 and the linear value is derived on the fly; jscan [px], mm [2 nB] ints
cudaError_t launch_jsub(const float* lin, const float* db, const float* dc01, const float* jscan, int* mm, uint8_t* out, int oph,
                        size_t px, int nB, float db_scale, float inv_db_scale, float thr, int sm_count, cudaStream_t st, int* launched);
// x += v over two arrays (either may be NULL): the dB offset of BscanFFTspinjnt's multiplyfactor (BscanFFTspinjnt.cpp:1860)
cudaError_t launch_add_const(float* a, size_t na, float* b, size_t nb, float v, int sm_count, cudaStream_t st, int* launched);
cudaError_t launch_jet(const uint8_t* in, uint8_t* out, size_t n, int sm_count, cudaStream_t st);  // applyColorMap(., COLORMAP_JET)
}  // namespace abcoct
