/*
 * abcoct.h - C ABI of the B200-native ABC-OCT B-scan reconstruction path.
 *
 * The reference (hn-88/FDOCT) has NO function / plugin / FFI boundary on this path: the
 * reconstruction is an inline block of main() (BscanFFT.cpp:953-958, 987-991, 1125-1255;
 * identical copies in BscanFFTspin.cpp:1095-1399, BscanFFTspinj.cpp:1635-1982,
 * BscanFFTpeak.cpp:1557-1873, BscanFFTwebcam.cpp:1044-1349; dark-frame variant
 * BscanDark.cpp:946-951, 1268-1393).  The "signature" of that block is the set of locals it
 * reads and writes; every entry point below cites the reference statements it replaces, so a
 * host app can delete those lines and make one call instead.
 *
 * Conventions: plain C, no C++ exceptions cross the ABI, `int` status with 0 == success
 * (like QHYCCD_SUCCESS, BscanFFT.cpp:730) and negative == error; abcoct_last_error() gives the
 * text.  A context is not thread-safe (the reference block is single-threaded).  There is no
 * CPU fallback: every compute entry point fails with ABCOCT_ERR_CUDA when no sm_100 device
 * is present.
 */
#ifndef ABCOCT_H
#define ABCOCT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ABCOCT_VERSION 100

enum {
  ABCOCT_OK = 0,
  ABCOCT_ERR_INVALID = -1,     /* bad argument / parameter combination the reference cannot run either */
  ABCOCT_ERR_UNSUPPORTED = -2, /* valid in the reference but not built yet (listed in DESIGN.md)        */
  ABCOCT_ERR_CUDA = -3,        /* CUDA runtime error, no device, wrong architecture                     */
  ABCOCT_ERR_STATE = -4,       /* e.g. process called before a background was set                       */
  ABCOCT_ERR_IO = -5           /* ini file could not be opened                                          */
};

/* Which positional .ini layout to parse (the parser is BscanFFT.cpp:395-484). */
enum {
  ABCOCT_INI_BSCANFFT = 0, /* BscanFFT.cpp:424-476, BscanFFTspin.cpp:439-491                      */
  ABCOCT_INI_SPINJ = 1,    /* BscanFFTspinj.cpp:864-918   (+ offlinetoolpath)                     */
  ABCOCT_INI_SPINJNT = 2,  /* BscanFFTspinjnt.cpp:769-829 (binvaluex/y, bscanbinx/y)              */
  ABCOCT_INI_DARK = 3,     /* BscanDark.cpp:434-486       (no offsets, + bandpass/lowpass)        */
  ABCOCT_INI_PEAK = 4,     /* BscanFFTpeak.cpp:1030-1080  (no offsets, + peakholdnumframes)       */
  ABCOCT_INI_WEBCAM = 5,   /* BscanFFTwebcam.cpp:458-508  (no offsets, + channelnum)              */
  ABCOCT_INI_SIM = 6       /* BscanFFTsim.cpp:316-360     (no offsets, stops after multiplier)    */
};

/* Mirrors the locals the block reads (BscanFFT.cpp:357-387 defaults, filled by the .ini parser). */
typedef struct abcoct_params {
  uint32_t w, h;             /* raw frame width / height in pixels (BscanFFT.cpp:430-432)                  */
  uint32_t bpp;              /* 8 or 16 (BscanFFT.cpp:428); 16-bit frames may hold 12-bit data              */
  uint32_t binx, biny;       /* binvalue (BscanFFT.cpp:446) or binvaluex/y (BscanFFTspinjnt.cpp:1553)       */
  uint32_t averages;         /* averagestoggle: frames averaged per B-scan (BscanFFT.cpp:450, 481, 1193)   */
  uint32_t numfftpoints;     /* N (BscanFFT.cpp:452)                                                       */
  uint32_t numdisplaypoints; /* D (BscanFFT.cpp:464)                                                       */
  double lambdamin, lambdamax; /* BscanFFT.cpp:479-480                                                     */
  int32_t mediann;           /* BscanFFT.cpp:470, 953                                                      */
  int32_t movavgn;           /* BscanFFT.cpp:462, 990                                                      */
  uint32_t fft_multiplier;   /* increasefftpointsmultiplier (BscanFFT.cpp:472, 1146)                       */
  uint8_t rowwisenormalize;  /* BscanFFT.cpp:474, 1126                                                     */
  uint8_t donotnormalize;    /* BscanFFT.cpp:476, 1128                                                     */
  uint8_t variant;           /* 0 = FFT (BscanFFT.cpp), 1 = DARK (BscanDark.cpp:1269 subtracts data_yd)    */
  uint8_t weight_mode;       /* 0 = reference quirk fractionalk[nearestkindex[q]] (BscanFFT.cpp:1170),
                                1 = corrected fractionalk[q]                                               */
  double bscanthreshold;     /* BscanFFT.cpp:385, 1247 (default -30.0)                                     */
  uint8_t clampupper;        /* BscanFFT.cpp:374, 1248                                                     */
  uint8_t bandpassfilter;    /* BscanDark.cpp:218-236 (inside the Fourier upsample only)                   */
  uint8_t lowpassfilter;     /* BscanDark.cpp:1070-1074: lpfilter (:119-167) on the captured dark / reference / sample frames */
  uint8_t output_rebin;      /* 1 = the caller is BscanFFTspinjnt, whose block re-bins the LINEAR B-scan before the log whenever
                                any of binx / biny / bscanbinx / bscanbiny exceeds 1 (resize INTER_AREA down by bscanbinx/y, times
                                multiplyfactor, resize INTER_CUBIC up by bscanbinx * binvaluey and bscanbiny,
                                BscanFFTspinjnt.cpp:835, 1856-1862).  Set by the ABCOCT_INI_SPINJNT parser.  Built for the shape the
                                shipped ini gives (binx > 1, biny = bscanbinx = bscanbiny = 1): both resizes are copies there and the
                                block is bscan *= multiplyfactor.  With a real resampling step the reference's own output is NaN
                                wherever the bicubic overshoot is negative; abcoct_create answers ABCOCT_ERR_UNSUPPORTED.         */
  uint8_t bscanbinx, bscanbiny; /* BscanFFTspinjnt.cpp:795-797; only looked at when output_rebin is set                       */
  uint8_t channelnum;        /* BscanFFTwebcam.cpp:412, 1016-1037: 0..2 = the caller passes the selected 8-bit plane (bpp = 8);
                                >= 3: frames are interleaved 8-bit BGR (3 bytes per pixel, what cap.read gives) and the library
                                sums the channels, scaled by 0.00130718954 like the reference's CV_64F mraw.  Needs bpp = 8,
                                mediann = 0 (cv::medianBlur throws on CV_64F in the reference) and no binning (not built)      */
  uint8_t reserved[1];
  double clamp_db;           /* 50.0 (BscanFFT.cpp:1252), 30.0 in BscanFFTspinjnt.cpp:1886                 */
} abcoct_params;

typedef struct abcoct_ctx abcoct_ctx; /* opaque; one per caller thread; may span 1..8 GPUs */

/* Fill *out with the reference's compile-time defaults (BscanFFT.cpp:357-390). */
void abcoct_params_default(abcoct_params* out);

/* Positional whitespace-token .ini parser: skips three tokens, then alternating comment / value
 * (BscanFFT.cpp:417-482).  A missing file is ABCOCT_ERR_IO and leaves the defaults in *out, like
 * "Unable to open ini file, using defaults." (BscanFFT.cpp:484). */
int abcoct_params_from_ini(const char* path, int ini_flavour, abcoct_params* out);

/* Replaces the one-time precompute BscanFFT.cpp:545-546, 615-698, 932-944 and owns the state the
 * block keeps between frames (data_yb/yp/yd, bscantransposed, indextemp).  gpu_ids == NULL means
 * device 0..ngpu-1.  Rejects what is undefined behaviour in the reference: numfftpoints <
 * fft_multiplier * (w / binx) (BscanFFT.cpp:1170 reads past fractionalk), numdisplaypoints >
 * numfftpoints (colRange throws, :1193) or < 6 (rows 4 and 5 are addressed at :1239, :1252).
 * Transform lengths with a fused plan (abcoct_list_plans), rows that are a multiple of 8 samples and
 * numdisplaypoints <= numfftpoints / 2 run in the fused kernels; every other numfftpoints = 2^a 3^b 5^c,
 * row width and numdisplaypoints <= numfftpoints runs on the generic kernel (abcoct_info.kernel_kind 3),
 * same results, a fraction of the throughput.  Other lengths: ABCOCT_ERR_UNSUPPORTED. */
int abcoct_create(const abcoct_params* params, const int* gpu_ids, int ngpu, abcoct_ctx** out);
void abcoct_destroy(abcoct_ctx* ctx);
const char* abcoct_last_error(const abcoct_ctx* ctx); /* ctx may be NULL: error of the last failed create */

/* Run-time state the reference's key handler mutates between frames (no re-creation needed):
 *   bscanthreshold   keys '[' / ']' (BscanFFT.cpp:1759-1775)
 *   clampupper       BscanFFT.cpp:374, 1248
 *   averages         key 'a' toggles averagestoggle between 1 and `averages` (BscanFFT.cpp:1874-1877); >= 1 */
int abcoct_set_threshold(abcoct_ctx* ctx, double bscanthreshold);
int abcoct_set_clampupper(abcoct_ctx* ctx, int on);
int abcoct_set_averages(abcoct_ctx* ctx, uint32_t averages);

/* Calibration state, oph x opw doubles with a row stride of `ld` elements.
 *   background = data_yb (BscanFFT.cpp:1050-1057; BscanDark.cpp:996), required before processing;
 *   pishift    = data_yp (BscanFFT.cpp:1081), NULL -> zeros (BscanFFT.cpp:563);
 *   dark       = data_yd (BscanDark.cpp:1045-1067), used by the DARK variant only.               */
int abcoct_set_background(abcoct_ctx* ctx, const double* yb, size_t ld);
int abcoct_set_pishift(abcoct_ctx* ctx, const double* yp, size_t ld);
int abcoct_set_dark(abcoct_ctx* ctx, const double* yd, size_t ld);
/* The capture done on keys b / p / o / r / t (BscanFFT.cpp:1041-1062, 1081; BscanDark.cpp:1045-1225): data_y of
 * `nframes` raw frames (after median, binning, convertTo, smoothmovavg) is accumulated, then normalised to [0.0001, 1]
 * row-wise (rowwisenormalize) and / or globally (!donotnormalize), else divided by nframes, with the reference's
 * `if (rowwise) ...; if (!donotnormalize) ...; else /n` structure; the DARK captures are low-pass filtered when
 * `lowpassfilter` is set.  The pi-shifted frame (which = 1) is a copy of ONE frame, normalised to [0, 1] row-wise / globally
 * under the same two switches (BscanFFT.cpp:1081, 1092-1096).  Runs on the host (once per key press).
 *   which: 0 background data_yb, 1 pishift data_yp, 2 dark data_yd, 3 reference arm data_yr, 4 sample arm data_ys. */
int abcoct_set_calibration_from_frames(abcoct_ctx* ctx, int which, const void* frames, size_t nframes,
                                       size_t stride_bytes);

/* Read a calibration frame back (oph x opw doubles, row stride `ld`, 0 = dense), e.g. to save data_yb as spectrum.ocv
 * like BscanFFTspinj.cpp:1788-1789.  which as in abcoct_set_calibration_from_frames; ABCOCT_ERR_STATE if it was never set. */
int abcoct_get_calibration(const abcoct_ctx* ctx, int which, double* out, size_t ld);

/* BscanDark's key 'b': data_yb = (data_yr - data_yd) + (data_ys - data_yd) (BscanDark.cpp:996) from the captures
 * which = 3, 4 and the current dark frame. */
int abcoct_compose_dark_background(abcoct_ctx* ctx);

/* Host-only (no GPU needed): the one-time precompute of BscanFFT.cpp:615-698 and :936-944 for `params`.
 * nearestkindex / fractionalk have numfftpoints entries, barthannwin has w / binx; any may be NULL. */
int abcoct_build_tables(const abcoct_params* params, int32_t* nearestkindex, double* fractionalk, double* barthannwin);

/* The bit-exact tables of BscanFFT.cpp:673-698 (N entries each) and the window of :936-944 (opw). */
int abcoct_get_tables(const abcoct_ctx* ctx, int32_t* nearestkindex, double* fractionalk);
int abcoct_get_window(const abcoct_ctx* ctx, double* barthannwin);

/* THE HOT PATH.  Replaces BscanFFT.cpp:953-958, 987-991, 1125-1255 for `nframes` consecutive frames.
 *   frames       nframes x h x w pixels (uint8 when bpp == 8, else uint16), HOST memory, rows
 *                `stride_bytes` apart (0 -> dense); pinned memory from abcoct_host_alloc is copied
 *                straight to the device, other memory goes through the internal pinned ring.
 *   nframes      must be a multiple of `averages`; nB = nframes / averages B-scans come out.
 *   bscan_u8     nB x D x oph, the display image bscandisp (BscanFFT.cpp:1254-1255).
 *   bscan_db     nullable, nB x D x oph float, bscandb after the DC-row mask (BscanFFT.cpp:1237-1240).
 * Synchronous: inputs may be freed and outputs are valid on return. */
int abcoct_process_bscans(abcoct_ctx* ctx, const void* frames, size_t nframes, size_t stride_bytes,
                          uint8_t* bscan_u8, float* bscan_db);

/* Same computation with every buffer already resident on GPU `gpu_index` (index into gpu_ids).
 * Asynchronous on `cuda_stream` (a cudaStream_t; NULL = the context's own stream, then the call
 * synchronises before returning).  d_frames must be 16-byte aligned with a 16-byte multiple row stride; the output
 * pointers may have any alignment (aligned ones take the vector-store path).  The asynchronous calls of ONE context share
 * its scratch / scheduler buffers: enqueue them on a single stream at a time (or order the streams with events), and do
 * not change the calibration while such work is in flight unless you synchronise first. */
int abcoct_process_bscans_device(abcoct_ctx* ctx, int gpu_index, const void* d_frames, size_t nframes,
                                 size_t stride_bytes, uint8_t* d_bscan_u8, float* d_bscan_db,
                                 void* cuda_stream);

/* ---- consumers of a finished B-scan (the statements that follow the block in the reference's loop) ----------------
 * Every image is nB x D x oph, depth-major like bscan_u8; all but bscan_u8 may be NULL.
 *   bscan_lin   float, the linear `bscan` Mat: mean magnitude + 1e-5 before the log and before the DC-row mask
 *               (BscanFFT.cpp:1220-1222) - what key 'j' copies into jscansave (:1292-1296).
 *   bscan_bgr   3 bytes per pixel, applyColorMap(bscandisp, COLORMAP_JET) (BscanFFT.cpp:1284), OpenCV's own table.
 *   jsub_u8     the 'Bscan subtracted' display of the J0 lock-in (BscanFFT.cpp:1225-1231, 1257-1267):
 *               max(bscan - jscansave, 0) + 0.001 -> ln * 20 / 2.303 -> max(., bscanthreshold) -> min-max normalise -> u8
 *               (no DC-row mask, no clampupper, as in the reference).  Needs abcoct_set_jscan, else ABCOCT_ERR_STATE.
 *   jsub_bgr    applyColorMap(jsub_u8, COLORMAP_JET) (BscanFFT.cpp:1268). */
typedef struct abcoct_outputs {
  uint8_t* bscan_u8;
  float* bscan_db;
  float* bscan_lin;
  uint8_t* bscan_bgr;
  uint8_t* jsub_u8;
  uint8_t* jsub_bgr;
  void* reserved[2]; /* must be NULL */
} abcoct_outputs;

/* Key 'j' (BscanFFT.cpp:1292-1296): keep one linear B-scan (D x oph floats, row stride `ld` elements, 0 = dense; a
 * bscan_lin output of an earlier call) as the lock-in reference.  NULL switches the lock-in off (key 'c', :1298-1303). */
int abcoct_set_jscan(abcoct_ctx* ctx, const float* jscan, size_t ld);

/* abcoct_process_bscans with the extra outputs above (host buffers; same ring, same multi-GPU split). */
int abcoct_process_bscans_ex(abcoct_ctx* ctx, const void* frames, size_t nframes, size_t stride_bytes,
                             const abcoct_outputs* out);
/* abcoct_process_bscans_device with the extra outputs above (device pointers on GPU `gpu_index`). */
int abcoct_process_bscans_device_ex(abcoct_ctx* ctx, int gpu_index, const void* d_frames, size_t nframes,
                                    size_t stride_bytes, const abcoct_outputs* d_out, void* cuda_stream);

/* Kernel timing of the device entry point (bench.py's roofline): every chunk enqueued by
 * abcoct_process_bscans_device after abcoct_timing_reset is bracketed by CUDA events ON THE LAUNCHING STREAM
 * (up to 512 chunks); abcoct_timing_read waits for them and returns the summed duration of the fused kernel
 * (reconstruction and display normalisation run in ONE launch, so norm_ms is the ~0 gap after it) and how many
 * chunks were timed. */
int abcoct_timing_reset(abcoct_ctx* ctx);
int abcoct_timing_read(abcoct_ctx* ctx, int gpu_index, uint32_t* nchunks, double* recon_ms, double* norm_ms);

/* Stage-level debug tap for the parity tests: data_ylin (BscanFFT.cpp:1151-1177), i.e. one frame after every
 * pre-processing stage and the lambda->k gather-lerp, oph x numfftpoints floats.  It runs the general pre-processing
 * kernels (prep_kernels.cu) whatever the configuration; the fused kernel never materialises this intermediate. */
int abcoct_debug_linearised(abcoct_ctx* ctx, const void* frame, size_t stride_bytes, float* ylin /* oph x N */);

/* Pinned host memory for zero-staging ingest (the caller's ring buffer). */
void* abcoct_host_alloc(size_t bytes);
void abcoct_host_free(void* p);

/* Introspection used by bench.py / tests: derived sizes and launch accounting. */
typedef struct abcoct_info {
  uint32_t opw, oph, M, N, D, averages;
  uint32_t fft_threads, fft_radix[3], groups_per_cta, ctas_per_sm, smem_bytes, regs_per_thread;
  uint32_t ngpu, sm_count;
  uint64_t kernel_launches;        /* kernels launched by this ctx so far                     */
  double last_recon_ms, last_norm_ms; /* mean per-chunk kernel times of the last abcoct_timing_read */
  uint32_t kernel_kind;            /* fused kernel in use: 0 thread group per row pair, 1 warp per A-scan + dB scratch, 2 warp per A-scan, dB rows resident in tensor memory, 3 generic (any N = 2^a 3^b 5^c, any row width, D <= N; run-time radices) */
  uint32_t slots_per_warp;         /* kernel_kind 2: dB row slots per warp                    */
} abcoct_info;
int abcoct_get_info(const abcoct_ctx* ctx, abcoct_info* out);

#ifdef __cplusplus
}
#endif
#endif /* ABCOCT_H */
