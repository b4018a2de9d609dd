// Harness around the reference's OWN processing block, compiled verbatim.  TEST INFRASTRUCTURE ONLY.
//
// oracle/build_ref.py cuts these line ranges out of /root/reference/BscanFFT.cpp into oracle/_ref/frag_*.inc at build time (checked
// against anchor strings, deleted again after the compile; the repository holds none of the reference's text):
//   frag_normalizerows  BscanFFT.cpp:88-97      normalizerows()
//   frag_helpers        BscanFFT.cpp:173-305    makeonlypositive(), zeropadrowwise(), smoothmovavg()
//   frag_tables         BscanFFT.cpp:615-698    lambda -> k tables (nearestkindex, fractionalk)
//   frag_window         BscanFFT.cpp:936-944    Bartlett-Hann window
//   frag_ingest1        BscanFFT.cpp:953-958    medianBlur + INTER_AREA binning
//   frag_ingest2        BscanFFT.cpp:987-991    convertTo(CV_64F) + smoothmovavg
//   frag_keys           BscanFFT.cpp:1000-1099  what keys 'b' (accumulate averagestoggle frames -> data_yb, normalise branches) and 'p'
//                                               (copy of one frame -> data_yp, normalised like data_y) do inside the frame loop
//   frag_block          BscanFFT.cpp:1125-1284  normalise, (y - yp) / yb, mean, window, upsample, gather-lerp, DFT, magnitude,
//                                               accumulate, dB, DC mask, threshold, clamp, min-max, u8, the J0 lock-in display
//                                               (:1225-1231, 1256-1268) and the JET colour images (:1267, 1284; the "^" marker of :1285 is not drawn)  (+ one closing brace)
// and, compiled a second time with -DREF_DARK into abcoct_ref_dark, the same ranges of /root/reference/BscanDark.cpp
// (82-91, 111-314 incl. lpfilter and the band-pass in zeropadrowwise, 614-697, 929-937, 946-951, 980-984, 993-1249: keys 'b'
// (data_yb composed from the three captures), 'o' / 'r' / 't' (dark / reference-arm / sample-arm captures, lpfilter) and 'p',
// 1268-1393: the block with the dark-frame subtraction `data_y = data_y - data_yd`).
// Everything in THIS file is the scaffolding main() has around those ranges: the declarations (same names and types as
// BscanFFT.cpp:349-613), the frame loop, and the state the key handler would set (data_yb / data_yp, BscanFFT.cpp:1027-1033, 1081).
// The OpenCV calls inside the fragments run in OpenCV itself, through cv2 (oracle/cvshim).
#include "opencv2/opencv.hpp"

using namespace cv;  // BscanFFT.cpp:86

#include "_ref/frag_normalizerows.inc"
#include "_ref/frag_helpers.inc"

// file output of the lock-in branch and of saveinterferograms (BscanFFT.cpp:1023, 1276-1278): nothing is written here
static void savematasdata(std::ofstream&, char*, Mat) {}
static void savematasimage(char*, char*, char*, Mat) {}

namespace py = pybind11;

static py::dict run_block(py::dict prm, py::array frames, py::object yb_in, py::object yp_in, py::object yd_in, py::object jscan_in) {
  // ---- declarations, BscanFFT.cpp:357-613 (names and types as there; camera / file / GUI state left out)
  unsigned int w = prm["w"].cast<unsigned>(), h = prm["h"].cast<unsigned>(), opw, oph;
  uint indextemp;
  uint averagestoggle = prm["averages"].cast<unsigned>();
  int binvalue = prm["binvalue"].cast<int>();
  int numfftpoints = prm["numfftpoints"].cast<int>();
  int numdisplaypoints = prm["numdisplaypoints"].cast<int>();
  bool saveframes = 0;
  int movavgn = prm["movavgn"].cast<int>();
  bool clampupper = prm["clampupper"].cast<bool>();
  bool jlockin = 0;
#ifdef REF_DARK
  bool jthresholding = 0;                                 // BscanDark.cpp
  bool bandpassfilter = prm["bandpassfilter"].cast<bool>();  // BscanDark.cpp:396
  Mat jmask, jmaskt;
#endif
  double lambdamin = prm["lambdamin"].cast<double>(), lambdamax = prm["lambdamax"].cast<double>();
  int mediann = prm["mediann"].cast<int>();
  uint increasefftpointsmultiplier = prm["fft_multiplier"].cast<unsigned>();
  double bscanthreshold = prm["bscanthreshold"].cast<double>();
  bool rowwisenormalize = prm["rowwisenormalize"].cast<bool>();
  bool donotnormalize = prm["donotnormalize"].cast<bool>();
  bool zeroisactive = 1;
  Mat m, opm, bscan, bscanlog, bscandb, bscandisp, bscantemp, bscantransposed;
  Mat tempmat;
  Mat mraw;
  Mat statusimg = Mat::zeros(cv::Size(600, 300), CV_64F);
  Mat secrowofstatusimgRHS = statusimg(Rect(300, 50, 300, 50));
  char textbuffer[80];
  opw = w / binvalue;  // :545
  oph = h / binvalue;  // :546
  Mat data_y(oph, opw, CV_64F);
  Mat data_ylin(oph, numfftpoints, CV_64F);
  Mat data_yb(oph, opw, CV_64F);
  Mat data_yp(oph, opw, CV_64F);
  Mat barthannwin(1, opw, CV_64F);
  data_yb = Mat::zeros(Size(opw, oph), CV_64F);  // :562
  data_yp = Mat::zeros(Size(opw, oph), CV_64F);  // :563
#ifdef REF_DARK
  Mat data_yd(oph, opw, CV_64F);
  data_yd = Mat::zeros(Size(opw, oph), CV_64F);
  if (!yd_in.is_none()) Mat(yd_in).copyTo(data_yd);  // key 'o' (BscanDark.cpp)
#endif
  Mat bscansave0[100];
  Mat bscansave1[100];
  Mat jscansave;
  // key handler state (BscanFFT.cpp:365-377, 555-566); prm["keys"][i] = the key pressed before frame i: 0 none, 1 'b', 2 'p',
  // and in BscanDark 3 'o' (dark), 4 'r' (reference arm), 5 't' (sample arm)
  bool saveinterferograms = 0, manualaveraging = 0;
  bool bkeypressed = 0, pkeypressed = 0;
  unsigned int indexi = 0;
  Mat baccum = Mat::zeros(Size(opw, oph), CV_64F);  // :564
  uint baccumcount = 0;                             // :565
  Mat interferogramsave0[100], interferogramsave1[100], interferogrambsave0[100], interferogrambsave1[100];
  Mat secrowofstatusimg = statusimg(Rect(0, 50, 600, 50));
  std::vector<int> keys;
  if (prm.contains("keys")) keys = prm["keys"].cast<std::vector<int>>();
#ifdef REF_DARK
  bool rkeypressed = 0, tkeypressed = 0, darkkeypressed = 0;
  bool lowpassfilter = prm.contains("lowpassfilter") && prm["lowpassfilter"].cast<bool>();  // BscanDark.cpp:397
  Mat data_yr = Mat::zeros(Size(opw, oph), CV_64F), data_ys = Mat::zeros(Size(opw, oph), CV_64F);
  char filename[20], pathname[140] = "", dirname[80] = "";
#endif
#ifndef REF_DARK
  Mat bscansublog, bscandispmanual, cmagI, cmagImanual, manualaccum;
  uint manualindexi = 0;
  char filename[20], filenamec[20], pathname[140] = "", dirname[80] = "";
  std::ofstream outfile;
  if (!jscan_in.is_none()) {  // key 'j' (BscanFFT.cpp:1292-1297): jscansave = a finished linear bscan, lock-in on
    Mat(jscan_in).copyTo(jscansave);
    jlockin = 1;
  }
#endif
  Mat positivediff;
  Mat magI;
  Scalar meanval;
  Mat lambdas, k, klinear;
  Mat diffk, slopes, fractionalk, nearestkindex;
  double kmin, kmax;
  double pi = 3.141592653589793;  // :609

#include "_ref/frag_tables.inc"

  indextemp = 0;                                                         // :931
  bscantransposed = Mat::zeros(Size(numdisplaypoints, oph), CV_64F);    // :932
#include "_ref/frag_window.inc"

  // what keys 'b' / 'p' leave behind (:1027-1033, :1081): the caller's calibration frames
  if (!yb_in.is_none()) Mat(yb_in).copyTo(data_yb);
  if (!yp_in.is_none()) Mat(yp_in).copyTo(data_yp);

  py::list disp, db, jdisp, jbgr, bgr;
  py::object last_ylin = py::none();
  const py::ssize_t nframes = frames.shape(0);
  for (py::ssize_t fi = 0; fi < nframes; ++fi) {
    mraw = Mat(py::reinterpret_borrow<py::object>(frames[py::int_(fi)]));  // GetQHYCCDLiveFrame(..., mraw.data), :949
    switch ((size_t)fi < keys.size() ? keys[fi] : 0) {  // waitKey of the previous iteration
      case 1: bkeypressed = 1; break;
      case 2: pkeypressed = 1; break;
#ifdef REF_DARK
      case 3: darkkeypressed = 1; break;
      case 4: rkeypressed = 1; break;
      case 5: tkeypressed = 1; break;
#endif
      default: break;
    }
    {
#include "_ref/frag_ingest1.inc"
#include "_ref/frag_ingest2.inc"
#include "_ref/frag_keys.inc"
#include "_ref/frag_block.inc"
      }  // closes `if (indextemp >= averagestoggle)` (the J0 lock-in display and the key handler follow in the reference)
    }
    if (indextemp == 0) {  // a B-scan was completed by this frame
      disp.append(ops().attr("copy")(bscandisp.arr));
      db.append(ops().attr("copy")(bscandb.arr));
#ifndef REF_DARK
      bgr.append(ops().attr("copy")(cmagI.arr));
      if (jlockin) {
        jdisp.append(ops().attr("copy")(bscandispmanual.arr));
        jbgr.append(ops().attr("copy")(cmagImanual.arr));
      }
#endif
      bscantransposed = Mat::zeros(Size(numdisplaypoints, oph), CV_64F);  // :1482
    }
  }
  py::dict out;
  out["bscandisp"] = disp;
  out["bscandb"] = db;
  out["cmagI"] = bgr;
  out["bscandispmanual"] = jdisp;
  out["cmagImanual"] = jbgr;
  out["data_yb"] = ops().attr("copy")(data_yb.arr);
  out["data_yp"] = ops().attr("copy")(data_yp.arr);
#ifdef REF_DARK
  // (BscanDark never clears bkeypressed: once 'b' was pressed, data_yb is recomposed on every frame, BscanDark.cpp:993-1003)
  out["capture_pending"] = pkeypressed || darkkeypressed || rkeypressed || tkeypressed;
  out["data_yd"] = ops().attr("copy")(data_yd.arr);
  out["data_yr"] = ops().attr("copy")(data_yr.arr);
  out["data_ys"] = ops().attr("copy")(data_ys.arr);
#else
  out["capture_pending"] = bkeypressed || pkeypressed;
#endif
  out["nearestkindex"] = ops().attr("copy")(nearestkindex.arr);
  out["fractionalk"] = ops().attr("copy")(fractionalk.arr);
  out["barthannwin"] = ops().attr("copy")(barthannwin.arr);
  out["data_ylin"] = ops().attr("copy")(data_ylin.arr);
  (void)kmin, (void)kmax, (void)zeroisactive, (void)saveframes, (void)jlockin, (void)textbuffer, (void)clampupper, (void)yd_in, (void)jscan_in;
  return out;
}

#ifndef REF_DARK
// BscanFFTwebcam.cpp:1018-1038, compiled verbatim: the frame cap.read() returned -> mraw (one 8-bit plane for channelnum < 3, else the
// CV_64F sum of the three planes * 0.00130718954).  Declarations as at BscanFFTwebcam.cpp:412, 578-580.
static py::object webcam_mraw(py::object frame_in, int channelnum_in) {
  int channelnum = channelnum_in;
  Mat frame(frame_in), mraw;
  Mat rgbchannels[3];
#include "_ref/frag_webcam.inc"
  return mraw.arr;
}
#endif

#ifndef REF_DARK
// BscanFFTspinjnt.cpp:1856-1862, compiled verbatim: the re-binning of the linear B-scan in front of the log.  multiplyfactor as at :835.
static py::object spinjnt_rebin(py::object bscan_in, int bscanbinx, int bscanbiny, int binvaluex, int binvaluey) {
  int multiplyfactor = bscanbinx * bscanbiny * binvaluex * binvaluey;  // BscanFFTspinjnt.cpp:835
  Mat bscan(ops().attr("copy")(bscan_in)), bscanbinned;
#include "_ref/frag_spinjnt_rebin.inc"
  return bscan.arr;
}
#endif

#ifdef REF_DARK
PYBIND11_MODULE(abcoct_ref_dark, mod) {
  mod.doc() = "the reference's processing block (BscanDark.cpp), compiled verbatim against oracle/cvshim";
  mod.def("lpfilter", [](py::object a) {  // BscanDark.cpp:119-167, applied to the captured calibration frames (:1070-1074)
    Mat mm(ops().attr("copy")(a));
    lpfilter(mm);
    return mm.arr;
  });
#else
PYBIND11_MODULE(abcoct_ref, mod) {
  mod.doc() = "the reference's processing block (BscanFFT.cpp), compiled verbatim against oracle/cvshim";
#endif
#ifndef REF_DARK
  mod.def("webcam_mraw", &webcam_mraw, py::arg("frame"), py::arg("channelnum"));
  mod.def("spinjnt_rebin", &spinjnt_rebin, py::arg("bscan"), py::arg("bscanbinx"), py::arg("bscanbiny"), py::arg("binvaluex"), py::arg("binvaluey"));
#endif
  mod.def("run_block", &run_block, py::arg("params"), py::arg("frames"), py::arg("yb") = py::none(), py::arg("yp") = py::none(),
          py::arg("yd") = py::none(), py::arg("jscan") = py::none());
  mod.def("opencv_version", []() { return ops().attr("opencv_version")().cast<std::string>(); });
}
