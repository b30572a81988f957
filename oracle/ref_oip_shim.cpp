// ref_oip_shim.cpp -- runs the REFERENCE'S OWN CODE (aux_separator.h, imageop.h, stitcher.h, included
// from /root/reference at build time, never copied) behind a C ABI, with stand-ins for the third-party
// headers it needs (oracle/shim/).  Used only to validate the oracle restatement and to generate golden
// fixtures in this container (tests/golden/make_golden_ref.py).  TEST INFRASTRUCTURE ONLY.
//
//   ref_auxsep        AuxSeparator(file).Separate()            ref aux_separator.h:193-245   (all of stage 1)
//   ref_inplace_rrc   IMO::InplaceRRC                          ref imageop.h:129-138
//   ref_prestitch     Stitcher::PreStitch + SectionaryRemap    ref stitcher.h:83-139, imageop.h:230-275
//                     (cv::remap = the oracle's restatement, itself pinned against cv2)
//   ref_stitch_raw    IMO::StitchBigRaw (RAW out)              ref imageop.h:277-363
//   ref_load_rrc_csv  IMO::LoadRRCParamFile                    ref imageop.h:140-192
//   ref_band_align    PreProcessor::LoadMSS + DoRRC4MSS + DoInterBandAlignment (both overloads)
//                                                              ref preproc.h:56-80, :202-222, :351-468
//                     (cv::remap = the oracle's restatement; cv::merge restated in the stub; cv::imwrite dumps the
//                     matrix bytes; NumCpp / GDAL stand-ins only have to parse)
#include <arpa/inet.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/types.h>
#include <unistd.h>

#include <cassert>
#include <climits>
#include <deque>
#include <filesystem>
#include <future>
#include <iostream>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>

#include "shim/libimsux/imsux.hxx"
#include "shim/libimsux/logger.h"

#include "shim/ref_cv_stub.hpp"

#define private public   // Stitcher keeps mDeltaX/mDeltaY private, PreProcessor its polynomial coefficients; the estimators
#define protected public // that fill them (phase correlation, N1/N2) are floating point and pinned by tolerance elsewhere
#include "aux_separator.h"
#include "stitcher.h"
#include "preproc.h"
#undef protected
#undef private

using namespace OIP;

static int in_dir(const char *dir, const std::function<void()> &fn)
{
    char old[4096];
    if (!getcwd(old, sizeof old)) return -100;
    if (chdir(dir)) return -101;
    int rc = 0;
    try { fn(); } catch (const std::exception &e) { fprintf(stderr, "ref shim: %s\n", e.what()); rc = -2; } catch (...) { rc = -1; }
    if (chdir(old)) return -102;
    return rc;
}

extern "C" int ref_auxsep(const char *aos_or_imdt_path, const char *workdir)
{
    return in_dir(workdir, [&] { AuxSeparator as(aos_or_imdt_path, 0); as.Separate(nullptr); });
}

extern "C" void ref_inplace_rrc(uint16_t *buf, int w, int h, const double *kb)
{
    IMO::InplaceRRC(buf, w, h, reinterpret_cast<const RRCParam *>(kb));
}

extern "C" int ref_prestitch(const char *pan1, const char *pan2, double dx, double dy, const char *workdir)
{
    return in_dir(workdir, [&] {
        Stitcher stt(pan1, pan2, "", "", 1, 1, STT_DEF_OVERLAPPX);
        stt.mDeltaX = dx; stt.mDeltaY = dy;   // what CalcSttParameters would have produced
        stt.PreStitch();                      // reads mRrcFilePAN2 (= pan2), writes <stem>.PRESTT.RAW into cwd
    });
}

extern "C" int ref_stitch_raw(const char *left, const char *right, const char *out, int fold_cols, const char *workdir)
{
    return in_dir(workdir, [&] { Stitcher::Stitch(left, right, out, fold_cols / 2); }); // ref main.cpp:189
}

extern "C" int ref_pixels_per_line(void) { return PIXELS_PER_LINE; }

// IMO::LoadRRCParamFile (ref imageop.h:140-192): kb receives `expected` {k,b} pairs; returns 0, or -2 when the
// reference throws (message on stderr)
extern "C" int ref_load_rrc_csv(const char *path, int expected, double *kb)
{
    try {
        RRCParam *p = IMO::LoadRRCParamFile(path, expected);
        memcpy(kb, p, sizeof(RRCParam) * (size_t)expected);
        delete[] p;
        return 0;
    } catch (const std::exception &e) { fprintf(stderr, "ref shim: %s\n", e.what()); return -2; } catch (...) { return -1; }
}

// the band-alignment leg of the default action (ref main.cpp:296-316 order: LoadMSS, DoRRC4MSS, [coefficients],
// DoInterBandAlignment) with the polynomial coefficients handed in instead of estimated.  pan_path only has to exist with
// 4x the MSS size (ref preproc.h:563-566).  rrc[b] == "" with do_rrc == 0 skips DoRRC4MSS (--no-rrc4mss).
// The aligned raster lands in <workdir>/<mss stem>.ALIGNED.TIFF as raw rows x 3072 x 4 u16 (imwrite stand-in).
extern "C" int ref_band_align(const char *pan_path, const char *mss_path, const char *const rrc[4], int do_rrc, const double cX[8],
                              const double cY[12], int lines_per_section, int line_offset, int overlap, int keep_leading,
                              const char *workdir)
{
    return in_dir(workdir, [&] {
        std::string rrcs[MSS_BANDS];
        for (int b = 0; b < MSS_BANDS; ++b) rrcs[b] = rrc && rrc[b] ? rrc[b] : "";
        PreProcessor pp(pan_path, mss_path, "", rrcs);
        pp.LoadMSS();
        if (do_rrc) pp.DoRRC4MSS();
        for (int b = 0; b < MSS_BANDS; ++b) {
            for (int k = 0; k < 2; ++k) pp.mDeltaXcoeffs[b][k] = cX[2 * b + k];
            for (int k = 0; k < 3; ++k) pp.mDeltaYcoeffs[b][k] = cY[3 * b + k];
        }
        pp.DoInterBandAlignment(lines_per_section, line_offset, overlap, keep_leading != 0, true);
    });
}
