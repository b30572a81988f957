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

#define private public   // Stitcher keeps mDeltaX/mDeltaY private; CalcSttParameters (phase correlation) is out of scope
#include "aux_separator.h"
#include "stitcher.h"
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
