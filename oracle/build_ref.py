#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY.  Recipe that compiles the UNMODIFIED reference hot-path sources where
they lie under /root/reference into oracle/_ref/ (git-ignored, travels to the GPU box):

  oracle/_ref/libref_mtrack.so   level A: mtracklib TUs + oracle/ref_shim.cpp (ctypes-callable stages)
  oracle/_ref/ref_rebvo          level B: the whole 3-thread REBVO class + oracle/ref_driver.cpp
                                 (CPU baseline / reference arm of bench.py, trajectory oracle)

Nothing is copied from the reference; TooN (third-party, vendored as a zip in the reference) is
unpacked into oracle/_ref/toon.  dgesvd_ comes from the OpenBLAS bundled with opencv-python-headless
(SURVEY.md section 8(c)).  If /root/reference is absent (GPU box) this script does nothing and the
prebuilt files are used.
"""
import glob
import os
import subprocess
import sys
import zipfile
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("REBVO_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")


def openblas():
    import importlib.util
    spec = importlib.util.find_spec("cv2")
    base = os.path.join(os.path.dirname(os.path.dirname(spec.origin)), "opencv_python_headless.libs")
    libs = glob.glob(os.path.join(base, "libopenblas*.so"))
    if not libs:
        raise RuntimeError("no bundled OpenBLAS found for dgesvd_")
    return base, libs[0]


SHIM_XLIB = """#pragma once
typedef struct _XDisplay Display; typedef unsigned long Window; typedef unsigned long Pixmap; typedef unsigned long Atom;
typedef struct _XGC *GC; typedef struct { int x; } XWindowAttributes; typedef union { int type; long pad[24]; } XEvent;
typedef struct _XImage XImage;
"""
SHIM_FIX = """#pragma once
#include <algorithm>
#include <cstdio>
#include <unistd.h>
#include <memory>
namespace std {
inline double max(float a,double b){return a>b?(double)a:b;} inline double max(double a,float b){return a>b?a:(double)b;}
inline double min(float a,double b){return a<b?(double)a:b;} inline double min(double a,float b){return a<b?a:(double)b;} }
"""
SHIM_GD = """#pragma once
typedef struct gdImageStruct {int sx, sy;} gdImage; typedef gdImage* gdImagePtr;
#ifdef __cplusplus
extern "C" {
#endif
gdImagePtr gdImageCreateTrueColor(int,int); void gdFree(void*); void gdImageSetPixel(gdImagePtr,int,int,int);
void* gdImageJpegPtr(gdImagePtr,int*,int); gdImagePtr gdImageCreateFromJpeg(FILE*); gdImagePtr gdImageCreateFromPng(FILE*);
gdImagePtr gdImageCreateFromWBMP(FILE*); gdImagePtr gdImageCreateFromGif(FILE*); int gdImageGetTrueColorPixel(gdImagePtr,int,int);
void gdImageDestroy(gdImagePtr);
#ifdef __cplusplus
}
#endif
#define gdTrueColor(r,g,b) (((r)<<16)+((g)<<8)+(b))
#define gdTrueColorGetRed(c) (((c)&0xFF0000)>>16)
#define gdTrueColorGetGreen(c) (((c)&0x00FF00)>>8)
#define gdTrueColorGetBlue(c) ((c)&0x0000FF)
"""
STUBS = """// stubs for never-called codec / camera functions (level B link only)
#include <cstdio>
#include "gd.h"
#include "VideoLib/video_io.h"
extern "C" {
gdImagePtr gdImageCreateTrueColor(int,int){return nullptr;} void gdFree(void*){} void gdImageSetPixel(gdImagePtr,int,int,int){}
void* gdImageJpegPtr(gdImagePtr,int*,int){return nullptr;} gdImagePtr gdImageCreateFromJpeg(FILE*){return nullptr;}
gdImagePtr gdImageCreateFromPng(FILE*){return nullptr;} gdImagePtr gdImageCreateFromWBMP(FILE*){return nullptr;}
gdImagePtr gdImageCreateFromGif(FILE*){return nullptr;} int gdImageGetTrueColorPixel(gdImagePtr,int,int){return 0;}
void gdImageDestroy(gdImagePtr){}
}
namespace rebvo {   // never-called v4l2 camera functions, prototypes: include/VideoLib/video_io.h:103-109
int CamaraInit(const char*, struct camera_context*, struct Size2D, uint){return -1;}
int CamaraWaitFrame(struct camera_context*){return -1;}
int CamaraClose(struct camera_context*){return 0;}
int CamaraGrabFrame(struct camera_context*, union RGB24Pixel*, struct timeval*){return -1;}
union RGB24Pixel* CamaraGrabBuffer(struct camera_context*, struct timeval*){return nullptr;}
int CamaraReleaseBuffer(struct camera_context*){return 0;}
int SavePPM(char*, union RGB24Pixel*, __u32, __u32){return 0;}
}
"""


def run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("reference compile failed")


def build(level_b=True, force=False):
    if not os.path.isdir(REF):
        return False
    so = os.path.join(OUT, "libref_mtrack.so")
    exe = os.path.join(OUT, "ref_rebvo")
    srcs_mine = [os.path.join(HERE, "ref_shim.cpp"), os.path.join(HERE, "ref_driver.cpp"), __file__]
    newest = max(os.path.getmtime(s) for s in srcs_mine if os.path.exists(s))
    if (not force and os.path.exists(so) and os.path.getmtime(so) > newest
            and (not level_b or (os.path.exists(exe) and os.path.getmtime(exe) > newest))):
        return True
    os.makedirs(os.path.join(OUT, "shim", "X11"), exist_ok=True)
    os.makedirs(os.path.join(OUT, "obj"), exist_ok=True)
    toon = os.path.join(OUT, "toon")
    if not os.path.isdir(os.path.join(toon, "TooN")):
        zipfile.ZipFile(os.path.join(REF, "TooN-2.2.zip")).extractall(toon)
        if not os.path.exists(os.path.join(toon, "TooN")):
            os.symlink("TooN-2.2", os.path.join(toon, "TooN"))
    open(os.path.join(OUT, "shim", "libv4l2.h"), "w").write("")
    open(os.path.join(OUT, "shim", "X11", "Xlib.h"), "w").write(SHIM_XLIB)
    open(os.path.join(OUT, "shim", "fix_gcc13.h"), "w").write(SHIM_FIX)
    open(os.path.join(OUT, "shim", "gd.h"), "w").write(SHIM_GD)
    open(os.path.join(OUT, "shim", "stubs.cpp"), "w").write(STUBS)
    blas_dir, blas = openblas()
    # the reference's own flags: rebvolib/Makefile:16-17  (-m64 -O2 -fPIC -std=c++11), no -march, no NE10
    cxx = ["g++", "-std=c++11", "-O2", "-m64", "-fPIC", "-w", "-include",
           os.path.join(OUT, "shim", "fix_gcc13.h"), "-I" + os.path.join(OUT, "shim"),
           "-I" + os.path.join(REF, "include"), "-I" + os.path.join(REF, "src"), "-I" + toon]
    link = [blas, "-Wl,-rpath," + blas_dir, "-Wl,--allow-shlib-undefined", "-lpthread"]

    a_srcs = [os.path.join(REF, "src/mtracklib", f + ".cpp") for f in
              ("sspace", "iigauss", "iimage", "edge_finder", "edge_tracker")]
    a_srcs += [os.path.join(REF, "src/UtilLib/ne10wrapper.cpp"), os.path.join(REF, "src/VideoLib/image_undistort.cpp"),
               os.path.join(REF, "src/CommLib/net_keypoint.cpp"), os.path.join(REF, "src/mtracklib/scaleestimator.cpp"),
               os.path.join(REF, "src/UtilLib/imugrabber.cpp"), os.path.join(REF, "src/UtilLib/configurator.cpp"),
               os.path.join(HERE, "ref_shim.cpp")]
    b_names = ["rebvo/rebvo", "rebvo/rebvo_first_t", "rebvo/rebvo_second_t", "rebvo/rebvo_third_t",
               "mtracklib/sspace", "mtracklib/iigauss", "mtracklib/iimage", "mtracklib/edge_finder",
               "mtracklib/edge_tracker", "mtracklib/global_tracker", "UtilLib/ne10wrapper",
               "mtracklib/keyframe", "mtracklib/kfvo", "mtracklib/pose_graph", "mtracklib/scaleestimator",
               "UtilLib/imugrabber", "UtilLib/configurator", "VideoLib/image_undistort",
               "CommLib/net_keypoint", "VideoLib/customcam", "VideoLib/simcam", "VideoLib/videocam",
               "VideoLib/video_encoder", "VideoLib/video_mfc", "VideoLib/video_mjpeg",
               "VideoLib/datasetcam", "VideoLib/v4lcam", "UtilLib/ttimer", "CommLib/udp_port",
               "visualizer/depth_filler"]
    b_srcs = [os.path.join(REF, "src", n + ".cpp") for n in b_names]
    b_srcs += [os.path.join(OUT, "shim", "stubs.cpp"), os.path.join(HERE, "ref_driver.cpp")]

    def obj(src, tag):
        o = os.path.join(OUT, "obj", tag + "_" + os.path.basename(src).replace(".cpp", ".o"))
        run(cxx + ["-c", src, "-o", o])
        return o

    jobs = [(s, "a") for s in a_srcs] + ([(s, "b") for s in b_srcs] if level_b else [])
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        objs = list(ex.map(lambda j: obj(*j), jobs))
    a_objs = objs[:len(a_srcs)]
    run(["g++", "-shared", "-o", so] + a_objs + link)
    if level_b:
        run(["g++", "-o", exe] + objs[len(a_srcs):] + link)
    return True


def build_shim_driver(force=False):
    """tests/shim_driver.cpp: the reference's call sequence compiled against include/rebvo_b200_shim.hpp with
    the reference's own non-hot-path headers (Image<>, cam_model, TooN).  Output: oracle/_ref/shim_driver."""
    if not os.path.isdir(REF):
        return False
    repo = os.path.dirname(HERE)
    src = os.path.join(repo, "tests", "shim_driver.cpp")
    exe = os.path.join(OUT, "shim_driver")
    lib = os.path.join(repo, "rebvo_b200", "librebvo_b200.so")
    deps = [src, os.path.join(repo, "include", "rebvo_b200_shim.hpp"), os.path.join(repo, "include", "rebvo_b200.h"), lib]
    if not force and os.path.exists(exe) and all(os.path.getmtime(exe) > os.path.getmtime(d) for d in deps):
        return True
    toon = os.path.join(OUT, "toon")
    run(["g++", "-std=c++11", "-O2", "-w", "-include", os.path.join(OUT, "shim", "fix_gcc13.h"),
         "-I" + os.path.join(OUT, "shim"), "-I" + os.path.join(repo, "include"), "-I" + os.path.join(REF, "include"),
         "-I" + toon, src, "-o", exe, lib, "-Wl,-rpath,$ORIGIN/../../rebvo_b200"])
    return True


HOT_PATH_TUS = ("mtracklib/sspace", "mtracklib/iigauss", "mtracklib/iimage", "mtracklib/edge_finder",
                "mtracklib/edge_tracker", "mtracklib/global_tracker")
FORWARDED_HEADERS = ("mtracklib/sspace.h", "mtracklib/edge_finder.h", "mtracklib/edge_tracker.h", "mtracklib/global_tracker.h")
B_NAMES = ["rebvo/rebvo", "rebvo/rebvo_first_t", "rebvo/rebvo_second_t", "rebvo/rebvo_third_t",
           "mtracklib/sspace", "mtracklib/iigauss", "mtracklib/iimage", "mtracklib/edge_finder",
           "mtracklib/edge_tracker", "mtracklib/global_tracker", "UtilLib/ne10wrapper",
           "mtracklib/keyframe", "mtracklib/kfvo", "mtracklib/pose_graph", "mtracklib/scaleestimator",
           "UtilLib/imugrabber", "UtilLib/configurator", "VideoLib/image_undistort",
           "CommLib/net_keypoint", "VideoLib/customcam", "VideoLib/simcam", "VideoLib/videocam",
           "VideoLib/video_encoder", "VideoLib/video_mfc", "VideoLib/video_mjpeg",
           "VideoLib/datasetcam", "VideoLib/v4lcam", "UtilLib/ttimer", "CommLib/udp_port",
           "visualizer/depth_filler"]


def build_shim_rebvo(force=False):
    """The drop-in claim, executed: every UNMODIFIED translation unit of the reference's library except the six hot-path
    ones (rebvo*.cpp with its three threads, keyframe / kfvo / pose_graph, scaleestimator, CommLib, VideoLib, ...) is
    compiled where it lies against include/rebvo_b200_shim.hpp -- through a copy of the reference's include tree in which
    only the four hot-path headers are replaced by forwarding includes -- and linked with
    librebvo_b200.so and the same driver as level B.  Output: oracle/_ref/shim_rebvo."""
    if not os.path.isdir(REF):
        return False
    repo = os.path.dirname(HERE)
    exe = os.path.join(OUT, "shim_rebvo")
    lib = os.path.join(repo, "rebvo_b200", "librebvo_b200.so")
    deps = [os.path.join(HERE, "ref_driver.cpp"), os.path.join(repo, "include", "rebvo_b200_shim.hpp"),
            os.path.join(repo, "include", "rebvo_b200.h"), lib, __file__]
    if not force and os.path.exists(exe) and all(os.path.getmtime(exe) > os.path.getmtime(d) for d in deps):
        return True
    ov = os.path.join(OUT, "include_overlay")
    inc = os.path.join(REF, "include")
    for d, _, fs in os.walk(inc):
        rel = os.path.relpath(d, inc)
        os.makedirs(os.path.join(ov, rel), exist_ok=True)
        for f in fs:
            r = os.path.normpath(os.path.join(rel, f))
            dst = os.path.join(ov, r)
            if os.path.lexists(dst):
                os.remove(dst)
            if r in FORWARDED_HEADERS:
                open(dst, "w").write("#pragma once\n#include <rebvo_b200_shim.hpp>\n")
            else:
                os.symlink(os.path.join(d, f), dst)   # untouched reference header
    toon = os.path.join(OUT, "toon")
    blas_dir, blas = openblas()
    cxx = ["g++", "-std=c++11", "-O2", "-m64", "-fPIC", "-w", "-include", os.path.join(OUT, "shim", "fix_gcc13.h"),
           "-I" + ov, "-I" + os.path.join(repo, "include"), "-I" + os.path.join(OUT, "shim"),
           "-I" + os.path.join(REF, "src"), "-I" + toon]
    srcs = [os.path.join(REF, "src", n + ".cpp") for n in B_NAMES if n not in HOT_PATH_TUS]
    srcs += [os.path.join(OUT, "shim", "stubs.cpp"), os.path.join(HERE, "ref_driver.cpp")]
    os.makedirs(os.path.join(OUT, "obj"), exist_ok=True)

    def obj(src):
        o = os.path.join(OUT, "obj", "s_" + os.path.basename(src).replace(".cpp", ".o"))
        run(cxx + ["-c", src, "-o", o])
        return o

    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        objs = list(ex.map(obj, srcs))
    run(["g++", "-o", exe] + objs + [lib, "-Wl,-rpath,$ORIGIN/../../rebvo_b200", blas, "-Wl,-rpath," + blas_dir,
                                      "-Wl,--allow-shlib-undefined", "-lpthread"])
    return True


if __name__ == "__main__":
    ok = build(level_b="--no-level-b" not in sys.argv, force="--force" in sys.argv)
    print("reference oracle built" if ok else "reference sources not present; using prebuilt oracle/_ref if any")
