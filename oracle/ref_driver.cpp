// TEST INFRASTRUCTURE ONLY -- never linked into, imported by or executed from the product path
// (bench.py's reference arm / cpu_baseline leg and the tests are the only callers).
//
// ref_driver.cpp: drives the UNMODIFIED reference `REBVO` class (three pipeline threads,
// src/rebvo/rebvo*.cpp) through its custom-camera API (include/rebvo/rebvo.h:548-567), the way
// app/rebvorun/main_custom_cam_example.cpp:52-82 does, on a raw frame file:
//
//   header  : int32 W, int32 H, int32 nframes
//   frame i : float64 timestamp, W*H*3 bytes RGB24
//
// and writes, per output callback (third thread, rebvo_third_t.cpp:329), one binary record
//   {double t; double Pos[3]; double PoseLie[3]; double Pose[9]; double V... ; int kn; int matches; double dtp0, dtp1, K, Kp}
// to the output file, then prints one JSON line with wall-clock fps.
//
// usage: ref_rebvo frames.bin out.bin key=value ...   (keys: the REBVOParameters fields set below)

#include "rebvo/rebvo.h"
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

using namespace rebvo;

struct OutRec {
    double t, Pos[3], PoseLie[3], Pose[9], Vel[3], RotLie[3];
    double dtp0, dtp1, K, Kp, s_rho_p;
    int kn, matches, est_ok, p_id;
    double Rot[9], RKp, dt;
};
static std::vector<OutRec> g_out;
static volatile int g_ncb = 0;

static bool callback(PipeBuffer &pb) {
    OutRec r;
    memset(&r, 0, sizeof(r));
    r.t = pb.t;
    for (int i = 0; i < 3; i++) {
        r.Pos[i] = pb.nav.Pos[i];
        r.PoseLie[i] = pb.nav.PoseLie[i];
        r.Vel[i] = pb.nav.Vel[i];
        r.RotLie[i] = pb.nav.RotLie[i];
        for (int j = 0; j < 3; j++) r.Pose[i * 3 + j] = pb.nav.Pose(i, j);
    }
    r.dtp0 = pb.dtp0;
    r.dtp1 = pb.dtp1;
    r.K = pb.K;
    r.Kp = pb.Kp;
    r.s_rho_p = pb.s_rho_p;
    r.kn = pb.ef->KNum();
    r.matches = pb.ef->NumMatches();
    r.est_ok = pb.EstimationOK;
    r.p_id = pb.p_id;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) r.Rot[i * 3 + j] = pb.nav.Rot(i, j);
    r.RKp = pb.RKp;
    r.dt = pb.dt;
    g_out.push_back(r);
    g_ncb++;
    return true;
}

int main(int argc, char **argv) {
    if (argc < 3) {
        fprintf(stderr, "usage: %s frames.bin out.bin [key=value ...]\n", argv[0]);
        return 2;
    }
    std::map<std::string, double> kv;
    std::map<std::string, std::string> skv;   // file names (ImuFile=..., SE3File=...)
    for (int i = 3; i < argc; i++) {
        char *eq = strchr(argv[i], '=');
        if (eq) {
            kv[std::string(argv[i], eq - argv[i])] = atof(eq + 1);
            skv[std::string(argv[i], eq - argv[i])] = std::string(eq + 1);
        }
    }
    auto get = [&](const char *k, double d) { return kv.count(k) ? kv[k] : d; };
    auto gets = [&](const char *k) { return skv.count(k) ? skv[k] : std::string(); };

    FILE *f = fopen(argv[1], "rb");
    if (!f) {
        fprintf(stderr, "cannot open %s\n", argv[1]);
        return 2;
    }
    int hdr[3];
    if (fread(hdr, 4, 3, f) != 3) return 2;
    const int W = hdr[0], H = hdr[1], NF = hdr[2];
    std::vector<double> ts(NF);
    std::vector<unsigned char> frames((size_t)NF * W * H * 3);
    for (int i = 0; i < NF; i++) {
        if (fread(&ts[i], 8, 1, f) != 1) return 2;
        if (fread(&frames[(size_t)i * W * H * 3], 1, (size_t)W * H * 3, f) != (size_t)W * H * 3) return 2;
    }
    fclose(f);

    // Every field must be set: REBVOParameters has no defaults (include/rebvo/rebvo.h:64-235).
    // Defaults below = app/rebvorun/GlobalConfig_EuRoC_2.txt with ImuMode=0, no undistort.
    REBVOParameters p;
    p.CameraType = 3;
    p.VideoNetHost = "127.0.0.1";
    p.VideoNetPort = 2708;
    p.VideoNetEnabled = false;
    p.BlockingUDP = false;
    p.VideoSave = 0;
    p.VideoSaveFile = "/tmp/ref_video.bin";
    p.VideoSaveBuffersize = 0;
    p.encoder_type = 0;
    p.encoder_dev = "";
    p.EdgeMapDelay = 0;
    p.SaveLog = get("SaveLog", 0) != 0;
    p.LogFile = gets("LogFile").empty() ? std::string("/tmp/ref_log.m") : gets("LogFile");
    p.TrayFile = gets("TrayFile").empty() ? std::string("/tmp/ref_tray.txt") : gets("TrayFile");
    p.TrackKeyFrames = false;
    p.KFSavePercent = 0.7;
    p.StereoAvaiable = false;
    p.DataSetFile = p.DataSetDir = p.DataSetFileStereo = p.DataSetDirStereo = "";
    p.CamTimeScale = 1;
    p.ImageSize = {(uint)W, (uint)H};
    p.z_f_x = get("ZfX", 458.654);
    p.z_f_y = get("ZfY", 457.296);
    p.pp_x = get("PPx", 367.215);
    p.pp_y = get("PPy", 248.375);
    p.kc = {get("KcR2", 0), get("KcR4", 0), get("KcR6", 0), get("KcP1", 0), get("KcP2", 0)};
    p.config_fps = get("FPS", 20);
    p.soft_fps = get("SoftFPS", 1e6);
    p.useUndistort = get("UseUndistort", 0) != 0;
    p.rotatedCam = false;
    p.CameraDevice = "";
    p.z_f_x_stereo = p.z_f_x;
    p.z_f_y_stereo = p.z_f_y;
    p.pp_x_stereo = p.pp_x;
    p.pp_y_stereo = p.pp_y;
    p.kc_stereo = {0, 0, 0, 0, 0};
    p.SimFile = "";
    p.sim_save_nframes = 0;
    p.simu_time_on = 0;
    p.simu_time_step = 0;
    p.simu_time_sweep = 0;
    p.simu_time_start = 0;
    // IMU fusion (config 3): ImuMode=2 reads the samples from a csv file "t,gx,gy,gz,ax,ay,az" (imugrabber.cpp:80-132)
    p.ImuMode = (int)get("ImuMode", 0);
    p.ImuFile = gets("ImuFile");
    p.SE3File = gets("SE3File");
    p.UseCamIMUSE3File = !p.SE3File.empty();
    p.ImuTimeScale = get("ImuTimeScale", 1);
    p.InitBias = get("InitBias", 0) != 0;
    p.InitBiasFrameNum = (int)get("InitBiasFrameNum", 10);
    p.BiasInitGuess = TooN::Zeros;
    p.GiroMeasStdDev = 1.6968e-4;
    p.GiroBiasStdDev = 1.9393e-5;
    p.AcelMeasStdDev = 2e-3;
    p.g_module = 9.8;
    p.g_module_uncer = 100e3;
    p.g_uncert = 2e-3;
    p.VBiasStdDev = 1e-7;
    p.ScaleStdDevMult = 1e-2;
    p.ScaleStdDevMax = 1e-4;
    p.ScaleStdDevInit = 1.2e-3;
    p.SampleTime = get("SampleTime", 0.005);
    p.CircBufferSize = 1000;
    p.TimeDesinc = 0;
    p.cpuSetAffinity = (int)get("SetAffinity", 0);
    p.cpu0 = (int)get("CPU0", 0);
    p.cpu1 = (int)get("CPU1", 1);
    p.cpu2 = (int)get("CPU2", 2);
    p.Sigma0 = get("Sigma0", 3.56359);
    p.KSigma = get("KSigma", 1.2599);
    p.DetectorPlaneFitSize = (int)get("DetectorPlaneFitSize", 2);
    p.DetectorPosNegThresh = get("DetectorPosNegThresh", 0.4);
    p.DetectorDoGThresh = get("DetectorDoGThresh", 0.095259868922420);
    p.ReferencePoints = (int)get("ReferencePoints", 15000);
    p.TrackPoints = (int)get("TrackPoints", 12000);
    p.MaxPoints = (int)get("MaxPoints", 40000);
    p.DetectorThresh = get("DetectorThresh", 0.01);
    p.DetectorAutoGain = get("DetectorAutoGain", 5e-7);
    p.DetectorMaxThresh = get("DetectorMaxThresh", 0.5);
    p.DetectorMinThresh = get("DetectorMinThresh", 0.005);
    p.MatchThreshold = (int)get("GlobalMatchThreshold", 500);
    p.SearchRange = get("SearchRange", 40);
    p.QCutOffNumBins = get("QCutOffNumBins", 100);
    p.QCutOffQuantile = get("QCutOffQuantile", 0.9);
    p.TrackerIterNum = (int)get("TrackerIterNum", 5);
    p.TrackerInitIterNum = (int)get("TrackerInitIterNum", 2);
    p.TrackerInitType = (int)get("TrackerInitType", 2);
    p.TrackerMatchThresh = get("TrackerMatchThresh", 0.5);
    p.LocationUncertaintyMatch = get("LocationUncertaintyMatch", 2);
    p.MatchThreshModule = get("MatchThreshModule", 1);
    p.MatchThreshAngle = get("MatchThreshAngle", 45);
    p.ReweigthDistance = get("ReweigthDistance", 2);
    p.MatchNumThresh = (uint)get("MatchNumThresh", 0);
    p.RegularizeThresh = get("RegularizeThresh", 0.5);
    p.ReshapeQAbsolute = get("ReshapeQAbsolute", 1e-4);
    p.ReshapeQRelative = get("ReshapeQRelative", 1.6968e-4);
    p.LocationUncertainty = get("LocationUncertainty", 1);
    p.DoReScaling = get("DoReScaling", 0);

    g_out.reserve(NF + 8);
    REBVO cf(p);
    cf.setOutputCallback(callback);
    if (!cf.Init()) {
        fprintf(stderr, "REBVO init failed\n");
        return 3;
    }
    const int warm = (int)get("Warmup", 0);
    auto t0 = std::chrono::steady_clock::now();
    int cb0 = 0;
    for (int i = 0; i < NF && cf.Running(); i++) {
        if (i == warm) {  // start the clock once `warm` frames went in
            t0 = std::chrono::steady_clock::now();
            cb0 = g_ncb;
        }
        std::shared_ptr<Image<RGB24Pixel>> ptr;
        while (!cf.requestCustomCamBuffer(ptr, ts[i], 0.5) && cf.Running()) {
        }
        memcpy((*ptr).Data(), &frames[(size_t)i * W * H * 3], (size_t)W * H * 3);
        cf.releaseCustomCamBuffer();
    }
    // An edge map reaches the callback only after it has served as old_buf, i.e. NF-2 callbacks.
    auto deadline = std::chrono::steady_clock::now() + std::chrono::seconds(30);
    while (g_ncb < NF - 2 && std::chrono::steady_clock::now() < deadline) usleep(1000);
    auto t1 = std::chrono::steady_clock::now();
    double wall = std::chrono::duration<double>(t1 - t0).count();
    int ncb = g_ncb;
    cf.CleanUp();

    FILE *fo = fopen(argv[2], "wb");
    if (fo) {
        int n = (int)g_out.size(), sz = (int)sizeof(OutRec);
        fwrite(&n, 4, 1, fo);
        fwrite(&sz, 4, 1, fo);
        fwrite(g_out.data(), sizeof(OutRec), g_out.size(), fo);
        fclose(fo);
    }
    double d0 = 0, d1 = 0;
    for (auto &r : g_out) {
        d0 += r.dtp0;
        d1 += r.dtp1;
    }
    int n = (int)g_out.size();
    printf("{\"frames_in\": %d, \"callbacks\": %d, \"timed_callbacks\": %d, \"wall_s\": %.6f, \"fps\": %.4f, "
           "\"mean_dtp0_ms\": %.4f, \"mean_dtp1_ms\": %.4f}\n",
           NF, ncb, ncb - cb0, wall, (ncb - cb0) / wall, n ? 1e3 * d0 / n : 0.0, n ? 1e3 * d1 / n : 0.0);
    return 0;
}
