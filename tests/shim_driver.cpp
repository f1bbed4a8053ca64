// shim_driver.cpp -- integration test of the drop-in boundary: the reference's per-frame call sequence
// (REBVO::FirstThr, src/rebvo/rebvo_first_t.cpp:259-272, and REBVO::SecondThread ImuMode=0,
// src/rebvo/rebvo_second_t.cpp:167-585) written against the shim classes of include/rebvo_b200_shim.hpp exactly
// as the reference writes it against its own mtracklib classes (same class / method names and arguments, TooN
// types, PipeBuffer-like slots).  Compiled with the reference's headers for the non-hot-path types
// (Image<>, cam_model, TooN); reads the raw frame file of oracle/ref_driver.cpp and writes the same record
// format, so tests/test_gpu_shim.py can compare it with the reference run and with rb_pipeline_push.
#include <TooN/so3.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "rebvo_b200_shim.hpp"

using namespace rebvo;
using namespace TooN;

struct OutRec {
    double t, Pos[3], PoseLie[3], Pose[9], Vel[3], RotLie[3];
    double dtp0, dtp1, K, Kp, s_rho_p;
    int kn, matches, est_ok, p_id;
    double Rot[9], RKp, dt;   // (layout of oracle/ref_driver.cpp's record; Rot / RKp stay zero here)
};

struct Slot {  // the hot-path members of PipeBuffer (include/rebvo/rebvo.h:312-351)
    sspace *ss;
    global_tracker *gt;
    edge_tracker *ef;
    Image<RGB24Pixel> *imgc;
    Image<float> *img;
    double t;
};

int main(int argc, char **argv) {
    if (argc < 3) return 2;
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 2;
    int hdr[3];
    if (fread(hdr, 4, 3, f) != 3) return 2;
    const int IW = hdr[0], IH = hdr[1], NF = hdr[2];
    // app/rebvorun/GlobalConfig_EuRoC_2.txt, TrackerInitType=2, no undistort
    cam_model::rad_tan_distortion kc = {0, 0, 0, 0, 0};
    cam_model cam({367.215f, 248.375f}, {458.654f, 457.296f}, kc, {(uint)IW, (uint)IH});
    const double Sigma0 = 3.56359, KSigma = 1.2599;
    const int DetectorPlaneFitSize = 2, ReferencePoints = 15000, MaxPoints = 40000, TrackPoints = 12000;
    const double DetectorPosNegThresh = 0.4, DetectorDoGThresh = 0.095259868922420, DetectorAutoGain = 5e-7,
                 DetectorMaxThresh = 0.5, DetectorMinThresh = 0.005;
    const double SearchRange = 40, QCutOffQuantile = 0.9, QCutOffNumBins = 100, TrackerMatchThresh = 0.5,
                 ReweigthDistance = 2, MatchThreshModule = 1, MatchThreshAngle = 45, LocationUncertaintyMatch = 2,
                 RegularizeThresh = 0.5, ReshapeQAbsolute = 1e-4, ReshapeQRelative = 1.6968e-4, LocationUncertainty = 1,
                 config_fps = 20;
    const int TrackerIterNum = 5, TrackerInitType = 2, TrackerInitIterNum = 2, MatchThreshold = 500;
    const uint MatchNumThresh = 0;

    const int NS = 3;
    std::vector<Slot> pipe(NS);
    for (Slot &pbuf : pipe) {  // rebvo.cpp:297-312
        pbuf.ss = new sspace(Sigma0, KSigma, cam.sz, 3);
        pbuf.ef = new edge_tracker(cam, 255 * 3);
        pbuf.gt = new global_tracker(pbuf.ef->GetCam());
        pbuf.img = new Image<float>(cam.sz);
        pbuf.imgc = new Image<RGB24Pixel>(cam.sz);
    }
    if (!pipe[0].ef->ok()) {
        fprintf(stderr, "rebvo_b200: no CUDA device / context failed\n");
        return 3;
    }

    // ---- SecondThread locals (:57-66)
    double Kp = 1, K = 1, error_vel = 0, error_score = 0, P_Kp = 5e-6;
    Vector<3> V = Zeros, W = Zeros, Pos = Zeros;
    Matrix<3, 3> R = Identity, Pose = Identity;
    Matrix<3, 3> P_V = Identity * 1e50, P_W = Identity * 1e-10;
    // ---- FirstThr locals (:92-94)
    int l_kl_num = 0;
    double tresh = 0.01;

    std::vector<OutRec> out;
    for (int n = 0; n < NF; n++) {
        Slot &pbuf = pipe[n % NS];
        double t;
        if (fread(&t, 8, 1, f) != 1) break;
        if (fread(pbuf.imgc->Data(), 3, (size_t)IW * IH, f) != (size_t)IW * IH) break;
        pbuf.t = t;
        // ---------------- FirstThr (:259-272)
        Image<float>::ConvertRGB2BW((*pbuf.img), *pbuf.imgc);
        pbuf.ss->build(*pbuf.img);
        pbuf.ef->detect(pbuf.ss, DetectorPlaneFitSize, DetectorPosNegThresh, DetectorDoGThresh, MaxPoints, tresh,
                        l_kl_num, ReferencePoints, DetectorAutoGain, DetectorMaxThresh, DetectorMinThresh);
        pbuf.ef->reEstimateThresh(TrackPoints, QCutOffNumBins);
        OutRec r;
        memset(&r, 0, sizeof(r));
        r.t = t;
        r.K = K;
        r.p_id = n;
        if (n == 0) {
            r.kn = pbuf.ef->KNum();
            out.push_back(r);
            continue;
        }
        // ---------------- SecondThread (:135-585), ImuMode=0
        Slot &new_buf = pbuf, &old_buf = pipe[(n - 1) % NS];
        bool EstimationOk = true;
        double dt_frame = (new_buf.t - old_buf.t);
        if (dt_frame < 0.001) dt_frame = 1 / config_fps;
        int klm_num = 0, num_kf_back_m = 0;
        P_V = Identity * 1e50;
        P_W = Identity * 1e50;
        R = Identity;
        double s_rho_q = old_buf.ef->EstimateQuantile(RHO_MIN, RHO_MAX, QCutOffQuantile, QCutOffNumBins);
        new_buf.gt->build_field(*new_buf.ef, SearchRange, new_buf.ef->getThresh());
        TooN::Matrix<6, 6, double> W_X;
        new_buf.gt->Minimizer_RV<double>(V, W, P_V, P_W, *old_buf.ef, TrackerMatchThresh, TrackerIterNum,
                                         TrackerInitType, ReweigthDistance, error_vel, error_score, s_rho_q,
                                         MatchNumThresh, TrackerInitIterNum, W_X);
        klm_num = old_buf.ef->FordwardMatch(new_buf.ef);
        SO3<> R0(W);
        R.T() = R0.get_matrix() * R.T();
        old_buf.ef->rotate_keylines(R0.get_matrix());
        if (util::isNaN(V) || util::isNaN(W)) {
            P_V = Identity * 1e50;
            V = Zeros;
            Kp = 1;
            P_Kp = 1e50;
            EstimationOk = false;
        } else {
            klm_num = new_buf.ef->directed_matching(V, P_V, R, old_buf.ef, num_kf_back_m, MatchThreshModule,
                                                    MatchThreshAngle, SearchRange, LocationUncertaintyMatch, false);
            if (klm_num < MatchThreshold) {
                P_V = Identity * 1e50;
                V = Zeros;
                Kp = 1;
                P_Kp = 10;
                EstimationOk = false;
            } else {
                new_buf.ef->Regularize_1_iter(RegularizeThresh);
                new_buf.ef->UpdateInverseDepthKalman(V, P_V, P_W, ReshapeQAbsolute, ReshapeQRelative,
                                                     LocationUncertainty);
                Kp = new_buf.ef->EstimateReScalingOpt(P_Kp, RHO_MAX, 1, false);
            }
        }
        Pose = Pose * R;
        Pos += -Pose * V * K;
        P_V /= dt_frame * dt_frame;
        Vector<3> PoseLie = SO3<>(Pose).ln(), RotLie = SO3<>(R).ln(), Vel = -V * K / dt_frame;
        for (int i = 0; i < 3; i++) {
            r.Pos[i] = Pos[i];
            r.PoseLie[i] = PoseLie[i];
            r.Vel[i] = Vel[i];
            r.RotLie[i] = RotLie[i];
            for (int j = 0; j < 3; j++) r.Pose[i * 3 + j] = Pose(i, j);
        }
        r.Kp = Kp;
        r.s_rho_p = s_rho_q;
        r.kn = new_buf.ef->KNum();
        r.matches = new_buf.ef->NumMatches();
        r.est_ok = EstimationOk;
        out.push_back(r);
    }
    fclose(f);
    // the callback-side view: iterate the AoS mirror like rebvo_nodelet.cpp:176-212 does
    double rho_sum = 0;
    int nk = 0;
    for (KeyLine &kl : *pipe[(NF - 1) % NS].ef) {
        rho_sum += kl.rho;
        nk++;
    }
    FILE *fo = fopen(argv[2], "wb");
    if (!fo) return 2;
    int cnt = (int)out.size(), sz = (int)sizeof(OutRec);
    fwrite(&cnt, 4, 1, fo);
    fwrite(&sz, 4, 1, fo);
    fwrite(out.data(), sizeof(OutRec), out.size(), fo);
    fclose(fo);
    printf("{\"frames\": %d, \"last_kn\": %d, \"mean_rho\": %.6f}\n", cnt, nk, nk ? rho_sum / nk : 0.0);
    for (Slot &pbuf : pipe) {
        delete pbuf.ss;
        delete pbuf.gt;
        delete pbuf.ef;
        delete pbuf.img;
        delete pbuf.imgc;
    }
    return 0;
}
