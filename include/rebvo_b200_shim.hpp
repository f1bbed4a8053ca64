// rebvo_b200_shim.hpp -- host-side C++ mirror of the reference's hot-path classes on top of the C ABI
// (include/rebvo_b200.h).  Compile it INSIDE the reference tree in place of
//     include/mtracklib/sspace.h, edge_finder.h, edge_tracker.h, global_tracker.h
// (INTEGRATION.md shows the four forwarding headers): REBVO::FirstThr (src/rebvo/rebvo_first_t.cpp:259-272),
// REBVO::SecondThread (src/rebvo/rebvo_second_t.cpp:172-487), the output callback users and the ROS nodelet
// keep their source unchanged -- same class names, method names, argument meaning and error behaviour.
//
// It depends only on headers that are NOT on the hot path and stay the reference's own:
//   "VideoLib/video_io.h" (Size2D, Point2DF, RGB24Pixel), "VideoLib/image.h" (Image<T>),
//   "UtilLib/cam_model.h" (cam_model) and TooN (vectors / matrices appear in the public signatures).
//
// Host mirror: consumers iterate 168-byte AoS KeyLine records (rebvo_nodelet.cpp:176-212,
// net_keypoint.cpp:29-75).  The device keeps SoA; the mirror is pulled lazily by operator[] / begin() and
// pushed back by commit_host() (only REBVO's reset loop writes keylines from the host,
// rebvo_second_t.cpp:610-613).
#ifndef REBVO_B200_SHIM_HPP
#define REBVO_B200_SHIM_HPP

#include <TooN/TooN.h>

#include <cstring>
#include <memory>
#include <stdexcept>
#include <vector>

#include "UtilLib/cam_model.h"
#include "VideoLib/image.h"
#include "VideoLib/video_io.h"
#include "rebvo_b200.h"

namespace rebvo {

constexpr double RHO_MAX = 20;   // include/mtracklib/edge_finder.h:38-40
constexpr double RHO_MIN = 1e-3;
constexpr double RhoInit = 1;
constexpr int KEYLINE_MAX = 50000;  // :43

typedef float DetectorImgType;

// struct KeyLine (include/mtracklib/edge_finder.h:45-91): identical layout to rb_keyline
struct KeyLine {
    int p_inx;
    Point2DF m_m, u_m;
    float n_m, score;
    Point2DF c_p;
    double rho, s_rho, rho_nr, s_rho_nr, rho0, s_rho0;
    Point2DF p_m, p_m_0;
    int m_id, m_id_f, m_id_kf, m_num;
    Point2DF m_m0;
    double n_m0;
    int p_id, n_id, net_id, stereo_m_id;
    double stereo_rho, stereo_s_rho;
};
static_assert(sizeof(KeyLine) == sizeof(rb_keyline) && sizeof(KeyLine) == 168, "KeyLine layout");

namespace b200 {
// One device context per (camera, scale-space plan); shared by all ring slots created with the same arguments.
// CUDA errors are sticky: ok() turns false and REBVO::Running() should follow it (SURVEY.md 8(b) error row).
struct Device {
    rb_ctx *ctx = nullptr;
    bool failed = false;
    Device(const cam_model &cam, double sigma0, double ksigma, int device = 0) {
        rb_camera c = {(int)cam.sz.w, (int)cam.sz.h, cam.pp.x, cam.pp.y, cam.zf.x, cam.zf.y};
        failed = rb_ctx_create(&ctx, device, &c, sigma0, ksigma, KEYLINE_MAX) != RB_OK;
    }
    ~Device() {
        if (ctx) rb_ctx_destroy(ctx);
    }
    bool ok() const { return ctx && !failed; }
    const char *error() const { return rb_last_error(ctx); }
    void check(int r) {
        if (r != RB_OK) failed = true;   // no exceptions on the hot path (reference convention)
    }
};
// REBVO::construct (rebvo.cpp:297-312) builds sspace(Sigma0,KSigma,size,3) first and edge_tracker(cam,255*3)
// second, per ring slot: the sspace constructor records the scale-space plan, the first object that knows the
// camera creates the device context, everything else attaches to it lazily.
struct Plan {
    double s0 = 0, ks = 0;
    bool set = false;
};
inline Plan &plan() {
    static Plan p;
    return p;
}
inline std::shared_ptr<Device> &device_slot() {
    static std::shared_ptr<Device> d;
    return d;
}
inline std::shared_ptr<Device> get_device(const cam_model *cam) {
    std::shared_ptr<Device> &d = device_slot();
    if (!d && cam && plan().set) d = std::make_shared<Device>(*cam, plan().s0, plan().ks);
    return d;
}
}  // namespace b200

class edge_finder;

// ---- sspace (include/mtracklib/sspace.h:30-64) ---------------------------------------------------------
class sspace {
    std::shared_ptr<b200::Device> dev;
    rb_map *map = nullptr;
    Size2D sz;
    Image<DetectorImgType> host[5];  // Img(0), Img(1), DoG, Dx, Dy mirrors, filled on demand
    friend class edge_finder;

    bool ensure() {
        if (map) return true;
        dev = b200::get_device(nullptr);
        if (!dev || !dev->ok()) return false;
        dev->check(rb_map_create(dev->ctx, &map));
        return map != nullptr;
    }
    Image<DetectorImgType> &plane(int which, int slot) {
        if (host[slot].bSize() != sz.w * sz.h) host[slot] = Image<DetectorImgType>(sz);
        if (ensure()) dev->check(rb_map_get_plane(map, which, host[slot].Data()));
        return host[slot];
    }

   public:
    sspace(double sigma0, double k_sigma, const Size2D &size, int bf_num) : sz(size) {
        if (bf_num != 3) throw std::invalid_argument("rebvo_b200: sspace is built for 3 box filters (rebvo.cpp:299)");
        b200::Plan &p = b200::plan();
        p.s0 = sigma0;
        p.ks = k_sigma;
        p.set = true;
    }
    ~sspace() {
        if (map) rb_map_destroy(map);
    }
    sspace(const sspace &) = delete;
    // sspace::build(Image<DetectorImgType>&) (sspace.cpp:52-60).  The gray image is produced by
    // Image<float>::ConvertRGB2BW on the host in the reference; both entry points are kept.
    void build(Image<DetectorImgType> &data) {
        if (!ensure()) return;
        dev->check(rb_map_upload_gray(map, data.Data()));
        dev->check(rb_map_dog_build(map));
    }
    void build_rgb(Image<RGB24Pixel> &rgb) {  // fused H2D + ConvertRGB2BW + build
        if (!ensure()) return;
        dev->check(rb_map_upload_rgb(map, (const uint8_t *)rgb.Data()));
        dev->check(rb_map_dog_build(map));
    }
    Image<DetectorImgType> &Img(int inx) { return plane(inx ? 1 : 0, inx ? 1 : 0); }
    Image<DetectorImgType> &ImgDOG() { return plane(2, 2); }
    Image<DetectorImgType> &ImgDx() { return plane(3, 3); }
    Image<DetectorImgType> &ImgDy() { return plane(4, 4); }
    rb_map *handle() { return ensure() ? map : nullptr; }
};

// ---- edge_finder / edge_tracker (include/mtracklib/edge_finder.h:96-166, edge_tracker.h:31-81) -----------
class edge_finder {
   protected:
    cam_model cam_mod;
    std::shared_ptr<b200::Device> dev;
    rb_map *map = nullptr;
    std::vector<KeyLine> host_kl;
    bool host_valid = false;
    int kn = 0;
    float reTunedThresh = 0;

    void pull() {
        if (host_valid) return;
        host_kl.resize(KEYLINE_MAX);
        dev->check(rb_map_sync_host_keylines(map, (rb_keyline *)host_kl.data(), KEYLINE_MAX, &kn));
        host_valid = true;
    }

   public:
    edge_finder(cam_model &cam, float max_i_value, int kl_num_max = KEYLINE_MAX) : cam_mod(cam) {
        (void)max_i_value;  // 255*3 (rebvo.cpp:300) is compiled into the detector
        (void)kl_num_max;
        dev = b200::get_device(&cam);
        if (!dev) throw std::runtime_error("rebvo_b200: construct the sspace (Sigma0, KSigma) before the edge_tracker, "
                                           "as REBVO::construct does (rebvo.cpp:299-300)");
        if (dev->ok()) dev->check(rb_map_create(dev->ctx, &map));
    }
    // edge_finder.cpp:42-52 (keyframes deep-copy an edge map, keyframe.cpp:28-35): device-side clone of keylines, mask
    // and, because the match field lives in the same ring-slot object here, of global_tracker's field as well
    edge_finder(const edge_finder &o)
        : cam_mod(o.cam_mod), dev(o.dev), host_valid(false), kn(o.kn), reTunedThresh(o.reTunedThresh) {
        if (o.map && dev && dev->ok()) dev->check(rb_map_clone(o.map, &map));
    }
    edge_finder &operator=(const edge_finder &) = delete;
    ~edge_finder() {
        if (map) rb_map_destroy(map);
    }
    rb_map *handle() { return map; }
    bool ok() const { return dev && dev->ok(); }

    // edge_finder::detect (edge_finder.cpp:342-365)
    void detect(sspace *ss, int plane_fit_size, double pos_neg_thresh, double dog_thresh, int kl_max, double &tresh,
                int &l_kl_num, int kl_ref = 0, double gain = 0, double thresh_max = 1e10, double thresh_min = 1 - 10) {
        rb_detect_params p = {plane_fit_size, pos_neg_thresh, dog_thresh, kl_max, kl_ref, gain, thresh_max, thresh_min};
        dev->check(rb_map_detect_ss(map, ss->handle(), &p, &tresh, &l_kl_num, &kn));
        host_valid = false;
    }
    int reEstimateThresh(int knum, int n) {  // edge_finder.cpp:373-405
        dev->check(rb_map_reestimate_thresh(map, knum, n, &reTunedThresh));
        return reTunedThresh;
    }
    cam_model &GetCam() { return cam_mod; }
    int KNum() const { return kn; }
    float getThresh() { return reTunedThresh; }
    KeyLine &operator[](uint inx) {
        pull();
        return host_kl[inx];
    }
    typedef KeyLine *iterator;
    iterator begin() {
        pull();
        return host_kl.data();
    }
    iterator end() {
        pull();
        return host_kl.data() + kn;
    }
    // push host-side edits (REBVO's reset loop, rebvo_second_t.cpp:610-613) back to the device
    void commit_host() {
        if (!host_valid) return;
        std::vector<int32_t> mask((size_t)cam_mod.sz.w * cam_mod.sz.h);
        dev->check(rb_map_get_mask(map, mask.data()));
        dev->check(rb_map_load_keylines(map, (const rb_keyline *)host_kl.data(), kn, mask.data()));
    }
    void invalidate_host() { host_valid = false; }
};

class edge_tracker : public edge_finder {
    int nmatch = 0;

   public:
    using edge_finder::edge_finder;

    static void m3(const TooN::Matrix<3, 3> &M, double *o) {
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) o[i * 3 + j] = M(i, j);
    }
    // edge_tracker.cpp:42-76
    void rotate_keylines(TooN::Matrix<3, 3> RotF) {
        double R[9];
        m3(RotF, R);
        dev->check(rb_map_rotate_keylines(map, R));
        host_valid = false;
    }
    // edge_tracker.cpp:302-374
    int directed_matching(TooN::Vector<3> Vel, TooN::Matrix<3, 3> RVel, TooN::Matrix<3, 3> BackRot, edge_tracker *et0,
                          int &kf_matchs, double min_thr_mod, double min_thr_ang, double max_radius,
                          double loc_uncertainty, bool stereo_mode, bool clear = false) {
        (void)stereo_mode;
        (void)clear;
        double V[3] = {Vel[0], Vel[1], Vel[2]}, RV[9], BR[9];
        m3(RVel, RV);
        m3(BackRot, BR);
        kf_matchs = 0;
        dev->check(rb_directed_matching(map, et0->map, V, RV, BR, min_thr_mod, min_thr_ang, max_radius,
                                        loc_uncertainty, &nmatch));
        host_valid = false;
        return nmatch;
    }
    // edge_tracker.cpp:380-436 ("this" is the OLD map, et the new one)
    int FordwardMatch(edge_tracker *et, bool clear = false) {
        (void)clear;
        int n = 0;
        dev->check(rb_forward_match(map, et->map, &n));
        et->nmatch = n;
        et->host_valid = false;
        return n;
    }
    // edge_tracker.cpp:695-724
    void UpdateInverseDepthKalman(TooN::Vector<3> vel, TooN::Matrix<3, 3> RVel, TooN::Matrix<3, 3> RW0,
                                  double ReshapeQAbsolute, double ReshapeQRelative, double LocationUncertainty) {
        (void)RVel;
        (void)RW0;
        (void)ReshapeQRelative;  // accepted but unused by the ARLU variant the reference runs
        double V[3] = {vel[0], vel[1], vel[2]};
        dev->check(rb_map_ekf_update(map, V, ReshapeQAbsolute, LocationUncertainty));
        host_valid = false;
    }
    // edge_tracker.cpp:1148-1186
    double EstimateQuantile(double s_rho_min, double s_rho_max, double percentile, int n) {
        double q = 1e3;
        dev->check(rb_map_quantile(map, s_rho_min, s_rho_max, percentile, n, &q));
        return q;
    }
    // edge_tracker.cpp:1104-1140
    double EstimateReScalingOpt(double &RKp, const double &s_rho_min, const uint &MatchNumMin, bool re_escale) {
        double Kp = 1;
        dev->check(rb_map_rescale_opt(map, s_rho_min, MatchNumMin, re_escale, &Kp, &RKp));
        if (re_escale) host_valid = false;
        return Kp;
    }
    // edge_tracker.cpp:87-148
    int Regularize_1_iter(double thresh) {
        int n = 0;
        dev->check(rb_map_regularize(map, thresh, &n));
        host_valid = false;
        return n;
    }
    int NumMatches() { return nmatch; }
    // IMU mode (config 3).  edge_tracker.cpp:1207-1301: linearised 6-DoF estimate over the forward matches of this map
    bool ExtRotVel(const TooN::Vector<3> &vel, TooN::Matrix<6, 6> &Wx, TooN::Matrix<6, 6> &Rx, TooN::Vector<6> &X,
                   const double &LocUncert, double HubReweigth) {
        double v[3] = {vel[0], vel[1], vel[2]}, wx[36], rx[36], x[6];
        int ok = 0;
        dev->check(rb_ext_rot_vel(map, v, wx, rx, x, LocUncert, HubReweigth, &ok));
        for (int i = 0; i < 6; i++) {
            X[i] = x[i];
            for (int j = 0; j < 6; j++) {
                Wx(i, j) = wx[i * 6 + j];
                Rx(i, j) = rx[i * 6 + j];
            }
        }
        return ok != 0;
    }
    // edge_tracker.cpp:1308-1338: gyroscope-prior correction of the roto-translation estimate (host algebra)
    static void BiasCorrect(TooN::Vector<6> &X, TooN::Matrix<6, 6> &Wx, TooN::Vector<3> &Gb, TooN::Matrix<3, 3> &Wb,
                            const TooN::Matrix<3, 3> &Rg, const TooN::Matrix<3, 3> &Rb) {
        double x[6], wx[36], gb[3], wb[9], rg[9], rb[9];
        for (int i = 0; i < 6; i++) {
            x[i] = X[i];
            for (int j = 0; j < 6; j++) wx[i * 6 + j] = Wx(i, j);
        }
        for (int i = 0; i < 3; i++) gb[i] = Gb[i];
        m3(Wb, wb);
        m3(Rg, rg);
        m3(Rb, rb);
        rb_bias_correct(x, wx, gb, wb, rg, rb);
        for (int i = 0; i < 6; i++) {
            X[i] = x[i];
            for (int j = 0; j < 6; j++) Wx(i, j) = wx[i * 6 + j];
        }
        for (int i = 0; i < 3; i++) {
            Gb[i] = gb[i];
            for (int j = 0; j < 3; j++) Wb(i, j) = wb[i * 3 + j];
        }
    }
    friend class global_tracker;
};

// ---- global_tracker (include/mtracklib/global_tracker.h:38-107) ------------------------------------------
class global_tracker {
    cam_model cam_mod;
    double max_r = 0;
    edge_tracker *klist_f = nullptr;

   public:
    global_tracker(cam_model &cam) : cam_mod(cam) {}
    // global_tracker.cpp:61-105: the field lives in the ring slot of `klist`
    void build_field(edge_tracker &klist, int radius, float min_mod = -1) {
        max_r = radius;
        klist_f = &klist;
        klist.dev->check(rb_map_build_field(klist.map, radius, min_mod));
    }
    const double &getMaxSRadius() { return max_r; }
    void SetEdgeTracker(edge_tracker *et) { klist_f = et; }
    // global_tracker.cpp:578-819
    template <class T, bool UsePriors = false>
    double Minimizer_RV(TooN::Vector<3> &Vel, TooN::Vector<3> &W0, TooN::Matrix<3, 3> &RVel, TooN::Matrix<3, 3> &RW0,
                        edge_tracker &klist, double match_thresh, int iter_max, int init_type,
                        double reweigth_distance, double &rel_error, double &rel_error_score, const double &max_s_rho,
                        const uint &MatchNumThresh, const double &init_iter, TooN::Matrix<6, 6, T> &W_X) {
        static_assert(!UsePriors, "rebvo_b200: the shipped configurations never enable priors");
        if (!klist_f || klist.KNum() <= 0) return 0;
        double V[3] = {Vel[0], Vel[1], Vel[2]}, W[3] = {W0[0], W0[1], W0[2]}, RV[9], RW[9], WX[36], score = 0;
        klist.dev->check(rb_minimizer_rv(klist_f->map, klist.map, V, W, RV, RW, match_thresh, iter_max, init_type,
                                         reweigth_distance, &rel_error, &rel_error_score, max_s_rho, MatchNumThresh,
                                         (int)init_iter, WX, &score));
        for (int i = 0; i < 3; i++) {
            Vel[i] = V[i];
            W0[i] = W[i];
            for (int j = 0; j < 3; j++) {
                RVel(i, j) = RV[i * 3 + j];
                RW0(i, j) = RW[i * 3 + j];
            }
        }
        for (int i = 0; i < 6; i++)
            for (int j = 0; j < 6; j++) W_X(i, j) = (T)WX[i * 6 + j];
        klist.invalidate_host();  // m_id_f of the old map was rewritten
        return score;
    }
    // IMU mode (config 3).  global_tracker.cpp:1036-1093: translation-only Levenberg-Marquardt (the reference
    // instantiates float and double; both forward to the float64 kernels)
    template <class T>
    double Minimizer_V(TooN::Vector<3> &Vel, TooN::Matrix<3, 3> &RVel, edge_tracker &klist, T match_thresh, int iter_max,
                       T s_rho_min, uint MatchNumThresh, double reweigth_distance, float min_mod) {
        if (!klist_f || klist.KNum() <= 0) return 0;
        double V[3] = {Vel[0], Vel[1], Vel[2]}, RV[9], score = 0;
        klist.dev->check(rb_minimizer_v(klist_f->map, klist.map, V, RV, (double)match_thresh, iter_max, (double)s_rho_min,
                                        MatchNumThresh, reweigth_distance, min_mod, &score));
        for (int i = 0; i < 3; i++) {
            Vel[i] = V[i];
            for (int j = 0; j < 3; j++) RVel(i, j) = RV[i * 3 + j];
        }
        klist.invalidate_host();
        return score;
    }
};

}  // namespace rebvo
#endif  // REBVO_B200_SHIM_HPP
