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

#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <vector>

#include "UtilLib/cam_model.h"
#include "UtilLib/ne10wrapper.h"   // as the reference's global_tracker.h does (kfvo.cpp uses Ne10:: through it)
#include "UtilLib/util.h"
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
// REBVO::construct (rebvo.cpp:297-312) builds, per ring slot and in one thread, sspace(Sigma0,KSigma,size,3) first and
// edge_tracker(cam,255*3) second.  The sspace constructor therefore only records its plan and queues itself (per
// thread); the edge_tracker constructor that follows knows the camera, finds or creates the device context of that
// (camera, Sigma0, KSigma) -- one per distinct triple in the process, shared by all ring slots -- and binds the queued
// scale spaces to it.  Nothing here is a process-wide singleton: several REBVO objects with different cameras or
// Sigma0 can live in one process.
struct Plan {
    double s0 = 0, ks = 0;
    bool set = false;
};
struct DeviceKey {
    unsigned w, h;
    float ppx, ppy, zfx, zfy;
    double s0, ks;
    bool operator==(const DeviceKey &o) const {
        return w == o.w && h == o.h && ppx == o.ppx && ppy == o.ppy && zfx == o.zfx && zfy == o.zfy && s0 == o.s0 && ks == o.ks;
    }
};
struct Registry {
    std::mutex mtx;
    std::vector<std::pair<DeviceKey, std::weak_ptr<Device>>> devs;
};
inline Registry &registry() {
    static Registry r;
    return r;
}
inline std::shared_ptr<Device> get_device(const cam_model &cam, double s0, double ks) {
    const DeviceKey k = {(unsigned)cam.sz.w, (unsigned)cam.sz.h, cam.pp.x, cam.pp.y, cam.zf.x, cam.zf.y, s0, ks};
    Registry &r = registry();
    std::lock_guard<std::mutex> lk(r.mtx);
    for (size_t i = 0; i < r.devs.size();) {
        std::shared_ptr<Device> d = r.devs[i].second.lock();
        if (!d) {
            r.devs.erase(r.devs.begin() + i);
            continue;
        }
        if (r.devs[i].first == k) return d;
        i++;
    }
    std::shared_ptr<Device> d = std::make_shared<Device>(cam, s0, ks);
    r.devs.push_back(std::make_pair(k, std::weak_ptr<Device>(d)));
    return d;
}
}  // namespace b200

class edge_finder;
class sspace;
namespace b200 {
inline std::vector<sspace *> &pending_sspaces() {   // scale spaces constructed in this thread and not yet bound to a device
    static thread_local std::vector<sspace *> v;
    return v;
}
}  // namespace b200

// ---- sspace (include/mtracklib/sspace.h:30-64) ---------------------------------------------------------
class sspace {
    std::shared_ptr<b200::Device> dev;
    rb_map *map = nullptr;
    Size2D sz;
    b200::Plan plan;
    Image<DetectorImgType> host[5];  // Img(0), Img(1), DoG, Dx, Dy mirrors, filled on demand
    friend class edge_finder;

    void unqueue() {
        std::vector<sspace *> &q = b200::pending_sspaces();
        for (size_t i = 0; i < q.size(); i++)
            if (q[i] == this) {
                q.erase(q.begin() + i);
                break;
            }
    }
    bool ensure() {
        if (map) return true;
        if (!dev || !dev->ok()) return false;   // not bound yet: no edge_tracker was constructed after this sspace
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
        plan.s0 = sigma0;
        plan.ks = k_sigma;
        plan.set = true;
        b200::pending_sspaces().push_back(this);
    }
    ~sspace() {
        unqueue();
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
    std::vector<KeyLine> shadow;      // what the device held when the mirror was pulled
    std::vector<int32_t> mask_host;   // Image<int> img_mask_kl mirror (search_match on the host, kfvo.cpp:725)
    bool host_valid = false, host_touched = false, mask_valid = false;
    int kn = 0;
    float reTunedThresh = 0;

    void pull() {
        if (host_valid || !map) return;
        host_kl.resize(KEYLINE_MAX);
        dev->check(rb_map_sync_host_keylines(map, (rb_keyline *)host_kl.data(), KEYLINE_MAX, &kn));
        shadow.assign(host_kl.begin(), host_kl.begin() + kn);
        host_valid = true;
        host_touched = false;
    }
    // The reference's callers write keylines through operator[] / iterators (the reset loop rebvo_second_t.cpp:610-613,
    // kfvo's augmentation).  Every method that works on the device first pushes such edits: the mirror is compared with
    // the copy taken when it was pulled, so an untouched or merely read mirror costs nothing but a flag test.
    void flush_host_edits() {
        if (!host_valid || !host_touched || !map) return;
        host_touched = false;
        if ((int)shadow.size() == kn && (kn == 0 || memcmp(shadow.data(), host_kl.data(), sizeof(KeyLine) * (size_t)kn) == 0)) return;
        commit_host();
    }
    void device_changed() {   // the device copy was rewritten: drop the mirrors
        host_valid = false;
        host_touched = false;
    }
    const std::vector<int32_t> &mask() {
        if (!mask_valid && map) {
            mask_host.resize((size_t)cam_mod.sz.w * cam_mod.sz.h);
            dev->check(rb_map_get_mask(map, mask_host.data()));
            mask_valid = true;
        }
        return mask_host;
    }

   public:
    edge_finder(cam_model &cam, float max_i_value, int kl_num_max = KEYLINE_MAX) : cam_mod(cam) {
        (void)max_i_value;  // 255*3 (rebvo.cpp:300) is compiled into the detector
        (void)kl_num_max;
        std::vector<sspace *> &q = b200::pending_sspaces();
        b200::Plan pl;
        for (size_t i = q.size(); i-- > 0;)
            if (q[i]->sz.w == cam.sz.w && q[i]->sz.h == cam.sz.h) {
                pl = q[i]->plan;
                break;
            }
        if (!pl.set) throw std::runtime_error("rebvo_b200: construct the sspace (Sigma0, KSigma) before the edge_tracker, "
                                              "as REBVO::construct does (rebvo.cpp:299-300)");
        dev = b200::get_device(cam, pl.s0, pl.ks);
        for (size_t i = 0; i < q.size();)   // bind the scale spaces of this plan that wait for a device
            if (q[i]->sz.w == cam.sz.w && q[i]->sz.h == cam.sz.h && q[i]->plan.s0 == pl.s0 && q[i]->plan.ks == pl.ks) {
                q[i]->dev = dev;
                q.erase(q.begin() + i);
            } else {
                i++;
            }
        if (dev->ok()) dev->check(rb_map_create(dev->ctx, &map));
    }
    // edge_finder.cpp:42-52 (keyframes deep-copy an edge map, keyframe.cpp:28-35): device-side clone of keylines, mask
    // and, because the match field lives in the same ring-slot object here, of global_tracker's field as well
    edge_finder(const edge_finder &o)
        : cam_mod(o.cam_mod), dev(o.dev), host_valid(false), kn(o.kn), reTunedThresh(o.reTunedThresh) {
        const_cast<edge_finder &>(o).flush_host_edits();
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
        if (!map) return;
        rb_detect_params p = {plane_fit_size, pos_neg_thresh, dog_thresh, kl_max, kl_ref, gain, thresh_max, thresh_min};
        dev->check(rb_map_detect_ss(map, ss->handle(), &p, &tresh, &l_kl_num, &kn));
        device_changed();
        mask_valid = false;
    }
    int reEstimateThresh(int knum, int n) {  // edge_finder.cpp:373-405
        if (!map) return 0;
        flush_host_edits();
        dev->check(rb_map_reestimate_thresh(map, knum, n, &reTunedThresh));
        return reTunedThresh;
    }
    cam_model &GetCam() { return cam_mod; }
    int KNum() const { return kn; }
    float getThresh() { return reTunedThresh; }
    KeyLine &operator[](uint inx) {
        pull();
        if (host_kl.size() <= inx) host_kl.resize(KEYLINE_MAX);
        host_touched = true;
        return host_kl[inx];
    }
    typedef KeyLine *iterator;
    iterator begin() {
        pull();
        host_touched = true;
        return host_kl.data();
    }
    iterator end() {
        pull();
        return host_kl.data() + kn;
    }
    // push the host mirror to the device (called automatically by flush_host_edits before device work)
    void commit_host() {
        if (!host_valid || !map) return;
        dev->check(rb_map_load_keylines(map, (const rb_keyline *)host_kl.data(), kn, nullptr));   // the mask stays
        shadow.assign(host_kl.begin(), host_kl.begin() + kn);
    }
    void invalidate_host() { device_changed(); }
    // edge_finder.cpp:411-436: raw KeyLine records, {int kn; KeyLine kl[kn];}
    void dumpToBinaryFile(std::ofstream &file) {
        pull();
        file.write((const char *)&kn, sizeof(kn));
        file.write((const char *)host_kl.data(), sizeof(KeyLine) * (size_t)kn);
        std::cout << "\nDumpped " << kn << " Keylines, " << sizeof(kn) + sizeof(KeyLine) * (size_t)kn << " bytes";
    }
    void readFromBinaryFile(std::ifstream &file) {
        file.read((char *)&kn, sizeof(kn));
        if (kn < 0) kn = 0;
        if (kn > KEYLINE_MAX) kn = KEYLINE_MAX;
        host_kl.resize(KEYLINE_MAX);
        file.read((char *)host_kl.data(), sizeof(KeyLine) * (size_t)kn);
        host_valid = true;
        commit_host();
    }
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
        if (!map) return;
        flush_host_edits();
        dev->check(rb_map_rotate_keylines(map, R));
        device_changed();
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
        if (stereo_mode) fprintf(stderr, "\nrebvo_b200: directed_matching: stereo mode is not available (StereoAvaiable=0 only)\n");
        if (!map || !et0->map) return nmatch = 0;
        flush_host_edits();
        et0->flush_host_edits();
        dev->check(rb_directed_matching(map, et0->map, V, RV, BR, min_thr_mod, min_thr_ang, max_radius,
                                        loc_uncertainty, &nmatch));
        device_changed();
        return nmatch;
    }
    // edge_tracker.cpp:380-436 ("this" is the OLD map, et the new one)
    int FordwardMatch(edge_tracker *et, bool clear = false) {
        (void)clear;
        int n = 0;
        if (!map || !et->map) return 0;
        flush_host_edits();
        et->flush_host_edits();
        dev->check(rb_forward_match(map, et->map, &n));
        et->nmatch = n;
        et->device_changed();
        return n;
    }
    // edge_tracker.cpp:695-724
    void UpdateInverseDepthKalman(TooN::Vector<3> vel, TooN::Matrix<3, 3> RVel, TooN::Matrix<3, 3> RW0,
                                  double ReshapeQAbsolute, double ReshapeQRelative, double LocationUncertainty) {
        (void)RVel;
        (void)RW0;
        (void)ReshapeQRelative;  // accepted but unused by the ARLU variant the reference runs
        double V[3] = {vel[0], vel[1], vel[2]};
        if (!map) return;
        flush_host_edits();
        dev->check(rb_map_ekf_update(map, V, ReshapeQAbsolute, LocationUncertainty));
        device_changed();
    }
    // edge_tracker.cpp:1148-1186
    double EstimateQuantile(double s_rho_min, double s_rho_max, double percentile, int n) {
        double q = 1e3;
        if (!map) return q;
        flush_host_edits();
        dev->check(rb_map_quantile(map, s_rho_min, s_rho_max, percentile, n, &q));
        return q;
    }
    // edge_tracker.cpp:1104-1140
    double EstimateReScalingOpt(double &RKp, const double &s_rho_min, const uint &MatchNumMin, bool re_escale) {
        double Kp = 1;
        if (!map) return Kp;
        flush_host_edits();
        dev->check(rb_map_rescale_opt(map, s_rho_min, MatchNumMin, re_escale, &Kp, &RKp));
        if (re_escale) device_changed();
        return Kp;
    }
    // edge_tracker.cpp:87-148
    int Regularize_1_iter(double thresh) {
        int n = 0;
        if (!map) return 0;
        flush_host_edits();
        dev->check(rb_map_regularize(map, thresh, &n));
        device_changed();
        return n;
    }
    int NumMatches() { return nmatch; }
    // IMU mode (config 3).  edge_tracker.cpp:1207-1301: linearised 6-DoF estimate over the forward matches of this map
    bool ExtRotVel(const TooN::Vector<3> &vel, TooN::Matrix<6, 6> &Wx, TooN::Matrix<6, 6> &Rx, TooN::Vector<6> &X,
                   const double &LocUncert, double HubReweigth) {
        double v[3] = {vel[0], vel[1], vel[2]}, wx[36], rx[36], x[6];
        int ok = 0;
        if (!map) return false;
        flush_host_edits();
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
    // edge_tracker.cpp:158-295 on the HOST mirror.  The per-frame path runs it inside rb_directed_matching; this public form
    // only serves the keyframe tracker (kfvo.cpp:725, TrackKeyFrames=1), which stays the reference's CPU code and probes one
    // keyline at a time against this map's mask and keylines.  Same tests in the same order as the reference.
    int search_match(KeyLine &k, TooN::Vector<3> Vel, TooN::Matrix<3, 3> RVel, TooN::Matrix<3, 3> BackRot, double min_thr_mod,
                     double min_thr_ang, double max_radius, double loc_uncertainty) {
        pull();
        const std::vector<int32_t> &msk = mask();
        if (msk.empty()) return -1;
        const double zf = cam_mod.zfm, cang_min_edge = cos(min_thr_ang * M_PI / 180.0);
        const TooN::Vector<3> p3 = BackRot * TooN::makeVector(k.p_m.x, k.p_m.y, zf);
        Point2DF p_m;
        p_m.x = p3[0] * zf / p3[2];
        p_m.y = p3[1] * zf / p3[2];
        const double k_rho = k.rho * zf / p3[2];
        const Point2DF pi0 = cam_mod.Hom2Img(p_m);
        double t_x = -(Vel[0] * zf - Vel[2] * p_m.x), t_y = -(Vel[1] * zf - Vel[2] * p_m.y);
        double norm_t = util::norm(t_x, t_y);
        const TooN::Vector<3> DrDv = TooN::makeVector(zf, zf, -p_m.x - p_m.y);
        const double sigma2_t = (DrDv.as_row() * RVel * DrDv.as_col())(0, 0);
        double dq_min, dq_max, dq_rho;
        int t_steps;
        if (norm_t > 1e-6) {
            t_x /= norm_t;
            t_y /= norm_t;
            dq_rho = norm_t * k_rho;
            dq_min = std::max(0.0, norm_t * (k_rho - k.s_rho)) - loc_uncertainty;
            dq_max = std::min(max_radius, norm_t * (k_rho + k.s_rho)) + loc_uncertainty;
            if (dq_rho > dq_max) {
                dq_rho = (dq_max + dq_min) / 2;
                t_steps = util::round2int_positive(dq_rho);
            } else {
                t_steps = util::round2int_positive(std::max(dq_max - dq_rho, dq_rho - dq_min));
            }
        } else {
            norm_t = k.n_m;
            t_x = k.m_m.x / norm_t;
            t_y = k.m_m.y / norm_t;
            norm_t = 1;
            dq_min = -max_radius - loc_uncertainty;
            dq_max = max_radius + loc_uncertainty;
            dq_rho = 0;
            t_steps = dq_max;
        }
        const int w = cam_mod.sz.w, h = cam_mod.sz.h;
        double tn = dq_rho, tp = dq_rho + 1;
        for (int t_i = 0; t_i < t_steps; t_i++, tp += 1, tn -= 1)
            for (int dir = 0; dir < 2; dir++) {
                const double t = dir ? tp : tn;
                if (dir ? t > dq_max : t < dq_min) continue;
                const int xi = (int)round(t_x * t + pi0.x), yi = (int)round(t_y * t + pi0.y);   // Image::GetIndexRC (image.h:121-126)
                if (xi >= w || yi >= h || xi < 0 || yi < 0) continue;
                const int j = msk[(size_t)yi * w + xi];
                if (j < 0) continue;
                const KeyLine &o = host_kl[j];
                const double cang = (o.m_m.x * k.m_m.x + o.m_m.y * k.m_m.y) / ((double)o.n_m * (double)k.n_m);
                if (cang < cang_min_edge || fabs((double)o.n_m / (double)k.n_m - 1) > min_thr_mod) continue;
                const double v_rho_dr = loc_uncertainty * loc_uncertainty + o.s_rho * o.s_rho * norm_t * norm_t + sigma2_t * o.rho * o.rho;
                if (util::square(t - norm_t * o.rho) > v_rho_dr) continue;
                return j;
            }
        return -1;
    }
    // Stereo entry points (edge_tracker.h:49-51,80; called under StereoAvaiable only, rebvo_second_t.cpp:471,484): no shipped
    // configuration enables stereo and the hot path scoped here is monocular -- they reject loudly instead of guessing.
    int directed_matching_stereo(TooN::Vector<3> &, TooN::Matrix<3, 3> &, edge_tracker *, double, double, double, double, double,
                                 double, double) {
        fprintf(stderr, "\nrebvo_b200: directed_matching_stereo is not available (StereoAvaiable must be 0)\n");
        return 0;
    }
    int search_match_stereo(KeyLine &, TooN::Vector<3> &, TooN::Matrix<3, 3> &, cam_model &, double, double, double, double,
                            double, double, double) {
        fprintf(stderr, "\nrebvo_b200: search_match_stereo is not available (StereoAvaiable must be 0)\n");
        return -1;
    }
    void fuseStereoDepth() { fprintf(stderr, "\nrebvo_b200: fuseStereoDepth is not available (StereoAvaiable must be 0)\n"); }
    friend class global_tracker;
};

// ---- global_tracker (include/mtracklib/global_tracker.h:38-107) ------------------------------------------
struct gt_field_data {   // global_tracker.h:33-36
    int dist;
    int ikl;
};
class global_tracker {
    cam_model cam_mod;
    double max_r = 0;
    edge_tracker *klist_f = nullptr;
    std::vector<gt_field_data> field_host;   // host mirror of the match field (Calc_f_J_Complete, kfvo.cpp:45,1503)
    bool field_valid = false;

    const std::vector<gt_field_data> &field() {
        if (!field_valid && klist_f && klist_f->map) {
            field_host.resize((size_t)cam_mod.sz.w * cam_mod.sz.h);
            klist_f->dev->check(rb_map_get_field(klist_f->map, (int32_t *)field_host.data()));
            field_valid = true;
        }
        return field_host;
    }

   public:
    global_tracker(cam_model &cam) : cam_mod(cam) {}
    // (copy: member-wise like the reference's global_tracker.cpp:42-47; keyframe.cpp:28-35 re-points it with SetEdgeTracker)
    // global_tracker.cpp:61-105: the field lives in the ring slot of `klist`
    void build_field(edge_tracker &klist, int radius, float min_mod = -1) {
        max_r = radius;
        klist_f = &klist;
        field_valid = false;
        if (!klist.map) return;
        klist.flush_host_edits();
        klist.dev->check(rb_map_build_field(klist.map, radius, min_mod));
    }
    // global_tracker.cpp:116-165 on the host mirrors (keyframe tracker only, kfvo.cpp:45,1503)
    template <class T>
    inline T Calc_f_J_Complete(int f_inx, T &df_dx, T &df_dy, KeyLine &kl, const Point2D<T> &p, const T &max_r_, const T &simil_mod,
                               const T &simil_cang, int &mnum, const T rho, const T rho_tol, T &fi) {
        const std::vector<gt_field_data> &f = field();
        if (f.empty() || f[f_inx].ikl < 0) {
            df_dx = 0;
            df_dy = 0;
            return max_r_;
        }
        KeyLine &f_kl = (*klist_f)[f[f_inx].ikl];
        const double cang = (kl.m_m.x * f_kl.m_m.x + kl.m_m.y * f_kl.m_m.y) / kl.n_m;
        if (cang < simil_cang || fabs(kl.n_m / f_kl.n_m - 1) > simil_mod ||
            fabs(rho - f_kl.rho) > rho_tol * (f_kl.s_rho + kl.s_rho * rho / kl.rho)) {
            df_dx = 0;
            df_dy = 0;
            return max_r_;
        }
        const T dx = p.x - f_kl.c_p.x, dy = p.y - f_kl.c_p.y;
        fi = (dx * f_kl.u_m.x + dy * f_kl.u_m.y);
        df_dx = f_kl.u_m.x;
        df_dy = f_kl.u_m.y;
        mnum++;
        kl.m_id_f = f[f_inx].ikl;
        return fi;
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
        if (!klist_f || klist.KNum() <= 0 || !klist.map || !klist_f->map) return 0;
        klist.flush_host_edits();
        klist_f->flush_host_edits();
        double V[3] = {Vel[0], Vel[1], Vel[2]}, W[3] = {W0[0], W0[1], W0[2]}, RV[9], RW[9], WX[36], score = 0;
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) {   // (outputs that keep the caller's value on an early return)
                RV[i * 3 + j] = RVel(i, j);
                RW[i * 3 + j] = RW0(i, j);
            }
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
        if (!klist_f || klist.KNum() <= 0 || !klist.map || !klist_f->map) return 0;
        klist.flush_host_edits();
        klist_f->flush_host_edits();
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
