// hostmath.cu -- C exports of the host/device algebra in lm.cuh so that the CPU test-suite can check it
// (SO3 exp/ln, LDL^T solve, symmetric pseudo-inverse solve) against the reference's TooN results without a GPU.
#include "lm.cuh"
#include "imu_filter.h"
#include <vector>

extern "C" {
void rb_hostmath_so3_exp(const double *w, double *R) { so3_exp(w, R); }
void rb_hostmath_so3_ln(const double *R, double *w) { so3_ln_of_matrix(R, w); }
void rb_hostmath_chol6_solve(const double *A, const double *b, double *x) {
    Chol6 ch;
    chol6_compute(A, &ch);
    chol6_backsub(&ch, b, x);
}
void rb_hostmath_chol6_inverse(const double *A, double *inv) {
    Chol6 ch;
    chol6_compute(A, &ch);
    chol6_inverse(&ch, inv);
}
void rb_hostmath_sym_svd_backsub(const double *A, int n, const double *b, double *x) { sym_svd_backsub(A, n, b, x); }
void rb_hostmath_solve_sym6_like_svd(const double *A, const double *b, double *x) { solve_sym6_like_svd(A, b, x); }
void rb_hostmath_mat3_inv(const double *A, double *B) { mat3_inv(A, B); }

// IMU-mode host filter chain (imu_filter.h) for the CPU parity test against the reference's ScaleEstimator / ImuGrabber
void *rb_hostmath_imu_hist_new() {
    rbimu::ImuFilterHist *h = new rbimu::ImuFilterHist;
    memset(h, 0, sizeof(*h));
    return h;
}
void rb_hostmath_imu_hist_free(void *h) { delete (rbimu::ImuFilterHist *)h; }
void rb_hostmath_est_acel_lsq4(void *h, const double *vel, double *acel, const double *R, double dt) {
    rbimu::est_acel_lsq4(*(rbimu::ImuFilterHist *)h, vel, acel, R, dt);
}
void rb_hostmath_mean_acel4(void *h, const double *s_acel, double *acel, const double *R) {
    rbimu::mean_acel4(*(rbimu::ImuFilterHist *)h, s_acel, acel, R);
}
double rb_hostmath_est_ka_gmek_bias(const double *s_acel, const double *f_acel, double kP, const double *Rot, double *X, double *P,
                                    const double *Qg, const double *Qrot, const double *Qbias, double QKp, double Rg,
                                    const double *Rs, const double *Rf, double *g_est, double *b_est, const double *Wvw,
                                    double *Xvw, double g_gravit) {
    return rbimu::est_ka_gmek_bias(s_acel, f_acel, kP, Rot, X, P, Qg, Qrot, Qbias, QKp, Rg, Rs, Rf, g_est, b_est, Wvw, Xvw, g_gravit);
}
// ImuGrabber dataset mode: integrate the samples (n x 7: t, gyro, accel) over consecutive frame intervals ts[0..nf-1]
// (first interval starts at 0 like FirstThr's t0); out: nf x 20 {n, dt, Rot[9], giro[3], acel[3], cacel[3]}
int rb_hostmath_imu_integrate(const double *samples, int n, const double *ts, int nf, double *out) {
    std::vector<rbimu::ImuSample> s(n);
    for (int i = 0; i < n; i++) {
        s[i].t = samples[i * 7];
        for (int k = 0; k < 3; k++) {
            s[i].giro[k] = samples[i * 7 + 1 + k];
            s[i].acel[k] = samples[i * 7 + 4 + k];
        }
    }
    rbimu::ImuBuffer b;
    b.init(s.data(), n, nullptr, nullptr);
    double t0 = 0;
    for (int f = 0; f < nf; f++) {
        const rbimu::ImuIntegral d = b.grab(t0, ts[f]);
        double *o = out + f * 20;
        o[0] = d.n;
        o[1] = d.dt;
        for (int k = 0; k < 9; k++) o[2 + k] = d.Rot[k];
        for (int k = 0; k < 3; k++) {
            o[11 + k] = d.giro[k];
            o[14 + k] = d.acel[k];
            o[17 + k] = d.cacel[k];
        }
        t0 = ts[f];
    }
    return 0;
}
}
