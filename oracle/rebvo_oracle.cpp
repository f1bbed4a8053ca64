// TEST INFRASTRUCTURE ONLY -- never linked into, imported by or executed from the product path.
//
// rebvo_oracle.cpp: plain C-style CPU restatement ("port") of the reference algorithm for REBVO's per-frame
// edge pipeline.  Every function cites the reference file:line it follows (paths relative to
// /root/reference).  It is pinned by tests/test_oracle_port.py against (a) the unmodified reference compiled
// into oracle/_ref/libref_mtrack.so and (b) the golden vectors under tests/golden/ that were generated from
// that build (tests/golden/make_golden.py).  Float32 stages, keyline records, masks and fields must match
// the reference bit for bit; the pairwise-tree sums of TryVelRot are restated in the reference's order, so
// JtJ / JtF / score match bitwise as well; only the 6x6 SVD solve differs (LAPACK dgesvd_ there, Gaussian
// elimination here) and is compared to 1e-9.  The IMU-mode rows (TryVel, Minimizer_V, ExtRotVel, BiasCorrect) are restated
// at the end of the file: TryVel is sequential double arithmetic and matches bitwise, the 3x3 solves follow
// util::Matrix3x3Inv over TooN's Gaussian-elimination determinant (1e-12), ExtRotVel's SVD is again an elimination solve.
//
// Third-party arithmetic restated from its published algorithm: TooN 2.2 (vendored in the reference as
// TooN-2.2.zip): SO3::exp (so3.h:254-285), Cholesky<6> (Cholesky.h), dot products accumulate from zero.
//
// Build: oracle/build_port.py  (g++ -O2 -ffp-contract=off, no -march: same FP behaviour as the reference's
// `-m64 -O2` build, rebvolib/Makefile:16-17).
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

struct OKeyLine {  // struct KeyLine, include/mtracklib/edge_finder.h:45-91 (168 bytes)
    int32_t p_inx;
    float m_m[2], u_m[2], n_m, score, c_p[2];
    int32_t pad0;
    double rho, s_rho, rho_nr, s_rho_nr, rho0, s_rho0;
    float p_m[2], p_m_0[2];
    int32_t m_id, m_id_f, m_id_kf, m_num;
    float m_m0[2];
    double n_m0;
    int32_t p_id, n_id, net_id, stereo_m_id;
    double stereo_rho, stereo_s_rho;
};
static_assert(sizeof(OKeyLine) == 168, "KeyLine layout");

static const double RHO_MAX = 20, RHO_MIN = 1e-3, RHO_INIT = 1;  // edge_finder.h:38-40
static const int MAX_IMG_VALUE = 765;                            // rebvo.cpp:300 edge_tracker(cam,255*3)
static const int KEYLINE_MAX = 50000;                            // edge_finder.h:43

struct OMap {
    int w, h;
    float ppx, ppy;
    double zfm;
    int box[2][3];
    double sigma_r[2];
    std::vector<float> gray, img0, img1, dog, dx, dy;
    std::vector<int> mask;
    std::vector<OKeyLine> kl;
    int kn;
    float retuned;
    int nmatch;
    std::vector<int> fdist, fikl;  // gt_field_data
    double max_r;
    unsigned frame_count;
    double pinv[3][25];
};

// ---- iigauss::iigauss (src/mtracklib/iigauss.cpp:43-81) ---------------------------------------------
static void box_plan(double sigma, int n, int *d, double *sigma_r) {
    double wideal = sqrt(12 * sigma * sigma / n + 1);
    int wl = (int)wideal;
    if ((wl / 2) * 2 == wl) wl--;
    int m = (int)round((3 * n + 4 * n * wl + n * wl * wl - 12 * sigma * sigma) / (4 + 4 * wl));
    int i = 0;
    for (; i < m && i < n; i++) d[i] = wl;
    for (; i < n; i++) d[i] = wl + 2;
    *sigma_r = sqrt((m * wl * wl + (n - m) * (wl + 2.0) * (wl + 2.0) - n) / 12.0);
}

// ---- iimage::load (src/mtracklib/iimage.cpp:53-71) ----------------------------------------------------
static void integral(const float *in, float *I, int w, int h) {
    for (int y = 0; y < h; y++) {
        I[y * w] = in[y * w];
        for (int x = 1; x < w; x++) I[y * w + x] = I[y * w + x - 1] + in[y * w + x];
    }
    for (int x = 0; x < w; x++)
        for (int y = 1; y < h; y++) I[y * w + x] += I[(y - 1) * w + x];
}

// ---- iimage::build_average (iimage.cpp:134-179): reciprocal of the clipped box area -----------------
static float box_div(int x, int y, int w, int h, int d) {
    int d2 = d / 2;
    int cx = x < d2 + 1 ? x + d2 + 1 : (x < w - d2 ? d : w - x + d2);
    int cy = y < d2 + 1 ? y + d2 + 1 : (y < h - d2 ? d : h - y + d2);
    float area = (float)(cx * cy);
    return (float)(1.0 / area);
}

// ---- iimage::average (iimage.cpp:86-128), the nine regions spelled out ----------------------------------
static void box_average(float *out, const float *I, int w, int h, int d) {
    int d2 = d / 2;
    float a = 1.0 / (d * d);
#define P(X, Y) I[(Y) * w + (X)]
#define DIV box_div(x, y, w, h, d)
    int x, y;
    for (y = 0; y < d2 + 1; y++) {
        for (x = 0; x < d2 + 1; x++) out[y * w + x] = P(x + d2, y + d2) * DIV;
        for (; x < w - d2; x++) out[y * w + x] = (P(x + d2, y + d2) - P(x - d2 - 1, y + d2)) * DIV;
        for (; x < w; x++) out[y * w + x] = (P(w - 1, y + d2) - P(x - d2 - 1, y + d2)) * DIV;
    }
    for (; y < h - d2; y++) {
        for (x = 0; x < d2 + 1; x++) out[y * w + x] = (P(x + d2, y + d2) - P(x + d2, y - d2 - 1)) * DIV;
        for (; x < w - d2; x++)
            out[y * w + x] =
                (P(x + d2, y + d2) - P(x - d2 - 1, y + d2) - P(x + d2, y - d2 - 1) + P(x - d2 - 1, y - d2 - 1)) * a;
        for (; x < w; x++)
            out[y * w + x] =
                (P(w - 1, y + d2) - P(x - d2 - 1, y + d2) - P(w - 1, y - d2 - 1) + P(x - d2 - 1, y - d2 - 1)) * DIV;
    }
    for (; y < h; y++) {
        for (x = 0; x < d2 + 1; x++) out[y * w + x] = (P(x + d2, h - 1) - P(x + d2, y - d2 - 1)) * DIV;
        for (; x < w - d2; x++)
            out[y * w + x] =
                (P(x + d2, h - 1) - P(x + d2, y - d2 - 1) - P(x - d2 - 1, h - 1) + P(x - d2 - 1, y - d2 - 1)) * DIV;
        for (; x < w; x++)
            out[y * w + x] =
                (P(w - 1, h - 1) - P(w - 1, y - d2 - 1) - P(x - d2 - 1, h - 1) + P(x - d2 - 1, y - d2 - 1)) * DIV;
    }
#undef P
#undef DIV
}

// ---- iigauss::smooth (iigauss.cpp:91-101) ---------------------------------------------------------------
static void smooth(const float *in, float *out, int w, int h, const int *box) {
    std::vector<float> I((size_t)w * h);
    integral(in, I.data(), w, h);
    for (int i = 0; i < 2; i++) {
        box_average(out, I.data(), w, h, box[i]);
        integral(out, I.data(), w, h);
    }
    box_average(out, I.data(), w, h, box[2]);
}

// plane-fit pseudo inverse (edge_finder.cpp:83-100; util::Matrix3x3Inv, include/UtilLib/toon_util.h:32-41)
static void plane_fit_pinv(double pinv[3][25]) {
    double Phi[25][3];
    int k = 0;
    for (int i = -2; i <= 2; i++)
        for (int j = -2; j <= 2; j++, k++) {
            Phi[k][0] = j;
            Phi[k][1] = i;
            Phi[k][2] = 1;
        }
    double A[3][3];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) {
            double s = 0;
            for (k = 0; k < 25; k++) s += Phi[k][r] * Phi[k][c];
            A[r][c] = s;
        }
    double B[3][3];
    B[0][0] = A[2][2] * A[1][1] - A[2][1] * A[1][2];
    B[0][1] = -(A[2][2] * A[0][1] - A[2][1] * A[0][2]);
    B[0][2] = A[1][2] * A[0][1] - A[1][1] * A[0][2];
    B[1][0] = -(A[2][2] * A[1][0] - A[2][0] * A[1][2]);
    B[1][1] = A[2][2] * A[0][0] - A[2][0] * A[0][2];
    B[1][2] = -(A[1][2] * A[0][0] - A[1][0] * A[0][2]);
    B[2][0] = A[2][1] * A[1][0] - A[2][0] * A[1][1];
    B[2][1] = -(A[2][1] * A[0][0] - A[2][0] * A[0][1]);
    B[2][2] = A[1][1] * A[0][0] - A[1][0] * A[0][1];
    double det = A[0][0] * (A[1][1] * A[2][2] - A[1][2] * A[2][1]) - A[0][1] * (A[1][0] * A[2][2] - A[1][2] * A[2][0]) +
                 A[0][2] * (A[1][0] * A[2][1] - A[1][1] * A[2][0]);
    for (int r = 0; r < 3; r++)
        for (k = 0; k < 25; k++) {
            double s = 0;
            for (int c = 0; c < 3; c++) s += (B[r][c] / det) * Phi[k][c];
            pinv[r][k] = s;
        }
}

// ---- TooN SO3::exp (so3.h:219-285) ------------------------------------------------------------------------
static void so3_exp(const double *w, double *R) {
    double theta_sq = 0;
    for (int i = 0; i < 3; i++) theta_sq += w[i] * w[i];
    double theta = sqrt(theta_sq), A, B;
    if (theta_sq < 1e-8) {
        A = 1.0 - (1.0 / 6.0) * theta_sq;
        B = 0.5;
    } else if (theta_sq < 1e-6) {
        B = 0.5 - 0.25 * (1.0 / 6.0) * theta_sq;
        A = 1.0 - theta_sq * (1.0 / 6.0) * (1.0 - (1.0 / 20.0) * theta_sq);
    } else {
        double it = 1.0 / theta;
        A = sin(theta) * it;
        B = (1 - cos(theta)) * (it * it);
    }
    double wx2 = w[0] * w[0], wy2 = w[1] * w[1], wz2 = w[2] * w[2];
    R[0] = 1.0 - B * (wy2 + wz2);
    R[4] = 1.0 - B * (wx2 + wz2);
    R[8] = 1.0 - B * (wx2 + wy2);
    double a = A * w[2], b = B * (w[0] * w[1]);
    R[1] = b - a;
    R[3] = b + a;
    a = A * w[1];
    b = B * (w[0] * w[2]);
    R[2] = b + a;
    R[6] = b - a;
    a = A * w[0];
    b = B * (w[1] * w[2]);
    R[5] = b - a;
    R[7] = b + a;
}

// ---- Ne10 scalar helpers (include/UtilLib/ne10wrapper.h:319-361): pairwise tree sum -----------------------
static double pairwise_add(const double *src, int pnum) {
    std::vector<double> b0(pnum / 2 + 1), b1(pnum / 2 + 1);
    const double *in = src;
    double *out = b0.data();
    double odd = 0;
    bool swap = true;
    while (pnum > 3) {
        odd += (pnum & 1) ? in[pnum - 1] : 0;
        pnum >>= 1;
        for (int k = 0; k < pnum; k++) out[k] = in[k] + in[k + pnum];
        if (swap) {
            in = b0.data();
            out = b1.data();
        } else {
            in = b1.data();
            out = b0.data();
        }
        swap = !swap;
    }
    for (int i = 0; i < pnum; i++) odd += in[i];
    return odd;
}
static double dot_product(const double *a, const double *b, int pnum) {  // ne10wrapper.h:307-314
    std::vector<double> prod(pnum);
    for (int k = 0; k < pnum; k++) prod[k] = a[k] * b[k];
    return pairwise_add(prod.data(), pnum);
}

// ---- Cholesky<6> (TooN Cholesky.h) ---------------------------------------------------------------------------
struct Chol {
    double c[36];
};
static void chol_compute(const double *M, Chol *ch) {
    memcpy(ch->c, M, sizeof(double) * 36);
    double *a = ch->c;
    for (int col = 0; col < 6; col++) {
        double inv_diag = 1;
        for (int row = col; row < 6; row++) {
            double val = a[row * 6 + col];
            for (int c2 = 0; c2 < col; c2++) val -= a[c2 * 6 + col] * a[row * 6 + c2];
            if (row == col) {
                a[row * 6 + col] = val;
                if (val == 0) return;
                inv_diag = 1 / val;
            } else {
                a[col * 6 + row] = val;
                a[row * 6 + col] = val * inv_diag;
            }
        }
    }
}
static void chol_backsub(const Chol *ch, const double *v, double *x) {
    const double *a = ch->c;
    double y[6];
    for (int i = 0; i < 6; i++) {
        double val = v[i];
        for (int j = 0; j < i; j++) val -= a[i * 6 + j] * y[j];
        y[i] = val;
    }
    for (int i = 0; i < 6; i++) y[i] /= a[i * 6 + i];
    for (int i = 5; i >= 0; i--) {
        double val = y[i];
        for (int j = i + 1; j < 6; j++) val -= a[j * 6 + i] * x[j];
        x[i] = val;
    }
}
static void chol_inverse(const Chol *ch, double *inv) {
    const double *a = ch->c;
    double y[36];
    for (int i = 0; i < 6; i++) {
        double val[6];
        for (int c = 0; c < 6; c++) val[c] = i == c ? 1.0 : 0.0;
        for (int j = 0; j < i; j++)
            for (int c = 0; c < 6; c++) val[c] -= a[i * 6 + j] * y[j * 6 + c];
        for (int c = 0; c < 6; c++) y[i * 6 + c] = val[c];
    }
    for (int i = 0; i < 6; i++) {
        double s = 1 / a[i * 6 + i];
        for (int c = 0; c < 6; c++) y[i * 6 + c] *= s;
    }
    for (int i = 5; i >= 0; i--) {
        double val[6];
        for (int c = 0; c < 6; c++) val[c] = y[i * 6 + c];
        for (int j = i + 1; j < 6; j++)
            for (int c = 0; c < 6; c++) val[c] -= a[j * 6 + i] * inv[j * 6 + c];
        for (int c = 0; c < 6; c++) inv[i * 6 + c] = val[c];
    }
}
// stand-in for SVD<>::backsub on the (well conditioned, symmetric positive definite) ApI of the init
// iterations (global_tracker.cpp:659-661): Gaussian elimination with partial pivoting
static void solve6(const double *A, const double *b, double *x) {
    double M[6][7];
    for (int i = 0; i < 6; i++) {
        for (int j = 0; j < 6; j++) M[i][j] = A[i * 6 + j];
        M[i][6] = b[i];
    }
    for (int c = 0; c < 6; c++) {
        int p = c;
        for (int r = c + 1; r < 6; r++)
            if (fabs(M[r][c]) > fabs(M[p][c])) p = r;
        if (p != c)
            for (int j = 0; j < 7; j++) {
                double t = M[c][j];
                M[c][j] = M[p][j];
                M[p][j] = t;
            }
        for (int r = c + 1; r < 6; r++) {
            double f = M[r][c] / M[c][c];
            for (int j = c; j < 7; j++) M[r][j] -= f * M[c][j];
        }
    }
    for (int i = 5; i >= 0; i--) {
        double s = M[i][6];
        for (int j = i + 1; j < 6; j++) s -= M[i][j] * x[j];
        x[i] = s / M[i][i];
    }
}

extern "C" {

void *orc_map_create(int w, int h, float ppx, float ppy, float zfx, float zfy, double sigma0, double ksigma) {
    OMap *m = new OMap;
    m->w = w;
    m->h = h;
    m->ppx = ppx;
    m->ppy = ppy;
    m->zfm = (zfx + zfy) / 2;  // cam_model.h:49, float arithmetic
    box_plan(sigma0, 3, m->box[0], &m->sigma_r[0]);                  // sspace.cpp:36-46
    box_plan(m->sigma_r[0] * ksigma, 3, m->box[1], &m->sigma_r[1]);
    size_t N = (size_t)w * h;
    m->gray.assign(N, 0);
    m->img0.assign(N, 0);
    m->img1.assign(N, 0);
    m->dog.assign(N, 0);
    m->dx.assign(N, 0);
    m->dy.assign(N, 0);
    m->mask.assign(N, -1);
    m->kl.resize(KEYLINE_MAX);
    memset(m->kl.data(), 0, sizeof(OKeyLine) * KEYLINE_MAX);
    m->kn = 0;
    m->retuned = 0;
    m->nmatch = 0;
    m->fdist.assign(N, 0);
    m->fikl.assign(N, -1);
    m->max_r = 0;
    m->frame_count = 0;
    plane_fit_pinv(m->pinv);
    return m;
}
void orc_map_destroy(void *p) { delete (OMap *)p; }
void orc_box_plan(void *p, int *d, double *s) {
    OMap *m = (OMap *)p;
    for (int f = 0; f < 2; f++) {
        for (int i = 0; i < 3; i++) d[f * 3 + i] = m->box[f][i];
        s[f] = m->sigma_r[f];
    }
}
// Image<float>::ConvertRGB2BW (include/VideoLib/image.h:197-203)
void orc_rgb2bw(void *p, const unsigned char *rgb) {
    OMap *m = (OMap *)p;
    for (size_t i = 0; i < m->gray.size(); i++) m->gray[i] = rgb[3 * i] + rgb[3 * i + 1] + rgb[3 * i + 2];
}
// sspace::build (sspace.cpp:52-85)
void orc_build(void *p) {
    OMap *m = (OMap *)p;
    int w = m->w, h = m->h;
    smooth(m->gray.data(), m->img0.data(), w, h, m->box[0]);
    smooth(m->gray.data(), m->img1.data(), w, h, m->box[1]);
    for (size_t k = 0; k < m->dog.size(); k++) m->dog[k] = m->img1[k] - m->img0[k];
    for (int y = 1; y < h - 1; y++)
        for (int x = 1; x < w - 1; x++) {
            m->dx[y * w + x] = m->img0[y * w + x + 1] - m->img0[y * w + x - 1];
            m->dy[y * w + x] = m->img0[(y + 1) * w + x] - m->img0[(y - 1) * w + x];
        }
}
void orc_get_plane(void *p, int which, float *out) {
    OMap *m = (OMap *)p;
    std::vector<float> *v = which == 0 ? &m->img0 : which == 1 ? &m->img1 : which == 2 ? &m->dog : which == 3 ? &m->dx
                                                                                         : which == 4 ? &m->dy
                                                                                                      : &m->gray;
    memcpy(out, v->data(), sizeof(float) * v->size());
}

// NextPoint (edge_finder.cpp:221-296)
static int next_point(int x, int y, const float *mm, const int *mask, int w) {
    float tx = -mm[1], ty = mm[0];
    int sx = ty > 0 ? (tx > 0 ? 1 : -1) : (tx < 0 ? -1 : 1);
    int sy = ty > 0 ? 1 : -1;
    int k;
    if ((k = mask[y * w + x + sx]) >= 0) return k;
    if ((k = mask[(y + sy) * w + x]) >= 0) return k;
    if ((k = mask[(y + sy) * w + x + sx]) >= 0) return k;
    return -1;
}

// edge_finder::detect = UpdateThresh + build_mask + join_edges (edge_finder.cpp:67-214, 304-365)
int orc_detect(void *p, int plane_fit, double pos_neg, double dog_thresh_d, int kl_max, double *tresh, int *l_kl_num,
               int kl_ref, double gain, double tmax, double tmin) {
    OMap *m = (OMap *)p;
    if (plane_fit != 2) return -1;
    if (gain > 0) {  // UpdateThresh :330-335
        *tresh -= gain * (double)(kl_ref - *l_kl_num);
        *tresh = *tresh > tmax ? tmax : (*tresh < tmin ? tmin : *tresh);
    }
    const int w = m->w, h = m->h, win_s = 2;
    const float per_hist = (float)pos_neg, grad_thesh = (float)*tresh, dog_thesh = (float)dog_thresh_d;
    if (kl_max > KEYLINE_MAX) kl_max = KEYLINE_MAX;
    int kn = 0;
    bool full = false;
    for (int y = win_s; y < h - win_s && !full; y++) {
        for (int x = win_s; x < w - win_s; x++) {
            int idx = y * w + x;
            m->mask[idx] = -1;
            float n2gI = m->dx[idx] * m->dx[idx] + m->dy[idx] * m->dy[idx];
            float t1 = grad_thesh * MAX_IMG_VALUE;
            if (n2gI < t1 * t1) continue;
            int pn = 0;
            double Y[25];
            for (int i = -win_s, k = 0; i <= win_s; i++)
                for (int j = -win_s; j <= win_s; j++, k++) {
                    float v = m->dog[(y + i) * w + x + j];
                    Y[k] = v;
                    if (v > 0) pn++;
                    else pn--;
                }
            if (fabs(pn) > ((float)((2.0 * win_s + 1.0) * (2.0 * win_s + 1.0))) * per_hist) continue;
            double th[3];
            for (int r = 0; r < 3; r++) {
                double s = 0;
                for (int k = 0; k < 25; k++) s += m->pinv[r][k] * Y[k];
                th[r] = s;
            }
            float xs = -th[0] * th[2] / (th[0] * th[0] + th[1] * th[1]);
            float ys = -th[1] * th[2] / (th[0] * th[0] + th[1] * th[1]);
            if (fabs(xs) > 0.5 || fabs(ys) > 0.5) continue;
            float mx = (float)th[0], my = (float)th[1];
            float n2_m = mx * mx + my * my;
            float t5 = grad_thesh * MAX_IMG_VALUE * dog_thesh;
            if (n2_m < t5 * t5) continue;
            OKeyLine &k = m->kl[kn];
            k.p_inx = idx;
            k.m_m[0] = mx;
            k.m_m[1] = my;
            k.n_m = sqrtf(n2_m);
            k.u_m[0] = k.m_m[0] / k.n_m;
            k.u_m[1] = k.m_m[1] / k.n_m;
            k.c_p[0] = x + xs;
            k.c_p[1] = y + ys;
            k.p_m[0] = k.c_p[0] - m->ppx;
            k.p_m[1] = k.c_p[1] - m->ppy;
            k.p_m_0[0] = k.p_m[0];
            k.p_m_0[1] = k.p_m[1];
            k.rho = k.rho0 = k.rho_nr = RHO_INIT;
            k.s_rho = k.s_rho0 = k.s_rho_nr = RHO_MAX;
            k.m_num = 0;
            k.n_id = k.p_id = k.net_id = -1;
            k.m_id = k.m_id_f = k.m_id_kf = -1;
            k.stereo_m_id = -1;
            k.stereo_rho = RHO_INIT;
            k.stereo_s_rho = RHO_MAX;
            m->mask[idx] = kn;
            if (++kn >= kl_max) {
                for (++idx; idx < w * h; idx++) m->mask[idx] = -1;
                full = true;
                break;
            }
        }
    }
    m->kn = kn;
    for (int i = 0; i < kn; i++) {  // join_edges :304-320
        int x = (int)(m->kl[i].c_p[0] + 0.5), y = (int)(m->kl[i].c_p[1] + 0.5);
        int j = next_point(x, y, m->kl[i].m_m, m->mask.data(), w);
        if (j < 0) continue;
        m->kl[j].p_id = i;
        m->kl[i].n_id = j;
    }
    *l_kl_num = kn;
    return kn;
}

// edge_finder::reEstimateThresh (edge_finder.cpp:373-405)
float orc_reestimate(void *p, int knum, int n) {
    OMap *m = (OMap *)p;
    if (m->kn <= 0) return m->retuned;
    float mx = m->kl[0].n_m, mn = m->kl[0].n_m;
    for (int i = 1; i < m->kn; i++) {
        if (m->kl[i].n_m > mx) mx = m->kl[i].n_m;
        if (m->kl[i].n_m < mn) mn = m->kl[i].n_m;
    }
    std::vector<int> histo(n + 1, 0);
    for (int i = 0; i < m->kn; i++) {
        int b = n * (mx - m->kl[i].n_m) / (mx - mn);
        b = b > n - 1 ? n - 1 : b;
        b = b < 0 ? 0 : b;
        histo[b]++;
    }
    int i = 0;
    for (int a = 0; i < n && a < knum; i++, a += histo[i])
        ;
    return m->retuned = mx - (float)i * (mx - mn) / (float)n;
}
int orc_knum(void *p) { return ((OMap *)p)->kn; }
void orc_get_keylines(void *p, void *out) {
    OMap *m = (OMap *)p;
    memcpy(out, m->kl.data(), sizeof(OKeyLine) * m->kn);
}
void orc_set_keylines(void *p, const void *in, int n) {
    OMap *m = (OMap *)p;
    memcpy(m->kl.data(), in, sizeof(OKeyLine) * n);
    m->kn = n;
}
void orc_get_mask(void *p, int *out) {
    OMap *m = (OMap *)p;
    memcpy(out, m->mask.data(), sizeof(int) * m->mask.size());
}
void orc_set_mask(void *p, const int *in) {
    OMap *m = (OMap *)p;
    memcpy(m->mask.data(), in, sizeof(int) * m->mask.size());
}

// edge_tracker::EstimateQuantile (edge_tracker.cpp:1148-1186)
double orc_quantile(void *p, double smin, double smax, double perc, int n) {
    OMap *m = (OMap *)p;
    std::vector<int> histo(n, 0);
    for (int k = 0; k < m->kn; k++) {
        int i = n * (m->kl[k].s_rho - smin) / (smax - smin);
        i = i > n - 1 ? n - 1 : i;
        i = i < 0 ? 0 : i;
        histo[i]++;
    }
    double s_rho = 1e3;
    for (int i = 0, a = 0; i < n; i++) {
        if (a > perc * m->kn) {
            s_rho = (double)i * (smax - smin) / (double)n + smin;
            break;
        }
        a += histo[i];
    }
    return s_rho;
}

static int index_rc(float x, float y, int w, int h) {  // Image::GetIndexRC (image.h:121-126)
    int xi = round(x), yi = round(y);
    if (xi >= w || yi >= h || xi < 0 || yi < 0) return -1;
    return yi * w + xi;
}

// global_tracker::build_field (global_tracker.cpp:61-105)
void orc_build_field(void *p, int radius, float min_mod) {
    OMap *m = (OMap *)p;
    m->max_r = radius;
    for (size_t i = 0; i < m->fikl.size(); i++) m->fikl[i] = -1;
    for (int ikl = 0; ikl < m->kn; ikl++) {
        OKeyLine &k = m->kl[ikl];
        if (min_mod > 0 && k.n_m < min_mod) continue;
        for (int t = -radius; t < radius; t++) {
            int inx = index_rc(k.u_m[0] * (float)t + k.c_p[0], k.u_m[1] * (float)t + k.c_p[1], m->w, m->h);
            if (inx < 0) continue;
            int at = abs(t);
            if (m->fikl[inx] >= 0 && at > m->fdist[inx]) continue;
            m->fdist[inx] = at;
            m->fikl[inx] = ikl;
        }
    }
}
void orc_get_field(void *p, int *out) {
    OMap *m = (OMap *)p;
    for (size_t i = 0; i < m->fikl.size(); i++) {
        out[2 * i] = m->fdist[i];
        out[2 * i + 1] = m->fikl[i];
    }
}

// global_tracker::TryVelRot<double,ReWeight,ProcJF,false> (global_tracker.cpp:285-543)
double orc_try_vel_rot(void *pn, void *po, const double *X, int RW, int PJ, double match_thresh, double s_rho_min,
                       unsigned match_num_thresh, double k_huber, const double *res_in, double *res_out, double *JtJ,
                       double *JtF) {
    OMap *mn = (OMap *)pn, *mo = (OMap *)po;
    const int K0 = mo->kn, pnum = (K0 + 3) & ~3;
    const double zf = mn->zfm, max_r = mn->max_r;
    double R0[9], RMf[9];
    so3_exp(X + 3, R0);
    double wz[3] = {0, 0, X[5]};
    so3_exp(wz, RMf);
    const double RM[4] = {RMf[0], RMf[1], RMf[3], RMf[4]};
    std::vector<double> P0(3 * pnum), Pt(3 * pnum), PI(3 * pnum), dfx(pnum, 0), dfy(pnum, 0), fm(pnum, 0);
    // KltoI3PMatrix (:552-570) + ProyI3Pto3PMatrix (ne10wrapper.h:413-424)
    for (int i = 0; i < pnum; i++) {
        double x = i < K0 ? mo->kl[i].p_m[0] : 0, y = i < K0 ? mo->kl[i].p_m[1] : 0, r = i < K0 ? mo->kl[i].rho : 1;
        double z = 1 / r, pz = (1 / zf) * z;
        P0[i] = pz * x;
        P0[pnum + i] = pz * y;
        P0[2 * pnum + i] = z;
    }
    for (int i = 0; i < pnum; i++) {  // SE3on3PMatrix (ne10wrapper.h:375-405)
        double x = P0[i], y = P0[pnum + i], z = P0[2 * pnum + i];
        for (int r = 0; r < 3; r++) {
            double t = R0[r * 3 + 0] * x;
            t += R0[r * 3 + 1] * y;
            t += R0[r * 3 + 2] * z;
            Pt[r * pnum + i] = X[r] + t;
        }
    }
    for (int i = 0; i < pnum; i++) {  // ProyP3toI3PMatrix (ne10wrapper.h:429-445)
        double rp = 1 / Pt[2 * pnum + i];
        double pz = zf * rp;
        PI[2 * pnum + i] = rp;
        PI[i] = pz * Pt[i];
        PI[pnum + i] = pz * Pt[pnum + i];
    }
    double fi = 0;
    unsigned mnt = match_num_thresh < mn->frame_count ? match_num_thresh : mn->frame_count;
    for (int i = 0; i < K0; i++) {
        OKeyLine &kl = mo->kl[i];
        kl.m_id_f = -1;
        if (kl.s_rho > s_rho_min || (unsigned)kl.m_num < mnt) continue;
        double px = PI[i] + mn->ppx, py = PI[pnum + i] + mn->ppy;
        int x = (int)(px + 0.5), y = (int)(py + 0.5);
        double weight = 1;
        if (RW && fabs(res_in[i]) > k_huber) weight = k_huber / fabs(res_in[i]);
        if (x < 1 || y < 1 || x >= mn->w - 1 || y >= mn->h - 1) {
            fm[i] = max_r;
            if (RW) fm[i] *= weight;
            res_out[i] = max_r;
            continue;
        }
        float mrx = RM[0] * kl.m_m[0] + RM[1] * kl.m_m[1];
        float mry = RM[2] * kl.m_m[0] + RM[3] * kl.m_m[1];
        int f = y * mn->w + x;
        double ret = max_r;  // Calc_f_J2 (:227-271)
        if (mn->fikl[f] >= 0) {
            OKeyLine &fk = mn->kl[mn->fikl[f]];
            double p_n2 = (kl.n_m * kl.n_m);  // Test_f_k (global_tracker.h:89-104)
            double p_esc = mrx * fk.m_m[0] + mry * fk.m_m[1];
            if (!(fabs(p_esc - p_n2) > match_thresh * p_n2)) {
                double dx = px - fk.c_p[0], dy = py - fk.c_p[1];
                fi = dx * fk.u_m[0] + dy * fk.u_m[1];
                dfx[i] = fk.u_m[0];
                dfy[i] = fk.u_m[1];
                kl.m_id_f = mn->fikl[f];
                ret = fi;
            }
        }
        fm[i] = ret;
        if (RW) {
            fm[i] *= weight;
            dfx[i] *= weight;
            dfy[i] *= weight;
        }
        res_out[i] = fi;  // keeps the previous match's fi on a miss (function-level variable, :341)
    }
    std::vector<double> Jm;
    if (PJ) {
        Jm.assign(6 * pnum, 0);
        for (int i = 0; i < pnum; i++) {  // :419-449
            double rp = PI[2 * pnum + i];
            double t = zf * rp;
            Jm[i] = t * dfx[i];
            Jm[pnum + i] = t * dfy[i];
            t = rp * PI[i];
            Jm[2 * pnum + i] = t * dfx[i];
            t = rp * PI[pnum + i];
            Jm[2 * pnum + i] += t * dfy[i];
            Jm[3 * pnum + i] = Jm[pnum + i] * Pt[2 * pnum + i];
            Jm[3 * pnum + i] += Jm[2 * pnum + i] * Pt[pnum + i];
            Jm[4 * pnum + i] = Jm[i] * Pt[2 * pnum + i];
            Jm[4 * pnum + i] += Jm[2 * pnum + i] * Pt[i];
            t = Jm[i] * Pt[pnum + i];
            Jm[5 * pnum + i] = -1 * t;
            Jm[5 * pnum + i] += Jm[pnum + i] * Pt[i];
        }
    }
    for (int i = 0; i < K0; i++) {  // :452-463 / :499-507
        double qvel = (zf * dfx[i] * X[0] + zf * dfy[i] * X[1] + (PI[i] * dfx[i] + PI[pnum + i] * dfy[i]) * X[2]);
        double q_rho = sqrt(mo->kl[i].s_rho * qvel * mo->kl[i].s_rho * qvel + 1);
        if (!RW) q_rho = mo->kl[i].s_rho;
        if (PJ)
            for (int j = 0; j < 6; j++) Jm[pnum * j + i] /= q_rho;
        fm[i] /= q_rho;
    }
    for (int i = K0; i < pnum; i++) {
        if (PJ)
            for (int j = 0; j < 6; j++) Jm[pnum * j + i] = 0;
        fm[i] = 0;
    }
    if (PJ) {
        for (int i = 0; i < 6; i++) {
            for (int j = i; j < 6; j++) JtJ[i * 6 + j] = dot_product(&Jm[pnum * i], &Jm[pnum * j], pnum);
            JtF[i] = dot_product(&Jm[pnum * i], fm.data(), pnum);
        }
        for (int i = 0; i < 2; i++) {  // :484-490
            JtF[i + 2] = -JtF[i + 2];
            for (int j = 0; j < 2; j++) {
                JtJ[(i + 0) * 6 + j + 2] = -JtJ[(i + 0) * 6 + j + 2];
                JtJ[(i + 2) * 6 + j + 4] = -JtJ[(i + 2) * 6 + j + 4];
            }
        }
        for (int i = 0; i < 6; i++)
            for (int j = i + 1; j < 6; j++) JtJ[j * 6 + i] = JtJ[i * 6 + j];
    }
    return dot_product(fm.data(), fm.data(), pnum);
}

// global_tracker::Minimizer_RV<double,false> (global_tracker.cpp:578-819)
double orc_minimizer_rv(void *pn, void *po, double *Vel, double *W0, double *RVel, double *RW0, double match_thresh,
                        int iter_max, int init_type, double reweight, double *rel_error, double *rel_error_score,
                        double max_s_rho, unsigned mnt, double init_iter, double *W_X) {
    OMap *mn = (OMap *)pn, *mo = (OMap *)po;
    if (mo->kn <= 0) return 0;
    const int pnum = (mo->kn + 3) & ~3;
    std::vector<double> Res0(pnum, 0), Res1(pnum, 0), Rest(pnum, 0);
    double *Residual = Res0.data(), *ResidualNew = Res1.data();
    double JtJ[36], JtF[6], JtJn[36], JtFn[6], ApI[36], h[6] = {0}, X[6], Xnew[6], Xt[6], nJtF[6];
    double F = 0, Fnew, F0 = 0, v = 2, tau = 1e-3, u = 0, gain;
    int eff_steps = 0;
    const double k_hubber = reweight;
    memset(JtJn, 0, sizeof(JtJn));
    memset(JtFn, 0, sizeof(JtFn));
    auto maxel = [](const double *M) {
        double m = M[0];
        for (int i = 1; i < 36; i++)
            if (M[i] > m) m = M[i];
        return m;
    };
    auto build = [&]() {
        for (int i = 0; i < 36; i++) ApI[i] = JtJ[i];
        for (int i = 0; i < 6; i++) {
            ApI[i * 6 + i] = JtJ[i * 6 + i] + u;
            nJtF[i] = -JtF[i];
        }
    };
    auto update = [&](bool with_den) {
        if (with_den) {
            double den = 0;
            for (int i = 0; i < 6; i++) den += (0.5 * h[i]) * (u * h[i] - JtF[i]);
            gain = (F - Fnew) / den;
        } else
            gain = F - Fnew;
        if (gain > 0) {
            F = Fnew;
            memcpy(X, Xnew, sizeof(X));
            memcpy(JtJ, JtJn, sizeof(JtJ));
            memcpy(JtF, JtFn, sizeof(JtF));
            double g = 2 * gain - 1, f = 1 - (g * g * g);
            u *= (0.33 > f ? 0.33 : f);
            v = 2;
            eff_steps++;
            return true;
        }
        u *= v;
        v *= 2;
        return false;
    };
    auto eval = [&](const double *Xe, int RW, int PJ, double *rin, double *rout, double *J, double *Fv) {
        return orc_try_vel_rot(pn, po, Xe, RW, PJ, match_thresh, max_s_rho, mnt, k_hubber, rin, rout, J, Fv);
    };
    if (init_type == 0) {
        memset(X, 0, sizeof(X));
    } else if (init_type == 1) {
        for (int i = 0; i < 3; i++) {
            X[i] = Vel[i];
            X[3 + i] = W0[i];
        }
    } else {
        double Ft = 0, F0t = 0, ut = 0, vt = 2;
        int eff_t = 0;
        for (int pass = 0; pass < 2; pass++) {
            double *rout = pass == 0 ? Rest.data() : ResidualNew;
            if (pass == 0) memset(X, 0, sizeof(X));
            else
                for (int i = 0; i < 3; i++) {
                    X[i] = Vel[i];
                    X[3 + i] = W0[i];
                }
            F = eval(X, 0, 1, Residual, rout, JtJ, JtF);
            F0 = F;
            u = tau * maxel(JtJ);
            if (pass == 1) v = 2;
            for (int i = 0; i < init_iter; i++) {
                build();
                solve6(ApI, nJtF, h);
                for (int k = 0; k < 6; k++) Xnew[k] = X[k] + h[k];
                if (i == init_iter - 1) {
                    Fnew = eval(Xnew, 0, 0, Residual, rout, JtJn, JtFn);
                    update(false);
                } else {
                    Fnew = eval(Xnew, 0, 1, Residual, rout, JtJn, JtFn);
                    update(true);
                }
            }
            if (pass == 0) {
                memcpy(Xt, X, sizeof(X));
                Ft = F;
                F0t = F0;
                ut = u;
                vt = v;
                eff_t = eff_steps;
                eff_steps = 0;
            } else {
                if (F > Ft) {
                    memcpy(X, Xt, sizeof(X));
                    F = Ft;
                    F0 = F0t;
                    u = ut;
                    v = vt;
                    eff_steps = eff_t;
                    ResidualNew = Rest.data();
                }
                double *t = ResidualNew;
                ResidualNew = Residual;
                Residual = t;
            }
        }
    }
    F0 = F = eval(X, 1, 1, Residual, ResidualNew, JtJ, JtF);
    u = tau * maxel(JtJ);
    v = 2;
    for (int it = 0; it < iter_max; it++) {
        build();
        Chol ch;
        chol_compute(ApI, &ch);
        chol_backsub(&ch, nJtF, h);
        for (int k = 0; k < 6; k++) Xnew[k] = X[k] + h[k];
        Fnew = eval(Xnew, 1, 1, Residual, ResidualNew, JtJn, JtFn);
        if (update(true)) {
            double *t = ResidualNew;
            ResidualNew = Residual;
            Residual = t;
        }
    }
    Chol ch;
    chol_compute(JtJ, &ch);
    double RRV[36];
    chol_inverse(&ch, RRV);
    for (int i = 0; i < 3; i++) {
        Vel[i] = X[i];
        W0[i] = X[3 + i];
        for (int j = 0; j < 3; j++) {
            RVel[i * 3 + j] = RRV[i * 6 + j];
            RW0[i * 3 + j] = RRV[(i + 3) * 6 + j + 3];
        }
    }
    memcpy(W_X, JtJ, sizeof(JtJ));
    if (eff_steps > 0) {
        double nh = 0, nx = 0;
        for (int i = 0; i < 6; i++) nh += h[i] * h[i];
        for (int i = 0; i < 6; i++) nx += X[i] * X[i];
        *rel_error = sqrt(nh) / (sqrt(nx) + 1e-30);
        *rel_error_score = F / F0;
    } else {
        *rel_error = 1e20;
        *rel_error_score = 1e20;
    }
    mn->frame_count++;
    return F;
}

// edge_tracker::FordwardMatch (edge_tracker.cpp:380-436)
int orc_forward_match(void *po, void *pn) {
    OMap *o = (OMap *)po, *e = (OMap *)pn;
    double nmatch = 0;
    for (int i = 0; i < o->kn; i++) {
        OKeyLine &k = o->kl[i];
        int f = k.m_id_f;
        if (f < 0 || f >= e->kn) continue;
        OKeyLine &t = e->kl[f];
        if (t.m_id >= 0 && t.rho > k.rho) continue;
        t.rho = k.rho;
        t.s_rho = k.s_rho;
        t.rho_nr = k.rho_nr;
        t.s_rho_nr = k.s_rho_nr;
        t.m_num = k.m_num + 1;
        t.m_id = i;
        t.p_m_0[0] = k.p_m[0];
        t.p_m_0[1] = k.p_m[1];
        t.m_m0[0] = k.m_m[0];
        t.m_m0[1] = k.m_m[1];
        t.n_m0 = k.n_m;
        t.m_id_kf = k.m_id_kf;
        nmatch++;
    }
    e->nmatch = nmatch;
    return nmatch;
}

// edge_tracker::rotate_keylines (edge_tracker.cpp:42-76)
void orc_rotate(void *p, const double *R) {
    OMap *m = (OMap *)p;
    const double zf = m->zfm;
    for (int i = 0; i < m->kn; i++) {
        OKeyLine &k = m->kl[i];
        double v[3] = {k.p_m[0] / zf, k.p_m[1] / zf, 1}, q[3];
        for (int r = 0; r < 3; r++) {
            double s = 0;
            for (int c = 0; c < 3; c++) s += R[r * 3 + c] * v[c];
            q[r] = s;
        }
        if (fabs(q[2]) > 0) {
            k.p_m[0] = q[0] / q[2] * zf;
            k.p_m[1] = q[1] / q[2] * zf;
            k.rho /= q[2];
            k.s_rho = k.s_rho / q[2];
        }
        double g[3] = {k.m_m[0], k.m_m[1], 0};
        for (int r = 0; r < 3; r++) {
            double s = 0;
            for (int c = 0; c < 3; c++) s += R[r * 3 + c] * g[c];
            q[r] = s;
        }
        k.m_m[0] = q[0];
        k.m_m[1] = q[1];
    }
}

// edge_tracker::search_match (edge_tracker.cpp:158-295); `o` = old map (mask + keylines searched)
static int search_match(OMap *o, const OKeyLine &k, const double *Vel, const double *RVel, const double *BR,
                        double min_thr_mod, double min_thr_ang, double max_radius, double loc_unc) {
    const double zf = o->zfm;
    const double cang_min_edge = cos(min_thr_ang * M_PI / 180.0);
    double dq_min = 0, dq_max = 0, t_x = 0, t_y = 0, dq_rho = 0;
    int t_steps = 0;
    double a[3] = {k.p_m[0], k.p_m[1], zf}, p3[3];
    for (int r = 0; r < 3; r++) {
        double s = 0;
        for (int c = 0; c < 3; c++) s += BR[r * 3 + c] * a[c];
        p3[r] = s;
    }
    float pmx = p3[0] * zf / p3[2], pmy = p3[1] * zf / p3[2];
    double k_rho = k.rho * zf / p3[2];
    float pi0x = pmx + o->ppx, pi0y = pmy + o->ppy;
    t_x = -(Vel[0] * zf - Vel[2] * pmx);
    t_y = -(Vel[1] * zf - Vel[2] * pmy);
    double norm_t = sqrt(t_x * t_x + t_y * t_y);
    double D[3] = {zf, zf, (double)(-pmx - pmy)}, row[3];
    for (int c = 0; c < 3; c++) {
        double s = 0;
        for (int r = 0; r < 3; r++) s += D[r] * RVel[r * 3 + c];
        row[c] = s;
    }
    double sigma2_t = 0;
    for (int c = 0; c < 3; c++) sigma2_t += row[c] * D[c];
    if (norm_t > 1e-6) {
        t_x /= norm_t;
        t_y /= norm_t;
        dq_rho = norm_t * k_rho;
        dq_min = fmax(0.0, norm_t * (k_rho - k.s_rho)) - loc_unc;
        dq_max = fmin(max_radius, norm_t * (k_rho + k.s_rho)) + loc_unc;
        if (dq_rho > dq_max) {
            dq_rho = (dq_max + dq_min) / 2;
            t_steps = (int)(dq_rho + 0.5);
        } else {
            t_steps = (int)(fmax(dq_max - dq_rho, dq_rho - dq_min) + 0.5);
        }
    } else {
        t_x = k.m_m[0];
        t_y = k.m_m[1];
        norm_t = k.n_m;
        t_x /= norm_t;
        t_y /= norm_t;
        norm_t = 1;
        dq_min = -max_radius - loc_unc;
        dq_max = max_radius + loc_unc;
        dq_rho = 0;
        t_steps = dq_max;
    }
    const double norm_m = k.n_m;
    double tn = dq_rho, tp = dq_rho + 1;
    for (int t_i = 0; t_i < t_steps; t_i++, tp += 1, tn -= 1) {
        for (int dir = 0; dir < 2; dir++) {
            double t;
            if (dir) {
                t = tp;
                if (t > dq_max) continue;
            } else {
                t = tn;
                if (t < dq_min) continue;
            }
            int inx = index_rc(t_x * t + pi0x, t_y * t + pi0y, o->w, o->h);
            if (inx < 0) continue;
            int j = o->mask[inx];
            if (j < 0) continue;
            const double norm_m0 = o->kl[j].n_m;
            double cang = (o->kl[j].m_m[0] * k.m_m[0] + o->kl[j].m_m[1] * k.m_m[1]) / (norm_m0 * norm_m);
            if (cang < cang_min_edge || fabs(norm_m0 / norm_m - 1) > min_thr_mod) continue;
            double s_rho = o->kl[j].s_rho, rho = o->kl[j].rho;
            double v_rho_dr = (loc_unc * loc_unc + s_rho * s_rho * norm_t * norm_t + sigma2_t * rho * rho);
            double e = t - norm_t * rho;
            if (e * e > v_rho_dr) continue;
            return j;
        }
    }
    return -1;
}

// edge_tracker::directed_matching (edge_tracker.cpp:302-374)
int orc_directed_matching(void *pn, void *po, const double *Vin, const double *RVin, const double *BR, double thr_mod,
                          double thr_ang, double max_radius, double loc_unc) {
    OMap *e = (OMap *)pn, *o = (OMap *)po;
    double Vel[3], t[9], RVel[9];
    for (int r = 0; r < 3; r++) {
        double s = 0;
        for (int c = 0; c < 3; c++) s += BR[r * 3 + c] * Vin[c];
        Vel[r] = s;
    }
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) {
            double s = 0;
            for (int k = 0; k < 3; k++) s += BR[r * 3 + k] * RVin[k * 3 + c];
            t[r * 3 + c] = s;
        }
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) {
            double s = 0;
            for (int k = 0; k < 3; k++) s += t[r * 3 + k] * BR[c * 3 + k];
            RVel[r * 3 + c] = s;
        }
    int nmatch = 0;
    for (int i = 0; i < e->kn; i++) {
        int j = search_match(o, e->kl[i], Vel, RVel, BR, thr_mod, thr_ang, max_radius, loc_unc);
        if (j < 0) continue;
        OKeyLine &k = e->kl[i], &q = o->kl[j];
        k.rho = q.rho;
        k.s_rho = q.s_rho;
        k.rho_nr = q.rho_nr;
        k.s_rho_nr = q.s_rho_nr;
        k.m_id = j;
        k.m_num = q.m_num + 1;
        k.p_m_0[0] = q.p_m[0];
        k.p_m_0[1] = q.p_m[1];
        k.m_m0[0] = q.m_m[0];
        k.m_m0[1] = q.m_m[1];
        k.n_m0 = q.n_m;
        k.m_id_kf = q.m_id_kf;
        nmatch++;
    }
    e->nmatch = nmatch;
    return nmatch;
}

// edge_tracker::Regularize_1_iter (edge_tracker.cpp:87-148)
int orc_regularize(void *p, double thresh) {
    OMap *m = (OMap *)p;
    int r_num = 0;
    std::vector<double> r(m->kn), s(m->kn);
    std::vector<char> set(m->kn, 0);
    for (int i = 0; i < m->kn; i++) {
        OKeyLine &k = m->kl[i];
        if (k.n_id < 0 || k.p_id < 0) continue;
        OKeyLine &kn = m->kl[k.n_id], &kp = m->kl[k.p_id];
        double d = kn.rho - kp.rho;
        if (d * d > kn.s_rho * kn.s_rho + kp.s_rho * kp.s_rho) continue;
        double alpha = (kn.m_m[0] * kp.m_m[0] + kn.m_m[1] * kp.m_m[1]) / (kn.n_m * kp.n_m);
        if (alpha - thresh < 0) continue;
        alpha = (alpha - thresh) / (1 - thresh);
        alpha /= fabs(kn.rho - kp.rho) / (kn.s_rho + kp.s_rho) + 1;
        double wr = 1 / (k.s_rho * k.s_rho), wrn = alpha / (kn.s_rho * kn.s_rho), wrp = alpha / (kp.s_rho * kp.s_rho);
        r[i] = (k.rho * wr + kn.rho * wrn + kp.rho * wrp) / (wr + wrn + wrp);
        s[i] = (k.s_rho * wr + kn.s_rho * wrn + kp.s_rho * wrp) / (wr + wrn + wrp);
        set[i] = 1;
        r_num++;
    }
    for (int i = 0; i < m->kn; i++)
        if (set[i]) {
            m->kl[i].rho = r[i];
            m->kl[i].s_rho = s[i];
        }
    return r_num;
}

// edge_tracker::UpdateInverseDepthKalman -> ...ARLU (edge_tracker.cpp:695-724, 954-1055)
void orc_ekf(void *p, const double *vel, double q_abs, double loc_unc) {
    OMap *m = (OMap *)p;
    const double zf = m->zfm;
    for (int i = 0; i < m->kn; i++) {
        OKeyLine &k = m->kl[i];
        if (k.m_id < 0) continue;
        double &rho = k.rho, &s_rho = k.s_rho;
        k.s_rho0 = s_rho;
        double qx = k.p_m[0], qy = k.p_m[1], q0x = k.p_m_0[0], q0y = k.p_m_0[1];
        double v_rho = s_rho * s_rho;
        double u_x = k.m_m0[0] / k.n_m0, u_y = k.m_m0[1] / k.n_m0;
        double Y = u_x * (qx - q0x) + u_y * (qy - q0y);
        double H = u_x * (vel[0] * zf - vel[2] * q0x) + u_y * (vel[1] * zf - vel[2] * q0y);
        double rho_p = 1 / (1.0 / rho + vel[2]);
        k.rho0 = rho_p;
        double F = 1 / (1 + rho * vel[2]);
        F = F * F;
        double p_p = F * v_rho * F + q_abs * q_abs;
        double e = Y - H * rho_p;
        double S = H * p_p * H + loc_unc * loc_unc;
        double K = p_p * H * (1 / S);
        rho = rho_p + (K * e);
        v_rho = (1 - K * H) * p_p;
        s_rho = sqrt(v_rho);
        if (rho < RHO_MIN) {
            s_rho += RHO_MIN - rho;
            rho = RHO_MIN;
        } else if (rho > RHO_MAX) {
            rho = RHO_MAX;
        } else if (std::isnan(rho) || std::isnan(s_rho) || std::isinf(rho) || std::isinf(s_rho)) {
            rho = RHO_INIT;
            s_rho = RHO_MAX;
        } else if (s_rho < 0) {
            rho = RHO_INIT;
            s_rho = RHO_MAX;
        }
    }
}

// edge_tracker::EstimateReScalingOpt (edge_tracker.cpp:1104-1140)
double orc_rescale(void *p, double *RKp, double s_rho_min, unsigned mnm, int re_escale) {
    OMap *m = (OMap *)p;
    if (m->kn <= 0) return 1;
    double Kp = 1;
    for (int iter = 0; iter < 5; iter++) {
        double rTr = 0, rTr0 = 0;
        for (int i = 0; i < m->kn; i++) {
            OKeyLine &k = m->kl[i];
            if ((unsigned)k.m_num < mnm || k.s_rho0 <= 0 || k.s_rho > s_rho_min) continue;
            rTr += k.rho * k.rho / (k.s_rho * k.s_rho + Kp * Kp * k.s_rho0 * k.s_rho0);
            rTr0 += k.rho0 * k.rho0 / (k.s_rho * k.s_rho + Kp * Kp * k.s_rho0 * k.s_rho0);
        }
        Kp = rTr0 > 0 ? sqrt(rTr / rTr0) : 1;
        *RKp = 1 / rTr0;
    }
    if (re_escale)
        for (int i = 0; i < m->kn; i++) {
            m->kl[i].rho /= Kp;
            m->kl[i].s_rho /= Kp;
        }
    return Kp;
}
void orc_so3_exp(const double *w, double *R) { so3_exp(w, R); }
void orc_set_frame_count(void *p, unsigned fc) { ((OMap *)p)->frame_count = fc; }

// =====================================================================================================
// IMU-mode rows (configuration 3)
// =====================================================================================================
// TooN::determinant of a 3x3 = determinant_gaussian_elimination (TooN/determinant.h:91-146): partial pivoting, the
// running product of the pivots, early exit on zero
static double det3_gauss(const double *Ain) {
    double A[3][3];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) A[i][j] = Ain[i * 3 + j];
    double det = 1;
    for (int i = 0; i < 3; i++) {
        int argmax = i;
        double maxval = fabs(A[i][i]);
        for (int ii = i + 1; ii < 3; ii++) {
            const double v = fabs(A[ii][i]);
            if (v > maxval) {
                maxval = v;
                argmax = ii;
            }
        }
        const double pivot = A[argmax][i];
        if (argmax != i) {
            det *= -1;
            for (int j = i; j < 3; j++) {
                const double t = A[i][j];
                A[i][j] = A[argmax][j];
                A[argmax][j] = t;
            }
        }
        det *= A[i][i];
        if (det == 0) return 0;
        for (int u = i + 1; u < 3; u++) {
            const double factor = A[u][i] / pivot;
            for (int j = i + 1; j < 3; j++) A[u][j] = A[u][j] - factor * A[i][j];
        }
    }
    return det;
}
// util::Matrix3x3Inv (include/UtilLib/toon_util.h:32-41): cofactors / TooN::determinant
static void mat3_inv(const double *A, double *B) {
    double t[9];
    t[0] = A[8] * A[4] - A[7] * A[5];
    t[1] = -(A[8] * A[1] - A[7] * A[2]);
    t[2] = A[5] * A[1] - A[4] * A[2];
    t[3] = -(A[8] * A[3] - A[6] * A[5]);
    t[4] = A[8] * A[0] - A[6] * A[2];
    t[5] = -(A[5] * A[0] - A[3] * A[2]);
    t[6] = A[7] * A[3] - A[6] * A[4];
    t[7] = -(A[7] * A[0] - A[6] * A[1]);
    t[8] = A[4] * A[0] - A[3] * A[1];
    const double det = det3_gauss(A);
    for (int i = 0; i < 9; i++) B[i] = t[i] / det;
}
// TooN fixed-size products: every element is a dot product accumulated from zero in index order
static void mat3_mul(const double *A, const double *B, double *C) {
    double t[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double r = 0;
            for (int k = 0; k < 3; k++) r += A[i * 3 + k] * B[k * 3 + j];
            t[i * 3 + j] = r;
        }
    for (int i = 0; i < 9; i++) C[i] = t[i];
}
static void mat3_vec(const double *A, const double *v, double *o) {
    double t[3];
    for (int i = 0; i < 3; i++) {
        double r = 0;
        for (int k = 0; k < 3; k++) r += A[i * 3 + k] * v[k];
        t[i] = r;
    }
    for (int i = 0; i < 3; i++) o[i] = t[i];
}

// global_tracker::TryVel<double> (global_tracker.cpp:829-934) with Calc_f_J (:174-220) and Test_f_k
// (global_tracker.h:89-104).  pn = the map whose field is searched, po = the keylines being projected; residuals is read
// and updated in place (K0 doubles).  Sequential sums in the reference's order.
double orc_try_vel(void *pn, void *po, const double *Vel, double match_thresh, double s_rho_min,
                   unsigned match_num_thresh, double *residuals, double reweigth_distance, float min_mod, double *JtJ,
                   double *JtF) {
    OMap *mn = (OMap *)pn, *mo = (OMap *)po;
    const double max_r = mn->max_r, zfm = mn->zfm;
    double score = 0, f;
    for (int i = 0; i < 9; i++) JtJ[i] = 0;
    for (int i = 0; i < 3; i++) JtF[i] = 0;
    double fi = 0;
    const unsigned mnt = match_num_thresh < mn->frame_count ? match_num_thresh : mn->frame_count;
    for (int ikl = 0; ikl < mo->kn; ikl++) {
        OKeyLine &kl = mo->kl[ikl];
        kl.m_id_f = -1;
        if (min_mod > 0 && kl.n_m < min_mod) continue;
        if (kl.s_rho > s_rho_min || (unsigned)kl.m_num < mnt) continue;
        double weight = 1;
        if (residuals[ikl] > reweigth_distance) weight = reweigth_distance / residuals[ikl];
        const double z_p = 1.0 / kl.rho + Vel[2];
        if (z_p <= 0) {
            f = (1 / (kl.s_rho)) * max_r * weight;
            score += f * f;
            continue;
        }
        const double rho_p = 1.0 / z_p;
        const double pjx = rho_p * (Vel[0] * zfm - Vel[2] * kl.p_m[0]) + kl.p_m[0];
        const double pjy = rho_p * (Vel[1] * zfm - Vel[2] * kl.p_m[1]) + kl.p_m[1];
        const double pix = pjx + mn->ppx, piy = pjy + mn->ppy;   // cam_mod.Hom2Img
        const int x = (int)(pix + 0.5), y = (int)(piy + 0.5);    // util::round2int_positive
        if (x < 1 || y < 1 || x >= mn->w - 1 || y >= mn->h - 1) {
            f = (1 / (kl.s_rho)) * max_r * weight;
            score += f * f;
            continue;
        }
        double df_dx, df_dy;
        {   // Calc_f_J<double>(y*w+x, df_dx, df_dy, kl, p_pji, max_r, match_thresh, mnum, fi)
            const int fidx = y * mn->w + x;
            bool hit = false;
            if (mn->fikl[fidx] >= 0) {
                const OKeyLine &fk = mn->kl[mn->fikl[fidx]];
                const double p_n2 = (kl.n_m * kl.n_m);                              // float product -> double
                const double p_esc = kl.m_m[0] * fk.m_m[0] + kl.m_m[1] * fk.m_m[1];  // float arithmetic -> double
                if (!(fabs(p_esc - p_n2) > match_thresh * p_n2)) {
                    const double dx = pix - fk.c_p[0], dy = piy - fk.c_p[1];
                    fi = (dx * fk.u_m[0] + dy * fk.u_m[1]);
                    df_dx = fk.u_m[0] / kl.s_rho;
                    df_dy = fk.u_m[1] / kl.s_rho;
                    kl.m_id_f = mn->fikl[fidx];
                    f = fi / kl.s_rho;
                    hit = true;
                }
            }
            if (!hit) {
                df_dx = 0;
                df_dy = 0;
                f = max_r / kl.s_rho;
            }
        }
        f *= weight;
        score += f * f;
        const double jx = rho_p * zfm * df_dx * weight;
        const double jy = rho_p * zfm * df_dy * weight;
        const double jz = -rho_p * (pjx * df_dx + pjy * df_dy) * weight;
        JtJ[0] += jx * jx;
        JtJ[4] += jy * jy;
        JtJ[8] += jz * jz;
        JtJ[1] += jx * jy;
        JtJ[2] += jx * jz;
        JtJ[5] += jy * jz;
        JtF[0] += jx * f;
        JtF[1] += jy * f;
        JtF[2] += jz * f;
        residuals[ikl] = fabs(fi);
    }
    JtJ[3] = JtJ[1];
    JtJ[6] = JtJ[2];
    JtJ[7] = JtJ[5];
    return score;
}

// global_tracker::Minimizer_V<double> (global_tracker.cpp:1036-1093)
double orc_minimizer_v(void *pn, void *po, double *Vel, double *RVel, double match_thresh, int iter_max, double s_rho_min,
                       unsigned match_num_thresh, double reweigth_distance, float min_mod) {
    OMap *mo = (OMap *)po;
    double JtJ[9], ApI[9], JtJnew[9], JtF[3], JtFnew[3], h[3], Vnew[3], inv[9];
    std::vector<double> residuals(mo->kn > 0 ? mo->kn : 1, 0.0);
    double F = orc_try_vel(pn, po, Vel, match_thresh, s_rho_min, match_num_thresh, residuals.data(), reweigth_distance,
                           min_mod, JtJ, JtF),
           Fnew;
    double v = 2, tau = 1e-3;
    double mx = JtJ[0];
    for (int i = 1; i < 9; i++)
        if (JtJ[i] > mx) mx = JtJ[i];
    double u = tau * mx;
    for (int lm_iter = 0; lm_iter < iter_max; lm_iter++) {
        for (int i = 0; i < 9; i++) ApI[i] = JtJ[i];
        for (int i = 0; i < 3; i++) ApI[i * 4] = JtJ[i * 4] + u;
        mat3_inv(ApI, inv);
        const double ng[3] = {-JtF[0], -JtF[1], -JtF[2]};
        mat3_vec(inv, ng, h);
        for (int i = 0; i < 3; i++) Vnew[i] = Vel[i] + h[i];
        Fnew = orc_try_vel(pn, po, Vnew, match_thresh, s_rho_min, match_num_thresh, residuals.data(), reweigth_distance,
                           min_mod, JtJnew, JtFnew);
        double den = 0;
        for (int i = 0; i < 3; i++) den += (0.5 * h[i]) * (u * h[i] - JtF[i]);
        const double gain = (F - Fnew) / den;
        if (gain > 0) {
            F = Fnew;
            for (int i = 0; i < 3; i++) Vel[i] = Vnew[i];
            for (int i = 0; i < 9; i++) JtJ[i] = JtJnew[i];
            for (int i = 0; i < 3; i++) JtF[i] = JtFnew[i];
            const double g = 1 - ((2 * gain - 1) * (2 * gain - 1) * (2 * gain - 1));
            u *= (0.33 > g ? 0.33 : g);
            v = 2;
        } else {
            u *= v;
            v *= 2;
        }
    }
    mat3_inv(JtJ, RVel);
    return F;
}

// edge_tracker::ExtRotVel(vel, Wx, Rx, X, LocUncert, HubReweigth) (edge_tracker.cpp:1207-1301).  The float / double mix of
// every expression follows the reference's declarations.  SVD<>::backsub / get_pinv are replaced by an elimination solve
// (the 6x6 system is well conditioned on tracked maps): X and Rx are compared to 1e-9, Wx = Phi^T Phi exactly.
int orc_ext_rot_vel(void *p, const double *vel, double *Wx, double *Rx, double *X, double LocUncert, double HubReweigth) {
    OMap *m = (OMap *)p;
    const double zf = m->zfm;
    int n = 0;
    for (int i = 0; i < m->kn; i++)
        if (m->kl[i].m_id >= 0) n++;
    std::vector<double> Phi((size_t)n * 6 + 6, 0.0), Y(n > 0 ? n : 1, 0.0);
    int j = 0;
    for (int i = 0; i < m->kn; i++) {
        const OKeyLine &k = m->kl[i];
        if (k.m_id < 0) continue;
        const float u_x = k.u_m[0], u_y = k.u_m[1];
        const double rho_t = 1 / (1 / k.rho + vel[2]);
        const float qt_x = k.p_m_0[0] + rho_t * (vel[0] * zf - vel[2] * k.p_m_0[0]);
        const float qt_y = k.p_m_0[1] + rho_t * (vel[1] * zf - vel[2] * k.p_m_0[1]);
        const double s_rho = k.s_rho;
        const float q_x = k.p_m[0], q_y = k.p_m[1];
        double *row = &Phi[(size_t)j * 6];
        row[0] = u_x * rho_t * zf;
        row[1] = u_y * rho_t * zf;
        row[2] = u_x * (-rho_t * q_x) + u_y * (-rho_t * q_y);
        row[3] = -u_x * q_x * q_y / zf - u_y * (zf + q_y * q_y / zf);
        row[4] = +u_y * q_x * q_y / zf + u_x * (zf + q_x * q_x / zf);
        row[5] = -u_x * q_y + u_y * q_x;
        Y[j] = u_x * (k.p_m[0] - qt_x) + u_y * (k.p_m[1] - qt_y);
        const float dqvel = u_x * (vel[0] * zf - vel[2] * k.p_m_0[0]) + u_y * (vel[1] * zf - vel[2] * k.p_m_0[1]);
        const float s_y = sqrt(s_rho * s_rho * dqvel * dqvel + LocUncert * LocUncert);
        double weigth = 1;
        if (fabs(Y[j]) > HubReweigth) weigth = fabs(Y[j]) / HubReweigth;
        for (int c = 0; c < 6; c++) row[c] /= s_y * weigth;
        Y[j] /= s_y * weigth;
        j++;
    }
    double JtJ[36], JtF[6];
    for (int a = 0; a < 6; a++) {   // Phi.T()*Phi, Phi.T()*Y: dot products over the rows, from zero, in row order
        for (int b = 0; b < 6; b++) {
            double r = 0;
            for (int q = 0; q < n; q++) r += Phi[(size_t)q * 6 + a] * Phi[(size_t)q * 6 + b];
            JtJ[a * 6 + b] = r;
        }
        double r = 0;
        for (int q = 0; q < n; q++) r += Phi[(size_t)q * 6 + a] * Y[q];
        JtF[a] = r;
    }
    solve6(JtJ, JtF, X);
    for (int c = 0; c < 6; c++) {
        double e[6] = {0, 0, 0, 0, 0, 0}, col[6];
        e[c] = 1;
        solve6(JtJ, e, col);
        for (int r = 0; r < 6; r++) Rx[r * 6 + c] = col[r];
    }
    for (int i = 0; i < 36; i++) Wx[i] = JtJ[i];
    for (int i = 0; i < 36; i++)
        if (Rx[i] != Rx[i]) return 0;
    for (int i = 0; i < 6; i++)
        if (X[i] != X[i]) return 0;
    return 1;
}

// edge_tracker::BiasCorrect (edge_tracker.cpp:1308-1338); all arguments in/out like the reference
void orc_bias_correct(double *X, double *Wx, double *Gb, double *Wb, const double *Rg, const double *Rb) {
    double Wg[9], t3[9], iWgWb[9];
    mat3_inv(Rg, Wg);
    mat3_inv(Wb, t3);
    for (int i = 0; i < 9; i++) t3[i] = t3[i] + Rb[i];
    mat3_inv(t3, Wb);   // Wb = Matrix3x3Inv(Matrix3x3Inv(Wb) + Rb)
    double Wxb[36];
    for (int i = 0; i < 36; i++) Wxb[i] = Wx[i];
    for (int i = 0; i < 9; i++) t3[i] = Wg[i] + Wb[i];
    mat3_inv(t3, iWgWb);
    double A[9], B[9];
    mat3_mul(iWgWb, Wg, A);                       // iWgWb*Wg
    for (int i = 0; i < 9; i++) A[i] = ((i % 4 == 0) ? 1.0 : 0.0) - A[i];   // Identity - iWgWb*Wg
    mat3_mul(Wg, A, B);                           // Wg*(Identity - iWgWb*Wg)
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) Wxb[(3 + r) * 6 + 3 + c] += B[r * 3 + c];
    double X1[6];
    for (int r = 0; r < 6; r++) {                 // X1 = Wx*X
        double acc = 0;
        for (int c = 0; c < 6; c++) acc += Wx[r * 6 + c] * X[c];
        X1[r] = acc;
    }
    double v3[3];
    mat3_mul(Wg, iWgWb, A);                       // Wg*iWgWb*Wb*Gb, left to right
    mat3_mul(A, Wb, B);
    mat3_vec(B, Gb, v3);
    for (int r = 0; r < 3; r++) X1[3 + r] += v3[r];
    Chol ch;
    chol_compute(Wxb, &ch);
    double inv[36];
    chol_inverse(&ch, inv);
    for (int r = 0; r < 6; r++) {                 // X = Cholesky<6>(Wxb).get_inverse()*X1
        double acc = 0;
        for (int c = 0; c < 6; c++) acc += inv[r * 6 + c] * X1[c];
        X[r] = acc;
    }
    double a3[3], b3[3];
    mat3_vec(Wg, X + 3, a3);                      // Gb = iWgWb*(Wg*X.slice<3,3>() + Wb*Gb)
    mat3_vec(Wb, Gb, b3);
    for (int r = 0; r < 3; r++) a3[r] = a3[r] + b3[r];
    mat3_vec(iWgWb, a3, Gb);
    for (int i = 0; i < 9; i++) Wb[i] = Wg[i] + Wb[i];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) Wx[(3 + r) * 6 + 3 + c] += Wg[r * 3 + c];
}

// =====================================================================================================
// image_undistort: constructor map (src/VideoLib/image_undistort.cpp:29-94) with
// cam_model::distortHom2Hom (include/UtilLib/cam_model.h:77-88), then undistort<true> on RGB24 through
// biInterp (include/VideoLib/image_undistort.h:63-78).  kc = {Kc2, Kc4, Kc6, P1, P2}.
// =====================================================================================================
static bool und_inx_valid(float fx, float fy, int w, int h) {
    // Image::isInxValid takes `const uint&`: the float argument converts to unsigned (x86-64 cvttss2si to 64 bit, low
    // half kept), so negative coordinates become huge and fail the bound
    const unsigned x = (unsigned)(long long)fx, y = (unsigned)(long long)fy;
    return x < (unsigned)w && y < (unsigned)h;
}
void orc_undistort_rgb(int w, int h, float ppx, float ppy, float zfx, float zfy, const double *kc,
                       const unsigned char *in, unsigned char *out) {
    const double zfm = (double)((zfx + zfy) / 2);   // cam_model: zfm((focal_dist.x+focal_dist.y)/2) in float
    const double Kc2 = kc[0], Kc4 = kc[1], Kc6 = kc[2], P1 = kc[3], P2 = kc[4];
    const float i_mult = (float)(1 << 16);
    for (int x = 0; x < w; x++)
        for (int y = 0; y < h; y++) {
            float qx = (float)x - ppx, qy = (float)y - ppy;   // cam.Img2Hom(Point2D<float>(x,y))
            {
                const double xp = qx / zfm, yp = qy / zfm;
                const double r2 = xp * xp + yp * yp;
                const double xpp = xp * (1 + r2 * (Kc2 + r2 * (Kc4 + r2 * Kc6))) + 2 * P1 * xp * yp + P2 * (r2 + 2 * xp * xp);
                const double ypp = yp * (1 + r2 * (Kc2 + r2 * (Kc4 + r2 * Kc6))) + P1 * (r2 + 2 * yp * yp) + 2 * P2 * xp * yp;
                qx = xpp * zfx;
                qy = ypp * zfy;
            }
            const float idx = qx + ppx, idy = qy + ppy;       // cam.Hom2Img(qd)
            const float p00x = floor(idx), p00y = floor(idy), p11x = floor(idx) + 1, p11y = floor(idy) + 1;
            const float tx[4] = {p00x, p11x, p00x, p11x}, ty[4] = {p00y, p00y, p11y, p11y};
            const float tw[4] = {(p11x - idx) * (p11y - idy), (idx - p00x) * (p11y - idy), (p11x - idx) * (idy - p00y),
                                 (idx - p00x) * (idy - p00y)};
            int num = 0, inx[4], iw[4];
            float wgt[4];
            for (int k = 0; k < 4; k++)
                if (und_inx_valid(tx[k], ty[k], w, h)) {
                    wgt[num] = tw[k];
                    inx[num] = index_rc(tx[k], ty[k], w, h);
                    num++;
                }
            if (num > 0) {
                float sum_w = 0;
                for (int i = 0; i < num; i++) sum_w += wgt[i];
                for (int i = 0; i < num; i++) {
                    wgt[i] /= sum_w;
                    iw[i] = wgt[i] * i_mult;
                }
            }
            int r = 0, g = 0, b = 0;                          // biInterp(Image<RGB24Pixel>&, inx)
            for (int i = 0; i < num; i++) {
                const unsigned char *p = in + 3 * (size_t)inx[i];
                r += iw[i] * p[0];
                g += iw[i] * p[1];
                b += iw[i] * p[2];
            }
            unsigned char *o = out + 3 * ((size_t)y * w + x);   // umap(x,y) -> out[inx]
            o[0] = (unsigned char)(r >> 16);
            o[1] = (unsigned char)(g >> 16);
            o[2] = (unsigned char)(b >> 16);
        }
}
}  // extern "C"
