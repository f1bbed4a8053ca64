// hostmath.cu -- C exports of the host/device algebra in lm.cuh so that the CPU test-suite can check it
// (SO3 exp/ln, LDL^T solve, symmetric pseudo-inverse solve) against the reference's TooN results without a GPU.
#include "lm.cuh"

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
}
