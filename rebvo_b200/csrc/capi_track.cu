// capi_track.cu -- stage-level C ABI of the tracker / mapper (include/rebvo_b200.h): thin synchronous
// wrappers that stage scalar arguments in device memory, enqueue the kernels of tracker.cu and read the
// scalar results back.  The per-frame flow (pipeline.cu) calls the *_enqueue functions directly and never
// synchronises between stages.
#include <math.h>

#include "common.cuh"
#include "tracker.cuh"

int rb_resolve_res_enqueue(rb_ctx *c, rb_map *fmap, rb_map *old, int buf);
int rb_try_vel_rot_enqueue(rb_ctx *c, rb_map *fmap, rb_map *old, const double *X_dev, int reweight, int procjf,
                           double match_thresh, double s_rho_min, unsigned int mnt, unsigned int fc,
                           double k_huber);

static double *args_dev(rb_ctx *c) { return (double *)((char *)c->dev_small + RB_DS_ARGS); }
static double *args_pin(rb_ctx *c) { return (double *)((char *)c->pinned + RB_DS_ARGS); }

static int stage_in(rb_ctx *c, const double *src, int n, int off) {
    memcpy(args_pin(c) + off, src, sizeof(double) * n);
    RB_CUDA(cudaMemcpyAsync(args_dev(c) + off, args_pin(c) + off, sizeof(double) * n, cudaMemcpyHostToDevice,
                            c->stream));
    return RB_OK;
}

extern "C" int rb_map_quantile(rb_map *m, double s_rho_min, double s_rho_max, double percentile, int nbins,
                               double *out) {
    if (!m) return RB_ERR_ARG;
    RB_ENTER(m->c);
    int r = rb_quantile_enqueue(m->c, m, s_rho_min, s_rho_max, percentile, nbins);
    if (r) return r;
    MapState s;
    if ((r = rb_read_map_state(m, &s))) return r;
    if (out) *out = s.s_rho_q;
    return RB_OK;
}

extern "C" int rb_map_build_field(rb_map *m, int radius, float min_mod) {
    if (!m) return RB_ERR_ARG;
    RB_ENTER(m->c);
    int r = rb_build_field_enqueue(m->c, m, radius, min_mod, false);
    if (r) return r;
    return rb_ctx_sync(m->c);
}

__global__ void k_field_unpack(const unsigned long long *__restrict__ f, int2 *__restrict__ out, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long k = f[i];
    int2 o;
    if (k == ~0ull) {
        o.x = 0;
        o.y = -1;
    } else {
        o.x = (int)(k >> 32);
        o.y = (int)(0xFFFFFFFFu - (unsigned int)(k & 0xFFFFFFFFull));
    }
    out[i] = o;
}

extern "C" int rb_map_get_field(rb_map *m, int32_t *out) {
    if (!m) return RB_ERR_ARG;
    RB_ENTER(m->c);
    rb_ctx *c = m->c;
    int2 *tmp = nullptr;
    RB_CUDA(cudaMalloc(&tmp, sizeof(int2) * (size_t)c->N));
    k_field_unpack<<<(unsigned)((c->N + 255) / 256), 256, 0, c->stream>>>(m->field, tmp, (size_t)c->N);
    c->launches++;
    cudaError_t e = cudaMemcpyAsync(out, tmp, sizeof(int2) * (size_t)c->N, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(tmp);
    if (e != cudaSuccess) {
        snprintf(c->err, sizeof(c->err), "get_field: %s", cudaGetErrorString(e));
        return RB_ERR_CUDA;
    }
    return RB_OK;
}

extern "C" int rb_map_set_frame_count(rb_map *m, uint32_t fc) {
    if (!m) return RB_ERR_ARG;
    RB_ENTER(m->c);
    rb_ctx *c = m->c;
    RB_CUDA(cudaMemcpyAsync(&m->st->frame_count, &fc, sizeof(fc), cudaMemcpyHostToDevice, c->stream));
    RB_CUDA(cudaStreamSynchronize(c->stream));
    return RB_OK;
}

extern "C" int rb_try_vel_rot(rb_map *fmap, rb_map *old, const double X[6], int reweight, int procjf,
                              double match_thresh, double s_rho_min, uint32_t match_num_thresh, double k_huber,
                              const double *res_in, double *res_out, double JtJ[36], double JtF[6], double *score) {
    if (!fmap || !old) return RB_ERR_ARG;
    RB_ENTER(fmap->c);
    rb_ctx *c = fmap->c;
    MapState so, sf;
    int r;
    if ((r = rb_read_map_state(old, &so))) return r;
    if ((r = rb_read_map_state(fmap, &sf))) return r;
    if ((r = stage_in(c, X, 6, 0))) return r;
    if (res_in && so.kn > 0)
        RB_CUDA(cudaMemcpyAsync(fmap->res[0], res_in, sizeof(double) * so.kn, cudaMemcpyHostToDevice, c->stream));
    else
        RB_CUDA(cudaMemsetAsync(fmap->res[0], 0, sizeof(double) * (size_t)c->kcap, c->stream));
    if ((r = rb_try_vel_rot_enqueue(c, fmap, old, args_dev(c), reweight, procjf, match_thresh, s_rho_min,
                                    match_num_thresh, sf.frame_count, k_huber)))
        return r;
    LMState *lmh = (LMState *)((char *)c->pinned + 4096);
    RB_CUDA(cudaMemcpyAsync(lmh, &fmap->ts->lm, sizeof(LMState), cudaMemcpyDeviceToHost, c->stream));
    if (res_out && so.kn > 0 && (r = rb_resolve_res_enqueue(c, fmap, old, 1))) return r;
    if (res_out && so.kn > 0)
        RB_CUDA(cudaMemcpyAsync(res_out, fmap->res[1], sizeof(double) * so.kn, cudaMemcpyDeviceToHost, c->stream));
    RB_CUDA(cudaStreamSynchronize(c->stream));
    if (JtJ) memcpy(JtJ, lmh->JtJn, sizeof(double) * 36);
    if (JtF) memcpy(JtF, lmh->JtFn, sizeof(double) * 6);
    if (score) *score = lmh->last_score;
    return RB_OK;
}

extern "C" int rb_minimizer_rv(rb_map *fmap, rb_map *old, double V[3], double W[3], double RVel[9], double RW0[9],
                               double match_thresh, int iter_max, int init_type, double reweight_distance,
                               double *rel_error, double *rel_error_score, double max_s_rho,
                               uint32_t match_num_thresh, int init_iter, double W_X[36], double *score) {
    if (!fmap || !old) return RB_ERR_ARG;
    RB_ENTER(fmap->c);
    rb_ctx *c = fmap->c;
    int r;
    double vw[6] = {V[0], V[1], V[2], W[0], W[1], W[2]};
    if ((r = stage_in(c, vw, 6, 0))) return r;
    rb_minimizer_args a;
    a.match_thresh = match_thresh;
    a.iter_max = iter_max;
    a.init_type = init_type;
    a.init_iter = init_iter;
    a.reweight_distance = reweight_distance;
    a.match_num_thresh = match_num_thresh;
    if ((r = rb_minimizer_enqueue(c, fmap, old, args_dev(c), &a, max_s_rho, false, 0, true))) return r;
    LMState *lmh = (LMState *)((char *)c->pinned + 4096);
    RB_CUDA(cudaMemcpyAsync(lmh, &fmap->ts->lm, sizeof(LMState), cudaMemcpyDeviceToHost, c->stream));
    if ((r = rb_minimizer_check_abort(c, fmap))) return r;   // (synchronises)
    if (lmh->no_keylines) {   // "if(klist.KNum()<=0) return 0;": every output keeps the caller's value
        if (score) *score = 0;
        return RB_OK;
    }
    memcpy(V, lmh->Vel, sizeof(double) * 3);
    memcpy(W, lmh->W0, sizeof(double) * 3);
    if (RVel) memcpy(RVel, lmh->RVel, sizeof(double) * 9);
    if (RW0) memcpy(RW0, lmh->RW0, sizeof(double) * 9);
    if (W_X) memcpy(W_X, lmh->W_X, sizeof(double) * 36);
    if (rel_error) *rel_error = lmh->rel_error;
    if (rel_error_score) *rel_error_score = lmh->rel_error_score;
    if (score) *score = lmh->score;
    return RB_OK;
}

extern "C" int rb_forward_match(rb_map *old, rb_map *neu, int *nmatch) {
    if (!old || !neu) return RB_ERR_ARG;
    RB_ENTER(old->c);
    int r = rb_forward_match_enqueue(old->c, old, neu);
    if (r) return r;
    MapState s;
    if ((r = rb_read_map_state(neu, &s))) return r;
    if (nmatch) *nmatch = s.fwd_match;
    return RB_OK;
}

extern "C" int rb_map_rotate_keylines(rb_map *m, const double R[9]) {
    if (!m) return RB_ERR_ARG;
    RB_ENTER(m->c);
    rb_ctx *c = m->c;
    int r;
    if ((r = stage_in(c, R, 9, 16))) return r;
    if ((r = rb_rotate_enqueue(c, m, args_dev(c) + 16))) return r;
    return rb_ctx_sync(c);
}

extern "C" int rb_directed_matching(rb_map *neu, rb_map *old, const double Vel[3], const double RVel[9],
                                    const double BackRot[9], double min_thr_mod, double min_thr_ang,
                                    double max_radius, double loc_uncertainty, int *nmatch) {
    if (!neu || !old) return RB_ERR_ARG;
    RB_ENTER(neu->c);
    rb_ctx *c = neu->c;
    // Vel=BackRot*Vel; RVel=BackRot*RVel*BackRot.T()  (edge_tracker.cpp:324-325), TooN dot order
    DMatchArgs a;
    for (int i = 0; i < 3; i++) {
        double s = 0;
        for (int k = 0; k < 3; k++) s += BackRot[i * 3 + k] * Vel[k];
        a.Vel[i] = s;
    }
    double t[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double s = 0;
            for (int k = 0; k < 3; k++) s += BackRot[i * 3 + k] * RVel[k * 3 + j];
            t[i * 3 + j] = s;
        }
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double s = 0;
            for (int k = 0; k < 3; k++) s += t[i * 3 + k] * BackRot[j * 3 + k];
            a.RVel[i * 3 + j] = s;
        }
    memcpy(a.BackRot, BackRot, sizeof(double) * 9);
    int r;
    if ((r = stage_in(c, (const double *)&a, sizeof(a) / sizeof(double), 32))) return r;
    if ((r = rb_directed_matching_enqueue(c, neu, old, (const DMatchArgs *)(args_dev(c) + 32), min_thr_mod,
                                          min_thr_ang, max_radius, loc_uncertainty, nullptr)))
        return r;
    MapState s;
    if ((r = rb_read_map_state(neu, &s))) return r;
    if (nmatch) *nmatch = s.nmatch;
    return RB_OK;
}

extern "C" int rb_map_regularize(rb_map *m, double thresh, int *r_num) {
    if (!m) return RB_ERR_ARG;
    RB_ENTER(m->c);
    int r = rb_regularize_enqueue(m->c, m, thresh, nullptr);
    if (r) return r;
    MapState s;
    if ((r = rb_read_map_state(m, &s))) return r;
    if (r_num) *r_num = s.reg_num;
    return RB_OK;
}

extern "C" int rb_map_ekf_update(rb_map *m, const double vel[3], double reshape_q_abs, double loc_uncertainty) {
    if (!m) return RB_ERR_ARG;
    RB_ENTER(m->c);
    rb_ctx *c = m->c;
    int r;
    if ((r = stage_in(c, vel, 3, 64))) return r;
    if ((r = rb_ekf_enqueue(c, m, args_dev(c) + 64, reshape_q_abs, loc_uncertainty, nullptr))) return r;
    return rb_ctx_sync(c);
}

extern "C" int rb_map_rescale_opt(rb_map *m, double s_rho_min, uint32_t match_num_min, int re_escale, double *Kp,
                                  double *RKp) {
    if (!m) return RB_ERR_ARG;
    RB_ENTER(m->c);
    int r = rb_rescale_enqueue(m->c, m, s_rho_min, match_num_min, re_escale, nullptr);
    if (r) return r;
    MapState s;
    if ((r = rb_read_map_state(m, &s))) return r;
    if (Kp) *Kp = s.Kp;
    if (RKp) *RKp = s.RKp;
    return RB_OK;
}
