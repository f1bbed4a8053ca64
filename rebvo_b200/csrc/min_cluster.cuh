// min_cluster.cuh -- global_tracker::Minimizer_RV (src/mtracklib/global_tracker.cpp:578-819) in ONE thread-block
// cluster.  Included by tracker.cu after the per-keyline pieces (tvr_body, lm_*).
//
// A minimisation is ~12 TryVelRot evaluations, each a sum over all old keylines followed by a 6x6 LM step whose result
// is the next evaluation's pose: a chain of grid-wide reductions.  Exchanging through L2 between ~65 independent
// blocks cost two L2 round trips + a 65 x 59-slot gather per evaluation (2/3 of the kernel).  Here the whole edge map
// lives in one cluster of MC_C CTAs (guaranteed co-resident by the cluster launch, so nothing can dead-lock when
// several pipelines share a GPU):
//   * CTA r owns the contiguous keylines [r*kpc, (r+1)*kpc); their pose-independent operands (back-projected point,
//     s_rho) and the three residual buffers Res0/Res1/Rest stay in shared memory for the whole minimisation;
//   * per evaluation every CTA reduces its 28 sums + stale-fi summary and sends them to EVERY CTA of the cluster with
//     st.async (distributed shared memory, completion counted on the receiver's mbarrier: no cluster-wide barrier,
//     no fence, L1 stays valid); every CTA adds the MC_C contributions in rank order and runs the LM step itself --
//     all CTAs compute bit-identical poses, so there is no request broadcast at all;
//   * the zero-initialised and the prior-initialised tries of init type 2 (:644-751) are independent of each other
//     (no re-weighting, different output buffers Rest / ResidualNew): both poses are evaluated in the same round and
//     their LM steps run on two warps side by side, so 2*(init_iter+1) dependent rounds become init_iter+1.
// Sums: per thread in keyline order, transposing warp butterfly, warps in order, ranks in order -- fixed, deterministic,
// not the reference's pairwise tree (tests: rel 1e-10 on JtJ/JtF, V/W abs 1e-9).
#pragma once

#define MC_C 16                      // CTAs per cluster (non-portable size, one CTA per SM)
#define MC_T 512
#define MC_NW (MC_T / 32)
#define MC_MAXJ 7                    // keylines per thread: kcap <= MC_C * MC_T * MC_MAXJ = 57344 >= KEYLINE_MAX
#define MC_NV (MC_MAXJ * MC_NW)      // "virtual warps" of a CTA (32 consecutive keylines each)
#define MC_PW 30                     // doubles per pose and CTA in the exchange: 28 sums, has-a-match, last matched fi
#define MC_XW 64
#define MC_SPIN_LIMIT (1ll << 29)    // ~0.27 s: a broken exchange aborts with NaN results instead of hanging the device
#define MC_BYTES_PER_KL 57           // x0,y0,z0,s_rho, 3 residual buffers (double) + 1 flag byte

#ifdef RB_TVR_PROF   // stamps: [CTA][round (15 = kernel level)][8]
#define MC_STAMP(e, k) do { if (threadIdx.x == 0) g_tvr_prof[(blockIdx.x * 16 + (e)) * 8 + (k)] = clock64(); } while (0)
#else
#define MC_STAMP(e, k) do { } while (0)
#endif

struct McPlan {
    int n, merge_round;              // merge_round: the round after which the better init try is picked (-1: none)
    unsigned char sa[MIN_MAX_EVALS]; // zero-init try's step of the round (STEP_NONE when the round has one pose)
    unsigned char sb[MIN_MAX_EVALS]; // prior-init try / main loop step
};

struct __align__(16) McSmem {
    double gather[2][MC_C][MC_XW];   // [round parity][source rank][value]
    double part[MC_NW][32];
    double xout[MC_XW];
    double tot[2][28];
    double req[2][16];               // per pose: R[9] V[3] RotM[4]
    int req_res[2][2];               // per pose: res_in, res_out
    double wcarry[3][MC_NV];         // per residual buffer: the stale fi that leading misses of a virtual warp inherit
    double vw_last[2][MC_NV];
    int vw_has[2][MC_NV];
    unsigned long long mbar[2];
    LMState lm, lmz;                 // main / prior-init chain, zero-init chain
    int abort;
};

__device__ __forceinline__ unsigned int mc_smem_u32(const void *p) { return (unsigned int)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned int mc_mapa(unsigned int addr, unsigned int rank) {
    unsigned int r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ unsigned int mc_cluster_rank() {
    unsigned int r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void mc_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mc_st_async_v2(unsigned int raddr, double a, double b, unsigned int rmbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f64 [%0], {%1, %2}, [%3];" ::"r"(raddr),
                 "d"(a), "d"(b), "r"(rmbar)
                 : "memory");
}
__device__ __forceinline__ void mc_st_remote_v2(unsigned int raddr, double a, double b) {
    asm volatile("st.shared::cluster.v2.f64 [%0], {%1, %2};" ::"r"(raddr), "d"(a), "d"(b) : "memory");
}
__device__ __forceinline__ void mc_mbar_init(unsigned long long *bar, unsigned int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mc_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mc_mbar_expect_tx(unsigned long long *bar, unsigned int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mc_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mc_mbar_try_wait(unsigned long long *bar, unsigned int parity) {
    unsigned int ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(mc_smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

struct McView {          // this CTA's slice of the edge map in dynamic shared memory
    double *x0, *y0, *z0, *s_rho;
    double *res[3];
    unsigned char *flag;  // 1: m_num < min(MatchNumThresh, FrameCount)
    int base, cnt, J;     // first keyline, keylines of this CTA, iterations per thread (uniform over the cluster)
};

// one TryVelRot evaluation of pose slot p over this CTA's keylines; leaves the CTA's 28 sums and stale-fi summary in
// sm.xout[p * MC_PW ..]
template <bool RW, bool PJ>
__device__ __forceinline__ void mc_eval_pose(McSmem &sm, const McView &v, int p, const KLSoA &old, const TvrConst &tc,
                                             const CamC &cam, const unsigned long long *__restrict__ field,
                                             const float4 *__restrict__ fpack, bool write_mid, int tid, int lane,
                                             int wid) {
    const double *sR = sm.req[p], *sV = sR + 9, *sRM = sR + 12;
    const int res_in = sm.req_res[p][0], res_out = sm.req_res[p][1];
    const bool has_rin = RW && res_in >= 0;
    const double *rin = has_rin ? v.res[res_in] : nullptr;
    double *rout = v.res[res_out];
    double acc[28];
#pragma unroll
    for (int k = 0; k < 28; k++) acc[k] = 0;
    for (int j = 0; j < v.J; j++) {
        const int li = j * MC_T + tid, i = v.base + li;
        const int vw = j * MC_NW + wid;
        const bool active = li < v.cnt;
        bool matched = false, need = false, wrote = false;
        double fi_own = 0, r_w = 0;
        if (active) {
            KlOp o;
            o.x0 = v.x0[li];
            o.y0 = v.y0[li];
            o.z0 = v.z0[li];
            o.s_rho = v.s_rho[li];
            o.m = __ldg(&old.m_m[i]);
            o.n_m = __ldg(&old.n_m[i]);
            o.m_num = v.flag[li] ? 0 : 0x7fffffff;   // the m_num part of the skip test (:356) does not change between evaluations
            double r_prev = 0.0;
            if (has_rin) {
                r_prev = rin[li];
                if ((unsigned long long)__double_as_longlong(r_prev) == RES_SENTINEL) r_prev = sm.wcarry[res_in][vw];
            }
            double pr[28];
            tvr_body<RW, PJ>(o, has_rin, r_prev, sR, sV, sRM, tc, cam, field, fpack, nullptr, write_mid ? old.m_id_f : nullptr,
                             i, pr, matched, need, fi_own, wrote, r_w);
            if (PJ) {
#pragma unroll
                for (int k = 0; k < 28; k++) acc[k] += pr[k];
            } else {
                acc[27] += pr[27];
            }
        }
        // "DResidualNew[ikl]=fi" keeps the fi of the last matched keyline before ikl (:341,399-408): in-warp scan here,
        // earlier warps / CTAs through wcarry once the round's exchange is complete
        const unsigned int bal = __ballot_sync(0xffffffffu, matched);
        const unsigned int lower = bal & ((1u << lane) - 1u);
        const double prev_fi = __shfl_sync(0xffffffffu, fi_own, lower ? 31 - __clz(lower) : 0);
        const double wl = __shfl_sync(0xffffffffu, fi_own, bal ? 31 - __clz(bal) : 0);
        if (lane == 0) {
            sm.vw_has[p][vw] = bal != 0;
            sm.vw_last[p][vw] = wl;
        }
        if (active) {
            if (matched) rout[li] = fi_own;
            else if (need) rout[li] = lower ? prev_fi : __longlong_as_double((long long)RES_SENTINEL);
            else if (wrote) rout[li] = r_w;
        }
    }
    // block sums in a fixed order
    __syncthreads();   // previous user of sm.part / sm.vw_* readers are done
    if (PJ) {
        const double w = warp_transpose_sum28(acc, lane);
        sm.part[wid][lane] = w;
    } else {
        double s = acc[27];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) sm.part[wid][27] = s;
    }
    __syncthreads();
    if (tid < 28) {
        double t = 0;
        if (PJ || tid == 27) {
            t = sm.part[0][tid];
#pragma unroll
            for (int ww = 1; ww < MC_NW; ww++) t += sm.part[ww][tid];
        }
        sm.xout[p * MC_PW + tid] = t;
    } else if (tid == 32) {
        int has = 0;
        double lastv = 0;
        for (int q = v.J * MC_NW - 1; q >= 0; q--)
            if (sm.vw_has[p][q]) {
                has = 1;
                lastv = sm.vw_last[p][q];
                break;
            }
        sm.xout[p * MC_PW + 28] = has ? 1.0 : 0.0;
        sm.xout[p * MC_PW + 29] = lastv;
    }
}

// LM steps of the two init tries evaluated side by side (global_tracker.cpp:651-683 and :700-732)
__device__ __forceinline__ void mc_lm_step_zero(LMState &z, int step) {
    switch (step) {
        case STEP_INIT_FIRST_ZERO:
            lm_take_first(z);
            if (z.init_iter > 0) {
                lm_solve(z, true);
                lm_request(z, z.Xnew, -1, z.iRt);
            }
            break;
        case STEP_INIT_ITER_ZERO:
            lm_update(z, true);
            lm_solve(z, true);
            lm_request(z, z.Xnew, -1, z.iRt);
            break;
        case STEP_INIT_LAST_ZERO:
            lm_update(z, false);
            break;
        default:
            break;
    }
}
__device__ __forceinline__ void mc_lm_step_main(LMState &s, int step, MapState *fst) {
    switch (step) {
        case STEP_INIT_FIRST_PRIOR:
            lm_take_first(s);
            s.v = 2;
            if (s.init_iter > 0) {
                lm_solve(s, true);
                lm_request(s, s.Xnew, -1, s.iRN);
            }
            break;
        case STEP_INIT_ITER_PRIOR:
            lm_update(s, true);
            lm_solve(s, true);
            lm_request(s, s.Xnew, -1, s.iRN);
            break;
        case STEP_INIT_LAST_PRIOR:
            lm_update(s, false);
            break;
        default:
            lm_step(s, step, fst);
            break;
    }
}
// "Save the scores in temporals" (:686-691) + "Check for the lowest score" (:734-747) once both tries are done
__device__ __forceinline__ void mc_lm_merge(LMState &s, const LMState &z) {
    for (int i = 0; i < 6; i++) s.Xt[i] = z.X[i];
    s.Ft = z.F;
    s.F0t = z.F0;
    s.ut = z.u;
    s.vt = z.v;
    s.eff_steps_t = z.eff_steps;
    lm_after_prior_pass(s);
}

template <int XCHG>   // 1: st.async + mbarrier complete_tx; 0: plain DSMEM stores + barrier.cluster
__global__ void __cluster_dims__(MC_C, 1, 1) __launch_bounds__(MC_T, 1)
    k_minimizer_cluster(KLSoA old, const MapState *__restrict__ old_st, const unsigned long long *__restrict__ field,
                        const float4 *__restrict__ fpack, MapState *f_st, LMState *lm_out, int *abort_out, CamC cam,
                        McPlan plan, MinSetup su, FrameState *post_fs, int kpc_cap) {
    MC_STAMP(15, 0);
    pdl_wait();
    pdl_launch();
    MC_STAMP(15, 1);
    extern __shared__ __align__(16) unsigned char mc_dyn[];
    __shared__ McSmem sm;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int rank = (int)mc_cluster_rank();
    const int K0 = old_st->kn;
    McView v;
    {
        double *d = reinterpret_cast<double *>(mc_dyn);
        v.x0 = d;
        v.y0 = d + kpc_cap;
        v.z0 = d + 2 * kpc_cap;
        v.s_rho = d + 3 * kpc_cap;
        v.res[0] = d + 4 * kpc_cap;
        v.res[1] = d + 5 * kpc_cap;
        v.res[2] = d + 6 * kpc_cap;
        v.flag = reinterpret_cast<unsigned char *>(d + 7 * kpc_cap);
        int kpc = (K0 + MC_C - 1) / MC_C;
        kpc = (kpc + 31) & ~31;               // whole virtual warps
        if (kpc > kpc_cap) kpc = kpc_cap;     // (the host only launches this kernel when MC_C * kpc_cap >= capacity)
        v.base = rank * kpc;
        v.cnt = K0 - v.base;
        v.cnt = v.cnt < 0 ? 0 : (v.cnt > kpc ? kpc : v.cnt);
        v.J = (kpc + MC_T - 1) / MC_T;
    }
    TvrConst tc;
    tc.max_r = su.max_r;
    tc.match_thresh = su.a.match_thresh;
    tc.k_huber = su.a.reweight_distance;
    tc.s_rho_min = su.s_rho_from_state ? old_st->s_rho_q : su.max_s_rho;
    {
        const unsigned int fc = su.fc_from_state ? f_st->frame_count : su.frame_count;
        tc.mnt = su.a.match_num_thresh < fc ? su.a.match_num_thresh : fc;
    }
    // ---- prologue: operands -> shared memory, LM state, barriers -----------------------------------------------
    for (int li = tid; li < v.cnt; li += MC_T) {
        const KlOp o = load_klop(old, v.base + li, cam);
        v.x0[li] = o.x0;
        v.y0[li] = o.y0;
        v.z0[li] = o.z0;
        v.s_rho[li] = o.s_rho;
        v.flag[li] = (unsigned int)o.m_num < tc.mnt ? 1 : 0;
        v.res[0][li] = 0.0;   // for (auto &r : Residual) r = 0   (:625)
    }
    const bool fused = plan.merge_round >= 0;
    if (tid == 0) {
        sm.abort = 0;
        lm_begin(sm.lm, old_st, f_st, su.VW, su.a, su.max_r, su.max_s_rho, su.s_rho_from_state, su.frame_count,
                 su.fc_from_state);
        if (fused) {   // the prior-initialised try starts beside the zero-initialised one (:696-700)
            for (int i = 0; i < 3; i++) {
                sm.lm.X[i] = sm.lm.Vel_in[i];
                sm.lm.X[3 + i] = sm.lm.W0_in[i];
            }
            lm_request(sm.lm, sm.lm.X, -1, sm.lm.iRN);
        }
        if (XCHG) {
            mc_mbar_init(&sm.mbar[0], 1);
            mc_mbar_init(&sm.mbar[1], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    } else if (tid == 32 && fused) {
        lm_begin(sm.lmz, old_st, f_st, su.VW, su.a, su.max_r, su.max_s_rho, su.s_rho_from_state, su.frame_count,
                 su.fc_from_state);   // X = 0, request {0, -1, Rest}
    }
    __syncthreads();
    mc_cluster_sync();   // every CTA's barriers exist before anybody sends
    MC_STAMP(15, 2);

    for (int e = 0; e < plan.n; e++) {
        const int sa = plan.sa[e], sb = plan.sb[e];
        const bool two = sa != STEP_NONE;
        const bool RW = sb >= STEP_MAIN_FIRST;
        const bool PJ = !(sb == STEP_INIT_LAST_ZERO || sb == STEP_INIT_LAST_PRIOR);
        const int par = e & 1;
        const unsigned int nval = two ? 2 * MC_PW : MC_PW;
        MC_STAMP(e, 0);
        // ---- poses of this round: R = exp(W), RotM = exp((0,0,W.z)) (:309-314), four exponentials on four warps ------
        if (tid == 0) {
            if (e > 0 && e - 1 == plan.merge_round) mc_lm_merge(sm.lm, sm.lmz);
            so3_exp(sm.lm.Xeval + 3, sm.req[0]);
            for (int k = 0; k < 3; k++) sm.req[0][9 + k] = sm.lm.Xeval[k];
            sm.req_res[0][0] = sm.lm.res_in;
            sm.req_res[0][1] = sm.lm.res_out;
            if (e > 0 && e - 1 == plan.merge_round) {   // the z rotation depends on the merge as well
                double wz[3] = {0, 0, sm.lm.Xeval[5]}, RMf[9];
                so3_exp(wz, RMf);
                sm.req[0][12] = RMf[0];
                sm.req[0][13] = RMf[1];
                sm.req[0][14] = RMf[3];
                sm.req[0][15] = RMf[4];
            }
            if (XCHG) mc_mbar_expect_tx(&sm.mbar[par], nval * MC_C * 8u);
        } else if (tid == 32) {
            if (!(e > 0 && e - 1 == plan.merge_round)) {
                double wz[3] = {0, 0, sm.lm.Xeval[5]}, RMf[9];
                so3_exp(wz, RMf);
                sm.req[0][12] = RMf[0];
                sm.req[0][13] = RMf[1];
                sm.req[0][14] = RMf[3];
                sm.req[0][15] = RMf[4];
            }
        } else if (tid == 64 && two) {
            so3_exp(sm.lmz.Xeval + 3, sm.req[1]);
            for (int k = 0; k < 3; k++) sm.req[1][9 + k] = sm.lmz.Xeval[k];
            sm.req_res[1][0] = sm.lmz.res_in;
            sm.req_res[1][1] = sm.lmz.res_out;
        } else if (tid == 96 && two) {
            double wz[3] = {0, 0, sm.lmz.Xeval[5]}, RMf[9];
            so3_exp(wz, RMf);
            sm.req[1][12] = RMf[0];
            sm.req[1][13] = RMf[1];
            sm.req[1][14] = RMf[3];
            sm.req[1][15] = RMf[4];
        }
        __syncthreads();
        MC_STAMP(e, 1);
        // ---- keylines ---------------------------------------------------------------------------------------
        const bool write_mid = e == plan.n - 1;
        if (RW) mc_eval_pose<true, true>(sm, v, 0, old, tc, cam, field, fpack, write_mid, tid, lane, wid);
        else if (PJ) mc_eval_pose<false, true>(sm, v, 0, old, tc, cam, field, fpack, write_mid, tid, lane, wid);
        else mc_eval_pose<false, false>(sm, v, 0, old, tc, cam, field, fpack, write_mid, tid, lane, wid);
        if (two) {
            if (PJ) mc_eval_pose<false, true>(sm, v, 1, old, tc, cam, field, fpack, false, tid, lane, wid);
            else mc_eval_pose<false, false>(sm, v, 1, old, tc, cam, field, fpack, false, tid, lane, wid);
        }
        __syncthreads();
        MC_STAMP(e, 2);
        // ---- all-to-all of the CTAs' sums through distributed shared memory ------------------------------------
        {
            const unsigned int np = nval / 2;
            for (unsigned int idx = tid; idx < np * MC_C; idx += MC_T) {
                const unsigned int dst = idx / np, pair = idx - dst * np;
                const unsigned int la = mc_smem_u32(&sm.gather[par][rank][2 * pair]);
                if (XCHG) mc_st_async_v2(mc_mapa(la, dst), sm.xout[2 * pair], sm.xout[2 * pair + 1],
                                         mc_mapa(mc_smem_u32(&sm.mbar[par]), dst));
                else mc_st_remote_v2(mc_mapa(la, dst), sm.xout[2 * pair], sm.xout[2 * pair + 1]);
            }
            if (XCHG) {
                const unsigned int ph = (unsigned int)(e >> 1) & 1u;
                const long long t0 = clock64();
                while (!mc_mbar_try_wait(&sm.mbar[par], ph)) {
                    if (clock64() - t0 > MC_SPIN_LIMIT) {
                        sm.abort = 1;
                        break;
                    }
                }
            } else {
                mc_cluster_sync();
            }
        }
        MC_STAMP(e, 3);
        // ---- totals in rank order; stale-fi carries of this CTA's virtual warps ---------------------------------
        if (tid < 64) {
            const int p = tid >> 5, k = tid & 31;
            if (k < 28 && (p == 0 || two)) {
                double t = sm.gather[par][0][p * MC_PW + k];
#pragma unroll
                for (int r = 1; r < MC_C; r++) t += sm.gather[par][r][p * MC_PW + k];
                sm.tot[p][k] = t;
            }
        } else {
            const int q = tid - 64, p = q / MC_NV, vw = q - p * MC_NV;
            if (p < (two ? 2 : 1) && vw < v.J * MC_NW) {
                double cy = 0;
                bool found = false;
                for (int w2 = vw - 1; w2 >= 0 && !found; w2--)
                    if (sm.vw_has[p][w2]) {
                        cy = sm.vw_last[p][w2];
                        found = true;
                    }
                for (int r = rank - 1; r >= 0 && !found; r--)
                    if (sm.gather[par][r][p * MC_PW + 28] != 0.0) {
                        cy = sm.gather[par][r][p * MC_PW + 29];
                        found = true;
                    }
                sm.wcarry[sm.req_res[p][1]][vw] = cy;
            }
        }
        __syncthreads();
        if (sm.abort) break;
        MC_STAMP(e, 4);
        // ---- LM steps (every CTA computes the same step on the same totals) ---------------------------------------
        if (tid == 0) {
            if (PJ) lm_ingest<true>(sm.lm, sm.tot[0]);
            else lm_ingest<false>(sm.lm, sm.tot[0]);
            mc_lm_step_main(sm.lm, sb, rank == 0 ? f_st : nullptr);
        } else if (tid == 32 && two) {
            if (PJ) lm_ingest<true>(sm.lmz, sm.tot[1]);
            else lm_ingest<false>(sm.lmz, sm.tot[1]);
            mc_lm_step_zero(sm.lmz, sa);
        }
        __syncthreads();
        MC_STAMP(e, 5);
    }
    // ---- epilogue ------------------------------------------------------------------------------------------------
    if (!sm.abort) lm_finalize_cov(sm.lm, tid);   // the six columns of Cholesky<6>(JtJ).get_inverse() side by side
    __syncthreads();
    if (sm.abort && tid == 0) {
        *abort_out = 1;
        for (int k = 0; k < 3; k++) sm.lm.Vel[k] = sm.lm.W0[k] = __longlong_as_double(0x7FF8000000000000ll);
    }
    __syncthreads();
    if (rank == 0) {
        const double *src = reinterpret_cast<const double *>(&sm.lm);
        double *dst = reinterpret_cast<double *>(lm_out);
        for (int k = tid; k < (int)(sizeof(LMState) / sizeof(double)); k += MC_T) dst[k] = src[k];
        if (tid == 0 && post_fs) d_frame_post_min(post_fs, sm.lm);   // folded one-thread stage of the per-frame pipeline
    }
    MC_STAMP(15, 3);
    mc_cluster_sync();   // no CTA exits while a peer may still send to it
    MC_STAMP(15, 4);
}
