// emu.cpp -- TEST INFRASTRUCTURE: compiles the product kernel source (uhc_b200/csrc/sim_core.h, env_step.h) as a host
// lane-loop emulation (-DUHC_EMU) so the warp-per-env algorithms can be checked against the oracle on a CPU-only box.
// Never loaded by the product path (uhc_b200/engine.py only loads the CUDA library).
#define UHC_EMU 1
#include <vector>
#include <cstring>
#include "../../include/uhc_b200.h"
#include "../../uhc_b200/csrc/env_step.h"

using namespace uhc;

template <class Real>
struct Emu {
    std::vector<Real> body_f, dof_f, dof_lim, hull, state, expert, shape;
    std::vector<int> lvl_tab, lvl_pack, hull_adr, hull_num, nbr, nbradr, parent, depth, child_adr, child, body_sub_end, ee, istate, clip_adr;
    EngineView<Real> ev; Work<Real> w; int E;
};

template <class Real, class T> static void cp(std::vector<Real> &d, const T *s, size_t n) { d.resize(n); for (size_t i = 0; i < n; i++) d[i] = (Real)s[i]; }

template <class Real>
static void set_cfg(EnvCfg<Real> &c, const UhcEnvCfg *h) {
    for (int i = 0; i < 4; i++) c.base_rot[i] = (Real)h->base_rot[i];
    c.rfc_scale = (Real)h->rfc_scale; c.rfc_lim = (Real)h->rfc_lim; c.rfc_rate = (Real)h->rfc_rate; c.body_diff_thresh = (Real)h->body_diff_thresh;
    c.meta_pd = h->meta_pd; c.env_episode_len = h->env_episode_len; c.trail_steps = h->trail_steps; c.newton_max_iter = h->newton_max_iter;
    for (int i = 0; i < 5; i++) { c.w[i] = (Real)h->w[i]; c.k[i] = (Real)h->k[i]; }
    c.newton_tol = (Real)h->newton_tol;
    c.reactive_v = 0; c.reactive_rate = 0; c.auto_reset = 0; c.t_min = h->t_min; c.t_max = h->t_max; c.reset_seed = h->reset_seed; c.num_clips = 0;
    c.rfc_mode = (h->rfc_mode == 1 || h->rfc_mode == 2) ? h->rfc_mode : 0; c.vf_dim = c.rfc_mode == 1 ? VF_BODY_DIM * NB : (c.rfc_mode == 2 ? 0 : 6); c.act_dim = NU + c.vf_dim + (h->meta_pd ? 2 * NSUB : 0);
    for (int b = 0; b < NB; b++) c.vf_slot[b] = (signed char)h->vf_slot[b];
    c.obs_v = (h->obs_v == 1 || h->obs_v == 3 || h->obs_v == 5 || h->obs_v == 6) ? h->obs_v : 2;
    c.fut_frames = h->fut_frames > 0 ? h->fut_frames : 10; c.fut_skip = h->fut_skip > 0 ? h->fut_skip : 10;
    c.has_shape = h->no_shape ? 0 : 1; c.obs_block = c.has_shape ? OBS_DIM : OBS_DIM - 17;
    c.obs_dim = c.obs_v == 1 ? OBS_DIM_V1 : (c.obs_v == 3 ? c.obs_block * c.fut_frames : c.obs_block);
    if (c.obs_v == 5 || c.obs_v == 6) c.obs_dim = (c.obs_v == 5 ? 636 : 384) + (c.has_shape ? 17 : 0);      // get_full_obs_v5 / v6
    c.term_body = (h->term_body == 1 || h->term_body == 2) ? h->term_body : 0; c.head_body = (h->head_body >= 0 && h->head_body < NB) ? h->head_body : 13;
    c.reward_mul = h->reward_mul ? 1 : 0;
}

template <class Real>
static Emu<Real> *create(const UhcModelHost *m, const UhcEnvCfg *cfg, int E) {
    Emu<Real> *e = new Emu<Real>();
    e->E = E;
    cp(e->body_f, m->body_f, NB * BODYF); cp(e->dof_f, m->dof_f, NV * 4); cp(e->hull, m->hull, (size_t)m->nvert * 3);
    cp(e->hull_adr, m->hull_adr, NB); cp(e->hull_num, m->hull_num, NB); cp(e->nbr, m->nbr, m->nnbr); cp(e->nbradr, m->nbradr, m->nvert + 1);
    cp(e->parent, m->parent, NB); cp(e->depth, m->depth, NB); cp(e->child_adr, m->child_adr, NB + 1); cp(e->child, m->child, NB - 1);
    cp(e->body_sub_end, m->body_sub_end, NB); cp(e->ee, m->ee, 5); cp(e->lvl_tab, m->lvl_tab, (MAXLEVEL + 1) * LVL_G * 5); cp(e->lvl_pack, m->lvl_pack, (MAXLEVEL + 1) * LVL_G);
    Model<Real> &M = e->ev.model;
    cp(e->dof_lim, m->dof_lim, NV * 4); M.dof_lim = e->dof_lim.data();
    M.body_f = e->body_f.data(); M.dof_f = e->dof_f.data(); M.hull = e->hull.data(); M.hull_adr = e->hull_adr.data(); M.hull_num = e->hull_num.data();
    M.nbr = e->nbr.data(); M.nbradr = e->nbradr.data(); M.parent = e->parent.data(); M.depth = e->depth.data(); M.child_adr = e->child_adr.data();
    M.child = e->child.data(); M.body_sub_end = e->body_sub_end.data();
    M.ee = e->ee.data(); M.lvl_tab = e->lvl_tab.data(); M.lvl_pack = e->lvl_pack.data(); M.topo_s = nullptr;
    M.dt = (Real)m->dt; M.margin = (Real)m->margin; M.mu = (Real)m->mu; M.solref0 = (Real)m->solref[0]; M.solref1 = (Real)m->solref[1];
    M.simp0 = (Real)m->solimp[0]; M.simp1 = (Real)m->solimp[1]; M.simp2 = (Real)m->solimp[2]; M.simp3 = (Real)m->solimp[3]; M.simp4 = (Real)m->solimp[4];
    M.gravz = (Real)m->gravz; M.nshape = 1; M.nvert = m->nvert;
    set_cfg(e->ev.cfg, cfg);
    e->ev.clip_model = nullptr; e->ev.clip_cdf = nullptr; e->ev.counters = nullptr; e->ev.ep_log = nullptr; e->ev.neutral = nullptr;
    e->state.assign((size_t)E * ST_SIZE, 0); e->istate.assign((size_t)E * SI_SIZE, 0);
    e->ev.num_envs = E; e->ev.state = e->state.data(); e->ev.istate = e->istate.data();
    return e;
}

template <class Real>
static void load_clips(Emu<Real> *e, int nclips, const int *len, const double *frames, const double *shape) {
    e->clip_adr.assign(nclips + 1, 0);
    for (int i = 0; i < nclips; i++) e->clip_adr[i + 1] = e->clip_adr[i] + len[i];
    cp(e->expert, frames, (size_t)e->clip_adr[nclips] * EX_SIZE); cp(e->shape, shape, (size_t)nclips * 17);
    e->ev.expert = e->expert.data(); e->ev.clip_adr = e->clip_adr.data(); e->ev.clip_shape = e->shape.data();
}

template <class Real, class TPT>
static int emu_forward_given_tau(const Model<Real> &m, const EnvCfg<Real> &cfg, Work<Real> &w, const TPT &tp, const double *fapp_d) {
    // same sequence as substep_dynamics' PH_SMOOTH / PH_NEWTON phases with caller-provided torques and applied force
    kin_rne_forward(m, w, tp);
    project_force(m, w, w.Fb, w.C, Real(1), (const Real *)nullptr);
    collide(m, w, tp);
    for (int i = 0; i < NV; i++) { Real f = -w.C[i] + (i < 6 ? (Real)fapp_d[i] : w.tau[i - 6]); w.fs[i] = f; w.as_[i] = f; }
    limit_setup(m, w);
    if (w.ncon == 0 && w.nlim == 0) { aba_solve(m, w, Real(0), false, w.as_); for (int i = 0; i < NV; i++) w.a[i] = w.as_[i]; return 0; }
    constraint_setup(m, w);
    Real gn2 = 0;
    const Real scale = newton_init(m, w, tp, &gn2);
    int iters = 0;
    while (iters < cfg.newton_max_iter && gn2 > cfg.newton_tol * cfg.newton_tol * scale * scale) {
        ++iters;
        aba_solve(m, w, Real(0), true, w.p);
        if (newton_advance(m, w, tp, &gn2)) break;
    }
    return iters;
}

extern "C" {
void *emu_create(const UhcModelHost *m, const UhcEnvCfg *cfg, int E, int precision) {
    return precision == 64 ? (void *)create<double>(m, cfg, E) : (void *)create<float>(m, cfg, E);
}
#define DISPATCH(h, prec, ...) do { if (prec == 64) { auto *e = (Emu<double> *)h; typedef double Real; (void)sizeof(Real); __VA_ARGS__; } else { auto *e = (Emu<float> *)h; typedef float Real; (void)sizeof(Real); __VA_ARGS__; } } while (0)
void emu_destroy(void *h, int prec) { DISPATCH(h, prec, delete e); }
void emu_load_clips(void *h, int prec, int nclips, const int *len, const double *frames, const double *shape) { DISPATCH(h, prec, load_clips(e, nclips, len, frames, shape)); }
void emu_reset(void *h, int prec, int env, int clip, int start, int len, const double *qpos, const double *qvel, double *obs) {
    DISPATCH(h, prec, {
        std::vector<Real> q, v; if (qpos) cp(q, qpos, NQ); if (qvel) cp(v, qvel, NV);
        env_reset_warp<Real, double>(e->ev, env, e->w, clip, start, len, qpos ? q.data() : nullptr, qvel ? v.data() : nullptr, obs);
    });
}
int emu_step(void *h, int prec, int env, const double *action, double *obs, double *reward, double *cinfo, int *fail, int *end, double *percent, double *torque) {
    int done = 0;
    DISPATCH(h, prec, done = (env_step_warp<Real, double>(e->ev, env, e->w, action, obs, reward, cinfo, fail, end, percent, torque)));
    return done;
}
void emu_get_state(void *h, int prec, int env, double *out, int *iout) {
    DISPATCH(h, prec, { for (int i = 0; i < ST_SIZE; i++) out[i] = (double)e->state[(size_t)env * ST_SIZE + i]; for (int i = 0; i < SI_SIZE; i++) iout[i] = e->istate[(size_t)env * SI_SIZE + i]; });
}
// forward pass only on an arbitrary state: returns dense M, C, qacc, xpos for unit tests
void emu_forward(void *h, int prec, const double *qpos, const double *qvel, const double *tau, const double *fapp, const double *aw, double *Msparse,
                 double *Cout, double *qacc, double *xpos, int *ncon, int *iters) {   // Msparse: unused (no joint-space matrix exists)
    DISPATCH(h, prec, {
        Work<Real> &w = e->w;
        for (int i = 0; i < NQ; i++) w.q[i] = (Real)qpos[i];
        for (int i = 0; i < NV; i++) { w.v[i] = (Real)qvel[i]; w.aw[i] = aw ? (Real)aw[i] : Real(0); }
        for (int i = 0; i < NU; i++) w.tau[i] = (Real)tau[i];
        // residual force goes through the action: vf = act[69:75] * rfc_scale rotated by the heading -> use a unit-scale, identity-heading cfg
        (void)fapp;
        for (int i = 0; i < ACT_DIM; i++) w.act[i] = 0;
        TOPO_DECL(e->ev.model);
        w.upper_contact = 0;
        // emulate "with_pd" minus the PD phase: torques are given
        *iters = emu_forward_given_tau<Real>(e->ev.model, e->ev.cfg, w, tp, fapp);
        (void)Msparse;
        for (int i = 0; i < NV; i++) { Cout[i] = (double)w.C[i]; qacc[i] = (double)w.a[i]; }
        for (int i = 0; i < 72; i++) xpos[i] = (double)(&w.xpos[0][0])[i];
        *ncon = w.ncon;
    });
}
}
