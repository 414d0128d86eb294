/* uhc_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * fp64, single-environment, dense-linear-algebra CPU restatement of the reference's hot path
 * (SURVEY.md section 8a rows a1-a10).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product path (uhc_b200/csrc) shares no code with it:
 * the algorithms here are deliberately the dense/naive ones (Jacobian-sum mass matrix, projected
 * Newton-Euler bias, dense Cholesky, dense Newton) so that the warp-parallel tree-sparse CUDA kernels are
 * checked against an independent implementation.
 *
 * PARITY STATUS -- read before trusting:
 *   * L0 physics (or_forward / or_step) restates MuJoCo 2.1.0's default pipeline for the model class of
 *     assets/mujoco_models/humanoid_smpl_neutral_mesh.xml.  MuJoCo is a closed binary absent from
 *     /root/reference and from this container, and the reference holds no test or recorded trajectory at
 *     that boundary  ==>  "parity unpinned" for or_forward/or_step against MuJoCo itself.  What IS pinned:
 *     physical invariants (tests/test_oracle_physics.py).
 *   * Everything above L0 (stable-PD torque, implicit RFC, body quats, termination, observation v2,
 *     world_rfc_implicit reward) is pinned against the reference's own Python, executed unmodified on top of
 *     a fake mujoco_py backed by or_forward/or_step (tools/ref_harness.py -> tests/golden/).
 *
 * Reference call sites restated (file:line under /root/reference):
 *   or_env_step        uhc/envs/humanoid_im.py:1192-1243 (step), :1145-1190 (do_simulation)
 *   or_compute_torque  uhc/envs/humanoid_im.py:1033-1076, :1014-1031 (stable PD, dense (M+Kd dt) solve)
 *   or_rfc_implicit    uhc/envs/humanoid_im.py:1136-1143
 *   or_body_quat       uhc/envs/humanoid_im.py:925-947
 *   or_body_diff       uhc/envs/humanoid_im.py:1408-1415
 *   or_obs_v2          uhc/envs/humanoid_im.py:419-503
 *   or_reward          uhc/losses/reward_function.py:12-88
 *   or_env_reset       uhc/khrylib/rl/envs/common/mujoco_env.py:95-113, uhc/envs/humanoid_im.py:1245-1299
 *   or_forward/or_step the mujoco-py calls at uhc/envs/humanoid_im.py:1177 (sim.step), :905 and
 *                      mujoco_env.py:113 (sim.forward); MuJoCo semantics per SURVEY.md Appendix B.
 */
#define _GNU_SOURCE
#include <math.h>
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#define NB 24
#define NQ 76
#define NV 75
#define NU 69
#define MAXCON 96
#define MAXROW (4 * MAXCON + NU)   /* pyramid rows + one joint-limit row per hinge */
#define OBS_DIM 657
#define ACT_DIM 105

typedef struct {
    int parent[NB];
    double offset[NB][3], mass[NB], ipos[NB][3], inertia[NB][9], invw[NB], armature[NV];
    double jkp[NU], jkd[NU], tlim[NU], diffw[NB];
    int ee[5];
    int nvert, vadr[NB], vnum[NB];
    double *vert; int *nbr, *nbradr;
    double dt, margin, mu, solref[2], solimp[5], gravity[3];
    double jnt_range[NU][2], dof_invw[NV];   /* hinge limits (the model's default is limited="true", humanoid_smpl_neutral_mesh.xml:9) and dof_invweight0 */
} OrModel;

typedef struct {
    double qpos[NQ], qvel[NV], qacc_warm[NV];
    /* results of the last forward pass; stale w.r.t. qpos/qvel after integration, exactly as mj_step leaves them */
    double xpos[NB][3], xquat[NB][4], xmat[NB][9], xipos[NB][3];
    double axis[NV][3];               /* world axis of every rotational dof (root: body axes) */
    double M[NV * NV], C[NV], qacc[NV], qacc_smooth[NV];
    double ctrl[NU], qfrc_applied[NV];
    int ncon, con_body[MAXCON]; double con_pos[MAXCON][3], con_dist[MAXCON];
    int nrow, newton_iters; double efc_force[MAXROW];
} OrData;

/* ------------------------------------------------------------------ small math */
static void qmul(const double *a, const double *b, double *o) { /* o = a (x) b, wxyz */
    double w = a[0]*b[0] - a[1]*b[1] - a[2]*b[2] - a[3]*b[3];
    double x = a[0]*b[1] + a[1]*b[0] + a[2]*b[3] - a[3]*b[2];
    double y = a[0]*b[2] - a[1]*b[3] + a[2]*b[0] + a[3]*b[1];
    double z = a[0]*b[3] + a[1]*b[2] - a[2]*b[1] + a[3]*b[0];
    o[0] = w; o[1] = x; o[2] = y; o[3] = z;
}
static void qinv(const double *q, double *o) { /* conj / |q|^2 : uhc/utils/transformation.py:1509-1520 */
    double n = q[0]*q[0] + q[1]*q[1] + q[2]*q[2] + q[3]*q[3];
    o[0] = q[0]/n; o[1] = -q[1]/n; o[2] = -q[2]/n; o[3] = -q[3]/n;
}
static void qnormalize(double *q) {
    double n = sqrt(q[0]*q[0] + q[1]*q[1] + q[2]*q[2] + q[3]*q[3]);
    if (n < 1e-300) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }
    for (int i = 0; i < 4; i++) q[i] /= n;
}
static void q2mat(const double *qin, double *m) { /* rotation of the NORMALISED quaternion (transformation.py:1344) */
    double q[4] = {qin[0], qin[1], qin[2], qin[3]};
    double n = q[0]*q[0] + q[1]*q[1] + q[2]*q[2] + q[3]*q[3];
    if (n < 1e-300) { memset(m, 0, 72); m[0] = m[4] = m[8] = 1; return; }
    double s = 2.0 / n;
    double xx = s*q[1]*q[1], yy = s*q[2]*q[2], zz = s*q[3]*q[3];
    double xy = s*q[1]*q[2], xz = s*q[1]*q[3], yz = s*q[2]*q[3];
    double wx = s*q[0]*q[1], wy = s*q[0]*q[2], wz = s*q[0]*q[3];
    m[0] = 1-yy-zz; m[1] = xy-wz;   m[2] = xz+wy;
    m[3] = xy+wz;   m[4] = 1-xx-zz; m[5] = yz-wx;
    m[6] = xz-wy;   m[7] = yz+wx;   m[8] = 1-xx-yy;
}
static void mv(const double *m, const double *v, double *o) {
    double a = m[0]*v[0]+m[1]*v[1]+m[2]*v[2], b = m[3]*v[0]+m[4]*v[1]+m[5]*v[2], c = m[6]*v[0]+m[7]*v[1]+m[8]*v[2];
    o[0] = a; o[1] = b; o[2] = c;
}
static void mtv(const double *m, const double *v, double *o) {
    double a = m[0]*v[0]+m[3]*v[1]+m[6]*v[2], b = m[1]*v[0]+m[4]*v[1]+m[7]*v[2], c = m[2]*v[0]+m[5]*v[1]+m[8]*v[2];
    o[0] = a; o[1] = b; o[2] = c;
}
static void cross(const double *a, const double *b, double *o) {
    double x = a[1]*b[2]-a[2]*b[1], y = a[2]*b[0]-a[0]*b[2], z = a[0]*b[1]-a[1]*b[0];
    o[0] = x; o[1] = y; o[2] = z;
}
static double dot3(const double *a, const double *b) { return a[0]*b[0]+a[1]*b[1]+a[2]*b[2]; }

/* dense Cholesky A = L L^T in place (lower), n<=NV; returns 0 ok / -1 not PD */
static int chol(double *A, int n) {
    for (int j = 0; j < n; j++) {
        double s = A[j*n+j];
        for (int k = 0; k < j; k++) s -= A[j*n+k]*A[j*n+k];
        if (s <= 0) return -1;
        double d = sqrt(s); A[j*n+j] = d;
        for (int i = j+1; i < n; i++) {
            double t = A[i*n+j];
            for (int k = 0; k < j; k++) t -= A[i*n+k]*A[j*n+k];
            A[i*n+j] = t/d;
        }
    }
    return 0;
}
static void chol_solve(const double *L, int n, double *b) {
    for (int i = 0; i < n; i++) { double s = b[i]; for (int k = 0; k < i; k++) s -= L[i*n+k]*b[k]; b[i] = s/L[i*n+i]; }
    for (int i = n-1; i >= 0; i--) { double s = b[i]; for (int k = i+1; k < n; k++) s -= L[k*n+i]*b[k]; b[i] = s/L[i*n+i]; }
}

/* ------------------------------------------------------------------ model / data */
OrModel *or_model_create(const int *parent, const double *offset, const double *mass, const double *ipos,
                         const double *inertia, const double *invw_tran, const double *armature, const double *jkp,
                         const double *jkd, const double *tlim, const double *diffw, const int *ee, int nvert,
                         const double *vert, const int *vadr, const int *vnum, const int *nbr, const int *nbradr,
                         double dt, double margin, double mu, const double *solref, const double *solimp,
                         const double *gravity) {
    OrModel *m = (OrModel *)calloc(1, sizeof(OrModel));
    memcpy(m->parent, parent, sizeof m->parent); memcpy(m->offset, offset, sizeof m->offset);
    memcpy(m->mass, mass, sizeof m->mass); memcpy(m->ipos, ipos, sizeof m->ipos);
    memcpy(m->inertia, inertia, sizeof m->inertia); memcpy(m->invw, invw_tran, sizeof m->invw);
    memcpy(m->armature, armature, sizeof m->armature); memcpy(m->jkp, jkp, sizeof m->jkp);
    memcpy(m->jkd, jkd, sizeof m->jkd); memcpy(m->tlim, tlim, sizeof m->tlim);
    memcpy(m->diffw, diffw, sizeof m->diffw); memcpy(m->ee, ee, sizeof m->ee);
    m->nvert = nvert; memcpy(m->vadr, vadr, sizeof m->vadr); memcpy(m->vnum, vnum, sizeof m->vnum);
    m->vert = (double *)malloc(sizeof(double)*3*nvert); memcpy(m->vert, vert, sizeof(double)*3*nvert);
    m->nbradr = (int *)malloc(sizeof(int)*(nvert+1)); memcpy(m->nbradr, nbradr, sizeof(int)*(nvert+1));
    m->nbr = (int *)malloc(sizeof(int)*nbradr[nvert]); memcpy(m->nbr, nbr, sizeof(int)*nbradr[nvert]);
    m->dt = dt; m->margin = margin; m->mu = mu;
    memcpy(m->solref, solref, sizeof m->solref); memcpy(m->solimp, solimp, sizeof m->solimp);
    memcpy(m->gravity, gravity, sizeof m->gravity);
    for (int j = 0; j < NU; j++) { m->jnt_range[j][0] = -M_PI; m->jnt_range[j][1] = M_PI; }
    for (int i = 0; i < NV; i++) m->dof_invw[i] = 1.0;
    return m;
}
void or_model_set_limits(OrModel *m, const double *jnt_range, const double *dof_invw) {
    memcpy(m->jnt_range, jnt_range, sizeof m->jnt_range); memcpy(m->dof_invw, dof_invw, sizeof m->dof_invw);
}
void or_model_free(OrModel *m) { if (m) { free(m->vert); free(m->nbr); free(m->nbradr); free(m); } }
OrData *or_data_create(void) { OrData *d = (OrData *)calloc(1, sizeof(OrData)); d->qpos[3] = 1; return d; }
void or_data_free(OrData *d) { free(d); }
/* field access for ctypes: 0 qpos 1 qvel 2 qacc_warm 3 xpos 4 xquat 5 xipos 6 M 7 C 8 qacc 9 ctrl 10 qfrc_applied
   11 con_pos 12 con_dist 13 efc_force 14 qacc_smooth 15 xmat */
double *or_field(OrData *d, int id) {
    switch (id) {
    case 0: return d->qpos; case 1: return d->qvel; case 2: return d->qacc_warm; case 3: return &d->xpos[0][0];
    case 4: return &d->xquat[0][0]; case 5: return &d->xipos[0][0]; case 6: return d->M; case 7: return d->C;
    case 8: return d->qacc; case 9: return d->ctrl; case 10: return d->qfrc_applied; case 11: return &d->con_pos[0][0];
    case 12: return d->con_dist; case 13: return d->efc_force; case 14: return d->qacc_smooth; case 15: return &d->xmat[0][0];
    }
    return NULL;
}
int or_ncon(const OrData *d) { return d->ncon; }
int or_con_body(const OrData *d, int i) { return d->con_body[i]; }
int or_newton_iters(const OrData *d) { return d->newton_iters; }

/* ------------------------------------------------------------------ kinematics */
static void or_kinematics(const OrModel *m, OrData *d) {
    /* root: free joint. xpos = qpos[0:3], xquat = normalised qpos[3:7] */
    double q0[4] = {d->qpos[3], d->qpos[4], d->qpos[5], d->qpos[6]};
    qnormalize(q0);
    memcpy(d->xpos[0], d->qpos, 24); memcpy(d->xquat[0], q0, 32); q2mat(q0, d->xmat[0]);
    for (int k = 0; k < 3; k++) { d->axis[3+k][0] = d->xmat[0][k]; d->axis[3+k][1] = d->xmat[0][3+k]; d->axis[3+k][2] = d->xmat[0][6+k]; }
    for (int k = 0; k < 3; k++) { d->axis[k][0] = d->axis[k][1] = d->axis[k][2] = 0; d->axis[k][k] = 1; }
    static const double LOC[3][3] = {{0,0,1},{0,1,0},{1,0,0}}; /* hinge order z, y, x */
    for (int b = 1; b < NB; b++) {
        int p = m->parent[b]; double r[3], q[4];
        mv(d->xmat[p], m->offset[b], r);
        for (int k = 0; k < 3; k++) d->xpos[b][k] = d->xpos[p][k] + r[k];
        memcpy(q, d->xquat[p], 32);
        for (int j = 0; j < 3; j++) {
            double R[9], ang = d->qpos[7+3*(b-1)+j], h = 0.5*ang, ql[4], qn[4];
            q2mat(q, R); mv(R, LOC[j], d->axis[6+3*(b-1)+j]); /* axis seen in the frame BEFORE this joint rotates */
            ql[0] = cos(h); ql[1] = sin(h)*LOC[j][0]; ql[2] = sin(h)*LOC[j][1]; ql[3] = sin(h)*LOC[j][2];
            qmul(q, ql, qn); memcpy(q, qn, 32);
        }
        qnormalize(q); memcpy(d->xquat[b], q, 32); q2mat(q, d->xmat[b]);
    }
    for (int b = 0; b < NB; b++) { double c[3]; mv(d->xmat[b], m->ipos[b], c); for (int k = 0; k < 3; k++) d->xipos[b][k] = d->xpos[b][k] + c[k]; }
}

/* translational / rotational Jacobian columns of body b at world point p, for every dof (zeros off the chain) */
static void or_jac(const OrModel *m, const OrData *d, int b, const double *p, double Jv[3][NV], double Jw[3][NV]) {
    memset(Jv, 0, sizeof(double)*3*NV); memset(Jw, 0, sizeof(double)*3*NV);
    for (int k = 0; k < 3; k++) Jv[k][k] = 1;
    for (int k = 0; k < 3; k++) {
        double r[3] = {p[0]-d->xpos[0][0], p[1]-d->xpos[0][1], p[2]-d->xpos[0][2]}, c[3];
        cross(d->axis[3+k], r, c);
        for (int i = 0; i < 3; i++) { Jw[i][3+k] = d->axis[3+k][i]; Jv[i][3+k] = c[i]; }
    }
    for (int a = b; a > 0; a = m->parent[a])
        for (int j = 0; j < 3; j++) {
            int dof = 6+3*(a-1)+j; double r[3] = {p[0]-d->xpos[a][0], p[1]-d->xpos[a][1], p[2]-d->xpos[a][2]}, c[3];
            cross(d->axis[dof], r, c);
            for (int i = 0; i < 3; i++) { Jw[i][dof] = d->axis[dof][i]; Jv[i][dof] = c[i]; }
        }
}

static void world_inertia(const OrModel *m, const OrData *d, int b, double *Iw) {
    double T[9]; const double *R = d->xmat[b], *I = m->inertia[b];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { double s = 0; for (int k = 0; k < 3; k++) s += R[i*3+k]*I[k*3+j]; T[i*3+j] = s; }
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { double s = 0; for (int k = 0; k < 3; k++) s += T[i*3+k]*R[j*3+k]; Iw[i*3+j] = s; }
}

/* joint-space inertia by the Jacobian sum  M = sum_b m Jv^T Jv + Jw^T I Jw  + armature */
static void or_mass_matrix(const OrModel *m, OrData *d) {
    static double Jv[3][NV], Jw[3][NV];
    memset(d->M, 0, sizeof d->M);
    for (int b = 0; b < NB; b++) {
        double Iw[9]; world_inertia(m, d, b, Iw); or_jac(m, d, b, d->xipos[b], Jv, Jw);
        for (int i = 0; i < NV; i++) {
            double IJ[3];
            for (int r = 0; r < 3; r++) IJ[r] = Iw[r*3]*Jw[0][i] + Iw[r*3+1]*Jw[1][i] + Iw[r*3+2]*Jw[2][i];
            for (int j = 0; j < NV; j++)
                d->M[i*NV+j] += m->mass[b]*(Jv[0][i]*Jv[0][j] + Jv[1][i]*Jv[1][j] + Jv[2][i]*Jv[2][j])
                              + IJ[0]*Jw[0][j] + IJ[1]*Jw[1][j] + IJ[2]*Jw[2][j];
        }
    }
    for (int i = 0; i < NV; i++) d->M[i*NV+i] += m->armature[i];
}

/* bias force C(q,v) (Coriolis, centrifugal, gravity): forward recursion of velocity-product accelerations,
   then projection of the per-body inertial wrenches through the Jacobians */
static void or_bias(const OrModel *m, OrData *d) {
    static double Jv[3][NV], Jw[3][NV];
    double w[NB][3], al[NB][3], ao[NB][3];
    mv(d->xmat[0], &d->qvel[3], w[0]); /* root angular velocity is body-frame in qvel */
    memset(al[0], 0, 24); memset(ao[0], 0, 24);
    for (int b = 1; b < NB; b++) {
        int p = m->parent[b]; double r[3], t[3], t2[3];
        for (int k = 0; k < 3; k++) r[k] = d->xpos[b][k] - d->xpos[p][k];
        cross(al[p], r, t); cross(w[p], r, t2); cross(w[p], t2, t2);
        for (int k = 0; k < 3; k++) { ao[b][k] = ao[p][k] + t[k] + t2[k]; w[b][k] = w[p][k]; al[b][k] = al[p][k]; }
        for (int j = 0; j < 3; j++) {
            int dof = 6+3*(b-1)+j; double qd = d->qvel[dof], wa[3];
            cross(w[b], d->axis[dof], wa);
            for (int k = 0; k < 3; k++) { al[b][k] += wa[k]*qd; w[b][k] += d->axis[dof][k]*qd; }
        }
    }
    memset(d->C, 0, sizeof d->C);
    for (int b = 0; b < NB; b++) {
        double c[3], t[3], t2[3], f[3], tau[3], Iw[9], Iwv[3], Ial[3];
        for (int k = 0; k < 3; k++) c[k] = d->xipos[b][k] - d->xpos[b][k];
        cross(al[b], c, t); cross(w[b], c, t2); cross(w[b], t2, t2);
        for (int k = 0; k < 3; k++) f[k] = m->mass[b]*(ao[b][k] + t[k] + t2[k] - m->gravity[k]);
        world_inertia(m, d, b, Iw); mv(Iw, w[b], Iwv); mv(Iw, al[b], Ial); cross(w[b], Iwv, t);
        for (int k = 0; k < 3; k++) tau[k] = Ial[k] + t[k];
        or_jac(m, d, b, d->xipos[b], Jv, Jw);
        for (int i = 0; i < NV; i++)
            d->C[i] += Jv[0][i]*f[0] + Jv[1][i]*f[1] + Jv[2][i]*f[2] + Jw[0][i]*tau[0] + Jw[1][i]*tau[1] + Jw[2][i]*tau[2];
    }
}

/* floor (z=0, normal +z) against each body hull: deepest vertex if within margin, plus up to three of its
   hull-graph neighbours that are also within margin, taken in ascending vertex order (SURVEY.md Appendix B:
   MuJoCo's plane/mesh multi-point rule; neighbour order is qhull-internal there, so this selection is a stated
   convention, not a verified match). dist = signed height of the vertex; pos = vertex - 0.5*dist*n. */
static void or_collide(const OrModel *m, OrData *d) {
    d->ncon = 0;
    for (int b = 0; b < NB; b++) {
        int best = -1; double bz = 1e300;
        for (int i = 0; i < m->vnum[b]; i++) {
            const double *v = &m->vert[3*(m->vadr[b]+i)];
            double z = d->xpos[b][2] + d->xmat[b][6]*v[0] + d->xmat[b][7]*v[1] + d->xmat[b][8]*v[2];
            if (z < bz) { bz = z; best = i; }
        }
        if (best < 0 || bz > m->margin) continue;
        int cand[4], nc = 0; cand[nc++] = best;
        int g = m->vadr[b] + best;
        for (int e = m->nbradr[g]; e < m->nbradr[g+1] && nc < 4; e++) {
            const double *v = &m->vert[3*(m->vadr[b]+m->nbr[e])];
            double z = d->xpos[b][2] + d->xmat[b][6]*v[0] + d->xmat[b][7]*v[1] + d->xmat[b][8]*v[2];
            if (z <= m->margin) cand[nc++] = m->nbr[e];
        }
        for (int c = 0; c < nc; c++) {
            const double *v = &m->vert[3*(m->vadr[b]+cand[c])]; double p[3];
            mv(d->xmat[b], v, p);
            for (int k = 0; k < 3; k++) p[k] += d->xpos[b][k];
            int n = d->ncon++;
            d->con_body[n] = b; d->con_dist[n] = p[2];
            d->con_pos[n][0] = p[0]; d->con_pos[n][1] = p[1]; d->con_pos[n][2] = p[2] - 0.5*p[2];
        }
    }
}

typedef struct { double a; double dA, dB; } Brk;
static int brk_cmp(const void *x, const void *y) { double a = ((const Brk *)x)->a, b = ((const Brk *)y)->a; return (a > b) - (a < b); }

/* constraint rows (pyramidal condim-3 contacts, 4 rows each), soft-constraint parameters, and the convex problem
       min_a  1/2 (a-a_s)^T M (a-a_s) + sum_i  1/2 D_i min(0, J_i a - aref_i)^2
   solved to machine precision by exact Newton with an exact (sorted-breakpoint) line search. */
static void or_constraint_solve(const OrModel *m, OrData *d) {
    static double J[MAXROW][NV], Jv[3][NV], Jw[3][NV], H[NV*NV];
    double aref[MAXROW], D[MAXROW], r[MAXROW];
    int nrow = 0;
    const double t1[3] = {0, 1, 0}, t2[3] = {-1, 0, 0}; /* contact frame of normal (0,0,1) */
    const double mu = m->mu, dmin = m->solimp[0], dmax = m->solimp[1], width = m->solimp[2], mid = m->solimp[3], power = m->solimp[4];
    const double tc = m->solref[0], dr = m->solref[1];
    const double kk = 1.0/(dmax*dmax*tc*tc*dr*dr), bb = 2.0/(dmax*tc);
    for (int c = 0; c < d->ncon; c++) {
        int b = d->con_body[c];
        or_jac(m, d, b, d->con_pos[c], Jv, Jw);
        double pos = d->con_dist[c] - m->margin; /* efc_pos - efc_margin */
        double x = fabs(pos)/width; if (x > 1) x = 1;
        double y;
        if (x < mid) y = pow(x/mid, power)*mid; /* a*x^p with a = 1/mid^(p-1) */
        else y = 1 - pow((1-x)/(1-mid), power)*(1-mid);
        double imp = dmin + y*(dmax-dmin);
        double diagApprox = m->invw[b]*(1 + mu*mu);
        double R0 = (1-imp)*diagApprox/imp; if (R0 < 1e-15) R0 = 1e-15;
        double Rpy = 2*mu*mu*R0;
        for (int e = 0; e < 4; e++) {
            const double *t = (e < 2) ? t1 : t2; double sg = (e & 1) ? -mu : mu, vel = 0;
            for (int i = 0; i < NV; i++) { J[nrow][i] = Jv[2][i] + sg*(t[0]*Jv[0][i] + t[1]*Jv[1][i] + t[2]*Jv[2][i]); vel += J[nrow][i]*d->qvel[i]; }
            aref[nrow] = -bb*vel - kk*imp*pos; D[nrow] = 1.0/Rpy; nrow++;
        }
    }
    /* joint limits (mj_instantiateLimit, margin 0): a hinge past its range gets one unilateral row J = +-e_dof, pos = distance to the limit (< 0),
       default solref / solimp, diagApprox = dof_invweight0; frictionless, so R = (1 - d)/d * diagApprox without the pyramid factor */
    for (int j = 0; j < NU; j++) {
        const double q = d->qpos[7+j], lo = m->jnt_range[j][0], hi = m->jnt_range[j][1];
        double dist, sg;
        if (q - lo < 0) { dist = q - lo; sg = 1; } else if (hi - q < 0) { dist = hi - q; sg = -1; } else continue;
        double x = fabs(dist)/width; if (x > 1) x = 1;
        double y = (x < mid) ? pow(x/mid, power)*mid : 1 - pow((1-x)/(1-mid), power)*(1-mid);
        double imp = dmin + y*(dmax-dmin);
        double R = (1-imp)*m->dof_invw[6+j]/imp; if (R < 1e-15) R = 1e-15;
        memset(J[nrow], 0, sizeof J[nrow]); J[nrow][6+j] = sg;
        aref[nrow] = -bb*sg*d->qvel[6+j] - kk*imp*dist; D[nrow] = 1.0/R; nrow++;
    }
    d->nrow = nrow; d->newton_iters = 0;
    if (nrow == 0) { memcpy(d->qacc, d->qacc_smooth, sizeof d->qacc); return; }

    double a[NV], g[NV], p[NV], Ma[NV], cost_w, cost_s;
    /* warm start: previous qacc unless the unconstrained acceleration is cheaper */
    for (int pass = 0; pass < 2; pass++) {
        const double *x = pass ? d->qacc_smooth : d->qacc_warm; double cst = 0;
        for (int i = 0; i < NV; i++) { double s = 0; for (int j = 0; j < NV; j++) s += d->M[i*NV+j]*(x[j]-d->qacc_smooth[j]); cst += 0.5*s*(x[i]-d->qacc_smooth[i]); }
        for (int k = 0; k < nrow; k++) { double s = -aref[k]; for (int i = 0; i < NV; i++) s += J[k][i]*x[i]; if (s < 0) cst += 0.5*D[k]*s*s; }
        if (pass) cost_s = cst; else cost_w = cst;
    }
    memcpy(a, (cost_w < cost_s) ? d->qacc_warm : d->qacc_smooth, sizeof a);

    for (int it = 0; it < 100; it++) {
        for (int k = 0; k < nrow; k++) { double s = -aref[k]; for (int i = 0; i < NV; i++) s += J[k][i]*a[i]; r[k] = s; }
        for (int i = 0; i < NV; i++) { double s = 0; for (int j = 0; j < NV; j++) s += d->M[i*NV+j]*(a[j]-d->qacc_smooth[j]); Ma[i] = s; g[i] = s; }
        memcpy(H, d->M, sizeof H);
        for (int k = 0; k < nrow; k++) if (r[k] < 0) {
            for (int i = 0; i < NV; i++) { if (J[k][i] == 0) continue; g[i] += J[k][i]*D[k]*r[k]; for (int j = 0; j < NV; j++) H[i*NV+j] += D[k]*J[k][i]*J[k][j]; }
        }
        double gn = 0, sc = 0; for (int i = 0; i < NV; i++) { gn += g[i]*g[i]; sc += d->M[i*NV+i]; }
        if (sqrt(gn) < 1e-11*sc) break;
        d->newton_iters = it+1;
        if (chol(H, NV)) break;
        for (int i = 0; i < NV; i++) p[i] = -g[i];
        chol_solve(H, NV, p);
        /* exact line search on f(al) = f0 + al*g.p' ... : derivative is piecewise linear and increasing */
        double A = 0, B = 0; /* f'(al) = A + B al on the current segment */
        for (int i = 0; i < NV; i++) { double s = 0; for (int j = 0; j < NV; j++) s += d->M[i*NV+j]*p[j]; A += Ma[i]*p[i]; B += s*p[i]; }
        Brk brk[MAXROW]; int nb = 0;
        for (int k = 0; k < nrow; k++) {
            double s = 0; for (int i = 0; i < NV; i++) s += J[k][i]*p[i];
            if (r[k] < 0) { A += D[k]*r[k]*s; B += D[k]*s*s; }
            if (s != 0) { double al = -r[k]/s; if (al > 0) { int act = r[k] < 0; brk[nb].a = al; brk[nb].dA = (act ? -1 : 1)*D[k]*r[k]*s; brk[nb].dB = (act ? -1 : 1)*D[k]*s*s; nb++; } }
        }
        qsort(brk, nb, sizeof(Brk), brk_cmp);
        double al = 0; int ib = 0;
        for (;;) {
            double cand = (B > 0) ? -A/B : 1e300;
            if (ib >= nb || cand <= brk[ib].a) { al = cand; break; }
            A += brk[ib].dA; B += brk[ib].dB; ib++;
        }
        if (!(al > 0) || al > 1e299) break;
        for (int i = 0; i < NV; i++) a[i] += al*p[i];
    }
    memcpy(d->qacc, a, sizeof a);
    for (int k = 0; k < nrow; k++) { double s = -aref[k]; for (int i = 0; i < NV; i++) s += J[k][i]*a[i]; d->efc_force[k] = (s < 0) ? -D[k]*s : 0; }
}

/* mj_forward: position -> velocity -> actuation -> smooth acceleration -> constraint */
void or_forward(const OrModel *m, OrData *d) {
    static double L[NV*NV];
    or_kinematics(m, d); or_mass_matrix(m, d); or_collide(m, d); or_bias(m, d);
    for (int i = 0; i < NV; i++) d->qacc_smooth[i] = d->qfrc_applied[i] - d->C[i] + (i >= 6 ? d->ctrl[i-6] : 0.0);
    memcpy(L, d->M, sizeof L);
    if (chol(L, NV) == 0) chol_solve(L, NV, d->qacc_smooth);
    or_constraint_solve(m, d);
}

/* semi-implicit Euler: v += dt a ; q advanced with the NEW v ; root quaternion by the exponential map of the
   body-frame angular velocity */
static void or_integrate(const OrModel *m, OrData *d) {
    double dt = m->dt;
    for (int i = 0; i < NV; i++) d->qvel[i] += dt*d->qacc[i];
    for (int k = 0; k < 3; k++) d->qpos[k] += dt*d->qvel[k];
    double w[3] = {d->qvel[3], d->qvel[4], d->qvel[5]}, n = sqrt(dot3(w, w)), ang = n*dt, dq[4] = {1, 0, 0, 0}, qn[4];
    if (ang > 1e-300) { double s = sin(0.5*ang)/n; dq[0] = cos(0.5*ang); dq[1] = w[0]*s; dq[2] = w[1]*s; dq[3] = w[2]*s; }
    qmul(&d->qpos[3], dq, qn); qnormalize(qn); memcpy(&d->qpos[3], qn, 32);
    for (int i = 6; i < NV; i++) d->qpos[i+1] += dt*d->qvel[i];
    memcpy(d->qacc_warm, d->qacc, sizeof d->qacc);
}
void or_step(const OrModel *m, OrData *d) { or_forward(m, d); or_integrate(m, d); }

/* =================================================================== environment level */
typedef struct {
    int len;                          /* frames in the clip slice */
    const double *qpos, *qvel, *wbpos, *wbquat, *bquat, *bangvel, *ee_wpos, *com; /* T x {76,75,72,96,96,72,15,3} */
    const double *body_com;           /* T x 72 per-body centre-of-mass positions (expert["body_com"], obs v1 only); may be NULL */
    double shape_obs[17];             /* beta[16], gender */
} OrExpert;

typedef struct {
    OrModel *m; OrData *d; OrExpert ex;
    int cur_t, start_ind, mode_train;
    double bquat[96], prev_bquat[96];
    double base_rot[4], rfc_scale, rfc_lim, rfc_rate, body_diff_thresh;
    int meta_pd, env_episode_len, trail_steps;
    double w[5], k[5];
    double torque[15][NU];            /* per-substep applied torque of the last step (curr_torque) */
    int rfc_mode;                     /* 0 = implicit root wrench (6 action dims), 1 = explicit per-body forces (cfg.residual_force_mode, humanoid_im.py:231-243) */
    int vf_dim;                       /* action dims of the residual force: 6, or 9 per body x 24 bodies */
    int vf_body[NB];                  /* explicit: model body of residual-force slot i (vf_bodies = SMPL_BONE_ORDER_NAMES, humanoid_im.py:236-237) */
    int obs_v;                        /* cfg.obs_v: 2 = get_full_obs_v2 (657 with the shape vector), 1 = get_full_obs_v1 (784: + per-body COM blocks, no shape),
                                         3 = get_full_obs_v3 (fut_frames v2 blocks against the expert frames cur_t + 1 + i * skip, humanoid_im.py:505-513) */
    int fut_frames, fut_skip, obs_dt; /* obs_dt: the delta_t of the v2 block being written */
    int no_shape;                     /* cfg.has_shape false: the v2 block has no shape vector (640 dims) */
    int term_body, head_body;         /* cfg.env_term_body (humanoid_im.py:1223-1229): 0 "body", 1 "root", 2 "Head" (head_body = its model body) */
    int reward_mul;                   /* reward_id world_rfc_implicit_v1_mul (reward_function.py:174-250) */
    double height_lb, head_height_lb; /* expert["height_lb"], ["head_height_lb"] (uhc/utils/tools.py:94-95): lowest root / head height of the loaded expert */
} OrEnv;

OrEnv *or_env_create(OrModel *m) {
    OrEnv *e = (OrEnv *)calloc(1, sizeof(OrEnv)); e->m = m; e->d = or_data_create();
    double br[4] = {0.7071, 0.7071, 0, 0}; memcpy(e->base_rot, br, 32);
    e->rfc_scale = 100; e->rfc_lim = 100; e->rfc_rate = 1; e->body_diff_thresh = 0.5; e->meta_pd = 1;
    e->env_episode_len = 100000; e->trail_steps = 0; e->mode_train = 1;
    double w[5] = {0.3, 0.1, 0.45, 0.1, 0.05}, k[5] = {2.0, 0.005, 5.0, 100.0, 1.0};
    memcpy(e->w, w, 40); memcpy(e->k, k, 40);
    e->rfc_mode = 0; e->vf_dim = 6; for (int b = 0; b < NB; b++) e->vf_body[b] = b;
    e->obs_v = 2; e->fut_frames = 10; e->fut_skip = 10; e->obs_dt = 0;
    return e;
}
/* residual_force_mode (copycat_config.py:105-109): explicit = contact point + force + torque per body (residual_force_torque = True,
   residual_force_bodies = "all", residual_force_bodies_num = 1: the released uhc_explicit.yml); action = [69 joint targets | vf | 30 meta-PD] */
void or_env_set_rfc_mode(OrEnv *e, int explicit_mode, const int *vf_body) {
    e->rfc_mode = explicit_mode == 1 ? 1 : (explicit_mode == 2 ? 2 : 0); e->vf_dim = e->rfc_mode == 1 ? 9*NB : (e->rfc_mode == 2 ? 0 : 6);   /* 2: cfg.residual_force false */
    if (vf_body) memcpy(e->vf_body, vf_body, sizeof e->vf_body);
}
int or_env_action_dim(const OrEnv *e) { return NU + e->vf_dim + (e->meta_pd ? 30 : 0); }
void or_env_set_obs_v(OrEnv *e, int obs_v, const double *body_com) { e->obs_v = (obs_v == 1 || obs_v == 3 || obs_v == 5 || obs_v == 6) ? obs_v : 2; e->ex.body_com = body_com; }
void or_env_set_future(OrEnv *e, int fut_frames, int skip) { e->fut_frames = fut_frames > 0 ? fut_frames : 10; e->fut_skip = skip > 0 ? skip : 10; }
void or_env_set_has_shape(OrEnv *e, int has_shape) { e->no_shape = !has_shape; }
void or_env_set_reward_mul(OrEnv *e, int on) { e->reward_mul = on ? 1 : 0; }
/* call after or_env_set_expert: the bounds are minima over the loaded expert (tools.py:94-95; torch_smpl_humanoid.py:250 for height_lb) */
void or_env_set_term_body(OrEnv *e, int mode, int head_body) {
    e->term_body = (mode == 1 || mode == 2) ? mode : 0; e->head_body = (head_body >= 0 && head_body < NB) ? head_body : 13;
    e->height_lb = e->head_height_lb = 1e30;
    for (int t = 0; t < e->ex.len; t++) {
        if (e->ex.qpos[(size_t)t*NQ + 2] < e->height_lb) e->height_lb = e->ex.qpos[(size_t)t*NQ + 2];
        if (e->ex.wbpos[(size_t)t*3*NB + 3*e->head_body + 2] < e->head_height_lb) e->head_height_lb = e->ex.wbpos[(size_t)t*3*NB + 3*e->head_body + 2];
    }
}
static int obs_block_dim(const OrEnv *e) { return e->no_shape ? 640 : 657; }
int or_env_obs_dim(const OrEnv *e) {
    if (e->obs_v == 5) return 636 + (e->no_shape ? 0 : 17);       /* get_full_obs_v5 */
    if (e->obs_v == 6) return 384 + (e->no_shape ? 0 : 17);       /* get_full_obs_v6 */
    return e->obs_v == 1 ? 784 : (e->obs_v == 3 ? obs_block_dim(e)*e->fut_frames : obs_block_dim(e));
}
void or_env_free(OrEnv *e) { if (e) { or_data_free(e->d); free(e); } }
OrData *or_env_data(OrEnv *e) { return e->d; }
void or_env_set_expert(OrEnv *e, int len, const double *qpos, const double *qvel, const double *wbpos, const double *wbquat,
                       const double *bquat, const double *bangvel, const double *ee_wpos, const double *com, const double *shape_obs) {
    e->ex.len = len; e->ex.qpos = qpos; e->ex.qvel = qvel; e->ex.wbpos = wbpos; e->ex.wbquat = wbquat; e->ex.bquat = bquat;
    e->ex.bangvel = bangvel; e->ex.ee_wpos = ee_wpos; e->ex.com = com; memcpy(e->ex.shape_obs, shape_obs, sizeof e->ex.shape_obs);
}
void or_env_config(OrEnv *e, const double *base_rot, double rfc_scale, double rfc_lim, double rfc_rate, double thresh,
                   int meta_pd, int env_episode_len, int trail_steps, const double *w, const double *k) {
    memcpy(e->base_rot, base_rot, 32); e->rfc_scale = rfc_scale; e->rfc_lim = rfc_lim; e->rfc_rate = rfc_rate;
    e->body_diff_thresh = thresh; e->meta_pd = meta_pd; e->env_episode_len = env_episode_len; e->trail_steps = trail_steps;
    memcpy(e->w, w, 40); memcpy(e->k, k, 40);
}
int or_env_cur_t(const OrEnv *e) { return e->cur_t; }
double *or_env_bquat(OrEnv *e) { return e->bquat; }
double *or_env_prev_bquat(OrEnv *e) { return e->prev_bquat; }
double *or_env_torque(OrEnv *e) { return &e->torque[0][0]; }

static int ex_index(const OrEnv *e, int t) { int i = e->start_ind + t; return i < e->ex.len - 1 ? i : e->ex.len - 1; } /* humanoid_im.py:1322 */

static void heading_q(const double *q, double *hq) { /* math_utils.py:134-139 */
    hq[0] = q[0]; hq[1] = 0; hq[2] = 0; hq[3] = q[3]; double n = sqrt(hq[0]*hq[0] + hq[3]*hq[3]); hq[0] /= n; hq[3] /= n;
}
static double heading(const double *q) { /* math_utils.py:176-183 */
    double w = q[0], z = q[3]; if (z < 0) { w = -w; z = -z; } double n = sqrt(w*w + z*z); return 2*acos(w/n);
}
static void remove_base_rot(const OrEnv *e, const double *q, double *o) { double bi[4]; qinv(e->base_rot, bi); qmul(q, bi, o); }
static void euler_zyx_quat(double e0, double e1, double e2, double *q) { /* quaternion_from_euler(.., 'rzyx') = Rz(e0) Ry(e1) Rx(e2) */
    double qz[4] = {cos(e0/2), 0, 0, sin(e0/2)}, qy[4] = {cos(e1/2), 0, sin(e1/2), 0}, qx[4] = {cos(e2/2), sin(e2/2), 0, 0}, t[4];
    qmul(qz, qy, t); qmul(t, qx, q);
}
void or_body_quat(const OrEnv *e, double *out) { /* humanoid_im.py:925-947 */
    const double *qp = e->d->qpos; memcpy(out, qp+3, 32);
    for (int b = 1; b < NB; b++) euler_zyx_quat(qp[7+3*(b-1)], qp[8+3*(b-1)], qp[9+3*(b-1)], out+4*b);
}
double or_body_diff(const OrEnv *e) { /* humanoid_im.py:1408-1415 ; stale body_xpos, expert wbpos at cur_t */
    const double *ew = e->ex.wbpos + 72*ex_index(e, e->cur_t); double s = 0; int n = 0;
    for (int b = 0; b < NB; b++) if (e->m->diffw[b] != 0) {
        double dx[3]; for (int k = 0; k < 3; k++) dx[k] = (e->d->xpos[b][k] - ew[3*b+k])*e->m->diffw[b];
        s += sqrt(dot3(dx, dx)); n++;
    }
    return s/n;
}

/* stable PD with the dense (M + Kd dt) solve -- M, C are whatever the last forward pass left in data (stale by one
   substep after the first), qpos/qvel are current: humanoid_im.py:1014-1076 */
static void or_compute_torque(OrEnv *e, const double *ctrl, int it, double *torque) {
    static double A[NV*NV];
    const OrModel *m = e->m; OrData *d = e->d; double dt = m->dt;
    const double *tq = e->ex.qpos + NQ*ex_index(e, e->cur_t + 1) + 7;
    double kp[NV] = {0}, kd[NV] = {0}, err[NV] = {0}, rhs[NV];
    double sp = 1, sd = 1;
    if (e->meta_pd) { sp = ctrl[NU+e->vf_dim+it] + 1; sd = ctrl[NU+e->vf_dim+it+15] + 1; if (sp < 0) sp = 0; if (sp > 10) sp = 10; if (sd < 0) sd = 0; if (sd > 10) sd = 10; }
    for (int j = 0; j < NU; j++) {
        double base = tq[j], q = d->qpos[7+j];
        while (base - q > M_PI) base -= 2*M_PI;
        while (base - q < -M_PI) base += 2*M_PI;
        kp[6+j] = m->jkp[j]*sp; kd[6+j] = m->jkd[j]*sd;
        err[6+j] = q + d->qvel[6+j]*dt - (base + ctrl[j]);
    }
    memcpy(A, d->M, sizeof A);
    for (int i = 0; i < NV; i++) { A[i*NV+i] += kd[i]*dt; rhs[i] = -d->C[i] - kp[i]*err[i] - kd[i]*d->qvel[i]; }
    chol(A, NV); chol_solve(A, NV, rhs);
    for (int j = 0; j < NU; j++) torque[j] = -kp[6+j]*err[6+j] - kd[6+j]*(d->qvel[6+j] + rhs[6+j]*dt);
}
static void or_rfc_implicit(OrEnv *e, const double *ctrl) { /* humanoid_im.py:1136-1143 */
    double vf[6], crq[4], hq[4], R[9], t[3];
    for (int i = 0; i < 6; i++) vf[i] = ctrl[NU+i]*e->rfc_scale*e->rfc_rate;
    remove_base_rot(e, e->d->qpos+3, crq); heading_q(crq, hq); q2mat(hq, R); mv(R, vf, t); memcpy(vf, t, 24);
    for (int i = 0; i < 6; i++) { if (vf[i] > e->rfc_lim) vf[i] = e->rfc_lim; if (vf[i] < -e->rfc_lim) vf[i] = -e->rfc_lim; e->d->qfrc_applied[i] = vf[i]; }
}

/* mj_applyFT restated: qfrc += J(point, body)^T [force; torque].  The reference calls it between two sim.step() (humanoid_im.py:1122-1130), so the
   Jacobian is built from the kinematics of the LAST forward pass (data.xpos / joint axes), not from the integrated qpos. */
void or_apply_ft(const OrModel *m, const OrData *d, const double *force, const double *torque, const double *point, int body, double *qfrc) {
    static double Jv[3][NV], Jw[3][NV];
    or_jac(m, d, body, point, Jv, Jw);
    for (int i = 0; i < NV; i++)
        qfrc[i] += Jv[0][i]*force[0] + Jv[1][i]*force[1] + Jv[2][i]*force[2] + Jw[0][i]*torque[0] + Jw[1][i]*torque[1] + Jw[2][i]*torque[2];
}
/* humanoid_im.py:1080-1132 with the release settings (no contact gating: residual_contact_only = False; no projection; one point per body):
   per body a contact point, a force and a torque in the BODY frame of the last forward pass (mujoco_env.py:171-180), scaled by
   residual_force_scale, applied through mj_applyFT; qfrc_applied is replaced, not accumulated. */
static void or_rfc_explicit(OrEnv *e, const double *ctrl) {
    OrData *d = e->d; double qfrc[NV] = {0};
    for (int i = 0; i < NB; i++) {
        const int b = e->vf_body[i]; const double *v = ctrl + NU + 9*i, *R = d->xmat[b];
        double f[3] = {v[3]*e->rfc_scale, v[4]*e->rfc_scale, v[5]*e->rfc_scale}, tq[3] = {v[6]*e->rfc_scale, v[7]*e->rfc_scale, v[8]*e->rfc_scale};
        double p[3], fw[3], tw[3];
        mv(R, v, p); mv(R, f, fw); mv(R, tq, tw);
        for (int k = 0; k < 3; k++) p[k] += d->xpos[b][k];
        or_apply_ft(e->m, d, fw, tw, p, b, qfrc);
    }
    memcpy(d->qfrc_applied, qfrc, sizeof qfrc);
}

/* get_full_obs_v2 (humanoid_im.py:419-503) and get_full_obs_v1 (:323-417), obs_coord = "root", obs_vel = "full".  v1 = v2 with two more blocks
   (per-body COM relative to the root, COM difference to the expert's body_com) between the joint-position blocks and the quaternions, and no shape vector. */
static void or_obs_block(const OrEnv *e, double *obs);
static void or_obs_v56(const OrEnv *e, double *obs);
void or_obs_v2(const OrEnv *ec, double *obs) {     /* get_obs: one block, or the v3 stack of v2 blocks */
    OrEnv *e = (OrEnv *)ec;
    if (e->obs_v == 5 || e->obs_v == 6) { or_obs_v56(e, obs); return; }
    if (e->obs_v != 3) { e->obs_dt = 0; or_obs_block(e, obs); return; }
    for (int f = 0; f < e->fut_frames; f++) { e->obs_dt = f*e->fut_skip; or_obs_block(e, obs + obs_block_dim(e)*f); }
    e->obs_dt = 0;
}
static void or_obs_block(const OrEnv *e, double *obs) {
    const OrData *d = e->d; double qpos[NQ], qvel[NV], R[9], t[3];
    memcpy(qpos, d->qpos, sizeof qpos); memcpy(qvel, d->qvel, sizeof qvel);
    q2mat(qpos+3, R); mtv(R, qvel, t); memcpy(qvel, t, 24);                       /* :425 */
    double crq[4], hq[4], hqi[4];
    remove_base_rot(e, qpos+3, crq); heading_q(crq, hq);
    int o = 0; memcpy(obs+o, hq, 32); o += 4;
    int ind = ex_index(e, e->cur_t + 1 + e->obs_dt);
    const double *tq = e->ex.qpos + NQ*ind, *twq = e->ex.wbquat + 96*ind, *tjp = e->ex.wbpos + 72*ind;
    double trq[4], dh[4], diff[NQ], ci[4];
    remove_base_rot(e, tq+3, trq);
    qinv(hq, hqi); qmul(hqi, crq, dh); memcpy(qpos+3, dh, 32);                    /* de_heading(curr_root_quat) :440 */
    memcpy(diff, tq, sizeof diff); diff[2] -= qpos[2];
    for (int i = 7; i < NQ; i++) diff[i] -= qpos[i];
    qinv(crq, ci); qmul(trq, ci, diff+3);
    memcpy(obs+o, tq+2, 74*8); o += 74; memcpy(obs+o, qpos+2, 74*8); o += 74; memcpy(obs+o, diff+2, 74*8); o += 74;
    q2mat(crq, R); mtv(R, qvel, t); memcpy(qvel, t, 24);                          /* second rotation :451 */
    memcpy(obs+o, qvel, 75*8); o += 75;
    double rel_h = heading(trq) - heading(crq);
    if (rel_h > M_PI) rel_h -= 2*M_PI;
    if (rel_h < -M_PI) rel_h += 2*M_PI;
    obs[o++] = rel_h;
    double rp[3] = {trq[0]-qpos[0], trq[1]-qpos[1], trq[2]-qpos[2]};             /* the kept bug :466 */
    mtv(R, rp, t); obs[o++] = t[0]; obs[o++] = t[1];
    /* joint positions rel. root, in root frame, laid out x-block / y-block / z-block (transform_vec_batch returns (3,N)) */
    for (int b = 0; b < NB; b++) { double r[3] = {d->xpos[b][0]-qpos[0], d->xpos[b][1]-qpos[1], d->xpos[b][2]-qpos[2]}; mtv(R, r, t); for (int k = 0; k < 3; k++) obs[o+24*k+b] = t[k]; }
    o += 72;
    for (int b = 0; b < NB; b++) { double r[3] = {tjp[3*b]-d->xpos[b][0], tjp[3*b+1]-d->xpos[b][1], tjp[3*b+2]-d->xpos[b][2]}; mtv(R, r, t); for (int k = 0; k < 3; k++) obs[o+24*k+b] = t[k]; }
    o += 72;
    if (e->obs_v == 1) {                                                           /* :383-393 */
        const double *tcom = e->ex.body_com + 72*ind;
        for (int b = 0; b < NB; b++) { double r[3] = {d->xipos[b][0]-qpos[0], d->xipos[b][1]-qpos[1], d->xipos[b][2]-qpos[2]}; mtv(R, r, t); for (int k = 0; k < 3; k++) obs[o+24*k+b] = t[k]; }
        o += 72;
        for (int b = 0; b < NB; b++) { double r[3] = {tcom[3*b]-d->xipos[b][0], tcom[3*b+1]-d->xipos[b][1], tcom[3*b+2]-d->xipos[b][2]}; mtv(R, r, t); for (int k = 0; k < 3; k++) obs[o+24*k+b] = t[k]; }
        o += 72;
    }
    int use_target = (d->xquat[0][0] == 0);                                        /* :485 */
    for (int b = 0; b < NB; b++) { const double *cq = use_target ? twq+4*b : d->xquat[b]; qmul(hqi, cq, obs+o+4*b); }
    o += 96;
    for (int b = 0; b < NB; b++) { const double *cq = use_target ? twq+4*b : d->xquat[b]; double iq[4]; qinv(cq, iq); double n = e->obs_v == 1 ? 1.0 : sqrt(cq[0]*cq[0]+cq[1]*cq[1]+cq[2]*cq[2]+cq[3]*cq[3]); for (int k = 0; k < 4; k++) iq[k] *= n; /* v2: inverse_batch divides by |q|, not |q|^2; v1: quaternion_inverse (:411) */ qmul(iq, twq+4*b, obs+o+4*b); }
    o += 96;
    if (e->obs_v != 1 && !e->no_shape) { memcpy(obs+o, e->ex.shape_obs, 17*8); o += 17; }
}

/* the "_new" heading helpers (uhc/utils/math_utils.py:169-207): yaw from the full quaternion, heading quaternion about z */
static double heading_new(const double *q) { return atan2(2*(q[0]*q[3] + q[1]*q[2]), 1 - 2*(q[2]*q[2] + q[3]*q[3])); }
static void heading_q_new(const double *q, double *hq) { double y = heading_new(q); hq[0] = cos(y/2); hq[1] = 0; hq[2] = 0; hq[3] = sin(y/2); }
/* get_full_obs_v5 (humanoid_im.py:505-594: "no diff, no heading" ablation of v2 on the _new heading helpers, velocity rotated ONCE) and get_full_obs_v6
   (:596-666: the concise one -- root offset / heading / relative root rotation, qvel, joint positions and their differences in the heading frame, local joint
   quaternions and their differences).  obs_coord = "root", obs_vel = "full".  Kept as the reference computes them: in v6 `transform_vec_batch_new(...)[1:]` slices
   the (3, 24) result, i.e. drops the x ROW (48 values remain), while the difference block drops the root BODY before the transform (3 x 23 = 69 values). */
static void or_obs_v56(const OrEnv *e, double *obs) {
    const OrData *d = e->d; double qpos[NQ], qvel[NV], R[9], t[3];
    memcpy(qpos, d->qpos, sizeof qpos); memcpy(qvel, d->qvel, sizeof qvel);
    const int ind = ex_index(e, e->cur_t + 1);
    const double *tq = e->ex.qpos + NQ*ind, *twq = e->ex.wbquat + 96*ind, *tjp = e->ex.wbpos + 72*ind, *tbq = e->ex.bquat + 96*ind;
    double crq[4], trq[4], hq[4], hqi[4], ci[4], relq[4];
    remove_base_rot(e, qpos+3, crq); remove_base_rot(e, tq+3, trq); heading_q_new(crq, hq); qinv(hq, hqi);
    qinv(crq, ci); qmul(trq, ci, relq);
    double rel_h = heading_new(trq) - heading_new(crq);
    if (rel_h > M_PI) rel_h -= 2*M_PI;
    if (rel_h < -M_PI) rel_h += 2*M_PI;
    int o = 0;
    if (e->obs_v == 5) {
        double dh[4], diff[NQ];
        qmul(hqi, crq, dh);                                                          /* de_heading_new(curr_root_quat) :528 */
        memcpy(diff, tq, sizeof diff); diff[2] -= qpos[2];
        memcpy(qpos+3, dh, 32);
        for (int i = 7; i < NQ; i++) diff[i] -= qpos[i];
        memcpy(diff+3, relq, 32);
        memcpy(obs+o, tq+2, 74*8); o += 74; memcpy(obs+o, qpos+2, 74*8); o += 74; memcpy(obs+o, diff+2, 74*8); o += 74;
        q2mat(crq, R); mtv(R, qvel, t); memcpy(qvel, t, 24);                          /* transform_vec_new(qvel[:3], curr_root_quat, "root") = v . R :540 */
        memcpy(obs+o, qvel, 75*8); o += 75;
        obs[o++] = rel_h;
        double rp[3] = {tq[0]-qpos[0], tq[1]-qpos[1], tq[2]-qpos[2]};
        mtv(R, rp, t); obs[o++] = t[0]; obs[o++] = t[1];
        for (int b = 0; b < NB; b++) { double r[3] = {d->xpos[b][0]-qpos[0], d->xpos[b][1]-qpos[1], d->xpos[b][2]-qpos[2]}; mtv(R, r, t); for (int k = 0; k < 3; k++) obs[o+24*k+b] = t[k]; }
        o += 72;
        for (int b = 0; b < NB; b++) { double r[3] = {tjp[3*b]-d->xpos[b][0], tjp[3*b+1]-d->xpos[b][1], tjp[3*b+2]-d->xpos[b][2]}; mtv(R, r, t); for (int k = 0; k < 3; k++) obs[o+24*k+b] = t[k]; }
        o += 72;
        int use_target = (d->xquat[0][0] == 0);
        for (int b = 0; b < NB; b++) { const double *cq = use_target ? twq+4*b : d->xquat[b]; qmul(hqi, cq, obs+o+4*b); }
        o += 96;
        for (int b = 0; b < NB; b++) { const double *cq = use_target ? twq+4*b : d->xquat[b]; double iq[4]; qinv(cq, iq); double n = sqrt(cq[0]*cq[0]+cq[1]*cq[1]+cq[2]*cq[2]+cq[3]*cq[3]); for (int k = 0; k < 4; k++) iq[k] *= n; qmul(iq, twq+4*b, obs+o+4*b); }
        o += 96;
    } else {
        q2mat(hq, R);
        double rp[3] = {tq[0]-qpos[0], tq[1]-qpos[1], tq[2]-qpos[2]};
        mtv(R, rp, t); memcpy(obs+o, t, 24); o += 3;                                  /* :623-626 */
        obs[o++] = rel_h;
        memcpy(obs+o, relq, 32); o += 4;
        mtv(R, qvel, t); memcpy(qvel, t, 24);
        memcpy(obs+o, qvel, 75*8); o += 75;
        for (int b = 0; b < NB; b++) { double r[3] = {d->xpos[b][0]-qpos[0], d->xpos[b][1]-qpos[1], d->xpos[b][2]-qpos[2]}; mtv(R, r, t); for (int k = 1; k < 3; k++) obs[o+24*(k-1)+b] = t[k]; }
        o += 48;                                                                      /* [1:] of the (3, 24) array: the y and z rows (:645) */
        for (int b = 1; b < NB; b++) { double r[3] = {tjp[3*b]-d->xpos[b][0], tjp[3*b+1]-d->xpos[b][1], tjp[3*b+2]-d->xpos[b][2]}; mtv(R, r, t); for (int k = 0; k < 3; k++) obs[o+23*k+(b-1)] = t[k]; }
        o += 69;
        double bq[96]; or_body_quat(e, bq);
        memcpy(obs+o, bq+4, 92*8); o += 92;
        for (int b = 1; b < NB; b++) { const double *cq = bq+4*b; double iq[4]; qinv(cq, iq); double n = sqrt(cq[0]*cq[0]+cq[1]*cq[1]+cq[2]*cq[2]+cq[3]*cq[3]); for (int k = 0; k < 4; k++) iq[k] *= n; qmul(iq, tbq+4*b, obs+o+4*(b-1)); }
        o += 92;
    }
    if (!e->no_shape) { memcpy(obs+o, e->ex.shape_obs, 17*8); o += 17; }
}

static void rot_from_quat(const double *q, double *rv) { /* transformation.py:362-372 */
    if (fabs(1.0 - q[0]) < 1e-6 || fabs(1 + q[0]) < 1e-6) { rv[0] = rv[1] = rv[2] = 0; return; }
    double ang = 2*acos(q[0]), s = sin(ang/2), ax[3] = {q[1]/s, q[2]/s, q[3]/s}, n = sqrt(dot3(ax, ax));
    for (int k = 0; k < 3; k++) rv[k] = ax[k]/n*ang;
}
double or_reward(const OrEnv *e, const double *action, double *cinfo) { /* reward_function.py:12-88 ; call after or_env_step */
    const OrData *d = e->d; int ind = ex_index(e, e->cur_t);
    const double *e_ee = e->ex.ee_wpos + 15*ind, *e_com = e->ex.com + 3*ind, *e_bq = e->ex.bquat + 96*ind, *e_bav = e->ex.bangvel + 72*ind;
    double cur_bq[96]; or_body_quat(e, cur_bq);
    double dt = e->m->dt*15, pose2 = 0, vel2 = 0, ee2 = 0, com2 = 0, vf2 = 0;
    for (int b = 0; b < NB; b++) {
        double iq[4], dq[4], rv[3], w = (b == 0) ? 1.0 : e->m->diffw[b];
        qinv(e_bq+4*b, iq); qmul(cur_bq+4*b, iq, dq);
        double c = dq[0]; if (c > 1) c = 1; if (c < -1) c = -1;
        double a = acos(c)*w; pose2 += a*a;
        qinv(e->prev_bquat+4*b, iq); qmul(cur_bq+4*b, iq, dq); rot_from_quat(dq, rv);
        /* world_rfc_implicit weights the angular-velocity error by jpos_diffw (reward_function.py:50-51); world_rfc_explicit does not (:253-341) */
        for (int k = 0; k < 3; k++) { double dv = (rv[k]/dt - e_bav[3*b+k])*(e->rfc_mode == 1 ? 1.0 : e->m->diffw[b]); vel2 += dv*dv; }
    }
    for (int i = 0; i < 5; i++) for (int k = 0; k < 3; k++) { double x = d->xpos[e->m->ee[i]][k] - e_ee[3*i+k]; ee2 += x*x; }
    for (int k = 0; k < 3; k++) { double x = d->xipos[0][k] - e_com[k]; com2 += x*x; }
    if (e->rfc_mode == 0) for (int i = 0; i < 6; i++) vf2 += action[NU+i]*action[NU+i];
    else if (e->rfc_mode == 1) for (int i = 0; i < NB; i++) for (int k = 3; k < 9; k++) vf2 += action[NU+9*i+k]*action[NU+9*i+k];   /* force + torque part of every body's slot (:321-327) */
    cinfo[0] = exp(-e->k[0]*pose2); cinfo[1] = exp(-e->k[1]*vel2); cinfo[2] = exp(-e->k[2]*ee2); cinfo[3] = exp(-e->k[3]*com2); cinfo[4] = e->rfc_mode == 2 ? 0.0 : exp(-e->k[4]*vf2);   /* residual_force off: vf_reward = 0.0 (reward_function.py:68-72) */
    if (e->reward_mul) return cinfo[0]*cinfo[1]*cinfo[2]*cinfo[3]*(e->w[4] != 0.0 ? cinfo[4] : 1.0);   /* :243-245 */
    double r = 0, ws = 0; for (int i = 0; i < 5; i++) { r += e->w[i]*cinfo[i]; ws += e->w[i]; }
    return r/ws;
}

/* reset: sim.reset() (warm start cleared), state <- expert frame 0, sim.forward(): mujoco_env.py:95-113, humanoid_im.py:1245 */
void or_env_reset(OrEnv *e, const double *qpos_override, const double *qvel_override, double *obs) {
    OrData *d = e->d;
    memset(d, 0, sizeof *d);
    e->cur_t = 0; e->start_ind = 0;
    d->qpos[3] = 1; /* sim.reset() -> qpos0: identity root quat, zero hinge angles */
    or_body_quat(e, e->bquat); /* reset_model reads bquat BEFORE set_state (humanoid_im.py:1277), i.e. at qpos0 */
    memcpy(d->qpos, qpos_override ? qpos_override : e->ex.qpos, sizeof d->qpos);
    memcpy(d->qvel, qvel_override ? qvel_override : e->ex.qvel, sizeof d->qvel);
    or_forward(e->m, d);
    if (obs) or_obs_v2(e, obs);
}
void or_env_set_bquat(OrEnv *e, const double *bq) { memcpy(e->bquat, bq, sizeof e->bquat); }

/* one 30 Hz control step; returns done, fills fail/end/percent. humanoid_im.py:1192-1243 */
int or_env_step(OrEnv *e, const double *action, double *obs, int *fail, int *end, double *percent) {
    OrData *d = e->d;
    memcpy(e->prev_bquat, e->bquat, sizeof e->bquat);
    for (int i = 0; i < 15; i++) {
        double tq[NU];
        or_compute_torque(e, action, i, tq);
        for (int j = 0; j < NU; j++) { if (tq[j] > e->m->tlim[j]) tq[j] = e->m->tlim[j]; if (tq[j] < -e->m->tlim[j]) tq[j] = -e->m->tlim[j]; d->ctrl[j] = tq[j]; e->torque[i][j] = tq[j]; }
        if (e->rfc_mode == 1) or_rfc_explicit(e, action); else if (e->rfc_mode == 0) or_rfc_implicit(e, action);
        or_step(e->m, d);
    }
    e->cur_t += 1;
    or_body_quat(e, e->bquat);
    double bd = or_body_diff(e);
    if (e->term_body == 1) *fail = d->qpos[2] < e->height_lb - 0.1;                              /* humanoid_im.py:1226 */
    else if (e->term_body == 2) *fail = d->xpos[e->head_body][2] < e->head_height_lb - 0.1;     /* :1221-1224: data.body_xpos of the last forward pass */
    else *fail = bd > e->body_diff_thresh;
    for (int i = 0; i < NQ; i++) if (!isfinite(d->qpos[i])) *fail = 1;
    *end = (e->cur_t >= e->env_episode_len) || (e->cur_t + e->start_ind >= e->ex.len + e->trail_steps - 1);
    *percent = (double)e->cur_t/(e->ex.len - 1);
    if (obs) or_obs_v2(e, obs);
    return *fail || *end;
}

/* kinetic + potential energy and momentum, for the invariant tests */
double or_energy(const OrModel *m, OrData *d, double *mom) {
    or_kinematics(m, d); or_mass_matrix(m, d);
    double ke = 0, pe = 0;
    for (int i = 0; i < NV; i++) for (int j = 0; j < NV; j++) ke += 0.5*d->qvel[i]*d->M[i*NV+j]*d->qvel[j];
    for (int b = 0; b < NB; b++) pe -= m->mass[b]*dot3(m->gravity, d->xipos[b]);
    if (mom) { static double Jv[3][NV], Jw[3][NV]; mom[0] = mom[1] = mom[2] = 0;
        for (int b = 0; b < NB; b++) { or_jac(m, d, b, d->xipos[b], Jv, Jw); for (int k = 0; k < 3; k++) for (int i = 0; i < NV; i++) mom[k] += m->mass[b]*Jv[k][i]*d->qvel[i]; } }
    return ke + pe;
}
