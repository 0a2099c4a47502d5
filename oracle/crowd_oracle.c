/*
 * oracle/crowd_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 * See crowd_oracle.h.  fp64 restatement of the reference's Python crowd step.
 */
#include "crowd_oracle.h"
#include "rvo2_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* Python-scalar arithmetic (`np.sqrt(dx**2 + dy**2)`): two products, one sum, each rounded. */
static double norm2(double x, double y) { return sqrt(x * x + y * y); }
/* numpy's `np.dot` / `@` / `np.linalg.norm` on float64 vectors go through the BLAS ddot kernel, which
 * accumulates with fused multiply-adds: acc = x0*y0; acc = fma(x1, y1, acc)  [probe: 20000/20000 random
 * vectors reproduce bit-for-bit, 83 % with the unfused form].  The reference's near-parallel branch
 * (`if not denom`, utils_plus.py:231) depends on that last bit, so the oracle mirrors it. */
static double npdot2(double x0, double x1, double y0, double y1) { return fma(x1, y1, x0 * y0); }
static double npnorm2(double x, double y) { return sqrt(fma(y, y, x * x)); }

/* ------------------------------------------------------------------------- */
/* crowd_sim_plus/envs/policy/orca.py:82-133 and orca_plus.py:29-90          */
static __thread OrcProbe *g_probe = NULL;

void orc_orca_predict(const OrcPolicyCfg *cfg, const double *self8, int n_others, const double *others,
                      int n_seg, const double *segs, double *out_v2, int *nbr_ids, int *n_nbr,
                      int *obst_nbr_ids, int *n_obst_nbr)
{
    const double px = self8[0], py = self8[1], vx = self8[2], vy = self8[3], radius = self8[4];
    const double gx = self8[5], gy = self8[6], v_pref = self8[7];
    const float nd = (float)cfg->neighbor_dist, th = (float)cfg->time_horizon, tho = (float)cfg->time_horizon_obst;
    RvoSim *sim = rvo_create((float)cfg->time_step, nd, cfg->max_neighbors, th, tho, (float)cfg->policy_radius,
                             (float)cfg->max_speed, 0.0f, 0.0f);
    if (cfg->policy == ORC_POLICY_ORCA_PLUS && n_seg > 0) { /* orca_plus.py:50-53 */
        for (int k = 0; k < n_seg; ++k) {
            float xy[4] = { (float)segs[4 * k], (float)segs[4 * k + 1], (float)segs[4 * k + 2], (float)segs[4 * k + 3] };
            rvo_add_obstacle(sim, xy, 2);
        }
        rvo_process_obstacles(sim);
    }
    /* orca.py:100-104: agent 0 = self, radius + 0.01 + safety_space evaluated in double then narrowed */
    rvo_add_agent(sim, (float)px, (float)py, nd, cfg->max_neighbors, th, tho,
                  (float)(radius + 0.01 + cfg->safety_space), (float)v_pref, (float)vx, (float)vy);
    for (int j = 0; j < n_others; ++j) {
        const double *o = others + 5 * j;
        rvo_add_agent(sim, (float)o[0], (float)o[1], nd, cfg->max_neighbors, th, tho,
                      (float)(o[4] + 0.01 + cfg->safety_space), (float)cfg->max_speed, (float)o[2], (float)o[3]);
    }
    /* preferred velocity: orca.py:113-115 / orca_plus.py:68-71 */
    const double dvx = gx - px, dvy = gy - py;
    const double speed = npnorm2(dvx, dvy); /* np.linalg.norm(velocity) */
    double pvx, pvy;
    if (cfg->policy == ORC_POLICY_ORCA_PLUS) {
        const double epsilon = 1e-3;
        if (speed > (v_pref - epsilon)) { pvx = dvx / speed * (v_pref - epsilon); pvy = dvy / speed * (v_pref - epsilon); }
        else { pvx = dvx; pvy = dvy; }
    } else {
        if (speed > 1) { pvx = dvx / speed; pvy = dvy / speed; }
        else { pvx = dvx; pvy = dvy; }
    }
    rvo_set_agent_pref_velocity(sim, 0, (float)pvx, (float)pvy);
    for (int j = 0; j < n_others; ++j) rvo_set_agent_pref_velocity(sim, j + 1, 0.0f, 0.0f);
    rvo_do_step(sim);
    float v[2];
    rvo_get_agent_velocity(sim, 0, v);
    out_v2[0] = (double)v[0]; out_v2[1] = (double)v[1];
    if (g_probe) { /* introspection for tests/test_orca_properties.py: agent 0's half-planes as RVO2 built them */
        OrcProbe *pr = g_probe;
        pr->n_lines = rvo_get_agent_num_orca_lines(sim, 0);
        pr->n_obst_lines = pr->n_lines - rvo_get_agent_num_agent_neighbors(sim, 0);
        for (int k = 0; k < pr->n_lines && k < ORC_PROBE_MAX_LINES; ++k) rvo_get_agent_orca_line(sim, 0, k, pr->lines + 4 * k);
        float pv[2];
        rvo_get_agent_pref_velocity(sim, 0, pv);
        pr->pref[0] = pv[0]; pr->pref[1] = pv[1];
        pr->max_speed = rvo_get_agent_max_speed(sim, 0);
    }
    if (n_nbr) {
        *n_nbr = rvo_get_agent_num_agent_neighbors(sim, 0);
        for (int k = 0; k < *n_nbr; ++k) nbr_ids[k] = rvo_get_agent_agent_neighbor(sim, 0, k) - 1; /* ob index */
    }
    if (n_obst_nbr) {
        *n_obst_nbr = rvo_get_agent_num_obstacle_neighbors(sim, 0);
        for (int k = 0; k < *n_obst_nbr; ++k) obst_nbr_ids[k] = rvo_get_agent_obstacle_neighbor(sim, 0, k);
    }
    rvo_destroy(sim);
}

/* ------------------------------------------------------------------------- */
/* crowd_sim_plus/envs/utils/utils_plus.py:21-42 */
void orc_closest_point_on_segment(double x1, double y1, double x2, double y2, double x3, double y3, double *out2)
{
    const double px = x2 - x1, py = y2 - y1;
    if (px == 0 && py == 0) { out2[0] = x1; out2[1] = y1; return; } /* reference returns a scalar here (quirk q12) */
    double u = ((x3 - x1) * px + (y3 - y1) * py) / (px * px + py * py);
    if (u > 1) u = 1; else if (u < 0) u = 0;
    out2[0] = x1 + u * px; out2[1] = y1 + u * py;
}

/* utils_plus.py:44-65 */
static void closest_point_on_segment_extended(double x1, double y1, double x2, double y2, double x3, double y3, double *out2)
{
    const double px = x2 - x1, py = y2 - y1;
    if (px == 0 && py == 0) { out2[0] = x1; out2[1] = y1; return; }
    const double u = ((x3 - x1) * px + (y3 - y1) * py) / (px * px + py * py);
    out2[0] = x1 + u * px; out2[1] = y1 + u * py;
}

/* utils_plus.py:73-95 */
double orc_point_to_segment_dist(double x1, double y1, double x2, double y2, double x3, double y3)
{
    const double px = x2 - x1, py = y2 - y1;
    if (px == 0 && py == 0) return npnorm2(x3 - x1, y3 - y1);
    double u = ((x3 - x1) * px + (y3 - y1) * py) / (px * px + py * py);
    if (u > 1) u = 1; else if (u < 0) u = 0;
    const double x = x1 + u * px, y = y1 + u * py;
    return npnorm2(x - x3, y - y3);
}

/* utils_plus.py:6-18 */
static void intersection_of_vec_line_and_2p_line(double vox, double voy, double vx, double vy, double x1, double y1,
                                                 double x2, double y2, double *out2)
{
    const double x3 = vox, y3 = voy, x4 = vox + vx, y4 = voy + vy;
    out2[0] = ((x1 * y2 - y1 * x2) * (x3 - x4) - (x1 - x2) * (x3 * y4 - y3 * x4)) /
              ((x1 - x2) * (y3 - y4) - (y1 - y2) * (x3 - x4));
    out2[1] = ((x1 * y2 - y1 * x2) * (y3 - y4) - (y1 - y2) * (x3 * y4 - y3 * x4)) /
              ((x1 - x2) * (y3 - y4) - (y1 - y2) * (x3 - x4));
}

/* utils_plus.py:205-338 with the z=0 component of the reference's 3-vectors dropped */
void orc_closest_distance_between_line_segments(const double *a0, const double *a1_in, const double *b0,
                                                const double *b1_in, double *out5)
{
    double a1[2] = { a1_in[0], a1_in[1] }, b1[2] = { b1_in[0], b1_in[1] };
    double A[2] = { a1[0] - a0[0], a1[1] - a0[1] }, B[2] = { b1[0] - b0[0], b1[1] - b0[1] };
    const double magA = npnorm2(A[0], A[1]), magB = npnorm2(B[0], B[1]);
    double _A[2], _B[2];
    if (magA < 1e-8) { a1[0] = a0[0]; a1[1] = a0[1]; A[0] = A[1] = _A[0] = _A[1] = 0.0; }
    else { _A[0] = A[0] / magA; _A[1] = A[1] / magA; }
    if (magB < 1e-8) { b1[0] = b0[0]; b1[1] = b0[1]; B[0] = B[1] = _B[0] = _B[1] = 0.0; }
    else { _B[0] = B[0] / magB; _B[1] = B[1] / magB; }

    const double cz = _A[0] * _B[1] - _A[1] * _B[0];
    const double ncross = sqrt(cz * cz);
    const double denom = ncross * ncross;
    double pA[2], pB[2];
#define RET(PA, PB) do { out5[0] = (PA)[0]; out5[1] = (PA)[1]; out5[2] = (PB)[0]; out5[3] = (PB)[1]; \
                         out5[4] = npnorm2((PA)[0] - (PB)[0], (PA)[1] - (PB)[1]); return; } while (0)
    if (denom == 0.0) {
        const double d0 = npdot2(_A[0], _A[1], b0[0] - a0[0], b0[1] - a0[1]);
        const double d1 = npdot2(_A[0], _A[1], b1[0] - a0[0], b1[1] - a0[1]);
        if (d0 <= 0 && 0 >= d1) {
            if (fabs(d0) < fabs(d1)) RET(a0, b0);
            RET(a0, b1);
        } else if (d0 >= magA && magA <= d1) {
            if (fabs(d0) < fabs(d1)) RET(a1, b0);
            RET(a1, b1);
        } else {
            double a0f[2], _Af[2];
            if (npnorm2(_A[0] - _B[0], _A[1] - _B[1]) < 1e-8 || magB < 1e-8) {
                a0f[0] = a0[0]; a0f[1] = a0[1]; _Af[0] = _A[0]; _Af[1] = _A[1];
            } else {
                a0f[0] = a1[0]; a0f[1] = a1[1]; _Af[0] = -_A[0]; _Af[1] = -_A[1];
            }
            const double d0f = npdot2(_Af[0], _Af[1], b0[0] - a0f[0], b0[1] - a0f[1]);
            if (d0f >= 0) {
                pB[0] = b0[0]; pB[1] = b0[1];
                const double t = npdot2(_Af[0], _Af[1], pB[0] - a0f[0], pB[1] - a0f[1]);
                pA[0] = a0f[0] + _Af[0] * t; pA[1] = a0f[1] + _Af[1] * t;
            } else {
                pA[0] = a0f[0]; pA[1] = a0f[1];
                const double t = npdot2(_B[0], _B[1], pA[0] - b0[0], pA[1] - b0[1]);
                pB[0] = b0[0] + _B[0] * t; pB[1] = b0[1] + _B[1] * t;
            }
            RET(pA, pB);
        }
    }
    /* lines criss-cross: det([t,_B,cross]) = cz * (t x _B) for z=0 vectors */
    const double t[2] = { b0[0] - a0[0], b0[1] - a0[1] };
    const double detA = cz * (t[0] * _B[1] - t[1] * _B[0]);
    const double detB = cz * (t[0] * _A[1] - t[1] * _A[0]);
    const double t0 = detA / denom, t1 = detB / denom;
    pA[0] = a0[0] + _A[0] * t0; pA[1] = a0[1] + _A[1] * t0;
    pB[0] = b0[0] + _B[0] * t1; pB[1] = b0[1] + _B[1] * t1;
    if (t0 < 0) { pA[0] = a0[0]; pA[1] = a0[1]; } else if (t0 > magA) { pA[0] = a1[0]; pA[1] = a1[1]; }
    if (t1 < 0) { pB[0] = b0[0]; pB[1] = b0[1]; } else if (t1 > magB) { pB[0] = b1[0]; pB[1] = b1[1]; }
    if (t0 < 0 || t0 > magA) {
        double dot = npdot2(_B[0], _B[1], pA[0] - b0[0], pA[1] - b0[1]);
        if (dot < 0) dot = 0; else if (dot > magB) dot = magB;
        pB[0] = b0[0] + _B[0] * dot; pB[1] = b0[1] + _B[1] * dot;
    }
    if (t1 < 0 || t1 > magB) {
        double dot = npdot2(_A[0], _A[1], pB[0] - a0[0], pB[1] - a0[1]);
        if (dot < 0) dot = 0; else if (dot > magA) dot = magA;
        pA[0] = a0[0] + _A[0] * dot; pA[1] = a0[1] + _A[1] * dot;
    }
    RET(pA, pB);
#undef RET
}

/* Agent.compute_position, crowd_sim_plus/envs/utils/agent_plus.py:175-185 */
static void compute_position(const double *pose3, int kin, const double *action2, double dt, double *out2)
{
    if (kin == ORC_KIN_HOLONOMIC) {
        out2[0] = pose3[0] + action2[0] * dt;
        out2[1] = pose3[1] + action2[1] * dt;
    } else {
        const double theta = pose3[2] + action2[1];
        out2[0] = pose3[0] + cos(theta) * action2[0] * dt;
        out2[1] = pose3[1] + sin(theta) * action2[0] * dt;
    }
}

/* CrowdSimPlus.constrain_agent_action_exact, crowd_sim_plus/envs/crowd_sim_plus.py:869-989 */
void orc_constrain_action(const double *pose3, double r, double dt, int kin, const double *action2, int n_seg,
                          const double *segs, double *out_action2)
{
    const double cur[2] = { pose3[0], pose3[1] };
    double fut[2];
    compute_position(pose3, kin, action2, dt, fut);
    const double mdir[2] = { fut[0] - cur[0], fut[1] - cur[1] };
    const double movement_mag = npnorm2(mdir[0], mdir[1]);
    double fin[2] = { action2[0], action2[1] };

    for (int k = 0; k < n_seg; ++k) {
        const double *L = segs + 4 * k;
        double cd[5];
        orc_closest_distance_between_line_segments(L, L + 2, cur, fut, cd);
        const double *pA = cd, *pB = cd + 2;
        const double closest_distance = cd[4];
        if (!(closest_distance - r < 0.0)) continue;

        double final_position[2];
        if ((npnorm2(pA[0] - L[0], pA[1] - L[1]) < 1e-8 || npnorm2(pA[0] - L[2], pA[1] - L[3]) < 1e-8) &&
            npnorm2(pA[0] - pB[0], pA[1] - pB[1]) > 1e-8) {
            /* collision with an end-point of the segment (:904-947) */
            const double dvec[2] = { pB[0] - cur[0], pB[1] - cur[1] };
            const double dir_mag = npnorm2(dvec[0], dvec[1]);
            double _d[2], redux;
            if (dir_mag > 0.0 && npnorm2(pA[0] - cur[0], pA[1] - cur[1]) - r < 1e-4 &&
                npdot2(mdir[0], mdir[1], pA[0] - cur[0], pA[1] - cur[1]) > -1e-8) {
                _d[0] = dvec[0] / dir_mag; _d[1] = dvec[1] / dir_mag;
                redux = dir_mag;
            } else if (dir_mag > 0.0) {
                _d[0] = dvec[0] / dir_mag; _d[1] = dvec[1] / dir_mag;
                const double arccos_value = npdot2(-dvec[0], -dvec[1], pA[0] - pB[0], pA[1] - pB[1]) / (dir_mag * closest_distance);
                const double clipped = arccos_value < -1.0 ? -1.0 : (arccos_value > 1.0 ? 1.0 : arccos_value);
                const double alpha = acos(clipped);
                if (alpha == M_PI) {
                    redux = r - closest_distance;
                } else {
                    const double gamma = asin(closest_distance * sin(alpha) / r);
                    const double beta = M_PI - alpha - gamma;
                    redux = r * sin(beta) / sin(alpha) + 1e-7;
                }
            } else {
                redux = 0.0; _d[0] = dvec[0]; _d[1] = dvec[1];
            }
            const double m = (dir_mag - redux) > 0 ? (dir_mag - redux) : 0;
            final_position[0] = cur[0] + _d[0] * m; final_position[1] = cur[1] + _d[1] * m;
        } else {
            /* collision with the interior: constrain against the infinite line (:948-967) */
            double cl[2];
            closest_point_on_segment_extended(L[0], L[1], L[2], L[3], cur[0], cur[1], cl);
            if (movement_mag > 0.0 && npnorm2(cl[0] - cur[0], cl[1] - cur[1]) - r < 1e-4 &&
                npdot2(mdir[0], mdir[1], cl[0] - cur[0], cl[1] - cur[1]) > -1e-8) {
                final_position[0] = cur[0]; final_position[1] = cur[1];
            } else if (movement_mag > 0.0) {
                double in[2];
                intersection_of_vec_line_and_2p_line(cur[0], cur[1], mdir[0], mdir[1], L[0], L[1], L[2], L[3], in);
                const double d_vec[2] = { in[0] - cur[0], in[1] - cur[1] };
                const double dc_0 = sqrt((cur[0] - cl[0]) * (cur[0] - cl[0]) + (cur[1] - cl[1]) * (cur[1] - cl[1]));
                double des_scaling = (dc_0 - (r + 1e-7)) / dc_0;
                if (!(des_scaling > 0.0)) des_scaling = 0.0; /* max(0.0, x) */
                final_position[0] = cur[0] + d_vec[0] * des_scaling; final_position[1] = cur[1] + d_vec[1] * des_scaling;
            } else {
                final_position[0] = cur[0]; final_position[1] = cur[1];
            }
        }
        /* make the new action; keep the slowest (:969-987) */
        if (kin == ORC_KIN_HOLONOMIC) {
            const double v_x = (final_position[0] - cur[0]) / dt, v_y = (final_position[1] - cur[1]) / dt;
            if ((v_x * v_x + v_y * v_y) < (fin[0] * fin[0] + fin[1] * fin[1])) { fin[0] = v_x; fin[1] = v_y; }
        } else {
            if (action2[0] > 0) {
                const double v = npnorm2(final_position[0] - cur[0], final_position[1] - cur[1]) / dt;
                if (v < fin[0]) { fin[0] = v; fin[1] = action2[1]; }
            } else {
                const double v = -npnorm2(final_position[0] - cur[0], final_position[1] - cur[1]) / dt;
                if (v > fin[0]) { fin[0] = v; fin[1] = action2[1]; }
            }
        }
    }
    out_action2[0] = fin[0]; out_action2[1] = fin[1];
}

/* ------------------------------------------------------------------------- */
/* crowd_sim_plus/envs/policy/social_force.py:38-94 */
void orc_sfm_predict(const OrcPolicyCfg *cfg, const double *self8, int n_others, const double *others, int n_seg,
                     const double *segs, double *out_v2)
{
    const double px = self8[0], py = self8[1], vx = self8[2], vy = self8[3], radius = self8[4];
    const double gx = self8[5], gy = self8[6], v_pref = self8[7];
    double delta_x = gx - px, delta_y = gy - py;
    double dist_to_goal = sqrt(delta_x * delta_x + delta_y * delta_y);
    dist_to_goal = dist_to_goal < 1e-6 ? 1.0 : dist_to_goal;
    const double desired_vx = (delta_x / dist_to_goal) * v_pref;
    const double desired_vy = (delta_y / dist_to_goal) * v_pref;
    const double curr_delta_vx = cfg->KI * (desired_vx - vx);
    const double curr_delta_vy = cfg->KI * (desired_vy - vy);
    double interaction_vx = 0, interaction_vy = 0;
    for (int j = 0; j < n_others; ++j) {
        const double *o = others + 5 * j;
        const double adjustment = fabs(cfg->sfm_radius - o[4]) + 0.01;
        delta_x = px - o[0]; delta_y = py - o[1];
        const double d = sqrt(delta_x * delta_x + delta_y * delta_y);
        const double e = cfg->A * exp((radius + o[4] + adjustment - d) / cfg->B);
        interaction_vx += e * (delta_x / d);
        interaction_vy += e * (delta_y / d);
    }
    for (int k = 0; k < n_seg; ++k) {
        const double *L = segs + 4 * k;
        double A_static, B_static;
        if (cfg->is_bottleneck && k >= 2) { A_static = cfg->A_bottleneck; B_static = cfg->B_bottleneck; }
        else { A_static = cfg->A_static; B_static = cfg->B_static; }
        double o2[2];
        orc_closest_point_on_segment(L[0], L[1], L[2], L[3], px, py, o2);
        delta_x = px - o2[0]; delta_y = py - o2[1];
        const double d = sqrt(delta_x * delta_x + delta_y * delta_y);
        const double e = A_static * exp((radius + 0.01 - d) / B_static);
        interaction_vx += e * (delta_x / d);
        interaction_vy += e * (delta_y / d);
    }
    const double total_delta_vx = (curr_delta_vx + interaction_vx) * cfg->time_step;
    const double total_delta_vy = (curr_delta_vy + interaction_vy) * cfg->time_step;
    const double new_vx = vx + total_delta_vx, new_vy = vy + total_delta_vy;
    const double act_norm = npnorm2(new_vx, new_vy); /* np.linalg.norm([new_vx, new_vy]) */
    if (act_norm > v_pref) { out_v2[0] = new_vx / act_norm * v_pref; out_v2[1] = new_vy / act_norm * v_pref; }
    else { out_v2[0] = new_vx; out_v2[1] = new_vy; }
}

/* ------------------------------------------------------------------------- */
/* Human.get_g_xy, crowd_sim_plus/envs/utils/human_plus.py:19-52 */
static void get_g_xy(const OrcDoorCfg *door, double px, double py, double fgx, double fgy, double *gx, double *gy)
{
    if (door && door->enabled) {
        const double ymin = py < fgy ? py : fgy, ymax = py > fgy ? py : fgy;
        if (ymin < door->door_y_mid_min && ymax > door->door_y_mid_max) {
            const double int_gx = door->door_x_mid;
            const double int_gy = 0.5 * (door->door_y_min + door->door_y_max);
            const double vec_norm = npnorm2(int_gx - px, int_gy - py);
            if (vec_norm <= door->door_width / 2.0) { *gx = fgx; *gy = fgy; }
            else { *gx = int_gx; *gy = int_gy; }
            return;
        }
    }
    *gx = fgx; *gy = fgy;
}

static double py_mod(double x, double y)
{
    double m = fmod(x, y);
    if (m != 0.0 && ((m < 0) != (y < 0))) m += y;
    return m;
}

/* Human.act (human_plus.py:103-116) for human i of env b: ob = other humans in index order, then the robot
 * (crowd_sim_plus.py:1047-1049) */
static void human_policy(const OrcPolicyCfg *pcfg, const OrcEnvState *st, int b, int i, double *others, double *out_v2,
                         int *nbr, int *nbr_cnt)
{
    const int H = st->H;
    const int o = b * H;
    double self8[8] = { st->px[o + i], st->py[o + i], st->vx[o + i], st->vy[o + i], st->radius[o + i],
                        st->gx[o + i], st->gy[o + i], st->vpref[o + i] };
    int n = 0;
    for (int j = 0; j < H; ++j) {
        if (j == i) continue;
        double *q = others + 5 * n++;
        q[0] = st->px[o + j]; q[1] = st->py[o + j]; q[2] = st->vx[o + j]; q[3] = st->vy[o + j]; q[4] = st->radius[o + j];
    }
    if (st->robot_visible) {
        double *q = others + 5 * n++;
        q[0] = st->rpx[b]; q[1] = st->rpy[b]; q[2] = st->rvx[b]; q[3] = st->rvy[b]; q[4] = st->rradius;
    }
    if (pcfg->policy == ORC_POLICY_SFM) {
        orc_sfm_predict(pcfg, self8, n, others, st->n_seg, st->segs, out_v2);
        if (nbr_cnt) *nbr_cnt = 0;
    } else {
        int ids[64], cnt = 0;
        int *idp = (pcfg->max_neighbors <= 64) ? ids : (int *)malloc(sizeof(int) * (size_t)pcfg->max_neighbors);
        orc_orca_predict(pcfg, self8, n, others, st->n_seg, st->segs, out_v2, idp, &cnt, NULL, NULL);
        if (nbr_cnt) {
            *nbr_cnt = cnt;
            for (int k = 0; k < cnt; ++k) { /* ob index -> agent id in the env (robot = H) */
                const int ob = idp[k];
                nbr[k] = (ob < H - 1) ? (ob < i ? ob : ob + 1) : H;
            }
        }
        if (idp != ids) free(idp);
    }
}

/* tiny pthread parallel-for over environments (no OpenMP runtime in the image) */
typedef void (*range_fn)(void *arg, int b0, int b1);
typedef struct { range_fn fn; void *arg; int b0, b1; } RangeJob;
static void *range_tramp(void *p) { RangeJob *j = (RangeJob *)p; j->fn(j->arg, j->b0, j->b1); return NULL; }
static void parallel_for(range_fn fn, void *arg, int B, int n_threads)
{
    if (n_threads <= 1 || B < 2) { fn(arg, 0, B); return; }
    if (n_threads > B) n_threads = B;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
    RangeJob *jobs = (RangeJob *)malloc(sizeof(RangeJob) * (size_t)n_threads);
    for (int t = 0; t < n_threads; ++t) {
        jobs[t].fn = fn; jobs[t].arg = arg;
        jobs[t].b0 = (int)((long long)B * t / n_threads); jobs[t].b1 = (int)((long long)B * (t + 1) / n_threads);
        pthread_create(&th[t], NULL, range_tramp, &jobs[t]);
    }
    for (int t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
    free(th); free(jobs);
}

typedef struct { const OrcPolicyCfg *pcfg; const OrcEnvState *st; double *out_v; int *nbr, *nbr_cnt; } PolicyArgs;
static void policy_range(void *p, int b0, int b1)
{
    PolicyArgs *a = (PolicyArgs *)p;
    const int H = a->st->H, MN = a->pcfg->max_neighbors;
    double *others = (double *)malloc(sizeof(double) * 5 * (size_t)(H + 1));
    for (int b = b0; b < b1; ++b)
        for (int i = 0; i < H; ++i)
            human_policy(a->pcfg, a->st, b, i, others, a->out_v + 2 * (b * H + i),
                         a->nbr ? a->nbr + (size_t)(b * H + i) * MN : NULL, a->nbr_cnt ? a->nbr_cnt + b * H + i : NULL);
    free(others);
}

void orc_policy_batch(const OrcPolicyCfg *pcfg, const OrcEnvState *st, double *out_v, int *nbr, int *nbr_cnt,
                      int n_threads)
{
    PolicyArgs a = { pcfg, st, out_v, nbr, nbr_cnt };
    parallel_for(policy_range, &a, st->B, n_threads);
}

/* CrowdSimPlus.step, crowd_sim_plus/envs/crowd_sim_plus.py:1025-1257 (update=True, non-SB3 observation) */
typedef struct {
    const OrcPolicyCfg *pcfg; const OrcDoorCfg *door; const OrcRewardCfg *rcfg; OrcEnvState *st;
    const double *robot_action; const unsigned char *active; double *reward, *dmin_out; int *flags, *nbr, *nbr_cnt;
} StepArgs;

static void step_range(void *p, int b0, int b1)
{
    StepArgs *A_ = (StepArgs *)p;
    const OrcPolicyCfg *pcfg = A_->pcfg; const OrcDoorCfg *door = A_->door; const OrcRewardCfg *rcfg = A_->rcfg;
    OrcEnvState *st = A_->st; const double *robot_action = A_->robot_action; const unsigned char *active = A_->active;
    double *reward = A_->reward, *dmin_out = A_->dmin_out; int *flags = A_->flags, *nbr = A_->nbr, *nbr_cnt = A_->nbr_cnt;
    const int H = st->H, MN = pcfg->max_neighbors;
    const double dt = pcfg->time_step;
    {
        double *others = (double *)malloc(sizeof(double) * 5 * (size_t)(H + 1));
        double *hact = (double *)malloc(sizeof(double) * 2 * (size_t)(H + 1));
        for (int b = b0; b < b1; ++b) {
            if (active && !active[b]) continue;
            const int o = b * H;
            /* human actions from the same pre-step state (:1044-1055) */
            for (int i = 0; i < H; ++i) {
                double a[2];
                human_policy(pcfg, st, b, i, others, a, nbr ? nbr + (size_t)(o + i) * MN : NULL,
                             nbr_cnt ? nbr_cnt + o + i : NULL);
                const double pose[3] = { st->px[o + i], st->py[o + i], st->theta[o + i] };
                orc_constrain_action(pose, st->radius[o + i], dt, ORC_KIN_HOLONOMIC, a, st->n_seg, st->segs, hact + 2 * i);
            }
            /* robot clamp + wall-collision flag (:1058-1064); quirk q11: compares only the first component */
            const double rpose[3] = { st->rpx[b], st->rpy[b], st->rtheta[b] };
            double ract[2];
            orc_constrain_action(rpose, st->rradius, dt, st->robot_kinematics, robot_action + 2 * b, st->n_seg, st->segs, ract);
            const int stat_collision = (robot_action[2 * b] != ract[0]);
            /* robot-human collision on NEXT positions, break at first hit (:1067-1080) */
            double dmin = INFINITY;
            int collision = 0;
            double rnext[2];
            compute_position(rpose, st->robot_kinematics, ract, dt, rnext);
            for (int i = 0; i < H; ++i) {
                const double x1 = st->px[o + i] + hact[2 * i] * dt, y1 = st->py[o + i] + hact[2 * i + 1] * dt;
                const double closest = npnorm2(rnext[0] - x1, rnext[1] - y1);
                if (closest < (st->rradius + st->radius[o + i])) { collision = 1; break; }
                else if (closest < dmin) dmin = closest;
            }
            /* frozen (:1083-1087) */
            int frozen;
            if (st->robot_kinematics == ORC_KIN_HOLONOMIC) frozen = sqrt(ract[0] * ract[0] + ract[1] * ract[1]) * dt < 0.01;
            else frozen = fabs(ract[0] * dt) < 0.01;
            /* goal / progress (:1090-1094) */
            const int reached = npnorm2(rnext[0] - st->rgx[b], rnext[1] - st->rgy[b]) < st->rradius;
            const double curr_dist = npnorm2(st->rgx[b] - rnext[0], st->rgy[b] - rnext[1]);
            double rew = 0.0;
            int f = 0;
            if (reached) { rew += rcfg->success_reward; f |= ORC_F_REACHED | ORC_F_DONE; }
            else if (st->global_time[b] >= rcfg->time_limit) { rew += rcfg->timeout; f |= ORC_F_TIMEOUT | ORC_F_DONE; }
            if (collision) { rew += rcfg->collision_penalty; f |= ORC_F_COLLISION; }
            if (stat_collision) { rew += rcfg->wall_collision_penalty; f |= ORC_F_WALL; }
            if (rcfg->discomfort && dmin < rcfg->discomfort_dist) {
                rew += (dmin - rcfg->discomfort_dist) * rcfg->discomfort_penalty_factor * dt;
                f |= ORC_F_DANGER;
            }
            if (rcfg->has_progress) {
                rew += (st->prev_dist[b] - curr_dist) * rcfg->progress_factor;
                st->prev_dist[b] = curr_dist;
            }
            if (frozen) { rew += rcfg->freezing_penalty; f |= ORC_F_FROZEN; }
            if (reward) reward[b] = rew;
            if (dmin_out) dmin_out[b] = dmin;
            if (flags) flags[b] = f;
            /* update (:1193-1206): Agent.step agent_plus.py:199-214, Human.step human_plus.py:118-120 */
            if (st->robot_kinematics == ORC_KIN_HOLONOMIC) {
                st->rpx[b] = rnext[0]; st->rpy[b] = rnext[1];
                st->rvx[b] = ract[0]; st->rvy[b] = ract[1];
                st->rtheta[b] = atan2(ract[1], ract[0]);
            } else {
                st->rpx[b] = rnext[0]; st->rpy[b] = rnext[1];
                const double un = py_mod(st->rtheta[b] + ract[1], 2 * M_PI);
                st->rtheta[b] = un > M_PI ? un - 2 * M_PI : un;
                st->rvx[b] = ract[0] * cos(st->rtheta[b]);
                st->rvy[b] = ract[0] * sin(st->rtheta[b]);
            }
            for (int i = 0; i < H; ++i) {
                st->px[o + i] = st->px[o + i] + hact[2 * i] * dt;
                st->py[o + i] = st->py[o + i] + hact[2 * i + 1] * dt;
                st->vx[o + i] = hact[2 * i]; st->vy[o + i] = hact[2 * i + 1];
                st->theta[o + i] = atan2(hact[2 * i + 1], hact[2 * i]);
                get_g_xy(door, st->px[o + i], st->py[o + i], st->fgx[o + i], st->fgy[o + i], &st->gx[o + i], &st->gy[o + i]);
            }
            st->global_time[b] += dt;
            for (int i = 0; i < H; ++i) {
                if (st->human_time[o + i] == 0 &&
                    npnorm2(st->px[o + i] - st->gx[o + i], st->py[o + i] - st->gy[o + i]) < st->radius[o + i])
                    st->human_time[o + i] = st->global_time[b];
            }
        }
        free(others); free(hact);
    }
}

void orc_env_step(const OrcPolicyCfg *pcfg, const OrcDoorCfg *door, const OrcRewardCfg *rcfg, OrcEnvState *st,
                  const double *robot_action, const unsigned char *active, double *reward, double *dmin,
                  int *flags, int *nbr, int *nbr_cnt, int n_threads)
{
    StepArgs a = { pcfg, door, rcfg, st, robot_action, active, reward, dmin, flags, nbr, nbr_cnt };
    parallel_for(step_range, &a, st->B, n_threads);
}


/* n_cases independent ORCA.predict / ORCAPlus.predict calls (case c: self8[c], the first n_others[c] rows of others[c, E, 5]) with the
 * half-planes, preferred velocity and max speed agent 0 ended up with, for the property tests. */
void orc_orca_probe(const OrcPolicyCfg *cfg, int n_cases, const double *self8, int E, const double *others, const int *n_others,
                    int n_seg, const double *segs, double *out_v, OrcProbe *probes)
{
    for (int c = 0; c < n_cases; ++c) {
        g_probe = probes + c;
        int ids[64], cnt = 0;
        orc_orca_predict(cfg, self8 + 8 * (size_t)c, n_others[c], others + 5 * (size_t)E * c, n_seg, segs, out_v + 2 * (size_t)c,
                         cfg->max_neighbors <= 64 ? ids : NULL, cfg->max_neighbors <= 64 ? &cnt : NULL, NULL, NULL);
        probes[c].n_agent_nbr = cnt;
        for (int k = 0; k < cnt && k < 16; ++k) probes[c].nbr_ids[k] = ids[k];
    }
    g_probe = NULL;
}
