"""Independent fp64 checkers for ORCA (tests/test_orca_properties.py).  TEST INFRASTRUCTURE.

RVO2 is not in /root/reference and cannot be installed: nothing can pin oracle/rvo2_oracle.c to the real library.  What
CAN be checked is that its outputs -- and the CUDA kernel's -- have the properties that DEFINE ORCA, derived here from the geometry
and not from RVO2's formulas:

  (i)   the new velocity is the point of  disc(max_speed) ∩ half-planes  closest to the preferred velocity (found by brute-force vertex
        enumeration in fp64), or -- when that set is empty -- it minimises the maximum violation of the agent half-planes subject to
        the obstacle half-planes and the disc (linearProgram3's contract; bisection + the same enumeration);
  (ii)  every agent half-plane passes through v_A + u/2 where u is the smallest change of the relative velocity that leaves the
        truncated velocity obstacle VO^tau_{A|B}, and is tangent to VO there; the VO boundary (cut-off arc + two legs) is built from
        angles (atan2 / asin), not from RVO2's dot-product branch tests;
  (ii') the same half-planes restated numerically from the reference's own CasADi transcription
        (sicnav/utils/mpc_utils/orca_casadi.py:200-268, the non-colliding branches; its colliding branch :269-287 is a different
        formula -- protrusion^2/dt -- and is NOT RVO2's, so it is not compared);
  (iii) every obstacle half-plane is a supporting line of the obstacle's velocity obstacle: moving with any permitted velocity for
        time_horizon_obst seconds keeps the agent at least `radius` away from that wall segment, and the line's own point touches
        the VO boundary (swept distance == radius).
"""
import numpy as np

# branch ids of oracle/rvo2_oracle.c's counters
BRANCHES = {
    "obst_already_covered": 0, "obst_collision_left_vertex": 1, "obst_collision_right_vertex": 2, "obst_collision_segment": 3,
    "obst_oblique_left": 4, "obst_oblique_right": 5, "obst_usual": 6, "obst_left_leg_foreign": 7, "obst_right_leg_foreign": 8,
    "obst_project_left_cutoff_circle": 9, "obst_project_right_cutoff_circle": 10, "obst_project_cutoff_line": 11,
    "obst_project_left_leg": 12, "obst_project_right_leg": 13, "obst_skip_foreign_left": 14, "obst_skip_foreign_right": 15,
    "agent_cutoff_circle": 20, "agent_left_leg": 21, "agent_right_leg": 22, "agent_collision": 23,
    "lp1_outside_disc": 30, "lp1_parallel_infeasible": 31, "lp1_interval_empty": 32, "lp2_pref_clamped_to_disc": 33,
    "lp3_invoked": 34, "lp3_parallel_same_direction": 35, "lp3_parallel_opposite": 36, "lp3_inner_lp_failed": 37,
    "lp1_direction_opt": 38, "lp1_clip_left": 39, "lp1_clip_right": 40, "lp1_parallel_skip": 41, "lp1_interior": 42,
}


# ------------------------------------------------------------------------------------------------ random populations
def random_cases(seed, n, E=10, crowded=0.35, colliding=0.03, rng=None):
    """n agents with up to E others.  `crowded`: share of cases drawn in a tight box (many binding constraints, LP3 fallbacks);
    `colliding`: share where one other already overlaps the agent (RVO2's collision branch)."""
    rng = rng or np.random.default_rng(seed)
    box = np.where(rng.random(n) < crowded, rng.uniform(0.8, 1.6, n), rng.uniform(2.0, 5.0, n))[:, None]
    pos = rng.uniform(-1, 1, (n, 2)) * box
    vel = rng.uniform(-1, 1, (n, 2))
    goal = rng.uniform(-4, 4, (n, 2))
    near = rng.random(n) < 0.1                        # short goal vector: preferred velocity not normalised (orca.py:113-115)
    goal[near] = pos[near] + rng.uniform(-0.6, 0.6, (near.sum(), 2))
    vpref = rng.uniform(0.5, 1.5, (n, 1))
    rad = rng.uniform(0.2, 0.35, (n, 1))
    self8 = np.concatenate([pos, vel, rad, goal, vpref], 1)
    opos = rng.uniform(-1, 1, (n, E, 2)) * box[:, None]
    ovel = rng.uniform(-1, 1, (n, E, 2))
    orad = rng.uniform(0.2, 0.35, (n, E, 1))
    # push overlapping others out to just beyond contact, except the deliberately colliding share
    d = opos - pos[:, None]
    dist = np.linalg.norm(d, axis=2, keepdims=True)
    need = rad[:, None] + orad + 0.02 + 0.05
    coll = rng.random((n, E, 1)) < colliding
    scale = np.where((dist < need) & ~coll, need / np.maximum(dist, 1e-9), 1.0)
    opos = pos[:, None] + d * scale
    n_others = rng.integers(0, E + 1, n).astype(np.int32)
    others = np.concatenate([opos, ovel, orad], 2)
    return self8, others, n_others


# ------------------------------------------------------------------------------------------------ (ii) agent half-planes
def _rot(v, ang):
    c, s = np.cos(ang), np.sin(ang)
    return np.stack([c * v[..., 0] - s * v[..., 1], s * v[..., 0] + c * v[..., 1]], -1)


def agent_lines_from_geometry(pA, vA, rA, pB, vB, rB, tau, dt):
    """Arrays over pairs.  Returns (point, direction, u, margin): the ORCA half-plane of A induced by B and `margin`, the gap between
    the best and second best boundary piece (small margin = the closest boundary point is not unique; such pairs are ambiguous)."""
    p = pB - pA
    v = vA - vB
    R = rA + rB
    d = np.linalg.norm(p, axis=-1)
    coll = d <= R
    T = np.where(coll, dt, tau)                                   # colliding pairs: the VO of ONE time step, cut-off disc only
    c = p / T[..., None]
    r = R / T
    w = v - c
    wl = np.linalg.norm(w, axis=-1)
    what = w / np.maximum(wl, 1e-300)[..., None]
    # --- cut-off arc: circle(c, r), the part facing the origin: directions within (pi/2 - alpha) of -p
    phi = np.arctan2(p[..., 1], p[..., 0])
    alpha = np.arcsin(np.clip(R / np.maximum(d, 1e-300), 0.0, 1.0))
    ang_w = np.arctan2(w[..., 1], w[..., 0])
    off = np.abs((ang_w - (phi + np.pi) + np.pi) % (2 * np.pi) - np.pi)          # angle between w and -p
    span = np.pi / 2 - alpha
    q_arc = c + r[..., None] * what
    n_arc = what
    dist_arc = np.where(coll | (off <= span), np.abs(wl - r), np.inf)
    # --- legs: rays from the tangent points, along angles phi +- alpha, starting at distance s0 from the origin
    s0 = np.sqrt(np.maximum(d * d - R * R, 0.0)) / T
    out = []
    for sign in (+1.0, -1.0):
        ell = np.stack([np.cos(phi + sign * alpha), np.sin(phi + sign * alpha)], -1)
        s = np.sum(v * ell, -1)
        foot = s[..., None] * ell
        dist = np.where(~coll & (s >= s0), np.linalg.norm(v - foot, axis=-1), np.inf)
        nrm = _rot(ell, sign * np.pi / 2)                         # outward normal of the cone on that leg
        out.append((dist, foot, nrm))
    dists = np.stack([dist_arc, out[0][0], out[1][0]], -1)
    pick = np.argmin(dists, -1)
    srt = np.sort(dists, -1)
    margin = srt[..., 1] - srt[..., 0]
    q = np.where((pick == 0)[..., None], q_arc, np.where((pick == 1)[..., None], out[0][1], out[1][1]))
    n = np.where((pick == 0)[..., None], n_arc, np.where((pick == 1)[..., None], out[0][2], out[1][2]))
    u = q - v
    point = vA + 0.5 * u
    direction = np.stack([n[..., 1], -n[..., 0]], -1)
    return point, direction, u, margin, pick, coll


def agent_lines_from_casadi_restatement(pA, vA, rA, pB, vB, rB, tau):
    """sicnav/utils/mpc_utils/orca_casadi.py:204-268 + :289-292 evaluated with numpy fp64 (non-colliding pairs only)."""
    rel_pos = pB - pA
    rel_vel = vA - vB
    dist_sq = np.sum(rel_pos * rel_pos, -1)
    comb_rad = rA + rB
    comb_rad_sq = comb_rad ** 2
    inv_t = 1.0 / tau
    w = rel_vel - inv_t * rel_pos
    w_len_sq = np.sum(w * w, -1)
    dot1 = np.sum(w * rel_pos, -1)
    cutoff = (dot1 < 0.0) & (dot1 ** 2 > comb_rad_sq * w_len_sq)
    w_len = np.sqrt(w_len_sq)
    unit_w = w / np.maximum(w_len, 1e-300)[..., None]
    dir_c = np.stack([unit_w[..., 1], -unit_w[..., 0]], -1)
    u_c = (comb_rad * inv_t - w_len)[..., None] * unit_w
    leg = np.sqrt(np.abs(dist_sq - comb_rad_sq))
    left = np.stack([rel_pos[..., 0] * leg - rel_pos[..., 1] * comb_rad, rel_pos[..., 0] * comb_rad + rel_pos[..., 1] * leg], -1) / dist_sq[..., None]
    right = -np.stack([rel_pos[..., 0] * leg + rel_pos[..., 1] * comb_rad, -rel_pos[..., 0] * comb_rad + rel_pos[..., 1] * leg], -1) / dist_sq[..., None]
    is_left = (rel_pos[..., 0] * w[..., 1] - rel_pos[..., 1] * w[..., 0]) > 0.0
    dir_l = np.where(is_left[..., None], left, right)
    u_l = np.sum(rel_vel * dir_l, -1)[..., None] * dir_l - rel_vel
    direction = np.where(cutoff[..., None], dir_c, dir_l)
    u = np.where(cutoff[..., None], u_c, u_l)
    return vA + 0.5 * u, direction, dist_sq > comb_rad_sq


# ------------------------------------------------------------------------------------------------ (i) the linear programs
def _candidates(P, D, r):
    """P, D [n,L,2] -> candidate vertices [n, K, 2] (pair intersections, line-circle intersections); invalid ones are NaN."""
    n, L, _ = P.shape
    cands = []
    iu, ju = np.triu_indices(L, 1)
    if len(iu):
        Pi, Di, Pj, Dj = P[:, iu], D[:, iu], P[:, ju], D[:, ju]
        den = Di[..., 0] * Dj[..., 1] - Di[..., 1] * Dj[..., 0]
        num = Dj[..., 0] * (Pi[..., 1] - Pj[..., 1]) - Dj[..., 1] * (Pi[..., 0] - Pj[..., 0])   # det(Dj, Pi - Pj)
        with np.errstate(divide="ignore", invalid="ignore"):
            t = np.where(np.abs(den) > 1e-12, num / den, np.nan)
        cands.append(Pi + t[..., None] * Di)
    b = np.sum(P * D, -1)
    disc = b * b + (r * r)[:, None] - np.sum(P * P, -1)
    with np.errstate(invalid="ignore"):
        sq = np.sqrt(np.where(disc >= 0, disc, np.nan))
    cands.append(P + (-b - sq)[..., None] * D)
    cands.append(P + (-b + sq)[..., None] * D)
    return np.concatenate(cands, 1)


def _feasible(X, P, D, valid, r, shift, tol):
    """X [n,K,2]; constraint k: det(D_k, P_k - x) <= shift_k + tol; |x| <= r + tol."""
    viol = D[:, None, :, 0] * (P[:, None, :, 1] - X[:, :, None, 1]) - D[:, None, :, 1] * (P[:, None, :, 0] - X[:, :, None, 0])
    ok = np.all((viol <= shift[:, None, :] + tol) | ~valid[:, None, :], -1)
    with np.errstate(invalid="ignore"):
        ok &= np.sum(X * X, -1) <= (r[:, None] + tol) ** 2
    return ok & np.all(np.isfinite(X), -1)


def project_onto_feasible_set(P, D, valid, r, q, tol=1e-9):
    """Closest point to q inside disc(r) ∩ {half-planes}: brute force over every vertex + the unconstrained / single-line optima.
    Returns (x [n,2], feasible [n])."""
    n, L, _ = P.shape
    Pm = np.where(valid[..., None], P, 0.0)
    Dm = np.where(valid[..., None], D, np.array([1.0, 0.0]))
    qn = np.linalg.norm(q, axis=-1)
    q_clip = q * np.minimum(1.0, r / np.maximum(qn, 1e-300))[:, None]
    t = np.sum((q[:, None] - Pm) * Dm, -1)
    proj = Pm + t[..., None] * Dm
    X = np.concatenate([q_clip[:, None], proj, _candidates(Pm, Dm, r)], 1)
    ok = _feasible(X, Pm, Dm, valid, r, np.zeros((n, L)), tol)
    dist = np.where(ok, np.linalg.norm(X - q[:, None], axis=-1), np.inf)
    best = np.argmin(dist, 1)
    return X[np.arange(n), best], np.isfinite(dist[np.arange(n), best])


def min_max_violation(P, D, valid, is_obst, r, iters=40):
    """min over x in disc ∩ obstacle half-planes of max over agent lines of det(D_i, P_i - x): bisection on the level d with the
    vertex enumeration as the emptiness test.  Returns d* [n] (inf where even the obstacle lines + disc are infeasible)."""
    n, L, _ = P.shape
    Pm = np.where(valid[..., None], P, 0.0)
    Dm = np.where(valid[..., None], D, np.array([1.0, 0.0]))
    agent = valid & ~is_obst
    Nl = np.stack([-Dm[..., 1], Dm[..., 0]], -1)            # left normal of each line (the permitted side)

    def nonempty(dv):
        # relaxing line i by d moves it by d towards the forbidden side: P_i - d * N_i
        Ps = Pm - (dv[:, None] * agent)[..., None] * Nl
        X = np.concatenate([np.zeros((n, 1, 2)), _candidates(Ps, Dm, r)], 1)
        return np.any(_feasible(X, Ps, Dm, valid, r, np.zeros((n, L)), 1e-9), 1)

    lo = np.zeros(n)
    hi = np.full(n, 4.0 * (np.max(np.linalg.norm(Pm, axis=-1), 1) + r + 1.0))
    base_ok = nonempty(hi)
    for _ in range(iters):
        mid = 0.5 * (lo + hi)
        ok = nonempty(mid)
        hi = np.where(ok, mid, hi)
        lo = np.where(ok, lo, mid)
    return np.where(base_ok, hi, np.inf)


def lp3_conditioning(P, D, agent):
    """largest distance from the origin of a pairwise intersection point of the agent lines (RVO2 treats |det| <= 1e-5 as parallel)"""
    n, L, _ = P.shape
    iu, ju = np.triu_indices(L, 1)
    if not len(iu):
        return np.zeros(n)
    Pi, Di, Pj, Dj = P[:, iu], D[:, iu], P[:, ju], D[:, ju]
    den = Di[..., 0] * Dj[..., 1] - Di[..., 1] * Dj[..., 0]
    num = Dj[..., 0] * (Pi[..., 1] - Pj[..., 1]) - Dj[..., 1] * (Pi[..., 0] - Pj[..., 0])
    both = agent[:, iu] & agent[:, ju] & (np.abs(den) > 1e-5)
    with np.errstate(divide="ignore", invalid="ignore"):
        X = Pi + np.where(both, num / den, 0.0)[..., None] * Di
    return np.where(both, np.linalg.norm(X, axis=-1), 0.0).max(1)


def violations(x, P, D, valid):
    """signed violation det(D_k, P_k - x) of every line (positive = violated); -inf where the slot is unused."""
    v = D[..., 0] * (P[..., 1] - x[:, None, 1]) - D[..., 1] * (P[..., 0] - x[:, None, 0])
    return np.where(valid, v, -np.inf)


def check_velocity(v, P, D, n_lines, n_obst, r, q, feas_tol=1e-5, opt_tol=1e-4, disc_tol=5e-5, chunk=4000):
    """Property (i) for a population.  v [n,2] new velocities, P/D [n,Lmax,2] half-planes (obstacle lines first), r max speeds, q
    preferred velocities.  Returns a dict of counts and worst-case errors; raises AssertionError listing offenders."""
    n = v.shape[0]
    stats = dict(n=n, feasible=0, infeasible=0, ill_conditioned=0, worst_feas=0.0, worst_disc=0.0, worst_opt=0.0, worst_pos=0.0, worst_lp3=0.0, lp3_kappa_gt_10=0, worst_disc_ill=0.0)
    order = np.argsort(n_lines, kind="stable")
    for s0 in range(0, n, chunk):
        idx = order[s0:s0 + chunk]
        L = max(int(n_lines[idx].max()), 1)
        Pc, Dc = P[idx, :L].astype(np.float64), D[idx, :L].astype(np.float64)
        valid = np.arange(L)[None, :] < n_lines[idx, None]
        Dc = np.where(valid[..., None], Dc / np.maximum(np.linalg.norm(Dc, axis=-1, keepdims=True), 1e-300), Dc)  # float32 unit vectors -> exact
        is_obst = np.arange(L)[None, :] < n_obst[idx, None]
        rc, qc, vc = r[idx], q[idx], v[idx]
        xb, feas = project_onto_feasible_set(Pc, Dc, valid, rc, qc)
        viol = violations(vc, Pc, Dc, valid)
        speed_excess = np.linalg.norm(vc, axis=-1) - rc
        # ---- feasible cases: v satisfies everything and is the projection of q
        f = np.where(feas)[0]
        if len(f):
            worst = viol[f].max(1)
            stats["worst_feas"] = max(stats["worst_feas"], float(worst.max()))
            stats["worst_disc"] = max(stats["worst_disc"], float(speed_excess[f].max()))
            assert np.all(worst <= feas_tol), f"half-plane violated by {worst.max():.3e} (case {idx[f[np.argmax(worst)]]})"
            assert np.all(speed_excess[f] <= disc_tol), f"max-speed disc violated by {speed_excess[f].max():.3e}"
            gap = np.linalg.norm(vc[f] - qc[f], axis=-1) - np.linalg.norm(xb[f] - qc[f], axis=-1)
            stats["worst_opt"] = max(stats["worst_opt"], float(gap.max()))
            assert np.all(gap <= opt_tol), f"not the closest point: {gap.max():.3e} farther than the brute-force optimum"
            pos = np.linalg.norm(vc[f] - xb[f], axis=-1)
            # conditioning: the optimum of the same problem with every constraint moved by +-2e-5 (fp32 noise of the lines)
            bad = np.where(pos > opt_tol)[0]
            if len(bad):
                b = f[bad]
                Pl = np.stack([-Dc[b][..., 1], Dc[b][..., 0]], -1)
                x_in, ok_in = project_onto_feasible_set(Pc[b] + 2e-5 * Pl, Dc[b], valid[b], rc[b] - 2e-5, qc[b])
                x_out, _ = project_onto_feasible_set(Pc[b] - 2e-5 * Pl, Dc[b], valid[b], rc[b] + 2e-5, qc[b])
                sens = np.where(ok_in, np.linalg.norm(x_in - x_out, axis=-1), np.inf)
                ill = sens > 0.5 * pos[bad]                       # the optimum itself moves that much under fp32-sized perturbations
                stats["ill_conditioned"] += int(ill.sum())
                assert np.all(ill), f"velocity differs from the fp64 projection by {pos[bad][~ill].max():.3e} in a well-conditioned case"
                pos = np.delete(pos, bad)
            if len(pos):
                stats["worst_pos"] = max(stats["worst_pos"], float(pos.max()))
            stats["feasible"] += len(f)
        # ---- infeasible cases: linearProgram3's contract
        g = np.where(~feas)[0]
        if len(g):
            dstar = min_max_violation(Pc[g], Dc[g], valid[g], is_obst[g], rc[g])
            agent_v = np.where(is_obst[g], -np.inf, viol[g]).max(1)
            obst_v = np.where(is_obst[g], viol[g], -np.inf).max(1)
            ok_hard = np.isfinite(dstar)                           # (obstacle lines + disc alone infeasible: RVO2 keeps the LP2 result)
            # float32 conditioning of linearProgram3: the line it projects for a pair (i, j) passes through the pair's intersection
            # point, kappa m/s from the origin; for half-planes within ~1e-2 rad of (anti-)parallel kappa is 10 .. 1000+ and
            # linearProgram1's discriminant  dot^2 + r^2 - |point|^2  cancels in float32 (absolute error ~6e-8 kappa^2), so the
            # result can leave the max-speed disc by that much (measured on 10^4 infeasible cases: <= 7.5e-6 for kappa <= 10,
            # 1.4e-4 at 30, 1.5e-2 at 300, 0.21 beyond 1000).  Inherited RVO2 arithmetic, kept for parity: the disc bound is
            # asserted with that error model, the min-max contract and the obstacle lines for EVERY case.
            kap = lp3_conditioning(Pc[g], Dc[g], valid[g] & ~is_obst[g])
            stats["lp3_kappa_gt_10"] += int((kap > 10.0).sum())
            stats["worst_disc"] = max(stats["worst_disc"], float(speed_excess[g][kap <= 10.0].max()) if (kap <= 10.0).any() else 0.0)
            stats["worst_disc_ill"] = max(stats["worst_disc_ill"], float(speed_excess[g].max()))
            assert np.all(obst_v[ok_hard] <= feas_tol + 4e-7 * kap[ok_hard] ** 2), f"LP3 result violates an obstacle line by {obst_v[ok_hard].max():.3e}"
            bound = np.maximum(disc_tol, 2e-5 + 4e-7 * kap ** 2)
            assert np.all(speed_excess[g] <= bound), f"LP3 result leaves the max-speed disc by {(speed_excess[g] - bound).max():.3e} more than float32 explains"
            excess = agent_v[ok_hard] - dstar[ok_hard]
            if len(excess):
                stats["worst_lp3"] = max(stats["worst_lp3"], float(excess.max()))
                assert np.all(excess <= opt_tol), f"LP3 result's max violation exceeds the fp64 minimum by {excess.max():.3e}"
            stats["infeasible"] += len(g)
    return stats


# ------------------------------------------------------------------------------------------------ (iii) obstacle half-planes
def _seg_seg_dist(a0, a1, b0, b1):
    """fp64 distance between segments a0-a1 and b0-b1 (arrays [...,2]): 0 if they cross, else the smallest end-point distance."""
    def pt_seg(p, s0, s1):
        d = s1 - s0
        L2 = np.sum(d * d, -1)
        t = np.clip(np.sum((p - s0) * d, -1) / np.maximum(L2, 1e-300), 0.0, 1.0)
        return np.linalg.norm(p - (s0 + t[..., None] * d), axis=-1)

    def cross(u, w):
        return u[..., 0] * w[..., 1] - u[..., 1] * w[..., 0]
    d1 = cross(b1 - b0, a0 - b0); d2 = cross(b1 - b0, a1 - b0); d3 = cross(a1 - a0, b0 - a0); d4 = cross(a1 - a0, b1 - a0)
    hit = (d1 * d2 < 0) & (d3 * d4 < 0)
    m = np.minimum(np.minimum(pt_seg(a0, b0, b1), pt_seg(a1, b0, b1)), np.minimum(pt_seg(b0, a0, a1), pt_seg(b1, a0, a1)))
    return np.where(hit, 0.0, m)


def swept_clearance(pA, w, T, seg0, seg1):
    """smallest distance between the wall segment and the agent centre moving from pA with velocity w for T seconds"""
    return _seg_seg_dist(pA, pA + w * T, seg0, seg1)
