// csrc/scene_kernels.cu -- batched, seeded scenario reset on the device.
//
// CrowdSimPlus.reset builds one scene per episode with numpy's default_rng(offset + case) and rejection sampling
// (crowd_sim_plus/envs/crowd_sim_plus.py:658-664; generate_random_human_position :425-451,
// generate_circle_crossing_human :454-481, generate_hallway_human :522-605, Human.get_g_xy human_plus.py:19-52).  For
// thousands of environments the Python generators dominate episode turnaround, so this kernel restates them with ONE
// THREAD PER ENVIRONMENT and writes the fp64 SoA crowd state in place:
//   * numpy's SeedSequence entropy hashing + PCG64 (XSL-RR 128/64) seeding and stream, so env b consumes exactly the
//     draws `default_rng(seed_b)` would hand the reference (`random()` = (u64 >> 11) * 2^-53; `uniform(a,b)` = a + (b-a) random()),
//   * the generators' draw order, collision tests and `eff_h *= 1.1` retries, statement by statement.
// Accept / reject decisions and every drawn number are identical to the host generator (snb/scenario.py, itself pinned to
// the reference's golden episodes); positions differ by at most an ulp where CUDA's cos / sin / atan2 round differently
// from the host libm.  Compiled with -fmad=false: Python float arithmetic never contracts, numpy's 2-element dot does
// (npnorm2 below).
#include <mutex>

#include "snb_common.h"

namespace {

struct SeedSeqPcg64 {
    unsigned __int128 state, inc;

    __device__ static uint32_t hashmix(uint32_t v, uint32_t &hc)
    {
        v ^= hc; hc *= 0x931e8875u; v *= hc; v ^= v >> 16;
        return v;
    }
    __device__ static uint32_t mix(uint32_t x, uint32_t y)
    {
        uint32_t r = 0xca01f9ddu * x - 0x4973f715u * y;
        r ^= r >> 16;
        return r;
    }
    __device__ void step() { state = state * (((unsigned __int128)0x2360ED051FC65DA4ull << 64) | 0x4385DF649FCCF645ull) + inc; }

    // np.random.default_rng(seed): SeedSequence(seed).generate_state(4, uint64) -> pcg64_srandom_r(initstate, initseq)
    __device__ void seed(uint64_t s)
    {
        uint32_t ent[2] = {(uint32_t)s, (uint32_t)(s >> 32)};
        const int n_ent = ent[1] ? 2 : 1;
        uint32_t hc = 0x43b0d7e5u, pool[4];
        for (int i = 0; i < 4; ++i) pool[i] = hashmix(i < n_ent ? ent[i] : 0u, hc);
        for (int i_src = 0; i_src < 4; ++i_src)
            for (int i_dst = 0; i_dst < 4; ++i_dst)
                if (i_src != i_dst) pool[i_dst] = mix(pool[i_dst], hashmix(pool[i_src], hc));
        uint32_t hb = 0x8b51f9ddu, w[8];
        for (int i = 0; i < 8; ++i) {
            uint32_t v = pool[i & 3];
            v ^= hb; hb *= 0x58f38dedu; v *= hb; v ^= v >> 16;
            w[i] = v;
        }
        const uint64_t u0 = w[0] | ((uint64_t)w[1] << 32), u1 = w[2] | ((uint64_t)w[3] << 32);
        const uint64_t u2 = w[4] | ((uint64_t)w[5] << 32), u3 = w[6] | ((uint64_t)w[7] << 32);
        const unsigned __int128 initstate = ((unsigned __int128)u0 << 64) | u1, initseq = ((unsigned __int128)u2 << 64) | u3;
        inc = (initseq << 1) | 1;
        state = 0;
        step();
        state += initstate;
        step();
    }
    __device__ uint64_t next64()
    {
        step();
        const uint64_t hi = (uint64_t)(state >> 64), lo = (uint64_t)state;
        const uint64_t x = hi ^ lo;
        const unsigned rot = (unsigned)(hi >> 58);
        return (x >> rot) | (x << ((64 - rot) & 63));
    }
    __device__ double random() { return (double)(next64() >> 11) * (1.0 / 9007199254740992.0); }
    __device__ double uniform(double a, double b) { return a + (b - a) * random(); }
};

__device__ __forceinline__ double npnorm2(double x, double y) { return sqrt(fma(y, y, x * x)); }   // np.linalg.norm((x, y))

// utils_plus.point_to_segment_dist (utils_plus.py:73-95)
__device__ double point_to_segment_dist(double x1, double y1, double x2, double y2, double x3, double y3)
{
    const double px = x2 - x1, py = y2 - y1;
    if (px == 0 && py == 0) return npnorm2(x3 - x1, y3 - y1);
    double u = ((x3 - x1) * px + (y3 - y1) * py) / (px * px + py * py);
    u = u > 1 ? 1 : (u < 0 ? 0 : u);
    return npnorm2(x1 + u * px - x3, y1 + u * py - y3);
}

// Human.get_g_xy (human_plus.py:19-52)
__device__ void door_goal(const SnbDoorCfg &door, bool door_rule, double px, double py, double fgx, double fgy, double &gx, double &gy)
{
    gx = fgx; gy = fgy;
    if (!door_rule) return;
    if (fmin(py, fgy) < door.door_y_mid_min && fmax(py, fgy) > door.door_y_mid_max) {
        const double igx = door.door_x_mid, igy = 0.5 * (door.door_y_min + door.door_y_max);
        if (npnorm2(igx - px, igy - py) <= door.door_width / 2.0) return;
        gx = igx; gy = igy;
    }
}

struct SceneArgs {
    SnbSceneCfg cfg;
    SnbDoorCfg door;
    SnbCrowdState st;
    const uint64_t *seeds;
    const double *segs;
    int n_seg;
    int32_t *n_draws;      // optional [B]: how many 64-bit draws env b consumed (parity evidence)
};

__global__ void scene_reset_kernel(const SceneArgs a)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.st.B) return;
    const int H = a.st.H;
    const SnbSceneCfg &c = a.cfg;
    const double PI = 3.141592653589793;
    const double r = c.human_radius;
    const double robot_px = 0.0, robot_py = -c.circle_radius, robot_gx = 0.0, robot_gy = c.circle_radius, robot_r = c.robot_radius;
    const bool door_rule = a.door.enabled && a.n_seg > 0;
    SeedSeqPcg64 rng;
    rng.seed(a.seeds[b]);
    int draws = 0;
    bool gave_up = false;          // the reference would spin forever on an over-crowded scene; the kernel stops and reports it
    const int MAX_TRIES = 200000;
    const size_t g0 = (size_t)b * H;
    for (int i = 0; i < H; ++i) {
        double v_pref = c.human_v_pref;
        double px, py, fgx, fgy, theta;
        if (c.rule == SNB_SCENE_CIRCLE_CROSSING) {
            // generate_circle_crossing_human (crowd_sim_plus.py:454-481)
            if (c.randomize_attributes) { v_pref = rng.uniform(0.5, 1.5); ++draws; }
            for (int tries = 0;; ++tries) {
                if (tries >= MAX_TRIES) { gave_up = true; break; }
                const double angle = rng.random() * PI * 2;
                const double px_noise = (rng.random() - 0.5) * v_pref;
                const double py_noise = (rng.random() - 0.5) * v_pref;
                draws += 3;
                px = c.circle_radius * cos(angle) + px_noise;
                py = c.circle_radius * sin(angle) + py_noise;
                bool collide = false;
                {   // [robot] + humans already placed; position and (current) goal both keep min_dist
                    const double md = r + robot_r + c.discomfort_dist;
                    collide = npnorm2(px - robot_px, py - robot_py) < md || npnorm2(px - robot_gx, py - robot_gy) < md;
                }
                for (int k = 0; k < i && !collide; ++k) {
                    const double md = r + a.st.radius[g0 + k] + c.discomfort_dist;
                    collide = npnorm2(px - a.st.px[g0 + k], py - a.st.py[g0 + k]) < md ||
                              npnorm2(px - a.st.gx[g0 + k], py - a.st.gy[g0 + k]) < md;
                }
                if (!collide) break;
            }
            fgx = -px; fgy = -py; theta = 0.0;
        } else {
            // generate_hallway_human (crowd_sim_plus.py:522-605)
            double eff_h = c.rect_height;
            for (int tries = 0;; ++tries) {
                if (tries >= MAX_TRIES) { gave_up = true; break; }
                if (c.randomize_attributes) { v_pref = rng.uniform(0.5, 1.5); ++draws; }
                const int dir_sign = rng.random() < 0.15 ? 1 : -1;
                const double right_num = dir_sign > 0 ? 0.8 : 1 - 0.8;
                const int wor_sign = rng.random() < right_num ? -1 : 1;
                double prob_cross = 0.3;
                if (rng.random() < right_num) prob_cross = 1 - prob_cross;
                const int cross_sign = rng.random() < prob_cross ? -wor_sign : wor_sign;
                px = rng.random() * 0.5 * wor_sign * (c.rect_width - r * 2);
                py = rng.random() * 0.25 * dir_sign * c.circle_radius * (eff_h - r * 2);
                draws += 6;
                bool collide = npnorm2(px - robot_px, py - robot_py) < r + robot_r + c.discomfort_dist;
                if (!collide) collide = npnorm2(px - robot_px, py - robot_py) < r + robot_r;
                for (int k = 0; k < i && !collide; ++k)
                    collide = npnorm2(px - a.st.px[g0 + k], py - a.st.py[g0 + k]) < r + a.st.radius[g0 + k];
                for (int s = 0; s < a.n_seg && !collide; ++s)
                    collide = fabs(point_to_segment_dist(a.segs[4 * s], a.segs[4 * s + 1], a.segs[4 * s + 2], a.segs[4 * s + 3], px, py)) < (r + 0.01);
                if (collide) { eff_h *= 1.1; continue; }
                fgx = rng.random() * 0.5 * cross_sign * (c.rect_width - r * 2);
                fgy = rng.random() * 0.5 * -dir_sign * c.circle_radius * (eff_h - r * 2);
                draws += 2;
                collide = npnorm2(fgx - robot_gx, fgy - robot_gy) < r + robot_r;
                for (int k = 0; k < i && !collide; ++k)
                    collide = npnorm2(fgx - a.st.gx[g0 + k], fgy - a.st.gy[g0 + k]) < r + a.st.radius[g0 + k];
                for (int s = 0; s < a.n_seg && !collide; ++s)
                    collide = fabs(point_to_segment_dist(a.segs[4 * s], a.segs[4 * s + 1], a.segs[4 * s + 2], a.segs[4 * s + 3], fgx, fgy)) < r;
                if (!collide) break;
                eff_h *= 1.1;
            }
            theta = atan2(fgy - py, fgx - px);
        }
        double gx, gy;
        door_goal(a.door, door_rule, px, py, fgx, fgy, gx, gy);
        const size_t g = g0 + i;
        a.st.px[g] = px; a.st.py[g] = py; a.st.vx[g] = 0.0; a.st.vy[g] = 0.0; a.st.theta[g] = theta;
        a.st.gx[g] = gx; a.st.gy[g] = gy; a.st.fgx[g] = fgx; a.st.fgy[g] = fgy;
        a.st.vpref[g] = v_pref; a.st.radius[g] = r; a.st.human_time[g] = 0.0;
    }
    // robot at (0, -R) heading for (0, R) (crowd_sim_plus.py:661), clocks
    if (a.st.E > 0) {
        const size_t e = (size_t)b * a.st.E;
        a.st.ex_px[e] = robot_px; a.st.ex_py[e] = robot_py; a.st.ex_vx[e] = 0.0; a.st.ex_vy[e] = 0.0; a.st.ex_radius[e] = robot_r;
    }
    if (a.st.rtheta) a.st.rtheta[b] = PI / 2;
    if (a.st.rgx) { a.st.rgx[b] = robot_gx; a.st.rgy[b] = robot_gy; }
    if (a.st.global_time) a.st.global_time[b] = 0.0;
    if (a.st.prev_dist) a.st.prev_dist[b] = npnorm2(robot_px - robot_gx, robot_py - robot_gy);
    if (a.n_draws) a.n_draws[b] = gave_up ? -1 : draws;
}

} // namespace

extern "C" int snb_scene_reset(const SnbSceneCfg *cfg, const SnbDoorCfg *door, const SnbCrowdState *state, const uint64_t *seeds_dev,
                               const double *segs_dev, int32_t n_seg, int32_t *n_draws_dev, void *stream)
{
    SNB_REQUIRE(cfg && state && seeds_dev, SNB_EINVAL, "scene reset: cfg/state/seeds is NULL");
    SNB_REQUIRE(cfg->rule == SNB_SCENE_CIRCLE_CROSSING || cfg->rule == SNB_SCENE_HALLWAY, SNB_EUNSUPPORTED,
                "scene reset: rule %d unsupported (square_crossing is broken in the reference, quirk q9)", cfg->rule);
    SNB_REQUIRE(state->B >= 0 && state->H >= 1, SNB_EINVAL, "scene reset: bad sizes B=%d H=%d", state->B, state->H);
    SNB_REQUIRE(state->px && state->py && state->vx && state->vy && state->theta && state->gx && state->gy && state->fgx && state->fgy &&
                state->vpref && state->radius && state->human_time, SNB_EINVAL, "scene reset: NULL human array");
    SNB_REQUIRE(n_seg == 0 || segs_dev, SNB_EINVAL, "scene reset: segs is NULL");
    if (state->B == 0) return SNB_OK;
    SceneArgs a;
    a.cfg = *cfg;
    if (door) a.door = *door; else { memset(&a.door, 0, sizeof(a.door)); }
    a.st = *state; a.seeds = seeds_dev; a.segs = segs_dev; a.n_seg = n_seg; a.n_draws = n_draws_dev;
    const int threads = 64;
    scene_reset_kernel<<<(state->B + threads - 1) / threads, threads, 0, (cudaStream_t)stream>>>(a);
    snb_count_launch();
    SNB_CUDA_TRY(cudaGetLastError());
    return SNB_OK;
}
