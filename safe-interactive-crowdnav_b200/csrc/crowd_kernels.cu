// csrc/crowd_kernels.cu -- the CrowdSimPlus per-agent step on sm_100a.
//
// One launch = one CrowdSimPlus.step for B independent environments (reference:
// crowd_sim_plus/envs/crowd_sim_plus.py:1025-1257).  Design:
//   * a CTA owns a tile of EPC consecutive environments; their fp64 SoA state is staged into shared memory
//     by 1-D TMA bulk copies (cp.async.bulk + mbarrier), one per state array;
//   * one WARP per human evaluates the policy: lanes = the other agents.  ORCA: each lane builds one
//     neighbour's half-plane, the incremental 2-D LP runs warp-cooperatively (the scan over previous lines in
//     linearProgram1 is a shuffle min/max reduction).  SFM: each lane adds one pairwise / wall force and the
//     sum is a shuffle reduction;
//   * phase 2 (one thread per agent / per environment): static-obstacle clamp, robot collision scan, reward,
//     flags, integration, written back coalesced.
// Numerics: this file is compiled with -fmad=false.  ORCA runs in fp32 exactly where Python-RVO2 does (every
// float op is a single IEEE op in the order of RVO2's Agent.cpp) so it is bit-identical to oracle/rvo2_oracle.c;
// everything else is fp64 like the reference's Python, with explicit fma() only where numpy's BLAS dot fuses.
/*
 * The ORCA parts of this file (half-plane construction, linearProgram1-3, neighbour selection, obstacle BSP walk) follow the RVO2 Library (v2.0.x: Agent.cpp, KdTree.cpp, RVOSimulator.cpp),
 *   Copyright 2008 University of North Carolina at Chapel Hill,
 *   licensed under the Apache License, Version 2.0 (http://www.apache.org/licenses/LICENSE-2.0).
 * RVO2 is distributed on an "AS IS" BASIS, WITHOUT WARRANTIES OR CONDITIONS OF ANY KIND; see the License for the specific
 * language governing permissions and limitations.  <https://gamma.cs.unc.edu/RVO2/>   This file is a derived restatement, not a copy.
 */
#include <cfloat>
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "snb_common.h"

#define RVO_EPSILON 0.00001f
#define FULL 0xffffffffu

// ---------------------------------------------------------------------------------------------------------
// obstacle data (host-built, device-resident)
// ---------------------------------------------------------------------------------------------------------
struct ObstVert { float px, py, ux, uy; int next, prev, convex, pad; };
struct BspNode { int obstacle, left, right, pad; };

struct SnbObstacles {
    int n_seg = 0;
    std::vector<double> segs;      // host copy [n_seg*4]
    std::vector<ObstVert> verts;   // after BSP splitting
    std::vector<BspNode> nodes;
    int root = -1;
    double *d_segs = nullptr;
    ObstVert *d_verts = nullptr;
    BspNode *d_nodes = nullptr;
};

// ---- host float helpers, same op order as RVO2's Vector2.h (this TU is built with -fmad=false; the host
// compiler gets -ffp-contract=off through -Xcompiler) ----
namespace hostrvo {
struct V2 { float x, y; };
static inline V2 sub(V2 a, V2 b) { return {a.x - b.x, a.y - b.y}; }
static inline float det(V2 a, V2 b) { return a.x * b.y - a.y * b.x; }
static inline float leftOf(V2 a, V2 b, V2 c) { return det(sub(a, c), sub(b, a)); }
static inline V2 normalize(V2 a) { const float inv = 1.0f / sqrtf(a.x * a.x + a.y * a.y); return {a.x * inv, a.y * inv}; }

// KdTree::buildObstacleTreeRecursive (RVO2 KdTree.cpp).  Returns node index or -1.
static int buildObstacleTree(SnbObstacles &o, const std::vector<int> &obstacles)
{
    if (obstacles.empty()) return -1;
    const size_t n = obstacles.size();
    size_t optimalSplit = 0, minLeft = n, minRight = n;
    auto worse_or_equal = [](size_t l, size_t r, size_t ml, size_t mr) {
        const size_t a1 = std::max(l, r), a2 = std::min(l, r), b1 = std::max(ml, mr), b2 = std::min(ml, mr);
        return !(a1 < b1 || (a1 == b1 && a2 < b2));
    };
    for (size_t i = 0; i < n; ++i) {
        size_t leftSize = 0, rightSize = 0;
        const V2 i1{o.verts[obstacles[i]].px, o.verts[obstacles[i]].py};
        const ObstVert &vi2 = o.verts[o.verts[obstacles[i]].next];
        const V2 i2{vi2.px, vi2.py};
        for (size_t j = 0; j < n; ++j) {
            if (i == j) continue;
            const V2 j1{o.verts[obstacles[j]].px, o.verts[obstacles[j]].py};
            const ObstVert &vj2 = o.verts[o.verts[obstacles[j]].next];
            const V2 j2{vj2.px, vj2.py};
            const float j1LeftOfI = leftOf(i1, i2, j1), j2LeftOfI = leftOf(i1, i2, j2);
            if (j1LeftOfI >= -RVO_EPSILON && j2LeftOfI >= -RVO_EPSILON) ++leftSize;
            else if (j1LeftOfI <= RVO_EPSILON && j2LeftOfI <= RVO_EPSILON) ++rightSize;
            else { ++leftSize; ++rightSize; }
            if (worse_or_equal(leftSize, rightSize, minLeft, minRight)) break;
        }
        if (!worse_or_equal(leftSize, rightSize, minLeft, minRight)) { minLeft = leftSize; minRight = rightSize; optimalSplit = i; }
    }
    std::vector<int> leftObst, rightObst;
    const size_t i = optimalSplit;
    const int I1 = obstacles[i];
    for (size_t j = 0; j < n; ++j) {
        if (i == j) continue;
        const int J1 = obstacles[j];
        const int J2 = o.verts[J1].next;
        const int I2 = o.verts[I1].next;
        const V2 i1{o.verts[I1].px, o.verts[I1].py}, i2{o.verts[I2].px, o.verts[I2].py};
        const V2 j1{o.verts[J1].px, o.verts[J1].py}, j2{o.verts[J2].px, o.verts[J2].py};
        const float j1LeftOfI = leftOf(i1, i2, j1), j2LeftOfI = leftOf(i1, i2, j2);
        if (j1LeftOfI >= -RVO_EPSILON && j2LeftOfI >= -RVO_EPSILON) leftObst.push_back(J1);
        else if (j1LeftOfI <= RVO_EPSILON && j2LeftOfI <= RVO_EPSILON) rightObst.push_back(J1);
        else {
            const float t = det(sub(i2, i1), sub(j1, i1)) / det(sub(i2, i1), sub(j1, j2));
            const V2 d = sub(j2, j1);
            ObstVert nv{};
            nv.px = j1.x + t * d.x; nv.py = j1.y + t * d.y;
            nv.prev = J1; nv.next = J2; nv.convex = 1;
            nv.ux = o.verts[J1].ux; nv.uy = o.verts[J1].uy;
            const int nid = (int)o.verts.size();
            o.verts.push_back(nv);
            o.verts[J1].next = nid; o.verts[J2].prev = nid;
            if (j1LeftOfI > 0.0f) { leftObst.push_back(J1); rightObst.push_back(nid); }
            else { rightObst.push_back(J1); leftObst.push_back(nid); }
        }
    }
    const int me = (int)o.nodes.size();
    o.nodes.push_back(BspNode{I1, -1, -1, 0});
    const int l = buildObstacleTree(o, leftObst);
    const int r = buildObstacleTree(o, rightObst);
    o.nodes[me].left = l; o.nodes[me].right = r;
    return me;
}
} // namespace hostrvo

extern "C" int snb_obstacles_create(SnbObstacles **out, const double *segs, int32_t n_seg)
{
    SNB_REQUIRE(out != nullptr, SNB_EINVAL, "snb_obstacles_create: out is NULL");
    SNB_REQUIRE(n_seg >= 0 && n_seg <= SNB_MAX_SEGMENTS, SNB_EUNSUPPORTED, "snb_obstacles_create: n_seg=%d beyond SNB_MAX_SEGMENTS", n_seg);
    SNB_REQUIRE(n_seg == 0 || segs != nullptr, SNB_EINVAL, "snb_obstacles_create: segs is NULL");
    SnbObstacles *o = new SnbObstacles();
    o->n_seg = n_seg;
    o->segs.assign(segs, segs + 4 * (size_t)n_seg);
    // RVOSimulator::addObstacle for every 2-vertex wall, in segment order (orca_plus.py:50-51)
    for (int k = 0; k < n_seg; ++k) {
        const hostrvo::V2 p0{(float)segs[4 * k], (float)segs[4 * k + 1]}, p1{(float)segs[4 * k + 2], (float)segs[4 * k + 3]};
        const int id0 = (int)o->verts.size(), id1 = id0 + 1;
        const hostrvo::V2 u0 = hostrvo::normalize(hostrvo::sub(p1, p0)), u1 = hostrvo::normalize(hostrvo::sub(p0, p1));
        o->verts.push_back(ObstVert{p0.x, p0.y, u0.x, u0.y, id1, id1, 1, 0});
        o->verts.push_back(ObstVert{p1.x, p1.y, u1.x, u1.y, id0, id0, 1, 0});
    }
    if (n_seg > 0) { // processObstacles (orca_plus.py:52-53)
        std::vector<int> all(o->verts.size());
        for (size_t i = 0; i < all.size(); ++i) all[i] = (int)i;
        o->root = hostrvo::buildObstacleTree(*o, all);
        cudaError_t e = cudaMalloc(&o->d_segs, sizeof(double) * 4 * (size_t)n_seg);
        if (e == cudaSuccess) e = cudaMalloc(&o->d_verts, sizeof(ObstVert) * o->verts.size());
        if (e == cudaSuccess) e = cudaMalloc(&o->d_nodes, sizeof(BspNode) * o->nodes.size());
        if (e == cudaSuccess) e = cudaMemcpy(o->d_segs, o->segs.data(), sizeof(double) * 4 * (size_t)n_seg, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(o->d_verts, o->verts.data(), sizeof(ObstVert) * o->verts.size(), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(o->d_nodes, o->nodes.data(), sizeof(BspNode) * o->nodes.size(), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            snb_set_error("snb_obstacles_create: %s", cudaGetErrorString(e));
            snb_obstacles_destroy(o);
            return SNB_ECUDA;
        }
    }
    *out = o;
    return SNB_OK;
}

extern "C" int snb_obstacles_destroy(SnbObstacles *o)
{
    if (!o) return SNB_OK;
    cudaFree(o->d_segs); cudaFree(o->d_verts); cudaFree(o->d_nodes);
    delete o;
    return SNB_OK;
}

extern "C" int32_t snb_obstacles_num_vertices(const SnbObstacles *o) { return o ? (int32_t)o->verts.size() : 0; }

extern "C" int snb_obstacles_get_vertex(const SnbObstacles *o, int32_t i, float *out7)
{
    SNB_REQUIRE(o && out7 && i >= 0 && i < (int)o->verts.size(), SNB_EINVAL, "snb_obstacles_get_vertex: bad index");
    const ObstVert &v = o->verts[i];
    out7[0] = v.px; out7[1] = v.py; out7[2] = v.ux; out7[3] = v.uy; out7[4] = (float)v.next; out7[5] = (float)v.prev; out7[6] = (float)v.convex;
    return SNB_OK;
}

// ---------------------------------------------------------------------------------------------------------
// device code
// ---------------------------------------------------------------------------------------------------------
struct CrowdParams {
    SnbPolicyCfg cfg;
    SnbDoorCfg door;
    SnbRewardCfg rcfg;
    SnbCrowdState st;
    const double *robot_action;
    const uint8_t *active;
    double *reward, *dmin, *out_v;
    int *flags, *nbr, *nbr_cnt, *status;
    int n_seg;
    const double *segs;
    int n_vert;
    const ObstVert *verts;
    const BspNode *nodes;
    int bsp_root;
    int epc;       // environments per CTA
    int full_step; // 0 = policy only, 1 = CrowdSimPlus.step(update=True), 2 = step(update=False) for n_actions candidate robot actions
    int n_actions; // full_step == 2: robot_action is [B, n_actions, 2]; reward / dmin / flags are [B, n_actions]
    double *next_h;     // full_step == 2: [B, H, 4] next observable human states (px, py, vx, vy)
    double *next_robot; // full_step == 2: [B, n_actions, 2] constrained next robot position (optional)
    int thread_mode;    // phase 1: 0 = one warp per human (small launches), 1 = one thread per human (large batches)
    int fast_lines;     // thread mode, plain ORCA: half-planes of every thread in shared memory (this many per thread), no local memory
    int lp3_mode;       // warp-owns-environments kernel: 0 = lp3_warp (loop per half-plane), 1 = lp3_warp_skip (ballot scans)
    double *log;        // optional state log [B, log_L, H + 1, 2]: slot log_slot <- positions before the step (humans, then the robot)
    int log_L, log_slot;
};

struct Line { float px, py, dx, dy; };

__device__ __forceinline__ float det2(float ax, float ay, float bx, float by) { return ax * by - ay * bx; }
__device__ __forceinline__ float dot2(float ax, float ay, float bx, float by) { return ax * bx + ay * by; }
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ Line shfl_line(const Line &l, int src) {
    Line r;
    r.px = __shfl_sync(FULL, l.px, src); r.py = __shfl_sync(FULL, l.py, src);
    r.dx = __shfl_sync(FULL, l.dx, src); r.dy = __shfl_sync(FULL, l.dy, src);
    return r;
}

// RVO2 linearProgram2 with linearProgram1 inlined, executed by a whole warp (Agent.cpp).
// Lane k holds line k (k < n <= 32).  `rx, ry` are warp-uniform.  The scan over the previous lines of
// linearProgram1 becomes a min/max shuffle reduction: tLeft only grows and tRight only shrinks along the scan,
// so "some prefix has tLeft > tRight" == "the final pair has", and min/max are exact, hence the result is
// bit-identical to the sequential program.
__device__ int lp2_warp(const Line &my, int n, float radius, float optx, float opty, bool dirOpt, float &rx, float &ry, int lane)
{
    if (dirOpt) { rx = optx * radius; ry = opty * radius; }
    else if (dot2(optx, opty, optx, opty) > radius * radius) {
        const float inv = 1.0f / sqrtf(dot2(optx, opty, optx, opty));
        const float nx = optx * inv, ny = opty * inv;
        rx = nx * radius; ry = ny * radius;
    } else { rx = optx; ry = opty; }

    for (int i = 0; i < n; ++i) {
        const Line li = shfl_line(my, i);
        if (det2(li.dx, li.dy, li.px - rx, li.py - ry) > 0.0f) {
            const float tx = rx, ty = ry;
            const float dotProduct = dot2(li.px, li.py, li.dx, li.dy);
            const float discriminant = dotProduct * dotProduct + radius * radius - dot2(li.px, li.py, li.px, li.py);
            bool fail = discriminant < 0.0f;
            float tLeft = 0.f, tRight = 0.f;
            if (!fail) {
                const float sq = sqrtf(discriminant);
                tLeft = -dotProduct - sq;
                tRight = -dotProduct + sq;
                float myL = -INFINITY, myR = INFINITY;
                bool pfail = false;
                if (lane < i) {
                    const float denominator = det2(li.dx, li.dy, my.dx, my.dy);
                    const float numerator = det2(my.dx, my.dy, li.px - my.px, li.py - my.py);
                    if (fabsf(denominator) <= RVO_EPSILON) { if (numerator < 0.0f) pfail = true; }
                    else {
                        const float t = numerator / denominator;
                        if (denominator >= 0.0f) myR = t; else myL = t;
                    }
                }
                tRight = fminf(tRight, warp_min(myR));
                tLeft = fmaxf(tLeft, warp_max(myL));
                fail = (__any_sync(FULL, pfail) != 0) || (tLeft > tRight);
            }
            if (fail) { rx = tx; ry = ty; return i; }
            if (dirOpt) {
                if (dot2(optx, opty, li.dx, li.dy) > 0.0f) { rx = li.px + tRight * li.dx; ry = li.py + tRight * li.dy; }
                else { rx = li.px + tLeft * li.dx; ry = li.py + tLeft * li.dy; }
            } else {
                const float t = dot2(li.dx, li.dy, optx - li.px, opty - li.py);
                if (t < tLeft) { rx = li.px + tLeft * li.dx; ry = li.py + tLeft * li.dy; }
                else if (t > tRight) { rx = li.px + tRight * li.dx; ry = li.py + tRight * li.dy; }
                else { rx = li.px + t * li.dx; ry = li.py + t * li.dy; }
            }
        }
    }
    return n;
}

// RVO2 linearProgram3, warp-cooperative.  `scratch` = 32 Lines of per-warp shared memory.
__device__ void lp3_warp(const Line &my, int n, int numObstLines, int beginLine, float radius, float &rx, float &ry,
                         int lane, Line *scratch)
{
    float distance = 0.0f;
    for (int i = beginLine; i < n; ++i) {
        const Line li = shfl_line(my, i);
        if (det2(li.dx, li.dy, li.px - rx, li.py - ry) > distance) {
            bool have = false;
            Line pl = my;
            if (lane < numObstLines) have = true;
            else if (lane < i) {
                const float determinant = det2(li.dx, li.dy, my.dx, my.dy);
                have = true;
                if (fabsf(determinant) <= RVO_EPSILON) {
                    if (dot2(li.dx, li.dy, my.dx, my.dy) > 0.0f) have = false;
                    else { pl.px = 0.5f * (li.px + my.px); pl.py = 0.5f * (li.py + my.py); }
                } else {
                    const float s = det2(my.dx, my.dy, li.px - my.px, li.py - my.py) / determinant;
                    pl.px = li.px + s * li.dx; pl.py = li.py + s * li.dy;
                }
                if (have) {
                    const float vx = my.dx - li.dx, vy = my.dy - li.dy;
                    const float inv = 1.0f / sqrtf(dot2(vx, vy, vx, vy));
                    pl.dx = vx * inv; pl.dy = vy * inv;
                }
            }
            const unsigned mask = __ballot_sync(FULL, have);
            const int np = __popc(mask);
            if (have) scratch[__popc(mask & ((1u << lane) - 1u))] = pl;
            __syncwarp();
            Line mine = (lane < np) ? scratch[lane] : Line{0.f, 0.f, 1.f, 0.f};
            __syncwarp();
            const float tx = rx, ty = ry;
            if (lp2_warp(mine, np, radius, -li.dy, li.dx, true, rx, ry, lane) < np) { rx = tx; ry = ty; }
            distance = det2(li.dx, li.dy, li.px - rx, li.py - ry);
        }
    }
}

__device__ __forceinline__ float redux_min(float v) { float r; asm volatile("redux.sync.min.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v)); return r; }
__device__ __forceinline__ float redux_max(float v) { float r; asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v)); return r; }

// linearProgram3 for the warp-owns-environments kernel.  Same arithmetic per half-plane as lp3_warp / lp3_serial, but the two
// sequential scans ("next half-plane the current point violates") are one ballot each instead of one loop trip per half-plane:
// the point only moves when a linearProgram1 succeeds, so every half-plane between two moves is tested against the same point
// and the first set bit of the ballot is exactly the half-plane the sequential loop would stop at.  Half-planes are read from
// shared memory as broadcasts (L[j * 32], the failing human's column of the warp's line store), min / max of the LP1 scan are
// single CREDUX instructions.  `scratch` = 32 float4 of per-warp shared memory for the projected half-planes.
__device__ void lp3_warp_skip(const float4 *L, int n, int beginLine, float radius, float &rx, float &ry, int lane, float4 *scratch)
{
    Line my; my.px = 0.f; my.py = 0.f; my.dx = 1.f; my.dy = 0.f;
    if (lane < n) { const float4 v = L[lane * 32]; my.px = v.x; my.py = v.y; my.dx = v.z; my.dy = v.w; }
    const unsigned lt = (1u << lane) - 1u;
    float distance = 0.0f;
    int i = beginLine;
    while (true) {
        const bool viol = lane >= i && lane < n && det2(my.dx, my.dy, my.px - rx, my.py - ry) > distance;
        const unsigned vm = __ballot_sync(FULL, viol);
        if (!vm) break;
        i = __ffs(vm) - 1;
        Line li; { const float4 v = L[i * 32]; li.px = v.x; li.py = v.y; li.dx = v.z; li.dy = v.w; }
        bool have = false;
        Line pl = my;
        if (lane < i) {
            const float determinant = det2(li.dx, li.dy, my.dx, my.dy);
            have = true;
            if (fabsf(determinant) <= RVO_EPSILON) {
                if (dot2(li.dx, li.dy, my.dx, my.dy) > 0.0f) have = false;
                else { pl.px = 0.5f * (li.px + my.px); pl.py = 0.5f * (li.py + my.py); }
            } else {
                const float sv = det2(my.dx, my.dy, li.px - my.px, li.py - my.py) / determinant;
                pl.px = li.px + sv * li.dx; pl.py = li.py + sv * li.dy;
            }
            if (have) {
                const float vx = my.dx - li.dx, vy = my.dy - li.dy;
                const float inv = 1.0f / sqrtf(dot2(vx, vy, vx, vy));
                pl.dx = vx * inv; pl.dy = vy * inv;
            }
        }
        const unsigned hm = __ballot_sync(FULL, have);
        const int np = __popc(hm);
        if (have) scratch[__popc(hm & lt)] = make_float4(pl.px, pl.py, pl.dx, pl.dy);
        __syncwarp();
        Line mine; mine.px = 0.f; mine.py = 0.f; mine.dx = 1.f; mine.dy = 0.f;
        if (lane < np) { const float4 v = scratch[lane]; mine.px = v.x; mine.py = v.y; mine.dx = v.z; mine.dy = v.w; }
        // linearProgram2(projected, radius, (-li.dy, li.dx), directionOpt = true)
        const float ox = -li.dy, oy = li.dx;
        float qx = ox * radius, qy = oy * radius;
        bool ok = true;
        int k = 0;
        while (true) {
            const bool vv = lane >= k && lane < np && det2(mine.dx, mine.dy, mine.px - qx, mine.py - qy) > 0.0f;
            const unsigned m2 = __ballot_sync(FULL, vv);
            if (!m2) break;
            k = __ffs(m2) - 1;
            Line lk; { const float4 v = scratch[k]; lk.px = v.x; lk.py = v.y; lk.dx = v.z; lk.dy = v.w; }
            const float dotProduct = dot2(lk.px, lk.py, lk.dx, lk.dy);
            const float discriminant = dotProduct * dotProduct + radius * radius - dot2(lk.px, lk.py, lk.px, lk.py);
            if (discriminant < 0.0f) { ok = false; break; }
            const float sq = sqrtf(discriminant);
            float tLeft = -dotProduct - sq;
            float tRight = -dotProduct + sq;
            float myL = -INFINITY, myR = INFINITY;
            bool pfail = false;
            if (lane < k) {
                const float denominator = det2(lk.dx, lk.dy, mine.dx, mine.dy);
                const float numerator = det2(mine.dx, mine.dy, lk.px - mine.px, lk.py - mine.py);
                if (fabsf(denominator) <= RVO_EPSILON) { if (numerator < 0.0f) pfail = true; }
                else {
                    const float t = numerator / denominator;
                    if (denominator >= 0.0f) myR = t; else myL = t;
                }
            }
            tRight = fminf(tRight, redux_min(myR));
            tLeft = fmaxf(tLeft, redux_max(myL));
            if ((__any_sync(FULL, pfail) != 0) || (tLeft > tRight)) { ok = false; break; }
            if (dot2(ox, oy, lk.dx, lk.dy) > 0.0f) { qx = lk.px + tRight * lk.dx; qy = lk.py + tRight * lk.dy; }
            else { qx = lk.px + tLeft * lk.dx; qy = lk.py + tLeft * lk.dy; }
            ++k;
        }
        __syncwarp();                                       // everybody has read `scratch` before the next projection overwrites it
        if (ok) { rx = qx; ry = qy; }
        distance = det2(li.dx, li.dy, li.px - rx, li.py - ry);
        ++i;
    }
}

// Agent ORCA half-plane for one neighbour (Agent::computeNewVelocity agent part).
__device__ Line agent_orca_line(float px, float py, float vx, float vy, float radius, float opx, float opy, float ovx,
                                float ovy, float orad, float invTimeHorizon, float timeStep)
{
    const float rpx = opx - px, rpy = opy - py;
    const float rvx = vx - ovx, rvy = vy - ovy;
    const float distSq = dot2(rpx, rpy, rpx, rpy);
    const float combinedRadius = radius + orad;
    const float combinedRadiusSq = combinedRadius * combinedRadius;
    Line line;
    float ux, uy;
    if (distSq > combinedRadiusSq) {
        const float wx = rvx - invTimeHorizon * rpx, wy = rvy - invTimeHorizon * rpy;
        const float wLengthSq = dot2(wx, wy, wx, wy);
        const float dotProduct1 = dot2(wx, wy, rpx, rpy);
        if (dotProduct1 < 0.0f && dotProduct1 * dotProduct1 > combinedRadiusSq * wLengthSq) {
            const float wLength = sqrtf(wLengthSq);
            const float inv = 1.0f / wLength;
            const float unx = wx * inv, uny = wy * inv;
            line.dx = uny; line.dy = -unx;
            const float s = combinedRadius * invTimeHorizon - wLength;
            ux = s * unx; uy = s * uny;
        } else {
            const float leg = sqrtf(distSq - combinedRadiusSq);
            const float inv = 1.0f / distSq;
            if (det2(rpx, rpy, wx, wy) > 0.0f) {
                line.dx = (rpx * leg - rpy * combinedRadius) * inv;
                line.dy = (rpx * combinedRadius + rpy * leg) * inv;
            } else {
                line.dx = -((rpx * leg + rpy * combinedRadius) * inv);
                line.dy = -((-rpx * combinedRadius + rpy * leg) * inv);
            }
            const float dotProduct2 = dot2(rvx, rvy, line.dx, line.dy);
            ux = dotProduct2 * line.dx - rvx; uy = dotProduct2 * line.dy - rvy;
        }
    } else {
        const float invTimeStep = 1.0f / timeStep;
        const float wx = rvx - invTimeStep * rpx, wy = rvy - invTimeStep * rpy;
        const float wLength = sqrtf(dot2(wx, wy, wx, wy));
        const float inv = 1.0f / wLength;
        const float unx = wx * inv, uny = wy * inv;
        line.dx = uny; line.dy = -unx;
        const float s = combinedRadius * invTimeStep - wLength;
        ux = s * unx; uy = s * uny;
    }
    line.px = vx + 0.5f * ux; line.py = vy + 0.5f * uy;
    return line;
}

// Obstacle ORCA half-plane for obstacle neighbour `o1i` (Agent::computeNewVelocity obstacle part).
// Returns false when RVO2 `continue`s without adding a line.
__device__ bool obstacle_orca_line(const ObstVert *verts, int o1i, float px, float py, float vx, float vy, float radius,
                                   float invT, Line &line)
{
    int o2i = verts[o1i].next;
    ObstVert o1 = verts[o1i], o2 = verts[o2i];
    const float rp1x = o1.px - px, rp1y = o1.py - py, rp2x = o2.px - px, rp2y = o2.py - py;
    const float distSq1 = dot2(rp1x, rp1y, rp1x, rp1y), distSq2 = dot2(rp2x, rp2y, rp2x, rp2y);
    const float radiusSq = radius * radius;
    const float ovx = o2.px - o1.px, ovy = o2.py - o1.py;
    const float s = dot2(-rp1x, -rp1y, ovx, ovy) / dot2(ovx, ovy, ovx, ovy);
    const float tx_ = -rp1x - s * ovx, ty_ = -rp1y - s * ovy;
    const float distSqLine = dot2(tx_, ty_, tx_, ty_);

    if (s < 0.0f && distSq1 <= radiusSq) {
        if (o1.convex) {
            line.px = 0.f; line.py = 0.f;
            const float inv = 1.0f / sqrtf(dot2(-rp1y, rp1x, -rp1y, rp1x));
            line.dx = -rp1y * inv; line.dy = rp1x * inv;
            return true;
        }
        return false;
    } else if (s > 1.0f && distSq2 <= radiusSq) {
        if (o2.convex && det2(rp2x, rp2y, o2.ux, o2.uy) >= 0.0f) {
            line.px = 0.f; line.py = 0.f;
            const float inv = 1.0f / sqrtf(dot2(-rp2y, rp2x, -rp2y, rp2x));
            line.dx = -rp2y * inv; line.dy = rp2x * inv;
            return true;
        }
        return false;
    } else if (s >= 0.0f && s < 1.0f && distSqLine <= radiusSq) {
        line.px = 0.f; line.py = 0.f; line.dx = -o1.ux; line.dy = -o1.uy;
        return true;
    }

    float llx, lly, rlx, rly; // left / right leg directions
    if (s < 0.0f && distSqLine <= radiusSq) {
        if (!o1.convex) return false;
        o2 = o1; o2i = o1i;
        const float leg1 = sqrtf(distSq1 - radiusSq);
        const float inv = 1.0f / distSq1;
        llx = (rp1x * leg1 - rp1y * radius) * inv; lly = (rp1x * radius + rp1y * leg1) * inv;
        rlx = (rp1x * leg1 + rp1y * radius) * inv; rly = (-rp1x * radius + rp1y * leg1) * inv;
    } else if (s > 1.0f && distSqLine <= radiusSq) {
        if (!o2.convex) return false;
        o1 = o2; o1i = o2i;
        const float leg2 = sqrtf(distSq2 - radiusSq);
        const float inv = 1.0f / distSq2;
        llx = (rp2x * leg2 - rp2y * radius) * inv; lly = (rp2x * radius + rp2y * leg2) * inv;
        rlx = (rp2x * leg2 + rp2y * radius) * inv; rly = (-rp2x * radius + rp2y * leg2) * inv;
    } else {
        if (o1.convex) {
            const float leg1 = sqrtf(distSq1 - radiusSq);
            const float inv = 1.0f / distSq1;
            llx = (rp1x * leg1 - rp1y * radius) * inv; lly = (rp1x * radius + rp1y * leg1) * inv;
        } else { llx = -o1.ux; lly = -o1.uy; }
        if (o2.convex) {
            const float leg2 = sqrtf(distSq2 - radiusSq);
            const float inv = 1.0f / distSq2;
            rlx = (rp2x * leg2 + rp2y * radius) * inv; rly = (-rp2x * radius + rp2y * leg2) * inv;
        } else { rlx = o1.ux; rly = o1.uy; }
    }

    const ObstVert leftNeighbor = verts[o1.prev];
    bool isLeftLegForeign = false, isRightLegForeign = false;
    if (o1.convex && det2(llx, lly, -leftNeighbor.ux, -leftNeighbor.uy) >= 0.0f) {
        llx = -leftNeighbor.ux; lly = -leftNeighbor.uy; isLeftLegForeign = true;
    }
    if (o2.convex && det2(rlx, rly, o2.ux, o2.uy) <= 0.0f) {
        rlx = o2.ux; rly = o2.uy; isRightLegForeign = true;
    }

    const float lcx = invT * (o1.px - px), lcy = invT * (o1.py - py);
    const float rcx = invT * (o2.px - px), rcy = invT * (o2.py - py);
    const float cvx = rcx - lcx, cvy = rcy - lcy;
    const bool same = (o1i == o2i);
    const float t = same ? 0.5f : dot2(vx - lcx, vy - lcy, cvx, cvy) / dot2(cvx, cvy, cvx, cvy);
    const float tLeft = dot2(vx - lcx, vy - lcy, llx, lly);
    const float tRight = dot2(vx - rcx, vy - rcy, rlx, rly);
    const float rs = radius * invT;

    if ((t < 0.0f && tLeft < 0.0f) || (same && tLeft < 0.0f && tRight < 0.0f)) {
        const float wx = vx - lcx, wy = vy - lcy;
        const float inv = 1.0f / sqrtf(dot2(wx, wy, wx, wy));
        const float unx = wx * inv, uny = wy * inv;
        line.dx = uny; line.dy = -unx;
        line.px = lcx + rs * unx; line.py = lcy + rs * uny;
        return true;
    } else if (t > 1.0f && tRight < 0.0f) {
        const float wx = vx - rcx, wy = vy - rcy;
        const float inv = 1.0f / sqrtf(dot2(wx, wy, wx, wy));
        const float unx = wx * inv, uny = wy * inv;
        line.dx = uny; line.dy = -unx;
        line.px = rcx + rs * unx; line.py = rcy + rs * uny;
        return true;
    }

    float distSqCutoff, distSqLeft, distSqRight;
    if (t < 0.0f || t > 1.0f || same) distSqCutoff = INFINITY;
    else { const float ax = vx - (lcx + t * cvx), ay = vy - (lcy + t * cvy); distSqCutoff = dot2(ax, ay, ax, ay); }
    if (tLeft < 0.0f) distSqLeft = INFINITY;
    else { const float ax = vx - (lcx + tLeft * llx), ay = vy - (lcy + tLeft * lly); distSqLeft = dot2(ax, ay, ax, ay); }
    if (tRight < 0.0f) distSqRight = INFINITY;
    else { const float ax = vx - (rcx + tRight * rlx), ay = vy - (rcy + tRight * rly); distSqRight = dot2(ax, ay, ax, ay); }

    if (distSqCutoff <= distSqLeft && distSqCutoff <= distSqRight) {
        line.dx = -o1.ux; line.dy = -o1.uy;
        line.px = lcx + rs * (-line.dy); line.py = lcy + rs * line.dx;
        return true;
    } else if (distSqLeft <= distSqRight) {
        if (isLeftLegForeign) return false;
        line.dx = llx; line.dy = lly;
        line.px = lcx + rs * (-line.dy); line.py = lcy + rs * line.dx;
        return true;
    } else {
        if (isRightLegForeign) return false;
        line.dx = -rlx; line.dy = -rly;
        line.px = rcx + rs * (-line.dy); line.py = rcy + rs * line.dx;
        return true;
    }
}

// distSqPointLineSegment (RVO2 Definitions / Vector2.h)
__device__ float dist_sq_point_segment(float ax, float ay, float bx, float by, float cx, float cy)
{
    const float r = dot2(cx - ax, cy - ay, bx - ax, by - ay) / dot2(bx - ax, by - ay, bx - ax, by - ay);
    if (r < 0.0f) return dot2(cx - ax, cy - ay, cx - ax, cy - ay);
    if (r > 1.0f) return dot2(cx - bx, cy - by, cx - bx, cy - by);
    const float qx = cx - (ax + r * (bx - ax)), qy = cy - (ay + r * (by - ay));
    return dot2(qx, qy, qx, qy);
}

// KdTree::queryObstacleTreeRecursive + Agent::insertObstacleNeighbor, iterative, run by ONE lane.
// ids/ds = per-warp shared scratch (SNB_MAX_ORCA_LINES entries).  Returns the neighbour count (-1 on overflow).
__device__ int obstacle_neighbors(const ObstVert *verts, const BspNode *nodes, int root, float px, float py, float rangeSq,
                                  int *ids, float *ds)
{
    int cnt = 0;
    int stack_node[48];
    unsigned char stack_stage[48];
    int sp = 0;
    if (root < 0) return 0;
    stack_node[0] = root; stack_stage[0] = 0; sp = 1;
    while (sp > 0) {
        const int node = stack_node[sp - 1];
        const int stage = stack_stage[sp - 1];
        const BspNode nd = nodes[node];
        const ObstVert o1 = verts[nd.obstacle];
        const ObstVert o2 = verts[o1.next];
        // leftOf(o1, o2, p) = det(o1 - p, o2 - o1)
        const float agentLeftOfLine = det2(o1.px - px, o1.py - py, o2.px - o1.px, o2.py - o1.py);
        if (stage == 0) {
            stack_stage[sp - 1] = 1;
            const int first = (agentLeftOfLine >= 0.0f ? nd.left : nd.right);
            if (first >= 0) { if (sp >= 48) return -1; stack_node[sp] = first; stack_stage[sp] = 0; ++sp; }
            continue;
        }
        --sp; // stage 1: this node, then maybe the far side
        const float ex = o2.px - o1.px, ey = o2.py - o1.py;
        const float distSqLine = agentLeftOfLine * agentLeftOfLine / dot2(ex, ey, ex, ey);
        if (distSqLine < rangeSq) {
            if (agentLeftOfLine < 0.0f) {
                const float distSq = dist_sq_point_segment(o1.px, o1.py, o2.px, o2.py, px, py);
                if (distSq < rangeSq) {
                    if (cnt >= SNB_MAX_ORCA_LINES) return -1;
                    int i = cnt++;
                    while (i != 0 && distSq < ds[i - 1]) { ds[i] = ds[i - 1]; ids[i] = ids[i - 1]; --i; }
                    ds[i] = distSq; ids[i] = nd.obstacle;
                }
            }
            const int other = (agentLeftOfLine >= 0.0f ? nd.right : nd.left);
            if (other >= 0) { if (sp >= 48) return -1; stack_node[sp] = other; stack_stage[sp] = 0; ++sp; }
        }
    }
    return cnt;
}

// KdTree::buildAgentTree + an unpruned queryAgentTreeRecursive from agent 0: writes, for every agent
// k of the throw-away simulator (0 = self, 1.. = `ob` order), its position in the visit sequence.  Only needed to
// order EXACTLY equidistant neighbours when the simulator has more than MAX_LEAF_SIZE=10 agents; run by one lane.
__device__ void kd_visit_rank(int n, const float *ax, const float *ay, unsigned char *vrank)
{
    unsigned char idx[SNB_MAX_AGENTS_PER_ENV + 1];
    struct Node { unsigned char begin, end, left, right; float minX, maxX, minY, maxY; };
    Node tree[2 * (SNB_MAX_AGENTS_PER_ENV + 1)];
    for (int i = 0; i < n; ++i) idx[i] = (unsigned char)i;
    // build (pre-order, explicit stack)
    unsigned char sb[40], se[40], sn[40];
    int sp = 0;
    sb[0] = 0; se[0] = (unsigned char)n; sn[0] = 0; sp = 1;
    while (sp > 0) {
        --sp;
        const int begin = sb[sp], end = se[sp], node = sn[sp];
        Node &t = tree[node];
        t.begin = (unsigned char)begin; t.end = (unsigned char)end; t.left = t.right = 0;
        t.minX = t.maxX = ax[idx[begin]]; t.minY = t.maxY = ay[idx[begin]];
        for (int i = begin + 1; i < end; ++i) {
            t.maxX = fmaxf(t.maxX, ax[idx[i]]); t.minX = fminf(t.minX, ax[idx[i]]);
            t.maxY = fmaxf(t.maxY, ay[idx[i]]); t.minY = fminf(t.minY, ay[idx[i]]);
        }
        if (end - begin > 10) {
            const bool isVertical = (t.maxX - t.minX > t.maxY - t.minY);
            const float splitValue = isVertical ? 0.5f * (t.maxX + t.minX) : 0.5f * (t.maxY + t.minY);
            int left = begin, right = end;
            while (left < right) {
                while (left < right && (isVertical ? ax[idx[left]] : ay[idx[left]]) < splitValue) ++left;
                while (right > left && (isVertical ? ax[idx[right - 1]] : ay[idx[right - 1]]) >= splitValue) --right;
                if (left < right) { const unsigned char tmp = idx[left]; idx[left] = idx[right - 1]; idx[right - 1] = tmp; ++left; --right; }
            }
            if (left == begin) { ++left; ++right; }
            t.left = (unsigned char)(node + 1);
            t.right = (unsigned char)(node + 2 * (left - begin));
            // push right first so that left is built first (order is irrelevant for the result)
            sb[sp] = (unsigned char)left; se[sp] = (unsigned char)end; sn[sp] = t.right; ++sp;
            sb[sp] = (unsigned char)begin; se[sp] = (unsigned char)left; sn[sp] = t.left; ++sp;
        }
    }
    // query from agent 0 without pruning: nearer child first (strict <), see queryAgentTreeRecursive
    const float qx = ax[0], qy = ay[0];
    int pos = 0;
    sn[0] = 0; sp = 1;
    while (sp > 0) {
        const int node = sn[--sp];
        const Node &t = tree[node];
        if (t.end - t.begin <= 10) {
            for (int i = t.begin; i < t.end; ++i) vrank[idx[i]] = (unsigned char)pos++;
        } else {
            const Node &L = tree[t.left], &R = tree[t.right];
            float a, b, c, d;
            a = fmaxf(0.0f, L.minX - qx); b = fmaxf(0.0f, qx - L.maxX); c = fmaxf(0.0f, L.minY - qy); d = fmaxf(0.0f, qy - L.maxY);
            const float distSqLeft = a * a + b * b + c * c + d * d;
            a = fmaxf(0.0f, R.minX - qx); b = fmaxf(0.0f, qx - R.maxX); c = fmaxf(0.0f, R.minY - qy); d = fmaxf(0.0f, qy - R.maxY);
            const float distSqRight = a * a + b * b + c * c + d * d;
            if (distSqLeft < distSqRight) { sn[sp++] = t.right; sn[sp++] = t.left; }
            else { sn[sp++] = t.left; sn[sp++] = t.right; }
        }
    }
}

// per-warp shared scratch
struct WarpScratch {
    Line lines[SNB_MAX_ORCA_LINES];
    int nb[SNB_MAX_AGENTS_PER_ENV];
    int obst_ids[SNB_MAX_ORCA_LINES];
    float obst_ds[SNB_MAX_ORCA_LINES];
    float ax[SNB_MAX_AGENTS_PER_ENV + 1], ay[SNB_MAX_AGENTS_PER_ENV + 1];
    unsigned char vrank[SNB_MAX_AGENTS_PER_ENV + 4];
};

// shared-memory view of the CTA's environment tile (all fp64)
struct Tile {
    double *px, *py, *vx, *vy, *rad, *gx, *gy, *vpref; // [epc*H]
    double *ex_px, *ex_py, *ex_vx, *ex_vy, *ex_rad;   // [epc*E]
    double *act;                                        // [epc*H*2] human actions
    double *segs;                                       // [n_seg*4]
};

// candidate c (0..n_others-1) of human i in local env e -> observable state (ob order: other humans, then extras)
__device__ __forceinline__ void load_other(const Tile &T, int H, int E, int e, int i, int c, double &opx, double &opy,
                                           double &ovx, double &ovy, double &orad, int &agent_id)
{
    if (c < H - 1) {
        const int j = (c < i) ? c : c + 1;
        const int k = e * H + j;
        opx = T.px[k]; opy = T.py[k]; ovx = T.vx[k]; ovy = T.vy[k]; orad = T.rad[k];
        agent_id = j;
    } else {
        const int x = c - (H - 1);
        const int k = e * E + x;
        opx = T.ex_px[k]; opy = T.ex_py[k]; ovx = T.ex_vx[k]; ovy = T.ex_vy[k]; orad = T.ex_rad[k];
        agent_id = H + x;
    }
}

// ORCA.predict / ORCAPlus.predict for one human, by one warp.  Returns the new velocity (warp-uniform).
__device__ void orca_predict_warp(const CrowdParams &P, const Tile &T, WarpScratch &W, int e, int i, int lane, int genv,
                                  float &out_vx, float &out_vy)
{
    const SnbPolicyCfg &cfg = P.cfg;
    const int H = P.st.H, E = P.st.E;
    const int n_others = H - 1 + P.st.n_obs_extras;
    const int k = e * H + i;
    const double dpx = T.px[k], dpy = T.py[k];
    const float px = (float)dpx, py = (float)dpy, vx = (float)T.vx[k], vy = (float)T.vy[k];
    const float radius = (float)(T.rad[k] + 0.01 + cfg.safety_space);   // orca.py:100
    const float maxSpeed = (float)T.vpref[k];
    const float neighborDist = (float)cfg.neighbor_dist;
    const float timeHorizon = (float)cfg.time_horizon, timeHorizonObst = (float)cfg.time_horizon_obst;
    const float timeStep = (float)cfg.time_step;
    const int maxNeighbors = cfg.max_neighbors;

    // preferred velocity in double, then narrowed (orca.py:113-123 / orca_plus.py:68-79)
    const double dvx = T.gx[k] - dpx, dvy = T.gy[k] - dpy;
    const double speed = sqrt(fma(dvy, dvy, dvx * dvx)); // np.linalg.norm: BLAS dot fuses
    double pvx, pvy;
    if (cfg.policy == SNB_POLICY_ORCA_PLUS) {
        const double vp = T.vpref[k] - 1e-3;
        if (speed > vp) { pvx = dvx / speed * vp; pvy = dvy / speed * vp; } else { pvx = dvx; pvy = dvy; }
    } else {
        if (speed > 1) { pvx = dvx / speed; pvy = dvy / speed; } else { pvx = dvx; pvy = dvy; }
    }
    const float prefx = (float)pvx, prefy = (float)pvy;

    // ---- agent neighbours: lanes = candidates ----
    float opx = 0.f, opy = 0.f, ovx = 0.f, ovy = 0.f, orad = 0.f;
    int agent_id = -1;
    bool in_range = false;
    float distSq = INFINITY;
    if (lane < n_others) {
        double a, b, c, d, r;
        load_other(T, H, E, e, i, lane, a, b, c, d, r, agent_id);
        opx = (float)a; opy = (float)b; ovx = (float)c; ovy = (float)d;
        orad = (float)(r + 0.01 + cfg.safety_space);
        const float ddx = px - opx, ddy = py - opy;
        distSq = dot2(ddx, ddy, ddx, ddy);
        in_range = (maxNeighbors > 0) && (distSq < neighborDist * neighborDist);
    }
    int rank = 0;
    bool tie = false;
    for (int j = 0; j < n_others; ++j) {
        const float dj = __shfl_sync(FULL, distSq, j);
        const bool inj = __shfl_sync(FULL, (int)in_range, j) != 0;
        if (inj && (dj < distSq || (dj == distSq && j < lane))) ++rank;
        if (inj && in_range && j != lane && dj == distSq) tie = true;
    }
    if (__any_sync(FULL, tie) && n_others + 1 > 10) {
        // exact distance tie in a simulator with a split kd-tree: order the tied agents by RVO2's visit sequence
        if (lane == 0) { W.ax[0] = px; W.ay[0] = py; }
        if (lane < n_others) { W.ax[lane + 1] = opx; W.ay[lane + 1] = opy; }
        __syncwarp();
        if (lane == 0) kd_visit_rank(n_others + 1, W.ax, W.ay, W.vrank);
        __syncwarp();
        const int myv = (lane < n_others) ? W.vrank[lane + 1] : 0;
        rank = 0;
        for (int j = 0; j < n_others; ++j) {
            const float dj = __shfl_sync(FULL, distSq, j);
            const bool inj = __shfl_sync(FULL, (int)in_range, j) != 0;
            const int vj = __shfl_sync(FULL, myv, j);
            if (inj && (dj < distSq || (dj == distSq && vj < myv))) ++rank;
        }
        __syncwarp();
    }
    const unsigned inmask = __ballot_sync(FULL, in_range);
    const int n_in = __popc(inmask);
    const int n_nb = n_in < maxNeighbors ? n_in : maxNeighbors;
    if (in_range && rank < n_nb) W.nb[rank] = lane;
    __syncwarp();
    if (P.nbr_cnt && lane == 0) P.nbr_cnt[genv * H + i] = n_nb;

    // ---- obstacle neighbours and lines (ORCAPlus only) ----
    int numObstLines = 0;
    bool overflow = false;
    if (cfg.policy == SNB_POLICY_ORCA_PLUS && P.n_vert > 0) {
        int n_on = 0;
        if (lane == 0) {
            const float rs = timeHorizonObst * maxSpeed + radius;
            n_on = obstacle_neighbors(P.verts, P.nodes, P.bsp_root, px, py, rs * rs, W.obst_ids, W.obst_ds);
        }
        n_on = __shfl_sync(FULL, n_on, 0);
        __syncwarp();
        if (n_on < 0) { overflow = true; n_on = 0; }
        const float invT = 1.0f / timeHorizonObst;
        Line cand = Line{0.f, 0.f, 1.f, 0.f};
        bool valid = false;
        float r1x = 0.f, r1y = 0.f, r2x = 0.f, r2y = 0.f;
        if (lane < n_on) {
            const int o1 = W.obst_ids[lane];
            const ObstVert a = P.verts[o1], b = P.verts[a.next];
            r1x = a.px - px; r1y = a.py - py; r2x = b.px - px; r2y = b.py - py;
            valid = obstacle_orca_line(P.verts, o1, px, py, vx, vy, radius, invT, cand);
        }
        // sequential "already covered" sweep (depends only on the lines accepted so far)
        bool accepted = false;
        for (int q = 0; q < n_on; ++q) {
            const float q1x = __shfl_sync(FULL, r1x, q), q1y = __shfl_sync(FULL, r1y, q);
            const float q2x = __shfl_sync(FULL, r2x, q), q2y = __shfl_sync(FULL, r2y, q);
            bool covers = false;
            if (accepted && lane < q) {
                covers = (det2(invT * q1x - cand.px, invT * q1y - cand.py, cand.dx, cand.dy) - invT * radius >= -RVO_EPSILON) &&
                         (det2(invT * q2x - cand.px, invT * q2y - cand.py, cand.dx, cand.dy) - invT * radius >= -RVO_EPSILON);
            }
            const bool covered = __any_sync(FULL, covers) != 0;
            if (lane == q) accepted = valid && !covered;
        }
        const unsigned amask = __ballot_sync(FULL, accepted);
        numObstLines = __popc(amask);
        if (accepted) W.lines[__popc(amask & ((1u << lane) - 1u))] = cand;
    }

    // ---- agent lines: lane q builds the half-plane of neighbour q ----
    int nLines = numObstLines + n_nb;
    if (nLines > SNB_MAX_ORCA_LINES) { overflow = true; nLines = SNB_MAX_ORCA_LINES; }
    {
        const int src = (lane < n_nb) ? W.nb[lane] : 0;
        const float qpx = __shfl_sync(FULL, opx, src), qpy = __shfl_sync(FULL, opy, src);
        const float qvx = __shfl_sync(FULL, ovx, src), qvy = __shfl_sync(FULL, ovy, src);
        const float qr = __shfl_sync(FULL, orad, src);
        const int qid = __shfl_sync(FULL, agent_id, src);
        if (lane < n_nb && numObstLines + lane < SNB_MAX_ORCA_LINES)
            W.lines[numObstLines + lane] = agent_orca_line(px, py, vx, vy, radius, qpx, qpy, qvx, qvy, qr, 1.0f / timeHorizon, timeStep);
        if (P.nbr && lane < maxNeighbors) P.nbr[(size_t)(genv * H + i) * maxNeighbors + lane] = (lane < n_nb) ? qid : -1;
    }
    __syncwarp();
    const Line my = (lane < nLines) ? W.lines[lane] : Line{0.f, 0.f, 1.f, 0.f};
    __syncwarp();
    if (overflow && lane == 0 && P.status) atomicExch(P.status, SNB_EOVERFLOW);

    float rx, ry;
    const int lineFail = lp2_warp(my, nLines, maxSpeed, prefx, prefy, false, rx, ry, lane);
    if (lineFail < nLines) lp3_warp(my, nLines, numObstLines, lineFail, maxSpeed, rx, ry, lane, W.lines);
    out_vx = rx; out_vy = ry;
}

// ---------------------------------------------------------------------------------------------------------------------
// Thread-serial ORCA (large batches): ONE THREAD per human runs RVO2's computeNeighbors + computeNewVelocity exactly as the
// sequential program does (Agent.cpp linearProgram1 / 2 / 3) -- same float operations in the same order as
// the warp-cooperative path above, hence the same bits.  The warp path minimises the latency of one small launch (32 lanes
// share one human); this path maximises throughput when there are enough humans to give every thread its own (ncu of the
// warp path at 262 144 envs: 24 % of the warp slots occupied, issue slots 43 % busy, most lanes idle behind <= 10 neighbours).
// Where a thread's ORCA lines live.  PtrLines: a per-thread array (local memory).  SmemLines: shared memory, interleaved so that
// line l of thread t is the float4 at base[l * CROWD_THREADS] (base already offset by t): conflict-free 16-byte accesses, and none of
// the local-memory traffic that the first thread-mode kernel spilled to L2 / DRAM (ncu r02: 1.8 GB written per 2^20-env step against
// 0.65 GB of state).
constexpr int CROWD_THREADS_C = 256;
struct PtrLines {
    const Line *p;
    __device__ __forceinline__ Line operator[](int i) const { return p[i]; }
};
template <int STRIDE>
struct SmemLinesT {
    float4 *base;
    __device__ __forceinline__ Line operator[](int i) const { const float4 v = base[i * STRIDE]; Line l; l.px = v.x; l.py = v.y; l.dx = v.z; l.dy = v.w; return l; }
    __device__ __forceinline__ void set(int i, const Line &l) const { base[i * STRIDE] = make_float4(l.px, l.py, l.dx, l.dy); }
};
typedef SmemLinesT<CROWD_THREADS_C> SmemLines;

template <class LA>
__device__ bool lp1_serial(const LA &lines, int lineNo, float radius, float optx, float opty, bool dirOpt, float &rx, float &ry)
{
    const Line li = lines[lineNo];
    const float dotProduct = dot2(li.px, li.py, li.dx, li.dy);
    const float discriminant = dotProduct * dotProduct + radius * radius - dot2(li.px, li.py, li.px, li.py);
    if (discriminant < 0.0f) return false;
    const float sq = sqrtf(discriminant);
    float tLeft = -dotProduct - sq;
    float tRight = -dotProduct + sq;
    for (int j = 0; j < lineNo; ++j) {
        const Line lj = lines[j];
        const float denominator = det2(li.dx, li.dy, lj.dx, lj.dy);
        const float numerator = det2(lj.dx, lj.dy, li.px - lj.px, li.py - lj.py);
        if (fabsf(denominator) <= RVO_EPSILON) {
            if (numerator < 0.0f) return false;
            continue;
        }
        const float t = numerator / denominator;
        if (denominator >= 0.0f) tRight = fminf(tRight, t);
        else tLeft = fmaxf(tLeft, t);
        if (tLeft > tRight) return false;
    }
    if (dirOpt) {
        if (dot2(optx, opty, li.dx, li.dy) > 0.0f) { rx = li.px + tRight * li.dx; ry = li.py + tRight * li.dy; }
        else { rx = li.px + tLeft * li.dx; ry = li.py + tLeft * li.dy; }
    } else {
        const float t = dot2(li.dx, li.dy, optx - li.px, opty - li.py);
        if (t < tLeft) { rx = li.px + tLeft * li.dx; ry = li.py + tLeft * li.dy; }
        else if (t > tRight) { rx = li.px + tRight * li.dx; ry = li.py + tRight * li.dy; }
        else { rx = li.px + t * li.dx; ry = li.py + t * li.dy; }
    }
    return true;
}

template <class LA>
__device__ int lp2_serial(const LA &lines, int n, float radius, float optx, float opty, bool dirOpt, float &rx, float &ry)
{
    if (dirOpt) { rx = optx * radius; ry = opty * radius; }
    else if (dot2(optx, opty, optx, opty) > radius * radius) {
        const float inv = 1.0f / sqrtf(dot2(optx, opty, optx, opty));
        const float nx = optx * inv, ny = opty * inv;
        rx = nx * radius; ry = ny * radius;
    } else { rx = optx; ry = opty; }
    for (int i = 0; i < n; ++i) {
        const Line li = lines[i];
        if (det2(li.dx, li.dy, li.px - rx, li.py - ry) > 0.0f) {
            const float tx = rx, ty = ry;
            if (!lp1_serial(lines, i, radius, optx, opty, dirOpt, rx, ry)) { rx = tx; ry = ty; return i; }
        }
    }
    return n;
}

template <class LA>
__device__ void lp3_serial(const LA &lines, int n, int numObstLines, int beginLine, float radius, float &rx, float &ry, Line *proj)
{
    float distance = 0.0f;
    for (int i = beginLine; i < n; ++i) {
        const Line li = lines[i];
        if (det2(li.dx, li.dy, li.px - rx, li.py - ry) > distance) {
            int np = 0;
            for (int j = 0; j < numObstLines; ++j) proj[np++] = lines[j];
            for (int j = numObstLines; j < i; ++j) {
                const Line lj = lines[j];
                Line pl;
                const float determinant = det2(li.dx, li.dy, lj.dx, lj.dy);
                if (fabsf(determinant) <= RVO_EPSILON) {
                    if (dot2(li.dx, li.dy, lj.dx, lj.dy) > 0.0f) continue;
                    pl.px = 0.5f * (li.px + lj.px); pl.py = 0.5f * (li.py + lj.py);
                } else {
                    const float sv = det2(lj.dx, lj.dy, li.px - lj.px, li.py - lj.py) / determinant;
                    pl.px = li.px + sv * li.dx; pl.py = li.py + sv * li.dy;
                }
                const float vx = lj.dx - li.dx, vy = lj.dy - li.dy;
                const float inv = 1.0f / sqrtf(dot2(vx, vy, vx, vy));
                pl.dx = vx * inv; pl.dy = vy * inv;
                proj[np++] = pl;
            }
            const float tx = rx, ty = ry;
            if (lp2_serial(PtrLines{proj}, np, radius, -li.dy, li.dx, true, rx, ry) < np) { rx = tx; ry = ty; }
            distance = det2(li.dx, li.dy, li.px - rx, li.py - ry);
        }
    }
}

// (goal and v_pref of the human are parameters: the warp-owns-environments kernel keeps them in registers, not in the tile)
__device__ void orca_predict_thread(const CrowdParams &P, const Tile &T, int e, int i, int genv, double goal_x, double goal_y, double v_pref,
                                    float &out_vx, float &out_vy)
{
    const SnbPolicyCfg &cfg = P.cfg;
    const int H = P.st.H, E = P.st.E;
    const int n_others = H - 1 + P.st.n_obs_extras;
    const int k = e * H + i;
    const double dpx = T.px[k], dpy = T.py[k];
    const float px = (float)dpx, py = (float)dpy, vx = (float)T.vx[k], vy = (float)T.vy[k];
    const float radius = (float)(T.rad[k] + 0.01 + cfg.safety_space);   // orca.py:100
    const float maxSpeed = (float)v_pref;
    const float neighborDist = (float)cfg.neighbor_dist;
    const float timeHorizon = (float)cfg.time_horizon, timeHorizonObst = (float)cfg.time_horizon_obst;
    const float timeStep = (float)cfg.time_step;
    const int maxNeighbors = cfg.max_neighbors;

    // preferred velocity in double, then narrowed (orca.py:113-123 / orca_plus.py:68-79)
    const double dvx = goal_x - dpx, dvy = goal_y - dpy;
    const double speed = sqrt(fma(dvy, dvy, dvx * dvx)); // np.linalg.norm: BLAS dot fuses
    double pvx, pvy;
    if (cfg.policy == SNB_POLICY_ORCA_PLUS) {
        const double vp = v_pref - 1e-3;
        if (speed > vp) { pvx = dvx / speed * vp; pvy = dvy / speed * vp; } else { pvx = dvx; pvy = dvy; }
    } else {
        if (speed > 1) { pvx = dvx / speed; pvy = dvy / speed; } else { pvx = dvx; pvy = dvy; }
    }
    const float prefx = (float)pvx, prefy = (float)pvy;

    // ---- agent neighbours: the (<= maxNeighbors) nearest inside neighborDist, sorted by (distSq, RVO2 visit order) ----
    float dsq[SNB_MAX_AGENTS_PER_ENV];
    unsigned inmask = 0;
    for (int c = 0; c < n_others; ++c) {
        double a, b, cc, d, r;
        int id;
        load_other(T, H, E, e, i, c, a, b, cc, d, r, id);
        const float ddx = px - (float)a, ddy = py - (float)b;
        dsq[c] = dot2(ddx, ddy, ddx, ddy);
        if (maxNeighbors > 0 && dsq[c] < neighborDist * neighborDist) inmask |= 1u << c;
    }
    bool tie = false;
    for (int c = 0; c < n_others && !tie; ++c)
        for (int j = c + 1; j < n_others; ++j)
            if (((inmask >> c) & 1u) && ((inmask >> j) & 1u) && dsq[j] == dsq[c]) { tie = true; break; }
    unsigned char vrank[SNB_MAX_AGENTS_PER_ENV + 4];
    const bool use_visit = tie && n_others + 1 > 10;
    if (use_visit) {
        // exact distance tie in a simulator with a split kd-tree: order the tied agents by RVO2's visit sequence
        float ax[SNB_MAX_AGENTS_PER_ENV + 1], ay[SNB_MAX_AGENTS_PER_ENV + 1];
        ax[0] = px; ay[0] = py;
        for (int c = 0; c < n_others; ++c) {
            double a, b, cc, d, r;
            int id;
            load_other(T, H, E, e, i, c, a, b, cc, d, r, id);
            ax[c + 1] = (float)a; ay[c + 1] = (float)b;
        }
        kd_visit_rank(n_others + 1, ax, ay, vrank);
    }
    const int n_in = __popc(inmask);
    const int n_nb = n_in < maxNeighbors ? n_in : maxNeighbors;
    int nb[SNB_MAX_AGENTS_PER_ENV];
    for (int c = 0; c < n_others; ++c) {
        if (!((inmask >> c) & 1u)) continue;
        int rank = 0;
        for (int j = 0; j < n_others; ++j) {
            if (!((inmask >> j) & 1u)) continue;
            const bool before = use_visit ? (vrank[j + 1] < vrank[c + 1]) : (j < c);
            if (dsq[j] < dsq[c] || (dsq[j] == dsq[c] && before)) ++rank;
        }
        if (rank < n_nb) nb[rank] = c;
    }
    if (P.nbr_cnt) P.nbr_cnt[genv * H + i] = n_nb;

    Line lines[SNB_MAX_ORCA_LINES];
    // ---- obstacle neighbours and lines (ORCAPlus only) ----
    int numObstLines = 0;
    bool overflow = false;
    if (cfg.policy == SNB_POLICY_ORCA_PLUS && P.n_vert > 0) {
        int obst_ids[SNB_MAX_ORCA_LINES];
        float obst_ds[SNB_MAX_ORCA_LINES];
        const float rs = timeHorizonObst * maxSpeed + radius;
        int n_on = obstacle_neighbors(P.verts, P.nodes, P.bsp_root, px, py, rs * rs, obst_ids, obst_ds);
        if (n_on < 0) { overflow = true; n_on = 0; }
        const float invT = 1.0f / timeHorizonObst;
        for (int q = 0; q < n_on; ++q) {
            const int o1 = obst_ids[q];
            const ObstVert a = P.verts[o1], b = P.verts[a.next];
            const float r1x = a.px - px, r1y = a.py - py, r2x = b.px - px, r2y = b.py - py;
            bool covered = false;
            for (int j = 0; j < numObstLines && !covered; ++j)
                covered = (det2(invT * r1x - lines[j].px, invT * r1y - lines[j].py, lines[j].dx, lines[j].dy) - invT * radius >= -RVO_EPSILON) &&
                          (det2(invT * r2x - lines[j].px, invT * r2y - lines[j].py, lines[j].dx, lines[j].dy) - invT * radius >= -RVO_EPSILON);
            if (covered) continue;
            Line cand;
            if (obstacle_orca_line(P.verts, o1, px, py, vx, vy, radius, invT, cand)) {
                if (numObstLines < SNB_MAX_ORCA_LINES) lines[numObstLines++] = cand; else overflow = true;
            }
        }
    }

    // ---- agent lines ----
    int nLines = numObstLines + n_nb;
    if (nLines > SNB_MAX_ORCA_LINES) { overflow = true; nLines = SNB_MAX_ORCA_LINES; }
    for (int r = 0; r < n_nb; ++r) {
        double a, b, cc, d, rr;
        int id;
        load_other(T, H, E, e, i, nb[r], a, b, cc, d, rr, id);
        if (numObstLines + r < SNB_MAX_ORCA_LINES)
            lines[numObstLines + r] = agent_orca_line(px, py, vx, vy, radius, (float)a, (float)b, (float)cc, (float)d,
                                                      (float)(rr + 0.01 + cfg.safety_space), 1.0f / timeHorizon, timeStep);
        if (P.nbr) P.nbr[(size_t)(genv * H + i) * maxNeighbors + r] = id;
    }
    if (P.nbr) for (int r = n_nb; r < maxNeighbors; ++r) P.nbr[(size_t)(genv * H + i) * maxNeighbors + r] = -1;
    if (overflow && P.status) atomicExch(P.status, SNB_EOVERFLOW);

    float rx, ry;
    const int lineFail = lp2_serial(PtrLines{lines}, nLines, maxSpeed, prefx, prefy, false, rx, ry);
    if (lineFail < nLines) {
        Line proj[SNB_MAX_ORCA_LINES];
        lp3_serial(PtrLines{lines}, nLines, numObstLines, lineFail, maxSpeed, rx, ry, proj);
    }
    out_vx = rx; out_vy = ry;
}

// The same program for the common shape -- plain ORCA (no obstacle lines), <= FAST_MAXO observed agents, <= FAST_LCAP neighbours --
// with nothing in local memory: distances and ranks in registers (loops unrolled over FAST_MAXO), the half-planes in shared memory.
// Same float operations in the same order as orca_predict_thread, hence the same bits; an exact distance tie that needs RVO2's
// kd-tree visit order (simulators with > 10 agents) is handed to the general function.
constexpr int FAST_MAXO = 12, FAST_LCAP = 12;
// Returns -1 when the velocity is final, else lineFail: linearProgram2 stopped at that line and linearProgram3 still has to run on
// (lines, n_nb, maxSpeed, the LP2 result in out_v).  The caller compacts those humans (~10 % of a crowd) onto the first threads of the
// CTA before running LP3, so that a warp no longer pays LP3's long serial loops whenever ONE of its 32 lanes needs them.
template <int STRIDE>
__device__ int orca_predict_thread_fast(const CrowdParams &P, const Tile &T, float4 *s_lines, int e, int i, int genv, float &out_vx, float &out_vy,
                                        int &out_n, float &out_speed)
{
    const SnbPolicyCfg &cfg = P.cfg;
    const int H = P.st.H, E = P.st.E;
    const int n_others = H - 1 + P.st.n_obs_extras;
    const int k = e * H + i;
    const double dpx = T.px[k], dpy = T.py[k];
    const float px = (float)dpx, py = (float)dpy, vx = (float)T.vx[k], vy = (float)T.vy[k];
    const float radius = (float)(T.rad[k] + 0.01 + cfg.safety_space);
    const float maxSpeed = (float)T.vpref[k];
    const float neighborDist = (float)cfg.neighbor_dist;
    const float timeHorizon = (float)cfg.time_horizon;
    const float timeStep = (float)cfg.time_step;
    const int maxNeighbors = cfg.max_neighbors;

    const double dvx = T.gx[k] - dpx, dvy = T.gy[k] - dpy;
    const double speed = sqrt(fma(dvy, dvy, dvx * dvx));
    double pvx, pvy;
    if (cfg.policy == SNB_POLICY_ORCA_PLUS) {
        const double vp = T.vpref[k] - 1e-3;
        if (speed > vp) { pvx = dvx / speed * vp; pvy = dvy / speed * vp; } else { pvx = dvx; pvy = dvy; }
    } else {
        if (speed > 1) { pvx = dvx / speed; pvy = dvy / speed; } else { pvx = dvx; pvy = dvy; }
    }
    const float prefx = (float)pvx, prefy = (float)pvy;

    float dsq[FAST_MAXO];
    unsigned inmask = 0;
#pragma unroll
    for (int c = 0; c < FAST_MAXO; ++c) {
        dsq[c] = INFINITY;
        if (c < n_others) {
            double a, b, cc, d, r;
            int id;
            load_other(T, H, E, e, i, c, a, b, cc, d, r, id);
            const float ddx = px - (float)a, ddy = py - (float)b;
            dsq[c] = dot2(ddx, ddy, ddx, ddy);
            if (maxNeighbors > 0 && dsq[c] < neighborDist * neighborDist) inmask |= 1u << c;
        }
    }
    bool tie = false;
#pragma unroll
    for (int c = 0; c < FAST_MAXO; ++c)
#pragma unroll
        for (int j = c + 1; j < FAST_MAXO; ++j)
            tie = tie || ((((inmask >> c) & (inmask >> j)) & 1u) && dsq[j] == dsq[c]);
    if (tie && n_others + 1 > 10) { orca_predict_thread(P, T, e, i, genv, T.gx[k], T.gy[k], T.vpref[k], out_vx, out_vy); return -1; }
    const int n_in = __popc(inmask);
    const int n_nb = n_in < maxNeighbors ? n_in : maxNeighbors;
    int rk[FAST_MAXO];
#pragma unroll
    for (int c = 0; c < FAST_MAXO; ++c) {
        int rank = 0;
#pragma unroll
        for (int j = 0; j < FAST_MAXO; ++j)
            rank += (((inmask >> j) & 1u) && (dsq[j] < dsq[c] || (dsq[j] == dsq[c] && j < c))) ? 1 : 0;
        rk[c] = ((inmask >> c) & 1u) ? rank : 0x7fffffff;
    }
    if (P.nbr_cnt) P.nbr_cnt[genv * H + i] = n_nb;

    const SmemLinesT<STRIDE> lines{s_lines};
    for (int r = 0; r < n_nb; ++r) {
        int csel = 0;
#pragma unroll
        for (int c = 0; c < FAST_MAXO; ++c) csel = rk[c] == r ? c : csel;
        double a, b, cc, d, rr;
        int id;
        load_other(T, H, E, e, i, csel, a, b, cc, d, rr, id);
        lines.set(r, agent_orca_line(px, py, vx, vy, radius, (float)a, (float)b, (float)cc, (float)d,
                                     (float)(rr + 0.01 + cfg.safety_space), 1.0f / timeHorizon, timeStep));
        if (P.nbr) P.nbr[(size_t)(genv * H + i) * maxNeighbors + r] = id;
    }
    if (P.nbr) for (int r = n_nb; r < maxNeighbors; ++r) P.nbr[(size_t)(genv * H + i) * maxNeighbors + r] = -1;

    float rx, ry;
    const int lineFail = lp2_serial(lines, n_nb, maxSpeed, prefx, prefy, false, rx, ry);
    out_vx = rx; out_vy = ry;
    out_n = n_nb; out_speed = maxSpeed;
    return lineFail < n_nb ? lineFail : -1;
}

// ORCA up to linearProgram2 for the warp-owns-environments kernel: the same float operations as orca_predict_thread_fast, fed from a
// float copy of the warp's agents (fs = px, py, vx, vy; fr = radius + 0.01 + safety_space, each narrowed once by its owner instead of
// once per observer).  Neighbour order: rank = number of candidates that sort before (distSq, then candidate index), counted once per
// unordered pair; candidates outside neighborDist carry distinct negative keys, so they sort first (their count is subtracted) and
// never compare equal.  The half-planes are then built in candidate order and stored at their rank -- no search for "the r-th
// nearest".  Returns lineFail (linearProgram3 still to run), -1 (velocity final) or -2 (exact distance tie in a simulator of more
// than 10 agents: RVO2's kd-tree visit order decides, the caller runs the general function).
__device__ int orca_predict_wo(const CrowdParams &P, const float4 *fs, const float *fr, int self, int ebase, int robot_idx, int i, int genv,
                               double dvx, double dvy, double v_pref, float4 *s_lines, float &out_vx, float &out_vy, int &out_n, float &out_speed)
{
    const SnbPolicyCfg &cfg = P.cfg;
    const int H = P.st.H;
    const int n_others = H - 1 + P.st.n_obs_extras;
    const float4 me = fs[self];
    const float px = me.x, py = me.y, vx = me.z, vy = me.w;
    const float radius = fr[self];
    const float maxSpeed = (float)v_pref;
    const float neighborDist = (float)cfg.neighbor_dist;
    const float timeHorizon = (float)cfg.time_horizon;
    const float timeStep = (float)cfg.time_step;
    const int maxNeighbors = cfg.max_neighbors;

    const double speed = sqrt(fma(dvy, dvy, dvx * dvx));
    double pvx, pvy;
    if (cfg.policy == SNB_POLICY_ORCA_PLUS) {
        const double vp = v_pref - 1e-3;
        if (speed > vp) { pvx = dvx / speed * vp; pvy = dvy / speed * vp; } else { pvx = dvx; pvy = dvy; }
    } else {
        if (speed > 1) { pvx = dvx / speed; pvy = dvy / speed; } else { pvx = dvx; pvy = dvy; }
    }
    const float prefx = (float)pvx, prefy = (float)pvy;

    float dsq[FAST_MAXO];
    int n_in = 0;
    const float nd2 = neighborDist * neighborDist;
#pragma unroll
    for (int c = 0; c < FAST_MAXO; ++c) {
        float d = -(float)(c + 1);
        if (c < n_others) {
            const int idx = (c < H - 1) ? ebase + ((c < i) ? c : c + 1) : robot_idx;
            const float4 o = fs[idx];
            const float ddx = px - o.x, ddy = py - o.y;
            const float v = dot2(ddx, ddy, ddx, ddy);
            if (maxNeighbors > 0 && v < nd2) { d = v; ++n_in; }
        }
        dsq[c] = d;
    }
    int rank[FAST_MAXO];
#pragma unroll
    for (int c = 0; c < FAST_MAXO; ++c) rank[c] = 0;
    bool tie = false;
#pragma unroll
    for (int c = 1; c < FAST_MAXO; ++c)
#pragma unroll
        for (int j = 0; j < c; ++j) {
            const bool le = dsq[j] <= dsq[c];
            rank[c] += le ? 1 : 0;
            rank[j] += le ? 0 : 1;
            tie = tie || (dsq[j] == dsq[c]);
        }
    if (tie && n_others + 1 > 10) return -2;
    const int n_excl = FAST_MAXO - n_in;
    unsigned long long packed = 0ull;                       // 4 bits per candidate: its rank among the neighbours, 15 = not one
#pragma unroll
    for (int c = 0; c < FAST_MAXO; ++c)
        packed |= (unsigned long long)(dsq[c] >= 0.0f ? (unsigned)(rank[c] - n_excl) : 15u) << (4 * c);
    const int n_nb = n_in < maxNeighbors ? n_in : maxNeighbors;
    if (P.nbr_cnt) P.nbr_cnt[genv * H + i] = n_nb;

    const SmemLinesT<32> lines{s_lines};
    const float invTH = 1.0f / timeHorizon;
    for (int c = 0; c < n_others; ++c) {
        const int r = (int)((unsigned)(packed >> (4 * c)) & 15u);
        if (r < n_nb) {
            const bool human = c < H - 1;
            const int j = human ? ((c < i) ? c : c + 1) : H;
            const int idx = human ? ebase + j : robot_idx;
            const float4 o = fs[idx];
            lines.set(r, agent_orca_line(px, py, vx, vy, radius, o.x, o.y, o.z, o.w, fr[idx], invTH, timeStep));
            if (P.nbr) P.nbr[(size_t)(genv * H + i) * maxNeighbors + r] = j;
        }
    }
    if (P.nbr) for (int r = n_nb; r < maxNeighbors; ++r) P.nbr[(size_t)(genv * H + i) * maxNeighbors + r] = -1;

    float rx, ry;
    const int lineFail = lp2_serial(lines, n_nb, maxSpeed, prefx, prefy, false, rx, ry);
    out_vx = rx; out_vy = ry;
    out_n = n_nb; out_speed = maxSpeed;
    return lineFail < n_nb ? lineFail : -1;
}

// utils_plus.closest_point_on_segment (utils_plus.py:21-42)
__device__ __forceinline__ void closest_point_on_segment(double x1, double y1, double x2, double y2, double x3, double y3,
                                                         double &ox, double &oy)
{
    const double px = x2 - x1, py = y2 - y1;
    if (px == 0 && py == 0) { ox = x1; oy = y1; return; } // quirk q12: unreachable with valid layouts
    double u = ((x3 - x1) * px + (y3 - y1) * py) / (px * px + py * py);
    if (u > 1) u = 1; else if (u < 0) u = 0;
    ox = x1 + u * px; oy = y1 + u * py;
}

// SFM.predict for one human, by one warp (social_force.py:38-94).  fp64 like the reference; the pairwise and wall
// forces are summed with a shuffle reduction (summation order differs from the reference's sequential += by
// rounding only; tolerance stated in the tests).
__device__ void sfm_predict_warp(const CrowdParams &P, const Tile &T, int e, int i, int lane, double &out_vx, double &out_vy)
{
    const SnbPolicyCfg &cfg = P.cfg;
    const int H = P.st.H, E = P.st.E;
    const int n_others = H - 1 + P.st.n_obs_extras;
    const int k = e * H + i;
    const double px = T.px[k], py = T.py[k], vx = T.vx[k], vy = T.vy[k], radius = T.rad[k];
    const double gx = T.gx[k], gy = T.gy[k], v_pref = T.vpref[k];
    double fx = 0.0, fy = 0.0;
    for (int c = lane; c < n_others + P.n_seg; c += 32) {
        if (c < n_others) {
            double opx, opy, ovx, ovy, orad; int id;
            load_other(T, H, E, e, i, c, opx, opy, ovx, ovy, orad, id);
            const double adjustment = fabs(cfg.sfm_radius - orad) + 0.01;
            const double dx = px - opx, dy = py - opy;
            const double d = sqrt(dx * dx + dy * dy);
            const double ee = cfg.A * exp((radius + orad + adjustment - d) / cfg.B);
            fx += ee * (dx / d); fy += ee * (dy / d);
        } else {
            const int s = c - n_others;
            const double *L = T.segs + 4 * s;
            double As, Bs;
            if (cfg.is_bottleneck && s >= 2) { As = cfg.A_bottleneck; Bs = cfg.B_bottleneck; } else { As = cfg.A_static; Bs = cfg.B_static; }
            double ox, oy;
            closest_point_on_segment(L[0], L[1], L[2], L[3], px, py, ox, oy);
            const double dx = px - ox, dy = py - oy;
            const double d = sqrt(dx * dx + dy * dy);
            const double ee = As * exp((radius + 0.01 - d) / Bs);
            fx += ee * (dx / d); fy += ee * (dy / d);
        }
    }
    fx = warp_sum(fx); fy = warp_sum(fy);
    double ddx = gx - px, ddy = gy - py;
    double dist_to_goal = sqrt(ddx * ddx + ddy * ddy);
    dist_to_goal = dist_to_goal < 1e-6 ? 1.0 : dist_to_goal;
    const double desired_vx = (ddx / dist_to_goal) * v_pref, desired_vy = (ddy / dist_to_goal) * v_pref;
    const double cdx = cfg.KI * (desired_vx - vx), cdy = cfg.KI * (desired_vy - vy);
    const double new_vx = vx + (cdx + fx) * cfg.time_step, new_vy = vy + (cdy + fy) * cfg.time_step;
    const double act_norm = sqrt(fma(new_vy, new_vy, new_vx * new_vx)); // np.linalg.norm
    if (act_norm > v_pref) { out_vx = new_vx / act_norm * v_pref; out_vy = new_vy / act_norm * v_pref; }
    else { out_vx = new_vx; out_vy = new_vy; }
}

// SFM.predict for one human by ONE thread (large batches): the reference's own sequential accumulation order
// (social_force.py:58-80: other agents in `ob` order, then the segments).
__device__ void sfm_predict_thread(const CrowdParams &P, const Tile &T, int e, int i, double &out_vx, double &out_vy)
{
    const SnbPolicyCfg &cfg = P.cfg;
    const int H = P.st.H, E = P.st.E;
    const int n_others = H - 1 + P.st.n_obs_extras;
    const int k = e * H + i;
    const double px = T.px[k], py = T.py[k], vx = T.vx[k], vy = T.vy[k], radius = T.rad[k];
    const double gx = T.gx[k], gy = T.gy[k], v_pref = T.vpref[k];
    double fx = 0.0, fy = 0.0;
    for (int c = 0; c < n_others; ++c) {
        double opx, opy, ovx, ovy, orad; int id;
        load_other(T, H, E, e, i, c, opx, opy, ovx, ovy, orad, id);
        const double adjustment = fabs(cfg.sfm_radius - orad) + 0.01;
        const double dx = px - opx, dy = py - opy;
        const double d = sqrt(dx * dx + dy * dy);
        const double ee = cfg.A * exp((radius + orad + adjustment - d) / cfg.B);
        fx += ee * (dx / d); fy += ee * (dy / d);
    }
    for (int s = 0; s < P.n_seg; ++s) {
        const double *L = T.segs + 4 * s;
        double As, Bs;
        if (cfg.is_bottleneck && s >= 2) { As = cfg.A_bottleneck; Bs = cfg.B_bottleneck; } else { As = cfg.A_static; Bs = cfg.B_static; }
        double ox, oy;
        closest_point_on_segment(L[0], L[1], L[2], L[3], px, py, ox, oy);
        const double dx = px - ox, dy = py - oy;
        const double d = sqrt(dx * dx + dy * dy);
        const double ee = As * exp((radius + 0.01 - d) / Bs);
        fx += ee * (dx / d); fy += ee * (dy / d);
    }
    double ddx = gx - px, ddy = gy - py;
    double dist_to_goal = sqrt(ddx * ddx + ddy * ddy);
    dist_to_goal = dist_to_goal < 1e-6 ? 1.0 : dist_to_goal;
    const double desired_vx = (ddx / dist_to_goal) * v_pref, desired_vy = (ddy / dist_to_goal) * v_pref;
    const double cdx = cfg.KI * (desired_vx - vx), cdy = cfg.KI * (desired_vy - vy);
    const double new_vx = vx + (cdx + fx) * cfg.time_step, new_vy = vy + (cdy + fy) * cfg.time_step;
    const double act_norm = sqrt(fma(new_vy, new_vy, new_vx * new_vx)); // np.linalg.norm
    if (act_norm > v_pref) { out_vx = new_vx / act_norm * v_pref; out_vy = new_vy / act_norm * v_pref; }
    else { out_vx = new_vx; out_vy = new_vy; }
}

// ---- fp64 geometry of the clamp (numpy semantics: np.dot / np.linalg.norm fuse, see oracle/crowd_oracle.c) ----
__device__ __forceinline__ double npdot2(double x0, double x1, double y0, double y1) { return fma(x1, y1, x0 * y0); }
__device__ __forceinline__ double npnorm2(double x, double y) { return sqrt(fma(y, y, x * x)); }

// utils_plus.closest_distance_between_line_segments (utils_plus.py:205-338), z = 0
__device__ void segseg(const double *a0, const double *a1_in, const double *b0, const double *b1_in, double *pA, double *pB, double &dist)
{
    double a1[2] = {a1_in[0], a1_in[1]}, b1[2] = {b1_in[0], b1_in[1]};
    double A[2] = {a1[0] - a0[0], a1[1] - a0[1]}, B[2] = {b1[0] - b0[0], b1[1] - b0[1]};
    const double magA = npnorm2(A[0], A[1]), magB = npnorm2(B[0], B[1]);
    double _A[2], _B[2];
    if (magA < 1e-8) { a1[0] = a0[0]; a1[1] = a0[1]; _A[0] = _A[1] = 0.0; } else { _A[0] = A[0] / magA; _A[1] = A[1] / magA; }
    if (magB < 1e-8) { b1[0] = b0[0]; b1[1] = b0[1]; _B[0] = _B[1] = 0.0; } else { _B[0] = B[0] / magB; _B[1] = B[1] / magB; }
    const double cz = _A[0] * _B[1] - _A[1] * _B[0];
    const double ncross = sqrt(cz * cz);
    const double denom = ncross * ncross;
#define SS_RET(PA, PB) do { pA[0] = (PA)[0]; pA[1] = (PA)[1]; pB[0] = (PB)[0]; pB[1] = (PB)[1]; \
                            dist = npnorm2(pA[0] - pB[0], pA[1] - pB[1]); return; } while (0)
    if (denom == 0.0) {
        const double d0 = npdot2(_A[0], _A[1], b0[0] - a0[0], b0[1] - a0[1]);
        const double d1 = npdot2(_A[0], _A[1], b1[0] - a0[0], b1[1] - a0[1]);
        if (d0 <= 0 && 0 >= d1) {
            if (fabs(d0) < fabs(d1)) SS_RET(a0, b0);
            SS_RET(a0, b1);
        } else if (d0 >= magA && magA <= d1) {
            if (fabs(d0) < fabs(d1)) SS_RET(a1, b0);
            SS_RET(a1, b1);
        } else {
            double a0f[2], _Af[2], qA[2], qB[2];
            if (npnorm2(_A[0] - _B[0], _A[1] - _B[1]) < 1e-8 || magB < 1e-8) { a0f[0] = a0[0]; a0f[1] = a0[1]; _Af[0] = _A[0]; _Af[1] = _A[1]; }
            else { a0f[0] = a1[0]; a0f[1] = a1[1]; _Af[0] = -_A[0]; _Af[1] = -_A[1]; }
            const double d0f = npdot2(_Af[0], _Af[1], b0[0] - a0f[0], b0[1] - a0f[1]);
            if (d0f >= 0) {
                qB[0] = b0[0]; qB[1] = b0[1];
                const double t = npdot2(_Af[0], _Af[1], qB[0] - a0f[0], qB[1] - a0f[1]);
                qA[0] = a0f[0] + _Af[0] * t; qA[1] = a0f[1] + _Af[1] * t;
            } else {
                qA[0] = a0f[0]; qA[1] = a0f[1];
                const double t = npdot2(_B[0], _B[1], qA[0] - b0[0], qA[1] - b0[1]);
                qB[0] = b0[0] + _B[0] * t; qB[1] = b0[1] + _B[1] * t;
            }
            SS_RET(qA, qB);
        }
    }
    const double t[2] = {b0[0] - a0[0], b0[1] - a0[1]};
    const double detA = cz * (t[0] * _B[1] - t[1] * _B[0]);
    const double detB = cz * (t[0] * _A[1] - t[1] * _A[0]);
    const double t0 = detA / denom, t1 = detB / denom;
    double qA[2] = {a0[0] + _A[0] * t0, a0[1] + _A[1] * t0}, qB[2] = {b0[0] + _B[0] * t1, b0[1] + _B[1] * t1};
    if (t0 < 0) { qA[0] = a0[0]; qA[1] = a0[1]; } else if (t0 > magA) { qA[0] = a1[0]; qA[1] = a1[1]; }
    if (t1 < 0) { qB[0] = b0[0]; qB[1] = b0[1]; } else if (t1 > magB) { qB[0] = b1[0]; qB[1] = b1[1]; }
    if (t0 < 0 || t0 > magA) {
        double dot = npdot2(_B[0], _B[1], qA[0] - b0[0], qA[1] - b0[1]);
        if (dot < 0) dot = 0; else if (dot > magB) dot = magB;
        qB[0] = b0[0] + _B[0] * dot; qB[1] = b0[1] + _B[1] * dot;
    }
    if (t1 < 0 || t1 > magB) {
        double dot = npdot2(_A[0], _A[1], qB[0] - a0[0], qB[1] - a0[1]);
        if (dot < 0) dot = 0; else if (dot > magA) dot = magA;
        qA[0] = a0[0] + _A[0] * dot; qA[1] = a0[1] + _A[1] * dot;
    }
    SS_RET(qA, qB);
#undef SS_RET
}

// Agent.compute_position (agent_plus.py:175-185)
__device__ __forceinline__ void compute_position(double px, double py, double theta, int kin, double a0, double a1, double dt,
                                                 double &ox, double &oy)
{
    if (kin == SNB_KIN_HOLONOMIC) { ox = px + a0 * dt; oy = py + a1 * dt; }
    else { const double th = theta + a1; ox = px + cos(th) * a0 * dt; oy = py + sin(th) * a0 * dt; }
}

// CrowdSimPlus.constrain_agent_action_exact (crowd_sim_plus.py:869-989)
__device__ void constrain_action(double px, double py, double theta, double r, double dt, int kin, double a0, double a1,
                                 int n_seg, const double *segs, double &o0, double &o1)
{
    const double PI = 3.14159265358979323846;
    const double cur[2] = {px, py};
    double fut[2];
    compute_position(px, py, theta, kin, a0, a1, dt, fut[0], fut[1]);
    const double mdir[2] = {fut[0] - cur[0], fut[1] - cur[1]};
    const double movement_mag = npnorm2(mdir[0], mdir[1]);
    double f0 = a0, f1 = a1;
    for (int k = 0; k < n_seg; ++k) {
        const double *L = segs + 4 * k;
        double pA[2], pB[2], closest_distance;
        segseg(L, L + 2, cur, fut, pA, pB, closest_distance);
        if (!(closest_distance - r < 0.0)) continue;
        double fin[2];
        if ((npnorm2(pA[0] - L[0], pA[1] - L[1]) < 1e-8 || npnorm2(pA[0] - L[2], pA[1] - L[3]) < 1e-8) &&
            npnorm2(pA[0] - pB[0], pA[1] - pB[1]) > 1e-8) {
            const double dvec[2] = {pB[0] - cur[0], pB[1] - cur[1]};
            const double dir_mag = npnorm2(dvec[0], dvec[1]);
            double _d[2], redux;
            if (dir_mag > 0.0 && npnorm2(pA[0] - cur[0], pA[1] - cur[1]) - r < 1e-4 &&
                npdot2(mdir[0], mdir[1], pA[0] - cur[0], pA[1] - cur[1]) > -1e-8) {
                _d[0] = dvec[0] / dir_mag; _d[1] = dvec[1] / dir_mag; redux = dir_mag;
            } else if (dir_mag > 0.0) {
                _d[0] = dvec[0] / dir_mag; _d[1] = dvec[1] / dir_mag;
                const double av = npdot2(-dvec[0], -dvec[1], pA[0] - pB[0], pA[1] - pB[1]) / (dir_mag * closest_distance);
                const double clipped = av < -1.0 ? -1.0 : (av > 1.0 ? 1.0 : av);
                const double alpha = acos(clipped);
                if (alpha == PI) redux = r - closest_distance;
                else {
                    const double gamma = asin(closest_distance * sin(alpha) / r);
                    const double beta = PI - alpha - gamma;
                    redux = r * sin(beta) / sin(alpha) + 1e-7;
                }
            } else { redux = 0.0; _d[0] = dvec[0]; _d[1] = dvec[1]; }
            const double m = (dir_mag - redux) > 0 ? (dir_mag - redux) : 0;
            fin[0] = cur[0] + _d[0] * m; fin[1] = cur[1] + _d[1] * m;
        } else {
            // closest_point_on_segment_extended (utils_plus.py:44-65)
            double cl[2];
            {
                const double sx = L[2] - L[0], sy = L[3] - L[1];
                if (sx == 0 && sy == 0) { cl[0] = L[0]; cl[1] = L[1]; }
                else { const double u = ((cur[0] - L[0]) * sx + (cur[1] - L[1]) * sy) / (sx * sx + sy * sy); cl[0] = L[0] + u * sx; cl[1] = L[1] + u * sy; }
            }
            if (movement_mag > 0.0 && npnorm2(cl[0] - cur[0], cl[1] - cur[1]) - r < 1e-4 &&
                npdot2(mdir[0], mdir[1], cl[0] - cur[0], cl[1] - cur[1]) > -1e-8) {
                fin[0] = cur[0]; fin[1] = cur[1];
            } else if (movement_mag > 0.0) {
                // intersection_of_vec_line_and_2p_line (utils_plus.py:6-18)
                const double x1 = L[0], y1 = L[1], x2 = L[2], y2 = L[3];
                const double x3 = cur[0], y3 = cur[1], x4 = cur[0] + mdir[0], y4 = cur[1] + mdir[1];
                const double den = (x1 - x2) * (y3 - y4) - (y1 - y2) * (x3 - x4);
                const double ix = ((x1 * y2 - y1 * x2) * (x3 - x4) - (x1 - x2) * (x3 * y4 - y3 * x4)) / den;
                const double iy = ((x1 * y2 - y1 * x2) * (y3 - y4) - (y1 - y2) * (x3 * y4 - y3 * x4)) / den;
                const double dc_0 = sqrt((cur[0] - cl[0]) * (cur[0] - cl[0]) + (cur[1] - cl[1]) * (cur[1] - cl[1]));
                double des = (dc_0 - (r + 1e-7)) / dc_0;
                if (!(des > 0.0)) des = 0.0;
                fin[0] = cur[0] + (ix - cur[0]) * des; fin[1] = cur[1] + (iy - cur[1]) * des;
            } else { fin[0] = cur[0]; fin[1] = cur[1]; }
        }
        if (kin == SNB_KIN_HOLONOMIC) {
            const double v_x = (fin[0] - cur[0]) / dt, v_y = (fin[1] - cur[1]) / dt;
            if ((v_x * v_x + v_y * v_y) < (f0 * f0 + f1 * f1)) { f0 = v_x; f1 = v_y; }
        } else {
            if (a0 > 0) { const double v = npnorm2(fin[0] - cur[0], fin[1] - cur[1]) / dt; if (v < f0) { f0 = v; f1 = a1; } }
            else { const double v = -npnorm2(fin[0] - cur[0], fin[1] - cur[1]) / dt; if (v > f0) { f0 = v; f1 = a1; } }
        }
    }
    o0 = f0; o1 = f1;
}

// Human.get_g_xy (human_plus.py:19-52)
__device__ __forceinline__ void get_g_xy(const SnbDoorCfg &door, double px, double py, double fgx, double fgy, double &gx, double &gy)
{
    if (door.enabled) {
        const double ymin = py < fgy ? py : fgy, ymax = py > fgy ? py : fgy;
        if (ymin < door.door_y_mid_min && ymax > door.door_y_mid_max) {
            const double igx = door.door_x_mid, igy = 0.5 * (door.door_y_min + door.door_y_max);
            const double vec_norm = npnorm2(igx - px, igy - py);
            if (vec_norm <= door.door_width / 2.0) { gx = fgx; gy = fgy; } else { gx = igx; gy = igy; }
            return;
        }
    }
    gx = fgx; gy = fgy;
}

__device__ __forceinline__ double py_mod(double x, double y)
{
    double m = fmod(x, y);
    if (m != 0.0 && ((m < 0) != (y < 0))) m += y;
    return m;
}

// ---- TMA 1-D bulk copy + mbarrier (sm_90+/sm_100a) ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t done = 0;
    do {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

#define CROWD_THREADS 256
static_assert(CROWD_THREADS == CROWD_THREADS_C, "SmemLines stride");

__global__ void __launch_bounds__(CROWD_THREADS, 3) crowd_step_kernel(const CrowdParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int H = P.st.H, E = P.st.E, B = P.st.B;
    const int epc = P.epc;
    const int env0 = blockIdx.x * epc;
    const int nenv = min(epc, B - env0);
    if (nenv <= 0) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int EA = ((epc * H + 1) & ~1); // padded to keep every array 16-byte aligned
    const int EX = ((epc * (E > 0 ? E : 1) + 1) & ~1);

    // carve shared memory
    double *sd = reinterpret_cast<double *>(smem_raw);
    Tile T;
    T.px = sd; sd += EA; T.py = sd; sd += EA; T.vx = sd; sd += EA; T.vy = sd; sd += EA;
    T.rad = sd; sd += EA; T.gx = sd; sd += EA; T.gy = sd; sd += EA; T.vpref = sd; sd += EA;
    T.ex_px = sd; sd += EX; T.ex_py = sd; sd += EX; T.ex_vx = sd; sd += EX; T.ex_vy = sd; sd += EX; T.ex_rad = sd; sd += EX;
    T.act = sd; sd += 2 * EA;
    T.segs = sd; sd += 4 * ((P.n_seg + 1) & ~1);
    double *s_next = sd; sd += 2 * EA;       // constrained human next positions
    uint64_t *bar = reinterpret_cast<uint64_t *>(sd); sd += 2;
    WarpScratch *WS = reinterpret_cast<WarpScratch *>(sd);              // warp mode only
    float4 *s_fast_lines = reinterpret_cast<float4 *>(sd);              // thread mode, fast path only (same bytes: the modes exclude each other)
    int4 *s_lp3 = reinterpret_cast<int4 *>(s_fast_lines + (size_t)P.fast_lines * CROWD_THREADS);   // LP3 work queue, one entry per thread at most
    int *s_lp3_count = reinterpret_cast<int *>(s_lp3 + CROWD_THREADS);
    if (P.fast_lines && tid == 0) *s_lp3_count = 0;

    // ---- stage the tile: 8 human arrays by TMA bulk copy when 16-byte aligned, else by plain loads ----
    const int nA = nenv * H;
    const size_t goff = (size_t)env0 * H;
    const uint32_t bytes = (uint32_t)(nA * sizeof(double));
    const double *src[8] = {P.st.px + goff, P.st.py + goff, P.st.vx + goff, P.st.vy + goff,
                            P.st.radius + goff, P.st.gx + goff, P.st.gy + goff, P.st.vpref + goff};
    double *dst[8] = {T.px, T.py, T.vx, T.vy, T.rad, T.gx, T.gy, T.vpref};
    bool aligned = (bytes % 16u) == 0;
#pragma unroll
    for (int a = 0; a < 8; ++a) aligned = aligned && ((reinterpret_cast<uintptr_t>(src[a]) & 15u) == 0);
    if (aligned) {
        if (tid == 0) {
            mbar_init(bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (tid == 0) {
            mbar_expect_tx(bar, 8u * bytes);
#pragma unroll
            for (int a = 0; a < 8; ++a) tma_bulk_g2s(dst[a], src[a], bytes, bar);
        }
    } else {
        for (int a = 0; a < 8; ++a)
            for (int k = tid; k < nA; k += blockDim.x) dst[a][k] = src[a][k];
    }
    // extras + segments: small, plain (read-only path) loads overlap the bulk copies
    for (int k = tid; k < nenv * E; k += blockDim.x) {
        const size_t g = (size_t)env0 * E + k;
        T.ex_px[k] = P.st.ex_px[g]; T.ex_py[k] = P.st.ex_py[g]; T.ex_vx[k] = P.st.ex_vx[g]; T.ex_vy[k] = P.st.ex_vy[g];
        T.ex_rad[k] = P.st.ex_radius[g];
    }
    for (int k = tid; k < 4 * P.n_seg; k += blockDim.x) T.segs[k] = P.segs[k];
    if (aligned) mbar_wait(bar, 0);
    __syncthreads();

    if (P.log) {                               // CrowdSimPlus.step's states.append: the positions this step starts from
        for (int k = tid; k < nA; k += blockDim.x) {
            const int e = k / H, h = k - e * H;
            double *o = P.log + (((size_t)(env0 + e) * P.log_L + P.log_slot) * (H + 1) + h) * 2;
            o[0] = T.px[k]; o[1] = T.py[k];
        }
        for (int k = tid; k < nenv; k += blockDim.x) {
            double *o = P.log + (((size_t)(env0 + k) * P.log_L + P.log_slot) * (H + 1) + H) * 2;
            o[0] = T.ex_px[k * E]; o[1] = T.ex_py[k * E];
        }
    }

    // ---- phase 1 (large batches): one thread per human ----
    if (P.thread_mode) {
        for (int task = tid; task < nA; task += blockDim.x) {
            const int e = task / H, i = task - e * H;
            const int genv = env0 + e;
            if (P.active && !P.active[genv]) continue;
            double ax, ay;
            if (P.cfg.policy == SNB_POLICY_SFM) {
                sfm_predict_thread(P, T, e, i, ax, ay);
                if (P.nbr_cnt) P.nbr_cnt[genv * H + i] = 0;
            } else {
                float fx, fy;
                if (P.fast_lines) {
                    int n_l; float spd;
                    const int fail = orca_predict_thread_fast<CROWD_THREADS_C>(P, T, s_fast_lines + threadIdx.x, e, i, genv, fx, fy, n_l, spd);
                    if (fail >= 0) {          // queue this human for the compacted LP3 pass below
                        const int slot = atomicAdd(s_lp3_count, 1);
                        s_lp3[slot] = make_int4(threadIdx.x | (task << 16), n_l | (fail << 8), __float_as_int(spd), 0);
                    }
                } else orca_predict_thread(P, T, e, i, genv, T.gx[e * H + i], T.gy[e * H + i], T.vpref[e * H + i], fx, fy);
                ax = (double)fx; ay = (double)fy;
            }
            T.act[2 * task] = ax; T.act[2 * task + 1] = ay;
            if (!P.full_step && P.out_v) { P.out_v[2 * (goff + task)] = ax; P.out_v[2 * (goff + task) + 1] = ay; }
        }
        if (P.fast_lines) {
            // ---- linearProgram3 for the humans whose LP2 failed, compacted onto the first threads of the CTA ----
            __syncthreads();
            const int n_q = *s_lp3_count;
            for (int q = tid; q < n_q; q += blockDim.x) {
                const int4 it = s_lp3[q];
                const int owner = it.x & 0xffff, task = it.x >> 16, n_l = it.y & 0xff, fail = it.y >> 8;
                float rx = (float)T.act[2 * task], ry = (float)T.act[2 * task + 1];     // the LP2 result (exactly representable: it was a float)
                Line proj[FAST_LCAP];
                lp3_serial(SmemLines{s_fast_lines + owner}, n_l, 0, fail, __int_as_float(it.z), rx, ry, proj);
                T.act[2 * task] = (double)rx; T.act[2 * task + 1] = (double)ry;
                if (!P.full_step && P.out_v) { P.out_v[2 * (goff + task)] = (double)rx; P.out_v[2 * (goff + task) + 1] = (double)ry; }
            }
        }
    } else
    // ---- phase 1: one warp per human ----
    for (int task = warp; task < nA; task += nwarps) {
        const int e = task / H, i = task - e * H;
        const int genv = env0 + e;
        if (P.active && !P.active[genv]) continue;
        double ax, ay;
        if (P.cfg.policy == SNB_POLICY_SFM) {
            sfm_predict_warp(P, T, e, i, lane, ax, ay);
            if (P.nbr_cnt && lane == 0) P.nbr_cnt[genv * H + i] = 0;
        } else {
            float fx, fy;
            orca_predict_warp(P, T, WS[warp], e, i, lane, genv, fx, fy);
            ax = (double)fx; ay = (double)fy;
        }
        if (lane == 0) {
            T.act[2 * task] = ax; T.act[2 * task + 1] = ay;
            if (!P.full_step && P.out_v) { P.out_v[2 * (goff + task)] = ax; P.out_v[2 * (goff + task) + 1] = ay; }
        }
    }
    if (!P.full_step) return;
    __syncthreads();

    // ---- phase 2a: one thread per human: static-obstacle clamp, next position ----
    const double dt = P.cfg.time_step;
    for (int task = tid; task < nA; task += blockDim.x) {
        const int e = task / H;
        const int genv = env0 + e;
        if (P.active && !P.active[genv]) continue;
        double c0 = T.act[2 * task], c1 = T.act[2 * task + 1];
        if (P.n_seg > 0) constrain_action(T.px[task], T.py[task], 0.0, T.rad[task], dt, SNB_KIN_HOLONOMIC, c0, c1, P.n_seg, T.segs, c0, c1);
        T.act[2 * task] = c0; T.act[2 * task + 1] = c1;
        s_next[2 * task] = T.px[task] + c0 * dt; s_next[2 * task + 1] = T.py[task] + c1 * dt;
    }
    __syncthreads();

    // ---- phase 2b: one thread per (environment, candidate robot action): robot clamp, collision scan, reward;
    //      update=True (one action): robot update and clocks; update=False: outputs only, the state is not touched ----
    const int nact = P.full_step == 2 ? P.n_actions : 1;
    for (int ea = tid; ea < nenv * nact; ea += blockDim.x) {
        const int e = ea / nact, a = ea - e * nact;
        const int genv = env0 + e;
        const size_t oidx = (size_t)genv * nact + a;
        if (P.active && !P.active[genv]) {
            // a frozen environment reports nothing: its terminal reward / flags were returned by the step that finished it
            if (P.full_step == 1) { if (P.reward) P.reward[oidx] = 0.0; if (P.flags) P.flags[oidx] = 0; }
            continue;
        }
        const double rpx = T.ex_px[e * E], rpy = T.ex_py[e * E], rrad = T.ex_rad[e * E];
        const double rtheta = P.st.rtheta[genv];
        const double ra0 = P.robot_action[2 * oidx], ra1 = P.robot_action[2 * oidx + 1];
        const int kin = P.st.robot_kinematics;
        double c0 = ra0, c1 = ra1;
        if (P.n_seg > 0) constrain_action(rpx, rpy, rtheta, rrad, dt, kin, ra0, ra1, P.n_seg, T.segs, c0, c1);
        const bool stat_collision = (ra0 != c0); // quirk q11: only the first component is compared
        double rnx, rny;
        compute_position(rpx, rpy, rtheta, kin, c0, c1, dt, rnx, rny);
        double dmin = INFINITY;
        bool collision = false;
        for (int i = 0; i < H; ++i) {
            const int k = e * H + i;
            const double closest = npnorm2(rnx - s_next[2 * k], rny - s_next[2 * k + 1]);
            if (closest < (rrad + T.rad[k])) { collision = true; break; }
            else if (closest < dmin) dmin = closest;
        }
        bool frozen;
        if (kin == SNB_KIN_HOLONOMIC) frozen = sqrt(c0 * c0 + c1 * c1) * dt < 0.01;
        else frozen = fabs(c0 * dt) < 0.01;
        const double rgx = P.st.rgx[genv], rgy = P.st.rgy[genv];
        const bool reached = npnorm2(rnx - rgx, rny - rgy) < rrad;
        const double curr_dist = npnorm2(rgx - rnx, rgy - rny);
        const double gt = P.st.global_time[genv];
        double rew = 0.0;
        int f = 0;
        if (reached) { rew += P.rcfg.success_reward; f |= SNB_F_REACHED | SNB_F_DONE; }
        else if (gt >= P.rcfg.time_limit) { rew += P.rcfg.timeout; f |= SNB_F_TIMEOUT | SNB_F_DONE; }
        if (collision) { rew += P.rcfg.collision_penalty; f |= SNB_F_COLLISION; }
        if (stat_collision) { rew += P.rcfg.wall_collision_penalty; f |= SNB_F_WALL; }
        if (P.rcfg.discomfort && dmin < P.rcfg.discomfort_dist) {
            rew += (dmin - P.rcfg.discomfort_dist) * P.rcfg.discomfort_penalty_factor * dt;
            f |= SNB_F_DANGER;
        }
        if (P.rcfg.has_progress) {
            rew += (P.st.prev_dist[genv] - curr_dist) * P.rcfg.progress_factor;
            if (P.full_step == 1) P.st.prev_dist[genv] = curr_dist;   // crowd_sim_plus.py:1136-1137: only when update
        }
        if (frozen) { rew += P.rcfg.freezing_penalty; f |= SNB_F_FROZEN; }
        if (P.reward) P.reward[oidx] = rew;
        if (P.dmin) P.dmin[oidx] = dmin;
        if (P.flags) P.flags[oidx] = f;
        if (P.full_step == 2) {
            if (P.next_robot) { P.next_robot[2 * oidx] = rnx; P.next_robot[2 * oidx + 1] = rny; }
            continue;
        }
        // Agent.step for the robot (agent_plus.py:199-214)
        const size_t gx = (size_t)genv * E;
        P.st.ex_px[gx] = rnx; P.st.ex_py[gx] = rny;
        if (kin == SNB_KIN_HOLONOMIC) {
            P.st.ex_vx[gx] = c0; P.st.ex_vy[gx] = c1;
            P.st.rtheta[genv] = atan2(c1, c0);
        } else {
            const double PI = 3.14159265358979323846;
            const double un = py_mod(rtheta + c1, 2 * PI);
            const double th = un > PI ? un - 2 * PI : un;
            P.st.rtheta[genv] = th;
            P.st.ex_vx[gx] = c0 * cos(th); P.st.ex_vy[gx] = c0 * sin(th);
        }
        P.st.global_time[genv] = gt + dt;
    }

    // ---- phase 2c: one thread per human: integrate, goal switch, arrival time; coalesced write-back ----
    for (int task = tid; task < nA; task += blockDim.x) {
        const int e = task / H;
        const int genv = env0 + e;
        if (P.active && !P.active[genv]) continue;
        const size_t g = goff + task;
        const double c0 = T.act[2 * task], c1 = T.act[2 * task + 1];
        const double nx = s_next[2 * task], ny = s_next[2 * task + 1];
        if (P.full_step == 2) {   // Agent.get_next_observable_state (agent_plus.py:86-98), holonomic humans
            if (P.next_h) { P.next_h[4 * g] = nx; P.next_h[4 * g + 1] = ny; P.next_h[4 * g + 2] = c0; P.next_h[4 * g + 3] = c1; }
            continue;
        }
        P.st.px[g] = nx; P.st.py[g] = ny; P.st.vx[g] = c0; P.st.vy[g] = c1;
        P.st.theta[g] = atan2(c1, c0);
        if (P.door.enabled) {     // without a door Human.set_g_xy returns the final goal, which gx / gy already hold: nothing to write
            double ngx, ngy;
            get_g_xy(P.door, nx, ny, P.st.fgx[g], P.st.fgy[g], ngx, ngy);
            P.st.gx[g] = ngx; P.st.gy[g] = ngy;
        }
    }
    if (P.full_step == 2) return;
    __syncthreads(); // phase 2b's global_time stores are visible to the CTA after this barrier
    for (int task = tid; task < nA; task += blockDim.x) {
        const int e = task / H;
        const int genv = env0 + e;
        if (P.active && !P.active[genv]) continue;
        const size_t g = goff + task;
        if (P.st.human_time[g] == 0 &&
            npnorm2(P.st.px[g] - P.st.gx[g], P.st.py[g] - P.st.gy[g]) < T.rad[task])
            P.st.human_time[g] = P.st.global_time[genv];
    }
}

// ---------------------------------------------------------------------------------------------------------
// crowd_orca_warp_kernel: the large-batch kernel of the simulator's common shape -- plain ORCA humans, no wall segments, one
// robot -- with NO CTA-wide barrier: every warp owns floor(32 / H) whole environments (lane = human), keeps their state in its own
// slice of shared memory and runs all phases of the step on it, synchronising with __syncwarp only.
//
// Why (ncu of the one-thread-per-human CTA kernel at 2^18 and 2^20 environments, profiles/r02_ncu_crowd_*): 39 % of the warp stall
// samples sat on the __syncthreads between phase 1 and phase 2 -- seven warps waiting for the one whose lanes drew the longest
// linear programs -- and issue slots were 40 % busy.  Here a warp that finishes early simply moves on to its own phase 2 and exits,
// and a new CTA's warps take its place.  linearProgram3 (needed by ~10 % of the humans, but by ~95 % of the warps) is not run by the
// lanes in lock-step: each human that needs it is handed to the whole warp (lp3_warp, lanes = half-planes, the shuffle-reduction LP
// of the small-launch path).  Same float operations per human as everywhere else => bit-identical results.
// ---------------------------------------------------------------------------------------------------------
constexpr int FW_WARPS = 4;
// per-warp shared memory: half-planes [line_cap][32] float4 (the next positions of phase 2 reuse their first 512 bytes), 32 float4 of
// LP3 scratch, the float copy of the warp's humans and robots (fs, fr), and the fp64 state the step itself needs (5 x 32 humans,
// 5 x robots).  7.6 KB per warp at H = 10, max_neighbors = 10 -> 7 CTAs of 4 warps per SM.
static __host__ __device__ inline int fw_exn(int H, int E) { return ((32 / H) * E + 1) & ~1; }
static __host__ __device__ inline size_t fw_warp_bytes(int line_cap, int H, int E)
{
    const int fsn = 32 + fw_exn(H, E);
    return (size_t)line_cap * 32 * sizeof(float4) + 32 * sizeof(float4) + (size_t)fsn * sizeof(float4) + (((size_t)fsn * sizeof(float) + 15) & ~(size_t)15) +
           (size_t)(5 * 32 + 5 * fw_exn(H, E)) * sizeof(double);
}

__global__ void __launch_bounds__(32 * FW_WARPS) crowd_orca_warp_kernel(const CrowdParams P)
{
    extern __shared__ __align__(16) unsigned char fw_smem[];
    const int H = P.st.H, E = P.st.E;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int epw = 32 / H;
    const int env0 = (blockIdx.x * FW_WARPS + w) * epw;
    const int nenv = min(epw, P.st.B - env0);
    if (nenv <= 0) return;                                  // the whole warp leaves together; nobody waits for it
    const int exn = fw_exn(H, E), fsn = 32 + exn;
    unsigned char *wb = fw_smem + (size_t)w * fw_warp_bytes(P.fast_lines, H, E);
    float4 *s_lines = reinterpret_cast<float4 *>(wb); wb += (size_t)P.fast_lines * 32 * sizeof(float4);
    float4 *scratch = reinterpret_cast<float4 *>(wb); wb += 32 * sizeof(float4);
    float4 *fs = reinterpret_cast<float4 *>(wb); wb += (size_t)fsn * sizeof(float4);
    float *fr = reinterpret_cast<float *>(wb); wb += ((size_t)fsn * sizeof(float) + 15) & ~(size_t)15;
    double *sd = reinterpret_cast<double *>(wb);
    Tile T;
    T.px = sd; sd += 32; T.py = sd; sd += 32; T.vx = sd; sd += 32; T.vy = sd; sd += 32; T.rad = sd; sd += 32;
    T.ex_px = sd; sd += exn; T.ex_py = sd; sd += exn; T.ex_vx = sd; sd += exn; T.ex_vy = sd; sd += exn; T.ex_rad = sd; sd += exn;
    T.gx = T.gy = T.vpref = nullptr;                        // goal and v_pref stay in their owner's registers
    T.act = nullptr; T.segs = nullptr;
    double *s_next = reinterpret_cast<double *>(s_lines);   // phase 2 only: the half-planes are dead by then

    const int nA = nenv * H;
    const bool live = lane < nA;
    const int e = live ? lane / H : 0, i = lane - e * H;
    const int genv = env0 + e;
    const size_t g = (size_t)env0 * H + lane;
    const bool on = live && !(P.active && !P.active[genv]);
    const double rad_pad = 0.01 + P.cfg.safety_space;
    double my_px = 0.0, my_py = 0.0, my_gx = 0.0, my_gy = 0.0, my_vpref = 0.0, my_rad = 0.0;
    if (live) {
        my_px = P.st.px[g]; my_py = P.st.py[g];
        const double vx = P.st.vx[g], vy = P.st.vy[g];
        my_rad = P.st.radius[g]; my_gx = P.st.gx[g]; my_gy = P.st.gy[g]; my_vpref = P.st.vpref[g];
        T.px[lane] = my_px; T.py[lane] = my_py; T.vx[lane] = vx; T.vy[lane] = vy; T.rad[lane] = my_rad;
        fs[lane] = make_float4((float)my_px, (float)my_py, (float)vx, (float)vy);
        fr[lane] = (float)(my_rad + rad_pad);
        if (P.log) {
            double *o = P.log + (((size_t)genv * P.log_L + P.log_slot) * (H + 1) + i) * 2;
            o[0] = my_px; o[1] = my_py;
        }
    }
    if (lane < nenv * E) {
        const size_t x = (size_t)env0 * E + lane;
        const double a = P.st.ex_px[x], b = P.st.ex_py[x], c = P.st.ex_vx[x], d = P.st.ex_vy[x], r = P.st.ex_radius[x];
        T.ex_px[lane] = a; T.ex_py[lane] = b; T.ex_vx[lane] = c; T.ex_vy[lane] = d; T.ex_rad[lane] = r;
        fs[32 + lane] = make_float4((float)a, (float)b, (float)c, (float)d);
        fr[32 + lane] = (float)(r + rad_pad);
        if (P.log) {                                        // E == 1 in this kernel: lane = environment, the extra is the robot
            double *o = P.log + (((size_t)(env0 + lane) * P.log_L + P.log_slot) * (H + 1) + H) * 2;
            o[0] = a; o[1] = b;
        }
    }
    __syncwarp();

    // ---- phase 1: ORCA per lane up to linearProgram2; linearProgram3 warp-cooperatively, one human at a time ----
    float fx = 0.0f, fy = 0.0f, spd = 0.0f;
    int fail = -1, n_l = 0;
    if (on) {
        fail = orca_predict_wo(P, fs, fr, lane, e * H, 32 + e * E, i, genv, my_gx - my_px, my_gy - my_py, my_vpref, s_lines + lane, fx, fy, n_l, spd);
        if (fail == -2) { orca_predict_thread(P, T, e, i, genv, my_gx, my_gy, my_vpref, fx, fy); fail = -1; }
    }
    unsigned need = __ballot_sync(FULL, fail >= 0);
    while (need) {
        const int src = __ffs(need) - 1;
        need &= need - 1;
        const int n = __shfl_sync(FULL, n_l, src), begin = __shfl_sync(FULL, fail, src);
        const float r = __shfl_sync(FULL, spd, src);
        float rx = __shfl_sync(FULL, fx, src), ry = __shfl_sync(FULL, fy, src);
        if (P.lp3_mode == 0) {
            Line my; my.px = 0.f; my.py = 0.f; my.dx = 1.f; my.dy = 0.f;
            if (lane < n) my = SmemLinesT<32>{s_lines + src}[lane];
            lp3_warp(my, n, 0, begin, r, rx, ry, lane, reinterpret_cast<Line *>(scratch));
        } else {
            lp3_warp_skip(s_lines + src, n, begin, r, rx, ry, lane, scratch);
        }
        if (lane == src) { fx = rx; fy = ry; }
    }
    const double c0 = (double)fx, c1 = (double)fy;
    if (!P.full_step) {
        if (on && P.out_v) { P.out_v[2 * g] = c0; P.out_v[2 * g + 1] = c1; }
        return;
    }

    // ---- phase 2a: no wall segments -> the action is not clamped; next position (Agent.compute_position) ----
    const double dt = P.cfg.time_step;
    double nx = 0.0, ny = 0.0;
    __syncwarp();                                           // every lane is done with the half-planes s_next overlays
    if (on) { nx = my_px + c0 * dt; ny = my_py + c1 * dt; s_next[2 * lane] = nx; s_next[2 * lane + 1] = ny; }
    __syncwarp();

    // ---- phase 2b: lane e < nenv is the robot of environment e (crowd_sim_plus.py:1058-1172, same expressions as crowd_step_kernel) ----
    double gt_new = 0.0;
    if (lane < nenv) {
        const int re = lane, renv = env0 + re;
        if (P.active && !P.active[renv]) {
            if (P.reward) P.reward[renv] = 0.0;
            if (P.flags) P.flags[renv] = 0;
        } else {
            const double rpx = T.ex_px[re * E], rpy = T.ex_py[re * E], rrad = T.ex_rad[re * E];
            const double rtheta = P.st.rtheta[renv];
            const double ra0 = P.robot_action[2 * (size_t)renv], ra1 = P.robot_action[2 * (size_t)renv + 1];
            const int kin = P.st.robot_kinematics;
            const double a0 = ra0, a1 = ra1;                      // no segments: constrain_agent_action_exact returns the action
            double rnx, rny;
            compute_position(rpx, rpy, rtheta, kin, a0, a1, dt, rnx, rny);
            double dmin = INFINITY;
            bool collision = false;
            for (int h = 0; h < H; ++h) {
                const int k = re * H + h;
                const double closest = npnorm2(rnx - s_next[2 * k], rny - s_next[2 * k + 1]);
                if (closest < (rrad + T.rad[k])) { collision = true; break; }
                else if (closest < dmin) dmin = closest;
            }
            bool frozen;
            if (kin == SNB_KIN_HOLONOMIC) frozen = sqrt(a0 * a0 + a1 * a1) * dt < 0.01;
            else frozen = fabs(a0 * dt) < 0.01;
            const double rgx = P.st.rgx[renv], rgy = P.st.rgy[renv];
            const bool reached = npnorm2(rnx - rgx, rny - rgy) < rrad;
            const double curr_dist = npnorm2(rgx - rnx, rgy - rny);
            const double gt = P.st.global_time[renv];
            double rew = 0.0;
            int f = 0;
            if (reached) { rew += P.rcfg.success_reward; f |= SNB_F_REACHED | SNB_F_DONE; }
            else if (gt >= P.rcfg.time_limit) { rew += P.rcfg.timeout; f |= SNB_F_TIMEOUT | SNB_F_DONE; }
            if (collision) { rew += P.rcfg.collision_penalty; f |= SNB_F_COLLISION; }
            if (P.rcfg.discomfort && dmin < P.rcfg.discomfort_dist) {
                rew += (dmin - P.rcfg.discomfort_dist) * P.rcfg.discomfort_penalty_factor * dt;
                f |= SNB_F_DANGER;
            }
            if (P.rcfg.has_progress) {
                rew += (P.st.prev_dist[renv] - curr_dist) * P.rcfg.progress_factor;
                P.st.prev_dist[renv] = curr_dist;
            }
            if (frozen) { rew += P.rcfg.freezing_penalty; f |= SNB_F_FROZEN; }
            if (P.reward) P.reward[renv] = rew;
            if (P.dmin) P.dmin[renv] = dmin;
            if (P.flags) P.flags[renv] = f;
            const size_t gx = (size_t)renv * E;
            P.st.ex_px[gx] = rnx; P.st.ex_py[gx] = rny;
            if (kin == SNB_KIN_HOLONOMIC) {
                P.st.ex_vx[gx] = a0; P.st.ex_vy[gx] = a1;
                P.st.rtheta[renv] = atan2(a1, a0);
            } else {
                const double PI = 3.14159265358979323846;
                const double un = py_mod(rtheta + a1, 2 * PI);
                const double th = un > PI ? un - 2 * PI : un;
                P.st.rtheta[renv] = th;
                P.st.ex_vx[gx] = a0 * cos(th); P.st.ex_vy[gx] = a0 * sin(th);
            }
            gt_new = gt + dt;
            P.st.global_time[renv] = gt_new;
        }
    }
    gt_new = __shfl_sync(FULL, gt_new, e);                  // the clock of my environment after this step (human arrival times)

    // ---- phase 2c: Human.step (integrate, heading, door goal) and arrival time; coalesced write-back straight from registers ----
    if (on) {
        P.st.px[g] = nx; P.st.py[g] = ny; P.st.vx[g] = c0; P.st.vy[g] = c1;
        P.st.theta[g] = atan2(c1, c0);
        double ggx = my_gx, ggy = my_gy;
        if (P.door.enabled) {
            get_g_xy(P.door, nx, ny, P.st.fgx[g], P.st.fgy[g], ggx, ggy);
            P.st.gx[g] = ggx; P.st.gy[g] = ggy;
        }
        if (P.st.human_time[g] == 0 && npnorm2(nx - ggx, ny - ggy) < my_rad) P.st.human_time[g] = gt_new;
    }
}

// ---------------------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------------------
/* ORCA phase 1 switches from one warp per human to one thread per human at this many humans per launch (B200, H = 10, r01:
 * 10 240 humans: warp 53 us vs thread 84 us; 40 960: 146 vs 60 us; 2.6 M: 7.07 vs 1.59 ms) */
#define SNB_THREAD_MODE_MIN_AGENTS 24576

static int choose_epc(int H)
{
    int epc = 48 / (H > 0 ? H : 1);
    if (epc < 1) epc = 1;
    if (epc > 16) epc = 16;
    if ((epc * H) & 1) epc += (epc > 1 ? -1 : 1); // keep epc*H even so that every tile offset is 16-byte aligned
    if ((epc * H) & 1) epc = 2;
    return epc;
}

static size_t crowd_smem_bytes(int epc, int H, int E, int n_seg, int fast_lines = 0)
{
    const size_t EA = (size_t)((epc * H + 1) & ~1);
    const size_t EX = (size_t)((epc * (E > 0 ? E : 1) + 1) & ~1);
    size_t d = 8 * EA + 5 * EX + 2 * EA + 4 * (size_t)((n_seg + 1) & ~1) + 2 * EA + 2;
    const size_t scratch = fast_lines ? (size_t)(fast_lines + 1) * CROWD_THREADS * sizeof(float4) + 16 : sizeof(WarpScratch) * (CROWD_THREADS / 32);
    return d * sizeof(double) + scratch + 16;
}

static int launch_crowd(const SnbPolicyCfg *cfg, const SnbDoorCfg *door, const SnbRewardCfg *rcfg, const SnbCrowdState *st,
                        const SnbObstacles *obs, const double *robot_action, const uint8_t *active, double *reward,
                        double *dmin, int *flags, double *out_v, int *nbr, int *nbr_cnt, int *status, int full_step, void *stream,
                        int n_actions = 1, double *next_h = nullptr, double *next_robot = nullptr, double *log = nullptr, int log_L = 0,
                        int log_slot = 0)
{
    SNB_REQUIRE(cfg && st, SNB_EINVAL, "crowd step: cfg/state is NULL");
    SNB_REQUIRE(st->B >= 0 && st->H >= 1 && st->E >= 0, SNB_EINVAL, "crowd step: bad sizes B=%d H=%d E=%d", st->B, st->H, st->E);
    SNB_REQUIRE(st->n_obs_extras >= 0 && st->n_obs_extras <= st->E, SNB_EINVAL, "crowd step: n_obs_extras=%d > E=%d", st->n_obs_extras, st->E);
    SNB_REQUIRE(st->H - 1 + st->n_obs_extras <= SNB_MAX_AGENTS_PER_ENV, SNB_EUNSUPPORTED,
                "crowd step: %d observed agents per human exceed SNB_MAX_AGENTS_PER_ENV=%d", st->H - 1 + st->n_obs_extras, SNB_MAX_AGENTS_PER_ENV);
    SNB_REQUIRE(cfg->policy >= SNB_POLICY_ORCA && cfg->policy <= SNB_POLICY_SFM, SNB_EINVAL, "crowd step: unknown policy %d", cfg->policy);
    SNB_REQUIRE(cfg->max_neighbors >= 0 && cfg->max_neighbors <= SNB_MAX_AGENTS_PER_ENV, SNB_EUNSUPPORTED, "crowd step: max_neighbors=%d unsupported", cfg->max_neighbors);
    SNB_REQUIRE(st->px && st->py && st->vx && st->vy && st->gx && st->gy && st->vpref && st->radius, SNB_EINVAL, "crowd step: NULL human array");
    SNB_REQUIRE(st->E == 0 || (st->ex_px && st->ex_py && st->ex_vx && st->ex_vy && st->ex_radius), SNB_EINVAL, "crowd step: NULL extras array");
    if (full_step) {
        SNB_REQUIRE(door && rcfg && robot_action, SNB_EINVAL, "env step: door/reward/robot_action is NULL");
        SNB_REQUIRE(st->E >= 1, SNB_EINVAL, "env step: the robot must be extra 0 (E >= 1)");
        SNB_REQUIRE(st->theta && st->fgx && st->fgy && st->human_time && st->rtheta && st->rgx && st->rgy && st->global_time,
                    SNB_EINVAL, "env step: NULL state array");
        SNB_REQUIRE(!rcfg->has_progress || st->prev_dist, SNB_EINVAL, "env step: prev_dist is NULL");
    } else {
        SNB_REQUIRE(out_v, SNB_EINVAL, "policy step: out_v is NULL");
    }
    if (st->B == 0) return SNB_OK;

    CrowdParams P;
    memset(&P, 0, sizeof(P));
    P.cfg = *cfg;
    if (door) P.door = *door;
    if (rcfg) P.rcfg = *rcfg;
    P.st = *st;
    P.robot_action = robot_action; P.active = active; P.reward = reward; P.dmin = dmin; P.flags = flags;
    P.out_v = out_v; P.nbr = nbr; P.nbr_cnt = nbr_cnt; P.status = status;
    if (obs && obs->n_seg > 0) {
        P.n_seg = obs->n_seg; P.segs = obs->d_segs;
        P.n_vert = (int)obs->verts.size(); P.verts = obs->d_verts; P.nodes = obs->d_nodes; P.bsp_root = obs->root;
    } else { P.bsp_root = -1; }
    // one thread per human when the launch has enough humans to fill the GPU that way (SNB_CROWD_MODE=warp|thread overrides)
    {
        const char *m = getenv("SNB_CROWD_MODE");
        if (m && m[0] == 't') P.thread_mode = 1;
        else if (m && m[0] == 'w') P.thread_mode = 0;
        else P.thread_mode = (long long)st->B * st->H >= SNB_THREAD_MODE_MIN_AGENTS;
    }
    P.epc = choose_epc(st->H);
    if (P.thread_mode) {                      // a CTA's 256 threads want ~256 humans: more environments per CTA
        int epc = CROWD_THREADS / st->H;
        if (epc < 1) epc = 1;
        if (epc > 64) epc = 64;
        if ((epc * st->H) & 1) epc += (epc > 1 ? -1 : 1);
        if (((epc * st->H) & 1) == 0 && crowd_smem_bytes(epc, st->H, st->E, P.n_seg) <= 96 * 1024) P.epc = epc;
    }
    P.full_step = full_step;
    P.n_actions = n_actions; P.next_h = next_h; P.next_robot = next_robot;
    P.log = log; P.log_L = log_L; P.log_slot = log_slot;
    if (P.thread_mode && P.n_vert == 0 && cfg->policy != SNB_POLICY_SFM && st->H - 1 + st->n_obs_extras <= FAST_MAXO &&
        cfg->max_neighbors >= 1 && cfg->max_neighbors <= FAST_LCAP && !getenv("SNB_CROWD_NO_FAST") &&
        crowd_smem_bytes(P.epc, st->H, st->E, P.n_seg, cfg->max_neighbors) <= 96 * 1024)
        P.fast_lines = cfg->max_neighbors;
    {
        const char *m = getenv("SNB_CROWD_LP3");
        P.lp3_mode = m ? atoi(m) : 1;
    }

    // the barrier-free kernel: plain ORCA, no walls, the simulator's single robot, update / policy-only (not the what-if look-ahead)
    if (P.fast_lines && P.n_seg == 0 && st->E == 1 && full_step != 2 && st->H <= 32 && !getenv("SNB_CROWD_NO_WARPOWN")) {
        const int epw = 32 / st->H;
        const size_t wsm = fw_warp_bytes(P.fast_lines, st->H, st->E) * FW_WARPS;
        static std::once_flag once_w;
        static cudaError_t attr_w = cudaSuccess;
        std::call_once(once_w, [] { attr_w = cudaFuncSetAttribute(crowd_orca_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024); });
        SNB_CUDA_TRY(attr_w);
        const int envs_per_cta = epw * FW_WARPS;
        crowd_orca_warp_kernel<<<(st->B + envs_per_cta - 1) / envs_per_cta, 32 * FW_WARPS, wsm, (cudaStream_t)stream>>>(P);
        snb_count_launch();
        SNB_CUDA_TRY(cudaGetLastError());
        return SNB_OK;
    }

    const size_t smem = crowd_smem_bytes(P.epc, st->H, st->E, P.n_seg, P.fast_lines);
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] { attr_err = cudaFuncSetAttribute(crowd_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024); });
    SNB_CUDA_TRY(attr_err);
    SNB_REQUIRE(smem <= 96 * 1024, SNB_EUNSUPPORTED, "crowd step: tile needs %zu B of shared memory", smem);
    const int grid = (st->B + P.epc - 1) / P.epc;
    crowd_step_kernel<<<grid, CROWD_THREADS, smem, (cudaStream_t)stream>>>(P);
    snb_count_launch();
    SNB_CUDA_TRY(cudaGetLastError());
    return SNB_OK;
}

extern "C" int snb_policy_step(const SnbPolicyCfg *cfg, const SnbCrowdState *state, const SnbObstacles *obs, double *out_v_dev,
                               int32_t *nbr_dev, int32_t *nbr_cnt_dev, int32_t *status_dev, void *stream)
{
    return launch_crowd(cfg, nullptr, nullptr, state, obs, nullptr, nullptr, nullptr, nullptr, nullptr, out_v_dev, nbr_dev,
                        nbr_cnt_dev, status_dev, 0, stream);
}

extern "C" int snb_env_step_logged(const SnbPolicyCfg *cfg, const SnbDoorCfg *door, const SnbRewardCfg *reward_cfg,
                                   const SnbCrowdState *state, const SnbObstacles *obs, const double *robot_action_dev,
                                   const uint8_t *active_dev, double *reward_dev, double *dmin_dev, int32_t *flags_dev,
                                   int32_t *nbr_dev, int32_t *nbr_cnt_dev, int32_t *status_dev, double *log_dev, int32_t L, int32_t slot,
                                   void *stream)
{
    SNB_REQUIRE(log_dev && L >= 1 && slot >= 0 && slot < L, SNB_EINVAL, "snb_env_step_logged: bad log / ring slot");
    SNB_REQUIRE(state && state->E == 1, SNB_EINVAL, "snb_env_step_logged: the simulator state has exactly one extra (the robot)");
    return launch_crowd(cfg, door, reward_cfg, state, obs, robot_action_dev, active_dev, reward_dev, dmin_dev, flags_dev,
                        nullptr, nbr_dev, nbr_cnt_dev, status_dev, 1, stream, 1, nullptr, nullptr, log_dev, L, slot);
}

extern "C" int snb_env_step(const SnbPolicyCfg *cfg, const SnbDoorCfg *door, const SnbRewardCfg *reward_cfg,
                            const SnbCrowdState *state, const SnbObstacles *obs, const double *robot_action_dev,
                            const uint8_t *active_dev, double *reward_dev, double *dmin_dev, int32_t *flags_dev,
                            int32_t *nbr_dev, int32_t *nbr_cnt_dev, int32_t *status_dev, void *stream)
{
    return launch_crowd(cfg, door, reward_cfg, state, obs, robot_action_dev, active_dev, reward_dev, dmin_dev, flags_dev,
                        nullptr, nbr_dev, nbr_cnt_dev, status_dev, 1, stream);
}

extern "C" int snb_env_whatif(const SnbPolicyCfg *cfg, const SnbDoorCfg *door, const SnbRewardCfg *reward_cfg,
                              const SnbCrowdState *state, const SnbObstacles *obs, const double *robot_actions_dev,
                              int32_t n_actions, const uint8_t *active_dev, double *reward_dev, double *dmin_dev,
                              int32_t *flags_dev, double *next_humans_dev, double *next_robot_dev, int32_t *status_dev, void *stream)
{
    SNB_REQUIRE(n_actions >= 1, SNB_EINVAL, "env what-if: n_actions=%d", n_actions);
    return launch_crowd(cfg, door, reward_cfg, state, obs, robot_actions_dev, active_dev, reward_dev, dmin_dev, flags_dev,
                        nullptr, nullptr, nullptr, status_dev, 2, stream, n_actions, next_humans_dev, next_robot_dev);
}

// ---- host-buffer single call: what policy.predict(state) does behind the rvo2 FFI ----
namespace {
struct PredictScratch {
    double *d = nullptr;   // 8 (self) + 5*32 (others) + 2 (out)
    int *i = nullptr;      // 32 nbr + 1 cnt + 1 status
    double *h = nullptr;   // pinned mirror
    int *hi = nullptr;
    std::mutex mu;
};
PredictScratch g_ps;
constexpr int PS_D = 8 + 5 * SNB_MAX_AGENTS_PER_ENV + 2;
constexpr int PS_I = SNB_MAX_AGENTS_PER_ENV + 2;
}

extern "C" int snb_policy_predict_host(const SnbPolicyCfg *cfg, const double *self8, int32_t n_others, const double *others5,
                                       int32_t n_seg, const double *segs, double *out_v2, int32_t *nbr_ids, int32_t *n_nbr)
{
    SNB_REQUIRE(cfg && self8 && out_v2, SNB_EINVAL, "snb_policy_predict_host: NULL argument");
    SNB_REQUIRE(n_others >= 0 && n_others <= SNB_MAX_AGENTS_PER_ENV, SNB_EUNSUPPORTED, "snb_policy_predict_host: n_others=%d unsupported", n_others);
    SNB_REQUIRE(n_others == 0 || others5, SNB_EINVAL, "snb_policy_predict_host: others is NULL");
    std::lock_guard<std::mutex> lk(g_ps.mu);
    if (!g_ps.d) {
        SNB_CUDA_TRY(cudaMalloc(&g_ps.d, sizeof(double) * PS_D));
        SNB_CUDA_TRY(cudaMalloc(&g_ps.i, sizeof(int) * PS_I));
        SNB_CUDA_TRY(cudaMallocHost(&g_ps.h, sizeof(double) * PS_D));
        SNB_CUDA_TRY(cudaMallocHost(&g_ps.hi, sizeof(int) * PS_I));
    }
    SnbObstacles *obs = nullptr;
    if (n_seg > 0) { const int rc = snb_obstacles_create(&obs, segs, n_seg); if (rc) return rc; }
    // SoA on the device: self arrays (1 each) then the extras arrays (n_others each)
    double *h = g_ps.h;
    const int N = n_others > 0 ? n_others : 1;
    for (int k = 0; k < 8; ++k) h[k] = self8[k];
    for (int j = 0; j < n_others; ++j)
        for (int c = 0; c < 5; ++c) h[8 + c * N + j] = others5[5 * j + c];
    cudaStream_t s = 0;
    cudaError_t e = cudaMemcpyAsync(g_ps.d, h, sizeof(double) * (8 + 5 * N), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemsetAsync(g_ps.i, 0, sizeof(int) * PS_I, s);
    int rc = SNB_OK;
    if (e == cudaSuccess) {
        SnbCrowdState st;
        memset(&st, 0, sizeof(st));
        st.B = 1; st.H = 1; st.E = n_others; st.n_obs_extras = n_others;
        double *d = g_ps.d;
        st.px = d + 0; st.py = d + 1; st.vx = d + 2; st.vy = d + 3; st.radius = d + 4; st.gx = d + 5; st.gy = d + 6; st.vpref = d + 7;
        st.ex_px = d + 8; st.ex_py = d + 8 + N; st.ex_vx = d + 8 + 2 * N; st.ex_vy = d + 8 + 3 * N; st.ex_radius = d + 8 + 4 * N;
        double *d_out = d + 8 + 5 * SNB_MAX_AGENTS_PER_ENV;
        rc = snb_policy_step(cfg, &st, obs, d_out, g_ps.i, g_ps.i + SNB_MAX_AGENTS_PER_ENV, g_ps.i + SNB_MAX_AGENTS_PER_ENV + 1, s);
        if (rc == SNB_OK) {
            e = cudaMemcpyAsync(h, d_out, sizeof(double) * 2, cudaMemcpyDeviceToHost, s);
            if (e == cudaSuccess) e = cudaMemcpyAsync(g_ps.hi, g_ps.i, sizeof(int) * PS_I, cudaMemcpyDeviceToHost, s);
            if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        }
    }
    if (obs) snb_obstacles_destroy(obs);
    if (rc) return rc;
    if (e != cudaSuccess) { snb_set_error("snb_policy_predict_host: %s", cudaGetErrorString(e)); return SNB_ECUDA; }
    if (g_ps.hi[SNB_MAX_AGENTS_PER_ENV + 1] != 0) { snb_set_error("snb_policy_predict_host: device capacity exceeded (ORCA lines / obstacle neighbours)"); return SNB_EOVERFLOW; }
    out_v2[0] = h[0]; out_v2[1] = h[1];
    if (n_nbr) {
        *n_nbr = g_ps.hi[SNB_MAX_AGENTS_PER_ENV];
        if (nbr_ids) for (int k = 0; k < *n_nbr; ++k) nbr_ids[k] = g_ps.hi[k] - 1; // agent id H+e with H=1 -> ob index e
    }
    return SNB_OK;
}
