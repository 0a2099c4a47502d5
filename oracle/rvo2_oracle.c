/*
 * The ORCA arithmetic restated here follows the RVO2 Library (v2.0.x: Agent.cpp, KdTree.cpp, RVOSimulator.cpp),
 *   Copyright 2008 University of North Carolina at Chapel Hill,
 *   licensed under the Apache License, Version 2.0 (http://www.apache.org/licenses/LICENSE-2.0).
 * RVO2 is distributed on an "AS IS" BASIS, WITHOUT WARRANTIES OR CONDITIONS OF ANY KIND; see the License for the specific
 * language governing permissions and limitations.  <https://gamma.cs.unc.edu/RVO2/>   This file is a derived restatement, not a copy.
 */
/*
 * oracle/rvo2_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  PARITY UNPINNED.
 * See rvo2_oracle.h for provenance.  Every function names the RVO2 routine
 *  and the reference call site it serves.
 */
#include "rvo2_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define RVO_EPSILON 0.00001f
#define MAX_LEAF_SIZE 10

typedef struct { float x, y; } V2;
typedef struct { V2 point, direction; } Line;

/* ---- Vector2.h (Appendix A preamble) ---- */
/* Branch coverage counters for tests/test_orca_properties.py (which geometric case of Agent::computeNewVelocity / linearProgram1-3
 * a test population actually exercised).  Pure instrumentation: no arithmetic depends on them. */
enum { RVO_NBRANCH = 48 };
static unsigned long long g_branch[RVO_NBRANCH];
static __thread int g_br_on = 0; /* counted only while agent 0 (the agent the policy reads back) is processed */
#define BR(k) do { if (g_br_on) __atomic_fetch_add(&g_branch[(k)], 1ULL, __ATOMIC_RELAXED); } while (0)
void rvo_branch_counters(unsigned long long *out, int reset)
{
    for (int i = 0; i < RVO_NBRANCH; ++i) {
        out[i] = __atomic_load_n(&g_branch[i], __ATOMIC_RELAXED);
        if (reset) __atomic_store_n(&g_branch[i], 0ULL, __ATOMIC_RELAXED);
    }
}

static inline V2 v2(float x, float y) { V2 r; r.x = x; r.y = y; return r; }
static inline V2 vadd(V2 a, V2 b) { return v2(a.x + b.x, a.y + b.y); }
static inline V2 vsub(V2 a, V2 b) { return v2(a.x - b.x, a.y - b.y); }
static inline V2 vneg(V2 a) { return v2(-a.x, -a.y); }
static inline V2 vscale(float s, V2 a) { return v2(s * a.x, s * a.y); }
static inline float vdot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
/* Vector2::operator/(float): multiply by the reciprocal */
static inline V2 vdiv(V2 a, float s) { const float inv = 1.0f / s; return v2(a.x * inv, a.y * inv); }
static inline float absSq(V2 a) { return vdot(a, a); }
static inline float vabs(V2 a) { return sqrtf(vdot(a, a)); }
static inline float det(V2 a, V2 b) { return a.x * b.y - a.y * b.x; }
static inline V2 normalize(V2 a) { return vdiv(a, vabs(a)); }
static inline float sqr(float a) { return a * a; }
static inline float leftOf(V2 a, V2 b, V2 c) { return det(vsub(a, c), vsub(b, a)); }
static float distSqPointLineSegment(V2 a, V2 b, V2 c)
{
    const float r = vdot(vsub(c, a), vsub(b, a)) / absSq(vsub(b, a));
    if (r < 0.0f) return absSq(vsub(c, a));
    if (r > 1.0f) return absSq(vsub(c, b));
    return absSq(vsub(c, vadd(a, vscale(r, vsub(b, a)))));
}

typedef struct {
    int isConvex;
    int next, prev;
    V2 point, unitDir;
    int id;
} Obstacle;

typedef struct { float distSq; int id; } Nb;

typedef struct {
    V2 position, velocity, prefVelocity, newVelocity;
    float neighborDist, timeHorizon, timeHorizonObst, radius, maxSpeed;
    int maxNeighbors;
    Nb *agentNb; int nAgentNb;
    Nb *obstNb; int nObstNb, capObstNb;
    Line *lines; int nLines, capLines;
} Agent;

typedef struct { int begin, end, left, right; float maxX, maxY, minX, minY; } AgentTreeNode;
typedef struct ObstNode { int obstacle; struct ObstNode *left, *right; } ObstNode;

struct RvoSim {
    float timeStep, globalTime;
    /* defaults (unused by the reference, which always passes explicit values) */
    float dNeighborDist, dTimeHorizon, dTimeHorizonObst, dRadius, dMaxSpeed; int dMaxNeighbors; V2 dVelocity;
    Agent *agents; int nAgents, capAgents;
    Obstacle *obst; int nObst, capObst;
    int *kdAgents; int nKdAgents;
    AgentTreeNode *agentTree;
    ObstNode *obstTree;
};

/* ---- RVOSimulator.cpp (A.1, A.2) ---- */
RvoSim *rvo_create(float time_step, float neighbor_dist, int max_neighbors, float time_horizon,
                   float time_horizon_obst, float radius, float max_speed, float vx, float vy)
{
    RvoSim *s = (RvoSim *)calloc(1, sizeof(RvoSim));
    s->timeStep = time_step;
    s->dNeighborDist = neighbor_dist; s->dMaxNeighbors = max_neighbors;
    s->dTimeHorizon = time_horizon; s->dTimeHorizonObst = time_horizon_obst;
    s->dRadius = radius; s->dMaxSpeed = max_speed; s->dVelocity = v2(vx, vy);
    return s;
}

static void freeObstTree(ObstNode *n) { if (!n) return; freeObstTree(n->left); freeObstTree(n->right); free(n); }

void rvo_destroy(RvoSim *s)
{
    if (!s) return;
    for (int i = 0; i < s->nAgents; ++i) { free(s->agents[i].agentNb); free(s->agents[i].obstNb); free(s->agents[i].lines); }
    free(s->agents); free(s->obst); free(s->kdAgents); free(s->agentTree); freeObstTree(s->obstTree); free(s);
}

int rvo_add_agent(RvoSim *s, float px, float py, float neighbor_dist, int max_neighbors, float time_horizon,
                  float time_horizon_obst, float radius, float max_speed, float vx, float vy)
{
    if (s->nAgents == s->capAgents) {
        s->capAgents = s->capAgents ? 2 * s->capAgents : 16;
        s->agents = (Agent *)realloc(s->agents, sizeof(Agent) * (size_t)s->capAgents);
    }
    Agent *a = &s->agents[s->nAgents];
    memset(a, 0, sizeof(*a));
    a->position = v2(px, py); a->velocity = v2(vx, vy);
    a->neighborDist = neighbor_dist; a->maxNeighbors = max_neighbors;
    a->timeHorizon = time_horizon; a->timeHorizonObst = time_horizon_obst;
    a->radius = radius; a->maxSpeed = max_speed;
    a->agentNb = (Nb *)malloc(sizeof(Nb) * (size_t)(max_neighbors > 0 ? max_neighbors : 1));
    return s->nAgents++;
}

static int pushObst(RvoSim *s)
{
    if (s->nObst == s->capObst) {
        s->capObst = s->capObst ? 2 * s->capObst : 32;
        s->obst = (Obstacle *)realloc(s->obst, sizeof(Obstacle) * (size_t)s->capObst);
    }
    memset(&s->obst[s->nObst], 0, sizeof(Obstacle));
    return s->nObst++;
}

/* RVOSimulator::addObstacle (A.2); reference call site orca_plus.py:50-51 */
int rvo_add_obstacle(RvoSim *s, const float *xy, int n)
{
    if (n < 2) return -1;
    const int obstacleNo = s->nObst;
    for (int i = 0; i < n; ++i) {
        const int id = pushObst(s);
        Obstacle *o = &s->obst[id];
        o->point = v2(xy[2 * i], xy[2 * i + 1]);
        if (i != 0) { o->prev = id - 1; s->obst[id - 1].next = id; }
        if (i == n - 1) { o->next = obstacleNo; s->obst[obstacleNo].prev = id; }
        const int inext = (i == n - 1 ? 0 : i + 1);
        o->unitDir = normalize(vsub(v2(xy[2 * inext], xy[2 * inext + 1]), o->point));
        if (n == 2) {
            o->isConvex = 1;
        } else {
            const int iprev = (i == 0 ? n - 1 : i - 1);
            o->isConvex = leftOf(v2(xy[2 * iprev], xy[2 * iprev + 1]), o->point,
                                 v2(xy[2 * inext], xy[2 * inext + 1])) >= 0.0f;
        }
        o->id = id;
    }
    return obstacleNo;
}

/* ---- KdTree.cpp: obstacle BSP (A.5) ---- */
static int pairLess(int a1, int a2, int b1, int b2) { return a1 < b1 || (a1 == b1 && a2 < b2); }
static int imax(int a, int b) { return a > b ? a : b; }
static int imin(int a, int b) { return a < b ? a : b; }

static ObstNode *buildObstacleTreeRecursive(RvoSim *s, const int *obstacles, int n)
{
    if (n == 0) return NULL;
    ObstNode *node = (ObstNode *)calloc(1, sizeof(ObstNode));
    int optimalSplit = 0, minLeft = n, minRight = n;

    for (int i = 0; i < n; ++i) {
        int leftSize = 0, rightSize = 0;
        const V2 i1 = s->obst[obstacles[i]].point, i2 = s->obst[s->obst[obstacles[i]].next].point;
        for (int j = 0; j < n; ++j) {
            if (i == j) continue;
            const V2 j1 = s->obst[obstacles[j]].point, j2 = s->obst[s->obst[obstacles[j]].next].point;
            const float j1LeftOfI = leftOf(i1, i2, j1), j2LeftOfI = leftOf(i1, i2, j2);
            if (j1LeftOfI >= -RVO_EPSILON && j2LeftOfI >= -RVO_EPSILON) ++leftSize;
            else if (j1LeftOfI <= RVO_EPSILON && j2LeftOfI <= RVO_EPSILON) ++rightSize;
            else { ++leftSize; ++rightSize; }
            if (!pairLess(imax(leftSize, rightSize), imin(leftSize, rightSize),
                          imax(minLeft, minRight), imin(minLeft, minRight))) break;
        }
        if (pairLess(imax(leftSize, rightSize), imin(leftSize, rightSize),
                     imax(minLeft, minRight), imin(minLeft, minRight))) {
            minLeft = leftSize; minRight = rightSize; optimalSplit = i;
        }
    }

    int *leftObst = (int *)malloc(sizeof(int) * (size_t)(minLeft + 1));
    int *rightObst = (int *)malloc(sizeof(int) * (size_t)(minRight + 1));
    int leftCounter = 0, rightCounter = 0;
    const int i = optimalSplit;
    const int I1 = obstacles[i];

    for (int j = 0; j < n; ++j) {
        if (i == j) continue;
        const int J1 = obstacles[j];
        const int J2 = s->obst[J1].next;
        const int I2 = s->obst[I1].next;
        const V2 i1 = s->obst[I1].point, i2 = s->obst[I2].point;
        const V2 j1 = s->obst[J1].point, j2 = s->obst[J2].point;
        const float j1LeftOfI = leftOf(i1, i2, j1), j2LeftOfI = leftOf(i1, i2, j2);
        if (j1LeftOfI >= -RVO_EPSILON && j2LeftOfI >= -RVO_EPSILON) {
            leftObst[leftCounter++] = J1;
        } else if (j1LeftOfI <= RVO_EPSILON && j2LeftOfI <= RVO_EPSILON) {
            rightObst[rightCounter++] = J1;
        } else {
            /* split obstacle j */
            const float t = det(vsub(i2, i1), vsub(j1, i1)) / det(vsub(i2, i1), vsub(j1, j2));
            const V2 splitpoint = vadd(j1, vscale(t, vsub(j2, j1)));
            const int nid = pushObst(s); /* may realloc: re-read through s->obst below */
            Obstacle *no = &s->obst[nid];
            no->point = splitpoint; no->prev = J1; no->next = J2; no->isConvex = 1;
            no->unitDir = s->obst[J1].unitDir; no->id = nid;
            s->obst[J1].next = nid; s->obst[J2].prev = nid;
            if (j1LeftOfI > 0.0f) { leftObst[leftCounter++] = J1; rightObst[rightCounter++] = nid; }
            else { rightObst[rightCounter++] = J1; leftObst[leftCounter++] = nid; }
        }
    }
    node->obstacle = I1;
    node->left = buildObstacleTreeRecursive(s, leftObst, leftCounter);
    node->right = buildObstacleTreeRecursive(s, rightObst, rightCounter);
    free(leftObst); free(rightObst);
    return node;
}

/* RVOSimulator::processObstacles -> KdTree::buildObstacleTree; call site orca_plus.py:52-53 */
void rvo_process_obstacles(RvoSim *s)
{
    freeObstTree(s->obstTree); s->obstTree = NULL;
    const int n = s->nObst;
    int *all = (int *)malloc(sizeof(int) * (size_t)(n + 1));
    for (int i = 0; i < n; ++i) all[i] = i;
    s->obstTree = buildObstacleTreeRecursive(s, all, n);
    free(all);
}

/* ---- KdTree.cpp: agent tree (A.3) ---- */
static void buildAgentTreeRecursive(RvoSim *s, int begin, int end, int node)
{
    AgentTreeNode *t = &s->agentTree[node];
    t->begin = begin; t->end = end;
    t->minX = t->maxX = s->agents[s->kdAgents[begin]].position.x;
    t->minY = t->maxY = s->agents[s->kdAgents[begin]].position.y;
    for (int i = begin + 1; i < end; ++i) {
        const V2 p = s->agents[s->kdAgents[i]].position;
        t->maxX = fmaxf(t->maxX, p.x); t->minX = fminf(t->minX, p.x);
        t->maxY = fmaxf(t->maxY, p.y); t->minY = fminf(t->minY, p.y);
    }
    if (end - begin > MAX_LEAF_SIZE) {
        const int isVertical = (t->maxX - t->minX > t->maxY - t->minY);
        const float splitValue = (isVertical ? 0.5f * (t->maxX + t->minX) : 0.5f * (t->maxY + t->minY));
        int left = begin, right = end;
#define COORD(k) (isVertical ? s->agents[s->kdAgents[(k)]].position.x : s->agents[s->kdAgents[(k)]].position.y)
        while (left < right) {
            while (left < right && COORD(left) < splitValue) ++left;
            while (right > left && COORD(right - 1) >= splitValue) --right;
            if (left < right) {
                const int tmp = s->kdAgents[left]; s->kdAgents[left] = s->kdAgents[right - 1]; s->kdAgents[right - 1] = tmp;
                ++left; --right;
            }
        }
#undef COORD
        if (left == begin) { ++left; ++right; }
        t->left = node + 1;
        t->right = node + 2 * (left - begin);
        const int l = t->left, r = t->right; /* t may not dangle: agentTree is not reallocated here */
        buildAgentTreeRecursive(s, begin, left, l);
        buildAgentTreeRecursive(s, left, end, r);
    }
}

static void buildAgentTree(RvoSim *s)
{
    if (s->nKdAgents < s->nAgents) {
        s->kdAgents = (int *)realloc(s->kdAgents, sizeof(int) * (size_t)s->nAgents);
        for (int i = s->nKdAgents; i < s->nAgents; ++i) s->kdAgents[i] = i;
        s->nKdAgents = s->nAgents;
        s->agentTree = (AgentTreeNode *)realloc(s->agentTree, sizeof(AgentTreeNode) * (size_t)(2 * s->nAgents - 1));
    }
    if (s->nKdAgents > 0) buildAgentTreeRecursive(s, 0, s->nKdAgents, 0);
}

/* Agent::insertAgentNeighbor (A.4) */
static void insertAgentNeighbor(RvoSim *s, int self, int other, float *rangeSq)
{
    Agent *a = &s->agents[self];
    if (self == other) return;
    const float distSq = absSq(vsub(a->position, s->agents[other].position));
    if (distSq < *rangeSq) {
        if (a->nAgentNb < a->maxNeighbors) { a->agentNb[a->nAgentNb].distSq = distSq; a->agentNb[a->nAgentNb].id = other; a->nAgentNb++; }
        int i = a->nAgentNb - 1;
        while (i != 0 && distSq < a->agentNb[i - 1].distSq) { a->agentNb[i] = a->agentNb[i - 1]; --i; }
        a->agentNb[i].distSq = distSq; a->agentNb[i].id = other;
        if (a->nAgentNb == a->maxNeighbors) *rangeSq = a->agentNb[a->nAgentNb - 1].distSq;
    }
}

static void queryAgentTreeRecursive(RvoSim *s, int self, float *rangeSq, int node)
{
    const AgentTreeNode *t = &s->agentTree[node];
    if (t->end - t->begin <= MAX_LEAF_SIZE) {
        for (int i = t->begin; i < t->end; ++i) insertAgentNeighbor(s, self, s->kdAgents[i], rangeSq);
    } else {
        const V2 p = s->agents[self].position;
        const AgentTreeNode *L = &s->agentTree[t->left], *R = &s->agentTree[t->right];
        const float distSqLeft = sqr(fmaxf(0.0f, L->minX - p.x)) + sqr(fmaxf(0.0f, p.x - L->maxX)) +
                                 sqr(fmaxf(0.0f, L->minY - p.y)) + sqr(fmaxf(0.0f, p.y - L->maxY));
        const float distSqRight = sqr(fmaxf(0.0f, R->minX - p.x)) + sqr(fmaxf(0.0f, p.x - R->maxX)) +
                                  sqr(fmaxf(0.0f, R->minY - p.y)) + sqr(fmaxf(0.0f, p.y - R->maxY));
        if (distSqLeft < distSqRight) {
            if (distSqLeft < *rangeSq) {
                queryAgentTreeRecursive(s, self, rangeSq, t->left);
                if (distSqRight < *rangeSq) queryAgentTreeRecursive(s, self, rangeSq, t->right);
            }
        } else {
            if (distSqRight < *rangeSq) {
                queryAgentTreeRecursive(s, self, rangeSq, t->right);
                if (distSqLeft < *rangeSq) queryAgentTreeRecursive(s, self, rangeSq, t->left);
            }
        }
    }
}

/* Agent::insertObstacleNeighbor (A.5) */
static void insertObstacleNeighbor(RvoSim *s, int self, int obstacle, float rangeSq)
{
    Agent *a = &s->agents[self];
    const Obstacle *o = &s->obst[obstacle];
    const float distSq = distSqPointLineSegment(o->point, s->obst[o->next].point, a->position);
    if (distSq < rangeSq) {
        if (a->nObstNb == a->capObstNb) {
            a->capObstNb = a->capObstNb ? 2 * a->capObstNb : 16;
            a->obstNb = (Nb *)realloc(a->obstNb, sizeof(Nb) * (size_t)a->capObstNb);
        }
        a->nObstNb++;
        int i = a->nObstNb - 1;
        while (i != 0 && distSq < a->obstNb[i - 1].distSq) { a->obstNb[i] = a->obstNb[i - 1]; --i; }
        a->obstNb[i].distSq = distSq; a->obstNb[i].id = obstacle;
    }
}

static void queryObstacleTreeRecursive(RvoSim *s, int self, float rangeSq, const ObstNode *node)
{
    if (node == NULL) return;
    const Obstacle *o1 = &s->obst[node->obstacle];
    const Obstacle *o2 = &s->obst[o1->next];
    const float agentLeftOfLine = leftOf(o1->point, o2->point, s->agents[self].position);
    queryObstacleTreeRecursive(s, self, rangeSq, (agentLeftOfLine >= 0.0f ? node->left : node->right));
    const float distSqLine = sqr(agentLeftOfLine) / absSq(vsub(o2->point, o1->point));
    if (distSqLine < rangeSq) {
        if (agentLeftOfLine < 0.0f) insertObstacleNeighbor(s, self, node->obstacle, rangeSq);
        queryObstacleTreeRecursive(s, self, rangeSq, (agentLeftOfLine >= 0.0f ? node->right : node->left));
    }
}

/* Agent::computeNeighbors (A.4) */
static void computeNeighbors(RvoSim *s, int self)
{
    Agent *a = &s->agents[self];
    a->nObstNb = 0;
    float rangeSq = sqr(a->timeHorizonObst * a->maxSpeed + a->radius);
    queryObstacleTreeRecursive(s, self, rangeSq, s->obstTree);
    a->nAgentNb = 0;
    if (a->maxNeighbors > 0) {
        rangeSq = sqr(a->neighborDist);
        queryAgentTreeRecursive(s, self, &rangeSq, 0);
    }
}

/* ---- Agent.cpp: linear programs (A.8) ---- */
static int linearProgram1(const Line *lines, int lineNo, float radius, V2 optVelocity, int directionOpt, V2 *result)
{
    const float dotProduct = vdot(lines[lineNo].point, lines[lineNo].direction);
    const float discriminant = sqr(dotProduct) + sqr(radius) - absSq(lines[lineNo].point);
    if (discriminant < 0.0f) { BR(30); return 0; }
    const float sqrtDiscriminant = sqrtf(discriminant);
    float tLeft = -dotProduct - sqrtDiscriminant;
    float tRight = -dotProduct + sqrtDiscriminant;
    for (int i = 0; i < lineNo; ++i) {
        const float denominator = det(lines[lineNo].direction, lines[i].direction);
        const float numerator = det(lines[i].direction, vsub(lines[lineNo].point, lines[i].point));
        if (fabsf(denominator) <= RVO_EPSILON) {
            if (numerator < 0.0f) { BR(31); return 0; }
            BR(41);
            continue;
        }
        const float t = numerator / denominator;
        if (denominator >= 0.0f) tRight = fminf(tRight, t);
        else tLeft = fmaxf(tLeft, t);
        if (tLeft > tRight) { BR(32); return 0; }
    }
    if (directionOpt) {
        BR(38);
        if (vdot(optVelocity, lines[lineNo].direction) > 0.0f)
            *result = vadd(lines[lineNo].point, vscale(tRight, lines[lineNo].direction));
        else
            *result = vadd(lines[lineNo].point, vscale(tLeft, lines[lineNo].direction));
    } else {
        const float t = vdot(lines[lineNo].direction, vsub(optVelocity, lines[lineNo].point));
        if (t < tLeft) { BR(39); *result = vadd(lines[lineNo].point, vscale(tLeft, lines[lineNo].direction)); }
        else if (t > tRight) { BR(40); *result = vadd(lines[lineNo].point, vscale(tRight, lines[lineNo].direction)); }
        else { BR(42); *result = vadd(lines[lineNo].point, vscale(t, lines[lineNo].direction)); }
    }
    return 1;
}

static int linearProgram2(const Line *lines, int n, float radius, V2 optVelocity, int directionOpt, V2 *result)
{
    if (directionOpt) *result = vscale(radius, optVelocity); /* optVelocity * radius */
    else if (absSq(optVelocity) > sqr(radius)) { BR(33); *result = vscale(radius, normalize(optVelocity)); }
    else *result = optVelocity;
    for (int i = 0; i < n; ++i) {
        if (det(lines[i].direction, vsub(lines[i].point, *result)) > 0.0f) {
            const V2 tempResult = *result;
            if (!linearProgram1(lines, i, radius, optVelocity, directionOpt, result)) {
                *result = tempResult;
                return i;
            }
        }
    }
    return n;
}

static void linearProgram3(const Line *lines, int n, int numObstLines, int beginLine, float radius, V2 *result)
{
    float distance = 0.0f;
    BR(34);
    Line *projLines = (Line *)malloc(sizeof(Line) * (size_t)(n + 1));
    for (int i = beginLine; i < n; ++i) {
        if (det(lines[i].direction, vsub(lines[i].point, *result)) > distance) {
            int np = 0;
            for (int j = 0; j < numObstLines; ++j) projLines[np++] = lines[j];
            for (int j = numObstLines; j < i; ++j) {
                Line line;
                const float determinant = det(lines[i].direction, lines[j].direction);
                if (fabsf(determinant) <= RVO_EPSILON) {
                    if (vdot(lines[i].direction, lines[j].direction) > 0.0f) { BR(35); continue; }
                    BR(36);
                    line.point = vscale(0.5f, vadd(lines[i].point, lines[j].point));
                } else {
                    line.point = vadd(lines[i].point,
                                      vscale(det(lines[j].direction, vsub(lines[i].point, lines[j].point)) / determinant,
                                             lines[i].direction));
                }
                line.direction = normalize(vsub(lines[j].direction, lines[i].direction));
                projLines[np++] = line;
            }
            const V2 tempResult = *result;
            if (linearProgram2(projLines, np, radius, v2(-lines[i].direction.y, lines[i].direction.x), 1, result) < np) {
                BR(37);
                *result = tempResult;
            }
            distance = det(lines[i].direction, vsub(lines[i].point, *result));
        }
    }
    free(projLines);
}

static void pushLine(Agent *a, Line l)
{
    if (a->nLines == a->capLines) {
        a->capLines = a->capLines ? 2 * a->capLines : 32;
        a->lines = (Line *)realloc(a->lines, sizeof(Line) * (size_t)a->capLines);
    }
    a->lines[a->nLines++] = l;
}

/* Agent::computeNewVelocity (A.6 obstacle lines, A.7 agent lines, A.8 LPs) */
static void computeNewVelocity(RvoSim *s, int self)
{
    Agent *a = &s->agents[self];
    a->nLines = 0;
    g_br_on = (self == 0);
    const float invTimeHorizonObst = 1.0f / a->timeHorizonObst;

    for (int i = 0; i < a->nObstNb; ++i) {
        int o1 = a->obstNb[i].id;
        int o2 = s->obst[o1].next;
        const V2 relativePosition1 = vsub(s->obst[o1].point, a->position);
        const V2 relativePosition2 = vsub(s->obst[o2].point, a->position);

        int alreadyCovered = 0;
        for (int j = 0; j < a->nLines; ++j) {
            if (det(vsub(vscale(invTimeHorizonObst, relativePosition1), a->lines[j].point), a->lines[j].direction) -
                        invTimeHorizonObst * a->radius >= -RVO_EPSILON &&
                det(vsub(vscale(invTimeHorizonObst, relativePosition2), a->lines[j].point), a->lines[j].direction) -
                        invTimeHorizonObst * a->radius >= -RVO_EPSILON) {
                alreadyCovered = 1;
                BR(0);
                break;
            }
        }
        if (alreadyCovered) continue;

        const float distSq1 = absSq(relativePosition1);
        const float distSq2 = absSq(relativePosition2);
        const float radiusSq = sqr(a->radius);
        const V2 obstacleVector = vsub(s->obst[o2].point, s->obst[o1].point);
        const float sP = vdot(vneg(relativePosition1), obstacleVector) / absSq(obstacleVector);
        const float distSqLine = absSq(vsub(vneg(relativePosition1), vscale(sP, obstacleVector)));
        Line line;

        if (sP < 0.0f && distSq1 <= radiusSq) {
            BR(1);
            if (s->obst[o1].isConvex) {
                line.point = v2(0.0f, 0.0f);
                line.direction = normalize(v2(-relativePosition1.y, relativePosition1.x));
                pushLine(a, line);
            }
            continue;
        } else if (sP > 1.0f && distSq2 <= radiusSq) {
            BR(2);
            if (s->obst[o2].isConvex && det(relativePosition2, s->obst[o2].unitDir) >= 0.0f) {
                line.point = v2(0.0f, 0.0f);
                line.direction = normalize(v2(-relativePosition2.y, relativePosition2.x));
                pushLine(a, line);
            }
            continue;
        } else if (sP >= 0.0f && sP < 1.0f && distSqLine <= radiusSq) {
            BR(3);
            line.point = v2(0.0f, 0.0f);
            line.direction = vneg(s->obst[o1].unitDir);
            pushLine(a, line);
            continue;
        }

        V2 leftLegDirection, rightLegDirection;
        if (sP < 0.0f && distSqLine <= radiusSq) {
            BR(4);
            if (!s->obst[o1].isConvex) { BR(16); continue; }
            o2 = o1;
            const float leg1 = sqrtf(distSq1 - radiusSq);
            leftLegDirection = vdiv(v2(relativePosition1.x * leg1 - relativePosition1.y * a->radius,
                                       relativePosition1.x * a->radius + relativePosition1.y * leg1), distSq1);
            rightLegDirection = vdiv(v2(relativePosition1.x * leg1 + relativePosition1.y * a->radius,
                                        -relativePosition1.x * a->radius + relativePosition1.y * leg1), distSq1);
        } else if (sP > 1.0f && distSqLine <= radiusSq) {
            BR(5);
            if (!s->obst[o2].isConvex) { BR(16); continue; }
            o1 = o2;
            const float leg2 = sqrtf(distSq2 - radiusSq);
            leftLegDirection = vdiv(v2(relativePosition2.x * leg2 - relativePosition2.y * a->radius,
                                       relativePosition2.x * a->radius + relativePosition2.y * leg2), distSq2);
            rightLegDirection = vdiv(v2(relativePosition2.x * leg2 + relativePosition2.y * a->radius,
                                        -relativePosition2.x * a->radius + relativePosition2.y * leg2), distSq2);
        } else {
            BR(6);
            if (s->obst[o1].isConvex) {
                const float leg1 = sqrtf(distSq1 - radiusSq);
                leftLegDirection = vdiv(v2(relativePosition1.x * leg1 - relativePosition1.y * a->radius,
                                           relativePosition1.x * a->radius + relativePosition1.y * leg1), distSq1);
            } else {
                leftLegDirection = vneg(s->obst[o1].unitDir);
            }
            if (s->obst[o2].isConvex) {
                const float leg2 = sqrtf(distSq2 - radiusSq);
                rightLegDirection = vdiv(v2(relativePosition2.x * leg2 + relativePosition2.y * a->radius,
                                            -relativePosition2.x * a->radius + relativePosition2.y * leg2), distSq2);
            } else {
                rightLegDirection = s->obst[o1].unitDir;
            }
        }

        const int leftNeighbor = s->obst[o1].prev;
        int isLeftLegForeign = 0, isRightLegForeign = 0;
        if (s->obst[o1].isConvex && det(leftLegDirection, vneg(s->obst[leftNeighbor].unitDir)) >= 0.0f) {
            leftLegDirection = vneg(s->obst[leftNeighbor].unitDir);
            isLeftLegForeign = 1;
            BR(7);
        }
        if (s->obst[o2].isConvex && det(rightLegDirection, s->obst[o2].unitDir) <= 0.0f) {
            rightLegDirection = s->obst[o2].unitDir;
            isRightLegForeign = 1;
            BR(8);
        }

        const V2 leftCutoff = vscale(invTimeHorizonObst, vsub(s->obst[o1].point, a->position));
        const V2 rightCutoff = vscale(invTimeHorizonObst, vsub(s->obst[o2].point, a->position));
        const V2 cutoffVec = vsub(rightCutoff, leftCutoff);

        const float t = (o1 == o2 ? 0.5f : vdot(vsub(a->velocity, leftCutoff), cutoffVec) / absSq(cutoffVec));
        const float tLeft = vdot(vsub(a->velocity, leftCutoff), leftLegDirection);
        const float tRight = vdot(vsub(a->velocity, rightCutoff), rightLegDirection);

        if ((t < 0.0f && tLeft < 0.0f) || (o1 == o2 && tLeft < 0.0f && tRight < 0.0f)) {
            BR(9);
            const V2 unitW = normalize(vsub(a->velocity, leftCutoff));
            line.direction = v2(unitW.y, -unitW.x);
            line.point = vadd(leftCutoff, vscale(a->radius * invTimeHorizonObst, unitW));
            pushLine(a, line);
            continue;
        } else if (t > 1.0f && tRight < 0.0f) {
            BR(10);
            const V2 unitW = normalize(vsub(a->velocity, rightCutoff));
            line.direction = v2(unitW.y, -unitW.x);
            line.point = vadd(rightCutoff, vscale(a->radius * invTimeHorizonObst, unitW));
            pushLine(a, line);
            continue;
        }

        const float distSqCutoff = ((t < 0.0f || t > 1.0f || o1 == o2) ? INFINITY
                                    : absSq(vsub(a->velocity, vadd(leftCutoff, vscale(t, cutoffVec)))));
        const float distSqLeft = ((tLeft < 0.0f) ? INFINITY
                                  : absSq(vsub(a->velocity, vadd(leftCutoff, vscale(tLeft, leftLegDirection)))));
        const float distSqRight = ((tRight < 0.0f) ? INFINITY
                                   : absSq(vsub(a->velocity, vadd(rightCutoff, vscale(tRight, rightLegDirection)))));

        if (distSqCutoff <= distSqLeft && distSqCutoff <= distSqRight) {
            BR(11);
            line.direction = vneg(s->obst[o1].unitDir);
            line.point = vadd(leftCutoff, vscale(a->radius * invTimeHorizonObst, v2(-line.direction.y, line.direction.x)));
            pushLine(a, line);
            continue;
        } else if (distSqLeft <= distSqRight) {
            if (isLeftLegForeign) { BR(14); continue; }
            BR(12);
            line.direction = leftLegDirection;
            line.point = vadd(leftCutoff, vscale(a->radius * invTimeHorizonObst, v2(-line.direction.y, line.direction.x)));
            pushLine(a, line);
            continue;
        } else {
            if (isRightLegForeign) { BR(15); continue; }
            BR(13);
            line.direction = vneg(rightLegDirection);
            line.point = vadd(rightCutoff, vscale(a->radius * invTimeHorizonObst, v2(-line.direction.y, line.direction.x)));
            pushLine(a, line);
            continue;
        }
    }

    const int numObstLines = a->nLines;
    const float invTimeHorizon = 1.0f / a->timeHorizon;

    for (int i = 0; i < a->nAgentNb; ++i) {
        const Agent *other = &s->agents[a->agentNb[i].id];
        const V2 relativePosition = vsub(other->position, a->position);
        const V2 relativeVelocity = vsub(a->velocity, other->velocity);
        const float distSq = absSq(relativePosition);
        const float combinedRadius = a->radius + other->radius;
        const float combinedRadiusSq = sqr(combinedRadius);
        Line line;
        V2 u;
        if (distSq > combinedRadiusSq) {
            const V2 w = vsub(relativeVelocity, vscale(invTimeHorizon, relativePosition));
            const float wLengthSq = absSq(w);
            const float dotProduct1 = vdot(w, relativePosition);
            if (dotProduct1 < 0.0f && sqr(dotProduct1) > combinedRadiusSq * wLengthSq) {
                BR(20);
                const float wLength = sqrtf(wLengthSq);
                const V2 unitW = vdiv(w, wLength);
                line.direction = v2(unitW.y, -unitW.x);
                u = vscale(combinedRadius * invTimeHorizon - wLength, unitW);
            } else {
                const float leg = sqrtf(distSq - combinedRadiusSq);
                if (det(relativePosition, w) > 0.0f) {
                    BR(21);
                    line.direction = vdiv(v2(relativePosition.x * leg - relativePosition.y * combinedRadius,
                                             relativePosition.x * combinedRadius + relativePosition.y * leg), distSq);
                } else {
                    BR(22);
                    line.direction = vneg(vdiv(v2(relativePosition.x * leg + relativePosition.y * combinedRadius,
                                                  -relativePosition.x * combinedRadius + relativePosition.y * leg), distSq));
                }
                const float dotProduct2 = vdot(relativeVelocity, line.direction);
                u = vsub(vscale(dotProduct2, line.direction), relativeVelocity);
            }
        } else {
            BR(23);
            const float invTimeStep = 1.0f / s->timeStep;
            const V2 w = vsub(relativeVelocity, vscale(invTimeStep, relativePosition));
            const float wLength = vabs(w);
            const V2 unitW = vdiv(w, wLength);
            line.direction = v2(unitW.y, -unitW.x);
            u = vscale(combinedRadius * invTimeStep - wLength, unitW);
        }
        line.point = vadd(a->velocity, vscale(0.5f, u));
        pushLine(a, line);
    }

    const int lineFail = linearProgram2(a->lines, a->nLines, a->maxSpeed, a->prefVelocity, 0, &a->newVelocity);
    if (lineFail < a->nLines)
        linearProgram3(a->lines, a->nLines, numObstLines, lineFail, a->maxSpeed, &a->newVelocity);
}

/* RVOSimulator::doStep (A.1); reference call site orca.py:128, orca_plus.py:84 */
void rvo_do_step(RvoSim *s)
{
    buildAgentTree(s);
    for (int i = 0; i < s->nAgents; ++i) { computeNeighbors(s, i); computeNewVelocity(s, i); }
    for (int i = 0; i < s->nAgents; ++i) {
        Agent *a = &s->agents[i];
        a->velocity = a->newVelocity;
        a->position = vadd(a->position, vscale(s->timeStep, a->velocity)); /* velocity_ * timeStep_ */
    }
    s->globalTime += s->timeStep;
}

void rvo_set_agent_position(RvoSim *s, int i, float x, float y) { s->agents[i].position = v2(x, y); }
void rvo_set_agent_velocity(RvoSim *s, int i, float x, float y) { s->agents[i].velocity = v2(x, y); }
void rvo_set_agent_pref_velocity(RvoSim *s, int i, float x, float y) { s->agents[i].prefVelocity = v2(x, y); }
void rvo_get_agent_position(const RvoSim *s, int i, float *o) { o[0] = s->agents[i].position.x; o[1] = s->agents[i].position.y; }
void rvo_get_agent_velocity(const RvoSim *s, int i, float *o) { o[0] = s->agents[i].velocity.x; o[1] = s->agents[i].velocity.y; }
void rvo_get_agent_pref_velocity(const RvoSim *s, int i, float *o) { o[0] = s->agents[i].prefVelocity.x; o[1] = s->agents[i].prefVelocity.y; }
float rvo_get_agent_max_speed(const RvoSim *s, int i) { return s->agents[i].maxSpeed; }
int rvo_get_num_agents(const RvoSim *s) { return s->nAgents; }
int rvo_get_num_obstacle_vertices(const RvoSim *s) { return s->nObst; }
float rvo_get_global_time(const RvoSim *s) { return s->globalTime; }
int rvo_get_agent_num_agent_neighbors(const RvoSim *s, int i) { return s->agents[i].nAgentNb; }
int rvo_get_agent_agent_neighbor(const RvoSim *s, int i, int k) { return s->agents[i].agentNb[k].id; }
int rvo_get_agent_num_obstacle_neighbors(const RvoSim *s, int i) { return s->agents[i].nObstNb; }
int rvo_get_agent_obstacle_neighbor(const RvoSim *s, int i, int k) { return s->agents[i].obstNb[k].id; }
int rvo_get_agent_num_orca_lines(const RvoSim *s, int i) { return s->agents[i].nLines; }
void rvo_get_agent_orca_line(const RvoSim *s, int i, int k, float *o)
{
    const Line *l = &s->agents[i].lines[k];
    o[0] = l->point.x; o[1] = l->point.y; o[2] = l->direction.x; o[3] = l->direction.y;
}
void rvo_get_obstacle_vertex(const RvoSim *s, int i, float *o)
{
    const Obstacle *b = &s->obst[i];
    o[0] = b->point.x; o[1] = b->point.y; o[2] = b->unitDir.x; o[3] = b->unitDir.y;
    o[4] = (float)b->next; o[5] = (float)b->prev; o[6] = (float)b->isConvex;
}
