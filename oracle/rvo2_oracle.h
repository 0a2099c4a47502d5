/*
 * The ORCA arithmetic restated here follows the RVO2 Library (v2.0.x: Agent.cpp, KdTree.cpp, RVOSimulator.cpp),
 *   Copyright 2008 University of North Carolina at Chapel Hill,
 *   licensed under the Apache License, Version 2.0 (http://www.apache.org/licenses/LICENSE-2.0).
 * RVO2 is distributed on an "AS IS" BASIS, WITHOUT WARRANTIES OR CONDITIONS OF ANY KIND; see the License for the specific
 * language governing permissions and limitations.  <https://gamma.cs.unc.edu/RVO2/>   This file is a derived restatement, not a copy.
 */
/*
 * oracle/rvo2_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU float32 restatement of the RVO2 library (v2.0.x: Agent.cpp, KdTree.cpp,
 * RVOSimulator.cpp, Vector2.h) as it is reached through Python-RVO2 from the
 * reference's call sites crowd_sim_plus/envs/policy/orca.py:95-129 and
 * orca_plus.py:45-85.  RVO2 is a third-party dependency of the reference that
 * is neither vendored nor pinned (reference README.md:62-68), so this file
 * restates the published algorithm.
 *
 * PARITY UNPINNED: the reference holds no test, fixture or golden vector for
 * ORCA outputs and the upstream rvo2 module cannot be installed here, so
 * nothing anchors this restatement to the real library beyond the published
 * algorithm.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may use anything under oracle/.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off, no -ffast-math, so
 * every float operation is a single IEEE-754 binary32 operation, the same
 * arithmetic the CUDA kernel performs with -fmad=false).
 */
#ifndef RVO2_ORACLE_H
#define RVO2_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct RvoSim RvoSim;

/* PyRVOSimulator(timeStep, neighborDist, maxNeighbors, timeHorizon,
 *                timeHorizonObst, radius, maxSpeed, velocity=(0,0))      */
RvoSim *rvo_create(float time_step, float neighbor_dist, int max_neighbors,
                   float time_horizon, float time_horizon_obst, float radius,
                   float max_speed, float vx, float vy);
void rvo_destroy(RvoSim *s);

int rvo_add_agent(RvoSim *s, float px, float py, float neighbor_dist,
                  int max_neighbors, float time_horizon,
                  float time_horizon_obst, float radius, float max_speed,
                  float vx, float vy);
/* xy = n vertices (x0,y0,x1,y1,...); returns index of first vertex or -1 */
int rvo_add_obstacle(RvoSim *s, const float *xy, int n);
void rvo_process_obstacles(RvoSim *s);

void rvo_set_agent_position(RvoSim *s, int i, float x, float y);
void rvo_set_agent_velocity(RvoSim *s, int i, float x, float y);
void rvo_set_agent_pref_velocity(RvoSim *s, int i, float x, float y);
void rvo_get_agent_position(const RvoSim *s, int i, float *out2);
void rvo_get_agent_velocity(const RvoSim *s, int i, float *out2);
void rvo_get_agent_pref_velocity(const RvoSim *s, int i, float *out2);
float rvo_get_agent_max_speed(const RvoSim *s, int i);
int rvo_get_num_agents(const RvoSim *s);
int rvo_get_num_obstacle_vertices(const RvoSim *s);
float rvo_get_global_time(const RvoSim *s);

void rvo_do_step(RvoSim *s);

/* introspection used by the parity tests (results of the LAST do_step) */
int rvo_get_agent_num_agent_neighbors(const RvoSim *s, int i);
int rvo_get_agent_agent_neighbor(const RvoSim *s, int i, int k);
int rvo_get_agent_num_obstacle_neighbors(const RvoSim *s, int i);
int rvo_get_agent_obstacle_neighbor(const RvoSim *s, int i, int k);
int rvo_get_agent_num_orca_lines(const RvoSim *s, int i);
/* out4 = point.x, point.y, direction.x, direction.y */
void rvo_get_agent_orca_line(const RvoSim *s, int i, int k, float *out4);
/* out7 = point.x point.y unitDir.x unitDir.y next prev isConvex (as float) */
void rvo_get_obstacle_vertex(const RvoSim *s, int i, float *out7);

/* branch coverage counters (48 slots, ids in tests/orca_props.py BRANCHES); reset != 0 clears them after the read */
void rvo_branch_counters(unsigned long long *out48, int reset);

#ifdef __cplusplus
}
#endif
#endif
