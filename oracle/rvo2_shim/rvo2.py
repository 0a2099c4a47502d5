"""oracle/rvo2_shim/rvo2.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  PARITY UNPINNED.

A module named `rvo2` exposing the Python-RVO2 surface the reference binds to
(`rvo2.PyRVOSimulator`, call sites crowd_sim_plus/envs/policy/orca.py:95-129 and
orca_plus.py:45-85; SURVEY.md 8b "rvo2 FFI"), implemented on oracle/liboracle.so.
Putting this directory on sys.path lets the reference's own orca.py / orca_plus.py
run UNMODIFIED in the build container (oracle/gen_golden.py, bench --impl reference).
Like the Cython original, every Python float is narrowed to C float on the way in
and widened on the way out.
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle_lib as _ol  # noqa: E402


class PyRVOSimulator:
    def __init__(self, timeStep, neighborDist, maxNeighbors, timeHorizon, timeHorizonObst, radius, maxSpeed,
                 velocity=(0, 0)):
        self._L = _ol.lib()
        self._s = self._L.rvo_create(timeStep, neighborDist, int(maxNeighbors), timeHorizon, timeHorizonObst, radius,
                                     maxSpeed, velocity[0], velocity[1])
        self._buf = (C.c_float * 8)()

    def __del__(self):
        if getattr(self, "_s", None):
            self._L.rvo_destroy(self._s)
            self._s = None

    def addAgent(self, pos, neighborDist, maxNeighbors, timeHorizon, timeHorizonObst, radius, maxSpeed,
                 velocity=(0, 0)):
        return self._L.rvo_add_agent(self._s, pos[0], pos[1], neighborDist, int(maxNeighbors), timeHorizon,
                                     timeHorizonObst, radius, maxSpeed, velocity[0], velocity[1])

    def addObstacle(self, vertices):
        n = len(vertices)
        arr = (C.c_float * (2 * n))(*[c for v in vertices for c in v])
        return self._L.rvo_add_obstacle(self._s, arr, n)

    def processObstacles(self):
        self._L.rvo_process_obstacles(self._s)

    def doStep(self):
        self._L.rvo_do_step(self._s)

    def setAgentPosition(self, i, p):
        self._L.rvo_set_agent_position(self._s, i, p[0], p[1])

    def setAgentVelocity(self, i, v):
        self._L.rvo_set_agent_velocity(self._s, i, v[0], v[1])

    def setAgentPrefVelocity(self, i, v):
        self._L.rvo_set_agent_pref_velocity(self._s, i, v[0], v[1])

    def getAgentPosition(self, i):
        self._L.rvo_get_agent_position(self._s, i, self._buf)
        return (self._buf[0], self._buf[1])

    def getAgentVelocity(self, i):
        self._L.rvo_get_agent_velocity(self._s, i, self._buf)
        return (self._buf[0], self._buf[1])

    def getAgentPrefVelocity(self, i):
        self._L.rvo_get_agent_pref_velocity(self._s, i, self._buf)
        return (self._buf[0], self._buf[1])

    def getAgentMaxSpeed(self, i):
        return self._L.rvo_get_agent_max_speed(self._s, i)

    def getNumAgents(self):
        return self._L.rvo_get_num_agents(self._s)

    def getGlobalTime(self):
        return self._L.rvo_get_global_time(self._s)

    # introspection beyond Python-RVO2 (used by the parity tests only)
    def getAgentNumAgentNeighbors(self, i):
        return self._L.rvo_get_agent_num_agent_neighbors(self._s, i)

    def getAgentAgentNeighbor(self, i, k):
        return self._L.rvo_get_agent_agent_neighbor(self._s, i, k)

    def getAgentNumORCALines(self, i):
        return self._L.rvo_get_agent_num_orca_lines(self._s, i)

    def getAgentORCALine(self, i, k):
        self._L.rvo_get_agent_orca_line(self._s, i, k, self._buf)
        return tuple(self._buf[j] for j in range(4))
