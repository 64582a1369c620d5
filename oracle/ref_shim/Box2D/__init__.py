"""Stand-in for box2d-py ~=2.3.5 (setup.py:14) -- TEST INFRASTRUCTURE ONLY (oracle side).

Just enough of the pybox2d API for the reference's car_racing/car_dynamics.py and
car_racing_multi_players.py to run unmodified.  The rigid-body arithmetic is NOT Box2D: it is the
restatement in oracle/car_oracle.c (b2Island::Solve with revolute joints, polygon mass data, the
sensor-overlap test), called through ctypes.  So running the reference on top of this module pins
the reference's PYTHON logic (Car.step, CarRacing.step/reset/_create_track, FrictionDetector)
against the C/CUDA restatements, while Box2D's own numerics stay 'restated from memory, unpinned'.

Restated pybox2d/Box2D behaviour: bodies and joints are prepended to the world lists and islands
are built by depth-first search from awake bodies (b2World::Solve), which fixes the order in which
the joints of a car are relaxed; ApplyForceToCenter / joint.motorSpeed wake bodies; forces are
cleared after every Step; sensor contacts fire Begin/EndContact inside Step before the solve.
Collisions between non-sensor fixtures exist only between the two cars of a two-car world; they go
through car_oracle.c's contact restatement (car_oracle_world_step_two), which owns the manifolds.
"""
import ctypes
import math
import os
import sys

import numpy as np

_ORACLE_DIR = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ORACLE_DIR not in sys.path:
    sys.path.insert(0, _ORACLE_DIR)
import car_oracle as _co  # noqa: E402

_L = _co.lib()
_f32 = np.float32


class _V2(ctypes.Structure):
    _fields_ = [("x", ctypes.c_float), ("y", ctypes.c_float)]


class _CBody(ctypes.Structure):
    _fields_ = [("p", _V2), ("qs", ctypes.c_float), ("qc", ctypes.c_float), ("local_center", _V2), ("c0", _V2),
                ("c", _V2), ("a0", ctypes.c_float), ("a", ctypes.c_float), ("v", _V2), ("w", ctypes.c_float),
                ("force", _V2), ("torque", ctypes.c_float), ("mass", ctypes.c_float), ("inv_mass", ctypes.c_float),
                ("I", ctypes.c_float), ("inv_I", ctypes.c_float), ("awake", ctypes.c_int),
                ("sleep_time", ctypes.c_float)]


class _CJoint(ctypes.Structure):
    _fields_ = [("a", ctypes.c_int), ("b", ctypes.c_int), ("local_anchor_a", _V2), ("local_anchor_b", _V2),
                ("reference_angle", ctypes.c_float), ("enable_motor", ctypes.c_int), ("enable_limit", ctypes.c_int),
                ("max_motor_torque", ctypes.c_float), ("motor_speed", ctypes.c_float), ("lower", ctypes.c_float),
                ("upper", ctypes.c_float), ("impulse", ctypes.c_float * 3), ("motor_impulse", ctypes.c_float),
                ("limit_state", ctypes.c_int), ("rA", _V2), ("rB", _V2), ("lcA", _V2), ("lcB", _V2),
                ("mA", ctypes.c_float), ("mB", ctypes.c_float), ("iA", ctypes.c_float), ("iB", ctypes.c_float),
                ("K", (ctypes.c_float * 3) * 3), ("motor_mass", ctypes.c_float)]


assert ctypes.sizeof(_CBody) == _L.car_oracle_sizeof_body(), (ctypes.sizeof(_CBody), _L.car_oracle_sizeof_body())
assert ctypes.sizeof(_CJoint) == _L.car_oracle_sizeof_joint(), (ctypes.sizeof(_CJoint), _L.car_oracle_sizeof_joint())
_L.car_oracle_body_init.argtypes = [ctypes.POINTER(_CBody), ctypes.c_float, ctypes.c_float, ctypes.c_float]
_L.car_oracle_body_set_mass.argtypes = [ctypes.POINTER(_CBody), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.c_int]
_L.car_oracle_island_solve.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_float,
                                       ctypes.c_float, ctypes.c_int, ctypes.c_int]
_L.car_oracle_world_step_two.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_void_p, ctypes.c_float, ctypes.c_float, ctypes.c_int, ctypes.c_int]
_L.car_oracle_contact_count.argtypes = [ctypes.c_void_p, ctypes.c_void_p]


class b2Vec2(object):
    __slots__ = ("x", "y")

    def __init__(self, x=0.0, y=None):
        if y is None:
            x, y = x
        self.x, self.y = float(_f32(x)), float(_f32(y))

    def __getitem__(self, i):
        return (self.x, self.y)[i]

    def __iter__(self):
        return iter((self.x, self.y))

    def __len__(self):
        return 2

    def __add__(self, o):
        return b2Vec2(self.x + o[0], self.y + o[1])

    __radd__ = __add__

    def __sub__(self, o):
        return b2Vec2(self.x - o[0], self.y - o[1])

    def __rsub__(self, o):
        return b2Vec2(o[0] - self.x, o[1] - self.y)

    def __mul__(self, s):
        return b2Vec2(self.x * s, self.y * s)

    __rmul__ = __mul__

    def __neg__(self):
        return b2Vec2(-self.x, -self.y)

    def __repr__(self):
        return "b2Vec2(%g,%g)" % (self.x, self.y)


class b2Transform(object):
    def __init__(self):
        self.position = b2Vec2(0, 0)
        self._angle = 0.0

    @property
    def angle(self):
        return self._angle

    @angle.setter
    def angle(self, a):
        self._angle = float(_f32(a))

    def __mul__(self, v):   # b2Mul(xf, v) in float32
        s, c = _f32(math.sin(_f32(self._angle))), _f32(math.cos(_f32(self._angle)))
        x, y = _f32(v[0]), _f32(v[1])
        px, py = _f32(self.position[0]), _f32(self.position[1])
        return b2Vec2(_f32(_f32(c * x) - _f32(s * y)) + px, _f32(_f32(s * x) + _f32(c * y)) + py)


class polygonShape(object):
    def __init__(self, vertices=None, box=None):
        self.vertices = list(vertices) if vertices is not None else []


class fixtureDef(object):
    def __init__(self, shape=None, density=0.0, friction=0.2, restitution=0.0, categoryBits=0x0001, maskBits=0xFFFF,
                 isSensor=False):
        self.shape, self.density, self.friction, self.restitution = shape, density, friction, restitution
        self.categoryBits, self.maskBits, self.isSensor = categoryBits, maskBits, isSensor


class revoluteJointDef(object):
    def __init__(self, **kw):
        self.referenceAngle = 0.0
        self.__dict__.update(kw)


class contactListener(object):
    def __init__(self):
        pass

    def BeginContact(self, contact):
        pass

    def EndContact(self, contact):
        pass


class _Fixture(object):
    def __init__(self, body, fd):
        raw = np.array([[float(_f32(x)), float(_f32(y))] for x, y in fd.shape.vertices], np.float32)
        hull = np.zeros((8, 2), np.float32)
        n = _L.car_oracle_convex_hull(raw.ctypes.data_as(ctypes.c_void_p), len(raw), hull.ctypes.data_as(ctypes.c_void_p))
        self.body = body
        self.shape = polygonShape(vertices=[(float(hull[i, 0]), float(hull[i, 1])) for i in range(n)])
        self._hull = hull[:n].copy()
        self.density, self.sensor = float(fd.density), bool(fd.isSensor)
        self.categoryBits, self.maskBits = fd.categoryBits, fd.maskBits

    def world_poly(self):
        b = self.body
        s, c = (b._c.qs, b._c.qc) if b._c is not None else (0.0, 1.0)
        px, py = b.position
        out = np.empty_like(self._hull)
        out[:, 0] = _f32(c) * self._hull[:, 0] - _f32(s) * self._hull[:, 1] + _f32(px)
        out[:, 1] = _f32(s) * self._hull[:, 0] + _f32(c) * self._hull[:, 1] + _f32(py)
        return np.ascontiguousarray(out, np.float32)


class _Body(object):
    def __init__(self, world, dynamic, position, angle, fixtures):
        self._world, self._dynamic = world, dynamic
        self.userData = None
        self._joints = []       # joint edges, most recent first
        self._island = False
        self._c = None
        if dynamic:
            self._c = _CBody()
            _L.car_oracle_body_init(ctypes.byref(self._c), position[0], position[1], angle)
        else:
            self._pos = b2Vec2(position)
        fds = fixtures if isinstance(fixtures, (list, tuple)) else [fixtures]
        self.fixtures = [_Fixture(self, fd) for fd in fds]
        if dynamic:
            polys = np.concatenate([f._hull for f in self.fixtures]).astype(np.float32)
            counts = np.array([len(f._hull) for f in self.fixtures], np.int32)
            dens = np.array([f.density for f in self.fixtures], np.float32)
            _L.car_oracle_body_set_mass(ctypes.byref(self._c), polys.ctypes.data_as(ctypes.c_void_p),
                                        counts.ctypes.data_as(ctypes.c_void_p), dens.ctypes.data_as(ctypes.c_void_p),
                                        len(counts))
        else:
            self._aabbs = None

    # ---- pybox2d surface used by the reference ----
    @property
    def position(self):
        return b2Vec2(self._c.p.x, self._c.p.y) if self._dynamic else self._pos

    @property
    def angle(self):
        return float(self._c.a) if self._dynamic else 0.0

    @property
    def linearVelocity(self):
        return b2Vec2(self._c.v.x, self._c.v.y) if self._dynamic else b2Vec2(0, 0)

    @property
    def angularVelocity(self):
        return float(self._c.w) if self._dynamic else 0.0

    @property
    def transform(self):
        t = b2Transform()
        t.position = self.position
        t.angle = self.angle
        return t

    @property
    def awake(self):
        return bool(self._c.awake) if self._dynamic else False

    def _set_awake(self, flag):
        c = self._c
        if flag:
            if not c.awake:
                c.awake, c.sleep_time = 1, 0.0
        else:
            c.awake, c.sleep_time = 0, 0.0
            c.v.x = c.v.y = c.w = c.force.x = c.force.y = c.torque = 0.0

    def GetWorldVector(self, v):
        s, c = _f32(self._c.qs), _f32(self._c.qc)
        x, y = _f32(v[0]), _f32(v[1])
        return b2Vec2(_f32(c * x) - _f32(s * y), _f32(s * x) + _f32(c * y))

    def ApplyForceToCenter(self, force, wake):
        if not self._dynamic:
            return
        if wake and not self._c.awake:
            self._set_awake(True)
        if self._c.awake:
            self._c.force.x = float(_f32(self._c.force.x) + _f32(force[0]))
            self._c.force.y = float(_f32(self._c.force.y) + _f32(force[1]))


class _Joint(object):
    def __init__(self, jd):
        self.bodyA, self.bodyB = jd.bodyA, jd.bodyB
        c = self._c = _CJoint()
        c.local_anchor_a.x, c.local_anchor_a.y = jd.localAnchorA
        c.local_anchor_b.x, c.local_anchor_b.y = jd.localAnchorB
        c.reference_angle = jd.referenceAngle
        c.enable_motor, c.enable_limit = int(jd.enableMotor), int(jd.enableLimit)
        c.max_motor_torque, c.motor_speed = jd.maxMotorTorque, jd.motorSpeed
        c.lower, c.upper = jd.lowerAngle, jd.upperAngle
        self._island = False

    @property
    def angle(self):   # b2RevoluteJoint::GetJointAngle
        return float(_f32(_f32(self.bodyB._c.a) - _f32(self.bodyA._c.a)) - _f32(self._c.reference_angle))

    @property
    def motorSpeed(self):
        return float(self._c.motor_speed)

    @motorSpeed.setter
    def motorSpeed(self, v):   # SetMotorSpeed wakes both bodies
        self.bodyA._set_awake(True)
        self.bodyB._set_awake(True)
        self._c.motor_speed = v


class _Contact(object):
    def __init__(self, fa, fb):
        self.fixtureA, self.fixtureB = fa, fb


class b2World(object):
    def __init__(self, gravity=(0, 0), contactListener=None, doSleep=True):
        self._listener = contactListener
        self.bodies = []     # head = most recently created (b2World::CreateBody prepends)
        self.joints = []
        self._touching = {}
        self._inv_dt0 = 0.0
        self._contact_store = ctypes.create_string_buffer(_L.car_oracle_sizeof_contact_store())

    def CreateDynamicBody(self, position=(0, 0), angle=0.0, fixtures=None, **kw):
        b = _Body(self, True, position, angle, fixtures)
        self.bodies.insert(0, b)
        return b

    def CreateStaticBody(self, position=(0, 0), fixtures=None, **kw):
        b = _Body(self, False, position, 0.0, fixtures)
        self.bodies.insert(0, b)
        return b

    def CreateJoint(self, jd):
        j = _Joint(jd)
        self.joints.insert(0, j)
        j.bodyA._joints.insert(0, (j, j.bodyB))
        j.bodyB._joints.insert(0, (j, j.bodyA))
        return j

    def DestroyBody(self, b):
        if b in self.bodies:
            self.bodies.remove(b)
        for j, _ in list(b._joints):
            if j in self.joints:
                self.joints.remove(j)

    # ---- b2ContactManager::Collide for (dynamic non-hull fixture, static sensor fixture) pairs ----
    def _collide(self):
        statics = [b for b in self.bodies if not b._dynamic]
        if not statics:
            return
        if getattr(self, "_static_cache", None) is None or self._static_cache[0] != len(statics):
            polys = [b.fixtures[0].world_poly() for b in statics]
            aabb = np.array([[p[:, 0].min(), p[:, 1].min(), p[:, 0].max(), p[:, 1].max()] for p in polys], np.float32)
            self._static_cache = (len(statics), statics, polys, aabb)
        _, statics, polys, aabb = self._static_cache
        m = np.float32(0.02)
        for b in self.bodies:
            if not b._dynamic:
                continue
            for f in b.fixtures:
                if f.categoryBits != 0x0020:     # wheels only: hull-tile callbacks have no effect in the reference
                    continue
                wp = f.world_poly()
                lo, hi = wp.min(axis=0), wp.max(axis=0)
                near = np.nonzero(~((aabb[:, 0] > hi[0] + m) | (aabb[:, 2] < lo[0] - m) | (aabb[:, 1] > hi[1] + m) |
                                    (aabb[:, 3] < lo[1] - m)))[0]
                now = set()
                for t in near:
                    if _L.car_oracle_polys_touch(wp.ctypes.data_as(ctypes.c_void_p), len(wp),
                                                 polys[t].ctypes.data_as(ctypes.c_void_p), len(polys[t])):
                        now.add(int(t))
                was = self._touching.get(f, set())
                # ascending tile creation index == descending block_id; the C/CUDA restatements iterate
                # block ids ascending, so do the same here
                for t in sorted(now - was, key=lambda i: statics[i].block_id if hasattr(statics[i], "block_id") else i):
                    self._listener.BeginContact(_Contact(statics[t].fixtures[0], f))
                for t in sorted(was - now, key=lambda i: statics[i].block_id if hasattr(statics[i], "block_id") else i):
                    self._listener.EndContact(_Contact(statics[t].fixtures[0], f))
                self._touching[f] = now

    def _joint_groups(self):
        """Bodies and joints of every joint-connected group in b2World::Solve's depth-first order
        (seeds in body-list order, joint edges most recent first), regardless of the awake state."""
        seen_b, seen_j, groups = set(), set(), []
        for seed in self.bodies:
            if not seed._dynamic or id(seed) in seen_b:
                continue
            gb, gj, stack = [], [], [seed]
            seen_b.add(id(seed))
            while stack:
                b = stack.pop()
                gb.append(b)
                for j, other in b._joints:
                    if id(j) in seen_j:
                        continue
                    gj.append(j)
                    seen_j.add(id(j))
                    if id(other) in seen_b:
                        continue
                    stack.append(other)
                    seen_b.add(id(other))
            groups.append((gb, gj))
        return groups

    def _step_two_cars(self, groups, dt, dt_ratio, vi, pi):
        """Two cars: collide + solve in C (car 0 = created first = LAST group of the body list)."""
        (b1, j1), (b0, j0) = groups
        arrs = []
        for gb, gj in ((b0, j0), (b1, j1)):
            cb = (_CBody * 5)(*[b._c for b in gb])
            idx = {id(b): i for i, b in enumerate(gb)}
            for j in gj:
                j._c.a, j._c.b = idx[id(j.bodyA)], idx[id(j.bodyB)]
            cj = (_CJoint * 4)(*[j._c for j in gj])
            arrs.append((cb, cj))
        _L.car_oracle_world_step_two(ctypes.cast(arrs[0][0], ctypes.c_void_p), ctypes.cast(arrs[0][1], ctypes.c_void_p),
                                     ctypes.cast(arrs[1][0], ctypes.c_void_p), ctypes.cast(arrs[1][1], ctypes.c_void_p),
                                     ctypes.cast(self._contact_store, ctypes.c_void_p), dt, dt_ratio, vi, pi)
        for (gb, gj), (cb, cj) in zip(((b0, j0), (b1, j1)), arrs):
            for i, b in enumerate(gb):
                ctypes.memmove(ctypes.byref(b._c), ctypes.byref(cb[i]), ctypes.sizeof(_CBody))
            for i, j in enumerate(gj):
                ctypes.memmove(ctypes.byref(j._c), ctypes.byref(cj[i]), ctypes.sizeof(_CJoint))

    def contact_count(self):
        return int(_L.car_oracle_contact_count(ctypes.cast(self._contact_store, ctypes.c_void_p), None))

    def _step_islands(self, dt, dt_ratio, velocity_iterations, position_iterations):
        for seed in self.bodies:                    # b2World::Solve
            if seed._island or not seed._dynamic or not seed.awake:
                continue
            island_b, island_j, stack = [], [], [seed]
            seed._island = True
            while stack:
                b = stack.pop()
                island_b.append(b)
                b._set_awake(True)
                for j, other in b._joints:
                    if j._island:
                        continue
                    island_j.append(j)
                    j._island = True
                    if other._island:
                        continue
                    stack.append(other)
                    other._island = True
            nb, nj = len(island_b), len(island_j)
            cb = (_CBody * nb)(*[b._c for b in island_b])
            idx = {id(b): i for i, b in enumerate(island_b)}
            for j in island_j:
                j._c.a, j._c.b = idx[id(j.bodyA)], idx[id(j.bodyB)]
            cj = (_CJoint * nj)(*[j._c for j in island_j])
            _L.car_oracle_island_solve(ctypes.cast(cb, ctypes.c_void_p), nb, ctypes.cast(cj, ctypes.c_void_p), nj, dt,
                                       dt_ratio, velocity_iterations, position_iterations)
            for i, b in enumerate(island_b):
                ctypes.memmove(ctypes.byref(b._c), ctypes.byref(cb[i]), ctypes.sizeof(_CBody))
            for i, j in enumerate(island_j):
                ctypes.memmove(ctypes.byref(j._c), ctypes.byref(cj[i]), ctypes.sizeof(_CJoint))

    def Step(self, dt, velocity_iterations, position_iterations):
        dt = float(_f32(dt))
        if self._listener is not None:
            self._collide()
        dt_ratio = float(_f32(self._inv_dt0) * _f32(dt))
        for b in self.bodies:
            b._island = False
        for j in self.joints:
            j._island = False
        groups = self._joint_groups()
        if len(groups) == 2 and all(len(gb) == 5 and len(gj) == 4 for gb, gj in groups):
            self._step_two_cars(groups, dt, dt_ratio, velocity_iterations, position_iterations)
        else:
            assert len([b for b in self.bodies if b._dynamic]) <= 5, "contacts are restated for two-car worlds only"
            self._step_islands(dt, dt_ratio, velocity_iterations, position_iterations)
        for b in self.bodies:                       # ClearForces
            if b._dynamic:
                b._c.force.x = b._c.force.y = b._c.torque = 0.0
        self._inv_dt0 = 1.0 / dt if dt > 0 else 0.0
