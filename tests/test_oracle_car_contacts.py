"""Car-car contacts of the C oracle (oracle/car_oracle.c, "mini Box2D, part 2": b2CollidePolygons, contact solver,
merged island) -- known answers worked out by hand for axis-aligned boxes, geometric properties of the manifold on
random poses, conservation laws of the solver in free flight, and the fixture recorded by running the reference's
own CarRacing.step on the stand-in Box2D with two cars steered into each other."""
import numpy as np
import pytest

from conftest import load_golden
import car_oracle as C

HW, HR = np.float32(14 * 0.02), np.float32(27 * 0.02)      # wheel box half extents (car_dynamics.py:20-21, 84-87)
RADIUS = 0.01                                              # b2_polygonRadius


def _inside(poly, pt, margin):
    """pt within `margin` of the convex CCW polygon."""
    n = len(poly)
    for i in range(n):
        e = poly[(i + 1) % n] - poly[i]
        nrm = np.array([e[1], -e[0]]) / np.hypot(*e)
        if np.dot(nrm, pt - poly[i]) > margin:
            return False
    return True


def _world(poly, pose):
    c, s = np.cos(pose[2]), np.sin(pose[2])
    return poly @ np.array([[c, s], [-s, c]]) + np.array(pose[:2])


def test_wheel_box_hull_order_and_known_manifold():
    box = C.fixture_polygon(4)
    assert np.array_equal(box, np.array([[HW, -HR], [HW, HR], [-HW, HR], [-HW, -HR]], np.float32))
    # two wheel boxes side by side, gap 0.01 (< 2 * polygonRadius): reference face = A's +x face (edge 0),
    # incident edge = B's -x face (edge 2: vertices 2 -> 3), two points
    gap = 0.01
    m = C.collide_fixtures(4, (0, 0, 0), 4, (2 * float(HW) + gap, 0, 0))
    assert m["count"] == 2 and m["type"] == 0
    assert np.allclose(m["normal"], [1, 0], atol=1e-7)
    # world-manifold points sit halfway between the two skins: x = HW + gap / 2, y = +-HR
    assert np.allclose(sorted(m["points"][:, 1]), [-HR, HR], atol=1e-6)
    assert np.allclose(m["points"][:, 0], float(HW) + gap / 2, atol=1e-6)
    # feature ids: face 0 of A against vertices 2 and 3 of B (typeA = face = 1, typeB = vertex = 0)
    assert sorted(m["ids"].tolist()) == sorted([0 | 2 << 8 | 1 << 16, 0 | 3 << 8 | 1 << 16])
    # separated by more than the two skins: no manifold
    assert C.collide_fixtures(4, (0, 0, 0), 4, (2 * float(HW) + 0.0201, 0, 0))["count"] == 0
    # B shifted up by HR: only half of the faces face each other -> one clipped point (vertex-face id) + one vertex
    m = C.collide_fixtures(4, (0, 0, 0), 4, (2 * float(HW) + gap, float(HR), 0))
    assert m["count"] == 2
    assert np.allclose(sorted(m["points"][:, 1]), [0.0, HR + 0.02], atol=1e-5) or np.allclose(sorted(m["points"][:, 1]), [0.0, HR], atol=0.021)


def test_manifold_geometry_on_random_poses():
    rng = np.random.default_rng(0)
    polys = [C.fixture_polygon(f).astype(np.float64) for f in range(5)]
    hits = 0
    for _ in range(4000):
        fa, fb = rng.integers(0, 5, 2)
        pa = (0.0, 0.0, rng.uniform(-np.pi, np.pi))
        d = rng.uniform(0.2, 3.2)
        th = rng.uniform(-np.pi, np.pi)
        pb = (d * np.cos(th), d * np.sin(th), rng.uniform(-np.pi, np.pi))
        m = C.collide_fixtures(int(fa), pa, int(fb), pb)
        wa, wb = _world(polys[fa], pa), _world(polys[fb], pb)
        # brute-force separating-axis distance (same measure the sensor test uses)
        def max_sep(A, B):
            best = -1e9
            for i in range(len(A)):
                e = A[(i + 1) % len(A)] - A[i]
                nrm = np.array([e[1], -e[0]]) / np.hypot(*e)
                best = max(best, min(np.dot(nrm, q - A[i]) for q in B))
            return best
        sep = max(max_sep(wa, wb), max_sep(wb, wa))
        if sep > 2 * RADIUS + 1e-5:
            assert m["count"] == 0
            continue
        if m["count"] == 0:
            continue            # clipped away (corner-corner near misses): allowed by the algorithm
        hits += 1
        assert m["count"] in (1, 2) and abs(np.hypot(*m["normal"]) - 1) < 1e-5
        if sep > -0.05:         # shallow contact: the normal points from A to B, the points lie on both skinned polygons
            ca, cb = wa.mean(0), wb.mean(0)
            assert np.dot(m["normal"], cb - ca) > 0
            for pt in m["points"]:
                assert _inside(wa, pt, 2 * RADIUS + abs(sep) + 1e-4) and _inside(wb, pt, 2 * RADIUS + abs(sep) + 1e-4)
    assert hits > 300


@pytest.mark.parametrize("case", ["head_on", "t_bone", "glancing", "rear_end"])
def test_free_flight_collision_conserves_momentum(case):
    pose, vel = {
        "head_on": ([(0, 0, 0), (0.3, 7.0, np.pi)], [(0, 6), (0, -6)]),
        "t_bone": ([(0, 0, 0), (5.0, 0.5, np.pi / 2)], [(0, 0), (-8, 0)]),
        "glancing": ([(0, 0, 0), (2.2, 6.0, np.pi + 0.1)], [(0, 5), (0, -5)]),
        "rear_end": ([(0, 0, 0), (0.2, -6.5, 0.05)], [(0, 2), (0, 9)]),
    }[case]
    T = 80
    out = C.free_collision(np.array(pose, np.float64), np.array(vel, np.float64), T)
    assert (out[:, 3] > 0).any(), "the cars never touched"
    p0 = out[0, :2]
    # internal impulses only: the linear momentum of the 10 bodies is conserved by the velocity solver (fp32);
    # angular momentum about the origin additionally drifts a little because the position pass (Baumgarte
    # pseudo-impulses on contacts and joints) moves centres without touching velocities
    scale = max(np.abs(p0).max(), 50.0)
    assert np.abs(out[:, :2] - p0).max() <= 2e-4 * scale
    assert np.abs(out[:, 2] - out[0, 2]).max() <= 1e-2 * max(abs(out[0, 2]), 200.0)
    # the hulls never pass through each other: the centres stay farther apart than the narrowest hull part allows
    dist = np.hypot(out[:, 5] - out[:, 7], out[:, 6] - out[:, 8])
    assert dist.min() > 1.0
    # restitution 0 and friction: kinetic energy cannot be checked here, but the cars must separate or rest,
    # i.e. contacts do not persist with growing penetration -- manifold points stay <= 2 per contact
    assert (out[:, 4] <= 2 * out[:, 3]).all()


def test_collision_rollout_matches_reference():
    """The reference's own CarRacing.step (two cars steered into each other) on the stand-in Box2D vs the C env."""
    g = load_golden("car_double_collision")
    track, border = C.create_track(g["draws"])
    env = C.CarOracleEnv(2, 1, None, render=False)
    env.reset(track, border, g["birth"])
    assert np.array_equal(env.get_state(), g["state0"])
    assert (g["contacts"] > 0).sum() >= 20
    for t, a in enumerate(g["actions"]):
        _, rew, done, _ = env.step(a)
        assert np.array_equal(env.get_state(), g["states"][t]), t
        assert env.contacts()[0] == g["contacts"][t], t
        assert np.array_equal(rew, g["rewards"][t]) and np.array_equal(done, g["dones"][t]), t


def test_cars_do_not_pass_through_each_other():
    rng = np.random.RandomState(3)
    track, border, _ = C.make_track(rng)
    env = C.CarOracleEnv(2, 1, None, render=False)
    env.reset(track, border, [0, 1])
    dmin, touched = 1e9, 0
    for t in range(120):
        env.step(np.array([[-0.35, 0.5], [0.35, 0.5]]))
        s = env.get_state()
        dmin = min(dmin, np.hypot(*(s[0, :2] - s[1, :2])))
        touched += env.contacts()[0] > 0
    assert touched > 20
    assert dmin > 2.0          # hull half-widths 1.2 + 1.2: side by side is as close as they get
