/*
 * car_oracle.c -- CPU restatement of the reference's cCarRacing stepping path.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/pong_oracle.c for the rules).
 *
 * What is restated from WHICH source (paths relative to /root/reference/competitive_rl/):
 *   car_racing/car_dynamics.py            Car.__init__ (55-129), gas/brake/steer (131-157), step (159-234)
 *   car_racing/car_racing_multi_players.py  FrictionDetector._contact (111-153), _create_track (262-452),
 *                                         reset (454-525), process_action (527-540), step (542-620),
 *                                         get_observation (622-634), render_indicators_for_pygame (645-670),
 *                                         render_road_for_observation_map (732-755), camera_view (764-789),
 *                                         camera_update (791-812), render "internal_rgb_array" (857-863)
 *   car_racing/pygame_rendering.py        vertical_ind / horiz_ind / draw_text (8-18)
 * and, because the reference's native arithmetic lives in box2d-py ~=2.3.5 (setup.py:14), which is
 * NOT in /root/reference and not installable here, the published Box2D 2.3 algorithms restated
 * from memory: b2PolygonShape::ComputeMass, b2Body::ResetMassData, b2World::Step / b2Island::Solve
 * (force integration, warm-started sequential impulses, 180 velocity + 60 position iterations,
 * translation/rotation clamps, island sleeping), b2RevoluteJoint (motor + limit + point),
 * sensor overlap for tile contacts.  PARITY UNPINNED for that part: no Box2D run is available to
 * check it against.  The Python-level logic IS pinned: oracle/ref_car_loader.py runs
 * the reference's own car_dynamics.py / car_racing_multi_players.py on top of a Box2D stand-in
 * that calls this file's solver, and tests/test_oracle_car.py compares.
 *
 * Deliberate simplifications (stated, tested as such):
 *   - car-car contacts (cCarRacingDouble) are evaluated for every allowed fixture pair each Step and
 *     enter the island in a canonical order (see "mini Box2D, part 2" below);
 *   - a sensor contact exists exactly while the wheel polygon and the tile polygon are closer
 *     than 2*b2_polygonRadius by the separating-axis measure, evaluated at the start of
 *     world.Step (Box2D: b2TestOverlap on contacts whose fat AABBs overlap);
 *   - the renderer evaluates the reference's paint-crop-rotate-blit pipeline per destination pixel
 *     with pygame 1.9's polygon fill / rotate rules restated from memory (see render_obs below).
 *
 * Build: gcc -O2 -ffp-contract=off -pthread -shared -fPIC (oracle/Makefile).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ */
/* constants                                                                                   */

#define STATE_W 96
#define STATE_H 96
#define SCALE 6.0
#define TRACK_RAD (900.0 / SCALE)
#define PLAYFIELD (2000.0 / SCALE)
#define FPS 50
#define TRACK_DETAIL_STEP (21.0 / SCALE)
#define TRACK_TURN_RATE 0.31
#define TRACK_WIDTH (40.0 / SCALE)
#define BORDER (8.0 / SCALE)
#define BORDER_MIN_COUNT 4
#define CHECKPOINTS 12
#define MAX_TRACK 512
#define MAX_TRACK_RAW 2600
#define MAX_CARS 2
#define OBS_SCALE ((10 / (100 / sqrt(96.0))) * 1.8)   /* CarRacing.obs_scale, :211 */

#define SIZE 0.02
#define ENGINE_POWER (100000000 * SIZE * SIZE)
#define WHEEL_MOMENT_OF_INERTIA (4000 * SIZE * SIZE)
#define FRICTION_LIMIT (1000000 * SIZE * SIZE)
#define WHEEL_R 27
#define WHEEL_W 14

/* Box2D 2.3 b2Settings.h */
#define B2_LINEAR_SLOP 0.005f
#define B2_ANGULAR_SLOP (2.0f / 180.0f * 3.14159265359f)
#define B2_POLYGON_RADIUS (2.0f * B2_LINEAR_SLOP)
#define B2_MAX_ANGULAR_CORRECTION (8.0f / 180.0f * 3.14159265359f)
#define B2_MAX_TRANSLATION 2.0f
#define B2_MAX_ROTATION (0.5f * 3.14159265359f)
#define B2_TIME_TO_SLEEP 0.5f
#define B2_LINEAR_SLEEP_TOL 0.01f
#define B2_ANGULAR_SLEEP_TOL (2.0f / 180.0f * 3.14159265359f)

typedef struct { float x, y; } V2;
typedef struct { float s, c; } Rot;

static V2 v2(float x, float y) { V2 r = {x, y}; return r; }
static V2 vadd(V2 a, V2 b) { return v2(a.x + b.x, a.y + b.y); }
static V2 vsub(V2 a, V2 b) { return v2(a.x - b.x, a.y - b.y); }
static V2 vscale(float s, V2 a) { return v2(s * a.x, s * a.y); }
static float vdot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
static float vcross(V2 a, V2 b) { return a.x * b.y - a.y * b.x; }
static V2 cross_sv(float s, V2 a) { return v2(-s * a.y, s * a.x); }
static Rot rot(float a) { Rot r = {sinf(a), cosf(a)}; return r; }
static V2 rmul(Rot q, V2 v) { return v2(q.c * v.x - q.s * v.y, q.s * v.x + q.c * v.y); }

/* ------------------------------------------------------------------------------------------ */
/* mini Box2D: bodies, polygon mass, revolute joints                                           */

typedef struct {
    V2 p; Rot q;                 /* m_xf */
    V2 local_center, c0, c;      /* m_sweep */
    float a0, a;
    V2 v; float w;
    V2 force; float torque;
    float mass, inv_mass, I, inv_I;
    int awake; float sleep_time;
} Body;

typedef struct {
    int a, b;                    /* body indices */
    V2 local_anchor_a, local_anchor_b;
    float reference_angle;
    int enable_motor, enable_limit;
    float max_motor_torque, motor_speed, lower, upper;
    /* solver state (b2RevoluteJoint members) */
    float impulse[3], motor_impulse;
    int limit_state;             /* 0 inactive, 1 atLower, 2 atUpper, 3 equal */
    V2 rA, rB, lcA, lcB;
    float mA, mB, iA, iB;
    float K[3][3];               /* K[col][row] like b2Mat33 ex, ey, ez */
    float motor_mass;
} RevJoint;

/* b2PolygonShape::ComputeMass (2.3) for a convex polygon given in CCW order */
static void polygon_mass(const V2* vs, int n, float density, float* mass, V2* center, float* I) {
    V2 c = v2(0.f, 0.f), s = v2(0.f, 0.f);
    float area = 0.f, inertia = 0.f;
    const float inv3 = 1.0f / 3.0f;
    for (int i = 0; i < n; ++i) s = vadd(s, vs[i]);
    s = vscale(1.0f / n, s);
    for (int i = 0; i < n; ++i) {
        V2 e1 = vsub(vs[i], s), e2 = vsub(vs[(i + 1) % n], s);
        float D = vcross(e1, e2);
        float tri = 0.5f * D;
        area += tri;
        c = vadd(c, vscale(tri * inv3, vadd(e1, e2)));
        float intx2 = e1.x * e1.x + e2.x * e1.x + e2.x * e2.x;
        float inty2 = e1.y * e1.y + e2.y * e1.y + e2.y * e2.y;
        inertia += (0.25f * inv3 * D) * (intx2 + inty2);
    }
    *mass = density * area;
    c = vscale(1.0f / area, c);
    *center = vadd(c, s);
    *I = density * inertia;
    *I += *mass * (vdot(*center, *center) - vdot(c, c));
}

/* b2Body::ResetMassData over a list of polygons attached to the body */
static void body_set_mass(Body* b, const V2* polys, const int* counts, const float* density, int n_polys) {
    float mass = 0.f, I = 0.f;
    V2 lc = v2(0.f, 0.f);
    const V2* p = polys;
    for (int k = 0; k < n_polys; ++k) {
        float m, i;
        V2 c;
        polygon_mass(p, counts[k], density[k], &m, &c, &i);
        mass += m;
        lc = vadd(lc, vscale(m, c));
        I += i;
        p += counts[k];
    }
    b->mass = mass;
    b->inv_mass = 1.0f / mass;
    lc = vscale(b->inv_mass, lc);
    I -= mass * vdot(lc, lc);
    b->I = I;
    b->inv_I = 1.0f / I;
    b->local_center = lc;
    b->c = b->c0 = vadd(rmul(b->q, lc), b->p);
}

static void body_init(Body* b, float x, float y, float angle) {
    memset(b, 0, sizeof *b);
    b->p = v2(x, y);
    b->q = rot(angle);
    b->a = b->a0 = angle;
    b->awake = 1;
}

static void body_set_awake(Body* b, int flag) {
    if (flag) {
        if (!b->awake) { b->awake = 1; b->sleep_time = 0.f; }
    } else {
        b->awake = 0; b->sleep_time = 0.f;
        b->v = v2(0.f, 0.f); b->w = 0.f; b->force = v2(0.f, 0.f); b->torque = 0.f;
    }
}

static float clampf(float a, float lo, float hi) { return a < lo ? lo : (a > hi ? hi : a); }

/* b2Mat33::Solve33 / Solve22 with K stored as columns ex, ey, ez */
static void solve33(float K[3][3], const float b[3], float out[3]) {
    const float* ex = K[0]; const float* ey = K[1]; const float* ez = K[2];
    /* det = dot(ex, cross(ey, ez)) */
    float cx = ey[1] * ez[2] - ey[2] * ez[1], cy = ey[2] * ez[0] - ey[0] * ez[2], cz = ey[0] * ez[1] - ey[1] * ez[0];
    float det = ex[0] * cx + ex[1] * cy + ex[2] * cz;
    if (det != 0.0f) det = 1.0f / det;
    /* x = det * dot(b, cross(ey, ez)) */
    out[0] = det * (b[0] * cx + b[1] * cy + b[2] * cz);
    /* y = det * dot(ex, cross(b, ez)) */
    float bx = b[1] * ez[2] - b[2] * ez[1], by = b[2] * ez[0] - b[0] * ez[2], bz = b[0] * ez[1] - b[1] * ez[0];
    out[1] = det * (ex[0] * bx + ex[1] * by + ex[2] * bz);
    /* z = det * dot(ex, cross(ey, b)) */
    float dx = ey[1] * b[2] - ey[2] * b[1], dy = ey[2] * b[0] - ey[0] * b[2], dz = ey[0] * b[1] - ey[1] * b[0];
    out[2] = det * (ex[0] * dx + ex[1] * dy + ex[2] * dz);
}

static V2 solve22(float K[3][3], V2 b) {
    float a11 = K[0][0], a12 = K[1][0], a21 = K[0][1], a22 = K[1][1];
    float det = a11 * a22 - a12 * a21;
    if (det != 0.0f) det = 1.0f / det;
    return v2(det * (a22 * b.x - a12 * b.y), det * (a11 * b.y - a21 * b.x));
}

/* b2RevoluteJoint::InitVelocityConstraints */
static void joint_init_velocity(RevJoint* j, const Body* bodies, V2* c, float* a, V2* v, float* w, float dt_ratio) {
    const Body* A = &bodies[j->a]; const Body* B = &bodies[j->b];
    j->lcA = A->local_center; j->lcB = B->local_center;
    j->mA = A->inv_mass; j->mB = B->inv_mass; j->iA = A->inv_I; j->iB = B->inv_I;
    float aA = a[j->a], aB = a[j->b];
    V2 vA = v[j->a], vB = v[j->b];
    float wA = w[j->a], wB = w[j->b];
    (void)c;
    Rot qA = rot(aA), qB = rot(aB);
    j->rA = rmul(qA, vsub(j->local_anchor_a, j->lcA));
    j->rB = rmul(qB, vsub(j->local_anchor_b, j->lcB));
    float mA = j->mA, mB = j->mB, iA = j->iA, iB = j->iB;
    int fixed_rotation = (iA + iB == 0.0f);
    j->K[0][0] = mA + mB + j->rA.y * j->rA.y * iA + j->rB.y * j->rB.y * iB;
    j->K[1][0] = -j->rA.y * j->rA.x * iA - j->rB.y * j->rB.x * iB;
    j->K[2][0] = -j->rA.y * iA - j->rB.y * iB;
    j->K[0][1] = j->K[1][0];
    j->K[1][1] = mA + mB + j->rA.x * j->rA.x * iA + j->rB.x * j->rB.x * iB;
    j->K[2][1] = j->rA.x * iA + j->rB.x * iB;
    j->K[0][2] = j->K[2][0];
    j->K[1][2] = j->K[2][1];
    j->K[2][2] = iA + iB;
    j->motor_mass = iA + iB;
    if (j->motor_mass > 0.0f) j->motor_mass = 1.0f / j->motor_mass;
    if (!j->enable_motor || fixed_rotation) j->motor_impulse = 0.0f;
    if (j->enable_limit && !fixed_rotation) {
        float joint_angle = aB - aA - j->reference_angle;
        if (fabsf(j->upper - j->lower) < 2.0f * B2_ANGULAR_SLOP) {
            j->limit_state = 3;
        } else if (joint_angle <= j->lower) {
            if (j->limit_state != 1) j->impulse[2] = 0.0f;
            j->limit_state = 1;
        } else if (joint_angle >= j->upper) {
            if (j->limit_state != 2) j->impulse[2] = 0.0f;
            j->limit_state = 2;
        } else {
            j->limit_state = 0;
            j->impulse[2] = 0.0f;
        }
    } else {
        j->limit_state = 0;
    }
    /* warm starting (always on in b2World::Step) */
    j->impulse[0] *= dt_ratio; j->impulse[1] *= dt_ratio; j->impulse[2] *= dt_ratio;
    j->motor_impulse *= dt_ratio;
    V2 P = v2(j->impulse[0], j->impulse[1]);
    vA = vsub(vA, vscale(mA, P));
    wA -= iA * (vcross(j->rA, P) + j->motor_impulse + j->impulse[2]);
    vB = vadd(vB, vscale(mB, P));
    wB += iB * (vcross(j->rB, P) + j->motor_impulse + j->impulse[2]);
    v[j->a] = vA; w[j->a] = wA; v[j->b] = vB; w[j->b] = wB;
}

/* b2RevoluteJoint::SolveVelocityConstraints */
static void joint_solve_velocity(RevJoint* j, V2* v, float* w, float dt) {
    V2 vA = v[j->a], vB = v[j->b];
    float wA = w[j->a], wB = w[j->b];
    float mA = j->mA, mB = j->mB, iA = j->iA, iB = j->iB;
    int fixed_rotation = (iA + iB == 0.0f);
    if (j->enable_motor && j->limit_state != 3 && !fixed_rotation) {
        float Cdot = wB - wA - j->motor_speed;
        float impulse = -j->motor_mass * Cdot;
        float old = j->motor_impulse;
        float max_impulse = dt * j->max_motor_torque;
        j->motor_impulse = clampf(old + impulse, -max_impulse, max_impulse);
        impulse = j->motor_impulse - old;
        wA -= iA * impulse;
        wB += iB * impulse;
    }
    if (j->enable_limit && j->limit_state != 0 && !fixed_rotation) {
        V2 Cdot1 = vsub(vsub(vadd(vB, cross_sv(wB, j->rB)), vA), cross_sv(wA, j->rA));
        float Cdot2 = wB - wA;
        float rhs[3] = {Cdot1.x, Cdot1.y, Cdot2}, imp[3];
        solve33(j->K, rhs, imp);
        imp[0] = -imp[0]; imp[1] = -imp[1]; imp[2] = -imp[2];
        if (j->limit_state == 3) {
            j->impulse[0] += imp[0]; j->impulse[1] += imp[1]; j->impulse[2] += imp[2];
        } else if (j->limit_state == 1) {
            float ni = j->impulse[2] + imp[2];
            if (ni < 0.0f) {
                V2 r = v2(-Cdot1.x + j->impulse[2] * j->K[2][0], -Cdot1.y + j->impulse[2] * j->K[2][1]);
                V2 red = solve22(j->K, r);
                imp[0] = red.x; imp[1] = red.y; imp[2] = -j->impulse[2];
                j->impulse[0] += red.x; j->impulse[1] += red.y; j->impulse[2] = 0.0f;
            } else {
                j->impulse[0] += imp[0]; j->impulse[1] += imp[1]; j->impulse[2] += imp[2];
            }
        } else {
            float ni = j->impulse[2] + imp[2];
            if (ni > 0.0f) {
                V2 r = v2(-Cdot1.x + j->impulse[2] * j->K[2][0], -Cdot1.y + j->impulse[2] * j->K[2][1]);
                V2 red = solve22(j->K, r);
                imp[0] = red.x; imp[1] = red.y; imp[2] = -j->impulse[2];
                j->impulse[0] += red.x; j->impulse[1] += red.y; j->impulse[2] = 0.0f;
            } else {
                j->impulse[0] += imp[0]; j->impulse[1] += imp[1]; j->impulse[2] += imp[2];
            }
        }
        V2 P = v2(imp[0], imp[1]);
        vA = vsub(vA, vscale(mA, P));
        wA -= iA * (vcross(j->rA, P) + imp[2]);
        vB = vadd(vB, vscale(mB, P));
        wB += iB * (vcross(j->rB, P) + imp[2]);
    } else {
        V2 Cdot = vsub(vsub(vadd(vB, cross_sv(wB, j->rB)), vA), cross_sv(wA, j->rA));
        V2 imp = solve22(j->K, v2(-Cdot.x, -Cdot.y));
        j->impulse[0] += imp.x; j->impulse[1] += imp.y;
        vA = vsub(vA, vscale(mA, imp));
        wA -= iA * vcross(j->rA, imp);
        vB = vadd(vB, vscale(mB, imp));
        wB += iB * vcross(j->rB, imp);
    }
    v[j->a] = vA; w[j->a] = wA; v[j->b] = vB; w[j->b] = wB;
}

/* b2RevoluteJoint::SolvePositionConstraints */
static int joint_solve_position(RevJoint* j, V2* c, float* a) {
    V2 cA = c[j->a], cB = c[j->b];
    float aA = a[j->a], aB = a[j->b];
    float angular_error = 0.0f, position_error;
    int fixed_rotation = (j->iA + j->iB == 0.0f);
    if (j->enable_limit && j->limit_state != 0 && !fixed_rotation) {
        float angle = aB - aA - j->reference_angle, limit_impulse = 0.0f;
        if (j->limit_state == 3) {
            float C = clampf(angle - j->lower, -B2_MAX_ANGULAR_CORRECTION, B2_MAX_ANGULAR_CORRECTION);
            limit_impulse = -j->motor_mass * C;
            angular_error = fabsf(C);
        } else if (j->limit_state == 1) {
            float C = angle - j->lower;
            angular_error = -C;
            C = clampf(C + B2_ANGULAR_SLOP, -B2_MAX_ANGULAR_CORRECTION, 0.0f);
            limit_impulse = -j->motor_mass * C;
        } else {
            float C = angle - j->upper;
            angular_error = C;
            C = clampf(C - B2_ANGULAR_SLOP, 0.0f, B2_MAX_ANGULAR_CORRECTION);
            limit_impulse = -j->motor_mass * C;
        }
        aA -= j->iA * limit_impulse;
        aB += j->iB * limit_impulse;
    }
    {
        Rot qA = rot(aA), qB = rot(aB);
        V2 rA = rmul(qA, vsub(j->local_anchor_a, j->lcA));
        V2 rB = rmul(qB, vsub(j->local_anchor_b, j->lcB));
        V2 C = vsub(vsub(vadd(cB, rB), cA), rA);
        position_error = sqrtf(vdot(C, C));
        float mA = j->mA, mB = j->mB, iA = j->iA, iB = j->iB;
        float K[3][3];
        K[0][0] = mA + mB + iA * rA.y * rA.y + iB * rB.y * rB.y;
        K[0][1] = -iA * rA.x * rA.y - iB * rB.x * rB.y;
        K[1][0] = K[0][1];
        K[1][1] = mA + mB + iA * rA.x * rA.x + iB * rB.x * rB.x;
        V2 imp = solve22(K, C);
        imp = v2(-imp.x, -imp.y);
        cA = vsub(cA, vscale(mA, imp));
        aA -= iA * vcross(rA, imp);
        cB = vadd(cB, vscale(mB, imp));
        aB += iB * vcross(rB, imp);
    }
    c[j->a] = cA; a[j->a] = aA; c[j->b] = cB; a[j->b] = aB;
    return position_error <= B2_LINEAR_SLOP && angular_error <= B2_ANGULAR_SLOP;
}

/* One island = bodies[0..nb) joined by joints[0..nj) (already in island order).
 * b2Island::Solve with gravity 0, no damping, no contacts. */
static void island_solve(Body* bodies, int nb, RevJoint* joints, int nj, float h, float dt_ratio, int vel_iters,
                         int pos_iters) {
    V2 c[8], v[8];
    float a[8], w[8];
    for (int i = 0; i < nb; ++i) {
        Body* b = &bodies[i];
        b->c0 = b->c; b->a0 = b->a;
        v[i] = vadd(b->v, vscale(h, vscale(b->inv_mass, b->force)));
        w[i] = b->w + h * b->inv_I * b->torque;
        v[i] = vscale(1.0f / (1.0f + h * 0.0f), v[i]);
        w[i] *= 1.0f / (1.0f + h * 0.0f);
        c[i] = b->c; a[i] = b->a;
    }
    for (int k = 0; k < nj; ++k) joint_init_velocity(&joints[k], bodies, c, a, v, w, dt_ratio);
    for (int it = 0; it < vel_iters; ++it)
        for (int k = 0; k < nj; ++k) joint_solve_velocity(&joints[k], v, w, h);
    for (int i = 0; i < nb; ++i) {
        V2 tr = vscale(h, v[i]);
        if (vdot(tr, tr) > B2_MAX_TRANSLATION * B2_MAX_TRANSLATION) v[i] = vscale(B2_MAX_TRANSLATION / sqrtf(vdot(tr, tr)), v[i]);
        float r = h * w[i];
        if (r * r > B2_MAX_ROTATION * B2_MAX_ROTATION) w[i] *= B2_MAX_ROTATION / fabsf(r);
        c[i] = vadd(c[i], vscale(h, v[i]));
        a[i] += h * w[i];
    }
    int position_solved = 0;
    for (int it = 0; it < pos_iters; ++it) {
        int ok = 1;
        for (int k = 0; k < nj; ++k) ok = joint_solve_position(&joints[k], c, a) && ok;
        if (ok) { position_solved = 1; break; }
    }
    for (int i = 0; i < nb; ++i) {
        Body* b = &bodies[i];
        b->c = c[i]; b->a = a[i]; b->v = v[i]; b->w = w[i];
        b->q = rot(b->a);
        b->p = vsub(b->c, rmul(b->q, b->local_center));
    }
    /* sleeping (allowSleep defaults to true) */
    float min_sleep = 3.4e38f;
    for (int i = 0; i < nb; ++i) {
        Body* b = &bodies[i];
        if (b->w * b->w > B2_ANGULAR_SLEEP_TOL * B2_ANGULAR_SLEEP_TOL ||
            vdot(b->v, b->v) > B2_LINEAR_SLEEP_TOL * B2_LINEAR_SLEEP_TOL) {
            b->sleep_time = 0.0f;
            min_sleep = 0.0f;
        } else {
            b->sleep_time += h;
            if (b->sleep_time < min_sleep) min_sleep = b->sleep_time;
        }
    }
    if (min_sleep >= B2_TIME_TO_SLEEP && position_solved)
        for (int i = 0; i < nb; ++i) body_set_awake(&bodies[i], 0);
}

/* ------------------------------------------------------------------------------------------ */
/* mini Box2D, part 2: polygon-polygon contacts between the two cars of cCarRacingDouble        */
/*                                                                                              */
/* Restated from the published Box2D 2.3.0 algorithm (from memory; PARITY UNPINNED like the     */
/* rest of the mini Box2D): b2PolygonShape::Set (normals, centroid), b2CollidePolygons with     */
/* b2FindMaxSeparation's hill climb / b2FindIncidentEdge / b2ClipSegmentToLine, b2Contact::     */
/* Update (impulses carried over by contact-feature id), b2ContactSolver (warm start, friction  */
/* then normal, 2-point block solver, Baumgarte position correction) and the merged island in   */
/* b2Island::Solve (joints before contacts in the velocity loop, contacts before joints in the  */
/* position loop).                                                                              */
/*                                                                                              */
/* Fixtures of a car: 0..3 = hull polygons (car_dynamics.py:63-68, category 0x0001 mask 0xFFFF),*/
/* 4..7 = wheels 0..3 (:94-96, category 0x0020 mask 0x001).  b2ContactFilter::ShouldCollide     */
/* therefore allows hull-hull and wheel-hull between the two cars but not wheel-wheel; fixtures  */
/* of one car never collide (wheels: mask; hull-wheel: joined with collideConnected = false).   */
/* All fixtures: friction 0.2 (b2FixtureDef default), restitution 0 -> b2MixFriction =          */
/* sqrt(0.2 * 0.2), b2MixRestitution = 0.                                                       */
/*                                                                                              */
/* Stated deviations from a real Box2D world: (1) a b2Contact object exists in Box2D only while */
/* the fat AABBs of two fixtures overlap; its manifold is empty unless the polygons are within  */
/* 2 * b2_polygonRadius, and an empty manifold carries no impulses, so evaluating the manifold  */
/* of every allowed pair at the start of each Step gives the same constraints.  (2) The order   */
/* of contacts inside the island follows the order in which Box2D's broad phase happened to     */
/* create them (proxy ids, move buffer); here it is canonical: fixture of car 0 major, fixture  */
/* of car 1 minor, and fixtureA is always car 0's.  (3) Merged-island joint order is car 1's    */
/* joints (3, 2, 1, 0) then car 0's (3, 2, 1, 0); Box2D's depth-first search can put one joint  */
/* of the second car first when it enters that car through a wheel contact.  Joints of the two  */
/* cars share no body, so only the position of that one joint relative to its siblings differs. */

#define CAR_FIXTURES 8
#define CAR_PAIRS 48                       /* 8 x 8 minus the 16 wheel-wheel pairs */
#define B2_VELOCITY_THRESHOLD 1.0f
#define B2_BAUMGARTE 0.2f
#define B2_MAX_LINEAR_CORRECTION 0.2f
#define B2_EPSILON 1.1920929e-07f

typedef struct { int n; V2 v[8], nrm[8]; V2 centroid; } Poly;
typedef struct { V2 p; Rot q; } Xf;

typedef struct {
    V2 local_point;
    float normal_impulse, tangent_impulse;
    uint32_t id;                 /* b2ContactFeature: indexA | indexB << 8 | typeA << 16 | typeB << 24 */
} MPoint;
typedef struct {
    MPoint pt[2];
    V2 local_normal, local_point;
    int type;                    /* 0 = e_faceA, 1 = e_faceB */
    int count;
} Manifold;
typedef struct { Manifold m[CAR_PAIRS]; } ContactStore;

static V2 vneg(V2 a) { return v2(-a.x, -a.y); }
static V2 xf_mul(Xf t, V2 v) { return v2((t.q.c * v.x - t.q.s * v.y) + t.p.x, (t.q.s * v.x + t.q.c * v.y) + t.p.y); }
static V2 xf_mulT(Xf t, V2 v) {
    float px = v.x - t.p.x, py = v.y - t.p.y;
    return v2(t.q.c * px + t.q.s * py, -t.q.s * px + t.q.c * py);
}
static V2 rmulT(Rot q, V2 v) { return v2(q.c * v.x + q.s * v.y, -q.s * v.x + q.c * v.y); }
static V2 cross_vs(V2 a, float s) { return v2(s * a.y, -s * a.x); }
static V2 normalized(V2 a) {
    float len = sqrtf(a.x * a.x + a.y * a.y);
    if (len < B2_EPSILON) return a;
    float inv = 1.0f / len;
    return v2(a.x * inv, a.y * inv);
}

/* b2PolygonShape::Set after the hull: edge normals and ComputeCentroid (pRef = origin) */
static void poly_set(Poly* p, const float* hull_xy, int n) {
    p->n = n;
    for (int i = 0; i < n; ++i) p->v[i] = v2(hull_xy[2 * i], hull_xy[2 * i + 1]);
    for (int i = 0; i < n; ++i) {
        V2 e = vsub(p->v[i + 1 < n ? i + 1 : 0], p->v[i]);
        p->nrm[i] = normalized(cross_vs(e, 1.0f));
    }
    V2 c = v2(0.f, 0.f);
    float area = 0.f;
    const float inv3 = 1.0f / 3.0f;
    for (int i = 0; i < n; ++i) {
        V2 p2 = p->v[i], p3 = p->v[i + 1 < n ? i + 1 : 0];
        float D = vcross(p2, p3);
        float tri = 0.5f * D;
        area += tri;
        c = vadd(c, vscale(tri * inv3, vadd(vadd(v2(0.f, 0.f), p2), p3)));
    }
    p->centroid = vscale(1.0f / area, c);
}

static float edge_separation(const Poly* p1, Xf xf1, int edge1, const Poly* p2, Xf xf2) {
    V2 n1w = rmul(xf1.q, p1->nrm[edge1]);
    V2 n1 = rmulT(xf2.q, n1w);
    int index = 0;
    float min_dot = 3.402823466e+38f;
    for (int i = 0; i < p2->n; ++i) {
        float d = vdot(p2->v[i], n1);
        if (d < min_dot) { min_dot = d; index = i; }
    }
    V2 v1 = xf_mul(xf1, p1->v[edge1]), v2w = xf_mul(xf2, p2->v[index]);
    return vdot(vsub(v2w, v1), n1w);
}

/* b2FindMaxSeparation (2.3.0): start at the edge most aligned with the centroid offset, climb */
static float find_max_separation(int* edge_index, const Poly* p1, Xf xf1, const Poly* p2, Xf xf2) {
    int count1 = p1->n;
    V2 d = vsub(xf_mul(xf2, p2->centroid), xf_mul(xf1, p1->centroid));
    V2 d_local1 = rmulT(xf1.q, d);
    int edge = 0;
    float max_dot = -3.402823466e+38f;
    for (int i = 0; i < count1; ++i) {
        float dt = vdot(p1->nrm[i], d_local1);
        if (dt > max_dot) { max_dot = dt; edge = i; }
    }
    float s = edge_separation(p1, xf1, edge, p2, xf2);
    int prev_edge = edge - 1 >= 0 ? edge - 1 : count1 - 1;
    float s_prev = edge_separation(p1, xf1, prev_edge, p2, xf2);
    int next_edge = edge + 1 < count1 ? edge + 1 : 0;
    float s_next = edge_separation(p1, xf1, next_edge, p2, xf2);
    int best_edge, increment;
    float best_sep;
    if (s_prev > s && s_prev > s_next) { increment = -1; best_edge = prev_edge; best_sep = s_prev; }
    else if (s_next > s) { increment = 1; best_edge = next_edge; best_sep = s_next; }
    else { *edge_index = edge; return s; }
    for (;;) {
        if (increment == -1) edge = best_edge - 1 >= 0 ? best_edge - 1 : count1 - 1;
        else edge = best_edge + 1 < count1 ? best_edge + 1 : 0;
        s = edge_separation(p1, xf1, edge, p2, xf2);
        if (s > best_sep) { best_edge = edge; best_sep = s; } else break;
    }
    *edge_index = best_edge;
    return best_sep;
}

typedef struct { V2 v; uint32_t id; } ClipVertex;
#define CF_ID(ia, ib, ta, tb) ((uint32_t)(ia) | ((uint32_t)(ib) << 8) | ((uint32_t)(ta) << 16) | ((uint32_t)(tb) << 24))
#define CF_VERTEX 0
#define CF_FACE 1

static int clip_segment_to_line(ClipVertex out[2], const ClipVertex in[2], V2 normal, float offset, int vertex_index_a) {
    int n_out = 0;
    float d0 = vdot(normal, in[0].v) - offset, d1 = vdot(normal, in[1].v) - offset;
    if (d0 <= 0.0f) out[n_out++] = in[0];
    if (d1 <= 0.0f) out[n_out++] = in[1];
    if (d0 * d1 < 0.0f) {
        float interp = d0 / (d0 - d1);
        out[n_out].v = vadd(in[0].v, vscale(interp, vsub(in[1].v, in[0].v)));
        out[n_out].id = CF_ID(vertex_index_a, (in[0].id >> 8) & 0xff, CF_VERTEX, CF_FACE);
        ++n_out;
    }
    return n_out;
}

/* b2CollidePolygons (2.3.0) */
static void collide_polygons(Manifold* m, const Poly* pa, Xf xfa, const Poly* pb, Xf xfb) {
    m->count = 0;
    const float total_radius = B2_POLYGON_RADIUS + B2_POLYGON_RADIUS;
    int edge_a = 0, edge_b = 0;
    float sep_a = find_max_separation(&edge_a, pa, xfa, pb, xfb);
    if (sep_a > total_radius) return;
    float sep_b = find_max_separation(&edge_b, pb, xfb, pa, xfa);
    if (sep_b > total_radius) return;
    const Poly *p1, *p2;
    Xf xf1, xf2;
    int edge1, flip;
    const float k_rel = 0.98f, k_abs = 0.001f;
    if (sep_b > k_rel * sep_a + k_abs) { p1 = pb; p2 = pa; xf1 = xfb; xf2 = xfa; edge1 = edge_b; m->type = 1; flip = 1; }
    else { p1 = pa; p2 = pb; xf1 = xfa; xf2 = xfb; edge1 = edge_a; m->type = 0; flip = 0; }
    /* b2FindIncidentEdge */
    ClipVertex incident[2];
    {
        V2 n1 = rmulT(xf2.q, rmul(xf1.q, p1->nrm[edge1]));
        int index = 0;
        float min_dot = 3.402823466e+38f;
        for (int i = 0; i < p2->n; ++i) {
            float d = vdot(n1, p2->nrm[i]);
            if (d < min_dot) { min_dot = d; index = i; }
        }
        int i1 = index, i2 = i1 + 1 < p2->n ? i1 + 1 : 0;
        incident[0].v = xf_mul(xf2, p2->v[i1]); incident[0].id = CF_ID(edge1, i1, CF_FACE, CF_VERTEX);
        incident[1].v = xf_mul(xf2, p2->v[i2]); incident[1].id = CF_ID(edge1, i2, CF_FACE, CF_VERTEX);
    }
    int iv1 = edge1, iv2 = edge1 + 1 < p1->n ? edge1 + 1 : 0;
    V2 v11 = p1->v[iv1], v12 = p1->v[iv2];
    V2 local_tangent = normalized(vsub(v12, v11));
    V2 local_normal = cross_vs(local_tangent, 1.0f);
    V2 plane_point = vscale(0.5f, vadd(v11, v12));
    V2 tangent = rmul(xf1.q, local_tangent);
    V2 normal = cross_vs(tangent, 1.0f);
    v11 = xf_mul(xf1, v11); v12 = xf_mul(xf1, v12);
    float front_offset = vdot(normal, v11);
    float side_offset1 = -vdot(tangent, v11) + total_radius;
    float side_offset2 = vdot(tangent, v12) + total_radius;
    ClipVertex clip1[2], clip2[2];
    if (clip_segment_to_line(clip1, incident, vneg(tangent), side_offset1, iv1) < 2) return;
    if (clip_segment_to_line(clip2, clip1, tangent, side_offset2, iv2) < 2) return;
    m->local_normal = local_normal;
    m->local_point = plane_point;
    int count = 0;
    for (int i = 0; i < 2; ++i) {
        float separation = vdot(normal, clip2[i].v) - front_offset;
        if (separation <= total_radius) {
            MPoint* cp = &m->pt[count];
            cp->local_point = xf_mulT(xf2, clip2[i].v);
            uint32_t id = clip2[i].id;
            if (flip) id = CF_ID((id >> 8) & 0xff, id & 0xff, (id >> 24) & 0xff, (id >> 16) & 0xff);
            cp->id = id;
            ++count;
        }
    }
    m->count = count;
}

/* b2Contact::Update for one pair: new manifold, impulses carried over by feature id.  Returns touching. */
static int contact_update(Manifold* m, const Poly* pa, Xf xfa, const Poly* pb, Xf xfb, int* touching_changed) {
    Manifold old = *m;
    int was = old.count > 0;
    collide_polygons(m, pa, xfa, pb, xfb);
    for (int i = 0; i < m->count; ++i) {
        MPoint* mp2 = &m->pt[i];
        mp2->normal_impulse = 0.0f; mp2->tangent_impulse = 0.0f;
        for (int j = 0; j < old.count; ++j)
            if (old.pt[j].id == mp2->id) {
                mp2->normal_impulse = old.pt[j].normal_impulse;
                mp2->tangent_impulse = old.pt[j].tangent_impulse;
                break;
            }
    }
    *touching_changed = (m->count > 0) != was;
    return m->count > 0;
}

/* b2ContactVelocityConstraint + b2ContactPositionConstraint of one touching pair */
typedef struct {
    int ia, ib;                  /* island body indices */
    Manifold* man;
    int point_count;             /* velocity constraint count (may drop to 1: ill-conditioned block) */
    V2 normal;
    V2 rA[2], rB[2];
    float normal_impulse[2], tangent_impulse[2], normal_mass[2], tangent_mass[2], velocity_bias[2];
    float K[2][2], NM[2][2];     /* [col][row] */
    float friction, restitution;
    float mA, mB, iA, iB;
    V2 lcA, lcB;
} ContactC;

static Xf xf_of(V2 c, float a, V2 local_center) {
    Xf t;
    t.q = rot(a);
    t.p = vsub(c, rmul(t.q, local_center));
    return t;
}

/* b2WorldManifold::Initialize */
static void world_manifold(const Manifold* m, Xf xfa, Xf xfb, V2* normal, V2 points[2]) {
    const float ra = B2_POLYGON_RADIUS, rb = B2_POLYGON_RADIUS;
    if (m->type == 0) {
        *normal = rmul(xfa.q, m->local_normal);
        V2 plane = xf_mul(xfa, m->local_point);
        for (int i = 0; i < m->count; ++i) {
            V2 clip = xf_mul(xfb, m->pt[i].local_point);
            V2 ca = vadd(clip, vscale(ra - vdot(vsub(clip, plane), *normal), *normal));
            V2 cb = vsub(clip, vscale(rb, *normal));
            points[i] = vscale(0.5f, vadd(ca, cb));
        }
    } else {
        *normal = rmul(xfb.q, m->local_normal);
        V2 plane = xf_mul(xfb, m->local_point);
        for (int i = 0; i < m->count; ++i) {
            V2 clip = xf_mul(xfa, m->pt[i].local_point);
            V2 cb = vadd(clip, vscale(rb - vdot(vsub(clip, plane), *normal), *normal));
            V2 ca = vsub(clip, vscale(ra, *normal));
            points[i] = vscale(0.5f, vadd(ca, cb));
        }
        *normal = vneg(*normal);
    }
}

/* b2ContactSolver constructor + InitializeVelocityConstraints + WarmStart for one contact */
static void contact_init(ContactC* cc, const V2* c, const float* a, V2* v, float* w, float dt_ratio) {
    const Manifold* m = cc->man;
    cc->point_count = m->count;
    for (int j = 0; j < m->count; ++j) {
        cc->normal_impulse[j] = dt_ratio * m->pt[j].normal_impulse;
        cc->tangent_impulse[j] = dt_ratio * m->pt[j].tangent_impulse;
    }
    const float mA = cc->mA, mB = cc->mB, iA = cc->iA, iB = cc->iB;
    V2 cA = c[cc->ia], cB = c[cc->ib];
    V2 vA = v[cc->ia], vB = v[cc->ib];
    float wA = w[cc->ia], wB = w[cc->ib];
    Xf xfa = xf_of(cA, a[cc->ia], cc->lcA), xfb = xf_of(cB, a[cc->ib], cc->lcB);
    V2 pts[2];
    world_manifold(m, xfa, xfb, &cc->normal, pts);
    for (int j = 0; j < cc->point_count; ++j) {
        cc->rA[j] = vsub(pts[j], cA);
        cc->rB[j] = vsub(pts[j], cB);
        float rnA = vcross(cc->rA[j], cc->normal), rnB = vcross(cc->rB[j], cc->normal);
        float k_normal = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
        cc->normal_mass[j] = k_normal > 0.0f ? 1.0f / k_normal : 0.0f;
        V2 tangent = cross_vs(cc->normal, 1.0f);
        float rtA = vcross(cc->rA[j], tangent), rtB = vcross(cc->rB[j], tangent);
        float k_tangent = mA + mB + iA * rtA * rtA + iB * rtB * rtB;
        cc->tangent_mass[j] = k_tangent > 0.0f ? 1.0f / k_tangent : 0.0f;
        cc->velocity_bias[j] = 0.0f;
        float v_rel = vdot(cc->normal, vsub(vsub(vadd(vB, cross_sv(wB, cc->rB[j])), vA), cross_sv(wA, cc->rA[j])));
        if (v_rel < -B2_VELOCITY_THRESHOLD) cc->velocity_bias[j] = -cc->restitution * v_rel;
    }
    if (cc->point_count == 2) {
        float rn1A = vcross(cc->rA[0], cc->normal), rn1B = vcross(cc->rB[0], cc->normal);
        float rn2A = vcross(cc->rA[1], cc->normal), rn2B = vcross(cc->rB[1], cc->normal);
        float k11 = mA + mB + iA * rn1A * rn1A + iB * rn1B * rn1B;
        float k22 = mA + mB + iA * rn2A * rn2A + iB * rn2B * rn2B;
        float k12 = mA + mB + iA * rn1A * rn2A + iB * rn1B * rn2B;
        const float k_max_cond = 1000.0f;
        if (k11 * k11 < k_max_cond * (k11 * k22 - k12 * k12)) {
            cc->K[0][0] = k11; cc->K[0][1] = k12; cc->K[1][0] = k12; cc->K[1][1] = k22;
            float A = k11, B = k12, C = k12, D = k22;   /* a = ex.x, b = ey.x, c = ex.y, d = ey.y */
            float det = A * D - B * C;
            if (det != 0.0f) det = 1.0f / det;
            cc->NM[0][0] = det * D; cc->NM[1][0] = -det * B; cc->NM[0][1] = -det * C; cc->NM[1][1] = det * A;
        } else {
            cc->point_count = 1;
        }
    }
    /* WarmStart */
    V2 tangent = cross_vs(cc->normal, 1.0f);
    for (int j = 0; j < cc->point_count; ++j) {
        V2 P = vadd(vscale(cc->normal_impulse[j], cc->normal), vscale(cc->tangent_impulse[j], tangent));
        wA -= iA * vcross(cc->rA[j], P);
        vA = vsub(vA, vscale(mA, P));
        wB += iB * vcross(cc->rB[j], P);
        vB = vadd(vB, vscale(mB, P));
    }
    v[cc->ia] = vA; w[cc->ia] = wA; v[cc->ib] = vB; w[cc->ib] = wB;
}

/* the constructor and InitializeVelocityConstraints run over ALL contacts before WarmStart runs over all
 * of them; InitializeVelocityConstraints reads velocities only for the restitution bias, which is zero
 * here (restitution 0), so doing init + warm start contact by contact gives the same numbers. */

/* b2ContactSolver::SolveVelocityConstraints for one contact */
static void contact_solve_velocity(ContactC* cc, V2* v, float* w) {
    const float mA = cc->mA, mB = cc->mB, iA = cc->iA, iB = cc->iB;
    V2 vA = v[cc->ia], vB = v[cc->ib];
    float wA = w[cc->ia], wB = w[cc->ib];
    V2 normal = cc->normal, tangent = cross_vs(normal, 1.0f);
    for (int j = 0; j < cc->point_count; ++j) {
        V2 dv = vsub(vsub(vadd(vB, cross_sv(wB, cc->rB[j])), vA), cross_sv(wA, cc->rA[j]));
        float vt = vdot(dv, tangent) - 0.0f;
        float lambda = cc->tangent_mass[j] * (-vt);
        float max_friction = cc->friction * cc->normal_impulse[j];
        float new_impulse = clampf(cc->tangent_impulse[j] + lambda, -max_friction, max_friction);
        lambda = new_impulse - cc->tangent_impulse[j];
        cc->tangent_impulse[j] = new_impulse;
        V2 P = vscale(lambda, tangent);
        vA = vsub(vA, vscale(mA, P));
        wA -= iA * vcross(cc->rA[j], P);
        vB = vadd(vB, vscale(mB, P));
        wB += iB * vcross(cc->rB[j], P);
    }
    if (cc->point_count == 1) {
        V2 dv = vsub(vsub(vadd(vB, cross_sv(wB, cc->rB[0])), vA), cross_sv(wA, cc->rA[0]));
        float vn = vdot(dv, normal);
        float lambda = -cc->normal_mass[0] * (vn - cc->velocity_bias[0]);
        float new_impulse = cc->normal_impulse[0] + lambda;
        if (!(new_impulse > 0.0f)) new_impulse = 0.0f;
        lambda = new_impulse - cc->normal_impulse[0];
        cc->normal_impulse[0] = new_impulse;
        V2 P = vscale(lambda, normal);
        vA = vsub(vA, vscale(mA, P));
        wA -= iA * vcross(cc->rA[0], P);
        vB = vadd(vB, vscale(mB, P));
        wB += iB * vcross(cc->rB[0], P);
    } else {
        V2 a = v2(cc->normal_impulse[0], cc->normal_impulse[1]);
        V2 dv1 = vsub(vsub(vadd(vB, cross_sv(wB, cc->rB[0])), vA), cross_sv(wA, cc->rA[0]));
        V2 dv2 = vsub(vsub(vadd(vB, cross_sv(wB, cc->rB[1])), vA), cross_sv(wA, cc->rA[1]));
        float vn1 = vdot(dv1, normal), vn2 = vdot(dv2, normal);
        V2 b = v2(vn1 - cc->velocity_bias[0], vn2 - cc->velocity_bias[1]);
        b = vsub(b, v2(cc->K[0][0] * a.x + cc->K[1][0] * a.y, cc->K[0][1] * a.x + cc->K[1][1] * a.y));
        V2 x;
        int solved = 0;
        for (;;) {
            x = vneg(v2(cc->NM[0][0] * b.x + cc->NM[1][0] * b.y, cc->NM[0][1] * b.x + cc->NM[1][1] * b.y));
            if (x.x >= 0.0f && x.y >= 0.0f) { solved = 1; break; }
            x.x = -cc->normal_mass[0] * b.x; x.y = 0.0f;
            vn1 = 0.0f; vn2 = cc->K[0][1] * x.x + b.y;
            if (x.x >= 0.0f && vn2 >= 0.0f) { solved = 1; break; }
            x.x = 0.0f; x.y = -cc->normal_mass[1] * b.y;
            vn1 = cc->K[1][0] * x.y + b.x; vn2 = 0.0f;
            if (x.y >= 0.0f && vn1 >= 0.0f) { solved = 1; break; }
            x.x = 0.0f; x.y = 0.0f;
            vn1 = b.x; vn2 = b.y;
            if (vn1 >= 0.0f && vn2 >= 0.0f) { solved = 1; break; }
            break;
        }
        if (solved) {
            V2 d = vsub(x, a);
            V2 P1 = vscale(d.x, normal), P2 = vscale(d.y, normal);
            vA = vsub(vA, vscale(mA, vadd(P1, P2)));
            wA -= iA * (vcross(cc->rA[0], P1) + vcross(cc->rA[1], P2));
            vB = vadd(vB, vscale(mB, vadd(P1, P2)));
            wB += iB * (vcross(cc->rB[0], P1) + vcross(cc->rB[1], P2));
            cc->normal_impulse[0] = x.x; cc->normal_impulse[1] = x.y;
        }
    }
    v[cc->ia] = vA; w[cc->ia] = wA; v[cc->ib] = vB; w[cc->ib] = wB;
}

/* b2ContactSolver::SolvePositionConstraints for one contact; returns its min separation */
static float contact_solve_position(const ContactC* cc, V2* c, float* a) {
    const Manifold* m = cc->man;
    const float mA = cc->mA, mB = cc->mB, iA = cc->iA, iB = cc->iB;
    V2 cA = c[cc->ia], cB = c[cc->ib];
    float aA = a[cc->ia], aB = a[cc->ib];
    float min_sep = 0.0f;
    for (int j = 0; j < m->count; ++j) {
        Xf xfa = xf_of(cA, aA, cc->lcA), xfb = xf_of(cB, aB, cc->lcB);
        V2 normal, point;
        float separation;
        if (m->type == 0) {
            normal = rmul(xfa.q, m->local_normal);
            V2 plane = xf_mul(xfa, m->local_point);
            V2 clip = xf_mul(xfb, m->pt[j].local_point);
            separation = vdot(vsub(clip, plane), normal) - B2_POLYGON_RADIUS - B2_POLYGON_RADIUS;
            point = clip;
        } else {
            normal = rmul(xfb.q, m->local_normal);
            V2 plane = xf_mul(xfb, m->local_point);
            V2 clip = xf_mul(xfa, m->pt[j].local_point);
            separation = vdot(vsub(clip, plane), normal) - B2_POLYGON_RADIUS - B2_POLYGON_RADIUS;
            point = clip;
            normal = vneg(normal);
        }
        V2 rA = vsub(point, cA), rB = vsub(point, cB);
        if (separation < min_sep) min_sep = separation;
        float C = clampf(B2_BAUMGARTE * (separation + B2_LINEAR_SLOP), -B2_MAX_LINEAR_CORRECTION, 0.0f);
        float rnA = vcross(rA, normal), rnB = vcross(rB, normal);
        float K = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
        float impulse = K > 0.0f ? -C / K : 0.0f;
        V2 P = vscale(impulse, normal);
        cA = vsub(cA, vscale(mA, P));
        aA -= iA * vcross(rA, P);
        cB = vadd(cB, vscale(mB, P));
        aB += iB * vcross(rB, P);
    }
    c[cc->ia] = cA; a[cc->ia] = aA; c[cc->ib] = cB; a[cc->ib] = aB;
    return min_sep;
}

/* b2Island::Solve over bodies[0..nb) (pointers), joints (indices into the island) and contacts */
static void island_solve_ex(Body** bodies, int nb, RevJoint** joints, int nj, ContactC* contacts, int nc, float h,
                            float dt_ratio, int vel_iters, int pos_iters) {
    V2 c[16], v[16];
    float a[16], w[16];
    Body flat[16];
    for (int i = 0; i < nb; ++i) {
        Body* b = bodies[i];
        b->c0 = b->c; b->a0 = b->a;
        v[i] = vadd(b->v, vscale(h, vscale(b->inv_mass, b->force)));
        w[i] = b->w + h * b->inv_I * b->torque;
        v[i] = vscale(1.0f / (1.0f + h * 0.0f), v[i]);
        w[i] *= 1.0f / (1.0f + h * 0.0f);
        c[i] = b->c; a[i] = b->a;
        flat[i] = *b;
    }
    for (int k = 0; k < nc; ++k) contact_init(&contacts[k], c, a, v, w, dt_ratio);
    for (int k = 0; k < nj; ++k) joint_init_velocity(joints[k], flat, c, a, v, w, dt_ratio);
    for (int it = 0; it < vel_iters; ++it) {
        for (int k = 0; k < nj; ++k) joint_solve_velocity(joints[k], v, w, h);
        for (int k = 0; k < nc; ++k) contact_solve_velocity(&contacts[k], v, w);
    }
    for (int k = 0; k < nc; ++k)       /* StoreImpulses */
        for (int j = 0; j < contacts[k].point_count; ++j) {
            contacts[k].man->pt[j].normal_impulse = contacts[k].normal_impulse[j];
            contacts[k].man->pt[j].tangent_impulse = contacts[k].tangent_impulse[j];
        }
    for (int i = 0; i < nb; ++i) {
        V2 tr = vscale(h, v[i]);
        if (vdot(tr, tr) > B2_MAX_TRANSLATION * B2_MAX_TRANSLATION) v[i] = vscale(B2_MAX_TRANSLATION / sqrtf(vdot(tr, tr)), v[i]);
        float r = h * w[i];
        if (r * r > B2_MAX_ROTATION * B2_MAX_ROTATION) w[i] *= B2_MAX_ROTATION / fabsf(r);
        c[i] = vadd(c[i], vscale(h, v[i]));
        a[i] += h * w[i];
    }
    int position_solved = 0;
    for (int it = 0; it < pos_iters; ++it) {
        float min_sep = 0.0f;
        for (int k = 0; k < nc; ++k) {
            float s = contact_solve_position(&contacts[k], c, a);
            if (s < min_sep) min_sep = s;
        }
        int contacts_ok = min_sep >= -3.0f * B2_LINEAR_SLOP;
        int joints_ok = 1;
        for (int k = 0; k < nj; ++k) joints_ok = joint_solve_position(joints[k], c, a) && joints_ok;
        if (contacts_ok && joints_ok) { position_solved = 1; break; }
    }
    float min_sleep = 3.4e38f;
    for (int i = 0; i < nb; ++i) {
        Body* b = bodies[i];
        b->c = c[i]; b->a = a[i]; b->v = v[i]; b->w = w[i];
        b->q = rot(b->a);
        b->p = vsub(b->c, rmul(b->q, b->local_center));
    }
    for (int i = 0; i < nb; ++i) {
        Body* b = bodies[i];
        if (b->w * b->w > B2_ANGULAR_SLEEP_TOL * B2_ANGULAR_SLEEP_TOL ||
            vdot(b->v, b->v) > B2_LINEAR_SLEEP_TOL * B2_LINEAR_SLEEP_TOL) {
            b->sleep_time = 0.0f;
            min_sleep = 0.0f;
        } else {
            b->sleep_time += h;
            if (b->sleep_time < min_sleep) min_sleep = b->sleep_time;
        }
    }
    if (min_sleep >= B2_TIME_TO_SLEEP && position_solved)
        for (int i = 0; i < nb; ++i) body_set_awake(bodies[i], 0);
}

/* ------------------------------------------------------------------------------------------ */
/* exported low-level hooks for the Box2D stand-in (oracle/ref_shim/Box2D): it keeps the bodies
 * and joints in flat arrays of these structs and calls back into the solver above.             */

int car_oracle_sizeof_body(void) { return (int)sizeof(Body); }
int car_oracle_sizeof_joint(void) { return (int)sizeof(RevJoint); }
void car_oracle_body_init(Body* b, float x, float y, float angle) { body_init(b, x, y, angle); }
void car_oracle_body_set_mass(Body* b, const float* polys_xy, const int* counts, const float* density, int n_polys) {
    body_set_mass(b, (const V2*)polys_xy, counts, density, n_polys);
}
void car_oracle_island_solve(Body* bodies, int nb, RevJoint* joints, int nj, float h, float dt_ratio, int vel_iters,
                             int pos_iters) {
    island_solve(bodies, nb, joints, nj, h, dt_ratio, vel_iters, pos_iters);
}

/* max separation of convex polygon B from the faces of convex polygon A (both CCW, world coords) */
static float max_separation(const V2* A, int na, const V2* B, int nb) {
    float best = -3.4e38f;
    for (int i = 0; i < na; ++i) {
        V2 e = vsub(A[(i + 1) % na], A[i]);
        float len = sqrtf(vdot(e, e));
        if (len < 1e-12f) continue;
        V2 n = v2(e.y / len, -e.x / len);   /* outward normal of a CCW polygon */
        float mn = 3.4e38f;
        for (int k = 0; k < nb; ++k) {
            float d = vdot(n, vsub(B[k], A[i]));
            if (d < mn) mn = d;
        }
        if (mn > best) best = mn;
    }
    return best;
}

/* sensor overlap used for wheel-tile contacts: polygons (with their b2_polygonRadius skins) touch */
int car_oracle_polys_touch(const float* a_xy, int na, const float* b_xy, int nb) {
    const V2* A = (const V2*)a_xy; const V2* B = (const V2*)b_xy;
    float s1 = max_separation(A, na, B, nb), s2 = max_separation(B, nb, A, na);
    float s = s1 > s2 ? s1 : s2;
    return s < 2.0f * B2_POLYGON_RADIUS;
}

/* b2PolygonShape::Set: convex hull (gift wrapping) of up to 8 points, CCW; returns the count */
int car_oracle_convex_hull(const float* in_xy, int n, float* out_xy) {
    V2 ps[8];
    int m = 0;
    for (int i = 0; i < n && i < 8; ++i) {   /* weld points closer than 0.5 * linearSlop */
        V2 p = v2(in_xy[2 * i], in_xy[2 * i + 1]);
        int unique = 1;
        for (int k = 0; k < m; ++k) {
            V2 d = vsub(p, ps[k]);
            if (vdot(d, d) < 0.5f * B2_LINEAR_SLOP * 0.5f * B2_LINEAR_SLOP) { unique = 0; break; }
        }
        if (unique) ps[m++] = p;
    }
    if (m < 3) return 0;
    int i0 = 0;
    for (int i = 1; i < m; ++i)
        if (ps[i].x > ps[i0].x || (ps[i].x == ps[i0].x && ps[i].y < ps[i0].y)) i0 = i;
    int hull[8], cnt = 0, ih = i0;
    for (;;) {
        hull[cnt] = ih;
        int ie = 0;
        for (int j = 1; j < m; ++j) {
            if (ie == ih) { ie = j; continue; }
            V2 r = vsub(ps[ie], ps[hull[cnt]]), v = vsub(ps[j], ps[hull[cnt]]);
            float c = vcross(r, v);
            if (c < 0.0f) ie = j;
            if (c == 0.0f && vdot(v, v) > vdot(r, r)) ie = j;
        }
        ++cnt;
        ih = ie;
        if (ie == i0) break;
    }
    for (int i = 0; i < cnt; ++i) { out_xy[2 * i] = ps[hull[i]].x; out_xy[2 * i + 1] = ps[hull[i]].y; }
    return cnt;
}

/* ------------------------------------------------------------------------------------------ */
/* track generator: CarRacing._create_track, car_racing_multi_players.py:262-452.  Pure float64  */
/* Python math; `draws` are the 2*CHECKPOINTS np_random.uniform values of this attempt in draw   */
/* order (noise_c, rad_c per checkpoint).  Returns the number of track points, 0 on failure.     */

typedef struct { double alpha, beta, x, y; } TrackPt;

/* red-white border on hard turns, car_racing_multi_players.py:383-397 (also run on a track loaded from JSON, :376-381) */
static void track_border(const double* beta, int n, int* border) {
    for (int k = 0; k < n; ++k) {
        int good = 1, oneside = 0;
        for (int neg = 0; neg < BORDER_MIN_COUNT; ++neg) {
            double b1 = beta[((k - neg - 0) % n + n) % n], b2 = beta[((k - neg - 1) % n + n) % n];
            good &= fabs(b1 - b2) > TRACK_TURN_RATE * 0.2;
            oneside += (b1 - b2 > 0) - (b1 - b2 < 0);
        }
        good &= abs(oneside) == BORDER_MIN_COUNT;
        border[k] = good;
    }
    for (int k = 0; k < n; ++k)
        for (int neg = 0; neg < BORDER_MIN_COUNT; ++neg) border[((k - neg) % n + n) % n] |= border[k];
}
/* border flags of a given track [n][4] (alpha, beta, x, y) */
void car_oracle_track_border(const double* track, int n, int* border_out) {
    double betas[MAX_TRACK];
    for (int k = 0; k < n && k < MAX_TRACK; ++k) betas[k] = track[4 * k + 1];
    track_border(betas, n < MAX_TRACK ? n : MAX_TRACK, border_out);
}

int car_oracle_create_track(const double* draws, double* out /* [MAX_TRACK][4] */, int* border_out) {
    double cp_alpha[CHECKPOINTS], cp_x[CHECKPOINTS], cp_y[CHECKPOINTS];
    double start_alpha = 0.0;
    for (int c = 0; c < CHECKPOINTS; ++c) {
        double noise = draws[2 * c];
        double alpha = 2 * M_PI * c / CHECKPOINTS + noise;
        double rad = draws[2 * c + 1];
        if (c == 0) { alpha = 0; rad = 1.5 * TRACK_RAD; }
        if (c == CHECKPOINTS - 1) {
            alpha = 2 * M_PI * c / CHECKPOINTS;
            start_alpha = 2 * M_PI * (-0.5) / CHECKPOINTS;
            rad = 1.5 * TRACK_RAD;
        }
        cp_alpha[c] = alpha; cp_x[c] = rad * cos(alpha); cp_y[c] = rad * sin(alpha);
    }
    static __thread TrackPt raw[MAX_TRACK_RAW];
    int n_raw = 0;
    double x = 1.5 * TRACK_RAD, y = 0, beta = 0;
    int dest_i = 0, laps = 0, no_freeze = 2500, visited_other_side = 0;
    for (;;) {
        double alpha = atan2(y, x);
        if (visited_other_side && alpha > 0) { laps += 1; visited_other_side = 0; }
        if (alpha < 0) { visited_other_side = 1; alpha += 2 * M_PI; }
        double dest_alpha, dest_x, dest_y;
        for (;;) {
            int failed = 1;
            for (;;) {
                dest_alpha = cp_alpha[dest_i % CHECKPOINTS]; dest_x = cp_x[dest_i % CHECKPOINTS]; dest_y = cp_y[dest_i % CHECKPOINTS];
                if (alpha <= dest_alpha) { failed = 0; break; }
                dest_i += 1;
                if (dest_i % CHECKPOINTS == 0) break;
            }
            if (!failed) break;
            alpha -= 2 * M_PI;
        }
        double r1x = cos(beta), r1y = sin(beta);
        double p1x = -r1y, p1y = r1x;
        double dest_dx = dest_x - x, dest_dy = dest_y - y;
        double proj = r1x * dest_dx + r1y * dest_dy;
        while (beta - alpha > 1.5 * M_PI) beta -= 2 * M_PI;
        while (beta - alpha < -1.5 * M_PI) beta += 2 * M_PI;
        double prev_beta = beta;
        proj *= SCALE;
        if (proj > 0.3) beta -= fmin(TRACK_TURN_RATE, fabs(0.001 * proj));
        if (proj < -0.3) beta += fmin(TRACK_TURN_RATE, fabs(0.001 * proj));
        x += p1x * TRACK_DETAIL_STEP;
        y += p1y * TRACK_DETAIL_STEP;
        if (n_raw >= MAX_TRACK_RAW) return 0;
        raw[n_raw].alpha = alpha; raw[n_raw].beta = prev_beta * 0.5 + beta * 0.5; raw[n_raw].x = x; raw[n_raw].y = y;
        n_raw++;
        if (laps > 4) break;
        no_freeze -= 1;
        if (no_freeze == 0) break;
    }
    int i1 = -1, i2 = -1, i = n_raw;
    for (;;) {
        i -= 1;
        if (i == 0) return 0;
        int pass = raw[i].alpha > start_alpha && raw[i - 1].alpha <= start_alpha;
        if (pass && i2 == -1) i2 = i;
        else if (pass && i1 == -1) { i1 = i; break; }
    }
    int n = (i2 - 1) - i1;
    if (n <= 0 || n > MAX_TRACK) return 0;
    const TrackPt* t = raw + i1;
    double fb = t[0].beta, fpx = cos(fb), fpy = sin(fb);
    double dx = fpx * (t[0].x - t[n - 1].x), dy = fpy * (t[0].y - t[n - 1].y);
    if (sqrt(dx * dx + dy * dy) > TRACK_DETAIL_STEP) return 0;
    /* red-white border on hard turns */
    int border[MAX_TRACK];
    double betas[MAX_TRACK];
    for (int k = 0; k < n; ++k) betas[k] = t[k].beta;
    track_border(betas, n, border);
    for (int k = 0; k < n; ++k) {
        out[4 * k] = t[k].alpha; out[4 * k + 1] = t[k].beta; out[4 * k + 2] = t[k].x; out[4 * k + 3] = t[k].y;
        if (border_out) border_out[k] = border[k];
    }
    return n;
}

/* ------------------------------------------------------------------------------------------ */
/* car + env                                                                                    */

static const float HULL_POLYS[4][8][2] = {
    {{-60, +130}, {+60, +130}, {+60, +110}, {-60, +110}},
    {{-15, +120}, {+15, +120}, {+20, +20}, {-20, 20}},
    {{+25, +20}, {+50, -10}, {+50, -40}, {+20, -90}, {-20, -90}, {-50, -40}, {-50, -10}, {-25, +20}},
    {{-50, -120}, {+50, -120}, {+50, -90}, {-50, -90}}};
static const int HULL_COUNTS[4] = {4, 4, 8, 4};
static const float WHEELPOS[4][2] = {{-55, +80}, {+55, +80}, {-55, -82}, {+55, -82}};

typedef struct {
    Body body[5];                /* island order: wheel3, hull, wheel0, wheel1, wheel2 (b2World::Solve DFS) */
    RevJoint joint[4];           /* island order: joint3, joint2, joint1, joint0 */
    double gas[4], brake[4], steer[4], phase[4], omega[4];
    uint32_t touching[4][MAX_TRACK / 32];   /* wheel.tiles */
    int n_touching[4];
    uint32_t visited[MAX_TRACK / 32];       /* tile.road_visited[car] */
    int last_block, has_block;              /* block_visited[car][-1] */
    int tile_visited_count;
    double reward, prev_reward;
    int done;
} Car;

static const int BODY_OF_WHEEL[4] = {2, 3, 4, 0};   /* wheel k -> index in Car.body */
#define HULL_BODY 1
static const int JOINT_OF_WHEEL[4] = {3, 2, 1, 0};

/* ---- two-car world: fixtures, b2ContactManager::Collide for the car-car pairs, b2World::Solve ---- */

static Poly g_fix_poly[CAR_FIXTURES];        /* body-local polygons: 0..3 hull (of the hull body), 4..7 wheel box */
static float g_fix_radius[CAR_FIXTURES];     /* bounding radius about the body origin (quick reject only) */
static int g_fix_ready = 0;

static void fixtures_init(void) {
    if (g_fix_ready) return;
    for (int k = 0; k < 4; ++k) {
        float raw[16], hull[16];
        for (int i = 0; i < HULL_COUNTS[k]; ++i) {
            raw[2 * i] = (float)(HULL_POLYS[k][i][0] * SIZE);
            raw[2 * i + 1] = (float)(HULL_POLYS[k][i][1] * SIZE);
        }
        int n = car_oracle_convex_hull(raw, HULL_COUNTS[k], hull);
        poly_set(&g_fix_poly[k], hull, n);
    }
    {
        const float wp[4][2] = {{-WHEEL_W, +WHEEL_R}, {+WHEEL_W, +WHEEL_R}, {+WHEEL_W, -WHEEL_R}, {-WHEEL_W, -WHEEL_R}};
        float raw[8], hull[16];
        for (int i = 0; i < 4; ++i) { raw[2 * i] = (float)(wp[i][0] * 1.0 * SIZE); raw[2 * i + 1] = (float)(wp[i][1] * 1.0 * SIZE); }
        int n = car_oracle_convex_hull(raw, 4, hull);
        for (int k = 4; k < 8; ++k) poly_set(&g_fix_poly[k], hull, n);
    }
    for (int k = 0; k < CAR_FIXTURES; ++k) {
        float r2 = 0.f;
        for (int i = 0; i < g_fix_poly[k].n; ++i) {
            float d = vdot(g_fix_poly[k].v[i], g_fix_poly[k].v[i]);
            if (d > r2) r2 = d;
        }
        g_fix_radius[k] = sqrtf(r2);
    }
    g_fix_ready = 1;
}

static int body_of_fixture(int f) { return f < 4 ? HULL_BODY : BODY_OF_WHEEL[f - 4]; }

/* b2World::Step (collide + solve) for the dynamic bodies of a two-car world.  b0/j0 = bodies and joints
 * of car 0 (created first) in island order, b1/j1 = car 1.  Tile sensors are handled by the caller. */
static void world_step_two(Body* b0, RevJoint* j0, Body* b1, RevJoint* j1, ContactStore* cs, float h, float dt_ratio,
                           int vel_iters, int pos_iters) {
    fixtures_init();
    int touching = 0, pair = 0;
    for (int fa = 0; fa < CAR_FIXTURES; ++fa)
        for (int fb = 0; fb < CAR_FIXTURES; ++fb) {
            if (fa >= 4 && fb >= 4) continue;
            Body* A = &b0[body_of_fixture(fa)];
            Body* B = &b1[body_of_fixture(fb)];
            Manifold* m = &cs->m[pair++];
            V2 d = vsub(B->p, A->p);
            float reach = g_fix_radius[fa] + g_fix_radius[fb] + 0.1f;
            int changed = 0, t = 0;
            if (vdot(d, d) > reach * reach) {           /* certainly separated: empty manifold */
                changed = m->count > 0;
                m->count = 0;
            } else {
                Xf xa = {A->p, A->q}, xb = {B->p, B->q};
                t = contact_update(m, &g_fix_poly[fa], xa, &g_fix_poly[fb], xb, &changed);
            }
            if (changed) { body_set_awake(A, 1); body_set_awake(B, 1); }
            touching |= t;
        }
    if (!touching) {
        for (int ci = 1; ci >= 0; --ci) {
            Body* b = ci ? b1 : b0;
            RevJoint* j = ci ? j1 : j0;
            int any_awake = 0;
            for (int i = 0; i < 5; ++i) any_awake |= b[i].awake;
            if (any_awake) {
                for (int i = 0; i < 5; ++i) body_set_awake(&b[i], 1);
                island_solve(b, 5, j, 4, h, dt_ratio, vel_iters, pos_iters);
            }
        }
        return;
    }
    int any_awake = 0;
    for (int i = 0; i < 5; ++i) any_awake |= b0[i].awake | b1[i].awake;
    if (!any_awake) return;
    Body* bodies[10];
    RevJoint* joints[8];
    for (int i = 0; i < 5; ++i) { bodies[i] = &b1[i]; bodies[5 + i] = &b0[i]; body_set_awake(&b1[i], 1); body_set_awake(&b0[i], 1); }
    for (int k = 0; k < 4; ++k) { joints[k] = &j1[k]; joints[4 + k] = &j0[k]; j0[k].a += 5; j0[k].b += 5; }
    ContactC cc[CAR_PAIRS];
    int nc = 0;
    pair = 0;
    for (int fa = 0; fa < CAR_FIXTURES; ++fa)
        for (int fb = 0; fb < CAR_FIXTURES; ++fb) {
            if (fa >= 4 && fb >= 4) continue;
            Manifold* m = &cs->m[pair++];
            if (m->count == 0) continue;
            ContactC* c = &cc[nc++];
            memset(c, 0, sizeof *c);
            c->man = m;
            c->ia = 5 + body_of_fixture(fa);
            c->ib = body_of_fixture(fb);
            const Body* A = bodies[c->ia]; const Body* B = bodies[c->ib];
            c->mA = A->inv_mass; c->iA = A->inv_I; c->lcA = A->local_center;
            c->mB = B->inv_mass; c->iB = B->inv_I; c->lcB = B->local_center;
            c->friction = sqrtf(0.2f * 0.2f);
            c->restitution = 0.0f;
        }
    island_solve_ex(bodies, 10, joints, 8, cc, nc, h, dt_ratio, vel_iters, pos_iters);
    for (int k = 0; k < 4; ++k) { j0[k].a -= 5; j0[k].b -= 5; }
}

int car_oracle_sizeof_contact_store(void) { return (int)sizeof(ContactStore); }
void car_oracle_world_step_two(Body* b0, RevJoint* j0, Body* b1, RevJoint* j1, ContactStore* cs, float h, float dt_ratio,
                               int vel_iters, int pos_iters) {
    world_step_two(b0, j0, b1, j1, cs, h, dt_ratio, vel_iters, pos_iters);
}
/* number of touching car-car contacts / manifold points in the store (diagnostics for the tests) */
int car_oracle_contact_count(const ContactStore* cs, int* n_points) {
    int nc = 0, np = 0;
    for (int i = 0; i < CAR_PAIRS; ++i) { nc += cs->m[i].count > 0; np += cs->m[i].count; }
    if (n_points) *n_points = np;
    return nc;
}

typedef struct {
    int n_cars, action_repeat;
    int n_track;
    double track[MAX_TRACK][4];
    int border[MAX_TRACK];
    float tile_poly[MAX_TRACK][5][2]; int tile_n[MAX_TRACK];   /* convex hulls, CCW */
    float tile_raw[MAX_TRACK][5][2];                            /* as listed (draw order) */
    float tile_aabb[MAX_TRACK][4];
    float kerb[MAX_TRACK][4][2];
    int tile_map[MAX_TRACK][5][2], kerb_map[MAX_TRACK][4][2];   /* road-map pixel coordinates (int-truncated) */
    Car car[MAX_CARS];
    ContactStore contacts;      /* car-car manifolds (two-car worlds) */
    int step_count;
    float inv_dt0;
    uint8_t obs[MAX_CARS][STATE_H * STATE_W];
    int lazy_render;            /* 1: step/reset do not render; car_oracle_obs renders on demand */
    const uint8_t* glyphs;      /* [11][8][4] bitmaps of "0123456789-" (non-AA COMIC 5 px) then [11] advances; may be NULL */
} CarEnv;

static void wheel_world_poly(const Body* b, V2 out[4]) {
    const float hw = WHEEL_W * (float)SIZE, hr = WHEEL_R * (float)SIZE;
    /* CCW order */
    V2 loc[4] = {{-hw, -hr}, {+hw, -hr}, {+hw, +hr}, {-hw, +hr}};
    for (int i = 0; i < 4; ++i) out[i] = vadd(rmul(b->q, loc[i]), b->p);
}

/* Car.__init__, car_dynamics.py:55-129 */
static void car_create(Car* c, double init_angle, double init_x, double init_y, int birth_place_index) {
    memset(c, 0, sizeof *c);
    init_x -= birth_place_index % 2 * 5;
    init_y -= floor(birth_place_index / 2) * 10;
    Body* hull = &c->body[HULL_BODY];
    body_init(hull, (float)init_x, (float)init_y, (float)init_angle);
    V2 polys[20];
    float dens[4] = {1.0f, 1.0f, 1.0f, 1.0f};
    int at = 0, counts[4];
    for (int k = 0; k < 4; ++k) {
        float raw[16], hullpts[16];
        for (int i = 0; i < HULL_COUNTS[k]; ++i) {
            raw[2 * i] = (float)(HULL_POLYS[k][i][0] * SIZE);
            raw[2 * i + 1] = (float)(HULL_POLYS[k][i][1] * SIZE);
        }
        int n = car_oracle_convex_hull(raw, HULL_COUNTS[k], hullpts);
        for (int i = 0; i < n; ++i) polys[at + i] = v2(hullpts[2 * i], hullpts[2 * i + 1]);
        counts[k] = n;
        at += n;
    }
    body_set_mass(hull, polys, counts, dens, 4);
    for (int k = 0; k < 4; ++k) {
        Body* w = &c->body[BODY_OF_WHEEL[k]];
        /* position = (init_x + wx*SIZE, init_y + wy*SIZE): NOT rotated by init_angle (car_dynamics.py:90) */
        body_init(w, (float)(init_x + WHEELPOS[k][0] * SIZE), (float)(init_y + WHEELPOS[k][1] * SIZE), (float)init_angle);
        const float hw = (float)(WHEEL_W * SIZE), hr = (float)(WHEEL_R * SIZE);
        V2 box[4] = {{+hw, -hr}, {+hw, +hr}, {-hw, +hr}, {-hw, -hr}};   /* hull order of b2PolygonShape::Set */
        float d = 0.1f;
        int cnt = 4;
        body_set_mass(w, box, &cnt, &d, 1);
        RevJoint* j = &c->joint[JOINT_OF_WHEEL[k]];
        memset(j, 0, sizeof *j);
        j->a = HULL_BODY; j->b = BODY_OF_WHEEL[k];
        j->local_anchor_a = v2((float)(WHEELPOS[k][0] * SIZE), (float)(WHEELPOS[k][1] * SIZE));
        j->local_anchor_b = v2(0.f, 0.f);
        j->enable_motor = 1; j->enable_limit = 1;
        j->max_motor_torque = (float)(180 * 900 * SIZE * SIZE);
        j->motor_speed = 0.f; j->lower = -0.4f; j->upper = +0.4f;
        j->reference_angle = 0.f;
    }
    c->last_block = 0; c->has_block = 0;
}

static double clipd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }
static double signd(double v) { return (v > 0) - (v < 0); }

/* Car.gas / brake / steer, car_dynamics.py:131-157 */
static void car_controls(Car* c, double steer, double gas, double brake) {
    c->steer[0] = steer; c->steer[1] = steer;
    gas = clipd(gas, 0, 1);
    for (int k = 2; k < 4; ++k) {
        double diff = gas - c->gas[k];
        if (diff > 0.1) diff = 0.1;
        c->gas[k] += diff;
    }
    for (int k = 0; k < 4; ++k) c->brake[k] = brake;
}

/* Car.step, car_dynamics.py:159-234 (skid particles are render-only and omitted) */
static void car_step(Car* c, double dt) {
    for (int k = 0; k < 4; ++k) {
        Body* w = &c->body[BODY_OF_WHEEL[k]];
        Body* hull = &c->body[HULL_BODY];
        RevJoint* j = &c->joint[JOINT_OF_WHEEL[k]];
        double joint_angle = (double)(w->a - hull->a - j->reference_angle);
        double dir = signd(c->steer[k] - joint_angle);
        double val = fabs(c->steer[k] - joint_angle);
        j->motor_speed = (float)(dir * fmin(50.0 * val, 3.0));
        /* joint.motorSpeed setter wakes both bodies (b2RevoluteJoint::SetMotorSpeed) */
        body_set_awake(hull, 1); body_set_awake(w, 1);
        double friction_limit = FRICTION_LIMIT * 0.6;
        if (c->n_touching[k] > 0) friction_limit = fmax(friction_limit, FRICTION_LIMIT * 1.0);
        V2 forw = rmul(w->q, v2(0.f, 1.f)), side = rmul(w->q, v2(1.f, 0.f));
        double vx = w->v.x, vy = w->v.y;
        double vf = (double)forw.x * vx + (double)forw.y * vy;
        double vs = (double)side.x * vx + (double)side.y * vy;
        c->omega[k] += dt * ENGINE_POWER * c->gas[k] / WHEEL_MOMENT_OF_INERTIA / (fabs(c->omega[k]) + 5.0);
        if (c->brake[k] >= 0.9) {
            c->omega[k] = 0;
        } else if (c->brake[k] > 0) {
            double bdir = -signd(c->omega[k]);
            double bval = 15 * c->brake[k];
            if (fabs(bval) > fabs(c->omega[k])) bval = fabs(c->omega[k]);
            c->omega[k] += bdir * bval;
        }
        c->phase[k] += c->omega[k] * dt;
        double vr = c->omega[k] * (WHEEL_R * SIZE);
        double f_force = -vf + vr, p_force = -vs;
        f_force *= 205000 * SIZE * SIZE;
        p_force *= 205000 * SIZE * SIZE;
        double force = sqrt(f_force * f_force + p_force * p_force);
        if (fabs(force) > friction_limit) {
            f_force /= force; p_force /= force;
            force = friction_limit;
            f_force *= force; p_force *= force;
        }
        c->omega[k] -= dt * f_force * (WHEEL_R * SIZE) / WHEEL_MOMENT_OF_INERTIA;
        /* ApplyForceToCenter((..), True): python floats -> b2Vec2 float32 */
        float fx = (float)(p_force * (double)side.x + f_force * (double)forw.x);
        float fy = (float)(p_force * (double)side.y + f_force * (double)forw.y);
        if (!w->awake) body_set_awake(w, 1);
        w->force = vadd(w->force, v2(fx, fy));
    }
}

/* FrictionDetector._contact (begin) for wheel k of car c touching tile t, :111-153 */
static void contact_begin(CarEnv* e, Car* c, int k, int t) {
    c->touching[k][t >> 5] |= 1u << (t & 31);
    c->n_touching[k] += 1;
    if (!((c->visited[t >> 5] >> (t & 31)) & 1u)) {
        int last_blk = c->has_block ? c->last_block : 0;
        if (t - last_blk < 50) {
            c->last_block = t; c->has_block = 1;
            c->reward += 1000.0 / e->n_track;
        }
        c->visited[t >> 5] |= 1u << (t & 31);
        c->tile_visited_count += 1;
    }
}

static void contact_end(Car* c, int k, int t) {
    c->touching[k][t >> 5] &= ~(1u << (t & 31));
    c->n_touching[k] -= 1;
}

/* b2ContactManager::Collide for the wheel-tile sensor pairs (see header for the simplification) */
static void world_collide(CarEnv* e) {
    for (int ci = 0; ci < e->n_cars; ++ci) {
        Car* c = &e->car[ci];
        for (int k = 0; k < 4; ++k) {
            const Body* w = &c->body[BODY_OF_WHEEL[k]];
            V2 wp[4];
            wheel_world_poly(w, wp);
            float minx = wp[0].x, maxx = wp[0].x, miny = wp[0].y, maxy = wp[0].y;
            for (int i = 1; i < 4; ++i) {
                if (wp[i].x < minx) minx = wp[i].x;
                if (wp[i].x > maxx) maxx = wp[i].x;
                if (wp[i].y < miny) miny = wp[i].y;
                if (wp[i].y > maxy) maxy = wp[i].y;
            }
            const float m = 2.0f * B2_POLYGON_RADIUS;
            for (int t = 0; t < e->n_track; ++t) {
                int was = (c->touching[k][t >> 5] >> (t & 31)) & 1u;
                int now = 0;
                const float* bb = e->tile_aabb[t];
                if (!(bb[0] > maxx + m || bb[2] < minx - m || bb[1] > maxy + m || bb[3] < miny - m))
                    now = car_oracle_polys_touch((const float*)wp, 4, &e->tile_poly[t][0][0], e->tile_n[t]);
                if (now && !was) contact_begin(e, c, k, t);
                else if (!now && was) contact_end(c, k, t);
            }
        }
    }
}

/* b2World::Step(1/FPS, 180, 60) */
static void world_step(CarEnv* e, float dt) {
    world_collide(e);
    float dt_ratio = e->inv_dt0 * dt;
    if (e->n_cars == 2) {
        world_step_two(e->car[0].body, e->car[0].joint, e->car[1].body, e->car[1].joint, &e->contacts, dt, dt_ratio,
                       6 * 30, 2 * 30);
    } else {
        for (int ci = 0; ci < e->n_cars; ++ci) {
            Car* c = &e->car[ci];
            int any_awake = 0;
            for (int i = 0; i < 5; ++i) any_awake |= c->body[i].awake;
            if (any_awake) {
                for (int i = 0; i < 5; ++i) body_set_awake(&c->body[i], 1);   /* island bodies are woken */
                island_solve(c->body, 5, c->joint, 4, dt, dt_ratio, 6 * 30, 2 * 30);
            }
        }
    }
    for (int ci = 0; ci < e->n_cars; ++ci)
        for (int i = 0; i < 5; ++i) { e->car[ci].body[i].force = v2(0.f, 0.f); e->car[ci].body[i].torque = 0.f; }   /* ClearForces */
    e->inv_dt0 = 1.0f / dt;
}

/* tiles from the track: _create_track tail, :399-445 */
static void build_tiles(CarEnv* e) {
    int n = e->n_track;
    for (int i = 0; i < n; ++i) {
        const double* p1 = e->track[i];
        const double* p2 = e->track[(i - 1 + n) % n];
        double b1 = p1[1], x1 = p1[2], y1 = p1[3], b2 = p2[1], x2 = p2[2], y2 = p2[3];
        double v[5][2] = {
            {x1 - TRACK_WIDTH * cos(b1), y1 - TRACK_WIDTH * sin(b1)},
            {x1 - TRACK_WIDTH / 2 * cos(b1 - M_PI / 2), y1 - TRACK_WIDTH / 2 * sin(b1 - M_PI / 2)},
            {x1 + TRACK_WIDTH * cos(b1), y1 + TRACK_WIDTH * sin(b1)},
            {x2 + TRACK_WIDTH * cos(b2), y2 + TRACK_WIDTH * sin(b2)},
            {x2 - TRACK_WIDTH * cos(b2), y2 - TRACK_WIDTH * sin(b2)}};
        float raw[10], hull[16];
        for (int k = 0; k < 5; ++k) {
            raw[2 * k] = (float)v[k][0]; raw[2 * k + 1] = (float)v[k][1];
            e->tile_raw[i][k][0] = raw[2 * k]; e->tile_raw[i][k][1] = raw[2 * k + 1];
            /* render_road_for_observation_map :749-753: (obs_scale * -v + world_size/2), int-truncated by pygame */
            e->tile_map[i][k][0] = (int)(OBS_SCALE * -v[k][0] + 5000.0);
            e->tile_map[i][k][1] = (int)(OBS_SCALE * -v[k][1] + 5000.0);
        }
        int cnt = car_oracle_convex_hull(raw, 5, hull);
        e->tile_n[i] = cnt;
        float minx = 3e38f, miny = 3e38f, maxx = -3e38f, maxy = -3e38f;
        for (int k = 0; k < cnt; ++k) {
            e->tile_poly[i][k][0] = hull[2 * k]; e->tile_poly[i][k][1] = hull[2 * k + 1];
            if (hull[2 * k] < minx) minx = hull[2 * k];
            if (hull[2 * k] > maxx) maxx = hull[2 * k];
            if (hull[2 * k + 1] < miny) miny = hull[2 * k + 1];
            if (hull[2 * k + 1] > maxy) maxy = hull[2 * k + 1];
        }
        e->tile_aabb[i][0] = minx; e->tile_aabb[i][1] = miny; e->tile_aabb[i][2] = maxx; e->tile_aabb[i][3] = maxy;
        if (e->border[i]) {
            double side = signd(b2 - b1);
            double kb[4][2] = {
                {x1 + side * TRACK_WIDTH * cos(b1), y1 + side * TRACK_WIDTH * sin(b1)},
                {x1 + side * (TRACK_WIDTH + BORDER) * cos(b1), y1 + side * (TRACK_WIDTH + BORDER) * sin(b1)},
                {x2 + side * (TRACK_WIDTH + BORDER) * cos(b2), y2 + side * (TRACK_WIDTH + BORDER) * sin(b2)},
                {x2 + side * TRACK_WIDTH * cos(b2), y2 + side * TRACK_WIDTH * sin(b2)}};
            for (int k = 0; k < 4; ++k) {
                e->kerb[i][k][0] = (float)kb[k][0]; e->kerb[i][k][1] = (float)kb[k][1];
                e->kerb_map[i][k][0] = (int)(OBS_SCALE * -kb[k][0] + 5000.0);
                e->kerb_map[i][k][1] = (int)(OBS_SCALE * -kb[k][1] + 5000.0);
            }
        }
    }
}


/* ---- test hooks for the contact restatement (tests/test_oracle_car_contacts.py) ---- */

/* b2CollidePolygons + b2WorldManifold for fixture fa of a body at (xa, ya, angle aa) against fixture fb of a body at
 * (xb, yb, ab); fixtures 0..3 = hull polygons, 4 = wheel box.  out: count, type, normal xy, world points (2 x xy),
 * ids (2).  Returns the point count. */
int car_oracle_collide_fixtures(int fa, float xa, float ya, float aa, int fb, float xb, float yb, float ab, double* out) {
    fixtures_init();
    Xf ta, tb;
    ta.p = v2(xa, ya); ta.q = rot(aa); tb.p = v2(xb, yb); tb.q = rot(ab);
    Manifold m;
    memset(&m, 0, sizeof m);
    collide_polygons(&m, &g_fix_poly[fa], ta, &g_fix_poly[fb], tb);
    V2 normal = v2(0.f, 0.f), pts[2] = {{0.f, 0.f}, {0.f, 0.f}};
    if (m.count) world_manifold(&m, ta, tb, &normal, pts);
    out[0] = m.count; out[1] = m.type; out[2] = normal.x; out[3] = normal.y;
    out[4] = pts[0].x; out[5] = pts[0].y; out[6] = pts[1].x; out[7] = pts[1].y;
    out[8] = m.pt[0].id; out[9] = m.pt[1].id;
    return m.count;
}

/* vertices of fixture f (body-local, hull order); returns the count */
int car_oracle_fixture_polygon(int f, float* out_xy) {
    fixtures_init();
    for (int i = 0; i < g_fix_poly[f].n; ++i) { out_xy[2 * i] = g_fix_poly[f].v[i].x; out_xy[2 * i + 1] = g_fix_poly[f].v[i].y; }
    return g_fix_poly[f].n;
}

/* Two cars in free flight (no wheel forces, no tiles): car k starts at pose[k] = (x, y, angle) with every body
 * moving at vel[k] = (vx, vy); `steps` world steps.  out[t] = total linear momentum (2, float64 sums of m * v),
 * total angular momentum about the origin, touching contacts, manifold points, the two hull positions (4) -> [steps][9]. */
void car_oracle_free_collision(const double* pose, const double* vel, int steps, double* out) {
    Car car[2];
    ContactStore cs;
    memset(&cs, 0, sizeof cs);
    for (int k = 0; k < 2; ++k) {
        car_create(&car[k], pose[3 * k + 2], pose[3 * k], pose[3 * k + 1], 0);
        Body* hull = &car[k].body[HULL_BODY];
        for (int w = 0; w < 4; ++w) {       /* put the wheels where the joints want them (Car.__init__ does not rotate them) */
            Body* b = &car[k].body[BODY_OF_WHEEL[w]];
            V2 anchor = vadd(rmul(hull->q, v2((float)(WHEELPOS[w][0] * SIZE), (float)(WHEELPOS[w][1] * SIZE))), hull->p);
            b->p = anchor; b->c = b->c0 = vadd(rmul(b->q, b->local_center), b->p);
        }
        for (int i = 0; i < 5; ++i) car[k].body[i].v = v2((float)vel[2 * k], (float)vel[2 * k + 1]);
    }
    float inv_dt0 = 0.0f;
    const float dt = 1.0f / FPS;
    for (int t = 0; t < steps; ++t) {
        world_step_two(car[0].body, car[0].joint, car[1].body, car[1].joint, &cs, dt, inv_dt0 * dt, 6 * 30, 2 * 30);
        inv_dt0 = 1.0f / dt;
        double px = 0, py = 0, L = 0;
        for (int k = 0; k < 2; ++k)
            for (int i = 0; i < 5; ++i) {
                const Body* b = &car[k].body[i];
                px += (double)b->mass * b->v.x; py += (double)b->mass * b->v.y;
                L += (double)b->mass * ((double)b->c.x * b->v.y - (double)b->c.y * b->v.x) + (double)b->I * b->w;
            }
        int np = 0;
        double* o = out + 9 * t;
        o[0] = px; o[1] = py; o[2] = L; o[3] = car_oracle_contact_count(&cs, &np); o[4] = np;
        o[5] = car[0].body[HULL_BODY].p.x; o[6] = car[0].body[HULL_BODY].p.y;
        o[7] = car[1].body[HULL_BODY].p.x; o[8] = car[1].body[HULL_BODY].p.y;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* renderer                                                                                     */

/* Observation of player `pi`: get_observation (:622-634) -> camera_update("rgb_array") (:791-812) ->
 * render(mode="internal_rgb_array") (:857-863) = camera_view (:764-789) + Car.draw_for_pygame
 * (car_dynamics.py:284-298) + render_indicators_for_pygame (:645-670); gray = trunc(0.299 R + 0.587 G +
 * 0.114 B) in float64.
 *
 * The reference paints the whole road once per reset into a 10 000 x 10 000 surface at obs_scale px/unit
 * (render_road_for_observation_map :732-755), crops 192 x 192 around the camera, rotates the crop
 * (pygame.transform.rotate) and centre-blits it onto the 96 x 96 screen.  This function evaluates the same
 * pipeline per destination pixel: screen pixel -> rotated-surface pixel -> (16.16 fixed point) source pixel
 * -> road-map pixel (U, V), whose colour is the last polygon in paint order that pygame's scanline fill
 * would have covered it with.  The third-party pieces restated here (pygame 1.9 draw_fillpoly, draw.rect
 * via polygon, transform.rotate, int truncation of float coordinates) are the same restatements as in
 * oracle/ref_shim/pygame, against which tests/golden/car_frames.npz pins this function bit for bit. */
typedef struct { int n; int vx[8], vy[8]; } IPoly;

/* pygame 1.9 draw_fillpoly: is pixel (U, V) filled? */
static int ipoly_covers(const IPoly* p, int U, int V) {
    int miny = p->vy[0], maxy = p->vy[0];
    for (int i = 1; i < p->n; ++i) { if (p->vy[i] < miny) miny = p->vy[i]; if (p->vy[i] > maxy) maxy = p->vy[i]; }
    if (V < miny || V > maxy) return 0;
    int xs[8], m = 0;
    for (int i = 0; i < p->n; ++i) {
        int i1 = i ? i - 1 : p->n - 1;
        int y1 = p->vy[i1], y2 = p->vy[i], x1 = p->vx[i1], x2 = p->vx[i];
        if (y1 > y2) { int t = y1; y1 = y2; y2 = t; t = x1; x1 = x2; x2 = t; }
        else if (y1 == y2) continue;
        if ((V >= y1 && V < y2) || (V == maxy && V > y1 && V <= y2)) xs[m++] = (V - y1) * (x2 - x1) / (y2 - y1) + x1;
    }
    for (int i = 1; i < m; ++i) { int k = xs[i], j = i - 1; while (j >= 0 && xs[j] > k) { xs[j + 1] = xs[j]; --j; } xs[j + 1] = k; }
    for (int k = 0; k + 1 < m; k += 2) if (U >= xs[k] && U <= xs[k + 1]) return 1;
    return 0;
}

static uint8_t gray_of(double r, double g, double b) { return (uint8_t)(r * 0.299 + g * 0.587 + b * 0.114); }

static void fill_ipoly(uint8_t* img, const IPoly* p, uint8_t val) {
    for (int y = 0; y < STATE_H; ++y)
        for (int x = 0; x < STATE_W; ++x)
            if (ipoly_covers(p, x, y)) img[y * STATE_W + x] = val;
}

/* pygame.draw.rect(screen, color, (x, y, w, h)): Rect truncates the floats; 1.9 draws it as the polygon
 * (l, t), (r, t), (r, b), (l, b) with r = x + w - 1, b = y + h - 1 */
static void draw_rect(uint8_t* img, double x, double y, double w, double h, uint8_t val) {
    int X = (int)x, Y = (int)y, W = (int)w, H = (int)h;
    IPoly p;
    p.n = 4;
    p.vx[0] = X; p.vy[0] = Y; p.vx[1] = X + W - 1; p.vy[1] = Y; p.vx[2] = X + W - 1; p.vy[2] = Y + H - 1; p.vx[3] = X; p.vy[3] = Y + H - 1;
    fill_ipoly(img, &p, val);
}

static float f32(double v) { return (float)v; }

static void render_obs(CarEnv* e, int pi) {
    const double obs_scale = OBS_SCALE;
    const Car* me = &e->car[pi];
    const Body* hull = &me->body[HULL_BODY];
    /* camera_update("rgb_array") */
    double angle = hull->a;
    double vx = hull->v.x, vy = hull->v.y;
    if (vx * vx + vy * vy > 0.5 * 0.5) angle = atan2(-vx, +vy);
    float fa = (float)angle, fs = sinf(fa), fc = cosf(fa);
    float camx = hull->p.x + (fc * 0.0f - fs * 16.0f), camy = hull->p.y + (fs * 0.0f + fc * 16.0f);   /* b2Vec2 arithmetic */
    /* camera_view(mode="rgb_array") */
    double pos0 = obs_scale * -(double)camx + 5000.0, pos1 = obs_scale * -(double)camy + 5000.0;
    int rx = (int)(pos0 - STATE_W), ry = (int)(pos1 - STATE_H);
    const int sw = 2 * STATE_W, sh = 2 * STATE_H;
    /* pygame.transform.rotate(view, 57.295779513 * angle) */
    double rad = (57.295779513 * angle) * .01745329251994329;
    double sangle = sin(rad), cangle = cos(rad);
    double cx = cangle * sw, cy = cangle * sh, sx = sangle * sw, sy = sangle * sh;
    double m1 = fmax(fmax(fmax(fabs(cx + sy), fabs(cx - sy)), fabs(-cx + sy)), fabs(-cx - sy));
    double m2 = fmax(fmax(fmax(fabs(sx + cy), fabs(sx - cy)), fabs(-sx + cy)), fabs(-sx - cy));
    int nx = (int)m1, ny = (int)m2;
    int cyi = ny / 2, xd = (sw - nx) * 32768, yd = (sh - ny) * 32768;
    int isin = (int)(sangle * 65536), icos = (int)(cangle * 65536);
    int ax = (nx * 32768) - (int)(cangle * (double)((nx - 1) * 32768));
    int ay = (ny * 32768) - (int)(sangle * (double)((nx - 1) * 32768));
    int xmaxval = (sw << 16) - 1, ymaxval = (sh << 16) - 1;
    int bx = -(nx >> 1) + STATE_W / 2, by = -(ny >> 1) + STATE_H / 2;   /* blit position of the rotated surface */
    uint8_t* img = e->obs[pi];
    const uint8_t g_grass = gray_of((int)(0.4 * 255), (int)(0.8 * 255), (int)(0.4 * 255));
    const uint8_t g_check = gray_of((int)(0.4 * 255), (int)(0.9 * 255), (int)(0.4 * 255));
    const double k = PLAYFIELD / 20.0;
    /* candidate tiles: anything that can reach the visible window (central 96 x 96 of the rotated crop) */
    static __thread int cand[MAX_TRACK];
    int n_cand = 0;
    const double reach = (48.0 * 1.4142135623730951 + 4.0) / obs_scale + 2.0 * TRACK_WIDTH + BORDER + TRACK_DETAIL_STEP;
    for (int t = e->n_track - 1; t >= 0; --t) {   /* paint order: tiles are created from i = n-1 down to 0 */
        double dx = e->track[t][2] - camx, dy = e->track[t][3] - camy;
        if (dx * dx + dy * dy <= reach * reach) cand[n_cand++] = t;
    }
    for (int Y = 0; Y < STATE_H; ++Y) {
        for (int X = 0; X < STATE_W; ++X) {
            int x = X - bx, y = Y - by;    /* pixel of the rotated surface under this screen pixel */
            uint8_t val = 0;               /* surfaces start black */
            if (x >= 0 && y >= 0 && x < nx && y < ny) {
                int dx = (ax + (isin * (cyi - y))) + xd + icos * x;
                int dy = (ay - (icos * (cyi - y))) + yd + isin * x;
                if (!(dx < 0 || dy < 0 || dx > xmaxval || dy > ymaxval)) {
                    int U = rx + (dx >> 16), V = ry + (dy >> 16);
                    val = g_grass;
                    for (int gx = -20; gx < 20; gx += 2)       /* checker squares, :735-746 */
                        for (int gy = -20; gy < 20; gy += 2) {
                            IPoly q;
                            q.n = 4;
                            double qx[4] = {k * gx + k, k * gx + 0, k * gx + 0, k * gx + k};
                            double qy[4] = {k * gy + 0, k * gy + 0, k * gy + k, k * gy + k};
                            for (int i = 0; i < 4; ++i) { q.vx[i] = (int)(obs_scale * -qx[i] + 5000.0); q.vy[i] = (int)(obs_scale * -qy[i] + 5000.0); }
                            if (U < q.vx[0] - 1 || U > q.vx[1] + 1 || V < q.vy[2] - 1 || V > q.vy[0] + 1) continue;
                            if (ipoly_covers(&q, U, V)) val = g_check;
                        }
                    for (int c = 0; c < n_cand; ++c) {
                        int t = cand[c];
                        IPoly q;
                        q.n = 5;
                        for (int i = 0; i < 5; ++i) { q.vx[i] = e->tile_map[t][i][0]; q.vy[i] = e->tile_map[t][i][1]; }
                        if (ipoly_covers(&q, U, V)) { double col = (int)(255 * (0.4 + 0.01 * (t % 3))); val = gray_of(col, col, col); }
                        if (e->border[t]) {
                            q.n = 4;
                            for (int i = 0; i < 4; ++i) { q.vx[i] = e->kerb_map[t][i][0]; q.vy[i] = e->kerb_map[t][i][1]; }
                            if (ipoly_covers(&q, U, V)) val = (t % 2 == 0) ? gray_of(255, 255, 255) : gray_of(255, 0, 0);
                        }
                    }
                }
            }
            img[Y * STATE_W + X] = val;
        }
    }
    /* cars: for k in cars: for obj in wheels + [hull]: for f in obj.fixtures: polygon (draw_for_pygame) */
    {
        float ta = f32(-angle), ts = sinf(ta), tc = cosf(ta);     /* tmp.angle = -angle */
        for (int ci = 0; ci < e->n_cars; ++ci) {
            const Car* cr = &e->car[ci];
            for (int part = 0; part < 8; ++part) {
                const Body* b = part < 4 ? &cr->body[BODY_OF_WHEEL[part]] : &cr->body[HULL_BODY];
                IPoly q;
                float lv[8][2];
                if (part < 4) {
                    const float hw = (float)(WHEEL_W * SIZE), hr = (float)(WHEEL_R * SIZE);
                    float box[4][2] = {{+hw, -hr}, {+hw, +hr}, {-hw, +hr}, {-hw, -hr}};
                    q.n = 4;
                    memcpy(lv, box, sizeof box);
                } else {
                    int f = part - 4;
                    q.n = HULL_COUNTS[f];
                    for (int i = 0; i < q.n; ++i) { lv[i][0] = (float)(HULL_POLYS[f][i][0] * SIZE); lv[i][1] = (float)(HULL_POLYS[f][i][1] * SIZE); }
                }
                for (int i = 0; i < q.n; ++i) {
                    float wx = (b->q.c * lv[i][0] - b->q.s * lv[i][1]) + b->p.x, wy = (b->q.s * lv[i][0] + b->q.c * lv[i][1]) + b->p.y;
                    float ox = wx - camx, oy = wy - camy;
                    float rx2 = (tc * ox - ts * oy) + 0.0f, ry2 = (ts * ox + tc * oy) + 0.0f;
                    float px = f32((double)rx2 * -obs_scale) + (float)(STATE_W / 2.0), py = f32((double)ry2 * -obs_scale) + (float)(STATE_H / 2.0);
                    q.vx[i] = (int)px; q.vy[i] = (int)py;
                }
                uint8_t col = part < 4 ? gray_of(0, 0, 0) : (ci == pi ? gray_of(0.8 * 255, 0, 0) : gray_of(0, 0, 255));
                fill_ipoly(img, &q, col);
            }
        }
    }
    /* HUD: render_indicators_for_pygame(width=96, height=96, scale=5) */
    const double W = STATE_W, H = STATE_H, s = W / 40.0, h = H / 40.0;
    double true_speed = sqrt((double)hull->v.x * hull->v.x + (double)hull->v.y * hull->v.y);
    draw_rect(img, 0, H - 4 * h, W, 4 * h * 1000, gray_of(0, 0, 0));
    draw_rect(img, 5 * s, H - h, s, h * (-0.02 * true_speed), gray_of(0, 0, 255));
    for (int wk = 0; wk < 4; ++wk)
        draw_rect(img, (7 + wk) * s, H - h, s, h * (-0.01 * me->omega[wk]), wk < 2 ? gray_of(0.0, 0, 255) : gray_of((int)(0.2 * 255), 0, 255));
    {
        const RevJoint* j0 = &me->joint[JOINT_OF_WHEEL[0]];
        double ja = (double)((me->body[BODY_OF_WHEEL[0]].a - hull->a) - j0->reference_angle);
        draw_rect(img, 20 * s, H - 2 * h, s * (10.0 * ja), 2 * h, gray_of(0, 255, 0));
        draw_rect(img, 30 * s, H - 2 * h, s * (0.8 * (double)hull->w), 2 * h, gray_of(255, 0, 0));
    }
    if (e->glyphs) {   /* draw_text("%05.0f" % reward) at (W/100, H - H/20), white, 5 px non-AA glyphs */
        char txt[32];
        snprintf(txt, sizeof txt, "%05.0f", me->reward);
        int pen = (int)(W / 100), y0 = (int)(H - H / 20);
        for (int i = 0; txt[i]; ++i) {
            int gi = txt[i] == '-' ? 10 : (txt[i] >= '0' && txt[i] <= '9' ? txt[i] - '0' : -1);
            if (gi < 0) continue;
            for (int gy = 0; gy < 8; ++gy)
                for (int gx = 0; gx < 4; ++gx)
                    if (e->glyphs[(gi * 8 + gy) * 4 + gx]) {
                        int px = pen + gx, py = y0 + gy;
                        if (px >= 0 && px < STATE_W && py >= 0 && py < STATE_H) img[py * STATE_W + px] = gray_of(255, 255, 255);
                    }
            pen += e->glyphs[11 * 8 * 4 + gi];
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* env API (ctypes)                                                                             */

CarEnv* car_oracle_create(int n_cars, int action_repeat, const uint8_t* glyphs) {
    CarEnv* e = (CarEnv*)calloc(1, sizeof(CarEnv));
    e->n_cars = n_cars;
    e->action_repeat = action_repeat > 0 ? action_repeat : 1;
    e->glyphs = glyphs;
    return e;
}
void car_oracle_destroy(CarEnv* e) { free(e); }

/* CarRacing.reset with an injected track ([n][4] alpha,beta,x,y as in the JSON track format
 * :376-381) and birth-place permutation (np.random.shuffle(arange(num_player)), :508-509). */
void car_oracle_reset(CarEnv* e, const double* track, const int* border, int n_track, const int* birth_place) {
    e->n_track = n_track;
    memcpy(e->track, track, sizeof(double) * 4 * (size_t)n_track);
    memcpy(e->border, border, sizeof(int) * (size_t)n_track);
    build_tiles(e);
    for (int k = 0; k < e->n_cars; ++k)
        car_create(&e->car[k], e->track[0][1], e->track[0][2], e->track[0][3], birth_place ? birth_place[k] : k);
    e->step_count = 0;
    e->inv_dt0 = 0.0f;
    memset(&e->contacts, 0, sizeof e->contacts);
    if (!e->lazy_render) for (int k = 0; k < e->n_cars; ++k) render_obs(e, k);   /* return self.step(None)[0] */
}

/* CarRacing.step, :542-620.  actions [n_cars][2]; out: step_rewards[n_cars], done[n_cars] */
void car_oracle_step(CarEnv* e, const double* actions, double* step_rewards, int* done, int* num_steps) {
    for (int k = 0; k < e->n_cars; ++k) {   /* process_action, :527-540 */
        double a0 = fmax(fmin(actions[2 * k], 1), -1), a1 = fmax(fmin(actions[2 * k + 1], 1), -1), a2;
        if (a1 > 0) a2 = 0; else { a2 = a1; a1 = 0; }
        car_controls(&e->car[k], -a0, fabs(a1), fabs(a2));
        step_rewards[k] = 0.0;
    }
    for (int rep = 0; rep < e->action_repeat; ++rep) {
        for (int k = 0; k < e->n_cars; ++k) {
            Car* c = &e->car[k];
            if (c->done) continue;
            car_step(c, 1.0 / FPS);
            c->reward -= 0.1 / e->action_repeat;
            step_rewards[k] += c->reward - c->prev_reward;
            c->prev_reward = c->reward;
            double x = c->body[HULL_BODY].p.x, y = c->body[HULL_BODY].p.y;
            if (c->tile_visited_count == e->n_track) c->done = 1;
            if (fabs(x) > PLAYFIELD || fabs(y) > PLAYFIELD) c->done = 1;
            if (e->step_count > 1000) c->done = 1;
        }
        world_step(e, 1.0f / FPS);
        e->step_count += 1;
    }
    for (int k = 0; k < e->n_cars; ++k) { if (!e->lazy_render) render_obs(e, k); done[k] = e->car[k].done; }
    *num_steps = e->step_count;
}

void car_oracle_set_lazy_render(CarEnv* e, int lazy) { e->lazy_render = lazy; }

/* Debug / tests (mirror of crl_car_set_state, same fp32 operations in the same order): put every car into the state
 * state[car][24] laid out as car_oracle_get_state writes it.  Used: hull x, y, angle, vx, vy, w; per wheel joint angle,
 * omega, gas; reward.  The wheels are placed on their joint anchors and move rigidly with the hull; joint impulses,
 * car-car manifolds and the wheels' tile sets are cleared. */
void car_oracle_set_state(CarEnv* e, const double* in) {
    for (int k = 0; k < e->n_cars; ++k) {
        Car* c = &e->car[k];
        const double* s = in + 24 * k;
        Body* h = &c->body[HULL_BODY];
        const float a = (float)s[2];
        const Rot q = rot(a);
        const V2 org = v2((float)s[0], (float)s[1]);
        const V2 com = vadd(org, rmul(q, h->local_center));
        const V2 v0 = v2((float)s[3], (float)s[4]);
        const float w0 = (float)s[5];
        h->a = h->a0 = a; h->q = q; h->p = org; h->c = h->c0 = com; h->v = v0; h->w = w0; h->awake = 1; h->sleep_time = 0.f;
        for (int w = 0; w < 4; ++w) {
            Body* b = &c->body[BODY_OF_WHEEL[w]];
            const V2 wp = vadd(org, rmul(q, v2((float)(WHEELPOS[w][0] * SIZE), (float)(WHEELPOS[w][1] * SIZE))));
            const V2 d = vsub(wp, com);
            b->a = b->a0 = a + (float)s[6 + 4 * w]; b->q = rot(b->a); b->p = wp; b->c = b->c0 = wp;
            b->v = vadd(v0, v2(-w0 * d.y, w0 * d.x)); b->w = w0; b->awake = 1; b->sleep_time = 0.f;
            c->omega[w] = s[7 + 4 * w];
            if (w >= 2) c->gas[w] = s[8 + 4 * w];
            c->n_touching[w] = 0;
            memset(c->touching[w], 0, sizeof c->touching[w]);
        }
        for (int j = 0; j < 4; ++j) {
            RevJoint* r = &c->joint[j];
            r->impulse[0] = r->impulse[1] = r->impulse[2] = 0.f; r->motor_impulse = 0.f; r->limit_state = 0; r->motor_speed = 0.f;
        }
        for (int w = 0; w < 4; ++w) { c->brake[w] = 0.0; c->steer[w] = 0.0; }
        c->reward = c->prev_reward = s[22];
    }
    memset(&e->contacts, 0, sizeof e->contacts);
}
/* touching car-car contacts / manifold points after the last step (diagnostics for the tests) */
int car_oracle_env_contacts(const CarEnv* e, int* n_points) { return car_oracle_contact_count(&e->contacts, n_points); }
const uint8_t* car_oracle_obs(CarEnv* e, int player) {
    if (e->lazy_render) render_obs(e, player);
    return e->obs[player];
}

/* state[car][24]: hull x, y, angle, vx, vy, w; wheel k: angle_k, omega_k (python), gas_k ; reward; tiles; done */
void car_oracle_get_state(const CarEnv* e, double* out) {
    for (int k = 0; k < e->n_cars; ++k) {
        const Car* c = &e->car[k];
        const Body* h = &c->body[HULL_BODY];
        double* s = out + 24 * k;
        s[0] = h->p.x; s[1] = h->p.y; s[2] = h->a; s[3] = h->v.x; s[4] = h->v.y; s[5] = h->w;
        for (int w = 0; w < 4; ++w) {
            const Body* b = &c->body[BODY_OF_WHEEL[w]];
            s[6 + 4 * w] = b->a - h->a; s[7 + 4 * w] = c->omega[w]; s[8 + 4 * w] = c->gas[w]; s[9 + 4 * w] = c->n_touching[w];
        }
        s[22] = c->reward; s[23] = c->tile_visited_count;
    }
}
