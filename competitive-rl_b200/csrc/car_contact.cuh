// car_contact.cuh -- car-car contacts of cCarRacingDouble (included by car_physics.cu only).
//
// Replaces, for the two cars of one env, what box2d-py ~=2.3.5 does inside b2World::Step when the
// reference's hull / wheel fixtures of different cars meet (car_dynamics.py:63-68 hull polygons,
// category 0x0001 mask 0xFFFF; :94-96 wheels, category 0x0020 mask 0x001 -> hull-hull and wheel-hull
// pairs collide, wheel-wheel pairs do not): b2CollidePolygons (2.3.0: b2FindMaxSeparation hill climb,
// b2FindIncidentEdge, b2ClipSegmentToLine), b2Contact::Update (impulses carried over by feature id),
// b2ContactSolver (warm start, friction then normal, 2-point block solver, Baumgarte position
// correction).  Restated from the published algorithm like the rest of the mini Box2D; the CPU
// restatement it is tested against is oracle/car_oracle.c ("mini Box2D, part 2"), which also lists
// the stated deviations (canonical contact order, manifolds evaluated for every allowed pair).
//
// Execution model: the two cars of an env sit on adjacent lanes.  Their bodies are exchanged through
// a small shared-memory block; the player-0 lane runs the (rare, sequential) contact code below on
// that block while the joints of both cars stay in registers of their own lanes.
#pragma once

namespace crl {

#define B2_VELOCITY_THRESHOLD 1.0f
#define B2_BAUMGARTE 0.2f
#define B2_MAX_LINEAR_CORRECTION 0.2f
#define B2_EPSILON 1.1920929e-07f

struct Xf { F2 p; Rot q; };
__device__ __forceinline__ F2 xf_mul(const Xf& t, F2 v) { return f2((t.q.c * v.x - t.q.s * v.y) + t.p.x, (t.q.s * v.x + t.q.c * v.y) + t.p.y); }
__device__ __forceinline__ F2 xf_mulT(const Xf& t, F2 v) {
    const float px = v.x - t.p.x, py = v.y - t.p.y;
    return f2(t.q.c * px + t.q.s * py, -t.q.s * px + t.q.c * py);
}
__device__ __forceinline__ F2 rmulT(Rot q, F2 v) { return f2(q.c * v.x + q.s * v.y, -q.s * v.x + q.c * v.y); }
__device__ __forceinline__ F2 cross_vs(F2 a, float s) { return f2(s * a.y, -s * a.x); }
__device__ __forceinline__ F2 neg(F2 a) { return f2(-a.x, -a.y); }
__device__ __forceinline__ F2 normalized(F2 a) {
    const float len = sqrtf(a.x * a.x + a.y * a.y);
    if (len < B2_EPSILON) return a;
    const float inv = 1.0f / len;
    return f2(a.x * inv, a.y * inv);
}
__device__ __forceinline__ Xf xf_of(F2 c, float a, F2 lc) {
    Xf t;
    t.q = make_rot(a);
    t.p = c - rmul(t.q, lc);
    return t;
}

// body-local polygon `s` (0..3 hull fixtures, 4 wheel box) from the constants block
struct PolyRef { const float *vx, *vy, *nx, *ny; int n; F2 centroid; };
__device__ __forceinline__ PolyRef poly_ref(const CarHullConst* K, int s) {
    PolyRef p;
    p.vx = K->fix_vx[s]; p.vy = K->fix_vy[s]; p.nx = K->fix_nx[s]; p.ny = K->fix_ny[s];
    p.n = K->fix_n[s]; p.centroid = f2(K->fix_cx[s], K->fix_cy[s]);
    return p;
}

__device__ float edge_separation(const PolyRef& p1, const Xf& xf1, int edge1, const PolyRef& p2, const Xf& xf2) {
    const F2 n1w = rmul(xf1.q, f2(p1.nx[edge1], p1.ny[edge1]));
    const F2 n1 = rmulT(xf2.q, n1w);
    int index = 0;
    float min_dot = 3.402823466e+38f;
    for (int i = 0; i < p2.n; ++i) {
        const float d = dot(f2(p2.vx[i], p2.vy[i]), n1);
        if (d < min_dot) { min_dot = d; index = i; }
    }
    const F2 v1 = xf_mul(xf1, f2(p1.vx[edge1], p1.vy[edge1])), v2 = xf_mul(xf2, f2(p2.vx[index], p2.vy[index]));
    return dot(v2 - v1, n1w);
}

__device__ float find_max_separation(int* edge_index, const PolyRef& p1, const Xf& xf1, const PolyRef& p2, const Xf& xf2) {
    const int count1 = p1.n;
    const F2 d = xf_mul(xf2, p2.centroid) - xf_mul(xf1, p1.centroid);
    const F2 d_local1 = rmulT(xf1.q, d);
    int edge = 0;
    float max_dot = -3.402823466e+38f;
    for (int i = 0; i < count1; ++i) {
        const float dt = dot(f2(p1.nx[i], p1.ny[i]), d_local1);
        if (dt > max_dot) { max_dot = dt; edge = i; }
    }
    float s = edge_separation(p1, xf1, edge, p2, xf2);
    const int prev_edge = edge - 1 >= 0 ? edge - 1 : count1 - 1;
    const float s_prev = edge_separation(p1, xf1, prev_edge, p2, xf2);
    const int next_edge = edge + 1 < count1 ? edge + 1 : 0;
    const float s_next = edge_separation(p1, xf1, next_edge, p2, xf2);
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

struct ClipVertex { F2 v; uint32_t id; };
__device__ __forceinline__ uint32_t cf_id(int ia, int ib, int ta, int tb) {
    return (uint32_t)ia | ((uint32_t)ib << 8) | ((uint32_t)ta << 16) | ((uint32_t)tb << 24);
}

__device__ int clip_segment_to_line(ClipVertex* out, const ClipVertex* in, F2 normal, float offset, int vertex_index_a) {
    int n_out = 0;
    const float d0 = dot(normal, in[0].v) - offset, d1 = dot(normal, in[1].v) - offset;
    if (d0 <= 0.0f) out[n_out++] = in[0];
    if (d1 <= 0.0f) out[n_out++] = in[1];
    if (d0 * d1 < 0.0f) {
        const float interp = d0 / (d0 - d1);
        out[n_out].v = in[0].v + interp * (in[1].v - in[0].v);
        out[n_out].id = cf_id(vertex_index_a, (in[0].id >> 8) & 0xff, 0, 1);
        ++n_out;
    }
    return n_out;
}

// b2CollidePolygons (2.3.0): fills the manifold fields of `m` (count, type, local normal / point, points, ids)
__device__ void collide_polygons(CarContact* m, const PolyRef& pa, const Xf& xfa, const PolyRef& pb, const Xf& xfb) {
    m->count = 0;
    const float total_radius = B2_POLYGON_RADIUS + B2_POLYGON_RADIUS;
    int edge_a = 0, edge_b = 0;
    const float sep_a = find_max_separation(&edge_a, pa, xfa, pb, xfb);
    if (sep_a > total_radius) return;
    const float sep_b = find_max_separation(&edge_b, pb, xfb, pa, xfa);
    if (sep_b > total_radius) return;
    const float k_rel = 0.98f, k_abs = 0.001f;
    const bool flip = sep_b > k_rel * sep_a + k_abs;
    const PolyRef& p1 = flip ? pb : pa;
    const PolyRef& p2 = flip ? pa : pb;
    const Xf& xf1 = flip ? xfb : xfa;
    const Xf& xf2 = flip ? xfa : xfb;
    const int edge1 = flip ? edge_b : edge_a;
    m->type = flip ? 1 : 0;
    ClipVertex incident[2];
    {
        const F2 n1 = rmulT(xf2.q, rmul(xf1.q, f2(p1.nx[edge1], p1.ny[edge1])));
        int index = 0;
        float min_dot = 3.402823466e+38f;
        for (int i = 0; i < p2.n; ++i) {
            const float d = dot(n1, f2(p2.nx[i], p2.ny[i]));
            if (d < min_dot) { min_dot = d; index = i; }
        }
        const int i1 = index, i2 = i1 + 1 < p2.n ? i1 + 1 : 0;
        incident[0].v = xf_mul(xf2, f2(p2.vx[i1], p2.vy[i1])); incident[0].id = cf_id(edge1, i1, 1, 0);
        incident[1].v = xf_mul(xf2, f2(p2.vx[i2], p2.vy[i2])); incident[1].id = cf_id(edge1, i2, 1, 0);
    }
    const int iv1 = edge1, iv2 = edge1 + 1 < p1.n ? edge1 + 1 : 0;
    F2 v11 = f2(p1.vx[iv1], p1.vy[iv1]), v12 = f2(p1.vx[iv2], p1.vy[iv2]);
    const F2 local_tangent = normalized(v12 - v11);
    const F2 local_normal = cross_vs(local_tangent, 1.0f);
    const F2 plane_point = 0.5f * (v11 + v12);
    const F2 tangent = rmul(xf1.q, local_tangent);
    const F2 normal = cross_vs(tangent, 1.0f);
    v11 = xf_mul(xf1, v11); v12 = xf_mul(xf1, v12);
    const float front_offset = dot(normal, v11);
    const float side_offset1 = -dot(tangent, v11) + total_radius;
    const float side_offset2 = dot(tangent, v12) + total_radius;
    ClipVertex clip1[2], clip2[2];
    if (clip_segment_to_line(clip1, incident, neg(tangent), side_offset1, iv1) < 2) return;
    if (clip_segment_to_line(clip2, clip1, tangent, side_offset2, iv2) < 2) return;
    m->lnx = local_normal.x; m->lny = local_normal.y; m->lpx = plane_point.x; m->lpy = plane_point.y;
    int count = 0;
    for (int i = 0; i < 2; ++i) {
        const float separation = dot(normal, clip2[i].v) - front_offset;
        if (separation <= total_radius) {
            const F2 lp = xf_mulT(xf2, clip2[i].v);
            m->px[count] = lp.x; m->py[count] = lp.y;
            uint32_t id = clip2[i].id;
            if (flip) id = cf_id((id >> 8) & 0xff, id & 0xff, (id >> 24) & 0xff, (id >> 16) & 0xff);
            m->id[count] = id;
            ++count;
        }
    }
    m->count = (uint8_t)count;
}

__device__ __forceinline__ F2 body_lc(const CarHullConst* K, int b) { return (b % 5 == 0) ? f2(K->hull_lcx, K->hull_lcy) : f2(0.f, 0.f); }
__device__ __forceinline__ float body_inv_mass(const CarHullConst* K, int b) { return (b % 5 == 0) ? K->hull_inv_mass : K->wheel_inv_mass; }
__device__ __forceinline__ float body_inv_I(const CarHullConst* K, int b) { return (b % 5 == 0) ? K->hull_inv_I : K->wheel_inv_I; }

// b2ContactManager::Collide for the 48 allowed fixture pairs of the two cars.  pose[b] = (cx, cy, angle) of
// body b (0..4 car 0: hull, wheels 0..3; 5..9 car 1).  Rewrites recs[0..n) in canonical pair order, carrying
// impulses over from the previous step's records; returns the number of touching contacts.
__device__ __noinline__ int car_contacts_collide(const CarHullConst* K, const float (*pose)[3], CarContact* recs, int n_old,
                                                 int* overflow) {
    Xf xf[10];
    for (int b = 0; b < 10; ++b) xf[b] = xf_of(f2(pose[b][0], pose[b][1]), pose[b][2], body_lc(K, b));
    // world centroids of the 8 fixtures of each car: quick reject of pairs farther apart than their bounding circles
    // plus 0.15 (a manifold needs a face separation <= 0.02; the car polygons have no corner sharper than ~87 degrees,
    // so a true distance of 0.15 leaves a face separation >= 0.1).  Rejecting is the same as an empty manifold.
    F2 wc[16];
    for (int car = 0; car < 2; ++car)
        for (int f = 0; f < 8; ++f) {
            const int b = 5 * car + (f < 4 ? 0 : f - 3), sh = f < 4 ? f : 4;
            wc[8 * car + f] = xf_mul(xf[b], f2(K->fix_cx[sh], K->fix_cy[sh]));
        }
    CarContact out[CAR_MAX_CONTACTS];
    int n_new = 0, pair = 0;
    for (int fa = 0; fa < 8; ++fa)
        for (int fb = 0; fb < 8; ++fb) {
            if (fa >= 4 && fb >= 4) continue;
            const int this_pair = pair++;
            const int ba = fa < 4 ? 0 : fa - 3, bb = 5 + (fb < 4 ? 0 : fb - 3);
            const int sa = fa < 4 ? fa : 4, sb = fb < 4 ? fb : 4;
            const F2 d = wc[8 + fb] - wc[fa];
            const float reach = K->fix_cradius[sa] + K->fix_cradius[sb] + 0.15f;
            if (dot(d, d) > reach * reach) continue;
            CarContact m;
            collide_polygons(&m, poly_ref(K, sa), xf[ba], poly_ref(K, sb), xf[bb]);
            if (m.count == 0) continue;
            if (n_new >= CAR_MAX_CONTACTS) { atomicAdd(overflow, 1); continue; }
            m.pair = (uint8_t)this_pair; m.ia = (uint8_t)ba; m.ib = (uint8_t)bb; m.vcount = m.count;
            for (int i = 0; i < m.count; ++i) { m.ni[i] = 0.f; m.ti[i] = 0.f; }
            for (int o = 0; o < n_old; ++o) {
                if (recs[o].pair != this_pair) continue;
                for (int i = 0; i < m.count; ++i)
                    for (int j = 0; j < recs[o].count; ++j)
                        if (recs[o].id[j] == m.id[i]) { m.ni[i] = recs[o].ni[j]; m.ti[i] = recs[o].ti[j]; break; }
            }
            out[n_new++] = m;
        }
    for (int k = 0; k < n_new; ++k) recs[k] = out[k];
    return n_new;
}

// b2WorldManifold::Initialize
__device__ void world_manifold(const CarContact& m, const Xf& xfa, const Xf& xfb, F2* normal, F2* points) {
    const float ra = B2_POLYGON_RADIUS, rb = B2_POLYGON_RADIUS;
    if (m.type == 0) {
        *normal = rmul(xfa.q, f2(m.lnx, m.lny));
        const F2 plane = xf_mul(xfa, f2(m.lpx, m.lpy));
        for (int i = 0; i < m.count; ++i) {
            const F2 clip = xf_mul(xfb, f2(m.px[i], m.py[i]));
            const F2 ca = clip + (ra - dot(clip - plane, *normal)) * (*normal);
            const F2 cb = clip - rb * (*normal);
            points[i] = 0.5f * (ca + cb);
        }
    } else {
        *normal = rmul(xfb.q, f2(m.lnx, m.lny));
        const F2 plane = xf_mul(xfb, f2(m.lpx, m.lpy));
        for (int i = 0; i < m.count; ++i) {
            const F2 clip = xf_mul(xfa, f2(m.px[i], m.py[i]));
            const F2 cb = clip + (rb - dot(clip - plane, *normal)) * (*normal);
            const F2 ca = clip - ra * (*normal);
            points[i] = 0.5f * (ca + cb);
        }
        *normal = neg(*normal);
    }
}

// b2ContactSolver constructor + InitializeVelocityConstraints + WarmStart, contact by contact (restitution is 0,
// so the only velocity-dependent term of the initialisation, the restitution bias, is 0 and the order is free).
// pose / vel: the shared-memory body block ((cx, cy, a) and (vx, vy, w) of the 10 bodies).
__device__ __noinline__ void car_contacts_init(const CarHullConst* K, CarContact* recs, int n, const float (*pose)[3],
                                               float (*vel)[3], float dt_ratio) {
    for (int k = 0; k < n; ++k) {
        CarContact c = recs[k];
        const int ia = c.ia, ib = c.ib;
        const float mA = body_inv_mass(K, ia), iA = body_inv_I(K, ia), mB = body_inv_mass(K, ib), iB = body_inv_I(K, ib);
        const F2 cA = f2(pose[ia][0], pose[ia][1]), cB = f2(pose[ib][0], pose[ib][1]);
        F2 vA = f2(vel[ia][0], vel[ia][1]), vB = f2(vel[ib][0], vel[ib][1]);
        float wA = vel[ia][2], wB = vel[ib][2];
        const Xf xfa = xf_of(cA, pose[ia][2], body_lc(K, ia)), xfb = xf_of(cB, pose[ib][2], body_lc(K, ib));
        F2 normal, pts[2];
        world_manifold(c, xfa, xfb, &normal, pts);
        c.nx = normal.x; c.ny = normal.y;
        const F2 tangent = cross_vs(normal, 1.0f);
        F2 rA[2], rB[2];
        for (int j = 0; j < c.count; ++j) {
            c.ni[j] = dt_ratio * c.ni[j];
            c.ti[j] = dt_ratio * c.ti[j];
            rA[j] = pts[j] - cA; rB[j] = pts[j] - cB;
            c.rAx[j] = rA[j].x; c.rAy[j] = rA[j].y; c.rBx[j] = rB[j].x; c.rBy[j] = rB[j].y;
            const float rnA = cross(rA[j], normal), rnB = cross(rB[j], normal);
            const float k_normal = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
            c.nmass[j] = k_normal > 0.0f ? 1.0f / k_normal : 0.0f;
            const float rtA = cross(rA[j], tangent), rtB = cross(rB[j], tangent);
            const float k_tangent = mA + mB + iA * rtA * rtA + iB * rtB * rtB;
            c.tmass[j] = k_tangent > 0.0f ? 1.0f / k_tangent : 0.0f;
        }
        c.vcount = c.count;
        if (c.count == 2) {
            const float rn1A = cross(rA[0], normal), rn1B = cross(rB[0], normal);
            const float rn2A = cross(rA[1], normal), rn2B = cross(rB[1], normal);
            const float k11 = mA + mB + iA * rn1A * rn1A + iB * rn1B * rn1B;
            const float k22 = mA + mB + iA * rn2A * rn2A + iB * rn2B * rn2B;
            const float k12 = mA + mB + iA * rn1A * rn2A + iB * rn1B * rn2B;
            const float k_max_cond = 1000.0f;
            if (k11 * k11 < k_max_cond * (k11 * k22 - k12 * k12)) {
                c.k11 = k11; c.k12 = k12; c.k22 = k22;
                float det = k11 * k22 - k12 * k12;
                if (det != 0.0f) det = 1.0f / det;
                c.nm00 = det * k22; c.nm10 = -det * k12; c.nm01 = -det * k12; c.nm11 = det * k11;
            } else {
                c.vcount = 1;
            }
        }
        for (int j = 0; j < c.vcount; ++j) {   // WarmStart
            const F2 P = c.ni[j] * normal + c.ti[j] * tangent;
            wA -= iA * cross(rA[j], P);
            vA = vA - mA * P;
            wB += iB * cross(rB[j], P);
            vB = vB + mB * P;
        }
        vel[ia][0] = vA.x; vel[ia][1] = vA.y; vel[ia][2] = wA;
        vel[ib][0] = vB.x; vel[ib][1] = vB.y; vel[ib][2] = wB;
        recs[k] = c;
    }
}

// A contact record by value: ten 16-byte loads issued back to back.  (Field-by-field access through the pointer makes
// every load wait for the previous store: the compiler cannot rule out that `vel` / `pose` alias the record.)
__device__ __forceinline__ CarContact load_contact(const CarContact* src) {
    static_assert(sizeof(CarContact) == 160, "CarContact is ten uint4");
    CarContact c;
    uint4* d = reinterpret_cast<uint4*>(&c);
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
#pragma unroll
    for (int i = 0; i < 10; ++i) d[i] = s4[i];
    return c;
}

// one friction (tangent) constraint of point J
template <int J>
__device__ __forceinline__ void contact_tangent(CarContact& c, F2 tangent, float friction, float mA, float iA, float mB, float iB,
                                                F2& vA, float& wA, F2& vB, float& wB) {
    const F2 rA = f2(c.rAx[J], c.rAy[J]), rB = f2(c.rBx[J], c.rBy[J]);
    const F2 dv = ((vB + cross_sv(wB, rB)) - vA) - cross_sv(wA, rA);
    const float vt = dot(dv, tangent) - 0.0f;
    float lambda = c.tmass[J] * (-vt);
    const float max_friction = friction * c.ni[J];
    const float new_impulse = clampf(c.ti[J] + lambda, -max_friction, max_friction);
    lambda = new_impulse - c.ti[J];
    c.ti[J] = new_impulse;
    const F2 P = lambda * tangent;
    vA = vA - mA * P;
    wA -= iA * cross(rA, P);
    vB = vB + mB * P;
    wB += iB * cross(rB, P);
}

// b2ContactSolver::SolveVelocityConstraints over all contacts of the env (one velocity iteration)
__device__ __noinline__ void car_contacts_solve_velocity(const CarHullConst* K, CarContact* recs, int n, float (*vel)[3]) {
    const float friction = 0.2f;   // b2MixFriction of two default fixtures: sqrtf(0.2f * 0.2f) == 0.2f in fp32
    const float mH = K->hull_inv_mass, iH = K->hull_inv_I, mW = K->wheel_inv_mass, iW = K->wheel_inv_I;
    for (int k = 0; k < n; ++k) {
        CarContact c = load_contact(recs + k);
        const int ia = c.ia, ib = c.ib;
        const float mA = (ia == 0) ? mH : mW, iA = (ia == 0) ? iH : iW, mB = (ib == 5) ? mH : mW, iB = (ib == 5) ? iH : iW;
        F2 vA = f2(vel[ia][0], vel[ia][1]), vB = f2(vel[ib][0], vel[ib][1]);
        float wA = vel[ia][2], wB = vel[ib][2];
        const F2 normal = f2(c.nx, c.ny), tangent = cross_vs(normal, 1.0f);
        const int vcount = c.vcount;
        contact_tangent<0>(c, tangent, friction, mA, iA, mB, iB, vA, wA, vB, wB);
        if (vcount == 2) contact_tangent<1>(c, tangent, friction, mA, iA, mB, iB, vA, wA, vB, wB);
        const F2 rA0 = f2(c.rAx[0], c.rAy[0]), rB0 = f2(c.rBx[0], c.rBy[0]);
        if (vcount == 1) {
            const F2 dv = ((vB + cross_sv(wB, rB0)) - vA) - cross_sv(wA, rA0);
            const float vn = dot(dv, normal);
            float lambda = -c.nmass[0] * (vn - 0.0f);
            float new_impulse = c.ni[0] + lambda;
            if (!(new_impulse > 0.0f)) new_impulse = 0.0f;
            lambda = new_impulse - c.ni[0];
            c.ni[0] = new_impulse;
            const F2 P = lambda * normal;
            vA = vA - mA * P;
            wA -= iA * cross(rA0, P);
            vB = vB + mB * P;
            wB += iB * cross(rB0, P);
        } else {
            const F2 rA1 = f2(c.rAx[1], c.rAy[1]), rB1 = f2(c.rBx[1], c.rBy[1]);
            const F2 a = f2(c.ni[0], c.ni[1]);
            const F2 dv1 = ((vB + cross_sv(wB, rB0)) - vA) - cross_sv(wA, rA0);
            const F2 dv2 = ((vB + cross_sv(wB, rB1)) - vA) - cross_sv(wA, rA1);
            float vn1 = dot(dv1, normal), vn2 = dot(dv2, normal);
            F2 b = f2(vn1 - 0.0f, vn2 - 0.0f);
            b = b - f2(c.k11 * a.x + c.k12 * a.y, c.k12 * a.x + c.k22 * a.y);
            F2 x;
            bool solved = false;
            for (;;) {
                x = neg(f2(c.nm00 * b.x + c.nm10 * b.y, c.nm01 * b.x + c.nm11 * b.y));
                if (x.x >= 0.0f && x.y >= 0.0f) { solved = true; break; }
                x.x = -c.nmass[0] * b.x; x.y = 0.0f;
                vn1 = 0.0f; vn2 = c.k12 * x.x + b.y;
                if (x.x >= 0.0f && vn2 >= 0.0f) { solved = true; break; }
                x.x = 0.0f; x.y = -c.nmass[1] * b.y;
                vn1 = c.k12 * x.y + b.x; vn2 = 0.0f;
                if (x.y >= 0.0f && vn1 >= 0.0f) { solved = true; break; }
                x.x = 0.0f; x.y = 0.0f;
                vn1 = b.x; vn2 = b.y;
                if (vn1 >= 0.0f && vn2 >= 0.0f) { solved = true; break; }
                break;
            }
            if (solved) {
                const F2 d = x - a;
                const F2 P1 = d.x * normal, P2 = d.y * normal;
                vA = vA - mA * (P1 + P2);
                wA -= iA * (cross(rA0, P1) + cross(rA1, P2));
                vB = vB + mB * (P1 + P2);
                wB += iB * (cross(rB0, P1) + cross(rB1, P2));
                c.ni[0] = x.x; c.ni[1] = x.y;
            }
        }
        vel[ia][0] = vA.x; vel[ia][1] = vA.y; vel[ia][2] = wA;
        vel[ib][0] = vB.x; vel[ib][1] = vB.y; vel[ib][2] = wB;
        *reinterpret_cast<float4*>(recs[k].ni) = make_float4(c.ni[0], c.ni[1], c.ti[0], c.ti[1]);   // ni[2], ti[2] are adjacent
    }
}

// b2ContactSolver::SolvePositionConstraints over all contacts; true when minSeparation >= -3 * linearSlop
__device__ __noinline__ bool car_contacts_solve_position(const CarHullConst* K, const CarContact* recs, int n, float (*pose)[3]) {
    float min_sep = 0.0f;
    const float mH = K->hull_inv_mass, iH = K->hull_inv_I, mW = K->wheel_inv_mass, iW = K->wheel_inv_I;
    const F2 lcH = f2(K->hull_lcx, K->hull_lcy);
    for (int k = 0; k < n; ++k) {
        const CarContact c = load_contact(recs + k);
        const int ia = c.ia, ib = c.ib;
        const float mA = (ia == 0) ? mH : mW, iA = (ia == 0) ? iH : iW, mB = (ib == 5) ? mH : mW, iB = (ib == 5) ? iH : iW;
        const F2 lcA = (ia == 0) ? lcH : f2(0.f, 0.f), lcB = (ib == 5) ? lcH : f2(0.f, 0.f);
        F2 cA = f2(pose[ia][0], pose[ia][1]), cB = f2(pose[ib][0], pose[ib][1]);
        float aA = pose[ia][2], aB = pose[ib][2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            if (j < c.count) {
                const Xf xfa = xf_of(cA, aA, lcA), xfb = xf_of(cB, aB, lcB);
                F2 normal, point;
                float separation;
                if (c.type == 0) {
                    normal = rmul(xfa.q, f2(c.lnx, c.lny));
                    const F2 plane = xf_mul(xfa, f2(c.lpx, c.lpy));
                    const F2 clip = xf_mul(xfb, f2(c.px[j], c.py[j]));
                    separation = dot(clip - plane, normal) - B2_POLYGON_RADIUS - B2_POLYGON_RADIUS;
                    point = clip;
                } else {
                    normal = rmul(xfb.q, f2(c.lnx, c.lny));
                    const F2 plane = xf_mul(xfb, f2(c.lpx, c.lpy));
                    const F2 clip = xf_mul(xfa, f2(c.px[j], c.py[j]));
                    separation = dot(clip - plane, normal) - B2_POLYGON_RADIUS - B2_POLYGON_RADIUS;
                    point = clip;
                    normal = neg(normal);
                }
                const F2 rA = point - cA, rB = point - cB;
                if (separation < min_sep) min_sep = separation;
                const float C = clampf(B2_BAUMGARTE * (separation + B2_LINEAR_SLOP), -B2_MAX_LINEAR_CORRECTION, 0.0f);
                const float rnA = cross(rA, normal), rnB = cross(rB, normal);
                const float Km = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
                const float impulse = Km > 0.0f ? -C / Km : 0.0f;
                const F2 P = impulse * normal;
                cA = cA - mA * P;
                aA -= iA * cross(rA, P);
                cB = cB + mB * P;
                aB += iB * cross(rB, P);
            }
        }
        pose[ia][0] = cA.x; pose[ia][1] = cA.y; pose[ia][2] = aA;
        pose[ib][0] = cB.x; pose[ib][1] = cB.y; pose[ib][2] = aB;
    }
    return min_sep >= -3.0f * B2_LINEAR_SLOP;
}

}  // namespace crl
