// r2d_narrow.cuh — per-pair narrowphase: two-pass SAT, reference-edge selection, incident-edge clipping and
// contact-point construction, as one host/device function evaluated by one thread per candidate pair.
//
// Follows  src/core/collision.zig:221-363 (performNarrowSAT / overlapSAT / clipLineToLine, CollisionPoint.init :38-51),
//          src/core/Bodies/Disc.zig:80-130, src/core/Bodies/Rectangle.zig:114-238, src/core/aabb.zig:12-19.
// The generic axis loop order of the reference is kept (disc: 1 axis, rectangle: 4 axes) because the hysteresis
// rules of overlapSAT make the winning axis order dependent.  Each body's cos/sin is evaluated once and its four
// world vertices once per pair: the reference re-evaluates `rotate2` every time, always with the same inputs, so the
// bits are identical.
#pragma once
#include "r2d_math.cuh"

namespace r2d {

constexpr uint32_t FLAG_STATIC = 1u;       // bit 0 of the per-body flags word
constexpr uint32_t FLAG_RECT = 2u;         // bit 1: shape type (0 disc, 1 rectangle)
constexpr uint32_t FLAG_LARGE = 4u;        // bit 2: wider than the fine broadphase cell (set by the host at upload)
constexpr uint32_t FLAG_WORLD_SHIFT = 8u;  // bits 8..31: world index inside a batch

// What the narrowphase needs to know about one body.
struct BodyView {
    v2 pos;
    float c, s;      // cos/sin of the angle
    float a, b;      // disc: radius, -; rectangle: half width, half height
    uint32_t flags;
    uint32_t id;
    v2 wv[4];        // rectangle: world vertices rot(local[k]) + pos, local = (-w,-h),(-w,h),(w,h),(w,-h) (Rectangle.zig:38-40)
    v2 en[4];        // rectangle: outward edge normals rot90ccw(normalize(wv[k+1] - wv[k])) — what getNormal (Rectangle.zig:132-151)
                     // and clipAgainstEdge (:220-227) both evaluate, always from the same vertices, so once is bit-identical
};

R2D_HD bool is_rect(const BodyView& b) { return (b.flags & FLAG_RECT) != 0; }

// pose = {pos.x, pos.y, cos, sin}; shape = {a, b (full width/height for rects), flags, id}
R2D_HD BodyView make_view(float px, float py, float c, float s, float shape_a, float shape_b, uint32_t flags,
                          uint32_t id) {
    BodyView v;
    v.pos = mk2(px, py);
    v.c = c;
    v.s = s;
    v.flags = flags;
    v.id = id;
    if (flags & FLAG_RECT) {
        const float w = fdiv(shape_a, 2.0f), h = fdiv(shape_b, 2.0f);  // Rectangle.zig:38-39
        v.a = w;
        v.b = h;
        v.wv[0] = add2(rotate_cs(mk2(-w, -h), c, s), v.pos);  // getWorldVertices :88-98 / localToWorld
        v.wv[1] = add2(rotate_cs(mk2(-w, h), c, s), v.pos);
        v.wv[2] = add2(rotate_cs(mk2(w, h), c, s), v.pos);
        v.wv[3] = add2(rotate_cs(mk2(w, -h), c, s), v.pos);
#pragma unroll
        for (int k = 0; k < 4; ++k) v.en[k] = rot90ccw(normalize2(sub2(v.wv[(k + 1) & 3], v.wv[k])));
    } else {
        v.a = shape_a;
        v.b = 0.0f;
        v.wv[0] = v.wv[1] = v.wv[2] = v.wv[3] = v.pos;
        v.en[0] = v.en[1] = v.en[2] = v.en[3] = mk2(0.0f, 0.0f);
    }
    return v;
}

// AABB.intersects (aabb.zig:12-19); aabb = {cx, cy, half_w, half_h}
R2D_HD bool aabb_intersects(float ax, float ay, float ahw, float ahh, float bx, float by, float bhw, float bhh) {
    const float dx = fabs_z(fsub(bx, ax));
    const float dy = fabs_z(fsub(by, ay));
    return (dx <= fadd(fadd(ahw, bhw), AABB_EPS_OVERLAP)) && (dy <= fadd(fadd(ahh, bhh), AABB_EPS_OVERLAP));
}

// closestPoint: Disc.zig:80-84, Rectangle.zig:114-130 (nearest VERTEX, strict <, first wins — Q11)
R2D_HD v2 closest_point(const BodyView& b, v2 p) {
    if (is_rect(b)) {
        float best = inf32();
        v2 best_pos = p;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float d2 = length2sq(sub2(b.wv[k], p));
            if (d2 < best) {
                best = d2;
                best_pos = b.wv[k];
            }
        }
        return best_pos;
    }
    const v2 n = normalize2(sub2(p, b.pos));
    return addmult2(b.pos, n, b.a);
}

struct Edge {
    v2 dir;      // the axis / outward normal
    v2 ea, eb;   // rectangle: the edge itself (world)
    v2 closest;  // disc: the other body's closest point, which the axis points at
};

// getNormal: Disc.zig:86-90 (axis toward the other body's closest point), Rectangle.zig:132-151.
// k is a compile-time-unrollable index wherever this is called, so wv[] / en[] stay in registers.
R2D_HD Edge get_normal(const BodyView& self, const BodyView& other, int k) {
    Edge e;
    if (is_rect(self)) {
        e.dir = self.en[k];
        e.ea = self.wv[k];
        e.eb = self.wv[(k + 1) & 3];
        e.closest = self.pos;
    } else {
        const v2 closest = closest_point(other, self.pos);
        e.dir = normalize2(sub2(closest, self.pos));
        e.ea = e.eb = self.pos;
        e.closest = closest;
    }
    return e;
}

// projectAlongAxis: Disc.zig:92-97, Rectangle.zig:153-169
R2D_HD void project(const BodyView& b, v2 n, float& lo, float& hi) {
    if (is_rect(b)) {
        lo = inf32();
        hi = -inf32();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float d = dot2(b.wv[k], n);
            if (d < lo) lo = d;
            if (d > hi) hi = d;
        }
    } else {
        const float mid = dot2(b.pos, n);
        lo = fsub(mid, b.a);
        hi = fadd(mid, b.a);
    }
}

struct SatResult {
    bool collides;
    v2 normal;
    float penetration;
    int normal_id;
    int ref_is_first;  // 1: key = (first arg of the pass that took it ... ) see overlap_sat
    // the unflipped axis of the pass that took it and the closest point it was aimed at: identifyCollisionPoints of a disc
    // reference evaluates getNormal / closestPoint again with the same arguments (Disc.zig:99-123) — same bits, kept here
    v2 axis_dir, axis_closest;
};

// normalShouldFlipSAT (collision.zig:221-226)
R2D_HD bool normal_should_flip(v2 n, const BodyView& ref, const BodyView& inc) {
    return dot2(n, sub2(inc.pos, ref.pos)) < 0.0f;
}

// overlapSAT (collision.zig:228-290).  `ref_tag` is stored into ret.ref_is_first whenever this pass takes the axis.
// The axis loop is NOT unrolled (code size: the narrowphase is instruction-fetch bound otherwise).  To keep the
// vertex / normal arrays in registers without dynamic indexing, R is a private copy whose arrays are rotated by one
// position per axis: axis k is always en[0] with the edge (wv[0], wv[1]).  project() takes min / max over all four
// vertices, which does not depend on their order, so every value is the one the indexed form computes.
R2D_HD bool overlap_sat(SatResult& ret, BodyView R, const BodyView& I, int ref_tag) {
    const float EPS = SAT_OVERLAP_THRESHOLD;
    const int nn = is_rect(R) ? 4 : 1;
#pragma unroll 1
    for (int k = 0; k < nn; ++k) {
        const Edge axis = get_normal(R, I, 0);
        v2 normal = axis.dir;
        bool flipped = false;
        if (normal_should_flip(normal, R, I)) {
            normal = negate_mul(normal);
            flipped = true;
        }
        float p1lo, p1hi, p2lo, p2hi;
        project(R, normal, p1lo, p1hi);
        project(I, normal, p2lo, p2hi);
        const float d1 = fsub(p1hi, p2lo);
        const float d2 = fsub(p2hi, p1lo);
        const float d = fmin_z(d1, d2);
        if (d <= -COLLISION_MARGIN) return false;

        if (!flipped && approx_eql2(normal, ret.normal, EPS)) {  // :254-268
            const float tref = dot2(R.pos, normal);
            const float tinc = dot2(I.pos, normal);
            if (tref < fsub(tinc, EPS)) {
                ret.penetration = d;
                ret.normal = normal;
                ret.normal_id = k;
                ret.ref_is_first = ref_tag;
                ret.axis_dir = axis.dir;
                ret.axis_closest = axis.closest;
            }
        }
        if (!flipped && approx_eql2(normal, negate2(ret.normal), EPS)) {  // :270-279
            const v2 diff = sub2(I.pos, R.pos);
            if (dot2(diff, normal) > fadd(dot2(diff, ret.normal), EPS)) {
                ret.penetration = d;
                ret.normal = normal;
                ret.normal_id = k;
                ret.ref_is_first = ref_tag;
                ret.axis_dir = axis.dir;
                ret.axis_closest = axis.closest;
            }
        }
        if (fadd(d, EPS) < ret.penetration) {  // :281-286
            ret.penetration = d;
            ret.normal = normal;
            ret.normal_id = k;
            ret.ref_is_first = ref_tag;
            ret.axis_dir = axis.dir;
            ret.axis_closest = axis.closest;
        }
        // next axis: rotate the private copy
        const v2 w0 = R.wv[0], e0 = R.en[0];
        R.wv[0] = R.wv[1]; R.wv[1] = R.wv[2]; R.wv[2] = R.wv[3]; R.wv[3] = w0;
        R.en[0] = R.en[1]; R.en[1] = R.en[2]; R.en[2] = R.en[3]; R.en[3] = e0;
    }
    return true;
}

// a ? x : y, field by field (register selects instead of two copies of the code that follows)
R2D_HD BodyView select_view(bool a, const BodyView& x, const BodyView& y) {
    BodyView v;
    v.pos = a ? x.pos : y.pos;
    v.c = a ? x.c : y.c;
    v.s = a ? x.s : y.s;
    v.a = a ? x.a : y.a;
    v.b = a ? x.b : y.b;
    v.flags = a ? x.flags : y.flags;
    v.id = a ? x.id : y.id;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        v.wv[k] = a ? x.wv[k] : y.wv[k];
        v.en[k] = a ? x.en[k] : y.en[k];
    }
    return v;
}

struct ContactPoint {
    v2 pos;      // stored mid-depth point
    float depth;
    v2 ref_r, inc_r;
};

struct Manifold {
    int collides;
    int ref_is_lo;   // 1 if the reference body is the lower-id body (o1)
    int normal_id;
    int n_points;
    v2 normal;
    ContactPoint pt[2];
};

// CollisionPoint.init (collision.zig:38-51)
R2D_HD ContactPoint make_point(v2 pos, float depth, const BodyView& ref, const BodyView& inc, v2 normal) {
    ContactPoint cp;
    const v2 middle = sub2(pos, scale2(normal, fdiv(depth, 2.0f)));
    cp.ref_r = sub2(middle, ref.pos);
    cp.inc_r = sub2(middle, inc.pos);
    cp.pos = middle;
    cp.depth = depth;
    return cp;
}

// clipLineToLine (collision.zig:330-363): line a is the reference edge, b the incident edge
R2D_HD void clip_line_to_line(v2 aa, v2 ab, v2 ba, v2 bb, v2& p1, v2& p2) {
    const v2 delta_a = sub2(ab, aa);
    const float a_len = length2(delta_a);
    const v2 tang = scale2(delta_a, fdiv(1.0f, a_len));
    p1 = ba;
    p2 = bb;
    {
        const float scalar = dot2(sub2(ba, aa), tang);
        if (!(scalar > 0.0f && scalar < a_len)) {
            const v2 clos = (scalar < fmul(0.5f, a_len)) ? aa : ab;
            const v2 delta_p = sub2(ba, bb);
            const float t = fdiv(-dot2(tang, sub2(bb, clos)), dot2(tang, delta_p));
            p1 = addmult2(bb, delta_p, t);
        }
    }
    {
        const float scalar = dot2(sub2(bb, aa), tang);
        if (!(scalar > 0.0f && scalar < a_len)) {
            const v2 clos = (scalar < fmul(0.5f, a_len)) ? aa : ab;
            const v2 delta_p = sub2(bb, ba);
            const float t = fdiv(-dot2(tang, sub2(ba, clos)), dot2(tang, delta_p));
            p2 = addmult2(ba, delta_p, t);
        }
    }
}

// identifyCollisionPoints: Disc.zig:99-123, Rectangle.zig:171-211 (+ clipAgainstEdge Disc.zig:125-130, Rectangle.zig:213-238)
R2D_HD void identify_points(Manifold& m, const BodyView& ref, const BodyView& inc, int normal_id, v2 axis_dir, v2 axis_closest) {
    m.n_points = 0;
    if (!is_rect(ref)) {
        // closestPoint(inc, ref.pos) and getNormal(ref, inc): exactly what the SAT pass that chose this reference evaluated
        const v2 pos = axis_closest;
        v2 normal = axis_dir;
        if (normal_should_flip(normal, ref, inc)) normal = negate_mul(normal);
        const float dot = dot2(normal, sub2(pos, ref.pos));
        const float depth = fsub(dot, ref.a);
        m.pt[0] = make_point(pos, depth, ref, inc, normal);
        m.n_points = 1;
        return;
    }
    Edge n;  // UNFLIPPED outward normal + edge of the reference face (Q13); selected without dynamic indexing
    n = get_normal(ref, inc, 0);
#pragma unroll
    for (int k = 1; k < 4; ++k)
        if (k == normal_id) n = get_normal(ref, inc, k);
    if (is_rect(inc)) {
        // Rectangle.clipAgainstEdge: incident edge = argmin_k dot(normal, outward_normal_k), strict <
        v2 best_a = inc.wv[0], best_b = inc.wv[0];
        float best_dot = inf32();
        v2 curr = inc.wv[0];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int ni = (i == 3) ? 0 : i + 1;
            const v2 next = inc.wv[ni];
            const v2 tentative = inc.en[i];  // rot90ccw(normalize(next - curr)), Rectangle.zig:224-225
            const float d = dot2(n.dir, tentative);
            if (d < best_dot) {
                best_a = curr;
                best_b = next;
                best_dot = d;
            }
            curr = next;
        }
        v2 ca, cb;
        clip_line_to_line(n.ea, n.eb, best_a, best_b, ca, cb);
        int i = 0;
        float dot = dot2(sub2(ca, n.ea), n.dir);
        if (dot < COLLISION_MARGIN) {
            m.pt[i] = make_point(ca, dot, ref, inc, n.dir);
            i += 1;
        }
        dot = dot2(sub2(cb, n.eb), n.dir);
        if (dot < COLLISION_MARGIN) {
            m.pt[i] = make_point(cb, dot, ref, inc, n.dir);
            i += 1;
        }
        m.n_points = i;
    } else {
        // Disc.clipAgainstEdge: point = pos + normal * (-radius)
        const v2 pos = addmult2(inc.pos, n.dir, -inc.a);
        const float dot = dot2(sub2(pos, n.ea), n.dir);
        m.pt[0] = make_point(pos, dot, ref, inc, n.dir);
        m.n_points = 1;
    }
}

// performNarrowSAT (collision.zig:299-320) + manifold construction (lib.zig:287-294).
// `lo` must be the lower-id body, `hi` the higher-id one (:304-307).
R2D_HD Manifold narrowphase(const BodyView& lo, const BodyView& hi) {
    Manifold m;
    m.collides = 0;
    m.n_points = 0;
    m.normal_id = 0;
    m.ref_is_lo = 1;
    SatResult ret;
    ret.collides = false;
    ret.penetration = inf32();
    ret.normal = mk2(u2f(0xAAAAAAAAu), u2f(0xAAAAAAAAu));  // `undefined` in the reference; never decides anything (Q12)
    ret.normal_id = 0;
    ret.ref_is_first = 1;
    ret.axis_dir = ret.axis_closest = mk2(0.0f, 0.0f);
    m.normal = ret.normal;
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {  // (lo, hi) then (hi, lo): unrolled, so that select_view folds away
        const bool first = pass == 0;
        if (!overlap_sat(ret, select_view(first, lo, hi), select_view(first, hi, lo), first ? 1 : 0)) return m;
    }
    m.collides = 1;
    m.ref_is_lo = ret.ref_is_first;
    m.normal_id = ret.normal_id;
    m.normal = ret.normal;
    const bool ref_lo = m.ref_is_lo != 0;
    identify_points(m, select_view(ref_lo, lo, hi), select_view(ref_lo, hi, lo), ret.normal_id, ret.axis_dir, ret.axis_closest);
    return m;
}

}  // namespace r2d
