// r2d_oracle.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A single-threaded, array-of-structs, line-by-line C++ restatement of resolve2d's `Solver.process` hot path.  It is
// the parity checker for the CUDA path and the CPU baseline of bench.py.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load it; nothing under resolve2d_b200/ does.
//
// Pinning: this file is checked bit for bit against outputs of the reference itself — its prebuilt binary
// demos/web/public/resolve2d.wasm executed by oracle/wasm_interp.cpp (tests/golden/make_wasm_golden.py ->
// tests/golden/wasm_golden.json, tests/test_oracle_wasm_golden.py: every step of six runs incl. body removal and
// other dt/sub_steps/iters) and against the survey's vectors (tests/golden/appendix_f.json,
// tests/test_oracle_golden.py: also candidate-pair sets and manifold lists).
//
// Every block cites the Zig source it restates (paths relative to /root/reference/src/core).  Build with
// -ffp-contract=off: all arithmetic is IEEE f32, one rounding per operation, no FMA (wasm semantics).
// Trig is the musl-derived sinf/cosf that Zig 0.14.1's compiler-rt supplies, never libm's.
//
// Deviations from the reference (declared): ids are u32 (reference u16, Bodies/RigidBody.zig:15); manifold keys use
// body indices instead of body pointers (collision.zig:12-15); joints whose body id is missing make process() fail
// before any state is touched.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <set>
#include <unordered_map>
#include <unordered_set>
#include <vector>

namespace {

// ---- simulation_constants.zig:3-22 ------------------------------------------------------------------------------
const float MIN_MANIFOLD_IMPULSE = 1e-4f;
const float BAUMGARTE = 0.02f;
const float BAUMGARTE_SLOP = 0.005f;
const float SAT_OVERLAP_THRESHOLD = 1e-4f;
const float COLLISION_MARGIN = 0.01f;
const float NMATH_WARN_DIVIDING_BELOW = 1e-3f;
const float AABB_EPS_OVERLAP = 0.01f;
const float CONSTRAINT_GRADIENT_DIVISION_LIMIT = 1e-4f;
const float ALLOWED_CONSTRAINT_VALUE = 1e-6f;

// ---- Zig builtins ------------------------------------------------------------------------------------------------
// @max/@min on f32 lower to compiler-rt fmaxf/fminf: NaN-ignoring, `if (x < y) y else x`.
static inline float zmax(float x, float y) {
    if (x != x) return y;
    if (y != y) return x;
    return (x < y) ? y : x;
}
static inline float zmin(float x, float y) {
    if (x != x) return y;
    if (y != y) return x;
    return (x < y) ? x : y;
}
static inline float zclamp(float v, float lo, float hi) { return zmax(lo, zmin(v, hi)); }  // std.math.clamp
static inline float zabs(float x) { return fabsf(x); }

// ---- compiler-rt trig (lib/compiler_rt/trig.zig, sin.zig, cos.zig, rem_pio2f.zig == musl) ---------------------------
static float cosdf(double x) {
    static const double C0 = -0x1.ffffffd0c5e81p-2, C1 = 0x1.55553e1053a42p-5, C2 = -0x1.6c087e80f1e27p-10,
                        C3 = 0x1.99342e0ee5069p-16;
    double z = x * x;
    double w = z * z;
    double r = C2 + z * C3;
    return (float)(((1.0 + z * C0) + w * C1) + (w * z) * r);
}
static float sindf(double x) {
    static const double S1 = -0x1.5555554cbac77p-3, S2 = 0x1.11110896efbb2p-7, S3 = -0x1.a00f9e2cae774p-13,
                        S4 = 0x1.6cd878c3b46a7p-19;
    double z = x * x;
    double w = z * z;
    double r = S3 + z * S4;
    double s = z * x;
    return (float)((x + s * (S1 + z * S2)) + s * w * r);
}
static int rem_pio2f_medium(float x, double* y) {
    static const double toint = 1.5 / 2.220446049250313e-16;  // 1.5/DBL_EPSILON
    static const double pio4 = 0x1.921fb6p-1;
    static const double invpio2 = 6.36619772367581382433e-01;
    static const double pio2_1 = 1.57079631090164184570e+00;
    static const double pio2_1t = 1.58932547735281966916e-08;
    double fn = (double)x * invpio2 + toint - toint;
    int n = (int32_t)fn;
    *y = x - fn * pio2_1 - fn * pio2_1t;
    if (*y < -pio4) {
        n--;
        fn--;
        *y = x - fn * pio2_1 - fn * pio2_1t;
    } else if (*y > pio4) {
        n++;
        fn++;
        *y = x - fn * pio2_1 - fn * pio2_1t;
    }
    return n;
}
static const double PIO2_1 = 1 * M_PI_2, PIO2_2 = 2 * M_PI_2, PIO2_3 = 3 * M_PI_2, PIO2_4 = 4 * M_PI_2;

static float zsinf(float x) {
    uint32_t ix;
    memcpy(&ix, &x, 4);
    int sign = ix >> 31;
    ix &= 0x7fffffff;
    if (ix <= 0x3f490fda) {
        if (ix < 0x39800000) return x;
        return sindf(x);
    }
    if (ix <= 0x407b53d1) {
        if (ix <= 0x4016cbe3) {
            if (sign) return -cosdf(x + PIO2_1);
            return cosdf(x - PIO2_1);
        }
        return sindf(sign ? -(x + PIO2_2) : -(x - PIO2_2));
    }
    if (ix <= 0x40e231d5) {
        if (ix <= 0x40afeddf) {
            if (sign) return cosdf(x + PIO2_3);
            return -cosdf(x - PIO2_3);
        }
        return sindf(sign ? x + PIO2_4 : x - PIO2_4);
    }
    if (ix >= 0x7f800000) return x - x;
    if (ix >= 0x4dc90fdb) return NAN;  // rem_pio2_large: |angle| > 4.2e8 rad, unreachable in any scene (documented)
    double y;
    int n = rem_pio2f_medium(x, &y);
    switch (n & 3) {
        case 0: return sindf(y);
        case 1: return cosdf(y);
        case 2: return sindf(-y);
        default: return -cosdf(y);
    }
}
static float zcosf(float x) {
    uint32_t ix;
    memcpy(&ix, &x, 4);
    int sign = ix >> 31;
    ix &= 0x7fffffff;
    if (ix <= 0x3f490fda) {
        if (ix < 0x39800000) return 1.0f;
        return cosdf(x);
    }
    if (ix <= 0x407b53d1) {
        if (ix > 0x4016cbe3) return -cosdf(sign ? x + PIO2_2 : x - PIO2_2);
        if (sign) return sindf(x + PIO2_1);
        return sindf(PIO2_1 - x);
    }
    if (ix <= 0x40e231d5) {
        if (ix > 0x40afeddf) return cosdf(sign ? x + PIO2_4 : x - PIO2_4);
        if (sign) return sindf(-x - PIO2_3);
        return sindf(x - PIO2_3);
    }
    if (ix >= 0x7f800000) return x - x;
    if (ix >= 0x4dc90fdb) return NAN;
    double y;
    int n = rem_pio2f_medium(x, &y);
    switch (n & 3) {
        case 0: return cosdf(y);
        case 1: return sindf(-y);
        case 2: return -cosdf(y);
        default: return sindf(y);
    }
}

// ---- nmath.zig ---------------------------------------------------------------------------------------------------
static inline bool approxEql(float a, float b, float eps) { return a > b - eps && a < b + eps; }  // :5-7

struct Vector2 {
    float x = 0.0f, y = 0.0f;
    Vector2() {}
    Vector2(float x_, float y_) : x(x_), y(y_) {}
    void add(Vector2 v) { x += v.x; y += v.y; }                         // :24-27
    void sub(Vector2 v) { x -= v.x; y -= v.y; }                         // :29-32
    void addmult(Vector2 v, float s) { x += v.x * s; y += v.y * s; }    // :39-42
    void negate() { x *= -1; y *= -1; }                                 // :48-51
};
static inline Vector2 add2(Vector2 a, Vector2 b) { return Vector2(a.x + b.x, a.y + b.y); }
static inline Vector2 sub2(Vector2 a, Vector2 b) { return Vector2(a.x - b.x, a.y - b.y); }
static inline Vector2 scale2(Vector2 a, float s) { return Vector2(a.x * s, a.y * s); }
static inline float dot2(Vector2 a, Vector2 b) { return a.x * b.x + a.y * b.y; }
static inline float cross2(Vector2 a, Vector2 b) { return a.x * b.y - a.y * b.x; }
static inline float length2sq(Vector2 a) { return dot2(a, a); }
static inline float length2(Vector2 a) { return sqrtf(length2sq(a)); }
static inline Vector2 normalize2(Vector2 a) {  // :95-103
    const float len = length2(a);
    if (len < NMATH_WARN_DIVIDING_BELOW) return Vector2();
    return scale2(a, 1 / len);
}
static inline Vector2 negate2(Vector2 a) { return Vector2(-a.x, -a.y); }
static inline Vector2 addmult2(Vector2 a, Vector2 b, float s) { return add2(a, scale2(b, s)); }
static inline bool approxEql2(Vector2 a, Vector2 b, float eps) { return approxEql(a.x, b.x, eps) && approxEql(a.y, b.y, eps); }
static inline Vector2 rotate2(Vector2 a, float angle) {  // :129-136
    const float c = zcosf(angle);
    const float s = zsinf(angle);
    return Vector2(a.x * c - a.y * s, a.x * s + a.y * c);
}
static inline float dist2(Vector2 a, Vector2 b) { return length2(sub2(a, b)); }
static inline Vector2 rotate90clockwise(Vector2 a) { return Vector2(a.y, -a.x); }
static inline Vector2 rotate90counterclockwise(Vector2 a) { return Vector2(-a.y, a.x); }

// ---- aabb.zig ----------------------------------------------------------------------------------------------------
struct AABB {
    Vector2 pos;
    float half_width = 0, half_height = 0;
    bool intersects(const AABB& other) const {  // :12-19
        const float EPS = AABB_EPS_OVERLAP;
        const float dx = zabs(other.pos.x - pos.x);
        const float dy = zabs(other.pos.y - pos.y);
        return (dx <= (half_width + other.half_width) + EPS) && (dy <= (half_height + other.half_height) + EPS);
    }
    void getVertices(Vector2 out[4]) const {  // :25-32
        out[0] = Vector2(pos.x - half_width, pos.y - half_height);
        out[1] = Vector2(pos.x + half_width, pos.y - half_height);
        out[2] = Vector2(pos.x + half_width, pos.y + half_height);
        out[3] = Vector2(pos.x - half_width, pos.y + half_height);
    }
};

// ---- Bodies/RigidBody.zig ------------------------------------------------------------------------------------------
enum BodyType { DISC = 0, RECTANGLE = 1 };
struct Props {  // :69-81
    Vector2 pos, momentum, force;
    float mass = 0;
    float angle = 0, ang_momentum = 0, torque = 0, inertia = 0, mu = 0;
};
struct Line {
    Vector2 a, b;
};
struct Edge {  // :22-26
    Vector2 dir;
    bool has_edge = false;
    Line edge;
};
struct Incident {  // :17-20
    bool is_edge = false;
    Line edge;
    Vector2 point;
};
struct RigidBody {  // :83-91 ; the vtable is a switch on `type`
    uint32_t id = 0;
    AABB aabb;
    bool is_static = false;
    size_t num_normals = 0;
    BodyType type = DISC;
    Props props;
    // Disc.zig:16 / Rectangle.zig:17-20 ("ptr" payload)
    float radius = 0;
    float width = 0, height = 0;
    Vector2 local_vertices[4];
};

struct CollisionPoint {  // collision.zig:22-51
    float accumulated_pn = 0, accumulated_pt = 0;
    Vector2 ref_r, inc_r, pos;
    float depth = 0, original_depth = 0;
    float mass_n = 0, mass_t = 0;
};
struct OptPoint {
    bool present = false;
    CollisionPoint p;
};

static CollisionPoint CollisionPoint_init(Vector2 pos, float depth, const RigidBody& ref, const RigidBody& inc, Vector2 normal) {
    const Vector2 middle = sub2(pos, scale2(normal, depth / 2));
    CollisionPoint cp;
    cp.accumulated_pn = 0;
    cp.accumulated_pt = 0;
    cp.ref_r = sub2(middle, ref.props.pos);
    cp.inc_r = sub2(middle, inc.props.pos);
    cp.pos = middle;
    cp.depth = depth;
    cp.original_depth = depth;
    cp.mass_n = 0;
    cp.mass_t = 0;
    return cp;
}

// ---- shape "vtable" -------------------------------------------------------------------------------------------------
static Vector2 localToWorld(const RigidBody& b, Vector2 pos) {  // RigidBody.zig:113-117
    const Vector2 r = rotate2(pos, b.props.angle);
    return add2(r, b.props.pos);
}
static void getWorldVertices(const RigidBody& b, Vector2 ret[4]) {  // Rectangle.zig:88-98
    for (int idx = 0; idx < 4; ++idx) {
        const Vector2 r = rotate2(b.local_vertices[idx], b.props.angle);
        ret[idx] = add2(r, b.props.pos);
    }
}
static void updateAABB(RigidBody& b) {
    if (b.type == DISC) {  // Disc.zig:63-66
        b.aabb.pos = b.props.pos;
        b.aabb.half_width = b.radius;
        b.aabb.half_height = b.radius;
    } else {  // Rectangle.zig:71-86
        float width = 0, height = 0;
        for (int k = 0; k < 4; ++k) {
            const Vector2 rot = rotate2(b.local_vertices[k], b.props.angle);
            if (rot.x > width) width = rot.x;
            if (rot.y > height) height = rot.y;
        }
        b.aabb.pos = b.props.pos;
        b.aabb.half_width = width;
        b.aabb.half_height = height;
    }
}
static Vector2 closestPoint(const RigidBody& b, Vector2 pos) {
    if (b.type == DISC) {  // Disc.zig:80-84
        const Vector2 normal = normalize2(sub2(pos, b.props.pos));
        return addmult2(b.props.pos, normal, b.radius);
    }
    // Rectangle.zig:114-130
    float best_dist2 = INFINITY;
    Vector2 best_pos = pos;
    Vector2 world_vertices[4];
    getWorldVertices(b, world_vertices);
    for (int k = 0; k < 4; ++k) {
        const float d2 = length2sq(sub2(world_vertices[k], pos));
        if (d2 < best_dist2) {
            best_dist2 = d2;
            best_pos = world_vertices[k];
        }
    }
    return best_pos;
}
static Edge getNormal(const RigidBody& self, const RigidBody& body, size_t iter) {
    Edge e;
    if (self.type == DISC) {  // Disc.zig:86-90
        const Vector2 closest = closestPoint(body, self.props.pos);
        e.dir = normalize2(sub2(closest, self.props.pos));
        return e;
    }
    // Rectangle.zig:132-151
    const size_t next_iter = (iter == 3) ? 0 : iter + 1;
    const Vector2 vert = self.local_vertices[iter];
    const Vector2 next_vert = self.local_vertices[next_iter];
    const Vector2 r1 = rotate2(vert, self.props.angle);
    const Vector2 a1 = add2(r1, self.props.pos);
    const Vector2 r2 = rotate2(next_vert, self.props.angle);
    const Vector2 a2 = add2(r2, self.props.pos);
    const Vector2 dir = normalize2(sub2(a2, a1));
    e.dir = rotate90counterclockwise(dir);
    e.has_edge = true;
    e.edge.a = a1;
    e.edge.b = a2;
    return e;
}
static void projectAlongAxis(const RigidBody& b, Vector2 normal, float out[2]) {
    if (b.type == DISC) {  // Disc.zig:92-97
        const float middle = dot2(b.props.pos, normal);
        const float rad = b.radius;
        out[0] = middle - rad;
        out[1] = middle + rad;
        return;
    }
    // Rectangle.zig:153-169
    float best_low = INFINITY, best_high = -INFINITY;
    for (int k = 0; k < 4; ++k) {
        const Vector2 r = rotate2(b.local_vertices[k], b.props.angle);
        const Vector2 world = add2(r, b.props.pos);
        const float dot = dot2(world, normal);
        if (dot < best_low) best_low = dot;
        if (dot > best_high) best_high = dot;
    }
    out[0] = best_low;
    out[1] = best_high;
}

static bool normalShouldFlipSAT(Vector2 normal, const RigidBody& reference, const RigidBody& incident) {  // collision.zig:221-226
    return dot2(normal, sub2(incident.props.pos, reference.props.pos)) < 0.0f;
}

static Line clipLineToLine(Line a, Line b) {  // collision.zig:330-363
    const Vector2 delta_a = sub2(a.b, a.a);
    const float a_len = length2(delta_a);
    const Vector2 tang = scale2(delta_a, 1 / a_len);
    Vector2 p1 = b.a, p2 = b.b;
    {
        const float scalar = dot2(sub2(b.a, a.a), tang);
        if (!(scalar > 0 && scalar < a_len)) {
            const Vector2 clos = (scalar < 0.5f * a_len) ? a.a : a.b;
            const Vector2 delta_p = sub2(b.a, b.b);
            const float t = -dot2(tang, sub2(b.b, clos)) / dot2(tang, delta_p);
            p1 = addmult2(b.b, delta_p, t);
        }
    }
    {
        const float scalar = dot2(sub2(b.b, a.a), tang);
        if (!(scalar > 0 && scalar < a_len)) {
            const Vector2 clos = (scalar < 0.5f * a_len) ? a.a : a.b;
            const Vector2 delta_p = sub2(b.b, b.a);
            const float t = -dot2(tang, sub2(b.a, clos)) / dot2(tang, delta_p);
            p2 = addmult2(b.a, delta_p, t);
        }
    }
    Line l;
    l.a = p1;
    l.b = p2;
    return l;
}

static Incident clipAgainstEdge(const RigidBody& self, const Edge& edge) {  // RigidBody.zig:123-125
    Incident inc;
    const Vector2 normal = edge.dir;
    if (self.type == DISC) {  // Disc.zig:125-130
        inc.is_edge = false;
        inc.point = addmult2(self.props.pos, normal, -self.radius);
        return inc;
    }
    // Rectangle.zig:213-238
    Line best_edge;
    float best_dot = INFINITY;
    Vector2 curr_world = localToWorld(self, self.local_vertices[0]);
    for (int i = 0; i < 4; ++i) {
        const int next_idx = (i == 3) ? 0 : i + 1;
        const Vector2 next_world = localToWorld(self, self.local_vertices[next_idx]);
        const Vector2 tangent = normalize2(sub2(next_world, curr_world));
        const Vector2 tentative_normal = rotate90counterclockwise(tangent);
        const float dot = dot2(normal, tentative_normal);
        if (dot < best_dot) {
            best_edge.a = curr_world;
            best_edge.b = next_world;
            best_dot = dot;
        }
        curr_world = next_world;
    }
    Line ref_line;
    ref_line.a = edge.edge.a;
    ref_line.b = edge.edge.b;
    const Line clipped = clipLineToLine(ref_line, best_edge);
    inc.is_edge = true;
    inc.edge = clipped;
    return inc;
}

static void identifyCollisionPoints(const RigidBody& self, const RigidBody& incident, size_t active_normal_iter, OptPoint ret[2]) {
    ret[0].present = false;
    ret[1].present = false;
    if (self.type == DISC) {  // Disc.zig:99-123
        const Vector2 pos = closestPoint(incident, self.props.pos);
        const Edge edge = getNormal(self, incident, active_normal_iter);
        Vector2 normal = edge.dir;
        if (normalShouldFlipSAT(normal, self, incident)) normal.negate();
        const float dot = dot2(normal, sub2(pos, self.props.pos));
        const float depth = dot - self.radius;
        ret[0].present = true;
        ret[0].p = CollisionPoint_init(pos, depth, self, incident, normal);
        return;
    }
    // Rectangle.zig:171-211
    const Edge n = getNormal(self, incident, active_normal_iter);
    const Incident incident_edge = clipAgainstEdge(incident, n);
    if (incident_edge.is_edge) {
        const Vector2 a = incident_edge.edge.a;
        const Vector2 b = incident_edge.edge.b;
        size_t i = 0;
        float dot = dot2(sub2(a, n.edge.a), n.dir);
        if (dot < COLLISION_MARGIN) {
            ret[i].present = true;
            ret[i].p = CollisionPoint_init(a, dot, self, incident, n.dir);
            i += 1;
        }
        dot = dot2(sub2(b, n.edge.b), n.dir);
        if (dot < COLLISION_MARGIN) {
            ret[i].present = true;
            ret[i].p = CollisionPoint_init(b, dot, self, incident, n.dir);
        }
    } else {
        const Vector2 pos = incident_edge.point;
        const float dot = dot2(sub2(pos, n.edge.a), n.dir);
        ret[0].present = true;
        ret[0].p = CollisionPoint_init(pos, dot, self, incident, n.dir);
    }
}

// ---- collision.zig: SAT ---------------------------------------------------------------------------------------------
struct SATResult {  // :292-298 ; key = body indices instead of pointers
    bool collides;
    Vector2 normal;
    float penetration;
    size_t reference_normal_id;
    size_t key_ref, key_inc;
};

static bool overlapSAT(SATResult* ret, const std::vector<RigidBody>& B, size_t reference_i, size_t incident_i) {  // :228-290
    const float EPS = SAT_OVERLAP_THRESHOLD;
    const RigidBody& reference = B[reference_i];
    const RigidBody& incident = B[incident_i];
    for (size_t iter_performed = 0; iter_performed < reference.num_normals;) {
        const Edge edge = getNormal(reference, incident, iter_performed);
        iter_performed += 1;
        Vector2 normal = edge.dir;
        bool flipped = false;
        if (normalShouldFlipSAT(normal, reference, incident)) {
            normal.negate();
            flipped = true;
        }
        float p1[2], p2[2];
        projectAlongAxis(reference, normal, p1);
        projectAlongAxis(incident, normal, p2);
        const float d1 = p1[1] - p2[0];
        const float d2 = p2[1] - p1[0];
        const float d = zmin(d1, d2);
        if (d <= -COLLISION_MARGIN) return false;

        if (!flipped && approxEql2(normal, ret->normal, EPS)) {
            const float tref = dot2(reference.props.pos, normal);
            const float tinc = dot2(incident.props.pos, normal);
            if (tref < tinc - EPS) {
                ret->penetration = d;
                ret->normal = normal;
                ret->reference_normal_id = iter_performed - 1;
                ret->key_ref = reference_i;
                ret->key_inc = incident_i;
            }
        }
        if (!flipped && approxEql2(normal, negate2(ret->normal), EPS)) {
            const Vector2 diff = sub2(incident.props.pos, reference.props.pos);
            if (dot2(diff, normal) > dot2(diff, ret->normal) + EPS) {
                ret->penetration = d;
                ret->normal = normal;
                ret->reference_normal_id = iter_performed - 1;
                ret->key_ref = reference_i;
                ret->key_inc = incident_i;
            }
        }
        if (d + EPS < ret->penetration) {
            ret->penetration = d;
            ret->normal = normal;
            ret->reference_normal_id = iter_performed - 1;
            ret->key_ref = reference_i;
            ret->key_inc = incident_i;
        }
    }
    return true;
}

static SATResult performNarrowSAT(const std::vector<RigidBody>& B, size_t b1, size_t b2) {  // :299-320
    SATResult ret;
    ret.collides = false;
    ret.penetration = INFINITY;
    // `undefined`: any finite value gives the same outcome; 0xAAAAAAAA is Zig's debug fill pattern.
    uint32_t und = 0xAAAAAAAAu;
    memcpy(&ret.normal.x, &und, 4);
    memcpy(&ret.normal.y, &und, 4);
    ret.reference_normal_id = 0;
    ret.key_ref = b1;
    ret.key_inc = b2;
    const uint32_t num1 = B[b1].id, num2 = B[b2].id;
    const size_t o1 = (num1 < num2) ? b1 : b2;
    const size_t o2 = (o1 == b1) ? b2 : b1;
    if (!overlapSAT(&ret, B, o1, o2)) return ret;
    if (!overlapSAT(&ret, B, o2, o1)) return ret;
    ret.collides = true;
    return ret;
}

// ---- collision.zig: CollisionManifold --------------------------------------------------------------------------------
struct CollisionManifold {  // :54-69
    size_t ref_body, inc_body;  // the CollisionKey (indices into `bodies`)
    size_t reference_normal_id = 0;
    Vector2 normal, tangent;
    OptPoint points[2];
    float prev_angle_1 = 0, prev_angle_2 = 0;
    Vector2 applied_linear_p1, applied_linear_p2;
    float applied_rot_p_1 = 0, applied_rot_p_2 = 0;
    float friction = 0;
    uint32_t color = 0;  // oracle-side colouring for the `permuted` Gauss-Seidel order (not in the reference)
    // R2D_OPT_WARM_START (README.md:60-61, not in the reference): the per-substep average of the impulses this contact had
    // accumulated at the end of the previous call, added to the FIRST update of this call as the initial guess
    bool warm_first = false;
    float warm_pn[2] = {0, 0}, warm_pt[2] = {0, 0};

    void updateTGSDepth(std::vector<RigidBody>& B) {  // :73-100
        RigidBody& b1 = B[ref_body];
        RigidBody& b2 = B[inc_body];
        for (int k = 0; k < 2; ++k) {
            if (!points[k].present) continue;
            CollisionPoint& point = points[k].p;  // by reference (Q26)
            if (approxEql(point.depth, point.original_depth, 1e-4f)) continue;
            const Vector2 r1 = point.ref_r, r2 = point.inc_r;
            const Vector2 r_rot_1 = rotate2(r1, b1.props.angle - prev_angle_1);
            const Vector2 r_rot_2 = rotate2(r2, b2.props.angle - prev_angle_2);
            const Vector2 a1 = add2(r_rot_1, b1.props.pos);
            const Vector2 a2 = add2(r_rot_2, b2.props.pos);
            const float depth = dot2(normal, sub2(a2, a1));
            point.depth = depth + point.original_depth;
            point.ref_r = r_rot_1;
            point.inc_r = r_rot_2;
        }
        prev_angle_1 = b1.props.angle;
        prev_angle_2 = b2.props.angle;
    }

    void preStep(std::vector<RigidBody>& B) {  // :102-133
        RigidBody& b1 = B[ref_body];
        RigidBody& b2 = B[inc_body];
        const float inv_m1 = b1.is_static ? 0 : (1 / b1.props.mass);
        const float inv_m2 = b2.is_static ? 0 : (1 / b2.props.mass);
        const float inv_mass = inv_m1 + inv_m2;
        const float inv_i1 = b1.is_static ? 0 : (1 / b1.props.inertia);
        const float inv_i2 = b2.is_static ? 0 : (1 / b2.props.inertia);
        tangent = rotate90clockwise(normal);
        friction = sqrtf(b1.props.mu * b2.props.mu);
        for (int k = 0; k < 2; ++k) {
            if (!points[k].present) continue;
            CollisionPoint& point = points[k].p;
            const Vector2 r1 = point.ref_r, r2 = point.inc_r;
            const float r1n = cross2(r1, normal);
            const float r2n = cross2(r2, normal);
            const float kn = inv_mass + inv_i1 * (r1n * r1n) + inv_i2 * (r2n * r2n);
            const float r1t = cross2(r1, tangent);
            const float r2t = cross2(r2, tangent);
            const float kt = inv_mass + inv_i1 * (r1t * r1t) + inv_i2 * (r2t * r2t);
            point.mass_n = (kn > 0.0f) ? (1 / kn) : 0.0f;
            point.mass_t = (kt > 0.0f) ? (1 / kt) : 0.0f;
        }
    }

    void calculateImpulses(std::vector<RigidBody>& B, float dt) {  // :135-218
        RigidBody& b1 = B[ref_body];
        RigidBody& b2 = B[inc_body];
        const float inv_m1 = b1.is_static ? 0 : (1 / b1.props.mass);
        const float inv_m2 = b2.is_static ? 0 : (1 / b2.props.mass);
        const float inv_i1 = b1.is_static ? 0 : (1 / b1.props.inertia);
        const float inv_i2 = b2.is_static ? 0 : (1 / b2.props.inertia);
        const Vector2 vlinear_1 = scale2(b1.props.momentum, inv_m1);
        const float omega1 = b1.props.ang_momentum * inv_i1;
        const Vector2 vlinear_2 = scale2(b2.props.momentum, inv_m2);
        const float omega2 = b2.props.ang_momentum * inv_i2;
        for (int k = 0; k < 2; ++k) {
            if (!points[k].present) continue;
            CollisionPoint& point = points[k].p;
            if (point.depth >= 0) {
                point.accumulated_pn = 0;
                point.accumulated_pt = 0;
                continue;
            }
            const Vector2 r1 = point.ref_r, r2 = point.inc_r;
            const Vector2 vrot_1 = Vector2(-r1.y * omega1, r1.x * omega1);
            const Vector2 v1 = add2(vlinear_1, vrot_1);
            const Vector2 vrot_2 = Vector2(-r2.y * omega2, r2.x * omega2);
            const Vector2 v2 = add2(vlinear_2, vrot_2);
            const Vector2 dv = sub2(v1, v2);
            const float bias = BAUMGARTE * zmax(0, (-point.depth) - BAUMGARTE_SLOP) / dt;
            float num = dot2(dv, normal) + bias;
            float pn = num * point.mass_n;
            if (warm_first) pn = pn + warm_pn[k];
            if (pn < MIN_MANIFOLD_IMPULSE) continue;
            num = dot2(dv, tangent);
            float pt = num * point.mass_t;
            if (warm_first) pt = pt + warm_pt[k];
            const float new_accumulated_pn = zmax(0, point.accumulated_pn + pn);
            const float applied_pn = new_accumulated_pn - point.accumulated_pn;
            point.accumulated_pn = new_accumulated_pn;
            const float max_pt = friction * zabs(point.accumulated_pn);
            const float new_accumulated_pt = zclamp(point.accumulated_pt + pt, -max_pt, max_pt);
            const float applied_pt = new_accumulated_pt - point.accumulated_pt;
            point.accumulated_pt = new_accumulated_pt;
            const Vector2 pn_vec = scale2(normal, applied_pn);
            const Vector2 pt_vec = scale2(tangent, applied_pt);
            const Vector2 dp = add2(pn_vec, pt_vec);
            if (!b1.is_static) {
                applied_linear_p1.sub(dp);
                applied_rot_p_1 -= cross2(r1, dp);
            }
            if (!b2.is_static) {
                applied_linear_p2.add(dp);
                applied_rot_p_2 += cross2(r2, dp);
            }
        }
        b1.props.momentum.add(applied_linear_p1);
        b1.props.ang_momentum += applied_rot_p_1;
        b2.props.momentum.add(applied_linear_p2);
        b2.props.ang_momentum += applied_rot_p_2;
        applied_linear_p1 = Vector2();
        applied_rot_p_1 = 0;
        applied_linear_p2 = Vector2();
        applied_rot_p_2 = 0;
        warm_first = false;
    }
};

// ---- Constraints/*.zig -----------------------------------------------------------------------------------------------
enum JointType { DISTANCE = 0, OFFSET_DISTANCE = 1, FIXED_POSITION = 2, MOTOR = 3 };
struct Constraint {
    JointType type;
    float power_max, power_min, beta;  // Constraint.zig:27-31
    uint32_t id1 = 0, id2 = 0;
    Vector2 r1, r2;
    float target_distance = 0;
    Vector2 target_position;
    float target_omega = 0;
    uint32_t color = 0;
};

// ---- SpatialHash.zig -------------------------------------------------------------------------------------------------
struct SpatialHash {
    std::vector<size_t> table;
    std::vector<size_t> body_indices;  // indices into `bodies` instead of *RigidBody
    size_t table_size;
    float cell_size;

    static size_t hash(size_t table_size, int64_t xi, int64_t yi) {  // :78-81
        const uint64_t h = (uint64_t)((int64_t)((uint64_t)xi * 92837111ull) ^ (int64_t)((uint64_t)yi * 689287499ull));
        return (size_t)(h % table_size);
    }
    template <class F>
    void iterateAABBHashes(const RigidBody& body, F onCell) const {  // :83-106
        Vector2 verts[4];
        body.aabb.getVertices(verts);
        const float minx = verts[0].x, miny = verts[0].y, maxx = verts[2].x, maxy = verts[2].y;
        const int64_t min_xi = (int64_t)floorf(minx / cell_size);
        const int64_t min_yi = (int64_t)floorf(miny / cell_size);
        const int64_t max_xi = (int64_t)floorf(maxx / cell_size);
        const int64_t max_yi = (int64_t)floorf(maxy / cell_size);
        for (int64_t yi = min_yi; yi <= max_yi; ++yi)
            for (int64_t xi = min_xi; xi <= max_xi; ++xi) onCell(hash(table_size, xi, yi));
    }
    SpatialHash(float cell_size_, size_t table_size_, const std::vector<RigidBody>& bodies)  // init :19-71
        : table(table_size_ + 1, 0), table_size(table_size_), cell_size(cell_size_) {
        for (const RigidBody& b : bodies) iterateAABBHashes(b, [&](size_t id) { table[id] += 1; });
        size_t start = 0;
        for (size_t id = 0; id < table_size; ++id) {
            start += table[id];
            table[id] = start;
        }
        table[table_size] = start;
        body_indices.assign(start, 0);
        for (size_t bi = 0; bi < bodies.size(); ++bi)
            iterateAABBHashes(bodies[bi], [&](size_t id) {
                table[id] -= 1;
                body_indices[table[id]] = bi;
            });
    }
    void query(const RigidBody& target, std::vector<size_t>* res) const {  // :108-137
        iterateAABBHashes(target, [&](size_t id) {
            const size_t curr = table[id];
            const size_t next = table[id + 1];
            if (curr != next)
                for (size_t idx = curr; idx < next; ++idx) res->push_back(body_indices[idx]);
        });
    }
};

struct PairHash {
    size_t operator()(const std::pair<size_t, size_t>& p) const { return p.first * 0x9E3779B97F4A7C15ull ^ (p.second + 0x7F4A7C15ull); }
};

// Colouring spec shared (as a specification, not as code) with the CUDA path — DESIGN.md "Gauss-Seidel order".
static uint32_t mix32(uint32_t x) {
    x ^= x >> 16;
    x *= 0x7feb352du;
    x ^= x >> 15;
    x *= 0x846ca68bu;
    x ^= x >> 16;
    return x;
}
static uint64_t contactPriority(uint32_t id_a, uint32_t id_b) {
    const uint32_t lo = id_a < id_b ? id_a : id_b, hi = id_a < id_b ? id_b : id_a;
    const uint32_t h = mix32(mix32(lo) ^ (hi * 0x9e3779b9u));
    return ((uint64_t)(h >> 12) << 32) | (uint64_t)(uint32_t)(lo + hi);
}

enum GsOrder { ORDER_REFERENCE = 0, ORDER_COLORED = 1 };
static const float SLEEP_LIN2 = 0.04f, SLEEP_ANG2 = 0.04f;   // (0.2 m/s)^2, (0.2 rad/s)^2

// ---- lib.zig: Solver ----------------------------------------------------------------------------------------------------
struct Solver {
    uint32_t current_body_id = 0;
    std::vector<RigidBody> bodies;                   // AutoArrayHashMap(Id, RigidBody): insertion order, swapRemove
    std::unordered_map<uint32_t, size_t> body_index;
    std::vector<float> force_generators;             // DownwardsGravity g values (Forces/DownwardsGravity.zig)
    std::vector<CollisionManifold> manifolds;        // AutoArrayHashMap(CollisionKey, CollisionManifold): insertion order
    std::unordered_map<std::pair<size_t, size_t>, size_t, PairHash> manifold_index;
    std::set<std::pair<uint32_t, uint32_t>> exclude_collision_pairs;
    std::vector<Constraint> constraints;
    float spatialhash_cell_width;
    size_t spatialhash_table_size_mult;
    int gs_order = ORDER_REFERENCE;
    // ---- roadmap options (README.md:59-64; not in the reference, off by default) ----
    bool opt_warm_start = false, opt_sleeping = false;
    uint32_t opt_sleep_calls = 30;
    struct WarmEntry { uint32_t normal_id, n_points; float pn[2], pt[2]; };
    std::unordered_map<uint64_t, WarmEntry> warm_store;   // key: ref id << 32 | inc id (stable ids, not pointers)
    std::unordered_map<uint32_t, uint32_t> sleep_counter; // id -> consecutive calls below the speed thresholds
    // instrumentation
    std::set<std::pair<uint32_t, uint32_t>> candidate_pairs;  // the set C of SURVEY A.2, logged at lib.zig:282
    size_t stat_entries = 0, stat_raw_candidates = 0;
    uint32_t stat_colors = 0, stat_joint_colors = 0, stat_dropped = 0;
    std::vector<size_t> manifold_order, joint_order;          // sweep order actually used by the last process()

    RigidBody* find(uint32_t id) {
        auto it = body_index.find(id);
        return it == body_index.end() ? nullptr : &bodies[it->second];
    }

    int process(float dt, size_t sub_steps, size_t collision_iters) {  // lib.zig:189-251
        for (const Constraint& c : constraints) {  // declared deviation: fail before touching state
            if (!find(c.id1)) return -2;
            if ((c.type == DISTANCE || c.type == OFFSET_DISTANCE) && !find(c.id2)) return -2;
        }
        const float f32_sub = (float)sub_steps;
        const float sub_dt = dt / f32_sub;
        // sleeping: a body that has been slow for opt_sleep_calls calls in a row and is not being pushed IS A STATIC BODY
        // for the duration of this call (broadphase filters, contacts, joints, integration)
        std::vector<size_t> asleep;
        if (opt_sleeping) {
            std::set<uint32_t> jointed;   // a body named by a joint never sleeps (a distance joint between two static bodies
            for (const Constraint& c : constraints) {   // divides by w1 + w2 = 0)
                jointed.insert(c.id1);
                if (c.type == DISTANCE || c.type == OFFSET_DISTANCE) jointed.insert(c.id2);
            }
            for (size_t i = 0; i < bodies.size(); ++i) {
                RigidBody& b = bodies[i];
                if (b.is_static) continue;
                uint32_t& cnt = sleep_counter[b.id];
                if (b.props.force.x != 0 || b.props.force.y != 0 || b.props.torque != 0) cnt = 0;   // user input wakes
                if (jointed.count(b.id)) cnt = 0;
                if (cnt >= opt_sleep_calls) {
                    b.is_static = true;
                    asleep.push_back(i);
                }
            }
        }
        updateManifolds();
        if (opt_warm_start)
            for (CollisionManifold& m : manifolds) {
                auto it = warm_store.find(((uint64_t)bodies[m.ref_body].id << 32) | bodies[m.inc_body].id);
                if (it == warm_store.end()) continue;
                const uint32_t np = (m.points[0].present ? 1u : 0u) + (m.points[1].present ? 1u : 0u);
                if (it->second.normal_id != (uint32_t)m.reference_normal_id || it->second.n_points != np) continue;
                m.warm_first = true;
                for (int k = 0; k < 2; ++k) {
                    m.warm_pn[k] = it->second.pn[k];
                    m.warm_pt[k] = it->second.pt[k];
                }
            }
        buildSweepOrder();
        for (size_t s = 0; s < sub_steps; ++s) {
            for (float g : force_generators)  // :200-205
                for (RigidBody& body : bodies) {
                    if (body.is_static) continue;  // DownwardsGravity.zig:35-39
                    body.props.force.addmult(Vector2(0, -g), body.props.mass);
                }
            for (RigidBody& body : bodies) {  // :207-216
                updateAABB(body);
                if (body.is_static) continue;
                body.props.momentum.addmult(body.props.force, sub_dt);
                body.props.ang_momentum += body.props.torque * sub_dt;
            }
            for (CollisionManifold& m : manifolds) {  // :218-224
                m.updateTGSDepth(bodies);
                m.preStep(bodies);
            }
            for (size_t it = 0; it < collision_iters; ++it) {  // :226-236
                for (size_t ci : joint_order) solveConstraint(constraints[ci], sub_dt);
                for (size_t mi : manifold_order) manifolds[mi].calculateImpulses(bodies, sub_dt);
            }
            for (RigidBody& body : bodies) {  // :238-249
                if (body.is_static) continue;
                Props& props = body.props;
                props.pos.addmult(props.momentum, sub_dt / props.mass);
                props.angle += props.ang_momentum * sub_dt / props.inertia;
                props.force = Vector2();
                props.torque = 0;
            }
        }
        if (opt_warm_start) {   // remember what every contact ended the call with, per substep
            warm_store.clear();
            for (const CollisionManifold& m : manifolds) {
                WarmEntry e{};
                e.normal_id = (uint32_t)m.reference_normal_id;
                e.n_points = (m.points[0].present ? 1u : 0u) + (m.points[1].present ? 1u : 0u);
                for (int k = 0; k < 2; ++k)
                    if (m.points[k].present) {
                        e.pn[k] = m.points[k].p.accumulated_pn / f32_sub;
                        e.pt[k] = m.points[k].p.accumulated_pt / f32_sub;
                    }
                warm_store[((uint64_t)bodies[m.ref_body].id << 32) | bodies[m.inc_body].id] = e;
            }
        }
        if (opt_sleeping) {
            for (size_t i : asleep) bodies[i].is_static = false;
            std::vector<unsigned char> moving(bodies.size(), 0), was_asleep(bodies.size(), 0);
            for (size_t i : asleep) was_asleep[i] = 1;
            for (size_t i = 0; i < bodies.size(); ++i) {
                RigidBody& b = bodies[i];
                if (b.is_static || was_asleep[i]) continue;
                const float vx = b.props.momentum.x / b.props.mass, vy = b.props.momentum.y / b.props.mass;
                const float w = b.props.ang_momentum / b.props.inertia;
                const bool slow = (vx * vx + vy * vy) < SLEEP_LIN2 && (w * w) < SLEEP_ANG2;
                uint32_t& cnt = sleep_counter[b.id];
                cnt = slow ? cnt + 1 : 0;
                moving[i] = slow ? 0 : 1;
            }
            // one hop: a body that moved fast in this call wakes what it touches (the woken bodies do not wake others yet)
            for (const CollisionManifold& m : manifolds) {
                if (moving[m.ref_body] && !bodies[m.inc_body].is_static) sleep_counter[bodies[m.inc_body].id] = 0;
                if (moving[m.inc_body] && !bodies[m.ref_body].is_static) sleep_counter[bodies[m.ref_body].id] = 0;
            }
        }
        return 0;
    }

    void updateManifolds() {  // lib.zig:253-299
        float cell = 4.0f;  // FIXME in the reference: hard-coded (Q2)
        size_t table = 2 * bodies.size();
        if (use_init_grid_params) {
            cell = spatialhash_cell_width;
            table = spatialhash_table_size_mult * bodies.size();
        }
        manifolds.clear();
        manifold_index.clear();
        candidate_pairs.clear();
        stat_raw_candidates = 0;
        stat_entries = 0;
        if (bodies.empty()) return;
        SpatialHash spatial(cell, table, bodies);
        stat_entries = spatial.body_indices.size();
        std::vector<size_t> queries;
        for (size_t i1 = 0; i1 < bodies.size(); ++i1) {
            const RigidBody& body1 = bodies[i1];
            queries.clear();
            spatial.query(body1, &queries);
            stat_raw_candidates += queries.size();
            for (size_t i2 : queries) {
                const RigidBody& body2 = bodies[i2];
                if (body1.is_static && body2.is_static) continue;
                if (i1 == i2) continue;
                if (exclude_collision_pairs.count({body1.id, body2.id})) continue;
                if (exclude_collision_pairs.count({body2.id, body1.id})) continue;
                if (manifold_index.count({i1, i2})) continue;
                if (manifold_index.count({i2, i1})) continue;
                if (!body1.aabb.intersects(body2.aabb)) continue;
                candidate_pairs.insert({std::min(body1.id, body2.id), std::max(body1.id, body2.id)});
                const SATResult sat = performNarrowSAT(bodies, i1, i2);
                if (!sat.collides) continue;
                CollisionManifold m;
                m.ref_body = sat.key_ref;
                m.inc_body = sat.key_inc;
                m.reference_normal_id = sat.reference_normal_id;
                m.normal = sat.normal;
                identifyCollisionPoints(bodies[sat.key_ref], bodies[sat.key_inc], sat.reference_normal_id, m.points);
                m.prev_angle_1 = bodies[sat.key_ref].props.angle;
                m.prev_angle_2 = bodies[sat.key_inc].props.angle;
                manifold_index[{sat.key_ref, sat.key_inc}] = manifolds.size();
                manifolds.push_back(m);
            }
        }
    }
    bool use_init_grid_params = false;

    // Sweep order of the iteration loop.  ORDER_REFERENCE: insertion order (lib.zig:227-235).  ORDER_COLORED: the
    // deterministic graph colouring specified in DESIGN.md, restated sequentially:
    //   contacts — greedy colouring in DESCENDING contactPriority(ref.id, inc.id); a manifold conflicts with every
    //              manifold that shares one of its NON-STATIC bodies; colour = smallest free; sweep by (colour, any).
    //   joints   — greedy colouring in list order; two joints conflict if they name a common body (static or not).
    void buildSweepOrder() {
        manifold_order.resize(manifolds.size());
        joint_order.resize(constraints.size());
        for (size_t i = 0; i < manifolds.size(); ++i) manifold_order[i] = i;
        for (size_t i = 0; i < constraints.size(); ++i) joint_order[i] = i;
        stat_colors = 0;
        stat_joint_colors = 0;
        stat_dropped = 0;
        for (CollisionManifold& m : manifolds) m.color = 0;
        for (Constraint& c : constraints) c.color = 0;
        if (gs_order != ORDER_COLORED) return;
        {
            std::vector<size_t> by_prio(manifolds.size());
            std::vector<uint64_t> prio(manifolds.size());
            for (size_t i = 0; i < manifolds.size(); ++i) {
                by_prio[i] = i;
                prio[i] = contactPriority(bodies[manifolds[i].ref_body].id, bodies[manifolds[i].inc_body].id);
            }
            std::sort(by_prio.begin(), by_prio.end(), [&](size_t a, size_t b) { return prio[a] > prio[b]; });
            std::vector<std::vector<uint32_t>> used(bodies.size());  // colours taken per body
            for (size_t mi : by_prio) {
                CollisionManifold& m = manifolds[mi];
                const size_t bs[2] = {m.ref_body, m.inc_body};
                uint32_t c = 0;
                for (;; ++c) {
                    bool taken = false;
                    for (size_t b : bs) {
                        if (bodies[b].is_static) continue;
                        if (std::find(used[b].begin(), used[b].end(), c) != used[b].end()) taken = true;
                    }
                    if (!taken) break;
                }
                // The CUDA path has 256 colours (a hard limit of the boundary, include/r2d_abi.h R2D_MAX_COLORS): a manifold
                // that finds none free — a body with more than 256 simultaneous contacts — is left out of this call's sweeps
                // (it keeps its place in the manifold list, colour R2D_COLOR_DROPPED).
                if (c >= 256) {
                    m.color = 0xFFFFFFFDu;
                    stat_dropped += 1;
                    continue;
                }
                m.color = c;
                for (size_t b : bs)
                    if (!bodies[b].is_static) used[b].push_back(c);
                stat_colors = std::max(stat_colors, c + 1);
            }
            std::stable_sort(manifold_order.begin(), manifold_order.end(),
                             [&](size_t a, size_t b) { return manifolds[a].color < manifolds[b].color; });
            while (!manifold_order.empty() && manifolds[manifold_order.back()].color == 0xFFFFFFFDu) manifold_order.pop_back();
        }
        {
            std::unordered_map<uint32_t, std::vector<uint32_t>> used;
            for (Constraint& c : constraints) {
                const bool two = (c.type == DISTANCE || c.type == OFFSET_DISTANCE);
                uint32_t col = 0;
                for (;; ++col) {
                    bool taken = false;
                    auto chk = [&](uint32_t id) {
                        auto& u = used[id];
                        if (std::find(u.begin(), u.end(), col) != u.end()) taken = true;
                    };
                    chk(c.id1);
                    if (two) chk(c.id2);
                    if (!taken) break;
                }
                c.color = col;
                used[c.id1].push_back(col);
                if (two && c.id2 != c.id1) used[c.id2].push_back(col);
                stat_joint_colors = std::max(stat_joint_colors, col + 1);
            }
            std::stable_sort(joint_order.begin(), joint_order.end(),
                             [&](size_t a, size_t b) { return constraints[a].color < constraints[b].color; });
        }
    }

    void solveConstraint(Constraint& ctr, float dt) {
        switch (ctr.type) {
            case DISTANCE: {  // DistanceJoint.zig:40-75
                const float epsilon = CONSTRAINT_GRADIENT_DIVISION_LIMIT;
                RigidBody* b1 = find(ctr.id1);
                RigidBody* b2 = find(ctr.id2);
                const float w1 = b1->is_static ? 0 : (1 / b1->props.mass);
                const float w2 = b2->is_static ? 0 : (1 / b2->props.mass);
                const Vector2 v1 = scale2(b1->props.momentum, w1);
                const Vector2 v2 = scale2(b2->props.momentum, w2);
                const Vector2 relative_v = sub2(v2, v1);
                const Vector2 normal = normalize2(sub2(b2->props.pos, b1->props.pos));
                const float dist_error = length2(sub2(b2->props.pos, b1->props.pos)) - ctr.target_distance;
                const float beta = ctr.beta;
                const float position_correction = beta * dist_error / zmax(ctr.target_distance, 0.001f);
                float J = -(dot2(relative_v, normal) + position_correction) / (w1 + w2);
                const float ndv = dot2(normal, relative_v);
                const float den = 1 / (zmax(zabs(ndv), epsilon));
                const float J_max = ctr.power_max * den;
                const float J_min = ctr.power_min * den;
                J = zclamp(J, J_min, J_max);
                const Vector2 dp = scale2(normal, J);
                b1->props.momentum.sub(dp);
                b2->props.momentum.add(dp);
            } break;
            case OFFSET_DISTANCE: {  // OffsetDistanceJoint.zig:44-105
                const float epsilon = CONSTRAINT_GRADIENT_DIVISION_LIMIT;
                RigidBody* b1 = find(ctr.id1);
                RigidBody* b2 = find(ctr.id2);
                const Vector2 a1 = localToWorld(*b1, ctr.r1);
                const Vector2 a2 = localToWorld(*b2, ctr.r2);
                const Vector2 normal = normalize2(sub2(a2, a1));
                const float inv_m1 = b1->is_static ? 0 : (1 / b1->props.mass);
                const float inv_m2 = b2->is_static ? 0 : (1 / b2->props.mass);
                const float inv_i1 = b1->is_static ? 0 : (1 / b1->props.inertia);
                const float inv_i2 = b2->is_static ? 0 : (1 / b2->props.inertia);
                const Vector2 vlinear_1 = scale2(b1->props.momentum, inv_m1);
                const float omega1 = b1->props.ang_momentum * inv_i1;
                const Vector2 vlinear_2 = scale2(b2->props.momentum, inv_m2);
                const float omega2 = b2->props.ang_momentum * inv_i2;
                const Vector2 rotated_r1 = rotate2(ctr.r1, b1->props.angle);
                const Vector2 v_at_a1 = Vector2(-rotated_r1.y * omega1, rotated_r1.x * omega1);
                const Vector2 v1 = add2(vlinear_1, v_at_a1);
                const Vector2 rotated_r2 = rotate2(ctr.r2, b2->props.angle);
                const Vector2 v_at_a2 = Vector2(-rotated_r2.y * omega2, rotated_r2.x * omega2);
                const Vector2 v2 = add2(vlinear_2, v_at_a2);
                const Vector2 dv = sub2(v2, v1);
                const float dist_error = length2(sub2(a2, a1)) - ctr.target_distance;
                const float beta = ctr.beta;
                const float position_correction = beta * dist_error / zmax(ctr.target_distance, epsilon);
                const float num = dot2(normal, dv) + position_correction;
                const float r1xn = cross2(rotated_r1, normal);
                const float r2xn = cross2(rotated_r2, normal);
                const float den = inv_m1 + inv_m2 + r1xn * r1xn * inv_i1 + r2xn * r2xn * inv_i2;
                float J = -num / den;
                const float ndv = dot2(normal, dv);
                const float den2 = 1 / (zmax(zabs(ndv), epsilon));
                const float J_max = ctr.power_max * den2;
                const float J_min = ctr.power_min * den2;
                J = zclamp(J, J_min, J_max);
                const Vector2 dp = scale2(normal, J);
                b1->props.momentum.sub(dp);
                b1->props.ang_momentum -= J * r1xn;
                b2->props.momentum.add(dp);
                b2->props.ang_momentum += J * r2xn;
            } break;
            case FIXED_POSITION: {  // FixedPositionJoint.zig:38-72
                const float epsilon = ALLOWED_CONSTRAINT_VALUE;
                RigidBody* b = find(ctr.id1);
                if (b->is_static) return;
                const float w = 1 / b->props.mass;
                const Vector2 v = scale2(b->props.momentum, w);
                const Vector2 delta_pos = sub2(ctr.target_position, b->props.pos);
                const float dist = length2(delta_pos);
                if (dist < ALLOWED_CONSTRAINT_VALUE) return;
                const Vector2 normal = scale2(delta_pos, 1 / dist);
                const float beta = ctr.beta;
                const float bias = beta * dist;
                const float relative_velocity = dot2(v, normal);
                float J = -(relative_velocity - bias) / w;
                const float den = 1 / (zmax(zabs(relative_velocity), epsilon));
                const float J_max = ctr.power_max * den;
                const float J_min = ctr.power_min * den;
                J = zclamp(J, J_min, J_max);
                const Vector2 dp = scale2(normal, J);
                b->props.momentum.add(dp);
            } break;
            case MOTOR: {  // MotorJoint.zig:38-64
                const float epsilon = ALLOWED_CONSTRAINT_VALUE;
                RigidBody* b = find(ctr.id1);
                if (b->is_static) return;
                const float omega = b->props.ang_momentum / b->props.inertia;
                const float C = omega - ctr.target_omega;
                if (zabs(C) < ALLOWED_CONSTRAINT_VALUE) return;
                const float beta = ctr.beta;
                const float bias = beta * C;
                float J = b->props.torque - bias;
                const float relative_velocity = b->props.torque / b->props.inertia;
                const float den = 1 / (zmax(zabs(relative_velocity), epsilon));
                const float J_max = ctr.power_max * den;
                const float J_min = ctr.power_min * den;
                J = zclamp(J, J_min, J_max);
                b->props.ang_momentum += J * dt;
            } break;
        }
    }

    // ---- EntityFactory (lib.zig:66-129) ---------------------------------------------------------------------------
    uint32_t appendBody(RigidBody body) {  // :66-71
        const uint32_t id = current_body_id;
        body.id = id;
        body_index[id] = bodies.size();
        bodies.push_back(body);
        current_body_id += 1;
        return id;
    }
    uint32_t makeDiscBody(Vector2 pos, Vector2 vel, float angle, float omega, float mu, bool is_density, float mass_value, float radius) {
        const float mass = is_density ? ((float)M_PI * radius * radius * mass_value) : mass_value;  // :74-77
        RigidBody b;  // Disc.init (Disc.zig:29-56)
        b.radius = radius;
        b.props.pos = pos;
        b.props.angle = angle;
        b.props.mass = mass;
        b.props.inertia = 0.5f * mass * radius * radius;
        b.props.mu = mu;
        b.num_normals = 1;
        b.type = DISC;
        updateAABB(b);
        b.props.momentum = scale2(vel, mass);           // :79
        b.props.ang_momentum = omega * b.props.inertia; // :80
        return appendBody(b);
    }
    uint32_t makeRectangleBody(Vector2 pos, Vector2 vel, float angle, float omega, float mu, bool is_density, float mass_value, float width, float height) {
        const float mass = is_density ? (width * height * mass_value) : mass_value;  // :85-88
        RigidBody b;  // Rectangle.init (Rectangle.zig:33-64)
        b.width = width;
        b.height = height;
        const float w = width / 2, h = height / 2;
        b.local_vertices[0] = Vector2(-w, -h);
        b.local_vertices[1] = Vector2(-w, h);
        b.local_vertices[2] = Vector2(w, h);
        b.local_vertices[3] = Vector2(w, -h);
        b.props.pos = pos;
        b.props.angle = angle;
        b.props.mass = mass;
        b.props.inertia = mass * (width * width + height * height) / 12;
        b.props.mu = mu;
        b.num_normals = 4;
        b.type = RECTANGLE;
        updateAABB(b);
        b.props.momentum = scale2(vel, mass);
        b.props.ang_momentum = omega * b.props.inertia;
        return appendBody(b);
    }
    bool removeRigidBody(uint32_t id) {  // :308-310 swapRemove
        auto it = body_index.find(id);
        if (it == body_index.end()) return false;
        const size_t idx = it->second;
        body_index.erase(it);
        if (idx != bodies.size() - 1) {
            bodies[idx] = bodies.back();
            body_index[bodies[idx].id] = idx;
        }
        bodies.pop_back();
        return true;
    }
    void clear() {  // :181-187 (Q16): exclusions and the id counter survive
        bodies.clear();
        body_index.clear();
        force_generators.clear();
        manifolds.clear();
        manifold_index.clear();
        constraints.clear();
    }
};

}  // namespace

// =====================================================================================================================
// C ABI for ctypes (tests, smoke, bench cpu_baseline).  Mirrors include/r2d_abi.h with the prefix orc_.
// =====================================================================================================================
extern "C" {

struct orc_body_opts {  // == r2d_body_opts
    float pos_x, pos_y, vel_x, vel_y, angle, omega, mu, mass_value;
    int32_t mass_is_density;
};
struct orc_body_desc {  // == r2d_body_desc
    orc_body_opts opts;
    int32_t shape;
    float a, b;
    int32_t is_static;
};
struct orc_joint_params {
    float power_max, power_min, beta;
};
struct orc_manifold {  // == r2d_manifold
    uint32_t ref_id, inc_id, normal_id, n_points;
    float normal_x, normal_y;
    float pos_x[2], pos_y[2], depth[2], ref_rx[2], ref_ry[2], inc_rx[2], inc_ry[2];
    uint32_t color;
};
struct orc_body_state {  // == r2d_body_state
    uint32_t id;
    int32_t shape, is_static;
    float pos_x, pos_y, angle, momentum_x, momentum_y, ang_momentum, force_x, force_y, torque, mass, inertia, mu;
    float aabb_x, aabb_y, aabb_half_w, aabb_half_h, shape_a, shape_b;
};
struct orc_step_stats {  // == r2d_step_stats
    uint32_t n_bodies, n_buckets, n_entries, n_pairs, n_manifolds, n_points, n_colors, n_color_rounds, n_joints,
        n_joint_colors, n_launches, n_dropped;
};

int orc_create(float cell_width, uint32_t table_mult, int /*device*/, void** out) {
    Solver* s = new Solver();
    s->spatialhash_cell_width = cell_width;
    s->spatialhash_table_size_mult = table_mult;
    *out = s;
    return 0;
}
int orc_destroy(void* h) {
    delete (Solver*)h;
    return 0;
}
int orc_clear(void* h) {
    ((Solver*)h)->clear();
    return 0;
}
// mode 0: reference behaviour (cell 4.0, 2N buckets); 1: honour cell_width/table_mult ("fast" mode of the CUDA path)
int orc_set_mode(void* h, int mode) {
    ((Solver*)h)->use_init_grid_params = (mode != 0);
    return 0;
}
// roadmap options (r2d_set_option): 1 warm start, 2 sleeping, 3 calls below the thresholds before a body sleeps
int orc_set_option(void* h, int option, uint32_t value) {
    Solver* s = (Solver*)h;
    if (option == 1) s->opt_warm_start = value != 0;
    else if (option == 2) s->opt_sleeping = value != 0;
    else if (option == 3) s->opt_sleep_calls = value;
    else return -4;
    return 0;
}
// 0: reference insertion order; 1: the coloured order the CUDA path sweeps in
int orc_set_gs_order(void* h, int order) {
    ((Solver*)h)->gs_order = order;
    return 0;
}

static uint32_t make_one(Solver* s, const orc_body_opts& o, int shape, float a, float b, int is_static) {
    uint32_t id;
    if (shape == 0)
        id = s->makeDiscBody(Vector2(o.pos_x, o.pos_y), Vector2(o.vel_x, o.vel_y), o.angle, o.omega, o.mu, o.mass_is_density != 0, o.mass_value, a);
    else
        id = s->makeRectangleBody(Vector2(o.pos_x, o.pos_y), Vector2(o.vel_x, o.vel_y), o.angle, o.omega, o.mu, o.mass_is_density != 0, o.mass_value, a, b);
    if (is_static) s->find(id)->is_static = true;
    return id;
}
int orc_make_disc(void* h, const orc_body_opts* o, float radius, uint32_t* out_id) {
    const uint32_t id = make_one((Solver*)h, *o, 0, radius, 0, 0);
    if (out_id) *out_id = id;
    return 0;
}
int orc_make_rect(void* h, const orc_body_opts* o, float width, float height, uint32_t* out_id) {
    const uint32_t id = make_one((Solver*)h, *o, 1, width, height, 0);
    if (out_id) *out_id = id;
    return 0;
}
int orc_make_bodies(void* h, const orc_body_desc* descs, size_t n, uint32_t* out_first_id) {
    Solver* s = (Solver*)h;
    if (out_first_id) *out_first_id = s->current_body_id;
    s->bodies.reserve(s->bodies.size() + n);
    for (size_t i = 0; i < n; ++i) make_one(s, descs[i].opts, descs[i].shape, descs[i].a, descs[i].b, descs[i].is_static);
    return 0;
}
int orc_make_gravity(void* h, float g) {
    ((Solver*)h)->force_generators.push_back(g);
    return 0;
}
static Constraint base_joint(JointType t, const orc_joint_params* p) {
    Constraint c;
    c.type = t;
    c.power_max = p ? p->power_max : INFINITY;
    c.power_min = p ? p->power_min : -INFINITY;
    c.beta = p ? p->beta : 10.0f;
    return c;
}
int orc_make_distance_joint(void* h, const orc_joint_params* p, uint32_t id1, uint32_t id2, float target, size_t* out_index) {
    Solver* s = (Solver*)h;
    Constraint c = base_joint(DISTANCE, p);
    c.id1 = id1;
    c.id2 = id2;
    c.target_distance = target;
    s->constraints.push_back(c);
    if (out_index) *out_index = s->constraints.size() - 1;
    return 0;
}
int orc_make_offset_distance_joint(void* h, const orc_joint_params* p, uint32_t id1, uint32_t id2, float r1x, float r1y, float r2x, float r2y, float target, size_t* out_index) {
    Solver* s = (Solver*)h;
    Constraint c = base_joint(OFFSET_DISTANCE, p);
    c.id1 = id1;
    c.id2 = id2;
    c.r1 = Vector2(r1x, r1y);
    c.r2 = Vector2(r2x, r2y);
    c.target_distance = target;
    s->constraints.push_back(c);
    if (out_index) *out_index = s->constraints.size() - 1;
    return 0;
}
int orc_make_fixed_position_joint(void* h, const orc_joint_params* p, uint32_t id, float tx, float ty, size_t* out_index) {
    Solver* s = (Solver*)h;
    Constraint c = base_joint(FIXED_POSITION, p);
    c.id1 = id;
    c.target_position = Vector2(tx, ty);
    s->constraints.push_back(c);
    if (out_index) *out_index = s->constraints.size() - 1;
    return 0;
}
int orc_make_motor_joint(void* h, const orc_joint_params* p, uint32_t id, float omega, size_t* out_index) {
    Solver* s = (Solver*)h;
    Constraint c = base_joint(MOTOR, p);
    c.id1 = id;
    c.target_omega = omega;
    s->constraints.push_back(c);
    if (out_index) *out_index = s->constraints.size() - 1;
    return 0;
}
int orc_exclude_pair(void* h, uint32_t id1, uint32_t id2) {  // lib.zig:124-129 (both orders, Q22)
    Solver* s = (Solver*)h;
    s->exclude_collision_pairs.insert({id1, id2});
    s->exclude_collision_pairs.insert({id2, id1});
    return 0;
}
int orc_remove_body(void* h, uint32_t id) { return ((Solver*)h)->removeRigidBody(id) ? 0 : -3; }
int orc_process(void* h, float dt, uint32_t sub_steps, uint32_t iters) { return ((Solver*)h)->process(dt, sub_steps, iters); }
int orc_step(void* h, float dt, uint32_t sub_steps, uint32_t iters) { return orc_process(h, dt, sub_steps, iters); }
int orc_synchronize(void*) { return 0; }
int orc_reorder(void*) { return 0; }                       // memory order is not a concept of the oracle
int orc_set_reorder_interval(void*, uint32_t) { return 0; }

int orc_num_bodies(void* h, size_t* out) {
    *out = ((Solver*)h)->bodies.size();
    return 0;
}
int orc_body_id_at(void* h, size_t i, uint32_t* out_id) {
    Solver* s = (Solver*)h;
    if (i >= s->bodies.size()) return -4;
    *out_id = s->bodies[i].id;
    return 0;
}
int orc_body_get(void* h, uint32_t id, orc_body_state* o) {
    RigidBody* b = ((Solver*)h)->find(id);
    if (!b) return -3;
    o->id = b->id;
    o->shape = b->type;
    o->is_static = b->is_static;
    o->pos_x = b->props.pos.x;
    o->pos_y = b->props.pos.y;
    o->angle = b->props.angle;
    o->momentum_x = b->props.momentum.x;
    o->momentum_y = b->props.momentum.y;
    o->ang_momentum = b->props.ang_momentum;
    o->force_x = b->props.force.x;
    o->force_y = b->props.force.y;
    o->torque = b->props.torque;
    o->mass = b->props.mass;
    o->inertia = b->props.inertia;
    o->mu = b->props.mu;
    o->aabb_x = b->aabb.pos.x;
    o->aabb_y = b->aabb.pos.y;
    o->aabb_half_w = b->aabb.half_width;
    o->aabb_half_h = b->aabb.half_height;
    o->shape_a = b->type == DISC ? b->radius : b->width;
    o->shape_b = b->type == DISC ? 0.0f : b->height;
    return 0;
}
#define ORC_SETTER(name, stmt)                 \
    int name {                                 \
        RigidBody* b = ((Solver*)h)->find(id); \
        if (!b) return -3;                     \
        stmt;                                  \
        return 0;                              \
    }
ORC_SETTER(orc_body_set_static(void* h, uint32_t id, int v), b->is_static = (v != 0))
ORC_SETTER(orc_body_set_pos(void* h, uint32_t id, float x, float y), b->props.pos = Vector2(x, y))
ORC_SETTER(orc_body_set_angle(void* h, uint32_t id, float a), b->props.angle = a)
ORC_SETTER(orc_body_set_momentum(void* h, uint32_t id, float x, float y), b->props.momentum = Vector2(x, y))
ORC_SETTER(orc_body_set_ang_momentum(void* h, uint32_t id, float l), b->props.ang_momentum = l)
ORC_SETTER(orc_body_set_force(void* h, uint32_t id, float x, float y), b->props.force = Vector2(x, y))
ORC_SETTER(orc_body_set_torque(void* h, uint32_t id, float t), b->props.torque = t)

int orc_read_bodies(void* h, uint32_t* ids, float* pos_xy, float* angle, float* momentum_xy, float* ang_momentum, float* aabb_xywh, size_t capacity) {
    Solver* s = (Solver*)h;
    if (capacity < s->bodies.size()) return -4;
    for (size_t i = 0; i < s->bodies.size(); ++i) {
        const RigidBody& b = s->bodies[i];
        if (ids) ids[i] = b.id;
        if (pos_xy) {
            pos_xy[2 * i] = b.props.pos.x;
            pos_xy[2 * i + 1] = b.props.pos.y;
        }
        if (angle) angle[i] = b.props.angle;
        if (momentum_xy) {
            momentum_xy[2 * i] = b.props.momentum.x;
            momentum_xy[2 * i + 1] = b.props.momentum.y;
        }
        if (ang_momentum) ang_momentum[i] = b.props.ang_momentum;
        if (aabb_xywh) {
            aabb_xywh[4 * i] = b.aabb.pos.x;
            aabb_xywh[4 * i + 1] = b.aabb.pos.y;
            aabb_xywh[4 * i + 2] = b.aabb.half_width;
            aabb_xywh[4 * i + 3] = b.aabb.half_height;
        }
    }
    return 0;
}
int orc_write_forces(void* h, const float* fxy_t, size_t n) {
    Solver* s = (Solver*)h;
    if (n > s->bodies.size()) return -4;
    for (size_t i = 0; i < n; ++i) {
        s->bodies[i].props.force = Vector2(fxy_t[3 * i], fxy_t[3 * i + 1]);
        s->bodies[i].props.torque = fxy_t[3 * i + 2];
    }
    return 0;
}
// Test helper: overwrite the dynamic state of all bodies (iteration order) — lets a full-size scene that was settled on
// the GPU be handed to the oracle for a bit-exact comparison of the following steps.
int orc_load_state(void* h, size_t n, const float* pos_xy, const float* angle, const float* momentum_xy, const float* ang_momentum,
                   const float* aabb_xywh) {
    Solver* s = (Solver*)h;
    if (n != s->bodies.size()) return -4;
    for (size_t i = 0; i < n; ++i) {
        RigidBody& b = s->bodies[i];
        b.props.pos = Vector2(pos_xy[2 * i], pos_xy[2 * i + 1]);
        b.props.angle = angle[i];
        b.props.momentum = Vector2(momentum_xy[2 * i], momentum_xy[2 * i + 1]);
        b.props.ang_momentum = ang_momentum[i];
        b.aabb.pos = Vector2(aabb_xywh[4 * i], aabb_xywh[4 * i + 1]);
        b.aabb.half_width = aabb_xywh[4 * i + 2];
        b.aabb.half_height = aabb_xywh[4 * i + 3];
    }
    return 0;
}
int orc_read_pairs(void* h, uint32_t* lo, uint32_t* hi, size_t capacity, size_t* out_n) {
    Solver* s = (Solver*)h;
    if (out_n) *out_n = s->candidate_pairs.size();
    if (!lo || !hi) return 0;
    if (capacity < s->candidate_pairs.size()) return -4;
    size_t i = 0;
    for (auto& p : s->candidate_pairs) {
        lo[i] = p.first;
        hi[i] = p.second;
        ++i;
    }
    return 0;
}
// Manifolds in discovery (insertion) order, as built by the last process() (depth = original depth).
int orc_read_manifolds(void* h, orc_manifold* out, size_t capacity, size_t* out_n) {
    Solver* s = (Solver*)h;
    if (out_n) *out_n = s->manifolds.size();
    if (!out) return 0;
    if (capacity < s->manifolds.size()) return -4;
    for (size_t i = 0; i < s->manifolds.size(); ++i) {
        const CollisionManifold& m = s->manifolds[i];
        orc_manifold& o = out[i];
        memset(&o, 0, sizeof(o));
        o.ref_id = s->bodies[m.ref_body].id;
        o.inc_id = s->bodies[m.inc_body].id;
        o.normal_id = (uint32_t)m.reference_normal_id;
        o.normal_x = m.normal.x;
        o.normal_y = m.normal.y;
        o.color = m.color;
        uint32_t n = 0;
        for (int k = 0; k < 2; ++k) {
            if (!m.points[k].present) continue;
            const CollisionPoint& p = m.points[k].p;
            o.pos_x[n] = p.pos.x;
            o.pos_y[n] = p.pos.y;
            o.depth[n] = p.original_depth;
            o.ref_rx[n] = p.ref_r.x;
            o.ref_ry[n] = p.ref_r.y;
            o.inc_rx[n] = p.inc_r.x;
            o.inc_ry[n] = p.inc_r.y;
            ++n;
        }
        o.n_points = n;
    }
    return 0;
}
int orc_read_joint_order(void* h, uint32_t* joint_index, uint32_t* joint_color, size_t capacity, size_t* out_n) {
    Solver* s = (Solver*)h;
    if (out_n) *out_n = s->joint_order.size();
    if (!joint_index) return 0;
    if (capacity < s->joint_order.size()) return -4;
    for (size_t i = 0; i < s->joint_order.size(); ++i) {
        joint_index[i] = (uint32_t)s->joint_order[i];
        if (joint_color) joint_color[i] = s->constraints[s->joint_order[i]].color;
    }
    return 0;
}
int orc_get_stats(void* h, orc_step_stats* out) {
    Solver* s = (Solver*)h;
    memset(out, 0, sizeof(*out));
    out->n_bodies = (uint32_t)s->bodies.size();
    out->n_buckets = (uint32_t)((s->use_init_grid_params ? s->spatialhash_table_size_mult : 2) * s->bodies.size());
    out->n_entries = (uint32_t)s->stat_entries;
    out->n_pairs = (uint32_t)s->candidate_pairs.size();
    out->n_manifolds = (uint32_t)s->manifolds.size();
    uint32_t k = 0;
    for (auto& m : s->manifolds) k += (m.points[0].present ? 1 : 0) + (m.points[1].present ? 1 : 0);
    out->n_points = k;
    out->n_colors = s->stat_colors;
    out->n_joints = (uint32_t)s->constraints.size();
    out->n_joint_colors = s->stat_joint_colors;
    out->n_dropped = s->stat_dropped;
    return 0;
}
uint64_t orc_raw_candidates(void* h) { return ((Solver*)h)->stat_raw_candidates; }
// Runs `steps` process() calls and returns wall seconds (steady clock) — the CPU-baseline timer of bench.py.
double orc_timed_steps(void* h, float dt, uint32_t sub_steps, uint32_t iters, uint32_t steps) {
    Solver* s = (Solver*)h;
    auto t0 = std::chrono::steady_clock::now();
    for (uint32_t i = 0; i < steps; ++i) s->process(dt, sub_steps, iters);
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}
// raw probes for known-answer tests
float orc_sinf(float x) { return zsinf(x); }
float orc_cosf(float x) { return zcosf(x); }
int orc_aabb_intersects(float ax, float ay, float ahw, float ahh, float bx, float by, float bhw, float bhh) {
    AABB a, b;
    a.pos = Vector2(ax, ay);
    a.half_width = ahw;
    a.half_height = ahh;
    b.pos = Vector2(bx, by);
    b.half_width = bhw;
    b.half_height = bhh;
    return a.intersects(b) ? 1 : 0;
}
uint64_t orc_cell_hash(uint64_t table_size, int64_t xi, int64_t yi) { return SpatialHash::hash(table_size, xi, yi); }

}  // extern "C"
