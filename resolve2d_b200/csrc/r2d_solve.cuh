// r2d_solve.cuh — per-element arithmetic of the substep loop: force accumulation + momentum integration + AABB refresh,
// manifold pre-step, the sequential-impulse contact update, the four joints and position integration.  One call =
// one element (body / manifold / joint); the kernels in r2d_kernels.cu supply the indexing and the memory traffic.
//
// Follows  src/core/lib.zig:199-250, src/core/collision.zig:102-218, src/core/Constraints/{DistanceJoint.zig:40-75,
// OffsetDistanceJoint.zig:44-105,FixedPositionJoint.zig:38-72,MotorJoint.zig:38-64}, src/core/Forces/DownwardsGravity.zig:35-39,
// src/core/Bodies/Disc.zig:63-66, src/core/Bodies/Rectangle.zig:71-86.
#pragma once
#include "r2d_math.cuh"
#include "r2d_narrow.cuh"

namespace r2d {

// ---- bodies -------------------------------------------------------------------------------------------------

// updateAABB half extents: Disc.zig:63-66 (r, r); Rectangle.zig:71-86 (max over rotated local vertices, from 0 — Q25)
R2D_HD void aabb_half_extents(uint32_t flags, float shape_a, float shape_b, float c, float s, float& hw, float& hh) {
    if (flags & FLAG_RECT) {
        const float w = fdiv(shape_a, 2.0f), h = fdiv(shape_b, 2.0f);
        const v2 lv[4] = {mk2(-w, -h), mk2(-w, h), mk2(w, h), mk2(w, -h)};
        float width = 0.0f, height = 0.0f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const v2 rot = rotate_cs(lv[k], c, s);
            if (rot.x > width) width = rot.x;
            if (rot.y > height) height = rot.y;
        }
        hw = width;
        hh = height;
    } else {
        hw = shape_a;
        hh = shape_a;
    }
}

// ---- contacts -----------------------------------------------------------------------------------------------

// Per-manifold constants of one process() call.  preStep (collision.zig:102-133) runs every substep in the reference
// but its inputs never change inside a call (updateTGSDepth is inert, Q5), so it is evaluated once.
struct ContactConst {
    v2 normal, tangent;
    float friction;
    float inv_m1, inv_m2, inv_i1, inv_i2;
};
struct ContactPointConst {
    v2 r1, r2;
    float mass_n, mass_t;
    float depth;
    float bias;   // Baumgarte bias of collision.zig:172; depth and dt are constant during one process() call
};
// bias = BAUMGARTE * max(0, -depth - SLOP) / dt   (collision.zig:172)
R2D_HD float contact_bias(float depth, float dt) {
    return fdiv(fmul(BAUMGARTE, fmax_z(0.0f, fsub(-depth, BAUMGARTE_SLOP))), dt);
}

R2D_HD ContactConst prestep_manifold(v2 normal, bool static1, bool static2, float mass1, float mass2, float inertia1,
                                     float inertia2, float mu1, float mu2) {
    ContactConst c;
    c.inv_m1 = static1 ? 0.0f : fdiv(1.0f, mass1);
    c.inv_m2 = static2 ? 0.0f : fdiv(1.0f, mass2);
    c.inv_i1 = static1 ? 0.0f : fdiv(1.0f, inertia1);
    c.inv_i2 = static2 ? 0.0f : fdiv(1.0f, inertia2);
    c.normal = normal;
    c.tangent = rot90cw(normal);            // :113
    c.friction = fsqrt(fmul(mu1, mu2));     // :114
    return c;
}
R2D_HD void prestep_point(const ContactConst& c, ContactPointConst& p) {  // :116-132
    const float inv_mass = fadd(c.inv_m1, c.inv_m2);
    const float r1n = cross2(p.r1, c.normal);
    const float r2n = cross2(p.r2, c.normal);
    const float kn = fadd(fadd(inv_mass, fmul(c.inv_i1, fmul(r1n, r1n))), fmul(c.inv_i2, fmul(r2n, r2n)));
    const float r1t = cross2(p.r1, c.tangent);
    const float r2t = cross2(p.r2, c.tangent);
    const float kt = fadd(fadd(inv_mass, fmul(c.inv_i1, fmul(r1t, r1t))), fmul(c.inv_i2, fmul(r2t, r2t)));
    p.mass_n = (kn > 0.0f) ? fdiv(1.0f, kn) : 0.0f;
    p.mass_t = (kt > 0.0f) ? fdiv(1.0f, kt) : 0.0f;
}

struct BodyVel {
    v2 mom;
    float ang;
};

// One contact point of calculateImpulses (collision.zig:146-206, the body of its point loop) up to — not including — the
// accumulation into the bodies' scratch impulses: returns false when the point is skipped (:154-158 depth >= 0, which
// also zeroes its accumulated impulses; :176 impulse below MIN_MANIFOLD_IMPULSE, Q6), else true with `dp` = the impulse
// (applied negatively to body 1, positively to body 2) and `acc` updated.  Both points of a manifold see the same,
// pre-loop body velocities (Q8).
// `warm` (R2D_OPT_WARM_START, not in the reference): the previous call's per-substep impulse of this contact, added to the
// FIRST update of the call as the initial guess; `use_warm` is false everywhere else (and the additions are not executed).
R2D_HD bool contact_point_impulse(const ContactConst& c, const ContactPointConst& p, v2& acc, v2 vlinear_1, float omega1,
                                  v2 vlinear_2, float omega2, v2& dp, bool use_warm = false, v2 warm = v2{0.0f, 0.0f}) {
    if (p.depth >= 0.0f) {  // :154-158
        acc = mk2(0.0f, 0.0f);
        return false;
    }
    const v2 r1 = p.r1, r2 = p.r2;
    const v2 vrot_1 = mk2(fmul(-r1.y, omega1), fmul(r1.x, omega1));
    const v2 v1 = add2(vlinear_1, vrot_1);
    const v2 vrot_2 = mk2(fmul(-r2.y, omega2), fmul(r2.x, omega2));
    const v2 v2_ = add2(vlinear_2, vrot_2);
    const v2 dv = sub2(v1, v2_);

    float num = fadd(dot2(dv, c.normal), p.bias);  // :172-173, bias evaluated once per call (contact_bias)
    float pn = fmul(num, p.mass_n);
    if (use_warm) pn = fadd(pn, warm.x);
    if (pn < MIN_MANIFOLD_IMPULSE) return false;  // :176 (Q6)

    num = dot2(dv, c.tangent);
    float pt = fmul(num, p.mass_t);
    if (use_warm) pt = fadd(pt, warm.y);

    const float new_acc_pn = fmax_z(0.0f, fadd(acc.x, pn));
    const float applied_pn = fsub(new_acc_pn, acc.x);
    acc.x = new_acc_pn;

    const float max_pt = fmul(c.friction, fabs_z(acc.x));
    const float new_acc_pt = clamp_z(fadd(acc.y, pt), -max_pt, max_pt);
    const float applied_pt = fsub(new_acc_pt, acc.y);
    acc.y = new_acc_pt;

    const v2 pn_vec = scale2(c.normal, applied_pn);
    const v2 pt_vec = scale2(c.tangent, applied_pt);
    dp = add2(pn_vec, pt_vec);
    return true;
}

// calculateImpulses (collision.zig:135-218).  acc[k] = {accumulated_pn, accumulated_pt} persists for the whole
// process() call (Q7).  Both points see the pre-loop velocities and are applied once at the end (Q8).
R2D_HD void solve_contact(const ContactConst& c, int n_points, const ContactPointConst* pts, v2* acc, bool static1,
                          bool static2, BodyVel& b1, BodyVel& b2, bool use_warm = false, const v2* warm = nullptr) {
    const v2 vlinear_1 = scale2(b1.mom, c.inv_m1);
    const float omega1 = fmul(b1.ang, c.inv_i1);
    const v2 vlinear_2 = scale2(b2.mom, c.inv_m2);
    const float omega2 = fmul(b2.ang, c.inv_i2);

    v2 lin1 = mk2(0.0f, 0.0f), lin2 = mk2(0.0f, 0.0f);
    float rot1 = 0.0f, rot2 = 0.0f;

#pragma unroll
    for (int k = 0; k < 2; ++k) {  // fully unrolled: pts[] / acc[] stay in registers (no local-memory arrays)
        if (k >= n_points) break;
        v2 dp;
        if (!contact_point_impulse(c, pts[k], acc[k], vlinear_1, omega1, vlinear_2, omega2, dp, use_warm,
                                   use_warm ? warm[k] : v2{0.0f, 0.0f}))
            continue;
        if (!static1) {
            lin1 = sub2(lin1, dp);
            rot1 = fsub(rot1, cross2(pts[k].r1, dp));
        }
        if (!static2) {
            lin2 = add2(lin2, dp);
            rot2 = fadd(rot2, cross2(pts[k].r2, dp));
        }
    }
    b1.mom = add2(b1.mom, lin1);   // :208-212
    b1.ang = fadd(b1.ang, rot1);
    b2.mom = add2(b2.mom, lin2);
    b2.ang = fadd(b2.ang, rot2);
}

// ---- joints -------------------------------------------------------------------------------------------------

struct JointBody {
    v2 pos;
    float angle;
    v2 mom;
    float ang;
    float mass, inertia;
    float torque;
    bool is_static;
};

// DistanceJoint.solve (DistanceJoint.zig:40-75).  Writes momentum even into static bodies (Q10).
R2D_HD void solve_distance(JointBody& b1, JointBody& b2, float target, float beta, float power_min, float power_max) {
    const float epsilon = CONSTRAINT_GRADIENT_DIVISION_LIMIT;
    const float w1 = b1.is_static ? 0.0f : fdiv(1.0f, b1.mass);
    const float w2 = b2.is_static ? 0.0f : fdiv(1.0f, b2.mass);
    const v2 v1 = scale2(b1.mom, w1);
    const v2 v2_ = scale2(b2.mom, w2);
    const v2 relative_v = sub2(v2_, v1);
    const v2 normal = normalize2(sub2(b2.pos, b1.pos));
    const float dist_error = fsub(length2(sub2(b2.pos, b1.pos)), target);
    const float position_correction = fdiv(fmul(beta, dist_error), fmax_z(target, 0.001f));
    float J = fdiv(-fadd(dot2(relative_v, normal), position_correction), fadd(w1, w2));
    const float ndv = dot2(normal, relative_v);
    const float den = fdiv(1.0f, fmax_z(fabs_z(ndv), epsilon));
    const float J_max = fmul(power_max, den);
    const float J_min = fmul(power_min, den);
    J = clamp_z(J, J_min, J_max);
    const v2 dp = scale2(normal, J);
    b1.mom = sub2(b1.mom, dp);
    b2.mom = add2(b2.mom, dp);
}

// OffsetDistanceJoint.solve (OffsetDistanceJoint.zig:44-105)
R2D_HD void solve_offset_distance(JointBody& b1, JointBody& b2, v2 r1, v2 r2, float target, float beta,
                                  float power_min, float power_max) {
    const float epsilon = CONSTRAINT_GRADIENT_DIVISION_LIMIT;
    float c1, s1, c2, s2;
    sincos_ref(b1.angle, &s1, &c1);
    sincos_ref(b2.angle, &s2, &c2);
    const v2 a1 = add2(rotate_cs(r1, c1, s1), b1.pos);  // localToWorld
    const v2 a2 = add2(rotate_cs(r2, c2, s2), b2.pos);
    const v2 normal = normalize2(sub2(a2, a1));
    const float inv_m1 = b1.is_static ? 0.0f : fdiv(1.0f, b1.mass);
    const float inv_m2 = b2.is_static ? 0.0f : fdiv(1.0f, b2.mass);
    const float inv_i1 = b1.is_static ? 0.0f : fdiv(1.0f, b1.inertia);
    const float inv_i2 = b2.is_static ? 0.0f : fdiv(1.0f, b2.inertia);
    const v2 vlinear_1 = scale2(b1.mom, inv_m1);
    const float omega1 = fmul(b1.ang, inv_i1);
    const v2 vlinear_2 = scale2(b2.mom, inv_m2);
    const float omega2 = fmul(b2.ang, inv_i2);
    const v2 rotated_r1 = rotate_cs(r1, c1, s1);
    const v2 v_at_a1 = mk2(fmul(-rotated_r1.y, omega1), fmul(rotated_r1.x, omega1));
    const v2 v1 = add2(vlinear_1, v_at_a1);
    const v2 rotated_r2 = rotate_cs(r2, c2, s2);
    const v2 v_at_a2 = mk2(fmul(-rotated_r2.y, omega2), fmul(rotated_r2.x, omega2));
    const v2 v2_ = add2(vlinear_2, v_at_a2);
    const v2 dv = sub2(v2_, v1);
    const float dist_error = fsub(length2(sub2(a2, a1)), target);
    const float position_correction = fdiv(fmul(beta, dist_error), fmax_z(target, epsilon));
    const float num = fadd(dot2(normal, dv), position_correction);
    const float r1xn = cross2(rotated_r1, normal);
    const float r2xn = cross2(rotated_r2, normal);
    // inv_m1 + inv_m2 + r1xn*r1xn*inv_i1 + r2xn*r2xn*inv_i2, left to right
    const float den = fadd(fadd(fadd(inv_m1, inv_m2), fmul(fmul(r1xn, r1xn), inv_i1)), fmul(fmul(r2xn, r2xn), inv_i2));
    float J = fdiv(-num, den);
    const float ndv = dot2(normal, dv);
    const float den2 = fdiv(1.0f, fmax_z(fabs_z(ndv), epsilon));
    const float J_max = fmul(power_max, den2);
    const float J_min = fmul(power_min, den2);
    J = clamp_z(J, J_min, J_max);
    const v2 dp = scale2(normal, J);
    b1.mom = sub2(b1.mom, dp);
    b1.ang = fsub(b1.ang, fmul(J, r1xn));
    b2.mom = add2(b2.mom, dp);
    b2.ang = fadd(b2.ang, fmul(J, r2xn));
}

// FixedPositionJoint.solve (FixedPositionJoint.zig:38-72) — linear only (Q20)
R2D_HD void solve_fixed_position(JointBody& b, v2 target, float beta, float power_min, float power_max) {
    const float epsilon = ALLOWED_CONSTRAINT_VALUE;
    if (b.is_static) return;
    const float w = fdiv(1.0f, b.mass);
    const v2 v = scale2(b.mom, w);
    const v2 delta_pos = sub2(target, b.pos);
    const float dist = length2(delta_pos);
    if (dist < ALLOWED_CONSTRAINT_VALUE) return;
    const v2 normal = scale2(delta_pos, fdiv(1.0f, dist));
    const float bias = fmul(beta, dist);
    const float relative_velocity = dot2(v, normal);
    float J = fdiv(-fsub(relative_velocity, bias), w);
    const float den = fdiv(1.0f, fmax_z(fabs_z(relative_velocity), epsilon));
    const float J_max = fmul(power_max, den);
    const float J_min = fmul(power_min, den);
    J = clamp_z(J, J_min, J_max);
    const v2 dp = scale2(normal, J);
    b.mom = add2(b.mom, dp);
}

// MotorJoint.solve (MotorJoint.zig:38-64) — power limit uses torque/inertia (Q19); dt = sub_dt
R2D_HD void solve_motor(JointBody& b, float target_omega, float beta, float power_min, float power_max, float dt) {
    const float epsilon = ALLOWED_CONSTRAINT_VALUE;
    if (b.is_static) return;
    const float omega = fdiv(b.ang, b.inertia);
    const float C = fsub(omega, target_omega);
    if (fabs_z(C) < ALLOWED_CONSTRAINT_VALUE) return;
    const float bias = fmul(beta, C);
    float J = fsub(b.torque, bias);
    const float relative_velocity = fdiv(b.torque, b.inertia);
    const float den = fdiv(1.0f, fmax_z(fabs_z(relative_velocity), epsilon));
    const float J_max = fmul(power_max, den);
    const float J_min = fmul(power_min, den);
    J = clamp_z(J, J_min, J_max);
    b.ang = fadd(b.ang, fmul(J, dt));
}

// ---- graph colouring priority ---------------------------------------------------------------------------------
// Jones-Plassmann priority of a manifold: a pure function of the two (world-local) body ids, so the colouring — and
// with it the Gauss-Seidel order — does not depend on how pairs happen to be listed.  Unique among manifolds that
// share a body: the low word (lo+hi mod 2^32) differs whenever the other endpoint differs.
R2D_HD uint32_t mix32(uint32_t x) {
    x ^= x >> 16;
    x *= 0x7feb352du;
    x ^= x >> 15;
    x *= 0x846ca68bu;
    x ^= x >> 16;
    return x;
}
constexpr int PRIO_ROUND_SHIFT = 52;  // bits 52..63: round tag; bits 32..51: hash; bits 0..31: lo+hi
R2D_HD uint64_t contact_priority(uint32_t id_lo, uint32_t id_hi) {
    const uint32_t h = mix32(mix32(id_lo) ^ (id_hi * 0x9e3779b9u));
    return ((uint64_t)(h >> 12) << 32) | (uint64_t)(uint32_t)(id_lo + id_hi);
}

}  // namespace r2d
