/* r2d_wasm_compat.c — the reference's flat wasm C ABI (/root/reference/src/wasm_root.zig:18-251: 43 `export fn` over ONE
 * global solver) re-exported, name for name, over libr2d_b200.so.  A host that binds the reference's wasm module
 * (demos/web/src/wasm_bridge.ts:43-81, demos/native via the same getters) can bind this library instead and gets the
 * B200 path without a line of new glue: `solverInit`, `setup_0_*`, `solverProcess`, the per-body getters / setters.
 *
 * Plain C over the public C ABI (include/r2d_abi.h) — nothing here touches the device or knows about CUDA.
 *
 * Pointers.  The reference hands out `*RigidBody` as integers; here a body "pointer" is its id + 1 (never 0) and the
 * "implementation pointer" of getRigidBodyImplementation is the same handle.  The reference's `usize` is wasm32's
 * u32; native hosts get size_t.  Ids stay u16 in this surface, as in the reference (wasm_root.zig:45,55,60).
 *
 * Reads are served from one snapshot per solverProcess(): the first getter after a step pulls the state of all bodies
 * with one bulk r2d_read_bodies call instead of a device round trip per getter (the TS bridge issues 17 getters per
 * body and frame); setters go straight through r2d_body_set_* and refresh the snapshot entry.
 *
 * The example scenes the reference exports (wasm_root.zig:67-85) are restated call for call from
 * src/examples/0_1_car_platformer.zig, 0_3_many_boxes.zig and utils.zig:6-54 (0_2 and 0_4 are empty in the reference:
 * their bodies are commented out), all arithmetic in f32 exactly as Zig evaluates it.
 */
#include <math.h>
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/r2d_abi.h"

#define EXPORT __attribute__((visibility("default")))

static r2d_solver* g_solver = NULL;
static int g_device = 0;

/* ---- per-step snapshot -------------------------------------------------------------------------------------------- */
static r2d_body_state* g_snap = NULL;   /* indexed by iteration order */
static size_t g_snap_n = 0, g_snap_cap = 0;
static int32_t* g_slot_of_id = NULL;    /* id -> iteration index, -1 = no such body */
static size_t g_slot_cap = 0;
static bool g_snap_valid = false;

static void snapshot_drop(void) { g_snap_valid = false; }

static bool snapshot_build(void) {
    size_t n = 0;
    if (!g_solver || r2d_num_bodies(g_solver, &n) != R2D_OK) return false;
    if (n > g_snap_cap) {
        free(g_snap);
        g_snap_cap = n + n / 4 + 16;
        g_snap = (r2d_body_state*)malloc(g_snap_cap * sizeof(r2d_body_state));
        if (!g_snap) {
            g_snap_cap = 0;
            return false;
        }
    }
    uint32_t max_id = 0;
    for (size_t i = 0; i < n; ++i) {
        uint32_t id = 0;
        if (r2d_body_id_at(g_solver, i, &id) != R2D_OK) return false;
        /* r2d_body_get reads the host mirror, which the library refreshes from the device ONCE after a step */
        if (r2d_body_get(g_solver, id, &g_snap[i]) != R2D_OK) return false;
        if (id > max_id) max_id = id;
    }
    if ((size_t)max_id + 1 > g_slot_cap) {
        free(g_slot_of_id);
        g_slot_cap = (size_t)max_id + 1 + 64;
        g_slot_of_id = (int32_t*)malloc(g_slot_cap * sizeof(int32_t));
        if (!g_slot_of_id) {
            g_slot_cap = 0;
            return false;
        }
    }
    for (size_t k = 0; k < g_slot_cap; ++k) g_slot_of_id[k] = -1;
    for (size_t i = 0; i < n; ++i) g_slot_of_id[g_snap[i].id] = (int32_t)i;
    g_snap_n = n;
    g_snap_valid = true;
    return true;
}

/* the reference's getters are `unreachable` on a bad pointer (UB in ReleaseFast); here they read zeros */
static const r2d_body_state* body_of(size_t ptr) {
    static const r2d_body_state zero;
    if (!g_snap_valid && !snapshot_build()) return &zero;
    if (ptr == 0 || ptr - 1 >= g_slot_cap || g_slot_of_id[ptr - 1] < 0) return &zero;
    return &g_snap[g_slot_of_id[ptr - 1]];
}
static r2d_body_state* body_mut(size_t ptr) { return (r2d_body_state*)body_of(ptr); }

/* ---- Solver functions (wasm_root.zig:18-64) ----------------------------------------------------------------------- */
EXPORT void r2dCompatSetDevice(int device) { g_device = device; }   /* (new) which GPU solverInit uses; default 0 */

EXPORT bool solverInit(float spatialhash_cell_width, size_t spatialhash_table_size_mult) {
    if (g_solver) return true;                                                        /* :19 */
    if (r2d_create(spatialhash_cell_width, (uint32_t)spatialhash_table_size_mult, g_device, &g_solver) != R2D_OK) {
        g_solver = NULL;
        return false;
    }
    snapshot_drop();
    return true;
}
EXPORT void solverDeinit(void) {
    if (!g_solver) return;
    r2d_destroy(g_solver);
    g_solver = NULL;
    snapshot_drop();
}
EXPORT bool solverProcess(float dt, size_t sub_steps, size_t collision_iters) {
    if (!g_solver) return false;
    snapshot_drop();
    return r2d_process(g_solver, dt, (uint32_t)sub_steps, (uint32_t)collision_iters) == R2D_OK;
}
EXPORT size_t solverGetRigidbodyPtrById(uint16_t id) { return (size_t)id + 1; }
EXPORT size_t solverGetNumBodies(void) {
    size_t n = 0;
    if (g_solver) r2d_num_bodies(g_solver, &n);
    return n;
}
EXPORT uint16_t solverGetBodyIdBasedOnIter(size_t iter_idx) {
    uint32_t id = 0;
    if (g_solver) r2d_body_id_at(g_solver, iter_idx, &id);
    return (uint16_t)id;
}
EXPORT bool solverRemoveBodyById(uint16_t id) {
    if (!g_solver) return false;
    snapshot_drop();
    return r2d_remove_body(g_solver, id) == R2D_OK;
}

/* ---- RigidBody basic properties (:88-118) ------------------------------------------------------------------------- */
EXPORT size_t getRigidBodyPtrFromId(uint16_t id) { return (size_t)id + 1; }
EXPORT uint16_t getRigidBodyIdFromPtr(size_t ptr) { return (uint16_t)(ptr - 1); }
EXPORT bool isRigidBodyStatic(size_t ptr) { return body_of(ptr)->is_static != 0; }
EXPORT size_t getRigidBodyNumNormals(size_t ptr) { return body_of(ptr)->shape == R2D_SHAPE_RECT ? 4 : 1; }   /* Disc.zig:49, Rectangle.zig:58 */
EXPORT size_t getRigidBodyType(size_t ptr) { return (size_t)body_of(ptr)->shape; }   /* RigidBodies { disc, rectangle } */
EXPORT uint32_t getRigidBodyImplementation(size_t ptr) { return (uint32_t)ptr; }

/* ---- AABB (:120-138) ---------------------------------------------------------------------------------------------- */
EXPORT float getRigidBodyAABBPosX(size_t ptr) { return body_of(ptr)->aabb_x; }
EXPORT float getRigidBodyAABBPosY(size_t ptr) { return body_of(ptr)->aabb_y; }
EXPORT float getRigidBodyAABBHalfWidth(size_t ptr) { return body_of(ptr)->aabb_half_w; }
EXPORT float getRigidBodyAABBHalfHeight(size_t ptr) { return body_of(ptr)->aabb_half_h; }

/* ---- kinematic properties (:141-234) ------------------------------------------------------------------------------- */
EXPORT float getRigidBodyPosX(size_t ptr) { return body_of(ptr)->pos_x; }
EXPORT float getRigidBodyPosY(size_t ptr) { return body_of(ptr)->pos_y; }
EXPORT float getRigidBodyMomentumX(size_t ptr) { return body_of(ptr)->momentum_x; }
EXPORT float getRigidBodyMomentumY(size_t ptr) { return body_of(ptr)->momentum_y; }
EXPORT void setRigidBodyMomentumX(size_t ptr, float value) {
    r2d_body_state* b = body_mut(ptr);
    if (!g_solver || b->id + 1 != ptr) return;
    b->momentum_x = value;
    r2d_body_set_momentum(g_solver, b->id, b->momentum_x, b->momentum_y);
}
EXPORT void setRigidBodyMomentumY(size_t ptr, float value) {
    r2d_body_state* b = body_mut(ptr);
    if (!g_solver || b->id + 1 != ptr) return;
    b->momentum_y = value;
    r2d_body_set_momentum(g_solver, b->id, b->momentum_x, b->momentum_y);
}
EXPORT float getRigidBodyAngularVelocity(size_t ptr) {   /* :171-174: ang_momentum / inertia */
    const r2d_body_state* b = body_of(ptr);
    return b->ang_momentum / b->inertia;
}
EXPORT float getRigidBodyForceX(size_t ptr) { return body_of(ptr)->force_x; }
EXPORT float getRigidBodyForceY(size_t ptr) { return body_of(ptr)->force_y; }
EXPORT void setRigidBodyForceX(size_t ptr, float value) {
    r2d_body_state* b = body_mut(ptr);
    if (!g_solver || b->id + 1 != ptr) return;
    b->force_x = value;
    r2d_body_set_force(g_solver, b->id, b->force_x, b->force_y);
}
EXPORT void setRigidBodyForceY(size_t ptr, float value) {
    r2d_body_state* b = body_mut(ptr);
    if (!g_solver || b->id + 1 != ptr) return;
    b->force_y = value;
    r2d_body_set_force(g_solver, b->id, b->force_x, b->force_y);
}
EXPORT float getRigidBodyMass(size_t ptr) { return body_of(ptr)->mass; }
EXPORT float getRigidBodyAngle(size_t ptr) { return body_of(ptr)->angle; }
EXPORT float getRigidBodyAngularMomentum(size_t ptr) { return body_of(ptr)->ang_momentum; }
EXPORT void setRigidBodyAngularMomentum(size_t ptr, float value) {
    r2d_body_state* b = body_mut(ptr);
    if (!g_solver || b->id + 1 != ptr) return;
    b->ang_momentum = value;
    r2d_body_set_ang_momentum(g_solver, b->id, value);
}
EXPORT float getRigidBodyTorque(size_t ptr) { return body_of(ptr)->torque; }
EXPORT void setRigidBodyTorque(size_t ptr, float value) {
    r2d_body_state* b = body_mut(ptr);
    if (!g_solver || b->id + 1 != ptr) return;
    b->torque = value;
    r2d_body_set_torque(g_solver, b->id, value);
}
EXPORT float getRigidBodyInertia(size_t ptr) { return body_of(ptr)->inertia; }
EXPORT float getRigidBodyFrictionCoeff(size_t ptr) { return body_of(ptr)->mu; }

/* ---- shape specific (:237-251) -------------------------------------------------------------------------------------- */
EXPORT float getDiscBodyRadiusAssumeType(size_t implementation_ptr) { return body_of(implementation_ptr)->shape_a; }
EXPORT float getRectangleBodyWidthAssumeType(size_t implementation_ptr) { return body_of(implementation_ptr)->shape_a; }
EXPORT float getRectangleBodyHeightAssumeType(size_t implementation_ptr) { return body_of(implementation_ptr)->shape_b; }

/* ---- example scenes (wasm_root.zig:67-85) ---------------------------------------------------------------------------- */
typedef struct { float x, y; } vec2;
static vec2 v2(float x, float y) { vec2 v = {x, y}; return v; }
static float dist2(vec2 a, vec2 b) {   /* nmath.dist2: sqrt(dx*dx + dy*dy), every operation in f32 */
    const float dx = a.x - b.x, dy = a.y - b.y;
    return sqrtf(dx * dx + dy * dy);
}
static bool g_ok;
static uint32_t mk_rect(vec2 pos, float w, float h, float mu, float angle, bool is_static, bool density, float mass_value) {
    r2d_body_opts o;
    memset(&o, 0, sizeof o);
    o.pos_x = pos.x; o.pos_y = pos.y; o.angle = angle; o.mu = mu; o.mass_value = mass_value; o.mass_is_density = density ? 1 : 0;
    uint32_t id = 0;
    if (r2d_make_rect(g_solver, &o, w, h, &id) != R2D_OK) g_ok = false;
    if (is_static && r2d_body_set_static(g_solver, id, 1) != R2D_OK) g_ok = false;
    return id;
}
static uint32_t mk_disc(vec2 pos, float r, float mu, bool is_static, bool density, float mass_value) {
    r2d_body_opts o;
    memset(&o, 0, sizeof o);
    o.pos_x = pos.x; o.pos_y = pos.y; o.mu = mu; o.mass_value = mass_value; o.mass_is_density = density ? 1 : 0;
    uint32_t id = 0;
    if (r2d_make_disc(g_solver, &o, r, &id) != R2D_OK) g_ok = false;
    if (is_static && r2d_body_set_static(g_solver, id, 1) != R2D_OK) g_ok = false;
    return id;
}
static void offset_joint(const r2d_joint_params* p, uint32_t a, uint32_t b, vec2 r1, vec2 r2, float dist) {
    if (r2d_make_offset_distance_joint(g_solver, p, a, b, r1.x, r1.y, r2.x, r2.y, dist, NULL) != R2D_OK) g_ok = false;
}
static void distance_joint(const r2d_joint_params* p, uint32_t a, uint32_t b, float dist) {
    if (r2d_make_distance_joint(g_solver, p, a, b, dist, NULL) != R2D_OK) g_ok = false;
}
static void exclude(uint32_t a, uint32_t b) {
    if (r2d_exclude_pair(g_solver, a, b) != R2D_OK) g_ok = false;
}

/* src/examples/utils.zig:6-54 */
static void car(vec2 pos) {
    const r2d_joint_params dflt = {INFINITY, -INFINITY, 10.0f};   /* Constraint.Parameters defaults */
    float mu = 0.4f;
    const vec2 t_pos = v2(pos.x - 5.0f, pos.y - 10.0f);
    const uint32_t bh = mk_rect(pos, 5.0f, 0.9f, mu, 0.0f, false, true, 1.0f);
    const vec2 bh_pos = pos;
    mu = 1.5f;
    const float rad = 1.0f;
    const vec2 wl_pos = v2(t_pos.x + 3.5f, t_pos.y + (9.8f - rad));
    const uint32_t wl = mk_disc(wl_pos, rad, mu, false, true, 1.0f);
    const vec2 wr_pos = v2(t_pos.x + 6.5f, wl_pos.y);
    const uint32_t wr = mk_disc(wr_pos, rad, mu, false, true, 1.0f);
    mu = 0.5f;
    const float dist_car = dist2(bh_pos, wl_pos);
    const vec2 bh2_pos = v2(t_pos.x + 5.25f, t_pos.y + 10.8f);
    const uint32_t bh2 = mk_rect(bh2_pos, 1.2f, 0.5f, mu, 0.0f, false, true, 1.0f);

    const r2d_joint_params params = {2.0f, -2.0f, 14.0f};
    offset_joint(&params, wl, bh, v2(0, 0), v2(-1.5f, 0), rad + 0.2f);
    offset_joint(&params, wr, bh, v2(0, 0), v2(1.5f, 0), rad + 0.2f);
    distance_joint(&params, wl, wr, 3.0f);
    distance_joint(&params, wl, bh, dist_car);
    distance_joint(&params, wr, bh, dist_car);
    exclude(bh, wl);
    exclude(bh, wr);
    exclude(bh, bh2);
    const vec2 p1 = v2(bh_pos.x + 0.25f, bh_pos.y + 0.0f);
    const float dist21 = dist2(p1, v2(bh2_pos.x + -1.0f, bh2_pos.y + 1.0f));
    const float dist22 = dist2(p1, v2(bh2_pos.x + 1.0f, bh2_pos.y + 1.0f));
    offset_joint(&dflt, bh, bh2, v2(0.25f, 0), v2(-1, 1), dist21);
    offset_joint(&dflt, bh, bh2, v2(0.25f, 0), v2(1, 1), dist22);
    const float dist23 = dist2(bh_pos, bh2_pos);
    const r2d_joint_params stiff = {INFINITY, -INFINITY, 100.0f};
    distance_joint(&stiff, bh, bh2, dist23);
}

/* src/examples/0_1_car_platformer.zig:12-259 (N = 111, 11 joints, 3 exclusion pairs) */
EXPORT bool setup_0_1_car_platformer(void) {
    if (!g_solver) return false;
    g_ok = true;
    snapshot_drop();
    const r2d_joint_params dflt = {INFINITY, -INFINITY, 10.0f};
    if (r2d_make_gravity(g_solver, 9.82f) != R2D_OK) return false;
    mk_rect(v2(0, -100), 1000, 20, 0.3f, 0, true, true, 1);
    mk_rect(v2(0, 0), 40, 10, 0.3f, 0, true, true, 1);
    mk_rect(v2(-19, 15), 2, 20, 0.3f, 0, true, true, 1);
    car(v2(5, 10));
    const float mu = 0.5f;
    for (int x = -15; x <= -12; ++x)
        for (int y = 9; y <= 13; ++y) mk_rect(v2((float)x, (float)y), 0.6f, 0.4f, mu, 0, false, true, 1);
    mk_rect(v2(0, 5), 0.8f, 0.4f, mu, 1.0f, true, true, 1);
    mk_rect(v2(4.1f, 5), 0.8f, 0.7f, mu, 3.0f, true, true, 1);
    mk_rect(v2(1, 5), 0.8f, 0.4f, mu, 0.5f, true, true, 1);
    mk_rect(v2(15, 5), 0.3f, 0.4f, mu, 3.0f, true, true, 1);
    mk_rect(v2(-9, 5), 0.9f, 0.3f, mu, -1.0f, true, true, 1);
    mk_rect(v2(-10, 5), 0.8f, 0.6f, mu, -2.0f, true, true, 1);
    mk_rect(v2(20, 12), 15, 0.5f, mu, 0, true, true, 1);
    mk_rect(v2(7, 14), 14, 0.5f, mu, -0.3f, true, true, 1);
    mk_rect(v2(30, 7), 20, 1.0f, mu, 0.25f, true, true, 1);
    uint32_t body = mk_rect(v2(35, 13), 10.5f, 0.5f, mu, 0, false, true, 1);
    if (r2d_make_fixed_position_joint(g_solver, &dflt, body, 35, 13, NULL) != R2D_OK) g_ok = false;
    mk_disc(v2(38, 19), 1.0f, mu, false, false, 100);
    mk_rect(v2(48, 9.5f), 15, 1, mu, 0, true, true, 1);
    mk_rect(v2(52, 12), 3, 3, mu, 0, false, false, 10);
    mk_rect(v2(76.2f, 3.3f), 40, 1.0f, mu, -0.3f, true, true, 1);
    for (int xp = 65; xp <= 70; ++xp)
        for (int yp = 9; yp <= 19; ++yp) mk_rect(v2((float)xp, (float)yp), 0.9f, 0.4f, 0.7f, 0, false, true, 1);
    mk_rect(v2(110, -2), 35, 1.0f, mu, 0, true, true, 1);
    mk_disc(v2(110, -4), 4.0f, mu, true, true, 1);
    body = mk_rect(v2(100, 5), 12.9f, 0.5f, mu, 0, false, true, 1);
    if (r2d_make_fixed_position_joint(g_solver, &dflt, body, 100, 5, NULL) != R2D_OK) g_ok = false;
    const r2d_joint_params motor = {100.0f, -100.0f, 100.0f};
    if (r2d_make_motor_joint(g_solver, &motor, body, 3.14f, NULL) != R2D_OK) g_ok = false;
    mk_disc(v2(104, 8), 2.0f, mu, false, true, 1);
    return g_ok;
}

/* src/examples/0_2_bridge_stress.zig:12-79: the body of the reference's setup is commented out */
EXPORT bool setup_0_2_bridge_stress(void) { return g_solver != NULL; }

/* src/examples/0_3_many_boxes.zig:12-61 (N = 523, no joints) */
EXPORT bool setup_0_3_many_boxes(void) {
    if (!g_solver) return false;
    g_ok = true;
    snapshot_drop();
    if (r2d_make_gravity(g_solver, 9.82f) != R2D_OK) return false;
    const float mu = 0.3f;
    mk_rect(v2(20, -5), 1000, 10, mu, 0, true, true, 5);
    {
        r2d_body_opts o;
        memset(&o, 0, sizeof o);
        o.pos_x = -40; o.pos_y = 10; o.vel_x = 70; o.vel_y = 10; o.omega = -6; o.mu = mu; o.mass_value = 100; o.mass_is_density = 0;
        if (r2d_make_rect(g_solver, &o, 8.0f, 8.0f, NULL) != R2D_OK) g_ok = false;
    }
    mk_rect(v2(20, 5), 4.0f, 1.0f, mu, 0, false, true, 5);
    for (int x = 10; x < 30; ++x) {
        const float xf = 2.0f * (float)x;
        for (int y = 10; y < 36; ++y) {
            if (y % 2 == 0)
                mk_rect(v2(xf, (float)y), 1.0f, 1.0f, mu, 0, false, true, 5);
            else
                mk_disc(v2(xf, (float)y), 0.5f, mu, false, true, 5);
        }
    }
    return g_ok;
}

/* src/examples/0_4_also_many_boxes.zig:12-67: commented out in the reference as well */
EXPORT bool setup_0_4_also_many_boxes(void) { return g_solver != NULL; }
