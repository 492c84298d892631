/* r2d_abi.h — C ABI of libr2d_b200.so: the drop-in boundary for resolve2d's per-step hot path
 * (`Solver.process`, /root/reference/src/core/lib.zig:189-251) on one NVIDIA B200 (sm_100a).
 *
 * The reference has no plugin ABI for this path; its two real surfaces are the Zig API
 * (`Solver` + `EntityFactory`, src/core/lib.zig:13-315) and the flat wasm C ABI over one global solver
 * (src/wasm_root.zig:18-251).  Every entry point below names the reference interface it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types.
 *   - every function returns an `int` status: R2D_OK (0) or a negative R2D_ERR_* (the Zig error union /
 *     the wasm shim's `bool` flattened, src/wasm_root.zig:39-43).
 *   - one opaque `r2d_solver*` per world.  A batch (`r2d_batch*`) owns many independent worlds that are
 *     stepped together on one GPU; `r2d_batch_world` hands out the per-world `r2d_solver*` used for scene
 *     construction and state access.
 *   - body ids are u32, assigned monotonically from 0 per world (reference: u16, src/core/Bodies/RigidBody.zig:15,
 *     lib.zig:66-71) — declared deviation, configs with > 65,535 bodies cannot exist in the reference.
 *   - there is NO CPU fallback: if no CUDA device is usable, r2d_create fails with R2D_ERR_NO_DEVICE.
 */
#ifndef R2D_ABI_H
#define R2D_ABI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define R2D_ABI_VERSION 2

/* ---- status codes ---------------------------------------------------------------------------- */
#define R2D_OK 0
#define R2D_ERR_OUT_OF_MEMORY (-1)     /* Zig error.OutOfMemory (host or device) */
#define R2D_ERR_INVALID_BODY_ID (-2)   /* Zig error.InvalidRigidBodyId: a joint names a removed body (DistanceJoint.zig:44-46) */
#define R2D_ERR_NO_SUCH_ID (-3)        /* Zig error.NoSuchIdExists (lib.zig:308-310) */
#define R2D_ERR_INVALID_ARGUMENT (-4)  /* NULL handle / out-of-range argument (reference: `unreachable`) */
#define R2D_ERR_NO_DEVICE (-5)         /* no usable CUDA device: the product path has no CPU fallback */
#define R2D_ERR_CUDA (-6)              /* CUDA runtime error; text via r2d_last_error() */
#define R2D_ERR_COLOR_OVERFLOW (-7)    /* the graph colouring did not terminate (internal limit of 4,000 rounds) */
#define R2D_ERR_BAD_STATE (-8)         /* e.g. r2d_process on a world that belongs to a multi-world batch */
#define R2D_ERR_GRID_RANGE (-9)        /* a body AABB covers an unreasonable number of grid cells (NaN/inf pose) */

/* HARD LIMIT of this boundary (the reference has none): the contacts of a body are swept in colour order and a body
 * can hold R2D_MAX_COLORS colours, so a non-static body with more than 256 simultaneous manifolds keeps its 256
 * highest-priority ones (contact_priority of the two ids); the others are reported by
 * r2d_read_manifolds with colour R2D_COLOR_DROPPED, counted in r2d_step_stats.n_dropped and left out of that call's
 * sweeps.  Everything else of the step — every other contact, joints, integration — proceeds normally. */
#define R2D_MAX_COLORS 256
#define R2D_COLOR_DROPPED 0xFFFFFFFDu

/* ---- shapes and joint kinds (reference enums: RigidBody.zig:28-31, Constraint.zig:15-20) ------ */
#define R2D_SHAPE_DISC 0
#define R2D_SHAPE_RECT 1

#define R2D_JOINT_DISTANCE 0
#define R2D_JOINT_OFFSET_DISTANCE 1
#define R2D_JOINT_FIXED_POSITION 2
#define R2D_JOINT_MOTOR 3

/* ---- broadphase mode ---------------------------------------------------------------------------
 * PARITY: cell = 4.0, table = 2*N — what the reference really does (lib.zig:254-255 ignores Solver.init's
 *         arguments).  Candidate-pair sets are bit-exact with the reference.
 * FAST:   honours cell_width / table_mult given to r2d_create (the reference's stated intent, README roadmap). */
#define R2D_MODE_PARITY 0
#define R2D_MODE_FAST 1
/* REFERENCE_ORDER (validation only, slow, single worlds): the PARITY broadphase and narrowphase, but the contacts are
 *         swept SEQUENTIALLY in the reference's own order — manifolds in the order updateManifolds discovers them
 *         (lib.zig:262-297 over the SpatialHash query order, replayed on the host), joints in list order — instead of the
 *         graph-colour order.  Gauss-Seidel is order dependent, so this is the mode whose body state equals the
 *         reference binary's bit for bit on every step (tests/golden/wasm_golden.json: 1,285 steps). */
#define R2D_MODE_REFERENCE_ORDER 2

typedef struct r2d_solver r2d_solver;
typedef struct r2d_batch r2d_batch;

/* EntityFactory.BodyOptions (lib.zig:35-46).  mass_is_density != 0 selects `.density`, else `.mass`. */
typedef struct r2d_body_opts {
    float pos_x, pos_y;
    float vel_x, vel_y;
    float angle;
    float omega;
    float mu;            /* reference default 0.5 */
    float mass_value;    /* density [kg/m^2] or mass [kg] */
    int32_t mass_is_density;
} r2d_body_opts;

/* One body for bulk creation: r2d_body_opts + geometry (DiscOptions / RectangleOptions, lib.zig:48-55). */
typedef struct r2d_body_desc {
    r2d_body_opts opts;
    int32_t shape;       /* R2D_SHAPE_* */
    float a, b;          /* disc: a = radius (b ignored); rect: a = width, b = height */
    int32_t is_static;   /* `body_unwrap().static = true` right after creation */
} r2d_body_desc;

/* Constraint.Parameters (Constraint.zig:27-31); defaults +inf, -inf, 10. */
typedef struct r2d_joint_params {
    float power_max;
    float power_min;
    float beta;
} r2d_joint_params;

/* Everything the reference exposes per body through `*RigidBody` (RigidBody.zig:69-91; the ~30 getters of
 * wasm_root.zig:88-251). */
typedef struct r2d_body_state {
    uint32_t id;
    int32_t shape;
    int32_t is_static;
    float pos_x, pos_y, angle;
    float momentum_x, momentum_y, ang_momentum;
    float force_x, force_y, torque;
    float mass, inertia, mu;
    float aabb_x, aabb_y, aabb_half_w, aabb_half_h;
    float shape_a, shape_b;
} r2d_body_state;

/* One contact manifold as produced by updateManifolds (lib.zig:287-294, collision.zig:22-69). */
typedef struct r2d_manifold {
    uint32_t ref_id, inc_id;
    uint32_t normal_id;        /* SATResult.reference_normal_id */
    uint32_t n_points;         /* 0..2 */
    float normal_x, normal_y;
    float pos_x[2], pos_y[2];  /* stored mid-depth point (CollisionPoint.pos) */
    float depth[2];
    float ref_rx[2], ref_ry[2];
    float inc_rx[2], inc_ry[2];
    uint32_t color;            /* graph colour used for the Gauss-Seidel sweep order (new; not in the reference) */
} r2d_manifold;

/* Counters of the last process() call (SURVEY §8 symbols). */
typedef struct r2d_step_stats {
    uint32_t n_bodies;     /* N */
    uint32_t n_buckets;    /* T */
    uint32_t n_entries;    /* E */
    uint32_t n_pairs;      /* P = |candidate set| */
    uint32_t n_manifolds;  /* M */
    uint32_t n_points;     /* K */
    uint32_t n_colors;     /* contact colours */
    uint32_t n_color_rounds;
    uint32_t n_joints;     /* C */
    uint32_t n_joint_colors;
    uint32_t n_launches;   /* kernels launched by the last process() */
    uint32_t n_dropped;    /* manifolds left out of the sweeps: no colour free (see R2D_MAX_COLORS); 0 in any sane scene */
} r2d_step_stats;

/* ---- library ---------------------------------------------------------------------------------- */
int r2d_abi_version(void);
const char* r2d_last_error(void);                 /* thread-local text of the last failure */
int r2d_device_count(int* out_count);

/* ---- Solver lifetime: Solver.init / deinit / clear (lib.zig:145-187; wasm solverInit/solverDeinit) ------ */
int r2d_create(float cell_width, uint32_t table_mult, int device, r2d_solver** out);
int r2d_destroy(r2d_solver* s);
int r2d_clear(r2d_solver* s);                     /* keeps exclusions and the id counter, like lib.zig:181-187 (Q16) */
int r2d_set_mode(r2d_solver* s, int mode);        /* R2D_MODE_*; default PARITY */
int r2d_set_stream(r2d_solver* s, void* cuda_stream); /* run on the caller's cudaStream_t (default: own stream) */
/* Options the reference lists on its roadmap but does not have (/root/reference/README.md:59-64).  They change results, so
 * they are OFF by default (parity with the reference is defined without them); the CPU oracle implements the same rules
 * and the CUDA path is bit-identical to it with the options on (tests/test_gpu_parity.py).
 *   R2D_OPT_WARM_START   "Persist manifolds across frames by keying with stable body IDs": every contact remembers, keyed by
 *                        (ref id, inc id), the impulses it had accumulated at the end of a call (per substep); a contact of
 *                        the next call with the same key, reference face and point count adds them to its FIRST update as
 *                        the initial guess of the iteration.  (ids < 2^24, worlds < 2^16.)
 *   R2D_OPT_SLEEPING     "Sleeping/awakening": a non-static body whose linear and angular speed stayed below 0.2 m/s and
 *                        0.2 rad/s for R2D_OPT_SLEEP_CALLS consecutive calls (default 30) and that has no user force or
 *                        torque IS A STATIC BODY for the duration of a call; a body that moved faster than the thresholds
 *                        in a call wakes every body it touched in that call (one hop per call).  A body named by a joint never sleeps. */
#define R2D_OPT_WARM_START 1
#define R2D_OPT_SLEEPING 2
#define R2D_OPT_SLEEP_CALLS 3
int r2d_set_option(r2d_solver* s, int option, uint32_t value);
/* Device memory keeps the bodies in a spatial (Morton) order so that bodies in contact are neighbours in HBM; ids,
 * iteration order and results are unaffected.  The order is re-derived from the current positions every `steps`
 * process() calls (default 1024, 0 = only when the scene is edited) or on demand — on the device while the state is resident
 * there (keys, radix sort, permutation: ~0.45 ms per 100k bodies), through the host (~6 ms) after an edit of the scene. */
int r2d_set_reorder_interval(r2d_solver* s, uint32_t steps);
int r2d_reorder(r2d_solver* s);

/* ---- EntityFactory (lib.zig:73-129) ------------------------------------------------------------------ */
int r2d_make_disc(r2d_solver* s, const r2d_body_opts* o, float radius, uint32_t* out_id);            /* makeDiscBody :73 */
int r2d_make_rect(r2d_solver* s, const r2d_body_opts* o, float width, float height, uint32_t* out_id); /* makeRectangleBody :84 */
int r2d_make_bodies(r2d_solver* s, const r2d_body_desc* descs, size_t n, uint32_t* out_first_id);    /* bulk form of the two above */
int r2d_make_gravity(r2d_solver* s, float g);                                                         /* makeDownwardsGravity :95 */
int r2d_make_distance_joint(r2d_solver* s, const r2d_joint_params* p, uint32_t id1, uint32_t id2,
                            float target_distance, size_t* out_index);                               /* :106 */
int r2d_make_offset_distance_joint(r2d_solver* s, const r2d_joint_params* p, uint32_t id1, uint32_t id2,
                                   float r1x, float r1y, float r2x, float r2y, float target_distance,
                                   size_t* out_index);                                               /* :100 */
int r2d_make_fixed_position_joint(r2d_solver* s, const r2d_joint_params* p, uint32_t id,
                                  float target_x, float target_y, size_t* out_index);                /* :112 */
int r2d_make_motor_joint(r2d_solver* s, const r2d_joint_params* p, uint32_t id, float target_omega,
                         size_t* out_index);                                                         /* :118 */
int r2d_exclude_pair(r2d_solver* s, uint32_t id1, uint32_t id2);                                      /* excludeCollisionPair :124 */
int r2d_remove_body(r2d_solver* s, uint32_t id);  /* removeRigidBody = swapRemove (lib.zig:308-310) */

/* ---- the hot path: Solver.process (lib.zig:189-251; wasm solverProcess) ------------------------------ */
int r2d_process(r2d_solver* s, float dt, uint32_t sub_steps, uint32_t collision_iters);
int r2d_step(r2d_solver* s, float dt, uint32_t sub_steps, uint32_t collision_iters); /* alias (north_star's name) */
/* process() followed by r2d_read_bodies() of ALL bodies (n = r2d_num_bodies) in one call and ONE host synchronisation: the
 * export of the new state is enqueued behind the step's last kernel (pinned destinations are written straight over PCIe).
 * Replaces the per-frame "solverProcess, then 17 getters per body" loop of demos/web/src/wasm_bridge.ts:43-81 /
 * demos/native/src/Renderer.zig:84-157.  Any output pointer may be NULL. */
int r2d_process_read(r2d_solver* s, float dt, uint32_t sub_steps, uint32_t collision_iters, uint32_t* ids, float* pos_xy,
                     float* angle, float* momentum_xy, float* ang_momentum, float* aabb_xywh, size_t n);
int r2d_synchronize(r2d_solver* s);               /* wait for the solver's stream */

/* ---- state access (replaces raw `*RigidBody`, lib.zig:23-31; wasm getters/setters :88-251) ------------ */
int r2d_num_bodies(r2d_solver* s, size_t* out);                              /* solverGetNumBodies */
int r2d_body_id_at(r2d_solver* s, size_t iter_index, uint32_t* out_id);      /* solverGetBodyIdBasedOnIter */
int r2d_body_get(r2d_solver* s, uint32_t id, r2d_body_state* out);
int r2d_body_set_static(r2d_solver* s, uint32_t id, int is_static);
int r2d_body_set_pos(r2d_solver* s, uint32_t id, float x, float y);
int r2d_body_set_angle(r2d_solver* s, uint32_t id, float angle);
int r2d_body_set_momentum(r2d_solver* s, uint32_t id, float mx, float my);
int r2d_body_set_ang_momentum(r2d_solver* s, uint32_t id, float l);
int r2d_body_set_force(r2d_solver* s, uint32_t id, float fx, float fy);
int r2d_body_set_torque(r2d_solver* s, uint32_t id, float torque);

/* Bulk SoA access in solver iteration order (slot i = i-th body of `bodies`).  Any output pointer may be NULL.
 * Buffers are HOST memory (pinned memory makes the copies asynchronous up to the final synchronize). */
int r2d_read_bodies(r2d_solver* s, uint32_t* ids, float* pos_xy, float* angle, float* momentum_xy,
                    float* ang_momentum, float* aabb_xywh, size_t capacity);
/* force_xy_torque: n x 3 floats {fx, fy, torque} per slot, written before the next process() (Q9: acts for one substep). */
int r2d_write_forces(r2d_solver* s, const float* force_xy_torque, size_t n);

/* Parity dumps of the last process(): the candidate set (sorted unique (lo_id, hi_id)) and the manifolds. */
int r2d_read_pairs(r2d_solver* s, uint32_t* lo_ids, uint32_t* hi_ids, size_t capacity, size_t* out_n);
int r2d_read_manifolds(r2d_solver* s, r2d_manifold* out, size_t capacity, size_t* out_n);
int r2d_read_joint_order(r2d_solver* s, uint32_t* joint_index, uint32_t* joint_color, size_t capacity, size_t* out_n);
int r2d_get_stats(r2d_solver* s, r2d_step_stats* out);

/* ---- batched independent worlds (north_star "Batched-worlds mode"; no reference counterpart) ---------- */
int r2d_batch_create(uint32_t n_worlds, float cell_width, uint32_t table_mult, int device, r2d_batch** out);
int r2d_batch_destroy(r2d_batch* b);
int r2d_batch_world(r2d_batch* b, uint32_t world, r2d_solver** out);   /* borrowed handle, owned by the batch */
int r2d_batch_num_worlds(r2d_batch* b, uint32_t* out);
int r2d_batch_set_mode(r2d_batch* b, int mode);
int r2d_batch_set_stream(r2d_batch* b, void* cuda_stream);
int r2d_batch_set_option(r2d_batch* b, int option, uint32_t value);   /* R2D_OPT_*, for every world of the batch */
int r2d_batch_set_reorder_interval(r2d_batch* b, uint32_t steps);
int r2d_batch_reorder(r2d_batch* b);
int r2d_batch_process(r2d_batch* b, float dt, uint32_t sub_steps, uint32_t collision_iters);
int r2d_batch_process_read(r2d_batch* b, float dt, uint32_t sub_steps, uint32_t collision_iters, uint32_t* ids, float* pos_xy,
                           float* angle, float* momentum_xy, float* ang_momentum, float* aabb_xywh, size_t n);
int r2d_batch_synchronize(r2d_batch* b);
int r2d_batch_num_bodies(r2d_batch* b, size_t* out);                   /* total over worlds */
/* bulk access over all worlds, world-major then slot order */
int r2d_batch_read_bodies(r2d_batch* b, uint32_t* ids, float* pos_xy, float* angle, float* momentum_xy,
                          float* ang_momentum, float* aabb_xywh, size_t capacity);
int r2d_batch_write_forces(r2d_batch* b, const float* force_xy_torque, size_t n);
int r2d_batch_get_stats(r2d_batch* b, r2d_step_stats* out);

/* ---- batched worlds sharded over several GPUs of one box (BASELINE config 5) -------------------
 * World w lives on shard floor(w * G / n_worlds) (contiguous blocks of worlds); shard k is an ordinary batch on
 * devices[k] with its own stream, resident there for the whole run.  One persistent host thread per shard issues its
 * work, so the shards step concurrently; there is NO collective and no peer traffic on the data path — worlds are
 * independent Solvers (every structure hangs off one `Solver`, lib.zig:132-142).  The same device may be named twice
 * (two shards on one GPU).  Bulk arrays are world-major over ALL worlds, i.e. the shards' arrays back to back. */
typedef struct r2d_sharded r2d_sharded;
int r2d_sharded_create(uint32_t n_worlds, const int* devices, uint32_t n_devices, float cell_width, uint32_t table_mult,
                       r2d_sharded** out);
int r2d_sharded_destroy(r2d_sharded* b);
int r2d_sharded_num_worlds(r2d_sharded* b, uint32_t* out);
int r2d_sharded_num_shards(r2d_sharded* b, uint32_t* out);
int r2d_sharded_shard(r2d_sharded* b, uint32_t shard, r2d_batch** out, uint32_t* first_world, uint32_t* n_worlds); /* borrowed */
int r2d_sharded_world(r2d_sharded* b, uint32_t world, r2d_solver** out);   /* borrowed handle (global world index) */
int r2d_sharded_set_mode(r2d_sharded* b, int mode);
int r2d_sharded_reorder(r2d_sharded* b);
int r2d_sharded_process(r2d_sharded* b, float dt, uint32_t sub_steps, uint32_t collision_iters);
int r2d_sharded_process_read(r2d_sharded* b, float dt, uint32_t sub_steps, uint32_t collision_iters, uint32_t* ids,
                             float* pos_xy, float* angle, float* momentum_xy, float* ang_momentum, float* aabb_xywh, size_t n);
int r2d_sharded_synchronize(r2d_sharded* b);
int r2d_sharded_num_bodies(r2d_sharded* b, size_t* out);
int r2d_sharded_read_bodies(r2d_sharded* b, uint32_t* ids, float* pos_xy, float* angle, float* momentum_xy,
                            float* ang_momentum, float* aabb_xywh, size_t capacity);
int r2d_sharded_write_forces(r2d_sharded* b, const float* force_xy_torque, size_t n);
int r2d_sharded_get_stats(r2d_sharded* b, r2d_step_stats* out);   /* counters summed over the shards */

/* ---- measurement hooks (bench.py's roofline leg) ----------------------------------------- */
#define R2D_KCLASS_BROADPHASE 0   /* pose/count, scan, fill, bucket sort, pair count/write */
#define R2D_KCLASS_NARROWPHASE 1
#define R2D_KCLASS_COLORING 2     /* colouring + colour partition + prestep */
#define R2D_KCLASS_INTEGRATE 3    /* forces+momentum (+AABB) and positions */
#define R2D_KCLASS_SOLVE_CONTACTS 4
#define R2D_KCLASS_SOLVE_JOINTS 5
#define R2D_KCLASS_COUNT 6
/* When enabled, every kernel launch of process() is bracketed by CUDA events on the solver's stream. */
int r2d_batch_profile_enable(r2d_batch* b, int enable);
int r2d_profile_enable(r2d_solver* s, int enable);
/* Accumulated since the last enable/reset: milliseconds and launch counts per kernel class (arrays of R2D_KCLASS_COUNT). */
int r2d_batch_profile_read(r2d_batch* b, double* ms, uint64_t* launches, int reset);
int r2d_profile_read(r2d_solver* s, double* ms, uint64_t* launches, int reset);

#ifdef __cplusplus
}
#endif
#endif /* R2D_ABI_H */
