// r2d_pipeline.cuh — device-state layout (structure-of-arrays in HBM) and the per-thread body of every kernel of
// the step pipeline.  Each `*_thread` function is what ONE CUDA thread does for ONE element; r2d_kernels.cu wraps
// them in grid-stride __global__ kernels.  They are host/device so that tests/emu/ can run the identical index and
// filter logic serially on a CPU against the oracle (test tooling only — the product has no CPU path).
//
// Reference mapping (src/core): broadphase = SpatialHash.zig:19-137 + the filters of lib.zig:273-282; narrowphase =
// collision.zig:221-363 (r2d_narrow.cuh); colouring is new (replaces the insertion-order sweep of lib.zig:226-236);
// solve/integrate = lib.zig:199-250, collision.zig:102-218, Constraints/*.zig (r2d_solve.cuh).
#pragma once
#include "r2d_math.cuh"
#include "r2d_narrow.cuh"
#include "r2d_solve.cuh"

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#else
// minimal vector types for the host-only test emulator
struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct uint2 { uint32_t x, y; };
struct int2 { int32_t x, y; };
static inline int2 make_int2(int32_t x, int32_t y) { int2 r; r.x = x; r.y = y; return r; }
struct alignas(16) int4 { int32_t x, y, z, w; };
static inline int4 make_int4(int32_t x, int32_t y, int32_t z, int32_t w) { int4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
struct alignas(16) uint4 { uint32_t x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { float4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
static inline uint2 make_uint2(uint32_t x, uint32_t y) { uint2 r; r.x = x; r.y = y; return r; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { uint4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
#endif

namespace r2d {

constexpr uint32_t COLOR_NONE = 0xFFFFFFFFu;       // pair slot holds no manifold
constexpr uint32_t COLOR_PENDING = 0xFFFFFFFEu;    // manifold not coloured yet
constexpr uint32_t COLOR_DROPPED = 0xFFFFFFFDu;    // manifold found none of the MAX_COLORS colours free on one of its bodies (a
                                                   // body with more than 256 simultaneous contacts): left out of this call's sweeps
constexpr uint32_t MAX_COLORS = 256;
constexpr uint32_t S_EMPTY = 0x400u;              // s_hdr.z: padding slot (colour segments are padded to whole warps)
constexpr uint32_t S_WARM = 0x800u;               // s_hdr.z: the record carries warm-start terms (s_warm0 / s_warm1)
constexpr uint32_t COLOR_ALIGN = 32;
constexpr uint32_t COLOR_WORDS = MAX_COLORS / 64;
constexpr uint32_t ADJ_CAP = 32;                   // manifolds a body lists in place for the dataflow colouring; further ones are chained
constexpr uint32_t FLOW_COLORS = 96;               // colours the 16-byte colouring word of a body can hold
constexpr uint32_t ADJ_MAX = FLOW_COLORS;          // a body with more manifolds needs more colours than that anyway: rounds
constexpr uint32_t MAX_COLOR_ROUNDS = 4000;        // < 2^12 (round tag field of the priority word)
constexpr uint32_t BIG_BODY_CELLS = 64;            // bodies covering more cells are walked by a whole CTA
constexpr int64_t MAX_BODY_CELLS = 1 << 22;        // beyond this the pose is garbage (NaN/inf): R2D_ERR_GRID_RANGE

// wait policy of the dataflow kernels (never affects results): lag <= WAIT_SPIN_LAG polls again at once, otherwise the
// warp sleeps lag * WAIT_SLEEP_UNIT ns (at most WAIT_SLEEP_MAX); the dataflow colouring sleeps FLOW_SLEEP_UNIT ns per
// manifold still ahead.  Values from the round-1 sweeps (profiles/tune_solver.py).
constexpr uint32_t WAIT_SPIN_LAG = 1u, WAIT_SLEEP_UNIT = 200u, WAIT_SLEEP_MAX = 4000u, FLOW_SLEEP_UNIT = 150u;

constexpr uint32_t ERR_COLOR_OVERFLOW = 1u;
constexpr uint32_t ERR_GRID_RANGE = 2u;
constexpr uint32_t ERR_ROUNDS = 4u;
constexpr uint32_t ERR_FLOW_STALL = 16u;       // dataflow colouring stalled (same: a bug, reported)
constexpr uint32_t ERR_FINE = 32u;             // a body flagged small covers more than 4 coarse cells (a bug, reported)
constexpr uint32_t ERR_STALL = 8u;             // dataflow sweep stalled (a bug, never data): reported instead of hanging

// Device-side counters of one process() call (one 128-byte block, copied to pinned host memory once per step).
struct Counters {
    uint32_t n_entries;      // E
    uint32_t n_pairs;        // P
    uint32_t n_manifolds;    // M
    uint32_t n_points;       // K
    uint32_t n_colors;
    uint32_t n_rounds;
    uint32_t err;
    uint32_t n_own_scan;     // n_colors * (W + 1): length of the owner-position scan
    uint32_t n_work;         // buckets holding 2..SMALL_BUCKET entries (a warp each in the pair kernels)
    uint32_t n_mid;          // SMALL_BUCKET+1..HEAVY_BUCKET entries (a warp or a CTA each, see medium_by_cta)
    uint32_t n_heavy;        // buckets holding more (a CTA each)
    uint32_t n_dropped;      // manifolds left without a colour (COLOR_DROPPED)
    uint32_t flow_abort;     // set by the narrowphase: a body has more than ADJ_MAX manifolds (or the chain pool is full), colour by rounds
    uint32_t flow_fail;      // set inside the dataflow colouring: more than FLOW_COLORS colours needed (or a stall)
    uint32_t flow_used;      // the colours of this step come from the dataflow colouring (masks in cstate, not in used)
    uint32_t tile_fallback;  // k_solve_tiles declined (a tile has too many bodies / tasks): the host runs k_solve_persistent
    uint32_t max_world_m;    // k_world_solve: most slots (contact points) found in one world of the batch (the host sizes the
    uint32_t n_big;          // bodies covering more than BIG_BODY_CELLS cells, listed by the count kernel (big_bodies)
    uint32_t n_adj_over;     // entries taken from adj_pool (manifolds beyond ADJ_CAP of a body)
    uint32_t broad_fallback; // shared-memory cache of the next call from it) | k_world_broad: a world's grid does not fit
};

// Everything the kernels need, passed by value.
struct Dev {
    // ---- bodies (NB slots; world-major, slot = insertion index inside its world) ------------------------------
    uint32_t n_bodies;
    float4* pos;      // x, y, angle, -
    float4* mom;      // momentum.x, momentum.y, ang_momentum, bits(version): contact updates applied in this substep
    float4* frc;      // force.x, force.y, torque, -
    float4* prop;     // mass, inertia, mu, -
    float4* shape;    // a, b, bits(flags), bits(id)   flags: bit0 static, bit1 rect, bits 8.. world
    float4* aabb;     // centre.x, centre.y, half_w, half_h   (as of the last refresh — stale on purpose, Q3)
    float4* pose;     // x, y, cos(angle), sin(angle)  (scratch of one process())
    float4* view;     // 4 per body, rectangles only: world vertices 0..3, outward edge normals 0..3 (scratch of one process())
    uint32_t* ncells; // grid cells covered by the body's AABB
    uint4* bkt;       // the body's bucket ids when it covers <= 4 cells (0xFFFFFFFF padded); scratch of one process()
    // ---- worlds ----------------------------------------------------------------------------------------------------
    uint32_t n_worlds;
    const uint32_t* world_base;   // n_worlds + 1
    const uint64_t* world_magic;  // n_worlds: hash_magic(table_mult * bodies of the world) (cell_hash_magic)
    const uint32_t* grav_off;     // n_worlds + 1
    const float* grav;            // g values, world-major
    float cell;                   // grid cell size
    uint32_t table_mult;          // buckets per body
    // ---- hashed grid -----------------------------------------------------------------------------------------------
    uint32_t n_buckets;           // T = table_mult * NB (world w owns [mult*base_w, mult*base_{w+1}))
    uint32_t* bucket_cnt;         // T + 1, all zero between kernels pairs count/fill
    uint32_t* bucket_start;       // T + 1 (exclusive scan of counts; [T] = E)
    uint32_t cap_entries;
    uint32_t* ent_body;           // E
    uint32_t* ent_key;            // E
    uint32_t* ent_off;            // T + 1: pairs emitted per BUCKET, then its exclusive scan ([T] = P)
    uint32_t* work;               // 2T: ids of the small buckets from the front of [0, T), of the heavy ones from its back
                                  // (work[T - 1 - k]), of the medium ones in [T, 2T); order irrelevant
    uint32_t* big_bodies;         // BIG_GLOBAL_LIST: bodies whose cells the fill kernel spreads over the whole grid
    uint32_t* hit_bits;           // 4 * cap_entries: 32-test ballots of the count pass, bucket b's words at 4 * bucket_start[b]
    const uint64_t* excl;         // sorted (lo_slot << 32 | hi_slot)
    uint32_t n_excl;
    // ---- fine grid (see "fine grid" below): small bodies list ONE home cell of width 1 / fine_inv in the buckets
    //      [n_buckets, 2 n_buckets); only FLAG_LARGE bodies are entered in the coarse buckets --------------------------
    uint32_t fine_on;             // 0: every body goes through the coarse buckets (the original pipeline)
    uint32_t ll_on;               // 1: the bucket pair kernels ran on the coarse (large-body) buckets, ent_off[T] = their pairs
    double fine_inv;              // 1 / fine cell width
    int4* fcell;                  // NB: home cell (x, y) of a small body, its fine bucket and its arrival rank in it (scratch of one process())
    float4* ent_aabb;             // E: the stored AABB of the body of a FINE entry (its ent_body carries the static flag in bit 31)
    uint4* fine_cand;             // 2 NB: the first 8 partners found by small body a (count pass -> write pass)
    uint32_t* pair_cnt;           // NB + 2: [0] = pairs of the bucket kernels, [a + 1] = pairs emitted by small body a; then
                                  // its exclusive scan: [a + 1] = first pair slot of body a, [NB + 1] = P
    // ---- candidate pairs / raw manifolds (P slots) ------------------------------------------------------------------
    uint32_t cap_pairs;
    uint2* pairs;                 // (owner slot, other slot)
    uint4* m_hdr;                 // ref slot, inc slot, n_points | normal_id << 8, dynamic mask (bit 0 ref, bit 1 inc)
    unsigned long long* m_prio;   // colouring priority of the manifold (contact_priority of the two ids)
    float4* m_g0;                 // normal.x, normal.y, p0.pos.x, p0.pos.y
    float4* m_g1;                 // p0.depth, p1.depth, p1.pos.x, p1.pos.y
    float4* m_r0;                 // p0: ref_r.x, ref_r.y, inc_r.x, inc_r.y
    float4* m_r1;                 // p1
    uint32_t* m_color;            // COLOR_NONE / COLOR_PENDING / colour
    // ---- colouring ---------------------------------------------------------------------------------------------------
    unsigned long long* maxprio0; // NB: highest pending priority seen on the body, even rounds
    unsigned long long* maxprio1; // NB: odd rounds
    unsigned long long* used;     // NB * COLOR_WORDS: colours taken on the body
    uint32_t* color_count;        // MAX_COLORS
    uint32_t* color_start;        // MAX_COLORS + 1
    uint32_t* color_cursor;       // MAX_COLORS
    uint32_t* round_left;         // MAX_COLOR_ROUNDS
    uint32_t* pend_cnt;           // MAX_COLOR_ROUNDS + 1: length of the pending list of each round (k_color, streamed manifolds)
    uint32_t* pend_list;          // 2 x cap_pairs: pair slots still pending, two lists that swap every round
    // dataflow colouring (single worlds): the manifolds of a body, and one 16-byte word per body that is both the wait
    // target and the data — {96-bit colour mask, manifolds coloured so far}
    uint32_t flow;                // 1: k_narrow lists every manifold on its non-static bodies (adj_*), k_color may use them
    uint32_t* adj_cnt;            // NB: manifolds on the body (more than ADJ_MAX: the colouring falls back to rounds)
    unsigned long long* adj_prio; // NB * ADJ_CAP: their priorities, in arrival order
    uint32_t* adj_head;           // NB: 1 + index of the body's first chained entry in adj_pool (0 = none); zeroed per step
    uint4* adj_pool;              // {priority lo, priority hi, 1 + next entry, -}: manifolds ADJ_CAP.. of a body (hub bodies)
    uint32_t adj_pool_cap;
    uint4* cstate;                // NB
    uint32_t own_words;           // W = ceil(NB / 32)
    uint32_t* own_bits;           // MAX_COLORS x W: bit (c, b) set iff body b owns a manifold of colour c (at most one)
    uint32_t* own_pos;            // MAX_COLORS x (W + 1) + 1: popcounts per word (+ the warp padding of the colour), then
                                  // their exclusive scan = slot of the first owner of each word in the colour-sorted array
    Counters* counters;
    // ---- solver records, grouped by colour (M slots) -----------------------------------------------------------------
    uint4* s_hdr;                 // ref slot, inc slot, n_points | static1 << 8 | static2 << 9 | S_EMPTY, pair slot
    float4* s_nf;                 // normal.x, normal.y, friction, -
    float4* s_inv;                // inv_m1, inv_m2, inv_i1, inv_i2
    float4* s_r0;                 // point 0: r1.x, r1.y, r2.x, r2.y
    float4* s_r1;
    float4* s_pm0;                // point 0: mass_n, mass_t, depth, Baumgarte bias (collision.zig:172; constant per call)
    float4* s_pm1;
    float2* s_acc0;               // point 0: accumulated_pn, accumulated_pt
    float2* s_acc1;
    uint32_t world_fused;         // 1: k_world_solve follows (batches of small worlds): it places and pre-steps the manifolds of
                                  // a world itself, so the colouring sets no owner bits and no partition kernel runs
    uint32_t tile_bodies;         // B > 0: k_solve_tiles will run with tiles of B consecutive body slots (else 0)
    uint32_t* body_shared;        // NB: 1 if a manifold owned by a body of ANOTHER tile touches the body (see k_solve_tiles)
    uint4* s_dep;                 // rank of this manifold among the contacts of its ref body, that body's contact count,
                                  // same for the inc body  (dataflow ordering of the sweep, see solve_contact_thread)
    // ---- roadmap options (README.md:59-64; off by default) ----------------------------------------
    // warm starting: open-addressing hash table of the previous call's contacts, keyed by stable ids
    uint32_t warm_on;
    uint32_t warm_mask;            // table size - 1 (a power of two)
    unsigned long long* warm_key;  // world << 48 | ref id << 24 | inc id; WARM_EMPTY = free
    uint32_t* warm_meta;           // n_points | normal_id << 8
    float4* warm_val;              // accumulated (pn, pt) of point 0 and point 1 at the end of the call, per substep
    float2* s_warm0;               // per record: warm terms of point 0 / point 1 (valid where S_WARM)
    float2* s_warm1;
    // sleeping: a slow body is a static body for the duration of a call
    uint32_t* sleep_cnt;           // NB: consecutive calls below the speed thresholds
    uint32_t* sleep_state;         // NB: bit 0 asleep in this call, bit 1 moved fast in this call
    // ---- joints, grouped by colour -----------------------------------------------------------------------------------
    uint32_t n_joints;
    const uint4* j_hdr;           // type, slot1, slot2, joint colour
    uint32_t n_joint_colors;
    const uint32_t* world_joint_start;  // NW + 1: the joints of world w are world_joint[start[w] .. start[w + 1]), in sweep order
    const uint32_t* world_joint;        // indices into j_hdr / j_par / j_vec (k_world_solve: joints of a world solved by its CTA)
    const float4* j_par;          // power_max, power_min, beta, target (distance | omega)
    const float4* j_vec;          // r1.x, r1.y, r2.x, r2.y  |  target.x, target.y, -, -
    // joints inside the dataflow sweep: the update sequence of a body in one iteration is [its joints in sweep order
    // (colour, list index)] then [its contacts in colour order]; versions count both
    uint32_t joints_flow;         // 1: k_solve_persistent sweeps the joints through the version words too (no grid barrier)
    const uint32_t* body_nj;      // NB: joints naming the body (null / unused when joints_flow == 0)
    const uint4* j_dep;           // per joint: rank among the joints of body 1, that body's joint count, same for body 2
    float sub_dt;                 // dt / sub_steps of the current process() call
    uint32_t color_smem;          // 1: maxprio / used point into shared memory (per-world colouring kernel)
};

R2D_HD uint32_t body_flags(const Dev& d, uint32_t i) { return f2u(d.shape[i].z); }
R2D_HD uint32_t body_id(const Dev& d, uint32_t i) { return f2u(d.shape[i].w); }

// ---- spatial order of the device slots (host: build_image; device: k_resort_*) ---------------------------------------------
// Sort key of a body inside its world: Morton code of the position on a 2 m lattice whose origin is the minimum over the
// world's bodies (NaN coordinates sort first).  One definition, evaluated by the host at upload and by the device at a
// re-sort, so that both derive the same order.
R2D_HD uint32_t morton16(uint32_t x, uint32_t y) {
    uint32_t v[2] = {x & 0xFFFFu, y & 0xFFFFu};
    for (int k = 0; k < 2; ++k) {
        v[k] = (v[k] | (v[k] << 8)) & 0x00FF00FFu;
        v[k] = (v[k] | (v[k] << 4)) & 0x0F0F0F0Fu;
        v[k] = (v[k] | (v[k] << 2)) & 0x33333333u;
        v[k] = (v[k] | (v[k] << 1)) & 0x55555555u;
    }
    return v[0] | (v[1] << 1);
}
R2D_HD uint32_t resort_key(float pos_x, float pos_y, float min_x, float min_y) {
    float fx = fmul(fsub(pos_x, min_x), 0.5f), fy = fmul(fsub(pos_y, min_y), 0.5f);
    if (!(fx >= 0.0f)) fx = 0.0f;
    if (!(fy >= 0.0f)) fy = 0.0f;
    if (fx > 65535.0f) fx = 65535.0f;
    if (fy > 65535.0f) fy = 65535.0f;
    return morton16((uint32_t)fx, (uint32_t)fy);
}
// order-preserving map float -> int for the per-world minimum (atomicMin on ints); NaNs are never passed in
R2D_HD int resort_float_order(float f) {
    const int i = (int)f2u(f);
    return i >= 0 ? i : (int)((uint32_t)i ^ 0x7FFFFFFFu);
}
R2D_HD float resort_float_unorder(int k) { return u2f((uint32_t)(k >= 0 ? k : (int)((uint32_t)k ^ 0x7FFFFFFFu))); }
constexpr int RESORT_NO_MIN = 0x7FFFFFFF;   // (the image of a NaN pattern: never produced by a finite or infinite coordinate)

// ---- atomics: real ones on the device, plain read-modify-write in the serial host emulator ---------------------------
R2D_HD uint32_t atomic_add_u32(uint32_t* p, uint32_t v) {
#if defined(__CUDA_ARCH__)
    return atomicAdd(p, v);
#else
    const uint32_t o = *p;
    *p = o + v;
    return o;
#endif
}
R2D_HD uint32_t atomic_sub_u32(uint32_t* p, uint32_t v) {
#if defined(__CUDA_ARCH__)
    return atomicSub(p, v);
#else
    const uint32_t o = *p;
    *p = o - v;
    return o;
#endif
}
R2D_HD void atomic_max_u32(uint32_t* p, uint32_t v) {
#if defined(__CUDA_ARCH__)
    atomicMax(p, v);
#else
    if (*p < v) *p = v;
#endif
}
R2D_HD uint32_t atomic_exch_u32(uint32_t* p, uint32_t v) {
#if defined(__CUDA_ARCH__)
    return atomicExch(p, v);
#else
    const uint32_t o = *p;
    *p = v;
    return o;
#endif
}
R2D_HD void atomic_or_u32(uint32_t* p, uint32_t v) {
#if defined(__CUDA_ARCH__)
    atomicOr(p, v);
#else
    *p |= v;
#endif
}
R2D_HD void atomic_max_u64(unsigned long long* p, unsigned long long v) {
#if defined(__CUDA_ARCH__)
    atomicMax(p, v);
#else
    if (*p < v) *p = v;
#endif
}

// index of the lowest ZERO bit (the word must have one)
R2D_HD uint32_t first_zero64(unsigned long long v) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__ffsll((long long)~v) - 1u;
#else
    return (uint32_t)__builtin_ctzll(~v);
#endif
}
R2D_HD uint32_t first_zero32(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__ffs((int)~v) - 1u;
#else
    return (uint32_t)__builtin_ctz(~v);
#endif
}
R2D_HD uint32_t popc64(unsigned long long v) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__popcll(v);
#else
    return (uint32_t)__builtin_popcountll(v);
#endif
}
// One 16-byte word per body holds {momentum.x, momentum.y, ang_momentum, version}.  The scalar .b128 relaxed accesses
// (LDG/STG.E.128.STRONG.GPU) are single-copy atomic at device scope, so whoever observes a version also observes the
// momentum written with it — no fence, no separate flag.  Plain accesses in the serial emulator.
R2D_HD float4 ld_body_word(const float4* p) {
#if defined(__CUDA_ARCH__)
    float4 v;
    asm volatile("{\n\t.reg .b128 t;\n\tld.relaxed.gpu.global.b128 t, [%4];\n\tmov.b128 {%0, %1, %2, %3}, t;\n\t}"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p)
                 : "memory");
    return v;
#else
    return *p;
#endif
}
R2D_HD void st_body_word(float4* p, float4 v) {
#if defined(__CUDA_ARCH__)
    asm volatile("{\n\t.reg .b128 t;\n\tmov.b128 t, {%1, %2, %3, %4};\n\tst.relaxed.gpu.global.b128 [%0], t;\n\t}" ::"l"(p),
                 "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
#else
    *p = v;
#endif
}
R2D_HD void backoff_ns(uint32_t ns) {
#if defined(__CUDA_ARCH__)
    __nanosleep(ns);
#else
    (void)ns;
#endif
}

// Loads of words that OTHER CTAs update inside the same cooperative kernel (between grid barriers): served from L2.
R2D_HD unsigned long long ld_shared_u64(const unsigned long long* p, bool in_smem = false) {
#if defined(__CUDA_ARCH__)
    if (in_smem) return *((const volatile unsigned long long*)p);  // per-world colouring: the words live in shared memory
    return __ldcg(p);
#else
    (void)in_smem;
    return *p;
#endif
}

// ---- broadphase ----------------------------------------------------------------------------------------------------
struct CellRange {
    int64_t min_xi, min_yi, max_xi, max_yi;
    uint32_t bucket_base;   // first bucket of the body's world
    uint32_t table_size;    // buckets of the body's world (T_w = table_mult * N_w)
    uint64_t magic;         // hash_magic(table_size)
    uint32_t nx;            // cells per row
    uint32_t count;         // total cells (0 if out of range)
};

// iterateAABBHashes (SpatialHash.zig:83-106): cell range of the STORED aabb
R2D_HD CellRange cell_range(const Dev& d, uint32_t i) {
    const float4 a = d.aabb[i];
    CellRange r;
    // getVertices (aabb.zig:25-32): min = pos - half, max = pos + half
    r.min_xi = cell_coord(fsub(a.x, a.z), d.cell);
    r.min_yi = cell_coord(fsub(a.y, a.w), d.cell);
    r.max_xi = cell_coord(fadd(a.x, a.z), d.cell);
    r.max_yi = cell_coord(fadd(a.y, a.w), d.cell);
    const uint32_t w = body_flags(d, i) >> FLAG_WORLD_SHIFT;
    const uint32_t b0 = d.world_base[w], b1 = d.world_base[w + 1];
    r.bucket_base = d.table_mult * b0;
    r.table_size = d.table_mult * (b1 - b0);
    r.magic = d.world_magic[w];
    const int64_t nx = r.max_xi - r.min_xi + 1, ny = r.max_yi - r.min_yi + 1;
    if (nx <= 0 || ny <= 0 || nx > MAX_BODY_CELLS || ny > MAX_BODY_CELLS || nx * ny > MAX_BODY_CELLS) {
        r.nx = 0;
        r.count = 0;
        // an empty range is legal only for NaN poses; flag everything else that is not a sane box
        if (!(nx <= 0 || ny <= 0) || a.x != a.x || a.y != a.y) atomic_or_u32(&d.counters->err, ERR_GRID_RANGE);
    } else {
        r.nx = (uint32_t)nx;
        r.count = (uint32_t)(nx * ny);
    }
    return r;
}
// k-th cell of the range in the reference's visiting order (yi outer ascending, xi inner ascending) -> global bucket
R2D_HD uint32_t cell_bucket(const CellRange& r, uint32_t k) {
    const int64_t yi = r.min_yi + (int64_t)(k / r.nx);
    const int64_t xi = r.min_xi + (int64_t)(k % r.nx);
    return r.bucket_base + (uint32_t)cell_hash_magic(xi, yi, (uint64_t)r.table_size, r.magic);
}

// The world vertices and edge normals of a rectangle are evaluated once per body and process() call (K2) and
// fetched here — one body is in several candidate pairs, and the four normalisations dominate make_view.
R2D_HD BodyView load_view(const Dev& d, uint32_t i) {
    const float4 p = d.pose[i];
    const float4 s = d.shape[i];
    BodyView v;
    v.pos = mk2(p.x, p.y);
    v.c = p.z;
    v.s = p.w;
    v.flags = f2u(s.z);
    v.id = f2u(s.w);
    if (v.flags & FLAG_RECT) {
        v.a = fdiv(s.x, 2.0f);  // Rectangle.zig:38-39
        v.b = fdiv(s.y, 2.0f);
        const float4 q0 = d.view[4 * (size_t)i], q1 = d.view[4 * (size_t)i + 1], q2 = d.view[4 * (size_t)i + 2],
                     q3 = d.view[4 * (size_t)i + 3];
        v.wv[0] = mk2(q0.x, q0.y); v.wv[1] = mk2(q0.z, q0.w); v.wv[2] = mk2(q1.x, q1.y); v.wv[3] = mk2(q1.z, q1.w);
        v.en[0] = mk2(q2.x, q2.y); v.en[1] = mk2(q2.z, q2.w); v.en[2] = mk2(q3.x, q3.y); v.en[3] = mk2(q3.z, q3.w);
    } else {
        v.a = s.x;
        v.b = 0.0f;
        v.wv[0] = v.wv[1] = v.wv[2] = v.wv[3] = v.pos;
        v.en[0] = v.en[1] = v.en[2] = v.en[3] = mk2(0.0f, 0.0f);
    }
    return v;
}
// K2 side of it
R2D_HD void store_view(const Dev& d, uint32_t i, const float4& pose) {
    const float4 s = d.shape[i];
    if (!(f2u(s.z) & FLAG_RECT)) return;
    const BodyView v = make_view(pose.x, pose.y, pose.z, pose.w, s.x, s.y, f2u(s.z), f2u(s.w));
    d.view[4 * (size_t)i] = make_float4(v.wv[0].x, v.wv[0].y, v.wv[1].x, v.wv[1].y);
    d.view[4 * (size_t)i + 1] = make_float4(v.wv[2].x, v.wv[2].y, v.wv[3].x, v.wv[3].y);
    d.view[4 * (size_t)i + 2] = make_float4(v.en[0].x, v.en[0].y, v.en[1].x, v.en[1].y);
    d.view[4 * (size_t)i + 3] = make_float4(v.en[2].x, v.en[2].y, v.en[3].x, v.en[3].y);
}
// ---- fine grid ---------------------------------------------------------------------------------------------------------
// The reference's candidate set is {pairs whose stored AABBs intersect} ∩ {pairs that share a hashed 4 m bucket}.  The
// first factor is geometry and can be found with any grid; the second is a lookup.  Searching it *through* the 4 m
// buckets costs n^2 tests per bucket (a settled pile puts ~47 bodies in every occupied bucket: 4.3 M tests for 260 k
// pairs at 100 k bodies).  So: every body that is narrower than the fine cell f (FLAG_LARGE clear; f ~ the widest
// common body, chosen by the host at upload) lists ONE home cell — floor(aabb centre / f) — in a second table.  Two
// small bodies whose AABBs intersect have home cells at most one apart, so a body finds its small partners in its own
// cell and the four "forward" neighbours (each pair exactly once), ~7 AABB tests per body instead of ~43, and accepts
// them if their remembered coarse bucket lists (d.bkt) share an entry — the reference's condition, hash collisions
// included.  Only FLAG_LARGE bodies (floor, walls, long bars) are entered in the coarse buckets; a small body looks them
// up through its own <= 4 coarse buckets, which IS the reference's query.  Large-large pairs come from the bucket
// kernels of the original pipeline run on the (now nearly empty) coarse buckets, and only when a dynamic large body exists.
R2D_HD bool body_is_small(const Dev& d, uint32_t flags) { return d.fine_on != 0u && !(flags & FLAG_LARGE); }
R2D_HD int32_t fine_coord(float v, double inv) {
    if (!(v == v)) return 0;
    double t = floor(dmul((double)v, inv));
    if (t > 1073741824.0) t = 1073741824.0;
    if (t < -1073741824.0) t = -1073741824.0;
    return (int32_t)t;
}
// bucket of a fine cell: row-major with a long odd stride, so that x-neighbours are neighbours in the table (coalesced
// lookups for bodies in spatial order); world w owns the fine buckets [T + mult base_w, T + mult base_{w+1})
R2D_HD uint32_t fine_bucket(const Dev& d, uint32_t world, int32_t cx, int32_t cy) {
    const uint32_t b0 = d.world_base[world], b1 = d.world_base[world + 1];
    const uint32_t h = (uint32_t)cx + (uint32_t)cy * 40503u;
    return d.n_buckets + d.table_mult * b0 + h % (d.table_mult * (b1 - b0));
}
R2D_HD uint32_t fine_tag(int32_t cx, int32_t cy) { return ((uint32_t)cx & 0xFFFFu) | ((uint32_t)cy << 16); }

// K2: pose cache + cell count (SpatialHash.zig:46-49).  Returns the cell range so the caller can walk big bodies
// cooperatively; small bodies are counted here.  With the fine grid a small body counts its home cell instead (and the
// returned range is empty).
R2D_HD CellRange count_body_thread(const Dev& d, uint32_t i, bool count_inline) {
    const float4 p = d.pos[i];
    const uint32_t flags = body_flags(d, i);
    // the rotation only enters through a rectangle's vertices and normals (store_view); a disc's pose carries no angle
    const bool rect = (flags & FLAG_RECT) != 0;
    float sn = 0.0f, cs = 1.0f;
    if (rect) sincos_ref(p.z, &sn, &cs);
    const float4 pose = make_float4(p.x, p.y, cs, sn);
    d.pose[i] = pose;
    store_view(d, i, pose);
    CellRange r = cell_range(d, i);
    d.ncells[i] = r.count;
    const bool small = body_is_small(d, flags);
    uint32_t bk[4] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
    if (r.count <= 4u) {  // small body: remember its buckets (pair de-duplication without touching the grid again)
        for (uint32_t k = 0; k < r.count; ++k) {
            bk[k] = cell_bucket(r, k);
            if (count_inline && !small) atomic_add_u32(&d.bucket_cnt[bk[k]], 1u);
        }
    } else if (small) {
        atomic_or_u32(&d.counters->err, ERR_FINE);
    } else if (count_inline) {
        for (uint32_t k = 0; k < r.count; ++k) atomic_add_u32(&d.bucket_cnt[cell_bucket(r, k)], 1u);
    }
    d.bkt[i] = make_uint4(bk[0], bk[1], bk[2], bk[3]);
    if (small) {
        const float4 a = d.aabb[i];
        const int32_t cx = fine_coord(a.x, d.fine_inv), cy = fine_coord(a.y, d.fine_inv);
        const uint32_t fb = fine_bucket(d, flags >> FLAG_WORLD_SHIFT, cx, cy);
        // the count's own atomic hands out the body's place in the bucket: the fill needs no second atomic (and no hash)
        const uint32_t rank = atomic_add_u32(&d.bucket_cnt[fb], 1u);
        d.fcell[i] = make_int4(cx, cy, (int32_t)fb, (int32_t)rank);
        r.count = 0;
    }
    return r;
}
// K4: fill (SpatialHash.zig:62-68): decrement-then-store; leaves bucket_cnt all zero again.
R2D_HD void fill_cell(const Dev& d, uint32_t i, uint32_t bucket, uint32_t key) {
    const uint32_t left = atomic_sub_u32(&d.bucket_cnt[bucket], 1u) - 1u;
    const uint32_t at = d.bucket_start[bucket] + left;
    if (at < d.cap_entries) {
        d.ent_body[at] = i;
        d.ent_key[at] = key;
    }
}
R2D_HD void fill_cell(const Dev& d, uint32_t i, uint32_t bucket) { fill_cell(d, i, bucket, bucket); }
// K4, small body with the fine grid: one entry, keyed by the low bits of its home cell, carrying what the pair test
// needs (AABB, static flag) so that the test does not have to chase the body
constexpr uint32_t ENT_STATIC = 0x80000000u;
R2D_HD void fill_fine(const Dev& d, uint32_t i) {
    const int4 c = d.fcell[i];
    const uint32_t bucket = (uint32_t)c.z, at = d.bucket_start[bucket] + (uint32_t)c.w;
    d.bucket_cnt[bucket] = 0u;   // (every body of the bucket writes the same zero) the counts are all zero again after the fill
    if (at < d.cap_entries) {
        d.ent_body[at] = i | ((body_flags(d, i) & FLAG_STATIC) ? ENT_STATIC : 0u);
        d.ent_key[at] = fine_tag(c.x, c.y);
        d.ent_aabb[at] = d.aabb[i];
    }
}
R2D_HD uint32_t bucket_end(const struct Dev& d, uint32_t bucket);
// K4b: make the bucket order deterministic (ascending slot, duplicates adjacent)
R2D_HD void sort_bucket_thread(const Dev& d, uint32_t bucket) {
    const uint32_t s = d.bucket_start[bucket];
    const uint32_t e = bucket_end(d, bucket);
    for (uint32_t a = s + 1; a < e; ++a) {
        const uint32_t v = d.ent_body[a];
        uint32_t b = a;
        while (b > s && d.ent_body[b - 1] > v) {
            d.ent_body[b] = d.ent_body[b - 1];
            --b;
        }
        d.ent_body[b] = v;
    }
}

R2D_HD bool pair_excluded(const Dev& d, uint32_t i, uint32_t j) {  // lib.zig:275-276 (both orders stored, Q22)
    if (d.n_excl == 0) return false;
    const uint64_t key = ((uint64_t)(i < j ? i : j) << 32) | (uint64_t)(i < j ? j : i);
    uint32_t lo = 0, hi = d.n_excl;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        const uint64_t v = d.excl[mid];
        if (v == key) return true;
        if (v < key)
            lo = mid + 1;
        else
            hi = mid;
    }
    return false;
}
// end of a bucket, clamped so that an over-capacity step (detected and redone by the host) never reads out of bounds
R2D_HD uint32_t bucket_end(const Dev& d, uint32_t bucket) {
    const uint32_t e = d.bucket_start[bucket + 1];
    return e < d.cap_entries ? e : d.cap_entries;
}
R2D_HD bool bucket_contains(const Dev& d, uint32_t bucket, uint32_t j) {
    uint32_t lo = d.bucket_start[bucket], hi = bucket_end(d, bucket);
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        const uint32_t v = d.ent_body[mid];
        if (v == j) return true;
        if (v < j)
            lo = mid + 1;
        else
            hi = mid;
    }
    return false;
}

// The candidate test for two distinct bodies i, j that both list bucket h (lib.zig:273-282 + de-duplication).
// A pair {i, j} sharing several buckets is accepted exactly once: in the lowest-numbered bucket they share; the test
// walks the cells of the body with fewer cells (the "owner"; ties: lower slot).  Filters are the reference's: both
// static, excluded, AABB — duplicates are removed by construction instead of by manifold-map probes.
// On success *out = (owner slot, other slot).
R2D_HD bool pair_candidate(const Dev& d, uint32_t h, uint32_t i, uint32_t j, const float4& ai, const float4& aj,
                           uint32_t fi, uint32_t fj, uint32_t nci, uint32_t ncj, const uint4& bi, const uint4& bj, uint2* out) {
    if (fi & fj & FLAG_STATIC) return false;                                               // :273
    if (!aabb_intersects(ai.x, ai.y, ai.z, ai.w, aj.x, aj.y, aj.z, aj.w)) return false;    // :282
    if (pair_excluded(d, i, j)) return false;                                              // :275-276
    const bool i_owns = nci < ncj || (nci == ncj && i < j);
    const uint32_t o = i_owns ? i : j, q = i_owns ? j : i;
    if (nci <= 4u && ncj <= 4u) {
        // both small: the lowest shared bucket follows from the two remembered bucket lists (pure ALU)
        const uint32_t a[4] = {bi.x, bi.y, bi.z, bi.w}, b[4] = {bj.x, bj.y, bj.z, bj.w};
        uint32_t lowest = 0xFFFFFFFFu;
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y)
                if (a[x] == b[y] && a[x] < lowest) lowest = a[x];
        if (lowest != h) return false;
    } else if ((i_owns ? nci : ncj) > 1) {  // a multi-cell body is involved: walk the owner's cells, bisect the buckets
        const CellRange r = cell_range(d, o);
        for (uint32_t k = 0; k < r.count; ++k) {
            const uint32_t h2 = cell_bucket(r, k);
            if (h2 < h && bucket_contains(d, h2, q)) return false;
        }
    }
    *out = make_uint2(o, q);
    return true;
}

// K5, one-thread-per-grid-entry form (used by the serial test emulator; the GPU uses the warp-per-bucket kernels of
// r2d_kernels.cuh, which accept exactly the same pairs): pairs of entry e = (bucket h, body i) that i owns.
// `out` == nullptr counts only.
R2D_HD uint32_t entry_pairs_thread(const Dev& d, uint32_t e, uint2* out) {
    const uint32_t h = d.ent_key[e], i = d.ent_body[e];
    const uint32_t bs = d.bucket_start[h], be = bucket_end(d, h);
    if (e > bs && d.ent_body[e - 1] == i) return 0;  // i lists this bucket more than once: first occurrence only
    const float4 ai = d.aabb[i];
    const uint32_t fi = body_flags(d, i);
    const uint32_t nci = d.ncells[i];
    uint32_t n = 0, prev = 0xFFFFFFFFu;
    for (uint32_t f = bs; f < be; ++f) {
        const uint32_t j = d.ent_body[f];
        if (j == prev) continue;
        prev = j;
        if (j == i) continue;                                        // :274
        const uint32_t ncj = d.ncells[j];
        if (!(nci < ncj || (nci == ncj && i < j))) continue;         // j owns this pair
        uint2 pr;
        if (!pair_candidate(d, h, i, j, ai, d.aabb[j], fi, body_flags(d, j), nci, ncj, d.bkt[i], d.bkt[j], &pr)) continue;
        if (out) out[n] = pr;
        ++n;
    }
    return n;
}

// K5 with the fine grid: the pairs emitted by small body a — its small partners from the home cell and the four forward
// neighbours in the fine table, then its large partners from its own coarse buckets.  The first PARK accepted partners are
// returned in `got` (if non-null); all of them are written to `out` as (a, partner) if non-null.  Returns how many.
R2D_HD bool bucket_lists_share(const uint4& x, const uint4& y) {
    const uint32_t a[4] = {x.x, x.y, x.z, x.w}, b[4] = {y.x, y.y, y.z, y.w};
    bool share = false;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) share = share || (a[i] != 0xFFFFFFFFu && a[i] == b[j]);
    return share;
}
template <uint32_t PARK = 8>
R2D_HD uint32_t fine_body_pairs(const Dev& d, uint32_t a, uint32_t* got, uint2* out) {
    const uint32_t fa = body_flags(d, a);
    const float4 aa = d.aabb[a];
    const int4 ca = d.fcell[a];
    const uint4 ba = d.bkt[a];
    const uint32_t world = fa >> FLAG_WORLD_SHIFT;
    const bool a_static = (fa & FLAG_STATIC) != 0;
    uint32_t n = 0;
    // ---- small partners: (0,0) (1,0) (-1,1) (0,1) (1,1) ----
    uint32_t fs[5], fe[5], ft[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {  // all lookups are issued before the first entry is needed
        const int32_t ox = (k == 0 || k == 3) ? 0 : ((k == 2) ? -1 : 1), oy = k >= 2 ? 1 : 0;
        const uint32_t h = fine_bucket(d, world, ca.x + ox, ca.y + oy);
        ft[k] = fine_tag(ca.x + ox, ca.y + oy);
        fs[k] = d.bucket_start[h];
        fe[k] = bucket_end(d, h);
    }
    const uint32_t bk[4] = {ba.x, ba.y, ba.z, ba.w};
    uint32_t cs[4], ce[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        bool use = bk[q] != 0xFFFFFFFFu;
#pragma unroll
        for (int q2 = 0; q2 < q; ++q2) use = use && bk[q2] != bk[q];  // two cells of a in one bucket: visited once
        cs[q] = use ? d.bucket_start[bk[q]] : 0u;
        ce[q] = use ? bucket_end(d, bk[q]) : 0u;
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        for (uint32_t e = fs[k]; e < fe[k]; ++e) {
            if (d.ent_key[e] != ft[k]) continue;               // another cell of the same bucket
            const uint32_t bw = d.ent_body[e], b = bw & ~ENT_STATIC;
            if (k == 0 && b <= a) continue;                    // same home cell: the lower slot emits the pair (:274 for b == a)
            if (a_static && (bw & ENT_STATIC)) continue;                                            // :273
            const float4 ab = d.ent_aabb[e];
            if (!aabb_intersects(aa.x, aa.y, aa.z, aa.w, ab.x, ab.y, ab.z, ab.w)) continue;         // :282
            if (!bucket_lists_share(ba, d.bkt[b])) continue;   // the reference only sees b through a shared hashed bucket
            if (pair_excluded(d, a, b)) continue;                                                   // :275-276
            if (got && n < PARK) got[n] = b;
            if (out) out[n] = make_uint2(a, b);
            ++n;
        }
    }
    // ---- large partners: only FLAG_LARGE bodies are entered in the coarse buckets ----
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        for (uint32_t e = cs[q]; e < ce[q]; ++e) {
            const uint32_t b = d.ent_body[e];
            bool seen = false;
            for (uint32_t e2 = cs[q]; e2 < e; ++e2) seen = seen || d.ent_body[e2] == b;   // b lists this bucket twice
#pragma unroll
            for (int q2 = 0; q2 < q; ++q2)                                                 // ... or an earlier bucket of a
                for (uint32_t e2 = cs[q2]; e2 < ce[q2] && !seen; ++e2) seen = seen || d.ent_body[e2] == b;
            if (seen) continue;
            if (a_static && (body_flags(d, b) & FLAG_STATIC)) continue;
            const float4 ab = d.aabb[b];
            if (!aabb_intersects(aa.x, aa.y, aa.z, aa.w, ab.x, ab.y, ab.z, ab.w)) continue;
            if (pair_excluded(d, a, b)) continue;
            if (got && n < PARK) got[n] = b;
            if (out) out[n] = make_uint2(a, b);
            ++n;
        }
    }
    return n;
}
// makes the order of a body's pairs independent of the order the fill's atomics landed in
R2D_HD void sort_item_pairs(uint2* out, uint32_t n) {
    for (uint32_t x = 1; x < n; ++x) {
        const uint2 v = out[x];
        uint32_t y = x;
        while (y > 0 && out[y - 1].y > v.y) {
            out[y] = out[y - 1];
            --y;
        }
        out[y] = v;
    }
}

// ---- narrowphase ---------------------------------------------------------------------------------------------------
// the manifold lists of the dataflow colouring (arrival order is irrelevant: only "how many have a higher priority" is used)
R2D_HD void adj_append(const Dev& d, uint32_t body, unsigned long long prio) {
    const uint32_t k = atomic_add_u32(&d.adj_cnt[body], 1u);
    if (k < ADJ_CAP) {
        d.adj_prio[(size_t)body * ADJ_CAP + k] = prio;
        return;
    }
    // a hub (a large body resting on many small ones): the rest of its list is a chain through a shared pool
    if (k < ADJ_MAX) {
        const uint32_t e = atomic_add_u32(&d.counters->n_adj_over, 1u);
        if (e < d.adj_pool_cap) {
            const uint32_t next = atomic_exch_u32(&d.adj_head[body], e + 1u);
            d.adj_pool[e] = make_uint4((uint32_t)prio, (uint32_t)(prio >> 32), next, 0u);   // (read by the colouring kernel only)
            return;
        }
    }
    atomic_or_u32(&d.counters->flow_abort, 1u);
}
// K6: one candidate pair -> raw manifold slot.  Returns the number of contact points, or -1 if SAT finds a gap.
R2D_HD int narrow_pair_thread(const Dev& d, uint32_t p) {
    const uint2 pr = d.pairs[p];
    // performNarrowSAT orders the two passes by id (collision.zig:304-307): the slots are put in id order BEFORE the views
    // are loaded, so that there is ONE call of narrowphase() — `a_lo ? narrowphase(va, vb) : narrowphase(vb, va)` inlines
    // it twice and a warp runs both copies with half of its lanes each (ncu: 15 of 32 lanes on an all-disc scene).
    const bool x_lo = body_id(d, pr.x) < body_id(d, pr.y);
    const uint32_t lo_slot = x_lo ? pr.x : pr.y, hi_slot = x_lo ? pr.y : pr.x;
    const BodyView vlo = load_view(d, lo_slot), vhi = load_view(d, hi_slot);
    const Manifold m = narrowphase(vlo, vhi);
    if (!m.collides) {
        d.m_color[p] = COLOR_NONE;
        return -1;
    }
    const uint32_t ref = m.ref_is_lo ? lo_slot : hi_slot, inc = m.ref_is_lo ? hi_slot : lo_slot;
    const uint32_t f_ref = m.ref_is_lo ? vlo.flags : vhi.flags, f_inc = m.ref_is_lo ? vhi.flags : vlo.flags;
    const uint32_t dyn = ((f_ref & FLAG_STATIC) ? 0u : 1u) | ((f_inc & FLAG_STATIC) ? 0u : 2u);
    d.m_hdr[p] = make_uint4(ref, inc, (uint32_t)m.n_points | ((uint32_t)m.normal_id << 8), dyn);
    const unsigned long long prio = contact_priority(vlo.id, vhi.id);  // (lower id, higher id)
    d.m_prio[p] = prio;
    if (d.flow) {
        if (dyn & 1u) adj_append(d, ref, prio);
        if (dyn & 2u) adj_append(d, inc, prio);
    }
    const ContactPoint z = {mk2(0, 0), 0.0f, mk2(0, 0), mk2(0, 0)};
    const ContactPoint p0 = m.n_points > 0 ? m.pt[0] : z, p1 = m.n_points > 1 ? m.pt[1] : z;
    d.m_g0[p] = make_float4(m.normal.x, m.normal.y, p0.pos.x, p0.pos.y);
    d.m_g1[p] = make_float4(p0.depth, p1.depth, p1.pos.x, p1.pos.y);
    d.m_r0[p] = make_float4(p0.ref_r.x, p0.ref_r.y, p0.inc_r.x, p0.inc_r.y);
    d.m_r1[p] = make_float4(p1.ref_r.x, p1.ref_r.y, p1.inc_r.x, p1.inc_r.y);
    d.m_color[p] = COLOR_PENDING;
    return m.n_points;
}

// ---- graph colouring (Jones-Plassmann with id-derived priorities; see contact_priority in r2d_solve.cuh) -----------------
R2D_HD void color_post(const Dev& d, uint32_t ref, uint32_t inc, bool dyn1, bool dyn2, uint64_t prio, uint32_t round) {
    unsigned long long* mp = (round & 1u) ? d.maxprio1 : d.maxprio0;
    const unsigned long long v = ((unsigned long long)round << PRIO_ROUND_SHIFT) | prio;
    if (dyn1) atomic_max_u64(&mp[ref], v);
    if (dyn2) atomic_max_u64(&mp[inc], v);
}
// One colouring round for a pending manifold whose header / priority the caller holds (registers across rounds).
// Returns 1 coloured now (colour in *out_color), 2 still pending (re-posted for round + 1), 3 dropped (no colour free).
R2D_HD int color_round_core(const Dev& d, uint32_t p, const uint4& h, uint64_t prio, uint32_t round, uint32_t* out_color) {
    const bool dyn1 = (h.w & 1u) != 0, dyn2 = (h.w & 2u) != 0;
    const unsigned long long mine = ((unsigned long long)round << PRIO_ROUND_SHIFT) | prio;
    const unsigned long long* mp = (round & 1u) ? d.maxprio1 : d.maxprio0;
    const bool sm = d.color_smem != 0u;
    const bool win = (!dyn1 || ld_shared_u64(&mp[h.x], sm) == mine) && (!dyn2 || ld_shared_u64(&mp[h.y], sm) == mine);
    if (!win) {
        color_post(d, h.x, h.y, dyn1, dyn2, prio, round + 1);
        return 2;
    }
    uint32_t color = MAX_COLORS;
    for (uint32_t w = 0; w < COLOR_WORDS; ++w) {
        unsigned long long u = 0;
        if (dyn1) u |= ld_shared_u64(&d.used[(size_t)h.x * COLOR_WORDS + w], sm);
        if (dyn2) u |= ld_shared_u64(&d.used[(size_t)h.y * COLOR_WORDS + w], sm);
        if (~u) {
            color = w * 64 + first_zero64(u);
            break;
        }
    }
    if (color >= MAX_COLORS) {   // every colour is taken on one of its bodies: the manifold sits this call out (R2D_MAX_COLORS)
        atomic_add_u32(&d.counters->n_dropped, 1u);
        d.m_color[p] = COLOR_DROPPED;
        return 3;
    }
    const unsigned long long bit = 1ull << (color & 63u);
    // unique winner per body and round: a plain read-modify-write cannot race
    if (dyn1) {
        unsigned long long* u = &d.used[(size_t)h.x * COLOR_WORDS + (color >> 6)];
        *u = ld_shared_u64(u, sm) | bit;
    }
    if (dyn2) {
        unsigned long long* u = &d.used[(size_t)h.y * COLOR_WORDS + (color >> 6)];
        *u = ld_shared_u64(u, sm) | bit;
    }
    d.m_color[p] = color;
    *out_color = color;
    return 1;
}
// Same, reading everything from memory.  Returns 0 if slot p holds nothing pending.
R2D_HD int color_round_thread(const Dev& d, uint32_t p, uint32_t round) {
    if (d.m_color[p] != COLOR_PENDING) return 0;
    uint32_t c;
    return color_round_core(d, p, d.m_hdr[p], d.m_prio[p], round, &c);
}

// ---- dataflow form of the same colouring ----------------------------------------------------------------------------
// Greedy colouring in descending priority means: a manifold takes the lowest colour free on both bodies once every
// manifold with a HIGHER priority on either body has taken its own.  With r_b = the number of such manifolds on body
// b (counted in the body's list), that is "wait until body b has coloured exactly r_b manifolds" — no rounds, no
// barriers, a manifold only waits for its own two bodies, exactly like the solver sweep.  The per-body word
// {mask[0..95], coloured} is read and written as ONE 16-byte relaxed access (see ld_body_word), and only one manifold
// per body is ever allowed to write, so there is no atomic and no fence.  Result: identical to the rounds above.
R2D_HD uint4 ld_cstate(const uint4* p) {
    const float4 v = ld_body_word(reinterpret_cast<const float4*>(p));
    return make_uint4(f2u(v.x), f2u(v.y), f2u(v.z), f2u(v.w));
}
R2D_HD void st_cstate(uint4* p, uint4 v) {
    st_body_word(reinterpret_cast<float4*>(p), make_float4(u2f(v.x), u2f(v.y), u2f(v.z), u2f(v.w)));
}
// ranks of a manifold on its two bodies: r1 | r2 << 8  (0 for a static body)
R2D_HD uint32_t flow_ranks(const Dev& d, const uint4& h, unsigned long long prio) {
    uint32_t r[2] = {0u, 0u};
    const uint32_t bs[2] = {h.x, h.y};
    for (int q = 0; q < 2; ++q) {
        if (!(h.w & (1u << q))) continue;
        uint32_t n = d.adj_cnt[bs[q]];
        const bool chained = n > ADJ_CAP;
        if (chained) n = ADJ_CAP;
        const unsigned long long* list = d.adj_prio + (size_t)bs[q] * ADJ_CAP;
        for (uint32_t k = 0; k < n; ++k) r[q] += list[k] > prio ? 1u : 0u;
        if (chained)
            for (uint32_t e = d.adj_head[bs[q]]; e != 0u && e <= d.adj_pool_cap;) {   // (at most ADJ_MAX - ADJ_CAP entries)
                const uint4 x = d.adj_pool[e - 1u];
                r[q] += (((unsigned long long)x.y << 32) | x.x) > prio ? 1u : 0u;
                e = x.z;
            }
    }
    return r[0] | (r[1] << 8);
}
// One probe.  Returns 1 (coloured, colour in *out), 0 (not ready; *lag = how many manifolds the slower body still
// has to colour first) or -1 (more than FLOW_COLORS colours needed: flow_fail is raised).
R2D_HD int flow_try(const Dev& d, uint32_t p, uint32_t ref, uint32_t inc, uint32_t dyn, uint32_t ranks, uint32_t* out, uint32_t* lag) {
    const bool dyn1 = (dyn & 1u) != 0, dyn2 = (dyn & 2u) != 0;
    uint4 w1 = make_uint4(0u, 0u, 0u, 0u), w2 = w1;
    if (dyn1) w1 = ld_cstate(&d.cstate[ref]);
    if (dyn2) w2 = ld_cstate(&d.cstate[inc]);
    const uint32_t lag1 = dyn1 ? (ranks & 0xFFu) - w1.w : 0u, lag2 = dyn2 ? (ranks >> 8) - w2.w : 0u;
    if (lag1 | lag2) {
        *lag = lag1 > lag2 ? lag1 : lag2;
        return 0;
    }
    const uint32_t u[3] = {w1.x | w2.x, w1.y | w2.y, w1.z | w2.z};
    uint32_t color = FLOW_COLORS;
    for (uint32_t w = 0; w < 3; ++w)
        if (~u[w]) {
            color = w * 32u + first_zero32(u[w]);
            break;
        }
    if (color >= FLOW_COLORS) {
        atomic_or_u32(&d.counters->flow_fail, 1u);
        return -1;
    }
    const uint32_t bit = 1u << (color & 31u), word = color >> 5;
    if (dyn1) st_cstate(&d.cstate[ref], make_uint4(w1.x | (word == 0 ? bit : 0u), w1.y | (word == 1 ? bit : 0u),
                                                   w1.z | (word == 2 ? bit : 0u), w1.w + 1u));
    if (dyn2) st_cstate(&d.cstate[inc], make_uint4(w2.x | (word == 0 ? bit : 0u), w2.y | (word == 1 ? bit : 0u),
                                                   w2.z | (word == 2 ? bit : 0u), w2.w + 1u));
    d.m_color[p] = color;
    *out = color;
    return 1;
}
// colours below `color` on body b and colours used on it at all (the dataflow order of the solver sweep)
R2D_HD void body_color_rank(const Dev& d, bool from_flow, uint32_t b, uint32_t color, uint32_t* rank, uint32_t* degree) {
    uint32_t rk = 0, dg = 0;
    if (from_flow) {
        const uint4 w = d.cstate[b];
        const uint32_t u[3] = {w.x, w.y, w.z};
        for (uint32_t k = 0; k < 3; ++k) {
            dg += popc64(u[k]);
            if (k < (color >> 5))
                rk += popc64(u[k]);
            else if (k == (color >> 5))
                rk += popc64(u[k] & ((1u << (color & 31u)) - 1u));
        }
    } else {
        for (uint32_t w = 0; w < COLOR_WORDS; ++w) {
            const unsigned long long u = d.used[(size_t)b * COLOR_WORDS + w];
            dg += popc64(u);
            if (w < (color >> 6))
                rk += popc64(u);
            else if (w == (color >> 6))
                rk += popc64(u & ((1ull << (color & 63u)) - 1ull));
        }
    }
    *rank = rk;
    *degree = dg;
}

// The "owner" of a manifold orders the colour-sorted solver records: the lower device slot of its non-static bodies.
// Colours on one body are pairwise distinct, so a body owns at most one manifold per colour and the position of a
// manifold inside its colour is the number of owners of that colour with a lower slot — a prefix popcount over a
// bitmap.  Device slots are in spatial (Morton) order, hence so are the records of every colour: a warp of the sweep
// touches bodies that are neighbours in memory.
R2D_HD uint32_t manifold_owner(const uint4& h) {
    const bool dyn1 = (h.w & 1u) != 0, dyn2 = (h.w & 2u) != 0;
    if (dyn1 && dyn2) return h.x < h.y ? h.x : h.y;
    return dyn1 ? h.x : h.y;
}
// the same for a manifold whose colour has just been decided (the colouring kernels set the bit on the spot: one launch less)
R2D_HD void owner_bit_set(const Dev& d, uint32_t ref, uint32_t inc, uint32_t dyn, uint32_t color) {
    if (d.world_fused) return;
    const uint32_t o = manifold_owner(make_uint4(ref, inc, 0u, dyn));
    atomic_or_u32(&d.own_bits[(size_t)color * d.own_words + (o >> 5)], 1u << (o & 31u));
}
R2D_HD void owner_bit_thread(const Dev& d, uint32_t p) {
    const uint32_t c = d.m_color[p];
    if (c >= MAX_COLORS || d.world_fused) return;
    const uint32_t o = manifold_owner(d.m_hdr[p]);
    atomic_or_u32(&d.own_bits[(size_t)c * d.own_words + (o >> 5)], 1u << (o & 31u));
}
// scan input: popcount per bitmap word, and one extra entry per colour that pads its segment to a whole warp
R2D_HD void owner_count_thread(const Dev& d, uint32_t k) {
    const uint32_t stride = d.own_words + 1u;
    const uint32_t c = k / stride, w = k % stride;
    if (w < d.own_words) {
        d.own_pos[k] = popc64(d.own_bits[(size_t)c * d.own_words + w]);
    } else {
        const uint32_t n = d.color_count[c];
        d.own_pos[k] = (COLOR_ALIGN - (n % COLOR_ALIGN)) % COLOR_ALIGN;
    }
}
R2D_HD uint32_t manifold_slot(const Dev& d, uint32_t p) {
    const uint32_t c = d.m_color[p];
    const uint32_t o = manifold_owner(d.m_hdr[p]);
    const uint32_t word = d.own_bits[(size_t)c * d.own_words + (o >> 5)];
    return d.own_pos[(size_t)c * (d.own_words + 1u) + (o >> 5)] + popc64(word & ((1u << (o & 31u)) - 1u));
}

// ---- warm starting (R2D_OPT_WARM_START) ---------------------------------------------------------------------------------
constexpr unsigned long long WARM_EMPTY = 0xFFFFFFFFFFFFFFFFull;
constexpr uint32_t WARM_MAX_ID = 1u << 24, WARM_MAX_WORLD = 1u << 16, WARM_PROBES = 128;
constexpr float SLEEP_LIN2 = 0.04f, SLEEP_ANG2 = 0.04f;   // (0.2 m/s)^2, (0.2 rad/s)^2: the Baumgarte push-out alone keeps a resting pile at ~0.1 m/s
R2D_HD unsigned long long warm_make_key(uint32_t world, uint32_t ref_id, uint32_t inc_id) {
    return ((unsigned long long)world << 48) | ((unsigned long long)ref_id << 24) | (unsigned long long)inc_id;
}
R2D_HD uint32_t warm_hash(unsigned long long key) {
    key ^= key >> 33;
    key *= 0xff51afd7ed558ccdull;
    key ^= key >> 33;
    return (uint32_t)key;
}
// the previous call's entry of this contact, if it had the same reference face and point count
R2D_HD bool warm_lookup(const Dev& d, unsigned long long key, uint32_t meta, float4* out) {
    uint32_t h = warm_hash(key) & d.warm_mask;
    for (uint32_t probe = 0; probe < WARM_PROBES; ++probe, h = (h + 1u) & d.warm_mask) {
        const unsigned long long k = d.warm_key[h];
        if (k == WARM_EMPTY) return false;
        if (k == key) {
            if (d.warm_meta[h] != meta) return false;
            *out = d.warm_val[h];
            return true;
        }
    }
    return false;
}

// ---- colour partition + pre-step (collision.zig:102-133, evaluated once per process(): inputs are constant, Q5) -----------
R2D_HD void gather_prestep_thread(const Dev& d, uint32_t p, uint32_t at) {
    const uint4 h = d.m_hdr[p];
    const uint32_t np = h.z & 0xFFu;
    const uint32_t f1 = body_flags(d, h.x), f2 = body_flags(d, h.y);
    const bool st1 = (f1 & FLAG_STATIC) != 0, st2 = (f2 & FLAG_STATIC) != 0;
    const float4 pr1 = d.prop[h.x], pr2 = d.prop[h.y];
    const float4 g0 = d.m_g0[p], g1 = d.m_g1[p];
    const ContactConst c = prestep_manifold(mk2(g0.x, g0.y), st1, st2, pr1.x, pr2.x, pr1.y, pr2.y, pr1.z, pr2.z);
    uint32_t warm_bit = 0u;
    if (d.warm_on) {
        float4 w;
        const uint32_t world = f1 >> FLAG_WORLD_SHIFT;
        if (warm_lookup(d, warm_make_key(world, body_id(d, h.x), body_id(d, h.y)), h.z, &w)) {
            warm_bit = S_WARM;
            d.s_warm0[at] = make_float2(w.x, w.y);
            d.s_warm1[at] = make_float2(w.z, w.w);
        }
    }
    d.s_hdr[at] = make_uint4(h.x, h.y, np | (st1 ? 0x100u : 0u) | (st2 ? 0x200u : 0u) | warm_bit, p);
    if (d.tile_bodies && !st1 && !st2 && h.x / d.tile_bodies != h.y / d.tile_bodies)
        d.body_shared[h.x > h.y ? h.x : h.y] = 1u;  // the owner is the lower slot: the other body is foreign to its tile
    // Colours on one body are pairwise distinct, so the body's sweep sequence is its colour set in ascending order:
    // rank = colours below mine, degree = colours used.
    {
        const uint32_t color = d.m_color[p];
        const bool from_flow = d.counters->flow_used != 0u;
        uint32_t rk[2] = {0, 0}, dg[2] = {0, 0};
        if (!st1) body_color_rank(d, from_flow, h.x, color, &rk[0], &dg[0]);
        if (!st2) body_color_rank(d, from_flow, h.y, color, &rk[1], &dg[1]);
        if (d.joints_flow) {   // the joints of a body come first in every iteration
            const uint32_t j1 = d.body_nj[h.x], j2 = d.body_nj[h.y];
            rk[0] += j1; dg[0] += j1; rk[1] += j2; dg[1] += j2;
        }
        d.s_dep[at] = make_uint4(rk[0], dg[0], rk[1], dg[1]);
    }
    d.s_nf[at] = make_float4(c.normal.x, c.normal.y, c.friction, 0.0f);
    d.s_inv[at] = make_float4(c.inv_m1, c.inv_m2, c.inv_i1, c.inv_i2);
    if (np > 0) {
        const float4 r = d.m_r0[p];
        ContactPointConst pc;
        pc.r1 = mk2(r.x, r.y);
        pc.r2 = mk2(r.z, r.w);
        pc.depth = g1.x;
        prestep_point(c, pc);
        d.s_r0[at] = r;
        d.s_pm0[at] = make_float4(pc.mass_n, pc.mass_t, pc.depth, contact_bias(pc.depth, d.sub_dt));
        d.s_acc0[at] = make_float2(0.0f, 0.0f);
    }
    if (np > 1) {
        const float4 r = d.m_r1[p];
        ContactPointConst pc;
        pc.r1 = mk2(r.x, r.y);
        pc.r2 = mk2(r.z, r.w);
        pc.depth = g1.y;
        prestep_point(c, pc);
        d.s_r1[at] = r;
        d.s_pm1[at] = make_float4(pc.mass_n, pc.mass_t, pc.depth, contact_bias(pc.depth, d.sub_dt));
        d.s_acc1[at] = make_float2(0.0f, 0.0f);
    }
}

// ---- solver sweeps --------------------------------------------------------------------------------------------------
// K11: one manifold (collision.zig:135-218).  Manifolds of one colour share no non-static body.
//
// DATAFLOW = false: the caller guarantees (kernel boundary / grid barrier) that all lower colours are done.
// DATAFLOW = true : no barrier between colours or iterations.  The 4th component of every body's momentum word is a
//   version: the number of contact updates applied to the body in this substep.  The k-th contact (in colour order) of
//   a body in iteration `it` runs when the version equals it * degree + k, and publishes momentum and version + 1 in
//   ONE 16-byte store.  Each body therefore sees exactly the same sequence of updates as with barriers — the result is
//   bit-identical — but a manifold only waits for its own two bodies.  All waiting threads are resident (cooperative
//   launch) and walk their manifolds in (iteration, colour) order, so the globally lowest pending manifold can always
//   run: no deadlock.  A stall would be a bug; it is reported through ERR_STALL instead of hanging the GPU.
template <bool DATAFLOW>
R2D_HD void solve_contact_thread(const Dev& d, uint32_t m, float sub_dt, uint32_t it = 0, bool first = false) {
    const uint4 h = d.s_hdr[m];
    const bool empty = (h.z & S_EMPTY) != 0;
    if (!DATAFLOW && empty) return;
    const int np = empty ? 0 : (int)(h.z & 0xFFu);
    const bool st1 = empty || (h.z & 0x100u) != 0, st2 = empty || (h.z & 0x200u) != 0;
    ContactConst c;
    ContactPointConst pts[2];
    v2 acc[2];
    v2 warm[2] = {mk2(0.0f, 0.0f), mk2(0.0f, 0.0f)};
    const bool use_warm = first && !empty && (h.z & S_WARM) != 0;   // first update of the call (R2D_OPT_WARM_START)
    if (use_warm) {
        const float2 w0 = d.s_warm0[m], w1 = d.s_warm1[m];
        warm[0] = mk2(w0.x, w0.y);
        warm[1] = mk2(w1.x, w1.y);
    }
    uint32_t e1 = 0, e2 = 0;
    float4 m1 = make_float4(0, 0, 0, 0), m2 = m1;
    if (!empty) {
        const float4 nf = d.s_nf[m];
        c.normal = mk2(nf.x, nf.y);
        c.tangent = rot90cw(c.normal);
        c.friction = nf.z;
        const float4 inv = d.s_inv[m];
        c.inv_m1 = inv.x;
        c.inv_m2 = inv.y;
        c.inv_i1 = inv.z;
        c.inv_i2 = inv.w;
        {  // point 0 is loaded unconditionally (almost every manifold has it): no dependent load level after the header
            const float4 r = d.s_r0[m], pm = d.s_pm0[m];
            const float2 a = d.s_acc0[m];
            pts[0].r1 = mk2(r.x, r.y);
            pts[0].r2 = mk2(r.z, r.w);
            pts[0].mass_n = pm.x;
            pts[0].mass_t = pm.y;
            pts[0].depth = pm.z;
            pts[0].bias = pm.w;
            acc[0] = mk2(a.x, a.y);
        }
        if (np > 1) {
            const float4 r = d.s_r1[m], pm = d.s_pm1[m];
            const float2 a = d.s_acc1[m];
            pts[1].r1 = mk2(r.x, r.y);
            pts[1].r2 = mk2(r.z, r.w);
            pts[1].mass_n = pm.x;
            pts[1].mass_t = pm.y;
            pts[1].depth = pm.z;
            pts[1].bias = pm.w;
            acc[1] = mk2(a.x, a.y);
        }
        if (DATAFLOW) {
            const uint4 dep = d.s_dep[m];
            e1 = it * dep.y + dep.x;
            e2 = it * dep.w + dep.z;
        }
        m1 = d.mom[h.x];  // static bodies: never written during a sweep
        m2 = d.mom[h.y];
    }
    if (DATAFLOW) {
        // The 32 manifolds of a warp have one colour (segments are padded to whole warps), hence no dependencies among
        // themselves.
#if defined(__CUDA_ARCH__)
        uint32_t spins = 0;
        // Stage 1: only lane 0 probes (one 16-byte load per warp instead of 64) until ITS body is ready; the other
        // lanes have the same colour and become ready at about the same time.
        for (;;) {
            uint32_t lag0 = 0u;
            if ((threadIdx.x & 31u) == 0u && !st1) lag0 = e1 - f2u(ld_body_word(&d.mom[h.x]).w);
            lag0 = __shfl_sync(0xffffffffu, lag0, 0);
            if (lag0 == 0u) break;
            if (lag0 > WAIT_SPIN_LAG) {
                const uint32_t ns = lag0 * WAIT_SLEEP_UNIT;
                backoff_ns(ns < WAIT_SLEEP_MAX ? ns : WAIT_SLEEP_MAX);
            }
            if ((++spins & 0xFFu) == 0u) {
                if (spins > (1u << 20)) atomicOr(&d.counters->err, ERR_STALL);
                const uint32_t flag = *((volatile uint32_t*)&d.counters->err) & ERR_STALL;
                if (__any_sync(0xffffffffu, flag != 0u)) return;
            }
        }
        // Stage 2: ready lanes update as they come (a few converged sub-groups per warp instead of waiting for the slowest lane).
        bool pending = !empty;
        while (__any_sync(0xffffffffu, pending)) {
            uint32_t lag = 0xffffffffu;
            if (pending) {
                if (!st1) m1 = ld_body_word(&d.mom[h.x]);
                if (!st2) m2 = ld_body_word(&d.mom[h.y]);
                lag = (st1 ? 0u : e1 - f2u(m1.w)) + (st2 ? 0u : e2 - f2u(m2.w));
            }
            if (pending && lag == 0u) {
                BodyVel b1 = {mk2(m1.x, m1.y), m1.z}, b2 = {mk2(m2.x, m2.y), m2.z};
                solve_contact(c, np, pts, acc, st1, st2, b1, b2, use_warm, warm);
                d.s_acc0[m] = make_float2(acc[0].x, acc[0].y);
                if (np > 1) d.s_acc1[m] = make_float2(acc[1].x, acc[1].y);
                if (!st1) st_body_word(&d.mom[h.x], make_float4(b1.mom.x, b1.mom.y, b1.ang, u2f(e1 + 1u)));
                if (!st2) st_body_word(&d.mom[h.y], make_float4(b2.mom.x, b2.mom.y, b2.ang, u2f(e2 + 1u)));
                pending = false;
            }
            const uint32_t best = __reduce_min_sync(0xffffffffu, pending ? lag : 0xffffffffu);
            if (best != 0xffffffffu && best > WAIT_SPIN_LAG) {
                const uint32_t ns = best * WAIT_SLEEP_UNIT;
                backoff_ns(ns < WAIT_SLEEP_MAX ? ns : WAIT_SLEEP_MAX);
            }
            if ((++spins & 0xFFu) == 0u) {
                if (spins > (1u << 20)) atomicOr(&d.counters->err, ERR_STALL);
                const uint32_t flag = *((volatile uint32_t*)&d.counters->err) & ERR_STALL;
                if (__any_sync(0xffffffffu, flag != 0u)) return;
            }
        }
        return;
#else
        if (!st1) m1 = ld_body_word(&d.mom[h.x]);
        if (!st2) m2 = ld_body_word(&d.mom[h.y]);
        const uint32_t lag = (st1 ? 0u : e1 - f2u(m1.w)) + (st2 ? 0u : e2 - f2u(m2.w));
        if (lag != 0u) {  // serial emulation: the order must already be right
            atomic_or_u32(&d.counters->err, ERR_STALL);
            return;
        }
#endif
    }
    if (empty) return;
    BodyVel b1 = {mk2(m1.x, m1.y), m1.z}, b2 = {mk2(m2.x, m2.y), m2.z};
    solve_contact(c, np, pts, acc, st1, st2, b1, b2, use_warm, warm);
    if (np > 0) d.s_acc0[m] = make_float2(acc[0].x, acc[0].y);
    if (np > 1) d.s_acc1[m] = make_float2(acc[1].x, acc[1].y);
    // static bodies receive a zero impulse in the reference (`momentum += 0`); not writing them is the same value
    if (DATAFLOW) {
        if (!st1) st_body_word(&d.mom[h.x], make_float4(b1.mom.x, b1.mom.y, b1.ang, u2f(e1 + 1u)));
        if (!st2) st_body_word(&d.mom[h.y], make_float4(b2.mom.x, b2.mom.y, b2.ang, u2f(e2 + 1u)));
    } else {
        if (!st1) d.mom[h.x] = make_float4(b1.mom.x, b1.mom.y, b1.ang, m1.w);
        if (!st2) d.mom[h.y] = make_float4(b2.mom.x, b2.mom.y, b2.ang, m2.w);
    }
}

R2D_HD JointBody load_joint_body(const Dev& d, uint32_t s) {
    const float4 p = d.pos[s], m = d.mom[s], pr = d.prop[s], f = d.frc[s];
    JointBody b;
    b.pos = mk2(p.x, p.y);
    b.angle = p.z;
    b.mom = mk2(m.x, m.y);
    b.ang = m.z;
    b.mass = pr.x;
    b.inertia = pr.y;
    b.torque = f.z;
    b.is_static = (body_flags(d, s) & FLAG_STATIC) != 0;
    return b;
}
R2D_HD void store_joint_body(const Dev& d, uint32_t s, const JointBody& b) {
    d.mom[s] = make_float4(b.mom.x, b.mom.y, b.ang, d.mom[s].w);
}
// K10: one joint of the current joint colour (Constraints/*.zig).  Joints of one colour name disjoint bodies.
R2D_HD void solve_joint_thread(const Dev& d, uint32_t j, float sub_dt) {
    const uint4 h = d.j_hdr[j];
    const float4 par = d.j_par[j], vec = d.j_vec[j];
    const float power_max = par.x, power_min = par.y, beta = par.z, target = par.w;
    switch (h.x) {
        case 0: {  // distance
            JointBody b1 = load_joint_body(d, h.y), b2 = load_joint_body(d, h.z);
            solve_distance(b1, b2, target, beta, power_min, power_max);
            store_joint_body(d, h.y, b1);
            store_joint_body(d, h.z, b2);
        } break;
        case 1: {  // offset distance
            JointBody b1 = load_joint_body(d, h.y), b2 = load_joint_body(d, h.z);
            solve_offset_distance(b1, b2, mk2(vec.x, vec.y), mk2(vec.z, vec.w), target, beta, power_min, power_max);
            store_joint_body(d, h.y, b1);
            store_joint_body(d, h.z, b2);
        } break;
        case 2: {  // fixed position
            JointBody b = load_joint_body(d, h.y);
            solve_fixed_position(b, mk2(vec.x, vec.y), beta, power_min, power_max);
            store_joint_body(d, h.y, b);
        } break;
        default: {  // motor
            JointBody b = load_joint_body(d, h.y);
            solve_motor(b, target, beta, power_min, power_max, sub_dt);
            store_joint_body(d, h.y, b);
        } break;
    }
}

// contacts of body b in this call (= colours used on it)
R2D_HD uint32_t body_contact_degree(const Dev& d, uint32_t b) {
    uint32_t rk, dg;
    body_color_rank(d, d.counters->flow_used != 0u, b, 0u, &rk, &dg);
    return dg;
}
#if defined(__CUDACC__)
// K10 inside the dataflow sweep: joint j of iteration `it` waits until each of its (non-static) bodies has received
// exactly the updates that precede it in that body's sequence — it * (joints + contacts of the body) + rank of the joint
// among the body's joints — and publishes momentum and version + 1 in ONE 16-byte store, like a contact.  Called by whole
// warps (`live` = this lane has a joint); ready lanes update as they come, every pass ends in a warp vote, so lanes of
// different joint colours in one warp cannot starve each other.
__device__ __forceinline__ void solve_joint_flow(const Dev& d, uint32_t j, bool live, float sub_dt, uint32_t it) {
    uint4 h = make_uint4(0u, 0u, 0u, 0u);
    float4 par = make_float4(0, 0, 0, 0), vec = par;
    uint32_t e1 = 0, e2 = 0;
    bool two = false, st1 = true, st2 = true;
    JointBody b1{}, b2{};
    if (live) {
        h = d.j_hdr[j];
        par = d.j_par[j];
        vec = d.j_vec[j];
        const uint4 dep = d.j_dep[j];
        two = h.x == 0u || h.x == 1u;
        b1 = load_joint_body(d, h.y);
        st1 = b1.is_static;
        e1 = it * (dep.y + body_contact_degree(d, h.y)) + dep.x;
        if (two) {
            b2 = load_joint_body(d, h.z);
            st2 = b2.is_static;
            e2 = it * (dep.w + body_contact_degree(d, h.z)) + dep.z;
        }
    }
    bool pending = live;
    uint32_t spins = 0;
    while (__any_sync(0xffffffffu, pending)) {
        if (pending) {
            float4 m1 = make_float4(0, 0, 0, 0), m2 = m1;
            uint32_t lag = 0u;
            if (!st1) {
                m1 = ld_body_word(&d.mom[h.y]);
                lag += e1 - f2u(m1.w);
            }
            if (two && !st2) {
                m2 = ld_body_word(&d.mom[h.z]);
                lag += e2 - f2u(m2.w);
            }
            if (lag == 0u) {
                if (!st1) { b1.mom = mk2(m1.x, m1.y); b1.ang = m1.z; }
                if (two && !st2) { b2.mom = mk2(m2.x, m2.y); b2.ang = m2.z; }
                const float power_max = par.x, power_min = par.y, beta = par.z, target = par.w;
                switch (h.x) {
                    case 0: solve_distance(b1, b2, target, beta, power_min, power_max); break;
                    case 1: solve_offset_distance(b1, b2, mk2(vec.x, vec.y), mk2(vec.z, vec.w), target, beta, power_min, power_max); break;
                    case 2: solve_fixed_position(b1, mk2(vec.x, vec.y), beta, power_min, power_max); break;
                    default: solve_motor(b1, target, beta, power_min, power_max, sub_dt); break;
                }
                // (two-body joints with a static body are swept with grid barriers instead: they write its momentum, Q10)
                if (!st1) st_body_word(&d.mom[h.y], make_float4(b1.mom.x, b1.mom.y, b1.ang, u2f(e1 + 1u)));
                if (two && !st2) st_body_word(&d.mom[h.z], make_float4(b2.mom.x, b2.mom.y, b2.ang, u2f(e2 + 1u)));
                pending = false;
            }
        }
        if ((++spins & 0xFFu) == 0u) {
            if (spins > (1u << 20)) atomicOr(&d.counters->err, ERR_STALL);
            const uint32_t flag = *((volatile uint32_t*)&d.counters->err) & ERR_STALL;
            if (__any_sync(0xffffffffu, flag != 0u)) return;
        }
    }
}
#endif

// ---- integrators ------------------------------------------------------------------------------------------------------
// Pure arithmetic of the three per-body steps (shared by every integrator flavour, so that they cannot drift apart).
// AABB refresh (lib.zig:210; Disc.zig:63-66, Rectangle.zig:71-86)
R2D_HD float4 refreshed_aabb(const float4& p, uint32_t flags, float shape_a, float shape_b) {
    float hw, hh;
    if (flags & FLAG_RECT) {
        float sn, cs;
        sincos_ref(p.z, &sn, &cs);
        aabb_half_extents(flags, shape_a, shape_b, cs, sn, hw, hh);
    } else {
        hw = shape_a;
        hh = shape_a;
    }
    return make_float4(p.x, p.y, hw, hh);
}
// gravity (DownwardsGravity.zig:35-39: force.addmult(g_vec = (0, -g), mass) per generator) + momentum integration
// (lib.zig:211-215).  `m.w` (the version word of the dataflow sweep) restarts at 0: contact updates are counted per substep.
R2D_HD void momentum_update(const Dev& d, uint32_t world, float4& m, float4 f, float mass, float sub_dt) {
    for (uint32_t g = d.grav_off[world]; g < d.grav_off[world + 1]; ++g) {
        f.x = fadd(f.x, fmul(0.0f, mass));
        f.y = fadd(f.y, fmul(-d.grav[g], mass));
    }
    m.x = fadd(m.x, fmul(f.x, sub_dt));
    m.y = fadd(m.y, fmul(f.y, sub_dt));
    m.z = fadd(m.z, fmul(f.z, sub_dt));
    m.w = u2f(0u);
}
// position integration (lib.zig:238-249)
R2D_HD void position_update(float4& p, const float4& m, float mass, float inertia, float sub_dt) {
    const float k = fdiv(sub_dt, mass);
    p.x = fadd(p.x, fmul(m.x, k));
    p.y = fadd(p.y, fmul(m.y, k));
    p.z = fadd(p.z, fdiv(fmul(m.z, sub_dt), inertia));
}

// K1: gravity (DownwardsGravity.zig:35-39) + AABB refresh (lib.zig:210) + momentum integration (lib.zig:211-215).
// The reference refreshes the AABB in every substep; only the last refresh is observable (the next process() and the
// accessors read it), so it is evaluated when `refresh_aabb` is set.
R2D_HD void integrate_forces_thread(const Dev& d, uint32_t i, float sub_dt, bool refresh_aabb) {
    const float4 s = d.shape[i];
    const uint32_t flags = f2u(s.z);
    if (refresh_aabb) d.aabb[i] = refreshed_aabb(d.pos[i], flags, s.x, s.y);
    if (flags & FLAG_STATIC) return;
    float4 m = d.mom[i];
    momentum_update(d, flags >> FLAG_WORLD_SHIFT, m, d.frc[i], d.prop[i].x, sub_dt);
    d.mom[i] = m;
    // force.xy is next read after the end-of-substep reset (positions kernel), so the gravity sum need not be stored
}
// K12: position integration and force reset (lib.zig:238-249)
R2D_HD void integrate_positions_thread(const Dev& d, uint32_t i, float sub_dt) {
    const uint32_t flags = body_flags(d, i);
    if (flags & FLAG_STATIC) return;
    const float4 m = d.mom[i], pr = d.prop[i];
    float4 p = d.pos[i];
    position_update(p, m, pr.x, pr.y, sub_dt);
    d.pos[i] = p;
    d.frc[i] = make_float4(0.0f, 0.0f, 0.0f, d.frc[i].w);
}

// K12 + K1 for K bodies at once (i0, i0 + stride, ...): exactly integrate_positions_thread followed by
// integrate_forces_thread — same expressions, same order — but with every load of all K bodies issued before the first
// dependent use.  The persistent solver streams a million bodies through 38 k threads: one body per iteration leaves a
// single round trip in flight per thread (433 us per process() on mixed1M); K = 4 keeps 20.
template <int K>
R2D_HD void integrate_batch(const Dev& d, uint32_t i0, uint32_t stride, float sub_dt, bool do_positions, bool do_forces,
                            bool refresh_aabb) {
    float4 sh[K], mo[K], pr[K], po[K], fr[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const uint32_t i = i0 + (uint32_t)k * stride;
        if (i < d.n_bodies) {
            sh[k] = d.shape[i];
            mo[k] = d.mom[i];
            pr[k] = d.prop[i];
            po[k] = d.pos[i];
            fr[k] = d.frc[i];
        }
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const uint32_t i = i0 + (uint32_t)k * stride;
        if (i >= d.n_bodies) continue;
        const uint32_t flags = f2u(sh[k].z);
        const bool is_static = (flags & FLAG_STATIC) != 0;
        float4 p = po[k], m = mo[k], f = fr[k];
        if (do_positions && !is_static) {   // lib.zig:238-249
            position_update(p, m, pr[k].x, pr[k].y, sub_dt);
            d.pos[i] = p;
            f = make_float4(0.0f, 0.0f, 0.0f, f.w);
            d.frc[i] = f;
        }
        if (!do_forces) continue;
        if (refresh_aabb) d.aabb[i] = refreshed_aabb(p, flags, sh[k].x, sh[k].y);   // lib.zig:210
        if (is_static) continue;
        momentum_update(d, flags >> FLAG_WORLD_SHIFT, m, f, pr[k].x, sub_dt);
        d.mom[i] = m;
    }
}

}  // namespace r2d
