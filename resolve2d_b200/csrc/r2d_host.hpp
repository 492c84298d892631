// r2d_host.hpp — host-side runtime shared by the CUDA backend (r2d_runtime.cu): the per-world registry that mirrors
// the reference's `Solver` containers (stable id -> dense slot, insertion order, swapRemove), scene construction
// arithmetic (EntityFactory), joint colouring, and flattening of all worlds of a batch into one structure-of-arrays
// image for upload.  Pure C++17 (no CUDA calls), so tests/emu/ can drive the same code on a CPU.
//
// Follows src/core/lib.zig:13-187,301-315 (EntityFactory, Solver.init/deinit/clear/removeRigidBody),
// Bodies/Disc.zig:29-56, Bodies/Rectangle.zig:33-64.
#pragma once
#include <algorithm>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <string>
#include <thread>
#include <unordered_map>
#include <utility>
#include <vector>

#include "../../include/r2d_abi.h"
#include "r2d_pipeline.cuh"

namespace r2d {
namespace host {

struct Body {
    uint32_t id = 0;
    int32_t shape = 0;
    bool is_static = false;
    float pos_x = 0, pos_y = 0, angle = 0;
    float mom_x = 0, mom_y = 0, ang_mom = 0;
    float force_x = 0, force_y = 0, torque = 0;
    float mass = 0, inertia = 0, mu = 0;
    float aabb_x = 0, aabb_y = 0, aabb_hw = 0, aabb_hh = 0;
    float a = 0, b = 0;  // disc: radius; rect: width, height
    uint32_t sleep_counter = 0;  // R2D_OPT_SLEEPING: consecutive calls below the speed thresholds
};

struct Joint {
    int32_t type = 0;
    uint32_t id1 = 0, id2 = 0;
    float power_max = 0, power_min = 0, beta = 0;
    float r1x = 0, r1y = 0, r2x = 0, r2y = 0;  // offset joint anchors | fixed joint target in (r1x, r1y)
    float target = 0;                          // distance | omega
};

inline bool joint_has_two_bodies(const Joint& j) {
    return j.type == R2D_JOINT_DISTANCE || j.type == R2D_JOINT_OFFSET_DISTANCE;
}

struct BatchBase;

// One world = the containers of the reference's Solver (lib.zig:132-142).
struct World {
    BatchBase* batch = nullptr;
    uint32_t index = 0;
    bool owns_batch = false;
    uint32_t current_body_id = 0;                       // lib.zig:134; survives clear() (Q16)
    std::vector<Body> bodies;                           // AutoArrayHashMap: insertion order, swapRemove
    std::unordered_map<uint32_t, uint32_t> slot_of;     // id -> slot
    std::vector<float> gravities;                       // force_generators (only DownwardsGravity exists)
    std::vector<Joint> joints;                          // constraints, list order = solve order in the reference
    std::set<std::pair<uint32_t, uint32_t>> excluded;   // exclude_collision_pairs, both orders (Q22)

    int find(uint32_t id) const {
        auto it = slot_of.find(id);
        return it == slot_of.end() ? -1 : (int)it->second;
    }

    // makeDiscBody / makeRectangleBody (lib.zig:73-93) + Disc.init / Rectangle.init + appendBody (lib.zig:66-71)
    uint32_t make_body(const r2d_body_opts& o, int shape, float a, float b, bool is_static) {
        Body bd;
        bd.shape = shape;
        bd.a = a;
        bd.b = (shape == R2D_SHAPE_RECT) ? b : 0.0f;
        float mass;
        if (shape == R2D_SHAPE_DISC) {
            const float pi = 3.14159265358979323846f;  // @as(f32, std.math.pi)
            mass = o.mass_is_density ? fmul(fmul(fmul(pi, a), a), o.mass_value) : o.mass_value;
            bd.inertia = fmul(fmul(fmul(0.5f, mass), a), a);                        // Disc.zig:42
        } else {
            mass = o.mass_is_density ? fmul(fmul(a, b), o.mass_value) : o.mass_value;
            bd.inertia = fdiv(fmul(mass, fadd(fmul(a, a), fmul(b, b))), 12.0f);     // Rectangle.zig:51
        }
        bd.mass = mass;
        bd.mu = o.mu;
        bd.pos_x = o.pos_x;
        bd.pos_y = o.pos_y;
        bd.angle = o.angle;
        bd.mom_x = fmul(o.vel_x, mass);       // lib.zig:79 / :90
        bd.mom_y = fmul(o.vel_y, mass);
        bd.ang_mom = fmul(o.omega, bd.inertia);
        bd.is_static = is_static;
        // updateAABB at init (Disc.zig:53, Rectangle.zig:62)
        const uint32_t flags = (shape == R2D_SHAPE_RECT) ? FLAG_RECT : 0u;
        aabb_half_extents(flags, bd.a, bd.b, cos_ref(bd.angle), sin_ref(bd.angle), bd.aabb_hw, bd.aabb_hh);
        bd.aabb_x = bd.pos_x;
        bd.aabb_y = bd.pos_y;
        bd.id = current_body_id;
        slot_of[bd.id] = (uint32_t)bodies.size();
        bodies.push_back(bd);
        current_body_id += 1;
        return bd.id;
    }
    bool remove_body(uint32_t id) {  // lib.zig:308-310: swapRemove — the last body moves into the hole (Q17)
        const int s = find(id);
        if (s < 0) return false;
        slot_of.erase(id);
        if ((size_t)s != bodies.size() - 1) {
            bodies[s] = bodies.back();
            slot_of[bodies[s].id] = (uint32_t)s;
        }
        bodies.pop_back();
        return true;
    }
    void clear() {  // lib.zig:181-187
        bodies.clear();
        slot_of.clear();
        gravities.clear();
        joints.clear();
    }
};

// All worlds of a batch flattened for the device (layout: r2d_pipeline.cuh `Dev`).
struct Image {
    std::vector<float4> pos, mom, frc, prop, shape, aabb;
    std::vector<uint32_t> sleep;   // per device slot: Body::sleep_counter
    std::vector<uint32_t> world_base, grav_off;
    std::vector<float> grav;
    std::vector<uint64_t> excl;
    // joints sorted by (colour, global list index)
    std::vector<uint4> j_hdr;
    std::vector<float4> j_par, j_vec;
    std::vector<uint32_t> joint_color_start;   // n_joint_colors + 1
    std::vector<uint32_t> joint_world, joint_local_index, joint_color;  // per sorted joint
    std::vector<uint32_t> world_joint_start, world_joint;              // the sorted joints of every world (indices), sweep order kept
    // joints inside the dataflow sweep (r2d_pipeline.cuh solve_joint_flow): per sorted joint {rank among the joints of its
    // first body, that body's joint count, same for the second body}; per device slot the number of joints naming the body.
    // `joints_flow_ok`: no two-body joint names a static body or the same body twice (those write a static body's momentum,
    // Q10, which the contacts read without waiting: they keep the barrier sweep).
    std::vector<uint4> j_dep;
    std::vector<uint32_t> body_nj;
    bool joints_flow_ok = false;
    uint32_t n_bodies = 0;
    // Fine broadphase grid (r2d_pipeline.cuh, "fine grid"): cell width chosen from the body sizes of this upload (0 = no
    // fine grid), the coarse cell it was chosen for, and how many bodies are too wide for it (FLAG_LARGE).
    float fine_cell = 0.0f, fine_for_cell = 0.0f;
    uint32_t n_large = 0, n_large_dynamic = 0;
    // Device slots are a spatial (Morton) permutation of the host's insertion-order slots inside every world, so that
    // bodies that touch are neighbours in memory (the gathers of the colouring and of the contact sweep coalesce).
    // Nothing observable depends on it: pairs, colours and the sweep order are functions of ids and geometry only.
    std::vector<uint32_t> dev_of_host;   // host global slot (world-major, insertion order) -> device slot
    std::vector<uint32_t> host_of_dev;   // device slot -> host global slot
    uint32_t dev_id(uint32_t dev) const { return f2u(shape[dev].w); }
    uint32_t dev_world(uint32_t dev) const { return f2u(shape[dev].z) >> FLAG_WORLD_SHIFT; }
};

inline float4 mkf4(float x, float y, float z, float w) { return make_float4(x, y, z, w); }

inline void body_to_image(const Body& b, uint32_t world, Image& im, size_t at, bool large) {
    const uint32_t flags = (b.is_static ? FLAG_STATIC : 0u) | (b.shape == R2D_SHAPE_RECT ? FLAG_RECT : 0u) |
                           (large ? FLAG_LARGE : 0u) | (world << FLAG_WORLD_SHIFT);
    im.pos[at] = mkf4(b.pos_x, b.pos_y, b.angle, 0.0f);
    im.mom[at] = mkf4(b.mom_x, b.mom_y, b.ang_mom, 0.0f);
    im.frc[at] = mkf4(b.force_x, b.force_y, b.torque, 0.0f);
    im.prop[at] = mkf4(b.mass, b.inertia, b.mu, 0.0f);
    im.shape[at] = mkf4(b.a, b.b, u2f(flags), u2f(b.id));
    im.aabb[at] = mkf4(b.aabb_x, b.aabb_y, b.aabb_hw, b.aabb_hh);
    im.sleep[at] = b.sleep_counter;
}

// Returns R2D_OK or R2D_ERR_INVALID_BODY_ID (a joint names a body that no longer exists — the reference fails inside
// process() with error.InvalidRigidBodyId, DistanceJoint.zig:44-46; here before any state is touched).
// Width bound of a body's AABB at any angle: disc 2 r, rectangle its diagonal; never below the stored AABB.
inline float body_width_bound(const Body& b) {
    float w = (b.shape == R2D_SHAPE_RECT) ? (float)std::sqrt((double)b.a * b.a + (double)b.b * b.b) : 2.0f * b.a;
    w = std::max(w, std::max(2.0f * b.aabb_hw, 2.0f * b.aabb_hh));
    return w == w ? w : INFINITY;
}
// Fine broadphase cell: the widest body that is at most 4 x the median width, plus a margin that covers
// AABB_EPS_OVERLAP and float rounding, capped below the coarse cell (a small body then covers <= 2 x 2 coarse cells).
constexpr float FINE_MARGIN = 0.05f;
inline void choose_fine_cell(const std::vector<std::unique_ptr<World>>& worlds, float coarse_cell, Image& im) {
    im.fine_cell = 0.0f;
    im.fine_for_cell = coarse_cell;
    std::vector<float> w;
    for (auto& W : worlds)
        for (const Body& b : W->bodies) w.push_back(body_width_bound(b));
    if (w.empty()) return;
    std::nth_element(w.begin(), w.begin() + w.size() / 2, w.end());
    const float median = w[w.size() / 2];
    const float cap = std::min(3.95f, coarse_cell - FINE_MARGIN);
    const float limit = std::min(4.0f * median, cap - FINE_MARGIN);
    float widest = 0.0f;
    for (float x : w)
        if (x <= limit) widest = std::max(widest, x);
    if (!(widest > 0.0f) || !(limit > 0.0f)) return;
    im.fine_cell = widest + FINE_MARGIN;
}
inline bool body_is_large(const Body& b, const Image& im) {
    return !(im.fine_cell > 0.0f && body_width_bound(b) + FINE_MARGIN <= im.fine_cell);
}

inline int build_slot_tables(const std::vector<std::unique_ptr<World>>& worlds, Image& im);
inline int build_image(const std::vector<std::unique_ptr<World>>& worlds, Image& im, float coarse_cell) {
    const size_t nw = worlds.size();
    choose_fine_cell(worlds, coarse_cell, im);
    im.n_large = im.n_large_dynamic = 0;
    im.world_base.assign(nw + 1, 0);
    im.grav_off.assign(nw + 1, 0);
    im.grav.clear();
    for (size_t w = 0; w < nw; ++w) {
        im.world_base[w + 1] = im.world_base[w] + (uint32_t)worlds[w]->bodies.size();
        for (float g : worlds[w]->gravities) im.grav.push_back(g);
        im.grav_off[w + 1] = (uint32_t)im.grav.size();
    }
    const size_t nb = im.world_base[nw];
    im.n_bodies = (uint32_t)nb;
    im.dev_of_host.resize(nb);
    im.host_of_dev.resize(nb);
    {
        std::vector<std::pair<uint64_t, uint32_t>> keyed;
        for (size_t w = 0; w < nw; ++w) {
            const World& W = *worlds[w];
            const uint32_t base = im.world_base[w];
            const size_t n = W.bodies.size();
            float minx = 0, miny = 0;
            bool any = false;
            for (const Body& b : W.bodies) {
                if (!(b.pos_x == b.pos_x) || !(b.pos_y == b.pos_y)) continue;
                if (!any || b.pos_x < minx) minx = b.pos_x;
                if (!any || b.pos_y < miny) miny = b.pos_y;
                any = true;
            }
            keyed.resize(n);
            for (size_t k = 0; k < n; ++k) {
                const Body& b = W.bodies[k];
                keyed[k] = {resort_key(b.pos_x, b.pos_y, minx, miny), (uint32_t)k};   // (shared with the device re-sort)
            }
            // stable LSD radix sort on the 32-bit Morton key (two 16-bit passes): the re-sort runs every reorder interval on
            // the host, std::stable_sort was most of its cost at 100 k bodies
            if (n > 2048) {
                std::vector<std::pair<uint64_t, uint32_t>> tmp(n);
                for (int pass = 0; pass < 2; ++pass) {
                    const int shift = 16 * pass;
                    std::vector<uint32_t> cnt(65537, 0u);
                    for (size_t k = 0; k < n; ++k) cnt[((keyed[k].first >> shift) & 0xFFFFu) + 1] += 1;
                    for (size_t b = 0; b < 65536; ++b) cnt[b + 1] += cnt[b];
                    for (size_t k = 0; k < n; ++k) tmp[cnt[(keyed[k].first >> shift) & 0xFFFFu]++] = keyed[k];
                    keyed.swap(tmp);
                }
            } else {
                std::stable_sort(keyed.begin(), keyed.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
            }
            for (size_t k = 0; k < n; ++k) {
                im.host_of_dev[base + k] = base + keyed[k].second;
                im.dev_of_host[base + keyed[k].second] = base + (uint32_t)k;
            }
        }
    }
    im.pos.resize(nb);
    im.mom.resize(nb);
    im.frc.resize(nb);
    im.prop.resize(nb);
    im.shape.resize(nb);
    im.aabb.resize(nb);
    im.sleep.resize(nb);
    for (size_t w = 0; w < nw; ++w) {
        const World& W = *worlds[w];
        const uint32_t base = im.world_base[w];
        for (size_t k = 0; k < W.bodies.size(); ++k) {   // in DEVICE order: six sequential streams, one gather per body
            const Body& bd = W.bodies[im.host_of_dev[base + k] - base];
            const bool large = body_is_large(bd, im);
            if (large) {
                im.n_large += 1;
                if (!bd.is_static) im.n_large_dynamic += 1;
            }
            body_to_image(bd, (uint32_t)w, im, base + k, large);
        }
    }
    return build_slot_tables(worlds, im);
}

// Everything in the image that names bodies by DEVICE slot and is not a per-body array: the exclusion list and the joints
// (slots, colours, dataflow ranks).  Rebuilt alone after a device-side re-sort (BatchBase::reorder), which permutes the
// body arrays on the device and hands back the new order.
inline int build_slot_tables(const std::vector<std::unique_ptr<World>>& worlds, Image& im) {
    const size_t nw = worlds.size();
    const size_t nb = im.n_bodies;
    im.excl.clear();
    struct GJ {
        Joint j;
        uint32_t world, local, s1, s2, color;
        bool st1, st2;
    };
    std::vector<GJ> gj;
    for (size_t w = 0; w < nw; ++w) {
        const World& W = *worlds[w];
        const uint32_t base = im.world_base[w];
        for (const auto& pr : W.excluded) {
            const int s1 = W.find(pr.first), s2 = W.find(pr.second);
            if (s1 < 0 || s2 < 0 || s1 == s2) continue;
            const uint32_t d1 = im.dev_of_host[base + (uint32_t)s1], d2 = im.dev_of_host[base + (uint32_t)s2];
            const uint64_t lo = std::min(d1, d2), hi = std::max(d1, d2);
            im.excl.push_back((lo << 32) | hi);
        }
        for (size_t k = 0; k < W.joints.size(); ++k) {
            const Joint& j = W.joints[k];
            const int s1 = W.find(j.id1);
            const int s2 = joint_has_two_bodies(j) ? W.find(j.id2) : s1;
            if (s1 < 0 || s2 < 0) return R2D_ERR_INVALID_BODY_ID;
            gj.push_back({j, (uint32_t)w, (uint32_t)k, im.dev_of_host[base + (uint32_t)s1], im.dev_of_host[base + (uint32_t)s2], 0u,
                          W.bodies[s1].is_static, W.bodies[s2].is_static});
        }
    }
    std::sort(im.excl.begin(), im.excl.end());
    im.excl.erase(std::unique(im.excl.begin(), im.excl.end()), im.excl.end());
    // Joint colouring: greedy in list order; two joints conflict when they name a common body (static or not — the
    // distance joints write momentum into static bodies too, Q10).  Sweep order = (colour, list index).
    {
        std::unordered_map<uint32_t, std::vector<uint32_t>> used;
        uint32_t n_colors = 0;
        for (GJ& g : gj) {
            uint32_t c = 0;
            for (;; ++c) {
                bool taken = false;
                for (uint32_t s : {g.s1, g.s2}) {
                    auto& u = used[s];
                    if (std::find(u.begin(), u.end(), c) != u.end()) taken = true;
                }
                if (!taken) break;
            }
            g.color = c;
            used[g.s1].push_back(c);
            if (g.s2 != g.s1) used[g.s2].push_back(c);
            n_colors = std::max(n_colors, c + 1);
        }
        std::stable_sort(gj.begin(), gj.end(), [](const GJ& a, const GJ& b) { return a.color < b.color; });
        const size_t nj = gj.size();
        im.j_hdr.resize(nj);
        im.j_par.resize(nj);
        im.j_vec.resize(nj);
        im.joint_world.resize(nj);
        im.joint_local_index.resize(nj);
        im.joint_color.resize(nj);
        im.joint_color_start.assign(n_colors + 1, 0);
        for (size_t k = 0; k < nj; ++k) {
            const GJ& g = gj[k];
            im.j_hdr[k] = make_uint4((uint32_t)g.j.type, g.s1, g.s2, g.color);
            im.j_par[k] = mkf4(g.j.power_max, g.j.power_min, g.j.beta, g.j.target);
            im.j_vec[k] = mkf4(g.j.r1x, g.j.r1y, g.j.r2x, g.j.r2y);
            im.joint_world[k] = g.world;
            im.joint_local_index[k] = g.local;
            im.joint_color[k] = g.color;
            im.joint_color_start[g.color + 1] += 1;
        }
        for (size_t c = 0; c < n_colors; ++c) im.joint_color_start[c + 1] += im.joint_color_start[c];
        // per world, in sweep order (the sorted order is (colour, list index): a stable counting sort by world keeps it)
        im.world_joint_start.assign(nw + 1, 0u);
        for (size_t k = 0; k < nj; ++k) im.world_joint_start[gj[k].world + 1] += 1;
        for (size_t w = 0; w < nw; ++w) im.world_joint_start[w + 1] += im.world_joint_start[w];
        im.world_joint.resize(nj);
        {
            std::vector<uint32_t> cur(im.world_joint_start.begin(), im.world_joint_start.end() - 1);
            for (size_t k = 0; k < nj; ++k) im.world_joint[cur[gj[k].world]++] = (uint32_t)k;
        }
        im.body_nj.assign(nb, 0u);
        im.j_dep.assign(nj, make_uint4(0u, 0u, 0u, 0u));
        im.joints_flow_ok = nj > 0;
        for (size_t k = 0; k < nj; ++k) {   // sweep order = sorted order: ranks are counted on the way
            const GJ& g = gj[k];
            const bool two = joint_has_two_bodies(g.j);
            im.j_dep[k].x = im.body_nj[g.s1]++;
            if (two) {
                if (g.s2 == g.s1 || g.st1 || g.st2) im.joints_flow_ok = false;
                if (g.s2 != g.s1) im.j_dep[k].z = im.body_nj[g.s2]++;
            }
        }
        for (size_t k = 0; k < nj; ++k) {
            im.j_dep[k].y = im.body_nj[gj[k].s1];
            im.j_dep[k].w = im.body_nj[gj[k].s2];
        }
    }
    return R2D_OK;
}

// R2D_MODE_REFERENCE_ORDER: the order in which the reference's updateManifolds (lib.zig:262-297) creates the manifolds of
// one world, replayed on the host from the AABBs the broadphase saw: SpatialHash.init (SpatialHash.zig:19-71: count,
// inclusive prefix, decrement-then-store fill, so a bucket lists its entries in REVERSE insertion order), then for every
// body in iteration order its query (:108-137: cells row-major, y outer; bucket contents appended, duplicates kept).  A
// pair becomes a manifold the first time the walk meets it.  `aabb[k]` = stored AABB (centre, half extents) of the k-th
// body in ITERATION order; `manifold_of(lo, hi)` = the manifold's index for that unordered pair of iteration indices, or
// -1.  Returns the manifold indices in creation order.
template <class Lookup>
inline std::vector<uint32_t> reference_manifold_order(const std::vector<float4>& aabb, float cell, uint32_t table_size,
                                                      size_t n_manifolds, Lookup manifold_of) {
    const size_t n = aabb.size();
    std::vector<uint32_t> order;
    order.reserve(n_manifolds);
    if (n == 0 || table_size == 0) return order;
    struct Range { int64_t x0, y0, x1, y1; };
    std::vector<Range> rg(n);
    for (size_t i = 0; i < n; ++i) {   // iterateAABBHashes :83-96 (getVertices: min = pos - half, max = pos + half)
        const float4 a = aabb[i];
        rg[i] = {cell_coord(fsub(a.x, a.z), cell), cell_coord(fsub(a.y, a.w), cell), cell_coord(fadd(a.x, a.z), cell),
                 cell_coord(fadd(a.y, a.w), cell)};
        if (rg[i].x1 - rg[i].x0 > MAX_BODY_CELLS || rg[i].y1 - rg[i].y0 > MAX_BODY_CELLS) rg[i] = {0, 0, -1, -1};   // NaN / inf pose
    }
    std::vector<uint32_t> table(table_size + 1, 0u);
    for (size_t i = 0; i < n; ++i)
        for (int64_t y = rg[i].y0; y <= rg[i].y1; ++y)
            for (int64_t x = rg[i].x0; x <= rg[i].x1; ++x) table[cell_hash(x, y, table_size)] += 1;
    uint32_t start = 0;
    for (uint32_t h = 0; h < table_size; ++h) {
        start += table[h];
        table[h] = start;
    }
    table[table_size] = start;
    std::vector<uint32_t> entries(start);
    for (size_t i = 0; i < n; ++i)
        for (int64_t y = rg[i].y0; y <= rg[i].y1; ++y)
            for (int64_t x = rg[i].x0; x <= rg[i].x1; ++x) entries[--table[cell_hash(x, y, table_size)]] = (uint32_t)i;
    std::vector<unsigned char> seen(n_manifolds, 0);
    for (size_t i = 0; i < n; ++i)
        for (int64_t y = rg[i].y0; y <= rg[i].y1; ++y)
            for (int64_t x = rg[i].x0; x <= rg[i].x1; ++x) {
                const uint64_t h = cell_hash(x, y, table_size);
                for (uint32_t e = table[h]; e < table[h + 1]; ++e) {
                    const uint32_t j = entries[e];
                    if (j == i) continue;
                    const long m = manifold_of(std::min<uint32_t>((uint32_t)i, j), std::max<uint32_t>((uint32_t)i, j));
                    if (m < 0 || seen[m]) continue;
                    seen[m] = 1;
                    order.push_back((uint32_t)m);
                }
            }
    return order;
}

struct RawManifold {   // one occupied pair slot of the last process(), slots are global
    uint32_t ref, inc, normal_id, n_points, color;
    float normal_x, normal_y;
    float pos_x[2], pos_y[2], depth[2], ref_rx[2], ref_ry[2], inc_rx[2], inc_ry[2];
};

enum BodyField { FIELD_POS = 0, FIELD_MOM = 1, FIELD_FRC = 2 };  // float4 arrays a setter may touch

// Backend-independent part of a batch: worlds + coherence protocol between the host mirror and the device arrays.
struct BatchBase {
    std::vector<std::unique_ptr<World>> worlds;
    float cell_width = 2.0f;
    uint32_t table_mult = 4;
    int mode = R2D_MODE_PARITY;
    // roadmap options (README.md:59-64), off by default: see include/r2d_abi.h R2D_OPT_*
    bool opt_warm_start = false, opt_sleeping = false;
    uint32_t opt_sleep_calls = 30;
    bool host_fresh = true;   // host mirror holds the truth
    bool dev_fresh = false;   // device arrays hold the truth (and the image layout is current)
    Image image;              // layout of the last upload (world_base etc.)
    r2d_step_stats stats{};

    virtual ~BatchBase() {}
    virtual int backend_upload() = 0;                                      // image -> device
    virtual int backend_download() = 0;                                    // device -> worlds[*].bodies
    virtual int backend_write(uint32_t gslot, BodyField f, int comp, int n, const float* v) = 0;
    virtual int backend_process(float dt, uint32_t sub_steps, uint32_t iters) = 0;
    // bulk access to device-resident state (valid while dev_fresh); [first, first + n) are global slots
    virtual int backend_read_bodies(uint32_t first, uint32_t n, uint32_t* ids, float* pos_xy, float* angle,
                                    float* momentum_xy, float* ang_momentum, float* aabb_xywh) = 0;
    virtual int backend_write_forces(uint32_t first, uint32_t n, const float* force_xy_torque) = 0;
    virtual int backend_read_pairs(std::vector<uint2>& slots) = 0;        // (owner, other) global slots, last process()
    virtual int backend_read_manifolds(std::vector<RawManifold>& out) = 0; // pair-slot order, last process()
    virtual int backend_sync() = 0;
    virtual int backend_set_stream(void* stream) = 0;
    virtual int backend_profile_enable(int on) = 0;
    virtual int backend_profile_read(double* ms, uint64_t* launches, int reset) = 0;
    // does the periodic spatial re-sort pay for this batch?  (not for batches of small worlds: their kernels keep a whole
    // world in shared memory, the order of its bodies in HBM is irrelevant, and the re-sort goes through the host)
    virtual bool backend_reorder_pays() const { return true; }
    // re-sort on the device (keys, radix sort and permutation of the body arrays there; the host only rebuilds the tables
    // that name device slots).  A backend without one answers REORDER_ON_HOST and reorder() goes through the host.
    static constexpr int REORDER_ON_HOST = -1000;
    virtual int backend_reorder() { return REORDER_ON_HOST; }

    float grid_cell() const { return mode != R2D_MODE_FAST ? 4.0f : cell_width; }      // lib.zig:254-255 (Q2)
    uint32_t grid_mult() const { return mode != R2D_MODE_FAST ? 2u : (table_mult ? table_mult : 1u); }

    int ensure_host() {
        if (!host_fresh) {
            const int st = backend_download();
            if (st != R2D_OK) return st;
            host_fresh = true;
        }
        return R2D_OK;
    }
    int touch_structure() {  // before any edit the device image cannot absorb in place
        const int st = ensure_host();
        if (st != R2D_OK) return st;
        dev_fresh = false;
        return R2D_OK;
    }
    int ensure_device() {
        if (!dev_fresh) {
            const int st = build_image(worlds, image, grid_cell());
            if (st != R2D_OK) return st;
            const int st2 = backend_upload();
            if (st2 != R2D_OK) return st2;
            dev_fresh = true;
        }
        return R2D_OK;
    }
    size_t total_bodies() const {
        size_t n = 0;
        for (auto& w : worlds) n += w->bodies.size();
        return n;
    }
    // scalar setter: edits whichever copies are fresh
    int set_field(World* w, uint32_t id, BodyField f, int comp, int n, const float* v) {
        const int s = w->find(id);
        if (s < 0) return R2D_ERR_NO_SUCH_ID;
        if (host_fresh) {
            Body& b = w->bodies[s];
            float* dst[3][3] = {{&b.pos_x, &b.pos_y, &b.angle}, {&b.mom_x, &b.mom_y, &b.ang_mom}, {&b.force_x, &b.force_y, &b.torque}};
            for (int k = 0; k < n; ++k) *dst[f][comp + k] = v[k];
        }
        if (dev_fresh) return backend_write(image.dev_of_host[image.world_base[w->index] + (uint32_t)s], f, comp, n, v);
        return R2D_OK;
    }
    // process() + bulk read of every body in ONE call (r2d_process_read): a backend may enqueue the export behind the
    // step's last kernel, so that the call has a single host synchronisation (sets readback_done); otherwise the
    // generic path reads after the step.
    struct Readback {
        uint32_t* ids = nullptr;
        float *pos_xy = nullptr, *angle = nullptr, *momentum_xy = nullptr, *ang_momentum = nullptr, *aabb_xywh = nullptr;
    };
    const Readback* readback = nullptr;
    bool readback_done = false;
    // process() calls between spatial re-sorts of the device order (0 = never); r2d_reorder() forces one.  With the state
    // resident on the device the re-sort happens there (backend_reorder: ~0.4 ms for 100 k bodies); a backend without that
    // path, a stale device image, or a changed fine cell go through the host (download, sort, upload: 5.6 ms for 100 k).
    uint32_t reorder_interval = 1024;
    uint32_t steps_since_upload = 0;
    // A backend re-sorted its body arrays itself: `host_of_new[j]` = host slot of the body now at device slot j.  Updates
    // the two maps, the shape image (ids and world indices by device slot) and everything that names bodies by device slot;
    // the backend then uploads the slot tables (exclusions, joints).
    int adopt_device_order(const std::vector<uint32_t>& host_of_new) {
        const size_t nb = image.n_bodies;
        if (host_of_new.size() != nb) return R2D_ERR_BAD_STATE;
        std::vector<float4> sh(nb);
        for (size_t j = 0; j < nb; ++j) sh[j] = image.shape[image.dev_of_host[host_of_new[j]]];
        image.shape.swap(sh);
        for (size_t j = 0; j < nb; ++j) {
            image.host_of_dev[j] = host_of_new[j];
            image.dev_of_host[host_of_new[j]] = (uint32_t)j;
        }
        return build_slot_tables(worlds, image);
    }
    int reorder_through_host() {
        const int st = ensure_host();
        if (st != R2D_OK) return st;
        dev_fresh = false;
        return R2D_OK;
    }
    int reorder() {  // re-derive the device order from the current positions
        if (dev_fresh && image.fine_for_cell == grid_cell()) {
            const int st = backend_reorder();
            if (st == R2D_OK) steps_since_upload = 0;
            if (st != REORDER_ON_HOST) return st;
        }
        return reorder_through_host();
    }
    bool poisoned = false;
    int process(float dt, uint32_t sub_steps, uint32_t iters, const Readback* rb = nullptr) {
        if (poisoned) return R2D_ERR_BAD_STATE;
        if (dev_fresh && reorder_interval && steps_since_upload >= reorder_interval && backend_reorder_pays()) {
            const int sr = reorder();
            if (sr != R2D_OK) return sr;
        }
        if (dev_fresh && image.fine_for_cell != grid_cell()) {  // the mode changed: the fine cell depends on the coarse one
            const int sr = reorder_through_host();
            if (sr != R2D_OK) return sr;
        }
        if (!dev_fresh) steps_since_upload = 0;
        steps_since_upload += 1;
        const int st = ensure_device();
        if (st != R2D_OK) return st;
        readback = rb;
        readback_done = false;
        const int st2 = backend_process(dt, sub_steps, iters);
        readback = nullptr;
        if (st2 == R2D_OK || st2 == R2D_ERR_COLOR_OVERFLOW) {
            host_fresh = false;
        } else if (st2 == R2D_ERR_CUDA) {
            // a stall or a CUDA failure: the state of the step is undefined on the device (every other status is raised
            // before a kernel has modified a body).  Re-upload from the host mirror if that is still the truth ...
            if (host_fresh)
                dev_fresh = false;
            else
                poisoned = true;   // ... else no good copy is left: process() reports R2D_ERR_BAD_STATE until r2d_clear()
        }
        if (st2 == R2D_OK && rb && !readback_done && image.n_bodies)
            return backend_read_bodies(0, image.n_bodies, rb->ids, rb->pos_xy, rb->angle, rb->momentum_xy, rb->ang_momentum, rb->aabb_xywh);
        return st2;
    }
};

}  // namespace host
}  // namespace r2d
