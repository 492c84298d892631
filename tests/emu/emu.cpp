// emu.cpp — TEST TOOLING, NOT PRODUCT CODE.
//
// Serial CPU driver for the per-thread kernel bodies of resolve2d_b200/csrc/r2d_pipeline.cuh: every `*_thread`
// function the CUDA kernels execute is run here in a plain loop, in the same pipeline order as r2d_runtime.cu, behind
// the same C API (prefix emu_).  Its only purpose is to let the CPU-only test tier (-m "not gpu") check the kernels'
// index / filter / colouring logic and the host registry against the oracle without a GPU.  Nothing under
// resolve2d_b200/ links or loads it; the product path has no CPU fallback.
#include <cstdio>
#include <cstring>
#include <numeric>

#include "../../resolve2d_b200/csrc/r2d_host.hpp"

namespace {

using namespace r2d;
using host::BatchBase;
using host::BodyField;
using host::RawManifold;

struct EmuBatch : BatchBase {
    Dev d{};
    std::vector<float4> pos, mom, frc, prop, shape, aabb, pose, view;
    std::vector<uint32_t> ncells, bucket_cnt, bucket_start, ent_body, ent_key, ent_off, m_color, pair_cnt;
    std::vector<int4> fcell;
    std::vector<float4> ent_aabb;
    std::vector<uint2> pairs;
    std::vector<uint4> m_hdr, s_hdr, bkt;
    std::vector<float4> m_g0, m_g1, m_r0, m_r1, s_nf, s_inv, s_r0, s_r1, s_pm0, s_pm1;
    std::vector<float2> s_acc0, s_acc1;
    std::vector<uint4> s_dep;
    std::vector<unsigned long long> maxprio0, maxprio1, used, m_prio, adj_prio;
    std::vector<uint32_t> adj_cnt, adj_head;
    std::vector<uint4> adj_pool;
    std::vector<uint4> cstate;
    std::vector<uint32_t> color_count, color_start, color_cursor, round_left, own_bits, own_pos;
    std::vector<uint64_t> world_magic;
    Counters counters{};
    uint32_t n_pairs_last = 0;

    int backend_upload() override {
        pos = image.pos;
        mom = image.mom;
        frc = image.frc;
        prop = image.prop;
        shape = image.shape;
        aabb = image.aabb;
        n_pairs_last = 0;
        return R2D_OK;
    }
    // the device-side re-sort (k_resort_min / k_resort_keys / radix sort / k_resort_gather of r2d_kernels.cuh), serially:
    // same key function, same stable order, then the host bookkeeping shared with the CUDA backend (adopt_device_order)
    int backend_reorder() override {
        if (getenv("R2D_EMU_RESORT") && atoi(getenv("R2D_EMU_RESORT")) == 0) return REORDER_ON_HOST;
        if (getenv("R2D_EMU_RESORT") && atoi(getenv("R2D_EMU_RESORT")) == 2) return R2D_ERR_BAD_STATE;   // (test: is this path taken?)
        const uint32_t nb = image.n_bodies, nw = (uint32_t)worlds.size();
        if (nb == 0) return REORDER_ON_HOST;
        std::vector<int> wmx(nw, RESORT_NO_MIN), wmy(nw, RESORT_NO_MIN);
        for (uint32_t i = 0; i < nb; ++i) {
            const float4 p = pos[i];
            if (!(p.x == p.x) || !(p.y == p.y)) continue;
            const uint32_t w = f2u(shape[i].z) >> FLAG_WORLD_SHIFT;
            wmx[w] = std::min(wmx[w], resort_float_order(p.x));
            wmy[w] = std::min(wmy[w], resort_float_order(p.y));
        }
        std::vector<std::pair<unsigned long long, uint32_t>> keyed(nb);
        for (uint32_t h = 0; h < nb; ++h) {
            const uint32_t s = image.dev_of_host[h];
            const uint32_t w = f2u(shape[s].z) >> FLAG_WORLD_SHIFT;
            const float mx = wmx[w] == RESORT_NO_MIN ? 0.0f : resort_float_unorder(wmx[w]);
            const float my = wmx[w] == RESORT_NO_MIN ? 0.0f : resort_float_unorder(wmy[w]);
            keyed[h] = {((unsigned long long)w << 32) | resort_key(pos[s].x, pos[s].y, mx, my), h};
        }
        std::stable_sort(keyed.begin(), keyed.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
        std::vector<uint32_t> order(nb);
        std::vector<float4> p2(nb), m2(nb), f2(nb), pr2(nb), s2(nb), a2(nb);
        for (uint32_t j = 0; j < nb; ++j) {
            order[j] = keyed[j].second;
            const uint32_t o = image.dev_of_host[order[j]];
            p2[j] = pos[o]; m2[j] = mom[o]; f2[j] = frc[o]; pr2[j] = prop[o]; s2[j] = shape[o]; a2[j] = aabb[o];
        }
        pos.swap(p2); mom.swap(m2); frc.swap(f2); prop.swap(pr2); shape.swap(s2); aabb.swap(a2);
        const int st = adopt_device_order(order);
        if (st != R2D_OK) return st;
        n_pairs_last = 0;
        // (always checked here: the order must be the one build_image derives from the same state)
        const int sd = backend_download();
        if (sd != R2D_OK) return sd;
        host::Image ref;
        const int sb = host::build_image(worlds, ref, grid_cell());
        if (sb != R2D_OK) return sb;
        return ref.host_of_dev == image.host_of_dev ? R2D_OK : R2D_ERR_BAD_STATE;
    }
    int backend_download() override {
        for (auto& w : worlds) {
            const uint32_t base = image.world_base[w->index];
            for (size_t s = 0; s < w->bodies.size(); ++s) {
                host::Body& b = w->bodies[s];
                const uint32_t ds = image.dev_of_host[base + s];
                const float4 p = pos[ds], m = mom[ds], f = frc[ds], a = aabb[ds];
                b.pos_x = p.x; b.pos_y = p.y; b.angle = p.z;
                b.mom_x = m.x; b.mom_y = m.y; b.ang_mom = m.z;
                b.force_x = f.x; b.force_y = f.y; b.torque = f.z;
                b.aabb_x = a.x; b.aabb_y = a.y; b.aabb_hw = a.z; b.aabb_hh = a.w;
            }
        }
        return R2D_OK;
    }
    int backend_write(uint32_t gslot, BodyField f, int comp, int n, const float* v) override {
        float4* arr = f == host::FIELD_POS ? pos.data() : (f == host::FIELD_MOM ? mom.data() : frc.data());
        float* p = &arr[gslot].x;
        for (int k = 0; k < n; ++k) p[comp + k] = v[k];
        return R2D_OK;
    }
    int backend_read_bodies(uint32_t first, uint32_t n, uint32_t* ids, float* pos_xy, float* angle, float* momentum_xy,
                            float* ang_momentum, float* aabb_xywh) override {
        for (uint32_t k = 0; k < n; ++k) {
            const uint32_t s = image.dev_of_host[first + k];
            if (ids) ids[k] = f2u(shape[s].w);
            if (pos_xy) { pos_xy[2 * k] = pos[s].x; pos_xy[2 * k + 1] = pos[s].y; }
            if (angle) angle[k] = pos[s].z;
            if (momentum_xy) { momentum_xy[2 * k] = mom[s].x; momentum_xy[2 * k + 1] = mom[s].y; }
            if (ang_momentum) ang_momentum[k] = mom[s].z;
            if (aabb_xywh) memcpy(aabb_xywh + 4 * k, &aabb[s], 16);
        }
        return R2D_OK;
    }
    int backend_write_forces(uint32_t first, uint32_t n, const float* f) override {
        for (uint32_t k = 0; k < n; ++k) {
            const uint32_t s = image.dev_of_host[first + k];
            frc[s].x = f[3 * k];
            frc[s].y = f[3 * k + 1];
            frc[s].z = f[3 * k + 2];
        }
        return R2D_OK;
    }
    int backend_read_pairs(std::vector<uint2>& out) override {
        out.assign(pairs.begin(), pairs.begin() + n_pairs_last);
        return R2D_OK;
    }
    int backend_read_manifolds(std::vector<RawManifold>& out) override {
        out.clear();
        for (uint32_t p = 0; p < n_pairs_last; ++p) {
            if (m_color[p] == COLOR_NONE) continue;
            RawManifold m{};
            m.ref = m_hdr[p].x;
            m.inc = m_hdr[p].y;
            m.n_points = m_hdr[p].z & 0xFF;
            m.normal_id = m_hdr[p].z >> 8;
            m.color = m_color[p];
            m.normal_x = m_g0[p].x;
            m.normal_y = m_g0[p].y;
            m.pos_x[0] = m_g0[p].z; m.pos_y[0] = m_g0[p].w;
            m.depth[0] = m_g1[p].x; m.depth[1] = m_g1[p].y;
            m.pos_x[1] = m_g1[p].z; m.pos_y[1] = m_g1[p].w;
            m.ref_rx[0] = m_r0[p].x; m.ref_ry[0] = m_r0[p].y; m.inc_rx[0] = m_r0[p].z; m.inc_ry[0] = m_r0[p].w;
            m.ref_rx[1] = m_r1[p].x; m.ref_ry[1] = m_r1[p].y; m.inc_rx[1] = m_r1[p].z; m.inc_ry[1] = m_r1[p].w;
            out.push_back(m);
        }
        return R2D_OK;
    }
    int backend_sync() override { return R2D_OK; }
    int backend_set_stream(void*) override { return R2D_OK; }
    int backend_profile_enable(int) override { return R2D_OK; }
    int backend_profile_read(double* ms, uint64_t* n, int) override {
        for (int k = 0; k < R2D_KCLASS_COUNT; ++k) { ms[k] = 0; n[k] = 0; }
        return R2D_OK;
    }

    static void exclusive_scan(uint32_t* a, uint32_t n) {  // a[0..n) -> exclusive prefix, a[n] = total
        uint32_t run = 0;
        for (uint32_t i = 0; i < n; ++i) {
            const uint32_t v = a[i];
            a[i] = run;
            run += v;
        }
        a[n] = run;
    }

    int backend_process(float dt, uint32_t S, uint32_t I) override {
        const uint32_t nb = image.n_bodies;
        stats = r2d_step_stats{};
        stats.n_bodies = nb;
        stats.n_joints = (uint32_t)image.j_hdr.size();
        stats.n_joint_colors = (uint32_t)image.joint_color_start.size() - 1;
        if (nb == 0) return R2D_OK;
        const float sub_dt = fdiv(dt, (float)S);   // lib.zig:190-191
        d.n_bodies = nb;
        d.color_smem = 0;
        d.sub_dt = sub_dt;
        d.pos = pos.data(); d.mom = mom.data(); d.frc = frc.data(); d.prop = prop.data(); d.shape = shape.data();
        d.aabb = aabb.data();
        pose.resize(nb); view.resize(4 * (size_t)nb); ncells.resize(nb); bkt.resize(nb);
        d.pose = pose.data(); d.view = view.data(); d.ncells = ncells.data(); d.bkt = bkt.data();
        d.n_worlds = (uint32_t)worlds.size();
        d.world_base = image.world_base.data(); d.grav_off = image.grav_off.data(); d.grav = image.grav.data();
        world_magic.resize(worlds.size());
        for (size_t w = 0; w < worlds.size(); ++w)
            world_magic[w] = hash_magic((uint64_t)grid_mult() * (image.world_base[w + 1] - image.world_base[w]));
        d.world_magic = world_magic.data();
        d.cell = grid_cell(); d.table_mult = grid_mult();
        d.n_buckets = d.table_mult * nb;
        d.excl = image.excl.data(); d.n_excl = (uint32_t)image.excl.size();
        counters = Counters{};
        d.counters = &counters;
        d.n_joints = (uint32_t)image.j_hdr.size();
        d.j_hdr = image.j_hdr.data(); d.j_par = image.j_par.data(); d.j_vec = image.j_vec.data();

        // ---- broadphase (K2..K5) ----
        // R2D_EMU_BROADPHASE=buckets emulates the original pipeline (every body through the coarse buckets)
        const bool want_fine = !(getenv("R2D_EMU_BROADPHASE") && std::string(getenv("R2D_EMU_BROADPHASE")) == "buckets");
        const bool fine = want_fine && image.fine_cell > 0.0f && image.fine_for_cell == grid_cell() &&
                          (image.n_large_dynamic == 0 || worlds.size() == 1);
        const bool ll = fine && image.n_large_dynamic > 0;
        d.fine_on = fine ? 1u : 0u; d.ll_on = ll ? 1u : 0u;
        d.fine_inv = fine ? 1.0 / (double)image.fine_cell : 0.0;
        fcell.resize(nb); d.fcell = fcell.data();
        const uint32_t TT = fine ? 2 * d.n_buckets : d.n_buckets;
        bucket_cnt.assign(TT + 1, 0); bucket_start.assign(TT + 2, 0);
        d.bucket_cnt = bucket_cnt.data(); d.bucket_start = bucket_start.data();
        for (uint32_t i = 0; i < nb; ++i) count_body_thread(d, i, true);
        memcpy(bucket_start.data(), bucket_cnt.data(), (TT + 1) * 4);
        exclusive_scan(bucket_start.data(), TT);
        const uint32_t E = bucket_start[TT];
        ent_body.assign(E + 1, 0); ent_key.assign(E + 1, 0); ent_aabb.resize(E + 1); d.ent_aabb = ent_aabb.data(); ent_off.assign(std::max(E, d.n_buckets) + 2, 0);
        d.cap_entries = E; d.ent_body = ent_body.data(); d.ent_key = ent_key.data(); d.ent_off = ent_off.data();
        for (uint32_t i = 0; i < nb; ++i) {
            if (body_is_small(d, body_flags(d, i))) {
                fill_fine(d, i);
                continue;
            }
            const CellRange r = cell_range(d, i);
            for (uint32_t k = 0; k < r.count; ++k) fill_cell(d, i, cell_bucket(r, k));
        }
        uint32_t P = 0;
        if (!fine || ll) {
            const uint32_t EC = bucket_start[d.n_buckets];  // entries of the coarse buckets
            for (uint32_t b = 0; b < d.n_buckets; ++b) sort_bucket_thread(d, b);
            for (uint32_t e = 0; e < EC; ++e) ent_off[e] = entry_pairs_thread(d, e, nullptr);
            exclusive_scan(ent_off.data(), EC);
            P = ent_off[EC];
        }
        const uint32_t P_ll = P;
        pair_cnt.assign(nb + 2, 0);
        d.pair_cnt = pair_cnt.data();
        if (fine) {
            pair_cnt[0] = P_ll;
            for (uint32_t a = 0; a < nb; ++a) {
                if (!body_is_small(d, body_flags(d, a))) continue;
                pair_cnt[a + 1] = fine_body_pairs(d, a, nullptr, nullptr);
            }
            exclusive_scan(pair_cnt.data(), nb + 1);
            P = pair_cnt[nb + 1];
        }
        pairs.assign(P + 1, make_uint2(0, 0));
        d.cap_pairs = P; d.pairs = pairs.data();
        if (!fine || ll) {
            const uint32_t EC = bucket_start[d.n_buckets];
            for (uint32_t e = 0; e < EC; ++e) entry_pairs_thread(d, e, pairs.data() + ent_off[e]);
        }
        if (fine) {
            for (uint32_t a = 0; a < nb; ++a) {
                if (!body_is_small(d, body_flags(d, a))) continue;
                const uint32_t at = pair_cnt[a + 1];
                sort_item_pairs(pairs.data() + at, fine_body_pairs(d, a, nullptr, pairs.data() + at));
            }
        }
        n_pairs_last = P;

        // ---- narrowphase (K6) ----
        m_hdr.assign(P + 1, make_uint4(0, 0, 0, 0)); m_color.assign(P + 1, COLOR_NONE);
        m_g0.resize(P + 1); m_g1.resize(P + 1); m_r0.resize(P + 1); m_r1.resize(P + 1);
        d.m_hdr = m_hdr.data(); d.m_g0 = m_g0.data(); d.m_g1 = m_g1.data(); d.m_r0 = m_r0.data(); d.m_r1 = m_r1.data();
        d.m_color = m_color.data();
        m_prio.assign(P + 1, 0); d.m_prio = m_prio.data();
        // dataflow colouring inputs (R2D_EMU_FLOW=0 emulates the rounds-only path)
        d.flow = (getenv("R2D_EMU_FLOW") && atoi(getenv("R2D_EMU_FLOW")) == 0) ? 0u : 1u;
        adj_cnt.assign(nb, 0); adj_prio.assign((size_t)nb * ADJ_CAP, 0); cstate.assign(nb, make_uint4(0, 0, 0, 0));
        d.adj_cnt = adj_cnt.data(); d.adj_prio = adj_prio.data(); d.cstate = cstate.data();
        adj_head.assign(nb, 0); adj_pool.assign((size_t)P / 2 + 1024, make_uint4(0, 0, 0, 0));
        d.adj_head = adj_head.data(); d.adj_pool = adj_pool.data(); d.adj_pool_cap = (uint32_t)adj_pool.size();
        uint32_t M = 0, K = 0;
        for (uint32_t p = 0; p < P; ++p) {
            const int np = narrow_pair_thread(d, p);
            if (np >= 0) { M += 1; K += (uint32_t)np; }
        }

        // ---- colouring (K8) ----
        maxprio0.assign(nb, 0); maxprio1.assign(nb, 0); used.assign((size_t)nb * COLOR_WORDS, 0);
        color_count.assign(MAX_COLORS, 0); color_start.assign(MAX_COLORS + 1, 0); color_cursor.assign(MAX_COLORS, 0);
        d.maxprio0 = maxprio0.data(); d.maxprio1 = maxprio1.data(); d.used = used.data();
        d.color_count = color_count.data(); d.color_start = color_start.data(); d.color_cursor = color_cursor.data();
        uint32_t rounds = 0, n_colors = 0;
        bool flow_done = false;
        if (d.flow && !counters.flow_abort) {
            // the dataflow colouring, probed serially: every pass colours what is ready (k_color's flow phase)
            std::vector<uint32_t> ranks(P, 0);
            uint32_t pending = 0;
            for (uint32_t p = 0; p < P; ++p)
                if (m_color[p] == COLOR_PENDING) {
                    ranks[p] = flow_ranks(d, m_hdr[p], m_prio[p]);
                    ++pending;
                }
            while (pending && !counters.flow_fail) {
                uint32_t done = 0;
                for (uint32_t p = 0; p < P; ++p) {
                    if (m_color[p] != COLOR_PENDING) continue;
                    uint32_t c = 0, lag = 0;
                    if (flow_try(d, p, m_hdr[p].x, m_hdr[p].y, m_hdr[p].w & 3u, ranks[p], &c, &lag) == 1) {
                        color_count[c] += 1;
                        n_colors = std::max(n_colors, c + 1);
                        ++done;
                    }
                }
                if (!done && !counters.flow_fail) return R2D_ERR_CUDA;  // would be a stall on the GPU
                pending -= done;
            }
            if (counters.flow_fail) {  // more than FLOW_COLORS colours: start over with the rounds
                for (uint32_t p = 0; p < P; ++p)
                    if (m_color[p] < MAX_COLORS) m_color[p] = COLOR_PENDING;
                color_count.assign(MAX_COLORS, 0);
                n_colors = 0;
            } else {
                flow_done = true;
                counters.flow_used = 1;
            }
        }
        for (uint32_t p = 0; p < P && !flow_done; ++p) {
            if (m_color[p] != COLOR_PENDING) continue;
            const uint4 h = m_hdr[p];
            color_post(d, h.x, h.y, (h.w & 1u) != 0, (h.w & 2u) != 0, d.m_prio[p], 1);
        }
        for (uint32_t round = 1; round < MAX_COLOR_ROUNDS && !flow_done; ++round) {
            // Serial emulation caveat: a thread of round r must not see round-r writes of `used` by another winner on a
            // shared body — there is none (unique winner per body and round) — nor round r+1 posts, which go to the
            // other maxprio array.  So a plain loop is equivalent to the parallel round.
            uint32_t left = 0;
            for (uint32_t p = 0; p < P; ++p) {
                const int r = color_round_thread(d, p, round);
                if (r == 2) ++left;
                if (r == 1) { color_count[m_color[p]] += 1; n_colors = std::max(n_colors, m_color[p] + 1); }
            }
            rounds = round;
            if (left == 0) break;
        }
        if (counters.err & ERR_GRID_RANGE) return R2D_ERR_GRID_RANGE;
        // ---- owner bitmaps -> positions (colour-sorted, owner-slot order) + pre-step ----
        counters.n_colors = n_colors;
        d.own_words = (nb + 31u) / 32u;
        own_bits.assign((size_t)MAX_COLORS * d.own_words, 0u);
        own_pos.assign((size_t)MAX_COLORS * (d.own_words + 1u) + 2u, 0u);
        d.own_bits = own_bits.data(); d.own_pos = own_pos.data();
        for (uint32_t p = 0; p < P; ++p) owner_bit_thread(d, p);
        const uint32_t n_scan = n_colors * (d.own_words + 1u);
        for (uint32_t k = 0; k < n_scan; ++k) owner_count_thread(d, k);
        exclusive_scan(own_pos.data(), n_scan);
        for (uint32_t c = 0; c <= n_colors; ++c) color_start[c] = own_pos[(size_t)c * (d.own_words + 1u)];
        const uint32_t MP = color_start[n_colors];  // padded manifold slots
        s_hdr.assign(MP + 1, make_uint4(0, 0, S_EMPTY, 0)); s_nf.resize(MP + 1); s_inv.resize(MP + 1); s_r0.resize(MP + 1);
        s_r1.resize(MP + 1); s_pm0.resize(MP + 1); s_pm1.resize(MP + 1); s_acc0.resize(MP + 1); s_acc1.resize(MP + 1);
        s_dep.resize(MP + 1); d.s_dep = s_dep.data();
        d.s_hdr = s_hdr.data(); d.s_nf = s_nf.data(); d.s_inv = s_inv.data(); d.s_r0 = s_r0.data(); d.s_r1 = s_r1.data();
        d.s_pm0 = s_pm0.data(); d.s_pm1 = s_pm1.data(); d.s_acc0 = s_acc0.data(); d.s_acc1 = s_acc1.data();
        {
            std::vector<char> taken(MP + 1, 0);
            for (uint32_t p = 0; p < P; ++p) {
                if (m_color[p] >= MAX_COLORS) continue;
                const uint32_t at = manifold_slot(d, p);
                if (at >= MP || taken[at] || at < color_start[m_color[p]] || at >= color_start[m_color[p] + 1]) return R2D_ERR_CUDA;
                taken[at] = 1;
                gather_prestep_thread(d, p, at);
            }
        }

        // ---- substeps (lib.zig:199-250) ----
        const auto& jcs = image.joint_color_start;
        for (uint32_t s = 0; s < S; ++s) {
            for (uint32_t i = 0; i < nb; ++i) integrate_forces_thread(d, i, sub_dt, s + 1 == S);
            for (uint32_t it = 0; it < I; ++it) {
                for (size_t c = 0; c + 1 < jcs.size(); ++c)
                    for (uint32_t j = jcs[c]; j < jcs[c + 1]; ++j) solve_joint_thread(d, j, sub_dt);
                for (uint32_t c = 0; c < n_colors; ++c)
                    for (uint32_t m = color_start[c]; m < color_start[c + 1]; ++m) solve_contact_thread<true>(d, m, sub_dt, it);
            }
            for (uint32_t i = 0; i < nb; ++i) integrate_positions_thread(d, i, sub_dt);
        }
        if (counters.err & ERR_STALL) return R2D_ERR_CUDA;  // dataflow bookkeeping (rank / degree) is wrong
        stats.n_buckets = d.n_buckets;
        stats.n_entries = E;
        stats.n_pairs = P;
        stats.n_manifolds = M;
        stats.n_points = K;
        stats.n_colors = n_colors;
        stats.n_color_rounds = rounds;
        stats.n_dropped = counters.n_dropped;
        return R2D_OK;
    }
};

}  // namespace

static r2d::host::BatchBase* r2d_new_backend(int, std::string&) { return new EmuBatch(); }
static int r2d_backend_device_count() { return 0; }
#define R2D_API(name) emu_##name
#include "../../resolve2d_b200/csrc/r2d_capi.inc"
