// r2d_tiles.cuh — the substep loop of ONE large world without joints (BASELINE config 2: the 100k-body pile): one spatial
// TILE of bodies per SM, its contact records resident in shared memory for the whole call.
//
// Device slots are in Morton order, so the bodies [t B, (t + 1) B) are a compact patch of the world and most manifolds
// touch two bodies of one patch.  CTA t runs ALL manifolds owned by its bodies (owner = the lower slot of the non-static
// bodies: per colour a contiguous range of the owner-ordered records, see manifold_owner) and keeps the momentum word of
// every body that no other CTA touches in shared memory.  Inside a tile the colours of a sweep are separated by
// __syncthreads(), exactly as in k_world_solve: a record of colour c finds its tile-local bodies finished with colours
// < c without asking.  Only bodies that a foreign CTA touches ("shared": marked by the partition kernel, about a quarter)
// and the bodies it owns nothing of stay in global memory with the 16-byte {momentum, version} word of the dataflow
// sweep (r2d_pipeline.cuh, solve_contact_thread<true>): the k-th contact (in colour order) of such a body in iteration
// `it` waits until the version is it * degree + k and publishes momentum and version + 1 in ONE 16-byte store.  A tile
// at phase (it, c) only ever waits for phases (it', c') < (it, c) of its neighbours, and all CTAs are resident
// (cooperative launch), so the lowest unfinished phase of the grid can always run: no deadlock.
// Same per-body update sequence (ascending colour) as every other solver flavour: bit-identical results.
//
// One lane per contact POINT, slots placed per colour with the two-point manifolds first (see r2d_world.cuh): every lane
// of every warp runs the same one-point chain, all lanes of a phase in ONE converged pass.  (Round 1's tile solver let
// every warp poll its 32 records' bodies and update the ready lanes as they came: 690 warp instructions per 32 records
// and 19 of 32 lanes active, for a 150-instruction update.)
//
// If a tile does not fit (bodies, slots or non-local references) the kernel declines before touching anything and the
// host launches k_solve_persistent on the same records.
#pragma once
#include <cooperative_groups.h>

#include "r2d_world.cuh"

namespace r2d {

constexpr int TILE_TPB = 512;
constexpr uint32_t TILE_MAX_BODIES = 1024;
constexpr size_t TILE_SMEM_BYTES = 232448 - 6144;   // all of it (5 KB of static shared memory for the colour tables)
// slot: hdr (8) + nfb (16) + r (16) + ma (16); non-local reference ("ext"): 2 x (slot, rank | degree << 16) + inv (16)
constexpr uint32_t TILE_SLOT_BYTES = 56, TILE_EXT_BYTES = 32;
constexpr uint32_t TILE_BODY_BYTES = 20;             // momentum word (w = 1 / mass) + 1 / inertia
// slot header x: local index of body 1 | local index of body 2 << 16 (valid where TL_LOC is set)
//             y: flags | ext index << 8
constexpr uint32_t TL_LOC1 = 1u, TL_LOC2 = 2u, TL_ST1 = 4u, TL_ST2 = 8u, TL_SKIP = 16u, TL_NOPT = 32u, TL_B = 64u, TL_A = 128u;

__host__ __device__ inline void tile_capacity(uint32_t max_slots, uint32_t& R, uint32_t& R_ext) {
    // two thirds of the slots may reference a non-local body
    const size_t room = TILE_SMEM_BYTES - (size_t)TILE_MAX_BODIES * TILE_BODY_BYTES - 64;
    R = (uint32_t)(room / (TILE_SLOT_BYTES + (2 * TILE_EXT_BYTES) / 3 + 1)) & ~3u;
    if (max_slots && R > max_slots) R = max_slots & ~3u;   // tests: force the decline
    R_ext = (2 * R / 3) & ~3u;
}

__global__ void __launch_bounds__(TILE_TPB) k_solve_tiles(Dev d, float sub_dt, uint32_t S, uint32_t I, uint32_t max_slots) {
    namespace cg = cooperative_groups;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t s_cnt[MAX_COLORS];      // per colour: one-point manifolds | two-point manifolds << 16
    __shared__ uint32_t s_cur[MAX_COLORS];      // fill cursors
    __shared__ uint32_t s_beg[MAX_COLORS + 1];  // first slot of the colour (even)
    __shared__ uint32_t s_cbeg[MAX_COLORS], s_cend[MAX_COLORS];   // the tile's records of every colour
    __shared__ uint32_t s_next, s_total;
    cg::grid_group grid = cg::this_grid();
    if (overflowed(d) || d.counters->err != 0u) return;  // uniform across the grid
    const uint32_t tid = threadIdx.x, nth = blockDim.x, lane = tid & 31u;
    const uint32_t B = d.tile_bodies;
    const uint32_t b0 = blockIdx.x * B < d.n_bodies ? blockIdx.x * B : d.n_bodies;
    const uint32_t b1 = b0 + B < d.n_bodies ? b0 + B : d.n_bodies;
    const uint32_t nb = b1 - b0, nc = d.counters->n_colors;
    uint32_t R, R_ext;
    tile_capacity(max_slots, R, R_ext);
    // ---- shared memory layout ----
    unsigned char* q = smem_raw;
    float4* t_mom = (float4*)q;     q += (size_t)TILE_MAX_BODIES * 16;   // momentum.x, momentum.y, ang_momentum, 1 / mass
    float4* s_nfb = (float4*)q;     q += (size_t)R * 16;                 // normal.x, normal.y, friction, bias
    float4* s_r = (float4*)q;       q += (size_t)R * 16;                 // r1.x, r1.y, r2.x, r2.y
    float4* s_ma = (float4*)q;      q += (size_t)R * 16;                 // mass_n, mass_t, accumulated_pn, accumulated_pt
    float4* x_inv = (float4*)q;     q += (size_t)R_ext * 16;             // inv_m1, inv_m2, inv_i1, inv_i2
    uint4* x_ref = (uint4*)q;       q += (size_t)R_ext * 16;             // slot 1, rank 1 | degree 1 << 16, slot 2, rank 2 | degree 2 << 16
    uint2* s_hdr = (uint2*)q;       q += (size_t)R * 8;
    float* t_ii = (float*)q;                                             // 1 / inertia of the tile's bodies
    auto is_local = [&](uint32_t slot, bool st) { return !st && slot >= b0 && slot < b1 && d.body_shared[slot] == 0u; };
    // ---- the tile's records per colour; slots and non-local references needed ----
    for (uint32_t c = tid; c < MAX_COLORS; c += nth) {
        s_cnt[c] = s_cur[c] = 0u;
        if (c < nc) {
            s_cbeg[c] = owner_rank(d, c, b0);
            s_cend[c] = (b1 < d.n_bodies) ? owner_rank(d, c, b1) : d.own_pos[(size_t)c * (d.own_words + 1u) + d.own_words];
        }
    }
    if (tid == 0) s_next = 0u;
    __syncthreads();
    for (uint32_t c = 0; c < nc; ++c)
        for (uint32_t m = s_cbeg[c] + tid; m < s_cend[c]; m += nth) {
            const uint4 h = d.s_hdr[m];
            const bool st1 = (h.z & 0x100u) != 0, st2 = (h.z & 0x200u) != 0;
            const uint32_t two = (h.z & 0xFFu) > 1u ? 1u : 0u;
            atomicAdd(&s_cnt[c], two ? 0x10000u : 1u);
            if (!(is_local(h.x, st1) && is_local(h.y, st2))) atomicAdd(&s_next, 1u + two);
        }
    __syncthreads();
    if (tid < 32u) {   // slots per colour, exclusive scan by one warp (8 colours per lane)
        constexpr uint32_t PER = MAX_COLORS / 32;
        uint32_t v[PER], sum = 0;
#pragma unroll
        for (uint32_t k = 0; k < PER; ++k) {
            const uint32_t cnt = s_cnt[tid * PER + k];
            v[k] = ((cnt & 0xFFFFu) + 2u * (cnt >> 16) + 1u) & ~1u;   // even: the next colour starts on an even slot
            sum += v[k];
        }
        uint32_t inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (tid >= (uint32_t)o) inc += t;
        }
        uint32_t run = inc - sum;
#pragma unroll
        for (uint32_t k = 0; k < PER; ++k) {
            s_beg[tid * PER + k] = run;
            run += v[k];
        }
        if (tid == 31u) {
            s_beg[MAX_COLORS] = run;
            s_total = run;
            // a colour population must fit the 16-bit halves of the counters
            if (run > R || s_next > R_ext || nb > TILE_MAX_BODIES || run >= 0xFFFFu) atomicOr(&d.counters->tile_fallback, 1u);
        }
    }
    __syncthreads();
    grid.sync();
    if (__ldcg(&d.counters->tile_fallback) != 0u) return;  // nothing has been modified: the host runs k_solve_persistent
    if (tid == 0) s_next = 0u;
    __syncthreads();
    // ---- stage the records: one slot per contact point ----
    for (uint32_t c = 0; c < nc; ++c) {
        const uint32_t lb = s_beg[c], n2 = s_cnt[c] >> 16;
        for (uint32_t m = s_cbeg[c] + tid; m < s_cend[c]; m += nth) {
            const uint4 h = d.s_hdr[m];
            const uint32_t np = h.z & 0xFFu;
            const bool st1 = (h.z & 0x100u) != 0, st2 = (h.z & 0x200u) != 0;
            const bool loc1 = is_local(h.x, st1), loc2 = is_local(h.y, st2);
            const float4 nf = d.s_nf[m], r0 = d.s_r0[m], pm0 = d.s_pm0[m];
            uint32_t flags = (loc1 ? TL_LOC1 : 0u) | (loc2 ? TL_LOC2 : 0u) | (st1 ? TL_ST1 : 0u) | (st2 ? TL_ST2 : 0u);
            const uint32_t idx = (loc1 ? h.x - b0 : 0u) | ((loc2 ? h.y - b0 : 0u) << 16);
            const uint32_t slot = np > 1u ? lb + (atomicAdd(&s_cur[c], 2u) & 0xFFFFu)
                                          : lb + 2u * n2 + (atomicAdd(&s_cur[c], 0x10000u) >> 16);
            uint32_t ext = 0;
            if (!(loc1 && loc2)) {
                ext = atomicAdd(&s_next, np > 1u ? 2u : 1u);
                const uint4 dep = d.s_dep[m];
                const uint4 ref = make_uint4(h.x, dep.x | (dep.y << 16), h.y, dep.z | (dep.w << 16));
                const float4 inv = d.s_inv[m];
                x_ref[ext] = ref;
                x_inv[ext] = inv;
                if (np > 1u) {
                    x_ref[ext + 1u] = ref;
                    x_inv[ext + 1u] = inv;
                }
            }
            const uint32_t f0 = np == 0u ? TL_NOPT : (pm0.z >= 0.0f ? TL_SKIP : 0u);
            s_hdr[slot] = make_uint2(idx, flags | f0 | (np > 1u ? TL_A : 0u) | (ext << 8));
            s_nfb[slot] = make_float4(nf.x, nf.y, nf.z, pm0.w);
            s_r[slot] = r0;
            s_ma[slot] = make_float4(pm0.x, pm0.y, 0.0f, 0.0f);
            if (np > 1u) {
                const float4 r1 = d.s_r1[m], pm1 = d.s_pm1[m];
                s_hdr[slot + 1u] = make_uint2(idx, flags | (pm1.z >= 0.0f ? TL_SKIP : 0u) | TL_B | ((ext + 1u) << 8));
                s_nfb[slot + 1u] = make_float4(nf.x, nf.y, nf.z, pm1.w);
                s_r[slot + 1u] = r1;
                s_ma[slot + 1u] = make_float4(pm1.x, pm1.y, 0.0f, 0.0f);
            }
        }
    }
    for (uint32_t i = tid; i < nb; i += nth) {
        const float4 pr = d.prop[b0 + i];
        t_ii[i] = (body_flags(d, b0 + i) & FLAG_STATIC) ? 0.0f : fdiv(1.0f, pr.y);   // prestep_manifold :104-111, per body
    }
    for (uint32_t s = 0; s < S; ++s) {
        for (uint32_t i = tid; i < nb; i += nth) {
            if (s == 0) integrate_forces_thread(d, b0 + i, sub_dt, S == 1);
            const float4 m = d.mom[b0 + i];
            const bool st = (body_flags(d, b0 + i) & FLAG_STATIC) != 0;
            t_mom[i] = make_float4(m.x, m.y, m.z, st ? 0.0f : fdiv(1.0f, d.prop[b0 + i].x));
        }
        __syncthreads();
        grid.sync();   // the version words of the shared bodies restart at 0 in every substep
        for (uint32_t it = 0; it < I; ++it)
            for (uint32_t c = 0; c < nc; ++c) {
                const uint32_t cnt = s_cnt[c];
                if (cnt == 0u) continue;   // uniform
                const uint32_t lb = s_beg[c], le = lb + (cnt & 0xFFFFu) + 2u * (cnt >> 16);
                for (uint32_t sb = lb + (tid & ~31u); sb < le; sb += nth) {
                    const uint32_t sl = sb + lane;
                    const bool live = sl < le;
                    const uint2 h = live ? s_hdr[sl] : make_uint2(0u, TL_NOPT | TL_ST1 | TL_ST2 | TL_B);
                    const uint32_t f = h.y;
                    const bool st1 = (f & TL_ST1) != 0, st2 = (f & TL_ST2) != 0;
                    const bool loc1 = (f & TL_LOC1) != 0, loc2 = (f & TL_LOC2) != 0;
                    const uint32_t i1 = h.x & 0xFFFFu, i2 = h.x >> 16;
                    float4 m1 = make_float4(0, 0, 0, 0), m2 = m1;
                    float inv_m1 = 0.0f, inv_m2 = 0.0f, inv_i1 = 0.0f, inv_i2 = 0.0f;
                    uint32_t g1 = 0, g2 = 0, e1 = 0, e2 = 0;
                    if (live) {
                        if (!(loc1 && loc2)) {
                            const uint4 ref = x_ref[f >> 8];
                            const float4 inv = x_inv[f >> 8];
                            g1 = ref.x;
                            g2 = ref.z;
                            e1 = it * (ref.y >> 16) + (ref.y & 0xFFFFu);
                            e2 = it * (ref.w >> 16) + (ref.w & 0xFFFFu);
                            inv_m1 = inv.x;
                            inv_m2 = inv.y;
                            inv_i1 = inv.z;
                            inv_i2 = inv.w;
                        }
                        if (loc1) {
                            m1 = t_mom[i1];
                            inv_m1 = m1.w;
                            inv_i1 = t_ii[i1];
                        } else if (st1) {
                            m1 = d.mom[g1];   // static bodies are never written during a sweep
                        } else {              // a body that another tile touches too: wait for its turn
                            uint32_t spins = 0;
                            for (;;) {
                                m1 = ld_body_word(&d.mom[g1]);
                                if (f2u(m1.w) == e1) break;
                                if ((++spins & 0x3FFu) == 0u) {   // a stall would be a bug: flag it and run to the end
                                    if (spins > (1u << 22)) atomicOr(&d.counters->err, ERR_STALL);
                                    if (*((volatile uint32_t*)&d.counters->err) & ERR_STALL) break;
                                }
                            }
                        }
                        if (loc2) {
                            m2 = t_mom[i2];
                            inv_m2 = m2.w;
                            inv_i2 = t_ii[i2];
                        } else if (st2) {
                            m2 = d.mom[g2];
                        } else {
                            uint32_t spins = 0;
                            for (;;) {
                                m2 = ld_body_word(&d.mom[g2]);
                                if (f2u(m2.w) == e2) break;
                                if ((++spins & 0x3FFu) == 0u) {
                                    if (spins > (1u << 22)) atomicOr(&d.counters->err, ERR_STALL);
                                    if (*((volatile uint32_t*)&d.counters->err) & ERR_STALL) break;
                                }
                            }
                        }
                    }
                    __syncwarp();
                    bool applied = false;
                    v2 dp = mk2(0.0f, 0.0f);
                    float c1 = 0.0f, c2 = 0.0f;
                    if (live && !(f & TL_NOPT)) {
                        const float4 nfb = s_nfb[sl], r = s_r[sl], ma = s_ma[sl];
                        ContactConst cc;
                        cc.normal = mk2(nfb.x, nfb.y);
                        cc.tangent = rot90cw(cc.normal);
                        cc.friction = nfb.z;
                        ContactPointConst pt;
                        pt.r1 = mk2(r.x, r.y);
                        pt.r2 = mk2(r.z, r.w);
                        pt.mass_n = ma.x;
                        pt.mass_t = ma.y;
                        pt.depth = (f & TL_SKIP) ? 0.0f : -1.0f;   // only the sign test of :154 looks at it (the bias is precomputed)
                        pt.bias = nfb.w;
                        v2 acc = mk2(ma.z, ma.w);
                        // calculateImpulses :139-145: pre-loop velocities of the two bodies
                        const v2 vl1 = scale2(mk2(m1.x, m1.y), inv_m1), vl2 = scale2(mk2(m2.x, m2.y), inv_m2);
                        const float om1 = fmul(m1.z, inv_i1), om2 = fmul(m2.z, inv_i2);
                        applied = contact_point_impulse(cc, pt, acc, vl1, om1, vl2, om2, dp);
                        *reinterpret_cast<float2*>(&s_ma[sl].z) = make_float2(acc.x, acc.y);
                        if (applied) {
                            c1 = cross2(pt.r1, dp);
                            c2 = cross2(pt.r2, dp);
                        }
                    }
                    // the second point's impulse goes to the lane of the first
                    const float bx = __shfl_down_sync(0xffffffffu, dp.x, 1), by = __shfl_down_sync(0xffffffffu, dp.y, 1);
                    const float bc1 = __shfl_down_sync(0xffffffffu, c1, 1), bc2 = __shfl_down_sync(0xffffffffu, c2, 1);
                    const bool b_applied = __shfl_down_sync(0xffffffffu, applied ? 1 : 0, 1) != 0 && (f & TL_A) != 0;
                    if (live && !(f & TL_B)) {
                        v2 lin1 = mk2(0.0f, 0.0f), lin2 = mk2(0.0f, 0.0f);
                        float rot1 = 0.0f, rot2 = 0.0f;
                        if (applied) {      // :197-206, point 0
                            if (!st1) {
                                lin1 = sub2(lin1, dp);
                                rot1 = fsub(rot1, c1);
                            }
                            if (!st2) {
                                lin2 = add2(lin2, dp);
                                rot2 = fadd(rot2, c2);
                            }
                        }
                        if (b_applied) {    // point 1
                            const v2 dq = mk2(bx, by);
                            if (!st1) {
                                lin1 = sub2(lin1, dq);
                                rot1 = fsub(rot1, bc1);
                            }
                            if (!st2) {
                                lin2 = add2(lin2, dq);
                                rot2 = fadd(rot2, bc2);
                            }
                        }
                        // :208-212; static bodies receive a zero impulse in the reference (`momentum += 0`): not written
                        if (!st1) {
                            const float nx = fadd(m1.x, lin1.x), ny = fadd(m1.y, lin1.y), na = fadd(m1.z, rot1);
                            if (loc1)
                                t_mom[i1] = make_float4(nx, ny, na, m1.w);
                            else
                                st_body_word(&d.mom[g1], make_float4(nx, ny, na, u2f(e1 + 1u)));
                        }
                        if (!st2) {
                            const float nx = fadd(m2.x, lin2.x), ny = fadd(m2.y, lin2.y), na = fadd(m2.z, rot2);
                            if (loc2)
                                t_mom[i2] = make_float4(nx, ny, na, m2.w);
                            else
                                st_body_word(&d.mom[g2], make_float4(nx, ny, na, u2f(e2 + 1u)));
                        }
                    }
                }
                __syncthreads();
            }
        grid.sync();   // every tile has added its impulses to the shared bodies
        // end of substep s fused with the start of substep s + 1 (per body, same thread)
        for (uint32_t i = tid; i < nb; i += nth) {
            const uint32_t b = b0 + i;
            if (!(body_flags(d, b) & FLAG_STATIC) && d.body_shared[b] == 0u) {
                const float4 m = t_mom[i];
                d.mom[b] = make_float4(m.x, m.y, m.z, 0.0f);
            }
            integrate_positions_thread(d, b, sub_dt);
            if (s + 1 < S) integrate_forces_thread(d, b, sub_dt, s + 2 == S);
        }
    }
    // accumulated impulses are not needed after the call (collision.zig:102-133 re-creates the manifolds)
}

}  // namespace r2d
