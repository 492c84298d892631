// r2d_world.cuh — batches of small independent worlds (BASELINE config 5: 4,096 worlds x 256 bodies): ONE CTA runs the
// whole substep loop of lib.zig:199-250 for ONE world out of shared memory and registers.
//
//   stage   bodies of the world: momentum, 1/mass, 1/inertia -> shared; pose, force, mass, inertia -> registers of the
//           thread that integrates the body                                                global -> on-chip, once
//           the world's manifolds (its candidate pairs are a contiguous slice of the pair list): counting sort by colour
//           in shared memory, preStep (collision.zig:102-133) and the Baumgarte bias evaluated on the way (inputs are
//           constant during a call, Q5); ONE SLOT PER CONTACT POINT                         global -> shared, once
//   S x {   positions of the previous substep + gravity + momentum (lib.zig:200-216,238-249)    registers / shared
//           I x colours x { calculateImpulses (collision.zig:135-218), __syncthreads() }        shared only   }
//   export  pos / momentum / force / AABB                                                 on-chip -> global, once
//
// No record, accumulated impulse or momentum word goes through L2 between the staging and the export, there is no
// inter-CTA synchronisation, and no separate partition / pre-step kernel (k_scan_owners, k_partition_prestep and the
// owner bitmaps are not used on this path).  Manifolds of one colour share no non-static body, so their order inside a
// colour is irrelevant: every body sees its contacts in ascending colour, as everywhere else — bit-identical results.
//
// One lane per contact POINT: the two points of a manifold see the same pre-loop velocities (Q8) and are independent
// until their impulses are summed, so they sit on two adjacent lanes; the second lane hands its impulse to the first
// (5 shuffles), which applies "point 0, then point 1" in the reference's order.  Every lane of every warp then runs the
// same ~130-instruction one-point chain — a sweep phase is bound by the latency of its longest chain, and with one
// THREAD per manifold nearly every warp held a two-point manifold and ran both points back to back.
//
// A world whose slots do not fit the shared-memory cache (sized by the host from the largest world of the previous
// call) runs the same code with its records in ITS OWN slice of the global record arrays (slot s -> array set s & 1,
// index p0 + s / 2: at most two slots per candidate pair), so capacity never limits correctness.
#pragma once
#include "r2d_pipeline.cuh"

namespace r2d {

constexpr uint32_t WORLD_MAX_BODIES = 512;
constexpr int WORLD_SOLVE_TPB = 128;
// ONE small world without joints (a scene of game size) runs on the same kernel with one CTA of 512 threads: a colour phase
// is then one pass of the CTA over shared memory instead of a round of L2 hand-offs between 148 tiles of a few bodies each
constexpr int WORLD_SINGLE_TPB = 512;
constexpr uint32_t WORLD_SINGLE_MAX_BODIES = 1024;   // (the slot header holds 10-bit body indices)
// slot header: ref | inc << 10 (world-local body slots) | flags << 20
constexpr uint32_t WS_ST1 = 1u << 20, WS_ST2 = 1u << 21;
constexpr uint32_t WS_SKIP = 1u << 22;    // depth >= 0: the point is skipped and its accumulated impulses are zeroed (:154-158)
constexpr uint32_t WS_NOPT = 1u << 23;    // a manifold without points (or a padding lane): contributes `momentum += 0` only
constexpr uint32_t WS_B = 1u << 24;       // second point of a manifold: the lane only feeds the lane before it
constexpr uint32_t WS_A = 1u << 25;       // first point of a two-point manifold
constexpr uint32_t WORLD_SLOT_BYTES = 4 + 3 * 16;    // hdr, nfb, r, ma
constexpr uint32_t WORLD_BODY_BYTES = 16 + 4 + 1;    // momentum word, 1 / inertia, static flag

__host__ __device__ inline size_t world_smem_bytes(uint32_t nb_cap, uint32_t R, bool joints = false) {
    // (nb_cap is a multiple of 4: the byte flags round up to 4 per body; worlds with joints also keep pose + torque per body)
    return (size_t)nb_cap * (joints ? 40 : 24) + (size_t)R * WORLD_SLOT_BYTES + 16;
}

// where the slots of a world live
template <bool SMEM>
struct WorldSlots;
template <>
struct WorldSlots<true> {
    uint32_t* hdr_;
    float4 *nfb_, *r_, *ma_;
    __device__ __forceinline__ uint32_t& hdr(uint32_t s) const { return hdr_[s]; }
    __device__ __forceinline__ float4& nfb(uint32_t s) const { return nfb_[s]; }   // normal.x, normal.y, friction, bias
    __device__ __forceinline__ float4& r(uint32_t s) const { return r_[s]; }       // r1.x, r1.y, r2.x, r2.y
    __device__ __forceinline__ float4& ma(uint32_t s) const { return ma_[s]; }     // mass_n, mass_t, accumulated_pn, accumulated_pt
};
template <>
struct WorldSlots<false> {   // even slots in one set of global arrays, odd slots in another, both at the world's pair slice
    uint32_t* hdr_[2];
    float4 *nfb_[2], *r_[2], *ma_[2];
    __device__ __forceinline__ uint32_t& hdr(uint32_t s) const { return hdr_[s & 1u][s >> 1]; }
    __device__ __forceinline__ float4& nfb(uint32_t s) const { return nfb_[s & 1u][s >> 1]; }
    __device__ __forceinline__ float4& r(uint32_t s) const { return r_[s & 1u][s >> 1]; }
    __device__ __forceinline__ float4& ma(uint32_t s) const { return ma_[s & 1u][s >> 1]; }
};

// r2d_process_read with page-locked destinations: the CTA writes the new state of its world straight into the caller's
// (mapped) host arrays as soon as the world is done — the transfer of the first worlds overlaps the substep loops of the
// later ones instead of starting after the whole batch (25 MB over PCIe for 4,096 worlds of 256 bodies).
struct WorldExport {
    float2* pos;        // null: no export
    float* angle;
    float2* mom;
    float* ang_mom;
    const uint32_t* host_of_dev;   // device slot -> host slot (the caller's order; a permutation inside every world)
};

// One joint of a world inside k_world_solve (Constraints/*.zig through the same solve_* functions as solve_joint_thread): the
// bodies' momentum words are the shared ones, pose and torque the copies the body phase of this substep published.
__device__ __noinline__ void world_solve_joint(const Dev& d, uint32_t j, uint32_t b0, float sub_dt, float4* s_mom, const float4* s_pose,
                                               const unsigned char* s_st) {
    const uint4 h = d.j_hdr[j];
    const float4 par = d.j_par[j], vec = d.j_vec[j];
    const float power_max = par.x, power_min = par.y, beta = par.z, target = par.w;
    auto load = [&](uint32_t slot) {
        const uint32_t i = slot - b0;
        const float4 p = s_pose[i], m = s_mom[i], pr = d.prop[slot];
        JointBody b;
        b.pos = mk2(p.x, p.y);
        b.angle = p.z;
        b.mom = mk2(m.x, m.y);
        b.ang = m.z;
        b.mass = pr.x;
        b.inertia = pr.y;
        b.torque = p.w;
        b.is_static = s_st[i] != 0;
        return b;
    };
    auto store = [&](uint32_t slot, const JointBody& b) {
        const uint32_t i = slot - b0;
        s_mom[i] = make_float4(b.mom.x, b.mom.y, b.ang, s_mom[i].w);
    };
    switch (h.x) {
        case 0: {  // distance
            JointBody b1 = load(h.y), b2 = load(h.z);
            solve_distance(b1, b2, target, beta, power_min, power_max);
            store(h.y, b1);
            store(h.z, b2);
        } break;
        case 1: {  // offset distance
            JointBody b1 = load(h.y), b2 = load(h.z);
            solve_offset_distance(b1, b2, mk2(vec.x, vec.y), mk2(vec.z, vec.w), target, beta, power_min, power_max);
            store(h.y, b1);
            store(h.z, b2);
        } break;
        case 2: {  // fixed position
            JointBody b = load(h.y);
            solve_fixed_position(b, mk2(vec.x, vec.y), beta, power_min, power_max);
            store(h.y, b);
        } break;
        default: {  // motor
            JointBody b = load(h.y);
            solve_motor(b, target, beta, power_min, power_max, sub_dt);
            store(h.y, b);
        } break;
    }
}

// Everything after the colour counts are known, for one world.  BPT = bodies a thread integrates (registers).
template <bool SMEM, int BPT, int WTPB, bool EXPORT, bool JOINTS>
__device__ __forceinline__ void world_run(const Dev& d, const WorldSlots<SMEM>& rc, uint32_t w, uint32_t b0, uint32_t nb,
                                          uint32_t p0, uint32_t p1, uint32_t nc, float sub_dt, uint32_t S, uint32_t I,
                                          float4* s_mom, float* s_ii, const unsigned char* s_st, const uint32_t* s_cnt,
                                          uint32_t* s_cur, const uint32_t* s_beg, float4 (&rp)[BPT], float4 (&rf)[BPT],
                                          const WorldExport& ex, float4* s_pose) {
    const uint32_t tid = threadIdx.x, nth = blockDim.x, lane = tid & 31u;
    // ---- place + preStep (collision.zig:102-133; once per call, Q5) ----
    for (uint32_t p = p0 + tid; p < p1; p += nth) {
        const uint32_t col = d.m_color[p];
        if (col >= MAX_COLORS) continue;
        const uint4 h = d.m_hdr[p];
        const float4 g0 = d.m_g0[p], g1 = d.m_g1[p], ra = d.m_r0[p];
        const uint32_t np = h.z & 0xFFu;
        const uint32_t i1 = h.x - b0, i2 = h.y - b0;
        const float mu1 = d.prop[h.x].z, mu2 = d.prop[h.y].z;
        ContactConst c;
        c.normal = mk2(g0.x, g0.y);
        c.tangent = rot90cw(c.normal);          // :113
        c.friction = fsqrt(fmul(mu1, mu2));     // :114
        c.inv_m1 = s_mom[i1].w;                 // static ? 0 : 1 / mass  (prestep_manifold, evaluated per body)
        c.inv_i1 = s_ii[i1];
        c.inv_m2 = s_mom[i2].w;
        c.inv_i2 = s_ii[i2];
        const uint32_t lb = s_beg[col], n2 = s_cnt[col] >> 16;
        const uint32_t base = i1 | (i2 << 10) | (s_st[i1] ? WS_ST1 : 0u) | (s_st[i2] ? WS_ST2 : 0u);
        ContactPointConst pc;
        pc.r1 = mk2(ra.x, ra.y);
        pc.r2 = mk2(ra.z, ra.w);
        pc.depth = g1.x;
        pc.mass_n = pc.mass_t = 0.0f;
        if (np > 0) prestep_point(c, pc);
        const float bias0 = np > 0 ? contact_bias(g1.x, sub_dt) : 0.0f;
        const uint32_t f0 = np == 0 ? WS_NOPT : (g1.x >= 0.0f ? WS_SKIP : 0u);
        if (np <= 1) {
            const uint32_t s = lb + 2u * n2 + (atomicAdd(&s_cur[col], 0x10000u) >> 16);
            rc.hdr(s) = base | f0;
            rc.nfb(s) = make_float4(c.normal.x, c.normal.y, c.friction, bias0);
            rc.r(s) = ra;
            rc.ma(s) = make_float4(pc.mass_n, pc.mass_t, 0.0f, 0.0f);
        } else {
            const uint32_t s = lb + (atomicAdd(&s_cur[col], 2u) & 0xFFFFu);
            rc.hdr(s) = base | f0 | WS_A;
            rc.nfb(s) = make_float4(c.normal.x, c.normal.y, c.friction, bias0);
            rc.r(s) = ra;
            rc.ma(s) = make_float4(pc.mass_n, pc.mass_t, 0.0f, 0.0f);
            const float4 rb = d.m_r1[p];
            pc.r1 = mk2(rb.x, rb.y);
            pc.r2 = mk2(rb.z, rb.w);
            pc.depth = g1.y;
            prestep_point(c, pc);
            rc.hdr(s + 1u) = base | (g1.y >= 0.0f ? WS_SKIP : 0u) | WS_B;
            rc.nfb(s + 1u) = make_float4(c.normal.x, c.normal.y, c.friction, contact_bias(g1.y, sub_dt));
            rc.r(s + 1u) = rb;
            rc.ma(s + 1u) = make_float4(pc.mass_n, pc.mass_t, 0.0f, 0.0f);
        }
    }
    // ---- substeps ----
    const uint32_t g_lo = d.grav_off[w], g_hi = d.grav_off[w + 1];
    __syncthreads();   // the pre-step above reads 1 / mass out of the momentum words the body phase below rewrites (same bits, but a race)
    for (uint32_t s = 0; s < S; ++s) {
#pragma unroll
        for (int k = 0; k < BPT; ++k) {
            const uint32_t i = tid + (uint32_t)k * WTPB;
            if (i >= nb) continue;
            const bool st = s_st[i] != 0;
            float4 m = s_mom[i];
            const float inv_m = m.w;
            if (s > 0 && !st) {   // end of substep s - 1 (lib.zig:238-249)
                position_update(rp[k], m, rp[k].w, rf[k].w, sub_dt);
                rf[k] = make_float4(0.0f, 0.0f, 0.0f, rf[k].w);
            }
            if (s + 1 == S) {     // only the last AABB refresh is observable (Q3)
                const float4 sh = d.shape[b0 + i];
                d.aabb[b0 + i] = refreshed_aabb(rp[k], f2u(sh.z), sh.x, sh.y);
            }
            if (JOINTS) s_pose[i] = make_float4(rp[k].x, rp[k].y, rp[k].z, rf[k].z);   // what the joints of this substep read (torque: MotorJoint)
            if (st) continue;
            float4 f = rf[k];
            const float mass = rp[k].w;
            for (uint32_t g = g_lo; g < g_hi; ++g) {   // momentum_update with the world's gravity list
                f.x = fadd(f.x, fmul(0.0f, mass));
                f.y = fadd(f.y, fmul(-d.grav[g], mass));
            }
            m.x = fadd(m.x, fmul(f.x, sub_dt));
            m.y = fadd(m.y, fmul(f.y, sub_dt));
            m.z = fadd(m.z, fmul(f.z, sub_dt));
            m.w = inv_m;
            s_mom[i] = m;
        }
        __syncthreads();
        for (uint32_t it = 0; it < I; ++it) {
            if (JOINTS) {   // lib.zig:226-231: the joints of the world in sweep order = (colour, list index), a barrier per colour
                const uint32_t j0 = d.world_joint_start[w], j1 = d.world_joint_start[w + 1];
                for (uint32_t jc = 0; jc < d.n_joint_colors; ++jc) {
                    for (uint32_t e = j0 + tid; e < j1; e += nth) {
                        const uint32_t j = d.world_joint[e];
                        if (d.j_hdr[j].w == jc) world_solve_joint(d, j, b0, sub_dt, s_mom, s_pose, s_st);
                    }
                    __syncthreads();
                }
            }
            for (uint32_t c = 0; c < nc; ++c) {
                const uint32_t cnt = s_cnt[c];
                if (cnt == 0u) continue;   // uniform
                const uint32_t lb = s_beg[c], le = lb + (cnt & 0xFFFFu) + 2u * (cnt >> 16);
                // whole warps (the two lanes of a manifold are in one warp: lb is even); idle lanes only shuffle
                for (uint32_t sb = lb + (tid & ~31u); sb < le; sb += nth) {
                    const uint32_t sl = sb + lane;
                    const bool live = sl < le;
                    const uint32_t h = live ? rc.hdr(sl) : (WS_NOPT | WS_ST1 | WS_ST2 | WS_B);
                    const uint32_t i1 = h & 0x3FFu, i2 = (h >> 10) & 0x3FFu;
                    const bool st1 = (h & WS_ST1) != 0, st2 = (h & WS_ST2) != 0;
                    bool applied = false;
                    v2 dp = mk2(0.0f, 0.0f);
                    float c1 = 0.0f, c2 = 0.0f;
                    float4 m1 = make_float4(0, 0, 0, 0), m2 = m1;
                    if (live) {
                        m1 = s_mom[i1];
                        m2 = s_mom[i2];
                        if (!(h & WS_NOPT)) {
                            const float4 nfb = rc.nfb(sl), r = rc.r(sl), ma = rc.ma(sl);
                            ContactConst cc;
                            cc.normal = mk2(nfb.x, nfb.y);
                            cc.tangent = rot90cw(cc.normal);
                            cc.friction = nfb.z;
                            ContactPointConst pt;
                            pt.r1 = mk2(r.x, r.y);
                            pt.r2 = mk2(r.z, r.w);
                            pt.mass_n = ma.x;
                            pt.mass_t = ma.y;
                            pt.depth = (h & WS_SKIP) ? 0.0f : -1.0f;   // only the sign test of :154 looks at it (the bias is precomputed)
                            pt.bias = nfb.w;
                            v2 acc = mk2(ma.z, ma.w);
                            // calculateImpulses :139-145: pre-loop velocities of the two bodies
                            const v2 vl1 = scale2(mk2(m1.x, m1.y), m1.w), vl2 = scale2(mk2(m2.x, m2.y), m2.w);
                            const float om1 = fmul(m1.z, s_ii[i1]), om2 = fmul(m2.z, s_ii[i2]);
                            applied = contact_point_impulse(cc, pt, acc, vl1, om1, vl2, om2, dp);
                            *reinterpret_cast<float2*>(&rc.ma(sl).z) = make_float2(acc.x, acc.y);
                            if (applied) {
                                c1 = cross2(pt.r1, dp);
                                c2 = cross2(pt.r2, dp);
                            }
                        }
                    }
                    // the second point's impulse goes to the lane of the first
                    const float bx = __shfl_down_sync(0xffffffffu, dp.x, 1), by = __shfl_down_sync(0xffffffffu, dp.y, 1);
                    const float bc1 = __shfl_down_sync(0xffffffffu, c1, 1), bc2 = __shfl_down_sync(0xffffffffu, c2, 1);
                    const bool b_applied = __shfl_down_sync(0xffffffffu, applied ? 1 : 0, 1) != 0 && (h & WS_A) != 0;
                    __syncwarp();   // the lane of a second point has read the two body words its neighbour is about to write
                    if (live && !(h & WS_B)) {
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
                        if (!st1) s_mom[i1] = make_float4(fadd(m1.x, lin1.x), fadd(m1.y, lin1.y), fadd(m1.z, rot1), m1.w);
                        if (!st2) s_mom[i2] = make_float4(fadd(m2.x, lin2.x), fadd(m2.y, lin2.y), fadd(m2.z, rot2), m2.w);
                    }
                }
                __syncthreads();
            }
        }
    }
    // ---- export ----
    if (S > 0) {
        float4 xa[BPT];
        float2 xb[BPT];
        uint32_t xj[BPT];
#pragma unroll
        for (int k = 0; k < BPT; ++k) {
            const uint32_t i = tid + (uint32_t)k * WTPB;
            xj[k] = 0xFFFFFFFFu;
            if (i >= nb) continue;
            const float4 m = s_mom[i];
            if (!s_st[i]) {
                position_update(rp[k], m, rp[k].w, rf[k].w, sub_dt);
                d.pos[b0 + i] = make_float4(rp[k].x, rp[k].y, rp[k].z, 0.0f);
                d.frc[b0 + i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                d.mom[b0 + i] = make_float4(m.x, m.y, m.z, 0.0f);
            } else if (JOINTS) {
                d.mom[b0 + i] = make_float4(m.x, m.y, m.z, 0.0f);   // distance joints write momentum into static bodies too (Q10)
            }
            if (EXPORT) {
                xa[k] = make_float4(rp[k].x, rp[k].y, m.x, m.y);
                xb[k] = make_float2(rp[k].z, m.z);
                xj[k] = ex.host_of_dev[b0 + i] - b0;
            }
        }
        if (EXPORT) {   // through shared memory into the caller's order, then whole lines to the host
            float4* s_xa = s_mom;                                   // 16 B per body
            float2* s_xb = reinterpret_cast<float2*>(s_ii);         // 1 / inertia and the static flags: 8 B per body, contiguous
            __syncthreads();
#pragma unroll
            for (int k = 0; k < BPT; ++k)
                if (xj[k] < nb) {
                    s_xa[xj[k]] = xa[k];
                    s_xb[xj[k]] = xb[k];
                }
            __syncthreads();
            for (uint32_t x = tid; x < nb; x += nth) {
                const float4 a = s_xa[x];
                const float2 b = s_xb[x];
                ex.pos[b0 + x] = make_float2(a.x, a.y);
                ex.mom[b0 + x] = make_float2(a.z, a.w);
                ex.angle[b0 + x] = b.x;
                ex.ang_mom[b0 + x] = b.y;
            }
        }
    }
}

template <int BPT, int WTPB = WORLD_SOLVE_TPB, bool EXPORT = false, bool JOINTS = false>
__global__ void __launch_bounds__(WTPB, WTPB > 128 ? 1 : (BPT == 2 ? 6 : 4)) k_world_solve(Dev d, float sub_dt, uint32_t S, uint32_t I,
                                                                                  uint32_t nb_cap, uint32_t R, WorldExport ex) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t s_cnt[MAX_COLORS];      // per colour: one-point manifolds | two-point manifolds << 16
    __shared__ uint32_t s_cur[MAX_COLORS];      // fill cursors: pairs from the front of the colour's slots, singles behind them
    __shared__ uint32_t s_beg[MAX_COLORS + 1];  // first slot of the colour (even)
    __shared__ uint32_t s_nc, s_fits;
    if (overflowed(d) || d.counters->err != 0u) return;   // an abandoned attempt leaves the body state untouched
    // ---- shared memory layout ----
    unsigned char* q = smem_raw;
    float4* s_mom = (float4*)q;   q += (size_t)nb_cap * 16;   // momentum.x, momentum.y, ang_momentum, 1 / mass (0: static)
    WorldSlots<true> sm;
    sm.nfb_ = (float4*)q;         q += (size_t)R * 16;
    sm.r_ = (float4*)q;           q += (size_t)R * 16;
    sm.ma_ = (float4*)q;          q += (size_t)R * 16;
    sm.hdr_ = (uint32_t*)q;       q += (size_t)R * 4;
    float* s_ii = (float*)q;      q += (size_t)nb_cap * 4;    // 1 / inertia (0: static)
    unsigned char* s_st = q;      q += (size_t)nb_cap * 4;    // 1: static
    float4* s_pose = (float4*)q;                              // JOINTS: x, y, angle, torque as of the start of the substep
    const uint32_t tid = threadIdx.x, nth = blockDim.x;
    for (uint32_t w = blockIdx.x; w < d.n_worlds; w += gridDim.x) {
        const uint32_t b0 = d.world_base[w], b1 = d.world_base[w + 1], nb = b1 - b0;
        // the world's candidate pairs: body-major list with the fine grid, bucket-major without
        // (a single world owns the whole list, whatever kernels wrote it)
        const uint32_t p0 = d.n_worlds == 1u ? 0u : (d.fine_on ? d.pair_cnt[b0 + 1] : d.ent_off[d.table_mult * b0]);
        const uint32_t p1 = d.n_worlds == 1u ? live_pairs(d) : (d.fine_on ? d.pair_cnt[b1 + 1] : d.ent_off[d.table_mult * b1]);
        for (uint32_t c = tid; c < MAX_COLORS; c += nth) s_cnt[c] = s_cur[c] = 0u;
        // ---- stage the bodies ----
        float4 rp[BPT], rf[BPT];   // pose + mass, force + inertia of the bodies this thread integrates
#pragma unroll
        for (int k = 0; k < BPT; ++k) {
            const uint32_t i = tid + (uint32_t)k * WTPB;
            rp[k] = rf[k] = make_float4(0, 0, 0, 0);
            if (i >= nb) continue;
            const float4 p = d.pos[b0 + i], m = d.mom[b0 + i], f = d.frc[b0 + i], pr = d.prop[b0 + i];
            const bool st = (body_flags(d, b0 + i) & FLAG_STATIC) != 0;
            rp[k] = make_float4(p.x, p.y, p.z, pr.x);
            rf[k] = make_float4(f.x, f.y, f.z, pr.y);
            s_mom[i] = make_float4(m.x, m.y, m.z, st ? 0.0f : fdiv(1.0f, pr.x));   // prestep_manifold :104-111, per body
            s_ii[i] = st ? 0.0f : fdiv(1.0f, pr.y);
            s_st[i] = st ? 1 : 0;
        }
        __syncthreads();
        // ---- counting sort of the world's manifolds by colour ----
        for (uint32_t p = p0 + tid; p < p1; p += nth) {
            const uint32_t c = d.m_color[p];
            if (c >= MAX_COLORS) continue;
            atomicAdd(&s_cnt[c], (d.m_hdr[p].z & 0xFFu) > 1u ? 0x10000u : 1u);
        }
        __syncthreads();
        if (tid < 32u) {   // slots per colour (1 per one-point, 2 per two-point manifold), exclusive scan by one warp (8 per lane)
            constexpr uint32_t PER = MAX_COLORS / 32;
            uint32_t v[PER], sum = 0, last = 0;
#pragma unroll
            for (uint32_t k = 0; k < PER; ++k) {
                const uint32_t cnt = s_cnt[tid * PER + k];
                v[k] = ((cnt & 0xFFFFu) + 2u * (cnt >> 16) + 1u) & ~1u;   // even: the next colour starts on an even slot
                sum += v[k];
                if (cnt) last = tid * PER + k + 1u;
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
            last = __reduce_max_sync(0xffffffffu, last);
            if (tid == 31u) {
                s_beg[MAX_COLORS] = run;
                s_nc = last;
                s_fits = run <= R ? 1u : 0u;
                atomicMax(&d.counters->max_world_m, run);
            }
        }
        __syncthreads();
        const uint32_t nc = s_nc;
        if (s_fits) {
            world_run<true, BPT, WTPB, EXPORT, JOINTS>(d, sm, w, b0, nb, p0, p1, nc, sub_dt, S, I, s_mom, s_ii, s_st, s_cnt, s_cur, s_beg, rp, rf, ex, s_pose);
        } else {   // the world's own slice of the global record arrays (see the header comment)
            WorldSlots<false> rg;
            rg.hdr_[0] = (uint32_t*)d.s_acc0 + p0;  rg.hdr_[1] = (uint32_t*)d.s_acc1 + p0;
            rg.nfb_[0] = d.s_nf + p0;               rg.nfb_[1] = d.s_inv + p0;
            rg.r_[0] = d.s_r0 + p0;                 rg.r_[1] = d.s_r1 + p0;
            rg.ma_[0] = d.s_pm0 + p0;               rg.ma_[1] = d.s_pm1 + p0;
            world_run<false, BPT, WTPB, EXPORT, JOINTS>(d, rg, w, b0, nb, p0, p1, nc, sub_dt, S, I, s_mom, s_ii, s_st, s_cnt, s_cur, s_beg, rp, rf, ex, s_pose);
        }
        __syncthreads();   // the next world of this CTA reuses the shared arrays
    }
}


// =====================================================================================================================
// Broadphase of a batch of small worlds: ONE CTA builds the grid of ONE world in shared memory and emits its candidate
// pairs — the same per-body functions as the device-wide kernels (count_body_thread, fill_fine, fill_cell,
// fine_body_pairs: SpatialHash.zig:19-137 + the filters of lib.zig:273-282), with the bucket tables, the grid entries,
// the stored AABBs and the remembered bucket lists of the world redirected to shared memory.  Count -> scan -> fill ->
// pair count -> scan -> pair write are separated by __syncthreads() instead of six kernel boundaries, no bucket counter
// or entry goes through L2, and nothing has to be zeroed in global memory.  The only inter-CTA step is the position of
// the world's pairs in the (world-major) pair list: a decoupled look-back over the per-world pair counts; worlds are
// handed out by an atomic ticket, so a CTA only ever waits for worlds that started before it.
// A world whose grid does not fit (entries of large bodies) raises `broad_fallback`: the attempt is abandoned (the
// broadphase does not modify body state) and the host redoes the step with the device-wide kernels.
// =====================================================================================================================
constexpr int WORLD_BROAD_TPB = 256;
constexpr uint32_t WORLD_BROAD_BIG = 32;   // large bodies of a world (floor, walls, bars), walked by the whole CTA
constexpr uint32_t WORLD_PARK = 16;        // partners a small body parks between the count and the write pass

__host__ __device__ inline size_t world_broad_smem_bytes(uint32_t nb_cap, uint32_t tw_cap, uint32_t ent_cap) {
    return (size_t)nb_cap * (16 + 16 + 16 + 4 * WORLD_PARK + 4) + (size_t)(2 * tw_cap + 1) * 8 + (size_t)ent_cap * 24 + 64;
}

__device__ __forceinline__ uint32_t cta_exclusive_scan(uint32_t v, uint32_t* s_warp, uint32_t* total) {   // blockDim <= 1024
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, nw = (blockDim.x + 31u) >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (uint32_t)o) inc += t;
    }
    __syncthreads();   // s_warp may still be read by the previous call
    if (lane == 31u) s_warp[warp] = inc;
    __syncthreads();
    uint32_t base = 0, sum = 0;
    for (uint32_t k = 0; k < nw; ++k) {
        const uint32_t x = s_warp[k];
        if (k < warp) base += x;
        sum += x;
    }
    *total = sum;
    return base + inc - v;
}

__global__ void __launch_bounds__(1024) k_world_broad(Dev d, uint32_t nb_cap, uint32_t tw_cap, uint32_t ent_cap,
                                                                 unsigned long long* state, uint32_t* ticket) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t s_warp[32], s_big[WORLD_BROAD_BIG];
    __shared__ uint32_t s_world, s_nbig, s_base;
    unsigned char* q = smem_raw;
    float4* s_aabb = (float4*)q;     q += (size_t)nb_cap * 16;
    uint4* s_bkt = (uint4*)q;        q += (size_t)nb_cap * 16;
    int4* s_fcell = (int4*)q;        q += (size_t)nb_cap * 16;
    uint32_t* s_cand = (uint32_t*)q; q += (size_t)nb_cap * 4 * WORLD_PARK;
    float4* s_eaabb = (float4*)q;    q += (size_t)ent_cap * 16;
    uint32_t* s_ebody = (uint32_t*)q; q += (size_t)ent_cap * 4;
    uint32_t* s_ekey = (uint32_t*)q; q += (size_t)ent_cap * 4;
    uint32_t* s_cnt = (uint32_t*)q;  q += (size_t)(2 * tw_cap + 1) * 4;
    uint32_t* s_start = (uint32_t*)q; q += (size_t)(2 * tw_cap + 1) * 4;
    uint32_t* s_pcnt = (uint32_t*)q;   // nb_cap + 1
    const uint32_t tid = threadIdx.x, nth = blockDim.x;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_world = atomicAdd(ticket, 1u);
        __syncthreads();
        const uint32_t w = s_world;
        if (w >= d.n_worlds) return;
        const uint32_t b0 = d.world_base[w], b1 = d.world_base[w + 1], nb = b1 - b0;
        const uint32_t tw = d.table_mult * nb;   // coarse buckets of the world; as many fine ones behind them
        // the per-body functions index with global slots / global bucket ids: shift the shared arrays accordingly
        Dev ds = d;
        ds.n_buckets = tw;                       // fine_bucket(): fine table right behind the world's coarse table
        ds.bucket_cnt = s_cnt - (size_t)d.table_mult * b0;
        ds.bucket_start = s_start - (size_t)d.table_mult * b0;
        ds.aabb = s_aabb - b0;
        ds.bkt = s_bkt - b0;
        ds.fcell = s_fcell - b0;
        ds.ent_body = s_ebody;
        ds.ent_key = s_ekey;
        ds.ent_aabb = s_eaabb;
        ds.cap_entries = ent_cap;
        for (uint32_t i = tid; i < nb; i += nth) s_aabb[i] = d.aabb[b0 + i];
        for (uint32_t k = tid; k < 2u * tw + 1u; k += nth) s_cnt[k] = 0u;
        if (tid == 0) s_nbig = 0u;
        __syncthreads();
        // ---- count (SpatialHash.zig:46-49) ----
        for (uint32_t i = tid; i < nb; i += nth) {
            const CellRange r = count_body_thread(ds, b0 + i, false);   // small bodies count their home cell themselves
            if (r.count == 0u) continue;
            // a large body: all of them are walked by the whole CTA (a thread walking 20 cells alone keeps 255 waiting)
            const uint32_t slot = atomicAdd(&s_nbig, 1u);
            if (slot < WORLD_BROAD_BIG) {
                s_big[slot] = i;
            } else {
                for (uint32_t k = 0; k < r.count; ++k) atomicAdd(&ds.bucket_cnt[cell_bucket(r, k)], 1u);
            }
        }
        __syncthreads();
        const uint32_t nbig = s_nbig < WORLD_BROAD_BIG ? s_nbig : WORLD_BROAD_BIG;
        for (uint32_t x = 0; x < nbig; ++x) {
            const CellRange r = cell_range(ds, b0 + s_big[x]);
            for (uint32_t k = tid; k < r.count; k += nth) atomicAdd(&ds.bucket_cnt[cell_bucket(r, k)], 1u);
        }
        __syncthreads();
        // ---- scan of the 2 tw bucket counts (:52-57) ----
        uint32_t n_entries = 0;
        {
            uint32_t carry = 0;
            for (uint32_t base = 0; base < 2u * tw; base += nth) {
                const uint32_t k = base + tid;
                const uint32_t v = k < 2u * tw ? s_cnt[k] : 0u;
                uint32_t total;
                const uint32_t ex = cta_exclusive_scan(v, s_warp, &total);
                if (k < 2u * tw) s_start[k] = carry + ex;
                carry += total;
            }
            if (tid == 0) s_start[2u * tw] = carry;
            n_entries = carry;
        }
        __syncthreads();
        const bool dead = n_entries > ent_cap;   // uniform: the grid of this world does not fit — the whole attempt is abandoned
        if (dead && tid == 0) atomicOr(&d.counters->broad_fallback, 1u);
        // ---- fill (:62-68) ----
        if (!dead) {
            for (uint32_t i = tid; i < nb; i += nth) {
                if (body_is_small(ds, body_flags(d, b0 + i))) {
                    fill_fine(ds, b0 + i);
                } else {
                    bool big = false;
                    for (uint32_t x = 0; x < nbig; ++x) big = big || s_big[x] == i;
                    if (big) continue;
                    const CellRange r = cell_range(ds, b0 + i);
                    for (uint32_t k = 0; k < r.count; ++k) fill_cell(ds, b0 + i, cell_bucket(r, k));
                }
            }
            for (uint32_t x = 0; x < nbig; ++x) {
                const CellRange r = cell_range(ds, b0 + s_big[x]);
                for (uint32_t k = tid; k < r.count; k += nth) fill_cell(ds, b0 + s_big[x], cell_bucket(r, k));
            }
            __syncthreads();
            // The fill's atomics decide the order inside a bucket; sorting every bucket by slot (they hold 0-3 entries)
            // makes the order in which a body meets its partners — and with it the pair list — a pure function of the state.
            for (uint32_t k = tid; k < 2u * tw; k += nth) {
                const uint32_t bs = s_start[k], be = s_start[k + 1];
                for (uint32_t x = bs + 1u; x < be; ++x) {
                    const uint32_t vb = s_ebody[x], vk = s_ekey[x];
                    const float4 va = s_eaabb[x];
                    uint32_t y = x;
                    while (y > bs && (s_ebody[y - 1] & ~ENT_STATIC) > (vb & ~ENT_STATIC)) {
                        s_ebody[y] = s_ebody[y - 1];
                        s_ekey[y] = s_ekey[y - 1];
                        s_eaabb[y] = s_eaabb[y - 1];
                        --y;
                    }
                    s_ebody[y] = vb;
                    s_ekey[y] = vk;
                    s_eaabb[y] = va;
                }
            }
        }
        __syncthreads();
        // ---- pairs of every small body: count, park the first WORLD_PARK partners ----
        for (uint32_t i = tid; i < nb; i += nth) {
            uint32_t got[WORLD_PARK];
            const uint32_t a = b0 + i;
            const bool live = !dead && body_is_small(ds, body_flags(d, a));
            const uint32_t n = live ? fine_body_pairs<WORLD_PARK>(ds, a, got, nullptr) : 0u;
            s_pcnt[i] = n;
#pragma unroll
            for (uint32_t x = 0; x < WORLD_PARK; x += 4)
                if (n > x) *reinterpret_cast<uint4*>(&s_cand[i * WORLD_PARK + x]) = make_uint4(got[x], got[x + 1], got[x + 2], got[x + 3]);
        }
        __syncthreads();
        // ---- scan of the pair counts inside the world; position of the world in the pair list by look-back ----
        uint32_t n_pairs_w = 0;
        {
            uint32_t carry = 0;
            for (uint32_t base = 0; base < nb; base += nth) {
                const uint32_t i = base + tid;
                const uint32_t v = i < nb ? s_pcnt[i] : 0u;
                uint32_t total;
                const uint32_t ex = cta_exclusive_scan(v, s_warp, &total);
                if (i < nb) s_cnt[i] = carry + ex;    // first pair of body i, relative to the world (the bucket counts are done with)
                carry += total;
            }
            n_pairs_w = carry;
        }
        if (tid < 32u) {
            const uint32_t lane = tid;
            uint32_t prefix = 0;
            if (w == 0) {
                if (lane == 0) atomicExch(&state[0], (2ull << 32) | n_pairs_w);
            } else {
                if (lane == 0) atomicExch(&state[w], (1ull << 32) | n_pairs_w);
                int first = (int)w - 1;
                for (;;) {
                    const int t = first - (int)lane;
                    unsigned long long x = 3ull << 32;   // beyond world 0: neutral, counts as "ready"
                    if (t >= 0) {
                        for (;;) {
                            x = *((volatile unsigned long long*)&state[t]);
                            if ((x >> 32) != 0ull) break;
                            __nanosleep(200);
                        }
                    }
                    const uint32_t is_prefix = __ballot_sync(0xffffffffu, (x >> 32) == 2ull);
                    const uint32_t upto = is_prefix ? (uint32_t)__ffs((int)is_prefix) - 1u : 31u;
                    uint32_t v = (t >= 0 && lane <= upto) ? (uint32_t)x : 0u;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                    prefix += v;
                    if (is_prefix || first < 32) break;
                    first -= 32;
                }
                if (lane == 0) atomicExch(&state[w], (2ull << 32) | (unsigned long long)(uint32_t)(prefix + n_pairs_w));
            }
            if (lane == 0) {
                s_base = prefix;
                atomicAdd(&d.counters->n_entries, n_entries);
                if (w == 0) d.pair_cnt[0] = 0u;
                if (w + 1 == d.n_worlds) {
                    d.pair_cnt[d.n_bodies + 1] = prefix + n_pairs_w;
                    d.counters->n_pairs = prefix + n_pairs_w;
                }
            }
        }
        __syncthreads();
        const uint32_t pbase = s_base;
        // ---- write, in the order the partners were met (deterministic: the buckets are sorted) ----
        for (uint32_t i = tid; i < nb; i += nth) {
            const uint32_t a = b0 + i;
            const uint32_t n = s_pcnt[i], at = pbase + s_cnt[i];
            d.pair_cnt[a + 1] = at;
            if (n == 0u || at + n > d.cap_pairs) continue;
            uint2* out = d.pairs + at;
            if (n <= WORLD_PARK) {
                for (uint32_t x = 0; x < n; ++x) out[x] = make_uint2(a, s_cand[i * WORLD_PARK + x]);
            } else {
                fine_body_pairs<WORLD_PARK>(ds, a, nullptr, out);
            }
        }
    }
}

}  // namespace r2d
