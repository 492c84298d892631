// r2d_world.cuh — batches of small independent worlds (BASELINE config 5: 4,096 worlds x 256 bodies): ONE CTA runs the
// whole substep loop of lib.zig:199-250 for ONE world out of shared memory.
//
//   stage   bodies of the world (pose, momentum, force, 1/mass, 1/inertia, mu)            global -> shared, once
//           the world's manifolds (its candidate pairs are a contiguous slice of the pair list): counting sort by colour
//           in shared memory, preStep (collision.zig:102-133) and the Baumgarte bias evaluated on the way (inputs are
//           constant during a call, Q5)                                                   global -> shared, once
//   S x {   positions of the previous substep + gravity + momentum (lib.zig:200-216,238-249)    shared only
//           I x colours x { calculateImpulses (collision.zig:135-218), __syncthreads() }        shared only   }
//   export  pos / momentum / force / AABB                                                 shared -> global, once
//
// No record, accumulated impulse or momentum word goes through L2 between the staging and the export, there is no
// inter-CTA synchronisation, and no separate partition / pre-step kernel (k_scan_owners, k_partition_prestep and the
// owner bitmaps are not used on this path).  Manifolds of one colour share no non-static body, so their order inside a
// colour is irrelevant: every body sees its contacts in ascending colour, as everywhere else — bit-identical results.
//
// A world whose manifolds do not fit the shared-memory cache (sized by the host from the largest world of the previous
// call) runs the same code with its records in ITS OWN slice [p0, p0 + M) of the global record arrays (M <= P: the slice
// of its candidate pairs), so capacity never limits correctness.
#pragma once
#include "r2d_pipeline.cuh"

namespace r2d {

constexpr uint32_t WORLD_MAX_BODIES = 512;
constexpr int WORLD_SOLVE_TPB = 128;
// record header: x = ref | inc << 16 (world-local body slots); y = flags | slot of point 1 << 8
constexpr uint32_t WS_NP_MASK = 3u, WS_ST1 = 4u, WS_ST2 = 8u, WS_D0 = 16u, WS_D1 = 32u;
constexpr uint32_t WORLD_REC_BYTES = 8 + 3 * 16;     // hdr, nfb, r0, ma0
constexpr uint32_t WORLD_PT1_BYTES = 2 * 16 + 4;     // r1, ma1, b1

struct WorldRecs {
    uint2* hdr;      // see above
    float4* nfb;     // normal.x, normal.y, friction, bias of point 0
    float4* r0;      // point 0: r1.x, r1.y, r2.x, r2.y
    float4* ma0;     // point 0: mass_n, mass_t, accumulated_pn, accumulated_pt
    float4* r1;      // point 1 (two-point manifolds only, own pool)
    float4* ma1;
    float* b1;       // bias of point 1
};

__host__ __device__ inline size_t world_smem_bytes(uint32_t nb_cap, uint32_t R, uint32_t R2) {
    return (size_t)nb_cap * (3 * 16 + 8 + 8) + (size_t)R * WORLD_REC_BYTES + (size_t)R2 * WORLD_PT1_BYTES + 16;
}

// one manifold of the current colour (collision.zig:135-218); `rc` is shared or global memory
__device__ __forceinline__ void world_sweep_record(const WorldRecs& rc, uint32_t l, float4* s_mom, const float2* s_inv) {
    const uint2 h = rc.hdr[l];
    const uint32_t i1 = h.x & 0xFFFFu, i2 = h.x >> 16;
    const float4 nfb = rc.nfb[l], r = rc.r0[l], ma = rc.ma0[l];
    const float4 m1 = s_mom[i1], m2 = s_mom[i2];
    const float2 v1 = s_inv[i1], v2_ = s_inv[i2];
    const int np = (int)(h.y & WS_NP_MASK);
    const bool st1 = (h.y & WS_ST1) != 0, st2 = (h.y & WS_ST2) != 0;
    ContactConst c;
    c.normal = mk2(nfb.x, nfb.y);
    c.tangent = rot90cw(c.normal);
    c.friction = nfb.z;
    c.inv_m1 = v1.x;
    c.inv_i1 = v1.y;
    c.inv_m2 = v2_.x;
    c.inv_i2 = v2_.y;
    ContactPointConst pts[2];
    v2 acc[2];
    pts[0].r1 = mk2(r.x, r.y);
    pts[0].r2 = mk2(r.z, r.w);
    pts[0].mass_n = ma.x;
    pts[0].mass_t = ma.y;
    pts[0].depth = (h.y & WS_D0) ? 0.0f : -1.0f;   // only the sign test of collision.zig:154 looks at it (the bias is precomputed)
    pts[0].bias = nfb.w;
    acc[0] = mk2(ma.z, ma.w);
    const uint32_t q = h.y >> 8;
    float4 mb = make_float4(0, 0, 0, 0);
    if (np > 1) {
        const float4 rb = rc.r1[q];
        mb = rc.ma1[q];
        pts[1].r1 = mk2(rb.x, rb.y);
        pts[1].r2 = mk2(rb.z, rb.w);
        pts[1].mass_n = mb.x;
        pts[1].mass_t = mb.y;
        pts[1].depth = (h.y & WS_D1) ? 0.0f : -1.0f;
        pts[1].bias = rc.b1[q];
        acc[1] = mk2(mb.z, mb.w);
    }
    BodyVel b1 = {mk2(m1.x, m1.y), m1.z}, b2 = {mk2(m2.x, m2.y), m2.z};
    solve_contact(c, np, pts, acc, st1, st2, b1, b2);
    if (np > 0) *reinterpret_cast<float2*>(&rc.ma0[l].z) = make_float2(acc[0].x, acc[0].y);
    if (np > 1) *reinterpret_cast<float2*>(&rc.ma1[q].z) = make_float2(acc[1].x, acc[1].y);
    // static bodies receive a zero impulse in the reference (`momentum += 0`); not writing them is the same value
    if (!st1) s_mom[i1] = make_float4(b1.mom.x, b1.mom.y, b1.ang, 0.0f);
    if (!st2) s_mom[i2] = make_float4(b2.mom.x, b2.mom.y, b2.ang, 0.0f);
}

// Everything after the colour counts are known, for one world whose records live in `rc` (shared memory, or the world's
// slice of the global record arrays).  Inlined once per storage so that the shared-memory copy uses LDS / STS.
__device__ __forceinline__ void world_run(const Dev& d, const WorldRecs& rc, uint32_t w, uint32_t b0, uint32_t nb, uint32_t p0,
                                          uint32_t p1, uint32_t nc, float sub_dt, uint32_t S, uint32_t I, float4* s_mom,
                                          float4* s_pos, float4* s_frc, float2* s_inv, const uint2* s_fm, uint32_t* s_cnt,
                                          const uint32_t* s_beg, uint32_t* s_n2p) {
    const uint32_t tid = threadIdx.x, nth = blockDim.x;
    // ---- place + preStep (collision.zig:102-133; once per call, Q5) ----
    for (uint32_t p = p0 + tid; p < p1; p += nth) {
        const uint32_t col = d.m_color[p];
        if (col >= MAX_COLORS) continue;
        const uint4 h = d.m_hdr[p];
        const float4 g0 = d.m_g0[p], g1 = d.m_g1[p], ra = d.m_r0[p];
        const uint32_t np = h.z & 0xFFu;
        const uint32_t i1 = h.x - b0, i2 = h.y - b0;
        const uint2 fm1 = s_fm[i1], fm2 = s_fm[i2];
        const float2 v1 = s_inv[i1], v2_ = s_inv[i2];
        ContactConst c;
        c.normal = mk2(g0.x, g0.y);
        c.tangent = rot90cw(c.normal);                      // :113
        c.friction = fsqrt(fmul(u2f(fm1.y), u2f(fm2.y)));   // :114
        c.inv_m1 = v1.x;
        c.inv_i1 = v1.y;
        c.inv_m2 = v2_.x;
        c.inv_i2 = v2_.y;
        const uint32_t l = s_beg[col] + atomicAdd(&s_cnt[col], 1u);
        uint32_t flags = np | ((fm1.x & FLAG_STATIC) ? WS_ST1 : 0u) | ((fm2.x & FLAG_STATIC) ? WS_ST2 : 0u);
        ContactPointConst pc;
        pc.r1 = mk2(ra.x, ra.y);
        pc.r2 = mk2(ra.z, ra.w);
        pc.depth = g1.x;
        pc.mass_n = pc.mass_t = 0.0f;
        if (np > 0) prestep_point(c, pc);
        if (g1.x >= 0.0f) flags |= WS_D0;
        rc.nfb[l] = make_float4(c.normal.x, c.normal.y, c.friction, np > 0 ? contact_bias(g1.x, sub_dt) : 0.0f);
        rc.r0[l] = ra;
        rc.ma0[l] = make_float4(pc.mass_n, pc.mass_t, 0.0f, 0.0f);
        if (np > 1) {
            const uint32_t slot = atomicAdd(s_n2p, 1u);
            const float4 rb = d.m_r1[p];
            pc.r1 = mk2(rb.x, rb.y);
            pc.r2 = mk2(rb.z, rb.w);
            pc.depth = g1.y;
            prestep_point(c, pc);
            if (g1.y >= 0.0f) flags |= WS_D1;
            rc.r1[slot] = rb;
            rc.ma1[slot] = make_float4(pc.mass_n, pc.mass_t, 0.0f, 0.0f);
            rc.b1[slot] = contact_bias(g1.y, sub_dt);
            flags |= slot << 8;
        }
        rc.hdr[l] = make_uint2(i1 | (i2 << 16), flags);
    }
    // ---- substeps ----
    const uint32_t g_lo = d.grav_off[w], g_hi = d.grav_off[w + 1];
    for (uint32_t s = 0; s < S; ++s) {
        for (uint32_t i = tid; i < nb; i += nth) {
            const uint2 fm = s_fm[i];
            const bool st = (fm.x & FLAG_STATIC) != 0;
            float4 p = s_pos[i], m = s_mom[i], f = s_frc[i];
            if (s > 0 && !st) {   // end of substep s - 1 (lib.zig:238-249)
                position_update(p, m, p.w, f.w, sub_dt);
                f = make_float4(0.0f, 0.0f, 0.0f, f.w);
                s_pos[i] = p;
                s_frc[i] = f;
            }
            if (s + 1 == S) {     // only the last AABB refresh is observable (Q3)
                const float4 sh = d.shape[b0 + i];
                d.aabb[b0 + i] = refreshed_aabb(p, fm.x, sh.x, sh.y);
            }
            if (st) continue;
            float mw = p.w;
            for (uint32_t g = g_lo; g < g_hi; ++g) {   // momentum_update with the world's gravity list
                f.x = fadd(f.x, fmul(0.0f, mw));
                f.y = fadd(f.y, fmul(-d.grav[g], mw));
            }
            m.x = fadd(m.x, fmul(f.x, sub_dt));
            m.y = fadd(m.y, fmul(f.y, sub_dt));
            m.z = fadd(m.z, fmul(f.z, sub_dt));
            s_mom[i] = m;
        }
        __syncthreads();
        for (uint32_t it = 0; it < I; ++it)
            for (uint32_t c = 0; c < nc; ++c) {
                const uint32_t lb = s_beg[c], le = s_beg[c + 1];
                if (lb == le) continue;   // uniform
                for (uint32_t l = lb + tid; l < le; l += nth) world_sweep_record(rc, l, s_mom, s_inv);
                __syncthreads();
            }
    }
    // ---- export ----
    for (uint32_t i = tid; i < nb; i += nth) {
        const uint2 fm = s_fm[i];
        if (fm.x & FLAG_STATIC) continue;
        float4 p = s_pos[i];
        const float4 m = s_mom[i], f = s_frc[i];
        if (S > 0) {
            position_update(p, m, p.w, f.w, sub_dt);
            d.pos[b0 + i] = make_float4(p.x, p.y, p.z, 0.0f);
            d.frc[b0 + i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            d.mom[b0 + i] = make_float4(m.x, m.y, m.z, 0.0f);
        }
    }
}

__global__ void __launch_bounds__(WORLD_SOLVE_TPB) k_world_solve(Dev d, float sub_dt, uint32_t S, uint32_t I, uint32_t nb_cap,
                                                                 uint32_t R, uint32_t R2) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t s_cnt[MAX_COLORS], s_beg[MAX_COLORS + 1];
    __shared__ uint32_t s_n2, s_nc, s_fits;
    if (overflowed(d) || d.counters->err != 0u) return;   // an abandoned attempt leaves the body state untouched
    // ---- shared memory layout ----
    unsigned char* q = smem_raw;
    float4* s_mom = (float4*)q;   q += (size_t)nb_cap * 16;   // momentum.x, momentum.y, ang_momentum, -
    float4* s_pos = (float4*)q;   q += (size_t)nb_cap * 16;   // x, y, angle, mass
    float4* s_frc = (float4*)q;   q += (size_t)nb_cap * 16;   // force.x, force.y, torque, inertia
    WorldRecs sm;
    sm.nfb = (float4*)q;          q += (size_t)R * 16;
    sm.r0 = (float4*)q;           q += (size_t)R * 16;
    sm.ma0 = (float4*)q;          q += (size_t)R * 16;
    sm.r1 = (float4*)q;           q += (size_t)R2 * 16;
    sm.ma1 = (float4*)q;          q += (size_t)R2 * 16;
    float2* s_inv = (float2*)q;   q += (size_t)nb_cap * 8;    // 1 / mass, 1 / inertia (0 for static bodies)
    uint2* s_fm = (uint2*)q;      q += (size_t)nb_cap * 8;    // flags, bits(mu)
    sm.hdr = (uint2*)q;           q += (size_t)R * 8;
    sm.b1 = (float*)q;
    const uint32_t tid = threadIdx.x, nth = blockDim.x;
    for (uint32_t w = blockIdx.x; w < d.n_worlds; w += gridDim.x) {
        const uint32_t b0 = d.world_base[w], b1 = d.world_base[w + 1], nb = b1 - b0;
        // the world's candidate pairs: body-major list with the fine grid, bucket-major without
        const uint32_t p0 = d.fine_on ? d.pair_cnt[b0 + 1] : d.ent_off[d.table_mult * b0];
        const uint32_t p1 = d.fine_on ? d.pair_cnt[b1 + 1] : d.ent_off[d.table_mult * b1];
        for (uint32_t c = tid; c < MAX_COLORS; c += nth) s_cnt[c] = 0u;
        if (tid == 0) s_n2 = 0u;
        // ---- stage the bodies ----
        for (uint32_t i = tid; i < nb; i += nth) {
            const float4 p = d.pos[b0 + i], m = d.mom[b0 + i], f = d.frc[b0 + i], pr = d.prop[b0 + i];
            const uint32_t flags = body_flags(d, b0 + i);
            const bool st = (flags & FLAG_STATIC) != 0;
            s_pos[i] = make_float4(p.x, p.y, p.z, pr.x);
            s_mom[i] = make_float4(m.x, m.y, m.z, 0.0f);
            s_frc[i] = make_float4(f.x, f.y, f.z, pr.y);
            s_inv[i] = make_float2(st ? 0.0f : fdiv(1.0f, pr.x), st ? 0.0f : fdiv(1.0f, pr.y));   // prestep_manifold, per body
            s_fm[i] = make_uint2(flags, f2u(pr.z));
        }
        __syncthreads();
        // ---- counting sort of the world's manifolds by colour ----
        for (uint32_t p = p0 + tid; p < p1; p += nth) {
            const uint32_t c = d.m_color[p];
            if (c >= MAX_COLORS) continue;
            atomicAdd(&s_cnt[c], 1u);
            if ((d.m_hdr[p].z & 0xFFu) > 1u) atomicAdd(&s_n2, 1u);
        }
        __syncthreads();
        if (tid < 32u) {   // exclusive scan of the colour populations by one warp (8 per lane)
            constexpr uint32_t PER = MAX_COLORS / 32;
            uint32_t v[PER], sum = 0, last = 0;
#pragma unroll
            for (uint32_t k = 0; k < PER; ++k) {
                v[k] = s_cnt[tid * PER + k];
                sum += v[k];
                if (v[k]) last = tid * PER + k + 1u;
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
                s_cnt[tid * PER + k] = 0u;   // becomes the fill cursor
            }
            last = __reduce_max_sync(0xffffffffu, last);
            if (tid == 31u) {
                s_beg[MAX_COLORS] = run;
                s_nc = last;
                const uint32_t n2 = s_n2;
                s_fits = (run <= R && n2 <= R2) ? 1u : 0u;
                s_n2 = 0u;                    // becomes the allocation cursor of the point-1 pool
                atomicMax(&d.counters->max_world_m, run);
                atomicMax(&d.counters->max_world_k2, n2);
            }
        }
        __syncthreads();
        const uint32_t nc = s_nc;
        if (s_fits) {
            world_run(d, sm, w, b0, nb, p0, p1, nc, sub_dt, S, I, s_mom, s_pos, s_frc, s_inv, s_fm, s_cnt, s_beg, &s_n2);
        } else {   // the world's own slice of the global record arrays (see the header comment)
            WorldRecs rg;
            rg.hdr = (uint2*)d.s_dep + p0;
            rg.nfb = d.s_nf + p0;
            rg.r0 = d.s_r0 + p0;
            rg.ma0 = d.s_pm0 + p0;
            rg.r1 = d.s_r1 + p0;
            rg.ma1 = d.s_pm1 + p0;
            rg.b1 = (float*)d.s_acc1 + p0;
            world_run(d, rg, w, b0, nb, p0, p1, nc, sub_dt, S, I, s_mom, s_pos, s_frc, s_inv, s_fm, s_cnt, s_beg, &s_n2);
        }
        __syncthreads();   // the next world of this CTA reuses the shared arrays
    }
}

}  // namespace r2d
