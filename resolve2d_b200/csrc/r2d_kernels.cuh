// r2d_kernels.cuh — the __global__ kernels of the step pipeline for sm_100a.  Each wraps a `*_thread` body from
// r2d_pipeline.cuh in a grid-stride loop; the cooperative parts (big-body cell walks, device-wide scan, the
// Jones-Plassmann colouring rounds, warp-aggregated colour partition) are written here.
//
// Launch shape: every kernel whose element count lives in device memory (E, P are only known on the GPU) is launched
// on a fixed grid that is a multiple of the SM count and strides over the elements, so process() needs a single
// host<->device round trip per call (r2d_runtime.cu).  All of this is HBM/L2-bound integer and f32 work: no tensor
// cores, coalesced float4/uint4 SoA accesses, scattered body gathers served from the 126 MB L2.
#pragma once
#include <cooperative_groups.h>
#include <cooperative_groups/reduce.h>
#include <cuda_runtime.h>

#include "r2d_pipeline.cuh"

namespace r2d {
namespace cg = cooperative_groups;

constexpr int TPB = 256;            // threads per CTA of the streaming kernels
constexpr int SOLVE_TPB = 128;      // colour sweeps: small CTAs spread thin colours over all SMs
constexpr int BIG_LIST = 32;        // big bodies a CTA can defer per pass

__device__ __forceinline__ uint32_t live_entries(const Dev& d) {
    const uint32_t e = d.counters->n_entries;
    return e < d.cap_entries ? e : d.cap_entries;
}
__device__ __forceinline__ uint32_t live_pairs(const Dev& d) {
    const uint32_t p = d.counters->n_pairs;
    return p < d.cap_pairs ? p : d.cap_pairs;
}
// an over-capacity attempt is abandoned (the host grows the buffers and redoes the broadphase)
__device__ __forceinline__ bool overflowed(const Dev& d) {
    return d.counters->n_entries > d.cap_entries || d.counters->n_pairs > d.cap_pairs;
}

// ---- K2 / K4: per-body cell walk; FILL = false counts (SpatialHash.zig:46-49), true fills (:62-68) -------------------------
template <bool FILL>
__global__ void __launch_bounds__(TPB) k_grid_cells(Dev d) {
    __shared__ uint32_t big_list[BIG_LIST];
    __shared__ uint32_t n_big;
    if (threadIdx.x == 0) n_big = 0;
    __syncthreads();
    for (uint32_t base = blockIdx.x * blockDim.x; base < d.n_bodies; base += gridDim.x * blockDim.x) {
        const uint32_t i = base + threadIdx.x;
        if (i < d.n_bodies) {
            CellRange r;
            if (FILL)
                r = cell_range(d, i);
            else
                r = count_body_thread(d, i, false);
            bool inline_walk = r.count <= BIG_BODY_CELLS;
            if (!inline_walk) {
                const uint32_t slot = atomicAdd(&n_big, 1u);
                if (slot < BIG_LIST)
                    big_list[slot] = i;
                else
                    inline_walk = true;
            }
            if (inline_walk) {
                for (uint32_t k = 0; k < r.count; ++k) {
                    const uint32_t b = cell_bucket(r, k);
                    if (FILL)
                        fill_cell(d, i, b);
                    else
                        atomicAdd(&d.bucket_cnt[b], 1u);
                }
            }
        }
        __syncthreads();
        const uint32_t nb = n_big < BIG_LIST ? n_big : BIG_LIST;
        for (uint32_t q = 0; q < nb; ++q) {  // e.g. the floor: hundreds of cells, walked by the whole CTA
            const uint32_t bi = big_list[q];
            const CellRange r = cell_range(d, bi);
            for (uint32_t k = threadIdx.x; k < r.count; k += blockDim.x) {
                const uint32_t b = cell_bucket(r, k);
                if (FILL)
                    fill_cell(d, bi, b);
                else
                    atomicAdd(&d.bucket_cnt[b], 1u);
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) n_big = 0;
        __syncthreads();
    }
}

// ---- K3: device-wide exclusive scan of u32 (reduce / spine / down-sweep) --------------------------------------------------
constexpr int SCAN_TPB = 256;
constexpr int SCAN_IPT = 8;
constexpr int SCAN_TILE = SCAN_TPB * SCAN_IPT;

__device__ __forceinline__ uint32_t scan_count(const uint32_t* n_ptr, uint32_t n_max) {
    if (!n_ptr) return n_max;
    const uint32_t n = *n_ptr;
    return n < n_max ? n : n_max;
}
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total) {
    __shared__ uint32_t warp_sums[SCAN_TPB / 32];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (uint32_t)o) inc += t;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < SCAN_TPB / 32 ? warp_sums[lane] : 0u;
#pragma unroll
        for (int o = 1; o < SCAN_TPB / 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= (uint32_t)o) w += t;
        }
        if (lane < SCAN_TPB / 32) warp_sums[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    const uint32_t warp_base = warp ? warp_sums[warp - 1] : 0u;
    if (total) *total = warp_sums[SCAN_TPB / 32 - 1];
    const uint32_t r = warp_base + inc - v;
    __syncthreads();
    return r;
}
__global__ void __launch_bounds__(SCAN_TPB) k_scan_reduce(const uint32_t* __restrict__ in, uint32_t* __restrict__ tile_sums,
                                                          const uint32_t* n_ptr, uint32_t n_max) {
    const uint32_t n = scan_count(n_ptr, n_max);
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_IPT;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_IPT; ++k)
        if (base + k < n) s += in[base + k];
    uint32_t total;
    block_exclusive_scan(s, &total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}
// one CTA: exclusive scan of the tile sums, writes the grand total to out[n] and (optionally) to *total_out
__global__ void __launch_bounds__(SCAN_TPB) k_scan_spine(uint32_t* tile_sums, const uint32_t* n_ptr, uint32_t n_max,
                                                         uint32_t* out, uint32_t* total_out) {
    const uint32_t n = scan_count(n_ptr, n_max);
    const uint32_t n_tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    __shared__ uint32_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n_tiles; base += SCAN_TPB) {
        const uint32_t t = base + threadIdx.x;
        const uint32_t v = t < n_tiles ? tile_sums[t] : 0u;
        uint32_t total;
        const uint32_t ex = block_exclusive_scan(v, &total);
        const uint32_t carry = carry_s;
        if (t < n_tiles) tile_sums[t] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out[n] = carry_s;
        if (total_out) *total_out = carry_s;
    }
}
__global__ void __launch_bounds__(SCAN_TPB) k_scan_down(const uint32_t* in, uint32_t* out, const uint32_t* __restrict__ tile_sums,
                                                        const uint32_t* n_ptr, uint32_t n_max) {
    const uint32_t n = scan_count(n_ptr, n_max);
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_IPT;
    if (blockIdx.x * SCAN_TILE >= n) return;
    uint32_t v[SCAN_IPT];
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_IPT; ++k) {
        v[k] = (base + k < n) ? in[base + k] : 0u;
        s += v[k];
    }
    uint32_t run = block_exclusive_scan(s, nullptr) + tile_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_IPT; ++k) {
        if (base + k < n) out[base + k] = run;
        run += v[k];
    }
}

// ---- K4b: per-bucket sort (deterministic bucket order) -----------------------------------------------------------------------
__global__ void __launch_bounds__(TPB) k_sort_buckets(Dev d) {
    for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < d.n_buckets; b += gridDim.x * blockDim.x)
        sort_bucket_thread(d, b);
}

// ---- K5: candidate pairs per grid entry; WRITE = false counts, true writes at the scanned offsets ----------------------------
template <bool WRITE>
__global__ void __launch_bounds__(TPB) k_pairs(Dev d) {
    const uint32_t n = live_entries(d);
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        if (!WRITE) {
            d.ent_off[e] = entry_pairs_thread(d, e, nullptr);
        } else {
            const uint32_t off = d.ent_off[e], cnt = d.ent_off[e + 1] - off;
            if (cnt && off + cnt <= d.cap_pairs) entry_pairs_thread(d, e, d.pairs + off);
        }
    }
}

// ---- K6: narrowphase, one thread per candidate pair --------------------------------------------------------------------------
__global__ void __launch_bounds__(TPB) k_narrow(Dev d) {
    if (overflowed(d)) return;
    const uint32_t n = live_pairs(d);
    uint32_t my_m = 0, my_k = 0;
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
        const int np = narrow_pair_thread(d, p);
        if (np >= 0) {
            my_m += 1;
            my_k += (uint32_t)np;
        }
    }
    my_m = cg::reduce(cg::tiled_partition<32>(cg::this_thread_block()), my_m, cg::plus<uint32_t>());
    my_k = cg::reduce(cg::tiled_partition<32>(cg::this_thread_block()), my_k, cg::plus<uint32_t>());
    if ((threadIdx.x & 31u) == 0 && (my_m | my_k)) {
        atomicAdd(&d.counters->n_manifolds, my_m);
        atomicAdd(&d.counters->n_points, my_k);
    }
}

// ---- K8: graph colouring, one cooperative launch ---------------------------------------------------------------------------------
// Jones-Plassmann rounds over the pending manifolds with one grid barrier per round.  Round r reads the priorities
// posted for r in maxprio[r & 1] and posts those of the losers for r + 1 into the other array.  The result equals a
// sequential greedy colouring in descending priority, so it is a pure function of the contact graph and the body ids.
__global__ void __launch_bounds__(TPB) k_color(Dev d) {
    cg::grid_group grid = cg::this_grid();
    const bool dead = overflowed(d);
    const uint32_t n = dead ? 0u : live_pairs(d);
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    for (uint32_t p = tid; p < n; p += nth) {
        if (d.m_color[p] != COLOR_PENDING) continue;
        const uint4 h = d.m_hdr[p];
        color_post(d, h.x, h.y, !(body_flags(d, h.x) & FLAG_STATIC), !(body_flags(d, h.y) & FLAG_STATIC),
                   manifold_priority(d, h.x, h.y), 1u);
    }
    grid.sync();
    uint32_t round = 1;
    for (; round < MAX_COLOR_ROUNDS; ++round) {
        uint32_t left = 0;
        for (uint32_t p = tid; p < n; p += nth) {
            const int r = color_round_thread(d, p, round);
            if (r == 2) {
                left += 1;
            } else if (r == 1) {
                // per-colour population, aggregated over the lanes of the warp that won the same colour
                const uint32_t c = d.m_color[p];
                const uint32_t peers = __match_any_sync(__activemask(), c);
                if ((threadIdx.x & 31u) == (uint32_t)(__ffs(peers) - 1)) {
                    atomicAdd(&d.color_count[c], (uint32_t)__popc(peers));
                    atomicMax(&d.counters->n_colors, c + 1u);
                }
            }
        }
        left = cg::reduce(cg::tiled_partition<32>(cg::this_thread_block()), left, cg::plus<uint32_t>());
        if ((threadIdx.x & 31u) == 0 && left) atomicAdd(&d.round_left[round], left);
        grid.sync();
        if (__ldcg(&d.round_left[round]) == 0) break;
    }
    if (tid == 0) {
        d.counters->n_rounds = round;
        if (round >= MAX_COLOR_ROUNDS) atomicOr(&d.counters->err, ERR_ROUNDS);
        const uint32_t nc = __ldcg(&d.counters->n_colors);
        uint32_t run = 0;
        for (uint32_t c = 0; c < nc; ++c) {  // colour segments start on warp boundaries (dataflow sweep)
            d.color_start[c] = run;
            run = (run + __ldcg(&d.color_count[c]) + COLOR_ALIGN - 1u) & ~(COLOR_ALIGN - 1u);
        }
        d.color_start[nc] = run;
    }
}

// ---- K9: group the manifolds by colour and evaluate the pre-step (collision.zig:102-133) ---------------------------------------
__global__ void __launch_bounds__(TPB) k_partition_prestep(Dev d) {
    if (overflowed(d)) return;
    const uint32_t n = live_pairs(d);
    if (blockIdx.x == 0) {  // mark the padding slots at the end of every colour segment
        const uint32_t nc = d.counters->n_colors;
        for (uint32_t k = threadIdx.x; k < nc * COLOR_ALIGN; k += blockDim.x) {
            const uint32_t c = k / COLOR_ALIGN;
            const uint32_t at = d.color_start[c] + d.color_count[c] + (k % COLOR_ALIGN);
            if (at < d.color_start[c + 1]) d.s_hdr[at] = make_uint4(0u, 0u, S_EMPTY, 0u);
        }
    }
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
        const uint32_t c = d.m_color[p];
        if (c >= MAX_COLORS) continue;
        // one cursor bump per (warp, colour)
        const uint32_t peers = __match_any_sync(__activemask(), c);
        const uint32_t lane = threadIdx.x & 31u, leader = (uint32_t)(__ffs(peers) - 1);
        uint32_t base = 0;
        if (lane == leader) base = atomicAdd(&d.color_cursor[c], (uint32_t)__popc(peers));
        base = __shfl_sync(peers, base, leader);
        const uint32_t at = d.color_start[c] + base + (uint32_t)__popc(peers & ((1u << lane) - 1u));
        gather_prestep_thread(d, p, at);
    }
}

// ---- substep kernels (lib.zig:199-250) ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TPB) k_integrate_forces(Dev d, float sub_dt, int refresh_aabb) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < d.n_bodies; i += gridDim.x * blockDim.x)
        integrate_forces_thread(d, i, sub_dt, refresh_aabb != 0);
}
__global__ void __launch_bounds__(TPB) k_integrate_positions(Dev d, float sub_dt) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < d.n_bodies; i += gridDim.x * blockDim.x)
        integrate_positions_thread(d, i, sub_dt);
}
__global__ void __launch_bounds__(SOLVE_TPB) k_solve_contacts(Dev d, uint32_t begin, uint32_t end, float sub_dt) {
    const uint32_t m = begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (m < end) solve_contact_thread<false>(d, m, sub_dt);
}
__global__ void __launch_bounds__(SOLVE_TPB) k_solve_joints(Dev d, uint32_t begin, uint32_t end, float sub_dt) {
    const uint32_t j = begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (j < end) solve_joint_thread(d, j, sub_dt);
}

// ---- the whole substep loop of lib.zig:199-250 as ONE persistent cooperative kernel ---------------------------------------------
// Per substep:  [positions of the previous substep + forces, per body]  | grid barrier |
//               I x { joint colours (a barrier each) ; ONE dataflow sweep over all contact colours }  | grid barrier |
// Inside a sweep there is no barrier between colours, nor between iterations when the world has no joints: manifolds
// synchronise through the per-body version words (solve_contact_thread<true>).  Thread t owns manifolds t, t + nth, ...
// of the colour-sorted array, i.e. it walks its manifolds in ascending colour, which the dataflow order requires.
// Colour ranges and counts are read from device memory, so the host never has to learn them before launching.
// An abandoned attempt (buffer overflow, colouring error) leaves the body state untouched.
__device__ __forceinline__ void stamp(const Dev& d, uint32_t slot) {
    if (blockIdx.x == 0 && threadIdx.x == 0 && slot < 12) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        d.counters->stamp[slot] = t;
        d.counters->n_stamps = slot + 1;
    }
}
__global__ void __launch_bounds__(TPB) k_solve_persistent(Dev d, float sub_dt, uint32_t S, uint32_t I,
                                                          const uint32_t* __restrict__ joint_color_start, uint32_t n_joint_colors) {
    cg::grid_group grid = cg::this_grid();
    if (overflowed(d) || d.counters->err != 0u) return;  // uniform across the grid
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    const uint32_t n_manifolds = d.color_start[d.counters->n_colors];
    stamp(d, 0);
    for (uint32_t s = 0; s < S; ++s) {
        if (s == 0)
            for (uint32_t i = tid; i < d.n_bodies; i += nth) integrate_forces_thread(d, i, sub_dt, S == 1);
        if (s == 0) stamp(d, 1);
        grid.sync();
        if (s == 0) stamp(d, 2);
        for (uint32_t it = 0; it < I; ++it) {
            for (uint32_t jc = 0; jc < n_joint_colors; ++jc) {
                const uint32_t b = joint_color_start[jc], e = joint_color_start[jc + 1];
                for (uint32_t j = b + tid; j < e; j += nth) solve_joint_thread(d, j, sub_dt);
                grid.sync();
            }
            for (uint32_t m = tid; m < n_manifolds; m += nth) solve_contact_thread<true>(d, m, sub_dt, it);
            if (s == 0) stamp(d, 3 + it);     // block 0 finished its part of sweep `it`
            if (n_joint_colors) grid.sync();  // joints of the next iteration read what the contacts wrote
        }
        if (!n_joint_colors) grid.sync();
        if (s == 0) stamp(d, 8);
        // end of substep s fused with the start of substep s + 1: both are per-body, same thread, no barrier needed
        for (uint32_t i = tid; i < d.n_bodies; i += nth) {
            integrate_positions_thread(d, i, sub_dt);
            if (s + 1 < S) integrate_forces_thread(d, i, sub_dt, s + 2 == S);
        }
        if (s == 0) stamp(d, 9);
    }
    stamp(d, 10);
}

// ---- boundary kernels: SoA export for bulk readback, force import ---------------------------------------------------------------
__global__ void __launch_bounds__(TPB) k_export_bodies(Dev d, uint32_t first, uint32_t n, uint32_t* ids, float2* pos, float* angle,
                                                       float2* mom, float* ang_mom, float4* aabb) {
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const uint32_t s = first + k;
        if (ids) ids[k] = body_id(d, s);
        if (pos || angle) {
            const float4 p = d.pos[s];
            if (pos) pos[k] = make_float2(p.x, p.y);
            if (angle) angle[k] = p.z;
        }
        if (mom || ang_mom) {
            const float4 m = d.mom[s];
            if (mom) mom[k] = make_float2(m.x, m.y);
            if (ang_mom) ang_mom[k] = m.z;
        }
        if (aabb) aabb[k] = d.aabb[s];
    }
}
__global__ void __launch_bounds__(TPB) k_import_forces(Dev d, uint32_t first, uint32_t n, const float* __restrict__ fxy_t) {
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const uint32_t s = first + k;
        d.frc[s] = make_float4(fxy_t[3 * k], fxy_t[3 * k + 1], fxy_t[3 * k + 2], d.frc[s].w);
    }
}

}  // namespace r2d
