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
#ifndef NARROW_CTAS
#define NARROW_CTAS 2   // 123 registers, no spills: 6 % faster than 3 CTAs per SM at 80 registers with spills (instruction-bound)
#endif
constexpr int SOLVE_TPB = 128;      // colour sweeps: small CTAs spread thin colours over all SMs
constexpr int BIG_LIST = 32;        // big bodies a CTA can defer per pass
constexpr int WORLD_TPB = 128;      // CTA-per-world kernels (batches of small worlds)

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
    return d.counters->n_entries > d.cap_entries || d.counters->n_pairs > d.cap_pairs || d.counters->broad_fallback != 0u;
}

// ---- K2 / K4: per-body cell walk; FILL = false counts (SpatialHash.zig:46-49), true fills (:62-68) -------------------------
// Bodies that cover more than BIG_BODY_CELLS cells (the floor of a 100k-body pile: 1,100 cells) are not walked by the thread
// that holds them.  The count pass lets its whole CTA walk them and appends them to a device-wide list; the fill pass
// spreads the cells of every listed body over the WHOLE grid (one CTA walking 1,100 cells with a dependent atomic and
// three stores each was the tail of the fill kernel: 19.7 of its 20 us).
constexpr uint32_t BIG_GLOBAL_LIST = 1024;   // bodies the device-wide list holds (more: walked by their CTA, as in the count pass)
template <bool FILL>
__global__ void __launch_bounds__(TPB) k_grid_cells(Dev d) {
    __shared__ uint32_t big_list[BIG_LIST];
    __shared__ uint32_t n_big;
    if (threadIdx.x == 0) n_big = 0;
    __syncthreads();
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < d.n_bodies; i += gridDim.x * blockDim.x) {
        CellRange r;
        if (!FILL) {
            r = count_body_thread(d, i, false);
        } else if (body_is_small(d, body_flags(d, i))) {
            fill_fine(d, i);
            r.count = 0;
        } else {
            r = cell_range(d, i);
        }
        bool inline_walk = r.count <= BIG_BODY_CELLS;
        if (!inline_walk) {  // e.g. the floor: hundreds of cells
            if (FILL) {      // on the device-wide list (complete: written by the count kernel)? then the whole grid walks it below
                bool listed = false;
                const uint32_t n_global = min(d.counters->n_big, BIG_GLOBAL_LIST);
                for (uint32_t q = 0; q < n_global; ++q) listed = listed || d.big_bodies[q] == i;
                if (listed) continue;
            } else {
                const uint32_t g = atomicAdd(&d.counters->n_big, 1u);
                if (g < BIG_GLOBAL_LIST) d.big_bodies[g] = i;
            }
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
    for (uint32_t q = 0; q < nb; ++q) {
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
    if (FILL) {
        const uint32_t n_global = min(d.counters->n_big, BIG_GLOBAL_LIST);
        for (uint32_t q = 0; q < n_global; ++q) {
            const uint32_t bi = d.big_bodies[q];
            const CellRange r = cell_range(d, bi);
            for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < r.count; k += gridDim.x * blockDim.x)
                fill_cell(d, bi, cell_bucket(r, k));
        }
    }
}

// ---- K3: device-wide exclusive scan of u32 ----------------------------------------------------------------------------------
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
// Single-pass chained scan (decoupled look-back): one launch instead of three.  Tile ids come from an atomic ticket,
// so a CTA only ever waits for tiles that started before it; every tile publishes (flag, value) as ONE 64-bit word
// (flag 1 = aggregate of the tile, 2 = inclusive prefix), which needs no fence.  `state` (n_tiles + 1 words, the last
// one is the ticket) must be zero on entry.  in == out is allowed.
// Returns false when the ticket drawn is beyond the last tile (nothing done).
template <class Load>
__device__ __forceinline__ bool scan_chained_body(Load load, uint32_t* out, uint32_t n, unsigned long long* state, uint32_t state_tiles,
                                                  uint32_t* total_out, uint32_t* first_cta) {
    __shared__ uint32_t s_tile, s_prefix;
    if (threadIdx.x == 0) s_tile = (uint32_t)atomicAdd(&state[state_tiles], 1ull);
    __syncthreads();
    const uint32_t tile = s_tile;
    if (first_cta) *first_cta = tile == 0u ? 1u : 0u;
    const uint32_t n_tiles = n ? (n + SCAN_TILE - 1) / SCAN_TILE : 1u;
    if (tile >= n_tiles) return false;
    const uint32_t base = tile * SCAN_TILE + threadIdx.x * SCAN_IPT;
    uint32_t v[SCAN_IPT];
    uint32_t sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_IPT; ++k) {
        v[k] = (base + k < n) ? load(base + k) : 0u;
        sum += v[k];
    }
    uint32_t aggregate;
    const uint32_t local = block_exclusive_scan(sum, &aggregate);
    if (threadIdx.x < 32u) {
        // decoupled look-back by one warp: lane j inspects tile (first - j); status 2 = inclusive prefix, 1 = aggregate only
        const uint32_t lane = threadIdx.x;
        uint32_t prefix = 0;
        if (tile == 0) {
            if (lane == 0) atomicExch(&state[0], (2ull << 32) | aggregate);
        } else {
            if (lane == 0) atomicExch(&state[tile], (1ull << 32) | aggregate);
            int first = (int)tile - 1;
            for (;;) {
                const int t = first - (int)lane;
                unsigned long long w = 3ull << 32;  // beyond tile 0: neutral, counts as "ready"
                if (t >= 0) {
                    do {
                        w = *((volatile unsigned long long*)&state[t]);
                    } while ((w >> 32) == 0ull);
                }
                const uint32_t is_prefix = __ballot_sync(0xffffffffu, (w >> 32) == 2ull);
                const uint32_t upto = is_prefix ? (uint32_t)__ffs((int)is_prefix) - 1u : 31u;  // lanes 0..upto contribute
                uint32_t v = (t >= 0 && lane <= upto) ? (uint32_t)w : 0u;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                prefix += v;
                if (is_prefix || first < 32) break;
                first -= 32;
            }
            if (lane == 0) atomicExch(&state[tile], (2ull << 32) | (unsigned long long)(uint32_t)(prefix + aggregate));
        }
        if (lane == 0) {
            s_prefix = prefix;
            if (tile == n_tiles - 1) {
                out[n] = prefix + aggregate;
                if (total_out) *total_out = prefix + aggregate;
            }
        }
    }
    __syncthreads();
    uint32_t run = s_prefix + local;
#pragma unroll
    for (int k = 0; k < SCAN_IPT; ++k) {
        if (base + k < n) out[base + k] = run;
        run += v[k];
    }
    return true;
}
__global__ void __launch_bounds__(SCAN_TPB) k_scan_chained(const uint32_t* in, uint32_t* out, const uint32_t* n_ptr, uint32_t n_max,
                                                           unsigned long long* state, uint32_t state_tiles, uint32_t* total_out) {
    scan_chained_body([in](uint32_t i) { return in[i]; }, out, scan_count(n_ptr, n_max), state, state_tiles, total_out, nullptr);
}
// The scan that places the manifolds (see manifold_owner): its input — the popcount of every owner-bitmap word, plus one
// entry per colour that pads the colour's segment to a whole warp — is computed on the fly (no k_owner_count launch), and
// so is the number of colours (no k_color_finish launch after the per-world colouring): every CTA derives it from the
// colour populations, the CTA with ticket 0 publishes it for the kernels that follow.
__global__ void __launch_bounds__(SCAN_TPB) k_scan_owners(Dev d, unsigned long long* state, uint32_t state_tiles) {
    __shared__ uint32_t s_nc;
    if (threadIdx.x == 0) s_nc = 0u;
    __syncthreads();
    for (uint32_t c = threadIdx.x; c < MAX_COLORS; c += blockDim.x)
        if (d.color_count[c]) atomicMax(&s_nc, c + 1u);
    __syncthreads();
    const uint32_t nc = s_nc, stride = d.own_words + 1u;
    const uint32_t n = overflowed(d) ? 0u : nc * stride;
    uint32_t first = 0;
    scan_chained_body(
        [&d, stride](uint32_t k) {
            const uint32_t c = k / stride, w = k % stride;
            if (w < d.own_words) return (uint32_t)__popc(d.own_bits[(size_t)c * d.own_words + w]);
            const uint32_t cnt = d.color_count[c];
            return (COLOR_ALIGN - (cnt % COLOR_ALIGN)) % COLOR_ALIGN;
        },
        d.own_pos, n, state, state_tiles, nullptr, &first);
    if (first && threadIdx.x == 0) {
        d.counters->n_colors = nc;
        d.counters->n_own_scan = nc * stride;
    }
}

// ---- K4b / K5: one WARP per bucket ------------------------------------------------------------------------------------------
// The bucket's entries are staged in shared memory.  k_sort_buckets rank-sorts them (ascending slot, duplicates
// adjacent) so that the bucket order — and with it the pair list — does not depend on the order the atomics of the
// fill landed in, and so that membership tests can bisect.  k_bucket_pairs then lets the lanes split the (a, b) tests
// of the bucket; a pair is counted (WRITE = false) or written at the bucket's scanned offset (WRITE = true) in
// (a, b) order, which makes the candidate list deterministic.
constexpr int BUCKET_WARPS = TPB / 32;
constexpr int BUCKET_CAP = 256;    // entries staged in shared memory; larger buckets read their entries from global memory
constexpr int SMALL_BUCKET = 16;   // up to here: a warp per bucket
constexpr int HEAVY_BUCKET = 40;   // above: a whole CTA per bucket (n^2 pair tests).  In between ("medium"): a CTA each when
                                   // there are too few of them to fill the GPU with warps (a dense pile), else a warp each
                                   // (thousands of small worlds)
constexpr uint32_t HIT_WORDS_PER_ENTRY = 4;  // ballot words kept per grid entry: n (n - 1) / 64 <= 4 n for n <= 257

// work lists: small buckets from the front of work[0, T), heavy ones from its back, medium ones in work[T, 2T)
enum { LIST_SMALL = 0, LIST_MEDIUM = 1, LIST_HEAVY = 2 };
template <int LIST>
__device__ __forceinline__ uint32_t list_count(const Dev& d) {
    return LIST == LIST_HEAVY ? d.counters->n_heavy : (LIST == LIST_MEDIUM ? d.counters->n_mid : d.counters->n_work);
}
template <int LIST>
__device__ __forceinline__ uint32_t list_bucket(const Dev& d, uint32_t w) {
    return LIST == LIST_HEAVY ? d.work[d.n_buckets - 1u - w] : (LIST == LIST_MEDIUM ? d.work[d.n_buckets + w] : d.work[w]);
}
// one THREAD per bucket: clear its pair counter and list it if it can produce pairs
__global__ void __launch_bounds__(TPB) k_list_buckets(Dev d) {
    for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < d.n_buckets; b += gridDim.x * blockDim.x) {
        const uint32_t bs = d.bucket_start[b], be = bucket_end(d, b);
        const uint32_t n = be > bs ? be - bs : 0u;
        d.ent_off[b] = 0u;
        if (n > (uint32_t)HEAVY_BUCKET)
            d.work[d.n_buckets - 1u - atomicAdd(&d.counters->n_heavy, 1u)] = b;
        else if (n > (uint32_t)SMALL_BUCKET)
            d.work[d.n_buckets + atomicAdd(&d.counters->n_mid, 1u)] = b;
        else if (n >= 2u)
            d.work[atomicAdd(&d.counters->n_work, 1u)] = b;
    }
}
// one WARP per listed bucket
__global__ void __launch_bounds__(TPB) k_sort_buckets(Dev d) {
    __shared__ uint32_t s_in[BUCKET_WARPS][BUCKET_CAP];
    const uint32_t lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t n_small = d.counters->n_work, n_mid = d.counters->n_mid, n_items = n_small + n_mid + d.counters->n_heavy;
    for (uint32_t w = warp; w < n_items; w += n_warps) {
        const uint32_t b = w < n_small ? list_bucket<LIST_SMALL>(d, w)
                                       : (w < n_small + n_mid ? list_bucket<LIST_MEDIUM>(d, w - n_small)
                                                              : list_bucket<LIST_HEAVY>(d, w - n_small - n_mid));
        const uint32_t bs = d.bucket_start[b], be = bucket_end(d, b);
        const uint32_t n = be > bs ? be - bs : 0u;
        if (n < 2u) continue;  // warp-uniform
        if (n > BUCKET_CAP) {
            if (lane == 0) sort_bucket_thread(d, b);
            continue;
        }
        for (uint32_t k = lane; k < n; k += 32) s_in[wib][k] = d.ent_body[bs + k];
        __syncwarp();
        for (uint32_t k = lane; k < n; k += 32) {
            const uint32_t v = s_in[wib][k];
            uint32_t rank = 0;
            for (uint32_t q = 0; q < n; ++q) {
                const uint32_t u = s_in[wib][q];
                rank += (u < v || (u == v && q < k)) ? 1u : 0u;
            }
            d.ent_body[bs + rank] = v;
        }
        __syncwarp();
    }
}

// (a, b), a < b, of the p-th pair in row-major order over the strict upper triangle of an n x n matrix
__device__ __forceinline__ void tri_decode(uint32_t p, uint32_t n, uint32_t& a, uint32_t& b) {
    const float fn = (float)(2 * n - 1);
    int ia = (int)((fn - sqrtf(fn * fn - 8.0f * (float)p)) * 0.5f);
    if (ia < 0) ia = 0;
    // row a starts at a*(2n-a-1)/2; correct the float estimate
    while ((uint32_t)(ia + 1) * (2 * n - (uint32_t)(ia + 1) - 1) / 2 <= p) ++ia;
    while ((uint32_t)ia * (2 * n - (uint32_t)ia - 1) / 2 > p) --ia;
    a = (uint32_t)ia;
    b = p - a * (2 * n - a - 1) / 2 + a + 1;
}

// advance (a, b) by `step` positions in that order (the caller guarantees the target index is < n(n-1)/2)
__device__ __forceinline__ void tri_advance(uint32_t n, uint32_t step, uint32_t& a, uint32_t& b) {
    b += step;
    while (b >= n) {
        b -= n;
        a += 1u;
        b += a + 1u;
    }
}

// One bucket, worked on by a TEAM of lanes: a warp (light buckets, 8 buckets per CTA) or the whole CTA (heavy buckets).
// Pair tests are taken in row-major order of the strict upper triangle, 32 consecutive tests per warp iteration; every
// warp of the team owns a contiguous range of test indices.  The COUNT pass evaluates the tests, stores the bucket's
// hit count in ent_off[b] and — for buckets staged in shared memory (n <= BUCKET_CAP) — the 32-test ballots in
// hit_bits[4 * bucket_start + test / 32]  (n (n - 1) / 64 <= 4 n words for n <= 257).  After the scan of ent_off the
// WRITE pass only has to expand those ballots: one word per team lane, a team-wide exclusive scan of the popcounts,
// and a tri_decode per HIT instead of per test (hits are ~6 % of the tests in a dense pile).  Oversized buckets
// (n > BUCKET_CAP) keep no ballots and repeat their tests in the write pass.
struct BucketItem {
    uint32_t b, bs, n;
};
template <int LIST>
__device__ __forceinline__ BucketItem bucket_item(const Dev& d, uint32_t w, bool dead) {
    BucketItem it;
    it.b = list_bucket<LIST>(d, w);
    it.bs = d.bucket_start[it.b];
    const uint32_t be = bucket_end(d, it.b);
    it.n = (be > it.bs && !dead) ? be - it.bs : 0u;
    return it;
}

// tests [p_begin, p_end) of an oversized bucket straight from global memory; returns the warp's hit count and, when
// `out_at` is valid, writes the hits at out_at, out_at + 1, ... in test order
__device__ __forceinline__ uint32_t bucket_tests_global(const Dev& d, const BucketItem& it, uint32_t p_begin, uint32_t p_end,
                                                        uint32_t lane, bool writing, uint32_t out_at) {
    uint32_t total = 0;
    for (uint32_t base = p_begin; base < p_end; base += 32) {
        const uint32_t p = base + lane;
        bool hit = false;
        uint2 pr = make_uint2(0u, 0u);
        if (p < p_end) {
            uint32_t a, k;
            tri_decode(p, it.n, a, k);
            const uint32_t i = d.ent_body[it.bs + a], j = d.ent_body[it.bs + k];
            const bool fa = a == 0 || d.ent_body[it.bs + a - 1] != i, fb = d.ent_body[it.bs + k - 1] != j;
            if (fa && fb && i != j)
                hit = pair_candidate(d, it.b, i, j, d.aabb[i], d.aabb[j], body_flags(d, i), body_flags(d, j), d.ncells[i],
                                     d.ncells[j], d.bkt[i], d.bkt[j], &pr);
        }
        const uint32_t votes = __ballot_sync(0xffffffffu, hit);
        if (writing && hit) {
            const uint32_t at = out_at + total + (uint32_t)__popc(votes & ((1u << lane) - 1u));
            if (at < d.cap_pairs) d.pairs[at] = pr;
        }
        total += (uint32_t)__popc(votes);
    }
    return total;
}

// LIST: which work list; HEAVY: the team is the whole CTA (else a warp)
template <int LIST, bool HEAVY>
__device__ __forceinline__ void bucket_count_part(const Dev& d) {
    constexpr int TEAMS = HEAVY ? 1 : BUCKET_WARPS;       // teams per CTA
    constexpr uint32_t TEAM = HEAVY ? TPB : 32;           // lanes per team
    constexpr uint32_t TEAM_WARPS = TEAM / 32;
    // entries a team can stage (the lists bound the bucket sizes)
    constexpr int CAP = LIST == LIST_HEAVY ? BUCKET_CAP : (LIST == LIST_MEDIUM ? HEAVY_BUCKET : SMALL_BUCKET);
    __shared__ uint32_t s_body[TEAMS][CAP];
    __shared__ uint32_t s_meta[TEAMS][CAP];  // flags (bit 0 static) | first-occurrence << 1 | ncells << 2
    __shared__ float4 s_aabb[TEAMS][CAP];
    __shared__ uint4 s_bkt[TEAMS][CAP];
    __shared__ uint32_t s_warp_total[BUCKET_WARPS];
    const uint32_t lane = threadIdx.x & 31u, warp_in_cta = threadIdx.x >> 5;
    const uint32_t tc = HEAVY ? 0u : warp_in_cta;                    // team inside the CTA
    const uint32_t tl = HEAVY ? threadIdx.x : lane;                  // lane inside the team
    const uint32_t tw = HEAVY ? warp_in_cta : 0u;                    // warp inside the team
    const uint32_t team = HEAVY ? blockIdx.x : (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t n_teams = HEAVY ? gridDim.x : (gridDim.x * blockDim.x) >> 5;
    const bool dead = d.counters->n_entries > d.cap_entries;  // the fill dropped entries: this attempt is redone
    const uint32_t n_items = list_count<LIST>(d);
    for (uint32_t w = team; w < n_items; w += n_teams) {
        const BucketItem it = bucket_item<LIST>(d, w, dead);
        const uint32_t n = it.n;
        const bool staged = n <= (uint32_t)CAP;
        if (staged) {
            for (uint32_t k = tl; k < n; k += TEAM) {
                const uint32_t body = d.ent_body[it.bs + k];
                const bool first = k == 0 || d.ent_body[it.bs + k - 1] != body;
                s_body[tc][k] = body;
                s_aabb[tc][k] = d.aabb[body];
                s_bkt[tc][k] = d.bkt[body];
                uint32_t nc = d.ncells[body];
                if (nc > 0x3FFFFFFFu) nc = 0x3FFFFFFFu;
                s_meta[tc][k] = (body_flags(d, body) & FLAG_STATIC) | (first ? 2u : 0u) | (nc << 2);
            }
        }
        if (HEAVY) __syncthreads(); else __syncwarp();
        const uint32_t n_pairs = n >= 2u ? n * (n - 1u) / 2u : 0u;
        // contiguous range of test indices of this warp (a multiple of 32, so that ballot words stay aligned)
        const uint32_t per_warp = ((n_pairs + TEAM_WARPS * 32u - 1u) / (TEAM_WARPS * 32u)) * 32u;
        const uint32_t p_begin = tw * per_warp, p_end = (p_begin + per_warp < n_pairs) ? p_begin + per_warp : n_pairs;
        uint32_t total = 0;
        if (!staged) {
            total = bucket_tests_global(d, it, p_begin, p_end, lane, false, 0u);
        } else if (p_begin < p_end) {
            uint32_t* bits = d.hit_bits + HIT_WORDS_PER_ENTRY * (size_t)it.bs;
            uint32_t a = 0, k = 0;
            if (p_begin + lane < p_end) tri_decode(p_begin + lane, n, a, k);
            for (uint32_t base = p_begin; base < p_end; base += 32) {
                const uint32_t p = base + lane;
                bool hit = false;
                if (p < p_end) {
                    const uint32_t ma = s_meta[tc][a], mb = s_meta[tc][k];
                    const uint32_t i = s_body[tc][a], j = s_body[tc][k];
                    uint2 pr;
                    if ((ma & mb & 2u) && i != j)
                        hit = pair_candidate(d, it.b, i, j, s_aabb[tc][a], s_aabb[tc][k], ma & 1u, mb & 1u, ma >> 2, mb >> 2,
                                             s_bkt[tc][a], s_bkt[tc][k], &pr);
                    if (p + 32u < p_end) tri_advance(n, 32u, a, k);
                }
                const uint32_t votes = __ballot_sync(0xffffffffu, hit);
                if (lane == 0) bits[base >> 5] = votes;
                total += (uint32_t)__popc(votes);
            }
        }
        if (!HEAVY) {
            if (lane == 0) d.ent_off[it.b] = total;
        } else {
            if (lane == 0) s_warp_total[tw] = total;
            __syncthreads();
            if (threadIdx.x == 0) {
                uint32_t sum = 0;
                for (uint32_t q = 0; q < TEAM_WARPS; ++q) sum += s_warp_total[q];
                d.ent_off[it.b] = sum;
            }
        }
        if (HEAVY) __syncthreads(); else __syncwarp();
    }
}

// heavy buckets first (a CTA each: the long poles), then the light ones (a warp each) fill the tail of the same launch
// medium buckets: CTA teams when a warp each would leave most of the GPU idle
__device__ __forceinline__ bool medium_by_cta(const Dev& d) { return d.counters->n_mid < 4u * gridDim.x; }

__global__ void __launch_bounds__(TPB) k_bucket_count(Dev d) {
    bucket_count_part<LIST_HEAVY, true>(d);
    __syncthreads();
    if (medium_by_cta(d))
        bucket_count_part<LIST_MEDIUM, true>(d);
    else
        bucket_count_part<LIST_MEDIUM, false>(d);
    __syncthreads();
    bucket_count_part<LIST_SMALL, false>(d);
}

template <int LIST, bool HEAVY>
__device__ __forceinline__ void bucket_write_part(const Dev& d) {
    constexpr int TEAMS = HEAVY ? 1 : BUCKET_WARPS;
    constexpr uint32_t TEAM = HEAVY ? TPB : 32;
    constexpr uint32_t TEAM_WARPS = TEAM / 32;
    constexpr int CAP = LIST == LIST_HEAVY ? BUCKET_CAP : (LIST == LIST_MEDIUM ? HEAVY_BUCKET : SMALL_BUCKET);
    __shared__ uint32_t s_body[TEAMS][CAP];
    __shared__ uint32_t s_nc[TEAMS][CAP];
    __shared__ uint32_t s_warp_total[BUCKET_WARPS];
    const uint32_t lane = threadIdx.x & 31u, warp_in_cta = threadIdx.x >> 5;
    const uint32_t tc = HEAVY ? 0u : warp_in_cta;
    const uint32_t tl = HEAVY ? threadIdx.x : lane;
    const uint32_t tw = HEAVY ? warp_in_cta : 0u;
    const uint32_t team = HEAVY ? blockIdx.x : (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t n_teams = HEAVY ? gridDim.x : (gridDim.x * blockDim.x) >> 5;
    const bool dead = d.counters->n_entries > d.cap_entries;
    const uint32_t n_items = list_count<LIST>(d);
    for (uint32_t w = team; w < n_items; w += n_teams) {
        const BucketItem it = bucket_item<LIST>(d, w, dead);
        const uint32_t n = it.n;
        const uint32_t n_pairs = n >= 2u ? n * (n - 1u) / 2u : 0u;
        const uint32_t out_base = d.ent_off[it.b];  // scanned: first pair slot of this bucket
        if (n <= (uint32_t)CAP) {
            for (uint32_t k = tl; k < n; k += TEAM) {
                const uint32_t body = d.ent_body[it.bs + k];
                s_body[tc][k] = body;
                s_nc[tc][k] = d.ncells[body];
            }
            // one ballot word per team lane and chunk (light buckets: <= 4 words; heavy ones: <= 1020, 256 per chunk)
            const uint32_t n_words = (n_pairs + 31u) >> 5;
            uint32_t running = out_base;
            for (uint32_t w0 = 0; w0 < n_words; w0 += TEAM) {
                const uint32_t wi = w0 + tl;
                const uint32_t word = wi < n_words ? d.hit_bits[HIT_WORDS_PER_ENTRY * (size_t)it.bs + wi] : 0u;
                const uint32_t cnt = (uint32_t)__popc(word);
                uint32_t at = cnt;  // inclusive scan over the team, then made exclusive
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, at, o);
                    if (lane >= (uint32_t)o) at += t;
                }
                uint32_t chunk_total = 0;
                if (HEAVY) {
                    if (lane == 31) s_warp_total[tw] = at;
                    __syncthreads();  // also orders the staging of s_body / s_nc before the first expansion
                    for (uint32_t q = 0; q < TEAM_WARPS; ++q) {
                        if (q < tw) at += s_warp_total[q];
                        chunk_total += s_warp_total[q];
                    }
                } else {
                    __syncwarp();
                }
                at = running + at - cnt;
                uint32_t rest = word;
                while (rest) {
                    const uint32_t bit = (uint32_t)__ffs((int)rest) - 1u;
                    rest &= rest - 1u;
                    uint32_t a, k;
                    tri_decode(wi * 32u + bit, n, a, k);
                    const uint32_t i = s_body[tc][a], j = s_body[tc][k];
                    const uint32_t nci = s_nc[tc][a], ncj = s_nc[tc][k];
                    const bool i_owns = nci < ncj || (nci == ncj && i < j);  // same orientation as pair_candidate
                    if (at < d.cap_pairs) d.pairs[at] = i_owns ? make_uint2(i, j) : make_uint2(j, i);
                    ++at;
                }
                running += chunk_total;
                if (HEAVY) __syncthreads();  // s_warp_total is reused by the next chunk
            }
        } else {
            const uint32_t per_warp = ((n_pairs + TEAM_WARPS * 32u - 1u) / (TEAM_WARPS * 32u)) * 32u;
            const uint32_t p_begin = tw * per_warp, p_end = (p_begin + per_warp < n_pairs) ? p_begin + per_warp : n_pairs;
            uint32_t out_at = out_base;
            if (HEAVY) {  // count first to learn where the warps of the team write
                const uint32_t total = bucket_tests_global(d, it, p_begin, p_end, lane, false, 0u);
                if (lane == 0) s_warp_total[tw] = total;
                __syncthreads();
                for (uint32_t q = 0; q < tw; ++q) out_at += s_warp_total[q];
            }
            bucket_tests_global(d, it, p_begin, p_end, lane, true, out_at);
        }
        if (HEAVY) __syncthreads(); else __syncwarp();
    }
}

__global__ void __launch_bounds__(TPB) k_bucket_write(Dev d) {
    bucket_write_part<LIST_HEAVY, true>(d);
    __syncthreads();
    if (medium_by_cta(d))
        bucket_write_part<LIST_MEDIUM, true>(d);
    else
        bucket_write_part<LIST_MEDIUM, false>(d);
    __syncthreads();
    bucket_write_part<LIST_SMALL, false>(d);
}

// ---- K5 with the fine grid: one thread per small body (see fine_body_pairs) ------------------------------------------------------
// WRITE = false runs the tests, counts (pair_cnt[a + 1]) and parks the first 8 partners of the body in `fine_cand`; after
// the scan WRITE = true only sorts the parked partners (ascending slot: the order of the list must not depend on the
// order the fill's atomics landed in) and emits them at pair_cnt[a + 1].  A body with more than 8 partners re-runs the tests.
template <bool WRITE>
__global__ void __launch_bounds__(TPB) k_fine_pairs(Dev d) {
    if (overflowed(d)) return;
    if (!WRITE && blockIdx.x == 0 && threadIdx.x == 0) d.pair_cnt[0] = d.ll_on ? d.ent_off[d.n_buckets] : 0u;
    for (uint32_t a = blockIdx.x * blockDim.x + threadIdx.x; a < d.n_bodies; a += gridDim.x * blockDim.x) {
        uint32_t got[8];
        if (!WRITE) {
            const bool live = body_is_small(d, body_flags(d, a));
            const uint32_t n = live ? fine_body_pairs(d, a, got, nullptr) : 0u;
            d.pair_cnt[a + 1] = n;
            uint4* park = d.fine_cand + 2 * (size_t)a;
            if (n > 0u) park[0] = make_uint4(got[0], n > 1u ? got[1] : 0u, n > 2u ? got[2] : 0u, n > 3u ? got[3] : 0u);
            if (n > 4u) park[1] = make_uint4(got[4], n > 5u ? got[5] : 0u, n > 6u ? got[6] : 0u, n > 7u ? got[7] : 0u);
        } else {
            const uint32_t at = d.pair_cnt[a + 1], n = d.pair_cnt[a + 2] - at;
            if (n == 0u || at + n > d.cap_pairs) continue;
            uint2* out = d.pairs + at;
            if (n <= 8u) {
                const uint4* park = d.fine_cand + 2 * (size_t)a;
                const uint4 q0 = park[0], q1 = n > 4u ? park[1] : make_uint4(0u, 0u, 0u, 0u);
                got[0] = q0.x; got[1] = q0.y; got[2] = q0.z; got[3] = q0.w;
                got[4] = q1.x; got[5] = q1.y; got[6] = q1.z; got[7] = q1.w;
#pragma unroll
                for (int x = 1; x < 8; ++x) {  // insertion sort in registers (fixed trip counts: no local memory)
#pragma unroll
                    for (int y = x; y > 0; --y)
                        if ((uint32_t)x < n && got[y - 1] > got[y]) {
                            const uint32_t t = got[y - 1];
                            got[y - 1] = got[y];
                            got[y] = t;
                        }
                }
#pragma unroll
                for (int x = 0; x < 8; ++x)
                    if ((uint32_t)x < n) out[x] = make_uint2(a, got[x]);
            } else {
                fine_body_pairs(d, a, nullptr, out);
                sort_item_pairs(out, n);
            }
        }
    }
}

// ---- K6: narrowphase, one thread per candidate pair --------------------------------------------------------------------------
template <uint32_t NARROW_PER_THREAD>
__global__ void __launch_bounds__(TPB, NARROW_CTAS) k_narrow(Dev d) {
    if (overflowed(d)) return;
    const uint32_t n = live_pairs(d);
    uint32_t my_m = 0, my_k = 0;
    // Each CTA takes NARROW_TILE consecutive pairs and re-deals them to its threads sorted by shape combination IN ID
    // ORDER (disc-disc, disc-rect, rect-disc, rect-rect; "first" = the lower id, whose axes the first SAT pass tests), so
    // that a warp runs one SAT flavour with the same trip counts in both passes instead of all of them (a disc owns one
    // axis, a rectangle four: mixing disc-rect with rect-disc halves the lanes in both passes).  The result of a pair
    // does not depend on which thread computes it.
    constexpr uint32_t NARROW_TILE = NARROW_PER_THREAD * TPB;   // 1, 2 or 4 pairs per thread: the host picks by pair count
    __shared__ uint32_t s_idx[NARROW_TILE];
    __shared__ uint32_t s_cnt[4];
    for (uint32_t base = blockIdx.x * NARROW_TILE; base < n; base += gridDim.x * NARROW_TILE) {
        if (threadIdx.x < 4) s_cnt[threadIdx.x] = 0u;
        __syncthreads();
        uint32_t t[NARROW_PER_THREAD], rank[NARROW_PER_THREAD];
#pragma unroll
        for (uint32_t k = 0; k < NARROW_PER_THREAD; ++k) {
            const uint32_t p = base + k * TPB + threadIdx.x;
            t[k] = 4u;
            if (p < n) {
                const uint2 pr = d.pairs[p];
                const float4 sx = d.shape[pr.x], sy = d.shape[pr.y];
                const uint32_t rx = (f2u(sx.z) & FLAG_RECT) ? 1u : 0u, ry = (f2u(sy.z) & FLAG_RECT) ? 1u : 0u;
                const bool x_lo = f2u(sx.w) < f2u(sy.w);
                t[k] = x_lo ? (rx * 2u + ry) : (ry * 2u + rx);
            }
        }
#pragma unroll
        for (uint32_t k = 0; k < NARROW_PER_THREAD; ++k) rank[k] = t[k] < 4u ? atomicAdd(&s_cnt[t[k]], 1u) : 0u;
        __syncthreads();
        const uint32_t c0 = s_cnt[0], c1 = s_cnt[1], c2 = s_cnt[2], c3 = s_cnt[3];
#pragma unroll
        for (uint32_t k = 0; k < NARROW_PER_THREAD; ++k)
            if (t[k] < 4u) s_idx[(t[k] == 0 ? 0u : (t[k] == 1 ? c0 : (t[k] == 2 ? c0 + c1 : c0 + c1 + c2))) + rank[k]] = base + k * TPB + threadIdx.x;
        __syncthreads();
        const uint32_t total = c0 + c1 + c2 + c3;
        for (uint32_t x = threadIdx.x; x < total; x += TPB) {
            const int np = narrow_pair_thread(d, s_idx[x]);
            if (np >= 0) {
                my_m += 1;
                my_k += (uint32_t)np;
            }
        }
        __syncthreads();
    }
    __shared__ uint32_t s_m, s_k;  // one pair of global atomics per CTA (same-address atomics serialise in L2)
    if (threadIdx.x == 0) s_m = s_k = 0u;
    __syncthreads();
    my_m = cg::reduce(cg::tiled_partition<32>(cg::this_thread_block()), my_m, cg::plus<uint32_t>());
    my_k = cg::reduce(cg::tiled_partition<32>(cg::this_thread_block()), my_k, cg::plus<uint32_t>());
    if ((threadIdx.x & 31u) == 0 && (my_m | my_k)) {
        atomicAdd(&s_m, my_m);
        atomicAdd(&s_k, my_k);
    }
    __syncthreads();
    if (threadIdx.x == 0 && (s_m | s_k)) {
        atomicAdd(&d.counters->n_manifolds, s_m);
        atomicAdd(&d.counters->n_points, s_k);
    }
}

// ---- K8: graph colouring, one cooperative launch ---------------------------------------------------------------------------------
// Jones-Plassmann rounds over the pending manifolds with one grid barrier per round.  Round r reads the priorities
// posted for r in maxprio[r & 1] and posts those of the losers for r + 1 into the other array.  The result equals a
// sequential greedy colouring in descending priority, so it is a pure function of the contact graph and the body ids.
constexpr int COLOR_REG_SLOTS = 2;  // pending manifolds a thread keeps in registers across the rounds
constexpr uint32_t COLOR_COMPACT_ROUND = 3;   // from this round on the streamed manifolds are kept in compacted lists
constexpr int FLOW_SLOTS = 8;       // manifolds per thread the dataflow colouring can hold in registers
constexpr int FLOW_BIG_SLOTS = 24;  // ... and in a compacted per-thread list in local memory (worlds of ~10^6 bodies)

__global__ void __launch_bounds__(TPB, 4) k_color(Dev d, uint32_t flow_reg_slots) {
    cg::grid_group grid = cg::this_grid();
    const bool dead = overflowed(d);
    const uint32_t n = dead ? 0u : live_pairs(d);
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    // Same-address global atomics serialise in L2, so the per-colour populations are histogrammed in shared memory
    // for the whole kernel and the "anything left?" flag costs at most one global atomic per CTA and round.
    __shared__ uint32_t s_hist[MAX_COLORS];
    for (uint32_t c = threadIdx.x; c < MAX_COLORS; c += blockDim.x) s_hist[c] = 0u;
    __syncthreads();
    // ---- dataflow colouring: no rounds, no barriers (see flow_try) ----
    // Every thread holds its manifolds p = tid + k * nth in registers and probes them round-robin; all threads are
    // resident (cooperative launch) and every pending manifold keeps being probed, so the pending manifold with the
    // globally highest priority — which is always ready — always gets its turn: no deadlock.
    // (flow_abort is only written by the narrowphase, so the decision is the same for every thread of the grid)
    const bool flow_any = d.flow != 0u && d.counters->flow_abort == 0u;
    // (flow_reg_slots = FLOW_SLOTS; 0 sends every world through the list flavour: R2D_FLOW_LIST=1, tests)
    const bool flow_big = flow_any && n > flow_reg_slots * nth && n <= (uint32_t)FLOW_BIG_SLOTS * nth;
    const bool flow = flow_any && n <= (uint32_t)FLOW_BIG_SLOTS * nth;
    if (flow_big) {
        // The same probing with the thread's pending manifolds in a list (local memory) that shrinks as they are coloured:
        // a pass of a thread gets shorter and shorter, so the few long chains (the contacts of a hub body colour one after
        // the other) are probed at the rate of a short list, not of FLOW_BIG_SLOTS gathers.
        uint32_t fa[FLOW_BIG_SLOTS], fb[FLOW_BIG_SLOTS], fm[FLOW_BIG_SLOTS], fp[FLOW_BIG_SLOTS];
        uint32_t cnt = 0;
        for (int k = 0; k < FLOW_BIG_SLOTS; ++k) {
            const uint32_t p = tid + (uint32_t)k * nth;
            if (p < n && d.m_color[p] == COLOR_PENDING) {
                const uint4 h = d.m_hdr[p];
                fa[cnt] = h.x;
                fb[cnt] = h.y;
                fm[cnt] = (h.w & 3u) | (flow_ranks(d, h, d.m_prio[p]) << 2);
                fp[cnt] = p;
                cnt += 1u;
            }
        }
        uint32_t idle = 0;
        for (;;) {
            bool progress = false;
            uint32_t min_lag = 0xFFFFFFFFu;
            for (uint32_t k = 0; k < cnt;) {
                uint32_t c = 0, lag = 0;
                const int r = flow_try(d, fp[k], fa[k], fb[k], fm[k] & 3u, (fm[k] >> 2) & 0xFFFFu, &c, &lag);
                if (r == 1) {
                    atomicAdd(&s_hist[c], 1u);
                    owner_bit_set(d, fa[k], fb[k], fm[k] & 3u, c);
                    cnt -= 1u;
                    fa[k] = fa[cnt];
                    fb[k] = fb[cnt];
                    fm[k] = fm[cnt];
                    fp[k] = fp[cnt];
                    progress = true;
                } else if (r == 0) {
                    min_lag = lag < min_lag ? lag : min_lag;
                    ++k;
                } else {
                    cnt = 0u;  // colour overflow: flow_fail is set, everybody leaves
                }
            }
            if (!__any_sync(0xffffffffu, cnt != 0u)) break;
            if (__any_sync(0xffffffffu, progress)) {
                idle = 0;
                continue;
            }
            const uint32_t warp_lag = __reduce_min_sync(0xffffffffu, min_lag);
            if (warp_lag > 1u) backoff_ns((warp_lag < 8u ? warp_lag - 1u : 7u) * FLOW_SLEEP_UNIT);
            if ((++idle & 63u) == 0u) {
                if (*((volatile uint32_t*)&d.counters->flow_fail)) break;
                if (idle > (1u << 20)) {
                    atomicOr(&d.counters->err, ERR_FLOW_STALL);
                    atomicOr(&d.counters->flow_fail, 1u);
                    break;
                }
            }
        }
    } else if (flow) {
        uint32_t fa[FLOW_SLOTS], fb[FLOW_SLOTS], fm[FLOW_SLOTS];  // ref slot, inc slot, dyn mask | ranks << 2 | pending << 31
        uint32_t n_pending = 0;
#pragma unroll
        for (int k = 0; k < FLOW_SLOTS; ++k) {
            const uint32_t p = tid + (uint32_t)k * nth;
            fm[k] = 0u;
            fa[k] = fb[k] = 0u;
            if (p < n && d.m_color[p] == COLOR_PENDING) {
                const uint4 h = d.m_hdr[p];
                fa[k] = h.x;
                fb[k] = h.y;
                fm[k] = (h.w & 3u) | (flow_ranks(d, h, d.m_prio[p]) << 2) | 0x80000000u;
                n_pending += 1u;
            }
        }
        // The warp stays converged: every pass ends in a warp vote, so a lane that waits for a manifold held by a
        // sibling lane (neighbouring pairs share bodies) can never spin ahead of the lane it waits for.
        uint32_t idle = 0;
        for (;;) {
            bool progress = false;
            uint32_t min_lag = 0xFFFFFFFFu;
#pragma unroll
            for (int k = 0; k < FLOW_SLOTS; ++k) {
                if (!(fm[k] & 0x80000000u)) continue;
                uint32_t c = 0, lag = 0;
                const int r = flow_try(d, tid + (uint32_t)k * nth, fa[k], fb[k], fm[k] & 3u, (fm[k] >> 2) & 0xFFFFu, &c, &lag);
                if (r == 1) {
                    atomicAdd(&s_hist[c], 1u);
                    owner_bit_set(d, fa[k], fb[k], fm[k] & 3u, c);
                    fm[k] = 0u;
                    n_pending -= 1u;
                    progress = true;
                } else if (r == 0) {
                    min_lag = lag < min_lag ? lag : min_lag;
                } else {
                    n_pending = 0u;  // colour overflow: flow_fail is set, everybody leaves
                }
            }
            if (!__any_sync(0xffffffffu, n_pending != 0u)) break;
            if (__any_sync(0xffffffffu, progress)) {
                idle = 0;
                continue;
            }
            const uint32_t warp_lag = __reduce_min_sync(0xffffffffu, min_lag);
            if (warp_lag > 1u) backoff_ns((warp_lag < 8u ? warp_lag - 1u : 7u) * FLOW_SLEEP_UNIT);
            if ((++idle & 63u) == 0u) {
                if (*((volatile uint32_t*)&d.counters->flow_fail)) break;
                if (idle > (1u << 20)) {  // a stall would be a bug; report it instead of hanging the GPU
                    atomicOr(&d.counters->err, ERR_FLOW_STALL);
                    atomicOr(&d.counters->flow_fail, 1u);
                    break;
                }
            }
        }
    }
    if (flow) {
        __syncthreads();
        for (uint32_t c = threadIdx.x; c < MAX_COLORS; c += blockDim.x)
            if (s_hist[c]) atomicAdd(&d.color_count[c], s_hist[c]);
        grid.sync();
        if (__ldcg(&d.counters->flow_fail) == 0u) {
            if (tid == 0) {
                uint32_t nc = 0;
                for (uint32_t c = 0; c < FLOW_COLORS; ++c)
                    if (__ldcg(&d.color_count[c])) nc = c + 1u;
                d.counters->n_colors = nc;
                d.counters->n_own_scan = nc * (d.own_words + 1u);
                d.counters->n_rounds = 0u;
                d.counters->flow_used = 1u;
            }
            return;
        }
        // abandoned: forget the partial result and colour by rounds (maxprio / used are still zero)
        for (uint32_t p = tid; p < n; p += nth)
            if (d.m_color[p] < MAX_COLORS) d.m_color[p] = COLOR_PENDING;
        for (size_t x = tid; x < (size_t)d.own_words * FLOW_COLORS; x += nth) d.own_bits[x] = 0u;   // owner bits of the partial result
        if (blockIdx.x == 0)
            for (uint32_t c = threadIdx.x; c < MAX_COLORS; c += blockDim.x) d.color_count[c] = 0u;
        for (uint32_t c = threadIdx.x; c < MAX_COLORS; c += blockDim.x) s_hist[c] = 0u;
        __syncthreads();
        grid.sync();
    }
    // ---- Jones-Plassmann rounds ----
    // The first COLOR_REG_SLOTS slots of a thread (p = tid + k * nth) live in registers for all rounds: header and
    // priority are read once, a round costs only the gathers of the two body words and the atomics.
    uint4 rh[COLOR_REG_SLOTS];
    unsigned long long rp[COLOR_REG_SLOTS];
    bool pend[COLOR_REG_SLOTS];
#pragma unroll
    for (int k = 0; k < COLOR_REG_SLOTS; ++k) {
        const uint32_t p = tid + (uint32_t)k * nth;
        pend[k] = p < n && d.m_color[p] == COLOR_PENDING;
        if (pend[k]) {
            rh[k] = d.m_hdr[p];
            rp[k] = d.m_prio[p];
            color_post(d, rh[k].x, rh[k].y, (rh[k].w & 1u) != 0, (rh[k].w & 2u) != 0, rp[k], 1u);
        }
    }
    for (uint32_t p = tid + COLOR_REG_SLOTS * nth; p < n; p += nth) {
        if (d.m_color[p] != COLOR_PENDING) continue;
        const uint4 h = d.m_hdr[p];
        color_post(d, h.x, h.y, (h.w & 1u) != 0, (h.w & 2u) != 0, d.m_prio[p], 1u);
    }
    grid.sync();
    uint32_t round = 1;
    for (; round < MAX_COLOR_ROUNDS; ++round) {
        int left = 0;
#pragma unroll
        for (int k = 0; k < COLOR_REG_SLOTS; ++k) {
            if (!pend[k]) continue;
            uint32_t c;
            const int r = color_round_core(d, tid + (uint32_t)k * nth, rh[k], rp[k], round, &c);
            if (r == 1) {
                atomicAdd(&s_hist[c], 1u);
                owner_bit_set(d, rh[k].x, rh[k].y, rh[k].w, c);
            }
            if (r == 2)
                left = 1;
            else
                pend[k] = false;
        }
        // Manifolds beyond the register slots (worlds with more than ~300k pairs) are streamed.  For the first
        // COLOR_COMPACT_ROUND rounds every slot is visited; from then on only the survivors, kept in two lists that swap
        // every round (a million-body pile is coloured within ~10 rounds except for the contacts of its few hub bodies,
        // which take one round per contact: 47 rounds over 2.7 M slots cost 1.35 ms on mixed1M).
        {
            const uint32_t* list_in = d.pend_list + (size_t)(round & 1u) * d.cap_pairs;
            uint32_t* list_out = d.pend_list + (size_t)((round + 1u) & 1u) * d.cap_pairs;
            const bool compact = round >= COLOR_COMPACT_ROUND;
            const bool from_list = round > COLOR_COMPACT_ROUND;
            const uint32_t first = from_list ? tid : tid + COLOR_REG_SLOTS * nth;
            const uint32_t count = from_list ? __ldcg(&d.pend_cnt[round]) : n;
            // whole warps: the survivors of a warp are appended with one atomic
            for (uint32_t x0 = first - (tid & 31u); x0 < count; x0 += nth) {
                const uint32_t x = x0 + (tid & 31u);
                uint32_t p = 0;
                int r = 0;
                if (x < count) {
                    p = from_list ? __ldcg(&list_in[x]) : x;
                    r = color_round_thread(d, p, round);
                }
                if (r == 2) {
                    left = 1;
                } else if (r == 1) {
                    atomicAdd(&s_hist[d.m_color[p]], 1u);
                    owner_bit_thread(d, p);
                }
                if (compact) {
                    const uint32_t votes = __ballot_sync(0xffffffffu, r == 2);
                    if (votes) {
                        uint32_t base = 0;
                        if ((tid & 31u) == 0u) base = atomicAdd(&d.pend_cnt[round + 1u], (uint32_t)__popc(votes));
                        base = __shfl_sync(0xffffffffu, base, 0);
                        if (r == 2) list_out[base + (uint32_t)__popc(votes & ((1u << (tid & 31u)) - 1u))] = p;
                    }
                }
            }
        }
        if (__syncthreads_or(left) && threadIdx.x == 0) atomicAdd(&d.round_left[round], 1u);
        grid.sync();
        if (__ldcg(&d.round_left[round]) == 0) break;
    }
    for (uint32_t c = threadIdx.x; c < MAX_COLORS; c += blockDim.x)
        if (s_hist[c]) atomicAdd(&d.color_count[c], s_hist[c]);
    grid.sync();
    if (tid == 0) {
        d.counters->n_rounds = round;
        if (round >= MAX_COLOR_ROUNDS) atomicOr(&d.counters->err, ERR_ROUNDS);
        uint32_t nc = 0;
        for (uint32_t c = 0; c < MAX_COLORS; ++c)
            if (__ldcg(&d.color_count[c])) nc = c + 1u;
        d.counters->n_colors = nc;
        d.counters->n_own_scan = nc * (d.own_words + 1u);
    }
}

// ---- K8 for batches of small worlds: one CTA colours one world with the per-body words in shared memory -------------------------
// World w's candidate pairs are the contiguous slots [ent_off[first bucket of w], ent_off[first bucket of w + 1]) (the
// pair list is bucket-major and a world owns a contiguous bucket range), and its conflict graph is closed, so the
// Jones-Plassmann rounds need only __syncthreads().  Same rounds, same priorities, same colours as k_color.
constexpr uint32_t COLOR_WORLD_MAX_BODIES = 512;

__global__ void __launch_bounds__(WORLD_TPB) k_color_worlds(Dev d) {
    __shared__ unsigned long long s_mp0[COLOR_WORLD_MAX_BODIES], s_mp1[COLOR_WORLD_MAX_BODIES];
    __shared__ unsigned long long s_used[COLOR_WORLD_MAX_BODIES * COLOR_WORDS];
    __shared__ uint32_t s_hist[MAX_COLORS];
    if (overflowed(d)) return;
    for (uint32_t c = threadIdx.x; c < MAX_COLORS; c += blockDim.x) s_hist[c] = 0u;
    uint32_t max_round = 0;
    for (uint32_t w = blockIdx.x; w < d.n_worlds; w += gridDim.x) {
        const uint32_t b0 = d.world_base[w], b1 = d.world_base[w + 1], nb = b1 - b0;
        // fine grid: the pair list is body-major instead (pair_cnt[b + 1] = first pair emitted by body b)
        const uint32_t p0 = d.fine_on ? d.pair_cnt[b0 + 1] : d.ent_off[d.table_mult * b0];
        const uint32_t p1 = d.fine_on ? d.pair_cnt[b1 + 1] : d.ent_off[d.table_mult * b1];
        for (uint32_t i = threadIdx.x; i < nb; i += blockDim.x) {
            s_mp0[i] = 0ull;
            s_mp1[i] = 0ull;
#pragma unroll
            for (uint32_t q = 0; q < COLOR_WORDS; ++q) s_used[i * COLOR_WORDS + q] = 0ull;
        }
        __syncthreads();
        Dev ds = d;
        ds.color_smem = 1u;
        ds.maxprio0 = s_mp0 - b0;
        ds.maxprio1 = s_mp1 - b0;
        ds.used = s_used - (size_t)b0 * COLOR_WORDS;
        // the first WORLD_REG_SLOTS slots of a thread stay in registers for all rounds (header + priority read once)
        constexpr int WORLD_REG_SLOTS = 6;
        uint4 rh[WORLD_REG_SLOTS];
        unsigned long long rp[WORLD_REG_SLOTS];
        bool pend[WORLD_REG_SLOTS];
#pragma unroll
        for (int k = 0; k < WORLD_REG_SLOTS; ++k) {
            const uint32_t p = p0 + threadIdx.x + (uint32_t)k * blockDim.x;
            pend[k] = p < p1 && d.m_color[p] == COLOR_PENDING;
            if (pend[k]) {
                rh[k] = d.m_hdr[p];
                rp[k] = d.m_prio[p];
                color_post(ds, rh[k].x, rh[k].y, (rh[k].w & 1u) != 0, (rh[k].w & 2u) != 0, rp[k], 1u);
            }
        }
        for (uint32_t p = p0 + threadIdx.x + WORLD_REG_SLOTS * blockDim.x; p < p1; p += blockDim.x) {
            if (d.m_color[p] != COLOR_PENDING) continue;
            const uint4 h = d.m_hdr[p];
            color_post(ds, h.x, h.y, (h.w & 1u) != 0, (h.w & 2u) != 0, d.m_prio[p], 1u);
        }
        __syncthreads();
        uint32_t round = 1;
        for (; round < MAX_COLOR_ROUNDS; ++round) {
            int left = 0;
#pragma unroll
            for (int k = 0; k < WORLD_REG_SLOTS; ++k) {
                if (!pend[k]) continue;
                uint32_t c;
                const int r = color_round_core(ds, p0 + threadIdx.x + (uint32_t)k * blockDim.x, rh[k], rp[k], round, &c);
                if (r == 1) {
                    atomicAdd(&s_hist[c], 1u);
                    owner_bit_set(d, rh[k].x, rh[k].y, rh[k].w, c);
                }
                if (r == 2)
                    left = 1;
                else
                    pend[k] = false;
            }
            for (uint32_t p = p0 + threadIdx.x + WORLD_REG_SLOTS * blockDim.x; p < p1; p += blockDim.x) {
                const int r = color_round_thread(ds, p, round);
                if (r == 2) {
                    left = 1;
                } else if (r == 1) {
                    atomicAdd(&s_hist[d.m_color[p]], 1u);
                    owner_bit_thread(d, p);
                }
            }
            if (!__syncthreads_or(left)) break;
        }
        if (round >= MAX_COLOR_ROUNDS && threadIdx.x == 0) atomicOr(&d.counters->err, ERR_ROUNDS);
        max_round = round > max_round ? round : max_round;
        __syncthreads();
    }
    for (uint32_t c = threadIdx.x; c < MAX_COLORS; c += blockDim.x)
        if (s_hist[c]) {
            atomicAdd(&d.color_count[c], s_hist[c]);
            if (d.world_fused) atomicMax(&d.counters->n_colors, c + 1u);   // (else k_scan_owners derives it)
        }
    if (threadIdx.x == 0) atomicMax(&d.counters->n_rounds, max_round);
}
// ---- K8 for batches of small worlds, sorted form ------------------------------------------------------------------------------
// "Greedy in descending priority" taken literally: the world's pending manifolds are sorted by priority (64-bit keys
// {priority, pair index}: 16 buckets by the top bits, a rank sort by one warp inside each) and one warp colours them
// in that order, 32 at a time — lowest colour free on both bodies, per-body colour masks in shared memory — while the
// Jones-Plassmann rounds above need ~26 block-wide rounds with 8 of 32 lanes active.  Same colours by construction (JP with unique priorities == sequential greedy; ties are only
// possible between manifolds that share no body and then the order does not matter).  Worlds with more than
// SEQ_WORLD_PAIRS candidate pairs take the rounds with the global per-body arrays instead.
constexpr uint32_t SEQ_WORLD_PAIRS = 1024;

__global__ void __launch_bounds__(WORLD_TPB) k_color_worlds_seq(Dev d, uint32_t smem_bodies, uint32_t export_used) {
    extern __shared__ unsigned long long s_used_dyn[];   // smem_bodies x COLOR_WORDS
    __shared__ unsigned long long s_key[SEQ_WORLD_PAIRS];
    __shared__ uint32_t s_pair[SEQ_WORLD_PAIRS];
    __shared__ unsigned short s_col[SEQ_WORLD_PAIRS];   // colour, or 0xFFFF: no colour free (COLOR_DROPPED)
    __shared__ uint32_t s_hist[MAX_COLORS];
    __shared__ uint32_t s_lanes[COLOR_WORLD_MAX_BODIES];   // per body: the lanes of the current chunk that touch it
    __shared__ unsigned short s_order[SEQ_WORLD_PAIRS];    // pair indices in descending priority
    __shared__ uint32_t s_bcnt[16], s_bfill[16], s_bstart[16];
    __shared__ uint32_t s_n;
    if (overflowed(d)) return;
    for (uint32_t c = threadIdx.x; c < MAX_COLORS; c += blockDim.x) s_hist[c] = 0u;
    for (uint32_t c = threadIdx.x; c < COLOR_WORLD_MAX_BODIES; c += blockDim.x) s_lanes[c] = 0u;
    uint32_t max_round = 0;
    for (uint32_t w = blockIdx.x; w < d.n_worlds; w += gridDim.x) {
        const uint32_t b0 = d.world_base[w], b1 = d.world_base[w + 1], nb = b1 - b0;
        const uint32_t p0 = d.fine_on ? d.pair_cnt[b0 + 1] : d.ent_off[d.table_mult * b0];
        const uint32_t p1 = d.fine_on ? d.pair_cnt[b1 + 1] : d.ent_off[d.table_mult * b1];
        const uint32_t np = p1 - p0;
        __syncthreads();
        if (np <= SEQ_WORLD_PAIRS && nb <= smem_bodies && nb <= COLOR_WORLD_MAX_BODIES) {
            if (threadIdx.x < 16u) s_bcnt[threadIdx.x] = s_bfill[threadIdx.x] = 0u;
            for (uint32_t i = threadIdx.x; i < nb * COLOR_WORDS; i += blockDim.x) s_used_dyn[i] = 0ull;
            __syncthreads();
            // ---- sort by priority, descending: 16 buckets by the top bits (a hash: uniform), then a rank sort inside each ----
            constexpr uint32_t PER_THREAD = SEQ_WORLD_PAIRS / WORLD_TPB;
            unsigned long long key[PER_THREAD];   // {priority (52 bits), pair index in the world (12 bits)}; 0 = not pending
#pragma unroll
            for (uint32_t k = 0; k < PER_THREAD; ++k) {
                const uint32_t i = threadIdx.x + k * WORLD_TPB, p = p0 + i;
                key[k] = 0ull;
                if (i < np && d.m_color[p] == COLOR_PENDING) {
                    const uint4 h = d.m_hdr[p];
                    s_pair[i] = (h.x - b0) | ((h.y - b0) << 12) | ((h.w & 3u) << 24);
                    key[k] = (d.m_prio[p] << 12) | (unsigned long long)i;
                    atomicAdd(&s_bcnt[(uint32_t)(key[k] >> 60)], 1u);
                }
            }
            __syncthreads();
            if (threadIdx.x == 0) {   // descending: bucket 15 first
                uint32_t run = 0;
                for (int b = 15; b >= 0; --b) {
                    s_bstart[b] = run;
                    run += s_bcnt[b];
                }
                s_n = run;
            }
            __syncthreads();
#pragma unroll
            for (uint32_t k = 0; k < PER_THREAD; ++k) {
                if (key[k] != 0ull) {   // (a priority is never 0: its low word is the sum of two distinct ids)
                    const uint32_t b = (uint32_t)(key[k] >> 60);
                    s_key[s_bstart[b] + atomicAdd(&s_bfill[b], 1u)] = key[k];
                }
            }
            __syncthreads();
            const uint32_t n = s_n;
            for (uint32_t b = threadIdx.x >> 5; b < 16u; b += blockDim.x >> 5) {   // a warp per bucket
                const uint32_t st = s_bstart[b], cnt = s_bcnt[b];
                for (uint32_t e = threadIdx.x & 31u; e < cnt; e += 32u) {
                    const unsigned long long mine = s_key[st + e];
                    uint32_t rank = 0;
                    for (uint32_t j = 0; j < cnt; ++j) rank += s_key[st + j] > mine ? 1u : 0u;   // keys are unique (pair index)
                    s_order[st + rank] = (unsigned short)((uint32_t)mine & 0xFFFu);
                }
            }
            __syncthreads();
            if (threadIdx.x < 32u) {
                // One warp walks the sorted list 32 manifolds at a time.  Inside a chunk a lane depends on the EARLIER lanes
                // that share one of its non-static bodies (found once per chunk, see s_lanes); lanes whose dependencies are
                // done colour in parallel — they touch disjoint bodies — so a chunk takes 2-3 passes instead of 32 steps.
                const uint32_t lane = threadIdx.x;
                for (uint32_t base = 0; base < n; base += 32u) {
                    const uint32_t k = base + lane;
                    const bool active = k < n;
                    const uint32_t i = active ? (uint32_t)s_order[k] : 0u, pr = active ? s_pair[i] : 0u;
                    const uint32_t l1 = pr & 0xFFFu, l2 = (pr >> 12) & 0xFFFu;
                    const bool dyn1 = ((pr >> 24) & 1u) != 0, dyn2 = ((pr >> 25) & 1u) != 0;
                    // lanes of this chunk on each body, collected in shared memory (two atomicOr per lane instead of 62 shuffles)
                    if (dyn1) atomicOr(&s_lanes[l1], 1u << lane);
                    if (dyn2) atomicOr(&s_lanes[l2], 1u << lane);
                    __syncwarp();
                    const uint32_t dep = ((dyn1 ? s_lanes[l1] : 0u) | (dyn2 ? s_lanes[l2] : 0u)) & ((1u << lane) - 1u);
                    __syncwarp();
                    if (dyn1) s_lanes[l1] = 0u;
                    if (dyn2) s_lanes[l2] = 0u;
                    bool pending = active;
                    for (;;) {
                        const uint32_t pend_mask = __ballot_sync(0xffffffffu, pending);
                        if (!pend_mask) break;
                        if (pending && !(dep & pend_mask)) {
                            uint32_t color = MAX_COLORS;
                            for (uint32_t q = 0; q < COLOR_WORDS; ++q) {
                                unsigned long long u = 0ull;
                                if (dyn1) u |= s_used_dyn[l1 * COLOR_WORDS + q];
                                if (dyn2) u |= s_used_dyn[l2 * COLOR_WORDS + q];
                                if (~u) {
                                    color = q * 64u + (uint32_t)__ffsll((long long)~u) - 1u;
                                    break;
                                }
                            }
                            if (color >= MAX_COLORS) {   // a body with more than MAX_COLORS contacts: the manifold sits this call out
                                atomicAdd(&d.counters->n_dropped, 1u);
                                s_col[i] = 0xFFFFu;
                            } else {
                                const unsigned long long bit = 1ull << (color & 63u);
                                if (dyn1) s_used_dyn[l1 * COLOR_WORDS + (color >> 6)] |= bit;
                                if (dyn2) s_used_dyn[l2 * COLOR_WORDS + (color >> 6)] |= bit;
                                s_col[i] = (unsigned short)color;
                                atomicAdd(&s_hist[color], 1u);
                            }
                            pending = false;
                        }
                        __syncwarp();
                    }
                }
            }
            __syncthreads();
            for (uint32_t k = threadIdx.x; k < n; k += blockDim.x) {
                const uint32_t i = (uint32_t)s_order[k], pr = s_pair[i];
                if (s_col[i] == 0xFFFFu) {
                    d.m_color[p0 + i] = COLOR_DROPPED;
                    continue;
                }
                d.m_color[p0 + i] = s_col[i];
                owner_bit_set(d, b0 + (pr & 0xFFFu), b0 + ((pr >> 12) & 0xFFFu), (pr >> 24) & 3u, s_col[i]);
            }
            if (export_used)   // the dataflow sweep derives rank / degree from the per-body masks (body_color_rank)
                for (uint32_t i = threadIdx.x; i < nb * COLOR_WORDS; i += blockDim.x) d.used[(size_t)b0 * COLOR_WORDS + i] = s_used_dyn[i];
        } else {
            // rounds on the global per-body arrays (zeroed at the start of the step); only this CTA touches this world
            for (uint32_t p = p0 + threadIdx.x; p < p1; p += blockDim.x) {
                if (d.m_color[p] != COLOR_PENDING) continue;
                const uint4 h = d.m_hdr[p];
                color_post(d, h.x, h.y, (h.w & 1u) != 0, (h.w & 2u) != 0, d.m_prio[p], 1u);
            }
            __syncthreads();
            uint32_t round = 1;
            for (; round < MAX_COLOR_ROUNDS; ++round) {
                int left = 0;
                for (uint32_t p = p0 + threadIdx.x; p < p1; p += blockDim.x) {
                    const int r = color_round_thread(d, p, round);
                    if (r == 2) {
                        left = 1;
                    } else if (r == 1) {
                        atomicAdd(&s_hist[d.m_color[p]], 1u);
                        owner_bit_thread(d, p);
                    }
                }
                if (!__syncthreads_or(left)) break;
            }
            if (round >= MAX_COLOR_ROUNDS && threadIdx.x == 0) atomicOr(&d.counters->err, ERR_ROUNDS);
            max_round = round > max_round ? round : max_round;
        }
    }
    __syncthreads();
    for (uint32_t c = threadIdx.x; c < MAX_COLORS; c += blockDim.x)
        if (s_hist[c]) {
            atomicAdd(&d.color_count[c], s_hist[c]);
            if (d.world_fused) atomicMax(&d.counters->n_colors, c + 1u);   // (else k_scan_owners derives it)
        }
    if (threadIdx.x == 0 && max_round) atomicMax(&d.counters->n_rounds, max_round);
}
// ---- K9: group the manifolds by colour and evaluate the pre-step (collision.zig:102-133) ---------------------------------------
__global__ void __launch_bounds__(TPB) k_partition_prestep(Dev d) {
    if (overflowed(d)) return;
    const uint32_t n = live_pairs(d);
    const uint32_t nc = d.counters->n_colors, stride = d.own_words + 1u;
    if (blockIdx.x == 0) {
        // colour segment starts (each begins on a warp boundary) and the padding slots at the end of every segment
        for (uint32_t c = threadIdx.x; c <= nc; c += blockDim.x) d.color_start[c] = d.own_pos[(size_t)c * stride];
        for (uint32_t k = threadIdx.x; k < nc * COLOR_ALIGN; k += blockDim.x) {
            const uint32_t c = k / COLOR_ALIGN;
            const uint32_t at = d.own_pos[(size_t)c * stride + d.own_words] + (k % COLOR_ALIGN);
            if (at < d.own_pos[(size_t)(c + 1u) * stride]) d.s_hdr[at] = make_uint4(0u, 0u, S_EMPTY, 0u);
        }
    }
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
        if (d.m_color[p] >= MAX_COLORS) continue;
        gather_prestep_thread(d, p, manifold_slot(d, p));
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
// Shared-memory cache of the solver records: thread t owns manifolds t, t + nth, ... for the whole kernel, so the first
// SOLVE_SMEM_SLOTS of them are copied once into shared memory (record k of thread x at index k * PSOLVE_TPB + x of each array)
// and every sweep reads its constants — and keeps its accumulated impulses — there; only the two body words of a
// manifold go through L2.  Records beyond the cache (worlds with more than ~245k manifolds) stream from global memory.
// threads per CTA of the persistent solver (one CTA per SM) x cached records per thread: 256 x 6 while most records fit the
// cache (512 x 3 measured the same on pile100k); 512 x 3 for worlds whose records mostly stream from HBM — twice the
// threads walk half as long a sequence of exposed record fetches each (mixed1M, 2.2 M manifolds)
constexpr int PSOLVE_TPB = 256, PSOLVE_TPB_BIG = 512;
constexpr int SOLVE_SMEM_SLOTS = 6, SOLVE_SMEM_SLOTS_BIG = 3;
constexpr int SOLVE_SMEM_BYTES_PER_RECORD = 6 * 16 + 8 + 2 * 16 + 8;  // hdr nf inv dep r0 pm0 | acc0 | r1 pm1 | acc1 = 144
constexpr size_t SOLVE_SMEM_BYTES = (size_t)SOLVE_SMEM_SLOTS * PSOLVE_TPB * SOLVE_SMEM_BYTES_PER_RECORD;

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_record_l2(const Dev& d, uint32_t m) {
    prefetch_l2(&d.s_hdr[m]);
    prefetch_l2(&d.s_nf[m]);
    prefetch_l2(&d.s_inv[m]);
    prefetch_l2(&d.s_dep[m]);
    prefetch_l2(&d.s_r0[m]);
    prefetch_l2(&d.s_pm0[m]);
    prefetch_l2(&d.s_acc0[m]);
    prefetch_l2(&d.s_r1[m]);     // (whether the record has a second point is in its header: not known yet)
    prefetch_l2(&d.s_pm1[m]);
    prefetch_l2(&d.s_acc1[m]);
}
template <int PT, int SLOTS>
__global__ void __launch_bounds__(PT) k_solve_persistent(Dev d, float sub_dt, uint32_t S, uint32_t I,
                                                          const uint32_t* __restrict__ joint_color_start, uint32_t n_joint_colors,
                                                          uint32_t smem_slots, uint32_t prefetch) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cg::grid_group grid = cg::this_grid();
    if (overflowed(d) || d.counters->err != 0u) return;  // uniform across the grid
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    const uint32_t n_manifolds = d.color_start[d.counters->n_colors];
    const bool jflow = d.joints_flow != 0u && n_joint_colors != 0u;
    // ---- stage my records ----
    Dev ds = d;  // same code path, record arrays redirected to shared memory
    {
        const size_t n = (size_t)SLOTS * PT;
        unsigned char* q = smem_raw;
        ds.s_hdr = (uint4*)q;    q += n * 16;
        ds.s_nf = (float4*)q;    q += n * 16;
        ds.s_inv = (float4*)q;   q += n * 16;
        ds.s_dep = (uint4*)q;    q += n * 16;
        ds.s_r0 = (float4*)q;    q += n * 16;
        ds.s_pm0 = (float4*)q;   q += n * 16;
        ds.s_r1 = (float4*)q;    q += n * 16;
        ds.s_pm1 = (float4*)q;   q += n * 16;
        ds.s_acc0 = (float2*)q;  q += n * 8;
        ds.s_acc1 = (float2*)q;
    }
    auto stage = [&]() {
        for (uint32_t k = 0; k < smem_slots; ++k) {
            const uint32_t m = tid + k * nth, l = k * PT + threadIdx.x;
            if (m >= n_manifolds) break;
            const uint4 h = d.s_hdr[m];
            ds.s_hdr[l] = h;
            if (h.z & S_EMPTY) continue;
            ds.s_nf[l] = d.s_nf[m];
            ds.s_inv[l] = d.s_inv[m];
            ds.s_dep[l] = d.s_dep[m];
            ds.s_r0[l] = d.s_r0[m];
            ds.s_pm0[l] = d.s_pm0[m];
            ds.s_acc0[l] = d.s_acc0[m];
            if ((h.z & 0xFFu) > 1u) {
                ds.s_r1[l] = d.s_r1[m];
                ds.s_pm1[l] = d.s_pm1[m];
                ds.s_acc1[l] = d.s_acc1[m];
            }
        }
    };
    // warm starting: the first sweep of the call reads its records (and their warm terms) from global memory; the cache is
    // filled after it, with the accumulated impulses that sweep left
    const bool warm = d.warm_on != 0u;
    if (!warm) stage();
    for (uint32_t s = 0; s < S; ++s) {
        if (s == 0)
            for (uint32_t i = tid; i < d.n_bodies; i += 4u * nth) integrate_batch<4>(d, i, nth, sub_dt, false, true, S == 1);
        grid.sync();
        for (uint32_t it = 0; it < I; ++it) {
            if (jflow) {   // joints through the version words, like the contacts: no barrier between colours or iterations
                for (uint32_t jb = tid & ~31u; jb < d.n_joints; jb += nth) {
                    const uint32_t j = jb + (tid & 31u);
                    solve_joint_flow(d, j, j < d.n_joints, sub_dt, it);
                }
            } else {
                for (uint32_t jc = 0; jc < n_joint_colors; ++jc) {
                    const uint32_t b = joint_color_start[jc], e = joint_color_start[jc + 1];
                    for (uint32_t j = b + tid; j < e; j += nth) solve_joint_thread(d, j, sub_dt);
                    grid.sync();
                }
            }
            const bool first = warm && s == 0 && it == 0;
            uint32_t k = 0;
            for (uint32_t m = tid; m < n_manifolds; m += nth, ++k) {
                // records beyond the cache stream from HBM: while this one waits for its bodies and runs, the lines of the
                // thread's next `prefetch` records are already on their way into L2 (a thread walks a fixed sequence)
                if (prefetch && k + prefetch >= smem_slots) {
                    const size_t mp = (size_t)m + (size_t)prefetch * nth;
                    if (mp < n_manifolds) prefetch_record_l2(d, (uint32_t)mp);
                }
                if (k < smem_slots && !first)
                    solve_contact_thread<true>(ds, k * PT + threadIdx.x, sub_dt, it);  // cached record
                else
                    solve_contact_thread<true>(d, m, sub_dt, it, first);
            }
            if (first) stage();
            if (n_joint_colors && !jflow) grid.sync();  // joints of the next iteration read what the contacts wrote
        }
        if (!n_joint_colors || jflow) grid.sync();
        // end of substep s fused with the start of substep s + 1: both are per-body, same thread, no barrier needed
        for (uint32_t i = tid; i < d.n_bodies; i += 4u * nth) integrate_batch<4>(d, i, nth, sub_dt, true, s + 1 < S, s + 2 == S);
    }
    if (warm && S * I > 0u)   // k_warm_save reads what every contact accumulated: flush the cached ones
        for (uint32_t k = 0; k < smem_slots; ++k) {
            const uint32_t m = tid + k * nth, l = k * PT + threadIdx.x;
            if (m >= n_manifolds) break;
            const uint4 h = ds.s_hdr[l];
            if (h.z & S_EMPTY) continue;
            d.s_acc0[m] = ds.s_acc0[l];
            if ((h.z & 0xFFu) > 1u) d.s_acc1[m] = ds.s_acc1[l];
        }
}

// ---- roadmap options (README.md:59-64; off by default) ----------------------------------------------------------------------
// warm starting: after the substep loop every contact leaves its accumulated impulses (per substep) in a hash table keyed
// by the stable ids of its two bodies; the next call's pre-step looks them up (gather_prestep_thread)
__global__ void __launch_bounds__(TPB) k_warm_save(Dev d, float inv_scale_div) {
    if (overflowed(d) || d.counters->err != 0u) return;
    const uint32_t n = d.color_start[d.counters->n_colors];
    for (uint32_t m = blockIdx.x * blockDim.x + threadIdx.x; m < n; m += gridDim.x * blockDim.x) {
        const uint4 h = d.s_hdr[m];
        if (h.z & S_EMPTY) continue;
        const uint32_t np = h.z & 0xFFu;
        const float2 a0 = np > 0u ? d.s_acc0[m] : make_float2(0.0f, 0.0f), a1 = np > 1u ? d.s_acc1[m] : make_float2(0.0f, 0.0f);
        const float4 v = make_float4(fdiv(a0.x, inv_scale_div), fdiv(a0.y, inv_scale_div), fdiv(a1.x, inv_scale_div), fdiv(a1.y, inv_scale_div));
        const unsigned long long key = warm_make_key(body_flags(d, h.x) >> FLAG_WORLD_SHIFT, body_id(d, h.x), body_id(d, h.y));
        const uint32_t meta = d.m_hdr[h.w].z;   // n_points | normal_id << 8
        uint32_t slot = warm_hash(key) & d.warm_mask;
        for (uint32_t probe = 0; probe < WARM_PROBES; ++probe, slot = (slot + 1u) & d.warm_mask) {
            const unsigned long long old = atomicCAS(&d.warm_key[slot], WARM_EMPTY, key);
            if (old == WARM_EMPTY || old == key) {
                d.warm_val[slot] = v;
                d.warm_meta[slot] = meta;
                break;
            }
        }
    }
}
// sleeping, start of a call: a body that has been slow for `calls` calls in a row and is not being pushed IS A STATIC BODY for
// the duration of this call — its static flag is set here and cleared by k_sleep_bodies
__global__ void __launch_bounds__(TPB) k_sleep_begin(Dev d, uint32_t calls) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < d.n_bodies; i += gridDim.x * blockDim.x) {
        const float4 sh = d.shape[i];
        const uint32_t flags = f2u(sh.z);
        uint32_t state = 0u;
        if (!(flags & FLAG_STATIC)) {
            const float4 f = d.frc[i];
            uint32_t cnt = d.sleep_cnt[i];
            if (f.x != 0.0f || f.y != 0.0f || f.z != 0.0f) d.sleep_cnt[i] = cnt = 0u;   // user input wakes
            // a body named by a joint never sleeps (a distance joint between two static bodies divides by w1 + w2 = 0)
            if (d.n_joints && d.body_nj[i] != 0u) d.sleep_cnt[i] = cnt = 0u;
            if (cnt >= calls) {
                state = 1u;
                d.shape[i] = make_float4(sh.x, sh.y, u2f(flags | FLAG_STATIC), sh.w);
            }
        }
        d.sleep_state[i] = state;
    }
}
// end of a call: sleepers get their flag back; everybody else is measured against the speed thresholds
__global__ void __launch_bounds__(TPB) k_sleep_bodies(Dev d) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < d.n_bodies; i += gridDim.x * blockDim.x) {
        const float4 sh = d.shape[i];
        const uint32_t flags = f2u(sh.z);
        if (d.sleep_state[i] & 1u) {
            d.shape[i] = make_float4(sh.x, sh.y, u2f(flags & ~FLAG_STATIC), sh.w);
            continue;
        }
        if (flags & FLAG_STATIC) continue;
        const float4 m = d.mom[i], pr = d.prop[i];
        const float vx = fdiv(m.x, pr.x), vy = fdiv(m.y, pr.x), w = fdiv(m.z, pr.y);
        const bool slow = fadd(fmul(vx, vx), fmul(vy, vy)) < SLEEP_LIN2 && fmul(w, w) < SLEEP_ANG2;
        d.sleep_cnt[i] = slow ? d.sleep_cnt[i] + 1u : 0u;
        d.sleep_state[i] = slow ? 0u : 2u;
    }
}
// one hop: a body that moved fast in this call wakes what it touches (the woken bodies do not wake others yet)
__global__ void __launch_bounds__(TPB) k_sleep_wake(Dev d) {
    if (overflowed(d)) return;
    const uint32_t n = live_pairs(d);
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
        if (d.m_color[p] == COLOR_NONE) continue;
        const uint4 h = d.m_hdr[p];
        // (true statics have no counter; a sleeper's flag has been cleared by k_sleep_bodies already)
        if ((d.sleep_state[h.x] & 2u) && !(body_flags(d, h.y) & FLAG_STATIC)) d.sleep_cnt[h.y] = 0u;
        if ((d.sleep_state[h.y] & 2u) && !(body_flags(d, h.x) & FLAG_STATIC)) d.sleep_cnt[h.x] = 0u;
    }
}

__device__ __forceinline__ uint32_t owner_rank(const Dev& d, uint32_t c, uint32_t slot) {  // owners of colour c below `slot`
    const uint32_t word = d.own_bits[(size_t)c * d.own_words + (slot >> 5)];
    return d.own_pos[(size_t)c * (d.own_words + 1u) + (slot >> 5)] + (uint32_t)__popc(word & ((1u << (slot & 31u)) - 1u));
}

// ---- K14: one world, one spatial TILE of bodies per SM — the momentum words of a tile live in shared memory ---------------------
// The dataflow sweep of k_solve_persistent is bound by the rate at which L2 serves scattered 16-byte body words (two
// loads and two stores per manifold and sweep).  Device slots are in Morton order, so the bodies [t B, (t + 1) B) are a
// compact patch of the world and most manifolds touch two bodies of the same patch.  CTA t therefore runs ALL
// manifolds owned by its bodies (per colour a contiguous range of the owner-ordered records, see manifold_owner) and
// keeps the momentum word of every body that no other CTA touches in shared memory, with its version counter next
// to it.  Bodies that a foreign CTA touches ("shared": marked by the partition kernel) stay in global memory with the
// 16-byte version protocol of k_solve_persistent.  Same per-body update sequence as everywhere else: bit-identical.
//
// A warp works on tasks of 32 consecutive records of ONE colour (no dependencies inside a warp) and takes its tasks in
// ascending (iteration, colour) order; all CTAs are resident (cooperative launch), so the lowest unfinished record of
// the grid can always run: no deadlock.  The first `cache_tasks` tasks of a CTA keep their records (and accumulated
// impulses) in shared memory for the whole kernel.
constexpr int TILE_TPB = 512;
constexpr uint32_t TILE_MAX_BODIES = 1024;
constexpr uint32_t TILE_MAX_TASKS = 512;
constexpr uint32_t TILE_LOC1 = 0x1000u, TILE_LOC2 = 0x2000u;  // s_hdr.z of a cached record: body word in shared memory
constexpr size_t TILE_FIXED_SMEM = (size_t)TILE_MAX_BODIES * 20 + (size_t)TILE_MAX_TASKS * 8 + (size_t)MAX_COLORS * 8;
constexpr size_t TILE_SMEM_BYTES = 232448 - 1024;  // all of it (the static part of the kernel is tiny)
constexpr uint32_t TILE_CACHE_TASKS = (uint32_t)((TILE_SMEM_BYTES - TILE_FIXED_SMEM) / (32 * 144));

// Hand-off of a tile-local body between two warps of the CTA: the version word is read with acquire and written with
// release semantics at block scope (ld.acquire.cta / st.release.cta), the momentum word it announces with ordinary
// (volatile: never cached in registers) accesses that those two order — the release/acquire pattern of the PTX memory model.
__device__ __forceinline__ uint32_t ld_acquire_shared_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_shared_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}
__device__ __forceinline__ float4 ld_volatile_shared_f4(const float4* p) {
    float4 v;
    asm volatile("ld.volatile.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "r"((uint32_t)__cvta_generic_to_shared(p))
                 : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_shared_f4(float4* p, float4 v) {
    asm volatile("st.volatile.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "f"(v.x), "f"(v.y),
                 "f"(v.z), "f"(v.w)
                 : "memory");
}

__global__ void __launch_bounds__(TILE_TPB) k_solve_tiles(Dev d, float sub_dt, uint32_t S, uint32_t I, uint32_t cache_tasks,
                                                          uint32_t max_tasks) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cg::grid_group grid = cg::this_grid();
    if (overflowed(d) || d.counters->err != 0u) return;  // uniform across the grid
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    const uint32_t B = d.tile_bodies;
    const uint32_t b0 = blockIdx.x * B < d.n_bodies ? blockIdx.x * B : d.n_bodies;
    const uint32_t b1 = b0 + B < d.n_bodies ? b0 + B : d.n_bodies;
    const uint32_t nb = b1 - b0, nc = d.counters->n_colors;
    // ---- shared memory layout ----
    unsigned char* q = smem_raw;
    float4* t_mom = (float4*)q;                 q += (size_t)TILE_MAX_BODIES * 16;
    uint32_t* t_ver = (uint32_t*)q;             q += (size_t)TILE_MAX_BODIES * 4;
    uint32_t* t_first = (uint32_t*)q;           q += (size_t)TILE_MAX_TASKS * 4;
    uint32_t* t_count = (uint32_t*)q;           q += (size_t)TILE_MAX_TASKS * 4;
    uint32_t* t_cbeg = (uint32_t*)q;            q += (size_t)MAX_COLORS * 4;
    uint32_t* t_cend = (uint32_t*)q;            q += (size_t)MAX_COLORS * 4;
    Dev ds = d;  // record arrays of the cached tasks
    {
        const size_t n = (size_t)cache_tasks * 32;
        ds.s_hdr = (uint4*)q;    q += n * 16;
        ds.s_nf = (float4*)q;    q += n * 16;
        ds.s_inv = (float4*)q;   q += n * 16;
        ds.s_dep = (uint4*)q;    q += n * 16;
        ds.s_r0 = (float4*)q;    q += n * 16;
        ds.s_pm0 = (float4*)q;   q += n * 16;
        ds.s_r1 = (float4*)q;    q += n * 16;
        ds.s_pm1 = (float4*)q;   q += n * 16;
        ds.s_acc0 = (float2*)q;  q += n * 8;
        ds.s_acc1 = (float2*)q;
    }
    __shared__ uint32_t s_n_tasks;
    // ---- task table: per colour the records owned by this tile, cut into warps ----
    for (uint32_t c = threadIdx.x; c < nc; c += blockDim.x) {
        t_cbeg[c] = owner_rank(d, c, b0);
        t_cend[c] = (b1 < d.n_bodies) ? owner_rank(d, c, b1) : d.own_pos[(size_t)c * (d.own_words + 1u) + d.own_words];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t k = 0;
        for (uint32_t c = 0; c < nc; ++c)
            for (uint32_t r = t_cbeg[c]; r < t_cend[c]; r += 32u, ++k)
                if (k < max_tasks) {
                    t_first[k] = r;
                    t_count[k] = t_cend[c] - r < 32u ? t_cend[c] - r : 32u;
                }
        s_n_tasks = k;
        if (k > max_tasks || nb > TILE_MAX_BODIES) atomicOr(&d.counters->tile_fallback, 1u);
    }
    __syncthreads();
    const uint32_t n_tasks = s_n_tasks;
    grid.sync();
    if (__ldcg(&d.counters->tile_fallback) != 0u) return;  // nothing has been modified: the host runs k_solve_persistent
    // ---- stage the cached records; remember which body words are tile-local ----
    for (uint32_t k = warp; k < n_tasks && k < cache_tasks; k += n_warps) {
        if (lane >= t_count[k]) continue;
        const uint32_t m = t_first[k] + lane, l = k * 32u + lane;
        uint4 h = d.s_hdr[m];
        const bool st1 = (h.z & 0x100u) != 0, st2 = (h.z & 0x200u) != 0;
        if (!st1 && h.x >= b0 && h.x < b1 && d.body_shared[h.x] == 0u) h.z |= TILE_LOC1;
        if (!st2 && h.y >= b0 && h.y < b1 && d.body_shared[h.y] == 0u) h.z |= TILE_LOC2;
        ds.s_hdr[l] = h;
        ds.s_nf[l] = d.s_nf[m];
        ds.s_inv[l] = d.s_inv[m];
        ds.s_dep[l] = d.s_dep[m];
        ds.s_r0[l] = d.s_r0[m];
        ds.s_pm0[l] = d.s_pm0[m];
        ds.s_acc0[l] = d.s_acc0[m];
        if ((h.z & 0xFFu) > 1u) {
            ds.s_r1[l] = d.s_r1[m];
            ds.s_pm1[l] = d.s_pm1[m];
            ds.s_acc1[l] = d.s_acc1[m];
        }
    }
    for (uint32_t s = 0; s < S; ++s) {
        for (uint32_t i = threadIdx.x; i < nb; i += blockDim.x) {
            if (s == 0) integrate_forces_thread(d, b0 + i, sub_dt, S == 1);
            t_mom[i] = d.mom[b0 + i];
            t_ver[i] = 0u;
        }
        __syncthreads();
        grid.sync();
        for (uint32_t it = 0; it < I; ++it) {
            for (uint32_t k = warp; k < n_tasks; k += n_warps) {
                const bool cached = k < cache_tasks;
                const Dev& rd = cached ? ds : d;
                const uint32_t x = cached ? k * 32u + lane : t_first[k] + lane;
                bool pending = lane < t_count[k];
                // ---- record ----
                uint4 h = make_uint4(0u, 0u, 0u, 0u);
                ContactConst c;
                ContactPointConst pts[2];
                v2 acc[2];
                int np = 0;
                bool st1 = true, st2 = true, loc1 = false, loc2 = false;
                uint32_t e1 = 0, e2 = 0;
                float4 m1 = make_float4(0, 0, 0, 0), m2 = m1;
                if (pending) {
                    h = rd.s_hdr[x];
                    np = (int)(h.z & 0xFFu);
                    st1 = (h.z & 0x100u) != 0;
                    st2 = (h.z & 0x200u) != 0;
                    if (cached) {
                        loc1 = (h.z & TILE_LOC1) != 0;
                        loc2 = (h.z & TILE_LOC2) != 0;
                    } else {
                        loc1 = !st1 && h.x >= b0 && h.x < b1 && d.body_shared[h.x] == 0u;
                        loc2 = !st2 && h.y >= b0 && h.y < b1 && d.body_shared[h.y] == 0u;
                    }
                    const float4 nf = rd.s_nf[x], inv = rd.s_inv[x];
                    c.normal = mk2(nf.x, nf.y);
                    c.tangent = rot90cw(c.normal);
                    c.friction = nf.z;
                    c.inv_m1 = inv.x;
                    c.inv_m2 = inv.y;
                    c.inv_i1 = inv.z;
                    c.inv_i2 = inv.w;
                    {
                        const float4 r = rd.s_r0[x], pm = rd.s_pm0[x];
                        const float2 a = rd.s_acc0[x];
                        pts[0].r1 = mk2(r.x, r.y);
                        pts[0].r2 = mk2(r.z, r.w);
                        pts[0].mass_n = pm.x;
                        pts[0].mass_t = pm.y;
                        pts[0].depth = pm.z;
                        pts[0].bias = pm.w;
                        acc[0] = mk2(a.x, a.y);
                    }
                    if (np > 1) {
                        const float4 r = rd.s_r1[x], pm = rd.s_pm1[x];
                        const float2 a = rd.s_acc1[x];
                        pts[1].r1 = mk2(r.x, r.y);
                        pts[1].r2 = mk2(r.z, r.w);
                        pts[1].mass_n = pm.x;
                        pts[1].mass_t = pm.y;
                        pts[1].depth = pm.z;
                        pts[1].bias = pm.w;
                        acc[1] = mk2(a.x, a.y);
                    }
                    const uint4 dep = rd.s_dep[x];
                    e1 = it * dep.y + dep.x;
                    e2 = it * dep.w + dep.z;
                    if (st1) m1 = d.mom[h.x];  // static bodies are never written during a sweep
                    if (st2) m2 = d.mom[h.y];
                }
                // ---- wait for both bodies, update, publish ----
                uint32_t spins = 0;
                while (__any_sync(0xffffffffu, pending)) {
                    if (pending) {
                        uint32_t lag = 0u;
                        if (!st1) {
                            if (loc1) {
                                lag += e1 - ld_acquire_shared_u32(&t_ver[h.x - b0]);
                            } else {
                                m1 = ld_body_word(&d.mom[h.x]);
                                lag += e1 - f2u(m1.w);
                            }
                        }
                        if (!st2) {
                            if (loc2) {
                                lag += e2 - ld_acquire_shared_u32(&t_ver[h.y - b0]);
                            } else {
                                m2 = ld_body_word(&d.mom[h.y]);
                                lag += e2 - f2u(m2.w);
                            }
                        }
                        if (lag == 0u) {
                            // (the acquire loads of the versions above order these reads after the momentum stores they announce)
                            if (loc1) m1 = ld_volatile_shared_f4(&t_mom[h.x - b0]);
                            if (loc2) m2 = ld_volatile_shared_f4(&t_mom[h.y - b0]);
                            BodyVel v1 = {mk2(m1.x, m1.y), m1.z}, v2_ = {mk2(m2.x, m2.y), m2.z};
                            solve_contact(c, np, pts, acc, st1, st2, v1, v2_);
                            rd.s_acc0[x] = make_float2(acc[0].x, acc[0].y);
                            if (np > 1) rd.s_acc1[x] = make_float2(acc[1].x, acc[1].y);
                            if (loc1) st_volatile_shared_f4(&t_mom[h.x - b0], make_float4(v1.mom.x, v1.mom.y, v1.ang, 0.0f));
                            if (loc2) st_volatile_shared_f4(&t_mom[h.y - b0], make_float4(v2_.mom.x, v2_.mom.y, v2_.ang, 0.0f));
                            if (!st1 && !loc1) st_body_word(&d.mom[h.x], make_float4(v1.mom.x, v1.mom.y, v1.ang, u2f(e1 + 1u)));
                            if (!st2 && !loc2) st_body_word(&d.mom[h.y], make_float4(v2_.mom.x, v2_.mom.y, v2_.ang, u2f(e2 + 1u)));
                            if (loc1 || loc2) {
                                if (loc1) st_release_shared_u32(&t_ver[h.x - b0], e1 + 1u);
                                if (loc2) st_release_shared_u32(&t_ver[h.y - b0], e2 + 1u);
                            }
                            pending = false;
                        }
                    }
                    if ((++spins & 0xFFu) == 0u) {
                        // a stall would be a bug: flag it and let everybody run to the end (a CTA that returned early
                        // would leave the others waiting at the grid barrier)
                        if (spins > (1u << 20)) atomicOr(&d.counters->err, ERR_STALL);
                        if (*((volatile uint32_t*)&d.counters->err) & ERR_STALL) pending = false;
                    }
                }
            }
        }
        __syncthreads();
        grid.sync();
        // end of substep s fused with the start of substep s + 1 (per body, same thread)
        for (uint32_t i = threadIdx.x; i < nb; i += blockDim.x) {
            const uint32_t b = b0 + i;
            if (!(body_flags(d, b) & FLAG_STATIC) && d.body_shared[b] == 0u) d.mom[b] = t_mom[i];
            integrate_positions_thread(d, b, sub_dt);
            if (s + 1 < S) integrate_forces_thread(d, b, sub_dt, s + 2 == S);
        }
    }
    // accumulated impulses of the cached records are not needed after the call (collision.zig:102-133 re-creates them)
}

// ---- R2D_MODE_REFERENCE_ORDER (validation, slow): the substep loop of lib.zig:199-250 with the SEQUENTIAL Gauss-Seidel sweep
// of the reference — joints in list order, then manifolds in the order updateManifolds created them (`order`: pair slots,
// replayed on the host by reference_manifold_order) — by ONE thread of one CTA; the per-body phases use the whole CTA.
__global__ void __launch_bounds__(TPB) k_solve_reference_order(Dev d, float sub_dt, uint32_t S, uint32_t I,
                                                               const uint32_t* __restrict__ order, uint32_t n_order,
                                                               const uint32_t* __restrict__ joint_list, uint32_t n_joints) {
    if (overflowed(d) || d.counters->err != 0u) return;
    // preStep (collision.zig:102-133) into the record slot of the manifold's own pair slot
    for (uint32_t k = threadIdx.x; k < n_order; k += blockDim.x) gather_prestep_thread(d, order[k], order[k]);
    __syncthreads();
    for (uint32_t s = 0; s < S; ++s) {
        for (uint32_t i = threadIdx.x; i < d.n_bodies; i += blockDim.x) integrate_forces_thread(d, i, sub_dt, s + 1 == S);
        __syncthreads();
        if (threadIdx.x == 0)
            for (uint32_t it = 0; it < I; ++it) {   // lib.zig:226-236
                for (uint32_t j = 0; j < n_joints; ++j) solve_joint_thread(d, joint_list[j], sub_dt);
                for (uint32_t k = 0; k < n_order; ++k) solve_contact_thread<false>(d, order[k], sub_dt);
            }
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < d.n_bodies; i += blockDim.x) integrate_positions_thread(d, i, sub_dt);
        __syncthreads();
    }
}

// ---- boundary kernels: SoA export for bulk readback, force import ---------------------------------------------------------------
// `slot_of[k]` = device slot of the k-th requested body (host order): the device order is a spatial permutation
__global__ void __launch_bounds__(TPB) k_export_bodies(Dev d, const uint32_t* __restrict__ slot_of, uint32_t n, uint32_t* ids,
                                                       float2* pos, float* angle, float2* mom, float* ang_mom, float4* aabb) {
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const uint32_t s = slot_of[k];
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
__global__ void __launch_bounds__(TPB) k_import_forces(Dev d, const uint32_t* __restrict__ slot_of, uint32_t n,
                                                       const float* __restrict__ fxy_t) {
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const uint32_t s = slot_of[k];
        d.frc[s] = make_float4(fxy_t[3 * k], fxy_t[3 * k + 1], fxy_t[3 * k + 2], d.frc[s].w);
    }
}

// ---- spatial re-sort of the device slots, on the device (BatchBase::reorder; the host's build_image derives the same order) -------
// k_resort_min: per world the minimum position over its bodies without a NaN coordinate (ordered-int atomicMin, one per warp
// when the warp holds one world).  k_resort_keys: in HOST slot order (ties of the stable sort keep insertion order, as on the
// host) key = world << 32 | resort_key.  A library radix sort (cub::DeviceRadixSort, r2d_runtime.cu) orders them.
// k_resort_gather: device slot j takes the body whose host slot the sort put there; the inverse map is written on the way.
__global__ void __launch_bounds__(TPB) k_resort_init(int2* wmin, uint32_t n_worlds) {
    for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < n_worlds; w += gridDim.x * blockDim.x)
        wmin[w] = make_int2(RESORT_NO_MIN, RESORT_NO_MIN);
}
__global__ void __launch_bounds__(TPB) k_resort_min(Dev d, int2* wmin) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;   // (grid covers the bodies: whole warps reach the shuffles)
    const bool in = i < d.n_bodies;
    const float4 p = in ? d.pos[i] : make_float4(0, 0, 0, 0);
    const uint32_t w = in ? body_flags(d, i) >> FLAG_WORLD_SHIFT : 0xFFFFFFFFu;
    const bool ok = in && p.x == p.x && p.y == p.y;
    int kx = ok ? resort_float_order(p.x) : RESORT_NO_MIN, ky = ok ? resort_float_order(p.y) : RESORT_NO_MIN;
    const uint32_t w0 = __shfl_sync(0xffffffffu, w, 0);
    if (__all_sync(0xffffffffu, !in || w == w0)) {
        kx = cg::reduce(cg::tiled_partition<32>(cg::this_thread_block()), kx, cg::less<int>());
        ky = cg::reduce(cg::tiled_partition<32>(cg::this_thread_block()), ky, cg::less<int>());
        if ((threadIdx.x & 31u) == 0 && w0 != 0xFFFFFFFFu && kx != RESORT_NO_MIN) {
            atomicMin(&wmin[w0].x, kx);
            atomicMin(&wmin[w0].y, ky);
        }
    } else if (ok) {
        atomicMin(&wmin[w].x, kx);
        atomicMin(&wmin[w].y, ky);
    }
}
__global__ void __launch_bounds__(TPB) k_resort_keys(Dev d, const uint32_t* __restrict__ dev_of_host, const int2* __restrict__ wmin,
                                                     unsigned long long* keys, uint32_t* vals) {
    for (uint32_t h = blockIdx.x * blockDim.x + threadIdx.x; h < d.n_bodies; h += gridDim.x * blockDim.x) {
        const uint32_t s = dev_of_host[h];
        const float4 p = d.pos[s];
        const uint32_t w = body_flags(d, s) >> FLAG_WORLD_SHIFT;
        const int2 m = wmin[w];
        const float mx = m.x == RESORT_NO_MIN ? 0.0f : resort_float_unorder(m.x), my = m.x == RESORT_NO_MIN ? 0.0f : resort_float_unorder(m.y);
        keys[h] = ((unsigned long long)w << 32) | resort_key(p.x, p.y, mx, my);
        vals[h] = h;
    }
}
struct ResortArrays {
    float4 *pos, *mom, *frc, *prop, *shape, *aabb;
    uint32_t* sleep_cnt;
};
__global__ void __launch_bounds__(TPB) k_resort_gather(Dev d, const uint32_t* __restrict__ host_of_new, const uint32_t* __restrict__ old_dev_of_host,
                                                       uint32_t* new_dev_of_host, ResortArrays o) {
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < d.n_bodies; j += gridDim.x * blockDim.x) {
        const uint32_t h = host_of_new[j], s = old_dev_of_host[h];
        o.pos[j] = d.pos[s];
        o.mom[j] = d.mom[s];
        o.frc[j] = d.frc[s];
        o.prop[j] = d.prop[s];
        o.shape[j] = d.shape[s];
        o.aabb[j] = d.aabb[s];
        o.sleep_cnt[j] = d.sleep_cnt[s];
        new_dev_of_host[h] = j;
    }
}

}  // namespace r2d

#include "r2d_world.cuh"
