// Not built.  See README.md in this directory.
// ---- K2-K5 in ONE cooperative launch (fine grid, no dynamic large bodies): count -> scan -> fill -> pair count -> scan ->
// pair write separated by grid barriers (~1.5 us each) instead of six kernel boundaries, the two scans done in place by
// the resident CTAs (same decoupled look-back, tickets), and the cells of the few bodies that cover more than
// BIG_BODY_CELLS cells (the floor of a 100k-body pile: 1,100 cells) spread over the WHOLE grid instead of walked by the
// one CTA that happened to hold the body — that walk was the tail of both grid kernels (19.7 us for 5 MB).
constexpr uint32_t FUSED_BIG_LIST = 256;   // big bodies the grid can walk together (more: they are walked by their own thread)

__global__ void __launch_bounds__(TPB, 4) k_broad_fused(Dev d, unsigned long long* scan_a, unsigned long long* scan_b,
                                                        uint32_t state_tiles, uint32_t* big_list) {
    cg::grid_group grid = cg::this_grid();
    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, gth = gridDim.x * blockDim.x;
    // ---- count (SpatialHash.zig:46-49) ----
    for (uint32_t i = gtid; i < d.n_bodies; i += gth) {
        const CellRange r = count_body_thread(d, i, false);   // small bodies count their home cell themselves
        if (r.count == 0u) continue;
        const uint32_t slot = r.count > BIG_BODY_CELLS ? atomicAdd(&d.counters->n_work, 1u) : FUSED_BIG_LIST;
        if (slot < FUSED_BIG_LIST) {
            big_list[slot] = i;
        } else {
            for (uint32_t k = 0; k < r.count; ++k) atomicAdd(&d.bucket_cnt[cell_bucket(r, k)], 1u);
        }
    }
    grid.sync();
    const uint32_t n_big = min(__ldcg(&d.counters->n_work), FUSED_BIG_LIST);
    for (uint32_t q = 0; q < n_big; ++q) {
        const CellRange r = cell_range(d, __ldcg(&big_list[q]));
        for (uint32_t k = gtid; k < r.count; k += gth) atomicAdd(&d.bucket_cnt[cell_bucket(r, k)], 1u);
    }
    grid.sync();
    // ---- bucket starts (:52-57) ----
    {
        const uint32_t* in = d.bucket_cnt;
        while (scan_chained_body([in](uint32_t i) { return in[i]; }, d.bucket_start, 2u * d.n_buckets, scan_a, state_tiles,
                                 &d.counters->n_entries, nullptr)) {
        }
    }
    grid.sync();
    // ---- fill (:62-68) ----
    for (uint32_t i = gtid; i < d.n_bodies; i += gth) {
        if (body_is_small(d, body_flags(d, i))) {
            fill_fine(d, i);
            continue;
        }
        const CellRange r = cell_range(d, i);
        if (r.count > BIG_BODY_CELLS) {
            bool listed = false;
            for (uint32_t q = 0; q < n_big; ++q) listed = listed || __ldcg(&big_list[q]) == i;
            if (listed) continue;
        }
        for (uint32_t k = 0; k < r.count; ++k) fill_cell(d, i, cell_bucket(r, k));
    }
    for (uint32_t q = 0; q < n_big; ++q) {
        const uint32_t bi = __ldcg(&big_list[q]);
        const CellRange r = cell_range(d, bi);
        for (uint32_t k = gtid; k < r.count; k += gth) fill_cell(d, bi, cell_bucket(r, k));
    }
    grid.sync();
    // ---- pairs of every small body: count, park the first 8 partners ----
    const bool dead = overflowed(d);   // the grid did not fit: the attempt is abandoned (uniform)
    if (gtid == 0) d.pair_cnt[0] = 0u;
    for (uint32_t a = gtid; a < d.n_bodies; a += gth) {
        uint32_t got[8];
        const bool live = !dead && body_is_small(d, body_flags(d, a));
        const uint32_t n = live ? fine_body_pairs(d, a, got, nullptr) : 0u;
        d.pair_cnt[a + 1] = n;
        uint4* park = d.fine_cand + 2 * (size_t)a;
        if (n > 0u) park[0] = make_uint4(got[0], n > 1u ? got[1] : 0u, n > 2u ? got[2] : 0u, n > 3u ? got[3] : 0u);
        if (n > 4u) park[1] = make_uint4(got[4], n > 5u ? got[5] : 0u, n > 6u ? got[6] : 0u, n > 7u ? got[7] : 0u);
    }
    grid.sync();
    {
        uint32_t* cnt = d.pair_cnt;
        while (scan_chained_body([cnt](uint32_t i) { return cnt[i]; }, cnt, d.n_bodies + 1u, scan_b, state_tiles,
                                 &d.counters->n_pairs, nullptr)) {
        }
    }
    grid.sync();
    // ---- write, partners in ascending slot order (the list must not depend on the order the fill's atomics landed in) ----
    if (overflowed(d)) return;
    for (uint32_t a = gtid; a < d.n_bodies; a += gth) {
        const uint32_t at = d.pair_cnt[a + 1], n = d.pair_cnt[a + 2] - at;
        if (n == 0u || at + n > d.cap_pairs) continue;
        uint2* out = d.pairs + at;
        if (n <= 8u) {
            uint32_t got[8];
            const uint4* park = d.fine_cand + 2 * (size_t)a;
            const uint4 q0 = park[0], q1 = n > 4u ? park[1] : make_uint4(0u, 0u, 0u, 0u);
            got[0] = q0.x; got[1] = q0.y; got[2] = q0.z; got[3] = q0.w;
            got[4] = q1.x; got[5] = q1.y; got[6] = q1.z; got[7] = q1.w;
#pragma unroll
            for (int x = 1; x < 8; ++x) {
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

