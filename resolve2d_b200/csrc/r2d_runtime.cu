// r2d_runtime.cu — the CUDA backend of libr2d_b200.so: device memory, stream, the launch sequence of one
// `Solver.process` call (src/core/lib.zig:189-251) and the extern "C" ABI of include/r2d_abi.h (via r2d_capi.inc).
//
// One process() =
//   updateManifolds (lib.zig:253-299):  k_grid_cells<count> -> scan -> k_grid_cells<fill> -> k_sort_buckets ->
//        k_pairs<count> -> scan -> k_pairs<write> -> k_narrow -> k_color (cooperative) -> k_partition_prestep
//   one small device->host copy (counters + colour offsets), the only synchronisation point of the call
//   S x { k_integrate_forces ; I x { joint colours ; contact colours } ; k_integrate_positions }   (lib.zig:199-250)
// There is no CPU fallback anywhere: a missing device is an error.
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "r2d_host.hpp"
#include "r2d_kernels.cuh"

namespace {

using namespace r2d;
using host::BatchBase;
using host::BodyField;
using host::RawManifold;

thread_local std::string g_cuda_error;

#define R2D_CUDA(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess) {                                                                         \
            char _buf[512];                                                                              \
            snprintf(_buf, sizeof(_buf), "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            g_cuda_error = _buf;                                                                         \
            return _e == cudaErrorMemoryAllocation ? R2D_ERR_OUT_OF_MEMORY : R2D_ERR_CUDA;               \
        }                                                                                                \
    } while (0)

#define R2D_TRY(expr)                 \
    do {                              \
        const int _st = (expr);       \
        if (_st != R2D_OK) return _st; \
    } while (0)

// growable device array
template <class T>
struct DBuf {
    T* p = nullptr;
    size_t cap = 0;
    ~DBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    // contents are NOT preserved on growth
    int reserve(size_t n, bool zero = false, cudaStream_t st = 0) {
        if (n <= cap) return R2D_OK;
        release();
        const size_t want = n + n / 4 + 64;
        R2D_CUDA(cudaMalloc((void**)&p, want * sizeof(T)));
        cap = want;
        if (zero) R2D_CUDA(cudaMemsetAsync(p, 0, want * sizeof(T), st));
        return R2D_OK;
    }
};

struct PinnedStep {            // what the host learns about a step, one D2H copy
    Counters counters;
    uint32_t color_start[MAX_COLORS + 1];
};

struct CudaBatch : BatchBase {
    int device = 0;
    int n_sms = 148;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    // r2d_write_forces copies and scatters on a side stream, so that the host->device transfer of a step's inputs
    // overlaps the broadphase / narrowphase / colouring of that step; the solver launch (the first reader of `frc`)
    // and every other entry point wait for it
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t forces_ready = nullptr;
    bool forces_pending = false;
    bool zero_copy = true;            // R2D_ZERO_COPY=0: bulk reads always go through the staging buffer
    DBuf<unsigned char> staging_f;
    int color_blocks = 0, solve_blocks = 0, pair_blocks = 0;
    uint32_t solve_smem_slots = SOLVE_SMEM_SLOTS;
    bool test_small_buffers = false;  // R2D_TEST_SMALL_BUFFERS=1 (tests): start with buffers that are too small
    bool world_cache_forced = false;
    uint32_t world_rec_cap = 0;       // R2D_WORLD_CACHE: slots per world kept in shared memory by k_world_solve, see world_cache()
    uint32_t seen_world_m = 0;        // slots of the largest world of the previous call
    bool persistent_solver = true;
    bool world_solver = true;       // CTA-per-world shared-memory solver when every world is small and there are no joints
    uint32_t max_world_bodies = 0;   // false: one launch per colour (kept for A/B measurements)
    // bodies
    DBuf<float4> pos, mom, frc, prop, shape, aabb, pose, view;
    DBuf<uint32_t> ncells, world_base, grav_off, joint_color_start, dev_of_host, host_of_dev, world_joint_start, world_joint;
    DBuf<float> grav;
    DBuf<uint64_t> excl;
    DBuf<uint4> j_hdr, bkt, j_dep;
    DBuf<uint32_t> body_nj;
    bool joints_flow = true;          // R2D_JOINTS_FLOW=0: joints swept with a grid barrier per joint colour (A/B, tests)
    DBuf<float4> j_par, j_vec;
    // grid
    DBuf<uint32_t> bucket_cnt, bucket_start, ent_body, ent_key, ent_off, tile_sums, work, hit_bits;
    DBuf<int4> fcell;                 // fine grid: home cell per small body
    DBuf<float4> ent_aabb;            // fine grid: AABB copies in entry order
    DBuf<uint4> fine_cand;            // fine grid: partners parked between the count and the write pass
    DBuf<uint32_t> pair_cnt;          // fine grid: pairs per small body, then their scan
    bool seq_world_coloring = true;   // R2D_WORLD_COLORING=rounds: Jones-Plassmann rounds per world instead of sort + sequential greedy
    bool fine_grid = true;            // R2D_BROADPHASE=buckets: every body through the coarse buckets (the original pipeline)
    bool fine_now = false, ll_now = false;
    bool world_broad = true;          // k_world_broad for batches of small worlds (R2D_WORLD_BROAD=0: device-wide grid kernels)
    bool world_broad_declined = false;  // a world's grid did not fit shared memory: device-wide kernels until the next upload
    uint32_t max_world_large_cells = 0; // grid entries of the large bodies of the widest world (bound from the upload)
    // roadmap options: warm-start table (keyed by stable ids, survives re-uploads), sleep counters
    // (two tables: the pre-step of a call reads the previous call's, k_warm_save fills the other one; they swap when the call
    // has succeeded, so an abandoned attempt loses nothing)
    DBuf<unsigned long long> warm_key[2];
    DBuf<uint32_t> warm_meta[2], sleep_cnt, sleep_state;
    DBuf<float4> warm_val[2];
    DBuf<float2> s_warm0, s_warm1;
    uint32_t warm_slots = 0;
    int warm_cur = 0;
    bool warm_saving = false;   // fill_dev points the table at the one being written
    DBuf<uint32_t> pend_list;
    DBuf<uint64_t> world_magic;
    DBuf<uint32_t> big_bodies;
    uint32_t magic_for_mult = 0;
    DBuf<uint32_t> ref_order, ref_joints;   // R2D_MODE_REFERENCE_ORDER: manifold / joint sweep order of the reference
    bool world_fused_now = false;     // this step is solved by k_world_solve (which also places and pre-steps the manifolds)
    // pairs / manifolds
    DBuf<uint2> pairs;
    DBuf<uint4> m_hdr, s_hdr;
    DBuf<float4> m_g0, m_g1, m_r0, m_r1, s_nf, s_inv, s_r0, s_r1, s_pm0, s_pm1;
    DBuf<float2> s_acc0, s_acc1;
    DBuf<uint4> s_dep;
    DBuf<uint32_t> m_color;
    // colouring
    DBuf<unsigned long long> m_prio;
    // everything that must be zero at the start of a step lives in ONE allocation (one memset per step)
    DBuf<unsigned char> zeroed;
    size_t zeroed_bytes = 0, scan_state_cap = 0;
    size_t off_counters = 0, off_color_misc = 0, off_scan = 0, off_maxprio0 = 0, off_maxprio1 = 0, off_used = 0, off_own_bits = 0;
    size_t off_adj_cnt = 0, off_adj_head = 0, off_cstate = 0, off_body_shared = 0, off_pend_cnt = 0;
    bool tile_solver = true;          // k_solve_tiles for single worlds without joints that fit (R2D_TILE_SOLVER=0: never)
    uint32_t tile_bodies_now = 0, tile_max_tasks = TILE_MAX_TASKS;
    bool tile_declined = false;
    DBuf<unsigned long long> adj_prio;
    DBuf<uint4> adj_pool;   // chained entries of bodies with more than ADJ_CAP manifolds (dataflow colouring)
    int solve_wide = -1;              // k_solve_persistent with 512 threads per CTA: -1 by manifold count, 0 never, 1 always (R2D_SOLVE_WIDE)
    uint32_t solve_prefetch = 2;      // k_solve_persistent: streamed records fetched into L2 this many records ahead (R2D_SOLVE_PREFETCH)
    bool world_joints = true;         // worlds with joints through k_world_solve too (R2D_WORLD_JOINTS=0: the device-wide persistent sweep)
    bool world_single = true;         // one world of <= 1,024 bodies without joints: k_world_solve with one CTA of 512 threads (R2D_WORLD_SINGLE=0: tiles)
    bool world_export = true;         // k_world_solve exports into page-locked read-back arrays itself (R2D_WORLD_EXPORT=0: separate kernel)
    bool world_exported = false;
    bool flow_list_only = false;      // R2D_FLOW_LIST=1 (tests): the list flavour of the dataflow colouring for every size
    bool flow_coloring = true, flow_now = false;   // dataflow colouring of single worlds (R2D_FLOW_COLORING=0: rounds only)
    unsigned long long* scan_state(int which) { return (unsigned long long*)(zeroed.p + off_scan) + (size_t)which * scan_state_cap; }
    DBuf<uint32_t> own_pos;
    // staging for the boundary copies
    DBuf<unsigned char> staging;
    PinnedStep* pinned = nullptr;
    size_t cap_entries = 0, cap_pairs = 0;
    uint32_t last_pairs = 0;
    uint32_t last_manifolds = 0;   // M of the previous call (picks the wide persistent sweep; `stats` is reset at the start of a call)
    Dev d{};
    // profiling
    bool profiling = false;
    struct Ev {
        cudaEvent_t a, b;
        int kclass;
    };
    std::vector<Ev> events;
    std::vector<cudaEvent_t> event_pool;
    double prof_ms[R2D_KCLASS_COUNT] = {0};
    uint64_t prof_n[R2D_KCLASS_COUNT] = {0};
    uint32_t launches = 0;

    ~CudaBatch() override {
        cudaSetDevice(device);
        if (stream) cudaStreamSynchronize(stream);
        for (auto& e : events) {
            cudaEventDestroy(e.a);
            cudaEventDestroy(e.b);
        }
        for (auto& e : event_pool) cudaEventDestroy(e);
        if (pinned) cudaFreeHost(pinned);
        if (own_stream) cudaStreamDestroy(own_stream);
        if (copy_stream) cudaStreamDestroy(copy_stream);
        if (forces_ready) cudaEventDestroy(forces_ready);
    }

    int init(int dev) {
        device = dev;
        R2D_CUDA(cudaSetDevice(device));
        cudaDeviceProp prop_{};
        R2D_CUDA(cudaGetDeviceProperties(&prop_, device));
        if (prop_.major < 10) {
            g_cuda_error = "libr2d_b200 is built for sm_100a (B200) only; found compute capability " +
                           std::to_string(prop_.major) + "." + std::to_string(prop_.minor);
            return R2D_ERR_NO_DEVICE;
        }
        n_sms = prop_.multiProcessorCount;
        R2D_CUDA(cudaStreamCreateWithFlags(&own_stream, cudaStreamNonBlocking));
        stream = own_stream;
        R2D_CUDA(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
        R2D_CUDA(cudaEventCreateWithFlags(&forces_ready, cudaEventDisableTiming));
        R2D_CUDA(cudaHostAlloc((void**)&pinned, sizeof(PinnedStep), cudaHostAllocDefault));
        int per_sm = 0;
        R2D_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_color, TPB, 0));
        color_blocks = std::max(1, std::min(per_sm, 4)) * n_sms;
        R2D_CUDA(cudaFuncSetAttribute(k_solve_persistent<PSOLVE_TPB, SOLVE_SMEM_SLOTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SOLVE_SMEM_BYTES));
        R2D_CUDA(cudaFuncSetAttribute(k_solve_persistent<PSOLVE_TPB_BIG, SOLVE_SMEM_SLOTS_BIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SOLVE_SMEM_BYTES));
        solve_blocks = n_sms;   // one CTA per SM (the record cache takes the whole shared memory)
        R2D_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_bucket_count, TPB, 0));
        pair_blocks = (per_sm < 1 ? 1 : per_sm) * n_sms;  // resident CTAs: a CTA per heavy bucket, a warp per light one, grid-stride
        R2D_CUDA(cudaFuncSetAttribute(k_solve_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TILE_SMEM_BYTES));
#define R2D_WS_ATTR(BPT, WTPB)                                                                                                             \
    R2D_CUDA(cudaFuncSetAttribute(k_world_solve<BPT, WTPB, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WORLD_SMEM_MAX)); \
    R2D_CUDA(cudaFuncSetAttribute(k_world_solve<BPT, WTPB, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WORLD_SMEM_MAX));  \
    R2D_CUDA(cudaFuncSetAttribute(k_world_solve<BPT, WTPB, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WORLD_SMEM_MAX));  \
    R2D_CUDA(cudaFuncSetAttribute(k_world_solve<BPT, WTPB, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WORLD_SMEM_MAX))
        R2D_WS_ATTR(2, WORLD_SOLVE_TPB);
        R2D_WS_ATTR(4, WORLD_SOLVE_TPB);
        R2D_WS_ATTR(2, WORLD_SINGLE_TPB);
#undef R2D_WS_ATTR
        // Flavour switches for A/B measurements and tests.  Every one of them selects between paths that produce
        // bit-identical results; they are read here once, never inside process().
        auto env_is = [](const char* name, const char* value) {
            const char* e = getenv(name);
            return e && std::string(e) == value;
        };
        if (env_is("R2D_SOLVER", "launches")) persistent_solver = false;     // one launch per colour
        if (env_is("R2D_WORLD_SOLVER", "0")) world_solver = false;           // batches through the single-world kernels
        if (env_is("R2D_FLOW_COLORING", "0")) flow_coloring = false;         // Jones-Plassmann rounds only
        if (env_is("R2D_TILE_SOLVER", "0")) tile_solver = false;             // k_solve_persistent instead of k_solve_tiles
        if (env_is("R2D_FLOW_LIST", "1")) flow_list_only = true;
        if (env_is("R2D_WORLD_EXPORT", "0")) world_export = false;
        if (env_is("R2D_WORLD_SINGLE", "0")) world_single = false;
        if (env_is("R2D_WORLD_JOINTS", "0")) world_joints = false;
        if (const char* e = getenv("R2D_SOLVE_WIDE")) solve_wide = atoi(e);
        if (const char* e = getenv("R2D_SOLVE_PREFETCH")) solve_prefetch = (uint32_t)atoi(e);
        if (env_is("R2D_DEVICE_RESORT", "0")) device_resort = false;         // the periodic re-sort through the host
        if (env_is("R2D_RESORT_CHECK", "1")) resort_check = true;            // (tests) device order == build_image's order
        if (const char* e = getenv("R2D_TILE_MAX_TASKS")) tile_max_tasks = std::min<uint32_t>((uint32_t)atoi(e), TILE_MAX_TASKS);
        if (env_is("R2D_ZERO_COPY", "0")) zero_copy = false;                 // bulk reads through the staging buffer
        if (env_is("R2D_BROADPHASE", "buckets")) fine_grid = false;          // every body through the hashed 4 m buckets
        if (env_is("R2D_WORLD_COLORING", "rounds")) seq_world_coloring = false;
        if (env_is("R2D_WORLD_BROAD", "0")) world_broad = false;
        if (env_is("R2D_JOINTS_FLOW", "0")) joints_flow = false;
        R2D_CUDA(cudaFuncSetAttribute(k_world_broad, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WORLD_SMEM_MAX));
        if (env_is("R2D_TEST_SMALL_BUFFERS", "1")) test_small_buffers = true;
        if (const char* e = getenv("R2D_WORLD_CACHE")) {                     // tests: records per world kept in shared memory
            world_rec_cap = (uint32_t)atoi(e);
            world_cache_forced = true;
        }
        return R2D_OK;
    }

    int join_forces() {  // the main stream continues after the pending force import
        if (forces_pending) {
            R2D_CUDA(cudaStreamWaitEvent(stream, forces_ready, 0));
            forces_pending = false;
        }
        return R2D_OK;
    }

    // ---- launch helpers -------------------------------------------------------------------------------------------
    cudaEvent_t get_event() {
        if (!event_pool.empty()) {
            cudaEvent_t e = event_pool.back();
            event_pool.pop_back();
            return e;
        }
        cudaEvent_t e;
        cudaEventCreate(&e);
        return e;
    }
    void prof_begin(int kclass) {
        if (!profiling) return;
        Ev ev{get_event(), get_event(), kclass};
        cudaEventRecord(ev.a, stream);
        events.push_back(ev);
    }
    void prof_end() {
        if (!profiling) return;
        cudaEventRecord(events.back().b, stream);
    }
    int grid_warp_per(size_t n_items) const {  // one warp per item (bucket): short dependent chains, idle warps exit
        size_t blocks = (n_items * 32 + TPB - 1) / TPB;
        if (blocks < 1) blocks = 1;
        if (blocks > (size_t)1 << 30) blocks = (size_t)1 << 30;
        return (int)blocks;
    }
    int grid_for(size_t n, int tpb = TPB) const {  // fixed-shape grids: a multiple of the SM count, grid-stride inside
        size_t blocks = (n + tpb - 1) / tpb;
        const size_t cap = (size_t)n_sms * 8;
        if (blocks > cap) blocks = cap;
        if (blocks < 1) blocks = 1;
        return (int)blocks;
    }
#define R2D_LAUNCH(kclass, kernel, grid, block, ...)            \
    do {                                                        \
        prof_begin(kclass);                                     \
        kernel<<<(grid), (block), 0, stream>>>(__VA_ARGS__);    \
        prof_end();                                             \
        launches += 1;                                          \
    } while (0)

    // exclusive scan of in[0..n) into out[0..n), out[n] = total (also *total_out); n = min(*n_ptr, n_max) if n_ptr.
    // One chained-scan launch; `which` selects one of the pre-zeroed state regions of this step.
    int scan(const uint32_t* in, uint32_t* out, const uint32_t* n_ptr, uint32_t n_max, uint32_t* total_out, int kclass, int which) {
        const uint32_t tiles = (n_max + SCAN_TILE - 1) / SCAN_TILE + 1;
        if (tiles + 1 > scan_state_cap) return R2D_ERR_CUDA;
        unsigned long long* st = scan_state(which);
        R2D_LAUNCH(kclass, k_scan_chained, tiles, SCAN_TPB, in, out, n_ptr, n_max, st, (uint32_t)(scan_state_cap - 1), total_out);
        return R2D_OK;
    }

    // ---- BatchBase backend ----------------------------------------------------------------------------------------
    template <class T>
    int up(DBuf<T>& dst, const std::vector<T>& src, size_t min_elems = 1) {
        R2D_TRY(dst.reserve(std::max(src.size(), min_elems)));
        if (!src.empty()) R2D_CUDA(cudaMemcpyAsync(dst.p, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice, stream));
        return R2D_OK;
    }
    int backend_upload() override {
        R2D_CUDA(cudaSetDevice(device));
        R2D_CUDA(cudaStreamSynchronize(copy_stream));  // a pending force import reads buffers that are replaced below
        forces_pending = false;
        int st;
        if ((st = up(pos, image.pos)) || (st = up(mom, image.mom)) || (st = up(frc, image.frc)) || (st = up(prop, image.prop)) ||
            (st = up(shape, image.shape)) || (st = up(aabb, image.aabb)) || (st = up(world_base, image.world_base)) ||
            (st = up(grav_off, image.grav_off)) || (st = up(grav, image.grav)) || (st = up(sleep_cnt, image.sleep)) ||
            (st = up(dev_of_host, image.dev_of_host)) || (st = up(host_of_dev, image.host_of_dev)) || (st = upload_slot_tables()))
            return st;
        R2D_CUDA(cudaStreamSynchronize(stream));  // the image vectors are pageable and may change after we return
        magic_for_mult = 0;   // (the bucket count of a world depends on the mode's table multiplier: refreshed in process())
        after_new_order();
        max_world_bodies = 0;
        for (size_t w = 0; w + 1 < image.world_base.size(); ++w)
            max_world_bodies = std::max(max_world_bodies, image.world_base[w + 1] - image.world_base[w]);
        return R2D_OK;
    }
    // what names bodies by device slot besides the per-body arrays (host::build_slot_tables)
    int upload_slot_tables() {
        int st;
        if ((st = up(excl, image.excl)) || (st = up(j_hdr, image.j_hdr)) || (st = up(j_par, image.j_par)) || (st = up(j_vec, image.j_vec)) ||
            (st = up(j_dep, image.j_dep)) || (st = up(body_nj, image.body_nj)) || (st = up(joint_color_start, image.joint_color_start)) ||
            (st = up(world_joint_start, image.world_joint_start)) || (st = up(world_joint, image.world_joint)))
            return st;
        return R2D_OK;
    }
    void after_new_order() {   // results of the last step that name device slots are void
        last_pairs = 0;
        tile_declined = false;
        world_broad_declined = false;
    }
    // ---- re-sort on the device: keys, radix sort, permutation of the seven per-body arrays; 4 B per body come back ----
    DBuf<unsigned long long> resort_keys[2];
    DBuf<uint32_t> resort_vals[2], resort_doh, resort_sleep;
    DBuf<int2> resort_min;
    DBuf<float4> resort_f4[6];
    DBuf<unsigned char> resort_tmp;
    std::vector<uint32_t> resort_order;   // host slot at every new device slot
    bool device_resort = true;            // R2D_DEVICE_RESORT=0: through the host
    bool resort_check = false;            // R2D_RESORT_CHECK=1 (tests): the order must equal the one build_image derives
    int backend_reorder() override {
        const uint32_t nb = image.n_bodies, nw = (uint32_t)worlds.size();
        if (!device_resort || nb == 0) return REORDER_ON_HOST;
        R2D_CUDA(cudaSetDevice(device));
        R2D_TRY(join_forces());   // (a pending force import scatters through the old dev_of_host)
        for (int k = 0; k < 2; ++k) {
            R2D_TRY(resort_keys[k].reserve(nb));
            R2D_TRY(resort_vals[k].reserve(nb));
        }
        for (auto& b : resort_f4) R2D_TRY(b.reserve(nb));   // (swapped with the body arrays below: every buffer holds >= nb)
        R2D_TRY(resort_sleep.reserve(nb));
        R2D_TRY(resort_doh.reserve(nb));
        R2D_TRY(resort_min.reserve(nw));
        int world_bits = 0;
        while (world_bits < 31 && (1u << world_bits) < nw) ++world_bits;
        size_t tmp_bytes = 0;
        R2D_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, resort_keys[0].p, resort_keys[1].p, resort_vals[0].p, resort_vals[1].p,
                                                 (int)nb, 0, 32 + world_bits, stream));
        R2D_TRY(resort_tmp.reserve(tmp_bytes + 256));
        fill_dev();
        k_resort_init<<<grid_for(nw), TPB, 0, stream>>>(resort_min.p, nw);
        k_resort_min<<<(nb + TPB - 1) / TPB, TPB, 0, stream>>>(d, resort_min.p);
        k_resort_keys<<<grid_for(nb), TPB, 0, stream>>>(d, dev_of_host.p, resort_min.p, resort_keys[0].p, resort_vals[0].p);
        R2D_CUDA(cub::DeviceRadixSort::SortPairs(resort_tmp.p, tmp_bytes, resort_keys[0].p, resort_keys[1].p, resort_vals[0].p, resort_vals[1].p,
                                                 (int)nb, 0, 32 + world_bits, stream));
        const ResortArrays out = {resort_f4[0].p, resort_f4[1].p, resort_f4[2].p, resort_f4[3].p, resort_f4[4].p, resort_f4[5].p, resort_sleep.p};
        k_resort_gather<<<grid_for(nb), TPB, 0, stream>>>(d, resort_vals[1].p, dev_of_host.p, resort_doh.p, out);
        resort_order.resize(nb);
        R2D_CUDA(cudaMemcpyAsync(resort_order.data(), resort_vals[1].p, (size_t)nb * 4, cudaMemcpyDeviceToHost, stream));
        R2D_CUDA(cudaStreamSynchronize(stream));
        R2D_CUDA(cudaGetLastError());
        auto swap_buf = [](auto& a, auto& b) {
            std::swap(a.p, b.p);
            std::swap(a.cap, b.cap);
        };
        swap_buf(pos, resort_f4[0]); swap_buf(mom, resort_f4[1]); swap_buf(frc, resort_f4[2]); swap_buf(prop, resort_f4[3]);
        swap_buf(shape, resort_f4[4]); swap_buf(aabb, resort_f4[5]); swap_buf(sleep_cnt, resort_sleep); swap_buf(dev_of_host, resort_doh);
        R2D_TRY(host_of_dev.reserve(nb));
        R2D_CUDA(cudaMemcpyAsync(host_of_dev.p, resort_vals[1].p, (size_t)nb * 4, cudaMemcpyDeviceToDevice, stream));
        const int bt = adopt_device_order(resort_order);   // host side: maps, shape image, exclusions and joints by device slot
        if (bt != R2D_OK) return bt;
        R2D_TRY(upload_slot_tables());
        R2D_CUDA(cudaStreamSynchronize(stream));
        after_new_order();
        if (resort_check) {   // the order build_image would derive from the same state
            R2D_TRY(backend_download());
            host::Image ref;
            const int st = host::build_image(worlds, ref, grid_cell());
            if (st != R2D_OK) return st;
            if (ref.host_of_dev != image.host_of_dev) return R2D_ERR_BAD_STATE;
        }
        return R2D_OK;
    }
    bool backend_reorder_pays() const override {
        const bool small_worlds = world_solver && worlds.size() >= (size_t)n_sms / 2 && max_world_bodies <= WORLD_MAX_BODIES;
        return !small_worlds;
    }
    int backend_download() override {
        R2D_CUDA(cudaSetDevice(device));
        const size_t nb = image.n_bodies;
        if (nb == 0) return R2D_OK;
        std::vector<float4> hp(nb), hm(nb), hf(nb), ha(nb);
        std::vector<uint32_t> hs(opt_sleeping ? nb : 0);
        R2D_TRY(join_forces());
        if (opt_sleeping) R2D_CUDA(cudaMemcpyAsync(hs.data(), sleep_cnt.p, nb * 4, cudaMemcpyDeviceToHost, stream));
        R2D_CUDA(cudaMemcpyAsync(hp.data(), pos.p, nb * 16, cudaMemcpyDeviceToHost, stream));
        R2D_CUDA(cudaMemcpyAsync(hm.data(), mom.p, nb * 16, cudaMemcpyDeviceToHost, stream));
        R2D_CUDA(cudaMemcpyAsync(hf.data(), frc.p, nb * 16, cudaMemcpyDeviceToHost, stream));
        R2D_CUDA(cudaMemcpyAsync(ha.data(), aabb.p, nb * 16, cudaMemcpyDeviceToHost, stream));
        R2D_CUDA(cudaStreamSynchronize(stream));
        for (auto& w : worlds) {
            const uint32_t base = image.world_base[w->index];
            for (size_t s = 0; s < w->bodies.size(); ++s) {
                host::Body& b = w->bodies[s];
                const uint32_t ds = image.dev_of_host[base + s];
                const float4 p = hp[ds], m = hm[ds], f = hf[ds], a = ha[ds];
                b.pos_x = p.x; b.pos_y = p.y; b.angle = p.z;
                b.mom_x = m.x; b.mom_y = m.y; b.ang_mom = m.z;
                b.force_x = f.x; b.force_y = f.y; b.torque = f.z;
                b.aabb_x = a.x; b.aabb_y = a.y; b.aabb_hw = a.z; b.aabb_hh = a.w;
                if (opt_sleeping) b.sleep_counter = hs[ds];
            }
        }
        return R2D_OK;
    }
    int backend_write(uint32_t gslot, BodyField f, int comp, int n, const float* v) override {
        R2D_CUDA(cudaSetDevice(device));
        R2D_TRY(join_forces());
        float4* arr = f == host::FIELD_POS ? pos.p : (f == host::FIELD_MOM ? mom.p : frc.p);
        float* dst = reinterpret_cast<float*>(arr + gslot) + comp;
        R2D_CUDA(cudaMemcpyAsync(dst, v, (size_t)n * 4, cudaMemcpyHostToDevice, stream));
        R2D_CUDA(cudaStreamSynchronize(stream));  // `v` lives on the caller's stack
        return R2D_OK;
    }
    // device address of a page-locked (UVA-mapped) host buffer, or null (pageable memory, R2D_ZERO_COPY=0)
    void* mapped_host(void* host) const {
        if (!host || !zero_copy) return nullptr;
        cudaPointerAttributes a{};
        if (cudaPointerGetAttributes(&a, host) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        return a.type == cudaMemoryTypeHost ? a.devicePointer : nullptr;
    }
    // enqueues the export (kernel + copies) of bodies [first, first + n) on the main stream; the caller synchronises
    int enqueue_read_bodies(uint32_t first, uint32_t n, uint32_t* ids, float* pos_xy, float* angle, float* momentum_xy,
                            float* ang_momentum, float* aabb_xywh) {
        R2D_CUDA(cudaSetDevice(device));
        // repack the float4 SoA into the caller's layout on the device, then one copy per requested array
        R2D_TRY(staging.reserve((size_t)n * 44 + 256));
        unsigned char* s = staging.p;
        size_t off = 0;
        auto carve = [&](size_t bytes) {  // 16-byte aligned sub-buffers
            unsigned char* q = s + off;
            off += (bytes + 15) & ~(size_t)15;
            return q;
        };
        uint32_t* d_ids = (uint32_t*)carve((size_t)n * 4);
        float2* d_pos = (float2*)carve((size_t)n * 8);
        float* d_ang = (float*)carve((size_t)n * 4);
        float2* d_mom = (float2*)carve((size_t)n * 8);
        float* d_l = (float*)carve((size_t)n * 4);
        float4* d_aabb = (float4*)carve((size_t)n * 16);
        fill_dev();
        // Pinned (page-locked, UVA-mapped) destinations are written by the export kernel itself: the stores stream over
        // PCIe as the warps finish, instead of one kernel followed by up to six separate DMA copies.  R2D_ZERO_COPY=0 or
        // any pageable destination: repack into the staging buffer and copy.
        auto mapped = [&](void* host) -> void* { return mapped_host(host); };
        void* m_ids = mapped(ids); void* m_pos = mapped(pos_xy); void* m_ang = mapped(angle); void* m_mom = mapped(momentum_xy);
        void* m_l = mapped(ang_momentum); void* m_aabb = mapped(aabb_xywh);
        const bool direct = (!ids || m_ids) && (!pos_xy || m_pos) && (!angle || m_ang) && (!momentum_xy || m_mom) &&
                            (!ang_momentum || m_l) && (!aabb_xywh || m_aabb);
        if (direct) {
            R2D_LAUNCH(R2D_KCLASS_INTEGRATE, k_export_bodies, grid_for(n), TPB, d, (const uint32_t*)dev_of_host.p + first, n, (uint32_t*)m_ids,
                       (float2*)m_pos, (float*)m_ang, (float2*)m_mom, (float*)m_l, (float4*)m_aabb);
            return R2D_OK;
        }
        R2D_LAUNCH(R2D_KCLASS_INTEGRATE, k_export_bodies, grid_for(n), TPB, d, (const uint32_t*)dev_of_host.p + first, n, ids ? d_ids : nullptr,
                   pos_xy ? d_pos : nullptr, angle ? d_ang : nullptr, momentum_xy ? d_mom : nullptr,
                   ang_momentum ? d_l : nullptr, aabb_xywh ? d_aabb : nullptr);
        if (ids) R2D_CUDA(cudaMemcpyAsync(ids, d_ids, (size_t)n * 4, cudaMemcpyDeviceToHost, stream));
        if (pos_xy) R2D_CUDA(cudaMemcpyAsync(pos_xy, d_pos, (size_t)n * 8, cudaMemcpyDeviceToHost, stream));
        if (angle) R2D_CUDA(cudaMemcpyAsync(angle, d_ang, (size_t)n * 4, cudaMemcpyDeviceToHost, stream));
        if (momentum_xy) R2D_CUDA(cudaMemcpyAsync(momentum_xy, d_mom, (size_t)n * 8, cudaMemcpyDeviceToHost, stream));
        if (ang_momentum) R2D_CUDA(cudaMemcpyAsync(ang_momentum, d_l, (size_t)n * 4, cudaMemcpyDeviceToHost, stream));
        if (aabb_xywh) R2D_CUDA(cudaMemcpyAsync(aabb_xywh, d_aabb, (size_t)n * 16, cudaMemcpyDeviceToHost, stream));
        return R2D_OK;
    }
    int backend_read_bodies(uint32_t first, uint32_t n, uint32_t* ids, float* pos_xy, float* angle, float* momentum_xy,
                            float* ang_momentum, float* aabb_xywh) override {
        R2D_TRY(enqueue_read_bodies(first, n, ids, pos_xy, angle, momentum_xy, ang_momentum, aabb_xywh));
        R2D_CUDA(cudaStreamSynchronize(stream));
        return R2D_OK;
    }
    int backend_write_forces(uint32_t first, uint32_t n, const float* f) override {
        R2D_CUDA(cudaSetDevice(device));
        if (staging_f.cap < (size_t)n * 12 + 256) {
            R2D_CUDA(cudaStreamSynchronize(copy_stream));  // an earlier import may still read the old buffer
            R2D_TRY(staging_f.reserve((size_t)n * 12 + 256));
        }
        // the main stream is idle here (every entry point synchronises it before returning), so the side stream may
        // touch `frc` right away; two imports in a row are ordered by the side stream itself
        R2D_CUDA(cudaMemcpyAsync(staging_f.p, f, (size_t)n * 12, cudaMemcpyHostToDevice, copy_stream));
        fill_dev();
        k_import_forces<<<grid_for(n), TPB, 0, copy_stream>>>(d, (const uint32_t*)dev_of_host.p + first, n, (const float*)staging_f.p);
        launches += 1;
        R2D_CUDA(cudaEventRecord(forces_ready, copy_stream));
        forces_pending = true;
        // pageable source buffers are consumed before cudaMemcpyAsync returns; pinned ones must stay valid until the
        // next synchronising call (r2d_process synchronises once per step)
        return R2D_OK;
    }
    int backend_read_pairs(std::vector<uint2>& out) override {
        R2D_CUDA(cudaSetDevice(device));
        out.resize(last_pairs);
        if (last_pairs) {
            R2D_CUDA(cudaMemcpyAsync(out.data(), pairs.p, (size_t)last_pairs * 8, cudaMemcpyDeviceToHost, stream));
            R2D_CUDA(cudaStreamSynchronize(stream));
        }
        return R2D_OK;
    }
    int backend_read_manifolds(std::vector<RawManifold>& out) override {
        R2D_CUDA(cudaSetDevice(device));
        out.clear();
        const size_t n = last_pairs;
        if (!n) return R2D_OK;
        std::vector<uint4> h(n);
        std::vector<float4> g0(n), g1(n), r0(n), r1(n);
        std::vector<uint32_t> col(n);
        R2D_CUDA(cudaMemcpyAsync(h.data(), m_hdr.p, n * 16, cudaMemcpyDeviceToHost, stream));
        R2D_CUDA(cudaMemcpyAsync(g0.data(), m_g0.p, n * 16, cudaMemcpyDeviceToHost, stream));
        R2D_CUDA(cudaMemcpyAsync(g1.data(), m_g1.p, n * 16, cudaMemcpyDeviceToHost, stream));
        R2D_CUDA(cudaMemcpyAsync(r0.data(), m_r0.p, n * 16, cudaMemcpyDeviceToHost, stream));
        R2D_CUDA(cudaMemcpyAsync(r1.data(), m_r1.p, n * 16, cudaMemcpyDeviceToHost, stream));
        R2D_CUDA(cudaMemcpyAsync(col.data(), m_color.p, n * 4, cudaMemcpyDeviceToHost, stream));
        R2D_CUDA(cudaStreamSynchronize(stream));
        for (size_t p = 0; p < n; ++p) {
            if (col[p] == COLOR_NONE) continue;
            RawManifold m{};
            m.ref = h[p].x;
            m.inc = h[p].y;
            m.n_points = h[p].z & 0xFF;
            m.normal_id = h[p].z >> 8;
            m.color = col[p];
            m.normal_x = g0[p].x; m.normal_y = g0[p].y;
            m.pos_x[0] = g0[p].z; m.pos_y[0] = g0[p].w;
            m.depth[0] = g1[p].x; m.depth[1] = g1[p].y;
            m.pos_x[1] = g1[p].z; m.pos_y[1] = g1[p].w;
            m.ref_rx[0] = r0[p].x; m.ref_ry[0] = r0[p].y; m.inc_rx[0] = r0[p].z; m.inc_ry[0] = r0[p].w;
            m.ref_rx[1] = r1[p].x; m.ref_ry[1] = r1[p].y; m.inc_rx[1] = r1[p].z; m.inc_ry[1] = r1[p].w;
            out.push_back(m);
        }
        return R2D_OK;
    }
    int backend_sync() override {
        R2D_CUDA(cudaSetDevice(device));
        R2D_CUDA(cudaStreamSynchronize(stream));
        return R2D_OK;
    }
    int backend_set_stream(void* s) override {
        R2D_CUDA(cudaSetDevice(device));
        R2D_TRY(join_forces());
        R2D_CUDA(cudaStreamSynchronize(stream));
        stream = s ? (cudaStream_t)s : own_stream;
        return R2D_OK;
    }
    int backend_profile_enable(int on) override {
        profiling = on != 0;
        return R2D_OK;
    }
    int backend_profile_read(double* ms, uint64_t* n, int reset) override {
        R2D_CUDA(cudaSetDevice(device));
        R2D_CUDA(cudaStreamSynchronize(stream));
        for (auto& e : events) {
            float t = 0;
            if (cudaEventElapsedTime(&t, e.a, e.b) == cudaSuccess) {
                prof_ms[e.kclass] += t;
                prof_n[e.kclass] += 1;
            }
            event_pool.push_back(e.a);
            event_pool.push_back(e.b);
        }
        events.clear();
        for (int k = 0; k < R2D_KCLASS_COUNT; ++k) {
            if (ms) ms[k] = prof_ms[k];
            if (n) n[k] = prof_n[k];
            if (reset) {
                prof_ms[k] = 0;
                prof_n[k] = 0;
            }
        }
        return R2D_OK;
    }

    void fill_dev() {
        d.n_bodies = image.n_bodies;
        d.pos = pos.p; d.mom = mom.p; d.frc = frc.p; d.prop = prop.p; d.shape = shape.p; d.aabb = aabb.p;
        d.pose = pose.p; d.view = view.p; d.ncells = ncells.p; d.bkt = bkt.p;
        d.n_worlds = (uint32_t)worlds.size();
        d.world_base = world_base.p; d.grav_off = grav_off.p; d.grav = grav.p; d.world_magic = world_magic.p; d.big_bodies = big_bodies.p;
        d.cell = grid_cell(); d.table_mult = grid_mult();
        d.n_buckets = d.table_mult * d.n_bodies;
        d.bucket_cnt = bucket_cnt.p; d.bucket_start = bucket_start.p;
        d.cap_entries = (uint32_t)cap_entries;
        d.ent_body = ent_body.p; d.ent_key = ent_key.p; d.ent_off = ent_off.p; d.work = work.p; d.hit_bits = hit_bits.p;
        d.excl = (const uint64_t*)excl.p; d.n_excl = (uint32_t)image.excl.size();
        d.fine_on = fine_now ? 1u : 0u; d.ll_on = ll_now ? 1u : 0u;
        d.fine_inv = fine_now ? 1.0 / (double)image.fine_cell : 0.0;
        d.fcell = fcell.p; d.pair_cnt = pair_cnt.p; d.ent_aabb = ent_aabb.p; d.fine_cand = fine_cand.p;
        d.cap_pairs = (uint32_t)cap_pairs;
        d.pairs = pairs.p; d.m_hdr = m_hdr.p; d.m_g0 = m_g0.p; d.m_g1 = m_g1.p; d.m_r0 = m_r0.p; d.m_r1 = m_r1.p;
        d.m_color = m_color.p; d.m_prio = m_prio.p;
        d.maxprio0 = (unsigned long long*)(zeroed.p + off_maxprio0);
        d.maxprio1 = (unsigned long long*)(zeroed.p + off_maxprio1);
        d.used = (unsigned long long*)(zeroed.p + off_used);
        uint32_t* color_misc = (uint32_t*)(zeroed.p + off_color_misc);
        d.color_count = color_misc;
        d.color_start = color_misc + MAX_COLORS;
        d.color_cursor = color_misc + 2 * MAX_COLORS + 1;
        d.round_left = color_misc + 3 * MAX_COLORS + 1;
        d.counters = (Counters*)(zeroed.p + off_counters);
        d.flow = flow_now ? 1u : 0u;
        d.adj_cnt = (uint32_t*)(zeroed.p + off_adj_cnt);
        d.adj_head = (uint32_t*)(zeroed.p + off_adj_head);
        d.adj_pool = adj_pool.p;
        d.adj_pool_cap = (uint32_t)std::min<size_t>(adj_pool.cap, 0x7FFFFFFFu);
        d.cstate = (uint4*)(zeroed.p + off_cstate);
        d.adj_prio = adj_prio.p;
        d.tile_bodies = tile_bodies_now;
        d.world_fused = world_fused_now ? 1u : 0u;
        d.warm_on = opt_warm_start ? 1u : 0u;
        d.warm_mask = warm_slots ? warm_slots - 1u : 0u;
        {
            const int t = warm_saving ? 1 - warm_cur : warm_cur;
            d.warm_key = warm_key[t].p; d.warm_meta = warm_meta[t].p; d.warm_val = warm_val[t].p;
        }
        d.s_warm0 = s_warm0.p; d.s_warm1 = s_warm1.p;
        d.sleep_cnt = sleep_cnt.p; d.sleep_state = sleep_state.p;
        d.body_shared = (uint32_t*)(zeroed.p + off_body_shared);
        d.pend_cnt = (uint32_t*)(zeroed.p + off_pend_cnt);
        d.pend_list = pend_list.p;
        d.own_words = (d.n_bodies + 31u) / 32u;
        d.own_bits = (uint32_t*)(zeroed.p + off_own_bits);
        d.own_pos = own_pos.p;
        d.s_hdr = s_hdr.p; d.s_nf = s_nf.p; d.s_inv = s_inv.p; d.s_r0 = s_r0.p; d.s_r1 = s_r1.p;
        d.s_pm0 = s_pm0.p; d.s_pm1 = s_pm1.p; d.s_acc0 = s_acc0.p; d.s_acc1 = s_acc1.p;
        d.s_dep = s_dep.p;
        d.n_joints = (uint32_t)image.j_hdr.size();
        d.j_hdr = j_hdr.p; d.j_par = j_par.p; d.j_vec = j_vec.p;
        d.n_joint_colors = (uint32_t)image.joint_color_start.size() - 1;
        d.world_joint_start = world_joint_start.p; d.world_joint = world_joint.p;
        d.j_dep = j_dep.p; d.body_nj = body_nj.p;
        // (a sleeper is a static body for the call, and two-body joints write a static body's momentum, Q10: barrier sweep)
        d.joints_flow = (joints_flow && image.joints_flow_ok && mode != R2D_MODE_REFERENCE_ORDER && !opt_sleeping) ? 1u : 0u;
        d.color_smem = 0;
    }

    int reserve_entries(size_t n) {
        int st;
        if ((st = ent_body.reserve(n)) || (st = ent_key.reserve(n)) || (st = ent_aabb.reserve(n))) return st;
        cap_entries = std::min(std::min(ent_body.cap, ent_key.cap), ent_aabb.cap);
        if ((st = hit_bits.reserve(HIT_WORDS_PER_ENTRY * cap_entries + 64))) return st;  // ballots of the staged buckets
        return R2D_OK;
    }
    int reserve_pairs(size_t n) {
        int st;
        const size_t pad = (size_t)MAX_COLORS * COLOR_ALIGN;  // colour segments are padded to whole warps
        n += pad;
        if ((st = pairs.reserve(n)) || (st = m_hdr.reserve(n)) || (st = m_g0.reserve(n)) || (st = m_g1.reserve(n)) ||
            (st = m_r0.reserve(n)) || (st = m_r1.reserve(n)) || (st = m_color.reserve(n)) || (st = m_prio.reserve(n)) || (st = s_hdr.reserve(n)) ||
            (st = s_nf.reserve(n)) || (st = s_inv.reserve(n)) || (st = s_r0.reserve(n)) || (st = s_r1.reserve(n)) ||
            (st = s_pm0.reserve(n)) || (st = s_pm1.reserve(n)) || (st = s_acc0.reserve(n)) || (st = s_acc1.reserve(n)) ||
            (st = s_dep.reserve(n)) || (st = pend_list.reserve(2 * n)) || (opt_warm_start && ((st = s_warm0.reserve(n)) || (st = s_warm1.reserve(n)))))
            return st;
        cap_pairs = pairs.cap;
        for (size_t c : {m_prio.cap, m_hdr.cap, m_g0.cap, m_g1.cap, m_r0.cap, m_r1.cap, m_color.cap, s_hdr.cap, s_nf.cap, s_inv.cap,
                         s_r0.cap, s_r1.cap, s_pm0.cap, s_pm1.cap, s_acc0.cap, s_acc1.cap, s_dep.cap})
            cap_pairs = std::min(cap_pairs, c);
        cap_pairs -= pad;
        return R2D_OK;
    }

    // Shared-memory budget of k_world_solve: room for the slots (one per contact point) of the largest world of the
    // previous call plus 1/32.  A world that still does not fit keeps its slots in its slice of the global record
    // arrays, so this is a performance choice only.
    static constexpr size_t WORLD_SMEM_MAX = 227 * 1024 - 4096;   // (the kernel also has 3 KB of static shared memory)
    size_t world_cache(uint32_t& nb_cap, uint32_t& R) {
        nb_cap = (max_world_bodies + 3u) & ~3u;
        if (world_cache_forced)
            R = world_rec_cap;
        else if (seen_world_m == 0)   // first call: a guess (about two contact points per body in a settled box)
            R = 2 * nb_cap + 32;
        else
            R = seen_world_m + seen_world_m / 32 + 4;
        R = (R + 3u) & ~3u;
        const bool joints = !image.j_hdr.empty();
        while (world_smem_bytes(nb_cap, R, joints) > WORLD_SMEM_MAX && R > 0) R = R / 2 & ~3u;
        return world_smem_bytes(nb_cap, R, joints);
    }

    int launch_persistent(float sub_dt, uint32_t S, uint32_t I) {
        prof_begin(R2D_KCLASS_SOLVE_CONTACTS);
        const uint32_t* jcs_dev = joint_color_start.p;
        uint32_t n_jc = (uint32_t)image.joint_color_start.size() - 1, S_ = S, I_ = I;
        float sd = sub_dt;
        // most records beyond the cache (sized from the previous call's manifold count): the wide flavour
        const bool wide = solve_wide == 1 || (solve_wide < 0 && (size_t)last_manifolds > 4 * (size_t)SOLVE_SMEM_SLOTS * PSOLVE_TPB * solve_blocks);
        uint32_t slots = wide ? std::min<uint32_t>(solve_smem_slots, SOLVE_SMEM_SLOTS_BIG) : solve_smem_slots, prefetch = solve_prefetch;
        void* args[] = {(void*)&d, (void*)&sd, (void*)&S_, (void*)&I_, (void*)&jcs_dev, (void*)&n_jc, (void*)&slots, (void*)&prefetch};
        if (wide)
            R2D_CUDA(cudaLaunchCooperativeKernel((void*)k_solve_persistent<PSOLVE_TPB_BIG, SOLVE_SMEM_SLOTS_BIG>, dim3(solve_blocks),
                                                 dim3(PSOLVE_TPB_BIG), args, SOLVE_SMEM_BYTES, stream));
        else
            R2D_CUDA(cudaLaunchCooperativeKernel((void*)k_solve_persistent<PSOLVE_TPB, SOLVE_SMEM_SLOTS>, dim3(solve_blocks), dim3(PSOLVE_TPB),
                                                 args, SOLVE_SMEM_BYTES, stream));
        prof_end();
        launches += 1;
        return R2D_OK;
    }

    int backend_process(float dt, uint32_t S, uint32_t I) override {
        R2D_CUDA(cudaSetDevice(device));
        const uint32_t nb = image.n_bodies;
        launches = 0;
        stats = r2d_step_stats{};
        stats.n_bodies = nb;
        stats.n_joints = (uint32_t)image.j_hdr.size();
        stats.n_joint_colors = (uint32_t)image.joint_color_start.size() - 1;
        if (nb == 0) return R2D_OK;
        R2D_TRY(big_bodies.reserve(BIG_GLOBAL_LIST));
        if (magic_for_mult != grid_mult()) {   // per world: the reciprocal that replaces the 64-bit modulo of the cell hash
            std::vector<uint64_t> mg(worlds.size());
            for (size_t w = 0; w < worlds.size(); ++w)
                mg[w] = hash_magic((uint64_t)grid_mult() * (image.world_base[w + 1] - image.world_base[w]));
            R2D_TRY(world_magic.reserve(mg.size()));
            R2D_CUDA(cudaMemcpyAsync(world_magic.p, mg.data(), mg.size() * 8, cudaMemcpyHostToDevice, stream));
            R2D_CUDA(cudaStreamSynchronize(stream));
            magic_for_mult = grid_mult();
        }
        const float sub_dt = dt / (float)S;  // lib.zig:190-191 (host f32 division, IEEE)
        d.sub_dt = sub_dt;
        const uint32_t T = grid_mult() * nb;
        int st;
        const size_t own_w = ((size_t)nb + 31) / 32;
        {   // layout of the zeroed arena for this body count
            auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
            const size_t max_scan = std::max<size_t>(2 * (size_t)T + 1, (own_w + 1) * MAX_COLORS + 2);
            scan_state_cap = std::max((max_scan + SCAN_TILE - 1) / SCAN_TILE + 4, worlds.size() + 4);   // (k_world_broad: one word per world)
            size_t o = 0;
            off_counters = o; o = align(o + sizeof(Counters));
            off_color_misc = o; o = align(o + (MAX_COLORS * 3 + 1 + MAX_COLOR_ROUNDS) * 4);
            off_scan = o; o = align(o + 4 * scan_state_cap * 8);
            off_maxprio0 = o; o = align(o + (size_t)nb * 8);
            off_maxprio1 = o; o = align(o + (size_t)nb * 8);
            off_used = o; o = align(o + (size_t)nb * COLOR_WORDS * 8);
            off_own_bits = o; o = align(o + own_w * MAX_COLORS * 4);
            off_adj_cnt = o; o = align(o + (size_t)nb * 4);
            off_adj_head = o; o = align(o + (size_t)nb * 4);
            off_cstate = o; o = align(o + (size_t)nb * 16);
            off_body_shared = o; o = align(o + (size_t)nb * 4);
            off_pend_cnt = o; o = align(o + (MAX_COLOR_ROUNDS + 2) * 4);
            zeroed_bytes = o;
            if ((st = zeroed.reserve(zeroed_bytes))) return st;
        }
        if ((st = own_pos.reserve((own_w + 1) * MAX_COLORS + 2)) ||
            (st = pose.reserve(nb)) || (st = view.reserve(4 * (size_t)nb)) || (st = ncells.reserve(nb)) || (st = bkt.reserve(nb)) ||  (st = bucket_cnt.reserve(2 * (size_t)T + 1, true, stream)) ||
            (st = bucket_start.reserve(2 * (size_t)T + 2)) || (st = fcell.reserve(nb)) || (st = fine_cand.reserve(2 * (size_t)nb)) || (st = pair_cnt.reserve((size_t)nb + 3)) || (st = ent_off.reserve((size_t)T + 2)) || (st = work.reserve(2 * (size_t)T + 2)))
            return st;
        const bool tiny = test_small_buffers;   // (tests) so that the grow-and-redo path runs
        if (cap_entries == 0 && (st = reserve_entries(tiny ? 64 : (size_t)nb * 3 + 4096))) return st;
        if (cap_pairs == 0 && (st = reserve_pairs(tiny ? 64 : (size_t)nb * 6 + 4096))) return st;

        // broadphase flavour: the fine grid whenever the upload found a usable fine cell; the bucket pair kernels then
        // only run for large-large pairs, i.e. when a dynamic large body exists.  Their pairs come first in the list, so
        // a batch (whose per-world kernels need a world's pairs contiguous) with dynamic large bodies keeps the buckets.
        fine_now = fine_grid && image.fine_cell > 0.0f && image.fine_for_cell == grid_cell() &&
                   (image.n_large_dynamic == 0 || worlds.size() == 1);
        ll_now = fine_now && image.n_large_dynamic > 0;
        // colouring flavour: CTA-per-world rounds in shared memory for batches of small worlds, else the dataflow
        // colouring while the candidate pairs fit its register slots (decided again on the device), else grid-wide rounds
        const bool many_small_worlds = world_solver && worlds.size() >= (size_t)n_sms / 2;
        const bool color_per_world = many_small_worlds && max_world_bodies <= COLOR_WORLD_MAX_BODIES;
        const size_t pairs_guess = last_pairs ? (size_t)last_pairs : (size_t)nb * 3;
        flow_now = flow_coloring && !color_per_world && pairs_guess <= (size_t)FLOW_BIG_SLOTS * color_blocks * TPB;
        if (flow_now && ((st = adj_prio.reserve((size_t)nb * ADJ_CAP)) || (st = adj_pool.reserve(cap_pairs / 2 + 1024)))) return st;
        // solver flavour: CTA-per-world for batches of small worlds; one spatial tile of bodies per SM for a single world
        // without joints that fits (k_solve_tiles may still decline on the device); else the persistent dataflow sweep
        if ((opt_warm_start || opt_sleeping) && (mode == R2D_MODE_REFERENCE_ORDER || !persistent_solver)) {
            g_cuda_error = "R2D_OPT_WARM_START / R2D_OPT_SLEEPING need the default solver (not REFERENCE_ORDER, not R2D_SOLVER=launches)";
            return R2D_ERR_BAD_STATE;
        }
        if (opt_warm_start) {   // table of the previous call's contacts (16 slots per body: load factor < 0.2 in a dense pile)
            if (worlds.size() >= WARM_MAX_WORLD) return R2D_ERR_INVALID_ARGUMENT;
            for (auto& w : worlds)
                if (w->current_body_id >= WARM_MAX_ID) return R2D_ERR_INVALID_ARGUMENT;
            uint32_t want = 1024;
            while (want < 16u * nb && want < (1u << 30)) want <<= 1;
            if (want > warm_slots) {   // (a grown table starts empty: one call without warm terms)
                for (int t = 0; t < 2; ++t) {
                    if ((st = warm_key[t].reserve(want)) || (st = warm_meta[t].reserve(want)) || (st = warm_val[t].reserve(want))) return st;
                    R2D_CUDA(cudaMemsetAsync(warm_key[t].p, 0xFF, (size_t)want * 8, stream));
                }
                warm_slots = want;
            }
            if (s_warm0.cap < s_hdr.cap && ((st = s_warm0.reserve(s_hdr.cap)) || (st = s_warm1.reserve(s_hdr.cap)))) return st;
        }
        if (opt_sleeping && (st = sleep_state.reserve(nb))) return st;
        // one CTA per world: batches with enough worlds to fill the GPU with 128-thread CTAs, or a FEW worlds (one world, a handful
        // of worlds per GPU of a sharded batch) each small enough for one CTA of 512 threads
        const bool single_small = world_single && worlds.size() < (size_t)n_sms / 2 && max_world_bodies <= WORLD_SINGLE_MAX_BODIES;
        // (worlds with joints: their CTA solves them between the contact colours, R2D_WORLD_JOINTS=0: the device-wide sweep)
        const bool world_has_joints = !image.j_hdr.empty();
        const bool use_world_solver = persistent_solver && world_solver && (!world_has_joints || world_joints) && !opt_warm_start &&
                                      ((max_world_bodies <= WORLD_MAX_BODIES && worlds.size() >= (size_t)n_sms / 2) || single_small);
        world_fused_now = use_world_solver;
        const uint32_t tile_b = (nb + (uint32_t)n_sms - 1) / (uint32_t)n_sms;
        const bool use_tile_solver = persistent_solver && tile_solver && !tile_declined && !use_world_solver && image.j_hdr.empty() &&
                                     !opt_warm_start && !opt_sleeping &&
                                     tile_b <= TILE_MAX_BODIES && nb >= (uint32_t)n_sms * 8;
        tile_bodies_now = use_tile_solver ? tile_b : 0u;
        // broadphase of a batch of small worlds: per-world CTAs with the grid in shared memory, when the fine grid applies
        // (no dynamic large bodies) and the tables of the largest world fit
        const uint32_t wb_nb_cap = (max_world_bodies + 3u) & ~3u, wb_tw_cap = grid_mult() * max_world_bodies;
        const uint32_t wb_ent_cap = wb_nb_cap + 256u;   // one entry per small body + the cells of the large ones
        const size_t wb_smem = world_broad_smem_bytes(wb_nb_cap, wb_tw_cap, wb_ent_cap);
        bool use_world_broad = world_broad && !world_broad_declined && fine_now && !ll_now && worlds.size() + 2 <= scan_state_cap &&
                               ((many_small_worlds && max_world_bodies <= WORLD_MAX_BODIES && wb_smem <= 100 * 1024) ||
                                (single_small && use_world_solver && wb_smem <= WORLD_SMEM_MAX));   // (one world: one CTA of 1,024 threads)
        if (opt_sleeping) {   // once per call, not per attempt: it edits the static flags
            fill_dev();
            if ((st = join_forces())) return st;   // user forces wake
            R2D_LAUNCH(R2D_KCLASS_INTEGRATE, k_sleep_begin, grid_for(nb), TPB, d, opt_sleep_calls);
        }
        for (int attempt = 0;; ++attempt) {
            fill_dev();
            world_exported = false;
            R2D_CUDA(cudaMemsetAsync(zeroed.p, 0, zeroed_bytes, stream));  // counters, colour tables, scan states, masks
            // ---- broadphase ----
            if (use_world_broad) {   // one CTA per world, the world's grid in shared memory (r2d_world.cuh)
                const uint32_t blocks = (uint32_t)std::min<size_t>(worlds.size(), (size_t)n_sms * 16);
                prof_begin(R2D_KCLASS_BROADPHASE);
                k_world_broad<<<blocks, single_small ? 1024 : WORLD_BROAD_TPB, wb_smem, stream>>>(d, wb_nb_cap, wb_tw_cap, wb_ent_cap, scan_state(0),
                                                                             (uint32_t*)(scan_state(0) + scan_state_cap - 1));
                prof_end();
                launches += 1;
            } else {
                R2D_LAUNCH(R2D_KCLASS_BROADPHASE, k_grid_cells<false>, grid_for(nb), TPB, d);
                if ((st = scan(d.bucket_cnt, d.bucket_start, nullptr, fine_now ? 2 * T : T, &d.counters->n_entries, R2D_KCLASS_BROADPHASE, 0))) return st;
                R2D_LAUNCH(R2D_KCLASS_BROADPHASE, k_grid_cells<true>, grid_for(nb), TPB, d);
                if (!fine_now || ll_now) {
                    R2D_LAUNCH(R2D_KCLASS_BROADPHASE, k_list_buckets, grid_for(T), TPB, d);
                    R2D_LAUNCH(R2D_KCLASS_BROADPHASE, k_sort_buckets, pair_blocks, TPB, d);
                    // pairs per BUCKET (one warp each) -> scan over the T buckets -> write at the scanned offsets
                    R2D_LAUNCH(R2D_KCLASS_BROADPHASE, k_bucket_count, pair_blocks, TPB, d);
                    if ((st = scan(d.ent_off, d.ent_off, nullptr, T, &d.counters->n_pairs, R2D_KCLASS_BROADPHASE, 1))) return st;
                    R2D_LAUNCH(R2D_KCLASS_BROADPHASE, k_bucket_write, pair_blocks, TPB, d);
                }
                if (fine_now) {
                    // pairs per small BODY (8 lanes each) -> scan over the bodies -> write at the scanned offsets
                    R2D_LAUNCH(R2D_KCLASS_BROADPHASE, k_fine_pairs<false>, grid_for(nb), TPB, d);
                    if ((st = scan(d.pair_cnt, d.pair_cnt, nullptr, nb + 1, &d.counters->n_pairs, R2D_KCLASS_BROADPHASE, 3))) return st;
                    R2D_LAUNCH(R2D_KCLASS_BROADPHASE, k_fine_pairs<true>, grid_for(nb), TPB, d);
                }
            }
            // ---- narrowphase ----
            // larger re-deal tiles sort the shape classes better; small pair counts keep one pair per thread (more CTAs)
            if (pairs_guess > (size_t)n_sms * 3 * TPB * 6)
                R2D_LAUNCH(R2D_KCLASS_NARROWPHASE, k_narrow<4>, grid_for((cap_pairs + 3) / 4), TPB, d);
            else if (pairs_guess > (size_t)n_sms * 3 * TPB * 3)
                R2D_LAUNCH(R2D_KCLASS_NARROWPHASE, k_narrow<2>, grid_for((cap_pairs + 1) / 2), TPB, d);
            else
                R2D_LAUNCH(R2D_KCLASS_NARROWPHASE, k_narrow<1>, grid_for(cap_pairs), TPB, d);
            if (mode == R2D_MODE_REFERENCE_ORDER) {   // validation: the reference's own sequential sweep order (see r2d_abi.h)
                R2D_CUDA(cudaMemcpyAsync(&pinned->counters, d.counters, sizeof(Counters), cudaMemcpyDeviceToHost, stream));
                R2D_CUDA(cudaStreamSynchronize(stream));
                const Counters c0 = pinned->counters;
                if (c0.n_entries > cap_entries || c0.n_pairs > cap_pairs) {
                    if (attempt >= 4) return R2D_ERR_OUT_OF_MEMORY;
                    if (c0.n_entries > cap_entries) {
                        if ((st = reserve_entries((size_t)c0.n_entries + c0.n_entries / 4))) return st;
                    } else if ((st = reserve_pairs((size_t)c0.n_pairs + c0.n_pairs / 4))) {
                        return st;
                    }
                    continue;
                }
                if ((st = join_forces())) return st;
                const uint32_t P = c0.n_pairs;
                std::vector<uint2> hp(P);
                std::vector<uint32_t> hc(P);
                std::vector<float4> ha(nb), ha_iter(nb);
                if (P) {
                    R2D_CUDA(cudaMemcpyAsync(hp.data(), pairs.p, (size_t)P * 8, cudaMemcpyDeviceToHost, stream));
                    R2D_CUDA(cudaMemcpyAsync(hc.data(), m_color.p, (size_t)P * 4, cudaMemcpyDeviceToHost, stream));
                }
                R2D_CUDA(cudaMemcpyAsync(ha.data(), aabb.p, (size_t)nb * 16, cudaMemcpyDeviceToHost, stream));
                R2D_CUDA(cudaStreamSynchronize(stream));
                for (uint32_t k = 0; k < nb; ++k) ha_iter[k] = ha[image.dev_of_host[k]];   // iteration order of the host registry
                std::unordered_map<uint64_t, uint32_t> slot_of_pair;
                for (uint32_t p = 0; p < P; ++p) {
                    if (hc[p] == COLOR_NONE) continue;   // SAT found a gap: no manifold
                    const uint32_t a = image.host_of_dev[hp[p].x], b = image.host_of_dev[hp[p].y];
                    slot_of_pair[((uint64_t)std::min(a, b) << 32) | std::max(a, b)] = p;
                }
                std::vector<uint32_t> pair_slots;   // manifold index -> pair slot, filled through the lookup below
                pair_slots.reserve(slot_of_pair.size());
                std::unordered_map<uint64_t, long> index_of_pair;
                for (auto& kv : slot_of_pair) {
                    index_of_pair[kv.first] = (long)pair_slots.size();
                    pair_slots.push_back(kv.second);
                }
                const std::vector<uint32_t> by_creation = host::reference_manifold_order(
                    ha_iter, grid_cell(), T, pair_slots.size(), [&](uint32_t lo, uint32_t hi) -> long {
                        auto it = index_of_pair.find(((uint64_t)lo << 32) | hi);
                        return it == index_of_pair.end() ? -1L : it->second;
                    });
                if (by_creation.size() != pair_slots.size()) {
                    g_cuda_error = "internal error: a manifold was not met by the replay of the reference's grid walk";
                    return R2D_ERR_CUDA;
                }
                std::vector<uint32_t> order(by_creation.size());
                for (size_t k = 0; k < order.size(); ++k) order[k] = pair_slots[by_creation[k]];
                const size_t nj = image.j_hdr.size();
                std::vector<uint32_t> joint_list(nj);   // list order: the device joints are sorted by colour
                for (size_t k = 0; k < nj; ++k) joint_list[image.joint_local_index[k]] = (uint32_t)k;
                if ((st = ref_order.reserve(order.size() + 1)) || (st = ref_joints.reserve(nj + 1))) return st;
                if (!order.empty()) R2D_CUDA(cudaMemcpyAsync(ref_order.p, order.data(), order.size() * 4, cudaMemcpyHostToDevice, stream));
                if (nj) R2D_CUDA(cudaMemcpyAsync(ref_joints.p, joint_list.data(), nj * 4, cudaMemcpyHostToDevice, stream));
                R2D_LAUNCH(R2D_KCLASS_SOLVE_CONTACTS, k_solve_reference_order, 1, TPB, d, sub_dt, S, I, (const uint32_t*)ref_order.p,
                           (uint32_t)order.size(), (const uint32_t*)ref_joints.p, (uint32_t)nj);
                R2D_CUDA(cudaMemcpyAsync(&pinned->counters, d.counters, sizeof(Counters), cudaMemcpyDeviceToHost, stream));
                R2D_CUDA(cudaStreamSynchronize(stream));   // (also keeps the host vectors alive until the copies are done)
                R2D_CUDA(cudaGetLastError());
                break;
            }
            // ---- colouring + partition + pre-step ----
            if (color_per_world) {
                const uint32_t blocks = (uint32_t)std::min<size_t>(worlds.size(), (size_t)n_sms * 8);
                if (seq_world_coloring) {
                    const uint32_t smem_bodies = max_world_bodies;
                    const uint32_t export_used = use_world_solver ? 0u : 1u;   // colour masks for the dataflow sweep
                    prof_begin(R2D_KCLASS_COLORING);
                    k_color_worlds_seq<<<blocks, WORLD_TPB, (size_t)smem_bodies * COLOR_WORDS * 8, stream>>>(d, smem_bodies, export_used);
                    prof_end();
                    launches += 1;
                } else {
                    R2D_LAUNCH(R2D_KCLASS_COLORING, k_color_worlds, blocks, WORLD_TPB, d);
                }
            } else {
                prof_begin(R2D_KCLASS_COLORING);
                uint32_t reg_slots = flow_list_only ? 0u : (uint32_t)FLOW_SLOTS;
                void* args[] = {(void*)&d, (void*)&reg_slots};
                R2D_CUDA(cudaLaunchCooperativeKernel((void*)k_color, dim3(color_blocks), dim3(TPB), args, 0, stream));
                prof_end();
                launches += 1;
            }
            // owner bitmaps -> popcounts -> scan = position of every manifold in the colour-sorted, spatially ordered records
            // (the owner bitmaps are set by the colouring kernels themselves, at the moment a manifold gets its colour);
            // k_world_solve places and pre-steps the manifolds of its world itself
            if (!use_world_solver) {   // popcounts of the owner bitmaps (and the number of colours) are computed inside the scan
                const uint32_t n_max = (uint32_t)((own_w + 1) * MAX_COLORS);
                const uint32_t tiles = (n_max + SCAN_TILE - 1) / SCAN_TILE + 1;
                if (tiles + 1 > scan_state_cap) return R2D_ERR_CUDA;
                R2D_LAUNCH(R2D_KCLASS_COLORING, k_scan_owners, tiles, SCAN_TPB, d, scan_state(2), (uint32_t)(scan_state_cap - 1));
                R2D_LAUNCH(R2D_KCLASS_COLORING, k_partition_prestep, grid_for(cap_pairs), TPB, d);
            }
            // ---- substeps: one persistent cooperative kernel (colour ranges are read on the device) ----
            if ((st = join_forces())) return st;  // forces written for this step have arrived (first reader of `frc`)
            if (use_world_solver) {
                uint32_t nb_cap = 0, R = 0;
                const size_t smem = world_cache(nb_cap, R);
                const uint32_t blocks = (uint32_t)std::min<size_t>(worlds.size(), (size_t)n_sms * 16);
                // r2d_process_read into page-locked arrays: every CTA exports its world as soon as it is done (WorldExport)
                WorldExport ex = {nullptr, nullptr, nullptr, nullptr, host_of_dev.p};
                if (readback && world_export && S > 0 && !readback->ids && !readback->aabb_xywh && readback->pos_xy && readback->angle &&
                    readback->momentum_xy && readback->ang_momentum) {
                    void* m[4] = {mapped_host(readback->pos_xy), mapped_host(readback->angle), mapped_host(readback->momentum_xy),
                                  mapped_host(readback->ang_momentum)};
                    if (m[0] && m[1] && m[2] && m[3]) {
                        ex.pos = (float2*)m[0];
                        ex.angle = (float*)m[1];
                        ex.mom = (float2*)m[2];
                        ex.ang_mom = (float*)m[3];
                    }
                }
                prof_begin(R2D_KCLASS_SOLVE_CONTACTS);
                const bool exporting = ex.pos != nullptr;
#define R2D_WORLD_SOLVE_J(BPT, WTPB, GRID, JOINTS)                                                                      \
    do {                                                                                                              \
        if (exporting)                                                                                                \
            k_world_solve<BPT, WTPB, true, JOINTS><<<GRID, WTPB, smem, stream>>>(d, sub_dt, S, I, nb_cap, R, ex);     \
        else                                                                                                          \
            k_world_solve<BPT, WTPB, false, JOINTS><<<GRID, WTPB, smem, stream>>>(d, sub_dt, S, I, nb_cap, R, ex);    \
    } while (0)
#define R2D_WORLD_SOLVE(BPT, WTPB, GRID)                                                                              \
    do {                                                                                                              \
        if (world_has_joints)                                                                                         \
            R2D_WORLD_SOLVE_J(BPT, WTPB, GRID, true);                                                                 \
        else                                                                                                          \
            R2D_WORLD_SOLVE_J(BPT, WTPB, GRID, false);                                                                \
    } while (0)
                if (single_small)
                    R2D_WORLD_SOLVE(2, WORLD_SINGLE_TPB, blocks);
                else if (max_world_bodies <= 2u * WORLD_SOLVE_TPB)
                    R2D_WORLD_SOLVE(2, WORLD_SOLVE_TPB, blocks);
                else
                    R2D_WORLD_SOLVE(4, WORLD_SOLVE_TPB, blocks);
#undef R2D_WORLD_SOLVE_J
#undef R2D_WORLD_SOLVE
                prof_end();
                launches += 1;
                world_exported = ex.pos != nullptr;
            } else if (use_tile_solver) {
                prof_begin(R2D_KCLASS_SOLVE_CONTACTS);
                uint32_t S_ = S, I_ = I, cache = TILE_CACHE_TASKS, max_tasks = tile_max_tasks;
                float sd = sub_dt;
                void* args[] = {(void*)&d, (void*)&sd, (void*)&S_, (void*)&I_, (void*)&cache, (void*)&max_tasks};
                R2D_CUDA(cudaLaunchCooperativeKernel((void*)k_solve_tiles, dim3(n_sms), dim3(TILE_TPB), args, TILE_SMEM_BYTES, stream));
                prof_end();
                launches += 1;
            } else if (persistent_solver) {
                if ((st = launch_persistent(sub_dt, S, I))) return st;
            }
            if (opt_warm_start) {   // (an abandoned attempt stores nothing — the kernel checks — and the tables do not swap)
                warm_saving = true;
                fill_dev();
                R2D_CUDA(cudaMemsetAsync(d.warm_key, 0xFF, (size_t)warm_slots * 8, stream));
                R2D_LAUNCH(R2D_KCLASS_COLORING, k_warm_save, grid_for(cap_pairs), TPB, d, (float)S);
                warm_saving = false;
                fill_dev();
            }
            // r2d_process_read: the export of the new state rides behind the solver, inside the same synchronisation
            if (readback && world_exported) {
                readback_done = true;   // (k_world_solve wrote the caller's arrays; the synchronisation below completes them)
            } else if (readback && persistent_solver) {
                const uint32_t keep = launches;
                if ((st = enqueue_read_bodies(0, nb, readback->ids, readback->pos_xy, readback->angle, readback->momentum_xy,
                                              readback->ang_momentum, readback->aabb_xywh)))
                    return st;
                launches = keep + 1;
                fill_dev();
                readback_done = true;
            }
            // ---- the one synchronisation point of the step: counters + colour offsets ----
            R2D_CUDA(cudaMemcpyAsync(&pinned->counters, d.counters, sizeof(Counters), cudaMemcpyDeviceToHost, stream));
            R2D_CUDA(cudaMemcpyAsync(pinned->color_start, d.color_start, (MAX_COLORS + 1) * 4, cudaMemcpyDeviceToHost, stream));
            R2D_CUDA(cudaStreamSynchronize(stream));
            R2D_CUDA(cudaGetLastError());
            const Counters& c = pinned->counters;
            if (c.broad_fallback && use_world_broad) {   // a world's grid did not fit shared memory: nothing was modified
                if (attempt >= 4) return R2D_ERR_CUDA;
                world_broad_declined = true;
                use_world_broad = false;
                continue;
            }
            if (c.n_entries > cap_entries || c.n_pairs > cap_pairs) {
                if (attempt >= 4) {
                    g_cuda_error = "broadphase buffers failed to converge";
                    return R2D_ERR_OUT_OF_MEMORY;
                }
                // P is only meaningful once all entries fit (the pair kernels do nothing on an attempt whose grid overflowed,
                // so n_pairs is then the scan of stale counts)
                if (c.n_entries > cap_entries) {
                    if ((st = reserve_entries((size_t)c.n_entries + c.n_entries / 4))) return st;
                } else if ((st = reserve_pairs((size_t)c.n_pairs + c.n_pairs / 4))) {
                    return st;
                }
                continue;
            }
            if (c.tile_fallback && use_tile_solver) {
                // k_solve_tiles declined before touching anything (a tile with too many bodies or records): run the
                // persistent sweep on the same records now, and stop trying tiles until the next upload
                tile_declined = true;
                if ((st = launch_persistent(sub_dt, S, I))) return st;
                if (readback && (st = enqueue_read_bodies(0, nb, readback->ids, readback->pos_xy, readback->angle, readback->momentum_xy,
                                                          readback->ang_momentum, readback->aabb_xywh)))
                    return st;   // the export enqueued above saw the state before this sweep
                R2D_CUDA(cudaMemcpyAsync(&pinned->counters, d.counters, sizeof(Counters), cudaMemcpyDeviceToHost, stream));
                R2D_CUDA(cudaStreamSynchronize(stream));
                R2D_CUDA(cudaGetLastError());
            }
            break;
        }
        if (opt_sleeping) {   // sleepers get their flag back, speeds are measured, movers wake what they touched
            fill_dev();
            R2D_LAUNCH(R2D_KCLASS_INTEGRATE, k_sleep_bodies, grid_for(nb), TPB, d);
            R2D_LAUNCH(R2D_KCLASS_INTEGRATE, k_sleep_wake, grid_for(cap_pairs), TPB, d);
            R2D_CUDA(cudaStreamSynchronize(stream));
        }
        const Counters c = pinned->counters;
        last_pairs = c.n_pairs;
        seen_world_m = c.max_world_m;
        stats.n_buckets = T;
        stats.n_entries = c.n_entries;
        stats.n_pairs = c.n_pairs;
        stats.n_manifolds = c.n_manifolds;
        last_manifolds = c.n_manifolds;
        stats.n_points = c.n_points;
        stats.n_colors = c.n_colors;
        stats.n_color_rounds = c.n_rounds;
        if (c.err & ERR_GRID_RANGE) {
            g_cuda_error = "a body AABB covers an unreasonable number of grid cells (NaN/inf pose?)";
            return R2D_ERR_GRID_RANGE;
        }
        if (c.err & ERR_FINE) {
            g_cuda_error = "internal error: a body classified as small covers more than 4 coarse cells";
            return R2D_ERR_CUDA;
        }
        if (c.err & ERR_FLOW_STALL) {
            g_cuda_error = "internal error: the dataflow colouring stalled (state of this step is undefined)";
            return R2D_ERR_CUDA;
        }
        if (c.err & ERR_STALL) {
            g_cuda_error = "internal error: the dataflow contact sweep stalled (state of this step is undefined)";
            return R2D_ERR_CUDA;
        }
        stats.n_dropped = c.n_dropped;
        if (c.err & ERR_ROUNDS) {
            g_cuda_error = "the colouring did not finish within its round limit";
            return R2D_ERR_COLOR_OVERFLOW;
        }
        // ---- substeps, one launch per colour (A/B path: R2D_SOLVER=launches) ----
        const uint32_t* cs = pinned->color_start;
        const auto& jcs = image.joint_color_start;
        for (uint32_t s = 0; s < S && !persistent_solver; ++s) {
            R2D_LAUNCH(R2D_KCLASS_INTEGRATE, k_integrate_forces, grid_for(nb), TPB, d, sub_dt, (int)(s + 1 == S));
            for (uint32_t it = 0; it < I; ++it) {
                for (size_t jc = 0; jc + 1 < jcs.size(); ++jc) {
                    const uint32_t n = jcs[jc + 1] - jcs[jc];
                    if (n) R2D_LAUNCH(R2D_KCLASS_SOLVE_JOINTS, k_solve_joints, (n + SOLVE_TPB - 1) / SOLVE_TPB, SOLVE_TPB, d, jcs[jc], jcs[jc + 1], sub_dt);
                }
                for (uint32_t col = 0; col < c.n_colors; ++col) {
                    const uint32_t n = cs[col + 1] - cs[col];
                    if (n) R2D_LAUNCH(R2D_KCLASS_SOLVE_CONTACTS, k_solve_contacts, (n + SOLVE_TPB - 1) / SOLVE_TPB, SOLVE_TPB, d, cs[col], cs[col + 1], sub_dt);
                }
            }
            R2D_LAUNCH(R2D_KCLASS_INTEGRATE, k_integrate_positions, grid_for(nb), TPB, d, sub_dt);
        }
        R2D_CUDA(cudaGetLastError());
        stats.n_launches = launches;
        if (opt_warm_start) warm_cur ^= 1;   // the table this call filled is the next call's input
        return R2D_OK;
    }
};

}  // namespace

static r2d::host::BatchBase* r2d_new_backend(int device, std::string& err) {
    int n = 0;
    const cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        err = std::string("no CUDA device available (") + cudaGetErrorString(e) + "): libr2d_b200 has no CPU fallback";
        return nullptr;
    }
    if (device < 0 || device >= n) {
        err = "device index out of range";
        return nullptr;
    }
    CudaBatch* b = new CudaBatch();
    if (b->init(device) != R2D_OK) {
        err = g_cuda_error;
        delete b;
        return nullptr;
    }
    return b;
}
static int r2d_backend_device_count() {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

#define R2D_API(name) r2d_##name
#define R2D_BACKEND_ERROR g_cuda_error
#include "r2d_capi.inc"
