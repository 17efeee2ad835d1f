// Host engine + C-ABI (include/iadhore_b200.h) of the B200 collinearity engine.
//
// Replaces the scheduling half of DataSet::parallelLevel2ADHoReDyn (levelTwo.cpp:177-292) and
// DataSet::parallelProfileSearch (higherLevel.cpp:252-326): instead of a pthread pool pulling
// weight-balanced packets, every GHM of a call is batched into three device stages
//   K1 build (count -> per-GHM tile scan -> fill)      build_kernels.cu
//   K2 cluster (persistent CTAs, heaviest GHM first)   cluster_kernel.cu
//   K3 pack (measure -> host size scan -> write)       pack_kernel.cu
// with persistent, growable device buffers and pinned staging.  No CPU fallback exists: every
// entry point fails if CUDA is unavailable.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/iadhore_b200.h"
#include "cluster_math.cuh"
#include "common.h"
#include "workspace.h"
#include "homolog.h"

// kernels (other translation units)
void launch_build_count(bool hl, const DatasetDev& ds, const ProfileDev& pf, GhmTask* tasks, int n_ghm,
                        const int64_t* tile_prefix, int64_t n_tiles, uint32_t* tile_cnt, int sm_count, cudaStream_t st);
void launch_tile_scan(const int64_t* tile_prefix, int n_ghm, const uint32_t* tile_cnt, uint32_t* tile_off,
                      cudaStream_t st);
void launch_build_fill(bool hl, const DatasetDev& ds, const ProfileDev& pf, GhmTask* tasks, int n_ghm,
                       const int64_t* tile_prefix, int64_t n_tiles, const uint32_t* tile_off, uint32_t* keys,
                       int sm_count, cudaStream_t st);
void launch_mask(uint8_t* flags, int64_t begin, int64_t end, int set, cudaStream_t st);
void launch_expand_count(const DatasetDev& ds, int64_t n_elems, const int32_t* pair_idx, const uint8_t* row_used, uint32_t* cnt,
                         int sm_count, cudaStream_t st);
void launch_expand_fill(const DatasetDev& ds, int64_t n_elems, const int32_t* pair_idx, const uint8_t* row_used, uint32_t* cursor,
                        const int64_t* dot_off, const uint32_t* cnt, uint32_t* keys, GhmTask* tasks, int n_ghm, int sm_count,
                        cudaStream_t st);
void launch_hl_expand_count(const DatasetDev& ds, const HlGroup* items, int n_items, const int32_t* task_idx, const int16_t* vis,
                            uint32_t* cnt, int sm_count, cudaStream_t st);
void launch_hl_expand_fill(const DatasetDev& ds, const HlGroup* items, int n_items, const int32_t* task_idx, const int16_t* vis,
                           uint32_t* cursor, const int64_t* dot_off, const uint32_t* cnt, uint32_t* keys, GhmTask* tasks, int n_ghm,
                           int sm_count, cudaStream_t st);
int expand_sort_max(void);
void launch_tandem_pairs(const int64_t* list_off, int n_lists, const int32_t* elem_list, const int32_t* gid, const uint8_t* flags,
                         const int64_t* pair_ptr, const int32_t* pair_ids, int n_genes, int gap, int64_t total, uint8_t* remapped,
                         int32_t* t3, cudaStream_t st);
void launch_tables(const int32_t* fam, const int32_t* gid, uint8_t* flags, const int64_t* list_off, int n_lists, int64_t n, int n_fam,
                   int32_t* fam_cnt, int32_t* gid_cnt, int32_t* cursor, int32_t* fam_ptr, int32_t* members, int32_t* elem_list,
                   uint32_t* mem_y, int32_t* mem_list, int32_t* self_slot, int32_t* plus_prefix, int32_t* tmp, int32_t* block_sum,
                   int sm_count, cudaStream_t st);
void launch_intern(const char* blob, const int64_t* off, int64_t n, int32_t* table, uint32_t table_size, int32_t* rep, int32_t* cmin,
                   int32_t* flag_scan, int32_t* block_sum, int32_t* total, int32_t* ids, cudaStream_t st);
void launch_fam_list(const int32_t* fam_ptr, const int32_t* members, const int64_t* list_off, int n_fam, int n_lists,
                     int64_t total, int32_t* table, cudaStream_t st);
void launch_tandem(const int64_t* list_off, int n_lists, const int32_t* elem_list, const int32_t* fam, const int32_t* gid,
                   const uint8_t* flags, int gap, int64_t total, uint8_t* remapped, int32_t* t3, cudaStream_t st);
void launch_cluster(GhmTask* tasks, GhmProf* profs, const int* order, int n_tasks, int* next_task, const uint32_t* dots, uint8_t* ws,
                    const DevParams& prm, int pool_factor, int* sm_slot, const DatasetDev& ds, const ProfileDev& pf, bool hl,
                    int grid, int nt, cudaStream_t st);
void launch_phase1_pass(GhmTask* tasks, const int* items, int n_items, int* next_item, const uint32_t* dots, uint8_t* ws,
                        const DevParams& prm, int pool_factor, int* sm_slot, int pass, int gi, int grid, int nt,
                        cudaStream_t st, unsigned long long* dbg, uint32_t* istate);
void launch_phase2_stage(GhmTask* tasks, const int* items, int n_items, int* next_item, const uint32_t* dots, uint8_t* ws,
                         const DevParams& prm, int pool_factor, int* sm_slot, const DatasetDev& ds, const ProfileDev& pf,
                         bool hl, int stage, int grid, int nt, cudaStream_t st, const uint32_t* istate);
cudaError_t launch_resident(GhmTask* tasks, const int* order, int n_tasks, int* next_task, const uint32_t* dots, uint8_t* ws,
                            const DevParams& prm, int pool_factor, int smem_bytes, const DatasetDev& ds, const ProfileDev& pf, bool hl,
                            int grid, int nt, cudaStream_t st, unsigned long long* dbg = nullptr);
int resident_ctas_per_sm(int nt, int smem_bytes);
bool pass_cfg_supported(int code);
void launch_link_tables(const LinkIn* links, const int64_t* link_ptr, int64_t n_links, const int32_t* elem_base, int n_prof, int64_t n_elems,
                        const ProfileDev& pf, unsigned long long* keys, int32_t* vals, uint32_t mask, int32_t* cols, int32_t* cnt,
                        int32_t* lk_ptr, int32_t* lk_col, int32_t* block_sum, cudaStream_t st);
void launch_homolog_filter(GhmTask* tasks, int n_tasks, uint8_t* ws, const DatasetDev& ds, int n_fam, const ProfileDev& pf,
                           const HomologDev* hdev, int pool_factor, int min_homologs, int sm_count, cudaStream_t st);
void launch_pack_write(bool hl, const GhmTask* tasks, const int* sel, const int64_t* out_off, int n_sel, uint8_t* ws,
                       const DatasetDev& ds, const ProfileDev& pf, int pool_factor, uint8_t* out, cudaStream_t st);

static thread_local std::string g_err;
static int fail(int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
#define CU(call)                                                                                       \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess)                                                                         \
            return fail(IADH_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// IADH_TRACE=1: print the wall time of every stage (synchronising after each; debugging aid)
#include <chrono>
static bool g_trace = getenv("IADH_TRACE") != nullptr;
static double g_t_last = 0;
static void trace_mark(cudaStream_t st, const char* label)
{
    if (!g_trace) return;
    const cudaError_t se = st ? cudaStreamSynchronize(st) : cudaSuccess;   // (null: host-only mark, the device keeps running)
    if (se != cudaSuccess) fprintf(stderr, "[iadh] !! %s after stage '%s'\n", cudaGetErrorString(se), label ? label : "(start)");
    const double now = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
    if (label) fprintf(stderr, "[iadh] %-28s %8.3f ms\n", label, (now - g_t_last) * 1e3);
    g_t_last = now;
}

// Process-wide block cache behind the growable buffers: contexts are cheap to create and destroy
// (a drop-in caller may build one per run), cudaMalloc / cudaMallocHost / cudaFree are not (and
// cudaFree synchronises the device).  Blocks return here instead of to the driver; iadh_trim()
// gives them back.
#include <mutex>
struct BlockCache {
    struct Blk {
        void* p;
        size_t cap;
        int dev;  // -1 = pinned host
    };
    std::mutex mu;
    std::vector<Blk> blocks;
    void* take(size_t bytes, int dev, size_t* cap_out)
    {
        std::lock_guard<std::mutex> lk(mu);
        int best = -1;
        for (int i = 0; i < (int)blocks.size(); i++)
            if (blocks[i].dev == dev && blocks[i].cap >= bytes && (best < 0 || blocks[i].cap < blocks[best].cap)) best = i;
        if (best < 0) return nullptr;
        void* p = blocks[best].p;
        *cap_out = blocks[best].cap;
        (dev < 0 ? held_host : held_dev) -= blocks[best].cap;
        blocks.erase(blocks.begin() + best);
        return p;
    }
    // The cache is bounded (IADH_CACHE_MAX_MB, default 64 GB of device blocks and a quarter of that in pinned
    // host blocks): a block that would push it over the bound goes back to the driver at once.
    size_t held_dev = 0, held_host = 0;
    static size_t limit_dev()
    {
        static const size_t v = getenv("IADH_CACHE_MAX_MB") ? (size_t)std::max(0, atoi(getenv("IADH_CACHE_MAX_MB"))) << 20 : (size_t)64 << 30;
        return v;
    }
    void give(void* p, size_t cap, int dev)
    {
        std::lock_guard<std::mutex> lk(mu);
        size_t& held = dev < 0 ? held_host : held_dev;
        const size_t limit = dev < 0 ? limit_dev() / 4 : limit_dev();
        if (held + cap > limit) {
            if (dev < 0) cudaFreeHost(p);
            else {
                int cur = 0;
                cudaGetDevice(&cur);
                cudaSetDevice(dev);
                cudaFree(p);
                cudaSetDevice(cur);
            }
            return;
        }
        held += cap;
        blocks.push_back({p, cap, dev});
    }
    void trim()
    {
        std::lock_guard<std::mutex> lk(mu);
        int cur = 0;
        cudaGetDevice(&cur);
        for (Blk& b : blocks) {
            if (b.dev < 0) cudaFreeHost(b.p);
            else {
                cudaSetDevice(b.dev);
                cudaFree(b.p);
            }
        }
        blocks.clear();
        held_dev = held_host = 0;
        cudaSetDevice(cur);
    }
};
static BlockCache g_cache;

// IADH_EFENCE=1 (debugging aid; compute-sanitizer is not available on every pool): device buffers
// come from the virtual-memory API, placed so that the buffer ENDS at the last mapped byte (up to the
// 256-byte alignment of its start); the next page is reserved but unmapped, so any access past the
// end of a buffer faults deterministically instead of silently touching a neighbour.  Blocks are
// never recycled in this mode.
#include <dlfcn.h>
static void* efence_alloc(size_t bytes)
{
    // the driver API is looked up at run time: the library must load on machines without a driver
    static void* drv = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (!drv) return nullptr;
    typedef CUresult (*gran_t)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags);
    typedef CUresult (*reserve_t)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long);
    typedef CUresult (*create_t)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long);
    typedef CUresult (*map_t)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long);
    typedef CUresult (*access_t)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t);
    static gran_t f_gran = (gran_t)dlsym(drv, "cuMemGetAllocationGranularity");
    static reserve_t f_reserve = (reserve_t)dlsym(drv, "cuMemAddressReserve");
    static create_t f_create = (create_t)dlsym(drv, "cuMemCreate");
    static map_t f_map = (map_t)dlsym(drv, "cuMemMap");
    static access_t f_access = (access_t)dlsym(drv, "cuMemSetAccess");
    if (!f_gran || !f_reserve || !f_create || !f_map || !f_access) return nullptr;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaFree(0);
    CUmemAllocationProp prop = {};
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = dev;
    size_t gran = 0;
    if (f_gran(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM) != CUDA_SUCCESS) return nullptr;
    const size_t need = (bytes + 255) & ~(size_t)255;
    const size_t mapped = (need + gran - 1) / gran * gran;
    CUdeviceptr va = 0;
    if (f_reserve(&va, mapped + gran, gran, 0, 0) != CUDA_SUCCESS) return nullptr;
    CUmemGenericAllocationHandle h;
    if (f_create(&h, mapped, &prop, 0) != CUDA_SUCCESS) return nullptr;
    // IADH_EFENCE=2: guard page BEFORE the buffer instead (catches negative indices)
    static const bool before = getenv("IADH_EFENCE") != nullptr && atoi(getenv("IADH_EFENCE")) == 2;
    const CUdeviceptr at = before ? va + gran : va;
    if (f_map(at, mapped, 0, h, 0) != CUDA_SUCCESS) return nullptr;
    CUmemAccessDesc acc = {};
    acc.location = prop.location;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    if (f_access(at, mapped, &acc, 1) != CUDA_SUCCESS) return nullptr;
    return (void*)(before ? at : va + (mapped - need));
}

// growable device buffer
struct DBuf {
    void* p = nullptr;
    size_t cap = 0;
    int dev = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) {
            // the old block may still be read by work in flight; it must be idle before another
            // context can take it from the cache (growth is rare: sizes converge after the first call)
            cudaDeviceSynchronize();
            if (!getenv("IADH_EFENCE")) g_cache.give(p, cap, dev);
        }
        p = nullptr;
        cap = 0;
        cudaGetDevice(&dev);
        cudaError_t e = cudaSuccess;
        static const char* efence = getenv("IADH_EFENCE");
        if (efence) {
            p = efence_alloc(bytes);
            cap = p ? bytes : 0;
            return p ? cudaSuccess : cudaErrorMemoryAllocation;
        }
        if (!(p = g_cache.take(bytes, dev, &cap))) {
            size_t want = bytes + bytes / 4 + 256;
            static const bool trace_alloc = getenv("IADH_TRACE_ALLOC") != nullptr;
            if (trace_alloc) fprintf(stderr, "[iadh] cache miss: cudaMalloc(%zu MB) for a request of %zu MB\n", want >> 20, bytes >> 20);
            // never change the device's address space under a running kernel: with peer access enabled
            // (NCCL in the same process) cudaMalloc also edits the peers' page tables
            cudaDeviceSynchronize();
            e = cudaMalloc(&p, want);
            if (e != cudaSuccess) {
                cudaGetLastError();
                g_cache.trim();
                want = bytes;
                e = cudaMalloc(&p, want);
            }
            if (e == cudaSuccess) cap = want;
            else p = nullptr;
        }
        // IADH_POISON=<byte>: debugging aid, every block handed out is filled with that byte first, so a
        // read of memory the kernels did not write shows up deterministically (fresh cudaMalloc pages
        // are zero, recycled ones hold old data)
        static const char* poison = getenv("IADH_POISON");
        if (p && poison) {
            cudaMemset(p, atoi(poison), cap);
            cudaDeviceSynchronize();  // (the memset runs on the legacy stream, our streams are non-blocking)
        }
        return e;
    }
    void release()
    {
        if (p && !getenv("IADH_EFENCE")) g_cache.give(p, cap, dev);
        p = nullptr;
        cap = 0;
    }
    template <typename T>
    T* as() { return (T*)p; }
};
// growable pinned host buffer
struct HBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        release();
        if ((p = g_cache.take(bytes, -1, &cap))) return cudaSuccess;
        size_t want = bytes + bytes / 4 + 256;
        static const bool trace_alloc = getenv("IADH_TRACE_ALLOC") != nullptr;
        if (trace_alloc) fprintf(stderr, "[iadh] cache miss: cudaMallocHost(%zu MB) for a request of %zu MB\n", want >> 20, bytes >> 20);
        cudaError_t e = cudaMallocHost(&p, want);
        if (e == cudaSuccess) cap = want;
        else p = nullptr;
        return e;
    }
    void release()
    {
        if (p) g_cache.give(p, cap, -1);
        p = nullptr;
        cap = 0;
    }
    template <typename T>
    T* as() { return (T*)p; }
};

// One pinned block holding the streams of several results (the profiles of a batch, the results of a gather): each
// result is a slice of it -- no host copy per result -- and the block goes back to the cache with the last of them.
struct SharedPinned {
    HBuf buf;
    int refs = 0;
    cudaEvent_t ready = nullptr;   // the copy into buf on the context's copy stream; null = complete
    // Or a view of a host shared-memory segment into which every rank of the box copies its own part of a gather
    // (single-node exchange, see shm_* below): complete when every rank's `done` tag of the slot equals `want`.
    uint8_t* ext = nullptr;
    struct CommState* comm = nullptr;
    int slot = -1;
    uint32_t want = 0;
    bool failed = false;
    uint8_t* base() { return ext ? ext : buf.as<uint8_t>(); }
    void wait_own()
    {
        if (ready) {
            cudaEventSynchronize(ready);
            cudaEventDestroy(ready);
            ready = nullptr;
        }
    }
    void wait();      // (own copy, then the other ranks' parts; defined with the exchange)
    void release();   // last reference gone: the pinned block returns to the cache / the shm slot is released
};

struct iadh_result {
    HBuf bytes;            // pinned: the device writes the packed stream straight into it (offset 4)
    std::vector<uint8_t> owned;   // (host-assembled streams: the bisected level-2 call)
    SharedPinned* shared = nullptr;   // or a slice [shared_off, shared_off + n_bytes) of a shared block
    size_t shared_off = 0;
    iadh_ctx* owner = nullptr;    // set while the context may still hold this result's stream on the device
    // The device-to-host copy of the stream runs on the context's copy stream and overlaps whatever the caller does
    // next (the next call's kernels included); iadh_result_bytes waits for it.  null = nothing in flight.
    mutable cudaEvent_t ready = nullptr;
    size_t n_bytes = 0;
    std::vector<int32_t> ghm_index;
    iadh_stats stats;
};

struct iadh_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t st = nullptr;
    cudaStream_t st_side[3] = {nullptr, nullptr, nullptr};  // K2 size classes run concurrently
    cudaEvent_t ev_fork = nullptr, ev_join[3] = {nullptr, nullptr, nullptr};
    // K2 size classes, heaviest first: GHMs with >= min_dots dots run on nt-thread CTAs, per_sm CTAs per SM
    struct SizeClass {
        int min_dots, nt, per_sm;
    } cls[4] = {{257, 256, 2}, {0, 64, 4}, {0, 32, 16}, {0, 32, 16}};  // (measured on config 3 / 4 profile searches and config 3 level 2)
    int n_classes = 2;
    int split_min = 256;  // both orientations need this many dots to be clustered by separate CTAs
    // batches of at least pass_min GHMs run phase 1 one pass per launch (instruction-cache friendly, many CTAs per SM)
    int pass_min = 4096, pass_nt = 3232, pass_per_sm = 32;
    bool pass_cfg_user = false;   // IADH_PASS_CFG given: no adaptation to the batch size
    // K1 family-expansion path (build_kernels.cu, third generation): level-2 batches of at least this many GHMs
    int expand_min = 1024;
    int hl_expand_min = 16;            // the same for profile GHMs (family mode): calls of at least this many GHMProfiles
    std::vector<HlGroup> hl_items;     // column groups of the running call's profiles
    int expand_sort_max = 0;   // 0 = the sort kernel's capacity (test knob: smaller values exercise the fall-back)
    std::vector<int32_t> h_pair_idx;   // (x list, y list) -> task of the running call
    std::vector<uint8_t> h_row_used;
    // a batch whose clustering workspace would exceed this is bisected (level-2 calls): set from the
    // device's memory at create time, IADH_WS_BUDGET_MB overrides
    int64_t ws_budget = (int64_t)1 << 62;
    int64_t pass_ratio_hl = 1500;   // the same threshold for profile GHMs (IADH_PASS_RATIO_HL)
    int64_t pass_ratio = 1500;  // (config 3: 1,276 -> fused, tail-bound either way; 1/8 of config 4: 3,680 -> passes)
    // workspace bytes per chunk.  Default: one chunk.  Measured on config 4: chunks small enough for L2
    // (256 MB of workspace) serialise 52 x 34 short launches and take 241 ms instead of 87 ms.
    int64_t pass_chunk_bytes = (int64_t)1 << 60;
    // resident K2 (the GHM's clustering state in shared memory, one launch per size class): GHMs whose
    // RsLayout fits res_max_bytes; the classes are shared-memory budgets per CTA
    // shared-memory resident K2: 0 off, 1 on, 2 (default) = for large batches of profile searches only (>= pass_min active
    // GHMs) -- measured on config 4: 32 searches 25.6 -> 19.9 ms, 8 searches 11.9 -> 12.6 ms; the level-2 call and single
    // searches are faster on the global-memory paths
    int resident = 2;
    int res_max_bytes = 200 * 1024;
    int res_wide_min = 600, res_wide_nt = 128;   // latency-bound calls: GHMs with >= res_wide_min dots get wide CTAs
    cudaStream_t st_res[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_res[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    DBuf d_items_res, d_items_fb, d_resdbg, d_istate;
    DBuf d_passctr;
    DBuf d_items1;
    cudaEvent_t user_ev[4] = {nullptr, nullptr, nullptr, nullptr};  // iadh_event_record slots
    cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_call = nullptr;   // end of a call (ev[7] = its start)
    // result copies: their own stream, so that a packed stream travels to the host while the next call computes;
    // ev_packed = the writer finished (copy may start), ev_copied = the last copy finished (d_out / the gather's
    // receive buffer may be overwritten)
    cudaStream_t st_copy = nullptr;
    cudaEvent_t ev_packed = nullptr, ev_copied = nullptr, ev_copied_recv = nullptr;
    bool copy_pending = false;        // a copy out of d_out may be in flight (ev_copied)
    bool copy_pending_recv = false;   // a copy out of the gather's receive buffer may be in flight (ev_copied_recv)
    // dataset (host mirror + device)
    int n_lists = 0;
    std::vector<int64_t> list_off;
    std::vector<uint8_t> flags;       // host mirror incl. mask bits
    std::vector<int32_t> masked_cnt;  // per list
    std::vector<int32_t> index_to_list;
    int n_fam = 0;
    bool pairs = false;   // table_type=pairs (iadh_create_pairs): families are gene ids, buckets hold partners
    DBuf d_list_off, d_fam, d_gid, d_num_id, d_flags, d_fam_ptr, d_fam_members, d_mem_y, d_plus_prefix, d_self_slot, d_fam_list;
    DBuf d_mem_list, d_elem_list, d_pair_idx, d_row_used, d_cur, d_hl_items, d_ghdr;
    DatasetDev ds;
    DevParams prm;
    // per-call buffers
    DBuf d_tasks, d_tile_prefix, d_tile_cnt, d_tile_off, d_dots, d_ws, d_items, d_prof, d_plan, d_cnt, d_outs, d_sel, d_out_off, d_out, d_counter;
    DBuf d_pf_col;  // the profile descriptor (column groups + raw arrays), one block
    HBuf h_tasks, h_misc, h_pf, h_init;
    DBuf d_init;
    // multi-GPU (iadh_comm_*): NCCL communicator, loaded at run time
    struct CommState* comm = nullptr;
    // the packed stream of the most recent single-stream call still sits in d_out (iadh_allgather_results sends
    // it from there instead of re-uploading the host copy)
    const iadh_result* last_result = nullptr;
    size_t last_result_body = 0;
    // same for a batch (iadh_profile_search_batch): its results' streams lie back to back in d_out
    std::vector<const iadh_result*> last_batch;
    size_t last_batch_body = 0;
    // iadh_profile_search_batch: host copy of the profile lengths of the running batch, device staging
    std::vector<int32_t> batch_L;
    DBuf d_vis;
    HBuf h_vis;
    // homolog-count filter of the running profile search (iadh_profile_search_filtered, homolog_kernel.cu): per-profile
    // CSR "element -> columns of its inherited Link partners" on the device; hf_active is set by the _filtered entry
    // points around run_batch
    bool hf_active = false;
    int64_t hf_dropped = 0;   // multiplicons the filter removed in the last run_batch
    DBuf d_links, d_link_ptr, d_elem_base, d_lk_keys, d_lk_vals, d_lk_cols, d_lk_cnt, d_lk_ptr, d_lk_col, d_lk_bsum, d_hdev;
    // build_dots scratch
    std::vector<int64_t> bd_cnt;
    std::vector<int32_t> bd_dots;
    int pool_factor = 8;
    int ctas_per_sm = 4;
};

// the one result whose packed stream is still in d_out (only it carries a back pointer to the context)
static void set_last_result(iadh_ctx* c, iadh_result* r, size_t body)
{
    if (c->last_result) const_cast<iadh_result*>(c->last_result)->owner = nullptr;
    for (const iadh_result* b : c->last_batch) const_cast<iadh_result*>(b)->owner = nullptr;
    c->last_batch.clear();
    c->last_batch_body = 0;
    c->last_result = r;
    c->last_result_body = body;
    if (r) r->owner = c;
}

// ---------------------------------------------------------------------------------------------
// host-side scheduling (mirrors of the reference's 1-thread orders)

// DataSet::sortGeneLists (DataSet.cpp:1088-1105): size descending; equal sizes come out in
// descending list index because the multimap is walked in reverse
static void sort_by_size(const std::vector<int64_t>& sizes, std::vector<int32_t>& v)
{
    v.resize(sizes.size());
    std::iota(v.begin(), v.end(), 0);
    std::stable_sort(v.begin(), v.end(), [&sizes](int a, int b) {
        if (sizes[a] != sizes[b]) return sizes[a] > sizes[b];
        return a > b;
    });
}

// DataSet::getSomeWorkL2 (levelTwo.cpp:113-175): anti-diagonal walk over the size-sorted lower
// triangle, each pair assigned to the least-loaded rank (getProcForPackage, :98-111)
static void plan_pairs(const std::vector<int64_t>& sizes, const std::vector<int32_t>& index_to_list, int rank,
                       int n_ranks, std::vector<int32_t>& xs, std::vector<int32_t>& ys)
{
    const unsigned n = (unsigned)sizes.size();
    std::vector<uint64_t> load(n_ranks, 0);
    unsigned cX = 0, cY = 0;
    while (cX < n) {
        const unsigned lX = index_to_list[cX], lY = index_to_list[cY];
        const unsigned sx = (unsigned)sizes[lX], sy = (unsigned)sizes[lY];
        const uint64_t weight = (unsigned)(sx * sy);  // 32-bit product, as the reference
        unsigned sp = 0;
        for (unsigned p = 1; p < load.size(); p++)
            if (load[p] < load[sp]) sp = p;
        load[sp] += weight;
        if ((int)sp == rank) {
            xs.push_back((int32_t)lX);
            ys.push_back((int32_t)lY);
        }
        if ((cX - cY) < 2) {
            if (cX == cY && cX == n - 1) break;
            const unsigned d = cX + cY + 1;
            cX = (d < n) ? d : n - 1;
            cY = (d < n) ? 0 : d - n + 1;
        } else {
            cX--;
            cY++;
        }
    }
}

// ParToolBox::statPartWorklSort (parallel.cpp:96-156): size-sorted lists dealt to the lightest package
static void plan_lists(const std::vector<int64_t>& sizes, const std::vector<int32_t>& index_to_list, int rank,
                       int n_ranks, std::vector<int32_t>& out)
{
    std::vector<uint64_t> load(n_ranks, 0);
    for (int32_t l : index_to_list) {
        unsigned sp = 0;
        for (unsigned p = 1; p < load.size(); p++)
            if (load[p] < load[sp]) sp = p;
        load[sp] += (uint64_t)sizes[l];
        if ((int)sp == rank) out.push_back(l);
    }
}

static std::vector<int64_t> ctx_sizes(const iadh_ctx* c)
{
    std::vector<int64_t> s(c->n_lists);
    for (int l = 0; l < c->n_lists; l++) s[l] = c->list_off[l + 1] - c->list_off[l];
    return s;
}
static void sort_lists(iadh_ctx* c) { sort_by_size(ctx_sizes(c), c->index_to_list); }
static void pair_sequence(const iadh_ctx* c, int rank, int n_ranks, std::vector<int32_t>& xs, std::vector<int32_t>& ys)
{
    plan_pairs(ctx_sizes(c), c->index_to_list, rank, n_ranks, xs, ys);
}
static void list_partition(const iadh_ctx* c, int rank, int n_ranks, std::vector<int32_t>& out)
{
    plan_lists(ctx_sizes(c), c->index_to_list, rank, n_ranks, out);
}

// ---------------------------------------------------------------------------------------------

// Result streams travel to the host on the copy stream in pieces of 2 MB: the calls that follow keep reading small
// blocks back (dot counts, outcomes) through the same copy engine, and behind one 300 MB transfer -- 8 ranks pulling
// the gathered level-2 result at once -- such a read-back waited tens of milliseconds (profile batch at 8 GPUs:
// 39 ms, 10 ms of it work).  In pieces, and with the copy stream at the lowest priority, it slips in between.
static cudaError_t copy_to_host_in_pieces(void* dst, const void* src, size_t bytes, cudaStream_t st)
{
    const size_t piece = (size_t)2 << 20;
    for (size_t o = 0; o < bytes; o += piece) {
        const cudaError_t e = cudaMemcpyAsync((uint8_t*)dst + o, (const uint8_t*)src + o, std::min(piece, bytes - o), cudaMemcpyDeviceToHost, st);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

static int upload(DBuf& d, const void* src, size_t bytes, cudaStream_t st)
{
    CU(d.ensure(bytes ? bytes : 16));
    if (bytes) CU(cudaMemcpyAsync(d.p, src, bytes, cudaMemcpyHostToDevice, st));
    return IADH_OK;
}

// stable sort of ids, descending by key(id): counting sort when the key range is small (dots per
// GHM are a few thousand; the batch can be 10^5 GHMs), comparison sort otherwise
template <typename KeyFn>
static void sort_desc_stable(std::vector<int>& ids, KeyFn key)
{
    uint64_t mx = 0;
    for (int id : ids) mx = std::max<uint64_t>(mx, key(id));
    if (mx > ((uint64_t)1 << 22) || ids.size() < 64) {
        std::stable_sort(ids.begin(), ids.end(), [&key](int a, int b) { return key(a) > key(b); });
        return;
    }
    std::vector<uint32_t> start((size_t)mx + 2, 0);
    for (int id : ids) start[mx - key(id) + 1]++;
    for (size_t k = 1; k < start.size(); k++) start[k] += start[k - 1];
    std::vector<int> out(ids.size());
    for (int id : ids) out[start[mx - key(id)]++] = id;
    ids.swap(out);
}

struct ProfileHost {
    ProfileDev dev;
    bool valid = false;
};

// column groups of a profile (see ProfileDev in common.h)
static int profile_groups(const iadh_ctx* c, const iadh_profile_desc* p, std::vector<int32_t>& col_ptr, std::vector<int32_t>& gfam,
                          std::vector<int32_t>& ggid, std::vector<uint8_t>& gstrand)
{
    const int S = p->n_segments, L = p->length;
    if (S < 1 || L < 0) return fail(IADH_ERR_ARG, "bad profile shape S=%d L=%d", S, L);
    if (L > IADH_MAX_COORD) return fail(IADH_ERR_CAPACITY, "profile longer than %d columns", IADH_MAX_COORD);
    if (L > 0 && (!p->gid || !p->num_id || !p->flags || (!c->pairs && !p->fam))) return fail(IADH_ERR_ARG, "profile descriptor with null arrays");
    col_ptr.assign(L + 1, 0);
    gfam.clear();
    ggid.clear();
    gstrand.clear();
    for (int x = 0; x < L; x++) {
        const size_t g0 = gfam.size();
        for (int s = 0; s < S; s++) {
            const size_t k = (size_t)s * L + x;
            if (p->flags[k] & IADH_F_GAP) continue;
            const int gid = p->gid[k];
            const int f = c->pairs ? gid : p->fam[k];
            const uint8_t st = p->flags[k] & IADH_F_STRAND;
            if (f < 0 || f >= c->n_fam) continue;  // family absent from the dataset: no partner possible
            size_t g = g0;
            for (; g < gfam.size(); g++)
                if (gfam[g] == f && gstrand[g] == st) break;
            if (g == gfam.size()) {
                gfam.push_back(f);
                ggid.push_back(gid);
                gstrand.push_back(st);
            } else if (ggid[g] != gid)
                ggid[g] = -1;
        }
        col_ptr[x + 1] = (int32_t)gfam.size();
    }
    return IADH_OK;
}

// the column groups of profile k as work items of the family-expansion build
static void append_hl_items(iadh_ctx* c, int k, const std::vector<int32_t>& col_ptr, const std::vector<int32_t>& gfam,
                            const std::vector<int32_t>& ggid, const std::vector<uint8_t>& gstrand)
{
    if (c->pairs) return;   // (pairs mode builds profile GHMs with the tile kernel)
    const int L = (int)col_ptr.size() - 1;
    for (int x = 0; x < L; x++)
        for (int g = col_ptr[x]; g < col_ptr[x + 1]; g++) {
            HlGroup it;
            it.fam = gfam[g];
            it.gid = ggid[g];
            it.x_strand = (uint32_t)x | (gstrand[g] ? 0x80000000u : 0u);
            it.prof = k;
            c->hl_items.push_back(it);
        }
}

// one profile laid out at h (host staging) / dv (the same bytes on the device); returns the bytes used
static size_t place_profile(const iadh_ctx* c, const iadh_profile_desc* p, const std::vector<int32_t>& col_ptr,
                            const std::vector<int32_t>& gfam, const std::vector<int32_t>& ggid, const std::vector<uint8_t>& gstrand,
                            uint8_t* h, uint8_t* dv, ProfileDev& dev)
{
    const int S = p->n_segments, L = p->length;
    const size_t n = (size_t)S * L, G = gfam.size();
    size_t o = 0;
    auto take = [&o](size_t bytes) {
        const size_t r = o;
        o += (bytes + 15) & ~(size_t)15;
        return r;
    };
    const size_t o_col = take(4 * col_ptr.size()), o_gfam = take(4 * G), o_ggid = take(4 * G), o_gstr = take(G),
                 o_fam = take(4 * n), o_gid = take(4 * n), o_num = take(4 * n), o_flg = take(n);
    if (h) {
        memcpy(h + o_col, col_ptr.data(), 4 * col_ptr.size());
        if (G) {
            memcpy(h + o_gfam, gfam.data(), 4 * G);
            memcpy(h + o_ggid, ggid.data(), 4 * G);
            memcpy(h + o_gstr, gstrand.data(), G);
        }
        if (n) {
            if (c->pairs) {  // bucket key = gene id; ids outside the pair table have no bucket
                int32_t* hf = (int32_t*)(h + o_fam);
                for (size_t k = 0; k < n; k++) hf[k] = (p->gid[k] >= 0 && p->gid[k] < c->n_fam) ? p->gid[k] : -1;
            } else
                memcpy(h + o_fam, p->fam, 4 * n);
            memcpy(h + o_gid, p->gid, 4 * n);
            memcpy(h + o_num, p->num_id, 4 * n);
            memcpy(h + o_flg, p->flags, n);
        }
        dev.S = S;
        dev.L = L;
        dev.col_ptr = (const int32_t*)(dv + o_col);
        dev.grp_fam = (const int32_t*)(dv + o_gfam);
        dev.grp_gid = (const int32_t*)(dv + o_ggid);
        dev.grp_strand = dv + o_gstr;
        dev.fam = (const int32_t*)(dv + o_fam);
        dev.gid = (const int32_t*)(dv + o_gid);
        dev.num_id = (const int32_t*)(dv + o_num);
        dev.flags = dv + o_flg;
        dev.batch = nullptr;
        dev.vis = nullptr;
    }
    return o;
}

// groups per profile column + raw arrays -> device (one pinned staging buffer, one H2D copy: the profile is
// re-sent on every search)
static int upload_profile(iadh_ctx* c, const iadh_profile_desc* p, ProfileHost& out, int64_t& h2d)
{
    std::vector<int32_t> col_ptr, gfam, ggid;
    std::vector<uint8_t> gstrand;
    int rc = profile_groups(c, p, col_ptr, gfam, ggid, gstrand);
    if (rc) return rc;
    c->hl_items.clear();
    append_hl_items(c, 0, col_ptr, gfam, ggid, gstrand);
    ProfileDev dummy;
    const size_t bytes = place_profile(c, p, col_ptr, gfam, ggid, gstrand, nullptr, nullptr, dummy);
    CU(c->h_pf.ensure(bytes + 16));
    CU(c->d_pf_col.ensure(bytes + 16));
    place_profile(c, p, col_ptr, gfam, ggid, gstrand, c->h_pf.as<uint8_t>(), c->d_pf_col.as<uint8_t>(), out.dev);
    CU(cudaMemcpyAsync(c->d_pf_col.p, c->h_pf.p, bytes, cudaMemcpyHostToDevice, c->st));
    h2d += (int64_t)bytes;
    out.valid = true;
    return IADH_OK;
}

// The inherited Link sets of a call's profiles (iadh_profile_search*_filtered) -> device tables of the homolog-count
// filter (homolog_kernel.cu): one CSR "profile element -> columns of its Link partners" over the concatenated elements
// of the profiles this rank searches, built by kernels from the raw (segment, numID) pairs.  link_ptr[k] .. link_ptr[k+1]
// = Links of profile k; mine[k] == 0: not searched by this rank (no table).  The profiles must already be on their way
// to the device (pf: the call's descriptor, as run_batch gets it).
static int upload_links(iadh_ctx* c, const iadh_profile_desc* profiles, int n_profiles, const uint8_t* mine, const iadh_link_t* links,
                        const int64_t* link_ptr, const ProfileDev& pf, iadh_stats& st)
{
    static_assert(sizeof(LinkIn) == sizeof(iadh_link_t), "Links are uploaded as the caller laid them out");
    if (!links || !link_ptr) return fail(IADH_ERR_ARG, "homolog filter: null Link arrays");
    if (link_ptr[0] < 0) return fail(IADH_ERR_ARG, "homolog filter: link_ptr[0] < 0");
    // this rank's profiles: their Links back to back on the device, their elements numbered consecutively
    std::vector<int64_t> lptr((size_t)n_profiles + 1, 0);
    std::vector<int32_t> ebase((size_t)n_profiles + 1, 0);
    int64_t n_elems = 0, nl = 0;
    for (int k = 0; k < n_profiles; k++) {
        const int64_t b = link_ptr[k], e = link_ptr[k + 1];
        if (e < b) return fail(IADH_ERR_ARG, "homolog filter: link_ptr not ascending at profile %d", k);
        if (!mine || mine[k]) {
            if (e == b)
                return fail(IADH_ERR_ARG, "homolog filter: profile %d without Links (a profile's multiplicon owns the Links of its anchor points)", k);
            const int S = profiles[k].n_segments, L = profiles[k].length;
            for (int64_t i = b; i < e; i++) {
                const iadh_link_t& l = links[i];
                if (l.segment_x < 0 || l.segment_x >= S || l.segment_y < 0 || l.segment_y >= S)
                    return fail(IADH_ERR_ARG, "homolog filter: Link %lld of profile %d names segment %d / %d of %d", (long long)(i - b), k, l.segment_x, l.segment_y, S);
            }
            n_elems += (int64_t)S * L;
            nl += e - b;
            if (n_elems > 0x3ffffff0ll || nl > 0x3ffffff0ll)
                return fail(IADH_ERR_CAPACITY, "homolog filter: more than 2^30 profile elements or Links in one call");
        }
        lptr[k + 1] = nl;
        ebase[k + 1] = (int32_t)n_elems;
    }
    uint32_t slots = 16;
    while ((int64_t)slots < 2 * n_elems) slots <<= 1;
    int rc;
    CU(c->d_links.ensure(sizeof(LinkIn) * (size_t)nl + 16));
    CU(c->d_lk_cols.ensure(8 * (size_t)nl + 16));
    CU(c->d_lk_col.ensure(8 * (size_t)nl + 16));
    CU(c->d_lk_cnt.ensure(4 * (size_t)n_elems + 16));
    CU(c->d_lk_ptr.ensure(4 * (size_t)(n_elems + 1) + 16));
    CU(c->d_lk_bsum.ensure(4 * (size_t)(n_elems / 1024 + 2) + 16));
    CU(c->d_lk_keys.ensure(8 * (size_t)slots));
    CU(c->d_lk_vals.ensure(4 * (size_t)slots));
    for (int k = 0; k < n_profiles; k++)
        if (lptr[k + 1] > lptr[k])
            CU(cudaMemcpyAsync(c->d_links.as<LinkIn>() + lptr[k], links + link_ptr[k], sizeof(LinkIn) * (size_t)(lptr[k + 1] - lptr[k]),
                               cudaMemcpyHostToDevice, c->st));
    if ((rc = upload(c->d_link_ptr, lptr.data(), 8 * lptr.size(), c->st))) return rc;
    if ((rc = upload(c->d_elem_base, ebase.data(), 4 * ebase.size(), c->st))) return rc;
    std::vector<HomologDev> hd((size_t)n_profiles);
    for (int k = 0; k < n_profiles; k++) {
        hd[k].lk_ptr = c->d_lk_ptr.as<int32_t>() + ebase[k];   // (profiles of other ranks: never dereferenced, no GHM names them)
        hd[k].lk_col = c->d_lk_col.as<int32_t>();
    }
    if ((rc = upload(c->d_hdev, hd.data(), sizeof(HomologDev) * hd.size(), c->st))) return rc;
    launch_link_tables(c->d_links.as<LinkIn>(), c->d_link_ptr.as<int64_t>(), nl, c->d_elem_base.as<int32_t>(), n_profiles, n_elems, pf,
                       c->d_lk_keys.as<unsigned long long>(), c->d_lk_vals.as<int32_t>(), slots - 1, c->d_lk_cols.as<int32_t>(),
                       c->d_lk_cnt.as<int32_t>(), c->d_lk_ptr.as<int32_t>(), c->d_lk_col.as<int32_t>(), c->d_lk_bsum.as<int32_t>(), c->st);
    CU(cudaGetLastError());
    st.h2d_bytes += (int64_t)(sizeof(LinkIn) * (size_t)nl + 12 * (size_t)(n_profiles + 1) + sizeof(HomologDev) * hd.size());
    st.kernel_launches += (n_elems > 0 ? 1 : 0) + (nl > 0 ? 3 : 0) + 3;   // hash, resolve, count, fill, 3 x scan
    return IADH_OK;
}

static const int RC_SPLIT = 1000;  // internal: batch exceeds the workspace budget

struct BatchOut {
    std::vector<GhmTask> tasks;  // final state (cnt, n_mult, out_bytes ...)
    HBuf packed;                 // pinned: int32 count + concatenated per-GHM blocks in seq order
    size_t packed_bytes = 0;     // incl. the 4-byte count
    cudaEvent_t ready = nullptr; // the copy into `packed` runs on the context's copy stream; set when one is in flight
    // groups of consecutive GHMs (the profiles of iadh_profile_search_batch) whose multiplicons form a stream of their
    // own inside `packed`: [int32 count][blocks] per group, group k at group_off[k] (n_groups + 1 entries)
    const std::vector<int>* group_begin = nullptr;   // in: first task of every group (+ the task count)
    std::vector<size_t> group_off;
    std::vector<int32_t> ghm_index;
    iadh_stats st;
    void wait()
    {
        if (ready) {
            cudaEventSynchronize(ready);
            cudaEventDestroy(ready);
            ready = nullptr;
        }
    }
    ~BatchOut()
    {
        wait();
        packed.release();
    }
};

// The task records stay on the device; only the fields that change hands travel, packed:
//   cnt (device -> host after the count pass), the plan (host -> device), the outcome (device -> host).
struct TaskPlan {
    int64_t dot_off, ws_off, ws_bytes;
    int32_t half_mask, half_needed;
};
struct TaskOut {
    int64_t out_bytes;
    int32_t status, n_mult, n_clusters, n_aps;
};
// Small read-backs (dot counts, outcomes, gather sizes) are written into pinned host memory BY A KERNEL (the pinned
// block is mapped into the device's address space: unified addressing), not by the copy engine: the copy engine may be
// busy for tens of milliseconds with a result stream on the copy stream (8 ranks pulling the gathered level-2 result
// at once), and a 1 MB read-back queued behind it stalled the profile batch that followed (36 ms instead of 10).
__global__ void readback_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst_host, size_t n_words)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += (size_t)gridDim.x * blockDim.x) dst_host[i] = src[i];
}
static cudaError_t readback(void* host_pinned, const void* dev, size_t bytes, cudaStream_t st)
{
    const size_t n_words = (bytes + 3) / 4;   // (blocks are allocated with slack: a rounded-up word is in range)
    if (n_words == 0) return cudaSuccess;
    readback_kernel<<<(unsigned)std::min<size_t>((n_words + 255) / 256, 592), 256, 0, st>>>((const uint32_t*)dev, (uint32_t*)host_pinned, n_words);
    return cudaGetLastError();
}
__global__ void gather_cnt_kernel(const GhmTask* __restrict__ tasks, int n, uint32_t* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        out[2 * i] = tasks[i].cnt[0];
        out[2 * i + 1] = tasks[i].cnt[1];
    }
}
__global__ void scatter_plan_kernel(GhmTask* __restrict__ tasks, int n, const TaskPlan* __restrict__ plan)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    GhmTask& t = tasks[i];
    const TaskPlan p = plan[i];
    t.dot_off = p.dot_off;
    t.ws_off = p.ws_off;
    t.ws_bytes = p.ws_bytes;
    t.half_mask = p.half_mask;
    t.half_needed = p.half_needed;
    t.half_done = 0;
    t.status = 0;
    t.n_mult = t.n_clusters = t.n_aps = 0;
    t.out_bytes = 0;
    for (int o = 0; o < 2; o++) t.half_nbc[o] = t.half_ncl[o] = t.half_pool[o] = t.half_status[o] = t.half_nk[o] = t.half_g3[o] = t.half_seed[o] = 0;
}
// fallback of resident GHMs: fresh outcome fields before the fused kernel re-runs them (items = index * 4 + mask)
__global__ void reset_subset_kernel(GhmTask* __restrict__ tasks, const int* __restrict__ items, int n)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    GhmTask& t = tasks[items[k] >> 2];
    t.half_done = 0;
    t.half_needed = 1;
    t.status = 0;
    t.n_mult = t.n_clusters = t.n_aps = 0;
    t.out_bytes = 0;
    for (int o = 0; o < 2; o++) t.half_nbc[o] = t.half_ncl[o] = t.half_pool[o] = t.half_status[o] = t.half_nk[o] = t.half_g3[o] = t.half_seed[o] = 0;
}
// what the host knows about a GHM before the build (the rest of the device record starts as zero)
struct TaskInit {
    int32_t x_list, y_list, x_object_id, y_object_id, level, seq, prof, pad;
    uint64_t area;
};
__global__ void init_tasks_kernel(GhmTask* __restrict__ tasks, int n, const TaskInit* __restrict__ init)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    GhmTask t;
    memset(&t, 0, sizeof t);
    const TaskInit v = init[i];
    t.x_list = v.x_list; t.y_list = v.y_list; t.x_object_id = v.x_object_id; t.y_object_id = v.y_object_id;
    t.level = v.level; t.seq = v.seq; t.prof = v.prof;
    t.area = v.area;
    tasks[i] = t;
}
__global__ void scatter_dot_off_kernel(GhmTask* __restrict__ tasks, int n, const int64_t* __restrict__ dot_off)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) tasks[i].dot_off = dot_off[i];
}
// the 4-byte multiplicon counts between the groups of a packed stream (run_batch): hdr = (byte offset, count) pairs
__global__ void group_header_kernel(uint8_t* __restrict__ out, const int64_t* __restrict__ hdr, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t v = (int32_t)hdr[2 * i + 1];
    uint8_t* p = out + hdr[2 * i];   // (unaligned: the blocks before it have odd sizes)
    p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24);
}
__global__ void gather_out_kernel(const GhmTask* __restrict__ tasks, int n, TaskOut* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    TaskOut o;
    o.out_bytes = tasks[i].out_bytes;
    o.status = tasks[i].status;
    o.n_mult = tasks[i].n_mult;
    o.n_clusters = tasks[i].n_clusters;
    o.n_aps = tasks[i].n_aps;
    out[i] = o;
}

// host plan -> device task records (pinned staging, one copy, one kernel)
static int send_plan(iadh_ctx* c, const std::vector<GhmTask>& tasks, iadh_stats& st)
{
    const int n = (int)tasks.size();
    if (n == 0) return IADH_OK;
    CU(c->h_tasks.ensure(sizeof(TaskPlan) * (size_t)n));
    CU(c->d_plan.ensure(sizeof(TaskPlan) * (size_t)n));
    TaskPlan* h = c->h_tasks.as<TaskPlan>();
    for (int i = 0; i < n; i++) {
        h[i].dot_off = tasks[i].dot_off;
        h[i].ws_off = tasks[i].ws_off;
        h[i].ws_bytes = tasks[i].ws_bytes;
        h[i].half_mask = tasks[i].half_mask;
        h[i].half_needed = tasks[i].half_needed;
    }
    CU(cudaMemcpyAsync(c->d_plan.p, h, sizeof(TaskPlan) * (size_t)n, cudaMemcpyHostToDevice, c->st));
    scatter_plan_kernel<<<(n + 255) / 256, 256, 0, c->st>>>(c->d_tasks.as<GhmTask>(), n, c->d_plan.as<TaskPlan>());
    st.h2d_bytes += (int64_t)sizeof(TaskPlan) * n;
    st.kernel_launches += 1;
    return IADH_OK;
}

// stages 1..4: build the dots of every task; on return tasks hold cnt/dot_off and d_dots the keys
static int build_stage(iadh_ctx* c, bool hl, const ProfileDev& pf, std::vector<GhmTask>& tasks, int64_t& total_dots,
                       iadh_stats& st, const std::function<int(std::vector<GhmTask>&)>& plan = nullptr)
{
    const int n = (int)tasks.size();
    int rc;
    // Path choice.  Many sparse level-2 GHMs: the family-expansion kernels (work = dots); otherwise (profile GHMs,
    // small calls, dense GHMs, a pair named twice) the tile kernels (work = GHM columns).
    const int nl = c->n_lists;
    const int64_t n_prof = hl ? (pf.batch ? (int64_t)c->batch_L.size() : 1) : 0;
    bool expand = hl ? (!c->pairs && n >= c->hl_expand_min && n_prof * nl <= ((int64_t)4 << 20))
                     : (n >= c->expand_min && (int64_t)nl * nl <= ((int64_t)64 << 20));
    if (expand) {
        // (x list | profile, y list) -> task; a pair named twice sends the call to the tile kernels
        c->h_pair_idx.assign((size_t)(hl ? n_prof : nl) * nl, -1);
        c->h_row_used.assign((size_t)nl, 0);
        for (int i = 0; i < n && expand; i++) {
            int32_t& slot = c->h_pair_idx[(size_t)(hl ? tasks[i].prof : tasks[i].x_list) * nl + tasks[i].y_list];
            if (slot >= 0) expand = false;
            slot = i;
            if (!hl) c->h_row_used[tasks[i].x_list] = 1;
        }
    }
    std::vector<int64_t> tile_prefix(n + 1, 0);
    auto make_tile_prefix = [&]() {
        for (int i = 0; i < n; i++) {
            // (a batch of profiles: pf.batch is a device pointer, the lengths come from the host copy)
            const int64_t nx = hl ? (pf.batch ? c->batch_L[tasks[i].prof] : pf.L) : (c->list_off[tasks[i].x_list + 1] - c->list_off[tasks[i].x_list]);
            tile_prefix[i + 1] = tile_prefix[i] + (nx + 31) / 32;
        }
    };
    if (!expand) make_tile_prefix();
    trace_mark(c->st, "host: tile prefix");
    // the device task records are built from what the host knows (40 of their 184 bytes), through pinned staging
    CU(c->d_tasks.ensure(sizeof(GhmTask) * (size_t)n + 16));
    CU(c->h_init.ensure((sizeof(TaskInit) + 8) * (size_t)(n + 1)));
    CU(c->d_init.ensure((sizeof(TaskInit) + 8) * (size_t)(n + 1)));
    {
        TaskInit* hi = c->h_init.as<TaskInit>();
        for (int i = 0; i < n; i++) {
            const GhmTask& t = tasks[i];
            hi[i].x_list = t.x_list; hi[i].y_list = t.y_list; hi[i].x_object_id = t.x_object_id; hi[i].y_object_id = t.y_object_id;
            hi[i].level = t.level; hi[i].seq = t.seq; hi[i].prof = t.prof; hi[i].pad = 0;
            hi[i].area = t.area;
        }
        int64_t* hp = (int64_t*)(hi + n);
        memcpy(hp, tile_prefix.data(), 8 * (size_t)(n + 1));
        CU(cudaMemcpyAsync(c->d_init.p, hi, sizeof(TaskInit) * (size_t)n + 8 * (size_t)(n + 1), cudaMemcpyHostToDevice, c->st));
        init_tasks_kernel<<<(n + 255) / 256, 256, 0, c->st>>>(c->d_tasks.as<GhmTask>(), n, c->d_init.as<TaskInit>());
        st.kernel_launches += 1;
    }
    int64_t* d_tile_prefix = (int64_t*)(c->d_init.as<TaskInit>() + n);
    st.h2d_bytes += (int64_t)sizeof(TaskInit) * n + 8 * (n + 1);
    CU(c->d_cnt.ensure(8 * (size_t)n + 16));
    CU(c->h_misc.ensure(8 * (size_t)n + 16));
    const int64_t n_elems = nl ? c->list_off[nl] : 0;
    if (expand) {
        if ((rc = upload(c->d_pair_idx, c->h_pair_idx.data(), 4 * c->h_pair_idx.size(), c->st))) return rc;
        if ((rc = upload(c->d_row_used, c->h_row_used.data(), c->h_row_used.size(), c->st))) return rc;
        st.h2d_bytes += 4 * (int64_t)c->h_pair_idx.size() + nl;
        if (hl) {
            if ((rc = upload(c->d_hl_items, c->hl_items.data(), sizeof(HlGroup) * c->hl_items.size(), c->st))) return rc;
            st.h2d_bytes += (int64_t)(sizeof(HlGroup) * c->hl_items.size());
        }
        CU(cudaMemsetAsync(c->d_cnt.p, 0, 8 * (size_t)n, c->st));
        CU(cudaEventRecord(c->ev[1], c->st));
        if (hl)
            launch_hl_expand_count(c->ds, c->d_hl_items.as<HlGroup>(), (int)c->hl_items.size(), c->d_pair_idx.as<int32_t>(), pf.vis,
                                   c->d_cnt.as<uint32_t>(), c->sm_count, c->st);
        else
            launch_expand_count(c->ds, n_elems, c->d_pair_idx.as<int32_t>(), c->d_row_used.as<uint8_t>(), c->d_cnt.as<uint32_t>(), c->sm_count, c->st);
        CU(cudaEventRecord(c->ev[2], c->st));
        CU(readback(c->h_misc.p, c->d_cnt.p, 8 * (size_t)n, c->st));
        CU(cudaStreamSynchronize(c->st));
        trace_mark(c->st, "K1 expand count + counts D2H");
        st.kernel_launches += 1;
        st.d2h_bytes += 8 * (int64_t)n;
        // the block sort of this path works in shared memory: a dense GHM sends the call to the tile kernels
        const uint32_t* hc = c->h_misc.as<uint32_t>();
        const uint32_t lim = (uint32_t)(c->expand_sort_max > 0 ? std::min(c->expand_sort_max, expand_sort_max()) : expand_sort_max());
        for (int i = 0; i < 2 * n && expand; i++)
            if (hc[i] > lim) expand = false;
        if (!expand) {
            make_tile_prefix();
            CU(cudaMemcpyAsync(d_tile_prefix, tile_prefix.data(), 8 * (size_t)(n + 1), cudaMemcpyHostToDevice, c->st));
            CU(cudaStreamSynchronize(c->st));   // (pageable source)
        }
    }
    const int64_t n_tiles = tile_prefix[n];
    if (!expand) {
        CU(c->d_tile_cnt.ensure(8 * (size_t)n_tiles + 16));
        CU(c->d_tile_off.ensure(8 * (size_t)n_tiles + 16));
        CU(cudaEventRecord(c->ev[1], c->st));
        launch_build_count(hl, c->ds, pf, c->d_tasks.as<GhmTask>(), n, d_tile_prefix, n_tiles,
                           c->d_tile_cnt.as<uint32_t>(), c->sm_count, c->st);
        launch_tile_scan(d_tile_prefix, n, c->d_tile_cnt.as<uint32_t>(), c->d_tile_off.as<uint32_t>(), c->st);
        CU(cudaEventRecord(c->ev[2], c->st));
        trace_mark(c->st, "K1 count + tile scan");
        st.kernel_launches += (n_tiles ? 1 : 0) + (n ? 1 : 0);
        gather_cnt_kernel<<<(n + 255) / 256, 256, 0, c->st>>>(c->d_tasks.as<GhmTask>(), n, c->d_cnt.as<uint32_t>());
        CU(readback(c->h_misc.p, c->d_cnt.p, 8 * (size_t)n, c->st));
        CU(cudaStreamSynchronize(c->st));  // (tile_prefix dies at return; the counts are read next)
        trace_mark(c->st, "counts D2H");
        st.d2h_bytes += 8 * (int64_t)n;
        st.kernel_launches += 1;
    }
    total_dots = 0;
    {
        const uint32_t* hc = c->h_misc.as<uint32_t>();
        for (int i = 0; i < n; i++) {
            tasks[i].cnt[0] = hc[2 * i];
            tasks[i].cnt[1] = hc[2 * i + 1];
            tasks[i].dot_off = total_dots;
            total_dots += (int64_t)tasks[i].cnt[0] + tasks[i].cnt[1];
        }
    }
    if (total_dots > 0x7fffffff00ll) return fail(IADH_ERR_CAPACITY, "too many dots in one call");
    CU(c->d_dots.ensure(4 * (size_t)total_dots + 16));
    trace_mark(c->st, "host: dot offsets");
    // the fill pass needs the dot offsets only: they go first, the fill kernel starts, and the host plans the
    // clustering stage (orders, workspaces, queue items: milliseconds for 10^5 GHMs) while it runs
    {
        CU(c->h_tasks.ensure(sizeof(TaskPlan) * (size_t)n + 16));   // (sized for the plan that follows)
        CU(c->d_plan.ensure(sizeof(TaskPlan) * (size_t)n + 16));
        int64_t* h = c->h_tasks.as<int64_t>();
        for (int i = 0; i < n; i++) h[i] = tasks[i].dot_off;
        CU(cudaMemcpyAsync(c->d_plan.p, h, 8 * (size_t)n, cudaMemcpyHostToDevice, c->st));
        scatter_dot_off_kernel<<<(n + 255) / 256, 256, 0, c->st>>>(c->d_tasks.as<GhmTask>(), n, c->d_plan.as<int64_t>());
        st.h2d_bytes += 8 * (int64_t)n;
        st.kernel_launches += 1;
    }
    if (expand) {
        CU(c->d_cur.ensure(8 * (size_t)n + 16));
        CU(cudaMemsetAsync(c->d_cur.p, 0, 8 * (size_t)n, c->st));
    }
    CU(cudaEventRecord(c->ev[3], c->st));
    if (expand) {
        // (d_plan holds the dot offsets until the plan is sent, which is stream-ordered after these kernels)
        if (hl)
            launch_hl_expand_fill(c->ds, c->d_hl_items.as<HlGroup>(), (int)c->hl_items.size(), c->d_pair_idx.as<int32_t>(), pf.vis,
                                  c->d_cur.as<uint32_t>(), c->d_plan.as<int64_t>(), c->d_cnt.as<uint32_t>(), c->d_dots.as<uint32_t>(),
                                  c->d_tasks.as<GhmTask>(), n, c->sm_count, c->st);
        else
            launch_expand_fill(c->ds, n_elems, c->d_pair_idx.as<int32_t>(), c->d_row_used.as<uint8_t>(), c->d_cur.as<uint32_t>(),
                               c->d_plan.as<int64_t>(), c->d_cnt.as<uint32_t>(), c->d_dots.as<uint32_t>(), c->d_tasks.as<GhmTask>(), n,
                               c->sm_count, c->st);
        st.kernel_launches += 3;
    } else {
        launch_build_fill(hl, c->ds, pf, c->d_tasks.as<GhmTask>(), n, d_tile_prefix, n_tiles,
                          c->d_tile_off.as<uint32_t>(), c->d_dots.as<uint32_t>(), c->sm_count, c->st);
        st.kernel_launches += n_tiles ? 1 : 0;
    }
    CU(cudaEventRecord(c->ev[4], c->st));
    st.n_dots = total_dots;
    if (plan) {
        // (the staging buffer of the offsets is reused by the plan: the copy above must have left it)
        if ((rc = plan(tasks))) return rc;
        trace_mark(nullptr, "host: plan (overlaps K1 fill)");
        CU(cudaEventSynchronize(c->ev[3]));
        if ((rc = send_plan(c, tasks, st))) return rc;
    }
    trace_mark(c->st, "K1 fill + plan H2D / scatter");
    return IADH_OK;
}

// the whole pipeline for a batch of GHMs (tasks[i].seq == i on entry)
static int run_batch(iadh_ctx* c, bool hl, const ProfileDev& pf, std::vector<GhmTask>& tasks, BatchOut& out)
{
    set_last_result(c, nullptr, 0);   // (d_out is about to be overwritten)
    iadh_stats& st = out.st;
    const int n = (int)tasks.size();
    CU(cudaEventRecord(c->ev[0], c->st));
    trace_mark(c->st, nullptr);
    int64_t total_dots = 0;
    int rc;
    if (n == 0) {   // nothing to search: empty streams (count 0), one per group
        const size_t ng = out.group_begin ? out.group_begin->size() - 1 : 1;
        out.tasks.clear();
        CU(out.packed.ensure(4 * ng));
        memset(out.packed.p, 0, 4 * ng);
        out.packed_bytes = 4 * ng;
        out.group_off.resize(ng + 1);
        for (size_t k = 0; k <= ng; k++) out.group_off[k] = 4 * k;
        return IADH_OK;
    }
    // host-side plan of the clustering stage: depends only on the dot counts, so it is made between
    // the count pass and the fill pass (inside build_stage) and travels to the device with dot_off
    int pool_factor = c->pool_factor;
    std::vector<int> order, items, items1, chunk_k, chunk_i1;
    std::vector<uint32_t> dcnt;
    std::vector<int64_t> wsb;       // per GHM: workspace bytes, orientation mask (compact copies: the task records are
    std::vector<uint8_t> hmask;     // 184 bytes apart and the loops below visit them in heaviest-first order)
    size_t ws_total = 0;
    bool pass_mode = false;
    int n_act = 0;
    // resident K2: GHMs whose state fits the shared memory of one CTA (RsLayout), grouped into launches by
    // (shared-memory budget, CTA width); force_global = outgrew a budget in an earlier attempt
    struct ResLaunch {
        int nt, smem, begin, count;
    };
    std::vector<ResLaunch> res_launches;
    std::vector<int> res_items, order_all;
    std::vector<uint8_t> is_res((size_t)n, 0), force_global((size_t)n, 0);
    static const int res_buckets[] = {8 << 10, 12 << 10, 16 << 10, 24 << 10, 32 << 10, 48 << 10, 64 << 10, 96 << 10, 128 << 10, 200 << 10};
    auto make_plan = [&](std::vector<GhmTask>& tasks) -> int {
        items.clear();
        items1.clear();
        ws_total = 0;
        // active GHMs: at least anchor_points dots (GHM.cpp:159), heaviest first
        order.clear();
        dcnt.resize(2 * (size_t)n);   // the sort keys, compact (the task records are 184 bytes apart)
        wsb.assign((size_t)n, 0);
        hmask.assign((size_t)n, 0);
        for (int i = 0; i < n; i++) {
            GhmTask& t = tasks[i];
            dcnt[2 * (size_t)i] = t.cnt[0];
            dcnt[2 * (size_t)i + 1] = t.cnt[1];
            const int64_t D = (int64_t)t.cnt[0] + t.cnt[1];
            t.status = 0;
            t.n_mult = t.n_clusters = t.n_aps = 0;
            t.out_bytes = 0;
            t.ws_off = 0;
            t.ws_bytes = 0;
            t.half_done = 0;
            t.half_mask = 0;
            t.half_needed = 1;
            if (D < (int64_t)c->prm.anchor_points || D == 0) continue;
            if (D > 0x0fffffff) return fail(IADH_ERR_CAPACITY, "GHM with %lld dots exceeds the per-GHM limit", (long long)D);
            if (!ws_pool_fits(D, pool_factor))
                return fail(IADH_ERR_CAPACITY, "GHM with %lld dots: anchor-point pool at factor %d exceeds 2^31 entries", (long long)D, pool_factor);
            const WsLayout L = ws_layout((int)D, pool_factor);
            t.ws_off = (int64_t)ws_total;
            t.ws_bytes = (int64_t)L.total;
            ws_total += L.total;
            // an orientation with fewer than 3 dots cannot seed a cluster: no queue item of its own
            t.half_mask = (t.cnt[0] >= 3 ? 1 : 0) | (t.cnt[1] >= 3 ? 2 : 0);
            t.half_needed = 1;
            wsb[i] = t.ws_bytes;
            hmask[i] = (uint8_t)t.half_mask;
            order.push_back(i);
        }
        trace_mark(nullptr, "  plan: workspaces");
        sort_desc_stable(order, [this_dcnt = dcnt.data()](int a) { return (uint64_t)this_dcnt[2 * (size_t)a] + this_dcnt[2 * (size_t)a + 1]; });
        trace_mark(nullptr, "  plan: order");
        order_all = order;
        // split off the GHMs of the resident path; `order` keeps the rest (global-memory kernels below)
        res_launches.clear();
        res_items.clear();
        std::fill(is_res.begin(), is_res.end(), 0);
        if ((c->resident == 1 || (c->resident == 2 && hl && pf.batch && (int)order.size() >= c->pass_min)) && c->prm.n_gaps > 0) {
            int64_t sum_d = 0, max_d = 1;
            for (int i : order) {
                const int64_t D = (int64_t)tasks[i].cnt[0] + tasks[i].cnt[1];
                sum_d += D;
                max_d = std::max(max_d, D);
            }
            // a throughput-bound batch runs every GHM on one warp (most GHMs in flight per SM); a latency-bound
            // call (a profile search: the heaviest GHM is the critical path) gives heavy GHMs wide CTAs
            const bool throughput = (int)order.size() >= c->pass_min && sum_d / max_d >= c->pass_ratio;
            std::map<std::pair<int, int>, std::vector<int>> classes;   // (-smem, nt) -> GHMs, heaviest first
            std::vector<int> rest;
            for (int i : order) {
                const int D0 = (int)tasks[i].cnt[0], D1 = (int)tasks[i].cnt[1];
                const int nt = (!throughput && D0 + D1 >= c->res_wide_min) ? c->res_wide_nt : 32;
                const int64_t need = (int64_t)D0 + D1 > 100000 ? ((int64_t)1 << 40) : (int64_t)rs_layout(D0, D1, nt, c->prm.rs_squeeze).total;
                if (force_global[i] || need > c->res_max_bytes) {
                    rest.push_back(i);
                    continue;
                }
                int b = 0;
                while (res_buckets[b] < need) b++;
                classes[std::make_pair(-std::min(res_buckets[b], c->res_max_bytes), nt)].push_back(i);
                is_res[i] = 1;
            }
            for (auto& kv : classes) {
                ResLaunch L;
                L.nt = kv.first.second;
                L.smem = -kv.first.first;
                L.begin = (int)res_items.size();
                L.count = (int)kv.second.size();
                res_items.insert(res_items.end(), kv.second.begin(), kv.second.end());
                res_launches.push_back(L);
            }
            order.swap(rest);
        }
        n_act = (int)order.size();
        // queue items: GHM * 4 + orientation mask.  The two orientations of a GHM are independent up to
        // the final enrichment; when both are heavy they get an item (a CTA) each, the heavier first --
        // that shortens the critical path of the heaviest GHMs, which is what a call waits for.
        items.reserve(2 * (size_t)n_act);
        // one pass per launch pays when the batch is throughput-bound: many GHMs and no single GHM
        // that would become the critical path on a 32-thread CTA (total dots >> heaviest GHM's dots)
        int64_t sum_d = 0, max_d = 1;
        for (int k = 0; k < n_act; k++) {
            const int64_t D = (int64_t)dcnt[2 * (size_t)order[k]] + dcnt[2 * (size_t)order[k] + 1];
            sum_d += D;
            max_d = std::max(max_d, D);
        }
        pass_mode = n_act >= c->pass_min && c->prm.n_gaps > 0 && sum_d / max_d >= (hl ? c->pass_ratio_hl : c->pass_ratio);
        if (pass_mode) {
            // phase 1 runs one pass per launch over (GHM, orientation) items, heaviest orientation first;
            // phase 2 through the fused kernel with "phase 2 only" items
            // The batch can be cut into chunks of GHMs (contiguous in the heaviest-first order,
            // IADH_PASS_CHUNK_MB of workspace each) whose 34 launches run before the next chunk starts, so
            // that a GHM's state is re-read from L2 instead of HBM between its passes; off by default
            // (see pass_chunk_bytes).
            items1.reserve(2 * (size_t)n_act);
            chunk_k.clear();
            chunk_i1.clear();
            int64_t acc = 0;
            for (int k = 0; k < n_act; k++) {
                const int i = order[k];
                if (k == 0 || acc + wsb[i] > c->pass_chunk_bytes) {
                    chunk_k.push_back(k);
                    chunk_i1.push_back((int)items1.size());
                    acc = 0;
                }
                acc += wsb[i];
                if (hmask[i] & 1) items1.push_back(i * 2);
                if (hmask[i] & 2) items1.push_back(i * 2 + 1);
                items.push_back(i * 4);   // (half_needed stays 1)
            }
            chunk_k.push_back(n_act);
            chunk_i1.push_back((int)items1.size());
            if (chunk_k.size() == 2)  // one chunk: heaviest orientation first over the whole batch
                sort_desc_stable(items1, [this_dcnt = dcnt.data()](int a) { return (uint64_t)this_dcnt[a]; });
        }
        for (int k = 0; k < n_act && !pass_mode; k++) {
            const int i = order[k];
            const uint32_t c0 = dcnt[2 * (size_t)i], c1 = dcnt[2 * (size_t)i + 1];
            if (hmask[i] == 3 && (int)std::min(c0, c1) >= c->split_min) {
                const int first = (c1 > c0) ? 2 : 1;
                items.push_back(i * 4 + first);
                items.push_back(i * 4 + (first ^ 3));
                tasks[i].half_needed = 2;
            } else
                items.push_back(i * 4 + hmask[i]);   // (half_needed stays 1)
        }
        trace_mark(nullptr, "  plan: queue items");
        // too much for one go: the caller bisects the batch (the count pass that led here is cheap)
        if ((int64_t)ws_total > c->ws_budget && n > 1 && !hl) return RC_SPLIT;
        CU(c->d_ws.ensure(ws_total + 256));
        trace_mark(nullptr, "  plan: workspace buffer");
        return IADH_OK;
    };
    if ((rc = build_stage(c, hl, pf, tasks, total_dots, st, make_plan))) return rc;

    int64_t st_fallbacks = 0;
    for (int attempt = 0;; attempt++) {
        if (attempt > 0) {  // rare: the clustering stage is rerun with a larger anchor-point pool
            if ((rc = make_plan(tasks))) return rc;
            if ((rc = send_plan(c, tasks, st))) return rc;
        }
        const int n_items = (int)items.size();
        if ((rc = upload(c->d_items, items.data(), 4 * (size_t)n_items, c->st))) return rc;
        if (pass_mode && (rc = upload(c->d_items1, items1.data(), 4 * items1.size(), c->st))) return rc;
        st.h2d_bytes += 4 * (int64_t)n_items + 4 * (int64_t)items1.size();
        CU(c->d_prof.ensure(sizeof(GhmProf) * (size_t)n + 16));
        CU(cudaMemsetAsync(c->d_prof.p, 0, sizeof(GhmProf) * (size_t)n, c->st));
        CU(c->d_counter.ensure(16 + 4 * 1024 + 4 * 64));  // 4 class queues + per-SM arrival counters + pass queues
        CU(cudaMemsetAsync(c->d_counter.p, 0, 16 + 4 * 1024 + 4 * 64, c->st));
        if (!res_items.empty() && (rc = upload(c->d_items_res, res_items.data(), 4 * res_items.size(), c->st))) return rc;
        st.h2d_bytes += 4 * (int64_t)res_items.size();
        trace_mark(c->st, "host: order/ws + uploads");
        CU(cudaEventRecord(c->ev[5], c->st));
        if (!res_launches.empty()) {
            // resident K2: one launch per (shared-memory budget, CTA width), largest budget first, on side
            // streams; each a persistent grid pulling its GHMs heaviest first
            CU(cudaEventRecord(c->ev_fork, c->st));
            int* rq = c->d_counter.as<int>() + 4 + 1024;   // queue heads (zeroed above)
            // IADH_PASS_TIMING=1 (measurement aid): the launches run one after the other on the main stream, timed
            static const bool res_timing = getenv("IADH_PASS_TIMING") != nullptr;
            std::vector<cudaEvent_t> rev;
            if (res_timing) {
                CU(c->d_resdbg.ensure(8 * 16 * res_launches.size() + 16));
                CU(cudaMemsetAsync(c->d_resdbg.p, 0, 8 * 16 * res_launches.size(), c->st));
            }
            for (size_t k = 0; k < res_launches.size(); k++) {
                const ResLaunch& L = res_launches[k];
                cudaStream_t s = res_timing ? c->st : c->st_res[k % 8];
                if (res_timing) {
                    cudaEvent_t e;
                    cudaEventCreate(&e);
                    cudaEventRecord(e, c->st);
                    rev.push_back(e);
                }
                if (k < 8 && !res_timing) CU(cudaStreamWaitEvent(s, c->ev_fork, 0));
                const int per_sm = std::max(1, resident_ctas_per_sm(L.nt, L.smem));
                DevParams rprm = c->prm;
                if (L.nt == 32 && !getenv("IADH_COOP"))   // one warp: always the all-lane merge / insert / statistics paths
                    for (int q = 0; q < 4; q++) rprm.coop[q] = 0;
                CU(launch_resident(c->d_tasks.as<GhmTask>(), c->d_items_res.as<int>() + L.begin, L.count, rq + (k % 60),
                                   c->d_dots.as<uint32_t>(), c->d_ws.as<uint8_t>(), rprm, pool_factor, L.smem, c->ds, pf, hl,
                                   std::min(L.count, c->sm_count * per_sm), L.nt, s,
                                   res_timing ? c->d_resdbg.as<unsigned long long>() + 16 * k : nullptr));
                st.kernel_launches += 1;
            }
            for (size_t k = 0; k < std::min<size_t>(8, res_launches.size()) && !res_timing; k++) {
                CU(cudaEventRecord(c->ev_res[k], c->st_res[k]));
                CU(cudaStreamWaitEvent(c->st, c->ev_res[k], 0));
            }
            if (res_timing) {
                cudaEvent_t e;
                cudaEventCreate(&e);
                cudaEventRecord(e, c->st);
                rev.push_back(e);
                cudaEventSynchronize(e);
                for (size_t k = 0; k < res_launches.size(); k++) {
                    float ms = 0;
                    cudaEventElapsedTime(&ms, rev[k], rev[k + 1]);
                    const ResLaunch& L = res_launches[k];
                    int64_t dsum = 0;
                    for (int q = 0; q < L.count; q++) dsum += (int64_t)tasks[res_items[L.begin + q]].cnt[0] + tasks[res_items[L.begin + q]].cnt[1];
                    fprintf(stderr, "[iadh] resident launch %zu: %d threads, %d KB, %d CTAs/SM, %d GHMs, %lld dots (heaviest %u): %.3f ms\n", k, L.nt,
                            L.smem >> 10, resident_ctas_per_sm(L.nt, L.smem), L.count, (long long)dsum,
                            tasks[res_items[L.begin]].cnt[0] + tasks[res_items[L.begin]].cnt[1], ms);
                    unsigned long long d[16];
                    cudaMemcpy(d, c->d_resdbg.as<unsigned long long>() + 16 * k, sizeof d, cudaMemcpyDeviceToHost);
                    fprintf(stderr, "[iadh]   Mcycles: stage %.1f seed %.1f enrich %.1f join %.1f enrich2 %.1f filter %.1f meta %.1f pack %.1f | gc %llu (%.1f Mcyc) realloc-copied %llu merges %llu inserts %llu seeds %llu stat-sum entries %llu failed %llu\n",
                            d[0] / 1e6, d[1] / 1e6, d[2] / 1e6, d[3] / 1e6, d[4] / 1e6, d[5] / 1e6, d[6] / 1e6, d[7] / 1e6, d[8], d[9] / 1e6, d[10],
                            d[11], d[12], d[13], d[14], d[15]);
                }
                for (cudaEvent_t ev : rev) cudaEventDestroy(ev);
            }
            trace_mark(c->st, "K2 resident");
        }
        {
            // size classes (order is sorted by dots, descending), launched concurrently: heavy GHMs get
            // wide CTAs and start first, the bulk runs on narrow CTAs (many GHMs in flight per SM --
            // the clustering of one GHM is a dependent chain that keeps about one warp busy)
            int* ctr = c->d_counter.as<int>();
            const int* ord = c->d_items.as<int>();
            if (pass_mode) {
                // single-warp CTAs: always the cooperative merge / insert / statistics paths (one code path,
                // all 32 lanes busy) unless IADH_COOP says otherwise
                DevParams pprm = c->prm;
                if (!getenv("IADH_COOP"))
                    for (int k = 0; k < 4; k++) pprm.coop[k] = 0;
                const int n_chunks = (int)chunk_k.size() - 1;
                CU(c->d_istate.ensure(8 * (size_t)n + 16));   // packed state word per (GHM, orientation) item (pop_batch)
                CU(cudaMemsetAsync(c->d_istate.p, 0, 8 * (size_t)n, c->st));
                CU(c->d_passctr.ensure(4 * 36 * (size_t)n_chunks + 80));
                CU(cudaMemsetAsync(c->d_passctr.p, 0, 4 * 36 * (size_t)n_chunks + 80, c->st));
                // IADH_PASS_TIMING=1 (measurement aid): device time of every pass / stage launch
                static const bool pass_timing = getenv("IADH_PASS_TIMING") != nullptr;
                std::vector<cudaEvent_t> pev;
                auto pmark = [&]() {
                    if (!pass_timing) return;
                    cudaEvent_t e;
                    cudaEventCreate(&e);
                    cudaEventRecord(e, c->st);
                    pev.push_back(e);
                };
                pmark();
                for (int ch = 0; ch < n_chunks; ch++) {
                    int* pc = c->d_passctr.as<int>() + 36 * ch;
                    const int* it1 = c->d_items1.as<int>() + chunk_i1[ch];
                    const int n1 = chunk_i1[ch + 1] - chunk_i1[ch];
                    // Launch configuration by batch size (measured on shares of config 4, K2 ms at 250 k / 63 k / 31 k /
                    // 16 k items: 32 CTAs per SM x 64 registers 60.7 / 23.3 / 18.5 / 15.6, 24 x 85 registers 67.5 / 22.0 / 17.0 /
                    // 14.4, 16 x 128 registers 81.1 / 23.5 / 16.2 / 13.6): a big batch is throughput-bound and wants every
                    // warp slot, a small one -- one rank's share at 4-8 GPUs -- is bound by its heaviest item in each of
                    // the 34 launches and wants that item's warp to run without register spills.
                    if (!c->pass_cfg_user) {
                        const int per_sm_items = n1 / std::max(1, c->sm_count);
                        if (per_sm_items >= 640) c->pass_nt = 3232, c->pass_per_sm = 32;
                        else if (per_sm_items >= 320) c->pass_nt = 3224, c->pass_per_sm = 24;
                        else c->pass_nt = 32, c->pass_per_sm = 16;
                    }
                    const int grid = std::min(n1, c->sm_count * c->pass_per_sm);
                    for (int gi = 0; gi < c->prm.n_gaps && gi < 10 && c->prm.gaps[gi] > 0 && n1 > 0; gi++)
                        for (int pass = 0; pass < 3; pass++) {
                            launch_phase1_pass(c->d_tasks.as<GhmTask>(), it1, n1, pc + gi * 3 + pass, c->d_dots.as<uint32_t>(),
                                               c->d_ws.as<uint8_t>(), pprm, pool_factor, ctr + 4, pass, gi, grid, c->pass_nt, c->st,
                                               g_trace ? (unsigned long long*)(c->d_passctr.as<int>() + 36 * n_chunks) : nullptr,
                                               c->d_istate.as<uint32_t>());
                            st.kernel_launches += 1;
                            pmark();
                        }
                    CU(cudaGetLastError());
                    const int* it2 = ord + chunk_k[ch];
                    const int n2 = chunk_k[ch + 1] - chunk_k[ch];
                    const int grid2 = std::min(n2, c->sm_count * c->pass_per_sm);
                    for (int stage = 0; stage < 4; stage++) {
                        launch_phase2_stage(c->d_tasks.as<GhmTask>(), it2, n2, pc + 30 + stage, c->d_dots.as<uint32_t>(),
                                            c->d_ws.as<uint8_t>(), pprm, pool_factor, ctr + 4, c->ds, pf, hl, stage, grid2,
                                            c->pass_nt, c->st, c->d_istate.as<uint32_t>());
                        st.kernel_launches += 1;
                        pmark();
                    }
                    CU(cudaGetLastError());
                }
                if (pass_timing && !pev.empty()) {
                    cudaEventSynchronize(pev.back());
                    fprintf(stderr, "[iadh] pass timing (ms), %d items / %d GHMs:", (int)items1.size(), n_act);
                    for (size_t k = 1; k < pev.size(); k++) {
                        float ms = 0;
                        cudaEventElapsedTime(&ms, pev[k - 1], pev[k]);
                        fprintf(stderr, "%s%.2f", ((k - 1) % 3 == 0) ? " | " : " ", ms);
                    }
                    fprintf(stderr, "\n");
                    for (cudaEvent_t e : pev) cudaEventDestroy(e);
                }
                trace_mark(c->st, "K2 pass-per-launch (all chunks)");
                if (g_trace) {
                    unsigned long long d[8];
                    cudaMemcpy(d, c->d_passctr.as<int>() + 36 * n_chunks, sizeof d, cudaMemcpyDeviceToHost);
                    fprintf(stderr, "[iadh]   join passes, CTA-cycles of the batch: init %.1fM reduce %.1fM locate %.1fM merge %.1fM erase %.1fM refresh %.1fM redo %.1fM, %llu merges\n",
                            d[0] / 1e6, d[1] / 1e6, d[2] / 1e6, d[3] / 1e6, d[4] / 1e6, d[5] / 1e6, d[6] / 1e6, d[7]);
                }
            }
            CU(cudaEventRecord(c->ev_fork, c->st));
            int begin = 0, launched = 0;
            for (int k = 0; k < c->n_classes && !pass_mode; k++) {
                // classes are cut by the GHM's total dots, so both orientations of a GHM share a class
                int end = begin;
                while (end < n_items) {
                    const GhmTask& t = tasks[items[end] >> 2];
                    if ((int64_t)t.cnt[0] + t.cnt[1] >= c->cls[k].min_dots || k == c->n_classes - 1) end++;
                    else break;
                }
                const int cnt = end - begin;
                if (cnt > 0) {
                    cudaStream_t s = (launched == 0) ? c->st : c->st_side[launched - 1];
                    if (launched) CU(cudaStreamWaitEvent(s, c->ev_fork, 0));
                    launch_cluster(c->d_tasks.as<GhmTask>(), c->d_prof.as<GhmProf>(), ord + begin, cnt, ctr + k, c->d_dots.as<uint32_t>(),
                                   c->d_ws.as<uint8_t>(), c->prm, pool_factor, ctr + 4, c->ds, pf, hl,
                                   std::min(cnt, c->sm_count * c->cls[k].per_sm),
                                   c->cls[k].nt, s);
                    if (launched) {
                        CU(cudaEventRecord(c->ev_join[launched - 1], s));
                        CU(cudaStreamWaitEvent(c->st, c->ev_join[launched - 1], 0));
                    }
                    launched++;
                }
                begin = end;
            }
            st.kernel_launches += launched;
        }
        CU(cudaEventRecord(c->ev[6], c->st));
        trace_mark(c->st, "K2 cluster");
        // (K3's measuring pass ran as the tail of every GHM inside K2)
        CU(c->d_outs.ensure(sizeof(TaskOut) * (size_t)n));
        CU(c->h_misc.ensure(sizeof(TaskOut) * (size_t)n));
        gather_out_kernel<<<(n + 255) / 256, 256, 0, c->st>>>(c->d_tasks.as<GhmTask>(), n, c->d_outs.as<TaskOut>());
        CU(readback(c->h_misc.p, c->d_outs.p, sizeof(TaskOut) * (size_t)n, c->st));
        CU(cudaStreamSynchronize(c->st));
        trace_mark(c->st, "outcomes D2H");
        st.d2h_bytes += (int64_t)sizeof(TaskOut) * n;
        st.kernel_launches += 1;
        bool any_failed = false;
        {
            const TaskOut* ho = c->h_misc.as<TaskOut>();
            for (int i = 0; i < n; i++) {
                tasks[i].out_bytes = ho[i].out_bytes;
                tasks[i].status = ho[i].status;
                tasks[i].n_mult = ho[i].n_mult;
                tasks[i].n_clusters = ho[i].n_clusters;
                tasks[i].n_aps = ho[i].n_aps;
                any_failed = any_failed || ho[i].status != 0;
            }
        }
        if (g_trace) {
            std::vector<GhmProf> pr((size_t)n);
            cudaMemcpy(pr.data(), c->d_prof.p, sizeof(GhmProf) * (size_t)n, cudaMemcpyDeviceToHost);
            long long tot[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            int hv = order.empty() ? -1 : order[0];
            for (int i : order) {
                for (int k = 0; k < 8; k++) tot[k] += pr[i].prof[k];
                if (pr[i].prof[5] > pr[hv].prof[5]) hv = i;
            }
            fprintf(stderr, "[iadh] K2 cycles (all GHMs): seed %.2fM enrich %.2fM join %.2fM filter %.2fM meta %.2fM total %.2fM\n",
                    tot[0] / 1e6, tot[1] / 1e6, tot[2] / 1e6, tot[3] / 1e6, tot[4] / 1e6, tot[5] / 1e6);
            if (hv >= 0)
                fprintf(stderr, "[iadh] slowest GHM %d (D=%u+%u, %d clusters, %d mult): seed %.2fM enrich %.2fM join %.2fM filter %.2fM meta %.2fM total %.2fM\n",
                        hv, tasks[hv].cnt[0], tasks[hv].cnt[1], tasks[hv].n_clusters, tasks[hv].n_mult, pr[hv].prof[0] / 1e6,
                        pr[hv].prof[1] / 1e6, pr[hv].prof[2] / 1e6, pr[hv].prof[3] / 1e6, pr[hv].prof[4] / 1e6, pr[hv].prof[5] / 1e6);
            if (hv >= 0)
                fprintf(stderr, "[iadh]   join of it: init %.2fM reduce %.2fM locate %.2fM merge %.2fM erase %.2fM refresh %.2fM redo %.2fM, %lld merges\n",
                        pr[hv].prof2[0] / 1e6, pr[hv].prof2[1] / 1e6, pr[hv].prof2[2] / 1e6, pr[hv].prof2[3] / 1e6,
                        pr[hv].prof2[4] / 1e6, pr[hv].prof2[5] / 1e6, pr[hv].prof2[6] / 1e6, pr[hv].prof2[7]);
        }
        // resident GHMs that outgrew a shared-memory budget are re-run on the global path (fused kernel, wide
        // CTAs), the others keep their results
        {
            std::vector<int> fb;
            bool global_failed = false;
            for (int i : order_all)
                if (any_failed && tasks[i].status != 0) {
                    if (is_res[i]) fb.push_back(i);
                    else global_failed = true;
                }
            for (int i : fb) force_global[i] = 1;
            if (!fb.empty() && !global_failed) {
                if (g_trace) fprintf(stderr, "[iadh] resident fallback: %d of %d GHMs (first: D=%u+%u status %d)\n", (int)fb.size(),
                                     (int)res_items.size(), tasks[fb[0]].cnt[0], tasks[fb[0]].cnt[1], tasks[fb[0]].status);
                std::vector<int> fitems;
                for (int i : fb) fitems.push_back(i * 4 + tasks[i].half_mask);
                if ((rc = upload(c->d_items_fb, fitems.data(), 4 * fitems.size(), c->st))) return rc;
                CU(cudaMemsetAsync(c->d_counter.p, 0, 16 + 4 * 1024 + 4 * 64, c->st));
                reset_subset_kernel<<<((int)fb.size() + 127) / 128, 128, 0, c->st>>>(c->d_tasks.as<GhmTask>(), c->d_items_fb.as<int>(), (int)fb.size());
                launch_cluster(c->d_tasks.as<GhmTask>(), c->d_prof.as<GhmProf>(), c->d_items_fb.as<int>(), (int)fb.size(), c->d_counter.as<int>(),
                               c->d_dots.as<uint32_t>(), c->d_ws.as<uint8_t>(), c->prm, pool_factor, c->d_counter.as<int>() + 4, c->ds, pf, hl,
                               std::min((int)fb.size(), c->sm_count * 2), 256, c->st);
                CU(cudaEventRecord(c->ev[6], c->st));
                gather_out_kernel<<<(n + 255) / 256, 256, 0, c->st>>>(c->d_tasks.as<GhmTask>(), n, c->d_outs.as<TaskOut>());
                CU(readback(c->h_misc.p, c->d_outs.p, sizeof(TaskOut) * (size_t)n, c->st));
                CU(cudaStreamSynchronize(c->st));
                st.d2h_bytes += (int64_t)sizeof(TaskOut) * n;
                st.h2d_bytes += 4 * (int64_t)fitems.size();
                st.kernel_launches += 3;
                const TaskOut* ho = c->h_misc.as<TaskOut>();
                for (int i : fb) {
                    tasks[i].out_bytes = ho[i].out_bytes;
                    tasks[i].status = ho[i].status;
                    tasks[i].n_mult = ho[i].n_mult;
                    tasks[i].n_clusters = ho[i].n_clusters;
                    tasks[i].n_aps = ho[i].n_aps;
                }
                st_fallbacks += (int64_t)fb.size();
            }
        }
        bool retry = false;
        for (int i : order_all)
            if (!any_failed) break;
            else if (tasks[i].status != 0) {
                if (g_trace && !retry)
                    fprintf(stderr, "[iadh] retry: GHM %d (D=%u+%u) status=%d clusters=%d pool_factor=%d\n", i, tasks[i].cnt[0],
                            tasks[i].cnt[1], tasks[i].status, tasks[i].n_clusters, pool_factor);
                retry = true;
            }
        if (!retry) break;
        if (attempt >= 3) return fail(IADH_ERR_CAPACITY, "anchor-point pool exhausted even at factor %d", pool_factor);
        pool_factor *= 4;
    }

    trace_mark(c->st, "tasks D2H");
    // homolog-count filter (GHMProfile.cpp:129-132) on the clustered GHMs: dropped multiplicons leave the output lists,
    // the sizes the measuring pass left in the task records are recomputed and read back once more
    c->hf_dropped = 0;
    if (hl && c->hf_active && n > 0) {
        launch_homolog_filter(c->d_tasks.as<GhmTask>(), n, c->d_ws.as<uint8_t>(), c->ds, c->n_fam, pf, c->d_hdev.as<HomologDev>(), pool_factor,
                              c->prm.anchor_points, c->sm_count, c->st);
        gather_out_kernel<<<(n + 255) / 256, 256, 0, c->st>>>(c->d_tasks.as<GhmTask>(), n, c->d_outs.as<TaskOut>());
        CU(readback(c->h_misc.p, c->d_outs.p, sizeof(TaskOut) * (size_t)n, c->st));
        CU(cudaStreamSynchronize(c->st));
        st.d2h_bytes += (int64_t)sizeof(TaskOut) * n;
        st.kernel_launches += 3;
        const TaskOut* ho = c->h_misc.as<TaskOut>();
        for (int i = 0; i < n; i++) {
            c->hf_dropped += tasks[i].n_mult - ho[i].n_mult;
            tasks[i].out_bytes = ho[i].out_bytes;
            tasks[i].n_mult = ho[i].n_mult;
            tasks[i].n_clusters = ho[i].n_clusters;
            tasks[i].n_aps = ho[i].n_aps;
        }
        trace_mark(c->st, "homolog filter");
    }
    // pack: sizes -> offsets -> write
    std::vector<int> sel;
    std::vector<int64_t> out_off;
    int64_t total_bytes = 0;
    st.n_ghm = n;
    // Groups: the device stream is the host stream minus its first 4 bytes, so group 0's count is written by the host
    // (make_result / the caller) and every later group starts with a 4-byte count written on the device.
    const int n_groups = out.group_begin ? (int)out.group_begin->size() - 1 : 1;
    std::vector<int64_t> ghdr;      // per later group: device offset of its header, and its count (pairs)
    out.group_off.assign((size_t)n_groups + 1, 0);
    int gk = 0;
    int32_t gcount = 0;
    auto next_group = [&]() {
        if (gk > 0) ghdr[2 * (size_t)(gk - 1) + 1] = gcount;
        gk++;
        gcount = 0;
        out.group_off[gk] = (size_t)total_bytes + 4;
        ghdr.push_back(total_bytes);
        ghdr.push_back(0);
        total_bytes += 4;
    };
    for (int i = 0; i < n; i++) {   // one pass: what is packed and where, multiplicon -> GHM index, totals
        while (gk + 1 < n_groups && i >= (*out.group_begin)[gk + 1]) next_group();
        const GhmTask& t = tasks[i];
        if (t.n_mult <= 0) continue;
        sel.push_back(i);
        out_off.push_back(total_bytes);
        total_bytes += t.out_bytes;
        for (int k = 0; k < t.n_mult; k++) out.ghm_index.push_back(t.seq);
        st.n_multiplicons += t.n_mult;
        gcount += t.n_mult;
        st.n_baseclusters += t.n_clusters;
        st.n_anchorpoints += t.n_aps;
    }
    while (gk + 1 < n_groups) next_group();
    if (gk > 0) ghdr[2 * (size_t)(gk - 1) + 1] = gcount;
    out.group_off[n_groups] = (size_t)total_bytes + 4;
    trace_mark(c->st, "host: sizes -> offsets");
    CU(out.packed.ensure((size_t)total_bytes + 4));
    trace_mark(c->st, "host: pinned result buffer");
    out.packed_bytes = (size_t)total_bytes + 4;
    if (total_bytes > 0) {
        if (c->copy_pending) {   // the previous result may still be travelling out of d_out
            if ((size_t)total_bytes + 16 > c->d_out.cap) CU(cudaEventSynchronize(c->ev_copied));   // (ensure() would hand the block on)
            else CU(cudaStreamWaitEvent(c->st, c->ev_copied, 0));
        }
        CU(c->d_out.ensure((size_t)total_bytes + 16));
        if (!sel.empty()) {
            if ((rc = upload(c->d_sel, sel.data(), 4 * sel.size(), c->st))) return rc;
            if ((rc = upload(c->d_out_off, out_off.data(), 8 * out_off.size(), c->st))) return rc;
            launch_pack_write(hl, c->d_tasks.as<GhmTask>(), c->d_sel.as<int>(), c->d_out_off.as<int64_t>(), (int)sel.size(),
                              c->d_ws.as<uint8_t>(), c->ds, pf, pool_factor, c->d_out.as<uint8_t>(), c->st);
            st.kernel_launches += 1;
        }
        if (!ghdr.empty()) {
            if ((rc = upload(c->d_ghdr, ghdr.data(), 8 * ghdr.size(), c->st))) return rc;
            const int nh = (int)(ghdr.size() / 2);
            group_header_kernel<<<(nh + 127) / 128, 128, 0, c->st>>>(c->d_out.as<uint8_t>(), c->d_ghdr.as<int64_t>(), nh);
            st.kernel_launches += 1;
            st.h2d_bytes += 8 * (int64_t)ghdr.size();
        }
        trace_mark(c->st, "K3 write");
        // the copy to the host runs on the copy stream: the caller's next call computes meanwhile
        CU(cudaEventRecord(c->ev_packed, c->st));
        CU(cudaStreamWaitEvent(c->st_copy, c->ev_packed, 0));
        CU(copy_to_host_in_pieces(out.packed.as<uint8_t>() + 4, c->d_out.p, (size_t)total_bytes, c->st_copy));
        CU(cudaEventRecord(c->ev_copied, c->st_copy));
        CU(cudaEventCreateWithFlags(&out.ready, cudaEventDisableTiming));
        CU(cudaEventRecord(out.ready, c->st_copy));
        c->copy_pending = true;
        st.h2d_bytes += 12 * (int64_t)sel.size();
        st.d2h_bytes += total_bytes;
    }
    if (g_trace) {
        out.wait();
        trace_mark(c->st, "result D2H");
    }
    out.tasks.swap(tasks);   // (the caller's vector is not used after the call)
    trace_mark(c->st, "host: collect");
    return IADH_OK;
}

static iadh_result* make_result(BatchOut& b)
{
    iadh_result* r = new iadh_result();
    const int32_t total = (int32_t)b.st.n_multiplicons;
    memcpy(b.packed.p, &total, 4);
    std::swap(r->bytes, b.packed);
    r->ready = b.ready;
    b.ready = nullptr;
    r->n_bytes = b.packed_bytes;
    r->ghm_index.swap(b.ghm_index);
    r->stats = b.st;
    return r;
}

static void init_task(GhmTask& t)
{
    memset(&t, 0, sizeof t);
}

// timing wrapper: total / build / cluster device times from CUDA events on the engine's stream
struct Timer {
    cudaEvent_t a, b;
};

static int run_l2(iadh_ctx* c, const int32_t* xs, const int32_t* ys, int64_t n, iadh_result** out)
{
    if (n < 0 || n > 0x3fffffff) return fail(IADH_ERR_ARG, "bad pair count");
    if (c->pairs)
        for (int m : c->masked_cnt)
            if (m)
                return fail(IADH_ERR_STATE, "level-2 call in pairs mode while elements are masked: the reference's pairs-mode "
                                          "build skips masked elements (ListElement.cpp:25), this one does not");
    std::vector<GhmTask> tasks((size_t)n);
    for (int64_t i = 0; i < n; i++) {
        GhmTask& t = tasks[i];
        init_task(t);
        if (xs[i] < 0 || xs[i] >= c->n_lists || ys[i] < 0 || ys[i] >= c->n_lists)
            return fail(IADH_ERR_ARG, "pair %lld names list %d/%d of %d", (long long)i, xs[i], ys[i], c->n_lists);
        t.x_list = xs[i];
        t.y_list = ys[i];
        t.x_object_id = xs[i];
        t.y_object_id = ys[i];
        t.level = 2;
        // GHM::prepareForStatisticalValidation (GHM.cpp:221-236): 32-bit unsigned arithmetic
        const unsigned nx = (unsigned)(c->list_off[xs[i] + 1] - c->list_off[xs[i]]);
        const unsigned ny = (unsigned)(c->list_off[ys[i] + 1] - c->list_off[ys[i]]);
        t.area = (xs[i] == ys[i]) ? (unsigned)(nx * (nx - 1) / 2) : (unsigned)(nx * ny);
        t.seq = (int32_t)i;
    }
    cudaEvent_t e0 = c->ev[7], e1 = c->ev_call;
    CU(cudaEventRecord(e0, c->st));
    BatchOut b;
    memset(&b.st, 0, sizeof b.st);
    ProfileDev nopf;
    memset(&nopf, 0, sizeof nopf);
    int rc = run_batch(c, false, nopf, tasks, b);
    double acc_build = 0, acc_cluster = 0;
    bool was_split = false;
    if (rc == RC_SPLIT) {
        // bisect until every part fits the workspace budget; parts are consecutive ranges of the pair
        // sequence, so their packed streams simply follow each other
        was_split = true;
        std::vector<std::pair<int64_t, int64_t>> todo(1, std::make_pair((int64_t)0, n));
        std::vector<uint8_t> bytes;
        memset(&b.st, 0, sizeof b.st);
        b.ghm_index.clear();
        rc = IADH_OK;
        while (!todo.empty() && rc == IADH_OK) {
            const std::pair<int64_t, int64_t> r = todo.back();
            todo.pop_back();
            std::vector<GhmTask> part(tasks.begin() + r.first, tasks.begin() + r.second);
            for (size_t i = 0; i < part.size(); i++) {
                const GhmTask keep = tasks[r.first + i];
                init_task(part[i]);
                part[i].x_list = keep.x_list; part[i].y_list = keep.y_list;
                part[i].x_object_id = keep.x_object_id; part[i].y_object_id = keep.y_object_id;
                part[i].level = keep.level; part[i].area = keep.area;
                part[i].seq = (int32_t)i;
            }
            BatchOut pb;
            memset(&pb.st, 0, sizeof pb.st);
            rc = run_batch(c, false, nopf, part, pb);
            if (rc == RC_SPLIT) {
                const int64_t mid = (r.first + r.second) / 2;
                todo.push_back(std::make_pair(mid, r.second));   // (stack: the first half is done first)
                todo.push_back(std::make_pair(r.first, mid));
                rc = IADH_OK;
                continue;
            }
            if (rc != IADH_OK) break;
            cudaStreamSynchronize(c->st);
            float m1 = 0, m2 = 0, m3 = 0;
            cudaEventElapsedTime(&m1, c->ev[1], c->ev[2]);
            cudaEventElapsedTime(&m2, c->ev[3], c->ev[4]);
            cudaEventElapsedTime(&m3, c->ev[5], c->ev[6]);
            acc_build += m1 + m2;
            acc_cluster += m3;
            pb.wait();
            bytes.insert(bytes.end(), pb.packed.as<uint8_t>() + 4, pb.packed.as<uint8_t>() + pb.packed_bytes);
            for (int32_t gidx : pb.ghm_index) b.ghm_index.push_back(gidx + (int32_t)r.first);
            b.st.n_ghm += pb.st.n_ghm; b.st.n_dots += pb.st.n_dots; b.st.n_multiplicons += pb.st.n_multiplicons;
            b.st.n_baseclusters += pb.st.n_baseclusters; b.st.n_anchorpoints += pb.st.n_anchorpoints;
            b.st.kernel_launches += pb.st.kernel_launches; b.st.h2d_bytes += pb.st.h2d_bytes; b.st.d2h_bytes += pb.st.d2h_bytes;
        }
        if (rc == IADH_OK) {
            if (b.packed.ensure(bytes.size() + 4) != cudaSuccess) rc = fail(IADH_ERR_CUDA, "out of pinned host memory");
            else {
                if (!bytes.empty()) memcpy(b.packed.as<uint8_t>() + 4, bytes.data(), bytes.size());
                b.packed_bytes = bytes.size() + 4;
            }
        }
    }
    cudaEventRecord(e1, c->st);
    cudaEventSynchronize(e1);
    if (rc == IADH_OK) {
        float ms = 0, m1 = 0, m2 = 0, m3 = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        b.st.total_ms = ms;
        if (was_split) {
            b.st.build_ms = acc_build;
            b.st.cluster_ms = acc_cluster;
        } else if (n > 0) {
            cudaEventElapsedTime(&m1, c->ev[1], c->ev[2]);
            cudaEventElapsedTime(&m2, c->ev[3], c->ev[4]);
            b.st.build_ms = m1 + m2;
            cudaEventElapsedTime(&m3, c->ev[5], c->ev[6]);
            b.st.cluster_ms = m3;
        }
        *out = make_result(b);
        set_last_result(c, was_split ? nullptr : *out, (*out)->n_bytes - 4);
    }
    return rc;
}

__global__ void omxn_kernel(const double* x, const int* n, long long count, double* out, int mode)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) out[i] = mode ? cm::pow_ref(x[i], n[i]) : cm::omxn(x[i], n[i]);
}

// No exception crosses the C boundary: the host vectors of a call can throw (bad_alloc, length_error)
#define IADH_TRY try {
#define IADH_CATCH                                                                                   \
    }                                                                                                \
    catch (const std::bad_alloc&) { return fail(IADH_ERR_CAPACITY, "out of host memory"); }          \
    catch (const std::exception& e_) { return fail(IADH_ERR_ARG, "exception: %s", e_.what()); }

// ---------------------------------------------------------------------------------------------
extern "C" {

const char* iadh_last_error(void) { return g_err.c_str(); }
const char* iadh_version(void) { return "iadhore_b200 0.1 (sm_100a)"; }

static int create_impl(const iadh_dataset_desc* d, const int64_t* pair_ptr, const int32_t* pair_ids, int32_t n_genes,
                       int device, iadh_ctx** out);
void iadh_destroy(iadh_ctx* ctx);

int iadh_create(const iadh_dataset_desc* d, int device, iadh_ctx** out)
{
    if (!d || (d->n_lists > 0 && !d->fam)) return fail(IADH_ERR_ARG, "iadh_create: bad dataset descriptor");
    IADH_TRY return create_impl(d, nullptr, nullptr, 0, device, out);
    IADH_CATCH
}

int iadh_create_pairs(const iadh_dataset_desc* d, const int64_t* pair_ptr, const int32_t* pair_ids, int32_t n_genes,
                      int device, iadh_ctx** out)
{
    if (!pair_ptr || n_genes < 0) return fail(IADH_ERR_ARG, "iadh_create_pairs: bad pair table");
    for (int32_t g = 0; g < n_genes; g++)
        if (pair_ptr[g + 1] < pair_ptr[g]) return fail(IADH_ERR_ARG, "iadh_create_pairs: pair_ptr is not ascending at gene %d", g);
    if (pair_ptr[0] != 0 || (pair_ptr[n_genes] > 0 && !pair_ids)) return fail(IADH_ERR_ARG, "iadh_create_pairs: bad pair table");
    for (int64_t k = 0; k < pair_ptr[n_genes]; k++)
        if (pair_ids[k] < 0 || pair_ids[k] >= n_genes)
            return fail(IADH_ERR_ARG, "iadh_create_pairs: partner id %d out of range (%d genes)", pair_ids[k], n_genes);
    IADH_TRY return create_impl(d, pair_ptr, pair_ids, n_genes, device, out);
    IADH_CATCH
}

}  // extern "C"

static int create_impl(const iadh_dataset_desc* d, const int64_t* pair_ptr, const int32_t* pair_ids, int32_t n_genes,
                       int device, iadh_ctx** out)
{
    const bool pairs = pair_ptr != nullptr;
    if (!d || !out || d->n_lists < 0 || (d->n_lists > 0 && (!d->list_off || !d->gid || !d->num_id || !d->flags)))
        return fail(IADH_ERR_ARG, "iadh_create: bad dataset descriptor");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(IADH_ERR_CUDA, "no CUDA device available (%s); this library has no CPU fallback",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(IADH_ERR_ARG, "device %d out of range (%d devices)", device, ndev);
    CU(cudaSetDevice(device));
    iadh_ctx* c = new iadh_ctx();
    // every early return below releases what was built so far (streams, events, device blocks)
    struct Guard {
        iadh_ctx* c;
        ~Guard()
        {
            if (c) iadh_destroy(c);
        }
    } guard{c};
    c->device = device;
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    int prio_lo = 0, prio_hi = 0;   // (numerically: lowest priority = largest value)
    CU(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    CU(cudaStreamCreateWithPriority(&c->st, cudaStreamNonBlocking, prio_hi));
    for (int i = 0; i < 8; i++) CU(cudaEventCreate(&c->ev[i]));
    CU(cudaEventCreate(&c->ev_call));
    CU(cudaStreamCreateWithPriority(&c->st_copy, cudaStreamNonBlocking, prio_lo));
    CU(cudaEventCreateWithFlags(&c->ev_packed, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_copied, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_copied_recv, cudaEventDisableTiming));
    for (int i = 0; i < 4; i++) CU(cudaEventCreate(&c->user_ev[i]));
    for (int i = 0; i < 3; i++) {
        CU(cudaStreamCreateWithFlags(&c->st_side[i], cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&c->ev_join[i], cudaEventDisableTiming));
    }
    CU(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    for (int i = 0; i < 8; i++) {
        CU(cudaStreamCreateWithFlags(&c->st_res[i], cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&c->ev_res[i], cudaEventDisableTiming));
    }
    c->n_lists = d->n_lists;
    c->list_off.assign(d->list_off, d->list_off + d->n_lists + 1);
    const int64_t total = d->n_lists ? c->list_off[d->n_lists] : 0;
    if (total >= 0x7fffffff) { return fail(IADH_ERR_CAPACITY, "more than 2^31 elements"); }
    c->pairs = pairs;
    // the bucket key of an element: its family, or in pairs mode its gene id
    const int32_t* key = pairs ? d->gid : d->fam;
    int max_fam = -1;
    for (int l = 0; l < d->n_lists; l++) {
        const int64_t sz = c->list_off[l + 1] - c->list_off[l];
        if (sz < 0) { return fail(IADH_ERR_ARG, "list %d has negative size", l); }
        if (sz > IADH_MAX_COORD) {
            return fail(IADH_ERR_CAPACITY,
                        "list %d has %lld remapped elements (> %d): the reference's 32-bit x*y sums overflow there "
                        "(BaseCluster.cpp:156-158), parity is undefined",
                        l, (long long)sz, IADH_MAX_COORD);
        }
    }
    for (int64_t i = 0; i < total; i++) {
        if (key[i] < 0) { return fail(IADH_ERR_ARG, "element %lld has a negative %s id", (long long)i, pairs ? "gene" : "family"); }
        if (pairs && key[i] >= n_genes) { return fail(IADH_ERR_ARG, "element %lld: gene id %d outside the pair table (%d genes)", (long long)i, key[i], n_genes); }
        max_fam = std::max(max_fam, key[i]);
    }
    c->n_fam = pairs ? n_genes : max_fam + 1;
    c->flags.resize((size_t)total);
    c->masked_cnt.assign(d->n_lists, 0);
    for (int l = 0; l < d->n_lists; l++)
        for (int64_t i = c->list_off[l]; i < c->list_off[l + 1]; i++) {
            uint8_t f = (d->flags[i] & IADH_F_STRAND) ? 1 : 0;
            if (d->flags[i] & IADH_F_MASKED) {
                f |= 4;
                c->masked_cnt[l]++;
            }
            c->flags[i] = f;
        }
    // (IADH_NO_FAM_TABLE: test knob for the binary-search fallback used when the table would be too large)
    const int64_t tbl = (int64_t)c->n_fam * (d->n_lists + 1);
    const bool have_table = tbl > 0 && tbl <= ((int64_t)128 << 20) && !getenv("IADH_NO_FAM_TABLE");
    int rc = IADH_OK;
    // Family mode: the join tables (family CSR, member records, strand prefix, own slots, repeated-ID flags) are built
    // on the device from the flat arrays (launch_tables, build_kernels.cu); IADH_HOST_TABLES=1 keeps the host
    // construction below (test knob; pairs mode always: its buckets come from the pair table).
    const bool device_tables = !pairs && !getenv("IADH_HOST_TABLES");
    if (device_tables) {
        int max_gid = -1;
        for (int64_t i = 0; i < total; i++) max_gid = std::max(max_gid, d->gid[i]);
        const size_t nf = (size_t)c->n_fam + 1, ng = (size_t)max_gid + 2, nb = (size_t)(std::max<int64_t>(total, (int64_t)nf) / 1024 + 4);
        DBuf t_cnt, t_tmp;   // zeroed counters [fam_cnt | cursor | gid_cnt], then scratch [tmp | block sums]
        if (!rc) rc = upload(c->d_list_off, c->list_off.data(), 8 * c->list_off.size(), c->st);
        if (!rc) rc = upload(c->d_fam, key, 4 * (size_t)total, c->st);
        if (!rc) rc = upload(c->d_gid, d->gid, 4 * (size_t)total, c->st);
        if (!rc) rc = upload(c->d_num_id, d->num_id, 4 * (size_t)total, c->st);
        if (!rc) rc = upload(c->d_flags, c->flags.data(), (size_t)total, c->st);
        if (!rc && (t_cnt.ensure(4 * (2 * nf + ng) + 16) != cudaSuccess || t_tmp.ensure(4 * ((size_t)total + nb) + 16) != cudaSuccess ||
                    c->d_fam_ptr.ensure(4 * nf + 16) != cudaSuccess || c->d_fam_members.ensure(4 * (size_t)total + 16) != cudaSuccess ||
                    c->d_elem_list.ensure(4 * (size_t)total + 16) != cudaSuccess || c->d_mem_y.ensure(4 * (size_t)total + 16) != cudaSuccess ||
                    c->d_mem_list.ensure(4 * (size_t)total + 16) != cudaSuccess || c->d_self_slot.ensure(4 * (size_t)total + 16) != cudaSuccess ||
                    c->d_plus_prefix.ensure(4 * (size_t)total + 16) != cudaSuccess))
            rc = fail(IADH_ERR_CUDA, "out of device memory");
        if (!rc) {
            cudaMemsetAsync(t_cnt.p, 0, 4 * (2 * nf + ng), c->st);
            int32_t* fam_cnt = t_cnt.as<int32_t>();
            launch_tables(c->d_fam.as<int32_t>(), c->d_gid.as<int32_t>(), c->d_flags.as<uint8_t>(), c->d_list_off.as<int64_t>(), d->n_lists, total,
                          c->n_fam, fam_cnt, fam_cnt + 2 * nf, fam_cnt + nf, c->d_fam_ptr.as<int32_t>(), c->d_fam_members.as<int32_t>(),
                          c->d_elem_list.as<int32_t>(), c->d_mem_y.as<uint32_t>(), c->d_mem_list.as<int32_t>(), c->d_self_slot.as<int32_t>(),
                          c->d_plus_prefix.as<int32_t>(), t_tmp.as<int32_t>(), t_tmp.as<int32_t>() + total, c->sm_count, c->st);
            if (cudaGetLastError() != cudaSuccess) rc = fail(IADH_ERR_CUDA, "dataset table kernels failed to launch");
        }
        if (!rc && cudaStreamSynchronize(c->st) != cudaSuccess) rc = fail(IADH_ERR_CUDA, "dataset tables failed on the device");
        t_cnt.release();
        t_tmp.release();
    } else {
    // family CSR: members ascending by global element index
    std::vector<int32_t> fam_ptr(c->n_fam + 1, 0), members;
    if (!pairs) {
        members.resize((size_t)total);
        for (int64_t i = 0; i < total; i++) fam_ptr[key[i] + 1]++;
        for (int f = 0; f < c->n_fam; f++) fam_ptr[f + 1] += fam_ptr[f];
        std::vector<int32_t> cur(fam_ptr.begin(), fam_ptr.end() - 1);
        for (int64_t i = 0; i < total; i++) members[cur[key[i]]++] = (int32_t)i;
    } else {
        // pairs mode: bucket g = positions of the partners of gene g, i.e. every element e with
        // gid[e] in pairs(g) (Gene::isPairWith(g, e), Gene.cpp:24-27).  Built through the transposed
        // relation (rev[p] = genes that list p as a partner; equal to pairs for a symmetric table):
        // walking the elements in ascending order and appending e to the bucket of every g in
        // rev[gid[e]] leaves each bucket ascending without a sort.
        const int64_t n_pairs = pair_ptr[n_genes];
        std::vector<int64_t> rev_ptr((size_t)n_genes + 1, 0);
        std::vector<int32_t> rev((size_t)n_pairs);
        for (int64_t k = 0; k < n_pairs; k++) rev_ptr[pair_ids[k] + 1]++;
        for (int g = 0; g < n_genes; g++) rev_ptr[g + 1] += rev_ptr[g];
        {
            std::vector<int64_t> cur(rev_ptr.begin(), rev_ptr.end() - 1);
            for (int g = 0; g < n_genes; g++)
                for (int64_t k = pair_ptr[g]; k < pair_ptr[g + 1]; k++) rev[cur[pair_ids[k]]++] = g;
        }
        std::vector<int64_t> cnt((size_t)n_genes + 1, 0);
        for (int64_t i = 0; i < total; i++)
            for (int64_t k = rev_ptr[key[i]]; k < rev_ptr[key[i] + 1]; k++) cnt[rev[k] + 1]++;
        for (int g = 0; g < n_genes; g++) {
            cnt[g + 1] += cnt[g];
            if (cnt[g + 1] >= 0x7fffffff) { return fail(IADH_ERR_CAPACITY, "more than 2^31 partner positions"); }
            fam_ptr[g + 1] = (int32_t)cnt[g + 1];
        }
        members.resize((size_t)cnt[n_genes]);
        {
            std::vector<int32_t> cur(fam_ptr.begin(), fam_ptr.end() - 1);
            for (int64_t i = 0; i < total; i++)
                for (int64_t k = rev_ptr[key[i]]; k < rev_ptr[key[i] + 1]; k++) members[cur[rev[k]]++] = (int32_t)i;
        }
        for (int g = 0; g < n_genes; g++)
            if (std::adjacent_find(members.begin() + fam_ptr[g], members.begin() + fam_ptr[g + 1]) != members.begin() + fam_ptr[g + 1]) {
                return fail(IADH_ERR_ARG, "pair table lists a partner of gene %d twice", g);
            }
    }
    const int64_t n_mem = (int64_t)members.size();
    // level-2 build tables (build_kernels.cu): member records, strand prefix counts, own slot of
    // every element, duplicate-gene-ID flag, and the (family, list) -> first slot table
    std::vector<uint32_t> mem_y((size_t)n_mem);
    std::vector<int32_t> plus_prefix((size_t)n_mem + 1, 0), self_slot((size_t)total), mem_list((size_t)n_mem), elem_list((size_t)total);
    {
        int l = 0;
        for (int64_t i = 0; i < total; i++) {
            while (i >= c->list_off[l + 1]) l++;
            elem_list[i] = l;
        }
        for (int64_t m = 0; m < n_mem; m++) {
            const int32_t e = members[m];
            const int strand = c->flags[e] & 1;
            mem_y[m] = ((uint32_t)(e - c->list_off[elem_list[e]]) << 1) | (uint32_t)strand;
            mem_list[m] = elem_list[e];
            plus_prefix[m + 1] = plus_prefix[m] + strand;
            if (!pairs) self_slot[e] = (int32_t)m;
        }
        if (pairs)  // the cut `y <= x` of the self comparison (GHM.cpp:125): partners before the element
            for (int64_t i = 0; i < total; i++)
                self_slot[i] = (int32_t)(std::lower_bound(members.begin() + fam_ptr[key[i]], members.begin() + fam_ptr[key[i] + 1],
                                                          (int32_t)i) - members.begin());
        int max_gid = -1;
        for (int64_t i = 0; i < total; i++) max_gid = std::max(max_gid, d->gid[i]);
        std::vector<uint8_t> seen((size_t)max_gid + 2, 0);
        for (int64_t i = 0; i < total; i++)
            if (d->gid[i] >= 0 && seen[d->gid[i]] < 2) seen[d->gid[i]]++;
        if (!pairs)  // (pairs mode has no gene-ID test: a gene is never in its own pair set, GenePairs.cpp:25)
            for (int64_t i = 0; i < total; i++)
                if (d->gid[i] < 0 || seen[d->gid[i]] > 1) c->flags[i] |= 8;
    }
    if (!rc) rc = upload(c->d_mem_y, mem_y.data(), 4 * mem_y.size(), c->st);
    if (!rc) rc = upload(c->d_plus_prefix, plus_prefix.data(), 4 * plus_prefix.size(), c->st);
    if (!rc) rc = upload(c->d_self_slot, self_slot.data(), 4 * self_slot.size(), c->st);
    if (!rc) rc = upload(c->d_mem_list, mem_list.data(), 4 * mem_list.size(), c->st);
    if (!rc) rc = upload(c->d_elem_list, elem_list.data(), 4 * elem_list.size(), c->st);
    if (!rc) rc = upload(c->d_list_off, c->list_off.data(), 8 * c->list_off.size(), c->st);
    if (!rc) rc = upload(c->d_fam, key, 4 * (size_t)total, c->st);
    if (!rc) rc = upload(c->d_gid, d->gid, 4 * (size_t)total, c->st);
    if (!rc) rc = upload(c->d_num_id, d->num_id, 4 * (size_t)total, c->st);
    if (!rc) rc = upload(c->d_flags, c->flags.data(), (size_t)total, c->st);
    if (!rc) rc = upload(c->d_fam_ptr, fam_ptr.data(), 4 * fam_ptr.size(), c->st);
    if (!rc) rc = upload(c->d_fam_members, members.data(), 4 * members.size(), c->st);
    }
    if (!rc && have_table) {  // the (family, list) table is built on the device from the CSR just uploaded
        if (c->d_fam_list.ensure(4 * (size_t)tbl) != cudaSuccess) rc = fail(IADH_ERR_CUDA, "out of device memory");
        else
            launch_fam_list(c->d_fam_ptr.as<int32_t>(), c->d_fam_members.as<int32_t>(), c->d_list_off.as<int64_t>(), c->n_fam,
                            d->n_lists, total, c->d_fam_list.as<int32_t>(), c->st);
    }
    if (!rc && cudaStreamSynchronize(c->st) != cudaSuccess) rc = fail(IADH_ERR_CUDA, "dataset upload failed");
    if (rc) return rc;
    c->ds.n_lists = c->n_lists;
    c->ds.list_off = c->d_list_off.as<int64_t>();
    c->ds.fam = c->d_fam.as<int32_t>();
    c->ds.gid = c->d_gid.as<int32_t>();
    c->ds.num_id = c->d_num_id.as<int32_t>();
    c->ds.flags = c->d_flags.as<uint8_t>();
    c->ds.fam_ptr = c->d_fam_ptr.as<int32_t>();
    c->ds.fam_members = c->d_fam_members.as<int32_t>();
    c->ds.mem_y = c->d_mem_y.as<uint32_t>();
    c->ds.plus_prefix = c->d_plus_prefix.as<int32_t>();
    c->ds.self_slot = c->d_self_slot.as<int32_t>();
    c->ds.mem_list = c->d_mem_list.as<int32_t>();
    c->ds.elem_list = c->d_elem_list.as<int32_t>();
    c->ds.fam_list = have_table ? c->d_fam_list.as<int32_t>() : nullptr;
    c->ds.pairs = pairs ? 1 : 0;
    sort_lists(c);
    // defaults = shipped datasetI/II settings, Bonferroni (Settings.cpp:4-9)
    memset(&c->prm, 0, sizeof c->prm);
    const int g30[10] = {3, 4, 5, 6, 8, 11, 14, 18, 23, 30};
    c->prm.n_gaps = 10;
    memcpy(c->prm.gaps, g30, sizeof g30);
    c->prm.gap_size = 30;
    c->prm.cluster_gap = 35;
    c->prm.anchor_points = 3;
    c->prm.mult_hyp = IADH_MH_BONFERRONI;
    c->prm.q_value = 0.75;
    c->prm.prob_cutoff = 0.01;
    // tuning / test knobs (the anchor-point pool grows automatically when too small)
    if (const char* e = getenv("IADH_POOL_FACTOR")) c->pool_factor = std::max(1, atoi(e));
    if (const char* e = getenv("IADH_CTAS_PER_SM")) c->ctas_per_sm = std::max(1, atoi(e));
    for (int k = 0; k < 4; k++) c->prm.coop[k] = (k < 3) ? 16 : 48;  // (measured: profile searches -5 % vs 48)
    if (const char* e = getenv("IADH_COOP")) sscanf(e, "%d,%d,%d,%d", &c->prm.coop[0], &c->prm.coop[1], &c->prm.coop[2], &c->prm.coop[3]);
    if (const char* e = getenv("IADH_SPLIT_MIN")) c->split_min = std::max(3, atoi(e));
    {
        size_t fr = 0, tot = 0;
        if (cudaMemGetInfo(&fr, &tot) == cudaSuccess) c->ws_budget = (int64_t)(tot / 2);
    }
    if (const char* e = getenv("IADH_WS_BUDGET_MB")) c->ws_budget = (int64_t)std::max(1, atoi(e)) << 20;
    if (const char* e = getenv("IADH_RESIDENT")) c->resident = atoi(e);
    if (const char* e = getenv("IADH_RES_SQUEEZE")) c->prm.rs_squeeze = atoi(e);
    if (const char* e = getenv("IADH_RES_MAX_KB")) c->res_max_bytes = std::min(222, std::max(4, atoi(e))) << 10;
    if (const char* e = getenv("IADH_RES_WIDE")) sscanf(e, "%d:%d", &c->res_wide_min, &c->res_wide_nt);
    if (const char* e = getenv("IADH_PASS_MIN")) c->pass_min = std::max(1, atoi(e));
    if (const char* e = getenv("IADH_EXPAND_MIN")) c->expand_min = std::max(1, atoi(e));
    if (const char* e = getenv("IADH_HL_EXPAND_MIN")) c->hl_expand_min = std::max(1, atoi(e));
    if (const char* e = getenv("IADH_EXPAND_SORT_MAX")) c->expand_sort_max = std::max(1, atoi(e));
    if (const char* e = getenv("IADH_PASS_RATIO")) c->pass_ratio = c->pass_ratio_hl = std::max(0, atoi(e));
    if (const char* e = getenv("IADH_PASS_RATIO_HL")) c->pass_ratio_hl = std::max(0, atoi(e));
    if (const char* e = getenv("IADH_PASS_CHUNK_MB")) c->pass_chunk_bytes = (int64_t)std::max(1, atoi(e)) << 20;
    if (const char* e = getenv("IADH_PASS_CFG")) {
        sscanf(e, "%d:%d", &c->pass_nt, &c->pass_per_sm);
        c->pass_cfg_user = true;
        if (!pass_cfg_supported(c->pass_nt) || c->pass_per_sm < 1)
            return fail(IADH_ERR_ARG, "IADH_PASS_CFG=%s: unsupported launch configuration (codes: 32 3224 3232 64 6412 6416 128)", e);
    }
    if (const char* e = getenv("IADH_CLASSES")) {  // "min_dots:threads:ctas_per_sm,..." heaviest class first
        int n = 0;
        const char* q = e;
        while (n < 4) {
            int a = 0, b = 0, d = 0;
            if (sscanf(q, "%d:%d:%d", &a, &b, &d) != 3) break;
            c->cls[n].min_dots = a;
            c->cls[n].nt = b;
            c->cls[n].per_sm = std::max(1, d);
            n++;
            q = strchr(q, ',');
            if (!q) break;
            q++;
        }
        if (n > 0) c->n_classes = n;
    }
    guard.c = nullptr;
    *out = c;
    return IADH_OK;
}

// ---------------------------------------------------------------------------------------------
// multi-GPU exchange: NCCL, looked up at run time (the library must load where NCCL is absent; a process that
// already loaded a libnccl.so.2 -- PyTorch's -- shares that copy)
#include <fcntl.h>
#include <nccl.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
struct CommState {
    void* lib = nullptr;
    ncclComm_t comm = nullptr;
    int rank = 0, n = 1;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    DBuf d_meta_send, d_meta_recv, d_send, d_recv;
    HBuf h_meta, h_stage, h_recv;
    // Single-node exchange through host shared memory (all ranks on one box): a gather's result is ONE segment
    // in /dev/shm that every rank maps; each rank copies only its own bodies from its device into their final place
    // and raises a tag; a reader waits for all tags.  Per step and box the results cross PCIe once instead of once
    // per rank (8 ranks x 320 MB of level-2 multiplicons: 2.6 GB -> 0.3 GB; the box's host ingest, ~80 GB/s for 8
    // GPUs together, was the longest part of the 8-GPU step).  NCCL still carries the sizes and acts as the barrier.
    static const int SHM_SLOTS = 4, SHM_MAX_RANKS = 64;
    struct ShmCtl {
        uint32_t done[SHM_MAX_RANKS][SHM_SLOTS];   // tag of the last gather whose part rank r has delivered into slot k
    };
    struct ShmSeg {
        uint8_t* base = nullptr;
        size_t cap = 0;
        uint32_t gen = 0;
        int outstanding = 0;   // live views (SharedPinned) of this rank
    };
    bool shm_ok = false;
    std::string shm_prefix;
    ShmCtl* ctl = nullptr;
    ShmSeg seg[SHM_SLOTS];
    uint32_t seq = 0;
};
void SharedPinned::wait()
{
    wait_own();
    if (!ext || failed) return;
    const double t0 = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
    for (int r = 0; r < comm->n; r++) {
        volatile uint32_t* tag = &comm->ctl->done[r][slot];
        for (long spin = 0; __atomic_load_n(tag, __ATOMIC_ACQUIRE) != want; spin++) {
            if (spin < 2000) sched_yield();
            else usleep(50);
            if ((spin & 1023) == 1023 &&
                std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count() - t0 > 300.0) {
                failed = true;
                return;
            }
        }
    }
}
void SharedPinned::release()
{
    wait_own();   // (our own copy targets memory that must stay mapped until it is over)
    if (ext) comm->seg[slot].outstanding--;
    else buf.release();
}
struct ShmDone {
    CommState::ShmCtl* ctl;
    int rank, slot;
    uint32_t tag;
};
static void CUDART_CB shm_done_cb(void* p)
{
    ShmDone* d = (ShmDone*)p;
    __atomic_store_n(&d->ctl->done[d->rank][d->slot], d->tag, __ATOMIC_RELEASE);
    delete d;
}
static void shm_unmap(CommState::ShmSeg& g)
{
    if (g.base) {
        cudaHostUnregister(g.base);
        munmap(g.base, g.cap);
    }
    g.base = nullptr;
    g.cap = 0;
}
// maps `bytes` of the named segment (rank 0 creates it first); returns null on failure
static void* shm_map(const std::string& name, size_t bytes, bool create)
{
    const int fd = shm_open(name.c_str(), create ? (O_CREAT | O_EXCL | O_RDWR) : O_RDWR, 0600);
    if (fd < 0) return nullptr;
    if (create && ftruncate(fd, (off_t)bytes) != 0) {
        close(fd);
        shm_unlink(name.c_str());
        return nullptr;
    }
    void* p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    return p == MAP_FAILED ? nullptr : p;
}
static void* nccl_lib()
{
    static void* lib = nullptr;
    if (lib) return lib;
    lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    return lib;
}
static int nccl_bind(CommState& s)
{
    s.lib = nccl_lib();
    if (!s.lib) return fail(IADH_ERR_STATE, "NCCL not found (libnccl.so.2): %s", dlerror());
    s.GetUniqueId = (decltype(s.GetUniqueId))dlsym(s.lib, "ncclGetUniqueId");
    s.CommInitRank = (decltype(s.CommInitRank))dlsym(s.lib, "ncclCommInitRank");
    s.AllGather = (decltype(s.AllGather))dlsym(s.lib, "ncclAllGather");
    s.CommDestroy = (decltype(s.CommDestroy))dlsym(s.lib, "ncclCommDestroy");
    s.GetErrorString = (decltype(s.GetErrorString))dlsym(s.lib, "ncclGetErrorString");
    if (!s.GetUniqueId || !s.CommInitRank || !s.AllGather || !s.CommDestroy || !s.GetErrorString)
        return fail(IADH_ERR_STATE, "libnccl.so.2 lacks an expected symbol");
    return IADH_OK;
}
#define NC(s, call)                                                                                              \
    do {                                                                                                         \
        ncclResult_t r_ = (call);                                                                                \
        if (r_ != ncclSuccess) return fail(IADH_ERR_CUDA, "%s failed: %s", #call, (s).GetErrorString(r_));        \
    } while (0)
// Communicators outlive contexts: a host program that builds a fresh context per run (cheap) must not pay
// ncclCommInitRank (hundreds of ms) again.  Keyed by the unique id; iadh_comm_destroy ends one explicitly.
static std::mutex g_comm_mu;
static std::map<std::string, CommState*> g_comms;
static void comm_release(iadh_ctx* c) { c->comm = nullptr; }   // detach only
static void comm_destroy(CommState* s)
{
    for (int k = 0; k < CommState::SHM_SLOTS; k++) shm_unmap(s->seg[k]);
    if (s->ctl) munmap(s->ctl, sizeof(CommState::ShmCtl));
    if (s->comm) s->CommDestroy(s->comm);
    DBuf* bufs[] = {&s->d_meta_send, &s->d_meta_recv, &s->d_send, &s->d_recv};
    for (DBuf* b : bufs) b->release();
    s->h_meta.release();
    s->h_stage.release();
    s->h_recv.release();
    delete s;
}

extern "C" {

void iadh_destroy(iadh_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    set_last_result(c, nullptr, 0);
    comm_release(c);
    if (c->st) cudaStreamSynchronize(c->st);
    DBuf* bufs[] = {&c->d_list_off, &c->d_fam, &c->d_gid, &c->d_num_id, &c->d_flags, &c->d_fam_ptr, &c->d_fam_members, &c->d_mem_y, &c->d_plus_prefix, &c->d_self_slot, &c->d_fam_list, &c->d_mem_list, &c->d_elem_list, &c->d_pair_idx, &c->d_row_used, &c->d_cur, &c->d_hl_items, &c->d_ghdr,
                    &c->d_tasks, &c->d_tile_prefix, &c->d_tile_cnt, &c->d_tile_off, &c->d_dots, &c->d_ws, &c->d_items, &c->d_items1, &c->d_passctr, &c->d_prof, &c->d_plan, &c->d_cnt, &c->d_outs,
                    &c->d_sel, &c->d_out_off, &c->d_out, &c->d_counter, &c->d_pf_col, &c->d_items_res, &c->d_items_fb, &c->d_resdbg, &c->d_istate, &c->d_vis, &c->d_init,
                    &c->d_links, &c->d_link_ptr, &c->d_elem_base, &c->d_lk_keys, &c->d_lk_vals, &c->d_lk_cols, &c->d_lk_cnt, &c->d_lk_ptr, &c->d_lk_col, &c->d_lk_bsum, &c->d_hdev};
    for (DBuf* b : bufs) b->release();
    c->h_tasks.release();
    c->h_misc.release();
    c->h_pf.release();
    c->h_vis.release();
    c->h_init.release();
    for (int i = 0; i < 8; i++)
        if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    if (c->ev_call) cudaEventDestroy(c->ev_call);
    if (c->st_copy) {
        cudaStreamSynchronize(c->st_copy);
        cudaStreamDestroy(c->st_copy);
    }
    if (c->ev_packed) cudaEventDestroy(c->ev_packed);
    if (c->ev_copied) cudaEventDestroy(c->ev_copied);
    if (c->ev_copied_recv) cudaEventDestroy(c->ev_copied_recv);
    for (int i = 0; i < 4; i++)
        if (c->user_ev[i]) cudaEventDestroy(c->user_ev[i]);
    if (c->st) cudaStreamDestroy(c->st);
    for (int i = 0; i < 3; i++) {
        if (c->st_side[i]) cudaStreamDestroy(c->st_side[i]);
        if (c->ev_join[i]) cudaEventDestroy(c->ev_join[i]);
    }
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    for (int i = 0; i < 8; i++) {
        if (c->st_res[i]) cudaStreamDestroy(c->st_res[i]);
        if (c->ev_res[i]) cudaEventDestroy(c->ev_res[i]);
    }
    delete c;
}

static int set_params_impl(iadh_ctx* c, const iadh_params* p);
int iadh_set_params(iadh_ctx* c, const iadh_params* p)
{
    IADH_TRY return set_params_impl(c, p);
    IADH_CATCH
}
static int set_params_impl(iadh_ctx* c, const iadh_params* p)
{
    if (!c || !p) return fail(IADH_ERR_ARG, "iadh_set_params: null argument");
    if (p->n_gaps < 0 || p->n_gaps > 10) return fail(IADH_ERR_ARG, "n_gaps must be 0..10");
    if (p->mult_hyp < 0 || p->mult_hyp > 2) return fail(IADH_ERR_ARG, "mult_hyp must be 0, 1 or 2");
    if (p->gap_size < 1 || p->cluster_gap < 1 || p->anchor_points < 3)
        return fail(IADH_ERR_ARG, "gap_size/cluster_gap must be >= 1 and anchor_points >= 3 (Settings.cpp:298-368)");
    c->prm.n_gaps = p->n_gaps;
    for (int i = 0; i < 10; i++) c->prm.gaps[i] = i < p->n_gaps ? p->gaps[i] : 0;
    c->prm.gap_size = p->gap_size;
    c->prm.cluster_gap = p->cluster_gap;
    c->prm.anchor_points = p->anchor_points;
    c->prm.mult_hyp = p->mult_hyp;
    c->prm.q_value = p->q_value;
    c->prm.prob_cutoff = p->prob_cutoff;
    return IADH_OK;
}

int iadh_level2_pairs(iadh_ctx* c, const int32_t* xs, const int32_t* ys, int64_t n, iadh_result** out)
{
    if (!c || !out || (n > 0 && (!xs || !ys))) return fail(IADH_ERR_ARG, "iadh_level2_pairs: null argument");
    CU(cudaSetDevice(c->device));
    IADH_TRY return run_l2(c, xs, ys, n, out);
    IADH_CATCH
}

int iadh_level2_all_pairs(iadh_ctx* c, int rank, int n_ranks, iadh_result** out)
{
    if (!c || !out) return fail(IADH_ERR_ARG, "iadh_level2_all_pairs: null argument");
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(IADH_ERR_ARG, "bad rank %d/%d", rank, n_ranks);
    CU(cudaSetDevice(c->device));
    IADH_TRY
    std::vector<int32_t> xs, ys;
    pair_sequence(c, rank, n_ranks, xs, ys);
    return run_l2(c, xs.data(), ys.data(), (int64_t)xs.size(), out);
    IADH_CATCH
}

static int mask_range_impl(iadh_ctx* c, int32_t list, int32_t begin, int32_t end);
int iadh_mask_range(iadh_ctx* c, int32_t list, int32_t begin, int32_t end)
{
    IADH_TRY return mask_range_impl(c, list, begin, end);
    IADH_CATCH
}
static int mask_range_impl(iadh_ctx* c, int32_t list, int32_t begin, int32_t end)
{
    if (!c) return fail(IADH_ERR_ARG, "null ctx");
    if (list < 0 || list >= c->n_lists) return fail(IADH_ERR_ARG, "mask_range: list %d out of range", list);
    const int64_t sz = c->list_off[list + 1] - c->list_off[list];
    if (begin < 0 || end >= sz) return fail(IADH_ERR_ARG, "mask_range: [%d,%d] outside list of %lld", begin, end, (long long)sz);
    if (end < begin) return IADH_OK;
    CU(cudaSetDevice(c->device));
    const int64_t o = c->list_off[list];
    for (int64_t i = o + begin; i <= o + end; i++)
        if (!(c->flags[i] & 4)) {
            c->flags[i] |= 4;
            c->masked_cnt[list]++;
        }
    launch_mask(c->d_flags.as<uint8_t>(), o + begin, o + end, 1, c->st);
    CU(cudaGetLastError());
    return IADH_OK;
}

static int clear_masks_impl(iadh_ctx* c);
int iadh_clear_masks(iadh_ctx* c)
{
    IADH_TRY return clear_masks_impl(c);
    IADH_CATCH
}
static int clear_masks_impl(iadh_ctx* c)
{
    if (!c) return fail(IADH_ERR_ARG, "null ctx");
    CU(cudaSetDevice(c->device));
    for (uint8_t& f : c->flags) f &= ~4;
    std::fill(c->masked_cnt.begin(), c->masked_cnt.end(), 0);
    if (!c->flags.empty()) {
        launch_mask(c->d_flags.as<uint8_t>(), 0, (int64_t)c->flags.size() - 1, 0, c->st);
        CU(cudaGetLastError());
    }
    return IADH_OK;
}

static int profile_search_impl(iadh_ctx* c, const iadh_profile_desc* p, const iadh_link_t* links, int64_t n_links, int rank, int n_ranks,
                               iadh_result** out);
int iadh_profile_search(iadh_ctx* c, const iadh_profile_desc* p, int rank, int n_ranks, iadh_result** out)
{
    IADH_TRY return profile_search_impl(c, p, nullptr, -1, rank, n_ranks, out);
    IADH_CATCH
}
int iadh_profile_search_filtered(iadh_ctx* c, const iadh_profile_desc* p, const iadh_link_t* links, int64_t n_links, int rank, int n_ranks,
                                 iadh_result** out)
{
    IADH_TRY
    if (!links || n_links <= 0) return fail(IADH_ERR_ARG, "iadh_profile_search_filtered: the Link set of the profile's multiplicon is required");
    const int rc = profile_search_impl(c, p, links, n_links, rank, n_ranks, out);
    if (c) c->hf_active = false;
    return rc;
    IADH_CATCH
}
// links == nullptr: no homolog filter
static int profile_search_impl(iadh_ctx* c, const iadh_profile_desc* p, const iadh_link_t* links, int64_t n_links, int rank, int n_ranks,
                               iadh_result** out)
{
    if (!c || !p || !out) return fail(IADH_ERR_ARG, "iadh_profile_search: null argument");
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(IADH_ERR_ARG, "bad rank %d/%d", rank, n_ranks);
    CU(cudaSetDevice(c->device));
    // (the context's own events: nothing to create, destroy or leak per call)
    cudaEvent_t e0 = c->ev[7], e1 = c->ev_call;
    CU(cudaEventRecord(e0, c->st));
    BatchOut b;
    memset(&b.st, 0, sizeof b.st);
    ProfileHost ph;
    int rc = upload_profile(c, p, ph, b.st.h2d_bytes);
    c->hf_active = false;
    if (rc == IADH_OK && links) {
        const int64_t lptr[2] = {0, n_links};
        rc = upload_links(c, p, 1, nullptr, links, lptr, ph.dev, b.st);
        c->hf_active = rc == IADH_OK;
    }
    std::vector<GhmTask> tasks;
    if (rc == IADH_OK) {
        std::vector<int32_t> part;
        list_partition(c, rank, n_ranks, part);
        static const char* only_list = getenv("IADH_ONLY_LIST");  // profiling aid: one GHMProfile per call
        for (int32_t l : part) {
            if (only_list && l != atoi(only_list)) continue;
            const unsigned sz = (unsigned)(c->list_off[l + 1] - c->list_off[l]);
            const unsigned masked = (unsigned)c->masked_cnt[l];
            if (sz - masked < (unsigned)c->prm.anchor_points) continue;  // higherLevel.cpp:40-41
            GhmTask t;
            init_task(t);
            t.x_list = -1;
            t.y_list = l;
            t.x_object_id = p->x_object_id;
            t.y_object_id = l;
            t.level = p->n_segments + 1;  // GHMProfile.cpp:14
            t.area = (unsigned)((unsigned)p->length * (sz - masked));  // GHM.cpp:228-234, level > 2
            t.seq = (int32_t)tasks.size();
            tasks.push_back(t);
        }
        rc = run_batch(c, true, ph.dev, tasks, b);
    }
    c->hf_active = false;
    cudaEventRecord(e1, c->st);
    cudaEventSynchronize(e1);
    if (rc == IADH_OK) {
        float ms = 0, m1 = 0, m2 = 0, m3 = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        b.st.total_ms = ms;
        if (!tasks.empty()) {
            cudaEventElapsedTime(&m1, c->ev[1], c->ev[2]);
            cudaEventElapsedTime(&m2, c->ev[3], c->ev[4]);
            b.st.build_ms = m1 + m2;
            cudaEventElapsedTime(&m3, c->ev[5], c->ev[6]);
            b.st.cluster_ms = m3;
        }
        *out = make_result(b);
        set_last_result(c, *out, (*out)->n_bytes - 4);
    }
    return rc;
}

__global__ void apply_vis_kernel(uint8_t* __restrict__ flags, const int16_t* __restrict__ vis, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && vis[i] != 0x7fff) flags[i] |= 4;
}

static int profile_search_batch_impl(iadh_ctx* c, const iadh_profile_desc* profiles, int32_t n_profiles, const iadh_link_t* links,
                              const int64_t* link_ptr, const iadh_mask_range_t* masks,
                              const int32_t* mask_ptr, int rank, int n_ranks, int shard_profiles, iadh_result** out);
int iadh_profile_search_batch(iadh_ctx* c, const iadh_profile_desc* profiles, int32_t n_profiles, const iadh_mask_range_t* masks,
                              const int32_t* mask_ptr, int rank, int n_ranks, int shard_profiles, iadh_result** out)
{
    IADH_TRY return profile_search_batch_impl(c, profiles, n_profiles, nullptr, nullptr, masks, mask_ptr, rank, n_ranks, shard_profiles, out);
    IADH_CATCH
}
int iadh_profile_search_batch_filtered(iadh_ctx* c, const iadh_profile_desc* profiles, int32_t n_profiles, const iadh_link_t* links,
                                       const int64_t* link_ptr, const iadh_mask_range_t* masks, const int32_t* mask_ptr, int rank,
                                       int n_ranks, int shard_profiles, iadh_result** out)
{
    IADH_TRY
    if (n_profiles > 0 && (!links || !link_ptr)) return fail(IADH_ERR_ARG, "iadh_profile_search_batch_filtered: the Link sets are required");
    const int rc = profile_search_batch_impl(c, profiles, n_profiles, links, link_ptr, masks, mask_ptr, rank, n_ranks, shard_profiles, out);
    if (c) c->hf_active = false;
    return rc;
    IADH_CATCH
}
// links == nullptr: no homolog filter
static int profile_search_batch_impl(iadh_ctx* c, const iadh_profile_desc* profiles, int32_t n_profiles, const iadh_link_t* links,
                              const int64_t* link_ptr, const iadh_mask_range_t* masks,
                              const int32_t* mask_ptr, int rank, int n_ranks, int shard_profiles, iadh_result** out)
{
    if (!c || !out || n_profiles < 0 || (n_profiles > 0 && (!profiles || !mask_ptr)))
        return fail(IADH_ERR_ARG, "iadh_profile_search_batch: null argument");
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(IADH_ERR_ARG, "bad rank %d/%d", rank, n_ranks);
    if (n_profiles > 0x7ffe) return fail(IADH_ERR_CAPACITY, "at most 32766 profiles per batch");
    if (n_profiles == 0) return IADH_OK;
    if (mask_ptr[0] != 0) return fail(IADH_ERR_ARG, "iadh_profile_search_batch: mask_ptr[0] must be 0");
    for (int k = 0; k < n_profiles; k++) {
        out[k] = nullptr;
        if (mask_ptr[k + 1] < mask_ptr[k] || (mask_ptr[k + 1] > 0 && !masks)) return fail(IADH_ERR_ARG, "iadh_profile_search_batch: bad mask_ptr");
        for (int m = mask_ptr[k]; m < mask_ptr[k + 1]; m++) {
            const iadh_mask_range_t& r = masks[m];
            if (r.list < 0 || r.list >= c->n_lists) return fail(IADH_ERR_ARG, "batch mask %d: list %d out of range", m, r.list);
            const int64_t sz = c->list_off[r.list + 1] - c->list_off[r.list];
            if (r.begin < 0 || r.end >= sz) return fail(IADH_ERR_ARG, "batch mask %d: [%d,%d] outside list of %lld", m, r.begin, r.end, (long long)sz);
        }
    }
    CU(cudaSetDevice(c->device));
    cudaEvent_t e0 = c->ev[7], e1 = c->ev_call;
    CU(cudaEventRecord(e0, c->st));
    const int64_t total = c->n_lists ? c->list_off[c->n_lists] : 0;
    BatchOut b;
    memset(&b.st, 0, sizeof b.st);
    // mask epochs + the GHMs of every profile, as the sequential calls would see them
    CU(c->h_vis.ensure(2 * (size_t)total + 16));
    CU(c->d_vis.ensure(2 * (size_t)total + 16));
    int16_t* vis = c->h_vis.as<int16_t>();
    for (int64_t i = 0; i < total; i++) vis[i] = (c->flags[i] & 4) ? 0 : 0x7fff;
    std::vector<int32_t> mc = c->masked_cnt, part_lists;
    if (shard_profiles) list_partition(c, 0, 1, part_lists);
    else list_partition(c, rank, n_ranks, part_lists);
    std::vector<GhmTask> tasks;
    std::vector<int> t_begin((size_t)n_profiles + 1, 0);
    std::vector<uint8_t> mine((size_t)n_profiles, 0);
    int n_mine = 0;
    for (int k = 0; k < n_profiles; k++) {
        for (int m = mask_ptr[k]; m < mask_ptr[k + 1]; m++) {
            const iadh_mask_range_t& r = masks[m];
            const int64_t o = c->list_off[r.list];
            for (int64_t i = o + r.begin; i <= o + r.end; i++)
                if (vis[i] == 0x7fff) {
                    vis[i] = (int16_t)k;
                    mc[r.list]++;
                }
        }
        t_begin[k] = (int)tasks.size();
        mine[k] = (!shard_profiles || k % n_ranks == rank) ? 1 : 0;
        if (!mine[k]) continue;
        n_mine++;
        const iadh_profile_desc& p = profiles[k];
        for (int32_t l : part_lists) {
            const unsigned sz = (unsigned)(c->list_off[l + 1] - c->list_off[l]);
            const unsigned masked = (unsigned)mc[l];
            if (sz - masked < (unsigned)c->prm.anchor_points) continue;  // higherLevel.cpp:40-41
            GhmTask t;
            init_task(t);
            t.x_list = -1;
            t.y_list = l;
            t.x_object_id = p.x_object_id;
            t.y_object_id = l;
            t.level = p.n_segments + 1;
            t.area = (unsigned)((unsigned)p.length * (sz - masked));
            t.prof = k;
            t.seq = (int32_t)tasks.size();
            tasks.push_back(t);
        }
    }
    t_begin[n_profiles] = (int)tasks.size();
    // all profiles of this rank in one staging block, followed by their descriptors
    int rc = IADH_OK;
    std::vector<std::vector<int32_t>> col_ptr((size_t)n_profiles), gfam((size_t)n_profiles), ggid((size_t)n_profiles);
    std::vector<std::vector<uint8_t>> gstrand((size_t)n_profiles);
    std::vector<size_t> p_off((size_t)n_profiles + 1, 0);
    ProfileDev dummy;
    c->batch_L.assign((size_t)n_profiles, 0);
    c->hl_items.clear();
    for (int k = 0; k < n_profiles && rc == IADH_OK; k++) {
        p_off[k + 1] = p_off[k];
        if ((rc = profile_groups(c, &profiles[k], col_ptr[k], gfam[k], ggid[k], gstrand[k]))) break;   // (validates every descriptor)
        c->batch_L[k] = profiles[k].length;
        if (mine[k]) append_hl_items(c, k, col_ptr[k], gfam[k], ggid[k], gstrand[k]);
        if (mine[k]) p_off[k + 1] += place_profile(c, &profiles[k], col_ptr[k], gfam[k], ggid[k], gstrand[k], nullptr, nullptr, dummy);
    }
    if (rc) return rc;
    const size_t desc_off = p_off[n_profiles], blob = desc_off + sizeof(ProfileDev) * (size_t)n_profiles;
    CU(c->h_pf.ensure(blob + 16));
    CU(c->d_pf_col.ensure(blob + 16));
    {
        uint8_t* h = c->h_pf.as<uint8_t>();
        uint8_t* dv = c->d_pf_col.as<uint8_t>();
        ProfileDev* hd = (ProfileDev*)(h + desc_off);
        for (int k = 0; k < n_profiles; k++) {
            memset(&hd[k], 0, sizeof(ProfileDev));
            if (mine[k]) place_profile(c, &profiles[k], col_ptr[k], gfam[k], ggid[k], gstrand[k], h + p_off[k], dv + p_off[k], hd[k]);
        }
        CU(cudaMemcpyAsync(dv, h, blob, cudaMemcpyHostToDevice, c->st));
        CU(cudaMemcpyAsync(c->d_vis.p, vis, 2 * (size_t)total, cudaMemcpyHostToDevice, c->st));
        b.st.h2d_bytes += (int64_t)blob + 2 * total;
    }
    ProfileDev pf;
    memset(&pf, 0, sizeof pf);
    pf.batch = (const ProfileDev*)(c->d_pf_col.as<uint8_t>() + desc_off);
    pf.vis = c->d_vis.as<int16_t>();
    set_last_result(c, nullptr, 0);
    b.group_begin = &t_begin;
    c->hf_active = false;
    if (links) {
        if ((rc = upload_links(c, profiles, n_profiles, mine.data(), links, link_ptr, pf, b.st))) return rc;
        c->hf_active = true;
    }
    rc = run_batch(c, true, pf, tasks, b);
    c->hf_active = false;
    // the masks of the whole batch now hold, as after the sequential calls
    if (rc == IADH_OK && total > 0 && mask_ptr[n_profiles] > 0) {
        apply_vis_kernel<<<(unsigned)((total + 255) / 256), 256, 0, c->st>>>(c->d_flags.as<uint8_t>(), c->d_vis.as<int16_t>(), total);
        for (int64_t i = 0; i < total; i++)
            if (vis[i] != 0x7fff) c->flags[i] |= 4;
        c->masked_cnt = mc;
    }
    cudaEventRecord(e1, c->st);
    cudaEventSynchronize(e1);
    if (rc != IADH_OK) return rc;
    float ms = 0, m1 = 0, m2 = 0, m3 = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (!tasks.empty()) {
        cudaEventElapsedTime(&m1, c->ev[1], c->ev[2]);
        cudaEventElapsedTime(&m2, c->ev[3], c->ev[4]);
        cudaEventElapsedTime(&m3, c->ev[5], c->ev[6]);
    }
    // one result per profile: slices of the batch's pinned stream (run_batch laid it out as one [count][blocks] stream
    // per profile), still on its way to the host -- iadh_result_bytes waits for the copy
    SharedPinned* sp = new SharedPinned();
    std::swap(sp->buf, b.packed);
    sp->ready = b.ready;
    b.ready = nullptr;
    sp->refs = n_profiles;
    const double share = n_mine > 0 ? 1.0 / n_mine : 0.0;
    for (int k = 0; k < n_profiles; k++) {
        iadh_result* r = new iadh_result();
        memset(&r->stats, 0, sizeof r->stats);
        for (int t = t_begin[k]; t < t_begin[k + 1]; t++) {
            const GhmTask& tk = b.tasks[t];
            r->stats.n_dots += (int64_t)tk.cnt[0] + tk.cnt[1];
            r->stats.n_multiplicons += tk.n_mult;
            if (tk.n_mult > 0) {
                r->stats.n_baseclusters += tk.n_clusters;
                r->stats.n_anchorpoints += tk.n_aps;
            }
            for (int q = 0; q < tk.n_mult; q++) r->ghm_index.push_back(t - t_begin[k]);
        }
        r->stats.n_ghm = t_begin[k + 1] - t_begin[k];
        r->shared = sp;
        r->shared_off = b.group_off[k];
        r->n_bytes = b.group_off[k + 1] - b.group_off[k];
        if (k == 0) {   // (the first 4 bytes of the block are not part of the device stream)
            const int32_t nm = (int32_t)r->stats.n_multiplicons;
            memcpy(sp->buf.p, &nm, 4);
        }
        if (mine[k]) {   // device times of the batch, spread evenly over this rank's profiles
            r->stats.total_ms = ms * share;
            r->stats.build_ms = (m1 + m2) * share;
            r->stats.cluster_ms = m3 * share;
            r->stats.h2d_bytes = (int64_t)(b.st.h2d_bytes * share);
            r->stats.d2h_bytes = (int64_t)(b.st.d2h_bytes * share);
        }
        out[k] = r;
    }
    const size_t pos = b.packed_bytes - 4;   // bytes of the device stream
    // the batch's stream is still in d_out, its results in order: iadh_allgather_results sends it from there
    for (int k = 0; k < n_profiles; k++) {
        c->last_batch.push_back(out[k]);
        out[k]->owner = c;
    }
    c->last_batch_body = pos;
    // launches are counted once, on the first profile this rank searched
    for (int k = 0; k < n_profiles; k++)
        if (mine[k]) {
            out[k]->stats.kernel_launches = b.st.kernel_launches + (mask_ptr[n_profiles] > 0 ? 1 : 0);
            break;
        }
    return IADH_OK;
}

const uint8_t* iadh_result_bytes(const iadh_result* r, size_t* n)
{
    if (n) *n = r ? r->n_bytes : 0;
    if (!r) return nullptr;
    if (r->ready) {   // the stream may still be on its way to the host
        cudaEventSynchronize(r->ready);
        cudaEventDestroy(r->ready);
        r->ready = nullptr;
    }
    if (r->shared) {
        r->shared->wait();
        if (r->shared->failed) {
            fail(IADH_ERR_CUDA, "gathered result: a rank did not deliver its part within the time limit");
            return nullptr;
        }
        return r->shared->base() + r->shared_off;
    }
    return r->owned.empty() ? (const uint8_t*)r->bytes.p : r->owned.data();
}
size_t iadh_result_size(const iadh_result* r) { return r ? r->n_bytes : 0; }
const int32_t* iadh_result_ghm_index(const iadh_result* r, size_t* n)
{
    if (n) *n = r ? r->ghm_index.size() : 0;
    return r ? r->ghm_index.data() : nullptr;
}
int iadh_result_stats(const iadh_result* r, iadh_stats* s)
{
    if (!r || !s) return fail(IADH_ERR_ARG, "null argument");
    *s = r->stats;
    return IADH_OK;
}
void iadh_free_result(iadh_result* r)
{
    if (!r) return;
    if (r->owner) set_last_result(r->owner, nullptr, 0);   // (it, or the batch it belongs to, is no longer "the last one")
    if (r->ready) {   // the pinned buffer goes back to the cache: the copy into it must be over
        cudaEventSynchronize(r->ready);
        cudaEventDestroy(r->ready);
        r->ready = nullptr;
    }
    r->bytes.release();
    if (r->shared && --r->shared->refs == 0) {
        r->shared->release();
        delete r->shared;
    }
    delete r;
}
void iadh_trim(void) { g_cache.trim(); }

int iadh_warmup(int device)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(IADH_ERR_CUDA, "no CUDA device available; this library has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(IADH_ERR_ARG, "device %d out of range (%d devices)", device, ndev);
    CU(cudaSetDevice(device));
    CU(cudaFree(0));  // creates the primary context (0.5-1 s in a fresh process)
    return IADH_OK;
}

int iadh_comm_unique_id(uint8_t* id_out, size_t n_bytes)
{
    if (!id_out || n_bytes < sizeof(ncclUniqueId)) return fail(IADH_ERR_ARG, "iadh_comm_unique_id: need %zu bytes", sizeof(ncclUniqueId));
    CommState s;
    int rc = nccl_bind(s);
    if (rc) return rc;
    ncclUniqueId id;
    NC(s, s.GetUniqueId(&id));
    memset(id_out, 0, n_bytes);
    memcpy(id_out, &id, sizeof id);
    return IADH_OK;
}

// one int64 per rank, to every rank (NCCL all-gather on the context's stream; also the barrier of the shm set-up)
static int comm_exchange(iadh_ctx* c, CommState& s, int64_t mine, std::vector<int64_t>& all)
{
    const int N = s.n;
    CU(s.h_meta.ensure(8 * (size_t)(N + 1) + 64));
    CU(s.d_meta_send.ensure(64));
    CU(s.d_meta_recv.ensure(8 * (size_t)N + 64));
    int64_t* h = s.h_meta.as<int64_t>();
    h[0] = mine;
    CU(cudaMemcpyAsync(s.d_meta_send.p, h, 8, cudaMemcpyHostToDevice, c->st));
    NC(s, s.AllGather(s.d_meta_send.p, s.d_meta_recv.p, 8, ncclChar, s.comm, c->st));
    CU(readback(h + 1, s.d_meta_recv.p, 8 * (size_t)N, c->st));
    CU(cudaStreamSynchronize(c->st));
    all.assign(h + 1, h + 1 + N);
    return IADH_OK;
}
static bool all_equal(const std::vector<int64_t>& v, int64_t x)
{
    for (int64_t y : v)
        if (y != x) return false;
    return true;
}
// decides, collectively, whether the ranks can exchange results through /dev/shm (same host, segment creatable)
static int shm_setup(iadh_ctx* c, CommState& s, const uint8_t* id)
{
    s.shm_ok = false;
    if (s.n < 2 || s.n > CommState::SHM_MAX_RANKS) return IADH_OK;
    if (const char* e = getenv("IADH_SHM_EXCHANGE"))
        if (atoi(e) == 0) return IADH_OK;   // (every rank sees the same environment under torchrun / mpirun; checked below anyway)
    // same host?  (host name + boot id, hashed)
    uint64_t hh = 1469598103934665603ull;
    {
        char buf[512];
        memset(buf, 0, sizeof buf);
        gethostname(buf, 255);
        if (FILE* f = fopen("/proc/sys/kernel/random/boot_id", "r")) {
            if (!fgets(buf + 256, 200, f)) buf[256] = 0;
            fclose(f);
        }
        for (size_t k = 0; k < sizeof buf; k++) hh = (hh ^ (uint8_t)buf[k]) * 1099511628211ull;
    }
    uint64_t hu = 1469598103934665603ull;
    for (size_t k = 0; k < sizeof(ncclUniqueId); k++) hu = (hu ^ id[k]) * 1099511628211ull;
    char name[64];
    snprintf(name, sizeof name, "/iadh_%016llx", (unsigned long long)hu);
    s.shm_prefix = name;
    std::vector<int64_t> all;
    int rc;
    if ((rc = comm_exchange(c, s, (int64_t)(hh >> 1), all))) return rc;
    const bool same_host = all_equal(all, (int64_t)(hh >> 1));
    // control block: rank 0 creates, barrier, the others map, barrier, rank 0 removes the name (mappings stay)
    const std::string ctl_name = s.shm_prefix + "_ctl";
    void* ctl = nullptr;
    if (same_host && s.rank == 0) {
        shm_unlink(ctl_name.c_str());
        ctl = shm_map(ctl_name, sizeof(CommState::ShmCtl), true);
        if (ctl) memset(ctl, 0, sizeof(CommState::ShmCtl));
    }
    if ((rc = comm_exchange(c, s, (same_host && (s.rank != 0 || ctl)) ? 1 : 0, all))) return rc;
    const bool go = all_equal(all, 1);
    if (go && s.rank != 0) ctl = shm_map(ctl_name, sizeof(CommState::ShmCtl), false);
    if ((rc = comm_exchange(c, s, ctl ? 1 : 0, all))) return rc;
    if (s.rank == 0) shm_unlink(ctl_name.c_str());
    if (all_equal(all, 1)) {
        s.ctl = (CommState::ShmCtl*)ctl;
        s.shm_ok = true;
    } else if (ctl)
        munmap(ctl, sizeof(CommState::ShmCtl));
    return IADH_OK;
}
// makes slot k hold at least `need` bytes on every rank (collective); false = the exchange falls back to NCCL
static int shm_grow(iadh_ctx* c, CommState& s, int k, size_t need, bool& ok)
{
    CommState::ShmSeg& g = s.seg[k];
    ok = true;
    if (g.cap >= need) return IADH_OK;
    shm_unmap(g);
    g.gen++;
    const size_t cap = ((need + need / 2 + ((size_t)2 << 20)) >> 21) << 21;
    char suffix[48];
    snprintf(suffix, sizeof suffix, "_d%d_%u", k, g.gen);
    const std::string name = s.shm_prefix + suffix;
    void* p = nullptr;
    std::vector<int64_t> all;
    int rc;
    if (s.rank == 0) {
        shm_unlink(name.c_str());
        p = shm_map(name, cap, true);
    }
    if ((rc = comm_exchange(c, s, (s.rank != 0 || p) ? 1 : 0, all))) return rc;
    const bool created = all_equal(all, 1);
    if (created && s.rank != 0) p = shm_map(name, cap, false);
    // pinned for this process's device: the copies into it run at full PCIe speed
    bool mine_ok = created && p && cudaHostRegister(p, cap, cudaHostRegisterPortable) == cudaSuccess;
    if (!mine_ok) cudaGetLastError();
    if ((rc = comm_exchange(c, s, mine_ok ? 1 : 0, all))) return rc;
    if (s.rank == 0) shm_unlink(name.c_str());
    if (all_equal(all, 1)) {
        g.base = (uint8_t*)p;
        g.cap = cap;
        return IADH_OK;
    }
    if (p) {
        if (mine_ok) cudaHostUnregister(p);
        munmap(p, cap);
    }
    ok = false;
    s.shm_ok = false;   // (a box that cannot provide the segment once is not asked again)
    return IADH_OK;
}

int iadh_comm_init(iadh_ctx* c, const uint8_t* id, size_t n_bytes, int rank, int n_ranks)
{
    if (!c || !id || n_bytes < sizeof(ncclUniqueId) || n_ranks < 1 || rank < 0 || rank >= n_ranks)
        return fail(IADH_ERR_ARG, "iadh_comm_init: bad argument");
    CU(cudaSetDevice(c->device));
    comm_release(c);
    const std::string key((const char*)id, sizeof(ncclUniqueId));
    {
        std::lock_guard<std::mutex> lk(g_comm_mu);
        auto it = g_comms.find(key);
        if (it != g_comms.end()) {
            if (it->second->rank != rank || it->second->n != n_ranks)
                return fail(IADH_ERR_ARG, "iadh_comm_init: this id was joined as rank %d/%d", it->second->rank, it->second->n);
            c->comm = it->second;
            return IADH_OK;
        }
    }
    CommState* s = new CommState();
    int rc = nccl_bind(*s);
    if (rc) {
        delete s;
        return rc;
    }
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof uid);
    const ncclResult_t r = s->CommInitRank(&s->comm, n_ranks, uid, rank);
    if (r != ncclSuccess) {
        rc = fail(IADH_ERR_CUDA, "ncclCommInitRank failed: %s", s->GetErrorString(r));
        delete s;
        return rc;
    }
    s->rank = rank;
    s->n = n_ranks;
    c->comm = s;
    {
        std::lock_guard<std::mutex> lk(g_comm_mu);
        g_comms[key] = s;
    }
    return shm_setup(c, *s, id);
}

int iadh_comm_destroy(iadh_ctx* c)
{
    if (!c) return fail(IADH_ERR_ARG, "null ctx");
    cudaSetDevice(c->device);
    CommState* s = c->comm;
    c->comm = nullptr;
    if (s) {
        cudaDeviceSynchronize();
        std::lock_guard<std::mutex> lk(g_comm_mu);
        for (auto it = g_comms.begin(); it != g_comms.end(); ++it)
            if (it->second == s) {
                g_comms.erase(it);
                break;
            }
        comm_destroy(s);
    }
    return IADH_OK;
}

// per result and rank: body bytes, multiplicons, GHMs, dots
static const int META = 5;   // (+ on result 0: this rank's live views of the shm slot of this gather, -1 = no shm)

static int allgather_results_impl(iadh_ctx* c, iadh_result* const* mine, int32_t n, iadh_result** all);
int iadh_allgather_results(iadh_ctx* c, iadh_result* const* mine, int32_t n, iadh_result** all)
{
    IADH_TRY return allgather_results_impl(c, mine, n, all);
    IADH_CATCH
}
static int allgather_results_impl(iadh_ctx* c, iadh_result* const* mine, int32_t n, iadh_result** all)
{
    if (!c || !mine || !all || n < 0) return fail(IADH_ERR_ARG, "iadh_allgather_results: bad argument");
    if (!c->comm) return fail(IADH_ERR_STATE, "iadh_allgather_results: no communicator (iadh_comm_init)");
    if (n == 0) return IADH_OK;
    for (int i = 0; i < n; i++) {
        if (!mine[i]) return fail(IADH_ERR_ARG, "iadh_allgather_results: null result %d", i);
        all[i] = nullptr;
    }
    CU(cudaSetDevice(c->device));
    CommState& s = *c->comm;
    const int N = s.n;
    // 1. sizes of every rank's streams
    CU(s.h_meta.ensure(8 * (size_t)META * n * (N + 1)));
    CU(s.d_meta_send.ensure(8 * (size_t)META * n));
    CU(s.d_meta_recv.ensure(8 * (size_t)META * n * N));
    int64_t* hm = s.h_meta.as<int64_t>();
    int64_t my_total = 0;
    for (int i = 0; i < n; i++) {
        hm[META * i] = (int64_t)mine[i]->n_bytes - 4;
        hm[META * i + 1] = mine[i]->stats.n_multiplicons;
        hm[META * i + 2] = mine[i]->stats.n_ghm;
        hm[META * i + 3] = mine[i]->stats.n_dots;
        hm[META * i + 4] = 0;
        my_total += hm[META * i];
    }
    const int slot_k = (int)(s.seq % CommState::SHM_SLOTS);
    hm[4] = s.shm_ok ? s.seg[slot_k].outstanding : -1;
    CU(cudaMemcpyAsync(s.d_meta_send.p, hm, 8 * (size_t)META * n, cudaMemcpyHostToDevice, c->st));
    NC(s, s.AllGather(s.d_meta_send.p, s.d_meta_recv.p, 8 * (size_t)META * n, ncclChar, s.comm, c->st));
    CU(readback(hm + (size_t)META * n, s.d_meta_recv.p, 8 * (size_t)META * n * N, c->st));
    CU(cudaStreamSynchronize(c->st));
    const std::vector<int64_t> meta_all(hm + (size_t)META * n, hm + (size_t)META * n * (N + 1));   // (the staging block is reused below)
    const int64_t* ha = meta_all.data();
    // 1b. single-node exchange through shared memory, when every rank can and nobody still reads this slot
    bool use_shm = s.shm_ok;
    for (int r = 0; r < N; r++) use_shm = use_shm && ha[(size_t)r * n * META + 4] == 0;
    if (use_shm) {
        std::vector<size_t> res_off((size_t)n + 1, 0);
        for (int i = 0; i < n; i++) {
            size_t bytes = 4;
            for (int k = 0; k < N; k++) bytes += (size_t)ha[((size_t)k * n + i) * META];
            res_off[i + 1] = res_off[i] + bytes;
        }
        bool ok = false;
        int rc = shm_grow(c, s, slot_k, res_off[n] + 64, ok);
        if (rc) return rc;
        if (ok) {
            CommState::ShmSeg& g = s.seg[slot_k];
            const uint32_t tag = s.seq + 1;
            // where my bodies are: still on the device (the stream of the last call / batch in d_out) or on the host
            bool on_device = (n == 1 && c->last_result == mine[0] && c->last_result_body == (size_t)my_total);
            if (!on_device && (size_t)n == c->last_batch.size() && c->last_batch_body == (size_t)(my_total + 4 * (int64_t)(n - 1))) {
                on_device = true;
                for (int i = 0; i < n; i++) on_device = on_device && c->last_batch[i] == mine[i];
            }
            SharedPinned* sp = new SharedPinned();
            sp->ext = g.base;
            sp->comm = &s;
            sp->slot = slot_k;
            sp->want = tag;
            sp->refs = n;
            g.outstanding++;
            if (on_device) {
                CU(cudaEventRecord(c->ev_packed, c->st));
                CU(cudaStreamWaitEvent(c->st_copy, c->ev_packed, 0));
            }
            size_t src = 0;   // position inside my device stream (body i > 0 behind a 4-byte hole)
            for (int i = 0; i < n; i++) {
                iadh_result* r = new iadh_result();
                memset(&r->stats, 0, sizeof r->stats);
                size_t o = res_off[i] + 4;
                for (int k = 0; k < N; k++) {
                    const int64_t* m = ha + ((size_t)k * n + i) * META;
                    if (k == s.rank) {
                        if (i) src += 4;
                        if (m[0] && on_device)
                            CU(copy_to_host_in_pieces(g.base + o, c->d_out.as<uint8_t>() + src, (size_t)m[0], c->st_copy));
                        else if (m[0]) {
                            size_t nb = 0;
                            const uint8_t* hp = iadh_result_bytes(mine[i], &nb);
                            if (hp) memcpy(g.base + o, hp + 4, (size_t)m[0]);
                        }
                        src += (size_t)m[0];
                    }
                    o += (size_t)m[0];
                    r->stats.n_multiplicons += m[1];
                    r->stats.n_ghm += m[2];
                    r->stats.n_dots += m[3];
                }
                if (s.rank == 0) {   // the counts in front of every result are rank 0's part
                    const int32_t nm = (int32_t)r->stats.n_multiplicons;
                    memcpy(g.base + res_off[i], &nm, 4);
                }
                r->shared = sp;
                r->shared_off = res_off[i];
                r->n_bytes = res_off[i + 1] - res_off[i];
                r->stats.d2h_bytes = ha[((size_t)s.rank * n + i) * META];   // (only this rank's bodies cross its PCIe link)
                all[i] = r;
            }
            // my tag goes up when my copies are over (host function on the copy stream: no thread of the caller involved)
            ShmDone* d = new ShmDone{s.ctl, s.rank, slot_k, tag};
            if (on_device) {
                CU(cudaEventRecord(c->ev_copied, c->st_copy));
                c->copy_pending = true;
                CU(cudaLaunchHostFunc(c->st_copy, shm_done_cb, d));
                CU(cudaEventCreateWithFlags(&sp->ready, cudaEventDisableTiming));
                CU(cudaEventRecord(sp->ready, c->st_copy));
            } else
                shm_done_cb(d);
            s.seq++;
            return IADH_OK;
        }
    }
    s.seq++;   // (gathers are numbered on every rank alike, whichever way they travel)
    // 2. payload: this rank's bodies in order, body i > 0 preceded by a 4-byte hole (the layout a batch leaves on the
    //    device: run_batch writes the per-profile counts there), padded to the longest rank (NCCL gathers equal slots)
    std::vector<int64_t> tot((size_t)N, 0);
    int64_t max_tot = 0;
    for (int r = 0; r < N; r++) {
        for (int i = 0; i < n; i++) tot[r] += ha[((size_t)r * n + i) * META];
        tot[r] += 4 * (int64_t)(n - 1);
        max_tot = std::max(max_tot, tot[r]);
    }
    my_total += 4 * (int64_t)(n - 1);
    const size_t slot = (size_t)((max_tot + 15) & ~(int64_t)15);
    if (slot > 0) {
        const void* send = nullptr;
        bool on_device = (n == 1 && c->last_result == mine[0] && c->last_result_body == (size_t)my_total);
        if (!on_device && (size_t)n == c->last_batch.size() && c->last_batch_body == (size_t)my_total) {
            on_device = true;
            for (int i = 0; i < n; i++) on_device = on_device && c->last_batch[i] == mine[i];
        }
        if (on_device && c->d_out.cap >= slot)
            send = c->d_out.p;   // still on the device: no re-upload (bytes past the stream are padding)
        else {
            // one pinned staging block, one copy (the results' own buffers may be pageable)
            CU(s.d_send.ensure(slot));
            CU(s.h_stage.ensure(slot));
            size_t o = 0;
            for (int i = 0; i < n; i++) {
                size_t nb = 0;
                const uint8_t* p = iadh_result_bytes(mine[i], &nb);
                if (i) o += 4;
                if (nb > 4) memcpy(s.h_stage.as<uint8_t>() + o, p + 4, nb - 4);
                o += nb - 4;
            }
            if (o) CU(cudaMemcpyAsync(s.d_send.p, s.h_stage.p, o, cudaMemcpyHostToDevice, c->st));
            send = s.d_send.p;
        }
        // an earlier gather's copy to the host may still be reading the receive buffer: the collective waits for it on
        // the device (the host only when the buffer has to grow)
        if (c->copy_pending_recv) {
            if (slot * (size_t)N > s.d_recv.cap) CU(cudaEventSynchronize(c->ev_copied_recv));
            else CU(cudaStreamWaitEvent(c->st, c->ev_copied_recv, 0));
        }
        CU(s.d_recv.ensure(slot * (size_t)N));
        NC(s, s.AllGather(send, s.d_recv.p, slot, ncclChar, s.comm, c->st));
    }
    // 3. results: for every i [summed count][body of rank 0][body of rank 1]..., all in one pinned block that the
    //    results share as slices; every body travels from the receive buffer to its place on the copy stream (the
    //    gathered streams reach the host while the next call computes), the counts are written by the host
    SharedPinned* sp = new SharedPinned();
    std::vector<size_t> res_off((size_t)n + 1, 0);
    for (int i = 0; i < n; i++) {
        size_t bytes = 4;
        for (int k = 0; k < N; k++) bytes += (size_t)ha[((size_t)k * n + i) * META];
        res_off[i + 1] = res_off[i] + bytes;
    }
    if (sp->buf.ensure(res_off[n] + 16) != cudaSuccess) {
        delete sp;
        return fail(IADH_ERR_CUDA, "out of pinned host memory");
    }
    sp->refs = n;
    CU(cudaEventRecord(c->ev_packed, c->st));
    CU(cudaStreamWaitEvent(c->st_copy, c->ev_packed, 0));
    std::vector<size_t> src_off((size_t)N, 0);   // position inside rank k's stream
    for (int i = 0; i < n; i++) {
        iadh_result* r = new iadh_result();
        memset(&r->stats, 0, sizeof r->stats);
        size_t o = res_off[i] + 4;
        for (int k = 0; k < N; k++) {
            const int64_t* m = ha + ((size_t)k * n + i) * META;
            if (i) src_off[k] += 4;
            if (m[0]) {
                // (adjacent pieces -- every rank's body of a single result -- would merge into one copy; with n > 1
                // the pieces of a rank are separated by the other results)
                cudaError_t e = copy_to_host_in_pieces(sp->buf.as<uint8_t>() + o, (uint8_t*)s.d_recv.p + slot * (size_t)k + src_off[k], (size_t)m[0],
                                                       c->st_copy);
                if (e != cudaSuccess) {
                    cudaStreamSynchronize(c->st_copy);
                    sp->buf.release();
                    delete sp;
                    delete r;
                    for (int q = 0; q < i; q++) {
                        all[q]->shared = nullptr;
                        delete all[q];
                        all[q] = nullptr;
                    }
                    return fail(IADH_ERR_CUDA, "%s", cudaGetErrorString(e));
                }
            }
            o += (size_t)m[0];
            src_off[k] += (size_t)m[0];
            r->stats.n_multiplicons += m[1];
            r->stats.n_ghm += m[2];
            r->stats.n_dots += m[3];
        }
        const int32_t nm = (int32_t)r->stats.n_multiplicons;
        memcpy(sp->buf.as<uint8_t>() + res_off[i], &nm, 4);
        r->shared = sp;
        r->shared_off = res_off[i];
        r->n_bytes = res_off[i + 1] - res_off[i];
        r->stats.d2h_bytes = (int64_t)r->n_bytes - 4;
        all[i] = r;
    }
    CU(cudaEventRecord(c->ev_copied_recv, c->st_copy));
    CU(cudaEventCreateWithFlags(&sp->ready, cudaEventDisableTiming));
    CU(cudaEventRecord(sp->ready, c->st_copy));
    c->copy_pending_recv = true;
    return IADH_OK;
}

int iadh_event_record(iadh_ctx* c, int slot)
{
    if (!c || slot < 0 || slot >= 4) return fail(IADH_ERR_ARG, "iadh_event_record: bad argument");
    CU(cudaSetDevice(c->device));
    CU(cudaEventRecord(c->user_ev[slot], c->st));
    return IADH_OK;
}
int iadh_event_elapsed_ms(iadh_ctx* c, int from_slot, int to_slot, double* ms)
{
    if (!c || !ms || from_slot < 0 || from_slot >= 4 || to_slot < 0 || to_slot >= 4)
        return fail(IADH_ERR_ARG, "iadh_event_elapsed_ms: bad argument");
    CU(cudaSetDevice(c->device));
    CU(cudaEventSynchronize(c->user_ev[to_slot]));
    float f = 0;
    CU(cudaEventElapsedTime(&f, c->user_ev[from_slot], c->user_ev[to_slot]));
    *ms = f;
    return IADH_OK;
}

static int build_dots_impl(iadh_ctx* c, const iadh_profile_desc* p, const int32_t* xs, const int32_t* ys, int64_t n,
                    const int64_t** cnt_out, const int32_t** dots_out, int64_t* n_dots_out);
int iadh_build_dots(iadh_ctx* c, const iadh_profile_desc* p, const int32_t* xs, const int32_t* ys, int64_t n,
                    const int64_t** cnt_out, const int32_t** dots_out, int64_t* n_dots_out)
{
    IADH_TRY return build_dots_impl(c, p, xs, ys, n, cnt_out, dots_out, n_dots_out);
    IADH_CATCH
}
static int build_dots_impl(iadh_ctx* c, const iadh_profile_desc* p, const int32_t* xs, const int32_t* ys, int64_t n,
                    const int64_t** cnt_out, const int32_t** dots_out, int64_t* n_dots_out)
{
    if (!c || !ys || !cnt_out || !dots_out || !n_dots_out || (!p && !xs) || n < 0) return fail(IADH_ERR_ARG, "iadh_build_dots: bad argument");
    CU(cudaSetDevice(c->device));
    ProfileHost ph;
    memset(&ph.dev, 0, sizeof ph.dev);
    int64_t h2d = 0;
    int rc;
    if (p && (rc = upload_profile(c, p, ph, h2d))) return rc;
    std::vector<GhmTask> tasks((size_t)n);
    for (int64_t i = 0; i < n; i++) {
        init_task(tasks[i]);
        tasks[i].x_list = p ? -1 : xs[i];
        tasks[i].y_list = ys[i];
        if (ys[i] < 0 || ys[i] >= c->n_lists || (!p && (xs[i] < 0 || xs[i] >= c->n_lists)))
            return fail(IADH_ERR_ARG, "build_dots: list out of range");
        tasks[i].seq = (int32_t)i;
    }
    iadh_stats st;
    memset(&st, 0, sizeof st);
    int64_t total = 0;
    if (n > 0 && (rc = build_stage(c, p != nullptr, ph.dev, tasks, total, st))) return rc;
    std::vector<uint32_t> keys((size_t)total);
    if (total) CU(cudaMemcpyAsync(keys.data(), c->d_dots.p, 4 * (size_t)total, cudaMemcpyDeviceToHost, c->st));
    CU(cudaStreamSynchronize(c->st));
    c->bd_cnt.resize(2 * (size_t)n);
    c->bd_dots.resize(2 * (size_t)total);
    for (int64_t i = 0; i < n; i++) {
        c->bd_cnt[2 * i] = tasks[i].cnt[0];
        c->bd_cnt[2 * i + 1] = tasks[i].cnt[1];
    }
    for (int64_t i = 0; i < total; i++) {
        c->bd_dots[2 * i] = KEY_X(keys[i]);
        c->bd_dots[2 * i + 1] = KEY_Y(keys[i]);
    }
    *cnt_out = c->bd_cnt.data();
    *dots_out = c->bd_dots.data();
    *n_dots_out = total;
    return IADH_OK;
}

static int build_only_impl(iadh_ctx* c, const iadh_profile_desc* p, const int32_t* xs, const int32_t* ys, int64_t n,
                    iadh_stats* out);
int iadh_build_only(iadh_ctx* c, const iadh_profile_desc* p, const int32_t* xs, const int32_t* ys, int64_t n,
                    iadh_stats* out)
{
    IADH_TRY return build_only_impl(c, p, xs, ys, n, out);
    IADH_CATCH
}
static int build_only_impl(iadh_ctx* c, const iadh_profile_desc* p, const int32_t* xs, const int32_t* ys, int64_t n,
                    iadh_stats* out)
{
    if (!c || !ys || !out || (!p && !xs) || n < 0) return fail(IADH_ERR_ARG, "iadh_build_only: bad argument");
    CU(cudaSetDevice(c->device));
    ProfileHost ph;
    memset(&ph.dev, 0, sizeof ph.dev);
    memset(out, 0, sizeof *out);
    int rc;
    if (p && (rc = upload_profile(c, p, ph, out->h2d_bytes))) return rc;
    std::vector<GhmTask> tasks((size_t)n);
    for (int64_t i = 0; i < n; i++) {
        init_task(tasks[i]);
        tasks[i].x_list = p ? -1 : xs[i];
        tasks[i].y_list = ys[i];
        if (ys[i] < 0 || ys[i] >= c->n_lists || (!p && (xs[i] < 0 || xs[i] >= c->n_lists)))
            return fail(IADH_ERR_ARG, "build_only: list out of range");
        tasks[i].seq = (int32_t)i;
    }
    int64_t total = 0;
    if (n > 0 && (rc = build_stage(c, p != nullptr, ph.dev, tasks, total, *out))) return rc;
    CU(cudaStreamSynchronize(c->st));
    out->n_ghm = n;
    if (n > 0) {
        float m1 = 0, m2 = 0;
        cudaEventElapsedTime(&m1, c->ev[1], c->ev[2]);
        cudaEventElapsedTime(&m2, c->ev[3], c->ev[4]);
        out->build_ms = m1 + m2;
        out->cluster_ms = m2;  // (fill pass alone)
        out->total_ms = m1 + m2;
    }
    return IADH_OK;
}

static int remap_tandems_impl(const iadh_dataset_desc* d, int32_t tandem_gap, int device, uint8_t* remapped_out);
int iadh_remap_tandems(const iadh_dataset_desc* d, int32_t tandem_gap, int device, uint8_t* remapped_out)
{
    IADH_TRY return remap_tandems_impl(d, tandem_gap, device, remapped_out);
    IADH_CATCH
}
static int remap_tandems_impl(const iadh_dataset_desc* d, int32_t tandem_gap, int device, uint8_t* remapped_out)
{
    if (!d || !remapped_out || d->n_lists < 0 || tandem_gap < 0 || (d->n_lists > 0 && (!d->list_off || !d->fam || !d->gid || !d->flags)))
        return fail(IADH_ERR_ARG, "iadh_remap_tandems: bad argument");
    const int64_t total = d->n_lists ? d->list_off[d->n_lists] : 0;
    if (total == 0) return IADH_OK;
    if (total >= 0x7fffffff) return fail(IADH_ERR_CAPACITY, "more than 2^31 elements");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(IADH_ERR_CUDA, "no CUDA device available; this library has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(IADH_ERR_ARG, "device %d out of range (%d devices)", device, ndev);
    CU(cudaSetDevice(device));
    std::vector<int32_t> elem_list((size_t)total);
    for (int l = 0; l < d->n_lists; l++) {
        if (d->list_off[l + 1] < d->list_off[l]) return fail(IADH_ERR_ARG, "list %d has negative size", l);
        for (int64_t i = d->list_off[l]; i < d->list_off[l + 1]; i++) elem_list[i] = l;
    }
    cudaStream_t st;
    CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    DBuf b_off, b_el, b_fam, b_gid, b_fl, b_rem, b_t3;
    int rc = IADH_OK;
    if (!rc) rc = upload(b_off, d->list_off, 8 * (size_t)(d->n_lists + 1), st);
    if (!rc) rc = upload(b_el, elem_list.data(), 4 * (size_t)total, st);
    if (!rc) rc = upload(b_fam, d->fam, 4 * (size_t)total, st);
    if (!rc) rc = upload(b_gid, d->gid, 4 * (size_t)total, st);
    if (!rc) rc = upload(b_fl, d->flags, (size_t)total, st);
    if (!rc && b_rem.ensure((size_t)total) != cudaSuccess) rc = fail(IADH_ERR_CUDA, "out of device memory");
    if (!rc && b_t3.ensure(4 * (size_t)total) != cudaSuccess) rc = fail(IADH_ERR_CUDA, "out of device memory");
    if (!rc) {
        launch_tandem(b_off.as<int64_t>(), d->n_lists, b_el.as<int32_t>(), b_fam.as<int32_t>(), b_gid.as<int32_t>(),
                      b_fl.as<uint8_t>(), tandem_gap, total, b_rem.as<uint8_t>(), b_t3.as<int32_t>(), st);
        if (cudaMemcpyAsync(remapped_out, b_rem.p, (size_t)total, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaStreamSynchronize(st) != cudaSuccess)
            rc = fail(IADH_ERR_CUDA, "tandem remapping failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    cudaStreamSynchronize(st);
    DBuf* bufs[] = {&b_off, &b_el, &b_fam, &b_gid, &b_fl, &b_rem, &b_t3};
    for (DBuf* b : bufs) b->release();
    cudaStreamDestroy(st);
    return rc;
}

static int intern_strings_impl(const char* blob, const int64_t* off, int64_t n, int device, int32_t* ids_out, int32_t* n_distinct_out);
int iadh_intern_strings(const char* blob, const int64_t* off, int64_t n, int device, int32_t* ids_out, int32_t* n_distinct_out)
{
    IADH_TRY return intern_strings_impl(blob, off, n, device, ids_out, n_distinct_out);
    IADH_CATCH
}
static int intern_strings_impl(const char* blob, const int64_t* off, int64_t n, int device, int32_t* ids_out, int32_t* n_distinct_out)
{
    if (n < 0 || (n > 0 && (!off || !ids_out)) || !n_distinct_out) return fail(IADH_ERR_ARG, "iadh_intern_strings: bad argument");
    *n_distinct_out = 0;
    if (n == 0) return IADH_OK;
    if (n >= 0x3fffffff) return fail(IADH_ERR_CAPACITY, "more than 2^30 strings");
    if (off[0] != 0) return fail(IADH_ERR_ARG, "iadh_intern_strings: off[0] must be 0");
    for (int64_t i = 0; i < n; i++)
        if (off[i + 1] < off[i]) return fail(IADH_ERR_ARG, "iadh_intern_strings: offsets are not ascending at string %lld", (long long)i);
    if (off[n] > 0 && !blob) return fail(IADH_ERR_ARG, "iadh_intern_strings: null blob");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(IADH_ERR_CUDA, "no CUDA device available; this library has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(IADH_ERR_ARG, "device %d out of range (%d devices)", device, ndev);
    CU(cudaSetDevice(device));
    uint32_t tsize = 1024;
    while ((int64_t)tsize < 2 * n) tsize <<= 1;   // load factor <= 0.5
    const size_t n_blocks = (size_t)((n + 1023) / 1024);
    cudaStream_t st;
    CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    DBuf b_blob, b_off, b_table, b_rep, b_cmin, b_scan, b_bsum, b_ids;
    int rc = IADH_OK;
    int32_t total = 0;
    if (!rc) rc = upload(b_blob, blob, (size_t)off[n], st);
    if (!rc) rc = upload(b_off, off, 8 * (size_t)(n + 1), st);
    if (!rc && (b_table.ensure(4 * (size_t)tsize) != cudaSuccess || b_rep.ensure(4 * (size_t)n) != cudaSuccess ||
                b_cmin.ensure(4 * (size_t)n) != cudaSuccess || b_scan.ensure(4 * (size_t)n) != cudaSuccess ||
                b_bsum.ensure(4 * n_blocks + 16) != cudaSuccess || b_ids.ensure(4 * (size_t)n) != cudaSuccess))
        rc = fail(IADH_ERR_CUDA, "out of device memory");
    if (!rc) {
        cudaMemsetAsync(b_table.p, 0xff, 4 * (size_t)tsize, st);   // -1: empty slot
        cudaMemsetAsync(b_cmin.p, 0x7f, 4 * (size_t)n, st);        // 0x7f7f7f7f: above every index
        int32_t* d_total = b_bsum.as<int32_t>() + n_blocks;
        launch_intern(b_blob.as<char>(), b_off.as<int64_t>(), n, b_table.as<int32_t>(), tsize, b_rep.as<int32_t>(), b_cmin.as<int32_t>(),
                      b_scan.as<int32_t>(), b_bsum.as<int32_t>(), d_total, b_ids.as<int32_t>(), st);
        if (cudaMemcpyAsync(ids_out, b_ids.p, 4 * (size_t)n, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaMemcpyAsync(&total, d_total, 4, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess)
            rc = fail(IADH_ERR_CUDA, "string interning failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    cudaStreamSynchronize(st);
    DBuf* bufs[] = {&b_blob, &b_off, &b_table, &b_rep, &b_cmin, &b_scan, &b_bsum, &b_ids};
    for (DBuf* b : bufs) b->release();
    cudaStreamDestroy(st);
    if (!rc) *n_distinct_out = total;
    return rc;
}

static int remap_tandems_pairs_impl(const iadh_dataset_desc* d, const int64_t* pair_ptr, const int32_t* pair_ids, int32_t n_genes,
                                    int32_t tandem_gap, int device, uint8_t* remapped_out);
int iadh_remap_tandems_pairs(const iadh_dataset_desc* d, const int64_t* pair_ptr, const int32_t* pair_ids, int32_t n_genes,
                             int32_t tandem_gap, int device, uint8_t* remapped_out)
{
    IADH_TRY return remap_tandems_pairs_impl(d, pair_ptr, pair_ids, n_genes, tandem_gap, device, remapped_out);
    IADH_CATCH
}
static int remap_tandems_pairs_impl(const iadh_dataset_desc* d, const int64_t* pair_ptr, const int32_t* pair_ids, int32_t n_genes,
                                    int32_t tandem_gap, int device, uint8_t* remapped_out)
{
    if (!d || !remapped_out || !pair_ptr || n_genes < 0 || d->n_lists < 0 || tandem_gap < 0 ||
        (d->n_lists > 0 && (!d->list_off || !d->gid || !d->flags)))
        return fail(IADH_ERR_ARG, "iadh_remap_tandems_pairs: bad argument");
    if (pair_ptr[0] != 0 || (pair_ptr[n_genes] > 0 && !pair_ids)) return fail(IADH_ERR_ARG, "iadh_remap_tandems_pairs: bad pair table");
    for (int32_t g = 0; g < n_genes; g++)
        if (pair_ptr[g + 1] < pair_ptr[g]) return fail(IADH_ERR_ARG, "iadh_remap_tandems_pairs: pair_ptr is not ascending at gene %d", g);
    const int64_t total = d->n_lists ? d->list_off[d->n_lists] : 0;
    if (total == 0) return IADH_OK;
    if (total >= 0x7fffffff) return fail(IADH_ERR_CAPACITY, "more than 2^31 elements");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(IADH_ERR_CUDA, "no CUDA device available; this library has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(IADH_ERR_ARG, "device %d out of range (%d devices)", device, ndev);
    CU(cudaSetDevice(device));
    std::vector<int32_t> elem_list((size_t)total);
    for (int l = 0; l < d->n_lists; l++) {
        if (d->list_off[l + 1] < d->list_off[l]) return fail(IADH_ERR_ARG, "list %d has negative size", l);
        for (int64_t i = d->list_off[l]; i < d->list_off[l + 1]; i++) elem_list[i] = l;
    }
    cudaStream_t st;
    CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    DBuf b_off, b_el, b_gid, b_fl, b_pp, b_pi, b_rem, b_t3;
    int rc = IADH_OK;
    if (!rc) rc = upload(b_off, d->list_off, 8 * (size_t)(d->n_lists + 1), st);
    if (!rc) rc = upload(b_el, elem_list.data(), 4 * (size_t)total, st);
    if (!rc) rc = upload(b_gid, d->gid, 4 * (size_t)total, st);
    if (!rc) rc = upload(b_fl, d->flags, (size_t)total, st);
    if (!rc) rc = upload(b_pp, pair_ptr, 8 * (size_t)(n_genes + 1), st);
    if (!rc) rc = upload(b_pi, pair_ids, 4 * (size_t)pair_ptr[n_genes], st);
    if (!rc && b_rem.ensure((size_t)total) != cudaSuccess) rc = fail(IADH_ERR_CUDA, "out of device memory");
    if (!rc && b_t3.ensure(4 * (size_t)total) != cudaSuccess) rc = fail(IADH_ERR_CUDA, "out of device memory");
    if (!rc) {
        launch_tandem_pairs(b_off.as<int64_t>(), d->n_lists, b_el.as<int32_t>(), b_gid.as<int32_t>(), b_fl.as<uint8_t>(), b_pp.as<int64_t>(),
                            b_pi.as<int32_t>(), n_genes, tandem_gap, total, b_rem.as<uint8_t>(), b_t3.as<int32_t>(), st);
        if (cudaMemcpyAsync(remapped_out, b_rem.p, (size_t)total, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaStreamSynchronize(st) != cudaSuccess)
            rc = fail(IADH_ERR_CUDA, "tandem remapping failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    cudaStreamSynchronize(st);
    DBuf* bufs[] = {&b_off, &b_el, &b_gid, &b_fl, &b_pp, &b_pi, &b_rem, &b_t3};
    for (DBuf* b : bufs) b->release();
    cudaStreamDestroy(st);
    return rc;
}

static int plan_lists_impl(const int64_t* sizes, int32_t n_lists, int rank, int n_ranks, int32_t* out, int64_t* n_out);
int iadh_plan_lists(const int64_t* sizes, int32_t n_lists, int rank, int n_ranks, int32_t* out, int64_t* n_out)
{
    IADH_TRY return plan_lists_impl(sizes, n_lists, rank, n_ranks, out, n_out);
    IADH_CATCH
}
static int plan_lists_impl(const int64_t* sizes, int32_t n_lists, int rank, int n_ranks, int32_t* out, int64_t* n_out)
{
    if (!sizes || !out || !n_out || n_lists < 0 || n_ranks < 1 || rank < 0 || rank >= n_ranks)
        return fail(IADH_ERR_ARG, "iadh_plan_lists: bad argument");
    std::vector<int64_t> sz(sizes, sizes + n_lists);
    std::vector<int32_t> itl, res;
    sort_by_size(sz, itl);
    plan_lists(sz, itl, rank, n_ranks, res);
    std::copy(res.begin(), res.end(), out);
    *n_out = (int64_t)res.size();
    return IADH_OK;
}

static int plan_pairs_impl(const int64_t* sizes, int32_t n_lists, int rank, int n_ranks, int32_t* out_x, int32_t* out_y,
                    int64_t* n_out);
int iadh_plan_pairs(const int64_t* sizes, int32_t n_lists, int rank, int n_ranks, int32_t* out_x, int32_t* out_y,
                    int64_t* n_out)
{
    IADH_TRY return plan_pairs_impl(sizes, n_lists, rank, n_ranks, out_x, out_y, n_out);
    IADH_CATCH
}
static int plan_pairs_impl(const int64_t* sizes, int32_t n_lists, int rank, int n_ranks, int32_t* out_x, int32_t* out_y,
                    int64_t* n_out)
{
    if (!sizes || !out_x || !out_y || !n_out || n_lists < 0 || n_ranks < 1 || rank < 0 || rank >= n_ranks)
        return fail(IADH_ERR_ARG, "iadh_plan_pairs: bad argument");
    std::vector<int64_t> sz(sizes, sizes + n_lists);
    std::vector<int32_t> itl, xs, ys;
    sort_by_size(sz, itl);
    plan_pairs(sz, itl, rank, n_ranks, xs, ys);
    std::copy(xs.begin(), xs.end(), out_x);
    std::copy(ys.begin(), ys.end(), out_y);
    *n_out = (int64_t)xs.size();
    return IADH_OK;
}

static int math_hook(iadh_ctx* c, const double* x, const int32_t* n, int64_t count, double* out, int mode);
int iadh_omxn(iadh_ctx* c, const double* x, const int32_t* n, int64_t count, double* out)
{
    return math_hook(c, x, n, count, out, 0);
}
int iadh_pow(iadh_ctx* c, const double* x, const int32_t* n, int64_t count, double* out)
{
    return math_hook(c, x, n, count, out, 1);
}
static int math_hook(iadh_ctx* c, const double* x, const int32_t* n, int64_t count, double* out, int mode)
{
    if (!c || !x || !n || !out || count < 0) return fail(IADH_ERR_ARG, "iadh_omxn/iadh_pow: bad argument");
    if (count == 0) return IADH_OK;
    CU(cudaSetDevice(c->device));
    // (three cached blocks, released on every path)
    struct Tmp {
        DBuf x, out, n;
        ~Tmp()
        {
            x.release();
            out.release();
            n.release();
        }
    } t;
    CU(t.x.ensure(8 * (size_t)count));
    CU(t.out.ensure(8 * (size_t)count));
    CU(t.n.ensure(4 * (size_t)count));
    CU(cudaMemcpyAsync(t.x.p, x, 8 * (size_t)count, cudaMemcpyHostToDevice, c->st));
    CU(cudaMemcpyAsync(t.n.p, n, 4 * (size_t)count, cudaMemcpyHostToDevice, c->st));
    omxn_kernel<<<(unsigned)((count + 255) / 256), 256, 0, c->st>>>(t.x.as<double>(), t.n.as<int>(), count, t.out.as<double>(), mode);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, t.out.p, 8 * (size_t)count, cudaMemcpyDeviceToHost, c->st));
    CU(cudaStreamSynchronize(c->st));
    return IADH_OK;
}

}  // extern "C"
