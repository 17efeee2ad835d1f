// TEST INFRASTRUCTURE: runs the phase-1 clustering kernels of cluster_kernel.cu on the CPU (cuda_emu.h) for
// a set of GHMs, once with the default join pass and once with the experimental two-phase join
// (phase1_join2_kernel, IADH_JOIN2), and compares the complete clustering state after the last gap size:
// cluster vectors, every cluster record field by field, anchor points and flags, and what is left of the
// two matrices.  The default join is validated on the GPU against the reference; equality here validates
// the CUDA source of the new kernel in a container without a GPU.
//
// input (little endian): i32 n_gaps, i32 gaps[10], f64 q, i32 n_ghm; per GHM: u32 cnt0, cnt1, u64 area,
// i32 level, u32 keys[cnt0 + cnt1] ((x << 16) | y, orientation 0 block first, each in (x, y) order)
#define IADH_HOST_EMU 1
#define EMU_NT 32
#include "cuda_emu.h"

#include "../../i-adhore_b200/csrc/cluster_kernel.cu"

#include "emu_runtime.inc"

struct GhmIn {
    uint32_t cnt[2];
    uint64_t area;
    int level;
    std::vector<uint32_t> keys;
};

static void dump_state(const GhmTask& task, const uint8_t* ws, int pool_factor, std::vector<uint8_t>& out)
{
    const int D = (int)(task.cnt[0] + task.cnt[1]);
    const WsLayout L = ws_layout(D, pool_factor);
    const ClusterRec* cl = (const ClusterRec*)(ws + L.cl);
    const uint32_t* pool = (const uint32_t*)(ws + L.pool);
    const uint8_t* pflag = ws + L.pflag;
    auto put = [&out](const void* p, size_t n) { out.insert(out.end(), (const uint8_t*)p, (const uint8_t*)p + n); };
    for (int o = 0; o < 2; o++) {
        const int nbc = task.half_nbc[o];
        put(&nbc, 4);
        put(&task.half_status[o], 4);
        const int* bc = (const int*)(ws + (o ? L.bc1 : L.bc0));
        for (int i = 0; i < nbc; i++) {
            const ClusterRec& c = cl[bc[i]];
            put(&c.a, 8 * 10);            // a, b, avg_x, var_x, mrss, x_end1, x_end2, y_end1, y_end2, p
            put(&c.sx, 8 * 5);
            put(&c.n, 4);
            put(&c.orient, 1);
            put(pool + c.off, 4 * (size_t)c.n);
            put(pflag + c.off, (size_t)c.n);
        }
        const uint32_t* key = (const uint32_t*)(ws + (o ? L.key1 : L.key0));
        const uint8_t* alive = ws + (o ? L.alive1 : L.alive0);
        for (uint32_t i = 0; i < task.cnt[o]; i++)
            if (alive[i]) put(&key[i], 4);
    }
}

static bool run_variant(const GhmIn& in, const DevParams& prm, int pool_factor, bool join2, std::vector<uint8_t>& state,
                        long long& merges)
{
    const int D = (int)(in.cnt[0] + in.cnt[1]);
    const WsLayout L = ws_layout(D, pool_factor);
    std::vector<uint8_t> ws(L.total + 256, 0);
    GhmTask task;
    memset(&task, 0, sizeof task);
    task.x_list = 0; task.y_list = 1; task.level = in.level;
    task.cnt[0] = in.cnt[0]; task.cnt[1] = in.cnt[1];
    task.area = in.area;
    std::vector<int> items;
    for (int o = 0; o < 2; o++)
        if (in.cnt[o] >= 3) { items.push_back(o); task.half_mask |= 1 << o; }
    task.half_needed = (int)items.size();
    std::vector<int> sm_slot(1024, 0);
    unsigned long long dbg[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint32_t istate[2] = {0, 0};
    if (items.empty()) return true;
    for (int gi = 0; gi < prm.n_gaps; gi++)
        for (int pass = 0; pass < 3; pass++) {
            int next_item = 0;
            if (pass == 2 && join2)
                emu::run_block(0, EMU_NT, [&]() { phase1_join2_kernel<32>(&task, items.data(), (int)items.size(), &next_item, ws.data(), prm, pool_factor, gi, dbg, istate); });
            else
                emu::run_block(0, EMU_NT, [&]() { phase1_pass_kernel<32, 32>(&task, items.data(), (int)items.size(), &next_item, in.keys.data(), ws.data(), prm, pool_factor, sm_slot.data(), pass, gi, dbg, istate); });
            if (getenv("EMU_TRACE")) {
                const ClusterRec* cl = (const ClusterRec*)(ws.data() + L.cl);
                for (int o = 0; o < 2; o++) {
                    const int* bc = (const int*)(ws.data() + (o ? L.bc1 : L.bc0));
                    long long aps = 0;
                    for (int i = 0; i < task.half_nbc[o]; i++) aps += cl[bc[i]].n;
                    int alive = 0;
                    const uint8_t* al = ws.data() + (o ? L.alive1 : L.alive0);
                    for (uint32_t i = 0; i < in.cnt[o]; i++) alive += al[i];
                    fprintf(stderr, "  gap %d pass %d o %d: %d clusters, %lld APs, %d dots alive of %u, status %d\n", prm.gaps[gi], pass, o,
                            task.half_nbc[o], aps, alive, in.cnt[o], task.half_status[o]);
                }
            }
        }
    merges = (long long)dbg[7];
    (void)getenv;
    if (task.half_status[0] || task.half_status[1]) return false;   // pool exhausted: caller retries with a larger pool
    dump_state(task, ws.data(), pool_factor, state);
    // EMU_DUMP=<file>: the same text the instrumented oracle writes with ORACLE_STUDY_DUMP (fidelity check
    // of the emulation itself: the default kernels, emulated, against the oracle's state after the gap loop)
    if (!join2)
        if (const char* path = getenv("EMU_DUMP")) {
            FILE* f = fopen(path, "a");
            const ClusterRec* cl = (const ClusterRec*)(ws.data() + L.cl);
            const uint32_t* pool = (const uint32_t*)(ws.data() + L.pool);
            fprintf(f, "GHM %u %u\n", in.cnt[0], in.cnt[1]);
            for (int o = 0; o < 2; o++) {
                const int* bc = (const int*)(ws.data() + (o ? L.bc1 : L.bc0));
                fprintf(f, " o%d clusters %d\n", o, task.half_nbc[o]);
                for (int i = 0; i < task.half_nbc[o]; i++) {
                    const ClusterRec& c = cl[bc[i]];
                    fprintf(f, "  c %d:", c.n);
                    for (int k = 0; k < c.n; k++) fprintf(f, " %d,%d", KEY_X(pool[c.off + k]), KEY_Y(pool[c.off + k]));
                    fprintf(f, "\n");
                }
                // (an orientation with fewer than 3 dots has no queue item: its matrix is never loaded)
                const bool loaded = in.cnt[o] >= 3;
                const uint32_t* key = loaded ? (const uint32_t*)(ws.data() + (o ? L.key1 : L.key0)) : in.keys.data() + (o ? in.cnt[0] : 0);
                const uint8_t* alive = ws.data() + (o ? L.alive1 : L.alive0);
                int left = 0;
                for (uint32_t i = 0; i < in.cnt[o]; i++) left += loaded ? alive[i] : 1;
                fprintf(f, "  left %d:", left);
                for (uint32_t i = 0; i < in.cnt[o]; i++)
                    if (!loaded || alive[i]) fprintf(f, " %d,%d", KEY_X(key[i]), KEY_Y(key[i]));
                fprintf(f, "\n");
            }
            fclose(f);
        }
    return true;
}

int main(int argc, char** argv)
{
    if (argc < 2) { fprintf(stderr, "usage: %s <ghms.bin>\n", argv[0]); return 2; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    DevParams prm;
    memset(&prm, 0, sizeof prm);
    int n_ghm = 0;
    if (fread(&prm.n_gaps, 4, 1, f) != 1 || fread(prm.gaps, 4, 10, f) != 10 || fread(&prm.q_value, 8, 1, f) != 1 ||
        fread(&n_ghm, 4, 1, f) != 1) return 2;
    prm.gap_size = prm.gaps[prm.n_gaps - 1];
    prm.cluster_gap = prm.gap_size + 5;
    prm.anchor_points = 3;
    prm.mult_hyp = 1;
    prm.prob_cutoff = 0.01;
    const int coop[4] = {16, 16, 16, 48};
    memcpy(prm.coop, coop, sizeof coop);
    long long bad = 0, total_merges = 0, total_clusters = 0;
    for (int g = 0; g < n_ghm; g++) {
        GhmIn in;
        if (fread(in.cnt, 4, 2, f) != 2 || fread(&in.area, 8, 1, f) != 1 || fread(&in.level, 4, 1, f) != 1) return 2;
        in.keys.resize(in.cnt[0] + in.cnt[1]);
        if (!in.keys.empty() && fread(in.keys.data(), 4, in.keys.size(), f) != in.keys.size()) return 2;
        for (int pool_factor = 8; pool_factor <= 512; pool_factor *= 4) {
            std::vector<uint8_t> a, b;
            long long ma = 0, mb = 0;
            const bool oka = run_variant(in, prm, pool_factor, false, a, ma);
            const bool okb = run_variant(in, prm, pool_factor, true, b, mb);
            if (!oka || !okb) continue;
            total_merges += ma;
            if (a.size() >= 4) { int n0; memcpy(&n0, a.data(), 4); total_clusters += n0; }
            if (a != b || ma != mb) {
                bad++;
                fprintf(stderr, "GHM %d (D = %u + %u): states differ (%zu vs %zu bytes, merges %lld vs %lld)\n", g, in.cnt[0], in.cnt[1],
                        a.size(), b.size(), ma, mb);
            }
            break;
        }
    }
    fclose(f);
    printf("{\"ghms\": %d, \"merges\": %lld, \"clusters_o0\": %lld, \"mismatching_ghms\": %lld, \"collectives\": %llu}\n", n_ghm, total_merges,
           total_clusters, bad, emu::n_collectives);
    return bad ? 1 : 0;
}
