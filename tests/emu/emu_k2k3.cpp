// TEST INFRASTRUCTURE: the fused clustering kernel (cluster_kernel<32>: phase 1, phase 2, K3 measuring pass)
// and the K3 writer (pack_write_kernel) compiled for the host and run under the one-warp emulation of
// cuda_emu.h, GHM by GHM, on the dot sets of a reference trace; the packed multiplicons must equal the
// reference's bytes (Multiplicon::packMultiplicons of the same GHM), p-values included.
//
// input (little endian): i32 n_gaps, gaps[10], gap_size, cluster_gap, anchor_points, mult_hyp; f64 q, prob_cutoff;
//   i32 n_lists; i64 list_off[n_lists+1]; i32 fam[T], gid[T], num_id[T]; u8 flags[T]   (T = list_off[n_lists])
//   i32 n_profiles; per profile: i32 S, L; i32 fam[S*L], gid[S*L], num_id[S*L]; u8 flags[S*L]
//   i32 n_ghm; per GHM: i32 profile (-1 = level 2), x_list, y_list, x_object_id, level; u64 area; u32 cnt0, cnt1;
//                       u32 keys[cnt0+cnt1]; i32 n_expected; u8 expected[n_expected]
//   (expected = the reference's packed blob of the GHM: i32 multiplicon count, then the multiplicons)
// Format flags in bits 8+ of the first word (n_gaps):
//   0x100  homolog-count filter (homolog_kernel.cu): every profile is followed by i32 n_links, n_links x {i32 segment_x,
//          segment_y, gene_x_id, gene_y_id} (the Link set of its multiplicon) and `expected` of a profile GHM holds only
//          the multiplicons the reference's own search kept (GHMProfile.cpp:129-132).  The link tables are built by the
//          emulated link_resolve / link_count / link_fill kernels (the scan between them on the host), the filter
//          kernel runs between the clustering kernel and the writer, as in the engine.
//   0x200  table_type=pairs: the dataset is followed by i32 n_genes, i64 pair_ptr[n_genes + 1], i32 pair_ids[]; the
//          partner buckets of iadh_create_pairs are rebuilt here (bucket g = positions of the partners of gene g)
#define IADH_HOST_EMU 1
#ifndef EMU_NT
#define EMU_NT 32        // block size of the emulated clustering kernel: 32, 64, 128 or 256
#endif
#if EMU_NT == 32
#define IADH_EMU_PT 32   // (and a one-warp writer; the larger builds keep the writer's own 128 threads)
#endif
#include "cuda_emu.h"
#include <algorithm>

#include "../../i-adhore_b200/csrc/cluster_kernel.cu"
#include "../../i-adhore_b200/csrc/pack_kernel.cu"
#include "../../i-adhore_b200/csrc/homolog_kernel.cu"

#include "emu_runtime.inc"

template <typename T>
static bool rd(FILE* f, T* p, size_t n) { return n == 0 || fread(p, sizeof(T), n, f) == n; }

int main(int argc, char** argv)
{
    if (argc < 2) { fprintf(stderr, "usage: %s <trace_ghms.bin>\n", argv[0]); return 2; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    DevParams prm;
    memset(&prm, 0, sizeof prm);
    if (!rd(f, &prm.n_gaps, 1) || !rd(f, prm.gaps, 10) || !rd(f, &prm.gap_size, 1) || !rd(f, &prm.cluster_gap, 1) ||
        !rd(f, &prm.anchor_points, 1) || !rd(f, &prm.mult_hyp, 1) || !rd(f, &prm.q_value, 1) || !rd(f, &prm.prob_cutoff, 1)) return 2;
    const int fmt = prm.n_gaps >> 8;
    prm.n_gaps &= 0xff;
    const int coop[4] = {16, 16, 16, 48};
    memcpy(prm.coop, coop, sizeof coop);
    int n_lists = 0;
    if (!rd(f, &n_lists, 1)) return 2;
    std::vector<int64_t> list_off(n_lists + 1);
    if (!rd(f, list_off.data(), list_off.size())) return 2;
    const size_t T = (size_t)list_off[n_lists];
    std::vector<int32_t> fam(T), gid(T), num_id(T);
    std::vector<uint8_t> flags(T);
    if (!rd(f, fam.data(), T) || !rd(f, gid.data(), T) || !rd(f, num_id.data(), T) || !rd(f, flags.data(), T)) return 2;
    DatasetDev ds;
    memset(&ds, 0, sizeof ds);
    ds.n_lists = n_lists;
    ds.list_off = list_off.data();
    ds.fam = fam.data();
    ds.gid = gid.data();
    ds.num_id = num_id.data();
    ds.flags = flags.data();
    // the dataset's bucket CSR: families (members ascending), or in pairs mode gene id -> positions of its partners
    std::vector<int32_t> fam_ptr, fam_members, ds_fam_pairs;
    int n_fam = 0;
    if (fmt & 2) {
        int n_genes = 0;
        if (!rd(f, &n_genes, 1)) return 2;
        std::vector<int64_t> pptr(n_genes + 1);
        if (!rd(f, pptr.data(), pptr.size())) return 2;
        std::vector<int32_t> pids((size_t)pptr[n_genes]);
        if (!rd(f, pids.data(), pids.size())) return 2;
        std::vector<std::vector<int32_t>> pos(n_genes);
        for (size_t i = 0; i < T; i++)
            if (gid[i] >= 0 && gid[i] < n_genes) pos[gid[i]].push_back((int32_t)i);
        n_fam = n_genes;
        fam_ptr.assign(n_genes + 1, 0);
        for (int g = 0; g < n_genes; g++) {
            std::vector<int32_t> b;
            for (int64_t k = pptr[g]; k < pptr[g + 1]; k++) b.insert(b.end(), pos[pids[k]].begin(), pos[pids[k]].end());
            std::sort(b.begin(), b.end());
            fam_members.insert(fam_members.end(), b.begin(), b.end());
            fam_ptr[g + 1] = (int32_t)fam_members.size();
        }
        ds_fam_pairs = gid;
        ds.fam = ds_fam_pairs.data();   // "family" = gene id (iadh_create_pairs)
        ds.pairs = 1;
    } else {
        for (size_t i = 0; i < T; i++) n_fam = std::max(n_fam, fam[i] + 1);
        fam_ptr.assign(n_fam + 1, 0);
        for (size_t i = 0; i < T; i++) fam_ptr[fam[i] + 1]++;
        for (int k = 0; k < n_fam; k++) fam_ptr[k + 1] += fam_ptr[k];
        fam_members.resize(T);
        std::vector<int32_t> cur(fam_ptr.begin(), fam_ptr.end() - 1);
        for (size_t i = 0; i < T; i++) fam_members[cur[fam[i]]++] = (int32_t)i;
    }
    ds.fam_ptr = fam_ptr.data();
    ds.fam_members = fam_members.data();
    struct Prof {
        int S, L;
        std::vector<int32_t> fam, gid, num_id;
        std::vector<uint8_t> flags;
        std::vector<LinkIn> links;
        int32_t lk_base = 0;   // first element of the profile in the batch's link tables
    };
    std::vector<int32_t> all_lk_ptr, all_lk_col;
    int n_prof = 0;
    if (!rd(f, &n_prof, 1)) return 2;
    std::vector<Prof> profs(n_prof);
    for (Prof& p : profs) {
        if (!rd(f, &p.S, 1) || !rd(f, &p.L, 1)) return 2;
        const size_t n = (size_t)p.S * p.L;
        p.fam.resize(n); p.gid.resize(n); p.num_id.resize(n); p.flags.resize(n);
        if (!rd(f, p.fam.data(), n) || !rd(f, p.gid.data(), n) || !rd(f, p.num_id.data(), n) || !rd(f, p.flags.data(), n)) return 2;
        if (fmt & 2)   // bucket key = gene id; ids outside the pair table have no bucket (place_profile, engine.cu)
            for (size_t k = 0; k < n; k++) p.fam[k] = (p.gid[k] >= 0 && p.gid[k] < n_fam) ? p.gid[k] : -1;
        if (fmt & 1) {
            int nl = 0;
            if (!rd(f, &nl, 1)) return 2;
            std::vector<int32_t> raw(4 * (size_t)nl);
            if (!rd(f, raw.data(), raw.size())) return 2;
            p.links.resize(nl);
            for (int i = 0; i < nl; i++) {
                LinkIn& r = p.links[i];
                r.seg_x = raw[4 * i]; r.seg_y = raw[4 * i + 1]; r.gene_x = raw[4 * i + 2]; r.gene_y = raw[4 * i + 3];
            }
        }
    }
    if ((fmt & 1) && n_prof > 0) {
        // the link tables of ALL profiles in one go, as launch_link_tables builds them for a batch (homolog_kernel.cu):
        // hash, resolve, count, scan, fill over the concatenated Links / elements.  Every profile is followed by an
        // empty slot (a profile of the batch that another rank searches: no Links, no elements here).
        const int np2 = 2 * n_prof;
        std::vector<ProfileDev> pdev(np2);
        std::vector<int64_t> link_ptr(np2 + 1, 0);
        std::vector<int32_t> elem_base(np2 + 1, 0);
        std::vector<LinkIn> all;
        for (int k = 0; k < n_prof; k++) {
            Prof& p = profs[k];
            ProfileDev& pf = pdev[2 * k];
            memset(&pf, 0, sizeof pf);
            pf.S = p.S; pf.L = p.L;
            pf.fam = p.fam.data(); pf.gid = p.gid.data(); pf.num_id = p.num_id.data(); pf.flags = p.flags.data();
            memset(&pdev[2 * k + 1], 0, sizeof(ProfileDev));
            all.insert(all.end(), p.links.begin(), p.links.end());
            link_ptr[2 * k + 1] = link_ptr[2 * k + 2] = (int64_t)all.size();
            elem_base[2 * k + 1] = elem_base[2 * k + 2] = elem_base[2 * k] + p.S * p.L;
        }
        const int64_t nl = (int64_t)all.size(), n = elem_base[np2];
        ProfileDev pfc;
        memset(&pfc, 0, sizeof pfc);
        pfc.batch = pdev.data();
        uint32_t slots = 16;
        while ((int64_t)slots < 2 * n) slots <<= 1;
        std::vector<unsigned long long> keys(slots, ~0ull);
        std::vector<int32_t> vals(slots, -1);
        std::vector<int32_t> cols(2 * (size_t)nl + 1, -7), cnt((size_t)n + 1, 0);
        all_lk_ptr.assign((size_t)n + 1, 0);
        all_lk_col.assign(2 * (size_t)nl + 1, -7);
        for (int64_t b = 0; b * 256 < n; b++) emu::run_block((unsigned)b, 256, [&]() { link_hash_kernel(elem_base.data(), np2, n, pfc, keys.data(), vals.data(), slots - 1); });
        for (int64_t b = 0; b * 256 < 2 * nl; b++)
            emu::run_block((unsigned)b, 256, [&]() { link_resolve_kernel(all.data(), link_ptr.data(), elem_base.data(), np2, pfc, keys.data(), vals.data(), slots - 1, cols.data()); });
        for (int64_t b = 0; b * 256 < nl; b++)
            emu::run_block((unsigned)b, 256, [&]() { link_count_kernel(all.data(), link_ptr.data(), elem_base.data(), np2, pfc, cols.data(), cnt.data()); });
        for (int64_t k = 0; k < n; k++) all_lk_ptr[k + 1] = all_lk_ptr[k] + cnt[k];
        std::fill(cnt.begin(), cnt.end(), 0);
        for (int64_t b = 0; b * 256 < nl; b++)
            emu::run_block((unsigned)b, 256, [&]() { link_fill_kernel(all.data(), link_ptr.data(), elem_base.data(), np2, pfc, cols.data(), all_lk_ptr.data(), cnt.data(), all_lk_col.data()); });
        for (int k = 0; k < n_prof; k++) profs[k].lk_base = elem_base[2 * k];
    }
    int n_ghm = 0;
    if (!rd(f, &n_ghm, 1)) return 2;
    long long bad = 0, n_mult = 0, n_bytes = 0, n_resident = 0, n_fallback = 0, n_before = 0;
    for (int gidx = 0; gidx < n_ghm; gidx++) {
        int32_t pi, xl, yl, xobj, level, n_exp;
        uint64_t area;
        uint32_t cnt[2];
        if (!rd(f, &pi, 1) || !rd(f, &xl, 1) || !rd(f, &yl, 1) || !rd(f, &xobj, 1) || !rd(f, &level, 1) || !rd(f, &area, 1) || !rd(f, cnt, 2)) return 2;
        const bool hl = pi >= 0;
        ProfileDev pf;
        memset(&pf, 0, sizeof pf);
        if (hl) {
            const Prof& p = profs[pi];
            pf.S = p.S; pf.L = p.L;
            pf.fam = p.fam.data(); pf.gid = p.gid.data(); pf.num_id = p.num_id.data(); pf.flags = p.flags.data();
        }
        std::vector<uint32_t> keys(cnt[0] + cnt[1]);
        if (!rd(f, keys.data(), keys.size()) || !rd(f, &n_exp, 1)) return 2;
        std::vector<uint8_t> expected(n_exp);
        if (!rd(f, expected.data(), expected.size())) return 2;
        std::vector<uint8_t> got(4, 0);
        const int D = (int)(cnt[0] + cnt[1]);
        if (D >= prm.anchor_points && D > 0) {
            for (int pool_factor = 8; pool_factor <= 2048; pool_factor *= 4) {
                const WsLayout L = ws_layout(D, pool_factor);
                std::vector<uint8_t> ws(L.total + 256, 0);
                GhmTask task;
                memset(&task, 0, sizeof task);
                task.x_list = xl; task.y_list = yl; task.x_object_id = xobj; task.y_object_id = yl; task.level = level;
                task.cnt[0] = cnt[0]; task.cnt[1] = cnt[1];
                task.area = area;
                task.half_mask = (cnt[0] >= 3 ? 1 : 0) | (cnt[1] >= 3 ? 2 : 0);
                task.half_needed = 1;
                GhmProf prof;
                memset(&prof, 0, sizeof prof);
                const int item = 0 * 4 + task.half_mask;
                int next = 0;
                std::vector<int> sm_slot(1024, 0);
                // EMU_RES=1: the resident kernel (the GHM's state in "shared memory": a host buffer of exactly
                // RsLayout::total bytes, so an overrun of a budget shows up under a memory checker); a GHM that
                // outgrows a budget (status != 0) falls through to the global-path kernel below, like the engine
                const char* rm = getenv("EMU_RES");
                bool done_resident = false;
                if (rm && atoi(rm) > 0) {
                    const RsLayout RL = rs_layout((int)cnt[0], (int)cnt[1], EMU_NT);
                    std::vector<uint8_t> smem(RL.total + 16, 0xcd);
                    rs_smem = smem.data();
                    const int order0 = 0;
                    constexpr int MB = EMU_NT == 32 ? 8 : EMU_NT == 64 ? 4 : EMU_NT == 128 ? 2 : 1;
                    emu::run_block(0, EMU_NT, [&]() { resident_kernel<EMU_NT, MB>(&task, &order0, 1, &next, keys.data(), ws.data(), prm, pool_factor, (int)RL.total, ds, pf, hl ? 1 : 0, nullptr); });
                    rs_smem = nullptr;
                    n_resident++;
                    if (task.status == 0) done_resident = true;
                    else {
                        n_fallback++;
                        task.status = 0; task.n_mult = 0; task.out_bytes = 0; task.n_clusters = 0; task.n_aps = 0;
                        next = 0;
                    }
                }
                if (!done_resident) {
#if EMU_NT == 32
                // EMU_PASS=1: the pass-per-launch path instead of the fused kernel (one launch per pass and gap
                // size over the (GHM, orientation) items, then the four phase-2 stages); EMU_PASS=2: the same
                // with the experimental two-phase join kernel (IADH_JOIN2)
                const char* pm = getenv("EMU_PASS");
                if (pm && atoi(pm) > 0) {
                    const bool join2 = atoi(pm) == 2;
                    std::vector<int> items1;
                    for (int o = 0; o < 2; o++)
                        if (cnt[o] >= 3) items1.push_back(o);
                    unsigned long long dbg[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                    uint32_t istate[2] = {0, 0};
                    for (int gi = 0; gi < prm.n_gaps && !items1.empty(); gi++)
                        for (int pass = 0; pass < 3; pass++) {
                            next = 0;
                            if (pass == 2 && join2)
                                emu::run_block(0, 32, [&]() { phase1_join2_kernel<32>(&task, items1.data(), (int)items1.size(), &next, ws.data(), prm, pool_factor, gi, dbg, istate); });
                            else
                                emu::run_block(0, 32, [&]() { phase1_pass_kernel<32, 32>(&task, items1.data(), (int)items1.size(), &next, keys.data(), ws.data(), prm, pool_factor, sm_slot.data(), pass, gi, dbg, istate); });
                        }
                    const int item2 = 0;
                    for (int stage = 0; stage < 4; stage++) {
                        next = 0;
                        emu::run_block(0, 32, [&]() { phase2_stage_kernel<32, 32>(&task, &item2, 1, &next, keys.data(), ws.data(), prm, pool_factor, sm_slot.data(), ds, pf, hl ? 1 : 0, stage, istate); });
                    }
                } else
#endif
                emu::run_block(0, EMU_NT, [&]() { cluster_kernel<EMU_NT>(&task, &prof, &item, 1, &next, keys.data(), ws.data(), prm, pool_factor, sm_slot.data(), ds, pf, hl ? 1 : 0); });
                }
                if (task.status != 0) continue;   // anchor-point pool exhausted: retry with a larger pool, like the engine
                if ((fmt & 1) && hl) {   // homolog-count filter between clustering and writer, as in run_batch
                    const Prof& p = profs[pi];
                    HomologDev hd;
                    hd.lk_ptr = all_lk_ptr.data() + p.lk_base;
                    hd.lk_col = all_lk_col.data();
                    n_before += task.n_mult;
                    emu::run_block(0, HT, [&]() { homolog_filter_kernel(&task, 1, ws.data(), ds, n_fam, pf, &hd, pool_factor, prm.anchor_points); });
                }
                int32_t nm = task.n_mult;
                memcpy(got.data(), &nm, 4);
                if (nm > 0) {
                    got.resize(4 + (size_t)task.out_bytes);
                    const int sel = 0;
                    const int64_t off = 0;
                    if (hl) emu::run_block(0, PT, [&]() { pack_write_kernel<true>(&task, &sel, &off, 1, ws.data(), ds, pf, pool_factor, got.data() + 4); });
                    else emu::run_block(0, PT, [&]() { pack_write_kernel<false>(&task, &sel, &off, 1, ws.data(), ds, pf, pool_factor, got.data() + 4); });
                }
                break;
            }
        }
        if (expected.size() < 4) expected.assign(4, 0);
        int32_t nm;
        memcpy(&nm, got.data(), 4);
        n_mult += nm;
        n_bytes += (long long)got.size() - 4;
        if (got != expected) {
            bad++;
            int32_t ne;
            memcpy(&ne, expected.data(), 4);
            size_t k = 0;
            while (k < got.size() && k < expected.size() && got[k] == expected[k]) k++;
            fprintf(stderr, "GHM %d (%s %d x %d, D = %u + %u): %d multiplicons / %zu bytes, reference %d / %zu, first difference at byte %zu\n", gidx, hl ? "profile" : "level 2", xl,
                    yl, cnt[0], cnt[1], nm, got.size(), ne, expected.size(), k);
        }
    }
    fclose(f);
    printf("{\"ghms\": %d, \"multiplicons\": %lld, \"bytes\": %lld, \"mismatching_ghms\": %lld, \"collectives\": %llu, \"resident\": %lld, \"resident_fallbacks\": %lld, \"before_homolog_filter\": %lld}\n", n_ghm, n_mult, n_bytes, bad,
           emu::n_collectives, n_resident, n_fallback, n_before);
    return bad ? 1 : 0;
}
