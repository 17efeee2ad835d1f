// TEST INFRASTRUCTURE: K0 / K1 of build_kernels.cu compiled for the host and run under cuda_emu.h (whole grids, block
// after block) on the dataset of a reference trace:
//   1. the join tables of iadh_create, built by the table kernels exactly in the order of launch_tables (family /
//      gene-ID histograms, exclusive scans with 1,024-thread blocks, scatter through an atomic cursor, bucket sorts,
//      member records, strand prefix, (family, list) table), compared field by field with a plain host construction;
//   2. the level-2 dot sets of every list pair of the trace through the TILE kernels (count, per-GHM tile scan, fill)
//      and through the FAMILY-EXPANSION kernels (count, fill through atomic cursors, block sorts) on those tables,
//      compared with the dots the unmodified reference's GHM::buildMatrix produced (GHM.cpp:34-85).
// Family mode.  Grids are a few blocks (the kernels stride over any grid size).
//
// input (little endian): i32 n_lists; i64 list_off[n_lists+1]; i32 fam[T], gid[T]; u8 flags[T] (bit 0 strand);
//   i32 n_ghm; per GHM: i32 x_list, y_list; u32 cnt0, cnt1; u32 keys[cnt0 + cnt1]  ((x << 16) | y, orientation 0 block first)
//   optional: i32 n_profiles; per profile: i32 n_masks, n_masks x (i32 list, begin, end) applied before it; i32 S, L;
//     i32 fam[S*L], gid[S*L]; u8 flags[S*L] (bit 0 strand, bit 1 gap); i32 n_ghm; per GHM: i32 y_list; u32 cnt0, cnt1; keys
//   3. (that section) the profile x list dot sets through build_kernel<HL> (tiles) and hl_expand_kernel (+ block sorts)
#define IADH_HOST_EMU 1
#define EMU_MAX_THREADS 1024
#define EMU_STACK_BYTES (256 << 10)
#include "cuda_emu.h"

#include <algorithm>

#include "../../i-adhore_b200/csrc/build_kernels.cu"

#include "emu_runtime.inc"

template <typename T>
static bool rd(FILE* f, T* p, size_t n) { return n == 0 || fread(p, sizeof(T), n, f) == n; }

// launch_scan_i32 (build_kernels.cu): block-local scan, scan of the block sums, add
static void scan_i32(const int32_t* in, int64_t n, int32_t* out, int32_t* block_sum)
{
    const int n_blocks = (int)((n + 1023) / 1024);
    if (n_blocks > 0) emu::launch(n_blocks, 1024, [&]() { scan_block_kernel(in, n, out, block_sum); });
    emu::launch(1, 1024, [&]() { scan_sums_kernel(block_sum, n_blocks); });
    emu::launch((unsigned)((n + 1 + 255) / 256), 256, [&]() { scan_add_kernel(out, n, block_sum, n_blocks); });
}

template <typename T>
static long long diff(const char* what, const std::vector<T>& a, const std::vector<T>& b, size_t n)
{
    long long bad = 0;
    for (size_t i = 0; i < n; i++) bad += a[i] != b[i];
    if (bad) fprintf(stderr, "table %s: %lld of %zu entries differ\n", what, bad, n);
    return bad;
}

int main(int argc, char** argv)
{
    if (argc < 2) { fprintf(stderr, "usage: %s <k1.bin>\n", argv[0]); return 2; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    int n_lists = 0;
    if (!rd(f, &n_lists, 1)) return 2;
    std::vector<int64_t> list_off(n_lists + 1);
    if (!rd(f, list_off.data(), list_off.size())) return 2;
    const int64_t T = list_off[n_lists];
    std::vector<int32_t> fam(T), gid(T);
    std::vector<uint8_t> flags(T);
    if (!rd(f, fam.data(), T) || !rd(f, gid.data(), T) || !rd(f, flags.data(), T)) return 2;
    int n_fam = 0, n_gid = 0;
    for (int64_t i = 0; i < T; i++) {
        n_fam = std::max(n_fam, fam[i] + 1);
        n_gid = std::max(n_gid, gid[i] + 1);
    }
    const int GRID = 3;   // a few blocks: every kernel strides over its grid

    // ---- 1. tables by the kernels, in the order of launch_tables -------------------------------------------------
    std::vector<int32_t> fam_cnt(n_fam + 1, 0), gid_cnt(n_gid + 1, 0), cursor(n_fam + 1, 0), fam_ptr(n_fam + 2, -1), members(T, -1),
        elem_list(T, -1), mem_list(T, -1), self_slot(T, -1), plus_prefix(T + 2, -1), tmp(T + 1, -1), block_sum(T / 1024 + n_fam / 1024 + 4, 0);
    std::vector<uint32_t> mem_y(T, 0xffffffffu);
    std::vector<uint8_t> dflags(flags);
    const unsigned g256 = (unsigned)((T + 255) / 256);
    emu::launch(g256, 256, [&]() { tab_count_kernel(fam.data(), gid.data(), T, fam_cnt.data(), gid_cnt.data()); });
    scan_i32(fam_cnt.data(), n_fam, fam_ptr.data(), block_sum.data());
    emu::launch(g256, 256, [&]() {
        tab_scatter_kernel(fam.data(), list_off.data(), n_lists, gid.data(), gid_cnt.data(), T, fam_ptr.data(), cursor.data(), members.data(),
                           elem_list.data(), dflags.data());
    });
    emu::launch(GRID, 256, [&]() { csr_sort_warp_kernel<8>(n_fam, fam_ptr.data(), (uint32_t*)members.data()); });
    emu::launch(GRID, 256, [&]() { csr_sort_warp_kernel<32>(n_fam, fam_ptr.data(), (uint32_t*)members.data()); });
    emu::launch(GRID, 256, [&]() { csr_sort_cta_kernel(n_fam, fam_ptr.data(), (uint32_t*)members.data()); });
    emu::launch(g256, 256, [&]() {
        tab_records_kernel(members.data(), T, elem_list.data(), list_off.data(), dflags.data(), mem_y.data(), mem_list.data(), self_slot.data(),
                           tmp.data());
    });
    scan_i32(tmp.data(), T, plus_prefix.data(), block_sum.data());
    std::vector<int32_t> fam_list((size_t)n_fam * (n_lists + 1), -1);
    emu::launch((n_fam + 127) / 128, 128, [&]() { fam_list_kernel(fam_ptr.data(), members.data(), list_off.data(), n_fam, n_lists, T, fam_list.data()); });

    // the same tables by plain loops
    long long table_bad = 0;
    {
        std::vector<int32_t> h_ptr(n_fam + 2, 0), h_mem(T), h_el(T), h_ml(T), h_self(T), h_pp(T + 2, 0), h_fl((size_t)n_fam * (n_lists + 1));
        std::vector<uint32_t> h_my(T);
        std::vector<uint8_t> h_flags(flags);
        for (int64_t i = 0; i < T; i++) h_ptr[fam[i] + 1]++;
        for (int k = 0; k < n_fam; k++) h_ptr[k + 1] += h_ptr[k];
        std::vector<int32_t> cur(h_ptr.begin(), h_ptr.begin() + n_fam);
        for (int l = 0; l < n_lists; l++)
            for (int64_t i = list_off[l]; i < list_off[l + 1]; i++) h_el[i] = l;
        for (int64_t i = 0; i < T; i++) h_mem[cur[fam[i]]++] = (int32_t)i;   // ascending: elements are visited in order
        std::vector<int> seen(n_gid + 1, 0);
        for (int64_t i = 0; i < T; i++)
            if (gid[i] >= 0) seen[gid[i]]++;
        for (int64_t i = 0; i < T; i++)
            if (gid[i] < 0 || seen[gid[i]] > 1) h_flags[i] |= 8;
        for (int64_t m = 0; m < T; m++) {
            const int e = h_mem[m], l = h_el[e];
            h_my[m] = ((uint32_t)(e - list_off[l]) << 1) | (uint32_t)(flags[e] & 1);
            h_ml[m] = l;
            h_self[e] = (int32_t)m;
            h_pp[m + 1] = h_pp[m] + (flags[e] & 1);
        }
        for (int fm = 0; fm < n_fam; fm++) {
            int m = h_ptr[fm];
            for (int l = 0; l <= n_lists; l++) {
                const int64_t lim = l < n_lists ? list_off[l] : T;
                while (m < h_ptr[fm + 1] && h_mem[m] < lim) m++;
                h_fl[(size_t)fm * (n_lists + 1) + l] = m;
            }
        }
        table_bad += diff("fam_ptr", fam_ptr, h_ptr, n_fam + 1);
        table_bad += diff("fam_members", members, h_mem, T);
        table_bad += diff("elem_list", elem_list, h_el, T);
        table_bad += diff("mem_y", mem_y, h_my, T);
        table_bad += diff("mem_list", mem_list, h_ml, T);
        table_bad += diff("self_slot", self_slot, h_self, T);
        table_bad += diff("plus_prefix", plus_prefix, h_pp, T + 1);
        table_bad += diff("flags", dflags, h_flags, T);
        table_bad += diff("fam_list", fam_list, h_fl, h_fl.size());
    }

    DatasetDev ds;
    memset(&ds, 0, sizeof ds);
    ds.n_lists = n_lists;
    ds.list_off = list_off.data();
    ds.fam = fam.data();
    ds.gid = gid.data();
    ds.num_id = gid.data();   // (not read by K1)
    ds.flags = dflags.data();
    ds.fam_ptr = fam_ptr.data();
    ds.fam_members = members.data();
    ds.mem_y = mem_y.data();
    ds.plus_prefix = plus_prefix.data();
    ds.self_slot = self_slot.data();
    ds.mem_list = mem_list.data();
    ds.elem_list = elem_list.data();
    ds.fam_list = fam_list.data();

    // ---- 2. the level-2 GHMs of the trace ---------------------------------------------------------------------------
    int n = 0;
    if (!rd(f, &n, 1)) return 2;
    std::vector<GhmTask> tasks(n);
    std::vector<std::vector<uint32_t>> want(n);
    std::vector<uint32_t> want_cnt(2 * (size_t)n);
    for (int i = 0; i < n; i++) {
        memset(&tasks[i], 0, sizeof(GhmTask));
        uint32_t c[2];
        if (!rd(f, &tasks[i].x_list, 1) || !rd(f, &tasks[i].y_list, 1) || !rd(f, c, 2)) return 2;
        want[i].resize(c[0] + c[1]);
        if (!rd(f, want[i].data(), want[i].size())) return 2;
        want_cnt[2 * i] = c[0];
        want_cnt[2 * i + 1] = c[1];
        tasks[i].seq = i;
    }
    long long dots = 0;
    auto compare = [&](const char* path, const std::vector<GhmTask>& tk, const std::vector<uint32_t>& keys) {
        long long bad = 0;
        for (int i = 0; i < n; i++) {
            bool ok = tk[i].cnt[0] == want_cnt[2 * i] && tk[i].cnt[1] == want_cnt[2 * i + 1];
            if (ok) ok = std::equal(want[i].begin(), want[i].end(), keys.begin() + tk[i].dot_off);
            if (!ok) {
                bad++;
                fprintf(stderr, "%s: GHM %d (%d x %d): %u + %u dots, reference %u + %u%s\n", path, i, tk[i].x_list, tk[i].y_list, tk[i].cnt[0],
                        tk[i].cnt[1], want_cnt[2 * i], want_cnt[2 * i + 1], "");
            }
        }
        return bad;
    };

    // tile path: launch_build_count, launch_tile_scan, (host: dot offsets), launch_build_fill
    long long tile_bad = 0;
    {
        std::vector<GhmTask> tk(tasks);
        std::vector<int64_t> tile_prefix(n + 1, 0);
        for (int i = 0; i < n; i++) {
            const int64_t nx = list_off[tk[i].x_list + 1] - list_off[tk[i].x_list];
            tile_prefix[i + 1] = tile_prefix[i] + (nx + 31) / 32;
        }
        const int64_t n_tiles = tile_prefix[n];
        std::vector<uint32_t> tile_cnt(2 * (size_t)n_tiles + 2, 0), tile_off(2 * (size_t)n_tiles + 2, 0);
        emu::launch(GRID, 256, [&]() { l2_build_kernel<false>(ds, tk.data(), n, tile_prefix.data(), n_tiles, tile_cnt.data(), nullptr, nullptr); });
        emu::launch((n + 7) / 8, 256, [&]() { tile_scan_kernel(tile_prefix.data(), n, tile_cnt.data(), tile_off.data()); });
        int64_t total = 0;
        for (int i = 0; i < n; i++) {
            tk[i].dot_off = total;
            total += (int64_t)tk[i].cnt[0] + tk[i].cnt[1];
        }
        std::vector<uint32_t> keys((size_t)total + 1, 0xdeadbeefu);
        emu::launch(GRID, 256, [&]() { l2_build_kernel<true>(ds, tk.data(), n, tile_prefix.data(), n_tiles, nullptr, tile_off.data(), keys.data()); });
        tile_bad = compare("tiles", tk, keys);
        dots = total;
    }

    // family-expansion path: launch_expand_count, (host: offsets), launch_expand_fill (scatter counts, fill, block sorts)
    long long expand_bad = 0;
    int expand_skipped = 0;
    {
        std::vector<GhmTask> tk(tasks);
        std::vector<int32_t> pair_idx((size_t)n_lists * n_lists, -1);
        std::vector<uint8_t> row_used(n_lists, 0);
        for (int i = 0; i < n; i++) {
            pair_idx[(size_t)tk[i].x_list * n_lists + tk[i].y_list] = i;
            row_used[tk[i].x_list] = 1;
        }
        std::vector<uint32_t> cnt(2 * (size_t)n + 2, 0), cursor2(2 * (size_t)n + 2, 0);
        emu::launch(GRID, 256, [&]() { l2_expand_kernel<false>(ds, T, pair_idx.data(), row_used.data(), cnt.data(), nullptr, nullptr, nullptr); });
        std::vector<int64_t> dot_off(n + 1, 0);
        for (int i = 0; i < n; i++) dot_off[i + 1] = dot_off[i] + cnt[2 * i] + cnt[2 * i + 1];
        // a block beyond the sort kernel's capacity sends the call to the tile kernels (build_stage, engine.cu)
        bool fits = true;
        for (int i = 0; i < 2 * n; i++) fits = fits && cnt[i] <= (uint32_t)SORT_MAX;
        if (!fits) {
            for (int i = 0; i < 2 * n; i++)
                if (cnt[i] != want_cnt[i]) expand_bad++;
            expand_skipped = 1;
        } else {
        for (int i = 0; i < n; i++) tk[i].dot_off = dot_off[i];
        std::vector<uint32_t> keys((size_t)dot_off[n] + 1, 0xdeadbeefu);
        emu::launch((n + 255) / 256, 256, [&]() { scatter_cnt_kernel(tk.data(), n, cnt.data()); });
        emu::launch(GRID, 256, [&]() { l2_expand_kernel<true>(ds, T, pair_idx.data(), row_used.data(), cursor2.data(), dot_off.data(), cnt.data(), keys.data()); });
        const int n_seg = 2 * n;
        emu::launch(GRID, 256, [&]() { seg_sort_warp_kernel<8>(n_seg, dot_off.data(), cnt.data(), keys.data()); });
        emu::launch(GRID, 256, [&]() { seg_sort_warp_kernel<32>(n_seg, dot_off.data(), cnt.data(), keys.data()); });
        emu::launch(GRID, 256, [&]() { seg_sort_cta_kernel(n_seg, dot_off.data(), cnt.data(), keys.data()); });
        expand_bad = compare("expansion", tk, keys);
        }
    }
    // ---- 3. the profile GHMs of the trace (optional section): GHMProfile::buildMatrix (GHMProfile.cpp:17-59) -------
    // per profile: the masks applied before its search (GeneList::mask, higherLevel.cpp:165-171), the aligned segments,
    // the lists searched with the reference's dots.  Column groups as profile_groups() of engine.cu builds them.
    long long hl_ghms = 0, hl_dots = 0, hl_tile_bad = 0, hl_expand_bad = 0;
    int n_prof = 0;
    if (rd(f, &n_prof, 1)) {
        for (int pi = 0; pi < n_prof; pi++) {
            int nm = 0, S = 0, L = 0, ng = 0;
            if (!rd(f, &nm, 1)) return 2;
            for (int k = 0; k < nm; k++) {
                int m[3];
                if (!rd(f, m, 3)) return 2;
                for (int64_t i = list_off[m[0]] + m[1]; i <= list_off[m[0]] + m[2]; i++) dflags[i] |= 4;
            }
            if (!rd(f, &S, 1) || !rd(f, &L, 1)) return 2;
            const size_t SL = (size_t)S * L;
            std::vector<int32_t> pfam(SL), pgid(SL);
            std::vector<uint8_t> pflags(SL);
            if (!rd(f, pfam.data(), SL) || !rd(f, pgid.data(), SL) || !rd(f, pflags.data(), SL) || !rd(f, &ng, 1)) return 2;
            std::vector<int32_t> col_ptr(L + 1, 0), gfam, ggid;
            std::vector<uint8_t> gstrand;
            std::vector<HlGroup> items;
            for (int x = 0; x < L; x++) {
                const size_t g0 = gfam.size();
                for (int sg = 0; sg < S; sg++) {
                    const size_t k = (size_t)sg * L + x;
                    if (pflags[k] & 2) continue;
                    const int fm = pfam[k];
                    const uint8_t st = pflags[k] & 1;
                    if (fm < 0 || fm >= n_fam) continue;
                    size_t g = g0;
                    for (; g < gfam.size(); g++)
                        if (gfam[g] == fm && gstrand[g] == st) break;
                    if (g == gfam.size()) {
                        gfam.push_back(fm);
                        ggid.push_back(pgid[k]);
                        gstrand.push_back(st);
                    } else if (ggid[g] != pgid[k])
                        ggid[g] = -1;
                }
                col_ptr[x + 1] = (int32_t)gfam.size();
                for (size_t g = g0; g < gfam.size(); g++) {
                    HlGroup it;
                    it.fam = gfam[g];
                    it.gid = ggid[g];
                    it.x_strand = (uint32_t)x | (gstrand[g] ? 0x80000000u : 0u);
                    it.prof = 0;
                    items.push_back(it);
                }
            }
            ProfileDev pf;
            memset(&pf, 0, sizeof pf);
            pf.S = S; pf.L = L;
            pf.col_ptr = col_ptr.data(); pf.grp_fam = gfam.data(); pf.grp_gid = ggid.data(); pf.grp_strand = gstrand.data();
            pf.fam = pfam.data(); pf.gid = pgid.data(); pf.num_id = pgid.data(); pf.flags = pflags.data();
            std::vector<GhmTask> ht(ng);
            std::vector<std::vector<uint32_t>> hw(ng);
            std::vector<uint32_t> hwc(2 * (size_t)ng);
            std::vector<int32_t> task_idx(n_lists, -1);
            for (int i = 0; i < ng; i++) {
                memset(&ht[i], 0, sizeof(GhmTask));
                uint32_t c[2];
                ht[i].x_list = -1;
                if (!rd(f, &ht[i].y_list, 1) || !rd(f, c, 2)) return 2;
                hw[i].resize(c[0] + c[1]);
                if (!rd(f, hw[i].data(), hw[i].size())) return 2;
                hwc[2 * i] = c[0]; hwc[2 * i + 1] = c[1];
                task_idx[ht[i].y_list] = i;
                hl_dots += c[0] + c[1];
            }
            hl_ghms += ng;
            auto cmp = [&](const std::vector<GhmTask>& tk, const std::vector<uint32_t>& keys) {
                long long bad = 0;
                for (int i = 0; i < ng; i++) {
                    bool ok = tk[i].cnt[0] == hwc[2 * i] && tk[i].cnt[1] == hwc[2 * i + 1];
                    if (ok) ok = std::equal(hw[i].begin(), hw[i].end(), keys.begin() + tk[i].dot_off);
                    bad += !ok;
                }
                return bad;
            };
            {   // tile path (build_kernel<HL>)
                std::vector<GhmTask> tk(ht);
                std::vector<int64_t> tile_prefix(ng + 1, 0);
                for (int i = 0; i < ng; i++) tile_prefix[i + 1] = tile_prefix[i] + (L + 31) / 32;
                const int64_t n_tiles = tile_prefix[ng];
                std::vector<uint32_t> tile_cnt(2 * (size_t)n_tiles + 2, 0), tile_off(2 * (size_t)n_tiles + 2, 0);
                if (n_tiles) emu::launch(GRID, 256, [&]() { build_kernel<true, false>(ds, pf, tk.data(), ng, tile_prefix.data(), n_tiles, tile_cnt.data(), nullptr, nullptr); });
                if (ng) emu::launch((ng + 7) / 8, 256, [&]() { tile_scan_kernel(tile_prefix.data(), ng, tile_cnt.data(), tile_off.data()); });
                int64_t total = 0;
                for (int i = 0; i < ng; i++) {
                    tk[i].dot_off = total;
                    total += (int64_t)tk[i].cnt[0] + tk[i].cnt[1];
                }
                std::vector<uint32_t> keys((size_t)total + 1, 0xdeadbeefu);
                if (n_tiles) emu::launch(GRID, 256, [&]() { build_kernel<true, true>(ds, pf, tk.data(), ng, tile_prefix.data(), n_tiles, nullptr, tile_off.data(), keys.data()); });
                hl_tile_bad += cmp(tk, keys);
            }
            {   // family-expansion path (hl_expand_kernel + block sorts)
                std::vector<GhmTask> tk(ht);
                const int n_items = (int)items.size();
                std::vector<uint32_t> cnt(2 * (size_t)ng + 2, 0), cur2(2 * (size_t)ng + 2, 0);
                if (n_items) emu::launch(GRID, 256, [&]() { hl_expand_kernel<false>(ds, items.data(), n_items, task_idx.data(), nullptr, cnt.data(), nullptr, nullptr, nullptr); });
                std::vector<int64_t> dot_off(ng + 1, 0);
                for (int i = 0; i < ng; i++) dot_off[i + 1] = dot_off[i] + cnt[2 * i] + cnt[2 * i + 1];
                for (int i = 0; i < ng; i++) tk[i].dot_off = dot_off[i];
                std::vector<uint32_t> keys((size_t)dot_off[ng] + 1, 0xdeadbeefu);
                if (ng) emu::launch((ng + 255) / 256, 256, [&]() { scatter_cnt_kernel(tk.data(), ng, cnt.data()); });
                if (n_items) emu::launch(GRID, 256, [&]() { hl_expand_kernel<true>(ds, items.data(), n_items, task_idx.data(), nullptr, cur2.data(), dot_off.data(), cnt.data(), keys.data()); });
                const int n_seg = 2 * ng;
                if (ng) {
                    emu::launch(GRID, 256, [&]() { seg_sort_warp_kernel<8>(n_seg, dot_off.data(), cnt.data(), keys.data()); });
                    emu::launch(GRID, 256, [&]() { seg_sort_warp_kernel<32>(n_seg, dot_off.data(), cnt.data(), keys.data()); });
                    emu::launch(GRID, 256, [&]() { seg_sort_cta_kernel(n_seg, dot_off.data(), cnt.data(), keys.data()); });
                }
                hl_expand_bad += cmp(tk, keys);
            }
        }
    }
    fclose(f);
    printf("{\"elements\": %lld, \"families\": %d, \"ghms\": %d, \"dots\": %lld, \"table_mismatches\": %lld, \"tile_mismatching_ghms\": %lld, "
           "\"expansion_mismatching_ghms\": %lld, \"expansion_fill_skipped\": %d, \"profiles\": %d, \"profile_ghms\": %lld, \"profile_dots\": %lld, "
           "\"profile_tile_mismatching_ghms\": %lld, \"profile_expansion_mismatching_ghms\": %lld}\n",
           (long long)T, n_fam, n, dots, table_bad, tile_bad, expand_bad, expand_skipped, n_prof, hl_ghms, hl_dots, hl_tile_bad, hl_expand_bad);
    return (table_bad || tile_bad || expand_bad || hl_tile_bad || hl_expand_bad) ? 1 : 0;
}
