// TEST INFRASTRUCTURE: the tandem-remapping kernels (tandem_kernel.cu: tandem_targets_kernel / tandem_targets_pairs_kernel,
// tandem_pass3_kernel) compiled for the host and run block by block under cuda_emu.h on the gene lists of a reference
// trace BEFORE remapping; the flags must equal Gene::isRemapped() as the unmodified reference left them
// (DataSet::remapTandems, DataSet.cpp:204-213).
//
// input (little endian): i32 tandem_gap, pairs_mode, n_lists; i64 list_off[n_lists+1]; i32 fam[T], gid[T]; u8 strand[T];
//   u8 expected[T]; pairs mode: i32 n_genes; i64 pair_ptr[n_genes+1]; i32 pair_ids[]
#define IADH_HOST_EMU 1
#include "cuda_emu.h"

#include "../../i-adhore_b200/csrc/tandem_kernel.cu"

#include "emu_runtime.inc"

template <typename T>
static bool rd(FILE* f, T* p, size_t n) { return n == 0 || fread(p, sizeof(T), n, f) == n; }

int main(int argc, char** argv)
{
    if (argc < 2) { fprintf(stderr, "usage: %s <lists.bin>\n", argv[0]); return 2; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    int gap = 0, pairs = 0, n_lists = 0;
    if (!rd(f, &gap, 1) || !rd(f, &pairs, 1) || !rd(f, &n_lists, 1)) return 2;
    std::vector<int64_t> list_off(n_lists + 1);
    if (!rd(f, list_off.data(), list_off.size())) return 2;
    const int64_t T = list_off[n_lists];
    std::vector<int32_t> fam(T), gid(T), elem_list(T), t3(T, -9);
    std::vector<uint8_t> flags(T), expected(T), remapped(T, 7);
    if (!rd(f, fam.data(), T) || !rd(f, gid.data(), T) || !rd(f, flags.data(), T) || !rd(f, expected.data(), T)) return 2;
    for (int l = 0; l < n_lists; l++)
        for (int64_t i = list_off[l]; i < list_off[l + 1]; i++) elem_list[i] = l;
    int n_genes = 0;
    std::vector<int64_t> pp;
    std::vector<int32_t> pi;
    if (pairs) {
        if (!rd(f, &n_genes, 1)) return 2;
        pp.resize(n_genes + 1);
        if (!rd(f, pp.data(), pp.size())) return 2;
        pi.resize((size_t)pp[n_genes]);
        if (!rd(f, pi.data(), pi.size())) return 2;
    }
    fclose(f);
    // as launch_tandem / launch_tandem_pairs: 256-thread blocks over the elements, then 128-thread blocks over the lists
    for (int64_t b = 0; b * 256 < T; b++)
        emu::run_block((unsigned)b, 256, [&]() {
            if (pairs)
                tandem_targets_pairs_kernel(list_off.data(), n_lists, elem_list.data(), gid.data(), flags.data(), pp.data(), pi.data(), n_genes, gap, T,
                                            remapped.data(), t3.data());
            else
                tandem_targets_kernel(list_off.data(), n_lists, elem_list.data(), fam.data(), gid.data(), flags.data(), gap, T, remapped.data(), t3.data());
        });
    for (int b = 0; b * 128 < n_lists; b++)
        emu::run_block((unsigned)b, 128, [&]() { tandem_pass3_kernel(list_off.data(), n_lists, remapped.data(), t3.data()); });
    long long bad = 0, n_remapped = 0;
    for (int64_t i = 0; i < T; i++) {
        bad += remapped[i] != expected[i];
        n_remapped += expected[i];
    }
    printf("{\"elements\": %lld, \"lists\": %d, \"remapped\": %lld, \"mismatching\": %lld}\n", (long long)T, n_lists, n_remapped, bad);
    return bad ? 1 : 0;
}
