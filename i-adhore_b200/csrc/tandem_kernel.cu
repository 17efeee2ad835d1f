// Tandem remapping on the device (family mode and, since round 2, table_type=pairs).
//
// Replaces GeneList::remapTandemsFamily (GeneList.cpp:143-200) with Gene::remapTo (Gene.cpp:51-62),
// driven per list by DataSet::remapTandems (DataSet.cpp:204-213): a gene is collapsed onto the
// nearest gene of its family (different ID, Gene.cpp:17-30) within tandem_gap + 1 positions, in three
// ordered passes: '+' onto earlier '+' (back to front), '-' onto later '-' (front to
// back), then any still unremapped gene onto any earlier gene; remapTo() also clears the target's
// flag.
//
// What is order-dependent and what is not:
//  * the partner a gene would be remapped to is state-independent (nearest family member in the
//    window) -> computed for every element in parallel (tandem_targets_kernel);
//  * passes 1 and 2 process a gene after every gene that could clear its flag (targets lie before /
//    after the gene, i.e. later in processing order), so their outcome is simply "has a same-strand
//    partner in the window";
//  * pass 3 is a true sequential recurrence (a cleared target is examined later and may remap
//    itself): one thread per list walks it back to front over the precomputed partners
//    (tandem_pass3_kernel) -- 46k steps for the longest legal list, lists in parallel.
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

__global__ void __launch_bounds__(256) tandem_targets_kernel(const int64_t* __restrict__ list_off, int n_lists,
                                                             const int32_t* __restrict__ elem_list,
                                                             const int32_t* __restrict__ fam,
                                                             const int32_t* __restrict__ gid,
                                                             const uint8_t* __restrict__ flags, int gap, int64_t total,
                                                             uint8_t* __restrict__ remapped, int32_t* __restrict__ t3)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int l = elem_list[i];
    const int64_t b = list_off[l], e = list_off[l + 1];
    const int f = fam[i], g = gid[i];
    const int s = flags[i] & 1;
    const int64_t lo = (i - (gap + 1) > b) ? i - (gap + 1) : b;
    const int64_t hi = (i + (gap + 1) < e - 1) ? i + (gap + 1) : e - 1;
    bool same = false;
    int32_t any = -1;
    // earlier genes, nearest first: pass 1 ('+' onto '+') and the partner of pass 3 (any strand)
    for (int64_t j = i - 1; j >= lo; j--) {
        if (fam[j] != f || gid[j] == g) continue;
        if (any < 0) any = (int32_t)j;
        if (s && (flags[j] & 1)) {
            same = true;
            break;
        }
        if (!s) break;  // a '-' gene only needs its pass-3 partner from this side
    }
    // later genes: pass 2 ('-' onto '-')
    if (!s)
        for (int64_t j = i + 1; j <= hi; j++)
            if (!(flags[j] & 1) && fam[j] == f && gid[j] != g) {
                same = true;
                break;
            }
    remapped[i] = same ? 1 : 0;
    t3[i] = any;
}

// table_type=pairs: GeneList::remapTandemsPairs (GeneList.cpp:84-141).  Same three passes; the predicate is
// Gene::isIndirectPairWith (Gene.cpp:33-49): j's gene is a partner of i's gene, or the two have a partner in common,
// and a gene without a pair set is never remapped (hasPairs, GeneList.cpp:88).  Pair sets = CSR over the dense gene ids
// (partners ascending), as iadh_create_pairs takes them.
__device__ __forceinline__ bool indirect_pair(const int64_t* __restrict__ pp, const int32_t* __restrict__ pi, int ga, int gb)
{
    int64_t a = pp[ga], ae = pp[ga + 1], b = pp[gb], be = pp[gb + 1];
    // direct: gb in pairs(ga)
    {
        int64_t lo = a, hi = ae;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (pi[mid] < gb) lo = mid + 1;
            else hi = mid;
        }
        if (lo < ae && pi[lo] == gb) return true;
    }
    // a common partner: merge of the two ascending lists
    while (a < ae && b < be) {
        const int x = pi[a], y = pi[b];
        if (x == y) return true;
        if (x < y) a++;
        else b++;
    }
    return false;
}
__global__ void __launch_bounds__(256) tandem_targets_pairs_kernel(const int64_t* __restrict__ list_off, int n_lists,
                                                                   const int32_t* __restrict__ elem_list,
                                                                   const int32_t* __restrict__ gid, const uint8_t* __restrict__ flags,
                                                                   const int64_t* __restrict__ pair_ptr,
                                                                   const int32_t* __restrict__ pair_ids, int n_genes, int gap,
                                                                   int64_t total, uint8_t* __restrict__ remapped, int32_t* __restrict__ t3)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    remapped[i] = 0;
    t3[i] = -1;
    const int g = gid[i];
    if (g < 0 || g >= n_genes || pair_ptr[g + 1] == pair_ptr[g]) return;   // !hasPairs()
    const int l = elem_list[i];
    const int64_t b = list_off[l], e = list_off[l + 1];
    const int s = flags[i] & 1;
    const int64_t lo = (i - (gap + 1) > b) ? i - (gap + 1) : b;
    const int64_t hi = (i + (gap + 1) < e - 1) ? i + (gap + 1) : e - 1;
    bool same = false;
    int32_t any = -1;
    for (int64_t j = i - 1; j >= lo; j--) {
        const int gj = gid[j];
        if (gj < 0 || gj >= n_genes || !indirect_pair(pair_ptr, pair_ids, g, gj)) continue;
        if (any < 0) any = (int32_t)j;
        if (s && (flags[j] & 1)) {
            same = true;
            break;
        }
        if (!s) break;
    }
    if (!s)
        for (int64_t j = i + 1; j <= hi; j++) {
            const int gj = gid[j];
            if (!(flags[j] & 1) && gj >= 0 && gj < n_genes && indirect_pair(pair_ptr, pair_ids, g, gj)) {
                same = true;
                break;
            }
        }
    remapped[i] = same ? 1 : 0;
    t3[i] = any;
}

__global__ void __launch_bounds__(128) tandem_pass3_kernel(const int64_t* __restrict__ list_off, int n_lists,
                                                           uint8_t* __restrict__ remapped, const int32_t* __restrict__ t3)
{
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n_lists) return;
    const int64_t b = list_off[l], e = list_off[l + 1];
    for (int64_t i = e - 1; i > b; i--) {
        if (remapped[i]) continue;
        const int32_t t = t3[i];
        if (t < 0) continue;
        remapped[i] = 1;
        remapped[t] = 0;
    }
}

}  // namespace

#ifndef IADH_HOST_EMU   // (tests/emu compiles the kernels above for the host)
void launch_tandem(const int64_t* list_off, int n_lists, const int32_t* elem_list, const int32_t* fam, const int32_t* gid,
                   const uint8_t* flags, int gap, int64_t total, uint8_t* remapped, int32_t* t3, cudaStream_t st)
{
    if (total <= 0) return;
    tandem_targets_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(list_off, n_lists, elem_list, fam, gid, flags, gap,
                                                                          total, remapped, t3);
    tandem_pass3_kernel<<<(n_lists + 127) / 128, 128, 0, st>>>(list_off, n_lists, remapped, t3);
}

void launch_tandem_pairs(const int64_t* list_off, int n_lists, const int32_t* elem_list, const int32_t* gid, const uint8_t* flags,
                         const int64_t* pair_ptr, const int32_t* pair_ids, int n_genes, int gap, int64_t total, uint8_t* remapped,
                         int32_t* t3, cudaStream_t st)
{
    if (total <= 0) return;
    tandem_targets_pairs_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(list_off, n_lists, elem_list, gid, flags, pair_ptr, pair_ids,
                                                                                n_genes, gap, total, remapped, t3);
    tandem_pass3_kernel<<<(n_lists + 127) / 128, 128, 0, st>>>(list_off, n_lists, remapped, t3);
}
#endif  // IADH_HOST_EMU
