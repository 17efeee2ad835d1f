// Drop-in shim: the ONLY translation unit a maintainer adds to i-ADHoRe 3.0 to run its hot path
// on a B200.  It supplies strong definitions of the two driver members
//
//     void DataSet::parallelLevel2ADHoReDyn()                          (levelTwo.cpp:177-292)
//     void DataSet::parallelProfileSearch(const Profile*, vector<..>&)  (higherLevel.cpp:252-326)
//
// that replace the pthread/MPI work distribution by calls into the C-ABI of
// include/iadhore_b200.h.  Everything else -- Settings, ingest, tandem remapping, the profile
// loop with its aligners and masks, every output table -- is the unmodified reference, compiled
// from its own sources.  Results cross the boundary in the reference's own wire format and are
// rehydrated with its own unpack constructors (Multiplicon.cpp:266-297), exactly as the reference
// does for multiplicons received from another MPI rank.
//
// This file includes the reference's headers; it contains none of the reference's code.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <set>
#include <ctime>
#include <map>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "DataSet.h"
#include "GeneList.h"
#include "GenePairs.h"
#include "Gene.h"
#include "ListElement.h"
#include "Multiplicon.h"
#include "Profile.h"
#include "Settings.h"
#include "parallel.h"

#include "iadhore_b200.h"

namespace {

struct Bridge {
    iadh_ctx* ctx = nullptr;
    const void* owner = nullptr;   // the gene lists this context was made from (a second DataSet gets its own)
    size_t owner_lists = 0;
    std::map<std::string, int> fam_ids, gene_ids;
    std::vector<std::vector<unsigned char> > masked;  // shadow of ListElement::isMasked per list
    std::vector<int32_t> fam_of, gid_of;               // by ListElement numID (profile elements are copies)
    int rank = 0, n_ranks = 1;
    double t_l2 = 0, t_hl = 0;
    long n_hl = 0;
    // wall-clock split of the profile searches (IADH_VERBOSE=1 prints it at exit)
    double w_mask = 0, w_desc = 0, w_call = 0, w_unpack = 0;
    long long n_ghm = 0, n_found = 0, n_returned = 0, n_prefiltered = 0;
    std::vector<std::vector<int32_t> > list_fam;       // family id per remapped element, per list
};
Bridge B;

double wall()
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

void report()
{
    if (B.ctx) {   // (atexit: streams, events and device blocks go back before the CUDA runtime shuts down)
        iadh_destroy(B.ctx);
        B.ctx = nullptr;
    }
    if (!getenv("IADH_VERBOSE")) return;
    fprintf(stderr,
            "[iadh] profile searches: %ld calls, %lld GHMProfiles, %lld multiplicons returned, %lld dropped before unpacking, %lld kept; "
            "wall: masks %.3fs, descriptor %.3fs, device call %.3fs (device time %.3fs), unpack+filter %.3fs\n",
            B.n_hl, B.n_ghm, B.n_returned, B.n_prefiltered, B.n_found, B.w_mask, B.w_desc, B.w_call, B.t_hl * 1e-3, B.w_unpack);
}

// The CUDA context is created on a helper thread while the reference's main() reads the gene lists
// and the family table (same order of magnitude); ensure_context() joins it.
struct Warmup {
    std::thread th;
    Warmup()
    {
        const char* dev = getenv("IADH_DEVICE");
        const int d = dev ? atoi(dev) : 0;
        th = std::thread([d]() { iadh_warmup(d); });
    }
    void join()
    {
        if (th.joinable()) th.join();
    }
    ~Warmup() { join(); }
};
Warmup g_warmup;

int intern(std::map<std::string, int>& m, const std::string& s)
{
    std::map<std::string, int>::iterator it = m.find(s);
    if (it != m.end()) return it->second;
    const int id = (int)m.size();
    m[s] = id;
    return id;
}

void check(int rc, const char* what)
{
    if (rc != IADH_OK) throw std::runtime_error(std::string(what) + ": " + iadh_last_error());
}

// Flatten the DataSet (after mapGenes + remapTandems + sortGeneLists) and create the device context.
void ensure_context(const std::vector<GeneList*>& genelists, const Settings& settings, GenePairs* genepairs)
{
    // one context per DataSet: a different set of gene lists (another DataSet in the same process) replaces it
    const void* owner = genelists.empty() ? (const void*)&genelists : (const void*)genelists[0];
    if (B.ctx && B.owner == owner && B.owner_lists == genelists.size()) return;
    if (B.ctx) {
        iadh_destroy(B.ctx);
        B = Bridge();
    }
    B.owner = owner;
    B.owner_lists = genelists.size();
    std::vector<int64_t> off(1, 0);
    std::vector<int32_t> fam, gid, num;
    std::vector<uint8_t> flags;
    const char* dev = getenv("IADH_DEVICE");   // (MPI build: one rank per GPU, e.g. IADH_DEVICE=$OMPI_COMM_WORLD_LOCAL_RANK)
    g_warmup.join();
    // Strings -> dense ids on the device (iadh_intern_strings), once: the family and the gene ID of every element of
    // every list, in list order.  Everything below -- and K1's join -- works on these integers; the reference would
    // hash the same strings again for every list pair (GHM.cpp:34-50).
    {
        std::string fblob, gblob;
        std::vector<int64_t> foff(1, 0), goff(1, 0);
        std::vector<int> ids;
        for (size_t l = 0; l < genelists.size(); l++) {
            const std::vector<ListElement*>& all = genelists[l]->getElements();
            for (size_t i = 0; i < all.size(); i++) {
                fblob += all[i]->getGene().getFamily();
                gblob += all[i]->getGene().getID();
                foff.push_back((int64_t)fblob.size());
                goff.push_back((int64_t)gblob.size());
                ids.push_back(all[i]->getNumID());
            }
        }
        const int64_t n = (int64_t)ids.size();
        std::vector<int32_t> fid((size_t)n), gidv((size_t)n);
        int32_t nf = 0, ng = 0;
        check(iadh_intern_strings(fblob.data(), foff.data(), n, dev ? atoi(dev) : 0, fid.data(), &nf), "iadh_intern_strings (families)");
        check(iadh_intern_strings(gblob.data(), goff.data(), n, dev ? atoi(dev) : 0, gidv.data(), &ng), "iadh_intern_strings (gene IDs)");
        for (int64_t i = 0; i < n; i++) {
            const int id = ids[i];
            if (id >= (int)B.fam_of.size()) {
                B.fam_of.resize(id + 1, -1);
                B.gid_of.resize(id + 1, -1);
            }
            B.fam_of[id] = fid[i];
            B.gid_of[id] = gidv[i];
            // ids count the distinct strings in order of first occurrence: the string -> id maps of the host side
            // (pairs tables, elements met later without a numID) are filled as new ids appear
            if (fid[i] == (int32_t)B.fam_ids.size()) B.fam_ids[fblob.substr((size_t)foff[i], (size_t)(foff[i + 1] - foff[i]))] = fid[i];
            if (gidv[i] == (int32_t)B.gene_ids.size()) B.gene_ids[gblob.substr((size_t)goff[i], (size_t)(goff[i + 1] - goff[i]))] = gidv[i];
        }
    }
    B.masked.resize(genelists.size());
    for (size_t l = 0; l < genelists.size(); l++) {
        const std::vector<ListElement*>& re = genelists[l]->getRemappedElements();
        for (size_t i = 0; i < re.size(); i++) {
            const int id = re[i]->getNumID();
            fam.push_back(B.fam_of[id]);
            gid.push_back(B.gid_of[id]);
            num.push_back(id);
            flags.push_back((re[i]->getOrientation() ? IADH_F_STRAND : 0) | (re[i]->isMasked() ? IADH_F_MASKED : 0));
        }
        B.list_fam.resize(genelists.size());
        B.list_fam[l].assign(fam.end() - re.size(), fam.end());
        B.masked[l].assign(re.size(), 0);
        for (size_t i = 0; i < re.size(); i++) B.masked[l][i] = re[i]->isMasked() ? 1 : 0;
        off.push_back((int64_t)fam.size());
    }
    iadh_dataset_desc d;
    d.n_lists = (int32_t)genelists.size();
    d.list_off = off.data();
    d.fam = fam.data();
    d.gid = gid.data();
    d.num_id = num.data();
    d.flags = flags.data();
    if (settings.useFamily())
        check(iadh_create(&d, dev ? atoi(dev) : 0, &B.ctx), "iadh_create");
    else {
        // table_type=pairs: the pair sets DataSet::getGenePairs left in `genepairs` (the very sets
        // Gene::pairs points to, DataSet.cpp:165-183), as a CSR over the interned gene ids
        const int n_genes = (int)B.gene_ids.size();
        std::vector<std::vector<int32_t> > ps(n_genes);
        for (std::map<std::string, int>::const_iterator it = B.gene_ids.begin(); it != B.gene_ids.end(); it++) {
            hash_set<std::string, stringhash>* set = genepairs ? genepairs->getPairsOf(it->first) : NULL;
            if (!set) continue;
            std::vector<int32_t>& v = ps[it->second];
            for (hash_set<std::string, stringhash>::const_iterator p = set->begin(); p != set->end(); p++) {
                std::map<std::string, int>::const_iterator q = B.gene_ids.find(*p);
                if (q != B.gene_ids.end()) v.push_back(q->second);
            }
            std::sort(v.begin(), v.end());
        }
        std::vector<int64_t> pptr(n_genes + 1, 0);
        std::vector<int32_t> pids;
        for (int g = 0; g < n_genes; g++) {
            pids.insert(pids.end(), ps[g].begin(), ps[g].end());
            pptr[g + 1] = (int64_t)pids.size();
        }
        check(iadh_create_pairs(&d, pptr.data(), pids.data(), n_genes, dev ? atoi(dev) : 0, &B.ctx), "iadh_create_pairs");
    }
    static bool registered = false;
    if (!registered) {
        atexit(report);
        registered = true;
    }

    iadh_params p;
    int gaps[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    settings.getGapSizes(gaps);
    p.n_gaps = 0;
    for (int i = 0; i < 10; i++) {
        p.gaps[i] = gaps[i] > 0 && p.n_gaps == i ? gaps[i] : 0;
        if (gaps[i] > 0 && p.n_gaps == i) p.n_gaps++;
    }
    p.gap_size = settings.getGapSize();
    p.cluster_gap = settings.getClusterGap();
    p.anchor_points = settings.getAnchorPoints();
    p.mult_hyp = (int)settings.getMultHypCorMethod();
    p.q_value = settings.getQValue();
    p.prob_cutoff = settings.getProbCutoff();
    check(iadh_set_params(B.ctx, &p), "iadh_set_params");
}

// masks are mutated by the reference between searches (higherLevel.cpp:165-171), only on the x- and
// y-list of the multiplicon under evaluation: forward the delta of those two lists
void sync_masks(const std::vector<GeneList*>& genelists, int lx, int ly)
{
    for (int pass = 0; pass < 2; pass++) {
        const int l = pass ? ly : lx;
        if (l < 0 || l >= (int)genelists.size() || (pass && ly == lx)) continue;
        const std::vector<ListElement*>& re = genelists[l]->getRemappedElements();
        std::vector<unsigned char>& sh = B.masked[l];
        size_t i = 0;
        while (i < re.size()) {
            if (re[i]->isMasked() && !sh[i]) {
                size_t j = i;
                while (j + 1 < re.size() && re[j + 1]->isMasked() && !sh[j + 1]) j++;
                check(iadh_mask_range(B.ctx, (int32_t)l, (int32_t)i, (int32_t)j), "iadh_mask_range");
                for (size_t k = i; k <= j; k++) sh[k] = 1;
                i = j + 1;
            } else
                i++;
        }
    }
}

// the return value of Multiplicon::createHomologs(useFamily, minHomologs) (Multiplicon.cpp:352-382),
// read from the hasHomolog flags the unpack constructor's createHomologs(useFamily, 0) left behind
bool enough_homologs(const Multiplicon& m, int minHomologs)
{
    const std::vector<GeneList*>& xs = m.getXSegments();
    for (size_t i = 0; i < xs.size(); i++) {
        int count = 0;
        for (unsigned j = 0; j < xs[i]->getSize(); j++) {
            if (xs[i]->getRemappedElements()[j]->hasHomolog) count++;
            if (count >= minHomologs) break;
        }
        if (count < minHomologs) return false;
    }
    const GeneList* ys = m.getYSegment();
    for (unsigned count = 0, j = 0; j < ys->getSize(); j++) {
        if (ys->getRemappedElements()[j]->hasHomolog) count++;
        if ((int)count >= minHomologs) return true;
    }
    return false;
}

// Cheap, conservative form of the homolog filter, applied to the packed stream BEFORE the unpack
// constructors run (they copy S+1 segments and build Link sets per multiplicon, and 40 % of what a search
// returns is then thrown away).  createHomologs (Multiplicon.cpp:352-382) marks an element of x-segment s
// inside [begin_x, end_x] as "has homolog" only if (a) a Link inherited from the parent multiplicon
// survives pruneLinks -- then the element is an end of an inherited Link of segment s -- or (b) addLinks
// finds a same-family element in [begin_y, end_y] of the y-list.  Counting the elements that satisfy the
// weaker "end of ANY inherited Link of segment s, or family present in the y-range" over-estimates the
// count, so a segment that stays below minHomologs is certain to fail the exact test.  When the parent
// has no Links, pruneLinks leaves the copied flags alone (Multiplicon.cpp:386-388): nothing is dropped.
// Family mode only.  Returns the filtered stream (same layout) in `out`.
struct ProfileLinks {
    bool any = false;
    std::vector<unsigned char> linked;  // S*L: element is an end of an inherited Link of its segment
};

void profile_links(const Profile* profile, int S, int L, const std::vector<int32_t>& num, ProfileLinks& pl)
{
    const std::set<Link>& hs = profile->getMultiplicon().getHomologs();
    pl.any = !hs.empty();
    pl.linked.assign((size_t)S * L, 0);
    if (!pl.any) return;
    std::vector<std::map<int, int> > col(S);   // numID -> column, per segment
    for (int s = 0; s < S; s++)
        for (int x = 0; x < L; x++)
            if (num[(size_t)s * L + x] >= 0) col[s][num[(size_t)s * L + x]] = x;
    for (std::set<Link>::const_iterator e = hs.begin(); e != hs.end(); e++) {
        for (int end = 0; end < 2; end++) {
            const int seg = end ? e->segmentY : e->segmentX;
            const int id = end ? e->geneYID : e->geneXID;
            if (seg < 0 || seg >= S) continue;
            std::map<int, int>::const_iterator it = col[seg].find(id);
            if (it != col[seg].end()) pl.linked[(size_t)seg * L + it->second] = 1;
        }
    }
}

long long prefilter_packed(const uint8_t* bytes, size_t nbytes, int S, int L, const std::vector<int32_t>& fam,
                           const std::vector<uint8_t>& flags, const ProfileLinks& pl, int minHomologs,
                           std::vector<char>& out)
{
    int32_t n = 0;
    memcpy(&n, bytes, 4);
    out.assign(4, 0);
    int32_t kept = 0;
    size_t o = 4;
    std::vector<int32_t> yfam;
    for (int32_t m = 0; m < n; m++) {
        const size_t begin = o;
        int32_t hdr[3], b[4], ncl;
        memcpy(hdr, bytes + o, 12);          // x object, y object, level
        memcpy(b, bytes + o + 13, 16);       // begin_x, end_x, begin_y, end_y (after the 1-byte redundant flag)
        memcpy(&ncl, bytes + o + 29, 4);
        o += 33;
        for (int32_t c = 0; c < ncl; c++) {
            int32_t nap;
            memcpy(&nap, bytes + o + 98, 4);
            o += 102 + (size_t)17 * nap;
        }
        bool drop = false;
        const int yl = hdr[1];
        if (pl.any && yl >= 0 && yl < (int)B.list_fam.size() && b[0] >= 0 && b[1] < L && b[2] >= 0 &&
            b[3] < (int)B.list_fam[yl].size()) {
            yfam.assign(B.list_fam[yl].begin() + b[2], B.list_fam[yl].begin() + b[3] + 1);
            std::sort(yfam.begin(), yfam.end());
            for (int s = 0; s < S && !drop; s++) {
                int count = 0;
                for (int x = b[0]; x <= b[1] && count < minHomologs; x++) {
                    const size_t k = (size_t)s * L + x;
                    if (flags[k] & IADH_F_GAP) continue;
                    if (pl.linked[k] || std::binary_search(yfam.begin(), yfam.end(), fam[k])) count++;
                }
                if (count < minHomologs) drop = true;
            }
        }
        if (drop) continue;
        out.insert(out.end(), (const char*)bytes + begin, (const char*)bytes + o);
        kept++;
    }
    memcpy(out.data(), &kept, 4);
    (void)nbytes;
    return n - kept;
}

// Multi-GPU (the reference's MPI build, one rank per GPU of the box): every rank searches its share
// (rank / n_ranks arguments of the library calls: the reference's own partitions, levelTwo.cpp:98-111 and
// parallel.cpp:96-156) and the packed multiplicons of all ranks are exchanged inside the library over NCCL
// (iadh_allgather_results) -- the MPI_Allgather + MPI_Allgatherv of levelTwo.cpp:225-269 / higherLevel.cpp:279-314.
// MPI is only used to hand rank 0's NCCL id to the others.  This block needs mpi.h: it is compiled in the
// reference's HAVE_MPI configuration only (not available in the authoring image, where bench.py / tools/dist_check.py
// drive the same library calls through torch.distributed instead).
void join_ranks()
{
#ifdef HAVE_MPI
    if (B.n_ranks > 1 || (int)ParToolBox::getNumProcesses() <= 1) return;
    B.rank = ParToolBox::getProcID();
    B.n_ranks = (int)ParToolBox::getNumProcesses();
    uint8_t id[128];
    if (B.rank == 0) check(iadh_comm_unique_id(id, sizeof id), "iadh_comm_unique_id");
    MPI_Bcast(id, (int)sizeof id, MPI_BYTE, 0, MPI_COMM_WORLD);
    check(iadh_comm_init(B.ctx, id, sizeof id, B.rank, B.n_ranks), "iadh_comm_init");
#endif
}
// this rank's result -> all ranks' results in rank order (single rank: unchanged)
iadh_result* gather(iadh_result* mine)
{
#ifdef HAVE_MPI
    if (B.n_ranks > 1) {
        iadh_result* all = NULL;
        check(iadh_allgather_results(B.ctx, &mine, 1, &all), "iadh_allgather_results");
        iadh_free_result(mine);
        return all;
    }
#endif
    return mine;
}

}  // namespace

// ---- replaces levelTwo.cpp:177-292 --------------------------------------------------------------
void DataSet::parallelLevel2ADHoReDyn()
{
    if (settings.getClusterType() != Collinear)
        throw std::runtime_error("the B200 path implements collinear detection only (cluster_type=collinear)");
    ensure_context(genelists, settings, genepairs);
    join_ranks();
    iadh_result* res = NULL;
    check(iadh_level2_all_pairs(B.ctx, B.rank, B.n_ranks, &res), "iadh_level2_all_pairs");
    res = gather(res);
    size_t nbytes = 0;
    const uint8_t* bytes = iadh_result_bytes(res, &nbytes);
    std::vector<Multiplicon*> mplicons;
    Multiplicon::unpackL2Multiplicons((const char*)bytes, mplicons, genelists, settings.useFamily());
    iadh_stats st;
    iadh_result_stats(res, &st);
    B.t_l2 += st.total_ms;
    iadh_free_result(res);
    addMultiplicons(mplicons);
    sortByMultipliconSize(multiplicons);
    if (getenv("IADH_VERBOSE"))
        fprintf(stderr, "[iadh] level 2: %lld GHMs, %lld dots, %lld multiplicons, device %.2f ms\n", (long long)st.n_ghm,
                (long long)st.n_dots, (long long)st.n_multiplicons, st.total_ms);
}

// ---- replaces higherLevel.cpp:252-326 -------------------------------------------------------------
void DataSet::parallelProfileSearch(const Profile* profile, std::vector<Multiplicon*>& target)
{
    profileUC = profile;
    const double w0 = wall();
    sync_masks(genelists, profile->getMultiplicon().getXObjectID(), profile->getMultiplicon().getYObjectID());
    const double w1 = wall();
    const std::vector<GeneList*>& segs = profile->getSegments();
    const int S = (int)segs.size();
    const int L = S ? (int)segs[0]->getRemappedElements().size() : 0;
    std::vector<int32_t> fam((size_t)S * L, -1), gid((size_t)S * L, -1), num((size_t)S * L, -1);
    std::vector<uint8_t> flags((size_t)S * L, IADH_F_GAP);
    for (int s = 0; s < S; s++) {
        const std::vector<ListElement*>& e = segs[s]->getRemappedElements();
        for (int x = 0; x < L; x++) {
            if (e[x]->isGap()) continue;
            const Gene& g = e[x]->getGene();
            const size_t k = (size_t)s * L + x;
            const int id = e[x]->getNumID();
            if (id >= 0 && id < (int)B.fam_of.size() && B.fam_of[id] >= 0) {
                fam[k] = B.fam_of[id];
                gid[k] = B.gid_of[id];
            } else {
                fam[k] = intern(B.fam_ids, g.getFamily());
                gid[k] = intern(B.gene_ids, g.getID());
            }
            num[k] = id;
            flags[k] = e[x]->getOrientation() ? IADH_F_STRAND : 0;
        }
    }
    iadh_profile_desc pd;
    pd.n_segments = S;
    pd.length = L;
    pd.x_object_id = segs[0]->getID();
    pd.fam = fam.data();
    pd.gid = gid.data();
    pd.num_id = num.data();
    pd.flags = flags.data();
    iadh_result* res = NULL;
    // The homolog-count filter of GHMProfile::setMultiplicons (GHMProfile.cpp:129-132) runs on the device when the
    // library gets the Link set the search inherits (Multiplicon.cpp:343-344): the stream then holds exactly the
    // multiplicons the reference's search returns.  IADH_SHIM_HOST_FILTER=1 keeps the filter on the host (the
    // unfiltered entry point + the code below), IADH_SHIM_VERIFY_FILTER=1 re-checks every returned multiplicon.
    static const bool host_filter = getenv("IADH_SHIM_HOST_FILTER") != NULL;
    static const bool verify_filter = getenv("IADH_SHIM_VERIFY_FILTER") != NULL;
    const std::set<Link>& inherited = profile->getMultiplicon().getHomologs();
    const bool device_filter = !host_filter && !inherited.empty();
    std::vector<iadh_link_t> links;
    if (device_filter) {
        links.reserve(inherited.size());
        for (std::set<Link>::const_iterator e = inherited.begin(); e != inherited.end(); e++) {
            iadh_link_t l;
            l.segment_x = e->segmentX;
            l.segment_y = e->segmentY;
            l.gene_x_id = e->geneXID;
            l.gene_y_id = e->geneYID;
            links.push_back(l);
        }
    }
    const double w2 = wall();
    if (device_filter)
        check(iadh_profile_search_filtered(B.ctx, &pd, links.data(), (int64_t)links.size(), B.rank, B.n_ranks, &res), "iadh_profile_search_filtered");
    else
        check(iadh_profile_search(B.ctx, &pd, B.rank, B.n_ranks, &res), "iadh_profile_search");
    res = gather(res);
    const double w3 = wall();
    size_t nbytes = 0;
    const uint8_t* bytes = iadh_result_bytes(res, &nbytes);
    std::vector<Multiplicon*> mplicons;
    static const bool use_prefilter = !getenv("IADH_SHIM_NO_PREFILTER");
    if (device_filter) {
        // (the unpack constructor's createHomologs(useFamily, 0) rebuilds the Link set of every kept multiplicon)
        Multiplicon::unpackHLMultiplicons((const char*)bytes, mplicons, *profile, genelists, settings.useFamily());
        if (verify_filter)
            for (size_t i = 0; i < mplicons.size(); i++)
                if (!enough_homologs(*mplicons[i], settings.getAnchorPoints()))
                    throw std::runtime_error("IADH_SHIM_VERIFY_FILTER: the device kept a multiplicon the host filter drops");
    } else if (use_prefilter && settings.useFamily() && nbytes > 4) {
        ProfileLinks pl;
        profile_links(profile, S, L, num, pl);
        std::vector<char> filtered;
        B.n_prefiltered += prefilter_packed(bytes, nbytes, S, L, fam, flags, pl, settings.getAnchorPoints(), filtered);
        Multiplicon::unpackHLMultiplicons(filtered.data(), mplicons, *profile, genelists, settings.useFamily());
    } else
        Multiplicon::unpackHLMultiplicons((const char*)bytes, mplicons, *profile, genelists, settings.useFamily());
    iadh_stats st;
    iadh_result_stats(res, &st);
    B.t_hl += st.total_ms;
    B.n_hl++;
    iadh_free_result(res);
    // the homolog-count filter of GHMProfile::setMultiplicons (GHMProfile.cpp:129-132), unless the device applied it
    for (size_t i = 0; i < mplicons.size(); i++) {
        if (device_filter || enough_homologs(*mplicons[i], settings.getAnchorPoints())) target.push_back(mplicons[i]);
        else delete mplicons[i];
    }
    B.n_ghm += st.n_ghm;
    B.n_found += (long long)target.size();
    B.n_returned += (long long)st.n_multiplicons;
    B.w_mask += w1 - w0;
    B.w_desc += w2 - w1;
    B.w_call += w3 - w2;
    B.w_unpack += wall() - w3;
}
