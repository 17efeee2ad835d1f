// TEST INFRASTRUCTURE -- not part of the product (see oracle_core.h).
//
// The oracle behind the same C interface as include/iadhore_b200.h (prefix oracle_
// instead of iadh_), so a parity test drives both with identical calls.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../include/iadhore_b200.h"
#include "oracle_core.h"

using namespace oracle;

struct oracle_ctx {
    std::vector<List> lists;
    std::vector<int> index_to_list;
    Params prm;
    int n_threads = 1;
    PairTable pairs;       // table_type=pairs (oracle_create_pairs)
    bool pairs_mode = false;
    const PairTable* pt() const { return pairs_mode ? &pairs : nullptr; }
};

struct oracle_result {
    std::vector<uint8_t> bytes;
    std::vector<int32_t> ghm_index;
    iadh_stats stats;
    // build_dots scratch
    std::vector<int64_t> cnt;
    std::vector<int32_t> dots;
};

static thread_local std::string g_err;

// DataSet.cpp:1088-1105: multimap<weight,list> walked in reverse => size descending,
// equal sizes in descending list index (T10)
static void sort_lists(oracle_ctx* c)
{
    std::multimap<uint64_t, unsigned> w;
    for (unsigned l = 0; l < c->lists.size(); l++) w.insert(std::make_pair((uint64_t)c->lists[l].size(), l));
    c->index_to_list.clear();
    for (std::multimap<uint64_t, unsigned>::reverse_iterator it = w.rbegin(); it != w.rend(); ++it)
        c->index_to_list.push_back(it->second);
}

// levelTwo.cpp:113-175 (pair walk) + 98-111 (least-loaded rank)
static void pair_sequence(const oracle_ctx* c, int rank, int n_ranks, std::vector<std::pair<int, int>>& out)
{
    unsigned n = c->lists.size();
    std::vector<uint64_t> wpp(n_ranks, 0);
    unsigned cX = 0, cY = 0;
    while (cX < n) {
        unsigned lX = c->index_to_list[cX], lY = c->index_to_list[cY];
        uint64_t weight = (unsigned)c->lists[lX].size() * (unsigned)c->lists[lY].size();
        unsigned sp = 0;
        for (unsigned p = 0; p < wpp.size(); p++)
            if (wpp[p] < wpp[sp]) sp = p;
        wpp[sp] += weight;
        if ((int)sp == rank) out.push_back(std::make_pair((int)lX, (int)lY));
        if ((cX - cY) < 2) {
            if (cX == cY && cX == n - 1) break;
            unsigned newD = cX + cY + 1;
            cX = (newD < n) ? newD : n - 1;
            cY = (newD < n) ? 0 : newD - n + 1;
        } else {
            cX--;
            cY++;
        }
    }
}

// parallel.cpp:96-156: LPT partition of the size-sorted lists over n_ranks packages
static void list_partition(const oracle_ctx* c, int rank, int n_ranks, std::vector<int>& out)
{
    std::multimap<uint64_t, unsigned> w;
    for (unsigned l = 0; l < c->lists.size(); l++) w.insert(std::make_pair((uint64_t)c->lists[l].size(), l));
    std::vector<uint64_t> wpp(n_ranks, 0);
    for (std::multimap<uint64_t, unsigned>::reverse_iterator it = w.rbegin(); it != w.rend(); ++it) {
        unsigned sp = 0;
        for (unsigned p = 0; p < wpp.size(); p++)
            if (wpp[p] < wpp[sp]) sp = p;
        wpp[sp] += it->first;
        if ((int)sp == rank) out.push_back((int)it->second);
    }
}

extern "C" {

int oracle_create(const iadh_dataset_desc* ds, int /*device*/, oracle_ctx** out)
{
    if (!ds || !out || ds->n_lists < 0) { g_err = "bad dataset"; return IADH_ERR_ARG; }
    oracle_ctx* c = new oracle_ctx();
    c->lists.resize(ds->n_lists);
    for (int l = 0; l < ds->n_lists; l++) {
        List& L = c->lists[l];
        L.id = l;
        for (int64_t i = ds->list_off[l]; i < ds->list_off[l + 1]; i++) {
            L.fam.push_back(ds->fam[i]);
            L.gid.push_back(ds->gid[i]);
            L.num_id.push_back(ds->num_id[i]);
            L.strand.push_back((ds->flags[i] & IADH_F_STRAND) ? 1 : 0);
            L.masked.push_back((ds->flags[i] & IADH_F_MASKED) ? 1 : 0);
        }
    }
    sort_lists(c);
    c->prm.n_gaps = gap_sizes(30, c->prm.gaps);
    *out = c;
    return IADH_OK;
}

// table_type=pairs counterpart of oracle_create: the family ids of the descriptor are ignored
int oracle_create_pairs(const iadh_dataset_desc* ds, const int64_t* pair_ptr, const int32_t* pair_ids, int32_t n_genes,
                        int device, oracle_ctx** out)
{
    if (!pair_ptr || n_genes < 0 || (pair_ptr[n_genes] > 0 && !pair_ids)) { g_err = "bad pair table"; return IADH_ERR_ARG; }
    iadh_dataset_desc d = *ds;
    std::vector<int32_t> zero;
    if (!d.fam) {
        zero.assign((size_t)(d.n_lists ? d.list_off[d.n_lists] : 0), 0);
        d.fam = zero.data();
    }
    int rc = oracle_create(&d, device, out);
    if (rc) return rc;
    (*out)->pairs.ptr.assign(pair_ptr, pair_ptr + n_genes + 1);
    (*out)->pairs.ids.assign(pair_ids, pair_ids + pair_ptr[n_genes]);
    (*out)->pairs_mode = true;
    return IADH_OK;
}

void oracle_destroy(oracle_ctx* c) { delete c; }

int oracle_set_threads(oracle_ctx* c, int n) { c->n_threads = n < 1 ? 1 : n; return IADH_OK; }

int oracle_set_params(oracle_ctx* c, const iadh_params* p)
{
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

}  // extern "C"

struct GhmOut {
    std::vector<uint8_t> bytes;
    int n_mult = 0;
    int64_t dots = 0, ncl = 0, nap = 0;
    double build_s = 0, run_s = 0;
};

static double now_s()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static void run_pair(const oracle_ctx* c, int lx, int ly, GhmOut& o)
{
    GHMInput g;
    double t0 = now_s();
    build_l2(c->lists[lx], c->lists[ly], lx == ly, g, c->pt());
    double t1 = now_s();
    o.dots = g.cnt[0] + g.cnt[1];
    std::vector<Cluster> cl;
    std::vector<Multiplicon> ms;
    run_collinear(g, c->prm, cl, ms);
    set_multiplicons_l2(c->lists[lx], c->lists[ly], cl, ms);
    o.build_s = t1 - t0;
    o.run_s = now_s() - t1;
    pack(cl, ms, o.bytes, o.n_mult);
    for (const Multiplicon& m : ms) {
        o.ncl += m.clusters.size();
        for (int ci : m.clusters) o.nap += cl[ci].aps.size();
    }
}

static void run_profile(const oracle_ctx* c, const Profile& P, int ly, GhmOut& o, const std::vector<Link>* links = nullptr)
{
    GHMInput g;
    double t0 = now_s();
    build_hl(P, c->lists[ly], g, c->pt());
    double t1 = now_s();
    o.dots = g.cnt[0] + g.cnt[1];
    std::vector<Cluster> cl;
    std::vector<Multiplicon> ms;
    run_collinear(g, c->prm, cl, ms);
    set_multiplicons_hl(P, c->lists[ly], cl, ms, c->pt());
    if (links) filter_homologs(P, c->lists[ly], *links, c->prm.anchor_points, ms, c->pt());   // GHMProfile.cpp:129-132
    o.build_s = t1 - t0;
    o.run_s = now_s() - t1;
    pack(cl, ms, o.bytes, o.n_mult);
    for (const Multiplicon& m : ms) {
        o.ncl += m.clusters.size();
        for (int ci : m.clusters) o.nap += cl[ci].aps.size();
    }
}

template <typename F>
static oracle_result* run_many(const oracle_ctx* c, size_t n, F fn)
{
    std::vector<GhmOut> outs(n);
    double t0 = now_s();
    int nt = c->n_threads;
    if (nt <= 1 || n < 2) {
        for (size_t i = 0; i < n; i++) fn(i, outs[i]);
    } else {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; t++)
            th.emplace_back([&, t]() {
                for (size_t i = t; i < n; i += nt) fn(i, outs[i]);
            });
        for (std::thread& t : th) t.join();
    }
    double t1 = now_s();
    oracle_result* r = new oracle_result();
    memset(&r->stats, 0, sizeof(r->stats));
    int32_t total = 0;
    r->bytes.resize(4);
    for (size_t i = 0; i < n; i++) {
        r->bytes.insert(r->bytes.end(), outs[i].bytes.begin(), outs[i].bytes.end());
        for (int k = 0; k < outs[i].n_mult; k++) r->ghm_index.push_back((int32_t)i);
        total += outs[i].n_mult;
        r->stats.n_dots += outs[i].dots;
        r->stats.n_baseclusters += outs[i].ncl;
        r->stats.n_anchorpoints += outs[i].nap;
        r->stats.build_ms += outs[i].build_s * 1e3;
        r->stats.cluster_ms += outs[i].run_s * 1e3;
    }
    memcpy(r->bytes.data(), &total, 4);
    r->stats.n_ghm = (int64_t)n;
    r->stats.n_multiplicons = total;
    r->stats.total_ms = (t1 - t0) * 1e3;
    return r;
}

extern "C" {

int oracle_level2_pairs(oracle_ctx* c, const int32_t* xs, const int32_t* ys, int64_t n, oracle_result** out)
{
    *out = run_many(c, (size_t)n, [&](size_t i, GhmOut& o) { run_pair(c, xs[i], ys[i], o); });
    return IADH_OK;
}

int oracle_level2_all_pairs(oracle_ctx* c, int rank, int n_ranks, oracle_result** out)
{
    std::vector<std::pair<int, int>> seq;
    pair_sequence(c, rank, n_ranks, seq);
    *out = run_many(c, seq.size(), [&](size_t i, GhmOut& o) { run_pair(c, seq[i].first, seq[i].second, o); });
    return IADH_OK;
}

int oracle_mask_range(oracle_ctx* c, int32_t list, int32_t begin, int32_t end)
{
    if (list < 0 || list >= (int)c->lists.size()) { g_err = "bad list"; return IADH_ERR_ARG; }
    for (int i = begin; i <= end; i++) c->lists[list].masked[i] = 1;
    return IADH_OK;
}

int oracle_clear_masks(oracle_ctx* c)
{
    for (List& l : c->lists) std::fill(l.masked.begin(), l.masked.end(), 0);
    return IADH_OK;
}

}  // extern "C"
static void to_profile(const iadh_profile_desc* p, Profile& P)
{
    P.S = p->n_segments;
    P.L = p->length;
    P.x_object_id = p->x_object_id;
    size_t n = (size_t)P.S * P.L;
    P.fam.assign(p->fam, p->fam + n);
    P.gid.assign(p->gid, p->gid + n);
    P.num_id.assign(p->num_id, p->num_id + n);
    P.flags.assign(p->flags, p->flags + n);
}

static int profile_search(oracle_ctx* c, const iadh_profile_desc* p, const std::vector<Link>* links, int rank, int n_ranks,
                          oracle_result** out)
{
    Profile P;
    to_profile(p, P);
    std::vector<int> part, lists;
    list_partition(c, rank, n_ranks, part);
    for (int l : part) {  // higherLevel.cpp:40-41
        unsigned masked = 0;
        for (uint8_t m : c->lists[l].masked) masked += m;
        if ((unsigned)c->lists[l].size() - masked >= (unsigned)c->prm.anchor_points) lists.push_back(l);
    }
    *out = run_many(c, lists.size(), [&](size_t i, GhmOut& o) { run_profile(c, P, lists[i], o, links); });
    return IADH_OK;
}
static int to_links(const iadh_link_t* links, int64_t n, std::vector<Link>& out)
{
    if (!links || n <= 0) { g_err = "the homolog filter needs the Link set of the profile's multiplicon"; return IADH_ERR_ARG; }
    out.resize((size_t)n);
    for (int64_t i = 0; i < n; i++) {
        out[i].seg_x = links[i].segment_x; out[i].seg_y = links[i].segment_y;
        out[i].gene_x = links[i].gene_x_id; out[i].gene_y = links[i].gene_y_id;
    }
    return IADH_OK;
}

extern "C" {
int oracle_profile_search(oracle_ctx* c, const iadh_profile_desc* p, int rank, int n_ranks, oracle_result** out)
{
    return profile_search(c, p, nullptr, rank, n_ranks, out);
}
// iadh_profile_search_filtered restated: the search with the homolog-count filter of GHMProfile.cpp:129-132 applied
int oracle_profile_search_filtered(oracle_ctx* c, const iadh_profile_desc* p, const iadh_link_t* links, int64_t n_links, int rank,
                                   int n_ranks, oracle_result** out)
{
    std::vector<Link> lk;
    if (int rc = to_links(links, n_links, lk)) return rc;
    return profile_search(c, p, &lk, rank, n_ranks, out);
}

// iadh_profile_search_batch restated: by definition the sequential calls (higherLevel.cpp:165-178 per iteration:
// mask the segments, then search).  shard_profiles: profile k is searched against all lists by rank k % n_ranks.
static int profile_search_batch(oracle_ctx* c, const iadh_profile_desc* profiles, int32_t n_profiles, const iadh_link_t* links,
                                const int64_t* link_ptr, const iadh_mask_range_t* masks, const int32_t* mask_ptr, int rank,
                                int n_ranks, int shard_profiles, oracle_result** out);
int oracle_profile_search_batch(oracle_ctx* c, const iadh_profile_desc* profiles, int32_t n_profiles, const iadh_mask_range_t* masks,
                                const int32_t* mask_ptr, int rank, int n_ranks, int shard_profiles, oracle_result** out)
{
    return profile_search_batch(c, profiles, n_profiles, nullptr, nullptr, masks, mask_ptr, rank, n_ranks, shard_profiles, out);
}
int oracle_profile_search_batch_filtered(oracle_ctx* c, const iadh_profile_desc* profiles, int32_t n_profiles, const iadh_link_t* links,
                                         const int64_t* link_ptr, const iadh_mask_range_t* masks, const int32_t* mask_ptr, int rank,
                                         int n_ranks, int shard_profiles, oracle_result** out)
{
    if (!links || !link_ptr) { g_err = "the homolog filter needs the Link sets"; return IADH_ERR_ARG; }
    return profile_search_batch(c, profiles, n_profiles, links, link_ptr, masks, mask_ptr, rank, n_ranks, shard_profiles, out);
}
}  // extern "C"
static int profile_search_batch(oracle_ctx* c, const iadh_profile_desc* profiles, int32_t n_profiles, const iadh_link_t* links,
                                const int64_t* link_ptr, const iadh_mask_range_t* masks, const int32_t* mask_ptr, int rank,
                                int n_ranks, int shard_profiles, oracle_result** out)
{
    for (int k = 0; k < n_profiles; k++) {
        for (int m = mask_ptr[k]; m < mask_ptr[k + 1]; m++) {
            const int rc = oracle_mask_range(c, masks[m].list, masks[m].begin, masks[m].end);
            if (rc) return rc;
        }
        if (shard_profiles && k % n_ranks != rank) {
            *(out + k) = run_many(c, 0, [&](size_t, GhmOut&) {});
            continue;
        }
        std::vector<Link> lk;
        if (links)
            if (int rc = to_links(links + link_ptr[k], link_ptr[k + 1] - link_ptr[k], lk)) return rc;
        const int rc = shard_profiles ? profile_search(c, profiles + k, links ? &lk : nullptr, 0, 1, out + k)
                                      : profile_search(c, profiles + k, links ? &lk : nullptr, rank, n_ranks, out + k);
        if (rc) return rc;
    }
    return IADH_OK;
}
extern "C" {

// The exchange of the multi-GPU product (iadh_comm_*, iadh_allgather_results) has no CPU counterpart to check:
// the oracle is a single process.  With one rank the gathered stream is the rank's own stream, which is all
// these entry points provide (so that host code written against the interface links against either library).
int oracle_comm_unique_id(uint8_t* id_out, size_t n_bytes)
{
    if (id_out) memset(id_out, 0, n_bytes);
    return IADH_OK;
}
int oracle_comm_init(oracle_ctx*, const uint8_t*, size_t, int rank, int n_ranks)
{
    if (rank != 0 || n_ranks != 1) { g_err = "the oracle is a single process: one rank only"; return IADH_ERR_STATE; }
    return IADH_OK;
}
int oracle_comm_destroy(oracle_ctx*) { return IADH_OK; }
int oracle_allgather_results(oracle_ctx*, oracle_result* const* mine, int32_t n, oracle_result** all)
{
    for (int i = 0; i < n; i++) {
        all[i] = new oracle_result();
        all[i]->bytes = mine[i]->bytes;
        all[i]->stats = mine[i]->stats;
    }
    return IADH_OK;
}

const uint8_t* oracle_result_bytes(const oracle_result* r, size_t* n)
{
    *n = r->bytes.size();
    return r->bytes.data();
}
size_t oracle_result_size(const oracle_result* r) { return r->bytes.size(); }
const int32_t* oracle_result_ghm_index(const oracle_result* r, size_t* n)
{
    *n = r->ghm_index.size();
    return r->ghm_index.data();
}
int oracle_result_stats(const oracle_result* r, iadh_stats* s) { *s = r->stats; return IADH_OK; }
void oracle_free_result(oracle_result* r) { delete r; }
void oracle_trim(void) {}
int oracle_warmup(int) { return IADH_OK; }

// GeneList::remapTandemsPairs (GeneList.cpp:84-141), restated sequentially: the three passes in the reference's order
// with Gene::isIndirectPairWith (Gene.cpp:33-49) over the CSR pair sets; a gene without pairs is never remapped.
int oracle_remap_tandems_pairs(const iadh_dataset_desc* d, const int64_t* pp, const int32_t* pi, int32_t n_genes, int32_t gap, int,
                               uint8_t* r)
{
    auto has = [&](int g) { return g >= 0 && g < n_genes && pp[g + 1] > pp[g]; };
    auto indirect = [&](int ga, int gb) {
        if (gb < 0 || gb >= n_genes) return false;
        if (std::binary_search(pi + pp[ga], pi + pp[ga + 1], gb)) return true;
        int64_t a = pp[ga], b = pp[gb];
        while (a < pp[ga + 1] && b < pp[gb + 1]) {
            if (pi[a] == pi[b]) return true;
            if (pi[a] < pi[b]) a++;
            else b++;
        }
        return false;
    };
    for (int l = 0; l < d->n_lists; l++) {
        const int64_t b = d->list_off[l], e = d->list_off[l + 1];
        const int n = (int)(e - b);
        const int32_t* gid = d->gid + b;
        const uint8_t* fl = d->flags + b;
        uint8_t* rem = r + b;
        for (int i = 0; i < n; i++) rem[i] = 0;
        for (int i = n - 1; i > 0; i--) {           // '+' onto earlier '+'
            if (!has(gid[i]) || !(fl[i] & 1)) continue;
            for (int j = i - 1; j >= 0 && i - j <= gap + 1; j--) {
                if (!(fl[j] & 1)) continue;
                if (indirect(gid[i], gid[j])) { rem[i] = 1; rem[j] = 0; break; }
            }
        }
        for (int i = 0; i < n - 1; i++) {           // '-' onto later '-'
            if (!has(gid[i]) || (fl[i] & 1)) continue;
            for (int j = i + 1; j < n && j - i <= gap + 1; j++) {
                if (fl[j] & 1) continue;
                if (indirect(gid[i], gid[j])) { rem[i] = 1; rem[j] = 0; break; }
            }
        }
        for (int i = n - 1; i > 0; i--) {           // anything left onto any earlier gene
            if (!has(gid[i]) || rem[i]) continue;
            for (int j = i - 1; j >= 0 && i - j <= gap + 1; j--)
                if (indirect(gid[i], gid[j])) { rem[i] = 1; rem[j] = 0; break; }
        }
    }
    return IADH_OK;
}

// Interning as the reference's string-keyed containers imply it (GeneFamily.cpp:12-16 map<string, string>,
// GHM.cpp:34-50 family-string hash): equal strings, equal ids; ids in order of first occurrence.
int oracle_intern_strings(const char* blob, const int64_t* off, int64_t n, int, int32_t* ids, int32_t* n_distinct)
{
    std::unordered_map<std::string, int32_t> seen;
    for (int64_t i = 0; i < n; i++) {
        const std::string s(blob + off[i], (size_t)(off[i + 1] - off[i]));
        auto it = seen.find(s);
        if (it == seen.end()) it = seen.emplace(s, (int32_t)seen.size()).first;
        ids[i] = it->second;
    }
    *n_distinct = (int32_t)seen.size();
    return IADH_OK;
}

// GeneList::remapTandemsFamily (GeneList.cpp:143-200), restated sequentially: the three passes in
// the reference's order, remapTo() (Gene.cpp:51-62) sets the gene's flag and clears the target's.
int oracle_remap_tandems(const iadh_dataset_desc* d, int32_t gap, int, uint8_t* r)
{
    for (int l = 0; l < d->n_lists; l++) {
        const int64_t b = d->list_off[l], e = d->list_off[l + 1];
        const int n = (int)(e - b);
        const int32_t* fam = d->fam + b;
        const int32_t* gid = d->gid + b;
        const uint8_t* fl = d->flags + b;
        uint8_t* rem = r + b;
        for (int i = 0; i < n; i++) rem[i] = 0;
        auto pair = [&](int i, int j) { return fam[i] == fam[j] && gid[i] != gid[j]; };
        for (int i = n - 1; i > 0; i--) {           // '+' onto earlier '+'
            if (!(fl[i] & 1)) continue;
            for (int j = i - 1; j >= 0 && i - j <= gap + 1; j--) {
                if (!(fl[j] & 1)) continue;
                if (pair(i, j)) { rem[i] = 1; rem[j] = 0; break; }
            }
        }
        for (int i = 0; i < n - 1; i++) {           // '-' onto later '-'
            if (fl[i] & 1) continue;
            for (int j = i + 1; j < n && j - i <= gap + 1; j++) {
                if (fl[j] & 1) continue;
                if (pair(i, j)) { rem[i] = 1; rem[j] = 0; break; }
            }
        }
        for (int i = n - 1; i > 0; i--) {           // anything still unremapped onto any earlier gene
            if (rem[i]) continue;
            for (int j = i - 1; j >= 0 && i - j <= gap + 1; j--)
                if (pair(i, j)) { rem[i] = 1; rem[j] = 0; break; }
        }
    }
    return IADH_OK;
}
// host-clock stand-ins of the device event hooks
static std::chrono::steady_clock::time_point g_ev[4];
int oracle_event_record(oracle_ctx*, int slot)
{
    if (slot < 0 || slot >= 4) return IADH_ERR_ARG;
    g_ev[slot] = std::chrono::steady_clock::now();
    return IADH_OK;
}
int oracle_event_elapsed_ms(oracle_ctx*, int a, int b, double* ms)
{
    if (a < 0 || a >= 4 || b < 0 || b >= 4 || !ms) return IADH_ERR_ARG;
    *ms = std::chrono::duration<double, std::milli>(g_ev[b] - g_ev[a]).count();
    return IADH_OK;
}

static oracle_result g_scratch;
int oracle_build_dots(oracle_ctx* c, const iadh_profile_desc* p, const int32_t* xs, const int32_t* ys,
                      int64_t n, const int64_t** cnt_out, const int32_t** dots_out, int64_t* n_dots_out)
{
    Profile P;
    if (p) to_profile(p, P);
    g_scratch.cnt.clear();
    g_scratch.dots.clear();
    for (int64_t i = 0; i < n; i++) {
        GHMInput g;
        if (p) build_hl(P, c->lists[ys[i]], g, c->pt());
        else build_l2(c->lists[xs[i]], c->lists[ys[i]], xs[i] == ys[i], g, c->pt());
        for (int o = 0; o < 2; o++) {
            g_scratch.cnt.push_back((int64_t)g.dots[o].size());
            for (const std::pair<int, int>& d : g.dots[o]) {
                g_scratch.dots.push_back(d.first);
                g_scratch.dots.push_back(d.second);
            }
        }
    }
    *cnt_out = g_scratch.cnt.data();
    *dots_out = g_scratch.dots.data();
    *n_dots_out = (int64_t)g_scratch.dots.size() / 2;
    return IADH_OK;
}

// CPU counterpart of the build-only measurement hook: GHM::buildMatrix / GHMProfile::buildMatrix
// restated (build_l2 / build_hl), timed on the host; build_ms = wall time of the builds.
int oracle_build_only(oracle_ctx* c, const iadh_profile_desc* p, const int32_t* xs, const int32_t* ys, int64_t n,
                      iadh_stats* out)
{
    Profile P;
    if (p) to_profile(p, P);
    memset(out, 0, sizeof *out);
    const auto t0 = std::chrono::steady_clock::now();
    for (int64_t i = 0; i < n; i++) {
        GHMInput g;
        if (p) build_hl(P, c->lists[ys[i]], g, c->pt());
        else build_l2(c->lists[xs[i]], c->lists[ys[i]], xs[i] == ys[i], g, c->pt());
        out->n_dots += (int64_t)(g.dots[0].size() + g.dots[1].size());
    }
    out->n_ghm = n;
    out->build_ms = out->total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return IADH_OK;
}

int oracle_omxn(oracle_ctx*, const double* x, const int32_t* n, int64_t count, double* out)
{
    for (int64_t i = 0; i < count; i++) out[i] = omxn(x[i], n[i]);
    return IADH_OK;
}

int oracle_pow(oracle_ctx*, const double* x, const int32_t* n, int64_t count, double* out)
{
    for (int64_t i = 0; i < count; i++) out[i] = std::pow(x[i], n[i]);  // BaseCluster.cpp:567
    return IADH_OK;
}

int oracle_gap_sizes(int gap_size, int32_t* gaps) { return gap_sizes(gap_size, gaps); }

static oracle_ctx* sizes_ctx(const int64_t* sizes, int32_t n)
{
    oracle_ctx* c = new oracle_ctx();
    c->lists.resize(n);
    for (int l = 0; l < n; l++) {
        c->lists[l].id = l;
        c->lists[l].fam.assign((size_t)sizes[l], 0);
    }
    sort_lists(c);
    return c;
}

int oracle_plan_lists(const int64_t* sizes, int32_t n_lists, int rank, int n_ranks, int32_t* out, int64_t* n_out)
{
    oracle_ctx* c = sizes_ctx(sizes, n_lists);
    std::vector<int> part;
    list_partition(c, rank, n_ranks, part);
    for (size_t i = 0; i < part.size(); i++) out[i] = part[i];
    *n_out = (int64_t)part.size();
    delete c;
    return IADH_OK;
}

int oracle_plan_pairs(const int64_t* sizes, int32_t n_lists, int rank, int n_ranks, int32_t* out_x, int32_t* out_y,
                      int64_t* n_out)
{
    oracle_ctx* c = sizes_ctx(sizes, n_lists);
    std::vector<std::pair<int, int>> seq;
    pair_sequence(c, rank, n_ranks, seq);
    for (size_t i = 0; i < seq.size(); i++) {
        out_x[i] = seq[i].first;
        out_y[i] = seq[i].second;
    }
    *n_out = (int64_t)seq.size();
    delete c;
    return IADH_OK;
}

const char* oracle_last_error(void) { return g_err.c_str(); }
const char* oracle_version(void) { return "oracle-port-1"; }

}  // extern "C"
