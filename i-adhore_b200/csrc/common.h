// Shared host/device definitions of the B200 collinearity engine.
//
// Vocabulary follows the reference: GHM (gene homology matrix), dots (homologous
// gene pairs, GHM.cpp:34-85), base clusters, anchor points (AP), multiplicons.
// Orientation index: 0 = OPP_ORIENT, 1 = SAME_ORIENT (GHM.h:27).
#pragma once
#include <stdint.h>

#define IADH_MAX_COORD 46340  // BaseCluster.cpp:156-158: x*y is a 32-bit product

// A dot / anchor point is packed as (x << 16) | y; both coordinates < 65536.
#define KEY_X(k) ((int)((k) >> 16))
#define KEY_Y(k) ((int)((k)&0xffffu))
#define MAKE_KEY(x, y) ((((uint32_t)(x)) << 16) | ((uint32_t)(y)))

struct DevParams {
    int n_gaps;
    int gaps[10];
    int gap_size;
    int cluster_gap;
    int anchor_points;
    int mult_hyp;  // 0 none, 1 Bonferroni, 2 FDR
    double q_value;
    double prob_cutoff;
    int coop[4];   // debug: thresholds of the cooperative stats / merge / insert / twist paths
    int rs_squeeze; // test knob: resident budgets that real GHMs outgrow (exercises the fallback to the global path)
};

// One GHM to cluster: either list x list (level 2) or profile x list.
struct GhmTask {
    int32_t x_list;       // L2: x list index; HL: -1
    int32_t y_list;
    int32_t x_object_id;  // Multiplicon x_objectID (GHM.cpp:588 / GHMProfile.cpp:13)
    int32_t y_object_id;
    int32_t level;        // 2 for L2, S+1 for HL
    uint32_t cnt[2];      // count_points[o] after the build (static afterwards)
    uint32_t pad0;
    uint64_t area;        // GHM.cpp:221-236 (32-bit arithmetic, widened)
    int64_t dot_off;      // offset (in keys) of this GHM's dots: [o=0 block][o=1 block]
    int64_t ws_off;       // byte offset of this GHM's clustering workspace
    int64_t ws_bytes;
    int32_t seq;          // position in the caller's GHM sequence
    int32_t status;       // out: 0 ok, 1 AP pool exhausted (host retries with a larger pool)
    int64_t out_bytes;    // out: packed size of this GHM's multiplicons
    int32_t n_mult;       // out
    int32_t n_clusters;   // out
    int32_t n_aps;        // out
    int32_t prof;         // HL batch (iadh_profile_search_batch): index of this GHM's profile; 0 otherwise
    // The two orientations are clustered by two CTAs up to the final enrichment (cluster_kernel.cu):
    int32_t half_mask;    // in: bit o = orientation o has its own queue item (>= 3 dots)
    int32_t half_needed;  // in: number of queue items of this GHM (1 or 2)
    int32_t half_done;    // device: items finished
    int32_t half_nbc[2], half_ncl[2], half_pool[2], half_status[2];  // device: state handed to phase 2
    int32_t half_nk[2];   // device: matrix sizes between the phase-2 launches of the pass-per-launch path
    int32_t pad2;
    // device: seeding state between the launches of the pass-per-launch path (seed_pass in cluster_kernel.cu):
    // smallest gap size with a 3-dot chain, and bit 0 = successor pointers computed, bit 1 = a dot was erased since
    int32_t half_g3[2], half_seed[2];
};

// Per-GHM cycle counters of K2 (debugging / profiling aid; a separate array so that the task
// records that travel between host and device stay small).
struct GhmProf {
    long long half_cycles[2];
    long long prof2[8];   // join sub-phases (init, reduce, locate, merge, erase, refresh, redo, #merges)
    long long prof[8];    // SM cycles per phase (seed, enrich, join, filter, metacluster, critical path, -, -)
};

// Base cluster record (BaseCluster.h:316-331).  Anchor points live in the GHM's AP
// pool at [off, off+n): keys in multiset order (x ascending, insertion order on ties,
// AnchorPoint.h:144-156) with a parallel backbone flag byte.
struct ClusterRec {
    double a, b, avg_x, var_x, mrss, x_end1, x_end2, y_end1, y_end2;
    double p;                          // random_probability
    long long sx, sy, sxy, sxx, syy;   // exact integer sums over the APs (r_squared, regression)
    int n;                             // anchor points
    int off, cap;                      // pool block
    int mult;                          // multiplicon index, -1 = none
    int mnext;                         // next cluster in the multiplicon's list
    unsigned char orient, was_twisted, dead, has_false;
    // filled by the pack stage (K3)
    int min_y, max_y;                  // Cluster::setBounds (Cluster.h:27-33)
    int n_false;                       // false anchor points added by GHMProfile::setMultiplicons
    int out_off;                       // byte offset of this cluster inside the GHM's packed block
};

struct MultRec {
    int head, tail, count;  // cluster list, join order
    int alive;
    int out_off;            // byte offset of the 33-byte header inside the GHM's packed block
    int pad[3];
};

// Profile column groups for the HL build (GHMProfile.cpp:17-59): per aligned column the
// distinct (family, strand) combinations of its non-gap segment elements; gid = the single
// gene id of the group or -1 when the group holds >= 2 different genes (then every y passes
// the ID != test of Gene.cpp:20).
struct ProfileDev {
    int S, L;
    const int32_t* col_ptr;    // L+1
    const int32_t* grp_fam;
    const int32_t* grp_gid;
    const uint8_t* grp_strand;
    // raw S*L arrays, for anchor-point gene ids (GHMProfile.cpp:86-112)
    const int32_t* fam;
    const int32_t* gid;
    const int32_t* num_id;
    const uint8_t* flags;
    // iadh_profile_search_batch: several profiles in one call.  batch[k] = descriptor of profile k (GhmTask::prof
    // selects it; its own batch / vis fields are unused) and vis[e] = index of the first profile of the batch
    // that sees dataset element e masked (0 = masked before the batch, 0x7fff = never): profile k sees e masked
    // iff vis[e] <= k, which is what the same calls one after the other -- masks applied between them
    // (higherLevel.cpp:165-171) -- would see.  Both null for a single-profile call.
    const ProfileDev* batch;
    const int16_t* vis;
};
// Work item of the family-expansion build of profile GHMs (build_kernels.cu): one column group of one profile.
struct HlGroup {
    int32_t fam, gid;      // the group's family and its single gene id (-1: several genes)
    uint32_t x_strand;     // column | strand << 31
    int32_t prof;          // profile of the batch (0 for a single-profile call)
};
// the profile a GHM belongs to
#ifdef __CUDACC__
__host__ __device__
#endif
static inline const ProfileDev& profile_of(const ProfileDev& pf, int prof) { return pf.batch ? pf.batch[prof] : pf; }

struct DatasetDev {
    int n_lists;
    const int64_t* list_off;
    const int32_t* fam;
    const int32_t* gid;
    const int32_t* num_id;
    uint8_t* flags;             // bit0 strand, bit2 masked, bit3 gene ID occurs more than once in the dataset
    const int32_t* fam_ptr;     // CSR over families: members = global element indices, ascending
    const int32_t* fam_members;
    // level-2 build (build_kernels.cu): per member slot (same order as fam_members)
    const uint32_t* mem_y;      // (coordinate inside its list << 1) | strand
    const int32_t* plus_prefix; // [slot] = members with strand 1 in slots [0, slot)  (n + 1 entries)
    const int32_t* self_slot;   // per element: its slot in fam_members
    const int32_t* mem_list;    // [slot] = list of the member (family-expansion build)
    const int32_t* elem_list;   // per element: its list
    const int32_t* fam_list;    // [f * (n_lists + 1) + l] = first slot of family f at or after list l; may be null
    // table_type=pairs (iadh_create_pairs): "family" = gene id and bucket f holds the positions of the
    // partners of gene f (Gene.cpp:24-27), so buckets overlap and an element is not in its own bucket;
    // fam[] = gid[], self_slot = first slot at or after the element, no gene-ID test anywhere
    int pairs;
};


static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
