// TEST INFRASTRUCTURE -- not part of the product (see oracle/README.md).
//
// CPU restatement ("port") of i-ADHoRe's collinear hot path, written from the
// algorithm description in SURVEY.md Appendix A with every function citing the
// reference file:line it follows.  It is pinned against the real reference by
// tests/test_oracle_vs_reference.py (golden traces produced by oracle/ref_harness.cpp
// running the unmodified reference objects).  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load this library.
#pragma once
#include <cstdint>
#include <set>
#include <string>
#include <utility>
#include <vector>

namespace oracle {

struct Params {
    int n_gaps = 0;
    int gaps[10] = {0};
    int gap_size = 30;
    int cluster_gap = 35;
    int anchor_points = 3;
    int mult_hyp = 1;  // 0 none, 1 Bonferroni, 2 FDR (Settings.h:23-27)
    double q_value = 0.75;
    double prob_cutoff = 0.01;
};

struct List {  // one GeneList after remapTandems(): remapped elements only
    int id = -1;
    std::vector<int32_t> fam, gid, num_id;
    std::vector<uint8_t> strand, masked;
    int size() const { return (int)fam.size(); }
};

struct Profile {  // S aligned segments of equal length L (gaps: flags bit1)
    int S = 0, L = 0, x_object_id = -1;
    std::vector<int32_t> fam, gid, num_id;  // S*L, segment-major
    std::vector<uint8_t> flags;             // bit0 strand, bit1 gap
};

// table_type=pairs: gene id -> ascending partner gene ids, i.e. GenePairs (GenePairs.cpp:3-48) after
// DataSet::getGenePairs dropped the partners that are in no gene list (DataSet.cpp:144-186)
struct PairTable {
    std::vector<int64_t> ptr;
    std::vector<int32_t> ids;
    // Gene::isPairWith, pairs branch (Gene.cpp:24-27): b's ID is in a's set
    bool is_pair(int a, int b) const;
};

struct AP {
    int gx = -1, gy = -1, x = 0, y = 0;
    bool real = true;
};

struct Cluster {
    std::vector<AP> aps;                        // multiset<AnchorPoint> order (AnchorPoint.h:144-156)
    std::vector<std::pair<int, int>> backbone;  // set<AnchorPoint>: unique x, ascending
    bool orient = false;
    double p = 0.0;
    double a = 0.0, b = 0.0, avg_x = 0.0, var_x = 0.0, mrss = 0.0;
    double x_end1 = 0.0, x_end2 = 0.0, y_end1 = 0.0, y_end2 = 0.0;
    bool was_twisted = false;
    int mult = -1;  // index into the GHM's multiplicon vector, -1 = none
    int begin_x = -1, end_x = -1, begin_y = -1, end_y = -1;
};

struct Multiplicon {
    int x_object_id = -1, y_object_id = -1, level = 2;
    int begin_x = -1, end_x = -1, begin_y = -1, end_y = -1;
    std::vector<int> clusters;  // indices into GHMResult::clusters, join order
};

typedef std::set<std::pair<int, int>> DotSet;  // map<int,set<int>>: (x asc, y asc)

struct GHMInput {
    DotSet dots[2];                  // [OPP_ORIENT=0], [SAME_ORIENT=1] (GHM.h:27)
    unsigned long long cnt[2] = {0, 0};
    unsigned long long area = 0;
    int level = 2;
    int x_object_id = -1, y_object_id = -1;
};

// GHM.cpp:34-85 (family mode); with a pair table GHM.cpp:105-147 (table_type=pairs)
void build_l2(const List& X, const List& Y, bool identical, GHMInput& g, const PairTable* pt = nullptr);
// GHMProfile.cpp:17-59 + ListElement.cpp:20-30 + Gene.cpp:17-30
void build_hl(const Profile& P, const List& Y, GHMInput& g, const PairTable* pt = nullptr);
// GHM.cpp:157-219 minus setMultiplicons; output = packed clusters / multiplicons
void run_collinear(GHMInput& g, const Params& prm, std::vector<Cluster>& clusters,
                   std::vector<Multiplicon>& mults);
// GHM.cpp:671-701 (bounds + gene ids); homolog links stay on the host side
void set_multiplicons_l2(const List& X, const List& Y, std::vector<Cluster>& clusters,
                         std::vector<Multiplicon>& mults);
// GHMProfile.cpp:68-136 up to (not including) createHomologs
void set_multiplicons_hl(const Profile& P, const List& Y, std::vector<Cluster>& clusters,
                         std::vector<Multiplicon>& mults, const PairTable* pt = nullptr);
// One Link of a multiplicon's homolog set (Multiplicon.h:22-52): gene numIDs in two segments, segment_x < segment_y
struct Link {
    int seg_x = 0, seg_y = 0, gene_x = -1, gene_y = -1;
};
// Multiplicon::createHomologs(useFamily, minHomologs) (Multiplicon.cpp:352-382) for a multiplicon found by searching
// profile P in list Y: extractXObject / extractYObject (:331-350) slice the S segments to [begin_x, end_x] and Y to
// [begin_y, end_y], pruneLinks (:384-421) keeps the inherited Links whose two ends survive the slicing, addLinks
// (:423-494) adds the x-segment x y-segment homologs; true iff every x-segment and the y-segment hold at least
// min_homologs elements with a homolog.  This is the filter of GHMProfile::setMultiplicons (GHMProfile.cpp:129-132).
// `links` = the homologs of the profile's multiplicon (never empty for a real profile; for an empty set the reference
// leaves the copied hasHomolog flags alone, :386-388 -- here they count as unset).
bool enough_homologs(const Profile& P, const List& Y, const std::vector<Link>& links, const Multiplicon& m, int min_homologs,
                     const PairTable* pt = nullptr);
// the filter applied to a GHM's multiplicons: the survivors, in order (GHMProfile.cpp:129-135)
void filter_homologs(const Profile& P, const List& Y, const std::vector<Link>& links, int min_homologs,
                     std::vector<Multiplicon>& mults, const PairTable* pt = nullptr);
// Multiplicon.cpp:210-240 / BaseCluster.cpp:751-797 / AnchorPoint.h:103-118, appended to `out`
void pack(const std::vector<Cluster>& clusters, const std::vector<Multiplicon>& mults,
          std::vector<uint8_t>& out, int& n_mult_total);

// Settings.cpp:585-598
int gap_sizes(int gap_size, int* gaps);
// hpmath.cpp:18-36
double omxn(double x, int n);

}  // namespace oracle
