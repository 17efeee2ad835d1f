// TEST INFRASTRUCTURE -- not part of the product (see oracle_core.h).
//
// Sequential CPU restatement of the collinear hot path of i-ADHoRe 3.0.  Written
// against SURVEY.md Appendix A; each function names the reference lines it
// follows.  Build with -ffp-contract=off: the reference is plain x86-64 -O3 and
// contains no fused multiply-adds, and every FP expression below keeps the
// reference's evaluation order.
#include "oracle_core.h"

#include <algorithm>
#ifdef ORACLE_STUDY
#include <atomic>
#include <memory>
#include <mutex>
#include <set>
#include <tuple>
#endif
#include <cfloat>
#include <cmath>
#include <cstring>
#include <map>

namespace oracle {

// ---------------------------------------------------------------------------
// Cluster.cpp:5-25
static inline int dpd_i(int x1, int y1, int x2, int y2)
{
    int dx = std::abs(x1 - x2), dy = std::abs(y1 - y2);
    return dx < dy ? 2 * dy - dx : 2 * dx - dy;
}
static inline double dpd_d(double x1, double y1, double x2, double y2)
{
    double dx = std::fabs(x1 - x2), dy = std::fabs(y1 - y2);
    return dx < dy ? 2 * dy - dx : 2 * dx - dy;
}

// Settings.cpp:585-598
int gap_sizes(int gap_size, int* gaps)
{
    int nr = 0;
    double step = (std::log(gap_size) / std::log(3) - 1) / 9;
    for (int i = 0; i < 10; i++) {
        int value = (int)std::round(std::pow(3, (1 + (i * step))));
        if (i == 0 || value != gaps[nr - 1]) gaps[nr++] = value;
    }
    for (int i = nr; i < 10; i++) gaps[i] = 0;
    return nr;
}

// hpmath.cpp:18-36
double omxn(double x, int n)
{
    if (std::abs(x) > 0.1) return std::pow(1.0 - x, n);
    const double epsilon = std::abs(x * DBL_EPSILON);
    double term = 1.0, result = term;
    int i = 0;
    while ((std::abs(term) > epsilon) && i != n) {
        i++;
        term = -term * (n - i + 1.0) / (i * 1.0) * x;
        result += term;
    }
    return result;
}

bool PairTable::is_pair(int a, int b) const
{
    if (a < 0 || a + 1 >= (int)ptr.size()) return false;
    return std::binary_search(ids.begin() + ptr[a], ids.begin() + ptr[a + 1], (int32_t)b);
}

// ---------------------------------------------------------------------------
// GHM.cpp:34-85: family-hash join, family mode.  Masks are ignored (T11).
// table_type=pairs (GHM.cpp:105-147): every x against every y that is neither a gap nor masked
// (ListElement::matchingPositions, ListElement.cpp:20-30) and whose ID is in x's pair set.
void build_l2(const List& X, const List& Y, bool identical, GHMInput& g, const PairTable* pt)
{
    g.dots[0].clear();
    g.dots[1].clear();
    g.cnt[0] = g.cnt[1] = 0;
    if (pt) {
        for (int x = 0; x < X.size(); x++)
            for (int y = 0; y < Y.size(); y++) {
                if (Y.masked[y]) continue;
                if (!pt->is_pair(X.gid[x], Y.gid[y])) continue;
                if (identical && y > x) continue;  // GHM.cpp:125
                int o = (X.strand[x] == Y.strand[y]) ? 1 : 0;
                g.dots[o].insert(std::make_pair(x, y));
                g.cnt[o]++;
            }
    } else {
    std::map<int, std::vector<int>> tree;  // family -> y positions, ascending
    for (int y = 0; y < Y.size(); y++) tree[Y.fam[y]].push_back(y);
    for (int x = 0; x < X.size(); x++) {
        std::map<int, std::vector<int>>::const_iterator it = tree.find(X.fam[x]);
        if (it == tree.end()) continue;
        for (int y : it->second) {
            if (identical && y > x) continue;      // GHM.cpp:64
            if (X.gid[x] == Y.gid[y]) continue;    // GHM.cpp:68 (ID strings, T15)
            int o = (X.strand[x] == Y.strand[y]) ? 1 : 0;
            g.dots[o].insert(std::make_pair(x, y));
            g.cnt[o]++;
        }
    }
    }
    // GHM.cpp:221-236, 32-bit unsigned arithmetic (T8); level 2 => no masked rows
    unsigned int nx = (unsigned)X.size(), ny = (unsigned)Y.size();
    g.area = identical ? (unsigned int)(nx * (nx - 1) / 2) : (unsigned int)(nx * ny);
    g.level = 2;
    g.x_object_id = X.id;
    g.y_object_id = Y.id;
}

// GHMProfile.cpp:17-59: every segment, every non-gap position against every
// non-gap, non-masked y element (ListElement.cpp:20-30), same family and a
// different gene ID (Gene.cpp:17-30); counted on first insert only.
void build_hl(const Profile& P, const List& Y, GHMInput& g, const PairTable* pt)
{
    g.dots[0].clear();
    g.dots[1].clear();
    g.cnt[0] = g.cnt[1] = 0;
    for (int s = 0; s < P.S; s++)
        for (int x = 0; x < P.L; x++) {
            int k = s * P.L + x;
            if (P.flags[k] & 2) continue;
            for (int y = 0; y < Y.size(); y++) {
                if (Y.masked[y]) continue;
                if (pt ? !pt->is_pair(P.gid[k], Y.gid[y]) : (P.fam[k] != Y.fam[y] || P.gid[k] == Y.gid[y])) continue;
                int o = ((P.flags[k] & 1) == Y.strand[y]) ? 1 : 0;
                if (g.dots[o].insert(std::make_pair(x, y)).second) g.cnt[o]++;
            }
        }
    // GHM.cpp:221-236 with level > 2: masked rows are subtracted; |X| = L
    unsigned int masked = 0;
    for (int y = 0; y < Y.size(); y++) masked += Y.masked[y] ? 1 : 0;
    unsigned int nx = (unsigned)P.L, ny = (unsigned)Y.size();
    g.area = (unsigned int)(nx * (ny - masked));
    g.level = P.S + 1;               // GHMProfile.cpp:14
    g.x_object_id = P.x_object_id;   // GHM base x_object = segment 0 (GHMProfile.cpp:13)
    g.y_object_id = Y.id;
}

// ---------------------------------------------------------------------------
// BaseCluster arithmetic

// multiset<AnchorPoint>::insert: after the last element with x <= X
static void ap_insert(Cluster& c, const AP& p)
{
    std::vector<AP>::iterator it = c.aps.begin();
    // upper_bound on x
    it = std::upper_bound(c.aps.begin(), c.aps.end(), p.x,
                          [](int x, const AP& q) { return x < q.x; });
    c.aps.insert(it, p);
    c.b = 0.0;  // BaseCluster.cpp:60,84
}
static void add_ap(Cluster& c, int x, int y)
{
    AP p;
    p.x = x;
    p.y = y;
    ap_insert(c, p);
}
// set<AnchorPoint>::insert on x: a second point with the same x is dropped
static void add_backbone(Cluster& c, int x, int y)
{
    std::vector<std::pair<int, int>>::iterator it = std::lower_bound(
        c.backbone.begin(), c.backbone.end(), x,
        [](const std::pair<int, int>& q, int xx) { return q.first < xx; });
    if (it != c.backbone.end() && it->first == x) return;
    c.backbone.insert(it, std::make_pair(x, y));
}

// BaseCluster.cpp:146-180 / 182-221.  Sums of int products (T7) held in doubles.
static double r2_from(const Cluster& c, const Cluster* other, int X, int Y)
{
    double sum_x = 0, sum_y = 0, sum_xy = 0, sum_x2 = 0, sum_y2 = 0;
    unsigned int n = 0;
    for (const AP& e : c.aps) {
        sum_x += e.x; sum_y += e.y; sum_xy += e.x * e.y; sum_x2 += e.x * e.x; sum_y2 += e.y * e.y;
        n++;
    }
    if (other) {
        for (const AP& e : other->aps) {
            sum_x += e.x; sum_y += e.y; sum_xy += e.x * e.y; sum_x2 += e.x * e.x; sum_y2 += e.y * e.y;
            n++;
        }
    } else if (X != 0 || Y != 0) {  // T6: (0,0) is silently ignored
        sum_x += X; sum_y += Y; sum_xy += X * Y; sum_x2 += X * X; sum_y2 += Y * Y;
        n++;
    }
    double value = ((sum_x2 - sum_x * sum_x / n) * (sum_y2 - sum_y * sum_y / n));
    if (value == 0) return -1;
    double r = (sum_xy - (sum_x * sum_y / n)) / std::sqrt(value);
    return r * r;
}

// BaseCluster.cpp:266-291
static void regression(Cluster& c)
{
    double N = (double)c.aps.size();
    if (N < 3.0) return;
    double sum_x = 0.0, sum_y = 0.0, sum_xy = 0.0, sum_x2 = 0.0;
    for (const AP& e : c.aps) {
        sum_x += e.x; sum_y += e.y; sum_xy += e.x * e.y; sum_x2 += e.x * e.x;
    }
    c.avg_x = sum_x / N;
    double avg_y = sum_y / N;
    c.var_x = (sum_x2 - N * c.avg_x * c.avg_x) / (N - 1.0);
    double cov_xy = (sum_xy - N * c.avg_x * avg_y) / (N - 1.0);
    c.b = cov_xy / c.var_x;
    c.a = avg_y - c.b * c.avg_x;
}

// BaseCluster.cpp:238-264
static void update_stats(Cluster& c)
{
    if (c.b != 0.0) return;
    int N = (int)c.aps.size();
    regression(c);
    c.var_x = 0.0;
    c.mrss = 0.0;
    for (const AP& e : c.aps) {
        double y_est = c.b * e.x + c.a;
        c.mrss += (e.y - y_est) * (e.y - y_est);
        c.var_x += (e.x - c.avg_x) * (e.x - c.avg_x);
    }
    c.mrss = c.mrss / (N - 2);
    c.var_x = c.var_x / (N - 1);
    c.x_end1 = (unsigned int)c.aps.front().x;
    c.x_end2 = (unsigned int)c.aps.back().x;
    c.y_end1 = c.x_end1 * c.b + c.a;
    c.y_end2 = c.x_end2 * c.b + c.a;
}

// BaseCluster.cpp:293-304
static double dist_to_point(const Cluster& c, int x, int y)
{
    if (x >= c.x_end1 && x <= c.x_end2) return 0;
    if ((double)x < c.x_end1) return dpd_d((double)x, (double)y, c.x_end1, c.y_end1);
    return dpd_d((double)x, (double)y, c.x_end2, c.y_end2);
}

// BaseCluster.cpp:306-336
static bool in_interval(Cluster& c, int x, int y)
{
    update_stats(c);
    double n = (double)c.aps.size();
    double pi = 1.96 * std::sqrt(c.mrss) *
                std::sqrt(1.0 + 1.0 / n + (x - c.avg_x) * (x - c.avg_x) / (n - 1) / c.var_x);
    double up = (c.a + c.b * x) + pi;
    double down = (c.a + c.b * x) - pi;
    return y >= down && y <= up;
}

// BaseCluster.cpp:353-364
static bool intersect(double x1, double y1, double x2, double y2, double x3, double y3,
                      double x4, double y4)
{
    double ud = (y4 - y3) * (x2 - x1) - (x4 - x3) * (y2 - y1);
    if (ud != 0) {
        double ua = ((x4 - x3) * (y1 - y3) - (y4 - y3) * (x1 - x3)) / ud;
        double ub = ((x2 - x1) * (y1 - y3) - (y2 - y1) * (x1 - x3)) / ud;
        if (ua >= 0 && ua <= 1 && ub >= 0 && ub <= 1) return true;
    }
    return false;
}
static inline double mind(double a, double b) { return (a < b) ? a : b; }
static inline double maxd(double a, double b) { return (a > b) ? a : b; }

// BaseCluster.cpp:376-396
static double point_line_dpd(double x1, double y1, double x2, double y2, double x3, double y3)
{
    double dx = x2 - x1, dy = y2 - y1, x0, y0;
    if ((dx == 0) && (dy == 0)) {
        x0 = x1;
        y0 = y1;
    } else {
        double t = ((x3 - x1) * dx + (y3 - y1) * dy) / (dx * dx + dy * dy);
        t = mind(maxd(0, t), 1);
        x0 = x1 + t * dx;
        y0 = y1 + t * dy;
    }
    return dpd_d(x0, y0, x3, y3);
}

// BaseCluster.cpp:398-465
static double dist_to_cluster(Cluster& A, Cluster& B)
{
    update_stats(A);
    update_stats(B);
    if (intersect(A.x_end1, A.y_end1, A.x_end2, A.y_end2, B.x_end1, B.y_end1, B.x_end2, B.y_end2))
        return 0;
    double lo = point_line_dpd(A.x_end1, A.y_end1, A.x_end2, A.y_end2, B.x_end1, B.y_end1);
    lo = mind(lo, point_line_dpd(A.x_end1, A.y_end1, A.x_end2, A.y_end2, B.x_end2, B.y_end2));
    lo = mind(lo, point_line_dpd(B.x_end1, B.y_end1, B.x_end2, B.y_end2, A.x_end1, A.y_end1));
    lo = mind(lo, point_line_dpd(B.x_end1, B.y_end1, B.x_end2, B.y_end2, A.x_end2, A.y_end2));
    const AP &s1 = A.aps.front(), &e1 = A.aps.back(), &s2 = B.aps.front(), &e2 = B.aps.back();
    lo = mind(lo, (double)dpd_i(s1.x, s1.y, s2.x, s2.y));
    lo = mind(lo, (double)dpd_i(s1.x, s1.y, e2.x, e2.y));
    lo = mind(lo, (double)dpd_i(e1.x, e1.y, s2.x, s2.y));
    lo = mind(lo, (double)dpd_i(e1.x, e1.y, e2.x, e2.y));
    lo = mind(lo, point_line_dpd(A.x_end1, A.y_end1, A.x_end2, A.y_end2, s2.x, s2.y));
    lo = mind(lo, point_line_dpd(A.x_end1, A.y_end1, A.x_end2, A.y_end2, e2.x, e2.y));
    lo = mind(lo, point_line_dpd(B.x_end1, B.y_end1, B.x_end2, B.y_end2, s1.x, s1.y));
    lo = mind(lo, point_line_dpd(B.x_end1, B.y_end1, B.x_end2, B.y_end2, e1.x, e1.y));
    return lo;
}

// BaseCluster.cpp:491-546
static bool partial_overlap(Cluster& A, Cluster& B)
{
    update_stats(A);
    update_stats(B);
    int x1min = A.aps.front().x, x1max = A.aps.back().x;
    for (const AP& e : B.aps)
        if (e.x >= x1min && e.x <= x1max)
            if (!in_interval(A, e.x, e.y)) return false;
    int x2min = B.aps.front().x, x2max = B.aps.back().x;
    for (const AP& e : A.aps)
        if (e.x >= x2min && e.x <= x2max)
            if (!in_interval(B, e.x, e.y)) return false;
    return true;
}

// BaseCluster.cpp:338-351
static void merge_with(Cluster& A, const Cluster& B)
{
    for (const AP& e : B.aps) add_ap(A, e.x, e.y);
    for (const std::pair<int, int>& q : B.backbone) add_backbone(A, q.first, q.second);
    A.b = 0.0;
    update_stats(A);
}

// BaseCluster.cpp:548-590 (including the never-advancing it2 quirk, T6)
static double calc_probability(Cluster& c, double area, double points)
{
    double density = (double)points / (double)area;
    double probability = 1.0;
    for (size_t i = 1; i < c.backbone.size(); i++) {
        double distance = dpd_d(c.backbone[i - 1].first, c.backbone[i - 1].second,
                                c.backbone[i].first, c.backbone[i].second);
        int cc = (int)std::ceil(0.5 * distance * distance);
        double p_i = 1.0 - std::pow(1.0 - density, cc);
        probability *= p_i;
    }
    // duplicates of the FIRST anchor point are erased; b is not touched
    if (!c.aps.empty()) {
        AP first = c.aps.front();
        std::vector<AP> keep;
        keep.push_back(first);
        for (size_t i = 1; i < c.aps.size(); i++)
            if (!(c.aps[i].x == first.x && c.aps[i].y == first.y)) keep.push_back(c.aps[i]);
        c.aps.swap(keep);
    }
    return probability;
}

// BaseCluster.cpp:700-719
static void twist(Cluster& c)
{
    int max_y = c.aps.front().y, min_y = c.aps.front().y;
    for (const AP& e : c.aps) {
        max_y = std::max(max_y, e.y);
        min_y = std::min(min_y, e.y);
    }
    for (AP& e : c.aps) e.y = max_y + min_y - e.y;  // x-keyed order is unchanged
    c.b = 0.0;
    update_stats(c);
}

static unsigned real_count(const Cluster& c)
{
    unsigned n = 0;
    for (const AP& e : c.aps) n += e.real ? 1 : 0;
    return n;
}

// ---------------------------------------------------------------------------
// GHM pipeline state

struct GHMState {
    GHMInput& g;
    const Params& prm;
    // clusters live in a pool; BC[o] holds pool indices in vector order
    std::vector<Cluster>& pool;
    std::vector<int> BC[2];
    std::vector<Multiplicon>& mults;
    GHMState(GHMInput& g_, const Params& p_, std::vector<Cluster>& pool_, std::vector<Multiplicon>& m_)
        : g(g_), prm(p_), pool(pool_), mults(m_) {}
};

#ifdef ORACLE_STUDY
#include "study/join_cache.inc"
#endif

// GHM.cpp:238-285
static void seed_chain(const DotSet& mat, Cluster& c, int gap)
{
    bool orientation = c.orient;
    int refX = c.aps.front().x, refY = c.aps.front().y;
    while (true) {
        int loX = refX + 1, hiX = refX + gap;
        int loY = orientation ? refY + 1 : refY - gap;
        int hiY = orientation ? refY + gap : refY - 1;
        int closestX = 0, closestY = 0, closestDpd = gap + 1;
        DotSet::const_iterator it = mat.lower_bound(std::make_pair(loX, INT32_MIN));
        for (; it != mat.end() && it->first <= hiX; ++it) {
            int x = it->first, y = it->second;
            if (y < loY || y > hiY) continue;
            int d = dpd_i(refX, refY, x, y);
            if (d < closestDpd) {
                closestDpd = d;
                closestX = x;
                closestY = y;
            }
        }
        if (closestDpd == gap + 1) return;
        add_ap(c, closestX, closestY);
        add_backbone(c, closestX, closestY);
        refX = closestX;
        refY = closestY;
    }
}

// GHM.cpp:287-324
static void seed_base_clusters(GHMState& s, int gap, int orientation)
{
    DotSet& mat = s.g.dots[orientation];
    DotSet::iterator it = mat.begin();
    while (it != mat.end()) {
        Cluster c;
        c.orient = orientation != 0;
        add_ap(c, it->first, it->second);
        add_backbone(c, it->first, it->second);
        seed_chain(mat, c, gap);
        if (c.aps.size() > 2 && r2_from(c, NULL, 0, 0) >= s.prm.q_value) {
            update_stats(c);
            std::pair<int, int> cur = *it;
            for (const AP& e : c.aps) mat.erase(std::make_pair(e.x, e.y));
            it = mat.upper_bound(cur);
            s.pool.push_back(c);
            s.BC[orientation].push_back((int)s.pool.size() - 1);
        } else {
            ++it;
        }
    }
}

// GHM.cpp:331-390
static void enrich_clusters(GHMState& s, int gap, int clusterOrientation, int ghmOrientation)
{
    std::vector<int> clusters(s.BC[clusterOrientation]);
    DotSet& mat = s.g.dots[ghmOrientation];
    const double q = s.prm.q_value;
    while (!clusters.empty()) {
        std::vector<int> next;
        std::vector<bool> changed(clusters.size(), false);
        DotSet::iterator it = mat.begin();
        while (it != mat.end()) {
            int x = it->first, y = it->second;
            double closestDistance = gap + 1;
            int closest = -1;
#ifdef ORACLE_STUDY
            if (study::cur) study::cur->enrich_tests += (long long)clusters.size();
#endif
            for (size_t k = 0; k < clusters.size(); k++) {
                Cluster& c = s.pool[clusters[k]];
                update_stats(c);
                double distance = dist_to_point(c, x, y);
                if (distance >= closestDistance) continue;
                if (r2_from(c, NULL, x, y) >= q && in_interval(c, x, y)) {
                    closestDistance = distance;
                    closest = (int)k;
                }
            }
            if (closest != -1) {
                Cluster& c = s.pool[clusters[closest]];
                add_ap(c, x, y);
                update_stats(c);
                if (!changed[closest]) {
                    next.push_back(clusters[closest]);
                    changed[closest] = true;
                }
                it = mat.erase(it);
            } else {
                ++it;
            }
        }
        clusters.swap(next);
    }
}

// GHM.cpp:406-444; returns positions in BC[orientI] / BC[orientJ]
static bool closest_clusters(GHMState& s, int gap, int orientI, int orientJ, int& ci, int& cj)
{
    std::vector<int>& bcI = s.BC[orientI];
    std::vector<int>& bcJ = s.BC[orientJ];
    double closestDistance = gap + 1;
    for (size_t i = 0; i < bcI.size(); i++) {
        size_t j = (orientI == orientJ) ? i + 1 : 0;
        for (; j < bcJ.size(); j++) {
            Cluster& cI = s.pool[bcI[i]];
            Cluster& cJ = s.pool[bcJ[j]];
            if (cI.mult >= 0 && cJ.mult >= 0 && cI.mult == cJ.mult) continue;
            double distance = dist_to_cluster(cI, cJ);
            if (distance < closestDistance) {
                if (partial_overlap(cI, cJ)) {
                    if (r2_from(cI, &cJ, 0, 0) >= s.prm.q_value) {
                        ci = (int)i;
                        cj = (int)j;
                        closestDistance = distance;
                    }
                }
            }
        }
    }
    return closestDistance < gap + 1;
}

#ifdef ORACLE_STUDY
static void study_join_init(GHMState& s, int gap, int orient)
{
    using namespace study;
    if (!cur) return;
    const std::vector<int>& bc = s.BC[orient];
    const int maxgap = s.prm.gaps[s.prm.n_gaps - 1];
    std::vector<Hull> h;
    for (int id : bc) h.push_back(hull(s.pool[id]));
    C.calls++;
    C.clusters += (long long)bc.size();
    C.h_calls[bucket(bc.size())]++;
    C.h_pairs[bucket(bc.size())] += (long long)bc.size() * ((long long)bc.size() - 1) / 2;
    long long mx = C.max_clusters.load();
    while ((long long)bc.size() > mx && !C.max_clusters.compare_exchange_weak(mx, (long long)bc.size())) {}
    for (int id : bc)
        if (cur->seen_clusters.insert(std::make_pair(id, (int)s.pool[id].aps.size())).second) C.changed++;
    cur->max_clusters = std::max(cur->max_clusters, (long long)bc.size());
    // warp-level view of the scan (32 lanes): today one warp takes a row and walks it in chunks of 32
    // later positions, and a chunk runs the expensive evaluation if ANY of its pairs is near; a two-phase
    // scan would first compact the near pairs of the whole call and evaluate them 32 at a time
    {
        long long near_call = 0;
        for (size_t i = 0; i + 1 < bc.size(); i++)
            for (size_t j0 = i + 1; j0 < bc.size(); j0 += 32) {
                bool any = false;
                for (size_t j = j0; j < std::min(bc.size(), j0 + 32); j++)
                    if (apart(h[i], h[j]) <= gap + 1) { any = true; near_call++; }
                C.warp_chunks++;
                if (any) C.warp_heavy_now++;
            }
        C.warp_heavy_2phase += (near_call + 31) / 32;
        C.warp_scan_2phase += ((long long)bc.size() * ((long long)bc.size() - 1) / 2 + 31) / 32;
    }
    for (size_t i = 0; i < bc.size(); i++)
        for (size_t j = i + 1; j < bc.size(); j++) {
            C.pairs++;
            cur->pairs++;
            const int d = apart(h[i], h[j]);
            if (d <= gap + 1) { C.near_gap++; cur->near++; }
            if (d <= maxgap + 1) {
                C.near_max++;
                if (cur->seen_pairs.insert(std::make_tuple(bc[i], (int)s.pool[bc[i]].aps.size(), bc[j], (int)s.pool[bc[j]].aps.size())).second)
                    C.first_max++;
            }
        }
}
#endif

#ifdef ORACLE_STUDY
// Lane-level CPU model of KT<32>::join_pass2 (i-adhore_b200/csrc/cluster_kernel.cu): the same flattened
// scan, 64-entry queue, 32-wide evaluation with a segmented minimum, classification after a merge and redo
// rows, with 32 "lanes" stepped in lock-step.  It proposes a merge sequence; join_clusters compares every
// proposal with the reference's pick (study::C.j2_merges / j2_mismatch).
struct J2Model {
    GHMState& s;
    int gap, o;
    std::vector<int>& bc;
    std::vector<double> rb_d;   // by pool id
    std::vector<int> rb_j, rank, pos;
    struct Box { int xlo, xhi, ylo, yhi; };
    std::vector<Box> box;       // by position
    std::vector<std::pair<int, int>> q;
    J2Model(GHMState& s_, int gap_, int o_) : s(s_), gap(gap_), o(o_), bc(s_.BC[o_]) {}
    Box make_box(Cluster& c)
    {
        update_stats(c);
        const int ys = c.aps.front().y, ye = c.aps.back().y;
        const double ylo = mind(mind(c.y_end1, c.y_end2), (double)std::min(ys, ye));
        const double yhi = maxd(maxd(c.y_end1, c.y_end2), (double)std::max(ys, ye));
        const double lim = 1e9;
        Box b;
        b.xlo = (c.x_end1 > -lim && c.x_end1 < lim) ? (int)std::floor(c.x_end1) : -(1 << 30);
        b.xhi = (c.x_end2 > -lim && c.x_end2 < lim) ? (int)std::ceil(c.x_end2) : (1 << 30);
        b.ylo = (ylo > -lim && ylo < lim) ? (int)std::floor(ylo) : -(1 << 30);
        b.yhi = (yhi > -lim && yhi < lim) ? (int)std::ceil(yhi) : (1 << 30);
        return b;
    }
    bool apart(int a, int b) const
    {
        const int cap = gap + 1;
        return (box[b].xlo - box[a].xhi > cap) || (box[a].xlo - box[b].xhi > cap) || (box[b].ylo - box[a].yhi > cap) ||
               (box[a].ylo - box[b].yhi > cap);
    }
    void drain(int cnt)
    {
        int row[32], idJ[32], rk[32];
        double d[32];
        for (int l = 0; l < 32; l++) {
            row[l] = -1 - l; idJ[l] = -1; rk[l] = 0x7fffffff; d[l] = 1e300;
            if (l < cnt) {
                row[l] = q[l].first;
                const int id = bc[q[l].second];
                Cluster& cI = s.pool[bc[row[l]]];
                Cluster& cJ = s.pool[id];
                const double dd = dist_to_cluster(cI, cJ);
                if (dd < gap + 1 && r2_from(cI, &cJ, 0, 0) >= s.prm.q_value && partial_overlap(cI, cJ)) {
                    d[l] = dd; idJ[l] = id; rk[l] = rank[id];
                }
            }
        }
        for (int off = 1; off < 32; off <<= 1) {
            int row2[32], id2[32], rk2[32];
            double d2[32];
            for (int l = 0; l < 32; l++) { const int m = (l + off < 32) ? l + off : l; row2[l] = row[m]; id2[l] = idJ[m]; rk2[l] = rk[m]; d2[l] = d[m]; }
            for (int l = 0; l < 32; l++)
                if (l + off < 32 && row2[l] == row[l] && id2[l] >= 0 && (idJ[l] < 0 || d2[l] < d[l] || (d2[l] == d[l] && rk2[l] < rk[l]))) {
                    d[l] = d2[l]; idJ[l] = id2[l]; rk[l] = rk2[l];
                }
        }
        for (int l = 0; l < cnt; l++) {
            const bool head = (l == 0 || row[l - 1] != row[l]);
            if (!head || idJ[l] < 0) continue;
            const int idI = bc[row[l]];
            const int cj = rb_j[idI];
            if (cj < 0 || d[l] < rb_d[idI] || (d[l] == rb_d[idI] && rk[l] < rank[cj])) { rb_d[idI] = d[l]; rb_j[idI] = idJ[l]; }
        }
        q.erase(q.begin(), q.begin() + cnt);
    }
    // one warp step: lanes in order, survivors appended, 32 evaluated as soon as 32 are queued
    void push(const bool* pass, const int* row, const int* j)
    {
        for (int l = 0; l < 32; l++)
            if (pass[l]) q.push_back(std::make_pair(row[l], j[l]));
        if (q.size() >= 32) drain(32);
    }
    void flush() { if (!q.empty()) drain((int)q.size()); }
    void start()
    {
        const int n = (int)bc.size();
        rb_d.assign(s.pool.size(), gap + 1); rb_j.assign(s.pool.size(), -1);
        rank.assign(s.pool.size(), -1); pos.assign(s.pool.size(), -1);
        box.resize(n);
        for (int i = 0; i < n; i++) { rank[bc[i]] = i; pos[bc[i]] = i; box[i] = make_box(s.pool[bc[i]]); }
        const long long total = (long long)n * (n - 1) / 2;
        int row[32], j[32];
        for (int l = 0; l < 32; l++) {
            row[l] = 0; j[l] = 1 + l;
            while (row[l] < n - 1 && j[l] >= n) { j[l] = j[l] - n + row[l] + 2; row[l]++; }
        }
        for (long long base = 0; base < total; base += 32) {
            bool pass[32];
            for (int l = 0; l < 32; l++) pass[l] = (row[l] < n - 1) && !apart(row[l], j[l]);
            push(pass, row, j);
            for (int l = 0; l < 32; l++) {
                j[l] += 32;
                while (row[l] < n - 1 && j[l] >= n) { j[l] = j[l] - n + row[l] + 2; row[l]++; }
            }
        }
        flush();
    }
    bool pick(int& pA, int& pB) const
    {
        const int n = (int)bc.size();
        double bd = gap + 1;
        long long bk = -1;
        for (int i = 0; i < n; i++) {
            const int id = bc[i];
            if (rb_j[id] < 0) continue;
            const long long k = ((long long)i << 32) | (unsigned)pos[rb_j[id]];
            if (bk < 0 || rb_d[id] < bd || (rb_d[id] == bd && k < bk)) { bd = rb_d[id]; bk = k; }
        }
        if (bk < 0) return false;
        pA = (int)(bk >> 32); pB = (int)(bk & 0xffffffff);
        return true;
    }
    // called after the caller merged position pB into pA and erased pB from bc
    void merged(int pA, int idA, int idB)
    {
        const int n = (int)bc.size();
        box.erase(box.begin() + pos[idB]);
        for (int i = 0; i < n; i++) pos[bc[i]] = i;
        box[pA] = make_box(s.pool[idA]);
        std::vector<int> redo;
        for (int base = 0; base < n; base += 32) {
            bool pass[32]; int row[32], j[32];
            for (int l = 0; l < 32; l++) {
                const int i = base + l;
                pass[l] = false; row[l] = i; j[l] = pA;
                if (i >= n) continue;
                const int id = bc[i];
                const int jj = rb_j[id];
                if (id == idA || jj == idA || jj == idB) {
                    rb_d[id] = gap + 1; rb_j[id] = -1;
                    if (i < n - 1) redo.push_back(i);
                } else if (i < pA)
                    pass[l] = !apart(i, pA);
            }
            push(pass, row, j);
        }
        for (int r : redo)
            for (int base = r + 1; base < n; base += 32) {
                bool pass[32]; int row[32], j[32];
                for (int l = 0; l < 32; l++) { row[l] = r; j[l] = base + l; pass[l] = (j[l] < n) && !apart(r, j[l]); }
                push(pass, row, j);
            }
        flush();
    }
};
#endif

// GHM.cpp:392-404
static void join_clusters(GHMState& s, int gap, int orient)
{
#ifdef ORACLE_STUDY
    study_join_init(s, gap, orient);
#endif
    int i, j;
#ifdef ORACLE_STUDY
    if (study::j2_model && s.BC[orient].size() >= 2) {
        std::unique_ptr<J2Model> mp(new J2Model(s, gap, orient));
        mp->start();
        while (true) {
            J2Model& m = *mp;
            int pA = -1, pB = -1;
            const bool mine = m.pick(pA, pB);
            const bool ref = closest_clusters(s, gap, orient, orient, i, j);
            if (mine != ref || (ref && (pA != i || pB != j))) {
                study::C.j2_mismatch++;
                if (!ref) break;
            }
            if (!ref) break;
            study::C.j2_merges++;
            study::C.merges++;
            if (study::cur) study::cur->merges++;
            const int idA = s.BC[orient][i], idB = s.BC[orient][j];
            merge_with(s.pool[idA], s.pool[idB]);
            s.BC[orient].erase(s.BC[orient].begin() + j);
            if (mine && pA == i && pB == j) m.merged(i, idA, idB);
            else { mp.reset(new J2Model(s, gap, orient)); mp->start(); }   // resynchronise after a mismatch
        }
        return;
    }
    // Speculative batches: at a batch start list every valid pair, take up to K disjoint ones in the
    // reference's order, then follow the reference's own sequence and count how long it agrees with the
    // speculation (= the merges a device could have carried out concurrently and validated afterwards).
    // The merges performed are always the reference's.
    if (study::spec_k > 0) {
        while (true) {
            std::vector<int>& bc = s.BC[orient];
            std::vector<std::tuple<double, int, int>> cand;
            for (size_t a = 0; a < bc.size(); a++)
                for (size_t b = a + 1; b < bc.size(); b++) {
                    Cluster& cI = s.pool[bc[a]];
                    Cluster& cJ = s.pool[bc[b]];
                    const double d = dist_to_cluster(cI, cJ);
                    if (d < gap + 1 && partial_overlap(cI, cJ) && r2_from(cI, &cJ, 0, 0) >= s.prm.q_value)
                        cand.push_back(std::make_tuple(d, (int)a, (int)b));
                }
            if (cand.empty()) break;
            std::sort(cand.begin(), cand.end());
            std::vector<std::pair<int, int>> spec;
            std::set<int> used;
            for (const auto& c : cand) {
                const int ia = bc[std::get<1>(c)], ib = bc[std::get<2>(c)];
                if (used.count(ia) || used.count(ib)) continue;
                used.insert(ia); used.insert(ib);
                spec.push_back(std::make_pair(ia, ib));
                if ((int)spec.size() >= study::spec_k) break;
            }
            // Validation criterion a device would use (no access to the reference's sequence): spec[t] is
            // the smallest pair among the clusters no earlier merge of the batch touched, so the reference
            // picks it at step t unless a pair involving one of the merged clusters A'_0..A'_{t-1} comes
            // before it in (distance, position i, position j) order.  Computed on copies; compared below
            // with what the reference really does.
            int crit = (int)spec.size();
            {
                std::vector<int> rank(s.pool.size(), -1);
                for (size_t a = 0; a < bc.size(); a++) rank[bc[a]] = (int)a;
                std::vector<Cluster> merged;
                std::vector<std::tuple<double, int, int>> key;  // (d, rank i, rank j) of every speculated pair
                for (size_t t = 0; t < spec.size(); t++) {
                    Cluster c = s.pool[spec[t].first];
                    merge_with(c, s.pool[spec[t].second]);
                    merged.push_back(c);
                    Cluster& cI = s.pool[spec[t].first];
                    Cluster& cJ = s.pool[spec[t].second];
                    key.push_back(std::make_tuple(dist_to_cluster(cI, cJ), rank[spec[t].first], rank[spec[t].second]));
                }
                for (size_t t = 1; t < spec.size() && crit == (int)spec.size(); t++) {
                    // the clusters present at step t: untouched ones and the members of spec[t..] as they
                    // are, spec[0..t-1] in merged form at the position of their first member
                    std::vector<std::pair<int, Cluster*>> cur_set;  // (rank, cluster)
                    std::set<int> gone;
                    for (size_t r = 0; r < t; r++) { gone.insert(spec[r].first); gone.insert(spec[r].second); }
                    for (int id : bc) if (!gone.count(id)) cur_set.push_back(std::make_pair(rank[id], &s.pool[id]));
                    for (size_t r = 0; r < t; r++) cur_set.push_back(std::make_pair(rank[spec[r].first], &merged[r]));
                    for (size_t r = 0; r < t && crit == (int)spec.size(); r++)
                        for (const auto& other : cur_set) {
                            if (other.second == &merged[r]) continue;
                            const bool first = rank[spec[r].first] < other.first;
                            Cluster& cI = first ? merged[r] : *other.second;
                            Cluster& cJ = first ? *other.second : merged[r];
                            const double d = dist_to_cluster(cI, cJ);
                            if (!(d < gap + 1)) continue;
                            const std::tuple<double, int, int> k2 = std::make_tuple(d, std::min(rank[spec[r].first], other.first),
                                                                                   std::max(rank[spec[r].first], other.first));
                            if (!(k2 < key[t])) continue;
                            if (partial_overlap(cI, cJ) && r2_from(cI, &cJ, 0, 0) >= s.prm.q_value) { crit = (int)t; break; }
                        }
                }
            }
            int accepted = 0;
            for (size_t t = 0; t < spec.size(); t++) {
                if (!closest_clusters(s, gap, orient, orient, i, j)) break;
                const bool same = (s.BC[orient][i] == spec[t].first && s.BC[orient][j] == spec[t].second);
                study::C.merges++;
                if (study::cur) study::cur->merges++;
                merge_with(s.pool[s.BC[orient][i]], s.pool[s.BC[orient][j]]);
                s.BC[orient].erase(s.BC[orient].begin() + j);
                if (!same) break;
                accepted++;
            }
            study::C.spec_batches++;
            study::C.spec_offered += (long long)spec.size();
            study::C.spec_accepted += accepted;
            if (crit > accepted) study::C.crit_unsafe++;       // the criterion would have accepted a wrong merge
            else if (crit < accepted) study::C.crit_lost += accepted - crit;   // safe, but accepted fewer
            if (study::cur) study::cur->batches++;
        }
        return;
    }
#endif
    while (closest_clusters(s, gap, orient, orient, i, j)) {
#ifdef ORACLE_STUDY
        study::C.merges++;
        if (study::cur) study::cur->merges++;
#endif
        merge_with(s.pool[s.BC[orient][i]], s.pool[s.BC[orient][j]]);
        s.BC[orient].erase(s.BC[orient].begin() + j);
    }
}

// GHM.cpp:446-469 (none), 481-503 (Bonferroni), 516-540 (FDR)
static void filter_clusters(GHMState& s)
{
    const double cutoff = s.prm.prob_cutoff;
    for (int orient = 0; orient < 2; orient++) {
        std::vector<int>& bc = s.BC[orient];
        if (s.prm.mult_hyp == 2) {
            for (int idx : bc) {
                Cluster& c = s.pool[idx];
                c.p = calc_probability(c, (double)s.g.area, (double)s.g.cnt[orient]);
            }
            std::vector<Cluster>& pool = s.pool;
            std::sort(bc.begin(), bc.end(), [&pool](int l, int r) { return pool[l].p < pool[r].p; });
            size_t k;
            for (k = 0; k < bc.size(); k++)
                if (s.pool[bc[k]].p > (int)(k + 1) * cutoff / s.g.cnt[orient]) break;
            bc.erase(bc.begin() + k, bc.end());  // NOT re-inserted into the matrix (T6)
        } else {
            size_t i = 0;
            while (i < bc.size()) {
                Cluster& c = s.pool[bc[i]];
                c.p = calc_probability(c, (double)s.g.area, (double)s.g.cnt[orient]);
                double thr = (s.prm.mult_hyp == 1) ? (cutoff / s.g.cnt[orient]) : cutoff;
                if (c.p > thr) {
                    for (const AP& e : c.aps) s.g.dots[orient].insert(std::make_pair(e.x, e.y));
                    bc.erase(bc.begin() + i);
                } else {
                    i++;
                }
            }
        }
    }
}

// GHM.cpp:542-625
static void cluster_clusters(GHMState& s)
{
    const int clusterGap = s.prm.cluster_gap;
    bool join_flag = true;
    while (join_flag) {
        join_flag = false;
        for (int k = 0; k < 4; k++) {
            int or_x = (k < 2) ? 1 : 0;
            int or_y = (k % 2 == 0) ? 1 : 0;
            bool was_twisted = (or_x != or_y);
            if (was_twisted)
                for (int idx : s.BC[or_y]) twist(s.pool[idx]);
            int pi = -1, pj = -1;
            bool found = closest_clusters(s, clusterGap, or_x, or_y, pi, pj);
            if (was_twisted)
                for (int idx : s.BC[or_y]) twist(s.pool[idx]);
            if (found && (!was_twisted || !s.pool[s.BC[or_x][pi]].was_twisted)) {
                int cI = s.BC[or_x][pi], cJ = s.BC[or_y][pj];
                s.pool[cJ].was_twisted = was_twisted;
                join_flag = true;
                if (s.pool[cI].mult >= 0 && s.pool[cJ].mult >= 0) {
                    int mi = s.pool[cI].mult, mj = s.pool[cJ].mult;
                    for (int c : s.mults[mj].clusters) {
                        s.pool[c].mult = mi;
                        s.mults[mi].clusters.push_back(c);
                    }
                    s.mults[mj].clusters.clear();
                } else if (s.pool[cI].mult >= 0 || s.pool[cJ].mult >= 0) {
                    if (s.pool[cI].mult < 0) std::swap(cI, cJ);
                    s.pool[cJ].mult = s.pool[cI].mult;
                    s.mults[s.pool[cI].mult].clusters.push_back(cJ);
                } else {
                    Multiplicon m;
                    m.x_object_id = s.g.x_object_id;
                    m.y_object_id = s.g.y_object_id;
                    m.level = s.g.level;
                    m.clusters.push_back(cI);
                    m.clusters.push_back(cJ);
                    s.mults.push_back(m);
                    s.pool[cI].mult = s.pool[cJ].mult = (int)s.mults.size() - 1;
                }
            }
        }
    }
    for (int k = 0; k < 2; k++)
        for (int idx : s.BC[k])
            if (s.pool[idx].mult < 0) {
                Multiplicon m;
                m.x_object_id = s.g.x_object_id;
                m.y_object_id = s.g.y_object_id;
                m.level = s.g.level;
                s.pool[idx].was_twisted = false;
                m.clusters.push_back(idx);
                s.mults.push_back(m);
                s.pool[idx].mult = (int)s.mults.size() - 1;
            }
    // drop empty multiplicons and those with too few real anchor points
    std::vector<Multiplicon> kept;
    for (Multiplicon& m : s.mults) {
        unsigned cnt = 0;
        for (int c : m.clusters) cnt += real_count(s.pool[c]);
        if (m.clusters.empty() || cnt < (unsigned)s.prm.anchor_points) continue;
        kept.push_back(m);
    }
    s.mults.swap(kept);
}

// GHM.cpp:157-219 (everything except setMultiplicons)
void run_collinear(GHMInput& g, const Params& prm, std::vector<Cluster>& clusters,
                   std::vector<Multiplicon>& mults)
{
#ifdef ORACLE_STUDY
    study::PerGhm study_scope;
    study::cur = &study_scope;
    study::C.ghms++;
    study_scope.dots = (long long)(g.dots[0].size() + g.dots[1].size());
    study_scope.level = g.level;
    struct StudyEnd { ~StudyEnd() { study::cur = nullptr; } } study_end;
#endif
    clusters.clear();
    mults.clear();
    if ((g.cnt[0] + g.cnt[1]) < (unsigned int)prm.anchor_points) return;
    GHMState s(g, prm, clusters, mults);
    for (int i = 0; i < 10 && i < prm.n_gaps && prm.gaps[i] > 0; i++) {
        seed_base_clusters(s, prm.gaps[i], 0);
        enrich_clusters(s, prm.gaps[i], 0, 0);
        join_clusters(s, prm.gaps[i], 0);
        seed_base_clusters(s, prm.gaps[i], 1);
        enrich_clusters(s, prm.gaps[i], 1, 1);
        join_clusters(s, prm.gaps[i], 1);
    }
#ifdef ORACLE_STUDY
    // ORACLE_STUDY_DUMP=<file>: the state after the gap-size loop (per orientation the cluster vector with
    // every cluster's anchor points in multiset order, then the dots left in the matrix) -- what
    // tests/emu/emu_join2 dumps for the emulated device kernels
    if (const char* path = getenv("ORACLE_STUDY_DUMP")) {
        static std::mutex mu;
        std::lock_guard<std::mutex> lk(mu);
        FILE* f = fopen(path, "a");
        if (f) {
            fprintf(f, "GHM %zu %zu\n", (size_t)g.cnt[0], (size_t)g.cnt[1]);
            for (int o = 0; o < 2; o++) {
                fprintf(f, " o%d clusters %zu\n", o, s.BC[o].size());
                for (int id : s.BC[o]) {
                    fprintf(f, "  c %zu:", s.pool[id].aps.size());
                    for (const AP& e : s.pool[id].aps) fprintf(f, " %d,%d", e.x, e.y);
                    fprintf(f, "\n");
                }
                fprintf(f, "  left %zu:", g.dots[o].size());
                for (const std::pair<int, int>& d : g.dots[o]) fprintf(f, " %d,%d", d.first, d.second);
                fprintf(f, "\n");
            }
            fclose(f);
        }
    }
#endif
    for (int rep = 0; rep < 2; rep++) {
        enrich_clusters(s, prm.gap_size, 0, 0);
        enrich_clusters(s, prm.gap_size, 1, 1);
        enrich_clusters(s, prm.gap_size, 1, 0);
        enrich_clusters(s, prm.gap_size, 0, 1);
        if (rep == 0) filter_clusters(s);
    }
    cluster_clusters(s);
}

// Cluster.h:27-33 + BaseCluster.cpp:99-140 + Multiplicon.cpp:135-193
static void set_bounds(std::vector<Cluster>& clusters, Multiplicon& m)
{
    bool first = true;
    for (int ci : m.clusters) {
        Cluster& c = clusters[ci];
        unsigned lx = c.aps.front().x, hx = c.aps.back().x;
        unsigned ly = c.aps.front().y, hy = c.aps.front().y;
        for (const AP& e : c.aps) {
            ly = std::min<unsigned>(ly, e.y);
            hy = std::max<unsigned>(hy, e.y);
        }
        if (first || real_count(c) > 0) {
            if (first) {
                m.begin_x = lx; m.end_x = hx; m.begin_y = ly; m.end_y = hy;
            } else {
                if (lx < (unsigned)m.begin_x) m.begin_x = lx;
                if (hx > (unsigned)m.end_x) m.end_x = hx;
                if (ly < (unsigned)m.begin_y) m.begin_y = ly;
                if (hy > (unsigned)m.end_y) m.end_y = hy;
            }
        }
        first = false;
        c.begin_x = lx; c.end_x = hx; c.begin_y = ly; c.end_y = hy;
    }
}

// GHM.cpp:671-701
void set_multiplicons_l2(const List& X, const List& Y, std::vector<Cluster>& clusters,
                         std::vector<Multiplicon>& mults)
{
    for (Multiplicon& m : mults) {
        set_bounds(clusters, m);
        for (int ci : m.clusters)
            for (AP& e : clusters[ci].aps) {
                e.gx = X.num_id[e.x];
                e.gy = Y.num_id[e.y];
            }
    }
}

// GHMProfile.cpp:68-136 (without the createHomologs filter, which needs host state)
void set_multiplicons_hl(const Profile& P, const List& Y, std::vector<Cluster>& clusters,
                         std::vector<Multiplicon>& mults, const PairTable* pt)
{
    for (Multiplicon& m : mults) {
        set_bounds(clusters, m);
        for (int ci : m.clusters) {
            Cluster& c = clusters[ci];
            // filterDuplicates (BaseCluster.cpp:63-75): unique (x,y), sorted by x then y
            std::vector<AP> uniq(c.aps);
            std::stable_sort(uniq.begin(), uniq.end(), [](const AP& l, const AP& r) {
                return l.x != r.x ? l.x < r.x : l.y < r.y;
            });
            uniq.erase(std::unique(uniq.begin(), uniq.end(),
                                   [](const AP& l, const AP& r) { return l.x == r.x && l.y == r.y; }),
                       uniq.end());
            c.aps.swap(uniq);
            std::vector<AP> falses;
            for (AP& e : c.aps) {
                bool first_match = false;
                for (int n = 0; n < P.S; n++) {
                    int k = n * P.L + e.x;
                    if (P.flags[k] & 2) continue;
                    if (!(pt ? pt->is_pair(P.gid[k], Y.gid[e.y]) : (P.fam[k] == Y.fam[e.y] && P.gid[k] != Y.gid[e.y]))) continue;
                    if (first_match) {
                        AP f;
                        f.gx = P.num_id[k]; f.gy = Y.num_id[e.y]; f.x = e.x; f.y = e.y; f.real = false;
                        falses.push_back(f);
                    } else {
                        e.gx = P.num_id[k];
                        e.gy = Y.num_id[e.y];
                        first_match = true;
                    }
                }
            }
            for (const AP& f : falses) ap_insert(c, f);  // sets b = 0.0 (BaseCluster.cpp:82-85)
        }
    }
}

// ---------------------------------------------------------------------------
// Multiplicon::createHomologs (Multiplicon.cpp:352-382) on the slices of extractXObject / extractYObject (:331-350)
bool enough_homologs(const Profile& P, const List& Y, const std::vector<Link>& links, const Multiplicon& m, int min_homologs,
                     const PairTable* pt)
{
    const int S = P.S, bx = m.begin_x, ex = m.end_x, by = m.begin_y, ey = m.end_y;
    const int wx = ex >= bx ? ex - bx + 1 : 0, wy = ey >= by ? ey - by + 1 : 0;   // GeneList(const GeneList&, begin, end), GeneList.cpp:55-63
    std::vector<std::vector<char>> hx(S, std::vector<char>(wx, 0));
    std::vector<char> hy(wy, 0);
    // pruneLinks (:384-421): numID -> element per x-segment slice; a Link stays while both ends are found
    if (!links.empty()) {
        std::vector<std::map<int, int>> le_map(S);
        for (int i = 0; i < S; i++)
            for (int j = 0; j < wx; j++) {
                const int k = i * P.L + bx + j;
                le_map[i][(P.flags[k] & 2) ? -1 : P.num_id[k]] = j;   // (a gap element carries numID -1, ListElement.cpp:5-9)
            }
        for (const Link& e : links) {
            if (e.seg_x < 0 || e.seg_x >= S || e.seg_y < 0 || e.seg_y >= S) continue;
            std::map<int, int>::const_iterator ix = le_map[e.seg_x].find(e.gene_x), iy = le_map[e.seg_y].find(e.gene_y);
            if (ix == le_map[e.seg_x].end() || iy == le_map[e.seg_y].end()) continue;
            hx[e.seg_x][ix->second] = 1;
            hx[e.seg_y][iy->second] = 1;
        }
    }
    // addLinks (:423-494)
    for (int jy = 0; jy < wy; jy++) {
        const int y = by + jy;
        for (int i = 0; i < S; i++)
            for (int j = 0; j < wx; j++) {
                const int k = i * P.L + bx + j;
                if (P.flags[k] & 2) continue;
                // pairs (:425-447): the x element's pair set holds the y gene; families (:448-493): same family and
                // not the same list element ("to have compatibility with the pairwise case", :484-485)
                const bool hom = pt ? pt->is_pair(P.gid[k], Y.gid[y]) : (P.fam[k] == Y.fam[y] && P.num_id[k] != Y.num_id[y]);
                if (!hom) continue;
                hx[i][j] = 1;
                hy[jy] = 1;
            }
    }
    // counts (:360-381)
    for (int i = 0; i < S; i++) {
        int count = 0;
        for (int j = 0; j < wx; j++) count += hx[i][j];
        if (count < min_homologs) return false;
    }
    int count = 0;
    for (int jy = 0; jy < wy; jy++) {
        count += hy[jy];
        if (count >= min_homologs) return true;
    }
    return false;
}

void filter_homologs(const Profile& P, const List& Y, const std::vector<Link>& links, int min_homologs,
                     std::vector<Multiplicon>& mults, const PairTable* pt)
{
    std::vector<Multiplicon> kept;
    for (const Multiplicon& m : mults)
        if (enough_homologs(P, Y, links, m, min_homologs, pt)) kept.push_back(m);
    mults.swap(kept);
}

// ---------------------------------------------------------------------------
template <typename T>
static void put(std::vector<uint8_t>& out, const T& v)
{
    const uint8_t* p = reinterpret_cast<const uint8_t*>(&v);
    out.insert(out.end(), p, p + sizeof(T));
}

void pack(const std::vector<Cluster>& clusters, const std::vector<Multiplicon>& mults,
          std::vector<uint8_t>& out, int& n_mult_total)
{
    for (const Multiplicon& m : mults) {
        put<int32_t>(out, m.x_object_id);
        put<int32_t>(out, m.y_object_id);
        put<int32_t>(out, m.level);
        put<uint8_t>(out, 0);  // isRedundant
        put<int32_t>(out, m.begin_x);
        put<int32_t>(out, m.end_x);
        put<int32_t>(out, m.begin_y);
        put<int32_t>(out, m.end_y);
        put<int32_t>(out, (int32_t)m.clusters.size());
        for (int ci : m.clusters) {
            const Cluster& c = clusters[ci];
            put<double>(out, c.p);
            put<uint8_t>(out, c.orient ? 1 : 0);
            put<double>(out, c.a);
            put<double>(out, c.b);
            put<double>(out, c.avg_x);
            put<double>(out, c.var_x);
            put<double>(out, c.mrss);
            put<double>(out, c.x_end1);
            put<double>(out, c.x_end2);
            put<double>(out, c.y_end1);
            put<double>(out, c.y_end2);
            put<uint8_t>(out, c.was_twisted ? 1 : 0);
            put<int32_t>(out, c.begin_x);
            put<int32_t>(out, c.end_x);
            put<int32_t>(out, c.begin_y);
            put<int32_t>(out, c.end_y);
            put<int32_t>(out, (int32_t)c.aps.size());
            for (const AP& e : c.aps) {
                put<int32_t>(out, e.gx);
                put<int32_t>(out, e.gy);
                put<int32_t>(out, e.x);
                put<int32_t>(out, e.y);
                put<uint8_t>(out, e.real ? 1 : 0);
            }
        }
        n_mult_total++;
    }
}

}  // namespace oracle

#ifdef ORACLE_STUDY
// per-GHM records since the last call (8 values each); returns the count, copies at most `cap` records
extern "C" long long oracle_study_records(long long* out, long long cap)
{
    std::lock_guard<std::mutex> lk(oracle::study::rec_mu);
    long long n = (long long)oracle::study::recs.size();
    for (long long i = 0; i < n && i < cap; i++)
        for (int k = 0; k < 8; k++) out[8 * i + k] = oracle::study::recs[i].v[k];
    oracle::study::recs.clear();
    return n;
}
extern "C" void oracle_study_set_spec(int k) { oracle::study::spec_k = k; }
extern "C" void oracle_study_set_j2_model(int on) { oracle::study::j2_model = on; }
extern "C" void oracle_study_counters(long long* out)
{
    using oracle::study::C;
    out[0] = C.ghms; out[1] = C.calls; out[2] = C.pairs; out[3] = C.near_gap; out[4] = C.near_max;
    out[5] = C.first_max; out[6] = C.clusters; out[7] = C.changed;
    out[8] = C.merges; out[9] = C.max_clusters;
    out[22] = C.spec_batches; out[23] = C.spec_offered; out[24] = C.spec_accepted;
    out[25] = C.crit_unsafe; out[26] = C.crit_lost;
    out[31] = C.j2_merges; out[32] = C.j2_mismatch;
    out[27] = C.warp_chunks; out[28] = C.warp_heavy_now; out[29] = C.warp_heavy_2phase; out[30] = C.warp_scan_2phase;
    for (int b = 0; b < 6; b++) { out[10 + b] = C.h_calls[b]; out[16 + b] = C.h_pairs[b]; }
}
#endif
