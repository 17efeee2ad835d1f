// TEST INFRASTRUCTURE -- not part of the product.
//
// Harness around the UNMODIFIED reference objects (built by oracle/Makefile from
// /root/reference/src into oracle/_ref/).  It drives the reference's own classes
// (DataSet, GHM, GHMProfile, Multiplicon, Profile) in the 1-thread order of
// levelTwo.cpp:113-175 / higherLevel.cpp:91-250 and dumps, for every GHM, the dot
// set after buildMatrix and the multiplicons in the reference's own wire format
// (Multiplicon.cpp:210-264).  The dump ("trace") is what pins oracle_core.cpp and
// the CUDA path.  Trace layout: see tests/trace_io.py.
//
// usage: ref_harness <ini> <trace.bin|-> [--no-hl] [--max-profiles N] [--max-pairs N] [--time]
//                    [--threads T] [--no-dump] [--dump-profiles file] [--dump-profile-links file]
// With --threads the GHMs of a stage are evaluated concurrently (each thread owns its GHM,
// exactly like the reference's work functions, DataSet.h:79,204); results are consumed in
// sequence order, so the trace is identical to the 1-thread one.  --no-dump (or "-" as the
// trace path) turns the harness into the CPU timing baseline used by bench.py.
//
// Private members of the reference classes are reached with the macro below;
// this TU is the only one compiled that way (object layout is unaffected).
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <map>
#include <set>
#include <list>
#include <queue>
#include <string>
#include <vector>
#include <deque>
#include <chrono>
#include <iostream>
#include <fstream>
#include <sstream>
#include <iomanip>
#include <algorithm>
#include <stdexcept>
#include <cmath>
#include <cassert>
#include <atomic>
#include <thread>
#include <functional>
#include <ext/hash_map>
#include <ext/hash_set>
#include <pthread.h>
#include "headers.h"
#define private public
#define protected public
#include "DataSet.h"
#include "Settings.h"
#include "GeneList.h"
#include "ListElement.h"
#include "Gene.h"
#include "GHM.h"
#include "GHMProfile.h"
#include "Multiplicon.h"
#include "BaseCluster.h"
#include "Profile.h"
#include "parallel.h"
#include "util.h"
#undef private
#undef protected

using namespace std;

typedef std::string Blob;
static void w32(Blob& b, int32_t v) { b.append((const char*)&v, 4); }
static void w8(Blob& b, uint8_t v) { b.append((const char*)&v, 1); }
static void wf64(Blob& b, double v) { b.append((const char*)&v, 8); }

static map<string, int> famIds, geneIds;
static int intern(map<string, int>& m, const string& s)
{
    map<string, int>::iterator it = m.find(s);
    if (it != m.end()) return it->second;
    int id = (int)m.size();
    m[s] = id;
    return id;
}

static double now()
{
    return chrono::duration<double>(chrono::steady_clock::now().time_since_epoch()).count();
}

static void dumpDots(Blob& b, const GHM& ghm)
{
    for (int o = 0; o < 2; o++) {
        int n = 0;
        for (map<int, set<int> >::const_iterator ix = ghm.matrix[o].begin(); ix != ghm.matrix[o].end(); ix++)
            n += ix->second.size();
        w32(b, n);
        for (map<int, set<int> >::const_iterator ix = ghm.matrix[o].begin(); ix != ghm.matrix[o].end(); ix++)
            for (set<int>::const_iterator iy = ix->second.begin(); iy != ix->second.end(); iy++) {
                w32(b, ix->first);
                w32(b, *iy);
            }
    }
}

static void dumpPacked(Blob& b, const vector<Multiplicon*>& mps)
{
    int sz = Multiplicon::getPackSize(mps);
    vector<char> buf(sz);
    Multiplicon::packMultiplicons(mps, buf.data());
    w32(b, sz);
    b.append(buf.data(), sz);
}

// GHMProfile whose setMultiplicons never drops a multiplicon (minHomologs = 0):
// this is the state the device path returns, before the host-side homolog filter
// of GHMProfile.cpp:129-132.
class GHMProfileNoFilter : public GHMProfile {
public:
    GHMProfileNoFilter(const Profile& p, const GeneList& y) : GHMProfile(p, y) {}
protected:
    virtual void setMultiplicons(bool useFamily, int minHomologs) {
        GHMProfile::setMultiplicons(useFamily, 0);
    }
};

struct ItemOut {
    Blob blob;
    vector<Multiplicon*> mps;
    double tBuild, tRun;
    long dots;
    ItemOut() : tBuild(0), tRun(0), dots(0) {}
};

static void parallel_for(size_t n, int threads, const std::function<void(size_t)>& fn)
{
    if (threads <= 1 || n < 2) {
        for (size_t i = 0; i < n; i++) fn(i);
        return;
    }
    std::atomic<size_t> next(0);
    vector<std::thread> th;
    for (int t = 0; t < threads; t++)
        th.emplace_back([&]() {
            while (true) {
                size_t i = next.fetch_add(1);
                if (i >= n) break;
                fn(i);
            }
        });
    for (size_t t = 0; t < th.size(); t++) th[t].join();
}

int main(int argc, char** argv)
{
    if (argc < 3) {
        fprintf(stderr, "usage: %s <ini> <trace.bin|-> [--no-hl] [--max-profiles N] [--time] [--threads T] [--no-dump]\n", argv[0]);
        return 2;
    }
    bool doHL = true, timing = false, dump = strcmp(argv[2], "-") != 0;
    // --dump-profiles <file>: only the aligned profiles the reference searches, in order, each preceded by the
    // masks applied before it (bench.py replays them on the device so that both arms search the same profiles):
    //   i32 n_masks, n_masks x (i32 list, begin, end); i32 S, L, x_object_id; S*L x (i32 numID or -1 = gap, u8 orientation)
    FILE* profOut = NULL;
    // --dump-profile-links <file>: beside --dump-profiles, per dumped profile the Link set its search inherits
    // (Multiplicon.cpp:343-344): i32 n_links, n_links x (i32 segment_x, segment_y, gene_x_id, gene_y_id)
    FILE* linkOut = NULL;
    long maxProfiles = -1, maxPairs = -1;
    int threads = 1;
    for (int i = 3; i < argc; i++) {
        if (!strcmp(argv[i], "--no-hl")) doHL = false;
        else if (!strcmp(argv[i], "--time")) timing = true;
        else if (!strcmp(argv[i], "--no-dump")) dump = false;
        else if (!strcmp(argv[i], "--max-profiles") && i + 1 < argc) maxProfiles = atol(argv[++i]);
        else if (!strcmp(argv[i], "--max-pairs") && i + 1 < argc) maxPairs = atol(argv[++i]);
        else if (!strcmp(argv[i], "--threads") && i + 1 < argc) threads = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--dump-profiles") && i + 1 < argc) profOut = fopen(argv[++i], "wb");
        else if (!strcmp(argv[i], "--dump-profile-links") && i + 1 < argc) linkOut = fopen(argv[++i], "wb");
    }

    ParToolBox::init(&argc, &argv);
    Settings settings(argv[1]);
    // silence the reference's chatter
    streambuf* coutbuf = cout.rdbuf();
    cout.rdbuf(NULL);
    DataSet ds(settings);
    ds.mapGenes();
    ds.remapTandems();
    ds.sortGeneLists();
    cout.rdbuf(coutbuf);

    FILE* out = NULL;
    if (dump) {
        out = fopen(argv[2], "wb");
        if (!out) { perror("trace"); return 1; }
    }
    Blob hdr;
    // "IADHTRC2" = table_type=pairs: the header carries the gene-pair sets after the dataset
    hdr.append(settings.useFamily() ? "IADHTRC1" : "IADHTRC2", 8);

    // ---- params -----------------------------------------------------------------
    int gaps[10];
    for (int i = 0; i < 10; i++) gaps[i] = 0;
    settings.getGapSizes(gaps);
    int ngaps = 0;
    while (ngaps < 10 && gaps[ngaps] > 0) ngaps++;
    w32(hdr, ngaps);
    for (int i = 0; i < 10; i++) w32(hdr, i < ngaps ? gaps[i] : 0);
    w32(hdr, settings.getGapSize());
    w32(hdr, settings.getClusterGap());
    w32(hdr, settings.getTandemGap());
    w32(hdr, settings.getAnchorPoints());
    w32(hdr, (int)settings.getMultHypCorMethod()); // None=0, Bonferroni=1, FDR=2 (Settings.h)
    w32(hdr, settings.level2Only() ? 1 : 0);
    wf64(hdr, settings.getQValue());
    wf64(hdr, settings.getProbCutoff());

    // ---- dataset ----------------------------------------------------------------
    const vector<GeneList*>& gl = ds.genelists;
    w32(hdr, (int)gl.size());
    for (size_t l = 0; l < gl.size(); l++) {
        const vector<ListElement*>& re = gl[l]->getRemappedElements();
        const vector<ListElement*>& all = gl[l]->getElements();
        w32(hdr, gl[l]->getID());
        w32(hdr, intern(famIds, "genome:" + gl[l]->getGenomeName())); // genome id (for `identical`)
        // raw (pre tandem-remap) elements: lets the python ingest be checked too
        w32(hdr, (int)all.size());
        for (size_t i = 0; i < all.size(); i++) {
            const Gene& g = all[i]->getGene();
            w32(hdr, intern(famIds, g.getFamily()));
            w32(hdr, intern(geneIds, g.getID()));
            w32(hdr, all[i]->getNumID());
            w8(hdr, all[i]->getOrientation() ? 1 : 0);
            w8(hdr, g.isRemapped() ? 1 : 0);
        }
        w32(hdr, (int)re.size());
        for (size_t i = 0; i < re.size(); i++) {
            const Gene& g = re[i]->getGene();
            w32(hdr, intern(famIds, g.getFamily()));
            w32(hdr, intern(geneIds, g.getID()));
            w32(hdr, re[i]->getNumID());
            w8(hdr, re[i]->getOrientation() ? 1 : 0);
        }
    }
    w32(hdr, (int)ds.indexToList.size());
    for (size_t i = 0; i < ds.indexToList.size(); i++) w32(hdr, ds.indexToList[i]);
    if (!settings.useFamily()) {
        // pair sets as DataSet::getGenePairs left them (DataSet.cpp:144-186), by interned gene id
        vector<vector<int> > ps(geneIds.size());
        for (size_t l = 0; l < gl.size(); l++) {
            const vector<ListElement*>& all = gl[l]->getElements();
            for (size_t i = 0; i < all.size(); i++) {
                const Gene& g = all[i]->getGene();
                if (!g.pairs) continue;
                vector<int>& v = ps[intern(geneIds, g.getID())];
                if (!v.empty()) continue;  // (a repeated gene ID shares one set)
                for (hash_set<string, stringhash>::const_iterator it = g.pairs->begin(); it != g.pairs->end(); it++)
                    v.push_back(intern(geneIds, *it));
                sort(v.begin(), v.end());
            }
        }
        w32(hdr, (int)ps.size());
        for (size_t g = 0; g < ps.size(); g++) {
            w32(hdr, (int)ps[g].size());
            for (size_t k = 0; k < ps[g].size(); k++) w32(hdr, ps[g][k]);
        }
    }
    if (dump) fwrite(hdr.data(), 1, hdr.size(), out);

    // ---- level 2: every list pair in the 1-thread order of getSomeWorkL2 ----------
    vector<pair<unsigned, unsigned> > pairs;
    {
        unsigned n = gl.size();
        unsigned cX = 0, cY = 0;
        while (cX < n) {
            pairs.push_back(make_pair(ds.indexToList[cX], ds.indexToList[cY]));
            // next pair (levelTwo.cpp:152-164)
            if ((cX - cY) < 2) {
                if (cX == cY && cX == n - 1) break;
                unsigned newD = cX + cY + 1;
                cX = (newD < n) ? newD : n - 1;
                cY = (newD < n) ? 0 : newD - n + 1;
            } else { cX--; cY++; }
        }
    }
    // bounded sample for benchmarks: only the first maxPairs pairs of the sequence
    if (maxPairs >= 0 && (size_t)maxPairs < pairs.size()) pairs.resize((size_t)maxPairs);
    vector<ItemOut> l2out(pairs.size());
    const bool useFamily = settings.useFamily();
    double wall0 = now();
    parallel_for(pairs.size(), threads, [&](size_t i) {
        ItemOut& o = l2out[i];
        unsigned lX = pairs[i].first, lY = pairs[i].second;
        double t0 = now();
        GHM ghm(*gl[lX], *gl[lY], false);
        ghm.buildMatrix(useFamily);
        double t1 = now();
        if (dump) {
            w32(o.blob, 1);
            w32(o.blob, lX);
            w32(o.blob, lY);
            dumpDots(o.blob, ghm);
        }
        o.dots = ghm.count_points[0] + ghm.count_points[1];
        double t2 = now();
        ghm.run(settings);
        double t3 = now();
        o.tBuild = t1 - t0;
        o.tRun = t3 - t2;
        ghm.getMultiplicons(o.mps);
        if (dump) dumpPacked(o.blob, o.mps);
    });
    double wallL2 = now() - wall0;
    double tBuild = 0, tRun = 0;
    long nDots = 0;
    vector<Multiplicon*> l2;
    for (size_t i = 0; i < l2out.size(); i++) {
        if (dump) fwrite(l2out[i].blob.data(), 1, l2out[i].blob.size(), out);
        Blob().swap(l2out[i].blob);
        tBuild += l2out[i].tBuild;
        tRun += l2out[i].tRun;
        nDots += l2out[i].dots;
        l2.insert(l2.end(), l2out[i].mps.begin(), l2out[i].mps.end());
    }
    if (timing)
        fprintf(stderr, "L2: pairs=%zu dots=%ld build=%.6fs run=%.6fs wall=%.6fs threads=%d multiplicons=%zu\n",
                pairs.size(), nDots, tBuild, tRun, wallL2, threads, l2.size());

    // ---- higher level: mirror of DataSet::profileDetection ------------------------------
    long nProfiles = 0, nGhmp = 0, nFound = 0;
    double tHLbuild = 0, tHLrun = 0, wallHL = 0;
    if (doHL && !settings.level2Only()) {
        cout.rdbuf(NULL);
        ds.sortByMultipliconSize(l2);
        deque<Multiplicon*> todo;
        for (size_t i = 0; i < l2.size(); i++) todo.push_front(l2[i]);
        unsigned profile_id = 1;
        unsigned mid = 1;
        const unsigned minAP = (unsigned)settings.getAnchorPoints();
        while (!todo.empty()) {
            if (maxProfiles >= 0 && nProfiles >= maxProfiles) break;
            Multiplicon* m = todo.front();
            todo.pop_front();
            try {
                GeneList& lY = *gl[m->getYObjectID()];
                m->createProfile(profile_id);
                const Profile* profile = m->getProfile();
                if (m->getLevel() == 2) {
                    if (ds.allMasked(lY, m->getBeginY(), m->getEndY())) {
                        m->align(settings.getAlignmentMethod(), settings.getMaxGapsInAlignment());
                        throw ProfileException("level-2 multiplicon is redundant");
                    }
                }
                if (ds.allMasked(lY, m->getBeginY(), m->getEndY()))
                    throw ProfileException("all elements masked");
                m->align(settings.getAlignmentMethod(), settings.getMaxGapsInAlignment());
                if (m->getLevel() > 2) m->checkAlignment(settings.getAnchorPoints());
                Blob pb;
                if (m->getLevel() == 2) {
                    GeneList& lX = *gl[m->getXObjectID()];
                    lX.mask(m->getBeginX(), m->getEndX());
                    w32(pb, 2); w32(pb, m->getXObjectID()); w32(pb, m->getBeginX()); w32(pb, m->getEndX());
                }
                lY.mask(m->getBeginY(), m->getEndY());
                w32(pb, 2); w32(pb, m->getYObjectID()); w32(pb, m->getBeginY()); w32(pb, m->getEndY());

                // profile descriptor
                const vector<GeneList*>& segs = profile->getSegments();
                int S = segs.size(), L = segs[0]->getRemappedElements().size();
                w32(pb, 3); w32(pb, S); w32(pb, L); w32(pb, segs[0]->getID());
                if (dump)
                    for (int s = 0; s < S; s++) {
                        const vector<ListElement*>& e = segs[s]->getRemappedElements();
                        if ((int)e.size() != L) { fprintf(stderr, "ragged profile\n"); return 1; }
                        for (int x = 0; x < L; x++) {
                            if (e[x]->isGap()) { w32(pb, -1); w32(pb, -1); w32(pb, -1); w8(pb, 2); continue; }
                            const Gene& g = e[x]->getGene();
                            w32(pb, intern(famIds, g.getFamily()));
                            w32(pb, intern(geneIds, g.getID()));
                            w32(pb, e[x]->getNumID());
                            w8(pb, e[x]->getOrientation() ? 1 : 0);
                        }
                    }
                if (dump) {
                    // the Link set the search inherits from the profile's multiplicon (Multiplicon.cpp:343-344): input
                    // of the homolog-count filter (createHomologs, Multiplicon.cpp:352-382)
                    const set<Link>& hs = profile->getMultiplicon().getHomologs();
                    w32(pb, 6); w32(pb, (int)hs.size());
                    for (set<Link>::const_iterator e = hs.begin(); e != hs.end(); e++) {
                        w32(pb, e->segmentX); w32(pb, e->segmentY); w32(pb, e->geneXID); w32(pb, e->geneYID);
                    }
                }
                if (dump) fwrite(pb.data(), 1, pb.size(), out);
                if (profOut) {
                    Blob q;
                    const bool l2m = m->getLevel() == 2;
                    w32(q, l2m ? 2 : 1);
                    if (l2m) { w32(q, m->getXObjectID()); w32(q, m->getBeginX()); w32(q, m->getEndX()); }
                    w32(q, m->getYObjectID()); w32(q, m->getBeginY()); w32(q, m->getEndY());
                    w32(q, S); w32(q, L); w32(q, segs[0]->getID());
                    for (int s = 0; s < S; s++) {
                        const vector<ListElement*>& e = segs[s]->getRemappedElements();
                        for (int x = 0; x < L; x++) {
                            if (e[x]->isGap()) { w32(q, -1); w8(q, 0); continue; }
                            w32(q, e[x]->getNumID());
                            w8(q, e[x]->getOrientation() ? 1 : 0);
                        }
                    }
                    fwrite(q.data(), 1, q.size(), profOut);
                }
                if (linkOut) {
                    Blob q;
                    const set<Link>& hs = profile->getMultiplicon().getHomologs();
                    w32(q, (int)hs.size());
                    for (set<Link>::const_iterator e = hs.begin(); e != hs.end(); e++) {
                        w32(q, e->segmentX); w32(q, e->segmentY); w32(q, e->geneXID); w32(q, e->geneYID);
                    }
                    fwrite(q.data(), 1, q.size(), linkOut);
                }
                nProfiles++;

                // the lists searched for this profile (higherLevel.cpp:36-41)
                vector<int> lists;
                for (size_t yi = 0; yi < ds.indexToList.size(); yi++) {
                    int lst = ds.indexToList[yi];
                    GeneList& g = *gl[lst];
                    if (g.getSize() - g.getNumberOfMaskedElements() >= minAP) lists.push_back(lst);
                }
                vector<ItemOut> hout(lists.size());
                double w0 = now();
                parallel_for(lists.size(), threads, [&](size_t i) {
                    ItemOut& o = hout[i];
                    GeneList& g = *gl[lists[i]];
                    vector<Multiplicon*> pre;
                    if (dump) {
                        // (a) unfiltered run: what the device path must reproduce
                        GHMProfileNoFilter ghm(*profile, g);
                        ghm.buildMatrix();
                        w32(o.blob, 4); w32(o.blob, lists[i]);
                        dumpDots(o.blob, ghm);
                        ghm.run(settings);
                        ghm.getMultiplicons(pre);
                        dumpPacked(o.blob, pre);
                    }
                    // (b) the reference's own run (with the homolog filter)
                    {
                        double t0 = now();
                        GHMProfile ghm(*profile, g);
                        ghm.buildMatrix();
                        double t1 = now();
                        ghm.run(settings);
                        double t2 = now();
                        o.tBuild = t1 - t0;
                        o.tRun = t2 - t1;
                        ghm.getMultiplicons(o.mps);
                    }
                    if (dump) {
                        // kept flags: the filtered result is an in-order subsequence of `pre`
                        w32(o.blob, (int)pre.size());
                        size_t k = 0;
                        for (size_t j = 0; j < pre.size(); j++) {
                            bool kept = false;
                            if (k < o.mps.size() && pre[j]->getPackSize() == o.mps[k]->getPackSize()) {
                                vector<char> a(pre[j]->getPackSize()), b(o.mps[k]->getPackSize());
                                pre[j]->pack(a.data()); o.mps[k]->pack(b.data());
                                kept = (a == b);
                            }
                            if (kept) k++;
                            w8(o.blob, kept ? 1 : 0);
                        }
                        if (k != o.mps.size()) { fprintf(stderr, "kept-matching failed\n"); exit(1); }
                        for (size_t j = 0; j < pre.size(); j++) delete pre[j];
                    }
                });
                wallHL += now() - w0;
                vector<Multiplicon*> found;
                for (size_t i = 0; i < hout.size(); i++) {
                    if (dump) fwrite(hout[i].blob.data(), 1, hout[i].blob.size(), out);
                    tHLbuild += hout[i].tBuild;
                    tHLrun += hout[i].tRun;
                    nGhmp++;
                    found.insert(found.end(), hout[i].mps.begin(), hout[i].mps.end());
                }
                ds.sortByMultipliconSize(found);
                nFound += found.size();
                if (dump) { Blob e; w32(e, 5); w32(e, (int)found.size()); fwrite(e.data(), 1, e.size(), out); }
                for (size_t i = 0; i < found.size(); i++) {
                    found[i]->setParentID(mid);
                    todo.push_front(found[i]);
                }
                m->setIsRedundant(false);
                m->setId(mid++);
                profile_id++;
            } catch (const ProfileException& e) {
                if (m->getLevel() == 2) { m->setIsRedundant(true); m->setId(mid++); }
                // (higher-level ones are deleted by the reference; we just leak them)
            }
        }
        cout.rdbuf(coutbuf);
    }
    if (dump) {
        Blob e;
        w32(e, 0);
        fwrite(e.data(), 1, e.size(), out);
        fclose(out);
    }
    if (profOut) fclose(profOut);
    if (linkOut) fclose(linkOut);
    if (timing)
        fprintf(stderr, "HL: profiles=%ld ghmprofiles=%ld build=%.6fs run=%.6fs wall=%.6fs threads=%d found=%ld\n",
                nProfiles, nGhmp, tHLbuild, tHLrun, wallHL, threads, nFound);
    return 0;
}
