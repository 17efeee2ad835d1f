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
// usage: ref_harness <ini> <trace.bin> [--no-hl] [--max-profiles N] [--time]
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

#include <cstdio>
#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include <deque>
#include <chrono>

using namespace std;

static FILE* out = NULL;
static void w32(int32_t v) { fwrite(&v, 4, 1, out); }
static void w8(uint8_t v) { fwrite(&v, 1, 1, out); }
static void wf64(double v) { fwrite(&v, 8, 1, out); }

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

static void dumpDots(const GHM& ghm)
{
    for (int o = 0; o < 2; o++) {
        int n = 0;
        for (map<int, set<int> >::const_iterator ix = ghm.matrix[o].begin(); ix != ghm.matrix[o].end(); ix++)
            n += ix->second.size();
        w32(n);
        for (map<int, set<int> >::const_iterator ix = ghm.matrix[o].begin(); ix != ghm.matrix[o].end(); ix++)
            for (set<int>::const_iterator iy = ix->second.begin(); iy != ix->second.end(); iy++) {
                w32(ix->first);
                w32(*iy);
            }
    }
}

static void dumpPacked(const vector<Multiplicon*>& mps)
{
    int sz = Multiplicon::getPackSize(mps);
    vector<char> buf(sz);
    Multiplicon::packMultiplicons(mps, buf.data());
    w32(sz);
    fwrite(buf.data(), 1, sz, out);
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

int main(int argc, char** argv)
{
    if (argc < 3) {
        fprintf(stderr, "usage: %s <ini> <trace.bin> [--no-hl] [--max-profiles N] [--time]\n", argv[0]);
        return 2;
    }
    bool doHL = true, timing = false;
    long maxProfiles = -1;
    for (int i = 3; i < argc; i++) {
        if (!strcmp(argv[i], "--no-hl")) doHL = false;
        else if (!strcmp(argv[i], "--time")) timing = true;
        else if (!strcmp(argv[i], "--max-profiles") && i + 1 < argc) maxProfiles = atol(argv[++i]);
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

    out = fopen(argv[2], "wb");
    if (!out) { perror("trace"); return 1; }
    fwrite("IADHTRC1", 1, 8, out);

    // ---- params -----------------------------------------------------------------
    int gaps[10];
    for (int i = 0; i < 10; i++) gaps[i] = 0;
    settings.getGapSizes(gaps);
    int ngaps = 0;
    while (ngaps < 10 && gaps[ngaps] > 0) ngaps++;
    w32(ngaps);
    for (int i = 0; i < 10; i++) w32(i < ngaps ? gaps[i] : 0);
    w32(settings.getGapSize());
    w32(settings.getClusterGap());
    w32(settings.getTandemGap());
    w32(settings.getAnchorPoints());
    w32((int)settings.getMultHypCorMethod()); // None=0, Bonferroni=1, FDR=2 (Settings.h)
    w32(settings.level2Only() ? 1 : 0);
    wf64(settings.getQValue());
    wf64(settings.getProbCutoff());

    // ---- dataset ----------------------------------------------------------------
    const vector<GeneList*>& gl = ds.genelists;
    w32((int)gl.size());
    for (size_t l = 0; l < gl.size(); l++) {
        const vector<ListElement*>& re = gl[l]->getRemappedElements();
        const vector<ListElement*>& all = gl[l]->getElements();
        w32(gl[l]->getID());
        w32(intern(famIds, "genome:" + gl[l]->getGenomeName())); // genome id (for `identical`)
        // raw (pre tandem-remap) elements: lets the python ingest be checked too
        w32((int)all.size());
        for (size_t i = 0; i < all.size(); i++) {
            const Gene& g = all[i]->getGene();
            w32(intern(famIds, g.getFamily()));
            w32(intern(geneIds, g.getID()));
            w32(all[i]->getNumID());
            w8(all[i]->getOrientation() ? 1 : 0);
            w8(g.isRemapped() ? 1 : 0);
        }
        w32((int)re.size());
        for (size_t i = 0; i < re.size(); i++) {
            const Gene& g = re[i]->getGene();
            w32(intern(famIds, g.getFamily()));
            w32(intern(geneIds, g.getID()));
            w32(re[i]->getNumID());
            w8(re[i]->getOrientation() ? 1 : 0);
        }
    }
    w32((int)ds.indexToList.size());
    for (size_t i = 0; i < ds.indexToList.size(); i++) w32(ds.indexToList[i]);

    // ---- level 2: every list pair in the 1-thread order of getSomeWorkL2 ----------
    double tBuild = 0, tRun = 0;
    long nPairs = 0, nDots = 0;
    vector<Multiplicon*> l2;
    {
        unsigned n = gl.size();
        unsigned cX = 0, cY = 0;
        while (cX < n) {
            unsigned lX = ds.indexToList[cX], lY = ds.indexToList[cY];
            double t0 = now();
            GHM ghm(*gl[lX], *gl[lY], false);
            ghm.buildMatrix(settings.useFamily());
            double t1 = now();
            w32(1);
            w32(lX);
            w32(lY);
            dumpDots(ghm);
            nDots += ghm.count_points[0] + ghm.count_points[1];
            double t2 = now();
            ghm.run(settings);
            double t3 = now();
            tBuild += t1 - t0;
            tRun += t3 - t2;
            nPairs++;
            vector<Multiplicon*> mps;
            ghm.getMultiplicons(mps);
            dumpPacked(mps);
            l2.insert(l2.end(), mps.begin(), mps.end());
            // next pair (levelTwo.cpp:152-164)
            if ((cX - cY) < 2) {
                if (cX == cY && cX == n - 1) { cX = n; break; }
                unsigned newD = cX + cY + 1;
                cX = (newD < n) ? newD : n - 1;
                cY = (newD < n) ? 0 : newD - n + 1;
            } else { cX--; cY++; }
        }
    }
    if (timing)
        fprintf(stderr, "L2: pairs=%ld dots=%ld build=%.6fs run=%.6fs multiplicons=%zu\n",
                nPairs, nDots, tBuild, tRun, l2.size());

    // ---- higher level: mirror of DataSet::profileDetection, 1 thread ----------------
    long nProfiles = 0, nGhmp = 0;
    double tHLbuild = 0, tHLrun = 0;
    if (doHL && !settings.level2Only()) {
        cout.rdbuf(NULL);
        ds.sortByMultipliconSize(l2);
        deque<Multiplicon*> todo;
        for (size_t i = 0; i < l2.size(); i++) todo.push_front(l2[i]);
        unsigned profile_id = 1;
        unsigned mid = 1;
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
                if (m->getLevel() == 2) {
                    GeneList& lX = *gl[m->getXObjectID()];
                    lX.mask(m->getBeginX(), m->getEndX());
                    w32(2); w32(m->getXObjectID()); w32(m->getBeginX()); w32(m->getEndX());
                }
                lY.mask(m->getBeginY(), m->getEndY());
                w32(2); w32(m->getYObjectID()); w32(m->getBeginY()); w32(m->getEndY());

                // profile descriptor
                const vector<GeneList*>& segs = profile->getSegments();
                int S = segs.size(), L = segs[0]->getRemappedElements().size();
                w32(3); w32(S); w32(L); w32(segs[0]->getID());
                for (int s = 0; s < S; s++) {
                    const vector<ListElement*>& e = segs[s]->getRemappedElements();
                    if ((int)e.size() != L) { fprintf(stderr, "ragged profile\n"); return 1; }
                    for (int x = 0; x < L; x++) {
                        if (e[x]->isGap()) { w32(-1); w32(-1); w32(-1); w8(2); continue; }
                        const Gene& g = e[x]->getGene();
                        w32(intern(famIds, g.getFamily()));
                        w32(intern(geneIds, g.getID()));
                        w32(e[x]->getNumID());
                        w8(e[x]->getOrientation() ? 1 : 0);
                    }
                }
                nProfiles++;

                vector<Multiplicon*> found;
                for (size_t yi = 0; yi < ds.indexToList.size(); yi++) {
                    int lst = ds.indexToList[yi];
                    GeneList& g = *gl[lst];
                    if (g.getSize() - g.getNumberOfMaskedElements() < (unsigned)settings.getAnchorPoints())
                        continue;
                    // (a) unfiltered run: what the device path must reproduce
                    vector<Multiplicon*> pre;
                    {
                        double t0 = now();
                        GHMProfileNoFilter ghm(*profile, g);
                        ghm.buildMatrix();
                        double t1 = now();
                        w32(4); w32(lst);
                        dumpDots(ghm);
                        ghm.run(settings);
                        ghm.getMultiplicons(pre);
                        dumpPacked(pre);
                        (void)t0; (void)t1;
                    }
                    // (b) the reference's own run (with the homolog filter)
                    vector<Multiplicon*> post;
                    {
                        double t0 = now();
                        GHMProfile ghm(*profile, g);
                        ghm.buildMatrix();
                        double t1 = now();
                        ghm.run(settings);
                        double t2 = now();
                        tHLbuild += t1 - t0; tHLrun += t2 - t1; nGhmp++;
                        ghm.getMultiplicons(post);
                    }
                    // kept flags: post is an in-order subsequence of pre
                    w32((int)pre.size());
                    size_t k = 0;
                    for (size_t i = 0; i < pre.size(); i++) {
                        bool kept = false;
                        if (k < post.size() && pre[i]->getPackSize() == post[k]->getPackSize()) {
                            vector<char> a(pre[i]->getPackSize()), b(post[k]->getPackSize());
                            pre[i]->pack(a.data()); post[k]->pack(b.data());
                            kept = (a == b);
                        }
                        if (kept) k++;
                        w8(kept ? 1 : 0);
                    }
                    if (k != post.size()) { fprintf(stderr, "kept-matching failed\n"); return 1; }
                    for (size_t i = 0; i < pre.size(); i++) delete pre[i];
                    found.insert(found.end(), post.begin(), post.end());
                }
                ds.sortByMultipliconSize(found);
                w32(5); w32((int)found.size());
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
    w32(0);
    fclose(out);
    if (timing)
        fprintf(stderr, "HL: profiles=%ld ghmprofiles=%ld build=%.6fs run=%.6fs\n",
                nProfiles, nGhmp, tHLbuild, tHLrun);
    return 0;
}
