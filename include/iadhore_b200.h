/*
 * iadhore_b200.h -- C-ABI of the B200-native collinearity engine.
 *
 * Drop-in boundary for i-ADHoRe 3.0's data-parallel hot path:
 * the reference has no plugin ABI, so these entry points are exactly what a shim
 * replacing DataSet::parallelLevel2ADHoReDyn (levelTwo.cpp:177-292) and
 * DataSet::parallelProfileSearch (higherLevel.cpp:252-326) binds.  Everything the
 * reference computes inside GHM::buildMatrix / GHMProfile::buildMatrix / GHM::run
 * (GHM.cpp:34-148,157-219; GHMProfile.cpp:17-136; BaseCluster.cpp; Cluster.cpp) runs on
 * the GPU behind these calls; results come back in the reference's own MPI wire format
 * (Multiplicon::packMultiplicons, Multiplicon.cpp:242-264), so the host rehydrates them
 * with the reference's unpack constructors (Multiplicon.cpp:19-96, 266-297).
 *
 * Plain C, plain pointers and sizes; no torch / CUDA types.  All functions return 0 on
 * success, non-zero on error (never throw); iadh_last_error() describes the last failure
 * of the calling thread.  A context is single-caller (the reference calls the two drivers
 * from the master thread only).  There is NO CPU fallback: iadh_create fails if no
 * CUDA device is usable.
 */
#ifndef IADHORE_B200_H
#define IADHORE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IADH_OK 0
#define IADH_ERR_ARG 1
#define IADH_ERR_CUDA 2
#define IADH_ERR_STATE 3
#define IADH_ERR_CAPACITY 4

/* element flag bits (dataset and profile descriptors) */
#define IADH_F_STRAND 1u /* ListElement::getOrientation() (ListElement.h) */
#define IADH_F_GAP 2u    /* profile only: ListElement::isGap()            */
#define IADH_F_MASKED 4u /* dataset only: ListElement::isMasked()         */

/* multiple-hypothesis correction, values of enum MultHypCor (Settings.h:23-27) */
#define IADH_MH_NONE 0
#define IADH_MH_BONFERRONI 1
#define IADH_MH_FDR 2

typedef struct iadh_ctx iadh_ctx;
typedef struct iadh_result iadh_result;

/* The DataSet after mapGenes() + remapTandems() + sortGeneLists() (iADHoRe.cpp:62-75),
 * flattened: list l owns elements [list_off[l], list_off[l+1]) = its
 * GeneList::getRemappedElements() in order.  List index == GeneList::getID().
 *   fam     dense id of Gene::getFamily()          (GHM.cpp:49,59)
 *   gid     dense id of the Gene::getID() string   (GHM.cpp:68; Gene.cpp:20)
 *   num_id  ListElement::getNumID()                (GHM.cpp:691)
 *   flags   IADH_F_STRAND | IADH_F_MASKED
 * Two lists are "identical" (GHM.cpp:19-22) iff they are the same list index. */
typedef struct {
    int32_t n_lists;
    const int64_t* list_off; /* n_lists + 1 */
    const int32_t* fam;
    const int32_t* gid;
    const int32_t* num_id;
    const uint8_t* flags;
} iadh_dataset_desc;

/* Settings surface consumed by the hot path (Settings.h:82-161, Settings.cpp:585-598). */
typedef struct {
    int32_t n_gaps;   /* number of valid entries of gaps[] (Settings::getGapSizes) */
    int32_t gaps[10];
    int32_t gap_size;
    int32_t cluster_gap;
    int32_t anchor_points;
    int32_t mult_hyp; /* IADH_MH_* */
    double q_value;
    double prob_cutoff;
} iadh_params;

/* One aligned Profile (Profile.h: getSegments(), all S segments have L positions):
 * element (s, x) is at index s*L + x.  flags = IADH_F_STRAND | IADH_F_GAP; fam/gid/num_id
 * are ignored for gaps.  x_object_id = getSegments()[0]->getID() (GHMProfile.cpp:13). */
typedef struct {
    int32_t n_segments; /* S; multiplicon level of the results is S + 1 */
    int32_t length;     /* L */
    int32_t x_object_id;
    const int32_t* fam;
    const int32_t* gid;
    const int32_t* num_id;
    const uint8_t* flags;
} iadh_profile_desc;

typedef struct {
    int64_t n_ghm;           /* GHMs evaluated (list pairs or profile x list)            */
    int64_t n_dots;          /* anchor-point candidates emitted by the build kernels      */
    int64_t n_multiplicons;  /* multiplicons in the packed stream                         */
    int64_t n_baseclusters;
    int64_t n_anchorpoints;
    double build_ms;         /* device time of the GHM build kernels (CUDA events)        */
    double cluster_ms;       /* device time of the clustering kernel                      */
    double total_ms;         /* device time of the whole call incl. copies                */
    int64_t kernel_launches; /* kernels launched by this call                             */
    int64_t h2d_bytes, d2h_bytes;
} iadh_stats;

/* Replaces: DataSet::remapTandems (DataSet.cpp:204-213) -> GeneList::remapTandemsFamily
 * (GeneList.cpp:143-200) with Gene::remapTo (Gene.cpp:51-62), family mode, on the device.
 * `lists` describes the gene lists BEFORE remapping (GeneList::getElements() in order; num_id is
 * ignored, flags = IADH_F_STRAND).  remapped_out[i] = Gene::isRemapped() of element i afterwards;
 * the remapped lists handed to iadh_create are the elements with remapped_out == 0, in order. */
int iadh_remap_tandems(const iadh_dataset_desc* lists, int32_t tandem_gap, int device,
                       uint8_t* remapped_out);

/* The same for table_type=pairs: GeneList::remapTandemsPairs (GeneList.cpp:84-141) with
 * Gene::isIndirectPairWith (Gene.cpp:33-49: a partner, or a partner in common) and the hasPairs test.  Pair sets as
 * for iadh_create_pairs: a CSR over the dense gene ids of lists->gid, partners ascending; lists->fam is ignored. */
int iadh_remap_tandems_pairs(const iadh_dataset_desc* lists, const int64_t* pair_ptr, const int32_t* pair_ids,
                             int32_t n_genes, int32_t tandem_gap, int device, uint8_t* remapped_out);

/* String interning on the device: dense ids for n byte strings (string i = blob[off[i] .. off[i+1]), off[0] = 0).
 * Equal strings get equal ids; ids number the distinct strings in order of first occurrence, so the result does not
 * depend on thread timing.  This is the hashing half of the reference's family join: DataSet::getGeneFamilies
 * (DataSet.cpp:188-202) maps every gene through the blast table's map<string, string> (GeneFamily.cpp:12-16, 38-48)
 * and GHM::buildMatrixFast hashes the family strings once more per list pair (GHM.cpp:34-50); a caller interns its
 * family strings (-> iadh_dataset_desc.fam) and gene IDs (-> .gid) once with this call, and K1 joins on integers.
 * No context needed; ids_out holds n entries. */
int iadh_intern_strings(const char* blob, const int64_t* off, int64_t n, int device, int32_t* ids_out,
                        int32_t* n_distinct_out);

/* Replaces: DataSet ownership of the flattened lists; called once after
 * DataSet::sortGeneLists() (iADHoRe.cpp:75).  `device` = CUDA ordinal. */
int iadh_create(const iadh_dataset_desc* ds, int device, iadh_ctx** out);
/* table_type=pairs (Settings.cpp `table_type`, DataSet::getGenePairs DataSet.cpp:144-186): the same
 * context for a dataset whose homology is a gene-pair relation instead of families.  Replaces the
 * slow path of GHM::buildMatrix (GHM.cpp:105-147) and the pairs branch of Gene::isPairWith
 * (Gene.cpp:24-27) inside GHMProfile::buildMatrix / setMultiplicons.  The pair sets are a CSR over
 * the dense gene ids of ds->gid: partners of gene g = pair_ids[pair_ptr[g] .. pair_ptr[g+1]), ascending,
 * symmetric, no self pairs -- i.e. Gene::pairs after getGenePairs.  ds->fam and the `fam` array of
 * iadh_profile_desc are ignored (may be NULL).  Every other entry point behaves as in family mode;
 * one difference inherited from the reference: the level-2 build honours masks in pairs mode
 * (ListElement::matchingPositions, ListElement.cpp:20-30), so iadh_level2_* refuses to run while
 * elements are masked (the reference never does that either: level 2 precedes all masking). */
int iadh_create_pairs(const iadh_dataset_desc* ds, const int64_t* pair_ptr, const int32_t* pair_ids, int32_t n_genes,
                      int device, iadh_ctx** out);
void iadh_destroy(iadh_ctx* ctx);

/* Replaces: the Settings getters read inside GHM::runCollinear (GHM.cpp:157-219). */
int iadh_set_params(iadh_ctx* ctx, const iadh_params* p);

/* Replaces: DataSet::level2ADHoRe over every list pair (levelTwo.cpp:32-96) in the
 * 1-thread order of getSomeWorkL2 (levelTwo.cpp:113-175) over lists sorted big->small
 * (DataSet.cpp:1088-1105).  rank/n_ranks shard the pair sequence by the greedy
 * least-loaded rule of getProcForPackage (levelTwo.cpp:98-111,142-148); the result holds
 * this rank's multiplicons in sequence order.  Use rank=0,n_ranks=1 for everything. */
int iadh_level2_all_pairs(iadh_ctx* ctx, int rank, int n_ranks, iadh_result** out);

/* Same work for an explicit pair list (x_list[i], y_list[i]); results in that order. */
int iadh_level2_pairs(iadh_ctx* ctx, const int32_t* x_list, const int32_t* y_list,
                      int64_t n_pairs, iadh_result** out);

/* Replaces: GeneList::mask (GeneList.cpp:283-287) as called at higherLevel.cpp:165-171. */
int iadh_mask_range(iadh_ctx* ctx, int32_t list, int32_t begin, int32_t end);
int iadh_clear_masks(iadh_ctx* ctx);

/* Replaces: DataSet::profileSearch over every list (higherLevel.cpp:32-48) for one
 * profile: lists big->small in the static partition of ParToolBox::statPartWorklSort
 * (parallel.cpp:96-156) for (rank, n_ranks); lists with fewer than anchor_points unmasked
 * elements are skipped (higherLevel.cpp:40-41).  Multiplicons are returned BEFORE the
 * homolog-count filter of GHMProfile.cpp:129-132, which needs the Link set of the profile's
 * multiplicon: iadh_profile_search_filtered (below) takes that set and applies the filter on
 * the device; with this entry point the caller applies it after unpacking. */
int iadh_profile_search(iadh_ctx* ctx, const iadh_profile_desc* profile, int rank, int n_ranks,
                        iadh_result** out);

/* Replaces: K consecutive iterations of the search half of DataSet::profileDetection (higherLevel.cpp:124-186:
 * mask the multiplicon's segments :165-171, then profileSearch :178) in one call -- the batched / speculative
 * path of SURVEY.md section 7 ("hard part 2").  Exactly equivalent to
 *     for k in 0 .. n_profiles-1:
 *         for m in mask_ptr[k] .. mask_ptr[k+1]-1:  iadh_mask_range(ctx, masks[m].list, masks[m].begin, masks[m].end)
 *         iadh_profile_search(ctx, &profiles[k], rank, n_ranks, &out[k])
 * (profile k sees the masks of sets 0..k and of earlier calls; afterwards all sets are applied to the context),
 * but every profile x list GHM of the batch is built and clustered in the same kernel launches, so the fixed
 * cost and the critical path of a search (one heavy GHM on one CTA) are paid once per batch, not per profile.
 * shard_profiles != 0: instead of sharding the lists of every profile over the ranks (parallel.cpp:96-156),
 * profile k is searched against ALL lists by rank k % n_ranks; the other ranks get an empty result for it.
 * A caller that speculates (profiles built before the results of earlier searches are known) keeps out[k] only
 * while the masks it assumed for k are the masks the reference's loop would have applied. */
typedef struct {
    int32_t list, begin, end;   /* GeneList::mask(begin, end) of list `list`, inclusive (GeneList.cpp:283-287) */
} iadh_mask_range_t;
int iadh_profile_search_batch(iadh_ctx* ctx, const iadh_profile_desc* profiles, int32_t n_profiles,
                              const iadh_mask_range_t* masks, const int32_t* mask_ptr, int rank, int n_ranks,
                              int shard_profiles, iadh_result** out);

/* Replaces: the homolog-count filter at the end of GHMProfile::setMultiplicons (GHMProfile.cpp:129-132), i.e. the return
 * value of Multiplicon::createHomologs(useFamily, minHomologs) (Multiplicon.cpp:352-382) with minHomologs =
 * Settings::getAnchorPoints() (GHM.cpp:218), evaluated on the device for every multiplicon of the search BEFORE the
 * stream is written: extractXObject / extractYObject (Multiplicon.cpp:331-350) slice the profile's segments to
 * [begin_x, end_x] and the list to [begin_y, end_y]; pruneLinks (:384-421) keeps the inherited Links whose two genes are
 * still inside the slices; addLinks (:423-494) adds the (x-segment element, y element) homologs; the multiplicon
 * survives iff every x-segment and the y-segment hold at least minHomologs elements with a homolog.  The result holds
 * exactly the multiplicons the reference's search returns (the `_filtered` entry points of the shim need no host-side
 * filter); multiplicons that fail never cross PCIe (4 in 10 on the config-3 shape).
 * links = profile->getMultiplicon().getHomologs() (Multiplicon.cpp:343-344: what extractXObject copies), any order;
 * segment indices are profile segments, gene ids are ListElement numIDs.  n_links must be > 0: a profile's multiplicon
 * always owns the Links of its anchor points (with an empty set the reference skips pruneLinks and keeps whatever
 * flags the copied elements had, Multiplicon.cpp:386-388 -- not reproducible from this interface). */
typedef struct {
    int32_t segment_x, segment_y;   /* Link::segmentX < Link::segmentY (Multiplicon.h:22-41) */
    int32_t gene_x_id, gene_y_id;   /* Link::geneXID, Link::geneYID                           */
} iadh_link_t;
int iadh_profile_search_filtered(iadh_ctx* ctx, const iadh_profile_desc* profile, const iadh_link_t* links, int64_t n_links,
                                 int rank, int n_ranks, iadh_result** out);
/* The batch with the same filter: links of profile k = links[link_ptr[k] .. link_ptr[k+1]). */
int iadh_profile_search_batch_filtered(iadh_ctx* ctx, const iadh_profile_desc* profiles, int32_t n_profiles,
                                       const iadh_link_t* links, const int64_t* link_ptr, const iadh_mask_range_t* masks,
                                       const int32_t* mask_ptr, int rank, int n_ranks, int shard_profiles, iadh_result** out);

/* Multi-GPU exchange, one process (rank) per GPU.  Replaces the MPI plumbing of the two drivers: ParToolBox
 * (parallel.cpp:13-31: rank / size) and the exchange step MPI_Allgather + MPI_Allgatherv of the packed
 * multiplicons (levelTwo.cpp:225-269, higherLevel.cpp:279-314), here one NCCL all-gather of sizes and one of
 * equal padded slots over NVLink.  iadh_comm_unique_id: rank 0 makes the id (128 bytes) and hands it to the other
 * ranks by whatever the host program has (MPI_Bcast, a file, torch.distributed); iadh_comm_init: every rank, the
 * same id.  iadh_allgather_results: for each i in [0, n), all[i] = the streams mine[i] of rank 0, 1, ..., n_ranks-1
 * back to back behind their summed multiplicon count -- what unpackL2Multiplicons / unpackHLMultiplicons read on
 * every rank after the reference's gather.  A single stream that is still on the device (the result of the call
 * just made) is sent from there; the gathered stream goes straight into the new result's pinned buffer.
 * iadh_result_ghm_index of a gathered result is empty (GHM positions are rank-local). */
int iadh_comm_unique_id(uint8_t* id_out, size_t n_bytes);
int iadh_comm_init(iadh_ctx* ctx, const uint8_t* id, size_t n_bytes, int rank, int n_ranks);
int iadh_comm_destroy(iadh_ctx* ctx);
int iadh_allgather_results(iadh_ctx* ctx, iadh_result* const* mine, int32_t n, iadh_result** all);

/* Packed stream, byte-for-byte Multiplicon::packMultiplicons (Multiplicon.cpp:242-264):
 * int32 nMultiplicons, then per multiplicon the 33-byte header, per base cluster the
 * 102-byte header (BaseCluster.cpp:751-797), per anchor point 17 bytes (AnchorPoint.h:103). */
const uint8_t* iadh_result_bytes(const iadh_result* r, size_t* n_bytes);
/* A call returns as soon as its packed stream is complete on the DEVICE; the copy into the result's pinned host
 * buffer runs on a copy stream of the context and overlaps whatever the caller does next (the next call's kernels
 * included).  iadh_result_bytes waits for that copy; iadh_result_size (bytes of the stream) never waits. */
size_t iadh_result_size(const iadh_result* r);
/* Per multiplicon: index of the GHM (position in the pair / list sequence) it came from. */
const int32_t* iadh_result_ghm_index(const iadh_result* r, size_t* n_multiplicons);
int iadh_result_stats(const iadh_result* r, iadh_stats* out);
void iadh_free_result(iadh_result* r);
/* Device and pinned-host blocks of destroyed contexts / freed results are kept in a process-wide
 * cache (contexts are cheap to re-create); iadh_trim returns them to the driver. */
void iadh_trim(void);

/* Creates the CUDA context of `device` (0.5-1 s in a fresh process).  Optional: a caller with host work
 * to do first (the reference reads its gene lists and family table for about that long) can run it
 * on a helper thread at start-up, so that iadh_create does not wait for it. */
int iadh_warmup(int device);

/* Device-side timing for benchmarks: CUDA events on the context's own stream (slots 0..3). */
int iadh_event_record(iadh_ctx* ctx, int slot);
int iadh_event_elapsed_ms(iadh_ctx* ctx, int from_slot, int to_slot, double* ms);

/* Test/measurement hook: the dot sets of the build kernels alone (GHM.cpp:34-85 /
 * GHMProfile.cpp:17-59) for explicit L2 pairs or, if profile != NULL, for profile x
 * y_list[i].  Writes per GHM counts cnt[2*i+o] and the concatenated (x,y) int32 pairs
 * (orientation 0 block then orientation 1 block per GHM, each sorted by x then y) into a
 * library-owned buffer valid until the next call on ctx. */
int iadh_build_dots(iadh_ctx* ctx, const iadh_profile_desc* profile, const int32_t* x_list,
                    const int32_t* y_list, int64_t n, const int64_t** cnt_out,
                    const int32_t** dots_out, int64_t* n_dots_out);

/* Measurement hook: the build kernels alone for the same arguments, dots left on the device.
 * out->n_dots, out->build_ms (count + scan + fill), out->cluster_ms = the fill pass alone. */
int iadh_build_only(iadh_ctx* ctx, const iadh_profile_desc* profile, const int32_t* x_list,
                    const int32_t* y_list, int64_t n, iadh_stats* out);

/* Host-only scheduling helpers (no GPU needed; what the two drivers above use internally).
 * sizes[l] = remapped length of list l.
 * iadh_plan_lists: DataSet::sortGeneLists order (DataSet.cpp:1088-1105) dealt to n_ranks
 *   packages as ParToolBox::statPartWorklSort does (parallel.cpp:96-156); writes this rank's
 *   lists (big -> small) to out[0..*n_out), out must hold n_lists entries.
 * iadh_plan_pairs: the pair walk of DataSet::getSomeWorkL2 (levelTwo.cpp:113-175) with the
 *   least-loaded-rank rule (levelTwo.cpp:98-111); out_x/out_y must hold n(n+1)/2 entries. */
int iadh_plan_lists(const int64_t* sizes, int32_t n_lists, int rank, int n_ranks, int32_t* out,
                    int64_t* n_out);
int iadh_plan_pairs(const int64_t* sizes, int32_t n_lists, int rank, int n_ranks, int32_t* out_x,
                    int32_t* out_y, int64_t* n_out);

/* hpmath::omxn (hpmath.cpp:18-36) evaluated on the device for n inputs (KAT hook). */
int iadh_omxn(iadh_ctx* ctx, const double* x, const int32_t* n, int64_t count, double* out);
/* pow(base, (double)n) as the reference's libm evaluates it inside
 * BaseCluster::calculateProbability (BaseCluster.cpp:567), on the device (KAT hook). */
int iadh_pow(iadh_ctx* ctx, const double* base, const int32_t* n, int64_t count, double* out);

const char* iadh_last_error(void);
const char* iadh_version(void);

#ifdef __cplusplus
}
#endif
#endif /* IADHORE_B200_H */
