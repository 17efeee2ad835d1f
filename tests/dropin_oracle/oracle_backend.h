/* TEST INFRASTRUCTURE.  Force-included (-include) when integration/dropin_shim.cpp is compiled for the
 * CPU-only check of the shim: every iadh_* entry point the shim binds is redirected to the oracle's
 * oracle_* export of the same signature (oracle/oracle_capi.cpp), so the drop-in binary can run -- and
 * its output tables be compared with the reference's -- in a container without a GPU.  The product
 * binary (integration/Makefile) never sees this header. */
#define iadh_create oracle_create
#define iadh_create_pairs oracle_create_pairs
#define iadh_destroy oracle_destroy
#define iadh_set_params oracle_set_params
#define iadh_level2_all_pairs oracle_level2_all_pairs
#define iadh_level2_pairs oracle_level2_pairs
#define iadh_mask_range oracle_mask_range
#define iadh_clear_masks oracle_clear_masks
#define iadh_profile_search oracle_profile_search
#define iadh_profile_search_filtered oracle_profile_search_filtered
#define iadh_result_bytes oracle_result_bytes
#define iadh_result_ghm_index oracle_result_ghm_index
#define iadh_result_stats oracle_result_stats
#define iadh_free_result oracle_free_result
#define iadh_warmup oracle_warmup
#define iadh_intern_strings oracle_intern_strings
#define iadh_last_error oracle_last_error
