// Device-side records of the homolog-count filter (homolog_kernel.cu), shared with the engine.
#pragma once
#include <stdint.h>

// One Link of a profile's multiplicon (Multiplicon.h:22-52) as the caller hands it over: layout of iadh_link_t
// (include/iadhore_b200.h), uploaded as is.
struct LinkIn {
    int32_t seg_x, seg_y, gene_x, gene_y;
};

// Per profile: CSR over its S * L elements (segment-major, like ProfileDev::num_id) -> columns of the other ends of the
// element's inherited Links.  lk_ptr points at the profile's first element; its values index lk_col.
struct HomologDev {
    const int32_t* lk_ptr;
    const int32_t* lk_col;
};
