/*
 * Minimal htslib-compatible declarations (our own code, NOT htslib).
 * TEST INFRASTRUCTURE ONLY: lets the unmodified ContextSV sources compile and
 * link in a container that has no htslib.  Only the names
 * the reference touches are declared; semantics are implemented in
 * ../shim.cpp (BGZF through zlib, BAM records, linear region scan).
 */
#ifndef CSV_SHIM_HTS_H
#define CSV_SHIM_HTS_H
#include <stdint.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef int64_t hts_pos_t;
#define HTS_IDX_NOCOOR (-2)
#define HTS_IDX_START  (-3)
#define HTS_IDX_REST   (-4)
#define HTS_IDX_NONE   (-5)

typedef struct htsFile htsFile;       /* opaque: defined in shim.cpp */
typedef struct hts_idx_t hts_idx_t;   /* opaque */
typedef struct hts_itr_t hts_itr_t;   /* opaque */

#define CSVSHIM_HTSLIB 1               /* lets code written for htslib know its htsFile is opaque here */
const char* hts_get_fn(htsFile* fp);   /* path the file was opened with (htsFile::fn in htslib) */
int hts_set_threads(htsFile* fp, int n);
void hts_idx_destroy(hts_idx_t* idx);
void hts_itr_destroy(hts_itr_t* itr);

#ifdef __cplusplus
}
#endif
#endif
