/* Minimal htslib-compatible VCF/BCF declarations (our own code; see hts.h).
 * The BCF side is inert in the shim: readers never open, so the reference's
 * CNV prediction sees "no SNPs" -- that part is outside the hot path. */
#ifndef CSV_SHIM_VCF_H
#define CSV_SHIM_VCF_H
#include "hts.h"
#include <math.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct bcf_hdr_t bcf_hdr_t;   /* opaque */
typedef struct bcf1_t {
    hts_pos_t pos;
    hts_pos_t rlen;
    int32_t rid;
    float qual;
} bcf1_t;

#define BCF_HT_INT 1
#define BCF_HT_REAL 2

static inline int bcf_float_is_missing(float f) { return isnan(f); }
int bcf_is_snp(bcf1_t* v);
int bcf_has_filter(const bcf_hdr_t* hdr, bcf1_t* line, char* filter);
int bcf_get_format_values(const bcf_hdr_t* hdr, bcf1_t* line, const char* tag, void** dst, int* ndst, int type);
int bcf_get_info_values(const bcf_hdr_t* hdr, bcf1_t* line, const char* tag, void** dst, int* ndst, int type);
#define bcf_get_format_int32(hdr, line, tag, dst, ndst) bcf_get_format_values(hdr, line, tag, (void**)(dst), ndst, BCF_HT_INT)
#define bcf_get_info_float(hdr, line, tag, dst, ndst) bcf_get_info_values(hdr, line, tag, (void**)(dst), ndst, BCF_HT_REAL)

#ifdef __cplusplus
}
#endif
#endif
