/* Minimal htslib-compatible synced-reader declarations (our own code; see vcf.h). */
#ifndef CSV_SHIM_SYNCED_BCF_READER_H
#define CSV_SHIM_SYNCED_BCF_READER_H
#include "vcf.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct bcf_sr_t {
    bcf_hdr_t* header;
} bcf_sr_t;

typedef struct bcf_srs_t {
    int require_index;
    int errnum;
    bcf_sr_t* readers;
    int nreaders;
} bcf_srs_t;

bcf_srs_t* bcf_sr_init(void);
void bcf_sr_destroy(bcf_srs_t* r);
int bcf_sr_set_threads(bcf_srs_t* r, int n);
int bcf_sr_add_reader(bcf_srs_t* r, const char* fname);
int bcf_sr_set_regions(bcf_srs_t* r, const char* regions, int is_file);
int bcf_sr_next_line(bcf_srs_t* r);
int bcf_sr_has_line(bcf_srs_t* r, int i);
bcf1_t* bcf_sr_get_line(bcf_srs_t* r, int i);
const char* bcf_sr_strerror(int errnum);

#ifdef __cplusplus
}
#endif
#endif
