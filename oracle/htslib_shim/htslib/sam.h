/* Minimal htslib-compatible SAM/BAM declarations (our own code; see hts.h). */
#ifndef CSV_SHIM_SAM_H
#define CSV_SHIM_SAM_H
#include "hts.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef htsFile samFile;

typedef struct sam_hdr_t {
    int32_t n_targets;
    uint32_t* target_len;
    char** target_name;
    size_t l_text;
    char* text;
} sam_hdr_t;
typedef sam_hdr_t bam_hdr_t;

/* CIGAR op codes and FLAG bits: values fixed by the SAM specification */
#define BAM_CMATCH 0
#define BAM_CINS 1
#define BAM_CDEL 2
#define BAM_CREF_SKIP 3
#define BAM_CSOFT_CLIP 4
#define BAM_CHARD_CLIP 5
#define BAM_CPAD 6
#define BAM_CEQUAL 7
#define BAM_CDIFF 8
#define BAM_CBACK 9
#define BAM_CIGAR_SHIFT 4
#define BAM_CIGAR_MASK 0xf
#define BAM_CIGAR_TYPE 0x3C1A7
#define bam_cigar_op(c) ((c) & BAM_CIGAR_MASK)
#define bam_cigar_oplen(c) ((c) >> BAM_CIGAR_SHIFT)
#define bam_cigar_type(o) (BAM_CIGAR_TYPE >> ((o) << 1) & 3)

#define BAM_FPAIRED 1
#define BAM_FPROPER_PAIR 2
#define BAM_FUNMAP 4
#define BAM_FMUNMAP 8
#define BAM_FREVERSE 16
#define BAM_FMREVERSE 32
#define BAM_FREAD1 64
#define BAM_FREAD2 128
#define BAM_FSECONDARY 256
#define BAM_FQCFAIL 512
#define BAM_FDUP 1024
#define BAM_FSUPPLEMENTARY 2048

typedef struct bam1_core_t {
    hts_pos_t pos;
    int32_t tid;
    uint16_t bin;
    uint8_t qual;
    uint8_t l_extranul;
    uint16_t flag;
    uint16_t l_qname;
    uint32_t n_cigar;
    int32_t l_qseq;
    int32_t mtid;
    hts_pos_t mpos;
    hts_pos_t isize;
} bam1_core_t;

typedef struct bam1_t {
    bam1_core_t core;
    uint64_t id;
    uint8_t* data;
    int l_data;
    uint32_t m_data;
} bam1_t;

#define bam_get_qname(b) ((char*)(b)->data)
#define bam_get_cigar(b) ((uint32_t*)((b)->data + (b)->core.l_qname))
#define bam_get_seq(b) ((b)->data + ((b)->core.n_cigar << 2) + (b)->core.l_qname)
#define bam_get_qual(b) ((b)->data + ((b)->core.n_cigar << 2) + (b)->core.l_qname + (((b)->core.l_qseq + 1) >> 1))
#define bam_seqi(s, i) ((s)[(i) >> 1] >> ((~(i) & 1) << 2) & 0xf)

extern const char seq_nt16_str[];

samFile* sam_open(const char* fn, const char* mode);
int sam_close(samFile* fp);
sam_hdr_t* sam_hdr_read(samFile* fp);
void sam_hdr_destroy(sam_hdr_t* h);
#define bam_hdr_destroy(h) sam_hdr_destroy(h)
int sam_hdr_name2tid(sam_hdr_t* h, const char* ref);
int bam_name2id(sam_hdr_t* h, const char* ref);
hts_idx_t* sam_index_load(samFile* fp, const char* fn);
hts_itr_t* sam_itr_querys(const hts_idx_t* idx, sam_hdr_t* hdr, const char* region);
hts_itr_t* sam_itr_queryi(const hts_idx_t* idx, int tid, hts_pos_t beg, hts_pos_t end);
int sam_itr_next(samFile* fp, hts_itr_t* itr, bam1_t* r);
bam1_t* bam_init1(void);
void bam_destroy1(bam1_t* b);
hts_pos_t bam_endpos(const bam1_t* b);

#ifdef __cplusplus
}
#endif
#endif
