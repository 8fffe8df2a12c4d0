/* In-memory record table served by the htslib shim as "mem:<name>".
 * TEST INFRASTRUCTURE ONLY (oracle side). */
#ifndef CSV_SHIM_MEM_H
#define CSV_SHIM_MEM_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct {
    int32_t n_targets;
    const char* const* target_name;
    const uint32_t* target_len;
    uint64_t n_reads;
    const int32_t* tid;        /* may be NULL: all contig 0 */
    const int32_t* pos0;
    const uint16_t* flag;
    const uint8_t* mapq;
    const uint64_t* cig_off;   /* [n_reads+1] */
    const uint32_t* cigar;
    const uint8_t* seq4;       /* optional 4-bit packed bases (BAM encoding), per-read byte ranges */
    const uint64_t* seq_off;   /* optional [n_reads] byte offset of each read in seq4 */
} csvshim_mem;
void csvshim_register_mem(const char* name, const csvshim_mem* m);
void csvshim_unregister_mem(const char* name);
#ifdef __cplusplus
}
#endif
#endif
