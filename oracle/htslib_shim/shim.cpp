/*
 * shim.cpp -- the handful of htslib entry points ContextSV links against,
 * re-implemented from the SAM/BAM/BGZF specification on top of zlib.
 *
 * TEST INFRASTRUCTURE ONLY (oracle side).  Our own code, not htslib.
 * Two record sources sit behind sam_open():
 *   "mem:<name>"  an in-process table registered with csvshim_register_mem()
 *                 (packed SoA records; lets the harness feed the unmodified
 *                 reference the very buffers the CUDA path receives);
 *   anything else a coordinate-sorted BAM file (BGZF blocks inflated with
 *                 zlib, records decoded per the BAM spec, CG:B,I long-CIGAR
 *                 promotion).  No .bai is needed: sam_index_load() scans the
 *                 file once and remembers where each contig starts.
 * The BCF side is inert ("no SNP file could be opened").
 */
#include "htslib/sam.h"
#include "htslib/synced_bcf_reader.h"
#include "shim_mem.h"

#include <zlib.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

extern "C" const char seq_nt16_str[] = "=ACMGRSVTWYHKDBN";

/* ------------------------------------------------------------ BGZF reader */
namespace {

/* One BGZF block: header parsed, payload read, inflated.  Used by the sequential reader and by the read-ahead threads. */
struct Block {
    int64_t coff = 0, next_coff = 0;
    std::vector<uint8_t> cbuf, ubuf;
    size_t clen = 0;
    bool eof = false, bad = false;

    /* reads the compressed block at coff (no inflate) */
    void fetch(FILE* f, int64_t at) {
        coff = at; eof = bad = false; ubuf.clear(); clen = 0;
        uint8_t h[18];
        if (fseeko(f, coff, SEEK_SET) != 0) { bad = true; return; }
        size_t got = fread(h, 1, 18, f);
        if (got == 0) { eof = true; return; }
        if (got < 18 || h[0] != 31 || h[1] != 139 || h[2] != 8 || !(h[3] & 4)) { bad = true; return; }
        unsigned xlen = h[10] | (h[11] << 8);
        std::vector<uint8_t> extra(xlen);
        memcpy(extra.data(), h + 12, xlen < 6 ? xlen : 6);
        if (xlen > 6 && fread(extra.data() + 6, 1, xlen - 6, f) != xlen - 6) { bad = true; return; }
        int bsize = -1;
        for (unsigned p = 0; p + 4 <= xlen;) {                 /* find the BC subfield */
            unsigned slen = extra[p + 2] | (extra[p + 3] << 8);
            if (extra[p] == 'B' && extra[p + 1] == 'C' && slen == 2) bsize = extra[p + 4] | (extra[p + 5] << 8);
            p += 4 + slen;
        }
        if (bsize < 0) { bad = true; return; }
        clen = (size_t)bsize + 1 - 12 - xlen - 8;              /* deflate payload */
        cbuf.resize(clen + 8);
        if (fseeko(f, coff + 12 + xlen, SEEK_SET) != 0 || fread(cbuf.data(), 1, clen + 8, f) != clen + 8) { bad = true; return; }
        next_coff = coff + bsize + 1;
    }
    void inflate_payload() {
        if (eof || bad) return;
        uint32_t isize; memcpy(&isize, cbuf.data() + clen + 4, 4);
        ubuf.resize(isize);
        if (!isize) return;
        z_stream zs; memset(&zs, 0, sizeof zs);
        if (inflateInit2(&zs, -15) != Z_OK) { bad = true; return; }
        zs.next_in = cbuf.data(); zs.avail_in = (uInt)clen;
        zs.next_out = ubuf.data(); zs.avail_out = isize;
        int rc = inflate(&zs, Z_FINISH);
        inflateEnd(&zs);
        if (rc != Z_STREAM_END) bad = true;
    }
};

/* hts_set_threads(fp, n > 1): one thread reads compressed blocks ahead of the consumer, n threads inflate them; the
 * consumer takes the blocks in file order.  Restarted by every seek. */
struct ReadAhead {
    enum { kRing = 256 };
    struct Slot { Block b; int state = 0; /* 0 free, 1 compressed, 2 inflating, 3 done */ };
    std::string path;
    FILE* f = nullptr;
    std::vector<Slot> ring = std::vector<Slot>(kRing);
    std::mutex m;
    std::condition_variable cv;
    uint64_t head = 0, tail = 0, work = 0;      /* consumer / reader / next slot to inflate */
    int64_t next_coff = 0;
    bool stop = false, ended = false;
    std::thread reader;
    std::vector<std::thread> workers;

    void start(const std::string& p, int64_t coff, int n_threads) {
        path = p; f = fopen(p.c_str(), "rb");
        if (!f) return;
        next_coff = coff; head = tail = work = 0; stop = ended = false;
        for (auto& s : ring) s.state = 0;
        reader = std::thread([this] {
            for (;;) {
                Slot* s;
                {
                    std::unique_lock<std::mutex> lk(m);
                    cv.wait(lk, [&] { return stop || tail - head < kRing; });
                    if (stop) return;
                    s = &ring[tail % kRing];
                }
                s->b.fetch(f, next_coff);
                const bool last = s->b.eof || s->b.bad;
                if (!last) next_coff = s->b.next_coff;
                {
                    std::lock_guard<std::mutex> lk(m);
                    s->state = 1; tail++;
                    if (last) ended = true;
                }
                cv.notify_all();
                if (last) return;
            }
        });
        for (int i = 0; i < n_threads; i++) workers.emplace_back([this] {
            for (;;) {
                Slot* s;
                {
                    std::unique_lock<std::mutex> lk(m);
                    cv.wait(lk, [&] { return stop || work < tail; });
                    if (stop) return;
                    s = &ring[work % kRing]; work++; s->state = 2;
                }
                s->b.inflate_payload();
                { std::lock_guard<std::mutex> lk(m); s->state = 3; }
                cv.notify_all();
            }
        });
    }
    /* next block in file order, swapped into `out`; false once the stream has ended and everything was handed over */
    bool next(Block& out) {
        std::unique_lock<std::mutex> lk(m);
        cv.wait(lk, [&] { return (head < tail && ring[head % kRing].state == 3) || (ended && head == tail); });
        if (head == tail) { out.eof = true; out.bad = false; out.ubuf.clear(); return false; }
        Slot& s = ring[head % kRing];
        std::swap(out, s.b);
        s.state = 0; head++;
        lk.unlock();
        cv.notify_all();
        return true;
    }
    void shutdown() {
        { std::lock_guard<std::mutex> lk(m); stop = true; }
        cv.notify_all();
        if (reader.joinable()) reader.join();
        for (auto& t : workers) if (t.joinable()) t.join();
        workers.clear();
        if (f) { fclose(f); f = nullptr; }
    }
    ~ReadAhead() { shutdown(); }
};

struct Bgzf {
    FILE* f = nullptr;
    std::string path;
    int threads = 0;                       /* hts_set_threads */
    std::unique_ptr<ReadAhead> ahead;      /* running read-ahead, positioned at next_coff */
    Block cur;
    std::vector<uint8_t>& ubuf = cur.ubuf;
    int64_t block_coff = 0;   /* file offset of the block in ubuf */
    int64_t next_coff = 0;    /* file offset of the next block     */
    size_t upos = 0;          /* read cursor inside ubuf           */
    bool eof = false;

    bool load_block() {
        if (threads > 1) {
            if (!ahead) { ahead.reset(new ReadAhead); ahead->start(path, next_coff, threads); }
            if (ahead->f) {
                ahead->next(cur);
                upos = 0;
                if (cur.eof) { eof = true; ubuf.clear(); return false; }
                if (cur.bad) return false;
                block_coff = cur.coff; next_coff = cur.next_coff;
                return true;
            }
        }
        cur.fetch(f, next_coff);
        upos = 0;
        if (cur.eof) { eof = true; ubuf.clear(); return false; }
        cur.inflate_payload();
        if (cur.bad) return false;
        block_coff = cur.coff; next_coff = cur.next_coff;
        return true;
    }
    /* read exactly n bytes; returns false at EOF/error */
    bool read(void* dst, size_t n) {
        uint8_t* d = (uint8_t*)dst;
        while (n) {
            if (upos == ubuf.size()) { if (!load_block()) return false; continue; }
            size_t k = ubuf.size() - upos; if (k > n) k = n;
            memcpy(d, ubuf.data() + upos, k); upos += k; d += k; n -= k;
        }
        return true;
    }
    uint64_t tell() {   /* virtual offset of the next byte */
        if (upos == ubuf.size()) return (uint64_t)next_coff << 16;
        return ((uint64_t)block_coff << 16) | (uint64_t)upos;
    }
    bool seek(uint64_t voff) {
        ahead.reset();                                     /* read-ahead restarts at the new offset */
        next_coff = (int64_t)(voff >> 16); eof = false;
        ubuf.clear(); upos = 0;
        size_t u = voff & 0xffff;
        if (u) { if (!load_block()) return false; upos = u; }
        return true;
    }
};

struct FileIndex { std::vector<uint64_t> first_voff; uint64_t records_voff = 0; };
std::mutex g_mu;
std::map<std::string, FileIndex> g_index_cache;
std::map<std::string, csvshim_mem> g_mem;

}  // namespace

struct htsFile {
    bool is_mem = false;
    std::string path;
    /* file source */
    Bgzf bz;
    uint64_t records_voff = 0;
    /* mem source */
    csvshim_mem mem{};
    /* header copy (both sources) */
    std::vector<std::string> names;
    std::vector<uint32_t> lens;
    std::string text;
};
struct hts_idx_t { FileIndex fi; bool is_mem = false; };
struct hts_itr_t {
    int tid; hts_pos_t beg, end; bool all;
    bool started = false;
    uint64_t mem_cursor = 0;
    uint64_t start_voff = 0;
};

/* --------------------------------------------------------- record decoding */
static void ensure_data(bam1_t* b, size_t n)
{
    if (n > b->m_data) {
        size_t m = n + (n >> 1) + 64;
        b->data = (uint8_t*)realloc(b->data, m); b->m_data = (uint32_t)m;
    }
}

/* ref length consumed by the CIGAR in b */
static hts_pos_t cigar_rlen(const uint32_t* cig, uint32_t n)
{
    hts_pos_t l = 0;
    for (uint32_t k = 0; k < n; k++) if (bam_cigar_type(bam_cigar_op(cig[k])) & 2) l += bam_cigar_oplen(cig[k]);
    return l;
}

/* returns 0 ok, -1 EOF, -2 error.  Reads one BAM record from the BGZF stream. */
static int read_file_record(htsFile* fp, bam1_t* b)
{
    uint32_t block_size;
    if (!fp->bz.read(&block_size, 4)) return fp->bz.eof ? -1 : -2;
    std::vector<uint8_t> raw(block_size);
    if (!fp->bz.read(raw.data(), block_size)) return -2;
    const uint8_t* p = raw.data();
    int32_t refID, pos, l_seq, next_refID, next_pos, tlen; uint16_t bin, n_cigar, flag; uint8_t l_read_name, mapq;
    memcpy(&refID, p, 4); memcpy(&pos, p + 4, 4); l_read_name = p[8]; mapq = p[9];
    memcpy(&bin, p + 10, 2); memcpy(&n_cigar, p + 12, 2); memcpy(&flag, p + 14, 2);
    memcpy(&l_seq, p + 16, 4); memcpy(&next_refID, p + 20, 4); memcpy(&next_pos, p + 24, 4); memcpy(&tlen, p + 28, 4);
    const uint8_t* name = p + 32;
    const uint8_t* cig = name + l_read_name;
    const uint8_t* seq = cig + 4u * n_cigar;
    size_t seq_bytes = ((size_t)l_seq + 1) / 2;
    const uint8_t* qual = seq + seq_bytes;
    const uint8_t* aux = qual + l_seq;
    const uint8_t* end = p + block_size;
    if (aux > end) return -2;

    /* CG:B,I long-CIGAR promotion (SAM spec section 4.2.2) */
    const uint8_t* real_cig = cig; uint32_t real_n = n_cigar;
    const uint8_t* cg_beg = nullptr; const uint8_t* cg_end = nullptr;
    if (n_cigar == 2) {
        uint32_t c0, c1; memcpy(&c0, cig, 4); memcpy(&c1, cig + 4, 4);
        if (bam_cigar_op(c0) == BAM_CSOFT_CLIP && (int32_t)bam_cigar_oplen(c0) == l_seq && bam_cigar_op(c1) == BAM_CREF_SKIP) {
            const uint8_t* a = aux;
            while (a + 3 <= end) {
                const uint8_t* tag = a; uint8_t ty = a[2]; const uint8_t* v = a + 3; size_t sz = 0;
                switch (ty) {
                    case 'A': case 'c': case 'C': sz = 1; break;
                    case 's': case 'S': sz = 2; break;
                    case 'i': case 'I': case 'f': sz = 4; break;
                    case 'Z': case 'H': sz = strlen((const char*)v) + 1; break;
                    case 'B': {
                        uint8_t sub = v[0]; uint32_t cnt; memcpy(&cnt, v + 1, 4);
                        size_t es = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : 4;
                        sz = 5 + es * cnt;
                        if (tag[0] == 'C' && tag[1] == 'G' && sub == 'I') { real_cig = v + 5; real_n = cnt; cg_beg = tag; cg_end = v + sz; }
                        break;
                    }
                    default: a = end; continue;
                }
                a = v + sz;
            }
        }
    }
    size_t l_qname = l_read_name; unsigned extranul = 0;
    while ((l_qname + extranul) % 4) extranul++;
    size_t aux_len = (size_t)(end - aux) - (cg_beg ? (size_t)(cg_end - cg_beg) : 0);
    size_t total = l_qname + extranul + 4u * real_n + seq_bytes + l_seq + aux_len;
    ensure_data(b, total);
    uint8_t* d = b->data;
    memcpy(d, name, l_qname); memset(d + l_qname, 0, extranul); d += l_qname + extranul;
    memcpy(d, real_cig, 4u * real_n); d += 4u * real_n;
    memcpy(d, seq, seq_bytes); d += seq_bytes;
    memcpy(d, qual, l_seq); d += l_seq;
    if (cg_beg) { memcpy(d, aux, cg_beg - aux); d += cg_beg - aux; memcpy(d, cg_end, end - cg_end); }
    else memcpy(d, aux, aux_len);
    b->l_data = (int)total;
    b->core.pos = pos; b->core.tid = refID; b->core.bin = bin; b->core.qual = mapq;
    b->core.l_extranul = (uint8_t)extranul; b->core.flag = flag; b->core.l_qname = (uint16_t)(l_qname + extranul);
    b->core.n_cigar = real_n; b->core.l_qseq = l_seq; b->core.mtid = next_refID; b->core.mpos = next_pos; b->core.isize = tlen;
    return 0;
}

static void fill_mem_record(const csvshim_mem& m, uint64_t i, bam1_t* b)
{
    char qn[32]; int ql = snprintf(qn, sizeof qn, "r%llu", (unsigned long long)i) + 1;
    unsigned extranul = 0; while ((ql + extranul) % 4) extranul++;
    uint64_t c0 = m.cig_off[i], c1 = m.cig_off[i + 1];
    uint32_t n_cigar = (uint32_t)(c1 - c0);
    /* query length implied by the CIGAR */
    int64_t qlen = 0;
    for (uint64_t k = c0; k < c1; k++) if (bam_cigar_type(bam_cigar_op(m.cigar[k])) & 1) qlen += bam_cigar_oplen(m.cigar[k]);
    size_t seq_bytes = ((size_t)qlen + 1) / 2;
    size_t total = ql + extranul + 4u * n_cigar + seq_bytes + (size_t)qlen;
    ensure_data(b, total);
    uint8_t* d = b->data;
    memcpy(d, qn, ql); memset(d + ql, 0, extranul); d += ql + extranul;
    memcpy(d, m.cigar + c0, 4u * n_cigar); d += 4u * n_cigar;
    if (m.seq4 && m.seq_off) {
        uint64_t s0 = m.seq_off[i];   /* byte offset of this read's packed bases */
        memcpy(d, m.seq4 + s0, seq_bytes);
    } else {
        /* deterministic filler: base code depends on (read, position) */
        for (size_t k = 0; k < seq_bytes; k++) {
            static const uint8_t code[4] = {1, 2, 4, 8};
            uint64_t h = (i * 0x9E3779B97F4A7C15ull) ^ (k * 0xC2B2AE3D27D4EB4Full);
            d[k] = (uint8_t)((code[(h >> 20) & 3] << 4) | code[(h >> 40) & 3]);
        }
    }
    d += seq_bytes;
    memset(d, 0xff, (size_t)qlen);
    b->l_data = (int)total;
    b->core.pos = m.pos0[i]; b->core.tid = m.tid ? m.tid[i] : 0; b->core.bin = 0; b->core.qual = m.mapq[i];
    b->core.l_extranul = (uint8_t)extranul; b->core.flag = m.flag[i]; b->core.l_qname = (uint16_t)(ql + extranul);
    b->core.n_cigar = n_cigar; b->core.l_qseq = (int32_t)qlen; b->core.mtid = -1; b->core.mpos = -1; b->core.isize = 0;
}

/* ------------------------------------------------------------------- API */
extern "C" {

void csvshim_register_mem(const char* name, const csvshim_mem* m)
{
    std::lock_guard<std::mutex> lk(g_mu);
    g_mem[name] = *m;
}
void csvshim_unregister_mem(const char* name)
{
    std::lock_guard<std::mutex> lk(g_mu);
    g_mem.erase(name);
}

samFile* sam_open(const char* fn, const char* mode)
{
    (void)mode;
    htsFile* fp = new htsFile;
    fp->path = fn;
    if (strncmp(fn, "mem:", 4) == 0) {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_mem.find(fn + 4);
        if (it == g_mem.end()) { delete fp; return nullptr; }
        fp->is_mem = true; fp->mem = it->second;
        for (int t = 0; t < fp->mem.n_targets; t++) { fp->names.push_back(fp->mem.target_name[t]); fp->lens.push_back(fp->mem.target_len[t]); }
        return fp;
    }
    fp->bz.f = fopen(fn, "rb");
    if (!fp->bz.f) { delete fp; return nullptr; }
    fp->bz.path = fn;
    return fp;
}

int sam_close(samFile* fp)
{
    if (!fp) return -1;
    fp->bz.ahead.reset();
    if (fp->bz.f) fclose(fp->bz.f);
    delete fp;
    return 0;
}

int hts_set_threads(htsFile* fp, int n)
{
    if (fp && !fp->is_mem) { fp->bz.ahead.reset(); fp->bz.threads = n; }     /* n > 1: BGZF blocks are inflated ahead of the reader by n threads */
    return 0;
}
const char* hts_get_fn(htsFile* fp) { return fp ? fp->path.c_str() : nullptr; }

sam_hdr_t* sam_hdr_read(samFile* fp)
{
    if (!fp) return nullptr;
    if (!fp->is_mem) {
        fp->bz.seek(0);
        char magic[4]; int32_t l_text, n_ref;
        if (!fp->bz.read(magic, 4) || memcmp(magic, "BAM\1", 4) != 0) return nullptr;
        if (!fp->bz.read(&l_text, 4)) return nullptr;
        fp->text.resize(l_text);
        if (l_text && !fp->bz.read(&fp->text[0], l_text)) return nullptr;
        if (!fp->bz.read(&n_ref, 4)) return nullptr;
        fp->names.clear(); fp->lens.clear();
        for (int i = 0; i < n_ref; i++) {
            int32_t l_name; uint32_t l_ref;
            if (!fp->bz.read(&l_name, 4)) return nullptr;
            std::string nm(l_name, '\0');
            if (!fp->bz.read(&nm[0], l_name)) return nullptr;
            nm.resize(strlen(nm.c_str()));
            if (!fp->bz.read(&l_ref, 4)) return nullptr;
            fp->names.push_back(nm); fp->lens.push_back(l_ref);
        }
        fp->records_voff = fp->bz.tell();
    }
    sam_hdr_t* h = (sam_hdr_t*)calloc(1, sizeof(sam_hdr_t));
    h->n_targets = (int32_t)fp->names.size();
    h->target_len = (uint32_t*)calloc(h->n_targets + 1, sizeof(uint32_t));
    h->target_name = (char**)calloc(h->n_targets + 1, sizeof(char*));
    for (int i = 0; i < h->n_targets; i++) { h->target_len[i] = fp->lens[i]; h->target_name[i] = strdup(fp->names[i].c_str()); }
    h->l_text = fp->text.size(); h->text = strdup(fp->text.c_str());
    return h;
}

void sam_hdr_destroy(sam_hdr_t* h)
{
    if (!h) return;
    for (int i = 0; i < h->n_targets; i++) free(h->target_name[i]);
    free(h->target_name); free(h->target_len); free(h->text); free(h);
}

int sam_hdr_name2tid(sam_hdr_t* h, const char* ref)
{
    if (!h) return -2;
    for (int i = 0; i < h->n_targets; i++) if (strcmp(h->target_name[i], ref) == 0) return i;
    return -1;
}
int bam_name2id(sam_hdr_t* h, const char* ref) { return sam_hdr_name2tid(h, ref); }

hts_idx_t* sam_index_load(samFile* fp, const char* fn)
{
    (void)fn;
    if (!fp) return nullptr;
    hts_idx_t* idx = new hts_idx_t;
    if (fp->is_mem) { idx->is_mem = true; return idx; }
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_index_cache.find(fp->path);
        if (it != g_index_cache.end()) { idx->fi = it->second; return idx; }
    }
    /* one linear scan: remember the virtual offset of the first record of each contig */
    htsFile* scan = sam_open(fp->path.c_str(), "r");
    if (!scan) { delete idx; return nullptr; }
    {   /* a .bai would make this scan unnecessary: keep it short, inflate with every core (at most 16) */
        unsigned hw = std::thread::hardware_concurrency();
        hts_set_threads(scan, (int)(hw > 16 ? 16 : (hw < 2 ? 0 : hw)));
    }
    sam_hdr_t* h = sam_hdr_read(scan);
    if (!h) { sam_close(scan); delete idx; return nullptr; }
    FileIndex fi; fi.first_voff.assign(h->n_targets + 1, UINT64_MAX); fi.records_voff = scan->records_voff;
    for (;;) {
        uint64_t voff = scan->bz.tell();
        uint32_t block_size; int32_t refID;
        if (!scan->bz.read(&block_size, 4)) break;
        if (!scan->bz.read(&refID, 4)) break;
        int slot = (refID >= 0 && refID < h->n_targets) ? refID : h->n_targets;
        if (fi.first_voff[slot] == UINT64_MAX) fi.first_voff[slot] = voff;
        /* skip the rest of the record */
        size_t left = block_size - 4; uint8_t tmp[4096];
        bool ok = true;
        while (left) { size_t k = left < sizeof tmp ? left : sizeof tmp; if (!scan->bz.read(tmp, k)) { ok = false; break; } left -= k; }
        if (!ok) break;
    }
    sam_hdr_destroy(h); sam_close(scan);
    {
        std::lock_guard<std::mutex> lk(g_mu);
        g_index_cache[fp->path] = fi;
    }
    idx->fi = fi;
    return idx;
}

void hts_idx_destroy(hts_idx_t* idx) { delete idx; }
void hts_itr_destroy(hts_itr_t* itr) { delete itr; }

hts_itr_t* sam_itr_queryi(const hts_idx_t* idx, int tid, hts_pos_t beg, hts_pos_t end)
{
    if (!idx) return nullptr;
    hts_itr_t* it = new hts_itr_t;
    it->tid = tid; it->beg = beg; it->end = end; it->all = (tid == HTS_IDX_START);
    if (!idx->is_mem) {
        if (it->all) it->start_voff = idx->fi.records_voff;
        else if (tid >= 0 && (size_t)tid < idx->fi.first_voff.size()) it->start_voff = idx->fi.first_voff[tid];
        else it->start_voff = UINT64_MAX;
    }
    return it;
}

hts_itr_t* sam_itr_querys(const hts_idx_t* idx, sam_hdr_t* hdr, const char* region)
{
    if (!idx || !hdr) return nullptr;
    int tid = sam_hdr_name2tid(hdr, region);
    hts_pos_t beg = 0, end = INT64_MAX;
    if (tid < 0) {
        /* name:beg-end (1-based inclusive) */
        std::string r(region); size_t c = r.rfind(':');
        if (c == std::string::npos) return nullptr;
        tid = sam_hdr_name2tid(hdr, r.substr(0, c).c_str());
        if (tid < 0) return nullptr;
        std::string rng = r.substr(c + 1); std::string clean;
        for (char ch : rng) if (ch != ',') clean += ch;
        size_t d = clean.find('-');
        beg = atoll(clean.substr(0, d).c_str()) - 1; if (beg < 0) beg = 0;
        if (d != std::string::npos && d + 1 < clean.size()) end = atoll(clean.substr(d + 1).c_str());
    }
    return sam_itr_queryi(idx, tid, beg, end);
}

int sam_itr_next(samFile* fp, hts_itr_t* itr, bam1_t* r)
{
    if (!fp || !itr) return -2;
    if (fp->is_mem) {
        const csvshim_mem& m = fp->mem;
        while (itr->mem_cursor < m.n_reads) {
            uint64_t i = itr->mem_cursor++;
            int tid = m.tid ? m.tid[i] : 0;
            if (!itr->all) {
                if (tid != itr->tid) continue;
                hts_pos_t p = m.pos0[i];
                if (p >= itr->end) continue;
                if (itr->beg > 0) {
                    hts_pos_t e = p + cigar_rlen(m.cigar + m.cig_off[i], (uint32_t)(m.cig_off[i + 1] - m.cig_off[i]));
                    if (e <= p) e = p + 1;
                    if (e <= itr->beg) continue;
                }
            }
            fill_mem_record(m, i, r);
            return 0;
        }
        return -1;
    }
    if (!itr->started) {
        itr->started = true;
        if (itr->start_voff == UINT64_MAX) return -1;
        if (!fp->bz.seek(itr->start_voff)) return -2;
    }
    for (;;) {
        int rc = read_file_record(fp, r);
        if (rc < 0) return rc;
        if (itr->all) return 0;
        if (r->core.tid != itr->tid) return -1;          /* coordinate-sorted: the contig is over */
        if (r->core.pos >= itr->end) return -1;
        if (itr->beg > 0 && bam_endpos(r) <= itr->beg) continue;
        return 0;
    }
}

bam1_t* bam_init1(void) { return (bam1_t*)calloc(1, sizeof(bam1_t)); }
void bam_destroy1(bam1_t* b) { if (b) { free(b->data); free(b); } }

hts_pos_t bam_endpos(const bam1_t* b)
{
    hts_pos_t rlen = (b->core.flag & BAM_FUNMAP) ? 0 : cigar_rlen(bam_get_cigar(b), b->core.n_cigar);
    if (rlen == 0) rlen = 1;
    return b->core.pos + rlen;
}

/* ------------------------------------------------ inert BCF side ("no SNPs") */
bcf_srs_t* bcf_sr_init(void) { return (bcf_srs_t*)calloc(1, sizeof(bcf_srs_t)); }
void bcf_sr_destroy(bcf_srs_t* r) { free(r); }
int bcf_sr_set_threads(bcf_srs_t*, int) { return 0; }
int bcf_sr_add_reader(bcf_srs_t*, const char*) { return -1; }
int bcf_sr_set_regions(bcf_srs_t*, const char*, int) { return -1; }
int bcf_sr_next_line(bcf_srs_t*) { return 0; }
int bcf_sr_has_line(bcf_srs_t*, int) { return 0; }
bcf1_t* bcf_sr_get_line(bcf_srs_t*, int) { return nullptr; }
const char* bcf_sr_strerror(int) { return "htslib shim: BCF readers are not implemented"; }
int bcf_is_snp(bcf1_t*) { return 0; }
int bcf_has_filter(const bcf_hdr_t*, bcf1_t*, char*) { return 0; }
int bcf_get_format_values(const bcf_hdr_t*, bcf1_t*, const char*, void**, int*, int) { return -1; }
int bcf_get_info_values(const bcf_hdr_t*, bcf1_t*, const char*, void**, int*, int) { return -1; }

}  // extern "C"
