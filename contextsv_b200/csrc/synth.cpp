// synth.cpp -- seeded synthetic long-read alignments, emitted directly as the
// packed SoA the hot path consumes (no network, no
// htslib, no real BAMs in the container).
//
// Host-only C++ (no CUDA).  Deterministic for a given (seed, parameters) and
// independent of the thread count: every read draws from its own counter-based
// stream.  Reads are coordinate-sorted by construction, like the indexed BAM
// the reference requires (cnv_caller.cpp:439, sv_caller.cpp:712).
#include "contextsv_b200.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

namespace {

struct Rng {   // splitmix64 stream keyed by (seed, stream id)
    uint64_t s;
    Rng(uint64_t seed, uint64_t stream) : s(seed ^ (stream * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull)) { next(); }
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }   // [0,1)
    uint64_t below(uint64_t n) { return n ? (uint64_t)(uni() * (double)n) : 0; }
    double normal() {
        double u1 = uni(), u2 = uni();
        if (u1 < 1e-300) u1 = 1e-300;
        return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
    }
    // geometric gap with per-base event rate p (>= 1)
    uint32_t gap(double p) {
        double u = uni(); if (u < 1e-300) u = 1e-300;
        double g = std::floor(std::log(u) / std::log1p(-p)) + 1.0;
        return g > 4e9 ? 4000000000u : (uint32_t)g;
    }
};

struct SV { uint32_t pos; uint32_t len; uint8_t is_ins; uint8_t het; };

struct Plan {
    csv_synth_params p;
    std::vector<uint32_t> contig_len;
    std::vector<uint64_t> read_base;      // first read index of each contig
    std::vector<std::vector<SV>> svs;     // per contig, sorted by pos
    double mean_len;
};

double profile_mean_len(const csv_synth_params& p)
{
    if (p.profile == 1) {   // ONT: lognormal, N50 = n50 => mu chosen so that length-weighted median = n50
        double sigma = 0.9;
        double mu = std::log(p.read_len_mean) - sigma * sigma;   // N50 = exp(mu + sigma^2)
        return std::exp(mu + 0.5 * sigma * sigma);
    }
    return p.read_len_mean;
}

uint32_t draw_len(const csv_synth_params& p, Rng& g)
{
    if (p.profile == 1) {
        double sigma = 0.9, mu = std::log(p.read_len_mean) - sigma * sigma;
        double l = std::exp(mu + sigma * g.normal());
        if (l < 500) l = 500; if (l > 1e6) l = 1e6;
        return (uint32_t)l;
    }
    double l = p.read_len_mean + p.read_len_sd * g.normal();
    double lo = p.read_len_mean / 3.0, hi = p.read_len_mean * 5.0 / 3.0;
    if (l < lo) l = lo; if (l > hi) l = hi;
    return (uint32_t)l;
}

Plan make_plan(const csv_synth_params* p, uint32_t n_contigs, const uint32_t* contig_len)
{
    Plan pl; pl.p = *p;
    pl.contig_len.assign(contig_len, contig_len + n_contigs);
    pl.mean_len = profile_mean_len(*p);
    pl.read_base.resize(n_contigs + 1);
    uint64_t acc = 0; uint64_t total_len = 0;
    for (uint32_t c = 0; c < n_contigs; c++) {
        pl.read_base[c] = acc;
        acc += (uint64_t)std::llround(p->coverage * (double)contig_len[c] / pl.mean_len);
        total_len += contig_len[c];
    }
    pl.read_base[n_contigs] = acc;
    // structural variants: n_sv spread over contigs in proportion to length
    pl.svs.resize(n_contigs);
    for (uint32_t c = 0; c < n_contigs; c++) {
        uint64_t n = total_len ? (uint64_t)std::llround((double)p->n_sv * (double)contig_len[c] / (double)total_len) : 0;
        Rng g(p->seed ^ 0x5356ull, 0x1000000ull + c);
        std::vector<SV>& v = pl.svs[c];
        uint32_t L = contig_len[c];
        for (uint64_t k = 0; k < n && L > 40000; k++) {
            SV s;
            s.pos = 10000 + (uint32_t)g.below(L - 30000);
            double lmin = std::log(50.0), lmax = std::log((double)(p->sv_len_max > 50 ? p->sv_len_max : 50));
            s.len = (uint32_t)std::exp(lmin + (lmax - lmin) * g.uni());
            if (s.len < 50) s.len = 50;
            if (g.uni() < p->frac_len50) s.len = 50;   // exercises the literal-ALT branch (sv_caller.cpp:587-591)
            s.is_ins = g.uni() < 0.5; s.het = g.uni() < 0.5;
            v.push_back(s);
        }
        std::sort(v.begin(), v.end(), [](const SV& a, const SV& b) { return a.pos < b.pos; });
        // keep SVs apart so that one read never sees overlapping events
        std::vector<SV> w; uint32_t last_end = 0;
        for (const SV& s : v) { if (s.pos > last_end + 200) { w.push_back(s); last_end = s.pos + (s.is_ins ? 0 : s.len); } }
        v.swap(w);
    }
    return pl;
}

inline uint32_t cig(uint32_t len, uint32_t op) { return (len << 4) | op; }

// Emit the alignment of read `ridx` on contig c.  If out == nullptr only counts ops.
struct ReadOut { int32_t pos0; uint16_t flag; uint8_t mapq; uint32_t n_ops; };

ReadOut gen_read(const Plan& pl, uint32_t c, uint64_t ridx, uint32_t* out)
{
    const csv_synth_params& p = pl.p;
    uint64_t k = ridx - pl.read_base[c], n = pl.read_base[c + 1] - pl.read_base[c];
    uint32_t L = pl.contig_len[c];
    Rng g(p.seed, ridx);
    ReadOut r;
    // stratified start: sorted by construction
    double slot = (double)L / (double)n;
    uint64_t p0 = (uint64_t)((double)k * slot) + g.below((uint64_t)std::max(1.0, slot));
    uint64_t pnext = (uint64_t)((double)(k + 1) * slot);
    if (k + 1 < n && p0 >= pnext && pnext > 0) p0 = pnext - 1;
    if (p0 >= L) p0 = L - 1;
    r.pos0 = (int32_t)p0;
    // flags / MAPQ mix
    uint16_t flag = (g.uni() < 0.5) ? 16 : 0;
    double u = g.uni();
    if (u < p.frac_supplementary) flag |= 0x800;
    else if (u < p.frac_supplementary + p.frac_secondary) flag |= 0x100;
    else if (u < p.frac_supplementary + p.frac_secondary + p.frac_dup) flag |= 0x400;
    else if (u < p.frac_supplementary + p.frac_secondary + p.frac_dup + p.frac_qcfail) flag |= 0x200;
    r.flag = flag;
    r.mapq = (g.uni() < p.frac_lowmapq) ? (uint8_t)g.below(20) : (uint8_t)(20 + g.below(41));
    uint32_t want = draw_len(p, g);
    bool hap = g.uni() < 0.5;   // which haplotype the read comes from (het SVs on hap 1 only)
    uint32_t nops = 0;
    const uint32_t mop = p.use_eqx ? 7u : 0u;   // '=' instead of 'M'
    auto emit = [&](uint32_t len, uint32_t op) {
        if (len == 0) return;
        if (out) out[nops] = cig(len, op);
        nops++;
    };
    // leading soft clip
    double uc = g.uni();
    if (uc < p.frac_softclip * 0.5) emit((g.uni() < 0.1) ? 50 : 50 + (uint32_t)g.below(2000), 4);
    else if (uc < p.frac_softclip * 0.5 + 0.01) emit(1 + (uint32_t)g.below(40), 4);
    // walk the reference span
    uint64_t ref = p0, ref_end = std::min<uint64_t>((uint64_t)p0 + want, L);
    const std::vector<SV>& svs = pl.svs[c];
    size_t si = std::lower_bound(svs.begin(), svs.end(), (uint32_t)(p0 + 1), [](const SV& a, uint32_t x) { return a.pos < x; }) - svs.begin();
    uint32_t m_run = 0;
    while (ref < ref_end) {
        uint64_t next_small = p.indel_rate > 0 ? ref + g.gap(p.indel_rate) : UINT64_MAX;
        uint64_t next_sv = UINT64_MAX;
        if (si < svs.size()) {
            int64_t jit = p.sv_jitter_sd > 0 ? (int64_t)std::llround(p.sv_jitter_sd * g.normal()) : 0;
            int64_t q = (int64_t)svs[si].pos + jit;
            if (q <= (int64_t)ref) q = (int64_t)ref + 1;
            next_sv = (uint64_t)q;
        }
        uint64_t stop = std::min(std::min(next_small, next_sv), ref_end);
        m_run += (uint32_t)(stop - ref); ref = stop;
        if (ref >= ref_end) break;
        if (next_sv <= next_small) {
            const SV& s = svs[si++];
            bool carry = !s.het || hap;
            if (carry) {
                if (s.is_ins) { emit(m_run, mop); m_run = 0; emit(s.len, 1); }
                else {
                    if (ref + s.len + 1 >= ref_end) { continue; }   // deletion would run off the read: skip it
                    emit(m_run, mop); m_run = 0; emit(s.len, 2); ref += s.len;
                }
                // at least one matched base after the event
                m_run += 1; ref += 1;
            }
        } else {
            uint32_t len = 1 + (uint32_t)g.below(p.indel_len_max > 0 ? p.indel_len_max : 1);
            bool ins = g.uni() < 0.5;
            if (!ins && ref + len + 1 >= ref_end) continue;
            emit(m_run, mop); m_run = 0;
            if (ins) emit(len, 1); else { emit(len, 2); ref += len; }
            m_run += 1; ref += 1;
        }
    }
    if (ref > ref_end) { /* the +1 after an event may overshoot by one base at the very end */
        uint32_t over = (uint32_t)(ref - ref_end);
        m_run = m_run > over ? m_run - over : 0;
    }
    emit(m_run ? m_run : 1, mop);
    // trailing soft clip: forced when the read ran off the contig end
    uint32_t hang = (uint32_t)(((uint64_t)p0 + want) - std::min<uint64_t>((uint64_t)p0 + want, L));
    double ut = g.uni();
    if (hang > 0) emit(hang, 4);
    else if (ut < p.frac_softclip * 0.5) emit((g.uni() < 0.1) ? 50 : 50 + (uint32_t)g.below(2000), 4);
    r.n_ops = nops;
    return r;
}

template <class F>
void parallel_for(uint64_t n, int threads, F f)
{
    if (threads < 1) threads = 1;
    if (n < 4096 || threads == 1) { f(0, n); return; }
    std::vector<std::thread> th;
    uint64_t chunk = (n + threads - 1) / threads;
    for (int t = 0; t < threads; t++) {
        uint64_t a = (uint64_t)t * chunk, b = std::min(n, a + chunk);
        if (a >= b) break;
        th.emplace_back([=] { f(a, b); });
    }
    for (auto& t : th) t.join();
}

}  // namespace

extern "C" {

void csv_synth_default_params(csv_synth_params* p)
{
    memset(p, 0, sizeof *p);
    p->seed = 20261018; p->profile = 0; p->coverage = 30.0;
    p->read_len_mean = 15000; p->read_len_sd = 2000;
    p->indel_rate = 0.002; p->indel_len_max = 3;
    p->n_sv = 400; p->sv_len_max = 10000; p->sv_jitter_sd = 0; p->frac_len50 = 0.05;
    p->frac_softclip = 0.02; p->frac_supplementary = 0.03; p->frac_secondary = 0.02;
    p->frac_dup = 0.01; p->frac_qcfail = 0.01; p->frac_lowmapq = 0.05; p->use_eqx = 0;
    p->threads = 0;
}

uint64_t csv_synth_num_reads(const csv_synth_params* p, uint32_t n_contigs, const uint32_t* contig_len)
{
    Plan pl = make_plan(p, n_contigs, contig_len);
    return pl.read_base[n_contigs];
}

int csv_synth_reads(const csv_synth_params* p, uint32_t n_contigs, const uint32_t* contig_len,
                    int32_t* tid, int32_t* pos0, uint16_t* flag, uint8_t* mapq, uint64_t* cig_off,
                    uint64_t* n_ops_out)
{
    Plan pl = make_plan(p, n_contigs, contig_len);
    uint64_t n = pl.read_base[n_contigs];
    int threads = p->threads > 0 ? p->threads : (int)std::max(1u, std::thread::hardware_concurrency());
    parallel_for(n, threads, [&](uint64_t a, uint64_t b) {
        uint32_t c = (uint32_t)(std::upper_bound(pl.read_base.begin(), pl.read_base.end(), a) - pl.read_base.begin() - 1);
        for (uint64_t i = a; i < b; i++) {
            while (i >= pl.read_base[c + 1]) c++;
            ReadOut r = gen_read(pl, c, i, nullptr);
            tid[i] = (int32_t)c; pos0[i] = r.pos0; flag[i] = r.flag; mapq[i] = r.mapq;
            cig_off[i + 1] = r.n_ops;
        }
    });
    cig_off[0] = 0;
    for (uint64_t i = 0; i < n; i++) cig_off[i + 1] += cig_off[i];
    *n_ops_out = cig_off[n];
    return 0;
}

int csv_synth_cigar(const csv_synth_params* p, uint32_t n_contigs, const uint32_t* contig_len,
                    const uint64_t* cig_off, uint32_t* cigar)
{
    Plan pl = make_plan(p, n_contigs, contig_len);
    uint64_t n = pl.read_base[n_contigs];
    int threads = p->threads > 0 ? p->threads : (int)std::max(1u, std::thread::hardware_concurrency());
    parallel_for(n, threads, [&](uint64_t a, uint64_t b) {
        uint32_t c = (uint32_t)(std::upper_bound(pl.read_base.begin(), pl.read_base.end(), a) - pl.read_base.begin() - 1);
        for (uint64_t i = a; i < b; i++) {
            while (i >= pl.read_base[c + 1]) c++;
            gen_read(pl, c, i, cigar + cig_off[i]);
        }
    });
    return 0;
}

}  // extern "C"
