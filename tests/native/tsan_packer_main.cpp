// ThreadSanitizer driver for the threaded BGZF read-ahead of the htslib shim under the host packer (test infrastructure):
//   tsan_packer <bam> <threads> <chrom>...   prints "<chrom> <records> <ops> <checksum>" per contig.
// Built by tests/test_host_packer.py with -fsanitize=thread from host_packer_harness.cpp + packed_reads.cpp + shim.cpp.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

extern "C" {
void hp_set_threads(int n);
void* hp_pack(const char* bam, const char* chrom, int keep_seq);
void hp_free(void* p);
uint64_t hp_size(void* p);
uint64_t hp_ops(void* p);
void hp_copy(void* p, int32_t* tid, int32_t* pos0, uint16_t* flag, uint8_t* mapq, uint64_t* cig_off, uint32_t* cigar, uint32_t* ref_end);
}

int main(int argc, char** argv)
{
    if (argc < 4) return 2;
    hp_set_threads(std::atoi(argv[2]));
    for (int rep = 0; rep < 2; rep++)                         // twice: the second round seeks back (read-ahead restarts)
        for (int a = argc - 1; a >= 3; a--) {                 // out of file order: every query seeks
            void* p = hp_pack(argv[1], argv[a], 1);
            if (!p) return 3;
            const uint64_t n = hp_size(p), ops = hp_ops(p);
            std::vector<int32_t> tid(n), pos0(n); std::vector<uint16_t> flag(n); std::vector<uint8_t> mapq(n);
            std::vector<uint64_t> off(n + 1); std::vector<uint32_t> cigar(ops ? ops : 1), ref_end(n);
            hp_copy(p, tid.data(), pos0.data(), flag.data(), mapq.data(), off.data(), cigar.data(), ref_end.data());
            uint64_t h = 1469598103934665603ull;
            auto mix = [&](uint64_t v) { h = (h ^ v) * 1099511628211ull; };
            for (uint64_t i = 0; i < n; i++) { mix((uint32_t)pos0[i]); mix(flag[i]); mix(mapq[i]); mix(off[i + 1]); mix(ref_end[i]); }
            for (uint64_t i = 0; i < ops; i++) mix(cigar[i]);
            if (rep == 1) std::printf("%s %llu %llu %016llx\n", argv[a], (unsigned long long)n, (unsigned long long)ops, (unsigned long long)h);
            hp_free(p);
        }
    return 0;
}
