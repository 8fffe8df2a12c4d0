"""Minimal BAM / FASTA writers for synthetic test inputs (pure Python + zlib; no htslib, no samtools).

Writes coordinate-sorted BAM (BGZF blocks of raw deflate, per the SAM/BAM specification) from the packed SoA
the generator produces, so that the reference CLI and the drop-in CLI can be run on the very same records.
"""
import struct
import zlib

import numpy as np

_QMASK = (1 << 0) | (1 << 1) | (1 << 4) | (1 << 7) | (1 << 8)
_RMASK = (1 << 0) | (1 << 2) | (1 << 3) | (1 << 7) | (1 << 8)


class BgzfWriter:
    def __init__(self, path, level=1):
        self.f = open(path, "wb")
        self.buf = bytearray()
        self.level = level

    def write(self, b):
        self.buf += b
        while len(self.buf) >= 0xff00:
            self._block(bytes(self.buf[:0xff00]))
            del self.buf[:0xff00]

    def _block(self, data):
        c = zlib.compressobj(self.level, zlib.DEFLATED, -15)
        comp = c.compress(data) + c.flush()
        bsize = len(comp) + 25
        hdr = struct.pack("<BBBBIBBHBBHH", 31, 139, 8, 4, 0, 0, 255, 6, 66, 67, 2, bsize)
        self.f.write(hdr + comp + struct.pack("<II", zlib.crc32(data) & 0xffffffff, len(data)))

    def close(self):
        if self.buf:
            self._block(bytes(self.buf))
        self._block(b"")       # EOF marker
        self.f.close()


def reg2bin(beg, end):
    end -= 1
    if beg >> 14 == end >> 14: return ((1 << 15) - 1) // 7 + (beg >> 14)
    if beg >> 17 == end >> 17: return ((1 << 12) - 1) // 7 + (beg >> 17)
    if beg >> 20 == end >> 20: return ((1 << 9) - 1) // 7 + (beg >> 20)
    if beg >> 23 == end >> 23: return ((1 << 6) - 1) // 7 + (beg >> 23)
    if beg >> 26 == end >> 26: return ((1 << 3) - 1) // 7 + (beg >> 26)
    return 0


def random_seq4(rng, qlen):
    """4-bit packed bases (A,C,G,T = 1,2,4,8) for a read of qlen bases."""
    codes = np.array([1, 2, 4, 8], np.uint8)[rng.integers(0, 4, qlen + (qlen & 1))]
    return ((codes[0::2] << 4) | codes[1::2]).astype(np.uint8)


def write_bam(path, reads, contig_names, contig_len, seed=0, seq4=None, seq_off=None, qnames=None):
    """qnames: optional query name per record (default r<i>: every record its own template).  A record with more than
    65 535 CIGAR ops is written the way htslib writes it (SAM spec 4.2.2): the real CIGAR in a CG:B,I tag and the
    placeholder <l_seq>S<ref_len>N in the 16-bit CIGAR field."""
    rng = np.random.default_rng(seed)
    w = BgzfWriter(path)
    text = "@HD\tVN:1.6\tSO:coordinate\n" + "".join("@SQ\tSN:%s\tLN:%d\n" % (n, l) for n, l in zip(contig_names, contig_len))
    hdr = b"BAM\x01" + struct.pack("<i", len(text)) + text.encode() + struct.pack("<i", len(contig_names))
    for n, l in zip(contig_names, contig_len):
        hdr += struct.pack("<i", len(n) + 1) + n.encode() + b"\0" + struct.pack("<i", int(l))
    w.write(hdr)
    cig = np.asarray(reads["cigar"]); off = np.asarray(reads["cig_off"]).astype(np.int64)
    tid = reads.get("tid")
    for i in range(int(reads["n_reads"])):
        c = cig[off[i]:off[i + 1]]
        qlen = int(((c >> 4) * ((_QMASK >> (c & 15)) & 1)).sum())
        rlen = int(((c >> 4) * ((_RMASK >> (c & 15)) & 1)).sum())
        name = (("r%d" % i) if qnames is None else str(qnames[i])).encode() + b"\0"
        pos = int(reads["pos0"][i])
        if seq4 is not None:
            s0 = int(seq_off[i]); sq = bytes(seq4[s0:s0 + (qlen + 1) // 2])
        else:
            sq = bytes(random_seq4(rng, qlen))
        aux = b""
        c_field = c
        if len(c) > 0xffff:
            aux = b"CGBI" + struct.pack("<I", len(c)) + c.astype("<u4").tobytes()
            c_field = np.array([(qlen << 4) | 4, (rlen << 4) | 3], np.uint32)
        rec = struct.pack("<iiBBHHHiiii", int(tid[i]) if tid is not None else 0, pos, len(name), int(reads["mapq"][i]),
                          reg2bin(pos, pos + max(rlen, 1)), len(c_field), int(reads["flag"][i]), qlen, -1, -1, 0)
        rec += name + c_field.astype("<u4").tobytes() + sq + b"\xff" * qlen + aux
        w.write(struct.pack("<i", len(rec)) + rec)
    w.close()


def write_fasta(path, contig_names, contig_len, seed=0):
    rng = np.random.default_rng(seed)
    with open(path, "w") as f:
        for n, l in zip(contig_names, contig_len):
            f.write(">%s\n" % n)
            s = np.array(list(b"ACGT"), np.uint8)[rng.integers(0, 4, int(l))].tobytes().decode()
            for k in range(0, len(s), 60):
                f.write(s[k:k + 60] + "\n")
