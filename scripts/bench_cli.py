"""CLI-level timing (SURVEY 8d ii, BASELINE configs[0] in small): the unmodified reference CLI (oracle/_ref/contextsv_ref)
against the same CLI with the hot path replaced at link time (oracle/_ref/contextsv_gpu) on one synthetic 30x HiFi
contig, `-c chr21`, wall clock, same BAM; the VCFs must be identical.  Both CLIs decode the BAM through the same
single-threaded shim (htslib is not in the image), once per pass, so this number is decode-bound on both sides -- it
shows what the drop-in buys a user today, not what the kernels do.  GPU box only.

  python scripts/bench_cli.py [--mbp 10] [--threads N]
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from contextsv_b200 import bamio, synth  # noqa: E402

REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def run(exe, d, out, threads, env=None):
    os.makedirs(out, exist_ok=True)
    cmd = [exe, "-b", d + "/x.bam", "-r", d + "/x.fa", "-s", d + "/snps.vcf", "-o", out, "--hmm", os.path.join(REF_DIR, "wgs.hmm"),
           "-c", "chr21", "-t", str(threads)]
    t0 = time.perf_counter()
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=dict(os.environ, **(env or {})))
    dt = time.perf_counter() - t0
    if env:
        sys.stderr.write("".join(l + "\n" for l in p.stdout.splitlines() if "[contextsv_b200]" in l or "lapsed" in l or "ime" in l.split(":")[0][-6:]))
    if p.returncode != 0 or "ContextSV finished successfully!" not in p.stdout:
        sys.stderr.write(p.stdout[-3000:])
        raise SystemExit("CLI failed: " + exe)
    with open(os.path.join(out, "output.vcf")) as f:
        vcf = [l for l in f if not l.startswith("##fileDate")]
    return dt, vcf


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mbp", type=float, default=10.0)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--reps", type=int, default=2)
    a = ap.parse_args()
    clen, names = [int(a.mbp * 1e6)], ["chr21"]
    with tempfile.TemporaryDirectory() as d:
        t0 = time.perf_counter()
        r = synth.generate(clen, seed=20261019, n_sv=int(200 * a.mbp / 46.7) + 10, coverage=30.0)
        bamio.write_bam(d + "/x.bam", r, names, clen, seed=1)
        bamio.write_fasta(d + "/x.fa", names, clen)
        open(d + "/snps.vcf", "w").write("##fileformat=VCFv4.2\n")
        t_gen = time.perf_counter() - t0
        res = {"workload": "synthetic 30x HiFi, one contig of %.1f Mbp, -c chr21" % a.mbp, "reads": int(r["n_reads"]), "cigar_ops": int(r["n_ops"]),
               "bam_bytes": os.path.getsize(d + "/x.bam"), "threads": a.threads, "generation_s": round(t_gen, 1)}
        t_ref, t_gpu, vcf_ref, vcf_gpu = [], [], None, None
        for rep in range(a.reps):
            dt, vcf_gpu = run(os.path.join(REF_DIR, "contextsv_gpu"), d, d + "/out_gpu", a.threads, {"CONTEXTSV_B200_STATS": "1"} if rep == 0 else None); t_gpu.append(dt)
            dt, vcf_ref = run(os.path.join(REF_DIR, "contextsv_ref"), d, d + "/out_ref", a.threads); t_ref.append(dt)
        res.update({"contextsv_ref_s": [round(x, 2) for x in t_ref], "contextsv_gpu_s": [round(x, 2) for x in t_gpu],
                    "vcf_identical": vcf_ref == vcf_gpu, "vcf_records": len([l for l in vcf_ref if not l.startswith("#")]),
                    "speedup_best": round(min(t_ref) / min(t_gpu), 2)})
        print(json.dumps(res))


if __name__ == "__main__":
    main()
