"""Diagnostic: the drop-in CLI with CONTEXTSV_GPUS=0,1 on a small synthetic BAM; prints the CLI's output."""
import os, subprocess, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from contextsv_b200 import bamio, synth
REF_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")
d = tempfile.mkdtemp()
clen, names = [150000, 90000, 60000], ["chr20", "chr21", "chr22"]
r = synth.generate(clen, seed=13, n_sv=70, coverage=20.0, frac_len50=0.2)
bamio.write_bam(d + "/x.bam", r, names, clen, seed=1)
bamio.write_fasta(d + "/x.fa", names, clen)
open(d + "/snps.vcf", "w").write("##fileformat=VCFv4.2\n")
for gpus in ("0", "0,1"):
    out = d + "/out_" + gpus.replace(",", "_")
    os.makedirs(out, exist_ok=True)
    cmd = [os.path.join(REF_DIR, "contextsv_gpu"), "-b", d + "/x.bam", "-r", d + "/x.fa", "-s", d + "/snps.vcf", "-o", out, "--hmm", os.path.join(REF_DIR, "wgs.hmm")]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=180, env=dict(os.environ, CONTEXTSV_GPUS=gpus, CONTEXTSV_B200_STATS="1"))
    print("==== CONTEXTSV_GPUS=%s rc=%d files=%s" % (gpus, p.returncode, os.listdir(out)))
    print(p.stdout[-3500:])
