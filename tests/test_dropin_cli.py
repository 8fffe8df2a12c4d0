"""GPU: the reference's own CLI with the three hot-path entry points replaced at link time by the C++ host
mirror (contextsv_b200/csrc/host/*_gpu.cpp -> C ABI -> CUDA) must write the same VCF as the unmodified
reference CLI on the same synthetic BAM.  Both binaries are built in the build container by
`make -C oracle ref dropin` (oracle/_ref/, prebuilt files travel to the GPU box)."""
import os
import subprocess

import pytest

from contextsv_b200 import bamio, synth

pytestmark = pytest.mark.gpu

REF_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")


def run_cli(exe, d, out, extra=(), env=None):
    os.makedirs(out, exist_ok=True)
    cmd = [exe, "-b", d + "/x.bam", "-r", d + "/x.fa", "-s", d + "/snps.vcf", "-o", out, "--hmm", os.path.join(REF_DIR, "wgs.hmm")] + list(extra)
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=180, env=dict(os.environ, **(env or {})))
    assert p.returncode == 0 and "ContextSV finished successfully!" in p.stdout, p.stdout[-2000:]
    with open(os.path.join(out, "output.vcf")) as f:
        return [l for l in f if not l.startswith("##fileDate")]


@pytest.mark.parametrize("extra,max_ops", [((), None), (("-c", "chr21"), None), ((), "3000"), (("-c", "chr21"), "1500")])
def test_cli_vcf_identical(tmp_path, extra, max_ops):
    """max_ops: the host mirror streams each chromosome in shards of at most that many CIGAR ops (what it does on its own
    once a chromosome holds more than a batch takes -- 60x ONT); the VCF must not change."""
    ref_exe, gpu_exe = os.path.join(REF_DIR, "contextsv_ref"), os.path.join(REF_DIR, "contextsv_gpu")
    if not (os.path.exists(ref_exe) and os.path.exists(gpu_exe)):
        pytest.skip("oracle/_ref CLIs not built (make -C oracle ref dropin)")
    d = str(tmp_path)
    clen, names = [150000, 90000], ["chr21", "chr22"]
    r = synth.generate(clen, seed=11, n_sv=60, coverage=20.0, frac_len50=0.2)
    bamio.write_bam(d + "/x.bam", r, names, clen, seed=1)
    bamio.write_fasta(d + "/x.fa", names, clen)
    open(d + "/snps.vcf", "w").write("##fileformat=VCFv4.2\n")
    want = run_cli(ref_exe, d, d + "/out_ref", extra)
    got = run_cli(gpu_exe, d, d + "/out_gpu", extra, {"CONTEXTSV_MAX_OPS": max_ops} if max_ops else None)
    assert len([l for l in want if not l.startswith("#")]) > 10
    assert got == want
