"""GPU: the reference's own CLI with the three hot-path entry points replaced at link time by the C++ host
mirror (contextsv_b200/csrc/host/*_gpu.cpp -> C ABI -> CUDA) must write the same VCF as the unmodified
reference CLI on the same synthetic BAM.  Both binaries are built in the build container by
`make -C oracle ref dropin` (oracle/_ref/, prebuilt files travel to the GPU box)."""
import os
import subprocess

import pytest

import numpy as np

from contextsv_b200 import bamio, synth
import util

pytestmark = pytest.mark.gpu

REF_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")


def run_cli(exe, d, out, extra=(), env=None):
    os.makedirs(out, exist_ok=True)
    cmd = [exe, "-b", d + "/x.bam", "-r", d + "/x.fa", "-s", d + "/snps.vcf", "-o", out, "--hmm", os.path.join(REF_DIR, "wgs.hmm")] + list(extra)
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=180, env=dict(os.environ, **(env or {})))
    assert p.returncode == 0 and "ContextSV finished successfully!" in p.stdout, p.stdout[-2000:]
    with open(os.path.join(out, "output.vcf")) as f:
        return [l for l in f if not l.startswith("##fileDate")]


@pytest.mark.parametrize("extra,max_ops", [((), None), (("-c", "chr21"), None), ((), "3000"), (("-c", "chr21"), "1500"), ((), "host_depth"), ((), "host_depth,2500")])
def test_cli_vcf_identical(tmp_path, extra, max_ops):
    """max_ops: the host mirror streams each chromosome in shards of at most that many CIGAR ops (what it does on its own
    once a chromosome holds more than a batch takes -- 60x ONT); the VCF must not change.  host_depth: CONTEXTSV_HOST_DEPTH=1,
    the depth map is also brought into the caller's vectors (narrow fetch) and its consumers read it there, like the reference."""
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
    env = {}
    for item in (max_ops or "").split(","):
        if item == "host_depth":
            env["CONTEXTSV_HOST_DEPTH"] = "1"
        elif item:
            env["CONTEXTSV_MAX_OPS"] = item
    got = run_cli(gpu_exe, d, d + "/out_gpu", extra, env or None)
    assert len([l for l in want if not l.startswith("#")]) > 10
    assert got == want


def test_cli_vcf_identical_on_two_gpus(tmp_path):
    """CONTEXTSV_GPUS=0,1: the depth pass hands contigs (and, with a small op budget, the shards of a contig) round-robin to
    one worker per device; depth consumers ask the device that holds the shard.  Same VCF.  Needs two GPUs (skipped on the
    single-GPU boxes of the regular run; `gpurun --gpus 2 -- python -m pytest tests/test_dropin_cli.py -m gpu`)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    ref_exe, gpu_exe = os.path.join(REF_DIR, "contextsv_ref"), os.path.join(REF_DIR, "contextsv_gpu")
    if not (os.path.exists(ref_exe) and os.path.exists(gpu_exe)):
        pytest.skip("oracle/_ref CLIs not built (make -C oracle ref dropin)")
    d = str(tmp_path)
    clen, names = [150000, 90000, 60000], ["chr20", "chr21", "chr22"]
    r = synth.generate(clen, seed=13, n_sv=70, coverage=20.0, frac_len50=0.2)
    bamio.write_bam(d + "/x.bam", r, names, clen, seed=1)
    bamio.write_fasta(d + "/x.fa", names, clen)
    open(d + "/snps.vcf", "w").write("##fileformat=VCFv4.2\n")
    want = run_cli(ref_exe, d, d + "/out_ref")
    assert len([l for l in want if not l.startswith("#")]) > 10
    for k, env in enumerate(({"CONTEXTSV_GPUS": "0,1"}, {"CONTEXTSV_GPUS": "0,1", "CONTEXTSV_MAX_OPS": "3000"}, {"CONTEXTSV_GPUS": "1"})):
        assert run_cli(gpu_exe, d, d + "/out_gpu%d" % k, env=env) == want, env


def split_bam(d, seed=23):
    clen, names = [400000, 260000], ["chr21", "chr22"]
    r = synth.generate(clen, seed=seed, n_sv=40, coverage=12.0, frac_len50=0.2)
    r2, qnames = util.add_split_events(r, clen, np.random.default_rng(seed), n_events=12)
    bamio.write_bam(d + "/x.bam", r2, names, clen, seed=1, qnames=qnames)
    bamio.write_fasta(d + "/x.fa", names, clen)
    open(d + "/snps.vcf", "w").write("##fileformat=VCFv4.2\n")
    return r2


@pytest.mark.parametrize("extra,max_ops", [((), None), (("-c", "chr22"), None), ((), "4000")])
def test_cli_split_reads_identical(tmp_path, extra, max_ops):
    """a BAM with split alignments (primary + supplementary records sharing a query name).  The
    drop-in decodes the file ONCE -- its findSplitSVSignatures works from what the depth pass parked (device-side
    record summaries, one batched DBSCAN1D launch sequence for all overlap groups) -- and must produce the reference's
    split candidates line for line and the reference's VCF byte for byte."""
    from oracle.oracle_py import Reference
    ref_exe, gpu_exe = os.path.join(REF_DIR, "contextsv_ref"), os.path.join(REF_DIR, "contextsv_gpu")
    if not (os.path.exists(ref_exe) and os.path.exists(gpu_exe)):
        pytest.skip("oracle/_ref CLIs not built (make -C oracle ref dropin)")
    d = str(tmp_path)
    split_bam(d)
    chrom = extra[1] if extra else ""
    n_ref = Reference().split_dump(d + "/x.bam", d + "/split_ref.txt", chrom=chrom)
    assert n_ref >= 6
    want = run_cli(ref_exe, d, d + "/out_ref", extra)
    env = {"CONTEXTSV_B200_DUMP_SPLIT": d + "/split_gpu.txt", "CONTEXTSV_B200_STATS": "1"}
    if max_ops:
        env["CONTEXTSV_MAX_OPS"] = max_ops
    got = run_cli(gpu_exe, d, d + "/out_gpu", extra, env)
    assert sorted(open(d + "/split_gpu.txt").read().splitlines()) == sorted(open(d + "/split_ref.txt").read().splitlines())
    assert any("SPLIT" in l for l in want if not l.startswith("#"))
    assert got == want
