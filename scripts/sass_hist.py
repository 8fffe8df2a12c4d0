"""SASS instruction histograms of the hot kernels, from the shipped library (cuobjdump -sass): the evidence for
bulk-TMA copies (UBLKCP), mbarriers (SYNCS), 256-bit stores (STG.E.ENL2.256 / .256), IDP.2A prefix sums and the absence
of local-memory traffic.  python scripts/sass_hist.py [kernel-regex ...] > profiles/r2_sass_hist.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "contextsv_b200", "libcontextsv_b200.so")


def main():
    pats = [re.compile(p) for p in (sys.argv[1:] or ["k_walk", "k_depth_tiles16", "k_span_carry", "k_pmax_chained", "k_sort_pass", "k_depth_narrow"])]
    out = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip()
            cur = name if any(p.search(name) for p in pats) else None
            if cur:
                kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            kernels[cur][m.group(2)] += 1
    print("# SASS instruction histograms (cuobjdump -sass contextsv_b200/libcontextsv_b200.so, sm_100a)\n")
    for name, c in kernels.items():
        total = sum(c.values())
        short = re.sub(r"\(.*", "", name)
        print("## `%s` -- %d instructions\n" % (short, total))
        groups = collections.Counter()
        for op, n in c.items():
            groups[op.split(".")[0]] += n
        print("by mnemonic: " + ", ".join("%s %d" % (k, v) for k, v in groups.most_common(24)) + "\n")
        notable = [(op, n) for op, n in sorted(c.items(), key=lambda x: -x[1])
                   if re.match(r"(UBLKCP|SYNCS|STG|LDG|STS|LDS|IDP|RED|ATOM|SHFL|REDUX|VIMNMX|LDL|STL|MATCH|VOTE|BAR|LDGSTS|UTMA|LOP3|SHF|IMAD|IADD3|PRMT)", op)]
        print("| instruction | count |\n|---|---|")
        for op, n in notable[:40]:
            print("| `%s` | %d |" % (op, n))
        spill = sum(n for op, n in c.items() if op.startswith(("LDL", "STL")))
        print("\nlocal-memory instructions (LDL/STL): %d\n" % spill)


if __name__ == "__main__":
    main()
