"""Per-kernel SASS evidence of the in-tree library: which kernels stage their tile with TMA (UTMALDG + mbarrier SYNCS),
prefetch through the bulk-copy unit (UBLKPF), and how much of their code is FP32 pipe work.

    python tools/sass_summary.py [path/to/libcvsteer_b200.so] > profiles/sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "cvsteer_b200", "libcvsteer_b200.so")
KEYS = ("UTMALDG", "SYNCS", "UBLKPF", "FFMA", "FADD", "FMUL", "MUFU", "LDS", "LDG", "STG", "ATOM", "RED", "BRA", "BAR")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts, order, cur, arch = {}, [], None, collections.Counter()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        m = re.search(r"arch = (sm_\w+)", line)
        if m:
            arch[m.group(1)] += 1
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            counts[cur]["TOTAL"] += 1
            for k in KEYS:
                if op.startswith(k):
                    counts[cur][k] += 1
    names = demangle(order)
    print(f"# {os.path.relpath(LIB, ROOT)}: {len(order)} kernels; cubin architectures: {dict(arch)}")
    print("# counts are STATIC instructions per kernel (cuobjdump -sass); UTMALDG = cp.async.bulk.tensor (TMA), SYNCS = mbarrier ops,")
    print("# UBLKPF = cp.async.bulk.prefetch.L2; no UTC*MMA / HMMA anywhere: a 9/13-tap stencil is not a contraction")
    print(f"{'kernel':110s} {'TOTAL':>6s} " + " ".join(f"{k:>7s}" for k in KEYS))
    tot = collections.Counter()
    for fn in sorted(order, key=lambda f: names[f]):
        nm = re.sub(r"\(CUtensorMap_st.*", "", names[fn]).replace("void cvs::", "").replace("cvs::", "")
        nm = nm.replace("(unsigned int)", "").replace("(bool)", "").replace("unsigned char", "u8")
        c = counts[fn]
        tot.update(c)
        print(f"{nm[:110]:110s} {c['TOTAL']:6d} " + " ".join(f"{c[k]:7d}" for k in KEYS))
    print(f"{'ALL KERNELS':110s} {tot['TOTAL']:6d} " + " ".join(f"{tot[k]:7d}" for k in KEYS))
    bad = [k for k in ("HMMA", "UTCHMMA", "HGMMA") if re.search(r"\b" + k, sass)]
    print("# tensor-core mnemonics present: " + (", ".join(bad) if bad else "none"))


if __name__ == "__main__":
    main()
