"""Per-kernel SASS mnemonic counts of libdiffvits_b200.so (cuobjdump -sass): which kernels are real tcgen05 / TMEM / TMA
code (UTCHMMA / LDTM / UTMALDG) and which are mma.sync / cp.async code (HMMA / LDGSTS). Runs without a GPU.

    python tools/sass_summary.py > profiles/sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "diff-vits_b200", "libdiffvits_b200.so")
MNEMONICS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMAPF", "SYNCS", "HMMA", "LDGSTS", "LDSM", "MUFU.EX2", "MUFU.RCP",
             "FFMA2", "LDG", "STG", "LDS", "STS", "BAR.SYNC"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            cur["_total"] += 1
            for mn in MNEMONICS:
                if op == mn or op.startswith(mn + "."):
                    cur[mn] += 1
    names = demangle(list(kernels))
    print(f"# SASS summary of {os.path.relpath(LIB, ROOT)} (sm_100a), instruction counts per kernel")
    print("# " + " ".join(f"{m:>8s}" for m in ["total"] + MNEMONICS) + "  kernel")
    for k, c in kernels.items():
        short = names[k].replace("(anonymous namespace)::", "").replace("void ", "")
        short = re.split(r"\((?=const |dv::|float|unsigned|int|__nv|CUtensor)", short)[0]
        short = short.replace("(bool)1", "true").replace("(bool)0", "false")
        print("  " + " ".join(f"{c.get(m, 0):8d}" for m in ["_total"] + MNEMONICS) + "  " + short)


if __name__ == "__main__":
    sys.exit(main())
