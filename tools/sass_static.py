#!/usr/bin/env python3
"""Static SASS evidence per kernel of libfoundpose_b200.so: counts of the Blackwell-specific mnemonics
(UTCHMMA = tcgen05.mma, UTMALDG/UTMASTG/UTMAREDG = TMA load / store / reduce, LDTM/STTM = tcgen05.ld/st, UTCBAR =
tcgen05.commit, SYNCS = mbarrier) next to the legacy ones (HMMA = mma.sync) - runs here, no GPU needed.

    python tools/sass_static.py > profiles/r02_sass_mnemonics.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "foundpose_b200", "libfoundpose_b200.so")
WANT = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTMASTG", "UTMAREDG", "LDTM", "STTM", "UTCBAR", "SYNCS", "HMMA", "FFMA",
        "FFMA2", "MUFU", "LDS", "STS", "LDG", "STG", "ATOM/RED"]


def main() -> None:
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    funcs = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"^void ", "", cur)
            cur = re.sub(r"fp::\(anonymous namespace\)::|fp::", "", cur)
            cur = re.sub(r"\(.*$", "", cur.replace("(int)", "").replace("(bool)", ""))
            funcs[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            op = m.group(2)
            base = op.split(".")[0]
            c = funcs[cur]
            c["total"] += 1
            if base == "UTCHMMA":
                c["UTCHMMA"] += 1
                if ".2CTA" in op:
                    c["UTCHMMA.2CTA"] += 1
            elif base in ("ATOM", "ATOMG", "ATOMS", "RED", "REDG"):
                c["ATOM/RED"] += 1
            elif base in WANT:
                c[base] += 1
    print(f"Static SASS mnemonic counts per kernel of `foundpose_b200/libfoundpose_b200.so` "
          f"(`cuobjdump -sass`, sm_100a; {len(funcs)} kernels)\n")
    print("| kernel | instrs | " + " | ".join(WANT) + " |")
    print("|---|---|" + "---|" * len(WANT))
    for name, c in funcs.items():
        print(f"| `{name}` | {c['total']} | " + " | ".join(str(c[w]) if c[w] else "" for w in WANT) + " |")
    tot = collections.Counter()
    for c in funcs.values():
        tot.update(c)
    print(f"\nWhole library: {tot['UTCHMMA']} UTCHMMA ({tot['UTCHMMA.2CTA']} of them .2CTA), {tot['UTMALDG']} UTMALDG, "
          f"{tot['UTMASTG']} UTMASTG, {tot['UTMAREDG']} UTMAREDG, {tot['LDTM']} LDTM, {tot['STTM']} STTM, "
          f"{tot['HMMA']} legacy HMMA.")


if __name__ == "__main__":
    main()
