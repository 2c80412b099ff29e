#!/usr/bin/env python
"""Per-kernel counts of the Blackwell-specific SASS mnemonics in libb200rag.so (cuobjdump -sass), the evidence
that the hot path is tcgen05 / TMEM / TMA code: UTCHMMA (tcgen05.mma), UTMALDG (TMA tensor load), UBLKCP (TMA bulk
copy), LDTM (tcgen05.ld), UTCBAR (tcgen05.commit), SYNCS (mbarrier), UCGABAR (cluster barrier), ATOMS / REDS
(shared-memory atomics).  Usage: python tools/sass_summary.py > profiles/rN_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "rag-dpo_b200", "b200rag", "libb200rag.so")
WATCH = ("UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "SYNCS",
         "UCGABAR", "ATOMS", "REDS", "RED", "ATOMG", "HMMA", "DFMA", "DADD", "DMUL", "LDG", "LDS", "STS")


def main():
    so = sys.argv[1] if len(sys.argv) > 1 else SO
    out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    demangle = {}
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            base = op.split(".")[0]
            if base in WATCH:
                # keep the qualifiers that matter (.2CTA, .MULTICAST, .2D, .x32 ...)
                keep = ".".join([base] + [p for p in op.split(".")[1:] if p in ("2CTA", "MULTICAST", "2D", "S", "G",
                                                                                   "ARV", "WAIT", "ADD", "x32", "x1",
                                                                                   "32x32b")])
                kernels[cur][keep] += 1
    names = list(kernels)
    try:
        dm = subprocess.run(["c++filt"] + names, capture_output=True, text=True, check=True).stdout.splitlines()
        demangle = dict(zip(names, dm))
    except Exception:
        pass
    print(f"# SASS summary of {os.path.relpath(so, ROOT)} (cuobjdump -sass, sm_100a): instructions per kernel")
    agg = collections.Counter()
    for k, cnt in kernels.items():
        name = demangle.get(k, k)
        name = re.sub(r"\(.*", "", name).replace("b200rag::", "").replace("void ", "")
        parts = [f"{op} x{n}" for op, n in sorted(cnt.items()) if op != "_total" and op.split(".")[0] not in
                 ("LDG", "LDS", "STS", "DFMA", "DADD", "DMUL")]
        fp64 = sum(n for op, n in cnt.items() if op.split(".")[0] in ("DFMA", "DADD", "DMUL"))
        print(f"{name}: {cnt['_total']} instr" + (f", fp64 x{fp64}" if fp64 else "") + ("; " + ", ".join(parts) if parts else ""))
        for op, n in cnt.items():
            if op != "_total":
                agg[op] += n
    print("\n# totals over all kernels")
    for op, n in sorted(agg.items()):
        print(f"{op}: {n}")


if __name__ == "__main__":
    main()
