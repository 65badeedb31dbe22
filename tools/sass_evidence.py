#!/usr/bin/env python
"""SASS / resource evidence for the built kernels (no GPU needed): which kernels use the TMA bulk-copy engine
(UBLKCP), mbarrier waits (SYNCS), global atomics / reductions, 128-bit loads and stores, plus registers and static
shared memory per kernel from cuobjdump --dump-resource-usage.  Mnemonics follow /opt/skills/guides/B200_PROFILING.md.

  python tools/sass_evidence.py > profiles/r1_s_sass_evidence.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "fgnn-artifacts_b200", "lib", "libfgnn_kernels.so")
PATTERNS = [("UBLKCP.S.G", r"\bUBLKCP\.S\.G"), ("UBLKCP.G.S", r"\bUBLKCP\.G\.S"), ("SYNCS(mbarrier)", r"\bSYNCS\."),
            ("ATOMG", r"\bATOMG\."), ("RED", r"\bREDG?\."), ("ATOMS", r"\bATOMS\."), ("LDG.128", r"\bLDG\.E\.[A-Z0-9.]*128"),
            ("STG.128", r"\bSTG\.E\.[A-Z0-9.]*128"), ("LDG", r"\bLDG\."), ("STG", r"\bSTG\."), ("SHFL", r"\bSHFL\."),
            ("IMAD.HI.U32", r"\bIMAD\.HI\.U32")]


def demangle(names):
    try:
        out = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
        return dict(zip(names, out))
    except OSError:
        return {n: n for n in names}


def short(name):
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = re.sub(r"^void ", "", name)
    return re.sub(r"\(.*$", "", name).replace("fgnn::", "")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts = collections.OrderedDict()
    fn = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            counts[fn] = collections.Counter()
            continue
        if fn is None:
            continue
        for key, pat in PATTERNS:
            if re.search(pat, line):
                counts[fn][key] += 1
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r"REG:(\d+).*?SHARED:(\d+).*?LOCAL:(\d+)", line)
        if m and cur:
            usage[cur] = tuple(int(x) for x in m.groups())
    names = demangle([n for n in counts if not n.startswith("_ZN3cub")])
    print("# %s (sm_100a) — cuobjdump -sass / --dump-resource-usage; cub:: library kernels omitted"
          % os.path.relpath(LIB, ROOT))
    print("# columns: kernel | regs | static smem B | local B | instruction counts")
    for fn, c in counts.items():
        if fn not in names:
            continue
        r = usage.get(fn, ("?", "?", "?"))
        body = " ".join("%s=%d" % (k, c[k]) for k, _ in PATTERNS if c[k])
        print("%-78s | %3s | %6s | %4s | %s" % (short(names[fn])[:78], r[0], r[1], r[2], body))
    bulk = sorted({short(names[f]) for f, c in counts.items() if f in names and (c["UBLKCP.S.G"] or c["UBLKCP.G.S"])})
    print("# kernels using the TMA bulk-copy engine (cp.async.bulk -> UBLKCP): %d" % len(bulk))
    spill = [short(names[f]) for f in counts if f in names and usage.get(f, (0, 0, 0))[2]]
    print("# kernels with local-memory (spill/stack) usage: %s" % (", ".join(sorted(set(spill))) or "none"))


if __name__ == "__main__":
    sys.exit(main())
