"""Count the Blackwell-specific SASS instructions per kernel of the shipped library.

    python bench_tools/sass_summary.py [> profiles/sass_summary.txt]

Evidence that the hot path runs on tcgen05 / TMEM / TMA (B200_PROFILING.md lists the mnemonics):
UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA tensor load / store,
UTCBAR = tcgen05.commit, SYNCS = mbarrier, UTCATOMSWS = TMEM allocation.
"""

import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "joshupscale_b200", "lib", "libJoshUpscale.so")
MNEMONICS = ["UTCHMMA", "UTCQMMA", "UTCOMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "UTCATOMSWS", "SYNCS",
             "HMMA", "FFMA2", "FFMA", "LDG", "STG", "LDS", "STS"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    ldd = subprocess.run(["ldd", LIB], capture_output=True, text=True).stdout
    counts = collections.OrderedDict()
    current = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = name.replace("(anonymous namespace)::", "").replace("void ", "")
            name = re.sub(r"\(.*", "", name).replace("ju::", "")
            current = counts.setdefault(name, collections.Counter())
            continue
        if current is None:
            continue
        m = re.search(r"/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1).split(".")[0]
            if op in MNEMONICS:
                current[op] += 1
            current["total"] += 1
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
    print(f"# {os.path.relpath(LIB, ROOT)}: cuobjdump -sass, instruction counts per kernel (arch {', '.join(arch)})")
    print("# linked libraries: " + ", ".join(sorted({ln.split()[0] for ln in ldd.splitlines() if ln.strip()})))
    cols = [m for m in MNEMONICS if any(c[m] for c in counts.values())]
    print(f"{'kernel':58s} " + " ".join(f"{c:>10s}" for c in cols) + f" {'total':>8s}")
    for name, c in counts.items():
        print(f"{name[:58]:58s} " + " ".join(f"{c[m]:10d}" for m in cols) + f" {c['total']:8d}")
    tot = collections.Counter()
    for c in counts.values():
        tot.update(c)
    print(f"{'ALL':58s} " + " ".join(f"{tot[m]:10d}" for m in cols) + f" {tot['total']:8d}")


if __name__ == "__main__":
    sys.exit(main())
