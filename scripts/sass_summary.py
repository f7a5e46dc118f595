"""SASS instruction summary of the in-tree library (runs without a GPU): per kernel family, the mnemonics that prove the
Blackwell-native paths (B200_PROFILING.md "What proves a Blackwell-native kernel"): UTC*MMA = tcgen05.mma, LDTM / STTM =
tcgen05.ld / st, UTMALDG / UTMASTG = TMA tensor loads / stores, HMMA = mma.sync (legacy tensor path), plus peer /
system-scope memory ops of the exchange kernel.   python scripts/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / "mimo_b200" / "libmimo_b200.so"
KEYS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "HMMA", "MUFU.EX2", "MUFU.RCP",
        "SYNCS", "LDG", "STG", "LDGSTS", "LD.E.STRONG.SYS", "ST.E.STRONG.SYS", "ATOM", "RED", "BAR.SYNC", "CCTL"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    fam = collections.defaultdict(collections.Counter)
    n_inst = collections.Counter()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(.*", "", name).replace("void ", "").replace("mimo::", "")
            cur = re.sub(r"<.*", "", cur)  # one row per kernel template
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(2)
            n_inst[cur] += 1
            for k in KEYS:
                if op.startswith(k) or (k.startswith(("LD.", "ST.")) and k in line):
                    fam[cur][k] += 1
    print(f"SASS summary of {LIB.name} (all template instantiations of a kernel summed)\n")
    cols = [k for k in KEYS if any(fam[f][k] for f in fam)]
    print(f"{'kernel':28s} {'instr':>8s} " + " ".join(f"{c[:9]:>9s}" for c in cols))
    for f in sorted(fam, key=lambda f: -n_inst[f]):
        print(f"{f[:28]:28s} {n_inst[f]:8d} " + " ".join(f"{fam[f][c]:9d}" for c in cols))
    tot = collections.Counter()
    for f in fam:
        tot.update(fam[f])
    print(f"{'TOTAL':28s} {sum(n_inst.values()):8d} " + " ".join(f"{tot[c]:9d}" for c in cols))


if __name__ == "__main__":
    sys.exit(main())
