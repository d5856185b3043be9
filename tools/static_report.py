"""Static evidence for profiles/: per kernel, ptxas registers / spills / smem (from build/ptxas.log) and the SASS mnemonics
that prove which hardware path it uses (cuobjdump -sass of the built library): UTCHMMA = tcgen05.mma, UTMALDG / UBLKCP =
TMA tensor / bulk copies, UTCBAR = tcgen05.commit, LDTM = tcgen05.ld, SYNCS = mbarrier ops, HMMA = legacy mma.sync
(see /opt/skills/guides/B200_PROFILING.md for the mnemonic list).  CPU only.

Usage: python tools/static_report.py > profiles/rNN_static.csv
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "dcs-net_b200")
MNEMONICS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "HMMA", "FFMA2", "FFMA", "MUFU",
             "LDG", "STG", "LDS", "STS", "ATOM", "RED", "BAR", "R2UR", "CALL"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    m = re.match(r"([A-Za-z0-9_:]+)(<.*>)?\(", name)
    if not m:
        return name[:80]
    targs = m.group(2) or ""
    targs = re.sub(r"\(([A-Za-z_:]+)\)(\d+)", r"\2", targs)
    return (m.group(1) + targs)[:90]


def ptxas_info():
    info = {}
    log = open(os.path.join(PKG, "build", "ptxas.log")).read()
    cur = None
    for line in log.splitlines():
        m = re.search(r"Compiling entry function '([^']+)'", line)
        if m:
            cur = m.group(1)
            info[cur] = {"regs": None, "spill_st": 0, "spill_ld": 0, "smem": 0, "barriers": 0}
            continue
        if cur is None:
            continue
        m = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m:
            info[cur]["spill_st"], info[cur]["spill_ld"] = int(m.group(1)), int(m.group(2))
        m = re.search(r"Used (\d+) registers", line)
        if m:
            info[cur]["regs"] = int(m.group(1))
            b = re.search(r"used (\d+) barriers", line)
            s = re.search(r"(\d+) bytes smem", line)
            info[cur]["barriers"] = int(b.group(1)) if b else 0
            info[cur]["smem"] = int(s.group(1)) if s else 0
    return info


def sass_counts():
    out = subprocess.run(["cuobjdump", "-sass", os.path.join(PKG, "libdcsnet_sm100a.so")], capture_output=True, text=True).stdout
    counts = {}
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?\w+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            counts[cur]["_total"] += 1
            for mn in MNEMONICS:
                if op == mn or (mn in ("LDG", "STG", "LDS", "STS", "ATOM", "RED", "BAR", "MUFU", "SYNCS", "CALL") and op.startswith(mn)):
                    counts[cur][mn] += 1
                    break
    return counts


def main():
    info, counts = ptxas_info(), sass_counts()
    names = sorted(set(info) | set(counts))
    dm = demangle(names)
    print(",".join(["kernel", "regs", "spill_store_B", "spill_load_B", "static_smem_B", "sass_instr"] + MNEMONICS))
    rows = []
    for n in names:
        i = info.get(n, {})
        c = counts.get(n, collections.Counter())
        rows.append((short(dm.get(n, n)), [i.get("regs"), i.get("spill_st"), i.get("spill_ld"), i.get("smem"), c["_total"]] +
                     [c[m] for m in MNEMONICS]))
    for k, v in sorted(rows):
        print(",".join(['"%s"' % k] + ["" if x is None else str(x) for x in v]))


if __name__ == "__main__":
    sys.exit(main())
