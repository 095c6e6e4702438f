#!/usr/bin/env python
"""Native-code evidence, regenerated from the built objects: per kernel the counts of the SASS mnemonics that show which
Blackwell units the code uses (cuobjdump -sass), and ptxas' register / spill / shared-memory figures (the logs the Makefile
keeps under vi-slam_b200/build/*.ptxas.log).  Writes profiles/sass_summary.txt and profiles/ptxas_summary.txt.
    UTC*MMA  = tcgen05.mma (UTCIMMA int8, UTCOMMA mxf4 block-scaled, UTCHMMA f16/tf32)     LDTM / STTM = tcgen05.ld / st (TMEM)
    UBLKCP   = cp.async.bulk (TMA engine, 1-D bulk copy)      UTMALDG / UTMASTG = cp.async.bulk.tensor (tensor-map TMA)
    DMMA     = mma.sync f64 (FP64 tensor cores)               SYNCS = mbarrier ops        UTCBAR = tcgen05.commit"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "vi-slam_b200", "build")
PAT = ["UTCIMMA", "UTCOMMA", "UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "UTMASTG", "DMMA", "HMMA",
       "IMMA", "SYNCS", "LDGSTS", "MUFU", "F2F", "POPC", "REDUX"]


def demangle(name):
    m = re.search(r"(_Z[A-Za-z0-9_]+)$", name)
    if m:
        name = m.group(1)
    try:
        out = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    except Exception:
        out = name
    out = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", out)
    out = re.sub(r"^void ", "", out)
    out = re.sub(r"\(.*\)$", "", out)          # drop the parameter list
    return out[:120]


def main():
    lines = ["# SASS mnemonic counts per kernel (cuobjdump -sass of vi-slam_b200/build/*.o, sm_100a); only kernels with at least one",
             "# of the listed instructions are shown.  Regenerate: python tools/sass_summary.py", ""]
    total = collections.Counter()
    for obj in sorted(glob.glob(os.path.join(BUILD, "*.o"))):
        sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
        cur, per = None, collections.OrderedDict()
        for ln in sass.splitlines():
            m = re.search(r"Function : (\S+)", ln)
            if m:
                cur = m.group(1)
                per[cur] = collections.Counter()
                continue
            if cur is None:
                continue
            m = re.search(r"^\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
            if m:
                op = m.group(1)
                per[cur]["_all"] += 1
                for p in PAT:
                    if op.startswith(p):
                        per[cur][p] += 1
        shown = False
        for fn, c in per.items():
            keys = [p for p in PAT if c[p] and p not in ("MUFU", "F2F", "LDGSTS", "REDUX", "POPC")]
            if not keys:
                continue
            if not shown:
                lines.append(f"== {os.path.basename(obj)}")
                shown = True
            name = demangle(fn)
            lines.append(f"  {name}")
            lines.append("      " + "  ".join(f"{p} x{c[p]}" for p in PAT if c[p]) + f"   ({c['_all']} instructions)")
            for p in PAT:
                total[p] += c[p]
    lines += ["", "== whole library: " + "  ".join(f"{p} x{total[p]}" for p in PAT if total[p])]
    open(os.path.join(ROOT, "profiles", "sass_summary.txt"), "w").write("\n".join(lines) + "\n")
    # ptxas
    out = ["# ptxas -v per kernel (vi-slam_b200/build/*.ptxas.log): registers, spills, static shared memory.  Regenerate: python tools/sass_summary.py", ""]
    for log in sorted(glob.glob(os.path.join(BUILD, "*.ptxas.log"))):
        txt = open(log).read()
        ents = re.findall(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers(?:, used (\d+) barriers)?(?:, (\d+) bytes smem)?", txt, re.S)
        if not ents:
            continue
        out.append(f"== {os.path.basename(log).replace('.ptxas.log', '.cu')}")
        for fn, stack, sst, sld, regs, bars, smem in ents:
            name = demangle(fn)
            out.append(f"  {regs:>3} regs  spill {sst}/{sld} B  stack {stack} B  smem {smem or 0:>6} B   {name}")
    open(os.path.join(ROOT, "profiles", "ptxas_summary.txt"), "w").write("\n".join(out) + "\n")
    print("wrote profiles/sass_summary.txt, profiles/ptxas_summary.txt")


if __name__ == "__main__":
    main()
