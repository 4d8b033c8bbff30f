#!/usr/bin/env python
"""Static evidence for kernels that have not been profiled on hardware yet: per-stage instruction mix of the
contraction kernels' main loops (from `cuobjdump -sass` of the built library) and ptxas' register / spill report.

    python tools/sass_mix.py > profiles/rNN/contraction_sass_mix.txt
"""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "easydistillation_b200", "libedk_sm100a.so")
KEEP = ("DMMA", "DFMA", "DMUL", "DADD", "LDS", "LDL", "STL", "UTMALDG", "UBLKCP", "SYNCS", "NOP")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    funcs, cur = {}, None
    for ln in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);", ln)
        if cur and m:
            funcs[cur].append((int(m.group(1), 16), m.group(2)))
    print("# main (stage) loop of every contraction kernel: the innermost backward branch that contains DMMAs")
    print("# one stage = 8 sites (gram_tma / gram_pw) or 8 site pairs = 16 sites (gram_pwf); counts are per warp and stage")
    print("# gram_pw / gram_pwf loops hold TWO copies of the stage body (full tiles: unguarded; edge tiles: f-blocks past Ne skipped),")
    print("# so their counts are twice what one stage executes")
    for name in sorted(funcs):
        if not re.search(r"gram_(pw|pwf|tma)_kernel", name):
            continue
        ins = funcs[name]
        best = None
        for a, t in ins:
            m = re.search(r"BRA\S*\s+(?:.*?)0x([0-9a-f]+)", t)
            if m and int(m.group(1), 16) < a:
                body = [x for x in ins if int(m.group(1), 16) <= x[0] <= a]
                ops = collections.Counter(re.sub(r"@!?U?P\d+\s+", "", x[1]).split()[0] for x in body)
                if any(k.startswith("DMMA") for k in ops) and (best is None or len(body) < best[0]):
                    best = (len(body), ops)
        if best is None:
            continue
        n, ops = best
        mix = {k: v for k, v in sorted(ops.items()) if k.split(".")[0] in KEEP}
        demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().split("(")[0]
        print(f"{demangled}: {n} instructions, {mix}")
    print()
    print("# separable contraction (form 4): whole-kernel instruction counts of the 13-mode, 8-pairs-per-stage instances.")
    print("# The kernel holds four copies of the stage body (1 or 2 blocks of L x 1 or 2 blocks of R per lane), each fully")
    print("# unrolled over the 8 site pairs of a stage; LDTM / STTM are tcgen05.ld / tcgen05.st (y accumulators in tensor")
    print("# memory), UTMALDG the TMA boxes, UTCATOMSWS / UTCBAR the tensor-memory allocation, no DMMA anywhere.")
    sep_keep = ("DFMA", "DADD", "DMUL", "LDS", "LDC", "LDTM", "STTM", "UTMALDG", "SYNCS", "LDL", "STL", "IMAD", "MOV", "LDG", "STG", "UTCATOMSWS", "UTCBAR")
    for name in sorted(funcs):
        if not re.search(r"gram_sepx1?_kernelILi2ELi4ELi8E", name):
            continue
        ops = collections.Counter(re.sub(r"@!?U?P\d+\s+", "", x[1]).split()[0].split(".")[0] for x in funcs[name])
        demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().split("(")[0]
        print(f"{demangled}: {len(funcs[name])} instructions, {({k: ops[k] for k in sep_keep if ops.get(k)})}")
    print()
    print("# ptxas -v (registers at launch; the MMA warps raise their budget to 232 with setmaxnreg)")
    for log in ("edk_gram.o.log", "edk_gram_pw.o.log", "edk_gram_sep_s24.o.log"):
        path = os.path.join(REPO, "easydistillation_b200", "build", log)
        if not os.path.exists(path):
            continue
        lines = open(path).read().splitlines()
        for i, ln in enumerate(lines):
            m = re.search(r"Compiling entry function '(\S+)'", ln)
            if m and re.search(r"gram_(pw|pwf|tma|sepx|sepx1)_kernel", m.group(1)):
                d = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
                print(d + ":", " | ".join(x.replace("ptxas info    : ", "").strip() for x in lines[i + 2:i + 4]))


if __name__ == "__main__":
    sys.exit(main())
