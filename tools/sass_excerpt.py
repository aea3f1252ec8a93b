#!/usr/bin/env python
"""Evidence tool: SASS of the backtrace kernels' inner loops out of the built library (cuobjdump -sass), with the instruction mix.

    python tools/sass_excerpt.py > profiles/r02_sass_inner_loops.txt

For each listed instantiation: the whole-kernel counts of the mnemonics that prove the design (UBLKCP = TMA bulk copy, SYNCS =
mbarrier, DFMA/DADD/DMUL = FP64 pipe, LDS/LDG, SHFL) and the hottest backward-branch loop -- the body with the most FP64
instructions -- printed verbatim: the two-levels-per-trip full-kick loop of the trace.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "numericalflowiteration_b200", "lib", "libnufi_b200.so")
WANT = [  # (template arguments <DIM, ILP, STAGED, POW2, XPP, ORDER>, what it is)
    ("1, 2, true, true, false, 4", "C1/C2: 1d, 2 points per thread, TMA-staged per-cell quadratics"),
    ("2, 2, true, true, true, 4", "C3: 2d, 2 points per thread, TMA-staged xpp levels"),
    ("3, 1, true, true, true, 4", "C4: 3d, TMA-staged xpp levels"),
    ("3, 1, true, true, false, 4", "C5-16: 3d, TMA-staged B-spline levels"),
    ("3, 1, false, true, false, 4", "C5-32/64: 3d, B-spline levels read through L1/L2"),
    ("3, 1, false, false, false, 6", "generic order 6, 3d (Cox-de Boor basis in registers)"),
]
MNEMONICS = ["DFMA", "DADD", "DMUL", "LDS", "LDG", "UBLKCP", "SYNCS", "SHFL", "BAR", "IMAD", "IADD3", "LOP3", "SEL", "MUFU", "F2I", "I2F", "BRA"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    funcs = re.split(r"\n\s*Function : ", sass)[1:]
    by_name = {}
    for f in funcs:
        mangled = f.split("\n", 1)[0].strip()
        dem = subprocess.run(["c++filt", mangled], capture_output=True, text=True).stdout.strip()
        m = re.search(r"backtrace_kernel<([^>]*)>", dem)
        if m:
            by_name[m.group(1)] = f
    print(f"# SASS excerpts of libnufi_b200.so (sm_100a), made by tools/sass_excerpt.py\n")
    for args, what in WANT:
        f = by_name.get(args)
        if f is None:
            print(f"## backtrace_kernel<{args}>: not in the library\n")
            continue
        lines = [ln for ln in f.splitlines() if re.search(r"/\*[0-9a-f]{4}\*/", ln)]
        insts = []
        for ln in lines:
            m = re.search(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m:
                insts.append((int(m.group(1), 16), m.group(2).strip()))
        mix = collections.Counter()
        for _, t in insts:
            op = re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0]
            mix[op] += 1
        print(f"## backtrace_kernel<{args}>  -- {what}")
        print(f"whole kernel: {len(insts)} instructions; " + ", ".join(f"{k} {mix[k]}" for k in MNEMONICS if mix[k]))
        # loops = backward branches; the hot one = the body that loads (LDS/LDG) with the highest FP64 density
        best = None
        addr_index = {a: i for i, (a, _) in enumerate(insts)}
        for i, (a, t) in enumerate(insts):
            m = re.search(r"BRA(?:\.\w+)*\s+(?:!?U?P\d+,\s*)?(?:`\()?(?:0x)?([0-9a-f]+)\)?", t)
            if "BRA" not in t or not m:
                continue
            try:
                tgt = int(m.group(1), 16)
            except ValueError:
                continue
            if tgt >= a or tgt not in addr_index:
                continue
            body = insts[addr_index[tgt]:i + 1]
            if len(body) > 1000 or not any(re.search(r"\b(LDS|LDG)", x) for _, x in body):
                continue
            n64 = sum(1 for _, x in body if re.match(r"(@!?U?P\d+\s+)?D(FMA|ADD|MUL)", x))
            nfma = sum(1 for _, x in body if re.match(r"(@!?U?P\d+\s+)?DFMA", x))
            if n64 >= 20 and nfma >= 8 and (best is None or n64 / len(body) > best[0] / len(best[1])):  # densest in FP64 = innermost trace loop
                best = (n64, body)
        if best:
            body = best[1]
            bm = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for _, t in body)
            print(f"hot loop ({len(body)} instructions = the full-kick trace loop, unrolled over levels and the points of a thread): " +
                  ", ".join(f"{k} {bm[k]}" for k in MNEMONICS if bm[k]))
            for a, t in (body if "--full" in sys.argv else body[:120]):
                print(f"    /*{a:04x}*/  {t}")
            if len(body) > 120 and "--full" not in sys.argv:
                print(f"    ... ({len(body) - 120} more instructions of the same loop body; run tools/sass_excerpt.py --full for all)")
        print()


if __name__ == "__main__":
    main()
