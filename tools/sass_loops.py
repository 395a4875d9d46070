#!/usr/bin/env python
"""sass_loops.py — instruction mix of the loops of one kernel, from its SASS (no GPU needed).

  python tools/sass_loops.py <object-or-so> <substring of the mangled kernel name> [--all]

Finds backward branches (loops), and for every loop body [target, branch] counts FP64-pipe
instructions (DFMA/DMUL/DADD/DSETP/DMNMX), FP32, integer/logic/select/move, branches, memory.  The
hot loop of a sweep kernel is the backward branch whose body holds most of the DFMAs; what matters
there is FP64 count (pipe cycles: 2 per warp-instruction per sub-partition on B200) against the
count of everything else (issue slots: 1 per instruction)."""
import re
import subprocess
import sys
from collections import Counter

FP64 = ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")


def kind(m):
    base = m.split(".")[0]
    if base in FP64:
        return "fp64"
    if base in ("MUFU",):
        return "sfu"
    if base.startswith(("FFMA", "FMUL", "FADD", "FSETP", "FSEL", "FMNMX", "HFMA2", "HADD2", "HMUL2")):
        return "fp32"
    if base.startswith(("LD", "ST", "ATOM", "RED", "LDG", "STG", "LDS", "STS", "LDL", "STL", "LDC", "ULDC")):
        return "mem"
    if base in ("BRA", "BSSY", "BSYNC", "EXIT", "RET", "CALL", "BRX", "JMP", "WARPSYNC", "BAR", "NOP", "BREAK", "YIELD"):
        return "ctrl"
    if base.startswith(("F2", "I2", "D2", "F2F")):
        return "cvt"
    return "alu"


def functions(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    cur, body = None, {}
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            body[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and cur:
            addr = int(m.group(1), 16)
            text = m.group(2).strip()
            text = re.sub(r"^@!?U?P\d+\s+", "", text)
            mn = text.split()[0]
            tgt = None
            mt = re.search(r"\b(BRA|BRA\.U|BRA\.DIV)\S*\s+.*?(0x[0-9a-f]+)", text)
            if mn.startswith("BRA"):
                mt = re.search(r"(0x[0-9a-f]+)\s*$", text)
                if mt:
                    tgt = int(mt.group(1), 16)
            body[cur].append((addr, mn, tgt, text))
    return body


def main():
    path, pat = sys.argv[1], sys.argv[2]
    show_all = "--all" in sys.argv
    dump = "--dump" in sys.argv
    for name, ins in functions(path).items():
        if pat not in name:
            continue
        print("==", name, "(%d instructions)" % len(ins))
        loops = [(t, a) for a, mn, t, _ in ins if t is not None and t <= a]
        rows = []
        for t, a in loops:
            c = Counter(kind(mn) for ad, mn, _, _ in ins if t <= ad <= a)
            rows.append((c["fp64"], t, a, c))
        rows.sort(reverse=True)
        for f64, t, a, c in (rows if show_all else rows[:6]):
            tot = sum(c.values())
            print("  loop 0x%05x..0x%05x: %4d instr | fp64 %4d | other %4d (alu %d, mem %d, ctrl %d, cvt %d, fp32 %d, sfu %d)"
                  % (t, a, tot, f64, tot - f64, c["alu"], c["mem"], c["ctrl"], c["cvt"], c["fp32"], c["sfu"]))
            if dump and (f64, t, a, c) == rows[-1 if "--last" in sys.argv else 0]:
                for ad, mn, _, text in ins:
                    if t <= ad <= a:
                        print("      %05x %s" % (ad, text))
        loc = [x for x in ins if x[1].startswith(("LDL", "STL"))]
        print("  local-memory instructions in the whole kernel: %d" % len(loc))


if __name__ == "__main__":
    main()
