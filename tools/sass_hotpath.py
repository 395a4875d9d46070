#!/usr/bin/env python
"""sass_hotpath.py — instruction mix along the fall-through path of one loop of a kernel (no GPU needed).

  python tools/sass_hotpath.py <object-or-so> <substring of the mangled kernel name> <loop start address (hex)>

Walks the SASS from the loop head: conditional branches are assumed not taken (the compilers place the rarely
taken tiers out of line), unconditional branches are followed, until the walk returns to the head.  Prints the
FP64 / other split of that path and, with -v, the path itself."""
import re
import subprocess
import sys

FP64 = ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")


def main():
    path, name, start = sys.argv[1], sys.argv[2], int(sys.argv[3], 16)
    verbose = "-v" in sys.argv
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    cur, ins = None, {}
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and cur and name in cur:
            ins[int(m.group(1), 16)] = m.group(2).strip()
    addrs = sorted(ins)
    nxt = {a: b for a, b in zip(addrs, addrs[1:])}
    pc, seen, counts, other = start, set(), {}, {}
    nfp = noth = 0
    while pc in ins and pc not in seen:
        seen.add(pc)
        text = ins[pc]
        pred = re.match(r"^@!?U?P\d+\s+", text)
        body = text[pred.end():] if pred else text
        mn = body.split()[0]
        base = mn.split(".")[0]
        if base in FP64:
            nfp += 1
        else:
            noth += 1
            other[base] = other.get(base, 0) + 1
        if verbose:
            print("%05x %s" % (pc, text))
        if base == "BRA" and not pred:
            tgt = int(re.search(r"0x([0-9a-f]+)", body).group(1), 16)
            if tgt == start:
                break
            pc = tgt
            continue
        if base == "BRA" and pred:
            tgt = int(re.search(r"0x([0-9a-f]+)", body).group(1), 16)
            if tgt == start:          # the loop's back edge
                break
        if base in ("EXIT", "RET"):
            break
        pc = nxt.get(pc)
    print("fall-through path from 0x%x: %d instructions, FP64 %d, other %d  %s" % (
        start, nfp + noth, nfp, noth, sorted(other.items(), key=lambda kv: -kv[1])))


if __name__ == "__main__":
    main()
