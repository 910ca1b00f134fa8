"""Static SASS statistics of librenv_b200.so (no GPU needed):

    python profiles/sass_stats.py [lib.so] [function-substring] [--loops]

Per kernel: instruction count, opcode histogram of the interesting classes, and (with --loops) every backward
branch with the length of the loop body it closes -- the quickest way to see what a source change did to the
inner loop of the fused rollout or of the sampler before spending GPU time.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INSTR = re.compile(r"^\s+/\*([0-9a-f]{4,})\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)(.*?);")


def functions(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True, check=True).stdout
    name, rows, result = None, [], {}
    for line in out.splitlines():
        if "Function :" in line:
            if name:
                result[name] = rows
            name, rows = line.split("Function :")[1].strip(), []
            continue
        m = INSTR.match(line)
        if m and name:
            rows.append((int(m.group(1), 16), m.group(3), m.group(4)))
    if name:
        result[name] = rows
    return result


def summarize(name, rows, loops):
    ops = collections.Counter(op.split(".")[0] for _, op, _ in rows)
    wide = sum(1 for _, op, _ in rows if op.startswith(("LDG.E.128", "STG.E.128")))
    local = sum(1 for _, op, _ in rows if op.startswith(("LDL", "STL")))
    print("%s\n  %d instructions; 128-bit global ld/st %d; local-memory ld/st %d" % (name, len(rows), wide, local))
    print("  " + ", ".join("%s %d" % kv for kv in ops.most_common(14)))
    if loops:
        for addr, op, rest in rows:
            if op.startswith("BRA"):
                m = re.search(r"0x([0-9a-f]+)", rest)
                if m and int(m.group(1), 16) <= addr:
                    lo = int(m.group(1), 16)
                    body = [r for r in rows if lo <= r[0] <= addr]
                    h = collections.Counter(o.split(".")[0] for _, o, _ in body)
                    print("  loop 0x%x..0x%x: %d instr: %s" % (lo, addr, len(body), ", ".join("%s %d" % kv for kv in h.most_common(8))))


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    lib = args[0] if args else os.path.join(ROOT, "random_envs_b200", "librenv_b200.so")
    want = args[1] if len(args) > 1 else ""
    for name, rows in functions(lib).items():
        if want in name:
            summarize(name, rows, "--loops" in sys.argv)


if __name__ == "__main__":
    main()
