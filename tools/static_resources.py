"""Static evidence that needs no GPU: per-kernel registers / spills / shared memory
from the ptxas logs of the build (build/*.ptxas.log, written by
tramp_b200/csrc/Makefile) and the counts of the SASS instructions that prove which
hardware paths the kernels use (B200_PROFILING.md: UBLKCP / UTMALDG = TMA, DMMA =
FP64 tensor pipe, UCGABAR = cluster barrier, SYNCS = mbarrier).

    python tools/static_resources.py > profiles/r01_static_resources.md
"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MNEMONICS = ("UBLKCP", "UTMALDG", "DMMA", "DFMA", "UCGABAR", "SYNCS", "MEMBAR", "LDG.E.128", "LDS.128", "ATOM", "RED")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def ptxas_rows():
    rows = []
    for log in sorted(glob.glob(os.path.join(ROOT, "build", "*.ptxas.log"))):
        text = open(log).read()
        for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n(?:.*\n)*?.*?(\d+) bytes stack frame, "
                             r"(\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers(.*)", text):
            name, stack, st, ld, regs, rest = m.groups()
            smem = re.search(r"(\d+) bytes smem", rest)
            rows.append((os.path.basename(log).replace(".ptxas.log", ".cu"), name, int(regs), int(stack), int(st),
                         int(ld), int(smem.group(1)) if smem else 0))
    return rows


def sass_counts():
    counts = {}
    for obj in sorted(glob.glob(os.path.join(ROOT, "build", "*.o"))):
        sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
        kernel = None
        for line in sass.split("\n"):
            m = re.match(r"\s*Function : (\S+)", line)
            if m:
                kernel = m.group(1)
                counts[kernel] = collections.Counter()
                continue
            if kernel:
                for mn in MNEMONICS:
                    if re.search(r"\b" + re.escape(mn), line):
                        counts[kernel][mn] += 1
    return counts


def short(name, limit=88):
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = re.sub(r"\(.*\)$", "", name)
    return name if len(name) <= limit else name[:limit - 1] + "…"


def main():
    rows = ptxas_rows()
    counts = sass_counts()
    names = demangle(sorted({r[1] for r in rows} | set(counts)))
    print("# Static resources of the sm_100a build (no GPU needed)\n")
    print("`python tools/static_resources.py`, from `build/*.ptxas.log` (`-Xptxas -v`) and `cuobjdump -sass build/*.o`.\n")
    print("## ptxas: registers, stack, spills, static shared memory\n")
    print("| file | kernel | registers | stack B | spill st/ld B | static smem B |\n|---|---|---|---|---|---|")
    for f, name, regs, stack, st, ld, smem in rows:
        print(f"| {f} | `{short(names[name])}` | {regs} | {stack} | {st}/{ld} | {smem} |")
    spilling = [r for r in rows if r[4] or r[5]]
    print(f"\n{len(rows)} kernels, {len(spilling)} with register spills.\n")
    print("## SASS: instructions that identify the hardware path\n")
    print("| kernel | " + " | ".join(MNEMONICS) + " |\n|---|" + "---|" * len(MNEMONICS))
    for kernel in sorted(counts, key=lambda k: names[k]):
        c = counts[kernel]
        if any(c[m] for m in ("UBLKCP", "UTMALDG", "DMMA", "UCGABAR", "SYNCS")):
            print(f"| `{short(names[kernel])}` | " + " | ".join(str(c[m]) if c[m] else "" for m in MNEMONICS) + " |")
    print("\n(kernels without TMA, DMMA, cluster-barrier or mbarrier instructions are left out of the second table)")


if __name__ == "__main__":
    main()
