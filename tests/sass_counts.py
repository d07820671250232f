"""Per-kernel SASS mnemonic counts of libaadff.so (CPU box; cuobjdump only):  python tests/sass_counts.py > profiles/r02_sass_counts.txt
Evidence that the shipped kernels are Blackwell-native: UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UBLKCP = cp.async.bulk,
UTCBAR = tcgen05.commit, UTCATOMSWS = tcgen05.alloc; HMMA (legacy mma.sync) must be absent."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "aberration-aware-depth-from-focus_b200", "libaadff.so")
WATCH = ["UTCHMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "UBLKCP", "UTMALDG", "SYNCS", "LDGSTS", "HMMA", "FFMA", "MUFU", "F2FP",
         "LDS", "STS", "LDG", "STG", "SHFL", "BAR", "ATOMG", "RED"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
    kernels, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = demangle(m.group(1))
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            kernels[cur][m.group(1).split(".")[0]] += 1
            kernels[cur]["_total"] += 1
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} | per-kernel mnemonic counts (sm_100a)")
    print("| kernel | instr | " + " | ".join(WATCH) + " |")
    print("|---|---|" + "---|" * len(WATCH))
    tot = collections.Counter()
    for name, c in kernels.items():
        short = name.replace("void aadff::", "")
        short = short[:short.rindex(">(") + 1] if ">(" in short else short.split("(")[0]
        short = short.replace("(int)", "").replace("(bool)", "")
        print(f"| {short} | {c['_total']} | " + " | ".join(str(c.get(w, 0)) for w in WATCH) + " |")
        tot.update(c)
    print(f"| ALL ({len(kernels)} kernels) | {tot['_total']} | " + " | ".join(str(tot.get(w, 0)) for w in WATCH) + " |")


if __name__ == "__main__":
    main()
