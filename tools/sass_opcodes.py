"""Per-kernel counts of the SASS mnemonics that prove the Blackwell paths (B200_PROFILING.md): tcgen05.mma -> UTC*MMA,
tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG/UTMAREDG/UBLKCP/UBLKRED, legacy tensor path -> HMMA.

    python tools/sass_opcodes.py > profiles/sass_opcodes_r2.txt
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gamer_b200", "lib", "libgamer_b200.so")
OPS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "UBLKRED", "UTCBAR", "USETMAXREG",
       "HMMA", "MUFU.EX2"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(.*", "", name.replace("(anonymous namespace)::", "").replace("void ", ""))
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for op in OPS:
            if re.search(r"(?<![A-Z])" + re.escape(op) + r"\b", line):      # (HMMA must not match inside UTCHMMA)
                counts[cur][op] += 1
    print(f"# {os.path.relpath(LIB, ROOT)}: SASS mnemonic counts per kernel (cuobjdump -sass), kernels without any of them omitted")
    print(f"# {'kernel':88s} " + " ".join(f"{o:>9s}" for o in OPS))
    for k, c in counts.items():
        if sum(c.values()) == 0:
            continue
        print(f"{k[:90]:90s} " + " ".join(f"{c.get(o, 0):9d}" for o in OPS))


if __name__ == "__main__":
    main()
