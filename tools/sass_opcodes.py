"""Per-kernel counts of the SASS opcodes that prove tcgen05 / TMEM / TMA in the in-tree library.

    python tools/sass_opcodes.py > profiles/r02_sass_opcodes.txt

cuobjdump -sass on fun_asr_gguf_b200/libfunasr_b200.so; UTCHMMA = tcgen05.mma (kind::f16), UTCQMMA = kind::f8f6f4,
LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA tensor load / store, UBLKCP = bulk copy, SYNCS = mbarrier.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "fun_asr_gguf_b200", "libfunasr_b200.so")
OPS = ["UTCHMMA", "UTCQMMA", "UTCOMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "HMMA", "FFMA", "MUFU"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, order, cur = collections.defaultdict(collections.Counter), [], None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", "-p", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            order.append(cur)
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            op = m.group(1)
            counts[cur]["_total"] += 1
            for o in OPS:
                if op == o or op.startswith(o + "."):
                    counts[cur][o] += 1
    print("# SASS opcode counts per kernel, " + os.path.relpath(LIB, ROOT) + " (sm_100a); columns: " + " ".join(OPS) + " | total")
    tot = collections.Counter()
    for k in order:
        c = counts[k]
        tot.update(c)
        print(f"{k[:90]:90s} " + " ".join(f"{c[o]:5d}" for o in OPS) + f" | {c['_total']:6d}")
    print(f"{'TOTAL':90s} " + " ".join(f"{tot[o]:5d}" for o in OPS) + f" | {tot['_total']:6d}")


if __name__ == "__main__":
    sys.exit(main())
