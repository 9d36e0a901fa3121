"""SASS opcode histogram of libt2h.so: whole library, then the Blackwell-specific opcodes per kernel.

    python tools/sass_histogram.py > profiles/r02_sass_opcodes.txt      (no GPU needed)
"""
import collections
import re
import subprocess
import sys

sys.path.insert(0, ".")
from tomosar2height_b200._lib import LIB_PATH  # noqa: E402

SPECIAL = ("UTCHMMA", "UTCQMMA", "UTCBAR", "UTMALDG", "UTMASTG", "LDTM", "STTM", "ELECT", "SYNCS", "LDGSTS", "UTCCP")
sass = subprocess.run(["cuobjdump", "-sass", LIB_PATH], capture_output=True, text=True, check=True).stdout
total, per_kernel, name, kernels = collections.Counter(), {}, None, 0
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = m.group(1)
        per_kernel[name] = collections.Counter()
        kernels += 1
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and name:
        total[m.group(1)] += 1
        per_kernel[name][m.group(1)] += 1
demangle = subprocess.run(["cu++filt"], input="\n".join(per_kernel), capture_output=True, text=True).stdout.splitlines()
print("# SASS opcode histogram of tomosar2height_b200/libt2h.so (sm_100a), round 2 (tools/sass_histogram.py)")
print("# cuobjdump -sass libt2h.so | opcode mnemonics (modifiers stripped), whole library then per kernel (tensor / TMA / TMEM opcodes only)")
print("# tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG, cp.async -> LDGSTS (see /opt/skills/guides/B200_PROFILING.md)\n")
print(f"## whole library: {kernels} kernels, {sum(total.values())} instructions")
for op, n in total.most_common(60):
    print(f"{n:8d}  {op}")
print("\n## Blackwell-specific opcodes per kernel")
for (mangled, ops), pretty in zip(per_kernel.items(), demangle):
    hits = {op: n for op, n in sorted(ops.items()) if op in SPECIAL}
    if any(op.startswith("UTC") or op.startswith("UTMA") for op in hits):
        print(pretty[:200])
        print("    " + "  ".join(f"{op}={n}" for op, n in hits.items()))
