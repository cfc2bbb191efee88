"""SASS evidence for the Blackwell-native paths: per kernel of libesmdiff_b200.so, the counts of the
tensor-core / TMEM / TMA mnemonics (B200_PROFILING.md "What proves a Blackwell-native kernel") and of the
packed-fp32 / 3-input-max instructions, plus the 12 most frequent opcodes.

    python tools/sass_hist.py > profiles/r2_sass_opcodes.md        (no GPU needed: cuobjdump -sass)
"""
import collections
import re
import subprocess
import sys
from pathlib import Path

LIB = Path(__file__).resolve().parent.parent / "esmdiff_b200" / "lib" / "libesmdiff_b200.so"
KEY = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "HMMA", "FFMA2", "FADD2", "FMUL2",
       "FMNMX3", "MUFU.EX2", "SYNCS", "ACQBULK", "UCGABAR_ARV", "UCGABAR_WAIT"]


def main():
    txt = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in txt.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(.*", "", cur).replace("void ", "").replace("esmdiff::", "")
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur:
            kernels[cur][m.group(1)] += 1
    print(f"`cuobjdump -sass {LIB.name}` (sm_100a), instruction counts per kernel.  UTCHMMA = tcgen05.mma, LDTM / STTM = "
          "tcgen05.ld / st, UTMALDG / UTMASTG = TMA tile load / store, SYNCS = mbarrier ops, ACQBULK = griddepcontrol.wait "
          "(programmatic dependent launch), UCGABAR = cluster barrier (CTA pairs); no HMMA (legacy mma.sync) anywhere.\n")
    print("| kernel | instructions | " + " | ".join(KEY) + " |")
    print("|---|---:|" + "---:|" * len(KEY))
    for name, c in kernels.items():
        tot = sum(c.values())
        if not name or not tot:
            continue
        cnt = [sum(v for k, v in c.items() if k.startswith(key)) for key in KEY]
        print(f"| `{name}` | {tot} | " + " | ".join(str(x) if x else "" for x in cnt) + " |")
    print("\nMost frequent opcodes of the three hot kernels:\n")
    for name, c in kernels.items():
        if any(t in name for t in ("gemm_bf16_tn_kernel<8", "gemm_bf16_tn_kernel<6", "attention_resident")):
            print(f"* `{name}`: " + ", ".join(f"{k} {v}" for k, v in c.most_common(12)))


if __name__ == "__main__":
    sys.exit(main())
