"""Per-kernel counts of the Blackwell tensor-core / TMA / TMEM SASS mnemonics in libocc_b200.so (cuobjdump -sass).
usage: python scripts/sass_summary.py [out.md]      (runs on the CPU box; no GPU needed)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "occformer_b200", "csrc", "libocc_b200.so")
PATTERNS = [("UTCHMMA / UTC*MMA (tcgen05.mma)", re.compile(r"\bUTC[A-Z]*MMA")), ("UTCBAR (tcgen05.commit)", re.compile(r"\bUTCBAR")),
            ("UTMALDG (TMA load)", re.compile(r"\bUTMALDG")), ("UTMASTG (TMA store)", re.compile(r"\bUTMASTG")),
            ("UTMAREDG (TMA reduce)", re.compile(r"\bUTMAREDG")), ("LDTM (tcgen05.ld)", re.compile(r"\bLDTM")),
            ("STTM (tcgen05.st)", re.compile(r"\bSTTM")), ("LDGSTS (cp.async)", re.compile(r"\bLDGSTS")),
            ("SYNCS (mbarrier)", re.compile(r"\bSYNCS"))]


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else None
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts = collections.OrderedDict()
    cur = None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(.*", "", cur)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for name, pat in PATTERNS:
            if pat.search(line):
                counts[cur][name] += 1
    names = [n for n, _ in PATTERNS]
    lines = ["# SASS summary of libocc_b200.so (sm_100a): tensor-core / TMA / TMEM instructions per kernel", "",
             "`cuobjdump -sass occformer_b200/csrc/libocc_b200.so`, counted by scripts/sass_summary.py; kernels without any of",
             "these instructions (SIMT kernels) are listed at the end.", "",
             "| kernel | " + " | ".join(names) + " |", "|---|" + "---:|" * len(names)]
    simt = []
    tot = collections.Counter()
    for k, c in counts.items():
        if not any(c[n] for n in names[:7]):
            simt.append(k)
            continue
        tot.update(c)
        lines.append(f"| `{k}` | " + " | ".join(str(c[n]) for n in names) + " |")
    lines.append("| **total** | " + " | ".join(str(tot[n]) for n in names) + " |")
    lines += ["", f"SIMT kernels ({len(simt)}): " + ", ".join(f"`{k}`" for k in simt)]
    text = "\n".join(lines) + "\n"
    if out:
        open(out, "w").write(text)
    print(text)


if __name__ == "__main__":
    main()
