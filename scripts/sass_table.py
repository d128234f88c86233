#!/usr/bin/env python
"""Per-kernel SASS evidence table: counts of the Blackwell-specific mnemonics in the built library.

    python scripts/sass_table.py > profiles/sass_r02.md

Runs `cuobjdump -sass` on neurosync_trainer_lite_b200/_lib/libnsf.so (no GPU needed) and counts, per kernel,
UTC*MMA (tcgen05.mma), UTMALDG / UTMASTG (TMA tensor copies), UBLKCP (cp.async.bulk), LDTM / STTM (tcgen05.ld / st),
HMMA (mma.sync, the legacy tensor path), LDSM (ldmatrix), plus the instruction total."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "neurosync_trainer_lite_b200", "_lib", "libnsf.so")
PATTERNS = [("UTC*MMA", r"\bUTC[A-Z]*MMA"), ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"), ("UBLKCP", r"\bUBLKCP"),
            ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("HMMA", r"\bHMMA"), ("LDSM", r"\bLDSM"), ("SYNCS", r"\bSYNCS"),
            ("FFMA", r"\bFFMA"), ("DFMA", r"\bDFMA")]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def short(name):
    name = re.sub(r"nsf::\(anonymous namespace\)::|nsf::<unnamed>::|nsf::", "", name)
    name = re.sub(r"^void ", "", name)
    return re.sub(r"\(.*$", "", name)


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    parts = re.split(r"\n\s*Function : ", txt)[1:]
    rows = []
    names = [p.split("\n", 1)[0].strip() for p in parts]
    dm = demangle(names)
    for raw, body in zip(names, parts):
        n_inst = len(re.findall(r"/\*[0-9a-f]{4,5}\*/", body))
        counts = [len(re.findall(pat, body)) for _, pat in PATTERNS]
        rows.append((short(dm.get(raw, raw)), n_inst, counts))
    rows.sort(key=lambda r: r[0])
    print("# SASS evidence, round 2 (`python scripts/sass_table.py`; `cuobjdump -sass` of `_lib/libnsf.so`, sm_100a)\n")
    print("Counts of the mnemonics that identify the execution path of each kernel: `UTC*MMA` = `tcgen05.mma`,")
    print("`UTMALDG` = TMA tensor load, `UBLKCP` = `cp.async.bulk`, `LDTM` = `tcgen05.ld`, `HMMA` = `mma.sync` (legacy tensor")
    print("path), `LDSM` = `ldmatrix`, `SYNCS` = mbarrier operations.  Kernels without any of them are CUDA-core kernels.\n")
    print("| kernel | instructions | " + " | ".join(n for n, _ in PATTERNS) + " |")
    print("|---|---|" + "---|" * len(PATTERNS))
    for name, n_inst, counts in rows:
        print(f"| `{name}` | {n_inst} | " + " | ".join(str(c) if c else "" for c in counts) + " |")


if __name__ == "__main__":
    sys.exit(main())
