"""Per-kernel SASS mnemonic counts of the shipped library (cuobjdump -sass), the evidence behind DESIGN.md's claims
about which kernels use TMA (UTMALDG / UBLKCP), mbarriers (SYNCS) and packed FP32 FMAs (FFMA2).

    python scripts/sass_excerpt.py > profiles/r02_sass_excerpt.txt
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gym_sted_b200", "libsted_b200.so")
COLS = [("FFMA2", r"\bFFMA2\b"), ("FFMA", r"\bFFMA\b"), ("DFMA", r"\bDFMA\b"), ("UTMALDG", r"UTMALDG"), ("UTMASTG", r"UTMASTG"),
        ("UBLKCP", r"UBLKCP"), ("SYNCS", r"SYNCS"), ("BAR", r"\bBAR\b"), ("ATOMS", r"\bATOMS\b"), ("RED/ATOMG", r"\bRED\b|\bATOMG\b"),
        ("MUFU", r"\bMUFU"), ("LDS", r"\bLDS\b"), ("STL/LDL", r"\bSTL\b|\bLDL\b")]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    parts = re.split(r"\n\s*Function : ", sass)[1:]
    rows = []
    for p in parts:
        name = p.split("\n", 1)[0].strip()
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        dem = dem.replace("(anonymous namespace)::", "")
        dem = (dem[5:] if dem.startswith("void ") else dem).split("(")[0]
        insts = len(re.findall(r"^\s*/\*[0-9a-f]+\*/\s+\S", p, flags=re.M))
        rows.append((dem, [len(re.findall(rx, p)) for _, rx in COLS], insts))
    print("# cuobjdump -sass gym_sted_b200/libsted_b200.so (sm_100a): mnemonic counts per kernel")
    print(f"{'kernel':46s}" + "".join(f"{c:>10s}" for c, _ in COLS) + f"{'insts':>8s}")
    for dem, counts, insts in sorted(rows):
        print(f"{dem[:46]:46s}" + "".join(f"{c:10d}" for c in counts) + f"{insts:8d}")


if __name__ == "__main__":
    sys.exit(main())
