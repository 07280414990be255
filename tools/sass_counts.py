"""SASS evidence: per object file, how many tcgen05 / TMA / shuffle instructions the compiled sm_100a
code contains (cuobjdump -sass over build/obj/*.o; mnemonics per /opt/skills/guides/B200_PROFILING.md):
  UTCHMMA = tcgen05.mma (kind::f16; ".2CTA" = cta_group::2), UTCQMMA = tcgen05.mma kind::f8f6f4 (the e5m2 cross-term MMAs
  of the F16Q8 mode), UTMALDG = TMA tensor load, UTMASTG = TMA tensor store, LDTM / STTM = tcgen05.ld / st,
  UTCBAR = tcgen05.commit, SHFL = warp shuffle, MUFU.EX2 = ex2.approx.
Usage: python tools/sass_counts.py > profiles/r2_sass_counts.txt"""
import os
import re
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
OBJ = os.path.join(ROOT, "build", "obj")
PAT = ["UTCHMMA", "UTCHMMA.2CTA", "UTCQMMA", "UTCQMMA.2CTA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "SHFL", "MUFU.EX2", "F2FP.F16", "F2FP.BF16", "LDG.E.128",
       "LDS.128"]


def main():
    if not os.path.isdir(OBJ):
        sys.exit("build first: python __graft_entry__.py")
    print("# " + " ".join(f"{p:>12s}" for p in PAT) + "  object / kernel")
    for f in sorted(os.listdir(OBJ)):
        if not f.endswith(".o"):
            continue
        sass = subprocess.run(["cuobjdump", "-sass", os.path.join(OBJ, f)], capture_output=True, text=True).stdout
        per = {}
        cur = None
        for line in sass.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                cur = m.group(1)
                per[cur] = dict.fromkeys(PAT, 0)
                continue
            if cur:
                for p in PAT:
                    if p in line:
                        per[cur][p] += 1
        tot = {p: sum(v[p] for v in per.values()) for p in PAT}
        print("  " + " ".join(f"{tot[p]:12d}" for p in PAT) + f"  {f} (all kernels)")
        for k, v in per.items():
            if v["UTCHMMA"] or v["UTMALDG"] or (f == "msda.o" and v["SHFL"]):
                name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip().split("(")[0][:70]
                print("  " + " ".join(f"{v[p]:12d}" for p in PAT) + f"    {name}")


if __name__ == "__main__":
    main()
