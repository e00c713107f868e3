"""Per-kernel counts of the SASS instructions that prove the Blackwell-native paths (B200_PROFILING.md): tcgen05.mma ->
UTC*MMA, tcgen05.ld -> LDTM, cp.async.bulk.tensor -> UTMALDG / UTMASTG, cp.async.bulk -> UBLKCP, mma.sync -> HMMA, cp.async ->
LDGSTS.  usage: python tools/sass_evidence.py > profiles/rNN_sass_evidence.txt   (runs cuobjdump -sass on the built .so)"""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "atspeed_b200", "libatspeed_b200.so")
PAT = {"UTC*MMA (tcgen05.mma)": r"\bUTC[A-Z]*MMA", "LDTM (tcgen05.ld)": r"\bLDTM", "UTMALDG (TMA tensor load)": r"\bUTMALDG",
       "UTMASTG (TMA tensor store)": r"\bUTMASTG", "UBLKCP (bulk copy)": r"\bUBLKCP", "UTCBAR (tcgen05.commit)": r"\bUTCBAR",
       "SYNCS (mbarrier)": r"\bSYNCS", "UCGABAR (cluster barrier)": r"\bUCGABAR", "HMMA (mma.sync)": r"\bHMMA",
       "LDSM (ldmatrix)": r"\bLDSM", "LDGSTS (cp.async)": r"\bLDGSTS"}


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", txt)[1:]
    print("# `cuobjdump -sass atspeed_b200/libatspeed_b200.so` (sm_100a): marker-instruction counts per kernel; kernels with no")
    print("# marker are plain SIMT (row-wise, beam tree, top-k).  Legend: " + "; ".join(PAT))
    rows = []
    for f in funcs:
        mangled = f.split("\n", 1)[0].strip()
        dem = subprocess.run(["cu++filt", mangled], capture_output=True, text=True).stdout.strip() or mangled
        dem = dem.replace("(int)", "").replace("atspeed::", "").replace("void ", "")
        dem = re.split(r"\((?![^<]*>)", dem)[0]
        n_inst = len(re.findall(r"^\s+/\*[0-9a-f]{4,5}\*/", f, flags=re.M))
        rows.append((dem, n_inst, {k: len(re.findall(v, f)) for k, v in PAT.items()}))
    for dem, n, c in sorted(rows, key=lambda r: (-sum(r[2].values()), r[0])):
        hits = ", ".join(f"{k.split(' ')[0]}={v}" for k, v in c.items() if v)
        print(f"{dem[:60]:<60} {n:>6} instr  {hits}")


if __name__ == "__main__":
    main()
