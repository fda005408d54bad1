"""Mnemonic counts and a few excerpt lines per kernel of libasw.so:
    python profiles/sass_counts.py stft_cc_rr_kernel srp_gather_ws_kernel gcc_fft_kernel > profiles/r02b_sass_counts.txt"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "acousticswarms_speech_b200", "lib", "libasw.so")
PICK = ("LDGSTS", "UBLKCP", "SYNCS", "BAR.SYNC", "LDS", "STS", "LDG.E", "STG.E", "LDL", "STL", "FADD2", "FMUL2", "FFMA2", "FADD", "FMUL",
        "FFMA", "SHFL", "MUFU", "UTMALDG", "UTCHMMA")
SHOW = ("UBLKCP", "SYNCS", "FFMA2", "LDGSTS", "STG.E.128", "LDS.128")


def main(names):
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    fn, counts, shown, total = None, None, None, 0
    res = []
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if fn:
                res.append((fn, total, counts, shown))
            fn = m.group(1) if any(n in m.group(1) for n in names) else None
            counts, shown, total = collections.Counter(), {}, 0
            continue
        if not fn:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if not m:
            continue
        total += 1
        op = m.group(1)
        for k in PICK:
            if op == k or op.startswith(k + "."):
                counts[k] += 1
                break
        for k in SHOW:
            if op.startswith(k) and k not in shown:
                shown[k] = line.split("/*", 2)[1].split("*/", 1)[1].strip()
    if fn:
        res.append((fn, total, counts, shown))
    print("# SASS mnemonic counts (cuobjdump -sass acousticswarms_speech_b200/lib/libasw.so, sm_100a): kernel | instructions | mnemonics")
    print("# UBLKCP = cp.async.bulk (1-D TMA bulk copy), SYNCS = mbarrier, LDGSTS = cp.async, F*2 = packed fp32x2, LDL/STL = spills")
    for fn, total, counts, shown in sorted(res):
        short = re.sub(r"^_ZN3asw\d+_GLOBAL__N__[0-9a-f]+_\d+_\w+?_cu_[0-9a-f]+\d\d", "", fn)
        print(f"{short[:70]:72s} {total:6d}  " + " ".join(f"{k}={v}" for k, v in counts.items()))
        for k, l in shown.items():
            print(f"      {l}")


if __name__ == "__main__":
    main(sys.argv[1:])
