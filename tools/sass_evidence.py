"""Static evidence that the kernels are Blackwell-native: per kernel of libdreamlab_b200.so, the
count of the SASS mnemonics B200_PROFILING.md lists (tcgen05.mma -> UTC*MMA, tcgen05.ld/st ->
LDTM/STTM, TMA -> UTMALDG/UTMASTG/UBLKCP; legacy HMMA must be absent).  Runs without a GPU:
  python tools/sass_evidence.py > profiles/r01_sass_evidence.txt"""
import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATTERNS = [("UTC*MMA", r"\bUTC[A-Z]*MMA\b"), ("LDTM", r"\bLDTM\b"), ("STTM", r"\bSTTM\b"),
            ("UTMALDG", r"\bUTMALDG\b"), ("UTMASTG", r"\bUTMASTG\b"), ("UBLKCP", r"\bUBLKCP\b"),
            ("SYNCS", r"\bSYNCS\b"), ("MUFU", r"\bMUFU\b"), ("HMMA", r"\bHMMA\b"), ("LDGSTS", r"\bLDGSTS\b")]


def main():
    sos = glob.glob(os.path.join(ROOT, "stable-diffusion-1.5-lcm-onnx-rknn2_b200", "*.so"))
    if not sos:
        sys.exit("build first: python -c 'import __graft_entry__ as g; g.build()'")
    sass = subprocess.run(["cuobjdump", "-sass", sos[0]], capture_output=True, text=True, check=True).stdout
    counts, name = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name)
            counts[name] = collections.Counter()
            continue
        if name is None or "/*" not in line:
            continue
        for key, pat in PATTERNS:
            if re.search(pat, line):
                counts[name][key] += 1
    arch = re.findall(r"arch = (sm_\w+)", sass)
    print(f"# {os.path.relpath(sos[0], ROOT)}: {len(counts)} kernels, arch {sorted(set(arch))}")
    print(f"{'kernel':<58}" + "".join(f"{k:>9}" for k, _ in PATTERNS))
    for n, c in sorted(counts.items(), key=lambda kv: -(kv[1]["UTC*MMA"] * 1000 + kv[1]["UTMALDG"] + kv[1]["UBLKCP"])):
        print(f"{n[:57]:<58}" + "".join(f"{c[k]:>9}" for k, _ in PATTERNS))
    total = collections.Counter()
    for c in counts.values():
        total.update(c)
    print(f"{'TOTAL':<58}" + "".join(f"{total[k]:>9}" for k, _ in PATTERNS))
    assert total["HMMA"] == 0, "legacy mma.sync path present"


if __name__ == "__main__":
    main()
