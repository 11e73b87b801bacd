"""SASS instruction histogram per kernel of liloc_b200/libliloc_gpu.so (-> profiles/sass_rNN.txt).

  python profiles/sass_hist.py [path/to/lib.so] > profiles/sass_r02.txt
"""
import collections
import re
import subprocess
import sys

OPS = ["FFMA", "FMUL", "FADD", "DFMA", "DMUL", "DADD", "F2F", "MUFU", "LDG", "STG", "LDS", "STS", "ATOMG", "ATOMS", "RED", "SHFL", "BAR",
       "MATCH", "UTMALDG", "UBLKCP", "HMMA", "LDL", "STL"]


def main():
    path = sys.argv[1] if len(sys.argv) > 1 else "liloc_b200/libliloc_gpu.so"
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    fn, hist, order = None, collections.defaultdict(collections.Counter), []
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1); order.append(fn); continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and fn:
            op = m.group(1)
            hist[fn]["instructions"] += 1
            for o in OPS:
                if op == o or op.startswith(o + ".") or (o in ("ATOMG", "ATOMS") and op.startswith(o)):
                    hist[fn][o] += 1
    for fn in order:
        h = hist[fn]
        print(fn)
        print("    instructions %d  " % h["instructions"] + "  ".join(f"{o}={h[o]}" for o in OPS))


if __name__ == "__main__":
    main()
