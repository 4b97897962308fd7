"""Mnemonic counts of the built library (cuobjdump -sass), whole library and per tcgen05 / TMA kernel.

    python benchmarks/sass_summary.py > profiles/r2c_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "torch-geometric-pool_b200", "tgp_b200", "libtgp_b200.so")
MNEMONICS = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "UTCCP", "SYNCS", "LDGSTS", "HMMA",
             "ELECT", "MATCH", "REDUX", "ATOM", "RED", "DADD", "SHFL", "LDL", "STL"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    demangle = {}
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            op = m.group(1)
            for k in MNEMONICS:
                if op == k or op.startswith(k + "."):  # whole mnemonic (HMMA does not count UTCHMMA)
                    per[cur][k] += 1
    names = list(per)
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    demangle = dict(zip(names, out))
    total = collections.Counter()
    for c in per.values():
        total.update(c)
    print(f"# SASS extract of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass, sm_100a)")
    print("# mnemonic counts over the whole library:")
    print("  " + "  ".join(f"{k}={total[k]}" for k in MNEMONICS))
    print("\n# kernels that use the tensor core / TMA / TMEM:")
    for n, c in per.items():
        if c["UTCHMMA"] or c["UTMALDG"] or c["UTMASTG"] or c["LDTM"] or c["STTM"]:
            print(demangle[n][:150])
            print("    " + "  ".join(f"{k}={c[k]}" for k in MNEMONICS if c[k]))


if __name__ == "__main__":
    main()
