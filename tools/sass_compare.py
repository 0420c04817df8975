#!/usr/bin/env python3
"""Compare the machine code of kernels between two builds (object files or shared libraries), function by function.

    python tools/sass_compare.py OLD NEW [--match cb_spmm_kernel]

Used when a validated kernel is refactored without access to a GPU: if `cuobjdump -sass` of every instantiation is
byte-identical before and after, the change cannot alter its behaviour or speed.  Exit code 1 when a function that
exists on both sides differs (functions present on one side only are listed, not counted as differences)."""
import argparse
import re
import subprocess
import sys


def functions(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    fn, cur, res = None, [], {}
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if fn:
                res[fn] = "\n".join(cur)
            # anonymous-namespace kernels carry a hash of the source path in their name
            fn, cur = re.sub(r"_GLOBAL__N__[0-9a-f]{8}_", "_GLOBAL__N__", m.group(1)), []
        elif fn:
            cur.append(" ".join(line.split()))        # the column alignment depends on the longest line of the cubin
    if fn:
        res[fn] = "\n".join(cur)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("old")
    ap.add_argument("new")
    ap.add_argument("--match", default="", help="only functions whose mangled name contains this")
    a = ap.parse_args()
    old, new = functions(a.old), functions(a.new)
    names = sorted(n for n in set(old) | set(new) if a.match in n)
    same = diff = 0
    for n in names:
        if n in old and n in new:
            if old[n] == new[n]:
                same += 1
            else:
                diff += 1
                print("DIFFERENT:", n)
    only_old = [n for n in names if n not in new]
    only_new = [n for n in names if n not in old]
    print(f"{same} functions identical, {diff} different, {len(only_old)} only in OLD, {len(only_new)} only in NEW")
    sys.exit(1 if diff else 0)


if __name__ == "__main__":
    main()
