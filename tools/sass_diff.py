#!/usr/bin/env python
"""Compare the SASS of every kernel in two builds of a cubin-bearing .so (or two extracted cubins), ignoring
label numbering: proves that a source change (a guard, a helper, an extra kernel) left the measured kernels'
machine code alone.   python tools/sass_diff.py old.so new.so"""
import hashlib
import os
import re
import subprocess
import sys
import tempfile


def kernels(so):
    td = tempfile.mkdtemp(prefix="sassdiff_")
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=td, check=True, stdout=subprocess.DEVNULL)
    out = {}
    for cub in sorted(os.listdir(td)):
        txt = subprocess.run(["nvdisasm", os.path.join(td, cub)], capture_output=True, text=True).stdout
        cur = None
        for line in txt.splitlines():
            if line.startswith(".text."):
                cur = re.sub(r"_GLOBAL__N__[0-9a-f]{8}_", "_GLOBAL__N__", line[6:].rstrip(":"))
                out[cur] = []
            elif cur and re.match(r"^\s*/\*[0-9a-f]{4,}\*/", line):
                t = re.sub(r"/\*[0-9a-f]+\*/", "", line)
                t = re.sub(r"\.L_x_\d+", ".L", t)
                t = re.sub(r"\$__internal_\d+_", "$__internal_", t)
                out[cur].append(t.strip())
    return {k: (hashlib.md5("\n".join(v).encode()).hexdigest(), len(v)) for k, v in out.items()}


def main():
    a, b = kernels(sys.argv[1]), kernels(sys.argv[2])
    same = [k for k in a if k in b and a[k] == b[k]]
    changed = [k for k in a if k in b and a[k] != b[k]]
    print(f"{len(same)} kernels identical, {len(changed)} changed, {len([k for k in b if k not in a])} new, "
          f"{len([k for k in a if k not in b])} removed")
    for k in changed:
        print("  changed:", k[:150], a[k][1], "->", b[k][1], "instructions")
    sys.exit(1 if changed else 0)


if __name__ == "__main__":
    main()
