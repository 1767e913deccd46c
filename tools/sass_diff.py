#!/usr/bin/env python
"""Per-kernel SASS comparison of two builds of libopb.so (cuobjdump -sass): lists the kernels whose
instruction streams differ.  Used to show that a refactor left a shipped kernel's machine code
untouched when no GPU is at hand to re-measure it.

    python tools/sass_diff.py old.so new.so [name-substring ...]
"""
import re
import subprocess
import sys


def kernels(path):
    txt = subprocess.run(["cuobjdump", "-sass", path], stdout=subprocess.PIPE, text=True, check=True).stdout
    out, name, body = {}, None, []
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if name:
                out[name] = body
            name, body = m.group(1), []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if m and name:
            body.append(re.sub(r"\s+", " ", m.group(1)))
    if name:
        out[name] = body
    return out


def canonical(body):
    """Register names replaced by their order of first appearance (per register class): two streams that differ
    only in ptxas' register assignment -- which is not stable across unrelated edits of the translation unit --
    compare equal."""
    maps = {}
    def sub(m):
        cls = m.group(1)
        d = maps.setdefault(cls, {})
        return "%s#%d" % (cls, d.setdefault(m.group(0), len(d)))
    return [re.sub(r"\b(UR|UP|R|P|B)(\d+)\b", sub, ins) for ins in body]


def main():
    a, b = kernels(sys.argv[1]), kernels(sys.argv[2])
    subs = sys.argv[3:]
    rc = 0
    for k in sorted(set(a) | set(b)):
        if subs and not any(s in k for s in subs):
            continue
        if k not in a:
            print("NEW      ", k, len(b[k]), "instructions")
        elif k not in b:
            print("REMOVED  ", k)
        elif a[k] == b[k]:
            print("IDENTICAL", k, len(a[k]), "instructions")
        elif canonical(a[k]) == canonical(b[k]):
            print("RENAMED  ", k, len(a[k]), "instructions (same stream, different register assignment)")
        else:
            rc = 1
            print("DIFFERENT", k, len(a[k]), "->", len(b[k]), "instructions")
    return rc


if __name__ == "__main__":
    sys.exit(main())
