#!/usr/bin/env python
"""Per-kernel fingerprints of the compiled SASS, to prove that an edit left a kernel's machine code untouched.

    python tools/sass_snapshot.py save  [file]     # default tools/_bin/sass_snapshot.json (git-ignored)
    python tools/sass_snapshot.py diff  [file]     # rebuild first; lists kernels whose SASS changed / appeared / vanished

Why: there is no GPU in the build container, so a kernel that was validated on a B200 must not change by accident when
host code, headers or OTHER template instantiations are edited.  The fingerprint is the sha1 of `cuobjdump -sass` of one
function with the address / encoding comments stripped.  ptxas does not always number (uniform) registers the same way in
two runs over the same source, so a second fingerprint with register numbers masked is kept: a kernel whose exact hash
differs but whose masked hash is equal is reported as RENAMED (same instruction stream), not CHANGED."""
import glob
import hashlib
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "codename-rvc-fork-3_b200", "_obj")
DEFAULT = os.path.join(ROOT, "tools", "_bin", "sass_snapshot.json")


def fingerprints():
    out = {}
    for obj in sorted(glob.glob(os.path.join(OBJ, "*.o"))):
        txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
        name, lines = None, []

        def flush():
            if name is not None:
                body = "\n".join(lines)
                masked = re.sub(r"\b(U?R|U?P)\d+\b", r"\1#", body)
                out[f"{os.path.basename(obj)}:{name}"] = {"sha1": hashlib.sha1(body.encode()).hexdigest(),
                                                         "sha1_masked": hashlib.sha1(masked.encode()).hexdigest(), "lines": len(lines)}

        for ln in txt.splitlines():
            m = re.search(r"Function : (\S+)", ln)
            if m:
                flush()
                name, lines = re.sub(r"_GLOBAL__N__[0-9a-f]+_", "_GLOBAL__N__", m.group(1)), []   # path-dependent hash
            elif name is not None:
                ln = re.sub(r"/\*.*?\*/", "", ln).strip()          # addresses and encodings
                if ln and not ln.startswith("identifier ="):
                    lines.append(ln)
        flush()
    return out


def main():
    cmd = sys.argv[1] if len(sys.argv) > 1 else "diff"
    path = sys.argv[2] if len(sys.argv) > 2 else DEFAULT
    cur = fingerprints()
    if cmd == "save":
        os.makedirs(os.path.dirname(path), exist_ok=True)
        json.dump(cur, open(path, "w"), indent=1, sort_keys=True)
        print(f"{len(cur)} kernels -> {path}")
        return 0
    old = {k: v for k, v in json.load(open(path)).items() if not k.startswith("_")}
    differs = [k for k in cur if k in old and cur[k]["sha1"] != old[k]["sha1"]]
    renamed = [k for k in differs if old[k].get("sha1_masked") == cur[k]["sha1_masked"]]
    changed = [k for k in differs if k not in renamed]
    new = [k for k in cur if k not in old]
    gone = [k for k in old if k not in cur]
    for tag, ks in (("CHANGED", changed), ("RENAMED", renamed), ("NEW", new), ("GONE", gone)):
        for k in ks:
            print(tag, k)
    print(f"{len(cur)} kernels: {len(changed)} changed, {len(renamed)} with renamed registers only, {len(new)} new, {len(gone)} gone")
    return 1 if changed or gone else 0


if __name__ == "__main__":
    sys.exit(main())
