#!/usr/bin/env python
"""Pins the machine code of the kernels whose GPU results are quoted in DESIGN.md.

The GPU is a scarce resource here: most edits happen without one.  This tool hashes the SASS
(`cuobjdump -sass`, per kernel) of the in-tree build and compares it with the hashes recorded when
those kernels were last run on a B200 (`profiles/sass_pins.json`), so that a refactor which was
meant to leave a measured kernel alone can be shown to have done so -- and one that did not is
noticed before its old numbers are quoted for new code.

    python tools/sass_pins.py            # check (exit 1 on a mismatch)
    python tools/sass_pins.py --record   # after re-validating on hardware: pin the current build
"""
import hashlib
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "feriphys_b200", "csrc", "_build")
PINS = os.path.join(ROOT, "profiles", "sass_pins.json")
# kernels that have NOT run on hardware in their current form are left out of the pins
UNPINNED = re.compile(r"nl_build_kernelILb1|nl_walk_kernelILi48|nl_walk_kernelILi64ELi5ELb1")


def kernel_hashes():
    out = {}
    for obj in sorted(os.listdir(BUILD)):
        if not obj.endswith(".o"):
            continue
        txt = subprocess.run(["cuobjdump", "-sass", os.path.join(BUILD, obj)], capture_output=True, text=True,
                             check=True).stdout
        for part in re.split(r"\n\s*Function : ", txt)[1:]:
            name, body = part.split("\n", 1)
            # anonymous-namespace symbols carry a per-file hash of the source path: drop it
            name = re.sub(r"_GLOBAL__N__[0-9a-f]+_\d+_(\w+?)_cu_[0-9a-f]+", r"anon_\1", name.strip())
            # (cuobjdump pads its columns to the longest line of the whole object: collapse the blanks)
            lines = [re.sub(r"\s+", " ", ln).strip() for ln in body.splitlines() if ln.strip()]
            out[f"{obj}:{name}"] = hashlib.sha256("\n".join(lines).encode()).hexdigest()
    return out


def nvcc_version():
    out = subprocess.run(["nvcc", "--version"], capture_output=True, text=True).stdout
    m = re.search(r"release [\d.]+, V([\d.]+)", out)
    return m.group(1) if m else "unknown"


def main():
    if "--record" in sys.argv:
        cur = kernel_hashes()
        pins = {k: v for k, v in cur.items() if not UNPINNED.search(k)}
        with open(PINS, "w") as fh:
            json.dump({"note": "SASS hashes of the kernels as last run on a B200 (tools/sass_pins.py)",
                       "nvcc": nvcc_version(), "kernels": pins}, fh, indent=1, sort_keys=True)
        print(f"recorded {len(pins)} kernels ({len(cur) - len(pins)} not yet run on hardware left out)")
        return 0
    with open(PINS) as fh:
        doc = json.load(fh)
    pins = doc["kernels"]
    if doc.get("nvcc") not in (None, nvcc_version()):
        print(f"pins were recorded with nvcc {doc['nvcc']}, this is {nvcc_version()}: another compiler, other code")
        return 2
    cur = kernel_hashes()
    bad = [k for k, v in pins.items() if cur.get(k) != v]
    for k in bad:
        print(("CHANGED " if k in cur else "MISSING ") + k)
    print(f"{len(pins) - len(bad)} of {len(pins)} pinned kernels unchanged; "
          f"{len([k for k in cur if k not in pins])} kernels not pinned")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
