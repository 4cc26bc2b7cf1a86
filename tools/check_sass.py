#!/usr/bin/env python
"""Checks the 256-bit vertex stores of the built library at the SASS level.

ptxas 12.9 turns `st.global.v8.f32` into `STG.E.ENL2.256 [addr], Ra, Rb` (two register quads).  In out-of-line device
functions we saw it emit a plain 32-bit `STG.E [addr], Ra` for the same PTX when the eight values sit in one aligned
register octet -- only the first word of the vertex reached memory.  The wide store is therefore used only by
EmitSinkWide, for the fast joins (inlined into the tile kernels), and this script proves, for the .so that ships, that
  * every stroke tile kernel that emits contains its two STG.E.ENL2.256 instructions (one per fast-join vertex), and
  * no predicated 32-bit STG is paired with the 128-bit stores of the narrow path (the miscompile's signature when the
    choice between the two was still made at run time), and
  * the EMIT and FUSED instantiations of a tile geometry hold the same number of them (fast joins: 2, the dense join
    pass's EmitSinkPacked: the base vertices of a join).
Usage: tools/check_sass.py [lib.so]   (exit code 1 on a violation; prints a short report)
"""
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def check(lib: Path):
    sass = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True, check=True).stdout
    fn = None
    wide = {}
    bad = []
    window = []
    for ln in sass.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            fn = m.group(1)
            window = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if not m:
            continue
        ins = m.group(2).strip()
        if "STG.E.ENL2.256" in ins:
            wide[fn] = wide.get(fn, 0) + 1
        window.append(ins)
        window = window[-4:]
        # signature: "@P STG.E desc[..][Ra.64], Rb" directly followed by "@!P ST(G).E.128 desc[..][Ra.64], Rb"
        if len(window) >= 2:
            a, b = window[-2], window[-1]
            ma = re.match(r"@(!?P\d) STG\.E desc\[\w+\]\[(R\d+)\.64\], (R\d+)$", a.strip())
            mb = re.match(r"@(!?P\d) STG?\.E\.128 desc\[\w+\]\[(R\d+)\.64\], (R\d+)$", b.strip())
            if ma and mb and ma.group(2) == mb.group(2) and ma.group(3) == mb.group(3) and ma.group(1) != mb.group(1):
                bad.append((fn, a, b))
    return wide, bad


def main():
    lib = Path(sys.argv[1]) if len(sys.argv) > 1 else ROOT / "lyon_b200" / "liblyon_cuda.so"
    wide, bad = check(lib)
    emitters = {f: n for f, n in wide.items() if "k_stroke_tile" in f or "k_stroke_wtile" in f}
    for f, n in sorted(emitters.items()):
        print(f"{n:3d} x STG.E.ENL2.256  {f[40:120]}")
    for f, a, b in bad:
        print("MISCOMPILED wide store in", f, "\n   ", a, "\n   ", b)
    # the fast-join pass writes two vertices: two wide stores per emitting instantiation, none lost to the miscompile
    ok = bool(emitters) and not bad and all(n >= 2 for n in emitters.values())
    # The EMIT and FUSED instantiations of one tile geometry compile the same emit code: a wide store that one of them
    # lost (the miscompile seen in an out-of-line arc emitter: STG.E.ENL2.256 in one, a 32-bit STG.E in the other) shows up
    # as different counts.
    groups = {}
    for f, n in emitters.items():
        m = re.search(r"k_stroke_tileILi(\d+)ELi([12])ELi(\d+)E", f)
        if m:
            groups.setdefault((m.group(1), m.group(3)), {})[m.group(2)] = n
    for key, modes in sorted(groups.items()):
        if len(set(modes.values())) > 1:
            print(f"MISMATCH: k_stroke_tile<{key[0]}, *, {key[1]}> has {modes} 256-bit stores in EMIT / FUSED")
            ok = False
    print("check_sass:", "ok" if ok else "FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
