"""Regenerates the path fixtures under tests/golden/ from the reference checkout (run in the build
container only; /root/reference does not exist on the GPU box).

  logo.path       <- /root/reference/assets/logo.path (the CLI README's example input, data asset)
  rust_logo.path  <- crates/extra/src/rust_logo.rs: the sequence of SvgPathBuilder calls of
                     `build_logo_path` rewritten as SVG path data (M/l/c/h/v/s/z commands with the same
                     decimal literals), so that parse_svg_path() rebuilds the same f32 path that
                     `Path::builder().with_svg()` + build_logo_path() builds in bench/tess.
"""
import re
import shutil
import sys
from pathlib import Path

REF = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
OUT = Path(__file__).resolve().parent

shutil.copyfile(REF / "assets" / "logo.path", OUT / "logo.path")

src = (REF / "crates/extra/src/rust_logo.rs").read_text()
CMD = {
    "move_to": "M", "line_to": "L", "cubic_bezier_to": "C", "horizontal_line_to": "H",
    "vertical_line_to": "V", "smooth_cubic_bezier_to": "S", "relative_cubic_bezier_to": "c",
    "relative_line_to": "l", "relative_horizontal_line_to": "h", "relative_vertical_line_to": "v",
    "smooth_relative_cubic_bezier_to": "s", "close": "z",
}
out = []
for m in re.finditer(r"path\.(\w+)\((.*?)\);", src, re.S):
    name, args = m.group(1), m.group(2)
    nums = re.findall(r"-?\d+\.?\d*(?:e-?\d+)?", args)
    out.append(CMD[name] + " " + " ".join(nums))
(OUT / "rust_logo.path").write_text("\n".join(out) + "\n")
print(f"wrote {len(out)} commands")
