#!/usr/bin/env python
"""Regenerates profiles/sass/ from the built library (no GPU needed): python tools/dump_sass.py

Splits `cuobjdump -sass libtcamcrf.so` into one listing per kernel instantiation the benchmarked paths run;
all_functions.txt lists every kernel in the library."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "tcam_wsol_video_b200", "csrc", "libtcamcrf.so")
OUT = os.path.join(ROOT, "profiles", "sass")
WANT = """build_kernelILi5Ef build_kernelILi5Eh build_kernelILi3Ef build_dedup_kernelILi5Ef build_dedup_kernelILi5Eh
build_dedup_kernelILi3Ef seed_fused_kernel neighbour_kernelILi5E neighbour_kernelILi3E
splat_kernelILi5ELi4ELb0E splat_kernelILi5ELi2ELb0E splat_kernelILi5ELi2ELb1E splat_rows_kernelILi5ELi3ELb0E
blur_kernelILi4ELi3E blur_kernelILi4ELi1E blur_kernelILi2ELi1E
slice_kernelILi5ELi4ELb0E slice_kernelILi5ELi2ELb0E loss_backward_kernel loss_backward_logits_kernel prepare_kernel
vertex_init_kernel temporal_max_kernel temporal_max_renorm_kernel prepare_std_cams_kernel seed_select_kernel
seed_labels_kernel otsu_roi_kernel""".split()


def strip_encodings(listing: str) -> str:
    """Mnemonics only: drops the hex encodings (the trailing comment and the continuation line of every instruction)."""
    out = []
    for line in listing.splitlines():
        if re.match(r"^\s*/\* 0x[0-9a-f]{16} \*/\s*$", line) or line.startswith("\t.headerflags"):
            continue
        out.append(re.sub(r"\s*/\* 0x[0-9a-f]{16} \*/\s*$", "", line))
    return "\n".join(out) + "\n"


def main():
    text = subprocess.run(["cuobjdump", "-sass", SO], check=True, capture_output=True, text=True).stdout
    parts = re.split(r"(?m)^(?=\s*Function : )", text)
    funcs = {}
    for part in parts:
        if "Function : " not in part:
            continue
        name = part.split("Function : ", 1)[1].split()[0]
        body = part.split("\n\t.......", 1)[0]   # drop cuobjdump's trailer after the function
        funcs[name] = part if body == part else body + "\n\t..........\n"
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "all_functions.txt"), "w") as f:
        f.write("\n".join(sorted(funcs)) + "\n")
    for w in WANT:
        hits = sorted(n for n in funcs if re.search(r"\d+" + re.escape(w) + r"E", n))
        if not hits:
            raise SystemExit(f"kernel not in the library: {w}")
        with open(os.path.join(OUT, w + ".sass"), "w") as f:
            f.write(strip_encodings(funcs[hits[0]]))
    print(f"{len(funcs)} kernels in the library, {len(WANT)} listings written to {OUT}")


if __name__ == "__main__":
    main()
