"""cuobjdump -sass of the hot kernels into profiles/sass_<kernel>.txt (one file per kernel family, every template
instance in it) plus an instruction-mix summary (profiles/sass_summary_<tag>.md).  Runs in the build container.

    python tools/dump_sass.py r2
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = {"k_disparity": "depth.o", "k_refine": "depth.o", "k_select_tile": "depth.o", "k_lm_solve": "lm.o", "k_lm_fill": "lm.o",
       "k_blur3": "pyramid.o", "k_pyrdown": "pyramid.o", "k_remap": "camera.o"}
# the template instances the KITTI step launches (the listings of the other instances differ only in constants)
KEEP = {"k_disparity": r"k_disparity<\(bool\)1, \(int\)1280, \(bool\)1>", "k_lm_solve": r"k_lm_solve<\(bool\)0>",
        "k_blur3": r"k_blur3<\(bool\)1, unsigned char>", "k_pyrdown": r"k_pyrdown<(unsigned char|float)>",
        "k_remap": r"k_remap<unsigned char>"}
WATCH = ["IMAD", "IDP", "FFMA", "FMUL", "FADD", "DFMA", "DADD", "DMUL", "LDG", "STG", "LDS", "STS", "SHFL", "REDUX", "VOTE",
         "BAR", "MUFU", "ATOM", "RED", "UTMALDG", "UTMASTG", "UBLKCP", "LDGSTS", "LDTM", "UTCMMA", "UTCHMMA", "SYNCS"]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
    out_dir = os.path.join(ROOT, "profiles")
    rows = []
    cache = {}
    for kern, obj in OBJ.items():
        path = os.path.join(ROOT, "odometry_b200", "build", obj)
        if path not in cache:
            cache[path] = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
        funcs = re.split(r"(?=\n\s*Function : )", cache[path])
        mine = [f for f in funcs if re.search(r"Function : \S*" + kern, f)]
        if kern in KEEP:
            def dem(f):
                return subprocess.run(["cu++filt", re.search(r"Function : (\S+)", f).group(1)], capture_output=True, text=True).stdout
            mine = [f for f in mine if re.search(KEEP[kern], dem(f))]
        if not mine:
            continue
        with open(os.path.join(out_dir, f"sass_{kern}.txt"), "w") as fh:
            fh.write(f"# cuobjdump -sass odometry_b200/build/{obj}, functions matching {kern} (sm_100a)\n")
            fh.write("".join(mine))
        for f in mine:
            name = re.search(r"Function : (\S+)", f).group(1)
            dem = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
            ops = collections.Counter()
            n = 0
            for m in re.finditer(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", f, re.M):
                n += 1
                op = m.group(1).split(".")[0]
                for w in WATCH:
                    if op.startswith(w):
                        ops[w] += 1
                        break
            rows.append((dem[:110], n, ops))
    with open(os.path.join(out_dir, f"sass_summary_{tag}.md"), "w") as fh:
        fh.write(f"# SASS instruction mix of the hot kernels ({tag}, sm_100a, static counts)\n\n")
        fh.write("Full listings: `profiles/sass_<kernel>.txt` (`python tools/dump_sass.py`).  No tensor-core (UTC*MMA / LDTM) "
                 "instructions are expected: nothing on the path is a dense contraction (DESIGN.md section 3).\n\n")
        fh.write("| kernel | SASS instrs | " + " | ".join(WATCH) + " |\n|---|---|" + "---|" * len(WATCH) + "\n")
        for name, n, ops in rows:
            fh.write(f"| `{name}` | {n} | " + " | ".join(str(ops.get(w, 0)) for w in WATCH) + " |\n")
    print(f"wrote {len(rows)} kernels")


if __name__ == "__main__":
    main()
