"""Registers, stack (spill) bytes and static shared memory of every kernel in libodo_b200.so (cuobjdump -res-usage),
demangled, as a markdown table.  CPU only:  python tools/resource_usage.py > profiles/resource_usage_r2.md"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "odometry_b200", "libodo_b200.so")


def main():
    out = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True, check=True).stdout
    rows = []
    src = "?"
    name = None
    for line in out.splitlines():
        m = re.match(r"identifier = .*/([a-z_]+\.cu)", line)
        if m:
            src = m.group(1)
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            name = m.group(1)
            continue
        m = re.match(r"\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", line)
        if m and name:
            rows.append((src, name, *map(int, m.groups())))
            name = None
    names = subprocess.run(["c++filt"], input="\n".join(r[1] for r in rows), capture_output=True, text=True).stdout.splitlines()
    print("# Resource usage of the kernels in libodo_b200.so (sm_100a, `cuobjdump -res-usage`)\n")
    print("STACK is the per-thread stack frame (spills and local arrays), SHARED the static shared memory; the persistent")
    print("kernels add dynamic shared memory at launch (k_lm_solve: its `Shared` struct; k_refine: slot state).\n")
    print("| source | kernel | registers | stack B | static smem B | local B |")
    print("|---|---|---|---|---|---|")
    for (src, _, reg, stack, sh, loc), dn in sorted(zip(rows, names), key=lambda t: (t[0][0], t[1])):
        dn = re.sub(r"\(anonymous namespace\)::", "", dn)
        dn = re.sub(r"^void ", "", dn)
        dn = re.sub(r"\(.*$", "", dn)
        print(f"| {src} | `{dn}` | {reg} | {stack} | {sh} | {loc} |")
    print(f"\n{len(rows)} kernels; {sum(1 for r in rows if r[3] > 0)} with a non-empty stack frame.")


if __name__ == "__main__":
    sys.exit(main())
