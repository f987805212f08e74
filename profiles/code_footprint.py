"""Static code-footprint profile of the K3 kernels (no GPU needed): every SASS instruction of a kernel's .text section is
attributed, through the -lineinfo records, to the source function whose body contains its line, so the table says where the
kernel's instruction bytes come from (inlined copies included).  K3 is bound by instruction-cache refills
(profiles/r01_k3b_icache.md), so bytes are the currency.

  python profiles/code_footprint.py [edgegraph3d_b200/libeg3d.so] > profiles/r01_k3_code_footprint.txt
"""
import bisect
import collections
import glob
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def source_functions(path):
    """[(first line, name)] of every function definition in a .cuh file (good enough: a line that starts a definition)."""
    out = []
    pat = re.compile(r"^(?:template\s*<[^>]*>\s*)?(?:static\s+|inline\s+|EG3D_HD_NI\s+|EG3D_HD\s+|__device__\s+|__global__\s+|__noinline__\s+|__forceinline__\s+|void\s+__launch_bounds__\([^)]*\)\s+)*"
                     r"[A-Za-z_][\w:<>\*&\s]*?\b([A-Za-z_]\w*)\s*\([^;]*$")
    for i, line in enumerate(open(path), 1):
        if line[:1] in " \t/#}" or "(" not in line or line.rstrip().endswith(";"):
            continue
        k = re.match(r"^__global__\s+void\s+(?:__launch_bounds__\([^)]*\)\s+)?([A-Za-z_]\w*)\s*\(", line)
        if k:
            out.append((i, k.group(1) + " [kernel body]"))
            continue
        m = pat.match(line)
        if m and m.group(1) not in ("if", "for", "while", "switch", "return", "sizeof"):
            out.append((i, m.group(1)))
    return out


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "edgegraph3d_b200", "libeg3d.so")
    funcs = {os.path.basename(p): source_functions(p) for p in glob.glob(os.path.join(ROOT, "edgegraph3d_b200", "csrc", "*.cuh")) + glob.glob(os.path.join(ROOT, "edgegraph3d_b200", "csrc", "*.cu"))}
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, capture_output=True)
        cubin = max(glob.glob(os.path.join(tmp, "*.cubin")), key=os.path.getsize)
        dis = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout
    kernel, cur = None, ("?", 0)
    per = collections.defaultdict(lambda: collections.Counter())
    for line in dis.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", line)
        if m:
            kernel = m.group(1)
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if kernel and re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+\S", line):
            f, ln = cur
            name = "?"
            if f in funcs and funcs[f]:
                k = bisect.bisect_right([a for a, _ in funcs[f]], ln) - 1
                name = funcs[f][k][1] if k >= 0 else "?"
            per[kernel][f"{name} ({f})"] += 16
    for kernel in sorted(per, key=lambda k: -sum(per[k].values())):
        if "k3" not in kernel:
            continue
        total = sum(per[kernel].values())
        print(f"\n{kernel}: {total} bytes of SASS")
        for name, b in per[kernel].most_common(40):
            print(f"  {b:8d} B  {100.0 * b / total:5.1f} %  {name}")


if __name__ == "__main__":
    main()
