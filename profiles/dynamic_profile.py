"""Dynamic profile of a kernel per SOURCE FUNCTION: joins the per-instruction counters of an ncu report's SASS page (executed warp
instructions, stall samples) with the -lineinfo attribution of the same build (nvdisasm), inlined copies included.
  python profiles/dynamic_profile.py <report.ncu-rep> <kernel regex> [libeg3d.so] > profiles/r02_k3b_dynamic_profile.txt
The report must have been captured from the library given (same SASS)."""
import bisect, collections, csv, glob, os, re, subprocess, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from code_footprint import source_functions, ROOT


def main():
    rep, kre = sys.argv[1], sys.argv[2]
    lib = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "edgegraph3d_b200", "libeg3d.so")
    funcs = {os.path.basename(p): source_functions(p) for p in glob.glob(os.path.join(ROOT, "edgegraph3d_b200", "csrc", "*.cu*"))}
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, capture_output=True)
        cubin = max(glob.glob(os.path.join(tmp, "*.cubin")), key=os.path.getsize)
        dis = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout
    kernel, cur = None, ("?", 0)
    attr = {}
    for line in dis.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", line)
        if m:
            kernel = m.group(1); continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+\S", line)
        if kernel and re.search(kre, kernel) and m:
            f, ln = cur
            name = "?"
            if f in funcs and funcs[f]:
                k = bisect.bisect_right([a for a, _ in funcs[f]], ln) - 1
                name = funcs[f][k][1] if k >= 0 else "?"
            attr[int(m.group(1), 16)] = f"{name} ({f})"
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    h = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
    H = rows[h]
    ia, ii, isamp, ini = H.index("Address"), H.index("Instructions Executed"), H.index("# Samples"), H.index("Warp Stall Sampling (Not-issued Samples)")
    noi = [i for i, x in enumerate(H) if x.strip() in ("stall_no_inst", "stall_no_instruction")]
    base = None
    inst, samp = collections.Counter(), collections.Counter()
    ilsb = H.index("stall_long_sb") if "stall_long_sb" in H else None
    inoi = H.index("stall_no_inst") if "stall_no_inst" in H else None
    lsb, nois = collections.Counter(), collections.Counter()
    for r in rows[h + 1:]:
        if len(r) <= max(ii, isamp) or not r[ia].startswith("0x"):
            if r and r[0] == "Kernel Name":
                break            # next launch of the same kernel: the first one is enough
            continue
        a = int(r[ia], 16)
        base = a if base is None else base
        f = attr.get(a - base, "?")
        inst[f] += int(r[ii] or 0); samp[f] += int(r[isamp] or 0)
        if ilsb is not None: lsb[f] += int(r[ilsb] or 0)
        if inoi is not None: nois[f] += int(r[inoi] or 0)
    ti, ts = sum(inst.values()), sum(samp.values())
    print(f"# {kre} in {os.path.basename(rep)}: {ti:.4g} executed warp instructions, {ts} stall samples; per source function (inlined copies included)")
    print(f"# {'function':44s} {'instructions':>14s} {'%':>6s} {'samples':>9s} {'% time':>7s} {'samples per 1e6 instr':>22s} {'no_inst %':>10s} {'long_sb %':>10s}")
    for f, n in samp.most_common(45):
        print(f"  {f:44s} {inst[f]:14d} {100.0 * inst[f] / ti:6.1f} {n:9d} {100.0 * n / ts:7.1f} {1e6 * n / max(1, inst[f]):22.1f} {100.0 * nois[f] / max(1, n):10.1f} {100.0 * lsb[f] / max(1, n):10.1f}")
    print(f"# all: no_instruction {100.0 * sum(nois.values()) / ts:.1f} % of the samples, long_scoreboard {100.0 * sum(lsb.values()) / ts:.1f} %")


if __name__ == "__main__":
    main()
