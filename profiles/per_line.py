"""Per-source-line instruction counts and stall samples of one kernel from an ncu report (source page) and the line
   table of the cubin that was profiled:
     python profiles/per_line.py <report.ncu-rep> <object or library holding ONE cubin with the kernel, e.g.
            bunmpc_b200/csrc/obj/inst_96_0.o> <mangled kernel name> [N_iter]
   Prints, per source line: warp-instructions (divided by N_iter if given), share of the stall samples, the dominant
   stall reasons and the opcode mix."""
import collections
import csv
import glob
import io
import os
import re
import subprocess
import sys
import tempfile


def line_table(so, kernel):
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=d, capture_output=True)
    sass = "".join(subprocess.run(["nvdisasm", "-g", "-c", f], capture_output=True, text=True).stdout
                   for f in sorted(glob.glob(d + "/*.cubin")))
    tab, cur, inside = {}, None, False
    for ln in sass.splitlines():
        if ln.startswith(".text."):
            inside = ln.strip() == ".text." + kernel + ":"
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            tab[int(m.group(1), 16)] = (cur, m.group(2))
    return tab


def main():
    rep, so, kernel = sys.argv[1:4]
    n_iter = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
    tab = line_table(so, kernel)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[1]
    iA, iS, iE, iN = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    stall_cols = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    base = int(rows[2][iA], 16)
    per = collections.defaultdict(collections.Counter)
    smp = collections.Counter()
    why = collections.defaultdict(collections.Counter)
    unmapped = 0
    for r in rows[2:]:
        off = int(r[iA], 16) - base
        toks = r[iS].strip().split()
        op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0] if toks else "?"
        if off not in tab:
            unmapped += 1
        line = tab.get(off, (("?", 0), ""))[0]
        per[line][op] += int(r[iE] or 0)
        smp[line] += int(r[iN] or 0)
        for i, name in stall_cols:
            why[line][name] += int(r[i] or 0)
    tot = sum(sum(c.values()) for c in per.values())
    tot_s = max(1, sum(smp.values()))
    allwhy = collections.Counter()
    for c in why.values():
        allwhy.update(c)
    print(f"total warp-instructions {tot / n_iter:.1f}   stall samples {tot_s}   (unmapped sass rows: {unmapped})")
    print("stall reasons overall: " + " ".join(f"{k}:{100.0 * v / max(1, sum(allwhy.values())):.1f}%" for k, v in allwhy.most_common(8)))
    for line, c in sorted(per.items(), key=lambda kv: -smp[kv[0]])[:60]:
        s = sum(c.values())
        mix = " ".join(f"{k}:{v / n_iter:.1f}" for k, v in c.most_common(5))
        w = " ".join(f"{k}:{100.0 * v / max(1, sum(why[line].values())):.0f}%" for k, v in why[line].most_common(3))
        print(f"{str(line):26s} inst {s / n_iter:9.1f}  samples {100.0 * smp[line] / tot_s:5.1f}%  [{w}]  {mix}")


if __name__ == "__main__":
    main()
