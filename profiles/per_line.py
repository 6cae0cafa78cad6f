"""Per-source-line instruction counts of one kernel from an ncu report (source page) and the line table of the
   cubin that was profiled:  python profiles/per_line.py <report.ncu-rep> <object or library holding ONE cubin with the kernel, e.g. bunmpc_b200/csrc/obj/inst_0_0.o> <mangled kernel name> [N_iter]
   Prints warp-instructions per source line (divided by N_iter if given), with the opcode mix of each line."""
import collections
import csv
import io
import re
import subprocess
import sys
import tempfile


def line_table(so, kernel):
    d = tempfile.mkdtemp()
    import os
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=d, capture_output=True)
    import glob
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
    iA, iS, iE = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed")
    base = int(rows[2][iA], 16)
    per = collections.defaultdict(collections.Counter)
    bad = 0
    for r in rows[2:]:
        off = int(r[iA], 16) - base
        src = r[iS].strip()
        toks = src.split()
        op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0] if toks else "?"
        if off not in tab or tab[off][1].split()[0:1] != src.split()[0:1] and not src.startswith("@"):
            bad += 0 if off in tab else 1
        line = tab.get(off, (("?", 0), ""))[0]
        per[line][op] += int(r[iE] or 0)
    tot = sum(sum(c.values()) for c in per.values())
    print(f"total {tot / n_iter:.1f}   (unmapped sass rows: {bad})")
    for line, c in sorted(per.items(), key=lambda kv: -sum(kv[1].values()))[:70]:
        s = sum(c.values())
        mix = " ".join(f"{k}:{v / n_iter:.1f}" for k, v in c.most_common(6))
        print(f"{str(line):28s} {s / n_iter:9.1f}  {mix}")


if __name__ == "__main__":
    main()
