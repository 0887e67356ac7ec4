#!/usr/bin/env python
"""Per-source-line summary of an ncu report: instructions executed and stall samples.

usage: ncu_lines.py report.ncu-rep kernel_regex [top_n [samples]]   (samples: sort by stall samples)
Runs `ncu -i report --page source --csv --print-source cuda,sass -k regex:<kernel_regex>` and
aggregates the rows that carry a CUDA line number.
"""
import csv, io, subprocess, sys

def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                          "-k", "regex:" + kern], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    fname, hdr, lines = None, None, []
    for r in rows:
        if not r: continue
        if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
        if r[0] == "Line No": hdr = r; continue
        if r[0] == "Function Name" or hdr is None: continue
        if r[0].strip().isdigit():
            d = dict(zip(hdr, r))
            try:
                inst = int(d["Instructions Executed"]); smp = int(d["# Samples"])
            except (KeyError, ValueError):
                continue
            lines.append((inst, smp, fname, int(r[0]), r[1].strip()[:90]))
    tot_i = sum(l[0] for l in lines) or 1
    tot_s = sum(l[1] for l in lines) or 1
    print(f"total warp-inst {tot_i}  samples {tot_s}")
    by_samples = len(sys.argv) > 4 and sys.argv[4] == "samples"
    for inst, smp, f, ln, src in sorted(lines, key=lambda l: (l[1], l[0]) if by_samples else l, reverse=True)[:top]:
        print(f"{100*inst/tot_i:5.1f}%i {100*smp/tot_s:5.1f}%s  {f}:{ln}  {src}")

if __name__ == "__main__":
    main()
