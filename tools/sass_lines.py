#!/usr/bin/env python3
"""Join an ncu SASS source page (dynamic instruction counts per SASS address) with nvdisasm's line
info of the matching cubin: dynamic instructions per CUDA source line / function.
usage: tools/sass_lines.py prof.ncu-rep file.cubin kernel_substring"""
import collections
import csv
import io
import re
import subprocess
import sys


def main(rep, cubin, kern):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = rows[2:]
    base = int(data[0][ix["Address"]], 16)
    dyn = {}
    for r in data:
        off = int(r[ix["Address"]], 16) - base
        dyn[off] = (int(r[ix["Instructions Executed"]]), r[ix["Source"]].strip(), int(r[ix["# Samples"]]))
    dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    # find the function
    cur_fn, cur_line, in_fn = None, ("?", 0), False
    per_line = collections.Counter()
    per_line_samples = collections.Counter()
    per_inl = collections.Counter()
    line_re = re.compile(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?')
    ins_re = re.compile(r'/\*([0-9a-f]{4,6})\*/\s+(.*?);')
    total = 0
    for ln in dis.split("\n"):
        if ln.startswith("//--------------------- .text."):
            in_fn = kern in ln
            continue
        if not in_fn:
            continue
        m = line_re.search(ln)
        if m:
            cur_line = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = ins_re.search(ln)
        if m:
            off = int(m.group(1), 16)
            if off in dyn:
                n, src, smp = dyn[off]
                per_line[cur_line] += n
                per_line_samples[cur_line] += smp
                total += n
    tot_s = sum(per_line_samples.values())
    print("total dynamic warp instructions", total)
    for (f, l), n in per_line.most_common(60):
        print("%6.2f%% instr %6.2f%% samples  %s:%d" % (100.0 * n / total, 100.0 * per_line_samples[(f, l)] / tot_s, f, l))


if __name__ == "__main__":
    main(*sys.argv[1:4])
