#!/usr/bin/env python3
"""Opcode-class histogram of one kernel.

  tools/sass_hist.py static  libort.so  <kernel substring>
        static instruction counts from `cuobjdump -sass`
  tools/sass_hist.py dynamic prof.ncu-rep [<kernel substring>]
        warp-level executed instruction counts from the SASS source page of an ncu report
        (`ncu --set full --import-source on`): every SASS instruction weighted by how often it ran
  tools/sass_hist.py lines   prof.ncu-rep file.cubin <kernel substring>
        the same per CUDA source line (needs the cubin of the profiled build: cuobjdump -xelf all)

Classes follow the pipes of the SM: fp64 (DFMA/DMUL/DADD/DSETP + the 64-bit MUFU seeds), imad
(integer multiply-add pipe, incl. IMAD used as a move), alu (LOP3/IADD3/SHF/ISETP/SEL/MOV/...),
pred (PLOP3, P2R, R2P), fsel (FSEL/FMNMX on register halves), fp32 (FFMA/FMUL/FADD/FSETP), mufu,
conv (I2F/F2I/F2F/FRND), warp (VOTE/SHFL/MATCH/REDUX/POPC), lsu (LDS/STS/LDG/STG/LDL/STL/RED/ATOM),
const (LDC/LDCU/ULDC), uniform (U* datapath), ctrl (BRA/BSSY/BSYNC/WARPSYNC/EXIT/NOP/...).
"""
import collections
import csv
import io
import re
import subprocess
import sys

CLASSES = [
    ("fp64", r"^(DFMA|DMUL|DADD|DSETP|DMNMX)"),
    ("mufu", r"^MUFU"),
    ("imad", r"^(IMAD|IMUL|IDP)"),
    ("fp32", r"^(FFMA|FMUL|FADD|FSETP|FCHK|FSWZ)"),
    ("fsel", r"^(FSEL|FMNMX)"),
    ("pred", r"^(PLOP3|P2R|R2P|PSETP)"),
    ("conv", r"^(I2F|F2I|F2F|FRND|I2I|I2FP|F2FP)"),
    ("warp", r"^(VOTE|VOTEU|SHFL|MATCH|REDUX|POPC|FLO|BREV|CREDUX)"),
    ("const", r"^(LDC|LDCU|ULDC)"),
    ("lsu", r"^(LDS|STS|LDG|STG|LDL|STL|RED|REDG|ATOM|ATOMS|ATOMG|LD|ST|LDSM|MEMBAR|CCTL|ERRBAR)"),
    ("uniform", r"^(U[A-Z0-9]+|R2UR|S2UR)"),
    ("ctrl", r"^(BRA|BRX|JMP|BSSY|BSYNC|WARPSYNC|EXIT|NOP|RET|CALL|BAR|YIELD|ENDCOLLECTIVE|BREAK|BMOV|DEPBAR|NANOSLEEP|KILL|BPT|ACQBULK|ELECT)"),
    ("alu", r"^(LOP3|LOP|IADD3|IADD|VIADD|SHF|SHL|SHR|LEA|ISETP|SEL|MOV|CS2R|S2R|VIMNMX|IMNMX|IABS|PRMT|BMSK|SGXT|VABSDIFF|ISCADD|HADD2|HFMA2|HMUL2)"),
]


def classify(op):
    for name, pat in CLASSES:
        if re.match(pat, op):
            return name
    return "other:" + op


def opcode(text):
    t = text.strip()
    t = re.sub(r"^@!?U?P[0-9T]+\s+", "", t)
    return t.split()[0].rstrip(";") if t else ""


def report(counts, total, title):
    by_class = collections.Counter()
    by_op = collections.Counter()
    for op, n in counts.items():
        by_class[classify(op)] += n
        by_op[op.split(".")[0]] += n
    print(title)
    print("total %d" % total)
    for c, n in by_class.most_common():
        print("  %-10s %14d  %6.2f %%" % (c, n, 100.0 * n / total))
    print("top opcodes:")
    for op, n in by_op.most_common(25):
        print("  %-14s %14d  %6.2f %%" % (op, n, 100.0 * n / total))


def static(lib, kern):
    names = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    counts, total, on, found = collections.Counter(), 0, False, None
    for ln in names.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            on = kern in m.group(1)
            found = m.group(1) if on else found
            continue
        if not on:
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(.*?);", ln)
        if m:
            op = opcode(m.group(1))
            if op:
                counts[op] += 1
                total += 1
    report(counts, total, "static SASS instructions of %s" % found)


def dynamic(rep, kern=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] +
                         (["--kernel-name", "regex:" + kern] if kern else []),
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr_i = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
    ix = {h: i for i, h in enumerate(rows[hdr_i])}
    counts, total, thr = collections.Counter(), 0, 0
    for r in rows[hdr_i + 1:]:
        if len(r) <= ix["Instructions Executed"]:
            continue
        try:
            n = int(r[ix["Instructions Executed"]])
        except ValueError:
            continue
        op = opcode(r[ix["Source"]])
        if not op:
            continue
        counts[op] += n
        total += n
        if "Thread Instructions Executed" in ix:
            thr += int(r[ix["Thread Instructions Executed"]] or 0)
    report(counts, total, "executed warp instructions, %s" % rep)
    if thr:
        print("threads per executed warp instruction: %.2f" % (thr / total))


def lines(rep, cubin, kern, top=45):
    """executed warp instructions per CUDA source line (nvdisasm -g line table of the cubin that was
    profiled), split by opcode class"""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr_i = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
    ix = {h: i for i, h in enumerate(rows[hdr_i])}
    data = [r for r in rows[hdr_i + 1:] if len(r) > ix["Instructions Executed"]]
    base = int(data[0][ix["Address"]], 16)
    dyn = {int(r[ix["Address"]], 16) - base: (int(r[ix["Instructions Executed"]]), opcode(r[ix["Source"]])) for r in data}
    dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    line_re = re.compile(r'//## File "([^"]+)", line (\d+)')
    ins_re = re.compile(r"/\*([0-9a-f]{4,6})\*/\s+(.*?);")
    per, cur, on, total = collections.defaultdict(collections.Counter), ("?", 0), False, 0
    for ln in dis.split("\n"):
        if ln.startswith("//--------------------- .text."):
            on = kern in ln
            continue
        if not on:
            continue
        m = line_re.search(ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = ins_re.search(ln)
        if m and int(m.group(1), 16) in dyn:
            n, op = dyn[int(m.group(1), 16)]
            per[cur][classify(op)] += n
            total += n
    print("executed warp instructions per source line, %s (%d total)" % (kern, total))
    for key, c in sorted(per.items(), key=lambda kv: -sum(kv[1].values()))[:top]:
        n = sum(c.values())
        print("  %5.2f %%  %-22s %s" % (100.0 * n / total, "%s:%d" % key,
                                       " ".join("%s=%.2f" % (k, 100.0 * v / total) for k, v in c.most_common(5))))


if __name__ == "__main__":
    if len(sys.argv) >= 5 and sys.argv[1] == "lines":
        lines(sys.argv[2], sys.argv[3], sys.argv[4])
    elif len(sys.argv) >= 4 and sys.argv[1] == "static":
        static(sys.argv[2], sys.argv[3])
    elif len(sys.argv) >= 3 and sys.argv[1] == "dynamic":
        dynamic(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
    else:
        sys.exit(__doc__)
