#!/usr/bin/env python3
"""Opcode-class histogram of one kernel.

  tools/sass_hist.py static  libort.so  <kernel substring>
        static instruction counts from `cuobjdump -sass`
  tools/sass_hist.py dynamic prof.ncu-rep [<kernel substring>]
        warp-level executed instruction counts from the SASS source page of an ncu report
        (`ncu --set full --import-source on`): every SASS instruction weighted by how often it ran

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
    ("lsu", r"^(LDS|STS|LDG|STG|LDL|STL|RED|REDG|ATOM|ATOMS|ATOMG|LD|ST|LDSM|MEMBAR|CCTL|ERRBAR)"),
    ("const", r"^(LDC|LDCU|ULDC)"),
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


if __name__ == "__main__":
    if len(sys.argv) >= 4 and sys.argv[1] == "static":
        static(sys.argv[2], sys.argv[3])
    elif len(sys.argv) >= 3 and sys.argv[1] == "dynamic":
        dynamic(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
    else:
        sys.exit(__doc__)
