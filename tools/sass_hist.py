#!/usr/bin/env python3
"""Static SASS opcode histogram of selected kernels of a built object / library (no GPU needed).

    tools/sass_hist.py lte-gnu-radio-code_b200/csrc/build/inst_n11.o 'tx_kernel<.*11, 16.*6, 1, false, true>' [...]

Prints, per matching kernel: instruction count, the FP32 scalar vs packed (FADD2/FMUL2/FFMA2) split, memory and
barrier instruction counts and the top opcodes.  The kernels are unrolled straight-line code inside their persistent
loops, so the static count tracks the per-symbol instruction count that bounds the N >= 1024 sizes.
`--json` emits one JSON object per kernel (used for the listings under profiles/)."""
import collections
import json
import re
import subprocess
import sys


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return [o.replace("ofdm::", "") for o in out]


def parse(path):
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    funcs, cur = [], None
    for line in txt.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = [m.group(1), collections.Counter()]
            funcs.append(cur)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            cur[1][m.group(1)] += 1
    names = demangle([f[0] for f in funcs])
    return [(n, f[1]) for n, f in zip(names, funcs)]


def summarise(name, ops):
    def tot(pred):
        return sum(v for k, v in ops.items() if pred(k))
    base = lambda k: k.split(".")[0]
    s = {
        "kernel": name[:200],
        "instructions": sum(ops.values()) - ops.get("NOP", 0),
        "fp32_scalar": tot(lambda k: base(k) in ("FADD", "FMUL", "FFMA")),
        "fp32_packed": tot(lambda k: base(k) in ("FADD2", "FMUL2", "FFMA2")),
        "FADD": tot(lambda k: base(k) == "FADD"), "FMUL": tot(lambda k: base(k) == "FMUL"),
        "FFMA": tot(lambda k: base(k) == "FFMA"), "FADD2": tot(lambda k: base(k) == "FADD2"),
        "FMUL2": tot(lambda k: base(k) == "FMUL2"), "FFMA2": tot(lambda k: base(k) == "FFMA2"),
        "MOV": tot(lambda k: base(k) in ("MOV", "IMAD") and (k == "MOV" or k.startswith("IMAD.MOV"))),
        "LDS": tot(lambda k: base(k) == "LDS"), "STS": tot(lambda k: base(k) == "STS"),
        "LDG": tot(lambda k: base(k) in ("LDG", "LDGSTS")), "STG": tot(lambda k: base(k) == "STG"),
        "BAR": tot(lambda k: base(k) == "BAR"), "SHFL": tot(lambda k: base(k) == "SHFL"),
        "bulk_async": tot(lambda k: base(k) in ("UBLKCP", "UBLKRED", "UTMALDG", "UTMASTG")),
        "tensor": tot(lambda k: base(k) in ("HMMA", "UTCMMA", "UTCHMMA", "IMMA", "DMMA")),
    }
    bases = collections.Counter()
    for k, v in ops.items():
        bases[base(k)] += v
    s["top"] = dict(bases.most_common(14))
    return s


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    as_json = "--json" in sys.argv
    path, pats = args[0], [re.compile(p) for p in args[1:]] or [re.compile(".")]
    for name, ops in parse(path):
        if not any(p.search(name) for p in pats):
            continue
        s = summarise(name, ops)
        if as_json:
            print(json.dumps(s))
        else:
            print("%s\n   instr %d | FP32 scalar %d (FADD %d FMUL %d FFMA %d) packed %d (FADD2 %d FMUL2 %d FFMA2 %d) | MOV %d | "
                  "LDS %d STS %d LDG %d STG %d BAR %d SHFL %d bulk %d" % (
                      s["kernel"], s["instructions"], s["fp32_scalar"], s["FADD"], s["FMUL"], s["FFMA"], s["fp32_packed"],
                      s["FADD2"], s["FMUL2"], s["FFMA2"], s["MOV"], s["LDS"], s["STS"], s["LDG"], s["STG"], s["BAR"],
                      s["SHFL"], s["bulk_async"]))


if __name__ == "__main__":
    main()
