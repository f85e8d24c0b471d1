#!/usr/bin/env python3
"""Dynamic SASS profile from an ncu report's source page: executed warp-instructions per opcode and the hottest lines.

    python tools/ncu_source.py gpurun_out/prof.ncu-rep [--top 40] [--dump out.txt]
"""
import argparse
import collections
import csv
import io
import subprocess


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--dump", default="")
    a = ap.parse_args()
    out = subprocess.run(["ncu", "-i", a.rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    c_src, c_ex, c_smp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    ops = collections.Counter()
    samples = collections.Counter()
    total = 0
    lines = []
    for r in rows[hi + 1:]:
        if len(r) <= c_ex:
            continue
        try:
            ex = int(r[c_ex])
        except ValueError:
            continue
        src = r[c_src].strip()
        toks = src.split()
        op = toks[0]
        if op.startswith("@") and len(toks) > 1:
            op = toks[1]
        op = op.split(".")[0].rstrip(";")
        ops[op] += ex
        smp = int(r[c_smp] or 0)
        samples[op] += smp
        total += ex
        stalls = sorted(((int(r[i] or 0), h) for i, h in stall_cols), reverse=True)[:2]
        lines.append((ex, smp, src, stalls))
    print("total warp-instructions executed: %d" % total)
    tot_s = sum(samples.values())
    for op, n in ops.most_common(a.top):
        print("  %-10s %12d  %5.1f%%   samples %5.1f%%" % (op, n, 100.0 * n / total, 100.0 * samples[op] / max(tot_s, 1)))
    if a.dump:
        with open(a.dump, "w") as f:
            for ex, smp, src, stalls in lines:
                f.write("%9d %6d  %-70s %s\n" % (ex, smp, src, " ".join("%s=%d" % (h, v) for v, h in stalls if v)))


if __name__ == "__main__":
    main()
