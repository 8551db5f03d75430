#!/usr/bin/env python
"""Stall samples and executed warp-instructions of an .ncu-rep, summed per CUDA SOURCE LINE (inlined code included).

usage: python profiles/ncu_lines.py gpurun_out/prof.ncu-rep [N]      (needs --import-source on, -lineinfo)
"""
import collections
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    path, hdr = "?", None
    samples = collections.Counter()
    insts = collections.Counter()
    text = {}
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            path = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            hdr = r
            i_s, i_i = hdr.index("# Samples"), hdr.index("Instructions Executed")
        elif hdr and len(r) == len(hdr):
            key = (path, r[0])
            try:
                samples[key] += int(r[i_s] or 0)
                insts[key] += int(r[i_i] or 0)
            except ValueError:
                continue
            text[key] = r[1].strip()
    total_s, total_i = sum(samples.values()) or 1, sum(insts.values()) or 1
    print("%d samples, %d warp instructions; hottest source lines:" % (total_s, total_i))
    for key, n in samples.most_common(top):
        print("  %5.1f%% samples %5.1f%% insts  %s:%s  %s" % (100.0 * n / total_s, 100.0 * insts[key] / total_i, key[0], key[1], text[key][:110]))


if __name__ == "__main__":
    main()
