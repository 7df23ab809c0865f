#!/usr/bin/env python
"""Top CUDA source lines of a kernel by stall samples / instructions, from
`ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:K > X.csv`."""
import csv
import sys


def main(path, top=25):
    rows = list(csv.reader(open(path)))
    cur_file = None
    out = []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] in ("Line No", "Function Name") or len(r) < 8 or r[2] != "-":
            continue
        try:
            out.append((int(r[6] or 0), int(r[7] or 0), cur_file, int(r[0]), r[1].strip()[:110]))
        except ValueError:
            pass
    tot_s = sum(o[0] for o in out) or 1
    tot_i = sum(o[1] for o in out) or 1
    print(f"total samples {tot_s}, total warp instructions {tot_i}")
    print("--- by stall samples")
    for s, i, f, ln, src in sorted(out, reverse=True)[:top]:
        print(f"{100*s/tot_s:5.1f}% smp {100*i/tot_i:5.1f}% ins  {f}:{ln}  {src}")
    print("--- by instructions")
    for s, i, f, ln, src in sorted(out, key=lambda o: -o[1])[:top]:
        print(f"{100*i/tot_i:5.1f}% ins {100*s/tot_s:5.1f}% smp  {f}:{ln}  {src}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
