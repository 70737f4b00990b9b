#!/usr/bin/env python3
"""Executed warp instructions and stall samples of one kernel aggregated by CUDA source line
(ncu source page, report captured with --import-source on)."""
import csv, sys, collections, subprocess
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass,cuda'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[2]
iex, isamp = hdr.index('Instructions Executed'), hdr.index('# Samples')
def num(x):
    try: return int(x)
    except ValueError: return 0
per = collections.Counter(); samp = collections.Counter(); text = {}
for r in rows[3:]:
    if len(r) <= iex: continue
    try: ln = int(r[0])
    except ValueError: continue
    per[ln] += num(r[iex]); samp[ln] += num(r[isamp]); text[ln] = r[1].strip()[:88]
tot, ts = sum(per.values()), sum(samp.values())
print(f"executed warp instructions {tot}, samples {ts}")
key = samp if len(sys.argv) > 3 and sys.argv[3] == 'samples' else per
for ln, _ in key.most_common(top):
    print(f"{ln:5d} instr {100*per[ln]/tot:5.2f}%  samples {100*samp[ln]/ts:5.2f}%  {text[ln]}")
