#!/usr/bin/env python3
"""Opcode histogram (executed warp instructions) of one kernel from ncu's SASS source page,
split into the regions between BAR.SYNC instructions (= the kernel's phases)."""
import csv, sys, collections, subprocess
rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ia, isrc, iex, ismp = hdr.index('Address'), hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
tot = collections.Counter(); reg = collections.defaultdict(collections.Counter); regsamp = collections.Counter()
region = 0; total = 0; n_static = 0
for r in rows[2:]:
    if len(r) <= iex or not r[iex]: continue
    src = r[isrc].strip()
    toks = src.split()
    if not toks: continue
    op = toks[1] if toks[0].startswith('@') else toks[0]
    base = op.split('.')[0]
    key = base
    if base in ('LDG','STG','LDS','STS','ATOMS','LDL','STL','IMAD','DADD','DMUL','DFMA','SHF','LOP3','PRMT','ISETP','IADD3','IADD','LEA','SEL','BRA','MOV','SHFL'):
        key = op if base in ('LDG','STG','LDS','STS','LDL','STL') else base
    ex = int(r[iex]); total += ex; n_static += 1
    tot[key] += ex; reg[region][key] += ex; regsamp[region] += int(r[ismp] or 0)
    if base == 'BAR': region += 1
print(f"static SASS instructions: {n_static}, executed warp instructions: {total}")
print("== whole kernel ==")
for k, v in tot.most_common(28): print(f"  {k:24s} {v:12d} {100*v/total:6.2f}%")
for rg in sorted(reg):
    s = sum(reg[rg].values())
    print(f"== region {rg} (up to BAR #{rg+1}): {s} = {100*s/total:.1f}% of instructions, {regsamp[rg]} samples ==")
    print("   " + ", ".join(f"{k}:{100*v/max(s,1):.0f}%" for k, v in reg[rg].most_common(12)))
