"""Aggregate the ncu source page (SASS level) of one kernel: samples per opcode, stall mix, smem wavefronts."""
import csv, sys, collections, subprocess, io
rep, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# first kernel block only
hdr = None; data = []
for r in rows:
    if r and r[0] == "Address":
        if hdr is not None: break
        hdr = r; continue
    if hdr is not None and len(r) == len(hdr): data.append(r)
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix['# Samples']]) for r in data)
print("instructions", len(data), "samples", tot)
agg = collections.Counter(); cnt = collections.Counter(); exe = collections.Counter(); st = collections.Counter()
wf = collections.Counter(); wfi = collections.Counter()
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
def opof(src):
    p = src.strip().split()
    op = p[1] if p[0].startswith('@') else p[0]
    return op.split('.')[0]
for r in data:
    op = opof(r[ix['Source']])
    s = int(r[ix['# Samples']]); agg[op] += s; cnt[op] += 1; exe[op] += int(r[ix['Instructions Executed']])
    for c in stall_cols: st[c] += int(r[ix[c]] or 0)
    wf[op] += int(r[ix['L1 Wavefronts Shared']] or 0); wfi[op] += int(r[ix['L1 Wavefronts Shared Ideal']] or 0)
print("op static executed samples share")
for op, s in agg.most_common(16): print(op, cnt[op], exe[op], s, "%.1f%%" % (100 * s / tot))
print("stalls", [(k, v) for k, v in st.most_common(9)])
print("smem wavefronts", dict(wf), "ideal", dict(wfi))
if len(sys.argv) > 3:
    top = sorted(data, key=lambda r: -int(r[ix['# Samples']]))[:int(sys.argv[3])]
    for r in top: print(r[ix['# Samples']], r[ix['Source']].strip()[:90], [ (c,r[ix[c]]) for c in stall_cols if int(r[ix[c]] or 0) > 0.3*int(r[ix['# Samples']])])
