"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by CUDA
source line: stall samples, instructions, shared-memory wavefronts."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = None
agg = {}
cur = None
for r in rows:
    if r and r[0] == 'Line No':
        hdr = r
        i_samp = hdr.index('# Samples'); i_inst = hdr.index('Instructions Executed')
        i_wf = hdr.index('L1 Wavefronts Shared'); i_ex = hdr.index('L1 Wavefronts Shared Excessive')
        continue
    if hdr is None or len(r) < len(hdr) - 2:
        continue
    if r[0] != '':
        cur = (int(r[0]), r[1].strip())
        agg.setdefault(cur, [0, 0, 0, 0])
        continue
    if cur is None: continue
    try:
        a = agg[cur]
        a[0] += int(r[i_samp]); a[1] += int(r[i_inst]); a[2] += int(r[i_wf]); a[3] += int(r[i_ex])
    except ValueError:
        pass
tot = sum(a[0] for a in agg.values()) or 1
toti = sum(a[1] for a in agg.values()) or 1
print('total samples', tot, 'instructions', toti)
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for (ln, src), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print('%5.1f%% samp %5.1f%% inst smemwf=%10d exc=%10d  L%-4d %s' % (100.*a[0]/tot, 100.*a[1]/toti, a[2], a[3], ln, src[:100]))
