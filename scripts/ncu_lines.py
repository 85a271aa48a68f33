"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump per source line (samples, instructions)."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
def I(x):
    try: return int(x)
    except ValueError: return 0
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
by = 1 if (len(sys.argv) > 3 and sys.argv[3] == 'inst') else 0  # sort by stall samples (default) or by instructions
cur_file = None; cur_line = None; cur_src = ''
agg = collections.OrderedDict()
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]; continue
    if len(r) < 8 or r[0] == 'Line No':
        continue
    if r[0] != '':
        cur_line = int(r[0]); cur_src = r[1].strip()
        continue
    key = (cur_file, cur_line)
    a = agg.setdefault(key, [0, 0, 0, cur_src])
    a[0] += I(r[4]); a[1] += I(r[7]); a[2] += I(r[8])
tot_s = sum(a[0] for a in agg.values()); tot_i = sum(a[1] for a in agg.values())
print('total samples', tot_s, 'total warp-inst', tot_i)
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][by])[:top]:
    thr = a[2] / a[1] if a[1] else 0
    print(f'{k[0]}:{k[1]:5d} samp {100*a[0]/tot_s:5.1f}% inst {100*a[1]/tot_i:5.1f}% thr {thr:4.1f} | {a[3][:90]}')
