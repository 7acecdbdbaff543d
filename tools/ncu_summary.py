import csv, sys
from collections import Counter
raw, src = sys.argv[1], sys.argv[2]
rows=list(csv.reader(open(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d={h:(u,v) for h,u,v in zip(hdr,units,vals)}
keys=['gpu__time_duration.sum','sm__cycles_active.avg','launch__registers_per_thread','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum',
'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed','dram__bytes_read.sum','dram__bytes_write.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct','smsp__sass_inst_executed_op_tmem_ldt.sum']
for k in keys:
    if k in d: print(f"{k:85s} {d[k]}")
items=[(h,float(d[h][1].replace(',',''))) for h in hdr if 'pcsamp_warps_issue_stalled' in h and not h.endswith('not_issued') and d[h][1] not in ('','n/a')]
items.sort(key=lambda x:-x[1]); tot=sum(v for _,v in items)
for h,v in items[:8]: print(f"{h:80s} {v:12.0f} {v/tot:6.3f}")
rows=list(csv.reader(open(src)))
hdr=rows[1]; body=rows[2:]
isrc=hdr.index('Source'); isamp=hdr.index('# Samples'); iex=hdr.index('Instructions Executed')
def s(i): return int(body[i][isamp])
tot=sum(s(i) for i in range(len(body)))
print('total samples',tot,'instructions',len(body))
thr=float(sys.argv[3]) if len(sys.argv)>3 else 0.012
for a in range(0,len(body),50):
    b=min(a+50,len(body)); sm=sum(s(i) for i in range(a,b))
    if sm/tot<thr: continue
    names=[(body[i][isrc].split()[1] if body[i][isrc].strip().startswith('@') else body[i][isrc].split()[0]) for i in range(a,b)]
    c=Counter(n.split('.')[0] for n in names)
    top=max(range(a,b), key=s)
    print(f"{a:5d}-{b:5d} samples={sm:6d} {sm/tot:6.3f} ex={body[a][iex]:>8s} {dict(c.most_common(4))} top: {s(top)} {body[top][isrc].strip()[:60]}")
