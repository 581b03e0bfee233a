import csv,sys
rows=list(csv.reader(open('/tmp/fr_raw.csv')))
hdr=rows[0]; units=rows[1]; vals=rows[2]
keys=('gpu__time_duration.sum','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','lts__t_sectors_srcunit_tex_op_write.sum','l1tex__t_requests_pipe_lsu_mem_global_op_st.sum','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed')
for i,h in enumerate(hdr):
    if h in keys or ('issue_stalled' in h and 'per_issue_active' in h and float(vals[i] or 0)>0.2):
        print(f"{h:90s} {vals[i]}")
rows=list(csv.reader(open('/tmp/fr_src.csv')))
hdr=rows[1]
isrc=hdr.index("Source"); ie=hdr.index("Instructions Executed"); iss=hdr.index("# Samples")
from collections import Counter
c=Counter(); cs=Counter(); tot=0
out=[]
for idx,r in enumerate(rows[2:]):
    if len(r)<=ie or r[ie] in("",): continue
    n=int(r[ie]); out.append((idx,n,int(r[iss] or 0),r[isrc]))
    if n==0: continue
    op=r[isrc].split()[0] if not r[isrc].startswith('@') else r[isrc].split()[1]
    c[op.split('.')[0]]+=n; cs[op.split('.')[0]]+=int(r[iss] or 0); tot+=n
print("total",tot)
for op,n in c.most_common(12): print(f"{op:12s} {n:12d} {100*n/tot:5.1f}%  samples {cs[op]}")
runs=[]
for idx,n,s,src in out:
    if runs and abs(runs[-1][1]-n)<=0.02*max(n,1): runs[-1][2]+=1; runs[-1][3]+=s; runs[-1][5]=idx; runs[-1][6]+=n
    else: runs.append([idx,n,1,s,src,idx,n])
for r in [r for r in runs if r[6]>8_000_000 or r[3]>3000]: print(f"sass#{r[0]:5d}-{r[5]:5d} exec/inst {r[1]:10d} x {r[2]:4d} insts = {r[6]/1e6:8.1f}M samples {r[3]:6d}  first: {r[4][:60]}")
