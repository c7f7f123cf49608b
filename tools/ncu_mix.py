import csv,collections,re,subprocess,sys
rep=sys.argv[1]; which=int(sys.argv[2]) if len(sys.argv)>2 else 0
raw=subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines())); hdr=rows[0]
want=["gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
"sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active","smsp__inst_executed.sum","smsp__issue_active.avg.pct_of_peak_sustained_active","l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum","l1tex__data_pipe_lsu_wavefronts_mem_shared.sum","sm__warps_active.avg.pct_of_peak_sustained_active","launch__registers_per_thread","launch__grid_size"]
stalls=[h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
for r in rows[2:]:
    print("="*60); print(r[hdr.index("Kernel Name")][:70])
    for k in want:
        if k in hdr: print(f"  {k:70s} {r[hdr.index(k)]}")
    st=sorted(((float(r[hdr.index(k)].replace(',','')),k) for k in stalls), reverse=True)[:8]
    print("   stalls:", ", ".join(f"{k.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio','')}={v:.2f}" for v,k in st))
src=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","sass","--launch-skip",str(which),"--launch-count","1"],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
st=[i for i,r in enumerate(rows) if r and r[0]=="Address"][0]
h=rows[st]; ia=h.index("Source"); ie=h.index("Instructions Executed"); isamp=h.index("# Samples")
ops=collections.Counter(); samp=collections.Counter(); tot=0; ts=0
for r in rows[st+1:]:
    try: n=int(r[ie]); s=int(r[isamp])
    except: continue
    m=re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_]+)',r[ia].strip()); op=m.group(2) if m else '?'
    ops[op]+=n; samp[op]+=s; tot+=n; ts+=s
print("total instr",tot)
for op,n in ops.most_common(22): print(f"{op:10s} {100*n/tot:5.1f}%  samples {100*samp[op]/ts:5.1f}%")
