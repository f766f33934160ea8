#!/usr/bin/env python
"""Attribute the instructions and stall samples of an ncu capture to source lines (no GPU needed):
    python tools/ncu_lines.py <report.ncu-rep> <cubin of the same build> <kernel name fragment> <source file> [top N]
The cubin comes from `cuobjdump -xelf all build/<unit>.o`; it must be the build the capture ran (the script counts opcode
mismatches between the report and the cubin)."""
import csv,re,sys,subprocess
rep, cubin, kern_pat, src_path = sys.argv[1:5]
topn = int(sys.argv[5]) if len(sys.argv)>5 else 40
out=subprocess.run(["nvdisasm","--print-line-info",cubin],capture_output=True,text=True).stdout.split('\n')
cur=None; off2line={}; off2ins={}; f=False
for l in out:
    if l.startswith('.text.'):
        f = kern_pat in l
        continue
    if not f: continue
    m=re.search(r'//## File "([^"]+)", line (\d+)',l)
    if m:
        cur=(m.group(1).split('/')[-1],int(m.group(2))); continue
    m=re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);',l)
    if m and cur:
        off2line[int(m.group(1),16)]=cur; off2ins[int(m.group(1),16)]=m.group(2)
csvt=subprocess.run(["ncu","-i",rep,"--page","source","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(csvt.split('\n')))
hdr=rows[1]; data=[]
for r in rows[2:]:
    if r and r[0]=="Kernel Name": break
    if len(r)==len(hdr): data.append(r)
ix={h:i for i,h in enumerate(hdr)}
def g(r,k):
    try: return float(r[ix[k]])
    except: return 0.0
base=int(data[0][ix["Address"]],16)
mism=sum(1 for r in data if r[ix["Source"]].strip().split()[-0:1]!=off2ins.get(int(r[ix["Address"]],16)-base,"").split()[0:1] and not r[ix["Source"]].strip().startswith('@'))
print("instr",len(data),"opcode mismatches",mism)
agg={}
for r in data:
    ln=off2line.get(int(r[ix["Address"]],16)-base,("?",0))
    a=agg.setdefault(ln,[0,0,0,0]); a[0]+=g(r,"Instructions Executed"); a[1]+=g(r,"# Samples"); a[2]+=g(r,"stall_short_sb"); a[3]+=g(r,"stall_long_sb")
ti=sum(v[0] for v in agg.values()); ts=sum(v[1] for v in agg.values())
print("exec",ti,"samples",ts)
src=open(src_path).read().split('\n')
for ln,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:topn]:
    text=src[ln[1]-1].strip()[:84] if ln[0]==src_path.split('/')[-1] and ln[1]>0 else ln[0]
    print(ln[0][:10],ln[1],"exec %.1f%%"%(100*v[0]/ti),"samp %.1f%%"%(100*v[1]/ts),"ssb %.1f%% lsb %.1f%%"%(100*v[2]/ts,100*v[3]/ts),"|",text)
