"""Joins `nvdisasm --print-line-info` SASS with an `ncu --page source --csv` dump and aggregates executed
warp-instructions and stall samples per source line.
Usage: ncu_by_line.py <kernel-substr> <unused> <top-n> <sass.txt> <source.csv>"""
import re,csv,collections,sys
kern=sys.argv[1]; srcfile=sys.argv[2]
cur=None; seq=[]; infn=False
for ln in open(sys.argv[4]):
    if ln.startswith('.text.'): infn = kern in ln; continue
    if not infn: continue
    m=re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur=(m.group(1).split('/')[-1], int(m.group(2)))
    m2=re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
    if m2: seq.append((cur, m2.group(2)))
rows=list(csv.reader(open(sys.argv[5])))
for i,r in enumerate(rows):
    if 'Source' in r and 'Instructions Executed' in r: hdr=r; start=i+1; break
ie=hdr.index('Instructions Executed'); st=hdr.index('Warp Stall Sampling (All Samples)')
counts=[]
for r in rows[start:]:
    try: counts.append((int(r[ie]), int(r[st] or 0)))
    except: pass
print(len(seq),'sass;',len(counts),'ncu')
agg=collections.Counter(); stall=collections.Counter()
for i in range(min(len(seq),len(counts))):
    agg[seq[i][0]]+=counts[i][0]; stall[seq[i][0]]+=counts[i][1]
tot=sum(agg.values()); stot=sum(stall.values())
src={}
import glob,os
for f in glob.glob('/root/repo/arkflow_b200/csrc/*'):
    src[os.path.basename(f)]=open(f,errors='ignore').read().split('\n')
print('total warp-instr',tot)
for (k,v) in agg.most_common(int(sys.argv[3])):
    f,l=k if k else ('?',0)
    text=src[f][l-1].strip()[:100] if f in src and l-1 < len(src[f]) else ''
    print(f"{v:10d} {100*v/tot:5.1f}% stall {100*stall[k]/max(stot,1):5.1f}%  {f}:{l}  {text}")
print("---- by stall share ----")
for (k,v) in stall.most_common(14):
    f,l=k if k else ('?',0)
    text=src[f][l-1].strip()[:100] if f in src and l-1 < len(src[f]) else ''
    print(f"stall {100*v/max(stot,1):5.1f}%  instr {100*agg[k]/tot:5.1f}%  {f}:{l}  {text}")
