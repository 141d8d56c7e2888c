import re,sys,csv,subprocess,collections
rep=sys.argv[1]; gfile=sys.argv[2]; fn=sys.argv[3]
raw=subprocess.run(["ncu","-i",rep,"--page","source","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines())); hdr=rows[1]; ci={h:i for i,h in enumerate(hdr)}
dyn=[]
for r in rows[2:]:
    try: dyn.append((float(r[ci["Instructions Executed"]]), float(r[ci["# Samples"]]), float(r[ci["Avg. Threads Executed"]] or 0), r[ci["Source"]]))
    except Exception: pass
# static mapping
lines=open(gfile).read().splitlines()
start=None
for i,l in enumerate(lines):
    if l.startswith('\t.section\t.text.'+fn): start=i
    elif start is not None and l.startswith('\t.section') and i>start: end=i; break
cur=None; stat=[]
stack=[]
for l in lines[start:end]:
    m=re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?',l)
    if m: cur=int(m.group(2)); continue
    if re.match(r'\s*/\*[0-9a-f]{4,5}\*/\s+\S',l): stat.append((cur,l.strip()))
print(len(dyn),len(stat))
assert len(dyn)==len(stat)
tot=sum(d[0] for d in dyn); tots=sum(d[1] for d in dyn)
byline=collections.defaultdict(lambda:[0,0])
for (ex,smp,thr,src),(ln,txt) in zip(dyn,stat):
    byline[ln][0]+=ex; byline[ln][1]+=smp
iters=float(sys.argv[4]) if len(sys.argv)>4 else 1
for ln,(ex,smp) in sorted(byline.items(), key=lambda x:-x[1][0])[:45]:
    print("line %5s instr%%=%5.1f per-iter=%6.1f time%%=%5.1f"%(ln,100*ex/tot,ex/iters,100*smp/tots))
