import csv, re, sys, collections
csv_path, kern, asm_path = sys.argv[1:4]
lines_of=[]; cur=None; infunc=False
for l in open(asm_path, errors="replace"):
    if l.startswith(".text.") or l.lstrip().startswith(".section\t.text."):
        infunc = kern in l; continue
    if not infunc: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur=(m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l): lines_of.append(cur)
rows=list(csv.reader(open(csv_path))); hdr=None; data=[]; started=False
for r in rows:
    if r and r[0]=="Kernel Name":
        if started: break
        continue
    if r and r[0]=="Address": hdr=r; started=True; continue
    if started: data.append(r)
ci={n:i for i,n in enumerate(hdr)}
S,X,T=ci["# Samples"],ci["Instructions Executed"],ci["Thread Instructions Executed"]
per=collections.defaultdict(lambda:[0,0,0]); tot=[0,0,0]
for k in range(min(len(data),len(lines_of))):
    v=(int(data[k][S] or 0),int(data[k][X] or 0),int(data[k][T] or 0))
    for j in range(3): per[lines_of[k]][j]+=v[j]; tot[j]+=v[j]
# per-file dump sorted by line
byfile=collections.defaultdict(list)
for (f,ln),v in per.items(): byfile[f].append((ln,v))
for f in sorted(byfile, key=lambda f:-sum(v[1] for _,v in byfile[f])):
    tf=sum(v[1] for _,v in byfile[f])
    print(f"### {f}: inst {100*tf/tot[1]:.1f}%  samples {100*sum(v[0] for _,v in byfile[f])/tot[0]:.1f}%")
    if '-v' in sys.argv:
        for ln,v in sorted(byfile[f]):
            if v[1]/tot[1]>0.002: print(f"    {ln:5d} inst {100*v[1]/tot[1]:5.2f}% samples {100*v[0]/tot[0]:5.2f}% thr {v[2]/max(v[1],1):.1f}")
