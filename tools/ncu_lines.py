import csv,io,subprocess,sys
rep=sys.argv[1]
out=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","cuda,sass"],capture_output=True,text=True).stdout
cur=None; rows=[]
for r in csv.reader(io.StringIO(out)):
    if len(r)==2 and r[0]=="File Path": cur=r[1].split("/")[-1]; continue
    if len(r)>6 and r[0].isdigit():
        try: rows.append((int(r[6] or 0), cur, int(r[0]), r[1].strip()[:110], int(r[7] or 0)))
        except Exception: pass
tot=sum(x[0] for x in rows) or 1
for s,f,l,src,inst in sorted(rows,reverse=True)[:int(sys.argv[2]) if len(sys.argv)>2 else 30]:
    print(f"{100*s/tot:5.1f}%  {f}:{l}  inst={inst}  {src}")
