"""Static instruction counts per sweep kernel of two builds of mif_poisson.o (cuobjdump -sass): python scripts/sass_count.py before.o after.o"""
import re, subprocess, sys, collections
def counts(obj):
    out=subprocess.run(["cuobjdump","-sass",obj],capture_output=True,text=True).stdout
    res={}; cur=None
    for line in out.splitlines():
        m=re.search(r"Function : (\S+)", line)
        if m:
            cur=subprocess.run(["c++filt",m.group(1)],capture_output=True,text=True).stdout.strip()
            cur=cur.replace("(anonymous namespace)::","").replace("void ","").replace("mifgpu::","")
            cur=re.sub(r"\(.*","",cur)
            res[cur]=collections.Counter(); continue
        m=re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if m and cur: res[cur][m.group(1).split(".")[0]]+=1
    return res
a=counts(sys.argv[1]); b=counts(sys.argv[2])
keys=["DADD","DMUL","DFMA","LDS","STS","SHFL","FSEL","SEL","LOP3","MOV","IMAD","LDL","STL"]
for k in b:
    if not any(s in k for s in ("dct512","dct1024")): continue
    ca,cb=a.get(k,{}),b[k]
    fa=sum(ca.get(x,0) for x in ("DADD","DMUL","DFMA")); fb=sum(cb.get(x,0) for x in ("DADD","DMUL","DFMA"))
    print(f"{k:42s} FP64 {fa:5d} -> {fb:5d}   total {sum(ca.values()):5d} -> {sum(cb.values()):5d}   " + " ".join(f"{x} {ca.get(x,0)}->{cb.get(x,0)}" for x in ("SHFL","LDS","STS","FSEL","LOP3","LDL","STL")))
