"""Join the SASS-level samples of an `ncu --import-source on` capture with source lines.

  ncu -i rep.ncu-rep --page source --csv --kernel-name regex:^k_pcr$ --print-source sass > pcr.csv
  nvcc -cubin ... -lineinfo k_pcr.cu   (or cuobjdump -xelf all librlsim_gpu.so) - the build the capture ran
  python profiles/linejoin.py k_pcrENS pcr.csv '/tmp/cub/*.cubin' 40 [source dir]

Prints the hottest source lines: share of the warp-state samples, samples, active lanes per instruction."""
import csv,re,subprocess,sys,collections,glob
kernel=sys.argv[1]; srccsv=sys.argv[2]; cubglob=sys.argv[3]
# sass samples by address
rows=list(csv.reader(open(srccsv)))
hdr=rows[1]; ai=hdr.index('Address'); smp=hdr.index('# Samples'); ie=hdr.index('Instructions Executed'); te=hdr.index('Thread Instructions Executed')
samples={}
addrs=[]
for r in rows[2:]:
    try: a=int(r[ai],16) if r[ai].lower().startswith('0x') else int(r[ai])
    except: continue
    addrs.append(a); samples[a]=(int(r[smp] or 0), int(r[ie] or 0), int(r[te] or 0))
base=min(addrs)
line_of={}
for cub in glob.glob(cubglob):
    out=subprocess.run(['nvdisasm','-g','-c',cub],capture_output=True,text=True).stdout
    if kernel not in out: continue
    infn=False; cur=None
    for ln in out.split('\n'):
        if ln.startswith('.text.') or re.match(r'\s*\.section\s+\.text\.',ln):
            infn = kernel in ln
        if not infn: continue
        m=re.search(r'//## File "([^"]+)", line (\d+)(.*)',ln)
        if m:
            cur=(m.group(1).split('/')[-1],int(m.group(2)),m.group(3)); continue
        m=re.search(r'/\*([0-9a-f]{4,})\*/',ln)
        if m and cur: line_of[int(m.group(1),16)]=cur
agg=collections.defaultdict(lambda:[0,0,0])
tot=0
for a,(n,i,t) in samples.items():
    k=line_of.get(a-base,('?',0,''))
    key=(k[0],k[1]); agg[key][0]+=n; agg[key][1]+=i; agg[key][2]+=t; tot+=n
src={}
for (f,l),v in sorted(agg.items(), key=lambda kv:-kv[1][0])[:int(sys.argv[4]) if len(sys.argv)>4 else 40]:
    if f not in src:
        try: src[f]=open((sys.argv[5] if len(sys.argv)>5 else '/root/repo/rlsim_b200/csrc')+'/'+f).read().split('\n')
        except: src[f]=[]
    text=src[f][l-1].strip()[:100] if 0<l<=len(src[f]) else ''
    print('%5.1f%% %7d  lanes %4.1f  %s:%d  %s'%(100*v[0]/tot, v[0], v[2]/max(v[1],1), f,l,text))
