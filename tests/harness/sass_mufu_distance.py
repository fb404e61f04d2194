"""Offline SASS check of the softmax schedule (no GPU): for the unmasked softmax tile of fa_fwd_kernel<128,1,false>, how many
instructions lie between each MUFU.EX2 and the first consumer of its result, and where the short ones sit.
   python tests/harness/sass_mufu_distance.py flash_attention_cuda_b200/libflashattn_b200.so"""
import re,collections,sys,subprocess,statistics
lib=sys.argv[1]
txt=subprocess.run(['cuobjdump','-sass',lib],capture_output=True,text=True).stdout
m=re.search(r'Function : \S*fa13fa_fwd_kernelILi128ELi1ELb0.*?(?=Function :)',txt,re.S)
lines=m.group(0).split('\n')
ins=[]
for l in lines:
    mm=re.match(r'\s+/\*[0-9a-f]{4,6}\*/\s+(.*?);',l)
    if mm: ins.append(mm.group(1).strip())
# unmasked softmax tile: the 4 consecutive-ish LDTM group whose following region has >= 90 MUFU before 2 SYNCS.ARRIVE: take the LAST such region start
idx=[i for i,t in enumerate(ins) if t.split()[0].startswith('LDTM') or (t.startswith('@') and 'LDTM' in t)]
best=None
for k in range(len(idx)):
    seg=ins[idx[k]:idx[k]+1400]
    cnt=0;end=None
    for j,t in enumerate(seg):
        if 'SYNCS.ARRIVE' in t:
            cnt+=1
            if cnt==2: end=j;break
    if end is None: continue
    seg=seg[:end+1]
    nm=sum(1 for t in seg if 'MUFU.EX2' in t); nx=sum(1 for t in seg if t.startswith('FMNMX3') or ' FMNMX3' in t)
    if nm>=90 and nx>=60 and not any('ISETP.GE.AND' in t and 'SEL' in t for t in seg):
        nsel=sum(1 for t in seg if t.split()[0]=='SEL' or ' SEL ' in t)
        if nsel<20: best=(idx[k],seg); break
start,seg=best
def regs(s): return [int(x) for x in re.findall(r'\bR(\d+)\b',s)]
pending={};dist=[]
for i,t in enumerate(seg):
    parts=t.split(None,1); op=parts[0]; rest=parts[1] if len(parts)>1 else ''
    if op.startswith('@'): op,rest=(rest.split(None,1)+[''])[:2]
    ops=[x.strip() for x in rest.split(',')]
    dst=regs(ops[0]) if ops else []
    wide=op in('FFMA2','FADD2','FMUL2')
    srcs=set()
    for o in ops[1:]:
        for r in regs(o):
            srcs.add(r)
            if wide: srcs.add(r+1)
    for r in list(pending):
        if r in srcs: dist.append(i-pending[r]); del pending[r]
    if op=='MUFU.EX2' and dst: pending[dst[0]]=i
c=collections.Counter(t.split()[1] if t.startswith('@') else t.split()[0] for t in seg)
print(lib.split('/')[-1],": softmax tile region",len(seg),"instructions; MUFU",c['MUFU.EX2'],"LDTM",sum(v for k,v in c.items() if k.startswith('LDTM')),"LDL",c.get('LDL',0),"STL",c.get('STL',0))
print("  MUFU -> first consumer distance: min",min(dist),"median",statistics.median(dist),"; within 8 instructions:",sum(1 for d in dist if d<8),"of",len(dist),"; within 16:",sum(1 for d in dist if d<16))
# positions of short-distance MUFUs (index in region, distance)
pending={};short=[]
for i,t in enumerate(seg):
    parts=t.split(None,1); op=parts[0]; rest=parts[1] if len(parts)>1 else ''
    if op.startswith('@'): op,rest=(rest.split(None,1)+[''])[:2]
    ops=[x.strip() for x in rest.split(',')]
    dst=regs(ops[0]) if ops else []
    wide=op in('FFMA2','FADD2','FMUL2')
    srcs=set()
    for o in ops[1:]:
        for r in regs(o):
            srcs.add(r)
            if wide: srcs.add(r+1)
    for r in list(pending):
        if r in srcs:
            if i-pending[r]<8: short.append((pending[r],i-pending[r],op))
            del pending[r]
    if op=='MUFU.EX2' and dst: pending[dst[0]]=i
sttm=[i for i,t in enumerate(seg) if 'STTM' in t]
print("  STTM at",sttm,"; short ones at (index,dist,consumer):",short)
