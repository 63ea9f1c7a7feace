import re,collections,sys
src={f:open('curobo_b200/csrc/'+f).read().split('\n') for f in ('cb200_kernels.cu','cb200_warp.cuh','cb200_math.cuh')}
def func_of(f,l):
    if f not in src: return f
    for i in range(l-1, -1, -1):
        t=src[f][i]
        m=re.match(r'^(?:static |CB_HD |__device__ |__global__ |__host__ |inline |__forceinline__ |__noinline__ )+.*?([A-Za-z_0-9]+)\(', t)
        if m and not t.startswith(' '): return m.group(1)
    return f
agg=collections.Counter(); lanes=collections.Counter()
for line in open(sys.argv[1]):
    m=re.match(r'\s*([\d.]+)\s+([\d.]+)%\s+([\d.]+)\s+(\S+):(\d+)',line)
    if not m: continue
    c=float(m.group(1)); ln=float(m.group(3)); f=m.group(4); l=int(m.group(5))
    k=func_of(f,l); agg[k]+=c; lanes[k]+=c*ln
for k,v in agg.most_common(16): print(f"{v:9.1f} instr/row  lanes={lanes[k]/v:5.1f}  {k}")
print("total", sum(agg.values()))
