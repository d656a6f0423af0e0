"""Per-shape tcgen05 kernel report from a koa_profile_dump file: time, TFLOP/s, and the ideal time by max(flops/peak, bytes/bw)."""
import sys
rows=[]
for l in open(sys.argv[1]):
    if l.startswith('#'): continue
    c,tag,m,n,k,cnt,ms,tf=l.split()
    c,tag,m,n,k,cnt=map(int,(c,tag,m,n,k,cnt)); ms=float(ms); tf=float(tf)
    if c==0:
        by=(m*k+n*k+m*n*(1+(1 if tag&4 else 0)+(1 if tag&8 else 0)+(1 if tag&128 else 0)))*2
    else:
        by=(k*m+k*n/(9 if tag&1 else 1))*2+m*n*4
    fl=2*m*n*k
    ideal=max(fl/1386e12, by/6536e9)*1e3
    rows.append((ms,c,tag,m,n,k,cnt,tf,ideal*cnt, by*cnt/ms/1e6))
rows.sort(reverse=True)
tot=sum(r[0] for r in rows); ti=sum(r[8] for r in rows)
print('total ms %.2f ideal %.2f'%(tot,ti))
top=int(sys.argv[2]) if len(sys.argv)>2 else 30
for r in rows[:top]: print('%.3f ms cls%d tag%3d m=%7d n=%5d k=%6d x%2d  %4.0f TF  ideal %.3f ms  %.0f GB/s'%r)
