import sys
rows = [l.split() for l in open(sys.argv[1]) if not l.startswith('#')]
L=int(sys.argv[2]) if len(sys.argv)>2 else 200
agg={}; prev=None
for r in rows:
    enc=int(r[1],16); fl=int(r[2],16)
    q=[None]+[int(x) for x in r[3:]]
    child=enc&0x07ffffff
    kind='leaf' if child<L else ('mul' if enc&(1<<28) else ('wait' if enc&(1<<30) else ('chain' if enc&(1<<29) else 'int')))
    if fl&(1<<29): kind+='+last'
    d=agg.setdefault(kind,{'n':0})
    d['n']+=1
    def add(k,v): d[k]=d.get(k,0)+v
    if q[2]>0:
        add('stage',q[2]-q[1]); add('land',q[3]-q[2]); add('bar',q[4]-q[3]); add('mm',q[5]-q[4]); add('bar2',q[6]-q[5]); add('tail',q[7]-q[6])
    else:
        add('body',q[7]-q[1])
    end=q[7]
    if fl&(1<<29):
        add('renorm+store',q[8]-q[7]); add('bar3',q[9]-q[8]); add('publish',q[10]-q[9]); end=q[10]
    if prev is not None: add('gap',q[1]-prev)
    prev=end
tot=0
for k,d in sorted(agg.items()):
    n=d['n']; t=sum(v for a,v in d.items() if a!='n'); tot+=t
    print(f'{k:12s} n={n:3d} total={t:8d} avg={t//n:6d}', {a:round(b/n) for a,b in d.items() if a!='n'})
print('sum',tot)
