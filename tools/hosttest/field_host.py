import random, subprocess, sys
sys.path.insert(0,'/root/repo/oracle')
from pyref import P, Q
rnd = random.Random(5)
cases=[]
def gen(mod, nbits):
    sp = [0,1,2,mod-1,mod-2,(1<<nbits)%mod, mod>>1, (1<<(nbits-32))-1]
    out=[]
    for _ in range(3000):
        v=[]
        for k in range(4):
            t = rnd.random()
            if t<0.15: v.append(rnd.choice(sp))
            elif t<0.25: v.append(rnd.getrandbits(rnd.choice([1,31,32,33,64,65,200])) % mod)
            else: v.append(rnd.randrange(mod))
        out.append(v)
    return out
lines=[]; exp=[]
for name,mod,nb in (('r',Q,256),('p',P,384)):
    Rinv = pow(1<<nb, -1, mod)
    for a,b,c,d in gen(mod,nb):
        # b, d may be any N-limb value
        if rnd.random()<0.2: b = rnd.getrandbits(nb)
        if rnd.random()<0.1: b = (1<<nb)-1
        w = nb//4
        lines.append("%s %0*x %0*x %0*x %0*x" % (name,w,a,w,b,w,c,w,d))
        exp.append((a*b*Rinv%mod, (a*b+c*d)*Rinv%mod, (a+(b%mod))%mod if b<mod else None, (a-b)%mod if b<mod else None))
out = subprocess.run(['/tmp/tf'], input="\n".join(lines)+"\n", capture_output=True, text=True).stdout.split("\n")
bad=0
for l,e,o in zip(lines,exp,out):
    got=[int(x,16) for x in o.split()]
    for k in range(4):
        if e[k] is not None and got[k]!=e[k] and not (k==1 and int(l.split()[2],16)>=(Q if l[0]=="r" else P)):
            bad+=1
            if bad<5: print("BAD",k,l[:60])
print("cases",len(lines),"bad",bad)
