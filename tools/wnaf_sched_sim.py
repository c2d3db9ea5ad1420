import random, sys
sys.path.insert(0,'oracle')
R_ORDER = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
def wnaf(c, w):
    d=[]
    while c:
        if c&1:
            u=c%(1<<(w+1))
            if u>(1<<w): u-= (1<<(w+1))
            c-=u
        else: u=0
        d.append(u); c>>=1
    return d
def ops(c,w=4):
    d=wnaf(c,w); seq=[]; found=False
    for x in reversed(d):
        if found: seq.append('D')
        if x!=0:
            found=True; seq.append('A')
    return seq
def sim(K, lanes=32, trials=200, cd=7, ca=16):
    tot=0; ideal=0; naive=0
    for t in range(trials):
        seqs=[ops(random.randrange(1,R_ORDER)) for _ in range(lanes)]
        pos=[0]*lanes
        time=0
        ideal+=sum((cd if o=='D' else ca) for s in seqs for o in s)/lanes
        while True:
            live=[l for l in range(lanes) if pos[l]<len(seqs[l])]
            if not live: break
            wa=[l for l in live if seqs[l][pos[l]]=='A']
            wd=[l for l in live if seqs[l][pos[l]]=='D']
            if len(wa)>=K or not wd:
                for l in wa: pos[l]+=1
                time+=ca
            else:
                for l in wd: pos[l]+=1
                time+=cd
        tot+=time
    return tot/trials, ideal/trials
for K in (1,4,8,12,16,20,24,28,32):
    t,i=sim(K)
    print(K, round(t), round(i), round(i/t,3))
print("ratio policy")
def sim2(alpha, lanes=32, trials=200, cd=7, ca=16):
    tot=0; ideal=0
    for t in range(trials):
        seqs=[ops(random.randrange(1,R_ORDER)) for _ in range(lanes)]
        pos=[0]*lanes
        time=0
        ideal+=sum((cd if o=='D' else ca) for s in seqs for o in s)/lanes
        while True:
            live=[l for l in range(lanes) if pos[l]<len(seqs[l])]
            if not live: break
            wa=[l for l in live if seqs[l][pos[l]]=='A']
            wd=[l for l in live if seqs[l][pos[l]]=='D']
            if not wd or len(wa)*alpha>=len(wd):
                for l in wa: pos[l]+=1
                time+=ca
            else:
                for l in wd: pos[l]+=1
                time+=cd
        tot+=time
    return tot/trials, ideal/trials
for a in (0.2,0.3,0.4,0.5,0.7,1.0,1.5):
    t,i=sim2(a); print(a, round(t), round(i), round(i/t,3))
print("k points per lane")
def sim3(kpts, alpha, lanes=32, trials=100, cd=7, ca=16):
    tot=0; ideal=0
    for t in range(trials):
        seqs=[[ops(random.randrange(1,R_ORDER)) for _ in range(kpts)] for _ in range(lanes)]
        pos=[[0]*kpts for _ in range(lanes)]
        time=0
        ideal+=sum((cd if o=='D' else ca) for L in seqs for s in L for o in s)/lanes
        while True:
            na=nd=0; any_live=False
            for l in range(lanes):
                has_a=has_d=False
                for j in range(kpts):
                    if pos[l][j]<len(seqs[l][j]):
                        any_live=True
                        if seqs[l][j][pos[l][j]]=='A': has_a=True
                        else: has_d=True
                na+=has_a; nd+=has_d
            if not any_live: break
            doA = nd==0 or na*alpha>=nd
            for l in range(lanes):
                # pick the point that is furthest behind among those wanting this op
                best=-1; bp=1<<30
                for j in range(kpts):
                    if pos[l][j]<len(seqs[l][j]) and (seqs[l][j][pos[l][j]]=='A')==doA:
                        if pos[l][j]<bp: bp=pos[l][j]; best=j
                if best>=0: pos[l][best]+=1
            time+=ca if doA else cd
        tot+=time
    return tot/trials, ideal/trials
for kp in (2,3):
    for a in (0.5,0.8,1.0,1.3):
        t,i=sim3(kp,a); print(kp, a, round(t), round(i), round(i/t,3))
