"""Round-2 study: an N-independent rule for the 8-GPU tail -- a Hamiltonian that crosses T liouville
evaluations at a Lie-step boundary is ALWAYS continued on a cluster of C CTAs (deterministic, so maps stay identical at
every N) -- simulated with the measured work map (profiles/work_map_r01_expm_onestep_256.npz), 65 us per liouville per
SM, cluster speed-ups s measured (1.44 / 1.87) and hypothetical.  Result: 0.64-0.77 at 8 GPUs with the measured s (0.72
today) and -7..-30 % at one GPU: rejected."""
import heapq, numpy as np, sys
import os
d=np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'profiles', 'work_map_r01_expm_onestep_256.npz'))
w=d['work'].astype(np.int64)
US=65e-6; OVER=30
def sim(ws, T, C, s, nsm=148, us=US):
    ncl=nsm//C; n=len(ws)
    state={'q':0}
    hq=[]; ev=[]; idle=set(); gathering={}; pend={}
    seq=[0]
    def push(t,c):
        seq[0]+=1; heapq.heappush(ev,(t,seq[0],c))
    for c in range(ncl*C): push(0.0,c)
    end=0.0
    def handle(t,c):
        cl=c//C
        if c in pend:
            hq.append(pend.pop(c))
        if cl in gathering:
            gathering[cl]+=1
        elif hq:
            gathering[cl]=1
            # wake idle peers
            for cc in range(cl*C,(cl+1)*C):
                if cc in idle and cc!=c:
                    idle.discard(cc); gathering[cl]+=1
        else:
            if state['q']<n:
                x=ws[state['q']]; state['q']+=1
                run=min(x,T)
                if x>T: pend[c]=x-T
                push(t+(run+OVER)*us,c)
            else:
                idle.add(c)
            return
        if gathering[cl]==C:
            del gathering[cl]
            if hq:
                rem=hq.pop(0)
                for cc in range(cl*C,(cl+1)*C): push(t+rem*us/s,cc)
            else:
                for cc in range(cl*C,(cl+1)*C): push(t,cc)
    while ev:
        t,_,c=heapq.heappop(ev)
        end=max(end,t)
        handle(t,c)
        # heavy items waiting and idle CTAs available -> wake one whole-idle cluster first
        while hq and idle:
            # choose cluster with most idle
            best=None;bn=0
            cnt={}
            for cc in idle: cnt[cc//C]=cnt.get(cc//C,0)+1
            cl=max(cnt,key=lambda k:cnt[k])
            if cl in gathering: 
                # shouldn't be idle in gathering cluster
                break
            cc=[x for x in idle if x//C==cl][0]
            idle.discard(cc)
            before=len(hq)
            handle(t,cc)
            if len(hq)==before and cl in gathering: break
    return end
ideal=lambda G:(w.sum()+OVER*65536)*US/148/G
if __name__=='__main__':
    for G in (1,8):
        base=max(sim(w[r::G],10**9,1,1.0) for r in range(G))
        print('G',G,'baseline ticket %.3f eff %.3f'%(base,ideal(G)/base))
        for C,s in ((2,1.44),(4,1.87),(2,1.7),(4,2.6),(4,3.0)):
            for T in (1000,2000,4000,8000,16000):
                t=max(sim(w[r::G],T,C,s) for r in range(G))
                print('  C',C,'s',s,'T',T,'time %.3f eff %.3f'%(t,ideal(G)/t))
