#!/usr/bin/env python3
"""Offline study (r02): how much further an insertion-based optimiser (Bittner et al. 2013: remove a subtree, re-insert it at the
best place found by branch and bound) would lower sum SA(internal) / SA(root) on top of the shipped tree rotations.
usage: reinsert_study.py gpurun_out/bvh_<scene>.npz ...   (dumps from tools/dump_bvh.py)"""
import sys, heapq, numpy as np, time

def area(lo,hi):
    d=np.maximum(hi-lo,0.0); return d[0]*d[1]+d[1]*d[2]+d[2]*d[0]
def load(path):
    d=np.load(path); nodes=d['nodes']; N=len(nodes)//2
    left={};right={};lo={};hi={};parent={}
    nxt=[N]
    def child(r):
        if r['link']>=0: return int(r['link'])
        i=nxt[0]; nxt[0]+=1; lo[i]=r['bmin'].astype(np.float64); hi[i]=r['bmax'].astype(np.float64); return i
    stack=[0]; order=[]
    while stack:
        p=stack.pop(); order.append(p)
        a,b=nodes[2*p],nodes[2*p+1]
        ca=child(a); cb=child(b); left[p]=ca; right[p]=cb; parent[ca]=p; parent[cb]=p
        for c in (ca,cb):
            if c<N: stack.append(c)
    for p in reversed(order):
        lo[p]=np.minimum(lo[left[p]],lo[right[p]]); hi[p]=np.maximum(hi[left[p]],hi[right[p]])
    parent[0]=-1
    return N,left,right,lo,hi,parent,order
def total(internal,lo,hi): return sum(area(lo[p],hi[p]) for p in internal)
def refit_up(p,left,right,lo,hi,parent):
    while p!=-1:
        nlo=np.minimum(lo[left[p]],lo[right[p]]); nhi=np.maximum(hi[left[p]],hi[right[p]])
        if (nlo==lo[p]).all() and (nhi==hi[p]).all(): break
        lo[p]=nlo; hi[p]=nhi; p=parent[p]
def run(path,passes=2):
    N,left,right,lo,hi,parent,order=load(path)
    internal=set(order); root=0
    A0=area(lo[root],hi[root])
    print(path.split('/')[-1],'start',round(total(internal,lo,hi)/A0,3))
    rng=np.random.RandomState(1)
    for ps in range(passes):
        nodes=[n for n in list(lo.keys()) if n!=root and parent.get(n,-1)!=root and parent.get(n,-1)!=-1]
        nodes.sort(key=lambda n:-area(lo[n],hi[n]))
        moved=0
        for S in nodes:
            P=parent[S]
            if P==root or P==-1: continue
            G=parent[P]; sib=right[P] if left[P]==S else left[P]
            # remove S and P: sibling takes P's place
            if left[G]==P: left[G]=sib
            else: right[G]=sib
            parent[sib]=G
            refit_up(G,left,right,lo,hi,parent)
            # find best insertion position for S (branch and bound)
            sa=area(lo[S],hi[S])
            best=(float('inf'),None)
            heap=[(0.0,0,root)]; cnt=1
            while heap:
                induced,_,x=heapq.heappop(heap)
                if induced+sa>=best[0]: break
                direct=area(np.minimum(lo[x],lo[S]),np.maximum(hi[x],hi[S]))
                c=induced+direct
                if c<best[0]: best=(c,x)
                if x in internal:
                    inc=induced+direct-area(lo[x],hi[x])
                    if inc+sa<best[0]:
                        for ch in (left[x],right[x]):
                            heapq.heappush(heap,(inc,cnt,ch)); cnt+=1
            X=best[1]
            # insert: P becomes parent of (X,S) at X's place
            XP=parent[X]
            if X==root:
                # keep root id stable: not handled; put back at sibling
                X=sib; XP=parent[X]
            if left[XP]==X: left[XP]=P
            else: right[XP]=P
            parent[P]=XP; left[P]=X; right[P]=S; parent[X]=P; parent[S]=P
            lo[P]=np.minimum(lo[X],lo[S]); hi[P]=np.maximum(hi[X],hi[S])
            refit_up(XP,left,right,lo,hi,parent)
            if X!=sib: moved+=1
        print('  pass',ps,'moved',moved,'cost',round(total(internal,lo,hi)/A0,3))
for p in sys.argv[1:]: 
    t=time.time(); run(p); print('  time',round(time.time()-t,1))
