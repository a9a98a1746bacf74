"""Randomised check of the host expander (expand.cpp: AVX-512 line-aligned / windowed / portable paths) against a numpy
restatement of the reference recurrence: random read lengths (empty reads, block-edge lengths, long reads), match and
chain-id densities, segment cuts, result-array skews, all PML widths; CPU only.
    python tools/fuzz_expand.py [seed] [seconds] [big]      COLBWT_NO_AVX512=1 / COLBWT_EXPAND_WINDOWED=1 select the other paths"""
import os, sys, time, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import col_bwt_b200 as cb
from test_compact import build_compact, dense_from_match_fast
BIG = len(sys.argv) > 3
rng=np.random.default_rng(int(sys.argv[1]) if len(sys.argv)>1 else 0)
t0=time.time(); n_cases=0; n_bases=0
while time.time()-t0 < float(sys.argv[2]) if len(sys.argv)>2 else 60:
    width=int(rng.choice([1,2,4]))
    kind=rng.integers(0,4)
    nreads=int(rng.integers(1,400)) * (40 if BIG else 1)
    if kind==0: lens=rng.integers(0, 255 if width==1 else 700, size=nreads)
    elif kind==1: lens=rng.choice([0,1,63,64,65,127,128,129,150,191,192,250], size=nreads)
    elif kind==2: lens=rng.integers(0, 255 if width==1 else 20000, size=max(1,nreads//20))
    else: lens=np.full(nreads, int(rng.integers(1,255)))
    if width==1: lens=np.minimum(lens,255)
    off=np.concatenate([[0],np.cumsum(lens)]).astype(np.uint64); n=int(off[-1])
    if n==0: continue
    pm=rng.choice([0.0,0.3,0.64,0.845,0.99,1.0]); pc=rng.choice([0.0,0.02,0.07,0.5,1.0])
    match=(rng.random(n)<pm).astype(np.uint8)
    cid=np.where(rng.random(n)<pc, rng.integers(1,256,size=n),0).astype(np.uint8)
    nr=len(lens)
    ncuts=int(rng.integers(0,4)); cuts=sorted(set([0,nr]+[int(x) for x in rng.integers(0,nr+1,size=ncuts)]))
    buf=build_compact(match,cid,off,cuts)
    want=dense_from_match_fast(match,off)
    dt={1:np.uint8,2:np.uint16,4:np.uint32}[width]
    sp=int(rng.choice([0,0,0,width,3*width,16,32])); sc=int(rng.choice([0,0,0,1,3,16,32]))
    pml=cb._aligned_empty(n,dt,skew=sp); c=cb._aligned_empty(n,np.uint8,skew=sc); pml[:]=0xAB; c[:]=0xEE
    rc=cb._L.colbwt_compact_expand(buf.ctypes.data, off.ctypes.data, nr, pml.ctypes.data, width, c.ctypes.data)
    assert rc==0
    ok=np.array_equal(pml.astype(np.uint32), want) and np.array_equal(c,cid)
    if not ok:
        print('MISMATCH', width, kind, nr, n, pm, pc, cuts, sp, sc); np.save('/tmp/fail_lens.npy', lens); sys.exit(1)
    c2=cb._aligned_empty(n,np.uint8,skew=sc); c2[:]=0xEE
    rc=cb._L.colbwt_compact_expand(buf.ctypes.data, off.ctypes.data, nr, None, width, c2.ctypes.data)
    assert rc==0 and np.array_equal(c2,cid)
    n_cases+=1; n_bases+=n
print('ok', n_cases, 'cases', n_bases, 'bases')
