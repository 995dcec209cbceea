import sys, ctypes as C, numpy as np, time
sys.path[:0]=['/root/repo','/root/repo/tests']
import oracle as O, corpus
L=C.CDLL('/tmp/g6_emu.so')
L.g6_emu_decode.argtypes=[C.c_int,C.c_void_p,C.c_uint32,C.c_void_p,C.c_uint64,C.c_int,C.POINTER(C.c_long),C.POINTER(C.c_long)]
L.g6_emu_decode.restype=C.c_long
def emu(codec, comp, cap, T=256):
    a=np.frombuffer(comp,dtype=np.uint8); out=np.zeros(70000,dtype=np.uint8); it=C.c_long(); mv=C.c_long()
    r=L.g6_emu_decode(codec,a.ctypes.data,len(a),out.ctypes.data,cap,T,C.byref(it),C.byref(mv))
    return r,out[:max(r,0)].tobytes(),it.value,mv.value
data=O.synth(64,65536)
for T in (256,512):
  for codec,comp in ((0,O.snappy_raw_compress),(2,O.lz4_block_compress)):
    its=[]
    for i in range(24):
        blk=data[i*65536:(i+1)*65536].tobytes()
        c=comp(blk)
        r,out,it,mv=emu(codec,c,65536,T)
        assert r==65536 and out==blk,(codec,i,r)
        its.append(it)
    print("T",T,"codec",codec,"iterations/block mean",np.mean(its),"warp-iterations",np.mean(its)*T/32)
cases=[d for d in corpus.edge_cases() if 0<len(d)<=65536]
for codec,comp in ((0,O.snappy_raw_compress),(2,O.lz4_block_compress)):
    acc=0
    for d in cases:
        c=comp(d)
        for T in (1,7,64,512):
            r,out,it,mv=emu(codec,c,len(d),T)
            if r==0: continue
            assert r==len(d) and out==d,(codec,len(d),T,r)
            acc+=1
    print("codec",codec,"edge cases accepted runs",acc,"of",len(cases)*4)
print("per block iterations (T=512, snappy):")
its=[]
for i in range(40):
    blk=data[i*65536:(i+1)*65536].tobytes(); c=O.snappy_raw_compress(blk)
    r,out,it,mv=emu(0,c,65536,512); its.append((it,len(c)))
print(its)
