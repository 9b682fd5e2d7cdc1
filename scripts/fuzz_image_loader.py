"""Host-side robustness: mutated PNG / HDR fixtures must be decoded or rejected with an error, never crash (CPU only)."""
import os, sys, random, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import restir_b200 as rb
random.seed(1)
src=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))+'/tests/golden/images'
files=[open(os.path.join(src,f),'rb').read() for f in sorted(os.listdir(src))]
tmp=tempfile.mkdtemp(); p=os.path.join(tmp,'f.bin')
ok=err=0
for it in range(6000):
    b=bytearray(random.choice(files))
    mode=random.random()
    if mode<0.3: b=b[:random.randrange(0,len(b))]
    elif mode<0.8:
        for _ in range(random.randrange(1,6)): b[random.randrange(len(b))]=random.randrange(256)
    else:
        i=random.randrange(len(b)); b[i:i]=bytes(random.randrange(256) for _ in range(random.randrange(1,40)))
    open(p,'wb').write(b)
    try:
        a=rb.load_image(p, bool(it&1)); ok+=1
        assert a.ndim==3 and a.shape[2]==3
    except rb.RestirError: err+=1
print("ok",ok,"rejected",err)
