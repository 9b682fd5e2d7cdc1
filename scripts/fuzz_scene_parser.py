"""Host-side robustness: mutated scene / OBJ texts must load or be rejected with an error, never crash (CPU only)."""
import os, sys, random, tempfile, shutil
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import restir_b200 as rb
random.seed(2)
g=np.load(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))+'/tests/golden/scene_file_textured.npz')
base=tempfile.mkdtemp()
for n,t in zip(g["file_names"], g["file_texts"]): open(os.path.join(base,str(n)),'w').write(str(t))
for f in os.listdir(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))+'/tests/golden/images'): shutil.copy(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))+'/tests/golden/images/'+f, base)
names=[str(n) for n in g["file_names"]]
ok=err=0
for it in range(1500):
    tmp=tempfile.mkdtemp()
    for f in os.listdir(base): shutil.copy(os.path.join(base,f), tmp)
    victim=random.choice(names)
    b=bytearray(open(os.path.join(tmp,victim),'rb').read())
    mode=random.random()
    if mode<0.25: b=b[:random.randrange(0,len(b))]
    elif mode<0.7:
        for _ in range(random.randrange(1,8)): b[random.randrange(len(b))]=random.choice(b" \n/-0123456789.exyzvfn\t\r#")
    else:
        lines=b.split(b"\n"); 
        for _ in range(random.randrange(1,4)):
            i=random.randrange(len(lines))
            if random.random()<0.5: del lines[i]
            else: lines.insert(i, random.choice(lines))
        b=bytearray(b"\n".join(lines))
    open(os.path.join(tmp,victim),'wb').write(b)
    try:
        sc=rb.Scene.from_file(os.path.join(tmp,"cornell_tex_file.txt")); sc.close(); ok+=1
    except rb.RestirError: err+=1
    shutil.rmtree(tmp)
print("ok",ok,"rejected",err)
