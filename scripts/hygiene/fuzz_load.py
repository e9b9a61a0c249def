import sys, os, tempfile, random; sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
import numpy as np, oracle as O
from annlite_b200.engine import Engine
from annlite_b200._lib import AnnbError
rng=np.random.default_rng(0); X=rng.standard_normal((800,16)).astype(np.float32)
cb=np.stack([X[rng.choice(800,16,replace=False),m*4:(m+1)*4] for m in range(4)]).astype(np.float32)
e=Engine(16,4,16,'euclidean',device=-1); e.init_graph(800,M=8,ef_construction=50)
e.add_items_with_tables(O.encode(X,cb),O.adc_table(X,cb),np.arange(800,dtype=np.uint64),num_threads=1)
d=tempfile.mkdtemp(); p=os.path.join(d,'g.hnsw'); e.save_index(p); good=open(p,'rb').read()
random.seed(1); ok=err=0
for it in range(int(sys.argv[1])):
    b=bytearray(good); mode=it%4
    if mode==0: b=b[:random.randrange(0,len(b))]
    elif mode==1:
        for _ in range(random.randrange(1,8)): b[random.randrange(0,96)]=random.randrange(256)      # header
    elif mode==2:
        for _ in range(random.randrange(1,20)): b[random.randrange(0,len(b))]=random.randrange(256)
    else:
        i=random.randrange(0,len(b)-8); b[i:i+8]=(random.getrandbits(64)).to_bytes(8,'little')
    q=os.path.join(d,'m.hnsw'); open(q,'wb').write(bytes(b))
    e2=Engine(16,4,16,'euclidean',device=-1)
    try:
        e2.load_index(q); ok+=1
        st=e2.get_graph(); e2.save_index(os.path.join(d,'o.hnsw'))
    except AnnbError as ex: err+=1
    except MemoryError: err+=1
    del e2
print('loaded',ok,'rejected',err)
