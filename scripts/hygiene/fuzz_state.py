import os, sys, random; sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
import numpy as np, oracle as O
from annlite_b200.engine import Engine
from annlite_b200._lib import AnnbError
rng=np.random.default_rng(0); X=rng.standard_normal((800,16)).astype(np.float32)
cb=np.stack([X[rng.choice(800,16,replace=False),m*4:(m+1)*4] for m in range(4)]).astype(np.float32)
e=Engine(16,4,16,'euclidean',device=-1); e.init_graph(800,M=8,ef_construction=50)
e.add_items_with_tables(O.encode(X,cb),O.adc_table(X,cb),np.arange(800,dtype=np.uint64),num_threads=1)
st=e.get_graph(); random.seed(2); ok=err=0
for it in range(int(sys.argv[1])):
    s=dict(st)
    for key in ('data_level0','link_lists','element_levels'):
        a=np.array(st[key]).copy()
        if random.random()<0.6 and a.size:
            v=a.view(np.uint8)
            for _ in range(random.randrange(1,6)): v[random.randrange(v.size)]=random.randrange(256)
        s[key]=a
    if random.random()<0.3: s['enterpoint_node']=random.randrange(0,5000)
    if random.random()<0.3: s['max_level']=random.randrange(-2,70)
    if random.random()<0.2: s['cur_element_count']=random.randrange(0,800)
    e2=Engine(16,4,16,'euclidean',device=-1)
    try:
        e2.set_graph(s); ok+=1
        e2.get_graph(); e2.mark_deleted(int(3)) if e2.element_count>3 else None
        # builder on top of an accepted state must not crash either
        e2.resize_index(e2.element_count+5)
        e2.add_items_with_tables(O.encode(X[:3],cb),O.adc_table(X[:3],cb),np.arange(9000,9003,dtype=np.uint64),num_threads=1)
    except AnnbError: err+=1
    del e2
print('accepted',ok,'rejected',err)
