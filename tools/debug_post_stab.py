import sys, numpy as np
sys.path.insert(0,'/root/repo')
from tests.util import small_pile, params, key_index
from adaptivemerging_b200.system import RigidBodySystem
from adaptivemerging_b200.ctypes_defs import contact_keys
from oracle.oracle import Oracle
blob=small_pile(); p=params(); p.enable_post_stabilization=1
gpu=RigidBodySystem(0).load(blob,p); cpu=Oracle(blob,p); gpu.record_orders(True)
for step in range(24):
    gpu.advanceTime(0.05)
    full,sweep,post=gpu.order(0),gpu.order(1),gpu.order(2)
    cpu.set_next_orders(full=full if len(full) else None, sweep=sweep if len(sweep) else None, post=post if len(post) else None)
    mism=cpu.step(0.05)
    g,o=gpu.bodies(),cpu.bodies()
    err={k:np.abs(g[k]-o[k]).max() for k in ('x','R','v','omega')}
    cg,co=gpu.contacts(),cpu.contacts()
    kg,ko=key_index(cg),key_index(co)
    common=[k for k in kg if k in ko]
    ig=np.array([kg[k][0] for k in common]); io=np.array([ko[k][0] for k in common])
    d={f:(np.abs(cg[f][ig]-co[f][io]).max() if len(common) else 0) for f in ('lambda','violation','prev_violation','point_w','lambda_warm')}
    print(step, 'mism',mism, 'nc',len(cg),len(co), 'orders',len(full),len(sweep),len(post), {k:float('%.2e'%v) for k,v in err.items()}, {k:float('%.2e'%v) for k,v in d.items()}, 'coll', (g['collection']>=0).sum(), (o['collection']>=0).sum(), 'ev', len(gpu.events()), len(cpu.events()))
print('gpu events', gpu.events().tolist()); print('cpu events', cpu.events().tolist())
w=np.abs(g['v']-o['v']).max(axis=1); i=int(w.argmax()); print('worst body', i, 'gpu v', g['v'][i], 'cpu v', o['v'][i], 'coll', g['collection'][i], o['collection'][i], 'sleep', g['sleeping'][i], o['sleeping'][i])
print('gpu coll', g['collection'].tolist()); print('cpu coll', o['collection'].tolist())
