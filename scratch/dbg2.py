import torch, numpy as np, sys
sys.path.insert(0,'/root/repo')
from tests import parity
shapes = {"blk.attn.weight": (300, 70), "wide.weight": (1, 1, 40000)}
for n in (16, 20):
    ref,res,_=parity.run_both(shapes,n,mask_p=None,svd_energy_threshold=0.9)
    job=res["job"]
    for name,rb in ref["bases"].items():
        dt,p=res["bases"]._index[name]
        S_ref=rb["singular_values"].numpy(); S_new=job._fetch()[dt]["sv"][p][:len(S_ref)]
        print(n,name,"k",res["bases"].meta(name)["k"],rb["k"])
        print("  ref",np.array2string(S_ref/S_ref[0],precision=6))
        print("  err/s1",np.array2string((S_new-S_ref)/S_ref[0],precision=2))
