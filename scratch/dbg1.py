import torch, numpy as np, sys
sys.path.insert(0,'/root/repo')
from tests import parity
from svd_quantization_task_merging_b200 import synth
ref,res,_=parity.run_both(synth.toy_shapes(),4,svd_energy_threshold=0.8,svd_fp16=False)
job=res["job"]
f=job._fetch()[torch.float32]
print("info",f["info"]); print("scal",f["scal"]); print("cbar",f["cbar"]); print("sv",f["sv"])
for name in ref["merged_state_dict"]:
    a=ref["merged_state_dict"][name]; b=res["merged_state_dict"][name].cpu()
    print(name,"ref nan",torch.isnan(a).sum().item(),"new nan",torch.isnan(b).sum().item(), "ref k",ref["bases"][name]["k"], ref["bases"][name]["singular_values"])
    for t in job.tasks:
        pl=ref["compressed"][name][t]["c_low_quant"]["payloads"]
        print("  ",t,[ (p["quantized"].tolist(),p["scale"].item(),p["zero_point"].item()) for p in pl], ref["compressed"][name][t]["c_low_fp32"])
        break
print("codes",f["codes"][0,0], f["qscale"][0,0], f["qzp"][0,0], f["coef"][0,0])
