import sys, torch
sys.path.insert(0, "/root/repo")
from mage_b200 import synthetic as syn
from mage_b200.config import instantiate_from_config
for L in (6, 32):
    params = syn.model_params("caterv2plus", frames_length=L)
    sd = syn.make_mage_state_dict(params)
    model = instantiate_from_config({"target": "modules.mage_model.MAGE", "params": params})
    model.load_state_dict(sd); model = model.to("cuda").eval()
    eng = model.engine()
    ae = syn.PatchLatentAE(**params["first_stage_config"]["params"])
    big = syn.make_batch(params, 64, seed=1234, text_len=20)
    noise = syn.make_noise(64, seed=99)
    z = ae.encode(big["images"][:, 0])
    ref = None
    for B in (1, 2, 4, 16, 64):
        out = eng.generate_continuous(z[:B].cuda(), big["text"][:B].cuda(), big["speed"][:B].cuda(), noise[:B].cuda())
        torch.cuda.synchronize()
        if ref is None:
            ref = out[:1].clone()
        d = (out[:1] - ref).abs()
        print(f"L={L} B={B}: row0 max diff vs B=1: {float(d.max()):.3e}; per slot {[f'{float(x):.1e}' for x in d[0].flatten(1).max(1)[0][:8]]}", flush=True)
