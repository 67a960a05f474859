import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from util import make_oracle, tiny_weights, Semantics, ulp_stats, Golden
from oracle import vae as ovae
from unimedvl_b200.autoencoder import AutoEncoder
from unimedvl_b200.engine import Engine
dims, sd, vsd = tiny_weights(vae=True)
eng = Engine(dims, max_tokens=512, max_seqs=4, kv_pages=64, enable_vae=True)
eng.load_state_dict(sd); vae = AutoEncoder(eng); vae.load_state_dict(vsd); eng.finalize()
o = make_oracle(Semantics.cuda, vae=True)
torch.manual_seed(0)
x = (torch.rand(1, 3, 32, 48) * 2 - 1).bfloat16()
ref = ovae.encoder(o.vae_sd, x, o.dims.vae, Semantics.cuda)
def enc(tag):
    got = vae.encode_moments(x.cuda()); torch.cuda.synchronize()
    print(tag, round(ulp_stats(got, ref)["rel_l2"], 5), flush=True)
enc("fresh")
enc("again")
z = Golden("t2i").t("vae.decode_in")
img = vae.decode(z.cuda()); torch.cuda.synchronize()
enc("after decode 8x8")
enc("again")
x2 = (torch.rand(1, 3, 64, 80) * 2 - 1).bfloat16()
vae.encode_moments(x2.cuda()); torch.cuda.synchronize()
enc("after encode 64x80")
