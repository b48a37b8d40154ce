"""Two bf16 Transformer.forward calls (cfg4, batch 256) for profiler runs.  Usage: python tools/run_transformer.py [cfg4|default]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from findtextcenternet_b200 import synthetic
import findtextcenternet_b200.models.transformer as T
name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
dims = dict(cfg4=dict(embed_dim=512, head_num=16, enc_block_num=16, dec_block_num=16, max_enc_seq_len=100, max_dec_seq_len=100), default=dict())[name]
B = 256 if name == "cfg4" else 64
cfg = T.ModelDimensions(**dims)
m = T.Transformer(**cfg.__dict__)
m.load_state_dict(synthetic.transformer_state_dict(0, **dims))
m = m.cuda().eval().set_precision("bf16")
m.weights_frozen = True
enc, dec, _ = synthetic.transformer_inputs(B, cfg.max_enc_seq_len, cfg.max_dec_seq_len, 0)
enc, dec = enc.cuda(), dec.cuda()
with torch.no_grad():
    for _ in range(2):
        m(enc, dec)
torch.cuda.synchronize()
print("done")
