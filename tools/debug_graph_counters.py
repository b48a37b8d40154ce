"""Debug: which device-side counters move when a train step is warmed up, captured and replayed (round 2: num_batches_tracked
came out one higher under Train1Graph than after the same number of eager steps)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from findtextcenternet_b200 import shard, synthetic, train
from findtextcenternet_b200.loss_func import CoVWeightingLoss
from findtextcenternet_b200.models.adamw_schedulefree import AdamWScheduleFree
from findtextcenternet_b200.models.detector import TextDetectorModel

model = TextDetectorModel(pre_weights=False)
model.load_state_dict(synthetic.detector_state_dict(0))
model.set_precision("fp32")
model = model.cuda().train()
model.detector.stochastic_depth_prob = 0.0
opt = AdamWScheduleFree([p for p in model.parameters() if p.requires_grad], lr=1e-4)
opt.train()
cov = CoVWeightingLoss(device="cuda", losses=train.TRAIN1_LOSSES)
batch = synthetic.train1_batch(2, seed=0, size=64, device="cuda")
fmask = model.get_fmask(batch["labelmap"], None)
bn = getattr(getattr(model.detector.backbone.features, "0"), "1")
bn2 = getattr(getattr(getattr(getattr(model.detector.backbone.features, "4"), "3").block, "0"), "1")


def show(tag):
    torch.cuda.synchronize()
    print(tag, "nbt0", int(bn.num_batches_tracked), "nbt_mid", int(bn2.num_batches_tracked), "cov_it", float(cov._it), flush=True)


show("init")
import findtextcenternet_b200.train as T
orig = T.train1_step
calls = [0]
def counted(*a, **k):
    calls[0] += 1
    return orig(*a, **k)
T.train1_step = counted
g = train.Train1Graph(model, opt, cov, 2, "cuda", size=64, warmup_batch=(batch["image"], batch["labelmap"], batch["idmap"], fmask), eager_steps=2)
show(f"after Train1Graph init (train1_step calls: {calls[0]})")
for i in range(3):
    g.step(batch["image"], batch["labelmap"], batch["idmap"], fmask)
    show(f"after replay {i + 1}")
