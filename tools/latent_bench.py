"""BASELINE config 5 (retrieval half): encoder-only latent extraction throughput at the headline shape and top-k
cosine search over 10k x 10k latents of dimension 1200.  Run on the GPU box."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sivae_b200  # noqa: E402

dev = "cuda"
torch.manual_seed(77)
net = sivae_b200.SoftIntroVAE(64, [[64, 1, 2], [128, 1, 2], [256, 2, 2]])
net.apply(sivae_b200.init_weights_he)
net.to(dev)
x = torch.rand(64, 1, 80, 96, 80, device=dev)
sivae_b200.extract_latents(net, x[:8], batch_size=8)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
lat = sivae_b200.extract_latents(net, x, mode="mu", batch_size=8)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"encoder-only latent extraction: {x.shape[0]} volumes 80x96x80 in {ms:.1f} ms = {x.shape[0] / ms * 1e3:.1f} volumes/s "
      f"(latent {tuple(lat.shape)})")
db = torch.randn(10000, 1200, device=dev)
sivae_b200.topk_similar(db[:64], db, k=10)
torch.cuda.synchronize()
for metric in ("cosine", "l2"):
    e0.record()
    sc, ix = sivae_b200.topk_similar(db, db, k=10, metric=metric)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    ok = bool((ix[:, 0].cpu() == torch.arange(10000, dtype=torch.int32)).all())
    print(f"top-10 {metric} search, 10000 queries x 10000 database x 1200 dims: {ms:.1f} ms "
          f"({2 * 1e4 * 1e4 * 1200 / ms / 1e9:.1f} TFLOP/s fp32), self-match first: {ok}")
