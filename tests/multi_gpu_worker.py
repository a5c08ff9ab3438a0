"""Worker of tests/test_multi_gpu.py (not a pytest module): one optimisation step of the hot path through the reference-facing
entry points; under torchrun the crops are sharded over the ranks and dRGB is all-reduced (aphantasia_b200/_dist.py). Writes
{loss_local, grad, params_after_adam} of rank 0 to argv[1].

    python tests/multi_gpu_worker.py out.pt H W S patch [sim]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tests/multi_gpu_worker.py out.pt H W S patch [sim]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    out, H, W, S, patch = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
    sim = sys.argv[6] if len(sys.argv) > 6 else 'mix'
    os.environ['APH_SYNC_SEED'] = '0'                    # explicit seeds below: every rank replays the same stream
    from aphantasia_b200 import _dist, transforms
    from aphantasia_b200.clip import CLIP, synthetic_visual_state_dict
    from aphantasia_b200.image import fft_image, to_valid_rgb
    from aphantasia_b200.utils import sim_func, slice_imgs
    st = _dist.init()
    torch.manual_seed(0); np.random.seed(0)
    params, image_f, _ = fft_image([1, 3, H, W], 0.07, 1.5, None)
    rgb_f = to_valid_rgb(image_f, colors=1.8)
    model = CLIP('ViT-B/%d' % patch, synthetic_visual_state_dict(patch=patch, seed=0), True)
    g = torch.Generator().manual_seed(1234)
    txt = torch.randn(1, 512, generator=g); txt = (10. * txt / txt.norm()).cuda()
    opt = torch.optim.Adam(params, 0.05, betas=(.0, .999))
    torch.manual_seed(1); np.random.seed(1)
    losses = []
    for step in range(2):
        crops = slice_imgs([rgb_f()], S, 224, transforms.transforms_fast, 'uniform', 0.4)[0]
        emb = model.encode_image(crops)
        loss = -1. * sim_func(txt, emb, sim)
        opt.zero_grad(); loss.backward()
        if step == 0:
            grad0 = params[0].grad.detach().clone()
        opt.step()
        # the local loss is a mean over the local shard: the N-rank weighted mean must equal the 1-rank loss
        w = torch.tensor([loss.item() * crops.shape[0] / S], device='cuda', dtype=torch.float64)
        _dist.all_reduce_sum_(w)
        losses.append(float(w.item()))
    torch.cuda.synchronize()
    assert not _dist.collective_error(), 'a rank barrier of the symmetric all-reduce timed out'
    if st['rank'] == 0:
        torch.save({'collective': _dist.collective_mode(), 'loss': losses, 'grad': grad0.cpu(), 'params': params[0].detach().cpu(), 'world': st['world'], 'local_crops': int(crops.shape[0])}, out)
    if st['world'] > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == '__main__':
    main()
