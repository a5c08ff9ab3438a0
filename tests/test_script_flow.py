"""GPU integration test: the call sequence of /root/reference/clip_fft.py main()/train(i) (clip_fft.py:92-315), reproduced
line by line against the drop-in module names (`aphantasia`, `clip`, `imageio` from dropin/). /root/reference itself does
not exist on the GPU box, so the script cannot be executed there; this test exercises exactly the entry points, argument
conventions and side effects the unmodified script relies on (per-step empty_cache(), second synth + D2H + JPEG, .pt save)."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_clip_fft_main_loop_against_dropin(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, 'dropin'))
    try:
        for k in [k for k in sys.modules if k == 'aphantasia' or k.startswith('aphantasia.') or k in ('clip', 'imageio', 'lpips')]:
            del sys.modules[k]
        from imageio import imread  # noqa: F401  (clip_fft.py:6)
        import clip
        import lpips  # noqa: F401
        from aphantasia.image import to_valid_rgb, fft_image, dwt_image  # noqa: F401
        from aphantasia.utils import slice_imgs, sim_func, img_list, txt_clean, checkout, old_torch
        from aphantasia import transforms
        from aphantasia.progress_bar import ProgressBar

        torch.manual_seed(0); np.random.seed(0)
        size, samples, steps = [224, 288], 16, 6                     # a.size is [H, W] (clip_fft.py:80)
        shape = [1, 3, *size]
        params, image_f, sz = fft_image(shape, 0.07, 1.5, None)      # :99
        image_f = to_valid_rgb(image_f, colors=1.8)                  # :101
        optimizer = torch.optim.Adam(params, 0.05, betas=(.0, .999)) # :115
        model_clip, _ = clip.load('ViT-B/32', jit=old_torch())       # :119
        modsize = model_clip.visual.input_resolution                 # :121
        assert modsize == 224
        samples = int(samples * 0.95)                                # :169
        trform_f = transforms.transforms_fast
        txt = 'red square|blue sky:0.5'
        txt_enc = []
        for subtxt in txt.split('|'):                                # :145-152
            wt = 1.
            if ':' in subtxt:
                subtxt, wt = subtxt.split(':'); wt = float(wt)
            emb = model_clip.encode_text(clip.tokenize(subtxt).cuda())
            txt_enc.append([emb.detach().clone(), wt])
        tempdir = str(tmp_path / txt_clean(txt)[:40]); os.makedirs(tempdir, exist_ok=True)
        pbar = ProgressBar(steps)
        losses = []
        for i in range(steps):                                       # train(i), :235-306
            loss = 0
            img_out = image_f(None)
            img_sliced = slice_imgs([img_out], samples, modsize, trform_f, 'uniform', 0.4)[0]
            out_enc = model_clip.encode_image(img_sliced)
            for enc, wt in txt_enc:
                loss += -1. * wt * sim_func(enc, out_enc, 'mix')
            del img_out, img_sliced, out_enc; torch.cuda.empty_cache()
            optimizer.zero_grad(); loss.backward(); optimizer.step()
            losses.append(loss.item())
            with torch.no_grad():
                img = image_f(contrast=1.1).cpu().numpy()[0]
            checkout(img, os.path.join(tempdir, '%04d.jpg' % i), verbose=False)
            pbar.upd()
        files = img_list(tempdir)
        assert len(files) == steps
        torch.save(params, os.path.join(tempdir, 'p.pt'))            # :315
        assert torch.load(os.path.join(tempdir, 'p.pt'))[0].shape == (1, 3, 224, 288 // 2 + 1, 2)
        assert all(np.isfinite(losses)) and losses[-1] < losses[0]   # the optimisation makes progress
        # resume from the saved spectrum (clip_fft.py -r file.pt -> resume_fft, image.py:143-145)
        p2, f2, _ = fft_image(shape, 1.0, 1.5, os.path.join(tempdir, 'p.pt'))
        assert torch.allclose(p2[0].detach(), params[0].detach())
    finally:
        sys.path.remove(os.path.join(ROOT, 'dropin'))
