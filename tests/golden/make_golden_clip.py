#!/usr/bin/env python
"""Golden vectors for the encoder of the validation loop: a SMALL configuration of the reference's own
`CLIP` class (`/root/reference/model/base/model.py:255-372`, imported by path - build container only) with seeded random
weights, its `encode_image` / `encode_text` outputs in float32, and the DCHMT head's codes (`model/DCHMT.py:8-26` +
`train/base.py:150-158`) on top of them.  The state dict travels with the outputs (a few hundred KB), so the GPU box can
check `cmh_b200.valid_loop.Clip` - same parameter names, fused attention - against it without the reference.

    python tests/golden/make_golden_clip.py      ->  tests/golden/clip_tiny.npz
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
CFG = dict(embed_dim=32, image_resolution=48, vision_layers=2, vision_width=64, vision_patch_size=16, context_length=16,
           vocab_size=128, transformer_width=64, transformer_heads=1, transformer_layers=2)


def main():
    sys.path.insert(0, REF)
    spec = importlib.util.spec_from_file_location("ref_clip_model", os.path.join(REF, "model", "base", "model.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    torch.manual_seed(20261018)
    clip = mod.CLIP(CFG["embed_dim"], CFG["image_resolution"], CFG["vision_layers"], CFG["vision_width"], CFG["vision_patch_size"],
                    CFG["context_length"], CFG["vocab_size"], CFG["transformer_width"], CFG["transformer_heads"],
                    CFG["transformer_layers"]).float().eval()
    g = torch.Generator().manual_seed(7)
    image = torch.randn(5, 3, CFG["image_resolution"], CFG["image_resolution"], generator=g)
    text = torch.randint(1, CFG["vocab_size"] - 2, (5, 12), generator=g)
    text[:, 0] = CFG["vocab_size"] - 2                       # start-of-text
    for i in range(5):
        text[i, 4 + i] = CFG["vocab_size"] - 1               # end-of-text = the highest id; the rest is padding
        text[i, 5 + i:] = 0
    with torch.no_grad():
        img_feat = clip.encode_image(image)
        txt_feat = clip.encode_text(text)
    # the DCHMT head on the image features: HashLayer.forward (model/DCHMT.py:20-26) + make_hash_code_DCHMT (train/base.py:150-158)
    bits = 16
    fc = torch.nn.Linear(CFG["embed_dim"], 128)
    hash_list = [torch.nn.Linear(128, 2) for _ in range(bits)]
    with torch.no_grad():
        embed = torch.relu(fc(img_feat))
        code = torch.stack([torch.softmax(l(embed), dim=-1) for l in hash_list]).permute(1, 0, 2)
        logits = torch.stack([l(embed) for l in hash_list]).permute(1, 0, 2)
        hash_code = torch.argmax(code, dim=-1)
        hash_code[torch.where(hash_code == 0)] = -1
    out = {f"sd/{k}": v.numpy() for k, v in clip.state_dict().items()}
    out.update(image=image.numpy(), text=text.numpy(), img_feat=img_feat.numpy(), txt_feat=txt_feat.numpy(),
               head_fc_w=fc.weight.detach().numpy(), head_fc_b=fc.bias.detach().numpy(),
               head_w=torch.stack([l.weight.detach() for l in hash_list]).numpy(),
               head_b=torch.stack([l.bias.detach() for l in hash_list]).numpy(),
               head_logits=logits.numpy(), head_code=hash_code.float().numpy(),
               cfg=np.array([CFG[k] for k in ("embed_dim", "image_resolution", "vision_layers", "vision_width", "vision_patch_size",
                                              "context_length", "vocab_size", "transformer_width", "transformer_heads",
                                              "transformer_layers")], dtype=np.int64))
    path = os.path.join(HERE, "clip_tiny.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
