"""Seeded random-init checkpoints with the reference's file layout and key sets.

The pretrained weights of the reference are not reachable (Baidu links, demo/pretrained_models.txt), so every configuration
runs with random-init weights shared by both sides.  Transforms: torch default inits (conv / PReLU), GDN and quantiser as
the reference constructs them.  Entropy net: the codec-time modules initialise with torch.rand (EntropyContextNew.py:245-249),
which overflows through 12 positive 5x5 layers; weights are drawn like the training net instead (kaiming-normal over the
causal half of the taps, zero bias, delta-net output bias 2, PReLU slope 0.25; MaskConstrain.py:30-32, model_zoo_v2.py:262)."""
import math
import os

import torch


def init_entropy_net(ent, seed=0):
    g = torch.Generator().manual_seed(seed)
    convs = [m for m in ent.net.modules() if hasattr(m, "weight") and m.weight.dim() == 5]
    for li, m in enumerate(convs):
        nb, nout, channel, k, _ = m.weight.shape
        fan_in = channel * k * k / 2.0                     # about half of the taps are causal
        w = torch.randn(m.weight.shape, generator=g) * math.sqrt(2.0 / fan_in)
        b = torch.zeros(m.bias.shape)
        if li == len(convs) - 1:
            b[1] = 2.0                                     # delta net: softplus-like offset keeps the scales away from 0
        with torch.no_grad():
            m.weight.copy_(w)
            m.bias.copy_(b)
            if m.relu is not None:
                m.relu.fill_(0.25)


def synthesize_checkpoints(model_dir, prex, valid_dim, device_id=0, seed=0):
    """Writes {prex}_encoder.pt, {prex}_decoder.pt, {prex}_ent.pt into model_dir; returns the three paths."""
    from .pseudo_codec import PseudoDecoder, PseudoEncoder
    os.makedirs(model_dir, exist_ok=True)
    torch.manual_seed(seed)
    enc = PseudoEncoder(valid_dim, device_id)
    torch.manual_seed(seed + 1)
    dec = PseudoDecoder(valid_dim, device_id)
    init_entropy_net(enc.ent, seed + 2)
    with torch.no_grad():
        # make the 8 quantiser centres span the sigmoid code range so that all symbols occur
        w = enc.quant.weight.data
        w[:, 0] = 0.06
        w[:, 1:] = math.log(0.125)
        dec.quant.weight.data.copy_(w)
    es, ds = enc.state_dict(), dec.state_dict()
    paths = [os.path.join(model_dir, "%s_%s.pt" % (prex, n)) for n in ("encoder", "decoder", "ent")]
    torch.save({k: v.cpu() for k, v in es.items() if k.startswith(("encoder.", "quant."))}, paths[0])
    torch.save({k: v.cpu() for k, v in ds.items() if k.startswith(("decoder.", "quant."))}, paths[1])
    torch.save({k: v.cpu() for k, v in es.items() if k.startswith("ent.")}, paths[2])
    return paths
