"""Two GPUs in ONE process (the reference API allows it: every op carries its device, pseudo_codec has --gpu-id, BaseOpModule
re-keys the native op when a module moves).  Engine state, pinned buffers, function attributes and SM counts are kept per device
(ADVICE r1): a codec on cuda:1 must produce the same bytes and symbols as on cuda:0, also when the two are used alternately.
Skipped on a single-GPU box."""
import os

import numpy as np
import pytest

from conftest import smooth_images

pytestmark = pytest.mark.gpu


def test_codec_on_two_devices_in_one_process(cuda, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from pseudocylindrical_convolution_b200 import pseudo_codec as pc
    from pseudocylindrical_convolution_b200.PCONV_operator import SphereSlice
    from pseudocylindrical_convolution_b200.random_init import synthesize_checkpoints
    H, W, VD = 256, 512, 56
    p_enc, p_dec, p_ent = synthesize_checkpoints(str(tmp_path), "4_56", VD, 0, seed=0)
    x = smooth_images(2, 3, H, W, seed=2)
    out = {}
    codecs = {}
    for gid in (0, 1):
        dev = "cuda:%d" % gid
        enc = pc.PseudoEncoder(VD, gid).to(dev)
        dec = pc.PseudoDecoder(VD, gid).to(dev)
        pc.load_models(enc, p_enc, p_ent, dev)
        pc.load_models(dec, p_dec, p_ent, dev)
        codecs[gid] = (enc, dec)
    for rnd in range(2):                                   # alternate between the devices
        for gid in (0, 1):
            enc, dec = codecs[gid]
            xt = torch.from_numpy(x).to("cuda:%d" % gid)
            names = [str(tmp_path / ("g%d_r%d_%d.bin" % (gid, rnd, i))) for i in range(2)]
            enc.encode_batch(xt, names)
            rec = dec.decode_batch(names, H, W)
            out[(gid, rnd)] = ([open(n, "rb").read() for n in names], rec.cpu().numpy())
    ref = out[(0, 0)]
    assert len(ref[0][0]) > 500
    for key, (streams, rec) in out.items():
        assert streams == ref[0], key
        assert np.array_equal(rec, ref[1]), key
    # BaseOpModule: a module built for GPU 0 follows .to('cuda:1')
    sl = SphereSlice(16, pad=0, opt=True, device=0)
    t0 = sl(torch.from_numpy(x).to("cuda:0"))
    sl = sl.to("cuda:1")
    assert list(sl.op) == [1] and sl.device_list == [1]
    t1 = sl(torch.from_numpy(x).to("cuda:1"))
    assert torch.equal(t0.cpu(), t1.cpu())
