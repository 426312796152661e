"""Codec drivers and command line (reference: pseudo_codec.py).

`python pseudo_codec.py --enc/--dec/--test ...` keeps the reference's flags and file conventions.  Classes
EntropyConvDBT, EntropyResidualBlockDBT, EntEncoder, EntDecoder, PseudoEncoder, PseudoDecoder keep their names,
constructor arguments and state_dict keys.  Differences: sizes are not hard-wired to 512x1024 (the code size is
derived from the input / passed to the decoder), only the rows the coder needs cross PCIe each wavefront step,
and --test evaluates PSNR / SSIM on the reference's 14 viewports with this package's own projector and SSIM kernels.
"""
import argparse
import math
import os
from collections import OrderedDict

import numpy as np
import torch
from torch import nn

from . import coder
from .PCONV_operator import (DExtract2, DExtract2Batch, DInput2, Dtow, EntropyAdd, EntropyBatchGmmTable, EntropyContextNew,
                             EntropyConv2Batch, EntropyCtxPadRun2, Extract, PseudoContextV2, PseudoDQUANT, PseudoFillV2,
                             PseudoQUANTV2, SphereSlice, SphereUslice)
from .model_zoo_v2 import ClipData, DecoderV2, EncoderV2

psnr_f = lambda xa: 10 * math.log10(1. / xa)

model_ssim_list = ['1_56', '2_56', '3_56', '4_56', '5_112', '6_112', '7_112', '8_192', '9_192']
ssim_channel_list = [56, 56, 56, 56, 112, 112, 112, 192, 192]
model_mse_list = ['1_56', '2_56', '3_56', '4_112', '5_112', '6_112', '7_112', '8_192', '9_192', '10_192']
mse_channel_list = [56, 56, 56, 112, 112, 112, 112, 192, 192, 192]
mse_model_dir = './demo/mse'
ssim_model_dir = './demo/ssim'


class EntropyConvDBT(nn.Module):
    """causal halo refresh + masked 5x5 conv of the three nets at the wavefront cells (reference :27-38)"""

    def __init__(self, batch, ngroups, cin, cout, hidden, npart, out_layer: bool, ctx: EntropyContextNew, device_id, act=True):
        super().__init__()
        pad_out = 0 if out_layer else 2
        self.pad = EntropyCtxPadRun2(2, npart, ngroups, ctx, not hidden, device=device_id)
        self.conv = EntropyConv2Batch(npart, ngroups, cin, cout, 5, ctx, 2, pad_out, batch=batch, hidden=hidden, act=act,
                                      device=device_id)

    def forward(self, x):
        return self.conv(self.pad(x))


class EntropyResidualBlockDBT(nn.Module):

    def __init__(self, batch, ngroups, cpn, npart, ctx: EntropyContextNew, device_id=0):
        super().__init__()
        self.conv1 = EntropyConvDBT(batch, ngroups, cpn, cpn, True, npart, False, ctx, device_id, True)
        self.conv2 = EntropyConvDBT(batch, ngroups, cpn, cpn, True, npart, False, ctx, device_id, True)
        self.add = EntropyAdd(npart, cpn * ngroups, ngroups, 2, ctx, device=device_id)

    def forward(self, x):
        return self.add(self.conv2(self.conv1(x)), x)


_STEPPED = (EntropyConv2Batch, EntropyCtxPadRun2, EntropyAdd, DInput2, DExtract2, DExtract2Batch)


@torch.no_grad()
def restart_entropy_network(m):
    if isinstance(m, _STEPPED):
        m.restart()


def _context_net(ngroup, npart, ctx, gid):
    return nn.Sequential(
        EntropyConvDBT(3, ngroup, 1, 3, False, npart, False, ctx, gid, True),
        *[EntropyResidualBlockDBT(3, ngroup, 3, npart, ctx, gid) for _ in range(5)],
        EntropyConvDBT(3, ngroup, 3, 3, True, npart, True, ctx, gid, False))


class _EntBase(nn.Module):
    def __init__(self, ngroup, npart, opt_f, bin_num, gid):
        super().__init__()
        self.cuda = 'cuda:{}'.format(gid)
        self.ctx2 = EntropyContextNew(npart, opt=opt_f, device=gid)
        self.ipt = DInput2(ngroup, npart, self.ctx2, 2, -3.5, 3, device=gid)
        self.npart, self.ngroup = npart, ngroup
        self.fill = PseudoFillV2(0, npart, self.ctx2, 0, device=gid)
        self.mcoder = None
        self.bias = (bin_num - 1) / 2.
        self.net = _context_net(ngroup, npart, self.ctx2, gid)
        self.ext = DExtract2Batch(npart, ngroup, self.ctx2, device=gid)
        self.gmm = EntropyBatchGmmTable(bin_num, self.bias, 3, 65536, device=gid)
        self.net = self.net.to(self.cuda)

    def engine(self):
        if getattr(self, "_engine", None) is None:
            from .wave_engine import WaveEngine
            self._engine = WaveEngine(self)
        return self._engine

    def start(self, code_name='./tmp/data'):
        self.apply(restart_entropy_network)
        self.mcoder = coder.coder(code_name)

    def _tables(self, prev):
        """one wavefront step on the GPU: returns (float CDF table tensor, symbol count of this step)"""
        b = self.ipt(prev)
        y = self.net(b)
        z, le = self.ext(y)
        vec = self.gmm(z, le)
        return b, vec, int(le[0].item())


class EntEncoder(_EntBase):
    """Entropy-codes the symbol tensor (npart, ngroup, h, w) with the wavefront context model (reference :68-114)."""

    def __init__(self, ngroup, npart=16, opt_f=True, bin_num=8, gid=0):
        super().__init__(ngroup, npart, opt_f, bin_num, gid)
        self.ext_label = DExtract2(npart, ngroup, True, self.ctx2, device=gid)

    def forward(self, data):
        from . import config
        with torch.no_grad():
            data = self.fill(data)
            h, w = data.shape[2:]
            self.ctx2.setup_context(w)
            self.mcoder.start_encoder()
            if config.WAVE_IMPL == 1:           # native wavefront engine: same kernels and CDFs, one native call
                self.engine().encode(data, self.mcoder)
                self.mcoder.end_encoder()
                return
            h_full = h * self.npart
            label = torch.zeros((1, 1, h_full, w), dtype=torch.float32).to(self.cuda)
            for _ in range(h_full + w + self.ngroup - 2):
                _, vec, ln = self._tables(label)
                label, _ = self.ext_label(data)
                if ln > 0:      # only the ln live rows cross PCIe (the reference copies both whole buffers, :112)
                    pred = vec[:ln].to(torch.int32).to('cpu')
                    tlabel = label.view(-1)[:ln].to(torch.int32).to('cpu')
                    self.mcoder.encodes(pred, 8, tlabel, ln)
            self.mcoder.end_encoder()


    def encode_batch(self, data, code_names):
        """data (nimg*npart, ngroup, h, w): the nimg images go through the wavefront together (native engine), each into
        its own bitstream file - byte-identical to coding them one by one."""
        with torch.no_grad():
            self.apply(restart_entropy_network)
            data = self.fill(data)
            self.ctx2.setup_context(data.shape[3])
            coders = [coder.coder(n) for n in code_names]
            for c in coders:
                c.start_encoder()
            self.engine().encode(data, coders)
            for c in coders:
                c.end_encoder()


class EntDecoder(_EntBase):
    """Inverse of EntEncoder: decodes (npart, ngroup, h, w) symbols from the bitstream (reference :117-160)."""

    def forward(self, h, w):
        from . import config
        with torch.no_grad():
            self.ctx2.setup_context(w)
            self.mcoder.start_decoder()
            if config.WAVE_IMPL == 1:
                code = self.engine().decode(h, w, self.mcoder, torch.device(self.cuda))
                return self.fill(code)
            h_full = h * self.npart
            pout = torch.zeros((1, 1, h_full, w), dtype=torch.float32).to(self.cuda)
            for _ in range(h_full + w + self.ngroup - 2):
                b, vec, ln = self._tables(pout)
                if ln > 0:
                    pred = vec[:ln].to(torch.int32).to('cpu')
                    sym = self.mcoder.decodes(pred, 8, ln)
                    pout.view(-1)[:ln].copy_(sym[:ln], non_blocking=False)
            # the last step's symbols never pass through DInput2; b holds every earlier plane
            b = self.ipt(pout)
            code = (b[:self.npart, :, 2:-2, 2:-2] + self.bias).contiguous()
            return self.fill(code)


    def decode_batch(self, h, w, code_names):
        """inverse of EntEncoder.encode_batch: returns (nimg*npart, ngroup, h, w)"""
        with torch.no_grad():
            self.apply(restart_entropy_network)
            self.ctx2.setup_context(w)
            coders = [coder.coder(n) for n in code_names]
            for c in coders:
                c.start_decoder()
            return self.fill(self.engine().decode(h, w, coders, torch.device(self.cuda)))


class PseudoEncoder(nn.Module):

    def __init__(self, valid_dim, device_id):
        super().__init__()
        npart, opt, channels, code_channels = 16, True, 192, 192
        quant_levels = 8
        self.slice = SphereSlice(npart, pad=0, opt=opt, device=device_id)
        self.ctx = PseudoContextV2(npart, opt, device=device_id)
        self.encoder = EncoderV2(channels, code_channels, npart, self.ctx, device_id).to('cuda:{}'.format(device_id))
        self.quant = PseudoQUANTV2(code_channels, 8, npart, self.ctx, device_id=device_id, ntop=2)
        self.ext = Extract(valid_dim)
        self.mean_val = (quant_levels - 1) / 2.
        self.dtw = Dtow(2, True, device_id)
        self.ent = EntEncoder(valid_dim // 4, npart, opt, quant_levels, gid=device_id)

    def symbols(self, x):
        """ERP image (1,3,H,W) -> symbol tensor (npart, valid_dim/4, H/128, W/8) fed to the entropy coder."""
        with torch.no_grad():
            return self.dtw(self.ext(self.code(x)[1]))

    def latent(self, x):
        """ERP image -> analysis-transform output (npart, 192, H/256, W/16) in [0,1], before quantisation.  The transform
        runs channels-last on the tcgen05 kernels (config.CONV_IMPL == 0, default) or as the NCHW fp32 exact-order path
        (CONV_IMPL == 1)."""
        from . import config
        with torch.no_grad():
            if config.CONV_IMPL == 1:
                return self.encoder(self.slice(x))
            from .transforms_nhwc import encoder_forward
            return encoder_forward(self.encoder, x.contiguous(), self.slice.op[x.device.index])

    def code(self, x):
        """ERP image -> (dequantised code, symbols)"""
        with torch.no_grad(), torch.cuda.device(x.device):
            return self.quant(self.latent(x))

    def forward(self, x, code_name):
        with torch.no_grad():
            hcode_i = self.symbols(x)
            self.ent.start(code_name)
            self.ent(hcode_i)

    def encode_batch(self, x, code_names):
        """x (N, 3, H, W) -> N bitstream files; transforms and wavefront run batched on the device."""
        with torch.no_grad():
            assert x.shape[0] == len(code_names)
            self.ent.encode_batch(self.symbols(x), code_names)


def _encode_images(self, images_u8, code_names):
    """Host-facing batch entry point: images_u8 is a uint8 (N, H, W, 3) tensor in (ideally pinned) HOST memory, as cv2.imread
    delivers it; the copy to the device, the /255 conversion (img2tensor, reference :215-217), the transforms, the wavefront and
    the host range coder all run inside this call; one bitstream file per image."""
    dev = next(self.encoder.parameters()).device
    with torch.no_grad():
        x = images_u8.to(dev, non_blocking=True).permute(0, 3, 1, 2).to(torch.float32).div_(255.).contiguous()
        self.encode_batch(x, code_names)


PseudoEncoder.encode_images = _encode_images


class PseudoDecoder(nn.Module):

    def __init__(self, valid_dim, device_id):
        super().__init__()
        self.npart, opt, self.channels, self.code_channels = 16, True, 192, 192
        quant_levels = 8
        self.valid_dim = valid_dim
        self.uslice = SphereUslice(self.npart, pad=0, opt=opt, device=device_id)
        self.ctx = PseudoContextV2(self.npart, opt, device=device_id)
        self.decoder = DecoderV2(self.channels, self.code_channels, self.npart, self.ctx, device_id).to('cuda:{}'.format(device_id))
        self.clip = ClipData()
        self.quant = PseudoDQUANT(self.code_channels, 8, self.npart, self.ctx, device_id=device_id)
        self.wtd = Dtow(2, False, device_id)
        self.ent = EntDecoder(self.valid_dim // 4, self.npart, opt, quant_levels, gid=device_id)

    def reconstruct(self, hcode_i):
        with torch.no_grad(), torch.cuda.device(hcode_i.device):
            code_i = self.wtd(hcode_i)
            code_ext = self.quant(code_i)
            n, _, h, w = code_ext.shape
            code_f = torch.zeros((n, self.code_channels, h, w)).type_as(code_ext)
            code_f[:, :self.valid_dim] = code_ext
            from . import config
            if config.CONV_IMPL == 1 or not self.decoder.widths_double(code_f, h, w):
                tx = self.uslice(self.decoder(code_f.contiguous()))
            else:
                from .transforms_nhwc import decoder_forward
                tx = decoder_forward(self.decoder, code_f.contiguous(), self.uslice.op[code_f.device.index])
            return self.clip(tx)

    def forward(self, code_name, height=512, width=1024):
        """The bitstream has no header: the image size is a decoder argument, 512x1024 by default
        like the reference (pseudo_codec.py:206)."""
        with torch.no_grad():
            self.ent.start(code_name)
            hcode_i = self.ent(height // 128, width // 8)
            return self.reconstruct(hcode_i)


def _decode_batch(self, code_names, height=512, width=1024):
    """N bitstream files -> (N, 3, H, W) reconstructions (batched wavefront + batched synthesis transform)."""
    with torch.no_grad():
        return self.reconstruct(self.ent.decode_batch(height // 128, width // 8, code_names))


PseudoDecoder.decode_batch = _decode_batch


def _decode_images(self, code_names, out_u8=None, height=512, width=1024):
    """Inverse of PseudoEncoder.encode_images: N bitstream files -> uint8 (N, H, W, 3) images in host memory (written into
    out_u8 when given, e.g. a pinned buffer).  tensor2img (reference :219-221) with the value range made explicit: the leaky
    ClipData output is clamped to [0, 255] before the truncating conversion."""
    with torch.no_grad():
        rec = self.decode_batch(code_names, height, width)
        img = rec.mul(255.).clamp_(0., 255.).to(torch.uint8).permute(0, 2, 3, 1).contiguous()
        if out_u8 is None:
            return img.cpu()
        out_u8.copy_(img, non_blocking=True)
        torch.cuda.current_stream(img.device).synchronize()
        return out_u8


PseudoDecoder.decode_images = _decode_images


def img2tensor(img, device):
    ts = torch.from_numpy(img.transpose(2, 0, 1).astype(np.float32)) / 255.
    return torch.unsqueeze(ts, 0).to(device).contiguous()


def tensor2img(data):
    img = (data[0] * 255.).to('cpu').detach().numpy().transpose(1, 2, 0)
    return img.astype(np.uint8)


def load_models(model: nn.Module, p1, p2, device):
    d2 = torch.load(p1, map_location=device)
    d1 = torch.load(p2, map_location=device)
    model.load_state_dict(OrderedDict(**d2, **d1))


def check_img(img, height=512, width=1024):
    """The reference resizes every input to 1024 x 512 (pseudo_codec.py:229-234); --height / --width generalise that (f3):
    the coded size must be a multiple of 256 x 128 (16 bands x 16-fold down-sampling; 8 code columns per context cell)."""
    import cv2
    h, w = img.shape[:2]
    if not (h == height and w == width):
        return cv2.resize(img, (width, height), interpolation=cv2.INTER_CUBIC)
    return img


def _check_size(height, width):
    assert height % 256 == 0 and width % 128 == 0 and height > 0 and width > 0, \
        'image size must be a multiple of 256 rows x 128 columns (got {}x{})'.format(height, width)


def _select(model_idx, mse):
    prex = model_mse_list[model_idx] if mse else model_ssim_list[model_idx]
    vd = mse_channel_list[model_idx] if mse else ssim_channel_list[model_idx]
    return prex, vd, (mse_model_dir if mse else ssim_model_dir)


def encoding(img_list, out_list, model_idx=0, mse=True, device_id=0, height=512, width=1024):
    import cv2
    _check_size(height, width)
    prex, vd, model_dir = _select(model_idx, mse)
    cuda = 'cuda:{}'.format(device_id)
    t1 = PseudoEncoder(vd, device_id=device_id).to(cuda)
    load_models(t1, '{}/{}_encoder.pt'.format(model_dir, prex), '{}/{}_ent.pt'.format(model_dir, prex), cuda)
    for fn, fo in zip(img_list, out_list):
        data = img2tensor(check_img(cv2.imread(fn), height, width), cuda)
        t1(data, fo)
        print('Encoding {}, bitrate: {:.3f}bpp'.format(fn, os.path.getsize(fo) * 8 / float(width) / float(height)))


def decoding(code_list, decoded_img_list, model_idx=0, mse=True, device_id=0, height=512, width=1024):
    import cv2
    _check_size(height, width)
    prex, vd, model_dir = _select(model_idx, mse)
    cuda = 'cuda:{}'.format(device_id)
    t1 = PseudoDecoder(vd, device_id=device_id).to(cuda)
    load_models(t1, '{}/{}_decoder.pt'.format(model_dir, prex), '{}/{}_ent.pt'.format(model_dir, prex), cuda)
    for fc, fo in zip(code_list, decoded_img_list):
        cv2.imwrite(fo, tensor2img(t1(fc, height, width)))
        print('Decoding {}, output to {}'.format(fc, fo))


def decoding_and_test(code_list, img_list, model_idx=0, mse=True, device_id=0, height=512, width=1024):
    """Decode and report bitrate, viewport PSNR and viewport SSIM like the reference (pseudo_codec.py:262-290): source and
    reconstruction are sampled on 14 rectilinear viewports of 171 x 256 pixels (MultiProject, fov 0.5 pi), PSNR from the
    mean squared difference, SSIM with an 11-tap Gaussian window."""
    import cv2
    from .PCONV_operator import MultiProject, SSIM
    from .PCONV_operator.pytorch_ssim import mean_squared_difference
    _check_size(height, width)
    prex, vd, model_dir = _select(model_idx, mse)
    cuda = 'cuda:{}'.format(device_id)
    t1 = PseudoDecoder(vd, device_id=device_id).to(cuda)
    load_models(t1, '{}/{}_decoder.pt'.format(model_dir, prex), '{}/{}_ent.pt'.format(model_dir, prex), cuda)
    pr1 = MultiProject(171, int(171 * 1.5), 0.5, False, device_id).to(cuda)
    pr2 = MultiProject(171, int(171 * 1.5), 0.5, False, device_id).to(cuda)
    sim_func = SSIM(11, 3).to(cuda)
    rt_list, pr_list, ssim_list = [], [], []
    for fc, fn in zip(code_list, img_list):
        rdata = t1(fc, height, width)
        data = img2tensor(check_img(cv2.imread(fn), height, width), cuda)
        x = pr1(data)
        y = pr2(rdata)
        pr = psnr_f(mean_squared_difference(x, y).item())
        vssim = sim_func(x, y).item()
        rt = os.path.getsize(fc) * 8 / float(width) / float(height)
        rt_list.append(rt)
        pr_list.append(pr)
        ssim_list.append(vssim)
        print('Decoding {}, compare it to {} \n Bitrate:{:.3f}bpp, PSNR:{:.2f}dB, SSIM:{:.4f}'.format(fc, fn, rt, pr, vssim))
    print('-' * 53 + '\nAverage Performance\n' + '-' * 53)
    rt, pr, vssim = float(np.mean(rt_list)), float(np.mean(pr_list)), float(np.mean(ssim_list))
    print('Bitrate:{:.3f}bpp, PSNR:{:.2f}dB, SSIM:{:.4f}'.format(rt, pr, vssim))
    return rt, pr, vssim


def read_list(fname):
    with open(fname) as f:
        return [line.rstrip('\n') for line in f.readlines()]


def check_models():
    assert os.path.exists('{}/{}_encoder.pt'.format(mse_model_dir, model_mse_list[0])), \
        'Please make sure the pretrained models for VMSE exists in the mse_model_dir'
    assert os.path.exists('{}/{}_encoder.pt'.format(ssim_model_dir, model_ssim_list[0])), \
        'Please make sure the pretrained models for VSSIM exists in the ssim_model_dir'


def main(argv=None):
    parser = argparse.ArgumentParser(description='Pseudo Convolution for 360 Image Compression')
    parser.add_argument('--img-list', nargs='*', help='The image list contains the input images for encoding and testing')
    parser.add_argument('--code-list', nargs='*', help='The code file list for codes')
    parser.add_argument('--out-list', nargs='*', help='The out list for saving decoded images.')
    parser.add_argument('--img-file', help='The file contains the input images for encoding and testing')
    parser.add_argument('--code-file', help='The file contains the list for codes')
    parser.add_argument('--out-file', help='The file  contains the names of decoded images.')
    parser.add_argument('--model-idx', type=int, default=0, help='Model index (0-9) for VMSE, (0-8) for VSSIM')
    parser.add_argument('--enc', action='store_true', default=False, help='Encoding flag, set for encoding phase.')
    parser.add_argument('--dec', action='store_true', default=False, help='Decoding flag, set for decoding phase.')
    parser.add_argument('--test', action='store_true', default=False, help='Testing flag, set for decoding and evalating the performance.')
    parser.add_argument('--ssim', action='store_true', default=False, help='Default with models optimized for VMSE, '
                        'set this flag for choosing the models optimized for VSSIM')
    parser.add_argument('--gpu-id', type=int, default=0, help='The graphic card id for encoding and decoding.')
    parser.add_argument('--height', type=int, default=512, help='Coded image height (multiple of 256); inputs are resized to it. '
                        'The bitstream has no header: pass the same size when decoding. Default 512 like the reference.')
    parser.add_argument('--width', type=int, default=1024, help='Coded image width (multiple of 128). Default 1024 like the reference.')
    args = parser.parse_args(argv)
    check_models()
    midx = args.model_idx
    if args.ssim:
        assert 0 <= midx < 9, '(0-8) for VSSIM'
    else:
        assert 0 <= midx < 10, '(0-9) for VMSE'
    assert args.enc or args.dec or args.test, 'Should set one flag, (--enc) for encoding, (--dec) for decoding, (--test) for testing.'
    pick = lambda lst, fil: lst if lst is not None else (read_list(fil) if fil is not None else None)
    img_list, code_list, out_list = pick(args.img_list, args.img_file), pick(args.code_list, args.code_file), pick(args.out_list, args.out_file)
    if args.enc:
        assert img_list is not None, 'No input images for encoding'
        assert code_list is not None, 'No code files for saving the codes'
        assert len(img_list) == len(code_list), 'The number of images and codes should be the same'
        encoding(img_list, code_list, midx, not args.ssim, args.gpu_id, args.height, args.width)
    else:
        assert code_list is not None, 'No code files for decoding'
        if args.dec:
            assert out_list is not None, 'No out files for saving the decoded images'
            assert len(code_list) == len(out_list), 'The number of codes and reconstructed images should be the same'
            decoding(code_list, out_list, midx, not args.ssim, args.gpu_id, args.height, args.width)
        else:
            assert img_list is not None, 'No source images for evaluation.'
            assert len(code_list) == len(img_list), 'The number of codes and corresponding source images should be the same'
            decoding_and_test(code_list, img_list, midx, not args.ssim, args.gpu_id, args.height, args.width)


if __name__ == '__main__':
    main()
