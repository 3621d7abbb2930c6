"""ORACLE (test infrastructure, not product code) — PARITY UNPINNED.

fp32 ``torch.nn`` restatement of the reference's network definitions (its ``models/networks.py``).
The reference mount ships NO source (``/root/reference/.gitignore:41-42`` ignores ``*.py``), so this
file follows, in order of authority:

  1. the launch flags  [REF test_start/start.sh:15-21,24; pretrainTrans.sh:11-13;
     train_start/pretrain_start.sh:25,29-37] — widths, depths, channel counts;
  2. README.md:101 "borrows heavily from pix2pixHD" — the GlobalGenerator / ResnetBlock /
     MultiscaleDiscriminator structure and the ``model.<idx>`` state_dict naming of public
     NVIDIA/pix2pixHD ``models/networks.py`` (un-vendored dependency, no version pinned; restated
     from its published structure);
  3. SPEC decisions D1-D14 (DESIGN.md §"Frozen decisions") for everything neither pins.

The arithmetic itself lives in PyTorch (the reference pins torch 1.1.0 through
``torchvision==0.3.0`` [REF requirment.txt:5]); Conv2d, ConvTranspose2d, ReflectionPad2d and
InstanceNorm2d(affine=False, eps=1e-5) semantics are unchanged in torch 2.11 used here.

There are no golden vectors in the reference; ``tests/golden/`` holds vectors
generated from THIS file by ``tests/golden/make_golden.py``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this package.
"""
from __future__ import annotations

import functools
from typing import List, Sequence

import torch
import torch.nn as nn

N_PARTS = 24            # DensePose body parts                      [SPEC D3]
UV_CHANNELS = 25 + 24 + 24   # part logits (bg + 24) + U + V         [SPEC D3]


def get_norm_layer(norm_type: str = "instance"):
    """pix2pixHD ``get_norm_layer`` [UPSTREAM, SURVEY Appendix C]."""
    if norm_type == "instance":
        return functools.partial(nn.InstanceNorm2d, affine=False)
    if norm_type == "batch":
        return functools.partial(nn.BatchNorm2d, affine=True)
    raise NotImplementedError("normalization layer [%s] is not found" % norm_type)


def weights_init(m: nn.Module) -> None:
    """pix2pixHD ``weights_init``: Conv* weight ~ N(0, 0.02) [UPSTREAM, SURVEY Appendix C]."""
    classname = m.__class__.__name__
    if classname.find("Conv") != -1 and hasattr(m, "weight"):
        m.weight.data.normal_(0.0, 0.02)
    elif classname.find("BatchNorm2d") != -1:
        m.weight.data.normal_(1.0, 0.02)
        m.bias.data.fill_(0)


class ResnetBlock(nn.Module):
    """x + IN(conv3(rpad1(ReLU(IN(conv3(rpad1(x)))))))   [UPSTREAM ResnetBlock; SURVEY §8 a5]."""

    def __init__(self, dim: int, norm_layer, activation=None):
        super().__init__()
        activation = activation or nn.ReLU(True)
        self.conv_block = nn.Sequential(
            nn.ReflectionPad2d(1), nn.Conv2d(dim, dim, kernel_size=3, padding=0), norm_layer(dim), activation,
            nn.ReflectionPad2d(1), nn.Conv2d(dim, dim, kernel_size=3, padding=0), norm_layer(dim),
        )

    def forward(self, x):
        return x + self.conv_block(x)


class GlobalGenerator(nn.Module):
    """pix2pixHD GlobalGenerator shape [UPSTREAM; widths from REF start.sh:15-17].

    rpad3-conv7-IN-ReLU -> n_down x (conv3 s2 p1-IN-ReLU) -> n_blocks x ResnetBlock
    -> n_down x (convT3 s2 p1 op1-IN-ReLU) -> rpad3-conv7 -> ``final`` activation.

    ``final``: 'tanh' (upstream), 'none' (UV generator, SPEC D3) or 'tanh_sigmoid_last'
    (RGB tanh + mask sigmoid, SPEC D9).
    """

    def __init__(self, input_nc: int, output_nc: int, ngf: int = 64, n_downsampling: int = 3, n_blocks: int = 9,
                 norm_layer=None, final: str = "tanh"):
        super().__init__()
        assert n_blocks >= 0
        norm_layer = norm_layer or get_norm_layer("instance")
        activation = nn.ReLU(True)
        model: List[nn.Module] = [nn.ReflectionPad2d(3), nn.Conv2d(input_nc, ngf, kernel_size=7, padding=0),
                                  norm_layer(ngf), activation]
        for i in range(n_downsampling):
            mult = 2 ** i
            model += [nn.Conv2d(ngf * mult, ngf * mult * 2, kernel_size=3, stride=2, padding=1),
                      norm_layer(ngf * mult * 2), activation]
        mult = 2 ** n_downsampling
        for _ in range(n_blocks):
            model += [ResnetBlock(ngf * mult, norm_layer=norm_layer, activation=activation)]
        for i in range(n_downsampling):
            mult = 2 ** (n_downsampling - i)
            model += [nn.ConvTranspose2d(ngf * mult, int(ngf * mult / 2), kernel_size=3, stride=2, padding=1,
                                         output_padding=1),
                      norm_layer(int(ngf * mult / 2)), activation]
        model += [nn.ReflectionPad2d(3), nn.Conv2d(ngf, output_nc, kernel_size=7, padding=0)]
        if final == "tanh":
            model += [nn.Tanh()]
        self.final = final
        self.input_nc, self.output_nc, self.ngf = input_nc, output_nc, ngf
        self.n_downsampling, self.n_blocks = n_downsampling, n_blocks
        self.model = nn.Sequential(*model)

    def forward(self, x):
        y = self.model(x)
        if self.final == "tanh_sigmoid_last":
            y = torch.cat([torch.tanh(y[:, :-1]), torch.sigmoid(y[:, -1:])], dim=1)
        return y


class NLayerDiscriminator(nn.Module):
    """pix2pixHD NLayerDiscriminator: kw=4, padw=2 [UPSTREAM, SURVEY Appendix C / §8 a8]."""

    def __init__(self, input_nc: int, ndf: int = 64, n_layers: int = 3, norm_layer=None, use_sigmoid: bool = False,
                 getIntermFeat: bool = False):
        super().__init__()
        norm_layer = norm_layer or get_norm_layer("instance")
        self.getIntermFeat = getIntermFeat
        self.n_layers = n_layers
        kw, padw = 4, 2
        sequence = [[nn.Conv2d(input_nc, ndf, kernel_size=kw, stride=2, padding=padw), nn.LeakyReLU(0.2, True)]]
        nf = ndf
        for _ in range(1, n_layers):
            nf_prev, nf = nf, min(nf * 2, 512)
            sequence += [[nn.Conv2d(nf_prev, nf, kernel_size=kw, stride=2, padding=padw), norm_layer(nf),
                          nn.LeakyReLU(0.2, True)]]
        nf_prev, nf = nf, min(nf * 2, 512)
        sequence += [[nn.Conv2d(nf_prev, nf, kernel_size=kw, stride=1, padding=padw), norm_layer(nf),
                      nn.LeakyReLU(0.2, True)]]
        sequence += [[nn.Conv2d(nf, 1, kernel_size=kw, stride=1, padding=padw)]]
        if use_sigmoid:
            sequence += [[nn.Sigmoid()]]
        if getIntermFeat:
            for n in range(len(sequence)):
                setattr(self, "model" + str(n), nn.Sequential(*sequence[n]))
        else:
            stream = []
            for n in range(len(sequence)):
                stream += sequence[n]
            self.model = nn.Sequential(*stream)

    def forward(self, x):
        if self.getIntermFeat:
            res = [x]
            for n in range(self.n_layers + 2):
                res.append(getattr(self, "model" + str(n))(res[-1]))
            return res[1:]
        return self.model(x)


class MultiscaleDiscriminator(nn.Module):
    """pix2pixHD MultiscaleDiscriminator [UPSTREAM; defaults num_D=2, n_layers=3, ndf=64, LSGAN]."""

    def __init__(self, input_nc: int, ndf: int = 64, n_layers: int = 3, norm_layer=None, use_sigmoid: bool = False,
                 num_D: int = 3, getIntermFeat: bool = False):
        super().__init__()
        self.num_D, self.n_layers, self.getIntermFeat = num_D, n_layers, getIntermFeat
        for i in range(num_D):
            netD = NLayerDiscriminator(input_nc, ndf, n_layers, norm_layer, use_sigmoid, getIntermFeat)
            if getIntermFeat:
                for j in range(n_layers + 2):
                    setattr(self, "scale" + str(i) + "_layer" + str(j), getattr(netD, "model" + str(j)))
            else:
                setattr(self, "layer" + str(i), netD.model)
        self.downsample = nn.AvgPool2d(3, stride=2, padding=[1, 1], count_include_pad=False)

    def singleD_forward(self, model, x):
        if self.getIntermFeat:
            result = [x]
            for m in model:
                result.append(m(result[-1]))
            return result[1:]
        return [model(x)]

    def forward(self, x):
        num_D = self.num_D
        result = []
        xd = x
        for i in range(num_D):
            if self.getIntermFeat:
                model = [getattr(self, "scale" + str(num_D - 1 - i) + "_layer" + str(j)) for j in range(self.n_layers + 2)]
            else:
                model = getattr(self, "layer" + str(num_D - 1 - i))
            result.append(self.singleD_forward(model, xd))
            if i != (num_D - 1):
                xd = self.downsample(xd)
        return result


def define_G(input_nc: int, output_nc: int, ngf: int, netG: str = "global", n_downsample_global: int = 3,
             n_blocks_global: int = 9, n_local_enhancers: int = 1, n_blocks_local: int = 3, norm: str = "instance",
             gpu_ids: Sequence[int] = ()):
    """pix2pixHD ``define_G`` signature [UPSTREAM; named by BASELINE.json].

    netG: 'global' (tanh RGB), 'temporal' (RGB tanh + mask sigmoid — the main generator named
    ``*_Temporal`` REF start.sh:7), 'translate' (UV generator "TransG", REF pretrainTrans.sh:13,
    73 raw channels), 'bg' (background refinement net, REF start.sh:20-21).
    """
    norm_layer = get_norm_layer(norm)
    final = {"global": "tanh", "bg": "tanh", "temporal": "tanh_sigmoid_last", "translate": "none"}.get(netG)
    if final is None:
        raise NotImplementedError("generator [%s] not implemented" % netG)
    net = GlobalGenerator(input_nc, output_nc, ngf, n_downsample_global, n_blocks_global, norm_layer, final=final)
    net.apply(weights_init)
    return net


def define_D(input_nc: int, ndf: int, n_layers_D: int, norm: str = "instance", use_sigmoid: bool = False,
             num_D: int = 1, getIntermFeat: bool = False, gpu_ids: Sequence[int] = ()):
    """pix2pixHD ``define_D`` signature [UPSTREAM; named by BASELINE.json]."""
    norm_layer = get_norm_layer(norm)
    net = MultiscaleDiscriminator(input_nc, ndf, n_layers_D, norm_layer, use_sigmoid, num_D, getIntermFeat)
    net.apply(weights_init)
    return net


class Vgg19(nn.Module):
    """pix2pixHD ``Vgg19``: torchvision ``vgg19().features`` cut into the five slices that end at relu1_1, relu2_1, relu3_1,
    relu4_1, relu5_1 [UPSTREAM models/networks.py Vgg19; torchvision is not needed: the stack is 13 Conv2d(3x3, padding 1) + ReLU and
    4 MaxPool2d(2, 2), parameter names ``features.<idx>``].  Weights: whatever the caller loads (random in the parity test)."""

    CFG = (64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M", 512, 512, 512, 512, "M", 512)
    CUTS = (2, 7, 12, 21, 30)

    def __init__(self):
        super().__init__()
        layers, cin = [], 3
        for v in self.CFG:
            if v == "M":
                layers.append(nn.MaxPool2d(2, 2))
            else:
                layers += [nn.Conv2d(cin, v, 3, padding=1), nn.ReLU(inplace=False)]
                cin = v
        self.features = nn.Sequential(*layers)
        for m in self.features:
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
                nn.init.zeros_(m.bias)

    def forward(self, x):
        feats, lo = [], 0
        for hi in self.CUTS:
            x = self.features[lo:hi](x)
            feats.append(x)
            lo = hi
        return feats
