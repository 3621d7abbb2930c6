"""ORACLE (test infrastructure) — PARITY UNPINNED.  A second, INDEPENDENT restatement of the network arithmetic in
plain numpy, written from the operator definitions (not from torch): ReflectionPad2d, Conv2d, ConvTranspose2d
(stride 2, padding 1, output_padding 1), InstanceNorm2d(affine=False, eps=1e-5, biased variance), ReLU / LeakyReLU(0.2) /
tanh / sigmoid, AvgPool2d(3, stride 2, padding 1, count_include_pad=False), assembled into pix2pixHD's GlobalGenerator
(with its ResnetBlock) and NLayerDiscriminator / MultiscaleDiscriminator [public NVIDIA/pix2pixHD models/networks.py;
SURVEY.md Appendix C].  tests/test_cpu_oracle.py checks oracle/networks.py (the torch.nn oracle) against this file on
the committed golden vectors: a mis-restatement in either one shows up as a disagreement.  Small cases only.
"""
from __future__ import annotations

import numpy as np


def reflect_pad(x: np.ndarray, p: int) -> np.ndarray:
    """ReflectionPad2d(p): index -k mirrors to k, index n-1+k mirrors to n-1-k (the edge is not repeated)."""
    n, c, h, w = x.shape
    ys = [abs(i) if i < h else 2 * (h - 1) - i for i in range(-p, h + p)]
    xs = [abs(j) if j < w else 2 * (w - 1) - j for j in range(-p, w + p)]
    return x[:, :, ys][:, :, :, xs]


def zero_pad(x: np.ndarray, p: int) -> np.ndarray:
    n, c, h, w = x.shape
    out = np.zeros((n, c, h + 2 * p, w + 2 * p), x.dtype)
    out[:, :, p:p + h, p:p + w] = x
    return out


def conv2d(x: np.ndarray, w: np.ndarray, b: np.ndarray, stride: int = 1) -> np.ndarray:
    """Cross-correlation (what nn.Conv2d computes) of an already padded input: out[o,y,x] = sum_{c,r,s} w[o,c,r,s] x[c, y*st+r, x*st+s]."""
    n, c, h, wd = x.shape
    o, c2, kh, kw = w.shape
    assert c == c2
    ho, wo = (h - kh) // stride + 1, (wd - kw) // stride + 1
    out = np.zeros((n, o, ho, wo), np.float64)
    for r in range(kh):
        for s in range(kw):
            patch = x[:, :, r:r + (ho - 1) * stride + 1:stride, s:s + (wo - 1) * stride + 1:stride]
            out += np.einsum("nchw,oc->nohw", patch, w[:, :, r, s])
    return out + b.reshape(1, -1, 1, 1)


def conv_transpose2d_s2(x: np.ndarray, w: np.ndarray, b: np.ndarray) -> np.ndarray:
    """ConvTranspose2d(k=3, stride=2, padding=1, output_padding=1): every input pixel (i, j) scatters x * w[c, o, r, s] to
    output (2i - 1 + r, 2j - 1 + s); output size 2H x 2W.  w has the ConvTranspose2d layout [Cin][Cout][kh][kw]."""
    n, c, h, wd = x.shape
    c2, o, kh, kw = w.shape
    assert c == c2 and kh == 3 and kw == 3
    out = np.zeros((n, o, 2 * h, 2 * wd), np.float64)
    for i in range(h):
        for j in range(wd):
            contrib = np.einsum("nc,cors->nors", x[:, :, i, j], w)
            for r in range(3):
                for s in range(3):
                    yy, xx = 2 * i - 1 + r, 2 * j - 1 + s
                    if 0 <= yy < 2 * h and 0 <= xx < 2 * wd:
                        out[:, :, yy, xx] += contrib[:, :, r, s]
    return out + b.reshape(1, -1, 1, 1)


def instance_norm(x: np.ndarray, eps: float = 1e-5) -> np.ndarray:
    mean = x.mean(axis=(2, 3), keepdims=True)
    var = ((x - mean) ** 2).mean(axis=(2, 3), keepdims=True)          # biased, as InstanceNorm2d
    return (x - mean) / np.sqrt(var + eps)


def relu(x):
    return np.maximum(x, 0.0)


def lrelu(x, a=0.2):
    return np.where(x > 0, x, a * x)


def avgpool3s2(x: np.ndarray) -> np.ndarray:
    """AvgPool2d(3, stride=2, padding=1, count_include_pad=False): mean over the in-bounds part of each 3x3 window."""
    n, c, h, w = x.shape
    ho, wo = (h + 2 - 3) // 2 + 1, (w + 2 - 3) // 2 + 1
    out = np.zeros((n, c, ho, wo), np.float64)
    for i in range(ho):
        for j in range(wo):
            y0, y1 = max(2 * i - 1, 0), min(2 * i + 2, h)
            x0, x1 = max(2 * j - 1, 0), min(2 * j + 2, w)
            out[:, :, i, j] = x[:, :, y0:y1, x0:x1].mean(axis=(2, 3))
    return out


def global_generator(x: np.ndarray, sd: dict, n_down: int, n_blocks: int, final: str, prefix: str = "") -> np.ndarray:
    """pix2pixHD GlobalGenerator forward from a state_dict with upstream key names (model.<idx>...)."""
    g = lambda k: np.asarray(sd[prefix + k], np.float64)
    h = np.asarray(x, np.float64)
    idx = 1                                                  # model.0 = ReflectionPad2d(3)
    h = relu(instance_norm(conv2d(reflect_pad(h, 3), g("model.%d.weight" % idx), g("model.%d.bias" % idx))))
    idx += 3
    for _ in range(n_down):                                  # Conv2d(3, stride 2, padding 1) - zero padding
        h = relu(instance_norm(conv2d(zero_pad(h, 1), g("model.%d.weight" % idx), g("model.%d.bias" % idx), stride=2)))
        idx += 3
    for _ in range(n_blocks):                                # ResnetBlock: x + IN(conv(rpad(ReLU(IN(conv(rpad(x)))))))
        t = relu(instance_norm(conv2d(reflect_pad(h, 1), g("model.%d.conv_block.1.weight" % idx), g("model.%d.conv_block.1.bias" % idx))))
        t = instance_norm(conv2d(reflect_pad(t, 1), g("model.%d.conv_block.5.weight" % idx), g("model.%d.conv_block.5.bias" % idx)))
        h = h + t
        idx += 1
    for _ in range(n_down):
        h = relu(instance_norm(conv_transpose2d_s2(h, g("model.%d.weight" % idx), g("model.%d.bias" % idx))))
        idx += 3
    idx += 1                                                 # ReflectionPad2d(3)
    h = conv2d(reflect_pad(h, 3), g("model.%d.weight" % idx), g("model.%d.bias" % idx))
    if final == "tanh":
        h = np.tanh(h)
    elif final == "tanh_sigmoid_last":
        h = np.concatenate([np.tanh(h[:, :-1]), 1.0 / (1.0 + np.exp(-h[:, -1:]))], axis=1)
    return h


def nlayer_discriminator(x: np.ndarray, sd: dict, scale: int, n_layers: int, prefix: str = ""):
    """NLayerDiscriminator (kw 4, padw 2) with getIntermFeat: list of the n_layers + 2 feature maps."""
    g = lambda k: np.asarray(sd[prefix + k], np.float64)
    feats = []
    h = np.asarray(x, np.float64)
    for j in range(n_layers + 2):
        w, b = g("scale%d_layer%d.0.weight" % (scale, j)), g("scale%d_layer%d.0.bias" % (scale, j))
        stride = 2 if j < n_layers else 1
        h = conv2d(zero_pad(h, 2), w, b, stride=stride)
        if 0 < j < n_layers + 1:
            h = instance_norm(h)
        if j < n_layers + 1:
            h = lrelu(h)
        feats.append(h)
    return feats


def multiscale_discriminator(x: np.ndarray, sd: dict, num_D: int, n_layers: int, prefix: str = ""):
    out, xd = [], np.asarray(x, np.float64)
    for i in range(num_D):
        out.append(nlayer_discriminator(xd, sd, num_D - 1 - i, n_layers, prefix))
        if i != num_D - 1:
            xd = avgpool3s2(xd)
    return out
