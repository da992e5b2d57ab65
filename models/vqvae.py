"""Drop-in `models/vqvae.py`: the reference's DualMotionVQVAE with its quantizer layer replaced by
the vqb200 CUDA engine.

Scripts written against the reference (`scripts/train_ablation.py:18-21`,
`scripts/deployment/export_motion.py:10-14`, `scripts/evaluation/analyze_latent_space.py:12-13`)
do `from models.vqvae import DualMotionVQVAE` and keep working: constructor signatures, the
forward dictionaries and every state_dict key are those of reference `models/vqvae.py:508-617`.

Only the quantizers (reference :10-259) are new code -- they come from the `vqb200` package.  The
encoders / decoders are stock `torch.nn` layers (out of scope as kernels) and
are re-declared here solely so that checkpoints load with identical keys.
"""
import math
import os
import sys

import torch
import torch.nn as nn

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

import vqb200  # noqa: E402
from vqb200 import VectorQuantizer, ResidualVQ, FSQ, LFQ, HybridVQ, IdentityVQ  # noqa: E402,F401


# ----------------------------------------------------------------------------------------------
# convolutional building blocks (reference :265-410)
# ----------------------------------------------------------------------------------------------
def _conv_bn_act(ch):
    return [nn.Conv1d(ch, ch, 3, 1, 1), nn.BatchNorm1d(ch), nn.LeakyReLU(0.2, inplace=True)]


class ResBlock1D(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.net = nn.Sequential(*_conv_bn_act(channels), *_conv_bn_act(channels))

    def forward(self, x):
        return x + self.net(x)


def _sinusoid_table(max_len, d_model):
    pos = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
    freq = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
    table = torch.zeros(max_len, d_model)
    table[:, 0::2] = torch.sin(pos * freq)
    table[:, 1::2] = torch.cos(pos * freq)
    return table


class PositionalEncoding(nn.Module):
    """Channel-major variant (unused by DualMotionVQVAE; kept for API parity, reference :280-291)."""

    def __init__(self, d_model, max_len=5000):
        super().__init__()
        self.register_buffer("pe", _sinusoid_table(max_len, d_model).unsqueeze(0).transpose(1, 2))

    def forward(self, x):
        return x + self.pe[:, :, :x.size(2)]


def _down(cin, cout):
    return [nn.Conv1d(cin, cout, 4, 2, 1), nn.LeakyReLU(0.2)]


class Encoder(nn.Module):
    """Strided conv encoders 'simple' / 'resnet' (T -> T/4), reference :293-325."""

    def __init__(self, input_dim, hidden_dim, arch="simple", num_res_layers=4):
        super().__init__()
        self.arch = arch
        if arch == "resnet":
            layers = _down(input_dim, hidden_dim)
            layers += [ResBlock1D(hidden_dim) for _ in range(num_res_layers)]
            layers += _down(hidden_dim, hidden_dim) + [ResBlock1D(hidden_dim)]
        else:
            layers = _down(input_dim, hidden_dim) + _down(hidden_dim, hidden_dim)
        self.model = nn.Sequential(*layers)

    def forward(self, x):
        return self.model(x)


class Decoder(nn.Module):
    """Mirror of `Encoder`, reference :327-365."""

    def __init__(self, input_dim, hidden_dim, arch="simple", num_res_layers=4):
        super().__init__()
        self.arch = arch
        if arch == "resnet":
            layers = [ResBlock1D(hidden_dim) for _ in range(num_res_layers)]
            layers += [nn.Upsample(scale_factor=2.0, mode="nearest"), nn.Conv1d(hidden_dim, hidden_dim, 3, 1, 1),
                       nn.LeakyReLU(0.2), ResBlock1D(hidden_dim),
                       nn.Upsample(scale_factor=2.0, mode="nearest"), nn.Conv1d(hidden_dim, input_dim, 3, 1, 1)]
        else:
            layers = [nn.ConvTranspose1d(hidden_dim, hidden_dim, 4, 2, 1), nn.LeakyReLU(0.2),
                      nn.ConvTranspose1d(hidden_dim, input_dim, 4, 2, 1)]
        self.model = nn.Sequential(*layers)

    def forward(self, x):
        return self.model(x)


class NoDownsampleEncoder(nn.Module):
    """Full-resolution ResNet encoder ([B,C,T] -> [B,hidden,T]), reference :370-391."""

    def __init__(self, input_dim, hidden_dim, num_res_layers=4):
        super().__init__()
        self.model = nn.Sequential(nn.Conv1d(input_dim, hidden_dim, kernel_size=3, stride=1, padding=1),
                                   nn.LeakyReLU(0.2, inplace=True))
        for i in range(num_res_layers):
            self.model.add_module(f"res_{i}", ResBlock1D(hidden_dim))
        self.model.add_module("final_conv", nn.Conv1d(hidden_dim, hidden_dim, 3, 1, 1))
        self.model.add_module("final_act", nn.LeakyReLU(0.2, inplace=True))

    def forward(self, x):
        return self.model(x)


class NoDownsampleDecoder(nn.Module):
    """Full-resolution ResNet decoder, reference :393-410."""

    def __init__(self, output_dim, hidden_dim, num_res_layers=4):
        super().__init__()
        self.model = nn.Sequential()
        for i in range(num_res_layers):
            self.model.add_module(f"res_{i}", ResBlock1D(hidden_dim))
        self.model.add_module("out_conv", nn.Conv1d(hidden_dim, output_dim, kernel_size=3, stride=1, padding=1))

    def forward(self, x):
        return self.model(x)


# ----------------------------------------------------------------------------------------------
# transformer encoder / decoder (reference :412-499)
# ----------------------------------------------------------------------------------------------
class TransformerPositionalEncoding(nn.Module):
    def __init__(self, d_model, max_len=5000):
        super().__init__()
        self.register_buffer("pe", _sinusoid_table(max_len, d_model).unsqueeze(0))     # (1, T, C)

    def forward(self, x):
        return x + self.pe[:, :x.size(1), :]


def _backbone(d_model, nhead, num_layers):
    layer = nn.TransformerEncoderLayer(d_model=d_model, nhead=nhead, dim_feedforward=512, batch_first=True)
    return nn.TransformerEncoder(layer, num_layers=num_layers)


class TransformerMotionEncoder(nn.Module):
    """[B,C,T] -> tokens -> mean-pool -> one latent per window, returned as the permuted view [B,hidden,1]."""

    def __init__(self, input_dim, hidden_dim, d_model=256, nhead=4, num_layers=4):
        super().__init__()
        self.input_proj = nn.Linear(input_dim, d_model)
        self.pe = TransformerPositionalEncoding(d_model)
        self.transformer = _backbone(d_model, nhead, num_layers)
        self.output_proj = nn.Linear(d_model, hidden_dim)

    def forward(self, x):
        h = self.pe(self.input_proj(x.permute(0, 2, 1)))
        h = self.transformer(h)
        h = self.output_proj(torch.mean(h, dim=1, keepdim=True))      # [B, 1, hidden]
        return h.permute(0, 2, 1)


class TransformerMotionDecoder(nn.Module):
    """[B,hidden,1] latent broadcast to `seq_len` tokens + positional code -> [B,out,seq_len]."""

    def __init__(self, output_dim, hidden_dim, d_model=256, nhead=4, num_layers=4, seq_len=64):
        super().__init__()
        self.seq_len = seq_len
        self.input_proj = nn.Linear(hidden_dim, d_model)
        self.pe = TransformerPositionalEncoding(d_model)
        self.transformer = _backbone(d_model, nhead, num_layers)
        self.output_proj = nn.Linear(d_model, output_dim)

    def forward(self, x):
        h = self.input_proj(x.permute(0, 2, 1)).repeat(1, self.seq_len, 1)
        h = self.transformer(self.pe(h))
        return self.output_proj(h).permute(0, 2, 1)


# ----------------------------------------------------------------------------------------------
# dual-encoder VQ-VAE (reference :508-617)
# ----------------------------------------------------------------------------------------------
def _make_quantizer(method, codebook_size, hidden_dim, n_layers):
    """method -> quantizer wiring of reference :540-560."""
    if method == "standard":
        return VectorQuantizer(codebook_size, hidden_dim, use_ema=False)
    if method == "ema":
        return VectorQuantizer(codebook_size, hidden_dim, use_ema=True)
    if method == "rvq":
        return ResidualVQ(num_quantizers=n_layers, num_embeddings=codebook_size, embedding_dim=hidden_dim, use_ema=True)
    if method == "fsq":
        return FSQ(levels=[8, 5, 5, 5], input_dim=hidden_dim, hidden_dim=hidden_dim)
    if method == "lfq":
        return LFQ(input_dim=hidden_dim, codebook_dim=10)
    if method == "hybrid":
        return HybridVQ(hidden_dim=hidden_dim, fsq_levels=[8, 5, 5, 5], vq_codebook_size=512)
    if method == "ae":
        return IdentityVQ()
    raise ValueError(f"Unknown quantization method: {method}")


class DualMotionVQVAE(nn.Module):
    def __init__(self, human_input_dim=263, robot_input_dim=29, hidden_dim=64, codebook_size=1024,
                 arch="transformer", method="hybrid", n_layers=4, window_size=64):
        super().__init__()
        self.arch = arch
        self.window_size = window_size
        if arch == "transformer":
            self.human_encoder = TransformerMotionEncoder(human_input_dim, hidden_dim, d_model=256, num_layers=4)
            self.robot_encoder = TransformerMotionEncoder(robot_input_dim, hidden_dim, d_model=256, num_layers=4)
        elif arch == "resnet_no_down":
            self.human_encoder = NoDownsampleEncoder(human_input_dim, hidden_dim)
            self.robot_encoder = NoDownsampleEncoder(robot_input_dim, hidden_dim)
        else:
            self.human_encoder = Encoder(human_input_dim, hidden_dim, arch=arch)
            self.robot_encoder = Encoder(robot_input_dim, hidden_dim, arch=arch)

        self.quantizer = _make_quantizer(method, codebook_size, hidden_dim, n_layers)

        if arch == "transformer":
            self.robot_decoder = TransformerMotionDecoder(robot_input_dim, hidden_dim, d_model=256, num_layers=4,
                                                          seq_len=window_size)
        elif arch == "resnet_no_down":
            self.robot_decoder = NoDownsampleDecoder(robot_input_dim, hidden_dim)
        else:
            self.robot_decoder = Decoder(robot_input_dim, hidden_dim, arch=arch)

    def _branch(self, x, encoder, out_key):
        z_e = encoder(x.permute(0, 2, 1))                              # [B, hidden, T']
        loss_vq, z_q, metrics = self.quantizer(z_e)
        y = self.robot_decoder(z_q)
        return {out_key: y.permute(0, 2, 1), "loss_vq": loss_vq, "metrics": metrics, "z_e": z_e}

    # ---- token export / decode-only path (not in the reference) --------------------------
    def encode_tokens(self, x_human=None, x_robot=None, digit_bits: int = 8):
        """Motion windows [B, T, dim] -> compact tokens (vqb200.tokens.TokenBatch); eval semantics, no state change."""
        import vqb200
        if (x_human is None) == (x_robot is None):
            raise ValueError("encode_tokens: pass exactly one of x_human / x_robot")
        enc, x = (self.human_encoder, x_human) if x_human is not None else (self.robot_encoder, x_robot)
        with torch.no_grad():
            z_e = enc(x.permute(0, 2, 1))
        return vqb200.tokens.encode(self.quantizer, z_e.contiguous(), digit_bits=digit_bits)

    def decode_tokens(self, tokens):
        """Tokens -> robot motion [B, T, robot_dim] through the quantizer's decode-only path and `robot_decoder`."""
        import vqb200
        with torch.no_grad():
            z_q = vqb200.tokens.decode(self.quantizer, tokens)
            return self.robot_decoder(z_q).permute(0, 2, 1)

    def forward(self, x_robot=None, x_human=None):
        outputs = {}
        if x_robot is not None:
            outputs["robot"] = self._branch(x_robot, self.robot_encoder, "recon")
        if x_human is not None:
            outputs["human"] = self._branch(x_human, self.human_encoder, "retargeted")
        return outputs
