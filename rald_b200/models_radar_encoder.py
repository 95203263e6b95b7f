"""Drop-in for the reference's ``model/models_radar_encoder.py`` (3-D conv radar-cube encoder).

Same class names, constructor arguments, parameter names / shapes / registration order and RNG consumption
order as the reference (so ``state_dict`` round-trips and seeded construction gives identical weights),
but ``Encoder.forward`` runs on the sm_100a kernels of librald_b200 (channels-last bf16 implicit-GEMM conv3d
on tcgen05 + fused GroupNorm/swish passes) instead of cuDNN / ATen. The parameters stay fp32 ``nn.Parameter``s
(the source of truth); packed bf16 device copies are rebuilt whenever they change.

Reference: model/models_radar_encoder.py:5-12 (Normalize, swish), :29-44 (Downsample), :46-100 (ResnetBlock),
:102-135 (AttnBlock), :137-241 (Encoder), :366-445 (RadarAutoencoder + factories).
The Decoder (radar-AE pre-training only) is kept as a parameter container; its forward is out
of scope and raises.
"""
from __future__ import annotations

import torch
import torch.nn as nn


def _group_norm(ch: int) -> nn.GroupNorm:
    return nn.GroupNorm(num_groups=32, num_channels=ch, eps=1e-6, affine=True)


def _conv3(cin: int, cout: int, k: int = 3, stride: int = 1, pad: int = 1) -> nn.Conv3d:
    return nn.Conv3d(cin, cout, kernel_size=k, stride=stride, padding=pad)


class ResnetBlock(nn.Module):
    """Parameter container: norm1, conv1, norm2, conv2 (+ nin_shortcut when channels change)."""

    def __init__(self, *, in_channels, out_channels=None, conv_shortcut=False, dropout=0.0, temb_channels=512):
        super().__init__()
        out_channels = in_channels if out_channels is None else out_channels
        self.in_channels, self.out_channels = in_channels, out_channels
        self.use_conv_shortcut = conv_shortcut
        self.norm1 = _group_norm(in_channels)
        self.conv1 = _conv3(in_channels, out_channels)
        if temb_channels > 0:
            self.temb_proj = nn.Linear(temb_channels, out_channels)
        self.norm2 = _group_norm(out_channels)
        self.dropout = nn.Dropout(dropout)
        self.conv2 = _conv3(out_channels, out_channels)
        if in_channels != out_channels:
            if conv_shortcut:
                self.conv_shortcut = _conv3(in_channels, out_channels)
            else:
                self.nin_shortcut = _conv3(in_channels, out_channels, k=1, pad=0)


class AttnBlock(nn.Module):
    def __init__(self, in_channels):
        super().__init__()
        self.in_channels = in_channels
        self.norm = _group_norm(in_channels)
        self.q = _conv3(in_channels, in_channels, k=1, pad=0)
        self.k = _conv3(in_channels, in_channels, k=1, pad=0)
        self.v = _conv3(in_channels, in_channels, k=1, pad=0)
        self.proj_out = _conv3(in_channels, in_channels, k=1, pad=0)


class Downsample(nn.Module):
    def __init__(self, in_channels, with_conv):
        super().__init__()
        self.with_conv = with_conv
        if with_conv:
            self.conv = _conv3(in_channels, in_channels, stride=2, pad=0)


class Upsample(nn.Module):
    def __init__(self, in_channels, with_conv):
        super().__init__()
        self.with_conv = with_conv
        if with_conv:
            self.conv = _conv3(in_channels, in_channels)


class Encoder(nn.Module):
    def __init__(self, *, ch=128, ch_mult=(1, 1, 2, 2, 4), num_res_blocks=2, attn_resolutions=((8, 4, 2),),
                 dropout=0.0, resamp_with_conv=True, in_channels=2, resolution=(128, 64, 32), z_channels=16,
                 **ignore_kwargs):
        super().__init__()
        self.ch, self.temb_ch = ch, 0
        self.ch_mult = tuple(ch_mult)
        self.num_resolutions = len(ch_mult)
        self.num_res_blocks = num_res_blocks
        self.resolution = resolution
        self.in_channels = in_channels
        self.z_channels = z_channels
        self.conv_in = _conv3(in_channels, ch)
        res = tuple(resolution)
        widths = (1,) + tuple(ch_mult)
        self.down = nn.ModuleList()
        c_prev = ch
        for lvl in range(self.num_resolutions):
            blocks, attns = nn.ModuleList(), nn.ModuleList()
            c_prev, c_out = ch * widths[lvl], ch * ch_mult[lvl]
            for _ in range(num_res_blocks):
                # construction order (block, then its attention) fixes the RNG stream of seeded init
                blocks.append(ResnetBlock(in_channels=c_prev, out_channels=c_out, temb_channels=0, dropout=dropout))
                c_prev = c_out
                if res in attn_resolutions:
                    attns.append(AttnBlock(c_prev))
            stage = nn.Module()
            stage.block = blocks
            stage.attn = attns
            if lvl != self.num_resolutions - 1:
                stage.downsample = Downsample(c_prev, resamp_with_conv)
                res = tuple(int(r / 2) for r in res)
            self.down.append(stage)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(in_channels=c_prev, out_channels=c_prev, temb_channels=0, dropout=dropout)
        self.mid.attn_1 = AttnBlock(c_prev)
        self.mid.block_2 = ResnetBlock(in_channels=c_prev, out_channels=c_prev, temb_channels=0, dropout=dropout)
        self.norm_out = _group_norm(c_prev)
        self.conv_out = _conv3(c_prev, z_channels)
        self._plan = None

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x: [B, Cin, R, A, E] fp32 -> [B, z, R/16, A/16, E/16] fp32 (reference :216-241)."""
        from .runtime_encoder import encoder_forward
        return encoder_forward(self, x, channels_last_out=False)

    def forward_channels_last(self, x_bdhwc: torch.Tensor) -> torch.Tensor:
        """x: [B, R, A, E, Cin] fp32 -> [B, R/16, A/16, E/16, z] fp32, skipping both permutes of
        EDMPrecond.process_radar_cond (model/models_radar_generation.py:384-387)."""
        from .runtime_encoder import encoder_forward
        return encoder_forward(self, x_bdhwc, channels_last_in=True, channels_last_out=True)


class Decoder(nn.Module):
    """Parameter container for checkpoint compatibility (radar-AE pre-training only; forward out of scope)."""

    def __init__(self, *, ch=128, out_ch=2, ch_mult=(1, 1, 2, 2, 4), num_res_blocks=2, attn_resolutions=(),
                 dropout=0.0, resamp_with_conv=True, in_channels=2, resolution=(128, 64, 32), z_channels=16,
                 give_pre_end=False, **ignore_kwargs):
        super().__init__()
        self.ch, self.temb_ch = ch, 0
        self.num_resolutions = len(ch_mult)
        self.num_res_blocks = num_res_blocks
        self.resolution = resolution
        self.in_channels = in_channels
        self.give_pre_end = give_pre_end
        c = ch * ch_mult[-1]
        res = tuple(int(r // 2 ** (self.num_resolutions - 1)) for r in resolution)
        self.z_shape = (1, z_channels, *res)
        self.conv_in = _conv3(z_channels, c)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(in_channels=c, out_channels=c, temb_channels=0, dropout=dropout)
        self.mid.attn_1 = AttnBlock(c)
        self.mid.block_2 = ResnetBlock(in_channels=c, out_channels=c, temb_channels=0, dropout=dropout)
        self.up = nn.ModuleList()
        for lvl in reversed(range(self.num_resolutions)):
            blocks, attns = nn.ModuleList(), nn.ModuleList()
            c_out = ch * ch_mult[lvl]
            for _ in range(num_res_blocks + 1):
                blocks.append(ResnetBlock(in_channels=c, out_channels=c_out, temb_channels=0, dropout=dropout))
                c = c_out
                if res in attn_resolutions:
                    attns.append(AttnBlock(c))
            stage = nn.Module()
            stage.block = blocks
            stage.attn = attns
            if lvl != 0:
                stage.upsample = Upsample(c, resamp_with_conv)
                res = res * 2  # (sic) tuple repetition as in the reference :325; never matches attn_resolutions=()
            self.up.insert(0, stage)
        self.norm_out = _group_norm(c)
        self.conv_out = _conv3(c, out_ch)

    def forward(self, z):
        raise NotImplementedError("rald_b200: the radar Decoder is outside the generation hot path (SURVEY.md §2 #2)")


class RadarAutoencoder(nn.Module):
    def __init__(self, *, basic_channel=128, ch_mult=(1, 1, 2, 2, 4), num_res_blocks=2, embed_dim=16):
        super().__init__()
        self.encoder = Encoder(ch=basic_channel, ch_mult=ch_mult, num_res_blocks=num_res_blocks, z_channels=embed_dim)
        self.decoder = Decoder(ch=basic_channel, ch_mult=ch_mult, num_res_blocks=num_res_blocks, z_channels=embed_dim)
        self.embed_dim = embed_dim

    def encode(self, x: torch.Tensor) -> torch.Tensor:
        return self.encoder(x)

    def decode(self, z: torch.Tensor) -> torch.Tensor:
        return self.decoder(z)

    def _encode(self, x: torch.Tensor) -> torch.Tensor:
        """[B, R, A, E, 2] -> [B, R/16, A/16, E/16, embed_dim] (reference :390-393)."""
        return self.encoder.forward_channels_last(x)

    def forward(self, inputs):
        raise NotImplementedError("rald_b200: radar autoencoder reconstruction is outside the generation hot path")


def create_autoencoder(basic_channel=128, ch_mult=(1, 1, 2, 2, 4), num_res_blocks=2, embed_dim=16):
    return RadarAutoencoder(basic_channel=basic_channel, ch_mult=ch_mult, num_res_blocks=num_res_blocks,
                            embed_dim=embed_dim)


def ae_ch128_mult5_n2_d16():
    return create_autoencoder(basic_channel=128)


def ae_ch64_mult5_n2_d16():
    return create_autoencoder(basic_channel=64)


def ae_ch16_mult5_n2_d16():
    return create_autoencoder(basic_channel=16)
