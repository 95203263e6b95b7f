"""Configuration helpers: an EasyDict-compatible mapping and the hyper-parameters of the reference's shipped
generation config (configs/generation/ge_indoor_cfg_aniso_mix_view_cone_unfreeze_enc_ints_only_eval.yml:116-148),
so the default models can be built where the reference tree / easydict are not installed (the GPU box)."""
from __future__ import annotations


class AttrDict(dict):
    """dict with attribute access, recursive wrapping and .get() — what the reference expects of EasyDict."""

    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in {**(d or {}), **kw}.items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, AttrDict):
            v = AttrDict(v)
        super().__setitem__(k, v)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    __setattr__ = __setitem__


# ar_model.configs of the shipped eval YAML (use_radar_enc anchors to `true` there)
DEFAULT_DENOISER_NAME = "kl_d512_m512_l32_d24_edm"
DEFAULT_AE_NAME = "kl_d512_m512_l32_mix"
DEFAULT_NUM_POINTS = 10000          # dataset.lidar.num_samples
DEFAULT_NUM_QUERY_POINTS = 500000   # eval.inference.num_query_points
RADAR_CUBE_SHAPE = (128, 64, 32, 2)  # range, azimuth, elevation, (intensity, doppler)


def default_denoiser_configs() -> AttrDict:
    return AttrDict(
        cond_type="radar", categories_num=5, use_radar_cond=True, use_radar_enc=True, unfreeze_radar_enc=True,
        input_radar_r_dim=128, input_radar_a_dim=8, input_radar_e_dim=2, input_radar_ch=2,
        enc_radar_r_dim=8, enc_radar_a_dim=4, enc_radar_e_dim=2, enc_radar_ch=16, enc_hidden_ch=64,
        radar_token_channel=512, sos_from_radar=True, use_radar_dopp=False)
