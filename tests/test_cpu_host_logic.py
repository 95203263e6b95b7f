"""CPU: host-side logic of the drop-in modules — constructor / factory surface, state_dict contract, weight packing
permutations, schedules, frame sharding and the gloo (world_size 2) gather."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from helpers import build_ae, build_denoiser
from oracle import rald_oracle as orc
from rald_b200 import gather, models_ae, models_radar_encoder, models_radar_generation, postproc
from rald_b200.config import AttrDict, default_denoiser_configs
from rald_b200.runtime_dit import geglu_pack_index


def test_factories_exist():
    for n in ("kl_d512_m512_l8_edm", "kl_d512_m512_l16_edm", "kl_d512_m512_l32_edm", "kl_d512_m512_l4_d24_edm",
              "kl_d512_m512_l8_d24_edm", "kl_d512_m512_l32_d24_edm", "kl_d512_m512_l32_d18_edm",
              "kl_d512_m512_l32_d12_edm", "EDMLoss", "edm_sampler", "StackedRandomGenerator", "EDMPrecond"):
        assert n in models_radar_generation.__dict__
    for n in ("kl_d512_m512_l512", "kl_d512_m512_l64", "kl_d512_m512_l32", "kl_d512_m512_l32_learn",
              "kl_d512_m512_l32_mix", "kl_d512_m512_l16", "kl_d512_m512_l8", "kl_d512_m512_l4", "kl_d512_m512_l2",
              "kl_d512_m512_l1", "ae_d512_m512", "ae_d64_m512", "KLAutoEncoder", "AutoEncoder"):
        assert n in models_ae.__dict__
    for n in ("ae_ch128_mult5_n2_d16", "ae_ch64_mult5_n2_d16", "ae_ch16_mult5_n2_d16", "Encoder", "RadarAutoencoder"):
        assert n in models_radar_encoder.__dict__


def test_denoiser_state_dict_contract():
    net = build_denoiser("kl_d512_m512_l32_d12_edm")
    sd = net.state_dict()
    assert sd["model.proj_in.weight"].shape == (512, 32)
    assert sd["model.transformer_blocks.0.ff.net.0.proj.weight"].shape == (4096, 512)
    assert sd["model.transformer_blocks.11.norm3.linear.bias"].shape == (1024,)
    assert sd["radar_enc.down.2.block.0.nin_shortcut.weight"].shape == (128, 64, 1, 1, 1)
    assert sd["radar_enc.down.4.attn.1.proj_out.weight"].shape == (256, 256, 1, 1, 1)
    assert sd["radar_r_emb.weight"].shape == (8, 512) and sd["radar_token_project.weight"].shape == (512, 16)
    assert all(v.dtype == torch.float32 for v in sd.values())
    names = [n for n, _ in net.named_parameters()]
    assert names[0] == "model.proj_in.weight" and names[-1] == "radar_token_project.bias"
    # strict round trip, as main_generation.py / utils/misc.py load checkpoints
    other = build_denoiser("kl_d512_m512_l32_d12_edm")
    other.load_state_dict(sd, strict=True)
    # zero-initialised proj_out straight from the constructor (reference :198)
    fresh = models_radar_generation.kl_d512_m512_l32_d12_edm(configs=default_denoiser_configs())
    assert float(fresh.model.proj_out.weight.abs().max()) == 0.0


def test_ae_state_dict_contract():
    ae = build_ae("kl_d512_m512_l32_mix", n=2048)
    sd = ae.state_dict()
    assert sd["point_embed.basis"].shape == (3, 24) and sd["point_embed.mlp.weight"].shape == (512, 51)
    assert sd["layers.23.1.fn.net.0.weight"].shape == (4096, 512)
    assert sd["decoder_cross_attn.fn.to_kv.weight"].shape == (1024, 512)
    assert sd["mean_fc.weight"].shape == (32, 512) and sd["to_outputs.weight"].shape == (1, 512)
    assert isinstance(ae, models_ae.KLAutoEncoder)
    assert "s_latents.weight" not in build_ae("kl_d512_m512_l32", n=2048).state_dict()


def test_use_radar_enc_false_sizes_embeddings_from_input_dims():
    cfg = default_denoiser_configs()
    cfg["use_radar_enc"] = False
    cfg["unfreeze_radar_enc"] = False
    net = models_radar_generation.kl_d512_m512_l32_d12_edm(configs=cfg)
    assert net.radar_r_emb.weight.shape == (128, 512) and net.radar_token_project.weight.shape == (512, 1)
    assert not hasattr(net, "radar_enc")


def test_attrdict_behaves_like_easydict():
    d = AttrDict({"a": {"b": 1}, "c": 2})
    assert d.a.b == 1 and d.get("missing", 5) == 5 and d["c"] == 2
    d.e = {"f": 3}
    assert d.e.f == 3
    with pytest.raises(AttributeError):
        _ = d.nope


def test_geglu_pack_index_is_a_permutation_pairing_value_and_gate():
    idx = geglu_pack_index(2048, "cpu")
    assert sorted(idx.tolist()) == list(range(4096))
    grp = idx.view(-1, 32)
    assert torch.equal(grp[:, :16] + 2048, grp[:, 16:])
    assert torch.equal(grp[:, :16].reshape(-1), torch.arange(2048))


def test_karras_schedule_matches_oracle():
    a = models_radar_generation.karras_schedule(18, 0.002, 80, 7)
    assert torch.equal(a, orc.karras_sigmas())


def test_stacked_generator_depends_only_on_seed():
    g1 = models_radar_generation.StackedRandomGenerator("cpu", [5, 9]).randn([2, 4, 3])
    g2 = models_radar_generation.StackedRandomGenerator("cpu", [9]).randn([1, 4, 3])
    assert torch.equal(g1[1], g2[0])
    assert torch.equal(g1, orc.stacked_randn([5, 9], [4, 3]))


def test_inverse_norm_constants():
    c = postproc.inverse_norm_constants([0, -90, -20, 15.8, 90, 20])
    assert c.dtype == np.float32
    assert np.allclose(c, [7.9, 90, 20, 7.9, 0, 0])
    iso = postproc.inverse_norm_constants([0, -90, -20, 15.8, 90, 20], True, True)
    assert np.allclose(iso[:3], 90)


def test_shard_frames_partitions_contiguously():
    for total in (1, 7, 64, 256):
        for world in (1, 2, 3, 8):
            spans = [gather.shard_frames(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gather_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    frames, cap = (2, 5) if rank == 0 else (1, 3)      # ragged: different frame counts and capacities
    pts = torch.full((frames, cap, 3), float(rank + 1))
    cnt = torch.tensor([4, 0] if rank == 0 else [3], dtype=torch.int32)
    pts[:, :, 1] = torch.arange(cap, dtype=torch.float32)[None, :]      # y = index of the point inside its frame
    g = gather.gather_point_clouds(pts, cnt)
    lat = gather.gather_latents(torch.full((frames, 4, 2), float(rank)))
    torch.save((g.points, g.counts, g.offsets, lat), os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_gather_point_clouds_gloo_world2(tmp_path):
    port = _free_port()
    mp.spawn(_gather_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(tmp_path / "r0.pt")
    r1 = torch.load(tmp_path / "r1.pt")
    for a, b in zip(r0, r1):
        assert torch.equal(a, b)               # every rank sees the whole batch
    pts, cnt, off, lat = r0
    g = gather.GatheredClouds(pts, cnt, off)
    assert cnt.tolist() == [4, 0, 3] and off.tolist() == [0, 4, 4]       # compact: rank totals 4 and 3, padded to 4
    assert pts.shape == (2 * 4, 3)
    f0, f1, f2 = g.frame(0), g.frame(1), g.frame(2)
    assert f0.shape == (4, 3) and f1.shape == (0, 3) and f2.shape == (3, 3)
    assert f0[:, 0].tolist() == [1.0] * 4 and f0[:, 1].tolist() == [0.0, 1.0, 2.0, 3.0]
    assert f2[:, 0].tolist() == [2.0] * 3 and f2[:, 1].tolist() == [0.0, 1.0, 2.0]
    assert lat.shape == (3, 4, 2) and lat[:, 0, 0].tolist() == [0.0, 0.0, 1.0]
