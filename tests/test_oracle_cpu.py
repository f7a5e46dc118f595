"""The oracle (oracle/torch_oracle.py) against the golden vectors the REFERENCE ITSELF produced
(oracle/pin_against_reference.py ran /root/reference/src verbatim and wrote tests/golden/)."""
import json

import torch

from oracle import torch_oracle as O


def test_context_windows_bit_exact(golden_dir):
    tables = json.loads((golden_dir / "integer_tables.json").read_text())
    for F, want in tables["windows"].items():
        assert O.uniform_windows(0, int(F), 24, 1, 4) == want
    assert len(tables["windows"]["48"]) == 3 and len(tables["windows"]["64"]) == 4 and len(tables["windows"]["150"]) == 8


def test_ddim_tables_bit_exact(golden_dir):
    tables = json.loads((golden_dir / "integer_tables.json").read_text())
    for N, want in tables["timesteps"].items():
        d = O.DDIM()
        assert [int(v) for v in d.set_timesteps(int(N))] == want
    assert tables["timesteps"]["20"][:3] == [999, 949, 899] and tables["timesteps"]["30"][:4] == [999, 966, 932, 899]
    d = O.DDIM()
    assert float(d.alphas_cumprod[999]) == 0.0  # zero terminal SNR: first step is x0 = -v
    for i, v in tables["alphas_cumprod_probe"].items():
        assert abs(float(d.alphas_cumprod[int(i)]) - v) < 1e-7
    d.set_timesteps(30)
    # prev_t = t - 1000 // N, not the next table entry (966 -> 933 although the table continues with 932)
    assert d.coefficients(966)[2] == float(d.alphas_cumprod[933] ** 0.5)


def test_first_ddim_step_is_minus_v():
    d = O.DDIM()
    d.set_timesteps(20)
    x, v = torch.randn(1, 4, 2, 8, 8), torch.randn(1, 4, 2, 8, 8)
    a_p = d.alphas_cumprod[949]
    want = a_p ** 0.5 * (-v) + (1 - a_p) ** 0.5 * x
    assert torch.allclose(d.step(v, 999, x), want, atol=1e-6)


def test_small_unet_matches_reference_golden(golden_dir):
    g = torch.load(golden_dir / "unet_small_read.pt")
    cfg = O.UNetConfig(block_out_channels=tuple(g["cfg"]))
    seed, f, hw = g["seed"], g["f"], g["hw"]
    sd_den = O.make_denoising_unet_sd(cfg, seed=seed)
    sd_ref = O.make_reference_unet_sd(cfg, seed=seed + 1)
    sd_pg = O.make_pose_guider_sd(seed=seed + 2, out_channels=cfg.block_out_channels[0])
    gen = torch.Generator().manual_seed(seed + 10)
    ref_lat = torch.randn(1, 4, hw, hw, generator=gen)
    emb = torch.randn(1, 1, cfg.cross_attention_dim, generator=gen)
    ehs = torch.cat([torch.zeros_like(emb), emb])
    x = torch.randn(1, 8, f, hw, hw, generator=gen).repeat(2, 1, 1, 1, 1)
    pose_img = torch.rand(1, 3, f, hw * 8, hw * 8, generator=gen)
    with torch.no_grad():
        banks = O.reference_unet_banks(sd_ref, ref_lat.repeat(2, 1, 1, 1), ehs, cfg)
        pose = O.pose_guider(sd_pg, pose_img)
        out = O.denoising_unet(sd_den, x, 499, ehs, pose.repeat(2, 1, 1, 1, 1), banks, cfg, cfg=True)
    want = g["out"].float()
    err = float((out - want).norm() / want.norm())
    assert err < 1e-3, err  # the fixture is stored in fp16


def test_bank_pairing_order():
    paths = O.transformer_paths(O.UNetConfig())
    assert paths[:6] == ["down_blocks.2.attentions.0", "down_blocks.2.attentions.1", "up_blocks.1.attentions.0",
                         "up_blocks.1.attentions.1", "up_blocks.1.attentions.2", "mid_block.attentions.0"]
    assert len(paths) == 16
