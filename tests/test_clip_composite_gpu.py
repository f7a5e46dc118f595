"""SURVEY.md §8f rows 3 and 4 on the GPU: the CLIP image encoder on the engine's kernels against the transformers module
it replaces (fp32 reference, same weights), and the run_edit.py blend chain against its numpy restatement (bit-exact)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_clip_vit_l14_engine_matches_transformers():
    from transformers import CLIPVisionConfig, CLIPVisionModelWithProjection

    from mimo_b200.clip_engine import CLIPVisionEngine
    torch.manual_seed(11)
    cfg = CLIPVisionConfig(hidden_size=1024, intermediate_size=4096, num_hidden_layers=24, num_attention_heads=16,
                           image_size=224, patch_size=14, projection_dim=768)
    m = CLIPVisionModelWithProjection(cfg).eval().cuda()
    sd16 = {k: v.half() for k, v in m.state_dict().items()}
    m.load_state_dict({k: v.float() for k, v in sd16.items()})  # fp32 math on the fp16-rounded weights
    px = torch.randn(2, 3, 224, 224, device="cuda").half()
    torch.backends.cuda.matmul.allow_tf32 = False
    with torch.no_grad():
        want = m(px.float()).image_embeds
        t16 = m.half()(px).image_embeds.float()
    got = CLIPVisionEngine(sd16, cfg, "cuda", torch.float16).image_embeds(px).float()
    rel = lambda a, b: float((a - b).norm() / b.norm())
    e_eng, e_ref = rel(got, want), rel(t16, want)
    print(f"CLIP ViT-L/14: engine {e_eng:.3e}  transformers-fp16 {e_ref:.3e} (vs fp32)")
    assert e_eng <= max(2e-3, e_ref), (e_eng, e_ref)


@pytest.mark.parametrize("occ,prev", [(False, False), (True, False), (False, True), (True, True)])
def test_composite_frame_is_bit_exact(occ, prev):
    from mimo_b200 import ops
    from oracle.composite_oracle import composite_frame
    rng = np.random.RandomState(5 + 2 * occ + prev)
    H, W = 270, 481
    canvas = rng.randint(0, 256, (H, W, 3)).astype(np.uint8)
    bk = rng.randint(0, 256, (H, W, 3)).astype(np.uint8)
    mask = rng.rand(H, W).astype(np.float32)
    mask[:40] = 0.0
    mask[-40:] = 1.0
    o = rng.randint(0, 256, (H, W)).astype(np.uint8) if occ else None
    v = rng.randint(0, 256, (H, W, 3)).astype(np.uint8) if occ else None
    p = rng.randint(0, 256, (H, W, 3)).astype(np.uint8) if prev else None
    for factor in (1 / 5, 3 / 5, 1.0):
        want = composite_frame(canvas, bk, mask, o, v, p, factor)
        d = lambda a: None if a is None else torch.from_numpy(a).cuda()
        got = ops.composite_frame(d(canvas), d(bk), d(mask), occ=d(o), vid=d(v), prev=d(p), factor=factor).cpu().numpy()
        assert np.array_equal(got, want), (occ, prev, factor, int((got != want).sum()))


def test_packed_weight_cache_reproduces_the_engine(tmp_path, monkeypatch):
    """MIMO_B200_WEIGHT_CACHE: the second construction maps the packed tensors from disk and must run bit-identically."""
    from mimo_b200 import engine as E
    from oracle import torch_oracle as O
    monkeypatch.setenv("MIMO_B200_WEIGHT_CACHE", str(tmp_path))
    widths, f, hw, seed = (128, 256, 512, 512), 2, 16, 31
    cfg = O.UNetConfig(block_out_channels=widths)
    sd_den, sd_ref = O.make_denoising_unet_sd(cfg, seed), O.make_reference_unet_sd(cfg, seed + 1)
    g = torch.Generator().manual_seed(seed)
    ref_lat = torch.randn(1, 4, hw, hw, generator=g).repeat(2, 1, 1, 1).half().cuda()
    emb = torch.randn(1, 1, 768, generator=g)
    ehs = torch.cat([torch.zeros_like(emb), emb]).half().cuda()
    x = torch.randn(1, 8, f, hw, hw, generator=g).repeat(2, 1, 1, 1, 1).half().cuda()
    outs, cached = [], []
    for _ in range(2):
        den = E.UNetEngine(sd_den, E.UNetSpec(block_out_channels=widths), "cuda")
        ref = E.UNetEngine(sd_ref, E.UNetSpec(block_out_channels=widths, in_channels=4, motion=False, out_head=False), "cuda")
        cached.append((den.from_cache, ref.from_cache))
        den.begin_clip(ehs, ref.write_banks(ref_lat, ehs, den), cfg=True, frames=f)
        outs.append(den.forward(x, 499, None).clone())
    assert cached == [(False, False), (True, True)]
    assert len(list(tmp_path.glob("unet-*.safetensors"))) == 2
    assert torch.equal(outs[0], outs[1])
